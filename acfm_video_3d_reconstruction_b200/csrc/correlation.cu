// correlation.cu — FlowNet-style correlation cost volume.
//
// Replaces the reference's only native code, the `correlation_cuda` extension of its frozen optical-flow network
// (/root/reference/multiframe/data/optical_flow/model/correlation_package/correlation_cuda_kernel.cu:46-147 forward,
// :150-334 backward; host side correlation_cuda.cc:9-104; module correlation.py:54-74; instantiated by
// MaskFlownet.py:116,416 as Correlation(pad_size=4, kernel_size=1, max_displacement=4, stride1=1, stride2=1)).
// The reference builds it for sm_50..sm_75 (+ compute_70 PTX) and runs three kernels per call: two NCHW -> padded NHWC
// copies and a one-warp-per-output-pixel dot product.
//
//   out[n, (tj+dr)*D + (ti+dr), oy, ox] = 1/(k*k*C) * sum_{j,i in [-kr,kr]} sum_c  P1[n, c, y1+j, x1+i] * P2[n, c, y1+j+tj*s2, x1+i+ti*s2]
//   P = input zero-padded by `pad`, (y1,x1) = (oy,ox)*s1 + md, dr = md/s2, D = 2 dr + 1, kr = (k-1)/2,
//   outH = ceil((H + 2 pad - 2 (kr + md)) / s1)          (correlation_cuda.cc:24-32)
//
// Here: no padded copies (bounds are tested while the tiles are staged), and for the network's configuration
// (k = 1, s1 = s2 = 1, D = 9) a register-tiled kernel: a CTA owns a 32 x 8 tile of output pixels, stages 8 channels of the
// in1 tile and of the in2 halo tile (16 x 40) in shared memory at a time (3-stage cp.async pipeline), and each thread accumulates 4 pixels x 3 rows of
// displacements x 9 columns of displacements = 108 dot products in registers — 10 LDS.128 per 108 FMA, against one shared /
// global load per FMA in the reference kernel.  Everything else goes through a plain one-thread-per-output kernel.
// fp32 accumulation like the reference (`float acc0`), divided by k*k*C at the end.  The backward is the exact adjoint
// (equal to the reference's for stride1 = 1; the reference only ever runs this op under no_grad, multiframe/main.py:386-411).
#include "common.cuh"

namespace {

struct CorrParams {
  const float* in1;
  const float* in2;
  float* out;
  int B, C, H, W, pad, k, md, s1, s2, kr, dr, D, outH, outW;
  float nelems, inv_nelems;
};

// ---- generic forward: one thread per output element ---------------------------------------------------
__global__ void __launch_bounds__(256) corr_fwd_generic_kernel(const CorrParams p) {
  const long long total = (long long)p.B * p.D * p.D * p.outH * p.outW;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int ox = (int)(idx % p.outW);
  const int oy = (int)((idx / p.outW) % p.outH);
  const int tc = (int)((idx / ((long long)p.outW * p.outH)) % (p.D * p.D));
  const int n = (int)(idx / ((long long)p.outW * p.outH * p.D * p.D));
  const int tj = tc / p.D - p.dr, ti = tc % p.D - p.dr;
  const int y1 = oy * p.s1 + p.md - p.pad, x1 = ox * p.s1 + p.md - p.pad;  // unpadded coordinates
  const int y2 = y1 + tj * p.s2, x2 = x1 + ti * p.s2;
  const size_t plane = (size_t)p.H * p.W;
  const float* a = p.in1 + (size_t)n * p.C * plane;
  const float* b = p.in2 + (size_t)n * p.C * plane;
  float acc = 0.0f;
  for (int j = -p.kr; j <= p.kr; ++j) {
    for (int i = -p.kr; i <= p.kr; ++i) {
      const int ya = y1 + j, xa = x1 + i, yb = y2 + j, xb = x2 + i;
      if (ya < 0 || ya >= p.H || xa < 0 || xa >= p.W || yb < 0 || yb >= p.H || xb < 0 || xb >= p.W) continue;  // zero padding
      const float* pa = a + (size_t)ya * p.W + xa;
      const float* pb = b + (size_t)yb * p.W + xb;
      for (int c = 0; c < p.C; ++c) acc = fmaf(pa[c * plane], pb[c * plane], acc);
    }
  }
  p.out[idx] = acc / p.nelems;
}

// 4-byte asynchronous global -> shared copy; !valid: the destination is zero-filled and the source is not read
__device__ __forceinline__ void cp_async_f32(float* dst_smem, const float* src, bool valid) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(smem_u32(dst_smem)), "l"(src), "r"(valid ? 4 : 0) : "memory");
}

__device__ __forceinline__ void cp_async_f32x4(float* dst_smem, const float* src, bool valid) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(dst_smem)), "l"(src), "r"(valid ? 16 : 0) : "memory");
}

// ---- the network's configuration: k = 1, s1 = s2 = 1, dr = 4 (81 displacements) ----------------------
constexpr int kTX = 32, kTY = 8, kCC = 8, kDR = 4, kD = 9;
constexpr int kHaloW = kTX + 2 * kDR, kHaloH = kTY + 2 * kDR;  // 40 x 16
constexpr int kCorrThreads = (kTX / 4) * kTY * 3;              // 4 pixels x 3 displacement rows per thread: 192

constexpr int kStages = 3;                                       // cp.async pipeline depth
constexpr int kS1Floats = kCC * kTY * kTX, kS2Floats = kCC * kHaloH * kHaloW;
constexpr int kCorrSmemBytes = kStages * (kS1Floats + kS2Floats) * 4;  // 84 KB: dynamic shared memory, 2 CTAs per SM

template <bool VEC>
__global__ void __launch_bounds__(kCorrThreads, 2) corr_fwd_d9_kernel(const CorrParams p) {
  extern __shared__ __align__(16) float corr_smem[];
  const int tid = threadIdx.x;
  const int gx = tid & 7, gy = (tid >> 3) & 7, tg = tid >> 6;  // pixel group (4 wide), row, displacement-row group
  const int n = blockIdx.z;
  const int ox0 = blockIdx.x * kTX, oy0 = blockIdx.y * kTY;
  const int off = p.md - p.pad;  // output pixel (oy, ox) sits at unpadded (oy + off, ox + off)
  const size_t plane = (size_t)p.H * p.W;
  const float* a = p.in1 + (size_t)n * p.C * plane;
  const float* b = p.in2 + (size_t)n * p.C * plane;

  float acc[3][4][kD];
#pragma unroll
  for (int t = 0; t < 3; ++t)
#pragma unroll
    for (int q = 0; q < 4; ++q)
#pragma unroll
      for (int i = 0; i < kD; ++i) acc[t][q][i] = 0.0f;

  // stage kCC channels starting at c0 into buffer `st`: the in1 tile and the in2 halo tile, zero outside the image (the
  // reference's zero padding).  The copies are asynchronous (cp.async, zero-filling form for the padding): all of a chunk's
  // loads are in flight at once, no register carries them past the 108 accumulators, and the next chunk's loads overlap this
  // chunk's arithmetic.  Unaligned shapes: 4-byte copies, one warp per (channel, row).
  auto stage = [&](int c0, int st) {
    float* s1 = corr_smem + st * (kS1Floats + kS2Floats);   // [kCC][kTY][kTX]
    float* s2 = s1 + kS1Floats;                              // [kCC][kHaloH][kHaloW]
    if (VEC) {
      // rows start 16-byte aligned (W, the tile origin and md - pad are multiples of 4): 16-byte copies, each wholly inside
      // or wholly outside the image
      for (int e = tid; e < kCC * kTY * (kTX / 4); e += kCorrThreads) {
        const int xv = e % (kTX / 4), y = (e / (kTX / 4)) % kTY, c = e / ((kTX / 4) * kTY);
        const int yy = oy0 + y + off, xx = ox0 + xv * 4 + off;
        const bool ok = (c0 + c < p.C) && yy >= 0 && yy < p.H && xx >= 0 && xx < p.W;
        cp_async_f32x4(s1 + (c * kTY + y) * kTX + xv * 4, a + (ok ? (size_t)(c0 + c) * plane + (size_t)yy * p.W + xx : 0), ok);
      }
      for (int e = tid; e < kCC * kHaloH * (kHaloW / 4); e += kCorrThreads) {
        const int xv = e % (kHaloW / 4), y = (e / (kHaloW / 4)) % kHaloH, c = e / ((kHaloW / 4) * kHaloH);
        const int yy = oy0 + y + off - kDR, xx = ox0 + xv * 4 + off - kDR;
        const bool ok = (c0 + c < p.C) && yy >= 0 && yy < p.H && xx >= 0 && xx < p.W;
        cp_async_f32x4(s2 + (c * kHaloH + y) * kHaloW + xv * 4, b + (ok ? (size_t)(c0 + c) * plane + (size_t)yy * p.W + xx : 0), ok);
      }
    } else {
      const int lane = tid & 31, wrp = tid >> 5;
      constexpr int kWarps = kCorrThreads / 32;
      for (int r = wrp; r < kCC * kTY; r += kWarps) {
        const int c = r / kTY, y = r - c * kTY;
        const int yy = oy0 + y + off, xx = ox0 + lane + off;
        const bool rowok = (c0 + c < p.C) && yy >= 0 && yy < p.H;
        const float* src = a + (rowok ? (size_t)(c0 + c) * plane + (size_t)yy * p.W : 0);
        const bool ok = rowok && xx >= 0 && xx < p.W;
        cp_async_f32(s1 + (c * kTY + y) * kTX + lane, src + (ok ? xx : 0), ok);
      }
      for (int r = wrp; r < kCC * kHaloH; r += kWarps) {
        const int c = r / kHaloH, y = r - c * kHaloH;
        const int yy = oy0 + y + off - kDR, xx = ox0 + lane + off - kDR;
        const bool rowok = (c0 + c < p.C) && yy >= 0 && yy < p.H;
        const float* src = b + (rowok ? (size_t)(c0 + c) * plane + (size_t)yy * p.W : 0);
        const bool ok0 = rowok && xx >= 0 && xx < p.W;
        cp_async_f32(s2 + (c * kHaloH + y) * kHaloW + lane, src + (ok0 ? xx : 0), ok0);
        if (lane < kHaloW - 32) {
          const bool ok1 = rowok && xx + 32 >= 0 && xx + 32 < p.W;
          cp_async_f32(s2 + (c * kHaloH + y) * kHaloW + 32 + lane, src + (ok1 ? xx + 32 : 0), ok1);
        }
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };

  // kStages - 1 chunks in flight ahead of the arithmetic; one commit group per iteration (empty past the last chunk), so
  // "all but the newest kStages - 1 groups have landed" always means "this iteration's chunk has landed"
#pragma unroll
  for (int s0 = 0; s0 < kStages - 1; ++s0) {
    if (s0 * kCC < p.C) stage(s0 * kCC, s0);
    else asm volatile("cp.async.commit_group;" ::: "memory");
  }
  int buf = 0, nbuf = kStages - 1;
  for (int c0 = 0; c0 < p.C; c0 += kCC) {
    if (c0 + (kStages - 1) * kCC < p.C) stage(c0 + (kStages - 1) * kCC, nbuf);
    else asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group %0;" ::"n"(kStages - 1) : "memory");
    __syncthreads();
    const float* s1 = corr_smem + buf * (kS1Floats + kS2Floats);
    const float* s2 = s1 + kS1Floats;
    // threads whose four pixels lie outside the output (small pyramid levels fill a fraction of the tile) only help staging
    const int nch = (oy0 + gy < p.outH && ox0 + gx * 4 < p.outW) ? kCC : 0;
#pragma unroll 2
    for (int c = 0; c < nch; ++c) {
      const float4 av = *reinterpret_cast<const float4*>(s1 + (c * kTY + gy) * kTX + gx * 4);
      const float aa[4] = {av.x, av.y, av.z, av.w};
#pragma unroll
      for (int t = 0; t < 3; ++t) {
        const float* row = s2 + (c * kHaloH + gy + tg * 3 + t) * kHaloW + gx * 4;
        const float4 v0 = *reinterpret_cast<const float4*>(row), v1 = *reinterpret_cast<const float4*>(row + 4),
                     v2 = *reinterpret_cast<const float4*>(row + 8);
        const float v[12] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w, v2.x, v2.y, v2.z, v2.w};
#pragma unroll
        for (int q = 0; q < 4; ++q)
#pragma unroll
          for (int i = 0; i < kD; ++i) acc[t][q][i] = fmaf(aa[q], v[q + i], acc[t][q][i]);
      }
    }
    __syncthreads();  // everyone is done with `buf` before a later iteration's copies overwrite it
    buf = (buf + 1 == kStages) ? 0 : buf + 1;
    nbuf = (nbuf + 1 == kStages) ? 0 : nbuf + 1;
  }

  const int oy = oy0 + gy, ox = ox0 + gx * 4;
  if (oy >= p.outH || ox >= p.outW) return;
  const size_t oplane = (size_t)p.outH * p.outW;
  float* o = p.out + (size_t)n * kD * kD * oplane + (size_t)oy * p.outW + ox;
  const bool vec = (ox + 3 < p.outW) && ((p.outW & 3) == 0) && ((((uintptr_t)p.out) & 15u) == 0);
#pragma unroll
  for (int t = 0; t < 3; ++t) {
#pragma unroll
    for (int i = 0; i < kD; ++i) {
      float* d = o + (size_t)((tg * 3 + t) * kD + i) * oplane;
      const float r0 = acc[t][0][i] * p.inv_nelems, r1 = acc[t][1][i] * p.inv_nelems, r2 = acc[t][2][i] * p.inv_nelems, r3 = acc[t][3][i] * p.inv_nelems;
      if (vec) {
        *reinterpret_cast<float4*>(d) = make_float4(r0, r1, r2, r3);
      } else {
        d[0] = r0;
        if (ox + 1 < p.outW) d[1] = r1;
        if (ox + 2 < p.outW) d[2] = r2;
        if (ox + 3 < p.outW) d[3] = r3;
      }
    }
  }
}

// ---- backward: exact adjoint, one thread per input element ---------------------------------------------
// grad_in1[n,c,y,x] = 1/nelems * sum over (tc, j, i) with (oy,ox)*s1 + md + (j,i) = (y,x) + pad of
//                     grad_out[n,tc,oy,ox] * P2[n,c, y + tj*s2, x + ti*s2]
// grad_in2[n,c,y,x] = 1/nelems * sum over (tc, j, i) with (oy,ox)*s1 + md + (j,i) + (tj,ti)*s2 = (y,x) + pad of
//                     grad_out[n,tc,oy,ox] * P1[n,c, y - tj*s2, x - ti*s2]
template <bool SECOND>
__global__ void __launch_bounds__(256) corr_bwd_kernel(const CorrParams p, const float* __restrict__ gout, float* __restrict__ gin) {
  const long long total = (long long)p.B * p.C * p.H * p.W;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int x = (int)(idx % p.W), y = (int)((idx / p.W) % p.H);
  const int c = (int)((idx / ((long long)p.W * p.H)) % p.C);
  const int n = (int)(idx / ((long long)p.W * p.H * p.C));
  const size_t plane = (size_t)p.H * p.W, oplane = (size_t)p.outH * p.outW;
  const float* other = (SECOND ? p.in1 : p.in2) + ((size_t)n * p.C + c) * plane;
  const float* go = gout + (size_t)n * p.D * p.D * oplane;
  float acc = 0.0f;
  for (int tc = 0; tc < p.D * p.D; ++tc) {
    const int dj = (tc / p.D - p.dr) * p.s2, di = (tc % p.D - p.dr) * p.s2;
    // the partner sample of the other input
    const int yo = SECOND ? y - dj : y + dj, xo = SECOND ? x - di : x + di;
    if (yo < 0 || yo >= p.H || xo < 0 || xo >= p.W) continue;  // zero padding
    const float w = other[(size_t)yo * p.W + xo];
    // (y1, x1) + (j, i) in padded coordinates of input 1
    const int Y = (SECOND ? y - dj : y) + p.pad, X = (SECOND ? x - di : x) + p.pad;
    for (int j = -p.kr; j <= p.kr; ++j) {
      const int ny = Y - j - p.md;
      if (ny < 0 || ny % p.s1 != 0) continue;
      const int oy = ny / p.s1;
      if (oy >= p.outH) continue;
      for (int i = -p.kr; i <= p.kr; ++i) {
        const int nx = X - i - p.md;
        if (nx < 0 || nx % p.s1 != 0) continue;
        const int ox = nx / p.s1;
        if (ox >= p.outW) continue;
        acc = fmaf(go[(size_t)tc * oplane + (size_t)oy * p.outW + ox], w, acc);
      }
    }
  }
  gin[idx] = acc / p.nelems;
}

int corr_setup(CorrParams& p, const char* who, int B, int C, int H, int W, int pad, int k, int md, int s1, int s2) {
  ACFM_REQUIRE(B >= 0 && C > 0 && H > 0 && W > 0, ACFM_ERR_BAD_ARG, "%s: bad sizes B=%d C=%d H=%d W=%d", who, B, C, H, W);
  ACFM_REQUIRE(pad >= 0 && k >= 1 && (k & 1) == 1 && md >= 0 && s1 >= 1 && s2 >= 1, ACFM_ERR_BAD_ARG,
               "%s: pad_size >= 0, odd kernel_size >= 1, max_displacement >= 0, strides >= 1 expected", who);
  p.B = B; p.C = C; p.H = H; p.W = W; p.pad = pad; p.k = k; p.md = md; p.s1 = s1; p.s2 = s2;
  p.kr = (k - 1) / 2; p.dr = md / s2; p.D = 2 * p.dr + 1;
  const int span_h = H + 2 * pad - 2 * (p.kr + md), span_w = W + 2 * pad - 2 * (p.kr + md);
  ACFM_REQUIRE(span_h > 0 && span_w > 0, ACFM_ERR_BAD_ARG, "%s: the padded input is smaller than the correlation border", who);
  ACFM_REQUIRE(p.kr <= md, ACFM_ERR_UNSUPPORTED, "%s: kernel radius %d > max_displacement %d reads outside the padded input in the reference too", who, p.kr, md);
  p.outH = (span_h + s1 - 1) / s1; p.outW = (span_w + s1 - 1) / s1;
  p.nelems = (float)(k * k * C);
  p.inv_nelems = (float)(1.0 / (double)(k * k * C));  // the tiled kernel multiplies (exact for power-of-two C, else within 1 ulp of the division)
  return ACFM_OK;
}

}  // namespace

extern "C" int acfm_correlation_out_shape(int H, int W, int pad_size, int kernel_size, int max_displacement, int stride1, int stride2,
                                          int* channels, int* out_h, int* out_w) {
  CorrParams p;
  const int rc = corr_setup(p, "acfm_correlation_out_shape", 1, 1, H, W, pad_size, kernel_size, max_displacement, stride1, stride2);
  if (rc != ACFM_OK) return rc;
  if (channels) *channels = p.D * p.D;
  if (out_h) *out_h = p.outH;
  if (out_w) *out_w = p.outW;
  return ACFM_OK;
}

extern "C" int acfm_correlation_fwd(const float* input1, const float* input2, int B, int C, int H, int W, int pad_size, int kernel_size,
                                    int max_displacement, int stride1, int stride2, float* output, void* stream) {
  CorrParams p;
  const int rc = corr_setup(p, "acfm_correlation_fwd", B, C, H, W, pad_size, kernel_size, max_displacement, stride1, stride2);
  if (rc != ACFM_OK) return rc;
  if (B == 0) return ACFM_OK;
  ACFM_REQUIRE(input1 && input2 && output, ACFM_ERR_BAD_ARG, "acfm_correlation_fwd: null pointer");
  p.in1 = input1; p.in2 = input2; p.out = output;
  cudaStream_t st = (cudaStream_t)stream;
#ifdef ACFM_TUNING
  static const bool force_generic = getenv("ACFM_CORR_GENERIC") != nullptr;  // tuning builds only (scripts/build_variant.sh)
#else
  const bool force_generic = false;
#endif
  if (kernel_size == 1 && stride1 == 1 && stride2 == 1 && p.dr == kDR && B <= 65535 && !force_generic) {
    const dim3 grid((p.outW + kTX - 1) / kTX, (p.outH + kTY - 1) / kTY, B);
    ACFM_REQUIRE(grid.y <= 65535, ACFM_ERR_UNSUPPORTED, "acfm_correlation_fwd: output too tall");
    const bool vec = (W & 3) == 0 && ((p.md - p.pad) & 3) == 0 && ((((uintptr_t)input1) | ((uintptr_t)input2)) & 15u) == 0;
    static std::atomic<int> smem_v[kAcfmMaxDevices], smem_s[kAcfmMaxDevices];
    if (vec) {
      ACFM_CUDA_OK(acfm_ensure_smem(corr_fwd_d9_kernel<true>, kCorrSmemBytes, smem_v));
      corr_fwd_d9_kernel<true><<<grid, kCorrThreads, kCorrSmemBytes, st>>>(p);
    } else {
      ACFM_CUDA_OK(acfm_ensure_smem(corr_fwd_d9_kernel<false>, kCorrSmemBytes, smem_s));
      corr_fwd_d9_kernel<false><<<grid, kCorrThreads, kCorrSmemBytes, st>>>(p);
    }
    ACFM_LAUNCH_OK("corr_fwd_d9_kernel");
  } else {
    const long long total = (long long)B * p.D * p.D * p.outH * p.outW;
    ACFM_REQUIRE((total + 255) / 256 < (1ll << 31), ACFM_ERR_UNSUPPORTED, "acfm_correlation_fwd: output too large");
    corr_fwd_generic_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(p);
    ACFM_LAUNCH_OK("corr_fwd_generic_kernel");
  }
  return ACFM_OK;
}

extern "C" int acfm_correlation_bwd(const float* input1, const float* input2, const float* grad_output, int B, int C, int H, int W,
                                    int pad_size, int kernel_size, int max_displacement, int stride1, int stride2, float* grad_input1,
                                    float* grad_input2, void* stream) {
  CorrParams p;
  const int rc = corr_setup(p, "acfm_correlation_bwd", B, C, H, W, pad_size, kernel_size, max_displacement, stride1, stride2);
  if (rc != ACFM_OK) return rc;
  if (B == 0) return ACFM_OK;
  ACFM_REQUIRE(input1 && input2 && grad_output && (grad_input1 || grad_input2), ACFM_ERR_BAD_ARG, "acfm_correlation_bwd: null pointer");
  p.in1 = input1; p.in2 = input2; p.out = nullptr;
  cudaStream_t st = (cudaStream_t)stream;
  const long long total = (long long)B * C * H * W;
  ACFM_REQUIRE((total + 255) / 256 < (1ll << 31), ACFM_ERR_UNSUPPORTED, "acfm_correlation_bwd: input too large");
  const unsigned blocks = (unsigned)((total + 255) / 256);
  if (grad_input1) corr_bwd_kernel<false><<<blocks, 256, 0, st>>>(p, grad_output, grad_input1);
  if (grad_input2) corr_bwd_kernel<true><<<blocks, 256, 0, st>>>(p, grad_output, grad_input2);
  ACFM_LAUNCH_OK("corr_bwd_kernel");
  return ACFM_OK;
}
