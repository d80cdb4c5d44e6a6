// skin.cu — handle deformation ("LBS") fused with the camera-multiplex projection.
//
// The reference solves, per frame, (L^T L + A^T A) X = L^T L m + A^T (A m + D)   (B*T identical 642^2
// Cholesky factorizations: /root/reference/multiframe/main.py:586-609, monocular/main.py:203-218),
// which is algebraically X = m + W D with W = (L^T L + A^T A)^-1 A^T  (V x K_h).
// W is computed once per step on the host side (one V x V solve with K_h right-hand sides); this file
// does the per-frame part as vertex-major passes:
//   pred_v[b]       = mean_v + W * delta[b]                          (b < NB = B*T frames)
//   ndc[g*NB + b]   = view(project(pred_v[b], cams[g*NB + b]))       (g < G hypotheses)
// and the matching backward.  HBM-bound: 12 B written per (render, vertex); W (V*K_h*4 B) and delta
// stay in L1/L2.
#include "common.cuh"

namespace {

constexpr int kThreads = 128;
constexpr int kMaxHandles = 128;

__device__ __forceinline__ void hamilton_s(const float* a, const float* b, float* o) {
  o[0] = fsub(fsub(fsub(fmul(a[0], b[0]), fmul(a[1], b[1])), fmul(a[2], b[2])), fmul(a[3], b[3]));
  o[1] = fsub(fadd(fadd(fmul(a[0], b[1]), fmul(a[1], b[0])), fmul(a[2], b[3])), fmul(a[3], b[2]));
  o[2] = fadd(fadd(fsub(fmul(a[0], b[2]), fmul(a[1], b[3])), fmul(a[2], b[0])), fmul(a[3], b[1]));
  o[3] = fadd(fsub(fadd(fmul(a[0], b[3]), fmul(a[1], b[2])), fmul(a[2], b[1])), fmul(a[3], b[0]));
}

// grid (ceil(V/kThreads), NB)
__global__ void __launch_bounds__(kThreads) skin_project_fwd_kernel(
    const float* __restrict__ mean_v, const float* __restrict__ Wm, const float* __restrict__ delta,
    const float* __restrict__ cams, int NB, int G, int V, int Kh, float offset_z, float sx, float sy, float z_add,
    float* __restrict__ pred_v, float* __restrict__ ndc) {
  __shared__ float sd[kMaxHandles * 3];
  const int b = blockIdx.y;
  const int v = blockIdx.x * kThreads + threadIdx.x;
  for (int i = threadIdx.x; i < Kh * 3; i += kThreads) sd[i] = delta[(size_t)b * Kh * 3 + i];
  __syncthreads();
  if (v >= V) return;
  float x0 = mean_v[v * 3], x1 = mean_v[v * 3 + 1], x2 = mean_v[v * 3 + 2];
  const float* w = Wm + (size_t)v * Kh;
  for (int k = 0; k < Kh; ++k) {
    const float wk = w[k];
    x0 = fmaf(wk, sd[k * 3], x0);
    x1 = fmaf(wk, sd[k * 3 + 1], x1);
    x2 = fmaf(wk, sd[k * 3 + 2], x2);
  }
  if (pred_v) {
    float* o = pred_v + ((size_t)b * V + v) * 3;
    o[0] = x0; o[1] = x1; o[2] = x2;
  }
  if (!ndc) return;
  for (int g = 0; g < G; ++g) {
    const int n = g * NB + b;
    const float* c = cams + (size_t)n * 7;  // broadcast load, L1-resident
    const float q[4] = {c[3], c[4], c[5], c[6]};
    const float qc[4] = {q[0], fmul(-1.0f, q[1]), fmul(-1.0f, q[2]), fmul(-1.0f, q[3])};
    const float xq[4] = {fmul(x0, 0.0f), x0, x1, x2};
    float t[4], r[4];
    hamilton_s(xq, qc, t);
    hamilton_s(q, t, r);
    const float px = fadd(fmul(c[0], r[1]), c[1]);
    const float py = fadd(fmul(c[0], r[2]), c[2]);
    const float pz = fadd(fmul(c[0], r[3]), offset_z);
    float* o = ndc + ((size_t)n * V + v) * 3;
    o[0] = fmul(sx, px);
    o[1] = fmul(sy, py);
    o[2] = (z_add != 0.0f) ? fadd(pz, z_add) : pz;
  }
}

// grad_delta[b,k,c] = sum_v W[v,k] * gp[b,v,c].  grid (NB); block (32*ceil(Kh/32), S slices)
__global__ void skin_bwd_delta_kernel(const float* __restrict__ Wm, const float* __restrict__ gp, int V, int Kh,
                                      float* __restrict__ grad_delta) {
  extern __shared__ float sred[];  // [S][Kx][3]
  const int b = blockIdx.x;
  const int k = threadIdx.x, s = threadIdx.y, S = blockDim.y, Kx = blockDim.x;
  float a0 = 0.f, a1 = 0.f, a2 = 0.f;
  if (k < Kh) {
    const float* g = gp + (size_t)b * V * 3;
#pragma unroll 4
    for (int v = s; v < V; v += S) {
      const float w = Wm[(size_t)v * Kh + k];
      a0 = fmaf(w, g[v * 3], a0);
      a1 = fmaf(w, g[v * 3 + 1], a1);
      a2 = fmaf(w, g[v * 3 + 2], a2);
    }
  }
  sred[(s * Kx + k) * 3] = a0; sred[(s * Kx + k) * 3 + 1] = a1; sred[(s * Kx + k) * 3 + 2] = a2;
  __syncthreads();
  if (s == 0 && k < Kh) {
    for (int j = 1; j < S; ++j) { a0 += sred[(j * Kx + k) * 3]; a1 += sred[(j * Kx + k) * 3 + 1]; a2 += sred[(j * Kx + k) * 3 + 2]; }
    float* o = grad_delta + ((size_t)b * Kh + k) * 3;
    o[0] = a0; o[1] = a1; o[2] = a2;
  }
}

// grad_W[v,k] = sum_b gp[b,v,:] . delta[b,k,:];  grad_mean[v,:] = sum_b gp[b,v,:].  A warp owns 32 consecutive (v,k) entries,
// the CTA's 8 warps each take every 8th frame and the partial sums meet in shared memory in a fixed order (one thread per
// entry walking all frames was a chain of 64 L2 latencies on 4 warps per SM).
__global__ void __launch_bounds__(256) skin_bwd_w_kernel(const float* __restrict__ gp, const float* __restrict__ delta,
                                                         int NB, int V, int Kh, float* __restrict__ grad_W,
                                                         float* __restrict__ grad_mean) {
  __shared__ float red[8][32][4];
  const int lane = threadIdx.x & 31, sl = threadIdx.x >> 5;
  const int i = blockIdx.x * 32 + lane;
  const bool live = i < V * Kh;
  const int v = live ? i / Kh : 0, k = live ? i - v * Kh : 0;
  float acc = 0.f, m0 = 0.f, m1 = 0.f, m2 = 0.f;
  if (live) {
#pragma unroll 4
    for (int b = sl; b < NB; b += 8) {
      const float* g = gp + ((size_t)b * V + v) * 3;
      const float* d = delta + ((size_t)b * Kh + k) * 3;
      const float g0 = g[0], g1 = g[1], g2 = g[2];
      acc = fmaf(g0, d[0], fmaf(g1, d[1], fmaf(g2, d[2], acc)));
      m0 += g0; m1 += g1; m2 += g2;
    }
  }
  red[sl][lane][0] = acc; red[sl][lane][1] = m0; red[sl][lane][2] = m1; red[sl][lane][3] = m2;
  __syncthreads();
  if (sl == 0 && live) {
#pragma unroll
    for (int j = 1; j < 8; ++j) { acc += red[j][lane][0]; m0 += red[j][lane][1]; m1 += red[j][lane][2]; m2 += red[j][lane][3]; }
    if (grad_W) grad_W[i] = acc;
    if (grad_mean && k == 0) { grad_mean[v * 3] = m0; grad_mean[v * 3 + 1] = m1; grad_mean[v * 3 + 2] = m2; }
  }
}

// ---- handle weights: softmax over VERTICES per handle (MeshNet.get_lbs, mesh_net.py:597-599) -----------------
// x, y (V,K) row-major; one CTA per handle column.  torch's softmax over a strided dim 0 costs ~50 us at (642,32).
__global__ void __launch_bounds__(256) softmax_cols_fwd_kernel(const float* __restrict__ x, int V, int K, float* __restrict__ y) {
  __shared__ float red[8];
  __shared__ float bc;
  const int k = blockIdx.x, tid = threadIdx.x;
  float m = -INFINITY;
  for (int v = tid; v < V; v += 256) m = fmaxf(m, x[(size_t)v * K + k]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((tid & 31) == 0) red[tid >> 5] = m;
  __syncthreads();
  if (tid == 0) { float a = red[0]; for (int w = 1; w < 8; ++w) a = fmaxf(a, red[w]); bc = a; }
  __syncthreads();
  m = bc;
  float s = 0.0f;
  for (int v = tid; v < V; v += 256) s += expf(x[(size_t)v * K + k] - m);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  __syncthreads();
  if ((tid & 31) == 0) red[tid >> 5] = s;
  __syncthreads();
  if (tid == 0) { float a = 0.f; for (int w = 0; w < 8; ++w) a += red[w]; bc = a; }
  __syncthreads();
  const float inv = 1.0f / bc;
  for (int v = tid; v < V; v += 256) y[(size_t)v * K + k] = expf(x[(size_t)v * K + k] - m) * inv;
}

// gx = y * (gy - sum_v gy*y)
__global__ void __launch_bounds__(256) softmax_cols_bwd_kernel(const float* __restrict__ y, const float* __restrict__ gy, int V, int K,
                                                               float* __restrict__ gx) {
  __shared__ float red[8];
  __shared__ float bc;
  const int k = blockIdx.x, tid = threadIdx.x;
  float s = 0.0f;
  for (int v = tid; v < V; v += 256) s += gy[(size_t)v * K + k] * y[(size_t)v * K + k];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((tid & 31) == 0) red[tid >> 5] = s;
  __syncthreads();
  if (tid == 0) { float a = 0.f; for (int w = 0; w < 8; ++w) a += red[w]; bc = a; }
  __syncthreads();
  const float d = bc;
  for (int v = tid; v < V; v += 256) gx[(size_t)v * K + k] = y[(size_t)v * K + k] * (gy[(size_t)v * K + k] - d);
}

}  // namespace

extern "C" int acfm_skin_project_fwd(const float* mean_v, const float* W, const float* delta, const float* cams, int NB,
                                     int G, int V, int Kh, float offset_z, float sx, float sy, float z_add,
                                     float* pred_v, float* ndc, void* stream) {
  ACFM_REQUIRE(NB >= 0 && G >= 0 && V >= 0 && Kh >= 0, ACFM_ERR_BAD_ARG, "acfm_skin_project_fwd: bad sizes");
  ACFM_REQUIRE(Kh <= kMaxHandles, ACFM_ERR_UNSUPPORTED, "acfm_skin_project_fwd: %d handles > %d", Kh, kMaxHandles);
  if (NB == 0 || V == 0) return ACFM_OK;
  ACFM_REQUIRE(mean_v && (W || Kh == 0) && (delta || Kh == 0), ACFM_ERR_BAD_ARG, "acfm_skin_project_fwd: null input");
  ACFM_REQUIRE(pred_v || ndc, ACFM_ERR_BAD_ARG, "acfm_skin_project_fwd: no output requested");
  ACFM_REQUIRE(!ndc || (cams && G > 0), ACFM_ERR_BAD_ARG, "acfm_skin_project_fwd: ndc output needs cams and G > 0");
  ACFM_REQUIRE(NB <= 65535, ACFM_ERR_UNSUPPORTED, "acfm_skin_project_fwd: NB=%d > 65535", NB);
  dim3 grid((V + kThreads - 1) / kThreads, NB);
  skin_project_fwd_kernel<<<grid, kThreads, 0, (cudaStream_t)stream>>>(mean_v, W, delta, cams, NB, G, V, Kh, offset_z, sx,
                                                                        sy, z_add, pred_v, ndc);
  ACFM_LAUNCH_OK("skin_project_fwd_kernel");
  return ACFM_OK;
}

extern "C" int acfm_skin_bwd(const float* W, const float* delta, const float* grad_pred_v, int NB, int V, int Kh,
                             float* grad_delta, float* grad_W, float* grad_mean_v, void* stream) {
  ACFM_REQUIRE(NB >= 0 && V >= 0 && Kh >= 0, ACFM_ERR_BAD_ARG, "acfm_skin_bwd: bad sizes");
  ACFM_REQUIRE(Kh <= kMaxHandles, ACFM_ERR_UNSUPPORTED, "acfm_skin_bwd: %d handles > %d", Kh, kMaxHandles);
  if (NB == 0 || V == 0) return ACFM_OK;
  ACFM_REQUIRE(grad_pred_v && (W || Kh == 0) && (delta || Kh == 0), ACFM_ERR_BAD_ARG, "acfm_skin_bwd: null input");
  cudaStream_t st = (cudaStream_t)stream;
  if (grad_delta && Kh > 0) {
    const int Kx = ((Kh + 31) / 32) * 32, S = 1024 / Kx;  // (32 vertex slices at Kh <= 32: 20 dependent loads per thread, not 80)
    dim3 block(Kx, S);
    skin_bwd_delta_kernel<<<NB, block, sizeof(float) * 3 * Kx * S, st>>>(W, grad_pred_v, V, Kh, grad_delta);
    ACFM_LAUNCH_OK("skin_bwd_delta_kernel");
  }
  if ((grad_W && Kh > 0) || grad_mean_v) {
    const int Kk = Kh > 0 ? Kh : 1;
    ACFM_REQUIRE(Kh > 0, ACFM_ERR_UNSUPPORTED, "acfm_skin_bwd: grad_mean_v with zero handles");
    skin_bwd_w_kernel<<<(V * Kk + 31) / 32, 256, 0, st>>>(grad_pred_v, delta, NB, V, Kh, grad_W, grad_mean_v);
    ACFM_LAUNCH_OK("skin_bwd_w_kernel");
  }
  return ACFM_OK;
}

extern "C" int acfm_softmax_cols_fwd(const float* x, int V, int K, float* y, void* stream) {
  ACFM_REQUIRE(V >= 0 && K >= 0, ACFM_ERR_BAD_ARG, "acfm_softmax_cols_fwd: bad sizes");
  if (V == 0 || K == 0) return ACFM_OK;
  ACFM_REQUIRE(x && y, ACFM_ERR_BAD_ARG, "acfm_softmax_cols_fwd: null pointer");
  softmax_cols_fwd_kernel<<<K, 256, 0, (cudaStream_t)stream>>>(x, V, K, y);
  ACFM_LAUNCH_OK("softmax_cols_fwd_kernel");
  return ACFM_OK;
}

extern "C" int acfm_softmax_cols_bwd(const float* y, const float* grad_y, int V, int K, float* grad_x, void* stream) {
  ACFM_REQUIRE(V >= 0 && K >= 0, ACFM_ERR_BAD_ARG, "acfm_softmax_cols_bwd: bad sizes");
  if (V == 0 || K == 0) return ACFM_OK;
  ACFM_REQUIRE(y && grad_y && grad_x, ACFM_ERR_BAD_ARG, "acfm_softmax_cols_bwd: null pointer");
  softmax_cols_bwd_kernel<<<K, 256, 0, (cudaStream_t)stream>>>(y, grad_y, V, K, grad_x);
  ACFM_LAUNCH_OK("softmax_cols_bwd_kernel");
  return ACFM_OK;
}
