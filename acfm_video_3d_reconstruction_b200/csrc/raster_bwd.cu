// raster_bwd.cu — backward of the rasterizer for gradients on dists, with the soft-silhouette blend
// backward fused in.  Replaces _C.rasterize_meshes_backward + the autograd of sigmoid_alpha_blend
// (PyTorch3D 0.3.0) as reached from NeuralRenderer.forward (/root/reference/multiframe/nnutils/nmr.py:143-172).
// Semantics: SURVEY.md §9.5-9.6.
#include "common.cuh"
#include "raster_common.cuh"

namespace {

// ---------------------------------------------------------------------------------------------
// Silhouette backward.  One CTA per (render, 32x32 region).  Pixels whose mask is exactly 0 have no
// fragments (every kept fragment has prob >= ~1e-4) and pixels with zero upstream gradient contribute
// nothing: both are dropped while the region's pixels are compacted into a shared list, so the fragment
// tensors of ~87% of the pixels are never read.  A lane then owns one ACTIVE pixel and walks its K
// fragments starting at a lane-dependent offset, so that neighbouring pixels (which see the same faces
// at the same depth rank) touch different vertices at the same time.  Shared-memory FLOAT atomics are CAS spin
// loops on sm_100 (ATOMS.CAST.SPIN) while 32-bit INTEGER adds are native fire-and-forget ATOMS.ADD, so the silhouette
// path accumulates in fixed point: a per-CTA power-of-two scale is derived from the largest |grad_mask (1 - mask)| of the
// region (every contribution with |q - p| <= kRmax is then bounded by 2^22), converted by one FFMA onto a magic constant (no
// F2I: that is an XU-pipe instruction the atomics would wait for) and added to ONE int32 per vertex component — the shared
// atomic unit retires about four lanes per cycle whatever the addresses (ncu: 5.2 wavefronts per ATOMS at 19 active lanes,
// with or without lanes meeting at a vertex), so the number of atomics is what the kernel pays for: four per fragment, not
// the eight of a high / low pair of planes.  An int32 cannot hold the worst case (2^16 contributions of 2^22), so the
// headroom is VERIFIED instead of assumed: every lane sums the magnitudes it adds, the CTA checks sum |contribution| * scale
// < 2^30 — a bound on every accumulator, wrap-around included: two's-complement sums are exact whenever the true result
// fits — and in the rare region that fails (thousands of large contributions) clears the accumulators and runs the region
// again with the scale that passes.  Exact to 2^-22 of the bound (2^-k less after a retry) and independent of the order of
// the additions.  The rare contribution beyond the bound (a
// face larger than kRmax on screen) goes straight to global memory as a float atomic.  Gradients reach HBM as one
// atomicAdd per touched (CTA, vertex, component) instead of PyTorch3D's 4-9 global atomics per fragment.
// ---------------------------------------------------------------------------------------------
struct BwdParams {
  const float* ndc;
  const void* faces;
  long long faces_stride;
  int N, V, F, H, W, K;
  float sigma;
  const long long* p2f;
  const float* dists;
  const float* mask;
  const float* grad_mask;   // d loss / d mask (N,H,W), or NULL when everything comes through grad_sums
  // fused silhouette losses: d loss / d (the four per-render sums of acfm_raster_fwd_losses), targets at render n % NB
  const float* grad_sums;   // (N,4) or NULL
  const float* loss_target; // (NB,H,W)
  const float* loss_edt;    // (NB,H,W) or NULL
  int NB;
  float headroom;           // fixed-point accumulators: largest admissible sum of scaled magnitudes (2^30)
  const float* grad_dists;  // FROM_MASK == false: upstream gradient per fragment (N,H,W,K)
  float* grad_ndc;
  int regions_x, regions_y;
  const int* work;  // optional: the forward's work lists (live regions by weight class); NULL = every region
  // lean mode: the forward's compact fragments [work-list slot][pixel of the region][K] instead of p2f / dists
  const unsigned short* lean_f;
  const float* lean_d;
};

struct BwdSmem {
  int off_verts, off_faces, off_acc, off_list, total;
  __host__ __device__ BwdSmem(int V, int F) {
    int o = 32;
    off_verts = o; o += ((V * 8 + 15) / 16) * 16;  // (x, y) per vertex
    off_faces = o; o += F * 8;
    off_acc = o; o += ((V * 8 + 15) / 16) * 16;  // (V,2) int32 (fixed point) or float
    off_list = o; o += kRegion * kRegion * 2;
    total = o;
  }
};

__device__ __forceinline__ float rcp_fast(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float ex2_fast(float x) {
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}

// Squared distance from p to segment ab with the clamped parameter t and the foot point q (fast arithmetic: the backward
// is held to 1e-3 relative, not to bit parity; only the CHOICE of the closest edge has to agree with the forward, and near a
// tie either choice has the same gradient to that tolerance).  dax, day = p - a.  Branch free: a degenerate edge (|ab|^2 = 0 or
// denormal) gives t = 0 * inf = NaN or +-inf, which the clamp turns into 0 or 1 — q = a = b either way.
__device__ __forceinline__ float seg_foot(float px, float py, float ax, float ay, float dax, float day, float bx, float by, float& t,
                                          float& qx, float& qy) {
  const float bax = bx - ax, bay = by - ay;
  const float l2 = fmaf(bax, bax, bay * bay);
  t = fmaf(bax, dax, bay * day) * rcp_fast(l2);
  t = fminf(fmaxf(t, 0.0f), 1.0f);
  qx = fmaf(t, bax, ax); qy = fmaf(t, bay, ay);
  const float dx = qx - px, dy = qy - py;
  return fmaf(dx, dx, dy * dy);
}

// fixed-point bound on |q - p| (NDC): the blur band is 0.03 wide and a pixel inside a face is at most the face's inradius from
// its closest edge, so faces up to ~1/8 of the screen wide stay in range.  The bound sets the quantum of the accumulation
// (bound / 2^22 per contribution): 1/16 with 22 bits resolves what 1/4 with 24 bits did.
constexpr float kRmax = 0.0625f;

// accumulation policy of one CTA: fixed point (silhouette path) or float CAS (general path).  add_edge() takes the gradient
// (gx, gy) of one fragment w.r.t. its foot point and splits it between the edge's vertices a, b by (1 - t, t).
constexpr float kFxMagic = 12582912.0f;       // 1.5 * 2^23: float -> int by one FFMA (|value| <= 2^22), no F2I on the XU pipe
constexpr int kFxMagicBits = 0x4B400000;
// (Measured worse at C2: a high and a low int32 plane per component — eight atomics per fragment, headroom by construction —
// 0.58 ms against 0.42 ms; {x, y} interleaved with one base address per vertex; giving a warp 32 pixels that
// lie far apart in the region instead of 32 neighbours — 0.57 -> 0.71 ms: neighbours reading the same face and vertex words is
// what keeps the shared loads cheap; starting every lane at its own depth rank — no change.)
struct AccFixed {
  int* acc; float scale; float* gout; bool far_ok;  // gout: global fallback for out-of-range contributions (first pass only)
  float sx = 0.0f, sy = 0.0f;                       // magnitudes this lane has added (headroom check)
  __device__ __forceinline__ void add1(int* a, float c) const {
    atomicAdd(a, __float_as_int(fmaf(c, scale, kFxMagic)) - kFxMagicBits);  // round to nearest even, like F2I.RN
  }
  __device__ __forceinline__ void add_edge(int ia, int ib, float t, float gx, float gy, bool in_range) {
    const float s = 1.0f - t;
    if (in_range) {
      int* a = acc + ia * 2;
      int* b = acc + ib * 2;
      add1(a, s * gx); add1(a + 1, s * gy); add1(b, t * gx); add1(b + 1, t * gy);
      sx += fabsf(gx); sy += fabsf(gy);
    } else if (far_ok) {
      atomicAdd(gout + ia * 3, s * gx); atomicAdd(gout + ia * 3 + 1, s * gy);
      atomicAdd(gout + ib * 3, t * gx); atomicAdd(gout + ib * 3 + 1, t * gy);
    }
  }
};
struct AccFloat {
  float* acc;
  __device__ __forceinline__ void add_edge(int ia, int ib, float t, float gx, float gy, bool) {
    const float s = 1.0f - t;
    atomicAdd(acc + ia * 2, s * gx); atomicAdd(acc + ia * 2 + 1, s * gy);
    atomicAdd(acc + ib * 2, t * gx); atomicAdd(acc + ib * 2 + 1, t * gy);
  }
};

// gradient of one fragment's squared distance w.r.t. the two vertices of its closest edge (PointLineDistanceBackward,
// SURVEY.md §9.6), accumulated into the CTA's per-vertex accumulator
template <typename Acc>
__device__ __forceinline__ void frag_grad(float px, float py, float2 v0, float2 v1, float2 v2, float g, int i0, int i1, int i2, Acc& acc) {
  float t01, t02, t12, qx01, qy01, qx02, qy02, qx12, qy12;
  const float d0x = px - v0.x, d0y = py - v0.y, d1x = px - v1.x, d1y = py - v1.y;
  const float d01 = seg_foot(px, py, v0.x, v0.y, d0x, d0y, v1.x, v1.y, t01, qx01, qy01);
  const float d02 = seg_foot(px, py, v0.x, v0.y, d0x, d0y, v2.x, v2.y, t02, qx02, qy02);
  const float d12 = seg_foot(px, py, v1.x, v1.y, d1x, d1y, v2.x, v2.y, t12, qx12, qy12);
  // same order as the forward's min: 01, then 02, then 12
  const bool c01 = d01 <= d02 && d01 <= d12, c02 = d02 <= d12;
  const float t = c01 ? t01 : (c02 ? t02 : t12);
  const float qx = c01 ? qx01 : (c02 ? qx02 : qx12), qy = c01 ? qy01 : (c02 ? qy02 : qy12);
  const float dm = c01 ? d01 : (c02 ? d02 : d12);
  const int ia = (c01 || c02) ? i0 : i1, ib = c01 ? i1 : i2;
  const float g2 = g + g;
  acc.add_edge(ia, ib, t, g2 * (qx - px), g2 * (qy - py), dm <= kRmax * kRmax);
}

// d loss / d mask at one pixel: the caller's grad_mask and / or the backward of the fused per-render sums
// { sum|m-t|, sum m t, sum (m+t-mt), sum edt m } (losses.cu: mask_sums_bwd_kernel has the same expression), so that a loss
// computed by acfm_raster_fwd_losses never materialises grad_mask.
__device__ __forceinline__ float upstream_grad(const BwdParams& p, int n, long long pix, int x, int y, float m) {
  float g = p.grad_mask ? p.grad_mask[pix] : 0.0f;
  if (p.grad_sums) {
    const float* gs = p.grad_sums + (size_t)n * 4;
    const size_t ti = ((size_t)(n % p.NB) * p.H + y) * p.W + x;
    const float t = p.loss_target[ti], d = m - t;
    g += gs[0] * (d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f)) + gs[1] * t + gs[2] * (1.f - t);
    if (p.loss_edt) g += gs[3] * p.loss_edt[ti];
  }
  return g;
}

// FROM_MASK: the upstream gradient is d loss / d mask and the blend backward (§9.5) is fused in;
// otherwise it is d loss / d dists per fragment (texture branch, general rasterize_meshes backward on dists).
// (71 registers, 22.6 KB of shared memory at the reference's templates: 7 CTAs per SM.  Capped at 63 registers for 8: no change;
// at 56 for 9: 0.42 -> 0.50 ms at C2.)
template <typename IdxT, bool FROM_MASK, bool LEAN = false>
__global__ void __launch_bounds__(128) raster_soft_bwd_kernel(const BwdParams p) {
  constexpr int NT = 128, NWARPS = 4;
  extern __shared__ __align__(16) unsigned char smem[];
  const BwdSmem L(p.V, p.F);
  int* nactive = reinterpret_cast<int*>(smem + 8);
  int* next_chunk = reinterpret_cast<int*>(smem + 12);
  ushort4* sfaces = reinterpret_cast<ushort4*>(smem + L.off_faces);
  float* accf = reinterpret_cast<float*>(smem + L.off_acc);
  int2* acci = reinterpret_cast<int2*>(smem + L.off_acc);
  float* ssum = reinterpret_cast<float*>(smem + 20);  // [2]: magnitudes added in x, in y (headroom check)
  float2* sxy = reinterpret_cast<float2*>(smem + L.off_verts);
  unsigned* gmax_bits = reinterpret_cast<unsigned*>(smem + 16);
  unsigned short* alist = reinterpret_cast<unsigned short*>(smem + L.off_list);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int regions = p.regions_x * p.regions_y;
  int unit = blockIdx.x;
  const size_t slot = blockIdx.x;  // position in the work lists: where the lean forward put this region's fragments
  if (p.work) {
    // the regions the forward found live (the others hold no fragment: nothing to differentiate), heaviest first — same
    // lists, same layout as raster_fwd_kernel reads (raster_fwd.cu: {.., live per weight class [4] at [4], lists at [8]})
    const int U = p.N * regions;
    int cls = 0;
#pragma unroll
    for (; cls < kWeightClasses; ++cls) {
      const int c = p.work[4 + cls];
      if (unit < c) break;
      unit -= c;
    }
    if (cls == kWeightClasses) return;
    unit = p.work[8 + cls * U + unit];
  }
  const int n = unit / regions;
  const int rg = unit - n * regions;
  const int px0 = (rg % p.regions_x) * kRegion, py0 = (rg / p.regions_x) * kRegion;
  const int K = p.K;

  if (tid == 0) { *nactive = 0; *next_chunk = 0; *gmax_bits = 0u; ssum[0] = 0.0f; ssum[1] = 0.0f; }
  __syncthreads();
  // ---- 1. compact the region's active pixels (coalesced reads of mask / grad_mask) -------------------
  for (int i0 = 0; i0 < kRegion * kRegion; i0 += NT) {
    const int i = i0 + tid;
    const int x = px0 + (i & (kRegion - 1)), y = py0 + (i / kRegion);
    bool act = false;
    float gabs = 0.0f;
    if (x < p.W && y < p.H) {
      const long long pix = ((long long)n * p.H + y) * p.W + x;
      if (FROM_MASK) {
        const float mv = p.mask[pix];
        gabs = mv != 0.0f ? fabsf(upstream_grad(p, n, pix, x, y, mv) * (1.0f - mv)) : 0.0f;  // |ga| sigma of the loop below
        act = (mv != 0.0f) && (gabs != 0.0f) && (gabs <= 3.0e38f);  // NaN / inf upstream gradients are dropped
        if (!act) gabs = 0.0f;
      } else {
        act = p.p2f[pix * K] >= 0;
      }
    }
    if (FROM_MASK) {
      gabs = fmaxf(gabs, __shfl_xor_sync(0xffffffffu, gabs, 16)); gabs = fmaxf(gabs, __shfl_xor_sync(0xffffffffu, gabs, 8));
      gabs = fmaxf(gabs, __shfl_xor_sync(0xffffffffu, gabs, 4)); gabs = fmaxf(gabs, __shfl_xor_sync(0xffffffffu, gabs, 2));
      gabs = fmaxf(gabs, __shfl_xor_sync(0xffffffffu, gabs, 1));
      if (lane == 0 && gabs > 0.0f) atomicMax(gmax_bits, __float_as_uint(gabs));  // non-negative floats order like their bits
    }
    const unsigned m = __ballot_sync(0xffffffffu, act);
    int base = 0;
    if (lane == 0 && m) base = atomicAdd(nactive, __popc(m));
    base = __shfl_sync(0xffffffffu, base, 0);
    if (act) alist[base + __popc(m & ((1u << lane) - 1u))] = (unsigned short)i;
  }
  __syncthreads();
  const int na = *nactive;
  if (na == 0) return;

  const long long fbase = (long long)n * p.faces_stride;
  for (int f = tid; f < p.F; f += NT) {
    const IdxT* fp = reinterpret_cast<const IdxT*>(p.faces) + fbase + (long long)f * 3;
    sfaces[f] = make_ushort4((unsigned short)fp[0], (unsigned short)fp[1], (unsigned short)fp[2], 0);
  }
  const float* gv3 = p.ndc + (size_t)n * p.V * 3;
  for (int i = tid; i < p.V; i += NT) {
    acci[i] = make_int2(0, 0);  // (as floats: 0.0f twice over)
    sxy[i] = make_float2(gv3[i * 3], gv3[i * 3 + 1]);  // x, y only: one 8-byte shared load per vertex in the fragment loop
  }
  __syncthreads();

  // ---- 2. one active pixel per lane, 32 at a time, chunks pulled dynamically --------------------------
  const float inv_sigma = 1.0f / p.sigma;
  const float sig_l2e = inv_sigma * 1.4426950408889634f;  // prob = 1 / (1 + 2^(d log2(e) / sigma))
  const float inv_w = 1.0f / (float)p.W, inv_h = 1.0f / (float)p.H;
  float* gout = p.grad_ndc + (size_t)n * p.V * 3;
  // |contribution| <= |grad_mask (1 - mask)| / sigma * 2 |q - p| (prob, t <= 1)  =>  bound = gmax / sigma * 2 kRmax, rounded up
  // to a power of two; scale maps the bound to 2^22
  float fx_scale = 1.0f;
  if (FROM_MASK) {
    int e;
    frexpf(__uint_as_float(*gmax_bits) * inv_sigma * (2.0f * kRmax), &e);  // bound < 2^e (gmax is finite and > 0 here)
    fx_scale = ldexpf(1.0f, 22 - max(e, -100));  // (clamped: a vanishing bound must not push the scale to infinity)
  }
  AccFixed accx{reinterpret_cast<int*>(acci), fx_scale, gout, true};
  AccFloat accl{accf};
  const int nchunks = (na + 31) / 32;
  const bool vec = LEAN || ((K & 3) == 0 && (((uintptr_t)p.p2f | (uintptr_t)p.dists | (FROM_MASK ? 0 : (uintptr_t)p.grad_dists)) & 15u) == 0);
  int passes = 1;
pass_again:
  while (true) {
    int c = 0;
    if (lane == 0) c = atomicAdd(next_chunk, 1);
    c = __shfl_sync(0xffffffffu, c, 0);
    if (c >= nchunks) break;
    const int a = c * 32 + lane;
    if (a >= na) continue;
    const int i = alist[a];
    const int xi = px0 + (i & (kRegion - 1)), yi = py0 + (i / kRegion);
    const long long pix = ((long long)n * p.H + yi) * p.W + xi;
    // pixel centre (PixToNdc); 2 ulp is plenty for the backward
    const float xf = fmaf((float)(2 * (p.W - 1 - xi) + 1), inv_w, -1.0f), yf = fmaf((float)(2 * (p.H - 1 - yi) + 1), inv_h, -1.0f);
    const long long* pf = LEAN ? nullptr : p.p2f + pix * K;
    const float* pd = LEAN ? p.lean_d + (slot * (kRegion * kRegion) + i) * K : p.dists + pix * K;
    const unsigned short* lf = LEAN ? p.lean_f + (slot * (kRegion * kRegion) + i) * K : nullptr;
    const long long nF = LEAN ? 0 : (long long)n * p.F;
    // one group of four fragment ids: packed ids n F + f (-1: none), or, lean, the render's own face ids (0xffff: none) widened
    auto load_ids = [&](int g4, longlong2& a, longlong2& b) {
      if constexpr (LEAN) {
        const uint2 w = *reinterpret_cast<const uint2*>(lf + 4 * g4);
        const unsigned f0 = w.x & 0xffffu, f1 = w.x >> 16, f2 = w.y & 0xffffu, f3 = w.y >> 16;
        a.x = f0 == 0xffffu ? -1ll : (long long)f0; a.y = f1 == 0xffffu ? -1ll : (long long)f1;
        b.x = f2 == 0xffffu ? -1ll : (long long)f2; b.y = f3 == 0xffffu ? -1ll : (long long)f3;
      } else {
        a = *reinterpret_cast<const longlong2*>(pf + 4 * g4);
        b = *reinterpret_cast<const longlong2*>(pf + 4 * g4 + 2);
      }
    };
    // fragments in groups of four (16-byte loads); the first group's loads are issued before the upstream gradient is formed
    // and every later group's while the one before it is processed: the kernel is bound by the latency of these loads and of
    // the shared atomics, so one group is always in flight.  Lanes start at different groups so that neighbouring pixels,
    // which see the same faces at the same depth rank, touch different vertices at the same time.
    const int groups = K >> 2;
    int g = vec ? lane % groups : 0;
    longlong2 fa = make_longlong2(-1, -1), fb = fa;
    float4 dd = make_float4(0.f, 0.f, 0.f, 0.f), gg = dd;
    if (vec) {
      load_ids(g, fa, fb);
      dd = *reinterpret_cast<const float4*>(pd + 4 * g);
      if (!FROM_MASK) gg = *reinterpret_cast<const float4*>(p.grad_dists + pix * K + 4 * g);
    }
    // d mask / d dist_k = -(prob_k / sigma) * prod_j (1 - prob_j)   (SURVEY.md §9.5); the product is 1 - mask, which the
    // forward already formed (its rounding only matters where the product, hence the gradient, is < 1e-7 of the largest)
    float ga = 1.0f;
    if (FROM_MASK) {
      const float mv = p.mask[pix];
      ga = -upstream_grad(p, n, pix, xi, yi, mv) * (1.0f - mv) * inv_sigma;
      if (ga == 0.0f) continue;
    }
    if (vec) {
      // (Measured without gain: combining the lanes that hit the same vertex with __match_any_sync + __reduce_add_sync before
      // the shared atomics — 0.63 -> 1.98 ms at C2, the match is far slower than the conflicts it removes; saving the closest
      // edge per fragment in the forward so that one foot point is computed instead of three — 0.64 -> 0.63 ms for +0.07 ms
      // of forward: the kernel waits on fragment loads and shared atomics, not on arithmetic.)
      for (int s = 0; s < groups; ++s) {
        const long long fid[4] = {fa.x, fa.y, fb.x, fb.y};
        const float dv[4] = {dd.x, dd.y, dd.z, dd.w};
        const float gv[4] = {gg.x, gg.y, gg.z, gg.w};
        g = (g + 1 == groups) ? 0 : g + 1;
        if (s + 1 < groups) {  // next group: in flight during this one's arithmetic
          load_ids(g, fa, fb);
          dd = *reinterpret_cast<const float4*>(pd + 4 * g);
          if (!FROM_MASK) gg = *reinterpret_cast<const float4*>(p.grad_dists + pix * K + 4 * g);
        }
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          if (fid[e] < 0) break;  // lists are front-packed
          const float d = dv[e];
          float gd = FROM_MASK ? ga * rcp_fast(1.0f + ex2_fast(d * sig_l2e)) : gv[e];
          if (gd == 0.0f) continue;
          if (signbit(d)) gd = -gd;  // dist = inside ? -|d| : |d|
          const ushort4 iv = sfaces[(int)((unsigned)fid[e] - (unsigned)nF)];  // (0 <= fid - nF < F: the low words suffice)
          if (FROM_MASK) frag_grad(xf, yf, sxy[iv.x], sxy[iv.y], sxy[iv.z], gd, iv.x, iv.y, iv.z, accx);
          else frag_grad(xf, yf, sxy[iv.x], sxy[iv.y], sxy[iv.z], gd, iv.x, iv.y, iv.z, accl);
        }
      }
    } else {
      int cnt = 0;
      while (cnt < K && pf[cnt] >= 0) ++cnt;  // lists are front-packed
      if (cnt == 0) continue;
      int k = (lane * 7) % cnt;
      for (int s = 0; s < cnt; ++s) {
        const float d = pd[k];
        const int f = (int)(pf[k] - nF);
        float gd = FROM_MASK ? ga * rcp_fast(1.0f + ex2_fast(d * sig_l2e)) : p.grad_dists[pix * K + k];
        k = (k + 1 == cnt) ? 0 : k + 1;
        if (gd == 0.0f) continue;
        if (signbit(d)) gd = -gd;
        const ushort4 iv = sfaces[f];
        if (FROM_MASK) frag_grad(xf, yf, sxy[iv.x], sxy[iv.y], sxy[iv.z], gd, iv.x, iv.y, iv.z, accx);
        else frag_grad(xf, yf, sxy[iv.x], sxy[iv.y], sxy[iv.z], gd, iv.x, iv.y, iv.z, accl);
      }
    }
  }
  if (FROM_MASK) {
    // headroom check: sum over the region of |gx| (|gy|) bounds the magnitude of every x (y) accumulator's true value
    float sx = accx.sx, sy = accx.sy;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { sx += __shfl_xor_sync(0xffffffffu, sx, o); sy += __shfl_xor_sync(0xffffffffu, sy, o); }
    if (lane == 0) { atomicAdd(&ssum[0], sx); atomicAdd(&ssum[1], sy); }
  }
  __syncthreads();
  if (FROM_MASK) {
    const float top = fmaxf(ssum[0], ssum[1]) * accx.scale;  // (+ half a unit of rounding per addition: far inside the 2x margin)
    if (top >= p.headroom && top <= 3.0e38f && ++passes <= 4) {  // (never for NaN / inf: garbage in, garbage out, but no loop)
      // run the region again with a scale that passes (the out-of-range contributions went to global memory already)
      int e;
      frexpf(top / p.headroom, &e);  // top / headroom < 2^e, e >= 1
      __syncthreads();                            // everyone has read ssum
      for (int i = tid; i < p.V; i += NT) acci[i] = make_int2(0, 0);
      if (tid == 0) { *next_chunk = 0; ssum[0] = 0.0f; ssum[1] = 0.0f; }
      accx.scale = ldexpf(accx.scale, -e); accx.far_ok = false; accx.sx = 0.0f; accx.sy = 0.0f;
      __syncthreads();
      goto pass_again;
    }
  }
  const float inv_scale = 1.0f / accx.scale;
  for (int i = tid; i < p.V * 2; i += NT) {
    const float a = FROM_MASK ? (float)reinterpret_cast<const int*>(acci)[i] * inv_scale : accf[i];
    if (a != 0.0f) atomicAdd(gout + (i >> 1) * 3 + (i & 1), a);
  }
}

}  // namespace

namespace {
int launch_bwd(const char* who, bool from_mask, const float* ndc, const void* faces, int faces_i64, int64_t faces_batch_stride,
               int N, int V, int F, int H, int W, int K, float sigma, const int64_t* pix_to_face, const float* dists,
               const float* mask, const float* grad_mask, const float* grad_dists, float* grad_ndc, const void* work, void* stream,
               const float* grad_sums = nullptr, const float* loss_target = nullptr, const float* loss_edt = nullptr, int NB = 1,
               const void* lean_workspace = nullptr) {
  ACFM_REQUIRE(N >= 0 && V >= 0 && F >= 0 && H > 0 && W > 0 && K >= 1, ACFM_ERR_BAD_ARG, "%s: bad sizes", who);
  ACFM_REQUIRE(!from_mask || sigma > 0.0f, ACFM_ERR_BAD_ARG, "%s: sigma must be > 0", who);
  ACFM_REQUIRE(faces_batch_stride == 0 || faces_batch_stride == (int64_t)F * 3, ACFM_ERR_BAD_ARG, "%s: faces_batch_stride must be 0 or F*3", who);
  if (N == 0 || V == 0) return ACFM_OK;
  const bool lean = lean_workspace != nullptr;
  ACFM_REQUIRE(ndc && faces && grad_ndc && (lean || (pix_to_face && dists)), ACFM_ERR_BAD_ARG, "%s: null pointer", who);
  ACFM_REQUIRE(!lean || (from_mask && work && (K & 3) == 0), ACFM_ERR_BAD_ARG, "%s: the lean backward needs the forward's workspace", who);
  ACFM_REQUIRE(from_mask ? (mask && (grad_mask || grad_sums)) : (grad_dists != nullptr), ACFM_ERR_BAD_ARG, "%s: null gradient pointer", who);
  ACFM_REQUIRE(!grad_sums || (loss_target && NB > 0 && N % NB == 0), ACFM_ERR_BAD_ARG, "%s: fused losses need the target and N %% NB == 0", who);
  ACFM_REQUIRE(F <= 65535 && V <= 65535, ACFM_ERR_UNSUPPORTED, "%s: V=%d, F=%d must be <= 65535", who, V, F);
  cudaStream_t st = (cudaStream_t)stream;
  ACFM_CUDA_OK(cudaMemsetAsync(grad_ndc, 0, sizeof(float) * 3 * (size_t)N * V, st));
  if (F == 0) return ACFM_OK;
  BwdParams p;
  p.ndc = ndc; p.faces = faces; p.faces_stride = faces_batch_stride;
  p.N = N; p.V = V; p.F = F; p.H = H; p.W = W; p.K = K; p.sigma = from_mask ? sigma : 1.0f;
  p.p2f = (const long long*)pix_to_face; p.dists = dists; p.mask = mask; p.grad_mask = grad_mask; p.grad_dists = grad_dists;
  p.grad_ndc = grad_ndc;
  p.headroom = ldexpf(1.0f, acfm_raster_bwd_headroom_bits());
  p.grad_sums = grad_sums; p.loss_target = loss_target; p.loss_edt = loss_edt; p.NB = NB > 0 ? NB : 1;
  p.work = (const int*)work;
  p.regions_x = (W + kRegion - 1) / kRegion; p.regions_y = (H + kRegion - 1) / kRegion;
  {
    const long long U = (long long)N * p.regions_x * p.regions_y;
    p.lean_d = (const float*)lean_workspace;                                            // layout: acfm_raster_fwd_lean
    p.lean_f = lean ? (const unsigned short*)((const float*)lean_workspace + U * (kRegion * kRegion) * K) : nullptr;
  }
  const BwdSmem L(V, F);
  ACFM_REQUIRE(L.total <= 227 * 1024, ACFM_ERR_UNSUPPORTED, "%s: needs %d B of shared memory (max 232448)", who, L.total);
  const long long ctas = (long long)N * p.regions_x * p.regions_y;
  ACFM_REQUIRE(ctas < (1ll << 31), ACFM_ERR_UNSUPPORTED, "%s: too many CTAs", who);
#define ACFM_LAUNCH_BWD(IDX, FM)                                                                                       \
  do {                                                                                                                 \
    static std::atomic<int> smem_set[kAcfmMaxDevices];                                                                 \
    ACFM_CUDA_OK(acfm_ensure_smem(raster_soft_bwd_kernel<IDX, FM>, L.total, smem_set));                                \
    raster_soft_bwd_kernel<IDX, FM><<<(int)ctas, 128, L.total, st>>>(p);                                               \
  } while (0)
#define ACFM_LAUNCH_BWD_LEAN(IDX)                                                                                      \
  do {                                                                                                                 \
    static std::atomic<int> smem_set[kAcfmMaxDevices];                                                                 \
    ACFM_CUDA_OK(acfm_ensure_smem(raster_soft_bwd_kernel<IDX, true, true>, L.total, smem_set));                        \
    raster_soft_bwd_kernel<IDX, true, true><<<(int)ctas, 128, L.total, st>>>(p);                                       \
  } while (0)
  if (lean) { if (faces_i64) ACFM_LAUNCH_BWD_LEAN(long long); else ACFM_LAUNCH_BWD_LEAN(int); }
  else if (faces_i64) { if (from_mask) ACFM_LAUNCH_BWD(long long, true); else ACFM_LAUNCH_BWD(long long, false); }
  else { if (from_mask) ACFM_LAUNCH_BWD(int, true); else ACFM_LAUNCH_BWD(int, false); }
#undef ACFM_LAUNCH_BWD_LEAN
#undef ACFM_LAUNCH_BWD
  ACFM_LAUNCH_OK("raster_soft_bwd_kernel");
  return ACFM_OK;
}
}  // namespace

extern "C" int acfm_raster_soft_bwd(const float* ndc, const void* faces, int faces_i64, int64_t faces_batch_stride,
                                    int N, int V, int F, int H, int W, int K, float sigma, const int64_t* pix_to_face,
                                    const float* dists, const float* mask, const float* grad_mask, float* grad_ndc,
                                    const void* fwd_workspace, void* stream) {
  return launch_bwd("acfm_raster_soft_bwd", true, ndc, faces, faces_i64, faces_batch_stride, N, V, F, H, W, K, sigma,
                    pix_to_face, dists, mask, grad_mask, nullptr, grad_ndc, fwd_workspace, stream);
}

extern "C" int acfm_raster_soft_bwd_train(const float* ndc, const void* faces, int faces_i64, int64_t faces_batch_stride, int N, int V,
                                          int F, int H, int W, int K, float sigma, const int64_t* pix_to_face, const float* dists,
                                          const float* mask, const float* grad_mask, const float* grad_sums,
                                          const float* target, const float* edt, int NB, float* grad_ndc, const void* fwd_workspace,
                                          void* stream) {
  return launch_bwd("acfm_raster_soft_bwd_train", true, ndc, faces, faces_i64, faces_batch_stride, N, V, F, H, W, K, sigma,
                    pix_to_face, dists, mask, grad_mask, nullptr, grad_ndc, fwd_workspace, stream, grad_sums, target, edt, NB);
}

extern "C" int acfm_raster_soft_bwd_lean(const float* ndc, const void* faces, int faces_i64, int64_t faces_batch_stride, int N, int V,
                                         int F, int H, int W, int K, float sigma, const float* mask, const float* grad_mask,
                                         const float* grad_sums, const float* target, const float* edt, int NB, float* grad_ndc,
                                         const void* lean_workspace, const void* fwd_workspace, void* stream) {
  ACFM_REQUIRE(lean_workspace && fwd_workspace, ACFM_ERR_BAD_ARG, "acfm_raster_soft_bwd_lean: needs the lean workspace and the workspace of the forward call");
  return launch_bwd("acfm_raster_soft_bwd_lean", true, ndc, faces, faces_i64, faces_batch_stride, N, V, F, H, W, K, sigma, nullptr,
                    nullptr, mask, grad_mask, nullptr, grad_ndc, fwd_workspace, stream, grad_sums, target, edt, NB, lean_workspace);
}

extern "C" int acfm_raster_dists_bwd(const float* ndc, const void* faces, int faces_i64, int64_t faces_batch_stride,
                                     int N, int V, int F, int H, int W, int K, const int64_t* pix_to_face,
                                     const float* dists, const float* grad_dists, float* grad_ndc, void* stream) {
  return launch_bwd("acfm_raster_dists_bwd", false, ndc, faces, faces_i64, faces_batch_stride, N, V, F, H, W, K, 0.0f,
                    pix_to_face, dists, nullptr, nullptr, grad_dists, grad_ndc, nullptr, stream);
}
