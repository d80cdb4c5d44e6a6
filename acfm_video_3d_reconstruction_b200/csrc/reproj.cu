// reproj.cu — fused reprojection losses on the projected mesh: visible-vertex bitmap, boundary loss,
// optical-flow loss, keypoint loss and the camera-hypothesis weighting.
//
// Replaces, in /root/reference/multiframe/nnutils/loss_utils.py (identical file in monocular/):
//   the `fi_maps -> faces_ -> unique -> scatter_` visibility block of bds_loss (:213-223) and
//   optical_flow_loss (:425-441)                                               -> acfm_visible_verts
//   bds_loss (:204-237): cdist^2 + invisible->1000 + topk(1) + masked sum      -> acfm_bds_loss_fwd/bwd
//   optical_flow_loss (:419-474): nearest grid_sample of the flow at the projected vertices, adjacent-frame
//   displacement, masked L1                                                    -> acfm_of_loss_fwd/bwd
//   kp_l2_loss (:341-356)                                                      -> acfm_kp_loss_fwd/bwd
// and the hypothesis softmax weighting of ShapeTrainer.forward (multiframe/main.py:735-746)
//                                                                              -> acfm_hypothesis_weight_fwd/bwd.
// All are tiny next to the rasterizer (a few KB per render); the point of fusing them is launch count:
// the reference spends ~60 small torch kernels, two device-wide `unique` sorts with host syncs and a
// (N,1000,V) cdist tensor on them per step.
#include "common.cuh"

namespace {

constexpr int kThreads = 256;

// ---------------------------------------------------------------------------------------------
// visible vertices: vis[n,v] = 1 iff v belongs to a face that is nearest at some pixel of render n
// ---------------------------------------------------------------------------------------------
template <typename IdxT>
__global__ void __launch_bounds__(kThreads) visible_verts_kernel(const long long* __restrict__ p2f, long long pix_stride,
                                                                 const IdxT* __restrict__ faces, long long faces_stride, int V,
                                                                 int F, int HW, float* __restrict__ vis) {
  const int n = blockIdx.y;
  const long long* src = p2f + (size_t)n * HW * pix_stride;
  const IdxT* fn = faces + (size_t)n * faces_stride;
  float* out = vis + (size_t)n * V;
  long long prev = -1;  // consecutive pixels of a thread's stride rarely repeat, neighbours in a warp do: dedupe below
  for (int i = blockIdx.x * kThreads + threadIdx.x; i < HW; i += gridDim.x * kThreads) {
    const long long id = src[(size_t)i * pix_stride];
    // one store per distinct face among the lanes that see it (neighbouring pixels share faces)
    const unsigned peers = __match_any_sync(__activemask(), id);
    if (id < 0 || id == prev || (__ffs(peers) - 1) != (int)(threadIdx.x & 31)) continue;
    prev = id;
    const long long f = id - (long long)n * F;  // packed id n*F + f
    if (f < 0 || f >= F) continue;
    const IdxT* t = fn + f * 3;
    out[(int)t[0]] = 1.0f; out[(int)t[1]] = 1.0f; out[(int)t[2]] = 1.0f;
  }
}

// ---------------------------------------------------------------------------------------------
// boundary loss.  grid (ceil(S / kThreads), N); verts + visibility of the render staged in shared memory
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads) bds_fwd_kernel(const float* __restrict__ verts, int vs, const float* __restrict__ vis,
                                                           const float* __restrict__ bds, const long long* __restrict__ sel,
                                                           int NB, int V, int P, int S, float* __restrict__ loss,
                                                           int* __restrict__ argmin) {
  extern __shared__ float sm[];  // x[V], y[V] of the visible vertices (compacted), then ids
  float* sx = sm;
  float* sy = sm + V;
  int* sid = reinterpret_cast<int*>(sm + 2 * V);
  __shared__ int nvis;
  __shared__ float red[kThreads / 32];
  const int n = blockIdx.y;
  if (threadIdx.x == 0) nvis = 0;
  __syncthreads();
  for (int v = threadIdx.x; v < V; v += kThreads) {
    if (vis[(size_t)n * V + v] != 0.0f) {
      const int k = atomicAdd(&nvis, 1);
      sx[k] = verts[((size_t)n * V + v) * vs];
      sy[k] = verts[((size_t)n * V + v) * vs + 1];
      sid[k] = v;
    }
  }
  __syncthreads();
  const int nv = nvis;
  const int j = blockIdx.x * kThreads + threadIdx.x;
  float contrib = 0.0f;
  if (j < S) {
    const long long pj = sel ? sel[j] : j;
    const float* b = bds + ((size_t)(n % NB) * P + pj) * 3;
    const float bx = b[0], by = b[1], bm = b[2];
    // invisible vertices count as distance 1000 (loss_utils.py:227): the minimum is min(1000 if any vertex is
    // invisible, nearest visible squared distance); ties resolve to the lowest vertex id for a stable gradient
    float best = (nv < V) ? 1000.0f : INFINITY;
    int bi = -1;
    for (int k = 0; k < nv; ++k) {
      const float dx = bx - sx[k], dy = by - sy[k];
      const float d = dx * dx + dy * dy;
      const int id = sid[k];
      if (d < best || (d == best && bi >= 0 && id < bi)) { best = d; bi = id; }
    }
    if (argmin) argmin[(size_t)n * S + j] = bi;
    contrib = best * bm;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) contrib += __shfl_xor_sync(0xffffffffu, contrib, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = contrib;
  __syncthreads();
  if (threadIdx.x == 0) {
    float a = 0.0f;
    for (int w = 0; w < kThreads / 32; ++w) a += red[w];
    if (gridDim.x == 1) loss[n] = a; else atomicAdd(loss + n, a);
  }
}

__global__ void __launch_bounds__(kThreads) bds_bwd_kernel(const float* __restrict__ verts, int vs, const float* __restrict__ bds,
                                                           const long long* __restrict__ sel, const int* __restrict__ argmin,
                                                           const float* __restrict__ grad_loss, int NB, int V, int P, int S,
                                                           float* __restrict__ grad_verts) {
  const int n = blockIdx.y;
  const int j = blockIdx.x * kThreads + threadIdx.x;
  if (j >= S) return;
  const int bi = argmin[(size_t)n * S + j];
  if (bi < 0) return;
  const long long pj = sel ? sel[j] : j;
  const float* b = bds + ((size_t)(n % NB) * P + pj) * 3;
  const float g = grad_loss[n] * b[2];
  if (g == 0.0f) return;
  const float* v = verts + ((size_t)n * V + bi) * vs;
  float* gv = grad_verts + ((size_t)n * V + bi) * vs;
  atomicAdd(gv, g * 2.0f * (v[0] - b[0]));
  atomicAdd(gv + 1, g * 2.0f * (v[1] - b[1]));
}

// ---------------------------------------------------------------------------------------------
// optical-flow loss.  One CTA per (sequence b, frame pair t -> t+1)
// ---------------------------------------------------------------------------------------------
// grid_sample(mode='nearest', align_corners=False, padding zeros) index of a normalised coordinate
__device__ __forceinline__ int nearest_index(float c, int size) {
  const float u = __fdiv_rn(__fsub_rn(__fmul_rn(__fadd_rn(c, 1.0f), (float)size), 1.0f), 2.0f);  // ((c+1)*size-1)/2
  return (int)nearbyintf(u);
}

__global__ void __launch_bounds__(kThreads) of_fwd_kernel(const float* __restrict__ proj, int ps, const float* __restrict__ vis,
                                                          const float* __restrict__ flows, int NBf, int T, int V, int H, int W,
                                                          float* __restrict__ loss, float* __restrict__ of_pred,
                                                          float* __restrict__ vis_out, float* __restrict__ samples) {
  const int b = blockIdx.x / (T - 1), t = blockIdx.x % (T - 1);
  const float* p0 = proj + ((size_t)(b * T + t) * V) * ps;        // current frame
  const float* p1 = proj + ((size_t)(b * T + t + 1) * V) * ps;    // next frame: visibility and flow are read here
  const float* fl = flows + ((size_t)((b % NBf) * T + t + 1) * H * W) * 2;
  const float* vz = vis + (size_t)(b * T + t + 1) * V;
  const size_t obase = ((size_t)b * (T - 1) + t) * V;
  float acc = 0.0f, cnt = 0.0f;
  for (int v = threadIdx.x; v < V; v += kThreads) {
    const float x1 = p1[(size_t)v * ps], y1 = p1[(size_t)v * ps + 1];
    const int ix = nearest_index(x1, W), iy = nearest_index(y1, H);
    float gx = 0.0f, gy = 0.0f;
    if (ix >= 0 && ix < W && iy >= 0 && iy < H) { gx = fl[((size_t)iy * W + ix) * 2]; gy = fl[((size_t)iy * W + ix) * 2 + 1]; }
    const bool vv = (fabsf(gx) + fabsf(gy) != 0.0f) && (vz[v] != 0.0f);
    const float m = vv ? 1.0f : 0.0f;
    // predicted_points_ = W * (p + 1) / 2 ; of_pred = current - next (loss_utils.py:454-459)
    const float wf = (float)W;
    const float cx = __fdiv_rn(__fmul_rn(wf, __fadd_rn(p0[(size_t)v * ps], 1.0f)), 2.0f);
    const float cy = __fdiv_rn(__fmul_rn(wf, __fadd_rn(p0[(size_t)v * ps + 1], 1.0f)), 2.0f);
    const float nx = __fdiv_rn(__fmul_rn(wf, __fadd_rn(x1, 1.0f)), 2.0f);
    const float ny = __fdiv_rn(__fmul_rn(wf, __fadd_rn(y1, 1.0f)), 2.0f);
    const float px = m * (cx - nx), py = m * (cy - ny);
    const float sx = m * gx, sy = m * gy;
    acc += fabsf(sx - px) + fabsf(sy - py);
    cnt += m;
    if (of_pred) { of_pred[(obase + v) * 2] = px; of_pred[(obase + v) * 2 + 1] = py; }
    if (samples) { samples[(obase + v) * 2] = sx; samples[(obase + v) * 2 + 1] = sy; }
    if (vis_out) vis_out[obase + v] = m;
  }
  __shared__ float red[2][kThreads / 32];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { acc += __shfl_xor_sync(0xffffffffu, acc, o); cnt += __shfl_xor_sync(0xffffffffu, cnt, o); }
  if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = acc; red[1][threadIdx.x >> 5] = cnt; }
  __syncthreads();
  if (threadIdx.x == 0) {
    float a = 0.0f, c = 0.0f;
    for (int w = 0; w < kThreads / 32; ++w) { a += red[0][w]; c += red[1][w]; }
    loss[(size_t)b * (T - 1) + t] = a / (float)H / (c + 1.0f);
  }
}

// grad_proj must be zeroed by the launcher: frame t receives from pair (t-1,t) as "next" and from (t,t+1) as "current"
__global__ void __launch_bounds__(kThreads) of_bwd_kernel(const float* __restrict__ vis_out, const float* __restrict__ of_pred,
                                                          const float* __restrict__ samples, const float* __restrict__ grad_loss,
                                                          int ps, int T, int V, int H, int W, float* __restrict__ grad_proj) {
  const int b = blockIdx.x / (T - 1), t = blockIdx.x % (T - 1);
  const size_t obase = ((size_t)b * (T - 1) + t) * V;
  __shared__ float s_cnt;
  __shared__ float red[kThreads / 32];
  float cnt = 0.0f;
  for (int v = threadIdx.x; v < V; v += kThreads) cnt += vis_out[obase + v];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = cnt;
  __syncthreads();
  if (threadIdx.x == 0) {
    float c = 0.0f;
    for (int w = 0; w < kThreads / 32; ++w) c += red[w];
    s_cnt = c;
  }
  __syncthreads();
  const float g = grad_loss[(size_t)b * (T - 1) + t] / (float)H / (s_cnt + 1.0f) * (0.5f * (float)W);
  float* g0 = grad_proj + ((size_t)(b * T + t) * V) * ps;
  float* g1 = grad_proj + ((size_t)(b * T + t + 1) * V) * ps;
  for (int v = threadIdx.x; v < V; v += kThreads) {
    if (vis_out[obase + v] == 0.0f) continue;
    // d|s - p|/dp = -sign(s - p); p = W/2 (cur - next)
    const float ex = samples[(obase + v) * 2] - of_pred[(obase + v) * 2];
    const float ey = samples[(obase + v) * 2 + 1] - of_pred[(obase + v) * 2 + 1];
    const float sx = ex > 0.f ? -1.f : (ex < 0.f ? 1.f : 0.f), sy = ey > 0.f ? -1.f : (ey < 0.f ? 1.f : 0.f);
    atomicAdd(g0 + (size_t)v * ps, g * sx); atomicAdd(g0 + (size_t)v * ps + 1, g * sy);
    atomicAdd(g1 + (size_t)v * ps, -g * sx); atomicAdd(g1 + (size_t)v * ps + 1, -g * sy);
  }
}

// ---------------------------------------------------------------------------------------------
// keypoint loss: one warp per render
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads) kp_fwd_kernel(const float* __restrict__ kp_pred, int ps, const float* __restrict__ kp_gt,
                                                          int N, int NB, int Kp, float* __restrict__ loss) {
  const int n = blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5);
  if (n >= N) return;
  const int lane = threadIdx.x & 31;
  float s = 0.0f, vs = 0.0f;
  for (int k = lane; k < Kp; k += 32) {
    const float* g = kp_gt + ((size_t)(n % NB) * Kp + k) * 3;
    const float* p = kp_pred + ((size_t)n * Kp + k) * ps;
    const float vis = g[2] > 0.0f ? 1.0f : 0.0f;
    s += (fabsf(p[0] - g[0]) + fabsf(p[1] - g[1])) * vis;
    vs += vis;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { s += __shfl_xor_sync(0xffffffffu, s, o); vs += __shfl_xor_sync(0xffffffffu, vs, o); }
  if (lane == 0) loss[n] = (s / (float)Kp) / (vs / (float)Kp + 1e-4f);
}

__global__ void __launch_bounds__(kThreads) kp_bwd_kernel(const float* __restrict__ kp_pred, int ps, const float* __restrict__ kp_gt,
                                                          const float* __restrict__ grad_loss, int N, int NB, int Kp,
                                                          float* __restrict__ grad_kp) {
  const int n = blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5);
  if (n >= N) return;
  const int lane = threadIdx.x & 31;
  float vs = 0.0f;
  for (int k = lane; k < Kp; k += 32) vs += kp_gt[((size_t)(n % NB) * Kp + k) * 3 + 2] > 0.0f ? 1.0f : 0.0f;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) vs += __shfl_xor_sync(0xffffffffu, vs, o);
  const float g = grad_loss[n] / (float)Kp / (vs / (float)Kp + 1e-4f);
  for (int k = lane; k < Kp; k += 32) {
    const float* gt = kp_gt + ((size_t)(n % NB) * Kp + k) * 3;
    const float* p = kp_pred + ((size_t)n * Kp + k) * ps;
    float* o = grad_kp + ((size_t)n * Kp + k) * ps;
    const float vis = gt[2] > 0.0f ? 1.0f : 0.0f;
    const float dx = p[0] - gt[0], dy = p[1] - gt[1];
    o[0] = g * vis * (dx > 0.f ? 1.f : (dx < 0.f ? -1.f : 0.f));
    o[1] = g * vis * (dy > 0.f ? 1.f : (dy < 0.f ? -1.f : 0.f));
    for (int c = 2; c < ps; ++c) o[c] = 0.0f;
  }
}

// ---------------------------------------------------------------------------------------------
// hypothesis weighting: probs = softmax(-L, dim 0) (detached); total = mean_m sum_g probs * L
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads) hyp_fwd_kernel(const float* __restrict__ L, int G, int M, float* __restrict__ probs,
                                                           float* __restrict__ total) {
  __shared__ float red[kThreads / 32];
  float acc = 0.0f;
  for (int m = blockIdx.x * kThreads + threadIdx.x; m < M; m += gridDim.x * kThreads) {
    float mx = -INFINITY;
    for (int g = 0; g < G; ++g) mx = fmaxf(mx, -L[(size_t)g * M + m]);
    float den = 0.0f;
    for (int g = 0; g < G; ++g) den += expf(-L[(size_t)g * M + m] - mx);
    float s = 0.0f;
    for (int g = 0; g < G; ++g) {
      const float l = L[(size_t)g * M + m];
      const float pr = expf(-l - mx) / den;
      probs[(size_t)g * M + m] = pr;
      s += pr * l;
    }
    acc += s;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float a = 0.0f;
    for (int w = 0; w < kThreads / 32; ++w) a += red[w];
    atomicAdd(total, a / (float)M);
  }
}

__global__ void __launch_bounds__(kThreads) hyp_bwd_kernel(const float* __restrict__ probs, const float* __restrict__ grad_total,
                                                           int GM, float inv_m, float* __restrict__ grad_L) {
  const float g = grad_total[0] * inv_m;
  for (int i = blockIdx.x * kThreads + threadIdx.x; i < GM; i += gridDim.x * kThreads) grad_L[i] = probs[i] * g;
}

}  // namespace

extern "C" int acfm_visible_verts(const int64_t* pix_to_face, int64_t pix_stride, const void* faces, int faces_i64,
                                  int64_t faces_batch_stride, int N, int V, int F, int HW, float* vis, void* stream) {
  ACFM_REQUIRE(N >= 0 && V >= 0 && F >= 0 && HW >= 0 && pix_stride >= 1, ACFM_ERR_BAD_ARG, "acfm_visible_verts: bad sizes");
  ACFM_REQUIRE(faces_batch_stride == 0 || faces_batch_stride == (int64_t)F * 3, ACFM_ERR_BAD_ARG, "acfm_visible_verts: faces_batch_stride must be 0 or F*3");
  if (N == 0 || V == 0) return ACFM_OK;
  ACFM_REQUIRE(vis, ACFM_ERR_BAD_ARG, "acfm_visible_verts: null output");
  ACFM_REQUIRE(N <= 65535, ACFM_ERR_UNSUPPORTED, "acfm_visible_verts: N=%d > 65535", N);
  cudaStream_t st = (cudaStream_t)stream;
  ACFM_CUDA_OK(cudaMemsetAsync(vis, 0, sizeof(float) * (size_t)N * V, st));
  if (F == 0 || HW == 0) return ACFM_OK;
  ACFM_REQUIRE(pix_to_face && faces, ACFM_ERR_BAD_ARG, "acfm_visible_verts: null input");
  const int chunks = max(1, min(64, HW / (kThreads * 4)));
  if (faces_i64)
    visible_verts_kernel<long long><<<dim3(chunks, N), kThreads, 0, st>>>((const long long*)pix_to_face, pix_stride, (const long long*)faces, faces_batch_stride, V, F, HW, vis);
  else
    visible_verts_kernel<int><<<dim3(chunks, N), kThreads, 0, st>>>((const long long*)pix_to_face, pix_stride, (const int*)faces, faces_batch_stride, V, F, HW, vis);
  ACFM_LAUNCH_OK("visible_verts_kernel");
  return ACFM_OK;
}

extern "C" int acfm_bds_loss_fwd(const float* verts, int vert_stride, const float* vis, const float* bds, const int64_t* sel,
                                 int N, int NB, int V, int P, int S, float* loss, int* argmin, void* stream) {
  ACFM_REQUIRE(N >= 0 && V >= 0 && P >= 0 && S >= 0 && (NB > 0 || N == 0) && vert_stride >= 2, ACFM_ERR_BAD_ARG, "acfm_bds_loss_fwd: bad sizes");
  if (N == 0) return ACFM_OK;
  ACFM_REQUIRE(N % NB == 0, ACFM_ERR_BAD_ARG, "acfm_bds_loss_fwd: N=%d is not a multiple of NB=%d", N, NB);
  ACFM_REQUIRE(loss, ACFM_ERR_BAD_ARG, "acfm_bds_loss_fwd: null output");
  ACFM_REQUIRE(N <= 65535, ACFM_ERR_UNSUPPORTED, "acfm_bds_loss_fwd: N=%d > 65535", N);
  cudaStream_t st = (cudaStream_t)stream;
  ACFM_CUDA_OK(cudaMemsetAsync(loss, 0, sizeof(float) * (size_t)N, st));
  if (S == 0) return ACFM_OK;
  ACFM_REQUIRE(verts && vis && bds, ACFM_ERR_BAD_ARG, "acfm_bds_loss_fwd: null input");
  const int smem = V * 12;
  ACFM_REQUIRE(smem <= 200 * 1024, ACFM_ERR_UNSUPPORTED, "acfm_bds_loss_fwd: V=%d too large", V);
  static std::atomic<int> smem_set[kAcfmMaxDevices];
  if (smem > 48 * 1024) ACFM_CUDA_OK(acfm_ensure_smem(bds_fwd_kernel, smem, smem_set));
  bds_fwd_kernel<<<dim3((S + kThreads - 1) / kThreads, N), kThreads, smem, st>>>(verts, vert_stride, vis, bds, (const long long*)sel, NB, V, P, S, loss, argmin);
  ACFM_LAUNCH_OK("bds_fwd_kernel");
  return ACFM_OK;
}

extern "C" int acfm_bds_loss_bwd(const float* verts, int vert_stride, const float* bds, const int64_t* sel, const int* argmin,
                                 const float* grad_loss, int N, int NB, int V, int P, int S, float* grad_verts, void* stream) {
  ACFM_REQUIRE(N >= 0 && V >= 0 && P >= 0 && S >= 0 && (NB > 0 || N == 0) && vert_stride >= 2, ACFM_ERR_BAD_ARG, "acfm_bds_loss_bwd: bad sizes");
  if (N == 0 || V == 0) return ACFM_OK;
  ACFM_REQUIRE(N % NB == 0, ACFM_ERR_BAD_ARG, "acfm_bds_loss_bwd: N=%d is not a multiple of NB=%d", N, NB);
  ACFM_REQUIRE(grad_verts, ACFM_ERR_BAD_ARG, "acfm_bds_loss_bwd: null output");
  ACFM_REQUIRE(N <= 65535, ACFM_ERR_UNSUPPORTED, "acfm_bds_loss_bwd: N=%d > 65535", N);
  cudaStream_t st = (cudaStream_t)stream;
  ACFM_CUDA_OK(cudaMemsetAsync(grad_verts, 0, sizeof(float) * (size_t)N * V * vert_stride, st));
  if (S == 0) return ACFM_OK;
  ACFM_REQUIRE(verts && bds && argmin && grad_loss, ACFM_ERR_BAD_ARG, "acfm_bds_loss_bwd: null input");
  bds_bwd_kernel<<<dim3((S + kThreads - 1) / kThreads, N), kThreads, 0, st>>>(verts, vert_stride, bds, (const long long*)sel, argmin, grad_loss, NB, V, P, S, grad_verts);
  ACFM_LAUNCH_OK("bds_bwd_kernel");
  return ACFM_OK;
}

extern "C" int acfm_of_loss_fwd(const float* proj, int proj_stride, const float* vis, const float* flows, int B, int NBf, int T,
                                int V, int H, int W, float* loss, float* of_pred, float* vis_out, float* samples, void* stream) {
  ACFM_REQUIRE(B >= 0 && T >= 1 && V >= 0 && H > 0 && W > 0 && proj_stride >= 2 && (NBf > 0 || B == 0), ACFM_ERR_BAD_ARG, "acfm_of_loss_fwd: bad sizes");
  if (B == 0 || T < 2) return ACFM_OK;
  ACFM_REQUIRE(B % NBf == 0, ACFM_ERR_BAD_ARG, "acfm_of_loss_fwd: B=%d is not a multiple of NBf=%d", B, NBf);
  ACFM_REQUIRE(proj && vis && flows && loss && vis_out && of_pred && samples, ACFM_ERR_BAD_ARG, "acfm_of_loss_fwd: null pointer");
  of_fwd_kernel<<<B * (T - 1), kThreads, 0, (cudaStream_t)stream>>>(proj, proj_stride, vis, flows, NBf, T, V, H, W, loss, of_pred, vis_out, samples);
  ACFM_LAUNCH_OK("of_fwd_kernel");
  return ACFM_OK;
}

extern "C" int acfm_of_loss_bwd(const float* vis_out, const float* of_pred, const float* samples, const float* grad_loss,
                                int proj_stride, int B, int T, int V, int H, int W, float* grad_proj, void* stream) {
  ACFM_REQUIRE(B >= 0 && T >= 1 && V >= 0 && H > 0 && W > 0 && proj_stride >= 2, ACFM_ERR_BAD_ARG, "acfm_of_loss_bwd: bad sizes");
  if (B == 0 || V == 0) return ACFM_OK;
  ACFM_REQUIRE(grad_proj, ACFM_ERR_BAD_ARG, "acfm_of_loss_bwd: null output");
  cudaStream_t st = (cudaStream_t)stream;
  ACFM_CUDA_OK(cudaMemsetAsync(grad_proj, 0, sizeof(float) * (size_t)B * T * V * proj_stride, st));
  if (T < 2) return ACFM_OK;
  ACFM_REQUIRE(vis_out && of_pred && samples && grad_loss, ACFM_ERR_BAD_ARG, "acfm_of_loss_bwd: null input");
  of_bwd_kernel<<<B * (T - 1), kThreads, 0, st>>>(vis_out, of_pred, samples, grad_loss, proj_stride, T, V, H, W, grad_proj);
  ACFM_LAUNCH_OK("of_bwd_kernel");
  return ACFM_OK;
}

extern "C" int acfm_kp_loss_fwd(const float* kp_pred, int pred_stride, const float* kp_gt, int N, int NB, int Kp, float* loss,
                                void* stream) {
  ACFM_REQUIRE(N >= 0 && Kp > 0 && pred_stride >= 2 && (NB > 0 || N == 0), ACFM_ERR_BAD_ARG, "acfm_kp_loss_fwd: bad sizes");
  if (N == 0) return ACFM_OK;
  ACFM_REQUIRE(N % NB == 0, ACFM_ERR_BAD_ARG, "acfm_kp_loss_fwd: N=%d is not a multiple of NB=%d", N, NB);
  ACFM_REQUIRE(kp_pred && kp_gt && loss, ACFM_ERR_BAD_ARG, "acfm_kp_loss_fwd: null pointer");
  kp_fwd_kernel<<<(N + kThreads / 32 - 1) / (kThreads / 32), kThreads, 0, (cudaStream_t)stream>>>(kp_pred, pred_stride, kp_gt, N, NB, Kp, loss);
  ACFM_LAUNCH_OK("kp_fwd_kernel");
  return ACFM_OK;
}

extern "C" int acfm_kp_loss_bwd(const float* kp_pred, int pred_stride, const float* kp_gt, const float* grad_loss, int N, int NB,
                                int Kp, float* grad_kp_pred, void* stream) {
  ACFM_REQUIRE(N >= 0 && Kp > 0 && pred_stride >= 2 && (NB > 0 || N == 0), ACFM_ERR_BAD_ARG, "acfm_kp_loss_bwd: bad sizes");
  if (N == 0) return ACFM_OK;
  ACFM_REQUIRE(N % NB == 0, ACFM_ERR_BAD_ARG, "acfm_kp_loss_bwd: N=%d is not a multiple of NB=%d", N, NB);
  ACFM_REQUIRE(kp_pred && kp_gt && grad_loss && grad_kp_pred, ACFM_ERR_BAD_ARG, "acfm_kp_loss_bwd: null pointer");
  kp_bwd_kernel<<<(N + kThreads / 32 - 1) / (kThreads / 32), kThreads, 0, (cudaStream_t)stream>>>(kp_pred, pred_stride, kp_gt, grad_loss, N, NB, Kp, grad_kp_pred);
  ACFM_LAUNCH_OK("kp_bwd_kernel");
  return ACFM_OK;
}

extern "C" int acfm_hypothesis_weight_fwd(const float* loss, int G, int M, float* probs, float* total, void* stream) {
  ACFM_REQUIRE(G > 0 && M > 0, ACFM_ERR_BAD_ARG, "acfm_hypothesis_weight_fwd: bad sizes");
  ACFM_REQUIRE(loss && probs && total, ACFM_ERR_BAD_ARG, "acfm_hypothesis_weight_fwd: null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  ACFM_CUDA_OK(cudaMemsetAsync(total, 0, sizeof(float), st));
  hyp_fwd_kernel<<<max(1, min(148, (M + kThreads - 1) / kThreads)), kThreads, 0, st>>>(loss, G, M, probs, total);
  ACFM_LAUNCH_OK("hyp_fwd_kernel");
  return ACFM_OK;
}

extern "C" int acfm_hypothesis_weight_bwd(const float* probs, const float* grad_total, int G, int M, float* grad_loss,
                                          void* stream) {
  ACFM_REQUIRE(G > 0 && M > 0, ACFM_ERR_BAD_ARG, "acfm_hypothesis_weight_bwd: bad sizes");
  ACFM_REQUIRE(probs && grad_total && grad_loss, ACFM_ERR_BAD_ARG, "acfm_hypothesis_weight_bwd: null pointer");
  const int GM = G * M;
  hyp_bwd_kernel<<<max(1, min(148, (GM + kThreads - 1) / kThreads)), kThreads, 0, (cudaStream_t)stream>>>(probs, grad_total, GM, 1.0f / (float)M, grad_loss);
  ACFM_LAUNCH_OK("hyp_bwd_kernel");
  return ACFM_OK;
}
