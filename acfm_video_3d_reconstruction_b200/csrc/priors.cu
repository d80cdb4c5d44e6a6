// priors.cu — shape priors on the deformed meshes: cotangent-Laplacian smoothing and local rigidity.
//
// Replaces pytorch3d.loss.mesh_laplacian_smoothing(mesh_3d, method="cot") (PyTorch3D 0.3.0; call site
// /root/reference/multiframe/main.py:699-704; its laplacian_cot is the function the reference also carries in
// nnutils/geom_utils.py:258-325) and loss_utils.locally_rigid_fn (/root/reference/multiframe/nnutils/loss_utils.py:150-164,
// call site main.py:714).  The reference evaluates both on the G-fold repeated batch through a packed sparse (sum V x sum V)
// Laplacian, a sparse matmul and gathers over `edges_packed`; here: one scatter pass over the faces and one pass over the
// vertices (smoothing), one pass over the edges (rigidity), per mesh, forward and backward.
//
// Smoothing, per mesh (weights are constants, as under the reference's no_grad):
//   w_ij = sum over faces containing edge (i,j) of cot(opposite angle)/4 ; s_i = sum_j w_ij ; nw_i = s_i > 0 ? 1/s_i : s_i
//   l_i = nw_i * sum_j w_ij v_j - v_i ;  loss = (1/N) sum_n (1/V) sum_i |l_i|
// Rigidity: loss = (1/N) sum_n sum_edges (|v_a - v_b| - |t_a - t_b|)^2.
#include "common.cuh"

namespace {

constexpr int kThreads = 256;

// cot/4 of the three angles of face f of mesh n: (cota, cotb, cotc) opposite v0, v1, v2 (geom_utils.py:281-301)
__device__ __forceinline__ void face_cots(const float* p0, const float* p1, const float* p2, float& cota, float& cotb, float& cotc) {
  const float ax = p1[0] - p2[0], ay = p1[1] - p2[1], az = p1[2] - p2[2];
  const float bx = p0[0] - p2[0], by = p0[1] - p2[1], bz = p0[2] - p2[2];
  const float cx = p0[0] - p1[0], cy = p0[1] - p1[1], cz = p0[2] - p1[2];
  const float A = sqrtf(ax * ax + ay * ay + az * az), B = sqrtf(bx * bx + by * by + bz * bz), C = sqrtf(cx * cx + cy * cy + cz * cz);
  const float s = 0.5f * (A + B + C);
  const float area = sqrtf(fmaxf(s * (s - A) * (s - B) * (s - C), 1e-12f));
  const float A2 = A * A, B2 = B * B, C2 = C * C;
  cota = (B2 + C2 - A2) / area / 4.0f;
  cotb = (A2 + C2 - B2) / area / 4.0f;
  cotc = (A2 + B2 - C2) / area / 4.0f;
}

// acc (N,V,4): xyz = sum_j w_ij v_j, w = sum_j w_ij.  grid (ceil(F/kThreads), N)
template <typename IdxT>
__global__ void __launch_bounds__(kThreads) lap_smooth_scatter_kernel(const float* __restrict__ verts, const IdxT* __restrict__ faces,
                                                                      long long faces_stride, int V, int F, float* __restrict__ acc) {
  const int f = blockIdx.x * kThreads + threadIdx.x, n = blockIdx.y;
  if (f >= F) return;
  const IdxT* t = faces + (size_t)n * faces_stride + (size_t)f * 3;
  const int i0 = (int)t[0], i1 = (int)t[1], i2 = (int)t[2];
  const float* vb = verts + (size_t)n * V * 3;
  const float *p0 = vb + (size_t)i0 * 3, *p1 = vb + (size_t)i1 * 3, *p2 = vb + (size_t)i2 * 3;
  float ca, cb, cc;
  face_cots(p0, p1, p2, ca, cb, cc);
  float* a = acc + (size_t)n * V * 4;
  // edge (v1,v2) weight ca, edge (v2,v0) weight cb, edge (v0,v1) weight cc; symmetric
#define ACFM_EDGE(I, J, PI, PJ, W)                                                                                         \
  atomicAdd(a + (size_t)(I) * 4, (W) * (PJ)[0]); atomicAdd(a + (size_t)(I) * 4 + 1, (W) * (PJ)[1]);                        \
  atomicAdd(a + (size_t)(I) * 4 + 2, (W) * (PJ)[2]); atomicAdd(a + (size_t)(I) * 4 + 3, (W));                              \
  atomicAdd(a + (size_t)(J) * 4, (W) * (PI)[0]); atomicAdd(a + (size_t)(J) * 4 + 1, (W) * (PI)[1]);                        \
  atomicAdd(a + (size_t)(J) * 4 + 2, (W) * (PI)[2]); atomicAdd(a + (size_t)(J) * 4 + 3, (W));
  ACFM_EDGE(i1, i2, p1, p2, ca)
  ACFM_EDGE(i2, i0, p2, p0, cb)
  ACFM_EDGE(i0, i1, p0, p1, cc)
#undef ACFM_EDGE
}

// per vertex: l = nw * acc.xyz - v ; loss[n] += |l| / V ; unit (N,V,4) = (l/|l|, nw) for the backward.  grid (chunks, N)
__global__ void __launch_bounds__(kThreads) lap_smooth_verts_kernel(const float* __restrict__ verts, const float* __restrict__ acc, int V,
                                                                    float* __restrict__ loss, float* __restrict__ unit) {
  const int n = blockIdx.y;
  float s = 0.0f;
  for (int i = blockIdx.x * kThreads + threadIdx.x; i < V; i += gridDim.x * kThreads) {
    const float4 a = reinterpret_cast<const float4*>(acc)[(size_t)n * V + i];
    const float* v = verts + ((size_t)n * V + i) * 3;
    const float nw = a.w > 0.0f ? 1.0f / a.w : a.w;
    const float lx = a.x * nw - v[0], ly = a.y * nw - v[1], lz = a.z * nw - v[2];
    const float len = sqrtf(lx * lx + ly * ly + lz * lz);
    s += len;
    const float inv = len > 0.0f ? 1.0f / len : 0.0f;
    if (unit) reinterpret_cast<float4*>(unit)[(size_t)n * V + i] = make_float4(lx * inv, ly * inv, lz * inv, nw);
  }
  __shared__ float red[kThreads / 32];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float a = 0.0f;
    for (int w = 0; w < kThreads / 32; ++w) a += red[w];
    atomicAdd(loss + n, a / (float)V);
  }
}

// grad_verts must hold the "- v_i" term already (written by lap_smooth_bwd_self_kernel); this adds the neighbour terms:
// for edge (i,j) weight w: grad_j += w nw_i g_i ; grad_i += w nw_j g_j, with g_i = grad_loss[n]/V * unit_i
template <typename IdxT>
__global__ void __launch_bounds__(kThreads) lap_smooth_bwd_kernel(const float* __restrict__ verts, const IdxT* __restrict__ faces,
                                                                  long long faces_stride, const float* __restrict__ unit,
                                                                  const float* __restrict__ grad_loss, int V, int F,
                                                                  float* __restrict__ grad_verts) {
  const int f = blockIdx.x * kThreads + threadIdx.x, n = blockIdx.y;
  if (f >= F) return;
  const IdxT* t = faces + (size_t)n * faces_stride + (size_t)f * 3;
  const int i0 = (int)t[0], i1 = (int)t[1], i2 = (int)t[2];
  const float* vb = verts + (size_t)n * V * 3;
  float ca, cb, cc;
  face_cots(vb + (size_t)i0 * 3, vb + (size_t)i1 * 3, vb + (size_t)i2 * 3, ca, cb, cc);
  const float gs = grad_loss[n] / (float)V;
  const float4* u = reinterpret_cast<const float4*>(unit) + (size_t)n * V;
  float* g = grad_verts + (size_t)n * V * 3;
  const float4 u0 = u[i0], u1 = u[i1], u2 = u[i2];
#define ACFM_EDGE(I, J, UI, UJ, W)                                                                                          \
  {                                                                                                                         \
    const float wi = (W) * (UI).w * gs, wj = (W) * (UJ).w * gs;                                                             \
    atomicAdd(g + (size_t)(J) * 3, wi * (UI).x); atomicAdd(g + (size_t)(J) * 3 + 1, wi * (UI).y); atomicAdd(g + (size_t)(J) * 3 + 2, wi * (UI).z); \
    atomicAdd(g + (size_t)(I) * 3, wj * (UJ).x); atomicAdd(g + (size_t)(I) * 3 + 1, wj * (UJ).y); atomicAdd(g + (size_t)(I) * 3 + 2, wj * (UJ).z); \
  }
  ACFM_EDGE(i1, i2, u1, u2, ca)
  ACFM_EDGE(i2, i0, u2, u0, cb)
  ACFM_EDGE(i0, i1, u0, u1, cc)
#undef ACFM_EDGE
}

__global__ void __launch_bounds__(kThreads) lap_smooth_bwd_self_kernel(const float* __restrict__ unit, const float* __restrict__ grad_loss,
                                                                       int V, float* __restrict__ grad_verts) {
  const int n = blockIdx.y;
  const float gs = grad_loss[n] / (float)V;
  for (int i = blockIdx.x * kThreads + threadIdx.x; i < V; i += gridDim.x * kThreads) {
    const float4 u = reinterpret_cast<const float4*>(unit)[(size_t)n * V + i];
    float* g = grad_verts + ((size_t)n * V + i) * 3;
    g[0] = -gs * u.x; g[1] = -gs * u.y; g[2] = -gs * u.z;
  }
}

// rigidity.  grid (chunks, N); loss[n] = sum_e (|v_a - v_b| - |t_a - t_b|)^2
template <typename IdxT>
__global__ void __launch_bounds__(kThreads) rigid_fwd_kernel(const float* __restrict__ verts, const float* __restrict__ tmpl, int NT,
                                                             const IdxT* __restrict__ edges, int V, int E, float* __restrict__ loss) {
  const int n = blockIdx.y;
  const float* v = verts + (size_t)n * V * 3;
  const float* t = tmpl + (size_t)(n % NT) * V * 3;
  float s = 0.0f;
  for (int e = blockIdx.x * kThreads + threadIdx.x; e < E; e += gridDim.x * kThreads) {
    const int a = (int)edges[(size_t)e * 2], b = (int)edges[(size_t)e * 2 + 1];
    const float dx = v[a * 3] - v[b * 3], dy = v[a * 3 + 1] - v[b * 3 + 1], dz = v[a * 3 + 2] - v[b * 3 + 2];
    const float tx = t[a * 3] - t[b * 3], ty = t[a * 3 + 1] - t[b * 3 + 1], tz = t[a * 3 + 2] - t[b * 3 + 2];
    const float d = sqrtf(dx * dx + dy * dy + dz * dz) - sqrtf(tx * tx + ty * ty + tz * tz);
    s += d * d;
  }
  __shared__ float red[kThreads / 32];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float a = 0.0f;
    for (int w = 0; w < kThreads / 32; ++w) a += red[w];
    atomicAdd(loss + n, a);
  }
}

// grad_verts (N,V,3) and, if not null, grad_tmpl (NT,V,3), both zeroed by the launcher
template <typename IdxT>
__global__ void __launch_bounds__(kThreads) rigid_bwd_kernel(const float* __restrict__ verts, const float* __restrict__ tmpl, int NT,
                                                             const IdxT* __restrict__ edges, const float* __restrict__ grad_loss, int V,
                                                             int E, float* __restrict__ grad_verts, float* __restrict__ grad_tmpl) {
  const int n = blockIdx.y;
  const float* v = verts + (size_t)n * V * 3;
  const float* t = tmpl + (size_t)(n % NT) * V * 3;
  float* gv = grad_verts + (size_t)n * V * 3;
  float* gt = grad_tmpl ? grad_tmpl + (size_t)(n % NT) * V * 3 : nullptr;
  const float g = grad_loss[n];
  for (int e = blockIdx.x * kThreads + threadIdx.x; e < E; e += gridDim.x * kThreads) {
    const int a = (int)edges[(size_t)e * 2], b = (int)edges[(size_t)e * 2 + 1];
    const float dx = v[a * 3] - v[b * 3], dy = v[a * 3 + 1] - v[b * 3 + 1], dz = v[a * 3 + 2] - v[b * 3 + 2];
    const float tx = t[a * 3] - t[b * 3], ty = t[a * 3 + 1] - t[b * 3 + 1], tz = t[a * 3 + 2] - t[b * 3 + 2];
    const float ld = sqrtf(dx * dx + dy * dy + dz * dz), lt = sqrtf(tx * tx + ty * ty + tz * tz);
    const float c = 2.0f * (ld - lt) * g;
    const float cv = ld > 0.0f ? c / ld : 0.0f;
    atomicAdd(gv + a * 3, cv * dx); atomicAdd(gv + a * 3 + 1, cv * dy); atomicAdd(gv + a * 3 + 2, cv * dz);
    atomicAdd(gv + b * 3, -cv * dx); atomicAdd(gv + b * 3 + 1, -cv * dy); atomicAdd(gv + b * 3 + 2, -cv * dz);
    if (gt) {
      const float ct = lt > 0.0f ? -c / lt : 0.0f;
      atomicAdd(gt + a * 3, ct * tx); atomicAdd(gt + a * 3 + 1, ct * ty); atomicAdd(gt + a * 3 + 2, ct * tz);
      atomicAdd(gt + b * 3, -ct * tx); atomicAdd(gt + b * 3 + 1, -ct * ty); atomicAdd(gt + b * 3 + 2, -ct * tz);
    }
  }
}

}  // namespace

extern "C" int acfm_laplacian_smoothing_fwd(const float* verts, const void* faces, int faces_i64, int64_t faces_batch_stride, int N,
                                            int V, int F, float* loss, float* unit, float* workspace, void* stream) {
  ACFM_REQUIRE(N >= 0 && V >= 0 && F >= 0, ACFM_ERR_BAD_ARG, "acfm_laplacian_smoothing_fwd: bad sizes");
  ACFM_REQUIRE(faces_batch_stride == 0 || faces_batch_stride == (int64_t)F * 3, ACFM_ERR_BAD_ARG, "acfm_laplacian_smoothing_fwd: faces_batch_stride must be 0 or F*3");
  if (N == 0) return ACFM_OK;
  ACFM_REQUIRE(loss && workspace && (V == 0 || verts) && (F == 0 || faces), ACFM_ERR_BAD_ARG, "acfm_laplacian_smoothing_fwd: null pointer (workspace: N*V*4 floats)");
  ACFM_REQUIRE(N <= 65535, ACFM_ERR_UNSUPPORTED, "acfm_laplacian_smoothing_fwd: N=%d > 65535", N);
  cudaStream_t st = (cudaStream_t)stream;
  ACFM_CUDA_OK(cudaMemsetAsync(loss, 0, sizeof(float) * (size_t)N, st));
  if (V == 0) return ACFM_OK;
  ACFM_CUDA_OK(cudaMemsetAsync(workspace, 0, sizeof(float) * 4 * (size_t)N * V, st));
  if (F > 0) {
    const dim3 grid((F + kThreads - 1) / kThreads, N);
    if (faces_i64) lap_smooth_scatter_kernel<long long><<<grid, kThreads, 0, st>>>(verts, (const long long*)faces, faces_batch_stride, V, F, workspace);
    else lap_smooth_scatter_kernel<int><<<grid, kThreads, 0, st>>>(verts, (const int*)faces, faces_batch_stride, V, F, workspace);
    ACFM_LAUNCH_OK("lap_smooth_scatter_kernel");
  }
  lap_smooth_verts_kernel<<<dim3(max(1, min(8, (V + kThreads - 1) / kThreads)), N), kThreads, 0, st>>>(verts, workspace, V, loss, unit);
  ACFM_LAUNCH_OK("lap_smooth_verts_kernel");
  return ACFM_OK;
}

extern "C" int acfm_laplacian_smoothing_bwd(const float* verts, const void* faces, int faces_i64, int64_t faces_batch_stride,
                                            const float* unit, const float* grad_loss, int N, int V, int F, float* grad_verts,
                                            void* stream) {
  ACFM_REQUIRE(N >= 0 && V >= 0 && F >= 0, ACFM_ERR_BAD_ARG, "acfm_laplacian_smoothing_bwd: bad sizes");
  ACFM_REQUIRE(faces_batch_stride == 0 || faces_batch_stride == (int64_t)F * 3, ACFM_ERR_BAD_ARG, "acfm_laplacian_smoothing_bwd: faces_batch_stride must be 0 or F*3");
  if (N == 0 || V == 0) return ACFM_OK;
  ACFM_REQUIRE(verts && unit && grad_loss && grad_verts && (F == 0 || faces), ACFM_ERR_BAD_ARG, "acfm_laplacian_smoothing_bwd: null pointer");
  ACFM_REQUIRE(N <= 65535, ACFM_ERR_UNSUPPORTED, "acfm_laplacian_smoothing_bwd: N=%d > 65535", N);
  cudaStream_t st = (cudaStream_t)stream;
  lap_smooth_bwd_self_kernel<<<dim3(max(1, min(8, (V + kThreads - 1) / kThreads)), N), kThreads, 0, st>>>(unit, grad_loss, V, grad_verts);
  ACFM_LAUNCH_OK("lap_smooth_bwd_self_kernel");
  if (F > 0) {
    const dim3 grid((F + kThreads - 1) / kThreads, N);
    if (faces_i64) lap_smooth_bwd_kernel<long long><<<grid, kThreads, 0, st>>>(verts, (const long long*)faces, faces_batch_stride, unit, grad_loss, V, F, grad_verts);
    else lap_smooth_bwd_kernel<int><<<grid, kThreads, 0, st>>>(verts, (const int*)faces, faces_batch_stride, unit, grad_loss, V, F, grad_verts);
    ACFM_LAUNCH_OK("lap_smooth_bwd_kernel");
  }
  return ACFM_OK;
}

extern "C" int acfm_edge_rigidity_fwd(const float* verts, const float* tmpl, const void* edges, int edges_i64, int N, int NT, int V,
                                      int E, float* loss, void* stream) {
  ACFM_REQUIRE(N >= 0 && V >= 0 && E >= 0 && (NT > 0 || N == 0), ACFM_ERR_BAD_ARG, "acfm_edge_rigidity_fwd: bad sizes");
  if (N == 0) return ACFM_OK;
  ACFM_REQUIRE(N % NT == 0, ACFM_ERR_BAD_ARG, "acfm_edge_rigidity_fwd: N=%d is not a multiple of NT=%d", N, NT);
  ACFM_REQUIRE(loss, ACFM_ERR_BAD_ARG, "acfm_edge_rigidity_fwd: null output");
  ACFM_REQUIRE(N <= 65535, ACFM_ERR_UNSUPPORTED, "acfm_edge_rigidity_fwd: N=%d > 65535", N);
  cudaStream_t st = (cudaStream_t)stream;
  ACFM_CUDA_OK(cudaMemsetAsync(loss, 0, sizeof(float) * (size_t)N, st));
  if (E == 0) return ACFM_OK;
  ACFM_REQUIRE(verts && tmpl && edges, ACFM_ERR_BAD_ARG, "acfm_edge_rigidity_fwd: null input");
  const dim3 grid(max(1, min(8, (E + kThreads - 1) / kThreads)), N);
  if (edges_i64) rigid_fwd_kernel<long long><<<grid, kThreads, 0, st>>>(verts, tmpl, NT, (const long long*)edges, V, E, loss);
  else rigid_fwd_kernel<int><<<grid, kThreads, 0, st>>>(verts, tmpl, NT, (const int*)edges, V, E, loss);
  ACFM_LAUNCH_OK("rigid_fwd_kernel");
  return ACFM_OK;
}

extern "C" int acfm_edge_rigidity_bwd(const float* verts, const float* tmpl, const void* edges, int edges_i64, const float* grad_loss,
                                      int N, int NT, int V, int E, float* grad_verts, float* grad_tmpl, void* stream) {
  ACFM_REQUIRE(N >= 0 && V >= 0 && E >= 0 && (NT > 0 || N == 0), ACFM_ERR_BAD_ARG, "acfm_edge_rigidity_bwd: bad sizes");
  if (N == 0 || V == 0) return ACFM_OK;
  ACFM_REQUIRE(N % NT == 0, ACFM_ERR_BAD_ARG, "acfm_edge_rigidity_bwd: N=%d is not a multiple of NT=%d", N, NT);
  ACFM_REQUIRE(grad_verts, ACFM_ERR_BAD_ARG, "acfm_edge_rigidity_bwd: null output");
  ACFM_REQUIRE(N <= 65535, ACFM_ERR_UNSUPPORTED, "acfm_edge_rigidity_bwd: N=%d > 65535", N);
  cudaStream_t st = (cudaStream_t)stream;
  ACFM_CUDA_OK(cudaMemsetAsync(grad_verts, 0, sizeof(float) * 3 * (size_t)N * V, st));
  if (grad_tmpl) ACFM_CUDA_OK(cudaMemsetAsync(grad_tmpl, 0, sizeof(float) * 3 * (size_t)NT * V, st));
  if (E == 0) return ACFM_OK;
  ACFM_REQUIRE(verts && tmpl && edges && grad_loss, ACFM_ERR_BAD_ARG, "acfm_edge_rigidity_bwd: null input");
  const dim3 grid(max(1, min(8, (E + kThreads - 1) / kThreads)), N);
  if (edges_i64) rigid_bwd_kernel<long long><<<grid, kThreads, 0, st>>>(verts, tmpl, NT, (const long long*)edges, grad_loss, V, E, grad_verts, grad_tmpl);
  else rigid_bwd_kernel<int><<<grid, kThreads, 0, st>>>(verts, tmpl, NT, (const int*)edges, grad_loss, V, E, grad_verts, grad_tmpl);
  ACFM_LAUNCH_OK("rigid_bwd_kernel");
  return ACFM_OK;
}
