// laplacian.cu — dense mesh Laplacian of the template, uniform or cotangent weights.
//
// Replaces geom_utils.mesh_laplacian / laplacian_cot
// (/root/reference/multiframe/nnutils/geom_utils.py:159-255,258-325) and, for method 'uniform',
// PyTorch3D's Meshes.laplacian_packed().  The multiframe trainer rebuilds the cotangent
// Laplacian every forward (multiframe/main.py:600-601) through a sparse tensor, a transpose-add, a sparse
// row sum and two densifications; here it is one scatter pass over the faces into the dense (V,V) matrix
// the deformation solve needs, plus one row pass for the diagonal.
//   uniform: L[i,j] = 1/deg(i) for every edge (i,j), L[i,i] = -1
//   cot:     L[i,j] = sum over the faces containing edge (i,j) of cot(angle opposite the edge)/4 computed as
//            (b^2 + c^2 - a^2) / area / 4 with Heron's area clamped at sqrt(1e-12);  L[i,i] = -sum_j L[i,j]
#include "common.cuh"

namespace {

constexpr int kThreads = 256;

template <typename IdxT>
__global__ void __launch_bounds__(kThreads) lap_scatter_kernel(const float* __restrict__ verts, const IdxT* __restrict__ faces,
                                                               int V, int F, int method, float* __restrict__ L) {
  const int f = blockIdx.x * kThreads + threadIdx.x;
  if (f >= F) return;
  const int i0 = (int)faces[f * 3], i1 = (int)faces[f * 3 + 1], i2 = (int)faces[f * 3 + 2];
  if (method == 0) {  // adjacency marks; the row pass turns them into 1/deg
    L[(size_t)i0 * V + i1] = 1.0f; L[(size_t)i1 * V + i0] = 1.0f;
    L[(size_t)i1 * V + i2] = 1.0f; L[(size_t)i2 * V + i1] = 1.0f;
    L[(size_t)i2 * V + i0] = 1.0f; L[(size_t)i0 * V + i2] = 1.0f;
    return;
  }
  const float* p0 = verts + (size_t)i0 * 3;
  const float* p1 = verts + (size_t)i1 * 3;
  const float* p2 = verts + (size_t)i2 * 3;
  // side lengths: A opposite v0, B opposite v1, C opposite v2 (geom_utils.py:281-285)
  const float ax = p1[0] - p2[0], ay = p1[1] - p2[1], az = p1[2] - p2[2];
  const float bx = p0[0] - p2[0], by = p0[1] - p2[1], bz = p0[2] - p2[2];
  const float cx = p0[0] - p1[0], cy = p0[1] - p1[1], cz = p0[2] - p1[2];
  const float A = sqrtf(ax * ax + ay * ay + az * az), B = sqrtf(bx * bx + by * by + bz * bz), C = sqrtf(cx * cx + cy * cy + cz * cz);
  const float s = 0.5f * (A + B + C);
  const float area = sqrtf(fmaxf(s * (s - A) * (s - B) * (s - C), 1e-12f));
  const float A2 = A * A, B2 = B * B, C2 = C * C;
  const float cota = (B2 + C2 - A2) / area / 4.0f, cotb = (A2 + C2 - B2) / area / 4.0f, cotc = (A2 + B2 - C2) / area / 4.0f;
  // L[v1,v2] += cota, L[v2,v0] += cotb, L[v0,v1] += cotc, and the transposes (geom_utils.py:303-318)
  atomicAdd(L + (size_t)i1 * V + i2, cota); atomicAdd(L + (size_t)i2 * V + i1, cota);
  atomicAdd(L + (size_t)i2 * V + i0, cotb); atomicAdd(L + (size_t)i0 * V + i2, cotb);
  atomicAdd(L + (size_t)i0 * V + i1, cotc); atomicAdd(L + (size_t)i1 * V + i0, cotc);
}

// one warp per row: diagonal / normalisation
__global__ void __launch_bounds__(kThreads) lap_rows_kernel(int V, int method, float* __restrict__ L) {
  const int i = blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5);
  if (i >= V) return;
  const int lane = threadIdx.x & 31;
  float* row = L + (size_t)i * V;
  float s = 0.0f;
  for (int j = lane; j < V; j += 32) if (j != i) s += row[j];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (method == 0) {
    const float w = s > 0.0f ? 1.0f / s : 0.0f;
    for (int j = lane; j < V; j += 32) if (j != i && row[j] != 0.0f) row[j] = w;
    if (lane == 0) row[i] = -1.0f;
  } else {
    if (lane == 0) row[i] = -s;
  }
}

}  // namespace

extern "C" int acfm_laplacian_fwd(const float* verts, const void* faces, int faces_i64, int V, int F, int method, float* L,
                                  void* stream) {
  ACFM_REQUIRE(V >= 0 && F >= 0 && (method == 0 || method == 1), ACFM_ERR_BAD_ARG, "acfm_laplacian_fwd: bad arguments");
  if (V == 0) return ACFM_OK;
  ACFM_REQUIRE(L && (F == 0 || (verts && faces)), ACFM_ERR_BAD_ARG, "acfm_laplacian_fwd: null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  ACFM_CUDA_OK(cudaMemsetAsync(L, 0, sizeof(float) * (size_t)V * V, st));
  if (F > 0) {
    if (faces_i64) lap_scatter_kernel<long long><<<(F + kThreads - 1) / kThreads, kThreads, 0, st>>>(verts, (const long long*)faces, V, F, method, L);
    else lap_scatter_kernel<int><<<(F + kThreads - 1) / kThreads, kThreads, 0, st>>>(verts, (const int*)faces, V, F, method, L);
    ACFM_LAUNCH_OK("lap_scatter_kernel");
  }
  lap_rows_kernel<<<(V + kThreads / 32 - 1) / (kThreads / 32), kThreads, 0, st>>>(V, method, L);
  ACFM_LAUNCH_OK("lap_rows_kernel");
  return ACFM_OK;
}
