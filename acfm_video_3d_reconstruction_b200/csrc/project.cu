// project.cu — weak-perspective camera-multiplex projection, forward and backward.
//
// Replaces geom_utils.orthographic_proj_withz / quat_rotate / hamilton_product
// (/root/reference/multiframe/nnutils/geom_utils.py:62-79,107-153) and the y-flip / view
// transform of NeuralRenderer.forward (/root/reference/multiframe/nnutils/nmr.py:144-149).
// ~40 elementwise torch kernels and their (N,V,4) temporaries in the reference; here one
// vertex-major, coalesced pass.  HBM-bound: 12 B in (L2-resident across hypotheses) + 12 B out
// per vertex.
#include "common.cuh"

namespace {

constexpr int kThreads = 256;

__device__ __forceinline__ void hamilton_strict(const float* a, const float* b, float* o) {
  // operand order of geom_utils.hamilton_product (geom_utils.py:104-128), one rounding per op
  o[0] = fsub(fsub(fsub(fmul(a[0], b[0]), fmul(a[1], b[1])), fmul(a[2], b[2])), fmul(a[3], b[3]));
  o[1] = fsub(fadd(fadd(fmul(a[0], b[1]), fmul(a[1], b[0])), fmul(a[2], b[3])), fmul(a[3], b[2]));
  o[2] = fadd(fadd(fsub(fmul(a[0], b[2]), fmul(a[1], b[3])), fmul(a[2], b[0])), fmul(a[3], b[1]));
  o[3] = fadd(fsub(fadd(fmul(a[0], b[3]), fmul(a[1], b[2])), fmul(a[2], b[1])), fmul(a[3], b[0]));
}

// grid: (ceil(V / kThreads), N)
__global__ void __launch_bounds__(kThreads) project_fwd_kernel(const float* __restrict__ verts,
                                                               const float* __restrict__ cams, int NB, int V,
                                                               float offset_z, float sx, float sy, float z_add,
                                                               float* __restrict__ out) {
  const int n = blockIdx.y;
  const int v = blockIdx.x * kThreads + threadIdx.x;
  __shared__ float c[7];
  if (threadIdx.x < 7) c[threadIdx.x] = cams[(size_t)n * 7 + threadIdx.x];
  __syncthreads();
  if (v >= V) return;
  const float* x = verts + ((size_t)(n % NB) * V + v) * 3;
  const float x0 = x[0], x1 = x[1], x2 = x[2];
  const float q[4] = {c[3], c[4], c[5], c[6]};
  const float qc[4] = {q[0], fmul(-1.0f, q[1]), fmul(-1.0f, q[2]), fmul(-1.0f, q[3])};
  const float xq[4] = {fmul(x0, 0.0f), x0, x1, x2};
  float t[4], r[4];
  hamilton_strict(xq, qc, t);
  hamilton_strict(q, t, r);
  const float px = fadd(fmul(c[0], r[1]), c[1]);
  const float py = fadd(fmul(c[0], r[2]), c[2]);
  const float pz = fadd(fmul(c[0], r[3]), offset_z);
  float* o = out + ((size_t)n * V + v) * 3;
  o[0] = fmul(sx, px);
  o[1] = fmul(sy, py);
  o[2] = (z_add != 0.0f) ? fadd(pz, z_add) : pz;
}

__device__ __forceinline__ void hamilton_bwd_a(const float* g, const float* b, float* ga) {
  ga[0] += g[0] * b[0] + g[1] * b[1] + g[2] * b[2] + g[3] * b[3];
  ga[1] += -g[0] * b[1] + g[1] * b[0] - g[2] * b[3] + g[3] * b[2];
  ga[2] += -g[0] * b[2] + g[1] * b[3] + g[2] * b[0] - g[3] * b[1];
  ga[3] += -g[0] * b[3] - g[1] * b[2] + g[2] * b[1] + g[3] * b[0];
}
__device__ __forceinline__ void hamilton_bwd_b(const float* g, const float* a, float* gb) {
  gb[0] += g[0] * a[0] + g[1] * a[1] + g[2] * a[2] + g[3] * a[3];
  gb[1] += -g[0] * a[1] + g[1] * a[0] + g[2] * a[3] - g[3] * a[2];
  gb[2] += -g[0] * a[2] - g[1] * a[3] + g[2] * a[0] + g[3] * a[1];
  gb[3] += -g[0] * a[3] + g[1] * a[2] - g[2] * a[1] + g[3] * a[0];
}

// grid: (chunks = ceil(V / kThreads), NB).  Each CTA walks the G = N/NB renders that share mesh b,
// so grad_verts needs no atomics; the 7 camera gradients are reduced per render with warp
// shuffles and (only when chunks > 1) one atomicAdd per CTA per component.
__global__ void __launch_bounds__(kThreads) project_bwd_kernel(const float* __restrict__ verts,
                                                               const float* __restrict__ cams,
                                                               const float* __restrict__ grad_out, int N, int NB,
                                                               int V, float sx, float sy,
                                                               float* __restrict__ grad_verts,
                                                               float* __restrict__ grad_cams) {
  const int b = blockIdx.y;
  const int v = blockIdx.x * kThreads + threadIdx.x;
  const int G = N / NB;
  const bool live = v < V;
  __shared__ float red[kThreads / 32][7];
  float x0 = 0.f, x1 = 0.f, x2 = 0.f;
  if (live) {
    const float* x = verts + ((size_t)b * V + v) * 3;
    x0 = x[0]; x1 = x[1]; x2 = x[2];
  }
  float gx[3] = {0.f, 0.f, 0.f};
  for (int g = 0; g < G; ++g) {
    const int n = g * NB + b;
    const float* c = cams + (size_t)n * 7;
    const float s = c[0];
    const float q[4] = {c[3], c[4], c[5], c[6]};
    float gc[7] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (live) {
      const float qc[4] = {q[0], -q[1], -q[2], -q[3]};
      const float xq[4] = {0.f, x0, x1, x2};
      float t[4], r[4];
      hamilton_strict(xq, qc, t);
      hamilton_strict(q, t, r);
      const float* go = grad_out + ((size_t)n * V + v) * 3;
      const float gp[3] = {sx * go[0], sy * go[1], go[2]};
      gc[0] = gp[0] * r[1] + gp[1] * r[2] + gp[2] * r[3];
      gc[1] = gp[0];
      gc[2] = gp[1];
      const float gr[4] = {0.f, s * gp[0], s * gp[1], s * gp[2]};
      float gq[4] = {0.f, 0.f, 0.f, 0.f}, gt[4] = {0.f, 0.f, 0.f, 0.f};
      hamilton_bwd_a(gr, t, gq);
      hamilton_bwd_b(gr, q, gt);
      float gxq[4] = {0.f, 0.f, 0.f, 0.f}, gqc[4] = {0.f, 0.f, 0.f, 0.f};
      hamilton_bwd_a(gt, qc, gxq);
      hamilton_bwd_b(gt, xq, gqc);
      gc[3] = gq[0] + gqc[0];
      gc[4] = gq[1] - gqc[1];
      gc[5] = gq[2] - gqc[2];
      gc[6] = gq[3] - gqc[3];
      gx[0] += gxq[1]; gx[1] += gxq[2]; gx[2] += gxq[3];
    }
    if (grad_cams) {
#pragma unroll
      for (int i = 0; i < 7; ++i) {
        float a = gc[i];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
        if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5][i] = a;
      }
      __syncthreads();
      if (threadIdx.x < 7) {
        float a = 0.f;
#pragma unroll
        for (int w = 0; w < kThreads / 32; ++w) a += red[w][threadIdx.x];
        if (gridDim.x == 1) grad_cams[(size_t)n * 7 + threadIdx.x] = a;
        else atomicAdd(grad_cams + (size_t)n * 7 + threadIdx.x, a);
      }
      __syncthreads();
    }
  }
  if (live && grad_verts) {
    float* o = grad_verts + ((size_t)b * V + v) * 3;
    o[0] = gx[0]; o[1] = gx[1]; o[2] = gx[2];
  }
}

}  // namespace

extern "C" int acfm_project_fwd(const float* verts, const float* cams, int N, int NB, int V, float offset_z,
                                float sx, float sy, float z_add, float* out, void* stream) {
  ACFM_REQUIRE(N >= 0 && V >= 0 && NB > 0 || N == 0, ACFM_ERR_BAD_ARG, "acfm_project_fwd: bad sizes N=%d NB=%d V=%d", N, NB, V);
  if (N == 0 || V == 0) return ACFM_OK;
  ACFM_REQUIRE(verts && cams && out, ACFM_ERR_BAD_ARG, "acfm_project_fwd: null pointer");
  ACFM_REQUIRE(N % NB == 0, ACFM_ERR_BAD_ARG, "acfm_project_fwd: N=%d is not a multiple of NB=%d", N, NB);
  ACFM_REQUIRE(N <= 65535, ACFM_ERR_UNSUPPORTED, "acfm_project_fwd: N=%d > 65535 renders per call", N);
  dim3 grid((V + kThreads - 1) / kThreads, N);
  project_fwd_kernel<<<grid, kThreads, 0, (cudaStream_t)stream>>>(verts, cams, NB, V, offset_z, sx, sy, z_add, out);
  ACFM_LAUNCH_OK("project_fwd_kernel");
  return ACFM_OK;
}

extern "C" int acfm_project_bwd(const float* verts, const float* cams, const float* grad_out, int N, int NB, int V,
                                float sx, float sy, float* grad_verts, float* grad_cams, void* stream) {
  ACFM_REQUIRE(N >= 0 && V >= 0 && (NB > 0 || N == 0), ACFM_ERR_BAD_ARG, "acfm_project_bwd: bad sizes");
  if (N == 0 || V == 0) return ACFM_OK;
  ACFM_REQUIRE(verts && cams && grad_out, ACFM_ERR_BAD_ARG, "acfm_project_bwd: null pointer");
  ACFM_REQUIRE(N % NB == 0, ACFM_ERR_BAD_ARG, "acfm_project_bwd: N=%d is not a multiple of NB=%d", N, NB);
  ACFM_REQUIRE(NB <= 65535, ACFM_ERR_UNSUPPORTED, "acfm_project_bwd: NB=%d > 65535", NB);
  const int chunks = (V + kThreads - 1) / kThreads;
  if (grad_cams && chunks > 1)
    ACFM_CUDA_OK(cudaMemsetAsync(grad_cams, 0, sizeof(float) * 7 * (size_t)N, (cudaStream_t)stream));
  dim3 grid(chunks, NB);
  project_bwd_kernel<<<grid, kThreads, 0, (cudaStream_t)stream>>>(verts, cams, grad_out, N, NB, V, sx, sy, grad_verts,
                                                                   grad_cams);
  ACFM_LAUNCH_OK("project_bwd_kernel");
  return ACFM_OK;
}
