// losses.cu — fused per-render losses on the rendered silhouette.
//
// Replaces loss_utils.l1_loss / iou / iou_loss / edt_loss
// (/root/reference/multiframe/nnutils/loss_utils.py:18-32,72-77,245-253) and the G-fold
// `masks.repeat(num_guesses,1,1)` the callers materialise (multiframe/main.py:644,716): one pass over
// the mask computes all four per-render sums, the targets are indexed n % NB.
// HBM-bound: 4 B (mask) per pixel per render; targets are L2-resident across hypotheses.
#include "common.cuh"

namespace {

constexpr int kThreads = 256;

// sums[n] = { sum|m-t|, sum m*t, sum (m+t-m*t), sum edt*m }.  grid (chunks, N); sums zeroed by the launcher
__global__ void __launch_bounds__(kThreads) mask_sums_kernel(const float* __restrict__ mask, const float* __restrict__ target,
                                                             const float* __restrict__ edt, int NB, int HW,
                                                             float* __restrict__ sums) {
  const int n = blockIdx.y;
  const float* m = mask + (size_t)n * HW;
  const float* t = target + (size_t)(n % NB) * HW;
  const float* e = edt ? edt + (size_t)(n % NB) * HW : nullptr;
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
  const int per = (HW + gridDim.x - 1) / gridDim.x;
  const int lo = blockIdx.x * per, hi = min(lo + per, HW);
  for (int i = lo + threadIdx.x; i < hi; i += kThreads) {
    const float mv = m[i], tv = t[i];
    s0 += fabsf(mv - tv);
    s1 += mv * tv;
    s2 += mv + tv - mv * tv;
    if (e) s3 += e[i] * mv;
  }
  __shared__ float red[kThreads / 32][4];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s0 += __shfl_xor_sync(0xffffffffu, s0, o); s1 += __shfl_xor_sync(0xffffffffu, s1, o);
    s2 += __shfl_xor_sync(0xffffffffu, s2, o); s3 += __shfl_xor_sync(0xffffffffu, s3, o);
  }
  if ((threadIdx.x & 31) == 0) { red[threadIdx.x >> 5][0] = s0; red[threadIdx.x >> 5][1] = s1; red[threadIdx.x >> 5][2] = s2; red[threadIdx.x >> 5][3] = s3; }
  __syncthreads();
  if (threadIdx.x < 4) {
    float a = 0.f;
#pragma unroll
    for (int w = 0; w < kThreads / 32; ++w) a += red[w][threadIdx.x];
    if (gridDim.x == 1) sums[(size_t)n * 4 + threadIdx.x] = a;
    else atomicAdd(sums + (size_t)n * 4 + threadIdx.x, a);
  }
}

// grad_mask = g0*sign(m-t) + g1*t + g2*(1-t) + g3*edt with g = grad_sums[n]
__global__ void __launch_bounds__(kThreads) mask_sums_bwd_kernel(const float* __restrict__ mask, const float* __restrict__ target,
                                                                 const float* __restrict__ edt, const float* __restrict__ grad_sums,
                                                                 int NB, int HW, float* __restrict__ grad_mask) {
  const int n = blockIdx.y;
  const float g0 = grad_sums[(size_t)n * 4], g1 = grad_sums[(size_t)n * 4 + 1], g2 = grad_sums[(size_t)n * 4 + 2],
              g3 = grad_sums[(size_t)n * 4 + 3];
  const float* m = mask + (size_t)n * HW;
  const float* t = target + (size_t)(n % NB) * HW;
  const float* e = edt ? edt + (size_t)(n % NB) * HW : nullptr;
  float* o = grad_mask + (size_t)n * HW;
  for (int i = blockIdx.x * kThreads + threadIdx.x; i < HW; i += gridDim.x * kThreads) {
    const float mv = m[i], tv = t[i];
    const float d = mv - tv;
    float g = g0 * (d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f)) + g1 * tv + g2 * (1.f - tv);
    if (e) g += g3 * e[i];
    o[i] = g;
  }
}

// per[n] = w_l1 * s0 / HW + w_iou * (1 - s1 / (s2 + 1e-6)) + w_edt * s3 / HW: l1_loss, iou_loss and edt_loss (reduce=False) of one
// render from its four sums, weighted and added as the callers do (multiframe/main.py:644-645,715-716)
__global__ void __launch_bounds__(kThreads) mask_loss_combine_kernel(const float* __restrict__ sums, int N, float inv_hw, float w_l1,
                                                                     float w_iou, float w_edt, float* __restrict__ per) {
  const int n = blockIdx.x * kThreads + threadIdx.x;
  if (n >= N) return;
  const float4 s = *reinterpret_cast<const float4*>(sums + (size_t)n * 4);
  float v = w_l1 * (s.x * inv_hw) + w_edt * (s.w * inv_hw);
  if (w_iou != 0.0f) v += w_iou * (1.0f - s.y / (s.z + 1e-6f));
  per[n] = v;
}
__global__ void __launch_bounds__(kThreads) mask_loss_combine_bwd_kernel(const float* __restrict__ sums, const float* __restrict__ grad_per,
                                                                         int N, float inv_hw, float w_l1, float w_iou, float w_edt,
                                                                         float* __restrict__ grad_sums) {
  const int n = blockIdx.x * kThreads + threadIdx.x;
  if (n >= N) return;
  const float4 s = *reinterpret_cast<const float4*>(sums + (size_t)n * 4);
  const float g = grad_per[n], d = 1.0f / (s.z + 1e-6f);
  *reinterpret_cast<float4*>(grad_sums + (size_t)n * 4) =
      make_float4(g * w_l1 * inv_hw, -g * w_iou * d, g * w_iou * s.y * d * d, g * w_edt * inv_hw);
}

}  // namespace

extern "C" int acfm_mask_loss_combine_fwd(const float* sums, int N, int HW, float w_l1, float w_iou, float w_edt, float* per,
                                          void* stream) {
  ACFM_REQUIRE(N >= 0 && HW > 0, ACFM_ERR_BAD_ARG, "acfm_mask_loss_combine_fwd: bad sizes");
  if (N == 0) return ACFM_OK;
  ACFM_REQUIRE(sums && per && (((uintptr_t)sums) & 15u) == 0, ACFM_ERR_BAD_ARG, "acfm_mask_loss_combine_fwd: null or misaligned pointer");
  mask_loss_combine_kernel<<<(N + kThreads - 1) / kThreads, kThreads, 0, (cudaStream_t)stream>>>(sums, N, 1.0f / (float)HW, w_l1, w_iou,
                                                                                                  w_edt, per);
  ACFM_LAUNCH_OK("mask_loss_combine_kernel");
  return ACFM_OK;
}

extern "C" int acfm_mask_loss_combine_bwd(const float* sums, const float* grad_per, int N, int HW, float w_l1, float w_iou, float w_edt,
                                          float* grad_sums, void* stream) {
  ACFM_REQUIRE(N >= 0 && HW > 0, ACFM_ERR_BAD_ARG, "acfm_mask_loss_combine_bwd: bad sizes");
  if (N == 0) return ACFM_OK;
  ACFM_REQUIRE(sums && grad_per && grad_sums && ((((uintptr_t)sums) | ((uintptr_t)grad_sums)) & 15u) == 0, ACFM_ERR_BAD_ARG,
               "acfm_mask_loss_combine_bwd: null or misaligned pointer");
  mask_loss_combine_bwd_kernel<<<(N + kThreads - 1) / kThreads, kThreads, 0, (cudaStream_t)stream>>>(sums, grad_per, N, 1.0f / (float)HW,
                                                                                                      w_l1, w_iou, w_edt, grad_sums);
  ACFM_LAUNCH_OK("mask_loss_combine_bwd_kernel");
  return ACFM_OK;
}

extern "C" int acfm_mask_sums_fwd(const float* mask, const float* target, const float* edt, int N, int NB, int HW,
                                  float* sums, void* stream) {
  ACFM_REQUIRE(N >= 0 && HW > 0 && (NB > 0 || N == 0), ACFM_ERR_BAD_ARG, "acfm_mask_sums_fwd: bad sizes");
  if (N == 0) return ACFM_OK;
  ACFM_REQUIRE(mask && target && sums, ACFM_ERR_BAD_ARG, "acfm_mask_sums_fwd: null pointer");
  ACFM_REQUIRE(N % NB == 0, ACFM_ERR_BAD_ARG, "acfm_mask_sums_fwd: N=%d is not a multiple of NB=%d", N, NB);
  ACFM_REQUIRE(N <= 65535, ACFM_ERR_UNSUPPORTED, "acfm_mask_sums_fwd: N=%d > 65535", N);
  cudaStream_t st = (cudaStream_t)stream;
  int chunks = 1;
  while (chunks * N < 592 && chunks * kThreads * 4 < HW) chunks *= 2;  // >= 4 CTAs per SM when N is small
  if (chunks > 1) ACFM_CUDA_OK(cudaMemsetAsync(sums, 0, sizeof(float) * 4 * (size_t)N, st));
  mask_sums_kernel<<<dim3(chunks, N), kThreads, 0, st>>>(mask, target, edt, NB, HW, sums);
  ACFM_LAUNCH_OK("mask_sums_kernel");
  return ACFM_OK;
}

extern "C" int acfm_mask_sums_bwd(const float* mask, const float* target, const float* edt, const float* grad_sums, int N,
                                  int NB, int HW, float* grad_mask, void* stream) {
  ACFM_REQUIRE(N >= 0 && HW > 0 && (NB > 0 || N == 0), ACFM_ERR_BAD_ARG, "acfm_mask_sums_bwd: bad sizes");
  if (N == 0) return ACFM_OK;
  ACFM_REQUIRE(mask && target && grad_sums && grad_mask, ACFM_ERR_BAD_ARG, "acfm_mask_sums_bwd: null pointer");
  ACFM_REQUIRE(N % NB == 0, ACFM_ERR_BAD_ARG, "acfm_mask_sums_bwd: N=%d is not a multiple of NB=%d", N, NB);
  ACFM_REQUIRE(N <= 65535, ACFM_ERR_UNSUPPORTED, "acfm_mask_sums_bwd: N=%d > 65535", N);
  const int chunks = max(1, min(16, HW / (kThreads * 8)));
  mask_sums_bwd_kernel<<<dim3(chunks, N), kThreads, 0, (cudaStream_t)stream>>>(mask, target, edt, grad_sums, NB, HW, grad_mask);
  ACFM_LAUNCH_OK("mask_sums_bwd_kernel");
  return ACFM_OK;
}
