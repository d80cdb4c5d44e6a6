// uvsample.cu — texture-flow bilinear UV sampling: the sampling step of TexturePredictorUV.forward
// (/root/reference/multiframe/nnutils/mesh_net.py:166-179; monocular/nnutils/mesh_net.py:166-180):
//   tex = grid_sample(uvimage (B,C,Hu,Wu), uv_sampler (1,F,T*T,2).repeat(B), align_corners=True)   [bilinear, zeros]
//   tex = tex.reshape(B,C,F,T,T).permute(0,2,3,4,1) ; tex = (tanh(tex) + 1) / 2
// fused into one gather pass writing the atlas (B,F,T,T,C) directly (the reference's permute leaves a
// strided view; values are identical).  The sampling grid is shared by all frames.
// HBM-bound: C*4 B written per atlas texel; the UV image (C*Hu*Wu*4 B per frame) is L2-resident.
#include "common.cuh"

namespace {

constexpr int kThreads = 256;

struct Taps { int x0, y0; float w00, w01, w10, w11; bool in00, in01, in10, in11; };

__device__ __forceinline__ Taps make_taps(float gx, float gy, int Hu, int Wu) {
  // grid_sampler_unnormalize(align_corners=True): ((coord + 1) / 2) * (size - 1)
  const float ix = ((gx + 1.0f) / 2.0f) * (float)(Wu - 1), iy = ((gy + 1.0f) / 2.0f) * (float)(Hu - 1);
  const float fx = floorf(ix), fy = floorf(iy);
  Taps t;
  t.x0 = (int)fx; t.y0 = (int)fy;
  const float ax = ix - fx, ay = iy - fy;
  t.w00 = (1.0f - ax) * (1.0f - ay); t.w01 = ax * (1.0f - ay); t.w10 = (1.0f - ax) * ay; t.w11 = ax * ay;
  const bool x0in = t.x0 >= 0 && t.x0 < Wu, x1in = t.x0 + 1 >= 0 && t.x0 + 1 < Wu;
  const bool y0in = t.y0 >= 0 && t.y0 < Hu, y1in = t.y0 + 1 >= 0 && t.y0 + 1 < Hu;
  t.in00 = x0in && y0in; t.in01 = x1in && y0in; t.in10 = x0in && y1in; t.in11 = x1in && y1in;
  return t;
}

// grid (ceil(P/kThreads), B);  out (B,P,C)
__global__ void __launch_bounds__(kThreads) uv_sample_fwd_kernel(const float* __restrict__ img, const float* __restrict__ grid,
                                                                 int C, int Hu, int Wu, int P, int apply_tanh,
                                                                 float* __restrict__ out) {
  const int b = blockIdx.y, pt = blockIdx.x * kThreads + threadIdx.x;
  if (pt >= P) return;
  const Taps t = make_taps(grid[pt * 2], grid[pt * 2 + 1], Hu, Wu);
  const size_t plane = (size_t)Hu * Wu;
  const float* im = img + (size_t)b * C * plane + (size_t)t.y0 * Wu + t.x0;
  float* o = out + ((size_t)b * P + pt) * C;
  for (int c = 0; c < C; ++c) {
    const float* q = im + c * plane;
    float v = 0.0f;
    if (t.in00) v += q[0] * t.w00;
    if (t.in01) v += q[1] * t.w01;
    if (t.in10) v += q[Wu] * t.w10;
    if (t.in11) v += q[Wu + 1] * t.w11;
    o[c] = apply_tanh ? (tanhf(v) + 1.0f) * 0.5f : v;
  }
}

__global__ void __launch_bounds__(kThreads) uv_sample_bwd_kernel(const float* __restrict__ out, const float* __restrict__ grid,
                                                                 const float* __restrict__ gout, int C, int Hu, int Wu, int P,
                                                                 int apply_tanh, float* __restrict__ gimg) {
  const int b = blockIdx.y, pt = blockIdx.x * kThreads + threadIdx.x;
  if (pt >= P) return;
  const Taps t = make_taps(grid[pt * 2], grid[pt * 2 + 1], Hu, Wu);
  const size_t plane = (size_t)Hu * Wu;
  float* gi = gimg + (size_t)b * C * plane + (size_t)t.y0 * Wu + t.x0;
  const size_t oi = ((size_t)b * P + pt) * C;
  for (int c = 0; c < C; ++c) {
    float g = gout[oi + c];
    if (apply_tanh) {
      const float th = 2.0f * out[oi + c] - 1.0f;  // tanh(v) recovered from the saved output
      g *= 0.5f * (1.0f - th * th);
    }
    if (g == 0.0f) continue;
    float* q = gi + c * plane;
    if (t.in00) atomicAdd(q, g * t.w00);
    if (t.in01) atomicAdd(q + 1, g * t.w01);
    if (t.in10) atomicAdd(q + Wu, g * t.w10);
    if (t.in11) atomicAdd(q + Wu + 1, g * t.w11);
  }
}

}  // namespace

extern "C" int acfm_uv_sample_fwd(const float* uvimage, const float* grid, int B, int C, int Hu, int Wu, int P, int apply_tanh,
                                  float* out, void* stream) {
  ACFM_REQUIRE(B >= 0 && C > 0 && Hu > 0 && Wu > 0 && P >= 0, ACFM_ERR_BAD_ARG, "acfm_uv_sample_fwd: bad sizes");
  if (B == 0 || P == 0) return ACFM_OK;
  ACFM_REQUIRE(uvimage && grid && out, ACFM_ERR_BAD_ARG, "acfm_uv_sample_fwd: null pointer");
  ACFM_REQUIRE(B <= 65535, ACFM_ERR_UNSUPPORTED, "acfm_uv_sample_fwd: B=%d > 65535", B);
  uv_sample_fwd_kernel<<<dim3((P + kThreads - 1) / kThreads, B), kThreads, 0, (cudaStream_t)stream>>>(uvimage, grid, C, Hu, Wu, P,
                                                                                                      apply_tanh, out);
  ACFM_LAUNCH_OK("uv_sample_fwd_kernel");
  return ACFM_OK;
}

extern "C" int acfm_uv_sample_bwd(const float* out, const float* grid, const float* grad_out, int B, int C, int Hu, int Wu,
                                  int P, int apply_tanh, float* grad_uvimage, void* stream) {
  ACFM_REQUIRE(B >= 0 && C > 0 && Hu > 0 && Wu > 0 && P >= 0, ACFM_ERR_BAD_ARG, "acfm_uv_sample_bwd: bad sizes");
  if (B == 0) return ACFM_OK;
  ACFM_REQUIRE(grad_uvimage, ACFM_ERR_BAD_ARG, "acfm_uv_sample_bwd: null pointer");
  ACFM_CUDA_OK(cudaMemsetAsync(grad_uvimage, 0, sizeof(float) * (size_t)B * C * Hu * Wu, (cudaStream_t)stream));
  if (P == 0) return ACFM_OK;
  ACFM_REQUIRE(out && grid && grad_out, ACFM_ERR_BAD_ARG, "acfm_uv_sample_bwd: null pointer");
  ACFM_REQUIRE(B <= 65535, ACFM_ERR_UNSUPPORTED, "acfm_uv_sample_bwd: B=%d > 65535", B);
  uv_sample_bwd_kernel<<<dim3((P + kThreads - 1) / kThreads, B), kThreads, 0, (cudaStream_t)stream>>>(out, grid, grad_out, C, Hu, Wu,
                                                                                                      P, apply_tanh, grad_uvimage);
  ACFM_LAUNCH_OK("uv_sample_bwd_kernel");
  return ACFM_OK;
}
