// camera.cu — camera-multiplex assembly: per (hypothesis g, frame b) embedding row -> 7-dof camera.
//
// Replaces the elementwise chain of ShapeTrainer.forward (/root/reference/multiframe/main.py:573-584):
//   scales = relu(scale_lr_decay * raw[0] + 1) + 1e-12 ; quats = normalize(raw[3:7]) ;
//   mirror_cameras (main.py:113-125; pytorch3d.transforms standardize_quaternion / quaternion_multiply
//   with matrix_to_quaternion(diag(-1,1,-1)) == (0,0,1,0): SURVEY.md §9.8) ;
//   transform_cameras (main.py:128-138),
// (~25 tiny torch kernels) with one pass forward and one backward.  raw / out are (G*NB,7)
// hypothesis-major; mirror_flag (NB) and transforms (NB,4) = [a, bx, by, flag] are indexed n % NB, which is
// what `.repeat(num_guesses)` expresses.
#include "common.cuh"

namespace {

constexpr int kThreads = 128;

__global__ void __launch_bounds__(kThreads) camera_fwd_kernel(const float* __restrict__ raw, const float* __restrict__ mirror,
                                                              const float* __restrict__ tf, int N, int NB, float lambda,
                                                              float* __restrict__ out) {
  const int n = blockIdx.x * kThreads + threadIdx.x;
  if (n >= N) return;
  const float* r = raw + (size_t)n * 7;
  float s = fmaxf(lambda * r[0] + 1.0f, 0.0f) + 1e-12f;
  float tx = r[1], ty = r[2];
  const float nrm = fmaxf(sqrtf(r[3] * r[3] + r[4] * r[4] + r[5] * r[5] + r[6] * r[6]), 1e-12f);  // F.normalize eps
  float q0 = r[3] / nrm, q1 = r[4] / nrm, q2 = r[5] / nrm, q3 = r[6] / nrm;
  if (mirror) {
    const float m = mirror[n % NB];
    // standardize, multiply by (0,0,1,0) on the left, standardize: (w,x,y,z) -> (-y, z, w, -x)
    const float s1 = q0 < 0.0f ? -1.0f : 1.0f;
    float m0 = -(s1 * q2), m1 = s1 * q3, m2 = s1 * q0, m3 = -(s1 * q1);
    if (m0 < 0.0f) { m0 = -m0; m1 = -m1; m2 = -m2; m3 = -m3; }
    tx = (1.0f - m) * tx + (-tx) * m;
    q0 = (1.0f - m) * q0 + m0 * m; q1 = (1.0f - m) * q1 + m1 * m;
    q2 = (1.0f - m) * q2 + m2 * m; q3 = (1.0f - m) * q3 + m3 * m;
  }
  if (tf) {
    const float* t = tf + (size_t)(n % NB) * 4;
    const float a = t[0], fl = t[3];
    s = (1.0f - fl) * s + (s * a) * fl;
    tx = (1.0f - fl) * tx + (tx * a + t[1]) * fl;
    ty = (1.0f - fl) * ty + (ty * a + t[2]) * fl;
  }
  float* o = out + (size_t)n * 7;
  o[0] = s; o[1] = tx; o[2] = ty; o[3] = q0; o[4] = q1; o[5] = q2; o[6] = q3;
}

__global__ void __launch_bounds__(kThreads) camera_bwd_kernel(const float* __restrict__ raw, const float* __restrict__ mirror,
                                                              const float* __restrict__ tf, const float* __restrict__ gout,
                                                              int N, int NB, float lambda, float* __restrict__ graw) {
  const int n = blockIdx.x * kThreads + threadIdx.x;
  if (n >= N) return;
  const float* r = raw + (size_t)n * 7;
  const float* g = gout + (size_t)n * 7;
  float gs = g[0], gtx = g[1], gty = g[2], g0 = g[3], g1 = g[4], g2 = g[5], g3 = g[6];
  if (tf) {
    const float* t = tf + (size_t)(n % NB) * 4;
    const float a = t[0], fl = t[3];
    const float k = (1.0f - fl) + a * fl;
    gs *= k; gtx *= k; gty *= k;
  }
  const float nrm_raw = sqrtf(r[3] * r[3] + r[4] * r[4] + r[5] * r[5] + r[6] * r[6]);
  const float nrm = fmaxf(nrm_raw, 1e-12f);
  const float q0 = r[3] / nrm, q1 = r[4] / nrm, q2 = r[5] / nrm, q3 = r[6] / nrm;
  if (mirror) {
    const float m = mirror[n % NB];
    gtx = (1.0f - m) * gtx - m * gtx;
    const float s1 = q0 < 0.0f ? -1.0f : 1.0f;
    const float s2 = (-(s1 * q2) < 0.0f) ? -1.0f : 1.0f;
    const float c = s1 * s2 * m;
    // mirrored quat = s2 * s1 * (-q2, q3, q0, -q1)
    const float n0 = (1.0f - m) * g0 + c * g2;
    const float n1 = (1.0f - m) * g1 - c * g3;
    const float n2 = (1.0f - m) * g2 - c * g0;
    const float n3 = (1.0f - m) * g3 + c * g1;
    g0 = n0; g1 = n1; g2 = n2; g3 = n3;
  }
  float* o = graw + (size_t)n * 7;
  o[0] = (lambda * r[0] + 1.0f > 0.0f) ? gs * lambda : 0.0f;
  o[1] = gtx; o[2] = gty;
  if (nrm_raw > 1e-12f) {
    const float dot = q0 * g0 + q1 * g1 + q2 * g2 + q3 * g3;
    o[3] = (g0 - q0 * dot) / nrm; o[4] = (g1 - q1 * dot) / nrm; o[5] = (g2 - q2 * dot) / nrm; o[6] = (g3 - q3 * dot) / nrm;
  } else {
    o[3] = g0 / nrm; o[4] = g1 / nrm; o[5] = g2 / nrm; o[6] = g3 / nrm;
  }
}

}  // namespace

extern "C" int acfm_camera_assemble_fwd(const float* raw, const float* mirror_flag, const float* transforms, int N, int NB,
                                        float scale_lr_decay, float* out, void* stream) {
  ACFM_REQUIRE(N >= 0 && (NB > 0 || N == 0), ACFM_ERR_BAD_ARG, "acfm_camera_assemble_fwd: bad sizes");
  if (N == 0) return ACFM_OK;
  ACFM_REQUIRE(raw && out, ACFM_ERR_BAD_ARG, "acfm_camera_assemble_fwd: null pointer");
  ACFM_REQUIRE(N % NB == 0, ACFM_ERR_BAD_ARG, "acfm_camera_assemble_fwd: N=%d is not a multiple of NB=%d", N, NB);
  camera_fwd_kernel<<<(N + kThreads - 1) / kThreads, kThreads, 0, (cudaStream_t)stream>>>(raw, mirror_flag, transforms, N, NB,
                                                                                          scale_lr_decay, out);
  ACFM_LAUNCH_OK("camera_fwd_kernel");
  return ACFM_OK;
}

extern "C" int acfm_camera_assemble_bwd(const float* raw, const float* mirror_flag, const float* transforms,
                                        const float* grad_out, int N, int NB, float scale_lr_decay, float* grad_raw,
                                        void* stream) {
  ACFM_REQUIRE(N >= 0 && (NB > 0 || N == 0), ACFM_ERR_BAD_ARG, "acfm_camera_assemble_bwd: bad sizes");
  if (N == 0) return ACFM_OK;
  ACFM_REQUIRE(raw && grad_out && grad_raw, ACFM_ERR_BAD_ARG, "acfm_camera_assemble_bwd: null pointer");
  ACFM_REQUIRE(N % NB == 0, ACFM_ERR_BAD_ARG, "acfm_camera_assemble_bwd: N=%d is not a multiple of NB=%d", N, NB);
  camera_bwd_kernel<<<(N + kThreads - 1) / kThreads, kThreads, 0, (cudaStream_t)stream>>>(raw, mirror_flag, transforms, grad_out,
                                                                                          N, NB, scale_lr_decay, grad_raw);
  ACFM_LAUNCH_OK("camera_bwd_kernel");
  return ACFM_OK;
}
