// shade.cu — texture branch of NeuralRenderer.forward after the hard rasterization:
// TexturesAtlas.sample_textures (nearest texel from clipped barycentrics) or per-vertex colour
// interpolation, ambient-only Phong (colour == texel) and softmax_rgb_blend, fused into one pass
// (/root/reference/multiframe/nnutils/nmr.py:173-200; PyTorch3D 0.3.0 renderer/mesh/textures.py,
// renderer/blending.py).  The reference materialises (N,H,W,K,3) texels and ~10
// elementwise temporaries; here each pixel is read once and one RGBA float4 is written.
// HBM-bound: 8K (pix_to_face) + 20K (bary, dists, zbuf) bytes read + 16 B written per pixel.
#include "common.cuh"

namespace {

constexpr int kThreads = 256;
constexpr int kMaxK = 8;

struct ShadeParams {
  const long long* p2f;
  const float* bary;
  const float* dists;
  const float* zbuf;
  int N, H, W, K;
  int mode;          // 0: atlas (N*F,R,R,3); 1: vertex colours (NC,V,3) + faces
  const float* tex;
  int R, V, F, NC;
  const void* faces;
  int faces_i64;
  long long faces_stride;
  float sigma, gamma, znear, zfar;
};

__device__ __forceinline__ void face_verts(const ShadeParams& p, int n, int f, int& i0, int& i1, int& i2) {
  const long long o = (long long)n * p.faces_stride + (long long)f * 3;
  if (p.faces_i64) {
    const long long* q = reinterpret_cast<const long long*>(p.faces) + o;
    i0 = (int)q[0]; i1 = (int)q[1]; i2 = (int)q[2];
  } else {
    const int* q = reinterpret_cast<const int*>(p.faces) + o;
    i0 = q[0]; i1 = q[1]; i2 = q[2];
  }
}

// TexturesAtlas.sample_textures index: w = trunc(bary01 * R); flip to the upper triangle of the RxR
// texel grid when (b0 + b1) * R - (wx + wy) > 1
__device__ __forceinline__ int atlas_index(float b0, float b1, int R) {
  int wx = (int)fmul(b0, (float)R), wy = (int)fmul(b1, (float)R);
  const bool below = fsub(fmul(fadd(b0, b1), (float)R), fadd((float)wx, (float)wy)) <= 1.0f;  // one rounding per torch op
  if (!below) { wx = R - 1 - wx; wy = R - 1 - wy; }
  wx = min(max(wx, 0), R - 1); wy = min(max(wy, 0), R - 1);  // guards bary == 1 exactly (reference would index out of range)
  return wy * R + wx;
}

__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }

template <bool BWD>
__global__ void __launch_bounds__(kThreads) shade_kernel(const ShadeParams p, float* __restrict__ rgba,
                                                         const float* __restrict__ grad_rgba, float* __restrict__ grad_tex,
                                                         float* __restrict__ grad_dists) {
  const long long npix = (long long)p.N * p.H * p.W;
  const long long pix = (long long)blockIdx.x * kThreads + threadIdx.x;
  if (pix >= npix) return;
  const int n = (int)(pix / ((long long)p.H * p.W));
  const int K = p.K;
  const float eps = 1e-10f;
  float prob[kMaxK], zinv[kMaxK];
  float zmax = eps, alpha = 1.0f;
  for (int k = 0; k < K; ++k) {
    const long long f = p.p2f[pix * K + k];
    const float m = f >= 0 ? 1.0f : 0.0f;
    prob[k] = sigmoidf_(-p.dists[pix * K + k] / p.sigma) * m;
    zinv[k] = (p.zfar - p.zbuf[pix * K + k]) / (p.zfar - p.znear) * m;
    zmax = fmaxf(zmax, zinv[k]);
    alpha *= (1.0f - prob[k]);
  }
  const float delta = fmaxf(expf((eps - zmax) / p.gamma), eps);
  float denom = delta, r = 0.f, g = 0.f, b = 0.f;
  float wnum[kMaxK];
  for (int k = 0; k < K; ++k) {
    wnum[k] = prob[k] * expf((zinv[k] - zmax) / p.gamma);
    denom += wnum[k];
  }
  float gr = 0.f, gg = 0.f, gb = 0.f, ga = 0.f;
  if (BWD) {
    const float4 go = reinterpret_cast<const float4*>(grad_rgba)[pix];
    gr = go.x; gg = go.y; gb = go.z; ga = go.w;
  }
  float cr[kMaxK], cg[kMaxK], cb[kMaxK];
  for (int k = 0; k < K; ++k) {
    const long long f = p.p2f[pix * K + k];
    cr[k] = cg[k] = cb[k] = 0.f;
    if (f < 0) continue;
    const float b0 = p.bary[(pix * K + k) * 3], b1 = p.bary[(pix * K + k) * 3 + 1], b2 = p.bary[(pix * K + k) * 3 + 2];
    if (p.mode == 0) {
      const float* t = p.tex + ((size_t)f * p.R * p.R + atlas_index(b0, b1, p.R)) * 3;
      cr[k] = t[0]; cg[k] = t[1]; cb[k] = t[2];
      if (BWD) {
        float* gt = grad_tex + ((size_t)f * p.R * p.R + atlas_index(b0, b1, p.R)) * 3;
        const float w = wnum[k] / denom;
        if (w != 0.f) { atomicAdd(gt, gr * w); atomicAdd(gt + 1, gg * w); atomicAdd(gt + 2, gb * w); }
      }
    } else {
      const int fl = (int)(f - (long long)n * p.F);
      int i0, i1, i2;
      face_verts(p, n, fl, i0, i1, i2);
      const size_t cbase = (size_t)(n % p.NC) * p.V * 3;
      const float* c0 = p.tex + cbase + (size_t)i0 * 3;
      const float* c1 = p.tex + cbase + (size_t)i1 * 3;
      const float* c2 = p.tex + cbase + (size_t)i2 * 3;
      cr[k] = b0 * c0[0] + b1 * c1[0] + b2 * c2[0];
      cg[k] = b0 * c0[1] + b1 * c1[1] + b2 * c2[1];
      cb[k] = b0 * c0[2] + b1 * c1[2] + b2 * c2[2];
      if (BWD) {
        const float w = wnum[k] / denom;
        if (w != 0.f) {
          float* g0 = grad_tex + cbase + (size_t)i0 * 3;
          float* g1 = grad_tex + cbase + (size_t)i1 * 3;
          float* g2 = grad_tex + cbase + (size_t)i2 * 3;
          atomicAdd(g0, gr * w * b0); atomicAdd(g0 + 1, gg * w * b0); atomicAdd(g0 + 2, gb * w * b0);
          atomicAdd(g1, gr * w * b1); atomicAdd(g1 + 1, gg * w * b1); atomicAdd(g1 + 2, gb * w * b1);
          atomicAdd(g2, gr * w * b2); atomicAdd(g2 + 1, gg * w * b2); atomicAdd(g2 + 2, gb * w * b2);
        }
      }
    }
    r += wnum[k] * cr[k]; g += wnum[k] * cg[k]; b += wnum[k] * cb[k];
  }
  r /= denom; g /= denom; b /= denom;
  if (!BWD) {
    reinterpret_cast<float4*>(rgba)[pix] = make_float4(r, g, b, 1.0f - alpha);
    return;
  }
  if (grad_dists) {
    // through prob only (zbuf enters via zinv - zmax, identically 0 for the K = 1 the reference uses)
    for (int k = 0; k < K; ++k) {
      float gd = 0.f;
      if (p.p2f[pix * K + k] >= 0) {
        const float e = expf((zinv[k] - zmax) / p.gamma);
        const float g_w = (gr * (cr[k] - r) + gg * (cg[k] - g) + gb * (cb[k] - b)) / denom;  // d loss / d wnum_k
        float others = 1.0f;
        for (int j = 0; j < K; ++j) if (j != k) others *= (1.0f - prob[j]);
        const float g_prob = g_w * e + ga * others;
        gd = g_prob * prob[k] * (1.0f - prob[k]) * (-1.0f / p.sigma);
      }
      grad_dists[pix * K + k] = gd;
    }
  }
}

int fill(ShadeParams& p, const int64_t* p2f, const float* bary, const float* dists, const float* zbuf, int N, int H, int W,
         int K, int mode, const float* tex, int R, int V, int F, int NC, const void* faces, int faces_i64,
         int64_t faces_stride, float sigma, float gamma, float znear, float zfar, const char* who) {
  ACFM_REQUIRE(N >= 0 && H > 0 && W > 0 && K >= 1, ACFM_ERR_BAD_ARG, "%s: bad sizes", who);
  ACFM_REQUIRE(K <= kMaxK, ACFM_ERR_UNSUPPORTED, "%s: faces_per_pixel K=%d > %d", who, K, kMaxK);
  ACFM_REQUIRE(mode == 0 || mode == 1, ACFM_ERR_BAD_ARG, "%s: mode must be 0 (atlas) or 1 (vertex colours)", who);
  ACFM_REQUIRE(sigma > 0.f && gamma > 0.f && zfar > znear, ACFM_ERR_BAD_ARG, "%s: bad blend parameters", who);
  if (N == 0) return ACFM_OK;
  ACFM_REQUIRE(p2f && bary && dists && zbuf && tex, ACFM_ERR_BAD_ARG, "%s: null pointer", who);
  ACFM_REQUIRE(mode == 0 ? R >= 1 : (faces && V > 0 && F > 0 && NC > 0 && N % NC == 0), ACFM_ERR_BAD_ARG, "%s: bad texture arguments", who);
  p.p2f = (const long long*)p2f; p.bary = bary; p.dists = dists; p.zbuf = zbuf;
  p.N = N; p.H = H; p.W = W; p.K = K; p.mode = mode; p.tex = tex; p.R = R; p.V = V; p.F = F; p.NC = NC > 0 ? NC : 1;
  p.faces = faces; p.faces_i64 = faces_i64; p.faces_stride = faces_stride;
  p.sigma = sigma; p.gamma = gamma; p.znear = znear; p.zfar = zfar;
  return ACFM_OK;
}

}  // namespace

extern "C" int acfm_shade_fwd(const int64_t* pix_to_face, const float* bary, const float* dists, const float* zbuf, int N,
                              int H, int W, int K, int mode, const float* tex, int R, int V, int F, int NC,
                              const void* faces, int faces_i64, int64_t faces_batch_stride, float sigma, float gamma,
                              float znear, float zfar, float* rgba, void* stream) {
  ShadeParams p;
  const int st = fill(p, pix_to_face, bary, dists, zbuf, N, H, W, K, mode, tex, R, V, F, NC, faces, faces_i64,
                      faces_batch_stride, sigma, gamma, znear, zfar, "acfm_shade_fwd");
  if (st != ACFM_OK || N == 0) return st;
  ACFM_REQUIRE(rgba, ACFM_ERR_BAD_ARG, "acfm_shade_fwd: null output");
  const long long npix = (long long)N * H * W;
  shade_kernel<false><<<(unsigned)((npix + kThreads - 1) / kThreads), kThreads, 0, (cudaStream_t)stream>>>(p, rgba, nullptr, nullptr, nullptr);
  ACFM_LAUNCH_OK("shade_kernel<fwd>");
  return ACFM_OK;
}

extern "C" int acfm_shade_bwd(const int64_t* pix_to_face, const float* bary, const float* dists, const float* zbuf, int N,
                              int H, int W, int K, int mode, const float* tex, int R, int V, int F, int NC,
                              const void* faces, int faces_i64, int64_t faces_batch_stride, float sigma, float gamma,
                              float znear, float zfar, const float* grad_rgba, float* grad_tex, int64_t grad_tex_numel,
                              float* grad_dists, void* stream) {
  ShadeParams p;
  const int st = fill(p, pix_to_face, bary, dists, zbuf, N, H, W, K, mode, tex, R, V, F, NC, faces, faces_i64,
                      faces_batch_stride, sigma, gamma, znear, zfar, "acfm_shade_bwd");
  if (st != ACFM_OK || N == 0) return st;
  ACFM_REQUIRE(grad_rgba && grad_tex && grad_tex_numel >= 0, ACFM_ERR_BAD_ARG, "acfm_shade_bwd: null pointer");
  ACFM_CUDA_OK(cudaMemsetAsync(grad_tex, 0, sizeof(float) * (size_t)grad_tex_numel, (cudaStream_t)stream));
  const long long npix = (long long)N * H * W;
  shade_kernel<true><<<(unsigned)((npix + kThreads - 1) / kThreads), kThreads, 0, (cudaStream_t)stream>>>(p, nullptr, grad_rgba, grad_tex, grad_dists);
  ACFM_LAUNCH_OK("shade_kernel<bwd>");
  return ACFM_OK;
}
