// handle_solve.cu — the per-step skinning matrix of the handle deformation, W = (L^T L + A^T A)^-1 A^T, by a Woodbury update
// of a once-inverted matrix, and its closed-form backward.  All arithmetic fp64 (cond(L^T L + c/V 11^T) ~ 1e4-1e5).
//
// Replaces, per training step, the reference's B*T identical 642 x 642 systems (repeat + 2 bmm + torch.cholesky +
// torch.cholesky_solve and their autograd: /root/reference/multiframe/main.py:586-609, monocular/main.py:203-218) — and round
// 1's own chain of ~40 cuBLAS / cuSOLVER fp64 micro-kernels (getrf, trsm, d884gemm, laswp: 0.3 ms per step of pure launch
// latency, profiles/launches_r04.md) — by three kernels forward and five backward.
//
// With lbs = softmax-over-vertices handle weights (V x Kh), A = lbs^T, U = [lbs, 1] (V x K1, K1 = Kh + 1),
// P~ = L^T L + (c/V) 1 1^T (constant; Pinv = P~^-1 computed once by the caller) and S = diag(I_Kh, -c/V):
//     M = L^T L + lbs lbs^T = P~ + U S U^T,   C = S^-1 + U^T Pinv U   (K1 x K1, symmetric, indefinite),
//     W = M^-1 lbs = (Pinv U) C^-1 [:, :Kh]                            (Woodbury; S^-1 E = E on the first Kh columns)
// Backward, for an upstream gW (V x Kh):   Z = M^-1 gW = Q - (Pinv U) C^-1 (U^T Q),  Q = Pinv gW,
//     d lbs = Z - Z (W^T lbs) - W (Z^T lbs).
//
// Data flow (R = 8 rows of V per CTA, nblk = ceil(V / R) CTAs; every reduction over V goes through per-CTA partial sums that
// ONE CTA adds up in a fixed order — deterministic, no floating-point atomics):
//   hs_rows_kernel<fwd>   PU = Pinv U                      + partial U^T PU                    -> Cpart[nblk][K1][K1]
//   hs_invert_kernel      C = S^-1 + sum Cpart; Gauss-Jordan with partial pivoting              -> Cinv[K1][K1]
//   hs_apply_kernel       W = PU Cinv[:, :Kh] (fp64 + fp32) + partial W^T lbs                   -> Apart[nblk][Kh][Kh]
//   hs_rows_kernel<bwd>   Q = Pinv gW                      + partial U^T Q                      -> Gpart[nblk][K1][Kh]
//   hs_small_kernel       T = Cinv (sum Gpart);  A = sum Apart
//   hs_z_kernel           Z = Q - PU T                     + partial Z^T lbs                    -> Bpart[nblk][Kh][Kh]
//   hs_reduce_kernel      B = sum Bpart
//   hs_grad_kernel        d lbs = Z - Z A - W B  (fp32)
#include "common.cuh"

namespace {

constexpr int kRows = 8;       // rows of V per CTA
constexpr int kThreads = 288;  // 8 rows x 33 columns (Kh = 32) + a few spare lanes; any Kh works through the strided loops

__device__ __forceinline__ double u_col(const float* lbs, int Kh, int k, int c) { return c < Kh ? (double)lbs[(size_t)k * Kh + c] : 1.0; }

// rows [r0, r0 + kRows) of  OUT = Pinv X,  X = [lbs, 1] (V x K1, forward) or gW (V x Kh, backward; no column of ones), and
// this CTA's partial of  U^T OUT  (K1 x NC).
template <bool kFwd>
__global__ void __launch_bounds__(kThreads) hs_rows_kernel(const double* __restrict__ Pinv, const float* __restrict__ X,
                                                           const float* __restrict__ lbs, int V, int Kh, double* __restrict__ out,
                                                           double* __restrict__ part) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* Ps = reinterpret_cast<double*>(smem_raw);  // [kRows][V]
  const int K1 = Kh + 1, NC = kFwd ? K1 : Kh;
  double* Os = Ps + (size_t)kRows * V;               // [kRows][NC]
  const int r0 = blockIdx.x * kRows, nr = min(kRows, V - r0);
  for (int e = threadIdx.x; e < nr * V; e += kThreads) Ps[e] = Pinv[(size_t)r0 * V + e];
  __syncthreads();
  for (int o = threadIdx.x; o < nr * NC; o += kThreads) {
    const int r = o / NC, c = o - r * NC;
    const double* pr = Ps + (size_t)r * V;
    double acc0 = 0.0, acc1 = 0.0, acc2 = 0.0, acc3 = 0.0;
    int k = 0;
    if (kFwd && c == Kh) {  // the column of ones: a row sum
      for (; k + 3 < V; k += 4) { acc0 += pr[k]; acc1 += pr[k + 1]; acc2 += pr[k + 2]; acc3 += pr[k + 3]; }
      for (; k < V; ++k) acc0 += pr[k];
    } else {
      const float* xc = X + c;
      for (; k + 3 < V; k += 4) {
        acc0 = fma(pr[k], (double)xc[(size_t)k * Kh], acc0);
        acc1 = fma(pr[k + 1], (double)xc[(size_t)(k + 1) * Kh], acc1);
        acc2 = fma(pr[k + 2], (double)xc[(size_t)(k + 2) * Kh], acc2);
        acc3 = fma(pr[k + 3], (double)xc[(size_t)(k + 3) * Kh], acc3);
      }
      for (; k < V; ++k) acc0 = fma(pr[k], (double)xc[(size_t)k * Kh], acc0);
    }
    const double v = (acc0 + acc1) + (acc2 + acc3);
    Os[o] = v;
    out[(size_t)(r0 + r) * NC + c] = v;
  }
  __syncthreads();
  double* pb = part + (size_t)blockIdx.x * K1 * NC;
  for (int e = threadIdx.x; e < K1 * NC; e += kThreads) {
    const int a = e / NC, b = e - a * NC;
    double s = 0.0;
    for (int r = 0; r < nr; ++r) s = fma(u_col(lbs, Kh, r0 + r, a), Os[r * NC + b], s);
    pb[e] = s;
  }
}

// C = S^-1 + sum of the partials, then C^-1 by Gauss-Jordan elimination with partial pivoting on [C | I] in shared memory.
// One CTA.  info[0] = 1 if a pivot vanished (singular C: the caller falls back to the direct solve).
__global__ void __launch_bounds__(1024) hs_invert_kernel(const double* __restrict__ part, int nblk, int K1, double sinv_last,
                                                         double* __restrict__ Cinv, int* __restrict__ info) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* M = reinterpret_cast<double*>(smem_raw);  // [K1][2 K1]
  __shared__ int piv_row;
  __shared__ double red_v[32];
  __shared__ int red_i[32];
  const int W2 = 2 * K1, tid = threadIdx.x, nt = blockDim.x;
  for (int e = tid; e < K1 * K1; e += nt) {
    double s = 0.0;
    for (int b = 0; b < nblk; ++b) s += part[(size_t)b * K1 * K1 + e];
    const int i = e / K1, j = e - i * K1;
    if (i == j) s += (i == K1 - 1) ? sinv_last : 1.0;
    M[i * W2 + j] = s;
    M[i * W2 + K1 + j] = (i == j) ? 1.0 : 0.0;
  }
  if (tid == 0) info[0] = 0;
  __syncthreads();
  for (int k = 0; k < K1; ++k) {
    // pivot: the largest |M[i][k]|, i >= k (first such row on ties: deterministic)
    double best = -1.0;
    int bi = k;
    for (int i = k + tid; i < K1; i += nt) {
      const double a = fabs(M[i * W2 + k]);
      if (a > best) { best = a; bi = i; }
    }
    for (int o = 16; o > 0; o >>= 1) {
      const double ob = __shfl_down_sync(0xffffffffu, best, o);
      const int oi = __shfl_down_sync(0xffffffffu, bi, o);
      if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
    }
    if ((tid & 31) == 0) { red_v[tid >> 5] = best; red_i[tid >> 5] = bi; }
    __syncthreads();
    if (tid == 0) {
      for (int w = 1; w < (nt >> 5); ++w)
        if (red_v[w] > best || (red_v[w] == best && red_i[w] < bi)) { best = red_v[w]; bi = red_i[w]; }
      piv_row = bi;
      if (!(best > 0.0)) info[0] = 1;
    }
    __syncthreads();
    const int pr = piv_row;
    if (pr != k)
      for (int j = tid; j < W2; j += nt) { const double t = M[k * W2 + j]; M[k * W2 + j] = M[pr * W2 + j]; M[pr * W2 + j] = t; }
    __syncthreads();
    const double inv = 1.0 / M[k * W2 + k];
    __syncthreads();
    for (int j = tid; j < W2; j += nt) M[k * W2 + j] *= inv;
    __syncthreads();
    // eliminate column k from every other row: element (i, j) -= M[i][k] * M[k][j]; column k itself last (it holds the factor)
    for (int e = tid; e < K1 * W2; e += nt) {
      const int i = e / W2, j = e - i * W2;
      if (i != k && j != k) M[e] = fma(-M[i * W2 + k], M[k * W2 + j], M[e]);
    }
    __syncthreads();
    for (int i = tid; i < K1; i += nt)
      if (i != k) M[i * W2 + k] = 0.0;
    __syncthreads();
  }
  for (int e = tid; e < K1 * K1; e += nt) Cinv[e] = M[(e / K1) * W2 + K1 + (e % K1)];
}

// W = PU Cinv[:, :Kh] for this CTA's rows (fp64 for the backward, fp32 for the caller) + partial W^T lbs (Kh x Kh)
__global__ void __launch_bounds__(kThreads) hs_apply_kernel(const double* __restrict__ PU, const double* __restrict__ Cinv,
                                                            const float* __restrict__ lbs, int V, int Kh, double* __restrict__ W64,
                                                            float* __restrict__ W32, double* __restrict__ Apart) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int K1 = Kh + 1;
  double* Cs = reinterpret_cast<double*>(smem_raw);  // [K1][Kh]
  double* Rs = Cs + (size_t)K1 * Kh;                 // [kRows][K1]  PU rows
  double* Ws = Rs + (size_t)kRows * K1;              // [kRows][Kh]
  const int r0 = blockIdx.x * kRows, nr = min(kRows, V - r0);
  for (int e = threadIdx.x; e < K1 * Kh; e += kThreads) Cs[e] = Cinv[(e / Kh) * K1 + (e % Kh)];
  for (int e = threadIdx.x; e < nr * K1; e += kThreads) Rs[e] = PU[(size_t)r0 * K1 + e];
  __syncthreads();
  for (int o = threadIdx.x; o < nr * Kh; o += kThreads) {
    const int r = o / Kh, j = o - r * Kh;
    double s = 0.0;
    for (int c = 0; c < K1; ++c) s = fma(Rs[r * K1 + c], Cs[c * Kh + j], s);
    Ws[o] = s;
    W64[(size_t)r0 * Kh + o] = s;
    W32[(size_t)r0 * Kh + o] = (float)s;
  }
  __syncthreads();
  double* pa = Apart + (size_t)blockIdx.x * Kh * Kh;
  for (int e = threadIdx.x; e < Kh * Kh; e += kThreads) {
    const int i = e / Kh, j = e - i * Kh;
    double s = 0.0;
    for (int r = 0; r < nr; ++r) s = fma(Ws[r * Kh + i], (double)lbs[(size_t)(r0 + r) * Kh + j], s);
    pa[e] = s;
  }
}

// T = Cinv (sum of Gpart)  (K1 x Kh)  and  A = sum of Apart  (Kh x Kh).  One CTA.
__global__ void __launch_bounds__(1024) hs_small_kernel(const double* __restrict__ Gpart, const double* __restrict__ Apart, int nblk,
                                                        int Kh, const double* __restrict__ Cinv, double* __restrict__ T,
                                                        double* __restrict__ A) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int K1 = Kh + 1, tid = threadIdx.x, nt = blockDim.x;
  double* G = reinterpret_cast<double*>(smem_raw);  // [K1][Kh]
  for (int e = tid; e < K1 * Kh; e += nt) {
    double s = 0.0;
    for (int b = 0; b < nblk; ++b) s += Gpart[(size_t)b * K1 * Kh + e];
    G[e] = s;
  }
  for (int e = tid; e < Kh * Kh; e += nt) {
    double s = 0.0;
    for (int b = 0; b < nblk; ++b) s += Apart[(size_t)b * Kh * Kh + e];
    A[e] = s;
  }
  __syncthreads();
  for (int e = tid; e < K1 * Kh; e += nt) {
    const int i = e / Kh, j = e - i * Kh;
    double s = 0.0;
    for (int c = 0; c < K1; ++c) s = fma(Cinv[i * K1 + c], G[c * Kh + j], s);
    T[e] = s;
  }
}

// Z = Q - PU T for this CTA's rows + partial Z^T lbs
__global__ void __launch_bounds__(kThreads) hs_z_kernel(const double* __restrict__ Q, const double* __restrict__ PU,
                                                        const double* __restrict__ T, const float* __restrict__ lbs, int V, int Kh,
                                                        double* __restrict__ Z, double* __restrict__ Bpart) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int K1 = Kh + 1;
  double* Ts = reinterpret_cast<double*>(smem_raw);  // [K1][Kh]
  double* Rs = Ts + (size_t)K1 * Kh;                 // [kRows][K1]
  double* Zs = Rs + (size_t)kRows * K1;              // [kRows][Kh]
  const int r0 = blockIdx.x * kRows, nr = min(kRows, V - r0);
  for (int e = threadIdx.x; e < K1 * Kh; e += kThreads) Ts[e] = T[e];
  for (int e = threadIdx.x; e < nr * K1; e += kThreads) Rs[e] = PU[(size_t)r0 * K1 + e];
  __syncthreads();
  for (int o = threadIdx.x; o < nr * Kh; o += kThreads) {
    const int r = o / Kh, j = o - r * Kh;
    double s = Q[(size_t)r0 * Kh + o];
    for (int c = 0; c < K1; ++c) s = fma(-Rs[r * K1 + c], Ts[c * Kh + j], s);
    Zs[o] = s;
    Z[(size_t)r0 * Kh + o] = s;
  }
  __syncthreads();
  double* pb = Bpart + (size_t)blockIdx.x * Kh * Kh;
  for (int e = threadIdx.x; e < Kh * Kh; e += kThreads) {
    const int i = e / Kh, j = e - i * Kh;
    double s = 0.0;
    for (int r = 0; r < nr; ++r) s = fma(Zs[r * Kh + i], (double)lbs[(size_t)(r0 + r) * Kh + j], s);
    pb[e] = s;
  }
}

__global__ void __launch_bounds__(1024) hs_reduce_kernel(const double* __restrict__ part, int nblk, int n, double* __restrict__ out) {
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < n; e += gridDim.x * blockDim.x) {
    double s = 0.0;
    for (int b = 0; b < nblk; ++b) s += part[(size_t)b * n + e];
    out[e] = s;
  }
}

// d lbs = Z - Z A - W B for this CTA's rows (fp32 out)
__global__ void __launch_bounds__(kThreads) hs_grad_kernel(const double* __restrict__ Z, const double* __restrict__ W64,
                                                           const double* __restrict__ A, const double* __restrict__ B, int V, int Kh,
                                                           float* __restrict__ g) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* As = reinterpret_cast<double*>(smem_raw);  // [Kh][Kh]
  double* Bs = As + (size_t)Kh * Kh;
  double* Zs = Bs + (size_t)Kh * Kh;                 // [kRows][Kh]
  double* Ws = Zs + (size_t)kRows * Kh;
  const int r0 = blockIdx.x * kRows, nr = min(kRows, V - r0);
  for (int e = threadIdx.x; e < Kh * Kh; e += kThreads) { As[e] = A[e]; Bs[e] = B[e]; }
  for (int e = threadIdx.x; e < nr * Kh; e += kThreads) { Zs[e] = Z[(size_t)r0 * Kh + e]; Ws[e] = W64[(size_t)r0 * Kh + e]; }
  __syncthreads();
  for (int o = threadIdx.x; o < nr * Kh; o += kThreads) {
    const int r = o / Kh, j = o - r * Kh;
    double s = Zs[o];
    for (int i = 0; i < Kh; ++i) s = fma(-Zs[r * Kh + i], As[i * Kh + j], fma(-Ws[r * Kh + i], Bs[i * Kh + j], s));
    g[(size_t)r0 * Kh + o] = (float)s;
  }
}

inline int hs_nblk(int V) { return (V + kRows - 1) / kRows; }

// workspace layout (doubles): PU | Cpart | Cinv | W64 | Apart | Q | Gpart | T | A | Z | Bpart | B | info(int, padded)
struct HsLayout {
  size_t PU, Cpart, Cinv, W64, Apart, Q, Gpart, T, A, Z, Bpart, B, info, total;
  HsLayout(int V, int Kh) {
    const size_t K1 = Kh + 1, nb = hs_nblk(V);
    size_t o = 0;
    PU = o; o += (size_t)V * K1;
    Cpart = o; o += nb * K1 * K1;
    Cinv = o; o += K1 * K1;
    W64 = o; o += (size_t)V * Kh;
    Apart = o; o += nb * Kh * Kh;
    Q = o; o += (size_t)V * Kh;
    Gpart = o; o += nb * K1 * Kh;
    T = o; o += K1 * Kh;
    A = o; o += (size_t)Kh * Kh;
    Z = o; o += (size_t)V * Kh;
    Bpart = o; o += nb * Kh * Kh;
    B = o; o += (size_t)Kh * Kh;
    info = o; o += 2;
    total = o;
  }
};

template <typename Kern>
int hs_smem(Kern kern, size_t bytes, std::atomic<int>* slot, const char* name) {
  ACFM_REQUIRE(bytes <= 227 * 1024, ACFM_ERR_UNSUPPORTED, "%s: needs %zu B of shared memory (V or the handle count is too large)", name, bytes);
  if (bytes > 48 * 1024) ACFM_CUDA_OK(acfm_ensure_smem(kern, (int)bytes, slot));
  return ACFM_OK;
}

}  // namespace

extern "C" int64_t acfm_handle_solve_workspace_bytes(int V, int Kh) {
  if (V <= 0 || Kh <= 0) return 0;
  return (int64_t)(HsLayout(V, Kh).total * sizeof(double));
}

extern "C" int acfm_handle_solve_fwd(const double* Pinv, const float* lbs, int V, int Kh, double c_over_V, float* W, void* workspace,
                                     int64_t workspace_bytes, void* stream) {
  ACFM_REQUIRE(V > 0 && Kh > 0 && Kh <= 128, ACFM_ERR_BAD_ARG, "acfm_handle_solve_fwd: bad sizes V=%d Kh=%d (Kh <= 128)", V, Kh);
  ACFM_REQUIRE(Pinv && lbs && W && workspace, ACFM_ERR_BAD_ARG, "acfm_handle_solve_fwd: null pointer");
  ACFM_REQUIRE(c_over_V > 0.0, ACFM_ERR_BAD_ARG, "acfm_handle_solve_fwd: c_over_V must be > 0");
  const HsLayout L(V, Kh);
  ACFM_REQUIRE(workspace_bytes >= (int64_t)(L.total * sizeof(double)) && (((uintptr_t)workspace) & 15u) == 0, ACFM_ERR_BAD_ARG,
               "acfm_handle_solve_fwd: workspace must be 16-byte aligned and hold acfm_handle_solve_workspace_bytes()");
  double* ws = (double*)workspace;
  cudaStream_t st = (cudaStream_t)stream;
  const int K1 = Kh + 1, nb = hs_nblk(V);
  static std::atomic<int> s_rows[kAcfmMaxDevices], s_inv[kAcfmMaxDevices], s_app[kAcfmMaxDevices];
  const size_t sm_rows = ((size_t)kRows * V + (size_t)kRows * K1) * sizeof(double);
  if (int rc = hs_smem(hs_rows_kernel<true>, sm_rows, s_rows, "acfm_handle_solve_fwd")) return rc;
  hs_rows_kernel<true><<<nb, kThreads, sm_rows, st>>>(Pinv, lbs, lbs, V, Kh, ws + L.PU, ws + L.Cpart);
  ACFM_LAUNCH_OK("hs_rows_kernel");
  const size_t sm_inv = (size_t)K1 * 2 * K1 * sizeof(double);
  if (int rc = hs_smem(hs_invert_kernel, sm_inv, s_inv, "acfm_handle_solve_fwd")) return rc;
  hs_invert_kernel<<<1, 1024, sm_inv, st>>>(ws + L.Cpart, nb, K1, -1.0 / c_over_V, ws + L.Cinv, (int*)(ws + L.info));
  ACFM_LAUNCH_OK("hs_invert_kernel");
  const size_t sm_app = ((size_t)K1 * Kh + (size_t)kRows * K1 + (size_t)kRows * Kh) * sizeof(double);
  if (int rc = hs_smem(hs_apply_kernel, sm_app, s_app, "acfm_handle_solve_fwd")) return rc;
  hs_apply_kernel<<<nb, kThreads, sm_app, st>>>(ws + L.PU, ws + L.Cinv, lbs, V, Kh, ws + L.W64, W, ws + L.Apart);
  ACFM_LAUNCH_OK("hs_apply_kernel");
  return ACFM_OK;
}

extern "C" int acfm_handle_solve_bwd(const double* Pinv, const float* lbs, const float* grad_W, int V, int Kh, float* grad_lbs,
                                     void* workspace, int64_t workspace_bytes, void* stream) {
  ACFM_REQUIRE(V > 0 && Kh > 0 && Kh <= 128, ACFM_ERR_BAD_ARG, "acfm_handle_solve_bwd: bad sizes V=%d Kh=%d (Kh <= 128)", V, Kh);
  ACFM_REQUIRE(Pinv && lbs && grad_W && grad_lbs && workspace, ACFM_ERR_BAD_ARG, "acfm_handle_solve_bwd: null pointer");
  const HsLayout L(V, Kh);
  ACFM_REQUIRE(workspace_bytes >= (int64_t)(L.total * sizeof(double)) && (((uintptr_t)workspace) & 15u) == 0, ACFM_ERR_BAD_ARG,
               "acfm_handle_solve_bwd: workspace must be the one acfm_handle_solve_fwd filled");
  double* ws = (double*)workspace;
  cudaStream_t st = (cudaStream_t)stream;
  const int K1 = Kh + 1, nb = hs_nblk(V);
  static std::atomic<int> s_rows[kAcfmMaxDevices], s_small[kAcfmMaxDevices], s_z[kAcfmMaxDevices], s_g[kAcfmMaxDevices];
  const size_t sm_rows = ((size_t)kRows * V + (size_t)kRows * Kh) * sizeof(double);
  if (int rc = hs_smem(hs_rows_kernel<false>, sm_rows, s_rows, "acfm_handle_solve_bwd")) return rc;
  hs_rows_kernel<false><<<nb, kThreads, sm_rows, st>>>(Pinv, grad_W, lbs, V, Kh, ws + L.Q, ws + L.Gpart);
  ACFM_LAUNCH_OK("hs_rows_kernel");
  const size_t sm_small = (size_t)K1 * Kh * sizeof(double);
  if (int rc = hs_smem(hs_small_kernel, sm_small, s_small, "acfm_handle_solve_bwd")) return rc;
  hs_small_kernel<<<1, 1024, sm_small, st>>>(ws + L.Gpart, ws + L.Apart, nb, Kh, ws + L.Cinv, ws + L.T, ws + L.A);
  ACFM_LAUNCH_OK("hs_small_kernel");
  const size_t sm_z = ((size_t)K1 * Kh + (size_t)kRows * K1 + (size_t)kRows * Kh) * sizeof(double);
  if (int rc = hs_smem(hs_z_kernel, sm_z, s_z, "acfm_handle_solve_bwd")) return rc;
  hs_z_kernel<<<nb, kThreads, sm_z, st>>>(ws + L.Q, ws + L.PU, ws + L.T, lbs, V, Kh, ws + L.Z, ws + L.Bpart);
  ACFM_LAUNCH_OK("hs_z_kernel");
  hs_reduce_kernel<<<1, 1024, 0, st>>>(ws + L.Bpart, nb, Kh * Kh, ws + L.B);
  ACFM_LAUNCH_OK("hs_reduce_kernel");
  const size_t sm_g = (2 * (size_t)Kh * Kh + 2 * (size_t)kRows * Kh) * sizeof(double);
  if (int rc = hs_smem(hs_grad_kernel, sm_g, s_g, "acfm_handle_solve_bwd")) return rc;
  hs_grad_kernel<<<nb, kThreads, sm_g, st>>>(ws + L.Z, ws + L.W64, ws + L.A, ws + L.B, V, Kh, grad_lbs);
  ACFM_LAUNCH_OK("hs_grad_kernel");
  return ACFM_OK;
}

// 1 if the last acfm_handle_solve_fwd on this workspace met a vanishing pivot (copies one int back: synchronises the stream;
// for tests and for the one-time check after construction, not for the per-step path)
extern "C" int acfm_handle_solve_singular(const void* workspace, int V, int Kh, void* stream) {
  const HsLayout L(V, Kh);
  int h = 0;
  if (cudaMemcpyAsync(&h, (const double*)workspace + L.info, sizeof(int), cudaMemcpyDeviceToHost, (cudaStream_t)stream) != cudaSuccess) return -1;
  if (cudaStreamSynchronize((cudaStream_t)stream) != cudaSuccess) return -1;
  return h;
}
