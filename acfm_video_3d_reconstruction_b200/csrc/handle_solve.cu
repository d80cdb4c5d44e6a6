// handle_solve.cu — the per-step skinning matrix of the handle deformation, W = (L^T L + A^T A)^-1 A^T, by a Woodbury update
// of a once-inverted matrix, and its closed-form backward.  All arithmetic fp64 (cond(L^T L + c/V 11^T) ~ 1e4-1e5).
//
// Replaces, per training step, the reference's B*T identical 642 x 642 systems (repeat + 2 bmm + torch.cholesky +
// torch.cholesky_solve and their autograd: /root/reference/multiframe/main.py:586-609, monocular/main.py:203-218) — and round
// 1's own chain of ~40 cuBLAS / cuSOLVER fp64 micro-kernels (getrf, trsm, d884gemm, laswp: 0.3 ms per step of pure launch
// latency, profiles/launches_r04.md) — by four kernels forward and seven backward.
//
// With lbs = softmax-over-vertices handle weights (V x Kh), A = lbs^T, U = [lbs, 1] (V x K1, K1 = Kh + 1),
// P~ = L^T L + (c/V) 1 1^T (constant; Pinv = P~^-1 computed once by the caller) and S = diag(I_Kh, -c/V):
//     M = L^T L + lbs lbs^T = P~ + U S U^T,   C = S^-1 + U^T Pinv U   (K1 x K1, symmetric, indefinite),
//     W = M^-1 lbs = (Pinv U) C^-1 [:, :Kh]                            (Woodbury; S^-1 E = E on the first Kh columns)
// Backward, for an upstream gW (V x Kh):   Z = M^-1 gW = Q - (Pinv U) C^-1 (U^T Q),  Q = Pinv gW,
//     d lbs = Z - Z (W^T lbs) - W (Z^T lbs).
//
// Arithmetic.  Scalar fp64 FMAs run at a small fraction of the fp32 rate on B200 (a first version of these kernels on DFMA took
// 25 us for the 642 x 642 x 32 product alone); the fp64 TENSOR cores do not: every product here is issued as
// mma.sync.m8n8k4.f64 (DMMA) on operand tiles staged in shared memory.  The (Kh+1)^2 inverse is a Gauss-Jordan elimination in
// fp32 (one SM, ordinary FFMA rate) polished by two Newton-Schulz steps X <- X + X (I - C X) in fp64 DMMA (the fp32 error
// ~1e-3 is squared twice).
//
// Data flow.  Row-parallel kernels take 8 rows of V per CTA; every reduction over V goes through per-block partial sums
// (32 rows per block) that the consumer adds up in block order — deterministic, no floating-point atomics:
//   hs_rows_kernel        PU = Pinv U          (the column Pinv 1 is a constant the caller supplies)
//   hs_gram_kernel        partials of U^T PU
//   hs_invert_kernel      C = S^-1 + sum; C^-1                                                       (one CTA)
//   hs_apply_kernel       W = PU Cinv[:, :Kh]  (fp64 + fp32)
//   hs_rows_kernel        Q = Pinv gW
//   hs_gram_kernel        partials of U^T Q
//   hs_small_kernel       T = Cinv (sum)                                                             (one CTA)
//   hs_apply_kernel       Z = Q - PU T
//   hs_gram_kernel x2     partials of lbs^T W (= A^T) and lbs^T Z (= B^T)
//   hs_grad_kernel        d lbs = Z - Z A - W B  (fp32; sums the partials of A^T, B^T itself)
// C = [[I + lbs^T Pinv lbs, b], [b^T, d]]: the leading Kh x Kh block is symmetric positive definite, so elimination in the
// natural order meets positive pivots there and the (non-zero) Schur complement of the last row at the end: no pivoting.
#include "common.cuh"

namespace {

constexpr int kRows = 8;          // rows of V per CTA of the row-parallel kernels
constexpr int kRowThreads = 128;  // 4 warps: the 8-column tiles of the output are dealt round-robin
constexpr int kGramRows = 32;     // rows of V per block of the reduction kernels
constexpr int kMaxKh = 64;        // the reference uses 16 / 32 / 64 handles

inline int hs_nblk(int V) { return (V + kRows - 1) / kRows; }
inline int hs_gblk(int V) { return (V + kGramRows - 1) / kGramRows; }
__host__ __device__ inline int pad8(int n) { return (n + 7) & ~7; }
// leading dimensions (in doubles) that keep the two DMMA operand patterns at the two-wavefront minimum of 64-bit shared loads:
// row-major A (lane -> [gi][k + tg]) wants ld = 4 mod 16, B / transposed A (lane -> [k + tg][gi]) wants ld = 8 mod 16
__host__ __device__ inline int ld_a(int n) { return ((n + 11) / 16) * 16 + 4; }   // smallest ld >= n with ld = 4 mod 16
__host__ __device__ inline int ld_b(int n) { return ((n + 7) / 16) * 16 + 8; }    // smallest ld >= n with ld = 8 mod 16

// D(8x8) += A(8x4) B(4x8) on the fp64 tensor cores.  Lane layout (gi = lane / 4, tg = lane % 4): a = A[gi][tg], b = B[tg][gi],
// c[0], c[1] = C[gi][2 tg], C[gi][2 tg + 1].
__device__ __forceinline__ void dmma(double (&c)[2], double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c[0]), "+d"(c[1]) : "d"(a), "d"(b));
}
// one 8x8 tile: C += A B over K (multiple of 4).  A row-major (AT = false: A(i,k) = As[i * lda + k]) or stored transposed
// (AT = true: A(i,k) = As[k * lda + i]); B row-major (BT = false: B(k,j) = Bs[k * ldb + j]) or transposed (B(k,j) = Bs[j * ldb + k]).
// As / Bs point at the tile's first row / column.
template <bool AT, bool BT>
__device__ __forceinline__ void tile_mma(const double* As, int lda, const double* Bs, int ldb, int K, double (&c)[2], int lane) {
  const int gi = lane >> 2, tg = lane & 3;
  for (int k = 0; k < K; k += 4) {
    const double a = AT ? As[(k + tg) * lda + gi] : As[gi * lda + k + tg];
    const double b = BT ? Bs[gi * ldb + k + tg] : Bs[(k + tg) * ldb + gi];
    dmma(c, a, b);
  }
}

// rows [r0, r0 + 8) of OUT = Pinv X, X (V x Kh) fp32 (lbs forward, grad_W backward); OUT has row stride ld (>= Kh); with
// `last` (forward) column Kh of OUT is set to last[r] (= Pinv 1).
// The operands of a chunk of k — the 8 row segments of Pinv (fp64) and the matching rows of X (fp32) — are brought into
// shared memory by the TMA unit in ONE round trip (cp.async.bulk + mbarrier; for V = 642, Kh = 32 the whole problem of the
// CTA, 41 KB + 82 KB, is one chunk), then 16 warps split the k steps: each keeps a private 8 x Kh accumulator (one DMMA per
// 8 columns per step, X converted on the fly) and the 16 partial tiles are added in warp order at the end.
constexpr int kRowsThreads = 512;
struct RowsSmem {
  int kc, off_x, off_red, total;  // chunk length (multiple of 4), byte offsets of the X chunk and of the reduction buffer
  __host__ __device__ RowsSmem(int V, int Kh) {
    const int Khp = pad8(Kh);
    int c = (160 * 1024) / (kRows * 8 + Kh * 4);
    c = min(c & ~3, (V + 3) & ~3);
    kc = c;
    off_x = 16 + kRows * kc * 8;
    const int x_bytes = ((kc * Kh * 4 + 15) / 16) * 16;
    off_red = off_x;  // the reduction buffer reuses the X chunk (dead by then)
    total = off_x + max(x_bytes, (kRowsThreads / 32) * kRows * Khp * 8);
  }
};

__global__ void __launch_bounds__(kRowsThreads) hs_rows_kernel(const double* __restrict__ Pinv, const float* __restrict__ X, int V, int Kh,
                                                               double* __restrict__ out, int ld, const double* __restrict__ last) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const RowsSmem L(V, Kh);
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw);
  double* As = reinterpret_cast<double*>(smem_raw + 16);       // [8][kc]   row segments of Pinv
  float* Xs = reinterpret_cast<float*>(smem_raw + L.off_x);    // [kc][Kh]
  double* red = reinterpret_cast<double*>(smem_raw + L.off_red);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, gi = lane >> 2, tg = lane & 3;
  const int r0 = blockIdx.x * kRows, nr = min(kRows, V - r0), Khp = pad8(Kh), ntile = Khp / 8, kc = L.kc;
  // bulk copies need 16-byte aligned addresses and sizes: true for even V / Kh % 4 == 0 / aligned tensors; else plain loads
  const bool bulk = ((V & 1) == 0) && ((Kh & 3) == 0) && ((((uintptr_t)Pinv) | ((uintptr_t)X)) & 15u) == 0;
  if (tid == 0) { mbar_init(bar, 1); mbar_fence_init(); }
  __syncthreads();
  double acc[kMaxKh / 8][2];
#pragma unroll
  for (int n = 0; n < kMaxKh / 8; ++n) acc[n][0] = acc[n][1] = 0.0;
  uint32_t phase = 0;
  for (int k0 = 0; k0 < V; k0 += kc) {
    const int nk = min(kc, V - k0), nk4 = (nk + 3) & ~3;
    if (bulk) {
      if (tid == 0) {
        const uint32_t rb = (uint32_t)nk * 8u, xb = (uint32_t)nk * (uint32_t)Kh * 4u;  // (nk even: k0, V even)
        mbar_expect_tx(bar, rb * (uint32_t)nr + xb);
        for (int r = 0; r < nr; ++r) tma_bulk_g2s(As + (size_t)r * kc, Pinv + (size_t)(r0 + r) * V + k0, rb, bar);
        tma_bulk_g2s(Xs, X + (size_t)k0 * Kh, xb, bar);
      }
    } else {
      for (int e = tid; e < nr * nk; e += kRowsThreads) As[(e / nk) * kc + (e % nk)] = Pinv[(size_t)(r0 + e / nk) * V + k0 + (e % nk)];
      for (int e = tid; e < nk * Kh; e += kRowsThreads) Xs[e] = X[(size_t)k0 * Kh + e];
    }
    // zero what the copies do not cover: rows past V, the k tail of the last chunk (up to 3 columns)
    for (int e = tid; e < kRows * (nk4 - nk); e += kRowsThreads) As[(e / (nk4 - nk)) * kc + nk + e % (nk4 - nk)] = 0.0;
    for (int e = tid; e < (kRows - nr) * nk4; e += kRowsThreads) As[(nr + e / nk4) * kc + e % nk4] = 0.0;
    for (int e = tid; e < (nk4 - nk) * Kh; e += kRowsThreads) Xs[nk * Kh + e] = 0.0f;
    if (bulk) mbar_wait(bar, phase);
    phase ^= 1u;
    __syncthreads();
    for (int k = 4 * warp; k < nk4; k += 4 * (kRowsThreads / 32)) {
      const double a = As[gi * kc + k + tg];
      const float* xr = Xs + (size_t)(k + tg) * Kh + gi;
#pragma unroll
      for (int n = 0; n < kMaxKh / 8; ++n)
        if (n < ntile) dmma(acc[n], a, (8 * n + gi < Kh) ? (double)xr[8 * n] : 0.0);
    }
    __syncthreads();  // the chunk is consumed: the next copies (or the reduction buffer) may overwrite it
  }
  // 16 partial tiles -> one, in warp order
#pragma unroll
  for (int n = 0; n < kMaxKh / 8; ++n)
    if (n < ntile) {
      red[((size_t)warp * kRows + gi) * Khp + 8 * n + 2 * tg] = acc[n][0];
      red[((size_t)warp * kRows + gi) * Khp + 8 * n + 2 * tg + 1] = acc[n][1];
    }
  __syncthreads();
  for (int e = tid; e < kRows * Khp; e += kRowsThreads) {
    const int r = e / Khp, c = e - r * Khp;
    if (r < nr && c < Kh) {
      double s = 0.0;
#pragma unroll
      for (int w = 0; w < kRowsThreads / 32; ++w) s += red[(size_t)w * kRows * Khp + e];
      out[(size_t)(r0 + r) * ld + c] = s;
    }
  }
  if (last && tid < nr) out[(size_t)(r0 + tid) * ld + Kh] = last[r0 + tid];
}

// part[blk][a][b] = sum over the block's 32 rows r of U[r][a] * R[r][b]:  U = [lbs, 1] (nL = Kh + 1) or lbs (nL = Kh), fp32;
// R (V x nR, row stride ldR) fp64.
__global__ void __launch_bounds__(512) hs_gram_kernel(const float* __restrict__ lbs, int Kh, int nL, const double* __restrict__ R, int ldR,
                                                      int nR, int V, double* __restrict__ part) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int nLp = pad8(nL), nRp = pad8(nR), ldu = ld_b(nLp), ldr = ld_b(nRp);
  double* Us = reinterpret_cast<double*>(smem_raw);  // [32][ldu]
  double* Rs = Us + (size_t)kGramRows * ldu;          // [32][ldr]
  const int r0 = blockIdx.x * kGramRows, nr = min(kGramRows, V - r0), tid = threadIdx.x, nt = blockDim.x;
  for (int e = tid; e < kGramRows * nLp; e += nt) {
    const int r = e / nLp, a = e - r * nLp;
    Us[r * ldu + a] = (r < nr && a < nL) ? (a < Kh ? (double)lbs[(size_t)(r0 + r) * Kh + a] : 1.0) : 0.0;
  }
  for (int e = tid; e < kGramRows * nRp; e += nt) {
    const int r = e / nRp, b = e - r * nRp;
    Rs[r * ldr + b] = (r < nr && b < nR) ? R[(size_t)(r0 + r) * ldR + b] : 0.0;
  }
  __syncthreads();
  double* pb = part + (size_t)blockIdx.x * nL * nR;
  const int lane = tid & 31, warp = tid >> 5, nw = nt >> 5, tn = nRp / 8, gi = lane >> 2, tg = lane & 3;
  for (int t = warp; t < (nLp / 8) * tn; t += nw) {
    const int ti = t / tn, tj = t - ti * tn;
    double c[2] = {0.0, 0.0};
    tile_mma<true, false>(Us + 8 * ti, ldu, Rs + 8 * tj, ldr, kGramRows, c, lane);
    const int a = 8 * ti + gi, b = 8 * tj + 2 * tg;
    if (a < nL && b < nR) pb[a * nR + b] = c[0];
    if (a < nL && b + 1 < nR) pb[a * nR + b + 1] = c[1];
  }
}

// the block partials of one element, added in block order (independent loads, four at a time)
__device__ __forceinline__ double hs_sum_parts(const double* __restrict__ part, int nblk, size_t stride, int e) {
  double s = 0.0;
  int b = 0;
  for (; b + 3 < nblk; b += 4) {
    const double p0 = part[(size_t)b * stride + e], p1 = part[(size_t)(b + 1) * stride + e], p2 = part[(size_t)(b + 2) * stride + e],
                 p3 = part[(size_t)(b + 3) * stride + e];
    s = (((s + p0) + p1) + p2) + p3;
  }
  for (; b < nblk; ++b) s += part[(size_t)b * stride + e];
  return s;
}

// D = alpha * A B + beta * Cin over square matrices padded to np (multiple of 8, np <= 72), all in shared memory with leading
// dimension ld; the CTA's 32 warps share the tiles (at most three each).  Every product is finished before anything is written
// (barrier inside), so D may alias A, B or Cin.  Cin may be nullptr.
__device__ void cta_matmul(const double* A, const double* B, const double* Cin, double* D, int np, int ld, double alpha, double beta,
                           int tid) {
  const int lane = tid & 31, warp = tid >> 5, tn = np / 8, gi = lane >> 2, tg = lane & 3;
  double c[3][2];
#pragma unroll
  for (int q = 0; q < 3; ++q) {
    const int t = warp + 32 * q;
    c[q][0] = c[q][1] = 0.0;
    if (t < tn * tn) {
      const int ti = t / tn, tj = t - ti * tn;
      tile_mma<false, false>(A + 8 * ti * ld, ld, B + 8 * tj, ld, np, c[q], lane);
      const int i = 8 * ti + gi, j = 8 * tj + 2 * tg;
      c[q][0] = alpha * c[q][0] + (Cin ? beta * Cin[i * ld + j] : 0.0);
      c[q][1] = alpha * c[q][1] + (Cin ? beta * Cin[i * ld + j + 1] : 0.0);
    }
  }
  __syncthreads();
#pragma unroll
  for (int q = 0; q < 3; ++q) {
    const int t = warp + 32 * q;
    if (t < tn * tn) {
      const int ti = t / tn, tj = t - ti * tn, i = 8 * ti + gi, j = 8 * tj + 2 * tg;
      D[i * ld + j] = c[q][0];
      D[i * ld + j + 1] = c[q][1];
    }
  }
  __syncthreads();
}

// C = S^-1 + sum of the partials; X0 = C^-1 by Gauss-Jordan in fp32 (two barriers per pivot: every thread reads the old
// entries it needs, then writes the new ones); two Newton-Schulz steps X <- X + X (I - C X) in fp64 on the tensor cores.
// One CTA of 1024 threads.  info[0] = 1 if a pivot is not positive where it must be / vanishes, or the polished inverse
// leaves a residual |I - C X| > 1e-6 (the caller can fall back to the direct solve).
constexpr int kInvPer = 10;  // rows per thread in the elimination: 1024 / (2 K1) column groups, K1 rows: 3 for 32 handles, 10 for 64
__global__ void __launch_bounds__(1024) hs_invert_kernel(const double* __restrict__ part, int nblk, int K1, double sinv_last,
                                                         double* __restrict__ Cinv, int* __restrict__ info) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int np = pad8(K1), ld = ld_b(np), W2 = 2 * K1, tid = threadIdx.x, nt = 1024, Kh = K1 - 1;
  double* Cs = reinterpret_cast<double*>(smem_raw);  // [np][ld]  C, padded with the identity
  double* Xs = Cs + (size_t)np * ld;                 // [np][ld]  current inverse
  double* Rs = Xs + (size_t)np * ld;                 // [np][ld]  residual I - C X
  float* M = reinterpret_cast<float*>(Rs + (size_t)np * ld);  // [K1][2 K1]  fp32 elimination
  __shared__ int bad;
  if (tid == 0) bad = 0;
  for (int e = tid; e < np * np; e += nt) {
    const int i = e / np, j = e - i * np;
    double s = (i == j) ? 1.0 : 0.0;
    if (i < K1 && j < K1) {
      s = hs_sum_parts(part, nblk, (size_t)K1 * K1, i * K1 + j);
      if (i == j) s += (i == Kh) ? sinv_last : 1.0;
      M[i * W2 + j] = (float)s;
      M[i * W2 + K1 + j] = (i == j) ? 1.0f : 0.0f;
    }
    Cs[i * ld + j] = s;
  }
  __syncthreads();
  // thread -> one column j of [C | I] and every G-th row of it (G = threads per column group), fixed over the pivots: the scaled
  // pivot-row entry is loaded once per pivot, each entry costs two loads, one FMA and one store
  const int G = nt / W2, myj = tid % W2, myg = tid / W2;   // threads past G * W2 idle in the elimination
  const bool worker = myg < G;
  for (int k = 0; k < K1; ++k) {
    const float p = M[k * W2 + k];
    if (tid == 0 && !(k < Kh ? p > 0.0f : p != 0.0f)) bad = 1;
    float nv[kInvPer];
    if (worker) {
      const float rk = M[k * W2 + myj] * (1.0f / p);
#pragma unroll
      for (int q = 0; q < kInvPer; ++q) {
        const int i = myg + q * G;
        if (i < K1) nv[q] = (i == k) ? rk : fmaf(-M[i * W2 + k], rk, M[i * W2 + myj]);
      }
    }
    __syncthreads();
    if (worker) {
#pragma unroll
      for (int q = 0; q < kInvPer; ++q) {
        const int i = myg + q * G;
        if (i < K1) M[i * W2 + myj] = nv[q];
      }
    }
    __syncthreads();
  }
  for (int e = tid; e < np * np; e += nt) {
    const int i = e / np, j = e - i * np;
    Xs[i * ld + j] = (i < K1 && j < K1) ? (double)M[i * W2 + K1 + j] : (i == j ? 1.0 : 0.0);
  }
  __syncthreads();
  for (int it = 0; it < 3; ++it) {
    cta_matmul(Cs, Xs, nullptr, Rs, np, ld, -1.0, 0.0, tid);      // R = -C X
    for (int i = tid; i < np; i += nt) Rs[i * ld + i] += 1.0;       // R = I - C X
    __syncthreads();
    if (it == 2) break;                                              // third pass: the residual of the polished inverse only
    cta_matmul(Xs, Rs, Xs, Xs, np, ld, 1.0, 1.0, tid);             // X = X + X R
  }
  double worst = 0.0;
  for (int e = tid; e < K1 * K1; e += nt) worst = fmax(worst, fabs(Rs[(e / K1) * ld + (e % K1)]));
  if (!(worst <= 1e-6)) bad = 1;   // (benign race: every writer stores the same value)
  for (int e = tid; e < K1 * K1; e += nt) Cinv[e] = Xs[(e / K1) * ld + (e % K1)];
  __syncthreads();
  if (tid == 0) info[0] = bad;
}

// OUT = Cin - A B (negate) or OUT = A B for this CTA's 8 rows: A = rows of PU (V x K1), B (K1 x Kh) = Cinv[:, :Kh] (row
// stride ldB) or T.  fp64 result, optionally also fp32.
__global__ void __launch_bounds__(kRowThreads) hs_apply_kernel(const double* __restrict__ PU, const double* __restrict__ B, int ldB,
                                                               const double* __restrict__ Cin, int negate, int V, int Kh,
                                                               double* __restrict__ out64, float* __restrict__ out32) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int K1 = Kh + 1, Kp = pad8(K1), Khp = pad8(Kh), lda = ld_a(Kp), ldb = ld_b(Khp);
  double* As = reinterpret_cast<double*>(smem_raw);  // [8][lda]   PU rows, K padded with zeros
  double* Bs = As + kRows * lda;                     // [Kp][ldb]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, r0 = blockIdx.x * kRows;
  for (int e = tid; e < kRows * Kp; e += kRowThreads) {
    const int r = e / Kp, k = e - r * Kp;
    As[r * lda + k] = (r0 + r < V && k < K1) ? PU[(size_t)(r0 + r) * K1 + k] : 0.0;
  }
  for (int e = tid; e < Kp * Khp; e += kRowThreads) {
    const int k = e / Khp, c = e - k * Khp;
    Bs[k * ldb + c] = (k < K1 && c < Kh) ? B[(size_t)k * ldB + c] : 0.0;
  }
  __syncthreads();
  const int gi = lane >> 2, tg = lane & 3, r = r0 + gi;
  for (int n = warp; 8 * n < Khp; n += 4) {
    double c[2] = {0.0, 0.0};
    tile_mma<false, false>(As, lda, Bs + 8 * n, ldb, Kp, c, lane);
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int col = 8 * n + 2 * tg + h;
      if (r < V && col < Kh) {
        const double v = negate ? Cin[(size_t)r * Kh + col] - c[h] : c[h];
        out64[(size_t)r * Kh + col] = v;
        if (out32) out32[(size_t)r * Kh + col] = (float)v;
      }
    }
  }
}

// T = Cinv (sum of the partials of U^T Q).  One CTA.
__global__ void __launch_bounds__(512) hs_small_kernel(const double* __restrict__ Gpart, int nblk, int Kh, const double* __restrict__ Cinv,
                                                       double* __restrict__ T) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int K1 = Kh + 1, Kp = pad8(K1), Khp = pad8(Kh), lda = ld_a(Kp), ldb = ld_b(Khp), tid = threadIdx.x, nt = blockDim.x;
  double* As = reinterpret_cast<double*>(smem_raw);  // [Kp][lda]  Cinv
  double* Bs = As + (size_t)Kp * lda;                // [Kp][ldb]  G
  for (int e = tid; e < Kp * Kp; e += nt) {
    const int i = e / Kp, j = e - i * Kp;
    As[i * lda + j] = (i < K1 && j < K1) ? Cinv[i * K1 + j] : 0.0;
  }
  for (int e = tid; e < Kp * Khp; e += nt) {
    const int k = e / Khp, c = e - k * Khp;
    Bs[k * ldb + c] = (k < K1 && c < Kh) ? hs_sum_parts(Gpart, nblk, (size_t)K1 * Kh, k * Kh + c) : 0.0;
  }
  __syncthreads();
  const int lane = tid & 31, warp = tid >> 5, nw = nt >> 5, tn = Khp / 8, gi = lane >> 2, tg = lane & 3;
  for (int t = warp; t < (Kp / 8) * tn; t += nw) {
    const int ti = t / tn, tj = t - ti * tn;
    double c[2] = {0.0, 0.0};
    tile_mma<false, false>(As + 8 * ti * lda, lda, Bs + 8 * tj, ldb, Kp, c, lane);
    const int i = 8 * ti + gi, j = 8 * tj + 2 * tg;
    if (i < K1 && j < Kh) T[i * Kh + j] = c[0];
    if (i < K1 && j + 1 < Kh) T[i * Kh + j + 1] = c[1];
  }
}

// d lbs = Z - Z A - W B for the CTA's 32 rows (fp32 out); A^T = lbs^T W and B^T = lbs^T Z come as block partials and are
// added up here, in block order.  Both products in one pass: [Z | W] (32 x 2 Kh) times [A ; B] (2 Kh x Kh).
__global__ void __launch_bounds__(512) hs_grad_kernel(const double* __restrict__ Z, const double* __restrict__ W64,
                                                      const double* __restrict__ ATpart, const double* __restrict__ BTpart, int nblk, int V,
                                                      int Kh, float* __restrict__ g) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int Khp = pad8(Kh), K2 = 2 * Khp, lda = ld_a(K2), ldt = ld_a(K2), tid = threadIdx.x, nt = blockDim.x;
  double* As = reinterpret_cast<double*>(smem_raw);  // [32][lda]   rows of [Z | W]
  double* Ts = As + (size_t)kGramRows * lda;         // [Khp][ldt]  Ts[j][i] = A[i][j], Ts[j][Khp + i] = B[i][j]  (B operand, transposed)
  const int r0 = blockIdx.x * kGramRows, nr = min(kGramRows, V - r0);
  for (int e = tid; e < Khp * Khp; e += nt) {
    const int j = e / Khp, i = e - j * Khp;
    const bool in = j < Kh && i < Kh;
    Ts[j * ldt + i] = in ? hs_sum_parts(ATpart, nblk, (size_t)Kh * Kh, j * Kh + i) : 0.0;
    Ts[j * ldt + Khp + i] = in ? hs_sum_parts(BTpart, nblk, (size_t)Kh * Kh, j * Kh + i) : 0.0;
  }
  for (int e = tid; e < kGramRows * Khp; e += nt) {
    const int r = e / Khp, c = e - r * Khp;
    const bool in = r < nr && c < Kh;
    As[r * lda + c] = in ? Z[(size_t)(r0 + r) * Kh + c] : 0.0;
    As[r * lda + Khp + c] = in ? W64[(size_t)(r0 + r) * Kh + c] : 0.0;
  }
  __syncthreads();
  const int lane = tid & 31, warp = tid >> 5, nw = nt >> 5, tn = Khp / 8, gi = lane >> 2, tg = lane & 3;
  for (int t = warp; t < (kGramRows / 8) * tn; t += nw) {
    const int ti = t / tn, tj = t - ti * tn;
    double c[2] = {0.0, 0.0};
    tile_mma<false, true>(As + 8 * ti * lda, lda, Ts + 8 * tj * ldt, ldt, K2, c, lane);
    const int r = 8 * ti + gi, j = 8 * tj + 2 * tg;
    if (r < nr && j < Kh) g[(size_t)(r0 + r) * Kh + j] = (float)(As[r * lda + j] - c[0]);
    if (r < nr && j + 1 < Kh) g[(size_t)(r0 + r) * Kh + j + 1] = (float)(As[r * lda + j + 1] - c[1]);
  }
}

// workspace layout (doubles): PU | Cpart | Cinv | W64 | Q | Gpart | T | Z | ATpart | BTpart | info(int, padded)
struct HsLayout {
  size_t PU, Cpart, Cinv, W64, Q, Gpart, T, Z, ATpart, BTpart, info, total;
  HsLayout(int V, int Kh) {
    const size_t K1 = Kh + 1, nb = hs_gblk(V);
    size_t o = 0;
    PU = o; o += (size_t)V * K1;
    Cpart = o; o += nb * K1 * K1;
    Cinv = o; o += K1 * K1;
    W64 = o; o += (size_t)V * Kh;
    Q = o; o += (size_t)V * Kh;
    Gpart = o; o += nb * K1 * Kh;
    T = o; o += K1 * Kh;
    Z = o; o += (size_t)V * Kh;
    ATpart = o; o += nb * Kh * Kh;
    BTpart = o; o += nb * Kh * Kh;
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

size_t sm_gram(int nL, int nR) { return (size_t)kGramRows * (ld_b(pad8(nL)) + ld_b(pad8(nR))) * sizeof(double); }
size_t sm_apply(int Kh) { return ((size_t)kRows * ld_a(pad8(Kh + 1)) + (size_t)pad8(Kh + 1) * ld_b(pad8(Kh))) * sizeof(double); }

}  // namespace

extern "C" int64_t acfm_handle_solve_workspace_bytes(int V, int Kh) {
  if (V <= 0 || Kh <= 0) return 0;
  return (int64_t)(HsLayout(V, Kh).total * sizeof(double));
}

extern "C" int acfm_handle_solve_fwd(const double* Pinv, const double* Pinv_ones, const float* lbs, int V, int Kh, double c_over_V,
                                     float* W, void* workspace, int64_t workspace_bytes, void* stream) {
  ACFM_REQUIRE(V > 0 && Kh > 0 && Kh <= kMaxKh, ACFM_ERR_BAD_ARG, "acfm_handle_solve_fwd: bad sizes V=%d Kh=%d (Kh <= 64)", V, Kh);
  ACFM_REQUIRE(Pinv && Pinv_ones && lbs && W && workspace, ACFM_ERR_BAD_ARG, "acfm_handle_solve_fwd: null pointer");
  ACFM_REQUIRE(c_over_V > 0.0, ACFM_ERR_BAD_ARG, "acfm_handle_solve_fwd: c_over_V must be > 0");
  const HsLayout L(V, Kh);
  ACFM_REQUIRE(workspace_bytes >= (int64_t)(L.total * sizeof(double)) && (((uintptr_t)workspace) & 15u) == 0, ACFM_ERR_BAD_ARG,
               "acfm_handle_solve_fwd: workspace must be 16-byte aligned and hold acfm_handle_solve_workspace_bytes()");
  double* ws = (double*)workspace;
  cudaStream_t st = (cudaStream_t)stream;
  const int K1 = Kh + 1, nb = hs_nblk(V), gb = hs_gblk(V);
  static std::atomic<int> s_rows[kAcfmMaxDevices], s_gram[kAcfmMaxDevices], s_inv[kAcfmMaxDevices], s_app[kAcfmMaxDevices];
  const size_t sm_rows = (size_t)RowsSmem(V, Kh).total;
  if (int rc = hs_smem(hs_rows_kernel, sm_rows, s_rows, "acfm_handle_solve_fwd")) return rc;
  hs_rows_kernel<<<nb, kRowsThreads, sm_rows, st>>>(Pinv, lbs, V, Kh, ws + L.PU, K1, Pinv_ones);
  ACFM_LAUNCH_OK("hs_rows_kernel");
  if (int rc = hs_smem(hs_gram_kernel, sm_gram(K1, K1), s_gram, "acfm_handle_solve_fwd")) return rc;
  hs_gram_kernel<<<gb, 512, sm_gram(K1, K1), st>>>(lbs, Kh, K1, ws + L.PU, K1, K1, V, ws + L.Cpart);
  ACFM_LAUNCH_OK("hs_gram_kernel");
  const int np = pad8(K1);
  const size_t sm_inv = 3 * (size_t)np * ld_b(np) * sizeof(double) + (size_t)K1 * 2 * K1 * sizeof(float);
  if (int rc = hs_smem(hs_invert_kernel, sm_inv, s_inv, "acfm_handle_solve_fwd")) return rc;
  hs_invert_kernel<<<1, 1024, sm_inv, st>>>(ws + L.Cpart, gb, K1, -1.0 / c_over_V, ws + L.Cinv, (int*)(ws + L.info));
  ACFM_LAUNCH_OK("hs_invert_kernel");
  if (int rc = hs_smem(hs_apply_kernel, sm_apply(Kh), s_app, "acfm_handle_solve_fwd")) return rc;
  hs_apply_kernel<<<nb, kRowThreads, sm_apply(Kh), st>>>(ws + L.PU, ws + L.Cinv, K1, nullptr, 0, V, Kh, ws + L.W64, W);
  ACFM_LAUNCH_OK("hs_apply_kernel");
  return ACFM_OK;
}

extern "C" int acfm_handle_solve_bwd(const double* Pinv, const float* lbs, const float* grad_W, int V, int Kh, float* grad_lbs,
                                     void* workspace, int64_t workspace_bytes, void* stream) {
  ACFM_REQUIRE(V > 0 && Kh > 0 && Kh <= kMaxKh, ACFM_ERR_BAD_ARG, "acfm_handle_solve_bwd: bad sizes V=%d Kh=%d (Kh <= 64)", V, Kh);
  ACFM_REQUIRE(Pinv && lbs && grad_W && grad_lbs && workspace, ACFM_ERR_BAD_ARG, "acfm_handle_solve_bwd: null pointer");
  const HsLayout L(V, Kh);
  ACFM_REQUIRE(workspace_bytes >= (int64_t)(L.total * sizeof(double)) && (((uintptr_t)workspace) & 15u) == 0, ACFM_ERR_BAD_ARG,
               "acfm_handle_solve_bwd: workspace must be the one acfm_handle_solve_fwd filled");
  double* ws = (double*)workspace;
  cudaStream_t st = (cudaStream_t)stream;
  const int K1 = Kh + 1, nb = hs_nblk(V), gb = hs_gblk(V);
  static std::atomic<int> s_rows[kAcfmMaxDevices], s_gram[kAcfmMaxDevices], s_small[kAcfmMaxDevices], s_app[kAcfmMaxDevices],
      s_g[kAcfmMaxDevices];
  const size_t sm_rows = (size_t)RowsSmem(V, Kh).total;
  if (int rc = hs_smem(hs_rows_kernel, sm_rows, s_rows, "acfm_handle_solve_bwd")) return rc;
  hs_rows_kernel<<<nb, kRowsThreads, sm_rows, st>>>(Pinv, grad_W, V, Kh, ws + L.Q, Kh, nullptr);
  ACFM_LAUNCH_OK("hs_rows_kernel");
  if (int rc = hs_smem(hs_gram_kernel, sm_gram(K1, K1), s_gram, "acfm_handle_solve_bwd")) return rc;
  hs_gram_kernel<<<gb, 512, sm_gram(K1, Kh), st>>>(lbs, Kh, K1, ws + L.Q, Kh, Kh, V, ws + L.Gpart);
  ACFM_LAUNCH_OK("hs_gram_kernel");
  const size_t sm_small = ((size_t)pad8(K1) * ld_a(pad8(K1)) + (size_t)pad8(K1) * ld_b(pad8(Kh))) * sizeof(double);
  if (int rc = hs_smem(hs_small_kernel, sm_small, s_small, "acfm_handle_solve_bwd")) return rc;
  hs_small_kernel<<<1, 512, sm_small, st>>>(ws + L.Gpart, gb, Kh, ws + L.Cinv, ws + L.T);
  ACFM_LAUNCH_OK("hs_small_kernel");
  if (int rc = hs_smem(hs_apply_kernel, sm_apply(Kh), s_app, "acfm_handle_solve_bwd")) return rc;
  hs_apply_kernel<<<nb, kRowThreads, sm_apply(Kh), st>>>(ws + L.PU, ws + L.T, Kh, ws + L.Q, 1, V, Kh, ws + L.Z, nullptr);
  ACFM_LAUNCH_OK("hs_apply_kernel");
  hs_gram_kernel<<<gb, 512, sm_gram(Kh, Kh), st>>>(lbs, Kh, Kh, ws + L.W64, Kh, Kh, V, ws + L.ATpart);
  hs_gram_kernel<<<gb, 512, sm_gram(Kh, Kh), st>>>(lbs, Kh, Kh, ws + L.Z, Kh, Kh, V, ws + L.BTpart);
  ACFM_LAUNCH_OK("hs_gram_kernel");
  const size_t sm_g = ((size_t)kGramRows + pad8(Kh)) * ld_a(2 * pad8(Kh)) * sizeof(double);
  if (int rc = hs_smem(hs_grad_kernel, sm_g, s_g, "acfm_handle_solve_bwd")) return rc;
  hs_grad_kernel<<<gb, 512, sm_g, st>>>(ws + L.Z, ws + L.W64, ws + L.ATpart, ws + L.BTpart, gb, V, Kh, grad_lbs);
  ACFM_LAUNCH_OK("hs_grad_kernel");
  return ACFM_OK;
}

// 1 if the last acfm_handle_solve_fwd on this workspace met a bad pivot or left a residual above 1e-6 (copies one int back:
// synchronises the stream; for tests and for the one-time check after construction, not for the per-step path)
extern "C" int acfm_handle_solve_singular(const void* workspace, int V, int Kh, void* stream) {
  const HsLayout L(V, Kh);
  int h = 0;
  if (cudaMemcpyAsync(&h, (const double*)workspace + L.info, sizeof(int), cudaMemcpyDeviceToHost, (cudaStream_t)stream) != cudaSuccess) return -1;
  if (cudaStreamSynchronize((cudaStream_t)stream) != cudaSuccess) return -1;
  return h;
}
