// targets.cu — per-step target maps from the ground-truth masks: exact Euclidean distance transforms, the barrier
// map and the boundary point lists.
//
// Replaces utils/image.py compute_dt / compute_dt_barrier / compute_boundaries
// (/root/reference/multiframe/utils/image.py:94-146; same file in monocular/), which ShapeTrainer.set_input runs on the
// CPU for every mask of every step (scipy.ndimage.distance_transform_edt twice per mask + skimage find_boundaries,
// multiframe/main.py:364-377) before copying the results to the GPU.  SURVEY.md §8f rank 1.
//
// EDT: exact, integer arithmetic.  Pass 1 (one thread per column) records for every pixel the squared vertical
// distance to the nearest FEATURE pixel of its column; pass 2 (one CTA per row, the row's values in shared memory)
// minimises (x-x')^2 + g^2[x'] searching outwards from x and stopping as soon as (x-x')^2 alone exceeds the best
// value.  scipy semantics: distance_transform_edt(a) is, for every non-zero element of a, the distance to the
// nearest zero element; dist_out = edt(1 - mask) (features: mask == 1), dist_in = edt(mask) (features: mask == 0).
// A map without any feature reproduces scipy's behaviour for that degenerate input (distances to the virtual
// point (row -1, col 0)).  Squared distances are < 2^24, so sqrt in fp64 and rounding to fp32 equals what the
// reference obtains with `torch.tensor(float64 array).float()`.
#include "common.cuh"

namespace {

constexpr int kThreads = 256;
constexpr int kInf = 1 << 29;

// g2[which][n][y][x] = squared vertical distance to the nearest feature in column x, kInf if none.
// which 0: features mask == 1 (for dist_out), which 1: features mask == 0 (for dist_in)
__global__ void __launch_bounds__(kThreads) edt_cols_kernel(const float* __restrict__ masks, int NB, int H, int W, int* __restrict__ g2) {
  const int x = blockIdx.x * kThreads + threadIdx.x;
  const int n = blockIdx.y;
  if (x >= W) return;
  const float* m = masks + (size_t)n * H * W + x;
  int* o0 = g2 + (size_t)n * H * W + x;
  int* o1 = g2 + ((size_t)NB + n) * H * W + x;
  int d0 = -1, d1 = -1;  // distance to the last feature seen going down (-1: none yet)
  for (int y = 0; y < H; ++y) {
    const float v = m[(size_t)y * W];
    d0 = (v == 1.0f) ? 0 : (d0 < 0 ? -1 : d0 + 1);
    d1 = (v == 0.0f) ? 0 : (d1 < 0 ? -1 : d1 + 1);
    o0[(size_t)y * W] = d0 < 0 ? kInf : d0 * d0;
    o1[(size_t)y * W] = d1 < 0 ? kInf : d1 * d1;
  }
  d0 = d1 = -1;
  for (int y = H - 1; y >= 0; --y) {
    const float v = m[(size_t)y * W];
    d0 = (v == 1.0f) ? 0 : (d0 < 0 ? -1 : d0 + 1);
    d1 = (v == 0.0f) ? 0 : (d1 < 0 ? -1 : d1 + 1);
    if (d0 >= 0) o0[(size_t)y * W] = min(o0[(size_t)y * W], d0 * d0);
    if (d1 >= 0) o1[(size_t)y * W] = min(o1[(size_t)y * W], d1 * d1);
  }
}

__device__ __forceinline__ int row_min(const int* g, int W, int x) {
  int best = g[x];
  for (int r = 1; r < W; ++r) {
    const int r2 = r * r;
    if (r2 >= best) break;
    if (x - r >= 0) best = min(best, r2 + g[x - r]);
    if (x + r < W) best = min(best, r2 + g[x + r]);
  }
  return best;
}

// grid (H, NB); dynamic shared memory 2*W ints
__global__ void __launch_bounds__(kThreads) edt_rows_kernel(const float* __restrict__ masks, const int* __restrict__ g2, int NB, int H,
                                                            int W, float k, int norm, float* __restrict__ edt_out,
                                                            float* __restrict__ barrier_out) {
  extern __shared__ int sg[];
  const int y = blockIdx.x, n = blockIdx.y;
  const int* r0 = g2 + ((size_t)n * H + y) * W;
  const int* r1 = g2 + (((size_t)NB + n) * H + y) * W;
  for (int x = threadIdx.x; x < W; x += kThreads) { sg[x] = r0[x]; sg[W + x] = r1[x]; }
  __syncthreads();
  const double size = (double)max(H, W);
  for (int x = threadIdx.x; x < W; x += kThreads) {
    int b0 = row_min(sg, W, x);
    int b1 = barrier_out ? row_min(sg + W, W, x) : 0;
    // no feature anywhere: scipy's result for an input without zeros is the distance to (row -1, col 0)
    if (b0 >= kInf) b0 = (y + 1) * (y + 1) + x * x;
    if (b1 >= kInf) b1 = (y + 1) * (y + 1) + x * x;
    const float v = masks[((size_t)n * H + y) * W + x];
    // distance_transform_edt is 0 on the zero elements of its input: 1 - mask is zero where mask == 1, mask where mask == 0
    const double dout = (v == 1.0f) ? 0.0 : sqrt((double)b0);
    const size_t o = ((size_t)n * H + y) * W + x;
    if (edt_out) edt_out[o] = (float)(norm ? dout / size : dout);
    if (barrier_out) {
      const double din = (v == 0.0f) ? 0.0 : sqrt((double)b1);
      const double diff = (dout - din) / size;
      barrier_out[o] = (float)(1.0 / (1.0 + exp((double)k * -diff)));
    }
  }
}

// boundary pixel (skimage find_boundaries, mode 'thick', connectivity 1): grey_dilation != grey_erosion over the cross
// footprint with scipy's default 'reflect' border, i.e. the 4 neighbours clamped to the image
__device__ __forceinline__ bool is_boundary(const float* m, int H, int W, int y, int x) {
  const float c = m[(size_t)y * W + x];
  const float u = m[(size_t)max(y - 1, 0) * W + x], d = m[(size_t)min(y + 1, H - 1) * W + x];
  const float l = m[(size_t)y * W + max(x - 1, 0)], r = m[(size_t)y * W + min(x + 1, W - 1)];
  const float mx = fmaxf(fmaxf(fmaxf(c, u), fmaxf(d, l)), r), mn = fminf(fminf(fminf(c, u), fminf(d, l)), r);
  return mx != mn;
}

// one warp per (mask, row): row_counts[n][y]
__global__ void __launch_bounds__(kThreads) bd_count_kernel(const float* __restrict__ masks, int NB, int H, int W, int* __restrict__ row_counts) {
  const int row = blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5);
  if (row >= NB * H) return;
  const int n = row / H, y = row - n * H, lane = threadIdx.x & 31;
  const float* m = masks + (size_t)n * H * W;
  int c = 0;
  for (int x0 = 0; x0 < W; x0 += 32) {
    const int x = x0 + lane;
    c += __popc(__ballot_sync(0xffffffffu, x < W && is_boundary(m, H, W, y, x)));
  }
  if (lane == 0) row_counts[row] = c;
}

// one CTA per mask: exclusive scan of the row counts -> totals[n] and, in place, row offsets
__global__ void __launch_bounds__(kThreads) bd_scan_kernel(int H, int* __restrict__ row_counts, int* __restrict__ totals) {
  __shared__ int carry;
  __shared__ int wsum[kThreads / 32];
  const int n = blockIdx.x;
  int* rc = row_counts + (size_t)n * H;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int y0 = 0; y0 < H; y0 += kThreads) {
    const int y = y0 + threadIdx.x;
    const int v = y < H ? rc[y] : 0;
    int incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, incl, o);
      if ((threadIdx.x & 31) >= o) incl += t;
    }
    if ((threadIdx.x & 31) == 31) wsum[threadIdx.x >> 5] = incl;
    __syncthreads();
    int base = carry;
    for (int w = 0; w < (threadIdx.x >> 5); ++w) base += wsum[w];
    if (y < H) rc[y] = base + incl - v;
    __syncthreads();
    if (threadIdx.x == kThreads - 1) carry = base + incl;
    __syncthreads();
  }
  if (threadIdx.x == 0) totals[n] = carry;
}

// one warp per (mask, row): points in row-major order, (x, y) normalised as utils/image.py:139-143, flag 1; padding
// entries are (-1, -1, 0) — the reference normalises its zero padding too
__global__ void __launch_bounds__(kThreads) bd_write_kernel(const float* __restrict__ masks, const int* __restrict__ row_offsets,
                                                            const int* __restrict__ totals, int NB, int H, int W, int max_bd,
                                                            float* __restrict__ out) {
  const int row = blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5);
  if (row >= NB * H) return;
  const int n = row / H, y = row - n * H, lane = threadIdx.x & 31;
  const float* m = masks + (size_t)n * H * W;
  float* o = out + (size_t)n * max_bd * 3;
  int base = row_offsets[row];
  const float fy = (float)(((double)y / (double)H - 0.5) * 2.0);
  for (int x0 = 0; x0 < W; x0 += 32) {
    const int x = x0 + lane;
    const bool b = x < W && is_boundary(m, H, W, y, x);
    const unsigned bal = __ballot_sync(0xffffffffu, b);
    if (b) {
      const int i = base + __popc(bal & ((1u << lane) - 1u));
      if (i < max_bd) {
        o[(size_t)i * 3] = (float)(((double)x / (double)W - 0.5) * 2.0);
        o[(size_t)i * 3 + 1] = fy;
        o[(size_t)i * 3 + 2] = 1.0f;
      }
    }
    base += __popc(bal);
  }
  if (y == H - 1) {  // padding of this mask
    for (int i = totals[n] + lane; i < max_bd; i += 32) { o[(size_t)i * 3] = -1.0f; o[(size_t)i * 3 + 1] = -1.0f; o[(size_t)i * 3 + 2] = 0.0f; }
  }
}

}  // namespace

extern "C" int acfm_edt_fwd(const float* masks, int NB, int H, int W, float k, int norm, float* edt_out, float* barrier_out,
                            int* workspace, void* stream) {
  ACFM_REQUIRE(NB >= 0 && H > 0 && W > 0, ACFM_ERR_BAD_ARG, "acfm_edt_fwd: bad sizes");
  ACFM_REQUIRE(H <= 4096 && W <= 4096, ACFM_ERR_UNSUPPORTED, "acfm_edt_fwd: H=%d, W=%d must be <= 4096 (exact fp32 squared distances)", H, W);
  if (NB == 0 || (!edt_out && !barrier_out)) return ACFM_OK;
  ACFM_REQUIRE(masks && workspace, ACFM_ERR_BAD_ARG, "acfm_edt_fwd: null pointer (workspace must hold 2*NB*H*W ints)");
  ACFM_REQUIRE(NB <= 65535, ACFM_ERR_UNSUPPORTED, "acfm_edt_fwd: NB=%d > 65535", NB);
  cudaStream_t st = (cudaStream_t)stream;
  edt_cols_kernel<<<dim3((W + kThreads - 1) / kThreads, NB), kThreads, 0, st>>>(masks, NB, H, W, workspace);
  ACFM_LAUNCH_OK("edt_cols_kernel");
  edt_rows_kernel<<<dim3(H, NB), kThreads, 2 * W * sizeof(int), st>>>(masks, workspace, NB, H, W, k, norm, edt_out, barrier_out);
  ACFM_LAUNCH_OK("edt_rows_kernel");
  return ACFM_OK;
}

extern "C" int acfm_boundaries_count(const float* masks, int NB, int H, int W, int* row_offsets, int* totals, void* stream) {
  ACFM_REQUIRE(NB >= 0 && H > 0 && W > 0, ACFM_ERR_BAD_ARG, "acfm_boundaries_count: bad sizes");
  if (NB == 0) return ACFM_OK;
  ACFM_REQUIRE(masks && row_offsets && totals, ACFM_ERR_BAD_ARG, "acfm_boundaries_count: null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  const int rows = NB * H;
  bd_count_kernel<<<(rows + kThreads / 32 - 1) / (kThreads / 32), kThreads, 0, st>>>(masks, NB, H, W, row_offsets);
  ACFM_LAUNCH_OK("bd_count_kernel");
  bd_scan_kernel<<<NB, kThreads, 0, st>>>(H, row_offsets, totals);
  ACFM_LAUNCH_OK("bd_scan_kernel");
  return ACFM_OK;
}

extern "C" int acfm_boundaries_write(const float* masks, const int* row_offsets, const int* totals, int NB, int H, int W,
                                     int max_bd, float* out, void* stream) {
  ACFM_REQUIRE(NB >= 0 && H > 0 && W > 0 && max_bd >= 0, ACFM_ERR_BAD_ARG, "acfm_boundaries_write: bad sizes");
  if (NB == 0 || max_bd == 0) return ACFM_OK;
  ACFM_REQUIRE(masks && row_offsets && totals && out, ACFM_ERR_BAD_ARG, "acfm_boundaries_write: null pointer");
  const int rows = NB * H;
  bd_write_kernel<<<(rows + kThreads / 32 - 1) / (kThreads / 32), kThreads, 0, (cudaStream_t)stream>>>(masks, row_offsets, totals, NB, H, W, max_bd, out);
  ACFM_LAUNCH_OK("bd_write_kernel");
  return ACFM_OK;
}
