// raster_fwd.cu — tile-binned rasterizer forward with the fused soft-silhouette blend.
//
// Replaces PyTorch3D 0.3.0 rasterize_meshes (coarse + fine CUDA kernels / naive CPU loop) and
// SoftSilhouetteShader / sigmoid_alpha_blend as reached from NeuralRenderer.forward
// (/root/reference/multiframe/nnutils/nmr.py:143-200) and OF_NeuralRenderer.forward (:224-238).
// Semantics: SURVEY.md §9.2-9.5.  Every value that reaches an output (barycentrics, depth, distances,
// inside / blur / depth decisions) is computed in strict IEEE fp32 in the CPU reference's operator
// order, so pix_to_face, zbuf and dists are bit-identical to oracle/.
//
// Three kernels per call when the caller passes a workspace (the split path, further down in this file), all on the caller's
// stream:
//   raster_prep_kernel   one CTA per render: mesh bounding box -> live regions (filed heaviest first by face count, so the
//                        rasterizer's grid ends on light regions) and runs of empty regions;
//   raster_fill_kernel   the -1 padding of the empty regions (87% of all fragment bytes at the reference's workloads), written
//                        by the TMA unit with cp.async.bulk shared -> global stores from one one-warp CTA per SM;
//   raster_fwd_kernel    one CTA per LIVE (render, 32x32-pixel region), described below — dispatched beside the padding kernel
//                        by programmatic dependent launch as soon as the padding CTAs are in place.
// Without a workspace raster_fwd_kernel runs on every region and pads the empty ones itself.
//
// raster_fwd_kernel, one CTA per (render, 32x32-pixel region):
//   1. the render's vertices (V*12 B) are staged into shared memory by the TMA unit (cp.async.bulk +
//      mbarrier); without the split path a CTA whose region lies outside the blur-expanded bounding box of
//      the mesh takes the pure fill path at once (~75% of the regions of the reference workloads);
//   2. all F faces are culled against the region (face-level skips of §9.4 + blur-expanded bbox),
//      survivors compacted with warp ballots and bucketed front to back by depth;
//   3. the first `cap` region faces get a RECORD in shared memory, set up once per region: vertices,
//      barycentric denominator, the refined reciprocals shared by the exact divisions, and three
//      conservative edge equations (see below);
//   4. warps pull 8x4-pixel tiles (one pixel per lane) from a shared counter, heaviest tile first (faces per tile are
//      counted while the records are written).  Per tile:
//      (a) scan, lane = face: tile-vs-bbox and tile-vs-edge-equation culling, one ballot per 32 faces;
//      (b) filter, lane = pixel, face uniform: exact bbox test + conservative edge test (6 FFMA), the
//          survivors' record ids are pushed on a per-lane queue;
//      (c) evaluate, lane = pixel, face per lane: each lane pops ITS OWN queue, so lanes only spend
//          instructions on (pixel, face) pairs that are real candidates (the face-uniform evaluation
//          of (b)'s survivors would run at ~30% lane utilisation); the exact §9.4 test — left right after the depth when the
//          fragment can no longer enter the pixel's full set — then the fragment joins the lane's K-nearest SET (K = 20:
//          unsorted, append or replace-the-farthest, no shifting; other K: round 1's sorted list) in shared memory;
//      (d) one rank pass puts every lane's set in depth order (K = 20), the silhouette is blended in that order, the fused
//          mask-loss sums are accumulated, and pix_to_face / zbuf / dists are written with 16-byte stores.
//   Region faces beyond `cap` (regions that see more than 256 faces) go through a face-uniform path
//   with on-the-fly set-up broadcast by warp shuffles.
//   Template variants: KT = 20 the K-nearest sets, KT = 1 the hard renders (nearest fragment in registers, faces behind it
//   dropped at the filter / scan stage), KT = 0 every other K (sorted lists); LEAN (acfm_raster_fwd_lean): step (d) writes
//   compact fragments of the covered pixels to a scratch for the backward instead of the padded (N,H,W,K) tensors.
//
// Conservative edge equations.  For edge i (opposite vertex i) with numerator n_i(p) (w_i = n_i/den),
// a pixel with sign(den)*n_i(p) < 0 lies outside that edge's line, at distance |n_i|/|e_i| from it and
// therefore at least that far from the triangle.  g_i(p) = A_i x + B_i y + C_i is n_i normalised by
// -(|e_i| T), T = 1.001 sqrt(blur) + delta, and shrunk by a bound of its own rounding error, so that
// g_i(p) < -1 proves "not inside, and squared distance > blur" with margin for the rounding of the exact
// path: such pixels can never be fragments.  The test only REJECTS; every accepted pair goes through
// the exact arithmetic.
//
// HBM-bound on the API-mandated (N,H,W,K) fragment tensors: 16K+4 bytes written per pixel.
#include "common.cuh"
#include <algorithm>

#include "raster_common.cuh"

namespace {

static_assert(kRegion / kTileW == 4 && (kRegion / kTileW) * (kRegion / kTileH) == 32, "a region is 4 x 8 tiles: one lane / one bit each");
constexpr int kZBuckets = 64;  // depth buckets of the region face list (front-to-back evaluation order)
constexpr int kQueue = 32;     // per-lane candidate queue depth (record ids, 1 byte each)
constexpr int kQueue1 = 4;     // ... at which the K = 1 renders drain: their depth culls feed on fresh nearest depths

struct RasterParams {
  const float* ndc;
  const void* faces;
  long long faces_stride;  // elements between renders (0: shared topology)
  int N, V, F, H, W, K;
  float blur, sq_blur, sigma;
  float k_eps;  // kEpsilon (acfm_set_raster_epsilon)
  int clip, cull;
  long long* p2f;
  float* zbuf;
  float* dists;
  float* bary;
  float* mask;
  float* vis;   // optional (N,V): 1 for the vertices of every pixel's nearest face (zeroed by the host entry)
  int regions_x, regions_y;
  int cap;      // face records per region
  int vec_ok;   // output pointers are 16-byte aligned
  int bulk_ok;  // rows of the fragment tensors are 16-byte aligned: the padding kernel may use bulk stores
  int pad_balance;  // bytes of padding per live region above which every padding CTA works (else 5/8 of them)
  // fused silhouette losses (optional): targets read at render n % NB, per-(render, region) partial sums out
  const float* loss_target;  // (NB,H,W) or NULL
  const float* loss_edt;     // (NB,H,W) or NULL
  int NB;
  float* loss_part;          // (N * regions, 4), zeroed by the host entry
  // lean mode (acfm_raster_fwd_lean): no fragment tensors; the fragments of the live regions go, compact, to the caller's
  // scratch for the backward: [work-list slot][pixel of the region (1024)][K] face ids (0xffff: none) and signed distances
  unsigned short* lean_f;
  float* lean_d;
  const int* work;  // split path, written by raster_prep_kernel: {live, fill runs, empty regions, -, live per weight class [4],
                    //   listR[4][N*regions], listF[N*regions][2]}
};

// slots of the padding pattern: one region row at K, at least 640 (bulk stores below ~2.5 KB reach a third of the HBM write rate)
__host__ __device__ inline int fill_pattern_slots(int K) { return max(kRegion * K, 640); }

// which (K, CTA size) pairs run the K-nearest-set specialisation of the rasterizer kernel (KT = K); the others run the sorted
// lists (KT = 0) or, for K = 1, the register z-buffer (KT = 1: the slab arrays are unused)
__host__ __device__ inline bool fwd_sets(int K, int nwarps) { return K == 20 && nwarps == 8; }

// shared-memory carve-up, identical on host and device
struct FwdSmem {
  int off_ndc, off_red, off_hist, off_tw, off_ts, off_rlist, off_recA, off_recB, off_union, off_verts, off_tmp, warp_bytes, total;
  int w_z, w_d, w_f, w_o, w_q, w_cnt, w_ord;  // offsets inside one warp's slab
  int KS;                                     // per-lane row stride in entries
  __host__ __device__ FwdSmem(int V, int F, int K, int nwarps, int cap) {
    int o = 32;  // mbarrier + counters
    off_ndc = o; o += 2 * kRegion * 4;  // pixel-centre NDC coordinates of the region's columns and rows
    off_red = o; o += nwarps * 32;
    off_hist = o; o += kZBuckets * 4;
    off_tw = o; o += 32 * 4;  // faces per 8x4 tile of the region (tile order: heaviest first)
    off_ts = o; o += 32 * 4 * 4;  // fused losses: four partial sums per tile (one slot per tile: order-independent)
    off_rlist = o; o += ((F * 2 + 15) / 16) * 16;  // ushort per region face
    off_recA = o; o += cap * 64;                   // 4 float4 arrays (scan / filter data)
    off_recB = o; o += cap * 64;                   // 4 float4 arrays (exact evaluation data)
    off_union = o;
    off_verts = o;
    const int vb = ((V * 12 + 16 + 15) / 16) * 16;
    off_tmp = o + vb;
    // K-nearest sets (K = 20 on 8-warp CTAs, see fwd_sets()): rows of K entries padded to 16 bytes of depth words with an
    // ODD number of 16-byte chunks: a quarter warp's LDS.128 of the same chunk index then falls into 8 distinct bank groups
    // (the rank pass reads the depths four at a time).  Sorted lists (every other K): an ODD row stride in words, so that the
    // 32 lanes' 4-byte accesses at the same list position fall into 32 distinct banks (a stride of 52 words at K = 50 made
    // every insertion step a 4-way conflict: half of all shared wavefronts of the C4 workload).
    if (fwd_sets(K, nwarps)) {
      KS = ((K + 3) / 4) * 4;
      if (((KS / 4) & 1) == 0) KS += 4;
    } else {
      KS = K | 1;
    }
    int w = 0;
    w_z = w; w += KS * 32 * 4;   // depth bits
    w_d = w; w += KS * 32 * 4;   // signed squared distance
    w_f = w; w += KS * 32 * 2;   // face id (ushort)
    w_o = w; w += KS * 32;       // slot of the k-th nearest entry (rank pass)
    w_q = w; w += kQueue * 32;
    w_cnt = w; w += 32;
    w_ord = w; w += 32;
    warp_bytes = w;
    total = off_union + max(nwarps * warp_bytes, vb + F * 4);  // the slabs alias the staging/bucketing scratch
  }
};

// fill `total` consecutive fragment slots starting at element `gbase` with the -1 padding
__device__ __forceinline__ void warp_fill_frag(const RasterParams& p, long long gbase, int total, int lane) {
  if (p.vec_ok && ((gbase | total) & 3) == 0) {
    int4* q = reinterpret_cast<int4*>(p.p2f + gbase);
    const int4 m1 = make_int4(-1, -1, -1, -1);
    for (int e = lane; e < (total >> 1); e += 32) q[e] = m1;
    float4* z = reinterpret_cast<float4*>(p.zbuf + gbase);
    float4* d = reinterpret_cast<float4*>(p.dists + gbase);
    const float4 f1 = make_float4(-1.f, -1.f, -1.f, -1.f);
    for (int e = lane; e < (total >> 2); e += 32) { z[e] = f1; d[e] = f1; }
  } else {
    for (int e = lane; e < total; e += 32) { p.p2f[gbase + e] = -1; p.zbuf[gbase + e] = -1.f; p.dists[gbase + e] = -1.f; }
  }
  if (p.bary) for (int e = lane; e < total * 3; e += 32) p.bary[gbase * 3 + e] = -1.f;
}

// fill a whole rectangle of pixels [x0,x1) x [y0,y1) of render n (all warps of the CTA)
template <int NWARPS>
__device__ __forceinline__ void cta_fill_rect(const RasterParams& p, int n, int x0, int x1, int y0, int y1, int warp, int lane) {
  const int npx = x1 - x0;
  for (int y = y0 + warp; y < y1; y += NWARPS) {
    const long long pix = ((long long)n * p.H + y) * p.W + x0;
    warp_fill_frag(p, pix * p.K, npx * p.K, lane);
    if (p.mask) for (int e = lane; e < npx; e += 32) p.mask[pix + e] = 0.0f;
  }
}

// ---- the exact per-(pixel, face) test of SURVEY.md §9.4 ---------------------------------------------
// Inputs: pixel centre, the face's vertices, den = edge(v2,v0,v1)+eps with its refined reciprocal, the
// refined reciprocals of the squared edge lengths and the "fast division is exact here" flags.
// Returns false if the pair produces no fragment; else the depth bits and the signed squared distance.
struct FaceB {
  float x0, y0, x1, y1, x2, y2, z0, z1, z2, den, yden, r01, r02, r12;
  int flags;  // face id | den_ok << 16 | l01_ok << 17 | l02_ok << 18 | l12_ok << 19
};

// generic version: every division individually guarded, degenerate edges handled (rare faces only).  Out of line, with the
// record passed BY VALUE and the result returned by value (depth bits, signed distance; depth bits 0xffffffff = no fragment),
// so that the hot caller keeps its record in registers: a by-reference signature forces the caller to spill the whole record
// to local memory on every pair, taken or not.
constexpr unsigned kNoFragment = 0xffffffffu;
struct FragZD { unsigned zbits; float sd; };
__device__ __forceinline__ bool eval_pair_generic_impl(const FaceB& r, float xf, float yf, int clip, float blur, float eps, unsigned& zbits, float& sd) {
  const float dx0 = fsub(xf, r.x0), dy0 = fsub(yf, r.y0), dx1 = fsub(xf, r.x1), dy1 = fsub(yf, r.y1);
  const float dx2 = fsub(xf, r.x2), dy2 = fsub(yf, r.y2);
  const float ex01 = fsub(r.x1, r.x0), ey01 = fsub(r.y1, r.y0), ex02 = fsub(r.x2, r.x0), ey02 = fsub(r.y2, r.y0);
  const float ex12 = fsub(r.x2, r.x1), ey12 = fsub(r.y2, r.y1);
  const bool den_ok = r.flags & 0x10000;
  const float w0 = fdiv_y(fsub(fmul(dx1, ey12), fmul(dy1, ex12)), r.den, r.yden, den_ok);  // edge(p,v1,v2)/den
  const float w1 = fdiv_y(fsub(fmul(dy2, ex02), fmul(dx2, ey02)), r.den, r.yden, den_ok);  // edge(p,v2,v0)/den
  const float w2 = fdiv_y(fsub(fmul(dx0, ey01), fmul(dy0, ex01)), r.den, r.yden, den_ok);  // edge(p,v0,v1)/den
  float c0w = w0, c1w = w1, c2w = w2;
  if (clip) {
    c0w = w0 > 0.0f ? w0 : 0.0f; c1w = w1 > 0.0f ? w1 : 0.0f; c2w = w2 > 0.0f ? w2 : 0.0f;
    float s = fadd(fadd(c0w, c1w), c2w);
    s = s > 1e-5f ? s : 1e-5f;
    c0w = fdiv_slow(c0w, s); c1w = fdiv_slow(c1w, s); c2w = fdiv_slow(c2w, s);
  }
  float pz = fadd(fadd(fmul(c0w, r.z0), fmul(c1w, r.z1)), fmul(c2w, r.z2));
  if (pz < 0.0f) return false;
  const float l01 = fadd(fmul(ex01, ex01), fmul(ey01, ey01));
  const float l02 = fadd(fmul(ex02, ex02), fmul(ey02, ey02));
  const float l12 = fadd(fmul(ex12, ex12), fmul(ey12, ey12));
  const float d01 = point_line_dist_h(dx0, dy0, dx1, dy1, r.x0, r.y0, xf, yf, ex01, ey01, l01, r.r01, r.flags & 0x20000, eps);
  const float d02 = point_line_dist_h(dx0, dy0, dx2, dy2, r.x0, r.y0, xf, yf, ex02, ey02, l02, r.r02, r.flags & 0x40000, eps);
  const float d12 = point_line_dist_h(dx1, dy1, dx2, dy2, r.x1, r.y1, xf, yf, ex12, ey12, l12, r.r12, r.flags & 0x80000, eps);
  const float dist = fminf(fminf(d01, d02), d12);
  const bool inside = w0 > 0.0f && w1 > 0.0f && w2 > 0.0f;
  if (!inside && dist >= blur) return false;
  pz = pz + 0.0f;  // canonicalise -0
  zbits = __float_as_uint(pz);
  sd = inside ? -dist : dist;
  return true;
}

__device__ __noinline__ FragZD eval_pair_generic(FaceB r, float xf, float yf, int clip, float blur, float eps) {
  FragZD o;
  o.zbits = kNoFragment; o.sd = 0.0f;
  unsigned z; float d;
  if (eval_pair_generic_impl(r, xf, yf, clip, blur, eps, z, d)) { o.zbits = z; o.sd = d; }
  return o;
}

// squared distance from p to q = a + clamp(t) * ba
__device__ __forceinline__ float seg_dist_t(float t, float ax, float ay, float bax, float bay, float px, float py) {
  const float tt = t < 0.0f ? 0.0f : (t > 1.0f ? 1.0f : t);
  const float qx = fadd(ax, fmul(tt, bax)), qy = fadd(ay, fmul(tt, bay));
  const float dx = fsub(px, qx), dy = fsub(py, qy);
  return fadd(fmul(dx, dx), fmul(dy, dy));
}

// Faces flagged kFaceFast (no degenerate edge, every denominator inside the fast-division window; practically all
// faces) take a straight-line path: p - v_i and the edge vectors, three barycentric quotients behind ONE operand
// guard, depth, three segment parameters behind one guard, distances.  Every product / difference is the same
// IEEE operation the reference performs (EdgeFunctionForward / PointLineDistanceForward), with common
// subexpressions shared; edge(p,v2,v0) uses -(v2-v0), whose negation commutes with rounding.
constexpr int kFaceFast = 0x100000;

// far_hi: depth bits beyond which a fragment cannot enter the lane's full set / list any more (0xffffffff while it has room):
// such pairs leave right after the depth, before the three point-segment distances (60 % of the arithmetic) — a round in which
// every lane holds such a pair, 19 % of the rounds at C2 (back layers met after the sets are full), skips them altogether.
__device__ __forceinline__ bool eval_pair(const FaceB& r, float xf, float yf, int clip, float blur, float eps, unsigned far_hi, unsigned& zbits, float& sd) {
  if (!(r.flags & kFaceFast)) {
    const FragZD o = eval_pair_generic(r, xf, yf, clip, blur, eps);
    zbits = o.zbits; sd = o.sd;
    return o.zbits != kNoFragment;
  }
  const float dx0 = fsub(xf, r.x0), dy0 = fsub(yf, r.y0), dx1 = fsub(xf, r.x1), dy1 = fsub(yf, r.y1);
  const float dx2 = fsub(xf, r.x2), dy2 = fsub(yf, r.y2);
  const float ex01 = fsub(r.x1, r.x0), ey01 = fsub(r.y1, r.y0), ex02 = fsub(r.x2, r.x0), ey02 = fsub(r.y2, r.y0);
  const float ex12 = fsub(r.x2, r.x1), ey12 = fsub(r.y2, r.y1);
  const float n0 = fsub(fmul(dx1, ey12), fmul(dy1, ex12));  // edge(p,v1,v2)
  const float n1 = fsub(fmul(dy2, ex02), fmul(dx2, ey02));  // edge(p,v2,v0)
  const float n2 = fsub(fmul(dx0, ey01), fmul(dy0, ex01));  // edge(p,v0,v1)
  float w0, w1, w2;
  if (div_safe3(n0, n1, n2)) {
    w0 = fdiv_fast(n0, r.den, r.yden); w1 = fdiv_fast(n1, r.den, r.yden); w2 = fdiv_fast(n2, r.den, r.yden);
  } else {
    w0 = fdiv_slow(n0, r.den); w1 = fdiv_slow(n1, r.den); w2 = fdiv_slow(n2, r.den);
  }
  float c0w = w0, c1w = w1, c2w = w2;
  if (clip) {
    c0w = w0 > 0.0f ? w0 : 0.0f; c1w = w1 > 0.0f ? w1 : 0.0f; c2w = w2 > 0.0f ? w2 : 0.0f;
    float s = fadd(fadd(c0w, c1w), c2w);
    s = s > 1e-5f ? s : 1e-5f;
    c0w = fdiv_slow(c0w, s); c1w = fdiv_slow(c1w, s); c2w = fdiv_slow(c2w, s);
  }
  float pz = fadd(fadd(fmul(c0w, r.z0), fmul(c1w, r.z1)), fmul(c2w, r.z2));
  if (pz < 0.0f) return false;
  if (__float_as_uint(pz + 0.0f) > far_hi) return false;  // (ties go on: the face id decides)
  const float l01 = fadd(fmul(ex01, ex01), fmul(ey01, ey01));
  const float l02 = fadd(fmul(ex02, ex02), fmul(ey02, ey02));
  const float l12 = fadd(fmul(ex12, ex12), fmul(ey12, ey12));
  const float m01 = fadd(fmul(ex01, dx0), fmul(ey01, dy0));
  const float m02 = fadd(fmul(ex02, dx0), fmul(ey02, dy0));
  const float m12 = fadd(fmul(ex12, dx1), fmul(ey12, dy1));
  float t01, t02, t12;
  if (div_safe3(m01, m02, m12)) {
    t01 = fdiv_fast(m01, l01, r.r01); t02 = fdiv_fast(m02, l02, r.r02); t12 = fdiv_fast(m12, l12, r.r12);
  } else {
    t01 = fdiv_slow(m01, l01); t02 = fdiv_slow(m02, l02); t12 = fdiv_slow(m12, l12);
  }
  const float d01 = seg_dist_t(t01, r.x0, r.y0, ex01, ey01, xf, yf);
  const float d02 = seg_dist_t(t02, r.x0, r.y0, ex02, ey02, xf, yf);
  const float d12 = seg_dist_t(t12, r.x1, r.y1, ex12, ey12, xf, yf);
  const float dist = fminf(fminf(d01, d02), d12);
  const bool inside = w0 > 0.0f && w1 > 0.0f && w2 > 0.0f;
  if (!inside && dist >= blur) return false;
  pz = pz + 0.0f;  // canonicalise -0
  zbits = __float_as_uint(pz);
  sd = inside ? -dist : dist;
  return true;
}

// Full per-face set-up from the three vertices (used once per region face, and by the overflow path).
struct FaceSetup {
  FaceB b;
  float bxmin, bxmax, bymin, bymax;
  float g[9];  // conservative edge equations (A,B,C) x 3
  float zcull;  // hard renders (blur == 0): a depth no fragment of the face can be nearer than; else 0 (never culls)
};

__device__ __forceinline__ void setup_face(FaceSetup& s, int f, float x0, float y0, float z0, float x1, float y1, float z1,
                                           float x2, float y2, float z2, float blur, float sq_blur, float eps) {
  FaceB& b = s.b;
  b.x0 = x0; b.y0 = y0; b.x1 = x1; b.y1 = y1; b.x2 = x2; b.y2 = y2; b.z0 = z0; b.z1 = z1; b.z2 = z2;
  s.bxmin = fsub(fminf(fminf(x0, x1), x2), sq_blur); s.bxmax = fadd(fmaxf(fmaxf(x0, x1), x2), sq_blur);
  s.bymin = fsub(fminf(fminf(y0, y1), y2), sq_blur); s.bymax = fadd(fmaxf(fmaxf(y0, y1), y2), sq_blur);
  const float area = edge_fn(x2, y2, x0, y0, x1, y1);
  b.den = fadd(area, eps);  // bary denominator
  // Depth cull of the hard renders (blur == 0: a fragment's pixel lies inside the face, so its barycentric weights are
  // positive and add up to area / (area + eps) — to 1 when they are clipped and renormalised — and its depth is at least that
  // sum times the nearest vertex's depth; the sum falls to 1/2 for faces barely above the degenerate-face threshold).  Margin
  // 1e-5 relative: two orders of magnitude over the rounding of the weights.  With a blur band the weights extrapolate beyond
  // the face and no such bound holds.
  const float zmin = fminf(fminf(z0, z1), z2);
  const float wsum = __fdividef(area, b.den);
  s.zcull = (blur == 0.0f && zmin > 0.0f && wsum > 0.0f) ? zmin * fminf(wsum, 1.0f) * 0.99999f : 0.0f;
  b.yden = rcp_refined(b.den);
  const float ex01 = fsub(x1, x0), ey01 = fsub(y1, y0), ex02 = fsub(x2, x0), ey02 = fsub(y2, y0);
  const float ex12 = fsub(x2, x1), ey12 = fsub(y2, y1);
  const float l01 = fadd(fmul(ex01, ex01), fmul(ey01, ey01));
  const float l02 = fadd(fmul(ex02, ex02), fmul(ey02, ey02));
  const float l12 = fadd(fmul(ex12, ex12), fmul(ey12, ey12));
  b.r01 = rcp_refined(l01); b.r02 = rcp_refined(l02); b.r12 = rcp_refined(l12);
  b.flags = f | (div_safe(b.den) ? 0x10000 : 0) | (div_safe(l01) ? 0x20000 : 0) | (div_safe(l02) ? 0x40000 : 0) |
            (div_safe(l12) ? 0x80000 : 0);
  if ((b.flags & 0xf0000) == 0xf0000 && l01 > eps && l02 > eps && l12 > eps) b.flags |= kFaceFast;
  // conservative edge equations (approximate arithmetic; see the header comment)
  const float mag = fmaxf(fmaxf(fmaxf(fabsf(x0), fabsf(x1)), fabsf(x2)), fmaxf(fmaxf(fabsf(y0), fabsf(y1)), fabsf(y2)));
  const float T = 1.001f * sq_blur + 1e-5f * (1.0f + mag);
  const float sgn = b.den > 0.0f ? 1.0f : -1.0f;
  const bool usable = fabsf(b.den) > 1e-6f && mag < 1e6f;
  // edge 0: through v1,v2 (n0 = dx1*ey12 - dy1*ex12); edge 1: through v2,v0 (n1 = dy2*ex02 - dx2*ey02);
  // edge 2: through v0,v1 (n2 = dx0*ey01 - dy0*ex01)
  const float ea[3] = {ey12, -ey02, ey01}, eb[3] = {-ex12, ex02, -ex01}, el[3] = {l12, l02, l01};
  const float qx[3] = {x1, x2, x0}, qy[3] = {y1, y2, y0};
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    float A = 0.0f, B = 0.0f, C = 0.0f;
    if (usable && el[i] > 1e-12f) {
      const float inv = sgn / (sqrtf(el[i]) * T);
      A = ea[i] * inv; B = eb[i] * inv;
      const float E = 4e-6f * (fabsf(A) * (1.0f + fabsf(qx[i])) + fabsf(B) * (1.0f + fabsf(qy[i])));
      const float sh = 1.0f / (1.0f + E);
      A *= sh; B *= sh;
      C = -(A * qx[i] + B * qy[i]);
    }
    s.g[i * 3] = A; s.g[i * 3 + 1] = B; s.g[i * 3 + 2] = C;
  }
}

// ---- per-pixel K-nearest SET (unsorted) -----------------------------------------------------------------------------
// Round 1 kept each pixel's fragments as a sorted list and inserted every new one by shifting.  The candidates of one
// surface layer arrive in an order unrelated to their depth order AT THE PIXEL (neighbouring faces' planes cross along the
// shared edge), so a mean insertion moved ~5 entries through a dependent LDS -> compare -> STS chain, and a warp waited for
// its longest shift every round: 36 % of the kernel's samples at 9 of 32 active lanes (profiles/raster_fwd_r04.md;
// scripts/sim_lanes.py reproduces the 9.7 lanes and shows that no face order or pixel grouping changes it).
// Now a lane only keeps the SET of its K nearest fragments: append while it has room (three stores, no load); once full, a
// nearer fragment replaces the farthest entry.  The farthest entry is looked up lazily (`stale`): after a replacement the
// old farthest key is still an upper bound, so later fragments beyond it are dropped at once and the set is rescanned only
// when a fragment falls below the bound.  The depth order is established once per tile by rank_entries(), in uniform
// control flow.  Built for K = KT = 20, the reference's faces_per_pixel (nmr.py:153): straight-line code that holds the 20
// depths in registers.  Every other K keeps round 1's sorted lists (list_insert below; the rank pass is quadratic in K).
struct KSet {
  unsigned* z;          // [KS] depth bits, 0xffffffff beyond cnt
  float* d;             // [KS]
  unsigned short* f;    // [KS]
  int cnt;
  int far_slot;                 // slot of the farthest entry (full sets; exact unless stale)
  unsigned long long far_key;   // its key (depth bits << 32 | face), or an upper bound of it when stale
  bool stale;
  float d1;                     // KT == 1: the set is this distance and far_key (all ones: empty); no shared memory
};

__device__ __forceinline__ unsigned umax3(unsigned a, unsigned b, unsigned c) { return max(max(a, b), c); }

// farthest entry of a FULL set (cnt == K = KT): depth first, then the larger face id (lexicographic (pz, f) as the
// reference's priority queue).  The maximum depth (3-input integer max), then the slots that hold it as a bit mask; more than
// one bit is a depth tie at the far end (rare), decided by the face ids.  Out of line, arguments and result by value: the
// caller keeps its set description in registers.  Returns (face, depth bits, slot).
template <int KT>
__device__ __noinline__ uint3 kset_rescan(const unsigned* sz, const unsigned short* sf) {
  static_assert(KT % 4 == 0 && KT <= 32, "whole 16-byte chunks, one mask word");
  unsigned zmax = 0u, eqm = 0u;
  uint4 v[KT / 4];
#pragma unroll
  for (int c = 0; c < KT / 4; ++c) v[c] = *reinterpret_cast<const uint4*>(sz + 4 * c);
#pragma unroll
  for (int c = 0; c < KT / 4; ++c) zmax = umax3(umax3(zmax, v[c].x, v[c].y), v[c].z, v[c].w);
#pragma unroll
  for (int c = 0; c < KT / 4; ++c) {  // (compare + predicated OR with an immediate: two instructions per entry)
    if (v[c].x == zmax) eqm |= 1u << (4 * c);
    if (v[c].y == zmax) eqm |= 1u << (4 * c + 1);
    if (v[c].z == zmax) eqm |= 1u << (4 * c + 2);
    if (v[c].w == zmax) eqm |= 1u << (4 * c + 3);
  }
  int slot = __ffs(eqm) - 1;
  unsigned fmax = sf[slot];
  eqm &= eqm - 1u;
  while (eqm) {
    const int j = __ffs(eqm) - 1;
    eqm &= eqm - 1u;
    if (sf[j] > fmax) { fmax = sf[j]; slot = j; }
  }
  return make_uint3(fmax, zmax, (unsigned)slot);
}

template <int KT>
__device__ __forceinline__ void kset_add(KSet& s, int K, unsigned zb, unsigned face, float sd) {
  const unsigned long long key = ((unsigned long long)zb << 32) | face;
  if (s.cnt < K) {
    const int slot = s.cnt++;
    s.z[slot] = zb; s.f[slot] = (unsigned short)face; s.d[slot] = sd;
    s.far_key = 0xffffffffffffffffull;  // full sets start with an unknown farthest entry
    s.stale = true;
  } else if (key < s.far_key) {
    if (s.stale) {
      const uint3 r = kset_rescan<KT>(s.z, s.f);
      s.far_key = ((unsigned long long)r.y << 32) | r.x;
      s.far_slot = (int)r.z;
      s.stale = false;
      if (key >= s.far_key) return;
    }
    s.z[s.far_slot] = zb; s.f[s.far_slot] = (unsigned short)face; s.d[s.far_slot] = sd;
    s.stale = true;  // far_key stays as an upper bound of the new farthest key
  }
}

// Depth order of every lane's set, all lanes at once: ord[r] = slot of the entry with r nearer entries.  Rank by counting on
// the 32-bit depths (slots beyond cnt hold 0xffffffff and never count as nearer); equal depths give equal ranks, which shows
// as a rank sum below cnt (cnt - 1) / 2 — only then the lane repeats the count on the exact (depth, face) keys.
// cmax = largest cnt of the warp.  No dependent shared-memory chain.
template <int KT>
__device__ __forceinline__ void rank_entries(const unsigned* z, const unsigned short* f, unsigned char* ord, int cnt, int cmax) {
  int rsum = 0;
  uint4 v[KT / 4];
#pragma unroll
  for (int c = 0; c < KT / 4; ++c) v[c] = *reinterpret_cast<const uint4*>(z + 4 * c);
#pragma unroll 2
  for (int i = 0; i < cmax; ++i) {
    const unsigned zi = z[i];  // 0xffffffff for i >= cnt (harmless: the rank is not used)
    int r = 0;
#pragma unroll
    for (int c = 0; c < KT / 4; ++c)
      r += (int)(v[c].x < zi) + (int)(v[c].y < zi) + (int)(v[c].z < zi) + (int)(v[c].w < zi);
    if (i < cnt) { ord[r] = (unsigned char)i; rsum += r; }
  }
  if (rsum != ((cnt * (cnt - 1)) >> 1)) {
    // depth ties in this pixel (0.6 % of the covered pixels of the reference templates): entries whose depth is unique keep
    // their rank; the tied ones are ranked again on the exact (depth, face) keys
#pragma unroll 1
    for (int i = 0; i < cnt; ++i) {
      const unsigned zi = z[i];
      int less = 0, same = 0;
#pragma unroll 1
      for (int j = 0; j < cnt; j += 4) {
        const uint4 v = *reinterpret_cast<const uint4*>(z + j);
        less += (int)(v.x < zi) + (int)(v.y < zi) + (int)(v.z < zi) + (int)(v.w < zi);
        same += (int)(v.x == zi) + (int)(v.y == zi) + (int)(v.z == zi) + (int)(v.w == zi);
      }
      if (same > 1) {
        const unsigned fi = f[i];
#pragma unroll 1
        for (int j = 0; j < cnt; ++j) less += (int)(z[j] == zi && f[j] < fi);
        ord[less] = (unsigned char)i;
      }
    }
  }
}

// ---- per-pixel K-nearest LIST (sorted), every K but 20 ------------------------------------------------------------------
// Round 1's structure, kept for the values of K the reference never uses: the same three arrays held in depth order by
// insertion (two entries per shared-memory round trip, the second speculative: the shift is a chain of dependent
// LDS -> compare -> STS, latency per entry is what counts).  Measured at 512^2: K = 32 3.3 ms / K = 50 4.8 ms against 5.1 / 5.7 ms
// for loop versions of the set operations above (their rank pass is quadratic in K).  The rank table is the identity.
__device__ __forceinline__ void list_insert(unsigned* lz, unsigned short* lf, float* ld, int K, int& cnt, unsigned long long& last,
                                            unsigned zb, unsigned face, float sd) {
  const unsigned long long key = ((unsigned long long)zb << 32) | face;
  int pos;
  if (cnt < K) {
    pos = cnt++;
  } else {
    if (key > last) return;  // not nearer than the current K-th
    pos = K - 1;
  }
  while (pos > 0) {
    const unsigned z1 = lz[pos - 1], f1 = lf[pos - 1];
    const float d1 = ld[pos - 1];
    const unsigned z2 = pos > 1 ? lz[pos - 2] : 0u, f2 = pos > 1 ? lf[pos - 2] : 0u;
    const float d2 = pos > 1 ? ld[pos - 2] : 0.0f;
    if ((((unsigned long long)z1 << 32) | f1) < key) break;
    lz[pos] = z1; lf[pos] = (unsigned short)f1; ld[pos] = d1;
    --pos;
    if (pos == 0 || (((unsigned long long)z2 << 32) | f2) < key) break;
    lz[pos] = z2; lf[pos] = (unsigned short)f2; ld[pos] = d2;
    --pos;
  }
  lz[pos] = zb; lf[pos] = (unsigned short)face; ld[pos] = sd;
  if (cnt == K) last = ((unsigned long long)lz[K - 1] << 32) | lf[K - 1];
}

// one accepted fragment into the lane's set (KT = 20) or list (any other K)
template <int KT>
__device__ __forceinline__ void frag_add(KSet& s, int K, unsigned zb, unsigned face, float sd) {
  if constexpr (KT == 1) {  // the nearest fragment, in registers
    const unsigned long long key = ((unsigned long long)zb << 32) | face;
    if (key < s.far_key) { s.far_key = key; s.d1 = sd; }
  } else if constexpr (KT > 0) kset_add<KT>(s, K, zb, face, sd);
  else list_insert(s.z, s.f, s.d, K, s.cnt, s.far_key, zb, face, sd);  // far_key holds the K-th key of a full list
}

// barrier of the rasterizer warps only (named barrier 1): the CTA's extra padding warp never joins it
template <int NT>
__device__ __forceinline__ void raster_sync() { asm volatile("bar.sync 1, %0;" ::"n"(NT) : "memory"); }

// stage_bulk_1d() of common.cuh for the rasterizer warps of a CTA that also holds the padding warp: same contract, with the
// named barrier instead of __syncthreads()
template <int NT>
__device__ __forceinline__ const float* raster_stage_verts(unsigned char* buf, const void* src, uint32_t bytes, uint64_t* bar, uint32_t phase) {
  const uintptr_t s = (uintptr_t)src;
  const uint32_t mis = (uint32_t)(s & 15u);
  unsigned char* dst = buf + mis;
  const uint32_t head = mis ? min(16u - mis, bytes) : 0u;
  const uint32_t body = (bytes - head) & ~15u;
  const uint32_t tail = bytes - head - body;
  if (threadIdx.x == 0 && body) {
    mbar_expect_tx(bar, body);
    tma_bulk_g2s(dst + head, (const unsigned char*)src + head, body, bar);
  }
  const uint32_t hw = head >> 2, tw = tail >> 2;  // at most 3 + 3 words, through the LSU
  if (threadIdx.x < hw) ((float*)dst)[threadIdx.x] = ((const float*)src)[threadIdx.x];
  if (threadIdx.x >= 32 && threadIdx.x < 32 + tw) {
    const uint32_t o = ((head + body) >> 2) + (threadIdx.x - 32);
    ((float*)dst)[o] = ((const float*)src)[o];
  }
  if (body) mbar_wait(bar, phase);
  raster_sync<NT>();
  return (const float*)dst;
}

// One (render, 32x32 region) unit, executed by the NWARPS rasterizer warps of a CTA (threads 0 .. NWARPS*32-1).
template <int NWARPS, typename IdxT, int KT, bool LEAN = false>
__device__ __forceinline__ void raster_unit(const RasterParams& p, unsigned char* smem, const int unit, const int slot = 0) {
  constexpr int NT = NWARPS * 32;
  const FwdSmem L(p.V, p.F, p.K, NWARPS, p.cap);
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem);
  int* rcount = reinterpret_cast<int*>(smem + 8);
  int* next_tile = reinterpret_cast<int*>(smem + 12);
  float* ndc_x = reinterpret_cast<float*>(smem + L.off_ndc);  // [kRegion] columns, then [kRegion] rows
  float* ndc_y = ndc_x + kRegion;
  float* red = reinterpret_cast<float*>(smem + L.off_red);
  int* hist = reinterpret_cast<int*>(smem + L.off_hist);
  int* tw = reinterpret_cast<int*>(smem + L.off_tw);
  float* tsum = reinterpret_cast<float*>(smem + L.off_ts);  // [32 tiles][4]
  unsigned short* rlist = reinterpret_cast<unsigned short*>(smem + L.off_rlist);
  float4* recA = reinterpret_cast<float4*>(smem + L.off_recA);  // [4][cap]: bbox | g0 g1.x | g1.yz g2.xy | g2.z
  float4* recB = reinterpret_cast<float4*>(smem + L.off_recB);  // [4][cap]: x0 y0 x1 y1 | x2 y2 z0 z1 | z2 den yden flags | r01 r02 r12
  const int cap = p.cap;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int regions = p.regions_x * p.regions_y;
  const int n = unit / regions;
  const int rg = unit - n * regions;
  const int px0 = (rg % p.regions_x) * kRegion, py0 = (rg / p.regions_x) * kRegion;
  const int px1 = min(px0 + kRegion, p.W), py1 = min(py0 + kRegion, p.H);
  const int K = KT > 0 ? KT : p.K;
  const long long fbase = (long long)n * p.faces_stride;
  const float* gverts = p.ndc + (size_t)n * p.V * 3;

  // (the mbarrier and the two counters were initialised by the kernel before the roles split)
  const float* sv = raster_stage_verts<NT>(smem + L.off_verts, gverts, (uint32_t)p.V * 12u, bar, 0);

  const float r_xhi = pix_to_ndc(p.W - 1 - px0, p.W), r_xlo = pix_to_ndc(p.W - 1 - (px1 - 1), p.W);
  const float r_yhi = pix_to_ndc(p.H - 1 - py0, p.H), r_ylo = pix_to_ndc(p.H - 1 - (py1 - 1), p.H);

  // ---- 1. mesh bounding box: regions that cannot be touched by any face skip the face scan ----------
  float zlo = INFINITY, zhi = -INFINITY;
  {
    float xmin = INFINITY, xmax = -INFINITY, ymin = INFINITY, ymax = -INFINITY;
    for (int v = tid; v < p.V; v += NT) {
      const float x = sv[v * 3], y = sv[v * 3 + 1], z = sv[v * 3 + 2];
      xmin = fminf(xmin, x); xmax = fmaxf(xmax, x); ymin = fminf(ymin, y); ymax = fmaxf(ymax, y);
      zlo = fminf(zlo, z); zhi = fmaxf(zhi, z);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      xmin = fminf(xmin, __shfl_xor_sync(0xffffffffu, xmin, o)); xmax = fmaxf(xmax, __shfl_xor_sync(0xffffffffu, xmax, o));
      ymin = fminf(ymin, __shfl_xor_sync(0xffffffffu, ymin, o)); ymax = fmaxf(ymax, __shfl_xor_sync(0xffffffffu, ymax, o));
      zlo = fminf(zlo, __shfl_xor_sync(0xffffffffu, zlo, o)); zhi = fmaxf(zhi, __shfl_xor_sync(0xffffffffu, zhi, o));
    }
    if (lane == 0) {
      red[warp * 8] = xmin; red[warp * 8 + 1] = xmax; red[warp * 8 + 2] = ymin; red[warp * 8 + 3] = ymax;
      red[warp * 8 + 4] = zlo; red[warp * 8 + 5] = zhi;
    }
    if (tid < kZBuckets) hist[tid] = 0;
    if (tid < 32) tw[tid] = 0;
    if (tid < 128) tsum[tid] = 0.0f;
    raster_sync<NT>();
#pragma unroll
    for (int w = 0; w < NWARPS; ++w) {
      xmin = fminf(xmin, red[w * 8]); xmax = fmaxf(xmax, red[w * 8 + 1]);
      ymin = fminf(ymin, red[w * 8 + 2]); ymax = fmaxf(ymax, red[w * 8 + 3]);
      zlo = fminf(zlo, red[w * 8 + 4]); zhi = fmaxf(zhi, red[w * 8 + 5]);
    }
    // same expansion and comparisons as the per-face test below, so this early-out is exact
    const bool outside = (r_xlo > fadd(xmax, p.sq_blur)) || (r_xhi < fsub(xmin, p.sq_blur)) ||
                         (r_ylo > fadd(ymax, p.sq_blur)) || (r_yhi < fsub(ymin, p.sq_blur));
    if (outside) {
      if constexpr (!LEAN) cta_fill_rect<NWARPS>(p, n, px0, px1, py0, py1, warp, lane);  // (lean: the mask was cleared by the host entry)
      return;
    }
  }

  // exact pixel-centre coordinates of the region's columns / rows (one division each, here only); entries past the image
  // edge repeat the last pixel.  Visible to everyone after the barriers of step 2.
  if (tid < 2 * kRegion) {
    const int i = tid & (kRegion - 1);
    if (tid < kRegion) ndc_x[i] = pix_to_ndc(p.W - 1 - min(px0 + i, p.W - 1), p.W);
    else ndc_y[i] = pix_to_ndc(p.H - 1 - min(py0 + i, p.H - 1), p.H);
  }

  // ---- 2. cull all faces against the region; bucket the survivors front to back ------------------------
  // Evaluating near faces first makes the per-pixel K-nearest lists fill with (almost) final entries, so
  // later candidates are appended or rejected by one compare instead of shifting entries.  The order
  // inside a bucket (and the order of atomics) is arbitrary: results do not depend on it, only the work.
  int* tmp = reinterpret_cast<int*>(smem + L.off_tmp);  // (bucket << 16 | face)
  const float zscale = (zhi > zlo) ? (float)kZBuckets / (3.0f * (zhi - zlo)) : 0.0f;
  for (int f0 = 0; f0 < p.F; f0 += NT) {
    const int f = f0 + tid;
    bool keep = false;
    int bucket = 0;
    if (f < p.F) {
      const IdxT* fp = reinterpret_cast<const IdxT*>(p.faces) + fbase + (long long)f * 3;
      const int i0 = (int)fp[0], i1 = (int)fp[1], i2 = (int)fp[2];
      const float x0 = sv[i0 * 3], y0 = sv[i0 * 3 + 1], z0 = sv[i0 * 3 + 2];
      const float x1 = sv[i1 * 3], y1 = sv[i1 * 3 + 1], z1 = sv[i1 * 3 + 2];
      const float x2 = sv[i2 * 3], y2 = sv[i2 * 3 + 1], z2 = sv[i2 * 3 + 2];
      const float zmax = fmaxf(fmaxf(z0, z1), z2);
      const float area = edge_fn(x0, y0, x1, y1, x2, y2);
      const bool skip = (zmax < 0.0f) || (p.cull && area < 0.0f) || (area <= p.k_eps && area >= -p.k_eps);
      const float bxmin = fsub(fminf(fminf(x0, x1), x2), p.sq_blur), bxmax = fadd(fmaxf(fmaxf(x0, x1), x2), p.sq_blur);
      const float bymin = fsub(fminf(fminf(y0, y1), y2), p.sq_blur), bymax = fadd(fmaxf(fmaxf(y0, y1), y2), p.sq_blur);
      keep = !skip && !(r_xlo > bxmax) && !(r_xhi < bxmin) && !(r_ylo > bymax) && !(r_yhi < bymin);
      bucket = min(kZBuckets - 1, max(0, (int)((z0 + z1 + z2 - 3.0f * zlo) * zscale)));
    }
    const unsigned m = __ballot_sync(0xffffffffu, keep);
    int base = 0;
    if (lane == 0 && m) base = atomicAdd(rcount, __popc(m));
    base = __shfl_sync(0xffffffffu, base, 0);
    if (keep) {
      tmp[base + __popc(m & ((1u << lane) - 1u))] = (bucket << 16) | f;
      atomicAdd(&hist[bucket], 1);
    }
  }
  raster_sync<NT>();
  const int nlist = *rcount;
  if (nlist == 0) {
    if constexpr (!LEAN) cta_fill_rect<NWARPS>(p, n, px0, px1, py0, py1, warp, lane);
    return;
  }
  if (warp == 0) {  // exclusive scan of the 64 bucket counts
    const int a = hist[lane * 2], b = hist[lane * 2 + 1];
    int incl = a + b;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int v = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += v;
    }
    hist[lane * 2] = incl - a - b;
    hist[lane * 2 + 1] = incl - b;
  }
  raster_sync<NT>();
  for (int e = tid; e < nlist; e += NT) {
    const int v = tmp[e];
    rlist[atomicAdd(&hist[v >> 16], 1)] = (unsigned short)(v & 0xffff);
  }
  raster_sync<NT>();

  // ---- 3. per-region face records; faces per tile ----------------------------------------------------------
  // While a face's record is written, its blur-expanded bounding box is tested against the region's 4 x 8 tiles (same
  // comparisons as the per-pixel test) and the hits are counted per tile: lane t of every warp accumulates tile t.
  // The counts order the tiles heaviest first (longest-processing-time-first over the CTA's warps: the warps of a
  // region finish closer together) and let tiles that no face touches go straight to the padding.
  const int nrec = min(nlist, cap);
  int tw_acc = 0;
  for (int j0 = 0; j0 < nrec; j0 += NT) {
    const int j = j0 + tid;
    unsigned tm = 0u;
    if (j < nrec) {
      const int f = rlist[j];
      const IdxT* fp = reinterpret_cast<const IdxT*>(p.faces) + fbase + (long long)f * 3;
      const int i0 = (int)fp[0], i1 = (int)fp[1], i2 = (int)fp[2];
      FaceSetup s;
      setup_face(s, f, sv[i0 * 3], sv[i0 * 3 + 1], sv[i0 * 3 + 2], sv[i1 * 3], sv[i1 * 3 + 1], sv[i1 * 3 + 2], sv[i2 * 3],
                 sv[i2 * 3 + 1], sv[i2 * 3 + 2], p.blur, p.sq_blur, p.k_eps);
      recA[j] = make_float4(s.bxmin, s.bxmax, s.bymin, s.bymax);
      recA[cap + j] = make_float4(s.g[0], s.g[1], s.g[2], s.g[3]);
      recA[2 * cap + j] = make_float4(s.g[4], s.g[5], s.g[6], s.g[7]);
      recA[3 * cap + j] = make_float4(s.g[8], s.zcull, 0.f, 0.f);
      recB[j] = make_float4(s.b.x0, s.b.y0, s.b.x1, s.b.y1);
      recB[cap + j] = make_float4(s.b.x2, s.b.y2, s.b.z0, s.b.z1);
      recB[2 * cap + j] = make_float4(s.b.z2, s.b.den, s.b.yden, __int_as_float(s.b.flags));
      recB[3 * cap + j] = make_float4(s.b.r01, s.b.r02, s.b.r12, 0.f);
      unsigned colm = 0u;
#pragma unroll
      for (int c = 0; c < kRegion / kTileW; ++c)
        if (!(ndc_x[c * kTileW + kTileW - 1] > s.bxmax) && !(ndc_x[c * kTileW] < s.bxmin)) colm |= 1u << c;
#pragma unroll
      for (int r = 0; r < kRegion / kTileH; ++r)
        if (!(ndc_y[r * kTileH + kTileH - 1] > s.bymax) && !(ndc_y[r * kTileH] < s.bymin)) tm |= colm << (r * (kRegion / kTileW));
    }
#pragma unroll
    for (int t = 0; t < 32; ++t) {
      const unsigned b = __ballot_sync(0xffffffffu, (tm >> t) & 1u);
      if (lane == t) tw_acc += __popc(b);
    }
  }
  if (tw_acc) atomicAdd(&tw[lane], tw_acc);
  raster_sync<NT>();  // the staging scratch (verts, tmp) is dead from here on: the warp slabs alias it

  // ---- 4. warps pull 8x4 tiles; no CTA-wide synchronisation from here on ----------------------------
  unsigned char* wslab = smem + L.off_union + warp * L.warp_bytes;
  const int KS = L.KS;
  KSet ks;
  ks.z = reinterpret_cast<unsigned*>(wslab + L.w_z) + lane * KS;
  ks.d = reinterpret_cast<float*>(wslab + L.w_d) + lane * KS;
  ks.f = reinterpret_cast<unsigned short*>(wslab + L.w_f) + lane * KS;
  unsigned char* ord = wslab + L.w_o + lane * KS;  // [KS] slot of the lane's k-th nearest entry
  unsigned char* queue = wslab + L.w_q;     // [slot][lane]
  unsigned char* cnts = wslab + L.w_cnt;    // [lane]
  const unsigned* wz = reinterpret_cast<const unsigned*>(wslab + L.w_z);  // the same arrays, pixel-major, for the output pass
  const float* wd = reinterpret_cast<const float*>(wslab + L.w_d);
  const unsigned short* wf = reinterpret_cast<const unsigned short*>(wslab + L.w_f);
  const unsigned char* wo = wslab + L.w_o;
  const int tiles_x = (px1 - px0 + kTileW - 1) / kTileW, tiles_y = (py1 - py0 + kTileH - 1) / kTileH;
  const int ntiles = tiles_x * tiles_y;
  const float inv_sigma_neg = p.sigma > 0.0f ? 1.0f / p.sigma : 0.0f;
  // tile order of this region, heaviest first (every warp derives the same table for itself: no further CTA barrier).
  // Lane t stands for tile t of the 4 x 8 grid; tiles outside the image sort last and are never reached (t < ntiles).
  unsigned char* ordv = wslab + L.w_ord;
  const int my_w = tw[lane];
  {
    const bool in_img = (lane & 3) < tiles_x && (lane >> 2) < tiles_y;
    const int key = in_img ? ((my_w + 1) << 5) | (31 - lane) : 0;  // ties: raster order
    int rank = 0;
#pragma unroll
    for (int o = 0; o < 32; ++o) rank += __shfl_sync(0xffffffffu, key, o) > key;
    ordv[rank] = (unsigned char)lane;  // keys of image tiles are distinct; the others collide past ntiles, harmlessly
    __syncwarp();
  }
  const bool has_overflow = nlist > nrec;

  while (true) {
    int t = 0;
    if (lane == 0) t = atomicAdd(next_tile, 1);
    t = __shfl_sync(0xffffffffu, t, 0);
    if (t >= ntiles) break;
    const int tbit = ordv[t];
    const int trow = tbit >> 2;
    const int tx0 = px0 + (tbit & 3) * kTileW, ty0 = py0 + trow * kTileH;
    const int xi = tx0 + (lane & 7), yi = ty0 + (lane >> 3);
    const bool valid = xi < p.W && yi < p.H;
    const int lx0 = tx0 - px0, ly0 = ty0 - py0;  // tile origin inside the region
    const float xf = ndc_x[lx0 + (lane & 7)], yf = ndc_y[ly0 + (lane >> 3)];
    const float t_xhi = ndc_x[lx0], t_xlo = ndc_x[lx0 + kTileW - 1];  // table entries past the image edge repeat the last pixel
    const float t_yhi = ndc_y[ly0], t_ylo = ndc_y[ly0 + kTileH - 1];

    int qn = 0;
    ks.cnt = 0; ks.far_slot = 0; ks.far_key = 0xffffffffffffffffull; ks.stale = true;  // (room: nothing is too far)
    ks.d1 = 0.0f;
    // K = 1: depth beyond which no face can matter to the TILE any more (every pixel has a nearer fragment); refreshed after
    // every drain of the queues, which is why these renders drain at depth kQueue1
    unsigned tile_far = 0xffffffffu;
    if constexpr (KT > 1) {
      // empty set: every depth slot "infinitely far" (rank pass), every rank slot a valid index (output pass; filled by
      // rank_entries()).  The sorted lists (KT = 0) need neither: their order is the slot order and cnt bounds every access.
      for (int j = 0; j < KS; j += 4) {
        *reinterpret_cast<uint4*>(ks.z + j) = make_uint4(0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu);
        *reinterpret_cast<unsigned*>(ord + j) = 0u;
      }
    }

    // (a)-(c): fill the per-lane queues face by face; drain them when one is full or the faces run out (one drain site)
    // a tile that no record's bounding box touches (and no overflow face could) skips the scan: straight to the padding
    int c0 = (!has_overflow && __shfl_sync(0xffffffffu, my_w, tbit) == 0) ? nrec : 0, cbase = 0;
    unsigned m = 0u;
    bool more = true;
    do {
      bool full = false;
      while (!full) {
        if (m == 0u) {
          if (c0 >= nrec) { more = false; break; }
          // (a) scan: lane = face; tile-level bbox + edge-equation culling
          const int j = c0 + lane;
          bool hit = false;
          if (j < nrec) {
            const float4 bb = recA[j];
            hit = !(t_xlo > bb.y) && !(t_xhi < bb.x) && !(t_ylo > bb.w) && !(t_yhi < bb.z);
            if (hit) {
              const float4 ga = recA[cap + j], gb = recA[2 * cap + j];
              const float4 g3 = recA[3 * cap + j];
              const float gc = g3.x;
              const float m0 = fmaf(ga.x, ga.x > 0.f ? t_xhi : t_xlo, fmaf(ga.y, ga.y > 0.f ? t_yhi : t_ylo, ga.z));
              const float m1 = fmaf(ga.w, ga.w > 0.f ? t_xhi : t_xlo, fmaf(gb.x, gb.x > 0.f ? t_yhi : t_ylo, gb.y));
              const float m2 = fmaf(gb.z, gb.z > 0.f ? t_xhi : t_xlo, fmaf(gb.w, gb.w > 0.f ? t_yhi : t_ylo, gc));
              hit = !(m0 < -1.0f) && !(m1 < -1.0f) && !(m2 < -1.0f);
              if constexpr (KT == 1) hit = hit && !(__float_as_uint(g3.y) > tile_far);  // behind what every pixel already has
            }
          }
          m = __ballot_sync(0xffffffffu, hit);
          cbase = c0;
          c0 += 32;
          continue;
        }
        // (b) filter: lane = pixel, face uniform; survivors are queued per lane
        const int jj = cbase + __ffs(m) - 1;
        m &= m - 1;
        const float4 bb = recA[jj], ga = recA[cap + jj], gb = recA[2 * cap + jj];
        const float4 g3 = recA[3 * cap + jj];
        const float gc = g3.x;
        bool cand = valid && !(xf > bb.y) && !(xf < bb.x) && !(yf > bb.w) && !(yf < bb.z);
        if constexpr (KT == 1) cand = cand && !(__float_as_uint(g3.y) > (unsigned)(ks.far_key >> 32));  // behind this pixel's fragment
        const float g0 = fmaf(ga.x, xf, fmaf(ga.y, yf, ga.z));
        const float g1 = fmaf(ga.w, xf, fmaf(gb.x, yf, gb.y));
        const float g2 = fmaf(gb.z, xf, fmaf(gb.w, yf, gc));
        cand = cand && !(g0 < -1.0f) && !(g1 < -1.0f) && !(g2 < -1.0f);
        if (cand) queue[(qn++) * 32 + lane] = (unsigned char)jj;
        full = __any_sync(0xffffffffu, qn == (KT == 1 ? kQueue1 : kQueue));
      }
      // (c) evaluate: each lane pops its own queue.  (Measured without gain: prefetching the next entry's record a round early;
      // letting a lane with an empty queue evaluate a candidate of its mirror lane (lane ^ 31) and hand the result back by
      // shuffles — 20 % fewer rounds in scripts/sim_lanes.py, but 2.41 -> 2.50 ms at C2: the shuffles, the selects and the
      // second add per round cost more than the idle lanes did.)
      const int qmax = __reduce_max_sync(0xffffffffu, qn);
      for (int i = 0; i < qmax; ++i) {
        if (i < qn) {
          const int j = queue[i * 32 + lane];
          const float4 b0 = recB[j], b1 = recB[cap + j], b2 = recB[2 * cap + j], b3 = recB[3 * cap + j];
          FaceB r;
          r.x0 = b0.x; r.y0 = b0.y; r.x1 = b0.z; r.y1 = b0.w; r.x2 = b1.x; r.y2 = b1.y; r.z0 = b1.z; r.z1 = b1.w;
          r.z2 = b2.x; r.den = b2.y; r.yden = b2.z; r.flags = __float_as_int(b2.w);
          r.r01 = b3.x; r.r02 = b3.y; r.r12 = b3.z;
          unsigned zb;
          float sd;
          if (eval_pair(r, xf, yf, p.clip, p.blur, p.k_eps, (unsigned)(ks.far_key >> 32), zb, sd)) frag_add<KT>(ks, K, zb, (unsigned)(r.flags & 0xffff), sd);
        }
      }
      qn = 0;
      if constexpr (KT == 1) tile_far = __reduce_max_sync(0xffffffffu, valid ? (unsigned)(ks.far_key >> 32) : 0u);
    } while (more);

    // overflow: region faces without a record (face-uniform evaluation, set-up broadcast by shuffles)
    for (int c0 = nrec; c0 < nlist; c0 += 32) {
      const int j = c0 + lane;
      bool hit = false;
      FaceSetup s;
      if (j < nlist) {
        const int f = rlist[j];
        const IdxT* fp = reinterpret_cast<const IdxT*>(p.faces) + fbase + (long long)f * 3;
        const int i0 = (int)fp[0], i1 = (int)fp[1], i2 = (int)fp[2];
        setup_face(s, f, gverts[i0 * 3], gverts[i0 * 3 + 1], gverts[i0 * 3 + 2], gverts[i1 * 3], gverts[i1 * 3 + 1],
                   gverts[i1 * 3 + 2], gverts[i2 * 3], gverts[i2 * 3 + 1], gverts[i2 * 3 + 2], p.blur, p.sq_blur, p.k_eps);
        hit = !(t_xlo > s.bxmax) && !(t_xhi < s.bxmin) && !(t_ylo > s.bymax) && !(t_yhi < s.bymin);
        if constexpr (KT == 1) hit = hit && !(__float_as_uint(s.zcull) > tile_far);
      }
      unsigned m = __ballot_sync(0xffffffffu, hit);
      while (m) {
        const int src = __ffs(m) - 1;
        m &= m - 1;
        FaceB r;
#define ACFM_BC(field) r.field = __shfl_sync(0xffffffffu, s.b.field, src)
        ACFM_BC(x0); ACFM_BC(y0); ACFM_BC(x1); ACFM_BC(y1); ACFM_BC(x2); ACFM_BC(y2); ACFM_BC(z0); ACFM_BC(z1); ACFM_BC(z2);
        ACFM_BC(den); ACFM_BC(yden); ACFM_BC(r01); ACFM_BC(r02); ACFM_BC(r12); ACFM_BC(flags);
#undef ACFM_BC
        const float axmin = __shfl_sync(0xffffffffu, s.bxmin, src), axmax = __shfl_sync(0xffffffffu, s.bxmax, src);
        const float aymin = __shfl_sync(0xffffffffu, s.bymin, src), aymax = __shfl_sync(0xffffffffu, s.bymax, src);
        if (!valid || xf > axmax || xf < axmin || yf > aymax || yf < aymin) continue;
        unsigned zb;
        float sd;
        if (eval_pair(r, xf, yf, p.clip, p.blur, p.k_eps, (unsigned)(ks.far_key >> 32), zb, sd)) frag_add<KT>(ks, K, zb, (unsigned)(r.flags & 0xffff), sd);
      }
    }

    // ---- (d) depth order, blend, write ------------------------------------------------------------------
    const int npx = min(kTileW, p.W - tx0);
    const int nrows = min(kTileH, p.H - ty0);
    const int cnt = KT == 1 ? (int)(ks.far_key != 0xffffffffffffffffull) : ks.cnt;
    const int cmax = __reduce_max_sync(0xffffffffu, cnt);
    if (cmax == 0) {
      if constexpr (!LEAN) {
        for (int row = 0; row < nrows; ++row) {
          const long long pix = ((long long)n * p.H + ty0 + row) * p.W + tx0;
          warp_fill_frag(p, pix * K, npx * K, lane);
        }
        if (p.mask && valid) p.mask[((long long)n * p.H + yi) * p.W + xi] = 0.0f;
      }
      continue;
    }
    if constexpr (KT > 1) {
      if (cmax > 1) rank_entries<KT>(ks.z, ks.f, ord, cnt, cmax);  // one entry: ord[0] = 0 from the reset above
    }
    if (p.vis && cnt > 0) {
      // visible vertices = vertices of the faces that are nearest at some pixel (the fi_maps -> unique -> scatter_ block
      // of bds_loss / optical_flow_loss, loss_utils.py:213-223,432-441): one lane per distinct face of the tile stores
      const unsigned fv = KT == 1 ? (unsigned)ks.far_key & 0xffffu : (unsigned)ks.f[KT > 1 ? ord[0] : 0];
      const unsigned peers = __match_any_sync(__activemask(), fv);
      if (__ffs(peers) - 1 == lane) {
        const IdxT* fp = reinterpret_cast<const IdxT*>(p.faces) + fbase + (long long)fv * 3;
        float* vo = p.vis + (size_t)n * p.V;
        vo[(int)fp[0]] = 1.0f; vo[(int)fp[1]] = 1.0f; vo[(int)fp[2]] = 1.0f;
      }
    }
    float l0 = 0.0f, l1 = 0.0f, l2 = 0.0f, l3 = 0.0f;
    if (p.mask && valid) {
      float alpha = 1.0f;
      for (int i = 0; i < cnt; ++i) {  // in depth order, like the reference's product (the bits of the mask do not depend on
                                       // the order in which the faces were met)
        // 1 - sigmoid(-d / sigma) = 1 / (1 + exp(-d / sigma)); fast exp / divide: ~2e-7 relative, the mask is held to 1e-5
        const float di = KT == 1 ? ks.d1 : ks.d[KT > 1 ? ord[i] : i];
        alpha *= __fdividef(1.0f, 1.0f + __expf(-di * inv_sigma_neg));
      }
      p.mask[((long long)n * p.H + yi) * p.W + xi] = 1.0f - alpha;
      if (p.loss_part) {
        // fused silhouette losses (l1 / iou / edt, loss_utils.py:18-32,72-77,245-253): this pixel's share of
        // sum(|m-t| - |t|), sum m t, sum (m - m t), sum edt m — the parts that vanish where the mask is 0, so that regions and
        // tiles without fragments contribute nothing and the sums over the bare target are added once per render afterwards
        const float m = 1.0f - alpha;
        const size_t ti = ((size_t)(n % p.NB) * p.H + yi) * p.W + xi;
        const float t = p.loss_target[ti];
        l0 = fabsf(m - t) - fabsf(t); l1 = m * t; l2 = m - m * t;
        l3 = p.loss_edt ? p.loss_edt[ti] * m : 0.0f;
      }
    }
    if (p.loss_part) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        l0 += __shfl_xor_sync(0xffffffffu, l0, o); l1 += __shfl_xor_sync(0xffffffffu, l1, o);
        l2 += __shfl_xor_sync(0xffffffffu, l2, o); l3 += __shfl_xor_sync(0xffffffffu, l3, o);
      }
      if (lane == 0) { tsum[tbit * 4] = l0; tsum[tbit * 4 + 1] = l1; tsum[tbit * 4 + 2] = l2; tsum[tbit * 4 + 3] = l3; }
    }
    const long long row_stride = (long long)p.W * K;                       // elements between image rows
    const long long tbase = (((long long)n * p.H + ty0) * p.W + tx0) * K;  // first element of the tile
    const long long nF = (long long)n * p.F;
    if constexpr (KT == 1) {
      // K = 1: every lane writes its own pixel from its registers (a row of the tile is 64 / 32 / 32 contiguous bytes)
      if (valid) {
        const long long g = ((long long)n * p.H + yi) * p.W + xi;
        const int fv = (int)((unsigned)ks.far_key & 0xffffu);
        p.p2f[g] = cnt ? nF + fv : -1ll;
        p.zbuf[g] = cnt ? __uint_as_float((unsigned)(ks.far_key >> 32)) : -1.f;
        p.dists[g] = cnt ? ks.d1 : -1.f;
        if (p.bary) {
          // barycentrics of the fragment, recomputed from the face id with the operator sequence of the evaluation
          float b0 = -1.f, b1 = -1.f, b2 = -1.f;
          if (cnt) {
            const IdxT* fp = reinterpret_cast<const IdxT*>(p.faces) + fbase + (long long)fv * 3;
            const int i0 = (int)fp[0], i1 = (int)fp[1], i2 = (int)fp[2];
            const float x0 = gverts[i0 * 3], y0 = gverts[i0 * 3 + 1], x1 = gverts[i1 * 3], y1 = gverts[i1 * 3 + 1];
            const float x2 = gverts[i2 * 3], y2 = gverts[i2 * 3 + 1];
            const float den = fadd(edge_fn(x2, y2, x0, y0, x1, y1), p.k_eps);
            b0 = fdiv(edge_fn(xf, yf, x1, y1, x2, y2), den);
            b1 = fdiv(edge_fn(xf, yf, x2, y2, x0, y0), den);
            b2 = fdiv(edge_fn(xf, yf, x0, y0, x1, y1), den);
            if (p.clip) {
              b0 = b0 > 0.0f ? b0 : 0.0f; b1 = b1 > 0.0f ? b1 : 0.0f; b2 = b2 > 0.0f ? b2 : 0.0f;
              float sw = fadd(fadd(b0, b1), b2);
              sw = sw > 1e-5f ? sw : 1e-5f;
              b0 = fdiv(b0, sw); b1 = fdiv(b1, sw); b2 = fdiv(b2, sw);
            }
          }
          float* bo = p.bary + g * 3;
          bo[0] = b0; bo[1] = b1; bo[2] = b2;
        }
      }
      continue;
    }
    if constexpr (LEAN) {
      // compact fragments for the backward: K face ids (2 B) and K distances per pixel of the region, in depth order, 0xffff
      // beyond cnt.  Pixels without a fragment have mask 0 and are never read back: nothing is written for them.
      // Every lane writes its own pixel with 16-byte stores (K = 20: 40 B of ids, 80 B of distances; rows are 8-byte / 16-byte
      // aligned: K * 2 and K * 4 bytes per pixel).
      if (valid && cnt > 0) {
        const size_t pixr = (size_t)slot * (kRegion * kRegion) + (size_t)(ly0 + (lane >> 3)) * kRegion + (lx0 + (lane & 7));
        unsigned short* of = p.lean_f + pixr * K;
        float* od = p.lean_d + pixr * K;
#pragma unroll
        for (int k = 0; k < KT; k += 4) {
          const unsigned so = *reinterpret_cast<const unsigned*>(ord + k);
          const int s0 = so & 0xffu, s1 = (so >> 8) & 0xffu, s2 = (so >> 16) & 0xffu, s3 = so >> 24;
          const unsigned f0 = k < cnt ? ks.f[s0] : 0xffffu, f1 = k + 1 < cnt ? ks.f[s1] : 0xffffu;
          const unsigned f2 = k + 2 < cnt ? ks.f[s2] : 0xffffu, f3 = k + 3 < cnt ? ks.f[s3] : 0xffffu;
          *reinterpret_cast<uint2*>(of + k) = make_uint2(f0 | (f1 << 16), f2 | (f3 << 16));
          *reinterpret_cast<float4*>(od + k) = make_float4(k < cnt ? ks.d[s0] : 0.f, k + 1 < cnt ? ks.d[s1] : 0.f,
                                                          k + 2 < cnt ? ks.d[s2] : 0.f, k + 3 < cnt ? ks.d[s3] : 0.f);
        }
      }
      __syncwarp();  // the sets are reused by the next tile
      continue;
    }
    cnts[lane] = (unsigned char)cnt;
    __syncwarp();
    if (p.vec_ok && (K & 3) == 0) {
      // p2f: two entries (16 B) per store; zbuf / dists: four entries (16 B) per store
      const int P2 = K >> 1, P4 = K >> 2;
      // 16-bit magics: exact for divisors <= 50 (brute-forced over e < 32 P); P2 <= 32 and P4 <= 16 here
      const unsigned d2 = (65536u + (unsigned)P2 - 1u) / (unsigned)P2, d4 = (65536u + (unsigned)P4 - 1u) / (unsigned)P4;
      for (int e = lane; e < 32 * P2; e += 32) {
        const int pxl = (int)(((unsigned)e * d2) >> 16);
        const int k = (e - pxl * P2) * 2;
        const int col = pxl & 7, row = pxl >> 3;
        if (col >= npx || row >= nrows) continue;
        const int c = cnts[pxl];
        const int o = pxl * KS;
        // two rank slots (the sorted lists' order is the slot order)
        const unsigned so = KT > 1 ? *reinterpret_cast<const unsigned short*>(wo + o + k) : (unsigned)(k | ((k + 1) << 8));
        const long long a = k < c ? nF + wf[o + (so & 0xffu)] : -1ll;
        const long long b = k + 1 < c ? nF + wf[o + (so >> 8)] : -1ll;
        longlong2 v; v.x = a; v.y = b;
        *reinterpret_cast<longlong2*>(p.p2f + tbase + row * row_stride + col * K + k) = v;
      }
      for (int e = lane; e < 32 * P4; e += 32) {
        const int pxl = (int)(((unsigned)e * d4) >> 16);
        const int k = (e - pxl * P4) * 4;
        const int col = pxl & 7, row = pxl >> 3;
        if (col >= npx || row >= nrows) continue;
        const int c = cnts[pxl];
        const int o = pxl * KS;
        int s0 = o + k, s1 = s0 + 1, s2 = s0 + 2, s3 = s0 + 3;
        if constexpr (KT > 1) {
          const unsigned so = *reinterpret_cast<const unsigned*>(wo + o + k);  // four rank slots
          s0 = o + (so & 0xffu); s1 = o + ((so >> 8) & 0xffu); s2 = o + ((so >> 16) & 0xffu); s3 = o + (so >> 24);
        }
        float4 z, d;
        z.x = k < c ? __uint_as_float(wz[s0]) : -1.f; d.x = k < c ? wd[s0] : -1.f;
        z.y = k + 1 < c ? __uint_as_float(wz[s1]) : -1.f; d.y = k + 1 < c ? wd[s1] : -1.f;
        z.z = k + 2 < c ? __uint_as_float(wz[s2]) : -1.f; d.z = k + 2 < c ? wd[s2] : -1.f;
        z.w = k + 3 < c ? __uint_as_float(wz[s3]) : -1.f; d.w = k + 3 < c ? wd[s3] : -1.f;
        const long long g = tbase + row * row_stride + col * K + k;
        *reinterpret_cast<float4*>(p.zbuf + g) = z;
        *reinterpret_cast<float4*>(p.dists + g) = d;
      }
    } else {
      // e / K == umulhi(e, ceil(2^32 / K)) for every e < 2^32 / K (brute-forced for K <= 150, e < 32 K); a 16-bit magic is
      // wrong for K = 51, 56, 60..63
      const unsigned kdiv = (unsigned)((0x100000000ull + (unsigned long long)K - 1ull) / (unsigned long long)K);
      for (int e = lane; e < 32 * K; e += 32) {
        const int pxl = K == 1 ? e : (int)__umulhi((unsigned)e, kdiv);
        const int k = e - pxl * K;
        const int col = pxl & 7, row = pxl >> 3;
        if (col >= npx || row >= nrows) continue;
        const int c = cnts[pxl];
        const int o = pxl * KS + (KT > 1 ? (int)wo[pxl * KS + k] : k);
        const long long g = tbase + row * row_stride + col * K + k;
        p.p2f[g] = k < c ? nF + wf[o] : -1ll;
        p.zbuf[g] = k < c ? __uint_as_float(wz[o]) : -1.f;
        p.dists[g] = k < c ? wd[o] : -1.f;
      }
    }
    if (p.bary) {
      // barycentrics of the surviving fragments are recomputed from the face id with the same operator
      // sequence as the evaluation above (bit-identical); only the hard (K = 1) texture path asks for them.
      for (int e = lane; e < 32 * K; e += 32) {
        const int pxl = e / K;
        const int k = e - pxl * K;
        const int col = pxl & 7, row = pxl >> 3;
        if (col >= npx || row >= nrows) continue;
        float b0 = -1.f, b1 = -1.f, b2 = -1.f;
        if (k < cnts[pxl]) {
          const int fv = (int)wf[pxl * KS + (KT > 1 ? (int)wo[pxl * KS + k] : k)];
          const IdxT* fp = reinterpret_cast<const IdxT*>(p.faces) + fbase + (long long)fv * 3;
          const int i0 = (int)fp[0], i1 = (int)fp[1], i2 = (int)fp[2];
          const float x0 = gverts[i0 * 3], y0 = gverts[i0 * 3 + 1], x1 = gverts[i1 * 3], y1 = gverts[i1 * 3 + 1];
          const float x2 = gverts[i2 * 3], y2 = gverts[i2 * 3 + 1];
          const float pxf = ndc_x[lx0 + col], pyf = ndc_y[ly0 + row];
          const float den = fadd(edge_fn(x2, y2, x0, y0, x1, y1), p.k_eps);
          b0 = fdiv(edge_fn(pxf, pyf, x1, y1, x2, y2), den);
          b1 = fdiv(edge_fn(pxf, pyf, x2, y2, x0, y0), den);
          b2 = fdiv(edge_fn(pxf, pyf, x0, y0, x1, y1), den);
          if (p.clip) {
            b0 = b0 > 0.0f ? b0 : 0.0f; b1 = b1 > 0.0f ? b1 : 0.0f; b2 = b2 > 0.0f ? b2 : 0.0f;
            float s = fadd(fadd(b0, b1), b2);
            s = s > 1e-5f ? s : 1e-5f;
            b0 = fdiv(b0, s); b1 = fdiv(b1, s); b2 = fdiv(b2, s);
          }
        }
        float* bo = p.bary + (tbase + row * row_stride + col * K + k) * 3;
        bo[0] = b0; bo[1] = b1; bo[2] = b2;
      }
    }
    __syncwarp();  // the lists are reused by the next tile
  }
  if (p.loss_part) {
    // the region's four partial sums: the 32 tile slots added in tile order (not in the order the warps happened to pull them)
    raster_sync<NT>();
    if (tid < 4) {
      float a = 0.0f;
#pragma unroll
      for (int t = 0; t < 32; ++t) a += tsum[t * 4 + tid];
      p.loss_part[(size_t)unit * 4 + tid] = a;
    }
  }
}

// ---- the two kernels of the split path ----------------------------------------------------------------------------------
// 87 % of all fragment bytes of the reference's workloads are the -1 padding of regions the mesh cannot touch.  It is written
// by raster_fill_kernel BESIDE the rasterizer: one-warp CTAs, one per SM, whose stores are issued by the TMA unit
// (cp.async.bulk shared -> global from a constant pattern in shared memory: one instruction per row piece, so the LSU / MIO
// queues stay with the co-resident rasterizer CTAs; a CTA of ordinary stores saturates them — measured 4.05-4.23 ms at C2, no
// better than one kernel doing everything).  The pattern holds max(32 K, 640) fragment slots (7.5 KB up to K = 20): such a
// CTA fits beside two resident rasterizer CTAs; 5 KB / 2.5 KB pieces reach the HBM write rate with one CTA per SM, 2 KB /
// 1 KB pieces a third of it.
//
// Placement matters more than anything else here: the padding CTAs must sit one per SM.  Round 1 launched the two kernels on
// two streams (fork / join events, high priority for the padding) and so depended on a launch race: whenever the
// rasterizer's grid reached the SMs first, the padding CTAs were packed onto the few SMs that freed up first and wrote at
// those SMs' rate only — 4x to 30x slower, time inversely proportional to the number of padding CTAs (measured at K = 24 and
// 48 with round 1's build, at every K when the rasterizer used all registers of an SM; profiles/raster_fwd_r05.md).  Now both
// kernels are launched on the caller's stream and overlap by PROGRAMMATIC DEPENDENT LAUNCH: the padding kernel starts on the
// idle GPU after raster_prep_kernel (its <= one-wave grid is spread one CTA per SM) and every CTA signals
// griddepcontrol.launch_dependents at once; the rasterizer kernel, launched with programmatic stream serialization, is
// dispatched when all of them have — i.e. exactly when the padding CTAs are in place.  The rasterizer does not read what the
// padding kernel writes, so it never waits for it, except for ONE thread of its last CTA, which executes griddepcontrol.wait
// before leaving: the rasterizer grid — and with it everything that follows on the stream — completes only after the
// padding is complete and visible.  No second stream, no events, nothing to pool or to leak.
constexpr int kFillThreads = 32;

__device__ __forceinline__ void bulk_s2g(void* dst, const void* src_smem, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(smem_u32(src_smem)), "r"(bytes)
               : "memory");
}

__device__ __forceinline__ void bulk_fill_row(void* dst, const void* pat_smem, int slots, int slot_bytes, int chunk) {
  for (int o = 0; o < slots; o += chunk)
    bulk_s2g(reinterpret_cast<unsigned char*>(dst) + (size_t)o * slot_bytes, pat_smem, (uint32_t)(min(chunk, slots - o) * slot_bytes));
}

__global__ void __maxnreg__(40) raster_fill_kernel(const RasterParams p, const int* ws) {  // (few registers: it shares SM sub-partitions with four 112-register rasterizer warps)
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");  // in place: the rasterizer's grid may be dispatched
  extern __shared__ __align__(128) unsigned char pat[];
  const int regions = p.regions_x * p.regions_y;
  const int countF = ws[1];
  const int* listF = ws + 8 + kWeightClasses * p.N * regions;
  const int lane = threadIdx.x;
  const int K = p.K, chunk = fill_pattern_slots(K);
  long long* pat8 = reinterpret_cast<long long*>(pat);                         // [chunk] int64 -1
  float* pat4 = reinterpret_cast<float*>(pat + (size_t)chunk * 8);             // [chunk] float -1
  if (p.bulk_ok) {
    for (int e = lane; e < chunk; e += kFillThreads) { pat8[e] = -1ll; pat4[e] = -1.0f; }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> visible to the bulk-copy engine
    __syncwarp();
  }
  // How many CTAs pad: five eighths of the grid (one CTA on most SMs) write ~4 TB/s, enough to finish under the rasterizer at
  // the reference's workloads and gentler on it than one per SM (C2, 74 / 92 / 110 / 148 CTAs: 2.95 / 2.79 / 2.80 / 2.85 ms);
  // the rest joins only when the padding would otherwise outlast the rasterizer (sparse views: few live regions, many bytes
  // to pad).  Estimate: a live region costs the rasterizer ~0.3 us of the GPU (K = 20, reference templates).
  const long long pad_bytes = (long long)ws[2] * (kRegion * kRegion) * (16ll * K + 4);
  const bool all_ctas = pad_bytes > (long long)ws[0] * (long long)p.pad_balance;
  const int nctas = all_ctas ? gridDim.x : max(1, ((int)gridDim.x * 5) >> 3);
  if ((int)blockIdx.x >= nctas) return;
  for (int w = blockIdx.x; w < countF; w += nctas) {
    const int unit = listF[2 * w], run = listF[2 * w + 1];
    const int n = unit / regions, rg = unit - n * regions;
    const int px0 = (rg % p.regions_x) * kRegion, py0 = (rg / p.regions_x) * kRegion;
    const int px1 = min(px0 + run * kRegion, p.W), py1 = min(py0 + kRegion, p.H);
    const int npx = px1 - px0, row_el = npx * K;
    if (p.bulk_ok && (row_el & 3) == 0) {
      if (py0 + lane < py1) {  // lane = pixel row of the run
        const long long g = (((long long)n * p.H + py0 + lane) * p.W + px0) * K;  // first fragment slot of the row
        bulk_fill_row(p.p2f + g, pat8, row_el, 8, chunk);
        bulk_fill_row(p.zbuf + g, pat4, row_el, 4, chunk);
        bulk_fill_row(p.dists + g, pat4, row_el, 4, chunk);
        if (p.bary) bulk_fill_row(p.bary + g * 3, pat4, row_el * 3, 4, chunk);
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      }
      if (p.mask)
        for (int y = py0; y < py1; ++y)
          for (int x = lane; x < npx; x += kFillThreads) p.mask[((long long)n * p.H + y) * p.W + px0 + x] = 0.0f;
    } else {
      cta_fill_rect<1>(p, n, px0, px1, py0, py1, 0, lane);
    }
  }
  asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");  // the pattern must outlive the reads; writes complete before exit
}

// One CTA (NWARPS warps) per (render, region) unit.
template <int NWARPS, typename IdxT, int KT, bool LEAN = false>
__global__ void __launch_bounds__(NWARPS * 32) __maxnreg__(NWARPS <= 8 ? 112 : 112) raster_fwd_kernel(const RasterParams p) {
  // (112 registers: two 8-warp CTAs leave every SM sub-partition room for one padding warp beside its four rasterizer warps;
  // the lean instantiation has no padding kernel beside it, but 128 registers gain it 0.6 %)
  extern __shared__ __align__(128) unsigned char smem[];
  const int regions = p.regions_x * p.regions_y;
  int unit = blockIdx.x;
  bool live = true;
  if (p.work) {
    // split path: only regions the mesh can touch are rendered (the others are padded by raster_fill_kernel), in four weight
    // classes, heaviest first (raster_prep_kernel): the grid's last CTAs are light ones, which shortens the kernel's tail
    const int U = p.N * regions;
    const int c0 = p.work[4], c1 = p.work[5], c2 = p.work[6], c3 = p.work[7];
    const int cls = unit < c0 ? 0 : (unit < c0 + c1 ? 1 : (unit < c0 + c1 + c2 ? 2 : 3));
    live = unit < c0 + c1 + c2 + c3;
    if (live) unit = p.work[8 + cls * U + (unit - (cls > 0 ? c0 : 0) - (cls > 1 ? c1 : 0) - (cls > 2 ? c2 : 0))];
  }
  if (live) {
    if (threadIdx.x == 0) {
      mbar_init(reinterpret_cast<uint64_t*>(smem), 1);
      mbar_fence_init();
      *reinterpret_cast<int*>(smem + 8) = 0;   // rcount
      *reinterpret_cast<int*>(smem + 12) = 0;  // next_tile
    }
    __syncthreads();
    raster_unit<NWARPS, IdxT, KT, LEAN>(p, smem, unit, (int)blockIdx.x);  // (slot = position in the work lists)
  }
  // the grid completes only after the padding kernel has (see above): one thread of the last CTA waits for it
  if (p.work && blockIdx.x == gridDim.x - 1 && threadIdx.x == 0) asm volatile("griddepcontrol.wait;" ::: "memory");
}

// choose the CTA size and the record capacity: 8-warp CTAs, two per SM (227 KB shared per SM, 1 KB reserved per CTA), with
// the largest record table that keeps two resident; 4-warp CTAs when the per-warp sets are too large for that (K > ~40).
// 10- and 12-warp CTAs were measured slower at every shape of the reference (profiles/README.md) and are not built.
int fwd_pick_config(int V, int F, int K, int* smem_bytes, int* cap_out) {
  int best_nw = 0, best_score = -1, best_cap = 0, best_smem = 0;
  for (int nw : {8, 4}) {
    for (int cap : {256, 224, 192, 160, 128, 96, 64}) {
      const FwdSmem l(V, F, K, nw, cap);
      if (l.total > 227 * 1024) continue;
      const int ctas = min(32, (228 * 1024) / (l.total + 1024));
      const int warps = min(64, ctas * nw);
      // a record table below ~192 entries overflows on ordinary views (a 32x32 region of the reference
      // templates sees 130-175 faces, up to ~500), so capacity comes first, then resident warps (capped at 16), then
      // fewer, larger tables
      const int score = min(cap, 192) * 10000 + min(warps, 16) * 100 + cap / 32;
      if (score > best_score) { best_score = score; best_nw = nw; best_cap = cap; best_smem = l.total; }
    }
  }
  *smem_bytes = best_smem;
  *cap_out = best_cap;
  return best_nw;
}

template <int NWARPS, typename IdxT, int KT, bool LEAN = false>
int launch_fwd(const RasterParams& p, int smem, int ctas, bool overlap, cudaStream_t st) {
  auto kern = raster_fwd_kernel<NWARPS, IdxT, KT, LEAN>;
  static std::atomic<int> smem_set[kAcfmMaxDevices];
  ACFM_CUDA_OK(acfm_ensure_smem(kern, smem, smem_set));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)ctas); cfg.blockDim = dim3(NWARPS * 32); cfg.dynamicSmemBytes = (size_t)smem; cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at; cfg.numAttrs = overlap ? 1 : 0;  // overlap: dispatch as soon as the padding kernel's CTAs are in place
  ACFM_CUDA_OK(cudaLaunchKernelEx(&cfg, kern, p));
  return ACFM_OK;
}


// ---- split path: region classification + concurrent fill -----------------------------------------------
// At the reference's workloads ~75% of the (render, region) units lie outside the blur-expanded bounding box of the mesh
// and 87% of all fragment bytes are -1 padding.  Inside the rasterizer kernel those units hold a 109 KB / 256-thread CTA
// slot while their stores drain (18% of the kernel's warp samples, profiles/raster_fwd_r01.md), slots the issue-bound
// units then lack.  The split path classifies the units first (raster_prep_kernel, one small CTA per render), gives the
// rasterizer kernel only the units the mesh can touch, and pads the others from raster_fill_kernel on a second stream:
// one-warp CTAs with a few KB of shared memory, which run in the resources the rasterizer leaves free, so the HBM write
// stream overlaps the arithmetic (C2: 3.97 -> 3.17 ms; the rasterizer alone 3.08 ms, the fill alone 1.33 ms).
constexpr int kPrepThreads = 128;

template <typename IdxT>
__global__ void __launch_bounds__(kPrepThreads) raster_prep_kernel(const RasterParams p, int* ws, int weigh) {
  extern __shared__ int rcnt[];  // faces per region (weigh != 0)
  __shared__ float red[kPrepThreads / 32][4];
  __shared__ float box[4];
  const int n = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float* v = p.ndc + (size_t)n * p.V * 3;
  float xmin = INFINITY, xmax = -INFINITY, ymin = INFINITY, ymax = -INFINITY;
  for (int i = tid; i < p.V; i += kPrepThreads) {
    const float x = v[i * 3], y = v[i * 3 + 1];
    xmin = fminf(xmin, x); xmax = fmaxf(xmax, x); ymin = fminf(ymin, y); ymax = fmaxf(ymax, y);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    xmin = fminf(xmin, __shfl_xor_sync(0xffffffffu, xmin, o)); xmax = fmaxf(xmax, __shfl_xor_sync(0xffffffffu, xmax, o));
    ymin = fminf(ymin, __shfl_xor_sync(0xffffffffu, ymin, o)); ymax = fmaxf(ymax, __shfl_xor_sync(0xffffffffu, ymax, o));
  }
  if (lane == 0) { red[warp][0] = xmin; red[warp][1] = xmax; red[warp][2] = ymin; red[warp][3] = ymax; }
  __syncthreads();
  if (tid == 0) {
#pragma unroll
    for (int w = 1; w < kPrepThreads / 32; ++w) {
      xmin = fminf(xmin, red[w][0]); xmax = fmaxf(xmax, red[w][1]); ymin = fminf(ymin, red[w][2]); ymax = fmaxf(ymax, red[w][3]);
    }
    // the same expansion as the per-face bounding boxes, so the classification is exact
    box[0] = fsub(xmin, p.sq_blur); box[1] = fadd(xmax, p.sq_blur); box[2] = fsub(ymin, p.sq_blur); box[3] = fadd(ymax, p.sq_blur);
  }
  __syncthreads();
  const float bxmin = box[0], bxmax = box[1], bymin = box[2], bymax = box[3];
  const int regions = p.regions_x * p.regions_y;
  const int U = p.N * regions;
  int* listR = ws + 8;
  int* listF = ws + 8 + kWeightClasses * U;
  // Faces per region (blur-expanded bounding boxes; approximate pixel arithmetic — the counts only ORDER the rasterizer's
  // work, heaviest regions first, they never decide what is rendered)
  if (weigh) {
    for (int r = tid; r < regions; r += kPrepThreads) rcnt[r] = 0;
    __syncthreads();
    const IdxT* fn = reinterpret_cast<const IdxT*>(p.faces) + (long long)n * p.faces_stride;
    const float hw = 0.5f * (float)p.W, hh = 0.5f * (float)p.H;
    for (int f = tid; f < p.F; f += kPrepThreads) {
      const int i0 = (int)fn[f * 3], i1 = (int)fn[f * 3 + 1], i2 = (int)fn[f * 3 + 2];
      const float x0 = v[i0 * 3], y0 = v[i0 * 3 + 1], x1 = v[i1 * 3], y1 = v[i1 * 3 + 1], x2 = v[i2 * 3], y2 = v[i2 * 3 + 1];
      const float fx0 = fminf(fminf(x0, x1), x2) - p.sq_blur, fx1 = fmaxf(fmaxf(x0, x1), x2) + p.sq_blur;
      const float fy0 = fminf(fminf(y0, y1), y2) - p.sq_blur, fy1 = fmaxf(fmaxf(y0, y1), y2) + p.sq_blur;
      if (!(fx0 <= 1.0f && fx1 >= -1.0f && fy0 <= 1.0f && fy1 >= -1.0f)) continue;  // off screen (or NaN)
      // pixel xi samples NDC x = 1 - (2 xi + 1) / W  =>  xi = (1 - x) W / 2 - 1/2
      const int ca = max(0, (int)floorf((1.0f - fminf(fx1, 1.0f)) * hw - 0.5f)) / kRegion;
      const int cb = min(p.W - 1, (int)ceilf((1.0f - fmaxf(fx0, -1.0f)) * hw - 0.5f)) / kRegion;
      const int ra = max(0, (int)floorf((1.0f - fminf(fy1, 1.0f)) * hh - 0.5f)) / kRegion;
      const int rb = min(p.H - 1, (int)ceilf((1.0f - fmaxf(fy0, -1.0f)) * hh - 0.5f)) / kRegion;
      if ((rb - ra + 1) * (cb - ca + 1) > 64) continue;  // a face this large weighs on every region alike; keeps the pass bounded
      for (int r = ra; r <= rb; ++r)
        for (int c = ca; c <= cb; ++c) atomicAdd(&rcnt[r * p.regions_x + c], 1);
    }
    __syncthreads();
  }
  for (int r0 = 0; r0 < regions; r0 += kPrepThreads) {
    const int rg = r0 + tid;
    bool live = false, inr = rg < regions;
    if (inr) {
      const int px0 = (rg % p.regions_x) * kRegion, py0 = (rg / p.regions_x) * kRegion;
      const int px1 = min(px0 + kRegion, p.W), py1 = min(py0 + kRegion, p.H);
      const float r_xhi = pix_to_ndc(p.W - 1 - px0, p.W), r_xlo = pix_to_ndc(p.W - 1 - (px1 - 1), p.W);
      const float r_yhi = pix_to_ndc(p.H - 1 - py0, p.H), r_ylo = pix_to_ndc(p.H - 1 - (py1 - 1), p.H);
      live = p.F > 0 && !((r_xlo > bxmax) || (r_xhi < bxmin) || (r_ylo > bymax) || (r_yhi < bymin));
    }
    // units for the rasterizer: one per live region; units for the fill kernel: horizontal RUNS of empty regions inside one
    // row of regions (and one warp's 32 consecutive regions), so that a bulk store can span the whole run
    const unsigned mr = __ballot_sync(0xffffffffu, inr && live), me = __ballot_sync(0xffffffffu, inr && !live);
    const int col = inr ? rg % p.regions_x : 0;
    const bool start = inr && !live && (lane == 0 || col == 0 || !((me >> (lane - 1)) & 1u));
    int run = 0;
    if (start) {
      const unsigned stop = ~(me >> lane);  // lowest set bit = first non-empty region at or after this lane
      run = min(stop ? __ffs(stop) - 1 : 32, p.regions_x - col);
    }
    const unsigned mf = __ballot_sync(0xffffffffu, start);
    int bf = 0;
    if (lane == 0) {
      if (mr) atomicAdd(&ws[0], __popc(mr));
      if (mf) bf = atomicAdd(&ws[1], __popc(mf));
      if (me) atomicAdd(&ws[2], __popc(me));  // empty regions (the runs' total length)
    }
    bf = __shfl_sync(0xffffffffu, bf, 0);
    const unsigned lt = (1u << lane) - 1u;
    // live regions by weight class (a 32 x 32 region of the reference templates sees 130-175 faces on average, up to ~500)
    const int wgt = (weigh && inr && live) ? rcnt[rg] : 0;
    const int cls = wgt >= 280 ? 0 : (wgt >= 200 ? 1 : (wgt >= 130 ? 2 : 3));
#pragma unroll
    for (int k = 0; k < kWeightClasses; ++k) {
      const unsigned mk = __ballot_sync(0xffffffffu, inr && live && cls == k);
      int bk = 0;
      if (lane == 0 && mk) bk = atomicAdd(&ws[4 + k], __popc(mk));
      bk = __shfl_sync(0xffffffffu, bk, 0);
      if (inr && live && cls == k) listR[k * U + bk + __popc(mk & lt)] = n * regions + rg;
    }
    if (start) {
      const int e = bf + __popc(mf & lt);
      listF[2 * e] = n * regions + rg;
      listF[2 * e + 1] = run;
    }
  }
}

// Tuning hooks (environment variables) exist only in builds with -DACFM_TUNING (scripts/build_variant.sh); the shipped
// library reads no environment.  ACFM_FWD_WARPS / ACFM_FWD_CAP override the launch configuration, ACFM_FWD_ONLY=raster|fill
// launches one of the two kernels only (timing: the outputs are incomplete), ACFM_FILL_CTAS / ACFM_FILL_PER_SM /
// ACFM_FILL_LSU / ACFM_PAD_BALANCE vary the padding kernel, ACFM_NO_PDL serialises the two kernels.
struct FwdTuning {
  int only = 0;  // 'r' / 'f'
  int fill_per_sm = 1, fill_abs = 0, fill_lsu = 0, no_pdl = 0;
  int pad_balance = 1200000;
};
const FwdTuning& fwd_tuning() {
  static const FwdTuning t = [] {
    FwdTuning v;
#ifdef ACFM_TUNING
    if (const char* e = getenv("ACFM_FWD_ONLY")) v.only = e[0];
    if (const char* e = getenv("ACFM_FILL_PER_SM")) v.fill_per_sm = std::max(1, atoi(e));
    if (const char* e = getenv("ACFM_FILL_CTAS")) v.fill_abs = atoi(e);
    if (const char* e = getenv("ACFM_PAD_BALANCE")) v.pad_balance = atoi(e);
    v.fill_lsu = getenv("ACFM_FILL_LSU") != nullptr;
    v.no_pdl = getenv("ACFM_NO_PDL") != nullptr;
#endif
    return v;
  }();
  return t;
}
int device_sm_count(int* sms) {  // cached per device
  static std::atomic<int> cache[kAcfmMaxDevices];
  int dev = 0;
  ACFM_CUDA_OK(cudaGetDevice(&dev));
  std::atomic<int>& c = cache[dev & (kAcfmMaxDevices - 1)];
  int v = c.load(std::memory_order_relaxed);
  if (v == 0) {
    ACFM_CUDA_OK(cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev));
    c.store(v, std::memory_order_relaxed);
  }
  *sms = v;
  return ACFM_OK;
}
int fwd_config(int V, int F, int K, int* smem, int* cap) {
  int nw = fwd_pick_config(V, F, K, smem, cap);
#ifdef ACFM_TUNING
  const char* ew = getenv("ACFM_FWD_WARPS");
  const char* ec = getenv("ACFM_FWD_CAP");
  if (ew || ec) {
    const int w = ew ? atoi(ew) : nw, c = ec ? atoi(ec) : *cap;
    if ((w == 4 || w == 8) && c >= 0 && c <= 256) {
      const FwdSmem l(V, F, K, w, c);
      if (l.total <= 227 * 1024) { nw = w; *cap = c; *smem = l.total; }
    }
  }
#endif
  return nw;
}

// ---- fused silhouette losses: the reductions around the rasterizer ---------------------------------------------------------
// base[b] = { sum |t_b|, sum t_b } over the bare target b (what the four sums are where the mask is 0); one CTA per target
__global__ void __launch_bounds__(256) loss_target_base_kernel(const float* __restrict__ target, int HW, float* __restrict__ base) {
  const float* t = target + (size_t)blockIdx.x * HW;
  float a = 0.0f, b = 0.0f;
  for (int i = threadIdx.x; i < HW; i += 256) { const float v = t[i]; a += fabsf(v); b += v; }
  __shared__ float red[8][2];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { a += __shfl_xor_sync(0xffffffffu, a, o); b += __shfl_xor_sync(0xffffffffu, b, o); }
  if ((threadIdx.x & 31) == 0) { red[threadIdx.x >> 5][0] = a; red[threadIdx.x >> 5][1] = b; }
  __syncthreads();
  if (threadIdx.x < 2) {
    float s = 0.0f;
#pragma unroll
    for (int w = 0; w < 8; ++w) s += red[w][threadIdx.x];
    base[blockIdx.x * 2 + threadIdx.x] = s;
  }
}

// sums[n] = { sum|m-t|, sum m t, sum (m+t-mt), sum edt m } = the render's region partials (fixed order) + the target's base
__global__ void __launch_bounds__(32) loss_reduce_kernel(const float* __restrict__ part, const float* __restrict__ base, int regions,
                                                         int NB, float* __restrict__ sums) {
  const int n = blockIdx.x, lane = threadIdx.x;
  float a0 = 0.0f, a1 = 0.0f, a2 = 0.0f, a3 = 0.0f;
  for (int r = lane; r < regions; r += 32) {
    const float4 v = *reinterpret_cast<const float4*>(part + ((size_t)n * regions + r) * 4);
    a0 += v.x; a1 += v.y; a2 += v.z; a3 += v.w;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    a0 += __shfl_xor_sync(0xffffffffu, a0, o); a1 += __shfl_xor_sync(0xffffffffu, a1, o);
    a2 += __shfl_xor_sync(0xffffffffu, a2, o); a3 += __shfl_xor_sync(0xffffffffu, a3, o);
  }
  if (lane == 0) {
    const float* b = base + (size_t)(n % NB) * 2;
    *reinterpret_cast<float4*>(sums + (size_t)n * 4) = make_float4(a0 + b[0], a1, a2 + b[1], a3);
  }
}

}  // namespace

namespace {
int raster_fwd_impl(const float* ndc, const void* faces, int faces_i64, int64_t faces_batch_stride, int N, int V, int F, int H,
                    int W, int K, float blur_radius, int clip_bary, int cull_backfaces, float sigma, int64_t* pix_to_face,
                    float* zbuf, float* dists, float* bary, float* mask, float* visible_verts, const float* loss_target,
                    const float* loss_edt, int NB, float* loss_sums, void* loss_workspace, int64_t loss_workspace_bytes,
                    void* workspace, int64_t workspace_bytes, void* stream, void* lean_workspace = nullptr, int64_t lean_bytes = 0);
}

extern "C" int acfm_raster_fwd_launch_info(int N, int V, int F, int H, int W, int K, int* smem_bytes, int* num_ctas,
                                           int* threads) {
  ACFM_REQUIRE(N >= 0 && V > 0 && F > 0 && H > 0 && W > 0 && K > 0, ACFM_ERR_BAD_ARG, "acfm_raster_fwd_launch_info: bad sizes");
  int smem = 0, cap = 0;
  const int nw = fwd_config(V, F, K, &smem, &cap);
  ACFM_REQUIRE(nw > 0, ACFM_ERR_UNSUPPORTED, "rasterizer needs more than 232448 B of shared memory for V=%d F=%d K=%d", V, F, K);
  if (smem_bytes) *smem_bytes = smem;
  if (num_ctas) *num_ctas = N * ((W + kRegion - 1) / kRegion) * ((H + kRegion - 1) / kRegion);
  if (threads) *threads = nw * 32;
  return ACFM_OK;
}

extern "C" int acfm_raster_fwd(const float* ndc, const void* faces, int faces_i64, int64_t faces_batch_stride, int N,
                               int V, int F, int H, int W, int K, float blur_radius, int clip_bary, int cull_backfaces,
                               float sigma, int64_t* pix_to_face, float* zbuf, float* dists, float* bary, float* mask,
                               float* visible_verts, void* workspace, int64_t workspace_bytes, void* stream) {
  return raster_fwd_impl(ndc, faces, faces_i64, faces_batch_stride, N, V, F, H, W, K, blur_radius, clip_bary, cull_backfaces, sigma,
                         pix_to_face, zbuf, dists, bary, mask, visible_verts, nullptr, nullptr, 0, nullptr, nullptr, 0, workspace,
                         workspace_bytes, stream);
}

extern "C" int64_t acfm_raster_loss_workspace_bytes(int N, int NB, int H, int W) {
  if (N <= 0 || NB <= 0 || H <= 0 || W <= 0) return 0;
  return 16 * (int64_t)N * ((W + kRegion - 1) / kRegion) * ((H + kRegion - 1) / kRegion) + 8 * (int64_t)NB;
}

extern "C" int acfm_raster_fwd_train(const float* ndc, const void* faces, int faces_i64, int64_t faces_batch_stride, int N, int V,
                                     int F, int H, int W, int K, float blur_radius, float sigma, int64_t* pix_to_face, float* zbuf,
                                     float* dists, float* mask, float* visible_verts, const float* target,
                                     const float* edt, int NB, float* loss_sums, void* loss_workspace, int64_t loss_workspace_bytes,
                                     void* workspace, int64_t workspace_bytes, void* stream) {
  ACFM_REQUIRE(mask, ACFM_ERR_BAD_ARG, "acfm_raster_fwd_train: null mask pointer");
  ACFM_REQUIRE(!loss_sums || (target && loss_workspace), ACFM_ERR_BAD_ARG, "acfm_raster_fwd_train: loss_sums needs target and loss_workspace");
  ACFM_REQUIRE(!loss_sums || (NB > 0 && N % NB == 0), ACFM_ERR_BAD_ARG, "acfm_raster_fwd_train: N=%d is not a multiple of NB=%d", N, NB);
  return raster_fwd_impl(ndc, faces, faces_i64, faces_batch_stride, N, V, F, H, W, K, blur_radius, 0, 0, sigma, pix_to_face, zbuf,
                         dists, nullptr, mask, visible_verts, loss_sums ? target : nullptr, loss_sums ? edt : nullptr, NB, loss_sums,
                         loss_workspace, loss_workspace_bytes, workspace, workspace_bytes, stream);
}

namespace {
int raster_fwd_impl(const float* ndc, const void* faces, int faces_i64, int64_t faces_batch_stride, int N, int V, int F, int H,
                    int W, int K, float blur_radius, int clip_bary, int cull_backfaces, float sigma, int64_t* pix_to_face,
                    float* zbuf, float* dists, float* bary, float* mask, float* visible_verts, const float* loss_target,
                    const float* loss_edt, int NB, float* loss_sums, void* loss_workspace, int64_t loss_workspace_bytes,
                    void* workspace, int64_t workspace_bytes, void* stream, void* lean_workspace, int64_t lean_bytes) {
  const bool lean = lean_workspace != nullptr;
  ACFM_REQUIRE(N >= 0 && V >= 0 && F >= 0 && H > 0 && W > 0, ACFM_ERR_BAD_ARG, "acfm_raster_fwd: bad sizes N=%d V=%d F=%d H=%d W=%d", N, V, F, H, W);
  ACFM_REQUIRE(K >= 1, ACFM_ERR_BAD_ARG, "acfm_raster_fwd: faces_per_pixel K=%d must be >= 1", K);
  ACFM_REQUIRE(K <= 64, ACFM_ERR_UNSUPPORTED, "acfm_raster_fwd: faces_per_pixel K=%d > 64 is not supported", K);
  ACFM_REQUIRE(blur_radius >= 0.0f, ACFM_ERR_BAD_ARG, "acfm_raster_fwd: blur_radius must be >= 0");
  ACFM_REQUIRE(!mask || sigma > 0.0f, ACFM_ERR_BAD_ARG, "acfm_raster_fwd: mask output requires sigma > 0");
  ACFM_REQUIRE(faces_batch_stride == 0 || faces_batch_stride == (int64_t)F * 3, ACFM_ERR_BAD_ARG, "acfm_raster_fwd: faces_batch_stride must be 0 or F*3");
  if (N == 0) return ACFM_OK;
  ACFM_REQUIRE(lean || (pix_to_face && zbuf && dists), ACFM_ERR_BAD_ARG, "acfm_raster_fwd: null output pointer");
  ACFM_REQUIRE((ndc || V == 0) && (faces || F == 0), ACFM_ERR_BAD_ARG, "acfm_raster_fwd: null input pointer");
  ACFM_REQUIRE(F <= 65535 && V <= 65535, ACFM_ERR_UNSUPPORTED, "acfm_raster_fwd: V=%d, F=%d must be <= 65535", V, F);
  RasterParams p;
  p.ndc = ndc; p.faces = faces; p.faces_stride = faces_batch_stride;
  p.N = N; p.V = V; p.F = F; p.H = H; p.W = W; p.K = K;
  p.blur = blur_radius; p.sq_blur = sqrtf(blur_radius); p.sigma = sigma;
  p.clip = clip_bary; p.cull = cull_backfaces;
  p.k_eps = acfm_raster_epsilon();
  p.loss_target = loss_target; p.loss_edt = loss_edt; p.NB = NB > 0 ? NB : 1; p.loss_part = nullptr;
  p.lean_f = nullptr; p.lean_d = nullptr;
  p.p2f = (long long*)pix_to_face; p.zbuf = zbuf; p.dists = dists; p.bary = bary; p.mask = mask; p.vis = visible_verts;
  p.regions_x = (W + kRegion - 1) / kRegion; p.regions_y = (H + kRegion - 1) / kRegion;
  p.vec_ok = ((((uintptr_t)pix_to_face) | ((uintptr_t)zbuf) | ((uintptr_t)dists)) & 15u) == 0;
  int smem = 0, cap = 0;
  const int nw = fwd_config(V, F, K, &smem, &cap);
  // the render's vertices (12 V bytes) and face list (4 F) are staged in one CTA's shared memory: meshes up to about
  // 12 V + 4 F <= 190 KB (e.g. 10 k vertices / 20 k faces) at K = 20; the 16-bit face ids allow 65 535 of each at most
  ACFM_REQUIRE(nw > 0, ACFM_ERR_UNSUPPORTED,
               "acfm_raster_fwd: V=%d F=%d K=%d needs more than the 232448 B of shared memory of one CTA (12 V + 4 F bytes of staged "
               "mesh + the per-pixel sets: about V <= 10000 / F <= 20000 at K = 20)", V, F, K);
  p.cap = cap;
  p.work = nullptr;
  const long long ctas = (long long)N * p.regions_x * p.regions_y;
  ACFM_REQUIRE(ctas < (1ll << 28), ACFM_ERR_UNSUPPORTED, "acfm_raster_fwd: too many CTAs");
  cudaStream_t st = (cudaStream_t)stream;
  if (lean) {
    // lean mode: the K-nearest-set kernel only (the reference's K = 20), work lists required, fragments to the compact scratch
    ACFM_REQUIRE(fwd_sets(K, nw), ACFM_ERR_UNSUPPORTED, "acfm_raster_fwd_lean: built for faces_per_pixel = 20 (got K=%d, V=%d, F=%d)", K, V, F);
    ACFM_REQUIRE(workspace && mask && F < 65535, ACFM_ERR_BAD_ARG, "acfm_raster_fwd_lean: needs the workspace, a mask output and F < 65535");
    ACFM_REQUIRE(lean_bytes >= ctas * (long long)(kRegion * kRegion) * K * 6 && (((uintptr_t)lean_workspace) & 15u) == 0, ACFM_ERR_BAD_ARG,
                 "acfm_raster_fwd_lean: lean workspace must be 16-byte aligned and hold acfm_raster_lean_workspace_bytes()");
    p.lean_d = (float*)lean_workspace;                                                        // [U][1024][K] f32
    p.lean_f = (unsigned short*)((float*)lean_workspace + ctas * (long long)(kRegion * kRegion) * K);   // [U][1024][K] u16
    ACFM_CUDA_OK(cudaMemsetAsync(mask, 0, sizeof(float) * (size_t)N * H * W, st));         // regions and tiles without fragments
  }
  if (visible_verts && V > 0) ACFM_CUDA_OK(cudaMemsetAsync(visible_verts, 0, sizeof(float) * (size_t)N * V, st));
  if (loss_sums) {
    // fused losses: region partials (zero where nothing is rendered) + per-target base sums, both in the caller's scratch
    ACFM_REQUIRE(loss_workspace_bytes >= 16 * ctas + 8 * (long long)NB && (((uintptr_t)loss_workspace) & 15u) == 0, ACFM_ERR_BAD_ARG,
                 "acfm_raster_fwd_train: loss workspace must be 16-byte aligned and hold acfm_raster_loss_workspace_bytes()");
    p.loss_part = (float*)loss_workspace;
    ACFM_CUDA_OK(cudaMemsetAsync(p.loss_part, 0, 16 * (size_t)ctas, st));
    loss_target_base_kernel<<<NB, 256, 0, st>>>(loss_target, H * W, p.loss_part + 4 * ctas);
    ACFM_LAUNCH_OK("loss_target_base_kernel");
  }
  // split path (see raster_prep_kernel): needs the caller's scratch; without it the rasterizer kernel pads the empty regions
  const FwdTuning& tune = fwd_tuning();
  p.bulk_ok = 0; p.pad_balance = tune.pad_balance;
  bool overlap = false;
  if (workspace) {
    ACFM_REQUIRE(workspace_bytes >= 32 + 24 * ctas && (((uintptr_t)workspace) & 15u) == 0, ACFM_ERR_BAD_ARG,
                 "acfm_raster_fwd: workspace must be 16-byte aligned and hold acfm_raster_fwd_workspace_bytes() = %lld bytes",
                 32 + 24 * ctas);
    int* ws = (int*)workspace;
    ACFM_CUDA_OK(cudaMemsetAsync(ws, 0, 32, st));
    const int nreg = p.regions_x * p.regions_y;
    const int weigh = nreg * 4 <= 40 * 1024 && F > 0;  // the per-region counters must fit the default shared-memory window
    if (faces_i64) raster_prep_kernel<long long><<<N, kPrepThreads, weigh ? nreg * 4 : 0, st>>>(p, ws, weigh);
    else raster_prep_kernel<int><<<N, kPrepThreads, weigh ? nreg * 4 : 0, st>>>(p, ws, weigh);
    ACFM_LAUNCH_OK("raster_prep_kernel");
    p.work = ws;
    int sms = 0;
    if (device_sm_count(&sms) != ACFM_OK) return ACFM_ERR_CUDA;
    // bulk stores need 16-byte aligned rows: aligned bases, W*K (and so every row start) a multiple of 4 fragment slots
    p.bulk_ok = p.vec_ok && (((long long)W * K) & 3) == 0 && (!bary || (((uintptr_t)bary) & 15u) == 0) && !tune.fill_lsu;
    // at most one wave, one CTA per SM: all of them are resident (and have signalled) before the rasterizer is dispatched
    const int fill_ctas = tune.fill_abs > 0 ? tune.fill_abs : (int)std::min<long long>(ctas, (long long)sms * tune.fill_per_sm);
    if (tune.only != 'r' && !lean) {
      raster_fill_kernel<<<fill_ctas, kFillThreads, fill_pattern_slots(K) * 12, st>>>(p, ws);
      ACFM_LAUNCH_OK("raster_fill_kernel");
      overlap = !tune.no_pdl;
    }
  }
  int rc = ACFM_ERR_UNSUPPORTED;
#define ACFM_FWD_CASE(NW, KT)                                                                          \
  rc = faces_i64 ? launch_fwd<NW, long long, KT>(p, smem, (int)ctas, overlap, st) : launch_fwd<NW, int, KT>(p, smem, (int)ctas, overlap, st)
  if (workspace && tune.only == 'f') rc = ACFM_OK;
  else if (lean) rc = faces_i64 ? launch_fwd<8, long long, 20, true>(p, smem, (int)ctas, false, st) : launch_fwd<8, int, 20, true>(p, smem, (int)ctas, false, st);
  else if (fwd_sets(K, nw)) ACFM_FWD_CASE(8, 20);  // the reference's faces_per_pixel: straight-line set operations
  else if (nw == 8 && K == 1) ACFM_FWD_CASE(8, 1);    // its hard renders (texture branch, OF_NeuralRenderer): z-buffer in registers
  else if (nw == 8) ACFM_FWD_CASE(8, 0);
  else if (nw == 4) ACFM_FWD_CASE(4, 0);
#undef ACFM_FWD_CASE
  ACFM_REQUIRE(rc != ACFM_ERR_UNSUPPORTED, ACFM_ERR_UNSUPPORTED, "acfm_raster_fwd: no launch configuration");
  if (rc == ACFM_OK && loss_sums) {
    loss_reduce_kernel<<<N, 32, 0, st>>>(p.loss_part, p.loss_part + 4 * ctas, p.regions_x * p.regions_y, NB, loss_sums);
    ACFM_LAUNCH_OK("loss_reduce_kernel");
  }
  return rc;
}
}  // namespace

extern "C" int64_t acfm_raster_lean_workspace_bytes(int N, int H, int W, int K) {
  if (N <= 0 || H <= 0 || W <= 0 || K <= 0) return 0;
  return (int64_t)N * ((W + kRegion - 1) / kRegion) * ((H + kRegion - 1) / kRegion) * (kRegion * kRegion) * K * 6;
}

extern "C" int acfm_raster_fwd_lean(const float* ndc, const void* faces, int faces_i64, int64_t faces_batch_stride, int N, int V, int F,
                                    int H, int W, int K, float blur_radius, float sigma, float* mask, float* visible_verts,
                                    const float* target, const float* edt, int NB, float* loss_sums, void* loss_workspace,
                                    int64_t loss_workspace_bytes, void* lean_workspace, int64_t lean_workspace_bytes, void* workspace,
                                    int64_t workspace_bytes, void* stream) {
  ACFM_REQUIRE(mask && lean_workspace && workspace, ACFM_ERR_BAD_ARG, "acfm_raster_fwd_lean: null mask / lean workspace / workspace pointer");
  ACFM_REQUIRE(!loss_sums || (target && loss_workspace), ACFM_ERR_BAD_ARG, "acfm_raster_fwd_lean: loss_sums needs target and loss_workspace");
  ACFM_REQUIRE(!loss_sums || (NB > 0 && N % NB == 0), ACFM_ERR_BAD_ARG, "acfm_raster_fwd_lean: N=%d is not a multiple of NB=%d", N, NB);
  return raster_fwd_impl(ndc, faces, faces_i64, faces_batch_stride, N, V, F, H, W, K, blur_radius, 0, 0, sigma, nullptr, nullptr, nullptr,
                         nullptr, mask, visible_verts, loss_sums ? target : nullptr, loss_sums ? edt : nullptr, NB, loss_sums,
                         loss_workspace, loss_workspace_bytes, workspace, workspace_bytes, stream, lean_workspace, lean_workspace_bytes);
}

extern "C" int64_t acfm_raster_fwd_workspace_bytes(int N, int H, int W) {
  if (N <= 0 || H <= 0 || W <= 0) return 0;
  return 32 + 24 * (int64_t)N * ((W + kRegion - 1) / kRegion) * ((H + kRegion - 1) / kRegion);
}
