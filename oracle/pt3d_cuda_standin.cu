/*
 * oracle/pt3d_cuda_standin.cu — TEST / BENCH INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * A plain restatement of the CUDA path PyTorch3D 0.3.0 runs for the reference's mask render
 * (/root/reference/multiframe/nnutils/nmr.py:143-172 -> MeshRasterizer -> _C.rasterize_meshes with bin_size=None, then
 * _C.rasterize_meshes_backward), as SURVEY.md section 2b lists it:
 *   coarse   RasterizeMeshesCoarseCuda: blocks walk chunks of 512 faces, one thread per face, blur-expanded bounding box
 *            against every bin (16 x 16 pixels at 256^2), a shared-memory bit mask per chunk, then one thread per bin
 *            reserves room in bin_faces with a global atomic and writes the face ids;
 *   fine     RasterizeMeshesFineCuda: ONE THREAD PER PIXEL walks its bin's face list, keeps the K nearest fragments in a
 *            per-thread local-memory array (replace-the-farthest), bubble-sorts them and writes pix_to_face (int64), zbuf,
 *            bary (3 floats) and dists;
 *   backward RasterizeMeshesBackwardCuda: one thread per pixel, per fragment recompute of the closest edge and up to four
 *            global fp32 atomicAdd into grad_face_verts (N*F,3,3).
 * The shader chain around it (verts_packed[faces_packed], sigmoid_alpha_blend and their autograd) stays torch ops, in
 * bench.py, as in PyTorch3D.  PyTorch3D itself is not in /root/reference and cannot be installed (SURVEY.md read-first 1):
 * this is a STAND-IN that gives BASELINE.json's ">= 50x the PyTorch3D CUDA path" a denominator on the same B200.  It is
 * compiled with -O3 for sm_100a (PyTorch3D's own wheels stop at sm_80 + PTX), uses the same arithmetic as the C oracle
 * (checked against it in tests/test_standin_gpu.py) and is timed at the library's defaults and, separately, with
 * max_faces_per_bin tightened to what the mesh needs (the default, max(10000, V/5), makes the fine kernel scan 10000 slots
 * per pixel) — bench.py reports both and says which is which.
 *
 * Only tests/ and bench.py (gpu_standin leg) load the library built from this file (oracle/Makefile -> oracle/_build/).
 */
#include <cuda_runtime.h>
#include <stdint.h>

namespace {

constexpr float kEps = 1e-8f;
constexpr int kMaxK = 64;  // per-thread fragment array (PyTorch3D: kMaxPointsPerPixel = 150)

__device__ __forceinline__ float pix_to_ndc(int i, int S) { return -1.0f + (2.0f * i + 1.0f) / S; }
__device__ __forceinline__ float edge_fn(float px, float py, float ax, float ay, float bx, float by) {
  return (px - ax) * (by - ay) - (py - ay) * (bx - ax);
}
__device__ __forceinline__ float point_line_dist(float px, float py, float ax, float ay, float bx, float by) {
  const float bax = bx - ax, bay = by - ay;
  const float l2 = bax * bax + bay * bay;
  if (l2 <= kEps) return (px - bx) * (px - bx) + (py - by) * (py - by);
  float t = (bax * (px - ax) + bay * (py - ay)) / l2;
  t = t < 0.f ? 0.f : (t > 1.f ? 1.f : t);
  const float qx = ax + t * bax, qy = ay + t * bay;
  return (px - qx) * (px - qx) + (py - qy) * (py - qy);
}

// ---- coarse -------------------------------------------------------------------------------------------------------------
__global__ void coarse_kernel(const float* __restrict__ face_verts, int N, int F, int H, int W, float blur, int bin_size,
                              int chunk_size, int M, int* __restrict__ faces_per_bin, int* __restrict__ bin_faces) {
  extern __shared__ unsigned int bits[];  // [bins_y * bins_x][chunk_size / 32]
  const int bx_n = 1 + (W - 1) / bin_size, by_n = 1 + (H - 1) / bin_size;
  const int words = chunk_size / 32;
  const float half_x = 1.0f / W, half_y = 1.0f / H;
  const float sq = sqrtf(blur);
  const int chunks_per_mesh = 1 + (F - 1) / chunk_size;
  for (int chunk = blockIdx.x; chunk < N * chunks_per_mesh; chunk += gridDim.x) {
    const int n = chunk / chunks_per_mesh, f0 = (chunk % chunks_per_mesh) * chunk_size;
    for (int i = threadIdx.x; i < by_n * bx_n * words; i += blockDim.x) bits[i] = 0u;
    __syncthreads();
    for (int f = threadIdx.x; f < chunk_size; f += blockDim.x) {
      if (f0 + f >= F) continue;
      const float* v = face_verts + ((size_t)n * F + f0 + f) * 9;
      const float xmin = fminf(fminf(v[0], v[3]), v[6]) - sq, xmax = fmaxf(fmaxf(v[0], v[3]), v[6]) + sq;
      const float ymin = fminf(fminf(v[1], v[4]), v[7]) - sq, ymax = fmaxf(fmaxf(v[1], v[4]), v[7]) + sq;
      if (fmaxf(fmaxf(v[2], v[5]), v[8]) < 0.f) continue;  // behind the camera
      for (int by = 0; by < by_n; ++by) {
        const float y_lo = pix_to_ndc(by * bin_size, H) - half_y, y_hi = pix_to_ndc((by + 1) * bin_size - 1, H) + half_y;
        const bool y_ov = (ymin <= y_hi) && (y_lo < ymax);
        for (int bx = 0; bx < bx_n; ++bx) {
          const float x_lo = pix_to_ndc(bx * bin_size, W) - half_x, x_hi = pix_to_ndc((bx + 1) * bin_size - 1, W) + half_x;
          if (y_ov && (xmin <= x_hi) && (x_lo < xmax)) atomicOr(&bits[(by * bx_n + bx) * words + (f >> 5)], 1u << (f & 31));
        }
      }
    }
    __syncthreads();
    for (int b = threadIdx.x; b < by_n * bx_n; b += blockDim.x) {
      int count = 0;
      for (int w = 0; w < words; ++w) count += __popc(bits[b * words + w]);
      const size_t bin = (size_t)n * by_n * bx_n + b;
      int next = atomicAdd(&faces_per_bin[bin], count);
      for (int f = 0; f < chunk_size; ++f)
        if ((bits[b * words + (f >> 5)] >> (f & 31)) & 1u) {
          if (next < M) bin_faces[bin * M + next] = f0 + f;
          ++next;
        }
    }
    __syncthreads();
  }
}

// ---- fine ---------------------------------------------------------------------------------------------------------------
struct Pix { float z; int64_t idx; float dist; float b0, b1, b2; };

__global__ void fine_kernel(const float* __restrict__ face_verts, const int* __restrict__ bin_faces, float blur, int bin_size, int N,
                            int F, int M, int H, int W, int K, int64_t* __restrict__ p2f, float* __restrict__ zbuf,
                            float* __restrict__ dists, float* __restrict__ bary) {
  const int B = 1 + (W - 1) / bin_size;
  const long long num = (long long)N * B * B * bin_size * bin_size;
  for (long long pid = blockIdx.x * (long long)blockDim.x + threadIdx.x; pid < num; pid += (long long)gridDim.x * blockDim.x) {
    long long i = pid;
    const int n = (int)(i / ((long long)B * B * bin_size * bin_size)); i %= (long long)B * B * bin_size * bin_size;
    const int by = (int)(i / (B * bin_size * bin_size)); i %= B * bin_size * bin_size;
    const int bx = (int)(i / (bin_size * bin_size)); i %= bin_size * bin_size;
    const int yi = (int)(i / bin_size) + by * bin_size, xi = (int)(i % bin_size) + bx * bin_size;
    if (yi >= H || xi >= W) continue;
    const float xf = pix_to_ndc(xi, W), yf = pix_to_ndc(yi, H);
    Pix q[kMaxK];
    int qn = 0, qmax_i = -1;
    float qmax_z = -1000.f;
    const int* bf = bin_faces + ((size_t)n * B * B + by * B + bx) * M;
    for (int m = 0; m < M; ++m) {
      const int f = bf[m];
      if (f < 0) continue;  // -1 is the sentinel (the list is not compacted to its length: every slot is read)
      const float* v = face_verts + ((size_t)n * F + f) * 9;
      const float x0 = v[0], y0 = v[1], z0 = v[2], x1 = v[3], y1 = v[4], z1 = v[5], x2 = v[6], y2 = v[7], z2 = v[8];
      const float area = edge_fn(x0, y0, x1, y1, x2, y2);
      const float sq = sqrtf(blur);
      const bool outside = xf > fmaxf(fmaxf(x0, x1), x2) + sq || xf < fminf(fminf(x0, x1), x2) - sq ||
                           yf > fmaxf(fmaxf(y0, y1), y2) + sq || yf < fminf(fminf(y0, y1), y2) - sq;
      if (outside || (area <= kEps && area >= -kEps) || fmaxf(fmaxf(z0, z1), z2) < 0.f) continue;
      const float den = edge_fn(x2, y2, x0, y0, x1, y1) + kEps;
      const float w0 = edge_fn(xf, yf, x1, y1, x2, y2) / den, w1 = edge_fn(xf, yf, x2, y2, x0, y0) / den,
                  w2 = edge_fn(xf, yf, x0, y0, x1, y1) / den;
      const float pz = w0 * z0 + w1 * z1 + w2 * z2;
      if (pz < 0.f) continue;
      const float d = fminf(fminf(point_line_dist(xf, yf, x0, y0, x1, y1), point_line_dist(xf, yf, x0, y0, x2, y2)),
                            point_line_dist(xf, yf, x1, y1, x2, y2));
      const bool inside = w0 > 0.f && w1 > 0.f && w2 > 0.f;
      if (!inside && d >= blur) continue;
      const Pix c = {pz, (int64_t)n * F + f, inside ? -d : d, w0, w1, w2};
      if (qn < K) {
        q[qn] = c;
        if (pz > qmax_z) { qmax_z = pz; qmax_i = qn; }
        ++qn;
      } else if (pz < qmax_z) {
        q[qmax_i] = c;
        qmax_z = pz;
        for (int j = 0; j < K; ++j)
          if (q[j].z > qmax_z) { qmax_z = q[j].z; qmax_i = j; }
      }
    }
    for (int a = 0; a + 1 < qn; ++a)  // BubbleSort
      for (int b = 0; b + 1 < qn - a; ++b)
        if (q[b + 1].z < q[b].z) { const Pix t = q[b]; q[b] = q[b + 1]; q[b + 1] = t; }
    const size_t o = (((size_t)n * H + (H - 1 - yi)) * W + (W - 1 - xi)) * K;
    for (int k = 0; k < qn; ++k) {
      p2f[o + k] = q[k].idx; zbuf[o + k] = q[k].z; dists[o + k] = q[k].dist;
      bary[(o + k) * 3] = q[k].b0; bary[(o + k) * 3 + 1] = q[k].b1; bary[(o + k) * 3 + 2] = q[k].b2;
    }
  }
}

// ---- backward (gradient on dists only: the silhouette shader uses nothing else) ------------------------------------------
__device__ __forceinline__ void seg_bwd(float px, float py, const float* a, const float* b, float g, float* ga, float* gb) {
  const float bax = b[0] - a[0], bay = b[1] - a[1];
  const float l2 = bax * bax + bay * bay;
  float t = (bax * (px - a[0]) + bay * (py - a[1])) / l2;
  t = t < 0.f ? 0.f : (t > 1.f ? 1.f : t);
  const float qx = (1.f - t) * a[0] + t * b[0], qy = (1.f - t) * a[1] + t * b[1];
  atomicAdd(ga, g * (1.f - t) * 2.f * (qx - px)); atomicAdd(ga + 1, g * (1.f - t) * 2.f * (qy - py));
  atomicAdd(gb, g * t * 2.f * (qx - px)); atomicAdd(gb + 1, g * t * 2.f * (qy - py));
}

__global__ void backward_kernel(const float* __restrict__ face_verts, const int64_t* __restrict__ p2f, const float* __restrict__ grad_dists,
                                int N, int H, int W, int K, float* __restrict__ grad_face_verts) {
  const long long num = (long long)N * H * W;
  for (long long pix = blockIdx.x * (long long)blockDim.x + threadIdx.x; pix < num; pix += (long long)gridDim.x * blockDim.x) {
    const int xi = W - 1 - (int)(pix % W), yi = H - 1 - (int)((pix / W) % H);
    const float xf = pix_to_ndc(xi, W), yf = pix_to_ndc(yi, H);
    for (int k = 0; k < K; ++k) {
      const int64_t f = p2f[pix * K + k];
      if (f < 0) break;
      const float* v = face_verts + f * 9;
      float* g = grad_face_verts + f * 9;
      const float den = edge_fn(v[6], v[7], v[0], v[1], v[3], v[4]) + kEps;
      const bool inside = edge_fn(xf, yf, v[3], v[4], v[6], v[7]) / den > 0.f && edge_fn(xf, yf, v[6], v[7], v[0], v[1]) / den > 0.f &&
                          edge_fn(xf, yf, v[0], v[1], v[3], v[4]) / den > 0.f;
      const float gd = inside ? -grad_dists[pix * K + k] : grad_dists[pix * K + k];
      if (gd == 0.f) continue;
      const float d01 = point_line_dist(xf, yf, v[0], v[1], v[3], v[4]), d02 = point_line_dist(xf, yf, v[0], v[1], v[6], v[7]),
                  d12 = point_line_dist(xf, yf, v[3], v[4], v[6], v[7]);
      if (d01 <= d02 && d01 <= d12) seg_bwd(xf, yf, v, v + 3, gd, g, g + 3);
      else if (d02 <= d01 && d02 <= d12) seg_bwd(xf, yf, v, v + 6, gd, g, g + 6);
      else seg_bwd(xf, yf, v + 3, v + 6, gd, g + 3, g + 6);
    }
  }
}

}  // namespace

// face_verts (N*F,3,3) device.  Workspace: faces_per_bin (N*B*B) int32 + bin_faces (N*B*B*M) int32, caller-allocated.
// Outputs must be pre-filled with -1 by the caller (torch.full, as PyTorch3D does).  Returns a cudaError_t.
extern "C" int acfm_standin_rasterize(const float* face_verts, int N, int F, int H, int W, float blur, int K, int bin_size, int M,
                                      int* faces_per_bin, int* bin_faces, int64_t* p2f, float* zbuf, float* dists, float* bary,
                                      void* stream) {
  if (K > kMaxK || H != W) return (int)cudaErrorInvalidValue;
  cudaStream_t st = (cudaStream_t)stream;
  const int B = 1 + (W - 1) / bin_size, chunk = 512;
  cudaMemsetAsync(faces_per_bin, 0, sizeof(int) * (size_t)N * B * B, st);
  cudaMemsetAsync(bin_faces, 0xff, sizeof(int) * (size_t)N * B * B * M, st);
  coarse_kernel<<<64, 512, (size_t)B * B * (chunk / 32) * 4, st>>>(face_verts, N, F, H, W, blur, bin_size, chunk, M, faces_per_bin, bin_faces);
  fine_kernel<<<1024, 64, 0, st>>>(face_verts, bin_faces, blur, bin_size, N, F, M, H, W, K, p2f, zbuf, dists, bary);
  return (int)cudaGetLastError();
}

extern "C" int acfm_standin_rasterize_backward(const float* face_verts, const int64_t* p2f, const float* grad_dists, int N, int H, int W,
                                               int K, float* grad_face_verts, void* stream) {
  backward_kernel<<<1024, 64, 0, (cudaStream_t)stream>>>(face_verts, p2f, grad_dists, N, H, W, K, grad_face_verts);
  return (int)cudaGetLastError();
}
