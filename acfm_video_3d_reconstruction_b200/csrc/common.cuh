// common.cuh — shared device/host helpers for libacfm_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/acfm_b200.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "libacfm_b200 is written for sm_100a (B200) only"
#endif

// ---------------------------------------------------------------------------------------------
// error plumbing (host)
// ---------------------------------------------------------------------------------------------
void acfm_set_error(const char* fmt, ...);

#define ACFM_REQUIRE(cond, code, ...)  \
  do {                                 \
    if (!(cond)) {                     \
      acfm_set_error(__VA_ARGS__);     \
      return (code);                   \
    }                                  \
  } while (0)

#define ACFM_CUDA_OK(expr)                                                                  \
  do {                                                                                      \
    cudaError_t e__ = (expr);                                                               \
    if (e__ != cudaSuccess) {                                                               \
      acfm_set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__, __LINE__); \
      return ACFM_ERR_CUDA;                                                                 \
    }                                                                                       \
  } while (0)

#define ACFM_LAUNCH_OK(name)                                                         \
  do {                                                                               \
    cudaError_t e__ = cudaGetLastError();                                            \
    if (e__ != cudaSuccess) {                                                        \
      acfm_set_error("launch of %s failed: %s", name, cudaGetErrorString(e__));      \
      return ACFM_ERR_CUDA;                                                          \
    }                                                                                \
  } while (0)

// Raise a kernel's dynamic shared-memory limit once per (kernel instantiation, device, size): keeps
// cudaFuncSetAttribute out of the steady-state launch path (host time; and nothing but launches inside a CUDA-graph
// capture).  `slot` is a per-instantiation static array, one entry per device.
#include <atomic>
constexpr int kAcfmMaxDevices = 64;
template <typename Kern>
inline cudaError_t acfm_ensure_smem(Kern kern, int bytes, std::atomic<int>* slot) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  std::atomic<int>& cur = slot[dev & (kAcfmMaxDevices - 1)];
  if (bytes <= cur.load(std::memory_order_acquire)) return cudaSuccess;
  e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (e == cudaSuccess) cur.store(bytes, std::memory_order_release);
  return e;
}

// ---------------------------------------------------------------------------------------------
// strict IEEE fp32: one rounding per operator, never contracted into FMA.  The rasterizer and
// the projection reproduce the reference's CPU operator order bit for bit.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float fmul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float fadd(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float fsub(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float fdiv(float a, float b) { return __fdiv_rn(a, b); }

// PyTorch3D kEpsilon: 1e-8 in the release the reference pins (0.3.0, as restated); releases before 0.2
// used 1e-30.  A run-time setting of the rasterizer (acfm_set_raster_epsilon), because the depth order of near-coplanar
// neighbours is sensitive to it (tests/test_oracle_variants.py) and the value cannot be checked against the upstream source here.
#define ACFM_K_EPS_DEFAULT 1e-8f
float acfm_raster_epsilon();  // host: the current setting (api.cu)
int acfm_raster_bwd_headroom_bits();  // host: acfm_set_raster_bwd_headroom_bits (api.cu; raster_bwd.cu)

// PixToNdc: -1 + (2 i + 1) / S
__device__ __forceinline__ float pix_to_ndc(int i, int S) {
  return fadd(-1.0f, fdiv(fadd((float)(2 * i), 1.0f), (float)S));
}

// EdgeFunctionForward(p, a, b)
__device__ __forceinline__ float edge_fn(float px, float py, float ax, float ay, float bx, float by) {
  return fsub(fmul(fsub(px, ax), fsub(by, ay)), fmul(fsub(py, ay), fsub(bx, ax)));
}

// ---------------------------------------------------------------------------------------------
// TMA (bulk async copy) + mbarrier, raw PTX.  1-D cp.async.bulk needs 16-byte aligned
// source, destination and size; stage_bulk_1d() below copies an arbitrary 4-byte aligned span by
// sending the aligned interior through the TMA unit and the (<16 B) head/tail through the LSU.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t phase) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra.uni WAIT_DONE;\n\t"
      "bra.uni WAIT_LOOP;\n\t"
      "WAIT_DONE:\n\t}" ::"r"(smem_u32(bar)), "r"(phase) : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// Copy `bytes` (multiple of 4) from 4-byte aligned global `src` into shared memory so that the
// copy lands at buf + (src & 15): returns that shared pointer.  `buf` must be 16-byte aligned and
// hold bytes + 16.  Call from all threads of the CTA; `bar` must be initialised (count 1) and
// `phase` is the parity to wait on.  All threads return after the data is visible.
__device__ __forceinline__ const float* stage_bulk_1d(unsigned char* buf, const void* src, uint32_t bytes,
                                                      uint64_t* bar, uint32_t phase) {
  const uintptr_t s = (uintptr_t)src;
  const uint32_t mis = (uint32_t)(s & 15u);
  unsigned char* dst = buf + mis;
  const uint32_t head = mis ? min(16u - mis, bytes) : 0u;
  const uint32_t body = (bytes - head) & ~15u;
  const uint32_t tail = bytes - head - body;
  if (threadIdx.x == 0) {
    if (body) {
      mbar_expect_tx(bar, body);
      tma_bulk_g2s(dst + head, (const unsigned char*)src + head, body, bar);
    }
  }
  // head/tail: at most 3 + 3 words, through the LSU
  const uint32_t hw = head >> 2, tw = tail >> 2;
  if (threadIdx.x < hw) ((float*)dst)[threadIdx.x] = ((const float*)src)[threadIdx.x];
  if (threadIdx.x >= 32 && threadIdx.x < 32 + tw) {
    const uint32_t o = ((head + body) >> 2) + (threadIdx.x - 32);
    ((float*)dst)[o] = ((const float*)src)[o];
  }
  if (body) mbar_wait(bar, phase);
  __syncthreads();
  return (const float*)dst;
}
