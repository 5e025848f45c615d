// Shared device/host helpers for the vggsfm_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include "../../include/vggsfm_b200.h"

namespace vgg {

void set_error(const char* fmt, ...);
extern thread_local long long g_launch_count;   // kernels launched by the current C-ABI call

#define VGG_CUDA_CHECK(expr)                                                              \
  do {                                                                                    \
    cudaError_t _e = (expr);                                                              \
    if (_e != cudaSuccess) {                                                              \
      vgg::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
      return VGG_ECUDA;                                                                   \
    }                                                                                     \
  } while (0)

#define VGG_LAUNCH_CHECK()                 \
  do {                                     \
    vgg::g_launch_count++;                 \
    VGG_CUDA_CHECK(cudaGetLastError());    \
  } while (0)

#define VGG_REQUIRE(cond, msg)                                   \
  do {                                                           \
    if (!(cond)) {                                               \
      vgg::set_error("%s:%d: %s", __FILE__, __LINE__, msg);      \
      return VGG_EINVAL;                                         \
    }                                                            \
  } while (0)

__host__ __device__ constexpr inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// bump allocator over the caller-provided workspace
struct Carver {
  char* base;
  size_t off = 0, cap;
  Carver(void* p, size_t c) : base(static_cast<char*>(p)), cap(c) {}
  template <typename T>
  T* take(size_t n) {
    off = align_up(off, 256);
    T* r = reinterpret_cast<T*>(base + off);
    off += n * sizeof(T);
    return r;
  }
  bool ok() const { return base == nullptr || off <= cap; }
};

// Destination table of the fused multi-GPU reduction (csrc/fabric.cu): world <= 1 means "not in use".
// Band structure of a sequential (video) problem on the device, valid for the duration of one solve (csrc/ba_solve.cu,
// compute_band_hint); all pointers null = dense.  Points are stored in creation order, so every range is contiguous.
//   rb_range[2 rb]  .. [2 rb + 1]   64-row k-block range of Zt outside which 128-column row block rb is zero
//   kb_rows[2 kb]   .. [2 kb + 1]   reduced-system row range [lo, hi) that the points of k-block kb can touch (band part)
//   fg_tracks[2 g]  .. [2 g + 1]    track range [lo, hi) visible to the 32-frame group g
struct BandDev {
  const int* rb_range;
  const int* kb_rows;
  const int* fg_tracks;
  int arrow_row;                     // first row of the dense arrow (shared intrinsics ...): always processed
};

struct FabricDev {
  int world, rank;
  double* peer[8];     // rank r's copy of the buffer being addressed (same layout on every rank, peer-mapped)
};

#ifdef __CUDACC__
// ---------------------------------------------------------------------------------------------
// TMA 1-D bulk copies + mbarrier (PTX ISA 8.x, sm_90+; SASS: UBLKCP / SYNCS)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE;\n"
      "bra WAIT_LOOP;\n"
      "DONE:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
// global -> shared, completion on mbarrier.  size multiple of 16, both addresses 16B aligned.
__device__ __forceinline__ void tma_load_1d(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
          smem_u32(smem_dst)),
      "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}
// shared -> global bulk store (bulk async-group completion)
__device__ __forceinline__ void tma_store_1d(void* gmem_dst, const void* smem_src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gmem_dst),
               "r"(smem_u32(smem_src)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void tma_store_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
template <int N>
__device__ __forceinline__ void tma_store_wait_all() {
  asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory");
}
// make generic-proxy shared-memory writes visible to the async proxy (TMA)
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// cp.async 16B (LDGSTS)
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async4(void* smem_dst, const void* gmem_src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(smem_dst)), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

// Warp reduce-scatter: every lane contributes v[0..KP), KP a power of two <= 32.  Returns in lane l
// the warp-wide sum of v[l & (KP-1)].  Costs KP-1 (+log2(32/KP)) 64-bit shuffles instead of 5*KP.
template <int KP>
__device__ __forceinline__ double warp_reduce_scatter(double (&v)[KP], int lane) {
#pragma unroll
  for (int off = KP / 2; off >= 1; off >>= 1) {
    const bool up = (lane & off) != 0;
#pragma unroll
    for (int i = 0; i < off; ++i) {
      const double mine = up ? v[i + off] : v[i];
      const double send = up ? v[i] : v[i + off];
      v[i] = mine + __shfl_xor_sync(0xffffffffu, send, off);
    }
  }
  double r = v[0];
#pragma unroll
  for (int off = KP; off < 32; off <<= 1) r += __shfl_xor_sync(0xffffffffu, r, off);
  return r;
}

__device__ __forceinline__ double warp_sum(double r) {
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) r += __shfl_xor_sync(0xffffffffu, r, off);
  return r;
}
__device__ __forceinline__ double warp_max(double r) {
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) r = fmax(r, __shfl_xor_sync(0xffffffffu, r, off));
  return r;
}
#endif  // __CUDACC__

}  // namespace vgg
