// Shared device/host helpers for librobseg_b200 (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdarg>
#include <cstdio>

#include "../../include/robseg_b200.h"

namespace robseg {

// ---- error plumbing: nothing throws across the C ABI -------------------------------------
void set_error(const char* fmt, ...);
int sm_count();  // cached multiProcessorCount of the current device (148 on B200)
// one-shot events of robseg_profile_next_kernel (thread-local): take_* returns the event and clears it
cudaEvent_t take_profile_start();
cudaEvent_t take_profile_stop();

#define ROBSEG_REQUIRE(cond, ...)          \
  do {                                     \
    if (!(cond)) {                         \
      ::robseg::set_error(__VA_ARGS__);    \
      return ROBSEG_EINVAL;                \
    }                                      \
  } while (0)

#define ROBSEG_CUDA(expr)                                                          \
  do {                                                                             \
    cudaError_t e_ = (expr);                                                       \
    if (e_ != cudaSuccess) {                                                       \
      ::robseg::set_error("%s failed: %s", #expr, cudaGetErrorString(e_));         \
      return static_cast<int>(e_);                                                 \
    }                                                                              \
  } while (0)

#define ROBSEG_LAUNCH_CHECK()                                                      \
  do {                                                                             \
    cudaError_t e_ = cudaGetLastError();                                           \
    if (e_ != cudaSuccess) {                                                       \
      ::robseg::set_error("kernel launch failed: %s", cudaGetErrorString(e_));     \
      return static_cast<int>(e_);                                                 \
    }                                                                              \
  } while (0)

// ---- device helpers -----------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}

// 3-D tiled TMA load global -> shared, completion on an mbarrier (SASS: UTMALDG).
__device__ __forceinline__ void tma_load_3d(void* smem_dst, const void* tmap, int c0, int c1,
                                            int c2, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(smem_u32(smem_dst)),
      "l"(tmap), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const void* tmap) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(tmap) : "memory");
}

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// Per-image class counters folded into the loss kernels' argmax pass (compute_iou_acc, semseg/attacker.py:9-52;
// the same counters robseg_pixel_hist returns): cnt = this image's [3][C] int64 block
// (intersection, target, prediction), updated with 64-bit reductions that return nothing (RED.E.ADD.64): integer,
// hence exact and order-independent.  Aggregation per warp row WITHOUT warp-match instructions (MATCH.ANY on 32
// distinct keys cost ~19 % of the loss kernel on uniformly random labels): twice, the first still-uncounted lane
// announces its class, every lane holding the same class is counted by that one lane (ballot + popc, one
// reduction); whatever is left after two rounds -- nothing on a coherent map, nothing on a two-class boundary row --
// issues its own reduction.  A pixel whose label is ignored contributes nothing (its prediction is ignored too,
// attacker.py:20).  Must be called by all 32 lanes.
// The reductions go to one of R copies of the [B][3][C] block (copy = blockIdx.x mod R, in the caller's workspace):
// all CTAs work on the same image at the same time, and with a single copy the L2 atomic units serialise on that
// image's 3*C addresses (uniformly random labels: +13 % on the loss kernel at C = 150, far more at C = 21).  R grows as
// C shrinks so that an image always spreads over >= ~2000 addresses.  A small kernel adds the copies into the caller's
// counts tensor afterwards (launch_counts_fold).
__host__ __device__ inline int count_replicas(int C) {
  int r = 8;
  while (r < 128 && r * 3 * C < 2048) r <<= 1;
  return r;
}
int launch_counts_fold(const unsigned long long* replicas, int R, int B, int C, int64_t* counts, cudaStream_t stream);
int launch_counts_zero(unsigned long long* replicas, size_t bytes, cudaStream_t stream);  // bytes: multiple of 16

__device__ __forceinline__ void count_keys(unsigned long long* arr, bool valid, int key) {
  const int lane = threadIdx.x & 31;
  unsigned left = __ballot_sync(0xffffffffu, valid);
#pragma unroll
  for (int round = 0; round < 2; ++round) {
    if (left == 0u) return;  // warp-uniform
    const int first = __ffs(left) - 1;
    const int k0 = __shfl_sync(0xffffffffu, key, first);
    const unsigned same = __ballot_sync(0xffffffffu, valid && key == k0) & left;
    if (lane == first) atomicAdd(arr + k0, (unsigned long long)__popc(same));
    left &= ~same;
  }
  if ((left >> lane) & 1u) atomicAdd(arr + key, 1ull);
}
__device__ __forceinline__ void count_pixel(unsigned long long* cnt, int C, bool valid, int t, int q) {
  // target and intersection share the label as their key: one extraction serves both arrays
  const int lane = threadIdx.x & 31;
  unsigned left = __ballot_sync(0xffffffffu, valid);
  const unsigned hits = __ballot_sync(0xffffffffu, valid && t == q);
#pragma unroll
  for (int round = 0; round < 2; ++round) {
    if (left == 0u) break;  // warp-uniform
    const int first = __ffs(left) - 1;
    const int k0 = __shfl_sync(0xffffffffu, t, first);
    const unsigned same = __ballot_sync(0xffffffffu, valid && t == k0) & left;
    if (lane == first) {
      atomicAdd(cnt + C + k0, (unsigned long long)__popc(same));
      const int h = __popc(same & hits);
      if (h) atomicAdd(cnt + k0, (unsigned long long)h);
    }
    left &= ~same;
  }
  if ((left >> lane) & 1u) {
    atomicAdd(cnt + C + t, 1ull);
    if (t == q) atomicAdd(cnt + t, 1ull);
  }
  count_keys(cnt + 2 * C, valid, q);  // prediction (only where the label is valid)
}
// All N pixels of every lane at once (t[j] < 0: pixel not counted).  Fast path for the common case on real label
// maps -- the whole warp tile lies inside one region, i.e. every (label, prediction) pair is the same: one compare
// per pixel, one vote, and lane 0 issues the (at most three) reductions for 32*N pixels.  The general path above costs
// ~45 instructions per pixel slot, which an issue-bound launch (C = 21: ~1500 instructions per tile) feels: +18 %.
template <int N>
__device__ __forceinline__ void count_pixels(unsigned long long* cnt, int C, const int (&t)[N], const int (&q)[N]) {
  if (C < 32768) {
    int key[N];
#pragma unroll
    for (int j = 0; j < N; ++j) key[j] = t[j] >= 0 ? ((t[j] << 16) | q[j]) : -1;
    const int k0 = __shfl_sync(0xffffffffu, key[0], 0);
    bool same = true;
#pragma unroll
    for (int j = 0; j < N; ++j) same = same && key[j] == k0;
    if (__all_sync(0xffffffffu, same)) {
      if ((threadIdx.x & 31) == 0 && k0 >= 0) {
        const int tt = k0 >> 16, qq = k0 & 0xffff;
        atomicAdd(cnt + C + tt, (unsigned long long)(32 * N));
        atomicAdd(cnt + 2 * C + qq, (unsigned long long)(32 * N));
        if (tt == qq) atomicAdd(cnt + tt, (unsigned long long)(32 * N));
      }
      return;
    }
  }
#pragma unroll
  for (int j = 0; j < N; ++j) count_pixel(cnt, C, t[j] >= 0, t[j], q[j]);
}

template <typename T>
__device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

}  // namespace robseg
