// Integer-exact metric counters: per-image confusion matrix / intersection / union histograms
// with warp-aggregated shared-memory atomics, and the SEA worst-case accuracy reduction.
//
// Replaces the 2*C-iteration masked reductions of compute_iou_acc (semseg/attacker.py:9-52),
// Metrics.update's bincount (semseg/metrics.py:27-33), eval_performance
// (tools/infer.py:86-116) and evalSEA (tools/worse_only.py:30-66,383-408).
// Bytes: 16 B per pixel (int64 pred + int64 label) read once; HBM-bound.
#include "common.cuh"

namespace robseg {

constexpr int kHistThreads = 256;
constexpr int kPairs = 4;                              // 16-byte loads in flight per array per thread
constexpr int kPxPerIter = kHistThreads * 2 * kPairs;  // 8 pixels per thread per iteration

// Leader aggregation: every lane compares its key with lane 0's; lane 0 adds the number of
// matching lanes with ONE shared-memory atomic, only the lanes that differ issue their own.
// Spatially coherent label maps (the normal case) collapse to one atomic per warp and slot;
// uniformly random keys degrade gracefully to one atomic per pixel (no match.any round trip).
__device__ __forceinline__ void leader_add(int* counters, int key) {
  const int k0 = __shfl_sync(0xffffffffu, key, 0);
  const bool same = key == k0;
  const unsigned m = __ballot_sync(0xffffffffu, same);
  if ((threadIdx.x & 31) == 0) {
    if (k0 >= 0) atomicAdd(counters + k0, __popc(m));
  } else if (!same && key >= 0) {
    atomicAdd(counters + key, 1);
  }
}

// FULL: shared [C*C] confusion tile, inter/tgt/prd derived from it at flush time (one atomic
// per pixel).  Otherwise 3*C counters (inter, tgt, prd) for class counts whose tile does not fit.
// grid (chunks, n_img); a block consumes px_per_block pixels of one image.
template <bool FULL>
__global__ void __launch_bounds__(kHistThreads)
    pixel_hist_kernel(const int64_t* __restrict__ pred, const int64_t* __restrict__ labels,
                      int n_lab_img, int64_t HW, int64_t px_per_block, int C, int ignore_index,
                      unsigned long long* hist, unsigned long long* hist_total,
                      unsigned long long* inter, unsigned long long* tgt,
                      unsigned long long* prd) {
  extern __shared__ int sh[];
  const int img = blockIdx.y;
  const int n_cnt = FULL ? C * C : 3 * C;
  for (int i = threadIdx.x; i < n_cnt; i += kHistThreads) sh[i] = 0;
  __syncthreads();
  const int64_t* pp = pred + (int64_t)img * HW;
  const int64_t* lp = labels + (int64_t)(img % n_lab_img) * HW;
  const int64_t p0 = (int64_t)blockIdx.x * px_per_block;
  int64_t p1 = p0 + px_per_block;
  if (p1 > HW) p1 = HW;
  const bool vec_ok = ((reinterpret_cast<uintptr_t>(pp) | reinterpret_cast<uintptr_t>(lp)) & 15) == 0 &&
                      (p0 & 1) == 0;
  // whole warps iterate together (uniform trip count) so the shuffles/ballots are well defined
  for (int64_t base = p0; base < p1; base += kPxPerIter) {
    int64_t tv[2 * kPairs], qv[2 * kPairs];
#pragma unroll
    for (int h = 0; h < kPairs; ++h) {
      // pair h of this thread: pixels base + h*512 + 2*tid, +1  (16-byte coalesced loads)
      const int64_t i = base + h * (2 * kHistThreads) + 2 * threadIdx.x;
      if (vec_ok && i + 1 < p1) {
        const longlong2 a = __ldcs(reinterpret_cast<const longlong2*>(lp + i));
        const longlong2 b = __ldcs(reinterpret_cast<const longlong2*>(pp + i));
        tv[2 * h] = a.x, tv[2 * h + 1] = a.y, qv[2 * h] = b.x, qv[2 * h + 1] = b.y;
      } else {
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const bool in = i + e < p1;
          tv[2 * h + e] = in ? __ldcs(lp + i + e) : -1;
          qv[2 * h + e] = in ? __ldcs(pp + i + e) : -1;
        }
      }
    }
#pragma unroll
    for (int e = 0; e < 2 * kPairs; ++e) {
      const int t = (tv[e] != ignore_index && tv[e] >= 0 && tv[e] < C) ? (int)tv[e] : -1;
      const int q = (qv[e] >= 0 && qv[e] < C) ? (int)qv[e] : -1;
      if constexpr (FULL) {
        leader_add(sh, (t >= 0 && q >= 0) ? t * C + q : -1);
      } else {
        leader_add(sh + C, t);                            // tgt
        leader_add(sh + 2 * C, (t >= 0) ? q : -1);        // prd (pred := ignore where target is)
        leader_add(sh, (t >= 0 && t == q) ? t : -1);      // inter
      }
    }
  }
  __syncthreads();
  if constexpr (FULL) {
    if (hist || hist_total) {
      for (int i = threadIdx.x; i < C * C; i += kHistThreads) {
        const int v = sh[i];
        if (v) {
          if (hist) atomicAdd(hist + (int64_t)img * C * C + i, (unsigned long long)v);
          if (hist_total) atomicAdd(hist_total + i, (unsigned long long)v);
        }
      }
    }
    if (inter || tgt || prd) {
      for (int c = threadIdx.x; c < C; c += kHistThreads) {
        int rs = 0, cs = 0;
        for (int k = 0; k < C; ++k) {
          const int kk = (k + c) % C;  // staggered start: conflict-free column walk
          rs += sh[c * C + kk], cs += sh[kk * C + c];
        }
        const int64_t o = (int64_t)img * C + c;
        if (inter && sh[c * C + c]) atomicAdd(inter + o, (unsigned long long)sh[c * C + c]);
        if (tgt && rs) atomicAdd(tgt + o, (unsigned long long)rs);
        if (prd && cs) atomicAdd(prd + o, (unsigned long long)cs);
      }
    }
  } else {
    for (int c = threadIdx.x; c < C; c += kHistThreads) {
      const int64_t o = (int64_t)img * C + c;
      if (inter && sh[c]) atomicAdd(inter + o, (unsigned long long)sh[c]);
      if (tgt && sh[C + c]) atomicAdd(tgt + o, (unsigned long long)sh[C + c]);
      if (prd && sh[2 * C + c]) atomicAdd(prd + o, (unsigned long long)sh[2 * C + c]);
    }
  }
}

// Fallback for class counts whose [C,C] tile does not fit in shared memory: warp-aggregated
// atomics straight to the global per-image histogram.
__global__ void __launch_bounds__(kHistThreads)
    pixel_hist_global_kernel(const int64_t* __restrict__ pred, const int64_t* __restrict__ labels,
                             int n_lab_img, int64_t HW, int64_t px_per_block, int C,
                             int ignore_index, unsigned long long* hist,
                             unsigned long long* hist_total) {
  const int img = blockIdx.y;
  const int64_t* pp = pred + (int64_t)img * HW;
  const int64_t* lp = labels + (int64_t)(img % n_lab_img) * HW;
  const int64_t p0 = (int64_t)blockIdx.x * px_per_block;
  int64_t p1 = p0 + px_per_block;
  if (p1 > HW) p1 = HW;
  for (int64_t base = p0 + (threadIdx.x & ~31); base < p1; base += kHistThreads) {
    const int64_t i = base + (threadIdx.x & 31);
    int t = -1, q = -1;
    if (i < p1) {
      const int64_t tv = __ldcs(lp + i), qv = __ldcs(pp + i);
      t = (tv != ignore_index && tv >= 0 && tv < C) ? (int)tv : -1;
      q = (qv >= 0 && qv < C) ? (int)qv : -1;
    }
    const bool active = t >= 0 && q >= 0;
    const unsigned mask = __ballot_sync(0xffffffffu, active);
    if (!active) continue;
    const int key = t * C + q;
    const unsigned peers = __match_any_sync(mask, key);
    if ((int)(threadIdx.x & 31) == __ffs(peers) - 1) {
      const unsigned long long n = (unsigned long long)__popc(peers);
      if (hist) atomicAdd(hist + (int64_t)img * C * C + key, n);
      if (hist_total) atomicAdd(hist_total + key, n);
    }
  }
}

// acc[a,n] = sum_c inter / sum_c tgt (fp32 division of exactly represented sums),
// worst[n] = min over attacks.  One thread per image.
__global__ void __launch_bounds__(256)
    sea_worst_acc_kernel(const int64_t* __restrict__ inter, const int64_t* __restrict__ tgt, int A,
                         int N, int C, float* acc_an, float* worst_n) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  float w = INFINITY;
  bool nan_seen = false;
  for (int a = 0; a < A; ++a) {
    const int64_t* ip = inter + ((int64_t)a * N + n) * C;
    const int64_t* tp = tgt + ((int64_t)a * N + n) * C;
    int64_t si = 0, st = 0;
    for (int c = 0; c < C; ++c) si += ip[c], st += tp[c];
    const float v = (float)si / (float)st;
    if (acc_an) acc_an[(int64_t)a * N + n] = v;
    nan_seen |= (v != v);
    w = fminf(w, v);
  }
  if (worst_n) worst_n[n] = nan_seen ? NAN : w;
}

}  // namespace robseg

using namespace robseg;

extern "C" int robseg_pixel_hist(const int64_t* pred, const int64_t* labels, int n_img,
                                 int n_lab_img, int64_t HW, int C, int ignore_index, int64_t* hist,
                                 int64_t* hist_total, int64_t* inter, int64_t* tgt, int64_t* prd,
                                 robseg_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  ROBSEG_REQUIRE(pred && labels, "NULL pointer");
  ROBSEG_REQUIRE(n_img > 0 && n_img <= 65535 && n_lab_img > 0 && HW > 0 && C > 0,
                 "bad shape n_img=%d HW=%lld C=%d", n_img, (long long)HW, C);
  ROBSEG_REQUIRE(hist || hist_total || inter || tgt || prd, "no output requested");
  const bool want_full = hist != nullptr || hist_total != nullptr;
  const size_t tile_bytes = (size_t)C * C * sizeof(int);
  const bool tile_fits = tile_bytes <= 200 * 1024;
  ROBSEG_REQUIRE((size_t)3 * C * sizeof(int) <= 200 * 1024, "C=%d too large", C);
  // pixels per block: a multiple of kPxPerIter, large enough to amortise the tile clear/flush,
  // small enough to give the grid ~2 blocks per SM
  // (one resident wave when the image count allows it: a second, nearly empty wave doubles the time)
  int chunks = (2 * sm_count()) / n_img;
  if (chunks < 1) chunks = 1;
  int64_t per_block = (HW + chunks - 1) / chunks;
  per_block = ((per_block + kPxPerIter - 1) / kPxPerIter) * kPxPerIter;
  if (per_block < 4 * kPxPerIter) per_block = 4 * kPxPerIter;
  dim3 grid((unsigned)((HW + per_block - 1) / per_block), n_img);
  auto u = [](int64_t* q) { return reinterpret_cast<unsigned long long*>(q); };
  if (tile_fits) {
    ROBSEG_CUDA(cudaFuncSetAttribute(pixel_hist_kernel<true>,
                                     cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tile_bytes));
    pixel_hist_kernel<true><<<grid, kHistThreads, tile_bytes, stream>>>(
        pred, labels, n_lab_img, HW, per_block, C, ignore_index, u(hist), u(hist_total), u(inter),
        u(tgt), u(prd));
  } else {
    // [C,C] tile does not fit: histogram through global atomics, counters through the 3C kernel
    if (want_full) {
      pixel_hist_global_kernel<<<grid, kHistThreads, 0, stream>>>(
          pred, labels, n_lab_img, HW, per_block, C, ignore_index, u(hist), u(hist_total));
      ROBSEG_LAUNCH_CHECK();
    }
    if (inter || tgt || prd) {
      pixel_hist_kernel<false><<<grid, kHistThreads, (size_t)3 * C * sizeof(int), stream>>>(
          pred, labels, n_lab_img, HW, per_block, C, ignore_index, nullptr, nullptr, u(inter),
          u(tgt), u(prd));
    }
  }
  ROBSEG_LAUNCH_CHECK();
  return 0;
}

extern "C" int robseg_sea_worst_acc(const int64_t* inter, const int64_t* tgt, int A, int N, int C,
                                    float* acc_an, float* worst_n, robseg_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  ROBSEG_REQUIRE(inter && tgt && (acc_an || worst_n), "NULL pointer");
  ROBSEG_REQUIRE(A > 0 && N > 0 && C > 0, "bad shape");
  sea_worst_acc_kernel<<<(N + 255) / 256, 256, 0, stream>>>(inter, tgt, A, N, C, acc_an, worst_n);
  ROBSEG_LAUNCH_CHECK();
  return 0;
}
