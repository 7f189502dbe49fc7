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
constexpr int kPxPerBlock = 256 * 32;  // pixels one block consumes

// Lanes holding the same key elect a leader that adds the group size: one shared-memory
// atomic per distinct key per warp instead of up to 32 colliding ones.
__device__ __forceinline__ void warp_agg_add(int* counters, int key, bool active) {
  const unsigned mask = __ballot_sync(0xffffffffu, active);
  if (!active) return;
  const unsigned peers = __match_any_sync(mask, key);
  const int leader = __ffs(peers) - 1;
  if ((int)(threadIdx.x & 31) == leader) atomicAdd(counters + key, __popc(peers));
}

// FULL: shared [C*C] confusion tile.  Otherwise 3*C counters (inter, tgt, prd).
template <bool FULL>
__global__ void __launch_bounds__(kHistThreads)
    pixel_hist_kernel(const int64_t* __restrict__ pred, const int64_t* __restrict__ labels,
                      int n_lab_img, int64_t HW, int C, int ignore_index,
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
  const int64_t p0 = (int64_t)blockIdx.x * kPxPerBlock;
  int64_t p1 = p0 + kPxPerBlock;
  if (p1 > HW) p1 = HW;
  // whole warps iterate together so the ballot/match masks are well defined
  for (int64_t base = p0 + (threadIdx.x & ~31); base < p1; base += kHistThreads) {
    const int64_t i = base + (threadIdx.x & 31);
    int t = -1, q = -1;
    if (i < p1) {
      const int64_t tv = __ldcs(lp + i), qv = __ldcs(pp + i);
      t = (tv != ignore_index && tv >= 0 && tv < C) ? (int)tv : -1;
      q = (qv >= 0 && qv < C) ? (int)qv : -1;
    }
    if constexpr (FULL) {
      warp_agg_add(sh, t * C + q, t >= 0 && q >= 0);
    } else {
      warp_agg_add(sh + C, t, t >= 0);                  // tgt
      warp_agg_add(sh + 2 * C, q, t >= 0 && q >= 0);    // prd (pred := ignore where target is)
      warp_agg_add(sh, t, t >= 0 && t == q);            // inter
    }
  }
  __syncthreads();
  if constexpr (FULL) {
    for (int i = threadIdx.x; i < C * C; i += kHistThreads) {
      const int v = sh[i];
      if (v) {
        if (hist) atomicAdd(hist + (int64_t)img * C * C + i, (unsigned long long)v);
        if (hist_total) atomicAdd(hist_total + i, (unsigned long long)v);
      }
    }
    if (inter || tgt || prd) {
      for (int c = threadIdx.x; c < C; c += kHistThreads) {
        int rs = 0, cs = 0;
        for (int k = 0; k < C; ++k) rs += sh[c * C + k], cs += sh[k * C + c];
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
                             int n_lab_img, int64_t HW, int C, int ignore_index,
                             unsigned long long* hist, unsigned long long* hist_total) {
  const int img = blockIdx.y;
  const int64_t* pp = pred + (int64_t)img * HW;
  const int64_t* lp = labels + (int64_t)(img % n_lab_img) * HW;
  const int64_t p0 = (int64_t)blockIdx.x * kPxPerBlock;
  int64_t p1 = p0 + kPxPerBlock;
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
  const bool full = hist != nullptr || hist_total != nullptr;
  size_t smem = (full ? (size_t)C * C : (size_t)3 * C) * sizeof(int);
  ROBSEG_REQUIRE((size_t)3 * C * sizeof(int) <= 200 * 1024, "C=%d too large", C);
  dim3 grid((unsigned)((HW + kPxPerBlock - 1) / kPxPerBlock), n_img);
  auto u = [](int64_t* q) { return reinterpret_cast<unsigned long long*>(q); };
  if (full && smem > 200 * 1024) {
    // [C,C] tile does not fit: histogram through global atomics, counters through the 3C kernel
    pixel_hist_global_kernel<<<grid, kHistThreads, 0, stream>>>(pred, labels, n_lab_img, HW, C,
                                                                ignore_index, u(hist), u(hist_total));
    ROBSEG_LAUNCH_CHECK();
    if (inter || tgt || prd) {
      smem = (size_t)3 * C * sizeof(int);
      pixel_hist_kernel<false><<<grid, kHistThreads, smem, stream>>>(
          pred, labels, n_lab_img, HW, C, ignore_index, nullptr, nullptr, u(inter), u(tgt), u(prd));
    }
  } else if (full) {
    ROBSEG_CUDA(cudaFuncSetAttribute(pixel_hist_kernel<true>,
                                     cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    pixel_hist_kernel<true><<<grid, kHistThreads, smem, stream>>>(
        pred, labels, n_lab_img, HW, C, ignore_index, u(hist), u(hist_total), u(inter), u(tgt), u(prd));
  } else {
    pixel_hist_kernel<false><<<grid, kHistThreads, smem, stream>>>(
        pred, labels, n_lab_img, HW, C, ignore_index, nullptr, nullptr, u(inter), u(tgt), u(prd));
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
