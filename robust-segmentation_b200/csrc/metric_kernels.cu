// Integer-exact metric counters: per-image confusion matrix (shared-memory tile, leader-
// aggregated atomics), intersection / target / prediction class counters (lane-private byte
// counters, no atomics in the pixel loop), and the SEA worst-case accuracy reduction.
//
// Replaces the 2*C-iteration masked reductions of compute_iou_acc (semseg/attacker.py:9-52),
// Metrics.update's bincount (semseg/metrics.py:27-33), eval_performance
// (tools/infer.py:86-116) and evalSEA (tools/worse_only.py:30-66,383-408).
// Bytes: 16 B per pixel (int64 pred + int64 label) read once; HBM-bound.
#include "common.cuh"

#include <cstdlib>

namespace robseg {

constexpr int kHistThreads = 256;

// One batch of a block: THREADS * 2 * kPairs consecutive pixels (kPairs 16-byte loads in flight
// per array and thread); pair h of thread tid is
// pixels h*2*THREADS + 2*tid, +1 (16-byte coalesced loads, a warp covers 512 contiguous bytes).
// Only the low words are kept: class ids and the ignore value fit in 32 bits.
template <int THREADS, int kPairs>
struct PxBatch {
  static constexpr int kN = 2 * kPairs;
  static constexpr int kPx = THREADS * 2 * kPairs;
  int t[2 * kPairs], q[2 * kPairs];

  // lp / pp point at the block's first pixel, n = pixels of the block, base = batch offset.
  __device__ __forceinline__ void load(const int64_t* __restrict__ lp,
                                       const int64_t* __restrict__ pp, int base, int n,
                                       bool vec_ok) {
    if (vec_ok && base + kPx <= n) {  // block-uniform: whole batch in range and 16-byte aligned
      longlong2 a[kPairs], b[kPairs];
#pragma unroll
      for (int h = 0; h < kPairs; ++h) {
        const int i = base + h * 2 * THREADS + 2 * (int)threadIdx.x;
        a[h] = __ldcs(reinterpret_cast<const longlong2*>(lp + i));
        b[h] = __ldcs(reinterpret_cast<const longlong2*>(pp + i));
      }
#pragma unroll
      for (int h = 0; h < kPairs; ++h) {
        t[2 * h] = (int)a[h].x, t[2 * h + 1] = (int)a[h].y;
        q[2 * h] = (int)b[h].x, q[2 * h + 1] = (int)b[h].y;
      }
    } else {
#pragma unroll
      for (int h = 0; h < kPairs; ++h) {
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int i = base + h * 2 * THREADS + 2 * (int)threadIdx.x + e;
          const bool in = i < n;
          t[2 * h + e] = in ? (int)__ldcs(lp + i) : -1;
          q[2 * h + e] = in ? (int)__ldcs(pp + i) : -1;
        }
      }
    }
  }
};

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

// Work split shared by both kernels: the n_img*HW pixels are one flat range cut into equal
// pieces of px_per_block (a multiple of the batch size), one per block, so every SM gets the same
// number of pixels whatever the image count.  A piece that crosses an image boundary is handled
// as one SEGMENT per image, each flushed to its own image's counters.
struct Segment {
  int64_t img, off;  // image, first pixel inside it
  int n;             // pixels
};
__device__ __forceinline__ Segment segment_at(int64_t f0, int64_t f1, int64_t HW) {
  Segment s;
  s.img = f0 / HW;
  s.off = f0 - s.img * HW;
  const int64_t left = HW - s.off;
  s.n = (int)(f1 - f0 < left ? f1 - f0 : left);
  return s;
}

// FULL: shared [C*C] confusion tile, inter/tgt/prd derived from it at flush time (one atomic
// per pixel).  Otherwise 3*C counters (inter, tgt, prd) for class counts too large for the
// atomic-free kernel below.  Batch i+1 is in flight while batch i is counted.
template <bool FULL, int kPairs>
__global__ void __launch_bounds__(kHistThreads)
    pixel_hist_kernel(const int64_t* __restrict__ pred, const int64_t* __restrict__ labels,
                      int n_lab_img, int64_t HW, int64_t total_px, int64_t px_per_block, int C,
                      int ignore_index, unsigned long long* hist, unsigned long long* hist_total,
                      unsigned long long* inter, unsigned long long* tgt,
                      unsigned long long* prd) {
  extern __shared__ int sh[];
  using Batch = PxBatch<kHistThreads, kPairs>;
  const int n_cnt = FULL ? C * C : 3 * C;
  int64_t f0 = (int64_t)blockIdx.x * px_per_block;
  const int64_t f1 = f0 + px_per_block < total_px ? f0 + px_per_block : total_px;

  auto count = [&](const Batch& b) {
#pragma unroll
    for (int e = 0; e < 2 * kPairs; ++e) {
      // one unsigned compare covers "negative or >= C"
      const int t = ((unsigned)b.t[e] < (unsigned)C && b.t[e] != ignore_index) ? b.t[e] : -1;
      const int q = ((unsigned)b.q[e] < (unsigned)C) ? b.q[e] : -1;
      if constexpr (FULL) {
        leader_add(sh, (t >= 0 && q >= 0) ? t * C + q : -1);
      } else {
        leader_add(sh + C, t);                        // tgt
        leader_add(sh + 2 * C, (t >= 0) ? q : -1);    // prd (pred := ignore where target is)
        leader_add(sh, (t >= 0 && t == q) ? t : -1);  // inter
      }
    }
  };

  while (f0 < f1) {  // block-uniform
    const Segment sg = segment_at(f0, f1, HW);
    const int n = sg.n;
    const int64_t img = sg.img;
    const int64_t* pp = pred + img * HW + sg.off;
    const int64_t* lp = labels + (img % n_lab_img) * HW + sg.off;
    const bool vec_ok =
        ((reinterpret_cast<uintptr_t>(pp) | reinterpret_cast<uintptr_t>(lp)) & 15) == 0;
    Batch A, B;
    A.load(lp, pp, 0, n, vec_ok);  // in flight while the tile is cleared
    for (int i = threadIdx.x; i < n_cnt; i += kHistThreads) sh[i] = 0;
    __syncthreads();
    // whole warps iterate together (uniform trip count) so the shuffles/ballots are defined
    for (int base = 0; base < n; base += 2 * Batch::kPx) {
      const bool more = base + Batch::kPx < n;
      if (more) B.load(lp, pp, base + Batch::kPx, n, vec_ok);
      count(A);
      if (base + 2 * Batch::kPx < n) A.load(lp, pp, base + 2 * Batch::kPx, n, vec_ok);
      if (more) count(B);
    }
    __syncthreads();
    if constexpr (FULL) {
      if (hist || hist_total) {
        for (int i = threadIdx.x; i < C * C; i += kHistThreads) {
          const int v = sh[i];
          if (v) {
            if (hist) atomicAdd(hist + img * C * C + i, (unsigned long long)v);
            if (hist_total) atomicAdd(hist_total + i, (unsigned long long)v);
          }
        }
      }
      if (inter || tgt || prd) {
        for (int c = threadIdx.x; c < C; c += kHistThreads) {
          int rs = 0, cs = 0;
          const int* row = sh + c * C;  // lanes stride C words: at most 2-way conflicts for even C
          const int* col = sh + c;      // lanes stride 1 word: conflict-free
#pragma unroll 4
          for (int k = 0; k < C; ++k) rs += row[k], cs += col[k * C];
          const int64_t o = img * C + c;
          if (inter && sh[c * C + c]) atomicAdd(inter + o, (unsigned long long)sh[c * C + c]);
          if (tgt && rs) atomicAdd(tgt + o, (unsigned long long)rs);
          if (prd && cs) atomicAdd(prd + o, (unsigned long long)cs);
        }
      }
    } else {
      for (int c = threadIdx.x; c < C; c += kHistThreads) {
        const int64_t o = img * C + c;
        if (inter && sh[c]) atomicAdd(inter + o, (unsigned long long)sh[c]);
        if (tgt && sh[C + c]) atomicAdd(tgt + o, (unsigned long long)sh[C + c]);
        if (prd && sh[2 * C + c]) atomicAdd(prd + o, (unsigned long long)sh[2 * C + c]);
      }
    }
    f0 += n;
    if (f0 < f1) __syncthreads();  // the tile is cleared again for the next image
  }
}

// ---- counters only (inter / tgt / prd), no atomics in the pixel loop ---------------------------
// Shared-memory atomics cost ~2 cycles per lane whatever the addresses, which caps an
// atomic-per-pixel histogram at ~1/3 of the HBM roofline.  Here every LANE owns private 8-bit
// counters: bin c of array a lives in byte (c&3) of word ((a*G + c/4)*32 + lane) of its warp's
// slab, so lane L only ever touches bank L -- plain conflict-free LDS/ADD/STS, no atomics, and
// the speed does not depend on the label distribution.  A lane adds at most one to a counter
// per pixel, so the bytes are folded into the block's int32 tile every <= 248 pixels per lane.
// Arrays: 0 = pixels with pred == label, 1 = label counts of the other pixels, 2 = their pred
// counts (one or two updates per pixel instead of three); tgt/prd are re-assembled at flush.
constexpr int kCntWarps = 4;
constexpr int kCntPairs = 4, kHistPairs = 4;  // batch depth: 8 measured the same (profiles/r01_hist.md)
constexpr int kCntThreads = 32 * kCntWarps;

__device__ __forceinline__ uint32_t lds_u8(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ void sts_u8(uint32_t addr, uint32_t v) {
  asm volatile("st.shared.u8 [%0], %1;" ::"r"(addr), "r"(v));
}

template <int kPairs>
__global__ void __launch_bounds__(kCntThreads)
    pixel_counts_kernel(const int64_t* __restrict__ pred, const int64_t* __restrict__ labels,
                        int n_lab_img, int64_t HW, int64_t total_px, int64_t px_per_block, int C,
                        int ignore_index, unsigned long long* inter, unsigned long long* tgt,
                        unsigned long long* prd) {
  extern __shared__ __align__(16) int sh[];
  using Batch = PxBatch<kCntThreads, kPairs>;
  constexpr int kBatchesPerFold = 240 / (2 * kPairs);  // <= 240 increments per lane between folds
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int G = (C + 3) >> 2;  // 4 bins per word
  int* tile = sh;              // [3*C] int32, block-wide
  uint32_t* slab = reinterpret_cast<uint32_t*>(sh + ((3 * C + 3) & ~3)) + (size_t)warp * 3 * G * 32;
  const uint32_t arr = (uint32_t)G * 128;  // bytes per array in the slab
  const uint32_t lane_base = smem_u32(slab) + 4 * lane;
  int64_t f0 = (int64_t)blockIdx.x * px_per_block;
  const int64_t f1 = f0 + px_per_block < total_px ? f0 + px_per_block : total_px;

  auto count = [&](const Batch& b) {
#pragma unroll
    for (int e = 0; e < 2 * kPairs; ++e) {
      const int t = b.t[e], q = b.q[e];
      const bool tv = (unsigned)t < (unsigned)C && t != ignore_index;
      const bool qv = (unsigned)q < (unsigned)C;
      const bool hit = tv && t == q;
      const uint32_t tt = tv ? (uint32_t)t : 0u, qq = qv ? (uint32_t)q : 0u;
      // invalid pixels add zero to bin 0: no divergence
      const uint32_t a1 = lane_base + (hit ? 0u : arr) + ((tt & ~3u) << 5) + (tt & 3u);
      const uint32_t a2 = lane_base + 2u * arr + ((qq & ~3u) << 5) + (qq & 3u);
      const uint32_t v1 = lds_u8(a1), v2 = lds_u8(a2);  // different arrays: never the same byte
      sts_u8(a1, v1 + (tv ? 1u : 0u));
      sts_u8(a2, v2 + ((tv && qv && !hit) ? 1u : 0u));
    }
  };
  // fold this warp's bytes into the block tile and clear them.  Lane L sums word row g = L, L+32,
  // ... over the 32 lane columns, starting at its own column (rotation keeps the banks distinct).
  auto fold = [&]() {
    __syncwarp();
    for (int g = lane; g < 3 * G; g += 32) {
      uint32_t lo = 0, hi = 0;  // 16-bit pairs: bins (0,2) and (1,3); 32 * 255 < 65536
#pragma unroll 8
      for (int j = 0; j < 32; ++j) {
        uint32_t* w = slab + g * 32 + ((j + lane) & 31);
        const uint32_t v = *w;
        *w = 0;
        lo += v & 0x00ff00ffu, hi += (v >> 8) & 0x00ff00ffu;
      }
      const int a = g / G, c = 4 * (g - a * G);
      int* dst = tile + a * C + c;
      if (lo & 0xffffu) atomicAdd(dst, (int)(lo & 0xffffu));
      if (c + 1 < C && (hi & 0xffffu)) atomicAdd(dst + 1, (int)(hi & 0xffffu));
      if (c + 2 < C && (lo >> 16)) atomicAdd(dst + 2, (int)(lo >> 16));
      if (c + 3 < C && (hi >> 16)) atomicAdd(dst + 3, (int)(hi >> 16));
    }
    __syncwarp();
  };

  for (int i = lane; i < 3 * G * 32; i += 32) slab[i] = 0;  // folds leave it zero afterwards
  while (f0 < f1) {  // block-uniform
    const Segment sg = segment_at(f0, f1, HW);
    const int n = sg.n;
    const int64_t* pp = pred + sg.img * HW + sg.off;
    const int64_t* lp = labels + (sg.img % n_lab_img) * HW + sg.off;
    const bool vec_ok =
        ((reinterpret_cast<uintptr_t>(pp) | reinterpret_cast<uintptr_t>(lp)) & 15) == 0;
    Batch A, B;
    A.load(lp, pp, 0, n, vec_ok);  // in flight while the counters are cleared
    for (int i = threadIdx.x; i < 3 * C; i += kCntThreads) tile[i] = 0;
    __syncthreads();
    int since_fold = 0;
    for (int base = 0; base < n; base += 2 * Batch::kPx) {
      const bool more = base + Batch::kPx < n;
      if (more) B.load(lp, pp, base + Batch::kPx, n, vec_ok);
      count(A);
      if (base + 2 * Batch::kPx < n) A.load(lp, pp, base + 2 * Batch::kPx, n, vec_ok);
      if (more) count(B);
      since_fold += 2;
      if (since_fold + 2 > kBatchesPerFold) fold(), since_fold = 0;
    }
    fold();
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += kCntThreads) {
      const int both = tile[c], t_only = tile[C + c], p_only = tile[2 * C + c];
      const int64_t o = sg.img * C + c;
      if (inter && both) atomicAdd(inter + o, (unsigned long long)both);
      if (tgt && both + t_only) atomicAdd(tgt + o, (unsigned long long)(both + t_only));
      if (prd && both + p_only) atomicAdd(prd + o, (unsigned long long)(both + p_only));
    }
    f0 += n;
    if (f0 < f1) __syncthreads();  // the tile is cleared again for the next image
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
  ROBSEG_REQUIRE(n_img > 0 && n_lab_img > 0 && HW > 0 && C > 0,
                 "bad shape n_img=%d HW=%lld C=%d", n_img, (long long)HW, C);
  ROBSEG_REQUIRE(hist || hist_total || inter || tgt || prd, "no output requested");
  const bool want_full = hist != nullptr || hist_total != nullptr;
  const size_t tile_bytes = (size_t)C * C * sizeof(int);
  const bool tile_fits = tile_bytes <= 200 * 1024;
  ROBSEG_REQUIRE((size_t)3 * C * sizeof(int) <= 200 * 1024, "C=%d too large", C);
  auto u = [](int64_t* q) { return reinterpret_cast<unsigned long long*>(q); };
  // one resident wave of `slots` equal pieces of the flat pixel range (a second, nearly empty
  // wave doubles the time); a piece is a multiple of the kernel's batch size
  const int64_t total_px = (int64_t)n_img * HW;
  auto flat_grid = [&](int slots, int batch_px, int64_t* per_block) {
    int64_t pb = (total_px + slots - 1) / slots;
    pb = ((pb + batch_px - 1) / batch_px) * batch_px;
    if (pb < 2 * (int64_t)batch_px) pb = 2 * (int64_t)batch_px;
    if (pb > ((int64_t)1 << 30)) pb = ((int64_t)1 << 30) / batch_px * batch_px;  // 32-bit offsets inside a piece
    *per_block = pb;
    return (unsigned)((total_px + pb - 1) / pb);
  };
  int64_t per_block = 0;
  constexpr int kCntPx = kCntThreads * 2 * kCntPairs, kHistPx = kHistThreads * 2 * kHistPairs;

  // counters: the atomic-free kernel while a warp's private slab stays <= 48 KB (C <= 512)
  const int G = (C + 3) / 4;
  const size_t cnt_smem = (size_t)((3 * C + 3) & ~3) * sizeof(int) + (size_t)kCntWarps * 3 * G * 128;
  const bool want_cnt = inter || tgt || prd;
  const bool cnt_private = want_cnt && (!want_full || !tile_fits) && cnt_smem <= 200 * 1024;
  if (cnt_private) {
    int per_sm = (int)((227 * 1024) / (cnt_smem + 1024));
    if (per_sm > 12) per_sm = 12;
    if (const char* e = getenv("ROBSEG_CNT_PER_SM")) per_sm = atoi(e) > 0 ? atoi(e) : per_sm;
    auto kern = pixel_counts_kernel<kCntPairs>;
    ROBSEG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)cnt_smem));
    const unsigned grid = flat_grid(per_sm * sm_count(), kCntPx, &per_block);
    kern<<<grid, kCntThreads, cnt_smem, stream>>>(pred, labels, n_lab_img, HW, total_px, per_block,
                                                  C, ignore_index, u(inter), u(tgt), u(prd));
    ROBSEG_LAUNCH_CHECK();
    if (!want_full) return 0;
  }
  if (want_full && tile_fits) {
    int per_sm = tile_bytes > 100 * 1024 ? 1 : (tile_bytes > 64 * 1024 ? 2 : 4);
    if (const char* e = getenv("ROBSEG_HIST_PER_SM")) per_sm = atoi(e) > 0 ? atoi(e) : per_sm;
    auto kern = pixel_hist_kernel<true, kHistPairs>;
    ROBSEG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)tile_bytes));
    const unsigned grid = flat_grid(per_sm * sm_count(), kHistPx, &per_block);
    kern<<<grid, kHistThreads, tile_bytes, stream>>>(pred, labels, n_lab_img, HW, total_px,
                                                     per_block, C, ignore_index, u(hist),
                                                     u(hist_total), u(inter), u(tgt), u(prd));
  } else {
    if (want_full) {
      // [C,C] tile does not fit: histogram through global atomics, grid (chunks, n_img)
      ROBSEG_REQUIRE(n_img <= 65535, "n_img=%d too large for C=%d", n_img, C);
      int chunks = (4 * sm_count()) / n_img;
      if (chunks < 1) chunks = 1;
      int64_t pb = (HW + chunks - 1) / chunks;
      pb = ((pb + kHistThreads - 1) / kHistThreads) * kHistThreads;
      dim3 grid((unsigned)((HW + pb - 1) / pb), n_img);
      pixel_hist_global_kernel<<<grid, kHistThreads, 0, stream>>>(
          pred, labels, n_lab_img, HW, pb, C, ignore_index, u(hist), u(hist_total));
      ROBSEG_LAUNCH_CHECK();
    }
    if (want_cnt && !cnt_private) {
      const unsigned grid = flat_grid(4 * sm_count(), kHistPx, &per_block);
      pixel_hist_kernel<false, kHistPairs><<<grid, kHistThreads, (size_t)3 * C * sizeof(int), stream>>>(
          pred, labels, n_lab_img, HW, total_px, per_block, C, ignore_index, nullptr, nullptr,
          u(inter), u(tgt), u(prd));
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
