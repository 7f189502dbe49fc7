// Fused per-pixel softmax + {CE, Mask-CE, Mask-CE-balanced, JS} loss + d(loss)/d(logits)
// + argmax + per-image partial sums, one read of the logits and one write of the gradient.
//
// Replaces the ATen chains of semseg/attacker.py:143-173,187-240,251-257 and the autograd
// of their per-image mean (formulas in section 10).
//
// Data layout.  Logits are NCHW: the softmax axis (C) has stride HW, pixels are contiguous.
// A *warp tile* is [C channels x 32*VEC pixels]; every lane owns VEC adjacent pixels and
// walks the channel axis, so each shared-memory row access is one conflict-free 128*VEC/…
// byte wavefront and every global store is a full coalesced line.
//
// Fast path (loss_tma_kernel): persistent CTAs, one per SM.  Warp 0 is a TMA producer: one
// elected lane streams warp tiles into a ring of shared-memory stages with 3-D tensor-map
// bulk copies (cp.async.bulk.tensor, SASS UTMALDG) over the [B][C][HW] view, box
// {32*VEC, C, 1}; completion is signalled on a per-stage "full" mbarrier.  Warps 1..W are
// consumers: each takes every W-th ring item, makes three passes over ITS OWN stage
// (max/argmax, sum-exp, gradient), writes the gradient straight from registers and releases
// the stage through the "empty" mbarrier.  No __syncthreads in the steady state.
//
// Generic path (loss_generic_kernel): any HW / alignment / C; each warp loads its tile with
// plain coalesced loads into a private stage and runs the same per-tile code.
//
// Determinism: per-tile partial sums are written to a workspace and reduced per image in a
// fixed order by loss_finalize_kernel (no float atomics), so repeated runs are bit-identical.
#include <cuda.h>
#include <cudaTypedefs.h>

#include <cstdlib>
#include <mutex>

#include "common.cuh"

namespace robseg {

struct LossParams {
  const void* logits;
  const int64_t* labels;
  const float* class_w;
  const float* grad_scale;
  const float* upstream;
  void* dlogits;
  float* loss_pix;
  int64_t* pred;
  float4* partials;
  unsigned long long* counts;  // [n_rep][B][3][C] inter / tgt / prd (zeroed by the launcher), or nullptr
  int64_t HW;
  int kind, ignore_index, B, C, n_rep;
  int tiles_per_img, num_tiles, n_slots, n_consumers;
  float inv_hw;
};

constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;

// ---- element access ------------------------------------------------------------------------
template <typename T, int VEC>
struct Vec;

template <int VEC>
struct Vec<float, VEC> {
  static __device__ __forceinline__ void lds(const float* p, float (&v)[VEC]) {
    if constexpr (VEC == 1) {
      v[0] = *p;
    } else if constexpr (VEC == 2) {
      float2 t = *reinterpret_cast<const float2*>(p);
      v[0] = t.x, v[1] = t.y;
    } else {
      static_assert(VEC == 4, "fp32 VEC in {1,2,4}");
      float4 t = *reinterpret_cast<const float4*>(p);
      v[0] = t.x, v[1] = t.y, v[2] = t.z, v[3] = t.w;
    }
  }
  static __device__ __forceinline__ void stg(float* p, const float (&v)[VEC]) {
    if constexpr (VEC == 1) {
      __stcs(p, v[0]);
    } else if constexpr (VEC == 2) {
      __stcs(reinterpret_cast<float2*>(p), make_float2(v[0], v[1]));
    } else {
      __stcs(reinterpret_cast<float4*>(p), make_float4(v[0], v[1], v[2], v[3]));
    }
  }
  static __device__ __forceinline__ float ld1(const float* p) { return *p; }
  static __device__ __forceinline__ void st1(float* p, float v) { *p = v; }
};

template <int VEC>
struct Vec<__nv_bfloat16, VEC> {
  using T = __nv_bfloat16;
  static __device__ __forceinline__ void unpack(uint32_t w, float& a, float& b) {
    a = __uint_as_float(w << 16);
    b = __uint_as_float(w & 0xffff0000u);
  }
  static __device__ __forceinline__ uint32_t pack(float a, float b) {
    __nv_bfloat162 t = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&t);
  }
  static __device__ __forceinline__ void lds(const T* p, float (&v)[VEC]) {
    if constexpr (VEC == 1) {
      v[0] = __bfloat162float(*p);
    } else if constexpr (VEC == 2) {
      unpack(*reinterpret_cast<const uint32_t*>(p), v[0], v[1]);
    } else if constexpr (VEC == 4) {
      uint2 t = *reinterpret_cast<const uint2*>(p);
      unpack(t.x, v[0], v[1]), unpack(t.y, v[2], v[3]);
    } else {
      static_assert(VEC == 8, "bf16 VEC in {1,2,4,8}");
      uint4 t = *reinterpret_cast<const uint4*>(p);
      unpack(t.x, v[0], v[1]), unpack(t.y, v[2], v[3]);
      unpack(t.z, v[4], v[5]), unpack(t.w, v[6], v[7]);
    }
  }
  static __device__ __forceinline__ void stg(T* p, const float (&v)[VEC]) {
    if constexpr (VEC == 1) {
      *p = __float2bfloat16_rn(v[0]);
    } else if constexpr (VEC == 2) {
      __stcs(reinterpret_cast<uint32_t*>(p), pack(v[0], v[1]));
    } else if constexpr (VEC == 4) {
      __stcs(reinterpret_cast<uint2*>(p), make_uint2(pack(v[0], v[1]), pack(v[2], v[3])));
    } else {
      __stcs(reinterpret_cast<uint4*>(p), make_uint4(pack(v[0], v[1]), pack(v[2], v[3]),
                                                      pack(v[4], v[5]), pack(v[6], v[7])));
    }
  }
  static __device__ __forceinline__ float ld1(const T* p) { return __bfloat162float(*p); }
  static __device__ __forceinline__ void st1(T* p, float v) { *p = __float2bfloat16_rn(v); }
};

// ---- one warp tile -------------------------------------------------------------------------
// tile: shared memory [C][32*VEC] of T, already filled (zeros beyond HW).  G warps share the
// tile, warp `half` walks channels [c_lo, c_hi) and the partial max / sum-exp / first-max
// index are combined through `xch` (shared) with a named barrier; G == 1: one warp, no exchange.
struct PairXch {  // per stage group, G == 2 only
  float* fmax;    // [2][ROW]
  float* fsum;    // [2][ROW]
  int* idx;       // [2][ROW]
  int bar_id;
};

__device__ __forceinline__ void pair_sync(int id) {
  asm volatile("bar.sync %0, 64;" ::"r"(id) : "memory");
}

__device__ __forceinline__ float fmax3(float a, float b, float c) {
  float r;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));  // SASS FMNMX3
  return r;
}

// STR (generic path, unaligned rows): the lane's VEC pixels are 32 apart in GLOBAL memory (lane,
// lane+32, ...) instead of adjacent, so every global access is a coalesced scalar one with no
// alignment requirement and a lane may be only partly inside the image.  The stage is still
// lane-major (pixel lane+32j at column lane*VEC+j: the 4-byte fill transposes), so shared memory
// is read with one vector load per row in both modes.
//
// bf16 with gradient (KEEP_E): pass 2 leaves e = 2^(z*log2e - mL), rounded to bf16, IN the stage (the warp
// owns it until it releases it), and pass 3 is one packed multiply per pair of logits (HMUL2.BF16) instead
// of unpack - FMA - ex2 - multiply - pack: the bf16 kernel was instruction-bound at 83 % of the roofline
// (19.7 instructions per logit, profiles/r01b_ncu_loss_bf16.md).  The gradient then carries three bf16
// roundings (e, coef/s, the product) instead of one: <= 0.6 % relative, inside the 1e-2 bf16 tolerance.
// (label, argmax) of a lane's VEC pixels, t = -1 where the label is ignored / out of range / beyond the image:
// what the per-image class counters (LossParams::counts) are updated from.  The TMA consumers count a tile while
// the NEXT one is being processed: the reductions are then long complete when the warp releases a stage
// (mbarrier.arrive has release semantics and would otherwise wait for their round trip to L2 once per tile --
// measured +19 % on the loss kernel with uniformly random labels when they were issued at the end of the tile).
template <int VEC>
struct TileCls {
  int t[VEC], q[VEC];
  int b;
};
template <int VEC>
__device__ __forceinline__ void count_tile(const LossParams& p, const TileCls<VEC>& k) {
  unsigned long long* cnt = p.counts + ((size_t)(blockIdx.x & (p.n_rep - 1)) * p.B + k.b) * 3 * p.C;
  count_pixels<VEC>(cnt, p.C, k.t, k.q);
}

// The labels of a warp tile (int64 -> int), ignore_index for pixels beyond the image.  Separate from
// process_tile so that the TMA consumers can issue these loads BEFORE they wait for their stage: the
// label latency then overlaps the wait instead of sitting between pass 1 and the first use.
template <int VEC, bool PARTIAL, bool STR>
__device__ __forceinline__ void tile_labels(const LossParams& p, int tile_idx, int lane, int (&y)[VEC]) {
  constexpr int ROW = 32 * VEC;
  constexpr int PS = STR ? 32 : 1;
  const int b = tile_idx / p.tiles_per_img;
  const int64_t px = (int64_t)(tile_idx - b * p.tiles_per_img) * ROW + (STR ? lane : lane * VEC);
  const int64_t pix = (int64_t)b * p.HW + px;
  const bool inb = PARTIAL ? (px < p.HW) : true;
  if constexpr (STR) {
#pragma unroll
    for (int j = 0; j < VEC; ++j)
      y[j] = (!PARTIAL || px + j * PS < p.HW) ? (int)__ldg(p.labels + pix + j * PS) : p.ignore_index;
  } else if (inb) {
    if constexpr (VEC == 1) {
      y[0] = (int)__ldg(p.labels + pix);
    } else {
      const longlong2* lp = reinterpret_cast<const longlong2*>(p.labels + pix);
#pragma unroll
      for (int j = 0; j < VEC / 2; ++j) {
        longlong2 t = __ldg(lp + j);
        y[2 * j] = (int)t.x, y[2 * j + 1] = (int)t.y;
      }
    }
  } else {
#pragma unroll
    for (int j = 0; j < VEC; ++j) y[j] = p.ignore_index;
  }
}

//
// OVF (generic path, fp32, with STR): the stage holds every channel row as the ALIGNED 16-byte chunks that
// cover the tile's pixels -- rows of ROW + 4 floats filled with 16-byte cp.async copies -- so pixel i of row c
// sits at column phase(c) + i, phase(c) = (ph0 + c * (HW & 3)) & 3.  Rows c, c+4, c+8, ... share their phase,
// so the unrolled loops (which start at multiples of 4) use four precomputed offsets; shared memory is read
// with conflict-free scalar loads (lane + 32 j).
//
// CREG > 0 (small class counts, C <= CREG; fp32, one warp per stage): the lane's C x VEC logits are read from the
// stage ONCE into registers and all three passes run on registers with fully unrolled channel loops (padding
// channels hold -inf: they lose every max, add exp(-inf) = 0 and never equal the maximum).  At C = 21 the
// shared-memory version issues ~38 instructions per logit -- per-pass row reads, loop and phase arithmetic, tails
// of the unrolled trips -- and is issue-bound at 72 % (473^2) / 85 % (472^2) of the roofline.
template <typename T, int VEC, int G, bool PARTIAL, bool STR = false, bool OVF = false, int CREG = 0>
__device__ __forceinline__ void process_tile(const LossParams& p, T* tile, int tile_idx, int lane, int half,
                                             const PairXch& xch, const int (&y)[VEC], TileCls<VEC>* cls = nullptr,
                                             int ph0 = 0) {
  static_assert(!STR || G == 1, "strided pixels: one warp per stage");
  static_assert(!OVF || (STR && sizeof(T) == 4), "over-fetched rows: fp32, strided pixel ownership");
  static_assert(CREG == 0 || (G == 1 && sizeof(T) == 4), "register-resident logits: fp32, one warp per stage");
  constexpr int ROW = 32 * VEC;
  constexpr int RS = OVF ? ROW + 4 : ROW;  // row stride of the stage in elements
  constexpr int PS = STR ? 32 : 1;  // distance between the lane's pixels
  constexpr int NACC = (VEC >= 4) ? 1 : 4 / VEC;  // independent accumulator sets per pixel
  constexpr int UNR = 8;
  constexpr int kNone = 0x7fffffff;
  const int C = p.C;
  const int c_lo = (G == 1) ? 0 : half * ((C + 1) >> 1);
  const int c_hi = (G == 1) ? C : (half == 0 ? ((C + 1) >> 1) : C);
  const int b = tile_idx / p.tiles_per_img;
  const int lane_px = STR ? lane : lane * VEC;
  const int64_t px = (int64_t)(tile_idx - b * p.tiles_per_img) * ROW + lane_px;
  // adjacent pixels: HW % VEC == 0, a lane is all in or all out; strided: per pixel
  const bool inb = PARTIAL ? (px < p.HW) : true;
  bool inj[VEC];
#pragma unroll
  for (int j = 0; j < VEC; ++j) inj[j] = (STR && PARTIAL) ? (px + j * PS < p.HW) : inb;
  const int64_t pix = (int64_t)b * p.HW + px;
  // shared memory is lane-major in both dense modes (the strided fill transposes): one vector load per row
  T* col = OVF ? tile + lane : tile + lane * VEC;
  int phs[4] = {0, 0, 0, 0};
  const int dph = OVF ? (int)(p.HW & 3) : 0;
  if constexpr (OVF) {
#pragma unroll
    for (int q = 0; q < 4; ++q) phs[q] = (ph0 + q * dph) & 3;
  }
  // row c of the stage -> the lane's VEC logits.  `slot` = c & 3 where the caller knows it at compile time
  // (unrolled trips that start at a multiple of 4), -1 otherwise.
  auto ldrow = [&](int c, int slot, float (&v)[VEC]) {
    if constexpr (OVF) {
      const int ph = slot >= 0 ? phs[slot & 3] : ((ph0 + c * dph) & 3);
      const float* rp = reinterpret_cast<const float*>(col) + c * RS + ph;
#pragma unroll
      for (int j = 0; j < VEC; ++j) v[j] = rp[32 * j];
    } else {
      Vec<T, VEC>::lds(col + c * ROW, v);
    }
  };
  const bool argmax_only = p.kind == ROBSEG_LOSS_ARGMAX;
  constexpr bool KEEP_E = sizeof(T) == 2 && VEC >= 2 && !STR;
  const bool keep_e = KEEP_E && p.dlogits != nullptr;
  using V = Vec<T, VEC>;
  auto stgv = [&](T* q, const float (&v)[VEC]) {
    if constexpr (STR) {
#pragma unroll
      for (int j = 0; j < VEC; ++j)
        if (inj[j]) {
          if constexpr (sizeof(T) == 4) __stcs(reinterpret_cast<float*>(q) + j * PS, v[j]);
          else V::st1(q + j * PS, v[j]);
        }
    } else {
      if (inb) V::stg(q, v);
    }
  };

  // ---- pass 1: channel max.  ARGMAX-only launches also track the index here (one pass);
  // loss launches find the first maximal channel by equality during pass 2.
  float mx[VEC];
  int amx[VEC];
  float zr[CREG > 0 ? CREG : 1][VEC];  // CREG: the lane's logits, -inf beyond C
  if constexpr (CREG > 0) {
#pragma unroll
    for (int c = 0; c < CREG; ++c) {
      if (c < C) {
        ldrow(c, c, zr[c]);
      } else {
#pragma unroll
        for (int j = 0; j < VEC; ++j) zr[c][j] = -INFINITY;
      }
    }
#pragma unroll
    for (int j = 0; j < VEC; ++j) mx[j] = zr[0][j], amx[j] = argmax_only ? 0 : kNone;
    if (argmax_only) {
#pragma unroll
      for (int c = 1; c < CREG; ++c)
#pragma unroll
        for (int j = 0; j < VEC; ++j) {
          const bool g = zr[c][j] > mx[j];  // ascending, strict: ties keep the lowest channel
          mx[j] = g ? zr[c][j] : mx[j];
          amx[j] = g ? c : amx[j];
        }
    } else {
#pragma unroll
      for (int c = 1; c + 1 < CREG; c += 2)
#pragma unroll
        for (int j = 0; j < VEC; ++j) mx[j] = fmax3(mx[j], zr[c][j], zr[c + 1][j]);
      if constexpr (CREG % 2 == 0) {
#pragma unroll
        for (int j = 0; j < VEC; ++j) mx[j] = fmaxf(mx[j], zr[CREG - 1][j]);
      }
    }
  } else if (argmax_only) {
    float m[NACC][VEC];
    int am[NACC][VEC];
#pragma unroll
    for (int a = 0; a < NACC; ++a)
#pragma unroll
      for (int j = 0; j < VEC; ++j) m[a][j] = -INFINITY, am[a][j] = c_lo;
    int c = c_lo;
#pragma unroll 1
    for (; c + UNR <= c_hi; c += UNR) {
#pragma unroll
      for (int u = 0; u < UNR; ++u) {
        float v[VEC];
        ldrow(c + u, u, v);
#pragma unroll
        for (int j = 0; j < VEC; ++j) {
          const bool g = v[j] > m[u % NACC][j];
          m[u % NACC][j] = g ? v[j] : m[u % NACC][j];
          am[u % NACC][j] = g ? (c + u) : am[u % NACC][j];
        }
      }
    }
#pragma unroll 1
    for (; c < c_hi; ++c) {
      float v[VEC];
      ldrow(c, -1, v);
#pragma unroll
      for (int j = 0; j < VEC; ++j) {
        const bool g = v[j] > m[0][j] || (v[j] == m[0][j] && c < am[0][j]);
        m[0][j] = g ? v[j] : m[0][j];
        am[0][j] = g ? c : am[0][j];
      }
    }
#pragma unroll
    for (int j = 0; j < VEC; ++j) {
      mx[j] = m[0][j], amx[j] = am[0][j];
#pragma unroll
      for (int a = 1; a < NACC; ++a) {
        const bool g = m[a][j] > mx[j] || (m[a][j] == mx[j] && am[a][j] < amx[j]);
        mx[j] = g ? m[a][j] : mx[j];
        amx[j] = g ? am[a][j] : amx[j];
      }
    }
  } else if constexpr (sizeof(T) == 2 && VEC >= 2) {
    // bf16: the maximum of packed pairs is exact, so the whole pass runs on the raw words
    // (HMNMX2.BF16, no unpack): 1 instruction per logit instead of 2
    constexpr int NW = VEC / 2;
    __nv_bfloat162 m2[2][NW];
    const __nv_bfloat162 ninf = __floats2bfloat162_rn(-INFINITY, -INFINITY);
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
      for (int k = 0; k < NW; ++k) m2[a][k] = ninf;
    auto ldw = [&](const T* q, __nv_bfloat162 (&wv)[NW]) {
      if constexpr (NW == 1) {
        wv[0] = *reinterpret_cast<const __nv_bfloat162*>(q);
      } else if constexpr (NW == 2) {
        const uint2 t = *reinterpret_cast<const uint2*>(q);
        wv[0] = *reinterpret_cast<const __nv_bfloat162*>(&t.x);
        wv[1] = *reinterpret_cast<const __nv_bfloat162*>(&t.y);
      } else {
        const uint4 t = *reinterpret_cast<const uint4*>(q);
        wv[0] = *reinterpret_cast<const __nv_bfloat162*>(&t.x);
        wv[1] = *reinterpret_cast<const __nv_bfloat162*>(&t.y);
        wv[2] = *reinterpret_cast<const __nv_bfloat162*>(&t.z);
        wv[3] = *reinterpret_cast<const __nv_bfloat162*>(&t.w);
      }
    };
    int c = c_lo;
#pragma unroll 1
    for (; c + UNR <= c_hi; c += UNR) {
#pragma unroll
      for (int u = 0; u < UNR; ++u) {
        __nv_bfloat162 wv[NW];
        ldw(col + (c + u) * ROW, wv);
#pragma unroll
        for (int k = 0; k < NW; ++k) m2[u & 1][k] = __hmax2(m2[u & 1][k], wv[k]);
      }
    }
#pragma unroll 1
    for (; c < c_hi; ++c) {
      __nv_bfloat162 wv[NW];
      ldw(col + c * ROW, wv);
#pragma unroll
      for (int k = 0; k < NW; ++k) m2[0][k] = __hmax2(m2[0][k], wv[k]);
    }
#pragma unroll
    for (int k = 0; k < NW; ++k) {
      const __nv_bfloat162 mm = __hmax2(m2[0][k], m2[1][k]);
      mx[2 * k] = __low2float(mm), mx[2 * k + 1] = __high2float(mm);
      amx[2 * k] = amx[2 * k + 1] = kNone;
    }
  } else {
    float m[NACC][VEC];
#pragma unroll
    for (int a = 0; a < NACC; ++a)
#pragma unroll
      for (int j = 0; j < VEC; ++j) m[a][j] = -INFINITY;
    int c = c_lo;
#pragma unroll 1
    for (; c + UNR <= c_hi; c += UNR) {
#pragma unroll
      for (int u = 0; u < UNR; u += 2) {
        float v0[VEC], v1[VEC];
        ldrow(c + u, u, v0);
        ldrow(c + u + 1, u + 1, v1);
#pragma unroll
        for (int j = 0; j < VEC; ++j)
          m[(u / 2) % NACC][j] = fmax3(m[(u / 2) % NACC][j], v0[j], v1[j]);
      }
    }
    if (c + 4 <= c_hi) {  // a half trip keeps the scalar tail below 4 rows (C = 21: 16 + 4 + 1)
#pragma unroll
      for (int u = 0; u < 4; u += 2) {
        float v0[VEC], v1[VEC];
        ldrow(c + u, u, v0);
        ldrow(c + u + 1, u + 1, v1);
#pragma unroll
        for (int j = 0; j < VEC; ++j)
          m[(u / 2) % NACC][j] = fmax3(m[(u / 2) % NACC][j], v0[j], v1[j]);
      }
      c += 4;
    }
#pragma unroll 1
    for (; c < c_hi; ++c) {
      float v[VEC];
      ldrow(c, -1, v);
#pragma unroll
      for (int j = 0; j < VEC; ++j) m[0][j] = fmaxf(m[0][j], v[j]);
    }
#pragma unroll
    for (int j = 0; j < VEC; ++j) {
      mx[j] = m[0][j];
#pragma unroll
      for (int a = 1; a < NACC; ++a) mx[j] = fmaxf(mx[j], m[a][j]);
      amx[j] = kNone;
    }
  }
  if constexpr (G == 2) {
#pragma unroll
    for (int j = 0; j < VEC; ++j) {
      xch.fmax[half * ROW + lane * VEC + j] = mx[j];
      xch.idx[half * ROW + lane * VEC + j] = amx[j];
    }
    pair_sync(xch.bar_id);
#pragma unroll
    for (int j = 0; j < VEC; ++j) {
      const float om = xch.fmax[(half ^ 1) * ROW + lane * VEC + j];
      const int oi = xch.idx[(half ^ 1) * ROW + lane * VEC + j];
      // ties between the halves go to the lower half (lower channel indices)
      const bool take = half == 0 ? (om > mx[j]) : (om >= mx[j]);
      amx[j] = take ? oi : amx[j];
      mx[j] = take ? om : mx[j];
    }
    pair_sync(xch.bar_id);  // both warps have read: the slots may be rewritten (pass 2 / next tile)
  }

  float loss_sum = 0.f, ce_sum = 0.f;
  int n_correct = 0, n_valid = 0;
  bool valid[VEC], hit[VEC];
  int ys[VEC];
  if (argmax_only) {
#pragma unroll
    for (int j = 0; j < VEC; ++j) {
      valid[j] = inj[j] && (y[j] != p.ignore_index) && (y[j] >= 0) && (y[j] < C);
      hit[j] = valid[j] && (amx[j] == y[j]);
      n_correct += hit[j], n_valid += valid[j];
    }
  } else {
    // ---- pass 2: sum of exp(z - max) + first maximal channel --------------------------------
    // exp(z-m) = 2^(z*log2e - mL) with mL = fl(m*log2e); the rounding residual of that
    // product is common to all channels (cancels in softmax) and is removed from the
    // log-sum-exp exactly through `resid`.  Channels are walked downwards so that the
    // predicated index write that lands last is the lowest maximal channel.
    float mL[VEC], resid[VEC];
#pragma unroll
    for (int j = 0; j < VEC; ++j) {
      mL[j] = mx[j] * kLog2e;
      resid[j] = fmaf(mx[j], kLog2e, -mL[j]);
    }
    // the label logit: read before pass 2 overwrites the stage (KEEP_E); otherwise after it, when the
    // labels have certainly arrived
    float zyv[VEC];
    if (keep_e) {
#pragma unroll
      for (int j = 0; j < VEC; ++j) {
        const bool v = inj[j] && (y[j] != p.ignore_index) && (y[j] >= 0) && (y[j] < C);
        zyv[j] = V::ld1(col + (v ? y[j] : 0) * ROW + j);
      }
      if constexpr (G == 2) pair_sync(xch.bar_id);  // the partner reads label logits from MY channel half too
    }
    float s[NACC][VEC];
#pragma unroll
    for (int a = 0; a < NACC; ++a)
#pragma unroll
      for (int j = 0; j < VEC; ++j) s[a][j] = 0.f;
    // e values of one row back into the stage as packed bf16 (KEEP_E only)
    auto put_e = [&](T* q, const float (&e)[VEC]) {
      if constexpr (KEEP_E) {
        using VB = Vec<__nv_bfloat16, VEC>;
        if constexpr (VEC == 2) {
          *reinterpret_cast<uint32_t*>(q) = VB::pack(e[0], e[1]);
        } else if constexpr (VEC == 4) {
          *reinterpret_cast<uint2*>(q) = make_uint2(VB::pack(e[0], e[1]), VB::pack(e[2], e[3]));
        } else {
          *reinterpret_cast<uint4*>(q) = make_uint4(VB::pack(e[0], e[1]), VB::pack(e[2], e[3]),
                                                    VB::pack(e[4], e[5]), VB::pack(e[6], e[7]));
        }
      }
    };
    int c = c_hi;
    if constexpr (CREG > 0) {
#pragma unroll
      for (int cc = CREG - 1; cc >= 0; --cc)
#pragma unroll
        for (int j = 0; j < VEC; ++j) {
          s[cc % NACC][j] += ex2_approx(fmaf(zr[cc][j], kLog2e, -mL[j]));
          if (zr[cc][j] == mx[j]) amx[j] = cc;  // walking downwards: the last hit is the lowest channel
        }
      c = c_lo;  // the shared-memory loops below have nothing left to do
    }
    // one half trip of 4 rows (rows c-1 .. c-4), walking downwards like the full trips
    auto half_trip = [&]() {
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        float v[VEC], e[VEC];
        ldrow(c - 1 - u, OVF ? (3 - u) : -1, v);
#pragma unroll
        for (int j = 0; j < VEC; ++j) {
          e[j] = ex2_approx(fmaf(v[j], kLog2e, -mL[j]));
          s[u % NACC][j] += e[j];
          if (v[j] == mx[j]) amx[j] = c - 1 - u;
        }
        if (keep_e) put_e(col + (c - 1 - u) * ROW, e);
      }
      c -= 4;
    };
    if constexpr (OVF) {
      // the rows above the last multiple of 4 first (still walking downwards), then a half trip down to a
      // multiple of UNR, so that every unrolled trip starts at a multiple of 4 and knows its rows' phases at
      // compile time
#pragma unroll 1
      for (; c > c_lo && (c & 3) != 0; --c) {
        float v[VEC];
        ldrow(c - 1, -1, v);
#pragma unroll
        for (int j = 0; j < VEC; ++j) {
          s[0][j] += ex2_approx(fmaf(v[j], kLog2e, -mL[j]));
          if (v[j] == mx[j]) amx[j] = c - 1;
        }
      }
      if ((c & 4) != 0 && c - 4 >= c_lo) half_trip();
    }
#pragma unroll 1
    for (; c - UNR >= c_lo; c -= UNR) {
#pragma unroll
      for (int u = 0; u < UNR; ++u) {
        float v[VEC], e[VEC];
        ldrow(c - 1 - u, OVF ? (UNR - 1 - u) : -1, v);
#pragma unroll
        for (int j = 0; j < VEC; ++j) {
          e[j] = ex2_approx(fmaf(v[j], kLog2e, -mL[j]));
          s[u % NACC][j] += e[j];
          if (v[j] == mx[j]) amx[j] = c - 1 - u;
        }
        if (keep_e) put_e(col + (c - 1 - u) * ROW, e);
      }
    }
    if constexpr (!OVF) {
      if (c - 4 >= c_lo) half_trip();
    }
#pragma unroll 1
    for (; c > c_lo; --c) {
      float v[VEC], e[VEC];
      ldrow(c - 1, -1, v);
#pragma unroll
      for (int j = 0; j < VEC; ++j) {
        e[j] = ex2_approx(fmaf(v[j], kLog2e, -mL[j]));
        s[0][j] += e[j];
        if (v[j] == mx[j]) amx[j] = c - 1;
      }
      if (keep_e) put_e(col + (c - 1) * ROW, e);
    }
    float st[VEC];
#pragma unroll
    for (int j = 0; j < VEC; ++j) {
      st[j] = s[0][j];
#pragma unroll
      for (int a = 1; a < NACC; ++a) st[j] += s[a][j];
    }
    if constexpr (G == 2) {
#pragma unroll
      for (int j = 0; j < VEC; ++j) {
        xch.fsum[half * ROW + lane * VEC + j] = st[j];
        xch.idx[half * ROW + lane * VEC + j] = amx[j];
      }
      pair_sync(xch.bar_id);
#pragma unroll
      for (int j = 0; j < VEC; ++j) {
        const float os = xch.fsum[(half ^ 1) * ROW + lane * VEC + j];
        const int oi = xch.idx[(half ^ 1) * ROW + lane * VEC + j];
        st[j] = half == 0 ? st[j] + os : os + st[j];  // same order in both warps
        amx[j] = min(amx[j], oi);
      }
      pair_sync(xch.bar_id);  // reads done before the partner's next-tile writes
    }

    // ---- per-pixel loss terms ---------------------------------------------------------------
    float kfac[VEC], sub[VEC], ey[VEC], loss[VEC];
    const float g_img = p.grad_scale ? __ldg(p.grad_scale + b) : p.inv_hw;
    bool any_grad = false;
#pragma unroll
    for (int j = 0; j < VEC; ++j) {
      valid[j] = inj[j] && (y[j] != p.ignore_index) && (y[j] >= 0) && (y[j] < C);
      hit[j] = valid[j] && (amx[j] == y[j]);
      ys[j] = valid[j] ? y[j] : 0;
      n_correct += hit[j], n_valid += valid[j];
      float zy;
      if constexpr (OVF) zy = reinterpret_cast<const float*>(col)[ys[j] * RS + ((ph0 + ys[j] * dph) & 3) + 32 * j];
      else zy = keep_e ? zyv[j] : V::ld1(col + ys[j] * ROW + j);
      const float ln_s = logf(st[j]) - resid[j] * kLn2;  // lse - m
      const float logp = (zy - mx[j]) - ln_s;            // log softmax_y  (<= 0)
      const float ce = valid[j] ? -logp : 0.f;
      float l, coef;
      if (p.kind == ROBSEG_LOSS_CE) {
        l = ce, coef = valid[j] ? 1.f : 0.f;
      } else if (p.kind == ROBSEG_LOSS_MASK_CE) {
        l = hit[j] ? ce : 0.f, coef = hit[j] ? 1.f : 0.f;
      } else if (p.kind == ROBSEG_LOSS_MASK_CE_BAL) {
        const float w = (p.class_w != nullptr) ? __ldg(p.class_w + ys[j]) : 1.f;
        l = hit[j] ? w * ce : 0.f, coef = hit[j] ? w : 0.f;
      } else {  // JS(softmax || one-hot), SURVEY section 10
        const float py = expf(logp);
        const float l1p = log1pf(py);
        l = valid[j] ? 0.5f * (2.f * kLn2 + py * logp - (1.f + py) * l1p) : 0.f;
        coef = valid[j] ? -0.5f * py * (logp - l1p) : 0.f;
      }
      loss[j] = l;
      loss_sum += l;
      ce_sum += ce;
      float gs = g_img;
      if (p.upstream != nullptr && inj[j]) gs *= __ldg(p.upstream + pix + j * PS);
      const float cg = coef * gs;
      kfac[j] = cg / st[j];
      sub[j] = cg;
      ey[j] = ex2_approx(fmaf(zy, kLog2e, -mL[j]));
      any_grad |= (cg != 0.f);
    }
    if (p.loss_pix != nullptr && inb && half == 0) {
      if constexpr (STR) {
#pragma unroll
        for (int j = 0; j < VEC; ++j)
          if (inj[j]) p.loss_pix[pix + j * PS] = loss[j];
      } else if constexpr (VEC == 1) {
        p.loss_pix[pix] = loss[0];
      } else if constexpr (VEC == 2) {
        *reinterpret_cast<float2*>(p.loss_pix + pix) = make_float2(loss[0], loss[1]);
      } else {
#pragma unroll
        for (int j = 0; j < VEC / 4; ++j)
          reinterpret_cast<float4*>(p.loss_pix + pix)[j] =
              make_float4(loss[4 * j], loss[4 * j + 1], loss[4 * j + 2], loss[4 * j + 3]);
      }
    }

    // ---- pass 3: gradient, written straight from registers ---------------------------------
    if (p.dlogits != nullptr) {
      T* gp = reinterpret_cast<T*>(p.dlogits) + ((int64_t)b * C + c_lo) * p.HW + px;
      const int64_t hw = p.HW;
      const bool warp_any = __any_sync(0xffffffffu, any_grad);
      if (!warp_any) {
        // every pixel of the tile is masked out (wrongly classified / ignored): zeros
        float zero[VEC];
#pragma unroll
        for (int j = 0; j < VEC; ++j) zero[j] = 0.f;
        if (inb) {
#pragma unroll 4
          for (c = c_lo; c < c_hi; ++c, gp += hw) stgv(gp, zero);
        }
      } else if (KEEP_E) {
        if constexpr (KEEP_E) {
          // the stage holds e (bf16): gradient = e * (coef / s), one packed multiply per pair
          constexpr int NW = VEC / 2;
          __nv_bfloat162 k2[NW];
#pragma unroll
          for (int k = 0; k < NW; ++k) k2[k] = __floats2bfloat162_rn(kfac[2 * k], kfac[2 * k + 1]);
          auto mul_store = [&](const T* q, T* g) {
            if constexpr (NW == 1) {
              const __nv_bfloat162 r = __hmul2(*reinterpret_cast<const __nv_bfloat162*>(q), k2[0]);
              if (inb) __stcs(reinterpret_cast<uint32_t*>(g), *reinterpret_cast<const uint32_t*>(&r));
            } else if constexpr (NW == 2) {
              const uint2 t = *reinterpret_cast<const uint2*>(q);
              const __nv_bfloat162 r0 = __hmul2(*reinterpret_cast<const __nv_bfloat162*>(&t.x), k2[0]);
              const __nv_bfloat162 r1 = __hmul2(*reinterpret_cast<const __nv_bfloat162*>(&t.y), k2[1]);
              if (inb)
                __stcs(reinterpret_cast<uint2*>(g), make_uint2(*reinterpret_cast<const uint32_t*>(&r0),
                                                               *reinterpret_cast<const uint32_t*>(&r1)));
            } else {
              const uint4 t = *reinterpret_cast<const uint4*>(q);
              const __nv_bfloat162 r0 = __hmul2(*reinterpret_cast<const __nv_bfloat162*>(&t.x), k2[0]);
              const __nv_bfloat162 r1 = __hmul2(*reinterpret_cast<const __nv_bfloat162*>(&t.y), k2[1]);
              const __nv_bfloat162 r2 = __hmul2(*reinterpret_cast<const __nv_bfloat162*>(&t.z), k2[2]);
              const __nv_bfloat162 r3 = __hmul2(*reinterpret_cast<const __nv_bfloat162*>(&t.w), k2[3]);
              if (inb)
                __stcs(reinterpret_cast<uint4*>(g),
                       make_uint4(*reinterpret_cast<const uint32_t*>(&r0), *reinterpret_cast<const uint32_t*>(&r1),
                                  *reinterpret_cast<const uint32_t*>(&r2), *reinterpret_cast<const uint32_t*>(&r3)));
            }
          };
#pragma unroll 8
          for (c = c_lo; c < c_hi; ++c, gp += hw) mul_store(col + c * ROW, gp);
          T* gy = reinterpret_cast<T*>(p.dlogits) + (int64_t)b * C * p.HW + px;
#pragma unroll
          for (int j = 0; j < VEC; ++j)
            if (sub[j] != 0.f && ys[j] >= c_lo && ys[j] < c_hi)
              V::st1(gy + (int64_t)ys[j] * p.HW + j, fmaf(ey[j], kfac[j], -sub[j]));
        }
      } else {
        c = c_lo;
        if constexpr (CREG > 0) {
#pragma unroll
          for (int cc = 0; cc < CREG; ++cc)
            if (cc < C) {
              float g[VEC];
#pragma unroll
              for (int j = 0; j < VEC; ++j) g[j] = ex2_approx(fmaf(zr[cc][j], kLog2e, -mL[j])) * kfac[j];
              stgv(gp, g);
              gp += hw;
            }
          c = c_hi;  // nothing left for the shared-memory loops
        }
#pragma unroll 1
        for (; c + UNR <= c_hi; c += UNR) {
          float v[UNR][VEC];
#pragma unroll
          for (int u = 0; u < UNR; ++u) ldrow(c + u, u, v[u]);
#pragma unroll
          for (int u = 0; u < UNR; ++u) {
            float g[VEC];
#pragma unroll
            for (int j = 0; j < VEC; ++j)
              g[j] = ex2_approx(fmaf(v[u][j], kLog2e, -mL[j])) * kfac[j];
            stgv(gp, g);
            gp += hw;
          }
        }
        if (c + 4 <= c_hi) {
          float v[4][VEC];
#pragma unroll
          for (int u = 0; u < 4; ++u) ldrow(c + u, u, v[u]);
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            float g[VEC];
#pragma unroll
            for (int j = 0; j < VEC; ++j)
              g[j] = ex2_approx(fmaf(v[u][j], kLog2e, -mL[j])) * kfac[j];
            stgv(gp, g);
            gp += hw;
          }
          c += 4;
        }
#pragma unroll 1
        for (; c < c_hi; ++c, gp += hw) {
          float v[VEC], g[VEC];
          ldrow(c, -1, v);
#pragma unroll
          for (int j = 0; j < VEC; ++j) g[j] = ex2_approx(fmaf(v[j], kLog2e, -mL[j])) * kfac[j];
          stgv(gp, g);
        }
        // the label channel: coef*(p_y - 1).  Same thread, same address, program order.
        T* gy = reinterpret_cast<T*>(p.dlogits) + (int64_t)b * C * p.HW + px;
#pragma unroll
        for (int j = 0; j < VEC; ++j)
          if (sub[j] != 0.f && ys[j] >= c_lo && ys[j] < c_hi)
            V::st1(gy + (int64_t)ys[j] * p.HW + j * PS, fmaf(ey[j], kfac[j], -sub[j]));
      }
    }
  }

  if (half == 0) {
    if (cls != nullptr) {  // warp-uniform: the caller counts (now or one tile later)
      cls->b = b;
#pragma unroll
      for (int j = 0; j < VEC; ++j) cls->t[j] = valid[j] ? y[j] : -1, cls->q[j] = amx[j];
    }
    if (p.pred != nullptr && inb) {
      if constexpr (STR) {
#pragma unroll
        for (int j = 0; j < VEC; ++j)
          if (inj[j]) p.pred[pix + j * PS] = amx[j];
      } else if constexpr (VEC == 1) {
        p.pred[pix] = amx[0];
      } else {
        longlong2* pp = reinterpret_cast<longlong2*>(p.pred + pix);
#pragma unroll
        for (int j = 0; j < VEC / 2; ++j) pp[j] = make_longlong2(amx[2 * j], amx[2 * j + 1]);
      }
    }
    // ---- per-tile partials (fixed shuffle order -> deterministic) --------------------------
    loss_sum = warp_sum(loss_sum);
    ce_sum = warp_sum(ce_sum);
    n_correct = warp_sum(n_correct);
    n_valid = warp_sum(n_valid);
    if (lane == 0)
      p.partials[tile_idx] =
          make_float4(loss_sum, ce_sum, __int_as_float(n_correct), __int_as_float(n_valid));
  }
}

// ---- fast path: persistent TMA ring --------------------------------------------------------
// Block = 1 producer warp + W*G consumer warps.  Consumer group w (G warps) owns K private
// stages; ring item `it` goes to group it % W, slot (it / W) % K.  A stage is only ever
// consumed by one group, so its mbarrier phases are observed in order (a parity wait is
// ambiguous for a waiter that is two phases away).
template <typename T, int VEC, int G, int CREG = 0>
__global__ void __launch_bounds__(G == 2 ? 1024 : 512, 1)
    loss_tma_kernel(const __grid_constant__ CUtensorMap tmap, const LossParams p) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  // keep the pointer derived from the __shared__ symbol so loads compile to LDS
  uint8_t* smem = smem_raw + ((128u - (smem_u32(smem_raw) & 127u)) & 127u);
  constexpr int ROW = 32 * VEC;
  const int W = p.n_consumers, K = p.n_slots, S = W * K;
  const uint32_t stage_bytes = (uint32_t)p.C * ROW * sizeof(T);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + (size_t)S * stage_bytes);
  uint64_t* empty = full + S;
  float* xbase = reinterpret_cast<float*>(empty + S);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmap);
    for (int s = 0; s < S; ++s) mbar_init(full + s, 1), mbar_init(empty + s, G);
    mbar_fence_init();
  }
  __syncthreads();

  if (warp == 0) {
    if (lane == 0) {  // ===== TMA producer =====
      int w = 0, k = 0;
      uint32_t round = 0;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
        const int s = w * K + k;
        mbar_wait(empty + s, (round & 1u) ^ 1u);
        mbar_arrive_expect_tx(full + s, stage_bytes);
        const int b = tile / p.tiles_per_img;
        const int px0 = (tile - b * p.tiles_per_img) * ROW;
        tma_load_3d(smem + (size_t)s * stage_bytes, &tmap, px0, 0, b, full + s);
        if (++w == W) {
          w = 0;
          if (++k == K) k = 0, ++round;
        }
      }
    }
  } else {  // ===== consumers =====
    const int w = (warp - 1) / G, half = (warp - 1) % G;
    PairXch xch{};
    if constexpr (G == 2) {
      xch.fmax = xbase + (size_t)w * 6 * ROW;
      xch.fsum = xch.fmax + 2 * ROW;
      xch.idx = reinterpret_cast<int*>(xch.fsum + 2 * ROW);
      xch.bar_id = 1 + w;
    }
    int k = 0;
    uint32_t round = 0;
    const bool has_partial = (p.HW % ROW) != 0;
    const bool counting = p.counts != nullptr && half == 0;
    TileCls<VEC> prev, now;
    bool pending = false;
    for (int tile = blockIdx.x + w * gridDim.x; tile < p.num_tiles; tile += W * gridDim.x) {
      const int s = w * K + k;
      const bool partial = has_partial && (tile % p.tiles_per_img) == p.tiles_per_img - 1;
      int y[VEC];  // label loads in flight while the stage arrives
      if (partial) tile_labels<VEC, true, false>(p, tile, lane, y);
      else tile_labels<VEC, false, false>(p, tile, lane, y);
      if (pending) count_tile<VEC>(p, prev);  // the previous tile's class counters (see TileCls)
      mbar_wait(full + s, round & 1u);
      T* st = reinterpret_cast<T*>(smem + (size_t)s * stage_bytes);
      TileCls<VEC>* out = counting ? &now : nullptr;
      if (partial)
        process_tile<T, VEC, G, true, false, false, CREG>(p, st, tile, lane, half, xch, y, out);
      else
        process_tile<T, VEC, G, false, false, false, CREG>(p, st, tile, lane, half, xch, y, out);
      if (counting) prev = now;
      pending = counting;
      if constexpr (sizeof(T) == 2)  // bf16 writes e into the stage: order those generic-proxy stores before the
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // TMA (async proxy) refill of the same bytes
      __syncwarp();
      if (lane == 0) mbar_arrive(empty + s);
      if (++k == K) k = 0, ++round;
    }
    if (pending) count_tile<VEC>(p, prev);
  }
}

// ---- generic path: any shape / alignment ---------------------------------------------------
// Shapes a tensor map cannot describe (HW*sizeof(T) not a multiple of 16 B -- e.g. the
// reference's 473x473 PASCAL-VOC crops -- or C > 256).  Each warp owns K stages of
// [C][32] elements and fills them with 4-byte cp.async copies (SASS LDGSTS: no register
// staging, a whole tile in flight per warp), double-buffered when shared memory allows, then
// runs the same per-tile code.
__device__ __forceinline__ void cp_async4(void* smem_dst, const void* gsrc, bool valid) {
  const int n = valid ? 4 : 0;  // src-size 0: zero-fill
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(smem_u32(smem_dst)), "l"(gsrc),
               "r"(n)
               : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

__global__ void __launch_bounds__(256) loss_generic_f32_kernel(const LossParams p, const int K) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, W = blockDim.x >> 5;
  const int stage_elems = p.C * 32;
  float* bufs = reinterpret_cast<float*>(smem_raw) + (size_t)warp * K * stage_elems;
  const float* logits = reinterpret_cast<const float*>(p.logits);
  auto issue = [&](int tile, float* dst) {
    const int b = tile / p.tiles_per_img;
    const int64_t px = (int64_t)(tile - b * p.tiles_per_img) * 32 + lane;
    const bool inb = px < p.HW;
    const float* src = logits + (int64_t)b * p.C * p.HW + (inb ? px : 0);
    uint32_t d = smem_u32(dst + lane);
    const int n = inb ? 4 : 0;  // src-size 0: zero-fill
    const int64_t hw = p.HW;
#pragma unroll 4
    for (int c = 0; c < p.C; ++c, src += hw, d += 128)
      asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(d), "l"(src), "r"(n) : "memory");
    cp_async_commit();
  };
  const int stride = gridDim.x * W;
  int tile = blockIdx.x * W + warp;
  int k = 0;
  TileCls<1> prev, now;
  bool pending = false;
  if (K == 2 && tile < p.num_tiles) issue(tile, bufs);
  for (; tile < p.num_tiles; tile += stride) {
    float* cur = bufs + (size_t)k * stage_elems;
    if (K == 2) {
      const int next = tile + stride;
      if (next < p.num_tiles) {
        issue(next, bufs + (size_t)(k ^ 1) * stage_elems);
        cp_async_wait<1>();
      } else {
        cp_async_wait<0>();
      }
      k ^= 1;
    } else {
      issue(tile, cur);
      cp_async_wait<0>();
    }
    __syncwarp();
    const int tin = tile % p.tiles_per_img;
    int y[1];
    if (pending) count_tile<1>(p, prev);  // one tile late: see TileCls
    TileCls<1>* out = p.counts != nullptr ? &now : nullptr;
    if ((int64_t)(tin + 1) * 32 <= p.HW) {
      tile_labels<1, false, false>(p, tile, lane, y);
      process_tile<float, 1, 1, false>(p, cur, tile, lane, 0, PairXch{}, y, out);
    } else {
      tile_labels<1, true, false>(p, tile, lane, y);
      process_tile<float, 1, 1, true>(p, cur, tile, lane, 0, PairXch{}, y, out);
    }
    if (out != nullptr) prev = now, pending = true;
    __syncwarp();
  }
  if (pending) count_tile<1>(p, prev);
}

// Wider tiles for the same case: [C][32*VEC] stages filled with 4-byte copies (lane L copies
// elements L, L+32, ... of every row, whatever the row alignment, into ITS columns L*VEC ..) and
// consumed with the lane's VEC pixels 32 apart, so labels, argmax map and gradient are coalesced
// scalar accesses while shared memory is read with vector loads.  The
// per-pixel work (label, log, loss terms) and the per-tile reductions are amortised over VEC
// times more pixels per lane than in the one-pixel kernel above (23 -> ~15 instructions per logit
// at C = 21).
template <int VEC>
__global__ void __launch_bounds__(256) loss_generic_strided_kernel(const LossParams p, const int K) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  constexpr int ROW = 32 * VEC;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, W = blockDim.x >> 5;
  const int stage_elems = p.C * ROW;
  float* bufs = reinterpret_cast<float*>(smem_raw) + (size_t)warp * K * stage_elems;
  const float* logits = reinterpret_cast<const float*>(p.logits);
  auto issue = [&](int tile, float* dst) {
    const int b = tile / p.tiles_per_img;
    const int64_t px = (int64_t)(tile - b * p.tiles_per_img) * ROW + lane;
    const float* src = logits + (int64_t)b * p.C * p.HW + px;
    int n[VEC];
#pragma unroll
    for (int k = 0; k < VEC; ++k) n[k] = px + 32 * k < p.HW ? 4 : 0;  // src-size 0: zero-fill
    if (n[0] == 0) src = logits;  // keep the (unused) address valid
    uint32_t d = smem_u32(dst + lane * VEC);  // transposed: pixel lane+32k -> column lane*VEC+k
    const int64_t hw = p.HW;
#pragma unroll 2
    for (int c = 0; c < p.C; ++c, src += hw, d += 4 * ROW) {
#pragma unroll
      for (int k = 0; k < VEC; ++k)
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(d + 4 * k),
                     "l"(n[k] ? src + 32 * k : src), "r"(n[k])
                     : "memory");
    }
    cp_async_commit();
  };
  const int stride = gridDim.x * W;
  int tile = blockIdx.x * W + warp;
  int k = 0;
  TileCls<VEC> prev, now;
  bool pending = false;
  if (K == 2 && tile < p.num_tiles) issue(tile, bufs);
  for (; tile < p.num_tiles; tile += stride) {
    float* cur = bufs + (size_t)k * stage_elems;
    if (K == 2) {
      const int next = tile + stride;
      if (next < p.num_tiles) {
        issue(next, bufs + (size_t)(k ^ 1) * stage_elems);
        cp_async_wait<1>();
      } else {
        cp_async_wait<0>();
      }
      k ^= 1;
    } else {
      issue(tile, cur);
      cp_async_wait<0>();
    }
    __syncwarp();
    const int tin = tile % p.tiles_per_img;
    int y[VEC];
    if (pending) count_tile<VEC>(p, prev);  // one tile late: see TileCls
    TileCls<VEC>* out = p.counts != nullptr ? &now : nullptr;
    if ((int64_t)(tin + 1) * ROW <= p.HW) {
      tile_labels<VEC, false, true>(p, tile, lane, y);
      process_tile<float, VEC, 1, false, true>(p, cur, tile, lane, 0, PairXch{}, y, out);
    } else {
      tile_labels<VEC, true, true>(p, tile, lane, y);
      process_tile<float, VEC, 1, true, true>(p, cur, tile, lane, 0, PairXch{}, y, out);
    }
    if (out != nullptr) prev = now, pending = true;
    __syncwarp();
  }
  if (pending) count_tile<VEC>(p, prev);
}

// The same case with 16-byte copies (one 4-byte cp.async per logit kept the copy pipe busy
// ~70 % of the time at the roofline rate -- LDGSTS costs ~8 cycles per warp instruction whatever its width).
// A channel row of a tile starts at an arbitrary 4-byte phase of global memory (row stride HW*4 bytes is not a
// multiple of 16), but the ALIGNED chunks that cover it can be copied 16 bytes at a time: ROW/4 + 1 chunks per row
// into a stage row of ROW + 4 floats, a flat (row, chunk) item list strided over the lanes so every copy
// instruction is full.  The consumer side (process_tile<..., STR, OVF>) reads pixel i of row c at column
// phase(c) + i.  Chunks beyond the tensor are zero-filled (src-size < 16); up to 3 floats of over-fetch per row
// come from the neighbouring rows' bytes of the same tensor, i.e. from L2.
template <int VEC, int CREG = 0>
__global__ void __launch_bounds__(256) loss_generic_ovf_kernel(const LossParams p, const int K, const int64_t n_total) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  constexpr int ROW = 32 * VEC, RS = ROW + 4, NCH = ROW / 4 + 1;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, W = blockDim.x >> 5;
  const int stage_elems = p.C * RS;
  float* bufs = reinterpret_cast<float*>(smem_raw) + (size_t)warp * K * stage_elems;
  const float* logits = reinterpret_cast<const float*>(p.logits);
  const int n_items = p.C * NCH;
  auto row0 = [&](int tile) {  // element offset of the tile's first pixel in channel 0 of its image
    const int b = tile / p.tiles_per_img;
    return (int64_t)b * p.C * p.HW + (int64_t)(tile - b * p.tiles_per_img) * ROW;
  };
  // Offsets inside an image fit 32 bits (the launcher checks C*HW < 2^31 - ROW): per item the address is
  // the tile's aligned base pointer + 4 * (((r0 + c*HW) & ~3) + 4k) with r0 = (image base + first pixel) & 3
  // carried into the 32-bit part; only a tile that can reach the end of the tensor takes the bounds-checked
  // copy size.
  auto issue = [&](int tile, float* dst) {
    const int b = tile / p.tiles_per_img;
    const int p0 = (tile - b * p.tiles_per_img) * ROW;
    const int64_t img = (int64_t)b * p.C * p.HW;
    const float* base = logits + (img & ~(int64_t)3);  // 16-byte aligned
    const int r0 = (int)(img & 3) + p0;
    const int hw = (int)p.HW;
    const uint32_t d0 = smem_u32(dst);
    const bool near_end = (img & ~(int64_t)3) + r0 + (int64_t)(p.C - 1) * hw + ROW + 8 > n_total;
    if (!near_end) {
#pragma unroll 4
      for (int i = lane; i < n_items; i += 32) {
        const int c = i / NCH, k = i - c * NCH;
        const int off = ((r0 + c * hw) & ~3) + 4 * k;
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d0 + 4u * (uint32_t)(c * RS + 4 * k)),
                     "l"(base + off)
                     : "memory");
      }
    } else {
      const int64_t left0 = n_total - (img & ~(int64_t)3);  // elements from `base` to the end of the tensor
      for (int i = lane; i < n_items; i += 32) {
        const int c = i / NCH, k = i - c * NCH;
        const int off = ((r0 + c * hw) & ~3) + 4 * k;
        const int64_t rem = left0 - off;
        const int nbytes = rem >= 4 ? 16 : (rem > 0 ? 4 * (int)rem : 0);  // src-size < 16: zero-fill the rest
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d0 + 4u * (uint32_t)(c * RS + 4 * k)),
                     "l"(base + (rem > 0 ? off : 0)), "r"(nbytes)
                     : "memory");
      }
    }
    cp_async_commit();
  };
  const int stride = gridDim.x * W;
  int tile = blockIdx.x * W + warp;
  int k = 0;
  TileCls<VEC> prev, now;
  bool pending = false;
  if (K == 2 && tile < p.num_tiles) issue(tile, bufs);
  for (; tile < p.num_tiles; tile += stride) {
    float* cur = bufs + (size_t)k * stage_elems;
    const int tin = tile % p.tiles_per_img;
    const bool full = (int64_t)(tin + 1) * ROW <= p.HW;
    int y[VEC];  // label loads in flight while the copies land
    if (full) tile_labels<VEC, false, true>(p, tile, lane, y);
    else tile_labels<VEC, true, true>(p, tile, lane, y);
    if (K == 2) {
      const int next = tile + stride;
      if (next < p.num_tiles) {
        issue(next, bufs + (size_t)(k ^ 1) * stage_elems);
        cp_async_wait<1>();
      } else {
        cp_async_wait<0>();
      }
      k ^= 1;
    } else {
      issue(tile, cur);
      cp_async_wait<0>();
    }
    __syncwarp();
    const int ph0 = (int)(row0(tile) & 3);
    if (pending) count_tile<VEC>(p, prev);  // one tile late: see TileCls
    TileCls<VEC>* out = p.counts != nullptr ? &now : nullptr;
    if (full) process_tile<float, VEC, 1, false, true, true, CREG>(p, cur, tile, lane, 0, PairXch{}, y, out, ph0);
    else process_tile<float, VEC, 1, true, true, true, CREG>(p, cur, tile, lane, 0, PairXch{}, y, out, ph0);
    if (out != nullptr) prev = now, pending = true;
    __syncwarp();
  }
  if (pending) count_tile<VEC>(p, prev);
}

template <typename T>
__global__ void __launch_bounds__(256) loss_generic_kernel(const LossParams p) {
  extern __shared__ uint8_t smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, W = blockDim.x >> 5;
  T* stage = reinterpret_cast<T*>(smem_raw) + (size_t)warp * p.C * 32;
  const T* logits = reinterpret_cast<const T*>(p.logits);
  TileCls<1> prev, now;
  bool pending = false;
  for (int tile = blockIdx.x * W + warp; tile < p.num_tiles; tile += gridDim.x * W) {
    const int b = tile / p.tiles_per_img;
    const int64_t px = (int64_t)(tile - b * p.tiles_per_img) * 32 + lane;
    const bool inb = px < p.HW;
    const T* src = logits + (int64_t)b * p.C * p.HW + px;
#pragma unroll 8
    for (int c = 0; c < p.C; ++c)
      stage[c * 32 + lane] = inb ? src[(int64_t)c * p.HW] : T(0.f);
    __syncwarp();
    int y[1];
    tile_labels<1, true, false>(p, tile, lane, y);
    if (pending) count_tile<1>(p, prev);  // one tile late: see TileCls
    TileCls<1>* out = p.counts != nullptr ? &now : nullptr;
    process_tile<T, 1, 1, true>(p, stage, tile, lane, 0, PairXch{}, y, out);
    if (out != nullptr) prev = now, pending = true;
    __syncwarp();
  }
  if (pending) count_tile<1>(p, prev);
}

// ---- fixed-order per-image reduction of the tile partials ----------------------------------
__global__ void __launch_bounds__(256)
    loss_finalize_kernel(const float4* __restrict__ partials, int tiles_per_img,
                         const float* __restrict__ grad_scale, double inv_hw, float* loss_img,
                         float* track_img, int32_t* correct_img, int32_t* valid_img) {
  __shared__ double sh_l[256], sh_c[256];
  __shared__ int sh_k[256], sh_v[256];
  const int b = blockIdx.x, t = threadIdx.x;
  double l = 0.0, ce = 0.0;
  int k = 0, v = 0;
  const float4* src = partials + (size_t)b * tiles_per_img;
  // 8 independent loads in flight per thread (the serial loop paid one L2 round trip per partial: 16 of them at
  // 512^2); the additions keep the ascending-i order, so the sums are bit-identical to the serial loop's
  constexpr int U = 8;
  for (int i0 = t; i0 < tiles_per_img; i0 += 256 * U) {
    float4 q[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int i = i0 + 256 * u;
      q[u] = i < tiles_per_img ? __ldcg(src + i) : make_float4(0.f, 0.f, __int_as_float(0), __int_as_float(0));
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (i0 + 256 * u < tiles_per_img)
        l += (double)q[u].x, ce += (double)q[u].y, k += __float_as_int(q[u].z), v += __float_as_int(q[u].w);
    }
  }
  sh_l[t] = l, sh_c[t] = ce, sh_k[t] = k, sh_v[t] = v;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (t < o) sh_l[t] += sh_l[t + o], sh_c[t] += sh_c[t + o], sh_k[t] += sh_k[t + o], sh_v[t] += sh_v[t + o];
    __syncthreads();
  }
  if (t == 0) {
    const double g = grad_scale ? (double)grad_scale[b] : inv_hw;
    if (loss_img) loss_img[b] = (float)(g * sh_l[0]);
    if (track_img) track_img[b] = (float)(sh_c[0] * inv_hw);
    if (correct_img) correct_img[b] = sh_k[0];
    if (valid_img) valid_img[b] = sh_v[0];
  }
}

// counts[b][i] = sum over the R replicas (see count_replicas).  Grid (B, ceil(3C / 32)), 8 warps: warp w adds replicas
// w, w+8, ... for 32 consecutive counters (coalesced 256-byte reads, R/8 independent loads per lane), then the warps'
// partial sums meet in shared memory.  (One thread per counter walking all R copies was latency-bound: ~20 us at R = 64.)
__global__ void __launch_bounds__(256)
    counts_fold_kernel(const unsigned long long* __restrict__ rep, int R, int B, int n, long long* __restrict__ counts) {
  __shared__ unsigned long long part[8][32];
  const int b = blockIdx.x, lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int i = blockIdx.y * 32 + lane;
  unsigned long long s = 0;
  if (i < n) {
#pragma unroll 4
    for (int r = w; r < R; r += 8) s += __ldcg(rep + ((size_t)r * B + b) * n + i);
  }
  part[w][lane] = s;
  __syncthreads();
  if (w == 0 && i < n) {
#pragma unroll
    for (int k = 1; k < 8; ++k) s += part[k][lane];
    counts[(size_t)b * n + i] = (long long)s;
  }
}
int launch_counts_fold(const unsigned long long* replicas, int R, int B, int C, int64_t* counts, cudaStream_t stream) {
  counts_fold_kernel<<<dim3(B, (3 * C + 31) / 32), 256, 0, stream>>>(replicas, R, B, 3 * C,
                                                                   reinterpret_cast<long long*>(counts));
  ROBSEG_LAUNCH_CHECK();
  return 0;
}
// Zeroing the replicas with a kernel of our own: cudaMemsetAsync may be served by a copy engine, and the hand-over
// between engines costs far more than the 0.2-2 MB memset itself (ROBSEG_COUNTS_MEMSET=1 restores it for comparison).
__global__ void __launch_bounds__(256) counts_zero_kernel(uint4* __restrict__ p, size_t n16) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n16; i += (size_t)gridDim.x * blockDim.x)
    p[i] = make_uint4(0u, 0u, 0u, 0u);
}
int launch_counts_zero(unsigned long long* replicas, size_t bytes, cudaStream_t stream) {
  if (getenv("ROBSEG_COUNTS_MEMSET")) {
    ROBSEG_CUDA(cudaMemsetAsync(replicas, 0, bytes, stream));
    return 0;
  }
  const size_t n16 = bytes / 16;
  int grid = (int)((n16 + 255) / 256);
  if (grid > sm_count() * 4) grid = sm_count() * 4;
  if (grid < 1) grid = 1;
  counts_zero_kernel<<<grid, 256, 0, stream>>>(reinterpret_cast<uint4*>(replicas), n16);
  ROBSEG_LAUNCH_CHECK();
  return 0;
}
static size_t count_replica_bytes(int B, int C) { return (size_t)count_replicas(C) * B * 3 * C * sizeof(int64_t); }

// ---- host side -----------------------------------------------------------------------------
static PFN_cuTensorMapEncodeTiled_v12000 get_encode() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) ==
            cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(f);
  });
  return fn;
}

// Small class counts (VOC 21, Cityscapes 19): logits kept in registers (process_tile CREG), 2 pixels per lane.
// ROBSEG_LOSS_CREG=0 restores the shared-memory passes.
constexpr int kRegChannels = 24;
static bool creg_enabled() {
  const char* e = getenv("ROBSEG_LOSS_CREG");
  return !(e && atoi(e) == 0);
}

constexpr size_t kSmemBudget = 227 * 1024 - 1024;  // ring + barriers + alignment slack
constexpr size_t kMaxDynSmem = 227 * 1024;         // opt-in dynamic shared memory per CTA on sm_100

static int pick_vec(int C, int esize, bool with_grad) {
  // widest per-lane vector whose stage ([C][32*VEC] elements) stays <= 40 KB, so at least 5
  // consumer warps fit; per-lane bytes are capped at 16.  Wider rows matter: measured on B200
  // at C=150 the 128 B-row layout tops out near 5.1 TB/s, 256 B rows reach 5.6 TB/s
  // (profiles/r01_loss_sweep.md).
  const int max_vec = 16 / esize;
  int vec = max_vec;
  // bf16 pays an unpack per element, so it keeps the smaller (<= 24 KB, more warps) stages.
  // Loss-only / argmax launches have no store stream and are warp-count bound: smaller stages.
  const size_t cap = (esize == 2 || !with_grad) ? 24 * 1024 : 40 * 1024;
  while (vec > (esize == 2 ? 2 : 1) && (size_t)C * 32 * vec * esize > cap) vec >>= 1;
  if (const char* e = getenv("ROBSEG_LOSS_VEC")) {
    const int v = atoi(e);
    if ((v == 1 || v == 2 || v == 4 || v == 8) && v <= max_vec && v >= (esize == 2 ? 2 : 1)) vec = v;
  }
  return vec;
}

template <typename T, int VEC, int G, int CREG = 0>
static int launch_tma(const LossParams& p0, cudaStream_t stream) {
  LossParams p = p0;
  constexpr int ROW = 32 * VEC;
  p.tiles_per_img = (int)((p.HW + ROW - 1) / ROW);
  p.num_tiles = p.B * p.tiles_per_img;
  const size_t stage = (size_t)p.C * ROW * sizeof(T);
  const size_t xch = G == 2 ? (size_t)6 * ROW * sizeof(float) : 0;  // per consumer group
  // exact accounting (stages + two mbarriers per stage + exchange areas + 128 B of alignment slack):
  // at C = 151 (the reference's ADE20K class count) six 38.7 KB stages fit with 288 bytes to spare --
  // a coarser budget dropped the kernel to five consumer warps there (76 % vs 92 % of the roofline).
  const size_t budget = kMaxDynSmem - 128;
  // W consumer groups (G warps each) x K private stages per group.  Warps first (the kernel
  // is latency-bound per warp), then deeper prefetch with whatever shared memory is left.
  int W = (int)(budget / (stage + 16 + xch));
  if (W > 15) W = 15;
  if (const char* e = getenv("ROBSEG_LOSS_WARPS")) W = atoi(e) > 0 && atoi(e) <= W ? atoi(e) : W;
  ROBSEG_REQUIRE(W >= 1, "C=%d: one stage does not fit in shared memory", p.C);
  int K = (int)((budget - W * xch) / (W * (stage + 16)));
  if (K > 4) K = 4;
  if (const char* e = getenv("ROBSEG_LOSS_SLOTS")) K = atoi(e) > 0 && atoi(e) <= K ? atoi(e) : K;
  const int S = W * K;
  p.n_slots = K, p.n_consumers = W;
  const size_t smem = S * stage + 2 * S * sizeof(uint64_t) + W * xch + 128;

  CUtensorMap tmap;
  const cuuint64_t dims[3] = {(cuuint64_t)p.HW, (cuuint64_t)p.C, (cuuint64_t)p.B};
  const cuuint64_t strides[2] = {(cuuint64_t)p.HW * sizeof(T), (cuuint64_t)p.HW * p.C * sizeof(T)};
  const cuuint32_t box[3] = {(cuuint32_t)ROW, (cuuint32_t)p.C, 1u};
  const cuuint32_t estr[3] = {1u, 1u, 1u};
  auto encode = get_encode();
  ROBSEG_REQUIRE(encode != nullptr, "cuTensorMapEncodeTiled entry point unavailable");
  CUresult r = encode(&tmap,
                      sizeof(T) == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16,
                      3, const_cast<void*>(p.logits), dims, strides, box, estr,
                      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                      CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  ROBSEG_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed (CUresult %d)", (int)r);

  auto kern = loss_tma_kernel<T, VEC, G, CREG>;
  ROBSEG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int grid = sm_count();
  const int max_useful = (p.num_tiles + W - 1) / W;
  if (grid > max_useful) grid = max_useful;
  if (grid < 1) grid = 1;
  kern<<<grid, 32 * (W * G + 1), smem, stream>>>(tmap, p);
  ROBSEG_LAUNCH_CHECK();
  return 0;
}

// G = 2 (two warps split the channel axis of one stage) when a stage is so large that fewer
// than 15 one-warp groups fit; only for per-lane vectors <= 8 bytes (register budget of a
// 1024-thread block).
template <typename T, int VEC>
static int launch_tma_pick(const LossParams& p, cudaStream_t stream) {
  const size_t stage = (size_t)p.C * 32 * VEC * sizeof(T);
  // Measured: splitting the channel axis between two warps does not beat one warp per stage
  // once the loads compile to LDS (the kernel is then bound by the memory system, not by
  // warp-level parallelism); kept selectable for experiments and covered by the tests.
  (void)stage;
  bool pair = false;
  if (const char* e = getenv("ROBSEG_LOSS_G")) pair = atoi(e) == 2;
  if constexpr (VEC * sizeof(T) <= 4) {
    if (pair) return launch_tma<T, VEC, 2>(p, stream);
  }
  return launch_tma<T, VEC, 1>(p, stream);
}

template <typename T>
static int launch_generic(const LossParams& p0, cudaStream_t stream, int* tiles_per_img) {
  LossParams p = p0;
  p.tiles_per_img = *tiles_per_img = (int)((p.HW + 31) / 32);
  p.num_tiles = p.B * p.tiles_per_img;
  const size_t stage = (size_t)p.C * 32 * sizeof(T);
  ROBSEG_REQUIRE(stage <= kSmemBudget, "C=%d too large for one shared-memory stage", p.C);
  if constexpr (sizeof(T) == 4) {
    // 16-byte over-fetch path (default when the logits pointer is 16-byte aligned; ROBSEG_LOSS_GENERIC_OVF=0
    // restores the 4-byte-copy kernels below): 2 pixels per lane while >= 16 warps per SM keep two stages each,
    // else 1 pixel per lane with as many warps as fit (two stages when >= 16 of them do).
    const char* ovf_env = getenv("ROBSEG_LOSS_GENERIC_OVF");
    if (reinterpret_cast<uintptr_t>(p.logits) % 16 == 0 && !(ovf_env && atoi(ovf_env) == 0)) {
      auto st_bytes = [&](int v) { return (size_t)p.C * (32 * v + 4) * sizeof(float); };
      int vec = (size_t)16 * 2 * st_bytes(2) <= kSmemBudget ? 2 : 1;
      if (ovf_env && (atoi(ovf_env) == 1 || atoi(ovf_env) == 2 || atoi(ovf_env) == 4)) vec = atoi(ovf_env);
      const size_t st = st_bytes(vec);
      // at least 4 warps with one stage each and 32-bit offsets inside an image, else the old kernels
      if (st * 4 <= kSmemBudget && (int64_t)p.C * p.HW < ((int64_t)1 << 31) - 1024) {
        int K = (size_t)16 * 2 * st <= kSmemBudget ? 2 : 1;
        if (vec == 4 && (size_t)8 * 2 * st <= kSmemBudget) K = 2;
        int warps_sm = (int)(kSmemBudget / (K * st));
        if (warps_sm > 32) warps_sm = 32;
        const int W = warps_sm >= 8 ? 8 : warps_sm;
        const int ctas_sm = warps_sm / W;
        const size_t smem = (size_t)W * K * st;
        auto kern = vec == 4 ? loss_generic_ovf_kernel<4> : vec == 2 ? loss_generic_ovf_kernel<2> : loss_generic_ovf_kernel<1>;
        if (vec == 2 && p.C <= kRegChannels && creg_enabled()) kern = loss_generic_ovf_kernel<2, kRegChannels>;
        ROBSEG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        p.tiles_per_img = *tiles_per_img = (int)((p.HW + 32 * vec - 1) / (32 * vec));
        p.num_tiles = p.B * p.tiles_per_img;
        int grid = sm_count() * ctas_sm;
        const int max_useful = (p.num_tiles + W - 1) / W;
        if (grid > max_useful) grid = max_useful;
        if (grid < 1) grid = 1;
        kern<<<grid, 32 * W, smem, stream>>>(p, K, (int64_t)p.B * p.C * p.HW);
        ROBSEG_LAUNCH_CHECK();
        return 0;
      }
    }
    // wide strided tiles while >= 16 warps per SM keep two stages each (C <= 13 at 4 pixels per
    // lane, <= 27 at 2; measured at 24x21x473x473: 0.27 ms at 2, 0.29-0.35 ms at 4 with 10 warps,
    // 0.33 ms at 1); otherwise one pixel per lane.  ROBSEG_LOSS_GENERIC_VEC forces a width that
    // fits 8 warps (tests, experiments).
    int vec = 4, min_warps = 16;
    if (const char* e = getenv("ROBSEG_LOSS_GENERIC_VEC")) vec = atoi(e), min_warps = 8;
    if (vec != 1 && vec != 2 && vec != 4) vec = 4;
    while (vec > 1 && (size_t)min_warps * 2 * stage * vec > kSmemBudget) vec >>= 1;
    if (vec > 1) {
      const size_t st = stage * vec;
      int warps_sm = (int)(kSmemBudget / (2 * st));
      if (warps_sm > 32) warps_sm = 32;
      const int W = 8, ctas_sm = warps_sm / W;
      const size_t smem = (size_t)W * 2 * st;
      auto kern = vec == 4 ? loss_generic_strided_kernel<4> : loss_generic_strided_kernel<2>;
      ROBSEG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      p.tiles_per_img = *tiles_per_img = (int)((p.HW + 32 * vec - 1) / (32 * vec));
      p.num_tiles = p.B * p.tiles_per_img;
      int grid = sm_count() * ctas_sm;
      const int max_useful = (p.num_tiles + W - 1) / W;
      if (grid > max_useful) grid = max_useful;
      if (grid < 1) grid = 1;
      kern<<<grid, 32 * W, smem, stream>>>(p, 2);
      ROBSEG_LAUNCH_CHECK();
      return 0;
    }
    // as many warps per SM as shared memory allows (<= 64), two stages per warp when they fit
    int warps_sm = (int)(kSmemBudget / stage);
    if (warps_sm > 64) warps_sm = 64;
    int K = (size_t)2 * stage * (warps_sm >= 32 ? 32 : warps_sm) <= kSmemBudget && warps_sm >= 16 ? 2 : 1;
    if (K == 2) warps_sm = (int)(kSmemBudget / (2 * stage)) > 64 ? 64 : (int)(kSmemBudget / (2 * stage));
    int W = warps_sm >= 8 ? 8 : warps_sm;  // warps per CTA
    int ctas_sm = warps_sm / W;
    if (ctas_sm < 1) ctas_sm = 1;
    const size_t smem = (size_t)W * K * stage;
    ROBSEG_CUDA(cudaFuncSetAttribute(loss_generic_f32_kernel,
                                     cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int grid = sm_count() * ctas_sm;
    const int max_useful = (p.num_tiles + W - 1) / W;
    if (grid > max_useful) grid = max_useful;
    if (grid < 1) grid = 1;
    loss_generic_f32_kernel<<<grid, 32 * W, smem, stream>>>(p, K);
    ROBSEG_LAUNCH_CHECK();
    return 0;
  }
  int W = (int)(kSmemBudget / 2 / stage);  // aim for two CTAs per SM
  if (W > 8) W = 8;
  if (W < 1) W = 1;
  const size_t smem = W * stage;
  auto kern = loss_generic_kernel<T>;
  ROBSEG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int grid = sm_count() * 2;
  const int max_useful = (p.num_tiles + W - 1) / W;
  if (grid > max_useful) grid = max_useful;
  if (grid < 1) grid = 1;
  kern<<<grid, 32 * W, smem, stream>>>(p);
  ROBSEG_LAUNCH_CHECK();
  return 0;
}

static int tiles_upper_bound(int B, int64_t HW) { return B * (int)((HW + 31) / 32); }

// shared with loss_up_kernel.cu: fixed-order reduction of per-tile partials into per-image outputs
int launch_loss_finalize(const float4* partials, int B, int tiles_per_img, const float* grad_scale,
                         int64_t HW, float* loss_img, float* track_img, int32_t* correct_img,
                         int32_t* valid_img, cudaStream_t stream) {
  loss_finalize_kernel<<<B, 256, 0, stream>>>(partials, tiles_per_img, grad_scale, 1.0 / (double)HW,
                                              loss_img, track_img, correct_img, valid_img);
  ROBSEG_LAUNCH_CHECK();
  return 0;
}

}  // namespace robseg

using namespace robseg;

static size_t loss_partial_bytes(int B, int64_t HW) {
  return ((size_t)tiles_upper_bound(B, HW) * sizeof(float4) + 255) / 256 * 256;
}
extern "C" size_t robseg_loss_workspace_bytes(int B, int C, int64_t HW, int dtype) {
  (void)dtype;
  if (B <= 0 || HW <= 0 || C <= 0) return 0;
  return loss_partial_bytes(B, HW) + count_replica_bytes(B, C);  // per-tile partials | counter replicas
}

static int loss_fwd_bwd_impl(const void* logits, int dtype, const int64_t* labels,
                             const float* class_w, int loss_kind, int ignore_index, int B,
                             int C, int64_t HW, const float* grad_scale,
                             const float* upstream_pix, void* dlogits, float* loss_pix,
                             int64_t* pred, float* loss_img, float* track_img,
                             int32_t* correct_img, int32_t* valid_img, int64_t* counts, void* workspace,
                             size_t workspace_bytes, robseg_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  // one-shot measurement events (robseg_profile_next_kernel): consumed by this call whatever its outcome
  const cudaEvent_t prof_start = take_profile_start(), prof_stop = take_profile_stop();
  ROBSEG_REQUIRE(logits && labels, "logits/labels must not be NULL");
  ROBSEG_REQUIRE(dtype == ROBSEG_F32 || dtype == ROBSEG_BF16, "unsupported dtype %d", dtype);
  ROBSEG_REQUIRE(loss_kind >= ROBSEG_LOSS_CE && loss_kind <= ROBSEG_LOSS_ARGMAX,
                 "unknown loss kind %d", loss_kind);
  ROBSEG_REQUIRE(B > 0 && C > 0 && HW > 0, "bad shape B=%d C=%d HW=%lld", B, C, (long long)HW);
  ROBSEG_REQUIRE((int64_t)B * ((HW + 31) / 32) < (1ll << 31), "too many tiles");
  ROBSEG_REQUIRE(workspace && workspace_bytes >= robseg_loss_workspace_bytes(B, C, HW, dtype),
                 "workspace too small (%zu bytes)", workspace_bytes);
  ROBSEG_REQUIRE(reinterpret_cast<uintptr_t>(workspace) % 16 == 0, "workspace must be 16B aligned");

  LossParams p{};
  p.logits = logits, p.labels = labels, p.class_w = class_w, p.grad_scale = grad_scale;
  p.upstream = upstream_pix, p.dlogits = dlogits, p.loss_pix = loss_pix, p.pred = pred;
  p.partials = static_cast<float4*>(workspace);
  p.HW = HW, p.kind = loss_kind, p.ignore_index = ignore_index, p.B = B, p.C = C;
  p.inv_hw = (float)(1.0 / (double)HW);
  if (counts != nullptr) {
    p.counts = reinterpret_cast<unsigned long long*>(static_cast<uint8_t*>(workspace) + loss_partial_bytes(B, HW));
    p.n_rep = count_replicas(C);
    const int zrc = launch_counts_zero(p.counts, count_replica_bytes(B, C), stream);
    if (zrc != 0) return zrc;
  }

  const int esize = dtype == ROBSEG_F32 ? 4 : 2;
  auto al16 = [](const void* q) { return q == nullptr || reinterpret_cast<uintptr_t>(q) % 16 == 0; };
  const bool tma_ok = (HW * esize) % 16 == 0 && C <= 256 && al16(logits) && al16(dlogits) &&
                      al16(labels) && al16(pred) && al16(loss_pix) && al16(upstream_pix) &&
                      (size_t)C * 32 * esize * (esize == 2 ? 2 : 1) * 2 <= kSmemBudget - 512;
  int rc;
  int tiles_per_img;
  if (prof_start) cudaEventRecord(prof_start, stream);
  if (tma_ok) {
    // (argmax-only launches are pure streaming reads: wide rows again)
    const int vec = pick_vec(C, esize, dlogits != nullptr || loss_kind == ROBSEG_LOSS_ARGMAX);
    tiles_per_img = (int)((HW + 32 * vec - 1) / (32 * vec));
    // (gradient launches only: loss-only / argmax launches are read-only streams that do better with the wide
    // 512-byte rows pick_vec gives them -- measured at 24x21x472^2: 0.120 vs 0.146 ms)
    if (dtype == ROBSEG_F32 && C <= kRegChannels && dlogits != nullptr && creg_enabled() && !getenv("ROBSEG_LOSS_VEC") &&
        !getenv("ROBSEG_LOSS_G")) {
      tiles_per_img = (int)((HW + 63) / 64);
      rc = launch_tma<float, 2, 1, kRegChannels>(p, stream);
    } else if (dtype == ROBSEG_F32) {
      rc = vec == 4 ? launch_tma_pick<float, 4>(p, stream)
                    : vec == 2 ? launch_tma_pick<float, 2>(p, stream)
                               : launch_tma_pick<float, 1>(p, stream);
    } else {
      rc = vec == 8 ? launch_tma_pick<__nv_bfloat16, 8>(p, stream)
                    : vec == 4 ? launch_tma_pick<__nv_bfloat16, 4>(p, stream)
                               : launch_tma_pick<__nv_bfloat16, 2>(p, stream);
    }
  } else {
    rc = dtype == ROBSEG_F32 ? launch_generic<float>(p, stream, &tiles_per_img)
                             : launch_generic<__nv_bfloat16>(p, stream, &tiles_per_img);
  }
  if (prof_stop) cudaEventRecord(prof_stop, stream);
  if (rc != 0) return rc;
  if (counts != nullptr) {
    rc = launch_counts_fold(p.counts, p.n_rep, B, C, counts, stream);
    if (rc != 0) return rc;
  }
  if (loss_img || track_img || correct_img || valid_img) {
    loss_finalize_kernel<<<B, 256, 0, stream>>>(p.partials, tiles_per_img, grad_scale,
                                                1.0 / (double)HW, loss_img, track_img,
                                                correct_img, valid_img);
    ROBSEG_LAUNCH_CHECK();
  }
  return 0;
}

extern "C" int robseg_loss_fwd_bwd(const void* logits, int dtype, const int64_t* labels,
                                   const float* class_w, int loss_kind, int ignore_index, int B,
                                   int C, int64_t HW, const float* grad_scale,
                                   const float* upstream_pix, void* dlogits, float* loss_pix,
                                   int64_t* pred, float* loss_img, float* track_img,
                                   int32_t* correct_img, int32_t* valid_img, void* workspace,
                                   size_t workspace_bytes, robseg_stream_t stream) {
  return loss_fwd_bwd_impl(logits, dtype, labels, class_w, loss_kind, ignore_index, B, C, HW, grad_scale,
                           upstream_pix, dlogits, loss_pix, pred, loss_img, track_img, correct_img, valid_img,
                           nullptr, workspace, workspace_bytes, stream);
}

extern "C" int robseg_loss_fwd_bwd_counts(const void* logits, int dtype, const int64_t* labels,
                                          const float* class_w, int loss_kind, int ignore_index, int B,
                                          int C, int64_t HW, const float* grad_scale,
                                          const float* upstream_pix, void* dlogits, float* loss_pix,
                                          int64_t* pred, float* loss_img, float* track_img,
                                          int32_t* correct_img, int32_t* valid_img, int64_t* counts,
                                          void* workspace, size_t workspace_bytes, robseg_stream_t stream) {
  ROBSEG_REQUIRE(counts != nullptr, "counts must not be NULL");
  return loss_fwd_bwd_impl(logits, dtype, labels, class_w, loss_kind, ignore_index, B, C, HW, grad_scale,
                           upstream_pix, dlogits, loss_pix, pred, loss_img, track_img, correct_img, valid_img,
                           counts, workspace, workspace_bytes, stream);
}
