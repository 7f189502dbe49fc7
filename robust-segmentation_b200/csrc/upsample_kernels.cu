// Bilinear up-sampling (align_corners=False) of the consumer's low-resolution logits to the
// input size, forward and backward: SURVEY.md section 8f rank 1, the caller-side neighbour of
// the loss kernel (semseg/models/uperforseg.py:416-418, semseg/models/segmenter.py:228).
//
// The stock ATen pair costs ~19 % of the GPU time of a SEA step on B200 (profiles/
// r01_launches_bench.md): the forward runs far below write bandwidth and the backward scatters
// with float atomics.  Here the forward is a pure streaming write (each thread produces four
// adjacent outputs, the 16x smaller input stays in L1/L2) and the backward is a deterministic
// GATHER: a CTA stages the output-gradient region that touches its tile of input cells in
// shared memory (coalesced, each element read once from HBM), reduces along x, then along y.
// No atomics, bit-reproducible.  Index and weight arithmetic follows ATen's
// area_pixel_compute_source_index: src = scale*(dst+0.5)-0.5 clamped at 0, i0 = floor(src),
// i1 = i0 + (i0 < in-1), w1 = src - i0, w0 = 1 - w1.
#include "common.cuh"

#include <cstdlib>

namespace robseg {

struct Tap {
  int i0, i1;
  float w0, w1;
};

// Plane groups (gridDim.z) of the kernels whose blocks loop over planes p = z, z + gz, ...: at most `want`, and
// then the smallest count that keeps the same number of loop trips, so every block runs ceil(planes / gz) or one
// fewer planes.  (With gz = want the trip counts were e.g. 3 and 4 -- 8192 planes over 2368 groups -- and the
// kernel's tail was a quarter of its duration; measured on the pow2 backward: 145 -> 117 us at 16x512x32^2 -> 128^2.)
static inline unsigned plane_groups(int64_t want, int64_t planes) {
  if (want < 1) want = 1;
  if (want > planes) want = planes;
  if (want > 65535) want = 65535;
  const int64_t trips = (planes + want - 1) / want;
  return (unsigned)((planes + trips - 1) / trips);
}

__device__ __forceinline__ Tap make_tap(int dst, float scale, int in_size) {
  float src = scale * ((float)dst + 0.5f) - 0.5f;
  src = src < 0.f ? 0.f : src;
  Tap t;
  t.i0 = min((int)src, in_size - 1);
  t.i1 = t.i0 + (t.i0 < in_size - 1 ? 1 : 0);
  t.w1 = src - (float)t.i0;
  t.w0 = 1.f - t.w1;
  return t;
}

// out[p, Y, X..X+3]; grid (ceil(H*ceil(W/4)/128), 1, plane groups), block 128: a thread owns one
// quad of the flattened plane (narrow planes still fill their blocks).  The taps depend only on
// (Y, X), so a thread computes them once and then streams over its planes.  For up-sampling
// ratios >= 2 the four outputs of a thread read at most three adjacent input columns: 6 loads
// per 16-byte store instead of 16.
__global__ void __launch_bounds__(128)
    upsample_fwd_kernel(const float* __restrict__ in, float* __restrict__ out, int64_t planes, int h,
                        int w, int H, int W, float sy, float sx) {
  const int quads = (W + 3) / 4;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= H * quads) return;
  const int Y = t / quads, X0 = (t - Y * quads) * 4;
  const Tap ty = make_tap(Y, sy, h);
  Tap tx[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) tx[j] = make_tap(min(X0 + j, W - 1), sx, w);
  const bool vec = (W % 4 == 0) && (reinterpret_cast<uintptr_t>(out) % 16 == 0);
  const int cbase = tx[0].i0;
  const bool narrow = tx[3].i1 - cbase <= 2;
  const int c1 = min(cbase + 1, w - 1), c2 = min(cbase + 2, w - 1);
  int o0[4], o1[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) o0[j] = tx[j].i0 - cbase, o1[j] = tx[j].i1 - cbase;
  for (int64_t p = blockIdx.z; p < planes; p += gridDim.z) {
    const float* r0 = in + (p * h + ty.i0) * w;
    const float* r1 = in + (p * h + ty.i1) * w;
    float v[4];
    if (narrow) {
      const float a0 = __ldg(r0 + cbase), a1 = __ldg(r0 + c1), a2 = __ldg(r0 + c2);
      const float b0 = __ldg(r1 + cbase), b1 = __ldg(r1 + c1), b2 = __ldg(r1 + c2);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float a = o0[j] == 0 ? a0 : (o0[j] == 1 ? a1 : a2);
        const float b = o1[j] == 0 ? a0 : (o1[j] == 1 ? a1 : a2);
        const float c = o0[j] == 0 ? b0 : (o0[j] == 1 ? b1 : b2);
        const float d = o1[j] == 0 ? b0 : (o1[j] == 1 ? b1 : b2);
        v[j] = ty.w0 * (tx[j].w0 * a + tx[j].w1 * b) + ty.w1 * (tx[j].w0 * c + tx[j].w1 * d);
      }
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float a = __ldg(r0 + tx[j].i0), b = __ldg(r0 + tx[j].i1);
        const float c = __ldg(r1 + tx[j].i0), d = __ldg(r1 + tx[j].i1);
        v[j] = ty.w0 * (tx[j].w0 * a + tx[j].w1 * b) + ty.w1 * (tx[j].w0 * c + tx[j].w1 * d);
      }
    }
    float* o = out + (p * H + Y) * W + X0;
    if (vec) {
      __stcs(reinterpret_cast<float4*>(o), make_float4(v[0], v[1], v[2], v[3]));
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (X0 + j < W) o[j] = v[j];
    }
  }
}

// Forward for any up-sampling ratio whose rows cannot be stored as aligned vectors (W % 4 != 0: the reference's
// 119 -> 473 PASCAL-VOC logits).  The quad kernel above then writes four 4-byte stores per thread at a 16-byte lane
// stride -- every store instruction touches 16 sectors partially -- and ran at 35 % of the roofline.  Here a lane owns
// ONE output column and a block a strip of 32 output rows: every store instruction of a warp is 128 contiguous bytes
// and the x-taps are per-lane constants.  A thread keeps the two x-interpolated input rows its current output row
// reads (top, bot) in registers and re-loads only when the walk crosses an input row (every ~ratio rows; the row that
// was "bottom" becomes "top").  Row taps come from a per-strip shared-memory table (one broadcast LDS.128 per row).
// ATen's operation order: w0y * (w0x*a + w1x*b) + w1y * (w0x*c + w1x*d).
// Three restructurings were measured and were slower (profiles/r02_upsample_walk_forward.md): staging every input row
// of the strip in shared memory (per-row or per-interval tables: 20-25 instructions per output) and prefetching the
// next input row into a third register.
constexpr int kWalkStrip = 32;
__global__ void __launch_bounds__(128)
    upsample_fwd_walk_kernel(const float* __restrict__ in, float* __restrict__ out, int64_t planes, int h, int w,
                             int H, int W, float sy, float sx) {
  __shared__ float4 ytab[kWalkStrip];  // (i0, i1 as int bits, w0, w1) per output row of the strip
  const int X = blockIdx.x * blockDim.x + threadIdx.x;
  const int Y0 = blockIdx.y * kWalkStrip, rows = min(kWalkStrip, H - Y0);
  if (threadIdx.x < rows) {
    const Tap t = make_tap(Y0 + threadIdx.x, sy, h);
    ytab[threadIdx.x] = make_float4(__int_as_float(t.i0), __int_as_float(t.i1), t.w0, t.w1);
  }
  __syncthreads();
  if (X >= W) return;
  const Tap tx = make_tap(X, sx, w);
  for (int64_t p = blockIdx.z; p < planes; p += gridDim.z) {
    const float* base = in + p * (int64_t)h * w;
    float* o = out + (p * H + Y0) * (int64_t)W + X;
    int cur0 = -1, cur1 = -1;
    float top = 0.f, bot = 0.f;
    for (int r = 0; r < rows; ++r, o += W) {
      const float4 t = ytab[r];
      const int i0 = __float_as_int(t.x), i1 = __float_as_int(t.y);
      if (i0 != cur0) {  // warp-uniform: all lanes walk the same rows
        if (i0 == cur1) {
          top = bot;
        } else {
          const float* rp = base + (int64_t)i0 * w;
          top = tx.w0 * __ldg(rp + tx.i0) + tx.w1 * __ldg(rp + tx.i1);
        }
        cur0 = i0;
      }
      if (i1 != cur1) {
        if (i1 == i0) {
          bot = top;
        } else {
          const float* rp = base + (int64_t)i1 * w;
          bot = tx.w0 * __ldg(rp + tx.i0) + tx.w1 * __ldg(rp + tx.i1);
        }
        cur1 = i1;
      }
      __stcs(o, t.z * top + t.w * bot);
    }
  }
}

static int launch_fwd_walk(const float* in, float* out, int64_t planes, int h, int w, int H, int W, float sy,
                           float sx, cudaStream_t stream) {
  int bs = 128;  // the block width that wastes the fewest lanes on the last block of a row
  for (int cand : {96, 64})
    if ((W + cand - 1) / cand * cand < (W + bs - 1) / bs * bs) bs = cand;
  const int gx = (W + bs - 1) / bs, gy = (H + kWalkStrip - 1) / kWalkStrip;
  int64_t gz = ((int64_t)sm_count() * 16 * (128 / bs) + (int64_t)gx * gy - 1) / ((int64_t)gx * gy);
  gz = plane_groups(gz, planes);
  upsample_fwd_walk_kernel<<<dim3(gx, gy, (unsigned)gz), bs, 0, stream>>>(in, out, planes, h, w, H, W, sy, sx);
  ROBSEG_LAUNCH_CHECK();
  return 0;
}

// Deterministic gather backward.  Block = 256 threads, tile TY x TX input cells; the block
// builds the per-cell tap tables once (they depend only on the cell), then loops over its
// planes: stage the output-gradient region in shared memory (each element read once, coalesced),
// reduce along x, reduce along y, write each input cell exactly once.
// Shared layout: region [RY][RXP] | xred [TX][RYP] | wx [TX][KX] | wy [TY][KY] | xs [TX] | ys [TY]
__global__ void __launch_bounds__(256)
    upsample_bwd_kernel(const float* __restrict__ gout, float* __restrict__ gin, int64_t planes, int C,
                        int64_t bs, int64_t cs, int h, int w, int H, int W, float sy, float sx, int TY,
                        int TX, int RY, int RX, int KY, int KX) {
  extern __shared__ float shm[];
  const int RXP = RX | 1, RYP = RY | 1;  // odd strides: conflict-free column walks
  float* region = shm;
  float* xred = region + (size_t)RY * RXP;
  float* wx = xred + (size_t)TX * RYP;
  float* wy = wx + (size_t)TX * KX;
  int* xs = reinterpret_cast<int*>(wy + (size_t)TY * KY);
  int* ys = xs + TX;
  const int cy0 = blockIdx.y * TY, cx0 = blockIdx.x * TX;
  const float inv_sy = 1.f / sy, inv_sx = 1.f / sx;
  // first output row/col whose taps can touch the tile: i1 >= c0  <=>  src > c0 - 1
  int Y0 = (int)floorf(((float)cy0 - 1.f + 0.5f) * inv_sy - 0.5f) - 1;
  int X0 = (int)floorf(((float)cx0 - 1.f + 0.5f) * inv_sx - 0.5f) - 1;
  Y0 = max(Y0, 0), X0 = max(X0, 0);
  X0 &= ~3;  // keep rows 16-byte aligned for the vector loads
  // ---- tap tables (once per block) ---------------------------------------------------------------
  for (int i = threadIdx.x; i < TX + TY; i += blockDim.x) {
    const bool isx = i < TX;
    const int c = isx ? cx0 + i : cy0 + (i - TX);
    const int in_size = isx ? w : h, out_size = isx ? W : H, K = isx ? KX : KY;
    const float inv_s = isx ? inv_sx : inv_sy, sc = isx ? sx : sy;
    float* wt = isx ? wx + (size_t)i * KX : wy + (size_t)(i - TX) * KY;
    int a = (int)floorf(((float)c - 1.f + 0.5f) * inv_s - 0.5f) - 1;
    a = max(a, isx ? X0 : Y0);
    for (int k = 0; k < K; ++k) {
      const int o = a + k;
      float wgt = 0.f;
      if (c < in_size && o < out_size) {
        const Tap t = make_tap(o, sc, in_size);
        wgt = (t.i0 == c ? t.w0 : 0.f) + (t.i1 == c ? t.w1 : 0.f);
      }
      wt[k] = wgt;
    }
    (isx ? xs[i] : ys[i - TX]) = a - (isx ? X0 : Y0);
  }
  const bool vec = (W % 4 == 0) && (RX % 4 == 0) && (reinterpret_cast<uintptr_t>(gout) % 16 == 0) &&
                   bs % 4 == 0 && cs % 4 == 0;
  for (int64_t p = blockIdx.z; p < planes; p += gridDim.z) {
    __syncthreads();  // tables ready / previous plane's xred consumed
    const float* src = gout + (p / C) * bs + (p % C) * cs;
    // ---- stage the region (zero beyond the image) --------------------------------------------
    if (vec) {
      const int nx4 = RX / 4;
      for (int i = threadIdx.x; i < RY * nx4; i += blockDim.x) {
        const int r = i / nx4, c4 = i - r * nx4;
        const int Y = Y0 + r, X = X0 + 4 * c4;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (Y < H && X < W) v = __ldcs(reinterpret_cast<const float4*>(src + (int64_t)Y * W + X));
        float* d = region + r * RXP + 4 * c4;
        d[0] = v.x, d[1] = v.y, d[2] = v.z, d[3] = v.w;
      }
    } else {
      for (int i = threadIdx.x; i < RY * RX; i += blockDim.x) {
        const int r = i / RX, c = i - r * RX;
        const int Y = Y0 + r, X = X0 + c;
        region[r * RXP + c] = (Y < H && X < W) ? __ldcs(src + (int64_t)Y * W + X) : 0.f;
      }
    }
    __syncthreads();
    // ---- reduce along x: lanes walk rows (odd stride), one cell column per group of RY threads -
    for (int i = threadIdx.x; i < RY * TX; i += blockDim.x) {
      const int cx = i / RY, r = i - cx * RY;
      const float* row = region + r * RXP + xs[cx];
      const float* wt = wx + (size_t)cx * KX;
      float acc = 0.f;
      for (int k = 0; k < KX; ++k) acc = fmaf(wt[k], row[k], acc);
      xred[cx * RYP + r] = acc;
    }
    __syncthreads();
    // ---- reduce along y and write each input cell once -------------------------------------------
    for (int i = threadIdx.x; i < TY * TX; i += blockDim.x) {
      const int cy = i / TX, cx = i - cy * TX;
      const int celly = cy0 + cy, cellx = cx0 + cx;
      if (celly >= h || cellx >= w) continue;
      const float* col = xred + cx * RYP + ys[cy];
      const float* wt = wy + (size_t)cy * KY;
      float acc = 0.f;
      for (int k = 0; k < KY; ++k) acc = fmaf(wt[k], col[k], acc);
      gin[(p * h + celly) * (int64_t)w + cellx] = acc;
    }
  }
}

// ---- exact power-of-two specialisations (H == R*h, W == R*w, R = 2, 4, 8, 16) -------------------
// x4 is UperNet's final logit up-sampling; x2 / x4 / x8 are the feature-pyramid up-samplings of
// its decode head (semseg/models/uperforseg.py:282-303); x16 is Segmenter's (segmenter.py:228).  Output block (R rows x R cols) <-> input
// cell (a, b): its taps only touch cells a-1..a+1 x b-1..b+1.  Dense 3-tap weight rows (zeros
// included, image borders folded in through make_tap) make both directions branch-free.
template <int R>
struct WR {
  float k[R][3];  // k[j][c]: weight of input (base-1+c) for output R*base+j
};

template <int R>
__device__ __forceinline__ WR<R> make_wr(int base, int in_size) {
  WR<R> r;
#pragma unroll
  for (int j = 0; j < R; ++j) {
    const Tap t = make_tap(R * base + j, 1.f / R, in_size);
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const int cell = base - 1 + c;
      r.k[j][c] = (t.i0 == cell ? t.w0 : 0.f) + (t.i1 == cell ? t.w1 : 0.f);
    }
  }
  return r;
}

// R adjacent floats, vector width min(R, 4); p is R*4-byte aligned (16 for R = 8)
template <int R>
__device__ __forceinline__ void store_row(float* p, const float (&v)[R]) {
  if constexpr (R == 2) {
    __stcs(reinterpret_cast<float2*>(p), make_float2(v[0], v[1]));
  } else {
#pragma unroll
    for (int q = 0; q < R / 4; ++q)
      __stcs(reinterpret_cast<float4*>(p) + q, make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]));
  }
}
template <int R>
__device__ __forceinline__ void load_row(const float* p, float (&v)[R]) {
  if constexpr (R == 2) {
    const float2 t = __ldcs(reinterpret_cast<const float2*>(p));
    v[0] = t.x, v[1] = t.y;
  } else {
#pragma unroll
    for (int q = 0; q < R / 4; ++q) {
      const float4 t = __ldcs(reinterpret_cast<const float4*>(p) + q);
      v[4 * q] = t.x, v[4 * q + 1] = t.y, v[4 * q + 2] = t.z, v[4 * q + 3] = t.w;
    }
  }
}

// grid (ceil(h*w/128), 1, plane groups), block 128: thread = one input cell (a, b) of the
// flattened plane (narrow planes still fill their blocks), produces the RxR output block from 9
// loads (separable: 3 rows x R horizontal outputs, then vertical), stores R coalesced rows.
template <int R>
__global__ void __launch_bounds__(128)
    upsample_fwd_pow2_kernel(const float* __restrict__ in, float* __restrict__ out, int64_t planes,
                             int h, int w) {
  // CW output columns per thread, R / CW threads per input cell.  A thread that owned all 8 columns of an x8 cell
  // wrote each output row with two 16-byte stores 32 bytes apart from its neighbour lane's: every store instruction
  // filled half of each 32-byte sector it touched, twice the L2 transactions per byte (61 % of the roofline at
  // 16x512x16^2 -> 128^2 against 81-90 % at x4).  With 4 columns per thread consecutive lanes write consecutive
  // 16-byte pieces: one instruction = 512 contiguous bytes of an output row.
  constexpr int CW = R < 4 ? R : 4, PARTS = R / CW;
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  const int cell = idx / PARTS, part = idx - cell * PARTS;
  if (cell >= h * w) return;
  const int a = cell / w, b = cell - a * w;
  const int H = R * h, W = R * w;
  const WR<R> ky = make_wr<R>(a, h);
  float kx[CW][3];  // my CW columns of the cell's x-weights
  {
    const WR<R> all = make_wr<R>(b, w);
#pragma unroll
    for (int j = 0; j < CW; ++j)
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        float v = all.k[j][c];
#pragma unroll
        for (int q = 1; q < PARTS; ++q) v = part == q ? all.k[q * CW + j][c] : v;
        kx[j][c] = v;
      }
  }
  const int cm = max(b - 1, 0), cp = min(b + 1, w - 1);
  const int rm = max(a - 1, 0), rp = min(a + 1, h - 1);
  for (int64_t p = blockIdx.z; p < planes; p += gridDim.z) {
    const float* base = in + p * (int64_t)h * w;
    float hx[3][CW];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      const float* row = base + (int64_t)(r == 0 ? rm : (r == 1 ? a : rp)) * w;
      const float v0 = __ldg(row + cm), v1 = __ldg(row + b), v2 = __ldg(row + cp);
#pragma unroll
      for (int j = 0; j < CW; ++j) hx[r][j] = kx[j][0] * v0 + kx[j][1] * v1 + kx[j][2] * v2;
    }
    float* o = out + (p * H + R * a) * (int64_t)W + R * b + part * CW;
#pragma unroll
    for (int i = 0; i < R; ++i) {
      float v[CW];
#pragma unroll
      for (int j = 0; j < CW; ++j)
        v[j] = ky.k[i][0] * hx[0][j] + ky.k[i][1] * hx[1][j] + ky.k[i][2] * hx[2][j];
      store_row<CW>(o + (int64_t)i * W, v);
    }
  }
}

// x2 with even h, w (the decode head's top-down path, uperforseg.py:282-303): one input cell per thread
// gives 16 bytes of output per 9 loads and ran at 53 % of the roofline, latency-bound (few bytes in flight
// per warp).  Here a thread owns a 2x2 block of input cells -> a 4x4 output block: 4 rows x (scalar +
// 8-byte + scalar) loads, 4 rows x one 16-byte store (a warp writes 512 contiguous bytes per row), and
// the next plane's 12 loads are issued before the current plane is reduced.  Every output has exactly two
// taps per axis at compile-time positions of the 4x4 neighbourhood (rows / columns a0-1 .. a0+2, clamped
// at the borders, where make_tap's weights (1, 0) / (w0, w1) reproduce ATen's clamped expression).
struct X2In {
  float v[4][4];
};
__device__ __forceinline__ void x2_load(X2In& q, const float* __restrict__ base, const int (&row)[4], int w,
                                        int cm, int b0, int cp) {
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const float* rp = base + (int64_t)row[r] * w;
    const float2 mid = __ldg(reinterpret_cast<const float2*>(rp + b0));
    q.v[r][0] = __ldg(rp + cm), q.v[r][1] = mid.x, q.v[r][2] = mid.y, q.v[r][3] = __ldg(rp + cp);
  }
}

__global__ void __launch_bounds__(128, 7)
    upsample_fwd_x2_kernel(const float* __restrict__ in, float* __restrict__ out, int64_t planes, int h,
                           int w) {
  const int hw2 = w >> 1;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (h >> 1) * hw2) return;
  const int a0 = 2 * (t / hw2), b0 = 2 * (t % hw2);
  const int H = 2 * h, W = 2 * w;
  // output j (0..3) along an axis: cell j/2, sub-position j%2; taps at neighbourhood slots
  // (j/2 + j%2, j/2 + j%2 + 1).  For x2 ATen's weights (make_tap) are the constants (w0, w1) = (.25, .75) for
  // even and (.75, .25) for odd outputs -- exactly, src = 0.5*(dst+0.5)-0.5 has no rounding -- except output 0
  // of the image, whose clamped source coordinate gives (1, 0): one run-time weight per axis.
  const float lx1 = b0 == 0 ? 0.f : 0.75f, lx0 = 1.f - lx1;
  const float ly1 = a0 == 0 ? 0.f : 0.75f, ly0 = 1.f - ly1;
  const float wx0[4] = {lx0, 0.75f, 0.25f, 0.75f}, wx1[4] = {lx1, 0.25f, 0.75f, 0.25f};
  const float wy0[4] = {ly0, 0.75f, 0.25f, 0.75f}, wy1[4] = {ly1, 0.25f, 0.75f, 0.25f};
  const int cm = max(b0 - 1, 0), cp = min(b0 + 2, w - 1);
  const int row[4] = {max(a0 - 1, 0), a0, a0 + 1, min(a0 + 2, h - 1)};
  int64_t p = blockIdx.z;
  if (p >= planes) return;
  X2In cur, nxt;
  x2_load(cur, in + p * (int64_t)h * w, row, w, cm, b0, cp);
  for (; p < planes; p += gridDim.z) {
    const int64_t pn = p + gridDim.z;
    if (pn < planes) x2_load(nxt, in + pn * (int64_t)h * w, row, w, cm, b0, cp);
    float hx[4][4];  // [neighbourhood row][output column]
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int s = (j >> 1) + (j & 1);
        hx[r][j] = wx0[j] * cur.v[r][s] + wx1[j] * cur.v[r][s + 1];
      }
    float* o = out + (p * H + 2 * a0) * (int64_t)W + 2 * b0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int s = (i >> 1) + (i & 1);
      float v[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) v[j] = wy0[i] * hx[s][j] + wy1[i] * hx[s + 1][j];
      __stcs(reinterpret_cast<float4*>(o + (int64_t)i * W), make_float4(v[0], v[1], v[2], v[3]));
    }
    cur = nxt;
  }
}

// x2 backward with even h, w: the mirror image of upsample_fwd_x2_kernel.  A thread owns a 2x2 block of input
// cells and gathers their gradient from the 6x6 output block that touches them: 6 rows x (scalar + 16-byte +
// scalar) loads -- the vector loads of a warp are contiguous, the two scalars are its neighbours' words -- then
// two 8-byte stores.  Flat thread -> cell mapping (16- and 32-wide pyramid planes fill their warps, which the
// walk-down kernel's one-lane-per-column layout does not: 18 of 32 lanes at w = 16), 18 independent loads in
// flight per thread, every input cell written once, no atomics, fixed summation order (bit-reproducible).
// Weights: an interior cell receives (.25, .75, .75, .25) from output rows 2a-1 .. 2a+2 (ATen's taps for x2,
// exact in fp32); at the image border the missing outer row weighs 0 and the clamped one 1 instead of .75.
__global__ void __launch_bounds__(128, 6)
    upsample_bwd_x2_kernel(const float* __restrict__ gout, float* __restrict__ gin, int64_t planes, int C,
                           int64_t bs, int64_t cs, int h, int w) {
  const int hw2 = w >> 1;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (h >> 1) * hw2) return;
  const int a0 = 2 * (t / hw2), b0 = 2 * (t % hw2);
  const int H = 2 * h, W = 2 * w;
  // run-time weights of the four border-sensitive taps per axis
  const float xa0 = b0 == 0 ? 0.f : 0.25f, xa1 = b0 == 0 ? 1.f : 0.75f;
  const float xb2 = b0 + 2 == w ? 1.f : 0.75f, xb3 = b0 + 2 == w ? 0.f : 0.25f;
  const float ya0 = a0 == 0 ? 0.f : 0.25f, ya1 = a0 == 0 ? 1.f : 0.75f;
  const float yb2 = a0 + 2 == h ? 1.f : 0.75f, yb3 = a0 + 2 == h ? 0.f : 0.25f;
  const int X0 = 2 * b0;                                    // first of my four aligned output columns
  const int xl = max(X0 - 1, 0), xr = min(X0 + 4, W - 1);   // clamped neighbours (weight 0 when clamped)
  int row[6];
#pragma unroll
  for (int r = 0; r < 6; ++r) row[r] = min(max(2 * a0 - 1 + r, 0), H - 1);
  for (int64_t p = blockIdx.z; p < planes; p += gridDim.z) {
    const float* base = gout + (p / C) * bs + (p % C) * cs;
    float c0[6], c1[6];  // x-reduced rows for my two cell columns
#pragma unroll
    for (int r = 0; r < 6; ++r) {
      const float* rp = base + (int64_t)row[r] * W;
      const float4 m = __ldcs(reinterpret_cast<const float4*>(rp + X0));
      const float l = __ldg(rp + xl), q = __ldg(rp + xr);
      c0[r] = ((xa0 * l + xa1 * m.x) + 0.75f * m.y) + 0.25f * m.z;
      c1[r] = ((0.25f * m.y + 0.75f * m.z) + xb2 * m.w) + xb3 * q;
    }
    float* o = gin + (p * h + a0) * (int64_t)w + b0;
    __stcs(reinterpret_cast<float2*>(o),
           make_float2(((ya0 * c0[0] + ya1 * c0[1]) + 0.75f * c0[2]) + 0.25f * c0[3],
                       ((ya0 * c1[0] + ya1 * c1[1]) + 0.75f * c1[2]) + 0.25f * c1[3]));
    __stcs(reinterpret_cast<float2*>(o + w),
           make_float2(((0.25f * c0[2] + 0.75f * c0[3]) + yb2 * c0[4]) + yb3 * c0[5],
                       ((0.25f * c1[2] + 0.75f * c1[3]) + yb2 * c1[4]) + yb3 * c1[5]));
  }
}

// grid (x groups, strips, plane groups), block = up to 8 warps side by side.  A warp covers
// `owned` (<= 30) input cell columns plus one halo lane on each side (the halo lanes only feed
// their neighbours), so the x-reduction needs no cross-warp traffic and no special cases.  A lane
// walks down the output rows of its strip of cell rows [a0, a1): R coalesced floats per output
// row straight from global memory, G rows in flight while the previous G are reduced;
// x-reduction in registers with two shuffles, y-reduction in three rotating accumulators; each
// input cell is written once.  No atomics.  Plane p = (p / C, p % C) of a [N, C, H, W] view with
// batch / channel strides bs / cs (elements): gradients of torch.cat slices are read in place.
template <int R>
__global__ void __launch_bounds__(256)
    upsample_bwd_pow2_kernel(const float* __restrict__ gout, float* __restrict__ gin, int64_t planes,
                             int C, int64_t bs, int64_t cs, int h, int w, int strip, int owned, int halo) {
  constexpr int G = R < 4 ? R : (R > 8 ? 2 : 4);  // output rows per prefetch group
  constexpr int KYS = (3 * R + 3) & ~3;       // padded floats of y-weights per cell row
  extern __shared__ __align__(16) float ky_tab[];  // [cell rows of the strip + halo][KYS]
  const int lane = threadIdx.x & 31;
  const int gw = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  // lane 0 / owned+1: halo columns (halo = 1).  A plane no wider than a warp needs none (halo = 0, lane = column):
  // the columns left of 0 / right of w-1 do not exist and the border cells' weights towards them are exactly zero,
  // so what the shuffles bring in from the unused lanes is multiplied out -- 32 of 32 lanes at w = 32 instead of 18.
  const int b = gw * owned + lane - halo;
  const int a0 = blockIdx.y * strip, a1 = min(a0 + strip, h);
  const int a_first = max(a0 - 1, 0), a_last = min(a1, h - 1);
  // the y-weights depend only on the cell row: build them once per block
  for (int i = threadIdx.x; i < (a_last - a_first + 1) * R; i += blockDim.x) {
    const int a = a_first + i / R, r = i % R;
    const Tap t = make_tap(R * a + r, 1.f / R, h);
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const int cell = a - 1 + c;
      ky_tab[(i / R) * KYS + r * 3 + c] = (t.i0 == cell ? t.w0 : 0.f) + (t.i1 == cell ? t.w1 : 0.f);
    }
  }
  __syncthreads();
  if (gw * owned >= w) return;  // whole warp past the image
  const int64_t W = (int64_t)R * w;
  const bool loads = b >= 0 && b < w && lane <= owned + 2 * halo - 1;
  const bool owns = loads && lane >= halo && lane < owned + halo;
  const int bcl = min(max(b, 0), w - 1);
  const WR<R> kx = make_wr<R>(bcl, w);  // transposed taps of my R outputs
  for (int64_t p = blockIdx.z; p < planes; p += gridDim.z) {
    const float* rowp = gout + (p / C) * bs + (p % C) * cs + (int64_t)R * a_first * W + R * bcl;
    float* outp = gin + (p * h + (a_first - 1)) * (int64_t)w + bcl;  // cell row a-1 of the first block row
    float acc_prev = 0.f, acc_cur = 0.f, acc_next = 0.f;  // cells a-1, a, a+1 of the current block row
    float v[G][R], nxt[G][R];
#pragma unroll
    for (int i = 0; i < G; ++i) {
#pragma unroll
      for (int j = 0; j < R; ++j) nxt[i][j] = 0.f;
      if (loads) load_row<R>(rowp + i * W, nxt[i]);
    }
    const float* kyp = ky_tab;
    for (int a = a_first; a <= a_last; ++a, kyp += KYS, outp += w) {
      float ky[KYS];
#pragma unroll
      for (int q = 0; q < KYS / 4; ++q) {
        const float4 t = *reinterpret_cast<const float4*>(kyp + 4 * q);
        ky[4 * q] = t.x, ky[4 * q + 1] = t.y, ky[4 * q + 2] = t.z, ky[4 * q + 3] = t.w;
      }
#pragma unroll
      for (int g = 0; g < R / G; ++g) {
        rowp += G * W;
#pragma unroll
        for (int i = 0; i < G; ++i)
#pragma unroll
          for (int j = 0; j < R; ++j) v[i][j] = nxt[i][j];
        if ((g + 1 < R / G || a < a_last) && loads) {  // prefetch the next G rows while these are reduced
#pragma unroll
          for (int i = 0; i < G; ++i) load_row<R>(rowp + i * W, nxt[i]);
        }
#pragma unroll
        for (int i = 0; i < G; ++i) {
          // partial sums of my R outputs for cells b-1, b, b+1
          float pl = 0.f, pc = 0.f, pr = 0.f;
#pragma unroll
          for (int j = 0; j < R; ++j) {
            pl = fmaf(kx.k[j][0], v[i][j], pl);
            pc = fmaf(kx.k[j][1], v[i][j], pc);
            pr = fmaf(kx.k[j][2], v[i][j], pr);
          }
          // lane-1's / lane+1's share for my cell (the shuffles hand lanes 0 / 31 their own value back: without
          // halo lanes those two own cells, whose missing neighbour contributes nothing)
          const float sl = __shfl_up_sync(0xffffffffu, pr, 1), sr = __shfl_down_sync(0xffffffffu, pl, 1);
          const float from_left = lane > 0 ? sl : 0.f, from_right = lane < 31 ? sr : 0.f;
          const float xr = (from_left + pc) + from_right;
          const int r = g * G + i;
          acc_prev = fmaf(ky[3 * r], xr, acc_prev);
          acc_cur = fmaf(ky[3 * r + 1], xr, acc_cur);
          acc_next = fmaf(ky[3 * r + 2], xr, acc_next);
        }
      }
      // cell row a-1 has now received everything (block rows a-2 .. a)
      if (owns && a - 1 >= a0 && a - 1 < a1) *outp = acc_prev;
      acc_prev = acc_cur, acc_cur = acc_next, acc_next = 0.f;
    }
    // the last cell row of the strip when the strip ends at the image border
    if (owns && a1 == h && h - 1 >= a0) gin[(p * h + (h - 1)) * (int64_t)w + b] = acc_prev;
  }
}

// ---- any up-sampling ratio >= 1 (e.g. 119 -> 473, the reference's PASCAL-VOC crops) -----------
// Same walk-down gather as the power-of-two kernel, with the output pixels assigned to cells by
// their FIRST tap: output X belongs to cell i0(X) (to the virtual cell -1 while the source
// coordinate is clamped at the left border), so it feeds cell own with weight wc and cell own+1
// with weight wr.  A lane owns one cell column and the <= RMAX outputs per row that belong to it;
// cell b needs its own partial sum plus its left neighbour's wr-sum: ONE shuffle, and only a left
// halo lane (lane 0).  Rows are walked the same way: the y-table holds, per output row, its cell
// row and the two weights; a cell row is emitted when the walk leaves it.
__device__ __forceinline__ int owner_of(int dst, float scale, int in_size) {
  const float src = scale * ((float)dst + 0.5f) - 0.5f;
  return src < 0.f ? -1 : min((int)src, in_size - 1);
}
// first output index whose owner is >= cell (monotone in dst), exact against owner_of
__device__ __forceinline__ int first_owned(int cell, float scale, float inv_scale, int in_size,
                                           int out_size) {
  int x = (int)ceilf(((float)cell + 0.5f) * inv_scale - 0.5f);
  x = max(0, min(x, out_size));
  while (x > 0 && owner_of(x - 1, scale, in_size) >= cell) --x;
  while (x < out_size && owner_of(x, scale, in_size) < cell) ++x;
  return x;
}

struct RowTap {
  int cell;      // owner cell row (-1: clamped rows, they only feed cell 0)
  float wc, wn;  // weights for cell `cell` and cell `cell + 1`
};

template <int RMAX>
__global__ void __launch_bounds__(256)
    upsample_bwd_walk_kernel(const float* __restrict__ gout, float* __restrict__ gin, int64_t planes,
                             int C, int64_t bs, int64_t cs, int h, int w, int H, int W, float sy,
                             float sx, int strip) {
  constexpr int G = RMAX <= 3 ? 4 : 2;  // output rows in flight
  extern __shared__ __align__(16) unsigned char walk_smem[];
  RowTap* ytab = reinterpret_cast<RowTap*>(walk_smem);
  const int lane = threadIdx.x & 31;
  const int gw = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int b = gw * 31 + lane - 1;  // lane 0: left halo (cell -1 = the clamped outputs for gw == 0)
  const int a0 = blockIdx.y * strip, a1 = min(a0 + strip, h);
  // output rows that touch cell rows [a0, a1): owners a0-1 .. a1-1
  const float inv_sy = 1.f / sy, inv_sx = 1.f / sx;
  const int Y0 = first_owned(a0 - 1, sy, inv_sy, h, H), Y1 = first_owned(a1, sy, inv_sy, h, H);
  for (int i = threadIdx.x; i < Y1 - Y0; i += blockDim.x) {
    const int Y = Y0 + i;
    const Tap t = make_tap(Y, sy, h);
    RowTap r;
    r.cell = owner_of(Y, sy, h);
    if (r.cell < 0) {
      r.wc = 0.f, r.wn = t.w0 + (t.i1 == 0 ? t.w1 : 0.f);  // clamped: everything goes to cell 0
    } else {
      r.wc = t.w0 + (t.i1 == t.i0 ? t.w1 : 0.f);
      r.wn = t.i1 == t.i0 ? 0.f : t.w1;
    }
    ytab[i] = r;
  }
  __syncthreads();
  if (gw * 31 >= w) return;  // whole warp past the image
  // my outputs per row: [x0, x0 + n), n <= RMAX (guaranteed by the host's choice of RMAX)
  const bool cell_ok = b >= -1 && b < w;
  const int x0 = cell_ok ? first_owned(b, sx, inv_sx, w, W) : 0;
  const int n = cell_ok ? first_owned(b + 1, sx, inv_sx, w, W) - x0 : 0;
  float wc[RMAX], wr[RMAX];
#pragma unroll
  for (int j = 0; j < RMAX; ++j) {
    wc[j] = wr[j] = 0.f;
    if (j < n) {
      const Tap t = make_tap(x0 + j, sx, w);
      if (b < 0) {
        wr[j] = t.w0 + (t.i1 == 0 ? t.w1 : 0.f);
      } else {
        wc[j] = t.w0 + (t.i1 == t.i0 ? t.w1 : 0.f);
        wr[j] = t.i1 == t.i0 ? 0.f : t.w1;
      }
    }
  }
  const bool owns = b >= 0 && b < w && lane >= 1;
  const int rows = Y1 - Y0;
  for (int64_t p = blockIdx.z; p < planes; p += gridDim.z) {
    const float* rowp = gout + (p / C) * bs + (p % C) * cs + (int64_t)Y0 * W + x0;
    float* outp = gin + p * (int64_t)h * w + b;
    float acc_cur = 0.f, acc_next = 0.f;
    int cur = rows > 0 ? ytab[0].cell : 0;
    float v[G][RMAX];
    auto load_rows = [&](int r0) {
#pragma unroll
      for (int i = 0; i < G; ++i)
#pragma unroll
        for (int j = 0; j < RMAX; ++j)
          v[i][j] = (j < n && r0 + i < rows) ? __ldcs(rowp + (int64_t)(r0 + i) * W + j) : 0.f;
    };
    load_rows(0);
    for (int r0 = 0; r0 < rows; r0 += G) {
      float xr[G];
#pragma unroll
      for (int i = 0; i < G; ++i) {
        float pc = 0.f, pr = 0.f;
#pragma unroll
        for (int j = 0; j < RMAX; ++j) pc = fmaf(wc[j], v[i][j], pc), pr = fmaf(wr[j], v[i][j], pr);
        xr[i] = pc + __shfl_up_sync(0xffffffffu, pr, 1);  // + lane-1's share for my cell
      }
      if (r0 + G < rows) load_rows(r0 + G);  // in flight while these rows are folded in
#pragma unroll
      for (int i = 0; i < G; ++i) {
        if (r0 + i < rows) {
          const RowTap t = ytab[r0 + i];
          if (t.cell != cur) {  // the walk leaves cell row `cur`: it has received everything
            if (owns && cur >= a0 && cur < a1) outp[(int64_t)cur * w] = acc_cur;
            acc_cur = acc_next, acc_next = 0.f, cur = t.cell;
          }
          acc_cur = fmaf(t.wc, xr[i], acc_cur);
          acc_next = fmaf(t.wn, xr[i], acc_next);
        }
      }
    }
    if (owns && cur >= a0 && cur < a1) outp[(int64_t)cur * w] = acc_cur;
    // (acc_next belongs to cell row a1: the next strip's, or none at the bottom border)
  }
}

template <int RMAX>
static int launch_bwd_walk(const float* gout, float* gin, int64_t planes, int C, int64_t bs, int64_t cs,
                           int h, int w, int H, int W, float sy, float sx, cudaStream_t stream) {
  const int warps_x = (w + 30) / 31;
  const int wpb = warps_x < 8 ? warps_x : 8;
  const int gx = (warps_x + wpb - 1) / wpb;
  int strip = h >= 128 ? 64 : 32;
  if (strip > h) strip = h;
  const int gy = (h + strip - 1) / strip;
  int64_t gz = ((int64_t)sm_count() * 16 + (int64_t)gx * gy - 1) / ((int64_t)gx * gy);
  gz = plane_groups(gz, planes);
  // rows per strip: (strip + 1) owners, each with <= ceil(1/sy) + 1 rows, plus the clamped rows
  const size_t rows_max = (size_t)((strip + 2) * (1.0 / sy + 1.0) + 8);
  const size_t smem = rows_max * sizeof(RowTap);
  if (smem > 160 * 1024) return -1;  // caller falls back
  auto kern = upsample_bwd_walk_kernel<RMAX>;
  if (smem > 48 * 1024)
    ROBSEG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kern<<<dim3(gx, gy, (unsigned)gz), 32 * wpb, smem, stream>>>(gout, gin, planes, C, bs, cs, h, w, H, W,
                                                              sy, sx, strip);
  ROBSEG_LAUNCH_CHECK();
  return 0;
}

}  // namespace robseg

using namespace robseg;

template <int R>
static int launch_fwd_pow2(const float* in, float* out, int64_t planes, int h, int w, cudaStream_t stream) {
  const int gx = (h * w * (R > 4 ? R / 4 : 1) + 127) / 128;  // R / 4 threads per cell beyond x4
  // ~32 resident blocks per SM worth of cell tiles; the rest of the parallelism is planes
  int64_t gz = ((int64_t)sm_count() * 32 + gx - 1) / gx;
  gz = plane_groups(gz, planes);
  upsample_fwd_pow2_kernel<R><<<dim3(gx, 1, (unsigned)gz), 128, 0, stream>>>(in, out, planes, h, w);
  ROBSEG_LAUNCH_CHECK();
  return 0;
}

static int launch_fwd_x2(const float* in, float* out, int64_t planes, int h, int w, cudaStream_t stream) {
  const int cells = (h >> 1) * (w >> 1);
  const int gx = (cells + 127) / 128;
  // 72 registers -> 7 resident blocks per SM: two full waves of blocks, the rest of the planes in the block's loop
  int64_t gz = ((int64_t)sm_count() * 14 + gx - 1) / gx;
  gz = plane_groups(gz, planes);
  upsample_fwd_x2_kernel<<<dim3(gx, 1, (unsigned)gz), 128, 0, stream>>>(in, out, planes, h, w);
  ROBSEG_LAUNCH_CHECK();
  return 0;
}

static int launch_bwd_x2(const float* gout, float* gin, int64_t planes, int C, int64_t bs, int64_t cs, int h,
                         int w, cudaStream_t stream) {
  const int cells = (h >> 1) * (w >> 1);
  const int gx = (cells + 127) / 128;
  int64_t gz = ((int64_t)sm_count() * 12 + gx - 1) / gx;  // two waves of 6 resident blocks per SM
  gz = plane_groups(gz, planes);
  upsample_bwd_x2_kernel<<<dim3(gx, 1, (unsigned)gz), 128, 0, stream>>>(gout, gin, planes, C, bs, cs, h, w);
  ROBSEG_LAUNCH_CHECK();
  return 0;
}

template <int R>
static int launch_bwd_pow2(const float* gout, float* gin, int64_t planes, int C, int64_t bs, int64_t cs,
                           int h, int w, cudaStream_t stream) {
  // cells per warp: spread the columns evenly over the fewest warps (<= 30 owned + 2 halo lanes)
  const int halo = (w <= 32 && !getenv("ROBSEG_UP_BWD_HALO")) ? 0 : 1;  // a warp that spans the plane needs no halo lanes
  const int warps_x = halo ? (w + 29) / 30 : 1;
  const int owned = (w + warps_x - 1) / warps_x;
  const int wpb = warps_x < 8 ? warps_x : 8;  // warps per block (exact cover when <= 8)
  const int gx = (warps_x + wpb - 1) / wpb;
  // cell rows per thread strip (2 halo block rows per strip): 64 / 32 when there are planes enough to fill the
  // GPU; with few planes the serial walk down a long strip is the critical path, so strips are halved until there
  // are >= 8 tiles per SM (2x21x128^2 cells, the configs[0] logits: 49 -> 18 us at strip 8).  Plane groups: up to
  // 64 blocks per SM worth of tiles, i.e. one plane per block up to ~9500 tiles -- the block loop over planes
  // buys nothing here (the y-weight table is cheap) and its uneven trip counts cost 8-25 % (profiles/
  // r02_kernel_brackets_and_trainer.md section 5).
  int strip = h >= 128 ? 64 : 32;
  int bps = 64;
  if (strip > h) strip = h;
  while (strip > 8 && (int64_t)gx * ((h + strip - 1) / strip) * planes < (int64_t)8 * sm_count()) strip >>= 1;
  if (const char* e = getenv("ROBSEG_UP_BWD_STRIP")) strip = atoi(e) > 0 ? atoi(e) : strip;
  if (const char* e = getenv("ROBSEG_UP_BWD_BPS")) bps = atoi(e) > 0 ? atoi(e) : bps;
  if (strip > h) strip = h;
  const int gy = (h + strip - 1) / strip;
  int64_t gz = ((int64_t)sm_count() * bps + (int64_t)gx * gy - 1) / ((int64_t)gx * gy);
  gz = plane_groups(gz, planes);
  const size_t ky_bytes = (size_t)(strip + 2) * ((3 * R + 3) & ~3) * sizeof(float);
  upsample_bwd_pow2_kernel<R><<<dim3(gx, gy, (unsigned)gz), 32 * wpb, ky_bytes, stream>>>(
      gout, gin, planes, C, bs, cs, h, w, strip, owned, halo);
  ROBSEG_LAUNCH_CHECK();
  return 0;
}

extern "C" int robseg_upsample_bilinear_fwd(const float* in, int64_t planes, int h, int w, float* out,
                                            int H, int W, robseg_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  ROBSEG_REQUIRE(in && out, "NULL pointer");
  ROBSEG_REQUIRE(planes > 0 && h > 0 && w > 0 && H > 0 && W > 0, "bad shape");
  const float sy = (float)h / (float)H, sx = (float)w / (float)W;
  if (H % h == 0 && W % w == 0 && H / h == W / w && reinterpret_cast<uintptr_t>(out) % 16 == 0 &&
      (int64_t)h * w < ((int64_t)1 << 30)) {
    switch (H / h) {
      case 2:
        if (h % 2 == 0 && w % 2 == 0 && reinterpret_cast<uintptr_t>(in) % 8 == 0 && !getenv("ROBSEG_UP_X2_CELL"))
          return launch_fwd_x2(in, out, planes, h, w, stream);
        return launch_fwd_pow2<2>(in, out, planes, h, w, stream);
      case 4: return launch_fwd_pow2<4>(in, out, planes, h, w, stream);
      // (planes of a few cells -- the head's 1x1 / 2x2 PSP pools -> 16x16 -- would leave most of a 128-thread
      // block idle: they take the generic kernel below, whose threads are quads of the OUTPUT plane)
      case 8:
        if (h * w >= 32) return launch_fwd_pow2<8>(in, out, planes, h, w, stream);
        break;
      case 16:  // SegMenter's class masks (segmenter.py:228); ROBSEG_UP_FWD_QUAD=1: the generic quad kernel
        if (h * w >= 32 && !getenv("ROBSEG_UP_FWD_QUAD")) return launch_fwd_pow2<16>(in, out, planes, h, w, stream);
        break;
      default: break;
    }
  }
  if ((W % 4 != 0 || reinterpret_cast<uintptr_t>(out) % 16 != 0) && H >= h && W >= w && !getenv("ROBSEG_UP_FWD_QUAD")) {
    // rows that cannot be written as aligned 16-byte vectors: one output column per lane, walking down a strip
    if ((int64_t)h * w < ((int64_t)1 << 30)) return launch_fwd_walk(in, out, planes, h, w, H, W, sy, sx, stream);
  }
  const int64_t quads = (int64_t)H * ((W + 3) / 4);
  ROBSEG_REQUIRE(quads < ((int64_t)1 << 30), "plane too large");
  const int gx = (int)((quads + 127) / 128);
  // ~32 resident blocks per SM worth of quads; the rest of the parallelism is planes
  int64_t gz = ((int64_t)sm_count() * 32 + gx - 1) / gx;
  gz = plane_groups(gz, planes);
  dim3 grid((unsigned)gx, 1, (unsigned)gz);
  upsample_fwd_kernel<<<grid, 128, 0, stream>>>(in, out, planes, h, w, H, W, sy, sx);
  ROBSEG_LAUNCH_CHECK();
  return 0;
}

extern "C" int robseg_upsample_bilinear_bwd_strided(const float* gout, int64_t N, int C,
                                                    int64_t batch_stride, int64_t chan_stride, int H,
                                                    int W, float* gin, int h, int w,
                                                    robseg_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  ROBSEG_REQUIRE(gout && gin, "NULL pointer");
  ROBSEG_REQUIRE(N > 0 && C > 0 && h > 0 && w > 0 && H > 0 && W > 0, "bad shape");
  ROBSEG_REQUIRE(batch_stride >= 0 && chan_stride >= (int64_t)H * W, "planes overlap");
  const int64_t planes = N * C, bs = batch_stride, cs = chan_stride;
  const float sy = (float)h / (float)H, sx = (float)w / (float)W;
  if (H % h == 0 && W % w == 0 && H / h == W / w && reinterpret_cast<uintptr_t>(gout) % 16 == 0 &&
      bs % 4 == 0 && cs % 4 == 0) {
    switch (H / h) {
      case 2:
        if (h % 2 == 0 && w % 2 == 0 && reinterpret_cast<uintptr_t>(gin) % 8 == 0 && !getenv("ROBSEG_UP_X2_CELL"))
          return launch_bwd_x2(gout, gin, planes, C, bs, cs, h, w, stream);
        return launch_bwd_pow2<2>(gout, gin, planes, C, bs, cs, h, w, stream);
      case 4: return launch_bwd_pow2<4>(gout, gin, planes, C, bs, cs, h, w, stream);
      case 8: return launch_bwd_pow2<8>(gout, gin, planes, C, bs, cs, h, w, stream);
      case 16: return launch_bwd_pow2<16>(gout, gin, planes, C, bs, cs, h, w, stream);
      default: break;
    }
  }
  if (H >= h && W >= w) {
    // walk-down gather for any up-sampling ratio whose per-cell output count fits the registers
    const int per_cell = (int)ceilf((float)W / (float)w) + 1;
    int rc = -1;
    if (per_cell <= 3) rc = launch_bwd_walk<3>(gout, gin, planes, C, bs, cs, h, w, H, W, sy, sx, stream);
    else if (per_cell <= 6) rc = launch_bwd_walk<6>(gout, gin, planes, C, bs, cs, h, w, H, W, sy, sx, stream);
    else if (per_cell <= 10) rc = launch_bwd_walk<10>(gout, gin, planes, C, bs, cs, h, w, H, W, sy, sx, stream);
    if (rc >= 0) return rc;
  }
  // anything else: tile of input cells per CTA, shrunk until the staged output region fits ~48 KB
  int TY = 8, TX = 32;
  if (TX > w) TX = w;
  if (TY > h) TY = h;
  int RY, RX, KY, KX;
  size_t smem;
  for (;;) {
    KY = (int)ceilf(2.f / sy) + 4, KX = (int)ceilf(2.f / sx) + 4;
    RY = (int)ceilf((TY + 2) / sy) + 4 + KY;
    RX = (((int)ceilf((TX + 2) / sx) + 8 + KX + 3) / 4) * 4;
    smem = ((size_t)RY * (RX | 1) + (size_t)TX * (RY | 1) + (size_t)TX * KX + (size_t)TY * KY + TX + TY) *
           sizeof(float);
    if (smem <= 48 * 1024 || (TY == 1 && TX == 1)) break;
    if (TX > 1 && (TX >= TY * 4 || TY == 1)) TX = (TX + 1) / 2; else TY = (TY + 1) / 2;
  }
  ROBSEG_REQUIRE(smem <= 200 * 1024, "up-sampling ratio too large (%dx%d -> %dx%d)", h, w, H, W);
  ROBSEG_CUDA(cudaFuncSetAttribute(upsample_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   (int)smem));
  const int gx = (w + TX - 1) / TX, gy = (h + TY - 1) / TY;
  int64_t gz = ((int64_t)sm_count() * 16 + (int64_t)gx * gy - 1) / ((int64_t)gx * gy);
  gz = plane_groups(gz, planes);
  dim3 grid((unsigned)gx, (unsigned)gy, (unsigned)gz);
  upsample_bwd_kernel<<<grid, 256, smem, stream>>>(gout, gin, planes, C, bs, cs, h, w, H, W, sy, sx, TY,
                                                   TX, RY, RX, KY, KX);
  ROBSEG_LAUNCH_CHECK();
  return 0;
}

extern "C" int robseg_upsample_bilinear_bwd(const float* gout, int64_t planes, int H, int W, float* gin,
                                            int h, int w, robseg_stream_t stream_) {
  ROBSEG_REQUIRE(planes > 0 && planes <= 0x7fffffff && H > 0 && W > 0, "bad shape");
  return robseg_upsample_bilinear_bwd_strided(gout, 1, (int)planes, 0, (int64_t)H * W, H, W, gin, h, w,
                                              stream_);
}
