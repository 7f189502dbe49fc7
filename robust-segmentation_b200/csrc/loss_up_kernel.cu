// Fused bilinear up-sampling (x R, align_corners=False) + per-pixel softmax + SEA loss + the gradient
// with respect to the LOW-RESOLUTION logits -- SURVEY.md section 8f rank 1.
//
// The reference's models end in
//     logits = nn.functional.interpolate(low, size=input.shape[2:], mode="bilinear", align_corners=False)
// (semseg/models/uperforseg.py:416-418, R = 4; semseg/models/segmenter.py:228, R = 16) and every attack
// iteration then makes ~7-25 passes over that [B,C,H,W] tensor (semseg/attacker.py:143-240).  Here the
// full-resolution logits and their gradient never exist: a warp interpolates its pixels' logits on the
// fly from a shared-memory table of the four surrounding low-resolution cells, makes the same three passes
// as loss_tma_kernel (max -> sum-exp + first-max index -> gradient) and reduces the gradient straight
// into the low-resolution cells.  HBM traffic per launch: the labels (8 B/pixel) plus the R^2-times
// smaller low-resolution tensors; the kernel is instruction-bound (ex2 + FMA), not memory-bound.
//
// Geometry.  With align_corners=False and an integer ratio R, output row Y reads low rows
// k = floor((Y - R/2) / R) and k+1 with weight ly = ((Y - R/2) mod R + 0.5) / R (ATen's
// area_pixel_compute_source_index; rows above R/2 / below H - R/2 clamp to the border row).  The R x R
// output pixels between four low-resolution cell centres -- a *dual cell* (k, l), k in [-1, h-1],
// l in [-1, w-1] -- share their four taps.  A warp tile is one dual row (R output rows) x 32 output
// columns = 32/R dual cells; lane = column, so the column weight lx is a per-lane constant, the row
// weights are compile-time constants, and one 16-byte shared-memory read per lane and channel fetches
// the taps (a, b-a, c, d-c):  z(row) = top + ly(row) * (bot - top).
//
// Gradient.  d low[cell] = sum over the pixels that read the cell of weight * dz.  Per channel a lane
// accumulates (sum_rows g, sum_rows ly*g) for its column; every 32 channels the warp transposes through
// shared memory (lane = channel) and finishes the column sums, giving the four corner contributions of
// each dual cell, written to a workspace [B][h+1][C][w+1][4].  A small gather kernel then adds the (up
// to four) contributions of every low-resolution cell in a fixed order: no atomics, bit-reproducible.
#include "common.cuh"

namespace robseg {

int launch_loss_finalize(const float4* partials, int B, int tiles_per_img, const float* grad_scale,
                         int64_t HW, float* loss_img, float* track_img, int32_t* correct_img,
                         int32_t* valid_img, cudaStream_t stream);  // loss_kernel.cu

struct LossUpParams {
  const float* low;       // [B][C][h][w]
  const int64_t* labels;  // [B][H][W]
  const float* class_w;
  const float* grad_scale;
  float4* contrib;  // [B][h+1][C][w+1] corner contributions (a,b,c,d), or nullptr (no gradient)
  int64_t* pred;
  float4* partials;
  unsigned long long* counts;  // [n_rep][B][3][C] inter / tgt / prd (zeroed by the launcher), or nullptr
  int kind, ignore_index, B, C, h, w, H, W, n_rep;
  int groups_x, tiles_per_img, num_tiles;
  float inv_hw;
};

constexpr float kLog2eU = 1.4426950408889634f;
constexpr float kLn2U = 0.6931471805599453f;
constexpr int kChunk = 32;        // channels per transpose round
constexpr int kChunkStride = 33;  // float2 per channel row (+1: lane = channel reads stay conflict-free)

__device__ __forceinline__ float fmax3u(float a, float b, float c) {
  float r;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
  return r;
}

template <int R>
__device__ __forceinline__ constexpr float row_w(int r) {
  return (r + 0.5f) / R;
}

template <int R>
__global__ void __launch_bounds__(256, 1) loss_up_kernel(const LossUpParams p) {
  constexpr int NDC = 32 / R;  // dual cells per warp tile
  constexpr int CU = R >= 16 ? 1 : (R == 8 ? 2 : 4);  // channels per loop trip (independent chains)
  extern __shared__ __align__(16) uint8_t smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, NW = blockDim.x >> 5;
  const int C = p.C;
  const size_t per_warp = (size_t)C * NDC * sizeof(float4) + (size_t)kChunk * kChunkStride * sizeof(float2);
  float4* tab = reinterpret_cast<float4*>(smem_raw + warp * per_warp);
  float2* buf = reinterpret_cast<float2*>(tab + (size_t)C * NDC);
  const int dc = lane / R;
  const bool argmax_only = p.kind == ROBSEG_LOSS_ARGMAX;
  // class counters: a tile is counted while the NEXT one is processed, so the global reductions are long complete
  // at the next warp barrier (issued right before it they cost ~25 % of a launch: the barrier waits for them)
  int cnt_t[R], cnt_q[R], cnt_b = 0;
  bool cnt_pending = false;
  auto count_prev = [&]() {
    unsigned long long* cnt = p.counts + ((size_t)(blockIdx.x & (p.n_rep - 1)) * p.B + cnt_b) * 3 * C;
    count_pixels<R>(cnt, C, cnt_t, cnt_q);
  };

  for (int tile = blockIdx.x * NW + warp; tile < p.num_tiles; tile += gridDim.x * NW) {
    const int b = tile / p.tiles_per_img;
    const int t = tile - b * p.tiles_per_img;
    const int kd = t / p.groups_x, j = t - kd * p.groups_x;  // dual row index k+1, column group
    const int k = kd - 1;
    const int y0 = min(max(k, 0), p.h - 1), y1 = min(max(k + 1, 0), p.h - 1);
    const float yscale = (k >= 0 && k < p.h - 1) ? 1.f : 0.f;
    const int l_first = (32 * j) / R - 1;  // dual column of lane 0
    const int Xp = 32 * j + lane, X = Xp - R / 2;
    const int l = l_first + dc;
    const bool lane_ok = X >= 0 && X < p.W;
    const float lx = (l >= 0 && l < p.w - 1) ? ((Xp % R) + 0.5f) / R : 0.f;
    const int Y0 = R * kd - R / 2;  // output row of r = 0 (may be negative / beyond H at the borders)

    // ---- taps of this tile's dual cells -> shared memory, as (a, b-a, c, d-c) ---------------------
    {
      const float* base = p.low + (size_t)b * C * p.h * p.w;
      const int n_items = C * NDC;
      constexpr int FU = 8;  // items (32 scalar loads) in flight per lane
      for (int i0 = 0; i0 < n_items; i0 += 32 * FU) {
        float va[FU], vb[FU], vc[FU], vd[FU];
#pragma unroll
        for (int u = 0; u < FU; ++u) {
          const int i = i0 + u * 32 + lane;
          if (i < n_items) {
            const int c = i / NDC, d = i - c * NDC;
            const int ll = l_first + d;
            const int x0 = min(max(ll, 0), p.w - 1), x1 = min(max(ll + 1, 0), p.w - 1);
            const float* pl = base + (size_t)c * p.h * p.w;
            va[u] = __ldg(pl + y0 * p.w + x0), vb[u] = __ldg(pl + y0 * p.w + x1);
            vc[u] = __ldg(pl + y1 * p.w + x0), vd[u] = __ldg(pl + y1 * p.w + x1);
          }
        }
#pragma unroll
        for (int u = 0; u < FU; ++u) {
          const int i = i0 + u * 32 + lane;
          if (i < n_items) tab[i] = make_float4(va[u], vb[u] - va[u], vc[u], vd[u] - vc[u]);
        }
      }
    }
    // labels: one coalesced row read per output row; rows / columns outside the image count as ignored
    int y[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const int Y = Y0 + r;
      y[r] = (lane_ok && Y >= 0 && Y < p.H) ? (int)__ldg(p.labels + ((size_t)b * p.H + Y) * p.W + X)
                                            : p.ignore_index;
    }
    __syncwarp();
    if (cnt_pending) count_prev();
    const float4* col = tab + dc;

    // ---- pass 1: channel maximum (ARGMAX launches also track the index: strict >, ascending) -------
    float m[R];
    int amx[R];
#pragma unroll
    for (int r = 0; r < R; ++r) m[r] = -INFINITY, amx[r] = 0;
    if (argmax_only) {
#pragma unroll 2
      for (int c = 0; c < C; ++c) {
        const float4 q = col[c * NDC];
        const float top = fmaf(lx, q.y, q.x), dif = fmaf(lx, q.w, q.z) - top;
#pragma unroll
        for (int r = 0; r < R; ++r) {
          const float z = fmaf(row_w<R>(r), dif, top);
          const bool g = z > m[r];
          m[r] = g ? z : m[r], amx[r] = g ? c : amx[r];
        }
      }
    } else {
      // few rows per lane (small R) leave few independent chains: unroll CU channel pairs per trip with
      // separate accumulators (ncu, x4: 44 % issue utilisation with one pair per trip, stalled on the
      // 4-cycle dependent-issue wait and the LDS / MUFU scoreboards)
      float m2[R];
#pragma unroll
      for (int r = 0; r < R; ++r) m2[r] = -INFINITY;
      int c = 0;
#pragma unroll 1
      for (; c + 2 * CU <= C; c += 2 * CU) {
#pragma unroll
        for (int u = 0; u < CU; ++u) {
          const float4 q0 = col[(c + 2 * u) * NDC], q1 = col[(c + 2 * u + 1) * NDC];
          const float top0 = fmaf(lx, q0.y, q0.x), dif0 = fmaf(lx, q0.w, q0.z) - top0;
          const float top1 = fmaf(lx, q1.y, q1.x), dif1 = fmaf(lx, q1.w, q1.z) - top1;
#pragma unroll
          for (int r = 0; r < R; ++r) {
            const float z0 = fmaf(row_w<R>(r), dif0, top0), z1 = fmaf(row_w<R>(r), dif1, top1);
            if (u & 1) m2[r] = fmax3u(m2[r], z0, z1);
            else m[r] = fmax3u(m[r], z0, z1);
          }
        }
      }
#pragma unroll 1
      for (; c < C; ++c) {
        const float4 q = col[c * NDC];
        const float top = fmaf(lx, q.y, q.x), dif = fmaf(lx, q.w, q.z) - top;
#pragma unroll
        for (int r = 0; r < R; ++r) m[r] = fmaxf(m[r], fmaf(row_w<R>(r), dif, top));
      }
#pragma unroll
      for (int r = 0; r < R; ++r) m[r] = fmaxf(m[r], m2[r]);
    }

    float loss_sum = 0.f, ce_sum = 0.f;
    int n_correct = 0, n_valid = 0;
    if (argmax_only) {
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const bool valid = (y[r] != p.ignore_index) && (y[r] >= 0) && (y[r] < C);
        n_valid += valid, n_correct += valid && (amx[r] == y[r]);
      }
    } else {
      // ---- pass 2: sum of exp(z - max) + first maximal channel (equality, walking downwards) -------
      constexpr int kNoIdx = 0x7fffffff;
      float mL[R], s[R];
      {
        float su[CU][R];
        int au[CU][R];
#pragma unroll
        for (int r = 0; r < R; ++r) mL[r] = m[r] * kLog2eU;
#pragma unroll
        for (int u = 0; u < CU; ++u)
#pragma unroll
          for (int r = 0; r < R; ++r) su[u][r] = 0.f, au[u][r] = kNoIdx;
        int c = C;
#pragma unroll 1
        for (; c - CU >= 0; c -= CU) {
#pragma unroll
          for (int u = 0; u < CU; ++u) {
            const int cu = c - 1 - u;
            const float4 q = col[cu * NDC];
            const float top = fmaf(lx, q.y, q.x), dif = fmaf(lx, q.w, q.z) - top;
#pragma unroll
            for (int r = 0; r < R; ++r) {
              const float z = fmaf(row_w<R>(r), dif, top);
              su[u][r] += ex2_approx(fmaf(z, kLog2eU, -mL[r]));
              au[u][r] = (z == m[r]) ? cu : au[u][r];  // walking downwards: the last hit is the lowest
            }
          }
        }
#pragma unroll 1
        for (; c > 0; --c) {
          const float4 q = col[(c - 1) * NDC];
          const float top = fmaf(lx, q.y, q.x), dif = fmaf(lx, q.w, q.z) - top;
#pragma unroll
          for (int r = 0; r < R; ++r) {
            const float z = fmaf(row_w<R>(r), dif, top);
            su[0][r] += ex2_approx(fmaf(z, kLog2eU, -mL[r]));
            au[0][r] = (z == m[r]) ? (c - 1) : au[0][r];
          }
        }
#pragma unroll
        for (int r = 0; r < R; ++r) {
          s[r] = su[0][r], amx[r] = au[0][r];
#pragma unroll
          for (int u = 1; u < CU; ++u) s[r] += su[u][r], amx[r] = min(amx[r], au[u][r]);
          amx[r] = amx[r] == kNoIdx ? 0 : amx[r];  // (only NaN logits have no channel equal to the maximum)
        }
      }
      // ---- per-pixel loss terms (same formulas as loss_kernel.cu) ---------------
      // kf = coef * g_b / sumexp multiplies exp(z - max); sub = coef * g_b is taken off the label channel
      float kf[R], kfl[R], sub[R];
      int ys[R];
      const float g_img = p.grad_scale ? __ldg(p.grad_scale + b) : p.inv_hw;
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const bool valid = (y[r] != p.ignore_index) && (y[r] >= 0) && (y[r] < C);
        const bool hit = valid && (amx[r] == y[r]);
        ys[r] = valid ? y[r] : -1;
        n_valid += valid, n_correct += hit;
        const float4 q = col[(valid ? y[r] : 0) * NDC];
        const float top = fmaf(lx, q.y, q.x), dif = fmaf(lx, q.w, q.z) - top;
        const float zy = fmaf(row_w<R>(r), dif, top);
        const float resid = fmaf(m[r], kLog2eU, -mL[r]);
        const float ln_s = logf(s[r]) - resid * kLn2U;  // lse - max
        const float logp = (zy - m[r]) - ln_s;
        const float ce = valid ? -logp : 0.f;
        float lo, coef;
        if (p.kind == ROBSEG_LOSS_CE) {
          lo = ce, coef = valid ? 1.f : 0.f;
        } else if (p.kind == ROBSEG_LOSS_MASK_CE) {
          lo = hit ? ce : 0.f, coef = hit ? 1.f : 0.f;
        } else if (p.kind == ROBSEG_LOSS_MASK_CE_BAL) {
          const float wgt = (p.class_w != nullptr && valid) ? __ldg(p.class_w + y[r]) : 1.f;
          lo = hit ? wgt * ce : 0.f, coef = hit ? wgt : 0.f;
        } else {  // JS(softmax || one-hot)
          const float py = expf(logp);
          const float l1p = log1pf(py);
          lo = valid ? 0.5f * (2.f * kLn2U + py * logp - (1.f + py) * l1p) : 0.f;
          coef = valid ? -0.5f * py * (logp - l1p) : 0.f;
        }
        loss_sum += lo, ce_sum += ce;
        const float cg = coef * g_img;
        sub[r] = cg;
        kf[r] = cg / s[r];
        kfl[r] = kf[r] * row_w<R>(r);
      }

      // ---- pass 3: gradient, reduced into the four corners of every dual cell ----------------------
      if (p.contrib != nullptr) {
        float4* out = p.contrib + (((size_t)b * (p.h + 1) + kd) * C) * (p.w + 1) + (l_first + 1);
        for (int c0 = 0; c0 < C; c0 += kChunk) {
          const int nc = min(kChunk, C - c0);
#pragma unroll CU
          for (int cc = 0; cc < nc; ++cc) {
            const float4 q = col[(c0 + cc) * NDC];
            const float top = fmaf(lx, q.y, q.x), dif = fmaf(lx, q.w, q.z) - top;
            float all = 0.f, bt = 0.f;
#pragma unroll
            for (int r = 0; r < R; ++r) {
              const float z = fmaf(row_w<R>(r), dif, top);
              const float e = ex2_approx(fmaf(z, kLog2eU, -mL[r]));
              all = fmaf(e, kf[r], all);
              bt = fmaf(e, kfl[r], bt);
            }
            buf[cc * kChunkStride + lane] = make_float2(all, bt);
          }
          // the label channel carries coef*(p_y - 1): take coef*g_b off where it falls in this chunk
          // (a lane only ever touches its own column of buf)
#pragma unroll
          for (int r = 0; r < R; ++r) {
            const int cc = ys[r] - c0;
            if (cc >= 0 && cc < nc && sub[r] != 0.f) {
              float2 v = buf[cc * kChunkStride + lane];
              v.x -= sub[r], v.y -= sub[r] * row_w<R>(r);
              buf[cc * kChunkStride + lane] = v;
            }
          }
          __syncwarp();
          if (lane < nc) {  // lane = channel: finish the column sums of each dual cell
            const float2* rowp = buf + lane * kChunkStride;
#pragma unroll
            for (int d = 0; d < NDC; ++d) {
              float sa = 0.f, sla = 0.f, sb = 0.f, slb = 0.f;
#pragma unroll
              for (int i = 0; i < R; ++i) {
                const float2 v = rowp[d * R + i];
                sa += v.x, sla = fmaf(row_w<R>(i), v.x, sla);
                sb += v.y, slb = fmaf(row_w<R>(i), v.y, slb);
              }
              const int ll = l_first + d;
              if (ll <= p.w - 1) {
                const float xs = (ll >= 0 && ll < p.w - 1) ? 1.f : 0.f;
                sla *= xs, slb *= xs;
                sb *= yscale, slb *= yscale;
                out[(size_t)(c0 + lane) * (p.w + 1) + d] =
                    make_float4((sa - sla) - (sb - slb), sla - slb, sb - slb, slb);
              }
            }
          }
          __syncwarp();
        }
      }
    }

    if (p.counts != nullptr) {  // warp-uniform; pixels outside the image carry y = ignore_index
#pragma unroll
      for (int r = 0; r < R; ++r)
        cnt_t[r] = ((y[r] != p.ignore_index) && (y[r] >= 0) && (y[r] < C)) ? y[r] : -1, cnt_q[r] = amx[r];
      cnt_b = b, cnt_pending = true;
    }
    if (p.pred != nullptr && lane_ok) {
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const int Y = Y0 + r;
        if (Y >= 0 && Y < p.H) p.pred[((size_t)b * p.H + Y) * p.W + X] = amx[r];
      }
    }
    loss_sum = warp_sum(loss_sum);
    ce_sum = warp_sum(ce_sum);
    n_correct = warp_sum(n_correct);
    n_valid = warp_sum(n_valid);
    if (lane == 0)
      p.partials[tile] = make_float4(loss_sum, ce_sum, __int_as_float(n_correct), __int_as_float(n_valid));
    __syncwarp();  // the next tile rewrites tab
  }
  if (cnt_pending) count_prev();
}

// dlow[b][c][y][x] = the corner contributions that target this cell, added in a fixed order.
// Dual cell (kd, ld) = (k+1, l+1): corner a -> (clamp k, clamp l), b -> (clamp k, l+1), c -> (k+1, clamp l),
// d -> (k+1, l+1); b / d are zero at the column borders, c / d at the row borders.
__global__ void __launch_bounds__(256) loss_up_gather_kernel(const float4* __restrict__ contrib, int C, int h,
                                                             int w, int64_t total, float* __restrict__ dlow) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int x = (int)(i % w);
    const int64_t r0 = i / w;
    const int y = (int)(r0 % h);
    const int64_t bc = r0 / h;
    const int c = (int)(bc % C);
    const int64_t b = bc / C;
    auto at = [&](int kd, int ld) { return __ldg(contrib + (((size_t)b * (h + 1) + kd) * C + c) * (w + 1) + ld); };
    // dual rows whose clamped k equals y: kd = y+1, plus kd = 0 when y == 0;  same for columns
    float acc = 0.f;
    const float4 q11 = at(y + 1, x + 1);
    acc += q11.x;
    if (x == 0) acc += at(y + 1, 0).x;
    if (y == 0) acc += at(0, x + 1).x;
    if (x == 0 && y == 0) acc += at(0, 0).x;
    if (x >= 1) {  // corner b of the dual column to the left
      acc += at(y + 1, x).y;
      if (y == 0) acc += at(0, x).y;
    }
    if (y >= 1) {  // corner c of the dual row above
      acc += at(y, x + 1).z;
      if (x == 0) acc += at(y, 0).z;
    }
    if (x >= 1 && y >= 1) acc += at(y, x).w;
    dlow[i] = acc;
  }
}

static size_t contrib_bytes(int B, int C, int h, int w) {
  return (size_t)B * (h + 1) * C * (w + 1) * sizeof(float4);
}
static int groups_x_of(int W, int R) { return (W + R / 2 + 31) / 32; }
static size_t partial_bytes(int B, int h, int W, int R) {
  return ((size_t)B * (h + 1) * groups_x_of(W, R) * sizeof(float4) + 255) / 256 * 256;
}

template <int R>
static int launch_up(LossUpParams p, cudaStream_t stream) {
  constexpr int NDC = 32 / R;
  p.groups_x = groups_x_of(p.W, R);
  p.tiles_per_img = (p.h + 1) * p.groups_x;
  p.num_tiles = p.B * p.tiles_per_img;
  const size_t per_warp = (size_t)p.C * NDC * sizeof(float4) + (size_t)kChunk * kChunkStride * sizeof(float2);
  int warps = (int)((227 * 1024) / per_warp);
  if (warps > 8) warps = 8;
  ROBSEG_REQUIRE(warps >= 1, "C=%d too large for the fused up-sampling loss kernel", p.C);
  const size_t smem = warps * per_warp;
  auto kern = loss_up_kernel<R>;
  ROBSEG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int grid = sm_count();
  const int max_useful = (p.num_tiles + warps - 1) / warps;
  if (grid > max_useful) grid = max_useful;
  kern<<<grid, 32 * warps, smem, stream>>>(p);
  ROBSEG_LAUNCH_CHECK();
  return 0;
}

}  // namespace robseg

using namespace robseg;

static size_t up_replica_bytes(int B, int C) { return (size_t)count_replicas(C) * B * 3 * C * sizeof(int64_t); }
extern "C" size_t robseg_loss_upsampled_workspace_bytes(int B, int C, int h, int w, int H, int W) {
  if (B <= 0 || C <= 0 || h <= 0 || w <= 0 || H % h != 0) return 0;
  const int R = H / h;
  return partial_bytes(B, h, W, R) + contrib_bytes(B, C, h, w) + up_replica_bytes(B, C);  // partials | corners | counters
}

static int loss_upsampled_impl(const float* low, const int64_t* labels, const float* class_w,
                               int loss_kind, int ignore_index, int B, int C, int h, int w,
                               int H, int W, const float* grad_scale, float* dlow,
                               int64_t* pred, float* loss_img, float* track_img,
                               int32_t* correct_img, int32_t* valid_img, int64_t* counts, void* workspace,
                               size_t workspace_bytes, robseg_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  // one-shot measurement events (robseg_profile_next_kernel): consumed by this call whatever its outcome
  const cudaEvent_t prof_start = take_profile_start(), prof_stop = take_profile_stop();
  ROBSEG_REQUIRE(low && labels, "low/labels must not be NULL");
  ROBSEG_REQUIRE(loss_kind >= ROBSEG_LOSS_CE && loss_kind <= ROBSEG_LOSS_ARGMAX, "unknown loss kind %d", loss_kind);
  ROBSEG_REQUIRE(B > 0 && C > 0 && h > 0 && w > 0, "bad shape B=%d C=%d h=%d w=%d", B, C, h, w);
  ROBSEG_REQUIRE(H % h == 0 && W % w == 0 && H / h == W / w, "output %dx%d is not an integer multiple of %dx%d", H, W, h, w);
  const int R = H / h;
  ROBSEG_REQUIRE(R == 2 || R == 4 || R == 8 || R == 16, "up-sampling ratio %d not in {2,4,8,16}", R);
  ROBSEG_REQUIRE(workspace && workspace_bytes >= robseg_loss_upsampled_workspace_bytes(B, C, h, w, H, W),
                 "workspace too small (%zu bytes)", workspace_bytes);
  ROBSEG_REQUIRE(reinterpret_cast<uintptr_t>(workspace) % 16 == 0, "workspace must be 16B aligned");
  ROBSEG_REQUIRE((int64_t)B * (h + 1) * groups_x_of(W, R) < (1ll << 31), "too many tiles");

  LossUpParams p{};
  p.low = low, p.labels = labels, p.class_w = class_w, p.grad_scale = grad_scale, p.pred = pred;
  p.partials = static_cast<float4*>(workspace);
  const bool want_grad = dlow != nullptr && loss_kind != ROBSEG_LOSS_ARGMAX;
  p.contrib = want_grad ? reinterpret_cast<float4*>(static_cast<uint8_t*>(workspace) + partial_bytes(B, h, W, R))
                        : nullptr;
  p.kind = loss_kind, p.ignore_index = ignore_index, p.B = B, p.C = C, p.h = h, p.w = w, p.H = H, p.W = W;
  p.inv_hw = (float)(1.0 / ((double)H * W));
  if (counts != nullptr) {
    p.counts = reinterpret_cast<unsigned long long*>(static_cast<uint8_t*>(workspace) + partial_bytes(B, h, W, R) +
                                                     contrib_bytes(B, C, h, w));
    p.n_rep = count_replicas(C);
    const int zrc = launch_counts_zero(p.counts, up_replica_bytes(B, C), stream);
    if (zrc != 0) return zrc;
  }
  if (prof_start) cudaEventRecord(prof_start, stream);
  int rc = R == 16 ? launch_up<16>(p, stream) : R == 8 ? launch_up<8>(p, stream)
           : R == 4 ? launch_up<4>(p, stream) : launch_up<2>(p, stream);
  if (prof_stop) cudaEventRecord(prof_stop, stream);
  if (rc != 0) return rc;
  if (counts != nullptr) {
    rc = launch_counts_fold(p.counts, p.n_rep, B, C, counts, stream);
    if (rc != 0) return rc;
  }
  if (want_grad) {
    const int64_t total = (int64_t)B * C * h * w;
    int grid = (int)((total + 255) / 256);
    if (grid > sm_count() * 16) grid = sm_count() * 16;
    loss_up_gather_kernel<<<grid, 256, 0, stream>>>(p.contrib, C, h, w, total, dlow);
    ROBSEG_LAUNCH_CHECK();
  }
  if (loss_img || track_img || correct_img || valid_img) {
    const int tiles_per_img = (h + 1) * groups_x_of(W, R);
    rc = launch_loss_finalize(p.partials, B, tiles_per_img, grad_scale, (int64_t)H * W, loss_img, track_img,
                              correct_img, valid_img, stream);
    if (rc != 0) return rc;
  }
  return 0;
}

extern "C" int robseg_loss_upsampled_fwd_bwd(const float* low, const int64_t* labels, const float* class_w,
                                             int loss_kind, int ignore_index, int B, int C, int h, int w,
                                             int H, int W, const float* grad_scale, float* dlow,
                                             int64_t* pred, float* loss_img, float* track_img,
                                             int32_t* correct_img, int32_t* valid_img, void* workspace,
                                             size_t workspace_bytes, robseg_stream_t stream) {
  return loss_upsampled_impl(low, labels, class_w, loss_kind, ignore_index, B, C, h, w, H, W, grad_scale, dlow,
                             pred, loss_img, track_img, correct_img, valid_img, nullptr, workspace,
                             workspace_bytes, stream);
}

extern "C" int robseg_loss_upsampled_fwd_bwd_counts(const float* low, const int64_t* labels,
                                                    const float* class_w, int loss_kind, int ignore_index,
                                                    int B, int C, int h, int w, int H, int W,
                                                    const float* grad_scale, float* dlow, int64_t* pred,
                                                    float* loss_img, float* track_img, int32_t* correct_img,
                                                    int32_t* valid_img, int64_t* counts, void* workspace,
                                                    size_t workspace_bytes, robseg_stream_t stream) {
  ROBSEG_REQUIRE(counts != nullptr, "counts must not be NULL");
  return loss_upsampled_impl(low, labels, class_w, loss_kind, ignore_index, B, C, h, w, H, W, grad_scale, dlow,
                             pred, loss_img, track_img, correct_img, valid_img, counts, workspace,
                             workspace_bytes, stream);
}
