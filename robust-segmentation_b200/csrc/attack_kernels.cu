// Elementwise attack-state kernels: APGD / PGD update, projection, device-side step-size and
// best-point bookkeeping, flag-driven row copies.  All HBM-bound streaming kernels: 128-bit
// coalesced loads/stores, one launch each, no host synchronisation.
//
// Reference arithmetic: semseg/attacker.py:388-410 (step), :485-551 (bookkeeping),
// :243-248 (check_oscillation), :683-690 (_project); semseg/val.py:169-172,210-213 (PGD).
// The update kernels replay the reference's fp32 op chain with explicit round-to-nearest
// intrinsics (no FMA contraction), so their output is bit-identical to the ATen chain.
#include "common.cuh"

namespace robseg {

__device__ __forceinline__ float sgn(float g) { return (float)(g > 0.f) - (float)(g < 0.f); }
__device__ __forceinline__ float clip01(float v) { return fminf(fmaxf(v, 0.f), 1.f); }

__device__ __forceinline__ float apgd_elem(float x, float xa, float xo, float g, float st,
                                           float eps, float a, float oma) {
  const float lo = __fsub_rn(x, eps), hi = __fadd_rn(x, eps);
  const float g2 = __fsub_rn(xa, xo);
  float z = __fadd_rn(xa, __fmul_rn(st, sgn(g)));
  z = clip01(fminf(fmaxf(z, lo), hi));
  float t = __fadd_rn(xa, __fmul_rn(__fsub_rn(z, xa), a));
  t = __fadd_rn(t, __fmul_rn(g2, oma));
  return clip01(fminf(fmaxf(t, lo), hi));
}

// grid (chunks, B); VEC4 = rows are 16B-aligned multiples of 4 floats
template <bool VEC4>
__global__ void __launch_bounds__(256)
    apgd_step_kernel(const float* __restrict__ x, const float* __restrict__ xa,
                     const float* __restrict__ xo, const float* __restrict__ g,
                     const float* __restrict__ step, float eps, float a, float oma,
                     int64_t n_per_img, float* __restrict__ out) {
  const int b = blockIdx.y;
  const float st = __ldg(step + b);
  const int64_t base = (int64_t)b * n_per_img;
  if constexpr (VEC4) {
    const int64_t n4 = n_per_img >> 2;
    const float4* x4 = reinterpret_cast<const float4*>(x + base);
    const float4* a4 = reinterpret_cast<const float4*>(xa + base);
    const float4* o4 = reinterpret_cast<const float4*>(xo + base);
    const float4* g4 = reinterpret_cast<const float4*>(g + base);
    float4* r4 = reinterpret_cast<float4*>(out + base);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4;
         i += (int64_t)gridDim.x * blockDim.x) {
      const float4 vx = __ldcs(x4 + i), va = __ldcs(a4 + i), vo = __ldcs(o4 + i), vg = __ldcs(g4 + i);
      float4 r;
      r.x = apgd_elem(vx.x, va.x, vo.x, vg.x, st, eps, a, oma);
      r.y = apgd_elem(vx.y, va.y, vo.y, vg.y, st, eps, a, oma);
      r.z = apgd_elem(vx.z, va.z, vo.z, vg.z, st, eps, a, oma);
      r.w = apgd_elem(vx.w, va.w, vo.w, vg.w, st, eps, a, oma);
      r4[i] = r;
    }
  } else {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_per_img;
         i += (int64_t)gridDim.x * blockDim.x)
      out[base + i] = apgd_elem(x[base + i], xa[base + i], xo[base + i], g[base + i], st, eps, a, oma);
  }
}

// The same update with the PREVIOUS iteration's flag-driven row copies folded in (semseg/attacker.py:
// 494-495, 523-525, 546-548): the step reads x_adv and grad anyway, so the rows selected by the
// bookkeeping kernel are stored to x_best_adv / x_best / grad_best from the registers that hold them,
// and restarted rows take x_best / grad_best as their x_adv / grad (written back: x_adv becomes x_old).
// Replaces two row_select launches per iteration whose source rows this kernel re-read one launch later.
template <bool VEC4>
__global__ void __launch_bounds__(256)
    apgd_step_fused_kernel(const float* __restrict__ x, float* xa_io, const float* __restrict__ xo,
                           float* g_io, const float* __restrict__ step, float eps, float a, float oma,
                           int B, int64_t n_per_img, float* __restrict__ out,
                           const int32_t* __restrict__ flags, float* x_best_adv, float* x_best,
                           float* grad_best) {
  const int b = blockIdx.y;
  const float st = __ldg(step + b);
  const bool f_adv = __ldg(flags + b) != 0, f_best = __ldg(flags + B + b) != 0;
  const bool f_restart = __ldg(flags + 2 * B + b) != 0 && !f_best;
  const int64_t base = (int64_t)b * n_per_img;
  if constexpr (VEC4) {
    const int64_t n4 = n_per_img >> 2;
    const float4* x4 = reinterpret_cast<const float4*>(x + base);
    float4* a4 = reinterpret_cast<float4*>(xa_io + base);
    const float4* o4 = reinterpret_cast<const float4*>(xo + base);
    float4* g4 = reinterpret_cast<float4*>(g_io + base);
    float4* r4 = reinterpret_cast<float4*>(out + base);
    float4* ba4 = reinterpret_cast<float4*>(x_best_adv + base);
    float4* bx4 = reinterpret_cast<float4*>(x_best + base);
    float4* bg4 = reinterpret_cast<float4*>(grad_best + base);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4;
         i += (int64_t)gridDim.x * blockDim.x) {
      const float4 vx = __ldcs(x4 + i), vo = __ldcs(o4 + i);
      float4 va = a4[i], vg = g4[i];
      if (f_adv) ba4[i] = va;
      if (f_best) bx4[i] = va, bg4[i] = vg;
      if (f_restart) {
        va = bx4[i], vg = bg4[i];
        a4[i] = va, g4[i] = vg;
      }
      float4 r;
      r.x = apgd_elem(vx.x, va.x, vo.x, vg.x, st, eps, a, oma);
      r.y = apgd_elem(vx.y, va.y, vo.y, vg.y, st, eps, a, oma);
      r.z = apgd_elem(vx.z, va.z, vo.z, vg.z, st, eps, a, oma);
      r.w = apgd_elem(vx.w, va.w, vo.w, vg.w, st, eps, a, oma);
      r4[i] = r;
    }
  } else {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_per_img;
         i += (int64_t)gridDim.x * blockDim.x) {
      float va = xa_io[base + i], vg = g_io[base + i];
      if (f_adv) x_best_adv[base + i] = va;
      if (f_best) x_best[base + i] = va, grad_best[base + i] = vg;
      if (f_restart) {
        va = x_best[base + i], vg = grad_best[base + i];
        xa_io[base + i] = va, g_io[base + i] = vg;
      }
      out[base + i] = apgd_elem(x[base + i], va, xo[base + i], vg, st, eps, a, oma);
    }
  }
}

// CUDA-graph form of the iteration's first launch: every per-iteration scalar comes
// from the device control block ctl (robseg_b200.h: [0] iteration, [2] eps bits), the buffers never rotate
// -- x_old <- x_adv and x_adv <- new point are written in place by the thread that read them -- so the same
// captured launch is valid for every iteration of every stage.  Includes the folded row copies above.
template <bool VEC4>
__global__ void __launch_bounds__(256)
    apgd_step_ctl_kernel(const float* __restrict__ x, float* xa_io, float* xo_io, float* g_io,
                         const float* __restrict__ step, const int32_t* __restrict__ ctl, int B,
                         int64_t n_per_img, const int32_t* __restrict__ flags, float* x_best_adv,
                         float* x_best, float* grad_best) {
  const int b = blockIdx.y;
  const float st = __ldg(step + b);
  const float a = __ldg(ctl + ROBSEG_CTL_ITER) == 0 ? 1.0f : 0.75f, oma = 1.0f - a;
  const float eps = __int_as_float(__ldg(ctl + ROBSEG_CTL_EPS));
  const bool f_adv = __ldg(flags + b) != 0, f_best = __ldg(flags + B + b) != 0;
  const bool f_restart = __ldg(flags + 2 * B + b) != 0 && !f_best;
  const int64_t base = (int64_t)b * n_per_img;
  if constexpr (VEC4) {
    const int64_t n4 = n_per_img >> 2;
    const float4* x4 = reinterpret_cast<const float4*>(x + base);
    float4* a4 = reinterpret_cast<float4*>(xa_io + base);
    float4* o4 = reinterpret_cast<float4*>(xo_io + base);
    float4* g4 = reinterpret_cast<float4*>(g_io + base);
    float4* ba4 = reinterpret_cast<float4*>(x_best_adv + base);
    float4* bx4 = reinterpret_cast<float4*>(x_best + base);
    float4* bg4 = reinterpret_cast<float4*>(grad_best + base);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4;
         i += (int64_t)gridDim.x * blockDim.x) {
      const float4 vx = __ldcs(x4 + i), vo = o4[i];
      float4 va = a4[i], vg = g4[i];
      if (f_adv) ba4[i] = va;
      if (f_best) bx4[i] = va, bg4[i] = vg;
      if (f_restart) {
        va = bx4[i], vg = bg4[i];
        g4[i] = vg;
      }
      float4 r;
      r.x = apgd_elem(vx.x, va.x, vo.x, vg.x, st, eps, a, oma);
      r.y = apgd_elem(vx.y, va.y, vo.y, vg.y, st, eps, a, oma);
      r.z = apgd_elem(vx.z, va.z, vo.z, vg.z, st, eps, a, oma);
      r.w = apgd_elem(vx.w, va.w, vo.w, vg.w, st, eps, a, oma);
      o4[i] = va;
      a4[i] = r;
    }
  } else {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_per_img;
         i += (int64_t)gridDim.x * blockDim.x) {
      float va = xa_io[base + i], vg = g_io[base + i];
      const float vo = xo_io[base + i];
      if (f_adv) x_best_adv[base + i] = va;
      if (f_best) x_best[base + i] = va, grad_best[base + i] = vg;
      if (f_restart) {
        va = x_best[base + i], vg = grad_best[base + i];
        g_io[base + i] = vg;
      }
      xo_io[base + i] = va;
      xa_io[base + i] = apgd_elem(x[base + i], va, vo, vg, st, eps, a, oma);
    }
  }
}

// out = clip01(x + clip(z-x, +-eps))   or, with noise, clip01(x + eps*noise)
__global__ void __launch_bounds__(256)
    project_kernel(const float* __restrict__ z, const float* __restrict__ x,
                   const float* __restrict__ noise, float eps, int64_t n, float* __restrict__ out) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
    const float xv = x[i];
    float r;
    if (noise != nullptr) {
      r = __fadd_rn(xv, __fmul_rn(eps, noise[i]));
    } else {
      const float d = fminf(fmaxf(__fsub_rn(z[i], xv), -eps), eps);
      r = __fadd_rn(xv, d);
    }
    out[i] = clip01(r);
  }
}

__global__ void __launch_bounds__(256)
    pgd_step_kernel(const float* __restrict__ X, float* __restrict__ delta,
                    const float* __restrict__ grad, float alpha, float eps, int mask_outside,
                    int clamp_next, int64_t n, float* __restrict__ x_next) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
    const float xv = X[i], d0 = delta[i];
    float g = grad[i];
    if (mask_outside) {
      const float s = __fadd_rn(xv, d0);
      if (!(s >= 0.f && s <= 1.f)) g = 0.f;
    }
    float d = __fadd_rn(d0, __fmul_rn(alpha, sgn(g)));
    d = __fsub_rn(clip01(__fadd_rn(xv, d)), xv);
    d = fminf(fmaxf(d, -eps), eps);
    delta[i] = d;
    if (x_next != nullptr) {
      const float s = __fadd_rn(xv, d);
      x_next[i] = clamp_next ? clip01(s) : s;
    }
  }
}

// One block, one thread per image (strided for B > blockDim).
__global__ void __launch_bounds__(1024)
    apgd_bookkeep_kernel(const int32_t* __restrict__ correct, const int32_t* __restrict__ valid,
                         const float* __restrict__ loss_indiv, float* acc, float* loss_best,
                         float* loss_best_last, float* reduced_last, float* step,
                         float* loss_steps, int n_iter, int iter, int check_k, int B, int64_t HW,
                         int early_stop, int32_t* flags, int32_t* done_flag, int32_t* done_host,
                         int32_t* ctl) {
  __shared__ int sh_any_nonzero;
  if (ctl != nullptr) {  // CUDA-graph form: iteration, its check window and the stage length live on the device
    iter = ctl[ROBSEG_CTL_ITER];
    n_iter = ctl[ROBSEG_CTL_NITER];
    check_k = ctl[ROBSEG_CTL_SCHED + iter];
  }
  const bool done = *done_flag != 0;
  if (threadIdx.x == 0) sh_any_nonzero = 0;
  __syncthreads();
  int local_nonzero = 0;
  for (int b = threadIdx.x; b < B; b += blockDim.x) {
    if (done) {
      flags[b] = 0, flags[B + b] = 0, flags[2 * B + b] = 0;
      continue;
    }
    // accuracy with ignored pixels counted correct (attacker.py:485-490)
    const float avg_acc = (float)((int64_t)correct[b] + (HW - (int64_t)valid[b])) / (float)HW;
    const float a_old = acc[b];
    const int ind_pred = avg_acc <= a_old;
    const float a_new = fminf(a_old, avg_acc);
    acc[b] = a_new;
    local_nonzero |= (a_new != 0.f);
    // best loss (attacker.py:519-526)
    const float y1 = loss_indiv[b];
    loss_steps[(int64_t)iter * B + b] = y1;
    float lb = loss_best[b];
    const int ind = y1 > lb;
    if (ind) lb = y1, loss_best[b] = y1;
    int reduce = 0;
    if (check_k > 0) {  // attacker.py:530-551
      float t = 0.f;
      for (int c = 0; c < check_k; ++c) {
        const int j = iter - c;
        int jm = j - 1;
        if (jm < 0) jm += n_iter;  // x[-1]: last row
        const float lj = (c == 0) ? y1 : loss_steps[(int64_t)j * B + b];
        t += (lj > loss_steps[(int64_t)jm * B + b]) ? 1.f : 0.f;
      }
      float osc = (t <= (float)check_k * 0.75f) ? 1.f : 0.f;
      const float no_impr = (1.f - reduced_last[b]) * ((loss_best_last[b] >= lb) ? 1.f : 0.f);
      osc = fmaxf(osc, no_impr);
      reduced_last[b] = osc;
      loss_best_last[b] = lb;
      if (osc > 0.f) step[b] = step[b] / 2.f, reduce = 1;
    }
    flags[b] = ind_pred, flags[B + b] = ind, flags[2 * B + b] = reduce;
  }
  if (local_nonzero) atomicOr(&sh_any_nonzero, 1);
  __syncthreads();
  if (threadIdx.x == 0 && !done && early_stop && sh_any_nonzero == 0) {
    *done_flag = 1;
    if (done_host != nullptr) {
      *reinterpret_cast<volatile int32_t*>(done_host) = 1;
      __threadfence_system();
    }
  }
  // every thread read ctl before the barrier above; the next replay sees the next iteration
  if (threadIdx.x == 0 && ctl != nullptr && iter + 1 < ROBSEG_CTL_MAX_ITER) ctl[ROBSEG_CTL_ITER] = iter + 1;
}

struct RowJobs {
  robseg_row_job j[ROBSEG_MAX_ROW_JOBS];
};

// grid (chunks, B, n_jobs)
__global__ void __launch_bounds__(256) row_select_kernel(const RowJobs jobs) {
  const robseg_row_job jb = jobs.j[blockIdx.z];
  const int b = blockIdx.y;
  if (__ldg(jb.flags + b) == 0) return;
  if (jb.unless != nullptr && __ldg(jb.unless + b) != 0) return;
  char* dst = static_cast<char*>(jb.dst) + (int64_t)b * jb.row_bytes;
  const char* src = static_cast<const char*>(jb.src) + (int64_t)b * jb.row_bytes;
  const bool v16 = ((reinterpret_cast<uintptr_t>(dst) | reinterpret_cast<uintptr_t>(src) |
                     (uintptr_t)jb.row_bytes) & 15) == 0;
  if (v16) {
    const int64_t n = jb.row_bytes >> 4;
    const uint4* s = reinterpret_cast<const uint4*>(src);
    uint4* d = reinterpret_cast<uint4*>(dst);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (int64_t)gridDim.x * blockDim.x)
      d[i] = s[i];
  } else {
    const int64_t n = jb.row_bytes >> 2;
    const uint32_t* s = reinterpret_cast<const uint32_t*>(src);
    uint32_t* d = reinterpret_cast<uint32_t*>(dst);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (int64_t)gridDim.x * blockDim.x)
      d[i] = s[i];
  }
}

static int ew_grid(int64_t n, int per_thread = 1) {
  int64_t blocks = (n + 256ll * per_thread - 1) / (256ll * per_thread);
  const int64_t cap = (int64_t)sm_count() * 32;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}

}  // namespace robseg

using namespace robseg;

extern "C" int robseg_apgd_step(const float* x, const float* x_adv, const float* x_old,
                                const float* grad, const float* step, float eps, float a,
                                float one_minus_a, int B, int64_t n_per_img, float* x_new,
                                robseg_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  ROBSEG_REQUIRE(x && x_adv && x_old && grad && step && x_new, "NULL pointer");
  ROBSEG_REQUIRE(B > 0 && B <= 65535 && n_per_img > 0, "bad shape B=%d n=%lld", B, (long long)n_per_img);
  ROBSEG_REQUIRE(x_new != x && x_new != x_adv && x_new != x_old && x_new != grad,
                 "x_new must not alias an input");
  const uintptr_t al = reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(x_adv) |
                       reinterpret_cast<uintptr_t>(x_old) | reinterpret_cast<uintptr_t>(grad) |
                       reinterpret_cast<uintptr_t>(x_new);
  const bool vec = (al % 16 == 0) && (n_per_img % 4 == 0);
  const int64_t work = vec ? n_per_img / 4 : n_per_img;
  int gx = (int)((work + 255) / 256);
  const int cap = (sm_count() * 32 + B - 1) / B;
  if (gx > cap) gx = cap < 1 ? 1 : cap;
  dim3 grid(gx, B);
  if (vec)
    apgd_step_kernel<true><<<grid, 256, 0, stream>>>(x, x_adv, x_old, grad, step, eps, a,
                                                     one_minus_a, n_per_img, x_new);
  else
    apgd_step_kernel<false><<<grid, 256, 0, stream>>>(x, x_adv, x_old, grad, step, eps, a,
                                                      one_minus_a, n_per_img, x_new);
  ROBSEG_LAUNCH_CHECK();
  return 0;
}

extern "C" int robseg_apgd_step_fused(const float* x, float* x_adv, const float* x_old, float* grad,
                                      const float* step, float eps, float a, float one_minus_a, int B,
                                      int64_t n_per_img, float* x_new, const int32_t* flags,
                                      float* x_best_adv, float* x_best, float* grad_best,
                                      robseg_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  ROBSEG_REQUIRE(x && x_adv && x_old && grad && step && x_new && flags && x_best_adv && x_best && grad_best,
                 "NULL pointer");
  ROBSEG_REQUIRE(B > 0 && B <= 65535 && n_per_img > 0, "bad shape B=%d n=%lld", B, (long long)n_per_img);
  ROBSEG_REQUIRE(x_new != x && x_new != x_adv && x_new != x_old && x_new != grad && x_new != x_best_adv &&
                     x_new != x_best && x_new != grad_best,
                 "x_new must not alias another buffer");
  const uintptr_t al = reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(x_adv) |
                       reinterpret_cast<uintptr_t>(x_old) | reinterpret_cast<uintptr_t>(grad) |
                       reinterpret_cast<uintptr_t>(x_new) | reinterpret_cast<uintptr_t>(x_best_adv) |
                       reinterpret_cast<uintptr_t>(x_best) | reinterpret_cast<uintptr_t>(grad_best);
  const bool vec = (al % 16 == 0) && (n_per_img % 4 == 0);
  const int64_t work = vec ? n_per_img / 4 : n_per_img;
  int gx = (int)((work + 255) / 256);
  const int cap = (sm_count() * 32 + B - 1) / B;
  if (gx > cap) gx = cap < 1 ? 1 : cap;
  dim3 grid(gx, B);
  if (vec)
    apgd_step_fused_kernel<true><<<grid, 256, 0, stream>>>(x, x_adv, x_old, grad, step, eps, a, one_minus_a, B,
                                                           n_per_img, x_new, flags, x_best_adv, x_best, grad_best);
  else
    apgd_step_fused_kernel<false><<<grid, 256, 0, stream>>>(x, x_adv, x_old, grad, step, eps, a, one_minus_a, B,
                                                            n_per_img, x_new, flags, x_best_adv, x_best, grad_best);
  ROBSEG_LAUNCH_CHECK();
  return 0;
}

extern "C" int robseg_project_linf(const float* z, const float* x, const float* noise, float eps,
                                   int64_t n, float* out, robseg_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  ROBSEG_REQUIRE(x && out && (z || noise), "NULL pointer");
  ROBSEG_REQUIRE(n > 0, "bad n");
  project_kernel<<<ew_grid(n), 256, 0, stream>>>(z, x, noise, eps, n, out);
  ROBSEG_LAUNCH_CHECK();
  return 0;
}

extern "C" int robseg_pgd_step(const float* X, float* delta, const float* grad, float alpha,
                               float eps, int mask_outside, int clamp_next, int64_t n,
                               float* x_next, robseg_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  ROBSEG_REQUIRE(X && delta && grad, "NULL pointer");
  ROBSEG_REQUIRE(n > 0, "bad n");
  pgd_step_kernel<<<ew_grid(n), 256, 0, stream>>>(X, delta, grad, alpha, eps, mask_outside,
                                                  clamp_next, n, x_next);
  ROBSEG_LAUNCH_CHECK();
  return 0;
}

extern "C" int robseg_apgd_bookkeep(const int32_t* correct, const int32_t* valid,
                                    const float* loss_indiv, float* acc, float* loss_best,
                                    float* loss_best_last, float* reduced_last, float* step,
                                    float* loss_steps, int n_iter, int iter, int check_k, int B,
                                    int64_t HW, int early_stop, int32_t* flags_out,
                                    int32_t* done_flag, int32_t* done_host,
                                    robseg_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  ROBSEG_REQUIRE(correct && valid && loss_indiv && acc && loss_best && loss_best_last &&
                     reduced_last && step && loss_steps && flags_out && done_flag,
                 "NULL pointer");
  ROBSEG_REQUIRE(B > 0 && n_iter > 0 && iter >= 0 && iter < n_iter && check_k >= 0 &&
                     check_k <= iter + 1 && HW > 0,
                 "bad arguments B=%d n_iter=%d iter=%d check_k=%d", B, n_iter, iter, check_k);
  int threads = ((B + 31) / 32) * 32;
  if (threads > 1024) threads = 1024;
  apgd_bookkeep_kernel<<<1, threads, 0, stream>>>(correct, valid, loss_indiv, acc, loss_best,
                                                  loss_best_last, reduced_last, step, loss_steps,
                                                  n_iter, iter, check_k, B, HW, early_stop,
                                                  flags_out, done_flag, done_host, nullptr);
  ROBSEG_LAUNCH_CHECK();
  return 0;
}

extern "C" int robseg_apgd_bookkeep_ctl(const int32_t* correct, const int32_t* valid,
                                        const float* loss_indiv, float* acc, float* loss_best,
                                        float* loss_best_last, float* reduced_last, float* step,
                                        float* loss_steps, int32_t* ctl, int B, int64_t HW,
                                        int early_stop, int32_t* flags_out, int32_t* done_flag,
                                        robseg_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  ROBSEG_REQUIRE(correct && valid && loss_indiv && acc && loss_best && loss_best_last && reduced_last &&
                     step && loss_steps && flags_out && done_flag && ctl,
                 "NULL pointer");
  ROBSEG_REQUIRE(B > 0 && HW > 0, "bad arguments B=%d", B);
  int threads = ((B + 31) / 32) * 32;
  if (threads > 1024) threads = 1024;
  apgd_bookkeep_kernel<<<1, threads, 0, stream>>>(correct, valid, loss_indiv, acc, loss_best,
                                                  loss_best_last, reduced_last, step, loss_steps, 0, 0, 0,
                                                  B, HW, early_stop, flags_out, done_flag, nullptr, ctl);
  ROBSEG_LAUNCH_CHECK();
  return 0;
}

extern "C" int robseg_apgd_step_ctl(const float* x, float* x_adv, float* x_old, float* grad,
                                    const float* step, const int32_t* ctl, int B, int64_t n_per_img,
                                    const int32_t* flags, float* x_best_adv, float* x_best,
                                    float* grad_best, robseg_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  ROBSEG_REQUIRE(x && x_adv && x_old && grad && step && ctl && flags && x_best_adv && x_best && grad_best,
                 "NULL pointer");
  ROBSEG_REQUIRE(B > 0 && B <= 65535 && n_per_img > 0, "bad shape B=%d n=%lld", B, (long long)n_per_img);
  ROBSEG_REQUIRE(x_adv != x && x_old != x && x_old != x_adv, "x / x_adv / x_old must be distinct buffers");
  const uintptr_t al = reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(x_adv) |
                       reinterpret_cast<uintptr_t>(x_old) | reinterpret_cast<uintptr_t>(grad) |
                       reinterpret_cast<uintptr_t>(x_best_adv) | reinterpret_cast<uintptr_t>(x_best) |
                       reinterpret_cast<uintptr_t>(grad_best);
  const bool vec = (al % 16 == 0) && (n_per_img % 4 == 0);
  const int64_t work = vec ? n_per_img / 4 : n_per_img;
  int gx = (int)((work + 255) / 256);
  const int cap = (sm_count() * 32 + B - 1) / B;
  if (gx > cap) gx = cap < 1 ? 1 : cap;
  dim3 grid(gx, B);
  if (vec)
    apgd_step_ctl_kernel<true><<<grid, 256, 0, stream>>>(x, x_adv, x_old, grad, step, ctl, B, n_per_img, flags,
                                                         x_best_adv, x_best, grad_best);
  else
    apgd_step_ctl_kernel<false><<<grid, 256, 0, stream>>>(x, x_adv, x_old, grad, step, ctl, B, n_per_img, flags,
                                                          x_best_adv, x_best, grad_best);
  ROBSEG_LAUNCH_CHECK();
  return 0;
}

extern "C" int robseg_row_select(const robseg_row_job* jobs_host, int n_jobs, int B,
                                 robseg_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  ROBSEG_REQUIRE(jobs_host && n_jobs > 0 && n_jobs <= ROBSEG_MAX_ROW_JOBS, "bad job count %d", n_jobs);
  ROBSEG_REQUIRE(B > 0 && B <= 65535, "bad B=%d", B);
  RowJobs jobs{};
  int64_t max_bytes = 0;
  for (int i = 0; i < n_jobs; ++i) {
    const robseg_row_job& j = jobs_host[i];
    ROBSEG_REQUIRE(j.dst && j.src && j.flags && j.row_bytes > 0 && j.row_bytes % 4 == 0,
                   "bad row job %d", i);
    jobs.j[i] = j;
    if (j.row_bytes > max_bytes) max_bytes = j.row_bytes;
  }
  int gx = (int)((max_bytes / 16 + 256 * 4 - 1) / (256 * 4));
  if (gx < 1) gx = 1;
  if (gx > 1024) gx = 1024;
  row_select_kernel<<<dim3(gx, B, n_jobs), 256, 0, stream>>>(jobs);
  ROBSEG_LAUNCH_CHECK();
  return 0;
}
