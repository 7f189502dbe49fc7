/*
 * robseg_b200.h -- C ABI of librobseg_b200.so: the B200 (sm_100a) attack-side hot path of
 * the Segmentation Ensemble Attack (SEA) / PIR-AT inner attack.
 *
 * The reference (nmndeep/Robust-Segmentation) is pure Python: its "plugin interface" for
 * this path is a set of module-level callables.  Each entry point
 * below replaces the ATen op chain behind one of them; the citation is the reference
 * file:line whose arithmetic the entry point reproduces.  The host-side mirror of the
 * reference call surface (the semseg/ modules under robust-segmentation_b200) binds these through
 * ctypes; INTEGRATION.md shows the stub.
 *
 * Conventions
 *  - every pointer is a DEVICE pointer unless its name ends in _host (functions named *_host
 *    take host pointers only and launch nothing);
 *  - the library allocates nothing persistent: the caller owns every buffer and workspace;
 *  - all launches go to the stream passed in (a cudaStream_t cast to void*), no internal
 *    synchronisation;
 *  - return value: 0 on success, otherwise a cudaError_t value or ROBSEG_EINVAL; a
 *    human-readable reason is available from robseg_last_error() (thread local);
 *  - nothing throws across the boundary;
 *  - logits are NCHW-contiguous [B, C, HW] (HW = H*W), labels int64 [B, HW], -1 = ignore.
 */
#ifndef ROBSEG_B200_H_
#define ROBSEG_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ROBSEG_ABI_VERSION 1
#define ROBSEG_EINVAL (-22)

typedef void* robseg_stream_t; /* cudaStream_t */

enum robseg_dtype { ROBSEG_F32 = 0, ROBSEG_BF16 = 1 };

/* semseg/attacker.py:251-257 criterion_dict keys ("ce"/"ce-avg" -> CE, "mask-ce-avg",
 * "mask-ce-bal", "js-avg"); ARGMAX = no loss, only the prediction/accuracy outputs. */
enum robseg_loss_kind {
  ROBSEG_LOSS_CE = 0,
  ROBSEG_LOSS_MASK_CE = 1,
  ROBSEG_LOSS_MASK_CE_BAL = 2,
  ROBSEG_LOSS_JS = 3,
  ROBSEG_LOSS_ARGMAX = 4
};

/* ABI version of the loaded library (== ROBSEG_ABI_VERSION). */
int robseg_version(void);

/* Reason of the last non-zero return on this thread ("" if none). */
const char* robseg_last_error(void);

/*
 * Measurement hook (bench.py's `roofline`): the NEXT robseg_loss_fwd_bwd* / robseg_loss_upsampled_fwd_bwd*
 * call on this thread records `start_event` immediately before and `stop_event` immediately after its
 * main kernel (loss_tma_kernel / loss_generic_* / loss_up_kernel) on the call's stream, leaving the small
 * counter-zeroing, counter-fold and finalize launches of the same call outside the bracket.  Both are
 * cudaEvent_t created by the caller with timing enabled; one-shot (cleared by the call that uses them);
 * pass NULL, NULL to cancel.  No effect on results.
 */
int robseg_profile_next_kernel(void* start_event, void* stop_event);

/* Bytes of scratch robseg_loss_fwd_bwd needs for a problem of this shape. */
size_t robseg_loss_workspace_bytes(int B, int C, int64_t HW, int dtype);

/*
 * Fused per-pixel softmax + loss + d(loss)/d(logits) + argmax + per-image partial sums:
 * ONE read of the logits, ONE write of the gradient.
 *
 * Replaces, per APGD / PGD iteration:
 *   semseg/attacker.py:143-152  masked_cross_entropy           (kind MASK_CE)
 *   semseg/attacker.py:155-173  masked_cross_entropy_balanced  (kind MASK_CE_BAL, class_w)
 *   semseg/attacker.py:187-234  js_div_fn / js_loss            (kind JS)
 *   semseg/attacker.py:252-253  "ce" / "ce-avg"                (kind CE)
 *   semseg/attacker.py:237-240  pixel_to_img_loss              (loss_img, track_img)
 *   semseg/attacker.py:347-350,462-469  autograd.grad of the summed per-image loss wrt logits
 *   semseg/attacker.py:353-361,473-475  track loss "ce-avg"    (track_img, same pass)
 *   semseg/attacker.py:370-373,485-490  argmax / per-image accuracy (pred, correct, valid)
 *   semseg/val.py:121-127       PIR-AT losses "pgd", "mask-ce-avg", "js-avg"
 *
 *  logits       [B,C,HW]  dtype f32 or bf16
 *  labels       [B,HW]    int64; == ignore_index -> pixel contributes nothing
 *  class_w      [C] f32   or NULL (MASK_CE_BAL only; NULL = unweighted)
 *  grad_scale   [B] f32   per-image upstream scale g_b, or NULL for 1/HW
 *  upstream_pix [B,HW] f32 optional per-pixel upstream gradient (multiplies g_b); NULL = 1
 *  dlogits      [B,C,HW]  same dtype as logits, or NULL (loss only)
 *                         dlogits[b,k,p] = g_b * up[b,p] * coef(b,p) * (softmax_k - 1[k=y])
 *  loss_pix     [B,HW] f32 or NULL   per-pixel loss (criterion_dict[...] output)
 *  pred         [B,HW] int64 or NULL argmax over C, lowest index on ties
 *  loss_img     [B] f32 or NULL      g_b * sum_p loss
 *  track_img    [B] f32 or NULL      (1/HW) * sum_p [y != ignore] (lse - z_y)
 *  correct_img  [B] int32 or NULL    #[argmax == y]
 *  valid_img    [B] int32 or NULL    #[y != ignore]
 */
int robseg_loss_fwd_bwd(const void* logits, int dtype, const int64_t* labels,
                        const float* class_w, int loss_kind, int ignore_index, int B, int C,
                        int64_t HW, const float* grad_scale, const float* upstream_pix,
                        void* dlogits, float* loss_pix, int64_t* pred, float* loss_img,
                        float* track_img, int32_t* correct_img, int32_t* valid_img,
                        void* workspace, size_t workspace_bytes, robseg_stream_t stream);

/*
 * The same fused loss taken THROUGH the consumer's final bilinear up-sampling (SURVEY.md section 8f
 * rank 1): logits = interpolate(low, size=(H,W), mode="bilinear", align_corners=False) as in
 * semseg/models/uperforseg.py:416-418 (H = 4h) and semseg/models/segmenter.py:228 (H = 16h), followed by
 * everything robseg_loss_fwd_bwd computes -- without ever materialising the [B,C,H,W] logits or their
 * gradient.  The kernel interpolates on the fly from the low-resolution tensor and returns
 *   dlow [B,C,h,w] = (d interpolate / d low)^T dlogits      (deterministic gather, no atomics)
 * so the caller continues with autograd.grad(low, x, grad_outputs=dlow).
 *  low [B,C,h,w] f32; labels [B,H,W] int64; H = R*h, W = R*w with R in {2,4,8,16};
 *  dlow may be NULL (loss / argmax only); pred [B,H,W] int64 or NULL; the per-image outputs as above.
 *  workspace: robseg_loss_upsampled_workspace_bytes(...) bytes, 16-byte aligned.
 */
size_t robseg_loss_upsampled_workspace_bytes(int B, int C, int h, int w, int H, int W);
int robseg_loss_upsampled_fwd_bwd(const float* low, const int64_t* labels, const float* class_w,
                                  int loss_kind, int ignore_index, int B, int C, int h, int w, int H,
                                  int W, const float* grad_scale, float* dlow, int64_t* pred,
                                  float* loss_img, float* track_img, int32_t* correct_img,
                                  int32_t* valid_img, void* workspace, size_t workspace_bytes,
                                  robseg_stream_t stream);

/*
 * The two fused-loss calls with the per-image class counters of compute_iou_acc
 * (semseg/attacker.py:9-52, called on every iteration's argmax at :496-498; the same counters
 * tools/infer.py:86-116 and evalSEA, tools/worse_only.py:30-66, derive from a prediction map) taken
 * in the kernel's argmax pass, so neither the int64 prediction map nor a robseg_pixel_hist launch
 * is needed to score an adversarial point:
 *   counts [B,3,C] int64 (written by the call): [b,0,c] = #{pred == label == c},
 *   [b,1,c] = #{label == c}, [b,2,c] = #{pred == c, label valid}; ignored pixels add nothing.
 * Exact (integer reductions into 8 copies held in the workspace -- the ..._workspace_bytes functions account
 * for them -- added up by a small kernel), identical to robseg_pixel_hist on the pred this call would return.
 * All other arguments as in robseg_loss_fwd_bwd / robseg_loss_upsampled_fwd_bwd.
 */
int robseg_loss_fwd_bwd_counts(const void* logits, int dtype, const int64_t* labels,
                               const float* class_w, int loss_kind, int ignore_index, int B, int C,
                               int64_t HW, const float* grad_scale, const float* upstream_pix,
                               void* dlogits, float* loss_pix, int64_t* pred, float* loss_img,
                               float* track_img, int32_t* correct_img, int32_t* valid_img,
                               int64_t* counts, void* workspace, size_t workspace_bytes,
                               robseg_stream_t stream);
int robseg_loss_upsampled_fwd_bwd_counts(const float* low, const int64_t* labels, const float* class_w,
                                         int loss_kind, int ignore_index, int B, int C, int h, int w,
                                         int H, int W, const float* grad_scale, float* dlow,
                                         int64_t* pred, float* loss_img, float* track_img,
                                         int32_t* correct_img, int32_t* valid_img, int64_t* counts,
                                         void* workspace, size_t workspace_bytes,
                                         robseg_stream_t stream);

/*
 * One L-inf APGD update for the whole batch, bit-exact with the fp32 op chain of
 * semseg/attacker.py:388-410:
 *   g2 = x_adv - x_old
 *   z  = clip01(min(max(x_adv + step_b*sign(grad), x-eps), x+eps))
 *   x_new = clip01(min(max(x_adv + (z-x_adv)*a + g2*(1-a), x-eps), x+eps))
 * x, x_adv, x_old, grad, x_new: f32 [B, n_per_img]; step: f32 [B].  x_new must not alias
 * the inputs (the caller rotates buffers: x_old <- x_adv <- x_new).
 */
int robseg_apgd_step(const float* x, const float* x_adv, const float* x_old, const float* grad,
                     const float* step, float eps, float a, float one_minus_a, int B,
                     int64_t n_per_img, float* x_new, robseg_stream_t stream);

/*
 * robseg_apgd_step with the PREVIOUS iteration's boolean-index row copies folded in
 * (semseg/attacker.py:494-495 x_best_adv[ind_pred] = x_adv; :523-525 x_best[ind] = x_adv, grad_best[ind] =
 * grad; :546-548 x_adv[fl] = x_best[fl], grad[fl] = grad_best[fl]), driven by the [3,B] flags
 * robseg_apgd_bookkeep wrote: the step reads x_adv and grad anyway, so selected rows are stored from
 * registers and restarted rows read x_best / grad_best instead (and are written back to x_adv / grad,
 * in place).  Same arithmetic as robseg_apgd_step.  After the LAST iteration the caller flushes the
 * pending copies with robseg_row_select.
 */
int robseg_apgd_step_fused(const float* x, float* x_adv, const float* x_old, float* grad,
                           const float* step, float eps, float a, float one_minus_a, int B,
                           int64_t n_per_img, float* x_new, const int32_t* flags, float* x_best_adv,
                           float* x_best, float* grad_best, robseg_stream_t stream);

/*
 * CUDA-graph forms of the two launches whose arguments change from one APGD iteration to the next
 * (one captured graph per iteration, semseg/attacker.py:385-551).  The
 * per-iteration scalars live in a DEVICE control block of int32 words that the bookkeeping kernel itself
 * advances, so the captured launches are identical for every iteration and every stage:
 *   ctl[ROBSEG_CTL_ITER]       iteration index i (a = 1.0 at i = 0, else 0.75: attacker.py:387)
 *   ctl[ROBSEG_CTL_NITER]      n_iter of the stage (rows of loss_steps in use)
 *   ctl[ROBSEG_CTL_EPS]        eps of the stage, float bits
 *   ctl[ROBSEG_CTL_SCHED + i]  window k if the step-size check fires at iteration i, else 0
 * robseg_apgd_step_ctl = robseg_apgd_step_fused with a / eps from ctl and NO buffer rotation: x_old <- x_adv
 * and x_adv <- new point are written in place (a captured graph needs fixed addresses).
 * robseg_apgd_bookkeep_ctl = robseg_apgd_bookkeep with (iter, check_k, n_iter) from ctl; it increments
 * ctl[ROBSEG_CTL_ITER] when it is done.  The host writes ctl once per stage.
 */
#define ROBSEG_CTL_ITER 0
#define ROBSEG_CTL_NITER 1
#define ROBSEG_CTL_EPS 2
#define ROBSEG_CTL_SCHED 8
#define ROBSEG_CTL_MAX_ITER 4096
int robseg_apgd_step_ctl(const float* x, float* x_adv, float* x_old, float* grad, const float* step,
                         const int32_t* ctl, int B, int64_t n_per_img, const int32_t* flags,
                         float* x_best_adv, float* x_best, float* grad_best, robseg_stream_t stream);
int robseg_apgd_bookkeep_ctl(const int32_t* correct, const int32_t* valid, const float* loss_indiv,
                             float* acc, float* loss_best, float* loss_best_last, float* reduced_last,
                             float* step, float* loss_steps, int32_t* ctl, int B, int64_t HW,
                             int early_stop, int32_t* flags_out, int32_t* done_flag,
                             robseg_stream_t stream);

/*
 * z <- clip01(x + clip(z - x, -eps, eps)): the stage hand-off of apgd_largereps
 * (semseg/attacker.py:683-690).  With noise != NULL computes the random start instead,
 * out = clip01(x + eps*noise) (semseg/attacker.py:292-294, noise = 2*rand-1).
 */
int robseg_project_linf(const float* z_or_null, const float* x, const float* noise_or_null,
                        float eps, int64_t n, float* out, robseg_stream_t stream);

/*
 * PIR-AT inner PGD update, semseg/val.py:169-172 and :210-213:
 *   delta <- clip(clip01(X + (delta + alpha*sign(grad))) - X, -eps, eps)      (in place)
 * mask_outside != 0 zeroes grad where X+delta is outside [0,1] first (the clamp backward
 * of val.py:151).  x_next (optional) receives X+delta, clamped to [0,1] if clamp_next.
 */
int robseg_pgd_step(const float* X, float* delta, const float* grad, float alpha, float eps,
                    int mask_outside, int clamp_next, int64_t n, float* x_next,
                    robseg_stream_t stream);

/*
 * Per-iteration APGD bookkeeping on the device (no host sync), semseg/attacker.py:485-551:
 * accuracy / best-loss tracking, oscillation check (:243-248), step-size halving and the
 * three row-selection flag vectors the row copies below consume.  One thread per image.
 *
 *  correct, valid [B] int32  from robseg_loss_fwd_bwd;  loss_indiv [B] f32 = track_img
 *  state (all [B] f32 unless noted, updated in place): acc, loss_best, loss_best_last,
 *        reduced_last, step;  loss_steps [n_iter,B] f32
 *  iter           iteration index i;  check_k = window k when the check fires at i, else 0
 *  flags_out      [3,B] int32: row 0 = avg_acc <= acc (x_best_adv/pred_best update),
 *                 row 1 = loss_indiv > loss_best (x_best/grad_best update),
 *                 row 2 = step halved & restart from x_best (only rows where row 1 is 0
 *                 need a copy)
 *  done_flag      [1] int32 device; set to 1 when early_stop and sum(acc)==0
 *                 (attacker.py:568-569).  Once set, every later call leaves the state
 *                 untouched and emits all-zero flags.
 *  done_host      optional mapped pinned host int32 mirror of done_flag (may be NULL)
 */
int robseg_apgd_bookkeep(const int32_t* correct, const int32_t* valid, const float* loss_indiv,
                         float* acc, float* loss_best, float* loss_best_last,
                         float* reduced_last, float* step, float* loss_steps, int n_iter,
                         int iter, int check_k, int B, int64_t HW, int early_stop,
                         int32_t* flags_out, int32_t* done_flag, int32_t* done_host,
                         robseg_stream_t stream);

/*
 * Flag-driven row copies (the boolean-index assignments of semseg/attacker.py:494-495,
 * 523-525,547-548) in one launch: for each job j and row b, if flags[j][b] != 0 (and, when
 * unless[j] != NULL, unless[j][b] == 0) copy row_bytes[j] bytes dst[j][b] <- src[j][b].
 * Up to ROBSEG_MAX_ROW_JOBS jobs; row sizes must be multiples of 4 bytes.  Jobs of one call run
 * concurrently: no job may write a buffer another job of the same call reads (the APGD restart
 * copies x_adv <- x_best therefore go in a second call, after x_best_adv <- x_adv).
 */
#define ROBSEG_MAX_ROW_JOBS 8
typedef struct {
  void* dst;
  const void* src;
  const int32_t* flags;  /* [B] */
  const int32_t* unless; /* [B] or NULL */
  int64_t row_bytes;
} robseg_row_job;
int robseg_row_select(const robseg_row_job* jobs_host, int n_jobs, int B,
                      robseg_stream_t stream);

/*
 * Per-image confusion / intersection / union counters with warp-aggregated shared-memory
 * atomics.  Integer-exact replacement for the 2*C masked reductions of compute_iou_acc
 * (semseg/attacker.py:9-52), Metrics.update (semseg/metrics.py:27-33, hist[target,pred]),
 * eval_performance (tools/infer.py:86-116) and evalSEA (tools/worse_only.py:30-66,383-394).
 *
 *  pred   [n_img,HW] int64;  labels [n_lab_img,HW] int64, image i uses labels[i % n_lab_img]
 *         (n_lab_img = N for A stacked attacks over the same N targets)
 *  hist   [n_img,C,C] int64 or NULL   += #[label==t, pred==p], label != ignore
 *  hist_total [C,C] int64 or NULL     += the same, summed over images
 *  inter, tgt, prd [n_img,C] int64 or NULL  += #[pred==c==label], #[label==c],
 *         #[pred==c, label != ignore].  Outputs ACCUMULATE: zero them first.
 */
int robseg_pixel_hist(const int64_t* pred, const int64_t* labels, int n_img, int n_lab_img,
                      int64_t HW, int C, int ignore_index, int64_t* hist, int64_t* hist_total,
                      int64_t* inter, int64_t* tgt, int64_t* prd, robseg_stream_t stream);

/*
 * SEA worst-case accuracy, tools/worse_only.py:396-408: acc[a,n] = sum_c inter / sum_c tgt
 * (fp32, as the reference), worst[n] = min_a acc[a,n].  inter/tgt: [A,N,C] int64.
 */
int robseg_sea_worst_acc(const int64_t* inter, const int64_t* tgt, int A, int N, int C,
                         float* acc_an, float* worst_n, robseg_stream_t stream);

/*
 * Bilinear up-sampling, align_corners=False, of [planes, h, w] fp32 to [planes, H, W] and its
 * adjoint (planes = B*C).  Replaces the consumer's final logit up-sampling
 * nn.functional.interpolate(logits, size=input.shape[2:], mode="bilinear", align_corners=False)
 * (semseg/models/uperforseg.py:416-418, semseg/models/segmenter.py:228), the same call inside
 * its decode head (uperforseg.py:193-198,282-303) and their autograd backward -- SURVEY.md
 * section 8f rank 1.  Exact ratios 2, 4 and 8 have streaming specialisations.  Index/weight arithmetic as ATen's
 * area_pixel_compute_source_index.  The backward is a deterministic gather (no atomics):
 * gin[p,y,x] = sum over the output pixels whose taps include (y,x) of weight * gout.
 */
int robseg_upsample_bilinear_fwd(const float* in, int64_t planes, int h, int w, float* out, int H,
                                 int W, robseg_stream_t stream);
int robseg_upsample_bilinear_bwd(const float* gout, int64_t planes, int H, int W, float* gin, int h,
                                 int w, robseg_stream_t stream);
/*
 * Same adjoint for a gradient that is a [N, C, H, W] VIEW with contiguous rows: plane (n, c)
 * starts at gout + n*batch_stride + c*chan_stride (elements).  The feature-pyramid
 * up-samplings of the decode head (semseg/models/uperforseg.py:282-303) feed torch.cat, so
 * their gradients arrive as channel slices of the concatenated gradient; this reads them in
 * place instead of through a .contiguous() copy.  gin is contiguous [N*C, h, w].
 */
int robseg_upsample_bilinear_bwd_strided(const float* gout, int64_t N, int C, int64_t batch_stride,
                                         int64_t chan_stride, int H, int W, float* gin, int h, int w,
                                         robseg_stream_t stream);

/*
 * HOST functions (every pointer is host memory): the exact arithmetic of the sequential greedy in
 * evalSEA.worst_case_miou (tools/worse_only.py:267-334) -- SURVEY.md section 8f rank 3.
 *
 * robseg_exact_mean_host: statistics.mean(values) -- the correctly rounded mean of the exact sum
 *   (the reference scores every candidate with it, :69-93).
 * robseg_sea_greedy_round_host: ONE round of the greedy over the images in order_host[N] (the caller
 *   shuffles with Python's `random`, as the reference does, :283-285).  cons_ints / cons_unions:
 *   [A,N,C] float64 exact counts; sel[N] current attack per image; run_int / run_union [C] running
 *   sums (re-rounded to float32 wherever the reference rebuilds a tensor from its lists,
 *   :311-316,323-326); *final_miou current value.  All four are updated in place.
 */
int robseg_exact_mean_host(const double* values_host, int64_t n, double* mean_host);
int robseg_sea_greedy_round_host(const double* cons_ints_host, const double* cons_unions_host, int A,
                                 int N, int C, const int32_t* order_host, int32_t* sel_host,
                                 double* run_int_host, double* run_union_host,
                                 double* final_miou_host);

#ifdef __cplusplus
}
#endif
#endif /* ROBSEG_B200_H_ */
