/*
 * e2enet_b200.h -- C ABI of the B200-native E2ENet hot-path library (libe2enet_b200.so).
 *
 * Plain C, no torch types: device pointers, sizes and a cudaStream_t (passed as void*).
 * Every entry point returns 0 on success or a negative error code; e2e_last_error()
 * returns a human readable message for the calling thread.  All buffers (incl. scratch)
 * are owned by the caller; the library only caches TMA tensor maps keyed by pointer+shape.
 * Entry points only enqueue work on `stream`; none synchronises.
 *
 * Activation layout ("C8"): bf16 [B][C/8][D][H][W][8] -- channel-blocked so that one
 * voxel's 8 channels are one 16-byte vector.  A [rows][8ch] slab of it is at the same
 * time a K-major UMMA operand (fwd / dgrad: K = channels) and an MN-major one (wgrad:
 * K = voxels), and a TMA box of it can be shifted per 8-channel block along D, which is
 * how the reference's depth shift (unetpp_d.py:45-59) is folded into the operand fetch.
 *
 * Reference interfaces replaced (file:line in boqian333/E2ENet-Medical):
 *   e2e_gather_gemm     torch_shift.forward + Conv3d fwd / dgrad, ConvTranspose3d fwd / dgrad,
 *                       1x1x1 seg heads      e2enet/network_architecture/unetpp_d.py:45-59,102-108,394-401,521-522
 *   e2e_gather_wgrad    Conv3d / ConvTranspose3d weight gradients (autograd of the above)
 *   e2e_in_*            InstanceNorm3d(affine) + LeakyReLU fwd/bwd      unetpp_d.py:99-100,111
 *   e2e_maxpool_*       MaxPool3d (down* modules)                       unetpp_d.py:523-524
 *   e2e_pack_weights    `w.data * mask` on weight load                  core_channel.py:427-434
 *   e2e_mask_*          Masking.apply_mask / kernel_death / kernel_growth / counts
 *                       e2enet/training/network_training/sparselearning/core_channel.py:338-347,427-434,647-666,721-739,861-876
 *   e2e_window_*        softmax + mirror + gaussian + `agg[:,tile] += ...`, `agg /= nb`, argmax
 *                       e2enet/network_architecture/neural_network.py:370-407,529-563
 */
#ifndef E2ENET_B200_H
#define E2ENET_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define E2E_MAX_SRC 4

#define E2E_OK 0
#define E2E_ERR_ARG -1
#define E2E_ERR_CUDA -2
#define E2E_ERR_UNSUPPORTED -3

const char* e2e_last_error(void);
int e2e_version(void);
/* "bf16" (libe2enet_b200.so, the default) or "fp16" (libe2enet_b200_fp16.so: the same sources compiled with -DE2E_FP16 --
 * activations, gradients and packed weights are IEEE fp16, the arithmetic of the reference's torch.cuda.amp loop,
 * nnUNetTrainer_simple.py:552-557).  Every "bf16" in this header reads "the library's 16-bit type". */
const char* e2e_precision(void);
/* number of kernels launched by this library in this process so far (bench: gpu_launches) */
long long e2e_launch_count(void);

/* ---------------------------------------------------------------- plan-driven gather GEMM */

/* 8 input channels (one C8 block) of source `src`, fetched at spatial offset (dd,dh,dw) */
typedef struct { int32_t src, blk, dd, dh, dw; } e2e_centry_t;
/* one filter tap: additional spatial offset applied to every channel entry */
typedef struct { int32_t dd, dh, dw; } e2e_tap_t;
/* one 8-column block of the GEMM result: where it is stored */
typedef struct { int32_t dst, blk, chmask, od, oh, ow; } e2e_colblk_t;

/*
 * out[o, n] = sum_{e, t, j}  src[cent[e].src][b, cent[e].blk, (o*is + iv + cent[e].off + tap[t].off), j]
 *                            * wpacked[e/2][t][e%2][n][j]
 * for every voxel o of the iteration grid (B, Do, Ho, Wo); out-of-range reads are 0.
 * Column block q (columns 8q..8q+7) goes to dst[cols[q].dst], channel block cols[q].blk,
 * voxel o*os + cols[q].off (skipped when outside the destination grid), channels selected
 * by cols[q].chmask.
 */
typedef struct {
  int32_t B, Di, Hi, Wi;       /* grid of every source tensor */
  int32_t Do, Ho, Wo;          /* iteration grid (GEMM M = B*Do*Ho*Wo) */
  int32_t isd, ish, isw;       /* input stride */
  int32_t ivd, ivh, ivw;       /* input offset of this call (variant) */
  int32_t Dd, Hd, Wd;          /* grid of every destination tensor */
  int32_t osd, osh, osw;       /* output stride */
  int32_t n_src;
  const void* src[E2E_MAX_SRC];      /* device, bf16 C8 */
  int32_t src_cb[E2E_MAX_SRC];       /* channel blocks of each source */
  int32_t n_cent;                    /* even */
  const e2e_centry_t* cents;         /* device */
  int32_t n_taps;
  const e2e_tap_t* taps;             /* device */
  const void* wpacked;               /* device bf16 [n_cent/2][n_taps][2][Npad][8] */
  int32_t Npad;                      /* multiple of 16 */
  const e2e_colblk_t* cols;          /* device, Npad/8 entries */
  int32_t n_dst;
  void* dst[E2E_MAX_SRC];
  int32_t dst_cb[E2E_MAX_SRC];       /* channel blocks (out_mode 0) or channel count (out_mode 1) */
  int32_t out_mode;                  /* 0: bf16 C8   1: fp32 NCDHW (dst[0], column n -> channel n) */
  int32_t impl;                      /* 0: mma.sync gather kernel   1: tcgen05/TMA kernel (halo / 1-tap forms) */
  int32_t col_bounds;                /* bit 0 / 1 / 2: a column block's destination depth / row / column can fall
                                        outside the destination grid and must be checked (7 is always safe) */
  /* InstanceNorm statistics fused into the epilogue (tcgen05 path, out_mode 0 only; may be null): per-slot partial
   * sums of the bf16-rounded result, stats[slot][b][{sum, sum of squares}][stats_ctot] fp32, channel index =
   * destination channel; slots = e2e_gather_gemm_stats_slots() (one per CTA of the launch; every slot is
   * written completely, no atomics: bit-reproducible).  Reduce with e2e_in_stats_final().  Replaces the separate
   * statistics pass over the conv output of ConvDropoutNormNonlin (unetpp_d.py:108-110). */
  float* stats;
  int32_t stats_ctot;
  /* tcgen05 path, out_mode 0 only: bit i set = the result is ADDED to what destination i already holds
   * (bf16(fp32(old) + fp32 accumulator)); the other destinations are overwritten.  The gradient fan-in of an
   * activation with several consumers (the fusion grid, unetpp_d.py:453-478) is summed in the data-gradient
   * epilogue instead of a separate add pass */
  int32_t accumulate;
} e2e_gemm_t;

int e2e_gather_gemm(const e2e_gemm_t* p, void* stream);
/* p[0..n): column chunks of one GEMM -- identical except wpacked / cols / Npad (a result wider than
 * 256 columns is split by the host plan); one kernel launch on the tcgen05 path */
int e2e_gather_gemm_multi(const e2e_gemm_t* p, int32_t n, void* stream);
/* slots the launch e2e_gather_gemm_multi(p, n) would write into p->stats; 0 when that launch cannot fuse the
 * statistics (mma.sync path, fp32 output): then run e2e_in_stats on the result instead */
int e2e_gather_gemm_stats_slots(const e2e_gemm_t* p, int32_t n);
/* 1 when e2e_gather_gemm_multi(p, n) is served by a tcgen05 kernel (then `stats` / `accumulate` are honoured) */
int e2e_gather_gemm_on_tcgen05(const e2e_gemm_t* p, int32_t n);

/*
 * dwp[e/2][t][e%2][n][j] += sum_o grad[b, n/8, o, n%8] * src[...][(o*is + iv + cent[e].off + tap[t].off), j]
 * (fp32, atomically accumulated: the caller zeroes dwp first).
 */
typedef struct {
  int32_t B, Di, Hi, Wi;
  int32_t Do, Ho, Wo;
  int32_t isd, ish, isw;
  int32_t ivd, ivh, ivw;
  int32_t n_src;
  const void* src[E2E_MAX_SRC];
  int32_t src_cb[E2E_MAX_SRC];
  int32_t n_cent;
  const e2e_centry_t* cents;
  int32_t n_taps;
  const e2e_tap_t* taps;
  const void* grad;                  /* device bf16 C8 on the iteration grid */
  int32_t grad_cb;
  int32_t Npad;                      /* multiple of 16, <= 8*grad_cb rounded up */
  float* dwp;                        /* device fp32 [n_cent/2][n_taps][2][Npad][8] (may be null in direct mode) */
  int32_t impl;
  /* direct mode (tcgen05 path only, see e2e_gather_wgrad_direct_ok): when grad_out is non-null the kernel adds its
   * result straight into the gradient in the reference's PARAMETER layout, grad_out[rowoff[n] + centoff[e*8+j] +
   * tapoff[t]] (the scatter of e2e_unpack_wgrad; the caller zeroes grad_out first), and dwp is not touched */
  float* grad_out;
  const int32_t* rowoff;             /* device, Npad entries (negative: padded column) */
  const int32_t* centoff;            /* device, n_cent*8 entries (negative: padded channel) */
  const int32_t* tapoff;             /* device, n_taps entries */
} e2e_wgrad_t;

int e2e_gather_wgrad(const e2e_wgrad_t* p, void* stream);
/* 1 when e2e_gather_wgrad(p) runs on a tcgen05 kernel and therefore honours grad_out */
int e2e_gather_wgrad_direct_ok(const e2e_wgrad_t* p);

/*
 * wpacked[e/2][t][e%2][n][j] = bf16( w[rowoff[n] + centoff[e*8+j] + tapoff[t]] * (mask ? mask[same] : 1) )
 * (0 where rowoff or centoff is negative).  This is where the DSFF mask meets the weights.
 * emask / rclass (both may be null): K entry e only feeds the columns whose class rclass[n] (0..31) has its bit set in
 * emask[e]; other (e, n) pairs pack 0.  (The stride-2 data gradient computes all four output parities in one GEMM:
 * an entry that reads d(raw) one row / column further only reaches the odd output rows / columns.)
 */
int e2e_pack_weights(const float* w, const float* mask, const int32_t* rowoff, const int32_t* centoff,
                     const int32_t* tapoff, const int32_t* emask, const int32_t* rclass, int32_t n_cent, int32_t n_taps,
                     int32_t Npad, void* wpacked, void* stream);
/* the same for many (plan, weight) pairs in one launch: jobs is a DEVICE array; job j packs items
 * [item_begin, item_begin + n_cent*n_taps*Npad) (one item = 8 bf16 = 16 bytes of `out`) */
typedef struct {
  const float* w;
  const float* mask;              /* may be null */
  const int32_t* rowoff;
  const int32_t* centoff;
  const int32_t* tapoff;
  int32_t n_cent, n_taps, Npad, pad_;
  void* out;
  int64_t item_begin;
  const int32_t* emask;           /* see e2e_pack_weights; may be null */
  const int32_t* rclass;
} e2e_pack_job_t;
int e2e_pack_weights_multi(const e2e_pack_job_t* jobs, int32_t n_jobs, int64_t total_items, void* stream);
/* grad[rowoff[n] + centoff[e*8+j] + tapoff[t]] = dwp[...]  (inverse scatter; every weight appears once) */
int e2e_unpack_wgrad(const float* dwp, const int32_t* rowoff, const int32_t* centoff, const int32_t* tapoff,
                     int32_t n_cent, int32_t n_taps, int32_t Npad, float* grad, void* stream);

/* ---------------------------------------------------------------- layout conversion */
int e2e_nc_to_c8(const float* x, void* y, int32_t B, int32_t C, int64_t V, void* stream);     /* fp32 NCDHW -> bf16 C8 (C padded to 8) */
int e2e_c8_to_nc(const void* x, float* y, int32_t B, int32_t C, int64_t V, void* stream);     /* bf16 C8 -> fp32 NCDHW */

/* stand-alone depth shift on a plain NCDHW tensor of 2- or 4-byte elements (torch_shift.forward, unetpp_d.py:45-59):
 * y[b,c,d,:] = x[b,c,d - sign*s_c,:] with zero fill, s_c = c / ceil(C/shift_size) - shift_size/2; sign = -1 is the
 * gradient.  (Inside the shift-conv the shift is folded into the operand fetch; this serves direct module calls.) */
int e2e_shift_depth(const void* x, void* y, int32_t elem_bytes, int32_t B, int32_t C, int32_t D, int64_t HW,
                    int32_t shift_size, int32_t sign, void* stream);

/* ---------------------------------------------------------------- InstanceNorm + LeakyReLU */
/* per (b, c): mean and rstd of raw over V voxels; partial: scratch fp32 [B*Cb][nchunk][16] */
int e2e_in_stats(const void* raw, int32_t B, int32_t Cb, int64_t V, float eps, float* partial, int32_t nchunk,
                 float* mean, float* rstd, void* stream);
/* mean / rstd from the per-slot partials a fused conv epilogue wrote (e2e_gemm_t.stats): fixed-order fp64 reduce */
int e2e_in_stats_final(const float* stats, int32_t n_slots, int32_t B, int32_t C, int64_t V, float eps, float* mean,
                       float* rstd, void* stream);
int e2e_in_apply(const void* raw, const float* mean, const float* rstd, const float* gamma, const float* beta,
                 float slope, int32_t B, int32_t Cb, int64_t V, void* out, void* stream);
/* e2e_in_stats_final + e2e_in_apply in ONE launch: every block forms its plane's mean / rstd from the epilogue slots
 * (stats[n_slots][B][2][Cb*8]) in the same fixed order; they are also stored to mean_out / rstd_out for the backward */
int e2e_in_apply_from_slots(const void* raw, const float* stats, int32_t n_slots, float eps, const float* gamma,
                            const float* beta, float slope, int32_t B, int32_t Cb, int64_t V, float* mean_out, float* rstd_out,
                            void* out, void* stream);
/* backward: sums[b][c] = {sum dz, sum dz*xhat}; then draw, dgamma, dbeta, dbias.
 * partial: scratch fp32 [B*Cb][nchunk][16] (pass 1) followed by [B*Cb][nchunk][8] (pass 2) = 24*B*Cb*nchunk floats */
int e2e_in_bwd(const void* dy, const void* raw, const float* mean, const float* rstd, const float* gamma,
               const float* beta, float slope, int32_t B, int32_t Cb, int64_t V, float* partial, int32_t nchunk,
               float* sums, void* draw, float* dgamma, float* dbeta, float* dbias, void* stream);

/* InstanceNorm + LeakyReLU fused with the MaxPool3d(kernel == stride) that consumes the same activation
 * (down* modules, unetpp_d.py:453-478,523-524): writes out (full resolution), pooled and its arg-max in one
 * pass; the backward folds the pooled gradient dyp into the norm backward (dy may be null = no other
 * consumer).  Requires the window to divide the grid and kd*kh*kw <= 8. */
int e2e_in_apply_pool(const void* raw, const float* mean, const float* rstd, const float* gamma, const float* beta,
                      float slope, int32_t B, int32_t Cb, int32_t D, int32_t H, int32_t W, int32_t kd, int32_t kh,
                      int32_t kw, void* out, void* pooled, uint8_t* argmax, void* stream);
int e2e_in_apply_pool_from_slots(const void* raw, const float* stats, int32_t n_slots, float eps, const float* gamma,
                                 const float* beta, float slope, int32_t B, int32_t Cb, int32_t D, int32_t H, int32_t W,
                                 int32_t kd, int32_t kh, int32_t kw, float* mean_out, float* rstd_out, void* out, void* pooled,
                                 uint8_t* argmax, void* stream);
/* partial: 24*B*Cb*nchunk floats, as e2e_in_bwd */
int e2e_in_bwd_pool(const void* dy, const void* dyp, const uint8_t* argmax, const void* raw, const float* mean,
                    const float* rstd, const float* gamma, const float* beta, float slope, int32_t B, int32_t Cb,
                    int32_t D, int32_t H, int32_t W, int32_t kd, int32_t kh, int32_t kw, float* partial, int32_t nchunk,
                    float* sums, void* draw, float* dgamma, float* dbeta, float* dbias, void* stream);

/* plane-resident backward (one kernel, 3 HBM passes instead of 5): a group of co-resident CTAs owns one (b, cb) plane
 * at a time, reduces sum dz / sum dz*xhat, synchronises on a counter, and re-reads its slice of dy / raw from the L2
 * to write draw.  dyp / argmax null = plain backward, else the pooled gradient is folded in (window (1,2,2) or
 * (2,2,2) dividing the grid).  scratch: e2e_in_bwd_scratch_floats(B, Cb, D*H*W) floats; results as e2e_in_bwd. */
int64_t e2e_in_bwd_scratch_floats(int32_t B, int32_t Cb, int64_t V);
int e2e_in_bwd_fused(const void* dy, const void* dyp, const uint8_t* argmax, const void* raw, const float* mean,
                     const float* rstd, const float* gamma, const float* beta, float slope, int32_t B, int32_t Cb,
                     int32_t D, int32_t H, int32_t W, int32_t kd, int32_t kh, int32_t kw, float* scratch, float* sums,
                     void* draw, float* dgamma, float* dbeta, float* dbias, void* stream);

/* ---------------------------------------------------------------- MaxPool3d (kernel == stride) */
int e2e_maxpool_fwd(const void* x, void* y, uint8_t* argmax, int32_t BCb, int32_t D, int32_t H, int32_t W,
                    int32_t kd, int32_t kh, int32_t kw, void* stream);
int e2e_maxpool_bwd(const void* dy, const uint8_t* argmax, void* dx, int32_t BCb, int32_t D, int32_t H, int32_t W,
                    int32_t kd, int32_t kh, int32_t kw, void* stream);
/* y += x  (gradient fan-in of an activation with several consumers) */
int e2e_add_inplace(void* y, const void* x, int64_t n_elems, void* stream);

/* ---------------------------------------------------------------- DSFF Masking */
/* multi-tensor apply_mask: w[i] *= m[i]; mom[i] *= m[i] (mom may be null); ptr tables are device arrays */
int e2e_mask_apply_multi(float* const* w, float* const* mom, const float* const* mask, const int64_t* numel,
                         int32_t n_tensors, int64_t max_numel, void* stream);
/* kernel L1: l1[a*C1+b] = nested fp32 sums over (kd,kh,kw <= 4) of |w| with the reference's association:
 * assoc 0 = torch CPU (left to right), assoc 1 = torch CUDA reduce order ((a0+a2)+a1 for 3 values) -- the reference's
 * Masking computes it on the GPU, so 1 is what reproduces its prune sets bit for bit */
int e2e_mask_kernel_l1(const float* w, int32_t n_kernels, int32_t kd, int32_t kh, int32_t kw, int32_t assoc, float* l1,
                       void* stream);
/* k-th smallest (0-based rank) of l1[0..n): exact radix select; result -> *thr (device) */
int e2e_mask_kth(const float* l1, int32_t n, int32_t rank, float* thr, uint32_t* scratch, void* stream);
/* mask[kernel,:] = 0 for l1 <= *thr; counts[0] = kernels alive after; counts[1] = dead kernels after */
int e2e_mask_kill(const float* l1, const float* thr, float* mask, int32_t n_kernels, int32_t ksize,
                  int32_t* counts, void* stream);
/* row-major ordered list of dead kernels (mask row sums < 1): dead[0..*n_dead) */
int e2e_mask_dead_list(const float* mask, int32_t n_kernels, int32_t ksize, int32_t* dead, int32_t* n_dead,
                       int32_t* scratch, void* stream);
/* mask[dead[pick[i]], :] = 1 */
int e2e_mask_grow(float* mask, const int32_t* dead, const int32_t* pick, int32_t n_pick, int32_t ksize, void* stream);
/* nnz[0] = #(mask != 0); fired |= mask; nnz[1] = #(fired != 0) */
int e2e_mask_counts(const float* mask, uint8_t* fired, int64_t numel, int32_t* nnz, void* stream);

/* ---------------------------------------------------------------- fused optimizer step */
/*
 * clip_grad_norm_ + Nesterov SGD + weight decay + Masking.apply_mask over pointer tables
 * (nnUNetTrainer_simple.py:367-371,560-564; core_channel.py:427-434).  `tensors` is a DEVICE array; hyper is a
 * DEVICE float[5] = {lr, momentum, weight_decay, max_norm (<= 0: no clipping), grad_scale (multiplies every
 * gradient first, e.g. 1/world_size or a GradScaler's inverse scale)} so captured graphs follow LR schedules.
 *   e2e_sgd_clip_coef: norm_coef (float[4]): [0] = ||m * g||_2 over all tensors with m = grad_scale (/ loss scale),
 *                      [1] = min(1, max_norm / (norm + 1e-6)), [2] = 1 if the norm is not finite, [3] = m;
 *                      partial: scratch of e2e_sgd_partial_count() floats.  scaler: null, or DEVICE float[5] =
 *                      {loss scale, clean steps, growth interval, backoff factor, growth factor} -- the reference's
 *                      GradScaler (nnUNetTrainer_simple.py:553-562) as device state: gradients are un-scaled on the
 *                      fly, a non-finite norm skips the update and multiplies the scale by the backoff factor,
 *                      `growth interval` clean steps multiply it by the growth factor.
 *   e2e_sgd_update:    g' = g * m * coef + wd * p;  buf = momentum * buf + g';
 *                      p = (p - lr * (nesterov ? g' + momentum * buf : buf)) * mask;  buf *= mask   (mask may be null);
 *                      skipped entirely when norm_coef[2] != 0.  Gradients are left unscaled in memory.
 */
typedef struct {
  float* p;
  const float* g;
  float* mom;                      /* momentum buffer (zero-initialised before the first step) */
  const float* mask;               /* DSFF mask of this tensor or null */
  int64_t numel;
} e2e_sgd_tensor_t;
int e2e_sgd_partial_count(int32_t n_tensors, int64_t max_numel);
int e2e_sgd_clip_coef(const e2e_sgd_tensor_t* tensors, int32_t n_tensors, int64_t max_numel, const float* hyper,
                      float* scaler, float* partial, float* norm_coef, void* stream);
int e2e_sgd_update(const e2e_sgd_tensor_t* tensors, int32_t n_tensors, int64_t max_numel, const float* hyper,
                   const float* norm_coef, int32_t nesterov, void* stream);

/* ---------------------------------------------------------------- deep-supervision loss statistics */
/*
 * logits fp32 [B][C][V], target fp32 [B][V] (class index as float, like the reference's target[:, 0]):
 * stats[b][c] = {sum_v p, sum_v p*[y==c], #{y==c}} and *ce_sum = sum -log p[y] are ACCUMULATED (caller
 * zeroes them) through per-block partials reduced in a fixed order -- no atomics, bit-reproducible: a one-ulp
 * difference here is amplified by the bf16 backward pass into ~1 % differences of the deepest weight gradients.
 * Replaces the softmax / one-hot / tp-fp-fn / log-softmax / nll passes of
 * e2enet/training/loss_functions/dice_loss.py:100-190,302-359 and crossentropy.py:4-11.
 */
int e2e_softmax_stats_partial_count(int32_t B, int32_t C, int64_t V);      /* floats of scratch `partial` */
int e2e_softmax_stats_fwd(const float* logits, const float* target, int32_t B, int32_t C, int64_t V,
                          float* partial, float* stats, float* ce_sum, void* stream);
/* DC_and_CE_loss from the statistics in one tiny launch (dice_loss.py:155-190, 302-359): loss[0] = weight_ce * ce_sum /
 * n_vox - weight_dice * mean dc, dc = (2 tp + smooth) / (S_p + S_y + smooth + 1e-8) over (b, c >= (do_bg ? 0 : 1))
 * (batch_dice: statistics summed over b first), and the coefficients of its gradient: gsp = dloss/dS_p [B][C],
 * gtp = dloss/dtp [B][C], gce = dloss/dce_sum [1].  n_vox = B * V. */
int e2e_dc_ce_from_stats(const float* stats, const float* ce_sum, int32_t B, int32_t C, int64_t n_vox, float smooth,
                         int32_t do_bg, int32_t batch_dice, float weight_ce, float weight_dice, float* loss, float* gsp,
                         float* gtp, float* gce, void* stream);
/* dlogits = s * (p_k (g_k - sum_j p_j g_j) + gce (p_k - [y==k])),  g_j = gsp[b][j] + gtp[b][j] [y==j]; gce: device scalar;
 * s = gscale[0] (device scalar: the upstream gradient of the loss, e.g. a GradScaler's scale) or 1 when gscale is null */
int e2e_softmax_stats_bwd(const float* logits, const float* target, const float* gsp, const float* gtp,
                          const float* gce, const float* gscale, int32_t B, int32_t C, int64_t V, float* dlogits,
                          void* stream);

/* ---------------------------------------------------------------- sliding-window accumulate */
/*
 * logits fp32 [ncls][px][py][pz] of one tile (possibly predicted on a flipped input: flip bit0=x,
 * bit1=y, bit2=z) -> softmax over classes (skipped when apply_softmax == 0: input already holds
 * the network's inference non-linearity), un-flip, * scale (1/num_mirrors) * gauss[px][py][pz]
 * (gauss may be null = 1), += into agg[ncls][X][Y][Z] at (x0,y0,z0); if add_weight, wsum[X][Y][Z] += gauss.
 */
int e2e_window_accumulate(const float* logits, const float* gauss, float* agg, float* wsum, int32_t ncls,
                          int32_t px, int32_t py, int32_t pz, int32_t X, int32_t Y, int32_t Z,
                          int32_t x0, int32_t y0, int32_t z0, int32_t flip, float scale, int32_t add_weight,
                          int32_t apply_softmax, void* stream);
/* the same with the network's last layer folded in: x = bf16 C8 [Cb][px][py][pz][8] feature map of one tile,
 * w = fp32 [ncls][C] weight of the 1x1x1 seg head (unetpp_d.py:394-401; rounded to bf16 like the GEMM operand):
 * logits = w . x per voxel on the CUDA cores, then softmax / un-flip / scale / gauss / += as above.  The fp32
 * logits tensor of the tile is never written or read. */
int e2e_window_head_accumulate(const void* x, int32_t Cb, const float* w, int32_t C, const float* gauss, float* agg,
                               float* wsum, int32_t ncls, int32_t px, int32_t py, int32_t pz, int32_t X, int32_t Y,
                               int32_t Z, int32_t x0, int32_t y0, int32_t z0, int32_t flip, float scale,
                               int32_t add_weight, void* stream);
/* agg /= wsum (in place, cropped region), seg = argmax over classes (first max wins) */
int e2e_window_finalize(float* agg, const float* wsum, int32_t ncls, int32_t X, int32_t Y, int32_t Z,
                        int64_t* seg, void* stream);
/* the same for the x-planes [x0, x1) only: the tiled predictor finalises (and starts copying to the host) the planes
 * that no later tile touches while the remaining tiles are still being computed (neural_network.py:396-426) */
int e2e_window_finalize_range(float* agg, const float* wsum, int32_t ncls, int32_t X, int32_t Y, int32_t Z, int32_t x0,
                              int32_t x1, int64_t* seg, void* stream);

/* ---------------------------------------------------------------- export: resample + argmax */
/*
 * probs fp32 [C][X][Y][Z] -> resampled to (Xo, Yo, Zo) with the pixel-centre map in = (out + 0.5) * n_in / n_out - 0.5
 * and edge clamping, per axis nearest (mode 0 = interpolation order 0) or linear (mode 1 = order 1); writes the
 * resampled probabilities (out_probs, may be null) and / or their arg-max over classes as uint8 labels (first
 * maximum wins, may be null).  Replaces resample_data_or_seg(is_seg=False) + argmax(0) of
 * e2enet/inference/segmentation_export.py:84-115 / e2enet/preprocessing/preprocessing.py:113-202.
 */
int e2e_resample_argmax(const float* probs, int32_t C, int32_t X, int32_t Y, int32_t Z, int32_t Xo, int32_t Yo, int32_t Zo,
                        int32_t mode_x, int32_t mode_y, int32_t mode_z, float* out_probs, uint8_t* labels, void* stream);

#ifdef __cplusplus
}
#endif
#endif
