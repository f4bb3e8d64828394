/* cmu_b200.h — C ABI of libcmu_b200.so: the B200 (sm_100a) kernels behind the CM-UNet pretraining / fine-tuning
 * hot path.  Plain pointers and sizes only; no torch types.  This is exactly the surface a binding of the
 * reference's modules would call (see INTEGRATION.md for the ctypes stub).
 *
 * Conventions
 *   - every function returns 0 on success, non-zero on error; the message is in cmu_last_error() (thread local).
 *   - all pointers are DEVICE pointers unless the name starts with h_; `stream` is a cudaStream_t passed as void*.
 *   - activations: NHWC bf16 ("act"), channel counts multiples of 64 (except the 1-channel input image and the
 *     2-channel head output, which stay fp32 in the reference's NCHW layout).
 *   - the caller owns every buffer including workspaces; the library never allocates device memory, keeps no
 *     pointers after a call returns and performs no host synchronisation.
 *   - unsupported shapes / architectures are errors (there is no CPU or library fallback).
 *
 * Each entry point cites the reference code it replaces (paths relative to the reference repository root;
 * CMU = Pretraining/CM-UNet/cmae/models, FT = Finetuning).
 */
#ifndef CMU_B200_H
#define CMU_B200_H

#ifdef __cplusplus
extern "C" {
#endif

/* ---- runtime ------------------------------------------------------------------------------------------- */
const char* cmu_last_error(void);
int cmu_version(void);
int cmu_device_check(void);               /* current device must be sm_100 */
long long cmu_launch_count(void);         /* kernels launched by this library so far (this process) */
int cmu_debug_set(int key, int value);    /* 0: 1 = CUDA-core cross-check path for the conv GEMMs (tests only)
                                             1: 64 = force 64-wide N tiles; 3: 1 = static tile schedule (A/B); 4: 1 = no resident weights (A/B);
                                             5: 1 = no CTA-pair kernel (A/B); 6: 1 = pair kernel with N = 128 only;
                                             7: 1 = 64-wide layers on the 1-CTA kernel; 8: 1 = no kernel-row stacking in the
                                             64-channel wgrad; 9: 1 = one epilogue warp group; 10: 1 = wgrad CTAs of one
                                             pixel range launched as a paced cluster (A/B, profiles/r1_k2_dram_traffic.md);
                                             11: 1 = ConvTranspose on the 1-CTA kernel (A/B); 12: 1 = full-depth TMA ring for
                                             one-tile wgrad CTAs (A/B); 13: unused; 14: KB of shared
                                             memory the K1 kernels leave free per SM, 15: 2 = two-stage K2 ring (A/B:
                                             co-residency of HBM-bound kernels of other streams) */

/* ---- a1  patch-mask generator: CMU/backbones/UNet_encoder.py:106-139 (create_random_patch_mask) ------------
 * d_state: uint32[cmu_mask_state_words()] = MT19937 key[624] + position, same content as numpy's
 * np.random.get_state()[1:3] (legacy global RNG used at UNet_encoder.py:124).  Bit-exact with the reference. */
int cmu_mask_state_words(void);
int cmu_mask_seed(unsigned int* d_state, unsigned int seed, void* stream);             /* == np.random.seed(seed) */
long long cmu_mask_workspace_bytes(int batch, int n_patches, int k_masked);            /* host-only planner */
int cmu_mask_generate(unsigned int* d_state, unsigned char* mask /* (batch,S,S) u8 or NULL: only advance the stream */,
                      int* perm_ws /* cmu_mask_workspace_bytes(batch, (S/ps)^2, K) bytes */,
                      int batch, int img_size, int patch_size, int k_masked, int n_shuffles, void* stream);

/* ---- a2/a3  DoubleConv pieces: CMU/backbones/UNet_encoder.py:18-30,141-158 == FT/model.py:16-26 ----------- */
/* weight packing (fp32 torch layouts -> bf16 GEMM operands); wf / wd may be NULL */
int cmu_pack_conv3x3_weights(const float* w /* (Cout,Cin,3,3) */, int cout, int cin, void* wf /* [9][Cout][Cin] */,
                             void* wd /* [9][Cin][Cout], rotated */, void* stream);
int cmu_pack_convT2x2_weights(const float* w /* (Cin,Cout,2,2) */, int cin, int cout, void* wf /* [4*Cout][Cin] */,
                              void* wd /* [Cin][4*Cout] */, void* stream);
int cmu_cast_f32_to_bf16(const float* x, void* y, long long n, void* stream);

/* first conv, Cin = 1, fused with x * (1 - mask[0]) (UNet_encoder.py:156, quirk Q1); stats_partial:
 * float[cmu_conv3x3_c1_grid()][2][64] per-block (sum, sumsq) of the fp32 outputs, or NULL */
int cmu_conv3x3_c1_grid(void);
int cmu_conv3x3_c1_fprop(const float* x /* (N,H,W) */, const unsigned char* mask0 /* (H,W) or NULL */,
                         const float* w /* (64,1,3,3) */, int cout, void* y /* act (N,H,W,64) */, float* stats_partial,
                         int n, int h, int w_, void* stream);
int cmu_conv3x3_c1_wgrad(const float* x, const unsigned char* mask0, const void* dy, int cout,
                         float* partial /* float[grid][576] */, float* dw /* (64,1,3,3) */, int accumulate, int n, int h,
                         int w_, void* stream);

/* conv3x3 pad 1 as tcgen05 implicit GEMM.  (x0|x1) = channel concat of two act tensors (munet_neck.py:48; x1 may be
 * NULL).  The conv bias is NOT applied (it cancels inside the train-mode BN that always follows; cmu_bn_finalize
 * folds it into running_mean / the eval shift).  stats_partial: float[*stats_grid][2][*stats_bn] per-CTA
 * (sum, sumsq), accumulated in fp32 over the bf16-ROUNDED outputs -- exactly the values y that BatchNorm later
 * normalises (torch's batch_norm also takes its statistics from the stored tensor); they differ from statistics of the
 * fp32 accumulators by the rounding noise only (<= 2e-3 relative on sumsq, tests/kernel_checks.py stats_sq_vs_fp32);
 * cmu_bn_finalize reduces the partial rows in fp64.  Size it with cmu_conv_max_grid() * 2 * max(128, Cout). */
int cmu_conv_max_grid(void);
int cmu_conv3x3_fprop(const void* x0, int c0, const void* x1, int c1, int n, int h, int w, const void* w_packed, int cout,
                      void* y, float* stats_partial, int* h_stats_grid, int* h_stats_bn, void* stream);
int cmu_conv3x3_dgrad(const void* dy, int cout, int n, int h, int w, const void* w_packed_dgrad, void* dx0, int c0,
                      void* dx1, int c1, void* stream);
/* dgrad that also delivers per-CTA column sums of its outputs (partial layout as stats_partial above); cmu_stats_colsum
 * reduces the SUM half over the CTAs for the first c_count channels: the bias gradient of the ConvTranspose2d whose
 * output was x0 (munet_neck.py:46-49) without a separate pass over dx0. */
int cmu_conv3x3_dgrad_sums(const void* dy, int cout, int n, int h, int w, const void* w_packed_dgrad, void* dx0, int c0,
                           void* dx1, int c1, float* sums_partial, int* h_sums_grid, int* h_sums_bn, void* stream);
int cmu_stats_colsum(const float* partial, int grid, int bn_tile, int c_total, int c_count, float* out, void* stream);
long long cmu_conv3x3_wgrad_workspace_bytes(int cin, int cout, int n, int h, int w);
int cmu_conv3x3_wgrad(const void* x0, int c0, const void* x1, int c1, const void* dy, int cout, int n, int h, int w,
                      float* workspace, long long workspace_bytes, float* dw /* (Cout,Cin,3,3) fp32 */, int accumulate,
                      void* stream);

/* BatchNorm2d (train: batch statistics, running stats momentum update with unbiased variance; eval: running stats)
 * -> per-channel scale/shift; then apply + ReLU (+ MaxPool2d(2), UNet_encoder.py:44-49) */
int cmu_bn_finalize(const float* partial, int grid, int bn_tile, int c, double count, const float* gamma, const float* beta,
                    const float* conv_bias, float* running_mean, float* running_var, float momentum, float eps,
                    int training, float* scale, float* shift, float* mean, float* rstd, void* stream);
int cmu_bn_relu_apply(const void* y, const float* scale, const float* shift, void* a, void* pooled /* or NULL */, int n,
                      int h, int w, int c, void* stream);
/* backward of ReLU + BN (+ pool routing): dy from da (+ dpool); sums = float[2][C] = (dbeta, dgamma) */
int cmu_bn_bwd_grid(void);
int cmu_bn_relu_bwd(const void* da, const void* dpool, const void* y, const float* scale, const float* shift,
                    const float* mean, const float* rstd, float* partial /* float[grid][2][C] */, float* sums, void* dy,
                    int n, int h, int w, int c, int training, void* stream);

/* ---- a5  UpBlock pieces: CMU/necks/munet_neck.py:25-49 == FT/model.py:57-81 ------------------------------ */
int cmu_convT2x2_fprop(const void* x, int cin, int n, int h, int w, const void* w_packed, int cout, const float* bias,
                       void* y /* act (N,2H,2W,Cout) */, void* stream);
int cmu_convT2x2_dgrad(const void* dy, int cout, int n, int h, int w, const void* w_packed_dgrad, int cin, void* dx,
                       void* stream);
long long cmu_convT2x2_wgrad_workspace_bytes(int cin, int cout, int n, int h, int w);
int cmu_convT2x2_wgrad(const void* x, int cin, const void* dy, int cout, int n, int h, int w, float* workspace,
                       long long workspace_bytes, float* dw /* (Cin,Cout,2,2) */, int accumulate, void* stream);
int cmu_colsum_bf16(const void* x /* (rows, C) bf16 */, long long rows, int c, float* partial /* float[grid][C] */,
                    float* out /* float[C] */, void* stream);   /* ConvTranspose2d bias gradient */

/* ---- a6  1x1 heads: munet_neck.py:72 (conv_last 64->2), cmunet.py:128-129 (reduce_channels 1024->256) ---- */
int cmu_conv1x1_fprop(const void* x, int cin, int n, int h, int w, const void* w_packed /* bf16 [Cout][Cin] */, int cout,
                      const float* bias, void* y, void* stream);
int cmu_head1x1_fprop(const void* a, const float* w, const float* b, float* out /* (N,2,H,W) fp32 */, int n, int h, int w_,
                      int cin, int cout, void* stream);
int cmu_head1x1_bwd(const void* a, const float* w, const float* dout, void* da, float* acc /* float[130] */, int n, int h,
                    int w_, int cin, int cout, void* stream);
/* decoder tail fused (munet_neck.py:48-49,72,81 == FT/model.py:79-81,131): BatchNorm apply + ReLU of the last 64-channel
 * conv folded into conv_last, forward and backward; the activated tensor `a` and its gradient never reach HBM.
 * y: raw conv output (act, 64 ch); scale/shift/mean/rstd from cmu_bn_finalize; w (2,64), b (2) = conv_last parameters.
 * bwd: partial = float[cmu_bn_relu_head_grid()][258] scratch; sums[2][64] = (dbeta, dgamma) of that BatchNorm;
 * acc[130] = d conv_last.weight (2,64) followed by d conv_last.bias (2); dy = gradient of y (act). */
int cmu_bn_relu_head_grid(void);
int cmu_bn_relu_head_fwd(const void* y, const float* scale, const float* shift, const float* w, const float* b, float* out,
                         int n, int h, int wd, int cin, int cout, void* stream);
int cmu_bn_relu_head_bwd(const void* y, const float* scale, const float* shift, const float* mean, const float* rstd,
                         const float* w, const float* dout, float* partial, float* sums, float* acc, void* dy, int n, int h,
                         int wd, int cin, int cout, int training, void* stream);

/* ---- a7  NonLinearNeck: CMU/necks/nonlinear_neck.py:88-103 ------------------------------------------------ */
long long cmu_sgemm_workspace_bytes(int m, int n, int k);
int cmu_sgemm(const float* a, long long sam, long long sak, const float* b, long long sbn, long long sbk, float* c,
              long long ldc, const float* bias, int m, int n, int k, int accumulate, float* workspace,
              long long workspace_bytes, void* stream);
int cmu_colsum(const float* x, int m, int n, float* out, int accumulate, void* stream);
/* tensor-core path for the big projector linears: out (qc,pc) fp32 = Q^T P (+ bias[pc]) over the rows of
 * Q (rows,qc) bf16 and P (rows,pc) bf16 -- the tcgen05 wgrad engine with the matrix rows as the reduction dim */
long long cmu_gemm_tn_workspace_bytes(int qc, int pc, long long rows);
int cmu_gemm_tn_bf16(const void* q, int qc, const void* p, int pc, long long rows, float* out, const float* bias,
                     int accumulate, float* workspace, long long workspace_bytes, void* stream);
/* fp32 (rows, cols) -> bf16 transposed (cols, rows) [+ optional untransposed bf16 copy y] */
int cmu_transpose_cast_bf16(const float* x, void* yt, void* y, long long rows, long long cols, void* stream);
int cmu_bn1d_stats(const float* x, int m, int c, float* stats /* [2][C] */, void* stream);
int cmu_bn1d_apply(const float* x, const float* stats, double count, int m, int c, const float* gamma, const float* beta,
                   float* running_mean, float* running_var, float momentum, float eps, int training, int relu, float* y,
                   float* mean, float* rstd, void* stream);
int cmu_bn1d_bwd_stats(const float* dy, const float* y, const float* x, const float* mean, const float* rstd, int m, int c,
                       int relu, float* sums, void* stream);
int cmu_bn1d_bwd_apply(const float* dy, const float* y, const float* x, const float* mean, const float* rstd,
                       const float* gamma, const float* sums, double count, int m, int c, int relu, float* dx, void* stream);
int cmu_channel_mean2(const float* x /* (N,2,HW) */, float* y /* (N,HW) */, int n, long long hw, void* stream);
int cmu_channel_mean2_bwd(const float* dx, float* dout, int n, long long hw, void* stream);   /* cmunet.py:126 */
int cmu_nhwc_to_nchw_bf16(const void* x, void* y, int n, int hw, int c, void* stream);        /* cmunet.py:130 */
int cmu_nhwc_to_nchw_f32(const void* x, float* y, int n, int hw, int c, void* stream);

/* ---- a9/a10  CMUNetPretrainHead: CMU/heads/cmunet_head.py:62-91 ------------------------------------------- */
int cmu_masked_mse_fwd(const float* x, const float* pred, long long pred_bstride, const unsigned char* mask, double* acc,
                       float rc_weight, float* loss, int b, int h, int w, void* stream);
int cmu_masked_mse_bwd(const float* x, const float* pred, long long pred_bstride, const unsigned char* mask,
                       const double* acc, const float* gscale, float* dpred, long long dpred_bstride, int b, int h, int w,
                       void* stream);
int cmu_l2_normalize_rows(const float* x, float* y, int rows, int dim, void* stream);
int cmu_infonce_fwd_bwd(const float* q, const float* z, int batch, int n_keys, int dim, int label_offset, float tau,
                        float ct_weight, float* loss_rows, float* loss, float* dq, void* stream);

/* ---- a17  MoCo-v2 queue head (BASELINE configs[3]): Pretraining/MoCo/pl_bolts/models/self_supervised/moco/
 * moco2_module.py:224-270 (logits = [q.k, q.Queue]/T, CE label 0), :160-175 (_dequeue_and_enqueue),
 * moco_data_module.py:65 (spatial mean).  The K x N negative logits are produced by cmu_conv1x1_fprop (queue rows as
 * pixels, normalised queries as weights) and dq_neg = P^T Queue by cmu_gemm_tn_bf16. */
int cmu_spatial_mean(const void* x /* act (N,HW,C) */, float* out /* (N,C) */, int n, int hw, int c, void* stream);
int cmu_spatial_mean_bwd(const float* dout, void* dx, int n, int hw, int c, void* stream);
int cmu_moco_prep(const float* q, const float* k, int n, int d, void* qh16, float* qh, float* qnorm, float* lpos,
                  void* stream);
int cmu_moco_softmax(const void* lt /* bf16 [K][N] */, const float* lpos, int k, int n, float temperature,
                     float* loss_rows, float* loss, float* ppos, void* p /* bf16 [K][N] or NULL */, void* stream);
int cmu_moco_dq(const float* dq_neg, const float* k, const float* qh, const float* qnorm, const float* ppos, int n, int d,
                float temperature, float* dq, void* stream);
/* fused form (the one Moco_v2 uses): logits never reach HBM.  cmu_conv1x1_fprop_exp stores E[k][n] = exp((Queue_k . q_n - 1)/T)
 * (bf16, the operand of the backward GEMM dq = E^T Queue) and per-CTA partial column sums; cmu_moco_finish turns them into
 * the loss, the probability of the positive and the per-query scale 1/(Z N T); cmu_moco_dq_scaled applies that scale. */
int cmu_conv1x1_fprop_exp(const void* x, int cin, int n, int h, int w, const void* w_packed, int cout, float inv_temperature,
                          void* y, float* sums_partial, int* sums_grid, int* sums_bn, void* stream);
int cmu_moco_finish(const float* sums_partial, int sums_grid, int sums_bn, int n, const float* lpos, float temperature,
                    float* loss_rows, float* loss, float* ppos, float* row_scale, void* stream);
int cmu_moco_dq_scaled(const float* dq_neg_unnorm, const float* row_scale, const float* k, const float* qh,
                       const float* qnorm, const float* ppos, int n, int d, float temperature, float* dq, void* stream);
int cmu_queue_enqueue(const float* keys, int n, int d, int k, int ptr, void* queue_rows /* bf16 [K][D] */,
                      float* queue_ref /* fp32 (D,K) reference layout or NULL */, void* stream);

/* ---- a12/a16  EMA (cmunet.py:78-92) and AdamW (configs/cmunet_config.py:76-91) ---------------------------- */
int cmu_ema_chunks(const long long* d_table /* [n][3] = dst, src, count */, int n_chunks, float momentum, void* stream);
int cmu_adamw_chunks(const long long* d_table /* [n][6] = p, g, m, v, count, decay */, int n_chunks, float lr, float beta1,
                     float beta2, float eps, float weight_decay, int step, float grad_scale, void* stream);
/* dynamic loss scaling of mmengine's AmpOptimWrapper(loss_scale='dynamic') (cmunet_config.py:76-78 == torch.amp.GradScaler):
 * d_amp = int32[5] device state {float scale, int found_inf, int optimizer steps taken, int growth tracker, float 1/scale
 * of the current gradients}.  cmu_adamw_chunks_amp = found-inf reduction over all gradients + scale update + AdamW that
 * unscales the gradients and leaves parameters and moments untouched after an overflow; no host synchronisation. */
int cmu_amp_init(int* d_amp, float init_scale, void* stream);
int cmu_adamw_chunks_amp(const long long* d_table, int n_chunks, float lr, float beta1, float beta2, float eps,
                         float weight_decay, int* d_amp, float growth, float backoff, int growth_interval, void* stream);
/* SGD with momentum (torch.optim.SGD semantics; MoCo-v2 optimizer, MOCO/moco2_module.py configure_optimizers) */
int cmu_sgd_chunks(const long long* d_table /* [n][4] = p, g, momentum buffer, count */, int n_chunks, float lr,
                   float momentum, float weight_decay, int first_step, void* stream);

/* ---- f3  GPU data pipeline: Pretraining/CM-UNet/cmae/datasets/cmunet_dataset.py:74-88 -------------------------
 * cmu_pil_resize_bicubic == `Image.fromarray(plane[y0:y0+h, x0:x0+w]).resize((out_w, out_h), Image.BICUBIC)` for n planes,
 * bit-exact with Pillow for dtype 0 (uint8, mode "L") and 1 (float32, mode "F") -- replaces cmunet_dataset.py:77-78 and
 * the crop + resize of RandomResizedCrop (cmae/datasets/pipelines/processing.py:570-590).  d_boxes: int32[n][4] =
 * x0, y0, w, h or NULL (whole plane); tmp: n * src_h * out_w elements of the same dtype.
 * cmu_aug_shift_flip_noise == RandomFlip decision + ShiftPixel (processing.py:97-121) + GaussNoise
 * (auto_augment.py:1148-1154): img = crop at (0, 0), img_t = crop at (ph, pw) + max(crop)/10 * noise, cast back to the
 * image dtype like `np.array(out, dtype=img.dtype)`; both returned as float32 model inputs.  d_params: int32[n][4] =
 * flip, ph, pw, 0.  noise: explicit float64 N(0,1) field [n][crop][crop] (exact parity) or NULL -> Philox4x32-10(seed). */
int cmu_pil_resize_bicubic(const void* src, int dtype, int n, int src_h, int src_w, const int* d_boxes, void* tmp,
                           void* dst, int out_h, int out_w, void* stream);
int cmu_aug_shift_flip_noise(const void* src, int dtype, int n, int src_h, int src_w, const int* d_params,
                             const double* noise, unsigned long long seed, int crop, float* img, float* img_t,
                             void* stream);

/* ---- a15  fine-tuning losses: FT/metrics.py:135-220,503-504 ----------------------------------------------- */
int cmu_seg_losses(const float* logits, const double* gt, double* acc /* double[4] */, double* out /* dice, iou, ce */,
                   float* dlogits /* or NULL */, const float* gscale, int n, int h, int w, double dice_eps, double beta,
                   double iou_eps, void* stream);
/* soft-clDice evaluation metric (FT/metrics.py:401-492, used by FT/train.py:464): prediction = [logit1 > logit0], target =
 * gt[:,1] (float64), soft skeletons by min/max-pool morphology with num_iter iterations (the reference fixes 10);
 * out[0] = 1 - 2 tprec tsens / (tprec + tsens), float64. */
long long cmu_soft_cldice_workspace_bytes(int n, int h, int w);
int cmu_soft_cldice(const float* logits, const double* gt, int n, int h, int w, int num_iter, double smooth, void* ws,
                    long long ws_bytes, double* out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* CMU_B200_H */
