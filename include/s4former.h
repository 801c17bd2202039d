/* s4former_b200 -- C ABI of the B200-native S4Former train-step kernels.
 *
 * Drop-in boundary: the reference is pure Python over PyTorch
 * (mmseg/mmcv; no native code, setup.py:200 `ext_modules=[]`), so what a maintainer binds is
 * this library from `torch.autograd.Function`s via ctypes.  Every entry
 * point takes raw device pointers (`tensor.data_ptr()`), plain sizes and a `cudaStream_t`;
 * nothing allocates, nothing throws: the return value is 0 or a negative error code and
 * `s4_last_error()` describes the failure.  Workspace sizes come from companion
 * `*_workspace()` functions.  All kernels are built for sm_100a only.
 *
 * dtype codes: 0 = float32, 1 = bfloat16.  Parameters, statistics, losses and logits are
 * always float32; `dtype` selects the activation storage type.
 *
 * Each group cites the reference lines it replaces (paths relative to the reference root).
 */
#ifndef S4FORMER_H_
#define S4FORMER_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#ifndef __CUDA_RUNTIME_H__
typedef struct CUstream_st* cudaStream_t;
#endif

#define S4_DTYPE_F32 0
#define S4_DTYPE_BF16 1

#define S4_ACT_NONE 0
#define S4_ACT_GELU 1

#define S4_BACKEND_AUTO 0  /* tcgen05 when the shape/dtype allows, else CUDA cores */
#define S4_BACKEND_SIMT 1  /* CUDA-core fp32-accumulate path (validation)           */
#define S4_BACKEND_TC 2    /* tcgen05/TMEM/TMA path, error if unsupported           */

/* ---- library ---------------------------------------------------------------------------- */
int s4_version(void);
int s4_built_arch(void);            /* 100 = sm_100a */
const char* s4_last_error(void);
long long s4_launch_count(void);     /* kernels launched by this library in this process */
/* Optional device timing per kernel family (bench.py's roofline leg): while enabled every
 * family brackets its launches with a CUDA event pair on the launching stream and books its
 * algorithmic work (unit 0 = FLOP, 1 = byte; 0 work = time only).  Reading synchronises. */
int s4_prof_enable(int on);          /* turning on clears the totals */
int s4_prof_num_kinds(void);
int s4_prof_get(int idx, char* name, int name_len, int* unit, double* ms, double* work,
                long long* launches);

/* ---- GEMM with fused epilogue -------------------------------------------------------------
 * Replaces: nn.Linear / in_proj / out_proj / FFN inside mmcv MultiheadAttention and FFN as
 * called from mmseg/models/backbones/vit.py:113-127; patch-embed Conv2d(3,768,16,16)
 * (mmseg/models/utils/embed.py:145-153,199-201) after s4_patchify; the batched QK^T / PV
 * contractions of nn.MultiheadAttention and all their backward forms.
 *
 *   C[z][m,n] = epi(alpha * sum_k A[z][m,k] * B[z][k,n]),  z = z1*nb2 + z2
 *   epi(v): v += bias[n]; v *= gelu'(aux[m,n]); pre[m,n] = v; v = gelu(v) (act);
 *           v += res[m,n]; v += C_old[m,n] (accumulate)
 * Element strides; aux / pre / res share C's strides and batch strides.
 */
typedef struct S4GemmParams {
  const void* a;
  const void* b;
  void* c;
  const float* bias;   /* [N] or NULL */
  const void* aux;     /* [M,N] pre-activations for the GELU-gradient epilogue, or NULL */
  const void* res;     /* [M,N] residual, or NULL */
  void* pre;           /* [M,N] receives the pre-activation, or NULL */
  int M, N, K;
  int nb1, nb2;        /* batch counts (>= 1) */
  long long a_sm, a_sk, a_b1, a_b2;
  long long b_sk, b_sn, b_b1, b_b2;
  long long c_sm, c_b1, c_b2;   /* C is unit-stride in n */
  float alpha;
  int act;             /* S4_ACT_* */
  int accumulate;      /* C += ... */
  int dtype;           /* dtype of A, B, aux, res, pre */
  int c_dtype;         /* dtype of C */
  int backend;         /* S4_BACKEND_* */
  int split_k;         /* tcgen05 path: >1 splits K over CTAs and accumulates atomically
                          into a float32 C (requires accumulate=1, c_dtype=f32, no epilogue);
                          <0 lets the library pick the split count (same requirements) */
  float* colsum;       /* optional [N]: colsum[n] += sum_m C[m,n] from the epilogue (tcgen05 path,
                          plain store, nb1 = nb2 = 1): the bias gradient of the layer whose dY this
                          GEMM produces.  NULL = off.  Check s4_gemm_uses_tc() first. */
} S4GemmParams;

int s4_gemm(const S4GemmParams* p, cudaStream_t stream);
/* 1 if s4_gemm would run this problem on the tcgen05 path */
int s4_gemm_uses_tc(const S4GemmParams* p);
/* Tile policy of the tcgen05 path: 0 = single-CTA tiles only, 1 = cost model (default; also the
 * S4_TC_PAIR environment variable), 2 = CTA-pair (cta_group::2, 256-row) tiles whenever the
 * problem allows them.  Returns the previous mode.  A tuning / test knob: results are the same
 * GEMM in every mode. */
int s4_set_tc_pair_mode(int mode);
/* Tile scheduling of the persistent tcgen05 GEMM / conv grid: 0 = static round-robin (default; also
 * S4_TC_SCHED), 1 = dynamic (CTA groups take tiles from a global counter).  Dynamic costs 1-2 us per
 * launch on an otherwise idle GPU and saves up to half of a launch's time when another kernel (NCCL's
 * all-reduce during the data-parallel backward) holds SMs: CTAs that become resident late find no
 * work left instead of starting a full static share late.  Returns the previous mode.  Same results
 * in both modes (split-K weight-gradient tiles are reduce-added in a different order). */
int s4_set_tc_sched(int mode);

/* ---- LayerNorm (+ row gather) ----------------------------------------------------------------
 * Replaces: ln1/ln2 (vit.py:119-120, eps 1e-6); head LayerNorm with the feature tap
 * (vit.py:556-562) and PatchMix un-shuffle (decode_heads/decode_head.py:186-212,
 * setr_up_head.py:96-103) folded in through `row_map` (dst row -> src row), NULL = identity. */
int s4_layernorm_fwd(const void* x, const int* row_map, const float* gamma, const float* beta,
                     void* y, float* mean, float* rstd, int rows, int D, float eps, int dtype,
                     cudaStream_t stream);
/* dx is written at the mapped source rows (pre-zero it when row_map skips rows); dres (optional,
 * indexed like dx) is added to it: the gradient arriving over the residual connection;
 * dgamma/dbeta are accumulated (+=).  dres_sum (optional fp32 [D], needs dres and no row_map) is
 * accumulated with the column sums of dres: the bias gradient of the linear layer whose output
 * joined the residual stream there (fc2 / out_proj bias of vit.py:113-127), for free. */
int s4_layernorm_bwd(const void* dy, const void* x, const int* row_map, const float* gamma,
                     const float* mean, const float* rstd, const void* dres, void* dx,
                     float* dgamma, float* dbeta, float* dres_sum, int rows, int D, int dtype,
                     cudaStream_t stream);

/* ---- attention with the patch-adaptive (PASA) bias ---------------------------------------------
 * Replaces: nn.MultiheadAttention core as driven by vit.py:119 with the additive float mask
 * built at vit.py:519-535.  The mask is never materialised: bias[b,h,q,k] =
 * w * gate[b,q] * u0[b,k] (rank 1, same for all heads and layers).
 * qkv: [B, L, 3*H*hd] packed as torch in_proj emits it; out: [B, L, H*hd].
 * u0/gate: [B, L] float32 or NULL; lse: [B,H,L] float32 (natural log).  bf16 with hd = 64 runs
 * the fused tcgen05 flash kernels (forward needs no workspace; backward: rowsum(dO*O) and an fp32
 * dQ accumulator); other cases run the composed GEMM + softmax path ([B,H,L,L] scratch). */
size_t s4_attention_workspace(int B, int H, int L, int hd, int dtype, int backend);
int s4_attention_fwd(const void* qkv, const float* u0, const float* gate, float bias_weight,
                     void* out, float* lse, void* workspace, size_t ws_bytes, int B, int H, int L,
                     int hd, int dtype, int backend, cudaStream_t stream);
int s4_attention_bwd(const void* dout, const void* qkv, const void* out, const float* lse,
                     const float* u0, const float* gate, float bias_weight, void* dqkv,
                     void* workspace, size_t ws_bytes, int B, int H, int L, int hd, int dtype,
                     int backend, cudaStream_t stream);

/* Debug / profiling hook (no reference counterpart): record a device-side event timeline of CTA
 * `block` of the following fused attention launches into dev_buf (4 roles x N x {id, clock64}
 * uint64, zero-filled by the caller; N is the return value).  NULL switches tracing off; the
 * production kernel instantiations carry no trace code.  See tools/attn_trace.py. */
int s4_attention_set_trace(void* dev_buf, int block);

/* ---- backbone glue -------------------------------------------------------------------------- */
/* img [B,Cin,H,W] f32 NCHW -> [B*gh*gw, Cin*P*P] (k = (c*P+ky)*P+kx), zero corner padding
 * (embed.py:58-80, 183-204). */
int s4_patchify(const float* img, void* out, int B, int Cin, int H, int W, int P, int dtype,
                cudaStream_t stream);
/* x[b,0]=cls+pos[0]; x[b,1+p]=tok[b,p]+pos[1+p]  (vit.py:486-487, 513) */
int s4_assemble_tokens(const void* tok, const float* cls, const float* pos, void* x, int B, int L,
                       int D, int dtype, cudaStream_t stream);
int s4_assemble_tokens_bwd(const void* dx, void* dtok, float* dcls, float* dpos, int B, int L,
                           int D, int dtype, cudaStream_t stream);
/* sum[c] += sum_r x[r,c]; sumsq[c] += sum_r x[r,c]^2 (sumsq may be NULL) */
int s4_colsum(const void* x, float* sum, float* sumsq, long long rows, int cols, int dtype,
              cudaStream_t stream);
int s4_cast(const void* x, void* y, long long n, int src_dtype, int dst_dtype,
            cudaStream_t stream);
int s4_transpose(const void* x, void* y, int batch, int rows, int cols, int dtype,
                 cudaStream_t stream);

/* ---- SETR-PUP head ---------------------------------------------------------------------------
 * Replaces: ConvModule(conv3x3 no-bias -> SyncBN -> ReLU) + Upsample(bilinear,
 * align_corners=False) stages and the 1x1 conv_seg (decode_heads/setr_up_head.py:49-77,
 * 92-111; decode_head.py:107-111, 311-316; ops/wrappers.py:30-51).  Activations are NHWC. */
/* y[B,H,W,Cout] = conv3x3(x[B,H,W,Cin], w), pad 1, stride 1.  w_packed: [Cout, 9*Cin] with
 * k = (ky*3+kx)*Cin + ci  (s4_pack_conv3x3_weight builds it and the dgrad form). */
int s4_conv3x3_fwd(const void* x, const void* w_packed, void* y, int B, int H, int W, int Cin,
                   int Cout, int dtype, int backend, cudaStream_t stream);
/* Same convolution, plus the BatchNorm batch statistics of its output in the same pass:
 * sum[c] += sum_pix y[pix,c], sumsq[c] += sum_pix y[pix,c]^2 (float32, caller-zeroed).  On the
 * tcgen05 path they come out of the GEMM epilogue (fp32 accumulators, before the bf16 rounding
 * of y); otherwise the library runs s4_colsum on y itself. */
int s4_conv3x3_fwd_stats(const void* x, const void* w_packed, void* y, float* sum, float* sumsq,
                         int B, int H, int W, int Cin, int Cout, int dtype, int backend,
                         cudaStream_t stream);
/* dx = conv3x3(dy, w_dgrad) with w_dgrad: [Cin, 9*Cout], taps flipped */
int s4_conv3x3_dgrad(const void* dy, const void* w_dgrad, void* dx, int B, int H, int W, int Cin,
                     int Cout, int dtype, int backend, cudaStream_t stream);
/* dw[Cout,Cin,3,3] (float32, torch layout) += sum over pixels */
int s4_conv3x3_wgrad(const void* x, const void* dy, float* dw, int B, int H, int W, int Cin,
                     int Cout, int dtype, int backend, cudaStream_t stream);
/* w [Cout,Cin,3,3] f32 -> fwd [Cout, 9*Cin] and dgrad [Cin, 9*Cout] in `dtype` */
int s4_pack_conv3x3_weight(const float* w, void* w_fwd, void* w_dgrad, int Cin, int Cout,
                           int dtype, cudaStream_t stream);
/* BatchNorm statistics from (sum, sumsq, count): mean, invstd (biased var, eps), the fused
 * affine scale = gamma*invstd, shift = beta - mean*scale, and the running-stat update with
 * momentum (unbiased var), as torch.nn.BatchNorm2d / SyncBatchNorm.  running_* may be NULL. */
int s4_bn_finalize(const float* sum, const float* sumsq, double count, float eps, float momentum,
                   const float* gamma, const float* beta, float* mean, float* invstd,
                   float* scale, float* shift, float* running_mean, float* running_var,
                   long long* num_batches_tracked /* +1 when non-NULL */, int C, cudaStream_t stream);
/* eval mode: scale/shift from running statistics */
int s4_bn_eval_affine(const float* running_mean, const float* running_var, const float* gamma,
                      const float* beta, float eps, float* scale, float* shift, int C,
                      cudaStream_t stream);
/* out[B,H*s,W*s,C] = bilinear_s(relu(x*scale[c] + shift[c])) */
int s4_bn_relu_upsample_fwd(const void* x, const float* scale, const float* shift, void* out,
                            int B, int H, int W, int C, int s, int dtype, cudaStream_t stream);
/* dact[B,H,W,C] = relu_mask * upsample^T(dout); also dsum[c] += sum dact, ddot[c] += sum dact*xhat */
int s4_bn_relu_upsample_bwd(const void* dout, const void* x, const float* scale,
                            const float* shift, const float* mean, const float* invstd,
                            void* dact, float* dsum, float* ddot, int B, int H, int W, int C,
                            int s, int dtype, cudaStream_t stream);
/* dy = gamma*invstd*(dact - dsum/n - xhat*ddot/n)  (training-mode BN backward) */
int s4_bn_bwd_apply(const void* dact, const void* x, const float* gamma, const float* mean,
                    const float* invstd, const float* dsum, const float* ddot, double count,
                    void* dy, long long rows, int C, int dtype, cudaStream_t stream);
/* z[B,H,W,NC] (f32) = conv1x1(relu(x*scale+shift)) + bias ; w [NC,C] f32 */
int s4_bn_relu_conv1x1_fwd(const void* x, const float* scale, const float* shift, const float* w,
                           const float* bias, float* z, long long rows, int C, int NC, int dtype,
                           cudaStream_t stream);
/* backward of the above: dact[rows,C] (relu-masked), dw[NC,C] +=, dbias[NC] +=, dsum/ddot += */
int s4_bn_relu_conv1x1_bwd(const float* dz, const void* x, const float* scale, const float* shift,
                           const float* mean, const float* invstd, const float* w, void* dact,
                           float* dw, float* dbias, float* dsum, float* ddot, long long rows,
                           int C, int NC, int dtype, cudaStream_t stream);
/* logits[B,NC,H*s,W*s] (NCHW f32) = bilinear_s(z[B,H,W,NC]) and its transpose */
int s4_upsample_logits_fwd(const float* z, float* logits, int B, int H, int W, int NC, int s,
                           cudaStream_t stream);
int s4_upsample_logits_bwd(const float* dlogits, float* dz, int B, int H, int W, int NC, int s,
                           cudaStream_t stream);

/* bf16 fast path of the last stage's backward (C in {64,128,256,512}, NC <= 24; s4_cls_supported):
 *   dz16 [B*H*W, 32] bf16 (classes >= NC zero) = bilinear_s^T(dlogits)
 *   reduce: dw[NC,C] +=, dbias[NC] +=, dsum[C] += sum da, ddot[C] += sum da*xhat,
 *           da = relu'(x*scale+shift) * (dz w)         (x is read once, nothing is written back)
 *   apply : dy = gamma*invstd*(da - dsum/count - xhat*ddot/count), da recomputed on the fly
 * (the SyncBN all-reduce of dsum/ddot happens between the two calls) */
int s4_cls_supported(int C, int NC, int dtype);
int s4_cls_upsample_bwd_padded(const float* dlogits, void* dz16, int B, int H, int W, int NC, int s,
                               cudaStream_t stream);
int s4_cls_bwd_reduce(const void* dz16, const void* x, const float* scale, const float* shift,
                      const float* mean, const float* invstd, const float* w, float* dw, float* dbias,
                      float* dsum, float* ddot, long long rows, int C, int NC, cudaStream_t stream);
int s4_cls_bwd_apply(const void* dz16, const void* x, const float* scale, const float* shift,
                     const float* mean, const float* invstd, const float* gamma, const float* w,
                     const float* dsum, const float* ddot, double count, void* dy, long long rows,
                     int C, int NC, cudaStream_t stream);

/* ---- pseudo labels and losses ------------------------------------------------------------------
 * s4_pseudo_label replaces encoder_decoder.py:888-901 (+ :541-542) and the patch unconfidence
 * of :547-555: hard = argmax or 255, conf = (max softmax > thr), u = mean_{patch}(1-conf).
 * s4_ce_ncr replaces CrossEntropyLoss (losses/cross_entropy_loss.py:45-61,
 * losses/utils.py:65-69, avg over ALL pixels) and the NCR loop of
 * encoder_decoder.py:936-954, forward and gradient in one pass. */
int s4_pseudo_label(const float* logits, long long* hard, long long* conf, float* u, int B, int C,
                    int H, int W, int patch, float threshold, cudaStream_t stream);
size_t s4_ce_ncr_workspace(int B, int H, int W);
/* loss_out[0..2] = (ce_weight/P * sum nll, ncr_weight/P * sum dist, #valid); dlogits (optional)
 * = g0*dloss0/dz + g1*dloss1/dz with (g0,g1) read from the device pointer grad_scale (NULL = 1,1) */
int s4_ce_ncr(const float* logits_s, const float* logits_t, const long long* label, float* dlogits,
              float* loss_out, const float* grad_scale, int B, int C, int H, int W,
              float ce_weight, float ncr_weight, int ignore_index, void* workspace,
              size_t ws_bytes, cudaStream_t stream);
/* Gradient fix-up for a dlogits produced together with the losses (unit upstream gradients):
 * applies the real upstream gradients grad_scale = (g_ce, g_ncr), read on the device -- nothing is
 * touched when they are (1, 1), a rescale when equal, a recompute from the logits when they
 * differ.  No host synchronisation. */
int s4_ce_ncr_grad_fixup(const float* logits_s, const float* logits_t, const long long* label,
                         float* dlogits, const float* grad_scale, int B, int C, int H, int W,
                         float ce_weight, float ncr_weight, int ignore_index, cudaStream_t stream);
int s4_scale_by_scalar(float* y, const float* scale_dev, size_t n, cudaStream_t stream);

/* ---- augmentation (host RNG, device gathers) --------------------------------------------------
 * generate_unsup_data.py:400-453 (CutMix with neighbour (i+1)%B) and :737-819 (PatchShuffle) */
int s4_cutmix(const float* img, const long long* label, const int* boxes_dev, float* out_img,
              long long* out_label, int B, int C, int H, int W, cudaStream_t stream);
int s4_patchshuffle(const float* img, const long long* perm_dev, float* out, int B, int C, int H,
                    int W, int block, cudaStream_t stream);

/* ---- multi-tensor EMA / SGD ---------------------------------------------------------------------
 * encoder_decoder.py:1044-1066 (t = m*t + (1-m)*s over all parameters and BN running stats) and
 * the SGD-momentum step mmcv's OptimizerHook drives.  Tables live in device memory. */
int s4_chunk_elems(void);
/* bf16_shadow (table of device pointers, entries or the table itself may be NULL): a bfloat16
 * copy of dst refreshed in the same pass (the tensor-core kernels read weights from it). */
int s4_ema_multi_tensor(void* const* dst_ptrs, void* const* src_ptrs, void* const* bf16_shadow,
                        const long long* sizes, const int* chunk_tensor, const long long* chunk_off,
                        int n_chunks, float momentum, float one_minus_momentum, cudaStream_t stream);
int s4_sgd_multi_tensor(void* const* params, void* const* grads, void* const* bufs,
                        void* const* bf16_shadow, const long long* sizes, const float* lrs,
                        const int* chunk_tensor, const long long* chunk_off, int n_chunks,
                        float momentum, float weight_decay, int first_step, cudaStream_t stream);
/* SGD-momentum step fused with the EMA-teacher update of the same parameters (SURVEY.md section
 * 8(f) rank 1: mmcv SGD + encoder_decoder.py:416-423, 1044-1066 in one sweep).  ema_params[i] may
 * be NULL (parameter without a teacher copy); ema_momentum[i] is that tensor's EMA momentum. */
int s4_sgd_ema_multi_tensor(void* const* params, void* const* grads, void* const* bufs,
                            void* const* bf16_shadow, void* const* ema_params,
                            void* const* ema_bf16_shadow, const float* ema_momentum,
                            const long long* sizes, const float* lrs, const int* chunk_tensor,
                            const long long* chunk_off, int n_chunks, float momentum,
                            float weight_decay, int first_step, cudaStream_t stream);

/* ---- inference / validation -----------------------------------
 * encoder_decoder.py:1068-1232 (slide / whole inference, rescale, softmax, flip, argmax) and
 * mmseg/core/evaluation/metrics.py:26-131 (intersect_and_union), on the device. */
/* mmseg/ops/wrappers.py:8-27 resize(mode='bilinear', align_corners=False), NCHW fp32, any scale */
int s4_resize_bilinear_nchw(const float* in, float* out, int planes, int IH, int IW, int OH, int OW,
                            cudaStream_t stream);
/* F.softmax(dim=1) (+ flip: 0 none, 1 horizontal, 2 vertical; :1196-1204) and argmax(dim=1) (:1221);
 * prob [B,C,H,W] and pred [B,H,W] int64 may each be NULL */
int s4_softmax_argmax_nchw(const float* logits, float* prob, long long* pred, int B, int C, int H,
                           int W, int flip, cudaStream_t stream);
/* slide_inference :1089-1093: preds[:, :, y1:y1+ch, x1:x1+cw] += crop; count[...] += 1 */
int s4_accumulate_crop(const float* crop, float* preds, float* count, int B, int C, int H, int W,
                       int y1, int x1, int ch, int cw, cudaStream_t stream);
/* metrics.py:26-91: hist3[0..C) intersect, [C..2C) prediction, [2C..3C) label counts (int64,
 * ACCUMULATED) over pixels with label != ignore_index */
int s4_intersect_union(const long long* pred, const long long* label, long long n, int num_classes,
                       long long ignore_index, long long* hist3, cudaStream_t stream);

/* ---- SyncBatchNorm statistics over NVLink peer memory -------------------
 * Replaces the cross-rank reduction inside torch.nn.SyncBatchNorm (norm_cfg type='SyncBN' of
 * configs/_base_/models/setr_pup.py; mmcv ConvModule at setr_up_head.py:57-64) for the [2, C] fp32
 * statistics of one layer: a one-shot all-reduce (sum) over symmetric buffers, one small kernel per
 * call, replayable inside a CUDA graph.  Every rank allocates s4_peer_allreduce_buffer_bytes() of
 * peer-mapped memory (zeroed once, before the first call on any rank); peer_bufs_dev is a DEVICE array
 * of `world` pointers to those buffers in rank order; seq_state is one zero-initialised uint32 of local
 * device memory per buffer set.  data [n] (n <= s4_peer_allreduce_max_elems()) is reduced in place;
 * all ranks add in rank order, so the sums are bit-identical everywhere.  All ranks must make the
 * same sequence of calls. */
long long s4_peer_allreduce_buffer_bytes(void);
int s4_peer_allreduce_max_elems(void);
int s4_peer_allreduce_f32(float* data, int n, const void* peer_bufs_dev, int rank, int world,
                          unsigned* seq_state, cudaStream_t stream);
/* SyncBN forward in one kernel: the same exchange of the LOCAL stats [2, C] = (sum, sumsq) followed by
 * s4_bn_finalize's arithmetic on the rank-ordered sums (count = total elements per channel over all
 * ranks).  stats is not modified. */
int s4_bn_finalize_peer(const float* stats, double count, float eps, float momentum,
                        const float* gamma, const float* beta, float* mean, float* invstd,
                        float* scale, float* shift, float* running_mean, float* running_var,
                        long long* num_batches_tracked, int C, const void* peer_bufs_dev, int rank,
                        int world, unsigned* seq_state, cudaStream_t stream);

/* ---- GPU-side input pipeline -------------------------------------
 * configs/setr/..._MT_w_ours.py:42-126 strong / weak / sup branch pipelines after RandomCrop / RandomFlip:
 * PhotoMetricDistortion (transforms.py:1165-1272), Normalize (:572-604), Pad (:484-565),
 * DefaultFormatBundle (formatting.py:202-225) for every branch in one launch.
 * crops[i] -> device uint8 [h_i, w_i, 3] BGR, labels[i] -> device uint8 [h_i, w_i] or NULL, crop_hw = {h_0, w_0, ...};
 * branch j reads crop crop_of[j] with the distortion draws pmd_params[j] (9 x 4 bytes:
 * do_brightness, beta, mode, do_contrast, alpha_c, do_saturation, alpha_s, do_hue, hue_delta);
 * mean: 3 floats, stdinv: 3 DOUBLES (= 1 / double(float(std)), what cv2.multiply is handed);
 * out_img [n_branches, 3, pad_h, pad_w] f32, out_label [n_branches, 1, pad_h, pad_w] i64 (may be NULL),
 * out_u8 (may be NULL): the distorted image before normalisation, [n_branches, pad_h, pad_w, 3]. */
int s4_branch_pipeline(const void* const* crops, const void* const* labels, const int* crop_hw,
                       const int* crop_of, const void* pmd_params, const float* mean, const double* stdinv,
                       int to_rgb, int seg_pad_val, float* out_img, long long* out_label, void* out_u8,
                       int n_branches, int pad_h, int pad_w, cudaStream_t stream);
int s4_pmd_params_size(void);

#ifdef __cplusplus
}
#endif
#endif /* S4FORMER_H_ */
