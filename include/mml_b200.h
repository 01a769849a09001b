/* mml_b200.h -- C ABI of libmml_b200.so: the B200 (sm_100a) kernels behind MML_Suite's late-fusion training step.
 *
 * The reference (TArsenii/task-specific-pretraining-multimodal, directory MML_Suite/) is pure PyTorch: it has no
 * native code and therefore no FFI of its own.  The boundary below is what a binding for the hot path replaces, op by
 * op; every entry point cites the reference call it stands in for (paths relative to MML_Suite/).  The Python host
 * side (task-specific-pretraining-multimodal_b200/*.py) binds these with ctypes; see INTEGRATION.md.
 *
 * Conventions
 *   - plain pointers and sizes only; no torch / C++ types.  All tensor memory is caller-owned DEVICE memory.
 *   - activations: NHWC bf16 (uint16_t storage).  conv weights: K,R,S,C ("KRSC", == torch channels_last of OIHW).
 *   - every kernel is enqueued on `stream` (a cudaStream_t passed as void*) and never synchronises, so a whole
 *     training step can be captured into a CUDA graph.
 *   - return value: 0 on success, negative mml_status on failure; text via mml_last_error().  Never throws/aborts.
 *   - one ctx per (process, device); not re-entrant.
 */
#ifndef MML_B200_H
#define MML_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct mml_ctx mml_ctx;

/* BatchNorm statistics accumulators are fp64 arrays [S][C][2] (sum, sum of squares), S = mml_bn_stat_slots(C) = clamp(1024 / C,
 * 2, 16): producers (conv / stem epilogues, the BN backward reduce) add their partial sums with fp64 atomics into slot
 * (CTA index % S) to spread the contention, every consumer sums the S slots (16 KB per BatchNorm whatever C is).  The caller
 * zeroes them once per step. */
int mml_bn_stat_slots(int C);

enum mml_status {
  MML_OK = 0,
  MML_ERR_INVALID = -1,     /* bad argument / unsupported geometry */
  MML_ERR_CUDA = -2,        /* a CUDA runtime / driver call failed */
  MML_ERR_UNSUPPORTED = -3, /* not an sm_100 device, missing driver entry point, ... */
};

/* conv geometry: x [N,H,W,C] -> y [N,P,Q,K], filter K x R x S x C, P = (H + 2*pad - R)/stride + 1 (nn.Conv2d, bias=False) */
typedef struct mml_conv_geom {
  int32_t N, H, W, C; /* input */
  int32_t K, R, S;    /* filter */
  int32_t stride, pad;
} mml_conv_geom;

/* ---- context ------------------------------------------------------------------------------------------------- */
int mml_version(void);
int mml_ctx_create(int device, mml_ctx** out);
void mml_ctx_destroy(mml_ctx* ctx);
const char* mml_last_error(const mml_ctx* ctx); /* ctx may be NULL: error of a failed mml_ctx_create */
int mml_ctx_sm_count(const mml_ctx* ctx);
/* number of kernels this library has launched since the ctx was created (bench.py's gpu_launches) */
int64_t mml_ctx_launch_count(const mml_ctx* ctx);
/* SMs the persistent convolution kernels may occupy from now on (0 or > SM count = all).  Read at launch time, so it can differ
 * per launch; the two-encoder step keeps a few SMs free for the image encoder's stream of small kernels. */
int mml_ctx_set_sm_budget(mml_ctx* ctx, int sms);
/* Programmatic dependent launch for the launches that follow (read at launch time): the next kernel of a stream is scheduled
 * while the previous one still runs and waits (griddepcontrol.wait) after its own set-up.  It shortens a chain of SMALL dependent
 * kernels (the ResNet34 image encoder: -7.5 % alone) but the pre-launched CTAs of LARGE grids hold SMs another stream could use
 * (two-encoder step: +3.6 % when applied to the audio encoder too), so the caller chooses per stream.  Default: on. */
int mml_ctx_set_pdl(mml_ctx* ctx, int enable);

/* A-B switches for experiments (not part of the reference surface): key 1 = use the halo conv kernel (default 1); key 2 = largest
 * thread-block cluster of the split-K convolution variant (1 = off (default), 2, 4, 8); key 3 = BatchNorm grids: 0 = fixed caps,
 * 1 = one resident wave, 2 = one resident wave of the SM budget (default); key 4 = fewest 128-pixel tiles per weight-gradient split (default 48);
 * key 5 = weight gradients of layers with at most this many pixel tiles use 128-wide output tiles (default 32, 0 = off); key 6 = the same
 * for fprop / dgrad with 64-wide tiles (default 0 = off) */
int mml_debug_set(int key, int value);

/* ---- a1: missing-modality mask -- data/base_dataset.py:70-72  sample[mod] = original * mask -------------------- */
/* y[b, :] = x[b, :] * mask[b]   (true IEEE multiply, bit-exact with torch CPU); reverse: x * -1 * (mask - 1) */
int mml_mask_apply_f32(mml_ctx*, const float* x, const float* mask, float* y, float* y_reverse, int64_t batch,
                       int64_t per_sample, void* stream);

/* ---- f4: the input path in front of the encoders, on the device (staging.cu) --------------------------------------------------
 * Mask draw: data/base_dataset.py:46-59 _initialise_missing_masks -> create_missing_mask(n_modalities, n, [P(present)]): one
 * independent Bernoulli(P(present)) per (sample, modality), drawn once per pattern.  masks[m*ld + (i - first_sample)] for
 * i in [first_sample, first_sample + count) = (u < p_present[m]) ? 1 : 0, u = (bits >> 8) * 2^-24 with bits = word (i % 4) of
 * Philox4x32-10(counter = {lo32(i/4), hi32(i/4), m, stream_id}, key = {lo32(seed), hi32(seed)}): a pure function of
 * (seed, stream_id, m, i), so any shard of the sample range (one rank's slice) reproduces the single-GPU draw bit for bit.
 * p_present: device fp32 [n_modalities]; stream_id: the pattern's index. */
int mml_missing_mask_draw(mml_ctx*, const float* p_present, float* masks, int n_modalities, int64_t first_sample, int64_t count, int64_t ld,
                          uint64_t seed, uint32_t stream_id, void* stream);
/* data/avmnist.py:193-224 __getitem__ looks the mask of sample idx up per item: out[m*batch + b] = masks[m*num_samples + sample_idx[b]];
 * an index outside [0, num_samples) yields 0 and sets *bad_index_flag (device int, may be NULL) to 1 */
int mml_missing_mask_gather(mml_ctx*, const float* masks, const int64_t* sample_idx, float* out, int n_modalities, int64_t num_samples,
                            int64_t batch, int* bad_index_flag, void* stream);
/* data/avmnist.py:188-191 _load_image for uint8 pixels: colormap -> uint8 RGBA -> PIL "L" -> float32 / 255 is a 256-entry table of the
 * pixel value (host-built, mml_b200.data.luma_lut): dst[i] = lut256[src[i]].  src and dst 16-byte aligned, any n. */
int mml_stage_u8_lut_f32(mml_ctx*, const uint8_t* src, const float* lut256, float* dst, int64_t n, void* stream);

/* ---- a2/a3: ResNetEncoder stem -- models/msa/networks/resnet.py:137 conv1 (7x7, stride 2, pad 3, C_in = 1) ----- */
/* x fp32 [B,H,W] (optionally multiplied by mask[b], same multiply as above), w fp32 [64][7][7] ->
 * y bf16 [B,P,Q,64]; stats (optional) fp64 [16][64][2] += (sum, sum of squares) of the stored y */
int mml_stem_fprop(mml_ctx*, const float* x, const float* mask, const float* w, uint16_t* y, double* stats, int B, int H,
                   int W, void* stream);
/* dw fp32 [64][49] = sum_{b,p,q} dy[b,p,q,k] * (x*mask)[b, 2p+r-3, 2q+s-3]  (overwrites dw) */
int mml_stem_wgrad(mml_ctx*, const float* x, const float* mask, const uint16_t* dy, float* dw, float* workspace,
                   int64_t workspace_bytes, int B, int H, int W, void* stream);
/* The same weight gradient with the stem BatchNorm's backward pass 2 folded in (resnet.py:137-138 conv1 -> bn1 and their autograd): g =
 * gradient w.r.t. the BatchNorm OUTPUT after the ReLU mask (what mml_stem_bn_pool_bwd(apply = 0) leaves in dx), bstat = its (sum g,
 * sum g*xhat).  Because the stem output is linear in the input patches, dx never has to exist:
 *   dw[k][t] = gamma_k invstd_k ( sum_p g x_t - mean(g)_k sum_p x_t - mean(g xhat)_k invstd_k ( sum_t' w[k][t'] sum_p x_t' x_t - mu_k sum_p x_t ) )
 * with the patch Gram matrix and the tap sums produced by the same tensor-core pass that forms sum_p g x_t.  w: the fp32 stem weights the
 * forward used; also stores dgamma = sum g*xhat and dbeta = sum g.  workspace: mml_stem_wgrad_workspace bytes, 8-byte aligned. */
int mml_stem_wgrad_bn(mml_ctx*, const float* x, const float* mask, const uint16_t* g, const float* w, const double* bstat, const float* mean,
                      const float* invstd, const float* gamma, float* dgamma, float* dbeta, float* dw, float* workspace,
                      int64_t workspace_bytes, int B, int H, int W, void* stream);
int64_t mml_stem_wgrad_workspace(const mml_ctx*, int B, int H, int W);

/* ---- a2-a4: 3x3 / 1x1 convolutions -- resnet.py:25,30,176 (nn.Conv2d fwd) and their autograd -------------------- */
/* tcgen05 implicit GEMM.  fprop: y = conv(x, w); stats (optional) fp64 [S][K][2] += per-channel (sum, sum of squares) of y */
int mml_conv_fprop(mml_ctx*, const mml_conv_geom* g, const uint16_t* x, const uint16_t* w_krsc, uint16_t* y, double* stats,
                   void* stream);
/* dgrad: dx [N,H,W,C] = conv_transpose(dy [N,P,Q,K], w); reads the SAME K,R,S,C weights as fprop (MN-major B operand) */
int mml_conv_dgrad(mml_ctx*, const mml_conv_geom* g, const uint16_t* dy, const uint16_t* w_krsc, uint16_t* dx, void* stream);
/* wgrad: dw_krsc fp32 [K][R][S][C] = sum_{n,p,q} dy * x   (OVERWRITES dw; deterministic: split partial sums go to `workspace`
 * with plain stores and are added in a fixed order -- the reference runs with cudnn.deterministic = True,
 * config/experiment_config.py:62-63).  workspace: device scratch of at least mml_conv_wgrad_workspace() bytes (may be NULL when
 * that is 0), private to the stream the call is enqueued on until the next call on that stream. */
int mml_conv_wgrad(mml_ctx*, const mml_conv_geom* g, const uint16_t* x, const uint16_t* dy, float* dw_krsc, float* workspace,
                   int64_t workspace_bytes, void* stream);
int64_t mml_conv_wgrad_workspace(const mml_ctx*, const mml_conv_geom* g); /* bytes; < 0: unsupported geometry */

/* ---- a5: BatchNorm2d (train / eval) + ReLU + residual -- resnet.py:26,31,138,177 and BasicBlock.forward :37-54 --- */
/* training mode, fused: y = relu?(bn(x) [+ res | + bn_r(res)]) with scale/shift derived in-kernel from the fp64 sums the conv
 * epilogue accumulated (stats [S][C][2]: 16 KB per CTA, no finalize launch); block 0 also saves mean / invstd for
 * backward and updates the running statistics (running = (1-m)*running + m*batch, unbiased var).  res NULL: none; rstats NULL:
 * identity residual; else the residual goes through its own training-mode BN (downsample path).  count == rows. */
int mml_bn_train_fwd(mml_ctx*, const uint16_t* x, const double* stats, const float* gamma, const float* beta, float* running_mean,
                     float* running_var, float* save_mean, float* save_invstd, const uint16_t* res, const double* rstats,
                     const float* rgamma, const float* rbeta, float* r_running_mean, float* r_running_var, float* r_save_mean,
                     float* r_save_invstd, uint16_t* y, int64_t rows, int C, int relu, float momentum, float eps, void* stream);
/* eval mode: scale/shift from running statistics */
int mml_bn_eval_coeffs(mml_ctx*, int C, const float* gamma, const float* beta, const float* running_mean,
                       const float* running_var, float eps, float* scale, float* shift, void* stream);
/* y = act(x*scale + shift [+ res*rscale + rshift]); res may be NULL; rscale NULL => identity residual */
int mml_bn_act_fwd(mml_ctx*, const uint16_t* x, const float* scale, const float* shift, const uint16_t* res,
                   const float* rscale, const float* rshift, uint16_t* y, int64_t rows, int C, int relu, void* stream);
/* backward of y = relu?(bn(x) [+ r]):  g = (dy1 [+ dy2]) * (y > 0 if relu);
 * pass 1: bstat fp64 [S][C][2] += (sum g, sum g*xhat) (caller zeroes it per step); g_out (optional, may alias dy1) = g as bf16 -- it
 * is the gradient of an identity skip path and the input of pass 2 (which then reads 2 tensors instead of 4) */
int mml_bn_bwd_reduce(mml_ctx*, const uint16_t* dy1, const uint16_t* dy2, const uint16_t* y, const uint16_t* x,
                      const float* mean, const float* invstd, double* bstat, uint16_t* g_out, int64_t rows, int C, int relu, void* stream);
/* pass 2: dx = gamma*invstd*(g - mean(g) - xhat*mean(g*xhat)) as bf16 (dx may alias g); coefficients straight from bstat;
 * dgamma = sum g*xhat, dbeta = sum g (written, may be NULL) */
int mml_bn_bwd_apply(mml_ctx*, const uint16_t* g, const uint16_t* x, const float* mean, const float* invstd, const float* gamma,
                     const double* bstat, float* dgamma, float* dbeta, uint16_t* dx, int64_t rows, int C, void* stream);

/* ---- pooling -- resnet.py:140 MaxPool2d(3,2,1), :149 AdaptiveAvgPool2d((1,1)) ---------------------------------- */
int mml_maxpool3x3s2_fwd(mml_ctx*, const uint16_t* x, uint16_t* y, uint8_t* argmax, int N, int H, int W, int C, void* stream);
/* dx = scatter of (dy [+ dy2]) to the argmax positions; dy2 may be NULL */
int mml_maxpool3x3s2_bwd(mml_ctx*, const uint16_t* dy, const uint16_t* dy2, const uint8_t* argmax, uint16_t* dx, int N, int H,
                         int W, int C, void* stream);
/* stem tail, fused: y = maxpool3x3s2(relu(bn(x))) without materialising the activation (resnet.py:138-140, :206-208).
 * train: stats != NULL (fp64 sums from mml_stem_fprop; saves mean/invstd, updates running stats); eval: scale/shift != NULL */
int mml_stem_bn_pool_fwd(mml_ctx*, const uint16_t* x, const double* stats, const float* gamma, const float* beta, float* running_mean,
                         float* running_var, float* save_mean, float* save_invstd, const float* scale, const float* shift, uint16_t* y,
                         uint8_t* argmax, int N, int H, int W, int C, float momentum, float eps, void* stream);
/* its backward: dx (grad of the raw stem output) from the pooled gradient(s); ReLU mask recomputed from x; two passes
 * (scatter + statistics, then the in-place apply).  apply == 0 stops after pass 1: dx holds g = scatter(dpool) * [bn(x) > 0] and bstat
 * the sums (sum g, sum g*xhat); dgamma / dbeta are then written by mml_stem_wgrad_bn, which folds pass 2 into the weight gradient */
int mml_stem_bn_pool_bwd(mml_ctx*, const uint16_t* dy, const uint16_t* dy2, const uint8_t* argmax, const uint16_t* x, const float* mean,
                         const float* invstd, const float* gamma, const float* beta, double* bstat, float* dgamma, float* dbeta, uint16_t* dx,
                         int N, int H, int W, int C, int apply, void* stream);
/* ---- ConvBlock encoders (MML_Suite/models/conv.py:16-59, models/avmnist.py:34-185: MNISTAudio / MNISTImage) ----
 * First convolution of a ConvBlock encoder: Conv2d(1, K, 3, stride 1, padding 1) over (x * mask) (fp32 [B][H][W], mask [B] or NULL), weights
 * fp32 [K][9], K in {8,16,32,64}; output NHWC bf16 with the channel dimension PADDED to 64 (channels >= K written as zeros) so that every later
 * layer is a 64-channel tensor-core convolution; stats as mml_conv_fprop with C = 64.  The bias is not an argument: see mml_bn_conv_bias_fold. */
int mml_conv3x3_c1_fprop(mml_ctx*, const float* x, const float* mask, const float* w, uint16_t* y, double* stats, int B, int H, int W, int K,
                         void* stream);
/* its weight gradient dw fp32 [K][9] (OVERWRITTEN; deterministic: per-CTA partials in `workspace`, added in a fixed order) */
int mml_conv3x3_c1_wgrad(mml_ctx*, const float* x, const float* mask, const uint16_t* dy, float* dw, float* workspace, int64_t workspace_bytes,
                         int B, int H, int W, int K, void* stream);
int64_t mml_conv3x3_c1_wgrad_workspace(const mml_ctx*, int B, int H, int K);
/* nn.MaxPool2d(kernel_size = k) (stride k, no padding, floor): x NHWC bf16 [B][H][W][C] -> y NHWC bf16 [B][H/k][W/k][C] and / or y_flat_nchw
 * fp32 [B][C*(H/k)*(W/k)] in nn.Flatten order (either may be NULL); argmax uint8 = r*k+s of the first maximum.  Backward takes the gradient in
 * exactly one of the two layouts and writes dx (zeros where the window maximum was elsewhere and in the rows / columns the floor drops). */
int mml_maxpool_k_fwd(mml_ctx*, const uint16_t* x, uint16_t* y, float* y_flat_nchw, uint8_t* argmax, int B, int H, int W, int C, int k,
                      void* stream);
int mml_maxpool_k_bwd(mml_ctx*, const uint16_t* dy, const float* dy_flat_nchw, const uint8_t* argmax, uint16_t* dx, int B, int H, int W, int C,
                      int k, void* stream);
/* Conv2d bias in front of a BatchNorm2d (conv.py:24-45): running_mean += momentum * bias (train; pass scale = shift = NULL) or
 * shift += bias * scale (eval coefficients of mml_bn_eval_coeffs; pass running_mean = NULL). */
int mml_bn_conv_bias_fold(mml_ctx*, const float* conv_bias, int C, float momentum, float* running_mean, const float* scale, float* shift,
                          void* stream);

int mml_avgpool_fwd(mml_ctx*, const uint16_t* x, float* y, int N, int HW, int C, void* stream);
int mml_avgpool_bwd(mml_ctx*, const float* dy, uint16_t* dx, int N, int HW, int C, void* stream);

/* ---- a6/a7: encoder fc x2 + concat + fusion MLP + dropout + softmax-CE -- avmnist.py:219-267, loss.py:98-148 ----- */
typedef struct mml_head_params {
  /* fp32 device pointers, torch nn.Linear layout [out][in] */
  const float *fcA_w, *fcA_b; /* [EA][FA] */
  const float *fcI_w, *fcI_b; /* [EI][FI] */
  const float *w0, *b0;       /* [H1][EA+EI]  net.0 */
  const float *w3, *b3;       /* [H2][H1]     net.3 */
  const float *w5, *b5;       /* [NC][H2]     net.5 */
  int32_t FA, FI, EA, EI, H1, H2, NC;
} mml_head_params;
typedef struct mml_head_grads {
  float *fcA_w, *fcA_b, *fcI_w, *fcI_b, *w0, *b0, *w3, *b3, *w5, *b5; /* written (not accumulated) */
} mml_head_grads;
/* scratch: fp32 [B][mml_head_scratch_per_sample()] kept between fwd and bwd */
int mml_head_scratch_per_sample(const mml_head_params* p);
/* pooledA [B][FA], pooledI [B][FI] fp32; labels int64 [B] (may be NULL: no loss); dropout_mask uint8 [B][H1] or NULL;
 * dropout_scale = 1/(1-p).  Outputs: logits [B][NC], loss_out[0] = mean CE, pred int32 [B] (argmax of softmax) */
int mml_head_fwd(mml_ctx*, const mml_head_params* p, const float* pooledA, const float* pooledI, const int64_t* labels,
                 const uint8_t* dropout_mask, float dropout_scale, float* scratch, float* logits, float* loss_out,
                 int32_t* pred, int B, void* stream);
/* backward of mean CE.  phases bit 0: data gradients (dpooledA [B][FA], dpooledI [B][FI], per-sample deltas into scratch);
 * bit 1: weight / bias gradients from those deltas (independent of the encoders' backward, so it can run on another
 * stream); 3 = both.  loss_scale multiplies dlogits. */
int mml_head_bwd(mml_ctx*, const mml_head_params* p, const mml_head_grads* g, const float* pooledA, const float* pooledI,
                 const int64_t* labels, const uint8_t* dropout_mask, float dropout_scale, float* scratch,
                 float loss_scale, float* dpooledA, float* dpooledI, int B, int phases, void* stream);
/* mean softmax cross-entropy over [B][NC] logits (CrossEntropyLoss() defaults, loss.py:48): dlogits = (softmax - onehot) * loss_scale / B,
 * row_loss [B] scratch, loss_out = mean, pred = argmax; labels / dlogits / row_loss / loss_out / pred optional. */
int mml_softmax_ce(mml_ctx*, const float* logits, const int64_t* labels, float* dlogits, float* row_loss, float* loss_out, int32_t* pred,
                   float loss_scale, int B, int NC, void* stream);
/* MonomodalEncoder tail (train_monomodal.py:64-92,224-232): emb = pooled W_fc^T + b_fc (the encoder's own fc, resnet.py:218),
 * logits = emb W_cls^T + b_cls, mean cross-entropy, argmax, dlogits = (softmax - onehot) * loss_scale / B.  labels / dlogits /
 * row_loss [B] / loss_out / pred are optional (forward only).  Backward: weight / bias gradients of both Linears (stored) and
 * dpooled [B][F]; demb [B][E] is scratch. */
int mml_mono_head_fwd(mml_ctx*, const float* pooled, const float* fc_w, const float* fc_b, const float* cls_w, const float* cls_b,
                      const int64_t* labels, float* emb, float* logits, float* dlogits, float* row_loss, float* loss_out, int32_t* pred,
                      float loss_scale, int B, int F, int E, int NC, void* stream);
int mml_mono_head_bwd(mml_ctx*, const float* pooled, const float* emb, const float* dlogits, const float* fc_w, const float* cls_w,
                      float* d_fc_w, float* d_fc_b, float* d_cls_w, float* d_cls_b, float* demb, float* dpooled, int B, int F, int E, int NC,
                      void* stream);
/* stand-alone nn.Linear forward (encoder fc outside the fused head, resnet.py:218): y [B][n_out] = x [B][n_in] W^T + b */
int mml_linear_fwd(mml_ctx*, const float* x, const float* w, const float* bias, float* y, int B, int n_in, int n_out, void* stream);
/* Philox-free counter RNG for the throughput path: mask[i] = hash(seed, *step_counter, i) >= p ? 1 : 0 */
int mml_dropout_mask(mml_ctx*, uint8_t* mask, int64_t n, float p, uint64_t seed, const int64_t* step_counter, void* stream);

/* ---- a11 (config 3): MMIMDb gated late fusion -- MML_Suite/models/mmimdb.py:20-245 ------------------------------- */
/* The Linear layers of this model run on mml_conv_fprop / _dgrad / _wgrad as 1x1 convolutions over [B,1,1,C] bf16 rows
 * (nn.Linear's [out][in] weight IS the K,R,S,C layout).  A bias is carried as one more input column: activations have a
 * constant 1 at column `in`, the weight row has the bias there (row pitch rounded up to 64), so fprop adds it and wgrad
 * produces its gradient.  The entry points below are everything between those GEMMs. */
enum { MML_BN1D_INPUT = 0, MML_BN1D_GATED = 1, MML_BN1D_MAXOUT = 2, MML_BN1D_MAX2 = 3 };
/* nn.BatchNorm1d over the batch (mmimdb.py:38,43,46,80) fused with the op that PRODUCES its input:
 *   INPUT : v[b][c] = x[b*ldx + c] * mask[b]        missing-modality mask (base_dataset.py:71); mask may be NULL
 *   GATED : v = gate[b]*h1 + (1-gate[b])*h2         GatedBiModalNetwork.forward, gated_bimodal.py:59 (fp32 [B][C]);
 *           with gate == NULL: v = mix_a*h1 + mix_b*h2  (MultimodalPooling "avg" = .5/.5, "sum" = 1/1; pooling.py:105-111)
 *   MAX2  : v = max(h1, h2)                          MultimodalPooling "max" (pooling.py:101-103)
 *   MAXOUT: v = max(pre[b][c], pre[b][C+c]) * (keep ? keep[b][c]*keep_scale : 1)   MaxOut (maxout.py:37-41) + Dropout;
 *           pre bf16 [B][2C] = both units' GEMM output side by side; keep uint8 [B][C] or NULL
 * train != 0: batch statistics (biased variance), running statistics updated with the unbiased one; else running stats.
 * Outputs: xhat fp32 [B][C] and invstd [C] (saved for backward, optional), y = gamma*xhat+beta as bf16 rows of pitch ldy
 * (the next GEMM's A operand) and / or fp32 [B][C]. */
typedef struct mml_bn1d_desc {
  int32_t mode, B, C, train;
  const float* x; const float* mask; int64_t ldx;
  const float* h1; const float* h2; const float* gate;
  const uint16_t* pre; const uint8_t* keep; float keep_scale; float momentum; float eps; float mix_a; float mix_b; float reserved;
  const float* gamma; const float* beta; float* running_mean; float* running_var;
  float* xhat; float* invstd; uint16_t* y_bf16; int64_t ldy; float* y_f32;
} mml_bn1d_desc;
int mml_bn1d_fwd(mml_ctx*, const mml_bn1d_desc*, void* stream);
/* backward of the same: dy bf16 rows (pitch lddy) -> dgamma, dbeta [C] (stored, not accumulated) and
 *   INPUT : nothing else;  GATED / MAX2: dz fp32 [B][C] (gradient of the mixed value, routed on by mml_gmu_bwd /
 *   mml_pool_bwd);  MAXOUT: dpre bf16 [B][2C] (winner takes the gradient, ties split). */
typedef struct mml_bn1d_bwd_desc {
  int32_t mode, B, C, reserved;
  const uint16_t* dy; int64_t lddy; const float* xhat; const float* gamma; const float* invstd;
  float* dgamma; float* dbeta;
  const uint16_t* pre; const uint8_t* keep; float keep_scale; float reserved2; uint16_t* dpre;
  float* dz;
} mml_bn1d_bwd_desc;
int mml_bn1d_bwd(mml_ctx*, const mml_bn1d_bwd_desc*, void* stream);
/* GMU (gated_bimodal.py:52-59): h1 = tanh(h1pre), h2 = tanh(h2pre) (bf16 [B][H] GEMM outputs -> fp32), and the scalar
 * gate[b] = sigmoid(wz . [h1|h2]).  Backward: dz fp32 [B][H] -> dh1pre, dh2pre bf16 and dwz [2H] (ACCUMULATED). */
int mml_gmu_fwd(mml_ctx*, const uint16_t* h1pre, const uint16_t* h2pre, const float* wz, float* h1, float* h2, float* gate, int B,
                int H, void* stream);
int mml_gmu_bwd(mml_ctx*, const float* dz, const float* h1, const float* h2, const float* gate, const float* wz, float* dwz,
                uint16_t* dh1pre, uint16_t* dh2pre, int B, int H, void* stream);
/* MultimodalPooling branches (pooling.py:92-98): h = dropout(tanh(pre + bias)) for both modalities; pre bf16 [B][H] (GEMM
 * outputs of proj_a / proj_b), keep uint8 [B][H] or NULL, outputs fp32 [B][H] (consumed by mml_bn1d_fwd GATED / MAX2).
 * Backward: dz fp32 [B][H] (from mml_bn1d_bwd) -> dpre bf16 for both branches + bias gradients [H] (stored).
 * kind: 0 = max (winner takes the gradient, ties split), 1 = linear mix with mix_a / mix_b.
 * comb (optional): bf16 [B][2H] = [h_a | h_b], the A operand of the attention / gate GEMM.  gate (optional, kind 1): per-sample
 * mix g / 1-g instead of mix_a / mix_b;  dcomb (optional): bf16 [B][2H] gradient that came back through that GEMM, added in. */
int mml_pool_fwd(mml_ctx*, const uint16_t* pre_a, const uint16_t* pre_b, const float* bias_a, const float* bias_b, const uint8_t* keep_a,
                 const uint8_t* keep_b, float keep_scale, float* h_a, float* h_b, uint16_t* comb, int B, int H, void* stream);
int mml_pool_bwd(mml_ctx*, const float* dz, const float* h_a, const float* h_b, const uint8_t* keep_a, const uint8_t* keep_b,
                 float keep_scale, int kind, float mix_a, float mix_b, const float* gate, const uint16_t* dcomb, uint16_t* dpre_a,
                 uint16_t* dpre_b, float* dbias_a, float* dbias_b, int B, int H, void* stream);
/* "attention" / "gated" pooling head (pooling.py:55-72,113-126): hid bf16 [B][Hd] = GEMM output of layer 0 (bias b0 added here),
 * t = tanh(hid + b0) (fp32, saved), s = W2 t + b2 with NS = 2 (attention: softmax over the two scores == (g, 1-g),
 * g = sigmoid(s0 - s1)) or NS = 1 (gated: g = sigmoid(s0)); gate[b] = g feeds mml_bn1d_fwd(GATED).  Backward: dz fp32 [B][H] ->
 * dw2 [NS][Hd], db2 [NS], db0 [Hd] (ACCUMULATED) and dhid bf16 [B][Hd] (dgrad / wgrad of layer 0 follow on the GEMM path). */
int mml_att_fwd(mml_ctx*, const uint16_t* hid, const float* b0, const float* w2, const float* b2, float* t, float* gate, int B, int Hd,
                int NS, void* stream);
int mml_att_bwd(mml_ctx*, const float* dz, const float* h_a, const float* h_b, const float* gate, const float* t, const float* w2,
                float* dw2, float* db2, float* db0, uint16_t* dhid, int B, int H, int Hd, int NS, void* stream);
/* classifier tail (mmimdb.py:47, loss.py:52, mmimdb.py:238-239): logits = xn W^T + b, loss = mean BCE-with-logits over
 * B x NC, dlogits = (sigmoid - y) * grad_scale / (B NC), pred = sigmoid(logit) > threshold.  labels / loss / dlogits /
 * pred are optional; scratch: mml_bce_head_scratch_floats(B) floats, zero-initialised once by the caller. */
int64_t mml_bce_head_scratch_floats(int B);
int mml_bce_head_fwd(mml_ctx*, const float* xn, const float* w, const float* bias, const float* labels, float* logits, float* loss,
                     float* dlogits, uint8_t* pred, float* scratch, float threshold, float grad_scale, int B, int H, int NC, void* stream);
int mml_bce_head_bwd(mml_ctx*, const float* dlogits, const float* xn, const float* w, float* dw, float* db, uint16_t* dxn, int B, int H,
                     int NC, void* stream);

/* ---- a12 (config 4): MOSI / UttFusion -- MML_Suite/models/msa/utt_fusion.py:106-198 ----------------------------------------- */
/* TextCNN convolutions (textcnn.py:29-49) run on mml_conv_fprop / mml_conv_wgrad: Conv2d(1, 128, (k, 768)).weight is a K,R,S,C tensor
 * with R = k, S = 1, C = 768 and the text input [B][T][768] is NHWC [B][T][1][768] (bf16 copy via mml_cast_f32_bf16). */
#define MML_CLIP_PARTIALS 256
/* one-layer batch_first nn.LSTM from zero state (lstm.py:17,62-64; gate order i,f,g,o): x fp32 [B][T][IN], w_ih [4H][IN], w_hh [4H][H],
 * b_ih, b_hh [4H] -> h_last [B][H] ("last" embedding); saved for BPTT: gates [B][T][4H] (activated), cs, hs [B][T][H].  H = 64, IN <= 32.
 * Backward: dh_last [B][H] -> dw_ih, dw_hh, db_ih, db_hh (ACCUMULATED); the input gets no gradient. */
int mml_lstm_fwd(mml_ctx*, const float* x, const float* w_ih, const float* w_hh, const float* b_ih, const float* b_hh, float* gates,
                 float* cs, float* hs, float* h_last, int B, int T, int IN, int H, void* stream);
int mml_lstm_bwd(mml_ctx*, const float* x, const float* w_hh, const float* gates, const float* cs, const float* hs, const float* dh_last,
                 float* dw_ih, float* dw_hh, float* db_ih, float* db_hh, int B, int T, int IN, int H, void* stream);
/* TextCNN conv_block tail (textcnn.py:51-58) + Dropout (:66): y[b][y_off + c] = keep * scale * max_t relu(conv[b][t][c] + bias[c]) over the
 * P valid positions, conv bf16 [B][P][C]; arg = arg-max position or -1 (ReLU inactive).  Backward: dconv bf16 [B][P][C] (zero except at
 * the arg-max), dbias [C] (ACCUMULATED).  y / dy / keep / arg are [B][ldy] (the concatenation of the three blocks). */
int mml_relumax_fwd(mml_ctx*, const uint16_t* conv, const float* bias, const uint8_t* keep, float keep_scale, float* y, int32_t* arg, int B,
                    int P, int C, int ldy, int y_off, void* stream);
int mml_relumax_bwd(mml_ctx*, const float* dy, const int32_t* arg, const uint8_t* keep, float keep_scale, uint16_t* dconv, float* dbias, int B,
                    int P, int C, int ldy, int y_off, void* stream);
/* small-batch dense layer (textcnn.py:25-28 embd, classifier.py:100-117): y = dropout(relu(x W^T + b)); x [B][ldx], y [B][ldy] (so the
 * concatenation of embeddings is a column offset), keep uint8 [B][N] or NULL.  Backward (dy [B][lddy] is overwritten by the gradient at the
 * pre-activation): dx [B][lddx] (optional), dw [N][K], db [N] (stored). */
int mml_dense_fwd(mml_ctx*, const float* x, int ldx, const float* w, const float* bias, const uint8_t* keep, float keep_scale, int relu,
                  float* y, int ldy, int B, int K, int N, void* stream);
int mml_dense_bwd(mml_ctx*, float* dy, int lddy, const float* y, int ldy, const uint8_t* keep, float keep_scale, int relu, const float* x, int ldx,
                  const float* w, float* dx, int lddx, float* dw, float* db, int B, int K, int N, void* stream);
/* torch.nn.utils.clip_grad_norm_ (utt_fusion.py:181-182) folded into the optimizer: norm = ||g||_2 * base_scale over the flat gradient
 * buffer, hyper[row][5] = base_scale * min(1, clip / (norm + 1e-6)) for rows 0..groups-1 (the Adam kernel multiplies gradients by it);
 * partial: MML_CLIP_PARTIALS doubles of scratch; norm_out (optional) receives the norm. */
int mml_clip_grad_scale(mml_ctx*, const float* g, int64_t n, float clip, float base_scale, float* hyper, int groups, double* partial,
                        float* norm_out, void* stream);

/* ---- a9: torch.optim.Adam (coupled weight decay) over the flat parameter buffer -- avmnist.py:303 ---------------- */
/* hyper (device, fp32[8]): lr, beta1, beta2, eps, weight_decay, grad_scale, -, -;  step (device int64[1]) holds the number
 * of completed steps: the update uses t = step + 1, and the counter is incremented on the device when advance_step != 0 (so a
 * captured graph advances it; a step split over several parameter ranges advances it with the last range only).
 * p/g/m/v fp32 [n]; p_bf16 (optional) gets bf16(p). */
int mml_adam_step(mml_ctx*, float* p, const float* g, float* m, float* v, uint16_t* p_bf16, int64_t n, const float* hyper,
                  int64_t* step, int advance_step, void* stream);
/* fp32 -> bf16 copy (shadow refresh after load_state_dict) */
int mml_cast_f32_bf16(mml_ctx*, const float* src, uint16_t* dst, int64_t n, void* stream);
/* exact widening of a bf16 GEMM output for an fp32 consumer (MonomodalEncoder around the MMIMDb encoders: encoder -> fp32 classifier) */
int mml_cast_bf16_f32(mml_ctx*, const uint16_t* src, float* dst, int64_t n, void* stream);
/* ---- e: data-parallel gradient all-reduce -- library-owned NCCL communicator (the reference has no distributed code) ------- */
/* One communicator per ctx.  Rank 0 creates the id (mml_comm_unique_id) and hands the 128 bytes to the other ranks by any
 * out-of-band channel (the Python binding uses torch.distributed's store); every rank then calls mml_comm_init.  max_ctas > 0
 * caps the CTAs NCCL may use (ncclConfig_t::maxCTAs): the all-reduces run under the audio encoder's backward and every SM they
 * take is one its persistent kernels lose; <= 0 leaves NCCL's default.  NCCL itself is dlopen'ed (the process's libnccl.so.2). */
#define MML_COMM_ID_BYTES 128
int mml_comm_unique_id(mml_ctx*, uint8_t* id_out /* [MML_COMM_ID_BYTES] */);
int mml_comm_init(mml_ctx*, const uint8_t* id /* [MML_COMM_ID_BYTES] */, int rank, int world, int max_ctas);
int mml_comm_world(const mml_ctx*); /* 0 until mml_comm_init */
/* in-place fp32 sum all-reduce of buf[0, count) over the communicator, enqueued on `stream` (CUDA-graph capturable); the
 * division by the world size happens in mml_adam_step (hyper[5]) */
int mml_allreduce_bucket(mml_ctx*, float* buf, int64_t count, void* stream);
int mml_comm_destroy(mml_ctx*);

/* ---- a13: FedAvg weighted aggregation (no reference implementation exists; McMahan et al.) ---------------------- */
/* out[i] = sum_k weights[k] * clients[k][i];  clients: DEVICE array of K device pointers; weights: device fp32[K] */
int mml_fedavg(mml_ctx*, const float* const* clients, const float* weights, int K, float* out, int64_t n, void* stream);
/* in-place scale (pre-scale for the allreduce variant): x *= weights[idx] */
int mml_scale_inplace(mml_ctx*, float* x, const float* weights, int idx, int64_t n, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MML_B200_H */
