/* discogan_b200.h -- C ABI of libdiscogan_b200.so (hand-written sm_100a kernels for the DiscoGAN train step).
 *
 * The reference (fasion-image-generator-project/discogan_modernized) has no FFI: every operation on its hot path
 * is a PyTorch library call made from model.py / image_translation.py.  Each entry point below names the reference
 * call site (file:line, relative to the reference tree) whose arithmetic it replaces.  The Python host layer
 * (discogan_modernized_b200/ops.py) binds these with ctypes; see INTEGRATION.md for the binding a reference
 * maintainer would add.
 *
 * Conventions: raw device pointers, explicit dims, a cudaStream_t; every function returns 0 on success and a
 * non-zero code otherwise (dg_last_error() gives the message); no allocation, no implicit synchronisation, no
 * cuDNN/cuBLAS, no CPU fallback.  Activations are NHWC bf16; images at the module boundary are NCHW fp32;
 * parameters, gradients, statistics and losses are fp32.  "big"/"small" are the two tensors a 4x4 stride-2 layer
 * connects: big = [B,2Hs,2Ws,Cb], small = [B,Hs,Ws,Cs]; the PyTorch weight viewed as W[Cs][Cb][4][4] is the
 * Conv2d weight [out=Cs][in=Cb] and equally the ConvTranspose2d weight [in=Cs][out=Cb].
 */
#ifndef DISCOGAN_B200_H
#define DISCOGAN_B200_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct CUstream_st* dg_stream_t; /* == cudaStream_t */

enum { DG_ACT_NONE = 0, DG_ACT_LRELU = 1, DG_ACT_RELU = 2 };

/* ---- runtime ----
 * The library keeps no mutable launch state: everything a launch depends on is an argument.  What is process-wide is
 * the launch counter (atomic), dg_last_error (thread-local) and per-device caches of immutable facts (SM count,
 * kernel attributes).  Any number of host threads / trainers / devices may call in concurrently. */
const char* dg_last_error(void);
int dg_version(void);
const char* dg_source_hash(void); /* sha256 of the CUDA sources this library was compiled from (set by the build) */
int dg_device_check(void); /* non-zero unless the current device is sm_100 */
long long dg_launch_count(void); /* kernels launched through this library so far (process-wide) */

/* ---- weights: fp32 W[Cs][Cb][4][4] -> bf16 Wd[Cs][16][Cb] (K-major for DOWN) and Wu[Cb][16][Cs] (for UP).
 * Either output may be NULL.  (Parameters: model.py:8-35,80-142.) */
int dg_pack_weights(const float* w, void* wd, void* wu, int Cs, int Cb, dg_stream_t stream);

/* all GEMM weights of a network in one launch: table = device int64 [n][6] {w, wd|0, wu|0, Cs, Cb, running end} */
int dg_pack_weights_multi(const long long* table, int n, long long total_items, dg_stream_t stream);

/* ---- layout converters for feature maps crossing the module boundary (model.py:69 returns NCHW fp32) ---- */
int dg_nhwc_bf16_to_nchw_f32(const void* x, float* y, int B, int HW, int C, dg_stream_t stream);
int dg_nchw_f32_to_nhwc_bf16(const float* x, void* y, int B, int HW, int C, dg_stream_t stream);

/* ---- tensor-core implicit-GEMM convolutions (tcgen05 + TMEM + TMA) ----
 * Per-call options, passed by pointer (NULL = all defaults):
 *   splitk_ws / splitk_ws_bytes  device fp32 workspace [output pixels][N] for split-K of the small-M / large-K layers;
 *                                NULL disables split-K.  Must not be shared by launches that can run concurrently.
 *   block_n   test hook: force the GEMM N tile (64/128/256; 0 = heuristic; 1 = the role-swapped kernel for <= 128
 *             output channels wherever it is eligible)
 *   pair      test hook: CTA pairs (tcgen05 cta_group::2, two M tiles per MMA): 1/0, -1 = default (on)
 *   wgrad_pair  same for the weight-gradient kernel (cta_group::2 pair kernel vs multicast cluster kernel)
 *   stat_accumulate  the *_stats entry points: 0 = stat_part receives one row of partial sums per CTA (finish with
 *             dg_bn_stats_finalize); 1 = stat_part is ONE zero-initialised accumulator pair float[2][N] that every CTA
 *             adds its sums to (red.global.add) -- consumed directly by dg_bn_act_fwd_acc, no finalize launch; split-K
 *             shapes are supported in this mode (the split-K finish kernel produces the sums)
 *   affine_scale / affine_shift / affine_act / affine_slope  (fprop / dgrad entry points) fused epilogue
 *             out = act(acc * scale[c] + shift[c]) per output channel, applied in fp32 before the single bf16 rounding:
 *             eval-mode BatchNorm (dg_bn_eval_coeffs gives scale = gamma/sqrt(var+eps), shift = beta - mean*scale) and the
 *             LeakyReLU / ReLU behind it folded into the convolution, inference.py:149,168-172.  NULL scale = off.
 * Results do not depend on block_n / pair / wgrad_pair (bit-identical except for split-K summation order). */
typedef struct dg_conv_opts {
  void* splitk_ws;
  size_t splitk_ws_bytes;
  int block_n;
  int pair;
  int wgrad_pair;
  int stat_accumulate;
  const float* affine_scale;
  const float* affine_shift;
  int affine_act;
  float affine_slope;
} dg_conv_opts;
int dg_conv_opts_check(const dg_conv_opts* opts); /* 0 if the options are well-formed */
/* nn.Conv2d(ci,co,4,2,1,bias=False) forward, model.py:11-31,84-103 (cuDNN fprop in the reference) */
int dg_conv4x4s2_fprop(const void* x_big, const void* wd, void* z_small, int B, int H, int W, int Cb, int Cs,
                       const dg_conv_opts* opts, dg_stream_t stream);
/* its data gradient (cuDNN bwd-data); also nn.ConvTranspose2d(ci,co,4,2,1) forward, model.py:118-138 */
int dg_conv4x4s2_dgrad(const void* dz_small, const void* wu, void* dx_big, int B, int Hs, int Ws, int Cs, int Cb,
                       const dg_conv_opts* opts, dg_stream_t stream);
/* forward convolutions with the BatchNorm statistics of the output fused in the epilogue: stat_part = float[2*rows*N]
 * (rows = dg_conv_stats_rows, N = output channels) holds per-CTA partial sums / sums of squares of the fp32
 * accumulators; finish with dg_bn_stats_finalize.  mode 0 = Conv2d fprop, 1 = ConvTranspose2d fprop.
 * dg_conv_stats_rows returns 0 for shapes that run split-K (use dg_conv4x4s2_fprop + dg_bn_stats there). */
int dg_conv_stats_rows(int mode, int B, int Hs, int Ws, int Cs, int Cb, const dg_conv_opts* opts);
int dg_conv4x4s2_fprop_stats(const void* x_big, const void* wd, void* z_small, float* stat_part, int B, int H, int W,
                             int Cb, int Cs, const dg_conv_opts* opts, dg_stream_t stream);
int dg_convT4x4s2_fprop_stats(const void* x_small, const void* wu, void* y_big, float* stat_part, int B, int Hs, int Ws,
                              int Cs, int Cb, const dg_conv_opts* opts, dg_stream_t stream);
/* same, fused with the LeakyReLU derivative of the (BN-less) layer that produced x: dx *= (mask>0 ? 1 : slope) */
int dg_conv4x4s2_dgrad_masked(const void* dz_small, const void* wu, void* dx_big, const void* mask, float slope, int B,
                              int Hs, int Ws, int Cs, int Cb, const dg_conv_opts* opts, dg_stream_t stream);
/* weight gradient (cuDNN bwd-filter): dw[Cs][Cb][4][4] = beta*dw + sum small (x) big; needs a workspace */
size_t dg_conv4x4s2_wgrad_workspace(int B, int Hs, int Ws, int Cs, int Cb);
int dg_conv4x4s2_wgrad(const void* small, const void* big, float* dw, float beta, int B, int Hs, int Ws, int Cs,
                       int Cb, void* ws, size_t ws_bytes, const dg_conv_opts* opts, dg_stream_t stream);
/* ConvTranspose2d(4,2,1) aliases: forward == dgrad, dgrad == fprop, wgrad == wgrad(small = x, big = dy) */
int dg_convT4x4s2_fprop(const void* x_small, const void* wu, void* y_big, int B, int Hs, int Ws, int Cs, int Cb,
                        const dg_conv_opts* opts, dg_stream_t stream);
int dg_convT4x4s2_dgrad(const void* dy_big, const void* wd, void* dx_small, int B, int H, int W, int Cb, int Cs,
                        const dg_conv_opts* opts, dg_stream_t stream);
int dg_convT4x4s2_wgrad(const void* x_small, const void* dy_big, float* dw, float beta, int B, int Hs, int Ws, int Cs,
                        int Cb, void* ws, size_t ws_bytes, const dg_conv_opts* opts, dg_stream_t stream);

/* ---- image-side 3-channel layers (direct kernels, fp32 NCHW image <-> bf16 NHWC 64 channels) ----
 * nn.Conv2d(3,64,4,2,1)+LeakyReLU(0.2), model.py:8-9,80-81 */
int dg_conv_c3_in_fwd(const float* x, const float* w, void* y, int B, int S, float slope, dg_stream_t stream);
int dg_conv_c3_in_bwd(const float* x, const float* w, const void* y, const void* dy, float* dx, int dx_accumulate,
                      float* dw, int B, int S, float slope, dg_stream_t stream);
/* nn.ConvTranspose2d(64,3,4,2,1)+Sigmoid, model.py:142-143 */
int dg_convT_c3_out_fwd(const void* x, const float* w, float* y, int B, int S, dg_stream_t stream);
int dg_convT_c3_out_bwd(const void* x, const float* w, const float* y, const float* dy, void* dx, float* dw, int B,
                        int S, dg_stream_t stream);

/* tensor-core path of the same two layers: the image is first repacked (once per use) into a zero-padded NHWC4 bf16
 * image [B,S+2,S+2,4] (optionally multiplied by y(1-y), the Sigmoid derivative); weights: wc bf16 [64][64],
 * wu3 bf16 [16][16][64] from dg_c3_pack_weights */
int dg_c3_pack_weights(const float* w, void* wc, void* wu3, dg_stream_t stream);
int dg_img_pad_nhwc4(const float* img, const float* img2 /* optional, added */, const float* yimg, void* out, int B,
                     int S, dg_stream_t stream);
int dg_c3_down_tc(const void* xp, const void* wc, void* y, int B, int S, int act, float slope, dg_stream_t stream);
int dg_c3_up_tc(const void* x64, const void* wu3, float* img, int B, int S, int sigmoid, int accumulate,
                dg_stream_t stream);
size_t dg_c3_wgrad_workspace(int B, int S);
int dg_c3_wgrad_tc(const void* v64, const void* xp, float* dw, float beta, int B, int S, void* ws, size_t ws_bytes,
                   dg_stream_t stream);

/* ---- 4x4 "valid" heads as skinny products against Wd[Ns][K=16*C] ----
 * nn.Conv2d(C,100,4,1,0) model.py:107, nn.Conv2d(C,1,4,1,0) model.py:35: small = big . Wd^T          (fc_down)
 * nn.ConvTranspose2d(100,C,4,1,0) model.py:114:                          big = small . Wd            (fc_up)
 * wgrad in PyTorch layout [Ns][C][4][4]:                                                            (fc_wgrad) */
int dg_fc_down(const void* big, const void* wd, void* small, int small_f32, int B, int Ns, int K, dg_stream_t stream);
int dg_fc_up(const void* small, int small_f32, const void* wd, void* big, int B, int Ns, int K, dg_stream_t stream);
int dg_fc_wgrad(const void* small, int small_f32, const void* big, float* dw, float beta, int B, int Ns, int C,
                dg_stream_t stream);

/* ---- BatchNorm2d (+LeakyReLU/ReLU), model.py:12-33,84-140 ----
 * stats = float[4*C] {mean, invstd, scale, shift}; scratch = float[dg_bn_scratch_floats(P,C)] */
size_t dg_bn_scratch_floats(long long P, int C);
int dg_bn_stats(const void* z, long long P, int C, const float* gamma, const float* beta, float eps, float momentum,
                float* stats, float* running_mean, float* running_var, float* scratch, dg_stream_t stream);
int dg_bn_stats_finalize(const float* part, int rows, long long P, int C, const float* gamma, const float* beta,
                         float eps, float momentum, float* stats, float* running_mean, float* running_var,
                         dg_stream_t stream);
int dg_bn_eval_coeffs(const float* gamma, const float* beta, const float* running_mean, const float* running_var,
                      float eps, int C, float* stats, dg_stream_t stream);
int dg_bn_act_fwd(const void* z, void* y, long long P, int C, const float* stats, int act, float slope,
                  dg_stream_t stream);
/* folded finalize (two launches fewer per BatchNorm layer and pass): reductions add into ONE zero-initialised
 * accumulator pair per layer -- acc = float[2*C] -- and the consumer derives its coefficients from the sums itself.
 * Producers: the *_stats convolutions with dg_conv_opts.stat_accumulate = 1, or dg_bn_stats_acc (sums of a bf16 tensor).
 * dg_bn_act_fwd_acc also writes stats[4*C] and updates the running statistics; dg_bn_act_bwd_acc (acc2 = float[2*C],
 * zeroed) writes dz and dgamma / dbeta (both may be NULL).  Summation order of the atomics is not reproducible. */
int dg_bn_stats_acc(const void* z, long long P, int C, float* acc, dg_stream_t stream);
int dg_bn_act_fwd_acc(const void* z, void* y, long long P, int C, const float* acc, const float* gamma, const float* beta,
                      float eps, float momentum, float* stats, float* running_mean, float* running_var, int act,
                      float slope, dg_stream_t stream);
int dg_bn_act_bwd_acc(const void* dy, const void* dy2, const float* bcast, float bcast_coef, long long bcast_rows,
                      const void* z, const float* stats, const float* gamma, long long P, int C, int act, float slope,
                      float* dgamma, float* dbeta, float grad_beta, void* dz, float* acc2, dg_stream_t stream);
int dg_bn_act_bwd(const void* dy, const void* dy2, const float* bcast, float bcast_coef, long long bcast_rows,
                  const void* y, const void* z, const float* stats, const float* gamma, long long P, int C, int act,
                  float slope, float* dgamma, float* dbeta, float grad_beta, void* dz, float* coefs, float* scratch,
                  dg_stream_t stream);

/* ---- losses ----
 * nn.Sigmoid (model.py:36) + nn.BCELoss in get_gan_loss, image_translation.py:146-168,268 */
int dg_gan_bce_fwd(const float* logit_real, const float* logit_fake, int B, float* p_real, float* p_fake, float* out2,
                   dg_stream_t stream);
int dg_gan_bce_bwd(const float* p_real, const float* p_fake, int B, float g_dis, float g_gen, float* dlogit_real,
                   float* dlogit_fake, dg_stream_t stream);
int dg_sigmoid_fwd(const float* x, float* y, int n, dg_stream_t stream);
int dg_sigmoid_bwd(const float* y, const float* dy, float* dx, int n, dg_stream_t stream);
/* nn.MSELoss, image_translation.py:267,349-350 */
size_t dg_reduce_scratch_floats(void);
int dg_mse_fwd(const float* a, const float* b, long long n, float* out, float* scratch, dg_stream_t stream);
int dg_mse_bwd(const float* a, const float* b, long long n, float g, float* da, int accumulate, dg_stream_t stream);
/* get_fm_loss, image_translation.py:136-144 (one feature map per call; accumulate sums the layers) */
int dg_fm_fwd(const void* real, const void* fake, int B, long long n, float* diff, float* out, int accumulate,
              float* scratch, dg_stream_t stream);
int dg_fm_bwd(const float* diff, int B, long long n, float g, void* dfeat, dg_stream_t stream);

/* ---- optim.Adam(lr, betas, weight_decay=1e-5), image_translation.py:275-287 ---- */
/* state: device float[4] {steps taken, 1-beta1^t, sqrt(1-beta2^t), -}; zero once, advanced by every call */
int dg_adam_step(float* p, const float* g, float* m, float* v, long long n, float lr, float beta1, float beta2,
                 float eps, float weight_decay, float* state, float grad_scale, dg_stream_t stream);
/* the same in two halves, for piecewise updates (one dg_adam_apply per gradient bucket as its exchange completes):
 * dg_adam_tick advances the step state once per optimiser step, dg_adam_apply updates a range with the current state */
int dg_adam_tick(float* state, float beta1, float beta2, dg_stream_t stream);
int dg_adam_apply(float* p, const float* g, float* m, float* v, long long n, float lr, float beta1, float beta2,
                  float eps, float weight_decay, const float* state, float grad_scale, dg_stream_t stream);

/* ---- input pipeline: read_images / DiscoGANDataset._load_and_process_image, dataset.py:37-73,238-261 ----
 * One launch preprocesses a batch of decoded uint8 RGB images resident in device memory into the trainer's input layout:
 * out = float [n][3][S][S] in [0,1].  table = device int64 [n][8] rows {src pointer (uint8 [H][W][3]), H, W, x0, crop_w,
 * mode, 0, 0}: columns [x0, x0+crop_w) are used (the reference crops halves of side-by-side pairs: domain 'A' = columns
 * [0,256), 'B' = [256,W)); mode 0 = cv2.resize(INTER_LINEAR) of the uint8 image, mode 1 = the domain-'A' edge-map
 * thickening (255-x, 3x3 dilate, 255-x, in float64) followed by cv2.resize on float64.  Bit-exact with the reference's
 * cv2 arithmetic in both modes (tests/test_dataset_gpu.py). */
int dg_preprocess_u8(const long long* table, int n, int S, float* out, dg_stream_t stream);

/* ---- debug-only SIMT versions of the tensor-core convolutions (never the product path) ---- */
int dg_simt_conv4x4s2_fprop(const void* x, const void* wd, void* z, int B, int H, int W, int Cb, int Cs,
                            dg_stream_t stream);
int dg_simt_conv4x4s2_dgrad(const void* dz, const void* wu, void* dx, int B, int Hs, int Ws, int Cs, int Cb,
                            dg_stream_t stream);
int dg_simt_conv4x4s2_wgrad(const void* small, const void* big, float* dw, float beta, int B, int Hs, int Ws, int Cs,
                            int Cb, dg_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* DISCOGAN_B200_H */
