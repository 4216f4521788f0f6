/*
 * cervix_b200 - C ABI of the B200-native DeepLabv3+ hot path.
 *
 * The reference (alanchou89/Multimodal-Prediction-and-Cervical-Lesion-Slice-Segmentation-
 * Based-on-Deep-Learning) has no FFI of its own: its hot path is torch.nn modules reaching
 * cuDNN/ATen.  This header is therefore the boundary a maintainer
 * would bind from the reference's Python side with ctypes; every entry
 * point below names the reference call site whose device work it replaces.  Paths are
 * relative to /root/reference/Segmentation/deeplabv3+/.
 *
 * Conventions
 *   - every function returns 0 on success, a negative CVX_E* code otherwise;
 *     cvx_last_error() returns a thread-local message for the last failure;
 *   - all data pointers are CALLER-OWNED DEVICE memory (torch tensor .data_ptr());
 *     the library never allocates, frees or retains them;
 *   - `stream` is a cudaStream_t passed as void* (torch.cuda.current_stream().cuda_stream);
 *     all work is enqueued on it, nothing synchronises the device;
 *   - activations are NHWC; `dtype` selects their storage (CVX_F32 or CVX_BF16), all
 *     accumulation is fp32 (fp64 for BatchNorm/loss reductions);
 *   - dense conv weights are consumed in the packed layout [kh*kw][C_out][C_in]
 *     (cvx_pack_weight), depthwise weights as fp32 [9][C].
 */
#ifndef CERVIX_B200_H_
#define CERVIX_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CVX_ABI_VERSION 1
#if defined(__GNUC__)
#define CVX_API __attribute__((visibility("default")))
#else
#define CVX_API
#endif

enum { CVX_F32 = 0, CVX_BF16 = 1 };
enum { CVX_ACT_NONE = 0, CVX_ACT_RELU = 1, CVX_ACT_RELU6 = 2 };
enum { CVX_OK = 0, CVX_EINVAL = -1, CVX_ECUDA = -2, CVX_EUNSUPPORTED = -3 };

/* Dense convolution geometry (nn.Conv2d with symmetric stride/padding/dilation). */
typedef struct cvx_conv_desc {
  int32_t n, h, w, cin;      /* input  NHWC */
  int32_t cout, kh, kw;      /* filter      */
  int32_t stride, pad, dil;
  int32_t ho, wo;            /* output spatial size */
  int32_t dtype;             /* CVX_F32 | CVX_BF16 : storage of x, packed w and y */
} cvx_conv_desc;

CVX_API int         cvx_abi_version(void);
CVX_API const char* cvx_last_error(void);
/* number of kernels this library has launched so far in this process */
CVX_API int64_t     cvx_launch_count(void);
/* Workspace contract switch.  Reduction entry points (BatchNorm statistics / backward sums, depthwise weight gradient,
 * loss statistics, GEMM-epilogue statistics) accumulate into caller-owned fp64 workspaces and clear them first.  A caller
 * that hands out workspaces from an arena it has already zeroed (one fill per training step instead of ~360 memset nodes)
 * sets on = 1 for the duration of that step; the library then skips its own cudaMemsetAsync.  Process-wide, not
 * per-stream: one training step at a time per process (one process per GPU). */
CVX_API int         cvx_set_ws_prezeroed(int on);
/* Profiling aid (tools/tc_trace.py), not part of the operator surface: enable != 0 makes the following CTA-pair tcgen05
 * forward / data-gradient launches leave clock64 phase marks of their first and last cluster; out64 (128 values, host
 * memory, nullable) receives the marks left so far after a device synchronise (slot layout: csrc/conv_tc.cu). */
CVX_API int         cvx_debug_tc_trace(int enable, unsigned long long* out64);
/* 1 if the running device is sm_100 (tcgen05/TMA kernels usable), 0 otherwise */
CVX_API int         cvx_device_is_sm100(void);

/* ---- layout plumbing -------------------------------------------------------------- */
/* NCHW fp32 -> NHWC dtype : entry of DeepLab.forward (nets/deeplabv3_plus.py:169). */
CVX_API int cvx_nchw_to_nhwc(const float* src, void* dst, int n, int c, int h, int w, int dtype, void* stream);
CVX_API int cvx_nhwc_to_nchw(const void* src, float* dst, int n, int c, int h, int w, int dtype, void* stream);
/* nn.Conv2d.weight OIHW fp32 -> packed [tap][cout][cin] dtype.  transpose_flip=1 gives the
 * data-gradient operand [flipped tap][cin][cout]. */
CVX_API int cvx_pack_weight(const float* w_oihw, void* dst, int cout, int cin, int kh, int kw, int dtype,
                    int transpose_flip, void* stream);
/* packed fp32 weight gradient [tap][cout][cin] -> OIHW fp32 (.grad layout) */
CVX_API int cvx_unpack_wgrad(const float* g_packed, float* g_oihw, int cout, int cin, int kh, int kw, void* stream);
/* depthwise nn.Conv2d.weight [C,1,3,3] <-> fp32 [9][C] */
CVX_API int cvx_pack_dw_weight(const float* w_c133, float* dst, int c, void* stream);
CVX_API int cvx_unpack_dw_wgrad(const float* g_9c, float* g_c133, int c, void* stream);
/* channel-slice copy: dst[row, dst_coff : dst_coff+c] = src[row, src_coff : src_coff+c]
 * (torch.cat / its backward on NHWC; nets/deeplabv3_plus.py:110,185) */
CVX_API int cvx_copy_channels(const void* src, int src_ld, int src_coff, void* dst, int dst_ld, int dst_coff,
                      int64_t rows, int c, int dtype, void* stream);

/* ---- dense convolution, generic SIMT path (any shape; the fp32 parity path) ----------- */
/* nn.Conv2d forward (xception.py:95,99,44,16; deeplabv3_plus.py:59-87,149-167) */
CVX_API int cvx_conv_fwd(const cvx_conv_desc* d, const void* x, const void* w_packed, const float* bias,
                 void* y, void* stream);
/* data gradient; w_packed_t is the transpose_flip=1 packing */
CVX_API int cvx_conv_dgrad(const cvx_conv_desc* d, const void* dy, const void* w_packed_t, void* dx, void* stream);
/* weight gradient, ACCUMULATED into fp32 dw_packed[tap][cout][cin] (caller zeroes it) */
CVX_API int cvx_conv_wgrad(const cvx_conv_desc* d, const void* x, const void* dy, float* dw_packed, void* stream);
/* bias gradient: dbias[c] = sum over rows of dy (overwrites) */
CVX_API int cvx_bias_grad(const void* dy, float* dbias, double* ws, int64_t rows, int c, int dtype, void* stream);

/* ---- dense convolution, tcgen05/TMEM/TMA implicit GEMM (bf16, sm_100a) ---------------- */
/* Same contracts as the three calls above; stride must be 1 (stride-2 1x1 convs are fed a
 * pre-subsampled input by the host).  Returns CVX_EUNSUPPORTED for shapes the tensor-core
 * path does not take (C_in or C_out not a multiple of 8, dtype != bf16). */
CVX_API int cvx_conv_fwd_tc(const cvx_conv_desc* d, const void* x, const void* w_packed, const float* bias,
                    void* y, void* stream);
CVX_API int cvx_conv_dgrad_tc(const cvx_conv_desc* d, const void* dy, const void* w_packed_t, void* dx, void* stream);
CVX_API int cvx_conv_wgrad_tc(const cvx_conv_desc* d, const void* x, const void* dy, float* dw_packed, void* stream);
/* Tuning knob of cvx_conv_fwd_tc / cvx_conv_dgrad_tc: CTA pairs per thread-block cluster (1, 2 or 4) that share one
 * weight tile through TMA multicast; 0 = the library's per-shape choice (2 pairs where the reduction rows are long and
 * not 128-byte aligned, e.g. 728 channels, else 1 - measured on B200).  force != 0 keeps an
 * explicit setting even for problems too small to fill the machine (tests).  CERVIX_TC_PAIRS in the environment sets
 * the initial value. */
CVX_API int cvx_conv_tc_set_pairs(int pairs, int force);
/* im2col of a narrow-channel input (C_in <= 4): patches[n,ho,wo,kpad], k = tap*C_in + ci, zero padded
 * to kpad (a multiple of 8).  Turns the ResNet-101 7x7/2 stem of the classifier's patch encoder
 * (MM/Graph_Structure(data_augmentation).py:136) into a 1x1 tensor-core GEMM. */
CVX_API int cvx_im2col_narrow(const cvx_conv_desc* d, const void* x, void* patches, int kpad, void* stream);
/* spatial subsample x[:, ::s, ::s, :] and its scatter-back (1x1 stride-2 skip convs,
 * xception.py:44) */
CVX_API int cvx_subsample(const void* x, void* y, int n, int h, int w, int c, int s, int dtype, void* stream);
CVX_API int cvx_subsample_bwd(const void* dy, void* dx, int n, int h, int w, int c, int s, int dtype, void* stream);

/* ---- depthwise 3x3 (xception.py:13, mobilenetv2.py:39,58) --------------------------- */
/* relu_in=1 applies ReLU to x on load (SeparableConv2d.relu0, xception.py:22-23) */
CVX_API int cvx_dwconv_fwd(const cvx_conv_desc* d, const void* x, const float* w9c, void* y, int relu_in, void* stream);
CVX_API int cvx_dwconv_bwd_data(const cvx_conv_desc* d, const void* dy, const float* w9c, const void* x,
                        void* dx, int relu_in, void* stream);
/* dw9c (fp32 [9][C]) is overwritten; ws needs 9*C doubles */
CVX_API int cvx_dwconv_bwd_weight(const cvx_conv_desc* d, const void* x, const void* dy, float* dw9c,
                          double* ws, int relu_in, void* stream);

/* ---- BatchNorm2d (+ residual add + activation) -------------------------------------- */
/* y = act(bn(x) + residual).  training=1: batch statistics, running buffers updated with
 * `momentum` (unbiased variance), save_mean/save_invstd written for backward.
 * training=0: running statistics.  ws: 2*C + 2 doubles of scratch (sums + grid-barrier counter). */
CVX_API int cvx_bn_forward(const void* x, const void* residual, void* y, const float* gamma, const float* beta,
                   float* running_mean, float* running_var, float* save_mean, float* save_invstd,
                   double* ws, int64_t rows, int c, int dtype, int act, int training,
                   float momentum, float eps, void* stream);
/* dx (and dres = d(act) if non-null), dgamma, dbeta (overwritten).  The activation mask (act != NONE) is read off y,
 * or - when beta is given (only valid if the forward had no residual; dres must be NULL) - recomputed from x with the
 * forward's own fma(x, gamma*invstd, beta - mean*gamma*invstd), so that y is not read at all (5 instead of 7 passes
 * over the tensor).  training=0 treats mean/invstd as constants. */
CVX_API int cvx_bn_backward(const void* dy, const void* x, const void* y, const float* gamma, const float* beta,
                    const float* save_mean, const float* save_invstd, void* dx, void* dres,
                    float* dgamma, float* dbeta, double* ws, int64_t rows, int c, int dtype,
                    int act, int training, void* stream);

/* ---- small bandwidth ops ------------------------------------------------------------ */
CVX_API int cvx_relu_fwd(const void* x, void* y, int64_t n, int dtype, void* stream);
CVX_API int cvx_relu_bwd(const void* dy, const void* y, void* dx, int64_t n, int dtype, void* stream);
CVX_API int cvx_add(const void* a, const void* b, void* out, int64_t n, int dtype, void* stream);
/* y[n,c] = scale * sum_hw x[n,hw,c]  (ASPP global pooling, deeplabv3_plus.py:101-102) */
CVX_API int cvx_spatial_reduce(const void* x, void* y, int n, int hw, int c, float scale, int dtype, void* stream);
/* y[n,hw,c] = scale * x[n,c]  (1x1 -> HxW "bilinear" broadcast, deeplabv3_plus.py:106) */
CVX_API int cvx_spatial_broadcast(const void* x, void* y, int n, int hw, int c, float scale, int dtype, void* stream);
/* F.interpolate(mode='bilinear', align_corners=True) on NHWC (deeplabv3_plus.py:184) */
CVX_API int cvx_upsample_fwd(const void* x, void* y, int n, int hi, int wi, int ho, int wo, int c, int dtype, void* stream);
CVX_API int cvx_upsample_bwd(const void* dy, void* dx, int n, int hi, int wi, int ho, int wo, int c, int dtype, void* stream);
/* The decoder's x4 upsample written straight into the concat buffer (deeplabv3_plus.py:184-185: F.interpolate followed
 * by torch.cat): channels [c_off, c_off + c) of y[n,ho,wo,c_total] = bilinear(x[n,hi,wi,c]); and its gradient read from
 * the same channel slice of dy[n,ho,wo,c_total].  c, c_total, c_off multiples of the 16-byte vector (8 bf16 / 4 fp32). */
CVX_API int cvx_upsample_into(const void* x, void* y, int n, int hi, int wi, int ho, int wo, int c, int c_total, int c_off,
                              int dtype, void* stream);
CVX_API int cvx_upsample_from_bwd(const void* dy, void* dx, int n, int hi, int wi, int ho, int wo, int c, int c_total,
                                  int c_off, int dtype, void* stream);
/* final upsample fused with the NHWC->NCHW fp32 conversion (deeplabv3_plus.py:187) */
CVX_API int cvx_upsample_to_nchw_fwd(const void* x, float* y, int n, int hi, int wi, int ho, int wo, int c, int dtype, void* stream);
CVX_API int cvx_upsample_to_nchw_bwd(const float* dy, void* dx, int n, int hi, int wi, int ho, int wo, int c, int dtype, void* stream);
/* nn.MaxPool2d(3, stride 2, padding 1) of the torchvision ResNet-101 patch encoder
 * (MM/Graph_Structure(data_augmentation).py:136); the gradient goes to the first maximum. */
CVX_API int cvx_maxpool3x3s2_fwd(const void* x, void* y, int n, int h, int w, int c, int dtype, void* stream);
CVX_API int cvx_maxpool3x3s2_bwd(const void* x, const void* y, const void* dy, void* dx, int n, int h, int w, int c,
                         int dtype, void* stream);
/* nn.Dropout (deeplabv3_plus.py:159,165): y = keep ? x / (1-p) : 0, where element i is kept iff a counter-based hash of
 * (seed, *step_dev, i) clears the threshold p * 2^32.  mask (nullable): one byte per element, written when given.
 * step_dev (nullable device int): mixed into the seed so that a CUDA-graph replay draws a new mask every step.
 * cvx_dropout_bwd_seeded recomputes the decisions from the same (seed, step_dev) instead of reading a stored mask. */
CVX_API int cvx_dropout_fwd(const void* x, void* y, uint8_t* mask, int64_t n, float p, uint64_t seed, const int* step_dev,
                    int dtype, void* stream);
CVX_API int cvx_dropout_bwd(const void* dy, const uint8_t* mask, void* dx, int64_t n, float p, int dtype, void* stream);
CVX_API int cvx_dropout_bwd_seeded(const void* dy, void* dx, int64_t n, float p, uint64_t seed, const int* step_dev, int dtype,
                                   void* stream);

/* ---- segmentation objective (nets/deeplabv3_training.py:9-56, utils/utils_metrics.py:13-35) */
/* One pass over fp32 NCHW logits [n,c,h,w] and int64 targets [n,h,w] (value c = ignore).
 * onehot (fp32 [n,h,w,c+1], may be NULL -> derived from target) feeds dice / f_score.
 * stats (double, 4 + 6*c entries, zeroed by the call):
 *   [0] sum w_t*nll  [1] sum w_t (valid)  [2] sum focal_i  [3] pixel count
 *   [4..4+c) tp  [4+c..) sum p  [4+2c..) sum t   then the same three for the hard mask. */
CVX_API int cvx_seg_loss_stats(const float* logits, const int64_t* target, const float* onehot,
                       const float* cls_weights, double* stats, int n, int c, int h, int w,
                       float focal_alpha, float focal_gamma, float threshold, void* stream);
/* results[0..3] = CE, focal, dice, f_score from stats (device -> device, fp32) */
CVX_API int cvx_seg_loss_finalize(const double* stats, float* results, int c, float beta, float smooth, void* stream);
/* dlogits = g[0]*dCE + g[1]*dFocal + g[2]*dDice, g = 3 device floats (upstream grads) */
CVX_API int cvx_seg_loss_grad(const float* logits, const int64_t* target, const float* onehot,
                      const float* cls_weights, const double* stats, const float* g, float* dlogits,
                      int n, int c, int h, int w, float focal_alpha, float focal_gamma,
                      float beta, float smooth, void* stream);

/* ---- fused Xception separable-conv chain (bf16, training mode; xception.py:9-31,33-73) -----------------
 * relu -> depthwise 3x3 -> bn1 -> pointwise 1x1 -> bn2 with NEITHER BatchNorm output materialised: bn1 is folded
 * into the pointwise weights, bn2 is applied on load by its consumer (csrc/sepconv.cu has the algebra).
 * All statistics buffers are fp64 [2][C] and are zeroed by the call that fills them. */
/* cvx_conv_fwd_tc with epilogue extras: y = conv + bias[c] + side_scale[c]*side[row][c]; stats += (sum y, sum y^2) */
CVX_API int cvx_conv_fwd_tc_ex(const cvx_conv_desc* d, const void* x, const void* w_packed, const float* bias,
                       const void* side, const float* side_scale, double* stats, void* y, void* stream);
/* Inference form: y = act(conv + bias[c] + side_scale[c]*side[row][c]) with act = CVX_ACT_NONE / CVX_ACT_RELU
 * (side_scale == NULL: plain residual add).  With an
 * eval-mode BatchNorm folded into w_packed / bias and the residual branch as the side input this is a whole
 * conv -> bn -> (+identity) -> relu group of torchvision's Bottleneck (the classifier's ResNet-101 patch encoder,
 * MM/Graph_Structure(data_augmentation).py:136-168) in one kernel. */
CVX_API int cvx_conv_fwd_tc_act(const cvx_conv_desc* d, const void* x, const void* w_packed, const float* bias,
                        const void* side, const float* side_scale, int act, void* y, void* stream);
/* cvx_conv_dgrad_tc with the same epilogue (side has the shape of dx) */
CVX_API int cvx_conv_dgrad_tc_ex(const cvx_conv_desc* d, const void* dy, const void* w_packed_t, const float* bias,
                         const void* side, const float* side_scale, void* dx, void* stream);
/* d = dw3x3(act(in_scale*x + in_shift)) (in_scale/in_shift nullable = identity; zero padding applies AFTER the
 * affine), stats (nullable) += (sum d, sum d^2) */
CVX_API int cvx_dwf_fwd(const cvx_conv_desc* d, const void* x, const float* w9c, const float* in_scale,
                const float* in_shift, int relu_in, void* y, double* stats, void* stream);
/* both gradients of that layer in one pass over (dd, x): g = dwT(dd') * 1[in_scale*x+in_shift > 0] (+ addend),
 * dw9c fp32 [9][C] (ws: 9*C doubles), sums (nullable) += (sum g, sum g*x).  The incoming gradient is
 * dd' = dd + negk*dside + kmean inside the image when dside/negk/kmean are given (bn1's backward applied on load:
 * dd = the pointwise conv's data gradient, dside = this layer's forward output), else dd itself. */
CVX_API int cvx_dwf_bwd(const cvx_conv_desc* d, const void* dd, const void* dside, const float* negk, const float* kmean,
                const void* x, const float* w9c, const float* in_scale, const float* in_shift, int relu_in,
                const void* addend, void* g, float* dw9c, double* ws, double* sums, void* stream);
/* stats = (sum x, sum x^2) over the rows of a [rows][c] bf16 tensor */
CVX_API int cvx_bn_stats(const void* x, double* stats, int64_t rows, int c, int dtype, void* stream);
/* batch statistics -> mean, invstd, scale = gamma*invstd, shift = beta - mean*scale; running buffers (nullable)
 * updated like nn.BatchNorm2d in training mode.  mean_offset (nullable): per-channel constant that was left out of
 * the measured tensor (it only moves the running mean). */
CVX_API int cvx_bn_affine(const double* stats, int64_t rows, const float* gamma, const float* beta, const float* mean_offset,
                  float* running_mean, float* running_var, float* mean, float* invstd, float* scale, float* shift, int c,
                  float momentum, float eps, void* stream);
/* wp[o][i] = bf16(W[o][i]*scale[i]), wpt = wp^T, bias[o] = sum_i W[o][i]*shift[i]  (1x1 conv weight W fp32 [cout][cin]) */
CVX_API int cvx_pw_fold(const float* w, const float* scale, const float* shift, void* wp, void* wpt, float* bias, int cout,
                int cin, void* stream);
/* y = act(scale*p + shift + res)  (res nullable) */
CVX_API int cvx_affine_act(const void* p, const void* res, void* y, const float* scale, const float* shift, int64_t rows,
                   int c, int act, void* stream);
/* sums = (sum g, sum g*p) with g = dy * act'(y)  (y nullable when act == NONE) */
CVX_API int cvx_bn_bwd_sums(const void* dy, const void* y, const void* p, double* sums, int64_t rows, int c, int act,
                    void* stream);
/* from those raw sums: dp = a*g + b*p + cc (BatchNorm backward as an affine map), dgamma, dbeta */
CVX_API int cvx_bn_bwd_coef(const double* sums, int64_t rows, const float* mean, const float* invstd, const float* gamma,
                    float* a, float* b, float* cc, float* dgamma, float* dbeta, int c, void* stream);
CVX_API int cvx_bn_bwd_affine(const void* dy, const void* y, const void* p, const float* a, const float* b, const float* cc,
                      void* dp, void* gout, int64_t rows, int c, int act, void* stream);
/* bn1's backward from the pointwise weight gradient G [cout][cin] w.r.t. the un-normalised input: dw = G*scale,
 * dgamma, dbeta (= 0), and the data-gradient epilogue vectors negk (side_scale) / kmean (bias).  colsum: cin floats. */
CVX_API int cvx_pw_bwd_coef(const float* g_packed, const float* w, const float* scale, const float* invstd,
                    const float* mean, int64_t rows, float* dw, float* colsum, float* dgamma, float* dbeta,
                    float* negk, float* kmean, int cout, int cin, void* stream);

/* ---- multimodal fusion head: fp32 row operators on [groups*seg, C] matrices ----------------------
 * (MultiModal Prediction/Four_Modal/my_mae_model.py:500-793; a "group" is one patient graph).
 * Linear layers use cvx_conv_fwd/dgrad/wgrad with a 1x1 geometry. */
/* mode 0 = torch_geometric LayerNorm(mode='graph') (my_mae_model.py:350,396,471-478): statistics over
 * all seg*C elements of a group, eps added to the std; mode 1 = nn.LayerNorm (mae_utils.py:112,118), seg = 1.
 * stats: 3 floats per group (mean, multiplier, std), kept for the backward. */
CVX_API int cvx_seg_layernorm_fwd(const float* x, const float* w, const float* b, float* y, float* stats, int groups,
                          int seg, int c, float eps, int mode, void* stream);
CVX_API int cvx_seg_layernorm_bwd(const float* dy, const float* x, const float* w, const float* stats, float* dx,
                          float* dw, float* db, int groups, int seg, int c, float eps, int mode, void* stream);
/* nn.GELU (erf form; mae_utils.py:38-55, my_mae_model.py:338-343) */
CVX_API int cvx_gelu_fwd(const float* x, float* y, int64_t n, void* stream);
CVX_API int cvx_gelu_bwd(const float* dy, const float* x, float* dx, int64_t n, void* stream);
/* SAGEConv mean aggregation over a fixed topology shared by all groups (CSR rowptr/col + per-edge weight):
 * out[g,i,:] = sum_e w[e] * x[g,col[e],:]  (my_mae_model.py:404-416,544; the transposed CSR gives the gradient) */
CVX_API int cvx_graph_gather(const float* x, float* out, int groups, int nodes, int c, const int* rowptr, const int* col,
                     const float* w, void* stream);
/* my_GlobalAttention (my_mae_model.py:35-63): att = softmax over the seg nodes of a group (+1e-16), pooled = sum att*x */
CVX_API int cvx_gate_pool_fwd(const float* x, const float* gate, float* pooled, float* att, int groups, int seg, int c,
                      void* stream);
CVX_API int cvx_gate_pool_bwd(const float* dpooled, const float* x, const float* att, float* dx, float* dgate, int groups,
                      int seg, int c, void* stream);
/* multi-head attention over <= 8 tokens (mae_utils.py:58-102): qkv [b,n,3,h,d] -> out [b,n,h*d], probs [b,h,n,n].
 * step_dev (nullable device int) is mixed into the attention-dropout seed, as in cvx_dropout_fwd, so that a CUDA-graph
 * replay draws a new mask every step; forward and backward of one step must see the same value. */
CVX_API int cvx_attn_small_fwd(const float* qkv, float* out, float* probs, int b, int n, int h, int d, float scale,
                       float drop_p, uint64_t seed, const int* step_dev, void* stream);
CVX_API int cvx_attn_small_bwd(const float* dout, const float* qkv, const float* probs, float* dqkv, int b, int n, int h,
                       int d, float scale, float drop_p, uint64_t seed, const int* step_dev, void* stream);
/* F.normalize(dim=1) (my_mae_model.py:679) */
CVX_API int cvx_l2norm_fwd(const float* x, float* y, float* norms, int rows, int c, void* stream);
CVX_API int cvx_l2norm_bwd(const float* dy, const float* y, const float* norms, float* dx, int rows, int c, void* stream);
/* token (un)shuffling of the masked auto-encoder (my_mae_model.py:143,318-335): y[i] = idx[i] >= 0 ? x[idx[i]] : fill */
CVX_API int cvx_rows_gather(const float* x, const int* idx, const float* fill, float* y, int rows, int c, void* stream);
/* backward of cvx_rows_gather in gather form (fixed summation order, writes every element of dx [src_rows, c] / dfill) */
CVX_API int cvx_rows_scatter_add(const float* dy, const int* idx, float* dx, float* dfill, int rows, int src_rows, int c,
                         void* stream);
/* ---- segment-table row operators: every modality branch of the fusion head in one launch -------------------------------
 * The node rows of all modality branches are stacked in one [rows, c] matrix (modality-major); segment s = one patient
 * graph of one modality = rows seg_start[s] .. +seg_len[s], with the LayerNorm affine of parameter set seg_set[s]
 * (cvx_param_sets: per-set parameter / gradient pointers and the row range of each set, passed by value).  Replaces the
 * per-modality calls of PyG LayerNorm (my_mae_model.py:396, 546), my_GlobalAttention (:35-63, 550, 657-674) and the
 * "node features += reconstructed token" adds (:636-649).  All index tables are device int32 arrays; sums have a fixed
 * order (no atomics). */
#define CVX_MAX_PARAM_SETS 8
typedef struct cvx_param_sets {
  const float* w[CVX_MAX_PARAM_SETS];
  const float* b[CVX_MAX_PARAM_SETS];
  float* dw[CVX_MAX_PARAM_SETS];            /* backward only (nullable) */
  float* db[CVX_MAX_PARAM_SETS];
  int row_start[CVX_MAX_PARAM_SETS + 1];    /* rows of set i: row_start[i] .. row_start[i + 1] */
  int sets;
} cvx_param_sets;
CVX_API int cvx_segtab_layernorm_fwd(const float* x, const cvx_param_sets* ps, float* y, float* stats, const int* seg_start,
                                     const int* seg_len, const int* seg_set, int segments, int c, float eps, int mode,
                                     void* stream);
CVX_API int cvx_segtab_layernorm_bwd(const float* dy, const float* x, const cvx_param_sets* ps, const float* stats, float* dx,
                                     const int* seg_start, const int* seg_len, const int* seg_set, const int* row_seg,
                                     int segments, int c, float eps, int mode, void* stream);
CVX_API int cvx_segtab_gate_pool_fwd(const float* x, const float* gate, float* pooled, float* att, const int* seg_start,
                                     const int* seg_len, int segments, int max_len, int c, void* stream);
CVX_API int cvx_segtab_gate_pool_bwd(const float* dpooled, const float* x, const float* att, float* dx, float* dgate,
                                     const int* seg_start, const int* seg_len, int segments, int max_len, int c, void* stream);
CVX_API int cvx_segtab_bcast_add(const float* x, const float* t, float* y, const int* row_seg, const int* tok_of_seg, int rows,
                                 int c, void* stream);
CVX_API int cvx_segtab_bcast_add_bwd(const float* dy, float* dt, const int* seg_of_tok, const int* seg_start,
                                     const int* seg_len, int tokens, int c, void* stream);
/* ---- grouped fp32 GEMM: the per-modality linear layers of the fusion head as ONE launch --------------------------------
 * Reference sites: the four SAGEConv (lin_l, lin_r), gate MLPs and per-modality head MLPs of fusion_model_mae_2
 * (MultiModal Prediction/Four_Modal/my_mae_model.py:404-416, 544, 550, 657-674, 706-769) - same topology per modality,
 * separate weights, each an nn.Linear call of its own in the reference.
 * Problem i computes C[m,n] = sum_k A(m,k) * B(n,k) (+ bias[n]) with A(m,k) = a[m*lda_m + k*lda_k], B(n,k) = b[n*ldb_n +
 * k*ldb_k], C(m,n) = c[m*ldc + n]; every operand must be contiguous along k or along its row index.  rowsum (nullable)
 * receives sum_k A(m,k) - the bias gradient when the problem is a weight gradient (A = dY^T).  Up to
 * CVX_MAX_GEMM_PROBLEMS problems per launch; the array is HOST memory, copied into the kernel's parameters (capturable).
 * fp32 FMA, one fixed-order sum per output element (bit-reproducible). */
#define CVX_MAX_GEMM_PROBLEMS 16
typedef struct cvx_gemm_problem {
  const float* a;
  const float* b;
  const float* bias;
  float* c;
  float* rowsum;
  int64_t lda_m, lda_k, ldb_n, ldb_k, ldc;
  int m, n, k, reserved;
} cvx_gemm_problem;
CVX_API int cvx_gemm_grouped(const cvx_gemm_problem* problems, int count, void* stream);
/* classifier objective (my_train(full).py:309-347): loss += weight * mean CE ; masked-row MSE */
CVX_API int cvx_softmax_ce(const float* logits, const int64_t* labels, float* loss, float* dlogits, int b, int k,
                   float weight, void* stream);
CVX_API int cvx_masked_mse(const float* a, const float* b, const uint8_t* sel, float* loss, float* da, float* db, int rows,
                   int c, float weight, float inv_count, void* stream);

/* ---- optimizer (train.py:472-476) ---------------------------------------------------- */
/* torch.optim.Adam semantics (L2 weight decay added to the gradient), fp32 state.
 * step_t is the 1-based step count; grad_scale multiplies the gradient first. */
CVX_API int cvx_adam_step(float* p, const float* g, float* m, float* v, int64_t n, float lr, float beta1,
                  float beta2, float eps, float weight_decay, int step_t, float grad_scale, void* stream);
/* same update with hyper = device float[6] {lr, beta1, beta2, eps, weight_decay, grad_scale} and the 1-based step
 * count in device memory: safe to capture in a CUDA graph and replay */
CVX_API int cvx_adam_step_dev(float* p, const float* g, float* m, float* v, int64_t n, const float* hyper, const int* step,
                      void* stream);
/* torch.optim.SGD(momentum, nesterov) semantics */
/* torch.optim.SGD(momentum, nesterov, weight_decay) (train.py:475) with lr / momentum / weight decay / gradient scale read
 * from device memory: hyper = {lr, momentum, -, -, weight_decay, grad_scale} (the layout of cvx_adam_step_dev), so a
 * captured CUDA graph follows the per-epoch learning-rate schedule (train.py:575).  buf starts zeroed. */
CVX_API int cvx_sgd_step_dev(float* p, const float* g, float* buf, int64_t n, const float* hyper, int nesterov, void* stream);
CVX_API int cvx_sgd_step(float* p, const float* g, float* buf, int64_t n, float lr, float momentum,
                 float weight_decay, int nesterov, int first_step, float grad_scale, void* stream);

/* Gather of n gradient tensors into one flat fp32 buffer in a single launch (the engine's replacement for autograd's
 * per-parameter accumulate kernels, utils_fit.py:88-90 `loss.backward()` + DDP's bucket copies, train.py:386):
 * src_ptrs[i] = device address of tensor i (0 = no gradient -> zeros), sizes[i] its element count, dst_offsets[i] its
 * place in dst (multiples of 4 floats); the (chunk_tensor, chunk_start) tables enumerate the
 * cvx_multi_gather_chunk()-float chunks of all tensors.  All tables live in device memory (graph-replay safe). */
CVX_API int cvx_multi_gather_chunk(void);
CVX_API int cvx_multi_gather(const int64_t* src_ptrs, const int32_t* chunk_tensor, const int32_t* chunk_start,
                             const int64_t* dst_offsets, const int64_t* sizes, int nchunks, float* dst, void* stream);

/* ---- inference post-processing (reference: deeplab.py:141-154 detect_image, :304-345 get_miou_png) ----------
 * softmax over the c class planes of ONE image's logits [c][h][w] (fp32, NCHW) -> crop (crop_y, crop_x, crop_h, crop_w:
 * the un-letterboxed region) -> bilinear resize to out_h x out_w with cv2.resize(INTER_LINEAR) coordinates ->
 * argmax.  cls[out_h][out_w] receives the class index; probs (nullable) the resized probabilities [out_h][out_w][c]. */
CVX_API int cvx_seg_postprocess(const float* logits, int c, int h, int w, int crop_y, int crop_x, int crop_h, int crop_w,
                                int out_h, int out_w, unsigned char* cls, float* probs, void* stream);
/* Confusion matrix for mIoU / mPA / accuracy (reference: utils_metrics.py:37-47 fast_hist): hist[gt][pred] += 1 over the
 * n pixels whose ground truth is < classes (255 / ignore labels drop out).  hist is [classes][classes] int64 and is
 * ACCUMULATED into, so one matrix collects a whole validation set without leaving the device. */
CVX_API int cvx_confusion_matrix(const unsigned char* pred, const unsigned char* gt, int64_t n, int classes, int64_t* hist,
                                 void* stream);

/* ---- input side -----------------------------------------------------------
 * Classifier patch pipeline (reference: MM/Graph_Structure(data_augmentation).py:145-161 - PIL resize to
 * new_size x new_size, (new_size/patch)^2 crops enumerated x-major, ToTensor + Normalize): images fp32 NCHW
 * [n,3,h,w] in [0,1] -> patches NHWC [n*k*k, patch, patch, 3] (k = new_size/patch, patch index = ix*k + iy) in
 * `dtype`, value = (bilinear(half-pixel centres, no antialias) - mean[c]) / std[c].  mean / std are HOST arrays of c
 * floats. */
CVX_API int cvx_split_patches(const float* images, void* patches, int n, int c, int h, int w, int new_size, int patch,
                              const float* mean, const float* std, int dtype, void* stream);
/* The same pipeline on the DECODED uint8 image [n,h,w,3], reproducing PIL.Image.resize(BILINEAR) exactly (Pillow
 * Resample.c: horizontal then vertical fixed-point pass with 22 fractional bits, uint8 rounding after each pass, a
 * triangle that widens when the image is reduced).  xmin/xcnt [new_size], xk [new_size][xksize] (likewise y*): first
 * source index, tap count and integer weights per output coordinate, device arrays built by the host
 * (multimodal/pil_resample.py).  value = ((u8 / 255) - mean[c]) / std[c] with IEEE divisions. */
CVX_API int cvx_split_patches_u8(const unsigned char* images, void* patches, int n, int h, int w, int c, int new_size,
                                 int patch, const int* xmin, const int* xcnt, const int* xk, int xksize, const int* ymin,
                                 const int* ycnt, const int* yk, int yksize, const float* mean, const float* std, int dtype,
                                 void* stream);
/* Tail of the segmentation loader on the device (reference: SEG/utils/dataloader.py:40-42 + utils/utils.py:63-65
 * preprocess_input): images_u8 (n_image_elems bytes, [n][h][w][3] as PIL / numpy hold them, 16-byte aligned) -> NHWC
 * activation in `dtype` scaled by 1/255 (nullable together with images_out); labels_u8 (n_pixels bytes, nullable together
 * with labels_out) -> int64 with values >= num_classes set to num_classes (the ignore label).  The one-hot expansion of :47 is implicit in
 * cvx_seg_loss_* (onehot == NULL). */
CVX_API int cvx_finish_batch_u8(const unsigned char* images_u8, void* images_out, int64_t n_image_elems,
                                const unsigned char* labels_u8, int64_t* labels_out, int64_t n_pixels, int num_classes,
                                int dtype, void* stream);

/* ---- training augmentation of the segmentation loader on the device ----------------------------
 * Reference: DeeplabDataset.get_random_data, Segmentation/deeplabv3+/utils/dataloader.py:55-154 (PIL bicubic / nearest
 * resize to a jittered size, flip, paste on a 128-grey canvas, cv2.GaussianBlur 5x5, cv2.warpAffine rotation, HSV gain
 * jitter) - reproduced bit for bit on uint8 data.  The random decisions and the size-dependent coefficient tables come
 * from the host (utils/dataloader.py); a batch is described by one cvx_aug_sample per image, in DEVICE memory. */
typedef struct cvx_aug_sample {
  int64_t src_off;        /* bytes into `src`: decoded image, uint8 RGB [ih][iw][3] */
  int64_t lab_off;        /* bytes into `src`: class map, uint8 [ih][iw] */
  int64_t tmp_off;        /* bytes into `tmp`: horizontal-pass result [ih][nw][3] (unused when iw == nw) */
  int32_t xtab, ytab;     /* index into `tables` (int32): per axis first source index [n], tap count [n], weights [n][taps] */
  int32_t xnn, ynn;       /* index into `tables`: nearest-neighbour source index per resized coordinate [nw], [nh] */
  int32_t rot;            /* index into `tables`: adelta[W], bdelta[W], x0[H], y0[H] of cv2.warpAffine (rotate != 0 only) */
  int32_t lut;            /* byte offset into `luts`: hue[256], sat[256], val[256]; -1 = no colour jitter (validation) */
  int32_t ih, iw, nh, nw; /* source size, resized size */
  int32_t xtaps, ytaps;   /* row length of the weight tables */
  int32_t dx, dy;         /* paste position of the resized image on the canvas (may be negative) */
  int32_t flip, blur, rotate;
  int32_t reserved;
} cvx_aug_sample;
/* pass 1: horizontal resample of every sample whose width changes -> tmp; max_elems = max over samples of ih * nw
 * (0: nothing to do).  Weights: Pillow's 22-bit fixed point. */
CVX_API int cvx_aug_resize_rows(const cvx_aug_sample* samples, int batch, const unsigned char* src, const int* tables,
                                unsigned char* tmp, int64_t max_elems, void* stream);
/* pass 2: vertical resample + flip + paste -> canvas uint8 [batch][h][w][3] (fill 128) and labels uint8 [batch][h][w]
 * (nearest resize, fill 0). */
CVX_API int cvx_aug_compose(const cvx_aug_sample* samples, int batch, const unsigned char* src, const unsigned char* tmp,
                            const int* tables, unsigned char* canvas, unsigned char* labels, int h, int w, void* stream);
/* cv2.GaussianBlur(img, (5, 5), 0) of the samples with blur != 0 -> out (same layout as canvas; other samples untouched) */
CVX_API int cvx_aug_blur5(const cvx_aug_sample* samples, int batch, const unsigned char* canvas, unsigned char* out, int h,
                          int w, void* stream);
/* rotation (samples with rotate != 0; bicubic for the image with border 128, nearest for the label with border 0) and the
 * HSV gain jitter (samples with lut >= 0) -> out_img, out_lab.  A sample reads `blurred` if blur != 0, else `canvas`.
 * cubic = OpenCV's 15-bit bicubic weights [32][32][16]; vec_cols = (w / 32) * 32: the columns OpenCV's vector loop covers. */
CVX_API int cvx_aug_rotate_jitter(const cvx_aug_sample* samples, int batch, const unsigned char* canvas,
                                  const unsigned char* blurred, const unsigned char* labels, const int* tables,
                                  const short* cubic, const unsigned char* luts, unsigned char* out_img,
                                  unsigned char* out_lab, int h, int w, int vec_cols, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* CERVIX_B200_H_ */
