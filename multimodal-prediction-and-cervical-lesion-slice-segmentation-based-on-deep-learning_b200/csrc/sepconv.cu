// Glue kernels of the fused Xception separable-conv chain (xception.py:9-31: relu -> depthwise 3x3 -> bn1 ->
// pointwise 1x1 -> bn2, activate_first=True, training mode).
//
// Neither BatchNorm output is materialised:
//   * bn1 (no ReLU after it) is FOLDED into the pointwise weights:   p = d . (W diag(scale1))^T + W shift1
//     and its backward needs no reduction pass over the activations, because
//         sum_pix dz * d  = sum_o W[o,i] * G[o,i]      with G = dp^T d   (the weight-gradient GEMM's output)
//         sum_pix dz      = 0                           (bn2's backward output dp sums to zero over the pixels)
//     so  dd = scale1 (.) dz - k (.) (d - mean1) comes out of the data-gradient GEMM's epilogue (conv_tc.cu, TcEpi.side);
//   * bn2 is applied on load by the consumer (dwconv_fused.cu), or by affine_act_kernel at the end of a block.
// Everything here is O(C) or O(C_out*C_in) work except the three streaming kernels at the bottom.
#include "colreduce.cuh"
#include "dwconv_fused.cuh"

namespace cvx {

// ---- BatchNorm statistics -> (mean, invstd, scale, shift) (+ running buffers) -----------------------------
// mean_offset (nullable): the tensor whose sums are in `acc` is the true BatchNorm input MINUS a per-channel constant
// (the folded bn1 shift that the pointwise GEMM leaves out, since this BatchNorm removes it again); only the
// running mean sees the difference.
__global__ void bn_affine_kernel(const double* __restrict__ acc, int64_t rows, const float* __restrict__ gamma,
                                 const float* __restrict__ beta, const float* __restrict__ mean_offset,
                                 float* running_mean, float* running_var,
                                 float* __restrict__ mean_out, float* __restrict__ invstd_out, float* __restrict__ scale,
                                 float* __restrict__ shift, int C, float momentum, float eps) {
  pdl_trigger();
  pdl_wait();
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const double n = (double)rows;
  const double mean = acc[c] / n;
  double var = acc[C + c] / n - mean * mean;
  if (var < 0) var = 0;
  const float invstd = (float)(1.0 / sqrt(var + (double)eps));
  mean_out[c] = (float)mean;
  invstd_out[c] = invstd;
  const float sc = gamma[c] * invstd;
  scale[c] = sc;
  shift[c] = beta[c] - (float)mean * sc;
  if (running_mean) {
    const double unbiased = rows > 1 ? var * n / (n - 1.0) : var;
    const double true_mean = mean + (mean_offset ? (double)mean_offset[c] : 0.0);
    running_mean[c] = (float)((1.0 - momentum) * running_mean[c] + momentum * true_mean);
    running_var[c] = (float)((1.0 - momentum) * running_var[c] + momentum * unbiased);
  }
}

// ---- fold bn1 into the pointwise weights: wp[o][i] = W[o][i]*scale[i] (bf16), wpt = wp^T, bias[o] = sum_i W[o][i]*shift[i]
__global__ void __launch_bounds__(128) pw_fold_kernel(const float* __restrict__ w, const float* __restrict__ scale,
                                                      const float* __restrict__ shift, __nv_bfloat16* __restrict__ wp,
                                                      __nv_bfloat16* __restrict__ wpt, float* __restrict__ bias, int cout,
                                                      int cin) {
  pdl_trigger();
  pdl_wait();
  __shared__ float sh[4];
  const int o = blockIdx.x;
  float part = 0.f;
  for (int i = threadIdx.x; i < cin; i += 128) {
    const float wv = w[(size_t)o * cin + i];
    const __nv_bfloat16 h = __float2bfloat16_rn(wv * scale[i]);
    wp[(size_t)o * cin + i] = h;
    wpt[(size_t)i * cout + o] = h;
    part = fmaf(wv, shift[i], part);
  }
  part = warp_sum(part);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = part;
  __syncthreads();
  if (threadIdx.x == 0) bias[o] = sh[0] + sh[1] + sh[2] + sh[3];
}

// ---- bn2 backward coefficients from the RAW sums (sum g, sum g*p):  dp = A*g + B*p + Cc ------------------------
__global__ void bn_bwd_coef_kernel(const double* __restrict__ sums, int64_t rows, const float* __restrict__ mean,
                                   const float* __restrict__ invstd, const float* __restrict__ gamma, float* __restrict__ A,
                                   float* __restrict__ Bc, float* __restrict__ Cc, float* __restrict__ dgamma,
                                   float* __restrict__ dbeta, int C) {
  pdl_trigger();
  pdl_wait();
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const double n = (double)rows;
  const double sg = sums[c], sgp = sums[C + c];
  const double mu = mean[c], is = invstd[c];
  const double dgam = is * (sgp - mu * sg);       // sum g * phat
  const double sc = (double)gamma[c] * is;
  const double b = -sc * is * dgam / n;
  A[c] = (float)sc;
  Bc[c] = (float)b;
  Cc[c] = (float)(-sc * sg / n - b * mu);
  dgamma[c] = (float)dgam;
  dbeta[c] = (float)sg;
}

// ---- bn1 backward from the pointwise weight gradient G[o][i] (w.r.t. the UN-normalised d) ------------------------
// colsum[i] += sum_o W[o][i]*G[o][i] ; dW[o][i] = scale[i]*G[o][i]
__global__ void __launch_bounds__(256) pw_bwd_coef_kernel(const float* __restrict__ G, const float* __restrict__ w,
                                                          const float* __restrict__ scale, float* __restrict__ dW,
                                                          float* __restrict__ colsum, int cout, int cin, int rows_per_block) {
  pdl_trigger();
  pdl_wait();
  __shared__ float red[8][33];
  const int cx = threadIdx.x & 31, ry = threadIdx.x >> 5;
  const int i = blockIdx.x * 32 + cx;
  const int o0 = blockIdx.y * rows_per_block;
  int o1 = o0 + rows_per_block;
  if (o1 > cout) o1 = cout;
  float part = 0.f;
  if (i < cin) {
    const float sc = scale[i];
    for (int o = o0 + ry; o < o1; o += 8) {
      const float g = G[(size_t)o * cin + i];
      part = fmaf(w[(size_t)o * cin + i], g, part);
      dW[(size_t)o * cin + i] = sc * g;
    }
  }
  red[ry][cx] = part;
  __syncthreads();
  if (ry == 0 && i < cin) {
    float s = 0.f;
#pragma unroll
    for (int r = 0; r < 8; ++r) s += red[r][cx];
    atomicAdd(colsum + i, s);
  }
}
// dgamma1 = invstd*colsum ; side_scale = -k ; bias = k*mean with k = scale*invstd*dgamma1/rows ; dbeta1 = 0
__global__ void pw_bwd_finalize_kernel(const float* __restrict__ colsum, const float* __restrict__ scale,
                                       const float* __restrict__ invstd, const float* __restrict__ mean, int64_t rows,
                                       float* __restrict__ dgamma, float* __restrict__ dbeta, float* __restrict__ negk,
                                       float* __restrict__ kmean, int C) {
  pdl_trigger();
  pdl_wait();
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const float dgam = invstd[c] * colsum[c];
  const float k = scale[c] * invstd[c] * dgam / (float)rows;
  dgamma[c] = dgam;
  dbeta[c] = 0.f;
  negk[c] = -k;
  kmean[c] = k * mean[c];
}

// ---- streaming kernels (bf16, one 16-byte channel vector per thread for its whole life) -------------------------
// y = act(scale*p + shift + res)
__global__ void __launch_bounds__(256) affine_act_kernel(const __nv_bfloat16* __restrict__ p, const __nv_bfloat16* __restrict__ res,
                                                         __nv_bfloat16* __restrict__ y, const float* __restrict__ scale,
                                                         const float* __restrict__ shift, int64_t rows, int C, int act,
                                                         int64_t stride_vecs) {
  pdl_trigger();
  pdl_wait();
  const int cvn = C / 8;
  const int64_t total = rows * cvn;
  int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (e >= stride_vecs) return;
  const int c0 = (int)(e % cvn) * 8;
  float sc[8], sh[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { sc[i] = scale[c0 + i]; sh[i] = shift[c0 + i]; }
  for (; e < total; e += stride_vecs) {
    Vec<__nv_bfloat16> v, r;
    v.load(p + e * 8);
    if (res) r.load(res + e * 8);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float o = fmaf(v.v[i], sc[i], sh[i]);
      if (res) o += r.v[i];
      v.v[i] = act_apply(o, act);
    }
    v.store(y + e * 8);
  }
}

struct RawStatsF {
  static constexpr int NACC = 2;
  const __nv_bfloat16* x;
  int C;
  __device__ __forceinline__ void operator()(int64_t row, int c0, float (&acc)[2][8]) const {
    Vec<__nv_bfloat16> v;
    v.load(x + row * C + c0);
#pragma unroll
    for (int i = 0; i < 8; ++i) { acc[0][i] += v.v[i]; acc[1][i] = fmaf(v.v[i], v.v[i], acc[1][i]); }
  }
};

// raw bn backward sums at the END of a block: g = dy * act'(y) ; acc[0] += g ; acc[1] += g*p
struct BnRawBwdF {
  static constexpr int NACC = 2;
  const __nv_bfloat16* dy;
  const __nv_bfloat16* y;  // activation output (mask source) or null
  const __nv_bfloat16* p;
  int C, act;
  __device__ __forceinline__ void operator()(int64_t row, int c0, float (&acc)[2][8]) const {
    Vec<__nv_bfloat16> g, yv, pv;
    g.load(dy + row * C + c0);
    pv.load(p + row * C + c0);
    if (act != CVX_ACT_NONE) yv.load(y + row * C + c0);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float dz = g.v[i];
      if (act != CVX_ACT_NONE) dz *= act_mask(yv.v[i], act);
      acc[0][i] += dz;
      acc[1][i] = fmaf(dz, pv.v[i], acc[1][i]);
    }
  }
};

// dp = A*g + B*p + Cc with g = dy * act'(y) ; optionally also writes g (the gradient of the residual branch)
__global__ void __launch_bounds__(256) bn_bwd_affine_kernel(const __nv_bfloat16* __restrict__ dy, const __nv_bfloat16* __restrict__ y,
                                                            const __nv_bfloat16* __restrict__ p, const float* __restrict__ A,
                                                            const float* __restrict__ Bc, const float* __restrict__ Cc,
                                                            __nv_bfloat16* __restrict__ dp, __nv_bfloat16* __restrict__ gout,
                                                            int64_t rows, int C, int act, int64_t stride_vecs) {
  pdl_trigger();
  pdl_wait();
  const int cvn = C / 8;
  const int64_t total = rows * cvn;
  int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (e >= stride_vecs) return;
  const int c0 = (int)(e % cvn) * 8;
  float a[8], b[8], cc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { a[i] = A[c0 + i]; b[i] = Bc[c0 + i]; cc[i] = Cc[c0 + i]; }
  for (; e < total; e += stride_vecs) {
    Vec<__nv_bfloat16> g, yv, pv;
    g.load(dy + e * 8);
    pv.load(p + e * 8);
    if (act != CVX_ACT_NONE) yv.load(y + e * 8);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float dz = g.v[i];
      if (act != CVX_ACT_NONE) dz *= act_mask(yv.v[i], act);
      g.v[i] = dz;
      pv.v[i] = fmaf(a[i], dz, fmaf(b[i], pv.v[i], cc[i]));
    }
    pv.store(dp + e * 8);
    if (gout) g.store(gout + e * 8);
  }
}

static inline void sep_stream_grid(int64_t total_vecs, int cvn, int* blocks, int64_t* stride) {
  int64_t want = (int64_t)kNumSMs * 8 * 256;
  if (want > total_vecs) want = total_vecs;
  int64_t s = ceil_div64(want, cvn) * cvn;
  *stride = s;
  *blocks = (int)ceil_div64(s, 256);
}

__global__ void d2f_kernel(const double* __restrict__ a, float* __restrict__ out, int n) {
  pdl_trigger();
  pdl_wait();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = (float)a[i];
}

}  // namespace cvx

using namespace cvx;

extern "C" {

int cvx_dwf_fwd(const cvx_conv_desc* d, const void* x, const float* w9c, const float* in_scale, const float* in_shift,
                int relu_in, void* y, double* stats, void* stream) {
  CVX_CHECK_ARG(d && x && w9c && y && (!in_scale == !in_shift), "dwf_fwd: bad arguments");
  const int rc = dwf_fwd_launch(d, x, w9c, in_scale, in_shift, relu_in, y, stats, as_stream(stream));
  if (rc == CVX_EUNSUPPORTED) set_error("dwf_fwd: needs bf16, 3x3, stride 1, dilation 1, pad 1, C %% 8 == 0");
  return rc;
}

int cvx_dwf_bwd(const cvx_conv_desc* d, const void* dd, const void* dside, const float* negk, const float* kmean,
                const void* x, const float* w9c, const float* in_scale, const float* in_shift, int relu_in,
                const void* addend, void* g, float* dw9c, double* ws, double* sums, void* stream) {
  CVX_CHECK_ARG(d && dd && x && w9c && g && dw9c && ws && (!in_scale == !in_shift) && (!dside == !negk) && (!dside == !kmean),
                "dwf_bwd: bad arguments");
  cudaStream_t st = as_stream(stream);
  const int rc = dwf_bwd_launch(d, dd, dside, negk, kmean, x, w9c, in_scale, in_shift, relu_in, addend, g, ws, sums, st);
  if (rc == CVX_EUNSUPPORTED) set_error("dwf_bwd: needs bf16, 3x3, stride 1, dilation 1, pad 1, C %% 8 == 0");
  if (rc) return rc;
  launch_pdl(d2f_kernel, dim3((9 * d->cin + 255) / 256), dim3(256), 0, st, ws, dw9c, 9 * d->cin);
  CVX_LAUNCH_OK();
  return CVX_OK;
}

int cvx_bn_stats(const void* x, double* stats, int64_t rows, int c, int dtype, void* stream) {
  CVX_CHECK_ARG(x && stats && rows > 0 && c > 0, "bn_stats: bad arguments");
  cudaStream_t st = as_stream(stream);
  CVX_WS_ZERO(stats, sizeof(double) * 2 * c, st);
  CVX_CHECK_ARG(dtype == CVX_BF16 && c % 8 == 0, "bn_stats: bf16 with C %% 8 == 0 only");
  return colreduce_launch<__nv_bfloat16, RawStatsF>(RawStatsF{(const __nv_bfloat16*)x, c}, rows, c, stats, st);
}

int cvx_bn_affine(const double* stats, int64_t rows, const float* gamma, const float* beta, const float* mean_offset,
                  float* running_mean, float* running_var, float* mean, float* invstd, float* scale, float* shift, int c,
                  float momentum, float eps, void* stream) {
  CVX_CHECK_ARG(stats && gamma && beta && mean && invstd && scale && shift && rows > 0 && c > 0, "bn_affine: bad arguments");
  launch_pdl(bn_affine_kernel, dim3((c + 127) / 128), dim3(128), 0, as_stream(stream), stats, rows, gamma, beta, mean_offset, running_mean,
                                                                   running_var, mean, invstd, scale, shift, c, momentum, eps);
  CVX_LAUNCH_OK();
  return CVX_OK;
}

int cvx_pw_fold(const float* w, const float* scale, const float* shift, void* wp, void* wpt, float* bias, int cout, int cin,
                void* stream) {
  CVX_CHECK_ARG(w && scale && shift && wp && wpt && bias && cout > 0 && cin > 0, "pw_fold: bad arguments");
  launch_pdl(pw_fold_kernel, dim3(cout), dim3(128), 0, as_stream(stream), w, scale, shift, (__nv_bfloat16*)wp, (__nv_bfloat16*)wpt, bias, cout, cin);
  CVX_LAUNCH_OK();
  return CVX_OK;
}

int cvx_affine_act(const void* p, const void* res, void* y, const float* scale, const float* shift, int64_t rows, int c,
                   int act, void* stream) {
  CVX_CHECK_ARG(p && y && scale && shift && rows > 0 && c > 0 && c % 8 == 0, "affine_act: bad arguments");
  int blocks; int64_t stride;
  sep_stream_grid(rows * (c / 8), c / 8, &blocks, &stride);
  launch_pdl(affine_act_kernel, dim3(blocks), dim3(256), 0, as_stream(stream), (const __nv_bfloat16*)p, (const __nv_bfloat16*)res,
                                                          (__nv_bfloat16*)y, scale, shift, rows, c, act, stride);
  CVX_LAUNCH_OK();
  return CVX_OK;
}

int cvx_bn_bwd_sums(const void* dy, const void* y, const void* p, double* sums, int64_t rows, int c, int act, void* stream) {
  CVX_CHECK_ARG(dy && p && sums && (act == CVX_ACT_NONE || y) && rows > 0 && c > 0 && c % 8 == 0, "bn_bwd_sums: bad arguments");
  cudaStream_t st = as_stream(stream);
  CVX_WS_ZERO(sums, sizeof(double) * 2 * c, st);
  // four resident 256-thread blocks per SM (61 registers): 42 -> 36 us on the middle-flow tensor against three
  return colreduce_launch<__nv_bfloat16, BnRawBwdF, 256, 4>(
      BnRawBwdF{(const __nv_bfloat16*)dy, (const __nv_bfloat16*)y, (const __nv_bfloat16*)p, c, act}, rows, c, sums, st);
}

int cvx_bn_bwd_coef(const double* sums, int64_t rows, const float* mean, const float* invstd, const float* gamma, float* a,
                    float* b, float* cc, float* dgamma, float* dbeta, int c, void* stream) {
  CVX_CHECK_ARG(sums && mean && invstd && gamma && a && b && cc && dgamma && dbeta && rows > 0 && c > 0, "bn_bwd_coef: bad arguments");
  launch_pdl(bn_bwd_coef_kernel, dim3((c + 127) / 128), dim3(128), 0, as_stream(stream), sums, rows, mean, invstd, gamma, a, b, cc, dgamma, dbeta, c);
  CVX_LAUNCH_OK();
  return CVX_OK;
}

int cvx_bn_bwd_affine(const void* dy, const void* y, const void* p, const float* a, const float* b, const float* cc, void* dp,
                      void* gout, int64_t rows, int c, int act, void* stream) {
  CVX_CHECK_ARG(dy && p && a && b && cc && dp && (act == CVX_ACT_NONE || y) && rows > 0 && c > 0 && c % 8 == 0,
                "bn_bwd_affine: bad arguments");
  int blocks; int64_t stride;
  sep_stream_grid(rows * (c / 8), c / 8, &blocks, &stride);
  launch_pdl(bn_bwd_affine_kernel, dim3(blocks), dim3(256), 0, as_stream(stream), (const __nv_bfloat16*)dy, (const __nv_bfloat16*)y,
                                                             (const __nv_bfloat16*)p, a, b, cc, (__nv_bfloat16*)dp,
                                                             (__nv_bfloat16*)gout, rows, c, act, stride);
  CVX_LAUNCH_OK();
  return CVX_OK;
}

int cvx_pw_bwd_coef(const float* g_packed, const float* w, const float* scale, const float* invstd, const float* mean,
                    int64_t rows, float* dw, float* colsum, float* dgamma, float* dbeta, float* negk, float* kmean, int cout,
                    int cin, void* stream) {
  CVX_CHECK_ARG(g_packed && w && scale && invstd && mean && dw && colsum && dgamma && dbeta && negk && kmean && rows > 0,
                "pw_bwd_coef: bad arguments");
  cudaStream_t st = as_stream(stream);
  CVX_CUDA_OK(cudaMemsetAsync(colsum, 0, sizeof(float) * cin, st));
  const int rpb = 64;
  dim3 grid((cin + 31) / 32, (cout + rpb - 1) / rpb);
  launch_pdl(pw_bwd_coef_kernel, dim3(grid), dim3(256), 0, st, g_packed, w, scale, dw, colsum, cout, cin, rpb);
  CVX_LAUNCH_OK();
  launch_pdl(pw_bwd_finalize_kernel, dim3((cin + 127) / 128), dim3(128), 0, st, colsum, scale, invstd, mean, rows, dgamma, dbeta, negk, kmean, cin);
  CVX_LAUNCH_OK();
  return CVX_OK;
}

}  // extern "C"
