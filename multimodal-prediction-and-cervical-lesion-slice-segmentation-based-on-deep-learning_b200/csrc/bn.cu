// BatchNorm2d (+ residual add + ReLU/ReLU6) forward/backward on NHWC, and bias gradients.
//
// Reference semantics: nn.BatchNorm2d as used at xception.py:14,17,45 /
// mobilenetv2.py:13,20 / deeplabv3_plus.py:61-85,151,157,162: training = biased batch
// variance for normalisation, unbiased into running_var, momentum-weighted running buffers.
// Bandwidth-bound: forward = 1 read (stats) + 1 read/1 write (apply); backward = 2 reads
// (reduce) + 2-3 reads/1-2 writes (apply).  Statistics are reduced in fp64.
#include <stdlib.h>

#include "colreduce.cuh"

namespace cvx {

template <typename T>
struct StatsF {
  static constexpr int NACC = 2;
  const T* x;
  int C;
  __device__ __forceinline__ void operator()(int64_t row, int c0, float (&acc)[2][Elem<T>::kVec]) const {
    Vec<T> v;
    v.load(x + row * C + c0);
#pragma unroll
    for (int i = 0; i < Vec<T>::N; ++i) {
      acc[0][i] += v.v[i];
      acc[1][i] = fmaf(v.v[i], v.v[i], acc[1][i]);
    }
  }
};

template <typename T>
struct SumF {
  static constexpr int NACC = 1;
  const T* x;
  int C;
  __device__ __forceinline__ void operator()(int64_t row, int c0, float (&acc)[1][Elem<T>::kVec]) const {
    Vec<T> v;
    v.load(x + row * C + c0);
#pragma unroll
    for (int i = 0; i < Vec<T>::N; ++i) acc[0][i] += v.v[i];
  }
};

template <typename T>
struct BnBwdF {
  static constexpr int NACC = 2;
  const T* dy;
  const T* x;
  const T* y;  // activation output: source of the mask unless beta is given; may be null when act == NONE
  const float* gamma;
  const float* beta;  // non-null: the mask is recomputed from x exactly as the forward computed y (no residual)
  const float* mean;
  const float* invstd;
  int C, act;
  struct Ctx {
    float mean[Elem<T>::kVec], is[Elem<T>::kVec], sc[Elem<T>::kVec], sh[Elem<T>::kVec];
  };
  __device__ __forceinline__ void init(int c0, Ctx& k) const {
#pragma unroll
    for (int i = 0; i < Elem<T>::kVec; ++i) {
      k.mean[i] = mean[c0 + i];
      k.is[i] = invstd[c0 + i];
      k.sc[i] = beta ? gamma[c0 + i] * k.is[i] : 0.f;
      k.sh[i] = beta ? beta[c0 + i] - k.mean[i] * k.sc[i] : 0.f;
    }
  }
  __device__ __forceinline__ void operator()(int64_t row, int c0, float (&acc)[2][Elem<T>::kVec], const Ctx& k) const {
    Vec<T> g, xv, yv;
    g.load(dy + row * C + c0);
    xv.load(x + row * C + c0);
    const bool from_y = act != CVX_ACT_NONE && beta == nullptr;
    if (from_y) yv.load(y + row * C + c0);
#pragma unroll
    for (int i = 0; i < Vec<T>::N; ++i) {
      float dz = g.v[i];
      if (act != CVX_ACT_NONE) dz *= act_mask(from_y ? yv.v[i] : fmaf(xv.v[i], k.sc[i], k.sh[i]), act);
      acc[0][i] += dz;
      acc[1][i] = fmaf(dz, xv.v[i] - k.mean[i], acc[1][i]);   // x 1/sigma in finish(): sum dz * xhat
    }
  }
  __device__ __forceinline__ void finish(float (&acc)[2][Elem<T>::kVec], const Ctx& k) const {
#pragma unroll
    for (int i = 0; i < Elem<T>::kVec; ++i) acc[1][i] *= k.is[i];
  }
};

__global__ void bn_finalize_kernel(const double* __restrict__ acc, int64_t rows, float* running_mean,
                                   float* running_var, float* save_mean, float* save_invstd, int C, int training,
                                   float momentum, float eps) {
  pdl_trigger();
  pdl_wait();
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  if (training) {
    const double n = (double)rows;
    const double mean = acc[c] / n;
    double var = acc[C + c] / n - mean * mean;
    if (var < 0) var = 0;
    save_mean[c] = (float)mean;
    save_invstd[c] = (float)(1.0 / sqrt(var + (double)eps));
    if (running_mean) {
      const double unbiased = rows > 1 ? var * n / (n - 1.0) : var;
      running_mean[c] = (float)((1.0 - momentum) * running_mean[c] + momentum * mean);
      running_var[c] = (float)((1.0 - momentum) * running_var[c] + momentum * unbiased);
    }
  } else {
    save_mean[c] = running_mean[c];
    save_invstd[c] = 1.0f / sqrtf(running_var[c] + eps);
  }
}

template <typename T>
__global__ void __launch_bounds__(256) bn_apply_kernel(const T* __restrict__ x, const T* __restrict__ res,
                                                       T* __restrict__ y, const float* __restrict__ gamma,
                                                       const float* __restrict__ beta,
                                                       const float* __restrict__ mean,
                                                       const float* __restrict__ invstd, int64_t rows, int C,
                                                       int act, int64_t stride_vecs) {
  pdl_trigger();
  pdl_wait();
  constexpr int VEC = Elem<T>::kVec;
  const int cvn = C / VEC;
  const int64_t total = rows * cvn;
  int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (e >= stride_vecs) return;  // stride is a multiple of cvn: each thread owns one channel vector
  const int c0 = (int)(e % cvn) * VEC;
  float sc[VEC], sh[VEC];
#pragma unroll
  for (int i = 0; i < VEC; ++i) {
    sc[i] = gamma[c0 + i] * invstd[c0 + i];
    sh[i] = beta[c0 + i] - mean[c0 + i] * sc[i];
  }
  for (; e < total; e += stride_vecs) {
    Vec<T> v, r;
    v.load(x + e * VEC);
    if (res) r.load(res + e * VEC);
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
      float o = fmaf(v.v[i], sc[i], sh[i]);
      if (res) o += r.v[i];
      v.v[i] = act_apply(o, act);
    }
    v.store(y + e * VEC);
  }
}

template <typename T>
__global__ void __launch_bounds__(256, 4) bn_bwd_apply_kernel(
    const T* __restrict__ dy, const T* __restrict__ x, const T* __restrict__ y, const float* __restrict__ gamma,
    const float* __restrict__ beta, const float* __restrict__ mean, const float* __restrict__ invstd,
    const double* __restrict__ acc, T* __restrict__ dx, T* __restrict__ dres, int64_t rows, int C, int act,
    int training, int64_t stride_vecs) {
  pdl_trigger();
  pdl_wait();
  constexpr int VEC = Elem<T>::kVec;
  const int cvn = C / VEC;
  const int64_t total = rows * cvn;
  int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (e >= stride_vecs) return;
  const int c0 = (int)(e % cvn) * VEC;
  // dx = k (dz - m1 - xhat m2) with xhat = (x - mu) is  ==  k dz + bx x + cx : three per-channel constants (+ the
  // forward's shift for the recomputed mask) instead of five - ncu showed this kernel at 33 % warp occupancy, limited
  // by its 80 registers; with 64 a fourth block is resident per SM
  float k[VEC], bx[VEC], cx[VEC], sh[VEC];
  const float inv_n = 1.0f / (float)rows;
  const bool from_y = act != CVX_ACT_NONE && beta == nullptr;   // else the mask is recomputed from x (forward's fma)
#pragma unroll
  for (int i = 0; i < VEC; ++i) {
    const float mu = mean[c0 + i], is = invstd[c0 + i];
    k[i] = gamma[c0 + i] * is;
    sh[i] = beta ? beta[c0 + i] - mu * k[i] : 0.f;
    const float m1 = training ? (float)(acc[c0 + i] * (double)inv_n) : 0.f;
    const float m2 = training ? (float)(acc[C + c0 + i] * (double)inv_n) : 0.f;
    bx[i] = -k[i] * m2 * is;
    cx[i] = -k[i] * m1 - bx[i] * mu;
  }
  for (; e < total; e += stride_vecs) {
    Vec<T> g, xv, yv;
    g.load(dy + e * VEC);
    xv.load(x + e * VEC);
    if (from_y) yv.load(y + e * VEC);
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
      float dz = g.v[i];
      if (act != CVX_ACT_NONE) dz *= act_mask(from_y ? yv.v[i] : fmaf(xv.v[i], k[i], sh[i]), act);
      g.v[i] = dz;
      xv.v[i] = fmaf(k[i], dz, fmaf(bx[i], xv.v[i], cx[i]));
    }
    xv.store(dx + e * VEC);
    if (dres) g.store(dres + e * VEC);
  }
}


__global__ void bn_param_grad_kernel(const double* __restrict__ acc, float* dgamma, float* dbeta, int C) {
  pdl_trigger();
  pdl_wait();
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  if (dbeta) dbeta[c] = (float)acc[c];
  if (dgamma) dgamma[c] = (float)acc[C + c];
}

__global__ void copy_d2f_kernel(const double* __restrict__ a, float* out, int n) {
  pdl_trigger();
  pdl_wait();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = (float)a[i];
}

// scalar fallback for narrow tensors whose channel count is not vector-aligned (C_out = 5 logits)
template <typename T>
__global__ void __launch_bounds__(256) bias_grad_small_kernel(const T* __restrict__ dy, double* __restrict__ out,
                                                              int64_t rows, int C) {
  float acc[16];
#pragma unroll
  for (int c = 0; c < 16; ++c) acc[c] = 0.f;
  for (int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; r < rows; r += (int64_t)gridDim.x * blockDim.x) {
#pragma unroll
    for (int c = 0; c < 16; ++c)
      if (c < C) acc[c] += Elem<T>::ld(dy + r * C + c);
  }
#pragma unroll
  for (int c = 0; c < 16; ++c) {
    if (c < C) {
      const float s = warp_sum(acc[c]);
      if ((threadIdx.x & 31) == 0) atomicAdd(out + c, (double)s);
    }
  }
}

// grid sizing for the "each thread owns a channel vector" streaming kernels
static inline void stream_grid(int64_t total_vecs, int cvn, int* blocks, int64_t* stride) {
  int64_t want = (int64_t)kNumSMs * 8 * 256;  // threads
  if (want > total_vecs) want = total_vecs;
  int64_t s = ceil_div64(want, cvn) * cvn;    // multiple of cvn
  *stride = s;
  *blocks = (int)ceil_div64(s, 256);
}

}  // namespace cvx

using namespace cvx;

extern "C" {

int cvx_bn_forward(const void* x, const void* residual, void* y, const float* gamma, const float* beta,
                   float* running_mean, float* running_var, float* save_mean, float* save_invstd, double* ws,
                   int64_t rows, int c, int dtype, int act, int training, float momentum, float eps,
                   void* stream) {
  CVX_CHECK_ARG(x && y && gamma && beta && save_mean && save_invstd && rows > 0 && c > 0, "bn_forward: bad arguments");
  CVX_CHECK_ARG(training || (running_mean && running_var), "bn_forward: eval mode needs running statistics");
  CVX_CHECK_ARG(!training || ws, "bn_forward: training mode needs the fp64 workspace");
  cudaStream_t st = as_stream(stream);
  const int vec = dtype == CVX_F32 ? 4 : 8;
  CVX_CHECK_ARG(c % vec == 0, "bn_forward: C=%d not a multiple of %d", c, vec);
  if (training) {
    CVX_WS_ZERO(ws, sizeof(double) * 2 * c, st);
    int rc = CVX_OK;
    CVX_DISPATCH_DTYPE(dtype, T, rc = (colreduce_launch<T, StatsF<T>>(StatsF<T>{(const T*)x, c}, rows, c, ws, st)));
    if (rc) return rc;
  }
  launch_pdl(bn_finalize_kernel, dim3((c + 127) / 128), dim3(128), 0, st, ws, rows, running_mean, running_var, save_mean, save_invstd,
                                                      c, training, momentum, eps);
  CVX_LAUNCH_OK();
  int blocks; int64_t stride;
  stream_grid(rows * (c / vec), c / vec, &blocks, &stride);
  CVX_DISPATCH_DTYPE(dtype, T, (launch_pdl(bn_apply_kernel<T>, dim3(blocks), dim3(256), 0, st, (const T*)x, (const T*)residual, (T*)y, gamma,
                                                                          beta, save_mean, save_invstd, rows, c, act,
                                                                          stride)));
  CVX_LAUNCH_OK();
  return CVX_OK;
}

int cvx_bn_backward(const void* dy, const void* x, const void* y, const float* gamma, const float* beta,
                    const float* save_mean, const float* save_invstd, void* dx, void* dres, float* dgamma,
                    float* dbeta, double* ws, int64_t rows, int c, int dtype, int act, int training, void* stream) {
  CVX_CHECK_ARG(dy && x && gamma && save_mean && save_invstd && dx && ws && rows > 0 && c > 0,
                "bn_backward: bad arguments");
  CVX_CHECK_ARG(act == CVX_ACT_NONE || y || beta, "bn_backward: the activation mask needs y, or beta to recompute it from x");
  CVX_CHECK_ARG(!(beta && dres), "bn_backward: the mask can only be recomputed from x when the forward had no residual");
  if (act == CVX_ACT_NONE) beta = nullptr;
  cudaStream_t st = as_stream(stream);
  const int vec = dtype == CVX_F32 ? 4 : 8;
  CVX_CHECK_ARG(c % vec == 0, "bn_backward: C=%d not a multiple of %d", c, vec);
  CVX_WS_ZERO(ws, sizeof(double) * 2 * c, st);
  int rc = CVX_OK;
  CVX_DISPATCH_DTYPE(dtype, T, rc = (colreduce_launch<T, BnBwdF<T>, 256, 3>(
                                   BnBwdF<T>{(const T*)dy, (const T*)x, (const T*)y, gamma, beta, save_mean, save_invstd, c, act},
                                   rows, c, ws, st)));
  if (rc) return rc;
  int blocks; int64_t stride;
  stream_grid(rows * (c / vec), c / vec, &blocks, &stride);
  CVX_DISPATCH_DTYPE(dtype, T, (launch_pdl(bn_bwd_apply_kernel<T>, dim3(blocks), dim3(256), 0, st, 
                                   (const T*)dy, (const T*)x, (const T*)y, gamma, beta, save_mean, save_invstd, ws, (T*)dx,
                                   (T*)dres, rows, c, act, training, stride)));
  CVX_LAUNCH_OK();
  if (dgamma || dbeta) {
    launch_pdl(bn_param_grad_kernel, dim3((c + 127) / 128), dim3(128), 0, st, ws, dgamma, dbeta, c);
    CVX_LAUNCH_OK();
  }
  return CVX_OK;
}

int cvx_bias_grad(const void* dy, float* dbias, double* ws, int64_t rows, int c, int dtype, void* stream) {
  CVX_CHECK_ARG(dy && dbias && ws && rows > 0 && c > 0, "bias_grad: bad arguments");
  cudaStream_t st = as_stream(stream);
  CVX_WS_ZERO(ws, sizeof(double) * c, st);
  int rc = CVX_OK;
  const int vec = dtype == CVX_F32 ? 4 : 8;
  if (c % vec != 0) {
    CVX_CHECK_ARG(c <= 16, "bias_grad: unaligned C=%d above the scalar path's limit of 16", c);
    int blocks = (int)(ceil_div64(rows, 256 * 8) > kNumSMs * 4 ? kNumSMs * 4 : ceil_div64(rows, 256 * 8));
    CVX_DISPATCH_DTYPE(dtype, T, (bias_grad_small_kernel<T><<<blocks, 256, 0, st>>>((const T*)dy, ws, rows, c)));
    CVX_LAUNCH_OK();
  } else {
    CVX_DISPATCH_DTYPE(dtype, T, rc = (colreduce_launch<T, SumF<T>>(SumF<T>{(const T*)dy, c}, rows, c, ws, st)));
    if (rc) return rc;
  }
  launch_pdl(copy_d2f_kernel, dim3((c + 127) / 128), dim3(128), 0, st, ws, dbias, c);
  CVX_LAUNCH_OK();
  return CVX_OK;
}

}  // extern "C"
