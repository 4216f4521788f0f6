// Depthwise 3x3 convolution on NHWC (stride 1/2, any dilation), forward / data-grad / weight-grad.
//
// Reference: nn.Conv2d(groups=C) at xception.py:13 and mobilenetv2.py:39,58.  Pure bandwidth
// work (9 MACs per element): every thread owns one 16-byte channel vector of one output
// pixel; a warp covers >= 512 contiguous bytes of each tap row.  relu_in fuses the
// SeparableConv2d.relu0 pre-activation (xception.py:22-23) into the loads.
#include "colreduce.cuh"

namespace cvx {

struct DwGeom {
  int n, h, w, c, stride, pad, dil, ho, wo;
};

template <typename T>
__global__ void __launch_bounds__(256) dw_fwd_kernel(const T* __restrict__ x, const float* __restrict__ w9c,
                                                     T* __restrict__ y, DwGeom g, int relu_in) {
  constexpr int VEC = Elem<T>::kVec;
  const int cvn = g.c / VEC;
  const int64_t total = (int64_t)g.n * g.ho * g.wo * cvn;
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < total;
       e += (int64_t)gridDim.x * blockDim.x) {
    const int c0 = (int)(e % cvn) * VEC;
    int64_t p = e / cvn;
    const int ox = (int)(p % g.wo);
    const int oy = (int)((p / g.wo) % g.ho);
    const int nn = (int)(p / ((int64_t)g.wo * g.ho));
    float acc[VEC];
#pragma unroll
    for (int i = 0; i < VEC; ++i) acc[i] = 0.f;
#pragma unroll
    for (int kh = 0; kh < 3; ++kh) {
      const int iy = oy * g.stride - g.pad + kh * g.dil;
      if (iy < 0 || iy >= g.h) continue;
#pragma unroll
      for (int kw = 0; kw < 3; ++kw) {
        const int ix = ox * g.stride - g.pad + kw * g.dil;
        if (ix < 0 || ix >= g.w) continue;
        Vec<T> v;
        v.load(x + (((int64_t)nn * g.h + iy) * g.w + ix) * g.c + c0);
        const float* wp = w9c + (kh * 3 + kw) * g.c + c0;
#pragma unroll
        for (int i = 0; i < VEC; ++i) {
          const float xv = relu_in ? fmaxf(v.v[i], 0.f) : v.v[i];
          acc[i] = fmaf(xv, __ldg(wp + i), acc[i]);
        }
      }
    }
    Vec<T> o;
#pragma unroll
    for (int i = 0; i < VEC; ++i) o.v[i] = acc[i];
    o.store(y + e * VEC);
  }
}

template <typename T>
__global__ void __launch_bounds__(256) dw_bwd_data_kernel(const T* __restrict__ dy, const float* __restrict__ w9c,
                                                          const T* __restrict__ x, T* __restrict__ dx, DwGeom g,
                                                          int relu_in) {
  constexpr int VEC = Elem<T>::kVec;
  const int cvn = g.c / VEC;
  const int64_t total = (int64_t)g.n * g.h * g.w * cvn;
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < total;
       e += (int64_t)gridDim.x * blockDim.x) {
    const int c0 = (int)(e % cvn) * VEC;
    int64_t p = e / cvn;
    const int ix = (int)(p % g.w);
    const int iy = (int)((p / g.w) % g.h);
    const int nn = (int)(p / ((int64_t)g.w * g.h));
    float acc[VEC];
#pragma unroll
    for (int i = 0; i < VEC; ++i) acc[i] = 0.f;
#pragma unroll
    for (int kh = 0; kh < 3; ++kh) {
      const int ny = iy + g.pad - kh * g.dil;
      if (ny < 0 || ny % g.stride != 0) continue;
      const int oy = ny / g.stride;
      if (oy >= g.ho) continue;
#pragma unroll
      for (int kw = 0; kw < 3; ++kw) {
        const int nx = ix + g.pad - kw * g.dil;
        if (nx < 0 || nx % g.stride != 0) continue;
        const int ox = nx / g.stride;
        if (ox >= g.wo) continue;
        Vec<T> v;
        v.load(dy + (((int64_t)nn * g.ho + oy) * g.wo + ox) * g.c + c0);
        const float* wp = w9c + (kh * 3 + kw) * g.c + c0;
#pragma unroll
        for (int i = 0; i < VEC; ++i) acc[i] = fmaf(v.v[i], __ldg(wp + i), acc[i]);
      }
    }
    Vec<T> o;
    if (relu_in) {
      Vec<T> xv;
      xv.load(x + e * VEC);
#pragma unroll
      for (int i = 0; i < VEC; ++i) o.v[i] = xv.v[i] > 0.f ? acc[i] : 0.f;
    } else {
#pragma unroll
      for (int i = 0; i < VEC; ++i) o.v[i] = acc[i];
    }
    o.store(dx + e * VEC);
  }
}

template <typename T>
struct DwWgradF {
  static constexpr int NACC = 9;
  const T* x;
  const T* dy;
  DwGeom g;
  int relu_in;
  __device__ __forceinline__ void operator()(int64_t row, int c0, float (&acc)[9][Elem<T>::kVec]) const {
    const int ox = (int)(row % g.wo);
    const int oy = (int)((row / g.wo) % g.ho);
    const int nn = (int)(row / ((int64_t)g.wo * g.ho));
    Vec<T> gv;
    gv.load(dy + row * g.c + c0);
#pragma unroll
    for (int kh = 0; kh < 3; ++kh) {
      const int iy = oy * g.stride - g.pad + kh * g.dil;
      if (iy < 0 || iy >= g.h) continue;
#pragma unroll
      for (int kw = 0; kw < 3; ++kw) {
        const int ix = ox * g.stride - g.pad + kw * g.dil;
        if (ix < 0 || ix >= g.w) continue;
        Vec<T> v;
        v.load(x + (((int64_t)nn * g.h + iy) * g.w + ix) * g.c + c0);
#pragma unroll
        for (int i = 0; i < Vec<T>::N; ++i) {
          const float xv = relu_in ? fmaxf(v.v[i], 0.f) : v.v[i];
          acc[kh * 3 + kw][i] = fmaf(gv.v[i], xv, acc[kh * 3 + kw][i]);
        }
      }
    }
  }
};

__global__ void dw_copy_d2f_kernel(const double* __restrict__ a, float* out, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = (float)a[i];
}

static int dw_check(const cvx_conv_desc* d, const char* who, DwGeom* g) {
  CVX_CHECK_ARG(d != nullptr, "%s: null descriptor", who);
  CVX_CHECK_ARG(d->kh == 3 && d->kw == 3 && d->cin == d->cout, "%s: only depthwise 3x3 is supported", who);
  CVX_CHECK_ARG(d->n > 0 && d->h > 0 && d->w > 0 && d->cin > 0 && d->stride > 0 && d->dil > 0 && d->pad >= 0,
                "%s: bad geometry", who);
  const int ho = (d->h + 2 * d->pad - d->dil * 2 - 1) / d->stride + 1;
  const int wo = (d->w + 2 * d->pad - d->dil * 2 - 1) / d->stride + 1;
  CVX_CHECK_ARG(ho == d->ho && wo == d->wo, "%s: inconsistent output size", who);
  const int vec = d->dtype == CVX_F32 ? 4 : 8;
  CVX_CHECK_ARG(d->cin % vec == 0, "%s: C=%d not a multiple of %d", who, d->cin, vec);
  *g = DwGeom{d->n, d->h, d->w, d->cin, d->stride, d->pad, d->dil, d->ho, d->wo};
  return CVX_OK;
}

static inline int dw_grid(int64_t total) {
  int64_t b = ceil_div64(total, 256);
  const int64_t cap = (int64_t)kNumSMs * 16;
  return (int)(b > cap ? cap : (b < 1 ? 1 : b));
}

}  // namespace cvx

using namespace cvx;

extern "C" {

int cvx_dwconv_fwd(const cvx_conv_desc* d, const void* x, const float* w9c, void* y, int relu_in, void* stream) {
  DwGeom g;
  if (int rc = dw_check(d, "dwconv_fwd", &g)) return rc;
  CVX_CHECK_ARG(x && w9c && y, "dwconv_fwd: null pointer");
  const int vec = d->dtype == CVX_F32 ? 4 : 8;
  const int64_t total = (int64_t)g.n * g.ho * g.wo * (g.c / vec);
  CVX_DISPATCH_DTYPE(d->dtype, T, (dw_fwd_kernel<T><<<dw_grid(total), 256, 0, as_stream(stream)>>>(
                                      (const T*)x, w9c, (T*)y, g, relu_in)));
  CVX_LAUNCH_OK();
  return CVX_OK;
}

int cvx_dwconv_bwd_data(const cvx_conv_desc* d, const void* dy, const float* w9c, const void* x, void* dx,
                        int relu_in, void* stream) {
  DwGeom g;
  if (int rc = dw_check(d, "dwconv_bwd_data", &g)) return rc;
  CVX_CHECK_ARG(dy && w9c && dx && (!relu_in || x), "dwconv_bwd_data: null pointer");
  const int vec = d->dtype == CVX_F32 ? 4 : 8;
  const int64_t total = (int64_t)g.n * g.h * g.w * (g.c / vec);
  CVX_DISPATCH_DTYPE(d->dtype, T, (dw_bwd_data_kernel<T><<<dw_grid(total), 256, 0, as_stream(stream)>>>(
                                      (const T*)dy, w9c, (const T*)x, (T*)dx, g, relu_in)));
  CVX_LAUNCH_OK();
  return CVX_OK;
}

int cvx_dwconv_bwd_weight(const cvx_conv_desc* d, const void* x, const void* dy, float* dw9c, double* ws,
                          int relu_in, void* stream) {
  DwGeom g;
  if (int rc = dw_check(d, "dwconv_bwd_weight", &g)) return rc;
  CVX_CHECK_ARG(x && dy && dw9c && ws, "dwconv_bwd_weight: null pointer");
  cudaStream_t st = as_stream(stream);
  CVX_CUDA_OK(cudaMemsetAsync(ws, 0, sizeof(double) * 9 * g.c, st));
  const int64_t rows = (int64_t)g.n * g.ho * g.wo;
  int rc = CVX_OK;
  CVX_DISPATCH_DTYPE(d->dtype, T, rc = (colreduce_launch<T, DwWgradF<T>>(
                                      DwWgradF<T>{(const T*)x, (const T*)dy, g, relu_in}, rows, g.c, ws, st)));
  if (rc) return rc;
  dw_copy_d2f_kernel<<<(9 * g.c + 255) / 256, 256, 0, st>>>(ws, dw9c, 9 * g.c);
  CVX_LAUNCH_OK();
  return CVX_OK;
}

}  // extern "C"
