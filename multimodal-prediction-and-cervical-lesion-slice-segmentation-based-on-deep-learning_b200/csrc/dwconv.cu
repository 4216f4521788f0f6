// Depthwise 3x3 convolution on NHWC (stride 1/2, any dilation), forward / data-grad / weight-grad.
//
// Reference: nn.Conv2d(groups=C) at xception.py:13 and mobilenetv2.py:39,58.  Pure bandwidth
// work (9 MACs per element).  Every thread owns ONE 16-byte channel vector for its whole
// life: the 9xVEC filter taps live in registers, and the thread strides over pixels so that a
// warp always touches >= 512 contiguous bytes per tap row; the nine tap loads of a pixel are
// issued back to back (raw 16-byte loads, converted while accumulating).  relu_in fuses the
// SeparableConv2d.relu0 pre-activation (xception.py:22-23) into the loads.
#include "colreduce.cuh"
#include "dwconv_tiled.cuh"

namespace cvx {

struct DwGeom {
  int n, h, w, c, stride, pad, dil, ho, wo;
};

template <typename T> struct RawVec { uint4 r; };

template <typename T>
__device__ __forceinline__ void raw_to_float(const uint4& r, float (&v)[Elem<T>::kVec]);
template <>
__device__ __forceinline__ void raw_to_float<float>(const uint4& r, float (&v)[4]) {
  v[0] = __uint_as_float(r.x); v[1] = __uint_as_float(r.y); v[2] = __uint_as_float(r.z); v[3] = __uint_as_float(r.w);
}
template <>
__device__ __forceinline__ void raw_to_float<__nv_bfloat16>(const uint4& r, float (&v)[8]) {
  const uint32_t u[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    v[2 * i] = __uint_as_float(u[i] << 16);
    v[2 * i + 1] = __uint_as_float(u[i] & 0xffff0000u);
  }
}

// MODE 0: forward (rows = output pixels, gather from x)      MODE 1: data gradient (rows = input pixels,
// gather from dy through the transposed stencil, optional relu mask from x)
template <typename T, int MODE>
__global__ void __launch_bounds__(128, 4) dw_stream_kernel(const T* __restrict__ src, const float* __restrict__ w9c,
                                                           const T* __restrict__ xmask, T* __restrict__ dst, DwGeom g,
                                                           int relu_in, int64_t stride_items) {
  constexpr int VEC = Elem<T>::kVec;
  const int cvn = g.c / VEC;
  const int64_t e0 = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (e0 >= stride_items) return;
  const int c0 = (int)(e0 % cvn) * VEC;
  float wreg[9][VEC];
#pragma unroll
  for (int t = 0; t < 9; ++t)
#pragma unroll
    for (int i = 0; i < VEC; ++i) wreg[t][i] = __ldg(w9c + t * g.c + c0 + i);

  const int rows_h = MODE == 0 ? g.ho : g.h, rows_w = MODE == 0 ? g.wo : g.w;
  const int src_h = MODE == 0 ? g.h : g.ho, src_w = MODE == 0 ? g.w : g.wo;
  const int sshift = g.stride == 2 ? 1 : 0, smask = g.stride - 1;  // stride is 1 or 2 (checked on the host)
  const int npix = g.n * rows_h * rows_w;  // < 2^31 (checked by the launcher)
  const int pstep = (int)(stride_items / cvn);
  for (int p = (int)(e0 / cvn); p < npix; p += pstep) {
    const int px = p % rows_w;
    const int t1 = p / rows_w;
    const int py = t1 % rows_h;
    const int nn = t1 / rows_h;
    uint4 raw[9];
    bool ok[9];
#pragma unroll
    for (int kh = 0; kh < 3; ++kh) {
      int sy;
      bool oky;
      if (MODE == 0) {
        sy = py * g.stride - g.pad + kh * g.dil;
        oky = sy >= 0 && sy < src_h;
      } else {
        const int ny = py + g.pad - kh * g.dil;
        oky = ny >= 0 && ((ny & smask) == 0);
        sy = ny >> sshift;
        oky = oky && sy < src_h;
      }
#pragma unroll
      for (int kw = 0; kw < 3; ++kw) {
        int sx;
        bool okx;
        if (MODE == 0) {
          sx = px * g.stride - g.pad + kw * g.dil;
          okx = sx >= 0 && sx < src_w;
        } else {
          const int nx = px + g.pad - kw * g.dil;
          okx = nx >= 0 && ((nx & smask) == 0);
          sx = nx >> sshift;
          okx = okx && sx < src_w;
        }
        const int t = kh * 3 + kw;
        ok[t] = oky && okx;
        raw[t] = ok[t] ? __ldg(reinterpret_cast<const uint4*>(src + (((int64_t)nn * src_h + sy) * src_w + sx) * g.c + c0))
                       : make_uint4(0, 0, 0, 0);
      }
    }
    float acc[VEC];
#pragma unroll
    for (int i = 0; i < VEC; ++i) acc[i] = 0.f;
#pragma unroll
    for (int t = 0; t < 9; ++t) {
      float v[VEC];
      raw_to_float<T>(raw[t], v);
#pragma unroll
      for (int i = 0; i < VEC; ++i) {
        const float xv = (MODE == 0 && relu_in) ? fmaxf(v[i], 0.f) : v[i];
        acc[i] = fmaf(xv, wreg[t][i], acc[i]);
      }
    }
    Vec<T> o;
    if (MODE == 1 && relu_in) {
      Vec<T> xv;
      xv.load(xmask + (int64_t)p * g.c + c0);
#pragma unroll
      for (int i = 0; i < VEC; ++i) o.v[i] = xv.v[i] > 0.f ? acc[i] : 0.f;
    } else {
#pragma unroll
      for (int i = 0; i < VEC; ++i) o.v[i] = acc[i];
    }
    o.store(dst + (int64_t)p * g.c + c0);
  }
}

// ---- stride 2, pad 1, dilation 1 data gradient (the three entry-flow down-sampling separable convs) ----------
// A thread owns a channel vector and a 2x2 quad of INPUT pixels (rows 2a, 2a+1; columns 2b, 2b+1).  With stride 2
// only taps of matching parity reach a pixel, so the quad needs exactly the 2x2 window dy[a..a+1][b..b+1] and 9
// multiply-adds per channel (the generic stencil walks all 9 taps of every pixel and masks 3/4 of them off):
//   dx[2a  ][2b  ] = w11 dy[a][b]                    dx[2a  ][2b+1] = w10 dy[a][b+1] + w12 dy[a][b]
//   dx[2a+1][2b  ] = w01 dy[a+1][b] + w21 dy[a][b]   dx[2a+1][2b+1] = w00 dy[a+1][b+1] + w02 dy[a+1][b]
//                                                                   + w20 dy[a][b+1]  + w22 dy[a][b]
// All eight loads (4 of dy, 4 of the relu mask source) are issued before the first use.
template <typename T>
__global__ void __launch_bounds__(128, 3) dw_s2_dgrad_kernel(const T* __restrict__ dy, const float* __restrict__ w9c,
                                                             const T* __restrict__ xmask, T* __restrict__ dx, DwGeom g,
                                                             int relu_in, int64_t stride_items) {
  constexpr int VEC = Elem<T>::kVec;
  const int cvn = g.c / VEC;
  const int64_t e0 = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (e0 >= stride_items) return;
  const int c0 = (int)(e0 % cvn) * VEC;
  float wreg[9][VEC];
#pragma unroll
  for (int t = 0; t < 9; ++t)
#pragma unroll
    for (int i = 0; i < VEC; ++i) wreg[t][i] = __ldg(w9c + t * g.c + c0 + i);
  const int qh = (g.h + 1) >> 1, qw = (g.w + 1) >> 1;
  const int nquads = g.n * qh * qw;
  const int qstep = (int)(stride_items / cvn);
  for (int q = (int)(e0 / cvn); q < nquads; q += qstep) {
    const int b = q % qw;
    const int t1 = q / qw;
    const int a = t1 % qh;
    const int nn = t1 / qh;
    const bool a1 = a + 1 < g.ho, b1 = b + 1 < g.wo;          // (a < ho and b < wo always: ho = ceil(h/2))
    const bool r1 = 2 * a + 1 < g.h, c1 = 2 * b + 1 < g.w;    // second row / column of the quad inside the image
    const T* dyp = dy + (((int64_t)nn * g.ho + a) * g.wo + b) * g.c + c0;
    const int64_t xoff = (((int64_t)nn * g.h + 2 * a) * g.w + 2 * b) * g.c + c0;
    const int64_t xrow = (int64_t)g.w * g.c;
    const uint4 z = make_uint4(0, 0, 0, 0);
    const uint4 d00 = __ldg(reinterpret_cast<const uint4*>(dyp));
    const uint4 d01 = b1 ? __ldg(reinterpret_cast<const uint4*>(dyp + g.c)) : z;
    const uint4 d10 = a1 ? __ldg(reinterpret_cast<const uint4*>(dyp + (int64_t)g.wo * g.c)) : z;
    const uint4 d11 = (a1 && b1) ? __ldg(reinterpret_cast<const uint4*>(dyp + (int64_t)g.wo * g.c + g.c)) : z;
    uint4 m00 = z, m01 = z, m10 = z, m11 = z;
    if (relu_in) {
      m00 = __ldg(reinterpret_cast<const uint4*>(xmask + xoff));
      if (c1) m01 = __ldg(reinterpret_cast<const uint4*>(xmask + xoff + g.c));
      if (r1) m10 = __ldg(reinterpret_cast<const uint4*>(xmask + xoff + xrow));
      if (r1 && c1) m11 = __ldg(reinterpret_cast<const uint4*>(xmask + xoff + xrow + g.c));
    }
    float v00[VEC], v01[VEC], v10[VEC], v11[VEC];
    raw_to_float<T>(d00, v00); raw_to_float<T>(d01, v01); raw_to_float<T>(d10, v10); raw_to_float<T>(d11, v11);
    Vec<T> o00, o01, o10, o11;
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
      o00.v[i] = wreg[4][i] * v00[i];
      o01.v[i] = fmaf(wreg[3][i], v01[i], wreg[5][i] * v00[i]);
      o10.v[i] = fmaf(wreg[1][i], v10[i], wreg[7][i] * v00[i]);
      o11.v[i] = fmaf(wreg[0][i], v11[i], fmaf(wreg[2][i], v10[i], fmaf(wreg[6][i], v01[i], wreg[8][i] * v00[i])));
    }
    if (relu_in) {
      float k00[VEC], k01[VEC], k10[VEC], k11[VEC];
      raw_to_float<T>(m00, k00); raw_to_float<T>(m01, k01); raw_to_float<T>(m10, k10); raw_to_float<T>(m11, k11);
#pragma unroll
      for (int i = 0; i < VEC; ++i) {
        o00.v[i] = k00[i] > 0.f ? o00.v[i] : 0.f;
        o01.v[i] = k01[i] > 0.f ? o01.v[i] : 0.f;
        o10.v[i] = k10[i] > 0.f ? o10.v[i] : 0.f;
        o11.v[i] = k11[i] > 0.f ? o11.v[i] : 0.f;
      }
    }
    o00.store(dx + xoff);
    if (c1) o01.store(dx + xoff + g.c);
    if (r1) o10.store(dx + xoff + xrow);
    if (r1 && c1) o11.store(dx + xoff + xrow + g.c);
  }
}

// ---- stride 1, dilation 1 (the 60 middle/exit-flow depthwise convs of Xception) -------------
// A thread owns a channel vector AND walks a strip of SX consecutive output columns with a
// sliding 3x3 window held in registers: 3 new 16-byte loads per output instead of 9, and the
// pixel decode / bounds logic amortised over the strip.  Consecutive strips are vertically
// adjacent rows, so neighbouring threads re-use each other's rows out of L1.
//   MODE 0 forward, MODE 1 data gradient (flipped taps, optional relu mask from x),
//   MODE 2 weight gradient (src = x, second = dy; 9xVEC accumulators reduced per block).
constexpr int kDwSX = 8;

template <typename T, int MODE>
__global__ void __launch_bounds__(128, 2) dw_s1_kernel(const T* __restrict__ src, const float* __restrict__ w9c,
                                                       const T* __restrict__ second, T* __restrict__ dst,
                                                       double* __restrict__ wgrad_out, int n, int h, int w, int c,
                                                       int relu_in, int64_t stride_items) {
  constexpr int VEC = Elem<T>::kVec;
  extern __shared__ float sm_acc[];  // MODE 2 only: [9][chunk*VEC]
  const int cvn = c / VEC;
  const int64_t e0 = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  const bool active = e0 < stride_items;
  const int cv = (int)(e0 % cvn), c0 = cv * VEC;
  float wreg[9][VEC];
#pragma unroll
  for (int t = 0; t < 9; ++t)
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
      if (MODE == 2) wreg[t][i] = 0.f;                                   // accumulators
      else wreg[t][i] = __ldg(w9c + (MODE == 1 ? 8 - t : t) * c + c0 + i);  // taps (flipped for dgrad)
    }
  const int strips_x = (w + kDwSX - 1) / kDwSX;
  const int nstrips = n * h * strips_x;
  const int sstep = (int)(stride_items / cvn);
  const bool relu_src = (MODE != 1) && relu_in;
  if (active) {
    for (int s = (int)(e0 / cvn); s < nstrips; s += sstep) {
      const int y = s % h;
      const int t1 = s / h;
      const int x0 = (t1 % strips_x) * kDwSX;
      const int nn = t1 / strips_x;
      const T* rowp[3];
      bool rok[3];
#pragma unroll
      for (int r = 0; r < 3; ++r) {
        const int yy = y - 1 + r;
        rok[r] = yy >= 0 && yy < h;
        rowp[r] = src + ((int64_t)(nn * h + (rok[r] ? yy : y)) * w) * c + c0;
      }
      float win[3][3][VEC];  // [column slot][row][lane element]
      auto load_col = [&](int slot, int xx) {
        const bool cok = xx >= 0 && xx < w;
#pragma unroll
        for (int r = 0; r < 3; ++r) {
          uint4 raw = (cok && rok[r]) ? __ldg(reinterpret_cast<const uint4*>(rowp[r] + (int64_t)xx * c))
                                      : make_uint4(0, 0, 0, 0);
          raw_to_float<T>(raw, win[slot][r]);
          if (relu_src) {
#pragma unroll
            for (int i = 0; i < VEC; ++i) win[slot][r][i] = fmaxf(win[slot][r][i], 0.f);
          }
        }
      };
      load_col(0, x0 - 1);
      load_col(1, x0);
#pragma unroll
      for (int j = 0; j < kDwSX; ++j) {
        const int xx = x0 + j;
        load_col((j + 2) % 3, xx + 1);
        if (xx < w) {
          const int64_t off = ((int64_t)(nn * h + y) * w + xx) * c + c0;
          if (MODE == 2) {
            float gv[VEC];
            raw_to_float<T>(__ldg(reinterpret_cast<const uint4*>(second + off)), gv);
#pragma unroll
            for (int kw = 0; kw < 3; ++kw)
#pragma unroll
              for (int r = 0; r < 3; ++r)
#pragma unroll
                for (int i = 0; i < VEC; ++i)
                  wreg[r * 3 + kw][i] = fmaf(gv[i], win[(j + kw) % 3][r][i], wreg[r * 3 + kw][i]);
          } else {
            float acc[VEC];
#pragma unroll
            for (int i = 0; i < VEC; ++i) acc[i] = 0.f;
#pragma unroll
            for (int kw = 0; kw < 3; ++kw)
#pragma unroll
              for (int r = 0; r < 3; ++r)
#pragma unroll
                for (int i = 0; i < VEC; ++i) acc[i] = fmaf(win[(j + kw) % 3][r][i], wreg[r * 3 + kw][i], acc[i]);
            Vec<T> o;
            if (MODE == 1 && relu_in) {
              Vec<T> xv;
              xv.load(second + off);
#pragma unroll
              for (int i = 0; i < VEC; ++i) o.v[i] = xv.v[i] > 0.f ? acc[i] : 0.f;
            } else {
#pragma unroll
              for (int i = 0; i < VEC; ++i) o.v[i] = acc[i];
            }
            o.store(dst + off);
          }
        }
      }
    }
  }
  if (MODE == 2) {
    // block-level combine: threads with the same (threadIdx % cvn_in_block) share a channel vector
    // only when blockDim is a multiple of cvn; in general use shared-memory atomics keyed by channel.
    const int chunk = cvn < (int)blockDim.x ? cvn : (int)blockDim.x;
    (void)chunk;
    for (int i = threadIdx.x; i < 9 * c; i += blockDim.x) sm_acc[i] = 0.f;
    __syncthreads();
    if (active) {
#pragma unroll
      for (int t = 0; t < 9; ++t)
#pragma unroll
        for (int i = 0; i < VEC; ++i) atomicAdd(&sm_acc[t * c + c0 + i], wreg[t][i]);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 9 * c; i += blockDim.x)
      if (sm_acc[i] != 0.f) atomicAdd(wgrad_out + i, (double)sm_acc[i]);
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
    constexpr int VEC = Elem<T>::kVec;
    const int ox = (int)(row % g.wo);
    const int oy = (int)((row / g.wo) % g.ho);
    const int nn = (int)(row / ((int64_t)g.wo * g.ho));
    const uint4 graw = __ldg(reinterpret_cast<const uint4*>(dy + row * g.c + c0));
    uint4 raw[9];
#pragma unroll
    for (int kh = 0; kh < 3; ++kh) {
      const int iy = oy * g.stride - g.pad + kh * g.dil;
      const bool oky = iy >= 0 && iy < g.h;
#pragma unroll
      for (int kw = 0; kw < 3; ++kw) {
        const int ix = ox * g.stride - g.pad + kw * g.dil;
        const bool ok = oky && ix >= 0 && ix < g.w;
        raw[kh * 3 + kw] = ok ? __ldg(reinterpret_cast<const uint4*>(x + (((int64_t)nn * g.h + iy) * g.w + ix) * g.c + c0))
                              : make_uint4(0, 0, 0, 0);
      }
    }
    float gv[VEC];
    raw_to_float<T>(graw, gv);
#pragma unroll
    for (int t = 0; t < 9; ++t) {
      float v[VEC];
      raw_to_float<T>(raw[t], v);
#pragma unroll
      for (int i = 0; i < VEC; ++i) {
        const float xv = relu_in ? fmaxf(v[i], 0.f) : v[i];
        acc[t][i] = fmaf(gv[i], xv, acc[t][i]);
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
  CVX_CHECK_ARG(d->stride == 1 || d->stride == 2, "%s: stride must be 1 or 2", who);
  const int vec = d->dtype == CVX_F32 ? 4 : 8;
  CVX_CHECK_ARG(d->cin % vec == 0, "%s: C=%d not a multiple of %d", who, d->cin, vec);
  *g = DwGeom{d->n, d->h, d->w, d->cin, d->stride, d->pad, d->dil, d->ho, d->wo};
  return CVX_OK;
}

// threads = a multiple of the channel-vector count, about 4 resident blocks of 128 per SM
static inline void dw_grid(int64_t total_items, int cvn, int* blocks, int64_t* stride) {
  int64_t want = (int64_t)kNumSMs * 4 * 128 * 2;
  if (want > total_items) want = total_items;
  const int64_t s = ceil_div64(want, cvn) * cvn;
  *stride = s;
  *blocks = (int)ceil_div64(s, 128);
}

static inline bool dw_is_s1(const DwGeom& g) {
  return g.stride == 1 && g.dil == 1 && g.pad == 1 && (int64_t)g.n * g.h * g.w < (1ll << 31);
}
// strips x channel-vectors, one wave of 3 resident 128-thread blocks per SM
static inline void dw_s1_grid(const DwGeom& g, int cvn, int* blocks, int64_t* stride) {
  const int64_t items = (int64_t)g.n * g.h * ((g.w + kDwSX - 1) / kDwSX) * cvn;
  int64_t want = (int64_t)kNumSMs * 2 * 128;
  if (want > items) want = items;
  const int64_t s = ceil_div64(want, cvn) * cvn;
  *stride = s;
  *blocks = (int)ceil_div64(s, 128);
}

}  // namespace cvx

using namespace cvx;

extern "C" {

int cvx_dwconv_fwd(const cvx_conv_desc* d, const void* x, const float* w9c, void* y, int relu_in, void* stream) {
  DwGeom g;
  if (int rc = dw_check(d, "dwconv_fwd", &g)) return rc;
  CVX_CHECK_ARG(x && w9c && y, "dwconv_fwd: null pointer");
  const int vec = d->dtype == CVX_F32 ? 4 : 8;
  int blocks; int64_t stride;
  CVX_CHECK_ARG((int64_t)g.n * g.h * g.w < (1ll << 31), "dwconv_fwd: tensor too large for 32-bit pixel indices");
  if (int rc = dw_tiled_launch(0, d, x, w9c, nullptr, y, nullptr, relu_in, as_stream(stream)); rc != CVX_EUNSUPPORTED) return rc;
  if (dw_is_s1(g)) {
    dw_s1_grid(g, g.c / vec, &blocks, &stride);
    CVX_DISPATCH_DTYPE(d->dtype, T, (dw_s1_kernel<T, 0><<<blocks, 128, 0, as_stream(stream)>>>(
                                        (const T*)x, w9c, nullptr, (T*)y, nullptr, g.n, g.h, g.w, g.c, relu_in, stride)));
    CVX_LAUNCH_OK();
    return CVX_OK;
  }
  dw_grid((int64_t)g.n * g.ho * g.wo * (g.c / vec), g.c / vec, &blocks, &stride);
  CVX_DISPATCH_DTYPE(d->dtype, T, (dw_stream_kernel<T, 0><<<blocks, 128, 0, as_stream(stream)>>>(
                                      (const T*)x, w9c, nullptr, (T*)y, g, relu_in, stride)));
  CVX_LAUNCH_OK();
  return CVX_OK;
}

int cvx_dwconv_bwd_data(const cvx_conv_desc* d, const void* dy, const float* w9c, const void* x, void* dx,
                        int relu_in, void* stream) {
  DwGeom g;
  if (int rc = dw_check(d, "dwconv_bwd_data", &g)) return rc;
  CVX_CHECK_ARG(dy && w9c && dx && (!relu_in || x), "dwconv_bwd_data: null pointer");
  const int vec = d->dtype == CVX_F32 ? 4 : 8;
  int blocks; int64_t stride;
  CVX_CHECK_ARG((int64_t)g.n * g.h * g.w < (1ll << 31), "dwconv_bwd_data: tensor too large for 32-bit pixel indices");
  if (int rc = dw_tiled_launch(1, d, dy, w9c, x, dx, nullptr, relu_in, as_stream(stream)); rc != CVX_EUNSUPPORTED) return rc;
  if (dw_is_s1(g)) {
    dw_s1_grid(g, g.c / vec, &blocks, &stride);
    CVX_DISPATCH_DTYPE(d->dtype, T, (dw_s1_kernel<T, 1><<<blocks, 128, 0, as_stream(stream)>>>(
                                        (const T*)dy, w9c, (const T*)x, (T*)dx, nullptr, g.n, g.h, g.w, g.c, relu_in, stride)));
    CVX_LAUNCH_OK();
    return CVX_OK;
  }
  if (g.stride == 2 && g.pad == 1 && g.dil == 1 && g.ho == (g.h + 1) / 2 && g.wo == (g.w + 1) / 2) {
    const int64_t items = (int64_t)g.n * ((g.h + 1) / 2) * ((g.w + 1) / 2) * (g.c / vec);
    int64_t want = (int64_t)kNumSMs * 3 * 128 * 4;
    if (want > items) want = items;
    stride = ceil_div64(want, g.c / vec) * (g.c / vec);
    blocks = (int)ceil_div64(stride, 128);
    CVX_DISPATCH_DTYPE(d->dtype, T, (dw_s2_dgrad_kernel<T><<<blocks, 128, 0, as_stream(stream)>>>(
                                        (const T*)dy, w9c, (const T*)x, (T*)dx, g, relu_in, stride)));
    CVX_LAUNCH_OK();
    return CVX_OK;
  }
  dw_grid((int64_t)g.n * g.h * g.w * (g.c / vec), g.c / vec, &blocks, &stride);
  CVX_DISPATCH_DTYPE(d->dtype, T, (dw_stream_kernel<T, 1><<<blocks, 128, 0, as_stream(stream)>>>(
                                      (const T*)dy, w9c, (const T*)x, (T*)dx, g, relu_in, stride)));
  CVX_LAUNCH_OK();
  return CVX_OK;
}

int cvx_dwconv_bwd_weight(const cvx_conv_desc* d, const void* x, const void* dy, float* dw9c, double* ws,
                          int relu_in, void* stream) {
  DwGeom g;
  if (int rc = dw_check(d, "dwconv_bwd_weight", &g)) return rc;
  CVX_CHECK_ARG(x && dy && dw9c && ws, "dwconv_bwd_weight: null pointer");
  cudaStream_t st = as_stream(stream);
  CVX_WS_ZERO(ws, sizeof(double) * 9 * g.c, st);
  const int64_t rows = (int64_t)g.n * g.ho * g.wo;
  int rc = CVX_OK;
  const int vec = d->dtype == CVX_F32 ? 4 : 8;
  if (int rc2 = dw_tiled_launch(2, d, x, nullptr, dy, nullptr, ws, relu_in, st); rc2 != CVX_EUNSUPPORTED) {
    if (rc2) return rc2;
    dw_copy_d2f_kernel<<<(9 * g.c + 255) / 256, 256, 0, st>>>(ws, dw9c, 9 * g.c);
    CVX_LAUNCH_OK();
    return CVX_OK;
  }
  if (dw_is_s1(g) && (size_t)9 * g.c * sizeof(float) <= 96 * 1024) {
    int blocks; int64_t stride;
    dw_s1_grid(g, g.c / vec, &blocks, &stride);
    const size_t smem = (size_t)9 * g.c * sizeof(float);
    if (d->dtype == CVX_F32) {
      static bool cfg = false;
      if (!cfg) { CVX_CUDA_OK(cudaFuncSetAttribute(dw_s1_kernel<float, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024)); cfg = true; }
    } else {
      static bool cfg = false;
      if (!cfg) { CVX_CUDA_OK(cudaFuncSetAttribute(dw_s1_kernel<__nv_bfloat16, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024)); cfg = true; }
    }
    CVX_DISPATCH_DTYPE(d->dtype, T, (dw_s1_kernel<T, 2><<<blocks, 128, smem, st>>>(
                                        (const T*)x, nullptr, (const T*)dy, nullptr, ws, g.n, g.h, g.w, g.c, relu_in, stride)));
    CVX_LAUNCH_OK();
    dw_copy_d2f_kernel<<<(9 * g.c + 255) / 256, 256, 0, st>>>(ws, dw9c, 9 * g.c);
    CVX_LAUNCH_OK();
    return CVX_OK;
  }
  CVX_DISPATCH_DTYPE(d->dtype, T, rc = (colreduce_launch<T, DwWgradF<T>, 128, 3>(
                                      DwWgradF<T>{(const T*)x, (const T*)dy, g, relu_in}, rows, g.c, ws, st)));
  if (rc) return rc;
  dw_copy_d2f_kernel<<<(9 * g.c + 255) / 256, 256, 0, st>>>(ws, dw9c, 9 * g.c);
  CVX_LAUNCH_OK();
  return CVX_OK;
}

}  // extern "C"
