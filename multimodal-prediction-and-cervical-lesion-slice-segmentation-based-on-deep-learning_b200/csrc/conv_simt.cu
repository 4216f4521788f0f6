// Generic dense convolution as a SIMT implicit GEMM (fp32 accumulate, any geometry).
//
// This is the exact-arithmetic path: fp32 storage for the fp32 parity configuration
// (BASELINE config 1) and the fallback for shapes the tcgen05 kernel does not take
// (C_in = 3 stem, C_out = 5 classifier).  GEMM view:
//   forward : M = N*Ho*Wo pixels, N = C_out, K = taps*C_in
//   dgrad   : M = N*H*W   pixels, N = C_in,  K = taps*C_out   (weights packed transposed+flipped)
//   wgrad   : M = C_out, N = C_in, K = N*Ho*Wo pixels, one GEMM per tap, split over pixels
#include "common.cuh"
#include "conv_narrow.cuh"

namespace cvx {

constexpr int BM = 64, BN = 64, BK = 16;

struct ConvGeom {
  int n, h, w, cin, cout, kh, kw, stride, pad, dil, ho, wo;
};

static ConvGeom geom_of(const cvx_conv_desc* d) {
  return ConvGeom{d->n, d->h, d->w, d->cin, d->cout, d->kh, d->kw, d->stride, d->pad, d->dil, d->ho, d->wo};
}

// MODE 0: forward  (rows = output pixels, reduce over taps x cin, read x)
// MODE 1: dgrad    (rows = input pixels,  reduce over taps x cout, read dy)
template <typename T, int MODE>
__global__ void __launch_bounds__(256) igemm_simt_kernel(const T* __restrict__ src, const T* __restrict__ wp,
                                                         const float* __restrict__ bias, T* __restrict__ dst,
                                                         ConvGeom g) {
  __shared__ float As[BK][BM + 4];
  __shared__ float Bs[BK][BN + 4];

  const int rows_h = MODE == 0 ? g.ho : g.h;
  const int rows_w = MODE == 0 ? g.wo : g.w;
  const int src_h = MODE == 0 ? g.h : g.ho;
  const int src_w = MODE == 0 ? g.w : g.wo;
  const int cred = MODE == 0 ? g.cin : g.cout;   // reduction channels (contiguous in src and wp)
  const int ncol = MODE == 0 ? g.cout : g.cin;   // output channels
  const int64_t M = (int64_t)g.n * rows_h * rows_w;
  const int taps = g.kh * g.kw;

  const int t = threadIdx.x;
  const int64_t m0 = (int64_t)blockIdx.x * BM;
  const int n0 = blockIdx.y * BN;

  // loader role: this thread stages 4 consecutive reduction channels of one row / one column
  const int lrow = t >> 2, lk = (t & 3) * 4;
  const int64_t am = m0 + lrow;
  const bool arow_ok = am < M;
  int a_n = 0, a_y = 0, a_x = 0;
  if (arow_ok) {
    a_x = (int)(am % rows_w);
    a_y = (int)((am / rows_w) % rows_h);
    a_n = (int)(am / ((int64_t)rows_w * rows_h));
  }
  const int bcol = n0 + lrow;
  const bool bcol_ok = bcol < ncol;

  // compute role
  const int ty = t >> 4, tx = t & 15;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int tap = 0; tap < taps; ++tap) {
    const int kh = tap / g.kw, kw = tap % g.kw;
    bool pix_ok = arow_ok;
    int sy, sx;
    if (MODE == 0) {
      sy = a_y * g.stride - g.pad + kh * g.dil;
      sx = a_x * g.stride - g.pad + kw * g.dil;
    } else {
      // tap index runs over the FLIPPED filter: original kh_o = KH-1-kh
      const int ny = a_y + g.pad - (g.kh - 1 - kh) * g.dil;
      const int nx = a_x + g.pad - (g.kw - 1 - kw) * g.dil;
      pix_ok = pix_ok && ny >= 0 && nx >= 0 && (ny % g.stride == 0) && (nx % g.stride == 0);
      sy = ny / g.stride;
      sx = nx / g.stride;
    }
    pix_ok = pix_ok && sy >= 0 && sy < src_h && sx >= 0 && sx < src_w;
    const T* arow = src + (((int64_t)a_n * src_h + sy) * src_w + sx) * cred;
    const T* brow = wp + ((int64_t)tap * ncol + bcol) * cred;

    // software pipeline: the global loads of k-block c0 + BK are issued before the multiply-adds of k-block c0, so
    // their latency hides behind the 16 x 16 FMAs per thread (the row operators' linear layers - 256 rows, K = 1024,
    // 32 CTAs on 148 SMs - were bound by exactly that latency, 64 exposed round trips per launch)
    float ra[4], rb[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int c = lk + i;
      ra[i] = (pix_ok && c < cred) ? Elem<T>::ld(arow + c) : 0.f;
      rb[i] = (bcol_ok && c < cred) ? Elem<T>::ld(brow + c) : 0.f;
    }
    for (int c0 = 0; c0 < cred; c0 += BK) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        As[lk + i][lrow] = ra[i];
        Bs[lk + i][lrow] = rb[i];
      }
      __syncthreads();
      if (c0 + BK < cred) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int c = c0 + BK + lk + i;
          ra[i] = (pix_ok && c < cred) ? Elem<T>::ld(arow + c) : 0.f;
          rb[i] = (bcol_ok && c < cred) ? Elem<T>::ld(brow + c) : 0.f;
        }
      }
#pragma unroll
      for (int kk = 0; kk < BK; ++kk) {
        const float4 a = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
        const float4 b = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
        const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
      }
      __syncthreads();
    }
  }

#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int64_t m = m0 + ty * 4 + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int c = n0 + tx * 4 + j;
      if (c < ncol) Elem<T>::st(dst + m * ncol + c, acc[i][j] + (bias ? bias[c] : 0.f));
    }
  }
}

// wgrad: dw[tap][co][ci] += sum_{m in split} dy[m][co] * x[gather(m,tap)][ci]
template <typename T>
__global__ void __launch_bounds__(256) wgrad_simt_kernel(const T* __restrict__ x, const T* __restrict__ dy,
                                                         float* __restrict__ dw, ConvGeom g, int splits) {
  __shared__ float As[BK][BM + 4];  // dy  [pixel][co]
  __shared__ float Bs[BK][BN + 4];  // x   [pixel][ci]
  const int64_t M = (int64_t)g.n * g.ho * g.wo;
  const int tap = blockIdx.z / splits, split = blockIdx.z % splits;
  const int kh = tap / g.kw, kw = tap % g.kw;
  const int co0 = blockIdx.x * BM, ci0 = blockIdx.y * BN;
  const int64_t per = ceil_div64(ceil_div64(M, splits), BK) * BK;
  const int64_t p_begin = per * split;
  const int64_t p_end = p_begin + per < M ? p_begin + per : M;

  const int t = threadIdx.x;
  const int lk = t >> 4, lc = (t & 15) * 4;  // loader: pixel lk, 4 consecutive channels from lc
  const int ty = t >> 4, tx = t & 15;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int64_t p0 = p_begin; p0 < p_end; p0 += BK) {
    const int64_t m = p0 + lk;
    bool ok = m < p_end;
    int64_t xpix = 0;
    if (ok) {
      const int ox = (int)(m % g.wo);
      const int oy = (int)((m / g.wo) % g.ho);
      const int nn = (int)(m / ((int64_t)g.wo * g.ho));
      const int iy = oy * g.stride - g.pad + kh * g.dil;
      const int ix = ox * g.stride - g.pad + kw * g.dil;
      const bool in = iy >= 0 && iy < g.h && ix >= 0 && ix < g.w;
      xpix = ((int64_t)nn * g.h + iy) * g.w + ix;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int co = co0 + lc + i, ci = ci0 + lc + i;
        As[lk][lc + i] = co < g.cout ? Elem<T>::ld(dy + m * g.cout + co) : 0.f;
        Bs[lk][lc + i] = (in && ci < g.cin) ? Elem<T>::ld(x + xpix * g.cin + ci) : 0.f;
      }
    } else {
#pragma unroll
      for (int i = 0; i < 4; ++i) { As[lk][lc + i] = 0.f; Bs[lk][lc + i] = 0.f; }
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      const float4 a = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
      const float4 b = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int co = co0 + ty * 4 + i;
    if (co >= g.cout) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int ci = ci0 + tx * 4 + j;
      if (ci < g.cin) atomicAdd(dw + ((int64_t)tap * g.cout + co) * g.cin + ci, acc[i][j]);
    }
  }
}

// fwd: y[n,oy,ox,:] = x[n,oy*s,ox*s,:]   bwd: dx[n,iy,ix,:] = (iy%s==0 && ix%s==0) ? dy[n,iy/s,ix/s,:] : 0
// VW = elements moved per thread (a 16-byte vector when the channel count allows, else 1)
template <typename T, int VW>
__global__ void subsample_kernel(const T* __restrict__ x, T* __restrict__ y, int n, int h, int w, int c, int s,
                                 int ho, int wo, int bwd) {
  const int cv = c / VW;
  const int64_t total = (bwd ? (int64_t)n * h * w : (int64_t)n * ho * wo) * cv;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int cc = (int)(i % cv) * VW;
    const int64_t p = i / cv;
    const T* src = nullptr;
    if (!bwd) {
      const int ox = (int)(p % wo), oy = (int)((p / wo) % ho), nn = (int)(p / ((int64_t)wo * ho));
      src = x + (((int64_t)nn * h + oy * s) * w + ox * s) * c + cc;
    } else {
      const int ix = (int)(p % w), iy = (int)((p / w) % h), nn = (int)(p / ((int64_t)w * h));
      if (iy % s == 0 && ix % s == 0 && iy / s < ho && ix / s < wo)
        src = x + (((int64_t)nn * ho + iy / s) * wo + ix / s) * c + cc;
    }
    T* dst = y + p * c + cc;
    if (VW == 1) {
      T v;
      Elem<T>::st(&v, 0.f);
      *dst = src ? *src : v;
    } else {
      *reinterpret_cast<uint4*>(dst) = src ? *reinterpret_cast<const uint4*>(src) : make_uint4(0, 0, 0, 0);
    }
  }
}

template <typename T>
static void subsample_launch(const T* x, T* y, int n, int h, int w, int c, int s, int ho, int wo, int bwd,
                             cudaStream_t st) {
  constexpr int V = Elem<T>::kVec;
  const bool vec = (c % V == 0) && ((uintptr_t)x % 16 == 0) && ((uintptr_t)y % 16 == 0);
  const int64_t total = (bwd ? (int64_t)n * h * w : (int64_t)n * ho * wo) * (vec ? c / V : c);
  const int64_t b = ceil_div64(total, 256);
  const int grid = (int)(b > (int64_t)kNumSMs * 16 ? (int64_t)kNumSMs * 16 : (b < 1 ? 1 : b));
  if (vec) subsample_kernel<T, V><<<grid, 256, 0, st>>>(x, y, n, h, w, c, s, ho, wo, bwd);
  else subsample_kernel<T, 1><<<grid, 256, 0, st>>>(x, y, n, h, w, c, s, ho, wo, bwd);
}

static int check_desc(const cvx_conv_desc* d, const char* who) {
  CVX_CHECK_ARG(d != nullptr, "%s: null descriptor", who);
  CVX_CHECK_ARG(d->n > 0 && d->h > 0 && d->w > 0 && d->cin > 0 && d->cout > 0 && d->kh > 0 && d->kw > 0 &&
                    d->stride > 0 && d->dil > 0 && d->pad >= 0,
                "%s: bad geometry", who);
  const int ho = (d->h + 2 * d->pad - d->dil * (d->kh - 1) - 1) / d->stride + 1;
  const int wo = (d->w + 2 * d->pad - d->dil * (d->kw - 1) - 1) / d->stride + 1;
  CVX_CHECK_ARG(ho == d->ho && wo == d->wo, "%s: ho/wo (%d,%d) inconsistent with geometry (%d,%d)", who, d->ho,
                d->wo, ho, wo);
  return CVX_OK;
}

}  // namespace cvx

using namespace cvx;

extern "C" {

int cvx_conv_fwd(const cvx_conv_desc* d, const void* x, const void* w_packed, const float* bias, void* y,
                 void* stream) {
  if (int rc = check_desc(d, "conv_fwd")) return rc;
  CVX_CHECK_ARG(x && w_packed && y, "conv_fwd: null pointer");
  if (int rc = narrow_conv_fwd(d, x, w_packed, bias, y, as_stream(stream)); rc != CVX_EUNSUPPORTED) return rc;
  const ConvGeom g = geom_of(d);
  const int64_t M = (int64_t)g.n * g.ho * g.wo;
  dim3 grid((unsigned)ceil_div64(M, BM), (g.cout + BN - 1) / BN);
  CVX_DISPATCH_DTYPE(d->dtype, T, (igemm_simt_kernel<T, 0><<<grid, 256, 0, as_stream(stream)>>>(
                                      (const T*)x, (const T*)w_packed, bias, (T*)y, g)));
  CVX_LAUNCH_OK();
  return CVX_OK;
}

int cvx_conv_dgrad(const cvx_conv_desc* d, const void* dy, const void* w_packed_t, void* dx, void* stream) {
  if (int rc = check_desc(d, "conv_dgrad")) return rc;
  CVX_CHECK_ARG(dy && w_packed_t && dx, "conv_dgrad: null pointer");
  if (int rc = narrow_conv_dgrad(d, dy, w_packed_t, dx, as_stream(stream)); rc != CVX_EUNSUPPORTED) return rc;
  const ConvGeom g = geom_of(d);
  const int64_t M = (int64_t)g.n * g.h * g.w;
  dim3 grid((unsigned)ceil_div64(M, BM), (g.cin + BN - 1) / BN);
  CVX_DISPATCH_DTYPE(d->dtype, T, (igemm_simt_kernel<T, 1><<<grid, 256, 0, as_stream(stream)>>>(
                                      (const T*)dy, (const T*)w_packed_t, nullptr, (T*)dx, g)));
  CVX_LAUNCH_OK();
  return CVX_OK;
}

int cvx_conv_wgrad(const cvx_conv_desc* d, const void* x, const void* dy, float* dw_packed, void* stream) {
  if (int rc = check_desc(d, "conv_wgrad")) return rc;
  CVX_CHECK_ARG(x && dy && dw_packed, "conv_wgrad: null pointer");
  if (int rc = narrow_conv_wgrad(d, x, dy, dw_packed, as_stream(stream)); rc != CVX_EUNSUPPORTED) return rc;
  const ConvGeom g = geom_of(d);
  const int64_t M = (int64_t)g.n * g.ho * g.wo;
  const int taps = g.kh * g.kw;
  const int tiles = ((g.cout + BM - 1) / BM) * ((g.cin + BN - 1) / BN) * taps;
  int splits = (2 * kNumSMs + tiles - 1) / tiles;
  // the row operators of the fusion head (M = patients x nodes <= a few thousand rows): one block per output tile, so
  // every element is one fixed-order sum and the head's gradients are bit-reproducible (no split-K atomics)
  if (M <= 4096) splits = 1;
  const int64_t max_splits = ceil_div64(M, 4 * BK);
  if (splits > max_splits) splits = (int)max_splits;
  if (splits < 1) splits = 1;
  CVX_CHECK_ARG((int64_t)taps * splits <= 65535, "conv_wgrad: grid.z too large");
  dim3 grid((g.cout + BM - 1) / BM, (g.cin + BN - 1) / BN, taps * splits);
  CVX_DISPATCH_DTYPE(d->dtype, T, (wgrad_simt_kernel<T><<<grid, 256, 0, as_stream(stream)>>>(
                                      (const T*)x, (const T*)dy, dw_packed, g, splits)));
  CVX_LAUNCH_OK();
  return CVX_OK;
}

int cvx_subsample(const void* x, void* y, int n, int h, int w, int c, int s, int dtype, void* stream) {
  CVX_CHECK_ARG(x && y && n > 0 && h > 0 && w > 0 && c > 0 && s > 0, "subsample: bad arguments");
  const int ho = (h - 1) / s + 1, wo = (w - 1) / s + 1;
  CVX_DISPATCH_DTYPE(dtype, T, (subsample_launch<T>((const T*)x, (T*)y, n, h, w, c, s, ho, wo, 0, as_stream(stream))));
  CVX_LAUNCH_OK();
  return CVX_OK;
}

int cvx_subsample_bwd(const void* dy, void* dx, int n, int h, int w, int c, int s, int dtype, void* stream) {
  CVX_CHECK_ARG(dy && dx && n > 0 && h > 0 && w > 0 && c > 0 && s > 0, "subsample_bwd: bad arguments");
  const int ho = (h - 1) / s + 1, wo = (w - 1) / s + 1;
  CVX_DISPATCH_DTYPE(dtype, T, (subsample_launch<T>((const T*)dy, (T*)dx, n, h, w, c, s, ho, wo, 1, as_stream(stream))));
  CVX_LAUNCH_OK();
  return CVX_OK;
}

}  // extern "C"
