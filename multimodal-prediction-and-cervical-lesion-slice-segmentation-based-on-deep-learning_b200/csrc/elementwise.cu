// Small bandwidth-bound ops: ReLU, add, ASPP global pooling / broadcast, bilinear
// (align_corners=True) upsampling forward/backward, dropout.
#include "common.cuh"

namespace cvx {

static inline int ew_grid(int64_t total, int block = 256) {
  int64_t b = ceil_div64(total, block);
  const int64_t cap = (int64_t)kNumSMs * 16;
  return (int)(b > cap ? cap : (b < 1 ? 1 : b));
}

#define CVX_GRID_STRIDE(i, total) \
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < (total); i += (int64_t)gridDim.x * blockDim.x)

template <typename T>
__global__ void relu_fwd_kernel(const T* __restrict__ x, T* __restrict__ y, int64_t nvec, int64_t n) {
  pdl_trigger();
  pdl_wait();
  constexpr int VEC = Elem<T>::kVec;
  CVX_GRID_STRIDE(i, nvec) {
    Vec<T> v;
    v.load(x + i * VEC);
#pragma unroll
    for (int k = 0; k < VEC; ++k) v.v[k] = fmaxf(v.v[k], 0.f);
    v.store(y + i * VEC);
  }
  // tail
  const int64_t t0 = nvec * VEC + blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (t0 < n) Elem<T>::st(y + t0, fmaxf(Elem<T>::ld(x + t0), 0.f));
}

template <typename T>
__global__ void relu_bwd_kernel(const T* __restrict__ dy, const T* __restrict__ y, T* __restrict__ dx, int64_t nvec,
                                int64_t n) {
  pdl_trigger();
  pdl_wait();
  constexpr int VEC = Elem<T>::kVec;
  CVX_GRID_STRIDE(i, nvec) {
    Vec<T> g, v;
    g.load(dy + i * VEC);
    v.load(y + i * VEC);
#pragma unroll
    for (int k = 0; k < VEC; ++k) g.v[k] = v.v[k] > 0.f ? g.v[k] : 0.f;
    g.store(dx + i * VEC);
  }
  const int64_t t0 = nvec * VEC + blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (t0 < n) Elem<T>::st(dx + t0, Elem<T>::ld(y + t0) > 0.f ? Elem<T>::ld(dy + t0) : 0.f);
}

template <typename T>
__global__ void add_kernel(const T* __restrict__ a, const T* __restrict__ b, T* __restrict__ o, int64_t nvec,
                           int64_t n) {
  pdl_trigger();
  pdl_wait();
  constexpr int VEC = Elem<T>::kVec;
  CVX_GRID_STRIDE(i, nvec) {
    Vec<T> u, v;
    u.load(a + i * VEC);
    v.load(b + i * VEC);
#pragma unroll
    for (int k = 0; k < VEC; ++k) u.v[k] += v.v[k];
    u.store(o + i * VEC);
  }
  const int64_t t0 = nvec * VEC + blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (t0 < n) Elem<T>::st(o + t0, Elem<T>::ld(a + t0) + Elem<T>::ld(b + t0));
}

// y[n,c] = scale * sum_p x[n,p,c] ; grid (ceil(c/256), n)
template <typename T>
__global__ void spatial_reduce_kernel(const T* __restrict__ x, T* __restrict__ y, int hw, int c, float scale) {
  const int cc = blockIdx.x * blockDim.x + threadIdx.x;
  if (cc >= c) return;
  const T* p = x + (size_t)blockIdx.y * hw * c + cc;
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
  int i = 0;
  for (; i + 3 < hw; i += 4) {
    s0 += Elem<T>::ld(p + (size_t)i * c);
    s1 += Elem<T>::ld(p + (size_t)(i + 1) * c);
    s2 += Elem<T>::ld(p + (size_t)(i + 2) * c);
    s3 += Elem<T>::ld(p + (size_t)(i + 3) * c);
  }
  for (; i < hw; ++i) s0 += Elem<T>::ld(p + (size_t)i * c);
  Elem<T>::st(y + (size_t)blockIdx.y * c + cc, (s0 + s1 + s2 + s3) * scale);
}

template <typename T>
__global__ void spatial_broadcast_kernel(const T* __restrict__ x, T* __restrict__ y, int64_t total, int hw, int c,
                                         float scale) {
  CVX_GRID_STRIDE(i, total) {
    const int cc = (int)(i % c);
    const int64_t nn = i / ((int64_t)hw * c);
    Elem<T>::st(y + i, Elem<T>::ld(x + nn * c + cc) * scale);
  }
}

// vector form for C % VEC == 0 (ASPP's pooled branch, [B,2048] -> [B,32,32,2048] in backward): a thread keeps one
// 16-byte channel vector of its image in registers and stores it to a strip of pixels - no per-element division
template <typename T>
__global__ void __launch_bounds__(256) spatial_broadcast_vec_kernel(const T* __restrict__ x, T* __restrict__ y, int hw, int c,
                                                                    float scale) {
  constexpr int VEC = Elem<T>::kVec;
  const int cvn = c / VEC;
  const int nn = blockIdx.z;
  T* out = y + (size_t)nn * hw * c;
  for (int cv = blockIdx.x * blockDim.x + threadIdx.x; cv < cvn; cv += gridDim.x * blockDim.x) {
    Vec<T> v;
    v.load(x + (size_t)nn * c + cv * VEC);
#pragma unroll
    for (int i = 0; i < VEC; ++i) v.v[i] *= scale;
    for (int p = blockIdx.y; p < hw; p += gridDim.y) v.store(out + (size_t)p * c + cv * VEC);
  }
}

// ---- bilinear, align_corners=True (same index arithmetic as ATen's upsample_bilinear2d) -----
struct Lerp {
  int i0, i1;
  float w0, w1;
};
__device__ __forceinline__ Lerp lerp_of(int o, float scale, int in_size) {
  const float r = scale * (float)o;
  Lerp l;
  l.i0 = (int)r;
  if (l.i0 > in_size - 1) l.i0 = in_size - 1;
  l.i1 = l.i0 + (l.i0 < in_size - 1 ? 1 : 0);
  l.w1 = r - (float)l.i0;
  l.w0 = 1.f - l.w1;
  return l;
}
static inline float lerp_scale(int in_size, int out_size) {
  return out_size > 1 ? (float)(in_size - 1) / (float)(out_size - 1) : 0.f;
}
// total weight with which input index `i` contributes to output index `o`
__device__ __forceinline__ float lerp_weight(int i, int o, float scale, int in_size) {
  const Lerp l = lerp_of(o, scale, in_size);
  return (l.i0 == i ? l.w0 : 0.f) + (l.i1 == i ? l.w1 : 0.f);
}
// candidate output range that can touch input index i
__device__ __forceinline__ void lerp_range(int i, float scale, int out_size, int* lo, int* hi) {
  if (scale <= 0.f) { *lo = 0; *hi = out_size - 1; return; }
  int a = (int)floorf((float)(i - 1) / scale) - 1;
  int b = (int)ceilf((float)(i + 1) / scale) + 1;
  *lo = a < 0 ? 0 : a;
  *hi = b > out_size - 1 ? out_size - 1 : b;
}

// One block row = one output row (blockIdx.x = image * ho + oy): the row's interpolation weights are block-uniform and
// the element index inside the row is 32-bit (the flat 64-bit index with three 64-bit divisions per 16-byte store made
// this kernel issue-bound at 1.5 TB/s).
template <typename T>
__global__ void __launch_bounds__(256) upsample_fwd_kernel(const T* __restrict__ x, T* __restrict__ y, int hi, int wi, int ho,
                                                           int wo, int c, float sh, float sw, int ldy, int yoff) {
  constexpr int VEC = Elem<T>::kVec;
  const int cvn = c / VEC;
  const int oy = blockIdx.x % ho, nn = blockIdx.x / ho;
  const Lerp ly = lerp_of(oy, sh, hi);
  const T* row0 = x + ((size_t)nn * hi + ly.i0) * wi * c;
  const T* row1 = x + ((size_t)nn * hi + ly.i1) * wi * c;
  T* out = y + ((size_t)nn * ho + oy) * wo * ldy + yoff;   // ldy > c: a channel slice of a wider (concat) tensor
  const int per_row = wo * cvn;
  for (int e = blockIdx.y * blockDim.x + threadIdx.x; e < per_row; e += gridDim.y * blockDim.x) {
    const int ox = e / cvn, c0 = (e - ox * cvn) * VEC;
    const Lerp lx = lerp_of(ox, sw, wi);
    Vec<T> a, b, cc, d, o;
    a.load(row0 + (size_t)lx.i0 * c + c0);
    b.load(row0 + (size_t)lx.i1 * c + c0);
    cc.load(row1 + (size_t)lx.i0 * c + c0);
    d.load(row1 + (size_t)lx.i1 * c + c0);
#pragma unroll
    for (int i = 0; i < VEC; ++i)
      o.v[i] = ly.w0 * (lx.w0 * a.v[i] + lx.w1 * b.v[i]) + ly.w1 * (lx.w0 * cc.v[i] + lx.w1 * d.v[i]);
    o.store(out + (size_t)ox * ldy + c0);
  }
}

// gradient (gather): one block row = one INPUT row (blockIdx.x = image * hi + iy), 32-bit indexing inside the row
template <typename T>
__global__ void __launch_bounds__(256) upsample_bwd_kernel(const T* __restrict__ dy, T* __restrict__ dx, int hi, int wi, int ho,
                                                           int wo, int c, float sh, float sw, int lddy, int dyoff) {
  constexpr int VEC = Elem<T>::kVec;
  const int cvn = c / VEC;
  const int iy = blockIdx.x % hi, nn = blockIdx.x / hi;
  int ylo, yhi;
  lerp_range(iy, sh, ho, &ylo, &yhi);
  const int per_row = wi * cvn;
  for (int e = blockIdx.y * blockDim.x + threadIdx.x; e < per_row; e += gridDim.y * blockDim.x) {
    const int ix = e / cvn, c0 = (e - ix * cvn) * VEC;
    int xlo, xhi;
    lerp_range(ix, sw, wo, &xlo, &xhi);
    float acc[VEC];
#pragma unroll
    for (int i = 0; i < VEC; ++i) acc[i] = 0.f;
    for (int oy = ylo; oy <= yhi; ++oy) {
      const float wy = lerp_weight(iy, oy, sh, hi);
      if (wy == 0.f) continue;
      const T* grow = dy + (((size_t)nn * ho + oy) * wo) * lddy + dyoff + c0;
      for (int ox = xlo; ox <= xhi; ++ox) {
        const float wgt = wy * lerp_weight(ix, ox, sw, wi);
        if (wgt == 0.f) continue;
        Vec<T> g;
        g.load(grow + (size_t)ox * lddy);
#pragma unroll
        for (int i = 0; i < VEC; ++i) acc[i] = fmaf(wgt, g.v[i], acc[i]);
      }
    }
    Vec<T> o;
#pragma unroll
    for (int i = 0; i < VEC; ++i) o.v[i] = acc[i];
    o.store(dx + (((size_t)nn * hi + iy) * wi) * c + (size_t)e * VEC);
  }
}

// low-res NHWC (any small C) -> full-res NCHW fp32
template <typename T>
__global__ void upsample_to_nchw_fwd_kernel(const T* __restrict__ x, float* __restrict__ y, int n, int hi, int wi,
                                            int ho, int wo, int c, float sh, float sw) {
  const int64_t total = (int64_t)n * ho * wo;
  CVX_GRID_STRIDE(p, total) {
    const int ox = (int)(p % wo), oy = (int)((p / wo) % ho), nn = (int)(p / ((int64_t)wo * ho));
    const Lerp ly = lerp_of(oy, sh, hi), lx = lerp_of(ox, sw, wi);
    const T* base = x + (size_t)nn * hi * wi * c;
    const T* p00 = base + ((size_t)ly.i0 * wi + lx.i0) * c;
    const T* p01 = base + ((size_t)ly.i0 * wi + lx.i1) * c;
    const T* p10 = base + ((size_t)ly.i1 * wi + lx.i0) * c;
    const T* p11 = base + ((size_t)ly.i1 * wi + lx.i1) * c;
    for (int cc = 0; cc < c; ++cc) {
      const float v = ly.w0 * (lx.w0 * Elem<T>::ld(p00 + cc) + lx.w1 * Elem<T>::ld(p01 + cc)) +
                      ly.w1 * (lx.w0 * Elem<T>::ld(p10 + cc) + lx.w1 * Elem<T>::ld(p11 + cc));
      y[(((size_t)nn * c + cc) * ho + oy) * wo + ox] = v;
    }
  }
}

template <typename T>
__global__ void upsample_to_nchw_bwd_kernel(const float* __restrict__ dy, T* __restrict__ dx, int n, int hi, int wi,
                                            int ho, int wo, int c, float sh, float sw) {
  // e enumerates [n][c][iy][ix] with ix fastest: the threads of a warp gather from adjacent windows of the SAME
  // rows of one dy plane (coalesced reads of the large fp32 NCHW gradient); the small NHWC result takes the
  // strided 2-byte stores instead.
  const int64_t total = (int64_t)n * hi * wi * c;
  CVX_GRID_STRIDE(e, total) {
    const int ix = (int)(e % wi);
    int64_t q = e / wi;
    const int iy = (int)(q % hi); q /= hi;
    const int cc = (int)(q % c), nn = (int)(q / c);
    int ylo, yhi, xlo, xhi;
    lerp_range(iy, sh, ho, &ylo, &yhi);
    lerp_range(ix, sw, wo, &xlo, &xhi);
    const float* plane = dy + ((size_t)nn * c + cc) * ho * wo;
    float acc = 0.f;
    for (int oy = ylo; oy <= yhi; ++oy) {
      const float wy = lerp_weight(iy, oy, sh, hi);
      if (wy == 0.f) continue;
      for (int ox = xlo; ox <= xhi; ++ox) {
        const float wgt = wy * lerp_weight(ix, ox, sw, wi);
        if (wgt != 0.f) acc = fmaf(wgt, __ldg(plane + (size_t)oy * wo + ox), acc);
      }
    }
    Elem<T>::st(dx + (((size_t)nn * hi + iy) * wi + ix) * c + cc, acc);
  }
}

// Separable form of the gradient above for the final x4 upsample (fp32 NCHW 512^2 planes -> NHWC 128^2): one block
// per (image, class, input row).  Phase 1 folds the <= 2*ceil(1/scale)+3 contributing output rows into one weighted
// row in shared memory (coalesced row reads of dy, one weight per row); phase 2 gathers each input column from that
// row.  ~(rows + columns) multiply-adds per result instead of rows x columns with the weights recomputed per tap:
// the gather kernel was bound by instruction issue (79 % issue-active at 6 % of DRAM bandwidth in ncu).
constexpr int kUpBwdMaxRows = 24;
template <typename T>
__global__ void __launch_bounds__(128) upsample_to_nchw_bwd_rows_kernel(const float* __restrict__ dy, T* __restrict__ dx,
                                                                        int hi, int wi, int ho, int wo, int c, float sh,
                                                                        float sw) {
  extern __shared__ float rowbuf[];          // [wo]
  __shared__ float wy[kUpBwdMaxRows];
  const int iy = blockIdx.x % hi;
  const int cc = (blockIdx.x / hi) % c;
  const int nn = blockIdx.x / (hi * c);
  int ylo, yhi;
  lerp_range(iy, sh, ho, &ylo, &yhi);
  const int nrows = yhi - ylo + 1;           // <= kUpBwdMaxRows (checked by the launcher)
  if (threadIdx.x < nrows) wy[threadIdx.x] = lerp_weight(iy, ylo + threadIdx.x, sh, hi);
  __syncthreads();
  const float* plane = dy + ((size_t)nn * c + cc) * ho * wo;
  for (int ox = threadIdx.x; ox < wo; ox += blockDim.x) {
    float acc = 0.f;
    for (int r = 0; r < nrows; ++r) {
      const float w = wy[r];
      if (w != 0.f) acc = fmaf(w, __ldg(plane + (size_t)(ylo + r) * wo + ox), acc);
    }
    rowbuf[ox] = acc;
  }
  __syncthreads();
  for (int ix = threadIdx.x; ix < wi; ix += blockDim.x) {
    int xlo, xhi;
    lerp_range(ix, sw, wo, &xlo, &xhi);
    float acc = 0.f;
    for (int ox = xlo; ox <= xhi; ++ox) {
      const float w = lerp_weight(ix, ox, sw, wi);
      if (w != 0.f) acc = fmaf(w, rowbuf[ox], acc);
    }
    Elem<T>::st(dx + (((size_t)nn * hi + iy) * wi + ix) * c + cc, acc);
  }
}

// ---- max pooling 3x3 stride 2 pad 1 (torchvision ResNet stem) ----------------------------
template <typename T>
__global__ void maxpool3x3s2_fwd_kernel(const T* __restrict__ x, T* __restrict__ y, int n, int h, int w, int c, int ho,
                                        int wo) {
  constexpr int VEC = Elem<T>::kVec;
  const int cvn = c / VEC;
  const int64_t total = (int64_t)n * ho * wo * cvn;
  CVX_GRID_STRIDE(e, total) {
    const int c0 = (int)(e % cvn) * VEC;
    int64_t p = e / cvn;
    const int ox = (int)(p % wo), oy = (int)((p / wo) % ho), nn = (int)(p / ((int64_t)wo * ho));
    float m[VEC];
#pragma unroll
    for (int i = 0; i < VEC; ++i) m[i] = -INFINITY;
#pragma unroll
    for (int kh = 0; kh < 3; ++kh) {
      const int iy = oy * 2 - 1 + kh;
      if (iy < 0 || iy >= h) continue;
#pragma unroll
      for (int kw = 0; kw < 3; ++kw) {
        const int ix = ox * 2 - 1 + kw;
        if (ix < 0 || ix >= w) continue;
        Vec<T> v;
        v.load(x + (((size_t)nn * h + iy) * w + ix) * c + c0);
#pragma unroll
        for (int i = 0; i < VEC; ++i) m[i] = fmaxf(m[i], v.v[i]);
      }
    }
    Vec<T> o;
#pragma unroll
    for (int i = 0; i < VEC; ++i) o.v[i] = m[i];
    o.store(y + e * VEC);
  }
}

// gradient goes to the FIRST maximal element of each window (ATen's tie rule, row-major scan)
template <typename T>
__global__ void maxpool3x3s2_bwd_kernel(const T* __restrict__ x, const T* __restrict__ y, const T* __restrict__ dy,
                                        T* __restrict__ dx, int n, int h, int w, int c, int ho, int wo) {
  constexpr int VEC = Elem<T>::kVec;
  const int cvn = c / VEC;
  const int64_t total = (int64_t)n * h * w * cvn;
  CVX_GRID_STRIDE(e, total) {
    const int c0 = (int)(e % cvn) * VEC;
    int64_t p = e / cvn;
    const int ix = (int)(p % w), iy = (int)((p / w) % h), nn = (int)(p / ((int64_t)w * h));
    Vec<T> xv;
    xv.load(x + e * VEC);
    float acc[VEC];
#pragma unroll
    for (int i = 0; i < VEC; ++i) acc[i] = 0.f;
    for (int oy = (iy + 1) / 2 - 1; oy <= (iy + 1) / 2; ++oy) {
      if (oy < 0 || oy >= ho || iy < oy * 2 - 1 || iy > oy * 2 + 1) continue;
      for (int ox = (ix + 1) / 2 - 1; ox <= (ix + 1) / 2; ++ox) {
        if (ox < 0 || ox >= wo || ix < ox * 2 - 1 || ix > ox * 2 + 1) continue;
        const size_t o = (((size_t)nn * ho + oy) * wo + ox) * c + c0;
        Vec<T> yv, gv;
        yv.load(y + o);
        gv.load(dy + o);
        // is (iy,ix) the first position of this window that attains the max?
#pragma unroll
        for (int i = 0; i < VEC; ++i) {
          if (xv.v[i] != yv.v[i]) continue;
          bool first = true;
          for (int kh = 0; kh < 3 && first; ++kh) {
            const int yy = oy * 2 - 1 + kh;
            if (yy < 0 || yy >= h) continue;
            for (int kw = 0; kw < 3; ++kw) {
              const int xx = ox * 2 - 1 + kw;
              if (xx < 0 || xx >= w) continue;
              if (yy == iy && xx == ix) { kh = 3; break; }
              if (Elem<T>::ld(x + (((size_t)nn * h + yy) * w + xx) * c + c0 + i) == yv.v[i]) { first = false; break; }
            }
          }
          if (first) acc[i] += gv.v[i];
        }
      }
    }
    Vec<T> o;
#pragma unroll
    for (int i = 0; i < VEC; ++i) o.v[i] = acc[i];
    o.store(dx + e * VEC);
  }
}

// ---- dropout -------------------------------------------------------------------------
__device__ __forceinline__ uint32_t mix64(uint64_t z) {
  z += 0x9e3779b97f4a7c15ull;
  z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
  z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
  return (uint32_t)((z ^ (z >> 31)) >> 32);
}

__device__ __forceinline__ uint32_t dropout_thresh(float p) {
  return (uint32_t)((double)p * 4294967296.0 > 4294967295.0 ? 4294967295.0 : (double)p * 4294967296.0);
}
// element i of the tensor is kept iff the counter-based hash of (seed, step, i) clears the threshold: the backward pass
// recomputes the decision instead of reading a stored mask (mask == nullptr: none is written - 1 byte per element of
// traffic saved in each direction)
__device__ __forceinline__ bool dropout_keep(uint64_t seed, int64_t i, uint32_t thresh) {
  return mix64(seed * 0x100000001b3ull + (uint64_t)i) >= thresh;
}

// FWD: y = keep ? x / (1 - p) : 0  (writes the mask when one is given); !FWD: the same map applied to dy.
// One 16-byte vector per thread and trip; the scalar tail (n % VEC) is handled by the first threads.
template <typename T>
__global__ void __launch_bounds__(256) dropout_kernel(const T* __restrict__ x, T* __restrict__ y, uint8_t* __restrict__ mask,
                                                      int64_t n, float p, uint64_t seed, const int* __restrict__ step_dev) {
  constexpr int VEC = Elem<T>::kVec;
  pdl_trigger();
  pdl_wait();
  if (step_dev) seed += 0x9e3779b97f4a7c15ull * (uint64_t)(*step_dev);  // per-step stream under CUDA-graph replay
  const float keep_scale = 1.f / (1.f - p);
  const uint32_t thresh = dropout_thresh(p);
  const int64_t nvec = n / VEC;
  CVX_GRID_STRIDE(v, nvec) {
    Vec<T> a;
    a.load(x + v * VEC);
    uint8_t m[VEC];
#pragma unroll
    for (int k = 0; k < VEC; ++k) {
      m[k] = dropout_keep(seed, v * VEC + k, thresh) ? 1 : 0;
      a.v[k] = m[k] ? a.v[k] * keep_scale : 0.f;
    }
    a.store(y + v * VEC);
    if (mask) {
#pragma unroll
      for (int k = 0; k < VEC; ++k) mask[v * VEC + k] = m[k];
    }
  }
  const int64_t t0 = nvec * VEC + blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (t0 < n) {
    const bool keep = dropout_keep(seed, t0, thresh);
    if (mask) mask[t0] = keep ? 1 : 0;
    Elem<T>::st(y + t0, keep ? Elem<T>::ld(x + t0) * keep_scale : 0.f);
  }
}

template <typename T>
__global__ void dropout_bwd_kernel(const T* __restrict__ dy, const uint8_t* __restrict__ mask, T* __restrict__ dx,
                                   int64_t n, float p) {
  pdl_trigger();
  pdl_wait();
  const float keep_scale = 1.f / (1.f - p);
  CVX_GRID_STRIDE(i, n) { Elem<T>::st(dx + i, mask[i] ? Elem<T>::ld(dy + i) * keep_scale : 0.f); }
}

}  // namespace cvx

using namespace cvx;

extern "C" {

int cvx_relu_fwd(const void* x, void* y, int64_t n, int dtype, void* stream) {
  CVX_CHECK_ARG(x && y && n > 0, "relu_fwd: bad arguments");
  CVX_CHECK_ARG((uintptr_t)x % 16 == 0 && (uintptr_t)y % 16 == 0, "relu_fwd: pointers must be 16-byte aligned");
  const int vec = dtype == CVX_F32 ? 4 : 8;
  const int64_t nvec = n / vec;
  CVX_DISPATCH_DTYPE(dtype, T, (launch_pdl(relu_fwd_kernel<T>, dim3(ew_grid(nvec)), dim3(256), 0, as_stream(stream), (const T*)x, (T*)y, nvec, n)));
  CVX_LAUNCH_OK();
  return CVX_OK;
}

int cvx_relu_bwd(const void* dy, const void* y, void* dx, int64_t n, int dtype, void* stream) {
  CVX_CHECK_ARG(dy && y && dx && n > 0, "relu_bwd: bad arguments");
  CVX_CHECK_ARG((uintptr_t)dy % 16 == 0 && (uintptr_t)y % 16 == 0 && (uintptr_t)dx % 16 == 0,
                "relu_bwd: pointers must be 16-byte aligned");
  const int vec = dtype == CVX_F32 ? 4 : 8;
  const int64_t nvec = n / vec;
  CVX_DISPATCH_DTYPE(dtype, T, (launch_pdl(relu_bwd_kernel<T>, dim3(ew_grid(nvec)), dim3(256), 0, as_stream(stream), (const T*)dy, (const T*)y, (T*)dx, nvec, n)));
  CVX_LAUNCH_OK();
  return CVX_OK;
}

int cvx_add(const void* a, const void* b, void* out, int64_t n, int dtype, void* stream) {
  CVX_CHECK_ARG(a && b && out && n > 0, "add: bad arguments");
  CVX_CHECK_ARG((uintptr_t)a % 16 == 0 && (uintptr_t)b % 16 == 0 && (uintptr_t)out % 16 == 0,
                "add: pointers must be 16-byte aligned");
  const int vec = dtype == CVX_F32 ? 4 : 8;
  const int64_t nvec = n / vec;
  CVX_DISPATCH_DTYPE(dtype, T, (launch_pdl(add_kernel<T>, dim3(ew_grid(nvec)), dim3(256), 0, as_stream(stream), (const T*)a, (const T*)b, (T*)out, nvec, n)));
  CVX_LAUNCH_OK();
  return CVX_OK;
}

int cvx_spatial_reduce(const void* x, void* y, int n, int hw, int c, float scale, int dtype, void* stream) {
  CVX_CHECK_ARG(x && y && n > 0 && hw > 0 && c > 0 && n <= 65535, "spatial_reduce: bad arguments");
  dim3 grid((c + 255) / 256, n);
  CVX_DISPATCH_DTYPE(dtype, T, (spatial_reduce_kernel<T><<<grid, 256, 0, as_stream(stream)>>>((const T*)x, (T*)y, hw, c, scale)));
  CVX_LAUNCH_OK();
  return CVX_OK;
}

int cvx_spatial_broadcast(const void* x, void* y, int n, int hw, int c, float scale, int dtype, void* stream) {
  CVX_CHECK_ARG(x && y && n > 0 && hw > 0 && c > 0, "spatial_broadcast: bad arguments");
  const int64_t total = (int64_t)n * hw * c;
  const int vec = dtype == CVX_F32 ? 4 : 8;
  if (c % vec == 0 && n <= 65535 && (((uintptr_t)x | (uintptr_t)y) & 15) == 0) {
    const int cvn = c / vec;
    const int gy = hw < 64 ? hw : 64;              // 64 pixel strips per image: n * 64 * ceil(cvn / 256) blocks
    const dim3 grid((unsigned)((cvn + 255) / 256), (unsigned)gy, (unsigned)n);
    CVX_DISPATCH_DTYPE(dtype, T, (spatial_broadcast_vec_kernel<T><<<grid, 256, 0, as_stream(stream)>>>((const T*)x, (T*)y, hw, c, scale)));
    CVX_LAUNCH_OK();
    return CVX_OK;
  }
  CVX_DISPATCH_DTYPE(dtype, T, (spatial_broadcast_kernel<T><<<ew_grid(total), 256, 0, as_stream(stream)>>>((const T*)x, (T*)y, total, hw, c, scale)));
  CVX_LAUNCH_OK();
  return CVX_OK;
}

static int upsample_fwd_launch(const void* x, void* y, int n, int hi, int wi, int ho, int wo, int c, int ldy, int yoff,
                               int dtype, void* stream, const char* who) {
  CVX_CHECK_ARG(x && y && n > 0 && hi > 0 && wi > 0 && ho > 0 && wo > 0 && c > 0, "%s: bad arguments", who);
  const int vec = dtype == CVX_F32 ? 4 : 8;
  CVX_CHECK_ARG(c % vec == 0 && ldy % vec == 0 && yoff % vec == 0 && yoff >= 0 && yoff + c <= ldy,
                "%s: C=%d, row pitch %d and channel offset %d must be multiples of %d", who, c, ldy, yoff, vec);
  CVX_CHECK_ARG((int64_t)n * ho < (1ll << 31) && (int64_t)wo * (c / vec) < (1ll << 24), "%s: tensor too large", who);
  const int per_row = wo * (c / vec);
  const dim3 grid((unsigned)(n * ho), (unsigned)((per_row + 1023) / 1024));   // 4 elements per thread
  CVX_DISPATCH_DTYPE(dtype, T, (upsample_fwd_kernel<T><<<grid, 256, 0, as_stream(stream)>>>(
                                   (const T*)x, (T*)y, hi, wi, ho, wo, c, lerp_scale(hi, ho), lerp_scale(wi, wo), ldy, yoff)));
  CVX_LAUNCH_OK();
  return CVX_OK;
}

static int upsample_bwd_launch(const void* dy, void* dx, int n, int hi, int wi, int ho, int wo, int c, int lddy, int dyoff,
                               int dtype, void* stream, const char* who) {
  CVX_CHECK_ARG(dy && dx && n > 0 && hi > 0 && wi > 0 && ho > 0 && wo > 0 && c > 0, "%s: bad arguments", who);
  const int vec = dtype == CVX_F32 ? 4 : 8;
  CVX_CHECK_ARG(c % vec == 0 && lddy % vec == 0 && dyoff % vec == 0 && dyoff >= 0 && dyoff + c <= lddy,
                "%s: C=%d, row pitch %d and channel offset %d must be multiples of %d", who, c, lddy, dyoff, vec);
  CVX_CHECK_ARG((int64_t)n * hi < (1ll << 31) && (int64_t)wi * (c / vec) < (1ll << 24), "%s: tensor too large", who);
  const int per_row = wi * (c / vec);
  const dim3 grid((unsigned)(n * hi), (unsigned)((per_row + 255) / 256));
  CVX_DISPATCH_DTYPE(dtype, T, (upsample_bwd_kernel<T><<<grid, 256, 0, as_stream(stream)>>>(
                                   (const T*)dy, (T*)dx, hi, wi, ho, wo, c, lerp_scale(hi, ho), lerp_scale(wi, wo), lddy, dyoff)));
  CVX_LAUNCH_OK();
  return CVX_OK;
}

int cvx_upsample_fwd(const void* x, void* y, int n, int hi, int wi, int ho, int wo, int c, int dtype, void* stream) {
  return upsample_fwd_launch(x, y, n, hi, wi, ho, wo, c, c, 0, dtype, stream, "upsample_fwd");
}

int cvx_upsample_bwd(const void* dy, void* dx, int n, int hi, int wi, int ho, int wo, int c, int dtype, void* stream) {
  return upsample_bwd_launch(dy, dx, n, hi, wi, ho, wo, c, c, 0, dtype, stream, "upsample_bwd");
}

int cvx_upsample_into(const void* x, void* y, int n, int hi, int wi, int ho, int wo, int c, int c_total, int c_off, int dtype,
                      void* stream) {
  return upsample_fwd_launch(x, y, n, hi, wi, ho, wo, c, c_total, c_off, dtype, stream, "upsample_into");
}

int cvx_upsample_from_bwd(const void* dy, void* dx, int n, int hi, int wi, int ho, int wo, int c, int c_total, int c_off,
                          int dtype, void* stream) {
  return upsample_bwd_launch(dy, dx, n, hi, wi, ho, wo, c, c_total, c_off, dtype, stream, "upsample_from_bwd");
}

int cvx_upsample_to_nchw_fwd(const void* x, float* y, int n, int hi, int wi, int ho, int wo, int c, int dtype,
                             void* stream) {
  CVX_CHECK_ARG(x && y && n > 0 && hi > 0 && wi > 0 && ho > 0 && wo > 0 && c > 0, "upsample_to_nchw_fwd: bad arguments");
  const int64_t total = (int64_t)n * ho * wo;
  CVX_DISPATCH_DTYPE(dtype, T, (upsample_to_nchw_fwd_kernel<T><<<ew_grid(total), 256, 0, as_stream(stream)>>>(
                                   (const T*)x, y, n, hi, wi, ho, wo, c, lerp_scale(hi, ho), lerp_scale(wi, wo))));
  CVX_LAUNCH_OK();
  return CVX_OK;
}

int cvx_upsample_to_nchw_bwd(const float* dy, void* dx, int n, int hi, int wi, int ho, int wo, int c, int dtype,
                             void* stream) {
  CVX_CHECK_ARG(dy && dx && n > 0 && hi > 0 && wi > 0 && ho > 0 && wo > 0 && c > 0, "upsample_to_nchw_bwd: bad arguments");
  const int64_t total = (int64_t)n * hi * wi * c;
  // separable row form when the contributing output rows of an input row fit the weight table and a row fits in
  // shared memory (scale <= 1: up-sampling); the generic gather kernel otherwise
  const float shf = lerp_scale(hi, ho);
  const int max_rows = shf > 0.f ? (int)(2.f / shf) + 5 : ho;
  if (ho >= hi && max_rows <= kUpBwdMaxRows && wo <= 8192 && (int64_t)n * c * hi < (1ll << 31)) {
    CVX_DISPATCH_DTYPE(dtype, T, (upsample_to_nchw_bwd_rows_kernel<T><<<(unsigned)(n * c * hi), 128, sizeof(float) * wo,
                                                                     as_stream(stream)>>>(
                                     dy, (T*)dx, hi, wi, ho, wo, c, shf, lerp_scale(wi, wo))));
    CVX_LAUNCH_OK();
    return CVX_OK;
  }
  CVX_DISPATCH_DTYPE(dtype, T, (upsample_to_nchw_bwd_kernel<T><<<ew_grid(total), 256, 0, as_stream(stream)>>>(
                                   dy, (T*)dx, n, hi, wi, ho, wo, c, lerp_scale(hi, ho), lerp_scale(wi, wo))));
  CVX_LAUNCH_OK();
  return CVX_OK;
}

int cvx_maxpool3x3s2_fwd(const void* x, void* y, int n, int h, int w, int c, int dtype, void* stream) {
  CVX_CHECK_ARG(x && y && n > 0 && h > 0 && w > 0 && c > 0, "maxpool_fwd: bad arguments");
  const int vec = dtype == CVX_F32 ? 4 : 8;
  CVX_CHECK_ARG(c % vec == 0, "maxpool_fwd: C=%d not a multiple of %d", c, vec);
  const int ho = (h + 2 - 3) / 2 + 1, wo = (w + 2 - 3) / 2 + 1;
  const int64_t total = (int64_t)n * ho * wo * (c / vec);
  CVX_DISPATCH_DTYPE(dtype, T, (maxpool3x3s2_fwd_kernel<T><<<ew_grid(total), 256, 0, as_stream(stream)>>>(
                                   (const T*)x, (T*)y, n, h, w, c, ho, wo)));
  CVX_LAUNCH_OK();
  return CVX_OK;
}

int cvx_maxpool3x3s2_bwd(const void* x, const void* y, const void* dy, void* dx, int n, int h, int w, int c, int dtype,
                         void* stream) {
  CVX_CHECK_ARG(x && y && dy && dx && n > 0 && h > 0 && w > 0 && c > 0, "maxpool_bwd: bad arguments");
  const int vec = dtype == CVX_F32 ? 4 : 8;
  CVX_CHECK_ARG(c % vec == 0, "maxpool_bwd: C=%d not a multiple of %d", c, vec);
  const int ho = (h + 2 - 3) / 2 + 1, wo = (w + 2 - 3) / 2 + 1;
  const int64_t total = (int64_t)n * h * w * (c / vec);
  CVX_DISPATCH_DTYPE(dtype, T, (maxpool3x3s2_bwd_kernel<T><<<ew_grid(total), 256, 0, as_stream(stream)>>>(
                                   (const T*)x, (const T*)y, (const T*)dy, (T*)dx, n, h, w, c, ho, wo)));
  CVX_LAUNCH_OK();
  return CVX_OK;
}

static int dropout_launch(const void* x, void* y, uint8_t* mask, int64_t n, float p, uint64_t seed, const int* step_dev, int dtype,
                          void* stream, const char* who) {
  CVX_CHECK_ARG(x && y && n > 0 && p >= 0.f && p < 1.f, "%s: bad arguments", who);
  CVX_CHECK_ARG((uintptr_t)x % 16 == 0 && (uintptr_t)y % 16 == 0, "%s: pointers must be 16-byte aligned", who);
  const int vec = dtype == CVX_F32 ? 4 : 8;
  const int64_t work = n / vec > 256 ? n / vec : 256;
  CVX_DISPATCH_DTYPE(dtype, T, (launch_pdl(dropout_kernel<T>, dim3(ew_grid(work)), dim3(256), 0, as_stream(stream), (const T*)x, (T*)y, mask, n, p, seed, step_dev)));
  CVX_LAUNCH_OK();
  return CVX_OK;
}

int cvx_dropout_fwd(const void* x, void* y, uint8_t* mask, int64_t n, float p, uint64_t seed, const int* step_dev,
                    int dtype, void* stream) {
  return dropout_launch(x, y, mask, n, p, seed, step_dev, dtype, stream, "dropout_fwd");
}

int cvx_dropout_bwd_seeded(const void* dy, void* dx, int64_t n, float p, uint64_t seed, const int* step_dev, int dtype,
                           void* stream) {
  return dropout_launch(dy, dx, nullptr, n, p, seed, step_dev, dtype, stream, "dropout_bwd_seeded");
}

int cvx_dropout_bwd(const void* dy, const uint8_t* mask, void* dx, int64_t n, float p, int dtype, void* stream) {
  CVX_CHECK_ARG(dy && dx && mask && n > 0 && p >= 0.f && p < 1.f, "dropout_bwd: bad arguments");
  CVX_DISPATCH_DTYPE(dtype, T, (launch_pdl(dropout_bwd_kernel<T>, dim3(ew_grid(n)), dim3(256), 0, as_stream(stream), (const T*)dy, mask, (T*)dx, n, p)));
  CVX_LAUNCH_OK();
  return CVX_OK;
}

}  // extern "C"
