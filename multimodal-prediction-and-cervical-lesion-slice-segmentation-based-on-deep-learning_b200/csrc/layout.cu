// ABI bookkeeping + layout plumbing kernels (NCHW<->NHWC, weight packing, channel slices).
#include <stdarg.h>
#include <string.h>

#include "common.cuh"
#include <stdlib.h>

namespace cvx {

static thread_local char g_err[512] = "";
unsigned long long g_kernel_launches = 0;
int g_ws_prezeroed = 0;
int g_pdl = [] { const char* e = getenv("CERVIX_PDL"); return e ? atoi(e) : 1; }();

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

// ---- [C][HW] <-> [HW][C] tile transposes, one image per blockIdx.z ----------------------
template <typename T>
__global__ void nchw_to_nhwc_kernel(const float* __restrict__ src, T* __restrict__ dst, int c, int hw) {
  __shared__ float tile[32][33];
  const int n = blockIdx.z;
  const int p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const float* s = src + (size_t)n * c * hw;
  T* d = dst + (size_t)n * c * hw;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    int cc = c0 + i, p = p0 + threadIdx.x;
    tile[i][threadIdx.x] = (cc < c && p < hw) ? s[(size_t)cc * hw + p] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    int p = p0 + i, cc = c0 + threadIdx.x;
    if (p < hw && cc < c) Elem<T>::st(d + (size_t)p * c + cc, tile[threadIdx.x][i]);
  }
}

// C <= 4 (the RGB image batch): one thread per pixel reads its C planes (coalesced) and writes C adjacent elements;
// the 32x32 tile transpose above would keep 3 of 32 lanes busy on the store side.
template <typename T>
__global__ void __launch_bounds__(256) nchw_to_nhwc_smallc_kernel(const float* __restrict__ src, T* __restrict__ dst,
                                                                  int c, int hw) {
  const float* s = src + (size_t)blockIdx.y * c * hw;
  T* d = dst + (size_t)blockIdx.y * c * hw;
  for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < hw; p += gridDim.x * blockDim.x) {
    float v[4];
#pragma unroll
    for (int cc = 0; cc < 4; ++cc) v[cc] = cc < c ? __ldg(s + (size_t)cc * hw + p) : 0.f;
#pragma unroll
    for (int cc = 0; cc < 4; ++cc)
      if (cc < c) Elem<T>::st(d + (size_t)p * c + cc, v[cc]);
  }
}

template <typename T>
__global__ void nhwc_to_nchw_kernel(const T* __restrict__ src, float* __restrict__ dst, int c, int hw) {
  __shared__ float tile[32][33];
  const int n = blockIdx.z;
  const int p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const T* s = src + (size_t)n * c * hw;
  float* d = dst + (size_t)n * c * hw;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    int p = p0 + i, cc = c0 + threadIdx.x;
    tile[i][threadIdx.x] = (p < hw && cc < c) ? Elem<T>::ld(s + (size_t)p * c + cc) : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    int cc = c0 + i, p = p0 + threadIdx.x;
    if (cc < c && p < hw) d[(size_t)cc * hw + p] = tile[threadIdx.x][i];
  }
}

// OIHW fp32 -> [tap][cout][cin] (or [flipped tap][cin][cout]) in T
template <typename T>
__global__ void pack_weight_kernel(const float* __restrict__ w, T* __restrict__ dst, int cout, int cin,
                                   int taps, int transpose_flip) {
  const int64_t total = (int64_t)taps * cout * cin;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    int tap, co, ci;
    if (!transpose_flip) {
      ci = (int)(i % cin); co = (int)((i / cin) % cout); tap = (int)(i / ((int64_t)cin * cout));
    } else {
      co = (int)(i % cout); ci = (int)((i / cout) % cin);
      tap = taps - 1 - (int)(i / ((int64_t)cin * cout));
    }
    Elem<T>::st(dst + i, w[((int64_t)co * cin + ci) * taps + tap]);
  }
}

__global__ void unpack_wgrad_kernel(const float* __restrict__ g, float* __restrict__ out, int cout, int cin,
                                    int taps) {
  const int64_t total = (int64_t)taps * cout * cin;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    int tap = (int)(i % taps);
    int ci = (int)((i / taps) % cin);
    int co = (int)(i / ((int64_t)taps * cin));
    out[i] = g[((int64_t)tap * cout + co) * cin + ci];
  }
}

__global__ void transpose_small_kernel(const float* __restrict__ src, float* __restrict__ dst, int rows, int cols) {
  // dst[c][r] = src[r][c]
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < rows * cols) {
    int r = i / cols, c = i % cols;
    dst[(size_t)c * rows + r] = src[i];
  }
}

template <typename T>
__global__ void copy_channels_kernel(const T* __restrict__ src, int src_ld, int src_coff, T* __restrict__ dst,
                                     int dst_ld, int dst_coff, int64_t rows, int c) {
  const int64_t total = rows * c;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    int64_t r = i / c;
    int cc = (int)(i - r * c);
    dst[r * dst_ld + dst_coff + cc] = src[r * src_ld + src_coff + cc];
  }
}

template <typename T>
__global__ void copy_channels_vec_kernel(const T* __restrict__ src, int src_ld, int src_coff, T* __restrict__ dst,
                                         int dst_ld, int dst_coff, int64_t rows, int c) {
  constexpr int V = Elem<T>::kVec;
  const int cv = c / V;
  const int64_t total = rows * cv;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    int64_t r = i / cv;
    int cc = (int)(i - r * cv) * V;
    *reinterpret_cast<uint4*>(dst + r * dst_ld + dst_coff + cc) =
        *reinterpret_cast<const uint4*>(src + r * src_ld + src_coff + cc);
  }
}

static inline int grid_for(int64_t total, int block, int max_blocks = kNumSMs * 16) {
  int64_t g = ceil_div64(total, block);
  if (g < 1) g = 1;
  return (int)(g > max_blocks ? max_blocks : g);
}

}  // namespace cvx

using namespace cvx;

extern "C" {

int cvx_abi_version(void) { return CVX_ABI_VERSION; }
const char* cvx_last_error(void) { return cvx::g_err; }
int64_t cvx_launch_count(void) { return (int64_t)cvx::g_kernel_launches; }
int cvx_set_ws_prezeroed(int on) {
  cvx::g_ws_prezeroed = on ? 1 : 0;
  return CVX_OK;
}

int cvx_device_is_sm100(void) {
  int dev = 0, major = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 0;
  if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) return 0;
  return major == 10 ? 1 : 0;
}

int cvx_nchw_to_nhwc(const float* src, void* dst, int n, int c, int h, int w, int dtype, void* stream) {
  CVX_CHECK_ARG(src && dst && n > 0 && c > 0 && h > 0 && w > 0, "nchw_to_nhwc: bad arguments");
  const int hw = h * w;
  if (c <= 4 && n <= 65535) {
    int bx = (hw + 1023) / 1024;
    const int cap = (kNumSMs * 8 + n - 1) / n;
    if (bx > cap) bx = cap;
    CVX_DISPATCH_DTYPE(dtype, T, (nchw_to_nhwc_smallc_kernel<T><<<dim3(bx, n), 256, 0, as_stream(stream)>>>(src, (T*)dst, c, hw)));
    CVX_LAUNCH_OK();
    return CVX_OK;
  }
  dim3 grid((hw + 31) / 32, (c + 31) / 32, n), block(32, 8);
  CVX_CHECK_ARG(grid.y <= 65535 && grid.z <= 65535, "nchw_to_nhwc: grid too large");
  CVX_DISPATCH_DTYPE(dtype, T, (nchw_to_nhwc_kernel<T><<<grid, block, 0, as_stream(stream)>>>(src, (T*)dst, c, hw)));
  CVX_LAUNCH_OK();
  return CVX_OK;
}

int cvx_nhwc_to_nchw(const void* src, float* dst, int n, int c, int h, int w, int dtype, void* stream) {
  CVX_CHECK_ARG(src && dst && n > 0 && c > 0 && h > 0 && w > 0, "nhwc_to_nchw: bad arguments");
  const int hw = h * w;
  dim3 grid((hw + 31) / 32, (c + 31) / 32, n), block(32, 8);
  CVX_CHECK_ARG(grid.y <= 65535 && grid.z <= 65535, "nhwc_to_nchw: grid too large");
  CVX_DISPATCH_DTYPE(dtype, T, (nhwc_to_nchw_kernel<T><<<grid, block, 0, as_stream(stream)>>>((const T*)src, dst, c, hw)));
  CVX_LAUNCH_OK();
  return CVX_OK;
}

int cvx_pack_weight(const float* w_oihw, void* dst, int cout, int cin, int kh, int kw, int dtype,
                    int transpose_flip, void* stream) {
  CVX_CHECK_ARG(w_oihw && dst && cout > 0 && cin > 0 && kh > 0 && kw > 0, "pack_weight: bad arguments");
  const int64_t total = (int64_t)kh * kw * cout * cin;
  CVX_DISPATCH_DTYPE(dtype, T, (pack_weight_kernel<T><<<grid_for(total, 256), 256, 0, as_stream(stream)>>>(
                                   w_oihw, (T*)dst, cout, cin, kh * kw, transpose_flip)));
  CVX_LAUNCH_OK();
  return CVX_OK;
}

int cvx_unpack_wgrad(const float* g_packed, float* g_oihw, int cout, int cin, int kh, int kw, void* stream) {
  CVX_CHECK_ARG(g_packed && g_oihw && cout > 0 && cin > 0, "unpack_wgrad: bad arguments");
  const int64_t total = (int64_t)kh * kw * cout * cin;
  unpack_wgrad_kernel<<<grid_for(total, 256), 256, 0, as_stream(stream)>>>(g_packed, g_oihw, cout, cin, kh * kw);
  CVX_LAUNCH_OK();
  return CVX_OK;
}

int cvx_pack_dw_weight(const float* w_c133, float* dst, int c, void* stream) {
  CVX_CHECK_ARG(w_c133 && dst && c > 0, "pack_dw_weight: bad arguments");
  transpose_small_kernel<<<(c * 9 + 255) / 256, 256, 0, as_stream(stream)>>>(w_c133, dst, c, 9);
  CVX_LAUNCH_OK();
  return CVX_OK;
}

int cvx_unpack_dw_wgrad(const float* g_9c, float* g_c133, int c, void* stream) {
  CVX_CHECK_ARG(g_9c && g_c133 && c > 0, "unpack_dw_wgrad: bad arguments");
  transpose_small_kernel<<<(c * 9 + 255) / 256, 256, 0, as_stream(stream)>>>(g_9c, g_c133, 9, c);
  CVX_LAUNCH_OK();
  return CVX_OK;
}

int cvx_copy_channels(const void* src, int src_ld, int src_coff, void* dst, int dst_ld, int dst_coff,
                      int64_t rows, int c, int dtype, void* stream) {
  CVX_CHECK_ARG(src && dst && rows > 0 && c > 0 && src_coff >= 0 && dst_coff >= 0 && src_coff + c <= src_ld &&
                    dst_coff + c <= dst_ld,
                "copy_channels: bad arguments");
  const int vec = dtype == CVX_F32 ? 4 : 8;
  const bool aligned = (c % vec == 0) && (src_ld % vec == 0) && (dst_ld % vec == 0) && (src_coff % vec == 0) &&
                       (dst_coff % vec == 0) && ((uintptr_t)src % 16 == 0) && ((uintptr_t)dst % 16 == 0);
  if (aligned) {
    CVX_DISPATCH_DTYPE(dtype, T, (copy_channels_vec_kernel<T><<<grid_for(rows * (c / vec), 256), 256, 0, as_stream(stream)>>>(
                                     (const T*)src, src_ld, src_coff, (T*)dst, dst_ld, dst_coff, rows, c)));
  } else {
    CVX_DISPATCH_DTYPE(dtype, T, (copy_channels_kernel<T><<<grid_for(rows * c, 256), 256, 0, as_stream(stream)>>>(
                                     (const T*)src, src_ld, src_coff, (T*)dst, dst_ld, dst_coff, rows, c)));
  }
  CVX_LAUNCH_OK();
  return CVX_OK;
}

}  // extern "C"
