// Input-side kernels: everything between the decoded image and the first
// convolution, as ONE pass each, writing the NHWC tensor the stem convolution reads.
//
//  * split_patches_kernel - the classifier's patch pipeline (reference: MM/Graph_Structure(data_augmentation).py
//    :145-161): bilinear resize of each colposcopic image to new_size^2, cut into (new_size/patch)^2 patches
//    enumerated x-major, ToTensor + Normalize(mean, std).  The reference does this with PIL one image at a time and
//    feeds the encoder one patch per call; here the fp32 NCHW image batch goes straight to [B*k*k, patch, patch, C]
//    NHWC in the encoder's compute type (no 1024^2 intermediate, no separate normalise / permute / layout passes).
//  * finish_batch_u8_kernel - the segmentation loader's tail (reference: SEG/utils/dataloader.py:40-42,
//    utils/utils.py:63-65 preprocess_input): uint8 HWC pixels -> /255 -> NHWC activation, and the uint8 class map ->
//    int64 with values >= num_classes clamped to num_classes (the ignore label).  Shipping the uint8 arrays and
//    finishing them here cuts the per-step host->device traffic 5x (4 B instead of 20 B per pixel); the one-hot
//    labels the reference builds on the host (:47) are derived inside the loss kernel.
#include "common.cuh"

namespace cvx {

// torch / PIL half-pixel bilinear coordinate (align_corners=False): src = (dst + 0.5) * scale - 0.5, clamped at 0;
// the second tap is clamped to the last row / column.
__device__ __forceinline__ void half_pixel_coord(int d, float scale, int src, int& i0, int& i1, float& f) {
  float s = ((float)d + 0.5f) * scale - 0.5f;
  if (s < 0.f) s = 0.f;
  int i = (int)s;
  if (i > src - 1) i = src - 1;
  i0 = i;
  i1 = i + 1 < src ? i + 1 : src - 1;
  f = s - (float)i;
}

constexpr int kMaxImgChannels = 4;

struct PatchNorm {
  float mean[kMaxImgChannels];
  float inv_std[kMaxImgChannels];
};

// One thread = kPx horizontally adjacent output pixels x all channels (kPx * C contiguous elements of the patch
// row).  Threads of a warp cover 32*kPx adjacent pixels of one patch row, so both the plane reads (<= 2 source rows
// per channel, contiguous) and the interleaved stores are coalesced.
template <typename T, int C, int kPx>
__global__ void __launch_bounds__(256) split_patches_kernel(const float* __restrict__ img, T* __restrict__ out, int h,
                                                            int w, int new_size, int patch, float sy, float sx,
                                                            PatchNorm nrm, int64_t total_groups) {
  const int k = new_size / patch;
  const int groups_per_row = patch / kPx;
  for (int64_t gidx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; gidx < total_groups;
       gidx += (int64_t)gridDim.x * blockDim.x) {
    // gidx enumerates [image][patch = ix*k + iy][py][px group]
    int64_t t = gidx;
    const int gx = (int)(t % groups_per_row); t /= groups_per_row;
    const int py = (int)(t % patch); t /= patch;
    const int pidx = (int)(t % (k * k)); t /= (k * k);
    const int64_t n = t;
    const int ix = pidx / k, iy = pidx % k;          // x-major enumeration (outer loop over columns, :157-160)
    const int oy = iy * patch + py;
    int y0, y1;
    float fy;
    half_pixel_coord(oy, sy, h, y0, y1, fy);
    const float* base = img + n * (int64_t)C * h * w;
    T* dst = out + (((n * (k * k) + pidx) * patch + py) * (int64_t)patch + gx * kPx) * C;
#pragma unroll
    for (int j = 0; j < kPx; ++j) {
      const int ox = ix * patch + gx * kPx + j;
      int x0, x1;
      float fx;
      half_pixel_coord(ox, sx, w, x0, x1, fx);
#pragma unroll
      for (int c = 0; c < C; ++c) {
        const float* pl = base + (int64_t)c * h * w;
        const float v00 = __ldg(pl + (int64_t)y0 * w + x0), v01 = __ldg(pl + (int64_t)y0 * w + x1);
        const float v10 = __ldg(pl + (int64_t)y1 * w + x0), v11 = __ldg(pl + (int64_t)y1 * w + x1);
        // torch's upsample_bilinear2d order: weights (1-fy)*((1-fx) a + fx b) + fy*((1-fx) c + fx d)
        const float top = (1.f - fx) * v00 + fx * v01, bot = (1.f - fx) * v10 + fx * v11;
        const float v = (1.f - fy) * top + fy * bot;
        Elem<T>::st(dst + j * C + c, (v - nrm.mean[c]) * nrm.inv_std[c]);
      }
    }
  }
}

// ---- Pillow-exact variant on the decoded uint8 image ------------------------------------------------------------------
// PIL.Image.resize(..., BILINEAR) on 8-bit images (Pillow Resample.c) is a separable fixed-point filter: horizontal pass
// first, rounded and clipped to uint8, then the vertical pass on those bytes; weights carry 22 fractional bits, the
// accumulator starts at one half; the triangle widens when the image is reduced (antialiasing) and is renormalised at the
// borders.  The per-axis tables (first tap, tap count, integer weights) come from the host (multimodal/pil_resample.py,
// checked bit for bit against the installed Pillow); here one thread produces one output pixel: for each of its source
// rows the horizontal sum -> uint8, then the vertical sum -> uint8, then ToTensor (/255) and Normalize ((v - mean) / std)
// with IEEE divisions, so an fp32 patch equals transforms.Normalize(ToTensor(patch)) of the reference bit for bit.
constexpr int kPilBits = 22;

struct PatchNormDiv {
  float mean[kMaxImgChannels];
  float std[kMaxImgChannels];
};

template <typename T, int C>
__global__ void __launch_bounds__(256) split_patches_u8_kernel(const uint8_t* __restrict__ img, T* __restrict__ out, int h,
                                                               int w, int new_size, int patch,
                                                               const int* __restrict__ xmin, const int* __restrict__ xcnt,
                                                               const int* __restrict__ xk, int xks,
                                                               const int* __restrict__ ymin, const int* __restrict__ ycnt,
                                                               const int* __restrict__ yk, int yks, PatchNormDiv nrm,
                                                               int64_t total) {
  const int k = new_size / patch;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    // idx enumerates [image][patch = ix*k + iy][py][px]: consecutive threads write consecutive pixels of a patch row
    int64_t t = idx;
    const int px = (int)(t % patch); t /= patch;
    const int py = (int)(t % patch); t /= patch;
    const int pidx = (int)(t % (k * k)); t /= (k * k);
    const int64_t n = t;
    const int ix = pidx / k, iy = pidx % k;          // x-major enumeration (outer loop over columns, :157-160)
    const int oy = iy * patch + py, ox = ix * patch + px;
    const int x0 = xmin[ox], nx = xcnt[ox], y0 = ymin[oy], ny = ycnt[oy];
    const int* kx = xk + (int64_t)ox * xks;
    const int* ky = yk + (int64_t)oy * yks;
    const uint8_t* base = img + n * (int64_t)h * w * C;
    int acc[C];
#pragma unroll
    for (int c = 0; c < C; ++c) acc[c] = 1 << (kPilBits - 1);
    for (int r = 0; r < ny; ++r) {
      const uint8_t* row = base + ((int64_t)(y0 + r) * w + x0) * C;
      int hs[C];
#pragma unroll
      for (int c = 0; c < C; ++c) hs[c] = 1 << (kPilBits - 1);
      for (int x = 0; x < nx; ++x) {
        const int wgt = __ldg(kx + x);
#pragma unroll
        for (int c = 0; c < C; ++c) hs[c] += (int)__ldg(row + x * C + c) * wgt;
      }
      const int wy = __ldg(ky + r);
#pragma unroll
      for (int c = 0; c < C; ++c) {
        int b = hs[c] >> kPilBits;                   // the horizontal pass's uint8 result (clip8)
        b = b < 0 ? 0 : (b > 255 ? 255 : b);
        acc[c] += b * wy;
      }
    }
    T* dst = out + idx * C;
#pragma unroll
    for (int c = 0; c < C; ++c) {
      int b = acc[c] >> kPilBits;
      b = b < 0 ? 0 : (b > 255 ? 255 : b);
      const float v = __fdiv_rn(__fsub_rn(__fdiv_rn((float)b, 255.f), nrm.mean[c]), nrm.std[c]);
      Elem<T>::st(dst + c, v);
    }
  }
}

// uint8 [n][h][w][c] -> T [n][h][w][c] * (1/255): 16 input bytes per thread.
template <typename T>
__global__ void __launch_bounds__(256) u8_to_unit_kernel(const uint8_t* __restrict__ src, T* __restrict__ dst,
                                                         int64_t nvec, int64_t n) {
  const float k = 1.f / 255.f;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < nvec; i += (int64_t)gridDim.x * blockDim.x) {
    const uint4 q = __ldg(reinterpret_cast<const uint4*>(src) + i);
    const uint32_t u[4] = {q.x, q.y, q.z, q.w};
    T* d = dst + i * 16;
#pragma unroll
    for (int g = 0; g < 16 / Vec<T>::N; ++g) {      // 16-byte stores (the destination is 16-byte aligned)
      Vec<T> o;
#pragma unroll
      for (int e = 0; e < Vec<T>::N; ++e) {
        const int j = g * Vec<T>::N + e;
        o.v[e] = (float)((u[j >> 2] >> (8 * (j & 3))) & 0xffu) * k;
      }
      o.store(d + g * Vec<T>::N);
    }
  }
  // tail (n not a multiple of 16)
  for (int64_t i = nvec * 16 + blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    Elem<T>::st(dst + i, (float)src[i] * k);
}

// uint8 class map -> int64, values >= num_classes become num_classes (dataloader.py:42)
__global__ void __launch_bounds__(256) u8_labels_kernel(const uint8_t* __restrict__ src, int64_t* __restrict__ dst,
                                                        int64_t n, int num_classes) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int v = src[i];
    dst[i] = v >= num_classes ? num_classes : v;
  }
}

static inline unsigned stream_grid(int64_t items, int threads) {
  int64_t b = ceil_div64(items, threads);
  const int64_t cap = (int64_t)kNumSMs * 8;
  if (b > cap) b = cap;
  return (unsigned)(b < 1 ? 1 : b);
}

}  // namespace cvx

using namespace cvx;

extern "C" int cvx_split_patches(const float* images, void* patches, int n, int c, int h, int w, int new_size, int patch,
                                 const float* mean, const float* std, int dtype, void* stream) {
  CVX_CHECK_ARG(images && patches && mean && std, "split_patches: null pointer");
  CVX_CHECK_ARG(n > 0 && h > 0 && w > 0, "split_patches: bad image batch %dx%dx%d", n, h, w);
  CVX_CHECK_ARG(c == 3, "split_patches: 3-channel images only (got %d)", c);
  CVX_CHECK_ARG(patch > 0 && new_size >= patch && new_size % patch == 0 && patch % 4 == 0,
                "split_patches: new_size %d must be a multiple of the patch size %d (itself a multiple of 4)", new_size, patch);
  PatchNorm nrm;
  for (int i = 0; i < kMaxImgChannels; ++i) { nrm.mean[i] = 0.f; nrm.inv_std[i] = 1.f; }
  for (int i = 0; i < c; ++i) {
    CVX_CHECK_ARG(std[i] != 0.f, "split_patches: std[%d] is zero", i);
    nrm.mean[i] = mean[i];
    nrm.inv_std[i] = 1.f / std[i];
  }
  // torch computes the scale in float as (float)in / out for align_corners=False
  const float sy = (float)h / (float)new_size, sx = (float)w / (float)new_size;
  const int64_t groups = (int64_t)n * new_size * (new_size / 4);
  const unsigned grid = stream_grid(groups, 256);
  CVX_DISPATCH_DTYPE(dtype, T, (split_patches_kernel<T, 3, 4><<<grid, 256, 0, as_stream(stream)>>>(
                                   images, (T*)patches, h, w, new_size, patch, sy, sx, nrm, groups)));
  CVX_LAUNCH_OK();
  return CVX_OK;
}

extern "C" int cvx_split_patches_u8(const unsigned char* images, void* patches, int n, int h, int w, int c, int new_size,
                                    int patch, const int* xmin, const int* xcnt, const int* xk, int xksize, const int* ymin,
                                    const int* ycnt, const int* yk, int yksize, const float* mean, const float* std,
                                    int dtype, void* stream) {
  CVX_CHECK_ARG(images && patches && xmin && xcnt && xk && ymin && ycnt && yk && mean && std, "split_patches_u8: null pointer");
  CVX_CHECK_ARG(n > 0 && h > 0 && w > 0, "split_patches_u8: bad image batch %dx%dx%d", n, h, w);
  CVX_CHECK_ARG(c == 3, "split_patches_u8: 3-channel images only (got %d)", c);
  CVX_CHECK_ARG(patch > 0 && new_size >= patch && new_size % patch == 0 && xksize > 0 && yksize > 0,
                "split_patches_u8: new_size %d must be a multiple of the patch size %d", new_size, patch);
  PatchNormDiv nrm;
  for (int i = 0; i < kMaxImgChannels; ++i) { nrm.mean[i] = 0.f; nrm.std[i] = 1.f; }
  for (int i = 0; i < c; ++i) {
    CVX_CHECK_ARG(std[i] != 0.f, "split_patches_u8: std[%d] is zero", i);
    nrm.mean[i] = mean[i];
    nrm.std[i] = std[i];
  }
  const int64_t total = (int64_t)n * new_size * new_size;
  const unsigned grid = stream_grid(total, 256);
  CVX_DISPATCH_DTYPE(dtype, T, (split_patches_u8_kernel<T, 3><<<grid, 256, 0, as_stream(stream)>>>(
                                   images, (T*)patches, h, w, new_size, patch, xmin, xcnt, xk, xksize, ymin, ycnt, yk, yksize,
                                   nrm, total)));
  CVX_LAUNCH_OK();
  return CVX_OK;
}

extern "C" int cvx_finish_batch_u8(const unsigned char* images_u8, void* images_out, int64_t n_image_elems,
                                   const unsigned char* labels_u8, int64_t* labels_out, int64_t n_pixels,
                                   int num_classes, int dtype, void* stream) {
  CVX_CHECK_ARG((images_u8 == nullptr) == (images_out == nullptr), "finish_batch_u8: images in/out must both be given or both be NULL");
  CVX_CHECK_ARG((labels_u8 == nullptr) == (labels_out == nullptr), "finish_batch_u8: labels in/out must both be given or both be NULL");
  CVX_CHECK_ARG(images_u8 || labels_u8, "finish_batch_u8: nothing to do");
  if (images_u8) {
    CVX_CHECK_ARG(n_image_elems > 0, "finish_batch_u8: bad image element count");
    CVX_CHECK_ARG(((uintptr_t)images_u8 & 15) == 0 && ((uintptr_t)images_out & 15) == 0,
                  "finish_batch_u8: image buffers must be 16-byte aligned");
    const int64_t nvec = n_image_elems / 16;
    const unsigned grid = stream_grid(nvec > 0 ? nvec : n_image_elems, 256);
    CVX_DISPATCH_DTYPE(dtype, T, (u8_to_unit_kernel<T><<<grid, 256, 0, as_stream(stream)>>>(images_u8, (T*)images_out, nvec,
                                                                                           n_image_elems)));
    CVX_LAUNCH_OK();
  }
  if (labels_u8) {
    CVX_CHECK_ARG(n_pixels > 0 && num_classes >= 1 && num_classes <= 255, "finish_batch_u8: bad label arguments");
    u8_labels_kernel<<<stream_grid(n_pixels, 256), 256, 0, as_stream(stream)>>>(labels_u8, labels_out, n_pixels, num_classes);
    CVX_LAUNCH_OK();
  }
  return CVX_OK;
}
