// Training augmentation of the segmentation loader on the device.
//
// Reference: DeeplabDataset.get_random_data, Segmentation/deeplabv3+/utils/dataloader.py:55-154 - per image on a CPU
// worker: PIL bicubic resize to a jittered size (label: nearest), optional flip, paste on a 128-grey canvas (label: 0),
// optional cv2.GaussianBlur 5x5, optional cv2.warpAffine rotation (bicubic, label nearest), HSV gain jitter through
// three 256-entry tables.  At ~750 images/s per GPU the four PIL/cv2 worker processes of train.py:281 cannot feed
// the step; here a batch of decoded uint8 images goes through the same arithmetic in four launches.
//
// Everything is 8-bit integer / fixed-point work and reproduces the libraries BIT FOR BIT (tests/test_augment.py against
// Pillow, OpenCV and the reference's own function):
//   * Pillow Resample.c: separable, horizontal pass first, 22-bit weights built by the host per (source, target) size
//     (utils/dataloader.py: resample_tables), uint8 rounding after each pass;
//   * Pillow Geometry.c nearest: source index tables from the host (running double sum);
//   * OpenCV GaussianBlur(5x5, sigma 0): taps (1,4,6,4,1)/16 per pass, exact in 8.8 fixed point, BORDER_REFLECT_101;
//   * OpenCV warpAffine: 10-bit fixed-point coordinates (host tables per angle), 5-bit phases, 15-bit 4x4 weights;
//   * OpenCV RGB2HSV (12-bit division tables) and HSV2RGB (float32; rows run in 32-pixel vectors with a fused
//     multiply-add and truncation, the ragged tail of a row in scalar code with rounding - both are reproduced).
// The kernels are HBM/L2-bound byte work (32 x 786 KB per batch): one thread per output pixel, a block per sample slice,
// per-sample parameters read from a descriptor array in device memory so that a batch is four launches whatever the
// source sizes are.
#include "common.cuh"

namespace cvx {

constexpr int kPilBits = 32 - 8 - 2;

__device__ __forceinline__ uint8_t clip8(int v) { return (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v)); }

// ---- pass 1: horizontal resample of every source row, tmp[ih][nw][3] (skipped by Pillow when the width is unchanged)
__global__ void __launch_bounds__(256) aug_resize_rows_kernel(const cvx_aug_sample* __restrict__ samples,
                                                              const uint8_t* __restrict__ src,
                                                              const int32_t* __restrict__ tables, uint8_t* __restrict__ tmp) {
  pdl_trigger();
  pdl_wait();
  const cvx_aug_sample s = samples[blockIdx.y];
  if (s.iw == s.nw) return;
  const int32_t* xmin = tables + s.xtab;
  const int32_t* xcnt = xmin + s.nw;
  const int32_t* xk = xcnt + s.nw;
  const uint8_t* in = src + s.src_off;
  uint8_t* out = tmp + s.tmp_off;
  const int64_t total = (int64_t)s.ih * s.nw;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int row = (int)(i / s.nw), xx = (int)(i % s.nw);
    const int x0 = xmin[xx], n = xcnt[xx];
    const int32_t* k = xk + (int64_t)xx * s.xtaps;
    const uint8_t* p = in + ((int64_t)row * s.iw + x0) * 3;
    int a0 = 1 << (kPilBits - 1), a1 = a0, a2 = a0;
    for (int t = 0; t < n; ++t) {
      const int w = k[t];
      a0 += p[3 * t] * w; a1 += p[3 * t + 1] * w; a2 += p[3 * t + 2] * w;
    }
    uint8_t* o = out + i * 3;
    o[0] = clip8(a0 >> kPilBits); o[1] = clip8(a1 >> kPilBits); o[2] = clip8(a2 >> kPilBits);
  }
}

// ---- pass 2: vertical resample + flip + paste on the canvas; label: nearest resize + flip + paste
__global__ void __launch_bounds__(256) aug_compose_kernel(const cvx_aug_sample* __restrict__ samples,
                                                          const uint8_t* __restrict__ src, const uint8_t* __restrict__ tmp,
                                                          const int32_t* __restrict__ tables, uint8_t* __restrict__ canvas,
                                                          uint8_t* __restrict__ labels, int H, int W) {
  pdl_trigger();
  pdl_wait();
  const cvx_aug_sample s = samples[blockIdx.y];
  const uint8_t* rows = s.iw == s.nw ? src + s.src_off : tmp + s.tmp_off;      // [ih][nw][3]
  const uint8_t* lab = src + s.lab_off;
  const int32_t* ymin = tables + s.ytab;
  const int32_t* ycnt = ymin + s.nh;
  const int32_t* yk = ycnt + s.nh;
  const int32_t* xnn = tables + s.xnn;
  const int32_t* ynn = tables + s.ynn;
  const bool vpass = s.ih != s.nh;
  uint8_t* cimg = canvas + (int64_t)blockIdx.y * H * W * 3;
  uint8_t* clab = labels + (int64_t)blockIdx.y * H * W;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < H * W; i += gridDim.x * blockDim.x) {
    const int y = i / W, x = i % W;
    const int ry = y - s.dy;
    int rx = x - s.dx;
    uint8_t r = 128, g = 128, b = 128, l = 0;
    if (ry >= 0 && ry < s.nh && rx >= 0 && rx < s.nw) {
      if (s.flip) rx = s.nw - 1 - rx;
      if (vpass) {
        const int y0 = ymin[ry], n = ycnt[ry];
        const int32_t* k = yk + (int64_t)ry * s.ytaps;
        const uint8_t* p = rows + ((int64_t)y0 * s.nw + rx) * 3;
        int a0 = 1 << (kPilBits - 1), a1 = a0, a2 = a0;
        for (int t = 0; t < n; ++t) {
          const int w = k[t];
          a0 += p[0] * w; a1 += p[1] * w; a2 += p[2] * w;
          p += (int64_t)s.nw * 3;
        }
        r = clip8(a0 >> kPilBits); g = clip8(a1 >> kPilBits); b = clip8(a2 >> kPilBits);
      } else {
        const uint8_t* p = rows + ((int64_t)ry * s.nw + rx) * 3;
        r = p[0]; g = p[1]; b = p[2];
      }
      l = lab[(int64_t)ynn[ry] * s.iw + xnn[rx]];
    }
    cimg[3 * i] = r; cimg[3 * i + 1] = g; cimg[3 * i + 2] = b;
    clab[i] = l;
  }
}

// ---- cv2.GaussianBlur(img, (5, 5), 0) for the samples that drew it (the others are read from `canvas` downstream)
__device__ __forceinline__ int reflect101(int i, int n) {
  if (i < 0) i = -i;
  if (i >= n) i = 2 * n - 2 - i;
  return i;
}

__global__ void __launch_bounds__(256) aug_blur5_kernel(const cvx_aug_sample* __restrict__ samples,
                                                        const uint8_t* __restrict__ canvas, uint8_t* __restrict__ out, int H,
                                                        int W) {
  pdl_trigger();
  pdl_wait();
  if (!samples[blockIdx.y].blur) return;
  const uint8_t* in = canvas + (int64_t)blockIdx.y * H * W * 3;
  uint8_t* o = out + (int64_t)blockIdx.y * H * W * 3;
  const int kw[5] = {1, 4, 6, 4, 1};
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < H * W; i += gridDim.x * blockDim.x) {
    const int y = i / W, x = i % W;
    int xs[5];
#pragma unroll
    for (int t = 0; t < 5; ++t) xs[t] = reflect101(x + t - 2, W) * 3;
    int a0 = 128, a1 = 128, a2 = 128;
#pragma unroll
    for (int u = 0; u < 5; ++u) {
      const uint8_t* row = in + (int64_t)reflect101(y + u - 2, H) * W * 3;
      int h0 = 0, h1 = 0, h2 = 0;
#pragma unroll
      for (int t = 0; t < 5; ++t) {
        h0 += kw[t] * row[xs[t]]; h1 += kw[t] * row[xs[t] + 1]; h2 += kw[t] * row[xs[t] + 2];
      }
      a0 += kw[u] * h0; a1 += kw[u] * h1; a2 += kw[u] * h2;
    }
    o[3 * i] = (uint8_t)(a0 >> 8); o[3 * i + 1] = (uint8_t)(a1 >> 8); o[3 * i + 2] = (uint8_t)(a2 >> 8);
  }
}

// ---- rotation (cv2.warpAffine, bicubic / nearest, constant border) + HSV gain jitter, fused (the jitter is pointwise)
constexpr int kAbBits = 10, kInterBits = 5, kCoefBits = 15, kHsvShift = 12;

__device__ __forceinline__ void rgb2hsv_u8(int r, int g, int b, const int* __restrict__ sdiv, const int* __restrict__ hdiv,
                                           int& h, int& s, int& v) {
  v = max(max(r, g), b);
  const int vmin = min(min(r, g), b);
  const int diff = v - vmin;
  s = (diff * sdiv[v] + (1 << (kHsvShift - 1))) >> kHsvShift;
  int hh = v == r ? g - b : (v == g ? b - r + 2 * diff : r - g + 4 * diff);
  hh = (hh * hdiv[diff] + (1 << (kHsvShift - 1))) >> kHsvShift;
  h = hh + (hh < 0 ? 180 : 0);
}

// OpenCV's float32 sector formula.  vec = the pixel lies in the part of the row OpenCV's AVX2 loop handles: fused
// multiply-add for 1 - s*h and truncation; otherwise separate multiply / subtract and round-to-nearest-even.
__device__ __forceinline__ void hsv2rgb_u8(int h, int s, int v, bool vec, uint8_t& r, uint8_t& g, uint8_t& b) {
  const float hf = __fmul_rn((float)h, (float)(6.0 / 180.0));
  const float sf = __fmul_rn((float)s, (float)(1.0 / 255.0));
  const float vf = __fmul_rn((float)v, (float)(1.0 / 255.0));
  const float sector = truncf(hf);
  const float fr = __fsub_rn(hf, sector);
  int sec = (int)sector;
  sec = sec >= 6 ? sec - 6 : sec;
  const float ifr = __fsub_rn(1.f, fr);
  const float a2 = vec ? __fmaf_rn(-sf, fr, 1.f) : __fsub_rn(1.f, __fmul_rn(sf, fr));
  const float a3 = vec ? __fmaf_rn(-sf, ifr, 1.f) : __fsub_rn(1.f, __fmul_rn(sf, ifr));
  const float t[4] = {vf, __fmul_rn(vf, __fsub_rn(1.f, sf)), __fmul_rn(vf, a2), __fmul_rn(vf, a3)};
  // sector -> table entries of (b, g, r)
  const int sel[6][3] = {{1, 3, 0}, {1, 0, 2}, {3, 0, 1}, {0, 2, 1}, {0, 1, 3}, {2, 1, 0}};
  const float fb = __fmul_rn(t[sel[sec][0]], 255.f), fg = __fmul_rn(t[sel[sec][1]], 255.f), fr8 = __fmul_rn(t[sel[sec][2]], 255.f);
  if (vec) {
    b = clip8((int)truncf(fb)); g = clip8((int)truncf(fg)); r = clip8((int)truncf(fr8));
  } else {
    b = clip8(__float2int_rn(fb)); g = clip8(__float2int_rn(fg)); r = clip8(__float2int_rn(fr8));
  }
}

__global__ void __launch_bounds__(256) aug_rotate_jitter_kernel(const cvx_aug_sample* __restrict__ samples,
                                                                const uint8_t* __restrict__ canvas,
                                                                const uint8_t* __restrict__ blurred,
                                                                const uint8_t* __restrict__ labels,
                                                                const int32_t* __restrict__ tables,
                                                                const int16_t* __restrict__ cubic,
                                                                const uint8_t* __restrict__ luts, uint8_t* __restrict__ out_img,
                                                                uint8_t* __restrict__ out_lab, int H, int W, int vec_cols) {
  __shared__ int sdiv[256], hdiv[256];
  __shared__ uint8_t lut[768];
  pdl_trigger();
  const cvx_aug_sample s = samples[blockIdx.y];
  // OpenCV's division tables: cvRound((255 << 12) / i), cvRound((180 << 12) / (6 i)) - IEEE double division, half to even
  for (int i = threadIdx.x; i < 256; i += blockDim.x) {
    sdiv[i] = i ? __double2int_rn((double)(255 << kHsvShift) / (double)i) : 0;
    hdiv[i] = i ? __double2int_rn((double)(180 << kHsvShift) / (6.0 * (double)i)) : 0;
  }
  pdl_wait();
  if (s.lut >= 0)
    for (int i = threadIdx.x; i < 768; i += blockDim.x) lut[i] = luts[s.lut + i];
  __syncthreads();
  const int64_t ioff = (int64_t)blockIdx.y * H * W * 3, loff = (int64_t)blockIdx.y * H * W;
  const uint8_t* in = (s.blur ? blurred : canvas) + ioff;
  const uint8_t* lin = labels + loff;
  const int32_t* adelta = tables + s.rot;
  const int32_t* bdelta = adelta + W;
  const int32_t* x0t = bdelta + W;
  const int32_t* y0t = x0t + H;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < H * W; i += gridDim.x * blockDim.x) {
    const int y = i / W, x = i % W;
    int r, g, b;
    uint8_t l;
    if (s.rotate) {
      // bicubic: phase rounding AB_SCALE / 32 / 2 = 16; nearest: AB_SCALE / 2 = 512
      const int X = (x0t[y] + 16 + adelta[x]) >> (kAbBits - kInterBits);
      const int Y = (y0t[y] + 16 + bdelta[x]) >> (kAbBits - kInterBits);
      const int sx = (X >> kInterBits) - 1, sy = (Y >> kInterBits) - 1;
      const int16_t* w = cubic + (((Y & 31) * 32 + (X & 31)) << 4);
      int a0 = 1 << (kCoefBits - 1), a1 = a0, a2 = a0;
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int yy = sy + u;
        const bool yok = yy >= 0 && yy < H;
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          const int xx = sx + t;
          const int wt = w[u * 4 + t];
          if (yok && xx >= 0 && xx < W) {
            const uint8_t* p = in + ((int64_t)yy * W + xx) * 3;
            a0 += p[0] * wt; a1 += p[1] * wt; a2 += p[2] * wt;
          } else {
            a0 += 128 * wt; a1 += 128 * wt; a2 += 128 * wt;
          }
        }
      }
      r = clip8(a0 >> kCoefBits); g = clip8(a1 >> kCoefBits); b = clip8(a2 >> kCoefBits);
      const int nx = (x0t[y] + 512 + adelta[x]) >> kAbBits, ny = (y0t[y] + 512 + bdelta[x]) >> kAbBits;
      l = (nx >= 0 && nx < W && ny >= 0 && ny < H) ? lin[ny * W + nx] : (uint8_t)0;
    } else {
      r = in[3 * i]; g = in[3 * i + 1]; b = in[3 * i + 2];
      l = lin[i];
    }
    uint8_t ro = (uint8_t)r, go = (uint8_t)g, bo = (uint8_t)b;
    if (s.lut >= 0) {
      int h, sa, v;
      rgb2hsv_u8(r, g, b, sdiv, hdiv, h, sa, v);
      hsv2rgb_u8(lut[h], lut[256 + sa], lut[512 + v], x < vec_cols, ro, go, bo);
    }
    out_img[ioff + 3 * i] = ro; out_img[ioff + 3 * i + 1] = go; out_img[ioff + 3 * i + 2] = bo;
    out_lab[loff + i] = l;
  }
}

static int aug_grid_x(int64_t elems) {
  int64_t b = (elems + 255) / 256;
  const int64_t cap = 8 * kNumSMs;        // per sample; a batch multiplies it by blockIdx.y
  return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

}  // namespace cvx

using namespace cvx;

extern "C" int cvx_aug_resize_rows(const cvx_aug_sample* samples, int batch, const unsigned char* src, const int* tables,
                                   unsigned char* tmp, int64_t max_elems, void* stream) {
  CVX_CHECK_ARG(samples && src && tables && batch > 0 && batch <= 65535, "aug_resize_rows: bad arguments");
  if (max_elems <= 0) return CVX_OK;      // no sample changes its width
  CVX_CHECK_ARG(tmp, "aug_resize_rows: scratch buffer missing");
  launch_pdl(aug_resize_rows_kernel, dim3(aug_grid_x(max_elems), batch), dim3(256), 0, as_stream(stream), samples, src, tables,
             tmp);
  CVX_LAUNCH_OK();
  return CVX_OK;
}

extern "C" int cvx_aug_compose(const cvx_aug_sample* samples, int batch, const unsigned char* src, const unsigned char* tmp,
                               const int* tables, unsigned char* canvas, unsigned char* labels, int h, int w, void* stream) {
  CVX_CHECK_ARG(samples && src && tables && canvas && labels && batch > 0 && batch <= 65535 && h > 0 && w > 0,
                "aug_compose: bad arguments");
  launch_pdl(aug_compose_kernel, dim3(aug_grid_x((int64_t)h * w), batch), dim3(256), 0, as_stream(stream), samples, src, tmp,
             tables, canvas, labels, h, w);
  CVX_LAUNCH_OK();
  return CVX_OK;
}

extern "C" int cvx_aug_blur5(const cvx_aug_sample* samples, int batch, const unsigned char* canvas, unsigned char* out, int h,
                             int w, void* stream) {
  CVX_CHECK_ARG(samples && canvas && out && batch > 0 && batch <= 65535 && h >= 3 && w >= 3, "aug_blur5: bad arguments");
  launch_pdl(aug_blur5_kernel, dim3(aug_grid_x((int64_t)h * w), batch), dim3(256), 0, as_stream(stream), samples, canvas, out,
             h, w);
  CVX_LAUNCH_OK();
  return CVX_OK;
}

extern "C" int cvx_aug_rotate_jitter(const cvx_aug_sample* samples, int batch, const unsigned char* canvas,
                                     const unsigned char* blurred, const unsigned char* labels, const int* tables,
                                     const short* cubic, const unsigned char* luts, unsigned char* out_img,
                                     unsigned char* out_lab, int h, int w, int vec_cols, void* stream) {
  CVX_CHECK_ARG(samples && canvas && blurred && labels && tables && cubic && luts && out_img && out_lab && batch > 0 &&
                    batch <= 65535 && h > 0 && w > 0,
                "aug_rotate_jitter: bad arguments");
  launch_pdl(aug_rotate_jitter_kernel, dim3(aug_grid_x((int64_t)h * w), batch), dim3(256), 0, as_stream(stream), samples,
             canvas, blurred, labels, tables, cubic, luts, out_img, out_lab, h, w, vec_cols);
  CVX_LAUNCH_OK();
  return CVX_OK;
}
