// Inference post-processing of the segmentation predictor (reference: deeplab.py:141-154 detect_image, :304-345
// get_miou_png): softmax over classes -> crop of the letterbox -> bilinear resize to the original image size ->
// argmax, as ONE kernel on the logits, so only the uint8 class map crosses PCIe (the reference copies the whole
// [H,W,C] fp32 probability tensor to the host and resizes it with cv2).
//
// The resize follows cv2.resize(..., INTER_LINEAR) on float data: pixel centres at (d + 0.5) * scale - 0.5, left /
// top taps clamped to the image with zero weight on the outside, horizontal pass first, all in fp32.
#include "common.cuh"

namespace cvx {

constexpr int kMaxClasses = 32;

__device__ __forceinline__ void cv_linear_coord(int d, float scale, int src, int& s0, int& s1, float& f) {
  float fx = ((float)d + 0.5f) * scale - 0.5f;
  int sx = (int)floorf(fx);
  fx -= (float)sx;
  if (sx < 0) { sx = 0; fx = 0.f; }
  if (sx >= src - 1) { sx = src - 1; fx = 0.f; }
  s0 = sx;
  s1 = sx + 1 < src ? sx + 1 : src - 1;
  f = fx;
}

__global__ void seg_postprocess_kernel(const float* __restrict__ logits, int c, int h, int w, int crop_y, int crop_x,
                                       int crop_h, int crop_w, int out_h, int out_w, float sy, float sx,
                                       uint8_t* __restrict__ cls, float* __restrict__ probs) {
  const int ox = blockIdx.x * blockDim.x + threadIdx.x;
  const int oy = blockIdx.y;
  if (ox >= out_w) return;
  int y0, y1, x0, x1;
  float fy, fx;
  cv_linear_coord(oy, sy, crop_h, y0, y1, fy);
  cv_linear_coord(ox, sx, crop_w, x0, x1, fx);
  const size_t plane = (size_t)h * w;
  const float* base = logits + (size_t)crop_y * w + crop_x;
  const size_t o00 = (size_t)y0 * w + x0, o01 = (size_t)y0 * w + x1, o10 = (size_t)y1 * w + x0, o11 = (size_t)y1 * w + x1;
  // softmax at the four taps (max-subtracted, as torch's)
  float m00 = -INFINITY, m01 = -INFINITY, m10 = -INFINITY, m11 = -INFINITY;
  for (int k = 0; k < c; ++k) {
    const float* pl = base + k * plane;
    m00 = fmaxf(m00, __ldg(pl + o00)); m01 = fmaxf(m01, __ldg(pl + o01));
    m10 = fmaxf(m10, __ldg(pl + o10)); m11 = fmaxf(m11, __ldg(pl + o11));
  }
  float e00[kMaxClasses], e01[kMaxClasses], e10[kMaxClasses], e11[kMaxClasses];
  float s00 = 0.f, s01 = 0.f, s10 = 0.f, s11 = 0.f;
#pragma unroll 1
  for (int k = 0; k < c; ++k) {
    const float* pl = base + k * plane;
    e00[k] = expf(__ldg(pl + o00) - m00); s00 += e00[k];
    e01[k] = expf(__ldg(pl + o01) - m01); s01 += e01[k];
    e10[k] = expf(__ldg(pl + o10) - m10); s10 += e10[k];
    e11[k] = expf(__ldg(pl + o11) - m11); s11 += e11[k];
  }
  int best = 0;
  float best_v = -INFINITY;
#pragma unroll 1
  for (int k = 0; k < c; ++k) {
    const float top = (e00[k] / s00) * (1.f - fx) + (e01[k] / s01) * fx;   // horizontal pass of both rows
    const float bot = (e10[k] / s10) * (1.f - fx) + (e11[k] / s11) * fx;
    const float v = top * (1.f - fy) + bot * fy;
    if (probs) probs[((size_t)oy * out_w + ox) * c + k] = v;
    if (v > best_v) { best_v = v; best = k; }   // first maximum, as numpy.argmax
  }
  cls[(size_t)oy * out_w + ox] = (uint8_t)best;
}

// Confusion matrix of a predicted class map against the ground truth (reference: utils_metrics.py:37-47 fast_hist =
// np.bincount(n * gt[k] + pred[k]) over the pixels k with 0 <= gt < n): per-block shared-memory histogram, one global
// atomic per non-empty cell and block.  Accumulates into hist, so a whole validation set is one [n][n] matrix on the
// device (the reference round-trips every prediction through a PNG file, callbacks.py:163-176).
constexpr int kMaxHistClasses = 32;

__global__ void __launch_bounds__(256) confusion_matrix_kernel(const uint8_t* __restrict__ pred,
                                                               const uint8_t* __restrict__ gt, int64_t n, int classes,
                                                               unsigned long long* __restrict__ hist) {
  __shared__ unsigned int sm[kMaxHistClasses * kMaxHistClasses];
  const int cells = classes * classes;
  for (int i = threadIdx.x; i < cells; i += blockDim.x) sm[i] = 0u;
  __syncthreads();
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int a = gt[i], b = pred[i];
    if (a < classes && b < classes) atomicAdd(&sm[a * classes + b], 1u);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < cells; i += blockDim.x)
    if (sm[i]) atomicAdd(hist + i, (unsigned long long)sm[i]);
}

}  // namespace cvx

using namespace cvx;

extern "C" int cvx_seg_postprocess(const float* logits, int c, int h, int w, int crop_y, int crop_x, int crop_h,
                                   int crop_w, int out_h, int out_w, unsigned char* cls, float* probs, void* stream) {
  CVX_CHECK_ARG(logits && cls, "seg_postprocess: null pointer");
  CVX_CHECK_ARG(c >= 1 && c <= kMaxClasses, "seg_postprocess: 1..%d classes supported (got %d)", kMaxClasses, c);
  CVX_CHECK_ARG(crop_h >= 1 && crop_w >= 1 && crop_y >= 0 && crop_x >= 0 && crop_y + crop_h <= h && crop_x + crop_w <= w,
                "seg_postprocess: crop %dx%d at (%d,%d) outside the %dx%d logits", crop_h, crop_w, crop_y, crop_x, h, w);
  CVX_CHECK_ARG(out_h >= 1 && out_w >= 1 && out_h <= 65535, "seg_postprocess: bad output size %dx%d", out_h, out_w);
  // cv2 computes the scale in double and the coordinate in float
  const float sy = (float)((double)crop_h / (double)out_h), sx = (float)((double)crop_w / (double)out_w);
  const dim3 grid((out_w + 127) / 128, out_h);
  seg_postprocess_kernel<<<grid, 128, 0, as_stream(stream)>>>(logits, c, h, w, crop_y, crop_x, crop_h, crop_w, out_h, out_w,
                                                             sy, sx, cls, probs);
  CVX_LAUNCH_OK();
  return CVX_OK;
}

extern "C" int cvx_confusion_matrix(const unsigned char* pred, const unsigned char* gt, int64_t n, int classes,
                                    int64_t* hist, void* stream) {
  CVX_CHECK_ARG(pred && gt && hist && n >= 0, "confusion_matrix: null pointer");
  CVX_CHECK_ARG(classes >= 1 && classes <= kMaxHistClasses, "confusion_matrix: 1..%d classes supported (got %d)",
                kMaxHistClasses, classes);
  if (n == 0) return CVX_OK;
  int64_t blocks = ceil_div64(n, 256 * 16);
  if (blocks > kNumSMs * 8) blocks = kNumSMs * 8;
  confusion_matrix_kernel<<<(unsigned)blocks, 256, 0, as_stream(stream)>>>(pred, gt, n, classes,
                                                                         reinterpret_cast<unsigned long long*>(hist));
  CVX_LAUNCH_OK();
  return CVX_OK;
}
