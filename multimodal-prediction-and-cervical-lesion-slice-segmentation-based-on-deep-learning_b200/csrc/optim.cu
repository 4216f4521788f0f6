// Fused optimizer steps (train.py:472-476: Adam(betas=(momentum,0.999)) / SGD(nesterov)).
#include "common.cuh"

namespace cvx {

__global__ void adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                            float* __restrict__ v, int64_t n, float lr, float b1, float b2, float eps, float wd,
                            float bc1, float bc2_sqrt, float gscale) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float gi = g[i] * gscale;
    const float pi = p[i];
    if (wd != 0.f) gi = fmaf(wd, pi, gi);
    const float mi = b1 * m[i] + (1.f - b1) * gi;
    const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
    m[i] = mi;
    v[i] = vi;
    const float denom = sqrtf(vi) / bc2_sqrt + eps;
    p[i] = pi - (lr / bc1) * (mi / denom);
  }
}

// graph-capturable variant: the step count and the hyper-parameters live in device memory, so a captured
// launch stays correct on every replay
__global__ void adam_dev_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                float* __restrict__ v, int64_t n, const float* __restrict__ hyper,
                                const int* __restrict__ step) {
  pdl_trigger();
  pdl_wait();
  const float lr = hyper[0], b1 = hyper[1], b2 = hyper[2], eps = hyper[3], wd = hyper[4], gscale = hyper[5];
  const float t = (float)(*step);
  const float bc1 = 1.f - powf(b1, t);
  const float bc2_sqrt = sqrtf(1.f - powf(b2, t));
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float gi = g[i] * gscale;
    const float pi = p[i];
    if (wd != 0.f) gi = fmaf(wd, pi, gi);
    const float mi = b1 * m[i] + (1.f - b1) * gi;
    const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
    m[i] = mi;
    v[i] = vi;
    p[i] = pi - (lr / bc1) * (mi / (sqrtf(vi) / bc2_sqrt + eps));
  }
}

// Gather of many gradient tensors into the flat gradient buffer in ONE launch (replaces one accumulate kernel per
// parameter, 435 for DeepLab-Xception): block b copies chunk b of tensor chunk_tensor[b]; a null source writes zeros
// (parameters that received no gradient).  Sources may be unaligned views, destinations are 16-byte aligned.
constexpr int kGatherChunk = 2048;   // floats per block

__global__ void __launch_bounds__(256) multi_gather_kernel(const int64_t* __restrict__ src_ptrs,
                                                           const int32_t* __restrict__ chunk_tensor,
                                                           const int32_t* __restrict__ chunk_start,
                                                           const int64_t* __restrict__ dst_offsets,
                                                           const int64_t* __restrict__ sizes, float* __restrict__ dst) {
  const int tsr = chunk_tensor[blockIdx.x];
  const int64_t start = chunk_start[blockIdx.x];
  const int64_t size = sizes[tsr];
  const float* src = reinterpret_cast<const float*>(src_ptrs[tsr]);
  float* out = dst + dst_offsets[tsr];
  int64_t end = start + kGatherChunk;
  if (end > size) end = size;
  if (src != nullptr && (reinterpret_cast<uintptr_t>(src) & 15) == 0) {
    const int64_t v0 = start / 4, v1 = end / 4;   // start is a multiple of the chunk size
    for (int64_t i = v0 + threadIdx.x; i < v1; i += blockDim.x)
      reinterpret_cast<float4*>(out)[i] = __ldg(reinterpret_cast<const float4*>(src) + i);
    for (int64_t i = v1 * 4 + threadIdx.x; i < end; i += blockDim.x) out[i] = src[i];
  } else {
    for (int64_t i = start + threadIdx.x; i < end; i += blockDim.x) out[i] = src ? src[i] : 0.f;
  }
}

__global__ void sgd_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ buf, int64_t n,
                           float lr, float mom, float wd, int nesterov, int first, float gscale) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float gi = g[i] * gscale;
    const float pi = p[i];
    if (wd != 0.f) gi = fmaf(wd, pi, gi);
    if (mom != 0.f) {
      const float b = first ? gi : mom * buf[i] + gi;
      buf[i] = b;
      gi = nesterov ? gi + mom * b : b;
    }
    p[i] = pi - lr * gi;
  }
}

// graph-capturable SGD: lr / momentum / weight decay / gradient scale from device memory (hyper layout shared with
// adam_dev_kernel: [lr, momentum, -, -, weight_decay, grad_scale]).  torch's first-step rule (buf = grad) is what a
// zero-initialised momentum buffer gives, so no step count is needed.
__global__ void sgd_dev_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ buf, int64_t n,
                               const float* __restrict__ hyper, int nesterov) {
  pdl_trigger();
  pdl_wait();
  const float lr = hyper[0], mom = hyper[1], wd = hyper[4], gscale = hyper[5];
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float gi = g[i] * gscale;
    const float pi = p[i];
    if (wd != 0.f) gi = fmaf(wd, pi, gi);
    if (mom != 0.f) {
      const float b = fmaf(mom, buf[i], gi);
      buf[i] = b;
      gi = nesterov ? fmaf(mom, b, gi) : b;
    }
    p[i] = pi - lr * gi;
  }
}

}  // namespace cvx

using namespace cvx;

extern "C" {

int cvx_adam_step(float* p, const float* g, float* m, float* v, int64_t n, float lr, float beta1, float beta2,
                  float eps, float weight_decay, int step_t, float grad_scale, void* stream) {
  CVX_CHECK_ARG(p && g && m && v && n > 0 && step_t >= 1, "adam_step: bad arguments");
  const double bc1 = 1.0 - pow((double)beta1, (double)step_t);
  const double bc2 = 1.0 - pow((double)beta2, (double)step_t);
  int blocks = (int)(ceil_div64(n, 256) > kNumSMs * 8 ? kNumSMs * 8 : ceil_div64(n, 256));
  adam_kernel<<<blocks, 256, 0, as_stream(stream)>>>(p, g, m, v, n, lr, beta1, beta2, eps, weight_decay, (float)bc1,
                                                     (float)sqrt(bc2), grad_scale);
  CVX_LAUNCH_OK();
  return CVX_OK;
}

int cvx_adam_step_dev(float* p, const float* g, float* m, float* v, int64_t n, const float* hyper, const int* step,
                      void* stream) {
  CVX_CHECK_ARG(p && g && m && v && hyper && step && n > 0, "adam_step_dev: bad arguments");
  int blocks = (int)(ceil_div64(n, 256) > kNumSMs * 8 ? kNumSMs * 8 : ceil_div64(n, 256));
  launch_pdl(adam_dev_kernel, dim3(blocks), dim3(256), 0, as_stream(stream), p, g, m, v, n, hyper, step);
  CVX_LAUNCH_OK();
  return CVX_OK;
}

int cvx_sgd_step_dev(float* p, const float* g, float* buf, int64_t n, const float* hyper, int nesterov, void* stream) {
  CVX_CHECK_ARG(p && g && buf && hyper && n > 0, "sgd_step_dev: bad arguments");
  int blocks = (int)(ceil_div64(n, 256) > kNumSMs * 8 ? kNumSMs * 8 : ceil_div64(n, 256));
  launch_pdl(sgd_dev_kernel, dim3(blocks), dim3(256), 0, as_stream(stream), p, g, buf, n, hyper, nesterov);
  CVX_LAUNCH_OK();
  return CVX_OK;
}

int cvx_multi_gather_chunk(void) { return kGatherChunk; }

int cvx_multi_gather(const int64_t* src_ptrs, const int32_t* chunk_tensor, const int32_t* chunk_start,
                     const int64_t* dst_offsets, const int64_t* sizes, int nchunks, float* dst, void* stream) {
  CVX_CHECK_ARG(src_ptrs && chunk_tensor && chunk_start && dst_offsets && sizes && dst && nchunks >= 0,
                "multi_gather: bad arguments");
  if (nchunks == 0) return CVX_OK;
  multi_gather_kernel<<<nchunks, 256, 0, as_stream(stream)>>>(src_ptrs, chunk_tensor, chunk_start, dst_offsets, sizes, dst);
  CVX_LAUNCH_OK();
  return CVX_OK;
}

int cvx_sgd_step(float* p, const float* g, float* buf, int64_t n, float lr, float momentum, float weight_decay,
                 int nesterov, int first_step, float grad_scale, void* stream) {
  CVX_CHECK_ARG(p && g && n > 0 && (momentum == 0.f || buf), "sgd_step: bad arguments");
  int blocks = (int)(ceil_div64(n, 256) > kNumSMs * 8 ? kNumSMs * 8 : ceil_div64(n, 256));
  sgd_kernel<<<blocks, 256, 0, as_stream(stream)>>>(p, g, buf, n, lr, momentum, weight_decay, nesterov, first_step,
                                                    grad_scale);
  CVX_LAUNCH_OK();
  return CVX_OK;
}

}  // extern "C"
