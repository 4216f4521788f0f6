// Internal launchers of the fused depthwise 3x3 kernels (dwconv_fused.cu); bf16, stride 1, dilation 1, pad 1 only
// (CVX_EUNSUPPORTED otherwise).  stats / sums: 2*C doubles, dw_out: 9*C doubles - zeroed by the launcher.
#pragma once
#include "common.cuh"
namespace cvx {
int dwf_fwd_launch(const cvx_conv_desc* d, const void* x, const float* w9c, const float* in_scale, const float* in_shift,
                   int relu_in, void* y, double* stats, cudaStream_t st);
// dside/negk/kmean (nullable together): the incoming gradient is dd + negk*dside + kmean inside the image
int dwf_bwd_launch(const cvx_conv_desc* d, const void* dd, const void* dside, const float* negk, const float* kmean,
                   const void* x, const float* w9c, const float* in_scale, const float* in_shift, int relu_in,
                   const void* addend, void* g, double* dw_out, double* sums, cudaStream_t st);
}  // namespace cvx
