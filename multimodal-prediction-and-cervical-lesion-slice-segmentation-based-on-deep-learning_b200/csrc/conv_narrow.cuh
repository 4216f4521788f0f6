// Internal launchers of the narrow-channel convolution kernels (conv_narrow.cu), used by the
// generic entry points in conv_simt.cu.  Each returns CVX_EUNSUPPORTED when the geometry is
// not one it specialises.
#pragma once
#include "common.cuh"

namespace cvx {
int narrow_conv_fwd(const cvx_conv_desc* d, const void* x, const void* wp, const float* bias, void* y, cudaStream_t st);
int narrow_conv_dgrad(const cvx_conv_desc* d, const void* dy, const void* wpt, void* dx, cudaStream_t st);
int narrow_conv_wgrad(const cvx_conv_desc* d, const void* x, const void* dy, float* dwp, cudaStream_t st);
}  // namespace cvx
