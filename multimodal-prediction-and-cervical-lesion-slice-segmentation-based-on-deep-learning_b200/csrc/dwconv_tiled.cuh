// Internal launcher of the TMA-staged depthwise 3x3 kernels (dwconv_tiled.cu).
#pragma once
#include "common.cuh"
namespace cvx {
// mode 0 forward, 1 data gradient, 2 weight gradient (ws: 9*C doubles, zeroed by the caller).
// Returns CVX_EUNSUPPORTED unless bf16, stride 1, dilation 1, pad 1.
int dw_tiled_launch(int mode, const cvx_conv_desc* d, const void* src, const float* w9c, const void* second, void* dst,
                    double* ws, int relu_in, cudaStream_t st);
}  // namespace cvx
