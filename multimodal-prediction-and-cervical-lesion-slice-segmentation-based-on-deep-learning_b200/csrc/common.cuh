// Shared helpers for the cervix_b200 kernels (sm_100a only).
#pragma once

#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/cervix_b200.h"

namespace cvx {

void set_error(const char* fmt, ...);
extern unsigned long long g_kernel_launches;  // kernels launched by this library (bench.py's gpu_launches)
extern int g_ws_prezeroed;                    // cvx_set_ws_prezeroed: fp64 workspaces arrive zeroed (caller's per-step arena)

// clear an accumulation workspace unless the caller has promised it is already zero
#define CVX_WS_ZERO(ptr, bytes, st)                                       \
  do {                                                                    \
    if (!cvx::g_ws_prezeroed) CVX_CUDA_OK(cudaMemsetAsync((ptr), 0, (bytes), (st))); \
  } while (0)

#define CVX_CHECK_ARG(cond, ...)                 \
  do {                                           \
    if (!(cond)) {                               \
      cvx::set_error(__VA_ARGS__);               \
      return CVX_EINVAL;                         \
    }                                            \
  } while (0)

#define CVX_CUDA_OK(expr)                                                            \
  do {                                                                               \
    cudaError_t _e = (expr);                                                         \
    if (_e != cudaSuccess) {                                                         \
      cvx::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return CVX_ECUDA;                                                              \
    }                                                                                \
  } while (0)

#define CVX_LAUNCH_OK()                                                              \
  do {                                                                               \
    ++cvx::g_kernel_launches;                                                        \
    cudaError_t _e = cudaPeekAtLastError();                                          \
    if (_e != cudaSuccess) {                                                         \
      cvx::set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(_e), __FILE__, __LINE__); \
      return CVX_ECUDA;                                                              \
    }                                                                                \
  } while (0)

constexpr int kNumSMs = 148;  // B200

// ---- programmatic dependent launch (PDL) ------------------------------------------------------------------------------
// A training step is ~1 500 kernels of 5 - 500 us; between two dependent kernels the GPU normally drains completely, flushes,
// and only then starts launching the next grid.  Kernels launched through launch_pdl() carry the programmatic-stream-
// serialization attribute: their CTAs may become resident while the previous kernel is still draining (as its CTAs exit),
// run their prologue (barrier / TMEM set-up, descriptor prefetch) and block in pdl_wait() until the previous grid has
// completed and its writes are visible.  Every such kernel calls pdl_trigger() first thing, so the kernel AFTER it may be
// scheduled early in turn.  Nothing may read or write global memory before pdl_wait().  CERVIX_PDL=0 turns the attribute
// off (the device-side instructions are then no-ops).
extern int g_pdl;
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

template <typename... KArgs, typename... Args>
static inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                                     Args&&... args) {
  static_assert(sizeof...(KArgs) == sizeof...(Args), "launch_pdl: argument count mismatch");
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = g_pdl ? 1 : 0;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

static inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }
__host__ __device__ static inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }

// ---- element access -----------------------------------------------------------------
template <typename T> struct Elem;
template <> struct Elem<float> {
  static constexpr int kVec = 4;  // elements per 16-byte vector
  __device__ __forceinline__ static float ld(const float* p) { return *p; }
  __device__ __forceinline__ static void st(float* p, float v) { *p = v; }
};
template <> struct Elem<__nv_bfloat16> {
  static constexpr int kVec = 8;
  __device__ __forceinline__ static float ld(const __nv_bfloat16* p) { return __bfloat162float(*p); }
  __device__ __forceinline__ static void st(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }
};

// A 16-byte vector of T viewed as kVec floats.
template <typename T> struct Vec;
template <> struct Vec<float> {
  static constexpr int N = 4;
  float v[4];
  __device__ __forceinline__ void load(const float* p) {
    float4 t = *reinterpret_cast<const float4*>(p);
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
  }
  __device__ __forceinline__ void store(float* p) const {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  }
};
template <> struct Vec<__nv_bfloat16> {
  static constexpr int N = 8;
  float v[8];
  __device__ __forceinline__ void load(const __nv_bfloat16* p) {
    uint4 t = *reinterpret_cast<const uint4*>(p);
    const uint32_t u[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      v[2 * i] = __uint_as_float(u[i] << 16);
      v[2 * i + 1] = __uint_as_float(u[i] & 0xffff0000u);
    }
  }
  __device__ __forceinline__ void store(__nv_bfloat16* p) const {
    uint32_t u[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
      u[i] = *reinterpret_cast<uint32_t*>(&h);
    }
    *reinterpret_cast<uint4*>(p) = make_uint4(u[0], u[1], u[2], u[3]);
  }
};

template <typename V>
__device__ __forceinline__ void vec_zero(V& a) {
#pragma unroll
  for (int i = 0; i < V::N; ++i) a.v[i] = 0.f;
}

__device__ __forceinline__ float act_apply(float v, int act) {
  if (act == CVX_ACT_RELU) return fmaxf(v, 0.f);
  if (act == CVX_ACT_RELU6) return fminf(fmaxf(v, 0.f), 6.f);
  return v;
}
// derivative mask evaluated on the activation OUTPUT y
__device__ __forceinline__ float act_mask(float y, int act) {
  if (act == CVX_ACT_RELU) return y > 0.f ? 1.f : 0.f;
  if (act == CVX_ACT_RELU6) return (y > 0.f && y < 6.f) ? 1.f : 0.f;
  return 1.f;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Dispatch a templated launcher on the storage dtype.
#define CVX_DISPATCH_DTYPE(dtype, T, ...)                         \
  do {                                                            \
    if ((dtype) == CVX_F32) {                                     \
      using T = float;                                            \
      __VA_ARGS__;                                                \
    } else if ((dtype) == CVX_BF16) {                             \
      using T = __nv_bfloat16;                                    \
      __VA_ARGS__;                                                \
    } else {                                                      \
      cvx::set_error("unknown dtype %d", (int)(dtype));           \
      return CVX_EINVAL;                                          \
    }                                                             \
  } while (0)

}  // namespace cvx
