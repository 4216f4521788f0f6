// Depthwise 3x3, stride 1, dilation 1 on NHWC bf16 with TMA halo staging (sm_100a).
//
// A CTA (128 threads = 8 channel vectors x 16 columns) walks over 8x16-pixel output tiles of
// one 64-channel chunk.  Each input tile (10 x 18 pixels x 64 ch, halo included) is fetched by
// ONE 4-D TMA box load whose out-of-bounds zero fill implements the padding and the channel
// tail; three tiles are in flight per CTA (mbarrier pipeline), so HBM latency is hidden without
// holding registers.  A thread slides a 3x3 register window down its column: 3 conflict-free
// 16-byte shared-memory loads + 72 FMAs per output vector, taps (or, for the weight gradient,
// the 9 partial sums) resident in registers.
//   mode 0  forward        y  = dw(relu?(x))
//   mode 1  data gradient  dx = dw_flipped(dy) * (x > 0 ?)         (second = x for the mask)
//   mode 2  weight gradient dw[t][c] = sum dy * relu?(x)[shifted]  (second = dy)
#include "dwconv_tiled.cuh"
#include "tma_utils.cuh"

namespace cvx {

constexpr int kTY = 8, kTX = 16, kDwStages = 3;
constexpr int kTileBytes = (kTY + 2) * (kTX + 2) * 128;  // 23040

struct DwTParams {
  int n, h, w, c;
  int tiles_x, tiles_y, ntiles;  // spatial tiles per launch
  int relu_in;
};

__device__ __forceinline__ void unpack8(const uint4& r, float (&v)[8]) {
  const uint32_t u[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    v[2 * i] = __uint_as_float(u[i] << 16);
    v[2 * i + 1] = __uint_as_float(u[i] & 0xffff0000u);
  }
}

template <int MODE>
__global__ void __launch_bounds__(128, MODE == 2 ? 2 : 3) dw_tiled_kernel(const __grid_constant__ CUtensorMap tmap,
                                                          const float* __restrict__ w9c,
                                                          const __nv_bfloat16* __restrict__ second,
                                                          __nv_bfloat16* __restrict__ dst, double* __restrict__ wgrad_out,
                                                          DwTParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~uintptr_t(127));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kDwStages * kTileBytes);
  const uint32_t smem_base = smem_u32(smem), bar0 = smem_u32(bars);

  const int t = threadIdx.x;
  const int cv = t & 7, col = t >> 3;
  const int cchunk = blockIdx.y;
  const int c0 = cchunk * 64 + cv * 8;
  const bool ch_ok = c0 < p.c;

  float wreg[9][8];
#pragma unroll
  for (int k = 0; k < 9; ++k)
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (MODE == 2) wreg[k][i] = 0.f;
      else wreg[k][i] = ch_ok ? __ldg(w9c + (MODE == 1 ? 8 - k : k) * p.c + c0 + i) : 0.f;
    }

  if (t == 0) {
    tma_prefetch_desc(&tmap);
    for (int s = 0; s < kDwStages; ++s) mbar_init(bar0 + 8 * s, 1);
    fence_barrier_init();
  }
  __syncthreads();

  const int my_tiles = (p.ntiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  auto issue = [&](int i) {  // thread 0 only
    const int tile = blockIdx.x + i * gridDim.x;
    const int tx = tile % p.tiles_x;
    const int t1 = tile / p.tiles_x;
    const int ty = t1 % p.tiles_y, img = t1 / p.tiles_y;
    const int s = i % kDwStages;
    mbar_expect_tx(bar0 + 8 * s, kTileBytes);
    tma_load_4d(smem_base + s * kTileBytes, &tmap, bar0 + 8 * s, cchunk * 64, tx * kTX - 1, ty * kTY - 1, img);
  };
  if (t == 0)
    for (int i = 0; i < kDwStages - 1 && i < my_tiles; ++i) issue(i);

  for (int i = 0; i < my_tiles; ++i) {
    if (t == 0 && i + kDwStages - 1 < my_tiles) {
      fence_proxy_async();  // generic-proxy reads of that slot (iteration i-1) precede this async write
      issue(i + kDwStages - 1);
    }
    const int s = i % kDwStages;
    mbar_wait(bar0 + 8 * s, (i / kDwStages) & 1);
    const uint8_t* tile_s = smem + s * kTileBytes;

    const int tile = blockIdx.x + i * gridDim.x;
    const int tx = tile % p.tiles_x;
    const int t1 = tile / p.tiles_x;
    const int ty = t1 % p.tiles_y, img = t1 / p.tiles_y;
    const int ox = tx * kTX + col, oy0 = ty * kTY;
    const bool col_ok = ch_ok && ox < p.w;

    // side input of this column (mask x for mode 1, dy for mode 2), all rows issued up front
    uint4 side[kTY];
    if (MODE != 0) {
#pragma unroll
      for (int r = 0; r < kTY; ++r) {
        const bool ok = col_ok && (oy0 + r) < p.h && (MODE == 2 || p.relu_in);
        side[r] = ok ? __ldg(reinterpret_cast<const uint4*>(second + (((size_t)img * p.h + oy0 + r) * p.w + ox) * p.c + c0))
                     : make_uint4(0, 0, 0, 0);
      }
    }

    float win[3][3][8];  // [row slot][kw][element]
    auto load_row = [&](int slot, int rr) {
#pragma unroll
      for (int kw = 0; kw < 3; ++kw) {
        const uint4 raw = *reinterpret_cast<const uint4*>(tile_s + ((rr * (kTX + 2) + col + kw) * 64 + cv * 8) * 2);
        unpack8(raw, win[slot][kw]);
        if (MODE != 1 && p.relu_in) {
#pragma unroll
          for (int e = 0; e < 8; ++e) win[slot][kw][e] = fmaxf(win[slot][kw][e], 0.f);
        }
      }
    };
    load_row(0, 0);
    load_row(1, 1);
#pragma unroll
    for (int r = 0; r < kTY; ++r) {
      load_row((r + 2) % 3, r + 2);
      if (MODE == 2) {
        float gv[8];
        unpack8(side[r], gv);  // zero outside the image / channel range
#pragma unroll
        for (int kh = 0; kh < 3; ++kh)
#pragma unroll
          for (int kw = 0; kw < 3; ++kw)
#pragma unroll
            for (int e = 0; e < 8; ++e)
              wreg[kh * 3 + kw][e] = fmaf(gv[e], win[(r + kh) % 3][kw][e], wreg[kh * 3 + kw][e]);
      } else {
        float acc[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[e] = 0.f;
#pragma unroll
        for (int kh = 0; kh < 3; ++kh)
#pragma unroll
          for (int kw = 0; kw < 3; ++kw)
#pragma unroll
            for (int e = 0; e < 8; ++e) acc[e] = fmaf(win[(r + kh) % 3][kw][e], wreg[kh * 3 + kw][e], acc[e]);
        if (MODE == 1 && p.relu_in) {
          float xm[8];
          unpack8(side[r], xm);
#pragma unroll
          for (int e = 0; e < 8; ++e) acc[e] = xm[e] > 0.f ? acc[e] : 0.f;
        }
        if (col_ok && (oy0 + r) < p.h) {
          uint32_t pk[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            __nv_bfloat162 hh = __floats2bfloat162_rn(acc[2 * e], acc[2 * e + 1]);
            pk[e] = *reinterpret_cast<uint32_t*>(&hh);
          }
          *reinterpret_cast<uint4*>(dst + (((size_t)img * p.h + oy0 + r) * p.w + ox) * p.c + c0) =
              make_uint4(pk[0], pk[1], pk[2], pk[3]);
        }
      }
    }
    __syncthreads();  // everyone is done with slot s before it is refilled
  }

  if (MODE == 2) {
    float* red = reinterpret_cast<float*>(smem);  // [9][64]
    for (int i = t; i < 9 * 64; i += 128) red[i] = 0.f;
    __syncthreads();
    // lanes l, l^8, l^16, l^24 of a warp share the channel vector: fold them first
#pragma unroll
    for (int k = 0; k < 9; ++k)
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        float v = wreg[k][e];
        v += __shfl_xor_sync(0xffffffffu, v, 8);
        v += __shfl_xor_sync(0xffffffffu, v, 16);
        if ((t & 31) < 8) atomicAdd(&red[k * 64 + cv * 8 + e], v);
      }
    __syncthreads();
    for (int i = t; i < 9 * 64; i += 128) {
      const int k = i / 64, cc = cchunk * 64 + (i % 64);
      if (cc < p.c) atomicAdd(wgrad_out + (size_t)k * p.c + cc, (double)red[i]);
    }
  }
}

int dw_tiled_launch(int mode, const cvx_conv_desc* d, const void* src, const float* w9c, const void* second, void* dst,
                    double* ws, int relu_in, cudaStream_t st) {
  static const bool disabled = getenv("CERVIX_DW_TILED_OFF") != nullptr;
  if (disabled || d->dtype != CVX_BF16 || d->stride != 1 || d->dil != 1 || d->pad != 1 || d->cin % 8 != 0)
    return CVX_EUNSUPPORTED;
  DwTParams p;
  p.n = d->n; p.h = d->h; p.w = d->w; p.c = d->cin;
  p.tiles_x = (d->w + kTX - 1) / kTX;
  p.tiles_y = (d->h + kTY - 1) / kTY;
  p.ntiles = d->n * p.tiles_x * p.tiles_y;
  p.relu_in = relu_in;
  CUtensorMap map;
  if (int rc = make_act_map(&map, src, d->n, d->h, d->w, d->cin, kTX + 2, kTY + 2, CU_TENSOR_MAP_SWIZZLE_NONE)) return rc;
  const int chunks = (d->cin + 63) / 64;
  int gx = (kNumSMs * (mode == 2 ? 2 : 3) + chunks - 1) / chunks;
  if (gx > p.ntiles) gx = p.ntiles;
  dim3 grid(gx, chunks);
  constexpr int smem = kDwStages * kTileBytes + 128 + 64;
  static bool configured = false;
  if (!configured) {
    CVX_CUDA_OK(cudaFuncSetAttribute(dw_tiled_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    CVX_CUDA_OK(cudaFuncSetAttribute(dw_tiled_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    CVX_CUDA_OK(cudaFuncSetAttribute(dw_tiled_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured = true;
  }
  const __nv_bfloat16* sec = (const __nv_bfloat16*)second;
  if (mode == 0) dw_tiled_kernel<0><<<grid, 128, smem, st>>>(map, w9c, nullptr, (__nv_bfloat16*)dst, nullptr, p);
  else if (mode == 1) dw_tiled_kernel<1><<<grid, 128, smem, st>>>(map, w9c, sec, (__nv_bfloat16*)dst, nullptr, p);
  else dw_tiled_kernel<2><<<grid, 128, smem, st>>>(map, nullptr, sec, nullptr, ws, p);
  CVX_LAUNCH_OK();
  return CVX_OK;
}

}  // namespace cvx
