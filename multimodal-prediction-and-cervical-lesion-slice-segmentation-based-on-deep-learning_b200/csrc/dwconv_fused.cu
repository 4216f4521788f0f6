// Depthwise 3x3 (stride 1, dilation 1, pad 1) on NHWC bf16, fused with its neighbours in an Xception
// separable-conv chain (xception.py:9-31,33-73):
//
//   forward   d = dw( act(s*x + t) )                       and  stats[0/1][c] += sum d, sum d^2
//       (s, t) is the not-yet-applied affine of the PREVIOUS BatchNorm (its output is never materialised),
//       act = ReLU (SeparableConv2d.relu0); the per-channel sums feed the NEXT BatchNorm (bn1).
//   backward  one kernel for both gradients of the same layer:
//       g  = dw_flipped(dd) * 1[s*x + t > 0] (+ addend)    gradient w.r.t. the previous BatchNorm's OUTPUT
//       dw[k][c] += sum dd * act(s*x + t)[shifted by tap k]
//       sums[0/1][c] += sum g, sum g * x                   the previous BatchNorm's backward reduction
//
// Tiling as in dwconv_tiled.cu: a CTA walks over 8x16-pixel tiles of one 64-channel chunk, every halo tile
// (10 x 18 pixels x 64 ch) arrives by ONE 4-D TMA box load (zero fill = padding / channel tail), 3 tiles in
// flight.  The backward CTA has two warpgroups reading the same staged tiles: threads 0-127 hold the 9 taps and
// produce g, threads 128-255 hold the 9 weight-gradient accumulators.
#include "dwconv_fused.cuh"
#include "tma_utils.cuh"

namespace cvx {

constexpr int kFY = 12, kFX = 16;                     // output tile; rows are walked 3 at a time (window period)
constexpr int kFTile = (kFY + 2) * (kFX + 2) * 128;   // 32256 bytes: one halo tile of 64 channels
constexpr int kFwdStages = 3;                         // forward: 3 tiles in flight, 2 CTAs per SM
static_assert(kFY % 3 == 0, "the row loop is unrolled over the 3-row register window");

struct DwFParams {
  int n, h, w, c;
  int tiles_x, tiles_y, ntiles;
  int relu_in;
};

__device__ __forceinline__ void unpack8f(const uint4& r, float (&v)[8]) {
  const uint32_t u[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    v[2 * i] = __uint_as_float(u[i] << 16);
    v[2 * i + 1] = __uint_as_float(u[i] & 0xffff0000u);
  }
}
// 8 bf16 -> 4 packed fp32 pairs: the arithmetic below runs on FFMA2 / FADD2 (two fp32 lanes per instruction,
// sm_100), which halves the issue slots of the 9-tap loops that bound these kernels
__device__ __forceinline__ void unpack4x2(const uint4& r, float2 (&v)[4]) {
  const uint32_t u[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) v[i] = make_float2(__uint_as_float(u[i] << 16), __uint_as_float(u[i] & 0xffff0000u));
}
__device__ __forceinline__ uint32_t pack_bf16x2(float2 v) {
  __nv_bfloat162 h = __floats2bfloat162_rn(v.x, v.y);
  return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ float2 unpack_bf16x2(uint32_t u) {
  return make_float2(__uint_as_float(u << 16), __uint_as_float(u & 0xffff0000u));
}

// reduce per-thread [K][4] float2 partials over the 16 column-threads that share a channel vector, then fp64 atomics
template <int K>
__device__ __forceinline__ void reduce_to_global(float2 (&part)[K][4], float* red /* [K][64] smem */, int t_in_group,
                                                 int cv, int cchunk, int C, double* out, int group_bar) {
  for (int i = t_in_group; i < K * 64; i += 128) red[i] = 0.f;
  asm volatile("bar.sync %0, 128;" ::"r"(group_bar) : "memory");
#pragma unroll
  for (int k = 0; k < K; ++k)
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      float v = (e & 1) ? part[k][e >> 1].y : part[k][e >> 1].x;
      v += __shfl_xor_sync(0xffffffffu, v, 8);
      v += __shfl_xor_sync(0xffffffffu, v, 16);
      if ((t_in_group & 31) < 8) atomicAdd(&red[k * 64 + cv * 8 + e], v);
    }
  asm volatile("bar.sync %0, 128;" ::"r"(group_bar) : "memory");
  for (int i = t_in_group; i < K * 64; i += 128) {
    const int k = i / 64, cc = cchunk * 64 + (i % 64);
    if (cc < C) atomicAdd(out + (size_t)k * C + cc, (double)red[i]);
  }
}

// one 16-byte channel vector of the (virtual) depthwise input act(s*x+t) at halo-tile position (rr, cc)
template <bool AFFINE>
__device__ __forceinline__ void load_virtual(const uint8_t* tile, int rr, int cc, int cv, const float2 (&sc)[4],
                                             const float2 (&sh)[4], bool relu, bool valid, float2 (&out)[4]) {
  unpack4x2(*reinterpret_cast<const uint4*>(tile + ((rr * (kFX + 2) + cc) * 64 + cv * 8) * 2), out);
  if (AFFINE) {
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      float2 v = __ffma2_rn(out[e], sc[e], sh[e]);
      if (relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); }
      out[e] = valid ? v : make_float2(0.f, 0.f);   // zero padding applies to the virtual tensor, not to the raw one
    }
  } else if (relu) {
#pragma unroll
    for (int e = 0; e < 4; ++e) { out[e].x = fmaxf(out[e].x, 0.f); out[e].y = fmaxf(out[e].y, 0.f); }
  }
}

// ------------------------------------------------------------------------------------------ forward
template <bool AFFINE>
__global__ void __launch_bounds__(128, 2) dwf_fwd_kernel(const __grid_constant__ CUtensorMap tmap,
                                                         const float* __restrict__ w9c, const float* __restrict__ in_scale,
                                                         const float* __restrict__ in_shift, __nv_bfloat16* __restrict__ dst,
                                                         double* __restrict__ stats, DwFParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~uintptr_t(127));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kFwdStages * kFTile);
  const uint32_t smem_base = smem_u32(smem), bar0 = smem_u32(bars);

  const int t = threadIdx.x;
  const int cv = t & 7, col = t >> 3;
  const int cchunk = blockIdx.y;
  const int c0 = cchunk * 64 + cv * 8;
  const bool ch_ok = c0 < p.c;
  const bool relu = p.relu_in != 0;

  float2 wreg[9][4], sc[4], sh[4], s12[2][4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    sc[i] = (AFFINE && ch_ok) ? make_float2(__ldg(in_scale + c0 + 2 * i), __ldg(in_scale + c0 + 2 * i + 1)) : make_float2(1.f, 1.f);
    sh[i] = (AFFINE && ch_ok) ? make_float2(__ldg(in_shift + c0 + 2 * i), __ldg(in_shift + c0 + 2 * i + 1)) : make_float2(0.f, 0.f);
    s12[0][i] = s12[1][i] = make_float2(0.f, 0.f);
  }
#pragma unroll
  for (int k = 0; k < 9; ++k)
#pragma unroll
    for (int i = 0; i < 4; ++i)
      wreg[k][i] = ch_ok ? make_float2(__ldg(w9c + k * p.c + c0 + 2 * i), __ldg(w9c + k * p.c + c0 + 2 * i + 1))
                         : make_float2(0.f, 0.f);

  if (t == 0) {
    tma_prefetch_desc(&tmap);
    for (int s = 0; s < kFwdStages; ++s) mbar_init(bar0 + 8 * s, 1);
    fence_barrier_init();
  }
  __syncthreads();

  const int my_tiles = (p.ntiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  auto issue = [&](int i) {  // thread 0 only
    const int tile = blockIdx.x + i * gridDim.x;
    const int tx = tile % p.tiles_x;
    const int t1 = tile / p.tiles_x;
    const int ty = t1 % p.tiles_y, img = t1 / p.tiles_y;
    const int s = i % kFwdStages;
    mbar_expect_tx(bar0 + 8 * s, kFTile);
    tma_load_4d(smem_base + s * kFTile, &tmap, bar0 + 8 * s, cchunk * 64, tx * kFX - 1, ty * kFY - 1, img);
  };
  if (t == 0)
    for (int i = 0; i < kFwdStages - 1 && i < my_tiles; ++i) issue(i);

  for (int i = 0; i < my_tiles; ++i) {
    if (t == 0 && i + kFwdStages - 1 < my_tiles) {
      fence_proxy_async();
      issue(i + kFwdStages - 1);
    }
    const int s = i % kFwdStages;
    mbar_wait(bar0 + 8 * s, (i / kFwdStages) & 1);
    const uint8_t* tile_s = smem + s * kFTile;

    const int tile = blockIdx.x + i * gridDim.x;
    const int tx = tile % p.tiles_x;
    const int t1 = tile / p.tiles_x;
    const int ty = t1 % p.tiles_y, img = t1 / p.tiles_y;
    const int ox = tx * kFX + col, oy0 = ty * kFY;
    const bool col_ok = ch_ok && ox < p.w;
    bool cvalid[3];
#pragma unroll
    for (int kw = 0; kw < 3; ++kw) cvalid[kw] = (ox - 1 + kw) >= 0 && (ox - 1 + kw) < p.w;

    float2 win[3][3][4];
    auto load_row = [&](int slot, int rr) {
      const int iy = oy0 - 1 + rr;
      const bool rvalid = iy >= 0 && iy < p.h;
#pragma unroll
      for (int kw = 0; kw < 3; ++kw)
        load_virtual<AFFINE>(tile_s, rr, col + kw, cv, sc, sh, relu, rvalid && cvalid[kw], win[slot][kw]);
    };
    load_row(0, 0);
    load_row(1, 1);
#pragma unroll 1
    for (int r3 = 0; r3 < kFY; r3 += 3) {
      if (oy0 + r3 >= p.h) break;  // the rest of this tile lies below the image
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        const int r = r3 + j;
        load_row((j + 2) % 3, r + 2);
        float2 acc[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) acc[e] = make_float2(0.f, 0.f);
#pragma unroll
        for (int kh = 0; kh < 3; ++kh)
#pragma unroll
          for (int kw = 0; kw < 3; ++kw)
#pragma unroll
            for (int e = 0; e < 4; ++e) acc[e] = __ffma2_rn(win[(j + kh) % 3][kw][e], wreg[kh * 3 + kw][e], acc[e]);
        if (col_ok && (oy0 + r) < p.h) {
          uint32_t pk[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            pk[e] = pack_bf16x2(acc[e]);
            const float2 ab = unpack_bf16x2(pk[e]);
            s12[0][e] = __fadd2_rn(s12[0][e], ab);
            s12[1][e] = __ffma2_rn(ab, ab, s12[1][e]);
          }
          *reinterpret_cast<uint4*>(dst + (((size_t)img * p.h + oy0 + r) * p.w + ox) * p.c + c0) =
              make_uint4(pk[0], pk[1], pk[2], pk[3]);
        }
      }
    }
    __syncthreads();
  }

  if (stats) reduce_to_global<2>(s12, reinterpret_cast<float*>(smem), t, cv, cchunk, p.c, stats, 1);
}

// ------------------------------------------------------------------------------------------ backward
// SIDE: the incoming gradient is not materialised either - it is assembled on load from the pointwise conv's
// data gradient e and the depthwise output d:  dd = e + negk*d + kmean  (bn1's backward, see sepconv.cu), zero
// outside the image.
template <bool AFFINE, bool SIDE>
__global__ void __launch_bounds__(256, 1) dwf_bwd_kernel(const __grid_constant__ CUtensorMap tmap_dd,
                                                         const __grid_constant__ CUtensorMap tmap_d,
                                                         const __grid_constant__ CUtensorMap tmap_x,
                                                         const float* __restrict__ w9c, const float* __restrict__ in_scale,
                                                         const float* __restrict__ in_shift, const float* __restrict__ negk,
                                                         const float* __restrict__ kmean,
                                                         const __nv_bfloat16* __restrict__ addend,
                                                         __nv_bfloat16* __restrict__ gout, double* __restrict__ dw_out,
                                                         double* __restrict__ sums, DwFParams p) {
  constexpr int kTiles = SIDE ? 3 : 2;
  constexpr int kStage = kTiles * kFTile;
  constexpr int kStages = SIDE ? 2 : 3;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~uintptr_t(127));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kStages * kStage);
  const uint32_t smem_base = smem_u32(smem), bar0 = smem_u32(bars);

  const int t = threadIdx.x;
  const int role = t >> 7;           // 0: data gradient, 1: weight gradient
  const int tg = t & 127;
  const int cv = tg & 7, col = tg >> 3;
  const int cchunk = blockIdx.y;
  const int c0 = cchunk * 64 + cv * 8;
  const bool ch_ok = c0 < p.c;
  const bool relu = p.relu_in != 0;

  // role 0: wreg = flipped taps ; role 1: wreg = weight-gradient accumulators
  float2 wreg[9][4], sc[4], sh[4], nk[4], km[4], s12[2][4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int c = c0 + 2 * i;
    sc[i] = (AFFINE && ch_ok) ? make_float2(__ldg(in_scale + c), __ldg(in_scale + c + 1)) : make_float2(1.f, 1.f);
    sh[i] = (AFFINE && ch_ok) ? make_float2(__ldg(in_shift + c), __ldg(in_shift + c + 1)) : make_float2(0.f, 0.f);
    nk[i] = (SIDE && ch_ok) ? make_float2(__ldg(negk + c), __ldg(negk + c + 1)) : make_float2(0.f, 0.f);
    km[i] = (SIDE && ch_ok) ? make_float2(__ldg(kmean + c), __ldg(kmean + c + 1)) : make_float2(0.f, 0.f);
    s12[0][i] = s12[1][i] = make_float2(0.f, 0.f);
  }
#pragma unroll
  for (int k = 0; k < 9; ++k)
#pragma unroll
    for (int i = 0; i < 4; ++i)
      wreg[k][i] = (role == 0 && ch_ok) ? make_float2(__ldg(w9c + (8 - k) * p.c + c0 + 2 * i), __ldg(w9c + (8 - k) * p.c + c0 + 2 * i + 1))
                                        : make_float2(0.f, 0.f);

  if (t == 0) {
    tma_prefetch_desc(&tmap_dd);
    if (SIDE) tma_prefetch_desc(&tmap_d);
    tma_prefetch_desc(&tmap_x);
    for (int s = 0; s < kStages; ++s) mbar_init(bar0 + 8 * s, 1);
    fence_barrier_init();
  }
  __syncthreads();

  const int my_tiles = (p.ntiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  auto issue = [&](int i) {  // thread 0 only
    const int tile = blockIdx.x + i * gridDim.x;
    const int tx = tile % p.tiles_x;
    const int t1 = tile / p.tiles_x;
    const int ty = t1 % p.tiles_y, img = t1 / p.tiles_y;
    const int s = i % kStages;
    const uint32_t dstp = smem_base + s * kStage;
    mbar_expect_tx(bar0 + 8 * s, kStage);
    tma_load_4d(dstp, &tmap_dd, bar0 + 8 * s, cchunk * 64, tx * kFX - 1, ty * kFY - 1, img);
    tma_load_4d(dstp + kFTile, &tmap_x, bar0 + 8 * s, cchunk * 64, tx * kFX - 1, ty * kFY - 1, img);
    if (SIDE) tma_load_4d(dstp + 2 * kFTile, &tmap_d, bar0 + 8 * s, cchunk * 64, tx * kFX - 1, ty * kFY - 1, img);
  };
  if (t == 0)
    for (int i = 0; i < kStages - 1 && i < my_tiles; ++i) issue(i);

  for (int i = 0; i < my_tiles; ++i) {
    if (t == 0 && i + kStages - 1 < my_tiles) {
      fence_proxy_async();
      issue(i + kStages - 1);
    }
    const int s = i % kStages;
    mbar_wait(bar0 + 8 * s, (i / kStages) & 1);
    const uint8_t* dd_s = smem + s * kStage;
    const uint8_t* x_s = dd_s + kFTile;
    const uint8_t* d_s = dd_s + 2 * kFTile;

    const int tile = blockIdx.x + i * gridDim.x;
    const int tx = tile % p.tiles_x;
    const int t1 = tile / p.tiles_x;
    const int ty = t1 % p.tiles_y, img = t1 / p.tiles_y;
    const int ox = tx * kFX + col, oy0 = ty * kFY;
    const bool col_ok = ch_ok && ox < p.w;
    bool cvalid[3];
#pragma unroll
    for (int kw = 0; kw < 3; ++kw) cvalid[kw] = (ox - 1 + kw) >= 0 && (ox - 1 + kw) < p.w;

    if (SIDE) {
      // ---- pre-pass, all 256 threads: assemble each halo element ONCE instead of once per tap and per role:
      //   dd  = e + negk*d + kmean  (bn1's backward)          -> written over the e tile
      //   xin = act(in_scale*x + in_shift)  (virtual input)   -> written over the d tile
      // both zero outside the image; thread t always handles channel vector t & 7 (= cv).
      uint8_t* e_w = smem + s * kStage;
      uint8_t* d_w = e_w + 2 * kFTile;
      const int ox0 = tx * kFX;
      for (int v = t; v < (kFY + 2) * (kFX + 2) * 8; v += 256) {
        const int pix = v >> 3;
        const int rr = pix / (kFX + 2), cc = pix - rr * (kFX + 2);
        const int iy = oy0 - 1 + rr, ix = ox0 - 1 + cc;
        if (iy > p.h) break;  // rows below the image feed nothing (warp-uniform up to the last partial row)
        const bool valid = iy >= 0 && iy < p.h && ix >= 0 && ix < p.w;
        const int off = (pix * 64 + cv * 8) * 2;
        uint32_t pd[4] = {0u, 0u, 0u, 0u}, px[4] = {0u, 0u, 0u, 0u};
        if (valid) {
          float2 ev[4], dv[4], xv[4];
          unpack4x2(*reinterpret_cast<const uint4*>(e_w + off), ev);
          unpack4x2(*reinterpret_cast<const uint4*>(d_w + off), dv);
          unpack4x2(*reinterpret_cast<const uint4*>(x_s + off), xv);
#pragma unroll
          for (int e2 = 0; e2 < 4; ++e2) {
            const float2 a = __ffma2_rn(nk[e2], dv[e2], __fadd2_rn(ev[e2], km[e2]));
            float2 b = AFFINE ? __ffma2_rn(xv[e2], sc[e2], sh[e2]) : xv[e2];
            if (relu) { b.x = fmaxf(b.x, 0.f); b.y = fmaxf(b.y, 0.f); }
            pd[e2] = pack_bf16x2(a);
            px[e2] = pack_bf16x2(b);
          }
        }
        *reinterpret_cast<uint4*>(e_w + off) = make_uint4(pd[0], pd[1], pd[2], pd[3]);
        *reinterpret_cast<uint4*>(d_w + off) = make_uint4(px[0], px[1], px[2], px[3]);
      }
      __syncthreads();
    }
    const uint8_t* xin_s = d_s;  // SIDE only: the transformed input tile

    auto load_plain = [&](const uint8_t* tile, int rr, int cc, float2 (&out)[4]) {
      unpack4x2(*reinterpret_cast<const uint4*>(tile + ((rr * (kFX + 2) + cc) * 64 + cv * 8) * 2), out);
    };

    float2 win[3][3][4];
    if (role == 0) {
      // ---- data gradient: g = sum_k dd[shifted] * w[8-k], masked by the ReLU of the (virtual) input
      auto load_row = [&](int slot, int rr) {
#pragma unroll
        for (int kw = 0; kw < 3; ++kw) load_plain(dd_s, rr, col + kw, win[slot][kw]);
      };
      load_row(0, 0);
      load_row(1, 1);
#pragma unroll 1
      for (int r3 = 0; r3 < kFY; r3 += 3) {
        if (oy0 + r3 >= p.h) break;  // the rest of this tile lies below the image
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          const int r = r3 + j;
          const bool out_ok = col_ok && (oy0 + r) < p.h;
          const size_t off = (((size_t)img * p.h + oy0 + r) * p.w + ox) * p.c + c0;
          uint4 add_raw = make_uint4(0, 0, 0, 0);
          if (addend && out_ok) add_raw = __ldg(reinterpret_cast<const uint4*>(addend + off));  // issued ahead of its use
          load_row((j + 2) % 3, r + 2);
          float2 acc[4], xc[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) acc[e] = make_float2(0.f, 0.f);
#pragma unroll
          for (int kh = 0; kh < 3; ++kh)
#pragma unroll
            for (int kw = 0; kw < 3; ++kw)
#pragma unroll
              for (int e = 0; e < 4; ++e) acc[e] = __ffma2_rn(win[(j + kh) % 3][kw][e], wreg[kh * 3 + kw][e], acc[e]);
          load_plain(x_s, r + 1, col + 1, xc);
          if (relu) {
            if (SIDE) {
              // xin = relu(..) >= 0 was stored as bf16: it is positive iff its bit pattern is non-zero
              const uint4 xt = *reinterpret_cast<const uint4*>(xin_s + (((r + 1) * (kFX + 2) + col + 1) * 64 + cv * 8) * 2);
              const uint32_t xu[4] = {xt.x, xt.y, xt.z, xt.w};
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                acc[e].x = (xu[e] & 0x0000ffffu) ? acc[e].x : 0.f;
                acc[e].y = (xu[e] & 0xffff0000u) ? acc[e].y : 0.f;
              }
            } else {
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const float2 pre = __ffma2_rn(xc[e], sc[e], sh[e]);
                acc[e].x = pre.x > 0.f ? acc[e].x : 0.f;
                acc[e].y = pre.y > 0.f ? acc[e].y : 0.f;
              }
            }
          }
          if (out_ok) {
            float2 av[4];
            if (addend) unpack4x2(add_raw, av);
            uint32_t pk[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              // the BatchNorm reduction uses g as the apply kernel will read it back (bf16), without the addend
              uint32_t u = pack_bf16x2(acc[e]);
              const float2 ab = unpack_bf16x2(u);
              s12[0][e] = __fadd2_rn(s12[0][e], ab);
              s12[1][e] = __ffma2_rn(ab, xc[e], s12[1][e]);
              if (addend) u = pack_bf16x2(__fadd2_rn(acc[e], av[e]));
              pk[e] = u;
            }
            *reinterpret_cast<uint4*>(gout + off) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
          }
        }
      }
    } else {
      // ---- weight gradient: dw[k] += dd[centre] * act(s*x+t)[shifted by k]
      auto load_row = [&](int slot, int rr) {
        const int iy = oy0 - 1 + rr;
        const bool rvalid = iy >= 0 && iy < p.h;
#pragma unroll
        for (int kw = 0; kw < 3; ++kw) {
          if (SIDE) load_plain(xin_s, rr, col + kw, win[slot][kw]);
          else load_virtual<AFFINE>(x_s, rr, col + kw, cv, sc, sh, relu, rvalid && cvalid[kw], win[slot][kw]);
        }
      };
      load_row(0, 0);
      load_row(1, 1);
#pragma unroll 1
      for (int r3 = 0; r3 < kFY; r3 += 3) {
        if (oy0 + r3 >= p.h) break;
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          const int r = r3 + j;
          load_row((j + 2) % 3, r + 2);
          float2 gv[4];  // dd at the output pixel: zero outside the image / channel range
          load_plain(dd_s, r + 1, col + 1, gv);
#pragma unroll
          for (int kh = 0; kh < 3; ++kh)
#pragma unroll
            for (int kw = 0; kw < 3; ++kw)
#pragma unroll
              for (int e = 0; e < 4; ++e)
                wreg[kh * 3 + kw][e] = __ffma2_rn(gv[e], win[(j + kh) % 3][kw][e], wreg[kh * 3 + kw][e]);
        }
      }
    }
    __syncthreads();  // both roles are done with slot s before it is refilled
  }

  float* red = reinterpret_cast<float*>(smem) + role * (9 * 64);
  if (role == 1) {
    reduce_to_global<9>(wreg, red, tg, cv, cchunk, p.c, dw_out, 2);
  } else if (sums) {
    reduce_to_global<2>(s12, red, tg, cv, cchunk, p.c, sums, 1);
  }
}

static bool dwf_supported(const cvx_conv_desc* d) {
  return d->dtype == CVX_BF16 && d->stride == 1 && d->dil == 1 && d->pad == 1 && d->kh == 3 && d->kw == 3 &&
         d->cin % 8 == 0 && d->cin == d->cout;
}

static void dwf_params(const cvx_conv_desc* d, int relu_in, DwFParams* p) {
  p->n = d->n; p->h = d->h; p->w = d->w; p->c = d->cin;
  p->tiles_x = (d->w + kFX - 1) / kFX;
  p->tiles_y = (d->h + kFY - 1) / kFY;
  p->ntiles = d->n * p->tiles_x * p->tiles_y;
  p->relu_in = relu_in;
}

int dwf_fwd_launch(const cvx_conv_desc* d, const void* x, const float* w9c, const float* in_scale, const float* in_shift,
                   int relu_in, void* y, double* stats, cudaStream_t st) {
  if (!dwf_supported(d)) return CVX_EUNSUPPORTED;
  DwFParams p;
  dwf_params(d, relu_in, &p);
  CUtensorMap map;
  if (int rc = make_act_map(&map, x, d->n, d->h, d->w, d->cin, kFX + 2, kFY + 2, CU_TENSOR_MAP_SWIZZLE_NONE)) return rc;
  const int chunks = (d->cin + 63) / 64;
  int gx = (kNumSMs * 2) / chunks;
  if (gx < 1) gx = 1;
  if (gx > p.ntiles) gx = p.ntiles;
  constexpr int smem = kFwdStages * kFTile + 128 + 64;
  static bool configured = false;
  if (!configured) {
    CVX_CUDA_OK(cudaFuncSetAttribute(dwf_fwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    CVX_CUDA_OK(cudaFuncSetAttribute(dwf_fwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured = true;
  }
  if (stats) CVX_CUDA_OK(cudaMemsetAsync(stats, 0, sizeof(double) * 2 * d->cin, st));
  if (in_scale)
    dwf_fwd_kernel<true><<<dim3(gx, chunks), 128, smem, st>>>(map, w9c, in_scale, in_shift, (__nv_bfloat16*)y, stats, p);
  else
    dwf_fwd_kernel<false><<<dim3(gx, chunks), 128, smem, st>>>(map, w9c, nullptr, nullptr, (__nv_bfloat16*)y, stats, p);
  CVX_LAUNCH_OK();
  return CVX_OK;
}

template <bool AFFINE, bool SIDE>
static int dwf_bwd_launch_t(const CUtensorMap& mdd, const CUtensorMap& md, const CUtensorMap& mx, const float* w9c,
                            const float* in_scale, const float* in_shift, const float* negk, const float* kmean,
                            const void* addend, void* g, double* dw_out, double* sums, const DwFParams& p, dim3 grid,
                            cudaStream_t st) {
  constexpr int smem = (SIDE ? 2 * 3 : 3 * 2) * kFTile + 128 + 64;
  static bool configured = false;
  if (!configured) {
    CVX_CUDA_OK(cudaFuncSetAttribute(dwf_bwd_kernel<AFFINE, SIDE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured = true;
  }
  dwf_bwd_kernel<AFFINE, SIDE><<<grid, 256, smem, st>>>(mdd, md, mx, w9c, in_scale, in_shift, negk, kmean,
                                                        (const __nv_bfloat16*)addend, (__nv_bfloat16*)g, dw_out, sums, p);
  CVX_LAUNCH_OK();
  return CVX_OK;
}

int dwf_bwd_launch(const cvx_conv_desc* d, const void* dd, const void* dside, const float* negk, const float* kmean,
                   const void* x, const float* w9c, const float* in_scale, const float* in_shift, int relu_in,
                   const void* addend, void* g, double* dw_out, double* sums, cudaStream_t st) {
  if (!dwf_supported(d)) return CVX_EUNSUPPORTED;
  DwFParams p;
  dwf_params(d, relu_in, &p);
  CUtensorMap mdd, md, mx;
  if (int rc = make_act_map(&mdd, dd, d->n, d->h, d->w, d->cin, kFX + 2, kFY + 2, CU_TENSOR_MAP_SWIZZLE_NONE)) return rc;
  if (int rc = make_act_map(&mx, x, d->n, d->h, d->w, d->cin, kFX + 2, kFY + 2, CU_TENSOR_MAP_SWIZZLE_NONE)) return rc;
  md = mx;
  if (dside)
    if (int rc = make_act_map(&md, dside, d->n, d->h, d->w, d->cin, kFX + 2, kFY + 2, CU_TENSOR_MAP_SWIZZLE_NONE)) return rc;
  const int chunks = (d->cin + 63) / 64;
  int gx = kNumSMs / chunks;  // one 256-thread CTA per SM
  if (gx < 1) gx = 1;
  if (gx > p.ntiles) gx = p.ntiles;
  const dim3 grid(gx, chunks);
  CVX_CUDA_OK(cudaMemsetAsync(dw_out, 0, sizeof(double) * 9 * d->cin, st));
  if (sums) CVX_CUDA_OK(cudaMemsetAsync(sums, 0, sizeof(double) * 2 * d->cin, st));
  if (dside) {
    if (in_scale) return dwf_bwd_launch_t<true, true>(mdd, md, mx, w9c, in_scale, in_shift, negk, kmean, addend, g, dw_out, sums, p, grid, st);
    return dwf_bwd_launch_t<false, true>(mdd, md, mx, w9c, nullptr, nullptr, negk, kmean, addend, g, dw_out, sums, p, grid, st);
  }
  if (in_scale) return dwf_bwd_launch_t<true, false>(mdd, md, mx, w9c, in_scale, in_shift, nullptr, nullptr, addend, g, dw_out, sums, p, grid, st);
  return dwf_bwd_launch_t<false, false>(mdd, md, mx, w9c, nullptr, nullptr, nullptr, nullptr, addend, g, dw_out, sums, p, grid, st);
}

}  // namespace cvx
