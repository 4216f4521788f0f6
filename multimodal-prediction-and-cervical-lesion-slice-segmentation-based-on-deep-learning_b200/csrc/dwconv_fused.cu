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
// A CTA (256 threads = 16 output columns x 16 channel quads) walks over 12x16-pixel tiles of one 64-channel chunk;
// every halo tile (14 x 18 pixels x 64 ch) arrives by ONE 4-D TMA box load (zero fill = padding / channel tail),
// 2-3 tiles in flight.  Each thread walks down the 12 rows of its column with a 3x3 register window, four channels
// wide; the arithmetic runs on FFMA2 / FADD2 (two fp32 lanes per instruction, sm_100) and all shared-memory traffic
// uses 32-bit shared addresses with immediate offsets, so the 9-tap loops, not address arithmetic, fill the issue slots.
#include "dwconv_fused.cuh"
#include "tma_utils.cuh"

namespace cvx {

constexpr int kFX = 16;                               // output tile width; the tile height FY is a template parameter
constexpr int kFRow = (kFX + 2) * 128;                // bytes of one halo row (18 pixels x 64 ch)
__host__ __device__ constexpr int ftile_bytes(int fy) { return (fy + 2) * kFRow; }   // one halo tile of 64 channels (32256 bytes at FY = 12)
constexpr int kFwdStages = 3;                         // forward: 3 tiles in flight
// Two shapes of each kernel: the tall tile (12 rows: 17 % halo rows, fewest bytes staged) and the short tile (6 rows:
// 33 % halo rows, but a stage is 18 KB so twice as many CTAs fit an SM).  The kernels sit at 8 - 16 resident warps per SM
// (ncu: 45 % issue-active, DRAM 24 %); the short-tile experiment tested whether residency beats halo traffic - it does
// not (see g_dwf_variant below); CERVIX_DWF_VARIANT selects among the shapes for A/B measurements.

struct DwFParams {
  int n, h, w, c;
  int tiles_x, tiles_y, ntiles;
  int relu_in;
};
// 1 = tall tiles, weights in registers (default: measured fastest), 0 = short tiles + weights in shared memory,
// 2 = short tiles, weights in registers under the same register cap (the compiler spills).  Measured on the middle-flow
// tensor [32,32,32,728], CUDA-graph timing (profiles/r02_dwf_variants.txt): backward 59 us (1) / 120 us (0) / 200 us (2),
// forward 44 / 57 / 68 us - residency bought with half-height tiles loses to the doubled per-tile overhead (TMA issue,
// barrier round trip, pre-pass and index arithmetic per tile) and to the spills under the 128-register cap.
static const int g_dwf_variant = [] { const char* e = getenv("CERVIX_DWF_VARIANT"); return e ? atoi(e) : 1; }();

__device__ __forceinline__ uint2 lds64(uint32_t addr) {
  uint2 v;
  asm volatile("ld.shared.v2.b32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(addr));
  return v;
}
// four fp32 (this thread's channel quad of one tap's weights) from shared memory as two packed pairs; volatile so the
// compiler re-reads them every row instead of pinning 36 registers on loop-invariant weights (WSMEM variants)
__device__ __forceinline__ void lds_w4(uint32_t addr, float2 (&w)[2]) {
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(w[0].x), "=f"(w[0].y), "=f"(w[1].x), "=f"(w[1].y) : "r"(addr));
}
__device__ __forceinline__ void sts64(uint32_t addr, uint2 v) {
  asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(addr), "r"(v.x), "r"(v.y) : "memory");
}
__device__ __forceinline__ uint32_t pack_bf16x2(float2 v) {
  __nv_bfloat162 h = __floats2bfloat162_rn(v.x, v.y);
  return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ float2 unpack_bf16x2(uint32_t u) {
  return make_float2(__uint_as_float(u << 16), __uint_as_float(u & 0xffff0000u));
}
// four bf16 channels at a shared address -> two packed fp32 pairs
__device__ __forceinline__ void load4(uint32_t addr, float2 (&v)[2]) {
  const uint2 r = lds64(addr);
  v[0] = unpack_bf16x2(r.x);
  v[1] = unpack_bf16x2(r.y);
}
__device__ __forceinline__ float2 relu2(float2 v) { return make_float2(fmaxf(v.x, 0.f), fmaxf(v.y, 0.f)); }

// reduce per-thread [K][2] float2 partials over the 16 column-threads that share a channel quad, then fp64 atomics.
// Fixed order inside the block (shuffle fold of the two half-warps, then the eight warps take turns on the 64-entry tile -
// no shared atomics), so a block's partial is bit-reproducible; the fp64 global sum of such partials is exact.
template <int K>
__device__ __forceinline__ void reduce4_to_global(float2 (&part)[K][2], float* red /* [K][64] smem */, int t, int cq,
                                                  int cchunk, int C, double* out) {
  for (int i = t; i < K * 64; i += 256) red[i] = 0.f;
  __syncthreads();
#pragma unroll
  for (int k = 0; k < K; ++k)
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      float v = (e & 1) ? part[k][e >> 1].y : part[k][e >> 1].x;
      v += __shfl_xor_sync(0xffffffffu, v, 16);
      if (e & 1) part[k][e >> 1].y = v; else part[k][e >> 1].x = v;
    }
  for (int w = 0; w < 8; ++w) {
    if ((t >> 5) == w && (t & 31) < 16) {
#pragma unroll
      for (int k = 0; k < K; ++k)
#pragma unroll
        for (int e = 0; e < 4; ++e) red[k * 64 + cq * 4 + e] += (e & 1) ? part[k][e >> 1].y : part[k][e >> 1].x;
    }
    __syncthreads();
  }
  for (int i = t; i < K * 64; i += 256) {
    const int k = i / 64, cc = cchunk * 64 + (i % 64);
    if (cc < C) atomicAdd(out + (size_t)k * C + cc, (double)red[i]);
  }
  __syncthreads();
}

// tile index -> (image, tile row, tile column)
__device__ __forceinline__ void tile_coords(const DwFParams& p, int tile, int& tx, int& ty, int& img) {
  tx = tile % p.tiles_x;
  const int t1 = tile / p.tiles_x;
  ty = t1 % p.tiles_y;
  img = t1 / p.tiles_y;
}

// ------------------------------------------------------------------------------------------ forward
// AFFINE: a pre-pass turns the staged halo tile into the virtual input act(s*x+t) ONCE per element (zero outside the
// image: the padding applies to the virtual tensor), instead of once per tap column in the window loads.
template <bool AFFINE, int FY, int MINB, bool WSMEM>
__global__ void __launch_bounds__(256, MINB) dwf_fwd_kernel(const __grid_constant__ CUtensorMap tmap,
                                                         const float* __restrict__ w9c, const float* __restrict__ in_scale,
                                                         const float* __restrict__ in_shift, __nv_bfloat16* __restrict__ dst,
                                                         double* __restrict__ stats, DwFParams p) {
  pdl_trigger();
  pdl_wait();
  constexpr int kFY = FY, kFTile = ftile_bytes(FY);
  static_assert(FY % 3 == 0, "the row loop is unrolled over the 3-row register window");
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~uintptr_t(127));
  const uint32_t smem_base = smem_u32(smem), bar0 = smem_base + kFwdStages * kFTile;

  const int t = threadIdx.x;
  const int cq = t & 15, col = t >> 4;
  const int cchunk = blockIdx.y;
  const int c0 = cchunk * 64 + cq * 4;
  const bool ch_ok = c0 < p.c;
  const bool relu = p.relu_in != 0;
  const uint32_t thr_off = col * 128 + cq * 8;   // this thread's (column, quad) inside a halo row

  float2 wreg[WSMEM ? 1 : 9][2], sc[2], sh[2], s12[2][2];
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int c = c0 + 2 * i;
    sc[i] = (AFFINE && ch_ok) ? make_float2(__ldg(in_scale + c), __ldg(in_scale + c + 1)) : make_float2(1.f, 1.f);
    sh[i] = (AFFINE && ch_ok) ? make_float2(__ldg(in_shift + c), __ldg(in_shift + c + 1)) : make_float2(0.f, 0.f);
    s12[0][i] = s12[1][i] = make_float2(0.f, 0.f);
  }
  // the chunk's 9 x 64 tap weights: in registers (36 per thread), or staged once in shared memory [tap][channel]
  const uint32_t wsm = bar0 + 64, w_thr = wsm + cq * 16;
  if (WSMEM) {
    float* wdst = reinterpret_cast<float*>(smem + kFwdStages * kFTile + 64);
    for (int i = t; i < 9 * 64; i += 256) {
      const int c = cchunk * 64 + (i & 63);
      wdst[i] = c < p.c ? __ldg(w9c + (i >> 6) * p.c + c) : 0.f;
    }
  } else {
#pragma unroll
    for (int k = 0; k < (WSMEM ? 1 : 9); ++k)
#pragma unroll
      for (int i = 0; i < 2; ++i)
        wreg[k][i] = ch_ok ? make_float2(__ldg(w9c + k * p.c + c0 + 2 * i), __ldg(w9c + k * p.c + c0 + 2 * i + 1))
                           : make_float2(0.f, 0.f);
  }

  if (t == 0) {
    tma_prefetch_desc(&tmap);
    for (int s = 0; s < kFwdStages; ++s) mbar_init(bar0 + 8 * s, 1);
    fence_barrier_init();
  }
  __syncthreads();

  const int my_tiles = (p.ntiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  auto issue = [&](int i) {  // thread 0 only
    int tx, ty, img;
    tile_coords(p, blockIdx.x + i * gridDim.x, tx, ty, img);
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
    const uint32_t tile_a = smem_base + s * kFTile;
    int tx, ty, img;
    tile_coords(p, blockIdx.x + i * gridDim.x, tx, ty, img);
    const int ox = tx * kFX + col, oy0 = ty * kFY;

    if (AFFINE) {
      // ---- pre-pass over the 14 x 18 x 16 quads of the halo tile, two elements per thread in flight; the element
      // index advances by 256 threads = 16 pixels, so a thread keeps its channel quad (v & 15 == cq)
      constexpr int kElems = (kFY + 2) * (kFX + 2) * 16;
      const int ox0 = tx * kFX - 1, oyb = oy0 - 1;
      for (int v = t; v < kElems; v += 512) {
        uint2 raw[2];
        bool valid[2];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          const int pix = (v + 256 * u) >> 4;
          const int rr = pix / (kFX + 2), cc = pix - rr * (kFX + 2);
          const int iy = oyb + rr, ix = ox0 + cc;
          valid[u] = v + 256 * u < kElems && iy >= 0 && iy < p.h && ix >= 0 && ix < p.w;
          raw[u] = valid[u] ? lds64(tile_a + (v + 256 * u) * 8) : make_uint2(0u, 0u);
        }
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          if (v + 256 * u >= kElems) break;
          uint2 out = make_uint2(0u, 0u);
          if (valid[u]) {
            float2 v0 = __ffma2_rn(unpack_bf16x2(raw[u].x), sc[0], sh[0]), v1 = __ffma2_rn(unpack_bf16x2(raw[u].y), sc[1], sh[1]);
            if (relu) { v0 = relu2(v0); v1 = relu2(v1); }
            out = make_uint2(pack_bf16x2(v0), pack_bf16x2(v1));
          }
          sts64(tile_a + (v + 256 * u) * 8, out);
        }
      }
      __syncthreads();
    }

    if (ch_ok && ox < p.w) {
      const uint32_t wa = tile_a + thr_off;
      float2 win[3][3][2];
      auto load_row = [&](int slot, uint32_t ra) {
#pragma unroll
        for (int kw = 0; kw < 3; ++kw) {
          load4(ra + kw * 128, win[slot][kw]);
          if (!AFFINE && relu) { win[slot][kw][0] = relu2(win[slot][kw][0]); win[slot][kw][1] = relu2(win[slot][kw][1]); }
        }
      };
      load_row(0, wa);
      load_row(1, wa + kFRow);
      uint32_t ra = wa + 2 * kFRow;
      __nv_bfloat16* out = dst + (((size_t)img * p.h + oy0) * p.w + ox) * p.c + c0;
      const size_t out_row = (size_t)p.w * p.c;
#pragma unroll 1
      for (int r3 = 0; r3 < kFY; r3 += 3) {
        if (oy0 + r3 >= p.h) break;  // the rest of this tile lies below the image
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          load_row((j + 2) % 3, ra);
          ra += kFRow;
          float2 acc[2] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f)};
#pragma unroll
          for (int kh = 0; kh < 3; ++kh)
#pragma unroll
            for (int kw = 0; kw < 3; ++kw) {
              float2 wk[2];
              if (WSMEM) lds_w4(w_thr + (kh * 3 + kw) * 256, wk);
              else { wk[0] = wreg[WSMEM ? 0 : kh * 3 + kw][0]; wk[1] = wreg[WSMEM ? 0 : kh * 3 + kw][1]; }
#pragma unroll
              for (int e = 0; e < 2; ++e) acc[e] = __ffma2_rn(win[(j + kh) % 3][kw][e], wk[e], acc[e]);
            }
          if (oy0 + r3 + j < p.h) {
            uint32_t pk[2];
#pragma unroll
            for (int e = 0; e < 2; ++e) {
              pk[e] = pack_bf16x2(acc[e]);
              const float2 ab = unpack_bf16x2(pk[e]);
              s12[0][e] = __fadd2_rn(s12[0][e], ab);
              s12[1][e] = __ffma2_rn(ab, ab, s12[1][e]);
            }
            *reinterpret_cast<uint2*>(out) = make_uint2(pk[0], pk[1]);
          }
          out += out_row;
        }
      }
    }
    __syncthreads();
  }

  if (stats) reduce4_to_global<2>(s12, reinterpret_cast<float*>(smem), t, cq, cchunk, p.c, stats);
}

// ------------------------------------------------------------------------------------------ backward
// The same 3x3 window of the incoming gradient dd feeds BOTH gradients:
//     g[q]        = sum_{i,j} dd[q + (i-1, j-1)] * w[2-i][2-j]          (data gradient, flipped taps)
//     dw[2-i][2-j] += dd[q + (i-1, j-1)] * xin[q]                        (weight gradient in scatter form)
// so per output element there are 3 shared-memory loads of dd, one of x, and 18 FMAs - no second warpgroup
// re-reading the tile, no role imbalance at the tile barrier.
// SIDE: the incoming gradient is not materialised - a pre-pass assembles it in place from the pointwise conv's data
// gradient e and the depthwise output d:  dd = e + negk*d + kmean  (bn1's backward, see sepconv.cu), zero outside
// the image (TMA's zero fill covers the plain case).
template <bool AFFINE, bool SIDE, int FY, int MINB, bool WSMEM>
__global__ void __launch_bounds__(256, MINB) dwf_bwd_kernel(const __grid_constant__ CUtensorMap tmap_dd,
                                                         const __grid_constant__ CUtensorMap tmap_d,
                                                         const __grid_constant__ CUtensorMap tmap_x,
                                                         const float* __restrict__ w9c, const float* __restrict__ in_scale,
                                                         const float* __restrict__ in_shift, const float* __restrict__ negk,
                                                         const float* __restrict__ kmean,
                                                         const __nv_bfloat16* __restrict__ addend,
                                                         __nv_bfloat16* __restrict__ gout, double* __restrict__ dw_out,
                                                         double* __restrict__ sums, DwFParams p) {
  pdl_trigger();
  pdl_wait();
  constexpr int kFY = FY, kFTile = ftile_bytes(FY);
  static_assert(FY % 3 == 0, "the row loop is unrolled over the 3-row register window");
  constexpr int kTiles = SIDE ? 3 : 2;
  constexpr int kStage = kTiles * kFTile;
  constexpr int kStages = SIDE ? 2 : 3;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~uintptr_t(127));
  const uint32_t smem_base = smem_u32(smem), bar0 = smem_base + kStages * kStage;

  const int t = threadIdx.x;
  const int cq = t & 15, col = t >> 4;   // channel quad inside the 64-channel chunk, output column inside the tile
  const int cchunk = blockIdx.y;
  const int c0 = cchunk * 64 + cq * 4;
  const bool ch_ok = c0 < p.c;           // C % 8 == 0: quads are whole
  const bool relu = p.relu_in != 0;
  const uint32_t thr_off = col * 128 + cq * 8;

  float2 wf[WSMEM ? 1 : 9][2], acc[9][2], sc[2], sh[2], nk[2], km[2], s12[2][2];
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int c = c0 + 2 * i;
    sc[i] = (AFFINE && ch_ok) ? make_float2(__ldg(in_scale + c), __ldg(in_scale + c + 1)) : make_float2(1.f, 1.f);
    sh[i] = (AFFINE && ch_ok) ? make_float2(__ldg(in_shift + c), __ldg(in_shift + c + 1)) : make_float2(0.f, 0.f);
    nk[i] = (SIDE && ch_ok) ? make_float2(__ldg(negk + c), __ldg(negk + c + 1)) : make_float2(0.f, 0.f);
    km[i] = (SIDE && ch_ok) ? make_float2(__ldg(kmean + c), __ldg(kmean + c + 1)) : make_float2(0.f, 0.f);
    s12[0][i] = s12[1][i] = make_float2(0.f, 0.f);
  }
  // window position k = (i_row, j_col) pairs with the flipped tap 8-k.  The 9 x 64 flipped weights of the chunk live in
  // registers (36 per thread) or, WSMEM, once in shared memory [position][channel] and are re-read every row
  const uint32_t wsm = bar0 + 64, w_thr = wsm + cq * 16;
  if (WSMEM) {
    float* wdst = reinterpret_cast<float*>(smem + kStages * kStage + 64);
    for (int i = t; i < 9 * 64; i += 256) {
      const int c = cchunk * 64 + (i & 63);
      wdst[i] = c < p.c ? __ldg(w9c + (8 - (i >> 6)) * p.c + c) : 0.f;
    }
    if (AFFINE && t < 128) {          // the previous BatchNorm's scale | shift of the chunk, rows 9 and 10 of the table
      const int c = cchunk * 64 + (t & 63);
      wdst[9 * 64 + t] = c < p.c ? __ldg((t < 64 ? in_scale : in_shift) + c) : (t < 64 ? 1.f : 0.f);
    }
  }
#pragma unroll
  for (int k = 0; k < 9; ++k)
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      if (!WSMEM)
        wf[WSMEM ? 0 : k][i] = ch_ok ? make_float2(__ldg(w9c + (8 - k) * p.c + c0 + 2 * i), __ldg(w9c + (8 - k) * p.c + c0 + 2 * i + 1))
                                     : make_float2(0.f, 0.f);
      acc[k][i] = make_float2(0.f, 0.f);
    }

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
    int tx, ty, img;
    tile_coords(p, blockIdx.x + i * gridDim.x, tx, ty, img);
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
    const uint32_t dd_a = smem_base + s * kStage;
    int tx, ty, img;
    tile_coords(p, blockIdx.x + i * gridDim.x, tx, ty, img);
    const int ox = tx * kFX + col, oy0 = ty * kFY;

    if (SIDE) {
      // ---- pre-pass: dd = e + negk*d + kmean over the halo tile, in place, zero outside the image; two elements per
      // thread in flight (element index advances by 256 threads = 16 pixels: the channel quad stays cq)
      constexpr int kElems = (kFY + 2) * (kFX + 2) * 16;
      const int ox0 = tx * kFX - 1, oyb = oy0 - 1;
      for (int v = t; v < kElems; v += 512) {
        uint2 e[2], d[2];
        bool valid[2];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          const int pix = (v + 256 * u) >> 4;
          const int rr = pix / (kFX + 2), cc = pix - rr * (kFX + 2);
          const int iy = oyb + rr, ix = ox0 + cc;
          valid[u] = v + 256 * u < kElems && iy >= 0 && iy < p.h && ix >= 0 && ix < p.w;
          e[u] = d[u] = make_uint2(0u, 0u);
          if (valid[u]) {
            e[u] = lds64(dd_a + (v + 256 * u) * 8);
            d[u] = lds64(dd_a + 2 * kFTile + (v + 256 * u) * 8);
          }
        }
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          if (v + 256 * u >= kElems) break;
          uint2 out = make_uint2(0u, 0u);
          if (valid[u]) {
            out.x = pack_bf16x2(__ffma2_rn(nk[0], unpack_bf16x2(d[u].x), __fadd2_rn(unpack_bf16x2(e[u].x), km[0])));
            out.y = pack_bf16x2(__ffma2_rn(nk[1], unpack_bf16x2(d[u].y), __fadd2_rn(unpack_bf16x2(e[u].y), km[1])));
          }
          sts64(dd_a + (v + 256 * u) * 8, out);
        }
      }
      __syncthreads();
    }

    if (ch_ok && ox < p.w) {
      const uint32_t wa = dd_a + thr_off;
      float2 win[3][3][2];
      auto load_row = [&](int slot, uint32_t ra) {
#pragma unroll
        for (int kw = 0; kw < 3; ++kw) load4(ra + kw * 128, win[slot][kw]);
      };
      load_row(0, wa);
      load_row(1, wa + kFRow);
      uint32_t ra = wa + 2 * kFRow;                         // next window row
      uint32_t xa = wa + kFTile + kFRow + 128;              // x at the output pixel (row r+1, column col+1)
      size_t off = (((size_t)img * p.h + oy0) * p.w + ox) * p.c + c0;
      const size_t out_row = (size_t)p.w * p.c;
      // the addend (gradient of the skip branch) is the one operand that comes straight from global memory: it is
      // fetched a whole 3-row group ahead, so its latency hides behind ~1 000 cycles of FMAs instead of one row's worth
      // (measured: the block-input launches took 114 us against 65 us for the same kernel without an addend)
      uint2 add_cur[3], add_nxt[3];
      auto load_add = [&](uint2 (&dst)[3], size_t o, int row0) {
#pragma unroll
        for (int j = 0; j < 3; ++j)
          dst[j] = (addend && oy0 + row0 + j < p.h) ? __ldg(reinterpret_cast<const uint2*>(addend + o + j * out_row))
                                                   : make_uint2(0u, 0u);
      };
      load_add(add_cur, off, 0);
#pragma unroll 1
      for (int r3 = 0; r3 < kFY; r3 += 3) {
        if (oy0 + r3 >= p.h) break;  // the rest of this tile lies below the image
        if (r3 + 3 < kFY) load_add(add_nxt, off + 3 * out_row, r3 + 3);
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          const bool row_ok = oy0 + r3 + j < p.h;
          const uint2 add_raw = add_cur[j];
          load_row((j + 2) % 3, ra);
          ra += kFRow;
          float2 xc[2], xin[2], g[2], scr[2], shr[2];
          load4(xa, xc);
          xa += kFRow;
          if (AFFINE && WSMEM) { lds_w4(w_thr + 9 * 256, scr); lds_w4(w_thr + 10 * 256, shr); }
          else { scr[0] = sc[0]; scr[1] = sc[1]; shr[0] = sh[0]; shr[1] = sh[1]; }
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            xin[e] = AFFINE ? __ffma2_rn(xc[e], scr[e], shr[e]) : xc[e];
            if (relu) xin[e] = relu2(xin[e]);
            if (!row_ok) xin[e] = make_float2(0.f, 0.f);   // rows below the image feed nothing
            g[e] = make_float2(0.f, 0.f);
          }
#pragma unroll
          for (int kh = 0; kh < 3; ++kh)
#pragma unroll
            for (int kw = 0; kw < 3; ++kw) {
              float2 wk[2];
              if (WSMEM) lds_w4(w_thr + (kh * 3 + kw) * 256, wk);
              else { wk[0] = wf[WSMEM ? 0 : kh * 3 + kw][0]; wk[1] = wf[WSMEM ? 0 : kh * 3 + kw][1]; }
#pragma unroll
              for (int e = 0; e < 2; ++e) {
                const float2 d = win[(j + kh) % 3][kw][e];
                g[e] = __ffma2_rn(d, wk[e], g[e]);
                acc[kh * 3 + kw][e] = __ffma2_rn(d, xin[e], acc[kh * 3 + kw][e]);
              }
            }
          if (row_ok) {
            uint32_t pk[2];
#pragma unroll
            for (int e = 0; e < 2; ++e) {
              if (relu) {
                g[e].x = xin[e].x > 0.f ? g[e].x : 0.f;
                g[e].y = xin[e].y > 0.f ? g[e].y : 0.f;
              }
              // the BatchNorm reduction uses g as the apply kernel will read it back (bf16), without the addend
              uint32_t u = pack_bf16x2(g[e]);
              const float2 ab = unpack_bf16x2(u);
              s12[0][e] = __fadd2_rn(s12[0][e], ab);
              s12[1][e] = __ffma2_rn(ab, xc[e], s12[1][e]);
              if (addend) u = pack_bf16x2(__fadd2_rn(g[e], unpack_bf16x2(e == 0 ? add_raw.x : add_raw.y)));
              pk[e] = u;
            }
            *reinterpret_cast<uint2*>(gout + off) = make_uint2(pk[0], pk[1]);
          }
          off += out_row;
        }
#pragma unroll
        for (int j = 0; j < 3; ++j) add_cur[j] = add_nxt[j];
      }
    }
    __syncthreads();  // everyone is done with slot s before it is refilled
  }

  // acc[k] pairs window position k with tap 8-k
  float2 dwp[9][2];
#pragma unroll
  for (int k = 0; k < 9; ++k) { dwp[k][0] = acc[8 - k][0]; dwp[k][1] = acc[8 - k][1]; }
  float* red = reinterpret_cast<float*>(smem);
  reduce4_to_global<9>(dwp, red, t, cq, cchunk, p.c, dw_out);
  if (sums) reduce4_to_global<2>(s12, red, t, cq, cchunk, p.c, sums);
}

static bool dwf_supported(const cvx_conv_desc* d) {
  return d->dtype == CVX_BF16 && d->stride == 1 && d->dil == 1 && d->pad == 1 && d->kh == 3 && d->kw == 3 &&
         d->cin % 8 == 0 && d->cin == d->cout;
}

static void dwf_params(const cvx_conv_desc* d, int relu_in, int fy, DwFParams* p) {
  p->n = d->n; p->h = d->h; p->w = d->w; p->c = d->cin;
  p->tiles_x = (d->w + kFX - 1) / kFX;
  p->tiles_y = (d->h + fy - 1) / fy;
  p->ntiles = d->n * p->tiles_x * p->tiles_y;
  p->relu_in = relu_in;
}

constexpr int kWsmBytes = 11 * 64 * 4;  // the chunk's tap weights (+ BatchNorm scale, shift) in shared memory (WSMEM variants)

template <bool AFFINE, int FY, int MINB, bool WSMEM>
static int dwf_fwd_launch_t(const cvx_conv_desc* d, const void* x, const float* w9c, const float* in_scale,
                            const float* in_shift, int relu_in, void* y, double* stats, cudaStream_t st) {
  DwFParams p;
  dwf_params(d, relu_in, FY, &p);
  CUtensorMap map;
  if (int rc = make_act_map(&map, x, d->n, d->h, d->w, d->cin, kFX + 2, FY + 2, CU_TENSOR_MAP_SWIZZLE_NONE)) return rc;
  const int chunks = (d->cin + 63) / 64;
  int gx = (kNumSMs * MINB) / chunks;   // MINB 256-thread CTAs per SM
  if (gx < 1) gx = 1;
  if (gx > p.ntiles) gx = p.ntiles;
  constexpr int smem = kFwdStages * ftile_bytes(FY) + 128 + 64 + (WSMEM ? kWsmBytes : 0);
  static bool configured = false;
  if (!configured) {
    CVX_CUDA_OK(cudaFuncSetAttribute(dwf_fwd_kernel<AFFINE, FY, MINB, WSMEM>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured = true;
  }
  launch_pdl(dwf_fwd_kernel<AFFINE, FY, MINB, WSMEM>, dim3(gx, chunks), dim3(256), smem, st, map, w9c, in_scale, in_shift,
             (__nv_bfloat16*)y, stats, p);
  CVX_LAUNCH_OK();
  return CVX_OK;
}

int dwf_fwd_launch(const cvx_conv_desc* d, const void* x, const float* w9c, const float* in_scale, const float* in_shift,
                   int relu_in, void* y, double* stats, cudaStream_t st) {
  if (!dwf_supported(d)) return CVX_EUNSUPPORTED;
  if (stats) CVX_WS_ZERO(stats, sizeof(double) * 2 * d->cin, st);
  if (g_dwf_variant == 1) {
    if (in_scale) return dwf_fwd_launch_t<true, 12, 2, false>(d, x, w9c, in_scale, in_shift, relu_in, y, stats, st);
    return dwf_fwd_launch_t<false, 12, 2, false>(d, x, w9c, nullptr, nullptr, relu_in, y, stats, st);
  }
  if (g_dwf_variant == 2) {
    if (in_scale) return dwf_fwd_launch_t<true, 6, 3, false>(d, x, w9c, in_scale, in_shift, relu_in, y, stats, st);
    return dwf_fwd_launch_t<false, 6, 3, false>(d, x, w9c, nullptr, nullptr, relu_in, y, stats, st);
  }
  if (g_dwf_variant == 3) {   // short tiles alone (no register cap, same residency as the tall kernel)
    if (in_scale) return dwf_fwd_launch_t<true, 6, 2, false>(d, x, w9c, in_scale, in_shift, relu_in, y, stats, st);
    return dwf_fwd_launch_t<false, 6, 2, false>(d, x, w9c, nullptr, nullptr, relu_in, y, stats, st);
  }
  if (g_dwf_variant == 4) {   // weights in shared memory alone (tall tiles)
    if (in_scale) return dwf_fwd_launch_t<true, 12, 2, true>(d, x, w9c, in_scale, in_shift, relu_in, y, stats, st);
    return dwf_fwd_launch_t<false, 12, 2, true>(d, x, w9c, nullptr, nullptr, relu_in, y, stats, st);
  }
  if (in_scale) return dwf_fwd_launch_t<true, 6, 3, true>(d, x, w9c, in_scale, in_shift, relu_in, y, stats, st);
  return dwf_fwd_launch_t<false, 6, 3, true>(d, x, w9c, nullptr, nullptr, relu_in, y, stats, st);
}

template <bool AFFINE, bool SIDE, int FY, int MINB, bool WSMEM>
static int dwf_bwd_launch_t(const cvx_conv_desc* d, const void* dd, const void* dside, const void* x, const float* w9c,
                            const float* in_scale, const float* in_shift, const float* negk, const float* kmean, int relu_in,
                            const void* addend, void* g, double* dw_out, double* sums, cudaStream_t st) {
  DwFParams p;
  dwf_params(d, relu_in, FY, &p);
  CUtensorMap mdd, md, mx;
  if (int rc = make_act_map(&mdd, dd, d->n, d->h, d->w, d->cin, kFX + 2, FY + 2, CU_TENSOR_MAP_SWIZZLE_NONE)) return rc;
  if (int rc = make_act_map(&mx, x, d->n, d->h, d->w, d->cin, kFX + 2, FY + 2, CU_TENSOR_MAP_SWIZZLE_NONE)) return rc;
  md = mx;
  if (dside)
    if (int rc = make_act_map(&md, dside, d->n, d->h, d->w, d->cin, kFX + 2, FY + 2, CU_TENSOR_MAP_SWIZZLE_NONE)) return rc;
  const int chunks = (d->cin + 63) / 64;
  int gx = (kNumSMs * MINB) / chunks;  // MINB 256-thread CTAs per SM
  if (gx < 1) gx = 1;
  if (gx > p.ntiles) gx = p.ntiles;
  const dim3 grid(gx, chunks);
  constexpr int smem = (SIDE ? 2 * 3 : 3 * 2) * ftile_bytes(FY) + 128 + 64 + (WSMEM ? kWsmBytes : 0);
  static bool configured = false;
  if (!configured) {
    CVX_CUDA_OK(cudaFuncSetAttribute(dwf_bwd_kernel<AFFINE, SIDE, FY, MINB, WSMEM>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured = true;
  }
  launch_pdl(dwf_bwd_kernel<AFFINE, SIDE, FY, MINB, WSMEM>, grid, dim3(256), smem, st, mdd, md, mx, w9c, in_scale, in_shift, negk,
             kmean, (const __nv_bfloat16*)addend, (__nv_bfloat16*)g, dw_out, sums, p);
  CVX_LAUNCH_OK();
  return CVX_OK;
}

template <int FY, int MINB, bool WSMEM>
static int dwf_bwd_dispatch(const cvx_conv_desc* d, const void* dd, const void* dside, const float* negk, const float* kmean,
                            const void* x, const float* w9c, const float* in_scale, const float* in_shift, int relu_in,
                            const void* addend, void* g, double* dw_out, double* sums, cudaStream_t st) {
  if (dside) {
    if (in_scale) return dwf_bwd_launch_t<true, true, FY, MINB, WSMEM>(d, dd, dside, x, w9c, in_scale, in_shift, negk, kmean, relu_in, addend, g, dw_out, sums, st);
    return dwf_bwd_launch_t<false, true, FY, MINB, WSMEM>(d, dd, dside, x, w9c, nullptr, nullptr, negk, kmean, relu_in, addend, g, dw_out, sums, st);
  }
  if (in_scale) return dwf_bwd_launch_t<true, false, FY, MINB, WSMEM>(d, dd, nullptr, x, w9c, in_scale, in_shift, nullptr, nullptr, relu_in, addend, g, dw_out, sums, st);
  return dwf_bwd_launch_t<false, false, FY, MINB, WSMEM>(d, dd, nullptr, x, w9c, nullptr, nullptr, nullptr, nullptr, relu_in, addend, g, dw_out, sums, st);
}

int dwf_bwd_launch(const cvx_conv_desc* d, const void* dd, const void* dside, const float* negk, const float* kmean,
                   const void* x, const float* w9c, const float* in_scale, const float* in_shift, int relu_in,
                   const void* addend, void* g, double* dw_out, double* sums, cudaStream_t st) {
  if (!dwf_supported(d)) return CVX_EUNSUPPORTED;
  CVX_WS_ZERO(dw_out, sizeof(double) * 9 * d->cin, st);
  if (sums) CVX_WS_ZERO(sums, sizeof(double) * 2 * d->cin, st);
  if (g_dwf_variant == 2)
    return dwf_bwd_dispatch<6, 2, false>(d, dd, dside, negk, kmean, x, w9c, in_scale, in_shift, relu_in, addend, g, dw_out, sums, st);
  if (g_dwf_variant == 3)
    return dwf_bwd_dispatch<6, 1, false>(d, dd, dside, negk, kmean, x, w9c, in_scale, in_shift, relu_in, addend, g, dw_out, sums, st);
  if (g_dwf_variant == 4)
    return dwf_bwd_dispatch<12, 1, true>(d, dd, dside, negk, kmean, x, w9c, in_scale, in_shift, relu_in, addend, g, dw_out, sums, st);
  if (g_dwf_variant == 1)
    return dwf_bwd_dispatch<12, 1, false>(d, dd, dside, negk, kmean, x, w9c, in_scale, in_shift, relu_in, addend, g, dw_out, sums, st);
  return dwf_bwd_dispatch<6, 2, true>(d, dd, dside, negk, kmean, x, w9c, in_scale, in_shift, relu_in, addend, g, dw_out, sums, st);
}

}  // namespace cvx
