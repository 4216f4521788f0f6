// Dense convolution as a tcgen05 / TMEM / TMA implicit GEMM (bf16 in, fp32 accumulate), sm_100a.
//
//   forward / dgrad : D[128 pixels][BN out-ch] += A[128 px][64 ch] * W[BN][64 ch]^T  per (tap, 64-ch block)
//       A is a 4-D TMA box (64 ch, BW, BH, 1 image) of the NHWC activation whose start
//       coordinate is shifted by the tap offset (kh*dil - pad, kw*dil - pad); TMA zero-fills
//       out-of-bounds elements, which IS the convolution's zero padding and also pads the
//       channel tails (728 = 11*64 + 24).  Both operands are K-major, SWIZZLE_128B.
//   wgrad           : D[128 co][BN ci] += dY[64 px][128 co]^T * X[64 px][BN ci]  per pixel tile
//       both operands are MN-major (pixel = K is the slow smem axis), split over pixel tiles,
//       accumulated into fp32 with red.global.add.
//
// One CTA = one accumulator tile; 4 warps: warp0 lane0 = TMA producer, warp1 lane0 = MMA
// issuer, warp2 = TMEM allocator, then all 4 warps drain TMEM (warp w owns lanes 32w..32w+31).
// Two CTAs are co-resident per SM (3 x 32 KB stages each) so one CTA's epilogue overlaps the
// other's main loop.
#include <cuda.h>
#include <stdlib.h>

#include "tma_utils.cuh"

namespace cvx {

// ------------------------------------------------------------------------------ PTX wrappers (tcgen05)
__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on an mbarrier once all previously issued tcgen05.mma of this thread have completed
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred != 0;
}

// UMMA shared-memory descriptor, SWIZZLE_128B, sm_100 version field = 1
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;  // descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;  // LayoutType::SWIZZLE_128B
  return d;
}
// kind::f16 instruction descriptor: bf16 x bf16 -> fp32
__host__ __device__ constexpr uint32_t make_idesc(int M, int N, int a_mn_major, int b_mn_major) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16) |
         ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// ------------------------------------------------------------------------------ fwd / dgrad
constexpr int kStages = 3;
constexpr int kABytes = 128 * 128;  // 128 rows x 64 bf16

struct TcFwdParams {
  int n, ho, wo, cout;       // output geometry (rows of the GEMM)
  int cin;                   // reduction channels
  int kh, kw, pad, dil;      // taps; input coordinate = out*stride - pad + k*dil
  int bw, bh;                // pixel tile (bw*bh == 128)
  int tiles_x, tiles_y;
  int stride;                // 1, or 2 (persistent forward kernel only: TMA element strides)
  int tile_n;                // output-channel tile (multiple of 16, <= 256), balanced over n_tiles
  int trace;                 // tools only (cvx_debug_tc_trace): the first and the last cluster leave clock64 marks
};

// phase marks of the CTA-pair kernel for tools/tc_trace.py: [0] first cluster, [64] last cluster.  Slots: 0 globaltimer
// at entry, 1 clock at entry, 2 set-up done, 3 dependency wait done, 4 first / 5 last TMA issue, 6 first operands landed,
// 8+t MMAs of tile t committed, 24+t accumulator of tile t seen by the epilogue, 40+t epilogue of tile t done,
// 56 statistics flushed, 57 clock at exit, 58 globaltimer at exit.
__device__ unsigned long long g_tc_trace[128];
__device__ __forceinline__ unsigned long long gtimer() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
#define TC_MARK(slot) do { if (tr) tr[slot] = (unsigned long long)clock64(); } while (0)

// Optional epilogue work of the forward / data-gradient kernels (persistent variants):
//   y[row][c] = acc + bias[c] + side_scale[c] * side[row][c]          (bias, side nullable)
//   stats[0][c] += sum_rows y ; stats[1][c] += sum_rows y^2           (of the bf16-rounded stored values; nullable)
// The side term turns the data gradient of a pointwise conv into the gradient through the BatchNorm that preceded
// it (sepconv.cu); the statistics feed the BatchNorm that follows the conv, saving one full read of y.
struct TcEpi {
  const float* bias;
  const __nv_bfloat16* side;
  const float* side_scale;
  double* stats;
  int relu;   // y = max(y, 0) after bias and side (inference: BatchNorm folded into the weights, residual as side)
};
constexpr int kEpiMaxNTiles = 8;
constexpr int kEpiStatsBytes = kEpiMaxNTiles * 256 * 2 * 4;  // per-CTA fp32 staging [n_tile][2][256]

// transpose-reduce over a warp: on return v[0] of lane L holds sum_{lanes} v[L]
__device__ __forceinline__ void warp_colsum32(float (&v)[32], int lane) {
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) {
    const bool up = (lane & off) != 0;
#pragma unroll
    for (int i = 0; i < off; ++i) {
      const float a = v[i], b = v[i + off];
      const float send = up ? a : b, keep = up ? b : a;
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
    }
  }
}

__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, 128;" ::: "memory"); }
__device__ __forceinline__ void epi_bar_sync256() { asm volatile("bar.sync 1, 256;" ::: "memory"); }
constexpr int kPairThreads = 320;   // CTA-pair forward kernel: producer warp + issuer warp + 8 epilogue warps

// Drain this warp's 32 rows x tile_n columns of one accumulator: TMEM -> (+bias, +side) -> bf16 -> global.
__device__ __forceinline__ void tc_epilogue_tile(const TcEpi& ep, uint32_t t_addr, int tile_n, int n0, int cout,
                                                 bool row_ok, __nv_bfloat16* yrow, const __nv_bfloat16* srow,
                                                 float* stats_sm, int lane) {
#pragma unroll 1
  for (int c0 = 0; c0 < tile_n; c0 += 32) {
    if (n0 + c0 >= cout) break;  // warp-uniform
    uint4 sv[4];
    if (ep.side) {
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        const int c = c0 + g * 8;
        sv[g] = (row_ok && n0 + c < cout && c < tile_n) ? __ldg(reinterpret_cast<const uint4*>(srow + c))
                                                        : make_uint4(0, 0, 0, 0);
      }
    }
    uint32_t r[32];
    tmem_ld32(t_addr + (uint32_t)c0, r);
    tmem_ld_wait();
    float vals[32];
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      const int c = c0 + g * 8;
      const bool ok = row_ok && n0 + c < cout && c < tile_n;
      const uint32_t su[4] = {sv[g].x, sv[g].y, sv[g].z, sv[g].w};
      uint32_t packed[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float v0 = __uint_as_float(r[g * 8 + 2 * j]), v1 = __uint_as_float(r[g * 8 + 2 * j + 1]);
        if (ok) {
          if (ep.bias) { v0 += __ldg(ep.bias + n0 + c + 2 * j); v1 += __ldg(ep.bias + n0 + c + 2 * j + 1); }
          if (ep.side) {
            v0 = fmaf(__ldg(ep.side_scale + n0 + c + 2 * j), __uint_as_float(su[j] << 16), v0);
            v1 = fmaf(__ldg(ep.side_scale + n0 + c + 2 * j + 1), __uint_as_float(su[j] & 0xffff0000u), v1);
          }
        }
        __nv_bfloat162 h = __floats2bfloat162_rn(v0, v1);
        packed[j] = *reinterpret_cast<uint32_t*>(&h);
        vals[g * 8 + 2 * j] = ok ? __uint_as_float(packed[j] << 16) : 0.f;
        vals[g * 8 + 2 * j + 1] = ok ? __uint_as_float(packed[j] & 0xffff0000u) : 0.f;
      }
      if (ok) *reinterpret_cast<uint4*>(yrow + c) = make_uint4(packed[0], packed[1], packed[2], packed[3]);
    }
    if (ep.stats) {
      float sq[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) sq[i] = vals[i] * vals[i];
      warp_colsum32(vals, lane);
      warp_colsum32(sq, lane);
      const int c = c0 + lane;
      // the four epilogue warps take turns (fixed order, no shared atomics); every warp of the CTA walks the same chunks
      for (int turn = 0; turn < 4; ++turn) {
        if (((threadIdx.x >> 5) & 3) == turn && n0 + c < cout && c < tile_n) {
          stats_sm[c] += vals[0];
          stats_sm[256 + c] += sq[0];
        }
        epi_bar_sync();
      }
    }
  }
}

// lean epilogue of the plain (bias-only) kernels
__device__ __forceinline__ void tc_epilogue_tile_plain(const float* __restrict__ bias, uint32_t t_addr, int tile_n, int n0,
                                                       int cout, bool row_ok, __nv_bfloat16* yrow, bool relu = false) {
#pragma unroll 1
  for (int c0 = 0; c0 < tile_n; c0 += 32) {
    if (n0 + c0 >= cout) break;  // warp-uniform
    uint32_t r[32];
    tmem_ld32(t_addr + (uint32_t)c0, r);
    tmem_ld_wait();
    if (row_ok) {
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        const int c = c0 + g * 8;
        if (n0 + c < cout && c < tile_n) {
          float bv[8];
          if (bias) {
            if ((reinterpret_cast<uintptr_t>(bias) & 15) == 0) {  // parameter views of a flat buffer may be unaligned
              const float4 b0 = __ldg(reinterpret_cast<const float4*>(bias + n0 + c));
              const float4 b1 = __ldg(reinterpret_cast<const float4*>(bias + n0 + c + 4));
              bv[0] = b0.x; bv[1] = b0.y; bv[2] = b0.z; bv[3] = b0.w; bv[4] = b1.x; bv[5] = b1.y; bv[6] = b1.z; bv[7] = b1.w;
            } else {
#pragma unroll
              for (int j = 0; j < 8; ++j) bv[j] = __ldg(bias + n0 + c + j);
            }
          }
          uint32_t packed[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            float v0 = __uint_as_float(r[g * 8 + 2 * j]), v1 = __uint_as_float(r[g * 8 + 2 * j + 1]);
            if (bias) { v0 += bv[2 * j]; v1 += bv[2 * j + 1]; }
            if (relu) { v0 = fmaxf(v0, 0.f); v1 = fmaxf(v1, 0.f); }
            __nv_bfloat162 h = __floats2bfloat162_rn(v0, v1);
            packed[j] = *reinterpret_cast<uint32_t*>(&h);
          }
          *reinterpret_cast<uint4*>(yrow + c) = make_uint4(packed[0], packed[1], packed[2], packed[3]);
        }
      }
    }
  }
}

// flush the per-CTA statistics staging to the fp64 accumulators (called by the 128 epilogue threads)
template <int NT = 128>
__device__ __forceinline__ void tc_epilogue_flush_stats(const TcEpi& ep, const float* stats_sm, int n_tiles, int tile_n,
                                                        int cout, int epi_tid) {
  if (NT == 256) epi_bar_sync256(); else epi_bar_sync();
  for (int i = epi_tid; i < n_tiles * 512; i += NT) {
    const float v = stats_sm[i];
    if (v == 0.f) continue;
    const int nt = i >> 9, which = (i >> 8) & 1, c = i & 255;
    const int col = nt * tile_n + c;
    if (c < tile_n && col < cout) atomicAdd(ep.stats + (size_t)which * cout + col, (double)v);
  }
}

template <int BN>
__global__ void __launch_bounds__(128) conv_tc_fwd_kernel(const __grid_constant__ CUtensorMap tmap_x,
                                                          const __grid_constant__ CUtensorMap tmap_w,
                                                          const float* __restrict__ bias,
                                                          __nv_bfloat16* __restrict__ y, TcFwdParams p) {
  constexpr int kBBytes = BN * 128;
  constexpr int kStageBytes = kABytes + kBBytes;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kStages * kStageBytes);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kStages + 1);
  const uint32_t smem_base = smem_u32(smem);
  const uint32_t full0 = smem_u32(bars), empty0 = full0 + 8 * kStages, tfull = full0 + 16 * kStages;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // tile coordinates
  int tile = blockIdx.x;
  const int tx = tile % p.tiles_x; tile /= p.tiles_x;
  const int ty = tile % p.tiles_y;
  const int img = tile / p.tiles_y;
  const int ox0 = tx * p.bw, oy0 = ty * p.bh;
  const int n0 = blockIdx.y * BN;
  const int kcb = (p.cin + 63) / 64;
  const int num_kb = p.kh * p.kw * kcb;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_x);
    tma_prefetch_desc(&tmap_w);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(full0 + 8 * s, 1);
      mbar_init(empty0 + 8 * s, 1);
    }
    mbar_init(tfull, 1);
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc(smem_u32(tmem_slot), BN);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      for (int kb = 0; kb < num_kb; ++kb) {
        const int s = kb % kStages;
        const uint32_t ph = (kb / kStages) & 1;
        mbar_wait(empty0 + 8 * s, ph ^ 1);
        const int tap = kb / kcb, cb = kb - tap * kcb;
        const int khi = tap / p.kw, kwi = tap - khi * p.kw;
        const uint32_t sa = smem_base + s * kStageBytes;
        mbar_expect_tx(full0 + 8 * s, kStageBytes);
        tma_load_4d(sa, &tmap_x, full0 + 8 * s, cb * 64, ox0 - p.pad + kwi * p.dil, oy0 - p.pad + khi * p.dil, img);
        tma_load_3d(sa + kABytes, &tmap_w, full0 + 8 * s, cb * 64, n0, tap);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc(128, BN, 0, 0);
      for (int kb = 0; kb < num_kb; ++kb) {
        const int s = kb % kStages;
        const uint32_t ph = (kb / kStages) & 1;
        mbar_wait(full0 + 8 * s, ph);
        tc_fence_after();
        const uint32_t sa = smem_base + s * kStageBytes;
        const uint32_t sb = sa + kABytes;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const uint64_t ad = make_smem_desc(sa + k * 32, 0, 1024);
          const uint64_t bd = make_smem_desc(sb + k * 32, 0, 1024);
          umma_bf16(tmem_base, ad, bd, idesc, (kb > 0 || k > 0) ? 1u : 0u);
        }
        umma_commit(empty0 + 8 * s);  // frees the smem stage when these MMAs retire
      }
      umma_commit(tfull);
    }
  }
  __syncwarp();

  // ---- epilogue: TMEM -> registers -> (bias) -> bf16 -> global ---------------------------
  mbar_wait(tfull, 0);
  tc_fence_after();
  const int row = warp * 32 + lane;
  const int oy = oy0 + row / p.bw, ox = ox0 + row % p.bw;
  const bool row_ok = oy < p.ho && ox < p.wo;
  __nv_bfloat16* yrow = y + (((size_t)img * p.ho + oy) * p.wo + ox) * p.cout + n0;
#pragma unroll 1
  for (int c0 = 0; c0 < BN; c0 += 32) {
    if (n0 + c0 >= p.cout) break;  // warp-uniform
    uint32_t r[32];
    tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, r);
    tmem_ld_wait();
    if (row_ok) {
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        const int c = c0 + g * 8;
        if (n0 + c < p.cout) {  // cout % 8 == 0: whole 8-column groups are in or out
          uint32_t packed[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            float v0 = __uint_as_float(r[g * 8 + 2 * j]), v1 = __uint_as_float(r[g * 8 + 2 * j + 1]);
            if (bias) { v0 += __ldg(bias + n0 + c + 2 * j); v1 += __ldg(bias + n0 + c + 2 * j + 1); }
            __nv_bfloat162 h = __floats2bfloat162_rn(v0, v1);
            packed[j] = *reinterpret_cast<uint32_t*>(&h);
          }
          *reinterpret_cast<uint4*>(yrow + c) = make_uint4(packed[0], packed[1], packed[2], packed[3]);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, BN);
}


// ------------------------------------------------------------------------------ fwd / dgrad, persistent
// One CTA per SM loops over output tiles (n-tile fastest, so the CTAs that run concurrently share
// the same activation tile through L2).  Warp roles: 0 = TMA producer, 1 = TMEM owner + MMA issuer,
// 2..5 = epilogue.  The accumulator is double-buffered in TMEM (2 x 256 columns): the epilogue of
// tile i overlaps the main loop of tile i+1, and the per-CTA prologue (barrier init, TMEM alloc,
// descriptor fetch, pipeline fill) is paid once per launch instead of once per tile.  BN = 256
// halves the shared-memory traffic per MMA relative to BN = 128 (A is re-read once per 256 columns).
constexpr int kPStages = 4;
constexpr int kPBN = 256;
constexpr int kPBBytes = kPBN * 128;
constexpr int kPStageBytes = kABytes + kPBBytes;


template <bool EXTRA>
__global__ void __launch_bounds__(192, 1) conv_tc_fwd_persistent_kernel(const __grid_constant__ CUtensorMap tmap_x,
                                                                        const __grid_constant__ CUtensorMap tmap_w,
                                                                        const TcEpi ep,
                                                                        __nv_bfloat16* __restrict__ y, TcFwdParams p,
                                                                        int n_tiles, int total_tiles) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kPStages * kPStageBytes);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kPStages + 4);
  const uint32_t smem_base = smem_u32(smem);
  const uint32_t full0 = smem_u32(bars), empty0 = full0 + 8 * kPStages;
  const uint32_t tfull0 = empty0 + 8 * kPStages, tempty0 = tfull0 + 16;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int kcb = (p.cin + 63) / 64;
  const int num_kb = p.kh * p.kw * kcb;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_x);
    tma_prefetch_desc(&tmap_w);
    for (int s = 0; s < kPStages; ++s) {
      mbar_init(full0 + 8 * s, 1);
      mbar_init(empty0 + 8 * s, 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull0 + 8 * a, 1);
      mbar_init(tempty0 + 8 * a, 4);  // one arrival per epilogue warp
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(smem_u32(tmem_slot), 2 * kPBN);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // warp-uniform producer / issuer loops, elect_one() around the TMA / tcgen05 instructions (see the pair kernel)
  if (warp == 0) {
    uint32_t s = 0, ph = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const int nt = tile % n_tiles;
      int mt = tile / n_tiles;
      const int tx = mt % p.tiles_x; mt /= p.tiles_x;
      const int ty = mt % p.tiles_y;
      const int img = mt / p.tiles_y;
      const int n0 = nt * p.tile_n;
      const int xb = tx * p.bw * p.stride - p.pad, yb = ty * p.bh * p.stride - p.pad;
      int cb = 0, kwi = 0, khi = 0;
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(empty0 + 8 * s, ph ^ 1);
        const uint32_t sa = smem_base + s * kPStageBytes;
        if (elect_one()) {
          mbar_expect_tx(full0 + 8 * s, kABytes + p.tile_n * 128);
          tma_load_4d(sa, &tmap_x, full0 + 8 * s, cb * 64, xb + kwi * p.dil, yb + khi * p.dil, img);
          tma_load_3d(sa + kABytes, &tmap_w, full0 + 8 * s, cb * 64, n0, khi * p.kw + kwi);
        }
        __syncwarp();
        if (++cb == kcb) { cb = 0; if (++kwi == p.kw) { kwi = 0; ++khi; } }
        if (++s == kPStages) { s = 0; ph ^= 1; }
      }
    }
  } else if (warp == 1) {
    uint32_t s = 0, ph = 0, tcount = 0;
    const int last_ksteps = (p.cin - (kcb - 1) * 64 >= 64) ? 4 : (p.cin - (kcb - 1) * 64 + 15) / 16;
    const uint64_t desc_hi = ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++tcount) {
      const int nt = tile % n_tiles;
      const int n0 = nt * p.tile_n;
      int n_eff = p.cout - n0;
      n_eff = n_eff > p.tile_n ? p.tile_n : ((n_eff + 15) & ~15);
      const uint32_t idesc = make_idesc(128, n_eff, 0, 0);
      const int acc = tcount & 1;
      mbar_wait(tempty0 + 8 * acc, ((tcount >> 1) & 1) ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * kPBN;
      int cb = 0;
      uint32_t accum = 0;
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(full0 + 8 * s, ph);
        tc_fence_after();
        const int ksteps = (cb == kcb - 1) ? last_ksteps : 4;
        const uint32_t sa = smem_base + s * kPStageBytes;
        const uint64_t ad = desc_hi | (uint64_t)((sa >> 4) & 0x3FFF);
        const uint64_t bd = desc_hi | (uint64_t)(((sa + kABytes) >> 4) & 0x3FFF);
        if (elect_one()) {
          for (int kk = 0; kk < ksteps; ++kk) umma_bf16(d_tmem, ad + 2 * kk, bd + 2 * kk, idesc, kk ? 1u : accum);
          umma_commit(empty0 + 8 * s);
        }
        __syncwarp();
        accum = 1;
        if (++cb == kcb) cb = 0;
        if (++s == kPStages) { s = 0; ph ^= 1; }
      }
      if (elect_one()) umma_commit(tfull0 + 8 * acc);
      __syncwarp();
    }
  } else {
    const int lg = warp & 3;  // TMEM lane group this warp may access
    const int epi_tid = threadIdx.x - 64;
    float* stats_sm = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + 256);
    if (EXTRA && ep.stats) {
      for (int i = epi_tid; i < n_tiles * 512; i += 128) stats_sm[i] = 0.f;
      epi_bar_sync();
    }
    uint32_t tcount = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++tcount) {
      const int nt = tile % n_tiles;
      int mt = tile / n_tiles;
      const int tx = mt % p.tiles_x; mt /= p.tiles_x;
      const int ty = mt % p.tiles_y;
      const int img = mt / p.tiles_y;
      const int n0 = nt * p.tile_n;
      const int row = lg * 32 + lane;
      const int oy = ty * p.bh + row / p.bw, ox = tx * p.bw + row % p.bw;
      const bool row_ok = oy < p.ho && ox < p.wo;
      const size_t roff = (((size_t)img * p.ho + oy) * p.wo + ox) * p.cout + n0;
      const int acc = tcount & 1;
      mbar_wait(tfull0 + 8 * acc, (tcount >> 1) & 1);
      tc_fence_after();
      const uint32_t t_addr = tmem_base + acc * kPBN + ((uint32_t)(lg * 32) << 16);
      if (EXTRA)
        tc_epilogue_tile(ep, t_addr, p.tile_n, n0, p.cout, row_ok, y + roff, ep.side ? ep.side + roff : nullptr,
                         stats_sm + nt * 512, lane);
      else
        tc_epilogue_tile_plain(ep.bias, t_addr, p.tile_n, n0, p.cout, row_ok, y + roff, ep.relu != 0);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty0 + 8 * acc);
    }
    if (EXTRA && ep.stats) tc_epilogue_flush_stats(ep, stats_sm, n_tiles, p.tile_n, p.cout, epi_tid);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 2 * kPBN);
}


// ------------------------------------------------------------------------------ fwd / dgrad, CTA pair (cta_group::2)
// Two CTAs of a cluster (same TPC) form one 256-pixel x 256-channel tile: each CTA stages its own
// 128 pixels of A and HALF of the weight tile (128 rows), and the leader issues
// tcgen05.mma.cta_group::2 (M = 256).  Per SM and per MMA the shared-memory traffic is half of the
// single-CTA kernel's (A: 4 KB + B: 4 KB per 128x256x16 half-MMA), which lifts the smem-bandwidth
// ceiling that bounds the 1-CTA kernel at ~70 % of the tensor peak, and every weight tile is
// fetched from L2 once per pair instead of once per CTA.
constexpr int k2Stages = 5;
constexpr int k2BBytes = 128 * 128;                 // half of a 256-row weight tile
constexpr int k2StageBytes = kABytes + k2BBytes;    // 32 KB

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t map_to_cta(uint32_t local_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  // default (.release.cta) semantics on purpose: a .release.cluster arrive compiles to MEMBAR.ALL.GPU + ERRBAR
  // on every call, which serialised the peer's producer loop (seen in the ncu source view)
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void tma_load_4d_2sm(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1,
                                                int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d_2sm(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1,
                                                int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tmem_alloc_2sm(uint32_t dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem_relinquish_2sm() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_2sm(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_bf16_2sm(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                              uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// commit -> arrive on the barrier at this smem offset in BOTH CTAs of the pair
__device__ __forceinline__ void umma_commit_2sm(uint32_t bar, uint16_t cta_mask = 3) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
      ::"r"(bar), "h"(cta_mask)
      : "memory");
}
// weight tile shared by several CTA pairs of one cluster: ONE L2 read lands in every CTA of `cta_mask` (same smem
// offset) and completes bytes on the full barrier of each destination's pair leader
__device__ __forceinline__ void tma_load_3d_2sm_mc(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1,
                                                   int c2, uint16_t cta_mask) {
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster"
      " [%0], [%1, {%3, %4, %5}], [%2], %6;"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "h"(cta_mask)
      : "memory");
}

// ---- TMA store helpers (epilogue staging -> global, clipped to the tensor bounds by the hardware)
__device__ __forceinline__ void tma_store_4d(const CUtensorMap* map, uint32_t src, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
               ::"l"(map), "r"(src), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }

constexpr int kEpiBufBytes = 128 * 128;  // one 128-pixel x 64-channel bf16 chunk, SWIZZLE_128B rows

// 32 accumulator columns of one row -> (+bias, +side) -> 16 packed bf16 pairs (EXTRA: zero where the row or the
// column lies outside the tensor, so the staged chunk can be summed for the statistics as it is).
// shared-memory plan of the CTA-pair forward kernel per epilogue variant (227 KB per CTA): the side variants trade
// pipeline stages for a 4-deep side-input staging ring (their epilogue, not the main loop, is the critical path)
template <int MODE>
struct TcFwdSmem {
  static constexpr int kSideBufs = 4;                    // side-input staging ring: fetched 3 chunks ahead
  static constexpr int kStages = (MODE == 3) ? 3 : ((MODE & 2) ? 4 : 5);
  static constexpr int kBytes = kStages * k2StageBytes + 2 * kEpiBufBytes + ((MODE & 2) ? kSideBufs * kEpiBufBytes : 0) +
                                1024 + 256 + ((MODE & 1) ? kEpiStatsBytes : 0);
};

// MODE bit 0: BatchNorm statistics of the output, bit 1: side input (each epilogue variant carries only its own code)
// the 32 bias values of a thread's column group, fetched by the caller BETWEEN the issue of the TMEM load and its wait
// so that their latency hides behind the accumulator read (inside epi_convert32 they were exposed once per chunk)
__device__ __forceinline__ void epi_load_bias32(const float* __restrict__ bias, int col0, int cout, float (&bv)[32]) {
#pragma unroll
  for (int j = 0; j < 32; ++j) bv[j] = 0.f;
  if (bias == nullptr) return;
  if ((reinterpret_cast<uintptr_t>(bias) & 15) == 0) {  // parameter views of a flat buffer may be unaligned
#pragma unroll
    for (int g = 0; g < 8; ++g) {
      if (col0 + g * 4 < cout) {           // cout % 8 == 0: whole groups of 4 are in or out
        const float4 b = __ldg(reinterpret_cast<const float4*>(bias + col0 + g * 4));
        bv[4 * g] = b.x; bv[4 * g + 1] = b.y; bv[4 * g + 2] = b.z; bv[4 * g + 3] = b.w;
      }
    }
  } else {
#pragma unroll
    for (int j = 0; j < 32; ++j)
      if (col0 + j < cout) bv[j] = __ldg(bias + col0 + j);
  }
}

template <int MODE>
__device__ __forceinline__ void epi_convert32(const TcEpi& ep, const uint32_t (&r)[32], int col0, int cout, bool row_ok,
                                              const uint4 (&side_raw)[4], const float (&bias32)[32],
                                              const float (&sscale32)[32], uint32_t (&packed)[16]) {
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    const int c = col0 + g * 8;          // absolute output channel of this group of 8
    const bool ok = c < cout;             // cout % 8 == 0: whole groups are in or out
    float bv[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) bv[j] = bias32[g * 8 + j];
    float sd[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) sd[j] = 0.f;
    if ((MODE & 2) && ok && row_ok) {
      const uint4 sv = side_raw[g];   // prefetched by the caller ahead of the TMEM load (one L2 round trip per chunk)
      const uint32_t su[4] = {sv.x, sv.y, sv.z, sv.w};
      float ss[8];   // per-channel scale of the side input (1 for a plain residual add), prefetched by the caller
#pragma unroll
      for (int j = 0; j < 8; ++j) ss[j] = sscale32[g * 8 + j];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        sd[2 * j] = ss[2 * j] * __uint_as_float(su[j] << 16);
        sd[2 * j + 1] = ss[2 * j + 1] * __uint_as_float(su[j] & 0xffff0000u);
      }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float v0 = __uint_as_float(r[g * 8 + 2 * j]) + bv[2 * j] + sd[2 * j];
      float v1 = __uint_as_float(r[g * 8 + 2 * j + 1]) + bv[2 * j + 1] + sd[2 * j + 1];
      if (ep.relu) { v0 = fmaxf(v0, 0.f); v1 = fmaxf(v1, 0.f); }
      __nv_bfloat162 h = __floats2bfloat162_rn(v0, v1);
      const uint32_t u = *reinterpret_cast<uint32_t*>(&h);
      packed[g * 4 + j] = ((MODE & 1) && !(ok && row_ok)) ? 0u : u;   // rows / channels outside the tensor must not reach the statistics
    }
  }
}

// kPairs CTA pairs form one cluster (launch attribute, 2*kPairs CTAs): they work on the SAME output-channel tile of
// kPairs consecutive pixel-tile pairs, so the weight tile is read from L2 once per cluster - every CTA fetches
// 1/kPairs of its half and multicasts it to the CTAs that hold the same half in the other pairs.  The operand
// stream L2 -> SM is what bounds this kernel (32 KB per CTA and k-block without sharing, ~6.3 KB/clk chip-wide).
template <int MODE, int kPairs>
__global__ void __launch_bounds__(kPairThreads, 1)
conv_tc_fwd_2cta_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_w,
                        const __grid_constant__ CUtensorMap tmap_y, const __grid_constant__ CUtensorMap tmap_s,
                        const TcEpi ep, TcFwdParams p, int n_tiles, int m_tiles, int total_pair_tiles) {
  constexpr int kSt = TcFwdSmem<MODE>::kStages;
  pdl_trigger();   // the next kernel's CTAs may take an SM as soon as this kernel's CTA on it has exited
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* epi_buf = smem + kSt * k2StageBytes;          // 2 x 16 KB output staging, 1024-aligned
  constexpr int kSB = TcFwdSmem<MODE>::kSideBufs;
  uint8_t* side_buf = epi_buf + 2 * kEpiBufBytes;        // MODE & 2: kSB x 16 KB side-input staging (TMA destination)
  uint64_t* bars = reinterpret_cast<uint64_t*>(side_buf + ((MODE & 2) ? kSB * kEpiBufBytes : 0));
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kSt + 4 + kSB);
  const uint32_t smem_base = smem_u32(smem);
  const uint32_t full0 = smem_u32(bars), empty0 = full0 + 8 * kSt;
  const uint32_t tfull0 = empty0 + 8 * kSt, tempty0 = tfull0 + 16, sfull0 = tempty0 + 16;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t crank = cluster_ctarank();   // 0 .. 2*kPairs-1
  const uint32_t rank = crank & 1;            // rank inside the CTA pair
  const uint32_t pidx = crank >> 1;           // pair inside the cluster
  const uint32_t lead_rank = crank & ~1u;
  const bool leader = rank == 0;
  const int pair = blockIdx.x / (2 * kPairs), npairs = gridDim.x / (2 * kPairs);   // cluster index / count
  const int kcb = (p.cin + 63) / 64;
  const int num_kb = p.kh * p.kw * kcb;
  unsigned long long* tr = nullptr;
  if (p.trace && crank == 0 && (pair == 0 || pair == npairs - 1)) tr = g_tc_trace + (pair == 0 ? 0 : 64);
  if (tr && threadIdx.x == 0) { tr[0] = gtimer(); tr[1] = (unsigned long long)clock64(); }

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_x);
    tma_prefetch_desc(&tmap_w);
    tma_prefetch_desc(&tmap_y);
    if (MODE & 2) {
      tma_prefetch_desc(&tmap_s);
      for (int i = 0; i < kSB; ++i) mbar_init(sfull0 + 8 * i, 1);
    }
    for (int s = 0; s < kSt; ++s) {
      mbar_init(full0 + 8 * s, 2);   // leader's expect_tx arrival + the peer producer's arrival
      mbar_init(empty0 + 8 * s, kPairs);  // one multicast commit per pair leader of the cluster
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull0 + 8 * a, 1);
      mbar_init(tempty0 + 8 * a, 16);  // 8 epilogue warps x 2 CTAs (leader's copy is the one used)
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc_2sm(smem_u32(tmem_slot), 2 * kPBN);
    tmem_relinquish_2sm();
  }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (threadIdx.x == 0) TC_MARK(2);
  pdl_wait();      // barriers, TMEM and descriptors are set up: now wait for the producer of this kernel's inputs
  if (threadIdx.x == 0) TC_MARK(3);

  // The producer and the issuer warps run their loops WARP-UNIFORMLY (all 32 lanes compute the same counters and
  // addresses and wait on the same barriers); only the TMA / tcgen05 instructions themselves sit under elect_one().
  // With the loops inside an `if (lane == 0)` the compiler must treat every operand as divergent and wraps each MMA
  // in ~25 instructions of ELECT / R2UR.BROADCAST loops (seen in the ncu source view: ~125 instructions and
  // 700-900 cycles per 64-channel k-block on the single issuing thread - the bound of every launch of this kernel).
  if (warp == 0) {
    // ---- TMA producer: running counters only (no divisions inside the k loop)
    uint32_t s = 0, ph = 0;
    const uint32_t tx_bytes = 2 * (kABytes + (p.tile_n >> 1) * 128);
    const uint32_t lead_full0 = map_to_cta(full0, lead_rank);
    const int bq_rows = p.tile_n / (2 * kPairs);   // weight rows per multicast box
    uint16_t bmask = 0;
#pragma unroll
    for (int q = 0; q < kPairs; ++q) bmask |= (uint16_t)(1u << (2 * q + rank));
    bool first_issue = true;
    for (int pt = pair; pt < total_pair_tiles; pt += npairs) {
      const int nt = pt % n_tiles;
      const int mtile = 2 * ((pt / n_tiles) * kPairs + (int)pidx) + (int)rank;
      int tx = 0, ty = 0, img = p.n;  // img == n -> every TMA coordinate is out of bounds (zero fill)
      if (mtile < m_tiles) {
        int mt = mtile;
        tx = mt % p.tiles_x; mt /= p.tiles_x;
        ty = mt % p.tiles_y;
        img = mt / p.tiles_y;
      }
      const int n0 = nt * p.tile_n;
      int n_eff = p.cout - n0;
      n_eff = n_eff > p.tile_n ? p.tile_n : ((n_eff + 15) & ~15);
      const int brow0 = n0 + (int)rank * (n_eff >> 1) + (int)pidx * bq_rows;
      const int xb = tx * p.bw - p.pad, yb = ty * p.bh - p.pad;
      int cb = 0, kwi = 0, khi = 0;
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(empty0 + 8 * s, ph ^ 1);
        const uint32_t sa = smem_base + s * k2StageBytes;
        const uint32_t lead_full = (full0 + 8 * s) & 0xFEFFFFFFu;  // the leader CTA's barrier (peer bit cleared)
        if (tr && lane == 0) { if (first_issue) TC_MARK(4); TC_MARK(5); }
        first_issue = false;
        if (elect_one()) {
          if (leader) mbar_expect_tx(full0 + 8 * s, tx_bytes);
          else mbar_arrive_cluster(lead_full0 + 8 * s);
          tma_load_4d_2sm(sa, &tmap_x, lead_full, cb * 64, xb + kwi * p.dil, yb + khi * p.dil, img);
          if (kPairs == 1)
            tma_load_3d_2sm(sa + kABytes, &tmap_w, lead_full, cb * 64, brow0, khi * p.kw + kwi);
          else
            tma_load_3d_2sm_mc(sa + kABytes + pidx * bq_rows * 128, &tmap_w, lead_full, cb * 64, brow0, khi * p.kw + kwi, bmask);
        }
        __syncwarp();
        if (++cb == kcb) { cb = 0; if (++kwi == p.kw) { kwi = 0; ++khi; } }
        if (++s == kSt) { s = 0; ph ^= 1; }
      }
    }
  } else if (warp == 1) {
    if (leader) {
      // ---- MMA issuer (elect_one() picks the same lane every time, so all MMAs and commits come from one thread)
      uint32_t s = 0, ph = 0, tcount = 0;
      const int last_ksteps = (p.cin - (kcb - 1) * 64 >= 64) ? 4 : (p.cin - (kcb - 1) * 64 + 15) / 16;
      const uint64_t desc_hi = ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
      const uint16_t all_mask = (uint16_t)((1u << (2 * kPairs)) - 1), pair_mask = (uint16_t)(3u << (2 * pidx));
      for (int pt = pair; pt < total_pair_tiles; pt += npairs, ++tcount) {
        const int nt = pt % n_tiles;
        const int n0 = nt * p.tile_n;
        int n_eff = p.cout - n0;
        n_eff = n_eff > p.tile_n ? p.tile_n : ((n_eff + 15) & ~15);
        const uint32_t idesc = make_idesc(256, n_eff, 0, 0);
        const int acc = tcount & 1;
        mbar_wait(tempty0 + 8 * acc, ((tcount >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * kPBN;
        int cb = 0;
        uint32_t accum = 0;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(full0 + 8 * s, ph);
          tc_fence_after();
          if (tr && lane == 0 && tcount == 0 && kb == 0) TC_MARK(6);
          const uint32_t sa = smem_base + s * k2StageBytes;
          const uint64_t ad = desc_hi | (uint64_t)((sa >> 4) & 0x3FFF);
          const uint64_t bd = desc_hi | (uint64_t)(((sa + kABytes) >> 4) & 0x3FFF);
          if (elect_one()) {
            if (cb != kcb - 1 || last_ksteps == 4) {   // full 64-channel block: four K=16 steps, unrolled
              umma_bf16_2sm(d_tmem, ad, bd, idesc, accum);
              umma_bf16_2sm(d_tmem, ad + 2, bd + 2, idesc, 1u);
              umma_bf16_2sm(d_tmem, ad + 4, bd + 4, idesc, 1u);
              umma_bf16_2sm(d_tmem, ad + 6, bd + 6, idesc, 1u);
            } else {
              for (int kk = 0; kk < last_ksteps; ++kk) umma_bf16_2sm(d_tmem, ad + 2 * kk, bd + 2 * kk, idesc, kk ? 1u : accum);
            }
            umma_commit_2sm(empty0 + 8 * s, all_mask);
          }
          __syncwarp();
          accum = 1;
          if (++cb == kcb) cb = 0;
          if (++s == kSt) { s = 0; ph ^= 1; }
        }
        if (elect_one()) umma_commit_2sm(tfull0 + 8 * acc, pair_mask);
        __syncwarp();
        if (tr && lane == 0 && tcount < 16) TC_MARK(8 + tcount);
      }
    }
  } else {
    // ---- epilogue: TMEM -> registers -> bf16 -> swizzled smem chunk (128 px x 64 ch) -> TMA store.
    // EIGHT warps: warp w may touch TMEM lanes 32*(w%4)..+31 only, so warps 2-5 take columns 0-31 of every 64-column
    // chunk and warps 6-9 columns 32-63 of the same rows - two warps per scheduler instead of one (the epilogue of
    // short-K launches is issue-bound: ~300 instructions per thread and chunk with one warp per scheduler).
    const int lg = warp & 3;
    const int half = (warp - 2) >> 2;
    const uint32_t lead_tempty0 = map_to_cta(tempty0, lead_rank);
    const int epi_tid = threadIdx.x - 64;
    const int row = lg * 32 + lane;
    float* stats_sm = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + 256);
    if (MODE & 1) {
      for (int i = epi_tid; i < n_tiles * 512; i += 256) stats_sm[i] = 0.f;
      epi_bar_sync256();
    }
    // MODE & 2: the side chunk (128 pixels x 64 channels, the box of the output store) is fetched by TMA kSB-1 chunks
    // ahead into a kSB-deep staging ring; epi_tid 0 walks the same (tile, chunk) sequence as the consumers below
    int s_pt = pair, s_q = 0;
    uint32_t s_count = 0;
    auto side_issue = [&]() {
      while (s_pt < total_pair_tiles) {
        const int nt = s_pt % n_tiles;
        const int mtile = 2 * ((s_pt / n_tiles) * kPairs + (int)pidx) + (int)rank;
        const int n0 = nt * p.tile_n;
        int ncols = p.cout - n0;
        ncols = ncols > p.tile_n ? p.tile_n : ncols;
        if (mtile >= m_tiles || s_q >= ((ncols + 63) >> 6)) { s_pt += npairs; s_q = 0; continue; }
        int mt = mtile;
        const int tx = mt % p.tiles_x; mt /= p.tiles_x;
        const int ty = mt % p.tiles_y;
        const int img = mt / p.tiles_y;
        const uint32_t b = s_count % kSB;
        mbar_expect_tx(sfull0 + 8 * b, kEpiBufBytes);
        tma_load_4d(smem_u32(side_buf) + b * kEpiBufBytes, &tmap_s, sfull0 + 8 * b, n0 + s_q * 64, tx * p.bw, ty * p.bh, img);
        ++s_q;
        ++s_count;
        return;
      }
    };
    if ((MODE & 2) && epi_tid == 0)
      for (int i = 0; i < kSB - 1; ++i) side_issue();
    uint32_t tcount = 0, chunk_count = 0;
    // MODE & 1: a warp's column sums of its 16 rows stay in registers (lane = one channel pair of each of the <= 4
    // chunks of an output-channel tile) and go to the shared staging only when the output-channel tile changes - once
    // per launch when C_out fits one tile.  (One shared fp32 atomic per value and chunk, eight warps on the same 64
    // addresses, is a CAS loop per add: it cost ~1 800 cycles per chunk and doubled the time of short-K launches.)
    float2 rs0 = make_float2(0.f, 0.f), rs1 = rs0, rs2 = rs0, rs3 = rs0, rq0 = rs0, rq1 = rs0, rq2 = rs0, rq3 = rs0;
    int stats_nt = -1;
    auto stats_flush = [&]() {
      if (stats_nt < 0) return;
      float* st = stats_sm + stats_nt * 512;
      const int n0f = stats_nt * p.tile_n;
      const float2 vs[4] = {rs0, rs1, rs2, rs3}, vq[4] = {rq0, rq1, rq2, rq3};
      // the eight warps add their registers in turn (fixed order, plain adds): a CTA's partial statistics are
      // bit-reproducible, and the fp64 global sum of those partials is exact.  The flush runs once per change of the
      // output-channel tile, CTA-uniformly, so the extra barriers are noise next to a tile's epilogue.
      for (int turn = 0; turn < 8; ++turn) {
        if (warp - 2 == turn) {
#pragma unroll
          for (int qq = 0; qq < 4; ++qq) {
            const int c = qq * 64 + 2 * lane;
            if (c < p.tile_n && n0f + c < p.cout) {
              st[c] += vs[qq].x; st[c + 1] += vs[qq].y;
              st[256 + c] += vq[qq].x; st[256 + c + 1] += vq[qq].y;
            }
          }
        }
        epi_bar_sync256();
      }
      rs0 = rs1 = rs2 = rs3 = rq0 = rq1 = rq2 = rq3 = make_float2(0.f, 0.f);
    };
    for (int pt = pair; pt < total_pair_tiles; pt += npairs, ++tcount) {
      const int nt = pt % n_tiles;
      const int mtile = 2 * ((pt / n_tiles) * kPairs + (int)pidx) + (int)rank;
      const bool tile_ok = mtile < m_tiles;
      if ((MODE & 1) && nt != stats_nt) { stats_flush(); stats_nt = nt; }
      int mt = tile_ok ? mtile : 0;
      const int tx = mt % p.tiles_x; mt /= p.tiles_x;
      const int ty = mt % p.tiles_y;
      const int img = mt / p.tiles_y;
      const int n0 = nt * p.tile_n;
      const int oy = ty * p.bh + row / p.bw, ox = tx * p.bw + row % p.bw;
      const bool row_ok = tile_ok && oy < p.ho && ox < p.wo;
      int ncols = p.cout - n0;
      ncols = ncols > p.tile_n ? p.tile_n : ncols;
      const int nchunks = (ncols + 63) >> 6;
      const int acc = tcount & 1;
      mbar_wait(tfull0 + 8 * acc, (tcount >> 1) & 1);
      tc_fence_after();
      if (tr && epi_tid == 0 && tcount < 16) TC_MARK(24 + tcount);
      const uint32_t t_addr = tmem_base + acc * kPBN + ((uint32_t)(lg * 32) << 16);
#pragma unroll 1
      for (int q = 0; q < nchunks; ++q) {
        const int col0 = n0 + q * 64;
        uint4 sd0[4];
        if ((MODE & 2) && tile_ok) {
          // this row's 128 bytes of the staged side chunk (zero where the box left the tensor)
          const uint32_t sb = chunk_count % kSB;
          mbar_wait(sfull0 + 8 * sb, (chunk_count / kSB) & 1);
          const uint8_t* rowp = side_buf + sb * kEpiBufBytes + row * 128;
          const int sw = row & 7;
#pragma unroll
          for (int g = 0; g < 4; ++g) sd0[g] = *reinterpret_cast<const uint4*>(rowp + (((half * 4 + g) ^ sw) << 4));
        }
        uint32_t r0[32];
        tmem_ld32(t_addr + (uint32_t)(q * 64 + half * 32), r0);
        float bias32[32], sscale32[32];
        epi_load_bias32(ep.bias, col0 + half * 32, p.cout, bias32);
        if (MODE & 2) {
          epi_load_bias32(ep.side_scale, col0 + half * 32, p.cout, sscale32);
          if (ep.side_scale == nullptr) {
#pragma unroll
            for (int j = 0; j < 32; ++j) sscale32[j] = 1.f;
          }
        }
        tmem_ld_wait();
        if (q == nchunks - 1) {  // the accumulator is in registers: hand it back to the MMA issuer
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_cluster(lead_tempty0 + 8 * acc);
        }
        if (!tile_ok) continue;  // CTA-uniform: the odd tile of the last pair
        uint32_t pk0[16];
        epi_convert32<MODE>(ep, r0, col0 + half * 32, p.cout, row_ok, sd0, bias32, sscale32, pk0);
        const uint32_t buf = (chunk_count & 1) * kEpiBufBytes;
        if (epi_tid == 0) bulk_wait_read<1>();  // the store that last read this buffer (two chunks ago) is done with it
        epi_bar_sync256();
        if ((MODE & 2) && epi_tid == 0) side_issue();   // every thread has read its side row: refill that buffer
        {
          uint8_t* rowp = epi_buf + buf + row * 128;
          const int sw = row & 7;
#pragma unroll
          for (int j = 0; j < 4; ++j)
            *reinterpret_cast<uint4*>(rowp + (((half * 4 + j) ^ sw) << 4)) =
                make_uint4(pk0[4 * j], pk0[4 * j + 1], pk0[4 * j + 2], pk0[4 * j + 3]);
        }
        fence_proxy_async();
        epi_bar_sync256();
        if (epi_tid == 0) {
          tma_store_4d(&tmap_y, smem_u32(epi_buf + buf), col0, tx * p.bw, ty * p.bh, img);
          bulk_commit();
        }
        if (MODE & 1) {
          // column sums of the staged (bf16-rounded) chunk: each of the 8 warps sums 16 rows, lane = one pair of
          // channels; a warp reads one whole 128-byte row per step, so the swizzled layout is conflict-free here too
          const uint8_t* bufp = epi_buf + buf;
          float2 sm = make_float2(0.f, 0.f), sq = make_float2(0.f, 0.f);
#pragma unroll 8
          for (int rr = 0; rr < 16; ++rr) {
            const int r = lg * 32 + half * 16 + rr;
            const uint32_t u = *reinterpret_cast<const uint32_t*>(bufp + r * 128 + ((((lane >> 2) ^ (r & 7)) << 4) | ((lane & 3) << 2)));
            const float2 v = make_float2(__uint_as_float(u << 16), __uint_as_float(u & 0xffff0000u));
            sm = __fadd2_rn(sm, v);
            sq = __ffma2_rn(v, v, sq);
          }
          if (q == 0) { rs0 = __fadd2_rn(rs0, sm); rq0 = __fadd2_rn(rq0, sq); }
          else if (q == 1) { rs1 = __fadd2_rn(rs1, sm); rq1 = __fadd2_rn(rq1, sq); }
          else if (q == 2) { rs2 = __fadd2_rn(rs2, sm); rq2 = __fadd2_rn(rq2, sq); }
          else { rs3 = __fadd2_rn(rs3, sm); rq3 = __fadd2_rn(rq3, sq); }
        }
        ++chunk_count;
      }
      if (tr && epi_tid == 0 && tcount < 16) TC_MARK(40 + tcount);
    }
    if (MODE & 1) stats_flush();
    if (epi_tid == 0) bulk_wait_read<0>();
    if (MODE & 1) tc_epilogue_flush_stats<256>(ep, stats_sm, n_tiles, p.tile_n, p.cout, epi_tid);
    if (epi_tid == 0) TC_MARK(56);
  }
  tc_fence_before();
  cluster_sync_all();
  if (warp == 1) tmem_dealloc_2sm(tmem_base, 2 * kPBN);
  if (tr && threadIdx.x == 0) { tr[57] = (unsigned long long)clock64(); tr[58] = gtimer(); }
}

// ------------------------------------------------------------------------------ wgrad
constexpr int kWStages = 4;
constexpr int kWBox = 64 * 128;  // 64 pixels x 64 bf16 channels

struct TcWgradParams {
  int n, ho, wo, cout, cin;
  int kh, kw, pad, dil;
  int bw, bh;                // pixel tile (bw*bh == 64)
  int tiles_x, tiles_y;
  int splits;
};

template <int BN>
__global__ void __launch_bounds__(128) conv_tc_wgrad_kernel(const __grid_constant__ CUtensorMap tmap_dy,
                                                            const __grid_constant__ CUtensorMap tmap_x,
                                                            float* __restrict__ dw, TcWgradParams p) {
  constexpr int kABytesW = 2 * kWBox;          // 128 co
  constexpr int kBBytesW = (BN / 64) * kWBox;  // BN ci
  constexpr int kStageBytes = kABytesW + kBBytesW;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kWStages * kStageBytes);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kWStages + 1);
  const uint32_t smem_base = smem_u32(smem);
  const uint32_t full0 = smem_u32(bars), empty0 = full0 + 8 * kWStages, tfull = full0 + 16 * kWStages;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int co0 = blockIdx.x * 128, ci0 = blockIdx.y * BN;
  const int taps = p.kh * p.kw;
  const int split = blockIdx.z / taps, tap = blockIdx.z - split * taps;  // split slowest: taps of a split share X/dY in L2
  const int khi = tap / p.kw, kwi = tap - khi * p.kw;
  const int ptiles = p.n * p.tiles_y * p.tiles_x;
  const int per = (ptiles + p.splits - 1) / p.splits;
  const int t_begin = split * per;
  const int t_end = min(t_begin + per, ptiles);
  const int num_kb = max(t_end - t_begin, 0);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_dy);
    tma_prefetch_desc(&tmap_x);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < kWStages; ++s) {
      mbar_init(full0 + 8 * s, 1);
      mbar_init(empty0 + 8 * s, 1);
    }
    mbar_init(tfull, 1);
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc(smem_u32(tmem_slot), BN);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (num_kb > 0) {
    if (warp == 0) {
      if (lane == 0) {
        for (int kb = 0; kb < num_kb; ++kb) {
          const int s = kb % kWStages;
          const uint32_t ph = (kb / kWStages) & 1;
          mbar_wait(empty0 + 8 * s, ph ^ 1);
          int t = t_begin + kb;
          const int tx = t % p.tiles_x; t /= p.tiles_x;
          const int ty = t % p.tiles_y;
          const int img = t / p.tiles_y;
          const int ox0 = tx * p.bw, oy0 = ty * p.bh;
          const uint32_t sa = smem_base + s * kStageBytes;
          mbar_expect_tx(full0 + 8 * s, kStageBytes);
          tma_load_4d(sa, &tmap_dy, full0 + 8 * s, co0, ox0, oy0, img);
          tma_load_4d(sa + kWBox, &tmap_dy, full0 + 8 * s, co0 + 64, ox0, oy0, img);
#pragma unroll
          for (int j = 0; j < BN / 64; ++j)
            tma_load_4d(sa + kABytesW + j * kWBox, &tmap_x, full0 + 8 * s, ci0 + j * 64,
                        ox0 - p.pad + kwi * p.dil, oy0 - p.pad + khi * p.dil, img);
        }
      }
    } else if (warp == 1) {
      if (lane == 0) {
        constexpr uint32_t idesc = make_idesc(128, BN, 1, 1);  // both operands MN-major
        for (int kb = 0; kb < num_kb; ++kb) {
          const int s = kb % kWStages;
          const uint32_t ph = (kb / kWStages) & 1;
          mbar_wait(full0 + 8 * s, ph);
          tc_fence_after();
          const uint32_t sa = smem_base + s * kStageBytes;
          const uint32_t sb = sa + kABytesW;
#pragma unroll
          for (int k = 0; k < 4; ++k) {  // 16 pixels per MMA = 16 smem rows of 128 B
            const uint64_t ad = make_smem_desc(sa + k * 2048, kWBox, 1024);
            const uint64_t bd = make_smem_desc(sb + k * 2048, kWBox, 1024);
            umma_bf16(tmem_base, ad, bd, idesc, (kb > 0 || k > 0) ? 1u : 0u);
          }
          umma_commit(empty0 + 8 * s);
        }
        umma_commit(tfull);
      }
    }
    __syncwarp();

    mbar_wait(tfull, 0);
    tc_fence_after();
    const int co = co0 + warp * 32 + lane;
    float* drow = dw + ((size_t)tap * p.cout + co) * p.cin + ci0;
#pragma unroll 1
    for (int c0 = 0; c0 < BN; c0 += 32) {
      if (ci0 + c0 >= p.cin) break;
      uint32_t r[32];
      tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, r);
      tmem_ld_wait();
      if (co < p.cout) {
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (ci0 + c0 + j < p.cin) atomicAdd(drow + c0 + j, __uint_as_float(r[j]));
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, BN);
}


// ------------------------------------------------------------------------------ wgrad, persistent (BN = 256)
// Work item = (128 output channels) x (256 input channels) x tap x pixel split.  Items are ordered
// with the pixel split slowest, so the CTAs that run at the same time read the same slice of dY / X
// and share it through L2; a CTA loops over items with a double-buffered TMEM accumulator so the
// red.global.add epilogue of one item overlaps the main loop of the next.  N = 256 halves the
// shared-memory operand traffic per MMA relative to the 128 x 128 kernel.
constexpr int kWPStages = 4;
constexpr int kWPABytes = 2 * kWBox;              // 128 co
constexpr int kWPBBytes = 4 * kWBox;              // 256 ci
constexpr int kWPStageBytes = kWPABytes + kWPBBytes;  // 48 KB per 64 pixels

struct TcWgradPParams {
  int n, ho, wo, cout, cin;
  int kh, kw, pad, dil;
  int bw, bh, tiles_x, tiles_y;
  int co_tiles, ci_tiles, splits, total_items;
  int ci_tile;               // input-channel tile (multiple of 16, <= 256), balanced over ci_tiles
};

__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

__global__ void __launch_bounds__(192, 1) conv_tc_wgrad_persistent_kernel(const __grid_constant__ CUtensorMap tmap_dy,
                                                                          const __grid_constant__ CUtensorMap tmap_x,
                                                                          float* __restrict__ dw, TcWgradPParams p) {
  pdl_trigger();
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kWPStages * kWPStageBytes);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kWPStages + 4);
  const uint32_t smem_base = smem_u32(smem);
  const uint32_t full0 = smem_u32(bars), empty0 = full0 + 8 * kWPStages;
  const uint32_t tfull0 = empty0 + 8 * kWPStages, tempty0 = tfull0 + 16;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int taps = p.kh * p.kw;
  const int ptiles = p.n * p.tiles_y * p.tiles_x;
  const int per = (ptiles + p.splits - 1) / p.splits;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_dy);
    tma_prefetch_desc(&tmap_x);
    for (int s = 0; s < kWPStages; ++s) {
      mbar_init(full0 + 8 * s, 1);
      mbar_init(empty0 + 8 * s, 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull0 + 8 * a, 1);
      mbar_init(tempty0 + 8 * a, 4);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(smem_u32(tmem_slot), 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();

  // item -> (co tile fastest, ci tile, tap, split slowest)
  auto decode = [&](int item, int& cot, int& cit, int& tap, int& t_begin, int& t_end) {
    cot = item % p.co_tiles; item /= p.co_tiles;
    cit = item % p.ci_tiles; item /= p.ci_tiles;
    tap = item % taps;
    const int split = item / taps;
    t_begin = split * per;
    t_end = min(t_begin + per, ptiles);
  };

  // warp-uniform producer / issuer loops, elect_one() around the TMA / tcgen05 instructions (see the pair kernels)
  if (warp == 0) {
    uint32_t s = 0, ph = 0;
    for (int item = blockIdx.x; item < p.total_items; item += gridDim.x) {
      int cot, cit, tap, t_begin, t_end;
      decode(item, cot, cit, tap, t_begin, t_end);
      const int khi = tap / p.kw, kwi = tap - khi * p.kw;
      const int co0 = cot * 128, ci0 = cit * p.ci_tile;
      int ci_n = p.cin - ci0;
      ci_n = ci_n > p.ci_tile ? p.ci_tile : ci_n;
      const int ci_boxes = (ci_n + 63) / 64;
      int tt = t_begin;
      int tx = tt % p.tiles_x; tt /= p.tiles_x;
      int ty = tt % p.tiles_y;
      int img = tt / p.tiles_y;
      for (int t = t_begin; t < t_end; ++t) {
        mbar_wait(empty0 + 8 * s, ph ^ 1);
        const int ox0 = tx * p.bw, oy0 = ty * p.bh;
        const uint32_t sa = smem_base + s * kWPStageBytes;
        if (elect_one()) {
          mbar_expect_tx(full0 + 8 * s, kWPABytes + ci_boxes * kWBox);
          tma_load_4d(sa, &tmap_dy, full0 + 8 * s, co0, ox0, oy0, img);
          tma_load_4d(sa + kWBox, &tmap_dy, full0 + 8 * s, co0 + 64, ox0, oy0, img);
          for (int j = 0; j < ci_boxes; ++j)
            tma_load_4d(sa + kWPABytes + j * kWBox, &tmap_x, full0 + 8 * s, ci0 + j * 64, ox0 - p.pad + kwi * p.dil,
                        oy0 - p.pad + khi * p.dil, img);
        }
        __syncwarp();
        if (++tx == p.tiles_x) { tx = 0; if (++ty == p.tiles_y) { ty = 0; ++img; } }
        if (++s == kWPStages) { s = 0; ph ^= 1; }
      }
    }
  } else if (warp == 1) {
    uint32_t s = 0, ph = 0, icount = 0;
    const uint64_t desc_hi = ((uint64_t)(kWBox >> 4) << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) |
                             ((uint64_t)2 << 61);
    for (int item = blockIdx.x; item < p.total_items; item += gridDim.x, ++icount) {
      int cot, cit, tap, t_begin, t_end;
      decode(item, cot, cit, tap, t_begin, t_end);
      const int ci0 = cit * p.ci_tile;
      int ci_n = p.cin - ci0;
      ci_n = ci_n > p.ci_tile ? p.ci_tile : ((ci_n + 15) & ~15);
      const uint32_t idesc = make_idesc(128, ci_n, 1, 1);  // both operands MN-major
      const int acc = icount & 1;
      mbar_wait(tempty0 + 8 * acc, ((icount >> 1) & 1) ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * 256;
      uint32_t accum = 0;
      for (int t = t_begin; t < t_end; ++t) {
        mbar_wait(full0 + 8 * s, ph);
        tc_fence_after();
        const uint32_t sa = smem_base + s * kWPStageBytes;
        const uint64_t ad = desc_hi | (uint64_t)((sa >> 4) & 0x3FFF);
        const uint64_t bd = desc_hi | (uint64_t)(((sa + kWPABytes) >> 4) & 0x3FFF);
        if (elect_one()) {
          umma_bf16(d_tmem, ad, bd, idesc, accum);
          umma_bf16(d_tmem, ad + (uint64_t)(2048 >> 4), bd + (uint64_t)(2048 >> 4), idesc, 1u);
          umma_bf16(d_tmem, ad + (uint64_t)(2 * (2048 >> 4)), bd + (uint64_t)(2 * (2048 >> 4)), idesc, 1u);
          umma_bf16(d_tmem, ad + (uint64_t)(3 * (2048 >> 4)), bd + (uint64_t)(3 * (2048 >> 4)), idesc, 1u);
          umma_commit(empty0 + 8 * s);
        }
        __syncwarp();
        accum = 1;
        if (++s == kWPStages) { s = 0; ph ^= 1; }
      }
      if (elect_one()) umma_commit(tfull0 + 8 * acc);
      __syncwarp();
    }
  } else {
    const int lg = warp & 3;
    uint32_t icount = 0;
    for (int item = blockIdx.x; item < p.total_items; item += gridDim.x, ++icount) {
      int cot, cit, tap, t_begin, t_end;
      decode(item, cot, cit, tap, t_begin, t_end);
      const int ci0 = cit * p.ci_tile;
      const int co = cot * 128 + lg * 32 + lane;
      const int acc = icount & 1;
      mbar_wait(tfull0 + 8 * acc, (icount >> 1) & 1);
      tc_fence_after();
      if (t_end > t_begin) {
        float* drow = dw + ((size_t)tap * p.cout + co) * p.cin + ci0;
        const uint32_t t_addr = tmem_base + acc * 256 + ((uint32_t)(lg * 32) << 16);
#pragma unroll 1
        for (int c0 = 0; c0 < p.ci_tile; c0 += 32) {
          if (ci0 + c0 >= p.cin) break;
          uint32_t r[32];
          tmem_ld32(t_addr + (uint32_t)c0, r);
          tmem_ld_wait();
          if (co < p.cout) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              if (ci0 + c0 + j < p.cin && c0 + j < p.ci_tile)  // cin % 8 == 0: whole groups of 4 are in or out
                red_add_v4(drow + c0 + j, __uint_as_float(r[j]), __uint_as_float(r[j + 1]), __uint_as_float(r[j + 2]),
                           __uint_as_float(r[j + 3]));
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty0 + 8 * acc);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

// ------------------------------------------------------------------------------ wgrad, CTA pair (cta_group::2)
// One pair = one (256 output channels) x (<= 256 input channels) accumulator per tap and pixel split: each CTA
// stages 64 pixels of ITS 128 output channels of dY and ITS half of the input channels of X (32 KB per 64 pixels
// instead of the single-CTA kernel's 48 KB for the same MMA work), which is what bounds this kernel: the L2 -> SM
// operand stream, not the tensor pipe.
constexpr int kW2Stages = 6;
constexpr int kW2ABytes = 2 * kWBox;                 // 128 co x 64 px
constexpr int kW2BBytes = 2 * kWBox;                 // up to 128 ci x 64 px
constexpr int kW2StageBytes = kW2ABytes + kW2BBytes;  // 32 KB

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(192, 1)
conv_tc_wgrad_2cta_kernel(const __grid_constant__ CUtensorMap tmap_dy, const __grid_constant__ CUtensorMap tmap_x,
                          float* __restrict__ dw, TcWgradPParams p) {
  pdl_trigger();
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kW2Stages * kW2StageBytes);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kW2Stages + 4);
  const uint32_t smem_base = smem_u32(smem);
  const uint32_t full0 = smem_u32(bars), empty0 = full0 + 8 * kW2Stages;
  const uint32_t tfull0 = empty0 + 8 * kW2Stages, tempty0 = tfull0 + 16;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;
  const int taps = p.kh * p.kw;
  const int ptiles = p.n * p.tiles_y * p.tiles_x;
  const int per = (ptiles + p.splits - 1) / p.splits;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_dy);
    tma_prefetch_desc(&tmap_x);
    for (int s = 0; s < kW2Stages; ++s) {
      mbar_init(full0 + 8 * s, 2);
      mbar_init(empty0 + 8 * s, 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull0 + 8 * a, 1);
      mbar_init(tempty0 + 8 * a, 8);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc_2sm(smem_u32(tmem_slot), 512);
    tmem_relinquish_2sm();
  }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();

  // item -> (co tile fastest, ci tile, tap, split slowest)
  auto decode = [&](int item, int& cot, int& cit, int& tap, int& t_begin, int& t_end) {
    cot = item % p.co_tiles; item /= p.co_tiles;
    cit = item % p.ci_tiles; item /= p.ci_tiles;
    tap = item % taps;
    const int split = item / taps;
    t_begin = split * per;
    t_end = min(t_begin + per, ptiles);
  };
  // input channels of one ci tile, rounded so that each CTA's half is a multiple of 16
  auto ci_count = [&](int ci0) {
    int ci_n = p.cin - ci0;
    return ci_n > p.ci_tile ? p.ci_tile : ((ci_n + 31) & ~31);
  };

  // producer and issuer loops are warp-uniform; only the TMA / tcgen05 instructions sit under elect_one() (see the
  // forward kernel: a divergent single-lane loop costs ~125 instructions per k-block on the issuing thread)
  if (warp == 0) {
    uint32_t s = 0, ph = 0;
    const uint32_t lead_full0 = map_to_cta(full0, 0);
    for (int item = pair; item < p.total_items; item += npairs) {
      int cot, cit, tap, t_begin, t_end;
      decode(item, cot, cit, tap, t_begin, t_end);
      const int khi = tap / p.kw, kwi = tap - khi * p.kw;
      const int ci0 = cit * p.ci_tile;
      const int half = ci_count(ci0) >> 1;
      const int co_r = cot * 256 + (int)rank * 128, ci_r = ci0 + (int)rank * half;
      const int nb = (half + 63) >> 6;
      const uint32_t tx_bytes = 2 * (kW2ABytes + nb * kWBox);
      int tt = t_begin;
      int tx = tt % p.tiles_x; tt /= p.tiles_x;
      int ty = tt % p.tiles_y;
      int img = tt / p.tiles_y;
      const int dx = kwi * p.dil - p.pad, dy = khi * p.dil - p.pad;
      for (int t = t_begin; t < t_end; ++t) {
        mbar_wait(empty0 + 8 * s, ph ^ 1);
        const int ox0 = tx * p.bw, oy0 = ty * p.bh;
        const uint32_t sa = smem_base + s * kW2StageBytes;
        const uint32_t lead_full = (full0 + 8 * s) & 0xFEFFFFFFu;
        if (elect_one()) {
          if (leader) mbar_expect_tx(full0 + 8 * s, tx_bytes);
          else mbar_arrive_cluster(lead_full0 + 8 * s);
          tma_load_4d_2sm(sa, &tmap_dy, lead_full, co_r, ox0, oy0, img);
          tma_load_4d_2sm(sa + kWBox, &tmap_dy, lead_full, co_r + 64, ox0, oy0, img);
          tma_load_4d_2sm(sa + kW2ABytes, &tmap_x, lead_full, ci_r, ox0 + dx, oy0 + dy, img);
          if (nb > 1) tma_load_4d_2sm(sa + kW2ABytes + kWBox, &tmap_x, lead_full, ci_r + 64, ox0 + dx, oy0 + dy, img);
        }
        __syncwarp();
        if (++tx == p.tiles_x) { tx = 0; if (++ty == p.tiles_y) { ty = 0; ++img; } }
        if (++s == kW2Stages) { s = 0; ph ^= 1; }
      }
    }
  } else if (warp == 1) {
    if (leader) {
      uint32_t s = 0, ph = 0, icount = 0;
      // MN-major operands: LBO = distance between 64-channel atoms, SBO = distance between 8-pixel groups
      const uint64_t desc_hi = ((uint64_t)(kWBox >> 4) << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) |
                               ((uint64_t)2 << 61);
      for (int item = pair; item < p.total_items; item += npairs, ++icount) {
        int cot, cit, tap, t_begin, t_end;
        decode(item, cot, cit, tap, t_begin, t_end);
        const uint32_t idesc = make_idesc(256, ci_count(cit * p.ci_tile), 1, 1);
        const int acc = icount & 1;
        mbar_wait(tempty0 + 8 * acc, ((icount >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * 256;
        uint32_t accum = 0;
        for (int t = t_begin; t < t_end; ++t) {
          mbar_wait(full0 + 8 * s, ph);
          tc_fence_after();
          const uint32_t sa = smem_base + s * kW2StageBytes;
          const uint64_t ad = desc_hi | (uint64_t)((sa >> 4) & 0x3FFF);
          const uint64_t bd = desc_hi | (uint64_t)(((sa + kW2ABytes) >> 4) & 0x3FFF);
          if (elect_one()) {   // 16 pixels per MMA = 16 smem rows of 128 B
            umma_bf16_2sm(d_tmem, ad, bd, idesc, accum);
            umma_bf16_2sm(d_tmem, ad + (uint64_t)(2048 >> 4), bd + (uint64_t)(2048 >> 4), idesc, 1u);
            umma_bf16_2sm(d_tmem, ad + (uint64_t)(2 * (2048 >> 4)), bd + (uint64_t)(2 * (2048 >> 4)), idesc, 1u);
            umma_bf16_2sm(d_tmem, ad + (uint64_t)(3 * (2048 >> 4)), bd + (uint64_t)(3 * (2048 >> 4)), idesc, 1u);
            umma_commit_2sm(empty0 + 8 * s);
          }
          __syncwarp();
          accum = 1;
          if (++s == kW2Stages) { s = 0; ph ^= 1; }
        }
        if (elect_one()) umma_commit_2sm(tfull0 + 8 * acc);
        __syncwarp();
      }
    }
  } else {
    const int lg = warp & 3;
    const uint32_t lead_tempty0 = map_to_cta(tempty0, 0);
    uint32_t icount = 0;
    for (int item = pair; item < p.total_items; item += npairs, ++icount) {
      int cot, cit, tap, t_begin, t_end;
      decode(item, cot, cit, tap, t_begin, t_end);
      const int ci0 = cit * p.ci_tile;
      const int co = cot * 256 + (int)rank * 128 + lg * 32 + lane;
      const int acc = icount & 1;
      mbar_wait(tfull0 + 8 * acc, (icount >> 1) & 1);
      tc_fence_after();
      if (t_end > t_begin) {
        float* drow = dw + ((size_t)tap * p.cout + co) * p.cin + ci0;
        const uint32_t t_addr = tmem_base + acc * 256 + ((uint32_t)(lg * 32) << 16);
#pragma unroll 1
        for (int c0 = 0; c0 < p.ci_tile; c0 += 32) {
          if (ci0 + c0 >= p.cin) break;
          uint32_t r[32];
          tmem_ld32(t_addr + (uint32_t)c0, r);
          tmem_ld_wait();
          if (co < p.cout) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              if (ci0 + c0 + j < p.cin && c0 + j < p.ci_tile)  // cin % 8 == 0: whole groups of 4 are in or out
                red_add_v4(drow + c0 + j, __uint_as_float(r[j]), __uint_as_float(r[j + 1]), __uint_as_float(r[j + 2]),
                           __uint_as_float(r[j + 3]));
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(lead_tempty0 + 8 * acc);
    }
  }
  tc_fence_before();
  cluster_sync_all();
  if (warp == 1) tmem_dealloc_2sm(tmem_base, 512);
}

// pixel splits for a persistent wgrad launch: fill `workers` CTAs (or pairs) in whole rounds, keeping every item
// long enough (>= 8 pixel tiles) that its fixed cost (accumulator drain, pipeline refill) stays small
static int pick_wgrad_splits(int out_tiles, int ptiles, int workers) {
  int best = 1;
  double best_cost = 1e30;
  const int max_splits = ptiles / 8 > 1 ? ptiles / 8 : 1;
  for (int s = 1; s <= max_splits && s <= 4096; ++s) {
    const long items = (long)out_tiles * s;
    const long rounds = (items + workers - 1) / workers;
    const int per = (ptiles + s - 1) / s;
    const double cost = (double)rounds * (per + 6.0);  // 6 pixel tiles ~ drain + refill of one item
    if (cost < best_cost - 1e-9) { best_cost = cost; best = s; }
    if (rounds > 8) break;
  }
  return best;
}

// ------------------------------------------------------------------------------ host side
// packed weights [taps][rows][k] bf16 with a (64, bn, 1) box
static int make_weight_map(CUtensorMap* m, const void* base, int taps, int rows, int k, int bn) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) { set_error("cuTensorMapEncodeTiled entry point not available"); return CVX_ECUDA; }
  cuuint64_t dims[3] = {(cuuint64_t)k, (cuuint64_t)rows, (cuuint64_t)taps};
  cuuint64_t strides[2] = {(cuuint64_t)k * 2, (cuuint64_t)rows * k * 2};
  cuuint32_t box[3] = {64, (cuuint32_t)bn, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(weights %dx%dx%d box %d) failed: %d", taps, rows, k, bn, (int)r); return CVX_ECUDA; }
  return CVX_OK;
}

// choose the pixel tile (bw x bh == pixels) that wastes the fewest rows on an ho x wo map
static void pick_tile(int ho, int wo, int pixels, int* bw, int* bh) {
  int best_bw = pixels, best_cost = INT32_MAX;
  for (int w = 8; w <= pixels && w <= 256; w *= 2) {
    const int h = pixels / w;
    if (h > 256) continue;
    const int cost = ((wo + w - 1) / w) * ((ho + h - 1) / h);
    if (cost < best_cost || (cost == best_cost && w > best_bw)) { best_cost = cost; best_bw = w; }
  }
  *bw = best_bw;
  *bh = pixels / best_bw;
}

static int tc_supported(const cvx_conv_desc* d, const char* who, bool allow_stride2 = false) {
  CVX_CHECK_ARG(d != nullptr, "%s: null descriptor", who);
  const bool stride_ok = d->stride == 1 || (allow_stride2 && d->stride == 2);
  if (d->dtype != CVX_BF16 || !stride_ok || d->cin % 8 != 0 || d->cout % 8 != 0) {
    set_error("%s: tensor-core path needs bf16, stride 1, C_in %% 8 == 0 and C_out %% 8 == 0 (got dtype=%d stride=%d cin=%d cout=%d)",
              who, d->dtype, d->stride, d->cin, d->cout);
    return CVX_EUNSUPPORTED;
  }
  const int ho = (d->h + 2 * d->pad - d->dil * (d->kh - 1) - 1) / d->stride + 1;
  const int wo = (d->w + 2 * d->pad - d->dil * (d->kw - 1) - 1) / d->stride + 1;
  CVX_CHECK_ARG(ho == d->ho && wo == d->wo && ho > 0 && wo > 0, "%s: inconsistent output size", who);
  return CVX_OK;
}

template <int BN>
static int launch_fwd(const CUtensorMap& mx, const CUtensorMap& mw, const float* bias, void* y, const TcFwdParams& p,
                      cudaStream_t st) {
  constexpr int smem = kStages * (kABytes + BN * 128) + 1024 + 256;
  static bool configured = false;
  if (!configured) {
    CVX_CUDA_OK(cudaFuncSetAttribute(conv_tc_fwd_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured = true;
  }
  dim3 grid(p.n * p.tiles_y * p.tiles_x, (p.cout + BN - 1) / BN);
  conv_tc_fwd_kernel<BN><<<grid, 128, smem, st>>>(mx, mw, bias, (__nv_bfloat16*)y, p);
  CVX_LAUNCH_OK();
  return CVX_OK;
}

// pairs per cluster of the forward / data-gradient kernel (cvx_conv_tc_set_pairs; CERVIX_TC_PAIRS at load time)
static int g_tc_pairs = [] {
  const char* e = getenv("CERVIX_TC_PAIRS");
  const int v = e ? atoi(e) : 0;
  return (v == 1 || v == 2 || v == 4) ? v : 0;   // 0 = per-shape choice
}();
static int g_tc_pairs_force = 0;

// launch the CTA-pair forward kernel with `kp` pairs per cluster (cluster size 2*kp as a launch attribute)
template <int MODE, int kPairs>
static int launch_fwd_pairs_t(const CUtensorMap& mx, const CUtensorMap& mw, const CUtensorMap& my, const CUtensorMap& ms,
                              const TcEpi& ep, const TcFwdParams& p, int n_tiles, int m_tiles, cudaStream_t st) {
  constexpr int smem = TcFwdSmem<MODE>::kBytes;
  auto kern = conv_tc_fwd_2cta_kernel<MODE, kPairs>;
  static int max_clusters = 0;
  cudaLaunchConfig_t cfg = {};
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2 * kPairs;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;   // see pdl_trigger / pdl_wait in the kernel
  attr[1].val.programmaticStreamSerializationAllowed = g_pdl ? 1 : 0;
  cfg.blockDim = dim3(kPairThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cfg.attrs = attr;
  cfg.numAttrs = 2;
  if (max_clusters == 0) {
    CVX_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    cfg.gridDim = dim3(kNumSMs / (2 * kPairs) * (2 * kPairs));
    int n = 0;
    CVX_CUDA_OK(cudaOccupancyMaxActiveClusters(&n, kern, &cfg));
    if (n < 1) { set_error("conv_tc: no cluster of %d CTAs fits on this device", 2 * kPairs); return CVX_ECUDA; }
    max_clusters = n < kNumSMs / (2 * kPairs) ? n : kNumSMs / (2 * kPairs);
  }
  const int total = ((m_tiles + 2 * kPairs - 1) / (2 * kPairs)) * n_tiles;
  const int clusters = total < max_clusters ? total : max_clusters;
  cfg.gridDim = dim3(clusters * 2 * kPairs);
  CVX_CUDA_OK(cudaLaunchKernelEx(&cfg, kern, mx, mw, my, ms, ep, p, n_tiles, m_tiles, total));
  CVX_LAUNCH_OK();
  return CVX_OK;
}

template <int kPairs>
static int launch_fwd_pairs_m(int mode, const CUtensorMap& mx, const CUtensorMap& mw, const CUtensorMap& my,
                              const CUtensorMap& ms, const TcEpi& ep, const TcFwdParams& p, int n_tiles, int m_tiles,
                              cudaStream_t st) {
  switch (mode) {
    case 0: return launch_fwd_pairs_t<0, kPairs>(mx, mw, my, ms, ep, p, n_tiles, m_tiles, st);
    case 1: return launch_fwd_pairs_t<1, kPairs>(mx, mw, my, ms, ep, p, n_tiles, m_tiles, st);
    case 2: return launch_fwd_pairs_t<2, kPairs>(mx, mw, my, ms, ep, p, n_tiles, m_tiles, st);
    default: return launch_fwd_pairs_t<3, kPairs>(mx, mw, my, ms, ep, p, n_tiles, m_tiles, st);
  }
}

static int launch_fwd_pairs(int kp, int mode, const CUtensorMap& mx, const CUtensorMap& mw, const CUtensorMap& my,
                            const CUtensorMap& ms, const TcEpi& ep, const TcFwdParams& p, int n_tiles, int m_tiles,
                            cudaStream_t st) {
  if (kp == 4) return launch_fwd_pairs_m<4>(mode, mx, mw, my, ms, ep, p, n_tiles, m_tiles, st);
  if (kp == 2) return launch_fwd_pairs_m<2>(mode, mx, mw, my, ms, ep, p, n_tiles, m_tiles, st);
  return launch_fwd_pairs_m<1>(mode, mx, mw, my, ms, ep, p, n_tiles, m_tiles, st);
}

static int g_tc_trace_on = 0;
// tools/tc_trace.py: enable != 0 arms the phase marks of the next CTA-pair forward / data-gradient launches; out64
// (128 values, host memory, nullable) receives the marks of the launches made so far (after a device synchronise)
extern "C" CVX_API int cvx_debug_tc_trace(int enable, unsigned long long* out64) {
  g_tc_trace_on = enable;
  if (out64) {
    CVX_CUDA_OK(cudaDeviceSynchronize());
    CVX_CUDA_OK(cudaMemcpyFromSymbol(out64, g_tc_trace, sizeof(unsigned long long) * 128));
  }
  return CVX_OK;
}

// rows = output pixels [n,ho,wo] ; src = [n,hs,ws,cred] ; wp = [taps][ncol][cred]
static int run_igemm(int n, int hs, int ws, int cred, int ho, int wo, int ncol, int kh, int kw, int pad, int dil,
                     const void* src, const void* wp, const TcEpi& ep, void* dst, cudaStream_t st, int stride = 1) {
  const float* bias = ep.bias;
  TcFwdParams p;
  p.n = n; p.ho = ho; p.wo = wo; p.cout = ncol; p.cin = cred; p.kh = kh; p.kw = kw; p.pad = pad; p.dil = dil;
  p.stride = stride;
  p.trace = g_tc_trace_on;
  {
    const int nt = (ncol + kPBN - 1) / kPBN;
    p.tile_n = ((ncol + nt - 1) / nt + 15) & ~15;   // balanced tiles: 304 -> 2 x 160 instead of 256 + 48
    if (p.tile_n > kPBN) p.tile_n = kPBN;
  }
  pick_tile(ho, wo, 128, &p.bw, &p.bh);
  p.tiles_x = (wo + p.bw - 1) / p.bw;
  p.tiles_y = (ho + p.bh - 1) / p.bh;
  static const bool use_v1 = getenv("CERVIX_TC_V1") != nullptr;
  static const bool use_2cta = getenv("CERVIX_TC_1CTA") == nullptr;  // CTA-pair kernel by default
  if (stride != 1) {
    CVX_CHECK_ARG(stride == 2 && p.bw * 2 <= 256 && p.bh * 2 <= 256 && !use_v1, "conv_tc: unsupported stride %d", stride);
  }
  if (use_2cta && stride == 1) {
    // the epilogue stores 64-channel chunks by TMA: tiles start on multiples of 64 channels
    {
      const int nt = (ncol + kPBN - 1) / kPBN;
      p.tile_n = ((ncol + nt - 1) / nt + 63) & ~63;
    }
    const int n_tiles = (ncol + p.tile_n - 1) / p.tile_n;
    const int m_tiles = n * p.tiles_y * p.tiles_x;
    CVX_CHECK_ARG(!ep.stats || n_tiles <= kEpiMaxNTiles, "conv_tc: fused statistics need C_out <= %d", kEpiMaxNTiles * kPBN);
    const int mode = (ep.stats ? 1 : 0) | (ep.side ? 2 : 0);
    // pairs per cluster: share the weight tile among 2 pairs when there are enough pixel tiles to keep the machine full
    // Pairs per cluster.  Default (g_tc_pairs == 0): two pairs share the weight tile by TMA multicast only where it
    // measured faster - long reduction rows that are NOT 128-byte aligned (728 channels = 1456 B: every 128-byte box
    // row straddles two L2 lines, so the operand stream is what bounds those launches; -12 % on the 728 -> 728
    // middle-flow GEMMs, +5..10 % everywhere else, profiles/r01_conv_shapes_pairs_v6.txt).
    int kp = g_tc_pairs ? g_tc_pairs : (((cred * 2) % 128 != 0 && cred >= 512) ? 2 : 1);
    while (kp > 1 && ((p.tile_n / (2 * kp)) % 8 != 0 || (!g_tc_pairs_force && m_tiles < 2 * kp * 16))) kp >>= 1;
    CUtensorMap mx, mw, my;
    if (int rc = make_act_map(&mx, src, n, hs, ws, cred, p.bw, p.bh)) return rc;
    if (int rc = make_weight_map(&mw, wp, kh * kw, ncol, cred, p.tile_n / (2 * kp))) return rc;
    if (int rc = make_act_map(&my, dst, n, ho, wo, ncol, p.bw, p.bh)) return rc;
    CUtensorMap ms = my;
    if (ep.side)
      if (int rc = make_act_map(&ms, ep.side, n, ho, wo, ncol, p.bw, p.bh)) return rc;
    return launch_fwd_pairs(kp, mode, mx, mw, my, ms, ep, p, n_tiles, m_tiles, st);
  }
  if (!use_v1) {
    CUtensorMap mx, mw;
    if (int rc = make_act_map(&mx, src, n, hs, ws, cred, p.bw, p.bh, CU_TENSOR_MAP_SWIZZLE_128B, stride)) return rc;
    if (int rc = make_weight_map(&mw, wp, kh * kw, ncol, cred, p.tile_n)) return rc;
    constexpr int smem = kPStages * kPStageBytes + 1024 + 256 + kEpiStatsBytes;
    static bool configured = false;
    if (!configured) {
      CVX_CUDA_OK(cudaFuncSetAttribute(conv_tc_fwd_persistent_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
      CVX_CUDA_OK(cudaFuncSetAttribute(conv_tc_fwd_persistent_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
      configured = true;
    }
    const int n_tiles = (ncol + p.tile_n - 1) / p.tile_n;
    const int total = n * p.tiles_y * p.tiles_x * n_tiles;
    const int grid = total < kNumSMs ? total : kNumSMs;
    CVX_CHECK_ARG(!ep.stats || n_tiles <= kEpiMaxNTiles, "conv_tc: fused statistics need C_out <= %d", kEpiMaxNTiles * kPBN);
    if (ep.side || ep.stats)
      conv_tc_fwd_persistent_kernel<true><<<grid, 192, smem, st>>>(mx, mw, ep, (__nv_bfloat16*)dst, p, n_tiles, total);
    else
      conv_tc_fwd_persistent_kernel<false><<<grid, 192, smem, st>>>(mx, mw, ep, (__nv_bfloat16*)dst, p, n_tiles, total);
    CVX_LAUNCH_OK();
    return CVX_OK;
  }
  CVX_CHECK_ARG(!ep.side && !ep.stats, "conv_tc: the legacy kernel (CERVIX_TC_V1) has no fused epilogue");
  const int bn = ncol <= 64 ? 64 : 128;
  CUtensorMap mx, mw;
  if (int rc = make_act_map(&mx, src, n, hs, ws, cred, p.bw, p.bh)) return rc;
  if (int rc = make_weight_map(&mw, wp, kh * kw, ncol, cred, bn)) return rc;
  return bn == 64 ? launch_fwd<64>(mx, mw, bias, dst, p, st) : launch_fwd<128>(mx, mw, bias, dst, p, st);
}

template <int BN>
static int launch_wgrad(const CUtensorMap& mdy, const CUtensorMap& mx, float* dw, const TcWgradParams& p,
                        cudaStream_t st) {
  constexpr int smem = kWStages * (2 * kWBox + (BN / 64) * kWBox) + 1024 + 256;
  static bool configured = false;
  if (!configured) {
    CVX_CUDA_OK(cudaFuncSetAttribute(conv_tc_wgrad_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured = true;
  }
  dim3 grid((p.cout + 127) / 128, (p.cin + BN - 1) / BN, p.kh * p.kw * p.splits);
  conv_tc_wgrad_kernel<BN><<<grid, 128, smem, st>>>(mdy, mx, dw, p);
  CVX_LAUNCH_OK();
  return CVX_OK;
}

}  // namespace cvx

using namespace cvx;

extern "C" {

int cvx_conv_tc_set_pairs(int pairs, int force) {
  CVX_CHECK_ARG(pairs == 0 || pairs == 1 || pairs == 2 || pairs == 4,
                "conv_tc_set_pairs: pairs must be 0 (per-shape default), 1, 2 or 4 (got %d)", pairs);
  g_tc_pairs = pairs;
  g_tc_pairs_force = force != 0;
  return CVX_OK;
}

int cvx_conv_fwd_tc(const cvx_conv_desc* d, const void* x, const void* w_packed, const float* bias, void* y,
                    void* stream) {
  if (int rc = tc_supported(d, "conv_fwd_tc", true)) return rc;
  CVX_CHECK_ARG(x && w_packed && y, "conv_fwd_tc: null pointer");
  const TcEpi ep{bias, nullptr, nullptr, nullptr, 0};
  return run_igemm(d->n, d->h, d->w, d->cin, d->ho, d->wo, d->cout, d->kh, d->kw, d->pad, d->dil, x, w_packed, ep, y,
                   as_stream(stream), d->stride);
}

int cvx_conv_fwd_tc_ex(const cvx_conv_desc* d, const void* x, const void* w_packed, const float* bias, const void* side,
                       const float* side_scale, double* stats, void* y, void* stream) {
  if (int rc = tc_supported(d, "conv_fwd_tc_ex", true)) return rc;
  CVX_CHECK_ARG(x && w_packed && y && (!side || side_scale), "conv_fwd_tc_ex: null pointer");
  if (stats) CVX_WS_ZERO(stats, sizeof(double) * 2 * d->cout, as_stream(stream));
  const TcEpi ep{bias, (const __nv_bfloat16*)side, side_scale, stats, 0};
  return run_igemm(d->n, d->h, d->w, d->cin, d->ho, d->wo, d->cout, d->kh, d->kw, d->pad, d->dil, x, w_packed, ep, y,
                   as_stream(stream), d->stride);
}

int cvx_conv_fwd_tc_act(const cvx_conv_desc* d, const void* x, const void* w_packed, const float* bias, const void* side,
                        const float* side_scale, int act, void* y, void* stream) {
  if (int rc = tc_supported(d, "conv_fwd_tc_act", true)) return rc;
  CVX_CHECK_ARG(x && w_packed && y, "conv_fwd_tc_act: null pointer");   // side_scale == NULL: plain residual add
  CVX_CHECK_ARG(act == CVX_ACT_NONE || act == CVX_ACT_RELU, "conv_fwd_tc_act: activation %d not supported", act);
  CVX_CHECK_ARG(!side || d->stride == 1, "conv_fwd_tc_act: the side input needs a stride-1 convolution");
  const TcEpi ep{bias, (const __nv_bfloat16*)side, side_scale, nullptr, act == CVX_ACT_RELU ? 1 : 0};
  return run_igemm(d->n, d->h, d->w, d->cin, d->ho, d->wo, d->cout, d->kh, d->kw, d->pad, d->dil, x, w_packed, ep, y,
                   as_stream(stream), d->stride);
}

int cvx_conv_dgrad_tc_ex(const cvx_conv_desc* d, const void* dy, const void* w_packed_t, const float* bias,
                         const void* side, const float* side_scale, void* dx, void* stream) {
  if (int rc = tc_supported(d, "conv_dgrad_tc_ex")) return rc;
  CVX_CHECK_ARG(dy && w_packed_t && dx && (!side || side_scale), "conv_dgrad_tc_ex: null pointer");
  const int pad_t = d->dil * (d->kh - 1) - d->pad;
  CVX_CHECK_ARG(pad_t >= 0 && d->kh == d->kw, "conv_dgrad_tc_ex: unsupported padding/filter");
  const TcEpi ep{bias, (const __nv_bfloat16*)side, side_scale, nullptr, 0};
  return run_igemm(d->n, d->ho, d->wo, d->cout, d->h, d->w, d->cin, d->kh, d->kw, pad_t, d->dil, dy, w_packed_t, ep, dx,
                   as_stream(stream));
}

int cvx_conv_dgrad_tc(const cvx_conv_desc* d, const void* dy, const void* w_packed_t, void* dx, void* stream) {
  if (int rc = tc_supported(d, "conv_dgrad_tc")) return rc;
  CVX_CHECK_ARG(dy && w_packed_t && dx, "conv_dgrad_tc: null pointer");
  // stride-1 data gradient == forward conv of dy with the flipped/transposed filter and
  // padding dil*(k-1) - pad
  const int pad_t = d->dil * (d->kh - 1) - d->pad;
  CVX_CHECK_ARG(pad_t >= 0 && d->kh == d->kw, "conv_dgrad_tc: unsupported padding/filter");
  const TcEpi ep{nullptr, nullptr, nullptr, nullptr, 0};
  return run_igemm(d->n, d->ho, d->wo, d->cout, d->h, d->w, d->cin, d->kh, d->kw, pad_t, d->dil, dy, w_packed_t, ep, dx,
                   as_stream(stream));
}

int cvx_conv_wgrad_tc(const cvx_conv_desc* d, const void* x, const void* dy, float* dw_packed, void* stream) {
  if (int rc = tc_supported(d, "conv_wgrad_tc")) return rc;
  CVX_CHECK_ARG(x && dy && dw_packed, "conv_wgrad_tc: null pointer");
  static const bool wgrad_v1 = getenv("CERVIX_TC_WGRAD_V1") != nullptr;
  if (!wgrad_v1) {
    static const bool wgrad_1cta = getenv("CERVIX_TC_WGRAD_1CTA") != nullptr;
    const bool pairk = !wgrad_1cta && d->cout > 128;  // the pair tile is 256 output channels tall
    TcWgradPParams q;
    q.n = d->n; q.ho = d->ho; q.wo = d->wo; q.cout = d->cout; q.cin = d->cin;
    q.kh = d->kh; q.kw = d->kw; q.pad = d->pad; q.dil = d->dil;
    pick_tile(d->ho, d->wo, 64, &q.bw, &q.bh);
    q.tiles_x = (d->wo + q.bw - 1) / q.bw;
    q.tiles_y = (d->ho + q.bh - 1) / q.bh;
    q.co_tiles = pairk ? (d->cout + 255) / 256 : (d->cout + 127) / 128;
    q.ci_tiles = (d->cin + 255) / 256;
    const int ci_round = pairk ? 31 : 15;
    q.ci_tile = ((d->cin + q.ci_tiles - 1) / q.ci_tiles + ci_round) & ~ci_round;
    if (q.ci_tile > 256) q.ci_tile = 256;
    q.ci_tiles = (d->cin + q.ci_tile - 1) / q.ci_tile;
    const int ptiles = q.n * q.tiles_y * q.tiles_x;
    const int out_tiles = q.co_tiles * q.ci_tiles * d->kh * d->kw;
    const int workers = pairk ? kNumSMs / 2 : kNumSMs;
    q.splits = pick_wgrad_splits(out_tiles, ptiles, workers);
    q.total_items = out_tiles * q.splits;
    CUtensorMap mdy, mx;
    if (int rc = make_act_map(&mdy, dy, d->n, d->ho, d->wo, d->cout, q.bw, q.bh)) return rc;
    if (int rc = make_act_map(&mx, x, d->n, d->h, d->w, d->cin, q.bw, q.bh)) return rc;
    if (pairk) {
      constexpr int smem2 = kW2Stages * kW2StageBytes + 1024 + 256;
      static bool configured2 = false;
      if (!configured2) {
        CVX_CUDA_OK(cudaFuncSetAttribute(conv_tc_wgrad_2cta_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem2));
        configured2 = true;
      }
      const int pairs = q.total_items < workers ? q.total_items : workers;
      launch_pdl(conv_tc_wgrad_2cta_kernel, dim3(2 * pairs), dim3(192), smem2, as_stream(stream), mdy, mx, dw_packed, q);
      CVX_LAUNCH_OK();
      return CVX_OK;
    }
    constexpr int smem = kWPStages * kWPStageBytes + 1024 + 256;
    static bool configured = false;
    if (!configured) {
      CVX_CUDA_OK(cudaFuncSetAttribute(conv_tc_wgrad_persistent_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
      configured = true;
    }
    const int grid = q.total_items < kNumSMs ? q.total_items : kNumSMs;
    launch_pdl(conv_tc_wgrad_persistent_kernel, dim3(grid), dim3(192), smem, as_stream(stream), mdy, mx, dw_packed, q);
    CVX_LAUNCH_OK();
    return CVX_OK;
  }
  TcWgradParams p;
  p.n = d->n; p.ho = d->ho; p.wo = d->wo; p.cout = d->cout; p.cin = d->cin;
  p.kh = d->kh; p.kw = d->kw; p.pad = d->pad; p.dil = d->dil;
  pick_tile(d->ho, d->wo, 64, &p.bw, &p.bh);
  p.tiles_x = (d->wo + p.bw - 1) / p.bw;
  p.tiles_y = (d->ho + p.bh - 1) / p.bh;
  const int bn = d->cin <= 64 ? 64 : 128;
  const int ptiles = p.n * p.tiles_y * p.tiles_x;
  const int out_tiles = ((d->cout + 127) / 128) * ((d->cin + bn - 1) / bn) * d->kh * d->kw;
  int splits = (2 * kNumSMs + out_tiles - 1) / out_tiles;
  if (splits > ptiles / 4) splits = ptiles / 4;  // keep >= 4 pixel tiles per CTA
  if (splits < 1) splits = 1;
  CVX_CHECK_ARG((int64_t)d->kh * d->kw * splits <= 65535, "conv_wgrad_tc: grid.z too large");
  p.splits = splits;
  CUtensorMap mdy, mx;
  if (int rc = make_act_map(&mdy, dy, d->n, d->ho, d->wo, d->cout, p.bw, p.bh)) return rc;
  if (int rc = make_act_map(&mx, x, d->n, d->h, d->w, d->cin, p.bw, p.bh)) return rc;
  return bn == 64 ? launch_wgrad<64>(mdy, mx, dw_packed, p, as_stream(stream))
                  : launch_wgrad<128>(mdy, mx, dw_packed, p, as_stream(stream));
}

}  // extern "C"
