// Convolutions with a very narrow channel dimension, where an implicit GEMM tile would be
// almost empty and the op is bandwidth-bound:
//   * "narrow in"  : the stem conv 3 -> 32, 3x3 stride 2 (xception.py:95, mobilenetv2.py:94)
//                    forward + weight gradient (the image needs no data gradient);
//   * "narrow out" : the classifier 1x1 conv 256 -> num_classes (deeplabv3_plus.py:167)
//                    forward + data gradient + weight gradient.
#include "colreduce.cuh"
#include "conv_narrow.cuh"

namespace cvx {

struct NGeom {
  int n, h, w, cin, cout, kh, kw, stride, pad, dil, ho, wo;
};

constexpr int kStemCout = 32;
constexpr int kStemMaxK = 36;  // taps * cin

// ---------------------------------------------------------------- narrow in: forward
template <typename T>
__global__ void __launch_bounds__(256) stem_fwd_kernel(const T* __restrict__ x, const T* __restrict__ wp,
                                                       T* __restrict__ y, NGeom g) {
  __shared__ float ws[kStemMaxK][kStemCout];  // [tap*cin + ci][co]
  const int K = g.kh * g.kw * g.cin;
  for (int i = threadIdx.x; i < K * kStemCout; i += blockDim.x) {
    const int k = i / kStemCout, co = i % kStemCout;
    const int tap = k / g.cin, ci = k % g.cin;
    ws[k][co] = Elem<T>::ld(wp + ((size_t)tap * kStemCout + co) * g.cin + ci);
  }
  __syncthreads();
  const int64_t npix = (int64_t)g.n * g.ho * g.wo;
  for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < npix; p += (int64_t)gridDim.x * blockDim.x) {
    const int ox = (int)(p % g.wo), oy = (int)((p / g.wo) % g.ho), nn = (int)(p / ((int64_t)g.wo * g.ho));
    float acc[kStemCout];
#pragma unroll
    for (int c = 0; c < kStemCout; ++c) acc[c] = 0.f;
    for (int kh = 0; kh < g.kh; ++kh) {
      const int iy = oy * g.stride - g.pad + kh * g.dil;
      if (iy < 0 || iy >= g.h) continue;
      for (int kw = 0; kw < g.kw; ++kw) {
        const int ix = ox * g.stride - g.pad + kw * g.dil;
        if (ix < 0 || ix >= g.w) continue;
        const T* px = x + (((size_t)nn * g.h + iy) * g.w + ix) * g.cin;
        const int kbase = (kh * g.kw + kw) * g.cin;
        for (int ci = 0; ci < g.cin; ++ci) {
          const float xv = Elem<T>::ld(px + ci);
          const float4* wrow = reinterpret_cast<const float4*>(ws[kbase + ci]);
#pragma unroll
          for (int j = 0; j < kStemCout / 4; ++j) {
            const float4 w4 = wrow[j];
            acc[4 * j] = fmaf(xv, w4.x, acc[4 * j]);
            acc[4 * j + 1] = fmaf(xv, w4.y, acc[4 * j + 1]);
            acc[4 * j + 2] = fmaf(xv, w4.z, acc[4 * j + 2]);
            acc[4 * j + 3] = fmaf(xv, w4.w, acc[4 * j + 3]);
          }
        }
      }
    }
    T* out = y + p * kStemCout;
    constexpr int V = Elem<T>::kVec;
#pragma unroll
    for (int j = 0; j < kStemCout / V; ++j) {
      Vec<T> o;
#pragma unroll
      for (int i = 0; i < V; ++i) o.v[i] = acc[j * V + i];
      o.store(out + j * V);
    }
  }
}

// ---------------------------------------------------------------- narrow in: weight gradient
// block = 288 threads; thread (k, cg) owns dW[k][4*cg .. 4*cg+3]; pixels staged 64 at a time.  Staging keeps the
// index arithmetic out of the inner loops: 64 threads decompose the pixel index once per batch, the gather threads
// keep their (tap, channel) for the whole kernel, and the dY rows are one linear copy.
template <typename T>
__global__ void __launch_bounds__(288) stem_wgrad_kernel(const T* __restrict__ x, const T* __restrict__ dy,
                                                         float* __restrict__ dw, NGeom g) {
  constexpr int P = 64;
  __shared__ float xs[P][kStemMaxK + 1];
  __shared__ __align__(16) float dys[P][kStemCout];
  __shared__ int py0[P], px0[P], pimg[P];
  const int K = g.kh * g.kw * g.cin;
  const int k = threadIdx.x >> 3, cg = threadIdx.x & 7;
  const bool owner = k < K;
  // gather role: kk = lane (fixed), pixel lanes = warp index
  const int gk = threadIdx.x & 31, gl = threadIdx.x >> 5;   // 9 warps
  int g_dy[2] = {0, 0}, g_dx[2] = {0, 0}, g_ci[2] = {0, 0};   // kk = gk and gk + 32 (K <= 36)
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    const int kk = gk + 32 * j;
    if (kk < K) {
      const int tap = kk / g.cin;
      g_ci[j] = kk - tap * g.cin;
      g_dy[j] = (tap / g.kw) * g.dil;
      g_dx[j] = (tap % g.kw) * g.dil;
    }
  }
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  const int64_t npix = (int64_t)g.n * g.ho * g.wo;
  for (int64_t p0 = (int64_t)blockIdx.x * P; p0 < npix; p0 += (int64_t)gridDim.x * P) {
    if (threadIdx.x < P) {
      const int64_t p = p0 + threadIdx.x;
      int iy0 = -(1 << 28), ix0 = 0, img = 0;   // far outside the image: the gather stores zeros
      if (p < npix) {
        const int64_t t = p / g.wo;
        const int ox = (int)(p - t * g.wo), oy = (int)(t % g.ho);
        img = (int)(t / g.ho);
        iy0 = oy * g.stride - g.pad;
        ix0 = ox * g.stride - g.pad;
      }
      py0[threadIdx.x] = iy0; px0[threadIdx.x] = ix0; pimg[threadIdx.x] = img;
    }
    for (int i = threadIdx.x; i < P * kStemCout; i += blockDim.x) {
      const int pp = i / kStemCout;
      (&dys[0][0])[i] = (p0 + pp < npix) ? Elem<T>::ld(dy + p0 * kStemCout + i) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int kk = gk + 32 * j;
      if (kk < K) {
        for (int pp = gl; pp < P; pp += 9) {
          const int iy = py0[pp] + g_dy[j], ix = px0[pp] + g_dx[j];
          float v = 0.f;
          if (iy >= 0 && iy < g.h && ix >= 0 && ix < g.w)
            v = Elem<T>::ld(x + (((size_t)pimg[pp] * g.h + iy) * g.w + ix) * g.cin + g_ci[j]);
          xs[pp][kk] = v;
        }
      }
    }
    __syncthreads();
    if (owner) {
#pragma unroll 8
      for (int pp = 0; pp < P; ++pp) {
        const float a = xs[pp][k];
        const float4 d = *reinterpret_cast<const float4*>(&dys[pp][cg * 4]);
        acc.x = fmaf(a, d.x, acc.x); acc.y = fmaf(a, d.y, acc.y);
        acc.z = fmaf(a, d.z, acc.z); acc.w = fmaf(a, d.w, acc.w);
      }
    }
    __syncthreads();
  }
  if (owner) {
    const int tap = k / g.cin, ci = k % g.cin;
    const float v[4] = {acc.x, acc.y, acc.z, acc.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) atomicAdd(dw + ((size_t)tap * kStemCout + cg * 4 + j) * g.cin + ci, v[j]);
  }
}

// ---------------------------------------------------------------- narrow in: weight gradient on warp MMAs (bf16)
// dW[kk][co] = sum_p patch[p][kk] * dY[p][co] with kk = tap*cin + ci < 32 and co < 32 is a 32 x 32 x (pixels) product:
// 8 mma.m16n8k16 per 16 pixels and warp.  The A fragments (patch^T, row = kk, column = pixel) are gathered straight
// from x into registers - a thread needs 4 values of kk (gid + 8j) at 4 pixels (2 tig, +1, +8, +9), 16 two-byte loads
// that hit L1 (each x element is used by ~2 windows) - and the B fragments (dY, 16 pixels x 32 channels = 1 KB,
// fetched with two coalesced 16-byte loads per lane) go through an 80-byte-pitch shared tile and ldmatrix.trans.
// The kernel reads x and dY once at streaming rate; the SIMT version above is bound by its shared-memory loads
// (2 LDS per 4 FMAs) and ran at 0.35 TB/s.
__device__ __forceinline__ uint32_t pack_raw_bf16(unsigned short lo, unsigned short hi) {
  return (uint32_t)lo | ((uint32_t)hi << 16);
}

__global__ void __launch_bounds__(128) stem_wgrad_mma_kernel(const __nv_bfloat16* __restrict__ x,
                                                             const __nv_bfloat16* __restrict__ dy,
                                                             float* __restrict__ dw, NGeom g) {
  constexpr int kPitch = 80;                                  // bytes per staged dY row (64 + 16: conflict-free ldmatrix)
  __shared__ __align__(16) uint8_t dys[4][16 * kPitch];
  __shared__ float red[32 * kStemCout];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int gid = lane >> 2, tig = lane & 3;
  const int K = g.kh * g.kw * g.cin;
  for (int i = threadIdx.x; i < 32 * kStemCout; i += blockDim.x) red[i] = 0.f;
  int koff[4], kdy[4], kdx[4];
  bool kval[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int kk = gid + 8 * j;
    kval[j] = kk < K;
    const int tap = kval[j] ? kk / g.cin : 0;
    const int ci = kval[j] ? kk - tap * g.cin : 0;
    kdy[j] = (tap / g.kw) * g.dil;
    kdx[j] = (tap % g.kw) * g.dil;
    koff[j] = (kdy[j] * g.w + kdx[j]) * g.cin + ci;
  }
  float acc[2][4][4];
#pragma unroll
  for (int a = 0; a < 2; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b)
#pragma unroll
      for (int c = 0; c < 4; ++c) acc[a][b][c] = 0.f;
  const int64_t npix = (int64_t)g.n * g.ho * g.wo;
  const int64_t nblk = (npix + 15) >> 4;
  const bool row_blocks = (g.wo & 15) == 0;                   // a 16-pixel block never straddles an output row
  const unsigned short* xr = reinterpret_cast<const unsigned short*>(x);
  const uint32_t tile = (uint32_t)__cvta_generic_to_shared(&dys[warp][0]);
  const uint32_t ld_addr = tile + ((lane & 7) + 8 * ((lane >> 3) & 1)) * kPitch + (lane >> 4) * 16;
  __syncthreads();
  for (int64_t blk = (int64_t)blockIdx.x * 4 + warp; blk < nblk; blk += (int64_t)gridDim.x * 4) {
    const int64_t p0 = blk << 4;
    // ---- dY tile -> shared (rows past the end are zero)
    uint4 dv[2];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const int i = lane + 32 * r, row = i >> 2, c16 = i & 3;
      dv[r] = (p0 + row < npix) ? __ldg(reinterpret_cast<const uint4*>(dy + (p0 + row) * kStemCout + c16 * 8))
                                : make_uint4(0, 0, 0, 0);
    }
    // ---- patch gather
    int oxb = 0, oyb = 0, nb = 0;
    if (row_blocks) {
      const int64_t t = p0 / g.wo;
      oxb = (int)(p0 - t * g.wo); oyb = (int)(t % g.ho); nb = (int)(t / g.ho);
    }
    unsigned short xv[4][4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int pl = 2 * tig + (q & 1) + 8 * (q >> 1);
      int ox, oy, nn;
      if (row_blocks) {
        ox = oxb + pl; oy = oyb; nn = nb;
      } else {
        const int64_t pp = p0 + pl, t = pp / g.wo;
        ox = (int)(pp - t * g.wo); oy = (int)(t % g.ho); nn = (int)(t / g.ho);
      }
      const bool pok = p0 + pl < npix;
      const int iy0 = oy * g.stride - g.pad, ix0 = ox * g.stride - g.pad;
      const int64_t base = (((int64_t)nn * g.h + iy0) * g.w + ix0) * g.cin;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int iy = iy0 + kdy[j], ix = ix0 + kdx[j];
        const bool ok = pok && kval[j] && iy >= 0 && iy < g.h && ix >= 0 && ix < g.w;
        xv[q][j] = ok ? __ldg(xr + base + koff[j]) : (unsigned short)0;
      }
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const int i = lane + 32 * r, row = i >> 2, c16 = i & 3;
      *reinterpret_cast<uint4*>(&dys[warp][row * kPitch + c16 * 16]) = dv[r];
    }
    __syncwarp();
    uint32_t bf[4][2];                                         // [n tile][k half]
#pragma unroll
    for (int h2 = 0; h2 < 2; ++h2) {
      asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
                   : "=r"(bf[2 * h2][0]), "=r"(bf[2 * h2][1]), "=r"(bf[2 * h2 + 1][0]), "=r"(bf[2 * h2 + 1][1])
                   : "r"(ld_addr + h2 * 32));
    }
#pragma unroll
    for (int mt = 0; mt < 2; ++mt) {
      const uint32_t a0 = pack_raw_bf16(xv[0][2 * mt], xv[1][2 * mt]);
      const uint32_t a1 = pack_raw_bf16(xv[0][2 * mt + 1], xv[1][2 * mt + 1]);
      const uint32_t a2 = pack_raw_bf16(xv[2][2 * mt], xv[3][2 * mt]);
      const uint32_t a3 = pack_raw_bf16(xv[2][2 * mt + 1], xv[3][2 * mt + 1]);
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, "
                     "{%0, %1, %2, %3};"
                     : "+f"(acc[mt][nt][0]), "+f"(acc[mt][nt][1]), "+f"(acc[mt][nt][2]), "+f"(acc[mt][nt][3])
                     : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(bf[nt][0]), "r"(bf[nt][1]));
      }
    }
    __syncwarp();                                              // the tile is rewritten by the next iteration
  }
  // ---- block reduction in shared memory, then one global atomic per (kk, co) and block
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
      const int r0 = 16 * mt + gid, c0 = 8 * nt + 2 * tig;
      atomicAdd(&red[r0 * kStemCout + c0], acc[mt][nt][0]);
      atomicAdd(&red[r0 * kStemCout + c0 + 1], acc[mt][nt][1]);
      atomicAdd(&red[(r0 + 8) * kStemCout + c0], acc[mt][nt][2]);
      atomicAdd(&red[(r0 + 8) * kStemCout + c0 + 1], acc[mt][nt][3]);
    }
  __syncthreads();
  for (int i = threadIdx.x; i < K * kStemCout; i += blockDim.x) {
    const int kk = i / kStemCout, co = i - kk * kStemCout;
    const int tap = kk / g.cin, ci = kk - tap * g.cin;
    atomicAdd(dw + ((size_t)tap * kStemCout + co) * g.cin + ci, red[i]);
  }
}

// ---------------------------------------------------------------- narrow out (1x1, C_out <= 8)
constexpr int kMaxNarrowOut = 8;
constexpr int kMaxNarrowCin = 1024;

// one warp per pixel: lanes split the channel vectors, shuffle-reduce the C_out dot products
template <typename T>
__global__ void __launch_bounds__(256) narrow_out_fwd_kernel(const T* __restrict__ x, const T* __restrict__ wp,
                                                             const float* __restrict__ bias, T* __restrict__ y,
                                                             int64_t npix, int cin, int cout) {
  constexpr int V = Elem<T>::kVec;
  extern __shared__ float wsm[];  // [cout][cin]
  for (int i = threadIdx.x; i < cout * cin; i += blockDim.x) wsm[i] = Elem<T>::ld(wp + i);
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const int cvn = cin / V;
  for (int64_t p = warp0; p < npix; p += nwarps) {
    float acc[kMaxNarrowOut];
#pragma unroll
    for (int c = 0; c < kMaxNarrowOut; ++c) acc[c] = 0.f;
    for (int cv = lane; cv < cvn; cv += 32) {
      Vec<T> v;
      v.load(x + p * cin + cv * V);
#pragma unroll
      for (int c = 0; c < kMaxNarrowOut; ++c) {
        if (c < cout) {
          const float* wr = wsm + c * cin + cv * V;
#pragma unroll
          for (int i = 0; i < V; ++i) acc[c] = fmaf(v.v[i], wr[i], acc[c]);
        }
      }
    }
#pragma unroll
    for (int c = 0; c < kMaxNarrowOut; ++c)
      if (c < cout) acc[c] = warp_sum(acc[c]);
    if (lane == 0) {
#pragma unroll
      for (int c = 0; c < kMaxNarrowOut; ++c)
        if (c < cout) Elem<T>::st(y + p * cout + c, acc[c] + (bias ? bias[c] : 0.f));
    }
  }
}

// dx[p][ci] = sum_co dy[p][co] * Wt[ci][co]  (packed_t layout [1][cin][cout])
template <typename T>
__global__ void __launch_bounds__(256) narrow_out_dgrad_kernel(const T* __restrict__ dy, const T* __restrict__ wpt,
                                                               T* __restrict__ dx, int64_t npix, int cin, int cout) {
  constexpr int V = Elem<T>::kVec;
  extern __shared__ float wsm[];  // [cin][cout]
  for (int i = threadIdx.x; i < cout * cin; i += blockDim.x) wsm[i] = Elem<T>::ld(wpt + i);
  __syncthreads();
  const int cvn = cin / V;
  const int64_t total = npix * cvn;
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t p = e / cvn;
    const int c0 = (int)(e - p * cvn) * V;
    float g[kMaxNarrowOut];
#pragma unroll
    for (int c = 0; c < kMaxNarrowOut; ++c) g[c] = c < cout ? Elem<T>::ld(dy + p * cout + c) : 0.f;
    Vec<T> o;
#pragma unroll
    for (int i = 0; i < V; ++i) {
      float s = 0.f;
#pragma unroll
      for (int c = 0; c < kMaxNarrowOut; ++c)
        if (c < cout) s = fmaf(g[c], wsm[(c0 + i) * cout + c], s);
      o.v[i] = s;
    }
    o.store(dx + e * V);
  }
}

// ---- bf16 fast paths of the class-logit conv (256 -> 5 at 128x128: one 268 MB read / write per call at batch 32).
// Weights live in registers (a lane owns one 8-channel vector of the input for all COUT outputs), eight pixels are in
// flight per warp, and the 8 x COUT partial sums are combined by a reduce-scatter butterfly: N/2 + N/4 + ... shuffles
// for N values instead of 5 per value.
template <int N, int OFF>
__device__ __forceinline__ void warp_reduce_scatter(float* v, int lane, int& base, int& cnt) {
  if constexpr (OFF > 0) {
    constexpr int H = (N + 1) / 2;
    const bool up = (lane & OFF) != 0;
#pragma unroll
    for (int i = 0; i < H; ++i) {
      const float lo = v[i], hi = (H + i < N) ? v[H + i] : 0.f;
      const float send = up ? lo : hi, keep = up ? hi : lo;
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, OFF);
    }
    if (up) { base += H; cnt = cnt > H ? cnt - H : 0; }
    else cnt = cnt < H ? cnt : H;
    warp_reduce_scatter<H, OFF / 2>(v, lane, base, cnt);
  }
}
template <int N, int OFF>
struct ScatterWidth { static constexpr int value = ScatterWidth<(N + 1) / 2, OFF / 2>::value; };
template <int N>
struct ScatterWidth<N, 0> { static constexpr int value = N; };

template <int COUT>
__global__ void __launch_bounds__(256) narrow_out_fwd_fast_kernel(const __nv_bfloat16* __restrict__ x,
                                                                  const __nv_bfloat16* __restrict__ wp,
                                                                  const float* __restrict__ bias,
                                                                  __nv_bfloat16* __restrict__ y, int64_t npix, int cin) {
  constexpr int PX = 8, NV = PX * COUT;
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const bool active = lane * 8 < cin;
  float w[COUT][8];
#pragma unroll
  for (int c = 0; c < COUT; ++c)
#pragma unroll
    for (int i = 0; i < 8; ++i) w[c][i] = active ? __bfloat162float(wp[c * cin + lane * 8 + i]) : 0.f;
  float bv[COUT];
#pragma unroll
  for (int c = 0; c < COUT; ++c) bv[c] = bias ? bias[c] : 0.f;

  for (int64_t p0 = warp0 * PX; p0 < npix; p0 += nwarps * PX) {
    uint4 raw[PX];
#pragma unroll
    for (int px = 0; px < PX; ++px)
      raw[px] = (active && p0 + px < npix) ? __ldg(reinterpret_cast<const uint4*>(x + (p0 + px) * cin + lane * 8))
                                           : make_uint4(0u, 0u, 0u, 0u);
    float v[NV];
#pragma unroll
    for (int px = 0; px < PX; ++px) {
      const uint32_t u[4] = {raw[px].x, raw[px].y, raw[px].z, raw[px].w};
      float xv[8];
#pragma unroll
      for (int i = 0; i < 4; ++i) { xv[2 * i] = __uint_as_float(u[i] << 16); xv[2 * i + 1] = __uint_as_float(u[i] & 0xffff0000u); }
#pragma unroll
      for (int c = 0; c < COUT; ++c) {
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) s = fmaf(xv[i], w[c][i], s);
        v[px * COUT + c] = s;
      }
    }
    int base = 0, cnt = NV;
    warp_reduce_scatter<NV, 16>(v, lane, base, cnt);
    constexpr int W = ScatterWidth<NV, 16>::value;
    const int64_t limit = (npix - p0) * COUT;   // outputs of this group that exist
#pragma unroll
    for (int i = 0; i < W; ++i) {
      const int idx = base + i;
      if (i < cnt && idx < limit) y[p0 * COUT + idx] = __float2bfloat16_rn(v[i] + bv[idx % COUT]);
    }
  }
}

// dx[p][ci] = sum_co dy[p][co] * Wt[ci][co]: a thread keeps its channel vector for its whole life (the grid stride is
// a multiple of the vectors per pixel), so its 8 x COUT weights are registers and the loop is 5 broadcast loads,
// 40 FMAs and one 16-byte store per pixel
template <int COUT>
__global__ void __launch_bounds__(256) narrow_out_dgrad_fast_kernel(const __nv_bfloat16* __restrict__ dy,
                                                                    const __nv_bfloat16* __restrict__ wpt,
                                                                    __nv_bfloat16* __restrict__ dx, int64_t npix, int cin) {
  const int cvn = cin / 8;
  const int64_t tid = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  const int64_t nthreads = (int64_t)gridDim.x * blockDim.x;
  const int cv = (int)(tid % cvn);
  float w[8][COUT];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int c = 0; c < COUT; ++c) w[i][c] = __bfloat162float(wpt[(cv * 8 + i) * COUT + c]);
  const int64_t pstep = nthreads / cvn;
#pragma unroll 2
  for (int64_t p = tid / cvn; p < npix; p += pstep) {
    float g[COUT];
#pragma unroll
    for (int c = 0; c < COUT; ++c) g[c] = __bfloat162float(dy[p * COUT + c]);
    uint32_t pk[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float s0 = 0.f, s1 = 0.f;
#pragma unroll
      for (int c = 0; c < COUT; ++c) { s0 = fmaf(g[c], w[2 * i][c], s0); s1 = fmaf(g[c], w[2 * i + 1][c], s1); }
      __nv_bfloat162 h = __floats2bfloat162_rn(s0, s1);
      pk[i] = *reinterpret_cast<uint32_t*>(&h);
    }
    *reinterpret_cast<uint4*>(dx + p * cin + cv * 8) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
  }
}

template <typename T>
struct NarrowWgradF {
  static constexpr int NACC = kMaxNarrowOut;
  const T* x;
  const T* dy;
  int cin, cout;
  __device__ __forceinline__ void operator()(int64_t row, int c0, float (&acc)[kMaxNarrowOut][Elem<T>::kVec]) const {
    Vec<T> v;
    v.load(x + row * cin + c0);
#pragma unroll
    for (int c = 0; c < kMaxNarrowOut; ++c) {
      if (c < cout) {
        const float g = Elem<T>::ld(dy + row * cout + c);
#pragma unroll
        for (int i = 0; i < Vec<T>::N; ++i) acc[c][i] = fmaf(g, v.v[i], acc[c][i]);
      }
    }
  }
};

__global__ void narrow_wgrad_finish_kernel(const double* __restrict__ acc, float* __restrict__ dw, int cin, int cout) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < cin * cout) dw[i] += (float)acc[i];  // acc is [co][cin] == packed [1][cout][cin]
}

// ---------------------------------------------------------------- narrow in: im2col for large filters
// patches[p][k], k = tap*cin + ci (zero beyond taps*cin up to kpad): turns the 7x7 stride-2 3->64
// ResNet stem (Graph_Structure(data_augmentation).py:136, torchvision resnet101.conv1) into a 1x1
// tensor-core GEMM with K = kpad.
template <typename T>
__global__ void im2col_narrow_kernel(const T* __restrict__ x, T* __restrict__ y, NGeom g, int kpad) {
  constexpr int V = Elem<T>::kVec;
  const int kv = kpad / V;
  const int K = g.kh * g.kw * g.cin;
  const int64_t total = (int64_t)g.n * g.ho * g.wo * kv;
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int k0 = (int)(e % kv) * V;
    const int64_t p = e / kv;
    const int ox = (int)(p % g.wo), oy = (int)((p / g.wo) % g.ho), nn = (int)(p / ((int64_t)g.wo * g.ho));
    Vec<T> o;
#pragma unroll
    for (int i = 0; i < V; ++i) {
      const int k = k0 + i;
      float v = 0.f;
      if (k < K) {
        const int tap = k / g.cin, ci = k - tap * g.cin;
        const int iy = oy * g.stride - g.pad + (tap / g.kw) * g.dil;
        const int ix = ox * g.stride - g.pad + (tap % g.kw) * g.dil;
        if (iy >= 0 && iy < g.h && ix >= 0 && ix < g.w) v = Elem<T>::ld(x + (((size_t)nn * g.h + iy) * g.w + ix) * g.cin + ci);
      }
      o.v[i] = v;
    }
    o.store(y + e * V);
  }
}

// Same result, row-wise: a thread keeps ONE 8-wide slice of the patch vector (its taps / channels are decoded once),
// a block walks whole output rows, so the inner loop is 8 bounds checks + loads and one 16-byte store per pixel.
template <typename T>
__global__ void __launch_bounds__(256) im2col_narrow_rows_kernel(const T* __restrict__ x, T* __restrict__ y, NGeom g,
                                                                 int kpad) {
  constexpr int V = Elem<T>::kVec;
  const int kv = kpad / V;                        // <= 32 (host-checked)
  const int K = g.kh * g.kw * g.cin;
  // (pixel lane, slice) = divmod(thread, kv): with kv = 19 slices (the ResNet stem's 7x7x3 -> 152) that keeps 247 of the
  // 256 threads busy; one warp per pixel lane left 13 of 32 lanes idle in this instruction-bound kernel
  const int kvi = threadIdx.x % kv, pl = threadIdx.x / kv;
  const int plN = 256 / kv;
  if (pl >= plN) return;
  int dy[V], dx[V], ci[V];
  bool kok[V];
#pragma unroll
  for (int i = 0; i < V; ++i) {
    const int k = kvi * V + i;
    kok[i] = k < K;
    const int tap = kok[i] ? k / g.cin : 0;
    ci[i] = kok[i] ? k - tap * g.cin : 0;
    dy[i] = (tap / g.kw) * g.dil - g.pad;
    dx[i] = (tap % g.kw) * g.dil - g.pad;
  }
  const int rows = g.n * g.ho;
  for (int row = blockIdx.x; row < rows; row += gridDim.x) {
    const int nn = row / g.ho, oy = row - nn * g.ho;
    const T* rp[V];
    bool rok[V];
#pragma unroll
    for (int i = 0; i < V; ++i) {
      const int iy = oy * g.stride + dy[i];
      rok[i] = kok[i] && iy >= 0 && iy < g.h;
      rp[i] = x + ((size_t)nn * g.h + (rok[i] ? iy : 0)) * g.w * g.cin + ci[i];
    }
    T* out = y + ((size_t)row * g.wo) * kpad + kvi * V;
    for (int ox = pl; ox < g.wo; ox += plN) {
      Vec<T> o;
#pragma unroll
      for (int i = 0; i < V; ++i) {
        const int ix = ox * g.stride + dx[i];
        o.v[i] = (rok[i] && ix >= 0 && ix < g.w) ? Elem<T>::ld(rp[i] + (size_t)ix * g.cin) : 0.f;
      }
      o.store(out + (size_t)ox * kpad);
    }
  }
}

static bool is_stem(const cvx_conv_desc* d) {
  return d->cin <= 4 && d->cout == kStemCout && d->kh * d->kw * d->cin <= kStemMaxK;
}
static bool is_narrow_out(const cvx_conv_desc* d) {
  const int vec = d->dtype == CVX_F32 ? 4 : 8;
  return d->cout <= kMaxNarrowOut && d->kh == 1 && d->kw == 1 && d->stride == 1 && d->pad == 0 && d->cin % vec == 0 &&
         d->cin <= kMaxNarrowCin;
}
static NGeom ngeom(const cvx_conv_desc* d) {
  return NGeom{d->n, d->h, d->w, d->cin, d->cout, d->kh, d->kw, d->stride, d->pad, d->dil, d->ho, d->wo};
}
static inline int cap_blocks(int64_t want, int per_sm) {
  const int64_t cap = (int64_t)kNumSMs * per_sm;
  return (int)(want > cap ? cap : (want < 1 ? 1 : want));
}

int narrow_conv_fwd(const cvx_conv_desc* d, const void* x, const void* wp, const float* bias, void* y, cudaStream_t st) {
  const int64_t npix = (int64_t)d->n * d->ho * d->wo;
  if (is_stem(d) && bias == nullptr) {
    const NGeom g = ngeom(d);
    CVX_DISPATCH_DTYPE(d->dtype, T, (stem_fwd_kernel<T><<<cap_blocks(ceil_div64(npix, 256), 8), 256, 0, st>>>(
                                        (const T*)x, (const T*)wp, (T*)y, g)));
    CVX_LAUNCH_OK();
    return CVX_OK;
  }
  if (is_narrow_out(d)) {
    if (d->dtype == CVX_BF16 && d->cout == 5 && d->cin <= 256) {   // the reference's 5-class head (deeplabv3_plus.py:167)
      narrow_out_fwd_fast_kernel<5><<<cap_blocks(ceil_div64(npix, 64), 8), 256, 0, st>>>(
          (const __nv_bfloat16*)x, (const __nv_bfloat16*)wp, bias, (__nv_bfloat16*)y, npix, d->cin);
      CVX_LAUNCH_OK();
      return CVX_OK;
    }
    const size_t smem = sizeof(float) * d->cin * d->cout;
    CVX_DISPATCH_DTYPE(d->dtype, T, (narrow_out_fwd_kernel<T><<<cap_blocks(ceil_div64(npix, 8), 8), 256, smem, st>>>(
                                        (const T*)x, (const T*)wp, bias, (T*)y, npix, d->cin, d->cout)));
    CVX_LAUNCH_OK();
    return CVX_OK;
  }
  return CVX_EUNSUPPORTED;
}

int narrow_conv_dgrad(const cvx_conv_desc* d, const void* dy, const void* wpt, void* dx, cudaStream_t st) {
  if (!is_narrow_out(d)) return CVX_EUNSUPPORTED;
  const int64_t npix = (int64_t)d->n * d->h * d->w;
  const int vec = d->dtype == CVX_F32 ? 4 : 8;
  if (d->dtype == CVX_BF16 && d->cout == 5 && 256 % (d->cin / 8) == 0) {
    narrow_out_dgrad_fast_kernel<5><<<cap_blocks(ceil_div64(npix * (d->cin / 8), 256 * 4), 8), 256, 0, st>>>(
        (const __nv_bfloat16*)dy, (const __nv_bfloat16*)wpt, (__nv_bfloat16*)dx, npix, d->cin);
    CVX_LAUNCH_OK();
    return CVX_OK;
  }
  const size_t smem = sizeof(float) * d->cin * d->cout;
  CVX_DISPATCH_DTYPE(d->dtype, T, (narrow_out_dgrad_kernel<T><<<cap_blocks(ceil_div64(npix * (d->cin / vec), 256), 8), 256, smem, st>>>(
                                      (const T*)dy, (const T*)wpt, (T*)dx, npix, d->cin, d->cout)));
  CVX_LAUNCH_OK();
  return CVX_OK;
}

// scratch for the narrow-out weight gradient lives in a small static device buffer per process
int narrow_conv_wgrad(const cvx_conv_desc* d, const void* x, const void* dy, float* dwp, cudaStream_t st) {
  if (is_stem(d)) {
    const NGeom g = ngeom(d);
    const int64_t npix = (int64_t)d->n * d->ho * d->wo;
    if (d->dtype == CVX_BF16 && d->kh * d->kw * d->cin <= 32 &&
        (int64_t)d->n * d->h * d->w * d->cin < (1ll << 40)) {
      stem_wgrad_mma_kernel<<<cap_blocks(ceil_div64(npix, 64), 8), 128, 0, st>>>(
          (const __nv_bfloat16*)x, (const __nv_bfloat16*)dy, dwp, g);
      CVX_LAUNCH_OK();
      return CVX_OK;
    }
    CVX_DISPATCH_DTYPE(d->dtype, T, (stem_wgrad_kernel<T><<<cap_blocks(ceil_div64(npix, 64), 4), 288, 0, st>>>(
                                        (const T*)x, (const T*)dy, dwp, g)));
    CVX_LAUNCH_OK();
    return CVX_OK;
  }
  if (is_narrow_out(d)) {
    static double* scratch = nullptr;
    if (!scratch) CVX_CUDA_OK(cudaMalloc(&scratch, sizeof(double) * kMaxNarrowOut * kMaxNarrowCin));
    CVX_CUDA_OK(cudaMemsetAsync(scratch, 0, sizeof(double) * kMaxNarrowOut * d->cin, st));
    const int64_t rows = (int64_t)d->n * d->ho * d->wo;
    int rc = CVX_OK;
    CVX_DISPATCH_DTYPE(d->dtype, T, rc = (colreduce_launch<T, NarrowWgradF<T>, 256, 2>(
                                        NarrowWgradF<T>{(const T*)x, (const T*)dy, d->cin, d->cout}, rows, d->cin, scratch, st)));
    if (rc) return rc;
    narrow_wgrad_finish_kernel<<<(d->cin * d->cout + 255) / 256, 256, 0, st>>>(scratch, dwp, d->cin, d->cout);
    CVX_LAUNCH_OK();
    return CVX_OK;
  }
  return CVX_EUNSUPPORTED;
}

}  // namespace cvx

extern "C" int cvx_im2col_narrow(const cvx_conv_desc* d, const void* x, void* patches, int kpad, void* stream) {
  using namespace cvx;
  CVX_CHECK_ARG(d && x && patches, "im2col_narrow: null pointer");
  const int vec = d->dtype == CVX_F32 ? 4 : 8;
  CVX_CHECK_ARG(d->cin <= 4 && kpad % vec == 0 && kpad >= d->kh * d->kw * d->cin, "im2col_narrow: bad geometry");
  const NGeom g = ngeom(d);
  const int64_t total = (int64_t)d->n * d->ho * d->wo * (kpad / vec);
  if (kpad / vec <= 32 && (int64_t)d->n * d->ho < (1 << 30)) {
    CVX_DISPATCH_DTYPE(d->dtype, T, (im2col_narrow_rows_kernel<T><<<cap_blocks((int64_t)d->n * d->ho, 8), 256, 0, as_stream(stream)>>>(
                                        (const T*)x, (T*)patches, g, kpad)));
    CVX_LAUNCH_OK();
    return CVX_OK;
  }
  CVX_DISPATCH_DTYPE(d->dtype, T, (im2col_narrow_kernel<T><<<cap_blocks(ceil_div64(total, 256), 16), 256, 0, as_stream(stream)>>>(
                                      (const T*)x, (T*)patches, g, kpad)));
  CVX_LAUNCH_OK();
  return CVX_OK;
}
