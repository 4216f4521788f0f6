// Grouped fp32 GEMM for the linear layers of the multimodal fusion head (BASELINE
// north_star: "the modality branches run as one grouped GEMM launch").
//
// The reference builds one SAGEConv / gate MLP / head MLP per modality (MultiModal Prediction/Four_Modal/my_mae_model.py
// :404-416, 544, 706-769) and runs them one nn.Linear at a time.  The modality branches have the same topology but their
// own weights, so a layer of the head is a GROUP of independent small GEMMs (M = patients x nodes <= a few thousand rows,
// N, K <= 2048).  One launch takes up to kMaxProblems problems - descriptors travel by value in the kernel parameters, so
// there is no device-side table to fill and the launch is CUDA-graph capturable - and its grid is the concatenation of all
// problems' 64x64 output tiles, which is what fills the 148 SMs (a single 256 x 512 layer has 32 tiles).
//
// Each problem is C[M,N] = A[M,K] * B[N,K]^T (+ bias[N]) with element strides for both operands, so the same kernel is
//   forward   Y  = X W^T + b        A = X  (k contiguous)   B = W   (k contiguous)
//   dgrad     dX = dY W             A = dY (k contiguous)   B = W^T (n contiguous)
//   wgrad     dW = dY^T X           A = dY^T (m contiguous) B = X^T (n contiguous)   + rowsum(A) = the bias gradient
// Arithmetic is fp32 FMA with fp32 accumulation (the head's parity bar against the reference is 2e-4, which bf16/tf32
// operands do not meet); every output element is one fixed-order sum computed by one thread - no split-K, no atomics -
// so the head's gradients stay bit-reproducible.  Weights are consumed in nn.Linear's own [out, in] layout: nothing is
// packed, transposed or unpacked around the GEMM.
#include "common.cuh"
#include <stdlib.h>

namespace cvx {

constexpr int kGN = 64, kGK = 32;   // two 16-deep fragments per operand and iteration: 8 loads in flight per thread

struct GemmBatch {
  cvx_gemm_problem p[CVX_MAX_GEMM_PROBLEMS];
  int tile_start[CVX_MAX_GEMM_PROBLEMS + 1];
  int count;
};

// stage a ROWS x 16 (k) slice of an operand tile into registers: 4 elements per thread of a 64-row slice (256 threads),
// or of a 32-row slice when only threads 0..127 take part.  k-contiguous operands are read as 4 consecutive k of one row
// (16-byte segments), row-contiguous operands as 4 consecutive rows of one k.
struct Frag { float v[4]; };

__device__ __forceinline__ Frag load_frag(const float* __restrict__ base, int64_t ld_r, int64_t ld_k, int r0, int rows,
                                          int k0, int K, int t, int tile_rows) {
  Frag f;
  if (ld_k == 1) {
    const int r = r0 + (t >> 2), k = k0 + (t & 3) * 4;
    const bool in = (t >> 2) < tile_rows && r < rows;
    const float* ptr = base + (int64_t)r * ld_r + k;
#pragma unroll
    for (int i = 0; i < 4; ++i) f.v[i] = (in && k + i < K) ? __ldg(ptr + i) : 0.f;
  } else {
    const int q = tile_rows >> 2;                          // threads per k: 16 (64 rows) or 8 (32 rows)
    const int k = k0 + t / q, r = r0 + (t % q) * 4;
    const bool in = t / q < 16 && k < K;
    const float* ptr = base + (int64_t)k * ld_k + (int64_t)r * ld_r;
#pragma unroll
    for (int i = 0; i < 4; ++i) f.v[i] = (in && r + i < rows) ? __ldg(ptr + (int64_t)i * ld_r) : 0.f;
  }
  return f;
}

template <int ROWS>
__device__ __forceinline__ void store_frag(float (*S)[ROWS + 4], const Frag& f, bool k_contig, int t) {
  if (k_contig) {
    const int r = t >> 2, k = (t & 3) * 4;
    if (r < ROWS) {
#pragma unroll
      for (int i = 0; i < 4; ++i) S[k + i][r] = f.v[i];
    }
  } else {
    constexpr int q = ROWS / 4;
    const int k = t / q, r = (t % q) * 4;
    if (k < 16) {
#pragma unroll
      for (int i = 0; i < 4; ++i) S[k][r + i] = f.v[i];
    }
  }
}

// BM = 64: 4 x 4 outputs per thread.  BM = 32: 2 x 4 outputs per thread - twice the tiles for layers whose 64-row tiling
// would leave most of the 148 SMs with one or two CTAs (the auto-encoder's and the heads' GEMMs have 24 - 200 tiles).
template <int BM>
__global__ void __launch_bounds__(256, 2) gemm_grouped_kernel(const __grid_constant__ GemmBatch batch) {
  constexpr int TM = BM / 16;
  __shared__ float As[kGK][BM + 4];
  __shared__ float Bs[kGK][kGN + 4];
  pdl_trigger();
  // which problem does this tile belong to?  (<= 16 entries: a linear scan of kernel parameters)
  int pi = 0;
  while (pi + 1 < batch.count && (int)blockIdx.x >= batch.tile_start[pi + 1]) ++pi;
  const cvx_gemm_problem& P = batch.p[pi];
  const int tile = blockIdx.x - batch.tile_start[pi];
  const int tiles_n = (P.n + kGN - 1) / kGN;
  const int m0 = (tile / tiles_n) * BM, n0 = (tile % tiles_n) * kGN;
  const int t = threadIdx.x, ty = t >> 4, tx = t & 15;
  const bool a_kc = P.lda_k == 1, b_kc = P.ldb_k == 1;

  float acc[TM][4];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  float rs[TM];
#pragma unroll
  for (int i = 0; i < TM; ++i) rs[i] = 0.f;                // row sums of A (bias gradient of the wgrad problems)
  const bool want_rs = P.rowsum != nullptr && n0 == 0 && tx == 0;
  pdl_wait();

  // software pipeline: the loads of k-block i+1 (two 16-deep fragments per operand) are in flight while block i is multiplied
  Frag fa0 = load_frag(P.a, P.lda_m, P.lda_k, m0, P.m, 0, P.k, t, BM), fa1 = load_frag(P.a, P.lda_m, P.lda_k, m0, P.m, 16, P.k, t, BM);
  Frag fb0 = load_frag(P.b, P.ldb_n, P.ldb_k, n0, P.n, 0, P.k, t, kGN), fb1 = load_frag(P.b, P.ldb_n, P.ldb_k, n0, P.n, 16, P.k, t, kGN);
  for (int k0 = 0; k0 < P.k; k0 += kGK) {
    store_frag<BM>(As, fa0, a_kc, t);
    store_frag<BM>(As + 16, fa1, a_kc, t);
    store_frag<kGN>(Bs, fb0, b_kc, t);
    store_frag<kGN>(Bs + 16, fb1, b_kc, t);
    __syncthreads();
    if (k0 + kGK < P.k) {
      fa0 = load_frag(P.a, P.lda_m, P.lda_k, m0, P.m, k0 + kGK, P.k, t, BM);
      fa1 = load_frag(P.a, P.lda_m, P.lda_k, m0, P.m, k0 + kGK + 16, P.k, t, BM);
      fb0 = load_frag(P.b, P.ldb_n, P.ldb_k, n0, P.n, k0 + kGK, P.k, t, kGN);
      fb1 = load_frag(P.b, P.ldb_n, P.ldb_k, n0, P.n, k0 + kGK + 16, P.k, t, kGN);
    }
    const int kmax = P.k - k0 < kGK ? ((P.k - k0 + 3) & ~3) : kGK;   // short last block: skip the all-zero tail
#pragma unroll 8
    for (int kk = 0; kk < kmax; ++kk) {
      float av[TM];
      if (TM == 4) {
        const float4 a = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
        av[0] = a.x; av[1] = a.y; av[TM - 2] = a.z; av[TM - 1] = a.w;
      } else {
        const float2 a = *reinterpret_cast<const float2*>(&As[kk][ty * 2]);
        av[0] = a.x; av[1] = a.y;
      }
      const float4 b = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
      const float bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
      if (want_rs) {
#pragma unroll
        for (int i = 0; i < TM; ++i) rs[i] += av[i];
      }
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    const int m = m0 + ty * TM + i;
    if (m >= P.m) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n < P.n) P.c[(int64_t)m * P.ldc + n] = acc[i][j] + (P.bias ? __ldg(P.bias + n) : 0.f);
    }
    if (want_rs) P.rowsum[m] = rs[i];
  }
}

}  // namespace cvx

using namespace cvx;

extern "C" int cvx_gemm_grouped(const cvx_gemm_problem* problems, int count, void* stream) {
  CVX_CHECK_ARG(problems && count > 0 && count <= CVX_MAX_GEMM_PROBLEMS, "gemm_grouped: 1..%d problems per launch (got %d)",
                CVX_MAX_GEMM_PROBLEMS, count);
  GemmBatch batch;
  int tiles64 = 0;
  for (int i = 0; i < count; ++i) {
    const cvx_gemm_problem& p = problems[i];
    CVX_CHECK_ARG(p.a && p.b && p.c && p.m > 0 && p.n > 0 && p.k > 0, "gemm_grouped: problem %d has a null operand or an empty shape", i);
    CVX_CHECK_ARG((p.lda_k == 1 || p.lda_m == 1) && (p.ldb_k == 1 || p.ldb_n == 1) && p.ldc >= p.n,
                  "gemm_grouped: problem %d: each operand must be contiguous along rows or along k", i);
    batch.p[i] = p;
    tiles64 += ((p.m + 63) / 64) * ((p.n + kGN - 1) / kGN);
  }
  // 64-row tiles when they already give every SM two CTAs, else 32-row tiles (twice as many, same arithmetic order)
  static const int force_bm = [] { const char* e = getenv("CERVIX_GEMM_BM"); return e ? atoi(e) : 0; }();
  const int bm = force_bm == 32 || force_bm == 64 ? force_bm : (tiles64 >= 2 * kNumSMs ? 64 : 32);
  int tiles = 0;
  for (int i = 0; i < count; ++i) {
    batch.tile_start[i] = tiles;
    tiles += ((batch.p[i].m + bm - 1) / bm) * ((batch.p[i].n + kGN - 1) / kGN);
  }
  batch.tile_start[count] = tiles;
  batch.count = count;
  if (bm == 64) launch_pdl(gemm_grouped_kernel<64>, dim3(tiles), dim3(256), 0, as_stream(stream), batch);
  else launch_pdl(gemm_grouped_kernel<32>, dim3(tiles), dim3(256), 0, as_stream(stream), batch);
  CVX_LAUNCH_OK();
  return CVX_OK;
}
