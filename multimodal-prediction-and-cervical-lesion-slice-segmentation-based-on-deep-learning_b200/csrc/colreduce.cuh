// Per-channel (column) reductions over an NHWC tensor viewed as [rows][C].
//
// Bandwidth-bound: every thread owns one 16-byte channel vector and walks down the rows, so
// a warp always reads whole contiguous 512-byte spans (4 independent rows in flight per
// thread); partial sums are combined once per block in shared memory in a FIXED order (row lane 0,
// 1, 2, ... - no shared atomics, so a block's partial is bit-reproducible) and then through fp64
// global atomics per channel.  The fp64 sum of a few hundred fp32 partials of comparable magnitude is
// exact (24-bit mantissas inside a 53-bit accumulator), hence independent of the arrival order.
// Few, long-lived blocks keep that epilogue negligible.
#pragma once

#include "common.cuh"

#include <type_traits>

namespace cvx {

// F must provide:  static constexpr int NACC;
//   __device__ void operator()(int64_t row, int c0, float (&acc)[NACC][VEC]) const;
// or, with per-thread channel constants held in registers for the whole walk (a thread owns one channel vector):
//   struct Ctx;  __device__ void init(int c0, Ctx&) const;
//   __device__ void operator()(int64_t row, int c0, float (&acc)[NACC][VEC], const Ctx&) const;
//   __device__ void finish(float (&acc)[NACC][VEC], const Ctx&) const;
template <typename F, typename = void> struct ColHasCtx : std::false_type {};
template <typename F> struct ColHasCtx<F, std::void_t<typename F::Ctx>> : std::true_type {};
template <typename T, typename F, int NT, int MINB>
__global__ void __launch_bounds__(NT, MINB) colreduce_kernel(F f, int64_t rows, int C, int colchunk_vecs,
                                                             int64_t rows_per_block, double* __restrict__ out) {
  constexpr int VEC = Elem<T>::kVec;
  constexpr int NACC = F::NACC;
  extern __shared__ float sm_acc[];  // [NACC][colchunk_vecs*VEC]
  pdl_trigger();
  const int chunk_elems = colchunk_vecs * VEC;
  for (int i = threadIdx.x; i < NACC * chunk_elems; i += NT) sm_acc[i] = 0.f;
  __syncthreads();
  pdl_wait();

  const int lanes = NT / colchunk_vecs;  // row lanes per block
  const int cv = threadIdx.x % colchunk_vecs;
  const int lane = threadIdx.x / colchunk_vecs;
  const int cvec = blockIdx.y * colchunk_vecs + cv;
  const bool active = lane < lanes && cvec * VEC < C;

  float acc[NACC][VEC];
#pragma unroll
  for (int a = 0; a < NACC; ++a)
#pragma unroll
    for (int i = 0; i < VEC; ++i) acc[a][i] = 0.f;

  if (active) {
    const int64_t r0 = (int64_t)blockIdx.x * rows_per_block;
    int64_t r1 = r0 + rows_per_block;
    if (r1 > rows) r1 = rows;
    if constexpr (ColHasCtx<F>::value) {
      typename F::Ctx ctx;
      f.init(cvec * VEC, ctx);
#pragma unroll 8
      for (int64_t r = r0 + lane; r < r1; r += lanes) f(r, cvec * VEC, acc, ctx);
      f.finish(acc, ctx);
    } else {
#pragma unroll 8
      for (int64_t r = r0 + lane; r < r1; r += lanes) f(r, cvec * VEC, acc);
    }
  }
  // fixed-order combine (no shared atomics).  When the chunk width divides the warp, the row lanes of a warp are first
  // folded by a shuffle tree; then the remaining contributors add their registers to the block tile one after the
  // other (the threads of one turn own distinct columns).
  int turn = lane, turns = lanes;
  bool contributes = active;
  if ((32 % colchunk_vecs) == 0) {
#pragma unroll
    for (int a = 0; a < NACC; ++a)
#pragma unroll
      for (int i = 0; i < VEC; ++i) {
        float v = acc[a][i];
        for (int o = 16; o >= colchunk_vecs; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        acc[a][i] = v;
      }
    turn = threadIdx.x >> 5;
    turns = NT / 32;
    contributes = active && (threadIdx.x & 31) < colchunk_vecs;
  }
  for (int l = 0; l < turns; ++l) {
    if (contributes && turn == l) {
#pragma unroll
      for (int a = 0; a < NACC; ++a)
#pragma unroll
        for (int i = 0; i < VEC; ++i) sm_acc[a * chunk_elems + cv * VEC + i] += acc[a][i];
    }
    __syncthreads();
  }
  for (int i = threadIdx.x; i < NACC * chunk_elems; i += NT) {
    const int a = i / chunk_elems, e = i % chunk_elems;
    const int c = blockIdx.y * chunk_elems + e;
    if (c < C) atomicAdd(out + (size_t)a * C + c, (double)sm_acc[i]);
  }
}

// Launch helper.  `out` ([NACC][C] doubles) must be zeroed by the caller.
template <typename T, typename F, int NT = 256, int MINB = 4>
static int colreduce_launch(const F& f, int64_t rows, int C, double* out, cudaStream_t stream) {
  constexpr int VEC = Elem<T>::kVec;
  if (C % VEC != 0) {
    set_error("channel count %d is not a multiple of the %d-element vector width", C, VEC);
    return CVX_EINVAL;
  }
  const int cvn = C / VEC;
  int maxchunk = (12288 / F::NACC) / VEC;  // keep the block's partial-sum tile within 48 KB
  if (maxchunk > NT) maxchunk = NT;
  // column chunk = channel vectors per block: pick the width that keeps the most threads busy (row lanes x chunk of NT,
  // times the fill of the last chunk) among chunks of >= 8 vectors (128 contiguous bytes per row segment); e.g. 728
  // channels = 91 vectors: one 91-wide chunk uses 182 of 256 threads, four 23-wide chunks use 253
  int colchunk = cvn < maxchunk ? cvn : maxchunk;
  {
    double best = -1.0;
    const int lo = cvn < 8 ? cvn : 8;
    for (int cc = colchunk; cc >= lo; --cc) {
      const int yc = (cvn + cc - 1) / cc;
      const double use = (double)((NT / cc) * cc) / NT * (double)cvn / (double)(yc * cc);
      if (use > best + 0.02) { best = use; colchunk = cc; }   // prefer wider chunks unless clearly better
    }
  }
  const int lanes = NT / colchunk;
  const int ychunks = (cvn + colchunk - 1) / colchunk;
  // one wave of resident blocks, each walking >= 16 rows per lane
  int64_t want_blocks = (int64_t)kNumSMs * MINB / ychunks;
  if (want_blocks < 1) want_blocks = 1;
  int64_t rpb = ceil_div64(rows, want_blocks);
  const int64_t min_rpb = (int64_t)lanes * 16;
  if (rpb < min_rpb) rpb = min_rpb;
  rpb = ceil_div64(rpb, lanes) * lanes;
  const int64_t gx = ceil_div64(rows, rpb);
  dim3 grid((unsigned)gx, ychunks);
  const size_t smem = sizeof(float) * F::NACC * colchunk * VEC;
  launch_pdl(colreduce_kernel<T, F, NT, MINB>, grid, dim3(NT), smem, stream, f, rows, C, colchunk, rpb, out);
  ++g_kernel_launches;
  cudaError_t e = cudaPeekAtLastError();
  if (e != cudaSuccess) {
    set_error("colreduce launch failed: %s", cudaGetErrorString(e));
    return CVX_ECUDA;
  }
  return CVX_OK;
}

}  // namespace cvx
