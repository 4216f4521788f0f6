// Small fp32 row-wise operators of the multimodal fusion head.
//
// The reference runs this head one patient at a time through hundreds of tiny ATen / PyG launches
// (MultiModal Prediction/Four_Modal/my_mae_model.py:500-793).  Here a batch of G patients is laid out
// as [G * nodes, C] row matrices and every graph-level operator works on fixed-size row segments, so
// one launch covers the whole batch.  Everything is latency/bandwidth bound and tiny (<= 16 x 512 floats
// per patient); one CTA per segment / graph, fp32 throughout.  Linear layers go through the 1x1 path of
// the convolution kernels (conv_simt.cu / conv_narrow.cu).
#include "common.cuh"

namespace cvx {

__device__ __forceinline__ float block_sum(float v, float* sh) {
  v = warp_sum(v);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31, nw = (blockDim.x + 31) >> 5;
  __syncthreads();
  if (l == 0) sh[w] = v;
  __syncthreads();
  float t = (threadIdx.x < nw) ? sh[threadIdx.x] : 0.f;
  if (w == 0) {
    t = warp_sum(t);
    if (l == 0) sh[0] = t;
  }
  __syncthreads();
  return sh[0];
}

// ---- segment LayerNorm -----------------------------------------------------------------------
// mode 0: PyG LayerNorm(mode='graph'): (x - mean) / (std + eps) over all seg*C elements of a segment
// mode 1: nn.LayerNorm: (x - mean) / sqrt(var + eps)   (use seg = 1)
// stats[g] = {mean, r (the multiplier), std}
__global__ void seg_layernorm_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                         const float* __restrict__ b, float* __restrict__ y, float* __restrict__ stats,
                                         int seg, int C, float eps, int mode) {
  pdl_trigger();
  pdl_wait();
  __shared__ float sh[32];
  const int g = blockIdx.x;
  const int n = seg * C;
  const float* xs = x + (size_t)g * n;
  float s = 0.f;
  for (int i = threadIdx.x; i < n; i += blockDim.x) s += xs[i];
  const float mean = block_sum(s, sh) / n;
  float v = 0.f;
  for (int i = threadIdx.x; i < n; i += blockDim.x) { const float d = xs[i] - mean; v = fmaf(d, d, v); }
  const float var = block_sum(v, sh) / n;
  const float sd = sqrtf(var);
  const float r = mode == 0 ? 1.f / (sd + eps) : rsqrtf(var + eps);
  if (threadIdx.x == 0) { stats[3 * g] = mean; stats[3 * g + 1] = r; stats[3 * g + 2] = sd; }
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const int c = i % C;
    y[(size_t)g * n + i] = (xs[i] - mean) * r * w[c] + b[c];
  }
}

// dx only; the affine gradients are a column reduction over all rows (ln_param_grad_kernel below, fixed summation order)
__global__ void seg_layernorm_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ x,
                                         const float* __restrict__ w, const float* __restrict__ stats,
                                         float* __restrict__ dx, int seg, int C, float eps, int mode) {
  pdl_trigger();
  pdl_wait();
  __shared__ float sh[32];
  const int g = blockIdx.x;
  const int n = seg * C;
  const float mean = stats[3 * g], r = stats[3 * g + 1], sd = stats[3 * g + 2];
  const float* xs = x + (size_t)g * n;
  const float* gs = dy + (size_t)g * n;
  float sg = 0.f, sgx = 0.f;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const float gi = gs[i] * w[i % C];
    sg += gi;
    sgx = fmaf(gi, xs[i] - mean, sgx);
  }
  const float Sg = block_sum(sg, sh);
  const float Sgx = block_sum(sgx, sh);
  // d r / d x_i = -K' * xc_i : PyG r = 1/(sd+eps) -> r^2/(n*sd) ; torch r = (var+eps)^-1/2 -> r^3/n
  const float K = mode == 0 ? (sd > 0.f ? Sgx * r * r / (n * sd) : 0.f) : Sgx * r * r * r / n;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const int c = i % C;
    const float xc = xs[i] - mean;
    dx[(size_t)g * n + i] = r * (gs[i] * w[c] - Sg / n) - xc * K;
  }
}

// dw[c] = sum_rows dy * (x - mean_g) * r_g ; db[c] = sum_rows dy.  32 channels x 8 row lanes per block; every partial is
// summed in a fixed order (no atomics), so the head's gradients are bit-reproducible run to run.
__global__ void __launch_bounds__(256) ln_param_grad_kernel(const float* __restrict__ dy, const float* __restrict__ x,
                                                            const float* __restrict__ stats, float* __restrict__ dw,
                                                            float* __restrict__ db, int rows, int seg, int C) {
  pdl_trigger();
  pdl_wait();
  __shared__ float pw[8][33], pb[8][33];
  const int cx = threadIdx.x & 31, ry = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + cx;
  float aw = 0.f, ab = 0.f;
  if (c < C) {
    for (int r = ry; r < rows; r += 8) {
      const int g = r / seg;
      const float gi = dy[(size_t)r * C + c];
      aw = fmaf(gi * (x[(size_t)r * C + c] - stats[3 * g]), stats[3 * g + 1], aw);
      ab += gi;
    }
  }
  pw[ry][cx] = aw;
  pb[ry][cx] = ab;
  __syncthreads();
  if (ry == 0 && c < C) {
    float sw = 0.f, sb = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) { sw += pw[k][cx]; sb += pb[k][cx]; }
    if (dw) dw[c] = sw;
    if (db) db[c] = sb;
  }
}

// ---- GELU (erf form, nn.GELU default) ------------------------------------------------------------
__global__ void gelu_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, int64_t n) {
  pdl_trigger();
  pdl_wait();
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float v = x[i];
    y[i] = 0.5f * v * (1.f + erff(v * 0.70710678118654752f));
  }
}
__global__ void gelu_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ x, float* __restrict__ dx, int64_t n) {
  pdl_trigger();
  pdl_wait();
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float v = x[i];
    const float cdf = 0.5f * (1.f + erff(v * 0.70710678118654752f));
    const float pdf = 0.3989422804014327f * expf(-0.5f * v * v);
    dx[i] = dy[i] * (cdf + v * pdf);
  }
}

// ---- fixed-topology graph gather: out[g,i,:] = sum_e w[e] * x[g, col[e], :], e in [rowptr[i], rowptr[i+1]) ----
__global__ void graph_gather_kernel(const float* __restrict__ x, float* __restrict__ out, int G, int nodes, int C,
                                    const int* __restrict__ rowptr, const int* __restrict__ col,
                                    const float* __restrict__ w) {
  pdl_trigger();
  pdl_wait();
  const int64_t total = (int64_t)G * nodes * C;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const int node = (int)((i / C) % nodes);
    const int g = (int)(i / ((int64_t)C * nodes));
    float acc = 0.f;
    for (int e = rowptr[node]; e < rowptr[node + 1]; ++e) acc = fmaf(w[e], x[((size_t)g * nodes + col[e]) * C + c], acc);
    out[i] = acc;
  }
}

// ---- gated attention pooling (my_GlobalAttention): one CTA per graph -------------------------------
__global__ void gate_pool_fwd_kernel(const float* __restrict__ x, const float* __restrict__ gate, float* __restrict__ pooled,
                                     float* __restrict__ att, int seg, int C) {
  pdl_trigger();
  pdl_wait();
  __shared__ float a[64];
  const int g = blockIdx.x;
  if (threadIdx.x == 0) {
    float mx = -INFINITY;
    for (int i = 0; i < seg; ++i) mx = fmaxf(mx, gate[g * seg + i]);
    float s = 0.f;
    for (int i = 0; i < seg; ++i) { a[i] = expf(gate[g * seg + i] - mx); s += a[i]; }
    for (int i = 0; i < seg; ++i) { a[i] = a[i] / (s + 1e-16f); att[g * seg + i] = a[i]; }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float acc = 0.f;
    for (int i = 0; i < seg; ++i) acc = fmaf(a[i], x[((size_t)g * seg + i) * C + c], acc);
    pooled[(size_t)g * C + c] = acc;
  }
}

__global__ void gate_pool_bwd_kernel(const float* __restrict__ dpooled, const float* __restrict__ x,
                                     const float* __restrict__ att, float* __restrict__ dx, float* __restrict__ dgate,
                                     int seg, int C) {
  pdl_trigger();
  pdl_wait();
  __shared__ float sh[32];
  __shared__ float datt[64];
  const int g = blockIdx.x;
  for (int i = 0; i < seg; ++i) {
    float s = 0.f;
    for (int c = threadIdx.x; c < C; c += blockDim.x) s = fmaf(dpooled[(size_t)g * C + c], x[((size_t)g * seg + i) * C + c], s);
    s = block_sum(s, sh);
    if (threadIdx.x == 0) datt[i] = s;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float dot = 0.f;
    for (int i = 0; i < seg; ++i) dot = fmaf(att[g * seg + i], datt[i], dot);
    for (int i = 0; i < seg; ++i) dgate[g * seg + i] = att[g * seg + i] * (datt[i] - dot);
  }
  for (int i = 0; i < seg; ++i)
    for (int c = threadIdx.x; c < C; c += blockDim.x)
      dx[((size_t)g * seg + i) * C + c] = att[g * seg + i] * dpooled[(size_t)g * C + c];
}

// ---- multi-head attention over <= 8 tokens: one warp per (batch, head) -------------------------------
// qkv [B, N, 3, H, D] -> out [B, N, H*D]; probs [B, H, N, N] saved (after dropout scaling)
constexpr int kMaxTok = 8;
__device__ __forceinline__ float hash_uniform(uint64_t z) {
  z += 0x9e3779b97f4a7c15ull;
  z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
  z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
  return (float)((z ^ (z >> 31)) >> 40) * (1.f / 16777216.f);
}

__global__ void attn_small_fwd_kernel(const float* __restrict__ qkv, float* __restrict__ out, float* __restrict__ probs,
                                      int B, int N, int H, int D, float scale, float drop_p, uint64_t seed,
                                      const int* __restrict__ step_dev) {
  pdl_trigger();
  pdl_wait();
  const int wid = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (wid >= B * H) return;
  if (step_dev) seed += 0x9e3779b97f4a7c15ull * (uint64_t)(*step_dev);  // per-step stream under CUDA-graph replay
  const int b = wid / H, h = wid % H;
  const size_t tok = (size_t)3 * H * D;  // stride between tokens
  const float* base = qkv + (size_t)b * N * tok + (size_t)h * D;
  float p[kMaxTok][kMaxTok];
  for (int i = 0; i < N; ++i)
    for (int j = 0; j < N; ++j) {
      float s = 0.f;
      for (int d = lane; d < D; d += 32) s = fmaf(base[i * tok + d], base[j * tok + (size_t)H * D + d], s);
      p[i][j] = warp_sum(s) * scale;
    }
  for (int i = 0; i < N; ++i) {
    float mx = -INFINITY, s = 0.f;
    for (int j = 0; j < N; ++j) mx = fmaxf(mx, p[i][j]);
    for (int j = 0; j < N; ++j) { p[i][j] = expf(p[i][j] - mx); s += p[i][j]; }
    for (int j = 0; j < N; ++j) {
      float v = p[i][j] / s;
      if (drop_p > 0.f) {
        const float u = hash_uniform(seed + (((uint64_t)wid * kMaxTok + i) * kMaxTok + j));
        v = u >= drop_p ? v / (1.f - drop_p) : 0.f;
      }
      p[i][j] = v;
      if (lane == 0) probs[(((size_t)b * H + h) * N + i) * N + j] = v;
    }
  }
  for (int i = 0; i < N; ++i)
    for (int d = lane; d < D; d += 32) {
      float acc = 0.f;
      for (int j = 0; j < N; ++j) acc = fmaf(p[i][j], base[j * tok + (size_t)2 * H * D + d], acc);
      out[((size_t)b * N + i) * H * D + (size_t)h * D + d] = acc;
    }
}

__global__ void attn_small_bwd_kernel(const float* __restrict__ dout, const float* __restrict__ qkv,
                                      const float* __restrict__ probs, float* __restrict__ dqkv, int B, int N, int H, int D,
                                      float scale, float drop_p, uint64_t seed, const int* __restrict__ step_dev) {
  pdl_trigger();
  pdl_wait();
  if (step_dev) seed += 0x9e3779b97f4a7c15ull * (uint64_t)(*step_dev);
  const int wid = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (wid >= B * H) return;
  const int b = wid / H, h = wid % H;
  const size_t tok = (size_t)3 * H * D;
  const float* base = qkv + (size_t)b * N * tok + (size_t)h * D;
  float* dbase = dqkv + (size_t)b * N * tok + (size_t)h * D;
  float p[kMaxTok][kMaxTok], dp[kMaxTok][kMaxTok];
  for (int i = 0; i < N; ++i)
    for (int j = 0; j < N; ++j) p[i][j] = probs[(((size_t)b * H + h) * N + i) * N + j];
  // dP = dOut V^T ; dV = P^T dOut
  for (int i = 0; i < N; ++i)
    for (int j = 0; j < N; ++j) {
      float s = 0.f;
      for (int d = lane; d < D; d += 32)
        s = fmaf(dout[((size_t)b * N + i) * H * D + (size_t)h * D + d], base[j * tok + (size_t)2 * H * D + d], s);
      dp[i][j] = warp_sum(s);
    }
  for (int j = 0; j < N; ++j)
    for (int d = lane; d < D; d += 32) {
      float acc = 0.f;
      for (int i = 0; i < N; ++i) acc = fmaf(p[i][j], dout[((size_t)b * N + i) * H * D + (size_t)h * D + d], acc);
      dbase[j * tok + (size_t)2 * H * D + d] = acc;
    }
  // through dropout and softmax: p_drop = m * p_soft / keep ; dS = p_soft * (dP' - sum_j p_soft dP')
  for (int i = 0; i < N; ++i) {
    float ps[kMaxTok], dps[kMaxTok], dot = 0.f;
    for (int j = 0; j < N; ++j) {
      float keep_scale = 1.f, soft = p[i][j];
      if (drop_p > 0.f) {
        // recover the pre-dropout probability is impossible for dropped entries from p alone: recompute softmax
        keep_scale = hash_uniform(seed + (((uint64_t)wid * kMaxTok + i) * kMaxTok + j)) >= drop_p ? 1.f / (1.f - drop_p) : 0.f;
      }
      ps[j] = soft;
      dps[j] = dp[i][j] * keep_scale;
    }
    if (drop_p > 0.f) {  // recompute the un-dropped softmax row from q, k
      float sc[kMaxTok], mx = -INFINITY, s = 0.f;
      for (int j = 0; j < N; ++j) {
        float t = 0.f;
        for (int d = lane; d < D; d += 32) t = fmaf(base[i * tok + d], base[j * tok + (size_t)H * D + d], t);
        sc[j] = warp_sum(t) * scale;
        mx = fmaxf(mx, sc[j]);
      }
      for (int j = 0; j < N; ++j) { sc[j] = expf(sc[j] - mx); s += sc[j]; }
      for (int j = 0; j < N; ++j) ps[j] = sc[j] / s;
    }
    for (int j = 0; j < N; ++j) dot = fmaf(ps[j], dps[j], dot);
    for (int j = 0; j < N; ++j) dp[i][j] = ps[j] * (dps[j] - dot) * scale;  // dS (w.r.t. q.k)
  }
  for (int i = 0; i < N; ++i)
    for (int d = lane; d < D; d += 32) {
      float aq = 0.f, ak = 0.f;
      for (int j = 0; j < N; ++j) {
        aq = fmaf(dp[i][j], base[j * tok + (size_t)H * D + d], aq);  // dQ_i = sum_j dS_ij K_j
        ak = fmaf(dp[j][i], base[j * tok + d], ak);                  // dK_i = sum_j dS_ji Q_j
      }
      dbase[i * tok + d] = aq;
      dbase[i * tok + (size_t)H * D + d] = ak;
    }
}

// ---- row L2 normalisation (F.normalize(dim=1)) ---------------------------------------------------
__global__ void l2norm_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, float* __restrict__ norms, int C) {
  pdl_trigger();
  pdl_wait();
  __shared__ float sh[32];
  const int r = blockIdx.x;
  float s = 0.f;
  for (int c = threadIdx.x; c < C; c += blockDim.x) { const float v = x[(size_t)r * C + c]; s = fmaf(v, v, s); }
  const float nrm = fmaxf(sqrtf(block_sum(s, sh)), 1e-12f);
  if (threadIdx.x == 0) norms[r] = nrm;
  for (int c = threadIdx.x; c < C; c += blockDim.x) y[(size_t)r * C + c] = x[(size_t)r * C + c] / nrm;
}
__global__ void l2norm_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ y, const float* __restrict__ norms,
                                  float* __restrict__ dx, int C) {
  pdl_trigger();
  pdl_wait();
  __shared__ float sh[32];
  const int r = blockIdx.x;
  float s = 0.f;
  for (int c = threadIdx.x; c < C; c += blockDim.x) s = fmaf(dy[(size_t)r * C + c], y[(size_t)r * C + c], s);
  const float dot = block_sum(s, sh);
  const float inv = 1.f / norms[r];
  for (int c = threadIdx.x; c < C; c += blockDim.x)
    dx[(size_t)r * C + c] = (dy[(size_t)r * C + c] - y[(size_t)r * C + c] * dot) * inv;
}

// ---- row gather: y[i,:] = (idx[i] >= 0) ? x[idx[i],:] : fill[:]  --------------------------------------
__global__ void rows_gather_kernel(const float* __restrict__ x, const int* __restrict__ idx, const float* __restrict__ fill,
                                   float* __restrict__ y, int rows, int C) {
  pdl_trigger();
  pdl_wait();
  const int64_t total = (int64_t)rows * C;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int r = (int)(i / C), c = (int)(i % C);
    const int s = idx[r];
    y[i] = s >= 0 ? x[(size_t)s * C + c] : (fill ? fill[c] : 0.f);
  }
}
// dx[s,:] = sum_{i: idx[i] == s} dy[i,:] ; dfill[:] = sum_{i: idx[i] < 0} dy[i,:].  Gather form of the scatter-add: one
// block per source row (block src_rows = the fill row), rows visited in increasing order -> no atomics, fixed order.
__global__ void rows_scatter_add_kernel(const float* __restrict__ dy, const int* __restrict__ idx, float* __restrict__ dx,
                                        float* __restrict__ dfill, int rows, int src_rows, int C) {
  pdl_trigger();
  pdl_wait();
  const int s = blockIdx.x;
  const bool fill = s == src_rows;
  float* out = fill ? dfill : dx + (size_t)s * C;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float acc = 0.f;
    for (int i = 0; i < rows; ++i) {
      const int t = idx[i];
      if (fill ? t < 0 : t == s) acc += dy[(size_t)i * C + c];
    }
    out[c] = acc;
  }
}

// ---- losses of the classifier (my_train(full).py:309-347) -----------------------------------------------
// mean cross entropy of [B, K] logits; loss accumulated (atomic) scaled by `weight`, dlogits written scaled
// (one block; the loss terms of a launch are summed in a fixed order and added to *loss by one thread - launches on a
// stream are ordered, so the accumulated objective is bit-reproducible)
__global__ void softmax_ce_kernel(const float* __restrict__ logits, const int64_t* __restrict__ labels,
                                  float* __restrict__ loss, float* __restrict__ dlogits, int B, int K, float weight) {
  pdl_trigger();
  pdl_wait();
  __shared__ float sh[32];
  float part = 0.f;
  for (int r = threadIdx.x; r < B; r += blockDim.x) {
    float mx = -INFINITY, s = 0.f;
    for (int k = 0; k < K; ++k) mx = fmaxf(mx, logits[r * K + k]);
    for (int k = 0; k < K; ++k) s += expf(logits[r * K + k] - mx);
    const float lse = mx + logf(s);
    const int t = (int)labels[r];
    part += weight * (lse - logits[r * K + t]) / B;
    if (dlogits)
      for (int k = 0; k < K; ++k) dlogits[r * K + k] = weight * (expf(logits[r * K + k] - lse) - (k == t ? 1.f : 0.f)) / B;
  }
  part = block_sum(part, sh);
  if (threadIdx.x == 0) *loss += part;
}

// masked MSE between rows of a and b: loss += weight * mean over (selected rows x C) ; da written (0 on unselected rows)
__global__ void masked_mse_kernel(const float* __restrict__ a, const float* __restrict__ b, const uint8_t* __restrict__ sel,
                                  float* __restrict__ loss, float* __restrict__ da, float* __restrict__ db_, int rows, int C,
                                  float weight, float inv_count) {
  pdl_trigger();
  pdl_wait();
  __shared__ float sh[32];
  float s = 0.f;          // one block walks every row: a single fixed-order sum, no atomics
  for (int r = 0; r < rows; ++r) {
    const bool on = sel[r] != 0;
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
      const float d = on ? a[(size_t)r * C + c] - b[(size_t)r * C + c] : 0.f;
      s = fmaf(d, d, s);
      if (da) da[(size_t)r * C + c] = 2.f * d * weight * inv_count;
      if (db_) db_[(size_t)r * C + c] = -2.f * d * weight * inv_count;
    }
  }
  s = block_sum(s, sh);
  if (threadIdx.x == 0) *loss += weight * s * inv_count;
}

// =====================================================================================================================
// Segment-table operators: ALL modalities of the head in one launch.
// The node rows of every modality branch are stacked in one [R, C] matrix (modality-major, then patient, then node); a
// "segment" is one patient graph of one modality: rows seg_start[s] .. +seg_len[s], belonging to parameter set
// seg_set[s] (= the modality, whose own LayerNorm affine / weights apply).  The ORDER of the segment ids is free, so the
// pooled output can be produced patient-major (the token order of the masked auto-encoder) or modality-major (the row
// blocks the per-modality head GEMMs need) without a gather.  Parameter sets arrive as a small by-value pointer table.
constexpr int kMaxSets = CVX_MAX_PARAM_SETS;

struct SetPtrs {
  const float* w[kMaxSets];
  const float* b[kMaxSets];
  float* dw[kMaxSets];
  float* db[kMaxSets];
  int row_start[kMaxSets + 1];   // rows of set i: row_start[i] .. row_start[i+1] (stacked modality-major)
  int sets;
};

__global__ void segtab_layernorm_fwd_kernel(const float* __restrict__ x, SetPtrs P, float* __restrict__ y,
                                            float* __restrict__ stats, const int* __restrict__ seg_start,
                                            const int* __restrict__ seg_len, const int* __restrict__ seg_set, int C,
                                            float eps, int mode) {
  pdl_trigger();
  pdl_wait();
  __shared__ float sh[32];
  const int g = blockIdx.x;
  const int n = seg_len[g] * C;
  const size_t off = (size_t)seg_start[g] * C;
  const float* xs = x + off;
  const float* w = P.w[seg_set[g]];
  const float* b = P.b[seg_set[g]];
  float s = 0.f;
  for (int i = threadIdx.x; i < n; i += blockDim.x) s += xs[i];
  const float mean = block_sum(s, sh) / n;
  float v = 0.f;
  for (int i = threadIdx.x; i < n; i += blockDim.x) { const float d = xs[i] - mean; v = fmaf(d, d, v); }
  const float var = block_sum(v, sh) / n;
  const float sd = sqrtf(var);
  const float r = mode == 0 ? 1.f / (sd + eps) : rsqrtf(var + eps);
  if (threadIdx.x == 0) { stats[3 * g] = mean; stats[3 * g + 1] = r; stats[3 * g + 2] = sd; }
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const int c = i % C;
    y[off + i] = (xs[i] - mean) * r * w[c] + b[c];
  }
}

__global__ void segtab_layernorm_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ x, SetPtrs P,
                                            const float* __restrict__ stats, float* __restrict__ dx,
                                            const int* __restrict__ seg_start, const int* __restrict__ seg_len,
                                            const int* __restrict__ seg_set, int C, float eps, int mode) {
  pdl_trigger();
  pdl_wait();
  __shared__ float sh[32];
  const int g = blockIdx.x;
  const int n = seg_len[g] * C;
  const size_t off = (size_t)seg_start[g] * C;
  const float mean = stats[3 * g], r = stats[3 * g + 1], sd = stats[3 * g + 2];
  const float* xs = x + off;
  const float* gs = dy + off;
  const float* w = P.w[seg_set[g]];
  float sg = 0.f, sgx = 0.f;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const float gi = gs[i] * w[i % C];
    sg += gi;
    sgx = fmaf(gi, xs[i] - mean, sgx);
  }
  const float Sg = block_sum(sg, sh);
  const float Sgx = block_sum(sgx, sh);
  const float K = mode == 0 ? (sd > 0.f ? Sgx * r * r / (n * sd) : 0.f) : Sgx * r * r * r / n;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const float xc = xs[i] - mean;
    dx[off + i] = r * (gs[i] * w[i % C] - Sg / n) - xc * K;
  }
}

// affine gradients of every parameter set: block (x = 32 channels, y = set); fixed summation order
__global__ void __launch_bounds__(256) segtab_ln_param_grad_kernel(const float* __restrict__ dy, const float* __restrict__ x,
                                                                   const float* __restrict__ stats,
                                                                   const int* __restrict__ row_seg, SetPtrs P, int C) {
  pdl_trigger();
  pdl_wait();
  __shared__ float pw[8][33], pb[8][33];
  const int cx = threadIdx.x & 31, ry = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + cx, set = blockIdx.y;
  const int r0 = P.row_start[set], r1 = P.row_start[set + 1];
  float aw = 0.f, ab = 0.f;
  if (c < C) {
    for (int r = r0 + ry; r < r1; r += 8) {
      const int g = row_seg[r];
      const float gi = dy[(size_t)r * C + c];
      aw = fmaf(gi * (x[(size_t)r * C + c] - stats[3 * g]), stats[3 * g + 1], aw);
      ab += gi;
    }
  }
  pw[ry][cx] = aw;
  pb[ry][cx] = ab;
  __syncthreads();
  if (ry == 0 && c < C) {
    float sw = 0.f, sb = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) { sw += pw[k][cx]; sb += pb[k][cx]; }
    if (P.dw[set]) P.dw[set][c] = sw;
    if (P.db[set]) P.db[set][c] = sb;
  }
}

// gated attention pooling with a segment table: CTA s pools rows seg_start[s] .. +seg_len[s] into pooled row s
__global__ void segtab_gate_pool_fwd_kernel(const float* __restrict__ x, const float* __restrict__ gate,
                                            float* __restrict__ pooled, float* __restrict__ att,
                                            const int* __restrict__ seg_start, const int* __restrict__ seg_len, int C) {
  pdl_trigger();
  pdl_wait();
  __shared__ float a[64];
  const int s = blockIdx.x, r0 = seg_start[s], seg = seg_len[s];
  if (threadIdx.x == 0) {
    float mx = -INFINITY;
    for (int i = 0; i < seg; ++i) mx = fmaxf(mx, gate[r0 + i]);
    float sum = 0.f;
    for (int i = 0; i < seg; ++i) { a[i] = expf(gate[r0 + i] - mx); sum += a[i]; }
    for (int i = 0; i < seg; ++i) { a[i] = a[i] / (sum + 1e-16f); att[r0 + i] = a[i]; }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float acc = 0.f;
    for (int i = 0; i < seg; ++i) acc = fmaf(a[i], x[((size_t)r0 + i) * C + c], acc);
    pooled[(size_t)s * C + c] = acc;
  }
}

__global__ void segtab_gate_pool_bwd_kernel(const float* __restrict__ dpooled, const float* __restrict__ x,
                                            const float* __restrict__ att, float* __restrict__ dx,
                                            float* __restrict__ dgate, const int* __restrict__ seg_start,
                                            const int* __restrict__ seg_len, int C) {
  pdl_trigger();
  pdl_wait();
  __shared__ float sh[32];
  __shared__ float datt[64];
  const int s = blockIdx.x, r0 = seg_start[s], seg = seg_len[s];
  for (int i = 0; i < seg; ++i) {
    float v = 0.f;
    for (int c = threadIdx.x; c < C; c += blockDim.x) v = fmaf(dpooled[(size_t)s * C + c], x[((size_t)r0 + i) * C + c], v);
    v = block_sum(v, sh);
    if (threadIdx.x == 0) datt[i] = v;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float dot = 0.f;
    for (int i = 0; i < seg; ++i) dot = fmaf(att[r0 + i], datt[i], dot);
    for (int i = 0; i < seg; ++i) dgate[r0 + i] = att[r0 + i] * (datt[i] - dot);
  }
  for (int i = 0; i < seg; ++i)
    for (int c = threadIdx.x; c < C; c += blockDim.x)
      dx[((size_t)r0 + i) * C + c] = att[r0 + i] * dpooled[(size_t)s * C + c];
}

// y[row] = x[row] + t[tok_of_seg[row_seg[row]]]  (token -1: nothing added) - "node features += reconstructed modality
// token" (my_mae_model.py:636-649) for every modality at once
__global__ void segtab_bcast_add_kernel(const float* __restrict__ x, const float* __restrict__ t, float* __restrict__ y,
                                        const int* __restrict__ row_seg, const int* __restrict__ tok_of_seg, int rows, int C) {
  pdl_trigger();
  pdl_wait();
  const int64_t total = (int64_t)rows * C;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int r = (int)(i / C), c = (int)(i % C);
    const int tok = tok_of_seg[row_seg[r]];
    y[i] = x[i] + (tok >= 0 ? t[(size_t)tok * C + c] : 0.f);
  }
}
// its gradient w.r.t. the tokens: dt[tok] = sum of dy over the rows of the segment that took token tok (one CTA per token,
// rows in order: no atomics); seg_of_tok[tok] = that segment or -1 (then dt[tok] = 0)
__global__ void segtab_bcast_add_bwd_kernel(const float* __restrict__ dy, float* __restrict__ dt,
                                            const int* __restrict__ seg_of_tok, const int* __restrict__ seg_start,
                                            const int* __restrict__ seg_len, int C) {
  pdl_trigger();
  pdl_wait();
  const int tok = blockIdx.x, s = seg_of_tok[tok];
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float acc = 0.f;
    if (s >= 0)
      for (int i = 0; i < seg_len[s]; ++i) acc += dy[((size_t)seg_start[s] + i) * C + c];
    dt[(size_t)tok * C + c] = acc;
  }
}

static inline int rgrid(int64_t total) {
  int64_t b = ceil_div64(total, 256);
  const int64_t cap = (int64_t)kNumSMs * 8;
  return (int)(b > cap ? cap : (b < 1 ? 1 : b));
}

}  // namespace cvx

using namespace cvx;

extern "C" {

int cvx_seg_layernorm_fwd(const float* x, const float* w, const float* b, float* y, float* stats, int groups, int seg, int c,
                          float eps, int mode, void* stream) {
  CVX_CHECK_ARG(x && w && b && y && stats && groups > 0 && seg > 0 && c > 0 && (mode == 0 || mode == 1), "seg_layernorm_fwd: bad arguments");
  launch_pdl(seg_layernorm_fwd_kernel, dim3(groups), dim3(256), 0, as_stream(stream), x, w, b, y, stats, seg, c, eps, mode);
  CVX_LAUNCH_OK();
  return CVX_OK;
}

int cvx_seg_layernorm_bwd(const float* dy, const float* x, const float* w, const float* stats, float* dx, float* dw,
                          float* db, int groups, int seg, int c, float eps, int mode, void* stream) {
  CVX_CHECK_ARG(dy && x && w && stats && dx && groups > 0 && seg > 0 && c > 0, "seg_layernorm_bwd: bad arguments");
  cudaStream_t st = as_stream(stream);
  launch_pdl(seg_layernorm_bwd_kernel, dim3(groups), dim3(256), 0, st, dy, x, w, stats, dx, seg, c, eps, mode);
  CVX_LAUNCH_OK();
  if (dw || db) {
    launch_pdl(ln_param_grad_kernel, dim3((c + 31) / 32), dim3(256), 0, st, dy, x, stats, dw, db, groups * seg, seg, c);
    CVX_LAUNCH_OK();
  }
  return CVX_OK;
}

static int fill_sets(SetPtrs* P, const cvx_param_sets* ps, const char* who) {
  CVX_CHECK_ARG(ps && ps->sets > 0 && ps->sets <= kMaxSets, "%s: 1..%d parameter sets", who, kMaxSets);
  P->sets = ps->sets;
  for (int i = 0; i < kMaxSets; ++i) {
    const bool on = i < ps->sets;
    P->w[i] = on ? ps->w[i] : nullptr; P->b[i] = on ? ps->b[i] : nullptr;
    P->dw[i] = on ? ps->dw[i] : nullptr; P->db[i] = on ? ps->db[i] : nullptr;
    P->row_start[i] = on ? ps->row_start[i] : 0;
  }
  for (int i = ps->sets; i <= kMaxSets; ++i) P->row_start[i] = ps->row_start[ps->sets];
  return CVX_OK;
}

int cvx_segtab_layernorm_fwd(const float* x, const cvx_param_sets* ps, float* y, float* stats, const int* seg_start,
                             const int* seg_len, const int* seg_set, int segments, int c, float eps, int mode, void* stream) {
  CVX_CHECK_ARG(x && y && stats && seg_start && seg_len && seg_set && segments > 0 && c > 0 && (mode == 0 || mode == 1),
                "segtab_layernorm_fwd: bad arguments");
  SetPtrs P;
  if (int rc = fill_sets(&P, ps, "segtab_layernorm_fwd")) return rc;
  for (int i = 0; i < P.sets; ++i) CVX_CHECK_ARG(P.w[i] && P.b[i], "segtab_layernorm_fwd: set %d has no affine parameters", i);
  launch_pdl(segtab_layernorm_fwd_kernel, dim3(segments), dim3(256), 0, as_stream(stream), x, P, y, stats, seg_start, seg_len, seg_set, c, eps, mode);
  CVX_LAUNCH_OK();
  return CVX_OK;
}

int cvx_segtab_layernorm_bwd(const float* dy, const float* x, const cvx_param_sets* ps, const float* stats, float* dx,
                             const int* seg_start, const int* seg_len, const int* seg_set, const int* row_seg, int segments,
                             int c, float eps, int mode, void* stream) {
  CVX_CHECK_ARG(dy && x && stats && dx && seg_start && seg_len && seg_set && row_seg && segments > 0 && c > 0,
                "segtab_layernorm_bwd: bad arguments");
  SetPtrs P;
  if (int rc = fill_sets(&P, ps, "segtab_layernorm_bwd")) return rc;
  cudaStream_t st = as_stream(stream);
  launch_pdl(segtab_layernorm_bwd_kernel, dim3(segments), dim3(256), 0, st, dy, x, P, stats, dx, seg_start, seg_len, seg_set, c, eps, mode);
  CVX_LAUNCH_OK();
  launch_pdl(segtab_ln_param_grad_kernel, dim3(dim3((c + 31) / 32, P.sets)), dim3(256), 0, st, dy, x, stats, row_seg, P, c);
  CVX_LAUNCH_OK();
  return CVX_OK;
}

int cvx_segtab_gate_pool_fwd(const float* x, const float* gate, float* pooled, float* att, const int* seg_start,
                             const int* seg_len, int segments, int max_len, int c, void* stream) {
  CVX_CHECK_ARG(x && gate && pooled && att && seg_start && seg_len && segments > 0 && max_len > 0 && max_len <= 64 && c > 0,
                "segtab_gate_pool_fwd: bad arguments (segments of at most 64 rows)");
  launch_pdl(segtab_gate_pool_fwd_kernel, dim3(segments), dim3(256), 0, as_stream(stream), x, gate, pooled, att, seg_start, seg_len, c);
  CVX_LAUNCH_OK();
  return CVX_OK;
}

int cvx_segtab_gate_pool_bwd(const float* dpooled, const float* x, const float* att, float* dx, float* dgate,
                             const int* seg_start, const int* seg_len, int segments, int max_len, int c, void* stream) {
  CVX_CHECK_ARG(dpooled && x && att && dx && dgate && seg_start && seg_len && segments > 0 && max_len > 0 && max_len <= 64 && c > 0,
                "segtab_gate_pool_bwd: bad arguments");
  launch_pdl(segtab_gate_pool_bwd_kernel, dim3(segments), dim3(256), 0, as_stream(stream), dpooled, x, att, dx, dgate, seg_start, seg_len, c);
  CVX_LAUNCH_OK();
  return CVX_OK;
}

int cvx_segtab_bcast_add(const float* x, const float* t, float* y, const int* row_seg, const int* tok_of_seg, int rows, int c,
                         void* stream) {
  CVX_CHECK_ARG(x && t && y && row_seg && tok_of_seg && rows > 0 && c > 0, "segtab_bcast_add: bad arguments");
  launch_pdl(segtab_bcast_add_kernel, dim3(rgrid((int64_t)rows * c)), dim3(256), 0, as_stream(stream), x, t, y, row_seg, tok_of_seg, rows, c);
  CVX_LAUNCH_OK();
  return CVX_OK;
}

int cvx_segtab_bcast_add_bwd(const float* dy, float* dt, const int* seg_of_tok, const int* seg_start, const int* seg_len,
                             int tokens, int c, void* stream) {
  CVX_CHECK_ARG(dy && dt && seg_of_tok && seg_start && seg_len && tokens > 0 && c > 0, "segtab_bcast_add_bwd: bad arguments");
  launch_pdl(segtab_bcast_add_bwd_kernel, dim3(tokens), dim3(128), 0, as_stream(stream), dy, dt, seg_of_tok, seg_start, seg_len, c);
  CVX_LAUNCH_OK();
  return CVX_OK;
}

int cvx_gelu_fwd(const float* x, float* y, int64_t n, void* stream) {
  CVX_CHECK_ARG(x && y && n > 0, "gelu_fwd: bad arguments");
  launch_pdl(gelu_fwd_kernel, dim3(rgrid(n)), dim3(256), 0, as_stream(stream), x, y, n);
  CVX_LAUNCH_OK();
  return CVX_OK;
}

int cvx_gelu_bwd(const float* dy, const float* x, float* dx, int64_t n, void* stream) {
  CVX_CHECK_ARG(dy && x && dx && n > 0, "gelu_bwd: bad arguments");
  launch_pdl(gelu_bwd_kernel, dim3(rgrid(n)), dim3(256), 0, as_stream(stream), dy, x, dx, n);
  CVX_LAUNCH_OK();
  return CVX_OK;
}

int cvx_graph_gather(const float* x, float* out, int groups, int nodes, int c, const int* rowptr, const int* col,
                     const float* w, void* stream) {
  CVX_CHECK_ARG(x && out && rowptr && col && w && groups > 0 && nodes > 0 && c > 0, "graph_gather: bad arguments");
  launch_pdl(graph_gather_kernel, dim3(rgrid((int64_t)groups * nodes * c)), dim3(256), 0, as_stream(stream), x, out, groups, nodes, c, rowptr, col, w);
  CVX_LAUNCH_OK();
  return CVX_OK;
}

int cvx_gate_pool_fwd(const float* x, const float* gate, float* pooled, float* att, int groups, int seg, int c, void* stream) {
  CVX_CHECK_ARG(x && gate && pooled && att && groups > 0 && seg > 0 && seg <= 64 && c > 0, "gate_pool_fwd: bad arguments");
  launch_pdl(gate_pool_fwd_kernel, dim3(groups), dim3(256), 0, as_stream(stream), x, gate, pooled, att, seg, c);
  CVX_LAUNCH_OK();
  return CVX_OK;
}

int cvx_gate_pool_bwd(const float* dpooled, const float* x, const float* att, float* dx, float* dgate, int groups, int seg,
                      int c, void* stream) {
  CVX_CHECK_ARG(dpooled && x && att && dx && dgate && groups > 0 && seg > 0 && seg <= 64 && c > 0, "gate_pool_bwd: bad arguments");
  launch_pdl(gate_pool_bwd_kernel, dim3(groups), dim3(256), 0, as_stream(stream), dpooled, x, att, dx, dgate, seg, c);
  CVX_LAUNCH_OK();
  return CVX_OK;
}

int cvx_attn_small_fwd(const float* qkv, float* out, float* probs, int b, int n, int h, int d, float scale, float drop_p,
                       uint64_t seed, const int* step_dev, void* stream) {
  CVX_CHECK_ARG(qkv && out && probs && b > 0 && n > 0 && n <= kMaxTok && h > 0 && d > 0 && drop_p >= 0.f && drop_p < 1.f,
                "attn_small_fwd: bad arguments (at most %d tokens)", kMaxTok);
  const int warps = b * h;
  launch_pdl(attn_small_fwd_kernel, dim3((warps * 32 + 127) / 128), dim3(128), 0, as_stream(stream), qkv, out, probs, b, n, h, d, scale, drop_p, seed, step_dev);
  CVX_LAUNCH_OK();
  return CVX_OK;
}

int cvx_attn_small_bwd(const float* dout, const float* qkv, const float* probs, float* dqkv, int b, int n, int h, int d,
                       float scale, float drop_p, uint64_t seed, const int* step_dev, void* stream) {
  CVX_CHECK_ARG(dout && qkv && probs && dqkv && b > 0 && n > 0 && n <= kMaxTok && h > 0 && d > 0, "attn_small_bwd: bad arguments");
  const int warps = b * h;
  launch_pdl(attn_small_bwd_kernel, dim3((warps * 32 + 127) / 128), dim3(128), 0, as_stream(stream), dout, qkv, probs, dqkv, b, n, h, d, scale, drop_p, seed, step_dev);
  CVX_LAUNCH_OK();
  return CVX_OK;
}

int cvx_l2norm_fwd(const float* x, float* y, float* norms, int rows, int c, void* stream) {
  CVX_CHECK_ARG(x && y && norms && rows > 0 && c > 0, "l2norm_fwd: bad arguments");
  launch_pdl(l2norm_fwd_kernel, dim3(rows), dim3(128), 0, as_stream(stream), x, y, norms, c);
  CVX_LAUNCH_OK();
  return CVX_OK;
}

int cvx_l2norm_bwd(const float* dy, const float* y, const float* norms, float* dx, int rows, int c, void* stream) {
  CVX_CHECK_ARG(dy && y && norms && dx && rows > 0 && c > 0, "l2norm_bwd: bad arguments");
  launch_pdl(l2norm_bwd_kernel, dim3(rows), dim3(128), 0, as_stream(stream), dy, y, norms, dx, c);
  CVX_LAUNCH_OK();
  return CVX_OK;
}

int cvx_rows_gather(const float* x, const int* idx, const float* fill, float* y, int rows, int c, void* stream) {
  CVX_CHECK_ARG(x && idx && y && rows > 0 && c > 0, "rows_gather: bad arguments");
  launch_pdl(rows_gather_kernel, dim3(rgrid((int64_t)rows * c)), dim3(256), 0, as_stream(stream), x, idx, fill, y, rows, c);
  CVX_LAUNCH_OK();
  return CVX_OK;
}

int cvx_rows_scatter_add(const float* dy, const int* idx, float* dx, float* dfill, int rows, int src_rows, int c,
                         void* stream) {
  CVX_CHECK_ARG(dy && idx && dx && rows > 0 && src_rows > 0 && c > 0, "rows_scatter_add: bad arguments");
  launch_pdl(rows_scatter_add_kernel, dim3(src_rows + (dfill ? 1 : 0)), dim3(128), 0, as_stream(stream), dy, idx, dx, dfill, rows, src_rows, c);
  CVX_LAUNCH_OK();
  return CVX_OK;
}

int cvx_softmax_ce(const float* logits, const int64_t* labels, float* loss, float* dlogits, int b, int k, float weight,
                   void* stream) {
  CVX_CHECK_ARG(logits && labels && loss && b > 0 && k > 0, "softmax_ce: bad arguments");
  launch_pdl(softmax_ce_kernel, dim3(1), dim3(128), 0, as_stream(stream), logits, labels, loss, dlogits, b, k, weight);
  CVX_LAUNCH_OK();
  return CVX_OK;
}

int cvx_masked_mse(const float* a, const float* b, const uint8_t* sel, float* loss, float* da, float* db, int rows, int c,
                   float weight, float inv_count, void* stream) {
  CVX_CHECK_ARG(a && b && sel && loss && rows > 0 && c > 0, "masked_mse: bad arguments");
  launch_pdl(masked_mse_kernel, dim3(1), dim3(512), 0, as_stream(stream), a, b, sel, loss, da, db, rows, c, weight, inv_count);
  CVX_LAUNCH_OK();
  return CVX_OK;
}

}  // extern "C"
