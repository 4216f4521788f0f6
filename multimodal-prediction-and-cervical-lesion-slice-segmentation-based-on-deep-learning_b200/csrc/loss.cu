// Fused segmentation objective: class-weighted CE, focal, dice and f_score statistics in ONE
// pass over the logits, and a single gradient pass for any weighted sum of the three losses.
//
// Reference (each a separate chain of ATen kernels over the 5x512x512 logits):
//   nets/deeplabv3_training.py:9-19  CE_Loss     nets/deeplabv3_training.py:21-36 Focal_Loss
//   nets/deeplabv3_training.py:38-56 Dice_loss   utils/utils_metrics.py:13-35     f_score
// Bandwidth-bound: logits (4*C B) + target (8 B) [+ one-hot 4*(C+1) B] read per pixel; the
// gradient pass reads the same and writes 4*C B.
#include "common.cuh"

namespace cvx {

constexpr int kMaxC = 16;

struct PixelSoftmax {
  float p[kMaxC];
  float logp_t;  // log-softmax at the target class (0 when ignored)
};

// CT > 0: class count known at compile time (CT = 5 is the reference's head): the per-class loops unroll to exactly
// CT iterations with no predicates; CT = 0 keeps the generic <= kMaxC form.
template <int CT>
__device__ __forceinline__ void softmax_px(const float* __restrict__ logits, int64_t base, int64_t cstride, int C,
                                           float* p, float* lse) {
  constexpr int NC = CT > 0 ? CT : kMaxC;
  float mx = -INFINITY;
#pragma unroll
  for (int c = 0; c < NC; ++c)
    if (c < C) { p[c] = __ldg(logits + base + c * cstride); mx = fmaxf(mx, p[c]); }
  float s = 0.f;
#pragma unroll
  for (int c = 0; c < NC; ++c)
    if (c < C) { p[c] = __expf(p[c] - mx); s += p[c]; }
  const float inv = 1.f / s;
  *lse = mx + __logf(s);
#pragma unroll
  for (int c = 0; c < NC; ++c)
    if (c < C) p[c] *= inv;
}

__device__ __forceinline__ float pow_gamma(float base, float gamma) {
  if (gamma == 2.f) return base * base;
  if (gamma == 1.f) return base;
  if (gamma == 0.f) return 1.f;
  return base > 0.f ? powf(base, gamma) : 0.f;
}

// stats layout: [0] sum w*nll [1] sum w [2] sum focal [3] pixels | tp[C] sp[C] st[C] | tph[C] sph[C] st[C]
template <int CT>
__global__ void __launch_bounds__(256) seg_loss_stats_kernel(const float* __restrict__ logits,
                                                             const int64_t* __restrict__ target,
                                                             const float* __restrict__ onehot,
                                                             const float* __restrict__ cls_w,
                                                             double* __restrict__ stats, int64_t npix, int64_t hw,
                                                             int C_rt, float alpha, float gamma, float thr) {
  constexpr int NC = CT > 0 ? CT : kMaxC;
  const int C = CT > 0 ? CT : C_rt;
  __shared__ float sm[4 + 6 * kMaxC];
  const int nstat = 4 + 6 * C;
  for (int i = threadIdx.x; i < nstat; i += blockDim.x) sm[i] = 0.f;
  __syncthreads();

  float a_ce = 0.f, a_w = 0.f, a_focal = 0.f, a_cnt = 0.f;
  float tp[NC], sp[NC], st[NC], tph[NC], sph[NC];
#pragma unroll
  for (int c = 0; c < NC; ++c) tp[c] = sp[c] = st[c] = tph[c] = sph[c] = 0.f;

  // grid = (pixel blocks, images): no 64-bit division per pixel
  const int64_t nn = blockIdx.y;
  for (int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; r < hw; r += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i = nn * hw + r;
    float p[NC], lse;
    softmax_px<CT>(logits, nn * C * hw + r, hw, C, p, &lse);
    const int t = (int)target[i];
    const bool valid = t >= 0 && t < C;
    a_cnt += 1.f;
    if (valid) {
      const float w = cls_w ? cls_w[t] : 1.f;
      const float logp_t = __ldg(logits + nn * C * hw + (int64_t)t * hw + r) - lse;
      const float u = w * logp_t;  // class-weighted log-prob ("logpt" of Focal_Loss)
      a_ce -= u;
      a_w += w;
      const float pt = __expf(u);
      a_focal -= pow_gamma(1.f - pt, gamma) * alpha * u;
    }
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      if (c < C) {
        const float tc = onehot ? onehot[i * (C + 1) + c] : (t == c ? 1.f : 0.f);
        const float hard = p[c] > thr ? 1.f : 0.f;
        tp[c] = fmaf(tc, p[c], tp[c]);
        sp[c] += p[c];
        st[c] += tc;
        tph[c] = fmaf(tc, hard, tph[c]);
        sph[c] += hard;
      }
    }
  }
  // warp -> block -> global
  a_ce = warp_sum(a_ce); a_w = warp_sum(a_w); a_focal = warp_sum(a_focal); a_cnt = warp_sum(a_cnt);
  // the eight warp leaders add in turn (fixed order, no shared atomics): a block's partial is bit-reproducible and the
  // fp64 global sum of the block partials is exact for these magnitudes
  const bool lead = (threadIdx.x & 31) == 0;
  float wv[5 * NC];
#pragma unroll
  for (int c = 0; c < NC; ++c) {
    wv[5 * c] = warp_sum(tp[c]); wv[5 * c + 1] = warp_sum(sp[c]); wv[5 * c + 2] = warp_sum(st[c]);
    wv[5 * c + 3] = warp_sum(tph[c]); wv[5 * c + 4] = warp_sum(sph[c]);
  }
  for (int turn = 0; turn < (int)(blockDim.x >> 5); ++turn) {
    if (lead && (int)(threadIdx.x >> 5) == turn) {
      sm[0] += a_ce; sm[1] += a_w; sm[2] += a_focal; sm[3] += a_cnt;
#pragma unroll
      for (int c = 0; c < NC; ++c) {
        if (c < C) {
          sm[4 + c] += wv[5 * c]; sm[4 + C + c] += wv[5 * c + 1]; sm[4 + 2 * C + c] += wv[5 * c + 2];
          sm[4 + 3 * C + c] += wv[5 * c + 3]; sm[4 + 4 * C + c] += wv[5 * c + 4]; sm[4 + 5 * C + c] += wv[5 * c + 2];
        }
      }
    }
    __syncthreads();
  }
  __syncthreads();
  for (int i = threadIdx.x; i < nstat; i += blockDim.x) atomicAdd(stats + i, (double)sm[i]);
}

__device__ __forceinline__ double dice_score(const double* tp, const double* sp, const double* st, int C, double beta,
                                             double smooth) {
  double acc = 0.0;
  const double b2 = beta * beta;
  for (int c = 0; c < C; ++c) {
    const double fp = sp[c] - tp[c], fn = st[c] - tp[c];
    acc += ((1 + b2) * tp[c] + smooth) / ((1 + b2) * tp[c] + b2 * fn + fp + smooth);
  }
  return acc / C;
}

__global__ void seg_loss_finalize_kernel(const double* __restrict__ stats, float* __restrict__ results, int C,
                                         float beta, float smooth) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  results[0] = (float)(stats[0] / stats[1]);
  results[1] = (float)(stats[2] / stats[3]);
  results[2] = (float)(1.0 - dice_score(stats + 4, stats + 4 + C, stats + 4 + 2 * C, C, beta, smooth));
  results[3] = (float)dice_score(stats + 4 + 3 * C, stats + 4 + 4 * C, stats + 4 + 5 * C, C, beta, smooth);
}

template <int CT>
__global__ void __launch_bounds__(256) seg_loss_grad_kernel(const float* __restrict__ logits,
                                                            const int64_t* __restrict__ target,
                                                            const float* __restrict__ onehot,
                                                            const float* __restrict__ cls_w,
                                                            const double* __restrict__ stats,
                                                            const float* __restrict__ gup, float* __restrict__ dlogits,
                                                            int64_t npix, int64_t hw, int C_rt, float alpha, float gamma,
                                                            float beta, float smooth) {
  constexpr int NC = CT > 0 ? CT : kMaxC;
  const int C = CT > 0 ? CT : C_rt;
  const float g_ce = gup[0], g_focal = gup[1], g_dice = gup[2];
  const float inv_wsum = (float)(1.0 / stats[1]);
  const float inv_npix = (float)(1.0 / stats[3]);
  const float b2 = beta * beta;
  float dA[NC], dB[NC];  // dice: dL/dp_c = -(dA[c]*t_c - dB[c])
#pragma unroll
  for (int c = 0; c < NC; ++c) {
    if (c < C) {
      const double tp = stats[4 + c], sp = stats[4 + C + c], st = stats[4 + 2 * C + c];
      const double D = b2 * st + sp + smooth;  // (1+b2)tp + b2*fn + fp + smooth with fn=st-tp, fp=sp-tp
      const double N = (1 + b2) * tp + smooth;
      dA[c] = (float)((1 + b2) / D / C);
      dB[c] = (float)(N / (D * D) / C);
    }
  }
  const int64_t nn = blockIdx.y;
  for (int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; r < hw; r += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i = nn * hw + r;
    const int64_t base = nn * C * hw + r;
    float p[NC], lse;
    softmax_px<CT>(logits, base, hw, C, p, &lse);
    const int t = (int)target[i];
    const bool valid = t >= 0 && t < C;
    float k_onehot = 0.f;  // coefficient of (1[k==t] - p_k)
    if (valid) {
      const float w = cls_w ? cls_w[t] : 1.f;
      const float logp_t = __ldg(logits + base + (int64_t)t * hw) - lse;
      const float u = w * logp_t;
      const float pt = __expf(u);
      const float om = 1.f - pt;
      const float dfdu = -alpha * (pow_gamma(om, gamma) - gamma * pow_gamma(om, gamma - 1.f) * pt * u);
      k_onehot = -g_ce * w * inv_wsum + g_focal * inv_npix * dfdu * w;
    }
    float G[NC], dot = 0.f;
    if (g_dice != 0.f) {
#pragma unroll
      for (int c = 0; c < NC; ++c) {
        if (c < C) {
          const float tc = onehot ? onehot[i * (C + 1) + c] : (t == c ? 1.f : 0.f);
          G[c] = -g_dice * (dA[c] * tc - dB[c]);
          dot = fmaf(p[c], G[c], dot);
        }
      }
    }
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      if (c < C) {
        float d = k_onehot * ((c == t ? 1.f : 0.f) - p[c]);
        if (g_dice != 0.f) d += p[c] * (G[c] - dot);
        dlogits[base + c * hw] = d;
      }
    }
  }
}

}  // namespace cvx

using namespace cvx;

extern "C" {

int cvx_seg_loss_stats(const float* logits, const int64_t* target, const float* onehot, const float* cls_weights,
                       double* stats, int n, int c, int h, int w, float focal_alpha, float focal_gamma,
                       float threshold, void* stream) {
  CVX_CHECK_ARG(logits && target && stats && n > 0 && c > 0 && c <= kMaxC && h > 0 && w > 0,
                "seg_loss_stats: bad arguments (C must be <= %d)", kMaxC);
  cudaStream_t st = as_stream(stream);
  CVX_WS_ZERO(stats, sizeof(double) * (4 + 6 * c), st);
  const int64_t npix = (int64_t)n * h * w;
  const int64_t hw = (int64_t)h * w;
  CVX_CHECK_ARG(n <= 65535, "seg_loss_stats: batch %d exceeds the grid's image dimension", n);
  int64_t bx = ceil_div64(hw, 256 * 4), cap = ceil_div64((int64_t)kNumSMs * 8, n);
  if (bx > cap) bx = cap;
  const dim3 grid((unsigned)bx, (unsigned)n);
  if (c == 5)
    seg_loss_stats_kernel<5><<<grid, 256, 0, st>>>(logits, target, onehot, cls_weights, stats, npix, hw, c, focal_alpha,
                                                   focal_gamma, threshold);
  else
    seg_loss_stats_kernel<0><<<grid, 256, 0, st>>>(logits, target, onehot, cls_weights, stats, npix, hw, c, focal_alpha,
                                                   focal_gamma, threshold);
  CVX_LAUNCH_OK();
  return CVX_OK;
}

int cvx_seg_loss_finalize(const double* stats, float* results, int c, float beta, float smooth, void* stream) {
  CVX_CHECK_ARG(stats && results && c > 0 && c <= kMaxC, "seg_loss_finalize: bad arguments");
  seg_loss_finalize_kernel<<<1, 32, 0, as_stream(stream)>>>(stats, results, c, beta, smooth);
  CVX_LAUNCH_OK();
  return CVX_OK;
}

int cvx_seg_loss_grad(const float* logits, const int64_t* target, const float* onehot, const float* cls_weights,
                      const double* stats, const float* g, float* dlogits, int n, int c, int h, int w,
                      float focal_alpha, float focal_gamma, float beta, float smooth, void* stream) {
  CVX_CHECK_ARG(logits && target && stats && g && dlogits && n > 0 && c > 0 && c <= kMaxC && h > 0 && w > 0,
                "seg_loss_grad: bad arguments");
  const int64_t npix = (int64_t)n * h * w;
  const int64_t hw = (int64_t)h * w;
  CVX_CHECK_ARG(n <= 65535, "seg_loss_grad: batch %d exceeds the grid's image dimension", n);
  int64_t bx = ceil_div64(hw, 256 * 2), cap = ceil_div64((int64_t)kNumSMs * 16, n);
  if (bx > cap) bx = cap;
  const dim3 grid((unsigned)bx, (unsigned)n);
  if (c == 5)
    seg_loss_grad_kernel<5><<<grid, 256, 0, as_stream(stream)>>>(logits, target, onehot, cls_weights, stats, g, dlogits,
                                                                 npix, hw, c, focal_alpha, focal_gamma, beta, smooth);
  else
    seg_loss_grad_kernel<0><<<grid, 256, 0, as_stream(stream)>>>(logits, target, onehot, cls_weights, stats, g, dlogits,
                                                                 npix, hw, c, focal_alpha, focal_gamma, beta, smooth);
  CVX_LAUNCH_OK();
  return CVX_OK;
}

}  // extern "C"
