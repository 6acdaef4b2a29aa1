// Fused segmentation loss / metric kernels and the fused clip + Adam step.
// Reference behaviour: deadtrees/loss/losses.py:124-141 (class2one_hot), :226-247 (DiceLoss),
// :273-291 (FocalLoss); deadtrees/loss/gdl.py:10-27 (GeneralizedDiceLoss); smp Fscore as used at
// deadtrees/network/segmodel.py:145-149,202-208; calculate_loss at segmodel.py:169-200; Adam at :420-425.
#include "common.cuh"

namespace {

constexpr int kThreads = 256;
constexpr int KMAX = 4;
constexpr float kEps = 1e-10f;

template <int K>
__device__ __forceinline__ void softmax_px(const float* __restrict__ z, int64_t plane, float (&p)[K]) {
  float m = z[0];
#pragma unroll
  for (int k = 1; k < K; ++k) m = fmaxf(m, z[k * plane]);
  float s = 0.f;
#pragma unroll
  for (int k = 0; k < K; ++k) { p[k] = expf(z[k * plane] - m); s += p[k]; }
#pragma unroll
  for (int k = 0; k < K; ++k) p[k] = p[k] / s;
}

// grid = (chunks, N).  sums[n][k][4], counts[k][3].
template <int K>
__global__ void loss_partials_kernel(const float* __restrict__ logits, const int64_t* __restrict__ labels, int64_t HW,
                                     double* __restrict__ sums, unsigned long long* __restrict__ counts,
                                     int* __restrict__ bad_label) {
  const int n = blockIdx.y;
  const float* zl = logits + static_cast<int64_t>(n) * K * HW;
  const int64_t* ll = labels + static_cast<int64_t>(n) * HW;
  float s_pt[K], s_p[K], s_t[K], s_f[K];
  unsigned c_tp[K], c_pr[K], c_gt[K];
#pragma unroll
  for (int k = 0; k < K; ++k) { s_pt[k] = s_p[k] = s_t[k] = s_f[k] = 0.f; c_tp[k] = c_pr[k] = c_gt[k] = 0u; }
  bool bad = false;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < HW;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    float p[K];
    softmax_px<K>(zl + i, HW, p);
    const int64_t lab = ll[i];
    if (lab < 0 || lab >= K) bad = true;
#pragma unroll
    for (int k = 0; k < K; ++k) {
      const bool t = (lab == k);
      const bool pr = p[k] > 0.5f;
      s_p[k] += p[k];
      if (t) {
        s_pt[k] += p[k];
        s_t[k] += 1.f;
        const float q = 1.f - p[k];
        s_f[k] += q * q * logf(p[k] + kEps);
      }
      c_pr[k] += pr;
      c_gt[k] += t;
      c_tp[k] += (pr && t);
    }
  }
  if (bad) atomicOr(bad_label, 1);
  __shared__ double sh[kThreads / 32][K][4];
  __shared__ unsigned shc[kThreads / 32][K][3];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int k = 0; k < K; ++k) {
    const double a = warp_sum(static_cast<double>(s_pt[k])), b = warp_sum(static_cast<double>(s_p[k]));
    const double c = warp_sum(static_cast<double>(s_t[k])), d = warp_sum(static_cast<double>(s_f[k]));
    const unsigned e = warp_sum(c_tp[k]), f = warp_sum(c_pr[k]), g = warp_sum(c_gt[k]);
    if (lane == 0) {
      sh[warp][k][0] = a; sh[warp][k][1] = b; sh[warp][k][2] = c; sh[warp][k][3] = d;
      shc[warp][k][0] = e; shc[warp][k][1] = f; shc[warp][k][2] = g;
    }
  }
  __syncthreads();
  if (threadIdx.x < K * 4) {
    const int k = threadIdx.x / 4, j = threadIdx.x % 4;
    double t = 0;
    for (int w = 0; w < kThreads / 32; ++w) t += sh[w][k][j];
    atomicAdd(&sums[(static_cast<int64_t>(n) * K + k) * 4 + j], t);
  } else if (threadIdx.x >= 32 && threadIdx.x < 32 + K * 3) {
    const int k = (threadIdx.x - 32) / 3, j = (threadIdx.x - 32) % 3;
    unsigned long long t = 0;
    for (int w = 0; w < kThreads / 32; ++w) t += shc[w][k][j];
    atomicAdd(&counts[k * 3 + j], t);
  }
}

// single block; thread-per-(n,k) work is tiny
// (round 2: one thread walking N * K * 4 doubles in global memory took 70 us of every training step; now the block stages
//  the partial sums in shared memory and zeroes `coef` in parallel, thread 0 keeps the serial arithmetic order - same bits)
__global__ void loss_finalize_kernel(const double* __restrict__ gsums, const long long* __restrict__ counts, int N,
                                     int K, int dice_mode, int use_focal, float* __restrict__ out,
                                     float* __restrict__ coef, float* __restrict__ focal_scale, int staged) {
  extern __shared__ double ssums[];
  const double* sums = gsums;
  if (staged) {
    for (int i = threadIdx.x; i < N * K * 4; i += blockDim.x) ssums[i] = gsums[i];
    sums = ssums;
  }
  for (int i = threadIdx.x; i < N * K * 2; i += blockDim.x) coef[i] = 0.f;
  __syncthreads();
  if (threadIdx.x != 0) return;
  float dice = 0.f, focal = 0.f;
  if (dice_mode == 1) {
    // DiceLoss over foreground classes: mean_{b,c}(1 - (2I + eps) / (U + eps))
    const int cnt = N * (K - 1);
    float acc = 0.f;
    for (int n = 0; n < N; ++n)
      for (int k = 1; k < K; ++k) {
        const double* s = sums + (static_cast<int64_t>(n) * K + k) * 4;
        const float I = static_cast<float>(s[0]), U = static_cast<float>(s[1]) + static_cast<float>(s[2]);
        acc += 1.f - (2.f * I + kEps) / (U + kEps);
        coef[(n * K + k) * 2 + 0] = -2.f / (U + kEps) / cnt;
        coef[(n * K + k) * 2 + 1] = (2.f * I + kEps) / ((U + kEps) * (U + kEps)) / cnt;
      }
    dice = cnt > 0 ? acc / cnt : 0.f;
  } else if (dice_mode == 2) {
    // GeneralizedDiceLoss: batch-global class weights 1 / (count^2 + 1e-9)
    float num = 0.f, den = 0.f, wk[KMAX];
    for (int k = 0; k < K; ++k) {
      double pt = 0, ps = 0, ts = 0;
      for (int n = 0; n < N; ++n) {
        const double* s = sums + (static_cast<int64_t>(n) * K + k) * 4;
        pt += s[0]; ps += s[1]; ts += s[2];
      }
      const long long cnt = static_cast<long long>(ts + 0.5);
      wk[k] = 1.0f / (static_cast<float>(cnt * cnt) + 1e-9f);
      num += wk[k] * static_cast<float>(pt);
      den += wk[k] * static_cast<float>(ts + ps);
    }
    dice = 1.f - 2.f * (num + 1e-9f) / (den + 1e-9f);
    for (int n = 0; n < N; ++n)
      for (int k = 0; k < K; ++k) {
        coef[(n * K + k) * 2 + 0] = -2.f * wk[k] / (den + 1e-9f);
        coef[(n * K + k) * 2 + 1] = 2.f * (num + 1e-9f) * wk[k] / ((den + 1e-9f) * (den + 1e-9f));
      }
  }
  float fs = 0.f;
  if (use_focal) {
    double f = 0, t = 0;
    for (int i = 0; i < N * K; ++i) { f += sums[i * 4 + 3]; t += sums[i * 4 + 2]; }
    fs = 1.f / (static_cast<float>(t) + kEps);
    focal = -static_cast<float>(f) * fs;
  }
  *focal_scale = fs;
  // smp Fscore (beta 1, eps 1e-7, threshold 0.5), micro-averaged over the kept channels
  float fscore[2];
  for (int with_bg = 0; with_bg < 2; ++with_bg) {
    long long tp = 0, pr = 0, gt = 0;
    for (int k = with_bg ? 0 : 1; k < K; ++k) { tp += counts[k * 3]; pr += counts[k * 3 + 1]; gt += counts[k * 3 + 2]; }
    const float ftp = static_cast<float>(tp), ffp = static_cast<float>(pr - tp), ffn = static_cast<float>(gt - tp);
    fscore[with_bg] = (2.f * ftp + 1e-7f) / (2.f * ftp + ffn + ffp + 1e-7f);
  }
  out[0] = dice; out[1] = focal; out[2] = dice + focal; out[3] = fscore[0]; out[4] = fscore[1];
  out[5] = out[6] = out[7] = 0.f;
}

template <int K>
__global__ void loss_backward_kernel(const float* __restrict__ logits, const int64_t* __restrict__ labels, int64_t HW,
                                     const float* __restrict__ coef, const float* __restrict__ focal_scale,
                                     float upstream, float* __restrict__ grad) {
  const int n = blockIdx.y;
  const float* zl = logits + static_cast<int64_t>(n) * K * HW;
  const int64_t* ll = labels + static_cast<int64_t>(n) * HW;
  float* gl = grad + static_cast<int64_t>(n) * K * HW;
  float a[K], b[K];
#pragma unroll
  for (int k = 0; k < K; ++k) { a[k] = coef[(n * K + k) * 2]; b[k] = coef[(n * K + k) * 2 + 1]; }
  const float fs = *focal_scale;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < HW;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    float p[K], g[K];
    softmax_px<K>(zl + i, HW, p);
    const int64_t lab = ll[i];
    float dot = 0.f;
#pragma unroll
    for (int k = 0; k < K; ++k) {
      g[k] = b[k];
      if (lab == k) {
        const float q = 1.f - p[k], pe = p[k] + kEps;
        g[k] += a[k] - fs * (-2.f * q * logf(pe) + q * q / pe);
      }
      dot += g[k] * p[k];
    }
#pragma unroll
    for (int k = 0; k < K; ++k) gl[k * HW + i] = upstream * p[k] * (g[k] - dot);
  }
}

// Boundary (surface) loss on the logits: L = scale_v * sum over (b, k in idc, pixels) softmax(z)_k * dist_k with
// scale_v = 1 / (N * |idc| * H * W)  (SurfaceLoss: mean of probs[:, idc] * dist_maps[:, idc], losses.py:250-270).
//   value pass:    one double partial per block, summed in a fixed order by the caller's tiny reduction
//   gradient pass: grad_j += scale_g * p_j * (dist_j [j in idc] - S),  S = sum_{k in idc} dist_k * p_k
template <int K>
__global__ void boundary_partials_kernel(const float* __restrict__ logits, const float* __restrict__ dist, int64_t HW,
                                         unsigned idc_mask, double* __restrict__ part) {
  const int n = blockIdx.y;
  const float* zl = logits + static_cast<int64_t>(n) * K * HW;
  const float* dl = dist + static_cast<int64_t>(n) * K * HW;
  float acc = 0.f;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < HW;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    float p[K];
    softmax_px<K>(zl + i, HW, p);
#pragma unroll
    for (int k = 0; k < K; ++k)
      if ((idc_mask >> k) & 1u) acc += p[k] * __ldg(dl + k * HW + i);
  }
  __shared__ double sh[kThreads];
  sh[threadIdx.x] = static_cast<double>(acc);
  __syncthreads();
  for (int o = kThreads / 2; o > 0; o >>= 1) {
    if (threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) part[static_cast<int64_t>(blockIdx.y) * gridDim.x + blockIdx.x] = sh[0];
}

__global__ void boundary_reduce_kernel(const double* __restrict__ part, int n, double scale, float* __restrict__ out) {
  __shared__ double sh[kThreads];
  double a = 0.0;
  for (int i = threadIdx.x; i < n; i += kThreads) a += part[i];
  sh[threadIdx.x] = a;
  __syncthreads();
  for (int o = kThreads / 2; o > 0; o >>= 1) {
    if (threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) *out = static_cast<float>(sh[0] * scale);
}

template <int K>
__global__ void boundary_backward_kernel(const float* __restrict__ logits, const float* __restrict__ dist, int64_t HW,
                                         unsigned idc_mask, float scale, float* __restrict__ grad) {
  const int n = blockIdx.y;
  const float* zl = logits + static_cast<int64_t>(n) * K * HW;
  const float* dl = dist + static_cast<int64_t>(n) * K * HW;
  float* gl = grad + static_cast<int64_t>(n) * K * HW;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < HW;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    float p[K], d[K];
    softmax_px<K>(zl + i, HW, p);
    float S = 0.f;
#pragma unroll
    for (int k = 0; k < K; ++k) {
      d[k] = ((idc_mask >> k) & 1u) ? __ldg(dl + k * HW + i) : 0.f;
      S += d[k] * p[k];
    }
#pragma unroll
    for (int k = 0; k < K; ++k) gl[k * HW + i] += scale * p[k] * (d[k] - S);
  }
}

// Generalized Wasserstein Dice loss (deadtrees/loss/gwdl.py:84-138, weighting "default"), on what SemSegment.calculate_loss
// hands it: the module is called with the SOFTMAX PROBABILITIES as `input` and applies softmax again itself
// (segmodel.py:176-178, gwdl.py:104) - q = softmax(p), p = softmax(logits), reproduced here (`twice`).
//   W[b,s]  = sum_c M[t][c] * q_c                      (Wasserstein distance to the label's class, M normalised to max 1)
//   u[s]    = sum_j (1 - W[j,s])                       (over the samples of the batch)
//   tp[b]   = sum_s alpha[b,s] * u[s],  alpha = 0 for the background class 0, 1 otherwise
//   err[b]  = sum_s W[b,s];     loss = mean_b ( 1 - (2 tp + eps) / (2 tp + err + eps) ),  eps = 2^-52
// The sum over j in tp is what the reference computes: compute_generalized_true_positive multiplies alpha (B, 1, S) with the
// distance map (B, S), which broadcasts to (B, B, S), and sums over dims [1, 2] (gwdl.py:187-205); for B = 1 it is the
// per-sample sum of the paper.  Passes: W map + err partials; u[s]; tp partials; finalize (loss, coef = {dL/dtp, dL/derr} and
// the per-pixel backward term A[s] = sum_i dL/dtp_i * alpha[i,s]); backward: dL/dW[j,s] = dL/derr_j - A[s], through both softmaxes.
struct GwdlMatrix { float m[KMAX * KMAX]; };

template <int K>
__device__ __forceinline__ void softmax_regs(const float (&z)[K], float (&p)[K]) {
  float mx = z[0];
#pragma unroll
  for (int k = 1; k < K; ++k) mx = fmaxf(mx, z[k]);
  float sum = 0.f;
#pragma unroll
  for (int k = 0; k < K; ++k) { p[k] = expf(z[k] - mx); sum += p[k]; }
  const float inv = 1.f / sum;
#pragma unroll
  for (int k = 0; k < K; ++k) p[k] *= inv;
}

// block-wide fixed-order sum of one float per thread into part[slot] (double)
__device__ __forceinline__ void gwdl_block_sum(float v, double* __restrict__ dst) {
  __shared__ double sh[kThreads];
  sh[threadIdx.x] = static_cast<double>(v);
  __syncthreads();
  for (int o = kThreads / 2; o > 0; o >>= 1) {
    if (threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) *dst = sh[0];
  __syncthreads();
}

// grid (chunks, N): omw[n][s] = 1 - W[n,s];  part[(n * chunks + c) * 2 + 1] = sum of W over the chunk
template <int K>
__global__ void gwdl_map_kernel(const float* __restrict__ logits, const int64_t* __restrict__ labels, int64_t HW,
                                GwdlMatrix M, int twice, float* __restrict__ omw, double* __restrict__ part) {
  const int n = blockIdx.y;
  const float* zl = logits + static_cast<int64_t>(n) * K * HW;
  const int64_t* ll = labels + static_cast<int64_t>(n) * HW;
  float* ol = omw + static_cast<int64_t>(n) * HW;
  float err = 0.f;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < HW;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    float p[K], q[K];
    softmax_px<K>(zl + i, HW, p);
    if (twice) softmax_regs<K>(p, q);
    const int t = static_cast<int>(ll[i]);
    float W = 0.f;
#pragma unroll
    for (int k = 0; k < K; ++k) W += M.m[(t < 0 || t >= K ? 0 : t) * K + k] * (twice ? q[k] : p[k]);
    err += W;
    ol[i] = 1.f - W;
  }
  gwdl_block_sum(err, part + (static_cast<int64_t>(n) * gridDim.x + blockIdx.x) * 2 + 1);
}

// u[s] = sum_j omw[j][s], in sample order
__global__ void gwdl_batch_sum_kernel(const float* __restrict__ omw, int N, int64_t HW, float* __restrict__ u) {
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < HW;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    float acc = 0.f;
    for (int j = 0; j < N; ++j) acc += omw[static_cast<int64_t>(j) * HW + i];
    u[i] = acc;
  }
}

// grid (chunks, N): part[(n * chunks + c) * 2] = sum over the chunk of alpha[n,s] * u[s]
__global__ void gwdl_tp_kernel(const int64_t* __restrict__ labels, const float* __restrict__ u, int64_t HW,
                               double* __restrict__ part) {
  const int n = blockIdx.y;
  const int64_t* ll = labels + static_cast<int64_t>(n) * HW;
  float tp = 0.f;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < HW;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x)
    if (ll[i] != 0) tp += u[i];
  gwdl_block_sum(tp, part + (static_cast<int64_t>(n) * gridDim.x + blockIdx.x) * 2);
}

// one block: per image the fixed-order sum of the chunk partials, the loss, and the backward coefficients
// coef[2b] = d loss / d tp_b, coef[2b+1] = d loss / d err_b
__global__ void gwdl_finalize_kernel(const double* __restrict__ part, int N, int chunks, float* __restrict__ loss,
                                     float* __restrict__ coef) {
  __shared__ float sh[kThreads];
  float acc = 0.f;
  const float eps = 2.220446049250313e-16f;
  for (int b = threadIdx.x; b < N; b += kThreads) {
    double tp = 0.0, err = 0.0;
    for (int c = 0; c < chunks; ++c) { tp += part[(static_cast<int64_t>(b) * chunks + c) * 2]; err += part[(static_cast<int64_t>(b) * chunks + c) * 2 + 1]; }
    const float ftp = static_cast<float>(tp), ferr = static_cast<float>(err);
    const float num = 2.f * ftp + eps, den = 2.f * ftp + ferr + eps;
    acc += 1.f - num / den;
    coef[2 * b] = -(2.f * den - 2.f * num) / (den * den) / static_cast<float>(N);
    coef[2 * b + 1] = num / (den * den) / static_cast<float>(N);
  }
  sh[threadIdx.x] = acc;
  __syncthreads();
  for (int o = kThreads / 2; o > 0; o >>= 1) {
    if (threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) *loss = sh[0] / static_cast<float>(N);
}

// amap[s] = sum_i coef[2i] * alpha[i,s]  (what every sample's W[.,s] contributes to through the true-positive terms)
__global__ void gwdl_alpha_map_kernel(const int64_t* __restrict__ labels, const float* __restrict__ coef, int N, int64_t HW,
                                      float* __restrict__ amap) {
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < HW;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    float acc = 0.f;
    for (int j = 0; j < N; ++j)
      if (labels[static_cast<int64_t>(j) * HW + i] != 0) acc += coef[2 * j];
    amap[i] = acc;
  }
}

template <int K>
__global__ void gwdl_backward_kernel(const float* __restrict__ logits, const int64_t* __restrict__ labels, int64_t HW,
                                     GwdlMatrix M, int twice, const float* __restrict__ coef, const float* __restrict__ amap,
                                     float weight, float* __restrict__ grad) {
  const int n = blockIdx.y;
  const float* zl = logits + static_cast<int64_t>(n) * K * HW;
  const int64_t* ll = labels + static_cast<int64_t>(n) * HW;
  float* gl = grad + static_cast<int64_t>(n) * K * HW;
  const float cb = coef[2 * n + 1];
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < HW;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    float p[K], q[K], g[K];
    softmax_px<K>(zl + i, HW, p);
    if (twice) softmax_regs<K>(p, q);
    const int t = static_cast<int>(ll[i]);
    const float* row = M.m + (t < 0 || t >= K ? 0 : t) * K;
    const float gW = weight * (cb - amap[i]);                         // d loss / d W[n,s]
    if (twice) {
      float W = 0.f;
#pragma unroll
      for (int k = 0; k < K; ++k) W += row[k] * q[k];
#pragma unroll
      for (int k = 0; k < K; ++k) g[k] = gW * q[k] * (row[k] - W);    // d loss / d p_k through q = softmax(p)
    } else {
#pragma unroll
      for (int k = 0; k < K; ++k) g[k] = gW * row[k];
    }
    float dot = 0.f;
#pragma unroll
    for (int k = 0; k < K; ++k) dot += g[k] * p[k];
#pragma unroll
    for (int k = 0; k < K; ++k) gl[k * HW + i] += p[k] * (g[k] - dot);  // through p = softmax(logits)
  }
}

// ---- confusion matrices of the validation / test epoch (segmodel.py:291-407) ----
// counts[0][t][p]: every pixel; counts[1][t][p]: pixels of the forest mask (lu == 1), the reference's "masked" matrices
// (torchmetrics confusion_matrix = bincount(target * K + pred), rows = target).  Integer histogram: per-warp private
// counters in shared memory, one 64-bit atomic per non-zero bin and block; ACCUMULATES into counts, so the batches of an
// epoch can be fed one by one without concatenating them.  bad: set when a label or prediction is outside [0, K).
constexpr int KCM = 16;

template <typename TP, typename TL>
__global__ void confusion_kernel(const TP* __restrict__ pred, const int64_t* __restrict__ target, const TL* __restrict__ lu,
                                 int64_t n, int K, unsigned long long* __restrict__ counts, int* __restrict__ bad) {
  extern __shared__ unsigned int cm_hist[];                       // [warps][2][K*K]
  const int KK = K * K, warps = blockDim.x >> 5;
  for (int i = threadIdx.x; i < warps * 2 * KK; i += blockDim.x) cm_hist[i] = 0u;
  __syncthreads();
  unsigned int* mine = cm_hist + (threadIdx.x >> 5) * 2 * KK;
  bool oob = false;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const long long t = target[i], p = static_cast<long long>(pred[i]);
    if (t < 0 || t >= K || p < 0 || p >= K) { oob = true; continue; }
    const int bin = static_cast<int>(t) * K + static_cast<int>(p);
    atomicAdd(&mine[bin], 1u);
    if (lu != nullptr && lu[i] == static_cast<TL>(1)) atomicAdd(&mine[KK + bin], 1u);
  }
  if (oob) atomicOr(bad, 1);
  __syncthreads();
  for (int b = threadIdx.x; b < 2 * KK; b += blockDim.x) {
    unsigned long long sum = 0;
    for (int w = 0; w < warps; ++w) sum += cm_hist[w * 2 * KK + b];
    if (sum) atomicAdd(&counts[b], sum);
  }
}

// ---- probability / one-hot API (the reference's loss callables take softmax output + one-hot) ----

__global__ void one_hot_kernel(const int64_t* __restrict__ labels, int K, int64_t HW, int64_t total,
                               int32_t* __restrict__ out, int* __restrict__ bad_label) {
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t n = i / HW, px = i % HW;
    const int64_t lab = labels[i];
    if (lab < 0 || lab >= K) atomicOr(bad_label, 1);
    for (int k = 0; k < K; ++k) out[(n * K + k) * HW + px] = (lab == k) ? 1 : 0;
  }
}

__global__ void softmax_nchw_kernel(const float* __restrict__ z, int K, int64_t HW, int64_t total,
                                    float* __restrict__ p) {
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t n = i / HW, px = i % HW;
    const float* zp = z + n * K * HW + px;
    float* pp = p + n * K * HW + px;
    float m = zp[0];
    for (int k = 1; k < K; ++k) m = fmaxf(m, zp[k * HW]);
    float s = 0.f;
    for (int k = 0; k < K; ++k) s += expf(zp[k * HW] - m);
    for (int k = 0; k < K; ++k) pp[k * HW] = expf(zp[k * HW] - m) / s;
  }
}

// grid = (chunks, N*K): one (n, k) plane per blockIdx.y.  sums[n][k][4] as in loss_partials_kernel,
// with the focal exponent gamma and a target that is either int32 one-hot or a float map.
template <typename TT>
__global__ void prob_partials_kernel(const float* __restrict__ probs, const TT* __restrict__ target, int64_t HW,
                                     float gamma, double* __restrict__ sums) {
  const int64_t plane = blockIdx.y;
  const float* pp = probs + plane * HW;
  const TT* tp = target + plane * HW;
  float s_pt = 0.f, s_p = 0.f, s_t = 0.f, s_f = 0.f;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < HW;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const float p = pp[i], t = static_cast<float>(tp[i]);
    s_pt += p * t;
    s_p += p;
    s_t += t;
    if (t != 0.f) {
      const float q = 1.f - p;
      const float w = gamma == 2.f ? q * q : powf(q, gamma);
      s_f += w * t * logf(p + kEps);
    }
  }
  __shared__ double sh[kThreads / 32][4];
  const double a = warp_sum(static_cast<double>(s_pt)), b = warp_sum(static_cast<double>(s_p));
  const double c = warp_sum(static_cast<double>(s_t)), d = warp_sum(static_cast<double>(s_f));
  if ((threadIdx.x & 31) == 0) {
    double* r = sh[threadIdx.x >> 5];
    r[0] = a; r[1] = b; r[2] = c; r[3] = d;
  }
  __syncthreads();
  if (threadIdx.x < 4) {
    double t = 0;
    for (int w = 0; w < kThreads / 32; ++w) t += sh[w][threadIdx.x];
    atomicAdd(&sums[plane * 4 + threadIdx.x], t);
  }
}

__global__ void sumsq_kernel(const float* __restrict__ g, int64_t n, double* __restrict__ out) {
  float acc = 0.f;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x)
    acc = fmaf(g[i], g[i], acc);
  double d = warp_sum(static_cast<double>(acc));
  __shared__ double sh[kThreads / 32];
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = d;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0;
    for (int w = 0; w < kThreads / 32; ++w) t += sh[w];
    atomicAdd(out, t);
  }
}

__global__ void adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                            float* __restrict__ v, int64_t n, float lr, float b1, float b2, float eps, float bc1,
                            float bc2_sqrt, const double* __restrict__ sumsq, float max_norm) {
  float clip = 1.f;
  if (max_norm > 0.f) {
    const float norm = static_cast<float>(sqrt(*sumsq));
    clip = fminf(1.f, max_norm / (norm + 1e-6f));
  }
  const float step_size = lr / bc1;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const float gi = g[i] * clip;
    const float mi = m[i] + (1.f - b1) * (gi - m[i]);            // torch: exp_avg.lerp_(grad, 1 - beta1)
    const float vi = b2 * v[i] + (1.f - b2) * gi * gi;           // exp_avg_sq.mul_(b2).addcmul_(g, g, 1 - b2)
    m[i] = mi;
    v[i] = vi;
    const float denom = sqrtf(vi) / bc2_sqrt + eps;
    p[i] -= step_size * (mi / denom);
  }
}

// Adam with ALL step state on the device, so the launch can live in a CUDA graph: state = {step (as float), lr};
// the step is skipped entirely (state untouched) when *loss is not finite - SemSegment.training_step returning None.
__global__ void adam_prepare_kernel(float* __restrict__ state, const float* __restrict__ loss, const double* __restrict__ sumsq,
                                    float b1, float b2, float max_norm, float* __restrict__ scal) {
  // skipped when the loss is not finite; with data parallelism the decision must be the same on every rank: a rank with a
  // non-finite loss contributes non-finite gradients, the all-reduced gradient norm is then non-finite everywhere
  const bool apply = (loss == nullptr || isfinite(*loss)) && (max_norm <= 0.f || isfinite(*sumsq));
  float step = state[0];
  if (apply) { step += 1.f; state[0] = step; }
  const double bc1 = 1.0 - pow(static_cast<double>(b1), static_cast<double>(step));
  const double bc2 = 1.0 - pow(static_cast<double>(b2), static_cast<double>(step));
  float clip = 1.f;
  if (max_norm > 0.f) clip = fminf(1.f, max_norm / (static_cast<float>(sqrt(*sumsq)) + 1e-6f));
  scal[0] = apply ? 1.f : 0.f;
  scal[1] = clip;
  scal[2] = state[1] / static_cast<float>(bc1);        // step size = lr / bias_correction1
  scal[3] = static_cast<float>(sqrt(bc2));
}

__global__ void adam_dev_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                float* __restrict__ v, int64_t n, float b1, float b2, float eps,
                                const float* __restrict__ scal) {
  if (scal[0] == 0.f) return;
  const float clip = scal[1], step_size = scal[2], bc2_sqrt = scal[3];
  for (int64_t i = (blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x) * 4; i < n;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x * 4) {
    if (i + 3 < n) {
      float4 p4 = *reinterpret_cast<float4*>(p + i), m4 = *reinterpret_cast<float4*>(m + i), v4 = *reinterpret_cast<float4*>(v + i);
      const float4 g4 = *reinterpret_cast<const float4*>(g + i);
      float* pp = &p4.x; float* mm = &m4.x; float* vv = &v4.x; const float* gg = &g4.x;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float gi = gg[j] * clip;
        mm[j] = mm[j] + (1.f - b1) * (gi - mm[j]);
        vv[j] = b2 * vv[j] + (1.f - b2) * gi * gi;
        pp[j] -= step_size * (mm[j] / (sqrtf(vv[j]) / bc2_sqrt + eps));
      }
      *reinterpret_cast<float4*>(p + i) = p4; *reinterpret_cast<float4*>(m + i) = m4; *reinterpret_cast<float4*>(v + i) = v4;
    } else {
      for (int64_t k = i; k < n; ++k) {
        const float gi = g[k] * clip;
        const float mi = m[k] + (1.f - b1) * (gi - m[k]);
        const float vi = b2 * v[k] + (1.f - b2) * gi * gi;
        m[k] = mi; v[k] = vi;
        p[k] -= step_size * (mi / (sqrtf(vi) / bc2_sqrt + eps));
      }
    }
  }
}

inline int grid_for(int64_t work) {
  int64_t blocks = (work + kThreads - 1) / kThreads;
  const int64_t cap = static_cast<int64_t>(dt_num_sms()) * 8;
  return static_cast<int>(blocks < 1 ? 1 : (blocks > cap ? cap : blocks));
}

template <typename TP>
static int confusion_launch(const TP* pred, const int64_t* target, const void* lu, int lu_elem, int64_t n, int K,
                            int64_t* counts, int32_t* bad, cudaStream_t s) {
  // every thread's 32-bit private counter is bounded by the pixels one block sees: keep blocks under 2^31 pixels
  int grid = grid_for(n);
  const size_t smem = static_cast<size_t>(kThreads / 32) * 2 * K * K * sizeof(unsigned int);
  unsigned long long* c = reinterpret_cast<unsigned long long*>(counts);
  if (lu == nullptr || lu_elem == 1)
    confusion_kernel<TP, uint8_t><<<grid, kThreads, smem, s>>>(pred, target, static_cast<const uint8_t*>(lu), n, K, c, bad);
  else if (lu_elem == 4)
    confusion_kernel<TP, int32_t><<<grid, kThreads, smem, s>>>(pred, target, static_cast<const int32_t*>(lu), n, K, c, bad);
  else
    confusion_kernel<TP, int64_t><<<grid, kThreads, smem, s>>>(pred, target, static_cast<const int64_t*>(lu), n, K, c, bad);
  return 0;
}

}  // namespace

extern "C" {

int dt_seg_loss_partials(const float* logits, const int64_t* labels, int N, int K, int H, int W, double* sums,
                         int64_t* counts, int32_t* bad_label, dt_stream_t stream) {
  DT_ARCH_GUARD();
  DT_REQUIRE(N > 0 && N <= 65535 && H > 0 && W > 0, DT_ERR_BAD_SHAPE, "dt_seg_loss_partials: bad shape");
  DT_REQUIRE(K >= 2 && K <= KMAX, DT_ERR_BAD_SHAPE, "dt_seg_loss_partials: K=%d (2..4)", K);
  const int64_t HW = static_cast<int64_t>(H) * W;
  int chunks = static_cast<int>((HW + kThreads * 8 - 1) / (kThreads * 8));
  const int cap = (dt_num_sms() * 8 + N - 1) / N;
  if (chunks > cap) chunks = cap;
  if (chunks < 1) chunks = 1;
  dim3 grid(chunks, N);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  unsigned long long* c = reinterpret_cast<unsigned long long*>(counts);
  switch (K) {
    case 2: loss_partials_kernel<2><<<grid, kThreads, 0, s>>>(logits, labels, HW, sums, c, bad_label); break;
    case 3: loss_partials_kernel<3><<<grid, kThreads, 0, s>>>(logits, labels, HW, sums, c, bad_label); break;
    default: loss_partials_kernel<4><<<grid, kThreads, 0, s>>>(logits, labels, HW, sums, c, bad_label); break;
  }
  DT_LAUNCH_CHECK();
  return DT_OK;
}

int dt_seg_loss_finalize(const double* sums, const int64_t* counts, int N, int K, int dice_mode, int use_focal,
                         float* out, float* coef, float* focal_scale, dt_stream_t stream) {
  DT_ARCH_GUARD();
  DT_REQUIRE(N > 0 && K >= 2 && K <= KMAX && dice_mode >= 0 && dice_mode <= 2, DT_ERR_BAD_SHAPE,
             "dt_seg_loss_finalize: bad arguments");
  const size_t bytes = static_cast<size_t>(N) * K * 4 * sizeof(double);
  const int staged = bytes <= 40 * 1024;
  loss_finalize_kernel<<<1, 256, staged ? bytes : 0, static_cast<cudaStream_t>(stream)>>>(
      sums, reinterpret_cast<const long long*>(counts), N, K, dice_mode, use_focal, out, coef, focal_scale, staged);
  DT_LAUNCH_CHECK();
  return DT_OK;
}

int dt_seg_loss_backward(const float* logits, const int64_t* labels, int N, int K, int H, int W, const float* coef,
                         const float* focal_scale, float upstream, float* grad_logits, dt_stream_t stream) {
  DT_ARCH_GUARD();
  DT_REQUIRE(N > 0 && N <= 65535 && H > 0 && W > 0 && K >= 2 && K <= KMAX, DT_ERR_BAD_SHAPE,
             "dt_seg_loss_backward: bad shape");
  const int64_t HW = static_cast<int64_t>(H) * W;
  int chunks = static_cast<int>((HW + kThreads * 4 - 1) / (kThreads * 4));
  if (chunks < 1) chunks = 1;
  dim3 grid(chunks, N);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  switch (K) {
    case 2: loss_backward_kernel<2><<<grid, kThreads, 0, s>>>(logits, labels, HW, coef, focal_scale, upstream, grad_logits); break;
    case 3: loss_backward_kernel<3><<<grid, kThreads, 0, s>>>(logits, labels, HW, coef, focal_scale, upstream, grad_logits); break;
    default: loss_backward_kernel<4><<<grid, kThreads, 0, s>>>(logits, labels, HW, coef, focal_scale, upstream, grad_logits); break;
  }
  DT_LAUNCH_CHECK();
  return DT_OK;
}

int dt_boundary_loss(const float* logits, const float* dist, int N, int K, int H, int W, unsigned idc_mask,
                     double* workspace, float* loss_out, dt_stream_t stream) {
  DT_ARCH_GUARD();
  DT_REQUIRE(N > 0 && N <= 65535 && H > 0 && W > 0 && K >= 2 && K <= KMAX && idc_mask != 0 && (idc_mask >> K) == 0,
             DT_ERR_BAD_SHAPE, "dt_boundary_loss: bad shape / class subset");
  const int64_t HW = static_cast<int64_t>(H) * W;
  int chunks = static_cast<int>((HW + kThreads * 4 - 1) / (kThreads * 4));
  if (chunks > 64) chunks = 64;
  if (chunks < 1) chunks = 1;
  dim3 grid(chunks, N);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  switch (K) {
    case 2: boundary_partials_kernel<2><<<grid, kThreads, 0, s>>>(logits, dist, HW, idc_mask, workspace); break;
    case 3: boundary_partials_kernel<3><<<grid, kThreads, 0, s>>>(logits, dist, HW, idc_mask, workspace); break;
    default: boundary_partials_kernel<4><<<grid, kThreads, 0, s>>>(logits, dist, HW, idc_mask, workspace); break;
  }
  DT_LAUNCH_CHECK();
  const double cnt = static_cast<double>(N) * __builtin_popcount(idc_mask) * static_cast<double>(HW);
  boundary_reduce_kernel<<<1, kThreads, 0, s>>>(workspace, chunks * N, 1.0 / cnt, loss_out);
  DT_LAUNCH_CHECK();
  return DT_OK;
}

int dt_boundary_loss_backward(const float* logits, const float* dist, int N, int K, int H, int W, unsigned idc_mask,
                              float weight, float* grad_logits, dt_stream_t stream) {
  DT_ARCH_GUARD();
  DT_REQUIRE(N > 0 && N <= 65535 && H > 0 && W > 0 && K >= 2 && K <= KMAX && idc_mask != 0 && (idc_mask >> K) == 0,
             DT_ERR_BAD_SHAPE, "dt_boundary_loss_backward: bad shape / class subset");
  const int64_t HW = static_cast<int64_t>(H) * W;
  int chunks = static_cast<int>((HW + kThreads * 4 - 1) / (kThreads * 4));
  if (chunks < 1) chunks = 1;
  dim3 grid(chunks, N);
  const float scale = weight / (static_cast<float>(N) * __builtin_popcount(idc_mask) * static_cast<float>(HW));
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  switch (K) {
    case 2: boundary_backward_kernel<2><<<grid, kThreads, 0, s>>>(logits, dist, HW, idc_mask, scale, grad_logits); break;
    case 3: boundary_backward_kernel<3><<<grid, kThreads, 0, s>>>(logits, dist, HW, idc_mask, scale, grad_logits); break;
    default: boundary_backward_kernel<4><<<grid, kThreads, 0, s>>>(logits, dist, HW, idc_mask, scale, grad_logits); break;
  }
  DT_LAUNCH_CHECK();
  return DT_OK;
}

static int gwdl_chunks(int64_t HW) {
  int chunks = static_cast<int>((HW + kThreads * 4 - 1) / (kThreads * 4));
  return chunks > 64 ? 64 : (chunks < 1 ? 1 : chunks);
}

static bool gwdl_matrix(const float* dist_matrix, int K, GwdlMatrix* M) {
  float mx = 0.f;
  for (int i = 0; i < K * K; ++i) mx = dist_matrix[i] > mx ? dist_matrix[i] : mx;
  if (!(mx > 0.f)) return false;
  for (int i = 0; i < KMAX * KMAX; ++i) M->m[i] = i < K * K ? dist_matrix[i] / mx : 0.f;     // normalised to max 1 (gwdl.py:72-77)
  return true;
}

// workspace: [2 * 64 * N doubles: chunk partials][N * HW floats: 1 - W][HW floats: u]
int64_t dt_gwdl_workspace(int N, int H, int W) {
  if (N <= 0 || H <= 0 || W <= 0) return 0;
  const int64_t HW = static_cast<int64_t>(H) * W;
  return static_cast<int64_t>(sizeof(double)) * 2 * 64 * N + static_cast<int64_t>(sizeof(float)) * (static_cast<int64_t>(N) + 1) * HW;
}

int dt_gwdl_loss(const float* logits, const int64_t* labels, int N, int K, int H, int W, const float* dist_matrix,
                 int softmax_twice, void* workspace, int64_t workspace_bytes, float* loss_out, float* coef,
                 dt_stream_t stream) {
  DT_ARCH_GUARD();
  DT_REQUIRE(N > 0 && N <= 65535 && H > 0 && W > 0 && K >= 2 && K <= KMAX && dist_matrix != nullptr, DT_ERR_BAD_SHAPE,
             "dt_gwdl_loss: bad shape");
  DT_REQUIRE(workspace != nullptr && workspace_bytes >= dt_gwdl_workspace(N, H, W), DT_ERR_BAD_SHAPE,
             "dt_gwdl_loss: workspace smaller than dt_gwdl_workspace()");
  GwdlMatrix M;
  DT_REQUIRE(gwdl_matrix(dist_matrix, K, &M), DT_ERR_BAD_SHAPE,
             "dt_gwdl_loss: the class distance matrix must have a positive maximum");
  const int64_t HW = static_cast<int64_t>(H) * W;
  const int chunks = gwdl_chunks(HW);
  double* part = static_cast<double*>(workspace);
  float* omw = reinterpret_cast<float*>(part + static_cast<size_t>(2) * 64 * N);
  float* u = omw + static_cast<size_t>(N) * HW;
  float* amap = coef + 2 * static_cast<size_t>(N);
  dim3 grid(chunks, N);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  switch (K) {
    case 2: gwdl_map_kernel<2><<<grid, kThreads, 0, s>>>(logits, labels, HW, M, softmax_twice, omw, part); break;
    case 3: gwdl_map_kernel<3><<<grid, kThreads, 0, s>>>(logits, labels, HW, M, softmax_twice, omw, part); break;
    default: gwdl_map_kernel<4><<<grid, kThreads, 0, s>>>(logits, labels, HW, M, softmax_twice, omw, part); break;
  }
  DT_LAUNCH_CHECK();
  gwdl_batch_sum_kernel<<<grid_for(HW), kThreads, 0, s>>>(omw, N, HW, u);
  DT_LAUNCH_CHECK();
  gwdl_tp_kernel<<<grid, kThreads, 0, s>>>(labels, u, HW, part);
  DT_LAUNCH_CHECK();
  gwdl_finalize_kernel<<<1, kThreads, 0, s>>>(part, N, chunks, loss_out, coef);
  DT_LAUNCH_CHECK();
  gwdl_alpha_map_kernel<<<grid_for(HW), kThreads, 0, s>>>(labels, coef, N, HW, amap);
  DT_LAUNCH_CHECK();
  return DT_OK;
}

int dt_gwdl_loss_backward(const float* logits, const int64_t* labels, int N, int K, int H, int W, const float* dist_matrix,
                          int softmax_twice, const float* coef, float weight, float* grad_logits, dt_stream_t stream) {
  DT_ARCH_GUARD();
  DT_REQUIRE(N > 0 && N <= 65535 && H > 0 && W > 0 && K >= 2 && K <= KMAX && dist_matrix != nullptr && coef != nullptr,
             DT_ERR_BAD_SHAPE, "dt_gwdl_loss_backward: bad shape");
  GwdlMatrix M;
  DT_REQUIRE(gwdl_matrix(dist_matrix, K, &M), DT_ERR_BAD_SHAPE,
             "dt_gwdl_loss_backward: the class distance matrix must have a positive maximum");
  const int64_t HW = static_cast<int64_t>(H) * W;
  int chunks = static_cast<int>((HW + kThreads * 4 - 1) / (kThreads * 4));
  if (chunks < 1) chunks = 1;
  dim3 grid(chunks, N);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const float* amap = coef + 2 * static_cast<size_t>(N);
  switch (K) {
    case 2: gwdl_backward_kernel<2><<<grid, kThreads, 0, s>>>(logits, labels, HW, M, softmax_twice, coef, amap, weight, grad_logits); break;
    case 3: gwdl_backward_kernel<3><<<grid, kThreads, 0, s>>>(logits, labels, HW, M, softmax_twice, coef, amap, weight, grad_logits); break;
    default: gwdl_backward_kernel<4><<<grid, kThreads, 0, s>>>(logits, labels, HW, M, softmax_twice, coef, amap, weight, grad_logits); break;
  }
  DT_LAUNCH_CHECK();
  return DT_OK;
}

int dt_confusion_matrix(const void* pred, int pred_elem, const int64_t* target, const void* lu, int lu_elem, int64_t n,
                        int K, int64_t* counts, int32_t* bad_label, dt_stream_t stream) {
  DT_ARCH_GUARD();
  DT_REQUIRE(n > 0 && K >= 2 && K <= KCM, DT_ERR_BAD_SHAPE, "dt_confusion_matrix: n=%lld K=%d (K in 2..16)",
             static_cast<long long>(n), K);
  DT_REQUIRE((pred_elem == 1 || pred_elem == 8) && (lu == nullptr || lu_elem == 1 || lu_elem == 4 || lu_elem == 8),
             DT_ERR_BAD_SHAPE, "dt_confusion_matrix: pred must be uint8 or int64, lu uint8 / int32 / int64");
  DT_REQUIRE(pred != nullptr && target != nullptr && counts != nullptr && bad_label != nullptr, DT_ERR_BAD_SHAPE,
             "dt_confusion_matrix: null pointer");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (pred_elem == 1) confusion_launch(static_cast<const uint8_t*>(pred), target, lu, lu_elem, n, K, counts, bad_label, s);
  else confusion_launch(static_cast<const int64_t*>(pred), target, lu, lu_elem, n, K, counts, bad_label, s);
  DT_LAUNCH_CHECK();
  return DT_OK;
}

int dt_class2one_hot(const int64_t* labels, int N, int K, int H, int W, int32_t* onehot, int32_t* bad_label,
                     dt_stream_t stream) {
  DT_ARCH_GUARD();
  DT_REQUIRE(N > 0 && K > 0 && H > 0 && W > 0, DT_ERR_BAD_SHAPE, "dt_class2one_hot: bad shape");
  const int64_t HW = static_cast<int64_t>(H) * W, total = HW * N;
  one_hot_kernel<<<grid_for(total), kThreads, 0, static_cast<cudaStream_t>(stream)>>>(labels, K, HW, total, onehot,
                                                                                     bad_label);
  DT_LAUNCH_CHECK();
  return DT_OK;
}

int dt_softmax_nchw(const float* logits, int N, int K, int H, int W, float* probs, dt_stream_t stream) {
  DT_ARCH_GUARD();
  DT_REQUIRE(N > 0 && K > 0 && H > 0 && W > 0, DT_ERR_BAD_SHAPE, "dt_softmax_nchw: bad shape");
  const int64_t HW = static_cast<int64_t>(H) * W, total = HW * N;
  softmax_nchw_kernel<<<grid_for(total), kThreads, 0, static_cast<cudaStream_t>(stream)>>>(logits, K, HW, total, probs);
  DT_LAUNCH_CHECK();
  return DT_OK;
}

int dt_prob_loss_partials(const float* probs, const void* target, int target_is_float, int N, int K, int H, int W,
                          float gamma, double* sums, dt_stream_t stream) {
  DT_ARCH_GUARD();
  DT_REQUIRE(N > 0 && K > 0 && H > 0 && W > 0 && static_cast<int64_t>(N) * K <= 65535, DT_ERR_BAD_SHAPE,
             "dt_prob_loss_partials: bad shape");
  const int64_t HW = static_cast<int64_t>(H) * W;
  int chunks = static_cast<int>((HW + kThreads * 8 - 1) / (kThreads * 8));
  const int cap = (dt_num_sms() * 8 + N * K - 1) / (N * K);
  if (chunks > cap) chunks = cap;
  if (chunks < 1) chunks = 1;
  dim3 grid(chunks, N * K);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (target_is_float)
    prob_partials_kernel<float><<<grid, kThreads, 0, s>>>(probs, static_cast<const float*>(target), HW, gamma, sums);
  else
    prob_partials_kernel<int32_t><<<grid, kThreads, 0, s>>>(probs, static_cast<const int32_t*>(target), HW, gamma, sums);
  DT_LAUNCH_CHECK();
  return DT_OK;
}

int dt_sumsq(const float* g, int64_t n, double* sumsq, dt_stream_t stream) {
  DT_ARCH_GUARD();
  DT_REQUIRE(n >= 0, DT_ERR_BAD_SHAPE, "dt_sumsq: n < 0");
  if (n == 0) return DT_OK;
  sumsq_kernel<<<grid_for(n), kThreads, 0, static_cast<cudaStream_t>(stream)>>>(g, n, sumsq);
  DT_LAUNCH_CHECK();
  return DT_OK;
}

int dt_adam_step(float* p, const float* g, float* m, float* v, int64_t n, float lr, float beta1, float beta2,
                 float eps, int step, const double* sumsq, float max_norm, dt_stream_t stream) {
  DT_ARCH_GUARD();
  DT_REQUIRE(n >= 0 && step >= 1, DT_ERR_BAD_SHAPE, "dt_adam_step: bad arguments");
  DT_REQUIRE(max_norm <= 0.f || sumsq != nullptr, DT_ERR_BAD_SHAPE, "dt_adam_step: clipping needs sumsq");
  if (n == 0) return DT_OK;
  // bias corrections in double on the host, as torch.optim.Adam does in Python floats
  const float bc1 = static_cast<float>(1.0 - pow(static_cast<double>(beta1), step));
  const float bc2_sqrt = static_cast<float>(sqrt(1.0 - pow(static_cast<double>(beta2), step)));
  adam_kernel<<<grid_for(n), kThreads, 0, static_cast<cudaStream_t>(stream)>>>(p, g, m, v, n, lr, beta1, beta2, eps,
                                                                               bc1, bc2_sqrt, sumsq, max_norm);
  DT_LAUNCH_CHECK();
  return DT_OK;
}

int dt_adam_step_dev(float* p, const float* g, float* m, float* v, int64_t n, float* state, float beta1, float beta2,
                     float eps, const double* sumsq, float max_norm, const float* loss, float* scratch4,
                     dt_stream_t stream) {
  DT_ARCH_GUARD();
  DT_REQUIRE(n >= 0 && state != nullptr && scratch4 != nullptr, DT_ERR_BAD_SHAPE, "dt_adam_step_dev: bad arguments");
  DT_REQUIRE(max_norm <= 0.f || sumsq != nullptr, DT_ERR_BAD_SHAPE, "dt_adam_step_dev: clipping needs sumsq");
  DT_REQUIRE((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m) |
              reinterpret_cast<uintptr_t>(v)) % 16 == 0, DT_ERR_BAD_ALIGN, "dt_adam_step_dev: buffers must be 16-byte aligned");
  if (n == 0) return DT_OK;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  adam_prepare_kernel<<<1, 1, 0, s>>>(state, loss, sumsq, beta1, beta2, max_norm, scratch4);
  DT_LAUNCH_CHECK();
  adam_dev_kernel<<<grid_for((n + 3) / 4), kThreads, 0, s>>>(p, g, m, v, n, beta1, beta2, eps, scratch4);
  DT_LAUNCH_CHECK();
  return DT_OK;
}

}  // extern "C"
