// CUDA-core kernels: fp32 "check mode" direct convolution (also usable on bf16 tensors for
// validation of the tensor-core path), maxpool 3x3/s2, segmentation head + argmax.
// Reference behaviour: smp.Unet(resnet34) forward as called at deadtrees/network/segmodel.py:214 and
// deadtrees/deployment/inference.py:60-62 (see include/deadtrees_b200.h).
#include "common.cuh"

namespace {

constexpr int kThreads = 256;

template <typename T> __device__ __forceinline__ float to_f(T v);
template <> __device__ __forceinline__ float to_f<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T from_f(float v);
template <> __device__ __forceinline__ float from_f<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 from_f<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

struct DirectParams {
  int N, H, W, C_in, C_x, upsample, C_out, R, S, stride, pad, relu, has_residual;
  int Ho, Wo, Kpad, stem;
};

// One thread = one output element (co fastest).  TA = activation type; TW = weight type:
// float weights are [tap][C_in][C_out]; bf16 weights are the tensor-core packing [C_out][Kpad].
template <typename TA, typename TW>
__global__ void conv_direct_kernel(DirectParams p, const TA* __restrict__ x, const TA* __restrict__ skip,
                                   const TW* __restrict__ w, const float* __restrict__ scale,
                                   const float* __restrict__ shift, const TA* __restrict__ residual,
                                   TA* __restrict__ y) {
  const int64_t total = static_cast<int64_t>(p.N) * p.Ho * p.Wo * p.C_out;
  const int C_s = p.C_in - p.C_x;
  const int Hx = p.upsample ? p.H >> 1 : p.H, Wx = p.upsample ? p.W >> 1 : p.W;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int co = static_cast<int>(i % p.C_out);
    int64_t r = i / p.C_out;
    const int wo = static_cast<int>(r % p.Wo); r /= p.Wo;
    const int ho = static_cast<int>(r % p.Ho);
    const int n = static_cast<int>(r / p.Ho);
    float acc = 0.f;
    for (int fr = 0; fr < p.R; ++fr) {
      const int hi = ho * p.stride + fr - p.pad;
      if (hi < 0 || hi >= p.H) continue;
      for (int fs = 0; fs < p.S; ++fs) {
        const int wi = wo * p.stride + fs - p.pad;
        if (wi < 0 || wi >= p.W) continue;
        const int tap = fr * p.S + fs;
        const TA* xp = p.upsample ? x + ((static_cast<int64_t>(n) * Hx + (hi >> 1)) * Wx + (wi >> 1)) * p.C_x
                                  : x + ((static_cast<int64_t>(n) * Hx + hi) * Wx + wi) * p.C_x;
        const TA* sp = C_s > 0 ? skip + ((static_cast<int64_t>(n) * p.H + hi) * p.W + wi) * C_s : nullptr;
        for (int ci = 0; ci < p.C_in; ++ci) {
          const float a = ci < p.C_x ? to_f<TA>(xp[ci]) : to_f<TA>(sp[ci - p.C_x]);
          float wv;
          if (sizeof(TW) == 4) {
            wv = to_f<TW>(w[(static_cast<int64_t>(tap) * p.C_in + ci) * p.C_out + co]);
          } else {
            const int k = p.stem ? fr * 32 + fs * 4 + ci : tap * p.C_in + ci;
            wv = to_f<TW>(w[static_cast<int64_t>(co) * p.Kpad + k]);
          }
          acc = fmaf(a, wv, acc);
        }
      }
    }
    float v = fmaf(acc, scale[co], shift[co]);
    if (p.has_residual) v += to_f<TA>(residual[i]);
    if (p.relu) v = fmaxf(v, 0.f);
    y[i] = from_f<TA>(v);
  }
}

template <typename TA>
__global__ void maxpool3x3s2_kernel(const TA* __restrict__ x, int N, int H, int W, int C, int Ho, int Wo,
                                    TA* __restrict__ y) {
  const int64_t total = static_cast<int64_t>(N) * Ho * Wo * C;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % C);
    int64_t r = i / C;
    const int wo = static_cast<int>(r % Wo); r /= Wo;
    const int ho = static_cast<int>(r % Ho);
    const int n = static_cast<int>(r / Ho);
    float m = -INFINITY;
    for (int dy = 0; dy < 3; ++dy) {
      const int hi = 2 * ho + dy - 1;
      if (hi < 0 || hi >= H) continue;
      for (int dx = 0; dx < 3; ++dx) {
        const int wi = 2 * wo + dx - 1;
        if (wi < 0 || wi >= W) continue;
        m = fmaxf(m, to_f<TA>(x[((static_cast<int64_t>(n) * H + hi) * W + wi) * C + c]));
      }
    }
    y[i] = from_f<TA>(m);
  }
}

// bf16 fast path: one thread = 8 channels (16 bytes) of one output pixel.
__global__ void maxpool3x3s2_bf16x8_kernel(const uint4* __restrict__ x, int N, int H, int W, int C8, int Ho, int Wo,
                                           uint4* __restrict__ y) {
  const int64_t total = static_cast<int64_t>(N) * Ho * Wo * C8;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % C8);
    int64_t r = i / C8;
    const int wo = static_cast<int>(r % Wo); r /= Wo;
    const int ho = static_cast<int>(r % Ho);
    const int n = static_cast<int>(r / Ho);
    __nv_bfloat162 m[4];
    const __nv_bfloat162 ninf = __float2bfloat162_rn(-INFINITY);
#pragma unroll
    for (int j = 0; j < 4; ++j) m[j] = ninf;
#pragma unroll
    for (int dy = 0; dy < 3; ++dy) {
      const int hi = 2 * ho + dy - 1;
      if (hi < 0 || hi >= H) continue;
#pragma unroll
      for (int dx = 0; dx < 3; ++dx) {
        const int wi = 2 * wo + dx - 1;
        if (wi < 0 || wi >= W) continue;
        const uint4 v = __ldg(x + ((static_cast<int64_t>(n) * H + hi) * W + wi) * C8 + c);
        const __nv_bfloat162* pv = reinterpret_cast<const __nv_bfloat162*>(&v);
#pragma unroll
        for (int j = 0; j < 4; ++j) m[j] = __hmax2(m[j], pv[j]);
      }
    }
    y[i] = *reinterpret_cast<uint4*>(m);
  }
}

// Segmentation head: 3x3 conv C -> K (+bias) with fused argmax / layout outputs.  One thread = one pixel.
template <typename TA, int K>
__global__ void head_kernel(const TA* __restrict__ x, int N, int H, int W, int C, const float* __restrict__ w,
                            const float* __restrict__ bias, float* __restrict__ logits_nchw,
                            TA* __restrict__ logits_nhwc, uint8_t* __restrict__ mask) {
  extern __shared__ float sw[];  // [9][C][K]
  for (int i = threadIdx.x; i < 9 * C * K; i += blockDim.x) sw[i] = w[i];
  __syncthreads();
  const int64_t total = static_cast<int64_t>(N) * H * W;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int xw = static_cast<int>(i % W);
    const int yh = static_cast<int>((i / W) % H);
    const int n = static_cast<int>(i / (static_cast<int64_t>(W) * H));
    float acc[K];
#pragma unroll
    for (int k = 0; k < K; ++k) acc[k] = 0.f;
    for (int dy = 0; dy < 3; ++dy) {
      const int hi = yh + dy - 1;
      if (hi < 0 || hi >= H) continue;
      for (int dx = 0; dx < 3; ++dx) {
        const int wi = xw + dx - 1;
        if (wi < 0 || wi >= W) continue;
        const TA* xp = x + ((static_cast<int64_t>(n) * H + hi) * W + wi) * C;
        const float* wp = sw + (dy * 3 + dx) * C * K;
        for (int c = 0; c < C; ++c) {
          const float a = to_f<TA>(xp[c]);
#pragma unroll
          for (int k = 0; k < K; ++k) acc[k] = fmaf(a, wp[c * K + k], acc[k]);
        }
      }
    }
    int best = 0;
    float bv = 0.f;
#pragma unroll
    for (int k = 0; k < K; ++k) {
      const float v = acc[k] + bias[k];
      if (logits_nchw) logits_nchw[((static_cast<int64_t>(n) * K + k) * H + yh) * W + xw] = v;
      if (logits_nhwc) logits_nhwc[i * K + k] = from_f<TA>(v);
      if (k == 0 || v > bv) { bv = v; best = k; }
    }
    if (mask) mask[i] = static_cast<uint8_t>(best);
  }
}

// 16 input channels as one vector load per tap
__device__ __forceinline__ void load16(const __nv_bfloat16* p, float (&a)[16]) {
  const uint4 v0 = __ldg(reinterpret_cast<const uint4*>(p)), v1 = __ldg(reinterpret_cast<const uint4*>(p) + 1);
  const uint32_t u[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float2 t = unpack_bf16x2(u[j]);
    a[2 * j] = t.x;
    a[2 * j + 1] = t.y;
  }
}
__device__ __forceinline__ void load16(const float* p, float (&a)[16]) {
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float4 t = __ldg(reinterpret_cast<const float4*>(p) + j);
    a[4 * j] = t.x; a[4 * j + 1] = t.y; a[4 * j + 2] = t.z; a[4 * j + 3] = t.w;
  }
}

// C == 16 fast path of the head: one thread = one pixel, vector loads, weights broadcast from smem.
// Accumulation order (tap-major, channel-minor, fp32 FMA) is the same as head_kernel's.
template <typename TA, int K>
__global__ void __launch_bounds__(256) head16_kernel(const TA* __restrict__ x, int N, int H, int W,
                                                     const float* __restrict__ w, const float* __restrict__ bias,
                                                     float* __restrict__ logits_nchw, TA* __restrict__ logits_nhwc,
                                                     uint8_t* __restrict__ mask) {
  __shared__ float sw[9 * 16 * K];
  for (int i = threadIdx.x; i < 9 * 16 * K; i += blockDim.x) sw[i] = w[i];
  __syncthreads();
  const int64_t total = static_cast<int64_t>(N) * H * W;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int xw = static_cast<int>(i % W);
    const int yh = static_cast<int>((i / W) % H);
    const int n = static_cast<int>(i / (static_cast<int64_t>(W) * H));
    float acc[K];
#pragma unroll
    for (int k = 0; k < K; ++k) acc[k] = 0.f;
#pragma unroll
    for (int dy = 0; dy < 3; ++dy) {
      const int hi = yh + dy - 1;
      if (hi < 0 || hi >= H) continue;
#pragma unroll
      for (int dx = 0; dx < 3; ++dx) {
        const int wi = xw + dx - 1;
        if (wi < 0 || wi >= W) continue;
        float a[16];
        load16(x + ((static_cast<int64_t>(n) * H + hi) * W + wi) * 16, a);
        const float* wp = sw + (dy * 3 + dx) * 16 * K;
#pragma unroll
        for (int c = 0; c < 16; ++c) {
#pragma unroll
          for (int k = 0; k < K; ++k) acc[k] = fmaf(a[c], wp[c * K + k], acc[k]);
        }
      }
    }
    int best = 0;
    float bv = 0.f;
#pragma unroll
    for (int k = 0; k < K; ++k) {
      const float v = acc[k] + bias[k];
      if (logits_nchw) logits_nchw[((static_cast<int64_t>(n) * K + k) * H + yh) * W + xw] = v;
      if (logits_nhwc) logits_nhwc[i * K + k] = from_f<TA>(v);
      if (k == 0 || v > bv) { bv = v; best = k; }
    }
    if (mask) mask[i] = static_cast<uint8_t>(best);
  }
}

__global__ void argmax_nchw_kernel(const float* __restrict__ logits, int N, int K, int64_t HW,
                                   uint8_t* __restrict__ mask) {
  const int64_t total = static_cast<int64_t>(N) * HW;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t n = i / HW, px = i % HW;
    const float* p = logits + n * K * HW + px;
    int best = 0;
    float bv = p[0];
    for (int k = 1; k < K; ++k) {
      const float v = p[k * HW];
      if (v > bv) { bv = v; best = k; }
    }
    mask[i] = static_cast<uint8_t>(best);
  }
}

// Per-pixel majority vote over M class-id masks (PyTorchEnsembleInference.run, deadtrees/deployment/inference.py:96-116:
// torch.mode over the stacked argmax masks).  torch.mode returns the SMALLEST of the most frequent values; class ids
// are < 256.  One thread = 16 consecutive pixels: one 16-byte load per model, counts by pairwise comparison
// (M <= 15), int64 (what the reference returns) or uint8 output.
template <typename OUT>
__global__ void mode_vote_kernel(const uint8_t* __restrict__ masks, int M, int64_t n, OUT* __restrict__ out) {
  const int64_t nvec = (n + 15) / 16;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < nvec;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t p0 = i * 16;
    uint8_t v[15][16];
    const bool fast = p0 + 16 <= n && n % 16 == 0;
    for (int m = 0; m < M; ++m) {
      if (fast) {
        *reinterpret_cast<uint4*>(v[m]) = __ldg(reinterpret_cast<const uint4*>(masks + m * n + p0));
      } else {
        for (int j = 0; j < 16; ++j) v[m][j] = p0 + j < n ? masks[m * n + p0 + j] : 0;
      }
    }
    for (int j = 0; j < 16; ++j) {
      if (p0 + j >= n) break;
      int best_count = 0, best_val = 256;
      for (int a = 0; a < M; ++a) {
        int c = 0;
        for (int b = 0; b < M; ++b) c += v[b][j] == v[a][j];
        if (c > best_count || (c == best_count && v[a][j] < best_val)) { best_count = c; best_val = v[a][j]; }
      }
      out[p0 + j] = static_cast<OUT>(best_val);
    }
  }
}

inline int grid_for(int64_t work) {
  int64_t blocks = (work + kThreads - 1) / kThreads;
  const int64_t cap = static_cast<int64_t>(dt_num_sms()) * 16;
  return static_cast<int>(blocks < 1 ? 1 : (blocks > cap ? cap : blocks));
}

}  // namespace

// Called by dt_conv2d_fwd (conv_tc.cu) for the fp32 check mode and the FORCE_DIRECT validation path.
int dt_conv2d_direct(const dt_conv_desc* d, int Ho, int Wo, int Kpad, int stem, const void* x, const void* skip,
                     const void* w, const float* scale, const float* shift, const void* residual, void* y,
                     cudaStream_t s) {
  DirectParams p{d->N, d->H, d->W, d->C_in, d->C_x, d->upsample, d->C_out, d->R, d->S, d->stride, d->pad,
                 d->relu, d->has_residual, Ho, Wo, Kpad, stem};
  const int64_t total = static_cast<int64_t>(d->N) * Ho * Wo * d->C_out;
  if (d->dtype == DT_F32) {
    conv_direct_kernel<float, float><<<grid_for(total), kThreads, 0, s>>>(
        p, static_cast<const float*>(x), static_cast<const float*>(skip), static_cast<const float*>(w), scale, shift,
        static_cast<const float*>(residual), static_cast<float*>(y));
  } else {
    conv_direct_kernel<__nv_bfloat16, __nv_bfloat16><<<grid_for(total), kThreads, 0, s>>>(
        p, static_cast<const __nv_bfloat16*>(x), static_cast<const __nv_bfloat16*>(skip),
        static_cast<const __nv_bfloat16*>(w), scale, shift, static_cast<const __nv_bfloat16*>(residual),
        static_cast<__nv_bfloat16*>(y));
  }
  DT_LAUNCH_CHECK();
  return DT_OK;
}

int dt_head_row(const void* x, int N, int H, int W, int K, const void* w_packed, const float* bias16, float* logits_nchw,
                void* logits_nhwc, uint8_t* mask, cudaStream_t s);
int dt_head_res(const void* x, int N, int H, int W, int K, const void* w_packed, const float* bias16,
                float* logits_nchw, void* logits_nhwc, uint8_t* mask, cudaStream_t s);

extern "C" {

int dt_maxpool3x3s2(const void* x, int N, int H, int W, int C, int dtype, void* y, dt_stream_t stream) {
  DT_ARCH_GUARD();
  DT_REQUIRE(N > 0 && H > 0 && W > 0 && C > 0, DT_ERR_BAD_SHAPE, "dt_maxpool3x3s2: bad shape");
  DT_REQUIRE(dtype == DT_BF16 || dtype == DT_F32, DT_ERR_BAD_SHAPE, "dt_maxpool3x3s2: dtype %d", dtype);
  const int Ho = (H + 2 - 3) / 2 + 1, Wo = (W + 2 - 3) / 2 + 1;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int64_t total = static_cast<int64_t>(N) * Ho * Wo * C;
  if (dtype == DT_BF16 && C % 8 == 0 && reinterpret_cast<uintptr_t>(x) % 16 == 0 &&
      reinterpret_cast<uintptr_t>(y) % 16 == 0) {
    maxpool3x3s2_bf16x8_kernel<<<grid_for(total / 8), kThreads, 0, s>>>(static_cast<const uint4*>(x), N, H, W, C / 8,
                                                                        Ho, Wo, static_cast<uint4*>(y));
  } else if (dtype == DT_BF16) {
    maxpool3x3s2_kernel<__nv_bfloat16><<<grid_for(total), kThreads, 0, s>>>(static_cast<const __nv_bfloat16*>(x), N,
                                                                            H, W, C, Ho, Wo,
                                                                            static_cast<__nv_bfloat16*>(y));
  } else {
    maxpool3x3s2_kernel<float><<<grid_for(total), kThreads, 0, s>>>(static_cast<const float*>(x), N, H, W, C, Ho, Wo,
                                                                    static_cast<float*>(y));
  }
  DT_LAUNCH_CHECK();
  return DT_OK;
}

int dt_head_fwd(const void* x, int x_dtype, int N, int H, int W, int C, int K, const float* w, const float* bias,
                float* logits_nchw, void* logits_nhwc, uint8_t* mask, dt_stream_t stream) {
  DT_ARCH_GUARD();
  DT_REQUIRE(N > 0 && H > 0 && W > 0 && C > 0 && C <= 64, DT_ERR_BAD_SHAPE, "dt_head_fwd: bad shape (C=%d)", C);
  DT_REQUIRE(K >= 1 && K <= 4, DT_ERR_BAD_SHAPE, "dt_head_fwd: K=%d (1..4)", K);
  DT_REQUIRE(x_dtype == DT_BF16 || x_dtype == DT_F32, DT_ERR_BAD_SHAPE, "dt_head_fwd: dtype %d", x_dtype);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int64_t total = static_cast<int64_t>(N) * H * W;
  const size_t smem = static_cast<size_t>(9) * C * K * sizeof(float);
  const bool fast = C == 16 && reinterpret_cast<uintptr_t>(x) % 16 == 0;
#define DT_HEAD(TA, KK)                                                                                       \
  if (fast)                                                                                                   \
    head16_kernel<TA, KK><<<grid_for(total), kThreads, 0, s>>>(static_cast<const TA*>(x), N, H, W, w, bias,   \
                                                               logits_nchw, static_cast<TA*>(logits_nhwc), mask); \
  else                                                                                                        \
    head_kernel<TA, KK><<<grid_for(total), kThreads, smem, s>>>(static_cast<const TA*>(x), N, H, W, C, w, bias, \
                                                               logits_nchw, static_cast<TA*>(logits_nhwc), mask)
#define DT_HEADK(TA)                 \
  switch (K) {                       \
    case 1: DT_HEAD(TA, 1); break;   \
    case 2: DT_HEAD(TA, 2); break;   \
    case 3: DT_HEAD(TA, 3); break;   \
    default: DT_HEAD(TA, 4); break;  \
  }
  if (x_dtype == DT_BF16) { DT_HEADK(__nv_bfloat16) } else { DT_HEADK(float) }
#undef DT_HEADK
#undef DT_HEAD
  DT_LAUNCH_CHECK();
  return DT_OK;
}

int dt_head_fwd_tc(const void* x, int N, int H, int W, int K, const void* w_packed, const float* bias16,
                   float* logits_nchw, void* logits_nhwc, uint8_t* mask, dt_stream_t stream) {
  DT_ARCH_GUARD();
  DT_REQUIRE(N > 0 && H > 0 && W > 0 && K >= 1 && K <= 4, DT_ERR_BAD_SHAPE, "dt_head_fwd_tc: bad shape");
  DT_REQUIRE(reinterpret_cast<uintptr_t>(x) % 16 == 0 && reinterpret_cast<uintptr_t>(w_packed) % 16 == 0,
             DT_ERR_BAD_ALIGN, "dt_head_fwd_tc: tensors must be 16-byte aligned");
  if (dt_row_kernels_enabled()) {   // row-streaming kernel (vertical taps in N) where the width allows; same results
    const int rr = dt_head_row(x, N, H, W, K, w_packed, bias16, logits_nchw, logits_nhwc, mask, static_cast<cudaStream_t>(stream));
    if (rr != DT_ERR_UNSUPPORTED) return rr;
  }
  const int rc = dt_head_res(x, N, H, W, K, w_packed, bias16, logits_nchw, logits_nhwc, mask,
                             static_cast<cudaStream_t>(stream));
  DT_REQUIRE(rc != DT_ERR_UNSUPPORTED, DT_ERR_BAD_SHAPE, "dt_head_fwd_tc: needs H %% 16 == 0 and W %% 8 == 0 (got %dx%d)", H, W);
  return rc;
}

int dt_argmax_nchw(const float* logits, int N, int K, int H, int W, uint8_t* mask, dt_stream_t stream) {
  DT_ARCH_GUARD();
  DT_REQUIRE(N > 0 && K > 0 && K <= 255 && H > 0 && W > 0, DT_ERR_BAD_SHAPE, "dt_argmax_nchw: bad shape");
  const int64_t HW = static_cast<int64_t>(H) * W;
  argmax_nchw_kernel<<<grid_for(N * HW), kThreads, 0, static_cast<cudaStream_t>(stream)>>>(logits, N, K, HW, mask);
  DT_LAUNCH_CHECK();
  return DT_OK;
}

int dt_mode_vote(const uint8_t* masks, int M, int64_t n, int out_int64, void* out, dt_stream_t stream) {
  DT_ARCH_GUARD();
  DT_REQUIRE(masks != nullptr && out != nullptr && M >= 1 && M <= 15 && n > 0, DT_ERR_BAD_SHAPE,
             "dt_mode_vote: M=%d (1..15) masks of n=%lld pixels", M, static_cast<long long>(n));
  DT_REQUIRE(reinterpret_cast<uintptr_t>(masks) % 16 == 0, DT_ERR_BAD_ALIGN, "dt_mode_vote: masks must be 16-byte aligned");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (out_int64)
    mode_vote_kernel<int64_t><<<grid_for((n + 15) / 16), kThreads, 0, s>>>(masks, M, n, static_cast<int64_t*>(out));
  else
    mode_vote_kernel<uint8_t><<<grid_for((n + 15) / 16), kThreads, 0, s>>>(masks, M, n, static_cast<uint8_t*>(out));
  DT_LAUNCH_CHECK();
  return DT_OK;
}

}  // extern "C"
