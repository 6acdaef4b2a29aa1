// Training-step kernels that are HBM-bound or generic (CUDA cores): train-mode BatchNorm forward / backward
// (two-stage deterministic per-channel reductions), ReLU masks, maxpool backward, nearest-x2 + concat and its
// backward, weight repacking, and generic direct dgrad / wgrad used by the fp32 check mode and by the layer
// shapes the tcgen05 kernels do not cover.
//
// Reference behaviour: autograd of smp.Unet(resnet34) in train mode as driven by SemSegment.training_step
// (deadtrees/network/segmodel.py:210-229): nn.BatchNorm2d (batch statistics, biased variance for the
// normalisation, unbiased for running_var, momentum 0.1, eps 1e-5), ReLU, MaxPool2d(3, 2, 1),
// F.interpolate(nearest x2) + torch.cat (decoder convention: network/extra/resunet/decoder.py:41-43).
#include "common.cuh"

namespace {

constexpr int kThreads = 256;

template <typename T> struct VecOf;
template <> struct VecOf<float> { static constexpr int N = 4; };
template <> struct VecOf<__nv_bfloat16> { static constexpr int N = 8; };

template <typename T> __device__ __forceinline__ void load_vec(const T* p, float* f);
template <> __device__ __forceinline__ void load_vec<float>(const float* p, float* f) {
  const float4 v = *reinterpret_cast<const float4*>(p);
  f[0] = v.x; f[1] = v.y; f[2] = v.z; f[3] = v.w;
}
template <> __device__ __forceinline__ void load_vec<__nv_bfloat16>(const __nv_bfloat16* p, float* f) {
  const uint4 v = *reinterpret_cast<const uint4*>(p);
  const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float2 t = unpack_bf16x2(w[j]);
    f[2 * j] = t.x; f[2 * j + 1] = t.y;
  }
}
template <typename T> __device__ __forceinline__ void store_vec(T* p, const float* f);
template <> __device__ __forceinline__ void store_vec<float>(float* p, const float* f) {
  *reinterpret_cast<float4*>(p) = make_float4(f[0], f[1], f[2], f[3]);
}
template <> __device__ __forceinline__ void store_vec<__nv_bfloat16>(__nv_bfloat16* p, const float* f) {
  *reinterpret_cast<uint4*>(p) = make_uint4(pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]), pack_bf16x2(f[4], f[5]),
                                             pack_bf16x2(f[6], f[7]));
}
template <typename T> __device__ __forceinline__ float to_f(T v);
template <> __device__ __forceinline__ float to_f<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T from_f(float v);
template <> __device__ __forceinline__ float from_f<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 from_f<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

inline int grid_for(int64_t work, int per_sm = 8) {
  int64_t blocks = (work + kThreads - 1) / kThreads;
  const int64_t cap = static_cast<int64_t>(dt_num_sms()) * per_sm;
  return static_cast<int>(blocks < 1 ? 1 : (blocks > cap ? cap : blocks));
}

// ---- per-channel two-value reductions over an NHWC tensor ---------------------------------------------------
// MODE 0: (sum y, sum y^2)                      -> BatchNorm batch statistics
// MODE 1: (sum gz, sum gz * yhat), gz = g * [a > 0], yhat = (y - mean) * invstd  -> BatchNorm backward
// Thread = one 16-byte vector of channels, walking pixels; block partials are written to part[2][gridDim.x][C]
// and summed in a fixed order by the finalize kernels (deterministic, no atomics).
// raw 16-byte vector of channels (8 bf16 or 4 floats): kept packed while several loads are in flight
template <typename T> __device__ __forceinline__ uint4 load_raw(const T* p) { return *reinterpret_cast<const uint4*>(p); }
template <typename T> __device__ __forceinline__ void unpack_raw(const uint4& v, float* f);
template <> __device__ __forceinline__ void unpack_raw<float>(const uint4& v, float* f) {
  f[0] = __uint_as_float(v.x); f[1] = __uint_as_float(v.y); f[2] = __uint_as_float(v.z); f[3] = __uint_as_float(v.w);
}
template <> __device__ __forceinline__ void unpack_raw<__nv_bfloat16>(const uint4& v, float* f) {
  const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float2 t = unpack_bf16x2(w[j]);
    f[2 * j] = t.x; f[2 * j + 1] = t.y;
  }
}

// U pixels per loop trip: their loads are issued together (packed registers) and accumulated in pixel order, so the sums are
// the same bits for every U.  With one pixel per trip the statistics pass had 16 KB of loads in flight per SM (2.6 TB/s).
template <typename T, int MODE, int U>
__global__ void channel_reduce_kernel(const T* __restrict__ y, const T* __restrict__ g, const T* __restrict__ a,
                                      const float* __restrict__ mean, const float* __restrict__ invstd,
                                      const float* __restrict__ fscale, const float* __restrict__ fshift, int64_t M, int C,
                                      float* __restrict__ part) {
  constexpr int VN = VecOf<T>::N;
  const int V = C / VN;                    // vectors per pixel (divides kThreads)
  const int lanes = kThreads / V;
  const int v = threadIdx.x % V, lane = threadIdx.x / V;
  float s[VN], q[VN], mu[VN], is[VN], fsc[VN], fsh[VN];
#pragma unroll
  for (int j = 0; j < VN; ++j) { s[j] = q[j] = 0.f; mu[j] = 0.f; is[j] = 1.f; fsc[j] = 0.f; fsh[j] = 1.f; }
  // ReLU mask recomputed from y when the forward had no residual: a > 0  <=>  fma(y, scale, shift) > 0 (the value
  // bn_apply_kernel clamped), so the activation tensor is not read again
  const bool mask_y = MODE == 1 && a == nullptr && fshift != nullptr;
  if (MODE == 1) {
#pragma unroll
    for (int j = 0; j < VN; ++j) { mu[j] = mean[v * VN + j]; is[j] = invstd[v * VN + j]; }
    if (mask_y) {
#pragma unroll
      for (int j = 0; j < VN; ++j) { fsc[j] = fscale[v * VN + j]; fsh[j] = fshift[v * VN + j]; }
    }
  }
  auto accumulate = [&](const uint4& ry, const uint4& rg, const uint4& ra) {
    float fy[VN];
    unpack_raw<T>(ry, fy);
    if (MODE == 0) {
#pragma unroll
      for (int j = 0; j < VN; ++j) { s[j] += fy[j]; q[j] = fmaf(fy[j], fy[j], q[j]); }
    } else {
      float fg[VN], fa[VN];
      unpack_raw<T>(rg, fg);
      if (a) unpack_raw<T>(ra, fa);
#pragma unroll
      for (int j = 0; j < VN; ++j) {
        const bool on = mask_y ? fmaf(fy[j], fsc[j], fsh[j]) > 0.f : (a == nullptr || fa[j] > 0.f);
        const float gz = on ? fg[j] : 0.f;
        s[j] += gz;
        q[j] = fmaf(gz, (fy[j] - mu[j]) * is[j], q[j]);
      }
    }
  };
  const int64_t stride = static_cast<int64_t>(gridDim.x) * lanes;
  int64_t p = static_cast<int64_t>(blockIdx.x) * lanes + lane;
  for (; p + (U - 1) * stride < M; p += U * stride) {
    uint4 ry[U], rg[U], ra[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t off = (p + u * stride) * C + v * VN;
      ry[u] = load_raw<T>(y + off);
      if (MODE == 1) {
        rg[u] = load_raw<T>(g + off);
        if (a) ra[u] = load_raw<T>(a + off);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) accumulate(ry[u], rg[u], ra[u]);
  }
  for (; p < M; p += stride) {
    const int64_t off = p * C + v * VN;
    uint4 ry = load_raw<T>(y + off), rg = make_uint4(0u, 0u, 0u, 0u), ra = rg;
    if (MODE == 1) {
      rg = load_raw<T>(g + off);
      if (a) ra = load_raw<T>(a + off);
    }
    accumulate(ry, rg, ra);
  }
  __shared__ float sh[2][kThreads][VN + 1];
#pragma unroll
  for (int j = 0; j < VN; ++j) { sh[0][threadIdx.x][j] = s[j]; sh[1][threadIdx.x][j] = q[j]; }
  __syncthreads();
  if (lane == 0) {
#pragma unroll
    for (int j = 0; j < VN; ++j) {
      float ts = 0.f, tq = 0.f;
      for (int l = 0; l < lanes; ++l) { ts += sh[0][l * V + v][j]; tq += sh[1][l * V + v][j]; }
      part[static_cast<int64_t>(blockIdx.x) * C + v * VN + j] = ts;
      part[(static_cast<int64_t>(gridDim.x) + blockIdx.x) * C + v * VN + j] = tq;
    }
  }
}

// DT_BN_REDUCE_UNROLL=0 selects one pixel per loop trip (A/B testing; the results are bit-identical)
static bool reduce_unrolled() {
  static const bool on = [] { const char* e = getenv("DT_BN_REDUCE_UNROLL"); return !(e && e[0] == '0'); }();
  return on;
}

// sum of the block partials of channel c by one 128-thread block (fixed thread -> partial assignment and a fixed-order
// tree: deterministic).  One warp per channel took ~10 us for the ~600 partials of a reduction - 92 such launches
// per training step; 128 threads have at most 5 partials each.
constexpr int kFinThreads = 128;
__device__ __forceinline__ void block_channel_sums(const float* __restrict__ part, int nblocks, int C, int c, double& s, double& q) {
  __shared__ double sh_s[kFinThreads], sh_q[kFinThreads];
  double ts = 0.0, tq = 0.0;
  for (int b = threadIdx.x; b < nblocks; b += kFinThreads) {
    ts += part[static_cast<int64_t>(b) * C + c];
    tq += part[(static_cast<int64_t>(nblocks) + b) * C + c];
  }
  sh_s[threadIdx.x] = ts; sh_q[threadIdx.x] = tq;
  __syncthreads();
  for (int o = kFinThreads / 2; o > 0; o >>= 1) {
    if (threadIdx.x < o) { sh_s[threadIdx.x] += sh_s[threadIdx.x + o]; sh_q[threadIdx.x] += sh_q[threadIdx.x + o]; }
    __syncthreads();
  }
  s = sh_s[0]; q = sh_q[0];
}

// batch statistics -> (scale, shift) of the normalisation, saved (mean, invstd), running-stat update
__global__ void bn_finalize_kernel(const float* __restrict__ part, int nblocks, int C, double M, const float* __restrict__ gamma,
                                   const float* __restrict__ beta, float eps, float momentum,
                                   float* __restrict__ running_mean, float* __restrict__ running_var,
                                   float* __restrict__ scale, float* __restrict__ shift, float* __restrict__ mean_out,
                                   float* __restrict__ invstd_out) {
  const int c = blockIdx.x;
  double s, q;
  block_channel_sums(part, nblocks, C, c, s, q);
  if (threadIdx.x != 0) return;
  const double mean = s / M;
  double var = q / M - mean * mean;
  if (var < 0.0) var = 0.0;
  const float invstd = static_cast<float>(1.0 / sqrt(var + static_cast<double>(eps)));
  const float sc = gamma[c] * invstd;
  scale[c] = sc;
  shift[c] = beta[c] - static_cast<float>(mean) * sc;
  mean_out[c] = static_cast<float>(mean);
  invstd_out[c] = invstd;
  if (running_mean) {
    const double unbiased = M > 1.0 ? var * M / (M - 1.0) : var;
    running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * static_cast<float>(mean);
    running_var[c] = (1.f - momentum) * running_var[c] + momentum * static_cast<float>(unbiased);
  }
}

// BatchNorm backward coefficients: dgamma, dbeta and c1 = dbeta / M, c2 = dgamma / M
__global__ void bn_bwd_finalize_kernel(const float* __restrict__ part, int nblocks, int C, double M,
                                       float* __restrict__ dgamma, float* __restrict__ dbeta, float* __restrict__ c1,
                                       float* __restrict__ c2) {
  const int c = blockIdx.x;
  double s, q;
  block_channel_sums(part, nblocks, C, c, s, q);
  if (threadIdx.x != 0) return;
  dbeta[c] = static_cast<float>(s);
  dgamma[c] = static_cast<float>(q);
  c1[c] = static_cast<float>(s / M);
  c2[c] = static_cast<float>(q / M);
}

// a = [relu]( y * scale + shift (+ residual) )
// The block size and the grid stride are multiples of V = C / VN, so a thread keeps ONE channel group: its coefficients
// live in registers (round 1 re-read them through __ldg for every vector - 16 / 48 extra load instructions per 16 bytes of
// output, which left both apply passes LSU-bound at half of the copy bandwidth) and two vectors per trip are in flight.
template <typename T>
__global__ void __launch_bounds__(kThreads)
bn_apply_kernel(const T* __restrict__ y, const float* __restrict__ scale, const float* __restrict__ shift,
                const T* __restrict__ residual, int relu, int64_t nvec, int C, T* __restrict__ out) {
  constexpr int VN = VecOf<T>::N;
  const int V = C / VN;
  const int c0 = (threadIdx.x % V) * VN;
  float sc[VN], sh[VN];
#pragma unroll
  for (int j = 0; j < VN; ++j) { sc[j] = __ldg(scale + c0 + j); sh[j] = __ldg(shift + c0 + j); }
  auto apply = [&](float* f, const float* r) {
#pragma unroll
    for (int j = 0; j < VN; ++j) {
      float t = fmaf(f[j], sc[j], sh[j]);
      if (residual) t += r[j];
      f[j] = relu ? fmaxf(t, 0.f) : t;
    }
  };
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  for (; i + stride < nvec; i += 2 * stride) {
    float f0[VN], f1[VN], r0[VN], r1[VN];
    load_vec<T>(y + i * VN, f0);
    load_vec<T>(y + (i + stride) * VN, f1);
    if (residual) { load_vec<T>(residual + i * VN, r0); load_vec<T>(residual + (i + stride) * VN, r1); }
    apply(f0, r0);
    apply(f1, r1);
    store_vec<T>(out + i * VN, f0);
    store_vec<T>(out + (i + stride) * VN, f1);
  }
  if (i < nvec) {
    float f0[VN], r0[VN];
    load_vec<T>(y + i * VN, f0);
    if (residual) load_vec<T>(residual + i * VN, r0);
    apply(f0, r0);
    store_vec<T>(out + i * VN, f0);
  }
}

// gz = g * [a > 0];  gy = scale * (gz - c1 - yhat * c2);  optional second output gz (identity branch of a block)
// (coefficients in registers - see bn_apply_kernel)
template <typename T>
__global__ void __launch_bounds__(kThreads, 3)
bn_bwd_apply_kernel(const T* __restrict__ g, const T* __restrict__ a, const T* __restrict__ y,
                    const float* __restrict__ mean, const float* __restrict__ invstd,
                    const float* __restrict__ scale, const float* __restrict__ c1,
                    const float* __restrict__ c2, const float* __restrict__ fshift, int64_t nvec, int C,
                    T* __restrict__ gy, T* __restrict__ gz_out) {
  constexpr int VN = VecOf<T>::N;
  const int V = C / VN;
  const int c0 = (threadIdx.x % V) * VN;
  const bool mask_y = a == nullptr && fshift != nullptr;
  float sc[VN], mu[VN], is[VN], k1[VN], k2[VN], fsh[VN];
#pragma unroll
  for (int j = 0; j < VN; ++j) {
    sc[j] = __ldg(scale + c0 + j); mu[j] = __ldg(mean + c0 + j); is[j] = __ldg(invstd + c0 + j);
    k1[j] = __ldg(c1 + c0 + j); k2[j] = __ldg(c2 + c0 + j);
    fsh[j] = mask_y ? __ldg(fshift + c0 + j) : 0.f;
  }
  auto apply = [&](float* fg, const float* fa, const float* fy, float* o) {
#pragma unroll
    for (int j = 0; j < VN; ++j) {
      const bool on = mask_y ? fmaf(fy[j], sc[j], fsh[j]) > 0.f : (a == nullptr || fa[j] > 0.f);
      const float gz = on ? fg[j] : 0.f;
      fg[j] = gz;
      const float yhat = (fy[j] - mu[j]) * is[j];
      o[j] = sc[j] * (gz - k1[j] - yhat * k2[j]);
    }
  };
  // one vector per trip: with two in flight the 48 coefficient registers push the kernel past 96 registers (two blocks per SM)
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < nvec; i += stride) {
    float g0[VN], a0[VN], y0[VN], o0[VN];
    load_vec<T>(g + i * VN, g0);
    if (a) load_vec<T>(a + i * VN, a0);
    load_vec<T>(y + i * VN, y0);
    apply(g0, a0, y0, o0);
    store_vec<T>(gy + i * VN, o0);
    if (gz_out) store_vec<T>(gz_out + i * VN, g0);
  }
}

// out = a + b (vectorised), used to merge gradient contributions
template <typename T>
__global__ void add_kernel(const T* __restrict__ a, const T* __restrict__ b, int64_t nvec, T* __restrict__ out) {
  constexpr int VN = VecOf<T>::N;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < nvec;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    float fa[VN], fb[VN];
    load_vec<T>(a + i * VN, fa);
    load_vec<T>(b + i * VN, fb);
#pragma unroll
    for (int j = 0; j < VN; ++j) fa[j] += fb[j];
    store_vec<T>(out + i * VN, fa);
  }
}

// MaxPool2d(3, 2, 1) backward, gather form: every input element sums the gradients of the (<= 4) windows whose
// FIRST maximum (row-major scan, as ATen's max_pool2d_with_indices) it is; plus an optional addend.
template <typename T>
__global__ void maxpool_bwd_kernel(const T* __restrict__ x, const T* __restrict__ gout, const T* __restrict__ addend,
                                   int N, int H, int W, int C, int Ho, int Wo, T* __restrict__ gx) {
  const int64_t total = static_cast<int64_t>(N) * H * W * C;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % C);
    int64_t r = i / C;
    const int w = static_cast<int>(r % W); r /= W;
    const int h = static_cast<int>(r % H);
    const int n = static_cast<int>(r / H);
    const T* xn = x + static_cast<int64_t>(n) * H * W * C + c;
    float acc = addend ? to_f<T>(addend[i]) : 0.f;
    // windows (ho, wo) with 2*ho - 1 <= h <= 2*ho + 1
    for (int ho = (h >> 1); ho <= ((h + 1) >> 1); ++ho) {
      if (ho < 0 || ho >= Ho) continue;
      for (int wo = (w >> 1); wo <= ((w + 1) >> 1); ++wo) {
        if (wo < 0 || wo >= Wo) continue;
        float best = -INFINITY;
        int bh = -1, bw = -1;
        for (int dy = 0; dy < 3; ++dy) {
          const int hi = 2 * ho + dy - 1;
          if (hi < 0 || hi >= H) continue;
          for (int dx = 0; dx < 3; ++dx) {
            const int wi = 2 * wo + dx - 1;
            if (wi < 0 || wi >= W) continue;
            const float v = to_f<T>(xn[(static_cast<int64_t>(hi) * W + wi) * C]);
            if (v > best || bh < 0) { best = v; bh = hi; bw = wi; }
          }
        }
        if (bh == h && bw == w) acc += to_f<T>(gout[((static_cast<int64_t>(n) * Ho + ho) * Wo + wo) * C + c]);
      }
    }
    gx[i] = from_f<T>(acc);
  }
}

// MaxPool2d(3, 2, 1) for the training step, index form.  Forward: one thread = one 16-byte channel vector of one
// output pixel; also stores, per element, the window position (dy * 3 + dx, 0..8) of the FIRST maximum in row-major
// scan order (ATen's max_pool2d_with_indices).  Backward (gather, no atomics, no re-read of the pool input): one thread
// = one channel vector of one INPUT pixel; it belongs to 1, 2 or 4 windows and takes the gradient of each window
// whose stored position names it.  Algorithmic bytes per input element-vector: <= 4 x (VN index bytes + 16 B gout)
// + 16 B addend + 16 B gx, against 9 x 16 B of x per window in the recompute form above.
template <typename T>
__global__ void maxpool_idx_fwd_kernel(const T* __restrict__ x, int N, int H, int W, int C, int Ho, int Wo,
                                       T* __restrict__ y, uint8_t* __restrict__ idx) {
  constexpr int VN = VecOf<T>::N;
  const int Cv = C / VN;
  const int64_t total = static_cast<int64_t>(N) * Ho * Wo * Cv;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int cv = static_cast<int>(i % Cv);
    int64_t r = i / Cv;
    const int wo = static_cast<int>(r % Wo); r /= Wo;
    const int ho = static_cast<int>(r % Ho);
    const int n = static_cast<int>(r / Ho);
    float best[VN];
    uint32_t code[VN];
#pragma unroll
    for (int j = 0; j < VN; ++j) { best[j] = 0.f; code[j] = 255u; }
#pragma unroll
    for (int dy = 0; dy < 3; ++dy) {
      const int hi = 2 * ho + dy - 1;
      if (hi < 0 || hi >= H) continue;
#pragma unroll
      for (int dx = 0; dx < 3; ++dx) {
        const int wi = 2 * wo + dx - 1;
        if (wi < 0 || wi >= W) continue;
        float v[VN];
        load_vec<T>(x + ((static_cast<int64_t>(n) * H + hi) * W + wi) * C + cv * VN, v);
#pragma unroll
        for (int j = 0; j < VN; ++j)
          if (code[j] == 255u || v[j] > best[j]) { best[j] = v[j]; code[j] = dy * 3 + dx; }
      }
    }
    store_vec<T>(y + i * VN, best);
    uint32_t packed[VN / 4];
#pragma unroll
    for (int q = 0; q < VN / 4; ++q)
      packed[q] = code[4 * q] | (code[4 * q + 1] << 8) | (code[4 * q + 2] << 16) | (code[4 * q + 3] << 24);
    if (VN == 8) *reinterpret_cast<uint2*>(idx + i * VN) = make_uint2(packed[0], packed[VN / 4 - 1]);
    else *reinterpret_cast<uint32_t*>(idx + i * VN) = packed[0];
  }
}

template <typename T>
__global__ void maxpool_idx_bwd_kernel(const uint8_t* __restrict__ idx, const T* __restrict__ gout,
                                       const T* __restrict__ addend, int N, int H, int W, int C, int Ho, int Wo,
                                       T* __restrict__ gx) {
  constexpr int VN = VecOf<T>::N;
  const int Cv = C / VN;
  const int64_t total = static_cast<int64_t>(N) * H * W * Cv;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int cv = static_cast<int>(i % Cv);
    int64_t r = i / Cv;
    const int w = static_cast<int>(r % W); r /= W;
    const int h = static_cast<int>(r % H);
    const int n = static_cast<int>(r / H);
    float acc[VN];
    if (addend != nullptr) {
      load_vec<T>(addend + i * VN, acc);
    } else {
#pragma unroll
      for (int j = 0; j < VN; ++j) acc[j] = 0.f;
    }
    // windows (ho, wo) with 2*ho - 1 <= h <= 2*ho + 1: ho = h/2 and, for odd h, also (h+1)/2
    const int ho1 = (h + 1) >> 1, wo1 = (w + 1) >> 1;
    for (int ho = h >> 1; ho <= ho1; ++ho) {
      if (ho >= Ho) continue;
      const uint32_t cy = static_cast<uint32_t>(h - 2 * ho + 1) * 3u;
      for (int wo = w >> 1; wo <= wo1; ++wo) {
        if (wo >= Wo) continue;
        const uint32_t want = cy + static_cast<uint32_t>(w - 2 * wo + 1);
        const int64_t o = (((static_cast<int64_t>(n) * Ho + ho) * Wo + wo) * Cv + cv) * VN;
        uint32_t packed[2];
        if (VN == 8) { const uint2 t = __ldg(reinterpret_cast<const uint2*>(idx + o)); packed[0] = t.x; packed[1] = t.y; }
        else { packed[0] = __ldg(reinterpret_cast<const uint32_t*>(idx + o)); packed[1] = 0u; }
        float g[VN];
        load_vec<T>(gout + o, g);
#pragma unroll
        for (int j = 0; j < VN; ++j)
          if (((packed[j >> 2] >> (8 * (j & 3))) & 255u) == want) acc[j] += g[j];
      }
    }
    store_vec<T>(gx + i * VN, acc);
  }
}

// G (N, 2Ho, 2Wo, C) = gy at the even positions, zero elsewhere.  The data gradient of a stride-2 convolution is the
// stride-1 convolution of G with the flipped, transposed weights (gx[h][w] = sum_{r,s} G[h+1-r][w+1-s] w[r][s]: the terms
// with an odd index vanish), which runs on the fast halo kernels; the gather-producer form of the transposed convolution
// ran at 16-74 TFLOP/s (0.65 ms of the training step for six layers).
template <typename T>
__global__ void zero_insert2x_kernel(const T* __restrict__ gy, int N, int Ho, int Wo, int C, T* __restrict__ out) {
  constexpr int VN = VecOf<T>::N;
  const int Cv = C / VN, W = 2 * Wo, H = 2 * Ho;
  const int64_t total = static_cast<int64_t>(N) * H * W * Cv;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int cv = static_cast<int>(i % Cv);
    int64_t r = i / Cv;
    const int w = static_cast<int>(r % W); r /= W;
    const int h = static_cast<int>(r % H);
    const int n = static_cast<int>(r / H);
    float f[VN];
    if ((h | w) & 1) {
#pragma unroll
      for (int j = 0; j < VN; ++j) f[j] = 0.f;
    } else {
      load_vec<T>(gy + ((static_cast<int64_t>(n) * Ho + (h >> 1)) * Wo + (w >> 1)) * C + cv * VN, f);
    }
    store_vec<T>(out + i * VN, f);
  }
}

// cat([nearest_x2(x_low), skip], C): (N, H/2, W/2, Cx) + (N, H, W, Cs) -> (N, H, W, Cx + Cs); 16-byte vectors
template <typename T>
__global__ void upsample_concat_kernel(const T* __restrict__ xl, const T* __restrict__ skip, int N, int H, int W, int Cx,
                                       int Cs, T* __restrict__ out) {
  constexpr int VN = VecOf<T>::N;
  const int Cv = (Cx + Cs) / VN, Cxv = Cx / VN;
  const int64_t total = static_cast<int64_t>(N) * H * W * Cv;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int cv = static_cast<int>(i % Cv);
    int64_t r = i / Cv;
    const int w = static_cast<int>(r % W); r /= W;
    const int h = static_cast<int>(r % H);
    const int n = static_cast<int>(r / H);
    const uint4* src = cv < Cxv
        ? reinterpret_cast<const uint4*>(xl + ((static_cast<int64_t>(n) * (H >> 1) + (h >> 1)) * (W >> 1) + (w >> 1)) * Cx) + cv
        : reinterpret_cast<const uint4*>(skip + ((static_cast<int64_t>(n) * H + h) * W + w) * Cs) + (cv - Cxv);
    reinterpret_cast<uint4*>(out)[i] = *src;
  }
}

// backward of the above: g_xl = sum over the 2x2 block of g_cat[..., :Cx]; g_skip = g_cat[..., Cx:]
template <typename T>
__global__ void unconcat_bwd_kernel(const T* __restrict__ gcat, int N, int H, int W, int Cx, int Cs, T* __restrict__ gxl,
                                    T* __restrict__ gskip) {
  constexpr int VN = VecOf<T>::N;
  const int C = Cx + Cs, Cxv = Cx / VN, Csv = Cs / VN;
  const int Hl = H >> 1, Wl = W >> 1;
  const int64_t n_low = static_cast<int64_t>(N) * Hl * Wl * Cxv;
  const int64_t n_skip = static_cast<int64_t>(N) * H * W * Csv;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n_low + n_skip;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    if (i < n_low) {
      const int cv = static_cast<int>(i % Cxv);
      int64_t r = i / Cxv;
      const int wl = static_cast<int>(r % Wl); r /= Wl;
      const int hl = static_cast<int>(r % Hl);
      const int n = static_cast<int>(r / Hl);
      float acc[VN];
#pragma unroll
      for (int j = 0; j < VN; ++j) acc[j] = 0.f;
#pragma unroll
      for (int dy = 0; dy < 2; ++dy)
#pragma unroll
        for (int dx = 0; dx < 2; ++dx) {
          float f[VN];
          load_vec<T>(gcat + ((static_cast<int64_t>(n) * H + 2 * hl + dy) * W + 2 * wl + dx) * C + cv * VN, f);
#pragma unroll
          for (int j = 0; j < VN; ++j) acc[j] += f[j];
        }
      store_vec<T>(gxl + i * VN, acc);
    } else {
      const int64_t k = i - n_low;
      const int cv = static_cast<int>(k % Csv);
      const int64_t pix = k / Csv;
      reinterpret_cast<uint4*>(gskip)[k] = *(reinterpret_cast<const uint4*>(gcat + pix * C + Cx) + cv);
    }
  }
}

// (N, K, H, W) fp32 -> (N, H, W, Kp) in T, channels >= K zero
template <typename T>
__global__ void nchw_to_nhwc_kernel(const float* __restrict__ x, int N, int K, int64_t HW, int Kp, T* __restrict__ out) {
  const int64_t total = static_cast<int64_t>(N) * HW * Kp;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int k = static_cast<int>(i % Kp);
    const int64_t r = i / Kp;
    const int64_t px = r % HW, n = r / HW;
    out[i] = from_f<T>(k < K ? x[(n * K + k) * HW + px] : 0.f);
  }
}

// Kp == 16 bf16 (the logits gradient entering the tensor-core head backward): one thread = one pixel, K coalesced plane
// reads and one 32-byte store (the scalar form above wrote 2 bytes per thread: 244 us for 64 tiles of 256 x 256)
__global__ void nchw_to_nhwc16_bf16_kernel(const float* __restrict__ x, int N, int K, int64_t HW, __nv_bfloat16* __restrict__ out) {
  const int64_t total = static_cast<int64_t>(N) * HW;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t n = i / HW, px = i - n * HW;
    float f[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) f[k] = k < K ? __ldg(x + (n * K + k) * HW + px) : 0.f;
    store_bf16x16(out + i * 16, f);
  }
}

// ---- weight repacking (fp32 OIHW master -> kernel layouts) ----------------------------------------------------
// mode 0: fp32 [tap][C_in_p][C_out]                                   (direct forward; C_in_p >= C_in, zero padded)
// mode 1: bf16 [C_out][Kpad], k = tap * C_in + ci                      (tcgen05 forward)
// mode 2: bf16 [C_out][256],  k = r * 32 + s * 4 + ci                  (tcgen05 stem, C_in <= 4, 7x7)
// mode 3: bf16 [C_in][Kpad],  k = (RS - 1 - tap) * Cop + co            (dgrad of a stride-1 conv run as a forward conv)
// mode 4: bf16 [C_in][Kpad],  k = tap * Cop + co                       (dgrad of a stride-2 conv, DT_CONV_TRANSPOSED)
__device__ __forceinline__ void pack_weight_element(const float* __restrict__ w, int C_out, int C_in, int R, int S, int mode,
                                                    int C_in_p, int Kpad, void* __restrict__ out, int64_t i) {
  const int RS = R * S;
  if (mode == 0) {
    const int co = static_cast<int>(i % C_out);
    const int ci = static_cast<int>((i / C_out) % C_in_p);
    const int tap = static_cast<int>(i / (static_cast<int64_t>(C_out) * C_in_p));
    static_cast<float*>(out)[i] = ci < C_in ? w[(static_cast<int64_t>(co) * C_in + ci) * RS + tap] : 0.f;
  } else if (mode == 1) {
    const int k = static_cast<int>(i % Kpad), co = static_cast<int>(i / Kpad);
    float v = 0.f;
    if (k < RS * C_in) { const int tap = k / C_in, ci = k % C_in; v = w[(static_cast<int64_t>(co) * C_in + ci) * RS + tap]; }
    static_cast<__nv_bfloat16*>(out)[i] = __float2bfloat16_rn(v);
  } else if (mode == 2) {
    const int k = static_cast<int>(i % Kpad), co = static_cast<int>(i / Kpad);
    const int r = k >> 5, s = (k & 31) >> 2, ci = k & 3;
    float v = 0.f;
    if (r < R && s < S && ci < C_in) v = w[(static_cast<int64_t>(co) * C_in + ci) * RS + r * S + s];
    static_cast<__nv_bfloat16*>(out)[i] = __float2bfloat16_rn(v);
  } else {   // modes 3 / 4: C_in_p = channel stride Cop of the gradient tensor (>= C_out)
    const int k = static_cast<int>(i % Kpad), ci = static_cast<int>(i / Kpad);
    float v = 0.f;
    if (k < RS * C_in_p) {
      const int tf = k / C_in_p, co = k % C_in_p;
      if (co < C_out) v = w[(static_cast<int64_t>(co) * C_in + ci) * RS + (mode == 3 ? RS - 1 - tf : tf)];
    }
    static_cast<__nv_bfloat16*>(out)[i] = __float2bfloat16_rn(v);
  }
}

__global__ void pack_weight_kernel(const float* __restrict__ w, int C_out, int C_in, int R, int S, int mode, int C_in_p,
                                   int Kpad, void* __restrict__ out, int64_t total) {
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x)
    pack_weight_element(w, C_out, C_in, R, S, mode, C_in_p, Kpad, out, i);
}

// every layout of every layer in ONE launch (the training step repacks ~93 weight tensors after each optimizer step):
// jobs[j] covers the output elements [start_j, start_{j+1}).  Work unit = one output ROW (Kpad elements: a C_out row of the
// forward layout, a C_in row of the dgrad layouts; 256-element chunks for the fp32 / stem layouts).  A block looks its row
// up once (binary search over the per-job row prefix in shared memory), stages the fp32 source of the row in shared memory -
// the forward row is ONE contiguous run of C_in * RS floats, a dgrad row is C_out runs of RS floats - and writes the row
// with consecutive threads on consecutive elements; shared-memory reads have stride RS (odd: conflict-free).  Against the
// thread-per-element form this removes the per-element search and 64-bit divisions and the 4-byte gathers.
constexpr int kPackStage = 8192;                         // floats of staging per block (32 KB): rows up to C * RS = 8192

__device__ __forceinline__ int pack_row_len(const dt_pack_job& jb) { return (jb.mode == 0 || jb.mode == 2) ? 256 : jb.Kpad; }

__global__ void pack_weights_batched_kernel(const dt_pack_job* __restrict__ jobs, int njobs, int64_t total) {
  extern __shared__ int64_t pk_sh[];                    // [njobs + 1] row prefix, then the staging floats
  int64_t* row_start = pk_sh;
  float* stage = reinterpret_cast<float*>(pk_sh + njobs + 1);
  if (threadIdx.x == 0) {
    int64_t rows = 0;
    for (int j = 0; j < njobs; ++j) {
      row_start[j] = rows;
      const int64_t count = (j + 1 < njobs ? jobs[j + 1].start : total) - jobs[j].start;
      const int len = pack_row_len(jobs[j]);
      rows += (count + len - 1) / len;
    }
    row_start[njobs] = rows;
  }
  __syncthreads();
  const int64_t rows = row_start[njobs];
  for (int64_t row = blockIdx.x; row < rows; row += gridDim.x) {
    int lo = 0, hi = njobs - 1;
    while (lo < hi) {
      const int mid = (lo + hi + 1) >> 1;
      if (row_start[mid] <= row) lo = mid; else hi = mid - 1;
    }
    const dt_pack_job jb = jobs[lo];
    const int r = static_cast<int>(row - row_start[lo]);
    const int RS = jb.R * jb.S;
    if (jb.mode == 0 || jb.mode == 2) {                  // small / check-mode layouts: element form on a 256-element chunk
      const int64_t count = (lo + 1 < njobs ? jobs[lo + 1].start : total) - jb.start;
      const int64_t i = static_cast<int64_t>(r) * 256 + threadIdx.x;
      if (threadIdx.x < 256 && i < count) pack_weight_element(jb.w, jb.C_out, jb.C_in, jb.R, jb.S, jb.mode, jb.C_in_p, jb.Kpad, jb.out, i);
      continue;
    }
    __nv_bfloat16* out = static_cast<__nv_bfloat16*>(jb.out) + static_cast<int64_t>(r) * jb.Kpad;
    const int X = jb.mode == 1 ? jb.C_in : jb.C_out;     // source entries of this row: X runs of RS floats
    const bool staged = X * RS <= kPackStage;
    if (staged) {
      __syncthreads();                                   // the previous row's readers are done with the staging buffer
      if (jb.mode == 1) {
        const float* src = jb.w + static_cast<int64_t>(r) * jb.C_in * RS;          // contiguous
        for (int e = threadIdx.x; e < X * RS; e += blockDim.x) stage[e] = src[e];
      } else {
        const float* src = jb.w + static_cast<int64_t>(r) * RS;                    // + co * C_in * RS + tap
        const int64_t cstride = static_cast<int64_t>(jb.C_in) * RS;
        for (int e = threadIdx.x; e < X * RS; e += blockDim.x) stage[e] = src[(e / RS) * cstride + e % RS];
      }
      __syncthreads();
    }
    if (jb.mode == 1) {
      for (int k = threadIdx.x; k < jb.Kpad; k += blockDim.x) {
        float v = 0.f;
        if (k < RS * jb.C_in) {
          const int tap = k / jb.C_in, ci = k - tap * jb.C_in;
          v = staged ? stage[ci * RS + tap] : jb.w[(static_cast<int64_t>(r) * jb.C_in + ci) * RS + tap];
        }
        out[k] = __float2bfloat16_rn(v);
      }
    } else {                                             // modes 3 / 4: row = input channel r, k = tf * Cop + co
      const int Cop = jb.C_in_p;
      for (int k = threadIdx.x; k < jb.Kpad; k += blockDim.x) {
        float v = 0.f;
        if (k < RS * Cop) {
          const int tf = k / Cop, co = k - tf * Cop;
          if (co < jb.C_out) {
            const int tap = jb.mode == 3 ? RS - 1 - tf : tf;
            v = staged ? stage[co * RS + tap] : jb.w[(static_cast<int64_t>(co) * jb.C_in + r) * RS + tap];
          }
        }
        out[k] = __float2bfloat16_rn(v);
      }
    }
  }
}

// ---- generic direct dgrad / wgrad (CUDA cores) ------------------------------------------------------------------
struct BwdGeo {
  int N, H, W, C_in, C_x;   // conv input (N, H, W, C_x) storage, C_in <= C_x real channels
  int C_out, R, S, stride, pad, Ho, Wo;
};

// gx[n,h,w,ci] = addend + sum_{r,s,co} gy[n,(h+pad-r)/st,(w+pad-s)/st,co] * w[co,ci,r,s]   (w: fp32 OIHW master)
template <typename T>
__global__ void dgrad_direct_kernel(BwdGeo p, const T* __restrict__ gy, const float* __restrict__ w,
                                    const T* __restrict__ addend, T* __restrict__ gx, int round_w) {
  const int64_t total = static_cast<int64_t>(p.N) * p.H * p.W * p.C_x;
  const int RS = p.R * p.S;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int ci = static_cast<int>(i % p.C_x);
    int64_t r = i / p.C_x;
    const int wx = static_cast<int>(r % p.W); r /= p.W;
    const int hx = static_cast<int>(r % p.H);
    const int n = static_cast<int>(r / p.H);
    float acc = addend ? to_f<T>(addend[i]) : 0.f;
    if (ci < p.C_in) {
      for (int fr = 0; fr < p.R; ++fr) {
        const int hn = hx + p.pad - fr;
        if (hn < 0 || hn % p.stride) continue;
        const int ho = hn / p.stride;
        if (ho >= p.Ho) continue;
        for (int fs = 0; fs < p.S; ++fs) {
          const int wn = wx + p.pad - fs;
          if (wn < 0 || wn % p.stride) continue;
          const int wo = wn / p.stride;
          if (wo >= p.Wo) continue;
          const T* gp = gy + ((static_cast<int64_t>(n) * p.Ho + ho) * p.Wo + wo) * p.C_out;
          const float* wp = w + static_cast<int64_t>(ci) * RS + fr * p.S + fs;
          // bf16 mode: weights rounded to bf16 like the tensor-core operands (fp32 accumulation either way)
          for (int co = 0; co < p.C_out; ++co) {
            const float wv = wp[static_cast<int64_t>(co) * p.C_in * RS];
            acc = fmaf(to_f<T>(gp[co]), round_w ? __bfloat162float(__float2bfloat16_rn(wv)) : wv, acc);
          }
        }
      }
    }
    gx[i] = from_f<T>(acc);
  }
}

// dw[co,ci,r,s] += sum_{n,ho,wo} gy[n,ho,wo,co] * x[n,ho*st+r-pad,wo*st+s-pad,ci]   (dw: fp32 OIHW, caller zeroes)
// grid = (pixel chunks, co tiles * ci tiles, taps); block tile 32 co x 32 ci, thread = 1 co x 4 ci; the pixel chunk is
// staged through shared memory 32 output pixels at a time.
template <typename T>
__global__ void wgrad_direct_kernel(BwdGeo p, const T* __restrict__ x, const T* __restrict__ gy, float* __restrict__ dw,
                                    int ci_tiles, int64_t M) {
  constexpr int TP = 32;
  __shared__ float sg[TP][33];
  __shared__ float sx[TP][33];
  const int tap = blockIdx.z, fr = tap / p.S, fs = tap % p.S;
  const int co0 = (blockIdx.y / ci_tiles) * 32, ci0 = (blockIdx.y % ci_tiles) * 32;
  const int col = threadIdx.x >> 3, cig = (threadIdx.x & 7) * 4;   // compute role
  const int lp = threadIdx.x >> 3, lc = (threadIdx.x & 7) * 4;     // load role: pixel lp, channels lc..lc+3
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  const int64_t per = (M + gridDim.x - 1) / gridDim.x;
  const int64_t m0 = blockIdx.x * per, m1 = min(M, m0 + per);
  for (int64_t mb = m0; mb < m1; mb += TP) {
    {
      const int64_t m = mb + lp;
      float g4[4] = {0.f, 0.f, 0.f, 0.f}, x4[4] = {0.f, 0.f, 0.f, 0.f};
      if (m < m1) {
        const int wo = static_cast<int>(m % p.Wo);
        const int64_t t = m / p.Wo;
        const int ho = static_cast<int>(t % p.Ho);
        const int n = static_cast<int>(t / p.Ho);
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (co0 + lc + j < p.C_out) g4[j] = to_f<T>(gy[m * p.C_out + co0 + lc + j]);
        const int hi = ho * p.stride + fr - p.pad, wi = wo * p.stride + fs - p.pad;
        if (hi >= 0 && hi < p.H && wi >= 0 && wi < p.W) {
          const T* xp = x + ((static_cast<int64_t>(n) * p.H + hi) * p.W + wi) * p.C_x;
#pragma unroll
          for (int j = 0; j < 4; ++j)
            if (ci0 + lc + j < p.C_in) x4[j] = to_f<T>(xp[ci0 + lc + j]);
        }
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) { sg[lp][lc + j] = g4[j]; sx[lp][lc + j] = x4[j]; }
    }
    __syncthreads();
#pragma unroll 8
    for (int q = 0; q < TP; ++q) {
      const float g = sg[q][col];
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[j] = fmaf(g, sx[q][cig + j], acc[j]);
    }
    __syncthreads();
  }
  const int co = co0 + col;
  if (co < p.C_out) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int ci = ci0 + cig + j;
      if (ci < p.C_in) atomicAdd(dw + (static_cast<int64_t>(co) * p.C_in + ci) * (p.R * p.S) + tap, acc[j]);
    }
  }
}

// db[k] += sum over pixels of g[pix][k]  (bias gradient of the head; C small).  Float atomics: the sum depends on the order
// in which the warps arrive (check-mode wgrad only; the training step uses the two-stage form below).
template <typename T>
__global__ void bias_grad_kernel(const T* __restrict__ g, int64_t M, int C, int K, float* __restrict__ db) {
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  for (int64_t m = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; m < M;
       m += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    for (int k = 0; k < K; ++k) acc[k] += to_f<T>(g[m * C + k]);
  }
  for (int k = 0; k < K; ++k) {
    const float t = warp_sum(acc[k]);
    if ((threadIdx.x & 31) == 0) atomicAdd(db + k, t);
  }
}

// The same sum, bit-reproducible from run to run: stage 1 leaves one partial per (block, channel) - summed over the block in
// a fixed order -, stage 2 adds the partials in block order.  (The atomic form made the head's bias gradient, through the
// gradient norm of the clip, the one source of run-to-run differences of a training step.)
template <typename T>
__global__ void channel_sum_partials_kernel(const T* __restrict__ g, int64_t M, int C, int K, float* __restrict__ part) {
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  for (int64_t m = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; m < M;
       m += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    for (int k = 0; k < K; ++k) acc[k] += to_f<T>(g[m * C + k]);
  }
  __shared__ float sh[kThreads / 32][4];
  for (int k = 0; k < 4; ++k) {
    const float t = warp_sum(acc[k]);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5][k] = t;
  }
  __syncthreads();
  if (threadIdx.x < 4) {
    float t = 0.f;
    for (int w = 0; w < kThreads / 32; ++w) t += sh[w][threadIdx.x];
    part[blockIdx.x * 4 + threadIdx.x] = t;
  }
}

__global__ void channel_sum_final_kernel(const float* __restrict__ part, int blocks, int K, float* __restrict__ out) {
  const int k = threadIdx.x;
  if (k >= K) return;
  double t = 0.0;
  for (int b = 0; b < blocks; ++b) t += static_cast<double>(part[b * 4 + k]);
  out[k] = static_cast<float>(t);
}

template <typename T, int MODE>
void launch_channel_reduce(int nb, cudaStream_t s, const T* y, const T* g, const T* a, const float* mean, const float* invstd,
                           const float* fscale, const float* fshift, int64_t M, int C, float* part) {
  constexpr int U = MODE == 0 ? 4 : 2;
  if (reduce_unrolled())
    channel_reduce_kernel<T, MODE, U><<<nb, kThreads, 0, s>>>(y, g, a, mean, invstd, fscale, fshift, M, C, part);
  else
    channel_reduce_kernel<T, MODE, 1><<<nb, kThreads, 0, s>>>(y, g, a, mean, invstd, fscale, fshift, M, C, part);
}

int reduce_blocks(int64_t M, int C, int vn) {
  const int lanes = kThreads / (C / vn);
  int64_t b = (M + lanes * 16 - 1) / (static_cast<int64_t>(lanes) * 16);
  const int64_t cap = static_cast<int64_t>(dt_num_sms()) * 4;
  if (b > cap) b = cap;
  return static_cast<int>(b < 1 ? 1 : b);
}

bool vec_ok(int C, int dtype) {
  const int vn = dtype == DT_BF16 ? 8 : 4;
  return C % vn == 0 && C / vn <= kThreads && kThreads % (C / vn) == 0;
}

}  // namespace

#define DT_DTYPE_SWITCH(dtype, CALL_F32, CALL_BF16) \
  do { if ((dtype) == DT_F32) { CALL_F32; } else { CALL_BF16; } } while (0)

extern "C" {

int dt_reduce_blocks(int64_t M, int C, int dtype) {
  if (M <= 0 || !vec_ok(C, dtype)) return DT_ERR_BAD_SHAPE;
  return reduce_blocks(M, C, dtype == DT_BF16 ? 8 : 4);
}

int dt_bn_train_stats(const void* y, int64_t M, int C, int dtype, const float* gamma, const float* beta, float eps,
                      float momentum, float* running_mean, float* running_var, float* scale, float* shift, float* mean,
                      float* invstd, float* workspace, dt_stream_t stream) {
  DT_ARCH_GUARD();
  DT_REQUIRE(M > 0 && (dtype == DT_F32 || dtype == DT_BF16) && vec_ok(C, dtype), DT_ERR_BAD_SHAPE,
             "dt_bn_train_stats: M=%lld C=%d dtype=%d", static_cast<long long>(M), C, dtype);
  DT_REQUIRE(reinterpret_cast<uintptr_t>(y) % 16 == 0, DT_ERR_BAD_ALIGN, "dt_bn_train_stats: y must be 16-byte aligned");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int nb = reduce_blocks(M, C, dtype == DT_BF16 ? 8 : 4);
  DT_DTYPE_SWITCH(dtype,
      (launch_channel_reduce<float, 0>(nb, s, static_cast<const float*>(y), nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, M, C, workspace)),
      (launch_channel_reduce<__nv_bfloat16, 0>(nb, s, static_cast<const __nv_bfloat16*>(y), nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, M, C, workspace)));
  DT_LAUNCH_CHECK();
  bn_finalize_kernel<<<C, kFinThreads, 0, s>>>(workspace, nb, C, static_cast<double>(M), gamma, beta, eps, momentum,
                                                     running_mean, running_var, scale, shift, mean, invstd);
  DT_LAUNCH_CHECK();
  return DT_OK;
}

int dt_bn_apply(const void* y, int64_t M, int C, int dtype, const float* scale, const float* shift, const void* residual,
                int relu, void* out, dt_stream_t stream) {
  DT_ARCH_GUARD();
  DT_REQUIRE(M > 0 && (dtype == DT_F32 || dtype == DT_BF16) && vec_ok(C, dtype), DT_ERR_BAD_SHAPE,
             "dt_bn_apply: M=%lld C=%d dtype=%d", static_cast<long long>(M), C, dtype);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int64_t nvec = M * C / (dtype == DT_BF16 ? 8 : 4);
  DT_DTYPE_SWITCH(dtype,
      (bn_apply_kernel<float><<<grid_for(nvec), kThreads, 0, s>>>(static_cast<const float*>(y), scale, shift, static_cast<const float*>(residual), relu, nvec, C, static_cast<float*>(out))),
      (bn_apply_kernel<__nv_bfloat16><<<grid_for(nvec), kThreads, 0, s>>>(static_cast<const __nv_bfloat16*>(y), scale, shift, static_cast<const __nv_bfloat16*>(residual), relu, nvec, C, static_cast<__nv_bfloat16*>(out))));
  DT_LAUNCH_CHECK();
  return DT_OK;
}

static int bn_train_bwd_impl(const void* g, const void* a, const void* y, int64_t M, int C, int dtype, const float* mean,
                             const float* invstd, const float* scale, const float* fshift, float* dgamma, float* dbeta,
                             void* gy, void* gz_out, float* workspace, dt_stream_t stream) {
  DT_ARCH_GUARD();
  DT_REQUIRE(M > 0 && (dtype == DT_F32 || dtype == DT_BF16) && vec_ok(C, dtype), DT_ERR_BAD_SHAPE,
             "dt_bn_train_bwd: M=%lld C=%d dtype=%d", static_cast<long long>(M), C, dtype);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int vn = dtype == DT_BF16 ? 8 : 4;
  const int nb = reduce_blocks(M, C, vn);
  float* c1 = workspace + static_cast<int64_t>(2) * nb * C;
  float* c2 = c1 + C;
  const int64_t nvec = M * C / vn;
  DT_DTYPE_SWITCH(dtype,
      (launch_channel_reduce<float, 1>(nb, s, static_cast<const float*>(y), static_cast<const float*>(g), static_cast<const float*>(a), mean, invstd, scale, fshift, M, C, workspace)),
      (launch_channel_reduce<__nv_bfloat16, 1>(nb, s, static_cast<const __nv_bfloat16*>(y), static_cast<const __nv_bfloat16*>(g), static_cast<const __nv_bfloat16*>(a), mean, invstd, scale, fshift, M, C, workspace)));
  DT_LAUNCH_CHECK();
  bn_bwd_finalize_kernel<<<C, kFinThreads, 0, s>>>(workspace, nb, C, static_cast<double>(M), dgamma, dbeta, c1, c2);
  DT_LAUNCH_CHECK();
  DT_DTYPE_SWITCH(dtype,
      (bn_bwd_apply_kernel<float><<<grid_for(nvec), kThreads, 0, s>>>(static_cast<const float*>(g), static_cast<const float*>(a), static_cast<const float*>(y), mean, invstd, scale, c1, c2, fshift, nvec, C, static_cast<float*>(gy), static_cast<float*>(gz_out))),
      (bn_bwd_apply_kernel<__nv_bfloat16><<<grid_for(nvec), kThreads, 0, s>>>(static_cast<const __nv_bfloat16*>(g), static_cast<const __nv_bfloat16*>(a), static_cast<const __nv_bfloat16*>(y), mean, invstd, scale, c1, c2, fshift, nvec, C, static_cast<__nv_bfloat16*>(gy), static_cast<__nv_bfloat16*>(gz_out))));
  DT_LAUNCH_CHECK();
  return DT_OK;
}

int dt_bn_train_bwd(const void* g, const void* a, const void* y, int64_t M, int C, int dtype, const float* mean,
                    const float* invstd, const float* scale, float* dgamma, float* dbeta, void* gy, void* gz_out,
                    float* workspace, dt_stream_t stream) {
  return bn_train_bwd_impl(g, a, y, M, C, dtype, mean, invstd, scale, nullptr, dgamma, dbeta, gy, gz_out, workspace, stream);
}

int dt_bn_train_bwd_relu(const void* g, const void* y, int64_t M, int C, int dtype, const float* mean, const float* invstd,
                         const float* scale, const float* shift, float* dgamma, float* dbeta, void* gy, void* gz_out,
                         float* workspace, dt_stream_t stream) {
  DT_REQUIRE(shift != nullptr, DT_ERR_BAD_SHAPE, "dt_bn_train_bwd_relu: the forward shift vector is required");
  return bn_train_bwd_impl(g, nullptr, y, M, C, dtype, mean, invstd, scale, shift, dgamma, dbeta, gy, gz_out, workspace, stream);
}

int dt_add(const void* a, const void* b, int64_t n, int dtype, void* out, dt_stream_t stream) {
  DT_ARCH_GUARD();
  const int vn = dtype == DT_BF16 ? 8 : 4;
  DT_REQUIRE(n > 0 && n % vn == 0 && (dtype == DT_F32 || dtype == DT_BF16), DT_ERR_BAD_SHAPE, "dt_add: n=%lld", static_cast<long long>(n));
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int64_t nvec = n / vn;
  DT_DTYPE_SWITCH(dtype,
      (add_kernel<float><<<grid_for(nvec), kThreads, 0, s>>>(static_cast<const float*>(a), static_cast<const float*>(b), nvec, static_cast<float*>(out))),
      (add_kernel<__nv_bfloat16><<<grid_for(nvec), kThreads, 0, s>>>(static_cast<const __nv_bfloat16*>(a), static_cast<const __nv_bfloat16*>(b), nvec, static_cast<__nv_bfloat16*>(out))));
  DT_LAUNCH_CHECK();
  return DT_OK;
}

int dt_maxpool3x3s2_bwd(const void* x, const void* gout, const void* addend, int N, int H, int W, int C, int dtype,
                        void* gx, dt_stream_t stream) {
  DT_ARCH_GUARD();
  DT_REQUIRE(N > 0 && H > 0 && W > 0 && C > 0 && (dtype == DT_F32 || dtype == DT_BF16), DT_ERR_BAD_SHAPE,
             "dt_maxpool3x3s2_bwd: bad shape");
  const int Ho = (H - 1) / 2 + 1, Wo = (W - 1) / 2 + 1;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int64_t total = static_cast<int64_t>(N) * H * W * C;
  DT_DTYPE_SWITCH(dtype,
      (maxpool_bwd_kernel<float><<<grid_for(total, 16), kThreads, 0, s>>>(static_cast<const float*>(x), static_cast<const float*>(gout), static_cast<const float*>(addend), N, H, W, C, Ho, Wo, static_cast<float*>(gx))),
      (maxpool_bwd_kernel<__nv_bfloat16><<<grid_for(total, 16), kThreads, 0, s>>>(static_cast<const __nv_bfloat16*>(x), static_cast<const __nv_bfloat16*>(gout), static_cast<const __nv_bfloat16*>(addend), N, H, W, C, Ho, Wo, static_cast<__nv_bfloat16*>(gx))));
  DT_LAUNCH_CHECK();
  return DT_OK;
}

int dt_maxpool3x3s2_idx(const void* x, int N, int H, int W, int C, int dtype, void* y, uint8_t* idx, dt_stream_t stream) {
  DT_ARCH_GUARD();
  const int vn = dtype == DT_BF16 ? 8 : 4;
  DT_REQUIRE(N > 0 && H > 0 && W > 0 && C > 0 && C % vn == 0 && (dtype == DT_F32 || dtype == DT_BF16), DT_ERR_BAD_SHAPE,
             "dt_maxpool3x3s2_idx: bad shape (C=%d must be a multiple of %d)", C, vn);
  DT_REQUIRE((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y) | reinterpret_cast<uintptr_t>(idx)) % 16 == 0,
             DT_ERR_BAD_ALIGN, "dt_maxpool3x3s2_idx: tensors must be 16-byte aligned");
  const int Ho = (H - 1) / 2 + 1, Wo = (W - 1) / 2 + 1;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int64_t total = static_cast<int64_t>(N) * Ho * Wo * (C / vn);
  DT_DTYPE_SWITCH(dtype,
      (maxpool_idx_fwd_kernel<float><<<grid_for(total, 16), kThreads, 0, s>>>(static_cast<const float*>(x), N, H, W, C, Ho, Wo, static_cast<float*>(y), idx)),
      (maxpool_idx_fwd_kernel<__nv_bfloat16><<<grid_for(total, 16), kThreads, 0, s>>>(static_cast<const __nv_bfloat16*>(x), N, H, W, C, Ho, Wo, static_cast<__nv_bfloat16*>(y), idx)));
  DT_LAUNCH_CHECK();
  return DT_OK;
}

int dt_maxpool3x3s2_bwd_idx(const uint8_t* idx, const void* gout, const void* addend, int N, int H, int W, int C, int dtype,
                            void* gx, dt_stream_t stream) {
  DT_ARCH_GUARD();
  const int vn = dtype == DT_BF16 ? 8 : 4;
  DT_REQUIRE(N > 0 && H > 0 && W > 0 && C > 0 && C % vn == 0 && (dtype == DT_F32 || dtype == DT_BF16), DT_ERR_BAD_SHAPE,
             "dt_maxpool3x3s2_bwd_idx: bad shape (C=%d must be a multiple of %d)", C, vn);
  DT_REQUIRE((reinterpret_cast<uintptr_t>(idx) | reinterpret_cast<uintptr_t>(gout) | reinterpret_cast<uintptr_t>(addend) |
              reinterpret_cast<uintptr_t>(gx)) % 16 == 0, DT_ERR_BAD_ALIGN, "dt_maxpool3x3s2_bwd_idx: tensors must be 16-byte aligned");
  const int Ho = (H - 1) / 2 + 1, Wo = (W - 1) / 2 + 1;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int64_t total = static_cast<int64_t>(N) * H * W * (C / vn);
  DT_DTYPE_SWITCH(dtype,
      (maxpool_idx_bwd_kernel<float><<<grid_for(total, 16), kThreads, 0, s>>>(idx, static_cast<const float*>(gout), static_cast<const float*>(addend), N, H, W, C, Ho, Wo, static_cast<float*>(gx))),
      (maxpool_idx_bwd_kernel<__nv_bfloat16><<<grid_for(total, 16), kThreads, 0, s>>>(idx, static_cast<const __nv_bfloat16*>(gout), static_cast<const __nv_bfloat16*>(addend), N, H, W, C, Ho, Wo, static_cast<__nv_bfloat16*>(gx))));
  DT_LAUNCH_CHECK();
  return DT_OK;
}

int dt_zero_insert2x(const void* gy, int N, int Ho, int Wo, int C, int dtype, void* out, dt_stream_t stream) {
  DT_ARCH_GUARD();
  const int vn = dtype == DT_BF16 ? 8 : 4;
  DT_REQUIRE(N > 0 && Ho > 0 && Wo > 0 && C > 0 && C % vn == 0 && (dtype == DT_F32 || dtype == DT_BF16), DT_ERR_BAD_SHAPE,
             "dt_zero_insert2x: bad shape (C=%d must be a multiple of %d)", C, vn);
  DT_REQUIRE((reinterpret_cast<uintptr_t>(gy) | reinterpret_cast<uintptr_t>(out)) % 16 == 0, DT_ERR_BAD_ALIGN,
             "dt_zero_insert2x: tensors must be 16-byte aligned");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int64_t total = static_cast<int64_t>(N) * 4 * Ho * Wo * (C / vn);
  DT_DTYPE_SWITCH(dtype,
      (zero_insert2x_kernel<float><<<grid_for(total, 16), kThreads, 0, s>>>(static_cast<const float*>(gy), N, Ho, Wo, C, static_cast<float*>(out))),
      (zero_insert2x_kernel<__nv_bfloat16><<<grid_for(total, 16), kThreads, 0, s>>>(static_cast<const __nv_bfloat16*>(gy), N, Ho, Wo, C, static_cast<__nv_bfloat16*>(out))));
  DT_LAUNCH_CHECK();
  return DT_OK;
}

int dt_upsample_concat(const void* x_low, const void* skip, int N, int H, int W, int Cx, int Cs, int dtype, void* out,
                       dt_stream_t stream) {
  DT_ARCH_GUARD();
  const int vn = dtype == DT_BF16 ? 8 : 4;
  DT_REQUIRE(N > 0 && H > 0 && W > 0 && H % 2 == 0 && W % 2 == 0 && Cx > 0 && Cx % vn == 0 && Cs >= 0 && Cs % vn == 0 &&
                 (Cs == 0 || skip != nullptr) && (dtype == DT_F32 || dtype == DT_BF16),
             DT_ERR_BAD_SHAPE, "dt_upsample_concat: bad shape (H=%d W=%d Cx=%d Cs=%d)", H, W, Cx, Cs);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int64_t total = static_cast<int64_t>(N) * H * W * ((Cx + Cs) / vn);
  DT_DTYPE_SWITCH(dtype,
      (upsample_concat_kernel<float><<<grid_for(total, 16), kThreads, 0, s>>>(static_cast<const float*>(x_low), static_cast<const float*>(skip), N, H, W, Cx, Cs, static_cast<float*>(out))),
      (upsample_concat_kernel<__nv_bfloat16><<<grid_for(total, 16), kThreads, 0, s>>>(static_cast<const __nv_bfloat16*>(x_low), static_cast<const __nv_bfloat16*>(skip), N, H, W, Cx, Cs, static_cast<__nv_bfloat16*>(out))));
  DT_LAUNCH_CHECK();
  return DT_OK;
}

int dt_upsample_concat_bwd(const void* g_cat, int N, int H, int W, int Cx, int Cs, int dtype, void* g_x_low, void* g_skip,
                           dt_stream_t stream) {
  DT_ARCH_GUARD();
  const int vn = dtype == DT_BF16 ? 8 : 4;
  DT_REQUIRE(N > 0 && H > 0 && W > 0 && H % 2 == 0 && W % 2 == 0 && Cx > 0 && Cx % vn == 0 && Cs >= 0 && Cs % vn == 0 &&
                 (Cs == 0 || g_skip != nullptr) && (dtype == DT_F32 || dtype == DT_BF16),
             DT_ERR_BAD_SHAPE, "dt_upsample_concat_bwd: bad shape (H=%d W=%d Cx=%d Cs=%d)", H, W, Cx, Cs);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int64_t total = static_cast<int64_t>(N) * (H / 2) * (W / 2) * (Cx / vn) + static_cast<int64_t>(N) * H * W * (Cs / vn);
  DT_DTYPE_SWITCH(dtype,
      (unconcat_bwd_kernel<float><<<grid_for(total, 16), kThreads, 0, s>>>(static_cast<const float*>(g_cat), N, H, W, Cx, Cs, static_cast<float*>(g_x_low), static_cast<float*>(g_skip))),
      (unconcat_bwd_kernel<__nv_bfloat16><<<grid_for(total, 16), kThreads, 0, s>>>(static_cast<const __nv_bfloat16*>(g_cat), N, H, W, Cx, Cs, static_cast<__nv_bfloat16*>(g_x_low), static_cast<__nv_bfloat16*>(g_skip))));
  DT_LAUNCH_CHECK();
  return DT_OK;
}

int dt_nchw_to_nhwc(const float* x, int N, int K, int H, int W, int Kp, int dtype, void* out, dt_stream_t stream) {
  DT_ARCH_GUARD();
  DT_REQUIRE(N > 0 && K > 0 && Kp >= K && H > 0 && W > 0 && (dtype == DT_F32 || dtype == DT_BF16), DT_ERR_BAD_SHAPE,
             "dt_nchw_to_nhwc: bad shape");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int64_t HW = static_cast<int64_t>(H) * W, total = N * HW * Kp;
  if (dtype == DT_BF16 && Kp == 16 && reinterpret_cast<uintptr_t>(out) % 32 == 0) {
    nchw_to_nhwc16_bf16_kernel<<<grid_for(N * HW, 16), kThreads, 0, s>>>(x, N, K, HW, static_cast<__nv_bfloat16*>(out));
    DT_LAUNCH_CHECK();
    return DT_OK;
  }
  DT_DTYPE_SWITCH(dtype,
      (nchw_to_nhwc_kernel<float><<<grid_for(total, 16), kThreads, 0, s>>>(x, N, K, HW, Kp, static_cast<float*>(out))),
      (nchw_to_nhwc_kernel<__nv_bfloat16><<<grid_for(total, 16), kThreads, 0, s>>>(x, N, K, HW, Kp, static_cast<__nv_bfloat16*>(out))));
  DT_LAUNCH_CHECK();
  return DT_OK;
}

int dt_channel_sum(const void* g, int64_t M, int C, int K, int dtype, float* workspace, float* out, dt_stream_t stream) {
  DT_ARCH_GUARD();
  DT_REQUIRE(M > 0 && C > 0 && K >= 1 && K <= 4 && K <= C && (dtype == DT_F32 || dtype == DT_BF16), DT_ERR_BAD_SHAPE,
             "dt_channel_sum: M=%lld C=%d K=%d (K <= 4)", static_cast<long long>(M), C, K);
  DT_REQUIRE(workspace != nullptr && out != nullptr, DT_ERR_BAD_SHAPE, "dt_channel_sum: null pointer");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  int blocks = grid_for(M, 2);
  if (blocks > 512) blocks = 512;                       // workspace: 2048 floats
  DT_DTYPE_SWITCH(dtype,
      (channel_sum_partials_kernel<float><<<blocks, kThreads, 0, s>>>(static_cast<const float*>(g), M, C, K, workspace)),
      (channel_sum_partials_kernel<__nv_bfloat16><<<blocks, kThreads, 0, s>>>(static_cast<const __nv_bfloat16*>(g), M, C, K, workspace)));
  DT_LAUNCH_CHECK();
  channel_sum_final_kernel<<<1, 32, 0, s>>>(workspace, blocks, K, out);
  DT_LAUNCH_CHECK();
  return DT_OK;
}

int dt_pack_conv_weight(const float* w_oihw, int C_out, int C_in, int R, int S, int mode, int C_in_p, int Kpad, void* out,
                        dt_stream_t stream) {
  DT_ARCH_GUARD();
  DT_REQUIRE(C_out > 0 && C_in > 0 && R > 0 && S > 0 && mode >= 0 && mode <= 4, DT_ERR_BAD_SHAPE, "dt_pack_conv_weight: bad arguments");
  int64_t total = 0;
  if (mode == 0) {
    DT_REQUIRE(C_in_p >= C_in, DT_ERR_BAD_SHAPE, "dt_pack_conv_weight: C_in_p < C_in");
    total = static_cast<int64_t>(R) * S * C_in_p * C_out;
  } else if (mode == 1) {
    DT_REQUIRE(Kpad >= R * S * C_in, DT_ERR_BAD_SHAPE, "dt_pack_conv_weight: Kpad too small");
    total = static_cast<int64_t>(C_out) * Kpad;
  } else if (mode == 2) {
    DT_REQUIRE(Kpad == 256 && C_in <= 4 && R == 7 && S == 7, DT_ERR_BAD_SHAPE, "dt_pack_conv_weight: stem packing needs 7x7, C_in <= 4, Kpad 256");
    total = static_cast<int64_t>(C_out) * Kpad;
  } else {
    DT_REQUIRE(C_in_p >= C_out && Kpad >= R * S * C_in_p, DT_ERR_BAD_SHAPE, "dt_pack_conv_weight: Cop / Kpad too small");
    total = static_cast<int64_t>(C_in) * Kpad;
  }
  pack_weight_kernel<<<grid_for(total), kThreads, 0, static_cast<cudaStream_t>(stream)>>>(w_oihw, C_out, C_in, R, S, mode,
                                                                                         C_in_p, Kpad, out, total);
  DT_LAUNCH_CHECK();
  return DT_OK;
}

int dt_pack_conv_weights_batched(const dt_pack_job* jobs_device, int njobs, int64_t total, dt_stream_t stream) {
  DT_ARCH_GUARD();
  DT_REQUIRE(jobs_device != nullptr && njobs > 0 && total > 0, DT_ERR_BAD_SHAPE, "dt_pack_conv_weights_batched: empty job list");
  DT_REQUIRE(njobs <= 4096, DT_ERR_BAD_SHAPE, "dt_pack_conv_weights_batched: at most 4096 jobs");
  const size_t smem = (static_cast<size_t>(njobs) + 1) * sizeof(int64_t) + kPackStage * sizeof(float);
  static bool attr_set = false;
  if (!attr_set) {
    DT_CUDA(cudaFuncSetAttribute(pack_weights_batched_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 static_cast<int>((4096 + 1) * sizeof(int64_t) + kPackStage * sizeof(float))));
    attr_set = true;
  }
  pack_weights_batched_kernel<<<dt_num_sms() * 6, kThreads, smem, static_cast<cudaStream_t>(stream)>>>(jobs_device, njobs, total);
  DT_LAUNCH_CHECK();
  return DT_OK;
}

int dt_conv2d_dgrad_direct(const void* gy, const float* w_oihw, const void* addend, int N, int H, int W, int C_in, int C_x,
                           int C_out, int R, int S, int stride, int pad, int dtype, int round_weights, void* gx,
                           dt_stream_t stream) {
  DT_ARCH_GUARD();
  DT_REQUIRE(N > 0 && H > 0 && W > 0 && C_in > 0 && C_x >= C_in && C_out > 0 && R > 0 && S > 0 && stride > 0 && pad >= 0 &&
                 (dtype == DT_F32 || dtype == DT_BF16),
             DT_ERR_BAD_SHAPE, "dt_conv2d_dgrad_direct: bad arguments");
  BwdGeo p{N, H, W, C_in, C_x, C_out, R, S, stride, pad, (H + 2 * pad - R) / stride + 1, (W + 2 * pad - S) / stride + 1};
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int64_t total = static_cast<int64_t>(N) * H * W * C_x;
  DT_DTYPE_SWITCH(dtype,
      (dgrad_direct_kernel<float><<<grid_for(total, 16), kThreads, 0, s>>>(p, static_cast<const float*>(gy), w_oihw, static_cast<const float*>(addend), static_cast<float*>(gx), 0)),
      (dgrad_direct_kernel<__nv_bfloat16><<<grid_for(total, 16), kThreads, 0, s>>>(p, static_cast<const __nv_bfloat16*>(gy), w_oihw, static_cast<const __nv_bfloat16*>(addend), static_cast<__nv_bfloat16*>(gx), round_weights)));
  DT_LAUNCH_CHECK();
  return DT_OK;
}

int dt_conv2d_wgrad_direct(const void* x, const void* gy, int N, int H, int W, int C_in, int C_x, int C_out, int R, int S,
                           int stride, int pad, int dtype, float* dw_oihw, float* dbias, dt_stream_t stream) {
  DT_ARCH_GUARD();
  DT_REQUIRE(N > 0 && H > 0 && W > 0 && C_in > 0 && C_x >= C_in && C_out > 0 && R > 0 && S > 0 && R * S <= 65535 &&
                 stride > 0 && pad >= 0 && (dtype == DT_F32 || dtype == DT_BF16),
             DT_ERR_BAD_SHAPE, "dt_conv2d_wgrad_direct: bad arguments");
  BwdGeo p{N, H, W, C_in, C_x, C_out, R, S, stride, pad, (H + 2 * pad - R) / stride + 1, (W + 2 * pad - S) / stride + 1};
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int64_t M = static_cast<int64_t>(N) * p.Ho * p.Wo;
  const int co_tiles = (C_out + 31) / 32, ci_tiles = (C_in + 31) / 32;
  int64_t chunks = (static_cast<int64_t>(dt_num_sms()) * 8 + co_tiles * ci_tiles * R * S - 1) / (co_tiles * ci_tiles * R * S);
  const int64_t max_chunks = (M + 255) / 256;
  if (chunks > max_chunks) chunks = max_chunks;
  if (chunks < 1) chunks = 1;
  DT_CUDA(cudaMemsetAsync(dw_oihw, 0, sizeof(float) * static_cast<size_t>(C_out) * C_in * R * S, s));
  dim3 grid(static_cast<unsigned>(chunks), co_tiles * ci_tiles, R * S);
  DT_DTYPE_SWITCH(dtype,
      (wgrad_direct_kernel<float><<<grid, kThreads, 0, s>>>(p, static_cast<const float*>(x), static_cast<const float*>(gy), dw_oihw, ci_tiles, M)),
      (wgrad_direct_kernel<__nv_bfloat16><<<grid, kThreads, 0, s>>>(p, static_cast<const __nv_bfloat16*>(x), static_cast<const __nv_bfloat16*>(gy), dw_oihw, ci_tiles, M)));
  DT_LAUNCH_CHECK();
  if (dbias) {
    DT_REQUIRE(C_out <= 4, DT_ERR_BAD_SHAPE, "dt_conv2d_wgrad_direct: bias gradient supports C_out <= 4 (the head)");
    DT_CUDA(cudaMemsetAsync(dbias, 0, sizeof(float) * C_out, s));
    DT_DTYPE_SWITCH(dtype,
        (bias_grad_kernel<float><<<grid_for(M, 2), kThreads, 0, s>>>(static_cast<const float*>(gy), M, C_out, C_out, dbias)),
        (bias_grad_kernel<__nv_bfloat16><<<grid_for(M, 2), kThreads, 0, s>>>(static_cast<const __nv_bfloat16*>(gy), M, C_out, C_out, dbias)));
    DT_LAUNCH_CHECK();
  }
  return DT_OK;
}

}  // extern "C"
