// HBM-bound tiler kernels: block split/merge, tile gather + normalise, mask stitch, blended stitch.
// Reference behaviour: deadtrees/utils/data_handling.py:9-34, deadtrees/deployment/tiler.py:105-170,
// deadtrees/data/deadtreedata.py:148-154 (see include/deadtrees_b200.h for the per-entry citations).
#include "common.cuh"

namespace {

constexpr int kThreads = 256;

inline int grid_for(int64_t work, int threads = kThreads) {
  int64_t blocks = (work + threads - 1) / threads;
  const int64_t cap = static_cast<int64_t>(dt_num_sms()) * 32;  // grid-stride beyond 32 CTAs / SM
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return static_cast<int>(blocks);
}

template <int U> struct Unit;
template <> struct Unit<1> { using T = uint8_t; };
template <> struct Unit<2> { using T = uint16_t; };
template <> struct Unit<4> { using T = uint32_t; };
template <> struct Unit<8> { using T = uint64_t; };
template <> struct Unit<16> { using T = uint4; };

// dst (nb, p, d, d) <- src (p, m, n); everything measured in units of U bytes along the row.
template <int U>
__global__ void make_blocks_kernel(const typename Unit<U>::T* __restrict__ src,
                                   typename Unit<U>::T* __restrict__ dst, int p, int m, int n_u, int d, int d_u) {
  const int nbx = n_u / d_u;
  const int64_t total = static_cast<int64_t>(p) * m * n_u;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    int64_t r = i;
    const int xu = static_cast<int>(r % d_u); r /= d_u;
    const int y = static_cast<int>(r % d); r /= d;
    const int c = static_cast<int>(r % p); r /= p;
    const int b = static_cast<int>(r);
    const int by = b / nbx, bx = b % nbx;
    dst[i] = src[(static_cast<int64_t>(c) * m + (by * d + y)) * n_u + bx * d_u + xu];
  }
}

// dst (m, n) <- src (nb, d, d)
template <int U>
__global__ void unmake_blocks_kernel(const typename Unit<U>::T* __restrict__ src,
                                     typename Unit<U>::T* __restrict__ dst, int m, int n_u, int d, int d_u) {
  const int nbx = n_u / d_u;
  const int64_t total = static_cast<int64_t>(m) * n_u;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int xg = static_cast<int>(i % n_u);
    const int yg = static_cast<int>(i / n_u);
    const int b = (yg / d) * nbx + xg / d_u;
    dst[i] = src[(static_cast<int64_t>(b) * d + (yg % d)) * d_u + (xg % d_u)];
  }
}

int pick_unit(const void* a, const void* b, int64_t row_bytes_a, int64_t row_bytes_b, int elem) {
  for (int u = 16; u > elem; u >>= 1) {
    if (row_bytes_a % u == 0 && row_bytes_b % u == 0 && reinterpret_cast<uintptr_t>(a) % u == 0 &&
        reinterpret_cast<uintptr_t>(b) % u == 0)
      return u;
  }
  return elem;
}

struct NormParams {
  float offset[4];
  float scale[4];
};

// One thread = 4 consecutive pixels of one tile row.  Fast paths read 4 bytes per channel (planar)
// or 12/16 contiguous bytes (interleaved RGB / RGBA); everything else falls back to guarded byte loads.
template <bool OUT_BF16>
__global__ void gather_normalize_kernel(const uint8_t* __restrict__ mosaic, int H, int W, int C, int64_t row_stride,
                                        int64_t pix_stride, int64_t chan_stride, int T, int step, int gx, int tile0,
                                        int ntiles, NormParams np, int pad, int out_hp, int out_wp,
                                        void* __restrict__ out) {
  const int q = T >> 2;
  const int64_t total = static_cast<int64_t>(ntiles) * T * q;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int x4 = static_cast<int>(i % q);
    const int y = static_cast<int>((i / q) % T);
    const int t = static_cast<int>(i / (static_cast<int64_t>(q) * T));
    const int cell = tile0 + t;
    const int gy0 = (cell / gx) * step + y;
    const int gx0 = (cell % gx) * step + x4 * 4;
    uint32_t px[4] = {0u, 0u, 0u, 0u};  // px[j] = bytes c0..c3 of pixel j
    if (gy0 < H && gx0 < W) {
      const uint8_t* base = mosaic + gy0 * row_stride + gx0 * pix_stride;
      const bool full = gx0 + 3 < W;
      if (full && pix_stride == 1 && ((reinterpret_cast<uintptr_t>(base) | chan_stride) & 3) == 0) {
        // planar: one aligned 32-bit load per channel
        for (int c = 0; c < C; ++c) {
          const uint32_t v = __ldg(reinterpret_cast<const uint32_t*>(base + c * chan_stride));
#pragma unroll
          for (int j = 0; j < 4; ++j) px[j] |= ((v >> (8 * j)) & 0xffu) << (8 * c);
        }
      } else if (full && chan_stride == 1 && pix_stride == 4 && C == 4 &&
                 (reinterpret_cast<uintptr_t>(base) & 15) == 0) {
        const uint4 v = __ldg(reinterpret_cast<const uint4*>(base));
        px[0] = v.x; px[1] = v.y; px[2] = v.z; px[3] = v.w;
      } else if (full && chan_stride == 1 && pix_stride == 3 && C == 3 &&
                 (reinterpret_cast<uintptr_t>(base) & 3) == 0) {
        const uint32_t a = __ldg(reinterpret_cast<const uint32_t*>(base));
        const uint32_t b = __ldg(reinterpret_cast<const uint32_t*>(base) + 1);
        const uint32_t c = __ldg(reinterpret_cast<const uint32_t*>(base) + 2);
        px[0] = a & 0xffffffu;
        px[1] = (a >> 24) | ((b & 0xffffu) << 8);
        px[2] = (b >> 16) | ((c & 0xffu) << 16);
        px[3] = c >> 8;
      } else {
        for (int j = 0; j < 4; ++j) {
          if (gx0 + j < W) {
            for (int c = 0; c < C; ++c)
              px[j] |= static_cast<uint32_t>(__ldg(base + j * pix_stride + c * chan_stride)) << (8 * c);
          }
        }
      }
    }
    float v[4][4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const float u = static_cast<float>((px[j] >> (8 * c)) & 0xffu);
        // subtract, then multiply by the reciprocal (albumentations Normalize); no FMA contraction
        v[j][c] = (c < C) ? __fmul_rn(__fsub_rn(u, np.offset[c]), np.scale[c]) : 0.0f;
      }
    }
    // output pixel index: dense (t, y, x) or inside a zero-bordered (t, out_hp, out_wp) frame at offset `pad`
    const int64_t opix = pad ? (static_cast<int64_t>(t) * out_hp + y + pad) * out_wp + x4 * 4 + pad : i * 4;
    if (OUT_BF16) {
      if (pad) {  // 8-byte aligned only
        uint2* o = reinterpret_cast<uint2*>(out) + opix;
#pragma unroll
        for (int j = 0; j < 4; ++j) o[j] = make_uint2(pack_bf16x2(v[j][0], v[j][1]), pack_bf16x2(v[j][2], v[j][3]));
        continue;
      }
      uint4* o = reinterpret_cast<uint4*>(out) + i * 2;
      o[0] = make_uint4(pack_bf16x2(v[0][0], v[0][1]), pack_bf16x2(v[0][2], v[0][3]),
                        pack_bf16x2(v[1][0], v[1][1]), pack_bf16x2(v[1][2], v[1][3]));
      o[1] = make_uint4(pack_bf16x2(v[2][0], v[2][1]), pack_bf16x2(v[2][2], v[2][3]),
                        pack_bf16x2(v[3][0], v[3][1]), pack_bf16x2(v[3][2], v[3][3]));
    } else {
      float4* o = reinterpret_cast<float4*>(out) + opix;
#pragma unroll
      for (int j = 0; j < 4; ++j) o[j] = make_float4(v[j][0], v[j][1], v[j][2], v[j][3]);
    }
  }
}

// bf16 fast path of the gather (the inference pipeline's K1): same arithmetic, organised for the memory system.
// The generic kernel is issue-bound (ncu: 221 instructions per 4 pixels, 76 % issue-active, 19 % DRAM); here
//   * blockIdx.y = tile and each thread walks R rows of it, so the tile -> mosaic coordinate math (one division by the
//     grid width through a host-side magic multiplier) is paid once per R rows;
//   * the source layout is a template parameter (the untaken paths and their address checks disappear);
//   * uint8 -> fp32 through PRMT + FADD (0x4B000000 | byte is the float 2^23 + byte; minus 2^23 is exact) instead of
//     I2F conversions; then the reference's subtract and multiply-by-reciprocal, unfused;
//   * a thread owns 4 consecutive pixels = 32 contiguous output bytes; in the stem frame those start 8 bytes off a
//     16-byte boundary (3 border pixels), so the first pixel of each lane travels to the previous lane and every lane
//     stores two ALIGNED 16-byte vectors {p1,p2} {p3,next p0}; only row ends store single 8-byte pixels.
// LAYOUT: 0 any strides (guarded byte loads)  1 interleaved RGB, 4-byte aligned rows  2 interleaved RGBA, 16-byte
// aligned  3 planar, 4-byte aligned rows and planes.  Requires q = T/4 to divide 256 (q_shift = log2 q).
template <int LAYOUT>
__global__ void __launch_bounds__(256)
gather_normalize_bf16_kernel(const uint8_t* __restrict__ mosaic, int H, int W, int C, int64_t row_stride,
                             int64_t pix_stride, int64_t chan_stride, int T, int q_shift, int step, int gx,
                             uint32_t gx_magic, int tile0, NormParams np, int pad, int out_hp, int out_wp,
                             uint2* __restrict__ out) {
  constexpr int R = 4;
  const int q = T >> 2;
  const int t = blockIdx.y;
  const int cell = tile0 + t;
  const int ty = static_cast<int>(__umulhi(static_cast<uint32_t>(cell), gx_magic));
  const int tx = cell - ty * gx;
  const int lin = blockIdx.x * 256 + threadIdx.x;
  const int x4 = lin & (q - 1);
  const int y0 = lin >> q_shift;
  const int ystep = (gridDim.x * 256) >> q_shift;
  const int gx0 = tx * step + x4 * 4;
  const bool col_in = gx0 < W;
  const bool full = gx0 + 3 < W;
  const int lane = threadIdx.x & 31;
  const uint8_t* col_base = mosaic + gx0 * pix_stride;
#pragma unroll 1
  for (int it = 0; it < R; ++it) {
    const int y = y0 + it * ystep;
    const bool live = y < T;
    const int gy0 = ty * step + y;
    uint32_t px[4] = {0u, 0u, 0u, 0u};  // px[j] = bytes c0..c3 of pixel j
    if (live && col_in && gy0 < H) {
      const uint8_t* base = col_base + gy0 * row_stride;
      if (LAYOUT == 1 && full) {
        const uint32_t a = __ldg(reinterpret_cast<const uint32_t*>(base));
        const uint32_t b = __ldg(reinterpret_cast<const uint32_t*>(base) + 1);
        const uint32_t c = __ldg(reinterpret_cast<const uint32_t*>(base) + 2);
        px[0] = a;                                  // byte 3 of each word is ignored below (C == 3)
        px[1] = __byte_perm(a, b, 0x4543);          // a.b3 b.b0 b.b1
        px[2] = __byte_perm(b, c, 0x4432);          // b.b2 b.b3 c.b0
        px[3] = c >> 8;
      } else if (LAYOUT == 2 && full) {
        const uint4 v = __ldg(reinterpret_cast<const uint4*>(base));
        px[0] = v.x; px[1] = v.y; px[2] = v.z; px[3] = v.w;
      } else if (LAYOUT == 3 && full) {
        for (int c = 0; c < C; ++c) {               // planar: one aligned 32-bit load per channel
          const uint32_t v = __ldg(reinterpret_cast<const uint32_t*>(base + c * chan_stride));
#pragma unroll
          for (int j = 0; j < 4; ++j) px[j] |= ((v >> (8 * j)) & 0xffu) << (8 * c);
        }
      } else {
        for (int j = 0; j < 4; ++j) {
          if (gx0 + j < W) {
            for (int c = 0; c < C; ++c)
              px[j] |= static_cast<uint32_t>(__ldg(base + j * pix_stride + c * chan_stride)) << (8 * c);
          }
        }
      }
    }
    uint2 o[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float v[4];
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        if (LAYOUT == 1 && c == 3) { v[c] = 0.f; continue; }      // RGB: the fourth channel is padding
        const float u = __fadd_rn(__uint_as_float(__byte_perm(px[j], 0x4B000000u, 0x7650 + c)), -8388608.0f);
        // subtract, then multiply by the reciprocal (albumentations Normalize); no FMA contraction.  Channels >= C
        // have offset = scale = 0: (u - 0) * 0 = +0, the padding value, whatever byte sits there
        v[c] = __fmul_rn(__fsub_rn(u, np.offset[c]), np.scale[c]);
      }
      o[j] = make_uint2(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]));
    }
    if (!pad) {          // dense (ntiles, T, T, 4): 32 aligned bytes per thread
      if (live) {
        uint4* dst = reinterpret_cast<uint4*>(out + ((static_cast<int64_t>(t) * T + y) * q + x4) * 4);
        dst[0] = make_uint4(o[0].x, o[0].y, o[1].x, o[1].y);
        dst[1] = make_uint4(o[2].x, o[2].y, o[3].x, o[3].y);
      }
      continue;
    }
    const uint32_t nx = __shfl_down_sync(0xffffffffu, o[0].x, 1);
    const uint32_t ny = __shfl_down_sync(0xffffffffu, o[0].y, 1);
    if (live) {
      uint2* dst = out + (static_cast<int64_t>(t) * out_hp + y + pad) * out_wp + x4 * 4 + pad;   // p0: 8 mod 16 bytes
      if (lane == 0 || x4 == 0) dst[0] = o[0];
      *reinterpret_cast<uint4*>(dst + 1) = make_uint4(o[1].x, o[1].y, o[2].x, o[2].y);
      if (lane < 31 && x4 + 1 < q) *reinterpret_cast<uint4*>(dst + 3) = make_uint4(o[3].x, o[3].y, nx, ny);
      else dst[3] = o[3];
    }
  }
}

// overlap-0 stitch: one thread = V consecutive pixels of one tile row.
template <int V>
__global__ void stitch_mask_kernel(const uint8_t* __restrict__ tiles, int T, int gx, int tile0, int ntiles,
                                   uint8_t* __restrict__ mosaic, int H, int W, int64_t row_stride) {
  const int q = T / V;
  const int64_t total = static_cast<int64_t>(ntiles) * T * q;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int xv = static_cast<int>(i % q);
    const int y = static_cast<int>((i / q) % T);
    const int t = static_cast<int>(i / (static_cast<int64_t>(q) * T));
    const int cell = tile0 + t;
    const int my = (cell / gx) * T + y;
    const int mx = (cell % gx) * T + xv * V;
    if (my >= H || mx >= W) continue;
    const uint8_t* s = tiles + (static_cast<int64_t>(t) * T + y) * T + xv * V;
    uint8_t* d = mosaic + my * row_stride + mx;
    if (V == 16 && mx + 16 <= W && (reinterpret_cast<uintptr_t>(d) & 15) == 0) {
      st_na_v4(d, ld_nc_v4(s));
    } else {
      for (int j = 0; j < V && mx + j < W; ++j) d[j] = s[j];
    }
  }
}

template <typename LT> __device__ __forceinline__ float load_logit(const LT* p);
template <> __device__ __forceinline__ float load_logit<float>(const float* p) { return __ldg(p); }
template <> __device__ __forceinline__ float load_logit<__nv_bfloat16>(const __nv_bfloat16* p) {
  return bf16_bits_to_float(__ldg(reinterpret_cast<const uint16_t*>(p)));
}

// gather-form blend: one thread = one mosaic pixel; visits the <= 4 covering tiles in (ty, tx) order.
template <typename LT, int K>
__global__ void stitch_blend_kernel(const LT* __restrict__ logits, int T, int step, int gy, int gx, int ty_base,
                                    const float* __restrict__ win, uint8_t* __restrict__ mask,
                                    float* __restrict__ blended, int H, int W, int row0, int nrows) {
  const int64_t total = static_cast<int64_t>(nrows) * W;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int x = static_cast<int>(i % W);
    const int y = row0 + static_cast<int>(i / W);
    int ty0 = (y - T + step) / step;  // ceil((y - T + 1) / step) for y - T + 1 > 0
    if (y - T + 1 <= 0) ty0 = 0;
    int tx0 = (x - T + step) / step;
    if (x - T + 1 <= 0) tx0 = 0;
    const int ty1 = min(gy - 1, y / step), tx1 = min(gx - 1, x / step);
    float acc[K];
#pragma unroll
    for (int k = 0; k < K; ++k) acc[k] = 0.0f;
    float wsum = 0.0f;
    for (int ty = ty0; ty <= ty1; ++ty) {
      const int ly = y - ty * step;
      const float wy = __ldg(win + ly);
      for (int tx = tx0; tx <= tx1; ++tx) {
        const int lx = x - tx * step;
        const float w = __fmul_rn(wy, __ldg(win + lx));
        const LT* p = logits + ((static_cast<int64_t>(ty - ty_base) * gx + tx) * T * T + static_cast<int64_t>(ly) * T + lx) * K;
#pragma unroll
        for (int k = 0; k < K; ++k) acc[k] = __fadd_rn(acc[k], __fmul_rn(load_logit<LT>(p + k), w));
        wsum = __fadd_rn(wsum, w);
      }
    }
    int best = 0;
    float bv = __fdiv_rn(acc[0], wsum);
    if (blended) blended[(static_cast<int64_t>(y) * W + x) * K] = bv;
#pragma unroll
    for (int k = 1; k < K; ++k) {
      const float v = __fdiv_rn(acc[k], wsum);
      if (blended) blended[(static_cast<int64_t>(y) * W + x) * K + k] = v;
      if (v > bv) { bv = v; best = k; }
    }
    mask[static_cast<int64_t>(y) * W + x] = static_cast<uint8_t>(best);
  }
}

// Vectorised form for bf16 logits with T, step multiples of 8: one thread = 8 consecutive mosaic pixels (the same
// covering tiles for all eight), 16-byte loads of 8 pixels x K logits per covering tile, one 8-byte mask store.
// Per-pixel arithmetic (order of the fp32 operations) is identical to stitch_blend_kernel.
template <int K>
__global__ void stitch_blend_v8_kernel(const __nv_bfloat16* __restrict__ logits, int T, int step, int gy, int gx,
                                       int ty_base, const float* __restrict__ win, uint8_t* __restrict__ mask,
                                       float* __restrict__ blended, int H, int W, int row0, int nrows) {
  const int W8 = (W + 7) >> 3;
  const int64_t total = static_cast<int64_t>(nrows) * W8;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int x0 = static_cast<int>(i % W8) << 3;
    const int y = row0 + static_cast<int>(i / W8);
    int ty0 = (y - T + step) / step;
    if (y - T + 1 <= 0) ty0 = 0;
    int tx0 = (x0 - T + step) / step;
    if (x0 - T + 1 <= 0) tx0 = 0;
    const int ty1 = min(gy - 1, y / step), tx1 = min(gx - 1, x0 / step);
    float acc[8][K], wsum[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      wsum[j] = 0.f;
#pragma unroll
      for (int k = 0; k < K; ++k) acc[j][k] = 0.f;
    }
    for (int ty = ty0; ty <= ty1; ++ty) {
      const int ly = y - ty * step;
      const float wy = __ldg(win + ly);
      for (int tx = tx0; tx <= tx1; ++tx) {
        const int lx = x0 - tx * step;
        const float4 wa = __ldg(reinterpret_cast<const float4*>(win + lx));
        const float4 wb = __ldg(reinterpret_cast<const float4*>(win + lx) + 1);
        const float wx[8] = {wa.x, wa.y, wa.z, wa.w, wb.x, wb.y, wb.z, wb.w};
        const uint4* src = reinterpret_cast<const uint4*>(
            logits + ((static_cast<int64_t>(ty - ty_base) * gx + tx) * T * T + static_cast<int64_t>(ly) * T + lx) * K);
        uint32_t raw[4 * K];   // 8 pixels x K bf16 = K uint4
#pragma unroll
        for (int q = 0; q < K; ++q) {
          const uint4 v = __ldg(src + q);     // L1-allocating: the K vectors of a thread share 32-byte sectors
          raw[4 * q] = v.x; raw[4 * q + 1] = v.y; raw[4 * q + 2] = v.z; raw[4 * q + 3] = v.w;
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float w = __fmul_rn(wy, wx[j]);
#pragma unroll
          for (int k = 0; k < K; ++k) {
            const int e = j * K + k;
            const uint32_t word = raw[e >> 1];
            const float z = (e & 1) ? __uint_as_float(word & 0xffff0000u) : __uint_as_float(word << 16);
            acc[j][k] = __fadd_rn(acc[j][k], __fmul_rn(z, w));
          }
          wsum[j] = __fadd_rn(wsum[j], w);
        }
      }
    }
    uint8_t best[8];
    if (blended == nullptr) {
      // mask only: x -> fl(x / wsum) is monotone, so the class is the first maximum of the un-normalised sums - except
      // that an EARLIER class whose sum is within rounding distance of the maximum can tie with it after the division
      // (first maximum wins, as np.argmax over the blended logits); only such near-ties pay for the IEEE divisions.
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        int b = 0;
        float bv = acc[j][0];
#pragma unroll
        for (int k = 1; k < K; ++k)
          if (acc[j][k] > bv) { bv = acc[j][k]; b = k; }
        bool near = false;
#pragma unroll
        for (int k = 0; k < K - 1; ++k)
          near = near || (k < b && bv - acc[j][k] <= 4.76837158e-7f * fmaxf(fabsf(bv), fabsf(acc[j][k])));
        if (near) {
          const float qb = __fdiv_rn(bv, wsum[j]);
          int nb = b;
#pragma unroll
          for (int k = K - 2; k >= 0; --k)       // static indices: acc stays in registers
            if (k < b && __fdiv_rn(acc[j][k], wsum[j]) == qb) nb = k;
          b = nb;
        }
        best[j] = static_cast<uint8_t>(b);
      }
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        int b = 0;
        float bv = __fdiv_rn(acc[j][0], wsum[j]);
        const bool live = x0 + j < W;
        if (live) blended[(static_cast<int64_t>(y) * W + x0 + j) * K] = bv;
#pragma unroll
        for (int k = 1; k < K; ++k) {
          const float v = __fdiv_rn(acc[j][k], wsum[j]);
          if (live) blended[(static_cast<int64_t>(y) * W + x0 + j) * K + k] = v;
          if (v > bv) { bv = v; b = k; }
        }
        best[j] = static_cast<uint8_t>(b);
      }
    }
    uint8_t* dst = mask + static_cast<int64_t>(y) * W + x0;
    if (x0 + 8 <= W && (reinterpret_cast<uintptr_t>(dst) & 7) == 0) {
      uint2 pk;
      pk.x = best[0] | (best[1] << 8) | (best[2] << 16) | (static_cast<uint32_t>(best[3]) << 24);
      pk.y = best[4] | (best[5] << 8) | (best[6] << 16) | (static_cast<uint32_t>(best[7]) << 24);
      *reinterpret_cast<uint2*>(dst) = pk;
    } else {
      for (int j = 0; j < 8 && x0 + j < W; ++j) dst[j] = best[j];
    }
  }
}

// ---- mask-only blended stitch in two warp-uniform passes (bf16 logits, T and step multiples of 8) ----------------
// ncu of the one-pass kernel above: 695 instructions per 8 pixels, local-memory traffic, 2 TB/s.  77 % of the pixels
// of the cfg2 mosaic are covered by ONE tile; for them  argmax_k fl(fl(z_k * w) / w)  is the first maximum of the raw
// bf16 logits (two different bf16 values differ by >= 2^-8 relative, the two fp32 roundings move them by 2^-24).
//   pass 1  every 8-pixel group covered by a single tile: K 16-byte loads, compares, one 8-byte store; groups inside
//           an overlap strip return at once (whole rows of blocks for the horizontal strips);
//   pass 2  only the overlap strips (rows [ty*step, ty*step+ov) and column groups [tx*step, tx*step+ov)): the blend
//           arithmetic of the reference order (multiply, add; first maximum of the sums; IEEE division on near-ties).
// Both passes are warp-uniform; the results equal stitch_blend_kernel bit for bit.
struct FastDiv {
  uint32_t magic;
  int d;
  __device__ __forceinline__ int div(int x) const { return static_cast<int>(__umulhi(static_cast<uint32_t>(x), magic)); }
};

template <int K>
__device__ __forceinline__ void unpack_group(const uint4* __restrict__ src, float (&z)[8][K]) {
  uint32_t raw[4 * K];
#pragma unroll
  for (int q = 0; q < K; ++q) {
    const uint4 v = __ldg(src + q);
    raw[4 * q] = v.x; raw[4 * q + 1] = v.y; raw[4 * q + 2] = v.z; raw[4 * q + 3] = v.w;
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) {
#pragma unroll
    for (int k = 0; k < K; ++k) {
      const int e = j * K + k;
      const uint32_t word = raw[e >> 1];
      z[j][k] = (e & 1) ? __uint_as_float(word & 0xffff0000u) : __uint_as_float(word << 16);
    }
  }
}

__device__ __forceinline__ void store_mask8(uint8_t* __restrict__ mask, int y, int x0, int W, const uint32_t (&best)[8]) {
  uint8_t* dst = mask + static_cast<int64_t>(y) * W + x0;
  if (x0 + 8 <= W && (reinterpret_cast<uintptr_t>(dst) & 7) == 0) {
    uint2 pk;
    pk.x = best[0] | (best[1] << 8) | (best[2] << 16) | (best[3] << 24);
    pk.y = best[4] | (best[5] << 8) | (best[6] << 16) | (best[7] << 24);
    *reinterpret_cast<uint2*>(dst) = pk;
  } else {
#pragma unroll
    for (int j = 0; j < 8; ++j) if (x0 + j < W) dst[j] = static_cast<uint8_t>(best[j]);
  }
}

template <int K>
__device__ __forceinline__ void load_raw(const __nv_bfloat16* __restrict__ p, uint4 (&raw)[K]) {
#pragma unroll
  for (int q = 0; q < K; ++q) raw[q] = __ldg(reinterpret_cast<const uint4*>(p) + q);
}

// Thread blocks are (256 / cb) rows x cb column groups, cb a power of two chosen by the host to fit the row length.
// Pass 1: a thread owns one 8-pixel column group in RS rows (rows_pb apart); the loads of all its rows are issued before
// the first compare (RS x K x 16 bytes in flight per thread).
template <int K, int RS>
__device__ __forceinline__ void stitch_single_body(const __nv_bfloat16* __restrict__ logits, int T, int step, FastDiv fd,
                                                   int gy, int gx, int ty_base, uint8_t* __restrict__ mask, int W, int row0,
                                                   int nrows, int cb_shift, int bx, int by) {
  const int rows_pb = 256 >> cb_shift;
  const int x0 = ((bx << cb_shift) + (threadIdx.x & ((1 << cb_shift) - 1))) << 3;
  if (x0 >= W) return;
  const int tx0 = (x0 - T + 1 <= 0) ? 0 : fd.div(x0 - T + step);
  const int tx1 = min(gx - 1, fd.div(x0));
  if (tx0 != tx1) return;                                  // a column group of a vertical overlap strip: pass 2
  const int lx = x0 - tx0 * step;
  const int rbase = by * rows_pb * RS + (threadIdx.x >> cb_shift);
  uint4 raw[RS][K];
  bool ok[RS];
#pragma unroll
  for (int i = 0; i < RS; ++i) {
    const int r = rbase + i * rows_pb;
    const int y = row0 + r;
    const int ty0 = (y - T + 1 <= 0) ? 0 : fd.div(y - T + step);
    const int ty1 = min(gy - 1, fd.div(y));
    ok[i] = r < nrows && ty0 == ty1;                       // else: a row of a horizontal overlap strip (pass 2)
    if (ok[i]) {
      const int ly = y - ty0 * step;
      load_raw<K>(logits + ((static_cast<int64_t>(ty0 - ty_base) * gx + tx0) * T * T + static_cast<int64_t>(ly) * T + lx) * K,
                  raw[i]);
    }
  }
#pragma unroll
  for (int i = 0; i < RS; ++i) {
    if (!ok[i]) continue;
    uint32_t words[4 * K];
#pragma unroll
    for (int q = 0; q < K; ++q) {
      words[4 * q] = raw[i][q].x; words[4 * q + 1] = raw[i][q].y; words[4 * q + 2] = raw[i][q].z; words[4 * q + 3] = raw[i][q].w;
    }
    uint32_t best[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      uint32_t bi = 0;
      float bv = 0.f;
#pragma unroll
      for (int k = 0; k < K; ++k) {
        const int e = j * K + k;
        const uint32_t word = words[e >> 1];
        const float z = (e & 1) ? __uint_as_float(word & 0xffff0000u) : __uint_as_float(word << 16);
        if (k == 0) bv = z;
        else if (z > bv) { bv = z; bi = k; }
      }
      best[j] = bi;
    }
    store_mask8(mask, row0 + rbase + i * rows_pb, x0, W, best);
  }
}

template <int K>
__device__ __forceinline__ void blend_accumulate(const uint4 (&raw)[K], float wy, const float* __restrict__ winx,
                                                 float (&acc)[8][K], float (&wsum)[8]) {
  const float4 wa = __ldg(reinterpret_cast<const float4*>(winx));
  const float4 wb = __ldg(reinterpret_cast<const float4*>(winx) + 1);
  const float wx[8] = {wa.x, wa.y, wa.z, wa.w, wb.x, wb.y, wb.z, wb.w};
  uint32_t words[4 * K];
#pragma unroll
  for (int q = 0; q < K; ++q) { words[4 * q] = raw[q].x; words[4 * q + 1] = raw[q].y; words[4 * q + 2] = raw[q].z; words[4 * q + 3] = raw[q].w; }
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float w = __fmul_rn(wy, wx[j]);
#pragma unroll
    for (int k = 0; k < K; ++k) {
      const int e = j * K + k;
      const uint32_t word = words[e >> 1];
      const float z = (e & 1) ? __uint_as_float(word & 0xffff0000u) : __uint_as_float(word << 16);
      acc[j][k] = __fadd_rn(acc[j][k], __fmul_rn(z, w));
    }
    wsum[j] = __fadd_rn(wsum[j], w);
  }
}

// first maximum of the blended logits of 8 pixels from the un-normalised sums (see stitch_blend_v8_kernel) -> mask
template <int K>
__device__ __forceinline__ void blend_argmax_store(const float (&acc)[8][K], const float (&wsum)[8], uint8_t* __restrict__ mask,
                                                   int y, int x0, int W) {
  uint32_t best[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    int b = 0;
    float bv = acc[j][0];
#pragma unroll
    for (int k = 1; k < K; ++k)
      if (acc[j][k] > bv) { bv = acc[j][k]; b = k; }
    bool near = false;
#pragma unroll
    for (int k = 0; k < K - 1; ++k)
      near = near || (k < b && bv - acc[j][k] <= 4.76837158e-7f * fmaxf(fabsf(bv), fabsf(acc[j][k])));
    if (near) {        // an earlier class can tie with the maximum after the division: first maximum wins
      const float qb = __fdiv_rn(bv, wsum[j]);
      int nb = b;
#pragma unroll
      for (int k = K - 2; k >= 0; --k)       // static indices: acc stays in registers
        if (k < b && __fdiv_rn(acc[j][k], wsum[j]) == qb) nb = k;
      b = nb;
    }
    best[j] = static_cast<uint32_t>(b);
  }
  store_mask8(mask, y, x0, W, best);
}

// part 0: rows = the rows of the horizontal strips (strip-major), columns = all column groups of the row;
// part 1: rows = mosaic rows (rows of horizontal strips return), columns = the column groups of the vertical strips.
// The covering tiles are visited in (ty, tx) order two at a time, both tiles' loads in flight together.
// One launch for both strip parts (round 2: two launches of ~18 us each, 44 % idle lanes in part 1 and 16 warps per SM left
// the strips at 1.8 TB/s while pass 1 ran at 5.9): blocks [0, n0) are part 0 (2-D block grid, bx0 blocks per block row),
// the rest part 1 with a LINEAR index over (row, column group) - no idle lanes -; three blocks per SM.  (Pass 1 as a
// third role of the same launch was slower: it inherits the strips' 80 registers and loses its occupancy.)
template <int K, int RS>
__global__ void __launch_bounds__(256)
stitch_single_kernel(const __nv_bfloat16* __restrict__ logits, int T, int step, FastDiv fd, int gy, int gx, int ty_base,
                     uint8_t* __restrict__ mask, int W, int row0, int nrows, int cb_shift) {
  stitch_single_body<K, RS>(logits, T, step, fd, gy, gx, ty_base, mask, W, row0, nrows, cb_shift, blockIdx.x, blockIdx.y);
}

template <int K>
__global__ void __launch_bounds__(256, 3)
stitch_strips_kernel(const __nv_bfloat16* __restrict__ logits, int T, int step, FastDiv fd, int gy, int gx, int ty_base,
                     const float* __restrict__ win, uint8_t* __restrict__ mask, int H, int W, int row0, int nrows,
                     int n0, int bx0, int first_strip, int nstrip_rows, int cb_shift, int groups, FastDiv gd) {
  const int ov = T - step;
  int y, x0;
  if (static_cast<int>(blockIdx.x) < n0) {
    const int by = blockIdx.x / bx0, bx = blockIdx.x - by * bx0;
    const int r = by * (256 >> cb_shift) + (threadIdx.x >> cb_shift);
    const int c = (bx << cb_shift) + (threadIdx.x & ((1 << cb_shift) - 1));
    if (r >= nstrip_rows) return;
    const int sq = r / ov;
    y = (first_strip + sq) * step + (r - sq * ov);          // tile row first_strip+sq overlaps the previous one here
    if (y < row0 || y >= row0 + nrows) return;
    x0 = c << 3;
    if (x0 >= W) return;
  } else {
    const int idx = (blockIdx.x - n0) * 256 + threadIdx.x;
    const int r = gd.div(idx);
    const int c = idx - r * groups;
    if (r >= nrows) return;
    y = row0 + r;
    const int ty0r = (y - T + 1 <= 0) ? 0 : fd.div(y - T + step);
    if (ty0r != min(gy - 1, fd.div(y))) return;             // done by part 0
    const int ov8 = ov >> 3;
    const int strip = c / ov8;
    x0 = (strip + 1) * step + ((c - strip * ov8) << 3);
    if (x0 >= W) return;
  }
  const int ty0 = (y - T + 1 <= 0) ? 0 : fd.div(y - T + step);
  const int tx0 = (x0 - T + 1 <= 0) ? 0 : fd.div(x0 - T + step);
  const int ty1 = min(gy - 1, fd.div(y)), tx1 = min(gx - 1, fd.div(x0));
  const int ntx = tx1 - tx0 + 1, n = (ty1 - ty0 + 1) * ntx;
  float acc[8][K], wsum[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    wsum[j] = 0.f;
#pragma unroll
    for (int k = 0; k < K; ++k) acc[j][k] = 0.f;
  }
  auto tile_ptr = [&](int i, int& ly, int& lx) {
    const int ty = ty0 + i / ntx, tx = tx0 + i % ntx;
    ly = y - ty * step; lx = x0 - tx * step;
    return logits + ((static_cast<int64_t>(ty - ty_base) * gx + tx) * T * T + static_cast<int64_t>(ly) * T + lx) * K;
  };
  for (int i = 0; i < n; i += 2) {
    int ly0, lx0, ly1 = 0, lx1 = 0;
    uint4 raw0[K], raw1[K];
    load_raw<K>(tile_ptr(i, ly0, lx0), raw0);
    const bool two = i + 1 < n;
    if (two) load_raw<K>(tile_ptr(i + 1, ly1, lx1), raw1);
    blend_accumulate<K>(raw0, __ldg(win + ly0), win + lx0, acc, wsum);
    if (two) blend_accumulate<K>(raw1, __ldg(win + ly1), win + lx1, acc, wsum);
  }
  blend_argmax_store<K>(acc, wsum, mask, y, x0, W);
}

// Strips with at most TWO covering tiles per axis (overlap < step), three warp-uniform roles in one launch (round 2, later):
// in the kernel above a warp of the horizontal strips spans a whole tile width, so every warp contains a corner group and ran
// the four-tile loop for all its lanes (ncu: 820 instructions per warp, issue slots 59 % busy, 2.5 TB/s).  Here
//   blocks [0, nH)        horizontal strips outside the vertical ones: tiles (ty, tx), (ty + 1, tx)        - 2-D block grid
//   blocks [nH, nH + nV)  vertical strips outside the horizontal ones: tiles (ty, tx), (ty, tx + 1)       - linear index
//   the rest              corners: four tiles in (ty, tx) order, two at a time                             - linear index
// Same tile order and arithmetic as the general kernel: identical bits.
template <int K>
__global__ void __launch_bounds__(256, 3)
stitch_strips3_kernel(const __nv_bfloat16* __restrict__ logits, int T, int step, FastDiv fd, int gy, int gx, int ty_base,
                      const float* __restrict__ win, uint8_t* __restrict__ mask, int H, int W, int row0, int nrows,
                      int nH, int nV, int bx0, int first_strip, int nstrip_rows, int cb_shift, int groups, FastDiv gd) {
  const int ov = T - step, ov8 = ov >> 3;
  int y, x0, ty0, tx0, role;
  if (static_cast<int>(blockIdx.x) < nH) {
    role = 0;
    const int by = blockIdx.x / bx0, bx = blockIdx.x - by * bx0;
    const int r = by * (256 >> cb_shift) + (threadIdx.x >> cb_shift);
    const int c = (bx << cb_shift) + (threadIdx.x & ((1 << cb_shift) - 1));
    if (r >= nstrip_rows) return;
    const int sq = r / ov;
    ty0 = first_strip + sq - 1;
    y = (ty0 + 1) * step + (r - sq * ov);                   // tile row ty0 + 1 overlaps tile row ty0 here
    if (y < row0 || y >= row0 + nrows) return;
    x0 = c << 3;
    if (x0 >= W) return;
    tx0 = (x0 - T + 1 <= 0) ? 0 : fd.div(x0 - T + step);
    if (tx0 != min(gx - 1, fd.div(x0))) return;             // a corner: third role
  } else {
    role = static_cast<int>(blockIdx.x) < nH + nV ? 1 : 2;
    const int idx = (blockIdx.x - (role == 1 ? nH : nH + nV)) * 256 + threadIdx.x;
    const int r = gd.div(idx);
    const int c = idx - r * groups;
    const int strip = c / ov8;
    x0 = (strip + 1) * step + ((c - strip * ov8) << 3);
    tx0 = strip;
    if (x0 >= W) return;
    if (role == 1) {
      if (r >= nrows) return;
      y = row0 + r;
      ty0 = (y - T + 1 <= 0) ? 0 : fd.div(y - T + step);
      if (ty0 != min(gy - 1, fd.div(y))) return;            // a row of a horizontal strip: roles 0 / 2
    } else {
      if (r >= nstrip_rows) return;
      const int sq = r / ov;
      ty0 = first_strip + sq - 1;
      y = (ty0 + 1) * step + (r - sq * ov);
      if (y < row0 || y >= row0 + nrows) return;
    }
  }
  float acc[8][K], wsum[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    wsum[j] = 0.f;
#pragma unroll
    for (int k = 0; k < K; ++k) acc[j][k] = 0.f;
  }
  auto tile_ptr = [&](int ty, int tx) {
    return logits + ((static_cast<int64_t>(ty - ty_base) * gx + tx) * T * T + static_cast<int64_t>(y - ty * step) * T +
                     (x0 - tx * step)) * K;
  };
  // second tile of the first pair: the lower tile (role 0) or the right-hand tile (roles 1, 2)
  const int ty_b = role == 0 ? ty0 + 1 : ty0, tx_b = role == 0 ? tx0 : tx0 + 1;
  {
    uint4 raw0[K], raw1[K];
    load_raw<K>(tile_ptr(ty0, tx0), raw0);
    load_raw<K>(tile_ptr(ty_b, tx_b), raw1);
    blend_accumulate<K>(raw0, __ldg(win + y - ty0 * step), win + x0 - tx0 * step, acc, wsum);
    blend_accumulate<K>(raw1, __ldg(win + y - ty_b * step), win + x0 - tx_b * step, acc, wsum);
  }
  if (role == 2) {
    uint4 raw0[K], raw1[K];
    load_raw<K>(tile_ptr(ty0 + 1, tx0), raw0);
    load_raw<K>(tile_ptr(ty0 + 1, tx0 + 1), raw1);
    const float wy = __ldg(win + y - (ty0 + 1) * step);
    blend_accumulate<K>(raw0, wy, win + x0 - tx0 * step, acc, wsum);
    blend_accumulate<K>(raw1, wy, win + x0 - (tx0 + 1) * step, acc, wsum);
  }
  blend_argmax_store<K>(acc, wsum, mask, y, x0, W);
}

// smallest waste of ceil(n / cb) * cb over cb in {32, 64, 128, 256}; ties go to the wider block row
inline int pick_cb_shift(int n) {
  int best_shift = 8, best_waste = (n + 255) / 256 * 256;
  for (int sh = 7; sh >= 5; --sh) {
    const int cb = 1 << sh, tot = (n + cb - 1) / cb * cb;
    if (tot < best_waste) { best_waste = tot; best_shift = sh; }
  }
  return best_shift;
}

template <int K>
void launch_stitch_two_pass(const __nv_bfloat16* lg, int T, int step, int gy, int gx, int ty_base, const float* win,
                            uint8_t* mask, int H, int W, int row0, int nrows, cudaStream_t s) {
  FastDiv fd;
  fd.d = step;
  fd.magic = static_cast<uint32_t>(((1ull << 32) + step - 1) / step);
  const int W8 = (W + 7) / 8, ov = T - step;
  const int sh = pick_cb_shift(W8), rows_pb = 256 >> sh;
  const int bx0 = (W8 + (1 << sh) - 1) >> sh;
  // pass 1 is DRAM-bound with one row per thread (5.9 TB/s; two or four rows per thread measured the same)
  stitch_single_kernel<K, 1><<<dim3(bx0, (nrows + rows_pb - 1) / rows_pb), 256, 0, s>>>(lg, T, step, fd, gy, gx, ty_base, mask,
                                                                                      W, row0, nrows, sh);
  if (ov == 0) return;
  // horizontal strips that intersect [row0, row0 + nrows): tile rows `first`..`last` (>= 1)
  int first = (row0 - ov + 1 <= 0) ? 1 : (row0 - ov + step) / step;          // smallest ty with ty*step + ov > row0
  if (first < 1) first = 1;
  int last = (row0 + nrows - 1) / step;
  if (last > gy - 1) last = gy - 1;
  int n0 = 0, srows = 0;
  if (last >= first) {
    srows = (last - first + 1) * ov;
    n0 = bx0 * ((srows + rows_pb - 1) / rows_pb);
  }
  const int groups = gx > 1 ? (gx - 1) * (ov / 8) : 0;
  const int n1 = static_cast<int>((static_cast<int64_t>(nrows) * groups + 255) / 256);   // items < 2^32 / groups
  if (n0 + n1 == 0) return;
  FastDiv gd;
  gd.d = groups > 0 ? groups : 1;
  gd.magic = static_cast<uint32_t>(((1ull << 32) + gd.d - 1) / gd.d);
  static const bool general = [] { const char* e = std::getenv("DT_STITCH_GENERAL"); return e && e[0] == '1'; }();
  if (ov < step && !general) {          // at most two covering tiles per axis: the three-role kernel
    const int n2 = static_cast<int>((static_cast<int64_t>(srows) * groups + 255) / 256);   // corners
    stitch_strips3_kernel<K><<<n0 + n1 + n2, 256, 0, s>>>(lg, T, step, fd, gy, gx, ty_base, win, mask, H, W, row0, nrows, n0,
                                                          n1, bx0, first, srows, sh, gd.d, gd);
    return;
  }
  stitch_strips_kernel<K><<<n0 + n1, 256, 0, s>>>(lg, T, step, fd, gy, gx, ty_base, win, mask, H, W, row0, nrows, n0, bx0,
                                                  first, srows, sh, gd.d, gd);
}

// (N, C_src, H, W) fp32 planes -> (N, H, W, 4) NHWC, first C channels kept (RGB slice), rest zero.
template <bool OUT_BF16>
__global__ void pack_nchw_kernel(const float* __restrict__ x, int N, int C_src, int C, int64_t HW,
                                 void* __restrict__ out) {
  const int64_t total = static_cast<int64_t>(N) * HW;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t n = i / HW, px = i % HW;
    float v[4] = {0.f, 0.f, 0.f, 0.f};
    for (int c = 0; c < C; ++c) v[c] = __ldg(x + (n * C_src + c) * HW + px);
    if (OUT_BF16)
      reinterpret_cast<uint2*>(out)[i] = make_uint2(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]));
    else
      reinterpret_cast<float4*>(out)[i] = make_float4(v[0], v[1], v[2], v[3]);
  }
}

// Same conversion into the stem's zero-bordered bf16 frame (N, H+6, W+8, 4) (conv_stem.cu): the interior sits at
// offset (3, 3); the kernel writes the WHOLE frame, borders included, so the caller need not clear it.
__global__ void pack_nchw_frame_kernel(const float* __restrict__ x, int N, int C_src, int C, int H, int W,
                                       uint2* __restrict__ out) {
  const int Hp = H + 6, Wp = W + 8;
  const int64_t HW = static_cast<int64_t>(H) * W;
  const int64_t total = static_cast<int64_t>(N) * Hp * Wp;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int wp = static_cast<int>(i % Wp);
    const int hp = static_cast<int>((i / Wp) % Hp);
    const int64_t n = i / (static_cast<int64_t>(Wp) * Hp);
    const int h = hp - 3, w = wp - 3;
    float v[4] = {0.f, 0.f, 0.f, 0.f};
    if (h >= 0 && h < H && w >= 0 && w < W)
      for (int c = 0; c < C; ++c) v[c] = __ldg(x + (n * C_src + c) * HW + static_cast<int64_t>(h) * W + w);
    out[i] = make_uint2(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]));
  }
}

}  // namespace

extern "C" {

int dt_pack_input_nchw(const float* x, int N, int C_src, int C, int H, int W, int out_dtype, void* out,
                       dt_stream_t stream) {
  DT_ARCH_GUARD();
  DT_REQUIRE(N > 0 && H > 0 && W > 0 && C >= 1 && C <= 4 && C_src >= C, DT_ERR_BAD_SHAPE,
             "dt_pack_input_nchw: N=%d C_src=%d C=%d", N, C_src, C);
  DT_REQUIRE(out_dtype == DT_BF16 || out_dtype == DT_F32, DT_ERR_BAD_SHAPE, "dt_pack_input_nchw: dtype %d", out_dtype);
  DT_REQUIRE(reinterpret_cast<uintptr_t>(out) % 16 == 0, DT_ERR_BAD_ALIGN, "dt_pack_input_nchw: out alignment");
  const int64_t HW = static_cast<int64_t>(H) * W;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (out_dtype == DT_BF16)
    pack_nchw_kernel<true><<<grid_for(N * HW), kThreads, 0, s>>>(x, N, C_src, C, HW, out);
  else
    pack_nchw_kernel<false><<<grid_for(N * HW), kThreads, 0, s>>>(x, N, C_src, C, HW, out);
  DT_LAUNCH_CHECK();
  return DT_OK;
}

int dt_pack_input_nchw_frame(const float* x, int N, int C_src, int C, int H, int W, void* out, dt_stream_t stream) {
  DT_ARCH_GUARD();
  DT_REQUIRE(N > 0 && H > 0 && W > 0 && C >= 1 && C <= 4 && C_src >= C, DT_ERR_BAD_SHAPE,
             "dt_pack_input_nchw_frame: N=%d C_src=%d C=%d", N, C_src, C);
  DT_REQUIRE(reinterpret_cast<uintptr_t>(out) % 16 == 0, DT_ERR_BAD_ALIGN, "dt_pack_input_nchw_frame: out alignment");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  pack_nchw_frame_kernel<<<grid_for(static_cast<int64_t>(N) * (H + 6) * (W + 8)), kThreads, 0, s>>>(
      x, N, C_src, C, H, W, static_cast<uint2*>(out));
  DT_LAUNCH_CHECK();
  return DT_OK;
}

int dt_make_blocks(const void* src, int p, int m, int n, int d, int elem_size, void* dst, dt_stream_t stream) {
  DT_ARCH_GUARD();
  DT_REQUIRE(p > 0 && m > 0 && n > 0 && d > 0 && m % d == 0 && n % d == 0, DT_ERR_BAD_SHAPE,
             "dt_make_blocks: (p,m,n)=(%d,%d,%d) not divisible by d=%d", p, m, n, d);
  DT_REQUIRE(elem_size == 1 || elem_size == 2 || elem_size == 4 || elem_size == 8, DT_ERR_BAD_SHAPE,
             "dt_make_blocks: elem_size %d", elem_size);
  DT_REQUIRE(reinterpret_cast<uintptr_t>(src) % elem_size == 0 && reinterpret_cast<uintptr_t>(dst) % elem_size == 0,
             DT_ERR_BAD_ALIGN, "dt_make_blocks: pointers not aligned to the element size");
  const int u = pick_unit(src, dst, static_cast<int64_t>(n) * elem_size, static_cast<int64_t>(d) * elem_size, elem_size);
  const int n_u = n * elem_size / u, d_u = d * elem_size / u;
  const int64_t total = static_cast<int64_t>(p) * m * n_u;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
#define DT_MB(U)                                                                                      \
  make_blocks_kernel<U><<<grid_for(total), kThreads, 0, s>>>(                                         \
      static_cast<const Unit<U>::T*>(src), static_cast<Unit<U>::T*>(dst), p, m, n_u, d, d_u)
  switch (u) {
    case 16: DT_MB(16); break;
    case 8: DT_MB(8); break;
    case 4: DT_MB(4); break;
    case 2: DT_MB(2); break;
    default: DT_MB(1); break;
  }
#undef DT_MB
  DT_LAUNCH_CHECK();
  return DT_OK;
}

int dt_unmake_blocks(const void* src, int d, int m, int n, int elem_size, void* dst, dt_stream_t stream) {
  DT_ARCH_GUARD();
  DT_REQUIRE(m > 0 && n > 0 && d > 0 && m % d == 0 && n % d == 0, DT_ERR_BAD_SHAPE,
             "dt_unmake_blocks: (m,n)=(%d,%d) not divisible by d=%d", m, n, d);
  DT_REQUIRE(elem_size == 1 || elem_size == 2 || elem_size == 4 || elem_size == 8, DT_ERR_BAD_SHAPE,
             "dt_unmake_blocks: elem_size %d", elem_size);
  DT_REQUIRE(reinterpret_cast<uintptr_t>(src) % elem_size == 0 && reinterpret_cast<uintptr_t>(dst) % elem_size == 0,
             DT_ERR_BAD_ALIGN, "dt_unmake_blocks: pointers not aligned to the element size");
  const int u = pick_unit(src, dst, static_cast<int64_t>(n) * elem_size, static_cast<int64_t>(d) * elem_size, elem_size);
  const int n_u = n * elem_size / u, d_u = d * elem_size / u;
  const int64_t total = static_cast<int64_t>(m) * n_u;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
#define DT_UB(U)                                                                                      \
  unmake_blocks_kernel<U><<<grid_for(total), kThreads, 0, s>>>(                                       \
      static_cast<const Unit<U>::T*>(src), static_cast<Unit<U>::T*>(dst), m, n_u, d, d_u)
  switch (u) {
    case 16: DT_UB(16); break;
    case 8: DT_UB(8); break;
    case 4: DT_UB(4); break;
    case 2: DT_UB(2); break;
    default: DT_UB(1); break;
  }
#undef DT_UB
  DT_LAUNCH_CHECK();
  return DT_OK;
}

int dt_tile_gather_normalize(const uint8_t* mosaic, int H, int W, int C, int64_t row_stride, int64_t pix_stride,
                             int64_t chan_stride, int tile, int step, int gx, int tile0, int ntiles,
                             const float* offset, const float* scale, int c_out, int out_dtype, int out_pad, void* out,
                             dt_stream_t stream) {
  DT_ARCH_GUARD();
  DT_REQUIRE(out_pad == 0 || out_pad == 3, DT_ERR_BAD_SHAPE, "dt_tile_gather_normalize: out_pad must be 0 or 3");
  const int out_hp = tile + 6, out_wp = tile + 8;
  DT_REQUIRE(C >= 1 && C <= 4 && c_out == 4, DT_ERR_BAD_SHAPE, "dt_tile_gather_normalize: C=%d c_out=%d (need C<=4, c_out==4)", C, c_out);
  DT_REQUIRE(tile > 0 && tile % 4 == 0 && step > 0 && step <= tile && gx > 0 && ntiles >= 0 && tile0 >= 0,
             DT_ERR_BAD_SHAPE, "dt_tile_gather_normalize: tile=%d step=%d gx=%d", tile, step, gx);
  DT_REQUIRE(out_dtype == DT_BF16 || out_dtype == DT_F32, DT_ERR_BAD_SHAPE, "dt_tile_gather_normalize: out_dtype %d", out_dtype);
  DT_REQUIRE(reinterpret_cast<uintptr_t>(out) % 16 == 0, DT_ERR_BAD_ALIGN, "dt_tile_gather_normalize: out must be 16-byte aligned");
  if (ntiles == 0) return DT_OK;
  NormParams np;
  for (int c = 0; c < 4; ++c) {
    np.offset[c] = c < C ? offset[c] : 0.f;
    np.scale[c] = c < C ? scale[c] : 0.f;
  }
  const int64_t total = static_cast<int64_t>(ntiles) * tile * (tile / 4);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int q = tile / 4;
  int q_shift = -1;
  for (int b = 0; b <= 8; ++b) if ((1 << b) == q) q_shift = b;
  // magic multiplier for cell / gx (exact while cell * gx < 2^32)
  const uint64_t cells = static_cast<uint64_t>(tile0) + ntiles;
  if (out_dtype == DT_BF16 && ntiles <= 65535 && q_shift >= 0 && cells * gx < (1ull << 32) && gx > 1) {
    const uint32_t magic = static_cast<uint32_t>(((1ull << 32) + gx - 1) / gx);
    const int rows_per_pass = 256 >> q_shift;
    const dim3 grid((tile + 4 * rows_per_pass - 1) / (4 * rows_per_pass), ntiles);
    const uintptr_t mp = reinterpret_cast<uintptr_t>(mosaic);
    int layout = 0;
    if (step % 4 == 0) {
      if (chan_stride == 1 && pix_stride == 3 && C == 3 && mp % 4 == 0 && row_stride % 4 == 0) layout = 1;
      else if (chan_stride == 1 && pix_stride == 4 && C == 4 && mp % 16 == 0 && row_stride % 16 == 0) layout = 2;
      else if (pix_stride == 1 && mp % 4 == 0 && row_stride % 4 == 0 && chan_stride % 4 == 0) layout = 3;
    }
#define DT_GATHER(L)                                                                                              \
  gather_normalize_bf16_kernel<L><<<grid, kThreads, 0, s>>>(mosaic, H, W, C, row_stride, pix_stride, chan_stride, tile, \
                                                           q_shift, step, gx, magic, tile0, np, out_pad, out_hp,  \
                                                           out_wp, static_cast<uint2*>(out))
    switch (layout) {
      case 1: DT_GATHER(1); break;
      case 2: DT_GATHER(2); break;
      case 3: DT_GATHER(3); break;
      default: DT_GATHER(0); break;
    }
#undef DT_GATHER
  } else if (out_dtype == DT_BF16)
    gather_normalize_kernel<true><<<grid_for(total), kThreads, 0, s>>>(mosaic, H, W, C, row_stride, pix_stride,
                                                                      chan_stride, tile, step, gx, tile0, ntiles, np,
                                                                      out_pad, out_hp, out_wp, out);
  else
    gather_normalize_kernel<false><<<grid_for(total), kThreads, 0, s>>>(mosaic, H, W, C, row_stride, pix_stride,
                                                                       chan_stride, tile, step, gx, tile0, ntiles, np,
                                                                       out_pad, out_hp, out_wp, out);
  DT_LAUNCH_CHECK();
  return DT_OK;
}

int dt_stitch_mask_u8(const uint8_t* tile_masks, int T, int gx, int tile0, int ntiles, uint8_t* mosaic_mask, int H,
                      int W, int64_t row_stride, dt_stream_t stream) {
  DT_ARCH_GUARD();
  DT_REQUIRE(T > 0 && gx > 0 && ntiles >= 0 && tile0 >= 0 && H > 0 && W > 0 && row_stride >= W, DT_ERR_BAD_SHAPE,
             "dt_stitch_mask_u8: T=%d gx=%d H=%d W=%d", T, gx, H, W);
  if (ntiles == 0) return DT_OK;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (T % 16 == 0 && reinterpret_cast<uintptr_t>(tile_masks) % 16 == 0) {
    const int64_t total = static_cast<int64_t>(ntiles) * T * (T / 16);
    stitch_mask_kernel<16><<<grid_for(total), kThreads, 0, s>>>(tile_masks, T, gx, tile0, ntiles, mosaic_mask, H, W, row_stride);
  } else {
    const int64_t total = static_cast<int64_t>(ntiles) * T * T;
    stitch_mask_kernel<1><<<grid_for(total), kThreads, 0, s>>>(tile_masks, T, gx, tile0, ntiles, mosaic_mask, H, W, row_stride);
  }
  DT_LAUNCH_CHECK();
  return DT_OK;
}

int dt_stitch_blend_argmax(const void* logits, int dtype, int K, int T, int overlap, int gy, int gx, int ty_base,
                           const float* win, uint8_t* mosaic_mask, float* blended, int H, int W, int row0,
                           int nrows, dt_stream_t stream) {
  DT_ARCH_GUARD();
  DT_REQUIRE(K >= 1 && K <= 4, DT_ERR_BAD_SHAPE, "dt_stitch_blend_argmax: K=%d (1..4)", K);
  DT_REQUIRE(T > 0 && overlap >= 0 && overlap < T && gy > 0 && gx > 0, DT_ERR_BAD_SHAPE,
             "dt_stitch_blend_argmax: T=%d overlap=%d grid=%dx%d", T, overlap, gy, gx);
  DT_REQUIRE(ty_base >= 0 && (row0 - T + 1 <= 0 ? 0 : (row0 - overlap) / (T - overlap)) >= ty_base, DT_ERR_BAD_SHAPE,
             "dt_stitch_blend_argmax: rows from %d need tile rows before ty_base=%d", row0, ty_base);
  DT_REQUIRE(row0 >= 0 && nrows >= 0 && row0 + nrows <= H && H <= (gy - 1) * (T - overlap) + T &&
                 W <= (gx - 1) * (T - overlap) + T,
             DT_ERR_BAD_SHAPE, "dt_stitch_blend_argmax: rows [%d,%d) of %dx%d not covered by the grid", row0, row0 + nrows, H, W);
  if (nrows == 0) return DT_OK;
  const int step = T - overlap;
  const int64_t total = static_cast<int64_t>(nrows) * W;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
#define DT_SB(LT, KK)                                                                                  \
  stitch_blend_kernel<LT, KK><<<grid_for(total), kThreads, 0, s>>>(static_cast<const LT*>(logits), T, step, gy, gx, \
                                                                  ty_base, win, mosaic_mask, blended, H, W, row0, nrows)
#define DT_SBK(LT)                     \
  switch (K) {                         \
    case 1: DT_SB(LT, 1); break;       \
    case 2: DT_SB(LT, 2); break;       \
    case 3: DT_SB(LT, 3); break;       \
    default: DT_SB(LT, 4); break;      \
  }
  const bool vec8 = dtype == DT_BF16 && T % 8 == 0 && step % 8 == 0 && reinterpret_cast<uintptr_t>(logits) % 16 == 0 &&
                    reinterpret_cast<uintptr_t>(win) % 16 == 0;
  if (vec8 && blended == nullptr && step > 1 &&
      static_cast<int64_t>(H > W ? H : W) * step < (1ll << 32)) {
    const __nv_bfloat16* lg = static_cast<const __nv_bfloat16*>(logits);
    switch (K) {
      case 1: launch_stitch_two_pass<1>(lg, T, step, gy, gx, ty_base, win, mosaic_mask, H, W, row0, nrows, s); break;
      case 2: launch_stitch_two_pass<2>(lg, T, step, gy, gx, ty_base, win, mosaic_mask, H, W, row0, nrows, s); break;
      case 3: launch_stitch_two_pass<3>(lg, T, step, gy, gx, ty_base, win, mosaic_mask, H, W, row0, nrows, s); break;
      default: launch_stitch_two_pass<4>(lg, T, step, gy, gx, ty_base, win, mosaic_mask, H, W, row0, nrows, s); break;
    }
  } else if (vec8) {
    const int64_t total8 = static_cast<int64_t>(nrows) * ((W + 7) / 8);
    const __nv_bfloat16* lg = static_cast<const __nv_bfloat16*>(logits);
    switch (K) {
      case 1: stitch_blend_v8_kernel<1><<<grid_for(total8), kThreads, 0, s>>>(lg, T, step, gy, gx, ty_base, win, mosaic_mask, blended, H, W, row0, nrows); break;
      case 2: stitch_blend_v8_kernel<2><<<grid_for(total8), kThreads, 0, s>>>(lg, T, step, gy, gx, ty_base, win, mosaic_mask, blended, H, W, row0, nrows); break;
      case 3: stitch_blend_v8_kernel<3><<<grid_for(total8), kThreads, 0, s>>>(lg, T, step, gy, gx, ty_base, win, mosaic_mask, blended, H, W, row0, nrows); break;
      default: stitch_blend_v8_kernel<4><<<grid_for(total8), kThreads, 0, s>>>(lg, T, step, gy, gx, ty_base, win, mosaic_mask, blended, H, W, row0, nrows); break;
    }
  } else if (dtype == DT_BF16) { DT_SBK(__nv_bfloat16) } else if (dtype == DT_F32) { DT_SBK(float) } else {
    DT_REQUIRE(false, DT_ERR_BAD_SHAPE, "dt_stitch_blend_argmax: dtype %d", dtype);
  }
#undef DT_SBK
#undef DT_SB
  DT_LAUNCH_CHECK();
  return DT_OK;
}

}  // extern "C"
