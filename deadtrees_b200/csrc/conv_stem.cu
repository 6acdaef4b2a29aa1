// ResNet stem: 7x7 / stride-2 / pad-3 convolution C_in(<=4, stored as 4) -> 64 with folded BN + ReLU,
// as a tcgen05 implicit GEMM whose im2col operand is produced by TMA alone.
//
// The input lives in a zero-bordered NHWC4 bf16 buffer (N, H+6, W+8, 4): 3 border rows / columns on the top / left
// (the conv padding - no out-of-bounds handling needed) and a row pitch of (W+8)*8 bytes.  A 5-D tensor map with
// OVERLAPPING strides views it as the im2col matrix
//     (32 elements, Wo, Ho, 7 filter rows, N)   strides: 2 B | 16 B | 2*pitch | pitch | image
// i.e. element (e, wo, ho, r, n) = x_pad[n][2*ho + r][2*wo + e/4][e%4]: the 8 pixels x 4 channels under filter row r
// of output pixel (ho, wo) are one contiguous 64-byte row (the 8th pixel meets a zero weight).  One TMA box per filter
// row delivers a 128-pixel x 64-byte K-major SWIZZLE_64B operand tile (measured to work with overlapping strides:
// scripts/exp/tma_overlap_exp.cu).  The 28 KB of weights [64][7*32] stay resident in shared memory.
//
//   warp 0: TMA producer   warp 1: TMEM alloc + MMA issuer (14 MMAs of K=16 per tile)   warps 2..5: epilogue
#include <cstring>
#include <mutex>

#include <cstdlib>

#include "common.cuh"

namespace {

constexpr int BM = 128, BN = 64;
constexpr int ROWS = 7;                         // filter rows = K chunks of 32 elements
constexpr int A_SUB = BM * 64;                  // one filter row of one tile: 128 x 64 B
constexpr int A_STAGE = ROWS * A_SUB;           // 56 KB
constexpr int W_SUB = BN * 64;                  // 64 x 64 B
constexpr int W_BYTES = ROWS * W_SUB;           // 28 KB
constexpr int STAGES = 3;
constexpr int kThreads = 192;

struct StemParams {
  int Ho, Wo, M_total, total_tiles, relu;
  __nv_bfloat16* y;
  const float* scale;
  const float* shift;
};

__device__ __forceinline__ void tma_load_5d(void* dst, const void* tmap, uint64_t* bar, int c0, int c1, int c2, int c3,
                                            int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
      :: "r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3),
         "r"(c4)
      : "memory");
}

__global__ void __launch_bounds__(kThreads, 1)
conv_stem_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b, const StemParams p) {
  constexpr int TMEM_COLS = 2 * BN;
  constexpr uint32_t idesc = umma_idesc_bf16(BM, BN);
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_w = smem;
  uint8_t* smem_a = smem + W_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_a + STAGES * A_STAGE);
  uint64_t* w_full = bars;
  uint64_t* full = w_full + 1;          // [STAGES]
  uint64_t* empty = full + STAGES;      // [STAGES]
  uint64_t* tmem_full = empty + STAGES; // [2]
  uint64_t* tmem_empty = tmem_full + 2; // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_a);
    tma_prefetch_desc(&tm_b);
    mbar_init(w_full, 1u);
    for (int i = 0; i < STAGES; ++i) { mbar_init(&full[i], 1u); mbar_init(&empty[i], 1u); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tmem_full[i], 1u); mbar_init(&tmem_empty[i], 128u); }
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);   // warp-uniform for the compiler (uniform registers)

  if (warp == 0) {
    if (lane == 0) {
      mbar_arrive_expect_tx(w_full, W_BYTES);
      for (int r = 0; r < ROWS; ++r) tma_load_2d(smem_w + r * W_SUB, &tm_b, w_full, r * 32, 0);
      int st = 0;
      uint32_t ph = 0;
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
        const int m0 = tile * BM;
        const int w0 = m0 % p.Wo, h0 = (m0 / p.Wo) % p.Ho, n0 = m0 / (p.Wo * p.Ho);
        mbar_wait(&empty[st], ph ^ 1u);
        mbar_arrive_expect_tx(&full[st], A_STAGE);
        for (int r = 0; r < ROWS; ++r)
          tma_load_5d(smem_a + st * A_STAGE + r * A_SUB, &tm_a, &full[st], 0, w0, h0, r, n0);
        if (++st == STAGES) { st = 0; ph ^= 1u; }
      }
    }
  } else if (warp == 1) {
    {   // whole warp: uniform bookkeeping; an elected lane issues (see conv_halo.cu)
      const uint64_t a_hi = umma_desc(0u, 512u, 4u);   // 64 B rows, SWIZZLE_64B
      const uint64_t b_d0 = umma_desc(smem_u32(smem_w), 512u, 4u);
      mbar_wait(w_full, 0);
      int st = 0, buf = 0;
      uint32_t ph = 0, pbuf = 0;
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
        mbar_wait(&tmem_empty[buf], pbuf ^ 1u);
        mbar_wait(&full[st], ph);
        tc_fence_after();
        const uint64_t a_d = a_hi + (smem_u32(smem_a + st * A_STAGE) >> 4);
        const uint32_t d_tmem = tmem_base + buf * BN;
        if (elect_one()) {
#pragma unroll
          for (int r = 0; r < ROWS; ++r) {
#pragma unroll
            for (int k = 0; k < 2; ++k)
              umma_bf16_ss(d_tmem, a_d + ((r * A_SUB + k * 32) >> 4), b_d0 + ((r * W_SUB + k * 32) >> 4), idesc,
                           (r | k) != 0 ? 1u : 0u);
          }
          umma_commit(&empty[st]);
          umma_commit(&tmem_full[buf]);
        }
        __syncwarp();
        if (++st == STAGES) { st = 0; ph ^= 1u; }
        if ((buf ^= 1) == 0) pbuf ^= 1u;
      }
    }
  } else {
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    int buf = 0;
    uint32_t pbuf = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
      const int m = tile * BM + row;
      const bool valid = m < p.M_total;
      mbar_wait(&tmem_full[buf], pbuf);
      tc_fence_after();
      const uint32_t t_row = tmem_base + buf * BN + (static_cast<uint32_t>(quarter * 32) << 16);
#pragma unroll
      for (int c0 = 0; c0 < BN; c0 += 16) {
        uint32_t v[16];
        tmem_ld_x16(t_row + c0, v);
        tmem_ld_wait();
        float f[16];
#pragma unroll
        for (int j = 0; j < 16; j += 4) {
          const float4 sc = __ldg(reinterpret_cast<const float4*>(p.scale + c0 + j));
          const float4 sh = __ldg(reinterpret_cast<const float4*>(p.shift + c0 + j));
          f[j] = fmaf(__uint_as_float(v[j]), sc.x, sh.x);
          f[j + 1] = fmaf(__uint_as_float(v[j + 1]), sc.y, sh.y);
          f[j + 2] = fmaf(__uint_as_float(v[j + 2]), sc.z, sh.z);
          f[j + 3] = fmaf(__uint_as_float(v[j + 3]), sc.w, sh.w);
        }
        if (p.relu) {
#pragma unroll
          for (int j = 0; j < 16; ++j) f[j] = fmaxf(f[j], 0.f);
        }
        if (valid) {
          store_bf16x16(p.y + static_cast<int64_t>(m) * BN + c0, f);
        }
      }
      tc_fence_before();
      mbar_arrive(&tmem_empty[buf]);
      if ((buf ^= 1) == 0) pbuf ^= 1u;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

}  // namespace

// ------------------------------------------------------------------------------------------------------------------
// Row form of the stem (output width a multiple of 128): the im2col operand is read IN PLACE from the raw input rows.
//
// The im2col tensor map above fetches 7 x (128 px x 64 B) = 56 KB per 128 output pixels - every input pixel crosses
// L2 -> SM up to 28 times (ncu: 1.06 GB of L2 reads per 135-tile launch, l1tex 70 % busy, 180 us against ~50 us for
// the 0.36 GB the layer reads and writes).  In the zero-bordered frame the 8 pixels x 4 channels under filter row r of
// output pixel wo start at byte 16*wo of input row 2*ho + r: consecutive output pixels are 16 bytes apart and the two
// 16-byte halves of a K = 16 step are 16 bytes apart - exactly the NO-SWIZZLE K-major canonical layout
// ((8 rows at 16 B, row groups at SBO = 128 B), (8 elements, chunks at LBO = 16 B)) with overlapping core matrices.
// So ONE bulk copy brings the 7 (contiguous) input rows of an output row into shared memory (14.8 KB at T = 256,
// 3.8x less than the im2col boxes) and the 14 MMAs of a tile point their A descriptors into it:
// start = row r + 16 * wo0 (+ 32 B for the second K step).  Weights stay SWIZZLE_64B as above.
//   warp 0: bulk-copy producer   warp 1: MMA issuer   warps 2..5: epilogue (same as above)
// ------------------------------------------------------------------------------------------------------------------
namespace {

struct StemRowParams {
  int Ho, Wo, total_rows, tiles_per_row, relu, stages;
  int R, chunks, total_items;            // work items = R consecutive output rows of one image (R = 1: single rows)
  __nv_bfloat16* pool;                   // POOL: (N, Ho/2, Wo/2, 64) max-pooled output
  int row_bytes, stage_bytes;            // one input row of the frame, 7 rows rounded up to 128 B
  int64_t image_bytes;                   // one frame image
  const uint8_t* x;
  __nv_bfloat16* y;
  const float* scale;
  const float* shift;
};

__device__ __forceinline__ void bulk_load(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               :: "r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(src)), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

constexpr int ROW_MAX_STAGES = 8;
constexpr int OUT_TILE = BM * BN * 2;        // 16 KB staging tile of the TMA output store

// POOL: the 3x3 / stride-2 / pad-1 max pooling that follows the stem (resnet.maxpool) is computed from the staged output
// rows: a CTA walks R consecutive output rows of one image (plus the row above them, recomputed and not stored), keeps the
// last three staged rows in a ring and emits pooled row y after output row 2y + 1.  Requires Wo == 128 (one tile per row).
template <bool POOL>
__global__ void __launch_bounds__(kThreads, 2)
conv_stem_rows_kernel(const __grid_constant__ CUtensorMap tm_b, const __grid_constant__ CUtensorMap tm_y,
                      const StemRowParams p) {
  constexpr int OUT_TILES = POOL ? 3 : 2;
  constexpr int TMEM_COLS = 2 * BN;
  constexpr uint32_t idesc = umma_idesc_bf16(BM, BN);
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_w = smem;
  uint8_t* smem_a = smem + W_BYTES;
  // two 16 KB staging tiles for the TMA output store (1024-byte aligned: SWIZZLE_128B), then the barriers
  uint8_t* smem_o = reinterpret_cast<uint8_t*>(
      (reinterpret_cast<uintptr_t>(smem_a + static_cast<size_t>(p.stages) * p.stage_bytes) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_o + OUT_TILES * OUT_TILE);
  uint64_t* w_full = bars;
  uint64_t* full = w_full + 1;                   // [ROW_MAX_STAGES]
  uint64_t* empty = full + ROW_MAX_STAGES;       // [ROW_MAX_STAGES]
  uint64_t* tmem_full = empty + ROW_MAX_STAGES;  // [2]
  uint64_t* tmem_empty = tmem_full + 2;          // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_b);
    tma_prefetch_desc(&tm_y);
    mbar_init(w_full, 1u);
    for (int i = 0; i < p.stages; ++i) { mbar_init(&full[i], 1u); mbar_init(&empty[i], 1u); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tmem_full[i], 1u); mbar_init(&tmem_empty[i], 128u); }
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);   // warp-uniform for the compiler (uniform registers)

  if (warp == 0) {
    if (lane == 0) {
      mbar_arrive_expect_tx(w_full, W_BYTES);
      for (int r = 0; r < ROWS; ++r) tma_load_2d(smem_w + r * W_SUB, &tm_b, w_full, r * 32, 0);
      int st = 0;
      uint32_t ph = 0;
      const uint32_t bytes = static_cast<uint32_t>(ROWS * p.row_bytes);
      for (int item = blockIdx.x; item < p.total_items; item += gridDim.x) {
        const int n = item / p.chunks, r0 = (item - n * p.chunks) * p.R;
        for (int ho = (POOL && r0 > 0) ? r0 - 1 : r0; ho < r0 + p.R; ++ho) {
          mbar_wait(&empty[st], ph ^ 1u);
          mbar_arrive_expect_tx(&full[st], bytes);
          bulk_load(smem_a + static_cast<size_t>(st) * p.stage_bytes,
                    p.x + n * p.image_bytes + static_cast<int64_t>(2 * ho) * p.row_bytes, bytes, &full[st]);
          if (++st == p.stages) { st = 0; ph ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    {   // whole warp: uniform bookkeeping; an elected lane issues (see conv_halo.cu)
      const uint64_t a_hi = umma_desc(0u, 128u, 0u);   // no swizzle: rows 16 B apart, 8-row groups 128 B, K chunks 16 B
      const uint64_t b_d0 = umma_desc(smem_u32(smem_w), 512u, 4u);
      mbar_wait(w_full, 0);
      int st = 0, buf = 0;
      uint32_t ph = 0, pbuf = 0;
      for (int item = blockIdx.x; item < p.total_items; item += gridDim.x) {
        const int r0 = (item % p.chunks) * p.R;
        for (int ho = (POOL && r0 > 0) ? r0 - 1 : r0; ho < r0 + p.R; ++ho) {
        mbar_wait(&full[st], ph);
        tc_fence_after();
        const uint32_t a_stage = smem_u32(smem_a + static_cast<size_t>(st) * p.stage_bytes);
        for (int t = 0; t < p.tiles_per_row; ++t) {
          mbar_wait(&tmem_empty[buf], pbuf ^ 1u);
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + buf * BN;
          const uint32_t a_tile = a_stage + t * (BM * 16);
          if (elect_one()) {
#pragma unroll
            for (int r = 0; r < ROWS; ++r) {
#pragma unroll
              for (int k = 0; k < 2; ++k)
                umma_bf16_ss(d_tmem, a_hi + (((a_tile + r * p.row_bytes + k * 32) & 0x3FFFFu) >> 4),
                             b_d0 + ((r * W_SUB + k * 32) >> 4), idesc, (r | k) != 0 ? 1u : 0u);
            }
            umma_commit(&tmem_full[buf]);
            if (t + 1 == p.tiles_per_row) umma_commit(&empty[st]);
          }
          __syncwarp();
          if ((buf ^= 1) == 0) pbuf ^= 1u;
        }
        if (++st == p.stages) { st = 0; ph ^= 1u; }
        }
      }
    }
  } else {
    // The 128 x 64 bf16 tile of an output row segment is 16 KB contiguous in global memory: the epilogue writes it into a
    // SWIZZLE_128B staging tile (16-byte chunk index ^ (row & 7): conflict-free) and ONE TMA store per tile moves it out,
    // instead of 16 warp-level 256-bit stores that each touch 32 different 128-byte lines (ncu: l1tex 82 % busy).
    const int quarter = warp & 3;
    const int lrow = quarter * 32 + lane;
    const bool issuer = threadIdx.x == 64;              // first epilogue thread (warp 2, lane 0)
    int buf = 0, ob = 0;
    uint32_t pbuf = 0;
    for (int item = blockIdx.x; item < p.total_items; item += gridDim.x) {
      const int n_img = item / p.chunks, r0 = (item - n_img * p.chunks) * p.R;
      for (int ho = (POOL && r0 > 0) ? r0 - 1 : r0; ho < r0 + p.R; ++ho) {
      const int row = n_img * p.Ho + ho;
      for (int t = 0; t < p.tiles_per_row; ++t) {
        const int m0 = row * p.Wo + t * BM;
        if (issuer) bulk_wait_group_read<OUT_TILES - 1>();   // the store that last read staging tile `ob` has drained it
        named_bar_sync(1, 128);
        mbar_wait(&tmem_full[buf], pbuf);
        tc_fence_after();
        const uint32_t t_row = tmem_base + buf * BN + (static_cast<uint32_t>(quarter * 32) << 16);
        uint8_t* orow = smem_o + ob * OUT_TILE + lrow * 128;
#pragma unroll
        for (int c0 = 0; c0 < BN; c0 += 16) {
          uint32_t v[16];
          tmem_ld_x16(t_row + c0, v);
          tmem_ld_wait();
          float f[16];
#pragma unroll
          for (int j = 0; j < 16; j += 4) {
            const float4 sc = __ldg(reinterpret_cast<const float4*>(p.scale + c0 + j));
            const float4 sh = __ldg(reinterpret_cast<const float4*>(p.shift + c0 + j));
            f[j] = fmaf(__uint_as_float(v[j]), sc.x, sh.x);
            f[j + 1] = fmaf(__uint_as_float(v[j + 1]), sc.y, sh.y);
            f[j + 2] = fmaf(__uint_as_float(v[j + 2]), sc.z, sh.z);
            f[j + 3] = fmaf(__uint_as_float(v[j + 3]), sc.w, sh.w);
          }
          if (p.relu) {
#pragma unroll
            for (int j = 0; j < 16; ++j) f[j] = fmaxf(f[j], 0.f);
          }
          const int ch = c0 >> 3;                       // 16-byte chunk index of channel c0 in the 128-byte row
          *reinterpret_cast<uint4*>(orow + (((ch) ^ (lrow & 7)) << 4)) =
              make_uint4(pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]), pack_bf16x2(f[4], f[5]), pack_bf16x2(f[6], f[7]));
          *reinterpret_cast<uint4*>(orow + (((ch + 1) ^ (lrow & 7)) << 4)) =
              make_uint4(pack_bf16x2(f[8], f[9]), pack_bf16x2(f[10], f[11]), pack_bf16x2(f[12], f[13]), pack_bf16x2(f[14], f[15]));
        }
        tc_fence_before();
        mbar_arrive(&tmem_empty[buf]);
        fence_proxy_async_smem();                       // generic-proxy writes visible to the TMA (async proxy)
        named_bar_sync(1, 128);
        if (issuer) {
          if (!POOL || ho >= r0) tma_store_2d(&tm_y, smem_o + ob * OUT_TILE, 0, m0);   // the recomputed row is not stored
          bulk_commit_group();                          // (an empty group keeps the ring's group count uniform)
        }
        if (POOL && (ho & 1) && ho > r0) {
          // pooled row y = max over output rows 2y-1 .. 2y+1 (ring tiles ob+1, ob+2, ob) and columns 2x-1 .. 2x+1.
          // thread = one 16-byte channel chunk of four consecutive pooled pixels: the eight lanes of a quarter warp read the
          // eight chunks of one staged pixel (conflict-free under the 128-byte swizzle) and write 128 contiguous bytes;
          // same __hmax2 as dt_maxpool3x3s2, max is exact in any order
          const int ch = lrow & 7, px0 = (lrow >> 3) * 4;
          const __nv_bfloat162 ninf = __float2bfloat162_rn(-INFINITY);
          __nv_bfloat162 col[9][4];                     // column maxima of staged columns 2*px0 - 1 .. 2*px0 + 7
#pragma unroll
          for (int cx = 0; cx < 9; ++cx) {
            const int sx = 2 * px0 - 1 + cx;
#pragma unroll
            for (int u = 0; u < 4; ++u) col[cx][u] = ninf;
            if (sx < 0) continue;
#pragma unroll
            for (int dy = 0; dy < 3; ++dy) {
              if (dy == 0 && ho < 2) continue;          // row -1 of the image
              const uint4 v = *reinterpret_cast<const uint4*>(smem_o + ((ob + 1 + dy) % 3) * OUT_TILE + sx * 128 +
                                                              ((ch ^ (sx & 7)) << 4));
              const __nv_bfloat162* pv = reinterpret_cast<const __nv_bfloat162*>(&v);
#pragma unroll
              for (int u = 0; u < 4; ++u) col[cx][u] = __hmax2(col[cx][u], pv[u]);
            }
          }
          uint8_t* dst = reinterpret_cast<uint8_t*>(p.pool) +
                         (((static_cast<int64_t>(n_img) * (p.Ho >> 1) + (ho >> 1)) * (p.Wo >> 1) + px0) * BN) * 2 + ch * 16;
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            __nv_bfloat162 m[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) m[u] = __hmax2(__hmax2(col[2 * i][u], col[2 * i + 1][u]), col[2 * i + 2][u]);
            *reinterpret_cast<uint4*>(dst + i * BN * 2) = *reinterpret_cast<const uint4*>(m);
          }
        }
        if (++ob == OUT_TILES) ob = 0;
        if ((buf ^= 1) == 0) pbuf ^= 1u;
      }
      }
    }
    if (issuer) bulk_wait_group<0>();                   // all stores complete before the CTA exits
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

}  // namespace

// x: zero-bordered (N, H+6, W+8, 4) bf16; w: bf16 [64][256] with k = r*32 + s*4 + c; y: (N, H/2, W/2, 64) bf16.
// Returns DT_ERR_UNSUPPORTED when the output grid cannot be tiled into 128-pixel boxes.
// pooled != nullptr: also write the 3x3 / s2 / pad-1 max pooling of y, (N, H/4, W/4, 64) (DT_ERR_UNSUPPORTED unless W/2 == 128).
int dt_conv_stem(const dt_conv_desc* d, const void* x, const void* w, const float* scale, const float* shift, void* y,
                 void* pooled, cudaStream_t s) {
  if (d->C_out != BN || d->H % 2 || d->W % 2) return DT_ERR_UNSUPPORTED;
  const int Ho = d->H / 2, Wo = d->W / 2;
  int bw, bh, bn;
  if (Wo >= BM) {
    if (Wo % BM) return DT_ERR_UNSUPPORTED;
    bw = BM; bh = 1; bn = 1;
  } else {
    if (BM % Wo) return DT_ERR_UNSUPPORTED;
    const int rows = BM / Wo;
    bw = Wo;
    if (Ho >= rows) { if (Ho % rows) return DT_ERR_UNSUPPORTED; bh = rows; bn = 1; }
    else { if (rows % Ho) return DT_ERR_UNSUPPORTED; bh = Ho; bn = rows / Ho; }
  }
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once_fn;
  std::call_once(once_fn, [] {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(ptr);
  });
  DT_REQUIRE(fn != nullptr, DT_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  const uint64_t Hp = d->H + 6, Wp = d->W + 8;
  CUtensorMap tm_a, tm_b;
  {
    const cuuint64_t dims[5] = {32, static_cast<cuuint64_t>(Wo), static_cast<cuuint64_t>(Ho), ROWS,
                                static_cast<cuuint64_t>(d->N)};
    const cuuint64_t str[4] = {16, 2 * Wp * 8, Wp * 8, Hp * Wp * 8};
    const cuuint32_t box[5] = {32, static_cast<cuuint32_t>(bw), static_cast<cuuint32_t>(bh), 1,
                               static_cast<cuuint32_t>(bn)};
    const cuuint32_t es[5] = {1, 1, 1, 1, 1};
    CUresult r = fn(&tm_a, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(x), dims, str, box, es,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    DT_REQUIRE(r == CUDA_SUCCESS, DT_ERR_CUDA, "stem im2col tensor map: CUresult %d", static_cast<int>(r));
  }
  {
    const cuuint64_t dims[2] = {256, BN};
    const cuuint64_t str[1] = {512};
    const cuuint32_t box[2] = {32, BN};
    const cuuint32_t es[2] = {1, 1};
    CUresult r = fn(&tm_b, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(w), dims, str, box, es,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    DT_REQUIRE(r == CUDA_SUCCESS, DT_ERR_CUDA, "stem weight tensor map: CUresult %d", static_cast<int>(r));
  }
  const bool want_pool = pooled != nullptr;
  if (want_pool && !(Wo == BM && Ho % 8 == 0 && !(d->flags & DT_CONV_NO_HALO))) return DT_ERR_UNSUPPORTED;
  if (Wo % BM == 0 && !(d->flags & DT_CONV_NO_HALO)) {
    // row form: the im2col operand is read in place from the raw input rows (no-swizzle descriptors)
    StemRowParams q;
    q.Ho = Ho; q.Wo = Wo; q.total_rows = d->N * Ho; q.tiles_per_row = Wo / BM; q.relu = d->relu;
    q.row_bytes = static_cast<int>(Wp * 8);
    q.stage_bytes = (ROWS * q.row_bytes + 127) / 128 * 128;
    q.image_bytes = static_cast<int64_t>(Hp) * Wp * 8;
    q.pool = static_cast<__nv_bfloat16*>(pooled);
    const int out_tiles = want_pool ? 3 : 2;
    // Two CTAs per SM when the stages fit in half the shared memory (T <= 256): the 14 MMAs of a row chain into one
    // accumulator (138 clk each, profiles/r02_mma_rate.txt), a second CTA's chain and epilogue overlap it.
    int ctas = 2;
    q.stages = (111 * 1024 - W_BYTES - out_tiles * OUT_TILE - 2048 - 256) / q.stage_bytes;
    if (q.stages < (want_pool ? 2 : 3) || getenv("DT_STEM_ONE_CTA")) {
      ctas = 1;
      q.stages = (200 * 1024 - W_BYTES - out_tiles * OUT_TILE - 1024) / q.stage_bytes;
    }
    if (q.stages > ROW_MAX_STAGES) q.stages = ROW_MAX_STAGES;
    if (ctas == 2 && q.stages > 4) q.stages = 4;
    if (q.stages >= 2 && (ROWS * q.row_bytes) % 16 == 0 && ROWS * q.row_bytes < (1 << 20)) {
      q.x = static_cast<const uint8_t*>(x);
      q.y = static_cast<__nv_bfloat16*>(y);
      q.scale = scale; q.shift = shift;
      const int slots = ctas * dt_num_sms();
      q.R = 1;
      if (want_pool) {   // rows per item: (items per CTA) x (R + 1 rows, one recomputed) is the cost of the slowest CTA
        long best = -1;
        for (int R = 32; R >= 8; R >>= 1) {
          if (Ho % R) continue;
          const long items = static_cast<long>(d->N) * (Ho / R);
          const long cost = ((items + slots - 1) / slots) * (R + 1);
          if (best < 0 || cost < best) { best = cost; q.R = R; }
        }
      }
      q.chunks = Ho / q.R;
      q.total_items = d->N * q.chunks;
      CUtensorMap tm_y;
      {
        const cuuint64_t ydims[2] = {BN, static_cast<cuuint64_t>(q.total_rows) * Wo};
        const cuuint64_t ystr[1] = {BN * 2};
        const cuuint32_t ybox[2] = {BN, BM};
        const cuuint32_t yes[2] = {1, 1};
        CUresult r = fn(&tm_y, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, y, ydims, ystr, ybox, yes, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        DT_REQUIRE(r == CUDA_SUCCESS, DT_ERR_CUDA, "stem output tensor map: CUresult %d", static_cast<int>(r));
      }
      const int smem_rows = W_BYTES + q.stages * q.stage_bytes + 1024 + out_tiles * OUT_TILE + 1024 + 256;
      static std::once_flag once_rows;
      static cudaError_t attr_rows = cudaSuccess;
      std::call_once(once_rows, [] {
        attr_rows = cudaFuncSetAttribute(conv_stem_rows_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 225 * 1024);
        if (attr_rows == cudaSuccess)
          attr_rows = cudaFuncSetAttribute(conv_stem_rows_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 225 * 1024);
      });
      DT_CUDA(attr_rows);
      const int grid_rows = q.total_items < slots ? q.total_items : slots;
      if (want_pool) conv_stem_rows_kernel<true><<<grid_rows, kThreads, smem_rows, s>>>(tm_b, tm_y, q);
      else conv_stem_rows_kernel<false><<<grid_rows, kThreads, smem_rows, s>>>(tm_b, tm_y, q);
      DT_LAUNCH_CHECK();
      return DT_OK;
    }
  }
  if (want_pool) return DT_ERR_UNSUPPORTED;
  StemParams p;
  p.Ho = Ho; p.Wo = Wo;
  p.M_total = d->N * Ho * Wo;
  p.total_tiles = (p.M_total + BM - 1) / BM;
  p.relu = d->relu;
  p.y = static_cast<__nv_bfloat16*>(y);
  p.scale = scale; p.shift = shift;
  constexpr int smem = W_BYTES + STAGES * A_STAGE + 1024 + 256;
  static std::once_flag once;
  static cudaError_t attr_err = cudaSuccess;
  std::call_once(once, [] {
    attr_err = cudaFuncSetAttribute(conv_stem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  });
  DT_CUDA(attr_err);
  const int grid = p.total_tiles < dt_num_sms() ? p.total_tiles : dt_num_sms();
  conv_stem_kernel<<<grid, kThreads, smem, s>>>(tm_a, tm_b, p);
  DT_LAUNCH_CHECK();
  return DT_OK;
}


// ------------------------------------------------------------------------------------------------------------------
// Weight gradient of the stem on the tensor cores.
//
//   dW[co][ci][r][s] = sum over output pixels p of gy[p][co] * x_pad[2*ho + r][2*wo + s][ci]
//
// The SAME overlapping-stride im2col tensor map as the forward kernel delivers, per filter row r, a tile of
// 128 pixels x 32 elements (8 pixels x 4 channels under the filter row) as 64-byte rows; read as an **MN-major**
// SWIZZLE_64B operand (N = 32, K = pixel) it multiplies the gy tile (MN-major SWIZZLE_128B, M = co, K = pixel):
// D_r[co][e] += gy^T . xcol_r.  Seven accumulators of 128 x 32 fp32 stay in TMEM across all tiles of a CTA;
// partial[split][r][co][e] is then reduced in a fixed order into the OIHW gradient (e = s*4 + ci; s = 7 and ci >= C_in
// are padding).   warp 0: TMA producer   warp 1: MMA issuer   warps 2..5: epilogue
// ------------------------------------------------------------------------------------------------------------------
namespace {

constexpr int WG_A = 128 * 128;                // gy tile: 128 pixels x 64 channels
constexpr int WG_BSUB = 128 * 64;              // one filter row of the im2col tile: 128 pixels x 64 B
constexpr int WG_STAGE = WG_A + ROWS * WG_BSUB;   // 72 KB
constexpr int WG_STAGES = 3;
constexpr int WG_TMEM = 256;                   // 7 x 32 columns

struct StemWgParams {
  int Ho, Wo, total_tiles, tiles_per_cta;
  float* partial;                              // [splits][7][64][32]
};

__device__ __forceinline__ uint64_t umma_desc_mn_sw(uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(layout & 7u) << 61;
  return d;
}

__global__ void __launch_bounds__(kThreads, 1)
conv_stem_wgrad_kernel(const __grid_constant__ CUtensorMap tm_g, const __grid_constant__ CUtensorMap tm_x,
                       const StemWgParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + WG_STAGES * WG_STAGE);
  uint64_t* full = bars;
  uint64_t* empty = full + WG_STAGES;
  uint64_t* done = empty + WG_STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tile_begin = blockIdx.x * p.tiles_per_cta;
  const int tile_end = min(p.total_tiles, tile_begin + p.tiles_per_cta);
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_g);
    tma_prefetch_desc(&tm_x);
    for (int i = 0; i < WG_STAGES; ++i) { mbar_init(&full[i], 1u); mbar_init(&empty[i], 1u); }
    mbar_init(done, 1u);
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, WG_TMEM);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);   // warp-uniform for the compiler (uniform registers)

  if (warp == 0) {
    if (lane == 0) {
      int st = 0;
      uint32_t ph = 0;
      for (int tile = tile_begin; tile < tile_end; ++tile) {
        const int m0 = tile * BM;
        const int w0 = m0 % p.Wo, h0 = (m0 / p.Wo) % p.Ho, n0 = m0 / (p.Wo * p.Ho);
        uint8_t* stage = smem + st * WG_STAGE;
        mbar_wait(&empty[st], ph ^ 1u);
        mbar_arrive_expect_tx(&full[st], WG_STAGE);
        tma_load_4d(stage, &tm_g, &full[st], 0, w0, h0, n0);
        for (int r = 0; r < ROWS; ++r)
          tma_load_5d(stage + WG_A + r * WG_BSUB, &tm_x, &full[st], 0, w0, h0, r, n0);
        if (++st == WG_STAGES) { st = 0; ph ^= 1u; }
      }
    }
  } else if (warp == 1) {
    {   // whole warp: warp-uniform bookkeeping (uniform registers); an elected lane issues (see conv_halo.cu)
      // bf16, A and B MN-major, M = 128 (rows 64..127 read whatever follows the gy tile and are discarded), N = 32
      constexpr uint32_t idesc_w = umma_idesc_bf16(128, 32) | (1u << 15) | (1u << 16);
      const uint64_t a_hi = umma_desc_mn_sw(WG_A, 1024u, 2u);      // 128 B rows, 8-row K groups 1024 B apart
      const uint64_t b_hi = umma_desc_mn_sw(16u, 512u, 4u);        // 64 B rows, 8-row K groups 512 B apart
      int st = 0;
      uint32_t ph = 0, accum = 0;
      for (int tile = tile_begin; tile < tile_end; ++tile) {
        mbar_wait(&full[st], ph);
        tc_fence_after();
        const uint32_t a_addr = smem_u32(smem + st * WG_STAGE);
        const uint32_t b_addr = a_addr + WG_A;
        if (elect_one()) {
#pragma unroll 1
          for (int r = 0; r < ROWS; ++r) {
#pragma unroll
            for (int k8 = 0; k8 < 8; ++k8)
              umma_bf16_ss(tmem_base + r * 32, a_hi + ((a_addr + k8 * 2048) >> 4),
                           b_hi + ((b_addr + r * WG_BSUB + k8 * 1024) >> 4), idesc_w, (accum | k8) != 0 ? 1u : 0u);
          }
          umma_commit(&empty[st]);
          if (tile + 1 == tile_end) umma_commit(done);
        }
        __syncwarp();
        accum = 1;
        if (++st == WG_STAGES) { st = 0; ph ^= 1u; }
      }
      if (tile_begin >= tile_end && elect_one()) umma_commit(done);
    }
  } else {
    const int quarter = warp & 3;
    const int co = quarter * 32 + lane;
    mbar_wait(done, 0);
    tc_fence_after();
    for (int r = 0; r < ROWS; ++r) {
#pragma unroll
      for (int c0 = 0; c0 < 32; c0 += 16) {
        uint32_t v[16];
        tmem_ld_x16(tmem_base + r * 32 + c0 + (static_cast<uint32_t>(quarter * 32) << 16), v);
        tmem_ld_wait();
        if (co < BN) {
          float4* dst = reinterpret_cast<float4*>(p.partial + ((static_cast<int64_t>(blockIdx.x) * ROWS + r) * BN + co) * 32 + c0);
#pragma unroll
          for (int j = 0; j < 4; ++j)
            dst[j] = make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]), __uint_as_float(v[4 * j + 2]),
                                 __uint_as_float(v[4 * j + 3]));
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, WG_TMEM);
  }
}

// dw[co][ci][r][s] = sum over splits of partial[split][r][co][s*4 + ci]
__global__ void stem_wgrad_reduce_kernel(const float* __restrict__ partial, int splits, int C_in, float* __restrict__ dw) {
  const int total = BN * C_in * 49;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int s = i % 7, r = (i / 7) % 7, ci = (i / 49) % C_in, co = i / (49 * C_in);
    float acc = 0.f;
    for (int sp = 0; sp < splits; ++sp) acc += partial[((static_cast<int64_t>(sp) * ROWS + r) * BN + co) * 32 + s * 4 + ci];
    dw[i] = acc;
  }
}

int stem_tiling(int Ho, int Wo, int* bw, int* bh, int* bn) {
  if (Wo >= BM) {
    if (Wo % BM) return DT_ERR_UNSUPPORTED;
    *bw = BM; *bh = 1; *bn = 1;
    return DT_OK;
  }
  if (BM % Wo) return DT_ERR_UNSUPPORTED;
  const int rows = BM / Wo;
  *bw = Wo;
  if (Ho >= rows) { if (Ho % rows) return DT_ERR_UNSUPPORTED; *bh = rows; *bn = 1; }
  else { if (rows % Ho) return DT_ERR_UNSUPPORTED; *bh = Ho; *bn = rows / Ho; }
  return DT_OK;
}

int stem_wgrad_splits(int total_tiles, int* tiles_per_cta) {
  int splits = dt_num_sms();           // one wave of single-CTA-per-SM (216 KB of shared memory each)
  if (splits > total_tiles) splits = total_tiles;
  *tiles_per_cta = (total_tiles + splits - 1) / splits;
  return (total_tiles + *tiles_per_cta - 1) / *tiles_per_cta;
}

}  // namespace

extern "C" int64_t dt_stem_wgrad_tc_workspace(int N, int H, int W) {
  int bw, bh, bn;
  if (N <= 0 || H <= 0 || W <= 0 || H % 2 || W % 2 || stem_tiling(H / 2, W / 2, &bw, &bh, &bn) != DT_OK || N % bn)
    return DT_ERR_UNSUPPORTED;
  int per;
  const int splits = stem_wgrad_splits(N * (H / 2) * (W / 2) / BM, &per);
  return static_cast<int64_t>(splits) * ROWS * BN * 32 * static_cast<int64_t>(sizeof(float));
}

// x_pad: zero-bordered (N, H+6, W+8, 4) bf16 frame; gy: (N, H/2, W/2, 64) bf16; dw: fp32 (64, C_in, 7, 7).
extern "C" int dt_stem_wgrad_tc(const void* x_pad, const void* gy, int N, int H, int W, int C_in, float* dw_oihw,
                                float* workspace, int64_t workspace_bytes, dt_stream_t stream) {
  DT_ARCH_GUARD();
  int bw, bh, bn;
  if (N <= 0 || H <= 0 || W <= 0 || H % 2 || W % 2 || C_in < 1 || C_in > 4 ||
      stem_tiling(H / 2, W / 2, &bw, &bh, &bn) != DT_OK || N % bn) {
    dt_set_error("dt_stem_wgrad_tc: unsupported shape N=%d H=%d W=%d C_in=%d", N, H, W, C_in);
    return DT_ERR_UNSUPPORTED;
  }
  const int Ho = H / 2, Wo = W / 2;
  StemWgParams p;
  p.Ho = Ho; p.Wo = Wo;
  p.total_tiles = N * Ho * Wo / BM;
  const int splits = stem_wgrad_splits(p.total_tiles, &p.tiles_per_cta);
  const int64_t need = static_cast<int64_t>(splits) * ROWS * BN * 32 * static_cast<int64_t>(sizeof(float));
  DT_REQUIRE(workspace != nullptr && workspace_bytes >= need, DT_ERR_BAD_SHAPE,
             "dt_stem_wgrad_tc: workspace of %lld bytes needed", static_cast<long long>(need));
  DT_REQUIRE((reinterpret_cast<uintptr_t>(x_pad) | reinterpret_cast<uintptr_t>(gy) | reinterpret_cast<uintptr_t>(workspace)) % 16 == 0,
             DT_ERR_BAD_ALIGN, "dt_stem_wgrad_tc: tensors must be 16-byte aligned");
  p.partial = workspace;
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once_fn;
  std::call_once(once_fn, [] {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(ptr);
  });
  DT_REQUIRE(fn != nullptr, DT_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  const uint64_t Hp = H + 6, Wp = W + 8;
  CUtensorMap tm_x, tm_g;
  {
    const cuuint64_t dims[5] = {32, static_cast<cuuint64_t>(Wo), static_cast<cuuint64_t>(Ho), ROWS, static_cast<cuuint64_t>(N)};
    const cuuint64_t str[4] = {16, 2 * Wp * 8, Wp * 8, Hp * Wp * 8};
    const cuuint32_t box[5] = {32, static_cast<cuuint32_t>(bw), static_cast<cuuint32_t>(bh), 1, static_cast<cuuint32_t>(bn)};
    const cuuint32_t es[5] = {1, 1, 1, 1, 1};
    CUresult r = fn(&tm_x, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(x_pad), dims, str, box, es,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    DT_REQUIRE(r == CUDA_SUCCESS, DT_ERR_CUDA, "stem wgrad im2col tensor map: CUresult %d", static_cast<int>(r));
  }
  {
    const cuuint64_t dims[4] = {BN, static_cast<cuuint64_t>(Wo), static_cast<cuuint64_t>(Ho), static_cast<cuuint64_t>(N)};
    const cuuint64_t str[3] = {BN * 2, static_cast<cuuint64_t>(Wo) * BN * 2, static_cast<cuuint64_t>(Ho) * Wo * BN * 2};
    const cuuint32_t box[4] = {BN, static_cast<cuuint32_t>(bw), static_cast<cuuint32_t>(bh), static_cast<cuuint32_t>(bn)};
    const cuuint32_t es[4] = {1, 1, 1, 1};
    CUresult r = fn(&tm_g, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(gy), dims, str, box, es,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    DT_REQUIRE(r == CUDA_SUCCESS, DT_ERR_CUDA, "stem wgrad gy tensor map: CUresult %d", static_cast<int>(r));
  }
  constexpr int smem = WG_STAGES * WG_STAGE + 1024 + 256;
  static std::once_flag once;
  static cudaError_t attr_err = cudaSuccess;
  std::call_once(once, [] {
    attr_err = cudaFuncSetAttribute(conv_stem_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  });
  DT_CUDA(attr_err);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  conv_stem_wgrad_kernel<<<splits, kThreads, smem, s>>>(tm_g, tm_x, p);
  DT_LAUNCH_CHECK();
  stem_wgrad_reduce_kernel<<<(BN * C_in * 49 + 255) / 256, 256, 0, s>>>(workspace, splits, C_in, dw_oihw);
  DT_LAUNCH_CHECK();
  return DT_OK;
}
