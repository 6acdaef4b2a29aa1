// Weight gradient of the 3x3 (stride 1 or 2, pad 1) and 1x1 / stride-2 convolutions on the sm_100a tensor cores
// (tcgen05, TMEM accumulators).
//
//   dW[co][ci][r][s] = sum over output pixels p of  gy[p][co] * x[stride * p + (r - pad, s - pad)][ci]
//
// GEMM view per filter tap: D_tap[M = co][N = ci] += A[co][K = pixel] * B_tap[ci][K = pixel].  Both operands are read
// straight from the NHWC bf16 tensors as **MN-major** UMMA operands: a TMA box (64 channels x pixels, SWIZZLE_128B)
// lands in shared memory as rows of 128 B (one pixel each) - exactly the MN-major canonical atom (64 MN elements x 8 K
// rows), with the pixel index as K.  No transposed copies of activations or gradients are ever made.
//   A  = gy tile: 8 (w) x 16 (h) pixels (= K 128) x 64-channel slabs, two slabs (leading-byte-offset apart) give M = 128
//   B  = x halo patch: (16+2) x (8+2) pixels x one 64-channel slab, loaded ONCE per tile; the taps of a tap group read it
//        in place through descriptors whose start is shifted by (dr*10 + ds) pixel rows and whose stride-byte-offset
//        (distance between the 8-pixel K groups = image rows) is 10 pixel rows - the same absolute-address-swizzle
//        property the forward halo kernel relies on (profiles/r01_halo_descriptor_experiment.txt).
//        stride 1: tap groups {0..4}, {5..8} of the one patch.   stride 2: one tap group per input parity plane
//        x[2i+pr][2j+pc] (TMA traversal stride 2): plane (1,1) serves taps (0,0) (0,2) (2,0) (2,2), (1,0) -> (0,1) (2,1),
//        (0,1) -> (1,0) (1,2), (0,0) -> (1,1); the 1x1 / stride-2 downsample is plane (0,0) alone.
// Channel counts below 64 are handled by TMA out-of-bounds zero fill (box wider than the tensor), so one kernel
// configuration (M 128, N 64) serves every layer; rows / columns beyond C_out / C_in are never written back.
// Images smaller than 16 rows (8x8 at the bottom of the encoder) put two images in one tile.
//
// Work decomposition: job = (128-wide co block, 64-wide ci slab, tap group); the pixel tiles are split over `splits`
// CTAs per job; every CTA keeps its <= 5 tap accumulators of 128 x 64 fp32 in TMEM across all its tiles and writes them
// once, with coalesced stores, to partial[split][tap][co][ci]; a second kernel sums the splits in a fixed order into the
// fp32 OIHW gradient (deterministic: no atomics).
//   warp 0: TMA producer   warp 1: TMEM alloc + MMA issuer   warps 2..5: epilogue
//
// Replaces the cuDNN wgrad kernels autograd reaches from SemSegment.training_step
// (deadtrees/network/segmodel.py:210-229).
#include <cstring>
#include <mutex>

#include "common.cuh"

namespace {

constexpr int TW = 8, TH = 16, PITCH = TW + 2;
constexpr int kThreads = 192;
constexpr int A_SLAB = 128 * 128;            // one 64-channel slab of a 128-pixel gy tile
constexpr int A_STAGE = 2 * A_SLAB;          // 32 KB
constexpr int B_STAGE = 26 * 1024;           // >= 2 images x 10 x 10 pixels x 128 B
constexpr int STAGES = 3;
constexpr int TMEM_COLS = 512;
constexpr int NCOL = 64;                     // UMMA N (one ci slab)
constexpr int MAX_GROUPS = 4, MAX_TAPS = 5;

struct WgGroup {
  int ntaps, pr, pc, pad_;
  int tap[MAX_TAPS];                // filter tap r*S+s this accumulator belongs to
  int off[MAX_TAPS];                // pixel-row offset of the tap's window inside the patch
};

struct WgParams {
  int N, H, W, C_in, C_out;         // H, W: OUTPUT (= gy) size
  int RS, stride;
  int th_img, imgs;                 // image rows per tile (8 or 16), images per tile (2 or 1)
  int tiles_w, tiles_h, total_tiles;
  int ci_slabs, ngroups, jobs, tiles_per_cta;
  int patch_bytes;
  float* partial;                   // [splits][RS][C_out][C_in]
  WgGroup grp[MAX_GROUPS];
};

// MN-major, 128-byte-swizzled operand descriptor: rows of 64 bf16 (one K index each), 8-row K groups `sbo` bytes
// apart, 64-element MN blocks `lbo` bytes apart.
__device__ __forceinline__ uint64_t umma_desc_mn(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;   // SWIZZLE_128B
  return d;
}

__global__ void __launch_bounds__(kThreads, 1)
conv_wgrad_kernel(const __grid_constant__ CUtensorMap tm_g, const __grid_constant__ CUtensorMap tm_x, const WgParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + STAGES * A_STAGE;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_b + STAGES * B_STAGE);
  uint64_t* full = bars;                 // [STAGES]
  uint64_t* empty = full + STAGES;       // [STAGES]
  uint64_t* done = empty + STAGES;       // [1]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int job = blockIdx.x % p.jobs, split = blockIdx.x / p.jobs;
  const int gi = job % p.ngroups;
  const int ci_slab = (job / p.ngroups) % p.ci_slabs;
  const int co_block = (job / p.ngroups) / p.ci_slabs;
  const WgGroup& grp = p.grp[gi];
  const int ntaps = grp.ntaps;
  const int tile_begin = split * p.tiles_per_cta;
  const int tile_end = min(p.total_tiles, tile_begin + p.tiles_per_cta);
  const int a_slabs = min(2, (p.C_out - co_block * 128 + 63) / 64);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_g);
    tma_prefetch_desc(&tm_x);
    for (int i = 0; i < STAGES; ++i) { mbar_init(&full[i], 1u); mbar_init(&empty[i], 1u); }
    mbar_init(done, 1u);
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
      int st = 0;
      uint32_t ph = 0;
      const int s = p.stride;
      for (int tile = tile_begin; tile < tile_end; ++tile) {
        int m = tile;
        const int w0 = (m % p.tiles_w) * TW; m /= p.tiles_w;
        const int h0 = (m % p.tiles_h) * p.th_img;
        const int n0 = (m / p.tiles_h) * p.imgs;
        mbar_wait(&empty[st], ph ^ 1u);
        mbar_arrive_expect_tx(&full[st], a_slabs * A_SLAB + p.patch_bytes);
        for (int sl = 0; sl < a_slabs; ++sl)
          tma_load_4d(smem_a + st * A_STAGE + sl * A_SLAB, &tm_g, &full[st], co_block * 128 + sl * 64, w0, h0, n0);
        // patch origin (h0 - 1, w0 - 1) in the coordinates of the (parity plane of the) input
        tma_load_4d(smem_b + st * B_STAGE, &tm_x, &full[st], ci_slab * 64, s * (w0 - 1) + grp.pc, s * (h0 - 1) + grp.pr, n0);
        if (++st == STAGES) { st = 0; ph ^= 1u; }
      }
    }
  } else if (warp == 1) {
    {   // whole warp: warp-uniform bookkeeping (uniform registers); an elected lane issues (see conv_halo.cu)
      // bf16 x bf16 -> fp32, A and B MN-major (bits 15 / 16), M = 128, N = 64
      constexpr uint32_t idesc = umma_idesc_bf16(128, NCOL) | (1u << 15) | (1u << 16);
      const uint64_t a_hi = umma_desc_mn(0u, A_SLAB, 1024u);
      const uint64_t b_hi = umma_desc_mn(0u, 16u, PITCH * 128u);
      // pixel-row offset of K step k8 (two image rows of 8 pixels) inside the patch
      uint32_t krow[8];
#pragma unroll
      for (int k8 = 0; k8 < 8; ++k8) {
        const int img = (2 * k8) / p.th_img, row = (2 * k8) % p.th_img;
        krow[k8] = static_cast<uint32_t>((img * (p.th_img + 2) + row) * PITCH);
      }
      int st = 0;
      uint32_t ph = 0, accum = 0;
      for (int tile = tile_begin; tile < tile_end; ++tile) {
        mbar_wait(&full[st], ph);
        tc_fence_after();
        const uint64_t a_d = a_hi + (smem_u32(smem_a + st * A_STAGE) >> 4);
        const uint32_t b_addr = smem_u32(smem_b + st * B_STAGE);
        if (elect_one()) {
#pragma unroll 1
          for (int t = 0; t < ntaps; ++t) {
            const uint32_t toff = static_cast<uint32_t>(grp.off[t]);
            const uint32_t d_tmem = tmem_base + t * NCOL;
#pragma unroll
            for (int k8 = 0; k8 < 8; ++k8) {
              const uint64_t b_d = b_hi + ((b_addr + (krow[k8] + toff) * 128u) >> 4);
              umma_bf16_ss(d_tmem, a_d + ((k8 * 2048) >> 4), b_d, idesc, (accum | k8) != 0 ? 1u : 0u);
            }
          }
          umma_commit(&empty[st]);
          if (tile + 1 == tile_end) umma_commit(done);
        }
        __syncwarp();
        accum = 1;
        if (++st == STAGES) { st = 0; ph ^= 1u; }
      }
      if (tile_begin >= tile_end && elect_one()) umma_commit(done);
    }
  } else {
    const int quarter = warp & 3;
    const int co = co_block * 128 + quarter * 32 + lane;
    mbar_wait(done, 0);
    tc_fence_after();
    for (int t = 0; t < ntaps; ++t) {
      float* dst = p.partial + ((static_cast<int64_t>(split) * p.RS + grp.tap[t]) * p.C_out + co) * p.C_in + ci_slab * 64;
#pragma unroll
      for (int c0 = 0; c0 < NCOL; c0 += 16) {
        uint32_t v[16];
        tmem_ld_x16(tmem_base + t * NCOL + c0 + (static_cast<uint32_t>(quarter * 32) << 16), v);
        tmem_ld_wait();
        if (co < p.C_out) {
#pragma unroll
          for (int j = 0; j < 16; j += 4) {   // C_in is a multiple of 4: whole float4s are inside or outside
            if (ci_slab * 64 + c0 + j < p.C_in)
              *reinterpret_cast<float4*>(dst + c0 + j) = make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]),
                                                                     __uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]));
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

// dw[co][ci][tap] = sum over splits (fixed order) of partial[split][tap][co][ci]
// Block (32, L): 32 float4 columns (4 consecutive ci) x L split lanes.  Lane l adds the splits l, l + L, l + 2L, ... with four
// independent accumulators (loads in flight together), the lanes are combined through shared memory in lane order: the
// order of every addition is a function of (splits, L) only - deterministic - while small layers with many splits (a
// 64 x 64 layer has 147 partials of 37 k floats) get L times the loads in flight of a thread-per-element loop.
__global__ void wgrad_reduce_kernel(const float* __restrict__ partial, int splits, int RS, int C_out, int C_in, int L,
                                    float* __restrict__ dw) {
  extern __shared__ float4 wr_sh[];                      // [blockDim.y][32]
  const int64_t plane = static_cast<int64_t>(C_out) * C_in;
  const int64_t plane4 = plane / 4, total4 = plane4 * RS;
  const int G = blockDim.y / L;                          // column groups of a block (L split lanes each)
  const int lane = threadIdx.y % L, group = threadIdx.y / L, col = threadIdx.x;
  const float4* p4 = reinterpret_cast<const float4*>(partial);
  const int64_t per_block = static_cast<int64_t>(G) * 32;
  for (int64_t base = blockIdx.x * per_block; base < total4; base += gridDim.x * per_block) {
    const int64_t i = base + group * 32 + col;
    const bool live = i < total4;
    const int64_t c4 = live ? i % plane4 : 0;           // float4 index inside a tap plane (ci fastest: coalesced reads)
    const int tap = live ? static_cast<int>(i / plane4) : 0;
    float4 acc[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) acc[u] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (live) {
      int sp = lane;
      for (; sp + 3 * L < splits; sp += 4 * L) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const float4 v = p4[(static_cast<int64_t>(sp + u * L) * RS + tap) * plane4 + c4];
          acc[u].x += v.x; acc[u].y += v.y; acc[u].z += v.z; acc[u].w += v.w;
        }
      }
      for (int u = 0; sp < splits; sp += L, ++u) {
        const float4 v = p4[(static_cast<int64_t>(sp) * RS + tap) * plane4 + c4];
        acc[u].x += v.x; acc[u].y += v.y; acc[u].z += v.z; acc[u].w += v.w;
      }
    }
    float4 t;
    t.x = (acc[0].x + acc[1].x) + (acc[2].x + acc[3].x);
    t.y = (acc[0].y + acc[1].y) + (acc[2].y + acc[3].y);
    t.z = (acc[0].z + acc[1].z) + (acc[2].z + acc[3].z);
    t.w = (acc[0].w + acc[1].w) + (acc[2].w + acc[3].w);
    if (L > 1) {
      wr_sh[threadIdx.y * 32 + col] = t;
      __syncthreads();
      if (lane == 0) {
        for (int l = 1; l < L; ++l) {
          const float4 v = wr_sh[(threadIdx.y + l) * 32 + col];
          t.x += v.x; t.y += v.y; t.z += v.z; t.w += v.w;
        }
      }
    }
    if (lane == 0 && live) {
      float* o = dw + c4 * 4 * RS + tap;
      o[0] = t.x; o[RS] = t.y; o[2 * RS] = t.z; o[3 * RS] = t.w;
    }
    if (L > 1) __syncthreads();
  }
}

// launch shape of the reduction: split lanes by how many partials there are and how few columns the layer has
static void launch_wgrad_reduce(const float* partial, int splits, int RS, int C_out, int C_in, float* dw, cudaStream_t s) {
  const int64_t total4 = static_cast<int64_t>(C_out) * C_in / 4 * RS;
  const int64_t cols32 = (total4 + 31) / 32;
  int L = 1;
  // few partials: one thread walks them all (four float4 loads in flight); many partials over few columns: up to 32 lanes
  if (splits > 16) L = cols32 >= 2 * dt_num_sms() ? (splits <= 64 ? 4 : 8) : (cols32 >= 96 ? 16 : 32);
  while (L > 1 && L * 4 > splits) L /= 2;                 // every lane gets at least four partials
  const int Y = L > 8 ? L : 8, G = Y / L;
  int64_t blocks = (cols32 + G - 1) / G;
  const int64_t cap = static_cast<int64_t>(dt_num_sms()) * (Y >= 16 ? 2 : 8);
  if (blocks > cap) blocks = cap;
  wgrad_reduce_kernel<<<static_cast<int>(blocks), dim3(32, Y), sizeof(float4) * 32 * Y, s>>>(partial, splits, RS, C_out, C_in, L, dw);
}

struct WgPlan {
  WgParams p;
  int splits;
  bool ok;
};

WgPlan make_plan(int N, int Ho, int Wo, int C_in, int C_out, int ksize, int stride) {
  WgPlan pl;
  memset(&pl, 0, sizeof(pl));
  WgParams& p = pl.p;
  const bool rows_ok = (Ho % TH == 0) || (Ho == 8 && N % 2 == 0);
  const bool kind_ok = (ksize == 3 && (stride == 1 || stride == 2)) || (ksize == 1 && stride == 2);
  pl.ok = N > 0 && Wo > 0 && Wo % TW == 0 && rows_ok && kind_ok && C_in > 0 && C_out > 0 && C_in % 4 == 0;
  if (!pl.ok) return pl;
  p.N = N; p.H = Ho; p.W = Wo; p.C_in = C_in; p.C_out = C_out;
  p.RS = ksize * ksize; p.stride = stride;
  p.th_img = Ho >= TH ? TH : Ho;
  p.imgs = TH / p.th_img;
  p.tiles_w = Wo / TW;
  p.tiles_h = Ho / p.th_img;
  p.total_tiles = p.tiles_w * p.tiles_h * (N / p.imgs);
  p.ci_slabs = (C_in + 63) / 64;
  if (ksize == 3 && stride == 1) {
    p.ngroups = 2;
    for (int tap = 0; tap < 9; ++tap) {
      WgGroup& g = p.grp[tap < 5 ? 0 : 1];
      g.tap[g.ntaps] = tap;
      g.off[g.ntaps] = (tap / 3) * PITCH + tap % 3;
      ++g.ntaps;
    }
  } else if (ksize == 3) {      // stride 2: one group per parity plane of the input
    p.ngroups = 4;
    for (int tap = 0; tap < 9; ++tap) {
      const int r = tap / 3, s = tap % 3;
      const int pr = (r + 1) & 1, pc = (s + 1) & 1;
      WgGroup& g = p.grp[pr * 2 + pc];
      g.pr = pr; g.pc = pc;
      g.tap[g.ntaps] = tap;
      g.off[g.ntaps] = ((r - 1 - pr) / 2 + 1) * PITCH + (s - 1 - pc) / 2 + 1;
      ++g.ntaps;
    }
  } else {                      // 1x1 / stride 2 / pad 0: plane (0, 0), window = the tile itself
    p.ngroups = 1;
    p.grp[0].ntaps = 1;
    p.grp[0].tap[0] = 0;
    p.grp[0].off[0] = PITCH + 1;
  }
  const int co_blocks = (C_out + 127) / 128;
  p.jobs = co_blocks * p.ci_slabs * p.ngroups;
  p.patch_bytes = p.imgs * (p.th_img + 2) * PITCH * 128;
  int splits = (2 * dt_num_sms() + p.jobs - 1) / p.jobs;      // about two waves of CTAs
  if (splits > p.total_tiles) splits = p.total_tiles;
  if (splits < 1) splits = 1;
  p.tiles_per_cta = (p.total_tiles + splits - 1) / splits;
  pl.splits = (p.total_tiles + p.tiles_per_cta - 1) / p.tiles_per_cta;
  return pl;
}

}  // namespace

int dt_encode_bf16_map(CUtensorMap* tm, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                       const uint32_t* box, const uint32_t* elem_strides);

// ------------------------------------------------------------------------------------------------------------------
// 3x3 / stride-1 layers with whole 16-row tiles: all nine taps of a (co block, ci slab) from ONE (C_out <= 32) or TWO
// (64-wide co blocks) tcgen05.mma per 16 pixels.
//
// With <= 32 channels a pixel is a 32- or 64-byte row, so the MN-major canonical layout (SWIZZLE_32B / _64B) has one
// "MN block" per pixel row and the block stride (leading byte offset) is free to ALIAS the same patch at a shifted pixel:
//   A = gy halo patch (18 rows x 8 px x CA ch), block i (i = 0..2) starts i image rows further: rows of A are
//       (i, co) = gy[q + (i-1) rows][co];  blocks 3.. of the M = 128 tile read further rows (garbage, never stored);
//   B = x halo patch (18 x 10 px x CB ch), block s' starts s' pixels further: columns (s', ci) = x[q + (s'-1) px][ci];
//   D[(i, co)][(s', ci)] += sum over the tile's pixels q  =  this tile's share of dW[co][ci][r = 2 - i][s = s']
// For 64-channel co blocks (SWIZZLE_128B rows) three blocks are 192 rows: a second MMA starts at block 2 (its upper
// half is garbage) and accumulates into its own TMEM columns.
// (every (gy pixel, x pixel) pair of a tap is met in exactly one tile; pairs that reach outside the image meet the zero
// fill of the TMA boxes).  The generic kernel above spends 9 MMAs of M 128 x N 64 on the same 16 pixels - for a 16-channel
// layer 32x more tensor work than the real 16 x 16 block; it ran the three 256^2-resolution decoder layers at 4..40
// TFLOP/s (890 us each).  This kernel is bound by reading x and gy once.
// The swizzle of both operands is a function of the absolute shared-memory address (profiles/
// r01_halo_descriptor_experiment.txt), which is what makes the aliased block strides legal.
// ------------------------------------------------------------------------------------------------------------------
namespace {

constexpr int NW_STAGES = 4;

struct NwParams {
  int N, H, W, C_in, C_out;
  int tiles_w, tiles_h, total_tiles, tiles_per_cta, n_slabs, co_blocks;
  float* partial;                   // [splits][9][C_out][C_in]
};

__device__ __forceinline__ uint64_t umma_desc_mn_any(uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(layout & 7u) << 61;
  return d;
}

template <int CA, int CB>
struct NwCfg {
  static constexpr int RA = CA * 2, RB = CB * 2;                      // bytes per pixel row
  static constexpr int BLK = 128 / CA;                                // MN blocks (image-row shifts) per M = 128 tile
  static constexpr int NM = (3 + BLK - 1) / BLK;                      // MMAs per 16 pixels: 1, or 2 for 64-wide co blocks
  static constexpr int A_PATCH = (TH + 2) * TW * RA;                  // 18 x 8 pixels
  static constexpr int B_PATCH = (TH + 2) * PITCH * RB;               // 18 x 10 pixels
  static constexpr int A_REGION = (A_PATCH + 1023) / 1024 * 1024;
  static constexpr int B_REGION = (B_PATCH + 1023) / 1024 * 1024;     // also absorbs the garbage blocks' over-read
  static constexpr int STAGE = A_REGION + B_REGION;
  static constexpr int NCOLS = 3 * CB;
  static constexpr int TMEM = NM * NCOLS <= 64 ? 64 : (NM * NCOLS <= 128 ? 128 : (NM * NCOLS <= 256 ? 256 : 512));
  static constexpr int SMEM = NW_STAGES * STAGE + 1024 + 256;
  static constexpr uint32_t LAYOUT_A = CA == 16 ? 6u : (CA == 32 ? 4u : 2u);   // SWIZZLE_32B / 64B / 128B
  static constexpr uint32_t LAYOUT_B = CB == 16 ? 6u : (CB == 32 ? 4u : 2u);
  static_assert((TH + NM * BLK) * TW * RA <= A_REGION + B_REGION, "garbage rows must stay inside the stage");
};

template <int CA, int CB>
__global__ void __launch_bounds__(kThreads, 1)
conv_wgrad_narrow_kernel(const __grid_constant__ CUtensorMap tm_g, const __grid_constant__ CUtensorMap tm_x,
                         const NwParams p) {
  using Cfg = NwCfg<CA, CB>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + NW_STAGES * Cfg::STAGE);
  uint64_t* full = bars;
  uint64_t* empty = full + NW_STAGES;
  uint64_t* done = empty + NW_STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int slab = blockIdx.x % p.n_slabs;
  const int cob = (blockIdx.x / p.n_slabs) % p.co_blocks, split = blockIdx.x / (p.n_slabs * p.co_blocks);
  const int tile_begin = split * p.tiles_per_cta;
  const int tile_end = min(p.total_tiles, tile_begin + p.tiles_per_cta);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_g);
    tma_prefetch_desc(&tm_x);
    for (int i = 0; i < NW_STAGES; ++i) { mbar_init(&full[i], 1u); mbar_init(&empty[i], 1u); }
    mbar_init(done, 1u);
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, Cfg::TMEM);
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
        int m = tile;
        const int w0 = (m % p.tiles_w) * TW; m /= p.tiles_w;
        const int h0 = (m % p.tiles_h) * TH;
        const int n0 = m / p.tiles_h;
        uint8_t* stage = smem + st * Cfg::STAGE;
        mbar_wait(&empty[st], ph ^ 1u);
        mbar_arrive_expect_tx(&full[st], Cfg::A_PATCH + Cfg::B_PATCH);
        tma_load_4d(stage, &tm_g, &full[st], cob * CA, w0, h0 - 1, n0);
        tma_load_4d(stage + Cfg::A_REGION, &tm_x, &full[st], slab * CB, w0 - 1, h0 - 1, n0);
        if (++st == NW_STAGES) { st = 0; ph ^= 1u; }
      }
    }
  } else if (warp == 1) {
    {   // whole warp: warp-uniform bookkeeping (uniform registers); an elected lane issues (see conv_halo.cu)
      constexpr uint32_t idesc = umma_idesc_bf16(128, Cfg::NCOLS) | (1u << 15) | (1u << 16);
      // A: MN blocks one image row (8 px) apart = the K groups' own stride; B: MN blocks one pixel apart
      const uint64_t a_hi = umma_desc_mn_any(TW * Cfg::RA, TW * Cfg::RA, Cfg::LAYOUT_A);
      const uint64_t b_hi = umma_desc_mn_any(Cfg::RB, PITCH * Cfg::RB, Cfg::LAYOUT_B);
      int st = 0;
      uint32_t ph = 0, accum = 0;
      for (int tile = tile_begin; tile < tile_end; ++tile) {
        mbar_wait(&full[st], ph);
        tc_fence_after();
        const uint32_t a_addr = smem_u32(smem + st * Cfg::STAGE);
        const uint32_t b_addr = a_addr + Cfg::A_REGION;
        if (elect_one()) {
#pragma unroll
          for (int k8 = 0; k8 < 8; ++k8) {      // 16 pixels = tile rows 2*k8, 2*k8 + 1
            const uint64_t a_d = a_hi + (((a_addr + 2 * k8 * TW * Cfg::RA) & 0x3FFFFu) >> 4);
            const uint64_t b_d = b_hi + (((b_addr + (2 * k8 + 1) * PITCH * Cfg::RB) & 0x3FFFFu) >> 4);
#pragma unroll
            for (int a = 0; a < Cfg::NM; ++a)   // second MMA: blocks BLK.. (image rows BLK further)
              umma_bf16_ss(tmem_base + a * Cfg::NCOLS, a_d + ((a * Cfg::BLK * TW * Cfg::RA) >> 4), b_d, idesc,
                           (accum | k8) != 0 ? 1u : 0u);
          }
          umma_commit(&empty[st]);
          if (tile + 1 == tile_end) umma_commit(done);
        }
        __syncwarp();
        accum = 1;
        if (++st == NW_STAGES) { st = 0; ph ^= 1u; }
      }
      if (tile_begin >= tile_end && elect_one()) umma_commit(done);
    }
  } else {
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;                 // accumulator row = (block, channel of the co block)
    mbar_wait(done, 0);
    tc_fence_after();
#pragma unroll
    for (int a = 0; a < Cfg::NM; ++a) {
      const int i = a * Cfg::BLK + row / CA;             // image-row shift of this accumulator row: tap row r = 2 - i
      const int co = cob * CA + row % CA;
      if (a * Cfg::BLK + (quarter * 32) / CA >= 3) continue;     // warp-uniform: only garbage rows in this quarter
#pragma unroll
      for (int sp = 0; sp < 3; ++sp) {
        const int tap = (2 - i) * 3 + sp;
        float* dst = p.partial + ((static_cast<int64_t>(split) * 9 + tap) * p.C_out + co) * p.C_in + slab * CB;
#pragma unroll
        for (int c0 = 0; c0 < CB; c0 += 16) {
          uint32_t v[16];
          tmem_ld_x16(tmem_base + a * Cfg::NCOLS + sp * CB + c0 + (static_cast<uint32_t>(quarter * 32) << 16), v);
          tmem_ld_wait();
          if (i < 3 && co < p.C_out) {
#pragma unroll
            for (int j = 0; j < 16; j += 4) {
              if (slab * CB + c0 + j < p.C_in)
                *reinterpret_cast<float4*>(dst + c0 + j) = make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]),
                                                                       __uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]));
            }
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::TMEM);
  }
}

struct NwPlan {
  NwParams p;
  int ca, cb, splits;
  bool ok;
};

NwPlan make_narrow_plan(int N, int Ho, int Wo, int C_in, int C_out, int ksize, int stride) {
  NwPlan pl;
  memset(&pl, 0, sizeof(pl));
  pl.ok = ksize == 3 && stride == 1 && N > 0 && Ho % TH == 0 && Wo % TW == 0 && C_out > 0 && C_in > 0 && C_in % 4 == 0;
  if (!pl.ok) return pl;
  NwParams& p = pl.p;
  pl.ca = C_out <= 16 ? 16 : (C_out <= 32 ? 32 : 64);
  pl.cb = C_in <= 16 ? 16 : (C_in <= 32 || C_in % 64 != 0 ? 32 : 64);
  p.N = N; p.H = Ho; p.W = Wo; p.C_in = C_in; p.C_out = C_out;
  p.tiles_w = Wo / TW; p.tiles_h = Ho / TH;
  p.total_tiles = p.tiles_w * p.tiles_h * N;
  p.n_slabs = (C_in + pl.cb - 1) / pl.cb;
  p.co_blocks = (C_out + pl.ca - 1) / pl.ca;
  int splits = dt_num_sms() / (p.n_slabs * p.co_blocks);   // one wave of one CTA per SM
  if (splits < 1) splits = 1;
  if (splits > p.total_tiles) splits = p.total_tiles;
  p.tiles_per_cta = (p.total_tiles + splits - 1) / splits;
  pl.splits = (p.total_tiles + p.tiles_per_cta - 1) / p.tiles_per_cta;
  return pl;
}

template <int CA, int CB>
int launch_narrow(const CUtensorMap& tm_g, const CUtensorMap& tm_x, const NwParams& p, int splits, cudaStream_t s) {
  using Cfg = NwCfg<CA, CB>;
  static std::once_flag once;
  static cudaError_t attr_err = cudaSuccess;
  std::call_once(once, [] {
    attr_err = cudaFuncSetAttribute(conv_wgrad_narrow_kernel<CA, CB>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM);
  });
  DT_CUDA(attr_err);
  conv_wgrad_narrow_kernel<CA, CB><<<p.n_slabs * p.co_blocks * splits, kThreads, Cfg::SMEM, s>>>(tm_g, tm_x, p);
  DT_LAUNCH_CHECK();
  return DT_OK;
}

}  // namespace

extern "C" int64_t dt_conv2d_wgrad_tc_workspace(int N, int Ho, int Wo, int C_in, int C_out, int ksize, int stride) {
  const NwPlan nw = make_narrow_plan(N, Ho, Wo, C_in, C_out, ksize, stride);
  if (nw.ok) return static_cast<int64_t>(nw.splits) * 9 * C_out * C_in * static_cast<int64_t>(sizeof(float));
  const WgPlan pl = make_plan(N, Ho, Wo, C_in, C_out, ksize, stride);
  if (!pl.ok) return DT_ERR_UNSUPPORTED;
  return static_cast<int64_t>(pl.splits) * pl.p.RS * C_out * C_in * static_cast<int64_t>(sizeof(float));
}

extern "C" int dt_conv2d_wgrad_tc(const void* x, const void* gy, int N, int Ho, int Wo, int C_in, int x_cstride, int C_out,
                                  int gy_cstride, int ksize, int stride, float* dw_oihw, float* workspace,
                                  int64_t workspace_bytes, dt_stream_t stream) {
  DT_ARCH_GUARD();
  NwPlan nw = make_narrow_plan(N, Ho, Wo, C_in, C_out, ksize, stride);
  if (nw.ok && x_cstride >= C_in && gy_cstride >= C_out && x_cstride % 8 == 0 && gy_cstride % 8 == 0) {
    NwParams& q = nw.p;
    const int64_t need = static_cast<int64_t>(nw.splits) * 9 * C_out * C_in * static_cast<int64_t>(sizeof(float));
    DT_REQUIRE(workspace != nullptr && workspace_bytes >= need, DT_ERR_BAD_SHAPE,
               "dt_conv2d_wgrad_tc: workspace of %lld bytes needed (dt_conv2d_wgrad_tc_workspace)", static_cast<long long>(need));
    DT_REQUIRE((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(gy) | reinterpret_cast<uintptr_t>(workspace)) % 16 == 0,
               DT_ERR_BAD_ALIGN, "dt_conv2d_wgrad_tc: tensors must be 16-byte aligned");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    q.partial = workspace;
    CUtensorMap tm_g, tm_x;
    {
      const uint64_t dims[4] = {static_cast<uint64_t>(gy_cstride), static_cast<uint64_t>(Wo), static_cast<uint64_t>(Ho),
                                static_cast<uint64_t>(N)};
      const uint64_t strides[3] = {static_cast<uint64_t>(gy_cstride) * 2, static_cast<uint64_t>(Wo) * gy_cstride * 2,
                                   static_cast<uint64_t>(Ho) * Wo * gy_cstride * 2};
      const uint32_t box[4] = {static_cast<uint32_t>(nw.ca), TW, TH + 2, 1};
      int rc = dt_encode_bf16_map(&tm_g, gy, 4, dims, strides, box, nullptr);
      if (rc != DT_OK) return rc;
    }
    {
      const uint64_t dims[4] = {static_cast<uint64_t>(x_cstride), static_cast<uint64_t>(Wo), static_cast<uint64_t>(Ho),
                                static_cast<uint64_t>(N)};
      const uint64_t strides[3] = {static_cast<uint64_t>(x_cstride) * 2, static_cast<uint64_t>(Wo) * x_cstride * 2,
                                   static_cast<uint64_t>(Ho) * Wo * x_cstride * 2};
      const uint32_t box[4] = {static_cast<uint32_t>(nw.cb), PITCH, TH + 2, 1};
      int rc = dt_encode_bf16_map(&tm_x, x, 4, dims, strides, box, nullptr);
      if (rc != DT_OK) return rc;
    }
    int rc;
    if (nw.ca == 16 && nw.cb == 16) rc = launch_narrow<16, 16>(tm_g, tm_x, q, nw.splits, s);
    else if (nw.ca == 16 && nw.cb == 32) rc = launch_narrow<16, 32>(tm_g, tm_x, q, nw.splits, s);
    else if (nw.ca == 16) rc = launch_narrow<16, 64>(tm_g, tm_x, q, nw.splits, s);
    else if (nw.ca == 32 && nw.cb == 16) rc = launch_narrow<32, 16>(tm_g, tm_x, q, nw.splits, s);
    else if (nw.ca == 32 && nw.cb == 32) rc = launch_narrow<32, 32>(tm_g, tm_x, q, nw.splits, s);
    else if (nw.ca == 32) rc = launch_narrow<32, 64>(tm_g, tm_x, q, nw.splits, s);
    else if (nw.cb == 16) rc = launch_narrow<64, 16>(tm_g, tm_x, q, nw.splits, s);
    else if (nw.cb == 32) rc = launch_narrow<64, 32>(tm_g, tm_x, q, nw.splits, s);
    else rc = launch_narrow<64, 64>(tm_g, tm_x, q, nw.splits, s);
    if (rc != DT_OK) return rc;
    launch_wgrad_reduce(workspace, nw.splits, 9, C_out, C_in, dw_oihw, s);
    DT_LAUNCH_CHECK();
    return DT_OK;
  }
  WgPlan pl = make_plan(N, Ho, Wo, C_in, C_out, ksize, stride);
  if (!pl.ok || x_cstride < C_in || gy_cstride < C_out || x_cstride % 8 != 0 || gy_cstride % 8 != 0) {
    dt_set_error("dt_conv2d_wgrad_tc: unsupported shape N=%d Ho=%d Wo=%d C_in=%d C_out=%d k=%d s=%d", N, Ho, Wo, C_in, C_out,
                 ksize, stride);
    return DT_ERR_UNSUPPORTED;
  }
  WgParams& p = pl.p;
  const int64_t need = static_cast<int64_t>(pl.splits) * p.RS * C_out * C_in * static_cast<int64_t>(sizeof(float));
  DT_REQUIRE(workspace != nullptr && workspace_bytes >= need, DT_ERR_BAD_SHAPE,
             "dt_conv2d_wgrad_tc: workspace of %lld bytes needed (dt_conv2d_wgrad_tc_workspace)", static_cast<long long>(need));
  DT_REQUIRE((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(gy) | reinterpret_cast<uintptr_t>(workspace)) % 16 == 0,
             DT_ERR_BAD_ALIGN, "dt_conv2d_wgrad_tc: tensors must be 16-byte aligned");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  p.partial = workspace;
  const int Hi = Ho * stride, Wi = Wo * stride;

  CUtensorMap tm_g, tm_x;
  {
    const uint64_t dims[4] = {static_cast<uint64_t>(gy_cstride), static_cast<uint64_t>(Wo), static_cast<uint64_t>(Ho),
                              static_cast<uint64_t>(N)};
    const uint64_t strides[3] = {static_cast<uint64_t>(gy_cstride) * 2, static_cast<uint64_t>(Wo) * gy_cstride * 2,
                                 static_cast<uint64_t>(Ho) * Wo * gy_cstride * 2};
    const uint32_t box[4] = {64, TW, static_cast<uint32_t>(p.th_img), static_cast<uint32_t>(p.imgs)};
    int rc = dt_encode_bf16_map(&tm_g, gy, 4, dims, strides, box, nullptr);
    if (rc != DT_OK) return rc;
  }
  {
    const uint32_t st = static_cast<uint32_t>(stride);
    const uint64_t dims[4] = {static_cast<uint64_t>(x_cstride), static_cast<uint64_t>(Wi), static_cast<uint64_t>(Hi),
                              static_cast<uint64_t>(N)};
    const uint64_t strides[3] = {static_cast<uint64_t>(x_cstride) * 2, static_cast<uint64_t>(Wi) * x_cstride * 2,
                                 static_cast<uint64_t>(Hi) * Wi * x_cstride * 2};
    const uint32_t box[4] = {64, PITCH * st, static_cast<uint32_t>(p.th_img + 2) * st, static_cast<uint32_t>(p.imgs)};
    const uint32_t estr[4] = {1, st, st, 1};
    int rc = dt_encode_bf16_map(&tm_x, x, 4, dims, strides, box, estr);
    if (rc != DT_OK) return rc;
  }
  constexpr int SMEM = STAGES * (A_STAGE + B_STAGE) + 1024 + 256;
  static std::once_flag once;
  static cudaError_t attr_err = cudaSuccess;
  std::call_once(once, [] {
    attr_err = cudaFuncSetAttribute(conv_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
  });
  DT_CUDA(attr_err);
  conv_wgrad_kernel<<<p.jobs * pl.splits, kThreads, SMEM, s>>>(tm_g, tm_x, p);
  DT_LAUNCH_CHECK();
  launch_wgrad_reduce(workspace, pl.splits, p.RS, C_out, C_in, dw_oihw, s);
  DT_LAUNCH_CHECK();
  return DT_OK;
}
