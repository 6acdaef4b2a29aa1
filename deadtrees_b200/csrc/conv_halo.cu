// 3x3 / stride-1 / pad-1 convolution with shared-memory resident input patches (tcgen05 implicit GEMM).
//
// One tile = 8 (w) x 16 (h) output pixels = the 128 rows of a UMMA M=128 tile.  For every slab of `CW` input
// channels ONE TMA box brings a (16+2) x (8+2) pixel halo patch into shared memory; the filter taps then read it
// in place: the A-operand descriptor of a tap starts `off` pixel rows into the patch and steps PITCH pixel rows
// between 8-pixel groups (stride-byte-offset).  This works because the UMMA shared-memory swizzle is a function
// of the absolute smem address (measured: scripts/halo_exp.py, profiles/r01_halo_descriptor_experiment.txt), so
// a shifted window of a TMA-swizzled patch is still a valid K-major operand.  Activations therefore cross
// L2 -> SM once instead of nine times.  Weights stream through a separate TMA ring, one 64-wide K chunk at a time.
//
// Two tile families share the kernel:
//   plain  : y = conv3x3(x)                      tile pixels = (h0+i, w0+j)
//   parity : y = conv3x3(cat[up2(x), skip])      tile pixels = (2(h0+i)+a, 2(w0+j)+b) for one parity class (a,b);
//            the up-sampled operand is the LOW-RES patch read at offset ((a+r-1)>>1, (b+s-1)>>1) - nine taps from
//            one patch - and the skip operand comes from the four stride-2 "plane" patches
//            skip[2i+pr, 2j+pc] (TMA traversal stride 2), each serving the taps with (a+r-1)&1 == pr, (b+s-1)&1 == pc.
//            Neither the up-sampled nor the concatenated tensor is ever materialised.
// The per-class tap lists / patch offsets are small tables in the kernel parameters.
//
//   warp 0: TMA producer (patches + weight chunks)   warp 1: TMEM alloc + MMA issuer   warps 2..5: epilogue
// Persistent CTAs, double-buffered TMEM accumulator (epilogue of tile i overlaps the main loop of tile i+1).
// K is accumulated patch-major - a different fp32 summation order than conv_tc.cu.
#include <cstring>
#include <mutex>

#include "common.cuh"

namespace {

constexpr int BM = 128;
constexpr int BK = 64;
constexpr int TW = 8, TH = 16;       // tile of output pixels (of one parity class in parity mode)
constexpr int kThreads = 192;
constexpr int MAX_A = 3;
constexpr int MAX_B = 8;
constexpr int PITCH = TW + 2;        // patch row pitch in pixels
constexpr int PATCH_PIX = (TH + 2) * PITCH;

struct PatchDesc {                   // one patch kind of one parity class
  uint8_t pr, pc, ntaps, pad;
  uint8_t tap[9];                    // filter tap r*3+s served by this patch (wide: one weight chunk per entry)
  uint8_t off[9];                    // pixel offset of that tap's window inside the patch, indexed like tap[]
  uint8_t off_by_tap[9];             // same offsets indexed by the natural tap number (narrow layers)
  uint8_t pad2[1];
};

struct HaloParams {
  int N, H, W, C_in, C_x, C_out;     // H, W: output (= virtual input) size
  int parity;                        // 0 plain, 1 parity classes
  int Hg, Wg;                        // pixel grid the tiles walk: (H, W) or (H/2, W/2)
  int n_xslab, n_sslab;              // slabs of the x operand / of the skip operand
  int a_slots, b_slots, a_slot_bytes;
  int relu, has_residual;
  int tiles_w, tiles_h, tiles_per_class, n_tiles, total_tiles;
  const __nv_bfloat16* residual;
  __nv_bfloat16* y;
  const float* scale;
  const float* shift;
  PatchDesc patch[4][5];             // [class][0] = x patch, [class][1..4] = skip planes
  int fold;                          // class-fused kernel: x operand with folded up-sampling weights (DT_CONV_UPS_FOLDED):
  int k_skip;                        //   K index of (class c, low-res pixel e, x channel ci) = (c*4+e)*C_x + ci, skip taps from k_skip
};

struct Geo {
  int n_tile, cls, n, h0, w0;
};
__device__ __forceinline__ Geo geo(const HaloParams& p, int tile) {
  Geo g;
  g.n_tile = tile % p.n_tiles;
  int m = tile / p.n_tiles;
  // parity layers: the four classes of one region are consecutive tiles (they run on neighbouring CTAs at the same
  // time and share the x patch and the four skip planes in L2)
  g.cls = p.parity ? (m & 3) : 0;
  if (p.parity) m >>= 2;
  g.w0 = (m % p.tiles_w) * TW; m /= p.tiles_w;
  g.h0 = (m % p.tiles_h) * TH;
  g.n = m / p.tiles_h;
  return g;
}

template <int BN, int CW>
__global__ void __launch_bounds__(kThreads, BN <= 128 ? 2 : 1)
conv_halo_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_s,
                 const __grid_constant__ CUtensorMap tm_b, const HaloParams p) {
  constexpr int B_BYTES = BN * BK * 2;
  constexpr int TMEM_COLS = 2 * BN < 32 ? 32 : 2 * BN;
  constexpr int ROW_BYTES = CW * 2;
  constexpr int PATCH_BYTES = PATCH_PIX * ROW_BYTES;
  constexpr int NARROW_CHUNKS = (9 * CW + BK - 1) / BK;    // weight chunks of a narrow (CW < 64) layer
  constexpr uint32_t idesc = umma_idesc_bf16(BM, BN);
  const int A_SLOTS = p.a_slots;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;                                   // a_slots x a_slot_bytes (1024-aligned each)
  uint8_t* smem_b = smem + A_SLOTS * p.a_slot_bytes;        // b_slots x B_BYTES
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_b + p.b_slots * B_BYTES);
  uint64_t* full_a = bars;                 // [MAX_A]
  uint64_t* empty_a = full_a + MAX_A;      // [MAX_A]
  uint64_t* full_b = empty_a + MAX_A;      // [MAX_B]
  uint64_t* empty_b = full_b + MAX_B;      // [MAX_B]
  uint64_t* tmem_full = empty_b + MAX_B;   // [2]
  uint64_t* tmem_empty = tmem_full + 2;    // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_a);
    tma_prefetch_desc(&tm_b);
    if (p.n_sslab > 0) tma_prefetch_desc(&tm_s);
    for (int i = 0; i < A_SLOTS; ++i) { mbar_init(&full_a[i], 1u); mbar_init(&empty_a[i], 1u); }
    for (int i = 0; i < p.b_slots; ++i) { mbar_init(&full_b[i], 1u); mbar_init(&empty_b[i], 1u); }
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
  const int patches_per_tile = p.n_xslab + 4 * p.n_sslab;   // patch instance pi: x slabs first, then skip planes

  if (warp == 0) {
    if (lane == 0) {
      int sa = 0, sb = 0;
      uint32_t pa = 0, pb = 0;
      // request patch instance `pi` of the tile with geometry `g`
      auto load_patch = [&](const Geo& g, int pi) {
        mbar_wait(&empty_a[sa], pa ^ 1u);
        mbar_arrive_expect_tx(&full_a[sa], PATCH_BYTES);
        uint8_t* dst = smem_a + sa * p.a_slot_bytes;
        if (pi < p.n_xslab) {
          tma_load_4d(dst, &tm_a, &full_a[sa], pi * CW, g.w0 - 1, g.h0 - 1, g.n);
        } else {
          const int s = pi - p.n_xslab;
          const PatchDesc& pd = p.patch[g.cls][1 + (s & 3)];
          tma_load_4d(dst, &tm_s, &full_a[sa], (s >> 2) * BK, 2 * (g.w0 - 1) + pd.pc, 2 * (g.h0 - 1) + pd.pr, g.n);
        }
        if (++sa == A_SLOTS) { sa = 0; pa ^= 1u; }
      };
      bool primed = false;
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
        const Geo g = geo(p, tile);
        for (int pi = 0; pi < patches_per_tile; ++pi) {
          if (!primed) { load_patch(g, pi); primed = true; }   // very first patch of this CTA
          const int kind = pi < p.n_xslab ? 0 : 1 + ((pi - p.n_xslab) & 3);
          const PatchDesc& pd = p.patch[g.cls][kind];
          // weight K coordinate of this patch's channels: k = tap * C_in + choff (+ channel)
          const int choff = pi < p.n_xslab ? pi * CW : p.C_x + ((pi - p.n_xslab) >> 2) * BK;
          const int nchunks = CW == BK ? pd.ntaps : NARROW_CHUNKS;
          // With three patch slots the next patch is requested before this patch's weight chunks; with two its
          // slot is still being read, so it is requested after the first few weight chunks instead.
          const int prefetch_at = A_SLOTS >= 3 ? 0 : (nchunks > 5 ? 4 : nchunks - 1);
          for (int q = 0; q < nchunks; ++q) {
            if (q == prefetch_at) {
              int npi = pi + 1, ntile = tile;
              if (npi == patches_per_tile) { npi = 0; ntile = tile + gridDim.x; }
              if (ntile < p.total_tiles) load_patch(geo(p, ntile), npi);
            }
            const int kcoord = CW == BK ? pd.tap[q] * p.C_in + choff : q * BK;
            mbar_wait(&empty_b[sb], pb ^ 1u);
            mbar_arrive_expect_tx(&full_b[sb], B_BYTES);
            tma_load_2d(smem_b + sb * B_BYTES, &tm_b, &full_b[sb], kcoord, g.n_tile * BN);
            if (++sb == p.b_slots) { sb = 0; pb ^= 1u; }
          }
        }
      }
    }
  } else if (warp == 1) {
    {
      // The whole warp runs the warp-uniform bookkeeping (barrier waits, descriptor arithmetic: uniform registers); one
      // elected lane issues the MMAs and commits.  Inside an `if (lane == 0)` region every operand took an R2UR / ELECT
      // round trip - ~100 clk of pure issue per MMA, more than the tensor pipe needs for N <= 64.
      const uint64_t a_hi = umma_desc(0u, PITCH * ROW_BYTES, CW == 64 ? 2u : (CW == 32 ? 4u : 6u));
      const uint64_t b_hi = umma_desc(0u, 1024u, 2u);
      int sa = 0, sb = 0, acc = 0;
      uint32_t pa = 0, pb = 0, pacc = 0;
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
        const int cls = p.parity ? ((tile / p.n_tiles) & 3) : 0;
        mbar_wait(&tmem_empty[acc], pacc ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        uint32_t accum = 0;
        for (int pi = 0; pi < patches_per_tile; ++pi) {
          const int kind = pi < p.n_xslab ? 0 : 1 + ((pi - p.n_xslab) & 3);
          const PatchDesc& pd = p.patch[cls][kind];
          mbar_wait(&full_a[sa], pa);
          const uint64_t a_d = a_hi + (smem_u32(smem_a + sa * p.a_slot_bytes) >> 4);
          const int nchunks = CW == BK ? pd.ntaps : NARROW_CHUNKS;
          for (int q = 0; q < nchunks; ++q) {
            mbar_wait(&full_b[sb], pb);
            tc_fence_after();
            const uint64_t b_d = b_hi + (smem_u32(smem_b + sb * B_BYTES) >> 4);
            if (CW == BK) {
              const uint64_t a_t = a_d + ((static_cast<uint32_t>(pd.off[q]) * ROW_BYTES) >> 4);
              if (elect_one()) {
#pragma unroll
                for (int k = 0; k < BK / 16; ++k) umma_bf16_ss(d_tmem, a_t + 2 * k, b_d + 2 * k, idesc, (accum | k) ? 1u : 0u);
                umma_commit(&empty_b[sb]);
              }
              accum = 1;
            } else {
              // narrow layer: chunk q holds K elements [64q, 64q+64) = taps (64q+16k)/CW in natural order
              const int ksteps = min(BK / 16, (9 * CW - q * BK) / 16);
              if (elect_one()) {
#pragma unroll
                for (int k = 0; k < BK / 16; ++k) {
                  if (k < ksteps) {
                    const int kk = q * BK + k * 16;
                    const uint32_t off = static_cast<uint32_t>(pd.off_by_tap[kk / CW]) * ROW_BYTES + (kk % CW) * 2;
                    umma_bf16_ss(d_tmem, a_d + (off >> 4), b_d + 2 * k, idesc, (accum | k) ? 1u : 0u);
                  }
                }
                umma_commit(&empty_b[sb]);
              }
              accum = 1;
            }
            __syncwarp();
            if (++sb == p.b_slots) { sb = 0; pb ^= 1u; }
          }
          if (elect_one()) umma_commit(&empty_a[sa]);
          __syncwarp();
          if (++sa == A_SLOTS) { sa = 0; pa ^= 1u; }
        }
        if (elect_one()) umma_commit(&tmem_full[acc]);
        __syncwarp();
        if ((acc ^= 1) == 0) pacc ^= 1u;
      }
    }
  } else {
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    constexpr int SC = BN < 64 ? BN : 64;
    int acc = 0;
    uint32_t pacc = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
      const Geo g = geo(p, tile);
      int oy = g.h0 + (row >> 3), ox = g.w0 + (row & 7);
      if (p.parity) { oy = 2 * oy + (g.cls >> 1); ox = 2 * ox + (g.cls & 1); }
      const int64_t out_off = ((static_cast<int64_t>(g.n) * p.H + oy) * p.W + ox) * p.C_out + g.n_tile * BN;
      uint4 res[SC / 8];
      if (p.has_residual) {
#pragma unroll
        for (int j = 0; j < SC / 16; ++j) ldg_v8(p.residual + out_off + 16 * j, res[2 * j], res[2 * j + 1]);
      }
      mbar_wait(&tmem_full[acc], pacc);
      tc_fence_after();
      const uint32_t t_row = tmem_base + acc * BN + (static_cast<uint32_t>(quarter * 32) << 16);
#pragma unroll 1
      for (int s0 = 0; s0 < BN; s0 += SC) {
        uint4 res_next[SC / 8];
        const bool more = s0 + SC < BN;
        if (p.has_residual && more) {
#pragma unroll
          for (int j = 0; j < SC / 16; ++j) ldg_v8(p.residual + out_off + s0 + SC + 16 * j, res_next[2 * j], res_next[2 * j + 1]);
        }
#pragma unroll
        for (int c0 = 0; c0 < SC; c0 += 16) {
          uint32_t v[16];
          tmem_ld_x16(t_row + s0 + c0, v);
          tmem_ld_wait();
          float f[16];
          const int co = g.n_tile * BN + s0 + c0;
#pragma unroll
          for (int j = 0; j < 16; j += 4) {
            const float4 sc = __ldg(reinterpret_cast<const float4*>(p.scale + co + j));
            const float4 sh = __ldg(reinterpret_cast<const float4*>(p.shift + co + j));
            f[j] = fmaf(__uint_as_float(v[j]), sc.x, sh.x);
            f[j + 1] = fmaf(__uint_as_float(v[j + 1]), sc.y, sh.y);
            f[j + 2] = fmaf(__uint_as_float(v[j + 2]), sc.z, sh.z);
            f[j + 3] = fmaf(__uint_as_float(v[j + 3]), sc.w, sh.w);
          }
          if (p.has_residual) {
            const uint32_t rr[8] = {res[c0 / 8].x, res[c0 / 8].y, res[c0 / 8].z, res[c0 / 8].w,
                                    res[c0 / 8 + 1].x, res[c0 / 8 + 1].y, res[c0 / 8 + 1].z, res[c0 / 8 + 1].w};
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float2 t = unpack_bf16x2(rr[j]);
              f[2 * j] += t.x;
              f[2 * j + 1] += t.y;
            }
          }
          if (p.relu) {
#pragma unroll
            for (int j = 0; j < 16; ++j) f[j] = fmaxf(f[j], 0.f);
          }
          store_bf16x16(p.y + out_off + s0 + c0, f);
        }
        if (more) {
#pragma unroll
          for (int j = 0; j < SC / 8; ++j) res[j] = res_next[j];
        }
      }
      tc_fence_before();
      mbar_arrive(&tmem_empty[acc]);
      if ((acc ^= 1) == 0) pacc ^= 1u;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

template <int BN, int CW>
int launch_halo(const CUtensorMap& tm_a, const CUtensorMap& tm_s, const CUtensorMap& tm_b, HaloParams& p,
                cudaStream_t s) {
  constexpr int B_BYTES = BN * BK * 2;
  constexpr int CTAS = BN <= 128 ? 2 : 1;
  // shared memory per CTA: ~111 KB when two CTAs share an SM, ~200 KB otherwise
  const int budget = (CTAS == 2 ? 111 : 200) * 1024 - 1024 - 512;
  p.a_slots = 3;
  int b_slots = (budget - 3 * p.a_slot_bytes) / B_BYTES;
  if (b_slots < 4) { p.a_slots = 2; b_slots = (budget - 2 * p.a_slot_bytes) / B_BYTES; }
  if (b_slots > MAX_B) b_slots = MAX_B;
  if (b_slots < 2) return DT_ERR_UNSUPPORTED;
  p.b_slots = b_slots;
  const int smem = p.a_slots * p.a_slot_bytes + b_slots * B_BYTES + 1024 + 512;
  static std::once_flag once;
  static cudaError_t attr_err = cudaSuccess;
  std::call_once(once, [] {
    attr_err = cudaFuncSetAttribute(conv_halo_kernel<BN, CW>, cudaFuncAttributeMaxDynamicSharedMemorySize, 225 * 1024);
  });
  DT_CUDA(attr_err);
  const int slots = dt_num_sms() * CTAS;
  const int grid = p.total_tiles < slots ? p.total_tiles : slots;
  conv_halo_kernel<BN, CW><<<grid, kThreads, smem, s>>>(tm_a, tm_s, tm_b, p);
  DT_LAUNCH_CHECK();
  return DT_OK;
}

// ------------------------------------------------------------------------------------------------------------------
// Class-fused variant for the up-sample + concat layers with C_out <= 64: ONE tile = the four parity classes of a
// low-res region (4 x 128 output pixels, four TMEM accumulators).  The five patches of the region (low-res x patch and
// the four stride-2 skip planes) and - for the x patch - the nine weight chunks are loaded once and shared by the four
// classes; in the class-per-tile kernel above every class re-loads all five patches (ncu, decoder.blocks.3.conv1:
// 1.4 GB of DRAM reads for 0.35 GB of operands, 187 KB of TMA traffic per 128 output pixels).  Per class the K order
// (x slabs, then planes, taps in table order) is unchanged, so the results are bit-identical to the kernel above.
// ------------------------------------------------------------------------------------------------------------------
template <int BN>
__global__ void __launch_bounds__(kThreads, 2)
conv_halo_quad_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_s,
                      const __grid_constant__ CUtensorMap tm_b, const HaloParams p) {
  constexpr int CW = BK;
  constexpr int B_BYTES = BN * BK * 2;
  constexpr int NBUF = 8 * BN <= 256 ? 2 : 1;              // two CTAs share the 512 TMEM columns of an SM
  constexpr int TMEM_COLS = NBUF * 4 * BN;
  constexpr int ROW_BYTES = CW * 2;
  constexpr int PATCH_BYTES = PATCH_PIX * ROW_BYTES;
  constexpr uint32_t idesc = umma_idesc_bf16(BM, BN);
  const int A_SLOTS = p.a_slots;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + A_SLOTS * p.a_slot_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_b + p.b_slots * B_BYTES);
  uint64_t* full_a = bars;
  uint64_t* empty_a = full_a + MAX_A;
  uint64_t* full_b = empty_a + MAX_A;
  uint64_t* empty_b = full_b + MAX_B;
  uint64_t* tmem_full = empty_b + MAX_B;   // [2]
  uint64_t* tmem_empty = tmem_full + 2;    // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_a);
    tma_prefetch_desc(&tm_b);
    if (p.n_sslab > 0) tma_prefetch_desc(&tm_s);
    for (int i = 0; i < A_SLOTS; ++i) { mbar_init(&full_a[i], 1u); mbar_init(&empty_a[i], 1u); }
    for (int i = 0; i < p.b_slots; ++i) { mbar_init(&full_b[i], 1u); mbar_init(&empty_b[i], 1u); }
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
  const int patches_per_tile = p.n_xslab + 4 * p.n_sslab;
  // region geometry: total_tiles = tiles_w * tiles_h * N (C_out == BN: one channel tile)
  auto region = [&](int tile, int& n, int& h0, int& w0) {
    w0 = (tile % p.tiles_w) * TW;
    const int m = tile / p.tiles_w;
    h0 = (m % p.tiles_h) * TH;
    n = m / p.tiles_h;
  };

  if (warp == 0) {
    if (lane == 0) {
      int sa = 0, sb = 0;
      uint32_t pa = 0, pb = 0;
      auto load_patch = [&](int tile, int pi) {
        int n, h0, w0;
        region(tile, n, h0, w0);
        mbar_wait(&empty_a[sa], pa ^ 1u);
        mbar_arrive_expect_tx(&full_a[sa], PATCH_BYTES);
        uint8_t* dst = smem_a + sa * p.a_slot_bytes;
        if (pi < p.n_xslab) {
          tma_load_4d(dst, &tm_a, &full_a[sa], pi * CW, w0 - 1, h0 - 1, n);
        } else {
          const int sidx = pi - p.n_xslab, plane = sidx & 3;
          tma_load_4d(dst, &tm_s, &full_a[sa], (sidx >> 2) * BK, 2 * (w0 - 1) + (plane & 1), 2 * (h0 - 1) + (plane >> 1), n);
        }
        if (++sa == A_SLOTS) { sa = 0; pa ^= 1u; }
      };
      auto load_chunk = [&](int kcoord) {
        mbar_wait(&empty_b[sb], pb ^ 1u);
        mbar_arrive_expect_tx(&full_b[sb], B_BYTES);
        tma_load_2d(smem_b + sb * B_BYTES, &tm_b, &full_b[sb], kcoord, 0);
        if (++sb == p.b_slots) { sb = 0; pb ^= 1u; }
      };
      bool primed = false;
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
        for (int pi = 0; pi < patches_per_tile; ++pi) {
          if (!primed) { load_patch(tile, pi); primed = true; }
          const int kind = pi < p.n_xslab ? 0 : 1 + ((pi - p.n_xslab) & 3);
          const int choff = pi < p.n_xslab ? pi * CW : p.C_x + ((pi - p.n_xslab) >> 2) * BK;
          const int prefetch_at = A_SLOTS >= 3 ? 0 : 4;
          int q = 0;
          auto maybe_prefetch = [&]() {
            if (q == prefetch_at) {
              int npi = pi + 1, ntile = tile;
              if (npi == patches_per_tile) { npi = 0; ntile = tile + gridDim.x; }
              if (ntile < p.total_tiles) load_patch(ntile, npi);
            }
            ++q;
          };
          if (kind == 0 && p.fold) {
            // folded up-sampling: four effective taps per class, each class its own summed weights
            for (int e = 0; e < 4; ++e)
              for (int cls = 0; cls < 4; ++cls) { maybe_prefetch(); load_chunk((cls * 4 + e) * p.C_x + pi * CW); }
          } else if (kind == 0) {
            for (int tap = 0; tap < 9; ++tap) { maybe_prefetch(); load_chunk(tap * p.C_in + choff); }
          } else {
            const int sl = ((pi - p.n_xslab) >> 2) * BK;
            for (int cls = 0; cls < 4; ++cls) {
              const PatchDesc& pd = p.patch[cls][kind];
              for (int j = 0; j < pd.ntaps; ++j) {
                maybe_prefetch();
                load_chunk(p.fold ? p.k_skip + pd.tap[j] * (p.C_in - p.C_x) + sl : pd.tap[j] * p.C_in + choff);
              }
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    {   // whole warp: uniform bookkeeping; an elected lane issues (see conv_halo_kernel)
      const uint64_t a_hi = umma_desc(0u, PITCH * ROW_BYTES, 2u);
      const uint64_t b_hi = umma_desc(0u, 1024u, 2u);
      int sa = 0, sb = 0, acc = 0;
      uint32_t pa = 0, pb = 0, pacc = 0;
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
        mbar_wait(&tmem_empty[acc], pacc ^ 1u);
        tc_fence_after();
        uint32_t started = 0;
        for (int pi = 0; pi < patches_per_tile; ++pi) {
          const int kind = pi < p.n_xslab ? 0 : 1 + ((pi - p.n_xslab) & 3);
          mbar_wait(&full_a[sa], pa);
          tc_fence_after();
          const uint64_t a_d = a_hi + (smem_u32(smem_a + sa * p.a_slot_bytes) >> 4);
          if (kind == 0 && p.fold) {
            for (int e = 0; e < 4; ++e) {
#pragma unroll
              for (int cls = 0; cls < 4; ++cls) {
                mbar_wait(&full_b[sb], pb);
                tc_fence_after();
                const uint64_t b_d = b_hi + (smem_u32(smem_b + sb * B_BYTES) >> 4);
                // low-res pixel (a-1+ey, b-1+ex) of class (a,b): patch offset (a+ey, b+ex)
                const uint32_t off = static_cast<uint32_t>(((cls >> 1) + (e >> 1)) * PITCH + (cls & 1) + (e & 1));
                const uint64_t a_t = a_d + ((off * ROW_BYTES) >> 4);
                const uint32_t d_tmem = tmem_base + (acc * 4 + cls) * BN;
                if (elect_one()) {
#pragma unroll
                  for (int k = 0; k < BK / 16; ++k)
                    umma_bf16_ss(d_tmem, a_t + 2 * k, b_d + 2 * k, idesc, (((started >> cls) & 1u) | k) != 0 ? 1u : 0u);
                  umma_commit(&empty_b[sb]);
                }
                __syncwarp();
                started |= 1u << cls;
                if (++sb == p.b_slots) { sb = 0; pb ^= 1u; }
              }
            }
          } else if (kind == 0) {
            for (int tap = 0; tap < 9; ++tap) {
              mbar_wait(&full_b[sb], pb);
              tc_fence_after();
              const uint64_t b_d = b_hi + (smem_u32(smem_b + sb * B_BYTES) >> 4);
              if (elect_one()) {
#pragma unroll
                for (int cls = 0; cls < 4; ++cls) {
                  const uint64_t a_t = a_d + ((static_cast<uint32_t>(p.patch[cls][0].off_by_tap[tap]) * ROW_BYTES) >> 4);
                  const uint32_t d_tmem = tmem_base + (acc * 4 + cls) * BN;
#pragma unroll
                  for (int k = 0; k < BK / 16; ++k)
                    umma_bf16_ss(d_tmem, a_t + 2 * k, b_d + 2 * k, idesc, (((started >> cls) & 1u) | k) != 0 ? 1u : 0u);
                }
                umma_commit(&empty_b[sb]);
              }
              __syncwarp();
              started = 0xFu;
              if (++sb == p.b_slots) { sb = 0; pb ^= 1u; }
            }
          } else {
            for (int cls = 0; cls < 4; ++cls) {
              const PatchDesc& pd = p.patch[cls][kind];
              const uint32_t d_tmem = tmem_base + (acc * 4 + cls) * BN;
              for (int j = 0; j < pd.ntaps; ++j) {
                mbar_wait(&full_b[sb], pb);
                tc_fence_after();
                const uint64_t b_d = b_hi + (smem_u32(smem_b + sb * B_BYTES) >> 4);
                const uint64_t a_t = a_d + ((static_cast<uint32_t>(pd.off[j]) * ROW_BYTES) >> 4);
                if (elect_one()) {
#pragma unroll
                  for (int k = 0; k < BK / 16; ++k)
                    umma_bf16_ss(d_tmem, a_t + 2 * k, b_d + 2 * k, idesc, (((started >> cls) & 1u) | k) != 0 ? 1u : 0u);
                  umma_commit(&empty_b[sb]);
                }
                __syncwarp();
                started |= 1u << cls;
                if (++sb == p.b_slots) { sb = 0; pb ^= 1u; }
              }
            }
          }
          if (elect_one()) umma_commit(&empty_a[sa]);
          __syncwarp();
          if (++sa == A_SLOTS) { sa = 0; pa ^= 1u; }
        }
        if (elect_one()) umma_commit(&tmem_full[acc]);
        __syncwarp();
        if (NBUF == 2) { if ((acc ^= 1) == 0) pacc ^= 1u; } else { pacc ^= 1u; }
      }
    }
  } else {
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    int acc = 0;
    uint32_t pacc = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
      int n, h0, w0;
      region(tile, n, h0, w0);
      mbar_wait(&tmem_full[acc], pacc);
      tc_fence_after();
#pragma unroll 1
      for (int cls = 0; cls < 4; ++cls) {
        const int oy = 2 * (h0 + (row >> 3)) + (cls >> 1), ox = 2 * (w0 + (row & 7)) + (cls & 1);
        const int64_t out_off = ((static_cast<int64_t>(n) * p.H + oy) * p.W + ox) * p.C_out;
        const uint32_t t_row = tmem_base + (acc * 4 + cls) * BN + (static_cast<uint32_t>(quarter * 32) << 16);
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
          store_bf16x16(p.y + out_off + c0, f);
        }
      }
      tc_fence_before();
      mbar_arrive(&tmem_empty[acc]);
      if (NBUF == 2) { if ((acc ^= 1) == 0) pacc ^= 1u; } else { pacc ^= 1u; }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

template <int BN>
int launch_halo_quad(const CUtensorMap& tm_a, const CUtensorMap& tm_s, const CUtensorMap& tm_b, HaloParams& p,
                     cudaStream_t s) {
  constexpr int B_BYTES = BN * BK * 2;
  const int budget = 111 * 1024 - 1024 - 512;
  p.a_slots = 3;
  int b_slots = (budget - 3 * p.a_slot_bytes) / B_BYTES;
  if (b_slots < 4) { p.a_slots = 2; b_slots = (budget - 2 * p.a_slot_bytes) / B_BYTES; }
  if (b_slots > MAX_B) b_slots = MAX_B;
  if (b_slots < 2) return DT_ERR_UNSUPPORTED;
  p.b_slots = b_slots;
  p.total_tiles = p.tiles_per_class;                   // one tile = the four classes of a region
  const int smem = p.a_slots * p.a_slot_bytes + b_slots * B_BYTES + 1024 + 512;
  static std::once_flag once;
  static cudaError_t attr_err = cudaSuccess;
  std::call_once(once, [] {
    attr_err = cudaFuncSetAttribute(conv_halo_quad_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, 225 * 1024);
  });
  DT_CUDA(attr_err);
  const int slots = dt_num_sms() * 2;
  const int grid = p.total_tiles < slots ? p.total_tiles : slots;
  conv_halo_quad_kernel<BN><<<grid, kThreads, smem, s>>>(tm_a, tm_s, tm_b, p);
  DT_LAUNCH_CHECK();
  return DT_OK;
}

// tap lists and patch-window offsets of every (class, patch kind)
void fill_patch_tables(HaloParams& p) {
  memset(p.patch, 0, sizeof(p.patch));
  for (int cls = 0; cls < 4; ++cls) {
    const int a = cls >> 1, b = cls & 1;
    for (int tap = 0; tap < 9; ++tap) {
      const int fr = tap / 3, fs = tap % 3;
      // x patch: plain conv reads the window shifted by (fr-1, fs-1); a parity class reads the low-res patch at
      // floor((parity + tap - 1) / 2)
      PatchDesc& px = p.patch[cls][0];
      const int dy = p.parity ? ((a + fr - 1) >> 1) : fr - 1, dx = p.parity ? ((b + fs - 1) >> 1) : fs - 1;
      px.tap[px.ntaps] = static_cast<uint8_t>(tap);
      px.off[px.ntaps] = static_cast<uint8_t>((dy + 1) * PITCH + dx + 1);
      px.off_by_tap[tap] = px.off[px.ntaps];
      ++px.ntaps;
      if (p.parity) {  // skip operand: plane (pr, pc) of the full-res tensor, shifted by (oy - pr) / 2
        const int oy = a + fr - 1, ox = b + fs - 1;
        const int pr = oy & 1, pc = ox & 1;
        PatchDesc& ps = p.patch[cls][1 + pr * 2 + pc];
        ps.pr = static_cast<uint8_t>(pr); ps.pc = static_cast<uint8_t>(pc);
        ps.tap[ps.ntaps] = static_cast<uint8_t>(tap);
        ps.off[ps.ntaps] = static_cast<uint8_t>(((oy - pr) / 2 + 1) * PITCH + (ox - pc) / 2 + 1);
        ++ps.ntaps;
      }
    }
  }
}

}  // namespace

int dt_encode_bf16_map(CUtensorMap* tm, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                       const uint32_t* box, const uint32_t* elem_strides);

// Returns DT_ERR_UNSUPPORTED when the layer does not fit the halo scheme (caller falls back to conv_tc.cu).
int dt_conv_halo(const dt_conv_desc* d, int BN, const void* x, const void* skip, const void* w, int Kpad,
                 const float* scale, const float* shift, const void* residual, void* y, cudaStream_t s) {
  const int C_s = d->C_in - d->C_x;
  const bool wide = d->C_in % BK == 0 && d->C_x % BK == 0;
  const bool narrow = (d->C_in == 32 || d->C_in == 16) && C_s == 0;
  const int parity = d->upsample ? 1 : 0;
  const int Hg = parity ? d->H / 2 : d->H, Wg = parity ? d->W / 2 : d->W;
  if (d->R != 3 || d->S != 3 || d->stride != 1 || d->pad != 1 || !(wide || narrow) || (!parity && C_s != 0) ||
      Wg % TW != 0 || Hg % TH != 0)
    return DT_ERR_UNSUPPORTED;
  HaloParams p;
  memset(&p, 0, sizeof(p));
  const int cw = wide ? BK : d->C_in;
  p.N = d->N; p.H = d->H; p.W = d->W; p.C_in = d->C_in; p.C_x = d->C_x; p.C_out = d->C_out;
  p.parity = parity; p.Hg = Hg; p.Wg = Wg;
  p.n_xslab = d->C_x / cw;
  p.n_sslab = C_s / BK;
  p.relu = d->relu; p.has_residual = d->has_residual;
  p.tiles_w = Wg / TW; p.tiles_h = Hg / TH; p.n_tiles = d->C_out / BN;
  p.tiles_per_class = p.tiles_w * p.tiles_h * d->N;
  p.total_tiles = p.tiles_per_class * (parity ? 4 : 1) * p.n_tiles;
  p.a_slot_bytes = (PATCH_PIX * cw * 2 + 1023) / 1024 * 1024;
  p.residual = static_cast<const __nv_bfloat16*>(residual);
  p.y = static_cast<__nv_bfloat16*>(y);
  p.scale = scale; p.shift = shift;
  fill_patch_tables(p);

  CUtensorMap tm_a, tm_s, tm_b;
  memset(&tm_s, 0, sizeof(tm_s));
  {
    const uint64_t dims[2] = {static_cast<uint64_t>(Kpad), static_cast<uint64_t>(d->C_out)};
    const uint64_t strides[1] = {static_cast<uint64_t>(Kpad) * 2};
    const uint32_t box[2] = {BK, static_cast<uint32_t>(BN)};
    int rc = dt_encode_bf16_map(&tm_b, w, 2, dims, strides, box, nullptr);
    if (rc != DT_OK) return rc;
  }
  {  // x operand at the resolution it is stored in (low-res for parity tiles)
    const uint64_t dims[4] = {static_cast<uint64_t>(d->C_x), static_cast<uint64_t>(Wg), static_cast<uint64_t>(Hg),
                              static_cast<uint64_t>(d->N)};
    const uint64_t strides[3] = {static_cast<uint64_t>(d->C_x) * 2, static_cast<uint64_t>(Wg) * d->C_x * 2,
                                 static_cast<uint64_t>(Hg) * Wg * d->C_x * 2};
    const uint32_t box[4] = {static_cast<uint32_t>(cw), PITCH, TH + 2, 1};
    int rc = dt_encode_bf16_map(&tm_a, x, 4, dims, strides, box, nullptr);
    if (rc != DT_OK) return rc;
  }
  if (p.n_sslab > 0) {  // skip operand: full-res tensor walked with traversal stride 2 (one parity plane per box)
    const uint64_t dims[4] = {static_cast<uint64_t>(C_s), static_cast<uint64_t>(d->W), static_cast<uint64_t>(d->H),
                              static_cast<uint64_t>(d->N)};
    const uint64_t strides[3] = {static_cast<uint64_t>(C_s) * 2, static_cast<uint64_t>(d->W) * C_s * 2,
                                 static_cast<uint64_t>(d->H) * d->W * C_s * 2};
    const uint32_t box[4] = {BK, 2 * PITCH, 2 * (TH + 2), 1};
    const uint32_t estr[4] = {1, 2, 2, 1};
    int rc = dt_encode_bf16_map(&tm_s, skip, 4, dims, strides, box, estr);
    if (rc != DT_OK) return rc;
  }
  if (d->flags & DT_CONV_UPS_FOLDED) {
    // x operand with folded up-sampling weights: only the class-fused kernel knows that packing
    if (!(parity && wide && !d->has_residual && d->C_out == BN && (BN == 32 || BN == 64)) || Kpad != 16 * d->C_x + 9 * C_s)
      return DT_ERR_UNSUPPORTED;
    p.fold = 1;
    p.k_skip = 16 * d->C_x;
  }
  if (parity && wide && !d->has_residual && d->C_out == BN && (p.fold || !(d->flags & DT_CONV_NO_QUAD))) {
    if (BN == 32) return launch_halo_quad<32>(tm_a, tm_s, tm_b, p, s);
    if (BN == 64) return launch_halo_quad<64>(tm_a, tm_s, tm_b, p, s);
  }
#define DT_HALO(BNV, CWV) \
  if (BN == BNV && cw == CWV) return launch_halo<BNV, CWV>(tm_a, tm_s, tm_b, p, s);
  DT_HALO(16, 64) DT_HALO(32, 64) DT_HALO(64, 64) DT_HALO(128, 64) DT_HALO(256, 64)
  DT_HALO(16, 32) DT_HALO(32, 32) DT_HALO(64, 32)
  DT_HALO(16, 16) DT_HALO(32, 16) DT_HALO(64, 16)
#undef DT_HALO
  return DT_ERR_UNSUPPORTED;
}
