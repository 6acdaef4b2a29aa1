// 3x3 / stride-1 / pad-1 convolution with a shared-memory resident input halo (tcgen05 implicit GEMM).
//
// One tile = 8 (w) x 16 (h) output pixels of one image = the 128 rows of a UMMA M=128 tile.  For every slab of
// `cw` input channels (64, or all of them when C_in is 32 / 16) ONE TMA box brings the (16+2) x (8+2) pixel halo
// patch into shared memory; the nine filter taps then read it in place: the A-operand descriptor of tap (r,s)
// starts (r*pitch + s) pixel rows into the patch and steps `pitch` pixel rows between 8-pixel groups
// (stride-byte-offset), so the activations cross L2 -> SM once instead of nine times.  Weights stream through a
// separate TMA ring, one 64-wide K chunk (tap, slab) at a time.
//   warp 0: TMA producer (A patches + B chunks)   warp 1: TMEM alloc + MMA issuer   warps 2..5: epilogue
// Persistent CTAs, one per SM, double-buffered TMEM accumulator (epilogue of tile i overlaps tile i+1).
// K is accumulated slab-major (slab, tap, channel) - a different fp32 summation order than conv_tc.cu.
#include <cstring>
#include <mutex>

#include "common.cuh"

namespace {

constexpr int BM = 128;
constexpr int BK = 64;
constexpr int TW = 8, TH = 16;       // tile of output pixels
constexpr int kThreads = 192;
constexpr int MAX_A = 3;
constexpr int MAX_B = 8;
constexpr int PITCH = 10;            // halo patch row pitch in pixels (TW + 2)

struct HaloParams {
  int N, H, W, C_in, C_out;
  int cw;                // channels per slab (row bytes = 2*cw)
  int nslab;             // C_in / cw
  int chunks_per_slab;   // B chunks (64 K elements) per slab: 9 (cw = 64) or ceil(9*C_in/64)
  int k_slab;            // K elements per slab = 9 * cw
  int pitch;             // halo patch row pitch in pixels (10; 16 only in scripts/halo_exp.py)
  int base_off;          // experiment: fill the descriptor's base-offset field from the start address
  int a_slots;           // 2 or 3 halo patch slots
  int relu, has_residual;
  int tiles_w, tiles_h, n_tiles, total_tiles;
  int b_slots, a_slot_bytes;
  const __nv_bfloat16* residual;
  __nv_bfloat16* y;
  const float* scale;
  const float* shift;
};

struct Geo {
  int n_tile, n, h0, w0;
};
__device__ __forceinline__ Geo geo(const HaloParams& p, int tile) {
  Geo g;
  g.n_tile = tile % p.n_tiles;
  int m = tile / p.n_tiles;
  g.w0 = (m % p.tiles_w) * TW; m /= p.tiles_w;
  g.h0 = (m % p.tiles_h) * TH;
  g.n = m / p.tiles_h;
  return g;
}

__device__ __forceinline__ uint64_t umma_desc_bo(uint32_t addr, uint32_t sbo, uint32_t layout, uint32_t base_off) {
  return umma_desc(addr, sbo, layout) | (static_cast<uint64_t>(base_off & 7u) << 49);
}

// issue the MMAs of one 64-wide K chunk; all offsets are compile-time when EXPERIMENT is false
template <int BN, int CW, int Q, bool EXPERIMENT>
__device__ __forceinline__ void issue_chunk(const HaloParams& p, uint32_t d_tmem, uint64_t a_d, uint32_t a_base,
                                            uint64_t b_d, uint32_t first) {
  constexpr uint32_t idesc = umma_idesc_bf16(BM, BN);
  constexpr int K_SLAB = 9 * CW;
  constexpr int KSTEPS = (K_SLAB - Q * BK) / 16 < BK / 16 ? (K_SLAB - Q * BK) / 16 : BK / 16;
#pragma unroll
  for (int k = 0; k < KSTEPS; ++k) {
    constexpr int dummy = 0; (void)dummy;
    const int kk = Q * BK + k * 16;                    // K index inside the slab: tap * CW + channel
    const int tap = kk / CW, ch = kk % CW;
    const int fr = tap / 3, fs = tap % 3;
    const uint32_t acc = (Q == 0 && k == 0) ? (first ? 0u : 1u) : 1u;
    if (EXPERIMENT) {
      const uint32_t a_addr = a_base + (fr * p.pitch + fs) * (CW * 2) + ch * 2;
      uint64_t d = umma_desc(a_addr, p.pitch * CW * 2, CW == 64 ? 2u : (CW == 32 ? 4u : 6u));
      if (p.base_off) d |= static_cast<uint64_t>((a_addr >> 7) & 7u) << 49;
      umma_bf16_ss(d_tmem, d, b_d + 2 * k, idesc, acc);
    } else {
      umma_bf16_ss(d_tmem, a_d + (((fr * PITCH + fs) * (CW * 2) + ch * 2) >> 4), b_d + 2 * k, idesc, acc);
    }
  }
}

template <int BN, int CW, bool EXPERIMENT>
__global__ void __launch_bounds__(kThreads, BN <= 128 ? 2 : 1)
conv_halo_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b,
                 const HaloParams p) {
  constexpr int B_BYTES = BN * BK * 2;
  constexpr int TMEM_COLS = 2 * BN < 32 ? 32 : 2 * BN;
  constexpr int CHUNKS = (9 * CW + BK - 1) / BK;
  const int A_SLOTS = p.a_slots;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;                                   // A_SLOTS x a_slot_bytes (1024-aligned each)
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
  const uint32_t tmem_base = *tmem_slot;
  constexpr int row_bytes = CW * 2;

  if (warp == 0) {
    if (lane == 0) {
      // The A patch of slab i+1 is requested before the weight chunks of slab i so that it is in flight while the
      // tensor core works through slab i (three A slots: previous / current / next).
      int sa = 0, sb = 0;
      uint32_t pa = 0, pb = 0;
      const int halo_bytes = (TH + 2) * p.pitch * row_bytes;
      bool primed = false;
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
        const Geo g = geo(p, tile);
        for (int slab = 0; slab < p.nslab; ++slab) {
          if (!primed) {  // very first patch of this CTA
            mbar_wait(&empty_a[sa], pa ^ 1u);
            mbar_arrive_expect_tx(&full_a[sa], halo_bytes);
            tma_load_4d(smem_a + sa * p.a_slot_bytes, &tm_a, &full_a[sa], slab * CW, g.w0 - 1, g.h0 - 1, g.n);
            if (++sa == A_SLOTS) { sa = 0; pa ^= 1u; }
            primed = true;
          }
          // With three patch slots the next patch (next slab, or slab 0 of this CTA's next tile) is requested
          // before this slab's weight chunks; with two slots its slot is still being read, so it is requested
          // after the first few weight chunks instead (by then the previous slab has been consumed).
          const int prefetch_at = A_SLOTS >= 3 ? 0 : (CHUNKS > 5 ? 4 : CHUNKS - 1);
          for (int q = 0; q < CHUNKS; ++q) {
            if (q == prefetch_at) {
              int nslab_i = slab + 1, ntile = tile;
              if (nslab_i == p.nslab) { nslab_i = 0; ntile = tile + gridDim.x; }
              if (ntile < p.total_tiles) {
                const Geo gn = geo(p, ntile);
                mbar_wait(&empty_a[sa], pa ^ 1u);
                mbar_arrive_expect_tx(&full_a[sa], halo_bytes);
                tma_load_4d(smem_a + sa * p.a_slot_bytes, &tm_a, &full_a[sa], nslab_i * CW, gn.w0 - 1, gn.h0 - 1, gn.n);
                if (++sa == A_SLOTS) { sa = 0; pa ^= 1u; }
              }
            }
            // wide: chunk q = tap q of this slab (k = q*C_in + slab*64); narrow: chunk q of the single slab
            const int kcoord = CW == BK ? q * p.C_in + slab * BK : q * BK;
            mbar_wait(&empty_b[sb], pb ^ 1u);
            mbar_arrive_expect_tx(&full_b[sb], B_BYTES);
            tma_load_2d(smem_b + sb * B_BYTES, &tm_b, &full_b[sb], kcoord, g.n_tile * BN);
            if (++sb == p.b_slots) { sb = 0; pb ^= 1u; }
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // This thread's instruction stream paces the tensor core: descriptor high words are loop invariants
      // and, with the chunk loop unrolled, every tap / k-step offset is an immediate.
      const uint64_t a_hi = umma_desc(0u, PITCH * row_bytes, CW == 64 ? 2u : (CW == 32 ? 4u : 6u));
      const uint64_t b_hi = umma_desc(0u, 1024u, 2u);
      int sa = 0, sb = 0, acc = 0;
      uint32_t pa = 0, pb = 0, pacc = 0;
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
        mbar_wait(&tmem_empty[acc], pacc ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int slab = 0; slab < p.nslab; ++slab) {
          mbar_wait(&full_a[sa], pa);
          const uint32_t a_base = smem_u32(smem_a + sa * p.a_slot_bytes);
          const uint64_t a_d = a_hi + (a_base >> 4);
          const uint32_t first = slab == 0 ? 1u : 0u;
#define DT_CHUNK(Q)                                                                          \
          if (Q < CHUNKS) {                                                                  \
            mbar_wait(&full_b[sb], pb);                                                      \
            tc_fence_after();                                                                \
            const uint64_t b_d = b_hi + (smem_u32(smem_b + sb * B_BYTES) >> 4);              \
            issue_chunk<BN, CW, (Q < CHUNKS ? Q : 0), EXPERIMENT>(p, d_tmem, a_d, a_base, b_d, first); \
            umma_commit(&empty_b[sb]);                                                       \
            if (++sb == p.b_slots) { sb = 0; pb ^= 1u; }                                     \
          }
          DT_CHUNK(0) DT_CHUNK(1) DT_CHUNK(2) DT_CHUNK(3) DT_CHUNK(4) DT_CHUNK(5) DT_CHUNK(6) DT_CHUNK(7) DT_CHUNK(8)
#undef DT_CHUNK
          umma_commit(&empty_a[sa]);
          if (++sa == A_SLOTS) { sa = 0; pa ^= 1u; }
        }
        umma_commit(&tmem_full[acc]);
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
      const int64_t out_off =
          ((static_cast<int64_t>(g.n) * p.H + g.h0 + (row >> 3)) * p.W + g.w0 + (row & 7)) * p.C_out + g.n_tile * BN;
      uint4 res[SC / 8];
      if (p.has_residual) {
#pragma unroll
        for (int j = 0; j < SC / 8; ++j) res[j] = __ldg(reinterpret_cast<const uint4*>(p.residual + out_off) + j);
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
          for (int j = 0; j < SC / 8; ++j)
            res_next[j] = __ldg(reinterpret_cast<const uint4*>(p.residual + out_off + s0 + SC) + j);
        }
#pragma unroll
        for (int c0 = 0; c0 < SC; c0 += 16) {
          uint32_t v[16];
          tmem_ld_x16(t_row + s0 + c0, v);
          tmem_ld_wait();
          float f[16];
          const int co = g.n_tile * BN + s0 + c0;
#pragma unroll
          for (int j = 0; j < 16; ++j)
            f[j] = fmaf(__uint_as_float(v[j]), __ldg(p.scale + co + j), __ldg(p.shift + co + j));
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
          uint4* op = reinterpret_cast<uint4*>(p.y + out_off + s0 + c0);
          op[0] = make_uint4(pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]), pack_bf16x2(f[4], f[5]),
                             pack_bf16x2(f[6], f[7]));
          op[1] = make_uint4(pack_bf16x2(f[8], f[9]), pack_bf16x2(f[10], f[11]), pack_bf16x2(f[12], f[13]),
                             pack_bf16x2(f[14], f[15]));
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

template <int BN, int CW, bool EXPERIMENT>
int launch_halo_impl(const CUtensorMap& tm_a, const CUtensorMap& tm_b, HaloParams& p, cudaStream_t s) {
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
    attr_err = cudaFuncSetAttribute(conv_halo_kernel<BN, CW, EXPERIMENT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    225 * 1024);
  });
  DT_CUDA(attr_err);
  const int slots = dt_num_sms() * CTAS;
  const int grid = p.total_tiles < slots ? p.total_tiles : slots;
  conv_halo_kernel<BN, CW, EXPERIMENT><<<grid, kThreads, smem, s>>>(tm_a, tm_b, p);
  DT_LAUNCH_CHECK();
  return DT_OK;
}

template <int BN, int CW>
int launch_halo(const CUtensorMap& tm_a, const CUtensorMap& tm_b, HaloParams& p, cudaStream_t s) {
  if (p.pitch != PITCH || p.base_off) return launch_halo_impl<BN, CW, true>(tm_a, tm_b, p, s);
  return launch_halo_impl<BN, CW, false>(tm_a, tm_b, p, s);
}

}  // namespace

int dt_encode_bf16_map(CUtensorMap* tm, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                       const uint32_t* box, const uint32_t* elem_strides);

// Returns DT_ERR_UNSUPPORTED when the layer does not fit the halo scheme (caller falls back to conv_tc.cu).
int dt_conv_halo(const dt_conv_desc* d, int BN, const void* x, const void* w, int Kpad, const float* scale,
                 const float* shift, const void* residual, void* y, cudaStream_t s) {
  const bool wide = d->C_in % BK == 0;
  const bool narrow = d->C_in == 32 || d->C_in == 16;
  if (d->R != 3 || d->S != 3 || d->stride != 1 || d->pad != 1 || d->upsample || d->C_x != d->C_in ||
      !(wide || narrow) || d->W % TW != 0 || d->H % TH != 0)
    return DT_ERR_UNSUPPORTED;
  HaloParams p;
  memset(&p, 0, sizeof(p));
  p.N = d->N; p.H = d->H; p.W = d->W; p.C_in = d->C_in; p.C_out = d->C_out;
  p.cw = wide ? BK : d->C_in;
  p.nslab = d->C_in / p.cw;
  p.k_slab = 9 * p.cw;
  p.chunks_per_slab = (p.k_slab + BK - 1) / BK;
  p.pitch = (d->flags & DT_CONV_HALO_P16) ? 16 : 10;
  p.base_off = (d->flags & DT_CONV_HALO_BASEOFF) ? 1 : 0;
  p.relu = d->relu; p.has_residual = d->has_residual;
  p.tiles_w = d->W / TW; p.tiles_h = d->H / TH; p.n_tiles = d->C_out / BN;
  p.total_tiles = p.tiles_w * p.tiles_h * d->N * p.n_tiles;
  p.a_slot_bytes = ((TH + 2) * p.pitch * p.cw * 2 + 1023) / 1024 * 1024;
  p.residual = static_cast<const __nv_bfloat16*>(residual);
  p.y = static_cast<__nv_bfloat16*>(y);
  p.scale = scale; p.shift = shift;

  CUtensorMap tm_a, tm_b;
  {
    const uint64_t dims[2] = {static_cast<uint64_t>(Kpad), static_cast<uint64_t>(d->C_out)};
    const uint64_t strides[1] = {static_cast<uint64_t>(Kpad) * 2};
    const uint32_t box[2] = {BK, static_cast<uint32_t>(BN)};
    int rc = dt_encode_bf16_map(&tm_b, w, 2, dims, strides, box, nullptr);
    if (rc != DT_OK) return rc;
  }
  {
    const uint64_t dims[4] = {static_cast<uint64_t>(d->C_in), static_cast<uint64_t>(d->W), static_cast<uint64_t>(d->H),
                              static_cast<uint64_t>(d->N)};
    const uint64_t strides[3] = {static_cast<uint64_t>(d->C_in) * 2, static_cast<uint64_t>(d->W) * d->C_in * 2,
                                 static_cast<uint64_t>(d->H) * d->W * d->C_in * 2};
    const uint32_t box[4] = {static_cast<uint32_t>(p.cw), static_cast<uint32_t>(p.pitch), TH + 2, 1};
    int rc = dt_encode_bf16_map(&tm_a, x, 4, dims, strides, box, nullptr);
    if (rc != DT_OK) return rc;
  }
#define DT_HALO(BNV, CWV) \
  if (BN == BNV && p.cw == CWV) return launch_halo<BNV, CWV>(tm_a, tm_b, p, s);
  DT_HALO(16, 64) DT_HALO(32, 64) DT_HALO(64, 64) DT_HALO(128, 64) DT_HALO(256, 64)
  DT_HALO(16, 32) DT_HALO(32, 32) DT_HALO(64, 32)
  DT_HALO(16, 16) DT_HALO(32, 16) DT_HALO(64, 16)
#undef DT_HALO
  return DT_ERR_UNSUPPORTED;
}
