// Implicit-GEMM convolution on the sm_100a tensor cores (tcgen05.mma, accumulators in TMEM).
//
// GEMM view: M = output pixels (128 per tile), N = C_out (BN per tile), K = R*S*C_in in chunks of 64 bf16.
//   B (weights [C_out][Kpad], K-major) : TMA 2-D tiles, SWIZZLE_128B
//   A (activations, NHWC bf16)         : TMA 4-D boxes, one per filter tap and 64-channel slab.  The box for
//       tap (r,s) is the tile's pixel box shifted by (r-pad, s-pad); rows/cols outside the image are
//       zero-filled by TMA (= the conv zero padding).  Variants:
//         - stride 2          : tensor map with traversal stride 2 (elementStrides)
//         - C_in = 32 / 16    : 64 B / 32 B rows (SWIZZLE_64B / SWIZZLE_32B), 2 / 4 taps per K chunk
//         - nearest-x2 + cat  : tiles are built per output-pixel parity class (h%2, w%2); for one class the
//                               up-sampled operand is a plain box of the low-res tensor shifted by
//                               ((a+r-1)>>1, (b+s-1)>>1) and the skip operand a stride-2 box of the skip tensor,
//                               so the concatenated / up-sampled tensor is never materialised
//       or a generic gather producer (4 warps) for everything else (7x7 stem, odd shapes).
//   D : 128 x BN fp32 in TMEM, double-buffered; epilogue = scale*acc+shift (+residual) (+ReLU) -> bf16 NHWC.
// Persistent CTAs (static round-robin over tiles); warp roles: w0 = TMA producer, w1 = TMEM allocator + MMA
// issuer, w2..5 = epilogue, w6..9 = A gather (gather mode only).  The epilogue of tile i overlaps the main loop
// of tile i+1 through the second TMEM accumulator.
//
// Replaces the cuDNN convolutions the reference reaches through smp.Unet.forward
// (deadtrees/network/segmodel.py:214, deadtrees/deployment/inference.py:60); layer list in SURVEY.md App. A.
#include <cstring>
#include <mutex>

#include "common.cuh"

int dt_conv2d_direct(const dt_conv_desc* d, int Ho, int Wo, int Kpad, int stem, const void* x, const void* skip,
                     const void* w, const float* scale, const float* shift, const void* residual, void* y,
                     cudaStream_t s);

namespace {

constexpr int BM = 128;            // output pixels per tile (UMMA M)
constexpr int BK = 64;             // bf16 elements per K chunk (128 bytes of weights per output channel)
constexpr int A_STAGE_BYTES = BM * BK * 2;
constexpr int kThreadsTma = 192;   // warps 0..5
constexpr int kThreadsGather = 320;  // + warps 6..9
constexpr int MAX_STAGES = 8;

struct TcParams {
  int H, W, C_in, C_x, C_s, ups, Hx, Wx;  // virtual input size, channel split, x storage size
  int Ho, Wo, C_out, R, S, stride, pad;
  int relu, has_residual, stem;
  int num_k_chunks, k_total;              // Kpad / 64, un-padded K
  int a_cw;                               // channels per A TMA load: min(64, C_in)
  int transposed;                         // 1: data gradient of a stride-2 conv (gather producer): x is gy, taps scatter
  int parity;                             // 1: tiles enumerate output-pixel parity classes (upsample convs)
  int Hg, Wg;                             // pixel grid the M tiles walk (output grid, or low-res grid in parity mode)
  int M_lim;                              // pixels per class (parity) or in total
  int m_tiles_per_class;
  int n_tiles, total_tiles, stages;
  const __nv_bfloat16* x;
  const __nv_bfloat16* skip;
  const __nv_bfloat16* residual;
  __nv_bfloat16* y;
  const float* scale;
  const float* shift;
};

struct TileGeo {
  int n_tile, m0, a, b;
};

__device__ __forceinline__ TileGeo tile_geo(const TcParams& p, int tile) {
  TileGeo g;
  g.n_tile = tile % p.n_tiles;
  int m_tile = tile / p.n_tiles;
  int cls = 0;
  if (p.parity) {   // the four classes of one region are consecutive tiles (shared operands stay in L2)
    cls = m_tile & 3;
    m_tile >>= 2;
  }
  g.a = cls >> 1;
  g.b = cls & 1;
  g.m0 = m_tile * BM;
  return g;
}

// element offset of output pixel `m` (index inside the tile's class) in y / residual, channel 0
__device__ __forceinline__ int64_t out_pixel_offset(const TcParams& p, const TileGeo& g, int m) {
  if (!p.parity) return static_cast<int64_t>(m) * p.C_out;
  const int wl = m % p.Wg, t = m / p.Wg;
  const int hl = t % p.Hg, n = t / p.Hg;
  return ((static_cast<int64_t>(n) * p.Ho + 2 * hl + g.a) * p.Wo + 2 * wl + g.b) * p.C_out;
}

template <int BN>
struct TcCfg {
  static constexpr int B_STAGE_BYTES = BN * BK * 2;
  static constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
  // BN = 256 needs all 512 TMEM columns for the two accumulators -> one CTA per SM, deeper ring
  static constexpr int SMEM_BUDGET = BN == 256 ? 200 * 1024 : 100 * 1024;
  static constexpr int STAGES_RAW = SMEM_BUDGET / STAGE_BYTES;
  static constexpr int MAX_ST = STAGES_RAW > MAX_STAGES ? MAX_STAGES : STAGES_RAW;
  static constexpr int CTAS_PER_SM = BN == 256 ? 1 : 2;
  static constexpr int TMEM_COLS = 2 * BN < 32 ? 32 : 2 * BN;
  static constexpr int BAR_BYTES = 256;
  static constexpr int smem_bytes(int stages) { return stages * STAGE_BYTES + 1024 + BAR_BYTES; }
};

template <int BN, bool A_TMA>
__global__ void __launch_bounds__(A_TMA ? kThreadsTma : kThreadsGather, TcCfg<BN>::CTAS_PER_SM)
conv_tc_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_s,
               const __grid_constant__ CUtensorMap tm_b, const TcParams p) {
  using Cfg = TcCfg<BN>;
  const int STAGES = p.stages;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + STAGES * A_STAGE_BYTES;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * Cfg::STAGE_BYTES);
  uint64_t* empty_bar = full_bar + MAX_STAGES;
  uint64_t* tmem_full_bar = empty_bar + MAX_STAGES;   // [2]
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;       // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_b);
    if (A_TMA) {
      tma_prefetch_desc(&tm_a);
      if (p.C_s > 0) tma_prefetch_desc(&tm_s);
    }
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], A_TMA ? 1u : 129u);  // expect_tx arrival (+128 gather threads)
      mbar_init(&empty_bar[s], 1u);                // one tcgen05.commit
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full_bar[i], 1u);            // one tcgen05.commit per tile
      mbar_init(&tmem_empty_bar[i], 128u);         // every epilogue thread
    }
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);   // warp-uniform for the compiler (uniform registers)

  if (warp == 0) {
    // ===================== TMA producer (one lane) =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      const int taps = p.R * p.S;
      const int sub_bytes = BM * p.a_cw * 2;                 // one A box
      const int subs_per_chunk = BK / p.a_cw;                // 1, 2 or 4 taps per K chunk
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
        const TileGeo g = tile_geo(p, tile);
        int n0 = 0, h0 = 0, w0 = 0;
        if (A_TMA) {
          w0 = g.m0 % p.Wg;
          h0 = (g.m0 / p.Wg) % p.Hg;
          n0 = g.m0 / (p.Wg * p.Hg);
        }
        for (int kc = 0; kc < p.num_k_chunks; ++kc) {
          mbar_wait(&empty_bar[stage], phase ^ 1u);
          uint8_t* a_dst = smem_a + stage * A_STAGE_BYTES;
          if (A_TMA) {
            int tap0, ci0, nsub;
            if (p.a_cw == BK) {
              tap0 = (kc * BK) / p.C_in;
              ci0 = kc * BK - tap0 * p.C_in;
              nsub = 1;
            } else {
              tap0 = kc * subs_per_chunk;
              ci0 = 0;
              nsub = min(subs_per_chunk, taps - tap0);
            }
            mbar_arrive_expect_tx(&full_bar[stage], Cfg::B_STAGE_BYTES + nsub * sub_bytes);
            for (int j = 0; j < nsub; ++j) {
              const int tap = tap0 + j;
              const int fr = tap / p.S, fs = tap - fr * p.S;
              if (ci0 < p.C_x) {
                int cx, cy;
                if (p.parity) {  // nearest-x2 source: low-res box shifted by floor((parity + tap - 1) / 2)
                  cx = w0 + ((g.b + fs - 1) >> 1);
                  cy = h0 + ((g.a + fr - 1) >> 1);
                } else {
                  cx = w0 * p.stride + fs - p.pad;
                  cy = h0 * p.stride + fr - p.pad;
                }
                tma_load_4d(a_dst + j * sub_bytes, &tm_a, &full_bar[stage], ci0, cx, cy, n0);
              } else {           // skip operand of a parity tile: stride-2 box of the full-res tensor
                tma_load_4d(a_dst + j * sub_bytes, &tm_s, &full_bar[stage], ci0 - p.C_x, 2 * w0 + g.b + fs - 1,
                            2 * h0 + g.a + fr - 1, n0);
              }
            }
          } else {
            mbar_arrive_expect_tx(&full_bar[stage], Cfg::B_STAGE_BYTES);
          }
          tma_load_2d(smem_b + stage * Cfg::B_STAGE_BYTES, &tm_b, &full_bar[stage], kc * BK, g.n_tile * BN);
          if (++stage == STAGES) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (whole warp keeps the uniform state, an elected lane issues) =====================
    {
      constexpr uint32_t idesc = umma_idesc_bf16(BM, BN);
      // A operand layout: 128 B rows (swizzle 128), or 64 B / 32 B rows with one sub-tile per tap.
      // This thread's instruction stream paces the tensor core, so everything that does not change per
      // MMA is hoisted: descriptor high words and the four per-k-step A offsets are loop invariants.
      const bool gather = !A_TMA;
      const int cw = gather ? BK : p.a_cw;
      const uint32_t a_layout = cw == 64 ? 2u : (cw == 32 ? 4u : 6u);
      const int sub_bytes = BM * cw * 2;
      const uint64_t a_hi = umma_desc(0u, 8u * cw * 2u, a_layout);
      const uint64_t b_hi = umma_desc(0u, 1024u, 2u);
      uint32_t a_off[BK / 16];
#pragma unroll
      for (int k = 0; k < BK / 16; ++k) a_off[k] = (((k * 16) / cw) * sub_bytes + ((k * 16) % cw) * 2) >> 4;
      const int tail_steps = gather ? BK / 16 : (p.k_total - (p.num_k_chunks - 1) * BK) / 16;
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
        mbar_wait(&tmem_empty_bar[acc], acc_phase ^ 1u);   // epilogue has drained this accumulator
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kc = 0; kc < p.num_k_chunks; ++kc) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint64_t a_d = a_hi + (smem_u32(smem_a + stage * A_STAGE_BYTES) >> 4);
          const uint64_t b_d = b_hi + (smem_u32(smem_b + stage * Cfg::B_STAGE_BYTES) >> 4);
          // gather mode zero-fills the padded tail of K; TMA mode simply skips it
          const int ksteps = kc + 1 < p.num_k_chunks ? BK / 16 : tail_steps;
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < BK / 16; ++k) {
              if (k < ksteps) umma_bf16_ss(d_tmem, a_d + a_off[k], b_d + 2 * k, idesc, (kc | k) != 0 ? 1u : 0u);
            }
            umma_commit(&empty_bar[stage]);  // frees this smem stage once the MMAs above have read it
            if (kc + 1 == p.num_k_chunks) umma_commit(&tmem_full_bar[acc]);  // accumulator complete
          }
          __syncwarp();
          if (++stage == STAGES) { stage = 0; phase ^= 1u; }
        }
        if ((acc ^= 1) == 0) acc_phase ^= 1u;
      }
    }
  } else if (warp < 6) {
    // ===================== epilogue: TMEM -> regs -> global =====================
    const int quarter = warp & 3;        // TMEM lanes [32*quarter, 32*quarter+32) are visible to this warp
    const int row = quarter * 32 + lane;
    constexpr int SC = BN < 64 ? BN : 64;  // columns per super-chunk (one 128 B line of bf16 per row)
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
      const TileGeo g = tile_geo(p, tile);
      const int m = g.m0 + row;
      const bool valid = m < p.M_lim;
      const int64_t out_off = out_pixel_offset(p, g, valid ? m : 0) + g.n_tile * BN;
      uint4 res[SC / 8];
      if (p.has_residual && valid) {       // issue the first residual line before waiting for the MMAs
#pragma unroll
        for (int j = 0; j < SC / 16; ++j) ldg_v8(p.residual + out_off + 16 * j, res[2 * j], res[2 * j + 1]);
      }
      mbar_wait(&tmem_full_bar[acc], acc_phase);
      tc_fence_after();
      const uint32_t t_row = tmem_base + acc * BN + (static_cast<uint32_t>(quarter * 32) << 16);
#pragma unroll 1
      for (int s0 = 0; s0 < BN; s0 += SC) {
        uint4 res_next[SC / 8];
        const bool more = s0 + SC < BN;
        if (p.has_residual && valid && more) {
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
          if (valid) {
            store_bf16x16(p.y + out_off + s0 + c0, f);
          }
        }
        if (more) {
#pragma unroll
          for (int j = 0; j < SC / 8; ++j) res[j] = res_next[j];
        }
      }
      tc_fence_before();
      mbar_arrive(&tmem_empty_bar[acc]);   // all TMEM reads of this accumulator are complete (wait::ld above)
      if ((acc ^= 1) == 0) acc_phase ^= 1u;
    }
  } else if (!A_TMA) {
    // ===================== A gather producer (128 threads) =====================
    const int tg = threadIdx.x - 192;
    const int j = tg & 7;        // 16-byte granule within the 128-byte K chunk
    const int rsub = tg >> 3;    // rows rsub + 16*it
    int stage = 0;
    uint32_t phase = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
      const TileGeo g = tile_geo(p, tile);
      int hb[8], wb[8], nb[8];
#pragma unroll
      for (int it = 0; it < 8; ++it) {
        const int m = g.m0 + rsub + 16 * it;
        if (m < p.M_lim) {
          const int wo = m % p.Wo, t = m / p.Wo;
          if (p.transposed) {   // output pixel (h, w) of the data gradient: source row = (h + pad - r) / stride
            hb[it] = (t % p.Ho) + p.pad;
            wb[it] = wo + p.pad;
          } else {
            hb[it] = (t % p.Ho) * p.stride - p.pad;
            wb[it] = wo * p.stride - p.pad;
          }
          nb[it] = t / p.Ho;
        } else {
          hb[it] = -100000; wb[it] = 0; nb[it] = 0;   // every tap lands out of bounds -> zeros
        }
      }
      for (int kc = 0; kc < p.num_k_chunks; ++kc) {
        const int k0 = kc * BK + j * 8;
        uint4 val[8];
        if (p.stem) {
          // k = r*32 + s*4 + c, C_in == 4: one granule = filter columns s0, s0+1 (4 channels each)
          const int fr = k0 >> 5, s0 = (k0 & 31) >> 2;
          const bool kvalid = fr < p.R;
#pragma unroll
          for (int it = 0; it < 8; ++it) {
            const int hi = hb[it] + fr, wi = wb[it] + s0;
            uint2 lo = make_uint2(0u, 0u), hi2 = make_uint2(0u, 0u);
            if (kvalid && hi >= 0 && hi < p.H) {
              const __nv_bfloat16* src = p.x + ((static_cast<int64_t>(nb[it]) * p.H + hi) * p.W + wi) * 4;
              if (wi >= 0 && wi < p.W) lo = __ldg(reinterpret_cast<const uint2*>(src));
              if (wi + 1 >= 0 && wi + 1 < p.W) hi2 = __ldg(reinterpret_cast<const uint2*>(src + 4));
            }
            val[it] = make_uint4(lo.x, lo.y, hi2.x, hi2.y);
          }
        } else {
          const int tap = k0 / p.C_in, ci = k0 - tap * p.C_in;
          const int fr = tap / p.S, fs = tap - fr * p.S;
          const bool kvalid = tap < p.R * p.S;
          const bool from_x = ci < p.C_x;
#pragma unroll
          for (int it = 0; it < 8; ++it) {
            int hi = hb[it] + fr, wi = wb[it] + fs;
            bool ok = kvalid;
            if (p.transposed) {   // stride-2 data gradient: only taps whose source coordinate is even contribute
              hi = hb[it] - fr; wi = wb[it] - fs;
              ok = ok && hi >= 0 && wi >= 0 && ((hi | wi) & 1) == 0;
              hi >>= 1; wi >>= 1;
            }
            val[it] = make_uint4(0u, 0u, 0u, 0u);
            if (ok && hi >= 0 && hi < p.H && wi >= 0 && wi < p.W) {
              const __nv_bfloat16* src =
                  from_x ? p.x + ((static_cast<int64_t>(nb[it]) * p.Hx + (hi >> p.ups)) * p.Wx + (wi >> p.ups)) * p.C_x + ci
                         : p.skip + ((static_cast<int64_t>(nb[it]) * p.H + hi) * p.W + wi) * p.C_s + (ci - p.C_x);
              val[it] = ld_nc_v4(src);
            }
          }
        }
        mbar_wait(&empty_bar[stage], phase ^ 1u);
        uint8_t* a_stage = smem_a + stage * A_STAGE_BYTES;
#pragma unroll
        for (int it = 0; it < 8; ++it) {
          const int r = rsub + 16 * it;
          *reinterpret_cast<uint4*>(a_stage + r * 128 + ((j ^ (r & 7)) << 4)) = val[it];
        }
        fence_proxy_async_smem();   // generic-proxy stores -> visible to the tensor-core (async) proxy
        mbar_arrive(&full_bar[stage]);
        if (++stage == STAGES) { stage = 0; phase ^= 1u; }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

// ---- host side ---------------------------------------------------------------------------------

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

// bf16 tensor map; inner box = `box[0]` elements (<= 64) selects the swizzle width (32 / 64 / 128 bytes)
int encode_bf16_map(CUtensorMap* tm, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                    const uint32_t* box, const uint32_t* elem_strides) {
  EncodeTiledFn fn = get_encode_fn();
  DT_REQUIRE(fn != nullptr, DT_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  if (elem_strides)
    for (int i = 0; i < rank; ++i) estr[i] = elem_strides[i];
  const CUtensorMapSwizzle sw = box[0] * 2 >= 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                                : (box[0] * 2 == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
  CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, static_cast<cuuint32_t>(rank), const_cast<void*>(base),
                  reinterpret_cast<const cuuint64_t*>(dims), reinterpret_cast<const cuuint64_t*>(strides_bytes),
                  reinterpret_cast<const cuuint32_t*>(box), estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  DT_REQUIRE(r == CUDA_SUCCESS, DT_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d (rank %d, box %u %u)",
             static_cast<int>(r), rank, box[0], box[1]);
  return DT_OK;
}

template <int BN, bool A_TMA>
int launch_tc(const CUtensorMap& tm_a, const CUtensorMap& tm_s, const CUtensorMap& tm_b, TcParams& p,
              cudaStream_t s) {
  using Cfg = TcCfg<BN>;
  static std::once_flag once;
  static cudaError_t attr_err = cudaSuccess;
  std::call_once(once, [] {
    attr_err = cudaFuncSetAttribute(conv_tc_kernel<BN, A_TMA>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    Cfg::smem_bytes(Cfg::MAX_ST));
  });
  DT_CUDA(attr_err);
  p.stages = Cfg::MAX_ST;
  const int slots = dt_num_sms() * Cfg::CTAS_PER_SM;
  const int grid = p.total_tiles < slots ? p.total_tiles : slots;
  conv_tc_kernel<BN, A_TMA><<<grid, A_TMA ? kThreadsTma : kThreadsGather, Cfg::smem_bytes(p.stages), s>>>(tm_a, tm_s,
                                                                                                          tm_b, p);
  DT_LAUNCH_CHECK();
  return DT_OK;
}

// 128-pixel tiles must be rectangular boxes (bw x bh x bn) of the (N, Hg, Wg) pixel grid
bool box_tiling(int Hg, int Wg, int* bw, int* bh, int* bn) {
  if (Wg >= BM) {
    if (Wg % BM) return false;
    *bw = BM; *bh = 1; *bn = 1;
    return true;
  }
  if (BM % Wg) return false;
  const int rows = BM / Wg;
  *bw = Wg;
  if (Hg >= rows) {
    if (Hg % rows) return false;
    *bh = rows; *bn = 1;
  } else {
    if (rows % Hg) return false;
    *bh = Hg; *bn = rows / Hg;
  }
  return true;
}

}  // namespace

int dt_encode_bf16_map(CUtensorMap* tm, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                       const uint32_t* box, const uint32_t* elem_strides) {
  return encode_bf16_map(tm, base, rank, dims, strides_bytes, box, elem_strides);
}
int dt_conv_halo(const dt_conv_desc* d, int BN, const void* x, const void* skip, const void* w, int Kpad,
                 const float* scale, const float* shift, const void* residual, void* y, cudaStream_t s);

int dt_conv_stem(const dt_conv_desc* d, const void* x, const void* w, const float* scale, const float* shift, void* y,
                 void* pooled, cudaStream_t s);
int dt_conv_res(const dt_conv_desc* d, const void* x, const void* w, int Kpad, const float* scale, const float* shift,
                const void* residual, void* y, cudaStream_t s);
int dt_conv_pair(const dt_conv_desc* d, const void* x, const void* w, int Kpad, const float* scale, const float* shift,
                 const void* residual, void* y, cudaStream_t s);
int dt_conv_row(const dt_conv_desc* d, const void* x, const void* w, int Kpad, const float* scale, const float* shift,
                const void* residual, void* y, cudaStream_t s);

extern "C" int dt_conv2d_fwd(const dt_conv_desc* d, const void* x, const void* skip, const void* w,
                             const float* scale, const float* shift, const void* residual, void* y,
                             dt_stream_t stream) {
  DT_ARCH_GUARD();
  DT_REQUIRE(d != nullptr, DT_ERR_BAD_SHAPE, "dt_conv2d_fwd: null descriptor");
  DT_REQUIRE(d->N > 0 && d->H > 0 && d->W > 0 && d->C_in > 0 && d->C_out > 0 && d->R > 0 && d->S > 0 &&
                 d->stride > 0 && d->pad >= 0 && d->C_x > 0 && d->C_x <= d->C_in,
             DT_ERR_BAD_SHAPE, "dt_conv2d_fwd: bad descriptor");
  DT_REQUIRE(!d->upsample || (d->H % 2 == 0 && d->W % 2 == 0), DT_ERR_BAD_SHAPE,
             "dt_conv2d_fwd: upsample needs even H, W");
  DT_REQUIRE(d->C_x == d->C_in || skip != nullptr, DT_ERR_BAD_SHAPE, "dt_conv2d_fwd: skip tensor missing");
  DT_REQUIRE(!d->has_residual || residual != nullptr, DT_ERR_BAD_SHAPE, "dt_conv2d_fwd: residual tensor missing");
  DT_REQUIRE(d->dtype == DT_BF16 || d->dtype == DT_F32, DT_ERR_BAD_SHAPE, "dt_conv2d_fwd: dtype %d", d->dtype);
  const bool transposed = (d->flags & DT_CONV_TRANSPOSED) != 0;
  // transposed: desc H, W are the OUTPUT size (the conv input whose gradient is computed); x is gy at the conv's output size
  const int Hsrc = (d->H + 2 * d->pad - d->R) / d->stride + 1, Wsrc = (d->W + 2 * d->pad - d->S) / d->stride + 1;
  const int Ho = transposed ? d->H : Hsrc, Wo = transposed ? d->W : Wsrc;
  DT_REQUIRE(Ho > 0 && Wo > 0 && Hsrc > 0 && Wsrc > 0, DT_ERR_BAD_SHAPE, "dt_conv2d_fwd: empty output");
  DT_REQUIRE(!transposed || (d->dtype == DT_BF16 && d->stride == 2 && !d->upsample && d->C_x == d->C_in &&
                             !(d->flags & (DT_CONV_X_PAD3 | DT_CONV_FORCE_DIRECT))),
             DT_ERR_BAD_SHAPE, "dt_conv2d_fwd: DT_CONV_TRANSPOSED is the bf16 data gradient of a stride-2 conv");
  cudaStream_t s = static_cast<cudaStream_t>(stream);

  const int stem = (d->C_in == 4 && d->R == 7 && d->S == 7) ? 1 : 0;
  const int Ktot = stem ? 7 * 32 : d->R * d->S * d->C_in;
  const int Kpad = stem ? 256 : (Ktot + BK - 1) / BK * BK;

  if (d->flags & DT_CONV_X_PAD3) {
    DT_REQUIRE(stem && d->dtype == DT_BF16 && d->stride == 2 && d->pad == 3, DT_ERR_BAD_SHAPE,
               "dt_conv2d_fwd: DT_CONV_X_PAD3 is the bf16 7x7/s2 stem layout");
    DT_REQUIRE(reinterpret_cast<uintptr_t>(y) % 32 == 0, DT_ERR_BAD_ALIGN,
               "dt_conv2d_fwd: the bf16 output must be 32-byte aligned (256-bit epilogue stores)");
    const int rc = dt_conv_stem(d, x, w, scale, shift, y, nullptr, s);
    DT_REQUIRE(rc != DT_ERR_UNSUPPORTED, DT_ERR_BAD_SHAPE, "dt_conv2d_fwd: stem output %dx%d cannot be tiled", Ho, Wo);
    return rc;
  }
  if (d->dtype == DT_F32 || (d->flags & DT_CONV_FORCE_DIRECT))
    return dt_conv2d_direct(d, Ho, Wo, Kpad, stem, x, skip, w, scale, shift, residual, y, s);

  // ---- tensor-core path ----
  DT_REQUIRE(stem || (d->C_in % 8 == 0 && d->C_x % 8 == 0), DT_ERR_BAD_SHAPE,
             "dt_conv2d_fwd: bf16 path needs C_in, C_x multiples of 8 (got %d, %d)", d->C_in, d->C_x);
  DT_REQUIRE(!stem || (!d->upsample && d->C_x == 4), DT_ERR_BAD_SHAPE, "dt_conv2d_fwd: bad stem descriptor");
  DT_REQUIRE(d->C_out % 16 == 0, DT_ERR_BAD_SHAPE, "dt_conv2d_fwd: bf16 path needs C_out %% 16 == 0 (got %d)", d->C_out);
  DT_REQUIRE((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y) | reinterpret_cast<uintptr_t>(w) |
              reinterpret_cast<uintptr_t>(skip) | reinterpret_cast<uintptr_t>(residual)) % 16 == 0,
             DT_ERR_BAD_ALIGN, "dt_conv2d_fwd: tensors must be 16-byte aligned");
  DT_REQUIRE((reinterpret_cast<uintptr_t>(y) | reinterpret_cast<uintptr_t>(residual)) % 32 == 0, DT_ERR_BAD_ALIGN,
             "dt_conv2d_fwd: the bf16 output and residual must be 32-byte aligned (256-bit epilogue accesses)");
  const int64_t M = static_cast<int64_t>(d->N) * Ho * Wo;
  DT_REQUIRE(M < (1LL << 31) - BM, DT_ERR_BAD_SHAPE, "dt_conv2d_fwd: too many output pixels");
  int BN = d->C_out;
  if (BN > 128) BN = (d->C_out % 256 == 0 && M >= static_cast<int64_t>(dt_num_sms()) * 2 * BM) ? 256 : 128;
  while (BN > 16 && d->C_out % BN != 0) BN >>= 1;   // e.g. 192 output channels (a data gradient): 3 tiles of 64
  DT_REQUIRE(d->C_out % BN == 0 && (BN == 16 || BN == 32 || BN == 64 || BN == 128 || BN == 256), DT_ERR_BAD_SHAPE,
             "dt_conv2d_fwd: unsupported C_out %d", d->C_out);

  if (d->flags & DT_CONV_UPS_FOLDED) {
    // up-sampling folded into per-class weights (dt_pack_conv_weight mode 5): only the resident-weight parity kernel
    DT_REQUIRE(d->upsample && !stem && !transposed, DT_ERR_BAD_SHAPE, "dt_conv2d_fwd: DT_CONV_UPS_FOLDED needs an up-sampled input");
    const int rcf = d->C_x == d->C_in
                        ? dt_conv_res(d, x, w, 16 * d->C_in, scale, shift, residual, y, s)
                        : dt_conv_halo(d, BN, x, skip, w, 16 * d->C_x + 9 * (d->C_in - d->C_x), scale, shift, residual, y, s);
    DT_REQUIRE(rcf != DT_ERR_UNSUPPORTED, DT_ERR_BAD_SHAPE, "dt_conv2d_fwd: no folded kernel for %d -> %d channels at %dx%d",
               d->C_in, d->C_out, d->H, d->W);
    return rcf;
  }
  if (!(d->flags & (DT_CONV_NO_HALO | DT_CONV_FORCE_GATHER)) && !stem && !transposed) {
    if (!(d->flags & DT_CONV_NO_ROW) && dt_row_kernels_enabled()) {   // narrow layers: row streaming, vertical taps in N
      const int rcr = dt_conv_row(d, x, w, Kpad, scale, shift, residual, y, s);
      if (rcr != DT_ERR_UNSUPPORTED) return rcr;
    }
    const int rc0 = dt_conv_res(d, x, w, Kpad, scale, shift, residual, y, s);   // weights resident in smem
    if (rc0 != DT_ERR_UNSUPPORTED) return rc0;
    if (d->flags & DT_CONV_PAIR) {      // wide layers on CTA pairs (tcgen05.mma.cta_group::2)
      const int rcp = dt_conv_pair(d, x, w, Kpad, scale, shift, residual, y, s);
      if (rcp != DT_ERR_UNSUPPORTED) return rcp;
    }
    const int rc = dt_conv_halo(d, BN, x, skip, w, Kpad, scale, shift, residual, y, s);
    if (rc != DT_ERR_UNSUPPORTED) return rc;   // not a 3x3/s1 layer of a fitting shape: per-tap path below
  }

  TcParams p;
  memset(&p, 0, sizeof(p));
  p.H = d->H; p.W = d->W; p.C_in = d->C_in; p.C_x = d->C_x; p.C_s = d->C_in - d->C_x; p.ups = d->upsample ? 1 : 0;
  p.Hx = d->upsample ? d->H / 2 : d->H; p.Wx = d->upsample ? d->W / 2 : d->W;
  if (transposed) { p.transposed = 1; p.H = p.Hx = Hsrc; p.W = p.Wx = Wsrc; }
  p.Ho = Ho; p.Wo = Wo; p.C_out = d->C_out; p.R = d->R; p.S = d->S; p.stride = d->stride; p.pad = d->pad;
  p.relu = d->relu; p.has_residual = d->has_residual; p.stem = stem;
  p.num_k_chunks = Kpad / BK;
  p.k_total = Ktot;
  p.a_cw = d->C_in >= BK ? BK : d->C_in;
  p.n_tiles = d->C_out / BN;
  p.x = static_cast<const __nv_bfloat16*>(x);
  p.skip = static_cast<const __nv_bfloat16*>(skip);
  p.residual = static_cast<const __nv_bfloat16*>(residual);
  p.y = static_cast<__nv_bfloat16*>(y);
  p.scale = scale; p.shift = shift;

  // ---- can the A operand come through TMA? ----
  int bw = 0, bh = 0, bn = 0;
  bool use_tma = !(d->flags & DT_CONV_FORCE_GATHER) && !stem && !transposed;
  const bool wide = d->C_in % BK == 0 && d->C_x % BK == 0;                       // 128 B rows, slabs of 64 channels
  const bool narrow = (d->C_in == 32 || d->C_in == 16) && d->C_x == d->C_in;     // 64 B / 32 B rows, one box per tap
  use_tma = use_tma && (wide || narrow) && (d->stride == 1 || d->stride == 2);
  if (use_tma && d->upsample) {
    use_tma = d->R == 3 && d->S == 3 && d->pad == 1 && d->stride == 1 && box_tiling(d->H / 2, d->W / 2, &bw, &bh, &bn);
    p.parity = 1; p.Hg = d->H / 2; p.Wg = d->W / 2;
  } else if (use_tma) {
    use_tma = p.C_s == 0 && box_tiling(Ho, Wo, &bw, &bh, &bn) && bw * d->stride <= 256 && bh * d->stride <= 256;
    p.parity = 0; p.Hg = Ho; p.Wg = Wo;
  }
  if (!use_tma) { p.parity = 0; p.Hg = Ho; p.Wg = Wo; }
  const int64_t m_class = p.parity ? static_cast<int64_t>(d->N) * p.Hg * p.Wg : M;
  p.M_lim = static_cast<int>(m_class);
  p.m_tiles_per_class = static_cast<int>((m_class + BM - 1) / BM);
  p.total_tiles = p.m_tiles_per_class * (p.parity ? 4 : 1) * p.n_tiles;

  CUtensorMap tm_a, tm_s, tm_b;
  memset(&tm_a, 0, sizeof(tm_a));
  memset(&tm_s, 0, sizeof(tm_s));
  {
    const uint64_t dims[2] = {static_cast<uint64_t>(Kpad), static_cast<uint64_t>(d->C_out)};
    const uint64_t strides[1] = {static_cast<uint64_t>(Kpad) * 2};
    const uint32_t box[2] = {BK, static_cast<uint32_t>(BN)};
    int rc = encode_bf16_map(&tm_b, w, 2, dims, strides, box, nullptr);
    if (rc != DT_OK) return rc;
  }
  if (use_tma) {
    const uint32_t st = static_cast<uint32_t>(d->stride);
    const uint64_t dims[4] = {static_cast<uint64_t>(p.C_x), static_cast<uint64_t>(p.Wx), static_cast<uint64_t>(p.Hx),
                              static_cast<uint64_t>(d->N)};
    const uint64_t strides[3] = {static_cast<uint64_t>(p.C_x) * 2, static_cast<uint64_t>(p.Wx) * p.C_x * 2,
                                 static_cast<uint64_t>(p.Hx) * p.Wx * p.C_x * 2};
    const uint32_t box[4] = {static_cast<uint32_t>(p.a_cw), bw * st, bh * st, static_cast<uint32_t>(bn)};
    const uint32_t estr[4] = {1, st, st, 1};
    int rc = encode_bf16_map(&tm_a, x, 4, dims, strides, box, estr);
    if (rc != DT_OK) return rc;
    if (p.C_s > 0) {  // parity tiles read the full-resolution skip tensor with traversal stride 2
      const uint64_t sdims[4] = {static_cast<uint64_t>(p.C_s), static_cast<uint64_t>(d->W), static_cast<uint64_t>(d->H),
                                 static_cast<uint64_t>(d->N)};
      const uint64_t sstr[3] = {static_cast<uint64_t>(p.C_s) * 2, static_cast<uint64_t>(d->W) * p.C_s * 2,
                                static_cast<uint64_t>(d->H) * d->W * p.C_s * 2};
      const uint32_t sbox[4] = {BK, static_cast<uint32_t>(2 * bw), static_cast<uint32_t>(2 * bh),
                                static_cast<uint32_t>(bn)};
      const uint32_t sestr[4] = {1, 2, 2, 1};
      rc = encode_bf16_map(&tm_s, skip, 4, sdims, sstr, sbox, sestr);
      if (rc != DT_OK) return rc;
    }
  }

#define DT_TC(BNV)                                                                             \
  case BNV:                                                                                    \
    return use_tma ? launch_tc<BNV, true>(tm_a, tm_s, tm_b, p, s) : launch_tc<BNV, false>(tm_a, tm_s, tm_b, p, s);
  switch (BN) {
    DT_TC(16)
    DT_TC(32)
    DT_TC(64)
    DT_TC(128)
    DT_TC(256)
  }
#undef DT_TC
  return DT_ERR_UNSUPPORTED;
}

// Stem + resnet.maxpool in one launch (conv_stem.cu, POOL variant): y as dt_conv2d_fwd with DT_CONV_X_PAD3, pooled =
// maxpool3x3/s2/pad1(y).  DT_ERR_UNSUPPORTED when the fused form does not apply (caller: dt_conv2d_fwd + dt_maxpool3x3s2).
extern "C" int dt_stem_pool_fwd(const dt_conv_desc* d, const void* x, const void* w, const float* scale, const float* shift,
                                void* y, void* pooled, dt_stream_t stream) {
  DT_ARCH_GUARD();
  DT_REQUIRE(d && x && w && scale && shift && y && pooled, DT_ERR_BAD_SHAPE, "dt_stem_pool_fwd: null argument");
  DT_REQUIRE(d->C_in == 4 && d->R == 7 && d->S == 7 && d->stride == 2 && d->pad == 3 && d->dtype == DT_BF16 &&
                 (d->flags & DT_CONV_X_PAD3) && !d->has_residual && !d->upsample,
             DT_ERR_BAD_SHAPE, "dt_stem_pool_fwd: not the bf16 7x7/s2 stem on the zero-bordered frame");
  DT_REQUIRE(reinterpret_cast<uintptr_t>(y) % 32 == 0 && reinterpret_cast<uintptr_t>(pooled) % 16 == 0, DT_ERR_BAD_ALIGN,
             "dt_stem_pool_fwd: outputs must be 32- / 16-byte aligned");
  return dt_conv_stem(d, x, w, scale, shift, y, pooled, static_cast<cudaStream_t>(stream));
}
