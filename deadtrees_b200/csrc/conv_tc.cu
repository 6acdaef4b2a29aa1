// Implicit-GEMM convolution on the sm_100a tensor cores (tcgen05.mma, accumulators in TMEM).
//
// GEMM view: M = N*Ho*Wo output pixels (128 per CTA), N = C_out (BN per CTA), K = R*S*C_in in
// chunks of 64 bf16 (= one 128-byte swizzled smem row per output pixel / output channel).
//   B (weights [C_out][Kpad], K-major)  : TMA 2-D tiles, SWIZZLE_128B
//   A (activations, NHWC bf16)          : either TMA 4-D halo boxes (stride-1, single-source convs: the
//                                         box for filter tap (r,s) is the output box shifted by
//                                         (r-pad, s-pad); out-of-bounds rows/cols are zero-filled by
//                                         TMA = the conv zero padding), or a generic gather producer
//                                         (4 warps) for strided convs, the 7x7 stem and the decoder's
//                                         virtual cat([nearest_x2(x), skip]) input.
//   D                                   : 128 x BN fp32 in TMEM; epilogue = scale*acc+shift
//                                         (+residual) (+ReLU) -> bf16 NHWC.
// Warp roles: w0 = TMA producer, w1 = TMEM allocator + MMA issuer, w2..5 = epilogue, w6..9 = A gather.
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

constexpr int BM = 128;            // output pixels per CTA (UMMA M)
constexpr int BK = 64;             // bf16 elements per K chunk (128 bytes)
constexpr int A_STAGE_BYTES = BM * BK * 2;
constexpr int kThreadsTma = 192;   // warps 0..5
constexpr int kThreadsGather = 320;  // + warps 6..9

struct TcParams {
  int H, W, C_in, C_x, C_s, ups, Hx, Wx;
  int Ho, Wo, C_out, R, S, stride, pad;
  int relu, has_residual, stem;
  int num_k_chunks, chunks_per_tap;
  int M_total, n_tiles;
  const __nv_bfloat16* x;
  const __nv_bfloat16* skip;
  const __nv_bfloat16* residual;
  __nv_bfloat16* y;
  const float* scale;
  const float* shift;
};

template <int BN>
struct TcCfg {
  static constexpr int B_STAGE_BYTES = BN * BK * 2;
  static constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
  static constexpr int STAGES_RAW = (100 * 1024) / STAGE_BYTES;
  static constexpr int STAGES = STAGES_RAW > 6 ? 6 : (STAGES_RAW < 2 ? 2 : STAGES_RAW);
  static constexpr int TMEM_COLS = BN < 32 ? 32 : BN;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*alignment slack*/ + 256 /*barriers*/;
};

template <int BN, bool A_TMA>
__global__ void __launch_bounds__(A_TMA ? kThreadsTma : kThreadsGather, 2)
conv_tc_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b, const TcParams p) {
  using Cfg = TcCfg<BN>;
  constexpr int STAGES = Cfg::STAGES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + STAGES * A_STAGE_BYTES;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * Cfg::STAGE_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full_bar = empty_bar + STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_tile = blockIdx.x % p.n_tiles;
  const int m_tile = blockIdx.x / p.n_tiles;
  const int m0 = m_tile * BM;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_b);
    if (A_TMA) tma_prefetch_desc(&tm_a);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], A_TMA ? 1u : 129u);  // expect_tx arrival (+128 gather threads)
      mbar_init(&empty_bar[s], 1u);                // one tcgen05.commit
    }
    mbar_init(tmem_full_bar, 1u);
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer (one lane) =====================
    if (lane == 0) {
      int n0 = 0, h0 = 0, w0 = 0;
      if (A_TMA) {
        w0 = m0 % p.Wo;
        h0 = (m0 / p.Wo) % p.Ho;
        n0 = m0 / (p.Wo * p.Ho);
      }
      int stage = 0;
      uint32_t phase = 0;
      for (int kc = 0; kc < p.num_k_chunks; ++kc) {
        mbar_wait(&empty_bar[stage], phase ^ 1u);
        mbar_arrive_expect_tx(&full_bar[stage], A_TMA ? Cfg::STAGE_BYTES : Cfg::B_STAGE_BYTES);
        tma_load_2d(smem_b + stage * Cfg::B_STAGE_BYTES, &tm_b, &full_bar[stage], kc * BK, n_tile * BN);
        if (A_TMA) {
          const int tap = kc / p.chunks_per_tap, cc = kc - tap * p.chunks_per_tap;
          const int fr = tap / p.S, fs = tap - fr * p.S;
          tma_load_4d(smem_a + stage * A_STAGE_BYTES, &tm_a, &full_bar[stage], cc * BK, w0 + fs - p.pad,
                      h0 + fr - p.pad, n0);
        }
        if (++stage == STAGES) { stage = 0; phase ^= 1u; }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (one lane) =====================
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(BM, BN);
      int stage = 0;
      uint32_t phase = 0;
      for (int kc = 0; kc < p.num_k_chunks; ++kc) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        const uint32_t a_addr = smem_u32(smem_a + stage * A_STAGE_BYTES);
        const uint32_t b_addr = smem_u32(smem_b + stage * Cfg::B_STAGE_BYTES);
#pragma unroll
        for (int k = 0; k < BK / 16; ++k) {
          umma_bf16_ss(tmem_base, umma_desc_sw128(a_addr + k * 32, 1024), umma_desc_sw128(b_addr + k * 32, 1024),
                       idesc, (kc | k) != 0 ? 1u : 0u);
        }
        umma_commit(&empty_bar[stage]);  // frees this smem stage once the MMAs above have read it
        if (++stage == STAGES) { stage = 0; phase ^= 1u; }
      }
      umma_commit(tmem_full_bar);        // accumulator complete
    }
  } else if (warp < 6) {
    // ===================== epilogue: TMEM -> regs -> global =====================
    const int quarter = warp & 3;        // TMEM lanes [32*quarter, 32*quarter+32) are visible to this warp
    const int row = quarter * 32 + lane;
    const int m = m0 + row;
    const bool valid = m < p.M_total;
    mbar_wait(tmem_full_bar, 0);
    tc_fence_after();
    const int64_t out_off = static_cast<int64_t>(m) * p.C_out + n_tile * BN;
#pragma unroll 1
    for (int c0 = 0; c0 < BN; c0 += 16) {
      uint32_t v[16];
      tmem_ld_x16(tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + c0, v);
      tmem_ld_wait();
      float f[16];
      const int co = n_tile * BN + c0;
#pragma unroll
      for (int j = 0; j < 16; ++j) f[j] = fmaf(__uint_as_float(v[j]), __ldg(p.scale + co + j), __ldg(p.shift + co + j));
      if (valid) {
        if (p.has_residual) {
          const uint4* rp = reinterpret_cast<const uint4*>(p.residual + out_off + c0);
          const uint4 r0 = __ldg(rp), r1 = __ldg(rp + 1);
          const uint32_t rr[8] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w};
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
        uint4* op = reinterpret_cast<uint4*>(p.y + out_off + c0);
        op[0] = make_uint4(pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]), pack_bf16x2(f[4], f[5]),
                           pack_bf16x2(f[6], f[7]));
        op[1] = make_uint4(pack_bf16x2(f[8], f[9]), pack_bf16x2(f[10], f[11]), pack_bf16x2(f[12], f[13]),
                           pack_bf16x2(f[14], f[15]));
      }
    }
  } else if (!A_TMA) {
    // ===================== A gather producer (128 threads) =====================
    const int tg = threadIdx.x - 192;
    const int j = tg & 7;        // 16-byte granule within the 128-byte K chunk
    const int rsub = tg >> 3;    // rows rsub + 16*it
    int hb[8], wb[8], nb[8];
#pragma unroll
    for (int it = 0; it < 8; ++it) {
      const int m = m0 + rsub + 16 * it;
      if (m < p.M_total) {
        const int wo = m % p.Wo, t = m / p.Wo;
        hb[it] = (t % p.Ho) * p.stride - p.pad;
        wb[it] = wo * p.stride - p.pad;
        nb[it] = t / p.Ho;
      } else {
        hb[it] = -100000; wb[it] = 0; nb[it] = 0;   // every tap lands out of bounds -> zeros
      }
    }
    int stage = 0;
    uint32_t phase = 0;
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
          const int hi = hb[it] + fr, wi = wb[it] + fs;
          val[it] = make_uint4(0u, 0u, 0u, 0u);
          if (kvalid && hi >= 0 && hi < p.H && wi >= 0 && wi < p.W) {
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
        const int row = rsub + 16 * it;
        *reinterpret_cast<uint4*>(a_stage + row * 128 + ((j ^ (row & 7)) << 4)) = val[it];
      }
      fence_proxy_async_smem();   // generic-proxy stores -> visible to the tensor-core (async) proxy
      mbar_arrive(&full_bar[stage]);
      if (++stage == STAGES) { stage = 0; phase ^= 1u; }
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

int encode_bf16_map(CUtensorMap* tm, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                    const uint32_t* box) {
  EncodeTiledFn fn = get_encode_fn();
  DT_REQUIRE(fn != nullptr, DT_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, static_cast<cuuint32_t>(rank), const_cast<void*>(base),
                  reinterpret_cast<const cuuint64_t*>(dims), reinterpret_cast<const cuuint64_t*>(strides_bytes),
                  reinterpret_cast<const cuuint32_t*>(box), estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  DT_REQUIRE(r == CUDA_SUCCESS, DT_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", static_cast<int>(r));
  return DT_OK;
}

template <int BN, bool A_TMA>
int launch_tc(const CUtensorMap& tm_a, const CUtensorMap& tm_b, const TcParams& p, int grid, cudaStream_t s) {
  using Cfg = TcCfg<BN>;
  static std::once_flag once;
  static cudaError_t attr_err = cudaSuccess;
  std::call_once(once, [] {
    attr_err = cudaFuncSetAttribute(conv_tc_kernel<BN, A_TMA>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    Cfg::SMEM_BYTES);
  });
  DT_CUDA(attr_err);
  conv_tc_kernel<BN, A_TMA><<<grid, A_TMA ? kThreadsTma : kThreadsGather, Cfg::SMEM_BYTES, s>>>(tm_a, tm_b, p);
  DT_LAUNCH_CHECK();
  return DT_OK;
}

// A is TMA-loadable when output boxes of 128 pixels are rectangular in (w, h, n) and map 1:1 onto input boxes.
bool tma_eligible(const dt_conv_desc* d, int Ho, int Wo) {
  if (d->stride != 1 || d->upsample || d->C_x != d->C_in || d->C_in % BK != 0) return false;
  if (Ho != d->H || Wo != d->W) return false;
  if (Wo >= BM) return Wo % BM == 0;
  if (BM % Wo != 0) return false;
  const int rows = BM / Wo;
  if (Ho >= rows) return Ho % rows == 0;
  return rows % Ho == 0;
}

}  // namespace

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
  const int Ho = (d->H + 2 * d->pad - d->R) / d->stride + 1;
  const int Wo = (d->W + 2 * d->pad - d->S) / d->stride + 1;
  DT_REQUIRE(Ho > 0 && Wo > 0, DT_ERR_BAD_SHAPE, "dt_conv2d_fwd: empty output");
  cudaStream_t s = static_cast<cudaStream_t>(stream);

  const int stem = (d->C_in == 4 && d->R == 7 && d->S == 7) ? 1 : 0;
  const int Ktot = stem ? 7 * 32 : d->R * d->S * d->C_in;
  const int Kpad = stem ? 256 : (Ktot + BK - 1) / BK * BK;

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
  int BN = d->C_out;
  if (BN > 128) BN = (d->C_out % 256 == 0 && static_cast<int64_t>(d->N) * Ho * Wo >= 148LL * 2 * BM) ? 256 : 128;
  DT_REQUIRE(d->C_out % BN == 0 && (BN == 16 || BN == 32 || BN == 64 || BN == 128 || BN == 256), DT_ERR_BAD_SHAPE,
             "dt_conv2d_fwd: unsupported C_out %d", d->C_out);

  TcParams p;
  p.H = d->H; p.W = d->W; p.C_in = d->C_in; p.C_x = d->C_x; p.C_s = d->C_in - d->C_x; p.ups = d->upsample ? 1 : 0;
  p.Hx = d->upsample ? d->H / 2 : d->H; p.Wx = d->upsample ? d->W / 2 : d->W;
  p.Ho = Ho; p.Wo = Wo; p.C_out = d->C_out; p.R = d->R; p.S = d->S; p.stride = d->stride; p.pad = d->pad;
  p.relu = d->relu; p.has_residual = d->has_residual; p.stem = stem;
  p.num_k_chunks = Kpad / BK;
  p.chunks_per_tap = stem ? 1 : (d->C_in >= BK ? d->C_in / BK : 1);
  const int64_t M = static_cast<int64_t>(d->N) * Ho * Wo;
  DT_REQUIRE(M < (1LL << 31) - BM, DT_ERR_BAD_SHAPE, "dt_conv2d_fwd: too many output pixels");
  p.M_total = static_cast<int>(M);
  p.n_tiles = d->C_out / BN;
  p.x = static_cast<const __nv_bfloat16*>(x);
  p.skip = static_cast<const __nv_bfloat16*>(skip);
  p.residual = static_cast<const __nv_bfloat16*>(residual);
  p.y = static_cast<__nv_bfloat16*>(y);
  p.scale = scale; p.shift = shift;
  const int m_tiles = static_cast<int>((M + BM - 1) / BM);
  const int grid = m_tiles * p.n_tiles;

  CUtensorMap tm_a, tm_b;
  memset(&tm_a, 0, sizeof(tm_a));
  {
    const uint64_t dims[2] = {static_cast<uint64_t>(Kpad), static_cast<uint64_t>(d->C_out)};
    const uint64_t strides[1] = {static_cast<uint64_t>(Kpad) * 2};
    const uint32_t box[2] = {BK, static_cast<uint32_t>(BN)};
    int rc = encode_bf16_map(&tm_b, w, 2, dims, strides, box);
    if (rc != DT_OK) return rc;
  }
  const bool use_tma = !(d->flags & DT_CONV_FORCE_GATHER) && !stem && tma_eligible(d, Ho, Wo);
  if (use_tma) {
    const int bw = Wo >= BM ? BM : Wo;
    const int bh = (BM / bw) >= Ho ? Ho : BM / bw;
    const int bn = BM / (bw * bh);
    const uint64_t dims[4] = {static_cast<uint64_t>(d->C_in), static_cast<uint64_t>(d->W), static_cast<uint64_t>(d->H),
                              static_cast<uint64_t>(d->N)};
    const uint64_t strides[3] = {static_cast<uint64_t>(d->C_in) * 2, static_cast<uint64_t>(d->W) * d->C_in * 2,
                                 static_cast<uint64_t>(d->H) * d->W * d->C_in * 2};
    const uint32_t box[4] = {BK, static_cast<uint32_t>(bw), static_cast<uint32_t>(bh), static_cast<uint32_t>(bn)};
    int rc = encode_bf16_map(&tm_a, x, 4, dims, strides, box);
    if (rc != DT_OK) return rc;
  }

#define DT_TC(BNV)                                                                             \
  case BNV:                                                                                    \
    return use_tma ? launch_tc<BNV, true>(tm_a, tm_b, p, grid, s) : launch_tc<BNV, false>(tm_a, tm_b, p, grid, s);
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
