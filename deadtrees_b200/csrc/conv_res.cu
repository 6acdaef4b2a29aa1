// 3x3 / stride-1 / pad-1 convolution for layers whose whole weight tensor fits in shared memory
// (C_in <= 64 and C_out <= 64: resnet layer1, the decoder's 64/32/16-channel tail) - tcgen05 implicit GEMM.
//
// The packed weights [C_out][9*C_in] are loaded ONCE per CTA and stay resident; the only per-tile traffic is
// one TMA box with the input halo patch.  A "super tile" is 8 (w) x 16*MT (h) output pixels = MT UMMA M=128
// tiles that share one patch of (16*MT+2) x (8+2) pixels and one set of pipeline hand-shakes, so the fixed cost
// per tile (barrier round trips, instruction issue) is amortised over MT*128 pixels - these layers are
// HBM / issue bound, not tensor bound.  Taps read the patch in place through shifted UMMA descriptors
// (see conv_halo.cu).  Parity mode (nearest-x2 up-sampled input without skip, decoder block 4): the MT = 4 M tiles of
// a super tile are the four output-pixel parity classes (a,b) of one 16 x 8 LOW-RES region; class (a,b) reads the
// shared low-res patch at ((a+r-1)>>1, (b+s-1)>>1) and writes pixels (2i+a, 2j+b), so the low-res input is read once.
//
//   warp 0: TMA producer   warp 1: TMEM alloc + MMA issuer   warps 2..5: epilogue (TMEM -> scale/shift/residual/ReLU)
// Persistent CTAs; TMEM holds two buffers of MT accumulators so the epilogue overlaps the next super tile.
#include <cstring>
#include <mutex>

#include "common.cuh"

namespace {

constexpr int BM = 128;
constexpr int BK = 64;
constexpr int TW = 8, TH = 16;
constexpr int kThreads = 192;
constexpr int MAX_A = 4;
constexpr int PITCH = TW + 2;

struct ResParams {
  int N, H, W, C_out;            // H, W: output size
  int parity, Hg, Wg;            // tile grid: (H, W) or (H/2, W/2)
  int a_slots, a_slot_bytes;
  int relu, has_residual;
  int tiles_w, tiles_h, total_tiles;
  const __nv_bfloat16* residual;
  __nv_bfloat16* y;
  const float* scale;
  const float* shift;
  // head epilogue (EPI == 1): the first K of the 16 output channels are the class logits
  int K;
  float* logits_nchw;
  __nv_bfloat16* logits_nhwc;
  uint8_t* mask;
};

struct Geo {
  int n, h0, w0;
};
template <int ROWS_PER_TILE>
__device__ __forceinline__ Geo geo(const ResParams& p, int tile) {
  Geo g;
  int m = tile;
  g.w0 = (m % p.tiles_w) * TW; m /= p.tiles_w;
  g.h0 = (m % p.tiles_h) * ROWS_PER_TILE;
  g.n = m / p.tiles_h;
  return g;
}

// pixel offset of tap (fr, fs)'s window inside the patch for M tile `mt` (compile-time in the unrolled issue loop)
template <bool PAR>
__device__ __forceinline__ constexpr int tap_pixel_offset(int mt, int tap) {
  const int fr = tap / 3, fs = tap % 3;
  if (PAR) {  // mt = parity class (a,b): floor((a + fr - 1) / 2) + 1 rows into the low-res patch
    const int a = mt >> 1, b = mt & 1;
    return (((a + fr + 1) >> 1)) * PITCH + ((b + fs + 1) >> 1);
  }
  return (mt * TH + fr) * PITCH + fs;
}

// FOLD (parity mode): nearest-x2 up-sampling folded into the weights.  For output parity class (a,b) the nine taps of the
// up-sampled operand touch only 2 x 2 low-res pixels, (a-1+ey, b-1+ex); the host sums the taps that share a pixel
// (dt_pack_conv_weight mode 5: K index = ((class * 4 + e) * C_in + ci), e = ey * 2 + ex) and the kernel issues 4 instead
// of 9 taps per class: 16 instead of 36 tap-MMAs per super tile on a layer that is bound by the A-operand reads.
template <int BN, int CW, int MT, bool PAR, int EPI, bool FOLD>
__global__ void __launch_bounds__(kThreads, 1)
conv_res_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b,
                const ResParams p) {
  static_assert(!FOLD || (PAR && MT == 4), "folded weights belong to the parity mode");
  constexpr int B_BYTES = BN * BK * 2;
  constexpr int NCH = FOLD ? (16 * CW) / BK : (9 * CW + BK - 1) / BK;   // weight chunks of 64 K elements
  constexpr int W_BYTES = NCH * B_BYTES;
  constexpr int ROW_BYTES = CW * 2;
  constexpr int TILE_ROWS = PAR ? TH : TH * MT;             // rows of the tile grid one super tile covers
  constexpr int PATCH_BYTES = (TILE_ROWS + 2) * PITCH * ROW_BYTES;
  constexpr int TMEM_USED = 2 * MT * BN;
  constexpr int TMEM_COLS = TMEM_USED <= 32 ? 32 : (TMEM_USED <= 64 ? 64 : (TMEM_USED <= 128 ? 128 : (TMEM_USED <= 256 ? 256 : 512)));
  constexpr uint32_t idesc = umma_idesc_bf16(BM, BN);
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_w = smem;                                   // resident weights, NCH chunks
  uint8_t* smem_a = smem + ((W_BYTES + 1023) / 1024) * 1024;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_a + p.a_slots * p.a_slot_bytes);
  uint64_t* w_full = bars;                 // [1]
  uint64_t* full_a = w_full + 1;           // [MAX_A]
  uint64_t* empty_a = full_a + MAX_A;      // [MAX_A]
  uint64_t* tmem_full = empty_a + MAX_A;   // [2]
  uint64_t* tmem_empty = tmem_full + 2;    // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_a);
    tma_prefetch_desc(&tm_b);
    mbar_init(w_full, 1u);
    for (int i = 0; i < p.a_slots; ++i) { mbar_init(&full_a[i], 1u); mbar_init(&empty_a[i], 1u); }
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
      for (int q = 0; q < NCH; ++q) tma_load_2d(smem_w + q * B_BYTES, &tm_b, w_full, q * BK, 0);
      int sa = 0;
      uint32_t pa = 0;
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
        const Geo g = geo<TILE_ROWS>(p, tile);
        mbar_wait(&empty_a[sa], pa ^ 1u);
        mbar_arrive_expect_tx(&full_a[sa], PATCH_BYTES);
        tma_load_4d(smem_a + sa * p.a_slot_bytes, &tm_a, &full_a[sa], 0, g.w0 - 1, g.h0 - 1, g.n);
        if (++sa == p.a_slots) { sa = 0; pa ^= 1u; }
      }
    }
  } else if (warp == 1) {
    {   // whole warp: warp-uniform bookkeeping in uniform registers; an elected lane issues (see conv_halo.cu)
      const uint64_t a_hi = umma_desc(0u, PITCH * ROW_BYTES, CW == 64 ? 2u : (CW == 32 ? 4u : 6u));
      const uint64_t b_d0 = umma_desc(smem_u32(smem_w), 1024u, 2u);
      mbar_wait(w_full, 0);
      int sa = 0, buf = 0;
      uint32_t pa = 0, pbuf = 0;
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
        mbar_wait(&tmem_empty[buf], pbuf ^ 1u);
        mbar_wait(&full_a[sa], pa);
        tc_fence_after();
        const uint64_t a_d = a_hi + (smem_u32(smem_a + sa * p.a_slot_bytes) >> 4);
        // k-step outer, M tile inner: consecutive MMAs go to DIFFERENT accumulators, so the tensor pipe never
        // waits on the read-modify-write latency of one accumulator (matters for the N = 16 / 32 layers)
        if (elect_one()) {
        if (FOLD) {
          // effective tap outer, class inner: consecutive MMAs go to different accumulators
#pragma unroll
          for (int e = 0; e < 4; ++e) {
#pragma unroll
            for (int ch = 0; ch < CW; ch += 16) {
#pragma unroll
              for (int cls = 0; cls < 4; ++cls) {
                const int kk = (cls * 4 + e) * CW + ch;                    // K index of the folded packing (compile time)
                const int off = ((cls >> 1) + (e >> 1)) * PITCH + (cls & 1) + (e & 1);   // low-res pixel (a-1+ey, b-1+ex)
                umma_bf16_ss(tmem_base + (buf * MT + cls) * BN, a_d + ((off * ROW_BYTES + ch * 2) >> 4),
                             b_d0 + (((kk / BK) * B_BYTES + (kk % BK) * 2) >> 4), idesc, (e == 0 && ch == 0) ? 0u : 1u);
              }
            }
          }
        }
#pragma unroll
        for (int q = 0; q < (FOLD ? 0 : NCH); ++q) {
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            const int kk = q * BK + k * 16;  // K index = tap * CW + channel (compile time)
            if (kk < 9 * CW) {
              const int tap = kk / CW, ch = kk % CW;
              const uint64_t b_k = b_d0 + ((q * B_BYTES + k * 32) >> 4);
#pragma unroll
              for (int mt = 0; mt < MT; ++mt)
                umma_bf16_ss(tmem_base + (buf * MT + mt) * BN,
                             a_d + ((tap_pixel_offset<PAR>(mt, tap) * ROW_BYTES + ch * 2) >> 4), b_k, idesc,
                             kk == 0 ? 0u : 1u);
            }
          }
        }
        umma_commit(&empty_a[sa]);
        umma_commit(&tmem_full[buf]);
        }
        __syncwarp();
        if (++sa == p.a_slots) { sa = 0; pa ^= 1u; }
        if ((buf ^= 1) == 0) pbuf ^= 1u;
      }
    }
  } else {
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    int buf = 0;
    uint32_t pbuf = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
      const Geo g = geo<TILE_ROWS>(p, tile);
      auto pixel_off = [&](int mt) {
        int oy = g.h0 + (PAR ? 0 : mt * TH) + (row >> 3), ox = g.w0 + (row & 7);
        if (PAR) { oy = 2 * oy + (mt >> 1); ox = 2 * ox + (mt & 1); }
        return ((static_cast<int64_t>(g.n) * p.H + oy) * p.W + ox) * p.C_out;
      };
      uint4 res[BN / 8], res_next[BN / 8];
      if (p.has_residual) {   // first residual line is in flight while the MMAs of this super tile run
        const int64_t o0 = pixel_off(0);
#pragma unroll
        for (int j = 0; j < BN / 16; ++j) ldg_v8(p.residual + o0 + 16 * j, res[2 * j], res[2 * j + 1]);
      }
      mbar_wait(&tmem_full[buf], pbuf);
      tc_fence_after();
#pragma unroll 1
      for (int mt = 0; mt < MT; ++mt) {
        const int64_t out_off = pixel_off(mt);
        const uint32_t t_row = tmem_base + (buf * MT + mt) * BN + (static_cast<uint32_t>(quarter * 32) << 16);
        if (EPI == 1) {
          // segmentation head: logits = acc + bias; first-max argmax on the fp32 values; three optional outputs
          uint32_t v[16];
          tmem_ld_x16(t_row, v);
          tmem_ld_wait();
          const int64_t pix = out_off / p.C_out;                 // (n*H + y)*W + x
          int best = 0;
          float bv = 0.f;
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            if (k >= p.K) break;
            const float z = __uint_as_float(v[k]) + __ldg(p.shift + k);
            if (p.logits_nhwc) p.logits_nhwc[pix * p.K + k] = __float2bfloat16_rn(z);
            if (p.logits_nchw) {
              const int64_t hw = static_cast<int64_t>(p.H) * p.W;
              const int64_t n = pix / hw;
              p.logits_nchw[(n * p.K + k) * hw + (pix - n * hw)] = z;
            }
            if (k == 0 || z > bv) { bv = z; best = k; }
          }
          if (p.mask) p.mask[pix] = static_cast<uint8_t>(best);
          continue;
        }
        if (p.has_residual && mt + 1 < MT) {
          const int64_t o1 = pixel_off(mt + 1);
#pragma unroll
          for (int j = 0; j < BN / 16; ++j) ldg_v8(p.residual + o1 + 16 * j, res_next[2 * j], res_next[2 * j + 1]);
        }
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
          store_bf16x16(p.y + out_off + c0, f);
        }
        if (p.has_residual) {
#pragma unroll
          for (int j = 0; j < BN / 8; ++j) res[j] = res_next[j];
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

template <int BN, int CW, int MT, bool PAR, int EPI = 0, bool FOLD = false>
int launch_res(const CUtensorMap& tm_a, const CUtensorMap& tm_b, ResParams& p, cudaStream_t s) {
  constexpr int B_BYTES = BN * BK * 2;
  constexpr int NCH = FOLD ? (16 * CW) / BK : (9 * CW + BK - 1) / BK;
  constexpr int W_BYTES = ((NCH * B_BYTES + 1023) / 1024) * 1024;
  constexpr int TMEM_USED = 2 * MT * BN;
  constexpr int TMEM_COLS = TMEM_USED <= 32 ? 32 : (TMEM_USED <= 64 ? 64 : (TMEM_USED <= 128 ? 128 : (TMEM_USED <= 256 ? 256 : 512)));
  constexpr int TILE_ROWS = PAR ? TH : TH * MT;
  p.a_slot_bytes = (((TILE_ROWS + 2) * PITCH * CW * 2) + 1023) / 1024 * 1024;
  // co-resident CTAs (their issue threads and epilogues work in parallel): as many as TMEM and shared memory allow
  int ctas = 512 / TMEM_COLS < 2 ? 512 / TMEM_COLS : 2, a_slots = 0;   // 3 CTAs / SM measured slower (HBM-bound layers)
  for (; ctas >= 1; --ctas) {
    a_slots = (225 * 1024 / ctas - 2048 - W_BYTES - 1280) / p.a_slot_bytes;
    if (a_slots >= 2) break;
  }
  if (ctas < 1) return DT_ERR_UNSUPPORTED;
  p.a_slots = a_slots > MAX_A ? MAX_A : a_slots;
  const int smem = W_BYTES + p.a_slots * p.a_slot_bytes + 1024 + 256;
  static std::once_flag once;
  static cudaError_t attr_err = cudaSuccess;
  std::call_once(once, [] {
    attr_err = cudaFuncSetAttribute(conv_res_kernel<BN, CW, MT, PAR, EPI, FOLD>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    225 * 1024);
  });
  DT_CUDA(attr_err);
  const int slots = dt_num_sms() * ctas;
  const int grid = p.total_tiles < slots ? p.total_tiles : slots;
  conv_res_kernel<BN, CW, MT, PAR, EPI, FOLD><<<grid, kThreads, smem, s>>>(tm_a, tm_b, p);
  DT_LAUNCH_CHECK();
  return DT_OK;
}

}  // namespace

int dt_encode_bf16_map(CUtensorMap* tm, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                       const uint32_t* box, const uint32_t* elem_strides);

// Returns DT_ERR_UNSUPPORTED when the layer does not fit (caller falls back to conv_halo.cu / conv_tc.cu).
int dt_conv_res(const dt_conv_desc* d, const void* x, const void* w, int Kpad, const float* scale, const float* shift,
                const void* residual, void* y, cudaStream_t s) {
  const int parity = d->upsample ? 1 : 0;
  const int Hg = parity ? d->H / 2 : d->H, Wg = parity ? d->W / 2 : d->W;
  const int cw = d->C_in, bn = d->C_out;
  if (d->R != 3 || d->S != 3 || d->stride != 1 || d->pad != 1 || d->C_x != d->C_in ||
      !(cw == 64 || cw == 32 || cw == 16) || !(bn == 64 || bn == 32 || bn == 16) || Wg % TW != 0 || Hg % TH != 0)
    return DT_ERR_UNSUPPORTED;
  // M tiles per patch: the four parity classes, or as many row blocks as TMEM (2 buffers x MT x BN columns <= 512)
  // and the image height allow
  int mt = parity ? 4 : (bn == 64 ? 2 : 4);
  while (!parity && mt > 1 && Hg % (TH * mt) != 0) mt >>= 1;
  if (parity && bn > 32) return DT_ERR_UNSUPPORTED;
  const int tile_rows = parity ? TH : TH * mt;
  ResParams p;
  memset(&p, 0, sizeof(p));
  p.N = d->N; p.H = d->H; p.W = d->W; p.C_out = d->C_out;
  p.parity = parity; p.Hg = Hg; p.Wg = Wg;
  p.relu = d->relu; p.has_residual = d->has_residual;
  p.tiles_w = Wg / TW; p.tiles_h = Hg / tile_rows;
  p.total_tiles = p.tiles_w * p.tiles_h * d->N;
  p.residual = static_cast<const __nv_bfloat16*>(residual);
  p.y = static_cast<__nv_bfloat16*>(y);
  p.scale = scale; p.shift = shift;
  CUtensorMap tm_a, tm_b;
  {
    const uint64_t dims[2] = {static_cast<uint64_t>(Kpad), static_cast<uint64_t>(d->C_out)};
    const uint64_t strides[1] = {static_cast<uint64_t>(Kpad) * 2};
    const uint32_t box[2] = {BK, static_cast<uint32_t>(bn)};
    int rc = dt_encode_bf16_map(&tm_b, w, 2, dims, strides, box, nullptr);
    if (rc != DT_OK) return rc;
  }
  {
    const uint64_t dims[4] = {static_cast<uint64_t>(cw), static_cast<uint64_t>(Wg), static_cast<uint64_t>(Hg),
                              static_cast<uint64_t>(d->N)};
    const uint64_t strides[3] = {static_cast<uint64_t>(cw) * 2, static_cast<uint64_t>(Wg) * cw * 2,
                                 static_cast<uint64_t>(Hg) * Wg * cw * 2};
    const uint32_t box[4] = {static_cast<uint32_t>(cw), PITCH, static_cast<uint32_t>(tile_rows + 2), 1};
    int rc = dt_encode_bf16_map(&tm_a, x, 4, dims, strides, box, nullptr);
    if (rc != DT_OK) return rc;
  }
  if (d->flags & DT_CONV_UPS_FOLDED) {
    // weights in the folded packing (mode 5): only the parity instances below take them
    if (!parity || Kpad != 16 * cw) return DT_ERR_UNSUPPORTED;
    if (bn == 16 && cw == 32) return launch_res<16, 32, 4, true, 0, true>(tm_a, tm_b, p, s);
    if (bn == 32 && cw == 32) return launch_res<32, 32, 4, true, 0, true>(tm_a, tm_b, p, s);
    if (bn == 16 && cw == 16) return launch_res<16, 16, 4, true, 0, true>(tm_a, tm_b, p, s);
    if (bn == 32 && cw == 64) return launch_res<32, 64, 4, true, 0, true>(tm_a, tm_b, p, s);
    return DT_ERR_UNSUPPORTED;
  }
#define DT_RES(BNV, CWV, MTV, PARV) \
  if (bn == BNV && cw == CWV && mt == MTV && parity == (PARV ? 1 : 0)) \
    return launch_res<BNV, CWV, MTV, PARV>(tm_a, tm_b, p, s);
#define DT_RES_MT(BNV, CWV) DT_RES(BNV, CWV, 1, false) DT_RES(BNV, CWV, 2, false) DT_RES(BNV, CWV, 4, false)
  DT_RES(64, 64, 1, false) DT_RES(64, 64, 2, false)
  DT_RES_MT(32, 64) DT_RES_MT(32, 32) DT_RES_MT(16, 32) DT_RES_MT(16, 16) DT_RES_MT(32, 16) DT_RES_MT(16, 64)
  DT_RES(16, 32, 4, true) DT_RES(32, 32, 4, true) DT_RES(16, 16, 4, true) DT_RES(16, 64, 4, true) DT_RES(32, 64, 4, true)
#undef DT_RES_MT
#undef DT_RES
  return DT_ERR_UNSUPPORTED;
}


// Segmentation head on the tensor cores: 3x3 conv 16 -> K (K <= 4, weights padded to 16 output channels, bf16
// [16][192]) + bias, with the argmax / layout outputs of dt_head_fwd fused into the epilogue.
int dt_head_res(const void* x, int N, int H, int W, int K, const void* w_packed, const float* bias16,
                float* logits_nchw, void* logits_nhwc, uint8_t* mask, cudaStream_t s) {
  if (W % TW != 0 || H % TH != 0 || K < 1 || K > 4) return DT_ERR_UNSUPPORTED;
  int mt = 4;
  while (mt > 1 && H % (TH * mt) != 0) mt >>= 1;
  ResParams p;
  memset(&p, 0, sizeof(p));
  p.N = N; p.H = H; p.W = W; p.C_out = 16;
  p.Hg = H; p.Wg = W;
  p.tiles_w = W / TW; p.tiles_h = H / (TH * mt);
  p.total_tiles = p.tiles_w * p.tiles_h * N;
  p.shift = bias16;
  p.K = K;
  p.logits_nchw = logits_nchw;
  p.logits_nhwc = static_cast<__nv_bfloat16*>(logits_nhwc);
  p.mask = mask;
  CUtensorMap tm_a, tm_b;
  {
    const uint64_t dims[2] = {192, 16};
    const uint64_t strides[1] = {192 * 2};
    const uint32_t box[2] = {BK, 16};
    int rc = dt_encode_bf16_map(&tm_b, w_packed, 2, dims, strides, box, nullptr);
    if (rc != DT_OK) return rc;
  }
  {
    const uint64_t dims[4] = {16, static_cast<uint64_t>(W), static_cast<uint64_t>(H), static_cast<uint64_t>(N)};
    const uint64_t strides[3] = {32, static_cast<uint64_t>(W) * 32, static_cast<uint64_t>(H) * W * 32};
    const uint32_t box[4] = {16, PITCH, static_cast<uint32_t>(TH * mt + 2), 1};
    int rc = dt_encode_bf16_map(&tm_a, x, 4, dims, strides, box, nullptr);
    if (rc != DT_OK) return rc;
  }
  if (mt == 4) return launch_res<16, 16, 4, false, 1>(tm_a, tm_b, p, s);
  if (mt == 2) return launch_res<16, 16, 2, false, 1>(tm_a, tm_b, p, s);
  return launch_res<16, 16, 1, false, 1>(tm_a, tm_b, p, s);
}
