// 3x3 / stride-1 / pad-1 convolution of the wide layers (C_in % 64 == 0, C_out % 128 == 0) on CTA PAIRS:
// tcgen05.mma.cta_group::2, M = 256 output pixels (two adjacent 8 x 16 tiles, one per CTA of a 2-CTA cluster), N = 128
// or 256 output channels.
//
// Why: in SS mode the operand fetch of one CTA runs at ~64 B/clk/SM (measured per-MMA times).  A single-CTA
// MMA of M 128 x N 256 x K 16 reads 4 KB of A + 8 KB of B = 192 clk of fetch for 128 clk of tensor work.  In a pair each
// CTA fetches its own A (its 128 pixels) and only HALF of B (N/2 weight rows): 8 KB per 128 tensor clocks.
//
// Structure (same in both CTAs unless noted; shared-memory layout identical, so one descriptor addresses both):
//   warp 0 lane 0  TMA producer: own halo patch per 64-channel slab (A), own half of every 64-wide weight chunk (B);
//                  all loads signal the LEADER's (rank 0) full barriers (cp.async.bulk.tensor .cta_group::2, barrier
//                  address with the peer bit cleared); the leader arms them with the bytes of both CTAs
//   warp 1 lane 0  leader only: MMA issue; tcgen05.commit .cta_group::2 multicast releases the ring slots / signals the
//                  accumulator in BOTH CTAs.  Warp 1 of both CTAs allocates / frees TMEM (cta_group::2)
//   warps 2..5     epilogue of the CTA's own 128 accumulator rows (as conv_halo.cu); arrive on the leader's tmem_empty
// K order per output element (slab-major, taps inside) is that of conv_halo.cu: identical results.
#include <cstring>
#include <mutex>

#include "common.cuh"

namespace {

constexpr int BM = 128, BK = 64;
constexpr int TW = 8, TH = 16, PITCH = TW + 2, PATCH_PIX = (TH + 2) * PITCH;
constexpr int kThreads = 192;
constexpr int A_SLOTS = 3;
constexpr int PATCH_BYTES = PATCH_PIX * BK * 2;                         // 23040
// 8 x 8 images (resnet layer4 at 256 x 256 tiles): an M tile is TWO whole images, stored row-interleaved with zero halos,
// [10 rows][2 images][10 pixels] - the 16 groups of 8 pixels (row y of image i = group 2y + i) keep the uniform 10-pixel
// stride the UMMA descriptor needs, a vertical tap moves 2 * PITCH pixels.  One TMA box over the map (C, x, n, y).
constexpr int PATCH8_BYTES = 10 * 2 * PITCH * BK * 2;                   // 25600
constexpr int A_SLOT_BYTES = (PATCH8_BYTES + 1023) / 1024 * 1024;       // 25600
constexpr int MAX_B = 8;
constexpr uint32_t kPeerMask = 0xFEFFFFFFu;   // clears the CTA-rank bit of a shared::cluster address: the even CTA's copy

struct PairParams {
  int N, H, W, C_in, C_out;
  int relu, has_residual;
  int pairs_w, tiles_h, n_tiles, total_tiles, n_slabs, b_slots;
  int img8;          // 1: 8 x 8 images, two per CTA tile (four per pair)
  const __nv_bfloat16* residual;
  __nv_bfloat16* y;
  const float* scale;
  const float* shift;
};

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tma_load_4d_2sm(void* dst, const void* tmap, uint64_t* bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      :: "r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(smem_u32(bar) & kPeerMask), "r"(c0), "r"(c1),
         "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void tma_load_2d_2sm(void* dst, const void* tmap, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      :: "r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(smem_u32(bar) & kPeerMask), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void umma_bf16_ss_2sm(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                                 uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      :: "r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// arrives (once all previously issued MMAs are complete) on the barrier at this offset in BOTH CTAs of the pair
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar) {
  const uint16_t mask = 3;
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               :: "r"(smem_u32(bar)), "h"(mask) : "memory");
}
__device__ __forceinline__ void mbar_arrive_leader(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" :: "r"(smem_u32(bar) & kPeerMask) : "memory");
}
__device__ __forceinline__ void tmem_alloc_2sm(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;"
               :: "r"(smem_u32(dst_smem)), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem_relinquish_2sm() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_2sm(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" :: "r"(taddr), "r"(ncols) : "memory");
}

struct PGeo {
  int n_tile, n, h0, w0;
};
__device__ __forceinline__ PGeo pgeo(const PairParams& p, int tile, int rank) {
  PGeo g;
  g.n_tile = tile % p.n_tiles;
  int m = tile / p.n_tiles;
  if (p.img8) {      // m = group of four images; this CTA takes two of them
    g.w0 = 0; g.h0 = 0; g.n = m * 4 + rank * 2;
    return g;
  }
  g.w0 = (m % p.pairs_w) * (2 * TW) + rank * TW; m /= p.pairs_w;
  g.h0 = (m % p.tiles_h) * TH;
  g.n = m / p.tiles_h;
  return g;
}

template <int BN>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, BN <= 128 ? 2 : 1)
conv_pair_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b, const PairParams p) {
  constexpr int B_HALF = (BN / 2) * BK * 2;                 // this CTA's half of a weight chunk
  constexpr int TMEM_COLS = 2 * BN;                         // two accumulator buffers
  constexpr uint32_t idesc = umma_idesc_bf16(2 * BM, BN);   // M = 256 across the pair
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + A_SLOTS * A_SLOT_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_b + p.b_slots * B_HALF);
  uint64_t* full_a = bars;                 // [A_SLOTS]   (used in the leader)
  uint64_t* empty_a = full_a + A_SLOTS;    // [A_SLOTS]
  uint64_t* full_b = empty_a + A_SLOTS;    // [MAX_B]     (used in the leader)
  uint64_t* empty_b = full_b + MAX_B;      // [MAX_B]
  uint64_t* tmem_full = empty_b + MAX_B;   // [2]
  uint64_t* tmem_empty = tmem_full + 2;    // [2]         (used in the leader: 256 arrivals)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int rank = static_cast<int>(cluster_ctarank());
  const bool leader = rank == 0;
  const int cluster_id = blockIdx.x >> 1, n_clusters = gridDim.x >> 1;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_a);
    tma_prefetch_desc(&tm_b);
    for (int i = 0; i < A_SLOTS; ++i) { mbar_init(&full_a[i], 1u); mbar_init(&empty_a[i], 1u); }
    for (int i = 0; i < p.b_slots; ++i) { mbar_init(&full_b[i], 1u); mbar_init(&empty_b[i], 1u); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tmem_full[i], 1u); mbar_init(&tmem_empty[i], 256u); }
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc_2sm(tmem_slot, TMEM_COLS);
    tmem_relinquish_2sm();
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();            // both CTAs' barriers are initialised before any remote arrive / TMA signal
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);   // warp-uniform for the compiler (uniform registers)

  if (warp == 0) {
    if (lane == 0) {
      int sa = 0, sb = 0;
      uint32_t pa = 0, pb = 0;
      auto load_patch = [&](int tile, int slab) {
        const PGeo g = pgeo(p, tile, rank);
        mbar_wait(&empty_a[sa], pa ^ 1u);
        if (leader) mbar_arrive_expect_tx(&full_a[sa], 2 * (p.img8 ? PATCH8_BYTES : PATCH_BYTES));
        if (p.img8) tma_load_4d_2sm(smem_a + sa * A_SLOT_BYTES, &tm_a, &full_a[sa], slab * BK, -1, g.n, -1);   // (c, x, n, y)
        else tma_load_4d_2sm(smem_a + sa * A_SLOT_BYTES, &tm_a, &full_a[sa], slab * BK, g.w0 - 1, g.h0 - 1, g.n);
        if (++sa == A_SLOTS) { sa = 0; pa ^= 1u; }
      };
      bool primed = false;
      for (int tile = cluster_id; tile < p.total_tiles; tile += n_clusters) {
        const PGeo g = pgeo(p, tile, rank);
        for (int slab = 0; slab < p.n_slabs; ++slab) {
          if (!primed) { load_patch(tile, slab); primed = true; }
          for (int tap = 0; tap < 9; ++tap) {
            if (tap == 0) {          // request the next patch before this patch's weight chunks
              int ns = slab + 1, nt = tile;
              if (ns == p.n_slabs) { ns = 0; nt = tile + n_clusters; }
              if (nt < p.total_tiles) load_patch(nt, ns);
            }
            mbar_wait(&empty_b[sb], pb ^ 1u);
            if (leader) mbar_arrive_expect_tx(&full_b[sb], 2 * B_HALF);
            tma_load_2d_2sm(smem_b + sb * B_HALF, &tm_b, &full_b[sb], tap * p.C_in + slab * BK,
                            g.n_tile * BN + rank * (BN / 2));
            if (++sb == p.b_slots) { sb = 0; pb ^= 1u; }
          }
        }
      }
    }
  } else if (warp == 1) {
    if (leader) {   // whole warp of the leader CTA: uniform bookkeeping; an elected lane issues (see conv_halo.cu)
      const uint64_t a_hi = umma_desc(0u, PITCH * BK * 2, 2u);
      const uint64_t b_hi = umma_desc(0u, 1024u, 2u);
      int sa = 0, sb = 0, acc = 0;
      uint32_t pa = 0, pb = 0, pacc = 0;
      const int rowp = p.img8 ? 2 * PITCH : PITCH;      // pixels between vertically adjacent rows of the patch
      for (int tile = cluster_id; tile < p.total_tiles; tile += n_clusters) {
        mbar_wait(&tmem_empty[acc], pacc ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        uint32_t accum = 0;
        for (int slab = 0; slab < p.n_slabs; ++slab) {
          mbar_wait(&full_a[sa], pa);
          tc_fence_after();
          const uint64_t a_d = a_hi + (smem_u32(smem_a + sa * A_SLOT_BYTES) >> 4);
          for (int tap = 0; tap < 9; ++tap) {
            mbar_wait(&full_b[sb], pb);
            tc_fence_after();
            const uint64_t b_d = b_hi + (smem_u32(smem_b + sb * B_HALF) >> 4);
            const uint64_t a_t = a_d + ((static_cast<uint32_t>((tap / 3) * rowp + tap % 3) * (BK * 2)) >> 4);
            if (elect_one()) {
#pragma unroll
              for (int k = 0; k < BK / 16; ++k) umma_bf16_ss_2sm(d_tmem, a_t + 2 * k, b_d + 2 * k, idesc, (accum | k) ? 1u : 0u);
              umma_commit_pair(&empty_b[sb]);
              if (tap == 8) {
                umma_commit_pair(&empty_a[sa]);
                if (slab + 1 == p.n_slabs) umma_commit_pair(&tmem_full[acc]);
              }
            }
            __syncwarp();
            accum = 1;
            if (++sb == p.b_slots) { sb = 0; pb ^= 1u; }
          }
          if (++sa == A_SLOTS) { sa = 0; pa ^= 1u; }
        }
        if ((acc ^= 1) == 0) pacc ^= 1u;
      }
    }
  } else {
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    constexpr int SC = 64;
    int acc = 0;
    uint32_t pacc = 0;
    for (int tile = cluster_id; tile < p.total_tiles; tile += n_clusters) {
      const PGeo g = pgeo(p, tile, rank);
      // 8 x 8 images: group (row >> 3) = 2 * y + image
      const int n_img = p.img8 ? g.n + ((row >> 3) & 1) : g.n;
      const bool live = n_img < p.N;                      // a last group of fewer than four images
      const int oy = p.img8 ? (row >> 4) : g.h0 + (row >> 3), ox = g.w0 + (row & 7);
      const int64_t out_off = ((static_cast<int64_t>(live ? n_img : 0) * p.H + oy) * p.W + ox) * p.C_out + g.n_tile * BN;
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
          if (live) store_bf16x16(p.y + out_off + s0 + c0, f);
        }
        if (more) {
#pragma unroll
          for (int j = 0; j < SC / 8; ++j) res[j] = res_next[j];
        }
      }
      tc_fence_before();
      mbar_arrive_leader(&tmem_empty[acc]);     // 128 threads of each CTA: 256 arrivals on the leader's barrier
      if ((acc ^= 1) == 0) pacc ^= 1u;
    }
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();            // the peer's accumulators are drained before the pair's TMEM is freed
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_2sm(tmem_base, TMEM_COLS);
  }
}

template <int BN>
int launch_pair(const CUtensorMap& tm_a, const CUtensorMap& tm_b, PairParams& p, cudaStream_t s) {
  constexpr int B_HALF = (BN / 2) * BK * 2;
  // N = 128: 256 TMEM columns and ~110 KB of shared memory per CTA, so two CTAs (of different pairs) share an SM and fill
  // each other's pipeline bubbles; N = 256 needs all 512 columns: one CTA per SM
  constexpr int CTAS = BN <= 128 ? 2 : 1;
  int b_slots = ((CTAS == 2 ? 111 : 216) * 1024 - A_SLOTS * A_SLOT_BYTES - 2048) / B_HALF;
  if (b_slots > MAX_B) b_slots = MAX_B;
  if (b_slots < 3) return DT_ERR_UNSUPPORTED;
  p.b_slots = b_slots;
  const int smem = A_SLOTS * A_SLOT_BYTES + b_slots * B_HALF + 1024 + 512;
  static std::once_flag once;
  static cudaError_t attr_err = cudaSuccess;
  std::call_once(once, [] {
    attr_err = cudaFuncSetAttribute(conv_pair_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, 225 * 1024);
  });
  DT_CUDA(attr_err);
  int clusters = dt_num_sms() / 2 * CTAS;
  if (clusters > p.total_tiles) clusters = p.total_tiles;
  conv_pair_kernel<BN><<<2 * clusters, kThreads, smem, s>>>(tm_a, tm_b, p);
  DT_LAUNCH_CHECK();
  return DT_OK;
}

}  // namespace

int dt_encode_bf16_map(CUtensorMap* tm, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                       const uint32_t* box, const uint32_t* elem_strides);

// Returns DT_ERR_UNSUPPORTED when the layer does not fit the pair scheme (caller falls back to conv_halo.cu).
int dt_conv_pair(const dt_conv_desc* d, const void* x, const void* w, int Kpad, const float* scale, const float* shift,
                 const void* residual, void* y, cudaStream_t s) {
  const bool img8 = d->H == 8 && d->W == 8;
  if (d->R != 3 || d->S != 3 || d->stride != 1 || d->pad != 1 || d->upsample || d->C_x != d->C_in || d->C_in % BK != 0 ||
      d->C_out % 128 != 0 || (!img8 && (d->H % TH != 0 || d->W % (2 * TW) != 0)) || Kpad != 9 * d->C_in)
    return DT_ERR_UNSUPPORTED;
  const int BN = d->C_out % 256 == 0 ? 256 : 128;
  PairParams p;
  memset(&p, 0, sizeof(p));
  p.N = d->N; p.H = d->H; p.W = d->W; p.C_in = d->C_in; p.C_out = d->C_out;
  p.relu = d->relu; p.has_residual = d->has_residual;
  p.pairs_w = d->W / (2 * TW); p.tiles_h = d->H / TH; p.n_tiles = d->C_out / BN;
  p.total_tiles = p.pairs_w * p.tiles_h * d->N * p.n_tiles;
  p.img8 = img8 ? 1 : 0;
  if (img8) p.total_tiles = ((d->N + 3) / 4) * p.n_tiles;
  p.n_slabs = d->C_in / BK;
  p.residual = static_cast<const __nv_bfloat16*>(residual);
  p.y = static_cast<__nv_bfloat16*>(y);
  p.scale = scale; p.shift = shift;
  CUtensorMap tm_a, tm_b;
  {
    const uint64_t dims[2] = {static_cast<uint64_t>(Kpad), static_cast<uint64_t>(d->C_out)};
    const uint64_t strides[1] = {static_cast<uint64_t>(Kpad) * 2};
    const uint32_t box[2] = {BK, static_cast<uint32_t>(BN / 2)};
    int rc = dt_encode_bf16_map(&tm_b, w, 2, dims, strides, box, nullptr);
    if (rc != DT_OK) return rc;
  }
  {
    const uint64_t dims[4] = {static_cast<uint64_t>(d->C_in), static_cast<uint64_t>(d->W), static_cast<uint64_t>(d->H),
                              static_cast<uint64_t>(d->N)};
    const uint64_t strides[3] = {static_cast<uint64_t>(d->C_in) * 2, static_cast<uint64_t>(d->W) * d->C_in * 2,
                                 static_cast<uint64_t>(d->H) * d->W * d->C_in * 2};
    const uint32_t box[4] = {BK, PITCH, TH + 2, 1};
    int rc;
    if (img8) {      // dimension order (c, x, n, y): one box = rows -1..8 of two images, row-interleaved in shared memory
      const uint64_t dims8[4] = {dims[0], dims[1], dims[3], dims[2]};
      const uint64_t strides8[3] = {strides[0], strides[2], strides[1]};
      const uint32_t box8[4] = {BK, PITCH, 2, 10};
      rc = dt_encode_bf16_map(&tm_a, x, 4, dims8, strides8, box8, nullptr);
    } else {
      rc = dt_encode_bf16_map(&tm_a, x, 4, dims, strides, box, nullptr);
    }
    if (rc != DT_OK) return rc;
  }
  return BN == 256 ? launch_pair<256>(tm_a, tm_b, p, s) : launch_pair<128>(tm_a, tm_b, p, s);
}
