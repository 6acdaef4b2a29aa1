// Fused decoder tail: decoder.blocks.4.conv2 (3x3, 16 -> 16, BatchNorm + ReLU) -> segmentation head (3x3, 16 -> K, + bias) ->
// logits / first-max argmax, ONE kernel (SURVEY.md 8b `dt_tail_fused`, layers Appendix A rows 45-46).
//
// Unfused, conv2 writes 16 channels x 65 536 pixels per tile (0.85 GB per 405-tile batch) and the head reads them back.  Here
// the intermediate never leaves the SM.  Both layers are row-streaming convolutions (conv_row.cu: an M tile = 128 pixels of
// one image row, the three vertical taps in the N dimension of one tcgen05.mma, accumulators of consecutive output rows in a
// ring of TMEM column slots):
//
//   stage 1  conv2: input rows arrive by TMA (one dense 130-pixel box per tile and row); a finished output row goes
//            TMEM -> registers -> scale / shift / ReLU -> bf16 -> SHARED memory: a ring of full-width rows (258 pixels with the
//            two zero halo pixels) written in the SWIZZLE_32B pattern the tensor core reads (16-byte chunk ^= address bit 7),
//            fence.proxy.async, mbarrier;
//   stage 2  head: its A operand is that ring (tile t of a row starts 128 * t pixels into it), its epilogue writes the logits
//            (fp32 NCHW / bf16 NHWC) and the mask.
// A work item is (image, 32 head rows): stage 1 computes the 34 rows the head needs (two halo rows recomputed per item, 6 %).
//
// One warp issues the MMAs of BOTH stages, interleaved tile by tile, so that four accumulators rotate (a chain into one
// accumulator costs 138 clk per MMA, four rotating 67: profiles/r02_mma_rate.txt); stage 2 walks the same item sequence LAG
// conv2 rows behind stage 1, across item boundaries, so it never waits for the epilogue it feeds.  Each stage has its own
// epilogue warpgroup (TMEM lanes are reachable from any warp with the same warp % 4).  The conv2 rows are rounded to bf16
// exactly where the unfused path stores them and every accumulation keeps its (r, s, ci) order: the logits are bit-identical
// to dt_conv2d_fwd + dt_head_fwd_tc (tests/test_gpu_conv_row.py).
//
//   warp 0: TMA producer   warp 1: TMEM alloc + MMA issue (whole warp, elected lane)   warps 2..5: conv2 epilogue
//   warps 6..9: head epilogue.  256 TMEM columns and ~108 KB of shared memory per CTA: two CTAs per SM.
//   Width 128 or 256 (full rows per CTA).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>

#include "common.cuh"

namespace {

constexpr int kThreads = 320;
constexpr int C = 16;                       // channels in, mid, out (head padded to 16)
constexpr int ROW_BYTES = C * 2;            // one pixel
constexpr int SL = 4;                       // accumulator ring slots per stage and tile
constexpr int LAG = 3;                      // stage 2 starts a conv2 row once LAG rows from it on are committed
constexpr int XS = 4;                       // input-row ring slots
constexpr int S2 = 6;                       // intermediate-row ring slots (>= LAG + 3)
constexpr int X_TILE = 5120;                // 130 pixels x 32 B, padded
constexpr int W_S = 2048;                   // weights of one horizontal tap: [3 * 16 rows][16] = 1536 B, padded
constexpr int MID_ROW = 9216;               // 258 pixels x 32 B = 8256, padded to a multiple of 1024
constexpr uint32_t LAYOUT = 6u;             // SWIZZLE_32B

struct TailParams {
  int N, H, W, K;
  int mt;                 // 128-pixel tiles per row (1 or 2)
  int R, chunks, total_items;
  const float* scale2;
  const float* shift2;
  const float* bias;      // 16 floats
  float* logits_nchw;
  __nv_bfloat16* logits_nhwc;
  uint8_t* mask;
  unsigned long long* dbg;   // DT_ROW_DEBUG: wait clocks per role of CTA 0
};

// item -> image, first head row, head rows, first conv2 row, conv2 rows (the rows the head reads, clipped to the image)
struct ItemGeo { int n, y0, n3, a2, n2; };
__device__ __forceinline__ ItemGeo decode(const TailParams& p, int item) {
  ItemGeo g;
  g.n = item / p.chunks;
  g.y0 = (item % p.chunks) * p.R;
  g.n3 = p.H - g.y0 < p.R ? p.H - g.y0 : p.R;
  g.a2 = g.y0 - 1 < 0 ? 0 : g.y0 - 1;
  const int b2 = g.y0 + g.n3 + 1 > p.H ? p.H : g.y0 + g.n3 + 1;
  g.n2 = b2 - g.a2;
  return g;
}

// One stage's walk over (item, input row) in the MMA warp.  Input index i <-> input row base - 1 + i; it feeds the output rows
// i - r (r = vertical tap) that lie inside the item.
struct Cursor {
  int item;               // >= total_items: exhausted
  int base, rows;         // first output row, output rows of the item
  int first_i, last_i;    // valid input indices (rows outside the image are zero padding: skipped)
  int i;
  uint32_t c0;            // ring position of the item's first output row = running count of output rows
  int opened, committed;  // accumulators opened / committed within the item
};

__device__ __forceinline__ void cursor_load(Cursor& c, const TailParams& p, int stage) {
  if (c.item >= p.total_items) return;
  const ItemGeo g = decode(p, c.item);
  c.base = stage == 1 ? g.a2 : g.y0;
  c.rows = stage == 1 ? g.n2 : g.n3;
  c.first_i = c.base == 0 ? 1 : 0;
  c.last_i = c.base + c.rows < p.H ? c.rows + 1 : c.rows;
  c.i = c.first_i;
  c.opened = c.committed = 0;
}

__device__ __forceinline__ void cursor_advance(Cursor& c, const TailParams& p, int stage) {
  if (++c.i > c.last_i) {
    c.c0 += static_cast<uint32_t>(c.rows);
    c.item += gridDim.x;
    cursor_load(c, p, stage);
  }
}

// What one input row of a stage issues: MMAs with N = taps * 16 into consecutive column slots, cut where the ring wraps
struct StepDesc {
  uint32_t col_a, id_a, id_b;
  uint32_t a_lo, b_lo;      // low words of the A / B descriptors (start address >> 4 | LBO); the high word is a constant
  int len_a, len_b, done;
};

// high word of the shared-memory descriptors used here (dense 8-pixel groups of 32-byte rows, SWIZZLE_32B; see umma_desc)
constexpr uint32_t DESC_HI = ((8u * ROW_BYTES) >> 4) | (1u << 14) | (LAYOUT << 29);
constexpr uint32_t DESC_LO = 1u << 16;
__device__ __forceinline__ void mma16(uint32_t col, uint32_t a_lo, uint32_t b_lo, uint32_t idesc) {
  umma_bf16_ss(col, (static_cast<uint64_t>(DESC_HI) << 32) | a_lo, (static_cast<uint64_t>(DESC_HI) << 32) | b_lo, idesc, 1u);
}

// The MMAs of one step: stage 1 and stage 2 interleaved tile by tile (four accumulators rotate)
template <int MT, bool S1, bool S2_>
__device__ __forceinline__ void issue_mmas(const StepDesc& d1, const StepDesc& d2, uint32_t t1, uint32_t t2) {
#pragma unroll
  for (int s = 0; s < 3; ++s)
#pragma unroll
    for (int t = 0; t < MT; ++t) {
      if (S1) {
        const uint32_t a = d1.a_lo + ((t * X_TILE + s * ROW_BYTES) >> 4), b = d1.b_lo + ((s * W_S) >> 4);
        mma16(d1.col_a + t * (SL * C), a, b, d1.id_a);
        if (d1.len_b > 0) mma16(t1 + t * (SL * C), a, b + d1.len_a * ((C * ROW_BYTES) >> 4), d1.id_b);
      }
      if (S2_) {
        const uint32_t a = d2.a_lo + ((t * 128 * ROW_BYTES + s * ROW_BYTES) >> 4), b = d2.b_lo + ((s * W_S) >> 4);
        mma16(d2.col_a + t * (SL * C), a, b, d2.id_a);
        if (d2.len_b > 0) mma16(t2 + t * (SL * C), a, b + d2.len_a * ((C * ROW_BYTES) >> 4), d2.id_b);
      }
    }
}

__device__ __forceinline__ StepDesc step_prepare(Cursor& c, uint32_t a_addr, uint32_t w_base, uint32_t t_col0,
                                                 uint64_t* acc_empty, bool dbg, long long& dbg_wait) {
  constexpr uint32_t idesc0 = umma_idesc_bf16(128, 0);
  const int i = c.i, rows = c.rows;
  const int hi = i < rows ? i : rows - 1;
  const long long t0 = dbg ? clock64() : 0;
  for (; c.opened <= hi; ++c.opened) {       // accumulators this row touches first: wait until the epilogue has cleared them
    const uint32_t k = c.c0 + c.opened;
    mbar_wait(&acc_empty[k % SL], ((k / SL) & 1u) ^ 1u);
  }
  if (dbg) dbg_wait += clock64() - t0;
  StepDesc d;
  const int r_lo = i - (rows - 1) > 0 ? i - (rows - 1) : 0, r_hi = i < 2 ? i : 2;
  const uint32_t pos_lo = SL - 1 - ((c.c0 + i - r_lo) % SL);
  const int n_taps = r_hi - r_lo + 1;
  d.len_a = n_taps < static_cast<int>(SL - pos_lo) ? n_taps : static_cast<int>(SL - pos_lo);
  d.len_b = n_taps - d.len_a;                    // taps past the wrap point start at column slot 0
  d.col_a = t_col0 + pos_lo * C;
  d.id_a = idesc0 | (static_cast<uint32_t>((d.len_a * C) >> 3) << 17);
  d.id_b = idesc0 | (static_cast<uint32_t>((d.len_b * C) >> 3) << 17);
  d.a_lo = DESC_LO | ((a_addr & 0x3FFFFu) >> 4);
  d.b_lo = DESC_LO | (((w_base + r_lo * C * ROW_BYTES) & 0x3FFFFu) >> 4);
  d.done = i == c.last_i ? rows - 1 : i - 2;     // output rows that have all their contributions after this step
  return d;
}

__device__ __forceinline__ uint32_t sw32(uint32_t addr) { return addr ^ (((addr >> 7) & 1u) << 4); }

__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" :: "r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

template <int MT>
__global__ void __launch_bounds__(kThreads, 2)
tail_fused_kernel(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_w2,
                  const __grid_constant__ CUtensorMap tm_wh, const TailParams p) {
  constexpr int X_SLOT = MT * X_TILE;
  constexpr int TMEM_COLS = 2 * MT * SL * C;             // 256 (MT = 2) / 128
  constexpr int T1 = 0, T2 = MT * SL * C;                // TMEM column base of stage 1 / stage 2
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_w2 = smem;                               // 3 x W_S
  uint8_t* smem_wh = smem_w2 + 3 * W_S;                  // 3 x W_S
  uint8_t* smem_x = smem_wh + 3 * W_S;                   // XS x X_SLOT
  uint8_t* smem_mid = smem_x + XS * X_SLOT;              // S2 x MID_ROW
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_mid + S2 * MID_ROW);
  uint64_t* w_full = bars;                 // [1]
  uint64_t* full_x = w_full + 1;           // [XS]
  uint64_t* empty_x = full_x + XS;         // [XS]
  uint64_t* acc1_full = empty_x + XS;      // [SL]
  uint64_t* acc1_empty = acc1_full + SL;   // [SL]
  uint64_t* acc2_full = acc1_empty + SL;   // [SL]
  uint64_t* acc2_empty = acc2_full + SL;   // [SL]
  uint64_t* mid_full = acc2_empty + SL;    // [S2]  128 arrivals (the epilogue threads that wrote the row)
  uint64_t* mid_empty = mid_full + S2;     // [S2]  the stage-2 MMAs that read the row are done
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(mid_empty + S2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool dbg = p.dbg != nullptr;
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_x);
    tma_prefetch_desc(&tm_w2);
    tma_prefetch_desc(&tm_wh);
    mbar_init(w_full, 1u);
    for (int i = 0; i < XS; ++i) { mbar_init(&full_x[i], 1u); mbar_init(&empty_x[i], 1u); }
    for (int i = 0; i < SL; ++i) {
      mbar_init(&acc1_full[i], 1u); mbar_init(&acc1_empty[i], 128u);
      mbar_init(&acc2_full[i], 1u); mbar_init(&acc2_empty[i], 128u);
    }
    for (int i = 0; i < S2; ++i) { mbar_init(&mid_full[i], 128u); mbar_init(&mid_empty[i], 1u); }
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, TMEM_COLS);
    tmem_relinquish();
  }
  // the intermediate rows' halo pixels (and everything else) start at zero; only pixels 0 .. W-1 are ever written
  for (int i = threadIdx.x; i < S2 * MID_ROW / 16; i += kThreads) st_shared_v4(smem_u32(smem_mid) + i * 16, 0u, 0u, 0u, 0u);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);
  if (warp >= 2 && warp < 6) {   // all accumulators start at zero: every MMA accumulates
    const uint32_t t_lane = tmem_base + (static_cast<uint32_t>((warp & 3) * 32) << 16);
#pragma unroll 1
    for (int c = 0; c < TMEM_COLS; c += 16) tmem_st_zero_x16(t_lane + c);
    tmem_st_wait();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();

  if (warp == 0) {
    // ---- producer: the conv2 input rows of every item, one 130-pixel box per tile ----
    if (lane == 0) {
      mbar_arrive_expect_tx(w_full, 2u * 3u * 3u * C * ROW_BYTES);
      for (int s = 0; s < 3; ++s) {
        tma_load_4d(smem_w2 + s * W_S, &tm_w2, w_full, 0, 0, 0, s);
        tma_load_4d(smem_wh + s * W_S, &tm_wh, w_full, 0, 0, 0, s);
      }
    }
    int sx = 0;
    uint32_t px = 0;
    for (int item = blockIdx.x; item < p.total_items; item += gridDim.x) {
      const ItemGeo g = decode(p, item);
      for (int i = 0; i <= g.n2 + 1; ++i) {
        const int q = g.a2 - 1 + i;
        if (q < 0 || q >= p.H) continue;
        if (lane == 0) {
          mbar_wait(&empty_x[sx], px ^ 1u);
          mbar_arrive_expect_tx(&full_x[sx], static_cast<uint32_t>(MT) * 130u * ROW_BYTES);
        }
        __syncwarp();
        if (lane < MT) tma_load_4d(smem_x + sx * X_SLOT + lane * X_TILE, &tm_x, &full_x[sx], 0, lane * 128 - 1, q, g.n);
        if (++sx == XS) { sx = 0; px ^= 1u; }
      }
    }
  } else if (warp == 1) {
    // ---- MMA issue of both stages (whole warp: warp-uniform state in uniform registers; an elected lane issues) ----
    const uint32_t w2_base = smem_u32(smem_w2), wh_base = smem_u32(smem_wh), mid_base = smem_u32(smem_mid);
    mbar_wait(w_full, 0);
    int sx = 0;
    uint32_t px = 0;
    long long w_e1 = 0, w_e2 = 0, w_x = 0, w_mid = 0;
    const long long dbg_t0 = clock64();
    Cursor c1, c2;
    c1.item = c2.item = blockIdx.x;
    c1.c0 = c2.c0 = 0u;
    cursor_load(c1, p, 1);
    cursor_load(c2, p, 2);
    uint32_t mid0 = 0;          // ring position (running count) of the first conv2 row of stage 2's item
    int a2_2 = c2.item < p.total_items ? decode(p, c2.item).a2 : 0, n2_2 = c2.item < p.total_items ? decode(p, c2.item).n2 : 0;
    while (c2.item < p.total_items) {
      const bool live1 = c1.item < p.total_items;
      // stage 2 reads conv2 row base - 1 + i of its item: ring position k2; go once LAG rows from it on are committed
      const uint32_t k2 = mid0 + static_cast<uint32_t>(c2.base - 1 + c2.i - a2_2);
      const bool do2 = !live1 || c1.c0 + static_cast<uint32_t>(c1.committed) >= k2 + LAG;
      StepDesc d1, d2;
      uint64_t* rel1 = nullptr;
      uint64_t* rel2 = nullptr;
      if (live1) {
        d1 = step_prepare(c1, smem_u32(smem_x + sx * X_SLOT), w2_base, tmem_base + T1, acc1_empty, dbg, w_e1);
        const long long tw = dbg ? clock64() : 0;
        mbar_wait(&full_x[sx], px);
        if (dbg) w_x += clock64() - tw;
        rel1 = &empty_x[sx];
        if (++sx == XS) { sx = 0; px ^= 1u; }
      }
      if (do2) {
        d2 = step_prepare(c2, mid_base + (k2 % S2) * MID_ROW, wh_base, tmem_base + T2, acc2_empty, dbg, w_e2);
        const long long tw = dbg ? clock64() : 0;
        mbar_wait(&mid_full[k2 % S2], (k2 / S2) & 1u);
        if (dbg) w_mid += clock64() - tw;
        rel2 = &mid_empty[k2 % S2];
      }
      tc_fence_after();
      if (elect_one()) {
        // three straight-line variants: the operands are moved to uniform registers once, the MMAs issue back to back
        if (live1 && do2) issue_mmas<MT, true, true>(d1, d2, tmem_base + T1, tmem_base + T2);
        else if (live1) issue_mmas<MT, true, false>(d1, d2, tmem_base + T1, tmem_base + T2);
        else issue_mmas<MT, false, true>(d1, d2, tmem_base + T1, tmem_base + T2);
        if (live1) {
          umma_commit(rel1);
          for (int cc = c1.committed; cc <= d1.done; ++cc) umma_commit(&acc1_full[(c1.c0 + cc) % SL]);
        }
        if (do2) {
          umma_commit(rel2);
          for (int cc = c2.committed; cc <= d2.done; ++cc) umma_commit(&acc2_full[(c2.c0 + cc) % SL]);
        }
      }
      __syncwarp();
      if (live1) {
        if (d1.done >= c1.committed) c1.committed = d1.done + 1;
        cursor_advance(c1, p, 1);
      }
      if (do2) {
        if (d2.done >= c2.committed) c2.committed = d2.done + 1;
        const int before = c2.item;
        cursor_advance(c2, p, 2);
        if (c2.item != before) {
          mid0 += static_cast<uint32_t>(n2_2);
          if (c2.item < p.total_items) {
            const ItemGeo g = decode(p, c2.item);
            a2_2 = g.a2;
            n2_2 = g.n2;
          }
        }
      }
    }
    if (dbg && blockIdx.x == 0 && lane == 0) {
      p.dbg[0] = clock64() - dbg_t0; p.dbg[1] = w_x; p.dbg[2] = w_e1; p.dbg[3] = w_mid; p.dbg[4] = w_e2;
    }
  } else if (warp < 6) {
    // ---- conv2 epilogue: thread = TMEM lane = pixel of the tile; TMEM -> BN / ReLU -> bf16 -> the shared-memory row ring ----
    const int quarter = warp & 3;
    const int m = quarter * 32 + lane;
    const uint32_t t_lane = tmem_base + T1 + (static_cast<uint32_t>(quarter * 32) << 16);
    const uint32_t mid_base = smem_u32(smem_mid);
    uint32_t c = 0;                   // running count of conv2 rows
    long long w_f1 = 0, w_me = 0;
    const long long dbg_t0 = clock64();
    for (int item = blockIdx.x; item < p.total_items; item += gridDim.x) {
      const int n2 = decode(p, item).n2;
#pragma unroll 1
      for (int o = 0; o < n2; ++o, ++c) {
        const uint32_t slot = c % SL, ms = c % S2;
        const long long tw = dbg ? clock64() : 0;
        mbar_wait(&acc1_full[slot], (c / SL) & 1u);
        const long long tw2 = dbg ? clock64() : 0;
        mbar_wait(&mid_empty[ms], ((c / S2) & 1u) ^ 1u);
        if (dbg) { w_f1 += tw2 - tw; w_me += clock64() - tw2; }
        tc_fence_after();
        uint32_t v[MT][16];
#pragma unroll
        for (int t = 0; t < MT; ++t) tmem_ld_x16(t_lane + (t * SL + (SL - 1 - slot)) * C, v[t]);
        tmem_ld_wait();
#pragma unroll
        for (int t = 0; t < MT; ++t) tmem_st_zero_x16(t_lane + (t * SL + (SL - 1 - slot)) * C);
#pragma unroll
        for (int t = 0; t < MT; ++t) {
          uint32_t h[8];
#pragma unroll
          for (int u = 0; u < 16; u += 4) {
            const float4 sc = __ldg(reinterpret_cast<const float4*>(p.scale2 + u));
            const float4 sh = __ldg(reinterpret_cast<const float4*>(p.shift2 + u));
            h[u / 2] = pack_bf16x2(fmaxf(fmaf(__uint_as_float(v[t][u]), sc.x, sh.x), 0.f),
                                   fmaxf(fmaf(__uint_as_float(v[t][u + 1]), sc.y, sh.y), 0.f));
            h[u / 2 + 1] = pack_bf16x2(fmaxf(fmaf(__uint_as_float(v[t][u + 2]), sc.z, sh.z), 0.f),
                                       fmaxf(fmaf(__uint_as_float(v[t][u + 3]), sc.w, sh.w), 0.f));
          }
          const uint32_t dst = mid_base + ms * MID_ROW + static_cast<uint32_t>(t * 128 + m + 1) * ROW_BYTES;
          st_shared_v4(sw32(dst), h[0], h[1], h[2], h[3]);
          st_shared_v4(sw32(dst + 16), h[4], h[5], h[6], h[7]);
        }
        fence_proxy_async_smem();          // the row is read by the tensor core (async proxy)
        mbar_arrive(&mid_full[ms]);
        tmem_st_wait();
        tc_fence_before();
        mbar_arrive(&acc1_empty[slot]);
      }
    }
    if (dbg && blockIdx.x == 0 && threadIdx.x == 64) { p.dbg[5] = clock64() - dbg_t0; p.dbg[6] = w_f1; p.dbg[7] = w_me; }
  } else {
    // ---- head epilogue: bias, logits in either layout, first-max argmax ----
    const int quarter = warp & 3;
    const int m = quarter * 32 + lane;
    const uint32_t t_lane = tmem_base + T2 + (static_cast<uint32_t>(quarter * 32) << 16);
    const int64_t hw = static_cast<int64_t>(p.H) * p.W;
    uint32_t c = 0;                   // running count of head rows
    long long w_f2 = 0;
    const long long dbg_t0 = clock64();
    for (int item = blockIdx.x; item < p.total_items; item += gridDim.x) {
      const ItemGeo g = decode(p, item);
#pragma unroll 1
      for (int o = 0; o < g.n3; ++o, ++c) {
        const uint32_t slot = c % SL;
        const long long tw = dbg ? clock64() : 0;
        mbar_wait(&acc2_full[slot], (c / SL) & 1u);
        if (dbg) w_f2 += clock64() - tw;
        tc_fence_after();
        uint32_t v[MT][16];
#pragma unroll
        for (int t = 0; t < MT; ++t) tmem_ld_x16(t_lane + (t * SL + (SL - 1 - slot)) * C, v[t]);
        tmem_ld_wait();
#pragma unroll
        for (int t = 0; t < MT; ++t) tmem_st_zero_x16(t_lane + (t * SL + (SL - 1 - slot)) * C);
        tmem_st_wait();
        tc_fence_before();
        mbar_arrive(&acc2_empty[slot]);
#pragma unroll
        for (int t = 0; t < MT; ++t) {
          const int64_t in_img = static_cast<int64_t>(g.y0 + o) * p.W + t * 128 + m;
          const int64_t pix = static_cast<int64_t>(g.n) * hw + in_img;
          int best = 0;
          float bv = 0.f;
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            if (k >= p.K) break;
            const float z = __uint_as_float(v[t][k]) + __ldg(p.bias + k);
            if (p.logits_nhwc) p.logits_nhwc[pix * p.K + k] = __float2bfloat16_rn(z);
            if (p.logits_nchw) p.logits_nchw[(static_cast<int64_t>(g.n) * p.K + k) * hw + in_img] = z;
            if (k == 0 || z > bv) { bv = z; best = k; }
          }
          if (p.mask) p.mask[pix] = static_cast<uint8_t>(best);
        }
      }
    }
    if (dbg && blockIdx.x == 0 && threadIdx.x == 192) { p.dbg[8] = clock64() - dbg_t0; p.dbg[9] = w_f2; }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

template <int MT>
int launch_tail(const CUtensorMap& tm_x, const CUtensorMap& tm_w2, const CUtensorMap& tm_wh, TailParams& p, cudaStream_t s) {
  const int smem = 1024 + 6 * W_S + XS * MT * X_TILE + S2 * MID_ROW + 512;
  static std::once_flag once;
  static cudaError_t attr_err = cudaSuccess;
  std::call_once(once, [] {
    attr_err = cudaFuncSetAttribute(tail_fused_kernel<MT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 225 * 1024);
  });
  DT_CUDA(attr_err);
  const int slots = dt_num_sms() * 2;
  const int grid = p.total_items < slots ? p.total_items : slots;
  static unsigned long long* dbg_buf = nullptr;
  const bool dbg = getenv("DT_ROW_DEBUG") != nullptr;
  if (dbg && !dbg_buf) cudaMalloc(&dbg_buf, 128);
  p.dbg = dbg ? dbg_buf : nullptr;
  tail_fused_kernel<MT><<<grid, kThreads, smem, s>>>(tm_x, tm_w2, tm_wh, p);
  DT_LAUNCH_CHECK();
  if (dbg) {
    unsigned long long h[10];
    cudaStreamSynchronize(s);
    cudaMemcpy(h, dbg_buf, sizeof(h), cudaMemcpyDeviceToHost);
    fprintf(stderr, "[tail N=%d %dx%d items=%d grid=%d] mma total %llu wait full_x %llu acc1_empty %llu mid_full %llu acc2_empty %llu | "
            "epi1 total %llu wait acc1_full %llu mid_empty %llu | epi2 total %llu wait acc2_full %llu\n", p.N, p.H, p.W, p.total_items,
            grid, h[0], h[1], h[2], h[3], h[4], h[5], h[6], h[7], h[8], h[9]);
  }
  return DT_OK;
}

}  // namespace

int dt_encode_bf16_map(CUtensorMap* tm, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                       const uint32_t* box, const uint32_t* elem_strides);

extern "C" int dt_tail_fused(const void* x, int N, int H, int W, int K, const void* w2_packed, const float* scale2,
                             const float* shift2, const void* wh_packed, const float* bias16, float* logits_nchw,
                             void* logits_nhwc, uint8_t* mask, dt_stream_t stream) {
  DT_ARCH_GUARD();
  DT_REQUIRE(x && w2_packed && wh_packed && scale2 && shift2 && bias16 && N > 0 && K >= 1 && K <= 4, DT_ERR_BAD_SHAPE,
             "dt_tail_fused: bad arguments (K=%d)", K);
  DT_REQUIRE((W == 128 || W == 256) && H >= 8, DT_ERR_UNSUPPORTED, "dt_tail_fused: width %d (128 or 256), height %d (>= 8)", W, H);
  DT_REQUIRE((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(w2_packed) | reinterpret_cast<uintptr_t>(wh_packed)) % 16 == 0,
             DT_ERR_BAD_ALIGN, "dt_tail_fused: tensors must be 16-byte aligned");
  TailParams p;
  memset(&p, 0, sizeof(p));
  p.N = N; p.H = H; p.W = W; p.K = K;
  p.mt = W / 128;
  p.R = H < 32 ? H : 32;
  p.chunks = (H + p.R - 1) / p.R;
  p.total_items = N * p.chunks;
  p.scale2 = scale2; p.shift2 = shift2; p.bias = bias16;
  p.logits_nchw = logits_nchw;
  p.logits_nhwc = static_cast<__nv_bfloat16*>(logits_nhwc);
  p.mask = mask;
  CUtensorMap tm_x, tm_w2, tm_wh;
  {
    const uint64_t dims[4] = {static_cast<uint64_t>(C), static_cast<uint64_t>(W), static_cast<uint64_t>(H), static_cast<uint64_t>(N)};
    const uint64_t strides[3] = {static_cast<uint64_t>(C) * 2, static_cast<uint64_t>(W) * C * 2, static_cast<uint64_t>(H) * W * C * 2};
    const uint32_t box[4] = {C, 130, 1, 1};
    int rc = dt_encode_bf16_map(&tm_x, x, 4, dims, strides, box, nullptr);
    if (rc != DT_OK) return rc;
  }
  {  // [16][192] (k = (r*3 + s)*16 + ci) seen as (ci, co, r, s): one box = the [3 * 16][16] matrix of a horizontal tap
    const uint64_t dims[4] = {C, C, 3, 3};
    const uint64_t strides[3] = {192 * 2, 3 * C * 2, C * 2};
    const uint32_t box[4] = {C, C, 3, 1};
    int rc = dt_encode_bf16_map(&tm_w2, w2_packed, 4, dims, strides, box, nullptr);
    if (rc == DT_OK) rc = dt_encode_bf16_map(&tm_wh, wh_packed, 4, dims, strides, box, nullptr);
    if (rc != DT_OK) return rc;
  }
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  return p.mt == 2 ? launch_tail<2>(tm_x, tm_w2, tm_wh, p, s) : launch_tail<1>(tm_x, tm_w2, tm_wh, p, s);
}
