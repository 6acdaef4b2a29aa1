// Row-streaming 3x3 / stride-1 / pad-1 convolution for the narrow layers (C_in, C_out in {16, 32, 64}) - tcgen05 implicit
// GEMM whose three VERTICAL taps ride in the N dimension of one MMA.
//
// Why: with C_out <= 64 a `tcgen05.mma` (M = 128, K = 16) is bound by the 4 KB of A it fetches from shared memory, not by
// the tensor pipe ("shared-memory operand fetch"): one MMA per tap reads every pixel nine times.  Here an
// M tile is 128 pixels of ONE image row q, and
//     D[pixel][(r, co)] += sum_ci x[q][pixel + s - 1][ci] * w[r][s][ci][co]          (one MMA per (s, 16 channels), N = 3 * C_out)
// where the column block r is the contribution of input row q to OUTPUT row q + 1 - r.  The accumulators of consecutive
// output rows sit in consecutive TMEM column slots in DESCENDING row order, so the (r = 0, 1, 2) blocks of one MMA land
// exactly on the accumulators of rows q + 1, q, q - 1: the vertical taps cost no extra A reads, every input row is
// fetched three times (the horizontal taps) instead of nine.  Per output accumulator the order of the additions is
// (r, s, ci) - the k order of the other kernels - so the results are bit-identical to conv_res / conv_halo.
//
// The epilogue warps read a finished row's accumulator (TMEM lane = pixel), apply scale / shift (+ residual) (+ ReLU), store
// the row, CLEAR the slot (tcgen05.st) and hand it back; every MMA runs with accumulate = 1.  Slots form a ring; an MMA
// whose three targets straddle the ring's wrap point is issued as two.
//
// A tile in shared memory: 16 groups of 8 pixels, each group with its own halo pixel left and right (pitch 10 / 12 pixels,
// a multiple of 128 bytes, one small TMA box per group, out-of-image pixels zero-filled by TMA) - the same pitched
// layout as the 8 x 16 halo patches, so the taps are shifted UMMA descriptors (conv_halo.cu).  Images 64 pixels wide put
// the same row of two images into one M tile (lanes 0-63 / 64-127).
//
// MT = 2 M tiles (neighbouring units, same rows) share a pipeline step and the weights; C_out <= 32 runs two CTAs per SM.
// Measured on B200 (405 tiles, per layer): layer1 64->64 @64: 145 / 183 us (8 x 16 tile kernel, without / with residual)
// -> 122 / 170; decoder.blocks.2.conv2 138 -> 118.  Two things mattered more than the MMA count: (1) NO indexed local
// arrays in the issuing thread (a local-memory load misses the small L1 next to the epilogue's streaming traffic:
// ~680 clk per MMA), (2) the issuing warp keeps its state warp-uniform and issues through elect.sync, so descriptors and
// TMEM addresses live in uniform registers (3 instead of ~25 instructions per MMA; DT_ROW_DEBUG=1 prints the per-role
// wait / busy clocks of block 0).
//
//   warp 0: TMA producer   warp 1: TMEM alloc + MMA issue (whole warp, elected lane)   warps 2..5: epilogue
#include <cstring>
#include <mutex>

#include "common.cuh"

namespace {

constexpr int kThreads = 192;
constexpr int GROUPS = 16;
constexpr int MAX_A = 8;
constexpr int MAX_SLOTS = 16;

struct RowParams {
  int N, H, W, C_out;
  int ipt;          // images per M tile: 2 when W == 64
  int col_blocks;   // 128-pixel column blocks per row (1 when W == 64)
  int R, chunks;    // output rows per work item, work items per image height
  int units;        // M-tile columns: image(-pair)s x column blocks; a work item covers MT consecutive units
  int total_items;
  int a_slots;
  int relu, has_residual;
  const __nv_bfloat16* residual;
  __nv_bfloat16* y;
  const float* scale;
  const float* shift;
  // head epilogue (EPI == 1): the first K of the 16 output channels are the class logits
  int K;
  float* logits_nchw;
  __nv_bfloat16* logits_nhwc;
  uint8_t* mask;
  unsigned long long* dbg;
};

struct Item {
  int u0, y0, rows;   // first unit, first output row, output rows
};
__device__ __forceinline__ Item decode(const RowParams& p, int item) {
  Item it;
  const int ch = item % p.chunks;
  it.u0 = (item / p.chunks);
  it.y0 = ch * p.R;
  it.rows = p.H - it.y0 < p.R ? p.H - it.y0 : p.R;
  return it;
}
// unit -> first image and first column of its M tile
__device__ __forceinline__ void unit_origin(const RowParams& p, int unit, int& n0, int& col0) {
  n0 = (unit / p.col_blocks) * p.ipt;
  col0 = (unit % p.col_blocks) * 128;
}

// DENSE (width a multiple of 128): an M tile is 128 consecutive pixels of a row, stored with its two halo pixels as one
// dense run of 130 pixels - ONE TMA box per tile and row.  Pitched (width 64, two images per tile): 16 groups of 8 pixels
// with private halos, one small box per group (the TMA unit serves ~1 box per 68 clk per SM, measured: 64 boxes of 320
// bytes per step made the producer the bottleneck of the 16-channel layers).
template <int CIN, int COUT, bool DENSE>
struct RowCfg {
  static constexpr int ROW_BYTES = CIN * 2;
  static constexpr int PITCH = DENSE ? 8 : (CIN == 16 ? 12 : 10);   // group pitch in pixels (pitched: a multiple of 128 bytes)
  static constexpr int GROUP_BYTES = PITCH * ROW_BYTES;
  static constexpr int A_TILE = DENSE ? ((130 * ROW_BYTES + 1023) / 1024) * 1024 : GROUPS * GROUP_BYTES;
  static constexpr int B_S = ((3 * COUT * ROW_BYTES + 1023) / 1024) * 1024;   // weights of one horizontal tap: [3 * COUT][CIN]
  static constexpr int W_BYTES = 3 * B_S;
  // MT M tiles (neighbouring units, same rows) advance in lockstep so that consecutive MMAs go to different accumulators:
  // a chain of MMAs into ONE accumulator runs at the MMA latency (measured ~300 clk for N = 48), not at the feed rate
  // Two CTAs per SM for C_out <= 32 (half of TMEM each): the MMAs of ONE CTA retire at ~130 clk apiece whatever their N
  // (measured, N = 48 ... 192), those of two CTAs overlap.
  static constexpr int CTAS = COUT == 64 ? 1 : 2;
  static constexpr int MT = 2;
  static constexpr int SLOTS = COUT == 16 ? 8 : 4;
  static constexpr int TMEM_COLS = MT * SLOTS * COUT;               // 256 / 256 / 512
  static constexpr uint32_t LAYOUT = CIN == 64 ? 2u : (CIN == 32 ? 4u : 6u);
  static constexpr int KSTEPS = CIN / 16;
};

template <int CIN, int COUT, int EPI, bool DENSE>
__global__ void __launch_bounds__(kThreads, 1)
conv_row_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b, const RowParams p) {
  using Cfg = RowCfg<CIN, COUT, DENSE>;
  constexpr int ROW_BYTES = Cfg::ROW_BYTES, GROUP_BYTES = Cfg::GROUP_BYTES, A_TILE = Cfg::A_TILE, B_S = Cfg::B_S;
  constexpr int SLOTS = Cfg::SLOTS, TMEM_COLS = Cfg::TMEM_COLS, KSTEPS = Cfg::KSTEPS, MT = Cfg::MT;
  constexpr int A_SLOT = MT * A_TILE;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_w = smem;
  uint8_t* smem_a = smem + Cfg::W_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_a + p.a_slots * A_SLOT);
  uint64_t* w_full = bars;                    // [1]
  uint64_t* full_a = w_full + 1;              // [MAX_A]
  uint64_t* empty_a = full_a + MAX_A;         // [MAX_A]
  uint64_t* acc_full = empty_a + MAX_A;       // [MAX_SLOTS]
  uint64_t* acc_empty = acc_full + MAX_SLOTS; // [MAX_SLOTS]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + MAX_SLOTS);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_a);
    tma_prefetch_desc(&tm_b);
    mbar_init(w_full, 1u);
    for (int i = 0; i < p.a_slots; ++i) { mbar_init(&full_a[i], 1u); mbar_init(&empty_a[i], 1u); }
    for (int i = 0; i < SLOTS; ++i) { mbar_init(&acc_full[i], 1u); mbar_init(&acc_empty[i], 128u); }
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  // broadcast from lane 0: tells the compiler the TMEM base is warp-uniform, so the MMA operands stay in uniform registers
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);
  if (warp >= 2) {   // all accumulators start at zero: every MMA accumulates
    const uint32_t t_lane = tmem_base + (static_cast<uint32_t>((warp & 3) * 32) << 16);
#pragma unroll 1
    for (int c = 0; c < TMEM_COLS; c += 16) tmem_st_zero_x16(t_lane + c);
    tmem_st_wait();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();

  if (warp == 0) {
    // ---- producer: one TMA box (10 pixels of one row) per group, lanes 0..15 ----
    if (lane == 0) {
      mbar_arrive_expect_tx(w_full, 3u * 3u * COUT * ROW_BYTES);
      for (int s = 0; s < 3; ++s) tma_load_4d(smem_w + s * B_S, &tm_b, w_full, 0, 0, 0, s);
    }
    int sa = 0;
    uint32_t pa = 0;
    long long dbg_w0 = 0;
    const long long dbg_t0 = clock64();
    for (int item = blockIdx.x; item < p.total_items; item += gridDim.x) {
      const Item it = decode(p, item);
      // lane < 16: group `lane` of every tile t of the item
      int g_n[MT], g_px[MT];
      uint32_t bytes = 0;
#pragma unroll
      for (int t = 0; t < MT; ++t) {
        int n0, col0;
        unit_origin(p, it.u0 * MT + t, n0, col0);
        const bool unit_ok = it.u0 * MT + t < p.units;
        g_n[t] = unit_ok ? n0 + (p.ipt == 2 ? (lane >> 3) : 0) : p.N;         // p.N = "no such image": nothing to load
        g_px[t] = (p.ipt == 2 ? (lane & 7) : lane) * 8 + col0 - 1;
        bytes += !unit_ok ? 0u : (DENSE ? 130u : ((p.ipt == 2 && n0 + 1 >= p.N) ? 8u : 16u) * 10u) * ROW_BYTES;
      }
      int d_n = p.N, d_px = 0;                                                  // DENSE: lane t loads tile t
      if (DENSE && lane < MT && it.u0 * MT + lane < p.units) {
        int n0, col0;
        unit_origin(p, it.u0 * MT + lane, n0, col0);
        d_n = n0;
        d_px = col0 - 1;
      }
      for (int i = 0; i <= it.rows + 1; ++i) {
        const int q = it.y0 - 1 + i;
        if (q < 0 || q >= p.H) continue;
        if (lane == 0) {
          const long long t0 = p.dbg ? clock64() : 0;
          mbar_wait(&empty_a[sa], pa ^ 1u);
          if (p.dbg) dbg_w0 += clock64() - t0;
          mbar_arrive_expect_tx(&full_a[sa], bytes);
        }
        __syncwarp();
        if (DENSE) {
          if (lane < MT && d_n < p.N)     // lane t: the 130-pixel run of tile t
            tma_load_4d(smem_a + sa * A_SLOT + lane * A_TILE, &tm_a, &full_a[sa], 0, d_px, q, d_n);
        } else if (lane < GROUPS) {
#pragma unroll
          for (int t = 0; t < MT; ++t)
            if (g_n[t] < p.N)
              tma_load_4d(smem_a + sa * A_SLOT + t * A_TILE + lane * GROUP_BYTES, &tm_a, &full_a[sa], 0, g_px[t], q, g_n[t]);
        }
        if (++sa == p.a_slots) { sa = 0; pa ^= 1u; }
      }
    }
    if (p.dbg && blockIdx.x == 0 && lane == 0) { p.dbg[0] = clock64() - dbg_t0; p.dbg[1] = dbg_w0; }
  } else if (warp == 1) {
    // ---- MMA issuer: the whole warp runs the (warp-uniform) bookkeeping so that descriptors, TMEM addresses and loop
    // state live in uniform registers; lane 0 issues.  (Inside an `if (lane == 0)` region every operand needed an
    // R2UR / ELECT round trip: ~200 clk per MMA of pure issue - measured - with the tensor pipe 80 % idle.)
    {
      const uint64_t a_hi = umma_desc(0u, GROUP_BYTES, Cfg::LAYOUT);
      const uint64_t b_hi = umma_desc(0u, 8 * ROW_BYTES, Cfg::LAYOUT);
      const uint32_t b_base = smem_u32(smem_w);
      constexpr uint32_t idesc0 = umma_idesc_bf16(128, 0);     // N is or-ed in per MMA
      mbar_wait(w_full, 0);
      int sa = 0;
      uint32_t pa = 0, c0 = 0;       // c0: running count of output rows of this CTA (ring position)
      long long dbg_we = 0, dbg_wf = 0, dbg_rows = 0;
      const long long dbg_t0 = clock64();
      for (int item = blockIdx.x; item < p.total_items; item += gridDim.x) {
        const Item it = decode(p, item);
        int opened = 0, committed = 0;
        const int n_tiles = p.units - it.u0 * MT < MT ? p.units - it.u0 * MT : MT;
        const int last_i = it.y0 + it.rows < p.H ? it.rows + 1 : it.rows;     // last input row that exists
        for (int i = 0; i <= it.rows + 1; ++i) {
          const int q = it.y0 - 1 + i;
          if (q < 0 || q >= p.H) continue;
          // accumulators this input row touches for the first time: wait until the epilogue has cleared them
          const int hi = i < it.rows ? i : it.rows - 1;
          long long t0 = p.dbg ? clock64() : 0;
          for (; opened <= hi; ++opened) {
            const uint32_t c = c0 + opened;
            mbar_wait(&acc_empty[c % SLOTS], ((c / SLOTS) & 1u) ^ 1u);
          }
          long long t1 = p.dbg ? clock64() : 0;
          mbar_wait(&full_a[sa], pa);
          if (p.dbg) { dbg_we += t1 - t0; dbg_wf += clock64() - t1; ++dbg_rows; }
          tc_fence_after();
          // vertical taps r with an output row inside this item: output i - r; consecutive r = ascending TMEM columns,
          // cut where the ring wraps
          const int r_lo = i - (it.rows - 1) > 0 ? i - (it.rows - 1) : 0, r_hi = i < 2 ? i : 2;
          const uint32_t pos_lo = SLOTS - 1 - ((c0 + i - r_lo) % SLOTS);      // column slot of tap r_lo
          const int n_taps = r_hi - r_lo + 1;
          const int len_a = n_taps < static_cast<int>(SLOTS - pos_lo) ? n_taps : static_cast<int>(SLOTS - pos_lo);
          const int len_b = n_taps - len_a;                                   // taps past the wrap point start at slot 0
          const uint32_t col_a = tmem_base + pos_lo * COUT, col_b = tmem_base;
          const uint32_t wb_a = b_base + r_lo * COUT * ROW_BYTES, wb_b = wb_a + len_a * COUT * ROW_BYTES;
          const uint32_t id_a = idesc0 | (static_cast<uint32_t>((len_a * COUT) >> 3) << 17);
          const uint32_t id_b = idesc0 | (static_cast<uint32_t>((len_b * COUT) >> 3) << 17);
          const uint64_t a_d0 = a_hi + (smem_u32(smem_a + sa * A_SLOT) >> 4);
          const uint64_t b_a0 = b_hi + (wb_a >> 4), b_b0 = b_hi + (wb_b >> 4);
          const int done = i == last_i ? it.rows - 1 : i - 2;     // output rows that have all their contributions now
          if (elect_one()) {
            if (len_b == 0) {
#pragma unroll
              for (int s = 0; s < 3; ++s)
#pragma unroll
                for (int k = 0; k < KSTEPS; ++k)
#pragma unroll
                  for (int t = 0; t < MT; ++t)      // tile inner: consecutive MMAs accumulate into different TMEM regions
                    if (t < n_tiles)
                      umma_bf16_ss(col_a + t * (SLOTS * COUT), a_d0 + ((t * A_TILE + s * ROW_BYTES + k * 32) >> 4),
                                   b_a0 + ((s * B_S + k * 32) >> 4), id_a, 1u);
            } else {
#pragma unroll
              for (int s = 0; s < 3; ++s)
#pragma unroll
                for (int k = 0; k < KSTEPS; ++k)
#pragma unroll
                  for (int t = 0; t < MT; ++t)
                    if (t < n_tiles) {
                      const uint64_t a_d = a_d0 + ((t * A_TILE + s * ROW_BYTES + k * 32) >> 4);
                      umma_bf16_ss(col_a + t * (SLOTS * COUT), a_d, b_a0 + ((s * B_S + k * 32) >> 4), id_a, 1u);
                      umma_bf16_ss(col_b + t * (SLOTS * COUT), a_d, b_b0 + ((s * B_S + k * 32) >> 4), id_b, 1u);
                    }
            }
            umma_commit(&empty_a[sa]);
            for (int cc = committed; cc <= done; ++cc) umma_commit(&acc_full[(c0 + cc) % SLOTS]);
          }
          if (done >= committed) committed = done + 1;
          __syncwarp();
          if (++sa == p.a_slots) { sa = 0; pa ^= 1u; }
        }
        c0 += it.rows;
      }
      if (p.dbg && blockIdx.x == 0 && lane == 0) { p.dbg[2] = clock64() - dbg_t0; p.dbg[3] = dbg_we; p.dbg[4] = dbg_wf; p.dbg[5] = dbg_rows; }
    }
  } else {
    // ---- epilogue: thread = TMEM lane = pixel of the M tile ----
    const int quarter = warp & 3;
    const int m = quarter * 32 + lane;
    const int g = m >> 3, j = m & 7;
    const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
    uint32_t c0 = 0;
    long long dbg_w = 0;
    const long long dbg_t0 = clock64();
    for (int item = blockIdx.x; item < p.total_items; item += gridDim.x) {
      const Item it = decode(p, item);
      int64_t pix0[MT];
      bool valid[MT];
#pragma unroll
      for (int t = 0; t < MT; ++t) {
        int n0, col0;
        unit_origin(p, it.u0 * MT + t, n0, col0);
        const int n = n0 + (p.ipt == 2 ? (g >> 3) : 0);
        const int x = col0 + (p.ipt == 2 ? (g & 7) : g) * 8 + j;
        valid[t] = it.u0 * MT + t < p.units && n < p.N;
        pix0[t] = (static_cast<int64_t>(valid[t] ? n : 0) * p.H + it.y0) * p.W + x;
      }
#pragma unroll 1
      for (int o = 0; o < it.rows; ++o) {
        const uint32_t c = c0 + o;
        const uint32_t slot = c % SLOTS;
        uint4 res[COUT / 8];
        if (EPI == 0 && p.has_residual && valid[0]) {
          const int64_t off0 = (pix0[0] + static_cast<int64_t>(o) * p.W) * p.C_out;
#pragma unroll
          for (int u = 0; u < COUT / 16; ++u) ldg_v8(p.residual + off0 + 16 * u, res[2 * u], res[2 * u + 1]);
        }
        const long long tw = p.dbg ? clock64() : 0;
        mbar_wait(&acc_full[slot], (c / SLOTS) & 1u);
        if (p.dbg) dbg_w += clock64() - tw;
        tc_fence_after();
#pragma unroll
        for (int t = 0; t < MT; ++t) {
          const int64_t pix = pix0[t] + static_cast<int64_t>(o) * p.W;
          const int64_t out_off = pix * p.C_out;
          const uint32_t t_row = t_lane + (t * SLOTS + (SLOTS - 1 - slot)) * COUT;
          if (EPI == 1) {
            uint32_t v[16];
            tmem_ld_x16(t_row, v);
            tmem_ld_wait();
            tmem_st_zero_x16(t_row);
            if (valid[t]) {
              int best = 0;
              float bv = 0.f;
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                if (k >= p.K) break;
                const float z = __uint_as_float(v[k]) + __ldg(p.shift + k);
                if (p.logits_nhwc) p.logits_nhwc[pix * p.K + k] = __float2bfloat16_rn(z);
                if (p.logits_nchw) {
                  const int64_t hw = static_cast<int64_t>(p.H) * p.W;
                  const int64_t nn = pix / hw;
                  p.logits_nchw[(nn * p.K + k) * hw + (pix - nn * hw)] = z;
                }
                if (k == 0 || z > bv) { bv = z; best = k; }
              }
              if (p.mask) p.mask[pix] = static_cast<uint8_t>(best);
            }
          } else {
            uint4 res_next[COUT / 8];
            if (p.has_residual && t + 1 < MT && valid[t + 1]) {     // the next tile's residual is in flight during this tile
              const int64_t off1 = (pix0[t + 1] + static_cast<int64_t>(o) * p.W) * p.C_out;
#pragma unroll
              for (int u = 0; u < COUT / 16; ++u) ldg_v8(p.residual + off1 + 16 * u, res_next[2 * u], res_next[2 * u + 1]);
            }
#pragma unroll
            for (int cc = 0; cc < COUT; cc += 16) {
              uint32_t v[16];
              tmem_ld_x16(t_row + cc, v);
              tmem_ld_wait();
              tmem_st_zero_x16(t_row + cc);
              float f[16];
#pragma unroll
              for (int u = 0; u < 16; u += 4) {
                const float4 sc = __ldg(reinterpret_cast<const float4*>(p.scale + cc + u));
                const float4 sh = __ldg(reinterpret_cast<const float4*>(p.shift + cc + u));
                f[u] = fmaf(__uint_as_float(v[u]), sc.x, sh.x);
                f[u + 1] = fmaf(__uint_as_float(v[u + 1]), sc.y, sh.y);
                f[u + 2] = fmaf(__uint_as_float(v[u + 2]), sc.z, sh.z);
                f[u + 3] = fmaf(__uint_as_float(v[u + 3]), sc.w, sh.w);
              }
              if (p.has_residual && valid[t]) {
                const uint32_t rr[8] = {res[cc / 8].x, res[cc / 8].y, res[cc / 8].z, res[cc / 8].w,
                                        res[cc / 8 + 1].x, res[cc / 8 + 1].y, res[cc / 8 + 1].z, res[cc / 8 + 1].w};
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                  const float2 h = unpack_bf16x2(rr[u]);
                  f[2 * u] += h.x;
                  f[2 * u + 1] += h.y;
                }
              }
              if (p.relu) {
#pragma unroll
                for (int u = 0; u < 16; ++u) f[u] = fmaxf(f[u], 0.f);
              }
              if (valid[t]) store_bf16x16(p.y + out_off + cc, f);
            }
            if (p.has_residual) {
#pragma unroll
              for (int u = 0; u < COUT / 8; ++u) res[u] = res_next[u];
            }
          }
        }
        tmem_st_wait();
        tc_fence_before();
        mbar_arrive(&acc_empty[slot]);
      }
      c0 += it.rows;
    }
    if (p.dbg && blockIdx.x == 0 && threadIdx.x == 64) { p.dbg[6] = clock64() - dbg_t0; p.dbg[7] = dbg_w; }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

template <int CIN, int COUT, int EPI, bool DENSE>
int launch_row(const CUtensorMap& tm_a, const CUtensorMap& tm_b, RowParams& p, cudaStream_t s) {
  using Cfg = RowCfg<CIN, COUT, DENSE>;
  const int ctas = Cfg::CTAS;
  const int a_slots = (225 * 1024 / ctas - 2048 - Cfg::W_BYTES - 1024) / (Cfg::MT * Cfg::A_TILE);
  if (a_slots < 3) return DT_ERR_UNSUPPORTED;
  p.a_slots = a_slots > MAX_A ? MAX_A : a_slots;
  const int smem = 1024 + Cfg::W_BYTES + p.a_slots * Cfg::MT * Cfg::A_TILE + 512;
  // rows per work item: the cost of a CTA is (items per CTA) * (R + 2) input rows - take the cheapest power-of-two split
  const int slots = dt_num_sms() * ctas;
  p.units = ((p.N + p.ipt - 1) / p.ipt) * p.col_blocks;
  const int per_height = (p.units + Cfg::MT - 1) / Cfg::MT;
  int best_R = p.H;
  long best_cost = -1;
  for (int R = p.H; R >= 8; R >>= 1) {
    const int chunks = (p.H + R - 1) / R;
    const long items = static_cast<long>(per_height) * chunks;
    const long cost = ((items + slots - 1) / slots) * (R + 2);
    if (best_cost < 0 || cost < best_cost) { best_cost = cost; best_R = R; }
  }
  p.R = best_R;
  p.chunks = (p.H + p.R - 1) / p.R;
  p.total_items = per_height * p.chunks;
  static std::once_flag once;
  static cudaError_t attr_err = cudaSuccess;
  std::call_once(once, [] {
    attr_err = cudaFuncSetAttribute(conv_row_kernel<CIN, COUT, EPI, DENSE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 225 * 1024);
  });
  DT_CUDA(attr_err);
  const int grid = p.total_items < slots ? p.total_items : slots;
  static unsigned long long* dbg_buf = nullptr;
  const bool dbg = getenv("DT_ROW_DEBUG") != nullptr;
  if (dbg && !dbg_buf) cudaMalloc(&dbg_buf, 64);
  p.dbg = dbg ? dbg_buf : nullptr;
  conv_row_kernel<CIN, COUT, EPI, DENSE><<<grid, kThreads, smem, s>>>(tm_a, tm_b, p);
  DT_LAUNCH_CHECK();
  if (dbg) {
    unsigned long long h[8];
    cudaStreamSynchronize(s);
    cudaMemcpy(h, dbg_buf, 64, cudaMemcpyDeviceToHost);
    fprintf(stderr, "[row %d->%d epi%d N=%d %dx%d R=%d items=%d grid=%d ctas=%d a_slots=%d] producer total %llu wait_empty %llu | mma total %llu "
            "wait_acc_empty %llu wait_full_a %llu rows %llu | epi total %llu wait_acc_full %llu\n", CIN, COUT, EPI, p.N, p.H, p.W, p.R,
            p.total_items, grid, ctas, p.a_slots, h[0], h[1], h[2], h[3], h[4], h[5], h[6], h[7]);
  }
  return DT_OK;
}

// A: (C, W, H, N) pixels, one box = 10 pixels of one row; B: the [C_out][Kpad] packing (k = (r*3 + s)*C_in + ci) seen as
// (ci, co, r, s) so that one box is the [3 * C_out][C_in] matrix of a horizontal tap, rows ordered (r, co)
int encode_row_maps(CUtensorMap* tm_a, CUtensorMap* tm_b, const void* x, const void* w, int Kpad, int N, int H, int W, int cin,
                    int cout, bool dense);

}  // namespace

int dt_encode_bf16_map(CUtensorMap* tm, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                       const uint32_t* box, const uint32_t* elem_strides);

namespace {
int encode_row_maps(CUtensorMap* tm_a, CUtensorMap* tm_b, const void* x, const void* w, int Kpad, int N, int H, int W, int cin,
                    int cout, bool dense) {
  {
    const uint64_t dims[4] = {static_cast<uint64_t>(cin), static_cast<uint64_t>(W), static_cast<uint64_t>(H),
                              static_cast<uint64_t>(N)};
    const uint64_t strides[3] = {static_cast<uint64_t>(cin) * 2, static_cast<uint64_t>(W) * cin * 2,
                                 static_cast<uint64_t>(H) * W * cin * 2};
    const uint32_t box[4] = {static_cast<uint32_t>(cin), dense ? 130u : 10u, 1, 1};
    int rc = dt_encode_bf16_map(tm_a, x, 4, dims, strides, box, nullptr);
    if (rc != DT_OK) return rc;
  }
  {
    const uint64_t dims[4] = {static_cast<uint64_t>(cin), static_cast<uint64_t>(cout), 3, 3};
    const uint64_t strides[3] = {static_cast<uint64_t>(Kpad) * 2, static_cast<uint64_t>(3 * cin) * 2,
                                 static_cast<uint64_t>(cin) * 2};
    const uint32_t box[4] = {static_cast<uint32_t>(cin), static_cast<uint32_t>(cout), 3, 1};
    int rc = dt_encode_bf16_map(tm_b, w, 4, dims, strides, box, nullptr);
    if (rc != DT_OK) return rc;
  }
  return DT_OK;
}
}  // namespace

// Returns DT_ERR_UNSUPPORTED when the layer does not fit (caller goes on to conv_res.cu / conv_halo.cu).
int dt_conv_row(const dt_conv_desc* d, const void* x, const void* w, int Kpad, const float* scale, const float* shift,
                const void* residual, void* y, cudaStream_t s) {
  const int cin = d->C_in, cout = d->C_out;
  if (d->R != 3 || d->S != 3 || d->stride != 1 || d->pad != 1 || d->C_x != d->C_in || d->upsample ||
      !(cin == 64 || cin == 32 || cin == 16) || !(cout == 64 || cout == 32 || cout == 16) ||
      !(d->W == 64 || d->W % 128 == 0) || d->H < 8)
    return DT_ERR_UNSUPPORTED;
  RowParams p;
  memset(&p, 0, sizeof(p));
  p.N = d->N; p.H = d->H; p.W = d->W; p.C_out = cout;
  p.ipt = d->W == 64 ? 2 : 1;
  p.col_blocks = d->W == 64 ? 1 : d->W / 128;
  p.relu = d->relu; p.has_residual = d->has_residual;
  p.residual = static_cast<const __nv_bfloat16*>(residual);
  p.y = static_cast<__nv_bfloat16*>(y);
  p.scale = scale; p.shift = shift;
  CUtensorMap tm_a, tm_b;
  const bool dense = d->W != 64;
  int rc = encode_row_maps(&tm_a, &tm_b, x, w, Kpad, d->N, d->H, d->W, cin, cout, dense);
  if (rc != DT_OK) return rc;
#define DT_ROW(CI, CO)                                                        \
  if (cin == CI && cout == CO)                                                \
    return dense ? launch_row<CI, CO, 0, true>(tm_a, tm_b, p, s) : launch_row<CI, CO, 0, false>(tm_a, tm_b, p, s);
  DT_ROW(64, 64) DT_ROW(64, 32) DT_ROW(64, 16) DT_ROW(32, 64) DT_ROW(32, 32) DT_ROW(32, 16) DT_ROW(16, 32) DT_ROW(16, 16)
#undef DT_ROW
  return DT_ERR_UNSUPPORTED;
}

// Segmentation head (3x3 conv 16 -> K <= 4, weights padded to 16 output channels, bf16 [16][192]) on the row kernel, with the
// bias / argmax / layout outputs fused into the epilogue.
int dt_head_row(const void* x, int N, int H, int W, int K, const void* w_packed, const float* bias16, float* logits_nchw,
                void* logits_nhwc, uint8_t* mask, cudaStream_t s) {
  if (!(W == 64 || W % 128 == 0) || H < 8 || K < 1 || K > 4) return DT_ERR_UNSUPPORTED;
  RowParams p;
  memset(&p, 0, sizeof(p));
  p.N = N; p.H = H; p.W = W; p.C_out = 16;
  p.ipt = W == 64 ? 2 : 1;
  p.col_blocks = W == 64 ? 1 : W / 128;
  p.shift = bias16;
  p.K = K;
  p.logits_nchw = logits_nchw;
  p.logits_nhwc = static_cast<__nv_bfloat16*>(logits_nhwc);
  p.mask = mask;
  CUtensorMap tm_a, tm_b;
  const bool dense = W != 64;
  int rc = encode_row_maps(&tm_a, &tm_b, x, w_packed, 192, N, H, W, 16, 16, dense);
  if (rc != DT_OK) return rc;
  return dense ? launch_row<16, 16, 1, true>(tm_a, tm_b, p, s) : launch_row<16, 16, 1, false>(tm_a, tm_b, p, s);
}
