// Micro-benchmark 2: what slows the MMAs of the row-streaming kernels (conv_row.cu, conv_tail.cu) below the rates of
// mma_rate.cu?  One warp issues N = 48 MMAs (M 128, K 16, SWIZZLE_32B operands) over n_acc accumulator regions, optionally
// with the row kernels' SLIDING column window (three 16-column slots of a ring, advancing one slot every third MMA), while
// four epilogue warps do what the row kernels' epilogues do to TMEM: tcgen05.ld x16 + tcgen05.st (clear) x16 on a column
// slot that is either inside the region the MMAs work on ("near": the ring slot next to the window) or far away from it.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I deadtrees_b200/csrc scripts/exp/mma_ldst.cu -o /tmp/mma_ldst
#include <cstdio>
#include <cstdlib>
#include "common.cuh"

void dt_set_error(const char*, ...) {}
int dt_check_device() { return 0; }

// epi: 0 no epilogue traffic, 1 ld+st near, 2 ld+st far, 3 ld only near, 4 st only near
__global__ void __launch_bounds__(192, 1) k(int n_acc, int iters, int slide, int epi, int region_cols, int tmem_cols, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t done;
  __shared__ uint32_t tmem_slot;
  __shared__ volatile int stop;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 48 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) { mbar_init(&done, 1u); fence_mbar_init(); stop = 0; }
  if (warp == 1) { tmem_alloc(&tmem_slot, tmem_cols); tmem_relinquish(); }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, tmem_slot, 0);
  const int slots = region_cols / 16;
  if (warp == 1) {
    const uint64_t a_d = umma_desc(smem_u32(smem), 256u, 6u);
    const uint64_t b_d = umma_desc(smem_u32(smem + 16384), 256u, 6u);
    const uint32_t idesc48 = umma_idesc_bf16(128, 48), idesc32 = umma_idesc_bf16(128, 32), idesc16 = umma_idesc_bf16(128, 16);
    long long t0 = 0, t1 = 0;
    for (int rep = 0; rep < 2; ++rep) {
      t0 = clock64();
      if (elect_one()) {
        int pos = 0;                      // first slot of the 3-slot window (slide mode)
        for (int i = 0; i < iters; ++i) {
#pragma unroll 1
          for (int a = 0; a < n_acc; ++a) {
            const uint32_t base = tmem_base + a * region_cols;
            if (!slide || pos + 3 <= slots) {
              umma_bf16_ss(base + (slide ? pos * 16 : 0), a_d + ((a * 4096) >> 4), b_d, idesc48, 1u);
            } else {                      // the window wraps around the ring: two MMAs, as in the kernels
              const int la = slots - pos;
              umma_bf16_ss(base + pos * 16, a_d + ((a * 4096) >> 4), b_d, la == 2 ? idesc32 : idesc16, 1u);
              umma_bf16_ss(base, a_d + ((a * 4096) >> 4), b_d + (la * 16 * 32 >> 4), la == 2 ? idesc16 : idesc32, 1u);
            }
          }
          if (slide && i % 3 == 2) pos = (pos + 1) % slots;
        }
        umma_commit(&done);
      }
      __syncwarp();
      mbar_wait(&done, rep & 1);
      t1 = clock64();
    }
    stop = 1;
    if (lane == 0 && blockIdx.x == 0) out[0] = t1 - t0;
  } else if (warp >= 2 && epi) {
    const uint32_t t_lane = tmem_base + (static_cast<uint32_t>((warp & 3) * 32) << 16);
    // near: the last slot of accumulator region 0 (slide = 0: the MMAs use columns 0..47 of the region, this is 48..63);
    // far: the last 16 columns of the allocation, outside every region
    const uint32_t col = epi == 2 ? tmem_cols - 16 : region_cols - 16;
    uint32_t acc = 0;
    while (!stop) {
      uint32_t v[16];
      if (epi != 4) { tmem_ld_x16(t_lane + col, v); tmem_ld_wait(); acc += v[0]; }
      if (epi != 3) { tmem_st_zero_x16(t_lane + col); tmem_st_wait(); }
    }
    if (acc == 0x12345678u) out[1] = acc;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, tmem_cols); }
}

int main() {
  long long* d_out;
  cudaMalloc(&d_out, 16);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  const int iters = 3000;
  const char* epi_name[5] = {"none", "ld+st near", "ld+st far", "ld near", "st near"};
  printf("%6s %6s %6s %12s %5s | %12s %12s\n", "n_acc", "slide", "region", "epilogue", "ctas", "clk/MMA(CTA)", "clk/MMA(SM)");
  for (int ctas = 1; ctas <= 2; ++ctas) {
    const int smem = ctas == 1 ? 100 * 1024 : 60 * 1024;
    const int tmem_cols = 512 / ctas;
    for (int region : {64, 128})
      for (int slide = 0; slide <= 1; ++slide)
        for (int n_acc : {1, 2, 4}) {
          if (n_acc * region + 16 > tmem_cols) continue;
          for (int epi = 0; epi < 5; ++epi) {
            if (region == 128 && (epi == 3 || epi == 4)) continue;
            for (int rep = 0; rep < 2; ++rep) {
              k<<<148 * ctas, 192, smem>>>(n_acc, iters / n_acc, slide, epi, region, tmem_cols, d_out);
              if (cudaDeviceSynchronize() != cudaSuccess) { printf("error %s\n", cudaGetErrorString(cudaGetLastError())); return 1; }
            }
            long long h = 0;
            cudaMemcpy(&h, d_out, 8, cudaMemcpyDeviceToHost);
            const double per = double(h) / double((iters / n_acc) * n_acc);
            printf("%6d %6d %6d %12s %5d | %12.1f %12.1f\n", n_acc, slide, region, epi_name[epi], ctas, per, per / ctas);
          }
        }
  }
  return 0;
}
