// Micro-benchmark: clocks per tcgen05.mma (cta_group::1, M = 128, K = 16, bf16, both operands from shared memory) as a
// function of N, of the number of accumulators the MMAs rotate over, and of the number of co-resident CTAs per SM.
// Operands are whatever the (zeroed) shared memory holds: only the issue / execution rate is measured.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I deadtrees_b200/csrc scripts/exp/mma_rate.cu -o /tmp/mma_rate
#include <cstdio>
#include <cstdlib>
#include "common.cuh"

void dt_set_error(const char*, ...) {}
int dt_check_device() { return 0; }

__global__ void __launch_bounds__(64, 1) mma_rate_kernel(int N, int n_acc, int iters, int a_bytes_step, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t done;
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 48 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) { mbar_init(&done, 1u); fence_mbar_init(); }
  if (warp == 1) { tmem_alloc(&tmem_slot, 512); tmem_relinquish(); }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, tmem_slot, 0);
  if (warp == 1) {
    const uint64_t a_d = umma_desc(smem_u32(smem), 1024u, 2u);                 // 128 rows x 128 B, SWIZZLE_128B
    const uint64_t b_d = umma_desc(smem_u32(smem + 16384), 1024u, 2u);         // up to 256 rows x 128 B
    const uint32_t idesc = umma_idesc_bf16(128, N);
    long long t0 = 0, t1 = 0;
    for (int rep = 0; rep < 2; ++rep) {        // rep 0 warms up
      t0 = clock64();
      if (elect_one()) {
        for (int i = 0; i < iters; ++i) {
#pragma unroll 1
          for (int a = 0; a < n_acc; ++a)
            umma_bf16_ss(tmem_base + a * N, a_d + (((i & 3) * 32 + a * a_bytes_step) >> 4), b_d + (((i & 3) * 32) >> 4), idesc, 1u);
        }
        umma_commit(&done);
      }
      __syncwarp();
      mbar_wait(&done, rep & 1);
      t1 = clock64();
    }
    if (lane == 0 && blockIdx.x == 0) out[0] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}

int main() {
  long long* d_out;
  cudaMalloc(&d_out, 8);
  cudaFuncSetAttribute(mma_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  const int iters = 2000;
  printf("%5s %6s %6s | %12s %12s\n", "N", "n_acc", "ctas", "clk/MMA(CTA)", "clk/MMA(SM)");
  for (int ctas = 1; ctas <= 2; ++ctas) {
    const int smem = ctas == 1 ? 100 * 1024 : 60 * 1024;     // one or two CTAs fit per SM
    for (int N : {16, 32, 48, 64, 96, 128, 192, 256}) {
      for (int n_acc : {1, 2, 4}) {
        if (n_acc * N > 512 / (ctas == 2 ? 1 : 1)) continue;
        mma_rate_kernel<<<148 * ctas, 64, smem>>>(N, n_acc, iters / n_acc, 0, d_out);
        long long h = 0;
        if (cudaMemcpy(&h, d_out, 8, cudaMemcpyDeviceToHost) != cudaSuccess) { printf("error %s\n", cudaGetErrorString(cudaGetLastError())); return 1; }
        const double per = double(h) / double((iters / n_acc) * n_acc);
        printf("%5d %6d %6d | %12.1f %12.1f\n", N, n_acc, ctas, per, per / ctas);
      }
    }
  }
  return 0;
}
