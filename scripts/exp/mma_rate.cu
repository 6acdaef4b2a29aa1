// Micro-benchmark: clocks per tcgen05.mma (cta_group::1, M = 128, K = 16, bf16, both operands from shared memory) as a
// function of N, of the number of accumulators the MMAs rotate over, and of the number of co-resident CTAs per SM.
// Operands are whatever the (zeroed) shared memory holds: only the issue / execution rate is measured.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I deadtrees_b200/csrc scripts/exp/mma_rate.cu -o /tmp/mma_rate
#include <cstdio>
#include <cstdlib>
#include "common.cuh"

void dt_set_error(const char*, ...) {}
int dt_check_device() { return 0; }

// sw32 = 1: both operands as 32-byte rows (SWIZZLE_32B, dense 8-row groups: the 16-channel layers' layout);
// commit_every > 0: a tcgen05.commit (to an mbarrier nobody waits on) after every commit_every-th MMA group
__global__ void __launch_bounds__(64, 1) mma_rate_kernel(int N, int n_acc, int iters, int a_bytes_step, int tmem_cols, int sw32,
                                                         int commit_every, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t done, sink;
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 48 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) { mbar_init(&done, 1u); mbar_init(&sink, 1u); fence_mbar_init(); }
  if (warp == 1) { tmem_alloc(&tmem_slot, tmem_cols); tmem_relinquish(); }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, tmem_slot, 0);
  if (warp == 1) {
    const uint64_t a_d = sw32 ? umma_desc(smem_u32(smem), 256u, 6u) : umma_desc(smem_u32(smem), 1024u, 2u);   // 128 rows
    const uint64_t b_d = sw32 ? umma_desc(smem_u32(smem + 16384), 256u, 6u) : umma_desc(smem_u32(smem + 16384), 1024u, 2u);
    const uint32_t idesc = umma_idesc_bf16(128, N);
    long long t0 = 0, t1 = 0;
    for (int rep = 0; rep < 2; ++rep) {        // rep 0 warms up
      t0 = clock64();
      if (elect_one()) {
        for (int i = 0; i < iters; ++i) {
#pragma unroll 1
          for (int a = 0; a < n_acc; ++a)
            umma_bf16_ss(tmem_base + a * N, a_d + (((i & 3) * 32 + a * a_bytes_step) >> 4), b_d + (((i & 3) * 32) >> 4), idesc, 1u);
          if (commit_every > 0 && i % commit_every == commit_every - 1) umma_commit(&sink);
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
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, tmem_cols); }
}

int main() {
  long long* d_out;
  cudaMalloc(&d_out, 8);
  cudaFuncSetAttribute(mma_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  const int iters = 2000;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  printf("%5s %6s %5s %5s %7s | %12s %12s %10s\n", "N", "n_acc", "ctas", "sw32", "commit", "clk/MMA(CTA)", "clk/MMA(SM)", "kernel_us");
  for (int ctas = 1; ctas <= 2; ++ctas) {
    const int smem = ctas == 1 ? 100 * 1024 : 60 * 1024;     // one or two CTAs fit per SM
    const int tmem_cols = 512 / ctas;                        // both CTAs of an SM must hold their TMEM at the same time
    for (int sw32 = 0; sw32 <= 1; ++sw32)
      for (int commit_every : {0, 4, 1})
        for (int N : {16, 48, 128, 256}) {
          for (int n_acc : {1, 2, 4}) {
            if (n_acc * N > tmem_cols) continue;
            if ((sw32 || commit_every) && N != 48) continue;
            float ms = 0;
            for (int rep = 0; rep < 2; ++rep) {
              cudaEventRecord(e0);
              mma_rate_kernel<<<148 * ctas, 64, smem>>>(N, n_acc, iters / n_acc, 0, tmem_cols, sw32, commit_every, d_out);
              cudaEventRecord(e1);
              cudaEventSynchronize(e1);
              cudaEventElapsedTime(&ms, e0, e1);
            }
            long long h = 0;
            if (cudaMemcpy(&h, d_out, 8, cudaMemcpyDeviceToHost) != cudaSuccess) { printf("error %s\n", cudaGetErrorString(cudaGetLastError())); return 1; }
            const double per = double(h) / double((iters / n_acc) * n_acc);
            printf("%5d %6d %5d %5d %7d | %12.1f %12.1f %10.1f\n", N, n_acc, ctas, sw32, commit_every, per, per / ctas, 1e3 * ms);
          }
        }
  }
  return 0;
}
