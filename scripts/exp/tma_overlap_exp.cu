// Experiment: does cuTensorMapEncodeTiled accept OVERLAPPING strides (sliding-window / im2col view), and does the
// TMA load return the expected bytes?  View of a padded NHWC4 bf16 image (pixel = 8 B):
//   dims  (32 elems, Wo, Ho, 7 filter rows, N)   strides (2 B) 16 B, 2*Wp*8 B, Wp*8 B, Hp*Wp*8 B
// element (e, wo, ho, r, n) = x[n][2*ho + r][2*wo + e/4][e%4]  -> a 64-byte row per (output pixel, filter row).
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <vector>
#include "../../deadtrees_b200/csrc/common.cuh"
void dt_set_error(const char*, ...) {}
int dt_check_device() { return 0; }

__global__ void k(const __grid_constant__ CUtensorMap tm, uint16_t* out, int r, int wo0, int ho0) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  __syncthreads();
  if (threadIdx.x == 0) {
    mbar_arrive_expect_tx(&bar, 128 * 64);
    asm volatile("cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3,%4,%5,%6,%7}], [%2];"
                 :: "r"(smem_u32(smem)), "l"(reinterpret_cast<uint64_t>(&tm)), "r"(smem_u32(&bar)), "r"(0), "r"(wo0), "r"(ho0), "r"(r), "r"(0) : "memory");
  }
  mbar_wait(&bar, 0);
  for (int i = threadIdx.x; i < 128 * 32; i += blockDim.x) out[i] = reinterpret_cast<uint16_t*>(smem)[i];
}

int main() {
  const int T = 64, Hp = T + 6, Wp = T + 8, N = 2, Wo = T / 2, Ho = T / 2;
  std::vector<uint16_t> h(size_t(N) * Hp * Wp * 4);
  for (size_t i = 0; i < h.size(); ++i) h[i] = uint16_t(i * 2654435761u >> 16);
  uint16_t *d, *o;
  cudaMalloc(&d, h.size() * 2); cudaMalloc(&o, 128 * 32 * 2);
  cudaMemcpy(d, h.data(), h.size() * 2, cudaMemcpyHostToDevice);
  void* fn = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
  typedef CUresult (*F)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  CUtensorMap tm;
  cuuint64_t dims[5] = {32, (cuuint64_t)Wo, (cuuint64_t)Ho, 7, (cuuint64_t)N};
  cuuint64_t str[4] = {16, (cuuint64_t)2 * Wp * 8, (cuuint64_t)Wp * 8, (cuuint64_t)Hp * Wp * 8};
  cuuint32_t box[5] = {32, 32, 4, 1, 1}, es[5] = {1, 1, 1, 1, 1};
  for (int sw = 0; sw < 2; ++sw) {
    CUresult r = ((F)fn)(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, d, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         sw ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode (swizzle %s): CUresult %d\n", sw ? "64B" : "none", (int)r);
    if (r != CUDA_SUCCESS) continue;
    const int fr = 3, wo0 = 0, ho0 = 8;
    k<<<1, 128, 128 * 64 + 1024>>>(tm, o, fr, wo0, ho0);
    cudaError_t e = cudaDeviceSynchronize();
    printf("kernel: %s\n", cudaGetErrorString(e));
    if (e != cudaSuccess) return 0;
    std::vector<uint16_t> got(128 * 32);
    cudaMemcpy(got.data(), o, got.size() * 2, cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int row = 0; row < 128; ++row) {
      const int wo = wo0 + row % 32, ho = ho0 + row / 32;
      for (int e2 = 0; e2 < 32; ++e2) {
        const uint16_t want = h[((size_t(0) * Hp + 2 * ho + fr) * Wp + 2 * wo) * 4 + e2];
        int chunk = e2 / 8, within = e2 % 8;
        int pchunk = sw ? (chunk ^ ((row >> 1) & 3)) : chunk;   // SWIZZLE_64B: 16B chunk ^= address bits [7,9)
        if (got[row * 32 + pchunk * 8 + within] != want) ++bad;
      }
    }
    printf("mismatches: %d / %d\n", bad, 128 * 32);
  }
  return 0;
}
