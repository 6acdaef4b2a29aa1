// Training-time augmentation + normalisation of a whole batch on the device: the dataloader's `train_transform`
// (deadtrees/data/deadtreedata.py:132-146) and the channel / class handling of `transform()` (:156-189) for one batch.
//
//   A.OneOf([HorizontalFlip, VerticalFlip], p=0.5) -> A.RandomRotate90(p=0.5) -> A.RandomBrightnessContrast(p=0.5,
//   brightness_limit=0.2, contrast_limit=0.15, brightness_by_max=False) -> A.Normalize(mean, std) -> ToTensorV2,
//   then image[0:in_channels], mask.long(), lu.long(), and mask > 1 -> 1 for two-class training.
//
// The random draws stay on the host (one tiny parameter row per sample); given the draws everything is deterministic
// byte / index work:
//   geometry    out(i, j) = in(f(r(i, j))): np.rot90 by `rot` quarter turns (counter-clockwise) applied after the flip
//   brightness  lut[v] = uint8(clip(float32(v) * alpha + float32(beta * mean(img)), 0, 255)), the uint8 look-up table of
//               albumentations' brightness / contrast adjustment; mean(img) = mean over EVERY byte of the sample's
//               image (all C channels, exact: integer sum / count in float64)
//   normalise   (float32(lut[v]) - offset[c]) * scale[c], offset = mean * 255, scale = 1 / (std * 255) in float32
// Pass 1 sums the bytes of each sample (uint64, exact); pass 2 does everything else in one read of the batch.
// HBM traffic per sample: H*W*(C + 2) bytes read twice-ish (C bytes in pass 1), H*W*(4*out_channels + 16) written.
#include "common.cuh"

namespace {

constexpr int kThreads = 256;

// grid (chunks, N): exact byte sum of sample n, one 64-bit atomic per block
__global__ void image_byte_sum_kernel(const uint8_t* __restrict__ img, int64_t bytes, int vectorised,
                                      unsigned long long* __restrict__ sums) {
  const uint8_t* p = img + static_cast<int64_t>(blockIdx.y) * bytes;
  unsigned int acc = 0;                                   // <= 255 * 16 bytes * iterations: blocks see < 2^24 bytes each
  const int64_t vecs = vectorised ? bytes / 16 : 0;       // vectorised: base 16-byte aligned and 16 | bytes (host checks)
  const uint4* p4 = reinterpret_cast<const uint4*>(p);
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < vecs;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const uint4 v = p4[i];
    // sum of the four bytes of a word: __vsadu4(x, 0) = |b0| + |b1| + |b2| + |b3|
    acc += __vsadu4(v.x, 0u) + __vsadu4(v.y, 0u) + __vsadu4(v.z, 0u) + __vsadu4(v.w, 0u);
  }
  if (blockIdx.x == 0)
    for (int64_t i = vecs * 16 + threadIdx.x; i < bytes; i += blockDim.x) acc += p[i];
  __shared__ unsigned long long sh[kThreads / 32];
  unsigned long long w = acc;
  for (int o = 16; o > 0; o >>= 1) w += __shfl_xor_sync(0xffffffffu, w, o);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = w;
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned long long t = 0;
    for (int i = 0; i < kThreads / 32; ++i) t += sh[i];
    if (t) atomicAdd(&sums[blockIdx.y], t);
  }
}

struct NormConst { float offset[4], scale[4]; };

// source pixel of output pixel (i, j): quarter turns first (inverse of np.rot90), then the flip
__device__ __forceinline__ void source_pixel(int i, int j, int H, int W, int flip, int rot, int& si, int& sj) {
  // np.rot90(m, k) on an (H, W) array (square here for odd k): k=1: out[i][j] = m[j][W-1-i]; k=2: m[H-1-i][W-1-j]; k=3: m[H-1-j][i]
  int a, b;
  switch (rot & 3) {
    case 1: a = j; b = W - 1 - i; break;
    case 2: a = H - 1 - i; b = W - 1 - j; break;
    case 3: a = H - 1 - j; b = i; break;
    default: a = i; b = j; break;
  }
  if (flip == 1) b = W - 1 - b;            // HorizontalFlip: out[i][j] = in[i][W-1-j]
  else if (flip == 2) a = H - 1 - a;       // VerticalFlip:   out[i][j] = in[H-1-i][j]
  si = a; sj = b;
}

// grid (ceil(W/32), ceil(H/8), N), block (32, 8): one thread = one output pixel
template <int C>
__global__ void train_transform_kernel(const uint8_t* __restrict__ img, const uint8_t* __restrict__ mask,
                                       const uint8_t* __restrict__ lu, int H, int W, int out_channels,
                                       const int32_t* __restrict__ geom, const double* __restrict__ bc,
                                       const unsigned long long* __restrict__ sums, NormConst nc, int merge_classes,
                                       float* __restrict__ out_img, int64_t* __restrict__ out_mask,
                                       int64_t* __restrict__ out_lu) {
  const int j = blockIdx.x * 32 + threadIdx.x, i = blockIdx.y * 8 + threadIdx.y, n = blockIdx.z;
  if (i >= H || j >= W) return;
  const int flip = geom[2 * n], rot = geom[2 * n + 1];
  const float alpha = static_cast<float>(bc[2 * n]);        // `lut *= alpha` on the float32 table
  const double beta = bc[2 * n + 1];
  int si, sj;
  source_pixel(i, j, H, W, flip, rot, si, sj);
  const int64_t HW = static_cast<int64_t>(H) * W;
  const int64_t src = static_cast<int64_t>(n) * HW + static_cast<int64_t>(si) * W + sj;
  const int64_t dst = static_cast<int64_t>(i) * W + j;
  // additive brightness term: beta * mean(img) in float64 as numpy computes it, then rounded to the float32 of the table
  float add = 0.f;
  if (beta != 0.0) {
    const double mean = static_cast<double>(sums[n]) / (static_cast<double>(HW) * C);
    add = static_cast<float>(beta * mean);
  }
  uint8_t px[C];
  if constexpr (C == 4) {
    const uint32_t w = reinterpret_cast<const uint32_t*>(img)[src];
    px[0] = w & 0xFF; px[1] = (w >> 8) & 0xFF; px[2] = (w >> 16) & 0xFF; px[3] = w >> 24;
  } else {
#pragma unroll
    for (int c = 0; c < C; ++c) px[c] = img[src * C + c];
  }
#pragma unroll
  for (int c = 0; c < C; ++c) {
    if (c >= out_channels) break;
    float v = static_cast<float>(px[c]);
    if (alpha != 1.f) v = __fmul_rn(v, alpha);
    if (beta != 0.0) v = __fadd_rn(v, add);
    v = fminf(fmaxf(v, 0.f), 255.f);
    const float q = static_cast<float>(static_cast<uint8_t>(v));            // astype(uint8): truncation
    out_img[(static_cast<int64_t>(n) * out_channels + c) * HW + dst] = __fmul_rn(__fsub_rn(q, nc.offset[c]), nc.scale[c]);
  }
  if (mask != nullptr) {
    int64_t m = mask[src];
    if (merge_classes && m > 1) m = 1;
    out_mask[static_cast<int64_t>(n) * HW + dst] = m;
  }
  if (lu != nullptr) out_lu[static_cast<int64_t>(n) * HW + dst] = lu[src];
}

}  // namespace

extern "C" {

int dt_train_transform(const uint8_t* images, const uint8_t* masks, const uint8_t* lus, int N, int H, int W, int C,
                       int out_channels, const int32_t* geom, const double* bc, const float* offset, const float* scale,
                       int merge_classes, uint64_t* sums, float* out_img, int64_t* out_mask, int64_t* out_lu,
                       dt_stream_t stream) {
  DT_ARCH_GUARD();
  DT_REQUIRE(N > 0 && N <= 65535 && H > 0 && W > 0 && (C == 1 || C == 3 || C == 4) && out_channels >= 1 && out_channels <= C,
             DT_ERR_BAD_SHAPE, "dt_train_transform: N=%d H=%d W=%d C=%d out_channels=%d (C in {1,3,4})", N, H, W, C, out_channels);
  DT_REQUIRE(H == W, DT_ERR_BAD_SHAPE, "dt_train_transform: quarter turns inside a batch need square tiles (H=%d W=%d)", H, W);
  DT_REQUIRE(images && geom && bc && offset && scale && sums && out_img && (!masks || out_mask) && (!lus || out_lu),
             DT_ERR_BAD_SHAPE, "dt_train_transform: null pointer");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int64_t bytes = static_cast<int64_t>(H) * W * C;
  DT_CUDA(cudaMemsetAsync(sums, 0, sizeof(uint64_t) * N, s));
  const bool aligned = (bytes % 16 == 0) && (reinterpret_cast<uintptr_t>(images) % 16 == 0);
  DT_REQUIRE(aligned || C != 4 || reinterpret_cast<uintptr_t>(images) % 4 == 0, DT_ERR_BAD_SHAPE,
             "dt_train_transform: 4-channel images must be 4-byte aligned");
  if (aligned) {
    int chunks = static_cast<int>((bytes / 16 + kThreads * 8 - 1) / (kThreads * 8));
    chunks = chunks < 1 ? 1 : (chunks > 256 ? 256 : chunks);
    image_byte_sum_kernel<<<dim3(chunks, N), kThreads, 0, s>>>(images, bytes, 1, reinterpret_cast<unsigned long long*>(sums));
  } else {   // odd sample sizes: one block per sample walks the bytes
    image_byte_sum_kernel<<<dim3(1, N), kThreads, 0, s>>>(images, bytes, 0, reinterpret_cast<unsigned long long*>(sums));
  }
  DT_LAUNCH_CHECK();
  NormConst nc;
  for (int c = 0; c < 4; ++c) { nc.offset[c] = c < out_channels ? offset[c] : 0.f; nc.scale[c] = c < out_channels ? scale[c] : 0.f; }
  dim3 grid((W + 31) / 32, (H + 7) / 8, N), block(32, 8);
  unsigned long long* su = reinterpret_cast<unsigned long long*>(sums);
  switch (C) {
    case 1: train_transform_kernel<1><<<grid, block, 0, s>>>(images, masks, lus, H, W, out_channels, geom, bc, su, nc, merge_classes, out_img, out_mask, out_lu); break;
    case 3: train_transform_kernel<3><<<grid, block, 0, s>>>(images, masks, lus, H, W, out_channels, geom, bc, su, nc, merge_classes, out_img, out_mask, out_lu); break;
    default: train_transform_kernel<4><<<grid, block, 0, s>>>(images, masks, lus, H, W, out_channels, geom, bc, su, nc, merge_classes, out_img, out_mask, out_lu); break;
  }
  DT_LAUNCH_CHECK();
  return DT_OK;
}

}  // extern "C"
