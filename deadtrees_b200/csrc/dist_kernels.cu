// Signed distance maps of the boundary loss on the device: the dataloader's `one_hot2dist`
// (deadtrees/loss/losses.py:159-178, called per sample at deadtrees/data/deadtreedata.py:182-185 with resolution [1, 1]).
//
//   res[k] = edt(~pos_k) * ~pos_k - (edt(pos_k) - 1) * pos_k     for every class k with at least one pixel, else 0
//
// with scipy's EXACT Euclidean distance transform: for a pixel of the mask, the distance to the nearest pixel outside it.
// Exactness makes this integer work: squared distances are integers, the reference takes their float64 square root.
// Two passes per (image, class), both over the label map (the one-hot tensor is implied):
//   pass 1 (columns)  v_in[y][x]  = vertical distance from (y, x) to the nearest pixel of class k in column x,
//                     v_out[y][x] = ... to the nearest pixel NOT of class k         (uint16, 0xFFFF = none in the column)
//   pass 2 (rows)     d2(y, x) = min over x' of (x - x')^2 + v[y][x']^2, v = v_out for pixels of class k, v_in otherwise;
//                     the row of v sits in shared memory and the search walks outwards from x until (x - x')^2 >= best.
// scipy's convention for a mask WITHOUT any outside pixel (a class covering the whole tile) is the distance to the
// virtual pixel (-1, 0); reproduced.  Values are rounded as the reference's assignment does: truncated towards zero
// into the int32 result of the dataloader path (`truncate`), or rounded to float32 (`dtype=np.float32`).
#include "common.cuh"

namespace {

constexpr int kThreads = 256;
constexpr uint16_t kNone = 0xFFFF;

// grid (ceil(W / 256), K, N): one thread = one column of one (image, class)
__global__ void dist_columns_kernel(const int64_t* __restrict__ labels, int K, int H, int W, uint16_t* __restrict__ v,
                                    int* __restrict__ flags) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int k = blockIdx.y, n = blockIdx.z;
  if (x >= W) return;
  const int64_t* lab = labels + static_cast<int64_t>(n) * H * W + x;
  const int64_t plane = static_cast<int64_t>(H) * W;
  uint16_t* v_in = v + (static_cast<int64_t>(n) * K + k) * 2 * plane + x;
  uint16_t* v_out = v_in + plane;
  int last_in = -1, last_out = -1;        // last row seen with a pixel of / not of class k
  bool any_in = false, any_out = false;
  for (int y = 0; y < H; ++y) {
    const bool in = lab[static_cast<int64_t>(y) * W] == k;
    if (in) { last_in = y; any_in = true; } else { last_out = y; any_out = true; }
    v_in[static_cast<int64_t>(y) * W] = last_in >= 0 ? static_cast<uint16_t>(y - last_in) : kNone;
    v_out[static_cast<int64_t>(y) * W] = last_out >= 0 ? static_cast<uint16_t>(y - last_out) : kNone;
  }
  last_in = last_out = -1;
  for (int y = H - 1; y >= 0; --y) {
    const bool in = lab[static_cast<int64_t>(y) * W] == k;
    if (in) last_in = y; else last_out = y;
    if (last_in >= 0) {
      const uint16_t d = static_cast<uint16_t>(last_in - y);
      if (d < v_in[static_cast<int64_t>(y) * W]) v_in[static_cast<int64_t>(y) * W] = d;
    }
    if (last_out >= 0) {
      const uint16_t d = static_cast<uint16_t>(last_out - y);
      if (d < v_out[static_cast<int64_t>(y) * W]) v_out[static_cast<int64_t>(y) * W] = d;
    }
  }
  if (any_in) atomicOr(flags + (n * K + k) * 2, 1);
  if (any_out) atomicOr(flags + (n * K + k) * 2 + 1, 1);
}

// grid (H, K, N): one block = one row of one (image, class); threads stride over x
__global__ void dist_rows_kernel(const int64_t* __restrict__ labels, int K, int H, int W, const uint16_t* __restrict__ v,
                                 const int* __restrict__ flags, int truncate, float* __restrict__ out) {
  extern __shared__ uint16_t row[];       // [2][W]: v_in, v_out of this row
  const int y = blockIdx.x, k = blockIdx.y, n = blockIdx.z;
  const int64_t plane = static_cast<int64_t>(H) * W;
  const uint16_t* v_in = v + (static_cast<int64_t>(n) * K + k) * 2 * plane + static_cast<int64_t>(y) * W;
  const uint16_t* v_out = v_in + plane;
  float* o = out + (static_cast<int64_t>(n) * K + k) * plane + static_cast<int64_t>(y) * W;
  const bool any_in = flags[(n * K + k) * 2] != 0, any_out = flags[(n * K + k) * 2 + 1] != 0;
  if (!any_in) {                          // class absent: the reference leaves the map at 0
    for (int x = threadIdx.x; x < W; x += blockDim.x) o[x] = 0.f;
    return;
  }
  for (int x = threadIdx.x; x < W; x += blockDim.x) { row[x] = v_in[x]; row[W + x] = v_out[x]; }
  __syncthreads();
  const int64_t* lab = labels + static_cast<int64_t>(n) * plane + static_cast<int64_t>(y) * W;
  for (int x = threadIdx.x; x < W; x += blockDim.x) {
    const bool in = lab[x] == k;
    double val;
    if (in && !any_out) {
      // no pixel outside the mask: scipy measures to the virtual pixel (-1, 0)
      val = -(sqrt(static_cast<double>(y + 1) * (y + 1) + static_cast<double>(x) * x) - 1.0);
    } else {
      const uint16_t* vv = in ? row + W : row;
      long long best = 1LL << 40;
      for (int d = 0; d < W; ++d) {
        const long long dd = static_cast<long long>(d) * d;
        if (dd >= best) break;
        if (x - d >= 0 && vv[x - d] != kNone) {
          const long long c = dd + static_cast<long long>(vv[x - d]) * vv[x - d];
          if (c < best) best = c;
        }
        if (d > 0 && x + d < W && vv[x + d] != kNone) {
          const long long c = dd + static_cast<long long>(vv[x + d]) * vv[x + d];
          if (c < best) best = c;
        }
      }
      const double dist = sqrt(static_cast<double>(best));
      val = in ? -(dist - 1.0) : dist;
    }
    o[x] = truncate ? static_cast<float>(static_cast<int>(val)) : static_cast<float>(val);
  }
}

}  // namespace

extern "C" {

int64_t dt_one_hot2dist_workspace(int N, int K, int H, int W) {
  if (N <= 0 || K <= 0 || H <= 0 || W <= 0) return DT_ERR_BAD_SHAPE;
  return static_cast<int64_t>(N) * K * 2 * H * W * static_cast<int64_t>(sizeof(uint16_t)) +
         static_cast<int64_t>(N) * K * 2 * static_cast<int64_t>(sizeof(int));
}

int dt_one_hot2dist(const int64_t* labels, int N, int K, int H, int W, int truncate, float* out, void* workspace,
                    int64_t workspace_bytes, dt_stream_t stream) {
  DT_ARCH_GUARD();
  DT_REQUIRE(N > 0 && N <= 65535 && K > 0 && K <= 65535 && H > 0 && H < 65535 && W > 0 && W <= 16384, DT_ERR_BAD_SHAPE,
             "dt_one_hot2dist: N=%d K=%d H=%d W=%d", N, K, H, W);
  const int64_t need = dt_one_hot2dist_workspace(N, K, H, W);
  DT_REQUIRE(workspace != nullptr && workspace_bytes >= need, DT_ERR_BAD_SHAPE,
             "dt_one_hot2dist: workspace of %lld bytes needed (dt_one_hot2dist_workspace)", static_cast<long long>(need));
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  uint16_t* v = static_cast<uint16_t*>(workspace);
  int* flags = reinterpret_cast<int*>(v + static_cast<int64_t>(N) * K * 2 * H * W);
  DT_REQUIRE(reinterpret_cast<uintptr_t>(flags) % 4 == 0, DT_ERR_BAD_ALIGN, "dt_one_hot2dist: workspace must be 4-byte aligned");
  DT_CUDA(cudaMemsetAsync(flags, 0, sizeof(int) * static_cast<size_t>(N) * K * 2, s));
  dist_columns_kernel<<<dim3((W + kThreads - 1) / kThreads, K, N), kThreads, 0, s>>>(labels, K, H, W, v, flags);
  DT_LAUNCH_CHECK();
  const int threads = W >= kThreads ? kThreads : ((W + 31) / 32 * 32);
  dist_rows_kernel<<<dim3(H, K, N), threads, 2 * W * sizeof(uint16_t), s>>>(labels, K, H, W, v, flags, truncate, out);
  DT_LAUNCH_CHECK();
  return DT_OK;
}

}  // extern "C"
