// binning.cu -- K3 (duplicateWithKeys) and K5 (identifyTileRanges).
//
// Restates SURVEY.md Appendix A.3 / A.4.  Two flavours of K3:
//   * key64  : the reference's layout -- one (tile << 32 | depth_bits, gaussian) pair per touched
//              tile, emitted in Gaussian-index order; a 64-bit sort on [0, 32+bit) follows.
//   * tiles  : two-level scheme -- Gaussians were already stably sorted by depth bits, thread k
//              handles the k-th nearest Gaussian and emits (tile, gaussian) in that order; a stable
//              sort on the tile id alone then yields EXACTLY the list the 64-bit sort would:
//              within a tile, depth ascending, ties by ascending Gaussian index.
// The rect is recomputed from (mean2D, radius) with the same explicit fp32 sequence as K1.
#include "common.cuh"

namespace gsr {

__global__ void __launch_bounds__(256)
duplicate_key64_kernel(int P, const float4* __restrict__ rec, const float* __restrict__ depths,
                       const uint32_t* __restrict__ offsets, const int32_t* __restrict__ radii,
                       int gx, int gy, uint64_t* __restrict__ keys, uint32_t* __restrict__ vals,
                       uint32_t cap, int32_t* __restrict__ status) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P) return;
  if (i == P - 1 && offsets[i] > cap) atomicExch(status + 1, 1);  // capacity overflow
  const int r = radii[i];
  if (r <= 0) return;
  uint32_t off = (i == 0) ? 0u : offsets[i - 1];
  const float4 q0 = __ldg(rec + 3 * (size_t)i);
  int x0, y0, x1, y1;
  get_rect(q0.x, q0.y, r, gx, gy, x0, y0, x1, y1);
  const uint64_t dbits = __float_as_uint(depths[i]);
  for (int y = y0; y < y1; y++)
    for (int x = x0; x < x1; x++) {
      if (off < cap) {
        keys[off] = ((uint64_t)(uint32_t)(y * gx + x) << 32) | dbits;
        vals[off] = (uint32_t)i;
      }
      off++;
    }
}

// Thread k handles Gaussian order[k] (depth rank k).  offsets[] is the inclusive scan of
// tiles_touched in that same order.
__global__ void __launch_bounds__(256)
duplicate_tiles_kernel(int P, const uint32_t* __restrict__ order, const float4* __restrict__ rec,
                       const uint32_t* __restrict__ offsets, const int32_t* __restrict__ radii,
                       int gx, int gy, uint32_t* __restrict__ tile_keys, uint32_t* __restrict__ vals,
                       uint32_t cap, int32_t* __restrict__ status) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= P) return;
  if (k == P - 1 && offsets[k] > cap) atomicExch(status + 1, 1);  // capacity overflow
  const uint32_t i = order[k];
  const int r = radii[i];
  if (r <= 0) return;  // culled Gaussians carry key 0xFFFFFFFF and sit at the end of `order`
  uint32_t off = (k == 0) ? 0u : offsets[k - 1];
  const float4 q0 = __ldg(rec + 3 * (size_t)i);
  int x0, y0, x1, y1;
  get_rect(q0.x, q0.y, r, gx, gy, x0, y0, x1, y1);
  for (int y = y0; y < y1; y++)
    for (int x = x0; x < x1; x++) {
      if (off < cap) {
        tile_keys[off] = (uint32_t)(y * gx + x);
        vals[off] = i;
      }
      off++;
    }
}

template <typename KeyT>
__device__ __forceinline__ uint32_t tile_of(KeyT k);
template <> __device__ __forceinline__ uint32_t tile_of<uint64_t>(uint64_t k) { return (uint32_t)(k >> 32); }
template <> __device__ __forceinline__ uint32_t tile_of<uint32_t>(uint32_t k) { return k; }

template <typename KeyT>
__global__ void __launch_bounds__(256)
tile_ranges_kernel(int64_t cap, const uint32_t* __restrict__ n_dev, const KeyT* __restrict__ keys,
                   uint2* __restrict__ ranges) {
  const int64_t N = n_dev ? min((int64_t)*n_dev, cap) : cap;
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= N) return;
  const uint32_t cur = tile_of<KeyT>(keys[idx]);
  if (idx == 0) {
    ranges[cur].x = 0;
  } else {
    const uint32_t prev = tile_of<KeyT>(keys[idx - 1]);
    if (cur != prev) {
      ranges[prev].y = (uint32_t)idx;
      ranges[cur].x = (uint32_t)idx;
    }
  }
  if (idx == N - 1) ranges[cur].y = (uint32_t)N;
}

cudaError_t launch_duplicate_key64(cudaStream_t s, int P, const float4* rec, const float* depths,
                                   const uint32_t* offsets, const int32_t* radii, int grid_x,
                                   int grid_y, uint64_t* keys, uint32_t* vals, int64_t cap, int32_t* status) {
  if (P == 0) return cudaSuccess;
  duplicate_key64_kernel<<<cdiv(P, 256), 256, 0, s>>>(P, rec, depths, offsets, radii, grid_x, grid_y, keys, vals,
                                                      (uint32_t)cap, status);
  count_launch();
  return cudaGetLastError();
}
cudaError_t launch_duplicate_tiles(cudaStream_t s, int P, const uint32_t* order, const float4* rec,
                                   const uint32_t* offsets, const int32_t* radii, int grid_x,
                                   int grid_y, uint32_t* tile_keys, uint32_t* vals, int64_t cap, int32_t* status) {
  if (P == 0) return cudaSuccess;
  duplicate_tiles_kernel<<<cdiv(P, 256), 256, 0, s>>>(P, order, rec, offsets, radii, grid_x, grid_y, tile_keys, vals,
                                                      (uint32_t)cap, status);
  count_launch();
  return cudaGetLastError();
}
cudaError_t launch_tile_ranges_u64(cudaStream_t s, int64_t N, const uint32_t* n_dev, const uint64_t* keys, int G, uint2* ranges) {
  cudaError_t e = cudaMemsetAsync(ranges, 0, (size_t)G * sizeof(uint2), s);
  if (e != cudaSuccess || N == 0) return e;
  tile_ranges_kernel<uint64_t><<<cdiv(N, 256), 256, 0, s>>>(N, n_dev, keys, ranges);
  count_launch();
  return cudaGetLastError();
}
cudaError_t launch_tile_ranges_u32(cudaStream_t s, int64_t N, const uint32_t* n_dev, const uint32_t* keys, int G, uint2* ranges) {
  cudaError_t e = cudaMemsetAsync(ranges, 0, (size_t)G * sizeof(uint2), s);
  if (e != cudaSuccess || N == 0) return e;
  tile_ranges_kernel<uint32_t><<<cdiv(N, 256), 256, 0, s>>>(N, n_dev, keys, ranges);
  count_launch();
  return cudaGetLastError();
}

}  // namespace gsr
