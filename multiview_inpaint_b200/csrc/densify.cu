// densify.cu -- the arena side of densification (the part DESIGN.md section 8 listed as open):
// one row-gather launch re-packs the parameter arena AND both Adam moment arenas after clone / split / prune.
//
// Replaces, in gs-simp/scene/gaussian_model.py, what densify_and_prune (:467-480) does to the six parameter tensors
// and their optimizer state through densification_postfix (:408-425) -> cat_tensors_to_optimizer (:385-406) twice
// (clone, split) and prune_points (:365-383) -> _prune_optimizer (:346-363) twice (split parents, final prune): four
// rounds of `torch.cat` / boolean-mask indexing over 6 tensors x {param, exp_avg, exp_avg_sq} = 72 torch kernels that
// each copy the whole model.  The host side (densify.py) composes the four steps into ONE index map
// src_row[new row] = old row; here every destination row is then written exactly once:
//     dst[i, :] = src[src_row[i], :]                       for parameters
//     dst[i, :] = i < n_keep_state ? src[src_row[i], :] : 0 for the Adam moments (cloned / split rows start from zero
//                                                           state: cat_tensors_to_optimizer appends zeros_like)
// for up to GSR_GATHER_MAX_SEGS segments (5 slices x 3 arenas) in the same launch.
//
// HBM-bound: (4 + 2 * 12 * row_f32 summed over segments) bytes per destination row (59 floats x 3 arenas at M = 16:
// 1.42 KB).  A CTA owns 64 consecutive destination rows (their source rows staged once in shared memory); within a
// segment the CTA's threads walk the 64 x row_f32 destination floats linearly, so stores are fully coalesced and
// loads are contiguous per source row (src_row is monotone inside each of the three blocks kept / clones / children,
// so neighbouring rows mostly share sectors).  Rows whose length is a multiple of 4 floats (SH rows of M = 4 / 16,
// rotations) move as float4.  Pure data movement: bit-exact.
#include "common.cuh"

namespace gsr {

namespace {

constexpr int GATHER_ROWS = 64;

struct GatherArgs {
  const float* src[GSR_GATHER_MAX_SEGS];
  float* dst[GSR_GATHER_MAX_SEGS];
  int row_f32[GSR_GATHER_MAX_SEGS];
  int zero_new[GSR_GATHER_MAX_SEGS];
  int n_segs;
};

__global__ void __launch_bounds__(256)
gather_rows_kernel(GatherArgs a, long long n_dst, long long n_keep_state, const int* __restrict__ src_row) {
  __shared__ int s_src[GATHER_ROWS];
  const long long n_tiles = (n_dst + GATHER_ROWS - 1) / GATHER_ROWS;
  for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const long long row0 = tile * GATHER_ROWS;
    const int rows = (int)min((long long)GATHER_ROWS, n_dst - row0);
    __syncthreads();  // WAR on s_src (previous tile)
    if (threadIdx.x < rows) s_src[threadIdx.x] = __ldg(src_row + row0 + threadIdx.x);
    __syncthreads();
    // rows of this tile that keep their optimizer state: [0, keep)
    const int keep = (int)max(0LL, min((long long)rows, n_keep_state - row0));
    for (int s = 0; s < a.n_segs; s++) {
      const int L = a.row_f32[s];
      const float* __restrict__ src = a.src[s];
      float* __restrict__ dst = a.dst[s];
      const int live = a.zero_new[s] ? keep : rows;   // rows below `live` copy, the rest are zero-filled
      if ((L & 3) == 0) {
        const int Lv = L >> 2;
        const float4* __restrict__ src4 = reinterpret_cast<const float4*>(src);
        float4* __restrict__ dst4 = reinterpret_cast<float4*>(dst) + (size_t)row0 * Lv;
        const int n = rows * Lv;
        for (int e = threadIdx.x; e < n; e += blockDim.x) {
          const int r = e / Lv, c = e - r * Lv;
          float4 v = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
          if (r < live) v = __ldg(src4 + (size_t)s_src[r] * Lv + c);
          dst4[e] = v;
        }
      } else {
        float* __restrict__ d = dst + (size_t)row0 * L;
        const int n = rows * L;
        for (int e = threadIdx.x; e < n; e += blockDim.x) {
          const int r = e / L, c = e - r * L;
          float v = 0.0f;
          if (r < live) v = __ldg(src + (size_t)s_src[r] * L + c);
          d[e] = v;
        }
      }
    }
  }
}

}  // namespace

cudaError_t launch_gather_rows(cudaStream_t stream, long long n_dst, long long n_keep_state, const int* src_row,
                               const gsr_gather_segment* segs, int n_segs) {
  GatherArgs a;
  a.n_segs = n_segs;
  for (int s = 0; s < n_segs; s++) {
    a.src[s] = segs[s].src;
    a.dst[s] = segs[s].dst;
    a.row_f32[s] = segs[s].row_f32;
    a.zero_new[s] = segs[s].zero_new;
  }
  for (int s = n_segs; s < GSR_GATHER_MAX_SEGS; s++) {
    a.src[s] = nullptr; a.dst[s] = nullptr; a.row_f32[s] = 0; a.zero_new[s] = 0;
  }
  const long long n_tiles = (n_dst + GATHER_ROWS - 1) / GATHER_ROWS;
  const int grid = (int)min(n_tiles, (long long)148 * 8);   // 8 resident CTAs of 256 threads per SM, grid-stride beyond
  gather_rows_kernel<<<grid, 256, 0, stream>>>(a, n_dst, n_keep_state, src_row);
  count_launch();
  return cudaGetLastError();
}

}  // namespace gsr
