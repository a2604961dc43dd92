// common.cuh -- shared declarations of the sm_100a rasterizer kernels.
//
// Floating-point contract (DESIGN.md "FP contract"): everything that feeds an integer output
// (radii, tile rects, depth keys) or the per-Gaussian geometry state is written with the explicit
// round-to-nearest intrinsics below, which nvcc never contracts or reassociates, in exactly the
// order the CPU oracle uses.  That is what makes radii / keys / sorted lists / ranges bit-exact.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/gsrast_b200.h"

namespace gsr {

constexpr int TILE_X = 16;
constexpr int TILE_Y = 16;
constexpr int TILE_PIXELS = TILE_X * TILE_Y;

__device__ __forceinline__ float MUL(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float ADD(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float SUB(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float DIV(float a, float b) { return __fdiv_rn(a, b); }
__device__ __forceinline__ float FMA(float a, float b, float c) { return __fmaf_rn(a, b, c); }
__device__ __forceinline__ float SQRT(float a) { return __fsqrt_rn(a); }

// SURVEY Appendix A.1 (== gs-simp/utils/sh_utils.py:26-43)
#define GSR_SH_C0 0.28209479177387814f
#define GSR_SH_C1 0.4886025119029199f
#define GSR_SH_C2_0 1.0925484305920792f
#define GSR_SH_C2_1 -1.0925484305920792f
#define GSR_SH_C2_2 0.31539156525252005f
#define GSR_SH_C2_3 -1.0925484305920792f
#define GSR_SH_C2_4 0.5462742152960396f
#define GSR_SH_C3_0 -0.5900435899266435f
#define GSR_SH_C3_1 2.890611442640554f
#define GSR_SH_C3_2 -0.4570457994644658f
#define GSR_SH_C3_3 0.3731763325901154f
#define GSR_SH_C3_4 -0.4570457994644658f
#define GSR_SH_C3_5 1.445305721320277f
#define GSR_SH_C3_6 -0.5900435899266435f

// Slack added to the per-Gaussian power cut-off ln(255*opacity) used for culling: anything the
// kernels skip has alpha < 1/255 with a margin ~1e3 x larger than the fp32 evaluation error of
// `power`, so skipping is exactly equivalent to the reference's `if (alpha < 1/255) continue`.
#define GSR_POWER_SLACK 0.05f

// Uniform per launch.  The matrices stay on the device (the reference hands them over as CUDA
// tensors, scene/cameras.py:60-63); each CTA copies the 35 floats into shared memory once, so no
// host round trip is needed to launch.
struct Camera {
  const float* view;    // (4,4) column-major W2C
  const float* proj;    // (4,4) column-major P*W2C
  const float* campos;  // (3,)
  float focal_x, focal_y, tan_fovx, tan_fovy, scale_modifier;
  int W, H, grid_x, grid_y;
};
// s_cam[0..15] view, [16..31] proj, [32..34] campos
__device__ __forceinline__ void load_camera(const Camera& cam, float* s_cam) {
  const int t = threadIdx.x;
  if (t < 16) s_cam[t] = __ldg(cam.view + t);
  else if (t < 32) s_cam[t] = __ldg(cam.proj + (t - 16));
  else if (t < 35) s_cam[t] = __ldg(cam.campos + (t - 32));
  __syncthreads();
}

// A.2 step 9 (getRect), explicit fp32 sequence shared by K1 and K3
__device__ __forceinline__ void get_rect(float px, float py, int max_radius, int gx, int gy,
                                         int& x0, int& y0, int& x1, int& y1) {
  const float r = (float)max_radius;
  x0 = min(gx, max(0, __float2int_rz(DIV(SUB(px, r), 16.0f))));
  y0 = min(gy, max(0, __float2int_rz(DIV(SUB(py, r), 16.0f))));
  x1 = min(gx, max(0, __float2int_rz(DIV(SUB(ADD(ADD(px, r), 16.0f), 1.0f), 16.0f))));
  y1 = min(gy, max(0, __float2int_rz(DIV(SUB(ADD(ADD(py, r), 16.0f), 1.0f), 16.0f))));
}

inline size_t align_up(size_t x, size_t a = 256) { return (x + a - 1) / a * a; }
inline int cdiv(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

// A.4 getHigherMsb (host)
inline uint32_t higher_msb(uint32_t n) {
  uint32_t msb = sizeof(n) * 4, step = msb;
  while (step > 1) {
    step /= 2;
    if (n >> msb) msb += step; else msb -= step;
  }
  if (n >> msb) msb++;
  return msb;
}
inline int ceil_log2(uint32_t n) {
  int b = 0;
  while ((1ull << b) < n) b++;
  return b;
}

// Views (independent segments) one batched launch covers: every front-end stage of a multi-view step is ONE launch
// over all views of the rank (blockIdx.y = view, or view-interleaved CTAs where the views share inputs).
constexpr int GSR_MAX_BATCH = 8;

// ---- private layout of the opaque scratch buffers (public view: gsr_layout) -------------------
struct GeomLayout {
  size_t rec, depths, clamped, tiles_touched, point_offsets;
  size_t depth_keys, order, depth_keys_alt, order_alt;    // two-level binning only (sorted result
                                                          // lands back in depth_keys / order)
  size_t rects;                                           // two-level only: ushort4[P] tile rects
  size_t status;                                          // int32[8]: trap, overflow, N (u64), work-list length, V
  size_t scan_temp, scan_temp_bytes;                      // ticket + look-back words of the scans (right after status)
  size_t temp, temp_bytes;                                // depth-sort temp (two-level binning only)
  size_t bytes;
};
struct ImageLayout {
  size_t final_T, n_contrib, ranges, tile_count, bytes;
};
struct BinningLayout {
  size_t point_list;       // uint32[N]  (final, sorted)
  size_t vals_alt;         // uint32[N]
  size_t keys_a, keys_b;   // KeyT[N] each (u32 tile ids, or u64 tile|depth)
  size_t temp, temp_bytes; // sort temp
  size_t big_items;        // uint4[bin_big_capacity(N)] work list of large rects (two-level only)
  size_t bytes;
};
GeomLayout geom_layout(int P, uint32_t flags);
ImageLayout image_layout(int W, int H);
BinningLayout binning_layout(int64_t N, int W, int H, uint32_t flags);

// experiment switches behind gsr_debug_set (never changed by the product path)
extern int g_rs_rank_mode;         // scan_sort.cu: 0 match_any, 1 ballots, 2 shared atomicOr (default)
extern int g_pre_min_blocks;       // preprocess.cu: resident CTAs per SM K1 is compiled for (4, 5 or 6 = default)
extern bool g_bwd_mma;             // blend_backward.cu: true (default) = tensor-core contraction of the nine sums, false = shuffle butterfly
extern int g_gather_bulk;          // densify.cu: 1 (default) wide rows of gsr_gather_rows move by cp.async.bulk, 0 thread path only
extern int g_bin_count_mode;       // binning.cu: 2 (default) per-CTA digit histograms + ranges off the sorted keys, 1 per-instance tile_count
                                   // atomics + tile_prepare (the previous scheme), 0 nothing (WRONG results: timing only)

// ---- host launchers (one per translation unit) --------------------------------------------------
void set_error(const char* msg);
void count_launch(int n = 1);  // bumps the counter behind gsr_kernel_launches()

// K1 outputs of one view.  status: int32[8] = {trap, overflow, N lo, N hi, work-list length, V (visible), -, -}
struct PreView {
  Camera cam;
  int32_t* radii;
  float4* rec;
  float* depths;
  uint8_t* clamped;
  uint32_t* tiles_touched;
  uint32_t* depth_keys;  // two-level binning only (else NULL)
  ushort4* rects;        // two-level binning only (else NULL)
  int32_t* status;
  float4* gacc;          // optional: the blend backward's 48-byte accumulators, cleared here (else NULL)
};
// One launch for all views: CTA b works on view b % nv, Gaussians [256 (b / nv), +256) -- the CTAs of the views
// of one Gaussian chunk are co-resident, so the (P,M,3) SH rows and the other parameters come from HBM once and
// from L2 for the other views.
cudaError_t launch_preprocess(cudaStream_t s, int P, int D, int M, const float* means3D,
                              const float* scales, const float* rotations, const float* opacities,
                              const float* shs, const float* cov3D_precomp,
                              const float* colors_precomp, float scale_modifier, int prefiltered,
                              const PreView* views, int nv, bool tight = false);
cudaError_t launch_mark_visible(cudaStream_t s, int P, const float* means3D, const float* view,
                                uint8_t* present);

size_t scan_temp_bytes(int64_t n);
cudaError_t launch_inclusive_scan(cudaStream_t s, int64_t n, const uint32_t* in,
                                  const uint32_t* gather, uint32_t* out, char* temp);
size_t sort_temp_bytes(int64_t n, int key_bytes, int end_bit);
// vals_in may be NULL: values are then the element indices 0..n-1.
// n_dev may be NULL; otherwise the element count is min(*n_dev, n) read on the device and `n` only
// sizes the grid and the temp storage (capacity).
// have_bases: the per-digit exclusive bases are already in temp (sort_hist_ptr) -- no histogram pass.
cudaError_t launch_sort_pairs_u32(cudaStream_t s, int64_t n, const uint32_t* n_dev, const uint32_t* keys_in,
                                  const uint32_t* vals_in, uint32_t* keys_out, uint32_t* vals_out,
                                  uint32_t* keys_alt, uint32_t* vals_alt, int end_bit, char* temp,
                                  bool have_bases = false);
inline uint32_t* sort_hist_ptr(char* temp) { return reinterpret_cast<uint32_t*>(temp); }
// One independent array of a batched (segmented) sort: blockIdx.y = segment.
template <typename KeyT>
struct SortSeg {
  int64_t n;                      // capacity (sizes the grid and the temp storage); element count if n_dev == NULL
  const uint32_t* n_dev;          // element count on the device: min(*n_dev, n)
  const uint32_t* n_dev_compact;  // compact sorts: the number of keys != ~0, read by the passes after the first
  const KeyT* keys_in;
  const uint32_t* vals_in;        // NULL: values are the element indices
  KeyT* keys_out;
  uint32_t* vals_out;
  KeyT* keys_alt;
  uint32_t* vals_alt;
  char* temp;                     // sort_temp_bytes(n, sizeof(KeyT), end_bit)
};
// compact: the first pass drops keys equal to 0xFFFFFFFF (culled Gaussians) -- the output holds *n_dev_compact pairs
// temp_zeroed: the caller has already cleared every segment's temp (launch_zero_regions) -- no memsets here
cudaError_t launch_sort_pairs_u32_batched(cudaStream_t s, const SortSeg<uint32_t>* segs, int nseg, int end_bit,
                                          int have_bases, bool compact, bool temp_zeroed);   // have_bases: 0 histogram pass,
                                          // 1 exclusive bases in temp, 2 digit COUNTS in temp (scanned here)
cudaError_t launch_sort_pairs_u64(cudaStream_t s, int64_t n, const uint32_t* n_dev, const uint64_t* keys_in,
                                  const uint32_t* vals_in, uint64_t* keys_out, uint32_t* vals_out,
                                  uint64_t* keys_alt, uint32_t* vals_alt, int end_bit, char* temp);

// `cap` = capacity of the key/value arrays: instances beyond it are dropped and status[1] is set
// (the host then re-runs the binning with a larger buffer).
cudaError_t launch_duplicate_key64(cudaStream_t s, int P, const float4* rec, const float* depths,
                                   const uint32_t* offsets, const int32_t* radii, int grid_x,
                                   int grid_y, uint64_t* keys, uint32_t* vals, int64_t cap, int32_t* status);
cudaError_t launch_duplicate_tiles(cudaStream_t s, int P, const uint32_t* order, const float4* rec,
                                   const uint32_t* offsets, const int32_t* radii, int grid_x,
                                   int grid_y, uint32_t* tile_keys, uint32_t* vals, int64_t cap, int32_t* status);
// Two-level scheme: fused scan (offsets in depth order) + load-balanced expansion + per-tile counts, one launch for
// all views (blockIdx.y = view).  Then ranges + the tile sort's digit bases.
struct BinView {
  const uint32_t* order;           // Gaussian ids in depth order, V = status[5] entries
  const ushort4* rects;            // [P] tile rects by Gaussian id
  uint32_t* offsets;               // [P] out: inclusive instance offsets in depth order (first V entries)
  uint32_t* tile_keys;             // [cap] out
  uint32_t* vals;                  // [cap] out
  uint32_t cap;
  uint32_t* tile_count;            // [G], zeroed by the caller (count mode 1)
  uint32_t* hist;                  // the tile sort's [pass][256] digit counters, zeroed by the caller (count mode 2)
  int end_bit;                     // tile-id width
  int gx;
  unsigned long long* lb_status;   // look-back words, zeroed by the caller (scan_temp + 256)
  uint32_t* ticket;                // zeroed by the caller (scan_temp)
  int32_t* status;                 // the view's status block; status[4] (work-list length) zeroed by the caller
  uint4* big_items;
  uint32_t big_cap;
  int P;                           // capacity of `order`
};
struct PrepView {
  int G;
  const uint32_t* tile_count;
  uint2* ranges;
  int passes, end_bit;
  uint32_t* hist;
};
struct RangeView {
  const uint32_t* keys;    // sorted tile ids
  const uint32_t* n_dev;   // N on the device
  int64_t cap;
  uint2* ranges;           // [G], every entry written
  int G;
};
cudaError_t launch_tile_ranges_views(cudaStream_t s, const RangeView* views, int nv);
cudaError_t launch_bin_expand(cudaStream_t s, const BinView* views, int nv, int count_mode);
int64_t bin_big_capacity(int64_t cap);
cudaError_t launch_tile_prepare(cudaStream_t s, const PrepView* views, int nv);
// Zeroes up to 4 * GSR_MAX_BATCH regions (4-byte aligned, sizes multiples of 4) with ONE launch.
struct ZeroRegion { void* ptr; size_t bytes; };
cudaError_t launch_zero_regions(cudaStream_t s, const ZeroRegion* regions, int n);
// N = min(*n_dev, cap) when n_dev is given, else cap
cudaError_t launch_tile_ranges_u64(cudaStream_t s, int64_t cap, const uint32_t* n_dev, const uint64_t* keys, int G, uint2* ranges);
cudaError_t launch_tile_ranges_u32(cudaStream_t s, int64_t cap, const uint32_t* n_dev, const uint32_t* keys, int G, uint2* ranges);

// K6 / K7 of all views of a batched step in one launch (blockIdx.y = view)
struct BlendFwdView {
  int W, H, grid_x, G;   // grid_x / G are filled in by the launcher
  const uint2* ranges;
  const uint32_t* point_list;
  const float4* rec;
  const float* depths;
  const float* bg;
  float* out_color;
  float* out_depth;
  float* final_T;
  uint32_t* n_contrib;
};
struct BlendBwdView {
  int W, H, grid_x, G;
  const uint2* ranges;
  const uint32_t* point_list;
  const float4* rec;
  const float* bg;
  const float* final_T;
  const uint32_t* n_contrib;
  const float* dL_dpix;
  float* gacc;  // [P][12]
};
cudaError_t launch_blend_forward(cudaStream_t s, const BlendFwdView* views, int nv, bool precise);
cudaError_t launch_blend_backward(cudaStream_t s, const BlendBwdView* views, int nv, bool precise);
cudaError_t launch_debug_approx_units(cudaStream_t s, const float* x, int n, float* out /*[n][2]*/);
cudaError_t launch_geom_backward(cudaStream_t s, int P, int D, int M, const float* means3D,
                                 const int32_t* radii, const float* shs, const uint8_t* clamped,
                                 const float* scales, const float* rotations,
                                 const float* cov3D_precomp, const float* colors_precomp,
                                 const Camera& cam, const float4* rec, const float* gacc,
                                 float* dL_dmean2D, float* dL_dconic, float* dL_dopacity,
                                 float* dL_dcolor, float* dL_dmean3D, float* dL_dcov3D,
                                 float* dL_dsh, float* dL_dscale, float* dL_drot, bool accumulate);

// Reference-structure stand-in (refstruct.cu, GSR_FLAG_REFERENCE): cub sort + thread-per-pixel blend
size_t ref_sort_temp_bytes(int64_t n, int end_bit);
cudaError_t launch_ref_sort_pairs(cudaStream_t s, int64_t n, const uint64_t* keys_in, const uint32_t* vals_in,
                                  uint64_t* keys_out, uint32_t* vals_out, int end_bit, char* temp, size_t temp_bytes);
cudaError_t launch_ref_blend_forward(cudaStream_t s, int W, int H, const uint2* ranges, const uint32_t* point_list,
                                     const float4* rec, const float* depths, const float* bg, float* out_color,
                                     float* out_depth, float* final_T, uint32_t* n_contrib);
cudaError_t launch_ref_blend_backward(cudaStream_t s, int W, int H, const uint2* ranges, const uint32_t* point_list,
                                      const float4* rec, const float* bg, const float* final_T,
                                      const uint32_t* n_contrib, const float* dL_dpix, float* gacc);

// One view's inputs to the multi-view per-Gaussian backward (geom_backward_multi.cu)
struct ViewGrad {
  const int32_t* radii;
  const uint8_t* clamped;
  const float4* rec;
  const float* gacc;
  const float* view;
  const float* proj;
  const float* campos;
  float* dL_dmean2D;  // optional (P,3)
  float focal_x, focal_y, tan_fovx, tan_fovy;
  int W, H;
};
bool geom_backward_multi_supported(int M);
cudaError_t launch_geom_backward_multi(cudaStream_t s, int P, int D, int M, const float* means3D, const float* shs,
                                       const float* scales, const float* rotations, float scale_modifier,
                                       const ViewGrad* views, int n_views, float* dL_dopacity,
                                       float* dL_dmean3D, float* dL_dsh, float* dL_dscale, float* dL_drot,
                                       float* grad_norm_accum, int32_t* visible_count, int32_t* max_radii,
                                       bool accumulate);
cudaError_t launch_nvls_allreduce(cudaStream_t s, char* mc, size_t off_f32, size_t n_f32, size_t off_add_s32,
                                  size_t n_add_s32, size_t off_max_s32, size_t n_max_s32, int rank, int world,
                                  int blocks, size_t sparse_first_f32, size_t sparse_rows, int sparse_row_f32);
cudaError_t launch_nvls_allreduce_plan(cudaStream_t s, char* mc, const gsr_nvls_plan& plan, int rank, int world, int blocks);
cudaError_t launch_p2p_allreduce_plan(cudaStream_t s, char* local, char* peer, const gsr_nvls_plan& in, int rank, int blocks);
size_t knn_temp_bytes(int P);
cudaError_t launch_knn3_mean_dist2(cudaStream_t s, int P, const float* points, float* mean_dist2, char* temp);
cudaError_t launch_view_stats(cudaStream_t s, int P, const int32_t* radii, const float* dL_dmean2D,
                              float* grad_norm_accum, int32_t* visible_count, int32_t* max_radii);

size_t loss_temp_bytes(int C, int H, int W);
cudaError_t launch_loss_forward(cudaStream_t s, int C, int H, int W, const float* img, const float* gt, float lambda,
                                float* out3, char* temp);
cudaError_t launch_loss_backward(cudaStream_t s, int C, int H, int W, const float* img, const float* gt, float lambda,
                                 const float* dL_dloss, const char* temp, float* dL_dimg);
cudaError_t launch_activate_forward(cudaStream_t s, int P, const float* raw_scale, const float* raw_rot,
                                    const float* raw_opacity, float* scale, float* rot, float* opacity);
cudaError_t launch_activate_backward(cudaStream_t s, int P, const float* raw_scale, const float* raw_rot,
                                     const float* raw_opacity, float* g_scale, float* g_rot, float* g_opacity);
cudaError_t launch_quantize_rgb8(cudaStream_t s, int C, int H, int W, const float* in, const float* affine, uint8_t* out);
cudaError_t launch_gather_rows(cudaStream_t s, long long n_dst, long long n_keep_state, const int* src_row,
                               const gsr_gather_segment* segs, int n_segs);
cudaError_t launch_adam(cudaStream_t s, const gsr_adam_segment* segs, int n_segs, int64_t step, double beta1,
                        double beta2, double eps);

}  // namespace gsr
