/*
 * gsrast_b200.h -- C ABI of the B200-native differentiable Gaussian-splatting rasterizer.
 *
 * This is the drop-in boundary for the ONE hot path of JiuTongBro/MultiView_Inpaint that
 * BASELINE.json names: the rasterizer behind
 *     /root/reference/gs-simp/gaussian_renderer/__init__.py:14   (import)
 *     /root/reference/gs-simp/gaussian_renderer/__init__.py:85-93 (the call)
 * In the reference that is the third-party `diff_gaussian_rasterization` extension (w-depth fork,
 * reference README.md:26; source not vendored).  Its pybind module `_C` exposes three functions --
 * rasterize_gaussians / rasterize_gaussians_backward / mark_visible -- which sit on a plain C++
 * core `CudaRasterizer::Rasterizer::{forward,backward,markVisible}` taking raw device pointers and
 * three `std::function<char*(size_t)>` buffer growers.  The entry points
 * below are that core, as `extern "C"`: plain pointers, sizes and C function pointers, no torch
 * types.  INTEGRATION.md shows the ctypes / pybind stub a maintainer of the reference would add.
 *
 * All pointers are DEVICE pointers unless the name ends in `_host`.  `stream` is a cudaStream_t
 * passed as void*.  Every function returns 0 on success, a positive cudaError_t value on a CUDA
 * failure, or one of the negative GSR_E_* codes; gsr_last_error() gives the message.
 * The library is stateless between calls except for a small per-device cache of pinned staging
 * memory; calls for one device must not run concurrently from several host threads.
 */
#ifndef GSRAST_B200_H
#define GSRAST_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GSR_VERSION 1

#define GSR_E_INVALID     (-1) /* bad argument (shape/alignment/null)                            */
#define GSR_E_PREFILTERED (-2) /* `prefiltered` set but a point was culled (the reference traps) */
#define GSR_E_ALLOC       (-3) /* a buffer-grower callback returned NULL                         */
#define GSR_E_OVERFLOW    (-4) /* num_rendered does not fit uint32 positions (GSR_MAX_INSTANCES)  */
#define GSR_MAX_INSTANCES 0xFFFF0000ll /* (tile, Gaussian) instances per view; the reference's offsets are uint32 too */

/* flags for gsr_forward */
#define GSR_FLAG_BINNING_KEY64 1u /* bin exactly like the reference: one 64-bit (tile|depth) onesweep
                                     sort (Appendix A.3/A.4).  Default (0) is the two-level scheme:
                                     depth-sort P Gaussians, emit in depth order, 2-pass tile sort.
                                     Both produce bit-identical point lists and tile ranges.          */
#define GSR_FLAG_PRECISE       2u /* blend with the oracle's exact op order, expf and IEEE division
                                     (parity build); default is ex2.approx / rcp.approx, ~2e-7 in colour */
#define GSR_FLAG_REFERENCE     4u /* ablation baseline, NOT the product path: kernels with the STRUCTURE of the
                                     public rasterizer the reference depends on (its source is absent from
                                     the reference tree): host round trip for N, one 64-bit key sort by
                                     cub::DeviceRadixSort, one thread per pixel with no culling, nine global
                                     atomicAdd per contributing (pixel, Gaussian) pair.  Same results as
                                     GSR_FLAG_PRECISE (forward bit-identical).  Implies BINNING_KEY64.     */
/* flag for gsr_backward: ADD the per-Gaussian parameter gradients (dL_dmean3D, dL_dsh, dL_dcolor,
 * dL_dopacity, dL_dcov3D, dL_dscale, dL_drot) into the output buffers instead of overwriting them, and
 * leave culled Gaussians untouched.  Used by the view-sharded multi-view step, where the outputs are
 * slices of one flat gradient arena summed over views and then all-reduced over ranks.
 * dL_dmean2D / dL_dconic stay per-view (densification reads the per-view norm, gaussian_model.py:483). */
#define GSR_FLAG_ACCUMULATE    8u
#define GSR_FLAG_ASYNC        16u /* gsr_forward: never block the host (see gsr_forward)           */
#define GSR_FLAG_SCRATCH_CLEARED 64u /* gsr_backward_blend / _views: the forward has cleared `scratch` (gsr_view_forward.backward_scratch)
                                       and nothing has written it since: skip the clear                                  */
#define GSR_FLAG_TIGHT_BINNING 32u /* two-level binning only (ignored with BINNING_KEY64 / REFERENCE): K1 stores, instead of
                                     the tile rect of the 3-sigma square (Appendix A.2 step 9), its intersection with the tiles
                                     the bounding box of the Gaussian's {alpha >= 1/255} ellipse reaches -- the very box the
                                     blend kernels test per 8x4 sub-tile.  An instance outside it has all eight sub-tile bits
                                     clear: it is sorted and staged, never evaluated.  The point list becomes the SUB-LIST of
                                     the reference's list (same order) without those instances, num_rendered counts what is
                                     left (69 % at the headline scene, less tile-sort and expansion work in proportion);
                                     colour, depth, radii and every gradient are unchanged -- bit for bit in the forward.
                                     The Python layers of this repository pass it by default (_C.DEFAULT_FLAGS); flags = 0 at
                                     the C ABI keeps the reference's literal lists.                                          */

/* Buffer grower, replaces `std::function<char*(size_t)>` of the reference core: must return a
 * device allocation of at least `bytes` bytes, 256-byte aligned, that stays alive until the
 * matching gsr_backward() has run. */
typedef char* (*gsr_alloc_fn)(void* user, size_t bytes);

/* Replaces CudaRasterizer::Rasterizer::forward (called by rasterize_gaussians, section 8b).
 * Shapes: means3D (P,3)  shs (P,M,3) or NULL  colors_precomp (P,3) or NULL  opacities (P,)
 *         scales (P,3) + rotations (P,4)  or  cov3D_precomp (P,6)
 *         viewmatrix/projmatrix (4,4) as stored by scene/cameras.py:60-62 (column-major),
 *         cam_pos (3,), background (3,)
 * Out:    out_color (3,H,W)  out_depth (1,H,W) [median depth, 15.0f where none: gen_seq.py:50]
 *         radii (P,) int32  *num_rendered_host = number of (tile, Gaussian) instances
 *
 * capacity_hint = 0: the binning buffer is sized exactly, which costs one host round trip in the
 *   middle of the pipeline (the reference does the same: cudaMemcpy of point_offsets[P-1]).
 * capacity_hint > 0 (an upper estimate of num_rendered, e.g. last frame's value + 25 %): binning
 *   and blend are queued for that capacity BEFORE N is known -- every kernel reads N on the device
 *   -- so the GPU never idles on the host; the call still returns the exact N and transparently
 *   re-bins at the exact size in the rare case N > capacity_hint.
 * GSR_FLAG_ASYNC (needs capacity_hint > 0): the call does not wait at all.  num_rendered_host must
 *   then be PINNED host memory for two int64: [0] = N, [1] = status bits (low word: prefiltered
 *   trap, high word: capacity overflow), valid after the caller synchronises the stream.  On
 *   overflow the outputs are invalid and the caller must call again with a larger hint.
 * The sorted point list always sits at offset 0 of the binning buffer.                            */
int gsr_forward(void* stream, gsr_alloc_fn geom_alloc, void* geom_user, gsr_alloc_fn binning_alloc,
                void* binning_user, gsr_alloc_fn image_alloc, void* image_user, int P, int D, int M,
                const float* background, int width, int height, const float* means3D,
                const float* shs, const float* colors_precomp, const float* opacities,
                const float* scales, float scale_modifier, const float* rotations,
                const float* cov3D_precomp, const float* viewmatrix, const float* projmatrix,
                const float* cam_pos, float tan_fovx, float tan_fovy, int prefiltered,
                float* out_color, float* out_depth, int32_t* radii, int64_t* num_rendered_host,
                int64_t capacity_hint, uint32_t flags);

/* ---- batched forward (extension; the reference renders one camera per call, gaussian_renderer/__init__.py:85-93) --
 * The forward of `n_views` views of the SAME Gaussians with ONE launch per pipeline stage for all views (K1 with the
 * views' CTAs interleaved so that the Gaussian parameters come from HBM once, the depth sort / expansion / tile sort as
 * segmented kernels, one blend launch over every view's tiles).  Always asynchronous (the GSR_FLAG_ASYNC contract of
 * gsr_forward): the caller owns every buffer -- sizes from gsr_get_layout(P, width, height, capacity, flags), 256-byte
 * aligned -- and reads result_host (PINNED int64[2]: N, status bits) after synchronising the stream; a view whose
 * N exceeds its capacity (or whose overflow bit is set) is invalid and must be redone with a larger capacity.
 * Two-level binning only (no GSR_FLAG_BINNING_KEY64 / GSR_FLAG_REFERENCE).  Results are bit-identical to gsr_forward
 * per view (same kernels: gsr_forward is the n_views == 1 case). */
typedef struct {
  const float* background;  /* (3,)                                                                       */
  const float* viewmatrix;  /* (4,4)                                                                      */
  const float* projmatrix;  /* (4,4)                                                                      */
  const float* cam_pos;     /* (3,)                                                                       */
  float tan_fovx, tan_fovy;
  int width, height;
  float* out_color;         /* (3,H,W)                                                                    */
  float* out_depth;         /* (1,H,W)                                                                    */
  int32_t* radii;           /* (P,)                                                                       */
  char* geom_buffer;        /* gsr_layout.geom_bytes                                                      */
  char* binning_buffer;     /* gsr_layout.binning_bytes for `capacity`                                    */
  char* image_buffer;       /* gsr_layout.image_bytes                                                     */
  int64_t capacity;         /* instances the binning buffer holds, 0 < capacity < GSR_MAX_INSTANCES       */
  int64_t* result_host;     /* pinned: [0] = N, [1] = status bits (low word trap, high word overflow)     */
  char* backward_scratch;   /* optional (NULL: none): the `scratch` this view's gsr_backward_blend_views call will get,
                               gsr_backward_scratch_bytes(P) bytes, 16-byte aligned.  K1 then clears it on the way (each CTA
                               the rows of its own 256 Gaussians: 144 MB of stores under a latency-bound kernel instead of
                               a memset of their own), and the backward is called with GSR_FLAG_SCRATCH_CLEARED.           */
} gsr_view_forward;
int gsr_forward_views(void* stream, int P, int D, int M, const float* means3D, const float* shs,
                      const float* colors_precomp, const float* opacities, const float* scales, float scale_modifier,
                      const float* rotations, const float* cov3D_precomp, int prefiltered,
                      const gsr_view_forward* views_host, int n_views, uint32_t flags);

/* Replaces CudaRasterizer::Rasterizer::backward (called by rasterize_gaussians_backward).
 * dL_dpix (3,H,W).  Every output is fully WRITTEN (zeros for culled Gaussians), callers need not
 * clear them:  dL_dmean2D (P,3) [.z = 0]  dL_dconic (P,4) [x,y,_,w]  dL_dopacity (P,)
 * dL_dcolor (P,3)  dL_dmean3D (P,3)  dL_dcov3D (P,6)  dL_dsh (P,M,3) or NULL when M == 0
 * dL_dscale (P,3)  dL_drot (P,4).  No gradient flows through depth.
 * dL_dconic may be NULL (the reference keeps it internal).  `scratch` is a device buffer of at
 * least gsr_backward_scratch_bytes(P) bytes (the packed per-Gaussian accumulator the blend
 * backward reduces into); the reference core needs none because it adds straight into its outputs. */
size_t gsr_backward_scratch_bytes(int P);
int gsr_backward(void* stream, int P, int D, int M, int64_t num_rendered, const float* background,
                 int width, int height, const float* means3D, const float* shs,
                 const float* colors_precomp, const float* scales, float scale_modifier,
                 const float* rotations, const float* cov3D_precomp, const float* viewmatrix,
                 const float* projmatrix, const float* cam_pos, float tan_fovx, float tan_fovy,
                 const int32_t* radii, const char* geom_buffer, const char* binning_buffer,
                 const char* image_buffer, const float* dL_dpix, float* dL_dmean2D,
                 float* dL_dconic, float* dL_dopacity, float* dL_dcolor, float* dL_dmean3D,
                 float* dL_dcov3D, float* dL_dsh, float* dL_dscale, float* dL_drot, char* scratch,
                 size_t scratch_bytes, uint32_t flags);

/* Replaces CudaRasterizer::Rasterizer::markVisible (mark_visible): present[i] = view z > 0.2 */
int gsr_mark_visible(void* stream, int P, const float* means3D, const float* viewmatrix,
                     const float* projmatrix, uint8_t* present);

/* ---- multi-view backward (extension; the reference has no counterpart: it renders one camera per
 * iteration, gs-simp/train.py:78,86) ----------------------------------------------------------------
 * gsr_backward == gsr_backward_blend (K7: fills the packed per-Gaussian accumulator `scratch`) followed by
 * the per-Gaussian chain rule (K8+K9).  For a batch of views of the SAME Gaussians the second half can run
 * once for all views: the parameters are read once, the gradients are summed over the views on chip and
 * written once (no read-modify-write of the gradient arrays, no memset), together with the densification
 * statistics.  Requires shs (not colors_precomp), scales + rotations (not cov3D_precomp) and
 * M in {1, 4, 16}; otherwise GSR_E_INVALID is returned and the caller runs gsr_backward per view.
 * GSR_FLAG_ACCUMULATE adds to the outputs / statistics instead of overwriting them. */
typedef struct {
  const int32_t* radii;      /* (P,) of that view's gsr_forward                                      */
  const char* geom_buffer;   /* that view's geometry buffer                                          */
  const char* scratch;       /* that view's accumulator, filled by gsr_backward_blend                */
  const float* viewmatrix;   /* (4,4)                                                                */
  const float* projmatrix;   /* (4,4)                                                                */
  const float* cam_pos;      /* (3,)                                                                 */
  float* dL_dmean2D;         /* optional per-view (P,3) output (NULL to skip)                        */
  float tan_fovx, tan_fovy;
  int width, height;
} gsr_view_grad;
int gsr_backward_blend(void* stream, int P, const float* background, int width, int height,
                       const char* geom_buffer, const char* binning_buffer, const char* image_buffer,
                       const float* dL_dpix, char* scratch, size_t scratch_bytes, uint32_t flags);
/* K7 of several views in one launch (one clear of all accumulators, one blend launch over every view's tiles). */
typedef struct {
  const float* background;
  int width, height;
  const char* geom_buffer;
  const char* binning_buffer;
  const char* image_buffer;
  const float* dL_dpix;      /* (3,H,W) */
  char* scratch;             /* gsr_backward_scratch_bytes(P), 16-byte aligned */
  size_t scratch_bytes;
} gsr_view_backward;
int gsr_backward_blend_views(void* stream, int P, const gsr_view_backward* views_host, int n_views, uint32_t flags);
int gsr_backward_geom_multi(void* stream, int P, int D, int M, const float* means3D, const float* shs,
                            const float* scales, float scale_modifier, const float* rotations,
                            const gsr_view_grad* views_host, int n_views, float* dL_dopacity,
                            float* dL_dmean3D, float* dL_dsh, float* dL_dscale, float* dL_drot,
                            float* grad_norm_accum, int32_t* visible_count, int32_t* max_radii,
                            uint32_t flags);
/* The same for Gaussians [g_begin, g_end) only (g_begin a multiple of 4); all pointers are those of the FULL
 * arrays.  Lets the caller pipeline chunks of the backward with chunks of the gradient all-reduce. */
int gsr_backward_geom_multi_range(void* stream, int P, int D, int M, const float* means3D, const float* shs,
                                  const float* scales, float scale_modifier, const float* rotations,
                                  const gsr_view_grad* views_host, int n_views, float* dL_dopacity,
                                  float* dL_dmean3D, float* dL_dsh, float* dL_dscale, float* dL_drot,
                                  float* grad_norm_accum, int32_t* visible_count, int32_t* max_radii,
                                  uint32_t flags, int g_begin, int g_end);

/* In-switch all-reduce (NVLink SHARP / NVLS) of one symmetric buffer that every rank has mapped through the
 * same MULTICAST address `multicast_ptr` (torch.distributed._symmetric_memory: handle.multicast_ptr).
 * Three segments, given as byte offsets into the buffer: float32 SUM (n_f32 a multiple of 4, 16-byte
 * aligned offset), int32 SUM, int32 MAX.  Rank `rank` of `world` reduces and re-broadcasts slice `rank` of
 * every segment with multimem.ld_reduce / multimem.st.  The caller must order the launch between two
 * cross-rank barriers on `stream`.  blocks <= 0 picks a default grid.
 * Row sparsity (optional, sparse_rows = 0 disables it): floats [sparse_first_f32, sparse_first_f32 +
 * sparse_rows * sparse_row_f32) of the float segment form a row-major matrix whose row g is known to be zero
 * on every rank whenever the SUM over ranks of element g of the int32 SUM segment is zero (the (P,M,3) SH
 * gradient vs the visibility count): those rows are skipped.  Needs sparse_row_f32 % 4 == 0. */
int gsr_nvls_all_reduce(void* stream, void* multicast_ptr, size_t off_f32, size_t n_f32, size_t off_add_s32,
                        size_t n_add_s32, size_t off_max_s32, size_t n_max_s32, int rank, int world, int blocks,
                        size_t sparse_first_f32, size_t sparse_rows, int sparse_row_f32);

/* The same collective for an arbitrary set of ranges of the symmetric buffer -- used to all-reduce the arena in
 * Gaussian-range chunks so that the reduction of chunk c overlaps the per-Gaussian backward of chunk c + 1.
 * All offsets are bytes from the multicast base, 16-byte aligned for float data; dense lengths are multiples of
 * 4 floats.  `rows` rows of `row_f32` floats starting at rows_off are reduced row-sparsely against the int32
 * counts starting at rows_count_off (row g is skipped when the SUM over ranks of count g is zero; that count
 * must lie inside the add_s32 range so that it is itself reduced); row widths other than 12 / 48 floats are
 * reduced densely. */
typedef struct {
  size_t dense_off[6];
  size_t dense_n_f32[6];
  int n_dense;
  size_t rows_off, rows;
  int row_f32;
  size_t rows_count_off;
  size_t add_s32_off, n_add_s32;
  size_t max_s32_off, n_max_s32;
} gsr_nvls_plan;
int gsr_nvls_all_reduce_plan(void* stream, void* multicast_ptr, const gsr_nvls_plan* plan, int rank, int world, int blocks);
/* The same plan between exactly TWO ranks without the switch: rank r (0 or 1) owns half of every segment, loads the peer's
 * elements over NVLink (`peer_ptr`: the peer's replica of the arena as mapped into this process, e.g. symmetric memory's
 * buffer_ptrs[1 - r] + offset; `local_ptr`: this rank's replica), adds its own and stores the sum into both.  Same barriers
 * around the launch as gsr_nvls_all_reduce_plan.  At two GPUs multimem moves the same bytes over the links at a lower rate. */
int gsr_p2p_all_reduce_plan(void* stream, void* local_ptr, void* peer_ptr, const gsr_nvls_plan* plan, int rank, int blocks);

/* Densification statistics of one view, fused (the reference does this in torch after backward():
 * gs-simp/scene/gaussian_model.py:482-484 `xyz_gradient_accum[vis] += norm(viewspace.grad[vis, :2])`,
 * `denom[vis] += 1`, and gs-simp/train.py:115 `max_radii2D[vis] = max(max_radii2D[vis], radii[vis])`,
 * vis = radii > 0).  dL_dmean2D is the (P,3) output of gsr_backward.  Any of the three accumulators
 * may be NULL. */
int gsr_accumulate_view_stats(void* stream, int P, const int32_t* radii, const float* dL_dmean2D,
                              float* grad_norm_accum, int32_t* visible_count, int32_t* max_radii);

/* ---- the training step either side of the rasterizer ---------------------
 * Fused photometric loss, replaces the torch expression of gs-simp/train.py:91-92 (also sds_train.py:117-118,
 * inpaint_rec.py:118-123):
 *     Ll1 = l1_loss(image, gt)                                   utils/loss_utils.py:17-18
 *     loss = (1 - lambda_dssim) * Ll1 + lambda_dssim * (1 - ssim(image, gt))    utils/loss_utils.py:33-62
 * image, gt (C,H,W) float32.  Forward writes out_loss3 = {Ll1, ssim, loss} (device, 3 floats) and keeps three
 * (C,H,W) derivative maps in `temp` (>= gsr_loss_temp_bytes bytes), which the backward consumes:
 * dL_dimage (C,H,W) = dL_dloss * d loss / d image; dL_dloss is a DEVICE scalar (NULL = 1.0, what loss.backward()
 * seeds).  dL_dimage is what gsr_backward takes as dL_dpix. */
size_t gsr_loss_temp_bytes(int C, int H, int W);
int gsr_loss_l1_ssim_forward(void* stream, int C, int H, int W, const float* image, const float* gt,
                             float lambda_dssim, float* out_loss3, char* temp, size_t temp_bytes);
int gsr_loss_l1_ssim_backward(void* stream, int C, int H, int W, const float* image, const float* gt,
                              float lambda_dssim, const float* dL_dloss, const char* temp, size_t temp_bytes,
                              float* dL_dimage);

/* Parameter activations of GaussianModel's getters (gs-simp/scene/gaussian_model.py:33-41,95-115), one kernel:
 * scales = exp(raw_scales) (P,3), rotations = normalize(raw_rotations) (P,4; v / max(|v|, 1e-12)),
 * opacities = sigmoid(raw_opacities) (P,).  Any of the three raw pointers may be NULL (that group is skipped).
 * Backward runs IN PLACE: g_* hold dL/d(activated) on entry (e.g. slices of the gradient arena after
 * gsr_backward_geom_multi) and dL/d(raw) on exit -- what autograd would hand to Adam for _scaling, _rotation,
 * _opacity.  rotations / g_rotations must be 16-byte aligned. */
int gsr_activate_forward(void* stream, int P, const float* raw_scales, const float* raw_rotations,
                         const float* raw_opacities, float* scales, float* rotations, float* opacities);
int gsr_activate_backward(void* stream, int P, const float* raw_scales, const float* raw_rotations,
                          const float* raw_opacities, float* g_scales, float* g_rotations, float* g_opacities);

/* Adam over up to 8 parameter segments in ONE launch; replaces `gaussians.optimizer.step()` (train.py:127) for the
 * optimizer of gaussian_model.py:154-165 = torch.optim.Adam(groups, lr=0.0, eps=1e-15): betas (0.9, 0.999), no
 * weight decay, no amsgrad, dense (every parameter moves every step, as in the reference).  `step` is the
 * 1-based step count after the increment (torch's state['step']).  A row-structured segment (row_len > 0) uses
 * `lr` for floats [0,row_split) of every row and `lr_rest` for [row_split,row_len): the (P,M,3) SH tensor holds
 * f_dc (lr = feature_lr) and f_rest (lr = feature_lr / 20) side by side, row_len = 3M, row_split = 3. */
typedef struct {
  float* param;
  const float* grad;
  float* exp_avg;
  float* exp_avg_sq;
  uint64_t n;        /* floats in the segment */
  double lr;         /* doubles, like the python floats torch keeps them as: 1 - beta2 and lr / (1 - beta1^t) are  */
  double lr_rest;    /* formed in double and rounded to fp32 once (a float 0.999 would put 1.3e-5 into 1 - beta2) */
  int row_len;       /* 0 = flat segment */
  int row_split;
} gsr_adam_segment;
int gsr_adam_step(void* stream, const gsr_adam_segment* segs_host, int n_segs, int64_t step, double beta1,
                  double beta2, double eps);

/* ---- densification re-pack of the flat arenas ------------------------------------------
 * What gs-simp/scene/gaussian_model.py does to the six parameter tensors and their Adam state when the model changes
 * size -- densify_and_prune :467-480 = cat_tensors_to_optimizer :385-406 (clone, split) + _prune_optimizer :346-363
 * (split parents, final prune), and prune_points :365-383 on its own -- as ONE row gather over all segments:
 *     dst[i*row_f32 + c] = (zero_new && i >= n_keep_state) ? 0 : src[src_row[i]*row_f32 + c],  i < n_dst, c < row_f32
 * src_row: device int32[n_dst], every entry in [0, n_src) (composed on the host side of the binding from the clone /
 * split / prune masks); n_keep_state: the leading destination rows that keep their optimizer state (surviving
 * original rows; cloned and split rows follow them and start from zero moments, as torch.zeros_like appends them).
 * Segments with zero_new = 0 are parameter slices, zero_new = 1 moment slices.  src and dst must not overlap;
 * slices whose row_f32 is a multiple of 4 must be 16-byte aligned.  Pure data movement: bit-exact. */
#define GSR_GATHER_MAX_SEGS 16
typedef struct {
  const float* src;
  float* dst;
  int row_f32;       /* floats per row (3, 3M, 1, 3, 4 for xyz / features / opacity / scaling / rotation) */
  int zero_new;      /* 1: rows >= n_keep_state are zero-filled instead of copied */
} gsr_gather_segment;
int gsr_gather_rows(void* stream, int64_t n_dst, int64_t n_src, int64_t n_keep_state, const int32_t* src_row,
                    const gsr_gather_segment* segs_host, int n_segs);

/* ---- image output of the inference loops ---------------------------------------------
 * The tensor half of `torchvision.utils.save_image(t, path)` as called after render() in gs-simp/render.py:36-39,
 * render_depth.py:39, gen_seq.py:45-55, vis_render.py:48-51: image (C,H,W) float32 with C = 3, or C = 1 (repeated to
 * three channels, as make_grid does) -> rgb8 (H,W,3) uint8 with u8 = trunc(clamp(x * 255 + 0.5, 0, 255)) (two fp32
 * roundings, as torch's mul(255).add_(0.5)).  `affine` (optional DEVICE float[2] = {lo, inv_range}) first maps
 * x -> (x - lo) * inv_range: scene/helpers.py:159-162 normalize_0_to_1 without a host round trip for min / max.
 * Bit-exact integer output.  The 3-bytes-per-pixel result is what an asynchronous PNG writer copies to the host. */
int gsr_quantize_rgb8(void* stream, int C, int H, int W, const float* image, const float* affine, uint8_t* rgb8);

/* ---- neighbour distances --------------------------------------------------------
 * Replaces `simple_knn._C.distCUDA2(points) -> Tensor[P]` (third-party simple-knn extension, source not in the
 * reference tree), imported at gs-simp/scene/gaussian_model.py:20 and called at :134, :546, :623:
 * mean_dist2[i] = (d1 + d2 + d3) / 3, the mean SQUARED distance from point i to its three nearest OTHER
 * points (coincident points count, with distance 0; fewer than four points leave FLT_MAX terms, as upstream).
 * points (P,3) float32, mean_dist2 (P,), temp >= gsr_knn_temp_bytes(P) bytes, all device memory.
 * Exact (not approximate) nearest neighbours: bit-identical to a brute-force scan with
 * d = fma(dz,dz, fma(dy,dy, dx*dx)). */
size_t gsr_knn_temp_bytes(int P);
int gsr_knn3_mean_dist2(void* stream, int P, const float* points, float* mean_dist2, char* temp, size_t temp_bytes);

/* ---- building blocks, exported so that parity tests can drive each stage through the C ABI ---- */

/* Stable LSD onesweep radix sort of (key,value) pairs on key bits [0,end_bit), 8-bit digits.
 * keys_in/vals_in are preserved.  `temp` must hold gsr_sort_temp_bytes(...) bytes. */
size_t gsr_sort_temp_bytes(int64_t n, int key_bytes /*4 or 8*/, int end_bit);
int gsr_sort_pairs_u64(void* stream, int64_t n, const uint64_t* keys_in, const uint32_t* vals_in,
                       uint64_t* keys_out, uint32_t* vals_out, int end_bit, char* temp,
                       size_t temp_bytes);
int gsr_sort_pairs_u32(void* stream, int64_t n, const uint32_t* keys_in, const uint32_t* vals_in,
                       uint32_t* keys_out, uint32_t* vals_out, int end_bit, char* temp,
                       size_t temp_bytes);

/* Single-pass (decoupled look-back) inclusive prefix sum of uint32; if `gather` is non-NULL the
 * input element i is in[gather[i]].  `temp` must hold gsr_scan_temp_bytes(n) bytes. */
size_t gsr_scan_temp_bytes(int64_t n);
int gsr_inclusive_scan_u32(void* stream, int64_t n, const uint32_t* in, const uint32_t* gather,
                           uint32_t* out, char* temp, size_t temp_bytes);

/* Byte offsets of the arrays inside the three opaque scratch buffers (for stage-level parity
 * tests and debuggers; the layout is otherwise private).  Unused entries are set to (size_t)-1. */
typedef struct {
  size_t rec;           /* geom: float4[3P]: {x,y,conic.x,conic.y} {conic.z,opacity,hx,hy} {r,g,b,power_cut} */
  size_t depths;        /* geom: float[P]  view-space z                                     */
  size_t clamped;       /* geom: uint8[P]  bit c set <=> colour channel c was clamped to 0  */
  size_t tiles_touched; /* geom: uint32[P]                                                  */
  size_t point_offsets; /* geom: uint32[P] inclusive scan (index order for KEY64, depth order otherwise) */
  size_t order;         /* geom: uint32[P] Gaussian ids by ascending depth bits (two-level only)        */
  size_t geom_bytes;
  size_t final_T;       /* image: float[H*W]   */
  size_t n_contrib;     /* image: uint32[H*W]  */
  size_t ranges;        /* image: uint2[G]     */
  size_t image_bytes;
  size_t point_list;    /* binning: uint32[N] sorted Gaussian ids (front of the buffer)    */
  size_t binning_bytes;
} gsr_layout;
int gsr_get_layout(int P, int width, int height, int64_t num_rendered, uint32_t flags, gsr_layout* out);

/* ---- measurement hooks (bench.py) ---- */
#define GSR_NUM_STAGES 10
/* stage ids: 0 preprocess(K1) 1 depth sort 2 scan(K2) 3 duplicate(K3) 4 tile/key sort(K4) 5 ranges(K5)
 *            6 blend forward(K6) 7 accumulator clear 8 blend backward(K7) 9 per-Gaussian backward(K8+K9) */
/* When enabled, gsr_forward/gsr_backward bracket every stage with CUDA events on the launching stream. */
void gsr_profile_enable(int on);
/* Synchronises the recorded events, ADDS each stage's elapsed milliseconds and launch-bracket count
 * into ms[GSR_NUM_STAGES] / counts[GSR_NUM_STAGES], and clears the recording. */
int gsr_profile_collect(double* ms_host, int64_t* counts_host);
/* Number of CUDA kernels this library has launched in this process (monotonic). */
uint64_t gsr_kernel_launches(void);
/* Diagnostics for the parity tests: evaluates on the device, for each x[i], the approximate units the
 * default (non-PRECISE) blend uses: out[2i] = rcp.approx.ftz(x[i]), out[2i+1] = ex2.approx.ftz(x[i]).
 * The branch-free blend backward relies on rcp.approx(1) == 1 and ex2.approx(0) == 1 being exact. */
int gsr_debug_approx_units(const float* x_dev, int n, float* out_dev, void* stream);

const char* gsr_last_error(void);
int gsr_version(void);

/* Experiment switches (tools/ only; the product path never calls this).  knob 0: warp ranking of the radix sort
 * (0 match_any, 1 ballots, 2 shared-memory atomicOr = default); knob 1: how the expansion counts its instances: 2 (default) per-CTA digit histograms for the
 * tile sort + tile ranges read off the sorted keys, 1 per-instance tile_count atomics (the previous scheme, identical results), 0 nothing
 * (results are then WRONG: timing experiments only); knob 2: resident CTAs per SM K1 is compiled for (4 / 5 / 6);
 * knob 3: reduction of the blend backward's nine sums: 1 (default) tensor-core contraction (mma.sync m16n8k8 tf32, hi + lo
 * split), 0 the round-1 shuffle butterfly -- same sums up to fp32 rounding, kept for A/B timing and the cross-check test;
 * knob 5: 1 (default) gsr_gather_rows moves its wide rows (SH segments) by cp.async.bulk, 0 by thread loads / stores only
 * (bit-identical: pure data movement). */
int gsr_debug_set(int knob, int value);

#ifdef __cplusplus
}
#endif
#endif
