/*
 * gsplat_oracle.h -- CPU restatement (plain C) of the differentiable Gaussian-splatting rasterizer
 * that /root/reference/gs-simp/gaussian_renderer/__init__.py:85-93 calls through
 * `diff_gaussian_rasterization` (the "w-depth" fork, README.md:26 of the reference).
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product package (multiview_inpaint_b200/) may include,
 * link or call this.  Allowed users: tests/, __graft_entry__.smoke(), bench.py's cpu_baseline /
 * `--impl reference` leg.
 *
 * PARITY STATUS: **parity unpinned** for the pipeline as a whole.  The rasterizer's source is a
 * third-party dependency (JonathonLuiten/diff-gaussian-rasterization-w-depth, no version pinned,
 * reference README.md:26, gs-simp/environment.yml:16) that is absent from /root/reference and from
 * this image, and the reference ships no tests or golden vectors for it.
 * The algorithm restated here is the published one written down in SURVEY.md Appendix A; the
 * sub-steps that DO exist in the reference tree are pinned against it by tests/golden/:
 *   - SH -> RGB        gs-simp/utils/sh_utils.py:57-112  + clamp gaussian_renderer/__init__.py:78
 *   - cov3D            gs-simp/utils/general_utils.py:66-112, scene/gaussian_model.py:27-31
 *   - camera matrices  gs-simp/utils/graphics_utils.py:38-70, scene/cameras.py:54-64
 *   - depth sentinel   gs-simp/gen_seq.py:50, vis_render.py:45  (15.0f)
 *
 * Floating-point contract: every expression is written as an explicit sequence of IEEE-754
 * binary32 mul / add / sub / fma / div / sqrt (compile with -ffp-contract=off, no fast-math), so
 * that the CUDA kernels, which spell the same sequence with __fmul_rn/__fadd_rn/__fmaf_rn, agree
 * BIT-FOR-BIT on everything up to the blend (radii, tile rects, depths, conics, colours, keys,
 * sorted lists, tile ranges).  The blend uses expf(), whose CUDA and glibc versions differ in the
 * last ulp, so colour/depth/grads are compared with the tolerances in BASELINE.json.
 */
#ifndef GSPLAT_ORACLE_H
#define GSPLAT_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GSO_BLOCK_X 16
#define GSO_BLOCK_Y 16

/* binary-search MSB used to size the sort's key range. */
uint32_t gso_higher_msb(uint32_t n);

/* Appendix A.2 (K1).  All arrays caller-allocated.  shs / colors_precomp / scales+rotations /
 * cov3D_precomp may be NULL exactly as in the reference call (gaussian_renderer/__init__.py:59-82).
 * Returns 0, or -1 if `prefiltered` was set and a point was culled (the reference traps). */
int gso_preprocess(int P, int D, int M,
                   const float* means3D, const float* scales, float scale_modifier,
                   const float* rotations, const float* opacities, const float* shs,
                   const float* cov3D_precomp, const float* colors_precomp,
                   const float* viewmatrix, const float* projmatrix, const float* campos,
                   int W, int H, float tan_fovx, float tan_fovy, int prefiltered,
                   /* out */
                   int32_t* radii, float* means2D /*2P*/, float* depths, float* cov3D /*6P*/,
                   float* rgb /*3P*/, float* conic_opacity /*4P*/, uint32_t* tiles_touched,
                   uint8_t* clamped /*3P*/);

/* Appendix A.3 (K2): inclusive prefix sum; returns the total (num_rendered) as 64-bit. */
int64_t gso_inclusive_scan(int P, const uint32_t* tiles_touched, uint32_t* point_offsets);

/* Appendix A.3 (K3): emit (tile<<32 | depth_bits, gaussian_id) for every touched tile. */
void gso_duplicate_with_keys(int P, const float* means2D, const float* depths,
                             const uint32_t* point_offsets, const int32_t* radii,
                             int W, int H, uint64_t* keys_unsorted, uint32_t* values_unsorted);

/* Appendix A.4 (K4): stable LSD radix sort of pairs on key bits [0, end_bit). */
void gso_sort_pairs(int64_t N, const uint64_t* keys_in, const uint32_t* vals_in,
                    uint64_t* keys_out, uint32_t* vals_out, int end_bit);

/* Appendix A.4 (K5): ranges[2*tile] = first, ranges[2*tile+1] = last+1; untouched tiles (0,0). */
void gso_identify_tile_ranges(int64_t N, const uint64_t* keys_sorted, int num_tiles,
                              uint32_t* ranges /*2G*/);

/* Appendix A.5 (K6): front-to-back alpha blend + median depth (sentinel 15.0f). */
void gso_blend_forward(int W, int H, const uint32_t* ranges, const uint32_t* point_list,
                       const float* means2D, const float* rgb, const float* depths,
                       const float* conic_opacity, const float* bg,
                       float* out_color /*3HW*/, float* out_depth /*HW*/,
                       float* final_T /*HW*/, uint32_t* n_contrib /*HW*/);

/* Appendix A.6 (K7): back-to-front gradient of the blend.  Outputs are ACCUMULATED into
 * (callers zero them): dL_dmean2D (3P, .x .y written), dL_dconic (4P: x,y,_,w), dL_dopacity (P),
 * dL_dcolors (3P). */
void gso_blend_backward(int W, int H, const uint32_t* ranges, const uint32_t* point_list,
                        const float* means2D, const float* rgb, const float* conic_opacity,
                        const float* bg, const float* final_T, const uint32_t* n_contrib,
                        const float* dL_dpixels /*3HW*/,
                        float* dL_dmean2D, float* dL_dconic, float* dL_dopacity,
                        float* dL_dcolors);

/* Tile-strided variants (tiles tile_start, tile_start+tile_step, ...): bench.py times a bounded
 * sample of the blend with these; pixels of unselected tiles are left untouched. */
void gso_blend_forward_tiles(int W, int H, const uint32_t* ranges, const uint32_t* point_list,
                             const float* means2D, const float* rgb, const float* depths,
                             const float* conic_opacity, const float* bg, float* out_color,
                             float* out_depth, float* final_T, uint32_t* n_contrib, int tile_start,
                             int tile_step);
void gso_blend_backward_tiles(int W, int H, const uint32_t* ranges, const uint32_t* point_list,
                              const float* means2D, const float* rgb, const float* conic_opacity,
                              const float* bg, const float* final_T, const uint32_t* n_contrib,
                              const float* dL_dpixels, float* dL_dmean2D, float* dL_dconic,
                              float* dL_dopacity, float* dL_dcolors, int tile_start, int tile_step);

/* Appendix A.7 (K8 + K9): per-Gaussian chain rule.  dL_dmeans3D (3P), dL_dcov3D (6P),
 * dL_dsh (3MP), dL_dscales (3P), dL_drots (4P) are WRITTEN for visible Gaussians (callers zero). */
void gso_preprocess_backward(int P, int D, int M,
                             const float* means3D, const int32_t* radii, const float* shs,
                             const uint8_t* clamped, const float* scales, const float* rotations,
                             float scale_modifier, const float* cov3D,
                             const float* viewmatrix, const float* projmatrix, const float* campos,
                             int W, int H, float tan_fovx, float tan_fovy,
                             const float* dL_dmean2D /*3P*/, const float* dL_dconic /*4P*/,
                             const float* dL_dcolors /*3P*/,
                             float* dL_dmeans3D, float* dL_dcov3D, float* dL_dsh,
                             float* dL_dscales, float* dL_drots);

/* markVisible (K10): view-space z > 0.2. */
void gso_mark_visible(int P, const float* means3D, const float* viewmatrix, uint8_t* present);

/* simple_knn.distCUDA2: mean squared distance to the three nearest other points (brute force, exact).
 * PARITY UNPINNED (third-party source absent).  queries may be NULL (= all points). */
void gso_knn3_mean_dist2(int P, const float* pts, const int32_t* queries, int nq, float* out);

int gso_num_threads(void);
void gso_set_num_threads(int n);

#ifdef __cplusplus
}
#endif
#endif
