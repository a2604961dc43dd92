// blend_forward.cu -- K6: per-16x16-tile front-to-back alpha blend with median depth.
//
// Restates SURVEY.md Appendix A.5.  The depth sentinel 15.0f and the (1,H,W) depth output are
// pinned by the reference callers gs-simp/gen_seq.py:50, vis_render.py:45, render_depth.py:37.
//
// B200 design (differs from the public kernel, same per-pixel recurrence and order):
//   * one CTA = one tile = 8 warps; warp w owns an 8x4 pixel sub-tile, so a warp is a compact
//     screen-space box and can be culled as a unit;
//   * the tile's list is staged 256 entries at a time into shared memory from the 48-byte blend
//     records written by K1 (3 x LDG.128 per entry, 2 sectors);
//   * while staging, each thread tests its Gaussian's alpha>=1/255 bounding box against the eight
//     sub-tiles and the CTA publishes, per warp, a 256-bit "touches my sub-tile" mask (ballots);
//     a warp then walks only the set bits (warp-uniform loop, no divergence on the skip);
//   * the exp is only evaluated for pairs whose power is above the Gaussian's cut-off.
//   Both culls are conservative (GSR_POWER_SLACK): they only drop pairs the reference would have
//   dropped with `alpha < 1/255`, so n_contrib / final_T / colour / depth are unchanged.
//   * saturated pixels: per-lane `done`, warp-level __all_sync to stop walking, CTA-level
//     __syncthreads_and to stop staging (the reference's __syncthreads_count early exit).
// The kernel is instruction-issue bound (ncu: issue slots ~89 % busy, DRAM ~1 %), so the inner
// loop is written for instruction count: 2 x LDS.128 + 5 FP + EX2 before the alpha test, the
// colour float4 only for contributing lanes (see blend_common.cuh).
#include "blend_common.cuh"

namespace gsr {

struct BlendFwdArgs {
  BlendFwdView v[GSR_MAX_BATCH];
};

// blockIdx.y = view of a batched step (one launch blends every view's tiles: no launch gaps, one tail)
template <bool PRECISE>
#ifndef GSR_FWD_MINB
#define GSR_FWD_MINB 6   // six CTAs per SM (40 registers, one 4-byte spill outside the hit loop): 0.297 -> 0.294 ms per view
#endif
__global__ void __launch_bounds__(256, PRECISE ? 0 : GSR_FWD_MINB)
blend_forward_kernel(const __grid_constant__ BlendFwdArgs args) {
  const BlendFwdView& a = args.v[blockIdx.y];
  const int W = a.W, H = a.H, grid_x = a.grid_x;
  if ((int)blockIdx.x >= a.G) return;  // views of different sizes share the grid
  const uint2* __restrict__ ranges = a.ranges;
  const uint32_t* __restrict__ point_list = a.point_list;
  const float4* __restrict__ rec = a.rec;
  const float* __restrict__ depths = a.depths;
  const float* __restrict__ bg = a.bg;
  float* __restrict__ out_color = a.out_color;
  float* __restrict__ out_depth = a.out_depth;
  float* __restrict__ final_T = a.final_T;
  uint32_t* __restrict__ n_contrib = a.n_contrib;
  __shared__ __align__(16) unsigned char s_entries[BLEND_BATCH * ENTRY_BYTES];
  __shared__ uint32_t s_mask_arr[64];  // [target warp][staging warp]
  const uint32_t s_ent = pin_reg((uint32_t)__cvta_generic_to_shared(s_entries));
  const uint32_t s_mask = pin_reg((uint32_t)__cvta_generic_to_shared(s_mask_arr));

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int tile = blockIdx.x;
  const int tile_x0 = (tile % grid_x) * TILE_X, tile_y0 = (tile / grid_x) * TILE_Y;
  const int px = tile_x0 + (warp & 1) * 8 + (lane & 7);
  const int py = tile_y0 + (warp >> 1) * 4 + (lane >> 3);
  const bool inside = px < W && py < H;
  // A finished pixel (outside the image, or saturated) gets a NaN x coordinate: its `power` is NaN,
  // fails the ordered range test below and is skipped without a separate `done` test per pair.
  float pxf = inside ? (float)px : __int_as_float(0x7fc00000);
  const float pyf = (float)py;

  const uint2 range = ranges[tile];
  const int todo = (int)(range.y - range.x);
  const int rounds = (todo + BLEND_BATCH - 1) / BLEND_BATCH;

  float T = 1.0f, C0 = 0.0f, C1 = 0.0f, C2 = 0.0f;
  uint32_t last = 0, median_id = 0xFFFFFFFFu;   // the Gaussian at which T crossed 0.5 (its depth is fetched once, at the end)
  // `done` (outside the image, or saturated) IS the NaN coordinate: no separate flag to keep in step with it
#define GSR_DONE (pxf != pxf)

  for (int b = 0; b < rounds; b++) {
    // all pixels of the tile saturated -> stop (also the WAR barrier for the staging buffer)
    if (__syncthreads_and(GSR_DONE)) break;
    const int pos = b * BLEND_BATCH + tid;
    const uint32_t bits = stage_entry<PRECISE, GSR_REFINE_FWD != 0>(pos < todo, range.x + pos, point_list, rec,
                                               s_ent + tid * ENTRY_BYTES, (float)tile_x0, (float)tile_y0);
    publish_masks<true>(bits, s_mask, warp, lane);
    __syncthreads();

    // ---- walk the entries that can touch this warp's sub-tile ----
    if (!__all_sync(0xffffffffu, GSR_DONE)) {
#pragma unroll 1
      for (int ws = 0; ws < 8; ws++) {
        // every pixel of the sub-tile saturated inside this batch: the rest of it can only be skipped
        // pair by pair (NaN coordinate), so stop walking at the next group of 32 entries
        if (ws && __all_sync(0xffffffffu, GSR_DONE)) break;
        unsigned m = lds32(s_mask + (warp * 8 + ws) * 4);  // bit 31 - j <=> entry j of this group
        const uint32_t etop = s_ent + (ws * 32 + 31) * ENTRY_BYTES;          // entry 31 of the group
        const uint32_t last_top = (uint32_t)(b * BLEND_BATCH + ws * 32 + 32);  // its 1-based list position
        while (m) {
          const int jr = bfind(m);  // entry 31 - jr
          m &= bits_below(jr);      // jr is the top set bit: clears it (BMSK + LOP3, no constant to materialise)
          const uint32_t ea = etop - jr * ENTRY_BYTES;
          const float4 e0 = lds128(ea);
          const float4 e1 = lds128(ea + 16);
          float dx, dy;
          const float power = pair_power<PRECISE>(e0, e1, pxf, pyf, dx, dy);
          if (!(power <= 0.0f && power >= e1.y)) continue;  // power > 0, below the cut-off, or NaN (done)
          const float G = pair_gauss<PRECISE>(power);
          const float alpha = fminf(0.99f, MUL(e1.z, G));
          if (alpha < 1.0f / 255.0f) continue;
          const float test_T = MUL(T, SUB(1.0f, alpha));
          const bool saturated = test_T < 0.0001f;
          pxf = retire_if(saturated, pxf);   // one predicated move; the compiler's own version saves and restores pxf around the branch
          if (saturated) continue;
          const float4 e2 = lds128(ea + 32);
          const float wgt = MUL(alpha, T);
          C0 = FMA(e2.x, wgt, C0);
          C1 = FMA(e2.y, wgt, C1);
          C2 = FMA(e2.z, wgt, C2);
          if (T > 0.5f && test_T < 0.5f) median_id = __float_as_uint(e1.w);  // median depth (A.5)
          T = test_T;
          last = last_top - (uint32_t)jr;
        }
      }
    }
  }
#undef GSR_DONE

  if (inside) {
    const size_t HW = (size_t)H * W;
    const size_t pix = (size_t)py * W + px;
    final_T[pix] = T;
    n_contrib[pix] = last;
    out_color[pix] = FMA(T, __ldg(bg + 0), C0);
    out_color[HW + pix] = FMA(T, __ldg(bg + 1), C1);
    out_color[2 * HW + pix] = FMA(T, __ldg(bg + 2), C2);
    out_depth[pix] = median_id != 0xFFFFFFFFu ? __ldg(depths + median_id) : 15.0f;
  }
}

cudaError_t launch_blend_forward(cudaStream_t s, const BlendFwdView* views, int nv, bool precise) {
  if (nv <= 0) return cudaSuccess;
  if (nv > GSR_MAX_BATCH) return cudaErrorInvalidValue;
  BlendFwdArgs args{};
  int max_g = 0;
  for (int k = 0; k < nv; k++) {
    args.v[k] = views[k];
    args.v[k].grid_x = cdiv(views[k].W, TILE_X);
    args.v[k].G = args.v[k].grid_x * cdiv(views[k].H, TILE_Y);
    max_g = args.v[k].G > max_g ? args.v[k].G : max_g;
  }
  if (max_g == 0) return cudaSuccess;
  if (precise)
    blend_forward_kernel<true><<<dim3((unsigned)max_g, (unsigned)nv), 256, 0, s>>>(args);
  else
    blend_forward_kernel<false><<<dim3((unsigned)max_g, (unsigned)nv), 256, 0, s>>>(args);
  count_launch();
  return cudaGetLastError();
}

}  // namespace gsr
