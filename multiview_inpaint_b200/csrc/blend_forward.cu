// blend_forward.cu -- K6: per-16x16-tile front-to-back alpha blend with median depth.
//
// Restates SURVEY.md Appendix A.5.  The depth sentinel 15.0f and the (1,H,W) depth output are
// pinned by the reference callers gs-simp/gen_seq.py:50, vis_render.py:45, render_depth.py:37.
//
// B200 design (differs from the public kernel, same per-pixel arithmetic and order):
//   * one CTA = one tile = 8 warps; warp w owns an 8x4 pixel sub-tile, so a warp is a compact
//     screen-space box and can be culled as a unit;
//   * the tile's list is staged 256 entries at a time into shared memory as the 48-byte blend
//     records written by K1 (3 x LDG.128 per entry, 2 sectors);
//   * while staging, each thread tests its Gaussian's alpha>=1/255 bounding box against the eight
//     sub-tiles and the CTA publishes, per warp, a 256-bit "touches my sub-tile" mask (ballots);
//     a warp then walks only the set bits (warp-uniform loop, no divergence on the skip);
//   * the exp is only evaluated for pairs whose power is above the Gaussian's cut-off.
//   Both culls are conservative (GSR_POWER_SLACK), i.e. they only drop pairs the reference would
//   have dropped with `alpha < 1/255`, so n_contrib / final_T / colour / depth are unchanged.
//   * saturated pixels: per-lane `done`, warp-level __all_sync to stop walking, CTA-level
//     __syncthreads_and to stop staging (the reference's __syncthreads_count early exit).
#include "common.cuh"

namespace gsr {

template <bool FAST_EXP>
__global__ void __launch_bounds__(256)
blend_forward_kernel(int W, int H, int grid_x, const uint2* __restrict__ ranges,
                     const uint32_t* __restrict__ point_list, const float4* __restrict__ rec,
                     const float* __restrict__ depths, const float* __restrict__ bg,
                     float* __restrict__ out_color, float* __restrict__ out_depth,
                     float* __restrict__ final_T, uint32_t* __restrict__ n_contrib) {
  __shared__ float4 s_q0[256];
  __shared__ float4 s_q1[256];
  __shared__ float4 s_q2[256];
  __shared__ uint32_t s_id[256];
  __shared__ uint32_t s_mask[8][8];  // [target warp][staging warp]

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int tile = blockIdx.x;
  const int tile_x0 = (tile % grid_x) * TILE_X, tile_y0 = (tile / grid_x) * TILE_Y;
  const int px = tile_x0 + (warp & 1) * 8 + (lane & 7);
  const int py = tile_y0 + (warp >> 1) * 4 + (lane >> 3);
  const bool inside = px < W && py < H;
  const float pxf = (float)px, pyf = (float)py;

  const uint2 range = ranges[tile];
  const int todo = (int)(range.y - range.x);
  const int rounds = (todo + 255) / 256;

  float T = 1.0f, C0 = 0.0f, C1 = 0.0f, C2 = 0.0f, Dm = 15.0f;
  uint32_t last = 0;
  bool done = !inside;

  for (int b = 0; b < rounds; b++) {
    // all pixels of the tile saturated -> stop (also the WAR barrier for the staging buffers)
    if (__syncthreads_and(done)) break;

    // ---- stage 256 list entries + cull masks ----
    const int pos = b * 256 + tid;
    uint32_t bits = 0;
    if (pos < todo) {
      const uint32_t id = point_list[range.x + pos];
      const float4 q0 = __ldg(rec + 3 * (size_t)id);
      const float4 q1 = __ldg(rec + 3 * (size_t)id + 1);
      const float4 q2 = __ldg(rec + 3 * (size_t)id + 2);
      s_q0[tid] = q0;
      s_q1[tid] = q1;
      s_q2[tid] = q2;
      s_id[tid] = id;
      // bounding box of {alpha >= 1/255} : [x - hx, x + hx] x [y - hy, y + hy]
      const float xlo = q0.x - q1.z, xhi = q0.x + q1.z, ylo = q0.y - q1.w, yhi = q0.y + q1.w;
      const float tx = (float)tile_x0, ty = (float)tile_y0;
      const uint32_t cx = ((xhi >= tx && xlo <= tx + 7.0f) ? 1u : 0u) |
                          ((xhi >= tx + 8.0f && xlo <= tx + 15.0f) ? 2u : 0u);
      uint32_t cy = 0;
#pragma unroll
      for (int r = 0; r < 4; r++)
        cy |= (yhi >= ty + 4.0f * r && ylo <= ty + 4.0f * r + 3.0f) ? (1u << r) : 0u;
      // warp v = (row r)*2 + col c
#pragma unroll
      for (int r = 0; r < 4; r++)
        if (cy & (1u << r)) bits |= cx << (2 * r);
    }
#pragma unroll
    for (int v = 0; v < 8; v++) {
      const unsigned m = __ballot_sync(0xffffffffu, (bits >> v) & 1u);
      if (lane == 0) s_mask[v][warp] = m;
    }
    __syncthreads();

    // ---- walk the entries that can touch this warp's sub-tile ----
    if (!__all_sync(0xffffffffu, done)) {
#pragma unroll 1
      for (int ws = 0; ws < 8; ws++) {
        unsigned m = s_mask[warp][ws];
        while (m) {
          const int j = __ffs(m) - 1;
          m &= m - 1;
          if (done) continue;
          const int e = ws * 32 + j;
          const float4 q0 = s_q0[e];
          const float4 q1 = s_q1[e];
          const float4 q2 = s_q2[e];
          const float dx = SUB(q0.x, pxf), dy = SUB(q0.y, pyf);
          const float q = FMA(MUL(q1.x, dy), dy, MUL(MUL(q0.z, dx), dx));
          const float power = FMA(-0.5f, q, -MUL(MUL(q0.w, dx), dy));
          if (power > 0.0f || power < q2.w) continue;
          const float G = FAST_EXP ? __expf(power) : expf(power);
          const float alpha = fminf(0.99f, MUL(q1.y, G));
          if (alpha < 1.0f / 255.0f) continue;
          const float test_T = MUL(T, SUB(1.0f, alpha));
          if (test_T < 0.0001f) {
            done = true;
            continue;
          }
          const float wgt = MUL(alpha, T);
          C0 = FMA(q2.x, wgt, C0);
          C1 = FMA(q2.y, wgt, C1);
          C2 = FMA(q2.z, wgt, C2);
          if (T > 0.5f && test_T < 0.5f) Dm = __ldg(depths + s_id[e]);  // median depth
          T = test_T;
          last = (uint32_t)(b * 256 + e + 1);
        }
      }
    }
  }

  if (inside) {
    const size_t HW = (size_t)H * W;
    const size_t pix = (size_t)py * W + px;
    final_T[pix] = T;
    n_contrib[pix] = last;
    out_color[pix] = FMA(T, __ldg(bg + 0), C0);
    out_color[HW + pix] = FMA(T, __ldg(bg + 1), C1);
    out_color[2 * HW + pix] = FMA(T, __ldg(bg + 2), C2);
    out_depth[pix] = Dm;
  }
}

cudaError_t launch_blend_forward(cudaStream_t s, int W, int H, const uint2* ranges,
                                 const uint32_t* point_list, const float4* rec, const float* depths,
                                 const float* bg, float* out_color, float* out_depth,
                                 float* final_T, uint32_t* n_contrib, bool fast_exp) {
  const int gx = cdiv(W, TILE_X), gy = cdiv(H, TILE_Y);
  if (gx * gy == 0) return cudaSuccess;
  if (fast_exp)
    blend_forward_kernel<true><<<gx * gy, 256, 0, s>>>(W, H, gx, ranges, point_list, rec, depths, bg,
                                                       out_color, out_depth, final_T, n_contrib);
  else
    blend_forward_kernel<false><<<gx * gy, 256, 0, s>>>(W, H, gx, ranges, point_list, rec, depths, bg,
                                                        out_color, out_depth, final_T, n_contrib);
  count_launch();
  return cudaGetLastError();
}

}  // namespace gsr
