// blend_backward.cu -- K7: back-to-front gradient of the tile blend.
//
// Restates SURVEY.md Appendix A.6 (per-pixel recurrences identical: T /= (1-alpha), accum_rec,
// last_alpha / last_color, the -T_final/(1-alpha) * <bg, dL_dpixel> term, 0.99 clamp and thresholds
// as pass-through, no gradient through depth).
//
// B200 design.  The public kernel issues 9 global atomicAdds per contributing (pixel, Gaussian)
// pair.  Here:
//   * same CTA/warp/sub-tile geometry and per-warp cull masks as K6, walked in reverse;
//   * the walk starts at the CTA's max n_contrib, not at the end of the tile list (entries past
//     every pixel's last contributor are never even staged);
//   * per (warp, Gaussian) the nine partial sums are reduced across the 32 lanes with a
//     transposing butterfly (9 shuffles for 8 values + 5 for the ninth) and ONE lane per value
//     issues a RED.ADD.F32 -- <= 9 L2 reductions per (warp, Gaussian) instead of 9 x 32;
//   * reductions land in a packed 48-byte accumulator per Gaussian (2 sectors):
//       [0..2] dL/dcolor   [3] A = sum dL_dG*G*dx   [4] B = sum dL_dG*G*dy
//       [5..7] dL/dconic (xx, xy, yy)               [8] dL/dopacity
//     dL/dmean2D = -(cx A + cy B) * W/2, -(cz B + cy A) * H/2 is formed once per Gaussian in K8/K9
//     (cx,cy,cz are per-Gaussian constants, so this is the same sum the reference accumulates).
#include "common.cuh"

namespace gsr {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Reduces v[0..7] over the warp; on return lane l holds the total of value index (l >> 2).
__device__ __forceinline__ float warp_transpose_reduce8(const float v[8], int lane) {
  const bool h16 = lane & 16;
  float a0 = h16 ? v[4] : v[0], a1 = h16 ? v[5] : v[1], a2 = h16 ? v[6] : v[2], a3 = h16 ? v[7] : v[3];
  const float b0 = h16 ? v[0] : v[4], b1 = h16 ? v[1] : v[5], b2 = h16 ? v[2] : v[6], b3 = h16 ? v[3] : v[7];
  a0 += __shfl_xor_sync(0xffffffffu, b0, 16);
  a1 += __shfl_xor_sync(0xffffffffu, b1, 16);
  a2 += __shfl_xor_sync(0xffffffffu, b2, 16);
  a3 += __shfl_xor_sync(0xffffffffu, b3, 16);
  const bool h8 = lane & 8;
  float c0 = h8 ? a2 : a0, c1 = h8 ? a3 : a1;
  const float d0 = h8 ? a0 : a2, d1 = h8 ? a1 : a3;
  c0 += __shfl_xor_sync(0xffffffffu, d0, 8);
  c1 += __shfl_xor_sync(0xffffffffu, d1, 8);
  const bool h4 = lane & 4;
  float e0 = h4 ? c1 : c0;
  const float f0 = h4 ? c0 : c1;
  e0 += __shfl_xor_sync(0xffffffffu, f0, 4);
  e0 += __shfl_xor_sync(0xffffffffu, e0, 2);
  e0 += __shfl_xor_sync(0xffffffffu, e0, 1);
  return e0;
}

template <bool FAST_EXP>
__global__ void __launch_bounds__(256)
blend_backward_kernel(int W, int H, int grid_x, const uint2* __restrict__ ranges,
                      const uint32_t* __restrict__ point_list, const float4* __restrict__ rec,
                      const float* __restrict__ bg, const float* __restrict__ final_T,
                      const uint32_t* __restrict__ n_contrib, const float* __restrict__ dL_dpix,
                      float* __restrict__ gacc) {
  __shared__ float4 s_q0[256];
  __shared__ float4 s_q1[256];
  __shared__ float4 s_q2[256];
  __shared__ uint32_t s_id[256];
  __shared__ uint32_t s_mask[8][8];
  __shared__ uint32_t s_wmax[8];

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int tile = blockIdx.x;
  const int tile_x0 = (tile % grid_x) * TILE_X, tile_y0 = (tile / grid_x) * TILE_Y;
  const int px = tile_x0 + (warp & 1) * 8 + (lane & 7);
  const int py = tile_y0 + (warp >> 1) * 4 + (lane >> 3);
  const bool inside = px < W && py < H;
  const float pxf = (float)px, pyf = (float)py;
  const size_t HW = (size_t)H * W;
  const size_t pix = (size_t)py * W + px;

  const uint2 range = ranges[tile];

  const float T_final = inside ? final_T[pix] : 0.0f;
  float T = T_final;
  const uint32_t last = inside ? n_contrib[pix] : 0u;
  float dLp0 = 0.0f, dLp1 = 0.0f, dLp2 = 0.0f;
  if (inside) {
    dLp0 = __ldg(dL_dpix + pix);
    dLp1 = __ldg(dL_dpix + HW + pix);
    dLp2 = __ldg(dL_dpix + 2 * HW + pix);
  }
  const float bg_dot = FMA(__ldg(bg + 2), dLp2, FMA(__ldg(bg + 1), dLp1, MUL(__ldg(bg + 0), dLp0)));
  float acc0 = 0.0f, acc1 = 0.0f, acc2 = 0.0f;   // accum_rec
  float lc0 = 0.0f, lc1 = 0.0f, lc2 = 0.0f;      // last_color
  float last_alpha = 0.0f;

  const uint32_t warp_last = __reduce_max_sync(0xffffffffu, last);
  if (lane == 0) s_wmax[warp] = warp_last;
  __syncthreads();
  uint32_t cta_last = 0;
#pragma unroll
  for (int w = 0; w < 8; w++) cta_last = max(cta_last, s_wmax[w]);
  if (cta_last == 0) return;  // nothing contributed anywhere in this tile (uniform exit)
  const int nb = (int)((cta_last + 255) / 256);

  for (int b = nb - 1; b >= 0; b--) {
    __syncthreads();  // WAR on the staging buffers
    const uint32_t pos = (uint32_t)b * 256 + tid;
    uint32_t bits = 0;
    if (pos < cta_last) {
      const uint32_t id = point_list[range.x + pos];
      const float4 q0 = __ldg(rec + 3 * (size_t)id);
      const float4 q1 = __ldg(rec + 3 * (size_t)id + 1);
      const float4 q2 = __ldg(rec + 3 * (size_t)id + 2);
      s_q0[tid] = q0;
      s_q1[tid] = q1;
      s_q2[tid] = q2;
      s_id[tid] = id;
      const float xlo = q0.x - q1.z, xhi = q0.x + q1.z, ylo = q0.y - q1.w, yhi = q0.y + q1.w;
      const float tx = (float)tile_x0, ty = (float)tile_y0;
      const uint32_t cx = ((xhi >= tx && xlo <= tx + 7.0f) ? 1u : 0u) |
                          ((xhi >= tx + 8.0f && xlo <= tx + 15.0f) ? 2u : 0u);
      uint32_t cy = 0;
#pragma unroll
      for (int r = 0; r < 4; r++)
        cy |= (yhi >= ty + 4.0f * r && ylo <= ty + 4.0f * r + 3.0f) ? (1u << r) : 0u;
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

    if (warp_last > (uint32_t)b * 256) {
#pragma unroll 1
      for (int ws = 7; ws >= 0; ws--) {
        unsigned m = s_mask[warp][ws];
        const uint32_t gbase = (uint32_t)b * 256 + ws * 32;
        if (gbase >= warp_last) continue;
        if (warp_last - gbase < 32) m &= (1u << (warp_last - gbase)) - 1;  // entries >= warp_last
        while (m) {
          const int j = 31 - __clz(m);
          m &= ~(1u << j);
          const int e = ws * 32 + j;
          const uint32_t epos = gbase + j;  // 0-based list position == contributor index
          const float4 q0 = s_q0[e];
          const float4 q1 = s_q1[e];
          const float4 q2 = s_q2[e];
          const float dx = SUB(q0.x, pxf), dy = SUB(q0.y, pyf);
          const float q = FMA(MUL(q1.x, dy), dy, MUL(MUL(q0.z, dx), dx));
          const float power = FMA(-0.5f, q, -MUL(MUL(q0.w, dx), dy));
          bool contrib = (epos < last) && !(power > 0.0f || power < q2.w);
          float G = 0.0f, alpha = 0.0f;
          if (contrib) {
            G = FAST_EXP ? __expf(power) : expf(power);
            alpha = fminf(0.99f, MUL(q1.y, G));
            contrib = !(alpha < 1.0f / 255.0f);
          }
          if (!__any_sync(0xffffffffu, contrib)) continue;

          float v[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
          float v8 = 0.0f;
          if (contrib) {
            const float one_m_alpha = SUB(1.0f, alpha);
            T = DIV(T, one_m_alpha);
            const float dch = MUL(alpha, T);
            acc0 = FMA(last_alpha, lc0, MUL(SUB(1.0f, last_alpha), acc0));
            acc1 = FMA(last_alpha, lc1, MUL(SUB(1.0f, last_alpha), acc1));
            acc2 = FMA(last_alpha, lc2, MUL(SUB(1.0f, last_alpha), acc2));
            lc0 = q2.x; lc1 = q2.y; lc2 = q2.z;
            float dL_dalpha = MUL(SUB(q2.x, acc0), dLp0);
            dL_dalpha = FMA(SUB(q2.y, acc1), dLp1, dL_dalpha);
            dL_dalpha = FMA(SUB(q2.z, acc2), dLp2, dL_dalpha);
            v[0] = MUL(dch, dLp0);
            v[1] = MUL(dch, dLp1);
            v[2] = MUL(dch, dLp2);
            dL_dalpha = MUL(dL_dalpha, T);
            last_alpha = alpha;
            dL_dalpha = FMA(DIV(-T_final, one_m_alpha), bg_dot, dL_dalpha);
            const float dL_dG = MUL(q1.y, dL_dalpha);
            const float gdx = MUL(G, dx), gdy = MUL(G, dy);
            v[3] = MUL(dL_dG, gdx);
            v[4] = MUL(dL_dG, gdy);
            v[5] = MUL(MUL(MUL(-0.5f, gdx), dx), dL_dG);
            v[6] = MUL(MUL(MUL(-0.5f, gdx), dy), dL_dG);
            v[7] = MUL(MUL(MUL(-0.5f, gdy), dy), dL_dG);
            v8 = MUL(G, dL_dalpha);
          }
          const float r8 = warp_transpose_reduce8(v, lane);
          const float r1 = warp_sum(v8);
          float* dst = gacc + (size_t)s_id[e] * 12;
          if ((lane & 3) == 0) atomicAdd(dst + (lane >> 2), r8);
          if (lane == 1) atomicAdd(dst + 8, r1);
        }
      }
    }
  }
}

cudaError_t launch_blend_backward(cudaStream_t s, int W, int H, const uint2* ranges,
                                  const uint32_t* point_list, const float4* rec, const float* bg,
                                  const float* final_T, const uint32_t* n_contrib,
                                  const float* dL_dpix, float* gacc, bool fast_exp) {
  const int gx = cdiv(W, TILE_X), gy = cdiv(H, TILE_Y);
  if (gx * gy == 0) return cudaSuccess;
  if (fast_exp)
    blend_backward_kernel<true><<<gx * gy, 256, 0, s>>>(W, H, gx, ranges, point_list, rec, bg, final_T,
                                                        n_contrib, dL_dpix, gacc);
  else
    blend_backward_kernel<false><<<gx * gy, 256, 0, s>>>(W, H, gx, ranges, point_list, rec, bg, final_T,
                                                         n_contrib, dL_dpix, gacc);
  count_launch();
  return cudaGetLastError();
}

}  // namespace gsr
