// blend_common.cuh -- shared pieces of K6 / K7: tile geometry, staging of list entries into shared
// memory, per-warp cull masks, and the pair evaluation in its two arithmetic flavours.
//
//   PRECISE = true  : the oracle's exact fp32 op order, expf(), IEEE division (parity build)
//   PRECISE = false : default.  The staging thread pre-scales the conic once per tile entry
//                     (A = -0.5 log2e cx, B = -log2e cy, C = -0.5 log2e cz) so a pair costs
//                     5 FP ops + one MUFU.EX2 (ex2.approx, rel. error 2^-22) instead of 7 FP ops +
//                     a 10-instruction expf.  Colour stays well inside the 1e-5 bar of
//                     BASELINE.json (measured ~2e-7).
//
// Shared-memory entry (48 bytes, one per staged list element, read by all lanes as broadcast):
//   +0  {x, y, A|cx, B|cy}     +16 {C|cz, power_cut, opacity, id}     +32 {r, g, b, -}
// Addresses are kept as 32-bit shared-window offsets and read with ld.shared.v4: nvcc otherwise
// re-derives the (cluster-aware) shared window base inside the hot loop (S2UR SR_CgaCtaId + ULEA
// per access in the first ncu capture, profiles/).
#pragma once

#include "common.cuh"

namespace gsr {

constexpr int BLEND_BATCH = 256;
constexpr int ENTRY_BYTES = 48;
#define GSR_LOG2E 1.4426950408889634f

__device__ __forceinline__ float4 lds128(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ void sts128(uint32_t addr, float4 v) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ void sts64(uint32_t addr, float a, float b) {
  asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(addr), "f"(a), "f"(b) : "memory");
}
__device__ __forceinline__ uint32_t lds32(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ void sts32(uint32_t addr, uint32_t v) {
  asm volatile("st.shared.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float rcp_approx(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// Exact ellipse-vs-sub-tile test (the bounding-box test above it passes 30 % of its hits to pairs no pixel of which can
// contribute -- elongated, diagonal splats; a CPU replay of the headline view, tests/analysis_blend_hits.py, finds that
// 93 % of those "dead" hits are such geometric misses).  A pair contributes only if power >= power_cut, i.e.
//     q(dx, dy) = a dx^2 + 2 b dx dy + c dy^2 <= -2 power_cut          (a, b, c the conic; d = mean - pixel)
// so sub-tile v can be dropped when the MINIMUM of q over the rectangle spanned by its pixel centres exceeds that bound.
// q is convex: the minimum is 0 when the mean lies inside the rectangle, otherwise it is attained on one of the four
// edges, where q is a 1-D quadratic minimised in closed form (clamped).  Conservative by construction (the rectangle is
// a superset of the pixel centres; power_cut carries GSR_POWER_SLACK = 0.05, i.e. 0.1 in q); when the terms of q get
// large enough for their fp32 rounding to approach that slack the test is skipped and the bounding-box bits stand.
// The terms of the four x-edges and eight y-edges are shared between the sub-tiles: ~220 instructions per staged entry,
// once per (tile, entry) -- against ~26 warp instructions saved per dead (warp, entry) hit in K6 and K7.
// Build switches (tools/build_variant.py): which kernels refine, and how the code is emitted.
#ifndef GSR_REFINE_FWD
#define GSR_REFINE_FWD 0
#endif
#ifndef GSR_REFINE_BWD
#define GSR_REFINE_BWD 1
#endif
#ifdef GSR_REFINE_NOINLINE
#define GSR_REFINE_ATTR static __noinline__
#else
#define GSR_REFINE_ATTR __forceinline__
#endif
__device__ GSR_REFINE_ATTR uint32_t ellipse_subtile_mask(float4 q0, float conic_z, float power_cut, float tile_x0,
                                                         float tile_y0) {
  const float a = q0.z, b = q0.w, c = conic_z;
  const float thr = -2.0f * power_cut;
  // d = mean - pixel over the whole tile: dx in [x - tx0 - 15, x - tx0], dy likewise
  const float X1 = q0.x - tile_x0, Y1 = q0.y - tile_y0;
  const float mx = fmaxf(fabsf(X1), fabsf(X1 - 15.0f)), my = fmaxf(fabsf(Y1), fabsf(Y1 - 15.0f));
  if (!(a * mx * mx + c * my * my + 2.0f * fabsf(b) * mx * my < 3.0e4f)) return 0xffu;   // rounding guard (also NaN)
  const float r1 = -b / c, r2 = -b / a;   // minimisers: dy*(dx) = r1 dx on a vertical edge, dx*(dy) = r2 dy on a horizontal one
  const float b2 = 2.0f * b;
  // vertical edges: dx = X1 - {0, 7, 8, 15}  (sub-tile column k spans pixel columns 8 k .. 8 k + 7)
  float xe[4], xu[4], xw[4], xz[4];
#pragma unroll
  for (int k = 0; k < 4; k++) {
    xe[k] = X1 - (float)((k >> 1) * 8 + (k & 1) * 7);
    xu[k] = r1 * xe[k];
    xw[k] = b2 * xe[k];
    xz[k] = a * xe[k] * xe[k];
  }
  uint32_t keep = 0;
#pragma unroll
  for (int r = 0; r < 4; r++) {
    const float yhi = Y1 - (float)(4 * r), ylo = yhi - 3.0f;            // dy range of sub-tile row r: [ylo, yhi]
    // horizontal edges of this row: dy = yhi, ylo
    const float u0 = r2 * yhi, w0 = b2 * yhi, z0 = c * yhi * yhi;
    const float u1 = r2 * ylo, w1 = b2 * ylo, z1 = c * ylo * ylo;
    const bool y_in = ylo <= 0.0f && yhi >= 0.0f;
#pragma unroll
    for (int k = 0; k < 2; k++) {
      const float xhi_ = xe[2 * k], xlo_ = xe[2 * k + 1];               // dx range of sub-tile column k: [xlo_, xhi_]
      float t, qmin;
      t = fminf(fmaxf(xu[2 * k], ylo), yhi);      qmin = fmaf(fmaf(c, t, xw[2 * k]), t, xz[2 * k]);
      t = fminf(fmaxf(xu[2 * k + 1], ylo), yhi);  qmin = fminf(qmin, fmaf(fmaf(c, t, xw[2 * k + 1]), t, xz[2 * k + 1]));
      t = fminf(fmaxf(u0, xlo_), xhi_);           qmin = fminf(qmin, fmaf(fmaf(a, t, w0), t, z0));
      t = fminf(fmaxf(u1, xlo_), xhi_);           qmin = fminf(qmin, fmaf(fmaf(a, t, w1), t, z1));
      const bool inside = y_in && xlo_ <= 0.0f && xhi_ >= 0.0f;
      if (inside || qmin <= thr) keep |= 1u << (2 * r + k);
    }
  }
  return keep;
}

// Stages list element `pos` (if valid) into entry slot `tid` and returns the 8-bit mask of warps
// (8x4 pixel sub-tiles, warp v = row*2 + col) whose box the Gaussian's {alpha >= 1/255} bounding
// box [x - hx, x + hx] x [y - hy, y + hy] overlaps.
template <bool PRECISE, bool REFINE_CULL>
__device__ __forceinline__ uint32_t stage_entry(bool valid, uint32_t list_index,
                                                const uint32_t* __restrict__ point_list,
                                                const float4* __restrict__ rec, uint32_t s_entry,
                                                float tile_x0, float tile_y0) {
  uint32_t bits = 0;
  if (valid) {
    const uint32_t id = __ldg(point_list + list_index);
    const float4 q0 = __ldg(rec + 3 * (size_t)id);      // x, y, conic.x, conic.y
    const float4 q1 = __ldg(rec + 3 * (size_t)id + 1);  // conic.z, opacity, hx, hy
    const float4 q2 = __ldg(rec + 3 * (size_t)id + 2);  // r, g, b, power_cut
    float4 e0, e1;
    if (PRECISE) {
      e0 = q0;
      e1 = make_float4(q1.x, q2.w, q1.y, __uint_as_float(id));
    } else {
      e0 = make_float4(q0.x, q0.y, (-0.5f * GSR_LOG2E) * q0.z, -GSR_LOG2E * q0.w);
      e1 = make_float4((-0.5f * GSR_LOG2E) * q1.x, GSR_LOG2E * q2.w, q1.y, __uint_as_float(id));
    }
    sts128(s_entry, e0);
    sts128(s_entry + 16, e1);
    sts128(s_entry + 32, q2);
    const float xlo = q0.x - q1.z, xhi = q0.x + q1.z, ylo = q0.y - q1.w, yhi = q0.y + q1.w;
    const uint32_t cx = ((xhi >= tile_x0 && xlo <= tile_x0 + 7.0f) ? 1u : 0u) |
                        ((xhi >= tile_x0 + 8.0f && xlo <= tile_x0 + 15.0f) ? 2u : 0u);
#pragma unroll
    for (int r = 0; r < 4; r++)
      if (yhi >= tile_y0 + 4.0f * r && ylo <= tile_y0 + 4.0f * r + 3.0f) bits |= cx << (2 * r);
    if (REFINE_CULL) bits &= ellipse_subtile_mask(q0, q1.x, q2.w, tile_x0, tile_y0);
  }
  return bits;
}

// Publishes, for every target warp v, the ballot of "entry touches v's sub-tile" over this warp's
// 32 staged entries: s_mask[v][warp].
// REVERSED: bit (31 - i) stands for entry i, so that a front-to-back walk takes the next entry with
// one count-leading-zeros (FLO.SH) instead of bit-reverse + find-leading-one.
template <bool REVERSED>
__device__ __forceinline__ void publish_masks(uint32_t bits, uint32_t s_mask, int warp, int lane) {
#pragma unroll
  for (int v = 0; v < 8; v++) {
    unsigned m = __ballot_sync(0xffffffffu, (bits >> v) & 1u);
    if (REVERSED) m = __brev(m);
    if (lane == 0) sts32(s_mask + (v * 8 + warp) * 4, m);
  }
}
// index of the most significant set bit (x != 0): one FLO
__device__ __forceinline__ int bfind(uint32_t x) {
  int r;
  asm("bfind.u32 %0, %1;" : "=r"(r) : "r"(x));
  return r;
}
// (1 << n) - 1 for n in [0, 31]: one BMSK
__device__ __forceinline__ uint32_t bits_below(int n) {
  uint32_t r;
  asm("bmsk.clamp.b32 %0, 0, %1;" : "=r"(r) : "r"(n));
  return r;
}
// x, or a quiet NaN when `cond`: one predicated move (a finished pixel's coordinate, see blend_forward.cu)
__device__ __forceinline__ float retire_if(bool cond, float x) {
  asm("{ .reg .pred p; setp.ne.s32 p, %1, 0; @p mov.b32 %0, 0x7fc00000; }" : "+f"(x) : "r"((int)cond));
  return x;
}
// keeps a shared-window address in a register (nvcc otherwise re-derives it -- S2UR + ULEA + MOV --
// inside the hot loop)
__device__ __forceinline__ uint32_t pin_reg(uint32_t x) {
  asm volatile("" : "+r"(x));
  return x;
}

// power (PRECISE) or power * log2e (fast) of entry (e0, e1) at pixel (pxf, pyf); also returns d.
template <bool PRECISE>
__device__ __forceinline__ float pair_power(const float4& e0, const float4& e1, float pxf, float pyf,
                                            float& dx, float& dy) {
  dx = SUB(e0.x, pxf);
  dy = SUB(e0.y, pyf);
  if (PRECISE) {
    const float q = FMA(MUL(e1.x, dy), dy, MUL(MUL(e0.z, dx), dx));
    return FMA(-0.5f, q, -MUL(MUL(e0.w, dx), dy));
  } else {
    const float t = fmaf(e0.w, dy, e0.z * dx);  // A dx + B dy
    return fmaf(e1.x * dy, dy, t * dx);         // dx (A dx + B dy) + C dy^2
  }
}
template <bool PRECISE>
__device__ __forceinline__ float pair_gauss(float power) {
  return PRECISE ? expf(power) : ex2_approx(power);
}

}  // namespace gsr
