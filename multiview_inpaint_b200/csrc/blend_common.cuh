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
//   +0  {x, y, A|cx, B|cy}     +16 {C|cz, opacity, power_cut, id}     +32 {r, g, b, -}
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

// Stages list element `pos` (if valid) into entry slot `tid` and returns the 8-bit mask of warps
// (8x4 pixel sub-tiles, warp v = row*2 + col) whose box the Gaussian's {alpha >= 1/255} bounding
// box [x - hx, x + hx] x [y - hy, y + hy] overlaps.
template <bool PRECISE>
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
      e1 = make_float4(q1.x, q1.y, q2.w, __uint_as_float(id));
    } else {
      e0 = make_float4(q0.x, q0.y, (-0.5f * GSR_LOG2E) * q0.z, -GSR_LOG2E * q0.w);
      e1 = make_float4((-0.5f * GSR_LOG2E) * q1.x, q1.y, GSR_LOG2E * q2.w, __uint_as_float(id));
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
