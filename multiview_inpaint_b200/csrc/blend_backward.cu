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
//     transposing butterfly (12 shuffles for the nine values) and ONE lane per value issues a
//     RED.ADD.F32 (one predicated instruction) -- 9 L2 reductions per (warp, Gaussian), not 9 x 32;
//   * reductions land in a packed 48-byte accumulator per Gaussian (2 sectors):
//       [0..2] dL/dcolor   [3] A = sum dL_dG*G*dx   [4] B = sum dL_dG*G*dy
//       [5..7] dL/dconic (xx, xy, yy)               [8] dL/dopacity
//     dL/dmean2D = -(cx A + cy B) * W/2, -(cz B + cy A) * H/2 is formed once per Gaussian in K8/K9
//     (cx,cy,cz are per-Gaussian constants, so this is the same sum the reference accumulates).
// Like K6 the kernel is instruction-issue bound; the default (non-PRECISE) flavour replaces the
// two IEEE divisions per pair by one rcp.approx and expf by ex2.approx (blend_common.cuh).
#include "blend_common.cuh"

namespace gsr {

// Reduces nine per-lane values over the warp with a transposing butterfly: 5 + 3 + 2 + 1 + 1 = 12
// shuffles (every stage halves the number of live values; the ninth rides along and is paired
// with the survivor of the other eight at the xor-2 stage).  On return lane l holds
//   (l & 2) == 0 : the total of value index (l >> 2)        (v0..v7)
//   (l & 2) != 0 : the total of v8
__device__ __forceinline__ float warp_transpose_reduce9(float v0, float v1, float v2, float v3,
                                                        float v4, float v5, float v6, float v7,
                                                        float v8, int lane) {
  const bool h16 = lane & 16;
  float a0 = h16 ? v4 : v0, a1 = h16 ? v5 : v1, a2 = h16 ? v6 : v2, a3 = h16 ? v7 : v3;
  const float b0 = h16 ? v0 : v4, b1 = h16 ? v1 : v5, b2 = h16 ? v2 : v6, b3 = h16 ? v3 : v7;
  a0 += __shfl_xor_sync(0xffffffffu, b0, 16);
  a1 += __shfl_xor_sync(0xffffffffu, b1, 16);
  a2 += __shfl_xor_sync(0xffffffffu, b2, 16);
  a3 += __shfl_xor_sync(0xffffffffu, b3, 16);
  v8 += __shfl_xor_sync(0xffffffffu, v8, 16);
  const bool h8 = lane & 8;
  float c0 = h8 ? a2 : a0, c1 = h8 ? a3 : a1;
  const float d0 = h8 ? a0 : a2, d1 = h8 ? a1 : a3;
  c0 += __shfl_xor_sync(0xffffffffu, d0, 8);
  c1 += __shfl_xor_sync(0xffffffffu, d1, 8);
  v8 += __shfl_xor_sync(0xffffffffu, v8, 8);
  const bool h4 = lane & 4;
  float e0 = h4 ? c1 : c0;
  const float f0 = h4 ? c0 : c1;
  e0 += __shfl_xor_sync(0xffffffffu, f0, 4);
  v8 += __shfl_xor_sync(0xffffffffu, v8, 4);
  const bool h2 = lane & 2;
  float k = h2 ? v8 : e0;
  const float snd = h2 ? e0 : v8;
  k += __shfl_xor_sync(0xffffffffu, snd, 2);
  k += __shfl_xor_sync(0xffffffffu, k, 1);
  return k;
}

// ---- the nine sums as ONE tensor-core contraction (default, non-PRECISE flavour) -------------------------------------
// Every one of the nine per-(warp, Gaussian) sums has the form  sum_p W(g, p) * F(p, f):
//     dL/dcolor_c  = sum_p dch(g,p)          * dL_dpixel_c(p)                       W1 = alpha * T,  F1 = dL_dpixel (3 cols)
//     S0,Su,..,Svv = sum_p G*dL_dalpha(g,p)  * {1, u, v, u^2, u v, v^2}(p)          W3 = G * dL_dalpha, FP = monomials of the
//                                                                                   pixel's offset (u, v) from the sub-tile centre
// and the sums the accumulator holds follow from the six moments with the Gaussian's own offset (X, Y) from that centre
// (dx = X - u, dy = Y - v):   sum w dx = X S0 - Su,   sum w dx^2 = X^2 S0 - 2 X Su + Suu,   sum w dx dy = X Y S0 - X Sv - Y Su + Suv, ...
// (u, v are half-integers |u| <= 3.5, |v| <= 1.5: no cancellation beyond what the per-pixel sums have, and the monomials
// are exact in tf32).  So instead of a 12-shuffle butterfly per (warp, Gaussian) -- ~45 of the ~106 instructions of a live
// hit -- a lane stores its two weights W1, W3 into a warp-private shared-memory tile [8 Gaussians][32 pixels], and every
// eighth live hit the warp runs D[feature][Gaussian] = F^T x W^T as mma.sync.m16n8k8 (tf32 inputs, fp32 accumulate):
// 4 k-steps x {hi, lo} x 2 matrices = 16 MMAs per 8 hits.  W (and dL_dpixel) are split hi + lo (hi = upper 19 bits,
// lo = x - hi, exact), so the products carry ~21 bits: the sums differ from the shuffle path by fp32 rounding only.
// Fragment layout (PTX ISA, m16n8k8 .tf32; q = lane >> 2, t = lane & 3):  A a0 = (row q, k t), a2 = (row q, k t + 4), rows
// q + 8 (a1, a3) are zero;  B b0 = (k t, col q), b1 = (k t + 4, col q);  D c0 = (row q, col 2 t), c1 = (row q, col 2 t + 1).
// k-step s maps k = t -> pixel 8 t + 2 s and k = t + 4 -> pixel 8 t + 2 s + 1 on BOTH operands, so a thread's fragments of
// all four k-steps are the 8 consecutive floats [8 t, 8 t + 8) of row q of a W tile: two LDS.128 (row stride 36 floats
// keeps them conflict free).  The epilogue transposes D through shared memory (slot-major), forms the nine values per
// Gaussian and issues them as 3 predicated RED instructions per 8 hits (was 8).
constexpr int MG_SLOTS = 8;                        // Gaussians per contraction (the MMA's N)
constexpr int MG_STRIDE = 36;                      // floats per row of a W tile [8 slots][32 pixels]
constexpr int MG_TILE_BYTES = MG_SLOTS * MG_STRIDE * 4;          // 1152
constexpr int MG_F_ROWS = 6;                       // feature rows kept (rows 6, 7 of the MMA's M are never read back)
constexpr int MG_F_STRIDE = 72;                    // floats per feature row: 32 pixels x {F1, FP} interleaved = 16 chunks of
                                                   // 16 bytes, the upper eight shifted by one chunk (bank swizzle), + pad
constexpr int MG_F_BYTES = MG_F_ROWS * MG_F_STRIDE * 4;          // 1728
constexpr int MG_INFO_OFF = 2 * MG_TILE_BYTES;                   // per warp: W1 tile | W3 tile | info[8] | feature table
constexpr int MG_F_OFF = MG_INFO_OFF + MG_SLOTS * 16;
constexpr int MG_WARP_BYTES = MG_F_OFF + MG_F_BYTES;             // 4160
constexpr int MG_M_STRIDE = 20;                    // floats per slot row of the transposed result (aliases the W1 tile)

// A = {a0, a1, a2, a3} (four consecutive registers, straight from one LDS.128), B = {b0, b1}
__device__ __forceinline__ void mma_m16n8k8_tf32(float (&d)[4], const float4& a, float b0, float b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(__float_as_uint(a.x)), "r"(__float_as_uint(a.y)), "r"(__float_as_uint(a.z)), "r"(__float_as_uint(a.w)),
        "r"(__float_as_uint(b0)), "r"(__float_as_uint(b1)));
}
__device__ __forceinline__ float tf32_hi(float x) { return __uint_as_float(__float_as_uint(x) & 0xffffe000u); }

// Contracts the `ns` filled slots of the warp's tiles and adds the nine sums of each into its Gaussian's accumulator.
// s_warp: the warp's region {W1 tile, W3 tile, info[8], feature table}; (cx, cy): sub-tile centre.
// The feature table interleaves F1 and FP per pixel, {F1[f][p], FP[f][p]}: one LDS.128 at pixel pair (8 t + 2 s, + 1) is
// the A fragment {a0, a1, a2, a3} of k-step s with F1 in rows 0..7 and FP in rows 8..15 of the MMA's M.  Multiplied with
// W1 the rows 0..7 (c0, c1) are the colour sums and rows 8..15 are discarded; multiplied with W3 it is the other way round.
__device__ __forceinline__ void flush_slots(uint32_t s_warp, int lane, uint32_t ns, float cx, float cy, float* __restrict__ gacc) {
  __syncwarp();
  const uint32_t q = (uint32_t)lane >> 2, t = (uint32_t)lane & 3;
  const uint32_t w_off = s_warp + (q * MG_STRIDE + 8 * t) * 4;
  // pixel pair p = 4 t + s sits at 16-byte chunk p + (p >> 3) of its row (= 4 t + (t >> 1) + s): with the row stride of 18
  // chunks the eight lanes of a quarter warp (two rows x four t) then hit eight different bank groups
  const uint32_t f_off = s_warp + MG_F_OFF + (min(q, (uint32_t)(MG_F_ROWS - 1)) * MG_F_STRIDE + 16 * t + 4 * (t >> 1)) * 4;
  float d1[4] = {0.0f, 0.0f, 0.0f, 0.0f}, d3[4] = {0.0f, 0.0f, 0.0f, 0.0f};
  const float4 w1a = lds128(w_off), w1b = lds128(w_off + 16);
  const float4 w3a = lds128(w_off + MG_TILE_BYTES), w3b = lds128(w_off + MG_TILE_BYTES + 16);
  const float w1[8] = {w1a.x, w1a.y, w1a.z, w1a.w, w1b.x, w1b.y, w1b.z, w1b.w};
  const float w3[8] = {w3a.x, w3a.y, w3a.z, w3a.w, w3b.x, w3b.y, w3b.z, w3b.w};
#pragma unroll
  for (int s = 0; s < 4; s++) {
    const float4 a = lds128(f_off + 16 * s);
    const float g0 = tf32_hi(w1[2 * s]), g1 = tf32_hi(w1[2 * s + 1]);
    mma_m16n8k8_tf32(d1, a, g0, g1);
    mma_m16n8k8_tf32(d1, a, w1[2 * s] - g0, w1[2 * s + 1] - g1);
    const float h0 = tf32_hi(w3[2 * s]), h1 = tf32_hi(w3[2 * s + 1]);
    mma_m16n8k8_tf32(d3, a, h0, h1);
    mma_m16n8k8_tf32(d3, a, w3[2 * s] - h0, w3[2 * s + 1] - h1);
  }
  __syncwarp();   // every lane has its W1 fragments: the W1 tile now becomes M[slot][16] (stride MG_M_STRIDE)
  {
    const uint32_t r0 = s_warp + ((2 * t) * MG_M_STRIDE + q) * 4, r1 = r0 + MG_M_STRIDE * 4;
    sts32(r0, __float_as_uint(d3[2]));        // rows 8..15 of FP|F1 x W3: moment q of slots 2 t, 2 t + 1
    sts32(r1, __float_as_uint(d3[3]));
    sts32(r0 + 32, __float_as_uint(d1[0]));   // rows 0..7 of F1 x W1: colour row q (0..2 hi parts, 3..5 lo parts)
    sts32(r1 + 32, __float_as_uint(d1[1]));
  }
  __syncwarp();
  const uint32_t slot = q, part = t;
  const uint32_t row = s_warp + slot * (MG_M_STRIDE * 4);
  const float4 m0 = lds128(row), m1 = lds128(row + 16), m2 = lds128(row + 32), m3 = lds128(row + 48);
  const float4 inf = lds128(s_warp + MG_INFO_OFF + slot * 16);                       // x, y, opacity, id
  const float S0 = m0.x, Su = m0.y, Sv = m0.z, Suu = m0.w, Suv = m1.x, Svv = m1.y;
  const float X = inf.x - cx, Y = inf.y - cy, op = inf.z, nh = -0.5f * inf.z;
  const float ax = fmaf(X, S0, -Su), ay = fmaf(Y, S0, -Sv);                          // sum W3 dx, sum W3 dy
  // branch-free: part 0..2 -> (dL/dcolor_part, [3 + part]), part 3 -> ([6], [7]) and [8]
  const float col = (part == 0 ? m2.x : (part == 1 ? m2.y : m2.z)) + (part == 0 ? m2.w : (part == 1 ? m3.x : m3.y));
  const float xy = nh * (fmaf(X, ay, Suv) - Y * Su);                                 // -0.5 sum w dx dy
  const float xx = nh * fmaf(X, ax - Su, Suu), yy = nh * fmaf(Y, ay - Sv, Svv);      // -0.5 sum w dx^2, -0.5 sum w dy^2
  const float va = part == 3 ? xy : col;
  const float vb = part == 0 ? op * ax : (part == 1 ? op * ay : (part == 2 ? xx : yy));
  const uint32_t oa = part == 3 ? 6u : part, ob = part == 3 ? 7u : 3u + part;
  if (slot < ns) {
    float* g = gacc + (size_t)__float_as_uint(inf.w) * 12;
    atomicAdd(g + oa, va);
    atomicAdd(g + ob, vb);
    if (part == 3) atomicAdd(g + 8, S0);                                             // dL/dopacity
  }
  __syncwarp();   // M consumed before the next hits overwrite the W1 tile
}

bool g_bwd_mma = true;   // gsr_debug_set knob 3: 0 = the shuffle-butterfly reduction (round-1 kernel), for A/B timing

struct BlendBwdArgs {
  BlendBwdView v[GSR_MAX_BATCH];
};

// blockIdx.y = view of a batched step
template <bool PRECISE, bool MMA>
__global__ void __launch_bounds__(256, MMA ? 4 : 0)
blend_backward_kernel(const __grid_constant__ BlendBwdArgs args) {
  static_assert(!(PRECISE && MMA), "the parity build keeps the oracle's op order: no tensor-core contraction");
  const BlendBwdView& a = args.v[blockIdx.y];
  const int W = a.W, H = a.H, grid_x = a.grid_x;
  if ((int)blockIdx.x >= a.G) return;
  const uint2* __restrict__ ranges = a.ranges;
  const uint32_t* __restrict__ point_list = a.point_list;
  const float4* __restrict__ rec = a.rec;
  const float* __restrict__ bg = a.bg;
  const float* __restrict__ final_T = a.final_T;
  const uint32_t* __restrict__ n_contrib = a.n_contrib;
  const float* __restrict__ dL_dpix = a.dL_dpix;
  float* __restrict__ gacc = a.gacc;
  __shared__ __align__(16) unsigned char s_entries[BLEND_BATCH * ENTRY_BYTES];
  __shared__ uint32_t s_mask_arr[64];
  __shared__ uint32_t s_wmax[8];
  const uint32_t s_ent = pin_reg((uint32_t)__cvta_generic_to_shared(s_entries));
  const uint32_t s_mask = pin_reg((uint32_t)__cvta_generic_to_shared(s_mask_arr));

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
  const float neg_Tf_bg = -T_final * bg_dot;
  float acc0 = 0.0f, acc1 = 0.0f, acc2 = 0.0f;   // accum_rec
  float lc0 = 0.0f, lc1 = 0.0f, lc2 = 0.0f;      // last_color
  float last_alpha = 0.0f;
  const bool red_lane = (lane & 3) == 0 || lane == 2;
  const uint32_t red_off = lane == 2 ? 8u : (uint32_t)(lane >> 2);

  // tensor-core path: the warp's W tiles and its feature table {F1 = this warp's dL_dpixel (rows 0..2 hi, 3..5 lo),
  // FP = monomials of the pixel's offset from the sub-tile centre}, interleaved per pixel
  uint32_t s_warp = 0, w_addr = 0, i_addr = 0, ns = 0;
  float sub_cx = 0.0f, sub_cy = 0.0f;
  if constexpr (MMA) {
    __shared__ __align__(16) unsigned char s_mma[8 * MG_WARP_BYTES];
    s_warp = pin_reg((uint32_t)__cvta_generic_to_shared(s_mma) + warp * MG_WARP_BYTES);
    const uint32_t f = s_warp + MG_F_OFF + lane * 8 + (lane >> 4) * 16;   // pair p = lane >> 1 at chunk p + (p >> 3)
    const float h0 = tf32_hi(dLp0), h1 = tf32_hi(dLp1), h2 = tf32_hi(dLp2);
    const float u = (float)(lane & 7) - 3.5f, v = (float)(lane >> 3) - 1.5f;
    const float f1[MG_F_ROWS] = {h0, h1, h2, dLp0 - h0, dLp1 - h1, dLp2 - h2};
    const float fp[MG_F_ROWS] = {1.0f, u, v, u * u, u * v, v * v};
#pragma unroll
    for (int r = 0; r < MG_F_ROWS; r++) {
      sts32(f + r * MG_F_STRIDE * 4, __float_as_uint(f1[r]));
      sts32(f + r * MG_F_STRIDE * 4 + 4, __float_as_uint(fp[r]));
    }
    w_addr = s_warp + lane * 4;
    i_addr = s_warp + MG_INFO_OFF;
    sub_cx = (float)(tile_x0 + (warp & 1) * 8) + 3.5f;
    sub_cy = (float)(tile_y0 + (warp >> 1) * 4) + 1.5f;
  }

  const uint32_t warp_last = __reduce_max_sync(0xffffffffu, last);
  if (lane == 0) s_wmax[warp] = warp_last;
  __syncthreads();
  uint32_t cta_last = 0;
#pragma unroll
  for (int w = 0; w < 8; w++) cta_last = max(cta_last, s_wmax[w]);
  if (cta_last == 0) return;  // nothing contributed anywhere in this tile (uniform exit)
  const int nb = (int)((cta_last + BLEND_BATCH - 1) / BLEND_BATCH);

  for (int b = nb - 1; b >= 0; b--) {
    __syncthreads();  // WAR on the staging buffer
    const uint32_t pos = (uint32_t)b * BLEND_BATCH + tid;
    const uint32_t bits = stage_entry<PRECISE, GSR_REFINE_BWD != 0>(pos < cta_last, range.x + pos, point_list, rec,
                                               s_ent + tid * ENTRY_BYTES, (float)tile_x0, (float)tile_y0);
    publish_masks<false>(bits, s_mask, warp, lane);
    __syncthreads();

    if (warp_last > (uint32_t)b * BLEND_BATCH) {
#pragma unroll 1
      for (int ws = 7; ws >= 0; ws--) {
        unsigned m = lds32(s_mask + (warp * 8 + ws) * 4);
        const uint32_t gbase = (uint32_t)b * BLEND_BATCH + ws * 32;
        if (gbase >= warp_last) continue;
        if (warp_last - gbase < 32) m &= (1u << (warp_last - gbase)) - 1;  // entries >= warp_last
        const uint32_t ebase = s_ent + ws * 32 * ENTRY_BYTES;
        // entry j of the group is list position gbase + j (0-based == contributor index): it lies in front of this
        // pixel's last contributor iff j < last - gbase (negative when the pixel stopped before this group)
        const int last_rel = (int)last - (int)gbase;
        while (m) {
          const int j = bfind(m);
          m &= bits_below(j);   // j is the top set bit: clears it (BMSK + LOP3)
          const uint32_t ea = ebase + j * ENTRY_BYTES;
          const float4 e0 = lds128(ea);
          const float4 e1 = lds128(ea + 16);
          float dx, dy;
          const float power = pair_power<PRECISE>(e0, e1, pxf, pyf, dx, dy);
          bool contrib = (j < last_rel) && !(power > 0.0f || power < e1.y);
          float G, alpha;
          if constexpr (MMA) {
            // Unconditional: a lane that fails the range test computes garbage (inf, NaN -> fminf gives 0.99) that the
            // selects below turn into exact zeros -- no predicated chain, no zero initialisation, and NO warp vote: after
            // the exact ellipse cull at staging ~97 % of the hits that reach this point have a contributing lane, so
            // skipping the rest costs more (vote + branch + the re-convergence the compiler wraps around it, every hit)
            // than letting them through as one more all-zero slot of the contraction.
            G = pair_gauss<PRECISE>(power);
            alpha = fminf(0.99f, MUL(e1.z, G));
            contrib = contrib && !(alpha < 1.0f / 255.0f);
          } else {
            G = 0.0f;
            alpha = 0.0f;
            if (contrib) {
              G = pair_gauss<PRECISE>(power);
              alpha = fminf(0.99f, MUL(e1.z, G));
              contrib = !(alpha < 1.0f / 255.0f);
            }
            if (!__any_sync(0xffffffffu, contrib)) continue;
          }

          const float4 e2 = lds128(ea + 32);
          float dch, dL_dalpha;
          if (PRECISE) {
            // the oracle's op order; lanes that do not contribute add exact zeros
            const float one_m_alpha = SUB(1.0f, alpha);
            const float Tn = DIV(T, one_m_alpha);
            const float dchT = MUL(DIV(-T_final, one_m_alpha), bg_dot);  // -T_final/(1-alpha)*<bg, dL_dpixel>
            dch = contrib ? MUL(alpha, Tn) : 0.0f;
            const float na0 = FMA(last_alpha, lc0, MUL(SUB(1.0f, last_alpha), acc0));
            const float na1 = FMA(last_alpha, lc1, MUL(SUB(1.0f, last_alpha), acc1));
            const float na2 = FMA(last_alpha, lc2, MUL(SUB(1.0f, last_alpha), acc2));
            dL_dalpha = MUL(SUB(e2.x, na0), dLp0);
            dL_dalpha = FMA(SUB(e2.y, na1), dLp1, dL_dalpha);
            dL_dalpha = FMA(SUB(e2.z, na2), dLp2, dL_dalpha);
            dL_dalpha = FMA(dL_dalpha, Tn, dchT);
            if (contrib) {
              T = Tn;
              acc0 = na0; acc1 = na1; acc2 = na2;
              lc0 = e2.x; lc1 = e2.y; lc2 = e2.z;
              last_alpha = alpha;
            } else {
              dL_dalpha = 0.0f;
            }
          } else {
            // Branch-free: a lane that does not contribute runs the same arithmetic with alpha = G = 0,
            // which leaves its state untouched (rcp.approx(1) == 1 exactly, R + 0 * d == R) and makes
            // every partial below an exact zero -- no predicated state moves.  The colour behind the
            // current Gaussian is carried as ONE value per channel, R = last_alpha * last_color +
            // (1 - last_alpha) * accum_rec, updated as R += alpha * (c - R) (same recurrence as A.6).
            if (!contrib) { alpha = 0.0f; G = 0.0f; }
            const float inv = rcp_approx(1.0f - alpha);
            const float d0 = e2.x - acc0, d1 = e2.y - acc1, d2 = e2.z - acc2;
            T *= inv;
            dL_dalpha = fmaf(fmaf(d2, dLp2, fmaf(d1, dLp1, d0 * dLp0)), T, neg_Tf_bg * inv);
            acc0 = fmaf(alpha, d0, acc0);
            acc1 = fmaf(alpha, d1, acc1);
            acc2 = fmaf(alpha, d2, acc2);
            dch = alpha * T;
          }
          if constexpr (MMA) {
            // this hit becomes slot `ns` of the warp's tiles: W1 = alpha * T, W3 = G * dL_dalpha (exact zeros for lanes
            // that do not contribute), plus the Gaussian's (x, y, opacity, id) for the epilogue
            sts32(w_addr, __float_as_uint(dch));
            sts32(w_addr + MG_TILE_BYTES, __float_as_uint(G * dL_dalpha));
            if (lane == 0) {   // {x, y} and {opacity, id} are register pairs of the two entry loads: two STS.64, no moves
              sts64(i_addr, e0.x, e0.y);
              sts64(i_addr + 8, e1.z, e1.w);
            }
            w_addr += MG_STRIDE * 4;
            i_addr += 16;
            if (++ns == MG_SLOTS) {
              flush_slots(s_warp, lane, MG_SLOTS, sub_cx, sub_cy, gacc);
              ns = 0;
              w_addr = s_warp + lane * 4;
              i_addr = s_warp + MG_INFO_OFF;
            }
            continue;
          }
          const float w = MUL(MUL(e1.z, dL_dalpha), G);  // dL_dG * G
          const float wx = MUL(w, dx), wy = MUL(w, dy);
          const float r9 = warp_transpose_reduce9(MUL(dch, dLp0), MUL(dch, dLp1), MUL(dch, dLp2), wx, wy,
                                                  MUL(-0.5f * dx, wx), MUL(-0.5f * dy, wx), MUL(-0.5f * dy, wy),
                                                  MUL(G, dL_dalpha), lane);
          // ONE predicated RED per (warp, Gaussian): lanes 0,4,..,28 carry values 0..7, lane 2 carries value 8
          if (red_lane) atomicAdd(gacc + (size_t)__float_as_uint(e1.w) * 12 + red_off, r9);
        }
      }
    }
  }
  if constexpr (MMA) {
    if (ns) flush_slots(s_warp, lane, ns, sub_cx, sub_cy, gacc);   // the warp's last, partial group (warp-uniform)
  }
}

cudaError_t launch_blend_backward(cudaStream_t s, const BlendBwdView* views, int nv, bool precise) {
  if (nv <= 0) return cudaSuccess;
  if (nv > GSR_MAX_BATCH) return cudaErrorInvalidValue;
  BlendBwdArgs args{};
  int max_g = 0;
  for (int k = 0; k < nv; k++) {
    args.v[k] = views[k];
    args.v[k].grid_x = cdiv(views[k].W, TILE_X);
    args.v[k].G = args.v[k].grid_x * cdiv(views[k].H, TILE_Y);
    max_g = args.v[k].G > max_g ? args.v[k].G : max_g;
  }
  if (max_g == 0) return cudaSuccess;
  if (precise) {
    blend_backward_kernel<true, false><<<dim3((unsigned)max_g, (unsigned)nv), 256, 0, s>>>(args);
  } else if (g_bwd_mma) {
    static bool carveout_set = false;   // 42 KB of static shared memory per CTA, four CTAs per SM: ask for the large carve-out
    if (!carveout_set) {
      cudaFuncSetAttribute(blend_backward_kernel<false, true>, cudaFuncAttributePreferredSharedMemoryCarveout,
                           (int)cudaSharedmemCarveoutMaxShared);
      carveout_set = true;
    }
    blend_backward_kernel<false, true><<<dim3((unsigned)max_g, (unsigned)nv), 256, 0, s>>>(args);
  } else {
    blend_backward_kernel<false, false><<<dim3((unsigned)max_g, (unsigned)nv), 256, 0, s>>>(args);
  }
  count_launch();
  return cudaGetLastError();
}

// diagnostics (gsr_debug_approx_units): the exact identities the branch-free fast path relies on
__global__ void debug_approx_units_kernel(const float* __restrict__ x, int n, float* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  out[2 * i] = rcp_approx(x[i]);
  out[2 * i + 1] = ex2_approx(x[i]);
}
cudaError_t launch_debug_approx_units(cudaStream_t s, const float* x, int n, float* out) {
  debug_approx_units_kernel<<<cdiv(n, 128), 128, 0, s>>>(x, n, out);
  count_launch();
  return cudaGetLastError();
}

}  // namespace gsr
