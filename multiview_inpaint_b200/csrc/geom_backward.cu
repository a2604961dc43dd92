// geom_backward.cu -- K8 (computeCov2D backward) + K9 (preprocess backward), fused.
//
// Restates SURVEY.md Appendix A.7.  One thread per Gaussian; consumes the packed accumulator
// K7 filled and writes EVERY output element (zeros for culled Gaussians), so callers never have to
// memset the (P,M,3) SH gradient.  Gradient conventions follow the public rasterizer:
//   * dL/dmean2D in NDC-scaled units (x 0.5 W, x 0.5 H)  -- what densify_grad_threshold is
//     calibrated against (gs-simp/arguments/__init__.py:93, scene/gaussian_model.py:482-484);
//   * frustum clamp masks (x_grad_mul / y_grad_mul);
//   * dL/dscale is d/d(scale_modifier * scale) (the modifier is not re-applied);
//   * dL/drot has no normalisation Jacobian (F.normalize is outside, gaussian_model.py:41).
// SH derivative polynomial: derivative of gs-simp/utils/sh_utils.py:74-100.
#include "common.cuh"
#include "sh_rows.cuh"

namespace gsr {

// `row` is this Gaussian's (M,3) SH block staged in shared memory: read as sh, then overwritten in
// place with dL/dsh (coefficients above the active degree get zero).
template <int DEG>
__device__ __forceinline__ void sh_backward(float* row, int M, float x, float y, float z,
                                            const float dRGB[3], float ddir[3]) {
  constexpr int NCO = (DEG + 1) * (DEG + 1);
  float w[NCO], dwx[NCO], dwy[NCO], dwz[NCO];
#pragma unroll
  for (int k = 0; k < NCO; k++) { dwx[k] = 0.f; dwy[k] = 0.f; dwz[k] = 0.f; }
  w[0] = GSR_SH_C0;
  if (DEG > 0) {
    w[1] = -GSR_SH_C1 * y; w[2] = GSR_SH_C1 * z; w[3] = -GSR_SH_C1 * x;
    dwy[1] = -GSR_SH_C1; dwz[2] = GSR_SH_C1; dwx[3] = -GSR_SH_C1;
  }
  if (DEG > 1) {
    const float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
    w[4] = GSR_SH_C2_0 * xy; w[5] = GSR_SH_C2_1 * yz; w[6] = GSR_SH_C2_2 * (2.f * zz - xx - yy);
    w[7] = GSR_SH_C2_3 * xz; w[8] = GSR_SH_C2_4 * (xx - yy);
    dwx[4] = GSR_SH_C2_0 * y;        dwy[4] = GSR_SH_C2_0 * x;
    dwy[5] = GSR_SH_C2_1 * z;        dwz[5] = GSR_SH_C2_1 * y;
    dwx[6] = GSR_SH_C2_2 * -2.f * x; dwy[6] = GSR_SH_C2_2 * -2.f * y; dwz[6] = GSR_SH_C2_2 * 4.f * z;
    dwx[7] = GSR_SH_C2_3 * z;        dwz[7] = GSR_SH_C2_3 * x;
    dwx[8] = GSR_SH_C2_4 * 2.f * x;  dwy[8] = GSR_SH_C2_4 * -2.f * y;
    if (DEG > 2) {
      w[9] = GSR_SH_C3_0 * y * (3.f * xx - yy);
      w[10] = GSR_SH_C3_1 * xy * z;
      w[11] = GSR_SH_C3_2 * y * (4.f * zz - xx - yy);
      w[12] = GSR_SH_C3_3 * z * (2.f * zz - 3.f * xx - 3.f * yy);
      w[13] = GSR_SH_C3_4 * x * (4.f * zz - xx - yy);
      w[14] = GSR_SH_C3_5 * z * (xx - yy);
      w[15] = GSR_SH_C3_6 * x * (xx - 3.f * yy);
      dwx[9] = GSR_SH_C3_0 * 6.f * xy;                 dwy[9] = GSR_SH_C3_0 * 3.f * (xx - yy);
      dwx[10] = GSR_SH_C3_1 * yz; dwy[10] = GSR_SH_C3_1 * xz; dwz[10] = GSR_SH_C3_1 * xy;
      dwx[11] = GSR_SH_C3_2 * -2.f * xy;               dwy[11] = GSR_SH_C3_2 * (4.f * zz - xx - 3.f * yy);
      dwz[11] = GSR_SH_C3_2 * 8.f * yz;
      dwx[12] = GSR_SH_C3_3 * -6.f * xz;               dwy[12] = GSR_SH_C3_3 * -6.f * yz;
      dwz[12] = GSR_SH_C3_3 * 3.f * (2.f * zz - xx - yy);
      dwx[13] = GSR_SH_C3_4 * (4.f * zz - 3.f * xx - yy); dwy[13] = GSR_SH_C3_4 * -2.f * xy;
      dwz[13] = GSR_SH_C3_4 * 8.f * xz;
      dwx[14] = GSR_SH_C3_5 * 2.f * xz;                dwy[14] = GSR_SH_C3_5 * -2.f * yz;
      dwz[14] = GSR_SH_C3_5 * (xx - yy);
      dwx[15] = GSR_SH_C3_6 * 3.f * (xx - yy);         dwy[15] = GSR_SH_C3_6 * -6.f * xy;
    }
  }
  ddir[0] = ddir[1] = ddir[2] = 0.f;
#pragma unroll
  for (int k = 0; k < NCO; k++) {
    if (k > 0) {
      const float s = row[3 * k] * dRGB[0] + row[3 * k + 1] * dRGB[1] + row[3 * k + 2] * dRGB[2];
      ddir[0] += dwx[k] * s;
      ddir[1] += dwy[k] * s;
      ddir[2] += dwz[k] * s;
    }
    row[3 * k + 0] = w[k] * dRGB[0];
    row[3 * k + 1] = w[k] * dRGB[1];
    row[3 * k + 2] = w[k] * dRGB[2];
  }
  for (int k = 3 * NCO; k < 3 * M; k++) row[k] = 0.f;
}

// Cooperative, fully coalesced transfer between the block's contiguous span of (rows x R) floats in
// global memory and the padded shared-memory tile (row stride R + 1 -> per-thread rows are bank
// conflict free).  Rows whose flag is 0 are never touched in global memory unless MODE == STORE_ALL.
enum { SPAN_LOAD = 0, SPAN_ADD = 1, SPAN_STORE_ALL = 2 };
template <int MODE>
__device__ __forceinline__ void span_xfer(float* __restrict__ g, float* s_tile, const uint8_t* s_vis,
                                          int rows, int R) {
  const int n = rows * R;
  const int n4 = n >> 2;
  float4* g4 = reinterpret_cast<float4*>(g);
  constexpr int U = 4;  // independent 16-byte global accesses in flight per thread
  for (int base = 0; base < n4; base += U * blockDim.x) {
    int row[U], col[U];
    bool act[U];
    float4 q[U];
#pragma unroll
    for (int u = 0; u < U; u++) {
      const int e4 = base + u * blockDim.x + threadIdx.x;
      act[u] = e4 < n4;
      row[u] = act[u] ? (4 * e4) / R : 0;
      col[u] = 4 * e4 - row[u] * R;
      if (act[u] && MODE != SPAN_STORE_ALL) {
        const int r_last = (4 * e4 + 3) / R;
        bool any = false;
        for (int r = row[u]; r <= r_last; r++) any |= (s_vis[r] != 0);
        act[u] = any;
      }
      q[u] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    if (MODE != SPAN_STORE_ALL) {
#pragma unroll
      for (int u = 0; u < U; u++)
        if (act[u]) q[u] = g4[base + u * blockDim.x + threadIdx.x];
    }
#pragma unroll
    for (int u = 0; u < U; u++) {
      if (!act[u]) continue;
      float v[4] = {q[u].x, q[u].y, q[u].z, q[u].w};
      int r = row[u], c = col[u];
#pragma unroll
      for (int k = 0; k < 4; k++) {
        float* sp = s_tile + r * (R + 1) + c;
        if (MODE == SPAN_LOAD) *sp = v[k];
        else if (MODE == SPAN_ADD) v[k] += s_vis[r] ? *sp : 0.f;
        else v[k] = s_vis[r] ? *sp : 0.f;
        if (++c == R) { c = 0; r++; }
      }
      if (MODE != SPAN_LOAD) g4[base + u * blockDim.x + threadIdx.x] = make_float4(v[0], v[1], v[2], v[3]);
    }
  }
  for (int e = 4 * n4 + threadIdx.x; e < n; e += blockDim.x) {  // tail when rows*R % 4 != 0
    const int row = e / R, col = e - row * R;
    float* sp = s_tile + row * (R + 1) + col;
    if (MODE == SPAN_LOAD) { if (s_vis[row]) *sp = g[e]; }
    else if (MODE == SPAN_ADD) { if (s_vis[row]) g[e] += *sp; }
    else g[e] = s_vis[row] ? *sp : 0.f;
  }
}


// SH backward on a 16-byte aligned staged row (M = 4 or 16): three float4 = four coefficients at a
// time, sh in, dL/dsh out in place.  Same polynomial as sh_backward<DEG> above.
template <int DEG, int MT>
__device__ __forceinline__ void sh_backward_rows(float* row, float x, float y, float z, const float dRGB[3], float ddir[3]) {
  constexpr int NCO = (DEG + 1) * (DEG + 1);
  float w[16], dwx[16], dwy[16], dwz[16];
#pragma unroll
  for (int k = 0; k < 16; k++) { w[k] = 0.f; dwx[k] = 0.f; dwy[k] = 0.f; dwz[k] = 0.f; }
  w[0] = GSR_SH_C0;
  if (DEG > 0) {
    w[1] = -GSR_SH_C1 * y; w[2] = GSR_SH_C1 * z; w[3] = -GSR_SH_C1 * x;
    dwy[1] = -GSR_SH_C1; dwz[2] = GSR_SH_C1; dwx[3] = -GSR_SH_C1;
  }
  if (DEG > 1) {
    const float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
    w[4] = GSR_SH_C2_0 * xy; w[5] = GSR_SH_C2_1 * yz; w[6] = GSR_SH_C2_2 * (2.f * zz - xx - yy);
    w[7] = GSR_SH_C2_3 * xz; w[8] = GSR_SH_C2_4 * (xx - yy);
    dwx[4] = GSR_SH_C2_0 * y;        dwy[4] = GSR_SH_C2_0 * x;
    dwy[5] = GSR_SH_C2_1 * z;        dwz[5] = GSR_SH_C2_1 * y;
    dwx[6] = GSR_SH_C2_2 * -2.f * x; dwy[6] = GSR_SH_C2_2 * -2.f * y; dwz[6] = GSR_SH_C2_2 * 4.f * z;
    dwx[7] = GSR_SH_C2_3 * z;        dwz[7] = GSR_SH_C2_3 * x;
    dwx[8] = GSR_SH_C2_4 * 2.f * x;  dwy[8] = GSR_SH_C2_4 * -2.f * y;
    if (DEG > 2) {
      w[9] = GSR_SH_C3_0 * y * (3.f * xx - yy);
      w[10] = GSR_SH_C3_1 * xy * z;
      w[11] = GSR_SH_C3_2 * y * (4.f * zz - xx - yy);
      w[12] = GSR_SH_C3_3 * z * (2.f * zz - 3.f * xx - 3.f * yy);
      w[13] = GSR_SH_C3_4 * x * (4.f * zz - xx - yy);
      w[14] = GSR_SH_C3_5 * z * (xx - yy);
      w[15] = GSR_SH_C3_6 * x * (xx - 3.f * yy);
      dwx[9] = GSR_SH_C3_0 * 6.f * xy;                 dwy[9] = GSR_SH_C3_0 * 3.f * (xx - yy);
      dwx[10] = GSR_SH_C3_1 * yz; dwy[10] = GSR_SH_C3_1 * xz; dwz[10] = GSR_SH_C3_1 * xy;
      dwx[11] = GSR_SH_C3_2 * -2.f * xy;               dwy[11] = GSR_SH_C3_2 * (4.f * zz - xx - 3.f * yy);
      dwz[11] = GSR_SH_C3_2 * 8.f * yz;
      dwx[12] = GSR_SH_C3_3 * -6.f * xz;               dwy[12] = GSR_SH_C3_3 * -6.f * yz;
      dwz[12] = GSR_SH_C3_3 * 3.f * (2.f * zz - xx - yy);
      dwx[13] = GSR_SH_C3_4 * (4.f * zz - 3.f * xx - yy); dwy[13] = GSR_SH_C3_4 * -2.f * xy;
      dwz[13] = GSR_SH_C3_4 * 8.f * xz;
      dwx[14] = GSR_SH_C3_5 * 2.f * xz;                dwy[14] = GSR_SH_C3_5 * -2.f * yz;
      dwz[14] = GSR_SH_C3_5 * (xx - yy);
      dwx[15] = GSR_SH_C3_6 * 3.f * (xx - yy);         dwy[15] = GSR_SH_C3_6 * -6.f * xy;
    }
  }
  ddir[0] = ddir[1] = ddir[2] = 0.f;
  float4* row4 = reinterpret_cast<float4*>(row);
  constexpr int GROUPS = MT / 4;              // 4 coefficients (3 float4) per group
  constexpr int GUSED = (NCO + 3) / 4;        // groups holding an active coefficient
#pragma unroll
  for (int gI = 0; gI < GROUPS; gI++) {
    if (gI < GUSED) {
      const float4 a = row4[3 * gI], b = row4[3 * gI + 1], c = row4[3 * gI + 2];
      const float v[12] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w, c.x, c.y, c.z, c.w};
      float o[12];
#pragma unroll
      for (int j = 0; j < 4; j++) {
        const int k = 4 * gI + j;
        if (k > 0 && k < NCO) {
          const float sdot = v[3 * j] * dRGB[0] + v[3 * j + 1] * dRGB[1] + v[3 * j + 2] * dRGB[2];
          ddir[0] += dwx[k] * sdot;
          ddir[1] += dwy[k] * sdot;
          ddir[2] += dwz[k] * sdot;
        }
        const float wk = k < NCO ? w[k] : 0.f;
        o[3 * j] = wk * dRGB[0]; o[3 * j + 1] = wk * dRGB[1]; o[3 * j + 2] = wk * dRGB[2];
      }
      row4[3 * gI] = make_float4(o[0], o[1], o[2], o[3]);
      row4[3 * gI + 1] = make_float4(o[4], o[5], o[6], o[7]);
      row4[3 * gI + 2] = make_float4(o[8], o[9], o[10], o[11]);
    } else {
      const float4 zf = make_float4(0.f, 0.f, 0.f, 0.f);
      row4[3 * gI] = zf; row4[3 * gI + 1] = zf; row4[3 * gI + 2] = zf;
    }
  }
}

// MT = compile-time M for the fast paths (1: no staging, 4 / 16: aligned rows), 0 = generic M.
template <bool ACC, int MT>
__global__ void __launch_bounds__(256)
geom_backward_kernel(int P, int D, int M, const float* __restrict__ means3D,
                     const int32_t* __restrict__ radii, const float* __restrict__ shs,
                     const uint8_t* __restrict__ clamped, const float* __restrict__ scales,
                     const float* __restrict__ rotations, const float* __restrict__ cov3D_precomp,
                     const float* __restrict__ colors_precomp, const Camera cam,
                     const float4* __restrict__ rec, const float* __restrict__ gacc,
                     float* __restrict__ dL_dmean2D, float* __restrict__ dL_dconic,
                     float* __restrict__ dL_dopacity, float* __restrict__ dL_dcolor,
                     float* __restrict__ dL_dmean3D, float* __restrict__ dL_dcov3D,
                     float* __restrict__ dL_dsh, float* __restrict__ dL_dscale,
                     float* __restrict__ dL_drot) {
  __shared__ float s_cam[35];
  __shared__ uint8_t s_vis[256];
  extern __shared__ __align__(16) float s_tile[];  // [256][row_stride(3M)] (fast) / [256][3M + 1] (generic)
  load_camera(cam, s_cam);
  const int tid = threadIdx.x;
  const int block_start = blockIdx.x * blockDim.x;
  const int i = block_start + tid;
  const int rows = min((int)blockDim.x, P - block_start);
  const bool valid = i < P;
  const bool vis = valid && radii[i] > 0;
  const int R = 3 * M;
  const bool use_sh = (colors_precomp == nullptr) && (dL_dsh != nullptr) && (shs != nullptr);
  s_vis[tid] = vis ? 1 : 0;
  __syncthreads();
  constexpr bool STAGED_FAST = (MT == 4 || MT == 16);
  const int nco = (D + 1) * (D + 1);
  if (use_sh) {
    // coefficients above the active degree have no effect on ddir: at degree 0 nothing is read
    if (STAGED_FAST) { if (D > 0) rows_load<3 * (STAGED_FAST ? MT : 4)>(shs + (size_t)block_start * R, s_tile, s_vis, rows); }
    else if (MT == 0) span_xfer<SPAN_LOAD>(const_cast<float*>(shs) + (size_t)block_start * R, s_tile, s_vis, rows, R);
    if (MT != 1) __syncthreads();
  }
  const size_t i3 = 3 * (size_t)i;

  if (valid && !vis) {
    dL_dmean2D[i3] = 0.f; dL_dmean2D[i3 + 1] = 0.f; dL_dmean2D[i3 + 2] = 0.f;
    if (dL_dconic) reinterpret_cast<float4*>(dL_dconic)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (!ACC) {  // accumulate mode: a culled Gaussian adds nothing to the arena
      dL_dopacity[i] = 0.f;
      dL_dcolor[i3] = 0.f; dL_dcolor[i3 + 1] = 0.f; dL_dcolor[i3 + 2] = 0.f;
      dL_dmean3D[i3] = 0.f; dL_dmean3D[i3 + 1] = 0.f; dL_dmean3D[i3 + 2] = 0.f;
#pragma unroll
      for (int k = 0; k < 6; k++) dL_dcov3D[6 * (size_t)i + k] = 0.f;
      dL_dscale[i3] = 0.f; dL_dscale[i3 + 1] = 0.f; dL_dscale[i3 + 2] = 0.f;
      reinterpret_cast<float4*>(dL_drot)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }

  if (vis) {
    // ---- accumulator -> per-Gaussian blend gradients ----
    const float4 g0 = reinterpret_cast<const float4*>(gacc)[3 * (size_t)i + 0];  // color rgb, A
    const float4 g1 = reinterpret_cast<const float4*>(gacc)[3 * (size_t)i + 1];  // B, conic xx, xy, yy
    const float4 g2 = reinterpret_cast<const float4*>(gacc)[3 * (size_t)i + 2];  // opacity
    const float4 q0 = __ldg(rec + 3 * (size_t)i);      // x, y, conic.x, conic.y
    const float4 q1 = __ldg(rec + 3 * (size_t)i + 1);  // conic.z, opacity
    const float A = g0.w, B = g1.x;
    const float gm2x = -(q0.z * A + q0.w * B) * (0.5f * (float)cam.W);
    const float gm2y = -(q1.x * B + q0.w * A) * (0.5f * (float)cam.H);
    const float gcx = g1.y, gcy = g1.z, gcw = g1.w;
    dL_dmean2D[i3] = gm2x; dL_dmean2D[i3 + 1] = gm2y; dL_dmean2D[i3 + 2] = 0.f;
    if (dL_dconic) reinterpret_cast<float4*>(dL_dconic)[i] = make_float4(gcx, gcy, 0.f, gcw);
    if (ACC) {
      dL_dopacity[i] += g2.x;
      if (colors_precomp != nullptr) { dL_dcolor[i3] += g0.x; dL_dcolor[i3 + 1] += g0.y; dL_dcolor[i3 + 2] += g0.z; }
    } else {
      dL_dopacity[i] = g2.x;
      dL_dcolor[i3] = g0.x; dL_dcolor[i3 + 1] = g0.y; dL_dcolor[i3 + 2] = g0.z;
    }

    const float m[3] = {__ldg(means3D + i3), __ldg(means3D + i3 + 1), __ldg(means3D + i3 + 2)};
    const float* V = s_cam;
    const float* PM = s_cam + 16;
    const float* campos = s_cam + 32;

    // ---- 3D covariance (recomputed; K1 does not store it) ----
    float c3[6];
    float R3[3][3], sc[3] = {0.f, 0.f, 0.f};
    float4 quat = make_float4(0.f, 0.f, 0.f, 0.f);
    if (cov3D_precomp != nullptr) {
#pragma unroll
      for (int k = 0; k < 6; k++) c3[k] = __ldg(cov3D_precomp + 6 * (size_t)i + k);
#pragma unroll
      for (int a = 0; a < 3; a++)
#pragma unroll
        for (int k = 0; k < 3; k++) R3[a][k] = 0.f;
    } else {
      quat = __ldg(reinterpret_cast<const float4*>(rotations) + i);
      const float r = quat.x, x = quat.y, y = quat.z, z = quat.w;
      R3[0][0] = 1.f - 2.f * (y * y + z * z); R3[0][1] = 2.f * (x * y - r * z); R3[0][2] = 2.f * (x * z + r * y);
      R3[1][0] = 2.f * (x * y + r * z); R3[1][1] = 1.f - 2.f * (x * x + z * z); R3[1][2] = 2.f * (y * z - r * x);
      R3[2][0] = 2.f * (x * z - r * y); R3[2][1] = 2.f * (y * z + r * x); R3[2][2] = 1.f - 2.f * (x * x + y * y);
      sc[0] = cam.scale_modifier * __ldg(scales + i3);
      sc[1] = cam.scale_modifier * __ldg(scales + i3 + 1);
      sc[2] = cam.scale_modifier * __ldg(scales + i3 + 2);
      float Mx[3][3];
#pragma unroll
      for (int a = 0; a < 3; a++)
#pragma unroll
        for (int k = 0; k < 3; k++) Mx[a][k] = R3[a][k] * sc[k];
      c3[0] = Mx[0][0] * Mx[0][0] + Mx[0][1] * Mx[0][1] + Mx[0][2] * Mx[0][2];
      c3[1] = Mx[0][0] * Mx[1][0] + Mx[0][1] * Mx[1][1] + Mx[0][2] * Mx[1][2];
      c3[2] = Mx[0][0] * Mx[2][0] + Mx[0][1] * Mx[2][1] + Mx[0][2] * Mx[2][2];
      c3[3] = Mx[1][0] * Mx[1][0] + Mx[1][1] * Mx[1][1] + Mx[1][2] * Mx[1][2];
      c3[4] = Mx[1][0] * Mx[2][0] + Mx[1][1] * Mx[2][1] + Mx[1][2] * Mx[2][2];
      c3[5] = Mx[2][0] * Mx[2][0] + Mx[2][1] * Mx[2][1] + Mx[2][2] * Mx[2][2];
    }

    float dmean[3];
    float dcov[6];
    // ---- K8: cov2D backward ----
    {
      float pv[3];
#pragma unroll
      for (int r = 0; r < 3; r++) pv[r] = V[r] * m[0] + V[4 + r] * m[1] + V[8 + r] * m[2] + V[12 + r];
      const float limx = 1.3f * cam.tan_fovx, limy = 1.3f * cam.tan_fovy;
      const float tz = pv[2];
      const float txtz = pv[0] / tz, tytz = pv[1] / tz;
      const float tx = fminf(limx, fmaxf(-limx, txtz)) * tz;
      const float ty = fminf(limy, fmaxf(-limy, tytz)) * tz;
      const float x_grad_mul = (txtz < -limx || txtz > limx) ? 0.f : 1.f;
      const float y_grad_mul = (tytz < -limy || tytz > limy) ? 0.f : 1.f;
      const float J00 = cam.focal_x / tz, J02 = -(cam.focal_x * tx) / (tz * tz);
      const float J11 = cam.focal_y / tz, J12 = -(cam.focal_y * ty) / (tz * tz);
      float T0[3], T1[3], v0[3], v1[3];
#pragma unroll
      for (int c = 0; c < 3; c++) {
        T0[c] = J00 * V[4 * c + 0] + J02 * V[4 * c + 2];
        T1[c] = J11 * V[4 * c + 1] + J12 * V[4 * c + 2];
      }
      const float S[3][3] = {{c3[0], c3[1], c3[2]}, {c3[1], c3[3], c3[4]}, {c3[2], c3[4], c3[5]}};
#pragma unroll
      for (int a = 0; a < 3; a++) {
        v0[a] = S[a][0] * T0[0] + S[a][1] * T0[1] + S[a][2] * T0[2];
        v1[a] = S[a][0] * T1[0] + S[a][1] * T1[1] + S[a][2] * T1[2];
      }
      const float a = T0[0] * v0[0] + T0[1] * v0[1] + T0[2] * v0[2] + 0.3f;
      const float b = T0[0] * v1[0] + T0[1] * v1[1] + T0[2] * v1[2];
      const float c = T1[0] * v1[0] + T1[1] * v1[1] + T1[2] * v1[2] + 0.3f;
      const float denom = a * c - b * b;
      const float denom2inv = 1.0f / ((denom * denom) + 0.0000001f);
      float dL_da = 0.f, dL_db = 0.f, dL_dc = 0.f;
      if (denom2inv != 0.f) {
        dL_da = denom2inv * (-c * c * gcx + 2.f * b * c * gcy + (denom - a * c) * gcw);
        dL_dc = denom2inv * (-a * a * gcw + 2.f * a * b * gcy + (denom - a * c) * gcx);
        dL_db = denom2inv * 2.f * (b * c * gcx - (denom + 2.f * b * b) * gcy + a * b * gcw);
        dcov[0] = T0[0] * T0[0] * dL_da + T0[0] * T1[0] * dL_db + T1[0] * T1[0] * dL_dc;
        dcov[3] = T0[1] * T0[1] * dL_da + T0[1] * T1[1] * dL_db + T1[1] * T1[1] * dL_dc;
        dcov[5] = T0[2] * T0[2] * dL_da + T0[2] * T1[2] * dL_db + T1[2] * T1[2] * dL_dc;
        dcov[1] = 2.f * T0[0] * T0[1] * dL_da + (T0[0] * T1[1] + T0[1] * T1[0]) * dL_db + 2.f * T1[0] * T1[1] * dL_dc;
        dcov[2] = 2.f * T0[0] * T0[2] * dL_da + (T0[0] * T1[2] + T0[2] * T1[0]) * dL_db + 2.f * T1[0] * T1[2] * dL_dc;
        dcov[4] = 2.f * T0[2] * T0[1] * dL_da + (T0[1] * T1[2] + T0[2] * T1[1]) * dL_db + 2.f * T1[1] * T1[2] * dL_dc;
      } else {
#pragma unroll
        for (int k = 0; k < 6; k++) dcov[k] = 0.f;
      }
      float dT0[3], dT1[3];
#pragma unroll
      for (int k = 0; k < 3; k++) {
        dT0[k] = 2.f * v0[k] * dL_da + v1[k] * dL_db;
        dT1[k] = 2.f * v1[k] * dL_dc + v0[k] * dL_db;
      }
      float dJ00 = 0.f, dJ02 = 0.f, dJ11 = 0.f, dJ12 = 0.f;
#pragma unroll
      for (int k = 0; k < 3; k++) {
        dJ00 += V[4 * k + 0] * dT0[k];
        dJ02 += V[4 * k + 2] * dT0[k];
        dJ11 += V[4 * k + 1] * dT1[k];
        dJ12 += V[4 * k + 2] * dT1[k];
      }
      const float itz = 1.f / tz, itz2 = itz * itz, itz3 = itz2 * itz;
      const float dtx = x_grad_mul * -cam.focal_x * itz2 * dJ02;
      const float dty = y_grad_mul * -cam.focal_y * itz2 * dJ12;
      const float dtz = -cam.focal_x * itz2 * dJ00 - cam.focal_y * itz2 * dJ11 +
                        (2.f * cam.focal_x * tx) * itz3 * dJ02 + (2.f * cam.focal_y * ty) * itz3 * dJ12;
      dmean[0] = V[0] * dtx + V[1] * dty + V[2] * dtz;
      dmean[1] = V[4] * dtx + V[5] * dty + V[6] * dtz;
      dmean[2] = V[8] * dtx + V[9] * dty + V[10] * dtz;
    }
    if (ACC) {
      if (cov3D_precomp != nullptr)
#pragma unroll
        for (int k = 0; k < 6; k++) dL_dcov3D[6 * (size_t)i + k] += dcov[k];
    } else {
#pragma unroll
      for (int k = 0; k < 6; k++) dL_dcov3D[6 * (size_t)i + k] = dcov[k];
    }

    // ---- K9: perspective projection of the mean ----
    {
      const float mw_den = PM[3] * m[0] + PM[7] * m[1] + PM[11] * m[2] + PM[15];
      const float m_w = 1.0f / (mw_den + 0.0000001f);
      const float mul1 = (PM[0] * m[0] + PM[4] * m[1] + PM[8] * m[2] + PM[12]) * m_w * m_w;
      const float mul2 = (PM[1] * m[0] + PM[5] * m[1] + PM[9] * m[2] + PM[13]) * m_w * m_w;
#pragma unroll
      for (int k = 0; k < 3; k++)
        dmean[k] += (PM[4 * k + 0] * m_w - PM[4 * k + 3] * mul1) * gm2x +
                    (PM[4 * k + 1] * m_w - PM[4 * k + 3] * mul2) * gm2y;
    }

    // ---- K9: SH backward on the staged row (sh in, dL/dsh out, in place) ----
    if (use_sh) {
      const float dorig[3] = {m[0] - campos[0], m[1] - campos[1], m[2] - campos[2]};
      const float sum2 = dorig[0] * dorig[0] + dorig[1] * dorig[1] + dorig[2] * dorig[2];
      const float inv_len = 1.0f / sqrtf(sum2);
      const float x = dorig[0] * inv_len, y = dorig[1] * inv_len, z = dorig[2] * inv_len;
      const uint8_t cb = clamped[i];
      const float dRGB[3] = {(cb & 1) ? 0.f : g0.x, (cb & 2) ? 0.f : g0.y, (cb & 4) ? 0.f : g0.z};
      float ddir[3] = {0.f, 0.f, 0.f};
      if (MT == 1) {  // M = 1 (the fork's default sh_degree = 0): three floats per Gaussian, no staging
        float* o = dL_dsh + i3;
        if (ACC) { o[0] += GSR_SH_C0 * dRGB[0]; o[1] += GSR_SH_C0 * dRGB[1]; o[2] += GSR_SH_C0 * dRGB[2]; }
        else { o[0] = GSR_SH_C0 * dRGB[0]; o[1] = GSR_SH_C0 * dRGB[1]; o[2] = GSR_SH_C0 * dRGB[2]; }
      } else if (STAGED_FAST) {
        constexpr int MF = STAGED_FAST ? MT : 4;
        float* row = s_tile + tid * row_stride(3 * MF);
        switch (D) {
          case 0: sh_backward_rows<0, MF>(row, x, y, z, dRGB, ddir); break;
          case 1: sh_backward_rows<1, MF>(row, x, y, z, dRGB, ddir); break;
          case 2: if (MF >= 9) sh_backward_rows<(MF >= 9 ? 2 : 0), MF>(row, x, y, z, dRGB, ddir); break;
          default: if (MF >= 16) sh_backward_rows<(MF >= 16 ? 3 : 0), MF>(row, x, y, z, dRGB, ddir); break;
        }
      } else {
        float* row = s_tile + tid * (R + 1);
        switch (D) {
          case 0: sh_backward<0>(row, M, x, y, z, dRGB, ddir); break;
          case 1: sh_backward<1>(row, M, x, y, z, dRGB, ddir); break;
          case 2: sh_backward<2>(row, M, x, y, z, dRGB, ddir); break;
          default: sh_backward<3>(row, M, x, y, z, dRGB, ddir); break;
        }
      }
      const float invsum32 = 1.0f / sqrtf(sum2 * sum2 * sum2);
      const float* v = dorig;
      dmean[0] += ((sum2 - v[0] * v[0]) * ddir[0] - v[1] * v[0] * ddir[1] - v[2] * v[0] * ddir[2]) * invsum32;
      dmean[1] += (-v[0] * v[1] * ddir[0] + (sum2 - v[1] * v[1]) * ddir[1] - v[2] * v[1] * ddir[2]) * invsum32;
      dmean[2] += (-v[0] * v[2] * ddir[0] - v[1] * v[2] * ddir[1] + (sum2 - v[2] * v[2]) * ddir[2]) * invsum32;
    }
    if (ACC) {
      dL_dmean3D[i3] += dmean[0]; dL_dmean3D[i3 + 1] += dmean[1]; dL_dmean3D[i3 + 2] += dmean[2];
    } else {
      dL_dmean3D[i3] = dmean[0]; dL_dmean3D[i3 + 1] = dmean[1]; dL_dmean3D[i3 + 2] = dmean[2];
    }

    // ---- K9: cov3D backward ----
    if (cov3D_precomp == nullptr) {
      const float Gs[3][3] = {{dcov[0], 0.5f * dcov[1], 0.5f * dcov[2]},
                              {0.5f * dcov[1], dcov[3], 0.5f * dcov[4]},
                              {0.5f * dcov[2], 0.5f * dcov[4], dcov[5]}};
      float dM[3][3], g[3][3];
#pragma unroll
      for (int a = 0; a < 3; a++)
#pragma unroll
        for (int k = 0; k < 3; k++)
          dM[a][k] = 2.0f * (Gs[a][0] * R3[0][k] + Gs[a][1] * R3[1][k] + Gs[a][2] * R3[2][k]) * sc[k];
#pragma unroll
      for (int k = 0; k < 3; k++) {
        const float ds = R3[0][k] * dM[0][k] + R3[1][k] * dM[1][k] + R3[2][k] * dM[2][k];
        if (ACC) dL_dscale[i3 + k] += ds; else dL_dscale[i3 + k] = ds;
      }
#pragma unroll
      for (int a = 0; a < 3; a++)
#pragma unroll
        for (int k = 0; k < 3; k++) g[a][k] = dM[a][k] * sc[k];
      const float r = quat.x, x = quat.y, y = quat.z, z = quat.w;
      float4 dq;
      dq.x = 2.f * z * (g[1][0] - g[0][1]) + 2.f * y * (g[0][2] - g[2][0]) + 2.f * x * (g[2][1] - g[1][2]);
      dq.y = 2.f * y * (g[0][1] + g[1][0]) + 2.f * z * (g[0][2] + g[2][0]) + 2.f * r * (g[2][1] - g[1][2]) - 4.f * x * (g[1][1] + g[2][2]);
      dq.z = 2.f * x * (g[0][1] + g[1][0]) + 2.f * r * (g[0][2] - g[2][0]) + 2.f * z * (g[1][2] + g[2][1]) - 4.f * y * (g[0][0] + g[2][2]);
      dq.w = 2.f * r * (g[1][0] - g[0][1]) + 2.f * x * (g[0][2] + g[2][0]) + 2.f * y * (g[1][2] + g[2][1]) - 4.f * z * (g[0][0] + g[1][1]);
      if (ACC) {
        const float4 o = reinterpret_cast<float4*>(dL_drot)[i];
        dq.x += o.x; dq.y += o.y; dq.z += o.z; dq.w += o.w;
      }
      reinterpret_cast<float4*>(dL_drot)[i] = dq;
    } else if (!ACC) {
      dL_dscale[i3] = 0.f; dL_dscale[i3 + 1] = 0.f; dL_dscale[i3 + 2] = 0.f;
      reinterpret_cast<float4*>(dL_drot)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }

  // ---- dL/dsh write-out: one coalesced pass over the block's contiguous (rows x 3M) span ----
  if (dL_dsh != nullptr && R > 0) {
    float* gspan = dL_dsh + (size_t)block_start * R;
    if (use_sh && MT == 1) {
      if (!ACC && valid && !vis) { gspan[3 * tid] = 0.f; gspan[3 * tid + 1] = 0.f; gspan[3 * tid + 2] = 0.f; }
    } else if (use_sh) {
      __syncthreads();
      if (STAGED_FAST) rows_store<3 * (STAGED_FAST ? MT : 4), ACC>(gspan, s_tile, s_vis, rows, 3 * nco);
      else if (ACC) span_xfer<SPAN_ADD>(gspan, s_tile, s_vis, rows, R);
      else span_xfer<SPAN_STORE_ALL>(gspan, s_tile, s_vis, rows, R);
    } else if (!ACC) {
      for (int e = tid; e < rows * R; e += blockDim.x) gspan[e] = 0.f;
    }
  }
}

__global__ void __launch_bounds__(256)
view_stats_kernel(int P, const int32_t* __restrict__ radii, const float* __restrict__ dL_dmean2D,
                  float* __restrict__ grad_norm_accum, int32_t* __restrict__ visible_count,
                  int32_t* __restrict__ max_radii) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P) return;
  const int32_t r = __ldg(radii + i);
  if (r <= 0) return;
  if (grad_norm_accum) {
    const float gx = __ldg(dL_dmean2D + 3 * (size_t)i), gy = __ldg(dL_dmean2D + 3 * (size_t)i + 1);
    grad_norm_accum[i] += sqrtf(gx * gx + gy * gy);
  }
  if (visible_count) visible_count[i] += 1;
  if (max_radii) max_radii[i] = max(max_radii[i], r);
}

cudaError_t launch_view_stats(cudaStream_t s, int P, const int32_t* radii, const float* dL_dmean2D,
                              float* grad_norm_accum, int32_t* visible_count, int32_t* max_radii) {
  if (P == 0) return cudaSuccess;
  view_stats_kernel<<<cdiv(P, 256), 256, 0, s>>>(P, radii, dL_dmean2D, grad_norm_accum, visible_count, max_radii);
  count_launch();
  return cudaGetLastError();
}

cudaError_t launch_geom_backward(cudaStream_t s, int P, int D, int M, const float* means3D,
                                 const int32_t* radii, const float* shs, const uint8_t* clamped,
                                 const float* scales, const float* rotations,
                                 const float* cov3D_precomp, const float* colors_precomp,
                                 const Camera& cam, const float4* rec, const float* gacc,
                                 float* dL_dmean2D, float* dL_dconic, float* dL_dopacity,
                                 float* dL_dcolor, float* dL_dmean3D, float* dL_dcov3D,
                                 float* dL_dsh, float* dL_dscale, float* dL_drot, bool accumulate) {
  if (P == 0) return cudaSuccess;
  const bool use_sh = colors_precomp == nullptr && dL_dsh != nullptr && shs != nullptr;
  const int mt = (M == 1 || M == 4 || M == 16) ? M : 0;
  const size_t row_floats = mt == 1 ? 0 : (mt ? row_stride(3 * M) : 3 * M + 1);
  const size_t smem = use_sh ? (size_t)256 * row_floats * sizeof(float) : 0;
  using KernT = void (*)(int, int, int, const float*, const int32_t*, const float*, const uint8_t*, const float*,
                         const float*, const float*, const float*, const Camera, const float4*, const float*, float*,
                         float*, float*, float*, float*, float*, float*, float*, float*);
  KernT kern;
  if (accumulate) kern = mt == 16 ? geom_backward_kernel<true, 16> : mt == 4 ? geom_backward_kernel<true, 4>
                       : mt == 1 ? geom_backward_kernel<true, 1> : geom_backward_kernel<true, 0>;
  else kern = mt == 16 ? geom_backward_kernel<false, 16> : mt == 4 ? geom_backward_kernel<false, 4>
            : mt == 1 ? geom_backward_kernel<false, 1> : geom_backward_kernel<false, 0>;
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
  }
  kern<<<cdiv(P, 256), 256, smem, s>>>(
      P, D, M, means3D, radii, shs, clamped, scales, rotations, cov3D_precomp, colors_precomp, cam,
      rec, gacc, dL_dmean2D, dL_dconic, dL_dopacity, dL_dcolor, dL_dmean3D, dL_dcov3D, dL_dsh,
      dL_dscale, dL_drot);
  count_launch();
  return cudaGetLastError();
}

}  // namespace gsr
