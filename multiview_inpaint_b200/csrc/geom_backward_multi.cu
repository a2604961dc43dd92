// geom_backward_multi.cu -- K8 + K9 for up to four views of ONE set of Gaussians in a single pass.
//
// The reference renders one camera per iteration and runs K8 + K9 once per view; the
// view-sharded multi-view step (multiview.py) therefore paid, per view, one read of every visible
// SH row and one read-modify-write of every visible gradient row of the arena.  Per Gaussian the
// parameters (mean, scale, rotation, SH row) do not depend on the view, and the parameter
// gradients are sums over views, so this kernel
//   * reads the per-Gaussian parameters ONCE,
//   * loops over the views' (packed accumulator of K7, blend record, radius, clamp bits, camera),
//   * sums dL/dmean3D, dL/dcov3D, dL/dopacity and the (M,3) SH gradient row on chip -- the cov3D
//     backward (dL/dscale, dL/drot) is linear in dL/dcov3D, so it runs once on the sum --
//   * and WRITES each gradient once (no atomics, no read-modify-write, no arena memset),
//   * plus the per-view densification statistics of gs-simp/scene/gaussian_model.py:482-484 and
//     train.py:115 (sum over views of ||dL/dmean2D_v||, visibility count, max radius).
// Math per (Gaussian, view) = geom_backward.cu; only the order of the
// floating-point additions over views differs from running the single-view kernel per view.
#include "common.cuh"
#include "sh_rows.cuh"

namespace gsr {

constexpr int GM_MAX_VIEWS = 4;

template <int NV>
struct MultiArgs {
  ViewGrad v[NV];
};

// SH basis value and its gradient w.r.t. the (unit) direction for coefficient K.
template <int K>
__device__ __forceinline__ void sh_basis(float x, float y, float z, float& w, float& dx, float& dy, float& dz) {
  const float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
  dx = 0.f; dy = 0.f; dz = 0.f;
  if (K == 0) { w = GSR_SH_C0; }
  else if (K == 1) { w = -GSR_SH_C1 * y; dy = -GSR_SH_C1; }
  else if (K == 2) { w = GSR_SH_C1 * z; dz = GSR_SH_C1; }
  else if (K == 3) { w = -GSR_SH_C1 * x; dx = -GSR_SH_C1; }
  else if (K == 4) { w = GSR_SH_C2_0 * xy; dx = GSR_SH_C2_0 * y; dy = GSR_SH_C2_0 * x; }
  else if (K == 5) { w = GSR_SH_C2_1 * yz; dy = GSR_SH_C2_1 * z; dz = GSR_SH_C2_1 * y; }
  else if (K == 6) { w = GSR_SH_C2_2 * (2.f * zz - xx - yy); dx = GSR_SH_C2_2 * -2.f * x; dy = GSR_SH_C2_2 * -2.f * y; dz = GSR_SH_C2_2 * 4.f * z; }
  else if (K == 7) { w = GSR_SH_C2_3 * xz; dx = GSR_SH_C2_3 * z; dz = GSR_SH_C2_3 * x; }
  else if (K == 8) { w = GSR_SH_C2_4 * (xx - yy); dx = GSR_SH_C2_4 * 2.f * x; dy = GSR_SH_C2_4 * -2.f * y; }
  else if (K == 9) { w = GSR_SH_C3_0 * y * (3.f * xx - yy); dx = GSR_SH_C3_0 * 6.f * xy; dy = GSR_SH_C3_0 * 3.f * (xx - yy); }
  else if (K == 10) { w = GSR_SH_C3_1 * xy * z; dx = GSR_SH_C3_1 * yz; dy = GSR_SH_C3_1 * xz; dz = GSR_SH_C3_1 * xy; }
  else if (K == 11) { w = GSR_SH_C3_2 * y * (4.f * zz - xx - yy); dx = GSR_SH_C3_2 * -2.f * xy; dy = GSR_SH_C3_2 * (4.f * zz - xx - 3.f * yy); dz = GSR_SH_C3_2 * 8.f * yz; }
  else if (K == 12) { w = GSR_SH_C3_3 * z * (2.f * zz - 3.f * xx - 3.f * yy); dx = GSR_SH_C3_3 * -6.f * xz; dy = GSR_SH_C3_3 * -6.f * yz; dz = GSR_SH_C3_3 * 3.f * (2.f * zz - xx - yy); }
  else if (K == 13) { w = GSR_SH_C3_4 * x * (4.f * zz - xx - yy); dx = GSR_SH_C3_4 * (4.f * zz - 3.f * xx - yy); dy = GSR_SH_C3_4 * -2.f * xy; dz = GSR_SH_C3_4 * 8.f * xz; }
  else if (K == 14) { w = GSR_SH_C3_5 * z * (xx - yy); dx = GSR_SH_C3_5 * 2.f * xz; dy = GSR_SH_C3_5 * -2.f * yz; dz = GSR_SH_C3_5 * (xx - yy); }
  else { w = GSR_SH_C3_6 * x * (xx - 3.f * yy); dx = GSR_SH_C3_6 * 3.f * (xx - yy); dy = GSR_SH_C3_6 * -6.f * xy; }
}

// One coefficient group (4 coefficients = 3 float4 of the staged row) for all views: reads the sh
// values, accumulates each view's dL/ddir, and overwrites the group in place with the summed dL/dsh.
template <int G, int NV>
__device__ __forceinline__ void sh_group_backward(float4* row4, int nco, const float (&dir)[NV][3],
                                                  const float (&dRGB)[NV][3], float (&ddir)[NV][3], bool read_sh) {
  float v[12];
  if (read_sh) {
    const float4 a = row4[3 * G], b = row4[3 * G + 1], c = row4[3 * G + 2];
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    v[8] = c.x; v[9] = c.y; v[10] = c.z; v[11] = c.w;
  } else {
#pragma unroll
    for (int k = 0; k < 12; k++) v[k] = 0.f;
  }
  float o[12];
#pragma unroll
  for (int k = 0; k < 12; k++) o[k] = 0.f;
#pragma unroll
  for (int vw = 0; vw < NV; vw++) {
    // a view in which the Gaussian is culled has dRGB == 0: it adds exact zeros
#pragma unroll
    for (int j = 0; j < 4; j++) {
      constexpr int K0 = 4 * G;
      const int k = K0 + j;
      if (k < nco) {
        float w, dx, dy, dz;
        // compile-time coefficient index: the j loop is unrolled
        if (j == 0) sh_basis<K0 + 0>(dir[vw][0], dir[vw][1], dir[vw][2], w, dx, dy, dz);
        else if (j == 1) sh_basis<K0 + 1>(dir[vw][0], dir[vw][1], dir[vw][2], w, dx, dy, dz);
        else if (j == 2) sh_basis<K0 + 2>(dir[vw][0], dir[vw][1], dir[vw][2], w, dx, dy, dz);
        else sh_basis<K0 + 3>(dir[vw][0], dir[vw][1], dir[vw][2], w, dx, dy, dz);
        if (k > 0) {
          const float sdot = v[3 * j] * dRGB[vw][0] + v[3 * j + 1] * dRGB[vw][1] + v[3 * j + 2] * dRGB[vw][2];
          ddir[vw][0] += dx * sdot;
          ddir[vw][1] += dy * sdot;
          ddir[vw][2] += dz * sdot;
        }
        o[3 * j] += w * dRGB[vw][0];
        o[3 * j + 1] += w * dRGB[vw][1];
        o[3 * j + 2] += w * dRGB[vw][2];
      }
    }
  }
  row4[3 * G] = make_float4(o[0], o[1], o[2], o[3]);
  row4[3 * G + 1] = make_float4(o[4], o[5], o[6], o[7]);
  row4[3 * G + 2] = make_float4(o[8], o[9], o[10], o[11]);
}

#ifndef GSR_GM_MINB
#define GSR_GM_MINB 5
#endif
#ifndef GSR_GM_THREADS
#define GSR_GM_THREADS 128
#endif
constexpr int GM_THREADS = GSR_GM_THREADS;   // Gaussians per CTA (128 threads, <= 102 registers, five CTAs per SM: measured
                                             // 0.167 vs 0.174 ms per view for 256 threads at 126 registers; 64 / 128 / 256 alone change nothing)
template <int NV, int MT, bool ACC>
__global__ void __launch_bounds__(GM_THREADS, GSR_GM_MINB)
geom_backward_multi_kernel(int P, int D, const float* __restrict__ means3D, const float* __restrict__ shs,
                           const float* __restrict__ scales, const float* __restrict__ rotations,
                           float scale_modifier, const MultiArgs<NV> args, float* __restrict__ dL_dopacity,
                           float* __restrict__ dL_dmean3D, float* __restrict__ dL_dsh,
                           float* __restrict__ dL_dscale, float* __restrict__ dL_drot,
                           float* __restrict__ grad_norm_accum, int32_t* __restrict__ visible_count,
                           int32_t* __restrict__ max_radii) {
  constexpr int R = 3 * MT;
  constexpr bool STAGED = (MT == 4 || MT == 16);
  __shared__ float s_cam[NV][36];
  __shared__ uint8_t s_vis[GM_THREADS];
  extern __shared__ __align__(16) float s_tile[];  // [GM_THREADS][row_stride(R)] when STAGED
  const int tid = threadIdx.x;
  const int block_start = blockIdx.x * GM_THREADS;
  const int i = block_start + tid;
  const int rows = min(GM_THREADS, P - block_start);
  const bool valid = i < P;
  for (int t = tid; t < NV * 35; t += GM_THREADS) {
    const int vw = t / 35, k = t - vw * 35;
    const ViewGrad& V = args.v[vw];
    s_cam[vw][k] = k < 16 ? __ldg(V.view + k) : (k < 32 ? __ldg(V.proj + (k - 16)) : __ldg(V.campos + (k - 32)));
  }
  int rad[NV];
  bool any = false;
#pragma unroll
  for (int vw = 0; vw < NV; vw++) {
    rad[vw] = valid ? __ldg(args.v[vw].radii + i) : 0;
    any |= rad[vw] > 0;
  }
  s_vis[tid] = any ? 1 : 0;
  __syncthreads();
  const int nco = (D + 1) * (D + 1);
  if (STAGED) {
    if (D > 0) rows_load<STAGED ? R : 12, GM_THREADS>(shs + (size_t)block_start * R, s_tile, s_vis, rows);
    __syncthreads();
  }
  const size_t i3 = 3 * (size_t)i;

  float dmean[3] = {0.f, 0.f, 0.f};
  float dscale[3] = {0.f, 0.f, 0.f};
  float4 dq = make_float4(0.f, 0.f, 0.f, 0.f);
  float dopac = 0.f;
  float dsh0[3] = {0.f, 0.f, 0.f};  // MT == 1 only
  float gnorm = 0.f;
  int nvis = 0, rmax = 0;

  if (any) {
    const float m[3] = {__ldg(means3D + i3), __ldg(means3D + i3 + 1), __ldg(means3D + i3 + 2)};
    const float4 quat = __ldg(reinterpret_cast<const float4*>(rotations) + i);
    const float qr = quat.x, qx = quat.y, qy = quat.z, qz = quat.w;
    float R3[3][3];
    R3[0][0] = 1.f - 2.f * (qy * qy + qz * qz); R3[0][1] = 2.f * (qx * qy - qr * qz); R3[0][2] = 2.f * (qx * qz + qr * qy);
    R3[1][0] = 2.f * (qx * qy + qr * qz); R3[1][1] = 1.f - 2.f * (qx * qx + qz * qz); R3[1][2] = 2.f * (qy * qz - qr * qx);
    R3[2][0] = 2.f * (qx * qz - qr * qy); R3[2][1] = 2.f * (qy * qz + qr * qx); R3[2][2] = 1.f - 2.f * (qx * qx + qy * qy);
    const float sc[3] = {scale_modifier * __ldg(scales + i3), scale_modifier * __ldg(scales + i3 + 1),
                         scale_modifier * __ldg(scales + i3 + 2)};
    float c3[6];
    {
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
    const float S[3][3] = {{c3[0], c3[1], c3[2]}, {c3[1], c3[3], c3[4]}, {c3[2], c3[4], c3[5]}};

    float dcov[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    float dRGB[NV][3], dir[NV][3], ddir[NV][3], inv_len3[NV], sum2v[NV];

#pragma unroll
    for (int vw = 0; vw < NV; vw++) {
      const ViewGrad& VA = args.v[vw];
      dRGB[vw][0] = dRGB[vw][1] = dRGB[vw][2] = 0.f;
      ddir[vw][0] = ddir[vw][1] = ddir[vw][2] = 0.f;
      dir[vw][0] = dir[vw][1] = 0.f; dir[vw][2] = 1.f;
      inv_len3[vw] = 0.f; sum2v[vw] = 1.f;
      if (rad[vw] > 0) {
        const float* V = s_cam[vw];
        const float* PM = s_cam[vw] + 16;
        const float* campos = s_cam[vw] + 32;
        const float4 g0 = __ldg(reinterpret_cast<const float4*>(VA.gacc) + 3 * (size_t)i + 0);  // color rgb, A
        const float4 g1 = __ldg(reinterpret_cast<const float4*>(VA.gacc) + 3 * (size_t)i + 1);  // B, conic xx, xy, yy
        const float4 g2 = __ldg(reinterpret_cast<const float4*>(VA.gacc) + 3 * (size_t)i + 2);  // opacity
        const float4 q0 = __ldg(VA.rec + 3 * (size_t)i);                                        // x, y, conic.x, conic.y
        const float4 q1 = __ldg(VA.rec + 3 * (size_t)i + 1);                                    // conic.z, opacity
        const uint8_t cb = __ldg(VA.clamped + i);
        const float A = g0.w, B = g1.x;
        const float gm2x = -(q0.z * A + q0.w * B) * (0.5f * (float)VA.W);
        const float gm2y = -(q1.x * B + q0.w * A) * (0.5f * (float)VA.H);
        const float gcx = g1.y, gcy = g1.z, gcw = g1.w;
        if (VA.dL_dmean2D) { VA.dL_dmean2D[i3] = gm2x; VA.dL_dmean2D[i3 + 1] = gm2y; VA.dL_dmean2D[i3 + 2] = 0.f; }
        gnorm += sqrtf(gm2x * gm2x + gm2y * gm2y);
        nvis += 1;
        rmax = max(rmax, rad[vw]);
        dopac += g2.x;
        dRGB[vw][0] = (cb & 1) ? 0.f : g0.x;
        dRGB[vw][1] = (cb & 2) ? 0.f : g0.y;
        dRGB[vw][2] = (cb & 4) ? 0.f : g0.z;

        // ---- K8: cov2D backward (geom_backward.cu) ----
        float pv[3];
#pragma unroll
        for (int r = 0; r < 3; r++) pv[r] = V[r] * m[0] + V[4 + r] * m[1] + V[8 + r] * m[2] + V[12 + r];
        const float limx = 1.3f * VA.tan_fovx, limy = 1.3f * VA.tan_fovy;
        const float tz = pv[2];
        const float txtz = pv[0] / tz, tytz = pv[1] / tz;
        const float tx = fminf(limx, fmaxf(-limx, txtz)) * tz;
        const float ty = fminf(limy, fmaxf(-limy, tytz)) * tz;
        const float x_grad_mul = (txtz < -limx || txtz > limx) ? 0.f : 1.f;
        const float y_grad_mul = (tytz < -limy || tytz > limy) ? 0.f : 1.f;
        const float J00 = VA.focal_x / tz, J02 = -(VA.focal_x * tx) / (tz * tz);
        const float J11 = VA.focal_y / tz, J12 = -(VA.focal_y * ty) / (tz * tz);
        float T0[3], T1[3], v0[3], v1[3];
#pragma unroll
        for (int c = 0; c < 3; c++) {
          T0[c] = J00 * V[4 * c + 0] + J02 * V[4 * c + 2];
          T1[c] = J11 * V[4 * c + 1] + J12 * V[4 * c + 2];
        }
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
          dcov[0] += T0[0] * T0[0] * dL_da + T0[0] * T1[0] * dL_db + T1[0] * T1[0] * dL_dc;
          dcov[3] += T0[1] * T0[1] * dL_da + T0[1] * T1[1] * dL_db + T1[1] * T1[1] * dL_dc;
          dcov[5] += T0[2] * T0[2] * dL_da + T0[2] * T1[2] * dL_db + T1[2] * T1[2] * dL_dc;
          dcov[1] += 2.f * T0[0] * T0[1] * dL_da + (T0[0] * T1[1] + T0[1] * T1[0]) * dL_db + 2.f * T1[0] * T1[1] * dL_dc;
          dcov[2] += 2.f * T0[0] * T0[2] * dL_da + (T0[0] * T1[2] + T0[2] * T1[0]) * dL_db + 2.f * T1[0] * T1[2] * dL_dc;
          dcov[4] += 2.f * T0[2] * T0[1] * dL_da + (T0[1] * T1[2] + T0[2] * T1[1]) * dL_db + 2.f * T1[1] * T1[2] * dL_dc;
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
        const float dtx = x_grad_mul * -VA.focal_x * itz2 * dJ02;
        const float dty = y_grad_mul * -VA.focal_y * itz2 * dJ12;
        const float dtz = -VA.focal_x * itz2 * dJ00 - VA.focal_y * itz2 * dJ11 +
                          (2.f * VA.focal_x * tx) * itz3 * dJ02 + (2.f * VA.focal_y * ty) * itz3 * dJ12;
        dmean[0] += V[0] * dtx + V[1] * dty + V[2] * dtz;
        dmean[1] += V[4] * dtx + V[5] * dty + V[6] * dtz;
        dmean[2] += V[8] * dtx + V[9] * dty + V[10] * dtz;

        // ---- K9: perspective projection of the mean ----
        const float mw_den = PM[3] * m[0] + PM[7] * m[1] + PM[11] * m[2] + PM[15];
        const float m_w = 1.0f / (mw_den + 0.0000001f);
        const float mul1 = (PM[0] * m[0] + PM[4] * m[1] + PM[8] * m[2] + PM[12]) * m_w * m_w;
        const float mul2 = (PM[1] * m[0] + PM[5] * m[1] + PM[9] * m[2] + PM[13]) * m_w * m_w;
#pragma unroll
        for (int k = 0; k < 3; k++)
          dmean[k] += (PM[4 * k + 0] * m_w - PM[4 * k + 3] * mul1) * gm2x +
                      (PM[4 * k + 1] * m_w - PM[4 * k + 3] * mul2) * gm2y;

        // direction for the SH phase
        const float d0 = m[0] - campos[0], d1 = m[1] - campos[1], d2 = m[2] - campos[2];
        const float sum2 = d0 * d0 + d1 * d1 + d2 * d2;
        const float inv_len = 1.0f / sqrtf(sum2);
        dir[vw][0] = d0 * inv_len; dir[vw][1] = d1 * inv_len; dir[vw][2] = d2 * inv_len;
        sum2v[vw] = sum2;
        inv_len3[vw] = 1.0f / sqrtf(sum2 * sum2 * sum2);
      }
    }

    // ---- SH backward over all views ----
    if (MT == 1) {
#pragma unroll
      for (int vw = 0; vw < NV; vw++) {
        dsh0[0] += GSR_SH_C0 * dRGB[vw][0]; dsh0[1] += GSR_SH_C0 * dRGB[vw][1]; dsh0[2] += GSR_SH_C0 * dRGB[vw][2];
      }
    } else if (STAGED) {
      float4* row4 = reinterpret_cast<float4*>(s_tile + tid * row_stride(STAGED ? R : 12));
      const bool read_sh = D > 0;
      sh_group_backward<0, NV>(row4, nco, dir, dRGB, ddir, read_sh);
      if (MT >= 8) sh_group_backward<(MT >= 8 ? 1 : 0), NV>(row4, nco, dir, dRGB, ddir, read_sh);
      if (MT >= 12) sh_group_backward<(MT >= 12 ? 2 : 0), NV>(row4, nco, dir, dRGB, ddir, read_sh);
      if (MT >= 16) sh_group_backward<(MT >= 16 ? 3 : 0), NV>(row4, nco, dir, dRGB, ddir, read_sh);
      // dL/ddir_v -> dL/dmean through normalize() (dnormvdv)
#pragma unroll
      for (int vw = 0; vw < NV; vw++) {
        const float l = sqrtf(sum2v[vw]);
        const float v0 = dir[vw][0] * l, v1 = dir[vw][1] * l, v2 = dir[vw][2] * l;  // un-normalised direction
        const float s2 = sum2v[vw], il3 = inv_len3[vw];
        dmean[0] += ((s2 - v0 * v0) * ddir[vw][0] - v1 * v0 * ddir[vw][1] - v2 * v0 * ddir[vw][2]) * il3;
        dmean[1] += (-v0 * v1 * ddir[vw][0] + (s2 - v1 * v1) * ddir[vw][1] - v2 * v1 * ddir[vw][2]) * il3;
        dmean[2] += (-v0 * v2 * ddir[vw][0] - v1 * v2 * ddir[vw][1] + (s2 - v2 * v2) * ddir[vw][2]) * il3;
      }
    }

    // ---- cov3D backward, once, on the summed dL/dcov3D ----
    {
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
      for (int k = 0; k < 3; k++) dscale[k] = R3[0][k] * dM[0][k] + R3[1][k] * dM[1][k] + R3[2][k] * dM[2][k];
#pragma unroll
      for (int a = 0; a < 3; a++)
#pragma unroll
        for (int k = 0; k < 3; k++) g[a][k] = dM[a][k] * sc[k];
      dq.x = 2.f * qz * (g[1][0] - g[0][1]) + 2.f * qy * (g[0][2] - g[2][0]) + 2.f * qx * (g[2][1] - g[1][2]);
      dq.y = 2.f * qy * (g[0][1] + g[1][0]) + 2.f * qz * (g[0][2] + g[2][0]) + 2.f * qr * (g[2][1] - g[1][2]) - 4.f * qx * (g[1][1] + g[2][2]);
      dq.z = 2.f * qx * (g[0][1] + g[1][0]) + 2.f * qr * (g[0][2] - g[2][0]) + 2.f * qz * (g[1][2] + g[2][1]) - 4.f * qy * (g[0][0] + g[2][2]);
      dq.w = 2.f * qr * (g[1][0] - g[0][1]) + 2.f * qx * (g[0][2] + g[2][0]) + 2.f * qy * (g[1][2] + g[2][1]) - 4.f * qz * (g[0][0] + g[1][1]);
    }
  }

  // ---- per-view dL/dmean2D of culled Gaussians (the reference zero-fills it) ----
  if (valid) {
#pragma unroll
    for (int vw = 0; vw < NV; vw++)
      if (rad[vw] <= 0 && args.v[vw].dL_dmean2D) {
        float* o = args.v[vw].dL_dmean2D + i3;
        o[0] = 0.f; o[1] = 0.f; o[2] = 0.f;
      }
  }

  // ---- parameter gradients and statistics: written once (ACC: added; rows no view saw are left alone) ----
  if (valid && (any || !ACC)) {
    // per-view norms summed (gaussian_model.py:483), visibility count (:484), max radius (train.py:115)
    if (grad_norm_accum) grad_norm_accum[i] = ACC ? grad_norm_accum[i] + gnorm : gnorm;
    if (visible_count) visible_count[i] = ACC ? visible_count[i] + nvis : nvis;
    if (max_radii) max_radii[i] = ACC ? max(max_radii[i], rmax) : rmax;
    if (ACC) {
      dL_dopacity[i] += dopac;
      dL_dmean3D[i3] += dmean[0]; dL_dmean3D[i3 + 1] += dmean[1]; dL_dmean3D[i3 + 2] += dmean[2];
      dL_dscale[i3] += dscale[0]; dL_dscale[i3 + 1] += dscale[1]; dL_dscale[i3 + 2] += dscale[2];
      float4 o = reinterpret_cast<float4*>(dL_drot)[i];
      o.x += dq.x; o.y += dq.y; o.z += dq.z; o.w += dq.w;
      reinterpret_cast<float4*>(dL_drot)[i] = o;
      if (MT == 1) { dL_dsh[i3] += dsh0[0]; dL_dsh[i3 + 1] += dsh0[1]; dL_dsh[i3 + 2] += dsh0[2]; }
    } else {
      dL_dopacity[i] = dopac;
      dL_dmean3D[i3] = dmean[0]; dL_dmean3D[i3 + 1] = dmean[1]; dL_dmean3D[i3 + 2] = dmean[2];
      dL_dscale[i3] = dscale[0]; dL_dscale[i3 + 1] = dscale[1]; dL_dscale[i3 + 2] = dscale[2];
      reinterpret_cast<float4*>(dL_drot)[i] = dq;
      if (MT == 1) { dL_dsh[i3] = dsh0[0]; dL_dsh[i3 + 1] = dsh0[1]; dL_dsh[i3 + 2] = dsh0[2]; }
    }
  }
  if (STAGED) {
    __syncthreads();
    rows_store<STAGED ? R : 12, ACC, GM_THREADS>(dL_dsh + (size_t)block_start * R, s_tile, s_vis, rows, 3 * nco);
  }
}

template <int NV, int MT>
static cudaError_t launch_nv_mt(cudaStream_t s, int P, int D, const float* means3D, const float* shs,
                                const float* scales, const float* rotations, float scale_modifier,
                                const ViewGrad* views, float* dL_dopacity, float* dL_dmean3D, float* dL_dsh,
                                float* dL_dscale, float* dL_drot, float* grad_norm_accum, int32_t* visible_count,
                                int32_t* max_radii, bool acc) {
  MultiArgs<NV> args;
  for (int v = 0; v < NV; v++) args.v[v] = views[v];
  const size_t smem = (MT == 4 || MT == 16) ? (size_t)GM_THREADS * row_stride(3 * MT) * sizeof(float) : 0;
  auto kern = acc ? geom_backward_multi_kernel<NV, MT, true> : geom_backward_multi_kernel<NV, MT, false>;
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
  }
  kern<<<cdiv(P, GM_THREADS), GM_THREADS, smem, s>>>(P, D, means3D, shs, scales, rotations, scale_modifier, args, dL_dopacity,
                                       dL_dmean3D, dL_dsh, dL_dscale, dL_drot, grad_norm_accum, visible_count,
                                       max_radii);
  count_launch();
  return cudaGetLastError();
}

template <int NV>
static cudaError_t launch_nv(cudaStream_t s, int P, int D, int M, const float* means3D, const float* shs,
                             const float* scales, const float* rotations, float scale_modifier,
                             const ViewGrad* views, float* dL_dopacity, float* dL_dmean3D, float* dL_dsh,
                             float* dL_dscale, float* dL_drot, float* grad_norm_accum, int32_t* visible_count,
                             int32_t* max_radii, bool acc) {
#define GSR_GM_CALL(MT)                                                                                             \
  return launch_nv_mt<NV, MT>(s, P, D, means3D, shs, scales, rotations, scale_modifier, views, dL_dopacity,        \
                              dL_dmean3D, dL_dsh, dL_dscale, dL_drot, grad_norm_accum, visible_count, max_radii, acc)
  if (M == 16) { GSR_GM_CALL(16); }
  if (M == 4) { GSR_GM_CALL(4); }
  if (M == 1) { GSR_GM_CALL(1); }
#undef GSR_GM_CALL
  return cudaErrorInvalidValue;
}

bool geom_backward_multi_supported(int M) { return M == 1 || M == 4 || M == 16; }

// views: host array of n_views descriptors.  Processed in batches of up to GM_MAX_VIEWS; the first
// batch assigns (unless `accumulate`), the following ones add.
cudaError_t launch_geom_backward_multi(cudaStream_t s, int P, int D, int M, const float* means3D, const float* shs,
                                       const float* scales, const float* rotations, float scale_modifier,
                                       const ViewGrad* views, int n_views, float* dL_dopacity,
                                       float* dL_dmean3D, float* dL_dsh, float* dL_dscale, float* dL_drot,
                                       float* grad_norm_accum, int32_t* visible_count, int32_t* max_radii,
                                       bool accumulate) {
  if (P == 0 || n_views == 0) return cudaSuccess;
  const ViewGrad* vg = views;
  bool acc = accumulate;
  for (int first = 0; first < n_views; first += GM_MAX_VIEWS) {
    const int nv = n_views - first < GM_MAX_VIEWS ? n_views - first : GM_MAX_VIEWS;
    cudaError_t e;
#define GSR_GM_NV(NV)                                                                                               \
  e = launch_nv<NV>(s, P, D, M, means3D, shs, scales, rotations, scale_modifier, vg + first, dL_dopacity, dL_dmean3D, \
                    dL_dsh, dL_dscale, dL_drot, grad_norm_accum, visible_count, max_radii, acc)
    switch (nv) {
      case 1: GSR_GM_NV(1); break;
      case 2: GSR_GM_NV(2); break;
      case 3: GSR_GM_NV(3); break;
      default: GSR_GM_NV(4); break;
    }
#undef GSR_GM_NV
    if (e != cudaSuccess) return e;
    acc = true;
  }
  return cudaSuccess;
}

}  // namespace gsr
