// preprocess.cu -- K1 (per-Gaussian EWA projection, SH->RGB, radii, tile rect) and K10 (markVisible).
//
// Restates SURVEY.md Appendix A.2; in-tree corroboration in the reference:
//   SH polynomial          gs-simp/utils/sh_utils.py:57-112, clamp gaussian_renderer/__init__.py:78
//   R(q), Sigma = L L^T    gs-simp/utils/general_utils.py:80-112, scene/gaussian_model.py:27-31
//   matrix conventions     gs-simp/scene/cameras.py:60-63, utils/graphics_utils.py:51-70
//
// One thread per Gaussian.  Inputs arrive in the reference's AoS layouts ((P,3), (P,4), (P,M,3));
// a warp's loads of consecutive Gaussians cover one contiguous span, so every fetched sector is
// fully used.  Outputs are one 48-byte blend record per Gaussian (three float4, gathered by the
// blend kernels as 2 sectors) plus SoA side arrays.
#include "common.cuh"

namespace gsr {

__device__ __forceinline__ void xform4x3(const float p[3], const float* m, float o[3]) {
#pragma unroll
  for (int r = 0; r < 3; r++)
    o[r] = ADD(FMA(m[8 + r], p[2], FMA(m[4 + r], p[1], MUL(m[r], p[0]))), m[12 + r]);
}
__device__ __forceinline__ float dot3(const float a[3], const float b[3]) {
  return FMA(a[2], b[2], FMA(a[1], b[1], MUL(a[0], b[0])));
}

__device__ __forceinline__ void rotation_matrix(float r, float x, float y, float z, float R[3][3]) {
  R[0][0] = FMA(-2.0f, FMA(y, y, MUL(z, z)), 1.0f);
  R[0][1] = MUL(2.0f, FMA(x, y, -MUL(r, z)));
  R[0][2] = MUL(2.0f, FMA(x, z, MUL(r, y)));
  R[1][0] = MUL(2.0f, FMA(x, y, MUL(r, z)));
  R[1][1] = FMA(-2.0f, FMA(x, x, MUL(z, z)), 1.0f);
  R[1][2] = MUL(2.0f, FMA(y, z, -MUL(r, x)));
  R[2][0] = MUL(2.0f, FMA(x, z, -MUL(r, y)));
  R[2][1] = MUL(2.0f, FMA(y, z, MUL(r, x)));
  R[2][2] = FMA(-2.0f, FMA(x, x, MUL(y, y)), 1.0f);
}

// SH basis weights for unit direction (x,y,z); identical op order to the oracle's sh_weights().
template <int MAXC>
__device__ __forceinline__ void sh_weights(int deg, float x, float y, float z, float* w) {
  w[0] = GSR_SH_C0;
  if (deg > 0) {
    w[1] = MUL(-GSR_SH_C1, y);
    w[2] = MUL(GSR_SH_C1, z);
    w[3] = MUL(-GSR_SH_C1, x);
    if (deg > 1) {
      const float xx = MUL(x, x), yy = MUL(y, y), zz = MUL(z, z);
      const float xy = MUL(x, y), yz = MUL(y, z), xz = MUL(x, z);
      w[4] = MUL(GSR_SH_C2_0, xy);
      w[5] = MUL(GSR_SH_C2_1, yz);
      w[6] = MUL(GSR_SH_C2_2, SUB(SUB(MUL(2.0f, zz), xx), yy));
      w[7] = MUL(GSR_SH_C2_3, xz);
      w[8] = MUL(GSR_SH_C2_4, SUB(xx, yy));
      if (deg > 2) {
        w[9] = MUL(MUL(GSR_SH_C3_0, y), FMA(3.0f, xx, -yy));
        w[10] = MUL(MUL(GSR_SH_C3_1, xy), z);
        w[11] = MUL(MUL(GSR_SH_C3_2, y), SUB(SUB(MUL(4.0f, zz), xx), yy));
        w[12] = MUL(MUL(GSR_SH_C3_3, z), FMA(-3.0f, yy, FMA(-3.0f, xx, MUL(2.0f, zz))));
        w[13] = MUL(MUL(GSR_SH_C3_4, x), SUB(SUB(MUL(4.0f, zz), xx), yy));
        w[14] = MUL(MUL(GSR_SH_C3_5, z), SUB(xx, yy));
        w[15] = MUL(MUL(GSR_SH_C3_6, x), FMA(-3.0f, yy, xx));
      }
    }
  }
}

__device__ __forceinline__ float ndc2pix(float v, int S) {
  // evaluated in double, as the reference does (double literals), without contraction
  const double t = __dmul_rn(__dadd_rn((double)v, 1.0), (double)S);
  return (float)__dmul_rn(__dadd_rn(t, -1.0), 0.5);
}

template <int DEG>
__device__ __forceinline__ void eval_sh(const float* __restrict__ sh, const float* w,
                                        float& r, float& g, float& b) {
  constexpr int NCO = (DEG + 1) * (DEG + 1);
  // 3*NCO contiguous floats; read as scalars through the read-only path (12-byte rows are not
  // 16-byte aligned for NCO == 1, and L1 serves the neighbouring lanes' bytes of each sector)
  r = MUL(w[0], __ldg(sh + 0));
  g = MUL(w[0], __ldg(sh + 1));
  b = MUL(w[0], __ldg(sh + 2));
#pragma unroll
  for (int k = 1; k < NCO; k++) {
    r = FMA(w[k], __ldg(sh + 3 * k + 0), r);
    g = FMA(w[k], __ldg(sh + 3 * k + 1), g);
    b = FMA(w[k], __ldg(sh + 3 * k + 2), b);
  }
}

// Same sum, same order, but the row is fetched with 16-byte loads (rows of 3M floats are 16-byte
// aligned when M % 4 == 0): a per-thread row walk touches 32 different 128-byte lines per warp
// instruction, so the L1 wavefront count -- the limiter of this kernel at SH degree 3, not DRAM --
// drops 4x against scalar loads.
template <int DEG>
__device__ __forceinline__ void eval_sh_vec(const float4* __restrict__ row4, const float* w,
                                            float& r, float& g, float& b) {
  constexpr int NCO = (DEG + 1) * (DEG + 1);
  constexpr int NV = (3 * NCO + 3) / 4;  // <= 3M/4: the over-read (if any) stays inside the row
  float v[4 * NV];
#pragma unroll
  for (int q = 0; q < NV; q++) {
    const float4 t = __ldg(row4 + q);
    v[4 * q] = t.x; v[4 * q + 1] = t.y; v[4 * q + 2] = t.z; v[4 * q + 3] = t.w;
  }
  r = MUL(w[0], v[0]);
  g = MUL(w[0], v[1]);
  b = MUL(w[0], v[2]);
#pragma unroll
  for (int k = 1; k < NCO; k++) {
    r = FMA(w[k], v[3 * k + 0], r);
    g = FMA(w[k], v[3 * k + 1], g);
    b = FMA(w[k], v[3 * k + 2], b);
  }
}

struct PreArgs {
  PreView v[GSR_MAX_BATCH];
  int nv;
  int tight;   // GSR_FLAG_TIGHT_BINNING: shrink the stored tile rect to the {alpha >= 1/255} bounding box
};

// Tile-index range [t0, t1) of the tiles (16 pixel centres 16 t .. 16 t + 15 each) that the interval [lo, hi] reaches:
// exactly the tiles for which blend_common.cuh::stage_entry's tests `hi >= tile0 && lo <= tile0 + 15` hold (the union of
// its sub-tile tests).  lo / 16 and hi / 16 are exact in fp32 (power-of-two scaling), floorf and the fraction are exact.
__device__ __forceinline__ void tile_span(float lo, float hi, int& t0, int& t1) {
  const float u = lo * 0.0625f, fu = floorf(u);
  t0 = __float2int_rz(fu) + ((u - fu) > 0.9375f ? 1 : 0);   // smallest t with 16 t + 15 >= lo
  t1 = min(__float2int_rz(floorf(hi * 0.0625f)), 1 << 24) + 1;   // one past the largest t with 16 t <= hi
}

int g_pre_min_blocks = 6;   // experiment switch (gsr_debug_set knob 2): resident CTAs per SM the kernel is compiled for

template <int MINB>
__global__ void __launch_bounds__(256, MINB)
preprocess_kernel(int P, int D, int M, const float* __restrict__ means3D,
                  const float* __restrict__ scales, const float* __restrict__ rotations,
                  const float* __restrict__ opacities, const float* __restrict__ shs,
                  const float* __restrict__ cov3D_precomp, const float* __restrict__ colors_precomp,
                  int prefiltered, const __grid_constant__ PreArgs args) {
  __shared__ float s_cam[35];
  __shared__ unsigned long long s_sum_tiles;
  __shared__ uint32_t s_sum_vis;
  if (threadIdx.x == 0) { s_sum_tiles = 0; s_sum_vis = 0; }   // published by the barrier inside load_camera
  // view-interleaved CTAs: the nv CTAs of one 256-Gaussian chunk are neighbours in launch order
  const int vw = (int)(blockIdx.x % (unsigned)args.nv);
  const PreView& pv_ = args.v[vw];
  const Camera& cam = pv_.cam;
  int32_t* __restrict__ radii = pv_.radii;
  float4* __restrict__ rec = pv_.rec;
  float* __restrict__ depths = pv_.depths;
  uint8_t* __restrict__ clamped = pv_.clamped;
  uint32_t* __restrict__ tiles_touched = pv_.tiles_touched;
  uint32_t* __restrict__ depth_keys = pv_.depth_keys;
  ushort4* __restrict__ rects = pv_.rects;
  int32_t* __restrict__ status = pv_.status;
  if (pv_.gacc != nullptr) {
    // the backward's accumulator rows of this CTA's Gaussians (48 bytes each, contiguous): three coalesced 16-byte stores
    // per thread, issued before anything else so that they drain under the loads below
    const size_t first = (size_t)(blockIdx.x / (unsigned)args.nv) * 256;
    const int n4 = 3 * (int)min((size_t)256, (size_t)P - first);
    float4* g4 = pv_.gacc + 3 * first;
    for (int t = threadIdx.x; t < n4; t += 256) g4[t] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
  }
  load_camera(cam, s_cam);
  const float* view = s_cam;
  const float* proj = s_cam + 16;
  const float* campos = s_cam + 32;
  const int i_raw = (int)(blockIdx.x / (unsigned)args.nv) * blockDim.x + threadIdx.x;
  const bool in_range = i_raw < P;   // threads past the end run on the last Gaussian and write nothing (no early
  const int i = in_range ? i_raw : P - 1;  // return: the whole CTA takes part in the reduction at the end)

  // defaults for culled Gaussians
  int32_t out_radius = 0;
  uint32_t out_tiles = 0;
  uint32_t out_key = 0xFFFFFFFFu;
  ushort4 out_rect = make_ushort4(0, 0, 0, 0);

  const float p[3] = {__ldg(means3D + 3 * (size_t)i), __ldg(means3D + 3 * (size_t)i + 1),
                      __ldg(means3D + 3 * (size_t)i + 2)};
  float pv[3];
  xform4x3(p, view, pv);
  bool alive = in_range && pv[2] > 0.2f;  // A.2 step 2 (the x/y frustum test is disabled upstream)
  if (in_range && !alive && prefiltered) atomicExch(status, 1);

  if (alive) {
    float ph[4];
#pragma unroll
    for (int r = 0; r < 4; r++)
      ph[r] = ADD(FMA(proj[8 + r], p[2], FMA(proj[4 + r], p[1], MUL(proj[r], p[0]))),
                  proj[12 + r]);
    const float pw = DIV(1.0f, ADD(ph[3], 0.0000001f));
    const float pprojx = MUL(ph[0], pw), pprojy = MUL(ph[1], pw);

    // ---- 3D covariance (A.2 step 4) ----
    float c3[6];
    if (cov3D_precomp != nullptr) {
#pragma unroll
      for (int k = 0; k < 6; k++) c3[k] = __ldg(cov3D_precomp + 6 * (size_t)i + k);
    } else {
      const float4 q = __ldg(reinterpret_cast<const float4*>(rotations) + i);
      float R[3][3], Mx[3][3];
      rotation_matrix(q.x, q.y, q.z, q.w, R);
      const float s[3] = {MUL(cam.scale_modifier, __ldg(scales + 3 * (size_t)i)),
                          MUL(cam.scale_modifier, __ldg(scales + 3 * (size_t)i + 1)),
                          MUL(cam.scale_modifier, __ldg(scales + 3 * (size_t)i + 2))};
#pragma unroll
      for (int a = 0; a < 3; a++)
#pragma unroll
        for (int k = 0; k < 3; k++) Mx[a][k] = MUL(R[a][k], s[k]);
      c3[0] = dot3(Mx[0], Mx[0]);
      c3[1] = dot3(Mx[0], Mx[1]);
      c3[2] = dot3(Mx[0], Mx[2]);
      c3[3] = dot3(Mx[1], Mx[1]);
      c3[4] = dot3(Mx[1], Mx[2]);
      c3[5] = dot3(Mx[2], Mx[2]);
    }

    // ---- EWA projection (A.2 step 5) ----
    const float limx = MUL(1.3f, cam.tan_fovx), limy = MUL(1.3f, cam.tan_fovy);
    const float tz = pv[2];
    const float tx = MUL(fminf(limx, fmaxf(-limx, DIV(pv[0], tz))), tz);
    const float ty = MUL(fminf(limy, fmaxf(-limy, DIV(pv[1], tz))), tz);
    const float tz2 = MUL(tz, tz);
    const float J00 = DIV(cam.focal_x, tz);
    const float J02 = DIV(-MUL(cam.focal_x, tx), tz2);
    const float J11 = DIV(cam.focal_y, tz);
    const float J12 = DIV(-MUL(cam.focal_y, ty), tz2);
    float T0[3], T1[3], v0[3], v1[3];
#pragma unroll
    for (int c = 0; c < 3; c++) {
      T0[c] = FMA(J02, view[4 * c + 2], MUL(J00, view[4 * c + 0]));
      T1[c] = FMA(J12, view[4 * c + 2], MUL(J11, view[4 * c + 1]));
    }
    const float S[3][3] = {{c3[0], c3[1], c3[2]}, {c3[1], c3[3], c3[4]}, {c3[2], c3[4], c3[5]}};
#pragma unroll
    for (int a = 0; a < 3; a++) {
      v0[a] = dot3(S[a], T0);
      v1[a] = dot3(S[a], T1);
    }
    const float ca = ADD(dot3(T0, v0), 0.3f);
    const float cb = dot3(T0, v1);
    const float cc = ADD(dot3(T1, v1), 0.3f);

    const float det = FMA(ca, cc, -MUL(cb, cb));
    if (det != 0.0f) {
      const float det_inv = DIV(1.0f, det);
      const float conx = MUL(cc, det_inv), cony = MUL(-cb, det_inv), conz = MUL(ca, det_inv);
      const float mid = MUL(0.5f, ADD(ca, cc));
      const float sq = SQRT(fmaxf(0.1f, FMA(mid, mid, -det)));
      const float lambda1 = ADD(mid, sq), lambda2 = SUB(mid, sq);
      const int my_radius = __float2int_rz(ceilf(MUL(3.0f, SQRT(fmaxf(lambda1, lambda2)))));
      const float pix_x = ndc2pix(pprojx, cam.W), pix_y = ndc2pix(pprojy, cam.H);
      int x0, y0, x1, y1;
      get_rect(pix_x, pix_y, my_radius, cam.grid_x, cam.grid_y, x0, y0, x1, y1);
      const uint32_t tiles = (uint32_t)((x1 - x0) * (y1 - y0));
      if (tiles != 0) {
        // ---- colour (A.2 step 10) ----
        float cr, cg, cbl;
        uint8_t clamp_bits = 0;
        if (colors_precomp == nullptr) {
          float dir[3] = {SUB(p[0], campos[0]), SUB(p[1], campos[1]), SUB(p[2], campos[2])};
          const float len = SQRT(dot3(dir, dir));
          const float dx = DIV(dir[0], len), dy = DIV(dir[1], len), dz = DIV(dir[2], len);
          float w[16];
          sh_weights<16>(D, dx, dy, dz, w);
          const float* sh = shs + (size_t)i * M * 3;
          if ((M & 3) == 0 && (reinterpret_cast<uintptr_t>(shs) & 15) == 0) {
            const float4* row4 = reinterpret_cast<const float4*>(sh);
            switch (D) {
              case 0: eval_sh_vec<0>(row4, w, cr, cg, cbl); break;
              case 1: eval_sh_vec<1>(row4, w, cr, cg, cbl); break;
              case 2: eval_sh_vec<2>(row4, w, cr, cg, cbl); break;
              default: eval_sh_vec<3>(row4, w, cr, cg, cbl); break;
            }
          } else {
            switch (D) {
              case 0: eval_sh<0>(sh, w, cr, cg, cbl); break;
              case 1: eval_sh<1>(sh, w, cr, cg, cbl); break;
              case 2: eval_sh<2>(sh, w, cr, cg, cbl); break;
              default: eval_sh<3>(sh, w, cr, cg, cbl); break;
            }
          }
          cr = ADD(cr, 0.5f);
          cg = ADD(cg, 0.5f);
          cbl = ADD(cbl, 0.5f);
          clamp_bits = (cr < 0.0f ? 1 : 0) | (cg < 0.0f ? 2 : 0) | (cbl < 0.0f ? 4 : 0);
          cr = fmaxf(cr, 0.0f);
          cg = fmaxf(cg, 0.0f);
          cbl = fmaxf(cbl, 0.0f);
        } else {
          cr = __ldg(colors_precomp + 3 * (size_t)i);
          cg = __ldg(colors_precomp + 3 * (size_t)i + 1);
          cbl = __ldg(colors_precomp + 3 * (size_t)i + 2);
        }
        const float opacity = __ldg(opacities + i);

        // ---- culling aids for the blend kernels (not part of the parity state) ----
        // alpha >= 1/255  <=>  power >= -ln(255*opacity) =: -tau.  Anything with power below
        // -(tau + slack) is skipped before the exp; the same ellipse bounds the pixels a warp's
        // 8x4 sub-tile can possibly receive: |dx| <= sqrt(2 tau' cov_xx), |dy| <= sqrt(2 tau' cov_yy).
        float hx = -1.0f, hy = -1.0f, power_cut = 1.0f;  // power_cut > 0  => always skipped
        const float tau = __logf(255.0f * opacity);
        if (tau > 0.0f && opacity > 0.0f) {
          const float taus = tau + GSR_POWER_SLACK;
          power_cut = -taus;
          hx = sqrtf(2.0f * taus * ca) * 1.001f + 0.01f;
          hy = sqrtf(2.0f * taus * cc) * 1.001f + 0.01f;
        }
        rec[3 * (size_t)i + 0] = make_float4(pix_x, pix_y, conx, cony);
        rec[3 * (size_t)i + 1] = make_float4(conz, opacity, hx, hy);
        rec[3 * (size_t)i + 2] = make_float4(cr, cg, cbl, power_cut);
        depths[i] = pv[2];
        clamped[i] = clamp_bits;
        out_radius = my_radius;
        uint32_t tiles_binned = tiles;
        if (args.tight) {
          // Only the tiles the {alpha >= 1/255} bounding box reaches (hx < 0: nothing can contribute).  The Gaussian stays
          // visible (radii, record, depth: the per-Gaussian backward and the statistics see what they saw before); with an
          // empty rect it takes no part in the depth sort either.
          tiles_binned = 0;
          if (hx >= 0.0f) {
            int bx0, bx1, by0, by1;
            tile_span(SUB(pix_x, hx), ADD(pix_x, hx), bx0, bx1);
            tile_span(SUB(pix_y, hy), ADD(pix_y, hy), by0, by1);
            x0 = max(x0, bx0); x1 = min(x1, bx1);
            y0 = max(y0, by0); y1 = min(y1, by1);
            if (x1 > x0 && y1 > y0) tiles_binned = (uint32_t)((x1 - x0) * (y1 - y0));
          }
        }
        if (tiles_binned != 0) {
          out_tiles = tiles_binned;
          out_key = __float_as_uint(pv[2]);
          out_rect = make_ushort4((unsigned short)x0, (unsigned short)y0, (unsigned short)x1, (unsigned short)y1);
        }
      }
    }
  }
  if (in_range) {
    radii[i] = out_radius;
    tiles_touched[i] = out_tiles;
    if (depth_keys != nullptr) depth_keys[i] = out_key;
    if (rects != nullptr) rects[i] = out_rect;  // two-level binning: the fused scan + expansion reads this
  }
  // N = sum of tiles_touched (status[2..3] as one uint64), known as soon as K1 ends: the host sizes the binning buffer
  // from it without waiting for the scan.  V = number of visible Gaussians (status[5]): the depth sort drops the culled
  // ones in its first pass and every later stage walks V entries.  Reduced per warp, then per CTA in shared memory:
  // two global atomics per CTA (one per warp on one address serialised in the L2: 94 k same-address atomics per view
  // cost K1 a quarter of its time once the second counter was added).
  __syncwarp();
  const uint32_t part = __reduce_add_sync(0xffffffffu, out_tiles);
  const uint32_t nvis = __popc(__ballot_sync(0xffffffffu, out_tiles != 0));
  if ((threadIdx.x & 31) == 0 && part != 0) {
    atomicAdd(&s_sum_tiles, (unsigned long long)part);
    atomicAdd(&s_sum_vis, nvis);
  }
  __syncthreads();
  if (threadIdx.x == 0 && s_sum_tiles != 0) {
    atomicAdd(reinterpret_cast<unsigned long long*>(status + 2), s_sum_tiles);
    atomicAdd(reinterpret_cast<uint32_t*>(status + 5), s_sum_vis);
  }
}

__global__ void __launch_bounds__(256)
mark_visible_kernel(int P, const float* __restrict__ means3D, const float* __restrict__ view,
                    uint8_t* __restrict__ present) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P) return;
  const float p[3] = {__ldg(means3D + 3 * (size_t)i), __ldg(means3D + 3 * (size_t)i + 1),
                      __ldg(means3D + 3 * (size_t)i + 2)};
  float m[16];
#pragma unroll
  for (int k = 0; k < 16; k++) m[k] = __ldg(view + k);
  float pv[3];
  xform4x3(p, m, pv);
  present[i] = pv[2] > 0.2f ? 1 : 0;
}

cudaError_t launch_preprocess(cudaStream_t s, int P, int D, int M, const float* means3D,
                              const float* scales, const float* rotations, const float* opacities,
                              const float* shs, const float* cov3D_precomp,
                              const float* colors_precomp, float scale_modifier, int prefiltered,
                              const PreView* views, int nv, bool tight) {
  if (P == 0 || nv <= 0) return cudaSuccess;
  if (nv > GSR_MAX_BATCH) return cudaErrorInvalidValue;
  PreArgs args{};
  args.nv = nv;
  args.tight = tight ? 1 : 0;
  for (int k = 0; k < nv; k++) {
    args.v[k] = views[k];
    args.v[k].cam.scale_modifier = scale_modifier;
  }
  const unsigned grid = (unsigned)cdiv(P, 256) * (unsigned)nv;
  if (g_pre_min_blocks >= 6)
    preprocess_kernel<6><<<grid, 256, 0, s>>>(P, D, M, means3D, scales, rotations, opacities, shs, cov3D_precomp, colors_precomp, prefiltered, args);
  else if (g_pre_min_blocks == 5)
    preprocess_kernel<5><<<grid, 256, 0, s>>>(P, D, M, means3D, scales, rotations, opacities, shs, cov3D_precomp, colors_precomp, prefiltered, args);
  else
    preprocess_kernel<4><<<grid, 256, 0, s>>>(P, D, M, means3D, scales, rotations, opacities, shs, cov3D_precomp, colors_precomp, prefiltered, args);
  count_launch();
  return cudaGetLastError();
}

cudaError_t launch_mark_visible(cudaStream_t s, int P, const float* means3D, const float* view,
                                uint8_t* present) {
  if (P == 0) return cudaSuccess;
  mark_visible_kernel<<<cdiv(P, 256), 256, 0, s>>>(P, means3D, view, present);
  count_launch();
  return cudaGetLastError();
}

}  // namespace gsr
