/*
 * gsplat_oracle.c -- see gsplat_oracle.h.  TEST INFRASTRUCTURE ONLY; "parity unpinned" for the
 * whole pipeline (third-party rasterizer absent from /root/reference), sub-steps pinned by
 * tests/golden/ against the reference's in-tree Python fragments.
 *
 * Build:  make -C oracle      (gcc -O2 -ffp-contract=off -mfma -fopenmp)
 *
 * Every function cites the SURVEY.md Appendix-A section it restates and, where one exists, the
 * reference file:line that corroborates it.
 */
#include "gsplat_oracle.h"

#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ---- explicit IEEE binary32 ops (no contraction: -ffp-contract=off) ------------------------ */
static inline float MUL(float a, float b) { return a * b; }
static inline float ADD(float a, float b) { return a + b; }
static inline float SUB(float a, float b) { return a - b; }
static inline float DIV(float a, float b) { return a / b; }
static inline float FMA(float a, float b, float c) { return fmaf(a, b, c); }
static inline float SQRT(float a) { return sqrtf(a); }
/* fminf/fmaxf return the non-NaN operand, like CUDA's min/max float overloads */
static inline float MINF(float a, float b) { return fminf(a, b); }
static inline float MAXF(float a, float b) { return fmaxf(a, b); }

/* float -> int32 with the semantics of PTX cvt.rzi.s32.f32 (truncate, saturate, NaN -> 0) */
static inline int32_t f2i_rz_sat(float f) {
  if (f != f) return 0;
  if (f >= 2147483648.0f) return INT32_MAX;
  if (f <= -2147483648.0f) return INT32_MIN;
  return (int32_t)f;
}
static inline int imin(int a, int b) { return a < b ? a : b; }
static inline int imax(int a, int b) { return a > b ? a : b; }

/* Appendix A.1; identical to gs-simp/utils/sh_utils.py:26-43 */
static const float SH_C0 = 0.28209479177387814f;
static const float SH_C1 = 0.4886025119029199f;
static const float SH_C2[5] = {1.0925484305920792f, -1.0925484305920792f, 0.31539156525252005f,
                               -1.0925484305920792f, 0.5462742152960396f};
static const float SH_C3[7] = {-0.5900435899266435f, 2.890611442640554f,  -0.4570457994644658f,
                               0.3731763325901154f,  -0.4570457994644658f, 1.445305721320277f,
                               -0.5900435899266435f};

int gso_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

/* bench.py --impl reference sets the thread count itself: torchrun exports OMP_NUM_THREADS=1 to its workers, which
 * made the reference arm of the multi-GPU runs 9x slower than its own single-process run (round-1 verdict). */
void gso_set_num_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}

/* A.4 getHigherMsb */
uint32_t gso_higher_msb(uint32_t n) {
  uint32_t msb = sizeof(n) * 4;
  uint32_t step = msb;
  while (step > 1) {
    step /= 2;
    if (n >> msb)
      msb += step;
    else
      msb -= step;
  }
  if (n >> msb) msb++;
  return msb;
}

/* A.2 step 2: column-major 4x4 applied to (p,1), three rows.  cameras.py:60 stores W2C transposed,
 * hence the m[0],m[4],m[8],m[12] indexing. */
static inline void xform4x3(const float* p, const float* m, float* o) {
  for (int r = 0; r < 3; r++)
    o[r] = ADD(FMA(m[8 + r], p[2], FMA(m[4 + r], p[1], MUL(m[r], p[0]))), m[12 + r]);
}
static inline void xform4x4(const float* p, const float* m, float* o) {
  for (int r = 0; r < 4; r++)
    o[r] = ADD(FMA(m[8 + r], p[2], FMA(m[4 + r], p[1], MUL(m[r], p[0]))), m[12 + r]);
}
static inline float dot3(const float* a, const float* b) {
  return FMA(a[2], b[2], FMA(a[1], b[1], MUL(a[0], b[0])));
}

/* A.2 step 4.  Rotation matrix == gs-simp/utils/general_utils.py:92-100 (no renormalisation here;
 * F.normalize runs upstream of the rasterizer, gaussian_model.py:41,100-101).
 * Sigma = (R S)(R S)^T == general_utils.py:103-112 + gaussian_model.py:27-31; packing order
 * [00,01,02,11,12,22] == strip_lowerdiag, general_utils.py:69-74. */
static void rotation_matrix(const float* q, float R[3][3]) {
  const float r = q[0], x = q[1], y = q[2], z = q[3];
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
static void compute_cov3D(const float* scale, float mod, const float* q, float* cov) {
  float R[3][3], Mx[3][3];
  rotation_matrix(q, R);
  const float s[3] = {MUL(mod, scale[0]), MUL(mod, scale[1]), MUL(mod, scale[2])};
  for (int a = 0; a < 3; a++)
    for (int k = 0; k < 3; k++) Mx[a][k] = MUL(R[a][k], s[k]);
  cov[0] = dot3(Mx[0], Mx[0]);
  cov[1] = dot3(Mx[0], Mx[1]);
  cov[2] = dot3(Mx[0], Mx[2]);
  cov[3] = dot3(Mx[1], Mx[1]);
  cov[4] = dot3(Mx[1], Mx[2]);
  cov[5] = dot3(Mx[2], Mx[2]);
}

/* A.2 step 5 helper shared by forward and K8: T0,T1 = rows of J * R_w2c, and Sigma*T0, Sigma*T1. */
typedef struct {
  float t[3];       /* clamped view-space point */
  float txtz, tytz; /* unclamped ratios (for the K8 clamp masks) */
  float T0[3], T1[3];
  float v0[3], v1[3]; /* Sigma T0, Sigma T1 */
  float a, b, c;      /* cov2D + 0.3 dilation */
} ewa_t;

static void ewa_project(const float* pview, float focal_x, float focal_y, float tan_fovx,
                        float tan_fovy, const float* cov3D, const float* V, ewa_t* e) {
  const float limx = MUL(1.3f, tan_fovx), limy = MUL(1.3f, tan_fovy);
  const float tz = pview[2];
  e->txtz = DIV(pview[0], tz);
  e->tytz = DIV(pview[1], tz);
  e->t[0] = MUL(MINF(limx, MAXF(-limx, e->txtz)), tz);
  e->t[1] = MUL(MINF(limy, MAXF(-limy, e->tytz)), tz);
  e->t[2] = tz;
  const float tz2 = MUL(tz, tz);
  const float J00 = DIV(focal_x, tz);
  const float J02 = DIV(-MUL(focal_x, e->t[0]), tz2);
  const float J11 = DIV(focal_y, tz);
  const float J12 = DIV(-MUL(focal_y, e->t[1]), tz2);
  /* R_w2c[r][c] = V[4c + r] */
  for (int c = 0; c < 3; c++) {
    e->T0[c] = FMA(J02, V[4 * c + 2], MUL(J00, V[4 * c + 0]));
    e->T1[c] = FMA(J12, V[4 * c + 2], MUL(J11, V[4 * c + 1]));
  }
  const float S[3][3] = {{cov3D[0], cov3D[1], cov3D[2]},
                         {cov3D[1], cov3D[3], cov3D[4]},
                         {cov3D[2], cov3D[4], cov3D[5]}};
  for (int a = 0; a < 3; a++) {
    e->v0[a] = dot3(S[a], e->T0);
    e->v1[a] = dot3(S[a], e->T1);
  }
  e->a = ADD(dot3(e->T0, e->v0), 0.3f);
  e->b = dot3(e->T0, e->v1);
  e->c = ADD(dot3(e->T1, e->v1), 0.3f);
}

/* A.2 step 10: SH basis weights for a unit direction; polynomial == sh_utils.py:74-100. */
static void sh_weights(int deg, float x, float y, float z, float* w) {
  w[0] = SH_C0;
  if (deg > 0) {
    w[1] = MUL(-SH_C1, y);
    w[2] = MUL(SH_C1, z);
    w[3] = MUL(-SH_C1, x);
    if (deg > 1) {
      const float xx = MUL(x, x), yy = MUL(y, y), zz = MUL(z, z);
      const float xy = MUL(x, y), yz = MUL(y, z), xz = MUL(x, z);
      w[4] = MUL(SH_C2[0], xy);
      w[5] = MUL(SH_C2[1], yz);
      w[6] = MUL(SH_C2[2], SUB(SUB(MUL(2.0f, zz), xx), yy));
      w[7] = MUL(SH_C2[3], xz);
      w[8] = MUL(SH_C2[4], SUB(xx, yy));
      if (deg > 2) {
        w[9] = MUL(MUL(SH_C3[0], y), FMA(3.0f, xx, -yy));
        w[10] = MUL(MUL(SH_C3[1], xy), z);
        w[11] = MUL(MUL(SH_C3[2], y), SUB(SUB(MUL(4.0f, zz), xx), yy));
        w[12] = MUL(MUL(SH_C3[3], z), FMA(-3.0f, yy, FMA(-3.0f, xx, MUL(2.0f, zz))));
        w[13] = MUL(MUL(SH_C3[4], x), SUB(SUB(MUL(4.0f, zz), xx), yy));
        w[14] = MUL(MUL(SH_C3[5], z), SUB(xx, yy));
        w[15] = MUL(MUL(SH_C3[6], x), FMA(-3.0f, yy, xx));
      }
    }
  }
}

static void unit_dir(const float* p, const float* campos, float* dir_orig, float* dir) {
  for (int k = 0; k < 3; k++) dir_orig[k] = SUB(p[k], campos[k]);
  const float len = SQRT(dot3(dir_orig, dir_orig));
  for (int k = 0; k < 3; k++) dir[k] = DIV(dir_orig[k], len);
}

/* A.2 step 8: ndc2Pix is evaluated in DOUBLE (double literals upstream) then narrowed. */
static inline float ndc2pix(float v, int S) {
  return (float)((((double)v + 1.0) * (double)S - 1.0) * 0.5);
}

/* A.2 step 9 */
static inline void get_rect(const float* p, int max_radius, int gx, int gy, int* rmin, int* rmax) {
  const float r = (float)max_radius;
  rmin[0] = imin(gx, imax(0, f2i_rz_sat(DIV(SUB(p[0], r), 16.0f))));
  rmin[1] = imin(gy, imax(0, f2i_rz_sat(DIV(SUB(p[1], r), 16.0f))));
  rmax[0] = imin(gx, imax(0, f2i_rz_sat(DIV(SUB(ADD(ADD(p[0], r), 16.0f), 1.0f), 16.0f))));
  rmax[1] = imin(gy, imax(0, f2i_rz_sat(DIV(SUB(ADD(ADD(p[1], r), 16.0f), 1.0f), 16.0f))));
}

/* ---- K1 ------------------------------------------------------------------------------------ */
int gso_preprocess(int P, int D, int M, const float* means3D, const float* scales,
                   float scale_modifier, const float* rotations, const float* opacities,
                   const float* shs, const float* cov3D_precomp, const float* colors_precomp,
                   const float* viewmatrix, const float* projmatrix, const float* campos, int W,
                   int H, float tan_fovx, float tan_fovy, int prefiltered, int32_t* radii,
                   float* means2D, float* depths, float* cov3D, float* rgb, float* conic_opacity,
                   uint32_t* tiles_touched, uint8_t* clamped) {
  const float focal_y = H / (2.0f * tan_fovy);
  const float focal_x = W / (2.0f * tan_fovx);
  const int gx = (W + GSO_BLOCK_X - 1) / GSO_BLOCK_X, gy = (H + GSO_BLOCK_Y - 1) / GSO_BLOCK_Y;
  int trapped = 0;
#pragma omp parallel for schedule(static) reduction(| : trapped)
  for (int i = 0; i < P; i++) {
    radii[i] = 0;
    tiles_touched[i] = 0;
    const float* p = means3D + 3 * (size_t)i;
    float pv[3];
    xform4x3(p, viewmatrix, pv);
    if (pv[2] <= 0.2f) { /* x/y frustum test is disabled upstream */
      if (prefiltered) trapped = 1;
      continue;
    }
    float ph[4];
    xform4x4(p, projmatrix, ph);
    const float pw = DIV(1.0f, ADD(ph[3], 0.0000001f));
    const float pproj[2] = {MUL(ph[0], pw), MUL(ph[1], pw)};

    const float* c3;
    if (cov3D_precomp) {
      c3 = cov3D_precomp + 6 * (size_t)i;
    } else {
      compute_cov3D(scales + 3 * (size_t)i, scale_modifier, rotations + 4 * (size_t)i,
                    cov3D + 6 * (size_t)i);
      c3 = cov3D + 6 * (size_t)i;
    }
    ewa_t e;
    ewa_project(pv, focal_x, focal_y, tan_fovx, tan_fovy, c3, viewmatrix, &e);
    const float det = FMA(e.a, e.c, -MUL(e.b, e.b));
    if (det == 0.0f) continue;
    const float det_inv = DIV(1.0f, det);
    const float conic[3] = {MUL(e.c, det_inv), MUL(-e.b, det_inv), MUL(e.a, det_inv)};
    const float mid = MUL(0.5f, ADD(e.a, e.c));
    const float sq = SQRT(MAXF(0.1f, FMA(mid, mid, -det)));
    const float lambda1 = ADD(mid, sq), lambda2 = SUB(mid, sq);
    const float my_radius = ceilf(MUL(3.0f, SQRT(MAXF(lambda1, lambda2))));
    const float pim[2] = {ndc2pix(pproj[0], W), ndc2pix(pproj[1], H)};
    int rmin[2], rmax[2];
    get_rect(pim, f2i_rz_sat(my_radius), gx, gy, rmin, rmax);
    if ((rmax[0] - rmin[0]) * (rmax[1] - rmin[1]) == 0) continue;

    if (!colors_precomp) {
      float dir_orig[3], dir[3], w[16];
      unit_dir(p, campos, dir_orig, dir);
      sh_weights(D, dir[0], dir[1], dir[2], w);
      const int nco = (D + 1) * (D + 1);
      const float* sh = shs + (size_t)i * M * 3;
      for (int c = 0; c < 3; c++) {
        float res = MUL(w[0], sh[c]);
        for (int k = 1; k < nco; k++) res = FMA(w[k], sh[3 * k + c], res);
        res = ADD(res, 0.5f);
        clamped[3 * (size_t)i + c] = (res < 0.0f);
        rgb[3 * (size_t)i + c] = MAXF(res, 0.0f); /* gaussian_renderer/__init__.py:78 */
      }
    } else {
      for (int c = 0; c < 3; c++) {
        rgb[3 * (size_t)i + c] = colors_precomp[3 * (size_t)i + c];
        clamped[3 * (size_t)i + c] = 0;
      }
    }
    depths[i] = pv[2];
    radii[i] = f2i_rz_sat(my_radius);
    means2D[2 * (size_t)i + 0] = pim[0];
    means2D[2 * (size_t)i + 1] = pim[1];
    conic_opacity[4 * (size_t)i + 0] = conic[0];
    conic_opacity[4 * (size_t)i + 1] = conic[1];
    conic_opacity[4 * (size_t)i + 2] = conic[2];
    conic_opacity[4 * (size_t)i + 3] = opacities[i];
    tiles_touched[i] = (uint32_t)((rmax[1] - rmin[1]) * (rmax[0] - rmin[0]));
  }
  return trapped ? -1 : 0;
}

/* ---- K2 ------------------------------------------------------------------------------------ */
int64_t gso_inclusive_scan(int P, const uint32_t* tiles_touched, uint32_t* point_offsets) {
  int64_t acc = 0;
  for (int i = 0; i < P; i++) {
    acc += tiles_touched[i];
    point_offsets[i] = (uint32_t)acc; /* upstream keeps uint32 offsets */
  }
  return acc;
}

/* ---- K3 ------------------------------------------------------------------------------------ */
void gso_duplicate_with_keys(int P, const float* means2D, const float* depths,
                             const uint32_t* point_offsets, const int32_t* radii, int W, int H,
                             uint64_t* keys_unsorted, uint32_t* values_unsorted) {
  const int gx = (W + GSO_BLOCK_X - 1) / GSO_BLOCK_X, gy = (H + GSO_BLOCK_Y - 1) / GSO_BLOCK_Y;
#pragma omp parallel for schedule(static)
  for (int i = 0; i < P; i++) {
    if (radii[i] <= 0) continue;
    uint64_t off = (i == 0) ? 0 : point_offsets[i - 1];
    int rmin[2], rmax[2];
    get_rect(means2D + 2 * (size_t)i, radii[i], gx, gy, rmin, rmax);
    uint32_t dbits;
    memcpy(&dbits, depths + i, 4);
    for (int y = rmin[1]; y < rmax[1]; y++)
      for (int x = rmin[0]; x < rmax[0]; x++) {
        uint64_t key = (uint64_t)(uint32_t)(y * gx + x);
        key = (key << 32) | dbits;
        keys_unsorted[off] = key;
        values_unsorted[off] = (uint32_t)i;
        off++;
      }
  }
}

/* ---- K4: stable LSD radix sort, 16-bit digits ---------------------------------------------- */
void gso_sort_pairs(int64_t N, const uint64_t* keys_in, const uint32_t* vals_in,
                    uint64_t* keys_out, uint32_t* vals_out, int end_bit) {
  if (N <= 0) return;
  const int passes = (end_bit + 15) / 16;
  uint64_t* kbuf[2] = {(uint64_t*)malloc((size_t)N * 8), (uint64_t*)malloc((size_t)N * 8)};
  uint32_t* vbuf[2] = {(uint32_t*)malloc((size_t)N * 4), (uint32_t*)malloc((size_t)N * 4)};
  memcpy(kbuf[0], keys_in, (size_t)N * 8);
  memcpy(vbuf[0], vals_in, (size_t)N * 4);
  size_t* hist = (size_t*)malloc(65536 * sizeof(size_t));
  int cur = 0;
  for (int p = 0; p < passes; p++) {
    const int shift = 16 * p;
    const int bits = (end_bit - shift) < 16 ? (end_bit - shift) : 16;
    const uint64_t mask = ((uint64_t)1 << bits) - 1;
    memset(hist, 0, 65536 * sizeof(size_t));
    const uint64_t* ki = kbuf[cur];
    const uint32_t* vi = vbuf[cur];
    uint64_t* ko = kbuf[cur ^ 1];
    uint32_t* vo = vbuf[cur ^ 1];
    for (int64_t i = 0; i < N; i++) hist[(ki[i] >> shift) & mask]++;
    size_t acc = 0;
    for (int b = 0; b < 65536; b++) {
      size_t c = hist[b];
      hist[b] = acc;
      acc += c;
    }
    for (int64_t i = 0; i < N; i++) {
      size_t d = hist[(ki[i] >> shift) & mask]++;
      ko[d] = ki[i];
      vo[d] = vi[i];
    }
    cur ^= 1;
  }
  memcpy(keys_out, kbuf[cur], (size_t)N * 8);
  memcpy(vals_out, vbuf[cur], (size_t)N * 4);
  free(hist);
  free(kbuf[0]);
  free(kbuf[1]);
  free(vbuf[0]);
  free(vbuf[1]);
}

/* ---- K5 ------------------------------------------------------------------------------------ */
void gso_identify_tile_ranges(int64_t N, const uint64_t* keys_sorted, int num_tiles,
                              uint32_t* ranges) {
  memset(ranges, 0, (size_t)num_tiles * 2 * sizeof(uint32_t));
  for (int64_t idx = 0; idx < N; idx++) {
    const uint32_t cur = (uint32_t)(keys_sorted[idx] >> 32);
    if (idx == 0) {
      ranges[2 * cur] = 0;
    } else {
      const uint32_t prev = (uint32_t)(keys_sorted[idx - 1] >> 32);
      if (cur != prev) {
        ranges[2 * prev + 1] = (uint32_t)idx;
        ranges[2 * cur] = (uint32_t)idx;
      }
    }
    if (idx == N - 1) ranges[2 * cur + 1] = (uint32_t)N;
  }
}

/* shared by K6/K7: power = -0.5(cx dx^2 + cz dy^2) - cy dx dy */
static inline float gauss_power(float cx, float cy, float cz, float dx, float dy) {
  const float q = FMA(MUL(cz, dy), dy, MUL(MUL(cx, dx), dx));
  return FMA(-0.5f, q, -MUL(MUL(cy, dx), dy));
}

/* ---- K6 ------------------------------------------------------------------------------------ */
void gso_blend_forward_tiles(int W, int H, const uint32_t* ranges, const uint32_t* point_list,
                             const float* means2D, const float* rgb, const float* depths,
                             const float* conic_opacity, const float* bg, float* out_color,
                             float* out_depth, float* final_T, uint32_t* n_contrib, int tile_start,
                             int tile_step) {
  const int gx = (W + GSO_BLOCK_X - 1) / GSO_BLOCK_X, gy = (H + GSO_BLOCK_Y - 1) / GSO_BLOCK_Y;
  const size_t HW = (size_t)H * W;
  const int n_sel = (gx * gy - tile_start + tile_step - 1) / tile_step;
#pragma omp parallel for schedule(dynamic, 4)
  for (int ts = 0; ts < n_sel; ts++) {
    const int tile = tile_start + ts * tile_step;
    const int bx = tile % gx, by = tile / gx;
    const uint32_t r0 = ranges[2 * tile], r1 = ranges[2 * tile + 1];
    for (int ly = 0; ly < GSO_BLOCK_Y; ly++)
      for (int lx = 0; lx < GSO_BLOCK_X; lx++) {
        const int px = bx * GSO_BLOCK_X + lx, py = by * GSO_BLOCK_Y + ly;
        if (px >= W || py >= H) continue;
        const float pxf = (float)px, pyf = (float)py; /* pixel centres at integer coords */
        float T = 1.0f, C[3] = {0.f, 0.f, 0.f}, Dm = 15.0f; /* sentinel: gen_seq.py:50 */
        uint32_t contributor = 0, last = 0;
        for (uint32_t k = r0; k < r1; k++) {
          contributor++;
          const uint32_t g = point_list[k];
          const float dx = SUB(means2D[2 * (size_t)g], pxf);
          const float dy = SUB(means2D[2 * (size_t)g + 1], pyf);
          const float* co = conic_opacity + 4 * (size_t)g;
          const float power = gauss_power(co[0], co[1], co[2], dx, dy);
          if (power > 0.0f) continue;
          const float alpha = MINF(0.99f, MUL(co[3], expf(power)));
          if (alpha < 1.0f / 255.0f) continue;
          const float test_T = MUL(T, SUB(1.0f, alpha));
          if (test_T < 0.0001f) break; /* done */
          const float wgt = MUL(alpha, T);
          for (int ch = 0; ch < 3; ch++) C[ch] = FMA(rgb[3 * (size_t)g + ch], wgt, C[ch]);
          if (T > 0.5f && test_T < 0.5f) Dm = depths[g]; /* median depth (w-depth fork) */
          T = test_T;
          last = contributor;
        }
        const size_t pix = (size_t)py * W + px;
        final_T[pix] = T;
        n_contrib[pix] = last;
        for (int ch = 0; ch < 3; ch++) out_color[ch * HW + pix] = FMA(T, bg[ch], C[ch]);
        out_depth[pix] = Dm;
      }
  }
}

void gso_blend_forward(int W, int H, const uint32_t* ranges, const uint32_t* point_list,
                       const float* means2D, const float* rgb, const float* depths,
                       const float* conic_opacity, const float* bg, float* out_color,
                       float* out_depth, float* final_T, uint32_t* n_contrib) {
  gso_blend_forward_tiles(W, H, ranges, point_list, means2D, rgb, depths, conic_opacity, bg, out_color,
                          out_depth, final_T, n_contrib, 0, 1);
}

/* ---- K7 ------------------------------------------------------------------------------------ */
void gso_blend_backward(int W, int H, const uint32_t* ranges, const uint32_t* point_list,
                        const float* means2D, const float* rgb, const float* conic_opacity,
                        const float* bg, const float* final_T, const uint32_t* n_contrib,
                        const float* dL_dpixels, float* dL_dmean2D, float* dL_dconic,
                        float* dL_dopacity, float* dL_dcolors) {
  gso_blend_backward_tiles(W, H, ranges, point_list, means2D, rgb, conic_opacity, bg, final_T, n_contrib,
                           dL_dpixels, dL_dmean2D, dL_dconic, dL_dopacity, dL_dcolors, 0, 1);
}

void gso_blend_backward_tiles(int W, int H, const uint32_t* ranges, const uint32_t* point_list,
                              const float* means2D, const float* rgb, const float* conic_opacity,
                              const float* bg, const float* final_T, const uint32_t* n_contrib,
                              const float* dL_dpixels, float* dL_dmean2D, float* dL_dconic,
                              float* dL_dopacity, float* dL_dcolors, int tile_start, int tile_step) {
  const int gx = (W + GSO_BLOCK_X - 1) / GSO_BLOCK_X, gy = (H + GSO_BLOCK_Y - 1) / GSO_BLOCK_Y;
  const size_t HW = (size_t)H * W;
  const int G = gx * gy;
  uint32_t N = 0;
  for (int t = 0; t < G; t++)
    if (ranges[2 * t + 1] > N) N = ranges[2 * t + 1];
  if (N == 0) return;
  /* per-(tile,Gaussian) partial sums in double -> deterministic, order-independent to ~1e-15.
   * 9 slots: color 0..2, mean2D 3..4, conic 5..7, opacity 8 */
  double* part = (double*)calloc((size_t)N * 9, sizeof(double));
  const float ddelx_dx = MUL(0.5f, (float)W), ddely_dy = MUL(0.5f, (float)H);
  const int n_sel = (G - tile_start + tile_step - 1) / tile_step;
#pragma omp parallel for schedule(dynamic, 4)
  for (int ts = 0; ts < n_sel; ts++) {
    const int tile = tile_start + ts * tile_step;
    const int bx = tile % gx, by = tile / gx;
    const uint32_t r0 = ranges[2 * tile], r1 = ranges[2 * tile + 1];
    if (r1 <= r0) continue;
    for (int ly = 0; ly < GSO_BLOCK_Y; ly++)
      for (int lx = 0; lx < GSO_BLOCK_X; lx++) {
        const int px = bx * GSO_BLOCK_X + lx, py = by * GSO_BLOCK_Y + ly;
        if (px >= W || py >= H) continue;
        const size_t pix = (size_t)py * W + px;
        const float pxf = (float)px, pyf = (float)py;
        const float T_final = final_T[pix];
        float T = T_final;
        const uint32_t last = n_contrib[pix];
        float accum_rec[3] = {0, 0, 0}, last_color[3] = {0, 0, 0}, last_alpha = 0.0f;
        const float dLp[3] = {dL_dpixels[pix], dL_dpixels[HW + pix], dL_dpixels[2 * HW + pix]};
        float bg_dot = 0.0f;
        for (int ch = 0; ch < 3; ch++) bg_dot = FMA(bg[ch], dLp[ch], bg_dot);
        /* contributor index of list slot k (0-based) is k - r0; processed iff < last */
        for (int64_t k = (int64_t)r0 + (int64_t)last - 1; k >= (int64_t)r0; k--) {
          const uint32_t g = point_list[k];
          const float dx = SUB(means2D[2 * (size_t)g], pxf);
          const float dy = SUB(means2D[2 * (size_t)g + 1], pyf);
          const float* co = conic_opacity + 4 * (size_t)g;
          const float power = gauss_power(co[0], co[1], co[2], dx, dy);
          if (power > 0.0f) continue;
          const float Gv = expf(power);
          const float alpha = MINF(0.99f, MUL(co[3], Gv));
          if (alpha < 1.0f / 255.0f) continue;
          T = DIV(T, SUB(1.0f, alpha));
          const float dch = MUL(alpha, T);
          double* ps = part + (size_t)k * 9;
          float dL_dalpha = 0.0f;
          for (int ch = 0; ch < 3; ch++) {
            const float c = rgb[3 * (size_t)g + ch];
            accum_rec[ch] = FMA(last_alpha, last_color[ch], MUL(SUB(1.0f, last_alpha), accum_rec[ch]));
            last_color[ch] = c;
            dL_dalpha = FMA(SUB(c, accum_rec[ch]), dLp[ch], dL_dalpha);
            ps[ch] += (double)MUL(dch, dLp[ch]);
          }
          dL_dalpha = MUL(dL_dalpha, T);
          last_alpha = alpha;
          dL_dalpha = FMA(DIV(-T_final, SUB(1.0f, alpha)), bg_dot, dL_dalpha);
          const float dL_dG = MUL(co[3], dL_dalpha);
          const float gdx = MUL(Gv, dx), gdy = MUL(Gv, dy);
          const float dG_ddelx = SUB(MUL(-gdx, co[0]), MUL(gdy, co[1]));
          const float dG_ddely = SUB(MUL(-gdy, co[2]), MUL(gdx, co[1]));
          ps[3] += (double)MUL(MUL(dL_dG, dG_ddelx), ddelx_dx);
          ps[4] += (double)MUL(MUL(dL_dG, dG_ddely), ddely_dy);
          ps[5] += (double)MUL(MUL(MUL(-0.5f, gdx), dx), dL_dG);
          ps[6] += (double)MUL(MUL(MUL(-0.5f, gdx), dy), dL_dG);
          ps[7] += (double)MUL(MUL(MUL(-0.5f, gdy), dy), dL_dG);
          ps[8] += (double)MUL(Gv, dL_dalpha);
        }
      }
  }
  /* fixed-order reduction per Gaussian: walk the selected tiles' instances in sorted order,
   * accumulate in double */
  int P = 0;
  for (int ts = 0; ts < n_sel; ts++) {
    const int tile = tile_start + ts * tile_step;
    for (uint32_t k = ranges[2 * tile]; k < ranges[2 * tile + 1]; k++)
      if ((int)point_list[k] + 1 > P) P = (int)point_list[k] + 1;
  }
  double* acc = (double*)calloc((size_t)(P > 0 ? P : 1) * 9, sizeof(double));
  for (int ts = 0; ts < n_sel; ts++) {
    const int tile = tile_start + ts * tile_step;
    for (uint32_t k = ranges[2 * tile]; k < ranges[2 * tile + 1]; k++) {
      double* a = acc + (size_t)point_list[k] * 9;
      const double* ps = part + (size_t)k * 9;
      for (int j = 0; j < 9; j++) a[j] += ps[j];
    }
  }
#pragma omp parallel for schedule(static)
  for (int g = 0; g < P; g++) {
    const double* a = acc + (size_t)g * 9;
    for (int ch = 0; ch < 3; ch++) dL_dcolors[3 * (size_t)g + ch] += (float)a[ch];
    dL_dmean2D[3 * (size_t)g + 0] += (float)a[3];
    dL_dmean2D[3 * (size_t)g + 1] += (float)a[4];
    dL_dconic[4 * (size_t)g + 0] += (float)a[5];
    dL_dconic[4 * (size_t)g + 1] += (float)a[6];
    dL_dconic[4 * (size_t)g + 3] += (float)a[7];
    dL_dopacity[g] += (float)a[8];
  }
  free(acc);
  free(part);
}

/* ---- K8 + K9 ------------------------------------------------------------------------------- */
void gso_preprocess_backward(int P, int D, int M, const float* means3D, const int32_t* radii,
                             const float* shs, const uint8_t* clamped, const float* scales,
                             const float* rotations, float scale_modifier, const float* cov3D,
                             const float* viewmatrix, const float* projmatrix, const float* campos,
                             int W, int H, float tan_fovx, float tan_fovy,
                             const float* dL_dmean2D, const float* dL_dconic,
                             const float* dL_dcolors, float* dL_dmeans3D, float* dL_dcov3D,
                             float* dL_dsh, float* dL_dscales, float* dL_drots) {
  const float focal_y = H / (2.0f * tan_fovy);
  const float focal_x = W / (2.0f * tan_fovx);
  const float* V = viewmatrix;
  const float* PM = projmatrix;
#pragma omp parallel for schedule(static)
  for (int i = 0; i < P; i++) {
    if (!(radii[i] > 0)) continue;
    const float* m = means3D + 3 * (size_t)i;
    const float* c3 = cov3D + 6 * (size_t)i;
    float dmean[3];

    /* --- K8 computeCov2D backward (A.7) --- */
    {
      float pv[3];
      xform4x3(m, V, pv);
      ewa_t e;
      ewa_project(pv, focal_x, focal_y, tan_fovx, tan_fovy, c3, V, &e);
      const float limx = 1.3f * tan_fovx, limy = 1.3f * tan_fovy;
      const float x_grad_mul = (e.txtz < -limx || e.txtz > limx) ? 0.0f : 1.0f;
      const float y_grad_mul = (e.tytz < -limy || e.tytz > limy) ? 0.0f : 1.0f;
      const float a = e.a, b = e.b, c = e.c;
      const float gx_ = dL_dconic[4 * (size_t)i + 0], gy_ = dL_dconic[4 * (size_t)i + 1],
                  gz_ = dL_dconic[4 * (size_t)i + 3];
      const float denom = a * c - b * b;
      const float denom2inv = 1.0f / ((denom * denom) + 0.0000001f);
      float dL_da = 0, dL_db = 0, dL_dc = 0;
      float* dcov = dL_dcov3D + 6 * (size_t)i;
      if (denom2inv != 0.0f) {
        dL_da = denom2inv * (-c * c * gx_ + 2 * b * c * gy_ + (denom - a * c) * gz_);
        dL_dc = denom2inv * (-a * a * gz_ + 2 * a * b * gy_ + (denom - a * c) * gx_);
        dL_db = denom2inv * 2 * (b * c * gx_ - (denom + 2 * b * b) * gy_ + a * b * gz_);
        const float* T0 = e.T0;
        const float* T1 = e.T1;
        dcov[0] = T0[0] * T0[0] * dL_da + T0[0] * T1[0] * dL_db + T1[0] * T1[0] * dL_dc;
        dcov[3] = T0[1] * T0[1] * dL_da + T0[1] * T1[1] * dL_db + T1[1] * T1[1] * dL_dc;
        dcov[5] = T0[2] * T0[2] * dL_da + T0[2] * T1[2] * dL_db + T1[2] * T1[2] * dL_dc;
        dcov[1] = 2 * T0[0] * T0[1] * dL_da + (T0[0] * T1[1] + T0[1] * T1[0]) * dL_db +
                  2 * T1[0] * T1[1] * dL_dc;
        dcov[2] = 2 * T0[0] * T0[2] * dL_da + (T0[0] * T1[2] + T0[2] * T1[0]) * dL_db +
                  2 * T1[0] * T1[2] * dL_dc;
        dcov[4] = 2 * T0[2] * T0[1] * dL_da + (T0[1] * T1[2] + T0[2] * T1[1]) * dL_db +
                  2 * T1[1] * T1[2] * dL_dc;
      } else {
        for (int k = 0; k < 6; k++) dcov[k] = 0;
      }
      /* dL/dT rows: 2 (Sigma T0) dL_da + (Sigma T1) dL_db ; 2 (Sigma T1) dL_dc + (Sigma T0) dL_db */
      float dT0[3], dT1[3];
      for (int k = 0; k < 3; k++) {
        dT0[k] = 2 * e.v0[k] * dL_da + e.v1[k] * dL_db;
        dT1[k] = 2 * e.v1[k] * dL_dc + e.v0[k] * dL_db;
      }
      /* R_w2c[r][k] = V[4k + r] */
      float dJ00 = 0, dJ02 = 0, dJ11 = 0, dJ12 = 0;
      for (int k = 0; k < 3; k++) {
        dJ00 += V[4 * k + 0] * dT0[k];
        dJ02 += V[4 * k + 2] * dT0[k];
        dJ11 += V[4 * k + 1] * dT1[k];
        dJ12 += V[4 * k + 2] * dT1[k];
      }
      const float tz = 1.0f / e.t[2], tz2 = tz * tz, tz3 = tz2 * tz;
      const float dtx = x_grad_mul * -focal_x * tz2 * dJ02;
      const float dty = y_grad_mul * -focal_y * tz2 * dJ12;
      const float dtz = -focal_x * tz2 * dJ00 - focal_y * tz2 * dJ11 +
                        (2 * focal_x * e.t[0]) * tz3 * dJ02 + (2 * focal_y * e.t[1]) * tz3 * dJ12;
      /* R_w2c^T applied (transformVec4x3Transpose) */
      dmean[0] = V[0] * dtx + V[1] * dty + V[2] * dtz;
      dmean[1] = V[4] * dtx + V[5] * dty + V[6] * dtz;
      dmean[2] = V[8] * dtx + V[9] * dty + V[10] * dtz;
    }

    /* --- K9: perspective projection of the mean --- */
    {
      float mh[4];
      xform4x4(m, PM, mh);
      const float m_w = 1.0f / (mh[3] + 0.0000001f);
      const float mul1 = (PM[0] * m[0] + PM[4] * m[1] + PM[8] * m[2] + PM[12]) * m_w * m_w;
      const float mul2 = (PM[1] * m[0] + PM[5] * m[1] + PM[9] * m[2] + PM[13]) * m_w * m_w;
      const float g2x = dL_dmean2D[3 * (size_t)i + 0], g2y = dL_dmean2D[3 * (size_t)i + 1];
      for (int k = 0; k < 3; k++)
        dmean[k] += (PM[4 * k + 0] * m_w - PM[4 * k + 3] * mul1) * g2x +
                    (PM[4 * k + 1] * m_w - PM[4 * k + 3] * mul2) * g2y;
    }

    /* --- K9: SH backward --- */
    if (shs) {
      float dir_orig[3], dir[3];
      unit_dir(m, campos, dir_orig, dir);
      const float x = dir[0], y = dir[1], z = dir[2];
      const float* sh = shs + (size_t)i * M * 3;
      float* dsh = dL_dsh + (size_t)i * M * 3;
      float dRGB[3];
      for (int c = 0; c < 3; c++)
        dRGB[c] = clamped[3 * (size_t)i + c] ? 0.0f : dL_dcolors[3 * (size_t)i + c];
      float w[16];
      sh_weights(D, x, y, z, w);
      const int nco = (D + 1) * (D + 1);
      for (int k = 0; k < nco; k++)
        for (int c = 0; c < 3; c++) dsh[3 * k + c] = w[k] * dRGB[c];
      /* d(weights)/d(dir) */
      float dwx[16] = {0}, dwy[16] = {0}, dwz[16] = {0};
      if (D > 0) {
        dwy[1] = -SH_C1;
        dwz[2] = SH_C1;
        dwx[3] = -SH_C1;
        if (D > 1) {
          const float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
          dwx[4] = SH_C2[0] * y;  dwy[4] = SH_C2[0] * x;
          dwy[5] = SH_C2[1] * z;  dwz[5] = SH_C2[1] * y;
          dwx[6] = SH_C2[2] * -2 * x; dwy[6] = SH_C2[2] * -2 * y; dwz[6] = SH_C2[2] * 4 * z;
          dwx[7] = SH_C2[3] * z;  dwz[7] = SH_C2[3] * x;
          dwx[8] = SH_C2[4] * 2 * x; dwy[8] = SH_C2[4] * -2 * y;
          if (D > 2) {
            dwx[9] = SH_C3[0] * 6 * xy;            dwy[9] = SH_C3[0] * 3 * (xx - yy);
            dwx[10] = SH_C3[1] * yz;               dwy[10] = SH_C3[1] * xz;  dwz[10] = SH_C3[1] * xy;
            dwx[11] = SH_C3[2] * -2 * xy;          dwy[11] = SH_C3[2] * (4 * zz - xx - 3 * yy);
            dwz[11] = SH_C3[2] * 8 * yz;
            dwx[12] = SH_C3[3] * -6 * xz;          dwy[12] = SH_C3[3] * -6 * yz;
            dwz[12] = SH_C3[3] * 3 * (2 * zz - xx - yy);
            dwx[13] = SH_C3[4] * (4 * zz - 3 * xx - yy); dwy[13] = SH_C3[4] * -2 * xy;
            dwz[13] = SH_C3[4] * 8 * xz;
            dwx[14] = SH_C3[5] * 2 * xz;           dwy[14] = SH_C3[5] * -2 * yz;
            dwz[14] = SH_C3[5] * (xx - yy);
            dwx[15] = SH_C3[6] * 3 * (xx - yy);    dwy[15] = SH_C3[6] * -6 * xy;
          }
        }
      }
      float ddir[3] = {0, 0, 0};
      for (int k = 1; k < nco; k++) {
        float s = 0;
        for (int c = 0; c < 3; c++) s += sh[3 * k + c] * dRGB[c];
        ddir[0] += dwx[k] * s;
        ddir[1] += dwy[k] * s;
        ddir[2] += dwz[k] * s;
      }
      /* through normalize: (I |v|^2 - v v^T) / |v|^3 */
      const float* v = dir_orig;
      const float sum2 = v[0] * v[0] + v[1] * v[1] + v[2] * v[2];
      const float invsum32 = 1.0f / sqrtf(sum2 * sum2 * sum2);
      dmean[0] += ((sum2 - v[0] * v[0]) * ddir[0] - v[1] * v[0] * ddir[1] - v[2] * v[0] * ddir[2]) * invsum32;
      dmean[1] += (-v[0] * v[1] * ddir[0] + (sum2 - v[1] * v[1]) * ddir[1] - v[2] * v[1] * ddir[2]) * invsum32;
      dmean[2] += (-v[0] * v[2] * ddir[0] - v[1] * v[2] * ddir[1] + (sum2 - v[2] * v[2]) * ddir[2]) * invsum32;
    }
    for (int k = 0; k < 3; k++) dL_dmeans3D[3 * (size_t)i + k] = dmean[k];

    /* --- K9: cov3D backward (only when scales/rotations were given) --- */
    if (scales) {
      const float* q = rotations + 4 * (size_t)i;
      float R[3][3];
      rotation_matrix(q, R);
      const float s[3] = {scale_modifier * scales[3 * (size_t)i + 0],
                          scale_modifier * scales[3 * (size_t)i + 1],
                          scale_modifier * scales[3 * (size_t)i + 2]};
      const float* dc = dL_dcov3D + 6 * (size_t)i;
      const float Gs[3][3] = {{dc[0], 0.5f * dc[1], 0.5f * dc[2]},
                              {0.5f * dc[1], dc[3], 0.5f * dc[4]},
                              {0.5f * dc[2], 0.5f * dc[4], dc[5]}};
      float Mx[3][3], dM[3][3];
      for (int a = 0; a < 3; a++)
        for (int k = 0; k < 3; k++) Mx[a][k] = R[a][k] * s[k];
      for (int a = 0; a < 3; a++)
        for (int k = 0; k < 3; k++)
          dM[a][k] = 2.0f * (Gs[a][0] * Mx[0][k] + Gs[a][1] * Mx[1][k] + Gs[a][2] * Mx[2][k]);
      /* NOTE: like the public rasterizer this is d/d(mod*scale); scale_modifier is not applied */
      for (int k = 0; k < 3; k++)
        dL_dscales[3 * (size_t)i + k] = R[0][k] * dM[0][k] + R[1][k] * dM[1][k] + R[2][k] * dM[2][k];
      float g[3][3];
      for (int a = 0; a < 3; a++)
        for (int k = 0; k < 3; k++) g[a][k] = dM[a][k] * s[k];
      const float r = q[0], x = q[1], y = q[2], z = q[3];
      float* dq = dL_drots + 4 * (size_t)i;
      dq[0] = 2 * z * (g[1][0] - g[0][1]) + 2 * y * (g[0][2] - g[2][0]) + 2 * x * (g[2][1] - g[1][2]);
      dq[1] = 2 * y * (g[0][1] + g[1][0]) + 2 * z * (g[0][2] + g[2][0]) + 2 * r * (g[2][1] - g[1][2]) -
              4 * x * (g[1][1] + g[2][2]);
      dq[2] = 2 * x * (g[0][1] + g[1][0]) + 2 * r * (g[0][2] - g[2][0]) + 2 * z * (g[1][2] + g[2][1]) -
              4 * y * (g[0][0] + g[2][2]);
      dq[3] = 2 * r * (g[1][0] - g[0][1]) + 2 * x * (g[0][2] + g[2][0]) + 2 * y * (g[1][2] + g[2][1]) -
              4 * z * (g[0][0] + g[1][1]);
    }
  }
}

/* ---- K10 ----------------------------------------------------------------------------------- */
void gso_mark_visible(int P, const float* means3D, const float* viewmatrix, uint8_t* present) {
#pragma omp parallel for schedule(static)
  for (int i = 0; i < P; i++) {
    float pv[3];
    xform4x3(means3D + 3 * (size_t)i, viewmatrix, pv);
    present[i] = pv[2] > 0.2f;
  }
}

/* ---- simple_knn.distCUDA2 -------------------------------------------------
 * Restates what the reference gets from `simple_knn._C.distCUDA2` (import gs-simp/scene/gaussian_model.py:20,
 * call sites :134, :546, :623).  The extension's source is not in the reference tree (third-party
 * simple-knn, unpinned, environment.yml:17): PARITY UNPINNED.  Published algorithm: for every point keep
 * the three smallest squared distances to the OTHER points (strict '<' insertion into a sorted triple that
 * starts at FLT_MAX) and return their mean; upstream only prunes the scan with conservative box tests, so
 * the value equals this brute-force scan.  Distance op order = what nvcc's default contraction makes of
 * d.x*d.x + d.y*d.y + d.z*d.z: fma(dz,dz, fma(dy,dy, dx*dx)).
 * queries == NULL: all P points (out has P entries); otherwise out[k] is for point queries[k]. */
void gso_knn3_mean_dist2(int P, const float* pts, const int32_t* queries, int nq, float* out) {
  const int n = queries ? nq : P;
#pragma omp parallel for schedule(dynamic, 16)
  for (int k = 0; k < n; k++) {
    const int i = queries ? queries[k] : k;
    const float qx = pts[3 * (size_t)i], qy = pts[3 * (size_t)i + 1], qz = pts[3 * (size_t)i + 2];
    float best[3] = {FLT_MAX, FLT_MAX, FLT_MAX};
    for (int j = 0; j < P; j++) {
      if (j == i) continue;
      const float dx = qx - pts[3 * (size_t)j], dy = qy - pts[3 * (size_t)j + 1], dz = qz - pts[3 * (size_t)j + 2];
      float d = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
      for (int t = 0; t < 3; t++)
        if (best[t] > d) { const float tmp = best[t]; best[t] = d; d = tmp; }
    }
    out[k] = ((best[0] + best[1]) + best[2]) / 3.0f;
  }
}
