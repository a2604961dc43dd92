// loss.cu -- fused photometric loss of the training step, forward and backward.
//
// Replaces, for the caller of the rasterizer, the torch expression every training script of the reference
// evaluates right after render():
//     Ll1  = l1_loss(image, gt)                      gs-simp/utils/loss_utils.py:17-18   abs(x - y).mean()
//     loss = (1 - lambda) * Ll1 + lambda * (1 - ssim(image, gt))
//                                                    gs-simp/train.py:91-92, sds_train.py:117-118, inpaint_rec.py:118-123
//     ssim: 11x11 Gaussian window, sigma 1.5, zero padding 5, per channel (depthwise conv2d, groups = C),
//           C1 = 0.01^2, C2 = 0.03^2, mean over every element     loss_utils.py:23-62
// which in torch is five depthwise 11x11 convolutions forward and their five transposes backward plus ~25
// elementwise kernels over (C,H,W) maps.  Here: two tiled kernels.
//
//   forward : a CTA stages a (32+10) x (32+10) window of x and y in shared memory (zeros outside the image = the
//             reference's zero padding), runs the separable filter (11 horizontal taps into shared memory, 11
//             vertical taps in registers; both passes register-blocked, 4 outputs per thread from 14 inputs)
//             for the five moments mu_x, mu_y, E[x^2], E[y^2], E[xy] at once,
//             evaluates the SSIM map and the three partial derivatives of it with respect to the FILTERED
//             moments that depend on x  (d/dmu_x, d/dE[x^2], d/dE[xy]), stores those three maps, and reduces
//             sum|x-y| and sum(ssim) per CTA (fixed order: deterministic).  A one-block kernel sums the CTA
//             partials in double and writes {Ll1, ssim, loss}.
//   backward: the adjoint of a symmetric zero-padded filter is the same filter, so
//             dL/dx = g * [ (1-lambda)/n * sign(x-y) - lambda/n * ( F[dmu] + 2x F[dxx] + y F[dxy] ) ]
//             with F = the same separable 11x11 filter applied to the three stored maps.
//
// Algorithmic HBM bytes per (pixel, channel): forward 8 read + 12 written, backward 20 read + 4 written = 44 B
// (132 B per RGB pixel); both kernels are HBM-streaming with an on-chip 11x11 stencil.
#include <math.h>

#include "common.cuh"

namespace gsr {

namespace {

constexpr int LT_X = 32;            // output tile
constexpr int LT_Y = 32;
constexpr int LHALO = 5;            // window_size // 2, loss_utils.py:46
constexpr int LWIN = 11;
constexpr int LS_Y = LT_Y + 2 * LHALO;   // 42 staged rows
constexpr int LS_W = 44;                 // staged row: 42 columns used, padded to a multiple of 4 floats (LDS.128)
constexpr int LH_W = LT_X;               // horizontally filtered row: 32 columns
constexpr int LTHREADS = 256;
constexpr int LROWS_PER_THREAD = LT_Y / (LTHREADS / LT_X);   // 4 vertically adjacent outputs per thread
constexpr float SSIM_C1 = 0.01f * 0.01f;  // loss_utils.py:57-58 (python doubles, applied to float tensors)
constexpr float SSIM_C2 = 0.03f * 0.03f;

__constant__ float c_win[LWIN];

// gaussian(11, 1.5), loss_utils.py:23-25: python-double exp -> float32 tensor -> divided by its float32 sum
void make_window(float* w) {
  float sum = 0.0f;
  for (int i = 0; i < LWIN; i++) {
    const double x = (double)(i - LWIN / 2);
    w[i] = (float)exp(-(x * x) / (2.0 * 1.5 * 1.5));
    sum += w[i];
  }
  for (int i = 0; i < LWIN; i++) w[i] = w[i] / sum;
}

cudaError_t upload_window() {
  // per device: __constant__ memory is per context
  static bool done[64] = {};
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  if (dev >= 0 && dev < 64 && done[dev]) return cudaSuccess;
  float w[LWIN];
  make_window(w);
  e = cudaMemcpyToSymbol(c_win, w, sizeof(w));
  if (e == cudaSuccess && dev >= 0 && dev < 64) done[dev] = true;
  return e;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// 16 consecutive floats of a staged row, 16-byte aligned: four LDS.128 (a quarter warp covers all 32 banks once)
__device__ __forceinline__ void load16(const float* row, float* v) {
#pragma unroll
  for (int q = 0; q < 4; q++) {
    const float4 t = *reinterpret_cast<const float4*>(row + 4 * q);
    v[4 * q] = t.x; v[4 * q + 1] = t.y; v[4 * q + 2] = t.z; v[4 * q + 3] = t.w;
  }
}

// Register blocking is what this kernel is about: the filter is smem-bandwidth bound when every output reads its
// 11 taps itself (99 LDS per pixel in the first version, profiles/r01_v5), so the horizontal pass computes 4
// adjacent outputs from 14 staged values and the vertical pass 4 adjacent outputs from 14 filtered rows.
// grid (ceil(W/32), ceil(H/32), C), 256 threads
__global__ void __launch_bounds__(LTHREADS)
ssim_l1_forward_kernel(int H, int W, const float* __restrict__ img, const float* __restrict__ gt,
                       float* __restrict__ d_mu, float* __restrict__ d_xx, float* __restrict__ d_xy,
                       double2* __restrict__ partials) {
  __shared__ __align__(16) float sx[LS_Y][LS_W];
  __shared__ __align__(16) float sy[LS_Y][LS_W];
  __shared__ __align__(16) float sh[5][LS_Y][LH_W];   // horizontally filtered moments
  __shared__ float s_red[2][LTHREADS / 32];

  const int t = threadIdx.x;
  const int x0 = blockIdx.x * LT_X, y0 = blockIdx.y * LT_Y;
  const size_t plane = (size_t)blockIdx.z * (size_t)H * (size_t)W;
  const float* ip = img + plane;
  const float* gp = gt + plane;

  for (int i = t; i < LS_Y * LS_W; i += LTHREADS) {
    const int r = i / LS_W, c = i - r * LS_W;
    const int gy = y0 + r - LHALO, gx = x0 + c - LHALO;
    float a = 0.0f, b = 0.0f;
    if (gy >= 0 && gy < H && gx >= 0 && gx < W) {     // zeros outside the image = the reference's zero padding
      a = __ldg(ip + (size_t)gy * W + gx);
      b = __ldg(gp + (size_t)gy * W + gx);
    }
    sx[r][c] = a;
    sy[r][c] = b;
  }
  __syncthreads();

  // horizontal pass: item = (row r, group of 4 output columns)
  for (int i = t; i < LS_Y * (LT_X / 4); i += LTHREADS) {
    const int r = i >> 3, c0 = (i & 7) * 4;
    float a[16], b[16];
    load16(&sx[r][c0], a);
    load16(&sy[r][c0], b);
    float m1[4] = {0.f, 0.f, 0.f, 0.f}, m2[4] = {0.f, 0.f, 0.f, 0.f}, m11[4] = {0.f, 0.f, 0.f, 0.f},
          m22[4] = {0.f, 0.f, 0.f, 0.f}, m12[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int k = 0; k < LWIN; k++) {
      const float w = c_win[k];
#pragma unroll
      for (int j = 0; j < 4; j++) {
        const float wa = w * a[j + k], wb = w * b[j + k];
        m1[j] += wa;
        m2[j] += wb;
        m11[j] = fmaf(wa, a[j + k], m11[j]);
        m22[j] = fmaf(wb, b[j + k], m22[j]);
        m12[j] = fmaf(wa, b[j + k], m12[j]);
      }
    }
    *reinterpret_cast<float4*>(&sh[0][r][c0]) = make_float4(m1[0], m1[1], m1[2], m1[3]);
    *reinterpret_cast<float4*>(&sh[1][r][c0]) = make_float4(m2[0], m2[1], m2[2], m2[3]);
    *reinterpret_cast<float4*>(&sh[2][r][c0]) = make_float4(m11[0], m11[1], m11[2], m11[3]);
    *reinterpret_cast<float4*>(&sh[3][r][c0]) = make_float4(m22[0], m22[1], m22[2], m22[3]);
    *reinterpret_cast<float4*>(&sh[4][r][c0]) = make_float4(m12[0], m12[1], m12[2], m12[3]);
  }
  __syncthreads();

  // vertical pass: thread = (column c, 4 adjacent output rows r0..r0+3): 14 filtered rows feed 4 outputs
  float acc_l1 = 0.0f, acc_ss = 0.0f;
  {
    const int c = t & 31, r0 = (t >> 5) * LROWS_PER_THREAD;
    float o[5][LROWS_PER_THREAD];
#pragma unroll
    for (int m = 0; m < 5; m++)
#pragma unroll
      for (int j = 0; j < LROWS_PER_THREAD; j++) o[m][j] = 0.f;
#pragma unroll
    for (int q = 0; q < LROWS_PER_THREAD + LWIN - 1; q++) {
      float v[5];
#pragma unroll
      for (int m = 0; m < 5; m++) v[m] = sh[m][r0 + q][c];
#pragma unroll
      for (int j = 0; j < LROWS_PER_THREAD; j++) {
        const int k = q - j;               // tap index of staged row q for output j
        if (k >= 0 && k < LWIN) {
          const float w = c_win[k];
#pragma unroll
          for (int m = 0; m < 5; m++) o[m][j] = fmaf(w, v[m], o[m][j]);
        }
      }
    }
    const int gx = x0 + c;
#pragma unroll
    for (int j = 0; j < LROWS_PER_THREAD; j++) {
      const int gy = y0 + r0 + j;
      if (gy < H && gx < W) {
        const float mu1 = o[0][j], mu2 = o[1][j], e11 = o[2][j], e22 = o[3][j], e12 = o[4][j];
        // loss_utils.py:49-60
        const float mu1_sq = mu1 * mu1, mu2_sq = mu2 * mu2, mu1_mu2 = mu1 * mu2;
        const float sigma1_sq = e11 - mu1_sq, sigma2_sq = e22 - mu2_sq, sigma12 = e12 - mu1_mu2;
        const float A = 2.0f * mu1_mu2 + SSIM_C1;
        const float B = 2.0f * sigma12 + SSIM_C2;
        const float Cc = mu1_sq + mu2_sq + SSIM_C1;
        const float D = sigma1_sq + sigma2_sq + SSIM_C2;
        const float inv_cd = 1.0f / (Cc * D);
        const float s = A * B * inv_cd;
        // derivatives of s with respect to the filtered moments (mu1, E[xx], E[xy])
        const float ds_dmu1 = 2.0f * mu2 * (B - A) * inv_cd + 2.0f * mu1 * s * (1.0f / D - 1.0f / Cc);
        const float ds_dxx = -s / D;
        const float ds_dxy = 2.0f * A * inv_cd;
        const size_t off = plane + (size_t)gy * W + gx;
        d_mu[off] = ds_dmu1;
        d_xx[off] = ds_dxx;
        d_xy[off] = ds_dxy;
        acc_ss += s;
        acc_l1 += fabsf(sx[r0 + j + LHALO][c + LHALO] - sy[r0 + j + LHALO][c + LHALO]);
      }
    }
  }
  acc_l1 = warp_sum(acc_l1);
  acc_ss = warp_sum(acc_ss);
  if ((t & 31) == 0) {
    s_red[0][t >> 5] = acc_l1;
    s_red[1][t >> 5] = acc_ss;
  }
  __syncthreads();
  if (t == 0) {
    double a = 0.0, b = 0.0;
#pragma unroll
    for (int w = 0; w < LTHREADS / 32; w++) {
      a += (double)s_red[0][w];
      b += (double)s_red[1][w];
    }
    const size_t bid = ((size_t)blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
    partials[bid] = make_double2(a, b);
  }
}

// one block: out = {Ll1, ssim, (1-lambda) Ll1 + lambda (1 - ssim)}
__global__ void __launch_bounds__(1024)
loss_finalize_kernel(const double2* __restrict__ partials, int n_partials, double inv_n, float lambda,
                     float* __restrict__ out) {
  __shared__ double s_a[32], s_b[32];
  double a = 0.0, b = 0.0;
  for (int i = threadIdx.x; i < n_partials; i += blockDim.x) {
    const double2 p = partials[i];
    a += p.x;
    b += p.y;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    a += __shfl_xor_sync(0xffffffffu, a, o);
    b += __shfl_xor_sync(0xffffffffu, b, o);
  }
  if ((threadIdx.x & 31) == 0) {
    s_a[threadIdx.x >> 5] = a;
    s_b[threadIdx.x >> 5] = b;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double ta = 0.0, tb = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); w++) {
      ta += s_a[w];
      tb += s_b[w];
    }
    const float l1 = (float)(ta * inv_n), ss = (float)(tb * inv_n);
    out[0] = l1;
    out[1] = ss;
    out[2] = (1.0f - lambda) * l1 + lambda * (1.0f - ss);   // train.py:92, in fp32 like the torch scalars
  }
}

__global__ void __launch_bounds__(LTHREADS)
ssim_l1_backward_kernel(int H, int W, const float* __restrict__ img, const float* __restrict__ gt,
                        const float* __restrict__ d_mu, const float* __restrict__ d_xx,
                        const float* __restrict__ d_xy, const float* __restrict__ dL_dloss, float l1_scale,
                        float ssim_scale, float* __restrict__ dL_dimg) {
  __shared__ __align__(16) float sm[3][LS_Y][LS_W];
  __shared__ __align__(16) float sh[3][LS_Y][LH_W];

  const int t = threadIdx.x;
  const int x0 = blockIdx.x * LT_X, y0 = blockIdx.y * LT_Y;
  const size_t plane = (size_t)blockIdx.z * (size_t)H * (size_t)W;

  for (int i = t; i < LS_Y * LS_W; i += LTHREADS) {
    const int r = i / LS_W, c = i - r * LS_W;
    const int gy = y0 + r - LHALO, gx = x0 + c - LHALO;
    float a = 0.0f, b = 0.0f, d = 0.0f;
    if (gy >= 0 && gy < H && gx >= 0 && gx < W) {
      const size_t o = plane + (size_t)gy * W + gx;
      a = __ldg(d_mu + o);
      b = __ldg(d_xx + o);
      d = __ldg(d_xy + o);
    }
    sm[0][r][c] = a;
    sm[1][r][c] = b;
    sm[2][r][c] = d;
  }
  __syncthreads();

  for (int i = t; i < LS_Y * (LT_X / 4); i += LTHREADS) {
    const int r = i >> 3, c0 = (i & 7) * 4;
#pragma unroll
    for (int m = 0; m < 3; m++) {
      float a[16];
      load16(&sm[m][r][c0], a);
      float o[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int k = 0; k < LWIN; k++) {
        const float w = c_win[k];
#pragma unroll
        for (int j = 0; j < 4; j++) o[j] = fmaf(w, a[j + k], o[j]);
      }
      *reinterpret_cast<float4*>(&sh[m][r][c0]) = make_float4(o[0], o[1], o[2], o[3]);
    }
  }
  __syncthreads();

  const float g = dL_dloss ? __ldg(dL_dloss) : 1.0f;
  {
    const int c = t & 31, r0 = (t >> 5) * LROWS_PER_THREAD;
    float o[3][LROWS_PER_THREAD];
#pragma unroll
    for (int m = 0; m < 3; m++)
#pragma unroll
      for (int j = 0; j < LROWS_PER_THREAD; j++) o[m][j] = 0.f;
#pragma unroll
    for (int q = 0; q < LROWS_PER_THREAD + LWIN - 1; q++) {
      float v[3];
#pragma unroll
      for (int m = 0; m < 3; m++) v[m] = sh[m][r0 + q][c];
#pragma unroll
      for (int j = 0; j < LROWS_PER_THREAD; j++) {
        const int k = q - j;
        if (k >= 0 && k < LWIN) {
          const float w = c_win[k];
#pragma unroll
          for (int m = 0; m < 3; m++) o[m][j] = fmaf(w, v[m], o[m][j]);
        }
      }
    }
    const int gx = x0 + c;
#pragma unroll
    for (int j = 0; j < LROWS_PER_THREAD; j++) {
      const int gy = y0 + r0 + j;
      if (gy < H && gx < W) {
        const size_t off = plane + (size_t)gy * W + gx;
        const float x = __ldg(img + off), y = __ldg(gt + off);
        const float diff = x - y;
        const float sgn = diff > 0.0f ? 1.0f : (diff < 0.0f ? -1.0f : 0.0f);   // torch.abs backward: sign, 0 at 0
        const float dssim = o[0][j] + 2.0f * x * o[1][j] + y * o[2][j];
        dL_dimg[off] = g * (l1_scale * sgn - ssim_scale * dssim);
      }
    }
  }
}

}  // namespace

size_t loss_temp_bytes(int C, int H, int W) {
  const size_t n = (size_t)C * (size_t)H * (size_t)W;
  const size_t blocks = (size_t)cdiv(W, LT_X) * (size_t)cdiv(H, LT_Y) * (size_t)C;
  return align_up(3 * n * sizeof(float)) + align_up(blocks * sizeof(double2));
}

cudaError_t launch_loss_forward(cudaStream_t s, int C, int H, int W, const float* img, const float* gt, float lambda,
                                float* out3, char* temp) {
  cudaError_t e = upload_window();
  if (e != cudaSuccess) return e;
  const size_t n = (size_t)C * (size_t)H * (size_t)W;
  float* maps = reinterpret_cast<float*>(temp);
  double2* partials = reinterpret_cast<double2*>(temp + align_up(3 * n * sizeof(float)));
  const dim3 grid(cdiv(W, LT_X), cdiv(H, LT_Y), C);
  ssim_l1_forward_kernel<<<grid, LTHREADS, 0, s>>>(H, W, img, gt, maps, maps + n, maps + 2 * n, partials);
  loss_finalize_kernel<<<1, 1024, 0, s>>>(partials, (int)(grid.x * grid.y * grid.z), 1.0 / (double)n, lambda, out3);
  count_launch(2);
  return cudaGetLastError();
}

cudaError_t launch_loss_backward(cudaStream_t s, int C, int H, int W, const float* img, const float* gt, float lambda,
                                 const float* dL_dloss, const char* temp, float* dL_dimg) {
  cudaError_t e = upload_window();
  if (e != cudaSuccess) return e;
  const size_t n = (size_t)C * (size_t)H * (size_t)W;
  const float* maps = reinterpret_cast<const float*>(temp);
  const dim3 grid(cdiv(W, LT_X), cdiv(H, LT_Y), C);
  const float inv_n = (float)(1.0 / (double)n);
  ssim_l1_backward_kernel<<<grid, LTHREADS, 0, s>>>(H, W, img, gt, maps, maps + n, maps + 2 * n, dL_dloss,
                                                    (1.0f - lambda) * inv_n, lambda * inv_n, dL_dimg);
  count_launch(1);
  return cudaGetLastError();
}

}  // namespace gsr
