// refstruct.cu -- reference-STRUCTURE stand-in for K4, K6 and K7 (GSR_FLAG_REFERENCE).
//
// Not the product path: an ablation baseline.  The reference's own CUDA rasterizer is a third-party
// extension whose source is not in /root/reference and cannot be built offline,
// so "the reference on the same B200" cannot be timed.  SURVEY.md section 8(d) asks for the next best
// thing: kernels with the STRUCTURE the public algorithm is known to have,
// written here from that description, running on the same device and data:
//   K4  the 64-bit (tile | depth) key sort done by the library call the reference makes
//       (cub::DeviceRadixSort::SortPairs over bits [0, 32 + bit)), host-known N;
//   K6  one thread per pixel, 16x16 block, the tile list staged 256 entries at a time, every pixel
//       evaluates every entry (no sub-tile culling, no power cut-off), expf, colours and depth fetched
//       from global memory for contributing pairs, __syncthreads_count early exit;
//   K7  the same walk in reverse over the WHOLE tile list, nine global atomicAdd per contributing
//       (pixel, Gaussian) pair (3 colour, 2 mean, 3 conic, 1 opacity).
// Arithmetic = the PRECISE flavour of the product kernels (same op order), so K6 here is bit-identical
// to blend_forward_kernel<true> and doubles as a second, independent implementation in the parity tests.
// Per-Gaussian state is read from the same 48-byte records K1 writes (the reference keeps separate
// means2D / conic_opacity / rgb arrays: same bytes per gather, 44 vs 48).
#include <cub/device/device_radix_sort.cuh>

#include "blend_common.cuh"

namespace gsr {

size_t ref_sort_temp_bytes(int64_t n, int end_bit) {
  size_t bytes = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, bytes, (const uint64_t*)nullptr, (uint64_t*)nullptr,
                                  (const uint32_t*)nullptr, (uint32_t*)nullptr, (int64_t)(n > 0 ? n : 1), 0, end_bit);
  return align_up(bytes);
}

cudaError_t launch_ref_sort_pairs(cudaStream_t s, int64_t n, const uint64_t* keys_in, const uint32_t* vals_in,
                                  uint64_t* keys_out, uint32_t* vals_out, int end_bit, char* temp, size_t temp_bytes) {
  if (n <= 0) return cudaSuccess;
  count_launch(2 + (end_bit + 7) / 8);  // histogram + scan + onesweep passes (CUB's own kernels)
  return cub::DeviceRadixSort::SortPairs(temp, temp_bytes, keys_in, keys_out, vals_in, vals_out, n, 0, end_bit, s);
}

__global__ void __launch_bounds__(256)
ref_blend_forward_kernel(int W, int H, int grid_x, const uint2* __restrict__ ranges,
                         const uint32_t* __restrict__ point_list, const float4* __restrict__ rec,
                         const float* __restrict__ depths, const float* __restrict__ bg,
                         float* __restrict__ out_color, float* __restrict__ out_depth,
                         float* __restrict__ final_T, uint32_t* __restrict__ n_contrib) {
  __shared__ uint32_t s_id[BLEND_BATCH];
  __shared__ float4 s_q0[BLEND_BATCH];   // x, y, conic.x, conic.y
  __shared__ float2 s_q1[BLEND_BATCH];   // conic.z, opacity
  const int tid = threadIdx.x;
  const int tile = blockIdx.x;
  const int px = (tile % grid_x) * TILE_X + (tid & 15);
  const int py = (tile / grid_x) * TILE_Y + (tid >> 4);
  const bool inside = px < W && py < H;
  const float pxf = (float)px, pyf = (float)py;
  const uint2 range = ranges[tile];
  int todo = (int)(range.y - range.x);
  const int rounds = (todo + BLEND_BATCH - 1) / BLEND_BATCH;

  bool done = !inside;
  float T = 1.0f, C0 = 0.0f, C1 = 0.0f, C2 = 0.0f, Dm = 15.0f;
  uint32_t contributor = 0, last = 0;
  for (int b = 0; b < rounds; b++, todo -= BLEND_BATCH) {
    if (__syncthreads_count(done) == BLEND_BATCH) break;
    const int pos = b * BLEND_BATCH + tid;
    if (pos < (int)(range.y - range.x)) {
      const uint32_t id = point_list[range.x + pos];
      s_id[tid] = id;
      s_q0[tid] = rec[3 * (size_t)id];
      const float4 q1 = rec[3 * (size_t)id + 1];
      s_q1[tid] = make_float2(q1.x, q1.y);
    }
    __syncthreads();
    for (int j = 0; !done && j < min(BLEND_BATCH, todo); j++) {
      contributor++;
      const float4 q0 = s_q0[j];
      const float2 q1 = s_q1[j];
      const float4 e0 = q0, e1 = make_float4(q1.x, q1.y, 0.f, 0.f);
      float dx, dy;
      const float power = pair_power<true>(e0, e1, pxf, pyf, dx, dy);
      if (power > 0.0f) continue;
      const float alpha = fminf(0.99f, MUL(q1.y, expf(power)));
      if (alpha < 1.0f / 255.0f) continue;
      const float test_T = MUL(T, SUB(1.0f, alpha));
      if (test_T < 0.0001f) {
        done = true;
        continue;
      }
      const uint32_t id = s_id[j];
      const float4 q2 = rec[3 * (size_t)id + 2];   // colour from global memory, as features[] upstream
      const float wgt = MUL(alpha, T);
      C0 = FMA(q2.x, wgt, C0);
      C1 = FMA(q2.y, wgt, C1);
      C2 = FMA(q2.z, wgt, C2);
      if (T > 0.5f && test_T < 0.5f) Dm = depths[id];
      T = test_T;
      last = contributor;
    }
  }
  if (inside) {
    const size_t HW = (size_t)H * W, pix = (size_t)py * W + px;
    final_T[pix] = T;
    n_contrib[pix] = last;
    out_color[pix] = FMA(T, bg[0], C0);
    out_color[HW + pix] = FMA(T, bg[1], C1);
    out_color[2 * HW + pix] = FMA(T, bg[2], C2);
    out_depth[pix] = Dm;
  }
}

__global__ void __launch_bounds__(256)
ref_blend_backward_kernel(int W, int H, int grid_x, const uint2* __restrict__ ranges,
                          const uint32_t* __restrict__ point_list, const float4* __restrict__ rec,
                          const float* __restrict__ bg, const float* __restrict__ final_T,
                          const uint32_t* __restrict__ n_contrib, const float* __restrict__ dL_dpix,
                          float* __restrict__ gacc) {
  __shared__ uint32_t s_id[BLEND_BATCH];
  __shared__ float4 s_q0[BLEND_BATCH];
  __shared__ float2 s_q1[BLEND_BATCH];
  __shared__ float4 s_q2[BLEND_BATCH];   // r, g, b
  const int tid = threadIdx.x;
  const int tile = blockIdx.x;
  const int px = (tile % grid_x) * TILE_X + (tid & 15);
  const int py = (tile / grid_x) * TILE_Y + (tid >> 4);
  const bool inside = px < W && py < H;
  const float pxf = (float)px, pyf = (float)py;
  const size_t HW = (size_t)H * W, pix = (size_t)py * W + px;
  const uint2 range = ranges[tile];
  const int total = (int)(range.y - range.x);
  const int rounds = (total + BLEND_BATCH - 1) / BLEND_BATCH;
  int todo = total;

  const float T_final = inside ? final_T[pix] : 0.0f;
  float T = T_final;
  uint32_t contributor = (uint32_t)total;
  const uint32_t last = inside ? n_contrib[pix] : 0u;
  float dLp0 = 0.f, dLp1 = 0.f, dLp2 = 0.f;
  if (inside) {
    dLp0 = dL_dpix[pix];
    dLp1 = dL_dpix[HW + pix];
    dLp2 = dL_dpix[2 * HW + pix];
  }
  const float bg_dot = FMA(bg[2], dLp2, FMA(bg[1], dLp1, MUL(bg[0], dLp0)));
  float acc0 = 0.f, acc1 = 0.f, acc2 = 0.f, lc0 = 0.f, lc1 = 0.f, lc2 = 0.f, last_alpha = 0.f;

  for (int b = 0; b < rounds; b++, todo -= BLEND_BATCH) {
    __syncthreads();
    const int pos = b * BLEND_BATCH + tid;   // back to front: entry range.y - 1 - pos
    if (pos < total) {
      const uint32_t id = point_list[range.y - 1 - pos];
      s_id[tid] = id;
      s_q0[tid] = rec[3 * (size_t)id];
      const float4 q1 = rec[3 * (size_t)id + 1];
      s_q1[tid] = make_float2(q1.x, q1.y);
      s_q2[tid] = rec[3 * (size_t)id + 2];
    }
    __syncthreads();
    for (int j = 0; inside && j < min(BLEND_BATCH, todo); j++) {
      contributor--;
      if (contributor >= last) continue;
      const float4 q0 = s_q0[j];
      const float2 q1 = s_q1[j];
      const float4 e1 = make_float4(q1.x, q1.y, 0.f, 0.f);
      float dx, dy;
      const float power = pair_power<true>(q0, e1, pxf, pyf, dx, dy);
      if (power > 0.0f) continue;
      const float G = expf(power);
      const float alpha = fminf(0.99f, MUL(q1.y, G));
      if (alpha < 1.0f / 255.0f) continue;
      const float4 q2 = s_q2[j];
      const float one_m_alpha = SUB(1.0f, alpha);
      T = DIV(T, one_m_alpha);
      const float dch = MUL(alpha, T);
      acc0 = FMA(last_alpha, lc0, MUL(SUB(1.0f, last_alpha), acc0));
      acc1 = FMA(last_alpha, lc1, MUL(SUB(1.0f, last_alpha), acc1));
      acc2 = FMA(last_alpha, lc2, MUL(SUB(1.0f, last_alpha), acc2));
      float dL_dalpha = MUL(SUB(q2.x, acc0), dLp0);
      dL_dalpha = FMA(SUB(q2.y, acc1), dLp1, dL_dalpha);
      dL_dalpha = FMA(SUB(q2.z, acc2), dLp2, dL_dalpha);
      dL_dalpha = FMA(dL_dalpha, T, MUL(DIV(-T_final, one_m_alpha), bg_dot));
      lc0 = q2.x; lc1 = q2.y; lc2 = q2.z;
      last_alpha = alpha;
      const float w = MUL(MUL(q1.y, dL_dalpha), G);   // dL_dG * G
      const float wx = MUL(w, dx), wy = MUL(w, dy);
      float* dst = gacc + (size_t)s_id[j] * 12;
      atomicAdd(dst + 0, MUL(dch, dLp0));
      atomicAdd(dst + 1, MUL(dch, dLp1));
      atomicAdd(dst + 2, MUL(dch, dLp2));
      atomicAdd(dst + 3, wx);
      atomicAdd(dst + 4, wy);
      atomicAdd(dst + 5, MUL(-0.5f * dx, wx));
      atomicAdd(dst + 6, MUL(-0.5f * dy, wx));
      atomicAdd(dst + 7, MUL(-0.5f * dy, wy));
      atomicAdd(dst + 8, MUL(G, dL_dalpha));
    }
  }
}

cudaError_t launch_ref_blend_forward(cudaStream_t s, int W, int H, const uint2* ranges, const uint32_t* point_list,
                                     const float4* rec, const float* depths, const float* bg, float* out_color,
                                     float* out_depth, float* final_T, uint32_t* n_contrib) {
  const int gx = cdiv(W, TILE_X), gy = cdiv(H, TILE_Y);
  if (gx * gy == 0) return cudaSuccess;
  ref_blend_forward_kernel<<<gx * gy, 256, 0, s>>>(W, H, gx, ranges, point_list, rec, depths, bg, out_color,
                                                   out_depth, final_T, n_contrib);
  count_launch();
  return cudaGetLastError();
}

cudaError_t launch_ref_blend_backward(cudaStream_t s, int W, int H, const uint2* ranges, const uint32_t* point_list,
                                      const float4* rec, const float* bg, const float* final_T,
                                      const uint32_t* n_contrib, const float* dL_dpix, float* gacc) {
  const int gx = cdiv(W, TILE_X), gy = cdiv(H, TILE_Y);
  if (gx * gy == 0) return cudaSuccess;
  ref_blend_backward_kernel<<<gx * gy, 256, 0, s>>>(W, H, gx, ranges, point_list, rec, bg, final_T, n_contrib,
                                                    dL_dpix, gacc);
  count_launch();
  return cudaGetLastError();
}

}  // namespace gsr
