// sh_rows.cuh -- cooperative staging of (rows x 3M) SH blocks between global and shared memory,
// shared by the single-view and the multi-view per-Gaussian backward kernels.
#pragma once

#include "common.cuh"

namespace gsr {

// ---- fast staging for R = 3M with R % 4 == 0 (M = 4, 16): compile-time geometry ------------------
// A row is R/4 whole float4s, so a 16-byte global access never straddles two Gaussians.  Rows sit in
// shared memory with an ODD stride in float4 units (13 for R = 48, 3 for R = 12): 16-byte aligned
// (LDS.128 / STS.128 on both sides) and conflict free for the per-thread row walk (the 8 threads of
// a quarter warp hit the 8 distinct 4-bank groups).
__host__ __device__ constexpr int row_stride(int R) { return ((R / 4) % 2 ? R / 4 : R / 4 + 1) * 4; }
template <int R, int NT = 256>
__device__ __forceinline__ void rows_load(const float* __restrict__ g, float* s_tile, const uint8_t* s_vis, int rows) {
  constexpr int Q = R / 4, STRIDE = row_stride(R), U = 4;
  const float4* g4 = reinterpret_cast<const float4*>(g);
  const int n4 = rows * Q;
  for (int base = threadIdx.x; base < n4; base += U * NT) {
    float4 q[U];
    int so[U];
#pragma unroll
    for (int u = 0; u < U; u++) {
      const int e4 = base + u * NT;
      const int row = e4 / Q;
      so[u] = -1;
      if (e4 < n4 && s_vis[row]) {
        so[u] = row * STRIDE + (e4 - row * Q) * 4;
        q[u] = __ldg(g4 + e4);
      }
    }
#pragma unroll
    for (int u = 0; u < U; u++)
      if (so[u] >= 0) *reinterpret_cast<float4*>(s_tile + so[u]) = q[u];
  }
}
// ACC: one vector reduction (RED.ADD.F32x4: fire and forget, the L2 does the read-modify-write) per
// 16 bytes of a visible row, columns >= cols_used (coefficients above the active degree) skipped.
// !ACC: every element of the span is written, zeros for culled rows.
template <int R, bool ACC, int NT = 256>
__device__ __forceinline__ void rows_store(float* __restrict__ g, const float* s_tile, const uint8_t* s_vis, int rows,
                                           int cols_used) {
  constexpr int Q = R / 4, STRIDE = row_stride(R);
  float4* g4 = reinterpret_cast<float4*>(g);
  const int n4 = rows * Q;
#pragma unroll 4
  for (int e4 = threadIdx.x; e4 < n4; e4 += NT) {
    const int row = e4 / Q, c = (e4 - row * Q) * 4;
    const bool v = s_vis[row] != 0;
    if (ACC) {
      if (v && c < cols_used) atomicAdd(g4 + e4, *reinterpret_cast<const float4*>(s_tile + row * STRIDE + c));
    } else {
      g4[e4] = v ? *reinterpret_cast<const float4*>(s_tile + row * STRIDE + c) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
}


}  // namespace gsr
