// nvls_allreduce.cu -- in-switch all-reduce of the multi-view gradient arena over NVLink / NVSwitch.
//
// New capability (the reference never communicates on this path).  After a
// rank has run its share of the views, its gradient arena and densification statistics must be summed
// (max-ed) over ranks.  The arena lives in SYMMETRIC memory with an NVSwitch MULTICAST mapping; rank r
// owns slice r of every segment and, for each 16 bytes of it, issues
//     multimem.ld_reduce.add.v4.f32   -- the switch reads the element from all R replicas and adds them
//     multimem.st.v4.f32              -- the switch writes the sum back into all R replicas
// so every byte crosses each GPU's links once in and once out (ring all-reduce: 2 (R-1)/R times) and no
// staging copies exist.  Integer segments use the scalar .add.s32 / .max.s32 forms (12 B per Gaussian).
// The caller brackets the launch with two cross-rank barriers on the stream (all replicas complete
// before anyone reads; all slices written before anyone continues).
#include "common.cuh"

namespace gsr {

__device__ __forceinline__ float4 mm_ld_reduce_add_f32x4(const void* mc) {
  float4 v;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(mc) : "memory");
  return v;
}
__device__ __forceinline__ void mm_st_f32x4(void* mc, float4 v) {
  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};"
               ::"l"(mc), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ int mm_ld_reduce_add_s32(const void* mc) {
  int v;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.s32 %0, [%1];" : "=r"(v) : "l"(mc) : "memory");
  return v;
}
__device__ __forceinline__ int mm_ld_reduce_max_s32(const void* mc) {
  int v;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.max.s32 %0, [%1];" : "=r"(v) : "l"(mc) : "memory");
  return v;
}
__device__ __forceinline__ void mm_st_s32(void* mc, int v) {
  asm volatile("multimem.st.relaxed.sys.global.s32 [%0], %1;" ::"l"(mc), "r"(v) : "memory");
}

// Dense part of the float segment: float4 indices [lo4, hi4), rank's slice of it, 4 reductions in flight.
__device__ __forceinline__ void reduce_dense_f32(float4* base, size_t lo4, size_t hi4, int rank, int world, size_t tid,
                                                 size_t nthreads) {
  const size_t n = hi4 - lo4;
  const size_t per = (n + world - 1) / world;
  const size_t lo = lo4 + min(n, per * rank), hi = min(hi4, lo + per);
  constexpr int U = 4;
  for (size_t i = lo + tid; i < hi; i += U * nthreads) {
    float4 v[U];
#pragma unroll
    for (int u = 0; u < U; u++)
      if (i + u * nthreads < hi) v[u] = mm_ld_reduce_add_f32x4(base + i + u * nthreads);
#pragma unroll
    for (int u = 0; u < U; u++)
      if (i + u * nthreads < hi) mm_st_f32x4(base + i + u * nthreads, v[u]);
  }
}

// Float4 indices [row_lo4, row_hi4) of the float segment form a (rows x row_f4) matrix -- the (P, M, 3) SH
// gradient, 81 % of the arena at M = 16 -- whose row g is all-zero on every rank exactly when the Gaussian
// was visible in no view of any rank, i.e. when the SUM over ranks of element g of the add_s32 segment (the
// visibility count) is zero.  Such rows are skipped.  A warp takes a group of 96 / row_f4 rows (96 float4,
// three per lane, lane-contiguous so every access is a coalesced 512-byte run); its first lanes fetch the
// group's counts with one 4-byte in-switch reduction each and broadcast them by shuffle.  (The count segment
// is concurrently being replaced by its own sum, which can only turn a positive value into a larger one, so
// the `!= 0` test is race free.)
__global__ void __launch_bounds__(512)
nvls_allreduce_kernel(char* __restrict__ mc, size_t off_f32, size_t n_f32x4, size_t off_add_s32, size_t n_add_s32,
                      size_t off_max_s32, size_t n_max_s32, int rank, int world, size_t row_lo4, size_t row_hi4,
                      unsigned row_f4) {
  const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t nthreads = (size_t)gridDim.x * blockDim.x;
  float4* base = reinterpret_cast<float4*>(mc + off_f32);
  if (row_f4 == 0) {
    reduce_dense_f32(base, 0, n_f32x4, rank, world, tid, nthreads);
  } else {
    reduce_dense_f32(base, 0, row_lo4, rank, world, tid, nthreads);
    reduce_dense_f32(base, row_hi4, n_f32x4, rank, world, tid, nthreads);
    const int* cnt = reinterpret_cast<const int*>(mc + off_add_s32);
    const unsigned lane = threadIdx.x & 31;
    const unsigned G = 96 / row_f4;                               // rows per group (8 at M = 16, 32 at M = 4)
    const size_t rows = (row_hi4 - row_lo4) / row_f4;
    const size_t groups = (rows + G - 1) / G;
    const size_t per = (groups + world - 1) / world;
    const size_t g_lo = min(groups, per * rank), g_hi = min(groups, g_lo + per);
    const size_t warp_id = tid >> 5, n_warps = nthreads >> 5;
    for (size_t g = g_lo + warp_id; g < g_hi; g += n_warps) {     // warp-uniform
      const size_t row0 = g * G;
      int c = 0;
      if (lane < G && row0 + lane < rows) c = mm_ld_reduce_add_s32(cnt + row0 + lane);
      float4* p = base + row_lo4 + row0 * row_f4;
      float4 v[3];
      bool live[3];
#pragma unroll
      for (int k = 0; k < 3; k++) {
        const unsigned q = lane + 32 * k;                         // float4 index inside the group
        const unsigned r = q / row_f4;
        const int cr = __shfl_sync(0xffffffffu, c, r);
        live[k] = cr != 0 && row0 + r < rows;
      }
#pragma unroll
      for (int k = 0; k < 3; k++)
        if (live[k]) v[k] = mm_ld_reduce_add_f32x4(p + lane + 32 * k);
#pragma unroll
      for (int k = 0; k < 3; k++)
        if (live[k]) mm_st_f32x4(p + lane + 32 * k, v[k]);
    }
  }
  {
    const size_t per = (n_add_s32 + world - 1) / world;
    const size_t lo = min(n_add_s32, per * rank), hi = min(n_add_s32, lo + per);
    int* base = reinterpret_cast<int*>(mc + off_add_s32);
    for (size_t i = lo + tid; i < hi; i += nthreads) mm_st_s32(base + i, mm_ld_reduce_add_s32(base + i));
  }
  {
    const size_t per = (n_max_s32 + world - 1) / world;
    const size_t lo = min(n_max_s32, per * rank), hi = min(n_max_s32, lo + per);
    int* base = reinterpret_cast<int*>(mc + off_max_s32);
    for (size_t i = lo + tid; i < hi; i += nthreads) mm_st_s32(base + i, mm_ld_reduce_max_s32(base + i));
  }
}

cudaError_t launch_nvls_allreduce(cudaStream_t s, char* mc, size_t off_f32, size_t n_f32, size_t off_add_s32,
                                  size_t n_add_s32, size_t off_max_s32, size_t n_max_s32, int rank, int world,
                                  int blocks, size_t sparse_first_f32, size_t sparse_rows, int sparse_row_f32) {
  if (blocks <= 0) blocks = 148 * 2;
  size_t row_lo4 = 0, row_hi4 = 0;
  unsigned row_f4 = 0;
  if (sparse_rows > 0 && (sparse_row_f32 == 12 || sparse_row_f32 == 48) && (sparse_first_f32 & 3) == 0 &&
      sparse_rows <= n_add_s32) {
    row_f4 = (unsigned)sparse_row_f32 / 4;
    row_lo4 = sparse_first_f32 / 4;
    row_hi4 = row_lo4 + sparse_rows * row_f4;
  }
  nvls_allreduce_kernel<<<blocks, 512, 0, s>>>(mc, off_f32, n_f32 / 4, off_add_s32, n_add_s32, off_max_s32, n_max_s32,
                                               rank, world, row_lo4, row_hi4, row_f4);
  count_launch();
  return cudaGetLastError();
}

}  // namespace gsr
