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

// Plan of one launch (by value): up to six dense float4 ranges, one row-sparse matrix, two int ranges; all
// positions are relative to the multicast base.
struct NvlsPlan {
  unsigned long long dense_lo4[6], dense_hi4[6];
  int n_dense;
  unsigned long long row_lo4, rows;   // float4 index of row 0, number of rows
  unsigned row_f4;                    // float4 per row (0: no sparse matrix)
  unsigned long long cnt_off;         // byte offset of the int32 count that belongs to row 0
  unsigned long long add_off, n_add, max_off, n_max;
};

// The sparse matrix is a (rows x row_f4) block of float4 -- the (P, M, 3) SH gradient, 81 % of the arena at
// M = 16 -- whose row g is all-zero on every rank exactly when the Gaussian was visible in no view of any
// rank, i.e. when the SUM over ranks of its visibility count is zero.  Such rows are skipped.  A warp takes a
// group of 96 / row_f4 rows (96 float4, three per lane, lane-contiguous so every access is a coalesced
// 512-byte run); its first lanes fetch the group's counts with one 4-byte in-switch reduction each and
// broadcast them by shuffle.  (The count segment is concurrently being replaced by its own sum, which can
// only turn a positive value into a larger one, so the `!= 0` test is race free.)
__global__ void __launch_bounds__(512)
nvls_allreduce_kernel(char* __restrict__ mc, const NvlsPlan pl, int rank, int world) {
  const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t nthreads = (size_t)gridDim.x * blockDim.x;
  float4* base = reinterpret_cast<float4*>(mc);
  for (int d = 0; d < pl.n_dense; d++) reduce_dense_f32(base, pl.dense_lo4[d], pl.dense_hi4[d], rank, world, tid, nthreads);
  if (pl.row_f4 != 0) {
    const unsigned row_f4 = pl.row_f4;
    const int* cnt = reinterpret_cast<const int*>(mc + pl.cnt_off);
    const unsigned lane = threadIdx.x & 31;
    const unsigned G = 96 / row_f4;                               // rows per group (8 at M = 16, 32 at M = 4)
    const size_t rows = pl.rows;
    const size_t groups = (rows + G - 1) / G;
    const size_t per = (groups + world - 1) / world;
    const size_t g_lo = min(groups, per * rank), g_hi = min(groups, g_lo + per);
    const size_t warp_id = tid >> 5, n_warps = nthreads >> 5;
    // The group's counts are fetched ONE ITERATION AHEAD: the count reduction is a full NVLink round trip, and issued in
    // front of the loads that depend on it, it would double the latency of every iteration of this latency-bound loop.
    auto fetch_counts = [&](size_t g) -> int {
      const size_t r = g * G + lane;
      return (g < g_hi && lane < G && r < rows) ? mm_ld_reduce_add_s32(cnt + r) : 0;
    };
    int c_next = fetch_counts(g_lo + warp_id);
    for (size_t g = g_lo + warp_id; g < g_hi; g += n_warps) {     // warp-uniform
      const size_t row0 = g * G;
      const int c = c_next;
      c_next = fetch_counts(g + n_warps);
      float4* p = base + pl.row_lo4 + row0 * row_f4;
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
    const size_t per = (pl.n_add + world - 1) / world;
    const size_t lo = min((size_t)pl.n_add, per * rank), hi = min((size_t)pl.n_add, lo + per);
    int* b = reinterpret_cast<int*>(mc + pl.add_off);
    for (size_t i = lo + tid; i < hi; i += nthreads) mm_st_s32(b + i, mm_ld_reduce_add_s32(b + i));
  }
  {
    const size_t per = (pl.n_max + world - 1) / world;
    const size_t lo = min((size_t)pl.n_max, per * rank), hi = min((size_t)pl.n_max, lo + per);
    int* b = reinterpret_cast<int*>(mc + pl.max_off);
    for (size_t i = lo + tid; i < hi; i += nthreads) mm_st_s32(b + i, mm_ld_reduce_max_s32(b + i));
  }
}

// ---- two ranks: peer-to-peer two-shot over NVLink -----------------------------------------------------------------------
// At two GPUs the switch has nothing to reduce that a peer load cannot: with multimem every 16 bytes still cross each GPU's
// links once out and once in (even the local replica's share goes through the switch), and the per-rank multimem rate -- not
// the links -- bounds the kernel (1.56 ms for the headline arena against NCCL's 1.42).  Here rank r owns half of every segment:
// it LOADS the peer's element over NVLink (the arena is symmetric memory: the peer's replica is mapped into this address
// space), adds its own, and STORES the sum into both replicas -- the same bytes on the links, plain loads / stores at the link
// rate, the same plan / row-sparsity / barriers as the in-switch kernel.  Peer accesses are volatile (no L1 line of a
// previous launch may answer them); ordering against the peer's kernels comes from the two cross-rank barriers around the launch.
__device__ __forceinline__ float4 ld_peer_f4(const float4* p) {
  float4 v;
  asm volatile("ld.volatile.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ int ld_peer_s32(const int* p) {
  int v;
  asm volatile("ld.volatile.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ float4 add4(float4 a, float4 b) { return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }

__device__ __forceinline__ void p2p_dense_f32(float4* loc, float4* peer, size_t lo4, size_t hi4, int rank, size_t tid, size_t nthreads) {
  const size_t n = hi4 - lo4;
  const size_t per = (n + 1) / 2;
  const size_t lo = lo4 + min(n, per * rank), hi = min(hi4, lo + per);
  constexpr int U = 4;
  for (size_t i = lo + tid; i < hi; i += U * nthreads) {
    float4 v[U];
#pragma unroll
    for (int u = 0; u < U; u++)
      if (i + u * nthreads < hi) v[u] = ld_peer_f4(peer + i + u * nthreads);
#pragma unroll
    for (int u = 0; u < U; u++)
      if (i + u * nthreads < hi) {
        const float4 r = add4(loc[i + u * nthreads], v[u]);
        loc[i + u * nthreads] = r;
        peer[i + u * nthreads] = r;
      }
  }
}

__global__ void __launch_bounds__(512)
p2p_allreduce2_kernel(char* __restrict__ loc_c, char* __restrict__ peer_c, const NvlsPlan pl, int rank) {
  const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t nthreads = (size_t)gridDim.x * blockDim.x;
  float4* loc = reinterpret_cast<float4*>(loc_c);
  float4* peer = reinterpret_cast<float4*>(peer_c);
  for (int d = 0; d < pl.n_dense; d++) p2p_dense_f32(loc, peer, pl.dense_lo4[d], pl.dense_hi4[d], rank, tid, nthreads);
  if (pl.row_f4 != 0) {
    // same walk as the in-switch kernel: a warp takes a group of 96 / row_f4 rows, the group's visibility counts (own + peer;
    // either may already hold the sum, which keeps a positive value positive) decide which rows move at all
    const unsigned row_f4 = pl.row_f4;
    const int* cnt_l = reinterpret_cast<const int*>(loc_c + pl.cnt_off);
    const int* cnt_p = reinterpret_cast<const int*>(peer_c + pl.cnt_off);
    const unsigned lane = threadIdx.x & 31;
    const unsigned G = 96 / row_f4;
    const size_t rows = pl.rows;
    const size_t groups = (rows + G - 1) / G;
    const size_t per = (groups + 1) / 2;
    const size_t g_lo = min(groups, per * rank), g_hi = min(groups, g_lo + per);
    const size_t warp_id = tid >> 5, n_warps = nthreads >> 5;
    auto fetch_counts = [&](size_t g) -> int {
      const size_t r = g * G + lane;
      return (g < g_hi && lane < G && r < rows) ? (ld_peer_s32(cnt_l + r) | ld_peer_s32(cnt_p + r)) : 0;   // counts are >= 0: OR != 0 <=> sum != 0
    };
    int c_next = fetch_counts(g_lo + warp_id);
    for (size_t g = g_lo + warp_id; g < g_hi; g += n_warps) {     // warp-uniform
      const size_t row0 = g * G;
      const int c = c_next;
      c_next = fetch_counts(g + n_warps);
      const size_t p0 = pl.row_lo4 + row0 * row_f4;
      float4 v[3];
      bool live[3];
#pragma unroll
      for (int k = 0; k < 3; k++) {
        const unsigned q = lane + 32 * k;
        const unsigned r = q / row_f4;
        const int cr = __shfl_sync(0xffffffffu, c, r);
        live[k] = cr != 0 && row0 + r < rows;
      }
#pragma unroll
      for (int k = 0; k < 3; k++)
        if (live[k]) v[k] = ld_peer_f4(peer + p0 + lane + 32 * k);
#pragma unroll
      for (int k = 0; k < 3; k++)
        if (live[k]) {
          const float4 r = add4(loc[p0 + lane + 32 * k], v[k]);
          loc[p0 + lane + 32 * k] = r;
          peer[p0 + lane + 32 * k] = r;
        }
    }
  }
  {
    const size_t per = (pl.n_add + 1) / 2;
    const size_t lo = min((size_t)pl.n_add, per * rank), hi = min((size_t)pl.n_add, lo + per);
    int* bl = reinterpret_cast<int*>(loc_c + pl.add_off);
    int* bp = reinterpret_cast<int*>(peer_c + pl.add_off);
    for (size_t i = lo + tid; i < hi; i += nthreads) {
      const int r = ld_peer_s32(bl + i) + ld_peer_s32(bp + i);
      bl[i] = r;
      bp[i] = r;
    }
  }
  {
    const size_t per = (pl.n_max + 1) / 2;
    const size_t lo = min((size_t)pl.n_max, per * rank), hi = min((size_t)pl.n_max, lo + per);
    int* bl = reinterpret_cast<int*>(loc_c + pl.max_off);
    int* bp = reinterpret_cast<int*>(peer_c + pl.max_off);
    for (size_t i = lo + tid; i < hi; i += nthreads) {
      const int r = max(ld_peer_s32(bl + i), ld_peer_s32(bp + i));
      bl[i] = r;
      bp[i] = r;
    }
  }
}

static cudaError_t make_plan(const gsr_nvls_plan& in, NvlsPlan& pl);

cudaError_t launch_p2p_allreduce_plan(cudaStream_t s, char* local, char* peer, const gsr_nvls_plan& in, int rank, int blocks) {
  if (blocks <= 0) blocks = 148 * 2;
  NvlsPlan pl{};
  cudaError_t e = make_plan(in, pl);
  if (e != cudaSuccess) return e;
  p2p_allreduce2_kernel<<<blocks, 512, 0, s>>>(local, peer, pl, rank);
  count_launch();
  return cudaGetLastError();
}

static cudaError_t make_plan(const gsr_nvls_plan& in, NvlsPlan& pl) {
  for (int d = 0; d < in.n_dense && d < 6; d++) {
    pl.dense_lo4[pl.n_dense] = in.dense_off[d] / 16;
    pl.dense_hi4[pl.n_dense] = in.dense_off[d] / 16 + in.dense_n_f32[d] / 4;
    pl.n_dense++;
  }
  if (in.rows > 0 && (in.row_f32 == 12 || in.row_f32 == 48)) {
    pl.row_f4 = (unsigned)in.row_f32 / 4;
    pl.row_lo4 = in.rows_off / 16;
    pl.rows = in.rows;
    pl.cnt_off = in.rows_count_off;
  } else if (in.rows > 0) {   // row width the sparse walk does not cover: reduce the matrix densely
    if (pl.n_dense >= 6) return cudaErrorInvalidValue;
    pl.dense_lo4[pl.n_dense] = in.rows_off / 16;
    pl.dense_hi4[pl.n_dense] = in.rows_off / 16 + (in.rows * (size_t)in.row_f32 + 3) / 4;
    pl.n_dense++;
  }
  pl.add_off = in.add_s32_off; pl.n_add = in.n_add_s32;
  pl.max_off = in.max_s32_off; pl.n_max = in.n_max_s32;
  return cudaSuccess;
}

cudaError_t launch_nvls_allreduce_plan(cudaStream_t s, char* mc, const gsr_nvls_plan& in, int rank, int world, int blocks) {
  if (blocks <= 0) blocks = 148 * 2;
  NvlsPlan pl{};
  cudaError_t e = make_plan(in, pl);
  if (e != cudaSuccess) return e;
  nvls_allreduce_kernel<<<blocks, 512, 0, s>>>(mc, pl, rank, world);
  count_launch();
  return cudaGetLastError();
}

cudaError_t launch_nvls_allreduce(cudaStream_t s, char* mc, size_t off_f32, size_t n_f32, size_t off_add_s32,
                                  size_t n_add_s32, size_t off_max_s32, size_t n_max_s32, int rank, int world,
                                  int blocks, size_t sparse_first_f32, size_t sparse_rows, int sparse_row_f32) {
  gsr_nvls_plan in{};
  if (sparse_rows > 0 && (sparse_row_f32 == 12 || sparse_row_f32 == 48) && (sparse_first_f32 & 3) == 0 &&
      sparse_rows <= n_add_s32) {
    const size_t mat = sparse_rows * (size_t)sparse_row_f32;
    in.n_dense = 2;
    in.dense_off[0] = off_f32;                                  in.dense_n_f32[0] = sparse_first_f32;
    in.dense_off[1] = off_f32 + 4 * (sparse_first_f32 + mat);   in.dense_n_f32[1] = n_f32 - sparse_first_f32 - mat;
    in.rows_off = off_f32 + 4 * sparse_first_f32;
    in.rows = sparse_rows;
    in.row_f32 = sparse_row_f32;
    in.rows_count_off = off_add_s32;
  } else {
    in.n_dense = 1;
    in.dense_off[0] = off_f32;
    in.dense_n_f32[0] = n_f32;
  }
  in.add_s32_off = off_add_s32; in.n_add_s32 = n_add_s32;
  in.max_s32_off = off_max_s32; in.n_max_s32 = n_max_s32;
  return launch_nvls_allreduce_plan(s, mc, in, rank, world, blocks);
}

}  // namespace gsr
