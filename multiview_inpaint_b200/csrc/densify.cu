// densify.cu -- the arena side of densification (the part DESIGN.md section 8 listed as open):
// one row-gather launch re-packs the parameter arena AND both Adam moment arenas after clone / split / prune.
//
// Replaces, in gs-simp/scene/gaussian_model.py, what densify_and_prune (:467-480) does to the six parameter tensors
// and their optimizer state through densification_postfix (:408-425) -> cat_tensors_to_optimizer (:385-406) twice
// (clone, split) and prune_points (:365-383) -> _prune_optimizer (:346-363) twice (split parents, final prune): four
// rounds of `torch.cat` / boolean-mask indexing over 6 tensors x {param, exp_avg, exp_avg_sq} = 72 torch kernels that
// each copy the whole model.  The host side (densify.py) composes the four steps into ONE index map
// src_row[new row] = old row; here every destination row is then written exactly once:
//     dst[i, :] = src[src_row[i], :]                       for parameters
//     dst[i, :] = i < n_keep_state ? src[src_row[i], :] : 0 for the Adam moments (cloned / split rows start from zero
//                                                           state: cat_tensors_to_optimizer appends zeros_like)
// for up to GSR_GATHER_MAX_SEGS segments (5 slices x 3 arenas) in the same launch.
//
// HBM-bound: (4 + 2 * 12 * row_f32 summed over segments) bytes per destination row (59 floats x 3 arenas at M = 16:
// 1.42 KB).  A CTA owns 64 consecutive destination rows (their source rows staged once in shared memory); within a
// segment the CTA's threads walk the 64 x row_f32 destination floats linearly, so stores are fully coalesced and
// loads are contiguous per source row (src_row is monotone inside each of the three blocks kept / clones / children,
// so neighbouring rows mostly share sectors).  Rows whose length is a multiple of 4 floats (SH rows of M = 4 / 16,
// rotations) move as float4.  Pure data movement: bit-exact.
//
// Bulk (TMA) copies for the wide rows.  The SH segments -- 192-byte rows at M = 16, 81 % of the bytes -- need no thread to
// touch the data: thread r of the CTA issues ONE `cp.async.bulk` global -> shared for source row src_row[row0 + r] (16-byte
// aligned, a multiple of 16 bytes), all 64 complete on one mbarrier (expect_tx = the bytes of the rows that copy), rows that
// start from zero state are filled by plain stores + `fence.proxy.async`, and ONE thread sends the CTA's 64 destination rows
// (contiguous: 12 KB) back with a single `cp.async.bulk` shared -> global.  The gather itself has no tile a tensor map could
// describe (rows are picked by index), so the 1-D form is the one that applies; it replaces 768 LDG.128 + 768 STG.128 and
// their address arithmetic per tile and segment by 64 + 1 copy instructions.  Segments whose rows are narrower than 48 bytes
// or not a multiple of 16 (xyz, opacity, scaling; rotations at 16 bytes would be 64 sixteen-byte TMA operations) keep the
// thread path.  gsr_debug_set knob 5 = 0 turns the bulk path off (A/B timing, cross-check in tests/test_densify_gpu.py).
#include "common.cuh"

namespace gsr {

namespace {

constexpr int GATHER_ROWS = 64;
constexpr int GATHER_TMA_MAX_ROW_BYTES = 192;

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("{ .reg .b64 st; mbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1; }" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                 : "=r"(done) : "r"(bar), "r"(parity) : "memory");
  } while (!done);
}
__device__ __forceinline__ void bulk_load(uint32_t smem_dst, const void* gsrc, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_dst), "l"(gsrc), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void bulk_store(void* gdst, uint32_t smem_src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(smem_src), "r"(bytes) : "memory");
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

struct GatherArgs {
  const float* src[GSR_GATHER_MAX_SEGS];
  float* dst[GSR_GATHER_MAX_SEGS];
  int row_f32[GSR_GATHER_MAX_SEGS];
  int zero_new[GSR_GATHER_MAX_SEGS];
  int bulk[GSR_GATHER_MAX_SEGS];   // rows of this segment move by cp.async.bulk (launch_gather_rows decides)
  int n_segs;
};

__global__ void __launch_bounds__(256)
gather_rows_kernel(GatherArgs a, long long n_dst, long long n_keep_state, const int* __restrict__ src_row) {
  __shared__ int s_src[GATHER_ROWS];
  __shared__ __align__(128) unsigned char s_rows[GATHER_ROWS * GATHER_TMA_MAX_ROW_BYTES];   // bulk path: the CTA's 64 rows of one segment
  __shared__ __align__(8) unsigned long long s_bar;
  const uint32_t rows_s = (uint32_t)__cvta_generic_to_shared(s_rows), bar = (uint32_t)__cvta_generic_to_shared(&s_bar);
  uint32_t parity = 0;
  if (threadIdx.x == 0) mbar_init(bar, 1);
  __syncthreads();
  const long long n_tiles = (n_dst + GATHER_ROWS - 1) / GATHER_ROWS;
  for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const long long row0 = tile * GATHER_ROWS;
    const int rows = (int)min((long long)GATHER_ROWS, n_dst - row0);
    __syncthreads();  // WAR on s_src (previous tile)
    if (threadIdx.x < rows) s_src[threadIdx.x] = __ldg(src_row + row0 + threadIdx.x);
    __syncthreads();
    // rows of this tile that keep their optimizer state: [0, keep)
    const int keep = (int)max(0LL, min((long long)rows, n_keep_state - row0));
    for (int s = 0; s < a.n_segs; s++) {
      const int L = a.row_f32[s];
      const float* __restrict__ src = a.src[s];
      float* __restrict__ dst = a.dst[s];
      const int live = a.zero_new[s] ? keep : rows;   // rows below `live` copy, the rest are zero-filled
      if (a.bulk[s]) {
        const uint32_t row_b = (uint32_t)L * 4;
        if (threadIdx.x == 0) {
          bulk_wait_read();                         // the previous bulk store has finished reading s_rows
          mbar_arrive_expect_tx(bar, (uint32_t)live * row_b);
        }
        __syncthreads();
        if ((int)threadIdx.x < live)
          bulk_load(rows_s + threadIdx.x * row_b, reinterpret_cast<const char*>(src) + (size_t)s_src[threadIdx.x] * row_b, row_b, bar);
        // rows that start from zero state: 16-byte stores through the generic proxy
        const int nz4 = (rows - live) * (L >> 2);
        for (int e = threadIdx.x; e < nz4; e += blockDim.x)
          reinterpret_cast<float4*>(s_rows + (size_t)live * row_b)[e] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
        mbar_wait(bar, parity);
        parity ^= 1u;
        fence_proxy_async_smem();                   // the zero rows, before the async proxy reads them
        __syncthreads();
        if (threadIdx.x == 0) bulk_store(reinterpret_cast<char*>(dst) + (size_t)row0 * row_b, rows_s, (uint32_t)rows * row_b);
        continue;
      }
      if ((L & 3) == 0) {
        const int Lv = L >> 2;
        const float4* __restrict__ src4 = reinterpret_cast<const float4*>(src);
        float4* __restrict__ dst4 = reinterpret_cast<float4*>(dst) + (size_t)row0 * Lv;
        const int n = rows * Lv;
        for (int e = threadIdx.x; e < n; e += blockDim.x) {
          const int r = e / Lv, c = e - r * Lv;
          float4 v = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
          if (r < live) v = __ldg(src4 + (size_t)s_src[r] * Lv + c);
          dst4[e] = v;
        }
      } else {
        float* __restrict__ d = dst + (size_t)row0 * L;
        const int n = rows * L;
        for (int e = threadIdx.x; e < n; e += blockDim.x) {
          const int r = e / L, c = e - r * L;
          float v = 0.0f;
          if (r < live) v = __ldg(src + (size_t)s_src[r] * L + c);
          d[e] = v;
        }
      }
    }
  }
  if (threadIdx.x == 0) bulk_wait_all();
}

}  // namespace

int g_gather_bulk = 1;   // gsr_debug_set knob 5

cudaError_t launch_gather_rows(cudaStream_t stream, long long n_dst, long long n_keep_state, const int* src_row,
                               const gsr_gather_segment* segs, int n_segs) {
  GatherArgs a;
  a.n_segs = n_segs;
  for (int s = 0; s < n_segs; s++) {
    a.src[s] = segs[s].src;
    a.dst[s] = segs[s].dst;
    a.row_f32[s] = segs[s].row_f32;
    a.zero_new[s] = segs[s].zero_new;
    const size_t row_b = (size_t)segs[s].row_f32 * 4;
    a.bulk[s] = g_gather_bulk && row_b >= 48 && row_b <= GATHER_TMA_MAX_ROW_BYTES && (row_b & 15) == 0 &&
                ((reinterpret_cast<uintptr_t>(segs[s].src) | reinterpret_cast<uintptr_t>(segs[s].dst)) & 15) == 0;
  }
  for (int s = n_segs; s < GSR_GATHER_MAX_SEGS; s++) {
    a.src[s] = nullptr; a.dst[s] = nullptr; a.row_f32[s] = 0; a.zero_new[s] = 0; a.bulk[s] = 0;
  }
  const long long n_tiles = (n_dst + GATHER_ROWS - 1) / GATHER_ROWS;
  const int grid = (int)min(n_tiles, (long long)148 * 8);   // 8 resident CTAs of 256 threads per SM, grid-stride beyond
  gather_rows_kernel<<<grid, 256, 0, stream>>>(a, n_dst, n_keep_state, src_row);
  count_launch();
  return cudaGetLastError();
}

}  // namespace gsr
