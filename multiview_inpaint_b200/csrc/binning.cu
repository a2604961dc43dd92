// binning.cu -- K3 (duplicateWithKeys) and K5 (identifyTileRanges).
//
// Restates SURVEY.md Appendix A.3 / A.4.  Two flavours of K3:
//   * key64  : the reference's layout -- one (tile << 32 | depth_bits, gaussian) pair per touched
//              tile, emitted in Gaussian-index order; a 64-bit sort on [0, 32+bit) follows.
//   * tiles  : two-level scheme -- Gaussians were already stably sorted by depth bits, thread k
//              handles the k-th nearest Gaussian and emits (tile, gaussian) in that order; a stable
//              sort on the tile id alone then yields EXACTLY the list the 64-bit sort would:
//              within a tile, depth ascending, ties by ascending Gaussian index.
// The rect is recomputed from (mean2D, radius) with the same explicit fp32 sequence as K1.
#include "common.cuh"

namespace gsr {

__global__ void __launch_bounds__(256)
duplicate_key64_kernel(int P, const float4* __restrict__ rec, const float* __restrict__ depths,
                       const uint32_t* __restrict__ offsets, const int32_t* __restrict__ radii,
                       int gx, int gy, uint64_t* __restrict__ keys, uint32_t* __restrict__ vals,
                       uint32_t cap, int32_t* __restrict__ status) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P) return;
  if (i == P - 1 && offsets[i] > cap) atomicExch(status + 1, 1);  // capacity overflow
  const int r = radii[i];
  if (r <= 0) return;
  uint32_t off = (i == 0) ? 0u : offsets[i - 1];
  const float4 q0 = __ldg(rec + 3 * (size_t)i);
  int x0, y0, x1, y1;
  get_rect(q0.x, q0.y, r, gx, gy, x0, y0, x1, y1);
  const uint64_t dbits = __float_as_uint(depths[i]);
  for (int y = y0; y < y1; y++)
    for (int x = x0; x < x1; x++) {
      if (off < cap) {
        keys[off] = ((uint64_t)(uint32_t)(y * gx + x) << 32) | dbits;
        vals[off] = (uint32_t)i;
      }
      off++;
    }
}

// Thread k handles Gaussian order[k] (depth rank k).  offsets[] is the inclusive scan of
// tiles_touched in that same order.
__global__ void __launch_bounds__(256)
duplicate_tiles_kernel(int P, const uint32_t* __restrict__ order, const float4* __restrict__ rec,
                       const uint32_t* __restrict__ offsets, const int32_t* __restrict__ radii,
                       int gx, int gy, uint32_t* __restrict__ tile_keys, uint32_t* __restrict__ vals,
                       uint32_t cap, int32_t* __restrict__ status) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= P) return;
  if (k == P - 1 && offsets[k] > cap) atomicExch(status + 1, 1);  // capacity overflow
  const uint32_t i = order[k];
  const int r = radii[i];
  if (r <= 0) return;  // culled Gaussians carry key 0xFFFFFFFF and sit at the end of `order`
  uint32_t off = (k == 0) ? 0u : offsets[k - 1];
  const float4 q0 = __ldg(rec + 3 * (size_t)i);
  int x0, y0, x1, y1;
  get_rect(q0.x, q0.y, r, gx, gy, x0, y0, x1, y1);
  for (int y = y0; y < y1; y++)
    for (int x = x0; x < x1; x++) {
      if (off < cap) {
        tile_keys[off] = (uint32_t)(y * gx + x);
        vals[off] = i;
      }
      off++;
    }
}


// =================================================================================================
// Fused K2 + K3 for the two-level scheme: ONE pass does the inclusive scan of tiles_touched in
// depth order (chained tiles, decoupled look-back), the load-balanced expansion into (tile, id)
// pairs, and the per-tile instance counts.
//   * K1 stores each Gaussian's tile rect as ushort4 {x0, y0, x1, y1} (zero area when culled), so
//     nothing is recomputed here and one 8-byte gather per Gaussian replaces radii + mean2D + rect;
//   * a warp expands 32 Gaussians at a time: output slot j of the round's contiguous span finds its
//     owner by a 5-step shuffle search over the round's inclusive counts, so the key / value stores
//     are fully coalesced however unequal the rects are (a thread-per-Gaussian loop serialises on
//     the largest splat of the warp and writes 32 scattered runs);
//   * tile_count[tile] += 1 per instance (L2 reductions): K5's ranges are the exclusive scan of
//     these counts and the tile sort's digit histograms are their marginals (tile_prepare_kernel),
//     which removes the histogram pass over the keys and the boundary-detection pass of K5.
// Emission order is unchanged: depth rank, then y-major over the rect (Appendix A.3).
// =================================================================================================
constexpr int BE_THREADS = 256;
#ifndef GSR_BE_ROUNDS
#define GSR_BE_ROUNDS 4
#endif
#ifndef GSR_BE_BIG
#define GSR_BE_BIG 64
#endif
constexpr int BE_ROUNDS = GSR_BE_ROUNDS;
constexpr int BE_TILE = BE_THREADS * BE_ROUNDS;  // == SCAN_TILE: the look-back status array is the scan's
constexpr uint32_t BE_BIG = GSR_BE_BIG;                  // rects with more tiles than this go to the work list
constexpr unsigned long long BE_FLAG_AGG = 1ull << 62;
constexpr unsigned long long BE_FLAG_INC = 2ull << 62;
constexpr unsigned long long BE_VAL_MASK = (1ull << 62) - 1;

// floor(q / w) = umulhi(q, magic(w)) whenever q * w < 2^32 (here q < 2^16 and w < 2^16); w == 1 -> q
__device__ __forceinline__ uint32_t div_magic(uint32_t w) { return w > 1 ? (0xFFFFFFFFu / w + 1u) : 0u; }
__device__ __forceinline__ uint32_t tile_of_slot(uint32_t q, uint32_t xy0, uint32_t w, uint32_t magic, uint32_t gx) {
  const uint32_t row = w > 1 ? __umulhi(q, magic) : q;
  return ((xy0 >> 16) + row) * gx + (xy0 & 0xffffu) + (q - row * w);
}

int g_bin_count_mode = 2;   // gsr_debug_set knob 1
struct BinArgs {
  BinView v[GSR_MAX_BATCH];
  int count;   // 0: nothing (timing experiments), 1: tile_count[tile] += 1 per instance, 2: per-CTA digit histograms
};

// Digit histograms of the tile sort, counted where the keys are made (mode 2): the CTA keeps [pass][256] counters in shared
// memory and adds its non-zero bins to the sort's histogram at the end -- ~500 L2 reductions per CTA instead of one per
// instance (3 300 per CTA at the headline; the per-instance tile_count atomics were 0.030 of the expansion's 0.079 ms:
// the L2's atomic rate, not contention).  Consecutive lanes hold consecutive tiles of one rect row: distinct low digits
// (no bank conflict), one shared high digit -- added once per warp when the vote says so.  Convergent call: every lane of
// the warp, `ok` = the lane holds an instance.
__device__ __forceinline__ void count_digits(uint32_t* s_h, bool ok, uint32_t t, int passes, int lane) {
  const unsigned act = __ballot_sync(0xffffffffu, ok);
  if (act == 0) return;
  if (ok) atomicAdd(&s_h[t & 0xffu], 1u);
  const int lead = __ffs(act) - 1;
  for (int p = 1; p < passes; p++) {
    const uint32_t d = (t >> (8 * p)) & 0xffu;
    const uint32_t d0 = __shfl_sync(0xffffffffu, d, lead);
    if (__all_sync(0xffffffffu, !ok || d == d0)) {
      if (lane == lead) atomicAdd(&s_h[p * 256 + d0], (uint32_t)__popc(act));
    } else if (ok) {
      atomicAdd(&s_h[p * 256 + d], 1u);
    }
  }
}
__device__ __forceinline__ void flush_digits(const uint32_t* s_h, uint32_t* __restrict__ hist, int passes) {
  __syncthreads();
  for (int i = threadIdx.x; i < passes * 256; i += blockDim.x)
    if (s_h[i]) atomicAdd(hist + i, s_h[i]);
}

// blockIdx.y = view.  Walks the V = status[5] visible Gaussians of `order` (depth order, culled ones were dropped by
// the depth sort's compacting first pass); the grid is sized for P and surplus CTAs retire before taking a ticket.
__global__ void __launch_bounds__(BE_THREADS)
bin_expand_kernel(const __grid_constant__ BinArgs args) {
  const BinView& a = args.v[blockIdx.y];
  const int P = (int)min((uint32_t)a.status[5], (uint32_t)a.P);
  if ((int64_t)blockIdx.x * BE_TILE >= P) return;
  const uint32_t* __restrict__ order = a.order;
  const ushort4* __restrict__ rects = a.rects;
  uint32_t* __restrict__ offsets = a.offsets;
  uint32_t* __restrict__ tile_keys = a.tile_keys;
  uint32_t* __restrict__ vals = a.vals;
  const uint32_t cap = a.cap;
  uint32_t* __restrict__ tile_count = a.tile_count;
  const int gx = a.gx;
  volatile unsigned long long* lb_status = a.lb_status;
  int32_t* __restrict__ status = a.status;
  uint4* __restrict__ big_items = a.big_items;
  const uint32_t big_cap = a.big_cap;
  __shared__ uint32_t s_tile;
  __shared__ uint32_t s_warp[BE_THREADS / 32];
  __shared__ uint32_t s_prefix;
  __shared__ uint32_t s_h[4 * 256];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int passes = (a.end_bit + 7) >> 3;   // <= 4 (launch_bin_expand checks)
  if (args.count == 2)
    for (int k = tid; k < passes * 256; k += BE_THREADS) s_h[k] = 0;
  if (tid == 0) s_tile = atomicAdd(a.ticket, 1u);
  __syncthreads();
  const uint32_t tile = s_tile;
  const int64_t kbase = (int64_t)tile * BE_TILE + warp * (32 * BE_ROUNDS) + lane;

  uint32_t id[BE_ROUNDS], xy0[BE_ROUNDS], wh[BE_ROUNDS], cnt[BE_ROUNDS], incl[BE_ROUNDS];
#pragma unroll
  for (int r = 0; r < BE_ROUNDS; r++) {
    const int64_t k = kbase + r * 32;
    id[r] = k < P ? __ldg(order + k) : 0u;
  }
#pragma unroll
  for (int r = 0; r < BE_ROUNDS; r++) {
    const int64_t k = kbase + r * 32;
    ushort4 q = make_ushort4(0, 0, 0, 0);
    if (k < P) q = __ldg(rects + id[r]);
    xy0[r] = (uint32_t)q.x | ((uint32_t)q.y << 16);
    const uint32_t w = (uint32_t)(q.z - q.x), h = (uint32_t)(q.w - q.y);
    wh[r] = w | (h << 16);
    cnt[r] = w * h;
  }
  // per-round inclusive scans over the warp, then the warp's running total
  uint32_t warp_total = 0;
#pragma unroll
  for (int r = 0; r < BE_ROUNDS; r++) {
    uint32_t x = cnt[r];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
      if (lane >= o) x += y;
    }
    incl[r] = x;
    warp_total += __shfl_sync(0xffffffffu, x, 31);
  }
  if (lane == 0) s_warp[warp] = warp_total;
  __syncthreads();
  if (warp == 0) {
    uint32_t wv = lane < BE_THREADS / 32 ? s_warp[lane] : 0;
#pragma unroll
    for (int o = 1; o < BE_THREADS / 32; o <<= 1) {
      const uint32_t y = __shfl_up_sync(0xffffffffu, wv, o);
      if (lane >= o) wv += y;
    }
    const uint32_t block_total = __shfl_sync(0xffffffffu, wv, BE_THREADS / 32 - 1);
    if (lane < BE_THREADS / 32) s_warp[lane] = wv;  // inclusive over warps
    unsigned long long excl = 0;
    if (tile == 0) {
      if (lane == 0) lb_status[0] = BE_FLAG_INC | block_total;
    } else {
      if (lane == 0) lb_status[tile] = BE_FLAG_AGG | block_total;
      int64_t look = (int64_t)tile - 1;
      while (true) {  // decoupled look-back, 32 predecessors per step
        const int64_t t = look - lane;
        unsigned long long st = BE_FLAG_INC;
        if (t >= 0) {
          do { st = lb_status[t]; } while ((st >> 62) == 0);
        }
        const unsigned inc_mask = __ballot_sync(0xffffffffu, (st & BE_FLAG_INC) != 0);
        const int first_inc = inc_mask ? (__ffs(inc_mask) - 1) : 32;
        unsigned long long val = (lane <= first_inc) ? (st & BE_VAL_MASK) : 0ull;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) val += __shfl_xor_sync(0xffffffffu, val, o);
        excl += val;
        if (inc_mask) break;
        look -= 32;
      }
      if (lane == 0) lb_status[tile] = BE_FLAG_INC | ((excl + block_total) & BE_VAL_MASK);
    }
    if (lane == 0) {
      s_prefix = (uint32_t)excl;
      // the tile holding the last Gaussian knows N: flag a capacity overflow
      if ((int64_t)tile * BE_TILE + BE_TILE >= P && (uint32_t)excl + block_total > cap) atomicExch(status + 1, 1);
    }
  }
  __syncthreads();
  uint32_t base = s_prefix + (warp > 0 ? s_warp[warp - 1] : 0);  // exclusive offset of this warp's round 0

#pragma unroll
  for (int r = 0; r < BE_ROUNDS; r++) {
    const int64_t k = kbase + r * 32;
    if (k < P) offsets[k] = base + incl[r];
    const uint32_t total = __shfl_sync(0xffffffffu, incl[r], 31);
    const uint32_t start = base + incl[r] - cnt[r];  // this Gaussian's first output slot
    // Large rects would serialise the warp (depth order puts the nearest = largest splats into the
    // same few CTAs): they are queued and expanded by bin_expand_big_kernel, one warp per item.
    const bool big = cnt[r] > BE_BIG;
    // Only rects whose first slot lies inside the capacity are queued: at most cap / (BE_BIG + 1) + 1 = big_cap of
    // them exist, so every slot below `cap` is written by somebody.  (When the speculative capacity is too small --
    // N > cap -- the rects beyond it used to compete for the work list's slots in arrival order and could push out
    // rects that start below cap: their key / value slots then kept whatever the recycled buffer held, and the blend
    // of the discarded speculative pass dereferenced those stale ids.)
    if (big && start < cap) {
      const uint32_t slot = atomicAdd(reinterpret_cast<uint32_t*>(status + 4), 1u);
      if (slot < big_cap) big_items[slot] = make_uint4(id[r], xy0[r], wh[r], start);
    }
    // inclusive scan of the counts this warp expands itself
    const uint32_t c_small = big ? 0u : cnt[r];
    uint32_t incl_s = c_small;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t y = __shfl_up_sync(0xffffffffu, incl_s, o);
      if (lane >= o) incl_s += y;
    }
    const uint32_t total_s = __shfl_sync(0xffffffffu, incl_s, 31);
    const uint32_t excl_s = incl_s - c_small;
    const uint32_t w = wh[r] & 0xffffu;
    const uint32_t magic = div_magic(w);
    for (uint32_t jb = 0; jb < total_s; jb += 32) {  // warp-uniform trip count
      const uint32_t j = jb + lane;
      int lo = 0;
#pragma unroll
      for (int step = 16; step > 0; step >>= 1) {
        const uint32_t v = __shfl_sync(0xffffffffu, incl_s, lo + step - 1);
        if (v <= j) lo += step;
      }
      lo = min(lo, 31);  // lanes past the end of the span (j >= total_s) are masked below
      const uint32_t o_excl = __shfl_sync(0xffffffffu, excl_s, lo);
      const uint32_t o_start = __shfl_sync(0xffffffffu, start, lo);
      const uint32_t o_xy = __shfl_sync(0xffffffffu, xy0[r], lo);
      const uint32_t o_w = __shfl_sync(0xffffffffu, w, lo);
      const uint32_t o_magic = __shfl_sync(0xffffffffu, magic, lo);
      const uint32_t o_id = __shfl_sync(0xffffffffu, id[r], lo);
      const uint32_t q = j - o_excl;
      const uint32_t off = o_start + q;
      const bool ok = j < total_s && off < cap;
      uint32_t t = 0;
      if (ok) {
        t = tile_of_slot(q, o_xy, o_w, o_magic, (uint32_t)gx);
        tile_keys[off] = t;
        vals[off] = o_id;
        if (args.count == 1) atomicAdd(tile_count + t, 1u);
      }
      if (args.count == 2) count_digits(s_h, ok, t, passes, lane);
    }
    base += total;
  }
  if (args.count == 2) flush_digits(s_h, a.hist, passes);
}

// Work-list items {id, x0 | y0 << 16, w | h << 16, first output slot}: one warp per item, coalesced.
__global__ void __launch_bounds__(256)
bin_expand_big_kernel(const __grid_constant__ BinArgs args) {
  const BinView& a = args.v[blockIdx.y];
  const uint4* __restrict__ big_items = a.big_items;
  uint32_t* __restrict__ tile_keys = a.tile_keys;
  uint32_t* __restrict__ vals = a.vals;
  uint32_t* __restrict__ tile_count = a.tile_count;
  const uint32_t cap = a.cap;
  const int gx = a.gx;
  const uint32_t n_items = min((uint32_t)a.status[4], a.big_cap);
  const uint32_t lane = threadIdx.x & 31;
  const uint32_t n_warps = gridDim.x * (blockDim.x >> 5);
  __shared__ uint32_t s_h[4 * 256];
  const int passes = (a.end_bit + 7) >> 3;
  if (args.count == 2) {
    if (blockIdx.x * (blockDim.x >> 5) >= n_items) return;   // no item for any warp of this CTA
    for (int k = threadIdx.x; k < passes * 256; k += blockDim.x) s_h[k] = 0;
    __syncthreads();
  }
  for (uint32_t it = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); it < n_items; it += n_warps) {
    const uint4 item = __ldg(big_items + it);
    const uint32_t w = item.z & 0xffffu, c = w * (item.z >> 16);
    const uint32_t magic = div_magic(w);
    for (uint32_t qb = 0; qb < c; qb += 32) {   // warp-uniform trip count
      const uint32_t q = qb + lane;
      const uint32_t off = item.w + q;
      const bool ok = q < c && off < cap;
      uint32_t t = 0;
      if (ok) {
        t = tile_of_slot(q, item.y, w, magic, (uint32_t)gx);
        tile_keys[off] = t;
        vals[off] = item.x;
        if (args.count == 1) atomicAdd(tile_count + t, 1u);
      }
      if (args.count == 2) count_digits(s_h, ok, t, passes, (int)lane);
    }
  }
  if (args.count == 2) flush_digits(s_h, a.hist, passes);
}

// Mode 2: the tile ranges are searched in the SORTED keys -- one warp per tile, 32-ary lower bound (five rounds of 32
// probes at N = 2^25), so the cost is G x log32(N) probes whatever N is (a boundary-detection pass over the keys cost
// 13 us per view at the headline and 0.75 ms at the 4K stress shape).  Every tile's range is written: (0, 0) when empty,
// as the reference's memset leaves untouched tiles.  blockIdx.y = view.
struct RangeArgs {
  RangeView v[GSR_MAX_BATCH];
};
__device__ __forceinline__ int64_t warp_lower_bound(const uint32_t* __restrict__ keys, int64_t lo, int64_t hi, uint32_t t, int lane) {
  while (hi > lo) {   // the answer lies in [lo, hi]; warp-uniform
    const int64_t step = (hi - lo + 31) >> 5;
    const int64_t pos = min(lo + (int64_t)(lane + 1) * step - 1, hi - 1);   // ascending probes, the last one at hi - 1
    const unsigned m = __ballot_sync(0xffffffffu, __ldg(keys + pos) >= t);
    if (m == 0) return hi;
    const int f = __ffs(m) - 1;                                             // first probe with key >= t
    const int64_t pf = min(lo + (int64_t)(f + 1) * step - 1, hi - 1);
    const int64_t pprev = f == 0 ? lo - 1 : min(lo + (int64_t)f * step - 1, hi - 1);   // last probe with key < t
    lo = pprev + 1;
    hi = pf;
  }
  return lo;
}
__global__ void __launch_bounds__(256) tile_ranges_views_kernel(const __grid_constant__ RangeArgs args) {
  const RangeView& a = args.v[blockIdx.y];
  const int t = (int)(blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5));
  if (t >= a.G) return;   // warp-uniform
  const int lane = threadIdx.x & 31;
  const int64_t N = min((int64_t)*a.n_dev, a.cap);
  const int64_t first = warp_lower_bound(a.keys, 0, N, (uint32_t)t, lane);
  const int64_t last = warp_lower_bound(a.keys, first, N, (uint32_t)t + 1u, lane);
  if (lane == 0) a.ranges[t] = last > first ? make_uint2((uint32_t)first, (uint32_t)last) : make_uint2(0u, 0u);
}

// One CTA: ranges = exclusive scan of the per-tile counts ((0,0) for untouched tiles, as the
// reference's memset leaves them), and the tile sort's per-digit exclusive bases (what the onesweep
// histogram + scan kernels would produce from the keys).
struct PrepArgs {
  PrepView v[GSR_MAX_BATCH];
};
__global__ void __launch_bounds__(1024)
tile_prepare_kernel(const __grid_constant__ PrepArgs args) {
  const PrepView& a = args.v[blockIdx.x];
  const int G = a.G, passes = a.passes, end_bit = a.end_bit;
  const uint32_t* __restrict__ tile_count = a.tile_count;
  uint2* __restrict__ ranges = a.ranges;
  uint32_t* __restrict__ hist = a.hist;
  __shared__ uint32_t s_h[4 * 256];
  __shared__ uint32_t s_w[32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int i = tid; i < passes * 256; i += 1024) s_h[i] = 0;
  __syncthreads();
  const int per = (G + 1023) / 1024;
  const int t0 = tid * per, t1 = min(G, t0 + per);
  uint32_t sum = 0;
  for (int t = t0; t < t1; t++) {
    const uint32_t c = tile_count[t];
    sum += c;
    if (c)
      for (int p = 0; p < passes; p++) {
        const int bits = min(8, end_bit - 8 * p);
        atomicAdd(&s_h[p * 256 + (((uint32_t)t >> (8 * p)) & ((1u << bits) - 1))], c);
      }
  }
  uint32_t x = sum;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
    if (lane >= o) x += y;
  }
  if (lane == 31) s_w[warp] = x;
  __syncthreads();
  if (warp == 0) {
    uint32_t v = s_w[lane];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t y = __shfl_up_sync(0xffffffffu, v, o);
      if (lane >= o) v += y;
    }
    s_w[lane] = v;
  }
  __syncthreads();
  uint32_t run = (warp > 0 ? s_w[warp - 1] : 0) + x - sum;
  for (int t = t0; t < t1; t++) {
    const uint32_t c = tile_count[t];
    ranges[t] = c ? make_uint2(run, run + c) : make_uint2(0u, 0u);
    run += c;
  }
  // exclusive scan of each digit histogram: warp p handles pass p (8 bins per lane)
  if (warp < passes) {
    uint32_t v[8], tot = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) { v[k] = s_h[warp * 256 + lane * 8 + k]; tot += v[k]; }
    uint32_t inc = tot;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t y = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += y;
    }
    uint32_t e = inc - tot;
#pragma unroll
    for (int k = 0; k < 8; k++) { hist[warp * 256 + lane * 8 + k] = e; e += v[k]; }
  }
}

template <typename KeyT>
__device__ __forceinline__ uint32_t tile_of(KeyT k);
template <> __device__ __forceinline__ uint32_t tile_of<uint64_t>(uint64_t k) { return (uint32_t)(k >> 32); }
template <> __device__ __forceinline__ uint32_t tile_of<uint32_t>(uint32_t k) { return k; }

template <typename KeyT>
__global__ void __launch_bounds__(256)
tile_ranges_kernel(int64_t cap, const uint32_t* __restrict__ n_dev, const KeyT* __restrict__ keys,
                   uint2* __restrict__ ranges) {
  const int64_t N = n_dev ? min((int64_t)*n_dev, cap) : cap;
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= N) return;
  const uint32_t cur = tile_of<KeyT>(keys[idx]);
  if (idx == 0) {
    ranges[cur].x = 0;
  } else {
    const uint32_t prev = tile_of<KeyT>(keys[idx - 1]);
    if (cur != prev) {
      ranges[prev].y = (uint32_t)idx;
      ranges[cur].x = (uint32_t)idx;
    }
  }
  if (idx == N - 1) ranges[cur].y = (uint32_t)N;
}

cudaError_t launch_duplicate_key64(cudaStream_t s, int P, const float4* rec, const float* depths,
                                   const uint32_t* offsets, const int32_t* radii, int grid_x,
                                   int grid_y, uint64_t* keys, uint32_t* vals, int64_t cap, int32_t* status) {
  if (P == 0) return cudaSuccess;
  duplicate_key64_kernel<<<cdiv(P, 256), 256, 0, s>>>(P, rec, depths, offsets, radii, grid_x, grid_y, keys, vals,
                                                      (uint32_t)cap, status);
  count_launch();
  return cudaGetLastError();
}
cudaError_t launch_duplicate_tiles(cudaStream_t s, int P, const uint32_t* order, const float4* rec,
                                   const uint32_t* offsets, const int32_t* radii, int grid_x,
                                   int grid_y, uint32_t* tile_keys, uint32_t* vals, int64_t cap, int32_t* status) {
  if (P == 0) return cudaSuccess;
  duplicate_tiles_kernel<<<cdiv(P, 256), 256, 0, s>>>(P, order, rec, offsets, radii, grid_x, grid_y, tile_keys, vals,
                                                      (uint32_t)cap, status);
  count_launch();
  return cudaGetLastError();
}
// capacity of the work list for `cap` instances: every queued rect has more than BE_BIG tiles
int64_t bin_big_capacity(int64_t cap) { return cap / (BE_BIG + 1) + 1; }

// The caller has zeroed every view's scan temp (ticket + look-back words), tile_count and status[4].
cudaError_t launch_bin_expand(cudaStream_t s, const BinView* views, int nv, int count_mode) {
  if (nv <= 0) return cudaSuccess;
  if (nv > GSR_MAX_BATCH) return cudaErrorInvalidValue;
  BinArgs args{};
  args.count = count_mode;
  int max_p = 0;
  for (int k = 0; k < nv; k++) {
    args.v[k] = views[k];
    max_p = views[k].P > max_p ? views[k].P : max_p;
    if (views[k].end_bit > 32) return cudaErrorInvalidValue;
  }
  if (max_p == 0) return cudaSuccess;
  const int64_t tiles = ((int64_t)max_p + BE_TILE - 1) / BE_TILE;
  bin_expand_kernel<<<dim3((unsigned)tiles, (unsigned)nv), BE_THREADS, 0, s>>>(args);
  const int big_blocks = nv >= 4 ? 148 : 148 * 4 / nv;
  bin_expand_big_kernel<<<dim3(big_blocks, (unsigned)nv), 256, 0, s>>>(args);
  count_launch(2);
  return cudaGetLastError();
}
cudaError_t launch_tile_prepare(cudaStream_t s, const PrepView* views, int nv) {
  if (nv <= 0) return cudaSuccess;
  if (nv > GSR_MAX_BATCH) return cudaErrorInvalidValue;
  PrepArgs args{};
  for (int k = 0; k < nv; k++) {
    args.v[k] = views[k];
    args.v[k].passes = (views[k].end_bit + 7) / 8;
    if (args.v[k].passes > 4) return cudaErrorInvalidValue;
  }
  tile_prepare_kernel<<<nv, 1024, 0, s>>>(args);
  count_launch();
  return cudaGetLastError();
}

cudaError_t launch_tile_ranges_views(cudaStream_t s, const RangeView* views, int nv) {
  if (nv <= 0) return cudaSuccess;
  if (nv > GSR_MAX_BATCH) return cudaErrorInvalidValue;
  RangeArgs args{};
  int max_g = 0;
  for (int k = 0; k < nv; k++) {
    args.v[k] = views[k];
    max_g = views[k].G > max_g ? views[k].G : max_g;
  }
  if (max_g <= 0) return cudaSuccess;
  tile_ranges_views_kernel<<<dim3((unsigned)cdiv(max_g, 8), (unsigned)nv), 256, 0, s>>>(args);
  count_launch();
  return cudaGetLastError();
}

// ---- one launch that zeroes every counter / look-back array / histogram a batched forward needs --------------------
constexpr int ZERO_MAX_REGIONS = 4 * GSR_MAX_BATCH;
struct ZeroArgs {
  uint32_t* ptr[ZERO_MAX_REGIONS];
  unsigned long long words[ZERO_MAX_REGIONS];
};
__global__ void __launch_bounds__(256) zero_regions_kernel(const __grid_constant__ ZeroArgs args) {
  uint32_t* __restrict__ p = args.ptr[blockIdx.y];
  const unsigned long long words = args.words[blockIdx.y];
  const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
  const unsigned long long t = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
  if ((reinterpret_cast<uintptr_t>(p) & 15) == 0) {
    const unsigned long long quads = words >> 2;
    for (unsigned long long i = t; i < quads; i += stride) reinterpret_cast<uint4*>(p)[i] = make_uint4(0, 0, 0, 0);
    for (unsigned long long i = (quads << 2) + t; i < words; i += stride) p[i] = 0;
  } else {
    for (unsigned long long i = t; i < words; i += stride) p[i] = 0;
  }
}
cudaError_t launch_zero_regions(cudaStream_t s, const ZeroRegion* regions, int n) {
  for (int base = 0; base < n; base += ZERO_MAX_REGIONS) {
    ZeroArgs args{};
    int m = 0;
    unsigned long long max_words = 0;
    for (int k = base; k < n && m < ZERO_MAX_REGIONS; k++) {
      if (!regions[k].ptr || regions[k].bytes == 0) continue;
      args.ptr[m] = reinterpret_cast<uint32_t*>(regions[k].ptr);
      args.words[m] = regions[k].bytes / 4;
      max_words = args.words[m] > max_words ? args.words[m] : max_words;
      m++;
    }
    if (m == 0) continue;
    unsigned long long blocks = (max_words / 4 + 255) / 256;
    if (blocks > 148 * 2) blocks = 148 * 2;
    if (blocks < 1) blocks = 1;
    zero_regions_kernel<<<dim3((unsigned)blocks, (unsigned)m), 256, 0, s>>>(args);
    count_launch();
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
  }
  return cudaSuccess;
}
cudaError_t launch_tile_ranges_u64(cudaStream_t s, int64_t N, const uint32_t* n_dev, const uint64_t* keys, int G, uint2* ranges) {
  cudaError_t e = cudaMemsetAsync(ranges, 0, (size_t)G * sizeof(uint2), s);
  if (e != cudaSuccess || N == 0) return e;
  tile_ranges_kernel<uint64_t><<<cdiv(N, 256), 256, 0, s>>>(N, n_dev, keys, ranges);
  count_launch();
  return cudaGetLastError();
}
cudaError_t launch_tile_ranges_u32(cudaStream_t s, int64_t N, const uint32_t* n_dev, const uint32_t* keys, int G, uint2* ranges) {
  cudaError_t e = cudaMemsetAsync(ranges, 0, (size_t)G * sizeof(uint2), s);
  if (e != cudaSuccess || N == 0) return e;
  tile_ranges_kernel<uint32_t><<<cdiv(N, 256), 256, 0, s>>>(N, n_dev, keys, ranges);
  count_launch();
  return cudaGetLastError();
}

}  // namespace gsr
