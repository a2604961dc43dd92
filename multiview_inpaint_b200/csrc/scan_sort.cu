// scan_sort.cu -- K2 (single-pass inclusive scan) and K4 (onesweep LSD radix sort), hand-written.
//
// Restates SURVEY.md Appendix A.3 (point_offsets = inclusive_sum(tiles_touched)) and A.4 (stable
// radix sort of (key,value) pairs on key bits [0, end_bit)); the reference calls
// cub::DeviceScan::InclusiveSum and cub::DeviceRadixSort::SortPairs for these.
// Both kernels are HBM-bound streaming passes:
//   scan      : 1 read + 1 write of n uint32, chained tiles with decoupled look-back
//   onesweep  : 1 histogram read of the keys, then per 8-bit digit ONE read + ONE write of every
//               pair; tile prefixes travel through a (tile x 256) status array with look-back.
// Tile ids are handed out by an atomic ticket so a tile's predecessors are always resident.
#include "common.cuh"

namespace gsr {

// =================================================================================================
// Inclusive scan (uint32), 256 threads x 4 items (one uint4 per thread), warp-wide look-back
// =================================================================================================
constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 4;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;
constexpr unsigned long long SCAN_FLAG_AGG = 1ull << 62;
constexpr unsigned long long SCAN_FLAG_INC = 2ull << 62;
constexpr unsigned long long SCAN_VAL_MASK = (1ull << 62) - 1;

size_t scan_temp_bytes(int64_t n) {
  const int64_t tiles = (n + SCAN_TILE - 1) / SCAN_TILE;
  return align_up(16) + align_up((size_t)tiles * 8);
}

__global__ void __launch_bounds__(SCAN_THREADS)
scan_kernel(const uint32_t* __restrict__ in, const uint32_t* __restrict__ gather,
            uint32_t* __restrict__ out, int64_t n, volatile unsigned long long* status,
            uint32_t* ticket) {
  __shared__ uint32_t s_tile;
  __shared__ uint32_t s_warp[SCAN_THREADS / 32];
  __shared__ uint32_t s_prefix;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) s_tile = atomicAdd(ticket, 1u);
  __syncthreads();
  const uint32_t tile = s_tile;
  const int64_t base = (int64_t)tile * SCAN_TILE + (int64_t)tid * SCAN_ITEMS;

  uint32_t v[SCAN_ITEMS];
  if (gather == nullptr && base + SCAN_ITEMS <= n && ((reinterpret_cast<uintptr_t>(in) & 15) == 0)) {
    const uint4 q = __ldg(reinterpret_cast<const uint4*>(in + base));
    v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w;
  } else {
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++) {
      const int64_t idx = base + k;
      v[k] = 0;
      if (idx < n) v[k] = gather ? __ldg(in + __ldg(gather + idx)) : __ldg(in + idx);
    }
  }
  v[1] += v[0]; v[2] += v[1]; v[3] += v[2];
  // warp inclusive scan of thread totals
  uint32_t x = v[3];
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
    if (lane >= o) x += y;
  }
  if (lane == 31) s_warp[warp] = x;
  __syncthreads();
  if (warp == 0) {
    uint32_t w = lane < SCAN_THREADS / 32 ? s_warp[lane] : 0;
#pragma unroll
    for (int o = 1; o < SCAN_THREADS / 32; o <<= 1) {
      const uint32_t y = __shfl_up_sync(0xffffffffu, w, o);
      if (lane >= o) w += y;
    }
    const uint32_t block_total = __shfl_sync(0xffffffffu, w, SCAN_THREADS / 32 - 1);
    if (lane < SCAN_THREADS / 32) s_warp[lane] = w;  // inclusive over warps
    // ---- decoupled look-back, 32 predecessors per step ----
    unsigned long long excl = 0;
    if (tile == 0) {
      if (lane == 0) status[0] = SCAN_FLAG_INC | block_total;
    } else {
      if (lane == 0) status[tile] = SCAN_FLAG_AGG | block_total;
      int64_t look = (int64_t)tile - 1;
      while (true) {
        const int64_t t = look - lane;
        unsigned long long st = SCAN_FLAG_INC;  // lanes before tile 0 read as "inclusive 0"
        if (t >= 0) {
          do { st = status[t]; } while ((st >> 62) == 0);
        }
        const unsigned inc_mask = __ballot_sync(0xffffffffu, (st & SCAN_FLAG_INC) != 0);
        const int first_inc = inc_mask ? (__ffs(inc_mask) - 1) : 32;
        unsigned long long val = (lane <= first_inc) ? (st & SCAN_VAL_MASK) : 0ull;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) val += __shfl_xor_sync(0xffffffffu, val, o);
        excl += val;
        if (inc_mask) break;
        look -= 32;
      }
      if (lane == 0) status[tile] = SCAN_FLAG_INC | ((excl + block_total) & SCAN_VAL_MASK);
    }
    if (lane == 0) s_prefix = (uint32_t)excl;
  }
  __syncthreads();
  const uint32_t offset = s_prefix + (warp > 0 ? s_warp[warp - 1] : 0) + (x - v[3]);
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; k++)
    if (base + k < n) out[base + k] = v[k] + offset;
}

cudaError_t launch_inclusive_scan(cudaStream_t s, int64_t n, const uint32_t* in,
                                  const uint32_t* gather, uint32_t* out, char* temp) {
  if (n <= 0) return cudaSuccess;
  const int64_t tiles = (n + SCAN_TILE - 1) / SCAN_TILE;
  uint32_t* ticket = reinterpret_cast<uint32_t*>(temp);
  auto* status = reinterpret_cast<unsigned long long*>(temp + align_up(16));
  cudaError_t e = cudaMemsetAsync(temp, 0, scan_temp_bytes(n), s);
  if (e != cudaSuccess) return e;
  scan_kernel<<<(unsigned)tiles, SCAN_THREADS, 0, s>>>(in, gather, out, n, status, ticket);
  count_launch();
  return cudaGetLastError();
}

// =================================================================================================
// Onesweep radix sort -- segmented: one launch sorts up to GSR_MAX_BATCH independent arrays (the views
// of a multi-view step), blockIdx.y = segment.  Per 8-bit digit ONE read + ONE write of every pair.
//   * COMPACT first pass: elements whose key is 0xFFFFFFFF (culled Gaussians) are dropped while they
//     are read -- they are not counted, ranked or written -- so the remaining passes and everything
//     downstream walk the V visible entries instead of all P (the count V lives on the device);
//   * stable ranking per warp with match_any; the per-item ranks are kept as packed 16-bit pairs
//     (<= 64 registers: four CTAs per SM);
//   * decoupled look-back eight predecessors at a time (eight independent L2 loads in flight per
//     digit thread instead of one dependent load per predecessor).
// =================================================================================================
constexpr int RS_THREADS = 256;
constexpr int RS_WARPS = RS_THREADS / 32;
constexpr int RS_RADIX = 256;
constexpr int RS_MAX_PASSES = 8;
constexpr int RS_LOOKBACK = 8;
// Look-back status word of one (tile, digit): 2 flag bits + the count / inclusive prefix.  32-bit words (30-bit
// prefix) below 2^30 elements; 64-bit words (62-bit prefix) from there up to the 2^32 - 1 elements that uint32
// positions can address -- the reference's own limit (its point_offsets are uint32).
constexpr int64_t RS_WIDE_FROM = 1ll << 30;
template <typename StatusT> struct RsStatus;
template <> struct RsStatus<uint32_t> {
  static constexpr int SHIFT = 30;
  static constexpr uint32_t AGG = 1u << 30, INC = 2u << 30, MASK = (1u << 30) - 1;
};
template <> struct RsStatus<unsigned long long> {
  static constexpr int SHIFT = 62;
  static constexpr unsigned long long AGG = 1ull << 62, INC = 2ull << 62, MASK = (1ull << 62) - 1;
};
static size_t rs_status_bytes(int64_t n) { return n >= RS_WIDE_FROM ? 8 : 4; }

template <typename KeyT> struct RsCfg;
#ifndef GSR_RS_ITEMS32
#define GSR_RS_ITEMS32 16
#endif
template <> struct RsCfg<uint32_t> { static constexpr int ITEMS = GSR_RS_ITEMS32; };
template <> struct RsCfg<uint64_t> { static constexpr int ITEMS = 12; };

static int rs_tile(int key_bytes) {
  return RS_THREADS * (key_bytes == 8 ? RsCfg<uint64_t>::ITEMS : RsCfg<uint32_t>::ITEMS);
}
static int rs_passes(int end_bit) { return (end_bit + 7) / 8; }

// temp layout: [hist: 8*256 u32][tickets: 8 u32 (padded)][status: passes * tiles * 256 words]
size_t sort_temp_bytes(int64_t n, int key_bytes, int end_bit) {
  const int passes = rs_passes(end_bit);
  const int64_t tiles = (n + rs_tile(key_bytes) - 1) / rs_tile(key_bytes);
  return align_up((size_t)RS_MAX_PASSES * RS_RADIX * 4) + align_up(64) +
         align_up((size_t)passes * (size_t)(tiles > 0 ? tiles : 1) * RS_RADIX * rs_status_bytes(n));
}

template <typename KeyT>
struct RsHistArgs {
  const KeyT* keys[GSR_MAX_BATCH];
  const uint32_t* n_dev[GSR_MAX_BATCH];
  int64_t n_cap[GSR_MAX_BATCH];
  uint32_t* hist[GSR_MAX_BATCH];
};

// COMPACT: keys equal to ~0 are not counted (they are dropped by the first onesweep pass).
// The low digits of a key are close to uniform: plain shared-memory atomics (no two lanes of a warp meet often).
// The most significant digit is highly repetitive (exponent byte of the depth, high bits of a tile id): its lanes
// aggregate with match_any first -- cheap there, because MATCH.ANY's cost grows with the number of DISTINCT values in
// the warp (it was the whole cost of this kernel, 316 us per 12 M keys, when every digit went through it).
template <typename KeyT, bool COMPACT>
__global__ void __launch_bounds__(256)
rs_histogram_kernel(const __grid_constant__ RsHistArgs<KeyT> args, int end_bit) {
  __shared__ uint32_t s_h[RS_MAX_PASSES * RS_RADIX];
  const int seg = blockIdx.y;
  const KeyT* __restrict__ keys = args.keys[seg];
  const int64_t n_cap = args.n_cap[seg];
  const int64_t n = args.n_dev[seg] ? min((int64_t)*args.n_dev[seg], n_cap) : n_cap;
  const int passes = (end_bit + 7) / 8;
  for (int i = threadIdx.x; i < passes * RS_RADIX; i += blockDim.x) s_h[i] = 0;
  __syncthreads();
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int lane = threadIdx.x & 31;
  const int top = passes - 1;
  const int top_bits = end_bit - 8 * top;
  // warp-uniform trip count so that the full-mask match below is legal
  for (int64_t b = (int64_t)blockIdx.x * blockDim.x + (threadIdx.x - lane); b < n; b += stride) {
    const int64_t i = b + lane;
    bool valid = i < n;
    const KeyT k = valid ? keys[i] : (KeyT)0;
    if (COMPACT) valid = valid && k != ~(KeyT)0;
    if (valid)
      for (int p = 0; p < top; p++) atomicAdd(&s_h[p * RS_RADIX + ((uint32_t)(k >> (8 * p)) & 0xffu)], 1u);
    // invalid lanes get a private pseudo-digit so they never pair up
    const uint32_t d = valid ? ((uint32_t)(k >> (8 * top)) & ((1u << top_bits) - 1)) : (0x100u | lane);
    const unsigned peers = __match_any_sync(0xffffffffu, d);
    if (valid && lane == __ffs(peers) - 1) atomicAdd(&s_h[top * RS_RADIX + d], __popc(peers));
  }
  __syncthreads();
  uint32_t* __restrict__ hist = args.hist[seg];
  for (int i = threadIdx.x; i < passes * RS_RADIX; i += blockDim.x)
    if (s_h[i]) atomicAdd(&hist[i], s_h[i]);
}

struct RsScanArgs {
  uint32_t* hist[GSR_MAX_BATCH];
};
// exclusive scan of each pass' 256-bin histogram, in place; one block of 256 threads per (pass, segment)
__global__ void __launch_bounds__(RS_RADIX) rs_scan_hist_kernel(const __grid_constant__ RsScanArgs args) {
  __shared__ uint32_t s_w[RS_RADIX / 32];
  uint32_t* h = args.hist[blockIdx.y] + blockIdx.x * RS_RADIX;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t c = h[tid];
  uint32_t x = c;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
    if (lane >= o) x += y;
  }
  if (lane == 31) s_w[warp] = x;
  __syncthreads();
  uint32_t wp = 0;
  for (int w = 0; w < warp; w++) wp += s_w[w];
  h[tid] = wp + x - c;
}

template <typename KeyT>
__device__ __forceinline__ uint32_t rs_digit(KeyT k, int shift, uint32_t mask) {
  return (uint32_t)(k >> shift) & mask;
}

template <typename KeyT>
struct RsPassSeg {
  const KeyT* keys_in;
  const uint32_t* vals_in;       // NULL: the element index is the value
  KeyT* keys_out;
  uint32_t* vals_out;
  const uint32_t* n_dev;         // element count on the device (NULL: n_cap)
  int64_t n_cap;                 // capacity: sizes the grid; surplus CTAs retire before taking a ticket
  const uint32_t* global_base;   // [256] exclusive digit bases of this pass
  void* status;                  // [tiles][256] look-back words (zeroed)
  uint32_t* ticket;              // zeroed
};
template <typename KeyT>
struct RsPassArgs {
  RsPassSeg<KeyT> seg[GSR_MAX_BATCH];
};

// Look-back words are read and written with relaxed GPU-scope accesses (a `volatile` access compiles to a
// system-scope strong one); each word carries its own flag, so no ordering with other memory is needed.
__device__ __forceinline__ uint32_t ld_status(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long ld_status(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_status(uint32_t* p, uint32_t v) {
  asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void st_status(unsigned long long* p, unsigned long long v) {
  asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// RANK: how the lanes of a warp find their equal-digit peers.
//   0  match_any         one instruction, but MATCH.ANY iterates over the DISTINCT values of the warp: ~30 of them with
//                        uniform 8-bit digits -- it was most of this kernel's time
//   1  eight ballots     fixed cost: one VOTE + one select per digit bit (what cub's onesweep does)
//   2  shared atomicOr   each lane ORs its lane bit into a per-(warp, digit) word and reads the word back
// Phases of a tile (4096 pairs of 32-bit keys):
//   1. keys in (warp-striped), 2. EARLY COUNTS: the tile's digit histogram by plain shared-memory atomics, published
//   as the tile's look-back aggregate BEFORE the ranking -- successors find it there instead of spinning while this
//   tile ranks (the spin was 11 % of the kernel's instructions), 3. stable ranking per warp, 4. look-back (eight
//   predecessors per round trip), 5. reorder through shared memory (values are loaded here, not held in registers
//   across the ranking), 6. run-contiguous scatter.  FULL tiles (every item takes part) run without per-item predicates.
template <typename KeyT, int ITEMS, bool COMPACT, int RANK, bool FULL>
__device__ __forceinline__ void rs_rank_items(const KeyT (&key)[ITEMS], uint32_t live, int shift, uint32_t mask, int lane, int warp,
                                              uint32_t* s_warp_cnt, uint32_t* s_match, uint32_t (&rank2)[ITEMS / 2]) {
  const uint32_t lt_mask = (1u << lane) - 1;
#pragma unroll
  for (int r = 0; r < ITEMS; r++) {
    const bool ok = FULL || ((live >> r) & 1u);
    // items that do not take part get a private pseudo-digit: they never pair up and touch no counter
    const uint32_t d = ok ? rs_digit(key[r], shift, mask) : (0x100u | (uint32_t)lane);
    unsigned peers;
    if (RANK == 0) {
      peers = __match_any_sync(0xffffffffu, d);
    } else if (RANK == 1) {
      peers = FULL ? 0xffffffffu : __ballot_sync(0xffffffffu, ok);   // items that take part
#pragma unroll
      for (int bit = 0; bit < 8; bit++) {
        const bool one = (d >> bit) & 1u;
        const unsigned m = __ballot_sync(0xffffffffu, one);
        peers &= one ? m : ~m;
      }
      if (!ok) peers = 1u << lane;
    } else {
      uint32_t* slot = s_match + warp * RS_RADIX + (d & 0xffu);
      if (ok) atomicOr(slot, 1u << lane);
      __syncwarp();
      peers = ok ? *slot : (1u << lane);
      __syncwarp();
      if (ok && lane == __ffs(peers) - 1) *slot = 0;   // ready for the next item (ordered by the __syncwarp below)
    }
    const int leader = __ffs(peers) - 1;
    uint32_t old = 0;
    if (ok && lane == leader) {
      old = s_warp_cnt[warp * RS_RADIX + d];
      s_warp_cnt[warp * RS_RADIX + d] = old + __popc(peers);
    }
    old = __shfl_sync(0xffffffffu, old, leader);
    const uint32_t rk = old + __popc(peers & lt_mask);
    if (r & 1) rank2[r >> 1] |= rk << 16; else rank2[r >> 1] = rk;
    __syncwarp();
  }
}

template <typename KeyT, int ITEMS, typename StatusT, bool COMPACT, int RANK>
#ifndef GSR_RS_MINB
#define GSR_RS_MINB 4
#endif
__global__ void __launch_bounds__(RS_THREADS, sizeof(KeyT) == 4 ? GSR_RS_MINB : 2)
rs_onesweep_kernel(const __grid_constant__ RsPassArgs<KeyT> args, int shift, int nbits) {
  using ST = RsStatus<StatusT>;
  constexpr int TILE = RS_THREADS * ITEMS;
  static_assert(ITEMS % 2 == 0 && TILE < 65535, "packed 16-bit ranks");
  const RsPassSeg<KeyT>& sg = args.seg[blockIdx.y];
  // the element count may live on the device (no host round trip): grid is sized by capacity and
  // surplus CTAs retire before taking a ticket
  const int64_t n = sg.n_dev ? min((int64_t)*sg.n_dev, sg.n_cap) : sg.n_cap;
  if ((int64_t)blockIdx.x * TILE >= n) return;
  const KeyT* __restrict__ keys_in = sg.keys_in;
  const uint32_t* __restrict__ vals_in = sg.vals_in;
  StatusT* status = reinterpret_cast<StatusT*>(sg.status);
  extern __shared__ __align__(16) unsigned char rs_smem[];
  KeyT* s_keys = reinterpret_cast<KeyT*>(rs_smem);                                  // [TILE]
  uint32_t* s_vals = reinterpret_cast<uint32_t*>(rs_smem + sizeof(KeyT) * TILE);    // [TILE]
  uint32_t* s_warp_cnt = s_vals + TILE;                                             // [WARPS][256]
  uint32_t* s_digit_off = s_warp_cnt + RS_WARPS * RS_RADIX;                         // [256]
  uint32_t* s_global_off = s_digit_off + RS_RADIX;                                  // [256]
  uint32_t* s_misc = s_global_off + RS_RADIX;                                       // [16]
  uint32_t* s_match = s_misc + 16;                                                  // [WARPS][256], RANK == 2 only

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t mask = (1u << nbits) - 1;
  if (tid == 0) s_misc[0] = atomicAdd(sg.ticket, 1u);
  for (int i = tid; i < RS_WARPS * RS_RADIX; i += RS_THREADS) {
    s_warp_cnt[i] = 0;
    if (RANK == 2) s_match[i] = 0;
  }
  s_digit_off[tid] = 0;   // early counts accumulate here (RS_THREADS == RS_RADIX)
  __syncthreads();
  const uint32_t tile = s_misc[0];
  const int64_t base = (int64_t)tile * TILE;
  const int64_t wbase = base + (int64_t)warp * 32 * ITEMS + lane;  // warp-striped arrangement
  const bool full = !COMPACT && base + TILE <= n;                   // CTA-uniform

  KeyT key[ITEMS];
  uint32_t live = 0;  // bit r: item r takes part (inside the array and, in a COMPACT pass, not a culled sentinel)
  if (full) {
#pragma unroll
    for (int r = 0; r < ITEMS; r++) key[r] = keys_in[wbase + r * 32];
    live = (1u << ITEMS) - 1;
  } else {
#pragma unroll
    for (int r = 0; r < ITEMS; r++) {
      const int64_t idx = wbase + r * 32;
      key[r] = ~(KeyT)0;
      if (idx < n) {
        key[r] = keys_in[idx];
        if (!COMPACT || key[r] != ~(KeyT)0) live |= 1u << r;
      }
    }
  }

  // ---- early counts: the tile's digit histogram, published as the look-back aggregate before the ranking ----
#pragma unroll
  for (int r = 0; r < ITEMS; r++)
    if ((live >> r) & 1u) atomicAdd(&s_digit_off[rs_digit(key[r], shift, mask)], 1u);
  __syncthreads();
  const uint32_t count = s_digit_off[tid];   // thread d owns digit d
  StatusT* my = status + (size_t)tile * RS_RADIX + tid;
  st_status(my, (tile == 0 ? ST::INC : ST::AGG) | (StatusT)count);

  // ---- stable ranking: per-warp digit counters + warp multi-split ----
  uint32_t rank2[ITEMS / 2];  // two 16-bit ranks per register
  if (full) rs_rank_items<KeyT, ITEMS, COMPACT, RANK, true>(key, live, shift, mask, lane, warp, s_warp_cnt, s_match, rank2);
  else rs_rank_items<KeyT, ITEMS, COMPACT, RANK, false>(key, live, shift, mask, lane, warp, s_warp_cnt, s_match, rank2);
  __syncthreads();

  // ---- thread d owns digit d: warp offsets, exclusive digit offsets inside the tile, look-back ----
  {
    const int d = tid;
    uint32_t run = 0;
#pragma unroll
    for (int w = 0; w < RS_WARPS; w++) {
      const uint32_t c = s_warp_cnt[w * RS_RADIX + d];
      s_warp_cnt[w * RS_RADIX + d] = run;
      run += c;
    }
    // exclusive scan of `count` over the 256 digits of this tile
    uint32_t x = count;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
      if (lane >= o) x += y;
    }
    if (lane == 31) s_misc[1 + warp] = x;
    __syncthreads();
    uint32_t wp = 0;
    for (int w = 0; w < warp; w++) wp += s_misc[1 + w];
    const uint32_t digit_off = wp + x - count;
    s_digit_off[d] = digit_off;
    if (d == RS_RADIX - 1) s_misc[10] = digit_off + count;  // elements of this tile that take part

    StatusT excl = 0;
    if (tile != 0) {
      int64_t t = (int64_t)tile - 1;
      bool found = false;
      while (!found) {
        StatusT st[RS_LOOKBACK];
#pragma unroll
        for (int j = 0; j < RS_LOOKBACK; j++)
          st[j] = (t - j >= 0) ? ld_status(status + (size_t)(t - j) * RS_RADIX + d) : ST::INC;  // before tile 0: "inclusive 0"
#pragma unroll
        for (int j = 0; j < RS_LOOKBACK; j++) {
          if (!found) {
            StatusT sv = st[j];
            while ((sv >> ST::SHIFT) == 0) sv = ld_status(status + (size_t)(t - j) * RS_RADIX + d);
            excl += sv & ST::MASK;
            found = (sv & ST::INC) != 0;
          }
        }
        t -= RS_LOOKBACK;
      }
      st_status(my, ST::INC | ((excl + (StatusT)count) & ST::MASK));
    }
    s_global_off[d] = sg.global_base[d] + (uint32_t)excl - digit_off;   // positions are < 2^32: mod-2^32 arithmetic
  }
  __syncthreads();

  // ---- reorder through shared memory so that the global scatter is run-wise contiguous ----
  if (full) {
    uint32_t val[ITEMS];
#pragma unroll
    for (int r = 0; r < ITEMS; r++) val[r] = vals_in ? vals_in[wbase + r * 32] : (uint32_t)(wbase + r * 32);
#pragma unroll
    for (int r = 0; r < ITEMS; r++) {
      const uint32_t d = rs_digit(key[r], shift, mask);
      const uint32_t rk = (r & 1) ? (rank2[r >> 1] >> 16) : (rank2[r >> 1] & 0xffffu);
      const uint32_t pos = s_digit_off[d] + s_warp_cnt[warp * RS_RADIX + d] + rk;
      s_keys[pos] = key[r];
      s_vals[pos] = val[r];
    }
  } else {
#pragma unroll
    for (int r = 0; r < ITEMS; r++) {
      if ((live >> r) & 1u) {
        const int64_t idx = wbase + r * 32;
        const uint32_t d = rs_digit(key[r], shift, mask);
        const uint32_t rk = (r & 1) ? (rank2[r >> 1] >> 16) : (rank2[r >> 1] & 0xffffu);
        const uint32_t pos = s_digit_off[d] + s_warp_cnt[warp * RS_RADIX + d] + rk;
        s_keys[pos] = key[r];
        s_vals[pos] = vals_in ? vals_in[idx] : (uint32_t)idx;
      }
    }
  }
  __syncthreads();
  const int n_valid = (int)s_misc[10];
  KeyT* __restrict__ keys_out = sg.keys_out;
  uint32_t* __restrict__ vals_out = sg.vals_out;
#pragma unroll 4
  for (int i = tid; i < n_valid; i += RS_THREADS) {
    const KeyT k = s_keys[i];
    const uint32_t d = rs_digit(k, shift, mask);
    const uint32_t g = s_global_off[d] + (uint32_t)i;  // mod-2^32 arithmetic (offset may have wrapped)
    keys_out[g] = k;
    vals_out[g] = s_vals[i];
  }
}

int g_rs_rank_mode = 2;   // see rs_onesweep_kernel (measured: 2 < 1 < 0 in time); gsr_debug_set(0, mode) switches it for experiments

template <typename KeyT, typename StatusT, bool COMPACT, int RANK>
static cudaError_t rs_launch_pass_r(cudaStream_t s, const RsPassArgs<KeyT>& pa, int nseg, int64_t max_tiles, int shift, int bits) {
  constexpr int ITEMS = RsCfg<KeyT>::ITEMS;
  constexpr int TILE = RS_THREADS * ITEMS;
  const size_t smem = sizeof(KeyT) * TILE + 4 * TILE + 4 * (RS_WARPS * RS_RADIX + 2 * RS_RADIX + 16) +
                      (RANK == 2 ? 4 * RS_WARPS * RS_RADIX : 0);
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(rs_onesweep_kernel<KeyT, ITEMS, StatusT, COMPACT, RANK>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  rs_onesweep_kernel<KeyT, ITEMS, StatusT, COMPACT, RANK><<<dim3((unsigned)max_tiles, (unsigned)nseg), RS_THREADS, smem, s>>>(pa, shift, bits);
  return cudaGetLastError();
}
template <typename KeyT, typename StatusT, bool COMPACT>
static cudaError_t rs_launch_pass(cudaStream_t s, const RsPassArgs<KeyT>& pa, int nseg, int64_t max_tiles, int shift, int bits) {
  switch (g_rs_rank_mode) {
    case 0: return rs_launch_pass_r<KeyT, StatusT, COMPACT, 0>(s, pa, nseg, max_tiles, shift, bits);
    case 1: return rs_launch_pass_r<KeyT, StatusT, COMPACT, 1>(s, pa, nseg, max_tiles, shift, bits);
    default: return rs_launch_pass_r<KeyT, StatusT, COMPACT, 2>(s, pa, nseg, max_tiles, shift, bits);
  }
}

// Sorts `nseg` independent arrays with one launch per pass.  In a `compact` sort the first pass drops keys equal
// to ~0 and the later passes read their element count from seg.n_dev_compact (the number of surviving keys).
template <typename KeyT>
static cudaError_t sort_pairs_batched(cudaStream_t s, const SortSeg<KeyT>* segs, int nseg, int end_bit,
                                      int have_bases, bool compact, bool temp_zeroed = false) {
  if (nseg <= 0) return cudaSuccess;
  if (nseg > GSR_MAX_BATCH) return cudaErrorInvalidValue;
  constexpr int ITEMS = RsCfg<KeyT>::ITEMS;
  constexpr int TILE = RS_THREADS * ITEMS;
  const int passes = rs_passes(end_bit);
  if (passes < 1 || passes > RS_MAX_PASSES) return cudaErrorInvalidValue;
  int64_t max_n = 0;
  for (int k = 0; k < nseg; k++) max_n = segs[k].n > max_n ? segs[k].n : max_n;
  if (max_n <= 0) return cudaSuccess;
  const bool wide = max_n >= RS_WIDE_FROM;
  for (int k = 0; k < nseg && nseg > 1; k++)
    if ((segs[k].n >= RS_WIDE_FROM) != wide) {  // mixed look-back word widths: sort the arrays one by one
      for (int j = 0; j < nseg; j++) {
        cudaError_t ej = sort_pairs_batched<KeyT>(s, segs + j, 1, end_bit, have_bases, compact, temp_zeroed);
        if (ej != cudaSuccess) return ej;
      }
      return cudaSuccess;
    }
  const size_t hist_bytes = align_up((size_t)RS_MAX_PASSES * RS_RADIX * 4);
  cudaError_t e;
  RsHistArgs<KeyT> ha{};
  RsScanArgs sa{};
  for (int k = 0; k < nseg; k++) {
    const SortSeg<KeyT>& g = segs[k];
    const size_t total = sort_temp_bytes(g.n > 0 ? g.n : 1, sizeof(KeyT), end_bit);
    // the caller has filled `hist` with exclusive digit bases when have_bases: clear tickets + status only
    if (!temp_zeroed) {
      e = have_bases ? cudaMemsetAsync(g.temp + hist_bytes, 0, total - hist_bytes, s) : cudaMemsetAsync(g.temp, 0, total, s);
      if (e != cudaSuccess) return e;
    }
    ha.keys[k] = g.keys_in;
    ha.n_dev[k] = g.n_dev;
    ha.n_cap[k] = g.n;
    ha.hist[k] = reinterpret_cast<uint32_t*>(g.temp);
    sa.hist[k] = ha.hist[k];
  }
  if (!have_bases) {
    int hist_blocks = (int)((max_n + 256 * 8 - 1) / (256 * 8));
    const int cap_blocks = 148 * 8 / nseg > 148 ? 148 * 8 / nseg : 148;   // ~8 resident CTAs per SM over all segments
    if (hist_blocks > cap_blocks) hist_blocks = cap_blocks;
    if (compact)
      rs_histogram_kernel<KeyT, true><<<dim3(hist_blocks, nseg), 256, 0, s>>>(ha, end_bit);
    else
      rs_histogram_kernel<KeyT, false><<<dim3(hist_blocks, nseg), 256, 0, s>>>(ha, end_bit);
    rs_scan_hist_kernel<<<dim3(passes, nseg), RS_RADIX, 0, s>>>(sa);
    count_launch(2);
  } else if (have_bases == 2) {   // the producer of the keys counted the digits (bin_expand_kernel): scan only
    rs_scan_hist_kernel<<<dim3(passes, nseg), RS_RADIX, 0, s>>>(sa);
    count_launch(1);
  }
  count_launch(passes);

  // ping-pong so that the LAST pass lands in (keys_out, vals_out); inputs are never written
  for (int p = 0; p < passes; p++) {
    const bool to_out = ((passes - 1 - p) % 2) == 0;
    const bool from_out = !to_out;  // for p > 0 the previous pass wrote the other pair
    RsPassArgs<KeyT> pa{};
    int64_t max_tiles = 1;
    for (int k = 0; k < nseg; k++) {
      const SortSeg<KeyT>& g = segs[k];
      const int64_t tiles = ((g.n > 0 ? g.n : 1) + TILE - 1) / TILE;
      max_tiles = tiles > max_tiles ? tiles : max_tiles;
      RsPassSeg<KeyT>& q = pa.seg[k];
      q.keys_in = p == 0 ? g.keys_in : (from_out ? g.keys_out : g.keys_alt);
      q.vals_in = p == 0 ? g.vals_in : (from_out ? g.vals_out : g.vals_alt);
      q.keys_out = to_out ? g.keys_out : g.keys_alt;
      q.vals_out = to_out ? g.vals_out : g.vals_alt;
      q.n_dev = (compact && p > 0) ? g.n_dev_compact : g.n_dev;
      q.n_cap = g.n;
      q.global_base = reinterpret_cast<uint32_t*>(g.temp) + p * RS_RADIX;
      uint32_t* tickets = reinterpret_cast<uint32_t*>(g.temp + hist_bytes);
      char* status = g.temp + hist_bytes + align_up(64);
      q.ticket = tickets + p;
      q.status = status + (size_t)p * (size_t)tiles * RS_RADIX * rs_status_bytes(g.n);
      if (g.n <= 0) q.n_cap = 0;
    }
    const int bits = (end_bit - 8 * p) < 8 ? (end_bit - 8 * p) : 8;
    const bool cpass = compact && p == 0;
    if (wide)
      e = cpass ? rs_launch_pass<KeyT, unsigned long long, true>(s, pa, nseg, max_tiles, 8 * p, bits)
                : rs_launch_pass<KeyT, unsigned long long, false>(s, pa, nseg, max_tiles, 8 * p, bits);
    else
      e = cpass ? rs_launch_pass<KeyT, uint32_t, true>(s, pa, nseg, max_tiles, 8 * p, bits)
                : rs_launch_pass<KeyT, uint32_t, false>(s, pa, nseg, max_tiles, 8 * p, bits);
    if (e != cudaSuccess) return e;
  }
  return cudaSuccess;
}

cudaError_t launch_sort_pairs_u32_batched(cudaStream_t s, const SortSeg<uint32_t>* segs, int nseg, int end_bit,
                                          int have_bases, bool compact, bool temp_zeroed) {
  return sort_pairs_batched<uint32_t>(s, segs, nseg, end_bit, have_bases, compact, temp_zeroed);
}

cudaError_t launch_sort_pairs_u32(cudaStream_t s, int64_t n, const uint32_t* n_dev, const uint32_t* keys_in,
                                  const uint32_t* vals_in, uint32_t* keys_out, uint32_t* vals_out,
                                  uint32_t* keys_alt, uint32_t* vals_alt, int end_bit, char* temp, bool have_bases) {
  if (n <= 0) return cudaSuccess;
  SortSeg<uint32_t> g{n, n_dev, nullptr, keys_in, vals_in, keys_out, vals_out, keys_alt, vals_alt, temp};
  return sort_pairs_batched<uint32_t>(s, &g, 1, end_bit, have_bases ? 1 : 0, false);
}
cudaError_t launch_sort_pairs_u64(cudaStream_t s, int64_t n, const uint32_t* n_dev, const uint64_t* keys_in,
                                  const uint32_t* vals_in, uint64_t* keys_out, uint32_t* vals_out,
                                  uint64_t* keys_alt, uint32_t* vals_alt, int end_bit, char* temp) {
  if (n <= 0) return cudaSuccess;
  SortSeg<uint64_t> g{n, n_dev, nullptr, keys_in, vals_in, keys_out, vals_out, keys_alt, vals_alt, temp};
  return sort_pairs_batched<uint64_t>(s, &g, 1, end_bit, 0, false);
}

}  // namespace gsr
