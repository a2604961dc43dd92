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
// Onesweep radix sort
// =================================================================================================
constexpr int RS_THREADS = 256;
constexpr int RS_WARPS = RS_THREADS / 32;
constexpr int RS_RADIX = 256;
constexpr int RS_MAX_PASSES = 8;
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
template <> struct RsCfg<uint32_t> { static constexpr int ITEMS = 16; };
template <> struct RsCfg<uint64_t> { static constexpr int ITEMS = 12; };

static int rs_tile(int key_bytes) {
  return RS_THREADS * (key_bytes == 8 ? RsCfg<uint64_t>::ITEMS : RsCfg<uint32_t>::ITEMS);
}
static int rs_passes(int end_bit) { return (end_bit + 7) / 8; }

// temp layout: [hist: passes*256 u32][tickets: 8 u32 (padded)][status: passes * tiles * 256 u32]
size_t sort_temp_bytes(int64_t n, int key_bytes, int end_bit) {
  const int passes = rs_passes(end_bit);
  const int64_t tiles = (n + rs_tile(key_bytes) - 1) / rs_tile(key_bytes);
  return align_up((size_t)RS_MAX_PASSES * RS_RADIX * 4) + align_up(64) +
         align_up((size_t)passes * (size_t)(tiles > 0 ? tiles : 1) * RS_RADIX * rs_status_bytes(n));
}

template <typename KeyT>
__global__ void __launch_bounds__(256)
rs_histogram_kernel(const KeyT* __restrict__ keys, int64_t n_cap, const uint32_t* __restrict__ n_dev,
                    int end_bit, uint32_t* __restrict__ hist) {
  __shared__ uint32_t s_h[RS_MAX_PASSES * RS_RADIX];
  const int64_t n = n_dev ? min((int64_t)*n_dev, n_cap) : n_cap;
  const int passes = (end_bit + 7) / 8;
  for (int i = threadIdx.x; i < passes * RS_RADIX; i += blockDim.x) s_h[i] = 0;
  __syncthreads();
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int lane = threadIdx.x & 31;
  // warp-uniform trip count so that the full-mask match below is legal
  for (int64_t b = (int64_t)blockIdx.x * blockDim.x + (threadIdx.x - lane); b < n; b += stride) {
    const int64_t i = b + lane;
    const bool valid = i < n;
    const KeyT k = valid ? keys[i] : (KeyT)0;
    for (int p = 0; p < passes; p++) {
      const int bits = min(8, end_bit - 8 * p);
      // invalid lanes get a private pseudo-digit so they never pair up
      const uint32_t d = valid ? ((uint32_t)(k >> (8 * p)) & ((1u << bits) - 1)) : (0x100u | lane);
      // warp-aggregate equal digits (tile-id keys are highly repetitive)
      const unsigned peers = __match_any_sync(0xffffffffu, d);
      if (valid && lane == __ffs(peers) - 1) atomicAdd(&s_h[p * RS_RADIX + d], __popc(peers));
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < passes * RS_RADIX; i += blockDim.x)
    if (s_h[i]) atomicAdd(&hist[i], s_h[i]);
}

// exclusive scan of each pass' 256-bin histogram, in place; one block of 256 threads per pass
__global__ void __launch_bounds__(RS_RADIX) rs_scan_hist_kernel(uint32_t* __restrict__ hist) {
  __shared__ uint32_t s_w[RS_RADIX / 32];
  uint32_t* h = hist + blockIdx.x * RS_RADIX;
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

template <typename KeyT, int ITEMS, typename StatusT>
__global__ void __launch_bounds__(RS_THREADS)
rs_onesweep_kernel(const KeyT* __restrict__ keys_in, const uint32_t* __restrict__ vals_in,
                   KeyT* __restrict__ keys_out, uint32_t* __restrict__ vals_out, int64_t n_cap,
                   const uint32_t* __restrict__ n_dev, int shift, int nbits,
                   const uint32_t* __restrict__ global_base, volatile StatusT* status,
                   uint32_t* ticket) {
  using ST = RsStatus<StatusT>;
  constexpr int TILE = RS_THREADS * ITEMS;
  // the element count may live on the device (no host round trip): grid is sized by capacity and
  // surplus CTAs retire before taking a ticket
  const int64_t n = n_dev ? min((int64_t)*n_dev, n_cap) : n_cap;
  if ((int64_t)blockIdx.x * TILE >= n) return;
  extern __shared__ __align__(16) unsigned char rs_smem[];
  KeyT* s_keys = reinterpret_cast<KeyT*>(rs_smem);                                  // [TILE]
  uint32_t* s_vals = reinterpret_cast<uint32_t*>(rs_smem + sizeof(KeyT) * TILE);    // [TILE]
  uint32_t* s_warp_cnt = s_vals + TILE;                                             // [WARPS][256]
  uint32_t* s_digit_off = s_warp_cnt + RS_WARPS * RS_RADIX;                         // [256]
  uint32_t* s_global_off = s_digit_off + RS_RADIX;                                  // [256]
  uint32_t* s_misc = s_global_off + RS_RADIX;                                       // [16]

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t mask = (1u << nbits) - 1;
  if (tid == 0) s_misc[0] = atomicAdd(ticket, 1u);
  for (int i = tid; i < RS_WARPS * RS_RADIX; i += RS_THREADS) s_warp_cnt[i] = 0;
  __syncthreads();
  const uint32_t tile = s_misc[0];
  const int64_t base = (int64_t)tile * TILE;
  const int64_t wbase = base + (int64_t)warp * 32 * ITEMS + lane;  // warp-striped arrangement

  KeyT key[ITEMS];
  uint32_t val[ITEMS];
#pragma unroll
  for (int r = 0; r < ITEMS; r++) {
    const int64_t idx = wbase + r * 32;
    if (idx < n) {
      key[r] = keys_in[idx];
      val[r] = vals_in ? vals_in[idx] : (uint32_t)idx;
    } else {
      key[r] = ~(KeyT)0;  // pads take the largest digit and sit at the very end of the tile
      val[r] = 0;
    }
  }

  // ---- stable ranking: per-warp digit counters + match_any multi-split ----
  uint32_t rank[ITEMS];
  const uint32_t lt_mask = (1u << lane) - 1;
#pragma unroll
  for (int r = 0; r < ITEMS; r++) {
    const uint32_t d = rs_digit(key[r], shift, mask);
    const unsigned peers = __match_any_sync(0xffffffffu, d);
    const int leader = __ffs(peers) - 1;
    uint32_t old = 0;
    if (lane == leader) {
      old = s_warp_cnt[warp * RS_RADIX + d];
      s_warp_cnt[warp * RS_RADIX + d] = old + __popc(peers);
    }
    old = __shfl_sync(0xffffffffu, old, leader);
    rank[r] = old + __popc(peers & lt_mask);
    __syncwarp();
  }
  __syncthreads();

  // ---- thread d owns digit d: warp offsets, tile count, look-back ----
  {
    const int d = tid;
    uint32_t run = 0;
#pragma unroll
    for (int w = 0; w < RS_WARPS; w++) {
      const uint32_t c = s_warp_cnt[w * RS_RADIX + d];
      s_warp_cnt[w * RS_RADIX + d] = run;
      run += c;
    }
    const uint32_t count = run;
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

    StatusT excl = 0;
    volatile StatusT* my = status + (size_t)tile * RS_RADIX + d;
    if (tile == 0) {
      *my = ST::INC | (StatusT)count;
    } else {
      *my = ST::AGG | (StatusT)count;
      int64_t t = (int64_t)tile - 1;
      while (true) {
        StatusT st;
        do { st = status[(size_t)t * RS_RADIX + d]; } while ((st >> ST::SHIFT) == 0);
        excl += st & ST::MASK;
        if (st & ST::INC) break;
        t--;
      }
      *my = ST::INC | ((excl + (StatusT)count) & ST::MASK);
    }
    s_global_off[d] = global_base[d] + (uint32_t)excl - digit_off;   // positions are < 2^32: mod-2^32 arithmetic
  }
  __syncthreads();

  // ---- reorder through shared memory so that the global scatter is run-wise contiguous ----
#pragma unroll
  for (int r = 0; r < ITEMS; r++) {
    const uint32_t d = rs_digit(key[r], shift, mask);
    const uint32_t pos = s_digit_off[d] + s_warp_cnt[warp * RS_RADIX + d] + rank[r];
    s_keys[pos] = key[r];
    s_vals[pos] = val[r];
  }
  __syncthreads();
  const int64_t remain = n - base;
  const int n_valid = remain < TILE ? (int)remain : TILE;
#pragma unroll 4
  for (int i = tid; i < n_valid; i += RS_THREADS) {
    const KeyT k = s_keys[i];
    const uint32_t d = rs_digit(k, shift, mask);
    const uint32_t g = s_global_off[d] + (uint32_t)i;  // mod-2^32 arithmetic (offset may have wrapped)
    keys_out[g] = k;
    vals_out[g] = s_vals[i];
  }
}

template <typename KeyT>
static cudaError_t sort_pairs_impl(cudaStream_t s, int64_t n, const uint32_t* n_dev, const KeyT* keys_in,
                                   const uint32_t* vals_in, KeyT* keys_out, uint32_t* vals_out,
                                   KeyT* keys_alt, uint32_t* vals_alt, int end_bit, char* temp,
                                   bool have_bases = false) {
  if (n <= 0) return cudaSuccess;
  constexpr int ITEMS = RsCfg<KeyT>::ITEMS;
  constexpr int TILE = RS_THREADS * ITEMS;
  const int passes = rs_passes(end_bit);
  if (passes < 1 || passes > RS_MAX_PASSES) return cudaErrorInvalidValue;
  const int64_t tiles = (n + TILE - 1) / TILE;
  uint32_t* hist = reinterpret_cast<uint32_t*>(temp);
  uint32_t* tickets = reinterpret_cast<uint32_t*>(temp + align_up((size_t)RS_MAX_PASSES * RS_RADIX * 4));
  uint32_t* status = reinterpret_cast<uint32_t*>(temp + align_up((size_t)RS_MAX_PASSES * RS_RADIX * 4) + align_up(64));
  cudaError_t e;
  if (have_bases) {  // the caller has filled `hist` with exclusive digit bases: clear tickets + status only
    const size_t skip = align_up((size_t)RS_MAX_PASSES * RS_RADIX * 4);
    e = cudaMemsetAsync(temp + skip, 0, sort_temp_bytes(n, sizeof(KeyT), end_bit) - skip, s);
    if (e != cudaSuccess) return e;
    count_launch(passes);
  } else {
    e = cudaMemsetAsync(temp, 0, sort_temp_bytes(n, sizeof(KeyT), end_bit), s);
    if (e != cudaSuccess) return e;
    int hist_blocks = (int)((n + 256 * 16 - 1) / (256 * 16));
    if (hist_blocks > 148 * 8) hist_blocks = 148 * 8;
    rs_histogram_kernel<KeyT><<<hist_blocks, 256, 0, s>>>(keys_in, n, n_dev, end_bit, hist);
    rs_scan_hist_kernel<<<passes, RS_RADIX, 0, s>>>(hist);
    count_launch(2 + passes);
  }

  const size_t smem = sizeof(KeyT) * TILE + 4 * TILE + 4 * (RS_WARPS * RS_RADIX + 2 * RS_RADIX + 16);
  const bool wide = n >= RS_WIDE_FROM;
  static bool attr_set[2] = {false, false};
  if (!attr_set[wide]) {
    e = wide ? cudaFuncSetAttribute(rs_onesweep_kernel<KeyT, ITEMS, unsigned long long>,
                                    cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)
             : cudaFuncSetAttribute(rs_onesweep_kernel<KeyT, ITEMS, uint32_t>,
                                    cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    attr_set[wide] = true;
  }
  // ping-pong so that the LAST pass lands in (keys_out, vals_out); inputs are never written
  const KeyT* kin = keys_in;
  const uint32_t* vin = vals_in;
  for (int p = 0; p < passes; p++) {
    const bool to_out = ((passes - 1 - p) % 2) == 0;
    KeyT* kout = to_out ? keys_out : keys_alt;
    uint32_t* vout = to_out ? vals_out : vals_alt;
    const int bits = (end_bit - 8 * p) < 8 ? (end_bit - 8 * p) : 8;
    if (wide)
      rs_onesweep_kernel<KeyT, ITEMS, unsigned long long><<<(unsigned)tiles, RS_THREADS, smem, s>>>(
          kin, vin, kout, vout, n, n_dev, 8 * p, bits, hist + p * RS_RADIX,
          reinterpret_cast<unsigned long long*>(status) + (size_t)p * tiles * RS_RADIX, tickets + p);
    else
      rs_onesweep_kernel<KeyT, ITEMS, uint32_t><<<(unsigned)tiles, RS_THREADS, smem, s>>>(
          kin, vin, kout, vout, n, n_dev, 8 * p, bits, hist + p * RS_RADIX,
          status + (size_t)p * tiles * RS_RADIX, tickets + p);
    e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    kin = kout;
    vin = vout;
  }
  return cudaSuccess;
}

cudaError_t launch_sort_pairs_u32(cudaStream_t s, int64_t n, const uint32_t* n_dev, const uint32_t* keys_in,
                                  const uint32_t* vals_in, uint32_t* keys_out, uint32_t* vals_out,
                                  uint32_t* keys_alt, uint32_t* vals_alt, int end_bit, char* temp, bool have_bases) {
  return sort_pairs_impl<uint32_t>(s, n, n_dev, keys_in, vals_in, keys_out, vals_out, keys_alt, vals_alt, end_bit, temp,
                                   have_bases);
}
cudaError_t launch_sort_pairs_u64(cudaStream_t s, int64_t n, const uint32_t* n_dev, const uint64_t* keys_in,
                                  const uint32_t* vals_in, uint64_t* keys_out, uint32_t* vals_out,
                                  uint64_t* keys_alt, uint32_t* vals_alt, int end_bit, char* temp) {
  return sort_pairs_impl<uint64_t>(s, n, n_dev, keys_in, vals_in, keys_out, vals_out, keys_alt, vals_alt, end_bit, temp);
}

}  // namespace gsr
