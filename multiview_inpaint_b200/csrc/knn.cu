// knn.cu -- mean squared distance to the 3 nearest neighbours of every point.
//
// Replaces `simple_knn._C.distCUDA2`, which the reference imports at
// gs-simp/scene/gaussian_model.py:20 and calls at :134 (initial scales from the SfM cloud), :546 and :623
// (scales of newly sampled points).  Its source is a third-party extension that is NOT in the reference tree
// (environment.yml:17 names a local path that does not exist); the published algorithm is: Morton-sort the
// points, and for each point keep the 3 smallest squared distances d = dx*dx + dy*dy + dz*dz (strict '<'
// insertion) over all other points, pruning boxes of consecutive sorted points conservatively.  The result is
// therefore the EXACT 3-NN value (best0 + best1 + best2) / 3 whatever the traversal, which is what makes a
// bit-exact brute-force oracle possible (oracle/gsplat_oracle.c gso_knn3_mean_dist2).
//
// B200 design (not the upstream structure, which walks every 1024-point box for every point through an
// index indirection):
//   * bounds by ordered-uint atomics, 63-bit Morton keys, the repo's own onesweep sort (scan_sort.cu),
//     points GATHERED into Morton order as float4 so that every later read is a coalesced 512-byte line;
//   * a 32-ary bounding-box tree over the sorted order, built bottom-up one warp per node;
//   * search: ONE WARP per leaf of 32 queries.  The warp seeds its 32 running top-3 lists from its own leaf,
//     then walks the tree top-down with a stack held in lanes (lane l = level l), pruning a node when its
//     box-to-box distance to the leaf's own box exceeds the largest current third-best of the warp.
//     A surviving leaf is staged in a 512-byte per-warp shared-memory slot and broadcast-read (one
//     LDS.128 per candidate), the insertion is branch-free (5 min/max).
// Pruning is exact: the box distance uses the same monotone fp32 op sequence as the point distance
// (SUB, MUL, FMA, FMA), so computed box distance <= computed distance of every pair it bounds, and a node
// whose box distance is >= the warp's largest third-best cannot change any list (insertion is strict '<').
// That also keeps degenerate clouds (many coincident points: third-best = 0) from visiting every leaf.
#include "common.cuh"

#include <float.h>

namespace gsr {

namespace {

constexpr int KNN_WARPS = 8;           // warps (= leaves) per CTA in the search
constexpr int KNN_MAX_LEVELS = 6;      // 32^6 leaves*32 points: far beyond 2^30
constexpr unsigned FULL = 0xffffffffu;

struct Box { float4 lo, hi; };
struct KnnTree {
  const Box* boxes[KNN_MAX_LEVELS];
  int count[KNN_MAX_LEVELS];
  int levels;
};

__device__ __forceinline__ uint32_t f2ord(float f) {
  const uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ord2f(uint32_t u) {
  return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}

// bounds[0..2] = min (ordered uint, initialised to 0xFFFFFFFF), bounds[4..6] = max (initialised to 0)
__global__ void __launch_bounds__(256) knn_bounds_kernel(const float* __restrict__ pts, int P, uint32_t* __restrict__ bounds) {
  float lo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, hi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < P; i += gridDim.x * blockDim.x) {
#pragma unroll
    for (int c = 0; c < 3; c++) {
      const float v = __ldg(pts + 3 * (size_t)i + c);
      lo[c] = fminf(lo[c], v);
      hi[c] = fmaxf(hi[c], v);
    }
  }
#pragma unroll
  for (int c = 0; c < 3; c++) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      lo[c] = fminf(lo[c], __shfl_xor_sync(FULL, lo[c], o));
      hi[c] = fmaxf(hi[c], __shfl_xor_sync(FULL, hi[c], o));
    }
  }
  if ((threadIdx.x & 31) == 0) {
#pragma unroll
    for (int c = 0; c < 3; c++) {
      atomicMin(bounds + c, f2ord(lo[c]));
      atomicMax(bounds + 4 + c, f2ord(hi[c]));
    }
  }
}

__device__ __forceinline__ uint64_t spread21(uint32_t v) {  // 21 bits -> every third bit
  uint64_t x = v & 0x1fffffu;
  x = (x | (x << 32)) & 0x001f00000000ffffull;
  x = (x | (x << 16)) & 0x001f0000ff0000ffull;
  x = (x | (x << 8)) & 0x100f00f00f00f00full;
  x = (x | (x << 4)) & 0x10c30c30c30c30c3ull;
  x = (x | (x << 2)) & 0x1249249249249249ull;
  return x;
}

// 63-bit Morton keys (21 bits per axis).  Upstream uses 10 bits per axis; a dense cluster that falls into one
// 1/1024 cell then keeps its input order, its leaves become spatially random and no box can be pruned inside
// it (measured: 137 ms instead of a few ms at 3 M clustered points).  The order only affects speed, never values.
__global__ void __launch_bounds__(256) knn_morton_kernel(const float* __restrict__ pts, int P, const uint32_t* __restrict__ bounds,
                                                         uint64_t* __restrict__ keys) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P) return;
  uint64_t code = 0;
#pragma unroll
  for (int c = 0; c < 3; c++) {
    const float lo = ord2f(__ldg(bounds + c)), hi = ord2f(__ldg(bounds + 4 + c));
    const float ext = hi - lo;
    const float t = ext > 0.f ? (__ldg(pts + 3 * (size_t)i + c) - lo) / ext : 0.f;
    const int q = min(2097151, max(0, (int)(t * 2097151.f)));   // NaN -> 0
    code |= spread21((uint32_t)q) << c;
  }
  keys[i] = code;
}

// sorted[i] = (points[order[i]], 1) for i < P; (+inf, +inf, +inf, 0) for the padding up to a multiple of 32
__global__ void __launch_bounds__(256) knn_gather_kernel(const float* __restrict__ pts, const uint32_t* __restrict__ order, int P,
                                                         int P32, float4* __restrict__ sorted) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P32) return;
  float4 v = make_float4(__int_as_float(0x7f800000), __int_as_float(0x7f800000), __int_as_float(0x7f800000), 0.f);
  if (i < P) {
    const size_t s = 3 * (size_t)__ldg(order + i);
    v = make_float4(__ldg(pts + s), __ldg(pts + s + 1), __ldg(pts + s + 2), 1.f);
  }
  sorted[i] = v;
}

__device__ __forceinline__ void warp_box_reduce(float4& lo, float4& hi) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    lo.x = fminf(lo.x, __shfl_xor_sync(FULL, lo.x, o));
    lo.y = fminf(lo.y, __shfl_xor_sync(FULL, lo.y, o));
    lo.z = fminf(lo.z, __shfl_xor_sync(FULL, lo.z, o));
    hi.x = fmaxf(hi.x, __shfl_xor_sync(FULL, hi.x, o));
    hi.y = fmaxf(hi.y, __shfl_xor_sync(FULL, hi.y, o));
    hi.z = fmaxf(hi.z, __shfl_xor_sync(FULL, hi.z, o));
  }
}

// level 0: one warp per leaf of 32 sorted points (padding excluded)
__global__ void __launch_bounds__(256) knn_leaf_boxes_kernel(const float4* __restrict__ sorted, int P, int n_leaf, Box* __restrict__ out) {
  const int leaf = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (leaf >= n_leaf) return;
  const int lane = threadIdx.x & 31, i = leaf * 32 + lane;
  const float inf = __int_as_float(0x7f800000);
  float4 lo = make_float4(inf, inf, inf, 0.f), hi = make_float4(-inf, -inf, -inf, 0.f);
  if (i < P) { const float4 p = sorted[i]; lo = p; hi = p; }
  warp_box_reduce(lo, hi);
  if (lane == 0) { out[leaf].lo = lo; out[leaf].hi = hi; }
}
// level k > 0: one warp per node over 32 boxes of the level below
__global__ void __launch_bounds__(256) knn_node_boxes_kernel(const Box* __restrict__ below, int n_below, int n_nodes, Box* __restrict__ out) {
  const int node = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (node >= n_nodes) return;
  const int lane = threadIdx.x & 31, i = node * 32 + lane;
  const float inf = __int_as_float(0x7f800000);
  float4 lo = make_float4(inf, inf, inf, 0.f), hi = make_float4(-inf, -inf, -inf, 0.f);
  if (i < n_below) { lo = below[i].lo; hi = below[i].hi; }
  warp_box_reduce(lo, hi);
  if (lane == 0) { out[node].lo = lo; out[node].hi = hi; }
}

// squared distance, the op sequence of the oracle: fma(dz,dz, fma(dy,dy, dx*dx))
__device__ __forceinline__ float dist2(float dx, float dy, float dz) { return FMA(dz, dz, FMA(dy, dy, MUL(dx, dx))); }
__device__ __forceinline__ float box_dist2(const float4& alo, const float4& ahi, const float4& blo, const float4& bhi) {
  const float gx = fmaxf(0.f, fmaxf(SUB(alo.x, bhi.x), SUB(blo.x, ahi.x)));
  const float gy = fmaxf(0.f, fmaxf(SUB(alo.y, bhi.y), SUB(blo.y, ahi.y)));
  const float gz = fmaxf(0.f, fmaxf(SUB(alo.z, bhi.z), SUB(blo.z, ahi.z)));
  return dist2(gx, gy, gz);
}

__global__ void __launch_bounds__(KNN_WARPS * 32)
knn3_search_kernel(const float4* __restrict__ sorted, const uint32_t* __restrict__ order, int P, KnnTree tree, float* __restrict__ out) {
  __shared__ float4 s_slot[KNN_WARPS][32];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int leaf = blockIdx.x * KNN_WARPS + warp;
  if (leaf >= tree.count[0]) return;          // warp-uniform; only __syncwarp below
  float4* slot = s_slot[warp];
  const int qi = leaf * 32 + lane;
  const bool valid = qi < P;
  const float4 q = sorted[qi];                // the array is padded to a multiple of 32
  const float inf = __int_as_float(0x7f800000);
  float b0 = FLT_MAX, b1 = FLT_MAX, b2 = FLT_MAX;

#define KNN_INSERT(d)                    \
  {                                      \
    const float _d = (d);                \
    b2 = fminf(b2, fmaxf(b1, _d));       \
    b1 = fminf(b1, fmaxf(b0, _d));       \
    b0 = fminf(b0, _d);                  \
  }

  // own leaf: every other point of the leaf (a point is never its own neighbour; coincident points are)
  slot[lane] = q;
  __syncwarp();
#pragma unroll 8
  for (int j = 0; j < 32; j++) {
    const float4 c = slot[j];
    float d = dist2(SUB(q.x, c.x), SUB(q.y, c.y), SUB(q.z, c.z));
    if (j == lane) d = inf;
    KNN_INSERT(d);
  }
  const Box qb = tree.boxes[0][leaf];
  // largest third-best of the warp's valid queries (non-negative floats order like their bit patterns)
  float rmax = __uint_as_float(__reduce_max_sync(FULL, valid ? __float_as_uint(b2) : 0u));

  const int T = tree.levels - 1;
  unsigned st_mask = 0;   // lane l: children of the current level-(l+1) node still to visit at level l
  int st_base = 0;        // lane l: index of the first of those children within level l
  // lane l also keeps level l's box array and node count (static indexing of the kernel parameter: no local copy)
  unsigned long long lv_boxes = 0;
  int lv_count = 0;
#pragma unroll
  for (int l = 0; l < KNN_MAX_LEVELS; l++)
    if (lane == l) { lv_boxes = (unsigned long long)tree.boxes[l]; lv_count = tree.count[l]; }
  {
    bool hit = false;
    const Box* top = reinterpret_cast<const Box*>(__shfl_sync(FULL, lv_boxes, T));
    if (lane < __shfl_sync(FULL, lv_count, T)) {
      const Box b = top[lane];
      hit = box_dist2(qb.lo, qb.hi, b.lo, b.hi) < rmax;
    }
    const unsigned m = __ballot_sync(FULL, hit);
    if (lane == T) { st_mask = m; st_base = 0; }
  }
  int cur = T;
  while (true) {
    const unsigned m = __shfl_sync(FULL, st_mask, cur);
    if (m == 0) {
      if (cur == T) break;
      ++cur;
      continue;
    }
    const int j = __ffs(m) - 1;
    if (lane == cur) st_mask = m & (m - 1);
    const int node = __shfl_sync(FULL, st_base, cur) + j;
    if (cur == 0) {
      if (node == leaf) continue;
      const float4 c4 = sorted[(size_t)node * 32 + lane];
      __syncwarp();
      slot[lane] = c4;
      __syncwarp();
#pragma unroll 8
      for (int k = 0; k < 32; k++) {
        const float4 c = slot[k];   // padding is +inf: its distance is +inf and is never inserted
        KNN_INSERT(dist2(SUB(q.x, c.x), SUB(q.y, c.y), SUB(q.z, c.z)));
      }
      rmax = __uint_as_float(__reduce_max_sync(FULL, valid ? __float_as_uint(b2) : 0u));
    } else {
      const int cb = node * 32;
      const int nchild = min(32, __shfl_sync(FULL, lv_count, cur - 1) - cb);
      const Box* below = reinterpret_cast<const Box*>(__shfl_sync(FULL, lv_boxes, cur - 1));
      bool hit = false;
      if (lane < nchild) {
        const Box b = below[cb + lane];
        hit = box_dist2(qb.lo, qb.hi, b.lo, b.hi) < rmax;
      }
      const unsigned cm = __ballot_sync(FULL, hit);
      if (lane == cur - 1) { st_mask = cm; st_base = cb; }
      --cur;
    }
  }
#undef KNN_INSERT
  if (valid) out[__ldg(order + qi)] = DIV(ADD(ADD(b0, b1), b2), 3.0f);
}

struct KnnLayout {
  size_t bounds, keys, keys_sorted, order, keys_alt, vals_alt, sort_temp, sorted, boxes, bytes;
  int count[KNN_MAX_LEVELS];
  size_t box_off[KNN_MAX_LEVELS];
  int levels;
};
KnnLayout knn_layout(int P) {
  KnnLayout L{};
  size_t o = 0;
  auto take = [&](size_t bytes) { const size_t r = o; o += align_up(bytes); return r; };
  const int P32 = (P + 31) / 32 * 32;
  L.bounds = take(32);
  L.keys = take((size_t)P * 8);
  L.keys_sorted = take((size_t)P * 8);
  L.order = take((size_t)P * 4);
  L.keys_alt = take((size_t)P * 8);
  L.vals_alt = take((size_t)P * 4);
  L.sort_temp = take(sort_temp_bytes(P, 8, 63));
  L.sorted = take((size_t)P32 * sizeof(float4));
  int n = P32 / 32, lv = 0;
  L.boxes = o;
  while (true) {
    L.count[lv] = n;
    L.box_off[lv] = take((size_t)n * sizeof(Box));
    lv++;
    if (n <= 32 || lv == KNN_MAX_LEVELS) break;
    n = (n + 31) / 32;
  }
  L.levels = lv;
  L.bytes = o;
  return L;
}

}  // namespace

size_t knn_temp_bytes(int P) { return P > 0 ? knn_layout(P).bytes : 0; }

cudaError_t launch_knn3_mean_dist2(cudaStream_t s, int P, const float* points, float* mean_dist2, char* temp) {
  if (P <= 0) return cudaSuccess;
  const KnnLayout L = knn_layout(P);
  if (L.count[L.levels - 1] > 32) return cudaErrorInvalidValue;   // cannot happen below 2^30 points
  const int P32 = (P + 31) / 32 * 32;
  uint32_t* bounds = reinterpret_cast<uint32_t*>(temp + L.bounds);
  uint64_t* keys = reinterpret_cast<uint64_t*>(temp + L.keys);
  uint64_t* keys_sorted = reinterpret_cast<uint64_t*>(temp + L.keys_sorted);
  uint32_t* order = reinterpret_cast<uint32_t*>(temp + L.order);
  float4* sorted = reinterpret_cast<float4*>(temp + L.sorted);
  cudaError_t e;
  if ((e = cudaMemsetAsync(bounds, 0xff, 16, s)) != cudaSuccess) return e;
  if ((e = cudaMemsetAsync(bounds + 4, 0x00, 16, s)) != cudaSuccess) return e;
  int blocks = cdiv(P, 256 * 8);
  if (blocks > 148 * 8) blocks = 148 * 8;
  knn_bounds_kernel<<<blocks, 256, 0, s>>>(points, P, bounds);
  knn_morton_kernel<<<cdiv(P, 256), 256, 0, s>>>(points, P, bounds, keys);
  count_launch(2);
  e = launch_sort_pairs_u64(s, P, nullptr, keys, nullptr, keys_sorted, order, reinterpret_cast<uint64_t*>(temp + L.keys_alt),
                            reinterpret_cast<uint32_t*>(temp + L.vals_alt), 63, temp + L.sort_temp);
  if (e != cudaSuccess) return e;
  knn_gather_kernel<<<cdiv(P32, 256), 256, 0, s>>>(points, order, P, P32, sorted);
  KnnTree tree{};
  tree.levels = L.levels;
  for (int l = 0; l < L.levels; l++) {
    tree.boxes[l] = reinterpret_cast<const Box*>(temp + L.box_off[l]);
    tree.count[l] = L.count[l];
  }
  knn_leaf_boxes_kernel<<<cdiv((int64_t)L.count[0] * 32, 256), 256, 0, s>>>(sorted, P, L.count[0], reinterpret_cast<Box*>(temp + L.box_off[0]));
  for (int l = 1; l < L.levels; l++)
    knn_node_boxes_kernel<<<cdiv((int64_t)L.count[l] * 32, 256), 256, 0, s>>>(tree.boxes[l - 1], L.count[l - 1], L.count[l],
                                                                             reinterpret_cast<Box*>(temp + L.box_off[l]));
  knn3_search_kernel<<<cdiv(L.count[0], KNN_WARPS), KNN_WARPS * 32, 0, s>>>(sorted, order, P, tree, mean_dist2);
  count_launch(2 + L.levels);
  return cudaGetLastError();
}

}  // namespace gsr
