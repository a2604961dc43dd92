// api.cu -- the extern "C" boundary declared in include/gsrast_b200.h.
//
// gsr_forward / gsr_backward / gsr_mark_visible stand where the reference's
// CudaRasterizer::Rasterizer::{forward,backward,markVisible} stand, driven
// by rasterize_gaussians / rasterize_gaussians_backward / mark_visible, which
// /root/reference/gs-simp/gaussian_renderer/__init__.py:85-93 reaches through
// GaussianRasterizer.forward.  Kernel order per view:
//   forward : K1 preprocess -> [depth sort] -> K2 scan -> (N to host) -> K3 duplicate -> K4 sort
//             -> K5 ranges -> K6 blend
//   backward: K7 blend backward -> K8+K9 per-Gaussian backward
#include <algorithm>
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include <nvtx3/nvToolsExt.h>   // header-only; a no-op unless a tool (nsys / ncu --nvtx) has injected itself

#include "common.cuh"

namespace gsr {

static thread_local std::string g_last_error;
void set_error(const char* msg) { g_last_error = msg; }

static std::atomic<uint64_t> g_launches{0};
void count_launch(int n) { g_launches.fetch_add((uint64_t)n, std::memory_order_relaxed); }

// ---- stage profiler: CUDA events on the launching stream around every stage ----
struct Profiler {
  bool on = false;
  std::vector<cudaEvent_t> pool;   // created lazily, reused
  std::vector<int> stage;          // stage id of mark i, -1 = end of a bracket
  size_t used = 0;
  void mark(cudaStream_t s, int st) {
    if (!on) return;
    if (used == pool.size()) {
      cudaEvent_t e;
      if (cudaEventCreate(&e) != cudaSuccess) return;
      pool.push_back(e);
    }
    cudaEventRecord(pool[used], s);
    if (stage.size() <= used) stage.push_back(st); else stage[used] = st;
    used++;
  }
};
static Profiler g_prof;

// NVTX ranges around the same stage brackets (GSR_NVTX=1): one range per pipeline stage on the calling host thread, so that
// `nsys` / `ncu --nvtx --nvtx-include "gsr/blend backward/"` can address a stage by name.  Host-side markers only: they cost
// two library calls per stage and nothing at all when no tool is attached.
static const char* const kStageNames[GSR_NUM_STAGES] = {"preprocess", "depth sort", "scan", "duplicate", "tile sort", "tile ranges",
                                                        "blend forward", "accumulator clear", "blend backward",
                                                        "per-Gaussian backward"};
struct NvtxStages {
  bool on = false, open = false;
  nvtxDomainHandle_t dom = nullptr;
  NvtxStages() {
    const char* e = getenv("GSR_NVTX");
    on = e && e[0] && e[0] != '0';
    if (on) dom = nvtxDomainCreateA("gsr");
  }
  void mark(int st) {
    if (!on) return;
    if (open) { nvtxDomainRangePop(dom); open = false; }
    if (st < 0 || st >= GSR_NUM_STAGES) return;
    nvtxEventAttributes_t a{};
    a.version = NVTX_VERSION;
    a.size = NVTX_EVENT_ATTRIB_STRUCT_SIZE;
    a.messageType = NVTX_MESSAGE_TYPE_ASCII;
    a.message.ascii = kStageNames[st];
    nvtxDomainRangePushEx(dom, &a);
    open = true;
  }
};
static thread_local NvtxStages g_nvtx;
#define PROF(st) do { g_prof.mark(s, (st)); g_nvtx.mark(st); } while (0)

static int fail_cuda(cudaError_t e, const char* where) {
  g_last_error = std::string(where) + ": " + cudaGetErrorString(e);
  return (int)e;
}
static int fail(int code, const char* msg) {
  g_last_error = msg;
  return code;
}

#define GSR_CUDA(expr, where)                         \
  do {                                                \
    cudaError_t _e = (expr);                          \
    if (_e != cudaSuccess) return fail_cuda(_e, where); \
  } while (0)

static bool refstruct(uint32_t flags) { return (flags & GSR_FLAG_REFERENCE) != 0; }
static bool key64(uint32_t flags) { return (flags & (GSR_FLAG_BINNING_KEY64 | GSR_FLAG_REFERENCE)) != 0; }

GeomLayout geom_layout(int P, uint32_t flags) {
  GeomLayout L{};
  size_t off = 0;
  const size_t n = (size_t)(P > 0 ? P : 1);
  auto take = [&](size_t bytes) { size_t o = off; off += align_up(bytes); return o; };
  L.rec = take(n * 48);
  L.depths = take(n * 4);
  L.clamped = take(n);
  L.tiles_touched = take(n * 4);
  L.point_offsets = take(n * 4);
  L.status = take(32);
  L.scan_temp_bytes = scan_temp_bytes(P);
  L.scan_temp = take(L.scan_temp_bytes);  // adjacent to status: one region to clear
  if (!key64(flags)) {
    L.depth_keys = take(n * 4);
    L.order = take(n * 4);
    L.depth_keys_alt = take(n * 4);
    L.order_alt = take(n * 4);
    L.rects = take(n * 8);
    L.temp_bytes = sort_temp_bytes(P, 4, 32);
    L.temp = take(L.temp_bytes);          // depth-sort temp (its own region: cleared up front with everything else)
  } else {
    L.depth_keys = L.order = L.depth_keys_alt = L.order_alt = L.rects = (size_t)-1;
    L.temp_bytes = 0;
    L.temp = (size_t)-1;
  }
  L.bytes = off;
  return L;
}

ImageLayout image_layout(int W, int H) {
  ImageLayout L{};
  size_t off = 0;
  const size_t hw = (size_t)W * H;
  const size_t G = (size_t)cdiv(W, TILE_X) * cdiv(H, TILE_Y);
  auto take = [&](size_t bytes) { size_t o = off; off += align_up(bytes); return o; };
  L.final_T = take(hw * 4);
  L.n_contrib = take(hw * 4);
  L.ranges = take((G > 0 ? G : 1) * 8);
  L.tile_count = take((G > 0 ? G : 1) * 4);
  L.bytes = off;
  return L;
}

static int tile_bits(int W, int H) {
  const uint32_t G = (uint32_t)cdiv(W, TILE_X) * (uint32_t)cdiv(H, TILE_Y);
  int b = ceil_log2(G);
  return b < 1 ? 1 : b;
}
static int key64_end_bit(int W, int H) {
  const uint32_t G = (uint32_t)cdiv(W, TILE_X) * (uint32_t)cdiv(H, TILE_Y);
  return 32 + (int)higher_msb(G);  // exactly the reference's bit range (Appendix A.4)
}

BinningLayout binning_layout(int64_t N, int W, int H, uint32_t flags) {
  // N is the CAPACITY the buffer is sized for.  The sorted list always lands at offset 0
  // (point_list), so the backward never needs to know the capacity.
  BinningLayout L{};
  size_t off = 0;
  const size_t n = (size_t)(N > 0 ? N : 1);
  auto take = [&](size_t bytes) { size_t o = off; off += align_up(bytes); return o; };
  const bool k64 = key64(flags);
  const int end_bit = k64 ? key64_end_bit(W, H) : tile_bits(W, H);
  const size_t kb = k64 ? 8 : 4;
  L.point_list = take(n * 4);
  L.vals_alt = take(n * 4);
  L.keys_a = take(n * kb);
  L.keys_b = take(n * kb);
  L.temp_bytes = sort_temp_bytes(N, (int)kb, end_bit);
  if (refstruct(flags)) L.temp_bytes = std::max(L.temp_bytes, ref_sort_temp_bytes(N, end_bit));
  L.temp = take(L.temp_bytes);
  L.big_items = k64 ? (size_t)-1 : take((size_t)bin_big_capacity(N) * 16);
  L.bytes = off;
  return L;
}

static Camera make_camera(int W, int H, const float* view_d, const float* proj_d, const float* campos_d,
                          float tan_fovx, float tan_fovy, float scale_modifier) {
  Camera c;
  c.view = view_d;
  c.proj = proj_d;
  c.campos = campos_d;
  c.focal_y = H / (2.0f * tan_fovy);
  c.focal_x = W / (2.0f * tan_fovx);
  c.tan_fovx = tan_fovx;
  c.tan_fovy = tan_fovy;
  c.scale_modifier = scale_modifier;
  c.W = W;
  c.H = H;
  c.grid_x = cdiv(W, TILE_X);
  c.grid_y = cdiv(H, TILE_Y);
  return c;
}

// pinned host staging for the one value that must reach the host per view: N (+ the trap flag)
struct Pinned {
  int64_t* result = nullptr;  // 2 x int64
};
static Pinned& pinned() {
  static thread_local Pinned p;
  if (!p.result) {
    void* ptr = nullptr;
    if (cudaMallocHost(&ptr, 64) == cudaSuccess) p.result = reinterpret_cast<int64_t*>(ptr);
  }
  return p;
}

}  // namespace gsr

using namespace gsr;

extern "C" {

const char* gsr_last_error(void) { return g_last_error.c_str(); }
uint64_t gsr_kernel_launches(void) { return g_launches.load(std::memory_order_relaxed); }
void gsr_profile_enable(int on) {
  g_prof.on = on != 0;
  // create the event pool up front so that no cudaEventCreate lands inside a timed region
  while (g_prof.on && g_prof.pool.size() < 8192) {
    cudaEvent_t e;
    if (cudaEventCreate(&e) != cudaSuccess) break;
    g_prof.pool.push_back(e);
  }
}
int gsr_profile_collect(double* ms_host, int64_t* counts_host) {
  if (!ms_host || !counts_host) return fail(GSR_E_INVALID, "gsr_profile_collect: null argument");
  for (size_t i = 0; i + 1 < g_prof.used; i++) {
    const int st = g_prof.stage[i];
    if (st < 0 || st >= GSR_NUM_STAGES) continue;
    GSR_CUDA(cudaEventSynchronize(g_prof.pool[i + 1]), "profile sync");
    float ms = 0.f;
    GSR_CUDA(cudaEventElapsedTime(&ms, g_prof.pool[i], g_prof.pool[i + 1]), "profile elapsed");
    ms_host[st] += ms;
    counts_host[st] += 1;
  }
  g_prof.used = 0;
  return 0;
}
int gsr_version(void) { return GSR_VERSION; }

int gsr_debug_set(int knob, int value) {
  switch (knob) {
    case 0: if (value < 0 || value > 2) return fail(GSR_E_INVALID, "gsr_debug_set: rank mode 0..2"); g_rs_rank_mode = value; return 0;
    case 1: if (value < 0 || value > 2) return fail(GSR_E_INVALID, "gsr_debug_set: count mode 0..2"); g_bin_count_mode = value; return 0;
    case 2: g_pre_min_blocks = value; return 0;
    case 3: g_bwd_mma = value != 0; return 0;
    case 5: g_gather_bulk = value != 0; return 0;
    default: return fail(GSR_E_INVALID, "gsr_debug_set: unknown knob");
  }
}

int gsr_debug_approx_units(const float* x_dev, int n, float* out_dev, void* stream) {
  if (n <= 0) return 0;
  if (!x_dev || !out_dev) return fail(GSR_E_INVALID, "gsr_debug_approx_units: null pointer");
  const cudaError_t e = launch_debug_approx_units((cudaStream_t)stream, x_dev, n, out_dev);
  return e == cudaSuccess ? 0 : fail_cuda(e, "gsr_debug_approx_units");
}

size_t gsr_loss_temp_bytes(int C, int H, int W) {
  if (C <= 0 || H <= 0 || W <= 0) return 0;
  return loss_temp_bytes(C, H, W);
}
int gsr_loss_l1_ssim_forward(void* stream, int C, int H, int W, const float* image, const float* gt,
                             float lambda_dssim, float* out_loss3, char* temp, size_t temp_bytes) {
  if (C <= 0 || H <= 0 || W <= 0 || C > 65535) return fail(GSR_E_INVALID, "gsr_loss_l1_ssim_forward: bad shape");
  if (!image || !gt || !out_loss3 || !temp || temp_bytes < loss_temp_bytes(C, H, W))
    return fail(GSR_E_INVALID, "gsr_loss_l1_ssim_forward: null argument or temp too small");
  if ((reinterpret_cast<uintptr_t>(temp) & 15) != 0) return fail(GSR_E_INVALID, "gsr_loss_l1_ssim_forward: temp not 16-byte aligned");
  GSR_CUDA(launch_loss_forward(reinterpret_cast<cudaStream_t>(stream), C, H, W, image, gt, lambda_dssim, out_loss3, temp),
           "loss forward");
  return 0;
}
int gsr_loss_l1_ssim_backward(void* stream, int C, int H, int W, const float* image, const float* gt,
                              float lambda_dssim, const float* dL_dloss, const char* temp, size_t temp_bytes,
                              float* dL_dimage) {
  if (C <= 0 || H <= 0 || W <= 0 || C > 65535) return fail(GSR_E_INVALID, "gsr_loss_l1_ssim_backward: bad shape");
  if (!image || !gt || !dL_dimage || !temp || temp_bytes < loss_temp_bytes(C, H, W))
    return fail(GSR_E_INVALID, "gsr_loss_l1_ssim_backward: null argument or temp too small");
  GSR_CUDA(launch_loss_backward(reinterpret_cast<cudaStream_t>(stream), C, H, W, image, gt, lambda_dssim, dL_dloss, temp,
                                dL_dimage), "loss backward");
  return 0;
}

int gsr_activate_forward(void* stream, int P, const float* raw_scales, const float* raw_rotations,
                         const float* raw_opacities, float* scales, float* rotations, float* opacities) {
  if (P < 0 || (P > 0 && ((raw_scales && !scales) || (raw_rotations && !rotations) || (raw_opacities && !opacities))))
    return fail(GSR_E_INVALID, "gsr_activate_forward: bad argument");
  if (((reinterpret_cast<uintptr_t>(raw_rotations) | reinterpret_cast<uintptr_t>(rotations)) & 15) != 0)
    return fail(GSR_E_INVALID, "gsr_activate_forward: rotations not 16-byte aligned");
  GSR_CUDA(launch_activate_forward(reinterpret_cast<cudaStream_t>(stream), P, raw_scales, raw_rotations, raw_opacities,
                                   scales, rotations, opacities), "activate forward");
  return 0;
}
int gsr_activate_backward(void* stream, int P, const float* raw_scales, const float* raw_rotations,
                          const float* raw_opacities, float* g_scales, float* g_rotations, float* g_opacities) {
  if (P < 0 || (P > 0 && ((raw_scales && !g_scales) || (raw_rotations && !g_rotations) || (raw_opacities && !g_opacities))))
    return fail(GSR_E_INVALID, "gsr_activate_backward: bad argument");
  if (((reinterpret_cast<uintptr_t>(raw_rotations) | reinterpret_cast<uintptr_t>(g_rotations)) & 15) != 0)
    return fail(GSR_E_INVALID, "gsr_activate_backward: rotations not 16-byte aligned");
  GSR_CUDA(launch_activate_backward(reinterpret_cast<cudaStream_t>(stream), P, raw_scales, raw_rotations, raw_opacities,
                                    g_scales, g_rotations, g_opacities), "activate backward");
  return 0;
}

int gsr_adam_step(void* stream, const gsr_adam_segment* segs_host, int n_segs, int64_t step, double beta1,
                  double beta2, double eps) {
  if (n_segs < 0 || n_segs > 8 || (n_segs > 0 && !segs_host) || step < 1 || !(beta1 >= 0.0 && beta1 < 1.0) ||
      !(beta2 >= 0.0 && beta2 < 1.0))
    return fail(GSR_E_INVALID, "gsr_adam_step: bad argument (1..8 segments, step >= 1, betas in [0,1))");
  for (int k = 0; k < n_segs; k++) {
    const gsr_adam_segment& g = segs_host[k];
    if (g.n == 0) continue;
    if (!g.param || !g.grad || !g.exp_avg || !g.exp_avg_sq) return fail(GSR_E_INVALID, "gsr_adam_step: null pointer in a segment");
    if (g.row_len < 0 || (g.row_len > 0 && (g.row_split < 0 || g.row_split > g.row_len)))
      return fail(GSR_E_INVALID, "gsr_adam_step: bad row structure");
  }
  GSR_CUDA(launch_adam(reinterpret_cast<cudaStream_t>(stream), segs_host, n_segs, step, beta1, beta2, eps), "adam");
  return 0;
}

int gsr_quantize_rgb8(void* stream, int C, int H, int W, const float* image, const float* affine, uint8_t* rgb8) {
  if ((C != 1 && C != 3) || H < 0 || W < 0) return fail(GSR_E_INVALID, "gsr_quantize_rgb8: C must be 1 or 3");
  if ((size_t)H * (size_t)W == 0) return 0;
  if (!image || !rgb8) return fail(GSR_E_INVALID, "gsr_quantize_rgb8: null argument");
  GSR_CUDA(launch_quantize_rgb8(reinterpret_cast<cudaStream_t>(stream), C, H, W, image, affine, rgb8), "quantize_rgb8");
  return 0;
}

int gsr_gather_rows(void* stream, int64_t n_dst, int64_t n_src, int64_t n_keep_state, const int32_t* src_row,
                    const gsr_gather_segment* segs_host, int n_segs) {
  if (n_dst < 0 || n_src < 0 || n_keep_state < 0 || n_segs < 0 || n_segs > GSR_GATHER_MAX_SEGS)
    return fail(GSR_E_INVALID, "gsr_gather_rows: bad size (n_segs <= GSR_GATHER_MAX_SEGS)");
  if (n_dst >= ((int64_t)1 << 31) || n_src >= ((int64_t)1 << 31)) return fail(GSR_E_OVERFLOW, "gsr_gather_rows: more than 2^31 - 1 rows");
  if (n_dst == 0 || n_segs == 0) return 0;
  if (n_src == 0) return fail(GSR_E_INVALID, "gsr_gather_rows: rows requested from an empty source");
  if (!src_row || !segs_host) return fail(GSR_E_INVALID, "gsr_gather_rows: null argument");
  for (int s = 0; s < n_segs; s++) {
    const gsr_gather_segment& g = segs_host[s];
    if (g.row_f32 < 0 || g.row_f32 > 4096) return fail(GSR_E_INVALID, "gsr_gather_rows: row_f32 out of range");
    if (g.row_f32 == 0) continue;
    if (!g.src || !g.dst) return fail(GSR_E_INVALID, "gsr_gather_rows: null segment pointer");
    if ((g.row_f32 & 3) == 0 && (((uintptr_t)g.src | (uintptr_t)g.dst) & 15))
      return fail(GSR_E_INVALID, "gsr_gather_rows: segments with row_f32 % 4 == 0 must be 16-byte aligned");
    const char* s0 = (const char*)g.src; const char* s1 = s0 + (size_t)n_src * g.row_f32 * 4;
    const char* d0 = (const char*)g.dst; const char* d1 = d0 + (size_t)n_dst * g.row_f32 * 4;
    if (s0 < d1 && d0 < s1) return fail(GSR_E_INVALID, "gsr_gather_rows: src and dst overlap");
  }
  GSR_CUDA(launch_gather_rows(reinterpret_cast<cudaStream_t>(stream), (long long)n_dst, (long long)n_keep_state, src_row,
                              segs_host, n_segs), "gather_rows");
  return 0;
}

size_t gsr_knn_temp_bytes(int P) { return knn_temp_bytes(P); }
int gsr_knn3_mean_dist2(void* stream, int P, const float* points, float* mean_dist2, char* temp, size_t temp_bytes) {
  if (P < 0) return fail(GSR_E_INVALID, "gsr_knn3_mean_dist2: P < 0");
  if (P == 0) return 0;
  if (P >= (1 << 30)) return fail(GSR_E_OVERFLOW, "gsr_knn3_mean_dist2: P >= 2^30");
  if (!points || !mean_dist2 || !temp || temp_bytes < knn_temp_bytes(P))
    return fail(GSR_E_INVALID, "gsr_knn3_mean_dist2: null argument or temp too small");
  GSR_CUDA(launch_knn3_mean_dist2(reinterpret_cast<cudaStream_t>(stream), P, points, mean_dist2, temp), "knn3_mean_dist2");
  return 0;
}

size_t gsr_backward_scratch_bytes(int P) { return align_up((size_t)(P > 0 ? P : 1) * 48); }
size_t gsr_sort_temp_bytes(int64_t n, int key_bytes, int end_bit) {
  // internal temp + alternate key/value buffers (inputs are preserved by the public entry points)
  const size_t nn = (size_t)(n > 0 ? n : 1);
  return sort_temp_bytes(n, key_bytes, end_bit) + align_up(nn * key_bytes) + align_up(nn * 4);
}
size_t gsr_scan_temp_bytes(int64_t n) { return scan_temp_bytes(n); }

int gsr_get_layout(int P, int width, int height, int64_t num_rendered, uint32_t flags, gsr_layout* out) {
  if (!out || P < 0 || width <= 0 || height <= 0 || num_rendered < 0) return fail(GSR_E_INVALID, "gsr_get_layout: bad argument");
  const GeomLayout g = geom_layout(P, flags);
  const ImageLayout im = image_layout(width, height);
  const BinningLayout b = binning_layout(num_rendered, width, height, flags);
  out->rec = g.rec;
  out->depths = g.depths;
  out->clamped = g.clamped;
  out->tiles_touched = g.tiles_touched;
  out->point_offsets = g.point_offsets;
  out->order = g.order;
  out->geom_bytes = g.bytes;
  out->final_T = im.final_T;
  out->n_contrib = im.n_contrib;
  out->ranges = im.ranges;
  out->image_bytes = im.bytes;
  out->point_list = b.point_list;
  out->binning_bytes = b.bytes;
  return 0;
}

}  // extern "C" (reopened below)

// ---- the forward pipeline over a batch of views (one launch per stage for all of them) ------------------------
namespace gsr {
struct FwdView {   // host-side working set of one view
  Camera cam;
  const float* bg;
  float* out_color;
  float* out_depth;
  int32_t* radii;
  char* geom;
  char* img;
  char* bin;       // set once the capacity is known
  char* scratch;   // optional: the backward's accumulator, cleared by K1
  int64_t cap;
  int G, end_bit;
};
static int32_t* status_of(const FwdView& v, const GeomLayout& gl) { return reinterpret_cast<int32_t*>(v.geom + gl.status); }

// K1 (+ depth sort in the two-level scheme, + scan in the literal one) for `nv` views.  `also_binning`: the binning
// buffers are already there (speculative / asynchronous path), so their counters are cleared by the same launch.
static int front_end(cudaStream_t s, FwdView* v, int nv, int P, int D, int M, const float* means3D, const float* shs,
                     const float* colors_precomp, const float* opacities, const float* scales, float scale_modifier,
                     const float* rotations, const float* cov3D_precomp, int prefiltered, uint32_t flags, bool also_binning);
static int bin_and_blend(cudaStream_t s, FwdView* v, int nv, int P, uint32_t flags, bool clear);

// what the expansion's counting scheme needs cleared: the per-tile counters (mode 1)
static ZeroRegion tile_counter_region(const FwdView& w, const ImageLayout& il) {
  return {w.img + il.tile_count, (size_t)w.G * 4};   // G words: cleared whatever scheme bin_and_blend picks
}
static void binning_zero_regions(const FwdView& w, const GeomLayout& gl, const ImageLayout& il, int P, uint32_t flags,
                                 std::vector<ZeroRegion>& z) {
  const BinningLayout bl = binning_layout(w.cap, w.cam.W, w.cam.H, flags);
  z.push_back({w.geom + gl.scan_temp, gl.scan_temp_bytes});
  z.push_back(tile_counter_region(w, il));
  z.push_back({w.bin + bl.temp, bl.temp_bytes});
  (void)P;
}

static int front_end(cudaStream_t s, FwdView* v, int nv, int P, int D, int M, const float* means3D, const float* shs,
                          const float* colors_precomp, const float* opacities, const float* scales, float scale_modifier,
                          const float* rotations, const float* cov3D_precomp, int prefiltered, uint32_t flags,
                          bool also_binning) {
  if (P <= 0 || nv <= 0) return 0;
  const bool k64 = key64(flags);
  const GeomLayout gl = geom_layout(P, flags);
  std::vector<ZeroRegion> z;
  std::vector<PreView> pv((size_t)nv);
  for (int k = 0; k < nv; k++) {
    const FwdView& w = v[k];
    const ImageLayout il = image_layout(w.cam.W, w.cam.H);
    if (!k64) z.push_back({w.geom + gl.temp, gl.temp_bytes});
    if (also_binning && !k64) {   // status block + the adjacent scan temp as one region
      z.push_back({w.geom + gl.status, gl.scan_temp + gl.scan_temp_bytes - gl.status});
      const BinningLayout bl = binning_layout(w.cap, w.cam.W, w.cam.H, flags);
      z.push_back(tile_counter_region(w, il));
      z.push_back({w.bin + bl.temp, bl.temp_bytes});
    } else {
      z.push_back({w.geom + gl.status, 32});
    }
    PreView& q = pv[(size_t)k];
    q.cam = w.cam;
    q.radii = w.radii;
    q.rec = reinterpret_cast<float4*>(w.geom + gl.rec);
    q.depths = reinterpret_cast<float*>(w.geom + gl.depths);
    q.clamped = reinterpret_cast<uint8_t*>(w.geom + gl.clamped);
    q.tiles_touched = reinterpret_cast<uint32_t*>(w.geom + gl.tiles_touched);
    q.depth_keys = k64 ? nullptr : reinterpret_cast<uint32_t*>(w.geom + gl.depth_keys);
    q.rects = k64 ? nullptr : reinterpret_cast<ushort4*>(w.geom + gl.rects);
    q.status = status_of(w, gl);
    q.gacc = reinterpret_cast<float4*>(w.scratch);
  }
  GSR_CUDA(launch_zero_regions(s, z.data(), (int)z.size()), "clear counters");
  PROF(0);
  GSR_CUDA(launch_preprocess(s, P, D, M, means3D, scales, rotations, opacities, shs, cov3D_precomp, colors_precomp,
                             scale_modifier, prefiltered, pv.data(), nv, !k64 && (flags & GSR_FLAG_TIGHT_BINNING) != 0), "preprocess");
  if (k64) {
    PROF(2);
    for (int k = 0; k < nv; k++)
      GSR_CUDA(launch_inclusive_scan(s, P, pv[(size_t)k].tiles_touched, nullptr,
                                     reinterpret_cast<uint32_t*>(v[k].geom + gl.point_offsets), v[k].geom + gl.scan_temp), "scan");
  } else {
    PROF(1);
    // Stable sort of (depth bits, index); culled Gaussians (key 0xFFFFFFFF) are DROPPED by the first pass, so the
    // sorted pairs are the V = status[5] visible ones.  32 bits = 4 passes (even): a -> b -> a -> b -> a, the result
    // lands back in (depth_keys, order); pass 0 takes the element index as the value.
    std::vector<SortSeg<uint32_t>> segs((size_t)nv);
    for (int k = 0; k < nv; k++) {
      char* g = v[k].geom;
      SortSeg<uint32_t>& q = segs[(size_t)k];
      q.n = P;
      q.n_dev = nullptr;
      q.n_dev_compact = reinterpret_cast<const uint32_t*>(status_of(v[k], gl) + 5);
      q.keys_in = reinterpret_cast<uint32_t*>(g + gl.depth_keys);
      q.vals_in = nullptr;
      q.keys_out = reinterpret_cast<uint32_t*>(g + gl.depth_keys);
      q.vals_out = reinterpret_cast<uint32_t*>(g + gl.order);
      q.keys_alt = reinterpret_cast<uint32_t*>(g + gl.depth_keys_alt);
      q.vals_alt = reinterpret_cast<uint32_t*>(g + gl.order_alt);
      q.temp = g + gl.temp;
    }
    GSR_CUDA(launch_sort_pairs_u32_batched(s, segs.data(), nv, 32, false, true, true), "depth sort");
  }
  PROF(-1);
  return 0;
}

// Binning + blend of `nv` views for the capacities in v[k].cap (v[k].bin allocated for them); every kernel reads the
// true N / V on the device.  `clear`: the binning counters were not cleared by front_end (exact-size path, re-binning).
static int bin_and_blend(cudaStream_t s, FwdView* v, int nv, int P, uint32_t flags, bool clear) {
  const bool k64 = key64(flags), ref = refstruct(flags);
  const GeomLayout gl = geom_layout(P > 0 ? P : 0, flags);
  std::vector<BlendFwdView> bv((size_t)nv);
  bool any = false;
  for (int k = 0; k < nv; k++) any = any || (P > 0 && v[k].cap > 0);
  if (any && !k64) {
    std::vector<ZeroRegion> z;
    std::vector<BinView> bins;
    std::vector<PrepView> preps;
    std::vector<SortSeg<uint32_t>> segs;
    std::vector<RangeView> rviews;
    // How the expansion counts (binning.cu): per-CTA digit histograms + a range search in the sorted keys (2) win where
    // rects are small -- headline 0.079 -> 0.061 + 0.009 ms, SVD orbit 0.109 -> 0.051 + 0.005 ms -- and lose 2 % where a few
    // hundred tiles per Gaussian make the per-instance reductions cheap next to the stores (4K stress shape): there, mode 1.
    int count_mode = g_bin_count_mode;
    if (count_mode == 2)
      for (int k = 0; k < nv; k++)
        if (v[k].cap > 32 * (int64_t)(P > 0 ? P : 1)) count_mode = 1;
    const int bases = count_mode == 2 ? 2 : 1;
    int end_bit = 0;
    bool same_bits = true;
    for (int k = 0; k < nv; k++) {
      FwdView& w = v[k];
      const ImageLayout il = image_layout(w.cam.W, w.cam.H);
      if (w.cap <= 0) {
        GSR_CUDA(cudaMemsetAsync(w.img + il.ranges, 0, (size_t)w.G * sizeof(uint2), s), "memset ranges");
        continue;
      }
      const BinningLayout bl = binning_layout(w.cap, w.cam.W, w.cam.H, flags);
      if (clear) {
        binning_zero_regions(w, gl, il, P, flags, z);
        z.push_back({status_of(w, gl) + 4, 4});   // work-list length
      }
      uint32_t* point_list = reinterpret_cast<uint32_t*>(w.bin + bl.point_list);
      uint32_t* vals_alt = reinterpret_cast<uint32_t*>(w.bin + bl.vals_alt);
      uint32_t* ka = reinterpret_cast<uint32_t*>(w.bin + bl.keys_a);
      uint32_t* kb = reinterpret_cast<uint32_t*>(w.bin + bl.keys_b);
      const int passes = (w.end_bit + 7) / 8;
      // emit into "a"; with an even pass count the sorted result lands back in "a", with an odd one in "b" -- pick
      // a so that the result is always point_list (offset 0)
      uint32_t* va = (passes & 1) ? vals_alt : point_list;
      uint32_t* valt = (passes & 1) ? va : vals_alt;
      uint32_t* kout = (passes & 1) ? kb : ka;
      uint32_t* kalt = (passes & 1) ? ka : kb;
      BinView b{};
      b.order = reinterpret_cast<const uint32_t*>(w.geom + gl.order);
      b.rects = reinterpret_cast<const ushort4*>(w.geom + gl.rects);
      b.offsets = reinterpret_cast<uint32_t*>(w.geom + gl.point_offsets);
      b.tile_keys = ka;
      b.vals = va;
      b.cap = (uint32_t)w.cap;
      b.tile_count = reinterpret_cast<uint32_t*>(w.img + il.tile_count);
      b.hist = sort_hist_ptr(w.bin + bl.temp);
      b.end_bit = w.end_bit;
      b.gx = w.cam.grid_x;
      b.ticket = reinterpret_cast<uint32_t*>(w.geom + gl.scan_temp);
      b.lb_status = reinterpret_cast<unsigned long long*>(w.geom + gl.scan_temp + align_up(16));
      b.status = status_of(w, gl);
      b.big_items = reinterpret_cast<uint4*>(w.bin + bl.big_items);
      b.big_cap = (uint32_t)bin_big_capacity(w.cap);
      b.P = P;
      bins.push_back(b);
      PrepView pr{};
      pr.G = w.G;
      pr.tile_count = b.tile_count;
      pr.ranges = reinterpret_cast<uint2*>(w.img + il.ranges);
      pr.end_bit = w.end_bit;
      pr.hist = sort_hist_ptr(w.bin + bl.temp);
      preps.push_back(pr);
      SortSeg<uint32_t> q{};
      q.n = w.cap;
      q.n_dev = reinterpret_cast<const uint32_t*>(status_of(w, gl) + 2);   // N (low word), written by K1
      q.n_dev_compact = nullptr;
      q.keys_in = ka;
      q.vals_in = va;
      q.keys_out = kout;
      q.vals_out = point_list;
      q.keys_alt = kalt;
      q.vals_alt = valt;
      q.temp = w.bin + bl.temp;
      segs.push_back(q);
      RangeView rv{};
      rv.keys = kout;
      rv.n_dev = q.n_dev;
      rv.cap = w.cap;
      rv.ranges = pr.ranges;
      rv.G = w.G;
      rviews.push_back(rv);
      if (end_bit == 0) end_bit = w.end_bit;
      same_bits = same_bits && end_bit == w.end_bit;
    }
    if (!z.empty()) GSR_CUDA(launch_zero_regions(s, z.data(), (int)z.size()), "clear binning counters");
    PROF(3);
    GSR_CUDA(launch_bin_expand(s, bins.data(), (int)bins.size(), count_mode), "scan + duplicate (depth order)");
    if (count_mode != 2) {
      PROF(5);
      GSR_CUDA(launch_tile_prepare(s, preps.data(), (int)preps.size()), "tile ranges + digit bases");
    }
    PROF(4);
    if (same_bits) {
      GSR_CUDA(launch_sort_pairs_u32_batched(s, segs.data(), (int)segs.size(), end_bit, bases, false, true), "tile sort");
    } else {  // views with different tile-id widths: one sort per view
      for (size_t k = 0; k < segs.size(); k++)
        GSR_CUDA(launch_sort_pairs_u32_batched(s, &segs[k], 1, preps[k].end_bit, bases, false, true), "tile sort");
    }
    if (count_mode == 2) {
      PROF(5);
      GSR_CUDA(launch_tile_ranges_views(s, rviews.data(), (int)rviews.size()), "tile ranges");
    }
  } else if (any) {  // the literal 64-bit key paths: one view at a time
    for (int k = 0; k < nv; k++) {
      FwdView& w = v[k];
      const ImageLayout il = image_layout(w.cam.W, w.cam.H);
      uint2* ranges = reinterpret_cast<uint2*>(w.img + il.ranges);
      if (w.cap <= 0) {
        GSR_CUDA(cudaMemsetAsync(ranges, 0, (size_t)w.G * sizeof(uint2), s), "memset ranges");
        continue;
      }
      const BinningLayout bl = binning_layout(w.cap, w.cam.W, w.cam.H, flags);
      const int end_bit = key64_end_bit(w.cam.W, w.cam.H);
      const int passes = (end_bit + 7) / 8;
      int32_t* status = status_of(w, gl);
      const uint32_t* n_dev = reinterpret_cast<const uint32_t*>(status + 2);
      uint32_t* point_list = reinterpret_cast<uint32_t*>(w.bin + bl.point_list);
      uint32_t* vals_alt = reinterpret_cast<uint32_t*>(w.bin + bl.vals_alt);
      uint32_t* va = (passes & 1) ? vals_alt : point_list;
      uint32_t* valt = (passes & 1) ? va : vals_alt;
      uint64_t* ka = reinterpret_cast<uint64_t*>(w.bin + bl.keys_a);
      uint64_t* kb = reinterpret_cast<uint64_t*>(w.bin + bl.keys_b);
      const float4* rec = reinterpret_cast<const float4*>(w.geom + gl.rec);
      const float* depths = reinterpret_cast<const float*>(w.geom + gl.depths);
      const uint32_t* offsets = reinterpret_cast<const uint32_t*>(w.geom + gl.point_offsets);
      if (ref) {  // reference structure: duplicateWithKeys -> cub SortPairs (unsorted -> sorted arrays) -> identifyTileRanges
        PROF(3);
        GSR_CUDA(launch_duplicate_key64(s, P, rec, depths, offsets, w.radii, w.cam.grid_x, w.cam.grid_y, ka, vals_alt, w.cap, status), "duplicateWithKeys");
        PROF(4);
        GSR_CUDA(launch_ref_sort_pairs(s, w.cap, ka, vals_alt, kb, point_list, end_bit, w.bin + bl.temp, bl.temp_bytes), "cub sort");
        PROF(5);
        GSR_CUDA(launch_tile_ranges_u64(s, w.cap, n_dev, kb, w.G, ranges), "identifyTileRanges");
      } else {
        PROF(3);
        GSR_CUDA(launch_duplicate_key64(s, P, rec, depths, offsets, w.radii, w.cam.grid_x, w.cam.grid_y, ka, va, w.cap, status), "duplicateWithKeys");
        uint64_t* kout = (passes & 1) ? kb : ka;
        uint64_t* kalt = (passes & 1) ? ka : kb;
        PROF(4);
        GSR_CUDA(launch_sort_pairs_u64(s, w.cap, n_dev, ka, va, kout, point_list, kalt, valt, end_bit, w.bin + bl.temp), "sort");
        PROF(5);
        GSR_CUDA(launch_tile_ranges_u64(s, w.cap, n_dev, kout, w.G, ranges), "identifyTileRanges");
      }
    }
  } else {
    for (int k = 0; k < nv; k++) {
      const ImageLayout il = image_layout(v[k].cam.W, v[k].cam.H);
      GSR_CUDA(cudaMemsetAsync(v[k].img + il.ranges, 0, (size_t)v[k].G * sizeof(uint2), s), "memset ranges");
    }
  }
  PROF(6);
  for (int k = 0; k < nv; k++) {
    const FwdView& w = v[k];
    const ImageLayout il = image_layout(w.cam.W, w.cam.H);
    BlendFwdView& q = bv[(size_t)k];
    q.W = w.cam.W;
    q.H = w.cam.H;
    q.ranges = reinterpret_cast<const uint2*>(w.img + il.ranges);
    q.point_list = reinterpret_cast<const uint32_t*>(w.bin);   // offset 0 for any capacity
    q.rec = reinterpret_cast<const float4*>(w.geom + gl.rec);
    q.depths = reinterpret_cast<const float*>(w.geom + gl.depths);
    q.bg = w.bg;
    q.out_color = w.out_color;
    q.out_depth = w.out_depth;
    q.final_T = reinterpret_cast<float*>(w.img + il.final_T);
    q.n_contrib = reinterpret_cast<uint32_t*>(w.img + il.n_contrib);
  }
  if (ref) {
    for (int k = 0; k < nv; k++)
      GSR_CUDA(launch_ref_blend_forward(s, bv[(size_t)k].W, bv[(size_t)k].H, bv[(size_t)k].ranges, bv[(size_t)k].point_list,
                                        bv[(size_t)k].rec, bv[(size_t)k].depths, bv[(size_t)k].bg, bv[(size_t)k].out_color,
                                        bv[(size_t)k].out_depth, bv[(size_t)k].final_T, bv[(size_t)k].n_contrib),
               "blend forward (reference structure)");
  } else {
    GSR_CUDA(launch_blend_forward(s, bv.data(), nv, (flags & GSR_FLAG_PRECISE) != 0), "blend forward");
  }
  PROF(-1);
  return 0;
}

static FwdView make_fwd_view(int W, int H, const float* view_d, const float* proj_d, const float* campos_d, float tan_fovx,
                             float tan_fovy, float scale_modifier, const float* bg, float* out_color, float* out_depth,
                             int32_t* radii, char* geom, char* img) {
  FwdView w{};
  w.cam = make_camera(W, H, view_d, proj_d, campos_d, tan_fovx, tan_fovy, scale_modifier);
  w.bg = bg;
  w.out_color = out_color;
  w.out_depth = out_depth;
  w.radii = radii;
  w.geom = geom;
  w.img = img;
  w.bin = nullptr;
  w.scratch = nullptr;
  w.cap = 0;
  w.G = w.cam.grid_x * w.cam.grid_y;
  w.end_bit = tile_bits(W, H);
  return w;
}
}  // namespace gsr

extern "C" {

int gsr_forward(void* stream, gsr_alloc_fn geom_alloc, void* geom_user, gsr_alloc_fn binning_alloc,
                void* binning_user, gsr_alloc_fn image_alloc, void* image_user, int P, int D, int M,
                const float* background, int width, int height, const float* means3D,
                const float* shs, const float* colors_precomp, const float* opacities,
                const float* scales, float scale_modifier, const float* rotations,
                const float* cov3D_precomp, const float* viewmatrix, const float* projmatrix,
                const float* cam_pos, float tan_fovx, float tan_fovy, int prefiltered,
                float* out_color, float* out_depth, int32_t* radii, int64_t* num_rendered_host,
                int64_t capacity_hint, uint32_t flags) {
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  if (P < 0 || width <= 0 || height <= 0 || D < 0 || D > 3) return fail(GSR_E_INVALID, "gsr_forward: bad P/width/height/degree");
  if (!geom_alloc || !binning_alloc || !image_alloc || !out_color || !out_depth || !background ||
      !viewmatrix || !projmatrix || !cam_pos || !num_rendered_host)
    return fail(GSR_E_INVALID, "gsr_forward: null argument");
  if (P > 0) {
    if (!means3D || !opacities || !radii) return fail(GSR_E_INVALID, "gsr_forward: null means3D/opacities/radii");
    if ((shs == nullptr) == (colors_precomp == nullptr))
      return fail(GSR_E_INVALID, "gsr_forward: provide exactly one of shs / colors_precomp");
    if (((scales == nullptr) || (rotations == nullptr)) == (cov3D_precomp == nullptr))
      return fail(GSR_E_INVALID, "gsr_forward: provide exactly one of (scales, rotations) / cov3D_precomp");
    if (shs && (D + 1) * (D + 1) > M) return fail(GSR_E_INVALID, "gsr_forward: sh degree needs more coefficients than M");
    if (rotations && (reinterpret_cast<uintptr_t>(rotations) & 15)) return fail(GSR_E_INVALID, "gsr_forward: rotations must be 16-byte aligned");
  }
  *num_rendered_host = 0;
  Pinned& pin = pinned();
  if (!pin.result) return fail(GSR_E_ALLOC, "gsr_forward: cudaMallocHost failed");

  const GeomLayout gl = geom_layout(P, flags);
  const ImageLayout il = image_layout(width, height);
  char* geom = geom_alloc(geom_user, gl.bytes);
  char* img = image_alloc(image_user, il.bytes);
  if (!geom || !img) return fail(GSR_E_ALLOC, "gsr_forward: geometry/image buffer allocation failed");
  FwdView w = make_fwd_view(width, height, viewmatrix, projmatrix, cam_pos, tan_fovx, tan_fovy, scale_modifier, background,
                            out_color, out_depth, radii, geom, img);
  int32_t* status = reinterpret_cast<int32_t*>(geom + gl.status);
  const bool k64 = key64(flags);

  int64_t N = 0;
  const bool async = (flags & GSR_FLAG_ASYNC) != 0;
  const bool ref = refstruct(flags);
  if (ref) {  // the reference learns N on the host mid-pipeline and sizes everything exactly
    if (async) return fail(GSR_E_INVALID, "gsr_forward: GSR_FLAG_REFERENCE excludes GSR_FLAG_ASYNC");
    capacity_hint = 0;
  }
  int64_t* result = async ? num_rendered_host : pin.result;  // [0] = N, [1] = status bits
  if (async && capacity_hint <= 0) return fail(GSR_E_INVALID, "gsr_forward: GSR_FLAG_ASYNC needs a capacity_hint");
  // Instance limit: uint32 positions (the reference's point_offsets are uint32 as well) for the default two-level
  // binning; the literal 64-bit-key paths (GSR_FLAG_BINNING_KEY64 / REFERENCE) keep index arithmetic validated to 2^30.
  const int64_t n_limit = k64 ? (1ll << 30) : GSR_MAX_INSTANCES;
  const bool speculate = capacity_hint > 0 && capacity_hint < n_limit;

  auto alloc_binning = [&](int64_t cap) -> int {
    const BinningLayout bl = binning_layout(cap, width, height, flags);
    w.bin = binning_alloc(binning_user, bl.bytes);
    w.cap = cap;
    if (!w.bin) return fail(GSR_E_ALLOC, "gsr_forward: binning buffer allocation failed");
    return 0;
  };
  if (speculate)
    if (int rc = alloc_binning(capacity_hint)) return rc;
  if (int rc = front_end(s, &w, 1, P, D, M, means3D, shs, colors_precomp, opacities, scales, scale_modifier, rotations,
                         cov3D_precomp, prefiltered, flags, speculate))
    return rc;
  // N lives in status[2..3] (uint64, written by K1); the kernels read its low word
  const uint32_t* n_dev = P > 0 ? reinterpret_cast<const uint32_t*>(status + 2) : nullptr;
  auto fetch_result = [&]() -> int {  // N and the status word to (pinned) host memory, asynchronously
    result[0] = 0;
    result[1] = 0;
    if (P > 0) {
      GSR_CUDA(cudaMemcpyAsync(result, n_dev, 8, cudaMemcpyDeviceToHost, s), "memcpy num_rendered");
      GSR_CUDA(cudaMemcpyAsync(result + 1, status, 8, cudaMemcpyDeviceToHost, s), "memcpy status");
    }
    return 0;
  };

  if (speculate) {
    // Speculative path: queue binning + blend for the hinted capacity BEFORE learning N, so the
    // GPU never idles on the host.  The host then waits for N only (copied right after the depth sort,
    // early in the queue) while the GPU keeps working, or does not wait at all (GSR_FLAG_ASYNC).
    static thread_local cudaEvent_t ev = nullptr;   // one event per host thread, reused (no create / destroy per view)
    if (!async) {
      if (int rc = fetch_result()) return rc;
      if (!ev) GSR_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming), "event create");
      GSR_CUDA(cudaEventRecord(ev, s), "event record");
    }
    if (int rc = bin_and_blend(s, &w, 1, P, flags, false)) return rc;
    if (async) {
      if (int rc = fetch_result()) return rc;   // includes the overflow bit set by the duplicate kernel
      return 0;                                 // caller checks result[0] <= capacity_hint && result[1] == 0 after a sync
    }
    GSR_CUDA(cudaEventSynchronize(ev), "sync num_rendered");
    N = result[0];  // full 64-bit sum: beyond the limit is rejected below
    if ((int32_t)(result[1] & 0xffffffff) != 0) return fail(GSR_E_PREFILTERED, "gsr_forward: point filtered by culling but 'prefiltered' was set");
    if (N >= n_limit) return fail(GSR_E_OVERFLOW, "gsr_forward: too many (tile, Gaussian) instances (limit 2^32 - 65536; 2^30 with GSR_FLAG_BINNING_KEY64 / GSR_FLAG_REFERENCE)");
    if (N > capacity_hint) {  // rare: the hint was too small, redo binning + blend at the exact size
      GSR_CUDA(cudaMemsetAsync(status + 1, 0, 4, s), "clear overflow");
      if (int rc = alloc_binning(N)) return rc;
      if (int rc = bin_and_blend(s, &w, 1, P, flags, true)) return rc;
    }
    *num_rendered_host = N;
    return 0;
  }

  // Exact-size path (what the reference does): one host round trip mid-pipeline.
  if (P > 0) {
    if (int rc = fetch_result()) return rc;
    GSR_CUDA(cudaStreamSynchronize(s), "sync num_rendered");
    N = result[0];  // full 64-bit sum: beyond the limit is rejected below
    if ((int32_t)(result[1] & 0xffffffff) != 0) return fail(GSR_E_PREFILTERED, "gsr_forward: point filtered by culling but 'prefiltered' was set");
    if (N >= n_limit) return fail(GSR_E_OVERFLOW, "gsr_forward: too many (tile, Gaussian) instances (limit 2^32 - 65536; 2^30 with GSR_FLAG_BINNING_KEY64 / GSR_FLAG_REFERENCE)");
  }
  *num_rendered_host = N;
  if (int rc = alloc_binning(N)) return rc;
  return bin_and_blend(s, &w, 1, P, flags, true);
}

int gsr_forward_views(void* stream, int P, int D, int M, const float* means3D, const float* shs,
                      const float* colors_precomp, const float* opacities, const float* scales, float scale_modifier,
                      const float* rotations, const float* cov3D_precomp, int prefiltered,
                      const gsr_view_forward* views_host, int n_views, uint32_t flags) {
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  if (P < 0 || D < 0 || D > 3 || n_views < 0) return fail(GSR_E_INVALID, "gsr_forward_views: bad P/degree/n_views");
  if (n_views == 0) return 0;
  if (!views_host) return fail(GSR_E_INVALID, "gsr_forward_views: null views");
  if (flags & (GSR_FLAG_BINNING_KEY64 | GSR_FLAG_REFERENCE))
    return fail(GSR_E_INVALID, "gsr_forward_views: the batched forward uses the two-level binning (no GSR_FLAG_BINNING_KEY64 / GSR_FLAG_REFERENCE)");
  if (P > 0) {
    if (!means3D || !opacities) return fail(GSR_E_INVALID, "gsr_forward_views: null means3D/opacities");
    if ((shs == nullptr) == (colors_precomp == nullptr))
      return fail(GSR_E_INVALID, "gsr_forward_views: provide exactly one of shs / colors_precomp");
    if (((scales == nullptr) || (rotations == nullptr)) == (cov3D_precomp == nullptr))
      return fail(GSR_E_INVALID, "gsr_forward_views: provide exactly one of (scales, rotations) / cov3D_precomp");
    if (shs && (D + 1) * (D + 1) > M) return fail(GSR_E_INVALID, "gsr_forward_views: sh degree needs more coefficients than M");
    if (rotations && (reinterpret_cast<uintptr_t>(rotations) & 15)) return fail(GSR_E_INVALID, "gsr_forward_views: rotations must be 16-byte aligned");
  }
  const GeomLayout gl = geom_layout(P, flags);
  std::vector<FwdView> vs((size_t)n_views);
  for (int k = 0; k < n_views; k++) {
    const gsr_view_forward& in = views_host[k];
    if (in.width <= 0 || in.height <= 0 || !in.background || !in.viewmatrix || !in.projmatrix || !in.cam_pos ||
        !in.out_color || !in.out_depth || !in.geom_buffer || !in.binning_buffer || !in.image_buffer || !in.result_host ||
        (P > 0 && !in.radii) || in.capacity <= 0 || in.capacity >= GSR_MAX_INSTANCES)
      return fail(GSR_E_INVALID, "gsr_forward_views: bad view descriptor (null pointer, size, or capacity outside (0, GSR_MAX_INSTANCES))");
    if ((reinterpret_cast<uintptr_t>(in.geom_buffer) | reinterpret_cast<uintptr_t>(in.binning_buffer) |
         reinterpret_cast<uintptr_t>(in.image_buffer)) & 255)
      return fail(GSR_E_INVALID, "gsr_forward_views: buffers must be 256-byte aligned");
    vs[(size_t)k] = make_fwd_view(in.width, in.height, in.viewmatrix, in.projmatrix, in.cam_pos, in.tan_fovx, in.tan_fovy,
                                  scale_modifier, in.background, in.out_color, in.out_depth, in.radii, in.geom_buffer,
                                  in.image_buffer);
    if (reinterpret_cast<uintptr_t>(in.backward_scratch) & 15)
      return fail(GSR_E_INVALID, "gsr_forward_views: backward_scratch must be 16-byte aligned");
    vs[(size_t)k].bin = in.binning_buffer;
    vs[(size_t)k].scratch = in.backward_scratch;
    vs[(size_t)k].cap = in.capacity;
    in.result_host[0] = 0;
    in.result_host[1] = 0;
  }
  for (int k0 = 0; k0 < n_views; k0 += GSR_MAX_BATCH) {
    const int nv = std::min(GSR_MAX_BATCH, n_views - k0);
    if (int rc = front_end(s, vs.data() + k0, nv, P, D, M, means3D, shs, colors_precomp, opacities, scales, scale_modifier,
                           rotations, cov3D_precomp, prefiltered, flags, true))
      return rc;
    if (int rc = bin_and_blend(s, vs.data() + k0, nv, P, flags, false)) return rc;
  }
  if (P > 0)
    for (int k = 0; k < n_views; k++) {   // N and the status words (trap, overflow) of every view, asynchronously
      const int32_t* status = status_of(vs[(size_t)k], gl);
      GSR_CUDA(cudaMemcpyAsync(views_host[k].result_host, status + 2, 8, cudaMemcpyDeviceToHost, s), "memcpy num_rendered");
      GSR_CUDA(cudaMemcpyAsync(views_host[k].result_host + 1, status, 8, cudaMemcpyDeviceToHost, s), "memcpy status");
    }
  return 0;
}

int gsr_backward(void* stream, int P, int D, int M, int64_t num_rendered, const float* background,
                 int width, int height, const float* means3D, const float* shs,
                 const float* colors_precomp, const float* scales, float scale_modifier,
                 const float* rotations, const float* cov3D_precomp, const float* viewmatrix,
                 const float* projmatrix, const float* cam_pos, float tan_fovx, float tan_fovy,
                 const int32_t* radii, const char* geom_buffer, const char* binning_buffer,
                 const char* image_buffer, const float* dL_dpix, float* dL_dmean2D,
                 float* dL_dconic, float* dL_dopacity, float* dL_dcolor, float* dL_dmean3D,
                 float* dL_dcov3D, float* dL_dsh, float* dL_dscale, float* dL_drot, char* scratch,
                 size_t scratch_bytes, uint32_t flags) {
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  if (P < 0 || width <= 0 || height <= 0 || num_rendered < 0) return fail(GSR_E_INVALID, "gsr_backward: bad P/width/height/num_rendered");
  if (P == 0) return 0;
  if (!means3D || !radii || !geom_buffer || !binning_buffer || !image_buffer || !dL_dpix || !dL_dmean2D ||
      !dL_dopacity || !dL_dcolor || !dL_dmean3D || !dL_dcov3D || !dL_dscale || !dL_drot || !scratch ||
      !background || !viewmatrix || !projmatrix || !cam_pos)
    return fail(GSR_E_INVALID, "gsr_backward: null argument");
  if (scratch_bytes < gsr_backward_scratch_bytes(P)) return fail(GSR_E_INVALID, "gsr_backward: scratch too small");
  if (shs && !dL_dsh) return fail(GSR_E_INVALID, "gsr_backward: dL_dsh missing");
  if ((reinterpret_cast<uintptr_t>(dL_drot) & 15) || (dL_dconic && (reinterpret_cast<uintptr_t>(dL_dconic) & 15)) ||
      (reinterpret_cast<uintptr_t>(scratch) & 15) || (rotations && (reinterpret_cast<uintptr_t>(rotations) & 15)) ||
      (shs && (reinterpret_cast<uintptr_t>(shs) & 15)) || (dL_dsh && (reinterpret_cast<uintptr_t>(dL_dsh) & 15)))
    return fail(GSR_E_INVALID, "gsr_backward: dL_drot / dL_dconic / dL_dsh / scratch / rotations / shs must be 16-byte aligned");
  const Camera cam = make_camera(width, height, viewmatrix, projmatrix, cam_pos, tan_fovx, tan_fovy, scale_modifier);

  const GeomLayout gl = geom_layout(P, flags);
  const ImageLayout il = image_layout(width, height);
  const BinningLayout bl = binning_layout(num_rendered, width, height, flags);
  const float4* rec = reinterpret_cast<const float4*>(geom_buffer + gl.rec);
  const uint8_t* clamped = reinterpret_cast<const uint8_t*>(geom_buffer + gl.clamped);
  const float* final_T = reinterpret_cast<const float*>(image_buffer + il.final_T);
  const uint32_t* n_contrib = reinterpret_cast<const uint32_t*>(image_buffer + il.n_contrib);
  const uint2* ranges = reinterpret_cast<const uint2*>(image_buffer + il.ranges);
  const uint32_t* point_list = reinterpret_cast<const uint32_t*>(binning_buffer + bl.point_list);
  float* gacc = reinterpret_cast<float*>(scratch);

  PROF(7);
  GSR_CUDA(cudaMemsetAsync(gacc, 0, (size_t)P * 48, s), "memset accumulator");
  PROF(8);
  // always launched: the tile ranges, not num_rendered, bound the work (num_rendered may be unknown
  // to the host after an asynchronous forward)
  if (refstruct(flags)) {
    GSR_CUDA(launch_ref_blend_backward(s, width, height, ranges, point_list, rec, background, final_T, n_contrib,
                                       dL_dpix, gacc), "blend backward (reference structure)");
  } else {
    BlendBwdView bv{};
    bv.W = width; bv.H = height; bv.ranges = ranges; bv.point_list = point_list; bv.rec = rec; bv.bg = background;
    bv.final_T = final_T; bv.n_contrib = n_contrib; bv.dL_dpix = dL_dpix; bv.gacc = gacc;
    GSR_CUDA(launch_blend_backward(s, &bv, 1, (flags & GSR_FLAG_PRECISE) != 0), "blend backward");
  }
  PROF(9);
  GSR_CUDA(launch_geom_backward(s, P, D, M, means3D, radii, shs, clamped, scales, rotations, cov3D_precomp,
                                colors_precomp, cam, rec, gacc, dL_dmean2D, dL_dconic, dL_dopacity, dL_dcolor,
                                dL_dmean3D, dL_dcov3D, dL_dsh, dL_dscale, dL_drot,
                                (flags & GSR_FLAG_ACCUMULATE) != 0), "per-Gaussian backward");
  PROF(-1);
  return 0;
}

int gsr_backward_blend(void* stream, int P, const float* background, int width, int height,
                       const char* geom_buffer, const char* binning_buffer, const char* image_buffer,
                       const float* dL_dpix, char* scratch, size_t scratch_bytes, uint32_t flags) {
  gsr_view_backward v{};
  v.background = background; v.width = width; v.height = height; v.geom_buffer = geom_buffer;
  v.binning_buffer = binning_buffer; v.image_buffer = image_buffer; v.dL_dpix = dL_dpix; v.scratch = scratch;
  v.scratch_bytes = scratch_bytes;
  return gsr_backward_blend_views(stream, P, &v, 1, flags);
}

int gsr_backward_blend_views(void* stream, int P, const gsr_view_backward* views_host, int n_views, uint32_t flags) {
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  if (P < 0 || n_views < 0) return fail(GSR_E_INVALID, "gsr_backward_blend: bad P/n_views");
  if (P == 0 || n_views == 0) return 0;
  if (!views_host) return fail(GSR_E_INVALID, "gsr_backward_blend: null views");
  const GeomLayout gl = geom_layout(P, flags);
  std::vector<BlendBwdView> bv((size_t)n_views);
  std::vector<ZeroRegion> z((size_t)n_views);
  for (int k = 0; k < n_views; k++) {
    const gsr_view_backward& in = views_host[k];
    if (in.width <= 0 || in.height <= 0) return fail(GSR_E_INVALID, "gsr_backward_blend: bad width/height");
    if (!in.background || !in.geom_buffer || !in.binning_buffer || !in.image_buffer || !in.dL_dpix || !in.scratch)
      return fail(GSR_E_INVALID, "gsr_backward_blend: null argument");
    if (in.scratch_bytes < gsr_backward_scratch_bytes(P) || (reinterpret_cast<uintptr_t>(in.scratch) & 15))
      return fail(GSR_E_INVALID, "gsr_backward_blend: scratch too small or misaligned");
    const ImageLayout il = image_layout(in.width, in.height);
    BlendBwdView& q = bv[(size_t)k];
    q.W = in.width;
    q.H = in.height;
    q.ranges = reinterpret_cast<const uint2*>(in.image_buffer + il.ranges);
    q.point_list = reinterpret_cast<const uint32_t*>(in.binning_buffer);   // point_list sits at offset 0 for any capacity
    q.rec = reinterpret_cast<const float4*>(in.geom_buffer + gl.rec);
    q.bg = in.background;
    q.final_T = reinterpret_cast<const float*>(in.image_buffer + il.final_T);
    q.n_contrib = reinterpret_cast<const uint32_t*>(in.image_buffer + il.n_contrib);
    q.dL_dpix = in.dL_dpix;
    q.gacc = reinterpret_cast<float*>(in.scratch);
    z[(size_t)k] = {in.scratch, (size_t)P * 48};
  }
  PROF(7);
  if (flags & GSR_FLAG_SCRATCH_CLEARED) {
    // K1 cleared the accumulators (gsr_view_forward.backward_scratch)
  } else if (n_views == 1)
    GSR_CUDA(cudaMemsetAsync(z[0].ptr, 0, z[0].bytes, s), "memset accumulator");
  else
    GSR_CUDA(launch_zero_regions(s, z.data(), n_views), "clear accumulators");
  PROF(8);
  if (refstruct(flags)) {
    for (int k = 0; k < n_views; k++)
      GSR_CUDA(launch_ref_blend_backward(s, bv[(size_t)k].W, bv[(size_t)k].H, bv[(size_t)k].ranges, bv[(size_t)k].point_list,
                                         bv[(size_t)k].rec, bv[(size_t)k].bg, bv[(size_t)k].final_T, bv[(size_t)k].n_contrib,
                                         bv[(size_t)k].dL_dpix, bv[(size_t)k].gacc), "blend backward (reference structure)");
  } else {
    for (int k0 = 0; k0 < n_views; k0 += GSR_MAX_BATCH)
      GSR_CUDA(launch_blend_backward(s, bv.data() + k0, std::min(GSR_MAX_BATCH, n_views - k0), (flags & GSR_FLAG_PRECISE) != 0),
               "blend backward");
  }
  PROF(-1);
  return 0;
}

int gsr_backward_geom_multi(void* stream, int P, int D, int M, const float* means3D, const float* shs,
                            const float* scales, float scale_modifier, const float* rotations,
                            const gsr_view_grad* views_host, int n_views, float* dL_dopacity,
                            float* dL_dmean3D, float* dL_dsh, float* dL_dscale, float* dL_drot,
                            float* grad_norm_accum, int32_t* visible_count, int32_t* max_radii,
                            uint32_t flags) {
  return gsr_backward_geom_multi_range(stream, P, D, M, means3D, shs, scales, scale_modifier, rotations, views_host,
                                       n_views, dL_dopacity, dL_dmean3D, dL_dsh, dL_dscale, dL_drot, grad_norm_accum,
                                       visible_count, max_radii, flags, 0, P);
}

int gsr_backward_geom_multi_range(void* stream, int P, int D, int M, const float* means3D, const float* shs,
                                  const float* scales, float scale_modifier, const float* rotations,
                                  const gsr_view_grad* views_host, int n_views, float* dL_dopacity,
                                  float* dL_dmean3D, float* dL_dsh, float* dL_dscale, float* dL_drot,
                                  float* grad_norm_accum, int32_t* visible_count, int32_t* max_radii,
                                  uint32_t flags, int g_begin, int g_end) {
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  if (P < 0 || n_views < 0 || D < 0 || D > 3) return fail(GSR_E_INVALID, "gsr_backward_geom_multi: bad P/n_views/degree");
  if (g_begin < 0 || g_end > P || g_begin > g_end || (g_begin & 3))
    return fail(GSR_E_INVALID, "gsr_backward_geom_multi_range: need 0 <= g_begin <= g_end <= P and g_begin % 4 == 0");
  if (P == 0 || n_views == 0 || g_begin == g_end) return 0;
  if (!geom_backward_multi_supported(M) || (D + 1) * (D + 1) > M)
    return fail(GSR_E_INVALID, "gsr_backward_geom_multi: needs M in {1, 4, 16} and (D+1)^2 <= M");
  if (!means3D || !shs || !scales || !rotations || !views_host || !dL_dopacity || !dL_dmean3D || !dL_dsh ||
      !dL_dscale || !dL_drot)
    return fail(GSR_E_INVALID, "gsr_backward_geom_multi: null argument (shs + scales/rotations inputs only)");
  if ((reinterpret_cast<uintptr_t>(dL_drot) & 15) || (reinterpret_cast<uintptr_t>(rotations) & 15) ||
      (reinterpret_cast<uintptr_t>(shs) & 15) || (reinterpret_cast<uintptr_t>(dL_dsh) & 15))
    return fail(GSR_E_INVALID, "gsr_backward_geom_multi: dL_drot / dL_dsh / rotations / shs must be 16-byte aligned");
  std::vector<ViewGrad> vs((size_t)n_views);
  const GeomLayout gl = geom_layout(P, flags);
  const size_t g0 = (size_t)g_begin;
  for (int v = 0; v < n_views; v++) {
    const gsr_view_grad& in = views_host[v];
    if (!in.radii || !in.geom_buffer || !in.scratch || !in.viewmatrix || !in.projmatrix || !in.cam_pos ||
        in.width <= 0 || in.height <= 0 || (reinterpret_cast<uintptr_t>(in.scratch) & 15))
      return fail(GSR_E_INVALID, "gsr_backward_geom_multi: bad view descriptor");
    ViewGrad& o = vs[(size_t)v];
    o.radii = in.radii + g0;
    o.clamped = reinterpret_cast<const uint8_t*>(in.geom_buffer + gl.clamped) + g0;
    o.rec = reinterpret_cast<const float4*>(in.geom_buffer + gl.rec) + 3 * g0;
    o.gacc = reinterpret_cast<const float*>(in.scratch) + 12 * g0;
    o.view = in.viewmatrix;
    o.proj = in.projmatrix;
    o.campos = in.cam_pos;
    o.dL_dmean2D = in.dL_dmean2D ? in.dL_dmean2D + 3 * g0 : nullptr;
    o.focal_x = in.width / (2.0f * in.tan_fovx);
    o.focal_y = in.height / (2.0f * in.tan_fovy);
    o.tan_fovx = in.tan_fovx;
    o.tan_fovy = in.tan_fovy;
    o.W = in.width;
    o.H = in.height;
  }
  PROF(9);
  GSR_CUDA(launch_geom_backward_multi(s, g_end - g_begin, D, M, means3D + 3 * g0, shs + (size_t)M * 3 * g0, scales + 3 * g0,
                                      rotations + 4 * g0, scale_modifier, vs.data(), n_views, dL_dopacity + g0,
                                      dL_dmean3D + 3 * g0, dL_dsh + (size_t)M * 3 * g0, dL_dscale + 3 * g0, dL_drot + 4 * g0,
                                      grad_norm_accum ? grad_norm_accum + g0 : nullptr,
                                      visible_count ? visible_count + g0 : nullptr, max_radii ? max_radii + g0 : nullptr,
                                      (flags & GSR_FLAG_ACCUMULATE) != 0),
           "multi-view per-Gaussian backward");
  PROF(-1);
  return 0;
}

int gsr_nvls_all_reduce(void* stream, void* multicast_ptr, size_t off_f32, size_t n_f32, size_t off_add_s32,
                        size_t n_add_s32, size_t off_max_s32, size_t n_max_s32, int rank, int world, int blocks,
                        size_t sparse_first_f32, size_t sparse_rows, int sparse_row_f32) {
  if (!multicast_ptr || world < 1 || rank < 0 || rank >= world || (n_f32 & 3) || (off_f32 & 15) || (off_add_s32 & 3) ||
      (off_max_s32 & 3) || (reinterpret_cast<uintptr_t>(multicast_ptr) & 15))
    return fail(GSR_E_INVALID, "gsr_nvls_all_reduce: bad argument (alignment / rank / multicast pointer)");
  GSR_CUDA(launch_nvls_allreduce(reinterpret_cast<cudaStream_t>(stream), reinterpret_cast<char*>(multicast_ptr), off_f32,
                                 n_f32, off_add_s32, n_add_s32, off_max_s32, n_max_s32, rank, world, blocks,
                                 sparse_first_f32, sparse_rows, sparse_row_f32),
           "nvls all-reduce");
  return 0;
}

int gsr_nvls_all_reduce_plan(void* stream, void* multicast_ptr, const gsr_nvls_plan* plan, int rank, int world, int blocks) {
  if (!multicast_ptr || !plan || world < 1 || rank < 0 || rank >= world || (reinterpret_cast<uintptr_t>(multicast_ptr) & 15) ||
      plan->n_dense < 0 || plan->n_dense > 6 || (plan->rows_off & 15) || (plan->rows_count_off & 3) ||
      (plan->add_s32_off & 3) || (plan->max_s32_off & 3) || plan->row_f32 < 0)
    return fail(GSR_E_INVALID, "gsr_nvls_all_reduce_plan: bad argument (alignment / rank / multicast pointer)");
  for (int d = 0; d < plan->n_dense; d++)
    if ((plan->dense_off[d] & 15) || (plan->dense_n_f32[d] & 3))
      return fail(GSR_E_INVALID, "gsr_nvls_all_reduce_plan: dense segments must be 16-byte aligned multiples of 4 floats");
  GSR_CUDA(launch_nvls_allreduce_plan(reinterpret_cast<cudaStream_t>(stream), reinterpret_cast<char*>(multicast_ptr), *plan,
                                      rank, world, blocks), "nvls all-reduce (plan)");
  return 0;
}

int gsr_p2p_all_reduce_plan(void* stream, void* local_ptr, void* peer_ptr, const gsr_nvls_plan* plan, int rank, int blocks) {
  if (!local_ptr || !peer_ptr || local_ptr == peer_ptr || !plan || rank < 0 || rank > 1 ||
      ((reinterpret_cast<uintptr_t>(local_ptr) | reinterpret_cast<uintptr_t>(peer_ptr)) & 15) ||
      plan->n_dense < 0 || plan->n_dense > 6 || (plan->rows_off & 15) || (plan->rows_count_off & 3) ||
      (plan->add_s32_off & 3) || (plan->max_s32_off & 3) || plan->row_f32 < 0)
    return fail(GSR_E_INVALID, "gsr_p2p_all_reduce_plan: bad argument (alignment / rank / pointers)");
  for (int d = 0; d < plan->n_dense; d++)
    if ((plan->dense_off[d] & 15) || (plan->dense_n_f32[d] & 3))
      return fail(GSR_E_INVALID, "gsr_p2p_all_reduce_plan: dense segments must be 16-byte aligned multiples of 4 floats");
  GSR_CUDA(launch_p2p_allreduce_plan(reinterpret_cast<cudaStream_t>(stream), reinterpret_cast<char*>(local_ptr),
                                     reinterpret_cast<char*>(peer_ptr), *plan, rank, blocks), "p2p all-reduce (plan)");
  return 0;
}

int gsr_accumulate_view_stats(void* stream, int P, const int32_t* radii, const float* dL_dmean2D,
                              float* grad_norm_accum, int32_t* visible_count, int32_t* max_radii) {
  if (P < 0 || (P > 0 && (!radii || (grad_norm_accum && !dL_dmean2D))))
    return fail(GSR_E_INVALID, "gsr_accumulate_view_stats: bad argument");
  GSR_CUDA(launch_view_stats(reinterpret_cast<cudaStream_t>(stream), P, radii, dL_dmean2D, grad_norm_accum,
                             visible_count, max_radii), "view stats");
  return 0;
}

int gsr_mark_visible(void* stream, int P, const float* means3D, const float* viewmatrix,
                     const float* projmatrix, uint8_t* present) {
  (void)projmatrix;  // the reference's frustum test only uses view-space z (x/y test disabled)
  if (P < 0 || (P > 0 && (!means3D || !viewmatrix || !present))) return fail(GSR_E_INVALID, "gsr_mark_visible: bad argument");
  GSR_CUDA(launch_mark_visible(reinterpret_cast<cudaStream_t>(stream), P, means3D, viewmatrix, present), "markVisible");
  return 0;
}

int gsr_sort_pairs_u64(void* stream, int64_t n, const uint64_t* keys_in, const uint32_t* vals_in,
                       uint64_t* keys_out, uint32_t* vals_out, int end_bit, char* temp, size_t temp_bytes) {
  if (n < 0 || end_bit < 1 || end_bit > 64) return fail(GSR_E_INVALID, "gsr_sort_pairs_u64: bad argument");
  if (n >= GSR_MAX_INSTANCES) return fail(GSR_E_OVERFLOW, "gsr_sort_pairs_u64: n >= 2^32 - 65536");
  if (n == 0) return 0;
  if (!keys_in || !keys_out || !vals_out || !temp || temp_bytes < gsr_sort_temp_bytes(n, 8, end_bit))
    return fail(GSR_E_INVALID, "gsr_sort_pairs_u64: null argument or temp too small");
  const size_t t0 = sort_temp_bytes(n, 8, end_bit);
  uint64_t* kalt = reinterpret_cast<uint64_t*>(temp + t0);
  uint32_t* valt = reinterpret_cast<uint32_t*>(temp + t0 + align_up((size_t)n * 8));
  GSR_CUDA(launch_sort_pairs_u64(reinterpret_cast<cudaStream_t>(stream), n, nullptr, keys_in, vals_in, keys_out, vals_out,
                                 kalt, valt, end_bit, temp), "sort_pairs_u64");
  return 0;
}

int gsr_sort_pairs_u32(void* stream, int64_t n, const uint32_t* keys_in, const uint32_t* vals_in,
                       uint32_t* keys_out, uint32_t* vals_out, int end_bit, char* temp, size_t temp_bytes) {
  if (n < 0 || end_bit < 1 || end_bit > 32) return fail(GSR_E_INVALID, "gsr_sort_pairs_u32: bad argument");
  if (n >= GSR_MAX_INSTANCES) return fail(GSR_E_OVERFLOW, "gsr_sort_pairs_u32: n >= 2^32 - 65536");
  if (n == 0) return 0;
  if (!keys_in || !keys_out || !vals_out || !temp || temp_bytes < gsr_sort_temp_bytes(n, 4, end_bit))
    return fail(GSR_E_INVALID, "gsr_sort_pairs_u32: null argument or temp too small");
  const size_t t0 = sort_temp_bytes(n, 4, end_bit);
  uint32_t* kalt = reinterpret_cast<uint32_t*>(temp + t0);
  uint32_t* valt = reinterpret_cast<uint32_t*>(temp + t0 + align_up((size_t)n * 4));
  GSR_CUDA(launch_sort_pairs_u32(reinterpret_cast<cudaStream_t>(stream), n, nullptr, keys_in, vals_in, keys_out, vals_out,
                                 kalt, valt, end_bit, temp), "sort_pairs_u32");
  return 0;
}

int gsr_inclusive_scan_u32(void* stream, int64_t n, const uint32_t* in, const uint32_t* gather,
                           uint32_t* out, char* temp, size_t temp_bytes) {
  if (n < 0) return fail(GSR_E_INVALID, "gsr_inclusive_scan_u32: bad n");
  if (n == 0) return 0;
  if (!in || !out || !temp || temp_bytes < scan_temp_bytes(n)) return fail(GSR_E_INVALID, "gsr_inclusive_scan_u32: null argument or temp too small");
  GSR_CUDA(launch_inclusive_scan(reinterpret_cast<cudaStream_t>(stream), n, in, gather, out, temp), "inclusive_scan");
  return 0;
}

}  // extern "C"
