// image_io.cu -- the output side of the inference loops: float (C,H,W) render -> 8-bit (H,W,3) image
// on the device, so that the device-to-host copy an asynchronous PNG writer drains carries 3 bytes per pixel, not 12.
//
// Replaces the tensor half of `torchvision.utils.save_image(t, path)` as the reference calls it after every render()
// (gs-simp/render.py:36-39, render_depth.py:39, gen_seq.py:45,53,55, vis_render.py:48-51; torchvision is a third-party
// dependency of the reference, unpinned, gs-simp/environment.yml): make_grid of a single image (a 1-channel image is
// repeated to 3 channels) followed by
//     ndarr = grid.mul(255).add_(0.5).clamp_(0, 255).permute(1, 2, 0).to("cpu", torch.uint8)
// i.e. per element u8 = trunc(clamp(fl(fl(x * 255) + 0.5), 0, 255)) -- two fp32 roundings (no fma), then truncation.
// Integer output: bit-exact against the numpy oracle.  NaN maps to 0.
// Optional affine pre-map x' = (x - lo) * inv_range covers scene/helpers.py:159-162 normalize_0_to_1 (render_depth.py:39)
// with lo / inv_range read from the device (no host sync for the min / max).
#include "common.cuh"

namespace gsr {

namespace {

__device__ __forceinline__ unsigned int quant_u8(float x) {
  float v = ADD(MUL(x, 255.0f), 0.5f);
  v = fminf(fmaxf(v, 0.0f), 255.0f);     // fmaxf(NaN, 0) = 0
  return __float2uint_rz(v);
}

// one thread per 4 horizontally adjacent pixels: three coalesced float4 plane loads, 12 bytes out (three 32-bit stores)
__global__ void __launch_bounds__(256)
quantize_rgb8_kernel(int C, size_t HW, const float* __restrict__ in, const float* __restrict__ affine,
                     uint8_t* __restrict__ out) {
  const size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t p0 = 4 * q;
  if (p0 >= HW) return;
  float lo = 0.0f, sc = 1.0f;
  if (affine) { lo = __ldg(affine); sc = __ldg(affine + 1); }
  const bool vec = (p0 + 3 < HW) && ((HW & 3) == 0) && ((reinterpret_cast<uintptr_t>(in) & 15) == 0);
  float v[3][4];
#pragma unroll
  for (int c = 0; c < 3; c++) {
    const float* plane = in + (size_t)(C == 1 ? 0 : c) * HW;
    if (vec) {
      const float4 t = __ldg(reinterpret_cast<const float4*>(plane + p0));
      v[c][0] = t.x; v[c][1] = t.y; v[c][2] = t.z; v[c][3] = t.w;
    } else {
#pragma unroll
      for (int j = 0; j < 4; j++) v[c][j] = p0 + j < HW ? __ldg(plane + p0 + j) : 0.0f;
    }
  }
  unsigned int b[12];
#pragma unroll
  for (int j = 0; j < 4; j++)
#pragma unroll
    for (int c = 0; c < 3; c++) {
      float x = v[c][j];
      if (affine) x = MUL(SUB(x, lo), sc);
      b[3 * j + c] = quant_u8(x);
    }
  uint8_t* o = out + 3 * p0;
  if (p0 + 3 < HW && (reinterpret_cast<uintptr_t>(out) & 3) == 0) {
    unsigned int* o32 = reinterpret_cast<unsigned int*>(o);       // 12 * q bytes: 4-byte aligned
#pragma unroll
    for (int k = 0; k < 3; k++) o32[k] = b[4 * k] | (b[4 * k + 1] << 8) | (b[4 * k + 2] << 16) | (b[4 * k + 3] << 24);
  } else {
    for (int j = 0; j < 12 && p0 + j / 3 < HW; j++) o[j] = (uint8_t)b[j];
  }
}

}  // namespace

cudaError_t launch_quantize_rgb8(cudaStream_t s, int C, int H, int W, const float* in, const float* affine, uint8_t* out) {
  const size_t HW = (size_t)H * (size_t)W;
  if (HW == 0) return cudaSuccess;
  const size_t quads = (HW + 3) / 4;
  quantize_rgb8_kernel<<<(unsigned int)((quads + 255) / 256), 256, 0, s>>>(C, HW, in, affine, out);
  count_launch();
  return cudaGetLastError();
}

}  // namespace gsr
