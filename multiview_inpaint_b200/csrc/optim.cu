// optim.cu -- the two per-Gaussian stages either side of the rasterizer in a training step:
//
//  * parameter activations, forward and backward, ONE kernel each -- what GaussianModel's getters do in five
//    torch kernels per render() call and autograd undoes in a dozen (gs-simp/scene/gaussian_model.py:33-41,95-115):
//        scales    = exp(_scaling)                torch.exp
//        rotations = normalize(_rotation)         torch.nn.functional.normalize: v / max(||v||_2, 1e-12)
//        opacity   = sigmoid(_opacity)            torch.sigmoid
//    In a multi-view step the activated values are shared by all views, so they are computed once per step
//    (not once per view as the reference's getters do) and the chain rule runs once on the summed gradients,
//    in place in the gradient arena.
//
//  * Adam over the flat parameter arena, ONE launch for all six parameter groups of
//    gs-simp/scene/gaussian_model.py:154-165 (torch.optim.Adam(l, lr=0.0, eps=1e-15), betas (0.9, 0.999), no
//    weight decay, no amsgrad), i.e. torch/optim/adam.py `_single_tensor_adam`:
//        m = lerp(m, g, 1-b1);  v = b2 v + (1-b2) g g;  p -= (lr / (1-b1^t)) * m / (sqrt(v)/sqrt(1-b2^t) + eps)
//    A segment may be row-structured: the (P,M,3) SH tensor keeps f_dc (row floats [0,3), lr = feature_lr) and
//    f_rest (row floats [3,3M), lr = feature_lr/20) in ONE tensor, which is what makes the reference's
//    per-view torch.cat((f_dc, f_rest)) (gaussian_model.py:107-111) disappear.
//
// Both are pure HBM streams: 64 B per Gaussian for an activation pass, 28 B per parameter for Adam
// (read p, g, m, v; write p, m, v) against ~76 B per parameter for the foreach implementation's passes.
#include <math.h>

#include "common.cuh"

namespace gsr {

namespace {

// ---------------------------------------------------------------- activations
__global__ void __launch_bounds__(256)
activate_forward_kernel(int P, const float* __restrict__ raw_scale, const float* __restrict__ raw_rot,
                        const float* __restrict__ raw_opacity, float* __restrict__ scale,
                        float* __restrict__ rot, float* __restrict__ opacity) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P) return;
  const size_t i3 = 3 * (size_t)i;
  if (raw_scale) {
#pragma unroll
    for (int k = 0; k < 3; k++) scale[i3 + k] = expf(__ldg(raw_scale + i3 + k));
  }
  if (raw_rot) {
    const float4 q = __ldg(reinterpret_cast<const float4*>(raw_rot) + i);
    // ||q||_2 then clamp_min(eps): F.normalize(p=2, dim=1, eps=1e-12)
    const float n = fmaxf(SQRT(ADD(ADD(ADD(MUL(q.x, q.x), MUL(q.y, q.y)), MUL(q.z, q.z)), MUL(q.w, q.w))), 1e-12f);
    reinterpret_cast<float4*>(rot)[i] = make_float4(DIV(q.x, n), DIV(q.y, n), DIV(q.z, n), DIV(q.w, n));
  }
  if (raw_opacity) opacity[i] = DIV(1.0f, ADD(1.0f, expf(-__ldg(raw_opacity + i))));
}

// In place: g_* hold dL/d(activated) on entry and dL/d(raw) on exit.
__global__ void __launch_bounds__(256)
activate_backward_kernel(int P, const float* __restrict__ raw_scale, const float* __restrict__ raw_rot,
                         const float* __restrict__ raw_opacity, float* __restrict__ g_scale,
                         float* __restrict__ g_rot, float* __restrict__ g_opacity) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P) return;
  const size_t i3 = 3 * (size_t)i;
  if (raw_scale) {
#pragma unroll
    for (int k = 0; k < 3; k++) g_scale[i3 + k] = g_scale[i3 + k] * expf(__ldg(raw_scale + i3 + k));   // grad * result
  }
  if (raw_rot) {
    const float4 q = __ldg(reinterpret_cast<const float4*>(raw_rot) + i);
    const float4 g = reinterpret_cast<const float4*>(g_rot)[i];
    const float norm = SQRT(ADD(ADD(ADD(MUL(q.x, q.x), MUL(q.y, q.y)), MUL(q.z, q.z)), MUL(q.w, q.w)));
    float4 o;
    if (norm >= 1e-12f) {   // clamp_min passes the gradient where norm >= eps
      const float inv = 1.0f / norm;
      const float yx = q.x * inv, yy = q.y * inv, yz = q.z * inv, yw = q.w * inv;
      const float d = yx * g.x + yy * g.y + yz * g.z + yw * g.w;
      o = make_float4((g.x - yx * d) * inv, (g.y - yy * d) * inv, (g.z - yz * d) * inv, (g.w - yw * d) * inv);
    } else {
      o = make_float4(g.x * 1e12f, g.y * 1e12f, g.z * 1e12f, g.w * 1e12f);
    }
    reinterpret_cast<float4*>(g_rot)[i] = o;
  }
  if (raw_opacity) {
    const float y = DIV(1.0f, ADD(1.0f, expf(-__ldg(raw_opacity + i))));
    g_opacity[i] = g_opacity[i] * (1.0f - y) * y;   // sigmoid_backward: grad * (1 - y) * y
  }
}

// ---------------------------------------------------------------- Adam
constexpr int ADAM_MAX_SEGS = 8;
constexpr int ADAM_THREADS = 256;
constexpr int ADAM_PER_BLOCK = ADAM_THREADS * 4 * 2;   // floats per CTA: two float4 per thread

struct AdamSeg {
  float* p;
  const float* g;
  float* m;
  float* v;
  unsigned long long n;
  unsigned int first_block;     // prefix sum of CTAs over the segments
  int row_len, row_split;       // row_len == 0: one learning rate
  float step_size, step_size_rest;   // lr / (1 - beta1^t), already negated
  int vec;                      // all four pointers 16-byte aligned
};
struct AdamArgs {
  AdamSeg seg[ADAM_MAX_SEGS];
  int n_segs;
  float beta1, beta2, one_minus_beta1, one_minus_beta2, inv_bc2_sqrt, eps;
};

__device__ __forceinline__ void adam_one(float& p, float g, float& m, float& v, float neg_step, const AdamArgs& A) {
  // exp_avg.lerp_(grad, 1 - beta1): weight < 0.5 -> start + weight * (end - start)
  m = fmaf(A.one_minus_beta1, g - m, m);
  // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, value = 1 - beta2)
  v = fmaf(A.one_minus_beta2 * g, g, v * A.beta2);
  // denom = (exp_avg_sq.sqrt() / bias_correction2_sqrt).add_(eps); param.addcdiv_(exp_avg, denom, value = -step_size)
  const float denom = __fsqrt_rn(v) * A.inv_bc2_sqrt + A.eps;
  p = fmaf(neg_step, __fdiv_rn(m, denom), p);
}

__global__ void __launch_bounds__(ADAM_THREADS)
adam_kernel(const AdamArgs A) {
  // which segment does this CTA belong to?  (<= 8 segments: a short scan of uniform values)
  int s = 0;
#pragma unroll
  for (int k = 1; k < ADAM_MAX_SEGS; k++)
    if (k < A.n_segs && blockIdx.x >= A.seg[k].first_block) s = k;
  const AdamSeg& S = A.seg[s];
  const unsigned long long base = (unsigned long long)(blockIdx.x - S.first_block) * ADAM_PER_BLOCK;
  if (S.vec) {
#pragma unroll
    for (int r = 0; r < 2; r++) {
      const unsigned long long e = base + ((unsigned long long)r * ADAM_THREADS + threadIdx.x) * 4;
      if (e + 3 < S.n) {
        float4 p = *reinterpret_cast<float4*>(S.p + e);
        const float4 g = __ldcs(reinterpret_cast<const float4*>(S.g + e));
        float4 m = *reinterpret_cast<float4*>(S.m + e);
        float4 v = *reinterpret_cast<float4*>(S.v + e);
        float st[4] = {S.step_size, S.step_size, S.step_size, S.step_size};
        if (S.row_len > 0) {
          const int c = (int)(e % (unsigned long long)S.row_len);
#pragma unroll
          for (int j = 0; j < 4; j++) {
            int cj = c + j;
            if (cj >= S.row_len) cj -= S.row_len;   // row_len >= 4 is checked on the host for vector segments
            st[j] = cj < S.row_split ? S.step_size : S.step_size_rest;
          }
        }
        adam_one(p.x, g.x, m.x, v.x, st[0], A);
        adam_one(p.y, g.y, m.y, v.y, st[1], A);
        adam_one(p.z, g.z, m.z, v.z, st[2], A);
        adam_one(p.w, g.w, m.w, v.w, st[3], A);
        *reinterpret_cast<float4*>(S.p + e) = p;
        *reinterpret_cast<float4*>(S.m + e) = m;
        *reinterpret_cast<float4*>(S.v + e) = v;
      } else {
        for (unsigned long long k = e; k < S.n && k < e + 4; k++) {
          float p = S.p[k], m = S.m[k], v = S.v[k];
          const float st = (S.row_len > 0 && (int)(k % (unsigned long long)S.row_len) >= S.row_split)
                               ? S.step_size_rest : S.step_size;
          adam_one(p, S.g[k], m, v, st, A);
          S.p[k] = p; S.m[k] = m; S.v[k] = v;
        }
      }
    }
  } else {
#pragma unroll
    for (int r = 0; r < 8; r++) {
      const unsigned long long k = base + (unsigned long long)r * ADAM_THREADS + threadIdx.x;
      if (k < S.n) {
        float p = S.p[k], m = S.m[k], v = S.v[k];
        const float st = (S.row_len > 0 && (int)(k % (unsigned long long)S.row_len) >= S.row_split)
                             ? S.step_size_rest : S.step_size;
        adam_one(p, S.g[k], m, v, st, A);
        S.p[k] = p; S.m[k] = m; S.v[k] = v;
      }
    }
  }
}

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

}  // namespace

cudaError_t launch_activate_forward(cudaStream_t s, int P, const float* raw_scale, const float* raw_rot,
                                    const float* raw_opacity, float* scale, float* rot, float* opacity) {
  if (P == 0) return cudaSuccess;
  activate_forward_kernel<<<cdiv(P, 256), 256, 0, s>>>(P, raw_scale, raw_rot, raw_opacity, scale, rot, opacity);
  count_launch();
  return cudaGetLastError();
}

cudaError_t launch_activate_backward(cudaStream_t s, int P, const float* raw_scale, const float* raw_rot,
                                     const float* raw_opacity, float* g_scale, float* g_rot, float* g_opacity) {
  if (P == 0) return cudaSuccess;
  activate_backward_kernel<<<cdiv(P, 256), 256, 0, s>>>(P, raw_scale, raw_rot, raw_opacity, g_scale, g_rot, g_opacity);
  count_launch();
  return cudaGetLastError();
}

cudaError_t launch_adam(cudaStream_t s, const gsr_adam_segment* segs, int n_segs, int64_t step, double beta1,
                        double beta2, double eps) {
  AdamArgs A;
  A.n_segs = 0;
  // torch keeps these as python doubles and rounds them to the tensors' dtype when the op is applied
  A.beta1 = (float)beta1;
  A.beta2 = (float)beta2;
  A.one_minus_beta1 = (float)(1.0 - beta1);
  A.one_minus_beta2 = (float)(1.0 - beta2);
  const double bc1 = 1.0 - pow(beta1, (double)step);
  const double bc2 = 1.0 - pow(beta2, (double)step);
  A.inv_bc2_sqrt = (float)(1.0 / sqrt(bc2));
  A.eps = (float)eps;
  unsigned long long blocks = 0;
  for (int k = 0; k < n_segs; k++) {
    const gsr_adam_segment& in = segs[k];
    if (in.n == 0) continue;
    AdamSeg& S = A.seg[A.n_segs++];
    S.p = in.param; S.g = in.grad; S.m = in.exp_avg; S.v = in.exp_avg_sq;
    S.n = in.n;
    S.first_block = (unsigned int)blocks;
    S.row_len = in.row_len;
    S.row_split = in.row_split;
    S.step_size = (float)(-in.lr / bc1);
    S.step_size_rest = (float)(-in.lr_rest / bc1);
    S.vec = aligned16(in.param) && aligned16(in.grad) && aligned16(in.exp_avg) && aligned16(in.exp_avg_sq) &&
            (in.row_len == 0 || in.row_len >= 4);
    blocks += (in.n + ADAM_PER_BLOCK - 1) / ADAM_PER_BLOCK;
    if (blocks >= (1ull << 31)) return cudaErrorInvalidValue;
  }
  if (A.n_segs == 0) return cudaSuccess;
  adam_kernel<<<(unsigned int)blocks, ADAM_THREADS, 0, s>>>(A);
  count_launch();
  return cudaGetLastError();
}

}  // namespace gsr
