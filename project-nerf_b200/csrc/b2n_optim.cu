// Per-step parameter update of the training loops as two passes over the parameters:
//   run.py:611-630 / :1112-1120, :1167-1178 / :1840-1859, :1940-1949 do, per step and per parameter tensor,
//   TV loss forward + backward over the flat hash tables (slice, sub, abs, mean, sign, ...), GradScaler.unscale_,
//   clip_grad_norm_ (norm pass + scale pass) and foreach-AdamW (7 more passes): ~100 launches and 1.25 ms of a 3.65 ms
//   Dual-Hash (C5) step on B200 (profiles: launch list), against 2.4 ms in the ray-marching kernels.
// pass 1  k_opt_prepare: g <- g / loss_scale + d/dp [ tv_weight * mean |p[i+1] - p[i]| ];  norm2[group] += sum g^2
// pass 2  k_opt_adamw  : g <- g * min(1, max_norm / (norm + 1e-6));  AdamW (torch.optim.AdamW arithmetic, fp32);
//                        skipped entirely when GradScaler found an inf.
// One launch per pass for up to B2N_OPT_MAX_TENSORS tensors (blockIdx.y = tensor); HBM-bound: 8 B/param in pass 1
// (16 with TV), 28 B/param in pass 2.
#include "b2n_common.cuh"

namespace b2n {
namespace opt {

struct Args {
  int n;
  b2n_opt_tensor t[B2N_OPT_MAX_TENSORS];
  const float* grad_scale;   // device scalar or nullptr: gradients are still multiplied by it
  const float* found_inf;    // device scalar or nullptr: non-zero -> no update this step
  const float* step_dev;     // device scalar or nullptr: the step count t of the bias corrections 1 - beta^t (a skipped
                             // step must not advance it, which only the device knows); nullptr: the descriptors' values
  float* norm2;              // device float[B2N_OPT_MAX_GROUPS]
  float max_norm[B2N_OPT_MAX_GROUPS];   // < 0: group not clipped
};

__device__ __forceinline__ float sgn(float x) { return (float)(x > 0.f) - (float)(x < 0.f); }

__global__ void __launch_bounds__(256) k_opt_prepare(const Args a) {
  const b2n_opt_tensor t = a.t[blockIdx.y];
  const float inv_scale = a.grad_scale ? 1.f / __ldg(a.grad_scale) : 1.f;
  const bool rewrite = t.tv_scale != 0.f || a.grad_scale != nullptr;
  float acc = 0.f;
  // 16-byte path (descriptor flag `reserved` = all pointers 16-byte aligned): 4 elements per thread and iteration -- with
  // scalar accesses the big tables ran at 1.8 TB/s (too few bytes in flight per thread: ncu, profiles/r1h)
  const int64_t n4 = t.reserved ? (t.n >> 2) : 0;
  for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < n4; k += (int64_t)gridDim.x * blockDim.x) {
    float4 g4 = reinterpret_cast<const float4*>(t.g)[k];
    float g[4] = {g4.x * inv_scale, g4.y * inv_scale, g4.z * inv_scale, g4.w * inv_scale};
    if (t.tv_scale != 0.f) {
      const float4 p4 = reinterpret_cast<const float4*>(t.p)[k];
      const int64_t i0 = 4 * k;
      const float pl = i0 > 0 ? t.p[i0 - 1] : 0.f, pr = i0 + 4 < t.n ? t.p[i0 + 4] : 0.f;
      const float s01 = sgn(p4.y - p4.x), s12 = sgn(p4.z - p4.y), s23 = sgn(p4.w - p4.z);
      const float sl = i0 > 0 ? sgn(p4.x - pl) : 0.f, sr = i0 + 4 < t.n ? sgn(pr - p4.w) : 0.f;
      g[0] = fmaf(t.tv_scale, sl - s01, g[0]);
      g[1] = fmaf(t.tv_scale, s01 - s12, g[1]);
      g[2] = fmaf(t.tv_scale, s12 - s23, g[2]);
      g[3] = fmaf(t.tv_scale, s23 - sr, g[3]);
    }
    if (rewrite) reinterpret_cast<float4*>(t.g)[k] = make_float4(g[0], g[1], g[2], g[3]);
    acc = fmaf(g[0], g[0], fmaf(g[1], g[1], fmaf(g[2], g[2], fmaf(g[3], g[3], acc))));
  }
  for (int64_t i = 4 * n4 + blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < t.n; i += (int64_t)gridDim.x * blockDim.x) {
    float g = t.g[i] * inv_scale;
    if (t.tv_scale != 0.f) {
      const float pc = t.p[i];
      float d = 0.f;
      if (i > 0) d += sgn(pc - t.p[i - 1]);
      if (i + 1 < t.n) d -= sgn(t.p[i + 1] - pc);
      g = fmaf(t.tv_scale, d, g);
    }
    if (rewrite) t.g[i] = g;
    acc = fmaf(g, g, acc);
  }
  if (t.clip_group < 0) return;
  acc = warp_sum(acc);
  __shared__ float part[8];
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;
    for (int w = 0; w < 8; ++w) s += part[w];
    if (s != 0.f) atomicAdd(a.norm2 + t.clip_group, s);
  }
}

// unscaled: pass 1 already divided the gradients by the loss scale (it ran); otherwise do it here
__global__ void __launch_bounds__(256) k_opt_adamw(const Args a, int unscaled) {
  if (a.found_inf && __ldg(a.found_inf) != 0.f) return;
  const b2n_opt_tensor t = a.t[blockIdx.y];
  float coef = (a.grad_scale && !unscaled) ? 1.f / __ldg(a.grad_scale) : 1.f;
  if (t.clip_group >= 0 && a.max_norm[t.clip_group] >= 0.f) {
    const float c = a.max_norm[t.clip_group] / (sqrtf(__ldg(a.norm2 + t.clip_group)) + 1e-6f);
    coef *= fminf(c, 1.f);
  }
  const float decay = 1.f - t.lr * t.weight_decay;
  float bc1 = t.bias_corr1, bc2 = t.bias_corr2;
  if (a.step_dev) {
    const float st = __ldg(a.step_dev);
    bc1 = 1.f - powf(t.beta1, st), bc2 = 1.f - powf(t.beta2, st);
  }
  const float step_size = t.lr / bc1;
  const float inv_sqrt_bc2 = 1.f / sqrtf(bc2);
  const float omb1 = 1.f - t.beta1, omb2 = 1.f - t.beta2;
  const int64_t n4 = t.reserved ? (t.n >> 2) : 0;
  for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < n4; k += (int64_t)gridDim.x * blockDim.x) {
    const float4 g4 = reinterpret_cast<const float4*>(t.g)[k];
    float4 p4 = reinterpret_cast<float4*>(t.p)[k], m4 = reinterpret_cast<float4*>(t.m)[k], v4 = reinterpret_cast<float4*>(t.v)[k];
    float gg[4] = {g4.x * coef, g4.y * coef, g4.z * coef, g4.w * coef};
    float pp[4] = {p4.x, p4.y, p4.z, p4.w}, mm[4] = {m4.x, m4.y, m4.z, m4.w}, vv[4] = {v4.x, v4.y, v4.z, v4.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      mm[j] = fmaf(gg[j] - mm[j], omb1, mm[j]);
      vv[j] = fmaf(gg[j] * gg[j], omb2, vv[j] * t.beta2);
      pp[j] = pp[j] * decay - step_size * (mm[j] / (sqrtf(vv[j]) * inv_sqrt_bc2 + t.eps));
    }
    reinterpret_cast<float4*>(t.p)[k] = make_float4(pp[0], pp[1], pp[2], pp[3]);
    reinterpret_cast<float4*>(t.m)[k] = make_float4(mm[0], mm[1], mm[2], mm[3]);
    reinterpret_cast<float4*>(t.v)[k] = make_float4(vv[0], vv[1], vv[2], vv[3]);
  }
  for (int64_t i = 4 * n4 + blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < t.n; i += (int64_t)gridDim.x * blockDim.x) {
    const float g = t.g[i] * coef;
    float p = t.p[i] * decay;
    float m = t.m[i], v = t.v[i];
    m = fmaf(g - m, 1.f - t.beta1, m);                   // exp_avg.lerp_(grad, 1 - beta1)
    v = fmaf(g * g, 1.f - t.beta2, v * t.beta2);         // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, 1 - beta2)
    const float denom = sqrtf(v) * inv_sqrt_bc2 + t.eps;
    p -= step_size * (m / denom);
    t.p[i] = p, t.m[i] = m, t.v[i] = v;
  }
}

static int fill(Args* a, const b2n_opt_tensor* tensors, int n) {
  a->n = n;
  for (int i = 0; i < n; ++i) {
    const b2n_opt_tensor& t = tensors[i];
    if (!t.p || !t.g || t.n <= 0 || t.clip_group >= B2N_OPT_MAX_GROUPS) return -1;
    a->t[i] = t;
    const uintptr_t bits = (uintptr_t)t.p | (uintptr_t)t.g | (uintptr_t)t.m | (uintptr_t)t.v;     // m, v may be NULL (pass 1)
    a->t[i].reserved = (bits & 15) == 0 ? 1 : 0;
  }
  return 0;
}

}  // namespace opt
}  // namespace b2n

using namespace b2n;

extern "C" int b2n_opt_prepare(const b2n_opt_tensor* tensors, int n_tensors, const float* grad_scale, float* norm2,
                               b2n_stream_t stream) {
  B2N_REQUIRE(tensors && n_tensors > 0 && n_tensors <= B2N_OPT_MAX_TENSORS && norm2, "1..B2N_OPT_MAX_TENSORS tensors, norm2 required");
  opt::Args a{};
  B2N_REQUIRE(opt::fill(&a, tensors, n_tensors) == 0, "bad tensor descriptor");
  a.grad_scale = grad_scale, a.norm2 = norm2;
  opt::k_opt_prepare<<<dim3(kSMs * 8, n_tensors), 256, 0, (cudaStream_t)stream>>>(a);
  return check_launch("b2n_opt_prepare");
}

extern "C" int b2n_opt_adamw(const b2n_opt_tensor* tensors, int n_tensors, const float* grad_scale, int already_unscaled,
                             const float* found_inf, const float* step_dev, const float* norm2, const float* max_norm_host,
                             b2n_stream_t stream) {
  B2N_REQUIRE(tensors && n_tensors > 0 && n_tensors <= B2N_OPT_MAX_TENSORS, "1..B2N_OPT_MAX_TENSORS tensors");
  opt::Args a{};
  B2N_REQUIRE(opt::fill(&a, tensors, n_tensors) == 0, "bad tensor descriptor");
  for (int i = 0; i < n_tensors; ++i) B2N_REQUIRE(tensors[i].m && tensors[i].v, "null moment buffer");
  a.grad_scale = grad_scale, a.found_inf = found_inf, a.step_dev = step_dev, a.norm2 = const_cast<float*>(norm2);
  for (int g = 0; g < B2N_OPT_MAX_GROUPS; ++g) a.max_norm[g] = (max_norm_host && norm2) ? max_norm_host[g] : -1.f;
  opt::k_opt_adamw<<<dim3(kSMs * 8, n_tensors), 256, 0, (cudaStream_t)stream>>>(a, already_unscaled);
  return check_launch("b2n_opt_adamw");
}
