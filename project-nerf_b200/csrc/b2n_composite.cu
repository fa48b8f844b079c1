// Alpha compositing forward / exact reverse-scan backward.  Replaces volume_render
// (src/renderer.py:204-237) and the mean_delta_x re-composite (src/renderer.py:363-380).
//
// One warp per ray; lane l owns samples l, l+32, ... (coalesced).  The exclusive transmittance
// product is a warp product-scan per 32-sample chunk with a carried prefix; the backward is the
// exact adjoint written as a reverse affine scan (no division by (1 - alpha + 1e-10)):
//     w_i = alpha_i T_i,  T_i = prod_{j<i} a_j,  a_j = 1 - alpha_j + 1e-10
//     R_{i-1} = alpha_i v_i + a_i R_i,  R_{N-1} = 0,  v_i = dL/dw_i
//     dL/dalpha_i = T_i (v_i - R_i),  dalpha_i/dsigma_i = delta_i exp(-sigma_i delta_i)
// Per-sample inputs are either dense [B*N] or compact (active samples only, addressed through
// the mask words + per-ray offsets produced by b2n_march_mask / b2n_march_scan), which removes
// the scatter-into-zeros pass of src/renderer.py:328-338.
#include "b2n_common.cuh"

namespace b2n {

struct RayCtx {
  float dn;       // |d|
  int64_t zbase;  // r*N
};

__device__ __forceinline__ float interval(const float* __restrict__ z, int64_t zbase, int s, int N, float zs, int lane,
                                          float dn) {
  // delta_s = z[s+1]-z[s] (1e10 for the last sample), times |d|   (renderer.py:213-215)
  float zn = __shfl_down_sync(0xffffffffu, zs, 1);
  if (lane == 31 && s + 1 < N) zn = z[zbase + s + 1];
  float d = (s == N - 1) ? 1e10f : __fsub_rn(zn, zs);
  return __fmul_rn(d, dn);
}

__device__ __forceinline__ float warp_incl_prod(float a, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    float t = __shfl_up_sync(0xffffffffu, a, o);
    if (lane >= o) a *= t;
  }
  return a;
}

__global__ void __launch_bounds__(256)
k_composite_fwd(const float* __restrict__ rgb, const float* __restrict__ sigma, const float* __restrict__ dx,
                const float* __restrict__ z, const float* __restrict__ rays_d, const float* __restrict__ bg,
                int bg_per_ray, const uint32_t* __restrict__ mask_words, const int32_t* __restrict__ ray_offset,
                int64_t B, int N, float* __restrict__ color, float* __restrict__ depth, float* __restrict__ acc,
                float* __restrict__ mean_dx) {
  const int lane = threadIdx.x & 31;
  const int64_t warps = (int64_t)gridDim.x * (blockDim.x >> 5);
  const int W = (N + 31) >> 5;
  const uint32_t lt = (1u << lane) - 1u;
  for (int64_t r = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); r < B; r += warps) {
    const float d0 = __ldg(rays_d + 3 * r), d1 = __ldg(rays_d + 3 * r + 1), d2 = __ldg(rays_d + 3 * r + 2);
    const float dn = sqrtf(d0 * d0 + d1 * d1 + d2 * d2);
    const int64_t zb = r * N;
    int64_t off = mask_words ? (int64_t)ray_offset[r] : zb;
    float T0 = 1.f, c0 = 0.f, c1 = 0.f, c2 = 0.f, dd = 0.f, aa = 0.f, m0 = 0.f, m1 = 0.f, m2 = 0.f;
    // a ray without a single active sample composites to exactly zero weights: skip its scan
    const bool empty_ray = mask_words && ray_offset[r + 1] == ray_offset[r];
    if (N > 1 && !empty_ray) {  // N == 1: the reference composites nothing (empty interval tensor)
      for (int k = 0; k < W; ++k) {
        const int s = (k << 5) + lane;
        const bool valid = s < N;
        uint32_t w = 0xffffffffu;
        if (mask_words) w = mask_words[r * W + k];
        const bool act = valid && ((w >> lane) & 1u);
        const int64_t idx = mask_words ? off + __popc(w & lt) : off + lane;
        const float zs = valid ? __ldcs(z + zb + s) : 0.f;
        const float dl = interval(z, zb, s, N, zs, lane, dn);
        float sg = 0.f, r0 = 0.f, r1 = 0.f, r2 = 0.f;
        if (act) {
          sg = __ldcs(sigma + idx);
          r0 = __ldcs(rgb + 3 * idx), r1 = __ldcs(rgb + 3 * idx + 1), r2 = __ldcs(rgb + 3 * idx + 2);
        }
        const float alpha = valid ? __fsub_rn(1.0f, expf(-__fmul_rn(sg, dl))) : 0.f;
        const float a = valid ? __fadd_rn(__fsub_rn(1.0f, alpha), 1e-10f) : 1.f;
        const float incl = warp_incl_prod(a, lane);
        float excl = __shfl_up_sync(0xffffffffu, incl, 1);
        if (lane == 0) excl = 1.f;
        const float wt = alpha * (T0 * excl);
        c0 += wt * r0, c1 += wt * r1, c2 += wt * r2;
        dd += wt * zs;
        aa += wt;
        if (dx && act) {
          m0 += wt * __ldcs(dx + 3 * idx), m1 += wt * __ldcs(dx + 3 * idx + 1), m2 += wt * __ldcs(dx + 3 * idx + 2);
        }
        T0 *= __shfl_sync(0xffffffffu, incl, 31);
        off += mask_words ? __popc(w) : 32;
      }
    }
    c0 = warp_sum(c0), c1 = warp_sum(c1), c2 = warp_sum(c2), dd = warp_sum(dd), aa = warp_sum(aa);
    if (dx) m0 = warp_sum(m0), m1 = warp_sum(m1), m2 = warp_sum(m2);
    if (lane == 0) {
      if (bg) {
        const float* b = bg + (bg_per_ray ? 3 * r : 0);
        const float rem = 1.0f - aa;
        c0 += rem * b[0], c1 += rem * b[1], c2 += rem * b[2];
      }
      color[3 * r] = c0, color[3 * r + 1] = c1, color[3 * r + 2] = c2;
      if (depth) depth[r] = dd;
      if (acc) acc[r] = aa;
      if (mean_dx) mean_dx[3 * r] = m0, mean_dx[3 * r + 1] = m1, mean_dx[3 * r + 2] = m2;
    }
  }
}

__global__ void __launch_bounds__(256)
k_composite_bwd(const float* __restrict__ rgb, const float* __restrict__ sigma, const float* __restrict__ dx,
                const float* __restrict__ z, const float* __restrict__ rays_d, const float* __restrict__ bg,
                int bg_per_ray, const uint32_t* __restrict__ mask_words, const int32_t* __restrict__ ray_offset,
                int64_t B, int N, const float* __restrict__ g_color, const float* __restrict__ g_depth,
                const float* __restrict__ g_acc, const float* __restrict__ g_mdx, float* __restrict__ g_rgb,
                float* __restrict__ g_sigma, float* __restrict__ g_dx) {
  const int lane = threadIdx.x & 31;
  const int64_t warps = (int64_t)gridDim.x * (blockDim.x >> 5);
  const int W = (N + 31) >> 5;  // <= 32
  const uint32_t lt = (1u << lane) - 1u;
  for (int64_t r = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); r < B; r += warps) {
    const float d0 = __ldg(rays_d + 3 * r), d1 = __ldg(rays_d + 3 * r + 1), d2 = __ldg(rays_d + 3 * r + 2);
    const float dn = sqrtf(d0 * d0 + d1 * d1 + d2 * d2);
    const int64_t zb = r * N;
    const int64_t off0 = mask_words ? (int64_t)ray_offset[r] : zb;
    const float gc0 = g_color ? g_color[3 * r] : 0.f, gc1 = g_color ? g_color[3 * r + 1] : 0.f,
                gc2 = g_color ? g_color[3 * r + 2] : 0.f;
    const float gd = g_depth ? g_depth[r] : 0.f;
    float ga = g_acc ? g_acc[r] : 0.f;
    if (bg) {  // color += (1 - acc) * bg
      const float* b = bg + (bg_per_ray ? 3 * r : 0);
      ga -= gc0 * b[0] + gc1 * b[1] + gc2 * b[2];
    }
    const float gm0 = (g_mdx && dx) ? g_mdx[3 * r] : 0.f, gm1 = (g_mdx && dx) ? g_mdx[3 * r + 1] : 0.f,
                gm2 = (g_mdx && dx) ? g_mdx[3 * r + 2] : 0.f;

    if (mask_words && ray_offset[r + 1] == ray_offset[r]) continue;  // no active sample: nothing to write
    if (N == 1) {  // nothing was composited: all sample gradients are zero
      const bool act = lane == 0 && (!mask_words || (mask_words[r * W] & 1u));
      if (act) {
        if (g_sigma) g_sigma[off0] = 0.f;
        if (g_rgb) g_rgb[3 * off0] = 0.f, g_rgb[3 * off0 + 1] = 0.f, g_rgb[3 * off0 + 2] = 0.f;
        if (g_dx && dx) g_dx[3 * off0] = 0.f, g_dx[3 * off0 + 1] = 0.f, g_dx[3 * off0 + 2] = 0.f;
      }
      continue;
    }

    // pass 1 (front to back): transmittance and compact offset entering every chunk.
    // lane k keeps the values of chunk k.
    float carryT = 1.f, myT = 1.f;
    int carryOff = 0, myOff = 0;
    for (int k = 0; k < W; ++k) {
      const int s = (k << 5) + lane;
      const bool valid = s < N;
      uint32_t w = 0xffffffffu;
      if (mask_words) w = mask_words[r * W + k];
      const bool act = valid && ((w >> lane) & 1u);
      if (lane == k) myT = carryT, myOff = carryOff;
      const int64_t idx = mask_words ? off0 + carryOff + __popc(w & lt) : off0 + carryOff + lane;
      const float zs = valid ? z[zb + s] : 0.f;
      const float dl = interval(z, zb, s, N, zs, lane, dn);
      const float sg = act ? sigma[idx] : 0.f;
      const float alpha = valid ? __fsub_rn(1.0f, expf(-__fmul_rn(sg, dl))) : 0.f;
      const float a = valid ? __fadd_rn(__fsub_rn(1.0f, alpha), 1e-10f) : 1.f;
      const float incl = warp_incl_prod(a, lane);
      carryT *= __shfl_sync(0xffffffffu, incl, 31);
      carryOff += mask_words ? __popc(w) : 32;
    }

    // pass 2 (back to front): reverse affine scan
    float Rend = 0.f;
    for (int k = W - 1; k >= 0; --k) {
      const int s = (k << 5) + lane;
      const bool valid = s < N;
      uint32_t w = 0xffffffffu;
      if (mask_words) w = mask_words[r * W + k];
      const bool act = valid && ((w >> lane) & 1u);
      const float Tin = __shfl_sync(0xffffffffu, myT, k);
      const int offk = __shfl_sync(0xffffffffu, myOff, k);
      const int64_t idx = mask_words ? off0 + offk + __popc(w & lt) : off0 + offk + lane;
      const float zs = valid ? z[zb + s] : 0.f;
      const float dl = interval(z, zb, s, N, zs, lane, dn);
      float sg = 0.f, r0 = 0.f, r1 = 0.f, r2 = 0.f, x0 = 0.f, x1 = 0.f, x2 = 0.f;
      if (act) {
        sg = __ldcs(sigma + idx);
        r0 = __ldcs(rgb + 3 * idx), r1 = __ldcs(rgb + 3 * idx + 1), r2 = __ldcs(rgb + 3 * idx + 2);
        if (dx) x0 = __ldcs(dx + 3 * idx), x1 = __ldcs(dx + 3 * idx + 1), x2 = __ldcs(dx + 3 * idx + 2);
      }
      const float e = expf(-__fmul_rn(sg, dl));
      const float alpha = valid ? __fsub_rn(1.0f, e) : 0.f;
      const float a = valid ? __fadd_rn(__fsub_rn(1.0f, alpha), 1e-10f) : 1.f;
      const float incl = warp_incl_prod(a, lane);
      float excl = __shfl_up_sync(0xffffffffu, incl, 1);
      if (lane == 0) excl = 1.f;
      const float T = Tin * excl;
      const float v = gc0 * r0 + gc1 * r1 + gc2 * r2 + gd * zs + ga + gm0 * x0 + gm1 * x1 + gm2 * x2;
      // suffix composition of f_s(R) = b_s + a_s R over lanes s..31
      float fa = a, fb = valid ? alpha * v : 0.f;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const float a2 = __shfl_down_sync(0xffffffffu, fa, o);
        const float b2 = __shfl_down_sync(0xffffffffu, fb, o);
        if (lane + o < 32) {
          fb = fb + fa * b2;
          fa = fa * a2;
        }
      }
      // R_s = (f_{s+1} o ... o f_31)(Rend)
      float na = __shfl_down_sync(0xffffffffu, fa, 1), nb = __shfl_down_sync(0xffffffffu, fb, 1);
      if (lane == 31) na = 1.f, nb = 0.f;
      const float Rs = nb + na * Rend;
      if (act) {
        const float wt = alpha * T;
        if (g_sigma) __stcs(g_sigma + idx, T * (v - Rs) * dl * e);
        if (g_rgb) __stcs(g_rgb + 3 * idx, wt * gc0), __stcs(g_rgb + 3 * idx + 1, wt * gc1), __stcs(g_rgb + 3 * idx + 2, wt * gc2);
        if (g_dx && dx) __stcs(g_dx + 3 * idx, wt * gm0), __stcs(g_dx + 3 * idx + 1, wt * gm1), __stcs(g_dx + 3 * idx + 2, wt * gm2);
      }
      const float A0 = __shfl_sync(0xffffffffu, fa, 0), B0 = __shfl_sync(0xffffffffu, fb, 0);
      Rend = B0 + A0 * Rend;
    }
  }
}


// ---------------------------------------------------------------------------------------------
// Register-resident variants for N <= 32*WC samples per ray (every shipped config: N = 64 / 128).
// All global loads of a ray (mask words, z, sigma, rgb, dx of every chunk) are issued before the
// first scan step, so one warp keeps ~5*WC independent 128-byte requests in flight instead of 5,
// and the backward re-uses z / sigma / exp(-sigma*delta) from its forward sweep instead of
// re-reading and re-computing them.  Arithmetic (and therefore every result bit) is the same as
// in the chunk-loop kernels above, which remain the path for longer rays.
// per-ray header (direction, mask words of the lanes, compact offsets): fetched ONE RAY AHEAD so that the
// dependent sample loads (sigma / rgb addresses come from the mask) can issue at the top of an iteration
struct RayHdr {
  float d0, d1, d2;
  uint32_t mw;
  int off0, off1;
};
__device__ __forceinline__ RayHdr load_hdr(const float* __restrict__ rays_d, const uint32_t* __restrict__ mask_words,
                                           const int32_t* __restrict__ ray_offset, int64_t r, int W, int lane) {
  RayHdr h;
  h.d0 = __ldg(rays_d + 3 * r), h.d1 = __ldg(rays_d + 3 * r + 1), h.d2 = __ldg(rays_d + 3 * r + 2);
  h.mw = 0xffffffffu, h.off0 = 0, h.off1 = 1;
  if (mask_words) {
    h.mw = lane < W ? __ldg(mask_words + r * W + lane) : 0u;
    h.off0 = __ldg(ray_offset + r), h.off1 = __ldg(ray_offset + r + 1);
  }
  return h;
}

template <int WC>
__global__ void __launch_bounds__(32)
k_composite_fwd_reg(const float* __restrict__ rgb, const float* __restrict__ sigma, const float* __restrict__ dx,
                    const float* __restrict__ z, const float* __restrict__ rays_d, const float* __restrict__ bg,
                    int bg_per_ray, const uint32_t* __restrict__ mask_words, const int32_t* __restrict__ ray_offset,
                    int64_t B, int N, float* __restrict__ color, float* __restrict__ depth, float* __restrict__ acc,
                    float* __restrict__ mean_dx) {
  const int lane = threadIdx.x;      // ONE warp per block: the ray index below is provably warp-uniform, so the
  const int W = (N + 31) >> 5;       // shuffles need no re-convergence code and addresses live in uniform registers
  const uint32_t lt = (1u << lane) - 1u;
  RayHdr nxt = load_hdr(rays_d, mask_words, ray_offset, blockIdx.x < B ? (int64_t)blockIdx.x : 0, W, lane);
  for (int64_t r = blockIdx.x; r < B; r += gridDim.x) {
    const RayHdr h = nxt;
    if (r + gridDim.x < B) nxt = load_hdr(rays_d, mask_words, ray_offset, r + gridDim.x, W, lane);
    const float dn = sqrtf(h.d0 * h.d0 + h.d1 * h.d1 + h.d2 * h.d2);
    const int64_t zb = r * N;
    int64_t off = mask_words ? (int64_t)h.off0 : zb;
    float T0 = 1.f, c0 = 0.f, c1 = 0.f, c2 = 0.f, dd = 0.f, aa = 0.f, m0 = 0.f, m1 = 0.f, m2 = 0.f;
    const bool empty_ray = mask_words && __all_sync(0xffffffffu, h.off1 == h.off0);
    if (N > 1 && !empty_ray) {
      const uint32_t mw = h.mw;
      float zs[WC], sg[WC], cr[WC][3], xr[WC][3];
      // ---- all loads first
#pragma unroll
      for (int k = 0; k < WC; ++k) {
        zs[k] = 0.f, sg[k] = 0.f;
        cr[k][0] = cr[k][1] = cr[k][2] = 0.f;
        xr[k][0] = xr[k][1] = xr[k][2] = 0.f;
        if (k < W) {
          const int s = (k << 5) + lane;
          const bool valid = s < N;
          const uint32_t w = __shfl_sync(0xffffffffu, mw, k);
          const bool act = valid && ((w >> lane) & 1u);
          const int64_t idx = mask_words ? off + __popc(w & lt) : off + lane;
          if (valid) zs[k] = __ldcs(z + zb + s);
          if (act) {
            sg[k] = __ldcs(sigma + idx);
            cr[k][0] = __ldcs(rgb + 3 * idx), cr[k][1] = __ldcs(rgb + 3 * idx + 1), cr[k][2] = __ldcs(rgb + 3 * idx + 2);
            if (dx) xr[k][0] = __ldcs(dx + 3 * idx), xr[k][1] = __ldcs(dx + 3 * idx + 1), xr[k][2] = __ldcs(dx + 3 * idx + 2);
          }
          off += mask_words ? __popc(w) : 32;
        }
      }
      // ---- per-sample alpha and attenuation of every chunk (independent of each other)
      float alpha[WC], incl[WC];
#pragma unroll
      for (int k = 0; k < WC; ++k) {
        const int s = (k << 5) + lane;
        const bool valid = s < N;
        float zn = __shfl_down_sync(0xffffffffu, zs[k], 1);
        const float z_next0 = __shfl_sync(0xffffffffu, zs[(k + 1 < WC) ? k + 1 : k], 0);
        if (lane == 31 && s + 1 < N) zn = z_next0;
        const float dl = __fmul_rn((s == N - 1) ? 1e10f : __fsub_rn(zn, zs[k]), dn);
        alpha[k] = valid ? __fsub_rn(1.0f, expf(-__fmul_rn(sg[k], dl))) : 0.f;
        incl[k] = valid ? __fadd_rn(__fsub_rn(1.0f, alpha[k]), 1e-10f) : 1.f;
      }
      // ---- the WC warp product-scans run interleaved: WC independent shuffles in flight per step
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
#pragma unroll
        for (int k = 0; k < WC; ++k) {
          const float t = __shfl_up_sync(0xffffffffu, incl[k], o);
          if (lane >= o) incl[k] *= t;
        }
      }
#pragma unroll
      for (int k = 0; k < WC; ++k) {
        float excl = __shfl_up_sync(0xffffffffu, incl[k], 1);
        if (lane == 0) excl = 1.f;
        const float wt = alpha[k] * (T0 * excl);
        c0 += wt * cr[k][0], c1 += wt * cr[k][1], c2 += wt * cr[k][2];
        dd += wt * zs[k];
        aa += wt;
        if (dx) m0 += wt * xr[k][0], m1 += wt * xr[k][1], m2 += wt * xr[k][2];
        T0 *= __shfl_sync(0xffffffffu, incl[k], 31);
      }
    }
    // ---- the 5 (8) ray sums reduce interleaved as well
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      c0 += __shfl_xor_sync(0xffffffffu, c0, o), c1 += __shfl_xor_sync(0xffffffffu, c1, o);
      c2 += __shfl_xor_sync(0xffffffffu, c2, o), dd += __shfl_xor_sync(0xffffffffu, dd, o);
      aa += __shfl_xor_sync(0xffffffffu, aa, o);
      if (dx) {
        m0 += __shfl_xor_sync(0xffffffffu, m0, o), m1 += __shfl_xor_sync(0xffffffffu, m1, o);
        m2 += __shfl_xor_sync(0xffffffffu, m2, o);
      }
    }
    if (lane == 0) {
      if (bg) {
        const float* b = bg + (bg_per_ray ? 3 * r : 0);
        const float rem = 1.0f - aa;
        c0 += rem * b[0], c1 += rem * b[1], c2 += rem * b[2];
      }
      color[3 * r] = c0, color[3 * r + 1] = c1, color[3 * r + 2] = c2;
      if (depth) depth[r] = dd;
      if (acc) acc[r] = aa;
      if (mean_dx) mean_dx[3 * r] = m0, mean_dx[3 * r + 1] = m1, mean_dx[3 * r + 2] = m2;
    }
  }
}

template <int WC>
__global__ void __launch_bounds__(32)
k_composite_bwd_reg(const float* __restrict__ rgb, const float* __restrict__ sigma, const float* __restrict__ dx,
                    const float* __restrict__ z, const float* __restrict__ rays_d, const float* __restrict__ bg,
                    int bg_per_ray, const uint32_t* __restrict__ mask_words, const int32_t* __restrict__ ray_offset,
                    int64_t B, int N, const float* __restrict__ g_color, const float* __restrict__ g_depth,
                    const float* __restrict__ g_acc, const float* __restrict__ g_mdx, float* __restrict__ g_rgb,
                    float* __restrict__ g_sigma, float* __restrict__ g_dx) {
  const int lane = threadIdx.x;      // one warp per block (see the forward kernel)
  const int W = (N + 31) >> 5;       // <= WC
  const uint32_t lt = (1u << lane) - 1u;
  RayHdr nxt = load_hdr(rays_d, mask_words, ray_offset, blockIdx.x < B ? (int64_t)blockIdx.x : 0, W, lane);
  for (int64_t r = blockIdx.x; r < B; r += gridDim.x) {
    const RayHdr h = nxt;
    if (r + gridDim.x < B) nxt = load_hdr(rays_d, mask_words, ray_offset, r + gridDim.x, W, lane);
    if (mask_words && __all_sync(0xffffffffu, h.off1 == h.off0)) continue;  // no active sample: nothing to write
    const float dn = sqrtf(h.d0 * h.d0 + h.d1 * h.d1 + h.d2 * h.d2);
    const int64_t zb = r * N;
    const int64_t off0 = mask_words ? (int64_t)h.off0 : zb;
    const float gc0 = g_color ? g_color[3 * r] : 0.f, gc1 = g_color ? g_color[3 * r + 1] : 0.f,
                gc2 = g_color ? g_color[3 * r + 2] : 0.f;
    const float gd = g_depth ? g_depth[r] : 0.f;
    float ga = g_acc ? g_acc[r] : 0.f;
    if (bg) {  // color += (1 - acc) * bg
      const float* b = bg + (bg_per_ray ? 3 * r : 0);
      ga -= gc0 * b[0] + gc1 * b[1] + gc2 * b[2];
    }
    const float gm0 = (g_mdx && dx) ? g_mdx[3 * r] : 0.f, gm1 = (g_mdx && dx) ? g_mdx[3 * r + 1] : 0.f,
                gm2 = (g_mdx && dx) ? g_mdx[3 * r + 2] : 0.f;
    const uint32_t mw = h.mw;
    // ---- all loads first
    float zs[WC], sg[WC], cr[WC][3], xr[WC][3];
    int64_t idxs[WC];
    bool acts[WC];
    {
      int64_t off = off0;
#pragma unroll
      for (int k = 0; k < WC; ++k) {
        zs[k] = 0.f, sg[k] = 0.f;
        cr[k][0] = cr[k][1] = cr[k][2] = 0.f;
        xr[k][0] = xr[k][1] = xr[k][2] = 0.f;
        idxs[k] = 0, acts[k] = false;
        if (k < W) {
          const int s = (k << 5) + lane;
          const bool valid = s < N;
          const uint32_t w = __shfl_sync(0xffffffffu, mw, k);
          acts[k] = valid && ((w >> lane) & 1u);
          idxs[k] = mask_words ? off + __popc(w & lt) : off + lane;
          if (valid) zs[k] = __ldcs(z + zb + s);
          if (acts[k]) {
            sg[k] = __ldcs(sigma + idxs[k]);
            cr[k][0] = __ldcs(rgb + 3 * idxs[k]), cr[k][1] = __ldcs(rgb + 3 * idxs[k] + 1), cr[k][2] = __ldcs(rgb + 3 * idxs[k] + 2);
            if (dx) xr[k][0] = __ldcs(dx + 3 * idxs[k]), xr[k][1] = __ldcs(dx + 3 * idxs[k] + 1), xr[k][2] = __ldcs(dx + 3 * idxs[k] + 2);
          }
          off += mask_words ? __popc(w) : 32;
        }
      }
    }
    // ---- per-sample interval, exp, attenuation and adjoint seed of every chunk (independent of each other)
    float dl[WC], ee[WC], incl[WC], fa[WC], fb[WC], vv[WC];
#pragma unroll
    for (int k = 0; k < WC; ++k) {
      const int s = (k << 5) + lane;
      const bool valid = s < N;
      float zn = __shfl_down_sync(0xffffffffu, zs[k], 1);
      const float z_next0 = __shfl_sync(0xffffffffu, zs[(k + 1 < WC) ? k + 1 : k], 0);
      if (lane == 31 && s + 1 < N) zn = z_next0;
      dl[k] = __fmul_rn((s == N - 1) ? 1e10f : __fsub_rn(zn, zs[k]), dn);
      ee[k] = expf(-__fmul_rn(sg[k], dl[k]));
      const float alpha = valid ? __fsub_rn(1.0f, ee[k]) : 0.f;
      fa[k] = valid ? __fadd_rn(__fsub_rn(1.0f, alpha), 1e-10f) : 1.f;
      incl[k] = fa[k];
      vv[k] = gc0 * cr[k][0] + gc1 * cr[k][1] + gc2 * cr[k][2] + gd * zs[k] + ga + gm0 * xr[k][0] + gm1 * xr[k][1] +
              gm2 * xr[k][2];
      fb[k] = valid ? alpha * vv[k] : 0.f;
    }
    // ---- WC product scans (transmittance) and WC suffix compositions of f_s(R) = b_s + a_s R, interleaved
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
#pragma unroll
      for (int k = 0; k < WC; ++k) {
        const float t = __shfl_up_sync(0xffffffffu, incl[k], o);
        const float a2 = __shfl_down_sync(0xffffffffu, fa[k], o);
        const float b2 = __shfl_down_sync(0xffffffffu, fb[k], o);
        if (lane >= o) incl[k] *= t;
        if (lane + o < 32) {
          fb[k] = fb[k] + fa[k] * b2;
          fa[k] = fa[k] * a2;
        }
      }
    }
    float Tin[WC];
    {
      float carryT = 1.f;
#pragma unroll
      for (int k = 0; k < WC; ++k) {
        Tin[k] = carryT;
        carryT *= __shfl_sync(0xffffffffu, incl[k], 31);
      }
    }
    // ---- back to front: R_s = (f_{s+1} o ... o f_last)(0), gradients
    float Rend = 0.f;
#pragma unroll
    for (int k = WC - 1; k >= 0; --k) {
      const int s = (k << 5) + lane;
      const bool valid = s < N;
      const float alpha = valid ? __fsub_rn(1.0f, ee[k]) : 0.f;
      float excl = __shfl_up_sync(0xffffffffu, incl[k], 1);
      if (lane == 0) excl = 1.f;
      const float T = Tin[k] * excl;
      float na = __shfl_down_sync(0xffffffffu, fa[k], 1), nb = __shfl_down_sync(0xffffffffu, fb[k], 1);
      if (lane == 31) na = 1.f, nb = 0.f;
      const float Rs = nb + na * Rend;
      if (acts[k]) {
        const float wt = alpha * T;
        const int64_t idx = idxs[k];
        if (g_sigma) __stcs(g_sigma + idx, T * (vv[k] - Rs) * dl[k] * ee[k]);
        if (g_rgb) __stcs(g_rgb + 3 * idx, wt * gc0), __stcs(g_rgb + 3 * idx + 1, wt * gc1), __stcs(g_rgb + 3 * idx + 2, wt * gc2);
        if (g_dx && dx) __stcs(g_dx + 3 * idx, wt * gm0), __stcs(g_dx + 3 * idx + 1, wt * gm1), __stcs(g_dx + 3 * idx + 2, wt * gm2);
      }
      const float A0 = __shfl_sync(0xffffffffu, fa[k], 0), B0 = __shfl_sync(0xffffffffu, fb[k], 0);
      Rend = B0 + A0 * Rend;
    }
  }
}

// one-warp blocks: 32 resident per SM, a few waves so that rays of different cost balance out
static inline unsigned warp_grid(int64_t B) {
  const int64_t cap = (int64_t)kSMs * 32 * 4;
  return (unsigned)(B < cap ? (B < 1 ? 1 : B) : cap);
}

static inline unsigned ray_grid(int64_t B) {
  int64_t blocks = (B + 7) / 8;
  const int64_t cap = (int64_t)kSMs * 8 * 4;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (unsigned)blocks;
}

}  // namespace b2n

using namespace b2n;

extern "C" int b2n_composite_fwd(const float* rgb, const float* sigma, const float* dx, const float* z,
                                 const float* rays_d, const float* bg, int bg_per_ray, const uint32_t* mask_words,
                                 const int32_t* ray_offset, int64_t B, int N, float* color, float* depth, float* acc,
                                 float* mean_dx, b2n_stream_t stream) {
  B2N_REQUIRE(B >= 0 && N > 0 && N <= 1024, "bad B or N (N <= 1024)");
  if (B == 0) return B2N_OK;
  B2N_REQUIRE(rgb && sigma && z && rays_d && color, "null pointer");
  B2N_REQUIRE(!mask_words || ray_offset, "compact layout needs ray_offset");
  B2N_REQUIRE(!mean_dx || dx, "mean_dx needs dx");
  if (N > 1 && N <= 128) {
    auto kern = N <= 64 ? k_composite_fwd_reg<2> : k_composite_fwd_reg<4>;
    kern<<<warp_grid(B), 32, 0, (cudaStream_t)stream>>>(rgb, sigma, dx, z, rays_d, bg, bg_per_ray, mask_words, ray_offset, B,
                                                       N, color, depth, acc, mean_dx);
  } else {
    k_composite_fwd<<<ray_grid(B), 256, 0, (cudaStream_t)stream>>>(rgb, sigma, dx, z, rays_d, bg, bg_per_ray, mask_words,
                                                                  ray_offset, B, N, color, depth, acc, mean_dx);
  }
  return check_launch("b2n_composite_fwd");
}

extern "C" int b2n_composite_bwd(const float* rgb, const float* sigma, const float* dx, const float* z,
                                 const float* rays_d, const float* bg, int bg_per_ray, const uint32_t* mask_words,
                                 const int32_t* ray_offset, int64_t B, int N, const float* g_color,
                                 const float* g_depth, const float* g_acc, const float* g_mean_dx, float* g_rgb,
                                 float* g_sigma, float* g_dx, b2n_stream_t stream) {
  B2N_REQUIRE(B >= 0 && N > 0 && N <= 1024, "bad B or N (N <= 1024)");
  if (B == 0) return B2N_OK;
  B2N_REQUIRE(rgb && sigma && z && rays_d, "null pointer");
  B2N_REQUIRE(!mask_words || ray_offset, "compact layout needs ray_offset");
  if (N > 1 && N <= 128) {
    auto kern = N <= 64 ? k_composite_bwd_reg<2> : k_composite_bwd_reg<4>;
    kern<<<warp_grid(B), 32, 0, (cudaStream_t)stream>>>(rgb, sigma, dx, z, rays_d, bg, bg_per_ray, mask_words, ray_offset, B,
                                                       N, g_color, g_depth, g_acc, g_mean_dx, g_rgb, g_sigma, g_dx);
  } else {
    k_composite_bwd<<<ray_grid(B), 256, 0, (cudaStream_t)stream>>>(rgb, sigma, dx, z, rays_d, bg, bg_per_ray, mask_words,
                                                                  ray_offset, B, N, g_color, g_depth, g_acc, g_mean_dx,
                                                                  g_rgb, g_sigma, g_dx);
  }
  return check_launch("b2n_composite_bwd");
}
