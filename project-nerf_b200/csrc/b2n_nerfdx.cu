// Input gradient of the 256-wide tcgen05 decoder (bf16 planes of b2n_nerf_mlp_bwd), on mma.sync.
#include "b2n_mma.cuh"

namespace b2n {
namespace ndx {

constexpr int THREADS = 128;

// fp32 matrix [rows_valid][cols_valid] (row stride ldw) -> bf16 smem [rows][cols + PAD], zero filled elsewhere
__device__ __forceinline__ void load_w(const float* __restrict__ W, int ldw, int rows_valid, int cols_valid, int rows,
                                       int cols, bf16* dst) {
  const int S = cols + PAD;
  for (int i = threadIdx.x; i < rows * cols; i += blockDim.x) {
    const int r = i / cols, c = i - r * cols;
    dst[r * S + c] = __float2bfloat16((r < rows_valid && c < cols_valid) ? __ldg(W + (size_t)r * ldw + c) : 0.f);
  }
}

static int persistent_grid(const void* kernel, size_t smem, int64_t warp_tiles) {
  int per_sm = 1;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, THREADS, smem);
  if (per_sm < 1) per_sm = 1;
  int64_t g = (int64_t)kSMs * per_sm;
  const int64_t blocks = (warp_tiles + THREADS / 32 - 1) / (THREADS / 32);
  if (g > blocks) g = blocks;
  if (g < 1) g = 1;
  return (int)g;
}

// ------------------------------------------------------------------------------ input gradient of the 256-wide decoder
// d x_enc [P, pos_dim] = dZ_0 W_0 + dZ_4 W_4[:, 256:256+pos_dim]   (NeRFDecoder, src/decoders.py:68-87: x feeds layer 0
// and, through the skip concat [h, x], layer 4).  dZ_l are the bf16 pre-activation gradient planes written by the
// tcgen05 backward chain (b2n_nerf_mlp_bwd); needed when the decoder input depends on a trainable deformation
// (Part 3: x_enc = gamma(x + delta_x), src/core.py:268-277).
template <int NTo>
__global__ void __launch_bounds__(THREADS) k_nerf_dx(const bf16* __restrict__ dz0, const bf16* __restrict__ dz4,
                                                     const float* __restrict__ W0, int ld0, const float* __restrict__ W4x,
                                                     int ld4, int pos_dim, int64_t P, float* __restrict__ gx, int ldg) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  bf16* sm = reinterpret_cast<bf16*>(smem_raw);
  constexpr int S = 8 * NTo + PAD;
  load_w(W0, ld0, 256, pos_dim, 256, 8 * NTo, sm);
  load_w(W4x, ld4, 256, pos_dim, 256, 8 * NTo, sm + 256 * S);
  __syncthreads();
  const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int64_t n_tiles = (P + 15) / 16;
  const int64_t wstride = (int64_t)gridDim.x * (THREADS / 32);
  for (int64_t tile = (int64_t)blockIdx.x * (THREADS / 32) + (threadIdx.x >> 5); tile < n_tiles; tile += wstride) {
    const int64_t pa = tile * 16 + g, pb = pa + 8;
    float c[NTo][4] = {};
#pragma unroll
    for (int q = 0; q < 4; ++q) {        // (plane, half): 128 gradient columns at a time
      const bf16* plane = (q < 2) ? dz0 : dz4;
      const int col0 = 128 * (q & 1);
      uint32_t a[8][4];
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int col = col0 + 16 * k + 2 * t;
        a[k][0] = pa < P ? __ldcs(reinterpret_cast<const uint32_t*>(plane + pa * 256 + col)) : 0u;
        a[k][1] = pb < P ? __ldcs(reinterpret_cast<const uint32_t*>(plane + pb * 256 + col)) : 0u;
        a[k][2] = pa < P ? __ldcs(reinterpret_cast<const uint32_t*>(plane + pa * 256 + col + 8)) : 0u;
        a[k][3] = pb < P ? __ldcs(reinterpret_cast<const uint32_t*>(plane + pb * 256 + col + 8)) : 0u;
      }
      gemm_dgrad<NTo, 8>(c, a, sm + ((q < 2) ? 0 : 256 * S) + col0 * S, S, lane);
    }
#pragma unroll
    for (int j = 0; j < NTo; ++j)
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int col = 8 * j + 2 * t + (i & 1);
        const int64_t p = (i & 2) ? pb : pa;
        if (p < P && col < pos_dim) gx[p * ldg + col] = c[j][i];
      }
  }
}

template <int NTo>
static int launch_nerf_dx(const bf16* dz0, const bf16* dz4, const float* W0, int ld0, const float* W4x, int ld4, int pos_dim,
                          int64_t P, float* gx, int ldg, cudaStream_t st) {
  constexpr size_t smem = (size_t)2 * 256 * (8 * NTo + PAD) * sizeof(bf16);
  auto k = k_nerf_dx<NTo>;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  const int grid = persistent_grid((const void*)k, smem, (P + 15) / 16);
  k<<<grid, THREADS, smem, st>>>(dz0, dz4, W0, ld0, W4x, ld4, pos_dim, P, gx, ldg);
  return check_launch("b2n_nerf_mlp_dx");
}

// ------------------------------------------------------------------------------ head weight gradients of the 256-wide decoder
// sigma_layer (1 x 256) and rgb_layer (3 x 128) (NeRFDecoder, src/decoders.py:60-66): dW = dz_small^T [H_7 | hv] over all
// points, db = column sums of dz_small -- two 4-row "GEMMs" that are pure streaming (768 bytes of bf16 planes per point
// against 12 multiply-adds per column): one pass over the two planes, 16-byte loads, one atomicAdd per output per CTA.
// (These were two cuBLAS split-K GEMMs plus a pad, a cast and a reduction in round 1: 6 % of the C1 step.)
constexpr int HW_WARPS = 16;
__global__ void __launch_bounds__(32 * HW_WARPS) k_nerf_head_wgrad(const float4* __restrict__ dzs, const uint4* __restrict__ h7,
                                                                   const uint4* __restrict__ hv, int64_t P,
                                                                   float* __restrict__ gw_sigma, float* __restrict__ gw_rgb,
                                                                   float* __restrict__ gb) {
  // a warp owns a point per step; lane l = 8 columns of H_7 (sigma head), lanes < 16 also 8 columns of hv (rgb head): a plane row
  // is 256 bf16 = 32 uint4, hv uses its first 128 columns = 16 uint4.  Two points per iteration keep four 16-byte loads in
  // flight per lane.  One persistent CTA per SM: the 644 outputs are combined through shared memory and leave as ONE atomicAdd per output per
  // CTA (the first version launched 8 x as many 4-warp CTAs: 760 k same-address atomics, 0.16 ms for a 0.035 ms stream)
  __shared__ float part[HW_WARPS][256 + 3 * 128 + 4];       // per-warp partial sums (plain stores: shared float atomics are CAS loops)
  const int warp = threadIdx.x >> 5, l = threadIdx.x & 31;
  float as[8] = {}, ar[3][8] = {};
  float bsum[4] = {};
  const int64_t stride = (int64_t)gridDim.x * HW_WARPS;
  auto fma_point = [&](const float4& d, const uint4& a, const uint4& b) {
    const __nv_bfloat162* a2 = reinterpret_cast<const __nv_bfloat162*>(&a);
    const __nv_bfloat162* b2 = reinterpret_cast<const __nv_bfloat162*>(&b);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 f = __bfloat1622float2(a2[j]);
      as[2 * j] += d.w * f.x, as[2 * j + 1] += d.w * f.y;
      const float2 h = __bfloat1622float2(b2[j]);          // lanes >= 16: zeros
      ar[0][2 * j] += d.x * h.x, ar[0][2 * j + 1] += d.x * h.y;
      ar[1][2 * j] += d.y * h.x, ar[1][2 * j + 1] += d.y * h.y;
      ar[2][2 * j] += d.z * h.x, ar[2][2 * j + 1] += d.z * h.y;
    }
    if (l == 0) bsum[0] += d.x, bsum[1] += d.y, bsum[2] += d.z, bsum[3] += d.w;
  };
  const uint4 zero4 = make_uint4(0u, 0u, 0u, 0u);
  int64_t p = (int64_t)blockIdx.x * HW_WARPS + warp;
  for (; p + stride < P; p += 2 * stride) {                            // two points: four 16-byte loads in flight per lane
    const int64_t q = p + stride;                                      // (four points were measured: slower)
    const float4 d0 = __ldg(dzs + p), d1 = __ldg(dzs + q);             // (d rgb_pre[3], d sigma_pre)
    const uint4 a0 = __ldcs(h7 + p * 32 + l), a1 = __ldcs(h7 + q * 32 + l);
    const uint4 b0 = l < 16 ? __ldcs(hv + p * 32 + l) : zero4, b1 = l < 16 ? __ldcs(hv + q * 32 + l) : zero4;
    fma_point(d0, a0, b0);
    fma_point(d1, a1, b1);
  }
  for (; p < P; p += stride) fma_point(__ldg(dzs + p), __ldcs(h7 + p * 32 + l), l < 16 ? __ldcs(hv + p * 32 + l) : zero4);
  float* mine = part[warp];
#pragma unroll
  for (int j = 0; j < 8; ++j) mine[8 * l + j] = as[j];
  if (l < 16) {
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
      for (int j = 0; j < 8; ++j) mine[256 + r * 128 + 8 * l + j] = ar[r][j];
  }
  if (l == 0)
    for (int j = 0; j < 4; ++j) mine[640 + j] = bsum[j];
  __syncthreads();
  for (int i = threadIdx.x; i < 644; i += blockDim.x) {
    float v = 0.f;
#pragma unroll
    for (int w = 0; w < HW_WARPS; ++w) v += part[w][i];
    float* dst = i < 256 ? gw_sigma + i : (i < 640 ? gw_rgb + (i - 256) : gb + (i - 640));
    atomicAdd(dst, v);
  }
}

}  // namespace ndx
}  // namespace b2n

using namespace b2n;
using namespace b2n::ndx;

extern "C" int b2n_nerf_mlp_dx(const void* dz0, const void* dz4, const float* W0, int ldw0, const float* W4x, int ldw4,
                               int pos_dim, int64_t P, float* g_x, int ldg, b2n_stream_t stream) {
  B2N_REQUIRE(P >= 0, "negative size");
  if (P == 0) return B2N_OK;
  B2N_REQUIRE(dz0 && dz4 && W0 && W4x && g_x, "null pointer");
  B2N_REQUIRE(pos_dim > 0 && pos_dim <= 96 && ldw0 >= pos_dim && ldw4 >= pos_dim && ldg >= pos_dim, "pos_dim <= 96 required");
  if (pos_dim <= 64)
    return launch_nerf_dx<8>((const bf16*)dz0, (const bf16*)dz4, W0, ldw0, W4x, ldw4, pos_dim, P, g_x, ldg, (cudaStream_t)stream);
  return launch_nerf_dx<12>((const bf16*)dz0, (const bf16*)dz4, W0, ldw0, W4x, ldw4, pos_dim, P, g_x, ldg, (cudaStream_t)stream);
}

// dz_small [P,4] fp32 (d rgb_pre[3], d sigma_pre) of b2n_nerf_mlp_bwd; h7 / hv: bf16 [P][256] planes 7 and 9 of b2n_nerf_mlp_fwd.
// ACCUMULATES: gw_sigma [256] = sum_p dz[p,3] H_7[p,:], gw_rgb [3][128] = sum_p dz[p,r] hv[p,:128], gb [4] = sum_p dz[p,:]
// (rgb biases 0..2, sigma bias 3).
extern "C" int b2n_nerf_mlp_head_wgrad(const float* dz_small, const void* h7_plane, const void* hv_plane, int64_t P,
                                       float* gw_sigma, float* gw_rgb, float* gb, b2n_stream_t stream) {
  B2N_REQUIRE(P >= 0, "negative size");
  if (P == 0) return B2N_OK;
  B2N_REQUIRE(dz_small && h7_plane && hv_plane && gw_sigma && gw_rgb && gb, "null pointer");
  B2N_REQUIRE(((reinterpret_cast<uintptr_t>(dz_small) | reinterpret_cast<uintptr_t>(h7_plane) |
                reinterpret_cast<uintptr_t>(hv_plane)) & 15) == 0, "16-byte aligned buffers required");
  int64_t blocks = (P + 2 * HW_WARPS - 1) / (2 * HW_WARPS);
  if (blocks > (int64_t)kSMs) blocks = kSMs;
  k_nerf_head_wgrad<<<(unsigned)blocks, 32 * HW_WARPS, 0, (cudaStream_t)stream>>>((const float4*)dz_small, (const uint4*)h7_plane,
                                                                      (const uint4*)hv_plane, P, gw_sigma, gw_rgb, gb);
  return check_launch("b2n_nerf_mlp_head_wgrad");
}
