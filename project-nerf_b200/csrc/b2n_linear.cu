// fp32 dense layers (exact-parity path): forward with fused bias + activation, data gradient
// with fused activation derivative, weight gradient as a split reduction over the points.
// Replaces torch.nn.Linear / cuBLAS in NeRFDecoder, DeformationNetwork, TimeModulationNetwork
// (src/decoders.py:55-66,176-189,347) and the matrices of the tinycudann FullyFusedMLPs
// (src/decoders.py:111-134,285-295) when running in fp32.  The bf16 tensor-core kernels
// (b2n_mlp64.cu, b2n_mlp256.cu) are the throughput path; this file is the 1e-4 path.
//
// One generic 128x64x16 register-tiled kernel, C[MxN] = A[MxK] * B[KxN], with strided
// accessors so that the three GEMM flavours share it:
//   fwd   : A = X  (k-contiguous),  B = W^T (k-contiguous)
//   dgrad : A = dY (k-contiguous),  B = W   (n-contiguous)
//   wgrad : A = dY^T (m-contiguous), B = X  (n-contiguous), K = points, split over gridDim.z
#include "b2n_common.cuh"

namespace b2n {

constexpr int BM = 128, BN = 64, BK = 16, TM = 8, TN = 4, NT = 256;

enum { EPI_FWD = 0, EPI_DGRAD = 1, EPI_WGRAD = 2 };

struct GemmArgs {
  const float* A; int64_t sa_m, sa_k;
  const float* B; int64_t sb_k, sb_n;
  float* C; int64_t ldc;
  int64_t M; int N; int64_t K;
  int64_t k_chunk;          // wgrad: reduction range per gridDim.z slice
  const float* bias;        // fwd
  const float* xact; int64_t ldxa;  // dgrad: activation output whose derivative masks C
  int act;
  int accumulate;
};

__device__ __forceinline__ float act_fwd(float v, int act) {
  if (act == B2N_ACT_RELU) return fmaxf(v, 0.f);
  if (act == B2N_ACT_SIGMOID) return 1.f / (1.f + expf(-v));
  return v;
}
__device__ __forceinline__ float act_grad_from_out(float y, int act) {
  if (act == B2N_ACT_RELU) return y > 0.f ? 1.f : 0.f;
  if (act == B2N_ACT_SIGMOID) return y * (1.f - y);
  return 1.f;
}

template <bool A_KCONT, bool B_KCONT, int EPI>
__global__ void __launch_bounds__(NT) k_gemm(const GemmArgs g) {
  __shared__ __align__(16) float As[BK][BM + 4];
  __shared__ __align__(16) float Bs[BK][BN + 4];
  const int t = threadIdx.x;
  const int64_t m0 = (int64_t)blockIdx.x * BM;
  const int n0 = blockIdx.y * BN;
  int64_t kbeg = 0, kend = g.K;
  if (EPI == EPI_WGRAD) {
    kbeg = (int64_t)blockIdx.z * g.k_chunk;
    kend = min(g.K, kbeg + g.k_chunk);
  }
  const int tm = (t / 16) * TM;  // 16 x 16 thread grid -> 128 x 64 outputs
  const int tn = (t % 16) * TN;
  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  for (int64_t k0 = kbeg; k0 < kend; k0 += BK) {
    // ---- A tile: BM x BK = 2048 elements, 8 per thread
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      int m, k;
      if (A_KCONT) { k = t % BK; m = t / BK + 16 * i; }
      else         { m = t % BM; k = t / BM + 2 * i; }
      const int64_t gm = m0 + m, gk = k0 + k;
      As[k][m] = (gm < g.M && gk < kend) ? __ldg(g.A + gm * g.sa_m + gk * g.sa_k) : 0.f;
    }
    // ---- B tile: BK x BN = 1024 elements, 4 per thread
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      int n, k;
      if (B_KCONT) { k = t % BK; n = t / BK + 16 * i; }
      else         { n = t % BN; k = t / BN + 4 * i; }
      const int gn = n0 + n;
      const int64_t gk = k0 + k;
      Bs[k][n] = (gn < g.N && gk < kend) ? __ldg(g.B + gk * g.sb_k + (int64_t)gn * g.sb_n) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      float a[TM], b[TN];
      const float4 a0 = *reinterpret_cast<const float4*>(&As[k][tm]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[k][tm + 4]);
      const float4 b0 = *reinterpret_cast<const float4*>(&Bs[k][tn]);
      a[0] = a0.x, a[1] = a0.y, a[2] = a0.z, a[3] = a0.w, a[4] = a1.x, a[5] = a1.y, a[6] = a1.z, a[7] = a1.w;
      b[0] = b0.x, b[1] = b0.y, b[2] = b0.z, b[3] = b0.w;
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }

#pragma unroll
  for (int i = 0; i < TM; ++i) {
    const int64_t gm = m0 + tm + i;
    if (gm >= g.M) continue;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      const int gn = n0 + tn + j;
      if (gn >= g.N) continue;
      float v = acc[i][j];
      float* c = g.C + gm * g.ldc + gn;
      if (EPI == EPI_FWD) {
        if (g.bias) v += __ldg(g.bias + gn);
        *c = act_fwd(v, g.act);
      } else if (EPI == EPI_DGRAD) {
        if (g.xact) v *= act_grad_from_out(g.xact[gm * g.ldxa + gn], g.act);
        *c = g.accumulate ? *c + v : v;
      } else {
        atomicAdd(c, v);
      }
    }
  }
}

// db[n] += sum_p dY[p, n]
__global__ void __launch_bounds__(256) k_colsum(const float* __restrict__ dY, int64_t ld, int64_t P, int N,
                                                int64_t rows_per_block, float* __restrict__ db) {
  const int n = blockIdx.y * 32 + (threadIdx.x & 31);
  const int64_t p0 = (int64_t)blockIdx.x * rows_per_block;
  const int64_t p1 = min(P, p0 + rows_per_block);
  float s = 0.f;
  if (n < N)
    for (int64_t p = p0 + (threadIdx.x >> 5); p < p1; p += 8) s += dY[p * ld + n];
  __shared__ float red[8][33];
  red[threadIdx.x >> 5][threadIdx.x & 31] = s;
  __syncthreads();
  if (threadIdx.x < 32 && n < N) {
    float tot = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) tot += red[i][threadIdx.x];
    atomicAdd(db + n, tot);
  }
}

__global__ void k_act_bwd(float* __restrict__ dY, int64_t lddy, const float* __restrict__ Y, int64_t ldy, int64_t P,
                          int N, int act) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P * N) return;
  const int64_t p = i / N;
  const int n = (int)(i - p * N);
  dY[p * lddy + n] *= act_grad_from_out(Y[p * ldy + n], act);
}

// softplus(h0 - 5), torch semantics (beta = 1, threshold = 20)   (decoders.py:153)
__global__ void k_sigma_head_fwd(const float* __restrict__ h, int64_t ld, int64_t P, float* __restrict__ sigma) {
  int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= P) return;
  const float v = h[p * ld] - 5.0f;
  sigma[p] = v > 20.f ? v : log1pf(expf(v));
}
__global__ void k_sigma_head_bwd(const float* __restrict__ h, int64_t ld, int64_t P, const float* __restrict__ gs,
                                 float* __restrict__ gh, int64_t ldg) {
  int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= P) return;
  const float v = h[p * ld] - 5.0f;
  const float d = v > 20.f ? 1.f : 1.f / (1.f + expf(-v));
  gh[p * ldg] += gs[p] * d;
}

}  // namespace b2n

using namespace b2n;

extern "C" int b2n_linear_fwd(const float* X, int ldx, const float* W, int ldw, const float* b, float* Y, int ldy,
                              int64_t P, int K, int N, int act, b2n_stream_t stream) {
  B2N_REQUIRE(P >= 0 && K > 0 && N > 0 && ldx >= K && ldw >= K && ldy >= N, "bad shape");
  B2N_REQUIRE(act >= 0 && act <= 2, "bad activation");
  if (P == 0) return B2N_OK;
  B2N_REQUIRE(X && W && Y, "null pointer");
  GemmArgs g{};
  g.A = X, g.sa_m = ldx, g.sa_k = 1;
  g.B = W, g.sb_k = 1, g.sb_n = ldw;
  g.C = Y, g.ldc = ldy, g.M = P, g.N = N, g.K = K, g.bias = b, g.act = act;
  dim3 grid(grid_for(P, BM), grid_for(N, BN));
  k_gemm<true, true, EPI_FWD><<<grid, NT, 0, (cudaStream_t)stream>>>(g);
  return check_launch("b2n_linear_fwd");
}

extern "C" int b2n_linear_dgrad(const float* dY, int lddy, const float* W, int ldw, const float* Xact, int ldxa,
                                int act, float* dX, int lddx, int64_t P, int K, int N, int accumulate,
                                b2n_stream_t stream) {
  B2N_REQUIRE(P >= 0 && K > 0 && N > 0 && lddy >= N && ldw >= K && lddx >= K, "bad shape");
  B2N_REQUIRE(act >= 0 && act <= 2 && (!Xact || ldxa >= K), "bad activation");
  if (P == 0) return B2N_OK;
  B2N_REQUIRE(dY && W && dX, "null pointer");
  GemmArgs g{};
  g.A = dY, g.sa_m = lddy, g.sa_k = 1;
  g.B = W, g.sb_k = ldw, g.sb_n = 1;   // B(k=n_out, n=k_in) = W[n_out, k_in]
  g.C = dX, g.ldc = lddx, g.M = P, g.N = K, g.K = N;
  g.xact = (act == B2N_ACT_NONE) ? nullptr : Xact, g.ldxa = ldxa, g.act = act, g.accumulate = accumulate;
  dim3 grid(grid_for(P, BM), grid_for(K, BN));
  k_gemm<true, false, EPI_DGRAD><<<grid, NT, 0, (cudaStream_t)stream>>>(g);
  return check_launch("b2n_linear_dgrad");
}

extern "C" int b2n_linear_wgrad(const float* dY, int lddy, const float* X, int ldx, float* dW, int lddw, float* db,
                                int64_t P, int K, int N, b2n_stream_t stream) {
  B2N_REQUIRE(P >= 0 && K > 0 && N > 0 && lddy >= N && ldx >= K && lddw >= K, "bad shape");
  if (P == 0) return B2N_OK;
  B2N_REQUIRE(dY && X && dW, "null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  GemmArgs g{};
  g.A = dY, g.sa_m = 1, g.sa_k = lddy;  // A(m=n_out, k=p)
  g.B = X, g.sb_k = ldx, g.sb_n = 1;    // B(k=p, n=k_in)
  g.C = dW, g.ldc = lddw, g.M = N, g.N = K, g.K = P;
  const int tiles = (int)(grid_for(N, BM) * grid_for(K, BN));
  int64_t splits = (4 * kSMs + tiles - 1) / tiles;
  const int64_t max_splits = (P + 255) / 256;
  if (splits > max_splits) splits = max_splits;
  if (splits < 1) splits = 1;
  g.k_chunk = ((P + splits - 1) / splits + BK - 1) / BK * BK;
  splits = (P + g.k_chunk - 1) / g.k_chunk;
  dim3 grid(grid_for(N, BM), grid_for(K, BN), (unsigned)splits);
  k_gemm<false, false, EPI_WGRAD><<<grid, NT, 0, st>>>(g);
  if (db) {
    const int64_t rpb = 2048;
    dim3 g2(grid_for(P, (int)rpb), grid_for(N, 32));
    k_colsum<<<g2, 256, 0, st>>>(dY, lddy, P, N, rpb, db);
  }
  return check_launch("b2n_linear_wgrad");
}

extern "C" int b2n_act_bwd(float* dY, int lddy, const float* Y, int ldy, int64_t P, int N, int act,
                           b2n_stream_t stream) {
  B2N_REQUIRE(P >= 0 && N > 0 && lddy >= N && ldy >= N && act >= 0 && act <= 2, "bad shape");
  if (P == 0 || act == B2N_ACT_NONE) return B2N_OK;
  B2N_REQUIRE(dY && Y, "null pointer");
  k_act_bwd<<<grid_for(P * N, 256), 256, 0, (cudaStream_t)stream>>>(dY, lddy, Y, ldy, P, N, act);
  return check_launch("b2n_act_bwd");
}

extern "C" int b2n_sigma_head_fwd(const float* h, int ldh, int64_t P, float* sigma, b2n_stream_t stream) {
  B2N_REQUIRE(P >= 0 && ldh > 0, "bad shape");
  if (P == 0) return B2N_OK;
  B2N_REQUIRE(h && sigma, "null pointer");
  k_sigma_head_fwd<<<grid_for(P, 256), 256, 0, (cudaStream_t)stream>>>(h, ldh, P, sigma);
  return check_launch("b2n_sigma_head_fwd");
}

extern "C" int b2n_sigma_head_bwd(const float* h, int ldh, int64_t P, const float* g_sigma, float* g_h, int ldg,
                                  b2n_stream_t stream) {
  B2N_REQUIRE(P >= 0 && ldh > 0 && ldg > 0, "bad shape");
  if (P == 0) return B2N_OK;
  B2N_REQUIRE(h && g_sigma && g_h, "null pointer");
  k_sigma_head_bwd<<<grid_for(P, 256), 256, 0, (cudaStream_t)stream>>>(h, ldh, P, g_sigma, g_h, ldg);
  return check_launch("b2n_sigma_head_bwd");
}
