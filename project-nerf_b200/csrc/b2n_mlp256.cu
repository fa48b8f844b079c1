// 256-wide vanilla NeRF decoder (NeRFDecoder.forward, src/decoders.py:68-87) on the 5th-gen
// tensor cores: tcgen05.mma (cta_group::1, M=128, N=256|128, K=16, bf16 x bf16 -> fp32 in TMEM).
//
//   h = x; for i in 0..7: (i == 4: h = [h, x]); h = relu(W_i h + b_i)
//   sigma = relu(w_s h + b_s); feat = W_f h + b_f; hv = relu(W_v [feat, d] + b_v); rgb = sigmoid(W_c hv + b_c)
//
// One persistent CTA per SM processes PAIRS of 128-point tiles.  The activations of both tiles
// never leave the SM: they sit in shared memory as bf16 in the canonical K-major SWIZZLE_128B
// UMMA layout (4 k-blocks of 128 rows x 128 B per tile) and are the A operand of the next layer.
// The weights are pre-packed once per step (b2n_nerf_mlp_pack) into a stream of 32 KB chunks
// that are byte images of the B operand tiles ([N rows x 64 k] bf16, same swizzle), so the
// producer needs no tensor map: one elected thread issues one 1-D bulk copy (cp.async.bulk ->
// mbarrier complete_tx) per chunk into a 2-stage ring, and every chunk is used for BOTH tiles
// (256 rows per weight fetch).  Accumulators: tile 0 in TMEM columns 0..255, tile 1 in 256..511.
//
// Warp roles (10 warps): warp 0 = weight producer, warp 1 = TMEM allocator + MMA issuer (one
// elected thread), warps 2..9 = epilogue (4 per tile; warp_id % 4 selects the TMEM lane quarter):
// tcgen05.ld 32x32b.x32 -> +bias, ReLU -> bf16 -> swizzled st.shared (next layer's A) and, for
// training, a bf16 copy to HBM (saved activations for the backward / weight-gradient GEMMs).
// The 256->1 density head and the 128->3 colour head are dot products inside the epilogue.
//
// Every mbarrier wait is bounded; on time-out the CTA raises an abort flag, stores an error code
// and drains, so a protocol bug cannot hang the GPU.
#include <cuda_bf16.h>
#include "b2n_common.cuh"

namespace b2n {
namespace m256 {

constexpr int HID = 256;
constexpr int KBLK_BYTES = 128 * 128;          // one k-block of one tile: 128 rows x 64 bf16
constexpr int ACT_BYTES = 4 * KBLK_BYTES;      // 128 x 256 bf16
constexpr int STAGE_BYTES = 256 * 128;         // one weight chunk: 256 rows x 64 bf16
constexpr int N_STAGES = 2;
constexpr int OFF_ACT = 0;                     // [2 tiles][ACT_BYTES]
constexpr int OFF_AUX = OFF_ACT + 2 * ACT_BYTES;       // [2 tiles][KBLK_BYTES]   x_enc, later d_enc
constexpr int OFF_RING = OFF_AUX + 2 * KBLK_BYTES;     // [N_STAGES][STAGE_BYTES]
constexpr int OFF_VEC = OFF_RING + N_STAGES * STAGE_BYTES;   // 256 floats bias + 384 floats head weights
constexpr int OFF_BAR = OFF_VEC + (256 + 384) * 4;
constexpr int SMEM_BYTES = OFF_BAR + 128;
static_assert(SMEM_BYTES <= 232448, "exceeds 227 KB of shared memory");

constexpr int N_THREADS = 320;
constexpr int EPI_THREADS = 256;
constexpr int MAX_STEPS = 12, MAX_CHUNKS = 64;

enum { EPI_RELU = 0, EPI_RELU_SIGMA = 1, EPI_LINEAR = 2, EPI_VIEW_RGB = 3,
       EPI_B_LINEAR = 4, EPI_B_MASK = 5, EPI_B_MASK_SIGMA = 6 };

struct Step {
  int n_act;     // k-chunks taken from the activation buffer (0 or 4)
  int aux_k16;   // k16 sub-steps taken from the aux buffer (0 = none, 4 = 64 columns, 2 = 32 columns)
  int n;         // output width of the step (256 or 128)
  int epi;       // epilogue kind
  int bias_off;  // offset into the bias vector
  int save_slot; // index of the saved-activation plane (or -1)
};
struct Plan {
  int n_steps;
  Step s[MAX_STEPS];
};

// ---------------------------------------------------------------------------------------- PTX
__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0;
}
// bounded wait: false = aborted (either this wait timed out or another role raised the flag)
__device__ __forceinline__ bool mbar_wait(uint32_t bar, uint32_t parity, volatile int* abort_flag, int* err, int code) {
  for (uint32_t it = 0; it < (1u << 22); ++it) {
    if (mbar_try(bar, parity)) return true;
    if ((it & 255) == 255 && *abort_flag) return false;
  }
  *abort_flag = 1;
  atomicCAS(err, 0, code);
  return false;
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void proxy_fence() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_mma(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void tc_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
        "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
        "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// K-major SWIZZLE_128B shared-memory matrix descriptor: rows of 128 B, 8-row atoms 1024 B apart.
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) |
         ((uint64_t)2 << 61);
}
// kind::f16 instruction descriptor: D = fp32, A = B = bf16, both K-major, M = 128
__device__ __forceinline__ uint32_t umma_idesc(int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}
// byte offset of 16-byte chunk c (0..7) of row r inside a swizzled 128-row x 128-byte block
__device__ __forceinline__ uint32_t swz(int r, int c) { return (uint32_t)(r * 128 + ((c ^ (r & 7)) << 4)); }

__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}

// stage `width` fp32 columns of one input row as bf16 into an aux k-block (zero padded to 64;
// column `one_col` (if >= 0) is set to 1 -- unused here, biases are added in the epilogue)
__device__ __forceinline__ void stage_row(const float* __restrict__ src, int width, bool valid, unsigned char* blk, int r) {
#pragma unroll 1
  for (int c = 0; c < 8; ++c) {
    float f[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int col = 8 * c + j;
      f[j] = (valid && col < width) ? __ldg(src + col) : 0.f;
    }
    *reinterpret_cast<uint4*>(blk + swz(r, c)) =
        make_uint4(pack_bf16(f[0], f[1]), pack_bf16(f[2], f[3]), pack_bf16(f[4], f[5]), pack_bf16(f[6], f[7]));
  }
}

struct FwdArgs {
  const float* x_enc; int pos_dim;
  const float* d_enc; int dir_dim;
  const unsigned char* packed;      // weight chunk stream
  const float* bias;                // concatenated biases, Step::bias_off indexes it
  const float* w_sigma;
  const float* w_rgb;
  const float* head_bias;           // device float[4]: b_sigma, b_rgb[0..2]
  int64_t P;
  float* rgb; float* sigma;
  __nv_bfloat16* save;              // [n_slots][P][256] or nullptr
  int* err;
  // backward only
  const __nv_bfloat16* fwd_planes;  // [10][P][256] saved by the forward (H0..H7, feat, hv)
  const float* g_rgb; const float* g_sigma;   // [P,3], [P]
  const float* rgb_out; const float* sigma_out;
  float* dz_small;                  // [P,4]: d(pre-sigmoid rgb)[3], d(pre-relu sigma)
  Plan plan;
};

template <bool BWD>
__global__ void __launch_bounds__(N_THREADS, 1) k_mlp256(const FwdArgs a) {
  extern __shared__ __align__(1024) unsigned char smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
  const uint32_t bar_full = s32(bars + 0);     // [N_STAGES]
  const uint32_t bar_empty = s32(bars + 2);    // [N_STAGES]
  const uint32_t bar_acc = s32(bars + 4);      // MMA -> epilogue: accumulators of the step complete
  const uint32_t bar_act = s32(bars + 5);      // epilogue -> MMA: A operands written, accumulators drained
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 6);
  volatile int* abort_flag = reinterpret_cast<volatile int*>(bars + 7);
  float* vec = reinterpret_cast<float*>(smem + OFF_VEC);

  if ((s32(smem) & 1023u) != 0) {  // SWIZZLE_128B operands need 1024-byte aligned tiles
    if (threadIdx.x == 0) atomicCAS(a.err, 0, 100);
    return;
  }
  if (threadIdx.x == 0) {
    for (int i = 0; i < N_STAGES; ++i) mbar_init(bar_full + 8 * i, 1), mbar_init(bar_empty + 8 * i, 1);
    mbar_init(bar_acc, 1);
    mbar_init(bar_act, EPI_THREADS);
    *abort_flag = 0;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s32(tmem_slot)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const int64_t n_pairs = (a.P + 255) / 256;
  const Plan& plan = a.plan;

  if (warp == 0) {
    // ================================ weight producer ================================
    if (lane == 0) {
      uint32_t use = 0;  // running chunk counter (ring position)
      for (int64_t pair = blockIdx.x; pair < n_pairs; pair += gridDim.x) {
        const unsigned char* src = a.packed;
        for (int s = 0; s < plan.n_steps; ++s) {
          const int nch = plan.s[s].n_act + (plan.s[s].aux_k16 ? 1 : 0);
          const uint32_t bytes = (uint32_t)plan.s[s].n * 128u;
          for (int c = 0; c < nch; ++c, ++use) {
            const uint32_t st = use % N_STAGES, ph = (use / N_STAGES) & 1;
            if (!mbar_wait(bar_empty + 8 * st, ph ^ 1, abort_flag, a.err, 1)) goto prod_done;
            mbar_expect_tx(bar_full + 8 * st, bytes);
            bulk_g2s(s32(smem + OFF_RING + st * STAGE_BYTES), src, bytes, bar_full + 8 * st);
            src += bytes;
          }
        }
      }
    }
  prod_done:;
  } else if (warp == 1) {
    // ================================ MMA issuer ================================
    if (lane == 0) {
      uint32_t use = 0, act_phase = 0;
      for (int64_t pair = blockIdx.x; pair < n_pairs; pair += gridDim.x) {
        for (int s = 0; s < plan.n_steps; ++s) {
          const Step& sp = plan.s[s];
          if (!mbar_wait(bar_act, act_phase & 1, abort_flag, a.err, 2)) goto mma_done;
          ++act_phase;
          tc_fence_after();
          const uint32_t idesc = umma_idesc(sp.n);
          const int nch = sp.n_act + (sp.aux_k16 ? 1 : 0);
          for (int c = 0; c < nch; ++c, ++use) {
            const uint32_t st = use % N_STAGES, ph = (use / N_STAGES) & 1;
            if (!mbar_wait(bar_full + 8 * st, ph, abort_flag, a.err, 3)) goto mma_done;
            tc_fence_after();
            const uint64_t bdesc = umma_desc(s32(smem + OFF_RING + st * STAGE_BYTES));
            const bool from_aux = c >= sp.n_act;
            const int nk = from_aux ? sp.aux_k16 : 4;
#pragma unroll 1
            for (int t = 0; t < 2; ++t) {
              const uint32_t abase = from_aux ? s32(smem + OFF_AUX + t * KBLK_BYTES)
                                              : s32(smem + OFF_ACT + t * ACT_BYTES + c * KBLK_BYTES);
              const uint64_t adesc = umma_desc(abase);
              for (int k = 0; k < nk; ++k)
                tc_mma(tmem + t * 256, adesc + 2 * k, bdesc + 2 * k, idesc, (c > 0 || k > 0) ? 1u : 0u);
            }
            tc_commit(bar_empty + 8 * st);   // frees the ring slot once these MMAs have read it
          }
          tc_commit(bar_acc);
        }
      }
    }
  mma_done:;
  } else {
    // ================================ epilogue (8 warps) ================================
    const int e = threadIdx.x - 64;            // 0..255
    const int t = e >> 7;                      // tile of this thread
    const int q = warp & 3;                    // TMEM lane quarter this warp may touch
    const int r = 32 * q + lane;               // row inside the tile
    unsigned char* act = smem + OFF_ACT + t * ACT_BYTES;
    unsigned char* aux = smem + OFF_AUX + t * KBLK_BYTES;
    const uint32_t trow = tmem + ((uint32_t)(32 * q) << 16) + t * 256;
    uint32_t acc_phase = 0;
    if (BWD) {  // head weights are needed by every pair: stage them once
      vec[e] = __ldg(a.w_sigma + e);
      vec[256 + e] = __ldg(a.w_rgb + e);
      if (e < 128) vec[512 + e] = __ldg(a.w_rgb + 256 + e);
      asm volatile("bar.sync 1, 256;" ::: "memory");
    }
    for (int64_t pair = blockIdx.x; pair < n_pairs; pair += gridDim.x) {
      const int64_t p = pair * 256 + t * 128 + r;
      const bool valid = p < a.P;
      float sig_acc = 0.f;      // fwd: density dot product; bwd: d(pre-relu sigma) of this row
      if (!BWD) {
        // ---- pre-step: stage the encoded position as the first A operand
        stage_row(a.x_enc + p * a.pos_dim, a.pos_dim, valid, aux, r);
      } else {
        // ---- pre-step: colour head backward -> dZ_view (128 wide) as the first A operand
        float dzr[3] = {0.f, 0.f, 0.f};
        if (valid) {
#pragma unroll
          for (int j = 0; j < 3; ++j) {
            const float y = __ldg(a.rgb_out + 3 * p + j);
            dzr[j] = __ldg(a.g_rgb + 3 * p + j) * y * (1.f - y);
          }
          sig_acc = __ldg(a.sigma_out + p) > 0.f ? __ldg(a.g_sigma + p) : 0.f;
          *reinterpret_cast<float4*>(a.dz_small + 4 * p) = make_float4(dzr[0], dzr[1], dzr[2], sig_acc);
        }
        const __nv_bfloat16* hv = a.fwd_planes + ((size_t)9 * a.P + (valid ? p : 0)) * HID;
        __nv_bfloat16* srow = (a.save && valid) ? a.save + ((size_t)0 * a.P + p) * HID : nullptr;
#pragma unroll 1
        for (int c = 0; c < 16; ++c) {
          uint4 hraw = valid ? __ldcs(reinterpret_cast<const uint4*>(hv + 8 * c)) : make_uint4(0, 0, 0, 0);
          const __nv_bfloat16* hb = reinterpret_cast<const __nv_bfloat16*>(&hraw);
          float f[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const int col = 8 * c + j;
            const float g = dzr[0] * vec[256 + col] + dzr[1] * vec[256 + 128 + col] + dzr[2] * vec[256 + 256 + col];
            f[j] = (__bfloat162float(hb[j]) > 0.f) ? g : 0.f;
          }
          const uint4 pk = make_uint4(pack_bf16(f[0], f[1]), pack_bf16(f[2], f[3]), pack_bf16(f[4], f[5]), pack_bf16(f[6], f[7]));
          *reinterpret_cast<uint4*>(act + ((8 * c) >> 6) * KBLK_BYTES + swz(r, ((8 * c) & 63) >> 3)) = pk;
          if (srow) __stcs(reinterpret_cast<uint4*>(srow + 8 * c), pk);
        }
      }
      proxy_fence();
      mbar_arrive(bar_act);
      for (int s = 0; s < plan.n_steps; ++s) {
        const Step& sp = plan.s[s];
        if (!BWD) {
          // stage this step's bias (and the head weights) for broadcast reads
          asm volatile("bar.sync 1, 256;" ::: "memory");
          if (e < sp.n) vec[e] = __ldg(a.bias + sp.bias_off + e);
          if (sp.epi == EPI_RELU_SIGMA) vec[256 + e] = __ldg(a.w_sigma + e);
          if (sp.epi == EPI_VIEW_RGB) {
            vec[256 + e] = __ldg(a.w_rgb + e);
            if (e < 128) vec[512 + e] = __ldg(a.w_rgb + 256 + e);
          }
          asm volatile("bar.sync 1, 256;" ::: "memory");
        }
        if (!mbar_wait(bar_acc, acc_phase & 1, abort_flag, a.err, 4)) goto epi_done;
        ++acc_phase;
        tc_fence_after();
        float rgb_acc[3] = {0.f, 0.f, 0.f};
        const bool relu = sp.epi != EPI_LINEAR;
        __nv_bfloat16* save_row = (a.save && sp.save_slot >= 0 && valid)
                                      ? a.save + ((size_t)sp.save_slot * a.P + p) * HID : nullptr;
        // bwd: the forward activation whose ReLU gates this gradient (plane index in bias_off)
        const __nv_bfloat16* mask_row = (BWD && sp.epi != EPI_B_LINEAR)
                                            ? a.fwd_planes + ((size_t)sp.bias_off * a.P + (valid ? p : 0)) * HID : nullptr;
        const bool last = (s + 1 == plan.n_steps);
#pragma unroll 1
        for (int cb = 0; cb < sp.n / 32; ++cb) {
          uint32_t v[32];
          tc_ld32(trow + 32 * cb, v);
          float f[32];
          if (!BWD) {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              float x = __uint_as_float(v[j]) + vec[32 * cb + j];
              f[j] = relu ? fmaxf(x, 0.f) : x;
            }
            if (sp.epi == EPI_RELU_SIGMA) {
#pragma unroll
              for (int j = 0; j < 32; ++j) sig_acc = fmaf(f[j], vec[256 + 32 * cb + j], sig_acc);
            } else if (sp.epi == EPI_VIEW_RGB) {
#pragma unroll
              for (int j = 0; j < 32; ++j) {
                rgb_acc[0] = fmaf(f[j], vec[256 + 32 * cb + j], rgb_acc[0]);
                rgb_acc[1] = fmaf(f[j], vec[256 + 128 + 32 * cb + j], rgb_acc[1]);
                rgb_acc[2] = fmaf(f[j], vec[256 + 256 + 32 * cb + j], rgb_acc[2]);
              }
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
            if (sp.epi == EPI_B_MASK_SIGMA) {   // + d(sigma_pre) * w_sigma  (density head, rank-1)
#pragma unroll
              for (int j = 0; j < 32; ++j) f[j] = fmaf(sig_acc, vec[32 * cb + j], f[j]);
            }
            if (mask_row) {
#pragma unroll
              for (int c4 = 0; c4 < 4; ++c4) {
                uint4 hraw = valid ? __ldcs(reinterpret_cast<const uint4*>(mask_row + 32 * cb + 8 * c4)) : make_uint4(0, 0, 0, 0);
                const __nv_bfloat16* hb = reinterpret_cast<const __nv_bfloat16*>(&hraw);
#pragma unroll
                for (int j = 0; j < 8; ++j)
                  if (!(__bfloat162float(hb[j]) > 0.f)) f[8 * c4 + j] = 0.f;
              }
            }
          }
#pragma unroll
          for (int c4 = 0; c4 < 4; ++c4) {
            const uint4 pk = make_uint4(pack_bf16(f[8 * c4], f[8 * c4 + 1]), pack_bf16(f[8 * c4 + 2], f[8 * c4 + 3]),
                                        pack_bf16(f[8 * c4 + 4], f[8 * c4 + 5]), pack_bf16(f[8 * c4 + 6], f[8 * c4 + 7]));
            const int col = 32 * cb + 8 * c4;            // first column of this 16-byte chunk
            if (sp.epi != EPI_VIEW_RGB && !(BWD && last))
              *reinterpret_cast<uint4*>(act + (col >> 6) * KBLK_BYTES + swz(r, (col & 63) >> 3)) = pk;
            if (save_row) __stcs(reinterpret_cast<uint4*>(save_row + col), pk);
          }
        }
        if (!BWD) {
          if (sp.epi == EPI_RELU_SIGMA && valid) a.sigma[p] = fmaxf(sig_acc + __ldg(a.head_bias), 0.f);
          if (sp.epi == EPI_VIEW_RGB && valid) {
#pragma unroll
            for (int j = 0; j < 3; ++j) a.rgb[3 * p + j] = 1.f / (1.f + expf(-(rgb_acc[j] + __ldg(a.head_bias + 1 + j))));
          }
          // the x block is dead after the skip layer (the step that consumed both act and aux):
          // re-use it for the encoded view direction of the view layer
          if (sp.n_act > 0 && sp.aux_k16 > 0 && sp.epi == EPI_RELU)
            stage_row(a.d_enc + p * a.dir_dim, a.dir_dim, valid, aux, r);
        }
        tc_fence_before();
        proxy_fence();
        if (!last) mbar_arrive(bar_act);
      }
    }
  epi_done:;
  }
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
  }
}

// ---------------------------------------------------------------------------------------- weight packing
struct PackChunk {
  const float* W;   // source matrix, row-major, leading dimension ldw
  // forward (trans = 0): tile row n <- source row n (n < n_real), tile k <- source column k0 + k (< k_real)
  // backward (trans = 1): tile row n <- source column n (n < n_real), tile k <- source row k0 + k (< k_real)
  int ldw, n_real, n_pad, k0, k_real, trans;
  int64_t dst_off;  // byte offset in the packed stream
};
struct PackArgs {
  int n_chunks;
  PackChunk c[MAX_CHUNKS];
  unsigned char* dst;
};

__global__ void k_mlp256_pack(const PackArgs a) {
  const PackChunk& c = a.c[blockIdx.x];
  unsigned char* dst = a.dst + c.dst_off;
  for (int i = threadIdx.x; i < c.n_pad * 8; i += blockDim.x) {
    const int n = i >> 3, ch = i & 7;
    float f[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int k = c.k0 + 8 * ch + j;
      const bool ok = n < c.n_real && k < c.k_real;
      f[j] = ok ? __ldg(c.trans ? c.W + (size_t)k * c.ldw + n : c.W + (size_t)n * c.ldw + k) : 0.f;
    }
    *reinterpret_cast<uint4*>(dst + swz(n, ch)) =
        make_uint4(pack_bf16(f[0], f[1]), pack_bf16(f[2], f[3]), pack_bf16(f[4], f[5]), pack_bf16(f[6], f[7]));
  }
}

}  // namespace m256
}  // namespace b2n

using namespace b2n;
using namespace b2n::m256;

// Forward plan of the reference architecture (8 x 256, skip at 4, view 128); pos_dim, dir_dim <= 64.
static void build_fwd_plan(Plan* pl) {
  int n = 0;
  auto add = [&](int n_act, int aux_k16, int width, int epi, int bias_off, int slot) {
    pl->s[n++] = Step{n_act, aux_k16, width, epi, bias_off, slot};
  };
  add(0, 4, 256, EPI_RELU, 0, 0);
  add(4, 0, 256, EPI_RELU, 256, 1);
  add(4, 0, 256, EPI_RELU, 512, 2);
  add(4, 0, 256, EPI_RELU, 768, 3);
  add(4, 4, 256, EPI_RELU, 1024, 4);        // skip layer: [h, x]
  add(4, 0, 256, EPI_RELU, 1280, 5);
  add(4, 0, 256, EPI_RELU, 1536, 6);
  add(4, 0, 256, EPI_RELU_SIGMA, 1792, 7);  // + density head
  add(4, 0, 256, EPI_LINEAR, 2048, 8);      // feature layer (no activation)
  add(4, 2, 128, EPI_VIEW_RGB, 2304, 9);    // view layer [feat, d] + colour head
  pl->n_steps = n;
}

extern "C" size_t b2n_nerf_mlp_packed_bytes(void) {
  // 1 + 3*4 + 5 + 3*4 + 4 chunks of 256 rows, 5 chunks of 128 rows
  return (size_t)(1 + 12 + 5 + 12 + 4) * 256 * 128 + (size_t)5 * 128 * 128;
}

// weights: the 12 nn.Linear weight matrices of NeRFDecoder in state_dict order:
// pts_layers[0..7], sigma_layer (unused here), feature_layer, view_layer, rgb_layer (unused here)
extern "C" int b2n_nerf_mlp_pack(const float* const* pts_w, const float* feature_w, const float* view_w, int pos_dim,
                                 int dir_dim, void* packed, b2n_stream_t stream) {
  B2N_REQUIRE(pts_w && feature_w && view_w && packed, "null pointer");
  B2N_REQUIRE(pos_dim > 0 && pos_dim <= 64 && dir_dim > 0 && dir_dim <= 32, "pos_dim <= 64 and dir_dim <= 32 required");
  PackArgs pa{};
  int n = 0;
  int64_t off = 0;
  auto add = [&](const float* W, int ldw, int n_real, int n_pad, int k0, int k_real) {
    pa.c[n++] = PackChunk{W, ldw, n_real, n_pad, k0, k_real, 0, off};
    off += (int64_t)n_pad * 128;
  };
  for (int l = 0; l < 8; ++l) {
    B2N_REQUIRE(pts_w[l], "null weight");
    if (l == 0) {
      add(pts_w[0], pos_dim, 256, 256, 0, pos_dim);
    } else {
      const int ld = (l == 4) ? 256 + pos_dim : 256;
      for (int c = 0; c < 4; ++c) add(pts_w[l], ld, 256, 256, 64 * c, 256);
      if (l == 4) add(pts_w[4], ld, 256, 256, 256, ld);
    }
  }
  for (int c = 0; c < 4; ++c) add(feature_w, 256, 256, 256, 64 * c, 256);
  for (int c = 0; c < 4; ++c) add(view_w, 256 + dir_dim, 128, 128, 64 * c, 256);
  add(view_w, 256 + dir_dim, 128, 128, 256, 256 + dir_dim);
  pa.n_chunks = n;
  pa.dst = (unsigned char*)packed;
  B2N_REQUIRE((size_t)off == b2n_nerf_mlp_packed_bytes(), "internal: packed size mismatch");
  k_mlp256_pack<<<n, 256, 0, (cudaStream_t)stream>>>(pa);
  return check_launch("b2n_nerf_mlp_pack");
}

extern "C" int b2n_nerf_mlp_fwd(const float* x_enc, int pos_dim, const float* d_enc, int dir_dim, const void* packed,
                                const float* bias, const float* w_sigma, const float* w_rgb, const float* head_bias,
                                int64_t P, float* rgb, float* sigma, void* save, int* err_flag, b2n_stream_t stream) {
  B2N_REQUIRE(P >= 0, "negative size");
  if (P == 0) return B2N_OK;
  B2N_REQUIRE(x_enc && d_enc && packed && bias && w_sigma && w_rgb && head_bias && rgb && sigma && err_flag,
              "null pointer");
  B2N_REQUIRE(pos_dim > 0 && pos_dim <= 64 && dir_dim > 0 && dir_dim <= 32, "pos_dim <= 64 and dir_dim <= 32 required");
  FwdArgs a{};
  a.x_enc = x_enc, a.pos_dim = pos_dim, a.d_enc = d_enc, a.dir_dim = dir_dim;
  a.packed = (const unsigned char*)packed, a.bias = bias, a.w_sigma = w_sigma, a.w_rgb = w_rgb;
  a.head_bias = head_bias;
  a.P = P, a.rgb = rgb, a.sigma = sigma, a.save = (__nv_bfloat16*)save, a.err = err_flag;
  build_fwd_plan(&a.plan);
  cudaFuncSetAttribute(k_mlp256<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
  const int64_t n_pairs = (P + 255) / 256;
  const int grid = (int)(n_pairs < kSMs ? n_pairs : kSMs);
  k_mlp256<false><<<grid, N_THREADS, SMEM_BYTES, (cudaStream_t)stream>>>(a);
  return check_launch("b2n_nerf_mlp_fwd");
}

// ---------------------------------------------------------------------------------------- backward (data gradients)
// Chain (all on the tensor cores, M = 128 x 2 tiles, N = 256):
//   pre : dZ_view = (d rgb_pre * W_rgb) (.) [hv > 0]                       (epilogue warps, 128 wide)
//   B1  : d feat  = dZ_view  * W_view[:, :256]            K = 128   -> dZ_feat (linear)
//   B2  : d h7    = dZ_feat  * W_feat + d sigma_pre * w_sigma, gated by H7  -> dZ7
//   B3..: d h_{l-1} = dZ_l * W_l[:, :256], gated by H_{l-1}    (l = 7 .. 1)  -> dZ_{l-1}
// Every dZ plane is written to HBM in bf16: the weight gradients dW_l = dZ_l^T In_l are plain
// [256 x P] x [P x 256] GEMMs done outside this kernel.  Step::bias_off holds the index of the
// forward plane that gates the step.
static void build_bwd_plan(Plan* pl) {
  int n = 0;
  auto add = [&](int n_act, int epi, int gate_plane, int slot) { pl->s[n++] = Step{n_act, 0, 256, epi, gate_plane, slot}; };
  add(2, EPI_B_LINEAR, 0, 1);        // d feat
  add(4, EPI_B_MASK_SIGMA, 7, 2);    // dZ7
  for (int l = 7; l >= 1; --l) add(4, EPI_B_MASK, l - 1, 2 + (8 - l));   // dZ6 .. dZ0 -> slots 3..9
  pl->n_steps = n;
}

extern "C" size_t b2n_nerf_mlp_packed_bwd_bytes(void) { return (size_t)(2 + 4 * 8) * 256 * 128; }

extern "C" int b2n_nerf_mlp_pack_bwd(const float* const* pts_w, const float* feature_w, const float* view_w, int pos_dim,
                                     int dir_dim, void* packed, b2n_stream_t stream) {
  B2N_REQUIRE(pts_w && feature_w && view_w && packed, "null pointer");
  B2N_REQUIRE(pos_dim > 0 && pos_dim <= 64 && dir_dim > 0 && dir_dim <= 32, "pos_dim <= 64 and dir_dim <= 32 required");
  PackArgs pa{};
  int n = 0;
  int64_t off = 0;
  // tile rows = input index of the layer (first 256 inputs), tile k = output index of the layer
  auto add = [&](const float* W, int ldw, int n_out, int k0) {
    pa.c[n++] = PackChunk{W, ldw, 256, 256, k0, n_out, 1, off};
    off += (int64_t)256 * 128;
  };
  for (int c = 0; c < 2; ++c) add(view_w, 256 + dir_dim, 128, 64 * c);
  for (int c = 0; c < 4; ++c) add(feature_w, 256, 256, 64 * c);
  for (int l = 7; l >= 1; --l) {
    B2N_REQUIRE(pts_w[l], "null weight");
    const int ld = (l == 4) ? 256 + pos_dim : 256;
    for (int c = 0; c < 4; ++c) add(pts_w[l], ld, 256, 64 * c);
  }
  pa.n_chunks = n;
  pa.dst = (unsigned char*)packed;
  B2N_REQUIRE((size_t)off == b2n_nerf_mlp_packed_bwd_bytes(), "internal: packed size mismatch");
  k_mlp256_pack<<<n, 256, 0, (cudaStream_t)stream>>>(pa);
  return check_launch("b2n_nerf_mlp_pack_bwd");
}

extern "C" int b2n_nerf_mlp_bwd(const void* packed_bwd, const float* w_sigma, const float* w_rgb, const void* fwd_planes,
                                const float* rgb, const float* sigma, const float* g_rgb, const float* g_sigma,
                                int64_t P, void* dz_planes, float* dz_small, int* err_flag, b2n_stream_t stream) {
  B2N_REQUIRE(P >= 0, "negative size");
  if (P == 0) return B2N_OK;
  B2N_REQUIRE(packed_bwd && w_sigma && w_rgb && fwd_planes && rgb && sigma && g_rgb && g_sigma && dz_planes &&
                  dz_small && err_flag, "null pointer");
  FwdArgs a{};
  a.packed = (const unsigned char*)packed_bwd, a.w_sigma = w_sigma, a.w_rgb = w_rgb;
  a.fwd_planes = (const __nv_bfloat16*)fwd_planes, a.rgb_out = rgb, a.sigma_out = sigma, a.g_rgb = g_rgb;
  a.g_sigma = g_sigma, a.P = P, a.save = (__nv_bfloat16*)dz_planes, a.dz_small = dz_small, a.err = err_flag;
  build_bwd_plan(&a.plan);
  cudaFuncSetAttribute(k_mlp256<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
  const int64_t n_pairs = (P + 255) / 256;
  const int grid = (int)(n_pairs < kSMs ? n_pairs : kSMs);
  k_mlp256<true><<<grid, N_THREADS, SMEM_BYTES, (cudaStream_t)stream>>>(a);
  return check_launch("b2n_nerf_mlp_bwd");
}
