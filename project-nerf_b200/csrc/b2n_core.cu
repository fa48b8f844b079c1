#include "b2n_common.cuh"

namespace b2n {
thread_local char g_err[512] = {0};
thread_local const int* g_active_rows = nullptr;
}

extern "C" int b2n_abi_version(void) { return B2N_ABI_VERSION; }
extern "C" const char* b2n_last_error(void) { return b2n::g_err; }
extern "C" int b2n_set_active_rows(const int* rows_device) {
  b2n::g_active_rows = rows_device;
  return B2N_OK;
}

// ---------------------------------------------------------------------------------------- fp32 rows -> zero-padded bf16 rows
// out[p, 0:kpad] = bf16(x[p, 0:width]) | 0   (kpad a multiple of 8 >= width).  The operand blocks of the tcgen05
// weight-gradient kernel (b2n_nerf_mlp_wgrad: x_bf16 / d_bf16) in ONE pass; torch needed a fill, a strided copy and a
// cast (3 kernels, 57 us for [262144, 63] against ~20 us here).
#include <cuda_bf16.h>
namespace b2n {
__global__ void k_pad_bf16(const float* __restrict__ x, int64_t P, int width, int kpad, __nv_bfloat16* __restrict__ out) {
  const int chunks = kpad >> 3;                                  // 16-byte output chunks per row
  const int64_t total = P * chunks;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t p = i / chunks;
    const int c = (int)(i - p * chunks) * 8;
    const float* src = x + p * width + c;
    float f[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) f[j] = (c + j < width) ? __ldg(src + j) : 0.f;
    __nv_bfloat162 h[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) h[j] = __floats2bfloat162_rn(f[2 * j], f[2 * j + 1]);
    *reinterpret_cast<uint4*>(out + p * kpad + c) = *reinterpret_cast<uint4*>(h);
  }
}
}  // namespace b2n

extern "C" int b2n_pad_bf16(const float* x, int64_t P, int width, int kpad, void* out, b2n_stream_t stream) {
  B2N_REQUIRE(P >= 0 && width > 0 && kpad >= width && (kpad & 7) == 0, "kpad must be a multiple of 8 and >= width");
  if (P == 0) return B2N_OK;
  B2N_REQUIRE(x && out, "null pointer");
  const int64_t total = P * (kpad >> 3);
  const unsigned grid = (unsigned)((total + 255) / 256 < (int64_t)b2n::kSMs * 16 ? (total + 255) / 256 : (int64_t)b2n::kSMs * 16);
  b2n::k_pad_bf16<<<grid, 256, 0, (cudaStream_t)stream>>>(x, P, width, kpad, (__nv_bfloat16*)out);
  return b2n::check_launch("b2n_pad_bf16");
}
