// Measurement aids (not on the product path): the L2 gather peak that SURVEY.md 8d asks the
// hash-encode kernels to be judged against.  MEASURED_PEAKS.json only has the HBM copy and the
// bf16 GEMM peaks; a multiresolution hash lookup is neither -- it is 8-byte gathers at random
// addresses of a table that lives in L2.
#include "b2n_common.cuh"

namespace b2n {

// every thread performs `per_thread` dependent-free random float2 loads (8 in flight) from `table`
__global__ void __launch_bounds__(256) k_gather_bench(const float2* __restrict__ table, uint32_t n_entries_mask,
                                                      int per_thread, float* __restrict__ sink) {
  uint32_t s = (blockIdx.x * blockDim.x + threadIdx.x) * 2654435761u + 12345u;
  float acc = 0.f;
  for (int i = 0; i < per_thread; i += 8) {
    float2 v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      s = s * 1664525u + 1013904223u;
      v[j] = __ldg(table + ((s >> 4) & n_entries_mask));
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) acc += v[j].x + v[j].y;
  }
  if (acc == 1.2345e-30f) *sink = acc;  // keep the loads alive
}

}  // namespace b2n

using namespace b2n;

// table: float2[n_entries] with n_entries a power of two; launches `blocks` CTAs of 256 threads.
extern "C" int b2n_debug_gather_bench(const float* table, int64_t n_entries, int blocks, int per_thread, float* sink,
                                      b2n_stream_t stream) {
  B2N_REQUIRE(table && sink && n_entries > 0 && (n_entries & (n_entries - 1)) == 0 && blocks > 0 && per_thread > 0 &&
                  per_thread % 8 == 0, "bad arguments");
  k_gather_bench<<<blocks, 256, 0, (cudaStream_t)stream>>>((const float2*)table, (uint32_t)(n_entries - 1), per_thread, sink);
  return check_launch("b2n_debug_gather_bench");
}
