// Multiresolution hash-grid device helpers (tinycudann HashGrid semantics) shared by the
// encode kernels (b2n_encode.cu) and the fused hash-encode + Instant-decoder forward (b2n_mlp64.cu).
#pragma once
#include "b2n_common.cuh"

namespace b2n {

// ----------------------------------------------------------------------------- hash grid
struct Levels {
  b2n_hash_level l[B2N_MAX_LEVELS];
};

__device__ __forceinline__ uint32_t corner_entry(const b2n_hash_level& lv, uint32_t cx, uint32_t cy, uint32_t cz) {
  uint32_t idx;
  if (lv.hashed) {
    idx = cx ^ (cy * 2654435761u) ^ (cz * 805459861u);
    idx &= lv.size - 1u;  // a hashed level always has size == 2^log2_hashmap_size
  } else {
    idx = cx + cy * lv.res + cz * lv.res * lv.res;
    if (idx >= lv.size) idx %= lv.size;  // only the +1 corners at x01 == 1 wrap
  }
  return lv.offset + idx;
}

// world coordinate -> unit cube: clamp((x + bound) / (2 bound), 0, 1)   (embeddings.py:86-87)
__device__ __forceinline__ float to_unit(float x, float bound, float two_bound, bool* inside) {
  // bound == 0: the input already lives in the unit cube (tcnn-style ``encoding(x01)`` call)
  const float xn = bound > 0.f ? __fdiv_rn(__fadd_rn(x, bound), two_bound) : x;
  *inside = (xn >= 0.f) && (xn <= 1.f);  // clamp passes gradient on the closed interval
  return fminf(fmaxf(xn, 0.f), 1.f);
}

struct Cell {
  uint32_t g[3];
  float w[3];
};

__device__ __forceinline__ Cell locate(const float x01[3], float scale) {
  Cell c;
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    const float pos = __fadd_rn(__fmul_rn(x01[d], scale), 0.5f);
    const float fl = floorf(pos);
    c.g[d] = (uint32_t)(int)fl;
    c.w[d] = pos - fl;
  }
  return c;
}

template <int F>
__device__ __forceinline__ void load_feat(const float* __restrict__ table, uint32_t e, float v[F]) {
  if (F == 2) {
    const float2 t = __ldg(reinterpret_cast<const float2*>(table) + e);
    v[0] = t.x, v[1] = t.y;
  } else if (F == 4) {
    const float4 t = __ldg(reinterpret_cast<const float4*>(table) + e);
    v[0] = t.x, v[1] = t.y, v[2] = t.z, v[3] = t.w;
  } else {
#pragma unroll
    for (int f = 0; f < F; ++f) v[f] = __ldg(table + (size_t)e * F + f);
  }
}

// One thread per (point, level) ITEM, item = p * L + level: consecutive lanes hold consecutive
// levels of the same point, so the feature row of a point (L*F floats) is written / read by
// adjacent lanes as one contiguous segment (full 32 B sectors, streaming cache hints) while the
// 8 gathers (or 8 reductions) of an item stay independent and in flight together.  The level
// table sits in shared memory because lanes index it divergently.  Streaming (.cs) hints on the
// per-point traffic keep the L2 for what is re-used: the table (forward) and its gradient
// (backward, where every red.global that misses L2 costs a 32 B DRAM fill).
struct SmemLevels {
  b2n_hash_level l[B2N_MAX_LEVELS];
};

__device__ __forceinline__ void stage_levels(const Levels& lv, int nl, SmemLevels* s) {
  if (threadIdx.x < nl) s->l[threadIdx.x] = lv.l[threadIdx.x];
  __syncthreads();
}

}  // namespace b2n
