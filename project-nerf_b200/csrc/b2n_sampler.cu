// GPU-resident training-ray sampler: replaces the CPU fancy-index gather + three
// H2D copies of BlenderDataset / DynamicDataset.sample_random_rays (src/dataset.py:147-171,
// :268-294).  The image stack lives in HBM as 8-bit RGBA (lossless: the PNGs are 8-bit and the
// reference's float conversion is v / 255), poses as fp32 4x4.  One thread per ray: pinhole
// direction ((x - W/2)/f, -(y - H/2)/f, -1), rotate by c2w[:3,:3], normalise, origin = c2w[:3,3]
// * scene_scale, target = rgba / 255, time = times[img].
#include "b2n_common.cuh"

namespace b2n {

__global__ void k_sample_rays(const float* __restrict__ poses, const uint8_t* __restrict__ images,
                              const float* __restrict__ times, const int64_t* __restrict__ img_idx,
                              const int64_t* __restrict__ pix_y, const int64_t* __restrict__ pix_x, int64_t B, int V,
                              int H, int W, float focal, float scene_scale, float* __restrict__ rays_o,
                              float* __restrict__ rays_d, float* __restrict__ target, float* __restrict__ t_out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B) return;
  const int64_t v = img_idx[i], y = pix_y[i], x = pix_x[i];
  const float* M = poses + 16 * v;
  // same fp32 expression order as the reference: (pix - W*0.5) / focal
  const float dx = __fdiv_rn(__fsub_rn((float)x, (float)W * 0.5f), focal);
  const float dy = -__fdiv_rn(__fsub_rn((float)y, (float)H * 0.5f), focal);
  const float dz = -1.0f;
  float d[3];
#pragma unroll
  for (int r = 0; r < 3; ++r) d[r] = M[4 * r] * dx + M[4 * r + 1] * dy + M[4 * r + 2] * dz;
  const float n = sqrtf(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    rays_d[3 * i + r] = __fdiv_rn(d[r], n);
    const float o = M[4 * r + 3];
    rays_o[3 * i + r] = scene_scale != 1.0f ? __fmul_rn(o, scene_scale) : o;
  }
  const uchar4 px = *reinterpret_cast<const uchar4*>(images + (((size_t)v * H + y) * W + x) * 4);
  float4 t;
  t.x = __fdiv_rn((float)px.x, 255.0f), t.y = __fdiv_rn((float)px.y, 255.0f);
  t.z = __fdiv_rn((float)px.z, 255.0f), t.w = __fdiv_rn((float)px.w, 255.0f);
  *reinterpret_cast<float4*>(target + 4 * i) = t;
  if (t_out) t_out[i] = times[v];
}

}  // namespace b2n

using namespace b2n;

extern "C" int b2n_sample_rays(const float* poses, const uint8_t* images_rgba8, const float* times,
                               const int64_t* img_idx, const int64_t* pix_y, const int64_t* pix_x, int64_t B, int V,
                               int H, int W, float focal, float scene_scale, float* rays_o, float* rays_d,
                               float* target_rgba, float* t_out, b2n_stream_t stream) {
  B2N_REQUIRE(B >= 0 && V > 0 && H > 0 && W > 0 && focal > 0.f, "bad shape");
  if (B == 0) return B2N_OK;
  B2N_REQUIRE(poses && images_rgba8 && img_idx && pix_y && pix_x && rays_o && rays_d && target_rgba, "null pointer");
  B2N_REQUIRE(!t_out || times, "t_out needs times");
  k_sample_rays<<<grid_for(B, 256), 256, 0, (cudaStream_t)stream>>>(poses, images_rgba8, times, img_idx, pix_y, pix_x, B,
                                                                   V, H, W, focal, scene_scale, rays_o, rays_d,
                                                                   target_rgba, t_out);
  return check_launch("b2n_sample_rays");
}
