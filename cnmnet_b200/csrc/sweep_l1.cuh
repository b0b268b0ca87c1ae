// One output of the plane-sweep cost volume with the four bilinear taps read through L1 from the zero-bordered float4
// texels of cnm_pack_source.  Shared by cost_volume_kernel (cost_volume.cu) and the volume producers of the fused
// conv1.0 kernel (conv_halo.cu): both compile this one body, so their outputs agree bit for bit.
//
// Reference semantics: depthnet/depthNet_model.py:208-223 (per plane {mul, add, div, grid_sample(bilinear, zeros,
// align_corners=False), sub, abs, sum over the colour channels}).
#pragma once
#include "common.cuh"

namespace cnm {

struct SweepPixelL1 {  // per-pixel invariants of the plane loop
  float a0, a1, a2;    // Hm . [x, y, 1]^T   (KRKiUV)
  float k0, k1, k2;    // K_s t, k2 with the reference's + 1e-6 folded in
  float r0, r1, r2;    // reference pixel
  const float4* tp;    // texel (0, 0) of the pair's packed source
  int pitch;           // texels per packed row (W + 4)
  float xmax, ymax;    // W, H
};

__device__ __forceinline__ void sweep_l1_setup(SweepPixelL1& s, const float* __restrict__ h, int x, int y) {
  const float fx = (float)x, fy = (float)y;
  s.a0 = fmaf(__ldg(h + 0), fx, fmaf(__ldg(h + 1), fy, __ldg(h + 2)));
  s.a1 = fmaf(__ldg(h + 3), fx, fmaf(__ldg(h + 4), fy, __ldg(h + 5)));
  s.a2 = fmaf(__ldg(h + 6), fx, fmaf(__ldg(h + 7), fy, __ldg(h + 8)));
  s.k0 = __ldg(h + 9);
  s.k1 = __ldg(h + 10);
  s.k2 = __ldg(h + 11) + 1e-6f;
}

// sample position of one plane, clamped to the zero border: ix in [-1, W], iy in [-1, H]
__device__ __forceinline__ void sweep_l1_position(const SweepPixelL1& s, float z, float& ix, float& iy) {
  const float px = fmaf(s.a0, z, s.k0), py = fmaf(s.a1, z, s.k1);
  const float den = fmaf(s.a2, z, s.k2);  // p_z + 1e-6
  float rcp;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rcp) : "f"(den));
  rcp = fmaf(fmaf(-den, rcp, 1.0f), rcp, rcp);  // one Newton step: ~0.5 ulp
  // grid_sample un-normalisation with align_corners=False: ix = u - 0.5
  ix = fmaf(px, rcp, -0.5f);
  iy = fmaf(py, rcp, -0.5f);
  ix = fminf(fmaxf(ix, -1.0f), s.xmax);  // NaN -> -1 (zero border)
  iy = fminf(fmaxf(iy, -1.0f), s.ymax);
}

// bilinear blend of the four texels around (ix, iy) against the reference pixel; `fetch(x0, y0, t00, t10, t01, t11)`
// delivers the texels (x0, y0), (x0+1, y0), (x0, y0+1), (x0+1, y0+1) -- from global memory or from a staged copy
template <typename Fetch>
__device__ __forceinline__ float sweep_l1_blend(const SweepPixelL1& s, float ix, float iy, Fetch fetch) {
  const float x0f = floorf(ix), y0f = floorf(iy);
  const float wx1 = ix - x0f, wy1 = iy - y0f;
  const float wx0 = 1.0f - wx1, wy0 = 1.0f - wy1;
  float4 t00, t10, t01, t11;
  fetch((int)x0f, (int)y0f, t00, t10, t01, t11);
  const float w00 = wx0 * wy0, w10 = wx1 * wy0, w01 = wx0 * wy1, w11 = wx1 * wy1;
  const float c0 = fmaf(t11.x, w11, fmaf(t01.x, w01, fmaf(t10.x, w10, fmaf(t00.x, w00, -s.r0))));
  const float c1 = fmaf(t11.y, w11, fmaf(t01.y, w01, fmaf(t10.y, w10, fmaf(t00.y, w00, -s.r1))));
  const float c2 = fmaf(t11.z, w11, fmaf(t01.z, w01, fmaf(t10.z, w10, fmaf(t00.z, w00, -s.r2))));
  return fabsf(c0) + fabsf(c1) + fabsf(c2);
}

// the four taps through L1 from the packed source
struct SweepFetchGlobal {
  const float4* tp;
  int pitch;
  __device__ __forceinline__ void operator()(int x0, int y0, float4& t00, float4& t10, float4& t01, float4& t11) const {
    const float4* t = tp + y0 * pitch + x0;
    t00 = __ldg(t); t10 = __ldg(t + 1); t01 = __ldg(t + pitch); t11 = __ldg(t + pitch + 1);
  }
};

__device__ __forceinline__ float sweep_l1_sample(const SweepPixelL1& s, float z) {
  float ix, iy;
  sweep_l1_position(s, z, ix, iy);
  return sweep_l1_blend(s, ix, iy, SweepFetchGlobal{s.tp, s.pitch});
}

// z_d of plane d, evaluated in fp64 like the reference's Python double (depthNet_model.py:193-196)
__device__ __forceinline__ float sweep_plane_z(double idmin, double idstep, int d) {
  return (float)(1.0 / (idmin + (double)d * idstep));
}

}  // namespace cnm
