// K1: camera pre-processing, source repack and the fused plane-sweep cost volume.
//
// Reference semantics: depthnet/depth_util.py:13-56 (process_camera_parameters),
// depthnet/depthNet_model.py:185-224 (getVolume: 64 x {mul, add, div, grid_sample, abs, sum}).
// Here: one launch for all (view, batch) pairs and all planes; the warped volume never exists.
//
// Data layout in HBM
//   ref      [B,3,H,W] fp32 (public NCHW), read once per (pair, plane group)
//   texels   [P][H+3][W+4] float4 = (r,g,b,0): source views repacked pixel-interleaved with a
//            zero border at row/col -1 and two trailing rows/cols, so that after clamping the
//            sample position to [-1, W] x [-1, H] every bilinear tap is an unpredicated 16 B
//            load and grid_sample's zeros padding falls out of the border texels.
//   out      NCHW fp32 [P,D,H,W] (parity layout) or act [P,H,W,D] bf16/fp32 (conv1.0 operand)
//
// Tried and rejected in round 1 (B200 measurements, C2: 256x192x64, 2 views, batch 5, NCHW fp32):
//   one pixel per thread, 4 x 16 B taps per plane (this file)            105 us   l1tex LSU wavefronts 89 % of peak
//   2x2 pixel quad per thread sharing a 3x3 texel block (9 loads / 4 px)  141 us   lanes 2 px apart: every load
//                                                                                  touches twice the 128 B lines
//   4x1 pixel column per thread sharing a 5x2 block (10 loads / 4 px)     127 us   ~30 % of neighbours off the
//                                                                                  expected anchor, 80 regs
// The kernel is bound by L1 data-pipe wavefronts (a 32-lane x 16 B tap request spans 4-6 lines), not by HBM.
#include "common.cuh"
#include "sweep_l1.cuh"

namespace cnm {

// ---------------------------------------------------------------------------------------------
// camera pre-processing: one thread per (view, batch) pair, fp64
// ---------------------------------------------------------------------------------------------
__device__ bool invert_n(double* a, double* inv, int n) {
  // Gauss-Jordan with partial pivoting on an n x n row-major matrix (n <= 4)
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < n; ++j) inv[i * n + j] = (i == j) ? 1.0 : 0.0;
  for (int c = 0; c < n; ++c) {
    int piv = c;
    double best = fabs(a[c * n + c]);
    for (int r = c + 1; r < n; ++r)
      if (fabs(a[r * n + c]) > best) { best = fabs(a[r * n + c]); piv = r; }
    if (best == 0.0) return false;
    if (piv != c)
      for (int j = 0; j < n; ++j) {
        double t = a[c * n + j]; a[c * n + j] = a[piv * n + j]; a[piv * n + j] = t;
        t = inv[c * n + j]; inv[c * n + j] = inv[piv * n + j]; inv[piv * n + j] = t;
      }
    double d = 1.0 / a[c * n + c];
    for (int j = 0; j < n; ++j) { a[c * n + j] *= d; inv[c * n + j] *= d; }
    for (int r = 0; r < n; ++r) {
      if (r == c) continue;
      double f = a[r * n + c];
      if (f == 0.0) continue;
      for (int j = 0; j < n; ++j) { a[r * n + j] -= f * a[c * n + j]; inv[r * n + j] -= f * inv[c * n + j]; }
    }
  }
  return true;
}

__global__ void camera_setup_kernel(const float* __restrict__ ref_cam, const float* __restrict__ src_cam,
                                    float* __restrict__ hk, int B, int V) {
  pdl_trigger();
  pdl_wait();
  int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= B * V) return;
  int b = p % B;
  const float* rc = ref_cam + (size_t)b * 32;
  const float* sc = src_cam + (size_t)p * 32;
  double Er[16], Eri[16], Es[16], Kr[9], Kri[9], Ks[9];
  for (int i = 0; i < 16; ++i) { Er[i] = rc[i]; Es[i] = sc[i]; }
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) { Kr[i * 3 + j] = rc[16 + i * 4 + j]; Ks[i * 3 + j] = sc[16 + i * 4 + j]; }
  bool ok = invert_n(Er, Eri, 4);
  ok = invert_n(Kr, Kri, 3) && ok;
  double T[16];
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j) {
      double s = 0;
      for (int k = 0; k < 4; ++k) s += Es[i * 4 + k] * Eri[k * 4 + j];
      T[i * 4 + j] = s;
    }
  double RK[9];  // R * Kr^-1
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      double s = 0;
      for (int k = 0; k < 3; ++k) s += T[i * 4 + k] * Kri[k * 3 + j];
      RK[i * 3 + j] = s;
    }
  float* o = hk + (size_t)p * 12;
  const double nan = __longlong_as_double(0x7ff8000000000000LL);
  for (int i = 0; i < 3; ++i) {
    for (int j = 0; j < 3; ++j) {
      double s = 0;
      for (int k = 0; k < 3; ++k) s += Ks[i * 3 + k] * RK[k * 3 + j];
      o[i * 3 + j] = (float)(ok ? s : nan);
    }
    double s = 0;
    for (int k = 0; k < 3; ++k) s += Ks[i * 3 + k] * T[k * 4 + 3];
    o[9 + i] = (float)(ok ? s : nan);
  }
}

// ---------------------------------------------------------------------------------------------
// source repack: NCHW planar -> zero-bordered pixel-interleaved float4
// ---------------------------------------------------------------------------------------------
__global__ void pack_source_kernel(const float* __restrict__ src, float4* __restrict__ tex, int N, int H, int W) {
  pdl_trigger();
  pdl_wait();
  const int pitch = W + 4, rows = H + 3;
  size_t total = (size_t)N * rows * pitch;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    int xx = (int)(i % pitch);
    size_t t = i / pitch;
    int yy = (int)(t % rows);
    int n = (int)(t / rows);
    int x = xx - 1, y = yy - 1;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (x >= 0 && x < W && y >= 0 && y < H) {
      const float* s = src + (size_t)n * 3 * H * W + (size_t)y * W + x;
      v.x = __ldg(s);
      v.y = __ldg(s + (size_t)H * W);
      v.z = __ldg(s + 2 * (size_t)H * W);
    }
    tex[i] = v;
  }
}

// ---------------------------------------------------------------------------------------------
// fused plane sweep
// ---------------------------------------------------------------------------------------------
constexpr int CV_TX = 32, CV_TY = 8;

template <int DPB, int LAYOUT>  // LAYOUT: 0 = NCHW fp32, 1 = act bf16, 2 = act fp32 (tf32), 3 = act split bf16, 4 = act fp16 + fp8
__global__ void __launch_bounds__(CV_TX* CV_TY)
    cost_volume_kernel(const float* __restrict__ ref, const float4* __restrict__ tex, const float* __restrict__ hk,
                       void* __restrict__ out, int B, int H, int W, int D, double idmin, double idstep) {
  pdl_trigger();
  pdl_wait();
  __shared__ float zs[DPB];
  const int groups = D / DPB;
  const int p = blockIdx.z / groups;
  const int d0 = (blockIdx.z % groups) * DPB;
  const int tid = threadIdx.y * CV_TX + threadIdx.x;
  if (tid < DPB) zs[tid] = sweep_plane_z(idmin, idstep, d0 + tid);
  __syncthreads();

  const int x = blockIdx.x * CV_TX + threadIdx.x;
  const int y = blockIdx.y * CV_TY + threadIdx.y;
  if (x >= W || y >= H) return;
  const int b = p % B;
  const size_t HW = (size_t)H * W;

  SweepPixelL1 sp;
  sweep_l1_setup(sp, hk + (size_t)p * 12, x, y);
  const float* rp = ref + (size_t)b * 3 * HW + (size_t)y * W + x;
  sp.r0 = __ldg(rp); sp.r1 = __ldg(rp + HW); sp.r2 = __ldg(rp + 2 * HW);
  sp.pitch = W + 4;
  sp.tp = tex + (size_t)p * (H + 3) * sp.pitch + sp.pitch + 1;  // texel (0,0)
  sp.xmax = (float)W; sp.ymax = (float)H;

  float res[DPB];
#pragma unroll
  for (int i = 0; i < DPB; ++i) res[i] = sweep_l1_sample(sp, zs[i]);

  if (LAYOUT == 0) {
    float* o = reinterpret_cast<float*>(out) + ((size_t)p * D + d0) * HW + (size_t)y * W + x;
#pragma unroll
    for (int i = 0; i < DPB; ++i) __stcs(o + (size_t)i * HW, res[i]);
  } else if (LAYOUT == 1) {
    __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(out) + ((size_t)p * HW + (size_t)y * W + x) * D + d0;
#pragma unroll
    for (int i = 0; i < DPB; i += 8) Vec16<__nv_bfloat16>::store(o + i, res + i);
  } else if (LAYOUT == 2) {
    float* o = reinterpret_cast<float*>(out) + ((size_t)p * HW + (size_t)y * W + x) * D + d0;
#pragma unroll
    for (int i = 0; i < DPB; i += 4) Vec16<float>::store(o + i, res + i);
  } else if (LAYOUT == 3) {  // split bf16: planes [d0, d0+DPB) of the pixel's hi row, the lo row 64 physical channels further
    bf16s* o = reinterpret_cast<bf16s*>(out) + ((size_t)p * HW + (size_t)y * W + x) * ActTraits<bf16s>::pstride(D) +
               ActTraits<bf16s>::phys(d0);
#pragma unroll
    for (int i = 0; i < DPB; i += 8) Vec16<bf16s>::store(o + i, res + i);
  } else {  // fp16 + fp8: fp16 row, then the lo8 / hi8 bytes of the chunk's fp8 row
    h16f8* o = reinterpret_cast<h16f8*>(out) + ((size_t)p * HW + (size_t)y * W + x) * ActTraits<h16f8>::pstride(D) +
               ActTraits<h16f8>::phys(d0);
#pragma unroll
    for (int i = 0; i < DPB; i += 8) Vec16<h16f8>::store(o + i, res + i);
  }
}

}  // namespace cnm

using namespace cnm;

extern "C" int cnm_camera_setup(const float* ref_cam, const float* src_cam, float* hk, int B, int V, void* stream) {
  CNM_REQUIRE(ref_cam && src_cam && hk, CNM_ERR_INVALID_ARGUMENT, "cnm_camera_setup: null pointer");
  CNM_REQUIRE(B > 0 && V > 0, CNM_ERR_INVALID_ARGUMENT, "cnm_camera_setup: B=%d V=%d", B, V);
  int n = B * V;
  launch_pdl(camera_setup_kernel, dim3((n + 63) / 64), dim3(64), 0, (cudaStream_t)stream, ref_cam, src_cam, hk, B, V);
  return after_launch("cnm_camera_setup");
}

extern "C" int64_t cnm_packed_source_bytes(int N, int H, int W) {
  if (N <= 0 || H <= 0 || W <= 0) return 0;
  return (int64_t)N * (H + 3) * (W + 4) * 16;
}

extern "C" int cnm_pack_source(const float* src, void* packed, int N, int H, int W, void* stream) {
  CNM_REQUIRE(src && packed, CNM_ERR_INVALID_ARGUMENT, "cnm_pack_source: null pointer");
  CNM_REQUIRE(N > 0 && H > 0 && W > 0, CNM_ERR_INVALID_ARGUMENT, "cnm_pack_source: bad shape");
  CNM_REQUIRE(((uintptr_t)packed & 15) == 0, CNM_ERR_INVALID_ARGUMENT, "cnm_pack_source: packed must be 16B aligned");
  size_t total = (size_t)N * (H + 3) * (W + 4);
  int blocks = (int)((total + 255) / 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  launch_pdl(pack_source_kernel, dim3(blocks), dim3(256), 0, (cudaStream_t)stream, src, (float4*)packed, N, H, W);
  return after_launch("cnm_pack_source");
}

extern "C" int cnm_cost_volume(const float* ref, const void* packed_src, const float* hk, void* out, int B, int V,
                               int H, int W, int D, double idepth_min, double idepth_max, int layout, int dtype,
                               void* stream) {
  CNM_REQUIRE(ref && packed_src && hk && out, CNM_ERR_INVALID_ARGUMENT, "cnm_cost_volume: null pointer");
  CNM_REQUIRE(B > 0 && V > 0 && H > 0 && W > 0, CNM_ERR_INVALID_ARGUMENT, "cnm_cost_volume: bad shape");
  CNM_REQUIRE(D >= 16 && D % 16 == 0, CNM_ERR_UNSUPPORTED_SHAPE, "cnm_cost_volume: D=%d must be a multiple of 16", D);
  CNM_REQUIRE(layout == CNM_VOLUME_NCHW_F32 || layout == CNM_VOLUME_ACT, CNM_ERR_INVALID_ARGUMENT,
              "cnm_cost_volume: layout %d", layout);
  CNM_REQUIRE(((uintptr_t)packed_src & 15) == 0 && ((uintptr_t)out & 15) == 0, CNM_ERR_INVALID_ARGUMENT,
              "cnm_cost_volume: pointers must be 16B aligned");
  const int P = B * V;
  constexpr int DPB = 16;
  CNM_REQUIRE((long long)P * (D / DPB) <= 65535, CNM_ERR_UNSUPPORTED_SHAPE, "cnm_cost_volume: too many pairs x planes");
  const double step = (idepth_max - idepth_min) / (double)(D - 1);
  dim3 grid((W + CV_TX - 1) / CV_TX, (H + CV_TY - 1) / CV_TY, P * (D / DPB));
  dim3 block(CV_TX, CV_TY);
  cudaStream_t s = (cudaStream_t)stream;
  const float4* tex = (const float4*)packed_src;
  if (layout == CNM_VOLUME_NCHW_F32)
    launch_pdl(cost_volume_kernel<DPB, 0>, dim3(grid), dim3(block), 0, s, ref, tex, hk, out, B, H, W, D, idepth_min, step);
  else if (dtype == CNM_BF16)
    launch_pdl(cost_volume_kernel<DPB, 1>, dim3(grid), dim3(block), 0, s, ref, tex, hk, out, B, H, W, D, idepth_min, step);
  else if (dtype == CNM_TF32)
    launch_pdl(cost_volume_kernel<DPB, 2>, dim3(grid), dim3(block), 0, s, ref, tex, hk, out, B, H, W, D, idepth_min, step);
  else if (dtype == CNM_BF16X3)
    launch_pdl(cost_volume_kernel<DPB, 3>, dim3(grid), dim3(block), 0, s, ref, tex, hk, out, B, H, W, D, idepth_min, step);
  else if (dtype == CNM_F16F8)
    launch_pdl(cost_volume_kernel<DPB, 4>, dim3(grid), dim3(block), 0, s, ref, tex, hk, out, B, H, W, D, idepth_min, step);
  else
    return fail(CNM_ERR_INVALID_ARGUMENT, "cnm_cost_volume: dtype %d", dtype);
  return after_launch("cnm_cost_volume");
}
