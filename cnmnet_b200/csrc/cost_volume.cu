// K1: camera pre-processing, source repack and the fused plane-sweep cost volume.
//
// Reference semantics: depthnet/depth_util.py:13-56 (process_camera_parameters),
// depthnet/depthNet_model.py:185-224 (getVolume: 64 x {mul, add, div, grid_sample, abs, sum}).
// Here: one launch for all (view, batch) pairs and all planes; the warped volume never exists.
//
// Data layout in HBM
//   ref      [B,3,H,W] fp32 (public NCHW), read once per (pair, plane group)
//   texels   [P][H+6][W+4] float4 = (r,g,b,0): source views repacked pixel-interleaved with a
//            zero border at row/col -1 and five/two trailing rows/cols, so that after clamping the
//            sample position to [-1, W] x [-1, H] every bilinear tap is an unpredicated 16 B
//            load and grid_sample's zeros padding falls out of the border texels.
//   out      NCHW fp32 [P,D,H,W] (parity layout) or act [P,H,W,D] bf16/fp32 (conv1.0 operand)
#include "common.cuh"

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
  const int pitch = W + 4, rows = H + 6;
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
// One thread owns a column of CV_Q = 4 vertically adjacent pixels (lanes = consecutive x, so every load is
// unit-stride across the warp).  For a plane the four sample positions are ~1 texel apart vertically, so their 16
// bilinear taps normally fall inside the 5 x 2 texel block anchored at the first pixel's (x0, y0): 10 loads instead
// of 16 (the kernel is bound by L1 data-pipe wavefronts, profiles/r01_cost_volume_ncu.md).  A pixel whose own anchor
// is not at the expected (x0, y0+q) re-loads its 4 taps (happens once per ~1/|scale-1| pixels).
constexpr int CV_TX = 32, CV_TY = 4, CV_Q = 4;  // a CTA covers 32 x 16 pixels

struct Tap4 {
  float4 t00, t10, t01, t11;
};

__device__ __forceinline__ float cv_cost(const Tap4& t, float wx1, float wy1, float r0, float r1, float r2) {
  const float wx0 = 1.0f - wx1, wy0 = 1.0f - wy1;
  const float w00 = wx0 * wy0, w10 = wx1 * wy0, w01 = wx0 * wy1, w11 = wx1 * wy1;
  const float c0 = fmaf(t.t11.x, w11, fmaf(t.t01.x, w01, fmaf(t.t10.x, w10, fmaf(t.t00.x, w00, -r0))));
  const float c1 = fmaf(t.t11.y, w11, fmaf(t.t01.y, w01, fmaf(t.t10.y, w10, fmaf(t.t00.y, w00, -r1))));
  const float c2 = fmaf(t.t11.z, w11, fmaf(t.t01.z, w01, fmaf(t.t10.z, w10, fmaf(t.t00.z, w00, -r2))));
  return fabsf(c0) + fabsf(c1) + fabsf(c2);
}

// sample position of a pixel with homography row products (a0,a1,a2) at plane depth z: clamped (ix, iy)
__device__ __forceinline__ void cv_project(float a0, float a1, float a2, float z, float k0, float k1, float k2e,
                                           float xmax, float ymax, float& ix, float& iy) {
  const float px = fmaf(a0, z, k0), py = fmaf(a1, z, k1);
  const float den = fmaf(a2, z, k2e);  // p_z + 1e-6
  float rcp;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rcp) : "f"(den));
  rcp = fmaf(fmaf(-den, rcp, 1.0f), rcp, rcp);  // one Newton step: ~0.5 ulp
  // grid_sample un-normalisation with align_corners=False: ix = u - 0.5; clamp into the zero border (NaN -> -1)
  ix = fminf(fmaxf(fmaf(px, rcp, -0.5f), -1.0f), xmax);
  iy = fminf(fmaxf(fmaf(py, rcp, -0.5f), -1.0f), ymax);
}

template <int DPB, int LAYOUT>  // LAYOUT: 0 = NCHW fp32, 1 = act bf16, 2 = act fp32
__global__ void __launch_bounds__(CV_TX* CV_TY)
    cost_volume_kernel(const float* __restrict__ ref, const float4* __restrict__ tex, const float* __restrict__ hk,
                       void* __restrict__ out, int B, int H, int W, int D, double idmin, double idstep) {
  __shared__ float zs[DPB];
  const int groups = D / DPB;
  const int p = blockIdx.z / groups;
  const int d0 = (blockIdx.z % groups) * DPB;
  const int tid = threadIdx.y * CV_TX + threadIdx.x;
  if (tid < DPB) zs[tid] = (float)(1.0 / (idmin + (double)(d0 + tid) * idstep));  // fp64 like the Python double
  __syncthreads();

  const int x = blockIdx.x * CV_TX + threadIdx.x;
  const int y = CV_Q * (blockIdx.y * CV_TY + threadIdx.y);
  if (x >= W || y >= H) return;
  const int b = p % B;
  const size_t HW = (size_t)H * W;

  const float* h = hk + (size_t)p * 12;
  const float h1 = __ldg(h + 1), h4 = __ldg(h + 4), h7 = __ldg(h + 7);
  const float fx = (float)x;
  const float b0 = fmaf(__ldg(h + 0), fx, __ldg(h + 2)), b1 = fmaf(__ldg(h + 3), fx, __ldg(h + 5)),
              b2 = fmaf(__ldg(h + 6), fx, __ldg(h + 8));
  // a = Hm . [x, y+q, 1]^T (KRKiUV) and the reference pixels of the column (rows past H are computed on a clamped
  // address and never stored)
  float a0[CV_Q], a1[CV_Q], a2[CV_Q], r[CV_Q][3];
  const float* rp = ref + (size_t)b * 3 * HW + x;
#pragma unroll
  for (int q = 0; q < CV_Q; ++q) {
    const float fy = (float)(y + q);
    a0[q] = fmaf(h1, fy, b0); a1[q] = fmaf(h4, fy, b1); a2[q] = fmaf(h7, fy, b2);
    const int yr = min(y + q, H - 1);
#pragma unroll
    for (int c = 0; c < 3; ++c) r[q][c] = __ldg(rp + c * HW + (size_t)yr * W);
  }
  const float k0 = __ldg(h + 9), k1 = __ldg(h + 10), k2 = __ldg(h + 11) + 1e-6f;

  const int pitch = W + 4;
  const float4* tp = tex + (size_t)p * (H + 6) * pitch + pitch + 1;  // texel (0,0)
  const float xmax = (float)W, ymax = (float)H;
  constexpr int NBUF = (LAYOUT == 0) ? 1 : DPB;
  float res[CV_Q][NBUF];

#pragma unroll
  for (int i = 0; i < DPB; ++i) {
    const float z = zs[i];
    int xi[CV_Q], yi[CV_Q];
    float wx[CV_Q], wy[CV_Q];
#pragma unroll
    for (int q = 0; q < CV_Q; ++q) {
      float ix, iy;
      cv_project(a0[q], a1[q], a2[q], z, k0, k1, k2, xmax, ymax, ix, iy);
      const float xf = floorf(ix), yf = floorf(iy);
      xi[q] = (int)xf; yi[q] = (int)yf;
      wx[q] = ix - xf; wy[q] = iy - yf;
    }
    // (CV_Q+1) x 2 texel block anchored at the first pixel
    const float4* t0 = tp + yi[0] * pitch + xi[0];
    float4 T[CV_Q + 1][2];
#pragma unroll
    for (int j = 0; j <= CV_Q; ++j) { T[j][0] = __ldg(t0 + j * pitch); T[j][1] = __ldg(t0 + j * pitch + 1); }
    float c[CV_Q];
#pragma unroll
    for (int q = 0; q < CV_Q; ++q) {
      Tap4 t;
      t.t00 = T[q][0]; t.t10 = T[q][1]; t.t01 = T[q + 1][0]; t.t11 = T[q + 1][1];
      if (q > 0 && (xi[q] != xi[0] || yi[q] != yi[0] + q)) {  // irregular neighbour: its own four taps
        const float4* tt = tp + yi[q] * pitch + xi[q];
        t.t00 = __ldg(tt); t.t10 = __ldg(tt + 1); t.t01 = __ldg(tt + pitch); t.t11 = __ldg(tt + pitch + 1);
      }
      c[q] = cv_cost(t, wx[q], wy[q], r[q][0], r[q][1], r[q][2]);
    }
    if (LAYOUT == 0) {
      float* o = reinterpret_cast<float*>(out) + ((size_t)p * D + d0 + i) * HW + (size_t)y * W + x;
#pragma unroll
      for (int q = 0; q < CV_Q; ++q)
        if (y + q < H) __stcs(o + (size_t)q * W, c[q]);
    } else {
#pragma unroll
      for (int q = 0; q < CV_Q; ++q) res[q][i % NBUF] = c[q];
    }
  }

  if (LAYOUT != 0) {
#pragma unroll
    for (int q = 0; q < CV_Q; ++q) {
      if (y + q >= H) continue;
      const size_t pix = (size_t)p * HW + (size_t)(y + q) * W + x;
      if (LAYOUT == 1) {
        __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(out) + pix * D + d0;
#pragma unroll
        for (int i = 0; i < DPB; i += 8) Vec16<__nv_bfloat16>::store(o + i, res[q] + i);
      } else {
        float* o = reinterpret_cast<float*>(out) + pix * D + d0;
#pragma unroll
        for (int i = 0; i < DPB; i += 4) Vec16<float>::store(o + i, res[q] + i);
      }
    }
  }
}

}  // namespace cnm

using namespace cnm;

extern "C" int cnm_camera_setup(const float* ref_cam, const float* src_cam, float* hk, int B, int V, void* stream) {
  CNM_REQUIRE(ref_cam && src_cam && hk, CNM_ERR_INVALID_ARGUMENT, "cnm_camera_setup: null pointer");
  CNM_REQUIRE(B > 0 && V > 0, CNM_ERR_INVALID_ARGUMENT, "cnm_camera_setup: B=%d V=%d", B, V);
  int n = B * V;
  camera_setup_kernel<<<(n + 63) / 64, 64, 0, (cudaStream_t)stream>>>(ref_cam, src_cam, hk, B, V);
  return after_launch("cnm_camera_setup");
}

extern "C" int64_t cnm_packed_source_bytes(int N, int H, int W) {
  if (N <= 0 || H <= 0 || W <= 0) return 0;
  return (int64_t)N * (H + 6) * (W + 4) * 16;
}

extern "C" int cnm_pack_source(const float* src, void* packed, int N, int H, int W, void* stream) {
  CNM_REQUIRE(src && packed, CNM_ERR_INVALID_ARGUMENT, "cnm_pack_source: null pointer");
  CNM_REQUIRE(N > 0 && H > 0 && W > 0, CNM_ERR_INVALID_ARGUMENT, "cnm_pack_source: bad shape");
  CNM_REQUIRE(((uintptr_t)packed & 15) == 0, CNM_ERR_INVALID_ARGUMENT, "cnm_pack_source: packed must be 16B aligned");
  size_t total = (size_t)N * (H + 6) * (W + 4);
  int blocks = (int)((total + 255) / 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  pack_source_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(src, (float4*)packed, N, H, W);
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
  CNM_REQUIRE((long long)P * (D / 8) <= 65535, CNM_ERR_UNSUPPORTED_SHAPE, "cnm_cost_volume: too many pairs x planes");
  const double step = (idepth_max - idepth_min) / (double)(D - 1);
  dim3 block(CV_TX, CV_TY);
  const unsigned gx = (unsigned)((W + CV_TX - 1) / CV_TX), gy = (unsigned)((H + CV_Q * CV_TY - 1) / (CV_Q * CV_TY));
  cudaStream_t s = (cudaStream_t)stream;
  const float4* tex = (const float4*)packed_src;
  if (layout == CNM_VOLUME_NCHW_F32)
    cost_volume_kernel<16, 0><<<dim3(gx, gy, P * (D / 16)), block, 0, s>>>(ref, tex, hk, out, B, H, W, D, idepth_min, step);
  else if (dtype == CNM_BF16)
    cost_volume_kernel<8, 1><<<dim3(gx, gy, P * (D / 8)), block, 0, s>>>(ref, tex, hk, out, B, H, W, D, idepth_min, step);
  else if (dtype == CNM_TF32)
    cost_volume_kernel<8, 2><<<dim3(gx, gy, P * (D / 8)), block, 0, s>>>(ref, tex, hk, out, B, H, W, D, idepth_min, step);
  else
    return fail(CNM_ERR_INVALID_ARGUMENT, "cnm_cost_volume: dtype %d", dtype);
  return after_launch("cnm_cost_volume");
}
