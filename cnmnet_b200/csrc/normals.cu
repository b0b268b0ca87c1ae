// K3 / K3b: depth -> point cloud -> k x k plane-fit normals, and the Combined Normal Map.
//
// Reference semantics: depthnet/inverse_warp.py:27-43 (pixel2cam), depthnet/depth_util.py:149-203
// (Unfold 9x9 -> bmm -> det -> inverse -> matmul -> norm), :205-238 and :243-278 (CNM).
// Here: per CTA a 32x16 pixel tile + halo of masked points in shared memory, the 9 second/first
// moment sums (xx,xy,xz,yy,yz,zz,x,y,z) as separable box sums, closed-form symmetric 3x3 solve.
// All of it in fp64: cond(AtA) ~ 1e4 makes fp32 lose 3 digits.
#include "common.cuh"

namespace cnm {

constexpr int NT_X = 32, NT_Y = 16;

__global__ void __launch_bounds__(NT_X* NT_Y)
    depth2normal_kernel(const float* __restrict__ depth, const float* __restrict__ Kinv, float* __restrict__ normals,
                        float* __restrict__ points, int H, int W, int k) {
  extern __shared__ double sm[];
  const int r = k / 2;
  const int HX = NT_X + 2 * r, HY = NT_Y + 2 * r;
  double* q = sm;                 // [HY][HX][3] masked points
  double* hs = sm + HY * HX * 3;  // [HY][NT_X][9] horizontal sums
  const int b = blockIdx.z;
  const int x0 = blockIdx.x * NT_X, y0 = blockIdx.y * NT_Y;
  const int tid = threadIdx.y * NT_X + threadIdx.x;
  const int nthr = NT_X * NT_Y;
  const float* Kb = Kinv + (size_t)b * 9;
  double Ki[9];
#pragma unroll
  for (int i = 0; i < 9; ++i) Ki[i] = (double)__ldg(Kb + i);
  const float* dp = depth + (size_t)b * H * W;

  // 1. masked points of the halo tile (zero outside the image = Unfold padding, depth_util.py:165)
  for (int i = tid; i < HY * HX; i += nthr) {
    const int hy = i / HX, hx = i % HX;
    const int y = y0 + hy - r, x = x0 + hx - r;
    double px = 0.0, py = 0.0, pz = 0.0;
    if (x >= 0 && x < W && y >= 0 && y < H) {
      const float d = __ldg(dp + (size_t)y * W + x);
      if (d > 0.0f && d < 10.0f) {  // NaN fails both (depth_util.py:162)
        const double dd = (double)d, fx = (double)x, fy = (double)y;
        px = (Ki[0] * fx + Ki[1] * fy + Ki[2]) * dd;
        py = (Ki[3] * fx + Ki[4] * fy + Ki[5]) * dd;
        pz = (Ki[6] * fx + Ki[7] * fy + Ki[8]) * dd;
      }
    }
    q[i * 3 + 0] = px; q[i * 3 + 1] = py; q[i * 3 + 2] = pz;
  }
  __syncthreads();

  // 2. horizontal k-tap sums of the 9 moments for every halo row and tile column
  for (int i = tid; i < HY * NT_X; i += nthr) {
    const int hy = i / NT_X, cx = i % NT_X;
    double s[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    const double* qq = q + (hy * HX + cx) * 3;
    for (int t = 0; t < k; ++t) {
      const double X = qq[t * 3], Y = qq[t * 3 + 1], Z = qq[t * 3 + 2];
      s[0] += X * X; s[1] += X * Y; s[2] += X * Z; s[3] += Y * Y; s[4] += Y * Z; s[5] += Z * Z;
      s[6] += X; s[7] += Y; s[8] += Z;
    }
    // moment-major so the vertical pass reads are conflict-free across a warp
#pragma unroll
    for (int j = 0; j < 9; ++j) hs[(j * HY + hy) * NT_X + cx] = s[j];
  }
  __syncthreads();

  const int x = x0 + threadIdx.x, y = y0 + threadIdx.y;
  if (x >= W || y >= H) return;
  // 3. vertical k-tap sums
  double m[9];
#pragma unroll
  for (int j = 0; j < 9; ++j) {
    double a = 0.0;
    const double* col = hs + (j * HY + threadIdx.y) * NT_X + threadIdx.x;
    for (int t = 0; t < k; ++t) a += col[t * NT_X];
    m[j] = a;
  }
  // 4. symmetric 3x3 solve  M n = s  with the reference's det guard (depth_util.py:185-200)
  const double xx = m[0], xy = m[1], xz = m[2], yy = m[3], yz = m[4], zz = m[5];
  const double c00 = yy * zz - yz * yz, c01 = xz * yz - xy * zz, c02 = xy * yz - xz * yy;
  const double det = xx * c00 + xy * c01 + xz * c02;
  double n0, n1, n2;
  if (!(det >= 1e-5)) {  // isnan(det) | det < 1e-5  ->  M := I
    n0 = m[6]; n1 = m[7]; n2 = m[8];
  } else {
    const double c11 = xx * zz - xz * xz, c12 = xz * xy - xx * yz, c22 = xx * yy - xy * xy;
    const double id = 1.0 / det;
    n0 = (c00 * m[6] + c01 * m[7] + c02 * m[8]) * id;
    n1 = (c01 * m[6] + c11 * m[7] + c12 * m[8]) * id;
    n2 = (c02 * m[6] + c12 * m[7] + c22 * m[8]) * id;
  }
  const double inv = 1.0 / (sqrt(n0 * n0 + n1 * n1 + n2 * n2) + 1e-5);
  const size_t HW = (size_t)H * W;
  float* np = normals + (size_t)b * 3 * HW + (size_t)y * W + x;
  np[0] = (float)(n0 * inv); np[HW] = (float)(n1 * inv); np[2 * HW] = (float)(n2 * inv);
  if (points) {  // unmasked P = Kinv.[x,y,1] * depth  (inverse_warp.py:40-43)
    const double dd = (double)__ldg(dp + (size_t)y * W + x), fx = (double)x, fy = (double)y;
    float* pp = points + (size_t)b * 3 * HW + (size_t)y * W + x;
    pp[0] = (float)(Ki[0] * fx + Ki[1] * fy + Ki[2]) * (float)dd;
    pp[HW] = (float)(Ki[3] * fx + Ki[4] * fy + Ki[5]) * (float)dd;
    pp[2 * HW] = (float)(Ki[6] * fx + Ki[7] * fy + Ki[8]) * (float)dd;
  }
}

// ---------------------------------------------------------------------------------------------
// CNM: sequential replace-by-plane-mean; one 1024-thread CTA per batch item
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ double block_sum(double v, double* scratch) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  __syncthreads();
  if (l == 0) scratch[w] = v;
  __syncthreads();
  double t = (threadIdx.x < (blockDim.x >> 5)) ? scratch[threadIdx.x] : 0.0;
  if (w == 0) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    if (l == 0) scratch[32] = t;
  }
  __syncthreads();
  return scratch[32];
}

__global__ void __launch_bounds__(1024)
    plane_regularize_kernel(float* __restrict__ normals, const uint8_t* __restrict__ segs,
                            const int32_t* __restrict__ planes_num, double* __restrict__ loss, int P, int H, int W) {
  __shared__ double scratch[33];
  const int b = blockIdx.x;
  const size_t HW = (size_t)H * W;
  float* n = normals + (size_t)b * 3 * HW;
  int np = planes_num[b];
  if (np > P) np = P;
  double loss_b = 0.0;
  for (int i = 0; i < np; ++i) {
    const uint8_t* m = segs + ((size_t)b * P + i) * HW;
    double s0 = 0, s1 = 0, s2 = 0, c = 0;
    for (size_t j = threadIdx.x; j < HW; j += blockDim.x)
      if (m[j]) { s0 += n[j]; s1 += n[HW + j]; s2 += n[2 * HW + j]; c += 1.0; }
    s0 = block_sum(s0, scratch); s1 = block_sum(s1, scratch); s2 = block_sum(s2, scratch);
    c = block_sum(c, scratch);
    const double m0 = s0 / c, m1 = s1 / c, m2 = s2 / c;  // 0/0 -> NaN like the reference; NOT renormalised
    if (loss) {
      // mean over ALL H*W pixels of 1 - cos(mean, n*mask); off-mask pixels contribute exactly 1
      const double nm = fmax(sqrt(m0 * m0 + m1 * m1 + m2 * m2), 1e-8);
      double acc = 0.0;
      for (size_t j = threadIdx.x; j < HW; j += blockDim.x) {
        double cs = 0.0;
        if (m[j]) {
          const double a0 = n[j], a1 = n[HW + j], a2 = n[2 * HW + j];
          cs = (a0 * m0 + a1 * m1 + a2 * m2) / (nm * fmax(sqrt(a0 * a0 + a1 * a1 + a2 * a2), 1e-8));
        }
        acc += 1.0 - cs;
      }
      loss_b += block_sum(acc, scratch) / (double)HW;
    }
    __syncthreads();
    for (size_t j = threadIdx.x; j < HW; j += blockDim.x)
      if (m[j]) { n[j] = (float)m0; n[HW + j] = (float)m1; n[2 * HW + j] = (float)m2; }
    __syncthreads();
  }
  if (loss && threadIdx.x == 0) atomicAdd(loss, loss_b);
}

__global__ void zero_double_kernel(double* p) { *p = 0.0; }

__global__ void intrinsics_inverse_kernel(const float* __restrict__ cam, float* __restrict__ Kinv, int B) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const float* K = cam + (size_t)b * 32 + 16;  // cam[b,1,:3,:3]
  const double a = K[0], bb = K[1], c = K[2], d = K[4], e = K[5], f = K[6], g = K[8], h = K[9], i = K[10];
  const double A = e * i - f * h, Bc = -(d * i - f * g), Cc = d * h - e * g;
  const double inv = 1.0 / (a * A + bb * Bc + c * Cc);
  float* o = Kinv + (size_t)b * 9;
  o[0] = (float)(A * inv); o[1] = (float)(-(bb * i - c * h) * inv); o[2] = (float)((bb * f - c * e) * inv);
  o[3] = (float)(Bc * inv); o[4] = (float)((a * i - c * g) * inv); o[5] = (float)(-(a * f - c * d) * inv);
  o[6] = (float)(Cc * inv); o[7] = (float)(-(a * h - bb * g) * inv); o[8] = (float)((a * e - bb * d) * inv);
}

__global__ void reciprocal_kernel(const float* __restrict__ in, float* __restrict__ out, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    out[i] = 1.0f / in[i];
}

// ---------------------------------------------------------------------------------------------
// depth-driven inverse warp (inverse_warp.py:46-118)
// ---------------------------------------------------------------------------------------------
__global__ void inverse_warp_kernel(const float* __restrict__ feat, const float* __restrict__ depth,
                                    const float* __restrict__ pose, const float* __restrict__ K,
                                    const float* __restrict__ Kinv, float* __restrict__ out, int C, int H, int W,
                                    int border) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y, b = blockIdx.z;
  if (x >= W || y >= H) return;
  const float* Kb = K + b * 9;
  const float* Ki = Kinv + b * 9;
  const float* Pm = pose + b * 12;
  const float d = depth[((size_t)b * H + y) * W + x];
  const float fx = (float)x, fy = (float)y;
  // cam = Kinv.[x,y,1] * depth ; p = (K.pose)[:, :3].cam + (K.pose)[:, 3]
  float cam[3], proj[12];
  for (int i = 0; i < 3; ++i) cam[i] = (Ki[i * 3] * fx + Ki[i * 3 + 1] * fy + Ki[i * 3 + 2]) * d;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 4; ++j)
      proj[i * 4 + j] = Kb[i * 3] * Pm[j] + Kb[i * 3 + 1] * Pm[4 + j] + Kb[i * 3 + 2] * Pm[8 + j];
  float p[3];
  for (int i = 0; i < 3; ++i) p[i] = proj[i * 4] * cam[0] + proj[i * 4 + 1] * cam[1] + proj[i * 4 + 2] * cam[2] + proj[i * 4 + 3];
  const float Z = fmaxf(p[2], 1e-3f);
  float xn = 2.0f * (p[0] / Z) / (float)(W - 1) - 1.0f;
  float yn = 2.0f * (p[1] / Z) / (float)(H - 1) - 1.0f;
  if (!border) {
    if (xn > 1.0f || xn < -1.0f) xn = 2.0f;
    if (yn > 1.0f || yn < -1.0f) yn = 2.0f;
  }
  // grid_sample, align_corners=False (today's default; the reference relies on it, inverse_warp.py:116)
  float ix = ((xn + 1.0f) * (float)W - 1.0f) * 0.5f, iy = ((yn + 1.0f) * (float)H - 1.0f) * 0.5f;
  if (border) {
    ix = fminf(fmaxf(ix, 0.0f), (float)(W - 1));
    iy = fminf(fmaxf(iy, 0.0f), (float)(H - 1));
  }
  const float x0f = floorf(ix), y0f = floorf(iy);
  const float wx1 = ix - x0f, wy1 = iy - y0f, wx0 = 1.0f - wx1, wy0 = 1.0f - wy1;
  const int xi = (int)x0f, yi = (int)y0f;
  const bool vx0 = xi >= 0 && xi < W, vx1 = xi + 1 >= 0 && xi + 1 < W;
  const bool vy0 = yi >= 0 && yi < H, vy1 = yi + 1 >= 0 && yi + 1 < H;
  const size_t HW = (size_t)H * W;
  for (int c = 0; c < C; ++c) {
    const float* f = feat + ((size_t)b * C + c) * HW;
    float v = 0.0f;
    if (vy0 && vx0) v += f[(size_t)yi * W + xi] * wx0 * wy0;
    if (vy0 && vx1) v += f[(size_t)yi * W + xi + 1] * wx1 * wy0;
    if (vy1 && vx0) v += f[(size_t)(yi + 1) * W + xi] * wx0 * wy1;
    if (vy1 && vx1) v += f[(size_t)(yi + 1) * W + xi + 1] * wx1 * wy1;
    out[((size_t)b * C + c) * HW + (size_t)y * W + x] = v;
  }
}

}  // namespace cnm

using namespace cnm;

extern "C" int cnm_depth2normal(const float* depth, const float* Kinv, float* normals, float* points, int B, int H,
                                int W, int k_size, void* stream) {
  CNM_REQUIRE(depth && Kinv && normals, CNM_ERR_INVALID_ARGUMENT, "cnm_depth2normal: null pointer");
  CNM_REQUIRE(B > 0 && H > 0 && W > 0, CNM_ERR_INVALID_ARGUMENT, "cnm_depth2normal: bad shape");
  CNM_REQUIRE(k_size >= 1 && k_size <= 15 && (k_size & 1), CNM_ERR_UNSUPPORTED_SHAPE,
              "cnm_depth2normal: k_size=%d must be odd and <= 15", k_size);
  CNM_REQUIRE(B <= 65535, CNM_ERR_UNSUPPORTED_SHAPE, "cnm_depth2normal: B too large");
  const int r = k_size / 2;
  const size_t smem = ((size_t)(NT_Y + 2 * r) * (NT_X + 2 * r) * 3 + (size_t)(NT_Y + 2 * r) * NT_X * 9) * sizeof(double);
  static size_t configured = 0;
  if (smem > configured) {
    cudaError_t e = cudaFuncSetAttribute(depth2normal_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return fail(CNM_ERR_CUDA, "cnm_depth2normal: smem attr: %s", cudaGetErrorString(e));
    configured = smem;
  }
  dim3 grid((W + NT_X - 1) / NT_X, (H + NT_Y - 1) / NT_Y, B), block(NT_X, NT_Y);
  depth2normal_kernel<<<grid, block, smem, (cudaStream_t)stream>>>(depth, Kinv, normals, points, H, W, k_size);
  return after_launch("cnm_depth2normal");
}

extern "C" int cnm_plane_regularize(float* normals, const uint8_t* segs, const int32_t* planes_num, double* loss,
                                    int B, int P, int H, int W, void* stream) {
  CNM_REQUIRE(normals && segs && planes_num, CNM_ERR_INVALID_ARGUMENT, "cnm_plane_regularize: null pointer");
  CNM_REQUIRE(B > 0 && P > 0 && H > 0 && W > 0, CNM_ERR_INVALID_ARGUMENT, "cnm_plane_regularize: bad shape");
  cudaStream_t s = (cudaStream_t)stream;
  if (loss) {
    zero_double_kernel<<<1, 1, 0, s>>>(loss);
    int rc = after_launch("cnm_plane_regularize(zero)");
    if (rc) return rc;
  }
  plane_regularize_kernel<<<B, 1024, 0, s>>>(normals, segs, planes_num, loss, P, H, W);
  return after_launch("cnm_plane_regularize");
}

extern "C" int cnm_intrinsics_inverse(const float* cam, float* Kinv, int B, void* stream) {
  CNM_REQUIRE(cam && Kinv && B > 0, CNM_ERR_INVALID_ARGUMENT, "cnm_intrinsics_inverse: bad argument");
  intrinsics_inverse_kernel<<<(B + 63) / 64, 64, 0, (cudaStream_t)stream>>>(cam, Kinv, B);
  return after_launch("cnm_intrinsics_inverse");
}

extern "C" int cnm_reciprocal(const float* idepth, float* depth, int64_t n, void* stream) {
  CNM_REQUIRE(idepth && depth && n > 0, CNM_ERR_INVALID_ARGUMENT, "cnm_reciprocal: bad argument");
  int blocks = (int)((n + 255) / 256);
  if (blocks > 148 * 8) blocks = 148 * 8;
  reciprocal_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(idepth, depth, n);
  return after_launch("cnm_reciprocal");
}

extern "C" int cnm_inverse_warp(const float* feat, const float* depth, const float* pose, const float* K,
                                const float* Kinv, float* out, int B, int C, int H, int W, int padding_mode,
                                void* stream) {
  CNM_REQUIRE(feat && depth && pose && K && Kinv && out, CNM_ERR_INVALID_ARGUMENT, "cnm_inverse_warp: null pointer");
  CNM_REQUIRE(B > 0 && C > 0 && H > 1 && W > 1, CNM_ERR_INVALID_ARGUMENT, "cnm_inverse_warp: bad shape");
  CNM_REQUIRE(padding_mode == 0 || padding_mode == 1, CNM_ERR_INVALID_ARGUMENT, "cnm_inverse_warp: padding_mode");
  dim3 block(32, 8), grid((W + 31) / 32, (H + 7) / 8, B);
  inverse_warp_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(feat, depth, pose, K, Kinv, out, C, H, W, padding_mode);
  return after_launch("cnm_inverse_warp");
}
