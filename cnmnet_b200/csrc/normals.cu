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
  pdl_trigger();
  pdl_wait();
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
// CNM (K3b): replace-by-plane-mean with the reference's SEQUENTIAL semantics (depth_util.py:205-238: plane i averages the
// normals as left by planes 0..i-1), computed in parallel over pixels.
//
// For a pixel, let its planes be i1 < i2 < ... (its bits in the mask stack).  Plane i sees that pixel's ORIGINAL normal
// if i is its first plane, otherwise the (float-rounded) mean of its previous plane.  Hence
//     mean_i = ( S_i + sum_{k<i} C(i,k) * float(mean_k) ) / count_i
// with S_i = sum of original normals over pixels whose first plane is i, C(i,k) = number of pixels of plane i whose
// previous plane is k.  Pass A accumulates S (fp64, fixed reduction order), count and C (integers) over all pixels,
// pass B solves the P <= 32 means in order, pass C walks every pixel's plane list once more: loss terms
// 1 - cos(mean_i, value seen by plane i) and the final value (mean of its last plane).  Disjoint masks (what the
// reference's loader produces, scannet/dataloader_batch.py:180-194) are the special case C = 0.
// ---------------------------------------------------------------------------------------------
constexpr int CNM_THREADS = 256, CNM_MAX_P = 32;

struct CnmWorkspace {
  double* partS;   // [B][NB][P][3]
  double* means;   // [B][P][3]
  double* lossp;   // [B][NB]
  unsigned* cnt;   // [B][P]
  unsigned* cm;    // [B][P][P]
};

__host__ __device__ inline int cnm_blocks_per_image(int H, int W) {
  const int nb = (H * W + CNM_THREADS * 4 - 1) / (CNM_THREADS * 4);
  return nb < 1 ? 1 : (nb > 64 ? 64 : nb);
}
static size_t cnm_align(size_t v) { return (v + 255) & ~(size_t)255; }
static size_t cnm_carve(void* base, int B, int P, int NB, CnmWorkspace* ws) {
  size_t off = 0;
  uint8_t* p = static_cast<uint8_t*>(base);
  auto take = [&](size_t bytes) { uint8_t* r = p ? p + off : nullptr; off += cnm_align(bytes); return r; };
  double* partS = reinterpret_cast<double*>(take((size_t)B * NB * P * 3 * sizeof(double)));
  double* means = reinterpret_cast<double*>(take((size_t)B * P * 3 * sizeof(double)));
  double* lossp = reinterpret_cast<double*>(take((size_t)B * NB * sizeof(double)));
  unsigned* cnt = reinterpret_cast<unsigned*>(take(((size_t)B * P + (size_t)B * P * P) * sizeof(unsigned)));
  if (ws) { ws->partS = partS; ws->means = means; ws->lossp = lossp; ws->cnt = cnt; ws->cm = cnt + (size_t)B * P; }
  return off;
}

__device__ __forceinline__ unsigned cnm_plane_bits(const uint8_t* __restrict__ segs, size_t HW, size_t j, int np) {
  unsigned bits = 0;
  for (int i = 0; i < np; ++i) bits |= (segs[(size_t)i * HW + j] != 0 ? 1u : 0u) << i;
  return bits;
}
__device__ __forceinline__ double warp_sum_ordered(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__global__ void __launch_bounds__(CNM_THREADS)
    cnm_accumulate_kernel(const float* __restrict__ normals, const uint8_t* __restrict__ segs,
                          const int32_t* __restrict__ planes_num, CnmWorkspace ws, int P, int H, int W) {
  __shared__ double s_S[CNM_THREADS / 32][CNM_MAX_P][3];
  __shared__ unsigned s_cnt[CNM_MAX_P], s_cm[CNM_MAX_P * CNM_MAX_P];
  const int b = blockIdx.y, NB = gridDim.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const size_t HW = (size_t)H * W;
  const int np = min(planes_num[b], P);
  for (int i = threadIdx.x; i < (CNM_THREADS / 32) * CNM_MAX_P * 3; i += CNM_THREADS) (&s_S[0][0][0])[i] = 0.0;
  for (int i = threadIdx.x; i < CNM_MAX_P; i += CNM_THREADS) s_cnt[i] = 0;
  for (int i = threadIdx.x; i < CNM_MAX_P * CNM_MAX_P; i += CNM_THREADS) s_cm[i] = 0;
  __syncthreads();
  const float* n = normals + (size_t)b * 3 * HW;
  const uint8_t* m = segs + (size_t)b * P * HW;
  const size_t stride = (size_t)NB * CNM_THREADS;
  const size_t iters = (HW + stride - 1) / stride;
  for (size_t it = 0; it < iters; ++it) {  // uniform trip count: every lane takes part in the ballots
    const size_t j = it * stride + (size_t)blockIdx.x * CNM_THREADS + threadIdx.x;
    const unsigned bits = j < HW ? cnm_plane_bits(m, HW, j, np) : 0u;
    unsigned any = __ballot_sync(0xffffffffu, bits != 0);
    if (!any) continue;
    float n0 = 0.f, n1 = 0.f, n2 = 0.f;
    if (bits) { n0 = n[j]; n1 = n[HW + j]; n2 = n[2 * HW + j]; }
    unsigned warp_bits = bits;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) warp_bits |= __shfl_xor_sync(0xffffffffu, warp_bits, o);
    while (warp_bits) {  // planes present in this warp, ascending
      const int i = __ffs(warp_bits) - 1;
      warp_bits &= warp_bits - 1;
      const bool in = (bits >> i) & 1u;
      const unsigned earlier = bits & ((1u << i) - 1u);
      if (in) {
        atomicAdd(&s_cnt[i], 1u);
        if (earlier) atomicAdd(&s_cm[i * CNM_MAX_P + (31 - __clz(earlier))], 1u);
      }
      const bool first = in && !earlier;
      if (__ballot_sync(0xffffffffu, first)) {
        const double a0 = warp_sum_ordered(first ? (double)n0 : 0.0), a1 = warp_sum_ordered(first ? (double)n1 : 0.0),
                     a2 = warp_sum_ordered(first ? (double)n2 : 0.0);
        if (lane == 0) { s_S[warp][i][0] += a0; s_S[warp][i][1] += a1; s_S[warp][i][2] += a2; }
      }
    }
  }
  __syncthreads();
  for (int t = threadIdx.x; t < P * 3; t += CNM_THREADS) {
    double acc = 0.0;
    for (int w = 0; w < CNM_THREADS / 32; ++w) acc += s_S[w][t / 3][t % 3];
    ws.partS[(((size_t)b * NB + blockIdx.x) * P) * 3 + t] = acc;
  }
  for (int t = threadIdx.x; t < P; t += CNM_THREADS)
    if (s_cnt[t]) atomicAdd(&ws.cnt[(size_t)b * P + t], s_cnt[t]);
  for (int t = threadIdx.x; t < P * P; t += CNM_THREADS) {
    const unsigned v = s_cm[(t / P) * CNM_MAX_P + t % P];
    if (v) atomicAdd(&ws.cm[(size_t)b * P * P + t], v);
  }
}

__global__ void __launch_bounds__(128)
    cnm_solve_kernel(const int32_t* __restrict__ planes_num, CnmWorkspace ws, int P, int NB) {
  __shared__ double S[CNM_MAX_P * 3];
  const int b = blockIdx.x;
  const int np = min(planes_num[b], P);
  for (int t = threadIdx.x; t < P * 3; t += blockDim.x) {
    double acc = 0.0;
    for (int k = 0; k < NB; ++k) acc += ws.partS[(((size_t)b * NB + k) * P) * 3 + t];  // fixed order
    S[t] = acc;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double* mean = ws.means + (size_t)b * P * 3;
    for (int i = 0; i < np; ++i) {
      double a0 = S[3 * i], a1 = S[3 * i + 1], a2 = S[3 * i + 2];
      for (int k = 0; k < i; ++k) {
        const unsigned c = ws.cm[((size_t)b * P + i) * P + k];
        if (c) {  // pixels of plane i that plane k rewrote: they hold float(mean_k)
          a0 += (double)c * (double)(float)mean[3 * k];
          a1 += (double)c * (double)(float)mean[3 * k + 1];
          a2 += (double)c * (double)(float)mean[3 * k + 2];
        }
      }
      const double cnt = (double)ws.cnt[(size_t)b * P + i];
      mean[3 * i] = a0 / cnt;  // 0/0 -> NaN like the reference; NOT renormalised
      mean[3 * i + 1] = a1 / cnt;
      mean[3 * i + 2] = a2 / cnt;
    }
  }
}

__global__ void __launch_bounds__(CNM_THREADS)
    cnm_apply_kernel(float* __restrict__ normals, const uint8_t* __restrict__ segs, const int32_t* __restrict__ planes_num,
                     CnmWorkspace ws, int want_loss, int P, int H, int W) {
  __shared__ double s_mean[CNM_MAX_P * 3], s_nm[CNM_MAX_P], scratch[33];
  const int b = blockIdx.y, NB = gridDim.x;
  const size_t HW = (size_t)H * W;
  const int np = min(planes_num[b], P);
  for (int t = threadIdx.x; t < np * 3; t += CNM_THREADS) s_mean[t] = ws.means[(size_t)b * P * 3 + t];
  __syncthreads();
  for (int t = threadIdx.x; t < np; t += CNM_THREADS)
    s_nm[t] = fmax(sqrt(s_mean[3 * t] * s_mean[3 * t] + s_mean[3 * t + 1] * s_mean[3 * t + 1] +
                        s_mean[3 * t + 2] * s_mean[3 * t + 2]), 1e-8);
  __syncthreads();
  float* n = normals + (size_t)b * 3 * HW;
  const uint8_t* m = segs + (size_t)b * P * HW;
  double lacc = 0.0;
  for (size_t j = (size_t)blockIdx.x * CNM_THREADS + threadIdx.x; j < HW; j += (size_t)NB * CNM_THREADS) {
    unsigned bits = cnm_plane_bits(m, HW, j, np);
    if (!bits) continue;
    float c0 = n[j], c1 = n[HW + j], c2 = n[2 * HW + j];
    while (bits) {
      const int i = __ffs(bits) - 1;
      bits &= bits - 1;
      const double m0 = s_mean[3 * i], m1 = s_mean[3 * i + 1], m2 = s_mean[3 * i + 2];
      if (want_loss) {  // 1 - cos(mean_i, the value plane i saw at this pixel)
        const double a0 = c0, a1 = c1, a2 = c2;
        lacc += 1.0 - (a0 * m0 + a1 * m1 + a2 * m2) / (s_nm[i] * fmax(sqrt(a0 * a0 + a1 * a1 + a2 * a2), 1e-8));
      }
      c0 = (float)m0; c1 = (float)m1; c2 = (float)m2;
    }
    n[j] = c0; n[HW + j] = c1; n[2 * HW + j] = c2;
  }
  if (want_loss) {
    // block sum in a fixed order
    double v = warp_sum_ordered(lacc);
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    if (l == 0) scratch[w] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
      double t = 0.0;
      for (int k = 0; k < CNM_THREADS / 32; ++k) t += scratch[k];
      ws.lossp[(size_t)b * NB + blockIdx.x] = t;
    }
  }
}

// loss = sum_b sum_i mean over ALL H*W pixels of (1 - cos): off-mask pixels contribute exactly 1 each
__global__ void cnm_loss_kernel(const int32_t* __restrict__ planes_num, CnmWorkspace ws, double* __restrict__ loss, int B,
                                int P, int NB, double HW) {
  if (threadIdx.x || blockIdx.x) return;
  double total = 0.0;
  for (int b = 0; b < B; ++b) {
    double acc = 0.0;
    for (int k = 0; k < NB; ++k) acc += ws.lossp[(size_t)b * NB + k];
    const int np = min(planes_num[b], P);
    for (int i = 0; i < np; ++i) {
      const double cnt = (double)ws.cnt[(size_t)b * P + i];
      // an empty mask below planes_num: the reference's mean is 0/0 = NaN, cosine_similarity is NaN at every pixel
      // and the whole loss becomes NaN (depth_util.py:222-231)
      acc += cnt > 0.0 ? HW - cnt : __longlong_as_double(0x7ff8000000000000LL);
    }
    total += acc / HW;
  }
  *loss = total;
}

__global__ void intrinsics_inverse_kernel(const float* __restrict__ cam, float* __restrict__ Kinv, int B) {
  pdl_trigger();
  pdl_wait();
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
  pdl_trigger();
  pdl_wait();
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
  static DynSmemLimit limit;
  cudaError_t e = limit.ensure(depth2normal_kernel, smem);
  if (e != cudaSuccess) return fail(CNM_ERR_CUDA, "cnm_depth2normal: smem attr: %s", cudaGetErrorString(e));
  dim3 grid((W + NT_X - 1) / NT_X, (H + NT_Y - 1) / NT_Y, B), block(NT_X, NT_Y);
  launch_pdl(depth2normal_kernel, dim3(grid), dim3(block), smem, (cudaStream_t)stream, depth, Kinv, normals, points, H, W, k_size);
  return after_launch("cnm_depth2normal");
}

extern "C" int64_t cnm_plane_regularize_workspace_bytes(int B, int P, int H, int W) {
  if (B <= 0 || P <= 0 || P > CNM_MAX_P || H <= 0 || W <= 0) return 0;
  return (int64_t)cnm_carve(nullptr, B, P, cnm_blocks_per_image(H, W), nullptr);
}

extern "C" int cnm_plane_regularize(float* normals, const uint8_t* segs, const int32_t* planes_num, double* loss,
                                    void* workspace, int B, int P, int H, int W, void* stream) {
  CNM_REQUIRE(normals && segs && planes_num && workspace, CNM_ERR_INVALID_ARGUMENT, "cnm_plane_regularize: null pointer");
  CNM_REQUIRE(B > 0 && P > 0 && H > 0 && W > 0, CNM_ERR_INVALID_ARGUMENT, "cnm_plane_regularize: bad shape");
  CNM_REQUIRE(P <= CNM_MAX_P, CNM_ERR_UNSUPPORTED_SHAPE, "cnm_plane_regularize: at most %d mask slots (got %d)", CNM_MAX_P, P);
  CNM_REQUIRE(((uintptr_t)workspace & 255) == 0, CNM_ERR_INVALID_ARGUMENT, "cnm_plane_regularize: workspace must be 256B aligned");
  cudaStream_t s = (cudaStream_t)stream;
  const int NB = cnm_blocks_per_image(H, W);
  CnmWorkspace ws;
  cnm_carve(workspace, B, P, NB, &ws);
  cudaError_t e = cudaMemsetAsync(ws.cnt, 0, ((size_t)B * P + (size_t)B * P * P) * sizeof(unsigned), s);
  if (e != cudaSuccess) return fail(CNM_ERR_CUDA, "cnm_plane_regularize: memset: %s", cudaGetErrorString(e));
  cnm_accumulate_kernel<<<dim3(NB, B), CNM_THREADS, 0, s>>>(normals, segs, planes_num, ws, P, H, W);
  int rc = after_launch("cnm_plane_regularize(accumulate)");
  if (rc) return rc;
  cnm_solve_kernel<<<B, 128, 0, s>>>(planes_num, ws, P, NB);
  rc = after_launch("cnm_plane_regularize(solve)");
  if (rc) return rc;
  cnm_apply_kernel<<<dim3(NB, B), CNM_THREADS, 0, s>>>(normals, segs, planes_num, ws, loss ? 1 : 0, P, H, W);
  rc = after_launch("cnm_plane_regularize(apply)");
  if (rc || !loss) return rc;
  cnm_loss_kernel<<<1, 32, 0, s>>>(planes_num, ws, loss, B, P, NB, (double)H * W);
  return after_launch("cnm_plane_regularize(loss)");
}

extern "C" int cnm_intrinsics_inverse(const float* cam, float* Kinv, int B, void* stream) {
  CNM_REQUIRE(cam && Kinv && B > 0, CNM_ERR_INVALID_ARGUMENT, "cnm_intrinsics_inverse: bad argument");
  launch_pdl(intrinsics_inverse_kernel, dim3((B + 63) / 64), dim3(64), 0, (cudaStream_t)stream, cam, Kinv, B);
  return after_launch("cnm_intrinsics_inverse");
}

extern "C" int cnm_reciprocal(const float* idepth, float* depth, int64_t n, void* stream) {
  CNM_REQUIRE(idepth && depth && n > 0, CNM_ERR_INVALID_ARGUMENT, "cnm_reciprocal: bad argument");
  int blocks = (int)((n + 255) / 256);
  if (blocks > 148 * 8) blocks = 148 * 8;
  launch_pdl(reciprocal_kernel, dim3(blocks), dim3(256), 0, (cudaStream_t)stream, idepth, depth, n);
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
