// Depth-evaluation metrics as one fused masked reduction per image.
//
// Reference semantics: eval.py:1009-1047 (cal_metrics) on top of utils/metric.py:149-362 -- pred clamped to
// [min_depth, max_depth], valid = min < gt < max and finite, then l1 / abs_relative / rmse / scale_invariant /
// sq_relative / rmse_log / ratio_threshold(1.25^k) on the valid pixels.  The reference makes nine numpy passes over the
// masked copies; here one pass accumulates the ten sums they share (fp64, fixed reduction order: deterministic) and a
// finishing step turns them into the nine numbers.
#include "common.cuh"

namespace cnm {

constexpr int MET_THREADS = 256, MET_SUMS = 10, MET_OUT = 10;

__global__ void __launch_bounds__(MET_THREADS)
    depth_metrics_partial_kernel(const float* __restrict__ pred, const float* __restrict__ gt, double* __restrict__ part,
                                 size_t HW, float dmin, float dmax) {
  __shared__ double red[MET_THREADS / 32][MET_SUMS];
  const int b = blockIdx.y, NB = gridDim.x;
  const float* p = pred + (size_t)b * HW;
  const float* g = gt + (size_t)b * HW;
  const double l1 = log(1.25), l2 = log(1.25 * 1.25), l3 = log(1.25 * 1.25 * 1.25);
  double s[MET_SUMS];
#pragma unroll
  for (int i = 0; i < MET_SUMS; ++i) s[i] = 0.0;
  for (size_t j = (size_t)blockIdx.x * MET_THREADS + threadIdx.x; j < HW; j += (size_t)NB * MET_THREADS) {
    const float gf = g[j];
    if (!(gf < dmax && gf > dmin) || isinf(gf)) continue;  // NaN fails the comparisons
    const double gd = (double)gf, pd = (double)fminf(fmaxf(p[j], dmin), dmax);
    const double d = gd - pd, ld = log(gd) - log(pd);
    s[0] += 1.0;
    s[1] += fabs(d);
    s[2] += fabs(d) / gd;
    s[3] += d * d;
    s[4] += d * d / gd;
    s[5] += ld;
    s[6] += ld * ld;
    s[7] += fabs(ld) < l1 ? 1.0 : 0.0;
    s[8] += fabs(ld) < l2 ? 1.0 : 0.0;
    s[9] += fabs(ld) < l3 ? 1.0 : 0.0;
  }
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
#pragma unroll
  for (int i = 0; i < MET_SUMS; ++i) {
    double v = s[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (l == 0) red[w][i] = v;
  }
  __syncthreads();
  if (threadIdx.x < MET_SUMS) {
    double v = 0.0;
    for (int k = 0; k < MET_THREADS / 32; ++k) v += red[k][threadIdx.x];
    part[((size_t)b * NB + blockIdx.x) * MET_SUMS + threadIdx.x] = v;
  }
}

// out[b] = {num_valid, l1, abs_relative, rmse, scale_invariant, sq_relative, rmse_log, a1, a2, a3}
__global__ void depth_metrics_finish_kernel(const double* __restrict__ part, double* __restrict__ out, int NB) {
  const int b = blockIdx.x;
  __shared__ double s[MET_SUMS];
  if (threadIdx.x < MET_SUMS) {
    double v = 0.0;
    for (int k = 0; k < NB; ++k) v += part[((size_t)b * NB + k) * MET_SUMS + threadIdx.x];
    s[threadIdx.x] = v;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double* o = out + (size_t)b * MET_OUT;
    const double n = s[0];
    o[0] = n;
    if (n == 0.0) {
      for (int i = 1; i < MET_OUT; ++i) o[i] = nan("");
      return;
    }
    o[1] = s[1] / n;
    o[2] = s[2] / n;
    o[3] = sqrt(s[3] / n);
    o[4] = sqrt(s[6] / n - (s[5] * s[5]) / (n * n));
    o[5] = s[4] / n;
    o[6] = sqrt(s[6] / n);
    o[7] = s[7] / n; o[8] = s[8] / n; o[9] = s[9] / n;
  }
}

static int met_blocks(size_t HW) {
  const size_t nb = (HW + MET_THREADS * 8 - 1) / (MET_THREADS * 8);
  return (int)(nb < 1 ? 1 : (nb > 64 ? 64 : nb));
}

}  // namespace cnm

using namespace cnm;

extern "C" int64_t cnm_depth_metrics_workspace_bytes(int B, int H, int W) {
  if (B <= 0 || H <= 0 || W <= 0) return 0;
  return (int64_t)B * met_blocks((size_t)H * W) * MET_SUMS * (int64_t)sizeof(double);
}

extern "C" int cnm_depth_metrics(const float* pred, const float* gt, double* out, void* workspace, int B, int H, int W,
                                 float min_depth, float max_depth, void* stream) {
  CNM_REQUIRE(pred && gt && out && workspace, CNM_ERR_INVALID_ARGUMENT, "cnm_depth_metrics: null pointer");
  CNM_REQUIRE(B > 0 && H > 0 && W > 0 && min_depth > 0.0f && max_depth > min_depth, CNM_ERR_INVALID_ARGUMENT,
              "cnm_depth_metrics: bad shape or depth range");
  CNM_REQUIRE(((uintptr_t)workspace & 7) == 0 && ((uintptr_t)out & 7) == 0, CNM_ERR_INVALID_ARGUMENT,
              "cnm_depth_metrics: out / workspace must be 8B aligned");
  const size_t HW = (size_t)H * W;
  const int NB = met_blocks(HW);
  cudaStream_t s = (cudaStream_t)stream;
  depth_metrics_partial_kernel<<<dim3(NB, B), MET_THREADS, 0, s>>>(pred, gt, (double*)workspace, HW, min_depth, max_depth);
  int rc = after_launch("cnm_depth_metrics(partial)");
  if (rc) return rc;
  depth_metrics_finish_kernel<<<B, 32, 0, s>>>((const double*)workspace, out, NB);
  return after_launch("cnm_depth_metrics(finish)");
}
