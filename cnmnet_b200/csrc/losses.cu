// Forward values of the reference's training losses as fused masked reductions (the first step
// toward training -- no gradients here).
//
// Reference semantics: depthnet/losses.py
//   IdepthLoss            :31-48    L1 (or L1 of log10) over pixels with gt > 0, pred > 0, both finite
//   IdepthwithProbLoss    :51-74    the same mask, mean of prob * |diff|  (log mode: 10 * |log10 diff|)
//   IdepthLoss_234        :7-28     0.1 * mean_k mean|disp_k - nearest-downsampled gt|, k = 2, 3, 4
//   surface_normal_loss   :77-122   mean (or prob-weighted mean) of 1 - cosine similarity over valid pixels + mean angle
// One pass per loss: per-block partial sums in fp64 with a fixed reduction order, a one-block finish (deterministic).
#include "common.cuh"

namespace cnm {

constexpr int LOSS_THREADS = 256, LOSS_MAX_BLOCKS = 148 * 4;

template <int K>
__device__ __forceinline__ void block_partials(double (&s)[K], double* __restrict__ part) {
  __shared__ double red[LOSS_THREADS / 32][K];
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
#pragma unroll
  for (int i = 0; i < K; ++i) {
    double v = s[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (l == 0) red[w][i] = v;
  }
  __syncthreads();
  if (threadIdx.x < K) {
    double v = 0.0;
    for (int k = 0; k < LOSS_THREADS / 32; ++k) v += red[k][threadIdx.x];
    part[(size_t)blockIdx.x * K + threadIdx.x] = v;
  }
}

template <int K>
__device__ __forceinline__ void sum_partials(const double* __restrict__ part, int nblocks, double* s) {
  // called by thread 0 of a single block
  for (int i = 0; i < K; ++i) s[i] = 0.0;
  for (int b = 0; b < nblocks; ++b)
    for (int i = 0; i < K; ++i) s[i] += part[(size_t)b * K + i];
}

// ---- IdepthLoss / IdepthwithProbLoss ----
__global__ void __launch_bounds__(LOSS_THREADS)
    idepth_loss_partial_kernel(const float* __restrict__ pred, const float* __restrict__ gt, const float* __restrict__ prob,
                               double* __restrict__ part, size_t n, int log_mode) {
  double s[2] = {0.0, 0.0};  // sum, count
  for (size_t i = (size_t)blockIdx.x * LOSS_THREADS + threadIdx.x; i < n; i += (size_t)gridDim.x * LOSS_THREADS) {
    const float p = pred[i], g = gt[i];
    if (!(g > 0.0f) || !(p > 0.0f) || isinf(g) || isinf(p)) continue;  // NaN fails the comparisons
    // the reference evaluates in fp32 (torch.log10 / abs on float tensors); the sums are kept in fp64 here
    float d = log_mode ? fabsf(log10f(p) - log10f(g)) : fabsf(p - g);
    if (prob) d = prob[i] * (log_mode ? 10.0f * d : d);
    s[0] += (double)d;
    s[1] += 1.0;
  }
  block_partials<2>(s, part);
}
__global__ void idepth_loss_finish_kernel(const double* __restrict__ part, int nblocks, double* __restrict__ out) {
  if (threadIdx.x) return;
  double s[2];
  sum_partials<2>(part, nblocks, s);
  out[0] = s[0] / s[1];  // mean over the masked pixels (NaN when the mask is empty, like torch)
  out[1] = s[1];
}

// ---- IdepthLoss_234 ----
__global__ void __launch_bounds__(LOSS_THREADS)
    idepth234_partial_kernel(const float* __restrict__ d2, const float* __restrict__ d3, const float* __restrict__ d4,
                             const float* __restrict__ gt, double* __restrict__ part, int B, int H, int W) {
  double s[3] = {0.0, 0.0, 0.0};
  const float* disp[3] = {d2, d3, d4};
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const int f = 2 << k, h = H / f, w = W / f;  // F.interpolate(mode='nearest'): source index = dst * f
    const size_t n = (size_t)B * h * w;
    for (size_t i = (size_t)blockIdx.x * LOSS_THREADS + threadIdx.x; i < n; i += (size_t)gridDim.x * LOSS_THREADS) {
      const int x = (int)(i % w), y = (int)((i / w) % h), b = (int)(i / ((size_t)w * h));
      s[k] += (double)fabsf(disp[k][i] - gt[((size_t)b * H + (size_t)y * f) * W + (size_t)x * f]);
    }
  }
  block_partials<3>(s, part);
}
__global__ void idepth234_finish_kernel(const double* __restrict__ part, int nblocks, double* __restrict__ out, double n2,
                                        double n3, double n4) {
  if (threadIdx.x) return;
  double s[3];
  sum_partials<3>(part, nblocks, s);
  out[0] = 0.1 * (s[0] / n2 + s[1] / n3 + s[2] / n4) / 3.0;
}

// ---- surface_normal_loss ----
__global__ void __launch_bounds__(LOSS_THREADS)
    normal_loss_partial_kernel(const float* __restrict__ pred, const float* __restrict__ gt, const uint8_t* __restrict__ valid,
                               const float* __restrict__ prob, double* __restrict__ part, int B, size_t HW) {
  double s[4] = {0.0, 0.0, 0.0, 0.0};  // sum (1-sim)[*prob], sum prob (or count), sum angle, count
  const size_t n = (size_t)B * HW;
  for (size_t i = (size_t)blockIdx.x * LOSS_THREADS + threadIdx.x; i < n; i += (size_t)gridDim.x * LOSS_THREADS) {
    const size_t b = i / HW, j = i % HW;
    if (valid && !valid[i]) continue;
    const float* p = pred + b * 3 * HW + j;
    const float* g = gt + b * 3 * HW + j;
    const float p0 = p[0], p1 = p[HW], p2 = p[2 * HW], g0 = g[0], g1 = g[HW], g2 = g[2 * HW];
    if (!isfinite(p0 + p1 + p2) || !isfinite(g0 + g1 + g2)) continue;
    // F.cosine_similarity(dim=1, eps=1e-8): x.y / (max(|x|, eps) * max(|y|, eps)), fp32 like the reference
    const float np_ = fmaxf(sqrtf(p0 * p0 + p1 * p1 + p2 * p2), 1e-8f), ng = fmaxf(sqrtf(g0 * g0 + g1 * g1 + g2 * g2), 1e-8f);
    const float sim = (p0 * g0 + p1 * g1 + p2 * g2) / (np_ * ng);
    const float wgt = prob ? prob[i] : 1.0f;
    s[0] += (double)((1.0f - sim) * wgt);
    s[1] += (double)wgt;
    s[2] += (double)acosf(fminf(fmaxf(sim, -1.0f), 1.0f));
    s[3] += 1.0;
  }
  block_partials<4>(s, part);
}
__global__ void normal_loss_finish_kernel(const double* __restrict__ part, int nblocks, double* __restrict__ out) {
  if (threadIdx.x) return;
  double s[4];
  sum_partials<4>(part, nblocks, s);
  out[0] = s[0] / s[1];                                  // mean, or sum((1-sim)*prob) / sum(prob)
  out[1] = s[2] / s[3] / 3.14159265358979323846 * 180.0;  // mean angle in degrees
  out[2] = s[3];
}

static int loss_blocks(size_t n) {
  const size_t nb = (n + (size_t)LOSS_THREADS * 8 - 1) / ((size_t)LOSS_THREADS * 8);
  return (int)(nb < 1 ? 1 : (nb > (size_t)LOSS_MAX_BLOCKS ? (size_t)LOSS_MAX_BLOCKS : nb));
}

}  // namespace cnm

using namespace cnm;

extern "C" int64_t cnm_loss_workspace_bytes(void) { return (int64_t)LOSS_MAX_BLOCKS * 4 * (int64_t)sizeof(double); }

extern "C" int cnm_idepth_loss(const float* pred, const float* gt, const float* prob, double* out, void* workspace,
                               int64_t n, int log_mode, void* stream) {
  CNM_REQUIRE(pred && gt && out && workspace && n > 0, CNM_ERR_INVALID_ARGUMENT, "cnm_idepth_loss: bad argument");
  const int nb = loss_blocks((size_t)n);
  cudaStream_t s = (cudaStream_t)stream;
  idepth_loss_partial_kernel<<<nb, LOSS_THREADS, 0, s>>>(pred, gt, prob, (double*)workspace, (size_t)n, log_mode);
  int rc = after_launch("cnm_idepth_loss(partial)");
  if (rc) return rc;
  idepth_loss_finish_kernel<<<1, 32, 0, s>>>((const double*)workspace, nb, out);
  return after_launch("cnm_idepth_loss(finish)");
}

extern "C" int cnm_idepth_loss_234(const float* disp2, const float* disp3, const float* disp4, const float* gt, double* out,
                                   void* workspace, int B, int H, int W, void* stream) {
  CNM_REQUIRE(disp2 && disp3 && disp4 && gt && out && workspace, CNM_ERR_INVALID_ARGUMENT, "cnm_idepth_loss_234: null pointer");
  CNM_REQUIRE(B > 0 && H > 0 && W > 0 && H % 16 == 0 && W % 16 == 0, CNM_ERR_UNSUPPORTED_SHAPE,
              "cnm_idepth_loss_234: H and W must be multiples of 16 (disp4 is 1/8 resolution), got %dx%d", H, W);
  const int nb = loss_blocks((size_t)B * H * W / 4);
  cudaStream_t s = (cudaStream_t)stream;
  idepth234_partial_kernel<<<nb, LOSS_THREADS, 0, s>>>(disp2, disp3, disp4, gt, (double*)workspace, B, H, W);
  int rc = after_launch("cnm_idepth_loss_234(partial)");
  if (rc) return rc;
  const double n = (double)B * H * W;
  idepth234_finish_kernel<<<1, 32, 0, s>>>((const double*)workspace, nb, out, n / 4, n / 16, n / 64);
  return after_launch("cnm_idepth_loss_234(finish)");
}

extern "C" int cnm_surface_normal_loss(const float* pred, const float* gt_normal, const uint8_t* valid, const float* prob,
                                       double* out, void* workspace, int B, int H, int W, void* stream) {
  CNM_REQUIRE(pred && gt_normal && out && workspace, CNM_ERR_INVALID_ARGUMENT, "cnm_surface_normal_loss: null pointer");
  CNM_REQUIRE(B > 0 && H > 0 && W > 0, CNM_ERR_INVALID_ARGUMENT, "cnm_surface_normal_loss: bad shape");
  const size_t HW = (size_t)H * W;
  const int nb = loss_blocks((size_t)B * HW);
  cudaStream_t s = (cudaStream_t)stream;
  normal_loss_partial_kernel<<<nb, LOSS_THREADS, 0, s>>>(pred, gt_normal, valid, prob, (double*)workspace, B, HW);
  int rc = after_launch("cnm_surface_normal_loss(partial)");
  if (rc) return rc;
  normal_loss_finish_kernel<<<1, 32, 0, s>>>((const double*)workspace, nb, out);
  return after_launch("cnm_surface_normal_loss(finish)");
}
