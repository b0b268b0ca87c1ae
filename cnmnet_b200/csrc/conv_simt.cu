// Validation twin of the tcgen05 implicit-GEMM convolution: same descriptor, same step list, same
// packed weights and act tensors, but plain fp32 FMAs on CUDA cores.  Used by the GPU tests to
// separate "the layer program / weight packing is wrong" from "the tensor-core kernel is wrong".
// Not used by the product pipeline.
#include "common.cuh"

namespace cnm {

struct SimtSrc {
  const void* ptr;
  int H, W, C;
};
struct SimtParams {
  SimtSrc srcs[4];
  int N, Ho, Wo, Cout, c_out_total, c_out_offset, stride, relu, num_steps, split;
  float acc_scale;
};

constexpr int SM_TM = 64, SM_TN = 64;

template <typename T>
__global__ void __launch_bounds__(256)
    conv_simt_kernel(SimtParams p, const cnm_conv_step* __restrict__ steps, const T* __restrict__ wp,
                     const float* __restrict__ bias, T* __restrict__ out) {
  constexpr int CH = 128 / sizeof(T);
  __shared__ float As[SM_TM][CH + 1];
  __shared__ float Bs[SM_TN][CH + 1];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const size_t M = (size_t)p.N * p.Ho * p.Wo;
  const size_t m0 = (size_t)blockIdx.x * SM_TM;
  const int n0 = blockIdx.y * SM_TN;
  const size_t Kt = (size_t)p.num_steps * CH;
  float acc[4][4] = {};
  for (int st = 0; st < p.num_steps; ++st) {
    const cnm_conv_step s = steps[st];
    const SimtSrc src = p.srcs[s.src];
    const int py = (s.sub >> 1) & 1, px = s.sub & 1;
    for (int i = threadIdx.x; i < SM_TM * CH; i += 256) {
      const int r = i / CH, c = i % CH;
      const size_t m = m0 + r;
      float v = 0.0f;
      if (m < M) {
        const int ox = (int)(m % p.Wo), oy = (int)((m / p.Wo) % p.Ho), n = (int)(m / ((size_t)p.Wo * p.Ho));
        int iy = oy + s.dy, ix = ox + s.dx;
        if (p.stride == 2) { iy = 2 * iy + py; ix = 2 * ix + px; }
        if (iy >= 0 && iy < src.H && ix >= 0 && ix < src.W)
          v = to_f<T>(reinterpret_cast<const T*>(src.ptr)[(((size_t)n * src.H + iy) * src.W + ix) * src.C + s.c0 + c]);
      }
      As[r][c] = v;
    }
    for (int i = threadIdx.x; i < SM_TN * CH; i += 256) {
      const int r = i / CH, c = i % CH;
      Bs[r][c] = (n0 + r < p.Cout) ? to_f<T>(wp[(size_t)(n0 + r) * Kt + (size_t)st * CH + c]) : 0.0f;
    }
    __syncthreads();
#pragma unroll 8
    for (int k = 0; k < CH; ++k) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) { a[i] = As[ty * 4 + i][k]; b[i] = Bs[tx * 4 + i][k]; }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const size_t m = m0 + ty * 4 + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int co = n0 + tx * 4 + j;
      if (co >= p.Cout) continue;
      float v = acc[i][j] + (bias ? __ldg(bias + co) : 0.0f);
      if (p.relu) v = fmaxf(v, 0.0f);
      if (p.split) {  // CNM_BF16X3 output: hi row, lo row 64 physical channels further
        __nv_bfloat16 hi, lo;
        split_bf16(v, hi, lo);
        T* o = out + m * p.c_out_total + p.c_out_offset + ActTraits<bf16s>::phys(co);
        o[0] = from_f<T>(__bfloat162float(hi));
        o[64] = from_f<T>(__bfloat162float(lo));
      } else {
        out[m * p.c_out_total + p.c_out_offset + co] = from_f<T>(v);
      }
    }
  }
}

// CNM_F16F8 twin: the first half of the K-steps are fp8 rows, the second half fp16 rows (64 channels) against fp16 weight columns; fp8 rows
// (128 e4m3 bytes: lo8 x 64 | hi8 x 64) against 128 e4m3 weight bytes; everything accumulates in fp32 and is scaled by
// 2^-acc_exp, exactly the arithmetic the two tensor-core passes perform.
__global__ void __launch_bounds__(256)
    conv_simt_f16f8_kernel(SimtParams p, const cnm_conv_step* __restrict__ steps, const uint8_t* __restrict__ wp,
                           const float* __restrict__ bias, h16f8* __restrict__ out) {
  constexpr int CH = 64;
  __shared__ float As[SM_TM][CH + 1];
  __shared__ float Bs[SM_TN][CH + 1];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const size_t M = (size_t)p.N * p.Ho * p.Wo;
  const size_t m0 = (size_t)blockIdx.x * SM_TM;
  const int n0 = blockIdx.y * SM_TN;
  const size_t Kbytes = (size_t)p.num_steps * 128;
  float acc[4][4] = {};
  for (int st = 0; st < p.num_steps; ++st) {
    const cnm_conv_step s = steps[st];
    const SimtSrc src = p.srcs[s.src];
    const int py = (s.sub >> 1) & 1, px = s.sub & 1;
    const bool f8 = st < (p.num_steps >> 1);
    for (int sub = 0; sub < (f8 ? 2 : 1); ++sub) {
      for (int i = threadIdx.x; i < SM_TM * CH; i += 256) {
        const int r = i / CH, c = i % CH;
        const size_t m = m0 + r;
        float v = 0.0f;
        if (m < M) {
          const int ox = (int)(m % p.Wo), oy = (int)((m / p.Wo) % p.Ho), n = (int)(m / ((size_t)p.Wo * p.Ho));
          int iy = oy + s.dy, ix = ox + s.dx;
          if (p.stride == 2) { iy = 2 * iy + py; ix = 2 * ix + px; }
          if (iy >= 0 && iy < src.H && ix >= 0 && ix < src.W) {
            const uint8_t* row = reinterpret_cast<const uint8_t*>(src.ptr) +
                                 ((((size_t)n * src.H + iy) * src.W + ix) * src.C + s.c0) * 2;
            v = f8 ? e4m3x2_to_float2(row[sub * 64 + c]).x : __half2float(reinterpret_cast<const __half*>(row)[c]);
          }
        }
        As[r][c] = v;
      }
      for (int i = threadIdx.x; i < SM_TN * CH; i += 256) {
        const int r = i / CH, c = i % CH;
        float v = 0.0f;
        if (n0 + r < p.Cout) {
          const uint8_t* row = wp + (size_t)(n0 + r) * Kbytes + (size_t)st * 128;
          v = f8 ? e4m3x2_to_float2(row[sub * 64 + c]).x : __half2float(reinterpret_cast<const __half*>(row)[c]);
        }
        Bs[r][c] = v;
      }
      __syncthreads();
#pragma unroll 8
      for (int k = 0; k < CH; ++k) {
        float a[4], b[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) { a[i] = As[ty * 4 + i][k]; b[i] = Bs[tx * 4 + i][k]; }
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
      }
      __syncthreads();
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const size_t m = m0 + ty * 4 + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int co = n0 + tx * 4 + j;
      if (co >= p.Cout) continue;
      float v = fmaf(acc[i][j], p.acc_scale, bias ? __ldg(bias + co) : 0.0f);
      if (p.relu) v = fmaxf(v, 0.0f);
      store_scalar<h16f8>(out + m * p.c_out_total + p.c_out_offset, co, v);
    }
  }
}

int check_conv_args(const char* fn, const cnm_conv_desc* d, const cnm_conv_src* srcs, const void* steps, const void* wp,
                    const void* out) {
  CNM_REQUIRE(d && srcs && steps && wp && out, CNM_ERR_INVALID_ARGUMENT, "%s: null pointer", fn);
  CNM_REQUIRE(d->dtype == CNM_BF16 || d->dtype == CNM_TF32 || d->dtype == CNM_BF16X3 || d->dtype == CNM_F16F8,
              CNM_ERR_INVALID_ARGUMENT, "%s: dtype %d", fn, d->dtype);
  CNM_REQUIRE(d->dtype == CNM_F16F8 ? (d->acc_exp >= -60 && d->acc_exp <= 60 && d->num_steps % 2 == 0) : d->acc_exp == 0,
              CNM_ERR_INVALID_ARGUMENT, "%s: acc_exp %d / num_steps %d do not fit dtype %d", fn, d->acc_exp, d->num_steps,
              d->dtype);
  CNM_REQUIRE(d->N > 0 && d->Ho > 0 && d->Wo > 0 && d->Cout > 0, CNM_ERR_INVALID_ARGUMENT, "%s: bad output shape", fn);
  CNM_REQUIRE(d->stride == 1 || d->stride == 2, CNM_ERR_UNSUPPORTED_SHAPE, "%s: stride %d", fn, d->stride);
  const bool fold = d->fold != 0;   // cnm_conv2d_halo / _halo2 only (they check the rest); up to 12 sources, views allowed
  CNM_REQUIRE(d->num_srcs >= 1 && d->num_srcs <= (fold ? 12 : 3), CNM_ERR_UNSUPPORTED_SHAPE, "%s: num_srcs %d (max %d)", fn,
              d->num_srcs, fold ? 12 : 3);
  CNM_REQUIRE(!fold || (d->Cout % 256 == 0 && d->stride == 1 && !d->head_partial && d->ksplit <= 1 && !d->balance &&
                        d->dtype != CNM_TF32),
              CNM_ERR_UNSUPPORTED_SHAPE, "%s: fold needs a stride-1 layer of 4 x (multiple of 64) channels, a 2-byte dtype, no head / split-K / balance", fn);
  CNM_REQUIRE(d->num_steps >= 1, CNM_ERR_INVALID_ARGUMENT, "%s: num_steps %d", fn, d->num_steps);
  CNM_REQUIRE((d->head_w == nullptr) == (d->head_partial == nullptr), CNM_ERR_INVALID_ARGUMENT,
              "%s: head_w and head_partial go together", fn);
  CNM_REQUIRE(!d->head_partial || (d->dtype != CNM_TF32 && ((uintptr_t)d->head_partial & 15) == 0 &&
                                   ((uintptr_t)d->head_w & 15) == 0),
              CNM_ERR_UNSUPPORTED_SHAPE, "%s: the fused head needs a 2-byte dtype and 16B-aligned buffers", fn);
  CNM_REQUIRE((d->side_disp == nullptr) == (d->side_w == nullptr), CNM_ERR_INVALID_ARGUMENT, "%s: side_disp and side_w go together", fn);
  CNM_REQUIRE(!d->side_disp || (d->dtype != CNM_TF32 && d->ksplit <= 1 && d->Ho % 2 == 0 && d->Wo % 2 == 0 && ((uintptr_t)d->side_w & 15) == 0),
              CNM_ERR_UNSUPPORTED_SHAPE, "%s: the side disparity input needs a 2-byte dtype, no split-K, an even output size and 16B-aligned weights", fn);
  const int ch = chunk_channels(d->dtype);
  CNM_REQUIRE(d->Cout % ch == 0 && d->c_out_total % ch == 0 && d->c_out_offset % ch == 0 &&
                  d->c_out_offset + phys_factor(d->dtype) * (fold ? d->Cout / 4 : d->Cout) <= d->c_out_total,
              CNM_ERR_UNSUPPORTED_SHAPE, "%s: Cout=%d / c_out_total=%d / offset=%d must be chunk (%d) aligned", fn,
              d->Cout, d->c_out_total, d->c_out_offset, ch);
  for (int i = 0; i < d->num_srcs; ++i) {
    CNM_REQUIRE(srcs[i].ptr && srcs[i].C > 0 && srcs[i].C % ch == 0, CNM_ERR_UNSUPPORTED_SHAPE,
                "%s: source %d channels %d must be a multiple of %d", fn, i, srcs[i].C, ch);
    CNM_REQUIRE(((uintptr_t)srcs[i].ptr & 127) == 0, CNM_ERR_INVALID_ARGUMENT, "%s: source %d must be 128B aligned", fn, i);
    const bool view = srcs[i].x_off || srcs[i].y_off || srcs[i].pitch_w || srcs[i].pitch_hw;
    CNM_REQUIRE(!view || fold, CNM_ERR_UNSUPPORTED_SHAPE, "%s: source %d is a view (fold layers of the row-halo kernels only)", fn, i);
    if (view)
      CNM_REQUIRE(srcs[i].x_off >= 0 && srcs[i].y_off >= 0 && srcs[i].W > 0 && srcs[i].H > 0 && srcs[i].x_off + srcs[i].W <= d->Wo &&
                      srcs[i].y_off + srcs[i].H <= d->Ho && srcs[i].pitch_w >= srcs[i].W && srcs[i].pitch_hw >= srcs[i].pitch_w * srcs[i].H,
                  CNM_ERR_INVALID_ARGUMENT, "%s: source %d: bad view", fn, i);
    else if (d->stride == 1)
      CNM_REQUIRE(srcs[i].apron >= 0 && srcs[i].apron <= 4 && srcs[i].H == d->Ho + 2 * srcs[i].apron && srcs[i].W == d->Wo,
                  CNM_ERR_UNSUPPORTED_SHAPE, "%s: source %d is %dx%d (apron %d), output %dx%d", fn, i, srcs[i].H, srcs[i].W,
                  srcs[i].apron, d->Ho, d->Wo);
    else
      CNM_REQUIRE((srcs[i].H + 1) / 2 == d->Ho && (srcs[i].W + 1) / 2 == d->Wo, CNM_ERR_UNSUPPORTED_SHAPE,
                  "%s: stride-2 source %d is %dx%d, output %dx%d", fn, i, srcs[i].H, srcs[i].W, d->Ho, d->Wo);
  }
  CNM_REQUIRE(((uintptr_t)wp & 127) == 0 && ((uintptr_t)out & 127) == 0, CNM_ERR_INVALID_ARGUMENT,
              "%s: wp/out must be 128B aligned", fn);
  return CNM_OK;
}

}  // namespace cnm

using namespace cnm;

extern "C" int cnm_conv2d_simt(const cnm_conv_desc* d, const cnm_conv_src* srcs, const cnm_conv_step* steps,
                               const void* wp, const float* bias, void* out, void* stream) {
  int rc = check_conv_args("cnm_conv2d_simt", d, srcs, steps, wp, out);
  if (rc) return rc;
  CNM_REQUIRE(!d->side_disp, CNM_ERR_UNSUPPORTED_SHAPE, "cnm_conv2d_simt: no side disparity input (list it as a taps source instead)");
  CNM_REQUIRE(!d->head_partial && d->ksplit <= 1, CNM_ERR_UNSUPPORTED_SHAPE,
              "cnm_conv2d_simt: no fused head / split-K in the validation twin");
  SimtParams p{};
  for (int i = 0; i < d->num_srcs; ++i) p.srcs[i] = SimtSrc{srcs[i].ptr, srcs[i].H, srcs[i].W, srcs[i].C};
  p.N = d->N; p.Ho = d->Ho; p.Wo = d->Wo; p.Cout = d->Cout;
  p.c_out_total = d->c_out_total; p.c_out_offset = d->c_out_offset;
  p.stride = d->stride; p.relu = d->relu; p.num_steps = d->num_steps;
  p.split = d->dtype == CNM_BF16X3;
  p.acc_scale = ldexpf(1.0f, -d->acc_exp);
  const size_t M = (size_t)d->N * d->Ho * d->Wo;
  dim3 grid((unsigned)((M + SM_TM - 1) / SM_TM), (d->Cout + SM_TN - 1) / SM_TN);
  if (d->dtype == CNM_F16F8)
    conv_simt_f16f8_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(p, steps, (const uint8_t*)wp, bias, (h16f8*)out);
  else if (d->dtype != CNM_TF32)
    conv_simt_kernel<__nv_bfloat16><<<grid, 256, 0, (cudaStream_t)stream>>>(p, steps, (const __nv_bfloat16*)wp, bias,
                                                                           (__nv_bfloat16*)out);
  else
    conv_simt_kernel<float><<<grid, 256, 0, (cudaStream_t)stream>>>(p, steps, (const float*)wp, bias, (float*)out);
  return after_launch("cnm_conv2d_simt");
}
