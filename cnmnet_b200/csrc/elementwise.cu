// K4 (refine-input aggregation) and the small bandwidth-bound kernels around the convolutions:
// layout converters, tap packers, bilinear x2 upsample, 1-channel 3x3 heads.
//
// Reference semantics: depthnet/depthNet_model.py:82-84 (depth_layer), :94,105 (nn.Upsample),
// :233,247-258 (torch.cat / F.upsample of the inputs), :332-333 (refine input), eval.py:659-663,
// 920-929 (view-group means).
#include <cstring>

#include "tc_common.cuh"

namespace cnm {

// ---- NCHW fp32 <-> act (NHWC) ----------------------------------------------------------------
template <typename T>
__global__ void nchw_to_act_kernel(const float* __restrict__ src, T* __restrict__ dst, int C, int H, int W, int ct,
                                   int co) {
  // tile transpose through smem: 32 pixels x 32 channels
  __shared__ float tile[32][33];
  const int n = blockIdx.z;
  const size_t HW = (size_t)H * W;
  const size_t p0 = (size_t)blockIdx.x * 32;
  const int c0 = blockIdx.y * 32;
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const int c = c0 + j;
    const size_t p = p0 + threadIdx.x;
    tile[j][threadIdx.x] = (c < C && p < HW) ? __ldg(src + ((size_t)n * C + c) * HW + p) : 0.0f;
  }
  __syncthreads();
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const size_t p = p0 + j;
    const int c = c0 + threadIdx.x;
    if (p < HW && c < C) store_scalar<T>(dst + ((size_t)n * HW + p) * ActTraits<T>::pstride(ct), co + c, tile[threadIdx.x][j]);
  }
}

template <typename T>
__global__ void act_to_nchw_kernel(const T* __restrict__ src, float* __restrict__ dst, int C, int H, int W, int ct,
                                   int co) {
  __shared__ float tile[32][33];
  const int n = blockIdx.z;
  const size_t HW = (size_t)H * W;
  const size_t p0 = (size_t)blockIdx.x * 32;
  const int c0 = blockIdx.y * 32;
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const size_t p = p0 + j;
    const int c = c0 + threadIdx.x;
    tile[j][threadIdx.x] = (p < HW && c < C) ? load_scalar<T>(src + ((size_t)n * HW + p) * ActTraits<T>::pstride(ct), co + c) : 0.0f;
  }
  __syncthreads();
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const int c = c0 + j;
    const size_t p = p0 + threadIdx.x;
    if (c < C && p < HW) dst[((size_t)n * C + c) * HW + p] = tile[threadIdx.x][j];
  }
}

// ---- tap packers -----------------------------------------------------------------------------
// One thread per pixel computes one K-chunk row (64 logical channels, 32 for tf32; only the first few are non-zero).
// The rows of a 128-thread block are contiguous in memory, so they are packed into shared memory (16-byte pieces,
// XOR-swizzled by pixel so that both phases are bank-conflict free) and written out as one coalesced stream; a thread
// storing its own 128 / 256 bytes directly would touch 32 different 256-byte segments per instruction.
constexpr int PACK_THREADS = 128;

// dst_block = first row of this block; `valid` rows exist (the last block of the grid may be partial)
template <typename T>
__device__ __forceinline__ void emit_rows(void* dst_block, const float* v, int valid) {
  constexpr int NP = RowPack<T>::PIECES;
  __shared__ uint4 stage[PACK_THREADS * NP];
  const int t = threadIdx.x;
#pragma unroll
  for (int j = 0; j < NP; ++j) stage[t * NP + (j ^ (t & (NP - 1)))] = RowPack<T>::piece(v, j);
  __syncthreads();
  uint4* out = reinterpret_cast<uint4*>(dst_block);
#pragma unroll
  for (int i = 0; i < NP; ++i) {
    const int l = i * PACK_THREADS + t, p = l / NP, j = (l % NP) ^ (p & (NP - 1));
    if (p < valid) out[p * NP + j] = stage[l];
  }
}

template <typename T, int KW, int CX>
__global__ void __launch_bounds__(PACK_THREADS)
    pack_row_taps_kernel(const float* __restrict__ x, T* __restrict__ dst, int N, int Nsrc, int H, int W) {
  pdl_trigger();
  pdl_wait();
  constexpr int CH = ActTraits<T>::CHUNK;
  static_assert(KW * CX <= CH, "taps must fit one chunk");
  const size_t HW = (size_t)H * W, total = (size_t)N * HW;
  const size_t i0 = (size_t)blockIdx.x * PACK_THREADS, i = i0 + threadIdx.x;
  float v[CH];
#pragma unroll
  for (int j = 0; j < CH; ++j) v[j] = 0.0f;
  if (i < total) {
    const int n = (int)(i / HW) % Nsrc;
    const size_t p = i % HW;
    const int y = (int)(p / W), xx = (int)(p % W);
    const float* row = x + (size_t)n * CX * HW + (size_t)y * W;
#pragma unroll
    for (int s = 0; s < KW; ++s) {
      const int xs = xx + s - KW / 2;
      const bool ok = xs >= 0 && xs < W;
#pragma unroll
      for (int c = 0; c < CX; ++c) v[s * CX + c] = ok ? __ldg(row + (size_t)c * HW + xs) : 0.0f;
    }
  }
  emit_rows<T>(dst + i0 * ActTraits<T>::pstride(CH), v, (int)(total - i0 < (size_t)PACK_THREADS ? total - i0 : (size_t)PACK_THREADS));
}

// three kernel rows per chunk: dst [N, H+2, W, CH], row j centred on image row j-1 (see include/cnmnet_b200.h)
template <typename T, int KW, int CX>
__global__ void __launch_bounds__(PACK_THREADS)
    pack_row_taps3_kernel(const float* __restrict__ x, T* __restrict__ dst, int N, int Nsrc, int H, int W) {
  pdl_trigger();
  pdl_wait();
  constexpr int CH = ActTraits<T>::CHUNK;
  static_assert(3 * KW * CX <= CH, "three rows of taps must fit one chunk");
  const size_t HW = (size_t)H * W, H2W = (size_t)(H + 2) * W, total = (size_t)N * H2W;
  const size_t i0 = (size_t)blockIdx.x * PACK_THREADS, i = i0 + threadIdx.x;
  float v[CH];
#pragma unroll
  for (int j = 0; j < CH; ++j) v[j] = 0.0f;
  if (i < total) {
    const int n = (int)(i / H2W) % Nsrc;
    const size_t p = i % H2W;
    const int j = (int)(p / W), xx = (int)(p % W);
#pragma unroll
    for (int q = 0; q < 3; ++q) {
      const int y = j - 2 + q;
      if (y < 0 || y >= H) continue;
      const float* row = x + (size_t)n * CX * HW + (size_t)y * W;
#pragma unroll
      for (int s = 0; s < KW; ++s) {
        const int xs = xx + s - KW / 2;
        const bool ok = xs >= 0 && xs < W;
#pragma unroll
        for (int c = 0; c < CX; ++c) v[(q * KW + s) * CX + c] = ok ? __ldg(row + (size_t)c * HW + xs) : 0.0f;
      }
    }
  }
  emit_rows<T>(dst + i0 * ActTraits<T>::pstride(CH), v, (int)(total - i0 < (size_t)PACK_THREADS ? total - i0 : (size_t)PACK_THREADS));
}

template <typename T>
__global__ void __launch_bounds__(PACK_THREADS)
    pack_disp_taps_kernel(const float* __restrict__ disp, T* __restrict__ dst, int N, int h, int w) {
  pdl_trigger();
  pdl_wait();
  constexpr int CH = ActTraits<T>::CHUNK;
  const int H = 2 * h, W = 2 * w;
  const size_t HW = (size_t)H * W, total = (size_t)N * HW;
  const size_t i0 = (size_t)blockIdx.x * PACK_THREADS, i = i0 + threadIdx.x;
  float v[CH];
#pragma unroll
  for (int j = 0; j < CH; ++j) v[j] = 0.0f;
  if (i < total) {
    const int n = (int)(i / HW);
    const size_t p = i % HW;
    const int y = (int)(p / W), x = (int)(p % W);
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
      for (int s = 0; s < 3; ++s) {
        const int yy = y + r - 1, xx = x + s - 1;
        if (yy >= 0 && yy < H && xx >= 0 && xx < W)
          v[3 * r + s] = __ldg(disp + ((size_t)n * h + (yy >> 1)) * w + (xx >> 1));  // nearest x2
      }
  }
  emit_rows<T>(dst + i0 * ActTraits<T>::pstride(CH), v, (int)(total - i0 < (size_t)PACK_THREADS ? total - i0 : (size_t)PACK_THREADS));
}

// ---- refine-input aggregation ------------------------------------------------------------------
// taps: channel 9c+3r+s of (id1, id2, |id1-id2|) with group means taken first
__device__ __forceinline__ float group_mean_at(const float* __restrict__ idepth, int g, int group, size_t BHW,
                                               size_t off) {
  float acc = __ldg(idepth + (size_t)g * BHW + off);
  for (int k = 1; k < group; ++k) acc += __ldg(idepth + (size_t)(g + 2 * k) * BHW + off);
  return group > 1 ? acc / (float)group : acc;
}

template <typename T>
__global__ void __launch_bounds__(PACK_THREADS)
    refine_taps_kernel(const float* __restrict__ idepth, T* __restrict__ taps, float* __restrict__ id_out, int B, int H,
                       int W, int group) {
  pdl_trigger();
  pdl_wait();
  constexpr int CH = ActTraits<T>::CHUNK;
  const size_t HW = (size_t)H * W, BHW = (size_t)B * HW;
  const size_t i0 = (size_t)blockIdx.x * PACK_THREADS, i = i0 + threadIdx.x;
  float v[CH];
#pragma unroll
  for (int j = 0; j < CH; ++j) v[j] = 0.0f;
  if (i < BHW) {
    const int b = (int)(i / HW);
    const size_t p = i % HW;
    const int y = (int)(p / W), x = (int)(p % W);
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
      for (int s = 0; s < 3; ++s) {
        const int yy = y + r - 1, xx = x + s - 1;
        if (yy >= 0 && yy < H && xx >= 0 && xx < W) {
          const size_t off = (size_t)b * HW + (size_t)yy * W + xx;
          const float a = group_mean_at(idepth, 0, group, BHW, off), c = group_mean_at(idepth, 1, group, BHW, off);
          v[3 * r + s] = a;
          v[9 + 3 * r + s] = c;
          v[18 + 3 * r + s] = fabsf(a - c);
          if (r == 1 && s == 1 && id_out) { id_out[off] = a; id_out[BHW + off] = c; }
        }
      }
  }
  emit_rows<T>(taps + i0 * ActTraits<T>::pstride(CH), v, (int)(BHW - i0 < (size_t)PACK_THREADS ? BHW - i0 : (size_t)PACK_THREADS));
}

// sum_act = mean_0(iconv) + mean_1(iconv) on act tensors [V*B,H,W,64]; one 16-byte vector per thread
template <typename T>
__global__ void refine_sum_act_kernel(const T* __restrict__ iconv, T* __restrict__ out, size_t per_view, int group) {
  pdl_trigger();
  pdl_wait();
  constexpr int VN = Vec16<T>::N;
  const size_t i = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * VN;
  if (i >= per_view) return;
  float m[2][VN];
#pragma unroll
  for (int g = 0; g < 2; ++g) {
    // one pixel = one 64-channel logical chunk, so the flat logical index maps to the physical one like a channel
    Vec16<T>::load(iconv + ActTraits<T>::phys((size_t)g * per_view + i), m[g]);
    for (int k = 1; k < group; ++k) {
      float t[VN];
      Vec16<T>::load(iconv + ActTraits<T>::phys((size_t)(g + 2 * k) * per_view + i), t);
#pragma unroll
      for (int j = 0; j < VN; ++j) m[g][j] += t[j];
    }
    if (group > 1)
#pragma unroll
      for (int j = 0; j < VN; ++j) m[g][j] /= (float)group;
  }
#pragma unroll
  for (int j = 0; j < VN; ++j) m[0][j] += m[1][j];
  Vec16<T>::store(out + ActTraits<T>::phys(i), m[0]);
}

// public-boundary variant: iconv NCHW fp32 [2,B,64,H,W] -> sum act [B,H,W,64]
template <typename T>
__global__ void refine_sum_nchw_kernel(const float* __restrict__ iconv, T* __restrict__ out, int B, int H, int W) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z;
  const size_t HW = (size_t)H * W;
  const size_t p0 = (size_t)blockIdx.x * 32;
  const int c0 = blockIdx.y * 32;
  const size_t view = (size_t)B * 64 * HW;
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const size_t p = p0 + threadIdx.x;
    const size_t off = ((size_t)b * 64 + c0 + j) * HW + p;
    tile[j][threadIdx.x] = (p < HW) ? __ldg(iconv + off) + __ldg(iconv + view + off) : 0.0f;
  }
  __syncthreads();
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const size_t p = p0 + j;
    if (p < HW) store_scalar<T>(out + ((size_t)b * HW + p) * ActTraits<T>::pstride(64), c0 + threadIdx.x, tile[threadIdx.x][j]);
  }
}

// ---- bilinear x2 upsample (align_corners=False): weights 0.25/0.75, clamped at the border ------
// one thread per (low-res pixel, 16-byte channel vector): 9 clamped loads in flight -> the 2x2 output quad
template <typename T>
__global__ void __launch_bounds__(256) upsample2x_kernel(const T* __restrict__ src, T* __restrict__ dst, int N, int h, int w, int C) {
  pdl_trigger();
  pdl_wait();
  constexpr int VN = Vec16<T>::N;
  const int cv = C / VN;
  const size_t total = (size_t)N * h * w * cv;
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int c = (int)(i % cv) * VN;
  size_t t = i / cv;
  const int x = (int)(t % w);
  t /= w;
  const int y = (int)(t % h);
  const int n = (int)(t / h);
  const int xm = max(x - 1, 0), xp = min(x + 1, w - 1), ym = max(y - 1, 0), yp = min(y + 1, h - 1);
  const size_t PC = ActTraits<T>::pstride(C), pc = ActTraits<T>::phys(c);
  const T* base = src + (size_t)n * h * w * PC + pc;
  float v[3][3][VN];
  const int ys[3] = {ym, y, yp}, xs[3] = {xm, x, xp};
#pragma unroll
  for (int a = 0; a < 3; ++a)
#pragma unroll
    for (int b = 0; b < 3; ++b) Vec16<T>::load(base + ((size_t)ys[a] * w + xs[b]) * PC, v[a][b]);
  // output (2y+dy, 2x+dx): source coordinate y - 0.25 (dy=0) or y + 0.25 (dy=1): rows (y-1,y) w (.25,.75) or (y,y+1) w (.75,.25)
  const int W2 = 2 * w;
  T* out = dst + (((size_t)n * 2 * h + 2 * y) * W2 + 2 * x) * PC + pc;
#pragma unroll
  for (int dy = 0; dy < 2; ++dy)
#pragma unroll
    for (int dx = 0; dx < 2; ++dx) {
      const float fx = dx ? 0.25f : 0.75f, fy = dy ? 0.25f : 0.75f;  // fraction toward the right / lower tap
      float o[VN];
#pragma unroll
      for (int j = 0; j < VN; ++j) {
        const float a00 = v[dy][dx][j], a01 = v[dy][dx + 1][j], a10 = v[dy + 1][dx][j], a11 = v[dy + 1][dx + 1][j];
        o[j] = a00 * ((1.f - fx) * (1.f - fy)) + a01 * (fx * (1.f - fy)) + a10 * ((1.f - fx) * fy) + a11 * (fx * fy);
      }
      Vec16<T>::store(out + ((size_t)dy * W2 + dx) * PC, o);
    }
}

// CNM_F16F8 specialisation: the generic kernel is instruction-bound in this layout (address-derived fp8 row pointers,
// 64-bit index arithmetic per load, saturation clamps); here every offset is a 32-bit byte offset inside the image, the
// hi8 bytes come from the fp16 pair through the f16x2 -> e4m3x2 conversion, and no clamp is needed (a convex
// combination of finite fp16 values stays finite).
__device__ __forceinline__ void load_h16f8_vec(const uint8_t* __restrict__ hi, const uint8_t* __restrict__ lo, float* f) {
  const uint4 v = __ldg(reinterpret_cast<const uint4*>(hi));
  const uint2 l = __ldg(reinterpret_cast<const uint2*>(lo));
  const uint32_t w[4] = {v.x, v.y, v.z, v.w};
  const uint32_t lw[4] = {l.x & 0xffffu, l.x >> 16, l.y & 0xffffu, l.y >> 16};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 h = __half22float2(*reinterpret_cast<const __half2*>(&w[i]));
    const float2 q = e4m3x2_to_float2(lw[i]);
    f[2 * i] = fmaf(q.x, 1.0f / F8_LO_SCALE, h.x);
    f[2 * i + 1] = fmaf(q.y, 1.0f / F8_LO_SCALE, h.y);
  }
}
__device__ __forceinline__ void store_h16f8_vec(uint8_t* __restrict__ hi, uint8_t* __restrict__ lo, const float* f) {
  uint32_t w[4], lw[4], hw[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const __half2 h = __floats2half2_rn(f[2 * i], f[2 * i + 1]);
    const float2 hf = __half22float2(h);
    w[i] = *reinterpret_cast<const uint32_t*>(&h);
    lw[i] = e4m3x2((f[2 * i] - hf.x) * F8_LO_SCALE, (f[2 * i + 1] - hf.y) * F8_LO_SCALE);
    const __half2 hs = __hmul2(h, __float2half2_rn(F8_HI_SCALE));  // exact unless far below the e4m3 range
    hw[i] = (uint32_t)__nv_cvt_halfraw2_to_fp8x2(*reinterpret_cast<const __half2_raw*>(&hs), __NV_SATFINITE, __NV_E4M3);
  }
  *reinterpret_cast<uint4*>(hi) = make_uint4(w[0], w[1], w[2], w[3]);
  *reinterpret_cast<uint2*>(lo) = make_uint2(lw[0] | (lw[1] << 16), lw[2] | (lw[3] << 16));
  *reinterpret_cast<uint2*>(lo + 64) = make_uint2(hw[0] | (hw[1] << 16), hw[2] | (hw[3] << 16));
}

__global__ void __launch_bounds__(256)
    upsample2x_h16f8_kernel(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, int N, int h, int w, int C) {
  pdl_trigger();
  pdl_wait();
  const int cv = C >> 3;
  const unsigned total = (unsigned)(N * h * w * cv);
  const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int c8 = (int)(i % cv);
  unsigned t = i / cv;
  const int x = (int)(t % w);
  t /= w;
  const int y = (int)(t % h);
  const int n = (int)(t / h);
  const unsigned PB = 4u * C;                                          // bytes per pixel (2 x 2 bytes per channel)
  const unsigned hi_off = (unsigned)(c8 >> 3) * 256u + (unsigned)(c8 & 7) * 16u;   // inside the pixel
  const unsigned lo_off = (unsigned)(c8 >> 3) * 256u + 128u + (unsigned)(c8 & 7) * 8u;
  const uint8_t* img = src + (size_t)n * h * w * PB;
  const int xs[3] = {max(x - 1, 0), x, min(x + 1, w - 1)}, ys[3] = {max(y - 1, 0), y, min(y + 1, h - 1)};
  float v[3][3][8];
#pragma unroll
  for (int a = 0; a < 3; ++a)
#pragma unroll
    for (int b = 0; b < 3; ++b) {
      const unsigned po = (unsigned)(ys[a] * w + xs[b]) * PB;
      load_h16f8_vec(img + po + hi_off, img + po + lo_off, v[a][b]);
    }
  const unsigned W2 = 2u * w;
  uint8_t* out = dst + (size_t)n * 4 * h * w * PB;
#pragma unroll
  for (int dy = 0; dy < 2; ++dy)
#pragma unroll
    for (int dx = 0; dx < 2; ++dx) {
      const float fx = dx ? 0.25f : 0.75f, fy = dy ? 0.25f : 0.75f;  // fraction toward the right / lower tap
      float o[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float a00 = v[dy][dx][j], a01 = v[dy][dx + 1][j], a10 = v[dy + 1][dx][j], a11 = v[dy + 1][dx + 1][j];
        o[j] = a00 * ((1.f - fx) * (1.f - fy)) + a01 * (fx * (1.f - fy)) + a10 * ((1.f - fx) * fy) + a11 * (fx * fy);
      }
      const unsigned po = ((unsigned)(2 * y + dy) * W2 + (unsigned)(2 * x + dx)) * PB;
      store_h16f8_vec(out + po + hi_off, out + po + lo_off, o);
    }
}

// ---- 1-channel 3x3 head --------------------------------------------------------------------------------------
// CTA = 32 x TR output pixels.  The (TR+2) x 34 halo of the input (all channels) is staged in shared memory by TMA
// (one or two 4-D boxes {<=256 ch, 34, TR+2, 1}; pixels outside the image are zero-filled = conv padding), so the
// staging costs one instruction and each input pixel crosses L2 -> SM ~1.3x.  LPP lanes per pixel, one 16-byte
// vector per lane and tap, the lane's 9 x VN weights in registers when one vector per lane covers all channels;
// fp32 accumulate, shuffle reduce, bias + scale * sigmoid epilogue.
struct alignas(64) HeadParams {
  CUtensorMap x_map;
  const float* w;
  float* out;
  float* logits;
  float bias, scale;
  int H, W, C, TR, PC, boxc, nbox;
};

template <typename T, int LPP>
__global__ void __launch_bounds__(256, 2) head3x3_kernel(const __grid_constant__ HeadParams p) {
  pdl_trigger();
  pdl_wait();
  extern __shared__ uint8_t hsm_raw[];
  constexpr int VN = Vec16<T>::N;
  const int C = p.C, PC = p.PC, TR = p.TR, npix = (TR + 2) * 34;
  const uint32_t base = (smem_u32(hsm_raw) + 127u) & ~127u;
  uint8_t* hsm = hsm_raw + (base - smem_u32(hsm_raw));
  T* tile = reinterpret_cast<T*>(hsm);                                   // [nbox][(TR+2)][34][boxc]
  const size_t tile_bytes = (size_t)npix * PC * sizeof(T);
  float* ws = reinterpret_cast<float*>(hsm + tile_bytes);                 // [9*C] in (r, s, c) order
  const uint32_t bar = base + (uint32_t)tile_bytes + (uint32_t)(9 * C * 4);
  const int n = blockIdx.z, x0 = blockIdx.x * 32, y0 = blockIdx.y * TR;
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    mbar_expect_tx(bar, (uint32_t)tile_bytes);
    for (int b = 0; b < p.nbox; ++b)
      tma_load_4d(base + (uint32_t)((size_t)b * npix * p.boxc * sizeof(T)), &p.x_map, bar, b * p.boxc, x0 - 1, y0 - 1, n);
  }
  for (int i = threadIdx.x; i < 9 * C; i += 256) ws[i] = __ldg(p.w + i);
  __syncthreads();
  mbar_wait(bar, 0);
  // physical channel q of halo pixel pi
  const int boxc = p.boxc;
  auto toff = [&](int pi, size_t q) { return (q / boxc) * ((size_t)npix * boxc) + (size_t)pi * boxc + (q % boxc); };

  const int sub = threadIdx.x % LPP, grp = threadIdx.x / LPP;
  constexpr int GROUPS = 256 / LPP;
  const bool wreg = (C == LPP * VN);
  float wr[9][VN];
  if (wreg) {
#pragma unroll
    for (int t = 0; t < 9; ++t)
#pragma unroll
      for (int j = 0; j < VN; ++j) wr[t][j] = ws[t * C + sub * VN + j];
  }
  for (int pi = grp; pi < TR * 32; pi += GROUPS) {  // uniform trip count: TR*32 is a multiple of GROUPS
    const int ty = pi / 32, tx = pi % 32;
    float acc = 0.0f;
    if (wreg) {
      const size_t q = ActTraits<T>::phys(sub * VN);
#pragma unroll
      for (int t = 0; t < 9; ++t) {
        float v[VN];
        Vec16<T>::load(tile + toff((ty + t / 3) * 34 + tx + t % 3, q), v);
#pragma unroll
        for (int j = 0; j < VN; ++j) acc = fmaf(v[j], wr[t][j], acc);
      }
    } else {
      for (int c = sub * VN; c < C; c += LPP * VN) {
        const size_t q = ActTraits<T>::phys(c);
#pragma unroll
        for (int t = 0; t < 9; ++t) {
          float v[VN];
          Vec16<T>::load(tile + toff((ty + t / 3) * 34 + tx + t % 3, q), v);
          const float* wp = ws + t * C + c;
#pragma unroll
          for (int j = 0; j < VN; j += 4) {
            const float4 w4 = *reinterpret_cast<const float4*>(wp + j);
            acc = fmaf(v[j], w4.x, acc);
            acc = fmaf(v[j + 1], w4.y, acc);
            acc = fmaf(v[j + 2], w4.z, acc);
            acc = fmaf(v[j + 3], w4.w, acc);
          }
        }
      }
    }
#pragma unroll
    for (int o = 1; o < LPP; o <<= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    const int yy = y0 + ty, xx = x0 + tx;
    if (sub == 0 && yy < p.H && xx < p.W) {
      const size_t pix = ((size_t)n * p.H + yy) * p.W + xx;
      const float l = acc + p.bias;
      if (p.logits) p.logits[pix] = l;
      p.out[pix] = p.scale / (1.0f + expf(-l));
    }
  }
}

// second half of a head fused into the producing convolution's epilogue (tc_common.cuh: epi_head_partial)
__global__ void __launch_bounds__(256) head_finish_kernel(const float* __restrict__ partial, int nslices, float bias,
                                                          float scale, float* __restrict__ out,
                                                          float* __restrict__ logits, int N, int H, int W) {
  pdl_trigger();
  pdl_wait();
  const size_t total = (size_t)N * H * W;
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int x = (int)(i % W), y = (int)((i / W) % H);
  float acc = bias;
#pragma unroll
  for (int r = 0; r < 3; ++r)
#pragma unroll
    for (int s = 0; s < 3; ++s) {
      const int yy = y + r - 1, xx = x + s - 1;
      if (yy >= 0 && yy < H && xx >= 0 && xx < W) {
        const float* q = partial + (i + (size_t)((r - 1) * W + (s - 1))) * HEAD_PARTIAL_STRIDE + 3 * r + s;
        float t = 0.0f;
        for (int k = 0; k < nslices; ++k) t += __ldg(q + (size_t)k * total * HEAD_PARTIAL_STRIDE);  // fixed order
        acc += t;
      }
    }
  if (logits) logits[i] = acc;
  out[i] = scale / (1.0f + expf(-acc));
}

// First half of a head as a stand-alone pass (layers whose head is not fused into the producing convolution): every
// pixel's C-vector is read and decoded ONCE and dotted with the nine tap weight vectors -> 12 floats per pixel (the
// layout epi_head_partial writes); head_finish_kernel does the 3x3 gather.  LPP lanes per pixel, 16-byte vectors.
template <typename T, int LPP>
__global__ void __launch_bounds__(256) head_partial_kernel(const T* __restrict__ x, const float* __restrict__ w,
                                                           float* __restrict__ partial, size_t npix, int C) {
  pdl_trigger();
  pdl_wait();
  extern __shared__ float hp_ws[];  // [9][C]
  constexpr int VN = Vec16<T>::N, GROUPS = 256 / LPP;
  for (int i = threadIdx.x; i < 9 * C; i += 256) hp_ws[i] = __ldg(w + i);
  __syncthreads();
  const int sub = threadIdx.x % LPP, grp = threadIdx.x / LPP;
  const size_t PC = ActTraits<T>::pstride(C);
  for (size_t pix = (size_t)blockIdx.x * GROUPS + grp; pix < npix; pix += (size_t)gridDim.x * GROUPS) {
    float acc[9];
#pragma unroll
    for (int t = 0; t < 9; ++t) acc[t] = 0.0f;
    const T* px = x + pix * PC;
    for (int c = sub * VN; c < C; c += LPP * VN) {
      float v[VN];
      Vec16<T>::load(px + ActTraits<T>::phys(c), v);
#pragma unroll
      for (int t = 0; t < 9; ++t) {
        const float* wp = hp_ws + t * C + c;
#pragma unroll
        for (int j = 0; j < VN; j += 4) {
          const float4 w4 = *reinterpret_cast<const float4*>(wp + j);
          acc[t] = fmaf(v[j], w4.x, acc[t]);
          acc[t] = fmaf(v[j + 1], w4.y, acc[t]);
          acc[t] = fmaf(v[j + 2], w4.z, acc[t]);
          acc[t] = fmaf(v[j + 3], w4.w, acc[t]);
        }
      }
    }
#pragma unroll
    for (int t = 0; t < 9; ++t)
#pragma unroll
      for (int o = 1; o < LPP; o <<= 1) acc[t] += __shfl_xor_sync(0xffffffffu, acc[t], o);
    if (sub == 0) {
      float4* o = reinterpret_cast<float4*>(partial + pix * HEAD_PARTIAL_STRIDE);
      o[0] = make_float4(acc[0], acc[1], acc[2], acc[3]);
      o[1] = make_float4(acc[4], acc[5], acc[6], acc[7]);
      o[2] = make_float4(acc[8], 0.0f, 0.0f, 0.0f);
    }
  }
}

template <typename T>
static int launch_head_partial(const T* x, const float* w, float* partial, size_t npix, int C, cudaStream_t s) {
  const int lanes = C / Vec16<T>::N;
  const size_t smem = (size_t)9 * C * sizeof(float);
  long blocks = 148 * 8;
#define CNM_HP(LPP)                                                                         \
  {                                                                                         \
    const long need = (long)((npix + 256 / LPP - 1) / (256 / LPP));                         \
    if (need < blocks) blocks = need;                                                       \
    launch_pdl(head_partial_kernel<T, LPP>, dim3((unsigned)blocks), dim3(256), smem, s, x, w, partial, npix, C); \
  }
  if (lanes >= 32) CNM_HP(32) else if (lanes >= 16) CNM_HP(16) else CNM_HP(8)
#undef CNM_HP
  return CNM_OK;
}

template <typename T, int LPP>
static int launch_head_lpp(const HeadParams& p, dim3 grid, size_t smem, cudaStream_t s) {
  static DynSmemLimit limit;
  cudaError_t e = limit.ensure(head3x3_kernel<T, LPP>, smem);
  if (e != cudaSuccess) return fail(CNM_ERR_CUDA, "cnm_head3x3: smem attribute: %s", cudaGetErrorString(e));
  launch_pdl(head3x3_kernel<T, LPP>, dim3(grid), dim3(256), smem, s, p);
  return CNM_OK;
}

template <typename T>
static int launch_head(const T* x, const float* w, float bias, float scale, float* out, float* logits, int N, int H,
                       int W, int C, int dtype, cudaStream_t s) {
  const int lanes = C / Vec16<T>::N;  // one 16-byte vector per lane
  const int lpp = lanes >= 32 ? 32 : (lanes >= 16 ? 16 : 8);
  HeadParams p;
  memset(&p, 0, sizeof(p));
  p.w = w; p.out = out; p.logits = logits; p.bias = bias; p.scale = scale;
  p.H = H; p.W = W; p.C = C;
  p.PC = (int)ActTraits<T>::pstride(C);
  p.boxc = p.PC < 256 ? p.PC : 256;
  p.nbox = p.PC / p.boxc;
  // rows per CTA: staged halo tile of ~50 KB (four CTAs per SM), ~100 KB when a pixel is wider than 128 B so that
  // the two halo rows stay a small fraction of the tile
  const size_t px_bytes = (size_t)p.PC * sizeof(T);
  int TR = (int)((px_bytes > 128 ? 100 : 50) * 1024 / (34 * px_bytes)) - 2;
  TR = TR > 8 ? 8 : (TR < 1 ? 1 : TR);
  p.TR = TR;
  const uint64_t pix = (uint64_t)p.PC * sizeof(T);
  const uint64_t dims[4] = {(uint64_t)p.PC, (uint64_t)W, (uint64_t)H, (uint64_t)N};
  const uint64_t str[3] = {pix, pix * W, pix * W * H};
  const uint32_t box[4] = {(uint32_t)p.boxc, 34u, (uint32_t)(TR + 2), 1u};
  int rc = encode_map(&p.x_map, dtype, 4, x, dims, str, box, /*swizzle128=*/false);
  if (rc) return rc;
  const size_t smem = 128 + (size_t)(TR + 2) * 34 * p.PC * sizeof(T) + (size_t)9 * C * 4 + 16;
  dim3 grid((W + 31) / 32, (H + TR - 1) / TR, N);
  if (N > 65535 || grid.y > 65535) return fail(CNM_ERR_UNSUPPORTED_SHAPE, "cnm_head3x3: grid too large");
  if (lpp == 32) return launch_head_lpp<T, 32>(p, grid, smem, s);
  if (lpp == 16) return launch_head_lpp<T, 16>(p, grid, smem, s);
  return launch_head_lpp<T, 8>(p, grid, smem, s);
}

}  // namespace cnm

using namespace cnm;

#define DISPATCH_DTYPE(dtype, ...)                                   \
  if ((dtype) == CNM_BF16) { using T = __nv_bfloat16; __VA_ARGS__; } \
  else if ((dtype) == CNM_TF32) { using T = float; __VA_ARGS__; }    \
  else if ((dtype) == CNM_BF16X3) { using T = bf16s; __VA_ARGS__; }  \
  else if ((dtype) == CNM_F16F8) { using T = h16f8; __VA_ARGS__; }   \
  else return fail(CNM_ERR_INVALID_ARGUMENT, "%s: dtype %d", __func__, (int)(dtype));

extern "C" int cnm_nchw_to_act(const float* src, void* dst, int N, int C, int H, int W, int c_dst_total,
                               int c_dst_offset, int dtype, void* stream) {
  CNM_REQUIRE(src && dst && N > 0 && C > 0 && H > 0 && W > 0, CNM_ERR_INVALID_ARGUMENT, "cnm_nchw_to_act: bad argument");
  CNM_REQUIRE(c_dst_offset >= 0 && c_dst_offset + C <= c_dst_total, CNM_ERR_INVALID_ARGUMENT,
              "cnm_nchw_to_act: channel slice out of range");
  dim3 grid((unsigned)(((size_t)H * W + 31) / 32), (C + 31) / 32, N), block(32, 8);
  DISPATCH_DTYPE(dtype, (nchw_to_act_kernel<T><<<grid, block, 0, (cudaStream_t)stream>>>(src, (T*)dst, C, H, W,
                                                                                         c_dst_total, c_dst_offset)));
  return after_launch("cnm_nchw_to_act");
}

extern "C" int cnm_act_to_nchw(const void* src, float* dst, int N, int C, int H, int W, int c_src_total,
                               int c_src_offset, int dtype, void* stream) {
  CNM_REQUIRE(src && dst && N > 0 && C > 0 && H > 0 && W > 0, CNM_ERR_INVALID_ARGUMENT, "cnm_act_to_nchw: bad argument");
  CNM_REQUIRE(c_src_offset >= 0 && c_src_offset + C <= c_src_total, CNM_ERR_INVALID_ARGUMENT,
              "cnm_act_to_nchw: channel slice out of range");
  dim3 grid((unsigned)(((size_t)H * W + 31) / 32), (C + 31) / 32, N), block(32, 8);
  DISPATCH_DTYPE(dtype, (act_to_nchw_kernel<T><<<grid, block, 0, (cudaStream_t)stream>>>((const T*)src, dst, C, H, W,
                                                                                         c_src_total, c_src_offset)));
  return after_launch("cnm_act_to_nchw");
}

extern "C" int cnm_pack_row_taps(const float* x, void* dst, int N, int Nsrc, int Cx, int H, int W, int kw, int dtype,
                                 void* stream) {
  CNM_REQUIRE(x && dst && N > 0 && Nsrc > 0 && Cx > 0 && H > 0 && W > 0, CNM_ERR_INVALID_ARGUMENT,
              "cnm_pack_row_taps: bad argument");
  CNM_REQUIRE(kw == 7 && Cx == 3, CNM_ERR_UNSUPPORTED_SHAPE,
              "cnm_pack_row_taps: only kw=7, Cx=3 (conv1.0 of DepthNet) is instantiated, got kw=%d Cx=%d", kw, Cx);
  const size_t total = (size_t)N * H * W;
  DISPATCH_DTYPE(dtype, (launch_pdl(pack_row_taps_kernel<T, 7, 3>, dim3((unsigned)((total + 127) / 128)), dim3(128), 0, (cudaStream_t)stream, x, (T*)dst, N, Nsrc, H, W)));
  return after_launch("cnm_pack_row_taps");
}

extern "C" int cnm_pack_row_taps3(const float* x, void* dst, int N, int Nsrc, int Cx, int H, int W, int kw, int dtype,
                                  void* stream) {
  CNM_REQUIRE(x && dst && N > 0 && Nsrc > 0 && Cx > 0 && H > 0 && W > 0, CNM_ERR_INVALID_ARGUMENT,
              "cnm_pack_row_taps3: bad argument");
  CNM_REQUIRE(kw == 7 && Cx == 3 && dtype != CNM_TF32, CNM_ERR_UNSUPPORTED_SHAPE,
              "cnm_pack_row_taps3: only kw=7, Cx=3 in the 2-byte dtypes (63 of 64 channels), got kw=%d Cx=%d dtype=%d", kw, Cx, dtype);
  const size_t total = (size_t)N * (H + 2) * W;
  const dim3 grid((unsigned)((total + 127) / 128)), block(128);
  cudaStream_t s = (cudaStream_t)stream;
  if (dtype == CNM_BF16) launch_pdl(pack_row_taps3_kernel<__nv_bfloat16, 7, 3>, grid, block, 0, s, x, (__nv_bfloat16*)dst, N, Nsrc, H, W);
  else if (dtype == CNM_BF16X3) launch_pdl(pack_row_taps3_kernel<bf16s, 7, 3>, grid, block, 0, s, x, (bf16s*)dst, N, Nsrc, H, W);
  else if (dtype == CNM_F16F8) launch_pdl(pack_row_taps3_kernel<h16f8, 7, 3>, grid, block, 0, s, x, (h16f8*)dst, N, Nsrc, H, W);
  else return fail(CNM_ERR_INVALID_ARGUMENT, "cnm_pack_row_taps3: dtype %d", dtype);
  return after_launch("cnm_pack_row_taps3");
}

extern "C" int cnm_pack_disp_taps(const float* disp, void* dst, int N, int h, int w, int dtype, void* stream) {
  CNM_REQUIRE(disp && dst && N > 0 && h > 0 && w > 0, CNM_ERR_INVALID_ARGUMENT, "cnm_pack_disp_taps: bad argument");
  const size_t total = (size_t)N * 4 * h * w;
  DISPATCH_DTYPE(dtype, (launch_pdl(pack_disp_taps_kernel<T>, dim3((unsigned)((total + 127) / 128)), dim3(128), 0, (cudaStream_t)stream, disp, (T*)dst, N, h, w)));
  return after_launch("cnm_pack_disp_taps");
}

extern "C" int cnm_refine_aggregate(const float* idepth, const void* iconv, void* sum_act, void* taps_act,
                                    float* id_out, int B, int H, int W, int group, int dtype, void* stream) {
  CNM_REQUIRE(idepth && iconv && sum_act && taps_act, CNM_ERR_INVALID_ARGUMENT, "cnm_refine_aggregate: null pointer");
  CNM_REQUIRE(B > 0 && H > 0 && W > 0 && group >= 1 && group <= 8, CNM_ERR_INVALID_ARGUMENT,
              "cnm_refine_aggregate: bad shape");
  const size_t px = (size_t)B * H * W, per_view = px * 64;
  cudaStream_t s = (cudaStream_t)stream;
  DISPATCH_DTYPE(dtype, {
    launch_pdl(refine_taps_kernel<T>, dim3((unsigned)((px + 127) / 128)), dim3(128), 0, s, idepth, (T*)taps_act, id_out, B, H, W, group);
    int rc = after_launch("cnm_refine_aggregate(taps)");
    if (rc) return rc;
    const size_t nv = per_view / Vec16<T>::N;
    launch_pdl(refine_sum_act_kernel<T>, dim3((unsigned)((nv + 255) / 256)), dim3(256), 0, s, (const T*)iconv, (T*)sum_act, per_view, group);
  });
  return after_launch("cnm_refine_aggregate(sum)");
}

extern "C" int cnm_refine_aggregate_nchw(const float* idepth, const float* iconv_nchw, void* sum_act, void* taps_act,
                                         int B, int H, int W, int dtype, void* stream) {
  CNM_REQUIRE(idepth && iconv_nchw && sum_act && taps_act, CNM_ERR_INVALID_ARGUMENT,
              "cnm_refine_aggregate_nchw: null pointer");
  CNM_REQUIRE(B > 0 && H > 0 && W > 0, CNM_ERR_INVALID_ARGUMENT, "cnm_refine_aggregate_nchw: bad shape");
  const size_t px = (size_t)B * H * W;
  cudaStream_t s = (cudaStream_t)stream;
  DISPATCH_DTYPE(dtype, {
    launch_pdl(refine_taps_kernel<T>, dim3((unsigned)((px + 127) / 128)), dim3(128), 0, s, idepth, (T*)taps_act, nullptr, B, H, W, 1);
    int rc = after_launch("cnm_refine_aggregate_nchw(taps)");
    if (rc) return rc;
    dim3 grid((unsigned)(((size_t)H * W + 31) / 32), 2, B), block(32, 8);
    refine_sum_nchw_kernel<T><<<grid, block, 0, s>>>(iconv_nchw, (T*)sum_act, B, H, W);
  });
  return after_launch("cnm_refine_aggregate_nchw(sum)");
}

extern "C" int cnm_upsample2x(const void* src, void* dst, int N, int h, int w, int C, int dtype, void* stream) {
  CNM_REQUIRE(src && dst && N > 0 && h > 0 && w > 0 && C > 0, CNM_ERR_INVALID_ARGUMENT, "cnm_upsample2x: bad argument");
  CNM_REQUIRE(C % 8 == 0, CNM_ERR_UNSUPPORTED_SHAPE, "cnm_upsample2x: C=%d must be a multiple of 8", C);
  if (dtype == CNM_F16F8 && C % 64 == 0 && (size_t)N * 4 * h * w * 4 * C < ((size_t)1 << 32) * 64) {
    // 32-bit offsets inside one image: 4*h*w*4*C bytes per output image must fit 32 bits
    CNM_REQUIRE((size_t)4 * h * w * 4 * C < ((size_t)1 << 32) && (size_t)N * h * w * (C / 8) < ((size_t)1 << 32),
                CNM_ERR_UNSUPPORTED_SHAPE, "cnm_upsample2x: image too large for 32-bit offsets");
    const size_t total = (size_t)N * h * w * (C / 8);
    launch_pdl(upsample2x_h16f8_kernel, dim3((unsigned)((total + 255) / 256)), dim3(256), 0, (cudaStream_t)stream, (const uint8_t*)src, (uint8_t*)dst, N, h, w, C);
    return after_launch("cnm_upsample2x");
  }
  DISPATCH_DTYPE(dtype, {
    const size_t total = (size_t)N * h * w * (C / Vec16<T>::N);
    launch_pdl(upsample2x_kernel<T>, dim3((unsigned)((total + 255) / 256)), dim3(256), 0, (cudaStream_t)stream, (const T*)src, (T*)dst, N, h, w, C);
  });
  return after_launch("cnm_upsample2x");
}

extern "C" int cnm_head_partial(const void* x, const float* w_tap_major, float* partial, int N, int H, int W, int C,
                                int dtype, void* stream) {
  CNM_REQUIRE(x && w_tap_major && partial && N > 0 && H > 0 && W > 0, CNM_ERR_INVALID_ARGUMENT,
              "cnm_head_partial: bad argument");
  CNM_REQUIRE(C % 64 == 0 && C <= 1024, CNM_ERR_UNSUPPORTED_SHAPE, "cnm_head_partial: C=%d must be a multiple of 64", C);
  CNM_REQUIRE(((uintptr_t)partial & 15) == 0, CNM_ERR_INVALID_ARGUMENT, "cnm_head_partial: partial must be 16B aligned");
  DISPATCH_DTYPE(dtype, {
    int rc = launch_head_partial<T>((const T*)x, w_tap_major, partial, (size_t)N * H * W, C, (cudaStream_t)stream);
    if (rc) return rc;
  });
  return after_launch("cnm_head_partial");
}

extern "C" int cnm_head_finish(const float* partial, int nslices, float bias, float scale, float* out, float* logits, int N,
                               int H, int W, void* stream) {
  CNM_REQUIRE(partial && out && N > 0 && H > 0 && W > 0 && nslices >= 1 && nslices <= 64, CNM_ERR_INVALID_ARGUMENT,
              "cnm_head_finish: bad argument");
  const size_t total = (size_t)N * H * W;
  launch_pdl(head_finish_kernel, dim3((unsigned)((total + 255) / 256)), dim3(256), 0, (cudaStream_t)stream, partial, nslices, bias, scale, out,
                                                                                       logits, N, H, W);
  return after_launch("cnm_head_finish");
}

extern "C" int cnm_head3x3(const void* x, const float* w, float bias, float scale, float* out, float* logits, int N,
                           int H, int W, int C, int dtype, void* stream) {
  CNM_REQUIRE(x && w && out && N > 0 && H > 0 && W > 0, CNM_ERR_INVALID_ARGUMENT, "cnm_head3x3: bad argument");
  CNM_REQUIRE(C % 64 == 0 && C <= 1024, CNM_ERR_UNSUPPORTED_SHAPE, "cnm_head3x3: C=%d must be a multiple of 64", C);
  DISPATCH_DTYPE(dtype, {
    int rc = launch_head<T>((const T*)x, w, bias, scale, out, logits, N, H, W, C, dtype, (cudaStream_t)stream);
    if (rc) return rc;
  });
  return after_launch("cnm_head3x3");
}
