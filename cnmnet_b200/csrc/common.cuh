// Shared helpers for the cnmnet_b200 CUDA sources (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_fp8.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <cstdarg>
#include <cstdio>

#include "../../include/cnmnet_b200.h"

namespace cnm {

extern thread_local char g_err[512];
extern std::atomic<long long> g_launches;

inline int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

inline int after_launch(const char* what) {
  g_launches.fetch_add(1, std::memory_order_relaxed);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(CNM_ERR_CUDA, "%s: %s", what, cudaGetErrorString(e));
  return CNM_OK;
}

// cudaFuncAttributeMaxDynamicSharedMemorySize is a per-DEVICE attribute: callers keep one `static DynSmemLimit` per kernel
// instantiation and raise the limit on whichever device is current (a process may drive several GPUs).
struct DynSmemLimit {
  int set[64] = {0};
  template <typename K>
  cudaError_t ensure(K kernel, size_t smem) {
    int dev = 0;
    cudaGetDevice(&dev);
    int& cur = set[dev & 63];
    if ((int)smem <= cur) return cudaSuccess;
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess) cur = (int)smem;
    return e;
  }
};

// ---- programmatic dependent launch -------------------------------------------------------------------------------
// The kernels of a pipeline step run back to back on one stream (as a CUDA graph).  Launched through launch_pdl they
// carry cudaLaunchAttributeProgrammaticStreamSerialization: the next kernel's CTAs may be scheduled while this one is
// still running -- as soon as all of its CTAs have called pdl_trigger() and an SM has room -- and run their prologue
// (barrier init, TMEM allocation, descriptor prefetch) up to pdl_wait(), which blocks until the previous kernel has
// completed and its writes are visible.  EVERY global-memory access of a kernel launched this way comes after its
// pdl_wait().  CNM_PDL=0 turns the attribute off (plain stream order); the device-side calls are no-ops then.
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

bool pdl_enabled();  // api.cu

template <typename... P, typename... A>
inline cudaError_t launch_pdl(void (*kernel)(P...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, A&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<P>(args)...);
}

#define CNM_REQUIRE(cond, code, ...) \
  do {                               \
    if (!(cond)) return ::cnm::fail(code, __VA_ARGS__); \
  } while (0)

// K-chunk in PHYSICAL channels (one 128-byte operand row) and bytes per physical element
inline int chunk_channels(int dtype) { return dtype == CNM_TF32 ? 32 : 64; }
inline int elem_bytes(int dtype) { return dtype == CNM_TF32 ? 4 : 2; }
// physical channels per logical channel: CNM_BF16X3 stores every value as a (hi, lo) pair of bf16, CNM_F16F8 as
// fp16 hi + two fp8 bytes
inline int phys_factor(int dtype) { return (dtype == CNM_BF16X3 || dtype == CNM_F16F8) ? 2 : 1; }

// ---- act element access (T = __nv_bfloat16 or float) ----
template <typename T>
__device__ __forceinline__ float to_f(T v);
template <>
__device__ __forceinline__ float to_f<float>(float v) { return v; }
template <>
__device__ __forceinline__ float to_f<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }

template <typename T>
__device__ __forceinline__ T from_f(float v);
// fp32 act tensors are tcgen05 kind::tf32 operands: store them tf32-rounded (nearest, ties away) so that the
// tensor core's truncation of the low 13 mantissa bits is exact and every consumer sees the same value
__device__ __forceinline__ float round_tf32(float v) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v));
  return __uint_as_float(r);
}
template <>
__device__ __forceinline__ float from_f<float>(float v) { return round_tf32(v); }
template <>
__device__ __forceinline__ __nv_bfloat16 from_f<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

// Element of a CNM_BF16X3 ("split bf16") act tensor.  A value v is stored as hi = bf16(v), lo = bf16(v - hi) (16 mantissa
// bits together).  Layout: logical channel c of a pixel lives at physical channel (c/64)*128 + c%64 (hi) and 64 further
// (lo), i.e. every 64-channel logical chunk is a 128-byte hi row followed by a 128-byte lo row -- both are ordinary
// bf16 tcgen05 operand rows, so a convolution runs as three bf16 passes (hi*Whi + lo*Whi + hi*Wlo) of the same kernels.
struct bf16s {
  __nv_bfloat16 v;
};

template <typename T>
struct ActTraits {
  static constexpr bool SPLIT = false;
  static constexpr int CHUNK = 128 / sizeof(T);  // logical channels per K-chunk
  __host__ __device__ static constexpr size_t phys(size_t c) { return c; }
  __host__ __device__ static constexpr size_t pstride(size_t C) { return C; }
};
template <>
struct ActTraits<bf16s> {
  static constexpr bool SPLIT = true;
  static constexpr int CHUNK = 64;
  __host__ __device__ static constexpr size_t phys(size_t c) { return ((c >> 6) << 7) + (c & 63); }
  __host__ __device__ static constexpr size_t pstride(size_t C) { return 2 * C; }
};

__device__ __forceinline__ void split_bf16(float v, __nv_bfloat16& hi, __nv_bfloat16& lo) {
  hi = __float2bfloat16_rn(v);
  lo = __float2bfloat16_rn(v - __bfloat162float(hi));
}

// Element of a CNM_F16F8 act tensor (include/cnmnet_b200.h): hi = fp16(v) in the chunk's 128-byte fp16 row; the next
// 128-byte row holds lo8 = e4m3((v - hi) * 2^7) for the 64 channels, then hi8 = e4m3(hi * 2^-4).  Readers other than
// the tensor cores reconstruct v = hi + lo8 * 2^-7.  Chunk rows start on 128-byte boundaries.
struct h16f8 {
  __half v;
};
template <>
struct ActTraits<h16f8> {
  static constexpr bool SPLIT = true;
  static constexpr int CHUNK = 64;
  __host__ __device__ static constexpr size_t phys(size_t c) { return ((c >> 6) << 7) + (c & 63); }
  __host__ __device__ static constexpr size_t pstride(size_t C) { return 2 * C; }
};
constexpr float F8_LO_SCALE = 128.0f, F8_HI_SCALE = 0.0625f;

__device__ __forceinline__ float sat_f16(float v) { return fminf(fmaxf(v, -65504.0f), 65504.0f); }
__device__ __forceinline__ uint32_t e4m3x2(float a, float b) {  // a -> low byte, b -> high byte
  return (uint32_t)__nv_cvt_float2_to_fp8x2(make_float2(a, b), __NV_SATFINITE, __NV_E4M3);
}
__device__ __forceinline__ float2 e4m3x2_to_float2(uint32_t v) {
  const __half2_raw h = __nv_cvt_fp8x2_to_halfraw2((__nv_fp8x2_storage_t)v, __NV_E4M3);
  return __half22float2(*reinterpret_cast<const __half2*>(&h));
}
// the three stored parts of one value
__device__ __forceinline__ void split_f16f8(float v, __half& hi, float& lo_scaled, float& hi_scaled) {
  hi = __float2half_rn(sat_f16(v));
  const float h = __half2float(hi);
  lo_scaled = (v - h) * F8_LO_SCALE;
  hi_scaled = h * F8_HI_SCALE;
}

// Two values -> their stored parts, in as few instructions as the ISA allows: one saturating packed f32 -> f16x2
// conversion, the lo residuals in fp32, hi8 from the fp16 pair (hi * 2^-4 is exact in fp16 unless far below the e4m3
// range, where both routes give 0).  Same results as two split_f16f8 calls for every non-NaN input.
__device__ __forceinline__ void pack2_f16f8(float a, float b, uint32_t& h2, uint32_t& lo2, uint32_t& hi2) {
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(h2) : "f"(b), "f"(a));  // a -> low half
  const __half2 h = *reinterpret_cast<const __half2*>(&h2);
  const float2 hf = __half22float2(h);
  lo2 = e4m3x2((a - hf.x) * F8_LO_SCALE, (b - hf.y) * F8_LO_SCALE);
  const __half2 hs = __hmul2(h, __float2half2_rn(F8_HI_SCALE));
  hi2 = (uint32_t)__nv_cvt_halfraw2_to_fp8x2(*reinterpret_cast<const __half2_raw*>(&hs), __NV_SATFINITE, __NV_E4M3);
}

// scalar access to logical channel c of the pixel starting at `px`
template <typename T>
__device__ __forceinline__ float load_scalar(const T* px, size_t c) { return to_f<T>(px[c]); }
template <>
__device__ __forceinline__ float load_scalar<bf16s>(const bf16s* px, size_t c) {
  const size_t q = ActTraits<bf16s>::phys(c);
  return __bfloat162float(px[q].v) + __bfloat162float(px[q + 64].v);
}
template <>
__device__ __forceinline__ float load_scalar<h16f8>(const h16f8* px, size_t c) {
  const size_t q = ActTraits<h16f8>::phys(c);
  const uint8_t* row8 = reinterpret_cast<const uint8_t*>(px + (q & ~(size_t)63) + 64);
  return __half2float(px[q].v) + e4m3x2_to_float2(row8[c & 63]).x * (1.0f / F8_LO_SCALE);
}
template <typename T>
__device__ __forceinline__ void store_scalar(T* px, size_t c, float v) { px[c] = from_f<T>(v); }
template <>
__device__ __forceinline__ void store_scalar<h16f8>(h16f8* px, size_t c, float v) {
  const size_t q = ActTraits<h16f8>::phys(c);
  uint8_t* row8 = reinterpret_cast<uint8_t*>(px + (q & ~(size_t)63) + 64);
  float lo, hs;
  split_f16f8(v, px[q].v, lo, hs);
  const uint32_t pr = e4m3x2(lo, hs);
  row8[c & 63] = (uint8_t)(pr & 0xff);
  row8[64 + (c & 63)] = (uint8_t)(pr >> 8);
}
template <>
__device__ __forceinline__ void store_scalar<bf16s>(bf16s* px, size_t c, float v) {
  const size_t q = ActTraits<bf16s>::phys(c);
  split_bf16(v, px[q].v, px[q + 64].v);
}

// 16-byte vector of act elements <-> floats
template <typename T>
struct Vec16;
template <>
struct Vec16<float> {
  static constexpr int N = 4;
  __device__ static __forceinline__ void load(const float* p, float* f) {
    float4 v = *reinterpret_cast<const float4*>(p);
    f[0] = v.x; f[1] = v.y; f[2] = v.z; f[3] = v.w;
  }
  __device__ static __forceinline__ void store(float* p, const float* f) {
    *reinterpret_cast<float4*>(p) = make_float4(round_tf32(f[0]), round_tf32(f[1]), round_tf32(f[2]), round_tf32(f[3]));
  }
};
template <>
struct Vec16<__nv_bfloat16> {
  static constexpr int N = 8;
  __device__ static __forceinline__ void load(const __nv_bfloat16* p, float* f) {
    uint4 v = *reinterpret_cast<const uint4*>(p);
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      f[2 * i] = __uint_as_float(w[i] << 16);
      f[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
    }
  }
  __device__ static __forceinline__ void store(__nv_bfloat16* p, const float* f) {
    uint32_t w[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      __nv_bfloat162 h = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
      w[i] = *reinterpret_cast<uint32_t*>(&h);
    }
    *reinterpret_cast<uint4*>(p) = make_uint4(w[0], w[1], w[2], w[3]);
  }
};

// split bf16: p points at the hi vector, the lo vector is 64 physical channels further
template <>
struct Vec16<bf16s> {
  static constexpr int N = 8;
  __device__ static __forceinline__ void load(const bf16s* p, float* f) {
    float lo[8];
    Vec16<__nv_bfloat16>::load(reinterpret_cast<const __nv_bfloat16*>(p), f);
    Vec16<__nv_bfloat16>::load(reinterpret_cast<const __nv_bfloat16*>(p) + 64, lo);
#pragma unroll
    for (int i = 0; i < 8; ++i) f[i] += lo[i];
  }
  __device__ static __forceinline__ void store(bf16s* p, const float* f) {
    float hi[8], lo[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      hi[i] = __bfloat162float(__float2bfloat16_rn(f[i]));
      lo[i] = f[i] - hi[i];
    }
    Vec16<__nv_bfloat16>::store(reinterpret_cast<__nv_bfloat16*>(p), hi);
    Vec16<__nv_bfloat16>::store(reinterpret_cast<__nv_bfloat16*>(p) + 64, lo);
  }
};

// fp16 + fp8: p points at the 8-channel hi vector (16 B); the matching 8 lo8 bytes sit in the fp8 row that follows the
// chunk's fp16 row, the 8 hi8 bytes 64 further.  The channel offset inside the chunk comes from the address (chunk rows
// are 128-byte aligned).
template <>
struct Vec16<h16f8> {
  static constexpr int N = 8;
  __device__ static __forceinline__ const uint8_t* row8(const h16f8* p) {
    const uintptr_t a = reinterpret_cast<uintptr_t>(p);
    const uintptr_t j = (a & 127) >> 1;  // channel offset of the vector inside its 64-channel chunk
    return reinterpret_cast<const uint8_t*>(a - 2 * j + 128 + j);
  }
  __device__ static __forceinline__ void load(const h16f8* p, float* f) {
    const uint4 v = *reinterpret_cast<const uint4*>(p);
    const uint2 l = *reinterpret_cast<const uint2*>(row8(p));
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
    const uint32_t lw[4] = {l.x & 0xffffu, l.x >> 16, l.y & 0xffffu, l.y >> 16};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float2 h = __half22float2(*reinterpret_cast<const __half2*>(&w[i]));
      const float2 lo = e4m3x2_to_float2(lw[i]);
      f[2 * i] = h.x + lo.x * (1.0f / F8_LO_SCALE);
      f[2 * i + 1] = h.y + lo.y * (1.0f / F8_LO_SCALE);
    }
  }
  __device__ static __forceinline__ void store(h16f8* p, const float* f) {
    uint32_t w[4], lo[4], hs[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) pack2_f16f8(f[2 * i], f[2 * i + 1], w[i], lo[i], hs[i]);
    *reinterpret_cast<uint4*>(p) = make_uint4(w[0], w[1], w[2], w[3]);
    uint8_t* r8 = const_cast<uint8_t*>(row8(p));
    *reinterpret_cast<uint2*>(r8) = make_uint2(lo[0] | (lo[1] << 16), lo[2] | (lo[3] << 16));
    *reinterpret_cast<uint2*>(r8 + 64) = make_uint2(hs[0] | (hs[1] << 16), hs[2] | (hs[3] << 16));
  }
};

// ---- one K-chunk row (64 logical channels; 32 for tf32) of fp32 values -> its stored 16-byte pieces ----------------
// PIECES x 16 bytes = 128 (bf16 / tf32) or 256 (the two-row layouts: hi row, then lo / fp8 row) bytes per pixel-chunk.
template <typename T>
struct RowPack;
template <>
struct RowPack<__nv_bfloat16> {  // 8 pieces of 8 channels
  static constexpr int PIECES = 8;
  __device__ static __forceinline__ uint4 piece(const float* v, int j) {
    uint32_t w[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const __nv_bfloat162 h = __floats2bfloat162_rn(v[8 * j + 2 * e], v[8 * j + 2 * e + 1]);
      w[e] = *reinterpret_cast<const uint32_t*>(&h);
    }
    return make_uint4(w[0], w[1], w[2], w[3]);
  }
};
template <>
struct RowPack<float> {  // 8 pieces of 4 channels, tf32-rounded
  static constexpr int PIECES = 8;
  __device__ static __forceinline__ uint4 piece(const float* v, int j) {
    return make_uint4(__float_as_uint(round_tf32(v[4 * j])), __float_as_uint(round_tf32(v[4 * j + 1])),
                      __float_as_uint(round_tf32(v[4 * j + 2])), __float_as_uint(round_tf32(v[4 * j + 3])));
  }
};
template <>
struct RowPack<bf16s> {  // 8 hi pieces, then 8 lo pieces
  static constexpr int PIECES = 16;
  __device__ static __forceinline__ uint4 piece(const float* v, int j) {
    uint32_t w[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      float a = v[8 * (j & 7) + 2 * e], b = v[8 * (j & 7) + 2 * e + 1];
      if (j >= 8) {
        a -= __bfloat162float(__float2bfloat16_rn(a));
        b -= __bfloat162float(__float2bfloat16_rn(b));
      }
      const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
      w[e] = *reinterpret_cast<const uint32_t*>(&h);
    }
    return make_uint4(w[0], w[1], w[2], w[3]);
  }
};
template <>
struct RowPack<h16f8> {  // 8 fp16 pieces (8 channels), 4 lo8 pieces (16 channels), 4 hi8 pieces
  static constexpr int PIECES = 16;
  __device__ static __forceinline__ uint4 piece(const float* v, int j) {
    uint32_t w[4];
    if (j < 8) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        uint32_t lo, hs;
        pack2_f16f8(v[8 * j + 2 * e], v[8 * j + 2 * e + 1], w[e], lo, hs);
      }
    } else {
      const int c0 = 16 * (j & 3);
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        uint32_t h, lo0, hs0, lo1, hs1;
        pack2_f16f8(v[c0 + 4 * e], v[c0 + 4 * e + 1], h, lo0, hs0);
        pack2_f16f8(v[c0 + 4 * e + 2], v[c0 + 4 * e + 3], h, lo1, hs1);
        w[e] = j < 12 ? (lo0 | (lo1 << 16)) : (hs0 | (hs1 << 16));
      }
    }
    return make_uint4(w[0], w[1], w[2], w[3]);
  }
};

}  // namespace cnm
