// tcgen05 / TMA / mbarrier PTX wrappers shared by the tensor-core convolution kernels (sm_100a).
#pragma once
#include <cuda.h>

#include "common.cuh"

namespace cnm {

// ---- PTX wrappers ------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile(
        "{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
  } while (!ok);
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2,
                                            int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_store_4d(const CUtensorMap* map, uint32_t src, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
               ::"l"(map), "r"(src), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}
// one lane of a converged warp; the branch on it stays warp-uniform for the compiler's uniform datapath
__device__ __forceinline__ bool elect_one() {
  uint32_t p;
  asm volatile("{\n.reg .pred P;\nelect.sync _|P, 0xffffffff;\nselp.u32 %0, 1, 0, P;\n}" : "=r"(p));
  return p != 0;
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
template <bool TF32>
__device__ __forceinline__ void tc_mma(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  if (TF32)
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}"
                 ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
                 : "memory");
  else
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}"
                 ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
                 : "memory");
}
// kind::f8f6f4 (e4m3 x e4m3 here): K = 32 elements = 32 bytes per instruction, twice the kind::f16 rate
__device__ __forceinline__ void tc_mma_f8(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f8f6f4 [%0], %1, %2, %3, p;\n}"
               ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
               : "memory");
}
// ---- thread-block clusters ----
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// multicast TMA load: the box lands at the same shared-memory offset in every CTA of `mask`, and each of them gets the
// complete_tx on its own barrier at the same offset
__device__ __forceinline__ void tma_load_2d_mc(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, uint16_t mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4}], [%2], %5;"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "h"(mask)
      : "memory");
}
// tcgen05.commit of this CTA's MMAs, arriving on the barrier at this offset in every CTA of `mask`
__device__ __forceinline__ void tc_commit_mc(uint32_t bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(bar), "h"(mask)
               : "memory");
}

// 32 lanes x 32 consecutive fp32 columns: thread t of the warp gets row (lane base + t)
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// the reverse: thread t of the warp writes 32 consecutive fp32 columns of row (lane base + t)
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
        "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]),
        "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]),
        "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ void epi_bar() { asm volatile("bar.sync 1, 128;" ::: "memory"); }

// K-major, 128B-swizzled smem tile: rows of 128 B, 8-row groups 1024 B apart
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr) {
  return (uint64_t)((saddr & 0x3ffff) >> 4) | ((uint64_t)1 << 16) /* LBO (unused when swizzled) */ |
         ((uint64_t)(1024 >> 4) << 32) /* SBO */ | ((uint64_t)1 << 46) /* sm100 descriptor version */ |
         ((uint64_t)2 << 61) /* SWIZZLE_128B */;
}

// ---- epilogue helpers ---------------------------------------------------------------------------------------------
// epi_finalize: accumulator columns of this thread's pixel -> value = relu(acc * scale + bias), in place (fp32 bits).
// range_ctr (debug, CNM_CHECK_RANGE=1 / cnm_set_range_check): counts (pixel, chunk) pairs holding a value beyond the fp16
// range, which CNM_F16F8 storage saturates silently otherwise.
template <int COLS>
__device__ __forceinline__ void epi_finalize(uint32_t* v, const float* bs, int relu, float scale,
                                             unsigned long long* range_ctr = nullptr) {
  bool over = false;
#pragma unroll
  for (int c = 0; c < COLS; ++c) {
    float t = fmaf(__uint_as_float(v[c]), scale, bs[c]);
    t = relu ? fmaxf(t, 0.0f) : t;
    over |= !(fabsf(t) <= 65504.0f);
    v[c] = __float_as_uint(t);
  }
  if (range_ctr && over) atomicAdd(range_ctr, 1ULL);
}

// Side disparity input (cnm_conv_desc.side_disp / side_w): the nine taps of the nearest-upsampled disparity around this
// thread's pixel (zero outside the image), then value = relu(acc * scale + bias + sum_t d[t] * w[t][c]).
__device__ __forceinline__ void side_disp_taps(const float* __restrict__ disp, int n, int py, int px, int Ho, int Wo, float* d) {
  const int h = Ho >> 1, w = Wo >> 1;
#pragma unroll
  for (int r = 0; r < 3; ++r)
#pragma unroll
    for (int s = 0; s < 3; ++s) {
      const int yy = py + r - 1, xx = px + s - 1;
      d[3 * r + s] = (yy >= 0 && yy < Ho && xx >= 0 && xx < Wo) ? __ldg(disp + ((size_t)n * h + (yy >> 1)) * w + (xx >> 1)) : 0.0f;
    }
}
template <int COLS>
__device__ __forceinline__ void epi_finalize_side(uint32_t* v, const float* bs, int relu, float scale, unsigned long long* range_ctr,
                                                  const float* d, const float* w, int wstride) {
  bool over = false;
#pragma unroll
  for (int c = 0; c < COLS; c += 4) {
    float t[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) t[e] = fmaf(__uint_as_float(v[c + e]), scale, bs[c + e]);
#pragma unroll
    for (int k = 0; k < 9; ++k) {
      const float4 w4 = *reinterpret_cast<const float4*>(w + k * wstride + c);
      t[0] = fmaf(d[k], w4.x, t[0]); t[1] = fmaf(d[k], w4.y, t[1]); t[2] = fmaf(d[k], w4.z, t[2]); t[3] = fmaf(d[k], w4.w, t[3]);
    }
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float u = relu ? fmaxf(t[e], 0.0f) : t[e];
      over |= !(fabsf(u) <= 65504.0f);
      v[c + e] = __float_as_uint(u);
    }
  }
  if (range_ctr && over) atomicAdd(range_ctr, 1ULL);
}

// epi_store_row: the finalised values of one chunk -> one 128-byte row of the 128B-swizzled staging tile.
// mode 0: plain bf16 / tf32 row.  mode 1 (CNM_BF16X3): part 0 = bf16 hi row, part 1 = bf16 lo row.
// mode 2 (CNM_F16F8): part 0 = fp16 row, part 1 = fp8 row [lo8 x 64 | hi8 x 64].
enum { EPI_PLAIN = 0, EPI_BF16X3 = 1, EPI_F16F8 = 2 };

template <bool TF32>
__device__ __forceinline__ void epi_store_row(const uint32_t* v, uint8_t* rowp, int row, int mode, int part) {
  auto val = [&](int c) { return __uint_as_float(v[c]); };
  if (TF32) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float f[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) f[e] = round_tf32(val(4 * j + e));  // store tf32-rounded activations
      *reinterpret_cast<uint4*>(rowp + ((j ^ (row & 7)) << 4)) =
          make_uint4(__float_as_uint(f[0]), __float_as_uint(f[1]), __float_as_uint(f[2]), __float_as_uint(f[3]));
    }
  } else if (mode == EPI_F16F8 && part == 1) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {  // 16 channels -> 16 lo8 bytes (piece j) and 16 hi8 bytes (piece j + 4)
      uint32_t lw[4], hw[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        uint32_t h, lo0, hs0, lo1, hs1;
        pack2_f16f8(val(16 * j + 4 * e), val(16 * j + 4 * e + 1), h, lo0, hs0);
        pack2_f16f8(val(16 * j + 4 * e + 2), val(16 * j + 4 * e + 3), h, lo1, hs1);
        lw[e] = lo0 | (lo1 << 16);
        hw[e] = hs0 | (hs1 << 16);
      }
      *reinterpret_cast<uint4*>(rowp + ((j ^ (row & 7)) << 4)) = make_uint4(lw[0], lw[1], lw[2], lw[3]);
      *reinterpret_cast<uint4*>(rowp + (((j + 4) ^ (row & 7)) << 4)) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
    }
  } else {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      uint32_t w[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        float t0 = val(8 * j + 2 * e), t1 = val(8 * j + 2 * e + 1);
        if (mode == EPI_F16F8) {
          asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(w[e]) : "f"(t1), "f"(t0));
        } else {
          if (part) {  // lo row of the split-bf16 output: what the hi row rounded away
            t0 -= __bfloat162float(__float2bfloat16_rn(t0));
            t1 -= __bfloat162float(__float2bfloat16_rn(t1));
          }
          const __nv_bfloat162 h = __floats2bfloat162_rn(t0, t1);
          w[e] = *reinterpret_cast<const uint32_t*>(&h);
        }
      }
      *reinterpret_cast<uint4*>(rowp + ((j ^ (row & 7)) << 4)) = make_uint4(w[0], w[1], w[2], w[3]);
    }
  }
}

// Fused 1-channel 3x3 head (desc.head_w): the nine per-tap dot products of one 64-channel chunk of this pixel's finalised
// fp32 output vector with the head weights, written as 12 floats (9 used) to dst.  ws points at the chunk's first weight,
// `stride` floats between taps (shared memory copy of the N tile, or the global [9][Cout] array).  Every 64-channel chunk
// of the layer writes its own slice of the partial buffer; cnm_head_finish sums the slices in order, gathers the nine
// neighbours' contributions, adds the bias and applies the sigmoid.
constexpr int HEAD_PARTIAL_STRIDE = 12;
__device__ __forceinline__ void epi_head_partial(const uint32_t* v, const float* __restrict__ ws, int stride, float* dst) {
  float acc[9];
#pragma unroll
  for (int t = 0; t < 9; ++t) acc[t] = 0.0f;
#pragma unroll
  for (int c = 0; c < 64; c += 4) {
#pragma unroll
    for (int t = 0; t < 9; ++t) {
      const float4 w4 = *reinterpret_cast<const float4*>(ws + t * stride + c);
      acc[t] = fmaf(__uint_as_float(v[c]), w4.x, acc[t]);
      acc[t] = fmaf(__uint_as_float(v[c + 1]), w4.y, acc[t]);
      acc[t] = fmaf(__uint_as_float(v[c + 2]), w4.z, acc[t]);
      acc[t] = fmaf(__uint_as_float(v[c + 3]), w4.w, acc[t]);
    }
  }
  float4* o = reinterpret_cast<float4*>(dst);
  o[0] = make_float4(acc[0], acc[1], acc[2], acc[3]);
  o[1] = make_float4(acc[4], acc[5], acc[6], acc[7]);
  o[2] = make_float4(acc[8], 0.0f, 0.0f, 0.0f);
}

// Balanced (stream-K) schedules (cnm_conv_desc.balance): the workspace is [flags, one int per CTA][slots, one per CTA];
// a slot holds the raw fp32 accumulators of one contributor piece: 128 rows x 256 columns (per-tap kernel) or two
// halves x 128 rows x 128 columns (row-halo kernels).
constexpr int BALANCE_FLAG_BYTES = 1024;
constexpr int BALANCE_SLOT_BYTES = 128 * 256 * 4;
constexpr long long BALANCE_WS_BYTES = BALANCE_FLAG_BYTES + 148LL * BALANCE_SLOT_BYTES;

// ---- balanced (stream-K) schedules: contributor pieces travel through an L2-resident slot -------------------------
// Slot layout [column][128 rows] fp32: a warp's 32 rows of one column are 128 contiguous bytes (coalesced stores by the
// contributor, conflict-free shared-memory reads by the finishing CTA).
__device__ __forceinline__ int ld_acquire_gpu(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_gpu(int* p, int v) {
  asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// 1-D bulk copy global -> this CTA's shared memory, completion counted in bytes on `bar`
__device__ __forceinline__ void bulk_load(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
// Contributor piece (128 epilogue threads): accumulator columns [0, ncols) of this thread's TMEM row -> slot, then the flag.
// tmem_row = address of column 0 in this warp's lane quarter; row = 0..127.
__device__ __forceinline__ void balance_park(uint32_t tmem_row, int ncols, float* slot, int row, int* flag, bool leader) {
#pragma unroll 1
  for (int c = 0; c < ncols; c += 32) {
    uint32_t v[32];
    tmem_ld32(tmem_row + c, v);
    tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 32; ++j) __stcg(slot + (size_t)(c + j) * 128 + row, __uint_as_float(v[j]));
  }
  // generic-proxy stores -> visible to the finishing CTA's bulk copy (async proxy) once it has seen the flag
  asm volatile("fence.proxy.async;" ::: "memory");
  __threadfence();
  epi_bar();
  if (leader) st_release_gpu(flag, 1);
}
// Head piece (128 epilogue threads), after the tile's own MMAs have completed (the operand rings are idle: this is the
// CTA's last segment): for every contributor, bulk-copy its slot into the ring memory and add it to the accumulator in
// TMEM, in contributor order.  `stride` = distance between consecutive contributors' CTA indices.
__device__ __forceinline__ void balance_gather(uint32_t tmem_row, int ncols, const float* slots, int* flags, int first, int stride,
                                               int ncontrib, uint8_t* ring, uint32_t bar, uint32_t& bar_phase, int row,
                                               bool leader) {
  const size_t slot_floats = (size_t)BALANCE_SLOT_BYTES / 4;
  for (int c = 0; c < ncontrib; ++c) {
    const int cta = first + c * stride;
    if (leader) {
      for (unsigned spins = 0; ld_acquire_gpu(flags + cta) == 0; ++spins) {
        __nanosleep(100);
        if (spins > (1u << 24)) __trap();   // ~2 s: the contributor never arrived (corrupt flags / workspace shared by concurrent launches)
      }
      asm volatile("fence.proxy.async;" ::: "memory");
      const uint32_t bytes = (uint32_t)ncols * 128u * 4u;
      mbar_expect_tx(bar, bytes);
      for (uint32_t o = 0; o < bytes; o += 32768u)
        bulk_load(smem_u32(ring) + o, reinterpret_cast<const uint8_t*>(slots + (size_t)cta * slot_floats) + o,
                  bytes - o < 32768u ? bytes - o : 32768u, bar);
    }
    mbar_wait(bar, bar_phase);
    bar_phase ^= 1u;
    const float* sm = reinterpret_cast<const float*>(ring);
#pragma unroll 1
    for (int c0 = 0; c0 < ncols; c0 += 32) {
      uint32_t v[32];
      tmem_ld32(tmem_row + c0, v);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = __float_as_uint(__uint_as_float(v[j]) + sm[(c0 + j) * 128 + row]);
      tmem_st32(tmem_row + c0, v);
    }
    tmem_st_wait();
    epi_bar();   // everyone is done with the ring copy before the next contributor's lands (and before flags are reset)
  }
  if (leader)
    for (int c = 0; c < ncontrib; ++c) flags[first + c * stride] = 0;
}

// host: cuTensorMapEncodeTiled through the runtime's driver entry point (libcuda is not linked)
int encode_map(CUtensorMap* map, int dtype, int rank, const void* ptr, const uint64_t* dims, const uint64_t* strides_bytes,
               const uint32_t* box, bool swizzle128 = true);
unsigned long long* range_counter();  // device counter of the current device, or nullptr when range checking is off
int check_conv_args(const char* fn, const cnm_conv_desc* d, const cnm_conv_src* srcs, const void* steps, const void* wp,
                    const void* out);

}  // namespace cnm
