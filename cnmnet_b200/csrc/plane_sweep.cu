// K1 (second generation): the fused plane-sweep cost volume with the source footprint staged in shared memory by TMA.
//
// Reference semantics: depthnet/depthNet_model.py:185-224 (getVolume: per plane {mul, add, div, grid_sample(bilinear,
// zeros, align_corners=False), sub, abs, sum}); cost_volume.cu holds the first-generation kernel (taps through L1).
//
// Why: round 1's kernel was bound by L1 data-pipe wavefronts -- 4 taps x 16 B per output, a 32-lane x 16 B request
// spans 5-7 lines at arbitrary alignment (ncu: 6.7 wavefronts per request, 89 % of the L1 peak, 105 us at C2).  In
// shared memory a planar fp32 channel-tap is ONE wavefront at any alignment (32 lanes x 4 B on consecutive banks),
// so an output costs 12 wavefronts instead of ~27, and the XU pipe (floor / float->int conversions) is relieved by
// doing the integer split with a magic-number add.
//
// Structure
//   CTA            32 x 8 pixels of one (view, batch) pair, ALL D planes (per-pixel invariants a = Hm.[x,y,1] and the
//                  reference pixel are computed once): 8 consumer warps, one pixel per thread, + 1 producer warp that
//                  only issues TMA.  No CTA-wide barrier in the plane loop: consumers wait on a stage's "full"
//                  mbarrier (TMA transaction count) and release it on its "empty" mbarrier, one arrival per warp.
//   plane group    8 consecutive planes share one source BOX: for a fixed pixel the sample point moves along the
//                  epipolar line as the plane changes, so the union of the footprints of the tile over 8 planes is a
//                  slightly larger rectangle than the tile itself.  Warp 0 finds it at kernel start: the homography is
//                  projective, the tile convex and 1/z monotonic, so the bounding box of the tile's 4 corners at the
//                  group's first and last plane bounds every sample (if p_z keeps its sign on those 8 corners).
//   TMA            the box is fetched as three 2-D tiles {64, BH} (one per colour channel, BH in {16, 24, 40}: the
//                  smallest tensor map that fits) straight from the caller's planar NCHW source -- no repack pass;
//                  out-of-image texels are ZERO-FILLED by TMA, which is grid_sample's zeros padding.  Two stages:
//                  the box of group g+1 lands while group g is being sampled.
//   sampling       per plane: 3 FMA (p = a z + kt), rcp + Newton, clamp into the box, integer / fraction split by
//                  adding 1.5 * 2^23, 12 x LDS.32 at immediate offsets from one address, 14 FMA/ADD.
//   levels         a group whose 8-plane footprint does not fit the largest box (fast-moving footprints: wide baselines)
//                  is cut into 2 x 4 or 4 x 2 planes per box; one that does not fit even then (extreme zoom, p_z
//                  changing sign) is sampled straight from global memory with per-tap predicates -- same arithmetic,
//                  CTA-uniform branch.
//
// Data layout   ref, src [.,3,H,W] fp32 planar (public NCHW);  z_d in kernel parameters (constant bank), evaluated in
//               fp64 on the host like the reference's Python double;  out as in cost_volume.cu.
#include <cstring>

#include "tc_common.cuh"

namespace cnm {

constexpr int PS_TX = 32, PS_TY = 8, PS_DP = 8;
constexpr int PS_BW = 64, PS_BH_MAX = 40;
constexpr int PS_NUM_BH = 3;
__host__ __device__ constexpr int ps_bh(int idx) { return idx == 0 ? 16 : (idx == 1 ? 24 : PS_BH_MAX); }
constexpr int PS_CH_WORDS = PS_BW * PS_BH_MAX;       // one colour channel of a stage
constexpr int PS_STAGE_BYTES = 3 * PS_CH_WORDS * 4;  // 30 720
constexpr int PS_MAX_D = 128;
constexpr int PS_MAX_GROUPS = PS_MAX_D / PS_DP;
constexpr int PS_STAGES = 2;
constexpr int PS_CONSUMER_WARPS = PS_TY, PS_THREADS = PS_TX * (PS_TY + 1);  // + the producer warp
constexpr int PS_OFF_GSTATE = PS_STAGES * PS_STAGE_BYTES + 64;
constexpr int PS_OFF_GBASE = PS_OFF_GSTATE + PS_MAX_GROUPS * 16;
constexpr int PS_OFF_NODES = PS_OFF_GBASE + 80;
constexpr int PS_OFF_FLAT = PS_OFF_NODES + 7 * PS_MAX_GROUPS * 16;
constexpr int PS_SMEM = PS_OFF_FLAT + 4 * PS_MAX_GROUPS * 16 + 1024;
static_assert(2 * PS_STAGES * 8 <= 64, "barrier block");
constexpr float PS_MAGIC = 12582912.0f;  // 1.5 * 2^23: (v + MAGIC) has integer spacing, low mantissa bits = round(v)
constexpr uint32_t PS_MAGIC_BITS = 0x4B400000u;

struct alignas(64) SweepParams {
  CUtensorMap maps[PS_NUM_BH];  // [W, H, 3P] fp32, box {64, BH, 1}
  const float* ref;
  const float* src;
  const float* hk;
  void* out;
  int B, H, W, D;
  float z[PS_MAX_D];
};

__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}

// mbarrier wait that does not burn issue slots: try_wait with a suspend-time hint parks the warp in hardware, the
// nanosleep backs the retry off (the profile of the plain spin loop showed 10 % of all issued instructions in it)
__device__ __forceinline__ void mbar_wait_parked(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  for (;;) {
    asm volatile(
        "{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\nselp.u32 %0, 1, 0, p;\n}"
        : "=r"(ok)
        : "r"(bar), "r"(parity), "r"(2000u)
        : "memory");
    if (ok) return;
    __nanosleep(64);
  }
}

// one fp32 from shared memory at a byte address + immediate (all 12 taps of a sample hang off one address register)
template <int IMM>
__device__ __forceinline__ float lds(uint32_t addr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1+%2];" : "=f"(v) : "r"(addr), "n"(IMM));
  return v;
}

// sample position of one plane: ix = u - 0.5 with u = (a0 z + k0) / (a2 z + k2 + 1e-6)   (depthNet_model.py:210-213 and
// grid_sample's align_corners=False un-normalisation)
__device__ __forceinline__ void sweep_position(float a0, float a1, float a2, float k0, float k1, float k2e, float z,
                                               float& ix, float& iy) {
  const float px = fmaf(a0, z, k0), py = fmaf(a1, z, k1);
  const float den = fmaf(a2, z, k2e);
  float rcp;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rcp) : "f"(den));
  rcp = fmaf(fmaf(-den, rcp, 1.0f), rcp, rcp);  // one Newton step: ~0.5 ulp
  ix = fmaf(px, rcp, -0.5f);
  iy = fmaf(py, rcp, -0.5f);
}

// one output from global memory with per-tap predicates (groups whose footprint does not fit a box)
__device__ __forceinline__ float sweep_gather_global(const float* __restrict__ sp, int H, int W, float ix, float iy,
                                                     float r0, float r1, float r2) {
  ix = fminf(fmaxf(ix, -1.0f), (float)W);  // NaN -> -1: all four taps outside
  iy = fminf(fmaxf(iy, -1.0f), (float)H);
  const float x0f = floorf(ix), y0f = floorf(iy);
  const float wx1 = ix - x0f, wy1 = iy - y0f;
  const int x0 = (int)x0f, y0 = (int)y0f;
  const bool vx0 = x0 >= 0 && x0 < W, vx1 = x0 + 1 >= 0 && x0 + 1 < W;
  const bool vy0 = y0 >= 0 && y0 < H, vy1 = y0 + 1 >= 0 && y0 + 1 < H;
  const float w11 = wx1 * wy1, w10 = wx1 - w11, w01 = wy1 - w11, w00 = (1.0f - wx1) - w01;
  const size_t HW = (size_t)H * W;
  const float* t = sp + (ptrdiff_t)y0 * W + x0;
  const float r[3] = {r0, r1, r2};
  float acc = 0.0f;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const float* tc = t + c * HW;
    const float t00 = (vx0 && vy0) ? __ldg(tc) : 0.0f, t10 = (vx1 && vy0) ? __ldg(tc + 1) : 0.0f;
    const float t01 = (vx0 && vy1) ? __ldg(tc + W) : 0.0f, t11 = (vx1 && vy1) ? __ldg(tc + W + 1) : 0.0f;
    acc += fabsf(fmaf(t11, w11, fmaf(t01, w01, fmaf(t10, w10, fmaf(t00, w00, -r[c])))));
  }
  return acc;
}

// one output from the staged box.  `cst` = shared-memory byte address of texel (0, 0) of the box minus the magic-number
// bits (see below); clamp bounds keep every tap inside the box whatever the arithmetic does.
__device__ __forceinline__ float sweep_sample_box(float ix, float iy, float cx_lo, float cx_hi, float cy_lo, float cy_hi,
                                                  uint32_t cst, float r0, float r1, float r2) {
  ix = fminf(fmaxf(ix, cx_lo), cx_hi);  // NaN -> cx_lo
  iy = fminf(fmaxf(iy, cy_lo), cy_hi);
  // integer / fraction split without the conversion pipe: tx = round(ix - 0.5) + MAGIC exactly; round(ix - 0.5) is
  // floor(ix) except within rounding of an integer, where the bilinear blend is continuous (weights -eps / 1 + eps)
  const float tx = (ix - 0.5f) + PS_MAGIC, ty = (iy - 0.5f) + PS_MAGIC;
  const float wx1 = ix - (tx - PS_MAGIC), wy1 = iy - (ty - PS_MAGIC);
  const uint32_t t = __float_as_uint(ty) * (4u * PS_BW) + (__float_as_uint(tx) << 2) + cst;
  const float w11 = wx1 * wy1, w10 = wx1 - w11, w01 = wy1 - w11, w00 = (1.0f - wx1) - w01;
  constexpr int ROW = 4 * PS_BW, CH = 4 * PS_CH_WORDS;
  const float c0 = fmaf(lds<ROW + 4>(t), w11, fmaf(lds<ROW>(t), w01, fmaf(lds<4>(t), w10, fmaf(lds<0>(t), w00, -r0))));
  const float c1 =
      fmaf(lds<CH + ROW + 4>(t), w11, fmaf(lds<CH + ROW>(t), w01, fmaf(lds<CH + 4>(t), w10, fmaf(lds<CH>(t), w00, -r1))));
  const float c2 = fmaf(lds<2 * CH + ROW + 4>(t), w11,
                        fmaf(lds<2 * CH + ROW>(t), w01, fmaf(lds<2 * CH + 4>(t), w10, fmaf(lds<2 * CH>(t), w00, -r2))));
  return fabsf(c0) + fabsf(c1) + fabsf(c2);
}

struct SweepPixel {  // per-thread invariants of the plane loop
  float a0, a1, a2, k0, k1, k2e, r0, r1, r2;
  int x, y, pair;
  bool active;
};

template <int LAYOUT>  // 0 = NCHW fp32, 1 = act bf16, 2 = act fp32 (tf32), 3 = act split bf16, 4 = act fp16 + fp8
__device__ __forceinline__ void sweep_store8(const SweepParams& p, const SweepPixel& px, int d0, const float* res) {
  const size_t HW = (size_t)p.H * p.W;
  const uint32_t pix = (uint32_t)(px.y * p.W + px.x);   // inside one image plane: 32 bits
  const int D = p.D;
  if (LAYOUT == 0) {
    // one byte pointer, advanced by the (uniform) plane stride: two integer instructions per store
    char* o = reinterpret_cast<char*>(p.out) + (((size_t)px.pair * D + d0) * HW + pix) * sizeof(float);
    const size_t plane_bytes = HW * sizeof(float);
    asm volatile("" : "+l"(o));  // one opaque per-thread pointer: keeps the compiler from re-deriving it per plane
#pragma unroll
    for (int i = 0; i < PS_DP; ++i) {
      __stcs(reinterpret_cast<float*>(o), res[i]);
      o += plane_bytes;
    }
  } else {
    const size_t pixel = (size_t)px.pair * HW + pix;
    if (LAYOUT == 1) {
      Vec16<__nv_bfloat16>::store(reinterpret_cast<__nv_bfloat16*>(p.out) + pixel * D + d0, res);
    } else if (LAYOUT == 2) {
      float* o = reinterpret_cast<float*>(p.out) + pixel * D + d0;
      Vec16<float>::store(o, res);
      Vec16<float>::store(o + 4, res + 4);
    } else if (LAYOUT == 3) {
      Vec16<bf16s>::store(reinterpret_cast<bf16s*>(p.out) + pixel * ActTraits<bf16s>::pstride(D) + ActTraits<bf16s>::phys(d0), res);
    } else {
      Vec16<h16f8>::store(reinterpret_cast<h16f8*>(p.out) + pixel * ActTraits<h16f8>::pstride(D) + ActTraits<h16f8>::phys(d0), res);
    }
  }
}

struct SweepCtx {  // CTA-uniform state of the plane loop
  const int4* flat;   // segments in plane order: box origin x, y, BH index (-1: no box), first plane
  int nflat;
  uint32_t base, bar0;  // stage s: "full" barrier at bar0 + 8 s, "empty" barrier at bar0 + 8 (PS_STAGES + s)
};

// One group of 8 planes as 8 / S segments of S planes per box, starting at flat segment n0 (consumer warps).
template <int S>
__device__ __forceinline__ void sweep_group(const SweepParams& p, const SweepPixel& px, const SweepCtx& c, int n0, int lane,
                                            float* res) {
  const int H = p.H, W = p.W;
  const float* sp = p.src + (size_t)px.pair * 3 * H * W;
#pragma unroll
  for (int sub = 0; sub < PS_DP / S; ++sub) {
    const int n = n0 + sub;
    const int st = n % PS_STAGES;
    const int4 bi = c.flat[n];
    mbar_wait_parked(c.bar0 + 8u * st, (uint32_t)(n / PS_STAGES) & 1u);  // the box has landed (or the segment has none)
    // a group whose footprint does not fit a box even at 2 planes per box (extreme zoom, p_z changing sign) is ONE
    // 8-plane segment without a box, sampled from global memory: only the S = 8 instantiation carries that path
    if (S < 8 || bi.z >= 0) {
      // x0 in [floor(ix) - 1, floor(ix)], taps x0 and x0 + 1: ix in [ox + 1, ox + BW - 2) keeps them inside the box
      const float cx_lo = fmaxf(-1.0f, (float)(bi.x + 1)), cx_hi = fminf((float)W, (float)(bi.x + PS_BW - 3));
      const float cy_lo = fmaxf(-1.0f, (float)(bi.y + 1)), cy_hi = fminf((float)H, (float)(bi.y + ps_bh(bi.z) - 3));
      // shared-memory byte address of tap (x0, y0) = stage + 4 * ((y0 - oy) * 64 + (x0 - ox)), from the float bits of
      // ty = y0 + MAGIC and tx = x0 + MAGIC
      const uint32_t cst = c.base + (uint32_t)(st * PS_STAGE_BYTES) -
                           4u * ((PS_MAGIC_BITS + (uint32_t)bi.y) * PS_BW + (PS_MAGIC_BITS + (uint32_t)bi.x));
#pragma unroll
      for (int i = 0; i < S; ++i) {
        float ix, iy;
        sweep_position(px.a0, px.a1, px.a2, px.k0, px.k1, px.k2e, p.z[bi.w + i], ix, iy);
        res[sub * S + i] = sweep_sample_box(ix, iy, cx_lo, cx_hi, cy_lo, cy_hi, cst, px.r0, px.r1, px.r2);
      }
    } else {
#pragma unroll
      for (int i = 0; i < S; ++i) {
        float ix, iy;
        sweep_position(px.a0, px.a1, px.a2, px.k0, px.k1, px.k2e, p.z[bi.w + i], ix, iy);
        res[sub * S + i] = px.active ? sweep_gather_global(sp, H, W, ix, iy, px.r0, px.r1, px.r2) : 0.0f;
      }
    }
    __syncwarp();  // every lane's loads of this stage have returned (their values feed the arithmetic above)
    if (lane == 0) mbar_arrive(c.bar0 + 8u * (PS_STAGES + st));  // this warp is done with stage st
  }
}

// The producer warp's lane 0: fills stage n % PS_STAGES with segment n's box as soon as all consumer warps have
// released what it held before.  Segments without a box complete the "full" barrier with a plain arrival, so that the
// consumers' phase arithmetic does not depend on which segments have one.
__device__ __forceinline__ void sweep_produce(const SweepParams& p, const SweepCtx& c, int pair) {
  for (int n = 0; n < c.nflat; ++n) {
    const int st = n % PS_STAGES;
    if (n >= PS_STAGES) mbar_wait_parked(c.bar0 + 8u * (PS_STAGES + st), (uint32_t)(n / PS_STAGES - 1) & 1u);
    const int4 bi = c.flat[n];
    const uint32_t full = c.bar0 + 8u * st;
    if (bi.z < 0) {
      mbar_arrive(full);
      continue;
    }
    mbar_expect_tx(full, 3u * (uint32_t)(PS_BW * ps_bh(bi.z) * 4));
#pragma unroll
    for (int ch = 0; ch < 3; ++ch)
      tma_load_3d(c.base + st * PS_STAGE_BYTES + ch * PS_CH_WORDS * 4, &p.maps[bi.z], full, bi.x, bi.y, 3 * pair + ch);
  }
}

template <int LAYOUT>
__global__ void __launch_bounds__(PS_THREADS, 3) plane_sweep_kernel(const __grid_constant__ SweepParams p) {
  pdl_trigger();
  pdl_wait();
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* base_ptr = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t bar0 = base + PS_STAGES * PS_STAGE_BYTES;
  // after the stages: barriers [64 B], per-group state [16 groups x 16 B: fit[3], level], group -> first flat segment
  // [17 ints], node table [15 G x 16 B], flat segment list [8 G x 16 B]
  int4* gstate = reinterpret_cast<int4*>(base_ptr + PS_OFF_GSTATE);
  int* gbase = reinterpret_cast<int*>(base_ptr + PS_OFF_GBASE);
  int4* seg = reinterpret_cast<int4*>(base_ptr + PS_OFF_NODES);
  int4* flat = reinterpret_cast<int4*>(base_ptr + PS_OFF_FLAT);
  const int tid = threadIdx.y * PS_TX + threadIdx.x;   // consumers: 0..255 (one pixel each); producer warp: 256..287
  const bool consumer = threadIdx.y < PS_TY;
  const int H = p.H, W = p.W;
  const int G = p.D / PS_DP;
  const size_t HW = (size_t)H * W;
  SweepPixel px;
  px.pair = blockIdx.z;
  px.x = blockIdx.x * PS_TX + threadIdx.x;
  px.y = blockIdx.y * PS_TY + threadIdx.y;
  px.active = consumer && px.x < W && px.y < H;

  const float* h = p.hk + (size_t)px.pair * 12;
  const float h0 = __ldg(h + 0), h1 = __ldg(h + 1), h2 = __ldg(h + 2), h3 = __ldg(h + 3), h4 = __ldg(h + 4),
              h5 = __ldg(h + 5), h6 = __ldg(h + 6), h7 = __ldg(h + 7), h8 = __ldg(h + 8);
  px.k0 = __ldg(h + 9); px.k1 = __ldg(h + 10); px.k2e = __ldg(h + 11) + 1e-6f;

  if (tid == 0) {
    for (int st = 0; st < PS_STAGES; ++st) {
      mbar_init(bar0 + 8u * st, 1);                                   // full: the producer's arrival (+ TMA bytes)
      mbar_init(bar0 + 8u * (PS_STAGES + st), PS_CONSUMER_WARPS);     // empty: one arrival per consumer warp
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    for (int i = 0; i < PS_NUM_BH; ++i) tma_prefetch_desc(&p.maps[i]);
  }
  if (tid < PS_MAX_GROUPS) gstate[tid] = make_int4(1, 1, 1, 0);  // fit at levels 0, 1, 2 until a node says otherwise
  if (tid == 0) gbase[PS_MAX_GROUPS] = 0;                        // "some group needs a finer level"
  __syncthreads();

  // ---- source boxes.  Level L (0, 1, 2) cuts the D planes into segments of S = 8 >> L planes; segment nodes are
  // numbered level by level (G, 2G, 4G of them).  One task = (node, tile corner): the corner's sample position at the
  // segment's first and last plane; four neighbouring lanes reduce to the bounding box.  Level 0 first (two warps); the
  // finer levels are only evaluated for tiles with a group whose 8-plane footprint does not fit (large baselines).
  auto box_tasks = [&](int node_begin, int node_end) {
    const int x_lo = blockIdx.x * PS_TX, y_lo = blockIdx.y * PS_TY;
    const int corner = tid & 3;
    const float cx = (float)((corner & 1) ? min(x_lo + PS_TX - 1, W - 1) : x_lo);
    const float cy = (float)((corner & 2) ? min(y_lo + PS_TY - 1, H - 1) : y_lo);
    const float a0 = fmaf(h0, cx, fmaf(h1, cy, h2)), a1 = fmaf(h3, cx, fmaf(h4, cy, h5)), a2 = fmaf(h6, cx, fmaf(h7, cy, h8));
    for (int n0 = node_begin; n0 < node_end; n0 += (PS_TX * PS_TY) / 4) {  // warp-uniform trip count (consumer threads)
      const int node = n0 + (tid >> 2);
      const bool live = node < node_end;
      const int nn = live ? node : node_end - 1;
      const int level = (nn >= G) + (nn >= 3 * G);
      const int j = nn - G * ((1 << level) - 1), S = PS_DP >> level;
      const float z0 = p.z[j * S], z1 = p.z[j * S + S - 1];
      const float d0 = fmaf(a2, z0, px.k2e), d1 = fmaf(a2, z1, px.k2e);
      float ix0, iy0, ix1, iy1;
      sweep_position(a0, a1, a2, px.k0, px.k1, px.k2e, z0, ix0, iy0);
      sweep_position(a0, a1, a2, px.k0, px.k1, px.k2e, z1, ix1, iy1);
      // p_z of one sign and away from zero at every corner (it is bilinear in (pixel, z): its extremes are here), so the
      // projective map is continuous on the tile; positions finite and small enough for the conversions below
      bool ok = (d0 > 1e-4f && d1 > 1e-4f) || (d0 < -1e-4f && d1 < -1e-4f);
      ok = ok && fabsf(ix0) < 1e6f && fabsf(iy0) < 1e6f && fabsf(ix1) < 1e6f && fabsf(iy1) < 1e6f;
      const float xm = (float)W, ym = (float)H;
      ix0 = fminf(fmaxf(ix0, -1.0f), xm); ix1 = fminf(fmaxf(ix1, -1.0f), xm);
      iy0 = fminf(fmaxf(iy0, -1.0f), ym); iy1 = fminf(fmaxf(iy1, -1.0f), ym);
      float mnx = fminf(ix0, ix1), mxx = fmaxf(ix0, ix1), mny = fminf(iy0, iy1), mxy = fmaxf(iy0, iy1);
      int oki = ok ? 1 : 0;
#pragma unroll
      for (int o = 1; o <= 2; o <<= 1) {
        mnx = fminf(mnx, __shfl_xor_sync(0xffffffffu, mnx, o));
        mxx = fmaxf(mxx, __shfl_xor_sync(0xffffffffu, mxx, o));
        mny = fminf(mny, __shfl_xor_sync(0xffffffffu, mny, o));
        mxy = fmaxf(mxy, __shfl_xor_sync(0xffffffffu, mxy, o));
        oki &= __shfl_xor_sync(0xffffffffu, oki, o);
      }
      if (corner == 0 && live) {
        // the taps of a sample at ix touch columns floor(ix) - 1 .. floor(ix) + 1; one more column on each side absorbs
        // the fp32 rounding of the corner positions
        // ... and the box starts on a multiple of 4 texels: TMA traps (illegal instruction) when the innermost start
        // coordinate is not 16-byte aligned
        const int ox = ((int)floorf(mnx) - 2) & ~3, oy = (int)floorf(mny) - 2;
        const int wneed = (int)floorf(mxx) + 4 - ox, hneed = (int)floorf(mxy) - (int)floorf(mny) + 6;
        int idx = -1;
        if (oki && wneed <= PS_BW) idx = hneed <= ps_bh(0) ? 0 : (hneed <= ps_bh(1) ? 1 : (hneed <= ps_bh(2) ? 2 : -1));
        seg[node] = make_int4(ox, oy, idx, j * S);
        if (idx < 0) {
          reinterpret_cast<int*>(&gstate[j >> level])[level] = 0;  // group of 8 planes this segment belongs to
          if (level == 0) gbase[PS_MAX_GROUPS] = 1;
        }
      }
    }
  };
  if (tid < 64) box_tasks(0, G);  // G <= 16 nodes x 4 corners
  // per-pixel invariants meanwhile
  const float fx = (float)px.x, fy = (float)px.y;
  px.a0 = fmaf(h0, fx, fmaf(h1, fy, h2));  // a = Hm . [x, y, 1]^T   (KRKiUV)
  px.a1 = fmaf(h3, fx, fmaf(h4, fy, h5));
  px.a2 = fmaf(h6, fx, fmaf(h7, fy, h8));
  px.r0 = px.r1 = px.r2 = 0.0f;
  if (px.active) {
    const float* rp = p.ref + (size_t)(px.pair % p.B) * 3 * HW + (size_t)px.y * W + px.x;
    px.r0 = __ldg(rp); px.r1 = __ldg(rp + HW); px.r2 = __ldg(rp + 2 * HW);
  }
  __syncthreads();
  if (gbase[PS_MAX_GROUPS]) {  // CTA-uniform
    if (consumer) box_tasks(G, 7 * G);
    __syncthreads();
  }
  // flat segment list: group g runs at the coarsest level (8, 4 or 2 planes per box) whose segments all have a box; a
  // group that has none even at 2 planes per box becomes one box-less segment of 8 planes (global-memory path)
  if (tid < G) {
    int first = 0;
    for (int hgrp = 0; hgrp <= tid; ++hgrp) {
      const int4 gs = gstate[hgrp];
      const int lv = gs.x ? 0 : (gs.y ? 1 : (gs.z ? 2 : 3));
      const int cnt = lv == 3 ? 1 : (1 << lv);
      if (hgrp == tid) {
        gstate[tid].w = lv;
        gbase[tid] = first;
        if (lv == 3) {
          flat[first] = make_int4(0, 0, -1, tid * PS_DP);
        } else {
          const int4* src_nodes = seg + G * ((1 << lv) - 1) + (tid << lv);
          for (int k = 0; k < cnt; ++k) flat[first + k] = src_nodes[k];
        }
      }
      first += cnt;
    }
    if (tid == G - 1) gbase[PS_MAX_GROUPS] = first;  // number of flat segments
  }
  __syncthreads();

  SweepCtx c;
  c.flat = flat;
  c.nflat = gbase[PS_MAX_GROUPS];
  c.base = base;
  c.bar0 = bar0;
  if (!consumer) {
    if (threadIdx.x == 0) sweep_produce(p, c, px.pair);
    return;
  }
  const int lane = threadIdx.x;
  for (int g = 0; g < G; ++g) {
    float res[PS_DP];
    const int lv = gstate[g].w, n0 = gbase[g];
    if (lv == 0 || lv == 3) sweep_group<8>(p, px, c, n0, lane, res);
    else if (lv == 1) sweep_group<4>(p, px, c, n0, lane, res);
    else sweep_group<2>(p, px, c, n0, lane, res);
    if (px.active) sweep_store8<LAYOUT>(p, px, g * PS_DP, res);
  }
}

template <int LAYOUT>
static int launch_sweep(const SweepParams& p, dim3 grid, cudaStream_t s) {
  static DynSmemLimit limit;
  cudaError_t e = limit.ensure(plane_sweep_kernel<LAYOUT>, PS_SMEM);
  if (e != cudaSuccess) return fail(CNM_ERR_CUDA, "cnm_plane_sweep: smem attribute: %s", cudaGetErrorString(e));
  launch_pdl(plane_sweep_kernel<LAYOUT>, dim3(grid), dim3(dim3(PS_TX, PS_TY + 1)), PS_SMEM, s, p);
  return after_launch("cnm_plane_sweep");
}

}  // namespace cnm

using namespace cnm;

extern "C" int cnm_plane_sweep(const float* ref, const float* src, const float* hk, void* out, int B, int V, int H, int W,
                               int D, double idepth_min, double idepth_max, int layout, int dtype, void* stream) {
  CNM_REQUIRE(ref && src && hk && out, CNM_ERR_INVALID_ARGUMENT, "cnm_plane_sweep: null pointer");
  CNM_REQUIRE(B > 0 && V > 0 && H > 0 && W > 0, CNM_ERR_INVALID_ARGUMENT, "cnm_plane_sweep: bad shape");
  CNM_REQUIRE(D >= PS_DP && D % PS_DP == 0 && D <= PS_MAX_D, CNM_ERR_UNSUPPORTED_SHAPE,
              "cnm_plane_sweep: D=%d must be a multiple of %d, at most %d (use cnm_cost_volume)", D, PS_DP, PS_MAX_D);
  CNM_REQUIRE(W % 4 == 0, CNM_ERR_UNSUPPORTED_SHAPE,
              "cnm_plane_sweep: W=%d must be a multiple of 4 (16-byte TMA row pitch; use cnm_cost_volume)", W);
  CNM_REQUIRE(layout == CNM_VOLUME_NCHW_F32 || layout == CNM_VOLUME_ACT, CNM_ERR_INVALID_ARGUMENT,
              "cnm_plane_sweep: layout %d", layout);
  CNM_REQUIRE(((uintptr_t)src & 15) == 0 && ((uintptr_t)out & 15) == 0, CNM_ERR_INVALID_ARGUMENT,
              "cnm_plane_sweep: src and out must be 16B aligned");
  const int P = B * V;
  CNM_REQUIRE(P <= 65535 && (H + PS_TY - 1) / PS_TY <= 65535, CNM_ERR_UNSUPPORTED_SHAPE, "cnm_plane_sweep: grid too large");
  SweepParams p;
  memset(&p, 0, sizeof(p));
  p.ref = ref; p.src = src; p.hk = hk; p.out = out;
  p.B = B; p.H = H; p.W = W; p.D = D;
  const double step = D > 1 ? (idepth_max - idepth_min) / (double)(D - 1) : 0.0;
  for (int d = 0; d < D; ++d) p.z[d] = (float)(1.0 / (idepth_min + (double)d * step));  // fp64 like the Python double
  const uint64_t dims[3] = {(uint64_t)W, (uint64_t)H, (uint64_t)3 * P};
  const uint64_t str[2] = {(uint64_t)W * 4, (uint64_t)W * H * 4};
  for (int i = 0; i < PS_NUM_BH; ++i) {
    const uint32_t box[3] = {(uint32_t)PS_BW, (uint32_t)ps_bh(i), 1u};
    int rc = encode_map(&p.maps[i], CNM_TF32, 3, src, dims, str, box, false);
    if (rc) return rc;
  }
  dim3 grid((W + PS_TX - 1) / PS_TX, (H + PS_TY - 1) / PS_TY, P);
  cudaStream_t s = (cudaStream_t)stream;
  if (layout == CNM_VOLUME_NCHW_F32) return launch_sweep<0>(p, grid, s);
  if (dtype == CNM_BF16) return launch_sweep<1>(p, grid, s);
  if (dtype == CNM_TF32) return launch_sweep<2>(p, grid, s);
  if (dtype == CNM_BF16X3) return launch_sweep<3>(p, grid, s);
  if (dtype == CNM_F16F8) return launch_sweep<4>(p, grid, s);
  return fail(CNM_ERR_INVALID_ARGUMENT, "cnm_plane_sweep: dtype %d", dtype);
}
