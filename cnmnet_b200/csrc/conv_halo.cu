// K2h: stride-1 convolution + folded BatchNorm + ReLU as implicit GEMM on tcgen05 with a ROW-HALO A slab and two
// pixel tiles per weight tile.  Same math and act/weight layouts as conv_tc.cu; what changes is operand delivery,
// the measured limiter of that kernel (profiles/r01_conv_tc_ncu.md: every tap re-read its 16 KB pixel tile from L2).
//
// Reference semantics: nn.Conv2d + nn.BatchNorm2d(eval) + nn.ReLU and the torch.cat of the inputs
// (depthnet/depthNet_model.py:19-112, 235-261, 335-365).
//
// Output tile     16 x 16 pixels = two M=128 UMMA tiles (rows 0-7 and 8-15), N = BN output channels.
// A slab          for a (source, channel chunk, dx): ONE TMA box {chunk, 16, 16+2p, 1} at (c0, x0+dx, y0-p): the k
//                   vertical taps of both pixel tiles are row windows of this slab, i.e. plain K-major 128B-swizzled
//                   operands starting at slab + (dy + 8*tile)*2048 B (1024-aligned, SBO = 1024: no descriptor tricks).
//                   L2->SM bytes per tap drop from 2 x 16 KB to (16+2p)/k x 2 KB  (3x3: 12 KB, 7x7: 6.3 KB).
// B tile          per tap {chunk, BN} of the packed weights, shared by the two pixel tiles (halves weight traffic).
// Pipelines       A ring (2-4 slabs) and B ring (3-6 tiles), both TMA -> mbarrier -> tcgen05.mma -> tcgen05.commit.
// Accumulators    2 x BN TMEM columns per tile pair; double-buffered when 4*BN <= 512.
// Warp roles      warp 0 A-slab TMA producer, warp 3 B-tile TMA producer, warps 1-2 MMA issuers (one per pixel tile;
//                   warp 2 also allocates TMEM), warps 4-7 epilogue.
#include <cstring>

#include "tc_common.cuh"

namespace cnm {

constexpr int HL_MAX_SRCS = 3;
constexpr int HL_TILE_BYTES = 128 * 128;
constexpr int HL_MAX_A = 4, HL_MAX_B = 8;

struct alignas(64) HaloParams {
  CUtensorMap a_maps[HL_MAX_SRCS * 4];  // [source][parity sub-image]
  CUtensorMap b_map;
  CUtensorMap out_map;
  const cnm_conv_group* groups;
  const float* bias;
  int num_groups, num_steps;
  int tiles_x, tiles_y, tiles_b, tiles_n;
  int a_stages, b_stages, slab_bytes;
  int out_stages;  // 1 or 2 staging tiles for the TMA stores (2 when the epilogue is the long pole: BN = 64)
  int map_pad[HL_MAX_SRCS * 4];   // rows of the slab above the tile, per map
  int map_rows[HL_MAX_SRCS * 4];  // slab height (16 + vertical tap span), per map
  int relu;
  int c_out_offset;
  int mode;         // EPI_PLAIN / EPI_BF16X3 / EPI_F16F8: the split modes store two 128-byte rows per 64-column chunk
  float acc_scale;  // 2^-acc_exp
  unsigned long long* range_ctr;  // debug counter of out-of-fp16-range outputs (NULL: off)
  const float* head_w;   // fused 1-channel 3x3 head (BN == Cout == 64 only): [9][64] weights, or NULL
  float* head_partial;   // [Cout/64 slices][N, Ho, Wo, 12] per-pixel tap dot products, one slice per 64-channel chunk
  int Ho, Wo, cout;
};

template <int BN, bool TF32>
__global__ void __launch_bounds__(256, 1) conv_halo_kernel(const __grid_constant__ HaloParams p) {
  constexpr int COLS_PER_CHUNK = TF32 ? 32 : 64;
  constexpr int CHUNK_CH = TF32 ? 32 : 64;
  constexpr int NUM_CHUNKS = BN / COLS_PER_CHUNK;
  constexpr int ACC_STAGES = (4 * BN <= 512) ? 2 : 1;
  constexpr uint32_t TMEM_COLS = (2 * BN * ACC_STAGES) < 32 ? 32 : 2 * BN * ACC_STAGES;
  constexpr int B_BYTES = BN * 128;
  constexpr uint32_t IDESC_SHAPE = (1u << 4) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  constexpr uint32_t IDESC = IDESC_SHAPE | ((TF32 ? 2u : 1u) << 7) | ((TF32 ? 2u : 1u) << 10);
  // CNM_F16F8: operand formats 0 = e4m3 (kind::f8f6f4) on the first half of the groups (the correction rows, accumulated
  // while the sum is still tiny) and 0 = fp16 (kind::f16) on the second half
  const bool mixed = !TF32 && p.mode == EPI_F16F8;
  const uint32_t idesc = mixed ? IDESC_SHAPE : IDESC;

  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* base_ptr = smem_raw + (base - smem_u32(smem_raw));
  // carve-up: [A ring][B ring][staging 16 KB][bias][barriers][tmem slot]
  const uint32_t off_b = (uint32_t)p.a_stages * p.slab_bytes;
  const uint32_t off_out = off_b + (uint32_t)p.b_stages * B_BYTES;
  const uint32_t off_bias = off_out + (uint32_t)p.out_stages * HL_TILE_BYTES;
  const uint32_t off_head = off_bias + BN * 4;  // fused-head weights of the N tile [9][BN]
  const uint32_t off_bar = off_head + (p.head_partial ? 9 * BN * 4 : 0);
  const uint32_t bar0 = base + off_bar;
  auto fullA = [&](int s) { return bar0 + 8u * s; };
  auto emptyA = [&](int s) { return bar0 + 8u * (HL_MAX_A + s); };
  auto fullB = [&](int s) { return bar0 + 8u * (2 * HL_MAX_A + s); };
  auto emptyB = [&](int s) { return bar0 + 8u * (2 * HL_MAX_A + HL_MAX_B + s); };
  auto tfull = [&](int a) { return bar0 + 8u * (2 * HL_MAX_A + 2 * HL_MAX_B + a); };
  auto tempty = [&](int a) { return bar0 + 8u * (2 * HL_MAX_A + 2 * HL_MAX_B + 2 + a); };
  const uint32_t off_tmem = off_bar + 8u * (2 * HL_MAX_A + 2 * HL_MAX_B + 4);
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(base_ptr + off_tmem);
  float* bias_s = reinterpret_cast<float*>(base_ptr + off_bias);
  float* head_s = reinterpret_cast<float*>(base_ptr + off_head);
  const bool fused_head = !TF32 && p.head_partial != nullptr;

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);  // warp-uniform for the compiler
  const int lane = threadIdx.x & 31;
  const int total_tiles = p.tiles_x * p.tiles_y * p.tiles_b * p.tiles_n;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.b_map);
    tma_prefetch_desc(&p.out_map);
    tma_prefetch_desc(&p.a_maps[0]);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < p.a_stages; ++s) { mbar_init(fullA(s), 1); mbar_init(emptyA(s), 2); }
    for (int s = 0; s < p.b_stages; ++s) { mbar_init(fullB(s), 1); mbar_init(emptyB(s), 2); }
    for (int a = 0; a < 2; ++a) { mbar_init(tfull(a), 2); mbar_init(tempty(a), 128); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(base + off_tmem), "r"(TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ================= TMA producer, A slabs =================
    // (separate from the B producer so that slab requests run a_stages groups ahead regardless of the B ring depth)
    if (lane == 0) {
      int sa = 0;
      uint32_t pa = 0;
      int4 g_next = __ldg(reinterpret_cast<const int4*>(p.groups));
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        int m = tile / p.tiles_n;
        const int x0 = (m % p.tiles_x) * 16;
        m /= p.tiles_x;
        const int y0 = (m % p.tiles_y) * 16;
        const int n0 = m / p.tiles_y;
        for (int g = 0; g < p.num_groups; ++g) {
          const int4 gr = g_next;  // src, c0, dx, ntaps
          g_next = __ldg(reinterpret_cast<const int4*>(p.groups + ((g + 1 == p.num_groups) ? 0 : g + 1)));
          const int pad = p.map_pad[gr.x];
          const uint32_t slab_tx = (uint32_t)p.map_rows[gr.x] * 2048u;
          mbar_wait(emptyA(sa), pa ^ 1);
          mbar_expect_tx(fullA(sa), slab_tx);
          tma_load_4d(base + sa * p.slab_bytes, &p.a_maps[gr.x], fullA(sa), gr.y, x0 + gr.z, y0 - pad, n0);
          if (++sa == p.a_stages) { sa = 0; pa ^= 1; }
        }
      }
    }
  } else if (warp == 3) {
    // ================= TMA producer, B (weight) tiles =================
    if (lane == 0) {
      int sb = 0;
      uint32_t pb = 0;
      const int num_steps = p.num_steps;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int nt = tile % p.tiles_n;
        for (int kstep = 0; kstep < num_steps; ++kstep) {
          mbar_wait(emptyB(sb), pb ^ 1);
          mbar_expect_tx(fullB(sb), B_BYTES);
          tma_load_2d(base + off_b + sb * B_BYTES, &p.b_map, fullB(sb), kstep * CHUNK_CH, nt * BN);
          if (++sb == p.b_stages) { sb = 0; pb ^= 1; }
        }
      }
    }
  } else if (warp == 1 || warp == 2) {
    // ================= MMA issuers =================
    // Two issuing lanes, one per pixel tile of the pair (warp 1 -> rows 0-7 / accumulator columns [0,BN), warp 2 -> rows
    // 8-15 / [BN,2BN)): a single lane needs ~100 instructions per weight step (barrier wait, descriptors, 8 MMAs,
    // commit), which at BN <= 128 is as long as the MMAs themselves take (profiles/r01_conv_halo_ncu.md).  Each
    // barrier that the MMAs release therefore expects two commits.
    // The whole warp walks the loop (all lanes poll the barriers) and one elected lane issues: this keeps the control
    // flow warp-uniform, so the compiler holds descriptors in uniform registers instead of wrapping every tcgen05
    // instruction in a per-lane election loop.
    {
      const int half = warp - 1;
      int sa = 0, sb = 0;
      uint32_t pa = 0, pb = 0;
      int it = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
        const int as = (ACC_STAGES == 2) ? (it & 1) : 0;
        const uint32_t aphase = (ACC_STAGES == 2) ? ((it >> 1) & 1) : (it & 1);
        mbar_wait(tempty(as), aphase ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)(as * 2 * BN + half * BN);
        uint32_t first = 1;
        for (int g = 0; g < p.num_groups; ++g) {
          const int ntaps = __ldg(&p.groups[g].ntaps);
          mbar_wait(fullA(sa), pa);
          const uint64_t adesc0 = umma_desc(base + sa * p.slab_bytes + half * 8 * 2048);
          for (int t = 0; t < ntaps; ++t) {
            mbar_wait(fullB(sb), pb);
            tc_fence_after();
            if (elect_one()) {
              const uint64_t bdesc = umma_desc(base + off_b + sb * B_BYTES);
              const uint64_t adesc = adesc0 + (uint64_t)(t * (2048 >> 4));
              if (mixed && g < (p.num_groups >> 1)) {
#pragma unroll
                for (int k = 0; k < 4; ++k)
                  tc_mma_f8(tmem_d, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (first && k == 0) ? 0u : 1u);
              } else {
#pragma unroll
                for (int k = 0; k < 4; ++k)
                  tc_mma<TF32>(tmem_d, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (first && k == 0) ? 0u : 1u);
              }
              tc_commit(emptyB(sb));
              if (t == ntaps - 1) tc_commit(emptyA(sa));
            }
            __syncwarp();
            first = 0;
            if (++sb == p.b_stages) { sb = 0; pb ^= 1; }
          }
          if (++sa == p.a_stages) { sa = 0; pa ^= 1; }
        }
        if (elect_one()) tc_commit(tfull(as));
        __syncwarp();
      }
    }
  } else if (warp >= 4) {
    // ================= epilogue =================
    const int q = warp - 4;
    const int row = q * 32 + lane;
    const int et = threadIdx.x - 128;
    const bool leader = (et == 0);
    int it = 0;
    int bias_tile = -1;
    uint32_t store_cnt = 0;

    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
      const int nt = tile % p.tiles_n;
      int m = tile / p.tiles_n;
      const int x0 = (m % p.tiles_x) * 16;
      m /= p.tiles_x;
      const int y0 = (m % p.tiles_y) * 16;
      const int n0 = m / p.tiles_y;
      const int as = (ACC_STAGES == 2) ? (it & 1) : 0;
      const uint32_t aphase = (ACC_STAGES == 2) ? ((it >> 1) & 1) : (it & 1);
      if (bias_tile != nt) {
        epi_bar();
        for (int i = et; i < BN; i += 128) bias_s[i] = p.bias ? __ldg(p.bias + nt * BN + i) : 0.0f;
        if (fused_head)
          for (int i = et; i < 9 * BN; i += 128) head_s[i] = __ldg(p.head_w + (i / BN) * p.cout + nt * BN + i % BN);
        bias_tile = nt;
        epi_bar();  // epi_finalize reads bias_s right after the TMEM load, before any other barrier
      }
      mbar_wait(tfull(as), aphase);
      tc_fence_after();
#pragma unroll 1
      for (int half = 0; half < 2; ++half) {
        const uint32_t tmem_row = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * 2 * BN + half * BN);
#pragma unroll 1
        for (int ch = 0; ch < NUM_CHUNKS; ++ch) {
          uint32_t v[COLS_PER_CHUNK];
          tmem_ld32(tmem_row + ch * COLS_PER_CHUNK, v);
          if (!TF32) tmem_ld32(tmem_row + ch * COLS_PER_CHUNK + 32, v + 32);
          tmem_ld_wait();
          epi_finalize<COLS_PER_CHUNK>(v, bias_s + ch * COLS_PER_CHUNK, p.relu, p.acc_scale, p.range_ctr);
          if (!TF32 && fused_head) {
            const int px = x0 + (row & 15), py = y0 + 8 * half + (row >> 4);
            if (px < p.Wo && py < p.Ho) {
              const size_t slice = (size_t)(nt * NUM_CHUNKS + ch) * p.tiles_b * p.Ho * p.Wo;  // one slice per 64-channel chunk
              epi_head_partial(v, head_s + ch * 64, BN,
                               p.head_partial + (slice + ((size_t)n0 * p.Ho + py) * p.Wo + px) * HEAD_PARTIAL_STRIDE);
            }
          }
          for (int part = 0; part < (p.mode ? 2 : 1); ++part) {
            // BN >= 128: one staging tile (smem goes to deeper operand rings; the epilogue runs under the next tile's
            // long main loop).  BN = 64: the K loop is short and the epilogue is the long pole, so two staging tiles
            // let the packing of one row overlap the TMA store of the previous one.
            uint8_t* stg = base_ptr + off_out + (p.out_stages == 2 ? (store_cnt & 1u) * HL_TILE_BYTES : 0u);
            ++store_cnt;
            if (leader) {
              if (p.out_stages == 2) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
              else asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            }
            epi_bar();
            epi_store_row<TF32>(v, stg + row * 128, row, p.mode, part);
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            epi_bar();
            if (leader) {
              tma_store_4d(&p.out_map, smem_u32(stg), p.c_out_offset + (p.mode ? 2 * (nt * BN + ch * CHUNK_CH) + part * 64 : nt * BN + ch * CHUNK_CH),
                                   x0, y0 + 8 * half, n0);
              asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
          }
        }
      }
      tc_fence_before();
      mbar_arrive(tempty(as));
    }
    if (leader) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}

template <int BN, bool TF32>
static int launch_halo(const HaloParams& p, int grid, int smem, cudaStream_t s) {
  static DynSmemLimit limit;
  cudaError_t e = limit.ensure(conv_halo_kernel<BN, TF32>, smem);
  if (e != cudaSuccess) return fail(CNM_ERR_CUDA, "cnm_conv2d_halo: smem attribute (%d B): %s", smem, cudaGetErrorString(e));
  conv_halo_kernel<BN, TF32><<<grid, 256, smem, s>>>(p);
  return after_launch("cnm_conv2d_halo");
}

}  // namespace cnm

using namespace cnm;

extern "C" int cnm_conv2d_halo(const cnm_conv_desc* d, const cnm_conv_src* srcs, const int32_t* map_info,
                               const cnm_conv_group* groups, int num_groups, const void* wp, const float* bias, void* out,
                               void* stream) {
  int rc = check_conv_args("cnm_conv2d_halo", d, srcs, groups, wp, out);
  if (rc) return rc;
  CNM_REQUIRE(d->ksplit <= 1, CNM_ERR_UNSUPPORTED_SHAPE, "cnm_conv2d_halo: split-K is a cnm_conv2d_tc feature");
  CNM_REQUIRE(map_info && num_groups > 0, CNM_ERR_INVALID_ARGUMENT, "cnm_conv2d_halo: bad group list");
  CNM_REQUIRE(d->Ho >= 16 && d->Wo >= 16, CNM_ERR_UNSUPPORTED_SHAPE,
              "cnm_conv2d_halo: output %dx%d smaller than the 16x16 tile (use cnm_conv2d_tc)", d->Ho, d->Wo);
  CNM_REQUIRE(d->Cout % 64 == 0, CNM_ERR_UNSUPPORTED_SHAPE, "cnm_conv2d_halo: Cout=%d must be a multiple of 64", d->Cout);
  const int eb = elem_bytes(d->dtype), ch = chunk_channels(d->dtype);
  HaloParams p;
  memset(&p, 0, sizeof(p));
  // map_info[src*4 + sub] = {rows above the tile, slab rows}; sub = 0 for stride 1, parity py*2+px for stride 2
  int rows_max = 0;
  const int nsub = d->stride == 2 ? 4 : 1;
  for (int i = 0; i < d->num_srcs; ++i)
    for (int sub = 0; sub < nsub; ++sub) {
      const int m = i * 4 + sub, pad = map_info[2 * m], rows = map_info[2 * m + 1];
      if (rows == 0) continue;  // sub-image not used by any group
      CNM_REQUIRE(pad >= 0 && pad <= 4 && rows >= 16 && rows <= 24, CNM_ERR_UNSUPPORTED_SHAPE,
                  "cnm_conv2d_halo: map %d has pad %d / %d slab rows", m, pad, rows);
      p.map_pad[m] = pad;
      p.map_rows[m] = rows;
      if (rows > rows_max) rows_max = rows;
    }
  CNM_REQUIRE(rows_max > 0, CNM_ERR_INVALID_ARGUMENT, "cnm_conv2d_halo: no operand maps");
  const int pmax = (rows_max - 16 + 1) / 2;
  p.slab_bytes = rows_max * 2048;
  p.tiles_x = (d->Wo + 15) / 16;
  p.tiles_y = (d->Ho + 15) / 16;
  p.tiles_b = d->N;
  const int m_tiles = p.tiles_x * p.tiles_y * p.tiles_b;
  // column tile: minimise waves x relative tile time.  BN=256 holds 2 x 256 accumulator columns = all of TMEM, so its
  // epilogue is not overlapped with the next tile's main loop (~10 % on top of twice the BN=128 tile time).
  int BN = 64;
  if (d->Cout % 128 == 0) {
    BN = 128;
    if (d->Cout % 256 == 0) {
      const long w256 = ((long)m_tiles * (d->Cout / 256) + 147) / 148, w128 = ((long)m_tiles * (d->Cout / 128) + 147) / 148;
      if (w256 * 20 <= w128 * 10) BN = 256;
    }
  }
  if (d->reserved == 64 || d->reserved == 128 || d->reserved == 256) {  // caller-selected tile (host autotuner)
    CNM_REQUIRE(d->Cout % d->reserved == 0, CNM_ERR_INVALID_ARGUMENT, "cnm_conv2d_halo: BN=%d does not divide Cout=%d",
                d->reserved, d->Cout);
    BN = d->reserved;
  }
  p.tiles_n = d->Cout / BN;
  p.groups = groups;
  p.bias = bias;
  p.num_groups = num_groups;
  p.num_steps = d->num_steps;
  p.relu = d->relu;
  p.c_out_offset = d->c_out_offset;
  p.mode = d->dtype == CNM_BF16X3 ? EPI_BF16X3 : (d->dtype == CNM_F16F8 ? EPI_F16F8 : EPI_PLAIN);
  p.acc_scale = ldexpf(1.0f, -d->acc_exp);
  p.range_ctr = d->dtype == CNM_F16F8 ? range_counter() : nullptr;
  p.head_w = d->head_w;
  p.head_partial = d->head_partial;
  p.Ho = d->Ho; p.Wo = d->Wo; p.cout = d->Cout;
  // smem budget: 227 KB - alignment slack - staging - bias - barriers.  Ring depths are sized to cover the loaded
  // L2 latency (~3-4k cycles): a B tile lasts 2 x 4 MMAs (1024 / 512 / 384 cycles at BN = 256 / 128 / 64), a slab
  // lasts ntaps B tiles.
  p.out_stages = BN == 64 ? 2 : 1;
  const int fixed = 1024 + p.out_stages * HL_TILE_BYTES + BN * 4 + (d->head_partial ? 9 * BN * 4 : 0) + 8 * (2 * HL_MAX_A + 2 * HL_MAX_B + 4) + 16;
  const int budget = 232448 - fixed;
  const int want_b = BN == 256 ? 4 : (BN == 128 ? 7 : 8);
  p.a_stages = (pmax >= 3 || BN == 256) ? 2 : (BN == 64 ? 4 : 3);
  p.b_stages = (budget - p.a_stages * p.slab_bytes) / (BN * 128);
  if (p.b_stages > want_b) p.b_stages = want_b;
  if (p.b_stages > HL_MAX_B) p.b_stages = HL_MAX_B;
  while (p.b_stages < 3 && p.a_stages > 2) {
    p.a_stages -= 1;
    p.b_stages = (budget - p.a_stages * p.slab_bytes) / (BN * 128);
  }
  CNM_REQUIRE(p.a_stages >= 2 && p.b_stages >= 2, CNM_ERR_UNSUPPORTED_SHAPE,
              "cnm_conv2d_halo: slab of %d B does not fit", p.slab_bytes);
  {  // spend what is left on more slabs
    const int left = budget - p.a_stages * p.slab_bytes - p.b_stages * BN * 128;
    const int extra = left / p.slab_bytes;
    p.a_stages += extra;
    if (p.a_stages > HL_MAX_A) p.a_stages = HL_MAX_A;
  }
  const int smem = fixed + p.a_stages * p.slab_bytes + p.b_stages * BN * 128;

  for (int i = 0; i < d->num_srcs; ++i) {
    const cnm_conv_src& s = srcs[i];
    const uint64_t pix = (uint64_t)s.C * eb;
    for (int sub = 0; sub < nsub; ++sub) {
      const int m = i * 4 + sub;
      if (p.map_rows[m] == 0) continue;
      const uint32_t box[4] = {(uint32_t)ch, 16u, (uint32_t)p.map_rows[m], 1u};
      if (d->stride == 1) {
        const uint64_t dims[4] = {(uint64_t)s.C, (uint64_t)s.W, (uint64_t)s.H, (uint64_t)d->N};
        const uint64_t str[3] = {pix, pix * s.W, pix * s.W * s.H};
        rc = encode_map(&p.a_maps[m], d->dtype, 4, s.ptr, dims, str, box);
      } else {  // parity sub-image (py, px): every other row / column of the source, doubled strides
        const int py = sub >> 1, px = sub & 1;
        const int ws = (s.W - px + 1) / 2, hs = (s.H - py + 1) / 2;
        const uint64_t dims[4] = {(uint64_t)s.C, (uint64_t)ws, (uint64_t)hs, (uint64_t)d->N};
        const uint64_t str[3] = {2 * pix, 2 * pix * s.W, pix * s.W * s.H};
        const uint8_t* ptr = (const uint8_t*)s.ptr + ((uint64_t)py * s.W + px) * pix;
        rc = encode_map(&p.a_maps[m], d->dtype, 4, ptr, dims, str, box);
      }
      if (rc) return rc;
    }
  }
  {
    const uint64_t K = (uint64_t)d->num_steps * ch;
    const uint64_t dims[2] = {K, (uint64_t)d->Cout};
    const uint64_t str[1] = {K * eb};
    const uint32_t bbox[2] = {(uint32_t)ch, (uint32_t)BN};
    rc = encode_map(&p.b_map, d->dtype, 2, wp, dims, str, bbox);
    if (rc) return rc;
  }
  {
    const uint64_t pix = (uint64_t)d->c_out_total * eb;
    const uint64_t dims[4] = {(uint64_t)d->c_out_total, (uint64_t)d->Wo, (uint64_t)d->Ho, (uint64_t)d->N};
    const uint64_t str[3] = {pix, pix * d->Wo, pix * d->Wo * d->Ho};
    const uint32_t box[4] = {(uint32_t)ch, 16u, 8u, 1u};
    rc = encode_map(&p.out_map, d->dtype, 4, out, dims, str, box);
    if (rc) return rc;
  }
  const int total = m_tiles * p.tiles_n;
  const int grid = total < 148 ? total : 148;
  cudaStream_t s = (cudaStream_t)stream;
  const bool tf32 = d->dtype == CNM_TF32;  // the 2-byte modes share the kind::f16 instantiations
  if (BN == 256) return tf32 ? launch_halo<256, true>(p, grid, smem, s) : launch_halo<256, false>(p, grid, smem, s);
  if (BN == 128) return tf32 ? launch_halo<128, true>(p, grid, smem, s) : launch_halo<128, false>(p, grid, smem, s);
  return tf32 ? launch_halo<64, true>(p, grid, smem, s) : launch_halo<64, false>(p, grid, smem, s);
}
