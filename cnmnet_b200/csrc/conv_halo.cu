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
//
// CTA-pair variant (conv_halo2_kernel, CG = 2): tcgen05.mma.cta_group::2 with M = 256 across the two SMs of a cluster.
// With both operands in shared memory a single-SM UMMA at N <= 128 is bound by the 128 B/clk shared-memory port
// (profiles/r01_conv_halo_ncu.md: 4 KB of A + 4 KB of B per 64-cycle MMA plus the TMA fill).  In cta_group::2 each SM
// supplies the 128 A rows of its own pixel block but only HALF of the weight tile (BN/2 rows), so B reads and B fill per
// SM halve.  Differences to the single-CTA instantiation (all under `if constexpr (CG == 2)` in ONE kernel body):
//   * cluster of 2 CTAs, both walk the same (N-tile, group, tap) sequence on two neighbouring 16x16 pixel blocks;
//   * every CTA TMA-loads its own A slabs and its half of each B tile; the loads signal the LEADER CTA's "full"
//     barriers (cp.async.bulk.tensor ... .cta_group::2 with the peer bit of the barrier address cleared);
//   * only the leader's warp 1 issues MMAs (D = 128 rows x BN columns in EACH CTA's TMEM); tcgen05.commit is
//     multicast to both CTAs' "empty" and "accumulator full" barriers;
//   * both CTAs run the epilogue on their own TMEM half; "accumulator empty" arrivals go to the leader's barrier.
#include <climits>
#include <cstring>

#include "sweep_l1.cuh"
#include "tc_common.cuh"

namespace cnm {

constexpr int HL_MAX_SRCS = 3;
constexpr int HL_TILE_BYTES = 128 * 128;
constexpr int HL_MAX_A = 4, HL_MAX_B = 8;
constexpr int HL_NUM_BARS = 2 * HL_MAX_A + 2 * HL_MAX_B + 4 + 2;  // A / B rings, accumulator full / empty, volume slots
constexpr int HL_SWEEP_MAX_D = 128;
constexpr int HL_SWEEP_THREADS = 256;  // volume-producer warps of the fused conv1.0 kernel (warps 8-15)
// registers per thread of the three roles once the fused kernel has re-balanced its 64 K (x 128 threads each: control
// warps 0-3, epilogue warps 4-7, two producer warpgroups).  ptxas allocates each role's code against its own limit.
constexpr int HL_SWEEP_REGS_CTRL = 104, HL_SWEEP_REGS_EPI = 200, HL_SWEEP_REGS_PROD = 104;
static_assert(HL_SWEEP_REGS_CTRL + HL_SWEEP_REGS_EPI + 2 * HL_SWEEP_REGS_PROD <= 512, "register file");
constexpr int HL_SWEEP_PLANES = 8;          // planes per staged box
constexpr int HL_SWEEP_BOX_TEXELS = 3072;  // float4 texels of the staged source box (48 KB)
constexpr int HL_SWEEP_SMEM = HL_SWEEP_MAX_D * 4 + 32 + HL_SWEEP_BOX_TEXELS * 16;  // z table, min / max records, box

// Fused plane sweep (SWEEP instantiations, cnm_conv2d_halo2_sweep): the cost-volume operand of conv1.0 never exists as a
// tensor.  Producer warps compute, per 16x16 output tile, the (16+2p)^2 halo region of the volume straight in the act
// layout into a per-CTA scratch slot (two slots per CTA, re-used every other tile: the lines stay in L2), and the A-slab
// TMA producer boxes its slabs out of that slot instead of out of a [P,H,W,D] tensor.
struct SweepFuseParams {
  CUtensorMap map;       // scratch as [2 * gridDim.x slots][R][R][2 D physical channels], box {64, 16, R, 1}
  const float* ref;      // [B,3,H,W] fp32
  const float4* tex;     // packed source views (cnm_pack_source)
  const float* hk;       // [P,12]
  uint8_t* scratch;
  double idmin, idstep;
  int B, D, H, W;
  int vol_map;           // the a_maps index (source * 4) the scratch stands in for
  int pad;               // rows / columns of the region above / left of the tile ((k - 1) / 2)
  int park;              // waiting TMA / epilogue warps back off with nanosleep
};

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
  const float* side_disp;   // 1-channel side input added in the epilogue (cnm_conv_desc.side_disp / side_w), or NULL
  const float* side_w;      // [9][Cout]
  // balanced schedule (cnm_conv_desc.balance, BAL instantiations): see HaloWalk
  // folded bilinear x2 upsampling (cnm_conv_desc.fold, FOLD instantiations): the layer runs on the low-res tensor with
  // 4 x fold_cout output channels, one block of fold_cout per output phase (a, b); phase maps address out[2y+a][2x+b]
  CUtensorMap out_phase[4];
  int fold_cout;
  int map_xoff[HL_MAX_SRCS * 4], map_yoff[HL_MAX_SRCS * 4];   // window origin of a source view (cnm_conv_src.x_off / y_off)
  float* bal_slots;      // [gridDim.x][HALVES][128][BN] fp32 raw accumulators of the contributor pieces
  int* bal_flags;        // [gridDim.x] 1 = the CTA's contributor piece is in its slot (reset by the tile's finishing CTA)
  SweepFuseParams sw;    // SWEEP instantiations only
};

// Work of one CTA (CG = 1) or CTA pair (CG = 2) as a sequence of segments (tile, groups [g0, g1)).
//   static (BAL = false): whole tiles first_tile, first_tile + stride, ...
//   balanced (BAL = true, stream-K): the launch's slab groups -- tiles x num_groups, tile-major -- are cut into one equal
//     contiguous range per unit.  A unit whose range starts inside a tile runs that CONTRIBUTOR piece first and parks the raw
//     accumulators in its CTAs' slots; the unit that owns the tile's first group (the HEAD piece, the last segment of its
//     range) adds the slots of the following units in index order and runs the normal epilogue.  Contributors never wait and
//     are done before the head unit has finished its own range.  240 tiles then cost 74 pairs 3.24 tile times instead of 4.
template <bool BAL>
struct HaloWalk {
  long long pos, end;
  int cur, stride, total, NG;
  int tile, g0, g1;
  __device__ __forceinline__ static long long cut(long long T, int u, int U) { return T * u / U; }
  __device__ __forceinline__ HaloWalk(int num_groups, int total_tiles, int unit, int units) {
    NG = num_groups;
    if (BAL) {
      const long long T = (long long)total_tiles * NG;
      pos = cut(T, unit, units); end = cut(T, unit + 1, units);
    } else {
      cur = unit; stride = units; total = total_tiles;
    }
  }
  __device__ __forceinline__ bool next() {
    if (BAL) {
      if (pos >= end) return false;
      tile = (int)(pos / NG);
      g0 = (int)(pos - (long long)tile * NG);
      const long long left = end - pos;
      g1 = (left < (long long)(NG - g0)) ? g0 + (int)left : NG;
      pos += g1 - g0;
      return true;
    }
    if (cur >= total) return false;
    tile = cur; g0 = 0; g1 = NG;
    cur += stride;
    return true;
  }
};

// ---- 2-CTA PTX wrappers --------------------------------------------------------------------------------------
constexpr uint32_t PEER_BIT_MASK = 0xFEFFFFFFu;  // clears the CTA-rank bit of a shared::cluster address: -> leader CTA

// TMA loads of a CTA pair: data lands in the executing CTA's smem, the transaction bytes are counted on the LEADER's barrier
__device__ __forceinline__ void tma2_load_4d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2,
                                             int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar & PEER_BIT_MASK), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma2_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar & PEER_BIT_MASK), "r"(c0), "r"(c1)
      : "memory");
}
// tcgen05.commit of the pair's MMAs, arriving on the barrier at this offset in BOTH CTAs
__device__ __forceinline__ void tc2_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(bar), "h"((uint16_t)3)
               : "memory");
}
template <bool TF32>
__device__ __forceinline__ void tc2_mma(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  if (TF32)
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n}"
                 ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
                 : "memory");
  else
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n}"
                 ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
                 : "memory");
}
__device__ __forceinline__ void tc2_mma_f8(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::2.kind::f8f6f4 [%0], %1, %2, %3, p;\n}"
               ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
               : "memory");
}
// arrive on the barrier at this offset in the leader CTA (from either CTA of the pair)
__device__ __forceinline__ void mbar_arrive_leader(uint32_t bar) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(bar & PEER_BIT_MASK) : "memory");
}

// mbarrier wait that leaves the issue slots to other warps while it waits (SWEEP kernels: the volume producers share the
// schedulers with the waiting TMA / epilogue warps)
__device__ __forceinline__ void mbar_wait_backoff(uint32_t bar, uint32_t parity, uint32_t ns) {
  uint32_t ok;
  for (;;) {
    asm volatile(
        "{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\nselp.u32 %0, 1, 0, p;\n}"
        : "=r"(ok)
        : "r"(bar), "r"(parity), "r"(1000u)
        : "memory");
    if (ok) return;
    __nanosleep(ns);
  }
}

// CG = 1: one CTA per 16x16 pixel block, two MMA-issuing warps (one per half of the block).  CG = 2: a CTA pair on two
// neighbouring blocks, one M = 256 MMA per half issued by the leader.
// TH = 16: the CTA's pixel block is 16x16 (two M = 128 halves sharing every weight tile).  TH = 8 (CTA pairs only): one
// 16x8 half per CTA -- twice the work items for maps with few blocks (the 48x64 / 24x32 levels), BN = 256 with two
// accumulator stages, and still a slab per (chunk, dx) instead of a pixel tile per tap.
// Border class of a 16x16 pixel block: bit 0 top, 1 bottom, 2 left, 3 right edge of the map.  A group whose ntaps field
// carries flags in bits 4-7 (FOLD instantiations: the border-correction groups of a folded upsampling) runs only on tiles
// whose class has all of them; for a CTA pair the class is the union of the two blocks (the block that does not need the
// group reads zeros through the view's TMA bounds).
__device__ __forceinline__ int halo_block_class(const HaloParams& p, int m) {
  const int x0 = (m % p.tiles_x) * 16, y0 = ((m / p.tiles_x) % p.tiles_y) * 16;
  return (y0 == 0 ? 1 : 0) | (y0 + 16 >= p.Ho ? 2 : 0) | (x0 == 0 ? 4 : 0) | (x0 + 16 >= p.Wo ? 8 : 0);
}

template <int BN, bool TF32, int CG, bool SWEEP = false, int TH = 16, bool BAL = false, bool FOLD = false, bool SIDE = false>
__device__ __forceinline__ void conv_halo_body(const HaloParams& p) {
  static_assert(!SIDE || (!TF32 && !SWEEP && !BAL && !FOLD && TH == 16), "side disparity input: plain 2-byte instantiations");
  pdl_trigger();
  constexpr int HALVES = TH / 8;
  static_assert(!FOLD || (!SWEEP && !TF32 && !BAL && TH == 16), "folded upsampling: 2-byte dtypes, 16x16 blocks, static schedule");
  static_assert(!BAL || (!SWEEP && !TF32 && TH == 16 && 4 * BN <= 512), "balanced schedule: 2-byte dtypes, 16x16 blocks, two accumulator stages");
  static_assert(TH == 16 || (TH == 8 && CG == 2 && !SWEEP), "16x8 blocks: CTA-pair instantiations");
  static_assert(!SWEEP || (!TF32 && 4 * BN <= 512 && TH == 16), "the fused plane sweep needs two accumulator stages (BN <= 128)");
  constexpr int COLS_PER_CHUNK = TF32 ? 32 : 64;
  constexpr int CHUNK_CH = TF32 ? 32 : 64;
  constexpr int NUM_CHUNKS = BN / COLS_PER_CHUNK;
  constexpr int ACC_STAGES = (2 * HALVES * BN <= 512) ? 2 : 1;
  constexpr uint32_t TMEM_COLS = (HALVES * BN * ACC_STAGES) < 32 ? 32 : HALVES * BN * ACC_STAGES;
  constexpr int B_BYTES = (BN / CG) * 128;  // CG = 2: this CTA's half of the weight tile
  constexpr uint32_t IDESC_SHAPE = (1u << 4) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)((128 * CG) >> 4) << 24);  // M = 128 per CTA
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
  const uint32_t off_side = off_head + (p.head_partial ? 9 * BN * 4 : 0);  // side-disparity weights of the N tile [9][BN]
  const uint32_t off_bar = off_side + (SIDE ? 9 * BN * 4 : 0);
  const uint32_t bar0 = base + off_bar;
  auto fullA = [&](int s) { return bar0 + 8u * s; };
  auto emptyA = [&](int s) { return bar0 + 8u * (HL_MAX_A + s); };
  auto fullB = [&](int s) { return bar0 + 8u * (2 * HL_MAX_A + s); };
  auto emptyB = [&](int s) { return bar0 + 8u * (2 * HL_MAX_A + HL_MAX_B + s); };
  auto tfull = [&](int a) { return bar0 + 8u * (2 * HL_MAX_A + 2 * HL_MAX_B + a); };
  auto tempty = [&](int a) { return bar0 + 8u * (2 * HL_MAX_A + 2 * HL_MAX_B + 2 + a); };
  auto vfull = [&](int s) { return bar0 + 8u * (2 * HL_MAX_A + 2 * HL_MAX_B + 4 + s); };  // SWEEP: volume slot s is written
  const uint32_t off_tmem = off_bar + 8u * HL_NUM_BARS;
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(base_ptr + off_tmem);
  float* zs = reinterpret_cast<float*>(base_ptr + ((off_tmem + 16 + 15) & ~15u));  // SWEEP: z_d of the D planes (16B aligned)
  float* bias_s = reinterpret_cast<float*>(base_ptr + off_bias);
  float* head_s = reinterpret_cast<float*>(base_ptr + off_head);
  float* side_s = reinterpret_cast<float*>(base_ptr + off_side);
  const bool fused_head = !TF32 && !SWEEP && p.head_partial != nullptr;
  constexpr bool side = SIDE;   // own instantiations: the extra epilogue state costs the others registers

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);  // warp-uniform for the compiler
  const int lane = threadIdx.x & 31;
  // CG = 2: a super-tile = the same N tile on two neighbouring pixel blocks (m = 2*mm + rank), one per CTA of the pair
  const uint32_t rank = CG == 2 ? cluster_ctarank() : 0u;
  const bool leader = (rank == 0);
  const int total_tiles = (p.tiles_x * p.tiles_y * p.tiles_b / CG) * p.tiles_n;
  const int first_tile = blockIdx.x / CG, tile_stride = gridDim.x / CG;
  auto tile_class = [&](int tile) {   // FOLD: which border groups this tile (pair) runs
    const int m = CG * (tile / p.tiles_n);
    return CG == 2 ? (halo_block_class(p, m) | halo_block_class(p, m + 1)) : halo_block_class(p, m);
  };

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.b_map);
    tma_prefetch_desc(&p.out_map);
    tma_prefetch_desc(&p.a_maps[0]);
  }
  if (warp == 1 && lane == 0) {
    // CG = 1: two issuing warps commit to every barrier the MMAs release; CG = 2: one issuer, 2 x 128 epilogue threads
    constexpr uint32_t ISSUERS = CG == 1 ? 2u : 1u;
    for (int s = 0; s < p.a_stages; ++s) { mbar_init(fullA(s), 1); mbar_init(emptyA(s), ISSUERS); }
    for (int s = 0; s < p.b_stages; ++s) { mbar_init(fullB(s), 1); mbar_init(emptyB(s), ISSUERS); }
    for (int a = 0; a < 2; ++a) { mbar_init(tfull(a), ISSUERS); mbar_init(tempty(a), 128u * CG); }
    if (SWEEP)
      for (int a = 0; a < 2; ++a) mbar_init(vfull(a), HL_SWEEP_THREADS);
    if (BAL) mbar_init(vfull(0), 1);   // balanced schedule: the bulk-copy barrier of balance_gather
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    if (CG == 2) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(base + off_tmem), "r"(TMEM_COLS)
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(base + off_tmem), "r"(TMEM_COLS)
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
  }
  tc_fence_before();
  if (CG == 2) cluster_sync_all();  // both CTAs' barriers are initialised and TMEM is allocated before any remote signal
  else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();  // everything above overlapped the previous kernel's tail; its outputs are visible from here on

  if (SWEEP) {
    // 512 threads share the register file at 128 each; the epilogue warpgroup needs ~200 (64 accumulator columns in
    // flight per thread), the TMA / MMA warps and the volume producers less: re-balance per warpgroup
    if (warp < 4) asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(HL_SWEEP_REGS_CTRL));
    else if (warp < 8) asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(HL_SWEEP_REGS_EPI));
    else asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(HL_SWEEP_REGS_PROD));
  }

  if (warp == 0) {
    // ================= TMA producer, A slabs =================
    // (separate from the B producer so that slab requests run a_stages groups ahead regardless of the B ring depth)
    if (lane == 0) {
      int sa = 0;
      uint32_t pa = 0;
      HaloWalk<BAL> w(p.num_groups, total_tiles, first_tile, tile_stride);
      bool more = w.next();
      int4 g_next = __ldg(reinterpret_cast<const int4*>(p.groups + (more ? w.g0 : 0)));
      int it = 0;
      for (; more; more = w.next(), ++it) {
        const int tile = w.tile;
        int m = CG * (tile / p.tiles_n) + (int)rank;
        const int x0 = (m % p.tiles_x) * 16;
        m /= p.tiles_x;
        const int y0 = (m % p.tiles_y) * TH;
        const int n0 = m / p.tiles_y;
        bool vol_ready = false;
        const int cls = FOLD ? tile_class(tile) : 0;
        for (int g = w.g0; g < w.g1; ++g) {
          const int4 gr = g_next;  // src, c0, dx, ntaps
          g_next = __ldg(reinterpret_cast<const int4*>(p.groups + ((g + 1 == w.g1) ? 0 : g + 1)));  // later segments start at group 0
          if (FOLD && (((gr.w >> 4) & 15) & ~cls)) continue;   // border group of another tile class
          const int pad = p.map_pad[gr.x];
          const uint32_t slab_tx = (uint32_t)p.map_rows[gr.x] * 2048u;
          if (SWEEP && p.sw.park) mbar_wait_backoff(emptyA(sa), pa ^ 1, 200);
          else mbar_wait(emptyA(sa), pa ^ 1);
          if (leader) mbar_expect_tx(fullA(sa), (uint32_t)CG * slab_tx);  // CG = 2: both CTAs' slabs
          const CUtensorMap* map = &p.a_maps[gr.x];
          int c1 = x0 + gr.z, c2 = y0 - pad, c3 = n0;
          if (FOLD) { c1 -= p.map_xoff[gr.x]; c2 -= p.map_yoff[gr.x]; }
          if (SWEEP && gr.x == p.sw.vol_map) {
            // the tile's region of the volume sits in this CTA's scratch slot (written by warps 8-15): columns
            // dx + pad ... of the region, all rows
            if (!vol_ready) {
              mbar_wait_backoff(vfull(it & 1), (uint32_t)(it >> 1) & 1u, 200);
              vol_ready = true;
            }
            map = &p.sw.map;
            c1 = gr.z + p.sw.pad; c2 = 0; c3 = (int)blockIdx.x * 2 + (it & 1);
          }
          if (CG == 2) tma2_load_4d(base + sa * p.slab_bytes, map, fullA(sa), gr.y, c1, c2, c3);
          else tma_load_4d(base + sa * p.slab_bytes, map, fullA(sa), gr.y, c1, c2, c3);
          if (++sa == p.a_stages) { sa = 0; pa ^= 1; }
        }
      }
    }
  } else if (warp == 3) {
    // ================= TMA producer, B (weight) tiles =================
    if (lane == 0) {
      int sb = 0;
      uint32_t pb = 0;
      HaloWalk<BAL> w(p.num_groups, total_tiles, first_tile, tile_stride);
      while (w.next()) {
        const int nt = w.tile % p.tiles_n;
        int ks0 = 0, ks1 = p.num_steps;
        if (BAL) {   // weight steps of the piece: bits 8.. of a group's ntaps field = index of its first weight step
          if (w.g0 > 0) ks0 = (__ldg(&p.groups[w.g0].ntaps) >> 8) & 0xFFFFF;
          if (w.g1 < w.NG) ks1 = (__ldg(&p.groups[w.g1].ntaps) >> 8) & 0xFFFFF;
        }
        if (FOLD) {   // group by group: the border groups of other tile classes are skipped, weight steps with them
          const int cls = tile_class(w.tile);
          for (int g = 0; g < w.NG; ++g) {
            const int nf = __ldg(&p.groups[g].ntaps);
            if (((nf >> 4) & 15) & ~cls) continue;
            for (int kstep = (nf >> 8) & 0xFFFFF; kstep < ((nf >> 8) & 0xFFFFF) + (nf & 15); ++kstep) {
              mbar_wait(emptyB(sb), pb ^ 1);
              if (leader) mbar_expect_tx(fullB(sb), (uint32_t)CG * B_BYTES);
              if (CG == 2) tma2_load_2d(base + off_b + sb * B_BYTES, &p.b_map, fullB(sb), kstep * CHUNK_CH, nt * BN + (int)rank * (BN / 2));
              else tma_load_2d(base + off_b + sb * B_BYTES, &p.b_map, fullB(sb), kstep * CHUNK_CH, nt * BN);
              if (++sb == p.b_stages) { sb = 0; pb ^= 1; }
            }
          }
          continue;
        }
        for (int kstep = ks0; kstep < ks1; ++kstep) {
          if (SWEEP && p.sw.park) mbar_wait_backoff(emptyB(sb), pb ^ 1, 100);
          else mbar_wait(emptyB(sb), pb ^ 1);
          if (leader) mbar_expect_tx(fullB(sb), (uint32_t)CG * B_BYTES);  // CG = 2: both halves
          if (CG == 2) tma2_load_2d(base + off_b + sb * B_BYTES, &p.b_map, fullB(sb), kstep * CHUNK_CH, nt * BN + (int)rank * (BN / 2));
          else tma_load_2d(base + off_b + sb * B_BYTES, &p.b_map, fullB(sb), kstep * CHUNK_CH, nt * BN);
          if (++sb == p.b_stages) { sb = 0; pb ^= 1; }
        }
      }
    }
  } else if (warp == 1 || (CG == 1 && warp == 2)) {
    // ================= MMA issuers =================
    // CG = 1: two issuing warps, one per half of the pixel block (warp 1 -> rows 0-7 / accumulator columns [0,BN),
    // warp 2 -> rows 8-15 / [BN,2BN)): a single lane needs ~100 instructions per weight step (barrier wait,
    // descriptors, 8 MMAs, commit), which at BN <= 128 is as long as the MMAs themselves take
    // (profiles/r01_conv_halo_ncu.md); every barrier the MMAs release therefore expects two commits.
    // CG = 2: the leader CTA's warp 1 issues one M = 256 MMA per half, spanning both CTAs.
    // The whole warp walks the loop (all lanes poll the barriers) and one elected lane issues: the control flow stays
    // warp-uniform, so the compiler holds descriptors in uniform registers instead of wrapping every tcgen05
    // instruction in a per-lane election loop.
    if (leader) {
      const int half = CG == 1 ? warp - 1 : 0;
      int sa = 0, sb = 0;
      uint32_t pa = 0, pb = 0;
      int it = 0;
      HaloWalk<BAL> w(p.num_groups, total_tiles, first_tile, tile_stride);
      for (; w.next(); ++it) {
        const int as = (ACC_STAGES == 2) ? (it & 1) : 0;
        const uint32_t aphase = (ACC_STAGES == 2) ? ((it >> 1) & 1) : (it & 1);
        mbar_wait(tempty(as), aphase ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)(as * HALVES * BN + half * BN);
        uint32_t first = 1;
        const int cls = FOLD ? tile_class(w.tile) : 0;
        for (int g = w.g0; g < w.g1; ++g) {
          const int nfield = __ldg(&p.groups[g].ntaps);
          if (FOLD && (((nfield >> 4) & 15) & ~cls)) continue;
          const int ntaps = nfield & 15;
          const int trows = (((nfield >> 28) & 7) + 1) * 2048;   // bytes between the row windows of consecutive taps
          mbar_wait(fullA(sa), pa);
          const uint32_t slab = base + sa * p.slab_bytes + half * 8 * 2048;
          for (int t = 0; t < ntaps; ++t) {
            mbar_wait(fullB(sb), pb);
            tc_fence_after();
            if (elect_one()) {
              const uint64_t bdesc = umma_desc(base + off_b + sb * B_BYTES);
              const uint64_t a0 = umma_desc(slab + t * trows), a1 = umma_desc(slab + t * trows + 8 * 2048);
              const bool f8 = mixed && g < (p.num_groups >> 1);
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                const uint32_t acc = (first && k == 0) ? 0u : 1u;
                if (CG == 2) {
                  if (f8) {
                    tc2_mma_f8(tmem_d, a0 + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, acc);
                    if (HALVES == 2) tc2_mma_f8(tmem_d + BN, a1 + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, acc);
                  } else {
                    tc2_mma<TF32>(tmem_d, a0 + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, acc);
                    if (HALVES == 2) tc2_mma<TF32>(tmem_d + BN, a1 + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, acc);
                  }
                } else {
                  if (f8) tc_mma_f8(tmem_d, a0 + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, acc);
                  else tc_mma<TF32>(tmem_d, a0 + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, acc);
                }
              }
              if (CG == 2) {
                tc2_commit(emptyB(sb));
                if (t == ntaps - 1) tc2_commit(emptyA(sa));
              } else {
                tc_commit(emptyB(sb));
                if (t == ntaps - 1) tc_commit(emptyA(sa));
              }
            }
            __syncwarp();
            first = 0;
            if (++sb == p.b_stages) { sb = 0; pb ^= 1; }
          }
          if (++sa == p.a_stages) { sa = 0; pa ^= 1; }
        }
        if (elect_one()) {
          if (CG == 2) tc2_commit(tfull(as));
          else tc_commit(tfull(as));
        }
        __syncwarp();
      }
    }
  } else if (SWEEP && warp >= 8) {
    // ================= fused plane sweep: volume producers =================
    // Per tile: the (16 + 2 pad)^2 region of the cost volume around it, all D planes, in the act layout (CNM_F16F8 rows:
    // fp16 | lo8 | hi8), zeros outside the image (= the convolution's padding).  A thread owns up to two pixels of the
    // region.  Eight planes at a time: (A) every thread computes its sample positions and the warps reduce their
    // bounding box; (L) the box of source texels is copied to shared memory with coalesced loads; (S) the samples are
    // blended from the staged copy -- two producer warps per scheduler cannot hide L2 latency behind four dependent
    // gathers per sample (ncu on the direct version: 0.3 instructions per cycle and scheduler, half of the stalls on
    // L1TEX), but a shared-memory tap returns in ~30 cycles.  A box that does not fit (extreme zoom) is sampled through
    // L1 instead; both paths return the same texels, so the rows equal cnm_cost_volume's bit for bit.
    // Slot it & 1 of the scratch is free again when the MMAs of tile it - 2 have completed (their commit on the
    // accumulator-full barrier of stage it & 1 also covers the TMA reads of that slot).
    const int pt = (int)threadIdx.x - 256, pw = warp - 8;
    const SweepFuseParams& sw = p.sw;
    const int D = sw.D, H = sw.H, W = sw.W;
    const int R = 16 + 2 * sw.pad, npx = R * R;
    const size_t px_bytes = (size_t)D * 4;  // 2 D physical 2-byte channels
    const size_t HW = (size_t)H * W;
    int* mm = reinterpret_cast<int*>(zs + HL_SWEEP_MAX_D);      // [2][4]: min x0, min y0, max x0, max y0 (group parity)
    float4* box = reinterpret_cast<float4*>(mm + 8);
    for (int i = pt; i < D; i += HL_SWEEP_THREADS) zs[i] = sweep_plane_z(sw.idmin, sw.idstep, i);
    if (pt < 8) mm[pt] = (pt & 2) ? INT_MIN : INT_MAX;
    asm volatile("bar.sync 2, 256;" ::: "memory");
    int it = 0;
    uint32_t grp = 0;  // running group count: parity selects the min / max record
    for (int tile = first_tile; tile < total_tiles; tile += tile_stride, ++it) {
      int m = CG * (tile / p.tiles_n) + (int)rank;
      const int x0 = (m % p.tiles_x) * 16;
      m /= p.tiles_x;
      const int y0 = (m % p.tiles_y) * TH;
      const int n0 = m / p.tiles_y;
      const int slot = it & 1;
      if (it >= 2) {
        if (lane == 0) mbar_wait_backoff(tfull(slot), (uint32_t)((it - 2) >> 1) & 1u, 200);
        __syncwarp();
      }
      uint8_t* sc = sw.scratch + ((size_t)((int)blockIdx.x * 2 + slot) * npx) * px_bytes;
      const float* h = sw.hk + (size_t)n0 * 12;
      const float* rimg = sw.ref + (size_t)(n0 % sw.B) * 3 * HW;
      const int pitch = W + 4;
      const float4* tp = sw.tex + (size_t)n0 * (H + 3) * pitch + pitch + 1;  // texel (0,0)
      SweepPixelL1 sp[2];
      h16f8* row[2];
      bool live[2];
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        const int idx = pt + k * HL_SWEEP_THREADS;
        const int ry = idx / R, rx = idx - ry * R;
        const int x = x0 - sw.pad + rx, y = y0 - sw.pad + ry;
        row[k] = reinterpret_cast<h16f8*>(sc + (size_t)idx * px_bytes);
        live[k] = idx < npx && x >= 0 && x < W && y >= 0 && y < H;
        sp[k].pitch = pitch; sp[k].tp = tp;
        sp[k].xmax = (float)W; sp[k].ymax = (float)H;
        if (live[k]) {
          sweep_l1_setup(sp[k], h, x, y);
          const float* rp = rimg + (size_t)y * W + x;
          sp[k].r0 = __ldg(rp); sp[k].r1 = __ldg(rp + HW); sp[k].r2 = __ldg(rp + 2 * HW);
        } else if (idx < npx) {
          uint4* z4 = reinterpret_cast<uint4*>(row[k]);
          for (int i = 0; i < (int)(px_bytes / 16); ++i) z4[i] = make_uint4(0u, 0u, 0u, 0u);
        }
      }
#pragma unroll 1
      for (int d0 = 0; d0 < D; d0 += HL_SWEEP_PLANES, ++grp) {
        // ---- (A) positions of this thread's samples, bounding box of their upper-left taps
        float ix[2][HL_SWEEP_PLANES], iy[2][HL_SWEEP_PLANES];
        int mnx = INT_MAX, mny = INT_MAX, mxx = INT_MIN, mxy = INT_MIN;
#pragma unroll
        for (int k = 0; k < 2; ++k)
          if (live[k]) {
#pragma unroll
            for (int i = 0; i < HL_SWEEP_PLANES; ++i) {
              sweep_l1_position(sp[k], zs[d0 + i], ix[k][i], iy[k][i]);
              const int tx = (int)floorf(ix[k][i]), ty = (int)floorf(iy[k][i]);
              mnx = min(mnx, tx); mxx = max(mxx, tx);
              mny = min(mny, ty); mxy = max(mxy, ty);
            }
          }
        int* rec = mm + 4 * (grp & 1u);
        mnx = __reduce_min_sync(0xffffffffu, mnx); mny = __reduce_min_sync(0xffffffffu, mny);
        mxx = __reduce_max_sync(0xffffffffu, mxx); mxy = __reduce_max_sync(0xffffffffu, mxy);
        if (lane == 0 && mnx <= mxx) {
          atomicMin(rec + 0, mnx); atomicMin(rec + 1, mny);
          atomicMax(rec + 2, mxx); atomicMax(rec + 3, mxy);
        }
        asm volatile("bar.sync 2, 256;" ::: "memory");   // also: every thread is done sampling the previous box
        const int bx = rec[0], by = rec[1];
        const int bw = rec[2] - bx + 2, bh = rec[3] - by + 2;   // taps x0, x0 + 1 / y0, y0 + 1
        const bool staged = bw * bh <= HL_SWEEP_BOX_TEXELS;      // CTA-uniform
        // ---- (L) the box, one row per warp and pass, 32 texels per load instruction
        if (staged) {
          for (int r = pw; r < bh; r += HL_SWEEP_THREADS / 32) {
            const float4* srow = tp + (by + r) * pitch + bx;
            for (int c = lane; c < bw; c += 32) box[r * bw + c] = __ldg(srow + c);
          }
        }
        asm volatile("bar.sync 2, 256;" ::: "memory");
        if (pt == 0) { rec[0] = INT_MAX; rec[1] = INT_MAX; rec[2] = INT_MIN; rec[3] = INT_MIN; }   // next use: group grp + 2
        // ---- (S) blend and store
        const float4* bo = box - (by * bw + bx);
#pragma unroll
        for (int k = 0; k < 2; ++k)
          if (live[k]) {
            float res[HL_SWEEP_PLANES];
            if (staged) {
#pragma unroll
              for (int i = 0; i < HL_SWEEP_PLANES; ++i)
                res[i] = sweep_l1_blend(sp[k], ix[k][i], iy[k][i],
                                        [&](int tx, int ty, float4& t00, float4& t10, float4& t01, float4& t11) {
                                          const float4* t = bo + ty * bw + tx;
                                          t00 = t[0]; t10 = t[1]; t01 = t[bw]; t11 = t[bw + 1];
                                        });
            } else {
#pragma unroll
              for (int i = 0; i < HL_SWEEP_PLANES; ++i)
                res[i] = sweep_l1_blend(sp[k], ix[k][i], iy[k][i], SweepFetchGlobal{tp, pitch});
            }
            h16f8* o = row[k] + ActTraits<h16f8>::phys(d0);
#pragma unroll
            for (int i = 0; i < HL_SWEEP_PLANES; i += 8) Vec16<h16f8>::store(o + i, res + i);
          }
      }
      // generic-proxy writes -> visible to the TMA (async proxy) reads the A producer issues after its wait
      asm volatile("fence.proxy.async;" ::: "memory");
      mbar_arrive(vfull(slot));
    }
  } else if (warp >= 4 && warp < 8) {
    // ================= epilogue =================
    const int q = warp - 4;
    const int row = q * 32 + lane;
    const int et = threadIdx.x - 128;
    const bool storer = (et == 0);
    int it = 0;
    int bias_tile = -1;
    uint32_t store_cnt = 0;
    uint32_t gather_phase = 0;

    HaloWalk<BAL> w(p.num_groups, total_tiles, first_tile, tile_stride);
    for (; w.next(); ++it) {
      const int tile = w.tile;
      const int nt = tile % p.tiles_n;
      int m = CG * (tile / p.tiles_n) + (int)rank;
      const int x0 = (m % p.tiles_x) * 16;
      m /= p.tiles_x;
      const int y0 = (m % p.tiles_y) * TH;
      const int n0 = m / p.tiles_y;
      const int as = (ACC_STAGES == 2) ? (it & 1) : 0;
      const uint32_t aphase = (ACC_STAGES == 2) ? ((it >> 1) & 1) : (it & 1);
      if (BAL && w.g0 > 0) {
        // contributor piece: raw accumulators of both halves -> this CTA's slot, then the flag
        mbar_wait(tfull(as), aphase);
        tc_fence_after();
        balance_park(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * HALVES * BN), HALVES * BN,
                     p.bal_slots + (size_t)blockIdx.x * (BALANCE_SLOT_BYTES / 4), row, p.bal_flags + blockIdx.x, storer);
        tc_fence_before();
        if (CG == 2) mbar_arrive_leader(tempty(as));
        else mbar_arrive(tempty(as));
        continue;
      }
      // head piece of a balanced tile: the next `ncontrib` units hold the rest of its groups (same CTA rank in each)
      int ncontrib = 0;
      if (BAL && w.g1 < w.NG) {
        const long long T = (long long)total_tiles * w.NG, tile_end = (long long)(tile + 1) * w.NG;
        while (HaloWalk<BAL>::cut(T, first_tile + 1 + ncontrib, tile_stride) < tile_end) ++ncontrib;
      }
      if (bias_tile != nt) {
        epi_bar();
        for (int i = et; i < BN; i += 128) bias_s[i] = p.bias ? __ldg(p.bias + nt * BN + i) : 0.0f;
        if (fused_head)
          for (int i = et; i < 9 * BN; i += 128) head_s[i] = __ldg(p.head_w + (i / BN) * p.cout + nt * BN + i % BN);
        if (side)
          for (int i = et; i < 9 * BN; i += 128) side_s[i] = __ldg(p.side_w + (i / BN) * p.cout + nt * BN + i % BN);
        bias_tile = nt;
        epi_bar();  // epi_finalize reads bias_s right after the TMEM load, before any other barrier
      }
      if (SWEEP && p.sw.park) mbar_wait_backoff(tfull(as), aphase, 500);
      else mbar_wait(tfull(as), aphase);
      tc_fence_after();
      if (BAL && ncontrib)   // the tile's other pieces -> added to the accumulator in TMEM (the rings are idle: last segment)
        balance_gather(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * HALVES * BN), HALVES * BN, p.bal_slots,
                       p.bal_flags, (int)blockIdx.x + CG, CG, ncontrib, base_ptr, vfull(0), gather_phase, row, storer);
#pragma unroll 1
      for (int half = 0; half < HALVES; ++half) {
        const uint32_t tmem_row = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * HALVES * BN + half * BN);
        float sd[9];
        if (side) side_disp_taps(p.side_disp, n0, y0 + 8 * half + (row >> 4), x0 + (row & 15), p.Ho, p.Wo, sd);
#pragma unroll 1
        for (int ch = 0; ch < NUM_CHUNKS; ++ch) {
          uint32_t v[COLS_PER_CHUNK];
          tmem_ld32(tmem_row + ch * COLS_PER_CHUNK, v);
          if (!TF32) tmem_ld32(tmem_row + ch * COLS_PER_CHUNK + 32, v + 32);
          tmem_ld_wait();
          if (side) epi_finalize_side<COLS_PER_CHUNK>(v, bias_s + ch * COLS_PER_CHUNK, p.relu, p.acc_scale, p.range_ctr, sd,
                                                      side_s + ch * COLS_PER_CHUNK, BN);
          else epi_finalize<COLS_PER_CHUNK>(v, bias_s + ch * COLS_PER_CHUNK, p.relu, p.acc_scale, p.range_ctr);
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
            if (storer) {
              if (p.out_stages == 2) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
              else asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            }
            epi_bar();
            epi_store_row<TF32>(v, stg + row * 128, row, p.mode, part);
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            epi_bar();
            if (storer) {
              if (FOLD) {   // column block -> (output phase, channel inside the phase)
                const int cg = nt * BN + ch * CHUNK_CH, ph = cg / p.fold_cout, cc = cg - ph * p.fold_cout;
                tma_store_4d(&p.out_phase[ph], smem_u32(stg), p.c_out_offset + (p.mode ? 2 * cc + part * 64 : cc), x0, y0 + 8 * half, n0);
              } else {
                tma_store_4d(&p.out_map, smem_u32(stg), p.c_out_offset + (p.mode ? 2 * (nt * BN + ch * CHUNK_CH) + part * 64 : nt * BN + ch * CHUNK_CH),
                             x0, y0 + 8 * half, n0);
              }
              asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
          }
        }
      }
      tc_fence_before();
      if (CG == 2) mbar_arrive_leader(tempty(as));  // 2 x 128 arrivals on the leader's barrier release the accumulator stage
      else mbar_arrive(tempty(as));
    }
    if (storer) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }

  if (SWEEP) {  // back to the launch allocation before the roles join
    if (warp < 4) asm volatile("setmaxnreg.inc.sync.aligned.u32 128;");
    else if (warp < 8) asm volatile("setmaxnreg.dec.sync.aligned.u32 128;");
    else asm volatile("setmaxnreg.inc.sync.aligned.u32 128;");
  }
  tc_fence_before();
  if (CG == 2) cluster_sync_all();  // the peer may still be signalling this CTA's barriers / reading its operands
  else __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    if (CG == 2) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
    else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}

template <int BN, bool TF32>
__global__ void __launch_bounds__(256, 1) conv_halo_kernel(const __grid_constant__ HaloParams p) {
  conv_halo_body<BN, TF32, 1>(p);
}
template <int BN, bool TF32>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(256, 1) conv_halo2_kernel(const __grid_constant__ HaloParams p) {
  conv_halo_body<BN, TF32, 2>(p);
}

// balanced (stream-K) schedule, 2-byte dtypes, column tile <= 128
template <int BN>
__global__ void __launch_bounds__(256, 1) conv_halob_kernel(const __grid_constant__ HaloParams p) {
  conv_halo_body<BN, false, 1, false, 16, true>(p);
}
template <int BN>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(256, 1) conv_halo2b_kernel(const __grid_constant__ HaloParams p) {
  conv_halo_body<BN, false, 2, false, 16, true>(p);
}

// side disparity input in the epilogue (cnm_conv_desc.side_disp), 2-byte dtypes, column tile <= 128
template <int BN>
__global__ void __launch_bounds__(256, 1) conv_halo_side_kernel(const __grid_constant__ HaloParams p) {
  conv_halo_body<BN, false, 1, false, 16, false, false, true>(p);
}
template <int BN>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(256, 1) conv_halo2_side_kernel(const __grid_constant__ HaloParams p) {
  conv_halo_body<BN, false, 2, false, 16, false, false, true>(p);
}

// folded bilinear x2 upsampling (cnm_conv_desc.fold), 2-byte dtypes, column tile 128
__global__ void __launch_bounds__(256, 1) conv_halof_kernel(const __grid_constant__ HaloParams p) {
  conv_halo_body<128, false, 1, false, 16, false, true>(p);
}
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(256, 1) conv_halo2f_kernel(const __grid_constant__ HaloParams p) {
  conv_halo_body<128, false, 2, false, 16, false, true>(p);
}

// 16x8 pixel blocks on CTA pairs (maps with few 16x16 blocks)
template <int BN>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(256, 1) conv_halo2h_kernel(const __grid_constant__ HaloParams p) {
  conv_halo_body<BN, false, 2, false, 8>(p);
}

// conv1.0 with the plane sweep fused in: CTA pairs + eight volume-producer warps per CTA
template <int BN>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(512, 1) conv_halo2_sweep_kernel(const __grid_constant__ HaloParams p) {
  conv_halo_body<BN, false, 2, true>(p);
}

template <int BN, bool TF32, int CG>
static int launch_halo(const HaloParams& p, int grid, int smem, cudaStream_t s) {
  static DynSmemLimit limit;
  const char* fn = CG == 2 ? "cnm_conv2d_halo2" : "cnm_conv2d_halo";
  cudaError_t e = CG == 2 ? limit.ensure(conv_halo2_kernel<BN, TF32>, smem) : limit.ensure(conv_halo_kernel<BN, TF32>, smem);
  if (e != cudaSuccess) return fail(CNM_ERR_CUDA, "%s: smem attribute (%d B): %s", fn, smem, cudaGetErrorString(e));
  if (CG == 2) launch_pdl(conv_halo2_kernel<BN, TF32>, dim3(grid), dim3(256), smem, s, p);
  else launch_pdl(conv_halo_kernel<BN, TF32>, dim3(grid), dim3(256), smem, s, p);
  return after_launch(fn);
}

}  // namespace cnm

using namespace cnm;

// Shared host side of cnm_conv2d_halo (CG = 1) and cnm_conv2d_halo2 (CG = 2)
static int64_t sweep_scratch_bytes(int D, int k) {
  const int64_t R = 16 + (k - 1);
  return (int64_t)2 * 148 * R * R * 4 * D;  // two slots per CTA, a full grid of CTAs, 4 D bytes per pixel (CNM_F16F8)
}

template <int CG>
static int conv2d_halo(const cnm_conv_desc* d, const cnm_conv_src* srcs, const int32_t* map_info,
                       const cnm_conv_group* groups, int num_groups, const void* wp, const float* bias, void* out,
                       void* stream, const cnm_sweep_fuse* sweep = nullptr) {
  const char* fn = sweep ? "cnm_conv2d_halo2_sweep" : (CG == 2 ? "cnm_conv2d_halo2" : "cnm_conv2d_halo");
  int rc = check_conv_args(fn, d, srcs, groups, wp, out);
  if (rc) return rc;
  CNM_REQUIRE(d->ksplit <= 1, CNM_ERR_UNSUPPORTED_SHAPE, "%s: split-K is a cnm_conv2d_tc feature", fn);
  CNM_REQUIRE(map_info && num_groups > 0, CNM_ERR_INVALID_ARGUMENT, "%s: bad group list", fn);
  CNM_REQUIRE(d->Ho >= 16 && d->Wo >= 16, CNM_ERR_UNSUPPORTED_SHAPE,
              "%s: output %dx%d smaller than the 16x16 tile (use cnm_conv2d_tc)", fn, d->Ho, d->Wo);
  CNM_REQUIRE(d->Cout % 64 == 0, CNM_ERR_UNSUPPORTED_SHAPE, "%s: Cout=%d must be a multiple of 64", fn, d->Cout);
  CNM_REQUIRE(d->tile_h == 0 || d->tile_h == 16 || (d->tile_h == 8 && CG == 2 && !sweep && d->dtype != CNM_TF32),
              CNM_ERR_UNSUPPORTED_SHAPE, "%s: tile_h = %d (8 = 16x8 blocks: cnm_conv2d_halo2, 2-byte dtypes)", fn, d->tile_h);
  const int TH = d->tile_h == 8 ? 8 : 16;
  const bool fold = d->fold != 0;
  CNM_REQUIRE(!fold || (!sweep && TH == 16), CNM_ERR_UNSUPPORTED_SHAPE, "%s: fold goes with 16x16 blocks and no fused sweep", fn);
  CNM_REQUIRE(CG == 1 || (((d->Wo + 15) / 16) * ((d->Ho + TH - 1) / TH) * d->N) % 2 == 0, CNM_ERR_UNSUPPORTED_SHAPE,
              "%s: needs an even number of 16x%d pixel blocks (CTA pairs)", fn, TH);
  const int eb = elem_bytes(d->dtype), ch = chunk_channels(d->dtype);
  HaloParams p;
  memset(&p, 0, sizeof(p));
  // map_info[src*4 + sub] = {rows above the tile, slab rows}; sub = 0 for stride 1, parity py*2+px for stride 2
  int rows_max = 0;
  const int nsub = d->stride == 2 ? 4 : 1;
  const int mstride = fold ? 1 : 4;   // fold: operand map = source index (up to 12 sources, stride 1)
  for (int i = 0; i < d->num_srcs; ++i)
    for (int sub = 0; sub < nsub; ++sub) {
      const int m = i * mstride + sub, pad = map_info[2 * m], rows = map_info[2 * m + 1];
      if (rows == 0) continue;  // sub-image not used by any group
      CNM_REQUIRE(pad >= 0 && pad <= 4 && rows >= 16 && rows <= 24, CNM_ERR_UNSUPPORTED_SHAPE,
                  "%s: map %d has pad %d / %d slab rows", fn, m, pad, rows);
      p.map_pad[m] = pad;
      p.map_rows[m] = rows - (16 - TH);   // map_info counts a 16-row block; a 16x8 block needs 8 rows less
      if (p.map_rows[m] > rows_max) rows_max = p.map_rows[m];
    }
  CNM_REQUIRE(rows_max > 0, CNM_ERR_INVALID_ARGUMENT, "%s: no operand maps", fn);
  const int pmax = (rows_max - TH + 1) / 2;
  p.slab_bytes = rows_max * 2048;
  p.tiles_x = (d->Wo + 15) / 16;
  p.tiles_y = (d->Ho + TH - 1) / TH;
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
    CNM_REQUIRE(d->Cout % d->reserved == 0, CNM_ERR_INVALID_ARGUMENT, "%s: BN=%d does not divide Cout=%d", fn,
                d->reserved, d->Cout);
    BN = d->reserved;
  }
  if (d->side_disp && BN == 256) BN = 128;   // the SIDE instantiations: column tile <= 128
  if (fold) BN = 128;   // the FOLD instantiations: two accumulator stages, a phase block (>= 64 channels) never straddles chunks
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
  p.side_disp = d->side_disp;
  p.side_w = d->side_w;
  CNM_REQUIRE(!d->side_disp || (!sweep && !fold && !d->balance && TH == 16), CNM_ERR_UNSUPPORTED_SHAPE,
              "%s: no side disparity input on fused-sweep / fold / balanced / 16x8-block layers", fn);
  p.Ho = d->Ho; p.Wo = d->Wo; p.cout = d->Cout;
  // smem budget: 227 KB - alignment slack - staging - bias - barriers.  Ring depths are sized to cover the loaded
  // L2 latency (~3-4k cycles): a B tile lasts 2 x 4 MMAs (1024 / 512 / 384 cycles at BN = 256 / 128 / 64), a slab
  // lasts ntaps B tiles.
  p.out_stages = BN == 64 ? 2 : 1;
  if (sweep) {
    // the fused plane sweep: see SweepFuseParams.  One N tile (every tile computes its region of the volume once), two
    // accumulator stages (the producers pace themselves on the accumulator-full barriers), CNM_F16F8 rows.
    const int vs = sweep->vol_src;
    CNM_REQUIRE(CG == 2 && d->stride == 1 && d->dtype == CNM_F16F8, CNM_ERR_UNSUPPORTED_SHAPE,
                "%s: stride-1 CNM_F16F8 layers on CTA pairs only", fn);
    CNM_REQUIRE(BN == d->Cout && BN <= 128, CNM_ERR_UNSUPPORTED_SHAPE, "%s: Cout=%d must be one tile of at most 128", fn, d->Cout);
    CNM_REQUIRE(!d->head_partial, CNM_ERR_UNSUPPORTED_SHAPE, "%s: no fused head on this layer", fn);
    CNM_REQUIRE(vs >= 0 && vs < d->num_srcs && p.map_rows[vs * 4] > 16, CNM_ERR_INVALID_ARGUMENT, "%s: vol_src %d", fn, vs);
    CNM_REQUIRE(sweep->ref && sweep->packed_src && sweep->hk && sweep->scratch, CNM_ERR_INVALID_ARGUMENT, "%s: null pointer", fn);
    CNM_REQUIRE(sweep->D >= 64 && sweep->D % 64 == 0 && sweep->D <= HL_SWEEP_MAX_D && srcs[vs].C == 2 * sweep->D,
                CNM_ERR_UNSUPPORTED_SHAPE, "%s: D=%d planes (64 or 128) against %d physical channels", fn, sweep->D, srcs[vs].C);
    CNM_REQUIRE(sweep->B > 0 && d->N % sweep->B == 0, CNM_ERR_INVALID_ARGUMENT, "%s: N=%d pairs, B=%d reference frames", fn, d->N, sweep->B);
    const int pad = p.map_pad[vs * 4], R = 16 + 2 * pad;
    CNM_REQUIRE(p.map_rows[vs * 4] == R, CNM_ERR_UNSUPPORTED_SHAPE, "%s: slab of %d rows for pad %d", fn, p.map_rows[vs * 4], pad);
    CNM_REQUIRE(((uintptr_t)sweep->scratch & 127) == 0 && sweep->scratch_bytes >= sweep_scratch_bytes(sweep->D, 2 * pad + 1),
                CNM_ERR_INVALID_ARGUMENT, "%s: scratch must be 128B aligned and hold cnm_sweep_scratch_bytes() = %lld bytes", fn,
                (long long)sweep_scratch_bytes(sweep->D, 2 * pad + 1));
    SweepFuseParams& sw = p.sw;
    sw.ref = sweep->ref; sw.tex = (const float4*)sweep->packed_src; sw.hk = sweep->hk; sw.scratch = (uint8_t*)sweep->scratch;
    sw.idmin = sweep->idepth_min;
    sw.idstep = (sweep->idepth_max - sweep->idepth_min) / (double)(sweep->D - 1);
    sw.B = sweep->B; sw.D = sweep->D; sw.H = d->Ho; sw.W = d->Wo;
    sw.vol_map = vs * 4; sw.pad = pad;
    sw.park = 1;
    const uint64_t pix = (uint64_t)sweep->D * 4;
    const uint64_t dims[4] = {(uint64_t)2 * sweep->D, (uint64_t)R, (uint64_t)R, (uint64_t)2 * 148};
    const uint64_t str[3] = {pix, pix * R, pix * R * R};
    const uint32_t box[4] = {(uint32_t)ch, 16u, (uint32_t)R, 1u};
    rc = encode_map(&sw.map, d->dtype, 4, sweep->scratch, dims, str, box);
    if (rc) return rc;
  }
  const int fixed = 1024 + p.out_stages * HL_TILE_BYTES + BN * 4 + (d->head_partial ? 9 * BN * 4 : 0) + (d->side_disp ? 9 * BN * 4 : 0) +
                    8 * HL_NUM_BARS + 16 +
                    (sweep ? HL_SWEEP_SMEM + 16 : 0);
  const int budget = 232448 - fixed;
  const int want_b = CG == 2 ? 8 : (BN == 256 ? 4 : (BN == 128 ? 7 : 8));
  const int b_tile = BN * 128 / CG;   // bytes of one weight tile (CG = 2: this CTA's half)
  p.a_stages = (pmax >= 3 || BN == 256) ? 2 : (BN == 64 ? 4 : 3);
  p.b_stages = (budget - p.a_stages * p.slab_bytes) / b_tile;
  if (p.b_stages > want_b) p.b_stages = want_b;
  if (p.b_stages > HL_MAX_B) p.b_stages = HL_MAX_B;
  while (p.b_stages < 3 && p.a_stages > 2) {
    p.a_stages -= 1;
    p.b_stages = (budget - p.a_stages * p.slab_bytes) / b_tile;
  }
  CNM_REQUIRE(p.a_stages >= 2 && p.b_stages >= 2, CNM_ERR_UNSUPPORTED_SHAPE,
              "%s: slab of %d B does not fit", fn, p.slab_bytes);
  {  // spend what is left on more slabs
    const int left = budget - p.a_stages * p.slab_bytes - p.b_stages * b_tile;
    const int extra = left / p.slab_bytes;
    p.a_stages += extra;
    if (p.a_stages > HL_MAX_A) p.a_stages = HL_MAX_A;
  }
  const int smem = fixed + p.a_stages * p.slab_bytes + p.b_stages * b_tile;

  for (int i = 0; i < d->num_srcs; ++i) {
    const cnm_conv_src& s = srcs[i];
    const uint64_t pix = (uint64_t)s.C * eb;
    for (int sub = 0; sub < nsub; ++sub) {
      const int m = i * mstride + sub;
      if (p.map_rows[m] == 0) continue;
      if (sweep && m == p.sw.vol_map) continue;  // the scratch map stands in for it
      const uint32_t box[4] = {(uint32_t)ch, 16u, (uint32_t)p.map_rows[m], 1u};
      if (d->stride == 1) {
        // a view keeps its parent's pitches; TMA zero-fills everything outside its H x W window
        const uint64_t pw = s.pitch_w ? (uint64_t)s.pitch_w : (uint64_t)s.W, phw = s.pitch_hw ? (uint64_t)s.pitch_hw : pw * s.H;
        const uint64_t dims[4] = {(uint64_t)s.C, (uint64_t)s.W, (uint64_t)s.H, (uint64_t)d->N};
        const uint64_t str[3] = {pix, pix * pw, pix * phw};
        p.map_xoff[m] = s.x_off; p.map_yoff[m] = s.y_off;
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
    const uint32_t bbox[2] = {(uint32_t)ch, (uint32_t)(BN / CG)};
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
    if (fold) {   // phase (a, b): out[n][2y + a][2x + b], addressed with low-res coordinates through doubled strides
      p.fold_cout = d->Cout / 4;
      const uint64_t W2 = 2 * (uint64_t)d->Wo, H2 = 2 * (uint64_t)d->Ho;
      const uint64_t pstr[3] = {2 * pix, 2 * pix * W2, pix * W2 * H2};
      for (int ph = 0; ph < 4; ++ph) {
        const uint8_t* base = (const uint8_t*)out + ((uint64_t)(ph >> 1) * W2 + (ph & 1)) * pix;
        rc = encode_map(&p.out_phase[ph], d->dtype, 4, base, dims, pstr, box);
        if (rc) return rc;
      }
    }
  }
  const int total = (m_tiles / CG) * p.tiles_n;  // CG = 2: super-tiles, one per CTA pair
  const int grid = CG * (total < 148 / CG ? total : 148 / CG);
  cudaStream_t s = (cudaStream_t)stream;
  const bool tf32 = d->dtype == CNM_TF32;  // the 2-byte modes share the kind::f16 instantiations
  if (d->side_disp) {
    static DynSmemLimit s64, s128, s2_64, s2_128;
    cudaError_t e;
    if (CG == 2) e = BN == 128 ? s2_128.ensure(conv_halo2_side_kernel<128>, smem) : s2_64.ensure(conv_halo2_side_kernel<64>, smem);
    else e = BN == 128 ? s128.ensure(conv_halo_side_kernel<128>, smem) : s64.ensure(conv_halo_side_kernel<64>, smem);
    if (e != cudaSuccess) return fail(CNM_ERR_CUDA, "%s: smem attribute (%d B): %s", fn, smem, cudaGetErrorString(e));
    if (CG == 2) {
      if (BN == 128) launch_pdl(conv_halo2_side_kernel<128>, dim3(grid), dim3(256), smem, s, p);
      else launch_pdl(conv_halo2_side_kernel<64>, dim3(grid), dim3(256), smem, s, p);
    } else {
      if (BN == 128) launch_pdl(conv_halo_side_kernel<128>, dim3(grid), dim3(256), smem, s, p);
      else launch_pdl(conv_halo_side_kernel<64>, dim3(grid), dim3(256), smem, s, p);
    }
    return after_launch(fn);
  }
  if (fold) {
    static DynSmemLimit f1, f2;
    cudaError_t e = CG == 2 ? f2.ensure(conv_halo2f_kernel, smem) : f1.ensure(conv_halof_kernel, smem);
    if (e != cudaSuccess) return fail(CNM_ERR_CUDA, "%s: smem attribute (%d B): %s", fn, smem, cudaGetErrorString(e));
    if (CG == 2) launch_pdl(conv_halo2f_kernel, dim3(grid), dim3(256), smem, s, p);
    else launch_pdl(conv_halof_kernel, dim3(grid), dim3(256), smem, s, p);
    return after_launch(fn);
  }
  if (d->balance) {
    CNM_REQUIRE(!sweep && !tf32 && TH == 16 && BN <= 128, CNM_ERR_UNSUPPORTED_SHAPE,
                "%s: balance needs a 2-byte dtype, 16x16 blocks and a column tile of at most 128 (got %d)", fn, BN);
    CNM_REQUIRE(d->splitk_ws && (((uintptr_t)d->splitk_ws) & 127) == 0, CNM_ERR_INVALID_ARGUMENT,
                "%s: balance needs a 128B-aligned, zero-initialised splitk_ws of cnm_conv_balance_ws_bytes() bytes", fn);
    p.bal_flags = reinterpret_cast<int*>(d->splitk_ws);
    p.bal_slots = d->splitk_ws + BALANCE_FLAG_BYTES / 4;
    CNM_REQUIRE(p.a_stages * p.slab_bytes + p.b_stages * b_tile >= 2 * 128 * BN * 4, CNM_ERR_UNSUPPORTED_SHAPE,
                "%s: operand rings of %d B cannot stage a contributor slot", fn, p.a_stages * p.slab_bytes + p.b_stages * b_tile);
    // one equal range of slab groups per CTA (pair); at least 4 groups each
    const long long T = (long long)total * num_groups;
    const int units = T / 4 < 148 / CG ? (int)(T / 4 > 0 ? T / 4 : 1) : 148 / CG;
    static DynSmemLimit b64, b128, b2_64, b2_128;
    cudaError_t e;
    if (CG == 2) e = BN == 128 ? b2_128.ensure(conv_halo2b_kernel<128>, smem) : b2_64.ensure(conv_halo2b_kernel<64>, smem);
    else e = BN == 128 ? b128.ensure(conv_halob_kernel<128>, smem) : b64.ensure(conv_halob_kernel<64>, smem);
    if (e != cudaSuccess) return fail(CNM_ERR_CUDA, "%s: smem attribute (%d B): %s", fn, smem, cudaGetErrorString(e));
    if (CG == 2) {
      if (BN == 128) launch_pdl(conv_halo2b_kernel<128>, dim3(CG * units), dim3(256), smem, s, p);
      else launch_pdl(conv_halo2b_kernel<64>, dim3(CG * units), dim3(256), smem, s, p);
    } else {
      if (BN == 128) launch_pdl(conv_halob_kernel<128>, dim3(units), dim3(256), smem, s, p);
      else launch_pdl(conv_halob_kernel<64>, dim3(units), dim3(256), smem, s, p);
    }
    return after_launch(fn);
  }
  if (sweep) {
    static DynSmemLimit limit64, limit128;
    cudaError_t e = BN == 128 ? limit128.ensure(conv_halo2_sweep_kernel<128>, smem) : limit64.ensure(conv_halo2_sweep_kernel<64>, smem);
    if (e != cudaSuccess) return fail(CNM_ERR_CUDA, "%s: smem attribute (%d B): %s", fn, smem, cudaGetErrorString(e));
    if (BN == 128) launch_pdl(conv_halo2_sweep_kernel<128>, dim3(grid), dim3(512), smem, s, p);
    else launch_pdl(conv_halo2_sweep_kernel<64>, dim3(grid), dim3(512), smem, s, p);
    return after_launch(fn);
  }
  if (TH == 8) {
    static DynSmemLimit l64, l128, l256;
    cudaError_t e = BN == 256 ? l256.ensure(conv_halo2h_kernel<256>, smem)
                              : (BN == 128 ? l128.ensure(conv_halo2h_kernel<128>, smem) : l64.ensure(conv_halo2h_kernel<64>, smem));
    if (e != cudaSuccess) return fail(CNM_ERR_CUDA, "%s: smem attribute (%d B): %s", fn, smem, cudaGetErrorString(e));
    if (BN == 256) launch_pdl(conv_halo2h_kernel<256>, dim3(grid), dim3(256), smem, s, p);
    else if (BN == 128) launch_pdl(conv_halo2h_kernel<128>, dim3(grid), dim3(256), smem, s, p);
    else launch_pdl(conv_halo2h_kernel<64>, dim3(grid), dim3(256), smem, s, p);
    return after_launch(fn);
  }
  if (BN == 256) return tf32 ? launch_halo<256, true, CG>(p, grid, smem, s) : launch_halo<256, false, CG>(p, grid, smem, s);
  if (BN == 128) return tf32 ? launch_halo<128, true, CG>(p, grid, smem, s) : launch_halo<128, false, CG>(p, grid, smem, s);
  return tf32 ? launch_halo<64, true, CG>(p, grid, smem, s) : launch_halo<64, false, CG>(p, grid, smem, s);
}

extern "C" int cnm_conv2d_halo(const cnm_conv_desc* d, const cnm_conv_src* srcs, const int32_t* map_info,
                               const cnm_conv_group* groups, int num_groups, const void* wp, const float* bias, void* out,
                               void* stream) {
  return conv2d_halo<1>(d, srcs, map_info, groups, num_groups, wp, bias, out, stream);
}

extern "C" int cnm_conv2d_halo2(const cnm_conv_desc* d, const cnm_conv_src* srcs, const int32_t* map_info,
                                const cnm_conv_group* groups, int num_groups, const void* wp, const float* bias, void* out,
                                void* stream) {
  return conv2d_halo<2>(d, srcs, map_info, groups, num_groups, wp, bias, out, stream);
}

extern "C" int64_t cnm_sweep_scratch_bytes(int D, int k) {
  if (D <= 0 || k < 1 || k > 9 || !(k & 1)) return 0;
  return sweep_scratch_bytes(D, k);
}

extern "C" int cnm_conv2d_halo2_sweep(const cnm_conv_desc* d, const cnm_conv_src* srcs, const int32_t* map_info,
                                      const cnm_conv_group* groups, int num_groups, const void* wp, const float* bias,
                                      void* out, const cnm_sweep_fuse* sweep, void* stream) {
  CNM_REQUIRE(sweep, CNM_ERR_INVALID_ARGUMENT, "cnm_conv2d_halo2_sweep: null sweep descriptor");
  return conv2d_halo<2>(d, srcs, map_info, groups, num_groups, wp, bias, out, stream, sweep);
}
