// K2: convolution + folded BatchNorm + ReLU as a persistent, warp-specialised implicit GEMM on
// the 5th-generation tensor cores (tcgen05.mma, accumulators in TMEM, operands staged by TMA).
//
// Reference semantics: nn.Conv2d + nn.BatchNorm2d(eval) + nn.ReLU and the torch.cat of the
// inputs (depthnet/depthNet_model.py:19-112, 235-261, 335-365).
//
// GEMM view      D[M = N*Ho*Wo pixels, Cout] = sum over K-steps of A_step[M, chunk] * W_step[Cout, chunk]^T
// K-step         = (source tensor, parity sub-image, tap shift (dx,dy), channel chunk c0); listed by
//                  the host, so concatenated inputs, taps and stride 2 are all the same loop.
// A operand      one 4-D TMA box {chunk ch, TW, TH, NB} (TW*TH*NB = 128 pixels) of the NHWC source at
//                  the shifted tile origin; out-of-image pixels are zero-filled by TMA (= conv padding).
//                  The box lands in smem as 128 rows x 128 B, 128B-swizzled = a K-major UMMA tile.
// B operand      2-D TMA box {chunk, BN} of the packed weights [Cout, K] (BN folded), same layout.
// MMA            tcgen05.mma cta_group::1, M = 128, N = BN, K = 32 bytes per instruction
//                  (kind::f16: 16 x bf16 / fp16, kind::tf32: 8 x tf32, kind::f8f6f4: 32 x e4m3 for the
//                  correction rows of CNM_F16F8); fp32 accumulators in TMEM, two accumulator stages
//                  so the epilogue of tile i overlaps the main loop of tile i+1.
// Epilogue       4 warps: tcgen05.ld -> * 2^-acc_exp + bias -> ReLU -> pack (bf16 / tf32 / hi+lo rows /
//                  fp16 + fp8 rows) -> swizzled smem -> TMA store into a channel slice of the NHWC
//                  output (TMA clips partial tiles).  Optional: the nine tap dot products of a fused
//                  1-channel head (Cout = 64), or raw accumulators to the split-K workspace.
// Split-K        desc.ksplit > 1: tile index gains a K-range dimension; splitk_finish_kernel sums the
//                  ranges in a fixed order and does the epilogue work.
// Multicast       CL = 2 / 4 (cnm_conv_desc.cluster): a cluster of CL CTAs works on CL pixel tiles of the SAME column
//                  tile and K range; every CTA loads 1/CL of each weight tile and multicasts it to all of them, so the
//                  L2 -> SM bytes per K-step drop from 16 + BN/8 KB to 16 + BN/(8 CL) KB.  These layers are bound by
//                  operand delivery (bytes in flight / loaded L2 latency), not by MMA time.
// Warp roles     warp 0 = TMA producer (one lane), warp 1 = MMA issuer (one lane), warp 2 = TMEM
//                  allocator, warp 3 idle, warps 4-7 = epilogue (TMEM lane quarter = warp - 4).
#include <cstdlib>
#include <cstring>
#include <mutex>

#include "tc_common.cuh"

namespace cnm {

constexpr int TC_MAX_SRCS = 3;
constexpr int TC_A_BYTES = 128 * 128;  // 128 pixels x 128 B
constexpr int TC_THREADS = 256;

struct alignas(64) TcParams {
  CUtensorMap a_maps[TC_MAX_SRCS * 4];
  CUtensorMap b_map;
  CUtensorMap out_map;
  const cnm_conv_step* steps;
  const float* bias;
  int num_steps;
  int tiles_x, tiles_y, tiles_b, tiles_n;
  int TW, TH, NB;
  int relu;
  int c_out_offset;
  int mode;         // EPI_PLAIN / EPI_BF16X3 / EPI_F16F8: the split modes store two 128-byte rows per 64-column chunk
  float acc_scale;  // 2^-acc_exp
  unsigned long long* range_ctr;  // debug counter of out-of-fp16-range outputs (NULL: off)
  const float* head_w;   // fused 1-channel 3x3 head (BN == Cout == 64 only): [9][64] weights, or NULL
  float* head_partial;   // [N, Ho, Wo, 12] per-pixel tap dot products
  int N, Ho, Wo;
  int ksplit;       // > 1: the K loop of every output tile is cut into ksplit ranges run by different CTAs
  float* ws;        // split-K workspace [ksplit][N*Ho*Wo][Cout] fp32 (raw accumulators; splitk_finish_kernel sums them)
  int cout;
  long long* dbg;   // CNM_CONV_TIMELINE=1: clock64 stamps of CTA 0 (tools/conv_timeline.py), else NULL
  // balanced mode (cnm_conv_desc.balance, BAL instantiations): see TcWalk
  const float* side_disp;   // 1-channel side input added in the epilogue (cnm_conv_desc.side_disp / side_w), or NULL
  const float* side_w;      // [9][Cout]; read through L1 (the layers that matter run the row-halo kernels)
  float* bal_slots;  // [gridDim.x][128][BN] fp32 raw accumulators of the contributor pieces
  int* bal_flags;    // [gridDim.x] 1 = the CTA's contributor piece is in its slot (reset by the tile's finishing CTA)
};

// Work of one CTA as a sequence of segments (tile, K-steps [k0, k1)).
//   static (BAL = false): work items first_q, first_q + stride, ... of the (pixel-tile group, column tile, split-K part)
//     index space; a split-K part is a K range whose raw accumulators go to the workspace for splitk_finish_kernel.
//   balanced (BAL = true): the launch's K-steps -- tiles x num_steps, tile-major -- are cut into gridDim.x equal contiguous
//     ranges, one per CTA (stream-K).  A CTA whose range starts inside a tile runs that CONTRIBUTOR piece first and parks
//     the raw accumulators in its slot; the CTA that owns the tile's first K-step (the HEAD piece, the last segment of its
//     range) adds the slots of the following CTAs in index order and runs the normal epilogue.  Contributors never wait,
//     and they finish their piece before the head CTA finishes its own, so nobody idles: a layer with 120 tiles keeps all
//     148 SMs busy for 0.81 tile times instead of 120 SMs for one.  The summation order is fixed, results are
//     deterministic (but differ in the last bits from the static schedule, which adds the K-steps in one chain).
template <bool BAL>
struct TcWalk {
  long long pos, end;   // balanced: position in the global K-step space
  int q, stride, total; // static: work item, stride, number of items
  int S, ksplit;
  int tile, part, k0, k1;
  __device__ __forceinline__ int kbeg(int pt) const { return pt >= ksplit ? S : ((S / 2) * pt / ksplit) * 2; }
  __device__ __forceinline__ static long long cut(long long T, int c, int G) { return T * c / G; }
  __device__ __forceinline__ TcWalk(const TcParams& p, int tiles, int first_q, int q_stride) {
    S = p.num_steps; ksplit = p.ksplit;
    if (BAL) {
      const long long T = (long long)tiles * S;
      pos = cut(T, (int)blockIdx.x, (int)gridDim.x); end = cut(T, (int)blockIdx.x + 1, (int)gridDim.x);
    } else {
      q = first_q; stride = q_stride; total = tiles * ksplit;
    }
  }
  __device__ __forceinline__ bool next() {
    if (BAL) {
      if (pos >= end) return false;
      tile = (int)(pos / S); part = 0;
      k0 = (int)(pos - (long long)tile * S);
      const long long left = end - pos;
      k1 = (left < (long long)(S - k0)) ? k0 + (int)left : S;
      pos += k1 - k0;
      return true;
    }
    if (q >= total) return false;
    part = q % ksplit; tile = q / ksplit;
    k0 = kbeg(part); k1 = kbeg(part + 1);
    q += stride;
    return true;
  }
  // first K-step of the segment after the current one (descriptor prefetch; any valid index if there is none)
  __device__ __forceinline__ int next_first() const { return BAL ? 0 : kbeg(q % ksplit) % S; }
};

template <int BN, int STAGES, bool TF32>
struct TcSmem {
  static constexpr int B_BYTES = BN * 128;
  static constexpr int STAGE_BYTES = TC_A_BYTES + B_BYTES;
  static constexpr int OUT_BYTES = 2 * TC_A_BYTES;  // double-buffered 128 x 128 B staging
  static constexpr int OFF_OUT = STAGES * STAGE_BYTES;
  static constexpr int OFF_BIAS = OFF_OUT + OUT_BYTES;
  static constexpr int OFF_HEAD = OFF_BIAS + BN * 4;                     // fused-head weights of the N tile [9][BN]
  static constexpr bool HEAD_IN_SMEM = BN == 64;                         // wider tiles fill shared memory: weights from L1
  static constexpr int OFF_BAR = OFF_HEAD + (HEAD_IN_SMEM ? 9 * BN * 4 : 0);
  static constexpr int NUM_BARS = 2 * STAGES + 5;   // + the bulk-copy barrier of the balanced schedule
  static constexpr int OFF_TMEM = OFF_BAR + NUM_BARS * 8;
  static constexpr int TOTAL = OFF_TMEM + 16 + 1024;  // + slack for manual 1024 B alignment
};

template <int BN, int STAGES, bool TF32, int CL, bool BAL = false, bool SIDE = false>
__device__ __forceinline__ void conv_tc_body(const TcParams& p) {
  static_assert(!SIDE || (!TF32 && CL == 1 && !BAL), "side disparity input: plain 2-byte instantiations");
  static_assert(!BAL || CL == 1, "balanced schedule: single CTAs");
  using L = TcSmem<BN, STAGES, TF32>;
  const bool stamp = p.dbg != nullptr && blockIdx.x == 0;
  if (stamp && threadIdx.x == 0) p.dbg[0] = clock64();
  pdl_trigger();
  constexpr uint16_t CL_MASK = (uint16_t)((1u << CL) - 1u);
  const uint32_t rank = CL > 1 ? cluster_ctarank() : 0u;
  // work item q of a cluster = (pixel-tile group, column tile, K range); CTA `rank` takes pixel tile CL * group + rank
  // (a tile index beyond the last one is a dummy: its TMA boxes lie outside the tensors -- zero-filled loads, clipped
  // stores -- but it walks the same K-steps, because the cluster shares the weight ring)
  const int first_q = (int)blockIdx.x / CL, q_stride = (int)gridDim.x / CL;
  constexpr int COLS_PER_CHUNK = TF32 ? 32 : 64;  // accumulator columns per 128-byte output row
  constexpr int CHUNK_CH = TF32 ? 32 : 64;        // channels per K-step
  constexpr int NUM_CHUNKS = BN / COLS_PER_CHUNK;
  constexpr uint32_t TMEM_COLS = 2 * BN;          // 128 / 256 / 512: powers of two >= 32
  constexpr uint32_t IDESC_SHAPE = (1u << 4) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  constexpr uint32_t IDESC = IDESC_SHAPE | ((TF32 ? 2u : 1u) << 7) | ((TF32 ? 2u : 1u) << 10);
  // CNM_F16F8: operand formats 0 = e4m3 (kind::f8f6f4) on the first half of the K-steps (the correction rows, accumulated
  // while the sum is still tiny) and 0 = fp16 (kind::f16) on the second half
  const bool mixed = !TF32 && p.mode == EPI_F16F8;
  const uint32_t idesc = mixed ? IDESC_SHAPE : IDESC;

  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* base_ptr = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t bar0 = base + L::OFF_BAR;
  auto full_bar = [&](int s) { return bar0 + 8u * s; };
  auto empty_bar = [&](int s) { return bar0 + 8u * (STAGES + s); };
  auto tfull_bar = [&](int a) { return bar0 + 8u * (2 * STAGES + a); };
  auto tempty_bar = [&](int a) { return bar0 + 8u * (2 * STAGES + 2 + a); };
  const uint32_t gather_bar = bar0 + 8u * (2 * STAGES + 4);
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(base_ptr + L::OFF_TMEM);
  float* bias_s = reinterpret_cast<float*>(base_ptr + L::OFF_BIAS);
  float* head_s = reinterpret_cast<float*>(base_ptr + L::OFF_HEAD);
  const bool fused_head = !TF32 && p.head_partial != nullptr;
  constexpr bool side = SIDE;   // own instantiations: the extra epilogue state costs the others registers

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);  // warp-uniform for the compiler
  const int lane = threadIdx.x & 31;
  const int m_groups = (p.tiles_x * p.tiles_y * p.tiles_b + CL - 1) / CL;
  const int num_tiles = m_groups * p.tiles_n;   // (pixel-tile group, column tile) pairs; x ksplit work items (TcWalk)

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.b_map);
    tma_prefetch_desc(&p.out_map);
    tma_prefetch_desc(&p.a_maps[0]);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), CL); }
    for (int a = 0; a < 2; ++a) { mbar_init(tfull_bar(a), 1); mbar_init(tempty_bar(a), 128); }
    mbar_init(gather_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(base + L::OFF_TMEM),
                 "r"(TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  if (CL > 1) cluster_sync_all();  // every CTA's barriers exist before a peer's multicast / commit can reach them
  else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (stamp && threadIdx.x == 0) p.dbg[1] = clock64();
  pdl_wait();  // everything above overlapped the previous kernel's tail; its outputs are visible from here on
  if (stamp && threadIdx.x == 0) p.dbg[2] = clock64();

  if (warp == 0) {
    // ================= TMA producer =================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      TcWalk<BAL> w(p, num_tiles, first_q, q_stride);
      bool more = w.next(), first_seg = true;
      int4 st_next = __ldg(reinterpret_cast<const int4*>(p.steps + (more ? w.k0 : 0)));
      int c0_next = __ldg(&p.steps[more ? w.k0 : 0].c0);
      for (; more; more = w.next(), first_seg = false) {
        const int nt = w.tile % p.tiles_n;
        int m = (w.tile / p.tiles_n) * CL + (int)rank;
        const int x0 = (m % p.tiles_x) * p.TW;
        m /= p.tiles_x;
        const int y0 = (m % p.tiles_y) * p.TH;
        const int n0 = (m / p.tiles_y) * p.NB;
        const int k_begin = w.k0, k_end = w.k1, k_after = w.next_first();
        for (int ks = k_begin; ks < k_end; ++ks) {
          const int4 st = st_next;  // src, sub, dx, dy
          const int c0 = c0_next;
          {  // prefetch the next step descriptor: its L2 latency overlaps the barrier wait below
            const int kn = (ks + 1 == k_end) ? k_after : ks + 1;
            st_next = __ldg(reinterpret_cast<const int4*>(p.steps + kn));
            c0_next = __ldg(&p.steps[kn].c0);
          }
          mbar_wait(empty_bar(stage), phase ^ 1);
          mbar_expect_tx(full_bar(stage), L::STAGE_BYTES);
          const uint32_t sa = base + stage * L::STAGE_BYTES;
          tma_load_4d(sa, &p.a_maps[st.x * 4 + st.y], full_bar(stage), c0, x0 + st.z, y0 + st.w, n0);
          if (CL > 1)  // this CTA's 1/CL of the weight rows, to every CTA of the cluster
            tma_load_2d_mc(sa + TC_A_BYTES + rank * (uint32_t)(L::B_BYTES / CL), &p.b_map, full_bar(stage), ks * CHUNK_CH,
                           nt * BN + (int)rank * (BN / CL), CL_MASK);
          else
            tma_load_2d(sa + TC_A_BYTES, &p.b_map, full_bar(stage), ks * CHUNK_CH, nt * BN);
          if (stamp && first_seg && ks == k_begin) p.dbg[3] = clock64();
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer =================
    // The whole warp walks the loop and one elected lane issues (warp-uniform control flow: descriptors stay in
    // uniform registers, the tcgen05 instructions are not wrapped in per-lane election loops).
    {
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      TcWalk<BAL> w(p, num_tiles, first_q, q_stride);
      for (; w.next(); ++it) {
        const int as = it & 1;
        const uint32_t aphase = (it >> 1) & 1;
        mbar_wait(tempty_bar(as), aphase ^ 1);  // epilogue has drained this accumulator stage
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)(as * BN);
        const int k_begin = w.k0, k_end = w.k1;
        for (int ks = k_begin; ks < k_end; ++ks) {
          mbar_wait(full_bar(stage), phase);
          if (stamp && lane == 0 && it == 0 && ks == k_begin) p.dbg[4] = clock64();
          tc_fence_after();
          if (elect_one()) {
            const uint32_t sa = base + stage * L::STAGE_BYTES;
            const uint64_t adesc = umma_desc(sa), bdesc = umma_desc(sa + TC_A_BYTES);
            if (mixed && ks < (p.num_steps >> 1)) {
#pragma unroll
              for (int k = 0; k < 4; ++k)
                tc_mma_f8(tmem_d, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (ks != k_begin || k) ? 1u : 0u);
            } else {
#pragma unroll
              for (int k = 0; k < 4; ++k)  // 4 x 32 B along K inside the 128 B swizzle row
                tc_mma<TF32>(tmem_d, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc,
                             (ks != k_begin || k) ? 1u : 0u);
            }
            // frees the smem stage when these MMAs retire (CL > 1: in every CTA -- the peers write into it)
            if (CL > 1) tc_commit_mc(empty_bar(stage), CL_MASK);
            else tc_commit(empty_bar(stage));
            if (ks == k_end - 1) tc_commit(tfull_bar(as));  // accumulator complete
          }
          __syncwarp();
          if (stamp && lane == 0 && it == 0 && ks == k_end - 1) p.dbg[5] = clock64();
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp >= 4) {
    // ================= epilogue =================
    const int q = warp - 4;               // TMEM lane quarter
    const int row = q * 32 + lane;        // pixel row of the tile handled by this thread
    const int et = threadIdx.x - 128;     // 0..127
    const bool leader = (et == 0);
    int it = 0;
    uint32_t chunk_counter = 0;
    int bias_tile = -1;
    uint32_t gather_phase = 0;

    TcWalk<BAL> w(p, num_tiles, first_q, q_stride);
    for (; w.next(); ++it) {
      const int nt = w.tile % p.tiles_n;
      int m = (w.tile / p.tiles_n) * CL + (int)rank;
      const int x0 = (m % p.tiles_x) * p.TW;
      m /= p.tiles_x;
      const int y0 = (m % p.tiles_y) * p.TH;
      const int n0 = (m / p.tiles_y) * p.NB;
      const int as = it & 1;
      const uint32_t aphase = (it >> 1) & 1;
      if (BAL && w.k0 > 0) {
        // contributor piece: raw accumulators -> this CTA's slot, then the flag
        mbar_wait(tfull_bar(as), aphase);
        tc_fence_after();
        balance_park(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * BN), BN,
                     p.bal_slots + (size_t)blockIdx.x * (BALANCE_SLOT_BYTES / 4), row, p.bal_flags + blockIdx.x, leader);
        tc_fence_before();
        mbar_arrive(tempty_bar(as));
        continue;
      }
      // head piece of a balanced tile: CTAs blockIdx.x + 1 ... + ncontrib hold the rest of its K range
      int ncontrib = 0;
      if (BAL && w.k1 < w.S) {
        const long long T = (long long)num_tiles * w.S, tile_end = (long long)(w.tile + 1) * w.S;
        while (TcWalk<BAL>::cut(T, (int)blockIdx.x + 1 + ncontrib, (int)gridDim.x) < tile_end) ++ncontrib;
      }
      if (!BAL && p.ksplit > 1) {
        // split-K: park the raw accumulators of this K range in the workspace; bias / ReLU / packing happen in
        // splitk_finish_kernel once all ranges are in
        mbar_wait(tfull_bar(as), aphase);
        tc_fence_after();
        const int px = x0 + row % p.TW, py = y0 + (row / p.TW) % p.TH, pn = n0 + row / (p.TW * p.TH);
        const bool inside = px < p.Wo && py < p.Ho && pn < p.N;
        const size_t mtot = (size_t)p.N * p.Ho * p.Wo;
        float* wrow = p.ws + ((size_t)w.part * mtot + ((size_t)pn * p.Ho + py) * p.Wo + px) * p.cout + nt * BN;
        const uint32_t tmem_row = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * BN);
#pragma unroll 1
        for (int c = 0; c < BN; c += 32) {
          uint32_t v[32];
          tmem_ld32(tmem_row + c, v);
          tmem_ld_wait();
          if (inside) {
#pragma unroll
            for (int j = 0; j < 32; j += 4)
              *reinterpret_cast<uint4*>(wrow + c + j) = make_uint4(v[j], v[j + 1], v[j + 2], v[j + 3]);
          }
        }
        tc_fence_before();
        mbar_arrive(tempty_bar(as));
        continue;
      }
      if (bias_tile != nt) {
        epi_bar();  // previous tile's bias reads are done
        for (int i = et; i < BN; i += 128) bias_s[i] = p.bias ? __ldg(p.bias + nt * BN + i) : 0.0f;
        if (fused_head && L::HEAD_IN_SMEM)
          for (int i = et; i < 9 * BN; i += 128) head_s[i] = __ldg(p.head_w + (i / BN) * p.cout + nt * BN + i % BN);
        bias_tile = nt;
        epi_bar();  // epi_finalize reads bias_s right after the TMEM load, before any other barrier
      }
      mbar_wait(tfull_bar(as), aphase);
      if (stamp && et == 0 && it == 0) p.dbg[6] = clock64();
      tc_fence_after();
      const uint32_t tmem_row = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * BN);
      if (BAL && ncontrib)   // the tile's other pieces -> added to the accumulator in TMEM (the rings are idle: last segment)
        balance_gather(tmem_row, BN, p.bal_slots, p.bal_flags, (int)blockIdx.x + 1, 1, ncontrib, base_ptr, gather_bar,
                       gather_phase, row, leader);
      if (stamp && et == 0 && it == 0) p.dbg[9] = clock64();
      float sd[9];
      if (side) {
        const int px = x0 + row % p.TW, py = y0 + (row / p.TW) % p.TH, pn = n0 + row / (p.TW * p.TH);
        side_disp_taps(p.side_disp, pn < p.N ? pn : p.N - 1, py, px, p.Ho, p.Wo, sd);
      }
#pragma unroll 1
      for (int ch = 0; ch < NUM_CHUNKS; ++ch) {
        uint32_t v[COLS_PER_CHUNK];
        tmem_ld32(tmem_row + ch * COLS_PER_CHUNK, v);
        if (!TF32) tmem_ld32(tmem_row + ch * COLS_PER_CHUNK + 32, v + 32);
        tmem_ld_wait();
        if (side) epi_finalize_side<COLS_PER_CHUNK>(v, bias_s + ch * COLS_PER_CHUNK, p.relu, p.acc_scale, p.range_ctr, sd,
                                                    p.side_w + nt * BN + ch * COLS_PER_CHUNK, p.cout);
        else epi_finalize<COLS_PER_CHUNK>(v, bias_s + ch * COLS_PER_CHUNK, p.relu, p.acc_scale, p.range_ctr);
        if (!TF32 && fused_head) {
          const int px = x0 + row % p.TW, py = y0 + (row / p.TW) % p.TH, pn = n0 + row / (p.TW * p.TH);
          if (px < p.Wo && py < p.Ho && pn < p.N) {
            const size_t slice = (size_t)(nt * NUM_CHUNKS + ch) * p.N * p.Ho * p.Wo;  // one slice per 64-channel chunk
            float* dst = p.head_partial + (slice + ((size_t)pn * p.Ho + py) * p.Wo + px) * HEAD_PARTIAL_STRIDE;
            if (L::HEAD_IN_SMEM) epi_head_partial(v, head_s + ch * 64, BN, dst);
            else epi_head_partial(v, p.head_w + nt * BN + ch * 64, p.cout, dst);
          }
        }
        for (int part = 0; part < (p.mode ? 2 : 1); ++part) {
          const uint32_t buf = chunk_counter & 1;
          uint8_t* stg = base_ptr + L::OFF_OUT + buf * TC_A_BYTES;
          // the TMA store that last read this staging buffer (two stores ago) must have finished reading
          if (leader) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
          epi_bar();
          epi_store_row<TF32>(v, stg + row * 128, row, p.mode, part);
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> async proxy
          epi_bar();
          if (leader) {
            tma_store_4d(&p.out_map, smem_u32(stg), p.c_out_offset + (p.mode ? 2 * (nt * BN + ch * CHUNK_CH) + part * 64 : nt * BN + ch * CHUNK_CH),
                                 x0, y0, n0);
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          }
          ++chunk_counter;
        }
      }
      tc_fence_before();
      mbar_arrive(tempty_bar(as));  // 128 arrivals release the accumulator stage
    }
    if (leader) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    if (stamp && et == 0) p.dbg[7] = clock64();
  }

  tc_fence_before();
  if (CL > 1) cluster_sync_all();  // a peer may still be writing this CTA's ring or arriving on its barriers
  else __syncthreads();
  if (stamp && threadIdx.x == 0) p.dbg[8] = clock64();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}

template <int BN, int STAGES, bool TF32>
__global__ void __launch_bounds__(TC_THREADS, 1) conv_tc_kernel(const __grid_constant__ TcParams p) {
  conv_tc_body<BN, STAGES, TF32, 1>(p);
}
// side disparity input in the epilogue (cnm_conv_desc.side_disp), 2-byte dtypes
template <int BN, int STAGES>
__global__ void __launch_bounds__(TC_THREADS, 1) conv_tc_side_kernel(const __grid_constant__ TcParams p) {
  conv_tc_body<BN, STAGES, false, 1, false, true>(p);
}
// balanced (stream-K) schedule, 2-byte dtypes
template <int BN, int STAGES>
__global__ void __launch_bounds__(TC_THREADS, 1) conv_tcb_kernel(const __grid_constant__ TcParams p) {
  conv_tc_body<BN, STAGES, false, 1, true>(p);
}
template <int BN, int STAGES>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(TC_THREADS, 1) conv_tc_mc2_kernel(const __grid_constant__ TcParams p) {
  conv_tc_body<BN, STAGES, false, 2>(p);
}
template <int BN, int STAGES>
__global__ void __cluster_dims__(4, 1, 1) __launch_bounds__(TC_THREADS, 1) conv_tc_mc4_kernel(const __grid_constant__ TcParams p) {
  conv_tc_body<BN, STAGES, false, 4>(p);
}

// Second half of a split-K convolution: sums the ksplit partial accumulators of every output element in a fixed order
// (deterministic), applies scale / bias / ReLU and writes the act tensor.  One thread per (pixel, 8 channels): 32-byte
// contiguous reads per thread and slice, a warp covers 1 KB.
template <typename T>
__global__ void __launch_bounds__(256)
    splitk_finish_kernel(const float* __restrict__ ws, int ksplit, size_t mtot, int cout, const float* __restrict__ bias,
                         float scale, int relu, uint8_t* __restrict__ out, size_t pix_bytes, size_t off_bytes,
                         unsigned long long* range_ctr) {
  pdl_trigger();
  pdl_wait();
  const int groups = cout >> 3;
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= mtot * groups) return;
  const int c = (int)(idx % groups) * 8;
  const size_t pixel = idx / groups;
  float v[8];
  bool over = false;
#pragma unroll
  for (int j = 0; j < 8; ++j) v[j] = 0.0f;
  for (int s = 0; s < ksplit; ++s) {
    const float4* src = reinterpret_cast<const float4*>(ws + ((size_t)s * mtot + pixel) * cout + c);
    const float4 a = __ldg(src), b = __ldg(src + 1);
    v[0] += a.x; v[1] += a.y; v[2] += a.z; v[3] += a.w;
    v[4] += b.x; v[5] += b.y; v[6] += b.z; v[7] += b.w;
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float t = fmaf(v[j], scale, bias ? __ldg(bias + c + j) : 0.0f);
    v[j] = relu ? fmaxf(t, 0.0f) : t;
    over |= !(fabsf(v[j]) <= 65504.0f);
  }
  if (range_ctr && over) atomicAdd(range_ctr, 1ULL);
  T* px = reinterpret_cast<T*>(out + pixel * pix_bytes + off_bytes) + ActTraits<T>::phys(c);
#pragma unroll
  for (int j = 0; j < 8; j += Vec16<T>::N) Vec16<T>::store(px + j, v + j);
}

// ---- host side ---------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(sym);
  });
  return fn;
}

int encode_map(CUtensorMap* map, int dtype, int rank, const void* ptr, const uint64_t* dims, const uint64_t* strides_bytes,
               const uint32_t* box, bool swizzle128) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return fail(CNM_ERR_CUDA, "cuTensorMapEncodeTiled is not available from the driver");
  cuuint64_t gd[5];
  cuuint64_t gs[5];
  cuuint32_t bx[5], es[5];
  for (int i = 0; i < rank; ++i) { gd[i] = dims[i]; bx[i] = box[i]; es[i] = 1; }
  for (int i = 0; i + 1 < rank; ++i) gs[i] = strides_bytes[i];
  CUresult r = enc(map, dtype == CNM_TF32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16,
                   (cuuint32_t)rank, const_cast<void*>(ptr), gd, gs, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(CNM_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
  return CNM_OK;
}

static int next_pow2(int v) {
  int p = 1;
  while (p < v) p <<= 1;
  return p;
}

template <int BN, int STAGES, bool TF32>
static int launch_tc(const TcParams& p, int grid, cudaStream_t s) {
  using L = TcSmem<BN, STAGES, TF32>;
  static DynSmemLimit limit;
  cudaError_t e = limit.ensure(conv_tc_kernel<BN, STAGES, TF32>, L::TOTAL);
  if (e != cudaSuccess) return fail(CNM_ERR_CUDA, "cnm_conv2d_tc: smem attribute: %s", cudaGetErrorString(e));
  launch_pdl(conv_tc_kernel<BN, STAGES, TF32>, dim3(grid), dim3(TC_THREADS), L::TOTAL, s, p);
  return after_launch("cnm_conv2d_tc");
}

template <int BN, int STAGES>
static int launch_tc_side(const TcParams& p, int grid, cudaStream_t s) {
  using L = TcSmem<BN, STAGES, false>;
  static DynSmemLimit limit;
  cudaError_t e = limit.ensure(conv_tc_side_kernel<BN, STAGES>, L::TOTAL);
  if (e != cudaSuccess) return fail(CNM_ERR_CUDA, "cnm_conv2d_tc: smem attribute: %s", cudaGetErrorString(e));
  launch_pdl(conv_tc_side_kernel<BN, STAGES>, dim3(grid), dim3(TC_THREADS), L::TOTAL, s, p);
  return after_launch("cnm_conv2d_tc");
}

template <int BN, int STAGES>
static int launch_tcb(const TcParams& p, int grid, cudaStream_t s) {
  using L = TcSmem<BN, STAGES, false>;
  static DynSmemLimit limit;
  cudaError_t e = limit.ensure(conv_tcb_kernel<BN, STAGES>, L::TOTAL);
  if (e != cudaSuccess) return fail(CNM_ERR_CUDA, "cnm_conv2d_tc: smem attribute: %s", cudaGetErrorString(e));
  launch_pdl(conv_tcb_kernel<BN, STAGES>, dim3(grid), dim3(TC_THREADS), L::TOTAL, s, p);
  return after_launch("cnm_conv2d_tc");
}

// multicast variant: grid = CL x (as many clusters as the device keeps resident, at most one per work item)
template <int BN, int STAGES, int CL>
static int launch_tc_mc(const TcParams& p, int total_q, cudaStream_t s) {
  using L = TcSmem<BN, STAGES, false>;
  static DynSmemLimit limit;
  static int max_clusters[64] = {0};
  auto kernel = CL == 2 ? conv_tc_mc2_kernel<BN, STAGES> : conv_tc_mc4_kernel<BN, STAGES>;
  cudaError_t e = limit.ensure(kernel, L::TOTAL);
  if (e != cudaSuccess) return fail(CNM_ERR_CUDA, "cnm_conv2d_tc: smem attribute: %s", cudaGetErrorString(e));
  int dev = 0;
  cudaGetDevice(&dev);
  int& mc = max_clusters[dev & 63];
  if (mc == 0) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(148 / CL * CL);
    cfg.blockDim = dim3(TC_THREADS);
    cfg.dynamicSmemBytes = L::TOTAL;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = CL; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    int n = 0;
    e = cudaOccupancyMaxActiveClusters(&n, kernel, &cfg);
    if (e != cudaSuccess || n <= 0) return fail(CNM_ERR_CUDA, "cnm_conv2d_tc: cluster occupancy query: %s", cudaGetErrorString(e));
    mc = n;
  }
  const int clusters = total_q < mc ? total_q : mc;
  launch_pdl(kernel, dim3(clusters * CL), dim3(TC_THREADS), L::TOTAL, s, p);
  return after_launch("cnm_conv2d_tc");
}

}  // namespace cnm

using namespace cnm;

static long long* g_tc_dbg_buf = nullptr;
static int g_tc_dbg_on = -1;

extern "C" int cnm_conv2d_tc(const cnm_conv_desc* d, const cnm_conv_src* srcs, const cnm_conv_step* steps,
                             const void* wp, const float* bias, void* out, void* stream) {
  int rc = check_conv_args("cnm_conv2d_tc", d, srcs, steps, wp, out);
  if (rc) return rc;
  const int eb = elem_bytes(d->dtype), ch = chunk_channels(d->dtype);
  CNM_REQUIRE(d->Cout % 64 == 0, CNM_ERR_UNSUPPORTED_SHAPE, "cnm_conv2d_tc: Cout=%d must be a multiple of 64", d->Cout);
  CNM_REQUIRE(!d->fold, CNM_ERR_UNSUPPORTED_SHAPE, "cnm_conv2d_tc: fold is a cnm_conv2d_halo / _halo2 feature");
  TcParams p;
  memset(&p, 0, sizeof(p));
  // pixel tile: TW x TH x NB = 128
  p.TW = d->Wo >= 16 ? 16 : 8;
  const int rows = 128 / p.TW;
  p.TH = next_pow2(d->Ho) < rows ? next_pow2(d->Ho) : rows;
  p.NB = 128 / (p.TW * p.TH);
  p.tiles_x = (d->Wo + p.TW - 1) / p.TW;
  p.tiles_y = (d->Ho + p.TH - 1) / p.TH;
  p.tiles_b = (d->N + p.NB - 1) / p.NB;
  const int m_tiles = p.tiles_x * p.tiles_y * p.tiles_b;
  // column-tile width: minimise waves x measured cycles per K-step.  The main loop is bound by operand delivery
  // (L2 -> smem bytes in flight / loaded latency), not by MMA issue: measured on B200 (profiles/r01_conv_tc_ncu.md)
  // a K-step costs ~800 cycles at BN=256 and ~575 at BN=128 and BN=64, so wide tiles win unless they cost a wave.
  int BN = 64;
  long best = -1;
  for (int cand = 256; cand >= 64; cand >>= 1) {
    if (d->Cout % cand) continue;
    const long tiles = (long)m_tiles * (d->Cout / cand);
    const long waves = (tiles + 147) / 148;
    const long cyc = cand == 256 ? 800 : (cand == 128 ? 575 : 560);
    const long cost = waves * cyc;
    if (best < 0 || cost < best) { best = cost; BN = cand; }
  }
  if (d->reserved == 64 || d->reserved == 128 || d->reserved == 256) {  // caller-selected tile (host autotuner)
    CNM_REQUIRE(d->Cout % d->reserved == 0, CNM_ERR_INVALID_ARGUMENT, "cnm_conv2d_tc: BN=%d does not divide Cout=%d",
                d->reserved, d->Cout);
    BN = d->reserved;
  }
  p.tiles_n = d->Cout / BN;
  p.steps = steps;
  p.bias = bias;
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
  p.N = d->N; p.Ho = d->Ho; p.Wo = d->Wo;
  p.ksplit = d->ksplit > 1 ? d->ksplit : 1;
  p.ws = d->splitk_ws;
  p.cout = d->Cout;
  if (g_tc_dbg_on < 0) {  // debugging aid: per-role time stamps of CTA 0 (cnm_debug_conv_timeline)
    const char* e = getenv("CNM_CONV_TIMELINE");
    g_tc_dbg_on = (e && e[0] == '1') ? 1 : 0;
    if (g_tc_dbg_on && (cudaMalloc(&g_tc_dbg_buf, 16 * sizeof(long long)) != cudaSuccess ||
                        cudaMemset(g_tc_dbg_buf, 0, 16 * sizeof(long long)) != cudaSuccess)) {
      g_tc_dbg_on = 0; g_tc_dbg_buf = nullptr; cudaGetLastError();
    }
  }
  p.dbg = g_tc_dbg_on == 1 ? g_tc_dbg_buf : nullptr;
  const bool balance = d->balance != 0;
  if (balance) {
    CNM_REQUIRE(p.ksplit == 1 && d->cluster <= 1 && d->dtype != CNM_TF32, CNM_ERR_UNSUPPORTED_SHAPE,
                "cnm_conv2d_tc: balance goes with ksplit <= 1, cluster <= 1 and a 2-byte dtype");
    CNM_REQUIRE(d->splitk_ws && (((uintptr_t)d->splitk_ws) & 127) == 0, CNM_ERR_INVALID_ARGUMENT,
                "cnm_conv2d_tc: balance needs a 128B-aligned, zero-initialised splitk_ws of cnm_conv_balance_ws_bytes() bytes");
    p.bal_flags = reinterpret_cast<int*>(d->splitk_ws);
    p.bal_slots = d->splitk_ws + BALANCE_FLAG_BYTES / 4;
  }
  if (p.ksplit > 1) {
    CNM_REQUIRE(d->splitk_ws && (((uintptr_t)d->splitk_ws) & 15) == 0, CNM_ERR_INVALID_ARGUMENT,
                "cnm_conv2d_tc: ksplit = %d needs a 16B-aligned splitk_ws", d->ksplit);
    CNM_REQUIRE(p.ksplit <= 16 && d->num_steps >= 4 * p.ksplit && !d->head_partial, CNM_ERR_UNSUPPORTED_SHAPE,
                "cnm_conv2d_tc: ksplit = %d does not fit %d K-steps (or a fused head)", d->ksplit, d->num_steps);
  }

  const uint32_t box[4] = {(uint32_t)ch, (uint32_t)p.TW, (uint32_t)p.TH, (uint32_t)p.NB};
  for (int i = 0; i < d->num_srcs; ++i) {
    const cnm_conv_src& s = srcs[i];
    const uint64_t pix = (uint64_t)s.C * eb;
    if (d->stride == 1) {
      const uint64_t dims[4] = {(uint64_t)s.C, (uint64_t)s.W, (uint64_t)s.H, (uint64_t)d->N};
      const uint64_t str[3] = {pix, pix * s.W, pix * s.W * s.H};
      rc = encode_map(&p.a_maps[i * 4], d->dtype, 4, s.ptr, dims, str, box);
      if (rc) return rc;
    } else {
      for (int sub = 0; sub < 4; ++sub) {
        const int py = sub >> 1, px = sub & 1;
        const int ws = (s.W - px + 1) / 2, hs = (s.H - py + 1) / 2;
        if (ws <= 0 || hs <= 0) continue;
        const uint64_t dims[4] = {(uint64_t)s.C, (uint64_t)ws, (uint64_t)hs, (uint64_t)d->N};
        const uint64_t str[3] = {2 * pix, 2 * pix * s.W, pix * s.W * s.H};
        const uint8_t* ptr = (const uint8_t*)s.ptr + ((uint64_t)py * s.W + px) * pix;
        rc = encode_map(&p.a_maps[i * 4 + sub], d->dtype, 4, ptr, dims, str, box);
        if (rc) return rc;
      }
    }
  }
  {
    const uint64_t K = (uint64_t)d->num_steps * ch;
    const uint64_t dims[2] = {K, (uint64_t)d->Cout};
    const uint64_t str[1] = {K * eb};
    const uint32_t bbox[2] = {(uint32_t)ch, (uint32_t)(BN / (d->cluster > 1 ? d->cluster : 1))};
    rc = encode_map(&p.b_map, d->dtype, 2, wp, dims, str, bbox);
    if (rc) return rc;
  }
  {
    const uint64_t pix = (uint64_t)d->c_out_total * eb;
    const uint64_t dims[4] = {(uint64_t)d->c_out_total, (uint64_t)d->Wo, (uint64_t)d->Ho, (uint64_t)d->N};
    const uint64_t str[3] = {pix, pix * d->Wo, pix * d->Wo * d->Ho};
    rc = encode_map(&p.out_map, d->dtype, 4, out, dims, str, box);
    if (rc) return rc;
  }
  const int total = m_tiles * p.tiles_n * p.ksplit;
  const int grid = total < 148 ? total : 148;
  cudaStream_t s = (cudaStream_t)stream;
  const bool tf32 = d->dtype == CNM_TF32;  // the 2-byte modes share the kind::f16 instantiations
  const int CL = d->cluster > 1 ? d->cluster : 1;
  if (d->side_disp) {
    CNM_REQUIRE(!balance && CL == 1, CNM_ERR_UNSUPPORTED_SHAPE, "cnm_conv2d_tc: no side disparity input on balanced / cluster launches");
    if (BN == 256) return launch_tc_side<256, 4>(p, grid, s);
    if (BN == 128) return launch_tc_side<128, 6>(p, grid, s);
    return launch_tc_side<64, 7>(p, grid, s);
  }
  if (balance) {
    // equal K ranges for all SMs; at least 8 K-steps per CTA
    const long long T = (long long)m_tiles * p.tiles_n * d->num_steps;
    const int g = T / 8 < 148 ? (int)(T / 8 > 0 ? T / 8 : 1) : 148;
    if (BN == 256) return launch_tcb<256, 4>(p, g, s);
    if (BN == 128) return launch_tcb<128, 6>(p, g, s);
    return launch_tcb<64, 7>(p, g, s);
  }
  if (CL > 1) {
    CNM_REQUIRE((CL == 2 || CL == 4) && !tf32 && BN >= 128, CNM_ERR_UNSUPPORTED_SHAPE,
                "cnm_conv2d_tc: cluster = %d needs a 2-byte dtype and a column tile of 128 or 256 (got %d)", d->cluster, BN);
    const int total_q = ((m_tiles + CL - 1) / CL) * p.tiles_n * p.ksplit;
    if (BN == 256) rc = CL == 2 ? launch_tc_mc<256, 4, 2>(p, total_q, s) : launch_tc_mc<256, 4, 4>(p, total_q, s);
    else rc = CL == 2 ? launch_tc_mc<128, 6, 2>(p, total_q, s) : launch_tc_mc<128, 6, 4>(p, total_q, s);
  } else if (BN == 256) rc = tf32 ? launch_tc<256, 4, true>(p, grid, s) : launch_tc<256, 4, false>(p, grid, s);
  else if (BN == 128) rc = tf32 ? launch_tc<128, 6, true>(p, grid, s) : launch_tc<128, 6, false>(p, grid, s);
  else rc = tf32 ? launch_tc<64, 7, true>(p, grid, s) : launch_tc<64, 7, false>(p, grid, s);
  if (rc || p.ksplit == 1) return rc;
  const size_t mtot = (size_t)d->N * d->Ho * d->Wo;
  const size_t out_vecs = mtot * (d->Cout / 8);
  const unsigned blocks = (unsigned)((out_vecs + 255) / 256);
  const size_t pix_bytes = (size_t)d->c_out_total * eb, off_bytes = (size_t)d->c_out_offset * eb;
#define CNM_FINISH(T) \
  launch_pdl(splitk_finish_kernel<T>, dim3(blocks), dim3(256), 0, s, p.ws, p.ksplit, mtot, d->Cout, bias, p.acc_scale, d->relu, (uint8_t*)out, pix_bytes, off_bytes, p.range_ctr)
  if (d->dtype == CNM_BF16) CNM_FINISH(__nv_bfloat16);
  else if (d->dtype == CNM_TF32) CNM_FINISH(float);
  else if (d->dtype == CNM_BF16X3) CNM_FINISH(bf16s);
  else CNM_FINISH(h16f8);
#undef CNM_FINISH
  return after_launch("cnm_conv2d_tc(split-K finish)");
}

extern "C" int64_t cnm_conv_balance_ws_bytes(void) { return BALANCE_WS_BYTES; }

// CNM_CONV_TIMELINE=1: the clock64 stamps CTA 0 of the last cnm_conv2d_tc launch left (synchronises the device)
extern "C" int cnm_debug_conv_timeline(long long* out12) {
  long long* out9 = out12;
  CNM_REQUIRE(out9, CNM_ERR_INVALID_ARGUMENT, "cnm_debug_conv_timeline: null pointer");
  CNM_REQUIRE(g_tc_dbg_buf, CNM_ERR_INVALID_ARGUMENT, "cnm_debug_conv_timeline: set CNM_CONV_TIMELINE=1 before the first convolution");
  cudaError_t e = cudaMemcpy(out9, g_tc_dbg_buf, 12 * sizeof(long long), cudaMemcpyDeviceToHost);
  if (e != cudaSuccess) return fail(CNM_ERR_CUDA, "cnm_debug_conv_timeline: %s", cudaGetErrorString(e));
  cudaMemset(g_tc_dbg_buf, 0, 16 * sizeof(long long));
  return CNM_OK;
}
