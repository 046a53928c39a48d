// K1: conv2d (3x3 s1/s2, 1x1) as an implicit GEMM on tcgen05 tensor cores.
//
//   M = B*Ho*Wo output pixels, N = Cout, K = sum over segments of taps*C.
//   A (activations, NHWC bf16) is never im2col'ed in memory: TMA tiled loads of shifted boxes land K-major,
//   128B-swizzled operand tiles in shared memory; TMA's out-of-bounds zero fill implements the conv padding, the
//   ragged image edge and the channel tail.  Stride-2 convs use the tensor map's element strides.  The
//   skip-connection torch.cat (legacy_unet.py:150) is a second K segment (another tensor map), never a copy.
//   B (weights) is a pre-packed [Cout][Ktot] K-major bf16 matrix, TMA-loaded as (64, BLOCK_N) boxes.
//   D accumulates in TMEM (fp32, 128 lanes x BLOCK_N columns); one elected thread issues tcgen05.mma.
//   Epilogue warps read TMEM with tcgen05.ld, add bias / per-sample time-embedding vector / residual, optionally
//   accumulate GroupNorm partial sums for the consumer norm, convert to bf16, stage the tile in 128B-swizzled
//   shared memory and write it with one TMA store per 64-channel slab (the store clips ragged tiles).
//
// Two kernels share this file's host code: the persistent per-tile kernel below (any tile box, stride 1/2, 1x1) and
// the rolling-row kernel of conv_rolling.cuh (stride-1 3x3 on rows >= 65 px, optional fused GroupNorm operand
// transform).  plan_conv() picks one from the shapes.  Reference op replaced: nn.Conv2d via ConvND.forward
// (src/nn/ops/convolution.py:53-54) and the adds around it in ResBlockND.forward (src/nn/blocks/residual.py:97-120).
#include <cstdarg>
#include <cstdlib>
#include <cstring>

#include "common.cuh"
#include "conv_epilogue.cuh"

namespace fm {

constexpr int kTileM = 128;
constexpr int kBlockK = 64;                      // 64 bf16 = 128 B = one swizzle row
constexpr int kABytes = kTileM * kBlockK * 2;    // 16 KB

struct alignas(64) ConvKernelParams {
  CUtensorMap src[FM_CONV_MAX_SEG];
  CUtensorMap wgt;
  CUtensorMap out;
  CUtensorMap out_up[3];            // out_upsample: the (dy, dx) = (0,1), (1,0), (1,1) phases of the 2x output
  CUtensorMap res;                  // residual, same geometry as the (un-upsampled) output (rolling kernel: TMA loads)
  int n_out;                        // 1, or 4 with out_upsample
  int nseg;
  int seg_c[FM_CONV_MAX_SEG];       // channels per segment
  int seg_taps[FM_CONV_MAX_SEG];    // 1 or 9
  int seg_koff[FM_CONV_MAX_SEG];    // K offset of the segment in the packed weight
  int Wt, Ht, Nt;                   // M-tile box, Wt*Ht*Nt == 128
  int tiles_w, tiles_h;             // tiles per image row / column
  int stride;
  int B, Ho, Wo, Cout;
  const float* bias;
  const float* addvec;
  int addvec_stride;
  const __nv_bfloat16* residual;
  float* gn_partial;                // [m_tiles*4][Cout/4][2] quad statistics, or null
  int log_wt, log_ht;               // Wt, Ht are powers of two
  int num_k_blocks;
  int dup_koff;                     // per-tile kernel, split-bf16 weights: K offset of the residual (lo) half of the
                                    // matrix - every A tile is multiplied by its hi AND lo weight tile (0: plain)
  int m_tiles, n_tiles;             // persistent kernel: real tile counts (M tiles padded to even for CTA pairs)
  // fused operand transform (XF kernels): per segment a*x+b table rows [B][.] and activation, or null
  const float* seg_na[FM_CONV_MAX_SEG];
  const float* seg_nb[FM_CONV_MAX_SEG];
  int seg_nstride[FM_CONV_MAX_SEG];
  int seg_nact[FM_CONV_MAX_SEG];
};

// MODE 0: one A tile per (tap, 64-channel block)  - any tile box, stride 1/2.
// MODE 1: "row" mode for stride-1 convs whose M tile is 128 consecutive pixels of one image row (Wt == 128):
//         ONE TMA load of the 130-pixel halo row per (kh, channel block) serves the three kw taps - the UMMA A
//         descriptor simply starts kw rows (kw*128 B) into the slot; the 128B swizzle is a function of the absolute
//         shared-memory address, so a row-shifted view of a TMA-written tile is still a valid SW128 K-major operand.
//         A traffic from L2 drops 3x (the per-tap form re-reads every input pixel nine times).
constexpr int kARowBytes = 128;
constexpr int kARowSlot = 17 * 1024;  // 130 rows * 128 B = 16640 B, rounded up to keep every slot 1024 B aligned
constexpr int kARowTx = 130 * 128;

// =================================================================================================================
// Persistent per-tile kernel.  One CTA (or CTA pair, CG = 2: tcgen05 cta_group::2, 256-row tiles, each CTA stages its
// own A rows and HALF of the weight tile) per SM walks a static round-robin list of output tiles; the TMA/MMA pipeline
// never drains between tiles and the accumulator is double-buffered in TMEM (2 x BLOCK_N columns), so the epilogue of
// tile i (TMEM -> registers -> bias/temb/residual/GroupNorm statistics -> bf16 -> TMA store) overlaps the main loop
// of tile i+1.  (A first one-tile-per-CTA version spent 31 % / 58 % (single / pair) of a K = 1152 conv in prologue,
// pipeline fill, epilogue and teardown.)
// =================================================================================================================
constexpr int kStageCols = 128;  // output columns staged (and TMA-stored) at a time

// MT = 2 (only with BLOCK_N = 128): every weight tile is used for TWO M tiles per CTA (accumulators side by side in
// TMEM, like one 256-column tile), halving the weight traffic per FLOP for the Cout = 128 layers, which are otherwise
// bound by L2->SM bandwidth (13.5 KB per 64-deep k-block per CTA against a 256-cycle MMA budget).
template <int BLOCK_N, int MODE, int CG, int MT>
struct PConvCfg {
  static constexpr int kBBytes = (BLOCK_N / CG) * kBlockK * 2;  // per CTA
  static constexpr int kASub = (MODE == 1) ? kARowSlot : kABytes;  // one M tile's A operand
  static constexpr int kASlot = kASub * MT;
  static constexpr int kATx = ((MODE == 1) ? kARowTx : kABytes) * MT;
  static constexpr int kAccCols = BLOCK_N * MT;                    // accumulator columns per buffer
  static constexpr int kHalfCols = (BLOCK_N < kStageCols) ? BLOCK_N : kStageCols;
  static constexpr int kOutBytes = kTileM * kHalfCols * 2;       // one staging buffer
  static constexpr int kOutBufs = 2;                              // store of half i overlaps the math of half i+1
  static constexpr int kTailBytes = 512 + BLOCK_N * 4;            // barriers + TMEM slot + per-tile bias vector
  static constexpr int kBudget = 227 * 1024 - 1024 - kTailBytes - kOutBufs * kOutBytes;
  static constexpr int kAStages = (MODE == 1) ? ((kBudget >= 3 * kASlot + 6 * kBBytes) ? 3 : 2)
                                              : ((kBudget / (kASlot + kBBytes)) > 8 ? 8 : (kBudget / (kASlot + kBBytes)));
  static constexpr int kBStagesRaw = (MODE == 1) ? (kBudget - kAStages * kASlot) / kBBytes : kAStages;
  static constexpr int kBStages = kBStagesRaw > 12 ? 12 : kBStagesRaw;
  static constexpr int kPipeBytes = kAStages * kASlot + kBStages * kBBytes;
  static constexpr int kNumBars = 2 * kAStages + 2 * kBStages + 4;
  static constexpr int kSmemBytes = 1024 + kPipeBytes + kOutBufs * kOutBytes + kTailBytes;
  static_assert(kAStages >= 2 && kBStages >= 3, "pipeline too shallow");
  static_assert(kNumBars * 8 + 8 <= 512, "barrier area");
  static_assert(kSmemBytes <= 227 * 1024, "shared memory budget");
};

constexpr int kEpiWarps = 8;                        // two warps per TMEM lane quadrant, each owning half the columns
constexpr int kEpiThreads = kEpiWarps * 32;
constexpr int kPConvThreads = 64 + kEpiThreads;     // + TMA producer warp + MMA warp

template <int BLOCK_N, int MODE, int CG, int MT>
__global__ void __launch_bounds__(kPConvThreads, 1)
conv_igemm_persistent_kernel(const __grid_constant__ ConvKernelParams p) {
  using Cfg = PConvCfg<BLOCK_N, MODE, CG, MT>;
  static_assert(MT == 1 || BLOCK_N == 128, "two M tiles per CTA only for 128-wide N tiles");
  const uint32_t cta_rank = (CG == 2) ? cluster_ctarank() : 0u;
  const bool is_leader = (cta_rank == 0);
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t smem_a0 = smem_base;
  const uint32_t smem_b0 = smem_base + Cfg::kAStages * Cfg::kASlot;
  const uint32_t smem_out = smem_base + Cfg::kPipeBytes;          // dedicated output staging (1024 B aligned)
  uint8_t* out_gen = smem_gen + Cfg::kPipeBytes;
  const uint32_t bar_base = smem_out + Cfg::kOutBufs * Cfg::kOutBytes;
  float* sbias = reinterpret_cast<float*>(out_gen + Cfg::kOutBufs * Cfg::kOutBytes + 512);  // [BLOCK_N]
  auto a_full = [&](int s) { return bar_base + 8u * s; };
  auto a_empty = [&](int s) { return bar_base + 8u * (Cfg::kAStages + s); };
  auto b_full = [&](int s) { return bar_base + 8u * (2 * Cfg::kAStages + s); };
  auto b_empty = [&](int s) { return bar_base + 8u * (2 * Cfg::kAStages + Cfg::kBStages + s); };
  auto t_full = [&](int b) { return bar_base + 8u * (2 * Cfg::kAStages + 2 * Cfg::kBStages + b); };
  auto t_empty = [&](int b) { return bar_base + 8u * (2 * Cfg::kAStages + 2 * Cfg::kBStages + 2 + b); };
  const uint32_t tmem_slot = bar_base + 8u * Cfg::kNumBars;
  volatile uint32_t* tmem_slot_gen =
      reinterpret_cast<volatile uint32_t*>(out_gen + Cfg::kOutBufs * Cfg::kOutBytes + 8 * Cfg::kNumBars);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  constexpr int kTmemCols = 2 * Cfg::kAccCols;  // double-buffered accumulator (<= 512)

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < p.nseg; ++s) tma_prefetch_desc(&p.src[s]);
    tma_prefetch_desc(&p.wgt);
    tma_prefetch_desc(&p.out);
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < 2 * Cfg::kAStages + 2 * Cfg::kBStages + 2; ++s) mbar_init(bar_base + 8u * s, 1);
      // accumulator-drained barriers: one arrival per epilogue warp of every CTA writing into this MMA's TMEM
      mbar_init(t_empty(0), kEpiWarps * CG);
      mbar_init(t_empty(1), kEpiWarps * CG);
      fence_barrier_init();
    }
    __syncwarp();
    if (CG == 2) tmem_alloc_pair<kTmemCols>(tmem_slot); else tmem_alloc<kTmemCols>(tmem_slot);
  }
  tc_fence_before();
  if (CG == 2) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_gen;
  // PDL: everything above is private set-up; from here on global memory of the preceding kernels is touched
  pdl_trigger();
  pdl_wait();

  // ---- static tile schedule: work unit = (M tile [pair], N tile), N fastest so neighbours share the A rows in L2
  const int units_m = p.m_tiles / (CG * MT);
  const int total_units = units_m * p.n_tiles;
  const int first_unit = (int)blockIdx.x / CG;
  const int unit_stride = (int)gridDim.x / CG;
  const int tiles_per_img_group = p.tiles_w * p.tiles_h;

  auto tile_coords = [&](int unit, int sub, int& m_tile, int& w0, int& h0, int& n0, int& ncol0) {
    const int um = unit / p.n_tiles;
    ncol0 = (unit - um * p.n_tiles) * BLOCK_N;
    m_tile = (um * CG + (int)cta_rank) * MT + sub;
    const int tn = m_tile / tiles_per_img_group;
    const int rem = m_tile - tn * tiles_per_img_group;
    const int th = rem / p.tiles_w;
    w0 = (rem - th * p.tiles_w) * p.Wt;
    h0 = th * p.Ht;
    n0 = tn * p.Nt;
  };

  if (warp == 0) {
    // ================= TMA producer (runs ahead across tiles) =================
    if (elect_one_sync()) {
      StageRing ra, rb;
      const uint32_t a_full_dst0 = (CG == 2) ? mapa_shared(a_full(0), 0) : a_full(0);  // pair: the leader's barriers
      const uint32_t b_full_dst0 = (CG == 2) ? mapa_shared(b_full(0), 0) : b_full(0);
      const bool arm = (CG == 1) || is_leader;  // pair: only the leader arms, with both CTAs' bytes
      for (int unit = first_unit; unit < total_units; unit += unit_stride) {
        int m_tile, w0, h0, n0, ncol0, w1 = 0, h1 = 0, n1 = 0;
        tile_coords(unit, 0, m_tile, w0, h0, n0, ncol0);
        if (MT == 2) tile_coords(unit, 1, m_tile, w1, h1, n1, ncol0);
        const int bcol = ncol0 + ((CG == 2) ? (int)cta_rank * (BLOCK_N / 2) : 0);
        for (int s = 0; s < p.nseg; ++s) {
          const int taps = p.seg_taps[s];
          const int C = p.seg_c[s];
          const int cblocks = (C + kBlockK - 1) / kBlockK;
          const int asteps = (MODE == 1) ? (taps == 9 ? 3 : 1) : taps;
          const int bsteps = (MODE == 1) ? (taps == 9 ? 3 : 1) : 1;
          int dh = (taps == 9) ? -1 : 0, dw = (MODE == 1 || taps == 9) ? -1 : 0;
          for (int as = 0; as < asteps; ++as) {
            for (int cb = 0; cb < cblocks; ++cb) {
              mbar_wait(a_empty(ra.idx), ra.phase ^ 1u);
              if (arm) mbar_expect_tx(a_full(ra.idx), CG * Cfg::kATx);
              const uint32_t bar = a_full_dst0 + 8u * ra.idx, dst = smem_a0 + ra.idx * Cfg::kASlot;
              if (CG == 2) {
                tma_load_4d_pair(&p.src[s], bar, dst, cb * kBlockK, w0 * p.stride + dw, h0 * p.stride + dh, n0);
                if (MT == 2)
                  tma_load_4d_pair(&p.src[s], bar, dst + Cfg::kASub, cb * kBlockK, w1 * p.stride + dw,
                                   h1 * p.stride + dh, n1);
              } else {
                tma_load_4d(&p.src[s], bar, dst, cb * kBlockK, w0 * p.stride + dw, h0 * p.stride + dh, n0);
                if (MT == 2)
                  tma_load_4d(&p.src[s], bar, dst + Cfg::kASub, cb * kBlockK, w1 * p.stride + dw, h1 * p.stride + dh,
                              n1);
              }
              ra.advance(Cfg::kAStages);
              // weight tiles of this A step: MODE 1 -> the three kw taps of row kh = as; MODE 0 -> tap = as
              int kcol = p.seg_koff[s] + ((MODE == 1) ? as * 3 : as) * C + cb * kBlockK;
              for (int bs = 0; bs < bsteps; ++bs, kcol += C) {
                // [lo,] hi tile: the small residual products enter the accumulator before the large ones of the step
                for (int kc = kcol + p.dup_koff; kc >= kcol; kc -= (p.dup_koff ? p.dup_koff : 1)) {
                  mbar_wait(b_empty(rb.idx), rb.phase ^ 1u);
                  if (arm) mbar_expect_tx(b_full(rb.idx), CG * Cfg::kBBytes);
                  if (CG == 2)
                    tma_load_2d_pair(&p.wgt, b_full_dst0 + 8u * rb.idx, smem_b0 + rb.idx * Cfg::kBBytes, kc, bcol);
                  else
                    tma_load_2d(&p.wgt, b_full_dst0 + 8u * rb.idx, smem_b0 + rb.idx * Cfg::kBBytes, kc, bcol);
                  rb.advance(Cfg::kBStages);
                }
              }
            }
            // next A step: MODE 1 walks kh (dh), MODE 0 walks the nine taps row-major
            if (MODE == 1) { ++dh; }
            else if (taps == 9) { if (++dw == 2) { dw = -1; ++dh; } }
          }
        }
      }
    }
  } else if (warp == 1 && is_leader) {
    // ================= MMA issuer =================
    // One thread feeds the tensor core(s); the body per weight tile stays at a few dozen instructions (ring cursors
    // instead of divisions, descriptors advanced by adding to their lower word) - measured: the longer form made the
    // issuing thread, not the tensor pipe, the pacer of 128-column tiles.
    constexpr uint32_t idesc = make_idesc_bf16_f32(kTileM * CG, BLOCK_N);
    if (elect_one_sync()) {
      StageRing ra, rb;
      const uint32_t a_lo0 = desc_lo_sw128(smem_a0), b_lo0 = desc_lo_sw128(smem_b0);
      int it = 0;
      for (int unit = first_unit; unit < total_units; unit += unit_stride, ++it) {
        const int buf = it & 1;
        mbar_wait(t_empty(buf), ((it >> 1) & 1) ^ 1u);  // epilogue(s) drained this accumulator buffer
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)(buf * Cfg::kAccCols);
        uint32_t accumulate = 0;
        for (int s = 0; s < p.nseg; ++s) {
          const int taps = p.seg_taps[s];
          const int cblocks = (p.seg_c[s] + kBlockK - 1) / kBlockK;
          const int asteps = (MODE == 1) ? (taps == 9 ? 3 : 1) : taps;
          const int bsteps = (MODE == 1) ? (taps == 9 ? 3 : 1) : 1;
          // MODE 1: tap kw reads the halo row starting kw pixels in (a 1x1 segment reads the centre, kw = 1)
          const uint32_t a_row0 = (MODE == 1 && taps != 9) ? (uint32_t)(kARowBytes >> 4) : 0u;
          for (int ac = 0; ac < asteps * cblocks; ++ac) {
            mbar_wait(a_full(ra.idx), ra.phase);
            uint32_t a_lo = a_lo0 + ra.idx * (uint32_t)(Cfg::kASlot >> 4) + a_row0;
            const int ndup = p.dup_koff ? 2 : 1;  // split-bf16 weights: the hi and the lo tile against the same A tile
            for (int bd = 0; bd < bsteps * ndup; ++bd, a_lo += (bd % ndup == 0) ? (uint32_t)(kARowBytes >> 4) : 0u) {
              mbar_wait(b_full(rb.idx), rb.phase);
              tc_fence_after();
              const uint32_t b_lo = b_lo0 + rb.idx * (uint32_t)(Cfg::kBBytes >> 4);
#pragma unroll
              for (int sub = 0; sub < MT; ++sub) {
                const uint32_t al = a_lo + (uint32_t)(sub * (Cfg::kASub >> 4));
                const uint32_t td = tmem_d + (uint32_t)(sub * BLOCK_N);
                if (CG == 2) {
                  umma_bf16_ss_pair_lh(td, al, b_lo, idesc, accumulate);  // first k-step of a unit overwrites
                  umma_bf16_ss_pair_lh(td, al + 2, b_lo + 2, idesc, 1u);
                  umma_bf16_ss_pair_lh(td, al + 4, b_lo + 4, idesc, 1u);
                  umma_bf16_ss_pair_lh(td, al + 6, b_lo + 6, idesc, 1u);
                } else {
                  umma_bf16_ss_lh(td, al, b_lo, idesc, accumulate);
                  umma_bf16_ss_lh(td, al + 2, b_lo + 2, idesc, 1u);
                  umma_bf16_ss_lh(td, al + 4, b_lo + 4, idesc, 1u);
                  umma_bf16_ss_lh(td, al + 6, b_lo + 6, idesc, 1u);
                }
              }
              accumulate = 1;
              if (CG == 2) umma_commit_pair(b_empty(rb.idx)); else umma_commit(b_empty(rb.idx));
              rb.advance(Cfg::kBStages);
            }
            if (CG == 2) umma_commit_pair(a_empty(ra.idx)); else umma_commit(a_empty(ra.idx));
            ra.advance(Cfg::kAStages);
          }
        }
        if (CG == 2) umma_commit_pair(t_full(buf)); else umma_commit(t_full(buf));
      }
    }
  } else if (warp >= 2) {
    // ================= epilogue (warps 2..9) =================
    const int quad = warp & 3;                 // TMEM lane quadrant this warp may access
    const int cgrp = (warp - 2) >> 2;          // which half of the staged columns this warp owns
    const int row = quad * 32 + lane;
    const int etid = threadIdx.x - 64;         // 0..255
    const bool store_issuer = (warp == 2) && elect_one_sync();
    const uint32_t t_empty_leader0 = (CG == 2) ? mapa_shared(t_empty(0), 0) : t_empty(0);
    const uint32_t t_empty_leader1 = (CG == 2) ? mapa_shared(t_empty(1), 0) : t_empty(1);
    constexpr int kHalves = BLOCK_N / Cfg::kHalfCols;
    constexpr int kWarpCols = Cfg::kHalfCols / 2;          // columns per warp per half (64, or 32 for BLOCK_N = 64)
    constexpr int kWarpChunks = kWarpCols / 8;             // 16-byte chunks per row per warp
    const bool vec_per_tile = (p.Nt == 1);                 // all rows of a tile belong to one image
    int it = 0, stage_use = 0;
    for (int unit = first_unit; unit < total_units; unit += unit_stride, ++it) {
     const int buf = it & 1;
#pragma unroll 1
     for (int sub = 0; sub < MT; ++sub) {
      int m_tile, w0, h0, n0, ncol_tile;
      tile_coords(unit, sub, m_tile, w0, h0, n0, ncol_tile);
      const int rw = row & (p.Wt - 1);
      const int rh = (row >> p.log_wt) & (p.Ht - 1);
      const int rn = row >> (p.log_wt + p.log_ht);
      const int ow = w0 + rw, oh = h0 + rh, on = n0 + rn;
      const bool valid = (ow < p.Wo) && (oh < p.Ho) && (on < p.B);

      // per-tile bias (+ per-image time-embedding vector) -> shared memory; the previous tile's readers are past
      // their last read (they all crossed the pre-store barrier of its last half)
      if (etid < BLOCK_N) {
        const int col = ncol_tile + etid;
        float bv = 0.f;
        if (col < p.Cout) {
          if (p.bias != nullptr) bv = __ldg(p.bias + col);
          if (p.addvec != nullptr && vec_per_tile && n0 < p.B) bv += __ldg(p.addvec + (size_t)n0 * p.addvec_stride + col);
        }
        sbias[etid] = bv;
      }

      mbar_wait(t_full(buf), (it >> 1) & 1);
      tc_fence_after();

#pragma unroll 1
      for (int half = 0; half < kHalves; ++half, ++stage_use) {
        const int ncol0 = ncol_tile + half * Cfg::kHalfCols;
        uint8_t* stg = out_gen + (stage_use & 1) * Cfg::kOutBytes;
        const uint32_t stg_u32 = smem_out + (stage_use & 1) * Cfg::kOutBytes;
        // the TMA store that last read THIS staging buffer (two halves ago) must be done reading it
        if (store_issuer) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
        asm volatile("bar.sync 1, 256;" ::: "memory");

        if (p.residual != nullptr) {
          // coalesced 16-byte reads of this warp's 32 rows x kWarpCols columns, staged into the swizzled tile
#pragma unroll
          for (int i0 = 0; i0 < kWarpChunks; i0 += 8) {
            // 32 rows x kWarpChunks chunks per warp = kWarpChunks loads per lane (4 for 64-wide tiles: the batch
            // must not run past the warp's own rows)
            constexpr int kU = kWarpChunks < 8 ? kWarpChunks : 8;
            uint4 buf4[kU];
#pragma unroll
            for (int u = 0; u < kU; ++u) {
              const int idx = (i0 + u) * 32 + lane;
              const int rl = idx / kWarpChunks, ch = cgrp * kWarpChunks + idx % kWarpChunks;
              const int rr = quad * 32 + rl;
              const int pw = w0 + (rr & (p.Wt - 1));
              const int ph = h0 + ((rr >> p.log_wt) & (p.Ht - 1));
              const int pn = n0 + (rr >> (p.log_wt + p.log_ht));
              const int col = ncol0 + ch * 8;
              buf4[u] = make_uint4(0, 0, 0, 0);
              if (pw < p.Wo && ph < p.Ho && pn < p.B && col < p.Cout)
                buf4[u] = *reinterpret_cast<const uint4*>(p.residual +
                                                          (((size_t)pn * p.Ho + ph) * p.Wo + pw) * p.Cout + col);
            }
#pragma unroll
            for (int u = 0; u < kU; ++u) {
              const int idx = (i0 + u) * 32 + lane;
              const int rl = idx / kWarpChunks, ch = cgrp * kWarpChunks + idx % kWarpChunks;
              const int rr = quad * 32 + rl;
              uint8_t* dst = stg + (ch >> 3) * (kTileM * 128) + rr * 128 + (((ch & 7) ^ (rr & 7)) * 16);
              *reinterpret_cast<uint4*>(dst) = buf4[u];
            }
          }
          __syncwarp();
        }

#pragma unroll 1
        for (int c0 = cgrp * kWarpCols; c0 < (cgrp + 1) * kWarpCols; c0 += 32) {
          uint32_t r[32];
          tmem_ld_32x32b_x32(tmem_base + ((uint32_t)(quad * 32) << 16) +
                                 (uint32_t)(buf * Cfg::kAccCols + sub * BLOCK_N + half * Cfg::kHalfCols + c0), r);
          tmem_ld_wait();
          const int col0 = ncol0 + c0;
          float v[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
          const int slab = c0 >> 6;
          const int chunk0 = (c0 & 63) >> 3;
          uint8_t* rowp = stg + slab * (kTileM * 128) + row * 128;
          epi_add_bias(v, sbias + half * Cfg::kHalfCols + c0);
          if (p.addvec != nullptr && !vec_per_tile && valid) {  // tiny images: several images per tile
#pragma unroll
            for (int j4 = 0; j4 < 8; ++j4) {
              if (col0 + j4 * 4 < p.Cout) {
                const float4 b = __ldg(reinterpret_cast<const float4*>(p.addvec + (size_t)on * p.addvec_stride + col0 + j4 * 4));
                v[j4 * 4 + 0] += b.x; v[j4 * 4 + 1] += b.y; v[j4 * 4 + 2] += b.z; v[j4 * 4 + 3] += b.w;
              }
            }
          }
          if (p.residual != nullptr) epi_add_residual(v, rowp, chunk0, row);
          if (p.gn_partial != nullptr) {
            // one row of partials per (M tile, lane quadrant): [(m_tile*4 + quad)][Cout/4][2], no atomics
            int vidx;
            const float red0 = epi_quad_stats(v, valid, lane, vidx);
            const int qcol = col0 + (vidx & 7) * 4;
            // (tiles of 2 or 4 whole images: a lane quadrant never straddles images, and row m_tile*4 + quad is row
            // quad % (4/Nt) of image n0 + quad / (4/Nt), i.e. the per-image rows stay contiguous)
            if ((lane & 1) == 0 && qcol < p.Cout && n0 + ((quad * p.Nt) >> 2) < p.B)
              p.gn_partial[(((size_t)m_tile * 4 + quad) * (p.Cout >> 2) + (qcol >> 2)) * 2 + (vidx >> 3)] = red0;
          }
          epi_pack_store(v, rowp, chunk0, row);
        }
        if (half == kHalves - 1 && sub == MT - 1) {
          // all TMEM reads of this accumulator buffer are done: hand it back to the MMA issuer (leader CTA)
          tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            const uint32_t bar = buf ? t_empty_leader1 : t_empty_leader0;
            asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(bar) : "memory");
          }
        }
        fence_proxy_async_smem();
        asm volatile("bar.sync 1, 256;" ::: "memory");
        if (store_issuer) {
#pragma unroll
          for (int slab = 0; slab < Cfg::kHalfCols / 64; ++slab) {
            if (ncol0 + slab * 64 < p.Cout)
              for (int u = 0; u < p.n_out; ++u)
                tma_store_4d(u == 0 ? &p.out : &p.out_up[u - 1], stg_u32 + slab * (kTileM * 128), ncol0 + slab * 64,
                             w0, h0, n0);
          }
          tma_store_commit();
        }
      }
     }
    }
    if (store_issuer) tma_store_wait_read0();
  }

  // ---- teardown ----
  tc_fence_before();
  if (CG == 2) cluster_sync_all(); else __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    if (CG == 2) tmem_dealloc_pair<kTmemCols>(tmem_base); else tmem_dealloc<kTmemCols>(tmem_base);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode() {
  static PFN_encodeTiled fn = nullptr;
  if (fn) return fn;
  void* sym = nullptr;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) != cudaSuccess ||
      qres != cudaDriverEntryPointSuccess || sym == nullptr)
    return nullptr;
  fn = reinterpret_cast<PFN_encodeTiled>(sym);
  return fn;
}

// 4D NHWC activation map: dims (C, W, H, N), box (64, bw, bh, bn), element strides (1, s, s, 1).
static int encode_act_map(CUtensorMap* m, const void* ptr, int C, int W, int H, int N, int bw, int bh, int bn,
                          int estride) {
  PFN_encodeTiled enc = get_encode();
  if (!enc) { set_error("cuTensorMapEncodeTiled unavailable"); return FM_ERR_NO_DEVICE; }
  cuuint64_t gdim[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
  cuuint64_t gstr[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
  cuuint32_t box[4] = {(cuuint32_t)kBlockK, (cuuint32_t)(bw * estride), (cuuint32_t)(bh * estride), (cuuint32_t)bn};
  cuuint32_t estr[4] = {1, (cuuint32_t)estride, (cuuint32_t)estride, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), gdim, gstr, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled(act C=%d W=%d H=%d N=%d box=%d,%d,%d stride=%d) failed: %d", C, W, H, N, bw, bh,
              bn, estride, (int)r);
    return (int)r;
  }
  return 0;
}

// Output map of one phase (dy, dx) of a nearest-2x upsampled store: the pixels (2h+dy, 2w+dx) of [N][2H][2W][C] seen as
// an [N][H][W][C] tensor with doubled pixel / row strides.
static int encode_up_phase_map(CUtensorMap* m, const void* ptr, int C, int W, int H, int N, int bw, int bh, int bn,
                               int dy, int dx) {
  PFN_encodeTiled enc = get_encode();
  if (!enc) { set_error("cuTensorMapEncodeTiled unavailable"); return FM_ERR_NO_DEVICE; }
  const char* base = static_cast<const char*>(ptr) + ((size_t)dy * 2 * W + dx) * C * 2;
  cuuint64_t gdim[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
  cuuint64_t gstr[3] = {(cuuint64_t)2 * C * 2, (cuuint64_t)4 * W * C * 2, (cuuint64_t)4 * H * W * C * 2};
  cuuint32_t box[4] = {(cuuint32_t)kBlockK, (cuuint32_t)bw, (cuuint32_t)bh, (cuuint32_t)bn};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<char*>(base), gdim, gstr, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled(upsampled out C=%d W=%d H=%d N=%d phase %d,%d) failed: %d", C, W, H, N, dy, dx,
              (int)r);
    return (int)r;
  }
  return 0;
}

static int encode_wgt_map(CUtensorMap* m, const void* ptr, int Ktot, int Cout, int block_n) {
  PFN_encodeTiled enc = get_encode();
  if (!enc) { set_error("cuTensorMapEncodeTiled unavailable"); return FM_ERR_NO_DEVICE; }
  cuuint64_t gdim[2] = {(cuuint64_t)Ktot, (cuuint64_t)Cout};
  cuuint64_t gstr[1] = {(cuuint64_t)Ktot * 2};
  cuuint32_t box[2] = {(cuuint32_t)kBlockK, (cuuint32_t)block_n};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), gdim, gstr, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled(weight K=%d Cout=%d) failed: %d", Ktot, Cout, (int)r);
    return (int)r;
  }
  return 0;
}

static int pow2_ceil(int v) {
  int p = 1;
  while (p < v) p <<= 1;
  return p;
}

template <int BLOCK_N, int MODE, int CG, int MT>
static int launch_conv_persistent(const ConvKernelParams& kp, cudaStream_t st) {
  using Cfg = PConvCfg<BLOCK_N, MODE, CG, MT>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(conv_igemm_persistent_kernel<BLOCK_N, MODE, CG, MT>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes);
    if (e != cudaSuccess) return check_cuda(e, "cudaFuncSetAttribute(conv_igemm_persistent)");
    attr_set = true;
  }
  const int units = (kp.m_tiles / (CG * MT)) * kp.n_tiles;
  int ctas = (sm_count() / CG) * CG;  // one CTA (pair) per SM (pair)
  if (ctas > units * CG) ctas = units * CG;
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(ctas);
  cfg.blockDim = dim3(kPConvThreads);
  cfg.dynamicSmemBytes = Cfg::kSmemBytes;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CG;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 2 : 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, conv_igemm_persistent_kernel<BLOCK_N, MODE, CG, MT>, kp);
  count_launch();
  if (e != cudaSuccess) return check_cuda(e, "cudaLaunchKernelEx(conv_igemm_persistent)");
  return 0;
}

}  // namespace fm

#include "conv_rolling.cuh"

namespace fm {

// Everything about a conv launch that follows from shapes, kernel sizes and the presence of operand transforms (no
// pointers): shared by the launcher and by fm_conv_stats_rows so the statistics workspace always matches the kernel.
struct ConvPlan {
  int Ho, Wo, Wt, Ht, Nt, tiles_w, tiles_h, tiles_n;
  bool row_mode, xf, rolling, pair;
  int block_n, n_tiles, mt, m_tiles;
  RollSched sch;
  int stats_rows;  // rows of GroupNorm partials per image (0: fused statistics unavailable)
};

static int plan_conv(const fm_conv_params* p, ConvPlan* pl) {
  FM_REQUIRE(p != nullptr, "conv: null params");
  FM_REQUIRE(p->nseg >= 1 && p->nseg <= FM_CONV_MAX_SEG, "conv: nseg=%d out of range", p->nseg);
  FM_REQUIRE(p->stride == 1 || p->stride == 2, "conv: stride must be 1 or 2 (got %d)", p->stride);
  FM_REQUIRE(p->B > 0 && p->H > 0 && p->W > 0, "conv: empty input %dx%dx%d", p->B, p->H, p->W);
  FM_REQUIRE(p->Cout > 0 && p->Cout % 8 == 0, "conv: Cout=%d must be a positive multiple of 8", p->Cout);
  memset(pl, 0, sizeof(*pl));
  pl->Ho = (p->H + p->stride - 1) / p->stride;
  pl->Wo = (p->W + p->stride - 1) / p->stride;
  // M-tile box: 128 output pixels as (Wt, Ht, Nt), powers of two
  pl->Wt = pow2_ceil(pl->Wo) < kTileM ? pow2_ceil(pl->Wo) : kTileM;
  const int rest = kTileM / pl->Wt;
  pl->Ht = pow2_ceil(pl->Ho) < rest ? pow2_ceil(pl->Ho) : rest;
  pl->Nt = rest / pl->Ht;
  pl->tiles_w = (pl->Wo + pl->Wt - 1) / pl->Wt;
  pl->tiles_h = (pl->Ho + pl->Ht - 1) / pl->Ht;
  pl->tiles_n = (p->B + pl->Nt - 1) / pl->Nt;
  bool any3 = false;
  for (int s = 0; s < p->nseg; ++s) {
    const fm_conv_seg& sg = p->seg[s];
    FM_REQUIRE(sg.C > 0 && sg.C % 8 == 0, "conv: segment %d channels=%d must be a positive multiple of 8", s, sg.C);
    FM_REQUIRE(sg.ksize == 1 || sg.ksize == 3, "conv: segment %d ksize=%d unsupported", s, sg.ksize);
    FM_REQUIRE(!(sg.ksize == 1 && p->stride != 1 && p->nseg > 1), "conv: fused 1x1 segment needs stride 1");
    if (sg.upsample) { set_error("conv: upsample-fused segments are not implemented yet"); return FM_ERR_UNSUPPORTED; }
    any3 |= (sg.ksize == 3);
    if (sg.norm_a != nullptr) {
      FM_REQUIRE(sg.norm_b != nullptr && sg.C % kBlockK == 0 && sg.norm_stride % 4 == 0 &&
                     ((uintptr_t)sg.norm_a & 15) == 0 && ((uintptr_t)sg.norm_b & 15) == 0,
                 "conv: segment %d operand transform needs C %% 64 == 0 and 16B-aligned a/b rows", s);
      pl->xf = true;
    }
  }
  // row mode (kw tap reuse): stride 1, M tile = 128 consecutive pixels of one row, at least one 3x3 segment
  pl->row_mode = p->stride == 1 && pl->Wt == kTileM && any3 && getenv("FMDM_CONV_NO_ROW_MODE") == nullptr;
  pl->block_n = (p->Cout > 128) ? 256 : (p->Cout > 64 ? 128 : 64);
  // rolling-row kernel: row-mode convs led by a 3x3 segment; N tiles of at most 128 columns
  pl->rolling = pl->row_mode && p->seg[0].ksize == 3;
  {
    const char* re = getenv("FMDM_CONV_ROLLING");  // 0/1: only where the operand transform is requested; 3: always
    const int rv = re ? atoi(re) : 2;
    if ((rv == 0 || rv == 1) && !pl->xf) pl->rolling = false;
    if (p->Cout > 128 && !pl->xf && rv != 3) pl->rolling = false;  // wide layers: the 256-column pair kernel unless fused
  }
  FM_REQUIRE(!pl->xf || pl->rolling, "conv: fused operand transform needs a leading 3x3 segment at stride 1 on image "
                                     "rows of >= 65 pixels (fm_conv_operand_norm_supported)");
  if (pl->rolling && pl->block_n == 256) pl->block_n = 128;
  // CTA pairs (cta_group::2): the 128/256-wide tiles when there are at least two M tiles; always for rolling rows
  const int m_tiles = pl->tiles_w * pl->tiles_h * pl->tiles_n;
  // Small problems (the 16x16 / 8x8 levels: a few dozen M tiles): 256-column tiles leave most SM pairs without a work
  // unit, so take the N tile that minimises (waves of work units) x (tile cost ~ columns + a fixed 64-column
  // equivalent of per-tile overhead).  Large problems keep 256 (many waves: fewer, fatter tiles win).
  if (!pl->rolling && pl->block_n == 256 && getenv("FMDM_CONV_NO_SMALL_N") == nullptr) {
    long best_cost = -1;
    int best_bn = 256;
    for (int bn = 256; bn >= 64; bn >>= 1) {
      const bool pr = bn >= 128 && m_tiles >= 2;
      const long units = (long)((m_tiles + (pr ? 1 : 0)) / (pr ? 2 : 1)) * ((p->Cout + bn - 1) / bn);
      const long slots = pr ? sm_count() / 2 : sm_count();
      const long cost = ((units + slots - 1) / slots) * (bn + 64);
      if (best_cost < 0 || cost < best_cost) best_cost = cost, best_bn = bn;
    }
    pl->block_n = best_bn;
  }
  pl->pair = (pl->block_n >= 128) && (m_tiles >= 2);
  {
    const char* pe = getenv("FMDM_CONV_PAIR");  // 0 disables, 1 (default) enables
    if (pe && atoi(pe) == 0) pl->pair = false;
  }
  if (pl->rolling) pl->pair = true;
  pl->n_tiles = (p->Cout + pl->block_n - 1) / pl->block_n;
  // two M tiles per CTA (MT = 2) for 128-wide N tiles when every SM pair still gets several work units
  pl->mt = (pl->pair && pl->block_n == 128 && m_tiles >= 8 * sm_count()) ? 2 : 1;
  {
    const char* me = getenv("FMDM_CONV_MT");  // 1 forces one M tile per CTA, 2 forces two wherever the variant exists
    if (me && atoi(me) == 1) pl->mt = 1;
    if (me && atoi(me) == 2 && pl->pair && pl->block_n == 128) pl->mt = 2;
  }
  const int m_round = (pl->pair ? 2 : 1) * pl->mt;
  pl->m_tiles = (m_tiles + m_round - 1) / m_round * m_round;
  if (pl->rolling) {
    pl->sch = roll_schedule(p->B, pl->Ho, pl->tiles_w, pl->n_tiles, sm_count() / 2);
    pl->stats_rows = pl->sch.chunks * pl->tiles_w * 4;      // one row per (strip, TMEM lane quadrant)
  } else {
    // one row per (M tile, lane quadrant); tiny images (2 or 4 per tile, >= 32 tile rows each): 4/Nt rows per image
    pl->stats_rows = (pl->Nt == 1) ? pl->tiles_w * pl->tiles_h * 4 : ((pl->Nt == 2 || pl->Nt == 4) ? 4 / pl->Nt : 0);
  }
  if (p->Cout % 4) pl->stats_rows = 0;
  return 0;
}

#include "conv_wgrad_tc.cuh"

}  // namespace fm

extern "C" int fm_conv_stats_rows(const fm_conv_params* p, int32_t* rows_per_image) {
  using namespace fm;
  FM_REQUIRE(rows_per_image != nullptr, "conv_stats_rows: null output");
  ConvPlan pl;
  if (int e = plan_conv(p, &pl)) return e;
  *rows_per_image = pl.stats_rows;
  return pl.stats_rows > 0 ? 0 : FM_ERR_UNSUPPORTED;
}

extern "C" int fm_conv_kernel_kind(const fm_conv_params* p) {
  using namespace fm;
  ConvPlan pl;
  if (int e = plan_conv(p, &pl)) return e;
  return pl.rolling ? (pl.xf ? 2 : 1) : 0;
}

extern "C" int fm_conv_operand_norm_supported(int32_t H, int32_t W, int32_t stride, int32_t has_3x3) {
  using namespace fm;
  if (H <= 0 || W <= 0) return 0;
  return (stride == 1 && has_3x3 && pow2_ceil(W) >= kTileM && getenv("FMDM_CONV_NO_ROW_MODE") == nullptr &&
          getenv("FMDM_CONV_NO_OPERAND_NORM") == nullptr) ? 1 : 0;
}

extern "C" int fm_conv2d_igemm_bf16(const fm_conv_params* p, fm_stream_t stream) {
  using namespace fm;
  if (int e = ensure_device()) return e;
  ConvPlan pl;
  if (int e = plan_conv(p, &pl)) return e;
  FM_REQUIRE(p->weight && p->out, "conv: null weight/out");
  FM_REQUIRE(((uintptr_t)p->weight & 15) == 0 && ((uintptr_t)p->out & 15) == 0, "conv: weight/out not 16B aligned");

  ConvKernelParams kp;
  memset(&kp, 0, sizeof(kp));
  kp.Wt = pl.Wt; kp.Ht = pl.Ht; kp.Nt = pl.Nt;
  kp.tiles_w = pl.tiles_w; kp.tiles_h = pl.tiles_h;
  kp.stride = p->stride;
  kp.B = p->B; kp.Ho = pl.Ho; kp.Wo = pl.Wo; kp.Cout = p->Cout;
  kp.nseg = p->nseg;
  int ktot = 0, nk = 0;
  for (int s = 0; s < p->nseg; ++s) {
    const fm_conv_seg& sg = p->seg[s];
    FM_REQUIRE(sg.src != nullptr && ((uintptr_t)sg.src & 15) == 0, "conv: segment %d source null/unaligned", s);
    kp.seg_na[s] = sg.norm_a;
    kp.seg_nb[s] = sg.norm_b;
    kp.seg_nstride[s] = sg.norm_stride;
    kp.seg_nact[s] = sg.norm_act;
    kp.seg_c[s] = sg.C;
    kp.seg_taps[s] = sg.ksize * sg.ksize;
    kp.seg_koff[s] = ktot;
    ktot += kp.seg_taps[s] * sg.C;
    nk += kp.seg_taps[s] * ((sg.C + kBlockK - 1) / kBlockK);
    if (int e = encode_act_map(&kp.src[s], sg.src, sg.C, p->W, p->H, p->B, pl.row_mode ? kTileM + 2 : kp.Wt, kp.Ht,
                               kp.Nt, p->stride))
      return e;
  }
  kp.num_k_blocks = nk;
  // Split-bf16 weights arrive as every source twice (hi segments, then the residual segments over the same tensors).
  // The per-tile kernel multiplies each A tile by both weight tiles instead of fetching it twice: half the operand
  // traffic of the narrow, launch-bound denoisers this mode exists for.
  kp.dup_koff = 0;
  if (!pl.rolling && p->nseg >= 2 && p->nseg % 2 == 0 && getenv("FMDM_CONV_NO_DUP") == nullptr) {
    const int h = p->nseg / 2;
    bool dup = true;
    for (int s = 0; s < h; ++s) {
      const fm_conv_seg &a = p->seg[s], &b = p->seg[s + h];
      dup = dup && a.src == b.src && a.C == b.C && a.ksize == b.ksize && a.norm_a == nullptr && b.norm_a == nullptr;
    }
    if (dup) {
      kp.nseg = h;
      kp.dup_koff = ktot / 2;
    }
  }
  const int block_n = pl.block_n;
  if (int e = encode_wgt_map(&kp.wgt, p->weight, ktot, p->Cout, pl.pair ? block_n / 2 : block_n)) return e;
  kp.n_out = 1;
  if (p->out_upsample) {
    for (int u = 0; u < 4; ++u)
      if (int e = encode_up_phase_map(u == 0 ? &kp.out : &kp.out_up[u - 1], p->out, p->Cout, pl.Wo, pl.Ho, p->B, kp.Wt,
                                      kp.Ht, kp.Nt, u >> 1, u & 1))
        return e;
    kp.n_out = 4;
  } else if (int e = encode_act_map(&kp.out, p->out, p->Cout, pl.Wo, pl.Ho, p->B, kp.Wt, kp.Ht, kp.Nt, 1)) {
    return e;
  }
  kp.bias = p->bias;
  kp.addvec = p->addvec;
  kp.addvec_stride = p->addvec_stride;
  FM_REQUIRE(p->addvec == nullptr || (p->addvec_stride % 4 == 0 && ((uintptr_t)p->addvec & 15) == 0),
             "conv: addvec must be 16B aligned with stride %% 4 == 0");
  FM_REQUIRE(p->bias == nullptr || ((uintptr_t)p->bias & 15) == 0, "conv: bias must be 16B aligned");
  kp.residual = reinterpret_cast<const __nv_bfloat16*>(p->residual);
  if (p->residual != nullptr && pl.rolling)
    if (int e = encode_act_map(&kp.res, p->residual, p->Cout, pl.Wo, pl.Ho, p->B, kp.Wt, kp.Ht, kp.Nt, 1)) return e;
  FM_REQUIRE(p->residual == nullptr || ((uintptr_t)p->residual & 15) == 0, "conv: residual must be 16B aligned");
  kp.gn_partial = p->gn_stats;
  kp.log_wt = 0; while ((1 << kp.log_wt) < kp.Wt) ++kp.log_wt;
  kp.log_ht = 0; while ((1 << kp.log_ht) < kp.Ht) ++kp.log_ht;
  FM_REQUIRE(p->gn_stats == nullptr || pl.stats_rows > 0,
             "conv: fused GroupNorm statistics need >= 32 tile rows per image and Cout %% 4 == 0 (fm_conv_stats_rows)");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  kp.n_tiles = pl.n_tiles;
  kp.m_tiles = pl.m_tiles;
  if (pl.rolling) {
    const bool res = p->residual != nullptr;  // two staging buffers only where the residual prefetch needs them
#define FM_RC(N, X) (res ? launch_conv_rolling<N, X, 2>(kp, pl.sch, st) : launch_conv_rolling<N, X, 1>(kp, pl.sch, st))
    if (pl.xf) return block_n == 64 ? FM_RC(64, 1) : FM_RC(128, 1);
    return block_n == 64 ? FM_RC(64, 0) : FM_RC(128, 0);
#undef FM_RC
  }
#define FM_PC(N, M, G) return launch_conv_persistent<N, M, G, 1>(kp, st)
  if (pl.pair && pl.mt == 2) {
    if (pl.row_mode) return launch_conv_persistent<128, 1, 2, 2>(kp, st);
    return launch_conv_persistent<128, 0, 2, 2>(kp, st);
  } else if (pl.pair) {
    if (pl.row_mode) { if (block_n == 128) FM_PC(128, 1, 2); else FM_PC(256, 1, 2); }
    else { if (block_n == 128) FM_PC(128, 0, 2); else FM_PC(256, 0, 2); }
  } else if (pl.row_mode) {
    if (block_n == 64) FM_PC(64, 1, 1); else if (block_n == 128) FM_PC(128, 1, 1); else FM_PC(256, 1, 1);
  } else {
    if (block_n == 64) FM_PC(64, 0, 1); else if (block_n == 128) FM_PC(128, 0, 1); else FM_PC(256, 0, 1);
  }
#undef FM_PC
}
