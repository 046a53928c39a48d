// K1, "rolling row" variant for stride-1 3x3 convs on images whose rows hold >= 65 pixels (the level-0/1/2 layers of
// the LDCT UNet: > 90 % of its FLOPs).  Included by conv_igemm.cu (shares ConvKernelParams and the host encoders).
//
// A CTA pair walks STRIPS: 128 consecutive pixels of `R` consecutive output rows (one strip per CTA, two per pair,
// tcgen05 cta_group::2, M = 256).  Input rows are streamed ONCE per strip: the 130-pixel halo row (kh-independent) of a
// 64-channel block lands in shared memory with one TMA load and feeds all nine taps - the three kw taps as row-shifted
// views of the slot (as in row mode) and the three kh taps by accumulating into the accumulators of three DIFFERENT
// output rows (o = r+1, r, r-1).  Four accumulators (4 x BLOCK_N TMEM columns) form a ring: three are live, the fourth
// is drained by the epilogue while the next input row streams.  Against the per-tile row mode this cuts the A operand
// traffic from L2 and - the point - the number of times an input element passes the operand transform by 3x, so the
// fused GroupNorm-apply + SiLU (XF) needs one pass per element and its four transform warps keep up with the MMAs.
//
// Warp roles: 0 weight (B) TMA producer, 1 MMA issuer (leader CTA), 2..9 epilogue, 10..13 operand transform (XF only),
// last warp: activation (A) TMA producer.  A and B have separate producer threads because one in-order thread cannot
// run the A loads further ahead than its B ring allows (11 tiles ~ 1.2 input rows); with the operand transform between
// the TMA landing and the MMA, A needs the deeper prefetch (measured: the MMA issuer waited on a_ready while the
// transform warps idled on a_full).
#pragma once

namespace fm {

constexpr int kRollAccBufs = 4;
// XF = 1: four extra "transform" warps rewrite every landed A slot in place with act(a[n,c]*x + b[n,c]) (GroupNorm apply
// + SiLU from the producer-side statistics) before the MMA issuer may read it, so the normalised activation tensor is
// never written to or re-read from HBM.  Halo rows/pixels that TMA zero-filled stay zero (the reference pads AFTER the
// activation).
constexpr int kXfWarps = 4;
constexpr int kXfThreads = kXfWarps * 32;

// 8 bf16 channels (one 16-byte chunk) through act(a*x+b); a, b pre-halved when kSilu (SiLU(2h) = h + h*tanh(h)).
template <bool kSilu>
__device__ __forceinline__ uint4 xf_chunk(uint4 u, const float (&a)[8], const float (&b)[8]) {
  uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    float2 h = ffma2(make_float2(__uint_as_float(w[j] << 16), __uint_as_float(w[j] & 0xffff0000u)),
                     make_float2(a[2 * j], a[2 * j + 1]), make_float2(b[2 * j], b[2 * j + 1]));
    if (kSilu) {
      float2 t;
      asm("tanh.approx.f32 %0, %1;" : "=f"(t.x) : "f"(h.x));
      asm("tanh.approx.f32 %0, %1;" : "=f"(t.y) : "f"(h.y));
      h = ffma2(h, t, h);
    }
    w[j] = pack_bf16x2(h.x, h.y);
  }
  return make_uint4(w[0], w[1], w[2], w[3]);
}

// OB = output staging buffers: 2 when the conv has a residual (the next tile's residual is prefetched into the other
// buffer), else 1 - the freed 32 KB deepen the weight ring from 7 to 11 stages (+2.5..5 % on residual-free convs).
template <int BLOCK_N, int XF, int OB>
struct RConvCfg {
  static_assert(BLOCK_N == 64 || BLOCK_N == 128, "four accumulators must fit the 512 TMEM columns");
  static constexpr int kBBytes = (BLOCK_N / 2) * kBlockK * 2;  // per CTA (each CTA of the pair stages half of N)
  static constexpr int kASlot = kARowSlot;
  static constexpr int kATx = kARowTx;
  static constexpr int kOutBytes = kTileM * BLOCK_N * 2;
  static constexpr int kOutBufs = OB;
  static constexpr int kTailBytes = 512 + BLOCK_N * 4;
  static constexpr int kBudget = 227 * 1024 - 1024 - kTailBytes - kOutBufs * kOutBytes;
  // six A slots: a 1x1 segment's slot holds only 4 MMAs (256 tensor cycles); with four, a run of such slots drained the
  // ring faster than TMA + transform refill it (level-0 ResBlock tails: +10 % with six, plain 3x3 convs unchanged)
  static constexpr int kAStages = 6;
  static constexpr int kBStagesRaw = (kBudget - kAStages * kASlot) / kBBytes;
  static constexpr int kBStages = kBStagesRaw > 12 ? 12 : kBStagesRaw;
  static constexpr int kPipeBytes = kAStages * kASlot + kBStages * kBBytes;
  static constexpr int kNumBars = 2 * kAStages + 2 * kBStages + 2 * kRollAccBufs + 2 + XF * kAStages;
  static constexpr int kSmemBytes = 1024 + kPipeBytes + kOutBufs * kOutBytes + kTailBytes;
  static constexpr int kTmemCols = kRollAccBufs * BLOCK_N;
  static_assert(kBStages >= 7, "weight pipeline too shallow");
  static_assert(kNumBars * 8 + 8 <= 512, "barrier area");
  static_assert(kSmemBytes <= 227 * 1024, "shared memory budget");
};

struct RollSched {
  int R;        // output rows per strip
  int chunks;   // row chunks per image
  int combos;   // (image, w-tile) combinations, padded to even so the two strips of a pair share their row chunk
  int pairs;    // strip pairs = chunks * combos / 2
};

template <int BLOCK_N, int XF, int OB>
__global__ void __launch_bounds__(kPConvThreads + XF * kXfThreads + 32, 1)
conv_rolling_kernel(const __grid_constant__ ConvKernelParams p, const RollSched sch) {
  using Cfg = RConvCfg<BLOCK_N, XF, OB>;
  const uint32_t cta_rank = cluster_ctarank();
  const bool is_leader = (cta_rank == 0);
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t smem_a0 = smem_base;
  const uint32_t smem_b0 = smem_base + Cfg::kAStages * Cfg::kASlot;
  const uint32_t smem_out = smem_base + Cfg::kPipeBytes;
  uint8_t* out_gen = smem_gen + Cfg::kPipeBytes;
  const uint32_t bar_base = smem_out + Cfg::kOutBufs * Cfg::kOutBytes;
  float* sbias = reinterpret_cast<float*>(out_gen + Cfg::kOutBufs * Cfg::kOutBytes + 512);  // [BLOCK_N]
  constexpr int kBarT = 2 * Cfg::kAStages + 2 * Cfg::kBStages;
  auto a_full = [&](int s) { return bar_base + 8u * s; };
  auto a_empty = [&](int s) { return bar_base + 8u * (Cfg::kAStages + s); };
  auto b_full = [&](int s) { return bar_base + 8u * (2 * Cfg::kAStages + s); };
  auto b_empty = [&](int s) { return bar_base + 8u * (2 * Cfg::kAStages + Cfg::kBStages + s); };
  auto t_full = [&](int b) { return bar_base + 8u * (kBarT + b); };
  auto t_empty = [&](int b) { return bar_base + 8u * (kBarT + kRollAccBufs + b); };
  auto r_full = [&](int b) { return bar_base + 8u * (kBarT + 2 * kRollAccBufs + b); };   // residual tile landed
  auto a_ready = [&](int s) { return bar_base + 8u * (kBarT + 2 * kRollAccBufs + 2 + s); };  // XF only
  const uint32_t tmem_slot = bar_base + 8u * Cfg::kNumBars;
  volatile uint32_t* tmem_slot_gen =
      reinterpret_cast<volatile uint32_t*>(out_gen + Cfg::kOutBufs * Cfg::kOutBytes + 8 * Cfg::kNumBars);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < p.nseg; ++s) tma_prefetch_desc(&p.src[s]);
    tma_prefetch_desc(&p.wgt);
    tma_prefetch_desc(&p.out);
    if (p.residual != nullptr) tma_prefetch_desc(&p.res);
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < kBarT + kRollAccBufs; ++s) mbar_init(bar_base + 8u * s, 1);
      for (int b = 0; b < kRollAccBufs; ++b) mbar_init(t_empty(b), kEpiWarps * 2);
      mbar_init(r_full(0), 1);
      mbar_init(r_full(1), 1);
      if (XF) for (int s = 0; s < Cfg::kAStages; ++s) mbar_init(a_ready(s), kXfWarps * 2);
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc_pair<Cfg::kTmemCols>(tmem_slot);
  }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_gen;
  // PDL: everything above is private set-up; from here on global memory of the preceding kernels is touched
  pdl_trigger();
  pdl_wait();

  // ---- static schedule: unit = (strip pair, N tile), N fastest --------------------------------------------------
  const int total_units = sch.pairs * p.n_tiles;
  const int first_unit = (int)blockIdx.x >> 1;
  const int unit_stride = (int)gridDim.x >> 1;
  // this CTA's strip of a unit: 128 pixels starting at w0 of rows [h_begin, h_end) of image n
  auto strip_coords = [&](int unit, int& tw, int& w0, int& n, int& h_begin, int& h_end, int& ncol0) {
    const int ps = unit / p.n_tiles;
    ncol0 = (unit - ps * p.n_tiles) * BLOCK_N;
    const int sid = ps * 2 + (int)cta_rank;
    const int chunk = sid / sch.combos;
    const int j = sid - chunk * sch.combos;
    n = j / p.tiles_w;
    tw = j - n * p.tiles_w;
    w0 = tw * kTileM;
    h_begin = chunk * sch.R;
    h_end = h_begin + sch.R;
    if (h_end > p.Ho) h_end = p.Ho;
    return chunk;
  };
  // taps of input row r that land on output rows of [h_begin, h_end): kh in [kh_lo, kh_hi], output row o = r + 1 - kh
  auto kh_range = [](int r, int h_begin, int h_end, int& kh_lo, int& kh_hi) {
    kh_lo = r + 2 - h_end;
    if (kh_lo < 0) kh_lo = 0;
    kh_hi = r + 1 - h_begin;
    if (kh_hi > 2) kh_hi = 2;
  };

  if (warp == 0) {
    // ================= weight (B) TMA producer =================
    if (elect_one_sync()) {
      StageRing rb;
      const uint32_t b_full_leader0 = mapa_shared(b_full(0), 0);
      const int bcol_off = (int)cta_rank * (BLOCK_N / 2);
      for (int unit = first_unit; unit < total_units; unit += unit_stride) {
        int tw, w0, n, h_begin, h_end, ncol0;
        strip_coords(unit, tw, w0, n, h_begin, h_end, ncol0);
        const int bcol = ncol0 + bcol_off;
        for (int r = h_begin - 1; r <= h_end; ++r) {
          int kh_lo, kh_hi;
          kh_range(r, h_begin, h_end, kh_lo, kh_hi);
          for (int s = 0; s < p.nseg; ++s) {
            const int taps = p.seg_taps[s];
            if (taps == 1 && (r < h_begin || r >= h_end)) continue;  // a 1x1 segment only feeds its own row
            const int C = p.seg_c[s];
            const int cblocks = (C + kBlockK - 1) / kBlockK;
            const int t_lo = (taps == 9) ? kh_lo * 3 : 0, t_n = (taps == 9) ? (kh_hi - kh_lo + 1) * 3 : 1;
            for (int cb = 0; cb < cblocks; ++cb) {
              int kcol = p.seg_koff[s] + t_lo * C + cb * kBlockK;
              for (int t = 0; t < t_n; ++t, kcol += C) {
                mbar_wait(b_empty(rb.idx), rb.phase ^ 1u);
                if (is_leader) mbar_expect_tx(b_full(rb.idx), 2 * Cfg::kBBytes);
                tma_load_2d_pair(&p.wgt, b_full_leader0 + 8u * rb.idx, smem_b0 + rb.idx * Cfg::kBBytes, kcol, bcol);
                rb.advance(Cfg::kBStages);
              }
            }
          }
        }
      }
    }
  } else if (warp == 2 + kEpiWarps + XF * kXfWarps) {
    // ================= activation (A) TMA producer: one 130-pixel halo row of 64 channels per slot =================
    if (elect_one_sync()) {
      StageRing ra;
      const uint32_t a_full_leader0 = mapa_shared(a_full(0), 0);
      for (int unit = first_unit; unit < total_units; unit += unit_stride) {
        int tw, w0, n, h_begin, h_end, ncol0;
        strip_coords(unit, tw, w0, n, h_begin, h_end, ncol0);
        for (int r = h_begin - 1; r <= h_end; ++r) {
          for (int s = 0; s < p.nseg; ++s) {
            if (p.seg_taps[s] == 1 && (r < h_begin || r >= h_end)) continue;
            const int cblocks = (p.seg_c[s] + kBlockK - 1) / kBlockK;
            for (int cb = 0; cb < cblocks; ++cb) {
              mbar_wait(a_empty(ra.idx), ra.phase ^ 1u);
              if (XF) {  // each CTA's transform warps wait for their own bytes
                mbar_expect_tx(a_full(ra.idx), Cfg::kATx);
                tma_load_4d(&p.src[s], a_full(ra.idx), smem_a0 + ra.idx * Cfg::kASlot, cb * kBlockK, w0 - 1, r, n);
              } else {
                if (is_leader) mbar_expect_tx(a_full(ra.idx), 2 * Cfg::kATx);
                tma_load_4d_pair(&p.src[s], a_full_leader0 + 8u * ra.idx, smem_a0 + ra.idx * Cfg::kASlot,
                                 cb * kBlockK, w0 - 1, r, n);
              }
              ra.advance(Cfg::kAStages);
            }
          }
        }
      }
    }
  } else if (warp == 1 && is_leader) {
    // ================= MMA issuer =================
    // One thread feeds the tensor cores of both SMs: the loop body per weight tile (4 MMAs = ~256 tensor cycles) is
    // kept to a few dozen instructions - ring cursors instead of divisions, descriptors advanced by adding to their
    // lower word.
    constexpr uint32_t idesc = make_idesc_bf16_f32(kTileM * 2, BLOCK_N);
    if (elect_one_sync()) {
      StageRing ra, rb;
      int oc0 = 0;  // output rows finished by this CTA pair before the current unit
      const uint32_t a_lo0 = desc_lo_sw128(smem_a0), b_lo0 = desc_lo_sw128(smem_b0);
      auto issue_tap = [&](uint32_t tmem_d, uint32_t a_lo, uint32_t first_acc) {
        mbar_wait(b_full(rb.idx), rb.phase);
        tc_fence_after();
        const uint32_t b_lo = b_lo0 + rb.idx * (uint32_t)(Cfg::kBBytes >> 4);
        umma_bf16_ss_pair_lh(tmem_d, a_lo, b_lo, idesc, first_acc);
        umma_bf16_ss_pair_lh(tmem_d, a_lo + 2, b_lo + 2, idesc, 1u);
        umma_bf16_ss_pair_lh(tmem_d, a_lo + 4, b_lo + 4, idesc, 1u);
        umma_bf16_ss_pair_lh(tmem_d, a_lo + 6, b_lo + 6, idesc, 1u);
        umma_commit_pair(b_empty(rb.idx));
        rb.advance(Cfg::kBStages);
      };
      for (int unit = first_unit; unit < total_units; unit += unit_stride) {
        int tw, w0, n, h_begin, h_end, ncol0;
        strip_coords(unit, tw, w0, n, h_begin, h_end, ncol0);
        for (int r = h_begin - 1; r <= h_end; ++r) {
          int kh_lo, kh_hi;
          kh_range(r, h_begin, h_end, kh_lo, kh_hi);
          const int ocr = oc0 + (r + 1 - h_begin);  // ring position of output row r+1 (kh = 0)
          if (r + 1 < h_end) {
            // output row r+1 starts accumulating with this input row: its ring slot must have been drained
            mbar_wait(t_empty(ocr & 3), ((ocr >> 2) & 1) ^ 1u);
            tc_fence_after();
          }
          for (int s = 0; s < p.nseg; ++s) {
            const int taps = p.seg_taps[s];
            if (taps == 1 && (r < h_begin || r >= h_end)) continue;
            const int cblocks = (p.seg_c[s] + kBlockK - 1) / kBlockK;
            for (int cb = 0; cb < cblocks; ++cb) {
              mbar_wait(XF ? a_ready(ra.idx) : a_full(ra.idx), ra.phase);
              const uint32_t a_lo = a_lo0 + ra.idx * (uint32_t)(Cfg::kASlot >> 4);
              if (taps == 9) {
                for (int kh = kh_lo; kh <= kh_hi; ++kh) {
                  const uint32_t tmem_d = tmem_base + (uint32_t)(((ocr - kh) & 3) * BLOCK_N);
                  // the very first MMA into an output row (kh = 0 of segment 0, channel block 0, kw = 0) overwrites
                  const uint32_t acc0 = (kh == 0 && s == 0 && cb == 0) ? 0u : 1u;
                  issue_tap(tmem_d, a_lo, acc0);                                // kw = 0: slot rows 0..127
                  issue_tap(tmem_d, a_lo + (kARowBytes >> 4), 1u);              // kw = 1: rows 1..128
                  issue_tap(tmem_d, a_lo + 2 * (kARowBytes >> 4), 1u);          // kw = 2: rows 2..129
                }
              } else {  // 1x1 segment: centre pixel, own row (kh = 1)
                issue_tap(tmem_base + (uint32_t)(((ocr - 1) & 3) * BLOCK_N), a_lo + (kARowBytes >> 4), 1u);
              }
              umma_commit_pair(a_empty(ra.idx));
              ra.advance(Cfg::kAStages);
            }
          }
          if (r - 1 >= h_begin) umma_commit_pair(t_full((ocr - 2) & 3));  // output row r-1 received its last tap
        }
        oc0 += h_end - h_begin;
      }
    }
  } else if (XF && warp >= 2 + kEpiWarps) {
    // ================= operand transform (warps 10..13) =================
    const int tt = (int)threadIdx.x - (64 + kEpiThreads);
    const int lc = tt & 7;   // this thread's 8-channel chunk of the 64-channel block
    const int r0 = tt >> 3;  // first slot row, step 16
    const uint32_t ready_bar0 = mapa_shared(a_ready(0), 0);
    int ia = 0;
    for (int unit = first_unit; unit < total_units; unit += unit_stride) {
      int tw, w0, n, h_begin, h_end, ncol0;
      strip_coords(unit, tw, w0, n, h_begin, h_end, ncol0);
      for (int r = h_begin - 1; r <= h_end; ++r) {
        for (int s = 0; s < p.nseg; ++s) {
          const int taps = p.seg_taps[s];
          if (taps == 1 && (r < h_begin || r >= h_end)) continue;
          const int cblocks = (p.seg_c[s] + kBlockK - 1) / kBlockK;
          const float* na = p.seg_na[s];
          const float* nb = p.seg_nb[s];
          const bool silu = p.seg_nact[s] != 0;
          const bool live = (na != nullptr) && ((unsigned)r < (unsigned)p.Ho) && (n < p.B);
          for (int cb = 0; cb < cblocks; ++cb, ++ia) {
            const int sa = ia % Cfg::kAStages;
            float a[8], b[8];
            if (live) {  // coefficient loads overlap the wait for the slot
              const size_t co = (size_t)n * p.seg_nstride[s] + cb * kBlockK + lc * 8;
              const float4 a0 = __ldg(reinterpret_cast<const float4*>(na + co));
              const float4 a1 = __ldg(reinterpret_cast<const float4*>(na + co + 4));
              const float4 b0 = __ldg(reinterpret_cast<const float4*>(nb + co));
              const float4 b1 = __ldg(reinterpret_cast<const float4*>(nb + co + 4));
              const float k = silu ? 0.5f : 1.0f;
              a[0] = a0.x * k; a[1] = a0.y * k; a[2] = a0.z * k; a[3] = a0.w * k;
              a[4] = a1.x * k; a[5] = a1.y * k; a[6] = a1.z * k; a[7] = a1.w * k;
              b[0] = b0.x * k; b[1] = b0.y * k; b[2] = b0.z * k; b[3] = b0.w * k;
              b[4] = b1.x * k; b[5] = b1.y * k; b[6] = b1.z * k; b[7] = b1.w * k;
            }
            mbar_wait(a_full(sa), (ia / Cfg::kAStages) & 1);
            if (live) {
              uint8_t* slot = smem_gen + sa * Cfg::kASlot;
              const int wbase = w0 - 1;  // input pixel of slot row 0
              // rows r0, r0+16, ..., r0+112 (< 128 for every thread): four chunks in flight, stores predicated on the
              // pixel lying inside the image (padding must stay zero)
#pragma unroll
              for (int i0 = 0; i0 < 8; i0 += 4) {
                uint4* ptr[4];
                uint4 u[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                  const int rr = r0 + 16 * (i0 + j);
                  ptr[j] = reinterpret_cast<uint4*>(slot + rr * 128 + ((lc ^ (rr & 7)) << 4));
                  u[j] = *ptr[j];
                }
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                  const int rr = r0 + 16 * (i0 + j);
                  const uint4 o = silu ? xf_chunk<true>(u[j], a, b) : xf_chunk<false>(u[j], a, b);
                  if ((unsigned)(wbase + rr) < (unsigned)p.Wo) *ptr[j] = o;
                }
              }
              if (r0 < 2) {  // slot rows 128, 129
                const int rr = r0 + 128;
                uint4* ptr = reinterpret_cast<uint4*>(slot + rr * 128 + ((lc ^ (rr & 7)) << 4));
                const uint4 o = silu ? xf_chunk<true>(*ptr, a, b) : xf_chunk<false>(*ptr, a, b);
                if ((unsigned)(wbase + rr) < (unsigned)p.Wo) *ptr = o;
              }
              fence_proxy_async_smem();  // generic-proxy writes -> visible to the tensor core's async-proxy reads
            }
            __syncwarp();
            if (lane == 0)
              asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(ready_bar0 + 8u * sa)
                           : "memory");
          }
        }
      }
    }
  } else if (warp >= 2) {
    // ================= epilogue (warps 2..9): one output row tile (128 pixels x BLOCK_N) at a time =================
    const int quad = warp & 3;
    const int cgrp = (warp - 2) >> 2;
    const int row = quad * 32 + lane;
    const int etid = threadIdx.x - 64;
    const bool store_issuer = (warp == 2) && elect_one_sync();
    const uint32_t t_empty_leader0 = mapa_shared(t_empty(0), 0);
    constexpr int kWarpCols = BLOCK_N / 2;
    constexpr int kWarpChunks = kWarpCols / 8;
    int oc = 0, stage_use = 0;
    // Residual tiles arrive by TMA, in the staging buffer's own swizzled layout, ONE TILE AHEAD of their use (issued by
    // the store thread while the previous tile is processed): with synchronous loads the epilogue paid a global-memory
    // latency per tile and became the pacer of every conv with an identity skip (+20 % run time at level 0).
    auto load_residual = [&](int use, int rw0, int rh0, int rn0, int rncol0) {
      const uint32_t dst = smem_out + (use & 1) * Cfg::kOutBytes, bar = r_full(use & 1);
      mbar_expect_tx(bar, Cfg::kOutBytes);
#pragma unroll
      for (int slab = 0; slab < BLOCK_N / 64; ++slab)  // slabs past Cout are zero-filled by TMA (bytes still counted)
        tma_load_4d(&p.res, bar, dst + slab * (kTileM * 128), rncol0 + slab * 64, rw0, rh0, rn0);
    };
    if (OB == 2 && store_issuer && p.residual != nullptr && first_unit < total_units) {
      int tw, w0, n0, h_begin, h_end, ncol0;
      strip_coords(first_unit, tw, w0, n0, h_begin, h_end, ncol0);
      load_residual(0, w0, h_begin, n0, ncol0);
    }
    for (int unit = first_unit; unit < total_units; unit += unit_stride) {
      int tw, w0, n0, h_begin, h_end, ncol0;
      const int strip_chunk = strip_coords(unit, tw, w0, n0, h_begin, h_end, ncol0);
      // GroupNorm partial statistics accumulate over the rows of the strip (one global row per strip and lane quadrant)
      float st_acc[kWarpCols / 32];
#pragma unroll
      for (int i = 0; i < kWarpCols / 32; ++i) st_acc[i] = 0.f;
      const size_t stat_row = ((size_t)(n0 * sch.chunks + strip_chunk) * p.tiles_w + tw) * 4 + quad;
      for (int h0 = h_begin; h0 < h_end; ++h0, ++oc, ++stage_use) {
        const int buf = oc & 3;
        const int ow = w0 + row;
        const bool valid = (ow < p.Wo) && (n0 < p.B);

        if (etid < BLOCK_N) {
          const int col = ncol0 + etid;
          float bv = 0.f;
          if (col < p.Cout) {
            if (p.bias != nullptr) bv = __ldg(p.bias + col);
            if (p.addvec != nullptr && n0 < p.B) bv += __ldg(p.addvec + (size_t)n0 * p.addvec_stride + col);
          }
          sbias[etid] = bv;
        }
        uint8_t* stg = out_gen + (stage_use & (OB - 1)) * Cfg::kOutBytes;
        const uint32_t stg_u32 = smem_out + (stage_use & (OB - 1)) * Cfg::kOutBytes;
        // the TMA store that last read THIS staging buffer (OB tiles ago) must be done reading it
        if (store_issuer) {
          if (OB == 2) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
          else asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        }
        asm volatile("bar.sync 1, 256;" ::: "memory");

        mbar_wait(t_full(buf), (oc >> 2) & 1);
        tc_fence_after();

        // pull this warp's whole share of the accumulator into registers and hand the ring slot back to the MMA issuer
        // at once: with three of the four accumulators live, the drain latency (not its throughput) would otherwise
        // stall the next input row
        uint32_t rg[kWarpCols];
#pragma unroll
        for (int i = 0; i < kWarpCols / 32; ++i)
          tmem_ld_32x32b_x32(tmem_base + ((uint32_t)(quad * 32) << 16) +
                                 (uint32_t)(buf * BLOCK_N + cgrp * kWarpCols + i * 32), rg + i * 32);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0)
          asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(t_empty_leader0 + 8u * buf)
                       : "memory");
        if (OB == 2 && p.residual != nullptr) {
          if (store_issuer) {
            // next tile of this CTA: next row of the strip, else the first row of its next strip
            int nw0 = w0, nh0 = h0 + 1, nn0 = n0, nncol0 = ncol0;
            bool has_next = nh0 < h_end;
            if (!has_next && unit + unit_stride < total_units) {
              int ntw, nh_end;
              strip_coords(unit + unit_stride, ntw, nw0, nn0, nh0, nh_end, nncol0);
              has_next = true;
            }
            if (has_next) {
              // the other staging buffer was last read by the previous tile's TMA store
              asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
              load_residual(stage_use + 1, nw0, nh0, nn0, nncol0);
            }
          }
          mbar_wait(r_full(stage_use & 1), (stage_use >> 1) & 1);
        }

#pragma unroll
        for (int i = 0; i < kWarpCols / 32; ++i) {
          const int c0 = cgrp * kWarpCols + i * 32;
          const int col0 = ncol0 + c0;
          float v[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(rg[i * 32 + j]);
          const int slab = c0 >> 6;
          const int chunk0 = (c0 & 63) >> 3;
          uint8_t* rowp = stg + slab * (kTileM * 128) + row * 128;
          epi_add_bias(v, sbias + c0);
          if (OB == 2 && p.residual != nullptr) epi_add_residual(v, rowp, chunk0, row);
          if (p.gn_partial != nullptr) {
            int vidx;
            st_acc[i] += epi_quad_stats(v, valid, lane, vidx);
            if (h0 == h_end - 1) {  // last row of the strip: one row of partials per (strip, lane quadrant)
              const int qcol = col0 + (vidx & 7) * 4;
              if ((lane & 1) == 0 && qcol < p.Cout && n0 < p.B)
                p.gn_partial[(stat_row * (p.Cout >> 2) + (qcol >> 2)) * 2 + (vidx >> 3)] = st_acc[i];
            }
          }
          epi_pack_store(v, rowp, chunk0, row);
        }
        fence_proxy_async_smem();
        asm volatile("bar.sync 1, 256;" ::: "memory");
        if (store_issuer) {
#pragma unroll
          for (int slab = 0; slab < BLOCK_N / 64; ++slab) {
            if (ncol0 + slab * 64 < p.Cout)
              for (int u = 0; u < p.n_out; ++u)
                tma_store_4d(u == 0 ? &p.out : &p.out_up[u - 1], stg_u32 + slab * (kTileM * 128), ncol0 + slab * 64,
                             w0, h0, n0);
          }
          tma_store_commit();
        }
      }
    }
    if (store_issuer) tma_store_wait_read0();
  }

  // ---- teardown ----
  tc_fence_before();
  cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_pair<Cfg::kTmemCols>(tmem_base);
  }
}

// Rows per strip: the largest R in {32, 16, 8, 4} whose static round-robin over the SM pairs loses <= 4 % to the last
// partial round (halo rows cost (R+2)/R operand loads and transforms, but no extra MMAs), else the best of them.
static RollSched roll_schedule(int B, int Ho, int tiles_w, int n_tiles, int sm_pairs) {
  RollSched best{};
  double best_cost = 1e30;
  const int combos = (B * tiles_w + 1) & ~1;
  for (int R = 32; R >= 4; R >>= 1) {
    const int chunks = (Ho + R - 1) / R;
    const int pairs = chunks * combos / 2;
    const long units = (long)pairs * n_tiles;
    const long rounds = (units + sm_pairs - 1) / sm_pairs;
    // time ~ rounds * (R rows of MMAs) with a small penalty for the two halo rows' operand traffic
    const double cost = (double)rounds * (R + 0.25 * 2.0);
    if (cost < best_cost * 0.96) {
      best_cost = cost;
      best = RollSched{R, chunks, combos, pairs};
    }
  }
  return best;
}

template <int BLOCK_N, int XF, int OB>
static int launch_conv_rolling(const ConvKernelParams& kp, const RollSched& sch, cudaStream_t st) {
  using Cfg = RConvCfg<BLOCK_N, XF, OB>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(conv_rolling_kernel<BLOCK_N, XF, OB>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         Cfg::kSmemBytes);
    if (e != cudaSuccess) return check_cuda(e, "cudaFuncSetAttribute(conv_rolling)");
    attr_set = true;
  }
  const int units = sch.pairs * kp.n_tiles;
  int ctas = (sm_count() / 2) * 2;
  if (ctas > units * 2) ctas = units * 2;
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(ctas);
  cfg.blockDim = dim3(kPConvThreads + XF * kXfThreads + 32);
  cfg.dynamicSmemBytes = Cfg::kSmemBytes;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 2 : 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, conv_rolling_kernel<BLOCK_N, XF, OB>, kp, sch);
  count_launch();
  if (e != cudaSuccess) return check_cuda(e, "cudaLaunchKernelEx(conv_rolling)");
  return 0;
}

}  // namespace fm
