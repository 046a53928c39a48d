// conv wgrad on the tcgen05 tensor cores (included by conv_igemm.cu, which owns the tensor-map encoders).
//
//   dW[co][ci][kh][kw] = sum_p dY[p][co] * X[p (+) (kh,kw)][ci]
//
// is a GEMM whose reduction (K) axis is the pixel axis, so BOTH operands are MN-major as they lie in HBM (NHWC: the
// channel = M/N index is contiguous).  TMA drops a [64 pixels][64 channels] box into shared memory as 64 rows of
// 128 B with the 128-byte swizzle -- exactly the canonical MN-major SWIZZLE_128B operand layout
// ((8,n),(8,k)):((1,LBO),(8,SBO)) in 16-byte units: 8-pixel groups are SBO = 1024 B apart, 64-channel column blocks
// LBO = 8192 B apart.  One tcgen05.mma (M=128 co, N=128 ci, K=16 pixels) therefore reads a [16 px][128 ch] slab of
// each tile; the instruction descriptor selects MN-major for A and B.  Conv padding / stride are TMA coordinates
// (out-of-bounds pixels are zero-filled, element stride 2 for the stride-2 downsampler).
//
// One CTA owns one (128 co x 128 ci) weight tile for one kernel row kh: the 3 kw taps are three TMEM accumulators
// (384 columns) fed by the same dY tile.  The pixel axis is split over CTAs (split-K); fp32 partials go to a workspace
// that wgrad_reduce_kernel folds in a fixed order (deterministic).
#pragma once

namespace wgtc {
constexpr int kPx = 64;                    // pixels (K) per pipeline stage
constexpr int kBlk = kPx * 128;            // one [64 px][64 ch] box: 8 KB
constexpr int kDyBytes = 2 * kBlk;         // 128 co
constexpr int kXBytes = 2 * kBlk;          // 128 ci, per kw tap
__host__ __device__ constexpr int stage_bytes(int nt) { return kDyBytes + nt * kXBytes; }
__host__ __device__ constexpr int stages(int nt) { return nt == 3 ? 3 : 6; }
constexpr int kThreads = 192;              // warp 0: TMA producer, warp 1: MMA issuer, warps 2-5: epilogue
}  // namespace wgtc

struct WgradTcParams {
  CUtensorMap dy;  // dims (Cout, Wo, Ho, B), box (64, Wc, R, 1)
  CUtensorMap x;   // dims (Cin, W, H, B), box (64, Wc, R, 1) with element stride `stride`
  float* part;     // [splits][KS*KS][Cout][Cin]
  int Cin, Cout, n_ci, n_co, KS, stride, pad;
  int chunks, chunks_per_split;
  int Wc, R, chunks_per_row, rowgroups;  // chunk = R rows x Wc pixels of one image (R * Wc == 64)
};

__device__ __forceinline__ uint32_t desc_lo_mn_sw128(uint32_t smem_addr) {
  return ((smem_addr & 0x3FFFFu) >> 4) | ((uint32_t)(wgtc::kBlk >> 4) << 16);  // LBO = next 64-channel column block
}

template <int NT>
__global__ void __launch_bounds__(wgtc::kThreads, 1) conv_wgrad_tc_kernel(const __grid_constant__ WgradTcParams p) {
  constexpr int kStages = wgtc::stages(NT);
  constexpr int kStage = wgtc::stage_bytes(NT);
  constexpr int kCols = NT == 3 ? 512 : 128;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const uint32_t smem_base = smem_u32(smem);
  const uint32_t bars = smem_base + kStages * kStage;  // full[kStages], empty[kStages], done, tmem slot
  const uint32_t full0 = bars, empty0 = bars + 8 * kStages, done_bar = bars + 16 * kStages;
  const uint32_t tmem_slot = done_bar + 8;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  int tile = blockIdx.x;
  const int ci_t = tile % p.n_ci;
  tile /= p.n_ci;
  const int co_t = tile % p.n_co;
  const int kh = tile / p.n_co;
  const int co0 = co_t * 128, ci0 = ci_t * 128;
  const int c_begin = blockIdx.y * p.chunks_per_split;
  const int c_end = min(p.chunks, c_begin + p.chunks_per_split);

  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(full0 + 8 * s, 1);
      mbar_init(empty0 + 8 * s, 1);
    }
    mbar_init(done_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<kCols>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(smem + kStages * kStage + 16 * kStages + 8);
  pdl_trigger();  // PDL: barrier init and TMEM allocation above overlap the predecessor's tail
  pdl_wait();

  if (warp == 0) {
    if (elect_one_sync()) {
      tma_prefetch_desc(&p.dy);
      tma_prefetch_desc(&p.x);
      StageRing ring;
      const int per_img = p.rowgroups * p.chunks_per_row;
      for (int c = c_begin; c < c_end; ++c) {
        mbar_wait(empty0 + 8 * ring.idx, ring.phase ^ 1u);
        const int b = c / per_img;
        const int rem = c - b * per_img;
        const int rg = rem / p.chunks_per_row;
        const int x0 = (rem - rg * p.chunks_per_row) * p.Wc, y0 = rg * p.R;
        const uint32_t fb = full0 + 8 * ring.idx;
        const uint32_t dst = smem_base + ring.idx * kStage;
        mbar_expect_tx(fb, kStage);
        tma_load_4d(&p.dy, fb, dst, co0, x0, y0, b);
        tma_load_4d(&p.dy, fb, dst + wgtc::kBlk, co0 + 64, x0, y0, b);
#pragma unroll
        for (int t = 0; t < NT; ++t) {
          const uint32_t xd = dst + wgtc::kDyBytes + t * wgtc::kXBytes;
          const int xi = x0 * p.stride + t - p.pad, yi = y0 * p.stride + kh - p.pad;
          tma_load_4d(&p.x, fb, xd, ci0, xi, yi, b);
          tma_load_4d(&p.x, fb, xd + wgtc::kBlk, ci0 + 64, xi, yi, b);
        }
        ring.advance(kStages);
      }
    }
  } else if (warp == 1) {
    if (elect_one_sync()) {
      constexpr uint32_t idesc = make_idesc_bf16_f32(128, 128) | (1u << 15) | (1u << 16);  // A and B MN-major
      StageRing ring;
      for (int c = c_begin; c < c_end; ++c) {
        mbar_wait(full0 + 8 * ring.idx, ring.phase);
        tc_fence_after();
        const uint32_t sdy = smem_base + ring.idx * kStage;
        const uint32_t a_lo = desc_lo_mn_sw128(sdy);
#pragma unroll
        for (int t = 0; t < NT; ++t) {
          const uint32_t b_lo = desc_lo_mn_sw128(sdy + wgtc::kDyBytes + t * wgtc::kXBytes);
#pragma unroll
          for (int k = 0; k < wgtc::kPx / 16; ++k)  // 16 pixels = two 8-row groups = 2048 B
            umma_bf16_ss_lh(tmem_base + t * 128, a_lo + k * (2048 >> 4), b_lo + k * (2048 >> 4), idesc,
                            (c > c_begin || k > 0) ? 1u : 0u);
        }
        umma_commit(empty0 + 8 * ring.idx);
        ring.advance(kStages);
      }
      umma_commit(done_bar);
    }
  } else {
    // epilogue: TMEM lane quarter (warp % 4) = 32 output channels, one row per thread
    mbar_wait(done_bar, 0);
    tc_fence_after();
    const int q = warp & 3;
    const int co = co0 + q * 32 + lane;
    const int taps = (NT == 3) ? 9 : 1;
#pragma unroll 1
    for (int t = 0; t < NT; ++t) {
      const int tap = (NT == 3) ? kh * 3 + t : 0;
      float* row = p.part + (((int64_t)blockIdx.y * taps + tap) * p.Cout + co) * p.Cin + ci0;
#pragma unroll 1
      for (int cb = 0; cb < 4; ++cb) {
        uint32_t r[32];
        tmem_ld_32x32b_x32(tmem_base + ((uint32_t)(q * 32) << 16) + t * 128 + cb * 32, r);
        tmem_ld_wait();
        if (co < p.Cout && ci0 + cb * 32 < p.Cin) {
          float4* dst = reinterpret_cast<float4*>(row + cb * 32);
#pragma unroll
          for (int j = 0; j < 8; ++j)
            dst[j] = make_float4(__uint_as_float(r[4 * j]), __uint_as_float(r[4 * j + 1]), __uint_as_float(r[4 * j + 2]),
                                 __uint_as_float(r[4 * j + 3]));
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<kCols>(tmem_base);
}

// Upper bound of the split-K partial count for a problem (0 = geometry unsupported): sizes the workspace.
int wgrad_tc_max_splits(int B, int Ho, int Wo, int Cin, int Cout, int ksize) {
  if (Cin % 64 || Cout % 64) return 0;
  int R;
  if (Wo % 64 == 0) R = 1;
  else if (Wo < 64 && 64 % Wo == 0 && Ho % (64 / Wo) == 0) R = 64 / Wo;
  else return 0;
  const int chunks = B * (Ho / R) * (Wo / (64 / R));
  const int base = ((Cin + 127) / 128) * ((Cout + 127) / 128) * ksize;
  int s = sm_count() / base;
  if (s < 1) s = 1;
  return s > chunks ? chunks : s;
}

// Returns 0 when launched, FM_ERR_UNSUPPORTED when the geometry does not fit this kernel (the caller falls back to the
// mma.sync kernel), another code on a real error.  `*splits_out` = number of split-K partials written.
int wgrad_tc_launch(const void* dy, const void* x, float* workspace, int64_t workspace_elems, int B, int H, int W,
                    int Cin, int Cout, int ksize, int stride, int* splits_out, cudaStream_t st) {
  const int Ho = (H + stride - 1) / stride, Wo = (W + stride - 1) / stride;
  if (Cin % 64 || Cout % 64) return FM_ERR_UNSUPPORTED;
  if ((H % stride) || (W % stride)) return FM_ERR_UNSUPPORTED;
  WgradTcParams p;
  if (Wo % 64 == 0) {
    p.Wc = 64, p.R = 1;
  } else if (Wo < 64 && 64 % Wo == 0 && Ho % (64 / Wo) == 0) {
    p.Wc = Wo, p.R = 64 / Wo;
  } else {
    return FM_ERR_UNSUPPORTED;
  }
  p.chunks_per_row = Wo / p.Wc;
  p.rowgroups = Ho / p.R;
  p.chunks = B * p.rowgroups * p.chunks_per_row;
  p.n_ci = (Cin + 127) / 128;
  p.n_co = (Cout + 127) / 128;
  p.KS = ksize, p.stride = stride, p.pad = ksize / 2;
  p.Cin = Cin, p.Cout = Cout;
  const int base = p.n_ci * p.n_co * ksize;
  int s = sm_count() / base;  // one CTA per SM, a single wave
  if (s < 1) s = 1;
  if (s > p.chunks) s = p.chunks;
  p.chunks_per_split = (p.chunks + s - 1) / s;
  const int splits = (p.chunks + p.chunks_per_split - 1) / p.chunks_per_split;
  if ((int64_t)splits * ksize * ksize * Cout * Cin > workspace_elems) return FM_ERR_UNSUPPORTED;
  p.part = workspace;
  if (int e = encode_act_map(&p.dy, dy, Cout, Wo, Ho, B, p.Wc, p.R, 1, 1)) return e;
  if (int e = encode_act_map(&p.x, x, Cin, W, H, B, p.Wc, p.R, 1, stride)) return e;
  dim3 grid(base, splits);
  if (ksize == 3) {
    constexpr int smem = wgtc::stages(3) * wgtc::stage_bytes(3) + 1024 + 256;
    static bool attr = false;
    if (!attr) {
      if (int e = check_cuda(cudaFuncSetAttribute(conv_wgrad_tc_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem), "wgrad_tc attr")) return e;
      attr = true;
    }
    launch_pdl(conv_wgrad_tc_kernel<3>, grid, dim3(wgtc::kThreads), smem, st, p);
  } else {
    constexpr int smem = wgtc::stages(1) * wgtc::stage_bytes(1) + 1024 + 256;
    static bool attr = false;
    if (!attr) {
      if (int e = check_cuda(cudaFuncSetAttribute(conv_wgrad_tc_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem), "wgrad_tc attr")) return e;
      attr = true;
    }
    launch_pdl(conv_wgrad_tc_kernel<1>, grid, dim3(wgtc::kThreads), smem, st, p);
  }
  FM_LAUNCH_CHECK("conv_wgrad_tc_kernel");
  *splits_out = splits;
  return 0;
}
