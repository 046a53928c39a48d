// K2: GroupNorm (+SiLU, +scale-shift) over NHWC bf16 activations with fp32 statistics.  HBM-bound: every thread
// moves 16 B (8 channels) per access, fully coalesced along the channel-contiguous NHWC rows; per-channel partial
// sums are reduced through shared memory and land in global memory with one atomicAdd per (block, group).
// Two sources are read as a virtual channel concat (the up-path torch.cat of legacy_unet.py:150 never exists in
// memory); the apply kernel writes the concatenated, normalised, activated tensor the consumer conv reads.
// Reference ops replaced: nn.GroupNorm + nn.SiLU (src/nn/blocks/residual.py:95-96,113-116).
#include <cstring>

#include "common.cuh"

namespace fm {

constexpr int kGnThreads = 256;

__device__ __forceinline__ uint4 ld_chunk(const uint4* __restrict__ x0, int c80, const uint4* __restrict__ x1,
                                          int c81, int64_t pix, int chunk) {
  return (chunk < c80) ? x0[pix * c80 + chunk] : x1[pix * c81 + (chunk - c80)];
}

// partial[n][blk][g] = (sum, sumsq) over this block's pixel slab (no atomics: the reduction order is fixed, so the
// whole network is bit-reproducible run to run)
__global__ void __launch_bounds__(kGnThreads) gn_stats_kernel(const uint4* __restrict__ x0, int c80,
                                                             const uint4* __restrict__ x1, int c81, int64_t HW,
                                                             int groups, float* __restrict__ partial,
                                                             int64_t pix_per_block) {
  pdl_enter();
  extern __shared__ float sm[];  // [ppi][C][2]
  const int tpp = c80 + c81;     // 8-channel chunks per pixel
  const int C = tpp * 8;
  const int ppi = kGnThreads / tpp;
  const int n = blockIdx.y;
  const int chunk = threadIdx.x % tpp;
  const int prow = threadIdx.x / tpp;
  const int64_t p_begin = (int64_t)blockIdx.x * pix_per_block;
  int64_t p_end = p_begin + pix_per_block;
  if (p_end > HW) p_end = HW;

  float s[8], q[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) s[j] = q[j] = 0.f;
  if (prow < ppi) {
    const int64_t base = (int64_t)n * HW;
    for (int64_t p = p_begin + prow; p < p_end; p += ppi) {
      const uint4 u = ld_chunk(x0, c80, x1, c81, base + p, chunk);
      const float2 f0 = unpack_bf16x2(u.x), f1 = unpack_bf16x2(u.y), f2 = unpack_bf16x2(u.z), f3 = unpack_bf16x2(u.w);
      const float v[8] = {f0.x, f0.y, f1.x, f1.y, f2.x, f2.y, f3.x, f3.y};
#pragma unroll
      for (int j = 0; j < 8; ++j) { s[j] += v[j]; q[j] = fmaf(v[j], v[j], q[j]); }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      sm[((size_t)prow * C + chunk * 8 + j) * 2 + 0] = s[j];
      sm[((size_t)prow * C + chunk * 8 + j) * 2 + 1] = q[j];
    }
  }
  __syncthreads();
  // one thread per group: fold pixel rows and the group's channels
  const int cg = C / groups;
  for (int g = threadIdx.x; g < groups; g += kGnThreads) {
    float ts = 0.f, tq = 0.f;
    for (int r = 0; r < ppi; ++r)
      for (int c = g * cg; c < (g + 1) * cg; ++c) {
        ts += sm[((size_t)r * C + c) * 2 + 0];
        tq += sm[((size_t)r * C + c) * 2 + 1];
      }
    float* dst = partial + (((size_t)n * gridDim.x + blockIdx.x) * groups + g) * 2;
    dst[0] = ts;
    dst[1] = tq;
  }
}

// stats[n][g] = (mean, rstd) from the per-block partial sums, accumulated in fp64 in a fixed order
__global__ void gn_finalize_kernel(const float* __restrict__ partial, int nblk, int groups, double inv_cnt, float eps,
                                   float* __restrict__ stats) {
  pdl_enter();
  const int n = blockIdx.x;
  for (int g = threadIdx.x; g < groups; g += blockDim.x) {
    double s = 0.0, q = 0.0;
    for (int b = 0; b < nblk; ++b) {
      const float* src = partial + (((size_t)n * nblk + b) * groups + g) * 2;
      s += (double)src[0];
      q += (double)src[1];
    }
    const double mean = s * inv_cnt;
    double var = q * inv_cnt - mean * mean;
    if (var < 0.0) var = 0.0;
    stats[((size_t)n * groups + g) * 2 + 0] = (float)mean;
    stats[((size_t)n * groups + g) * 2 + 1] = (float)(1.0 / sqrt(var + (double)eps));
  }
}

// stats[n][g] = (mean, rstd) from the channel-quad partial sums the conv epilogues wrote
// (partial_s[(n*rows_s + r)][C_s/4][2]); the two sources form the virtual concat.  fp64, fixed order.
__global__ void __launch_bounds__(128) gn_finalize_partials_kernel(const float* __restrict__ p0, int rows0, int nq0,
                                                                  const float* __restrict__ p1, int rows1, int nq1,
                                                                  int groups, double inv_cnt, float eps,
                                                                  float* __restrict__ stats,
                                                                  const float* __restrict__ gamma,
                                                                  const float* __restrict__ beta,
                                                                  const float* __restrict__ scale_shift,
                                                                  int64_t ss_stride, float* __restrict__ ab) {
  pdl_enter();
  __shared__ double ss[128], sq[128];
  __shared__ float s_mean, s_rstd;
  const int g = blockIdx.x, n = blockIdx.y;
  const int qpg = (nq0 + nq1) / groups;  // quads per group
  const int q_begin = g * qpg;
  double s = 0.0, q = 0.0;
  for (int k = 0; k < qpg; ++k) {
    const int qi = q_begin + k;
    const float* base;
    int rows, nq, ql;
    if (qi < nq0) { base = p0; rows = rows0; nq = nq0; ql = qi; } else { base = p1; rows = rows1; nq = nq1; ql = qi - nq0; }
    const float* src = base + ((size_t)n * rows * nq + ql) * 2;
    for (int r = threadIdx.x; r < rows; r += 128) {
      const float2 v = *reinterpret_cast<const float2*>(src + (size_t)r * nq * 2);
      s += (double)v.x;
      q += (double)v.y;
    }
  }
  ss[threadIdx.x] = s;
  sq[threadIdx.x] = q;
  __syncthreads();
  for (int o = 64; o > 0; o >>= 1) {
    if (threadIdx.x < o) { ss[threadIdx.x] += ss[threadIdx.x + o]; sq[threadIdx.x] += sq[threadIdx.x + o]; }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const double mean = ss[0] * inv_cnt;
    double var = sq[0] * inv_cnt - mean * mean;
    if (var < 0.0) var = 0.0;
    const float rstd = (float)(1.0 / sqrt(var + (double)eps));
    if (stats != nullptr) {
      stats[((size_t)n * groups + g) * 2 + 0] = (float)mean;
      stats[((size_t)n * groups + g) * 2 + 1] = rstd;
    }
    s_mean = (float)mean;
    s_rstd = rstd;
  }
  if (ab == nullptr) return;
  __syncthreads();
  // the group's channels in the affine form consumed by the conv operand transform (same arithmetic as gn_apply)
  const int C = (nq0 + nq1) * 4, cg = qpg * 4;
  for (int c = g * cg + threadIdx.x; c < (g + 1) * cg; c += 128) {
    float a = s_rstd * gamma[c];
    float b = beta[c] - s_mean * a;
    if (scale_shift != nullptr) {
      const float sc = 1.0f + scale_shift[(size_t)n * ss_stride + c];
      const float sh = scale_shift[(size_t)n * ss_stride + C + c];
      a *= sc;
      b = b * sc + sh;
    }
    ab[((size_t)n * 2 + 0) * C + c] = a;
    ab[((size_t)n * 2 + 1) * C + c] = b;
  }
}

// ab[n][0][c], ab[n][1][c] from finished (mean, rstd) statistics
__global__ void __launch_bounds__(256) gn_affine_kernel(const float* __restrict__ stats, const float* __restrict__ gamma,
                                                       const float* __restrict__ beta,
                                                       const float* __restrict__ scale_shift, int64_t ss_stride,
                                                       int C, int groups, float* __restrict__ ab) {
  pdl_enter();
  const int n = blockIdx.x;
  const int cg = C / groups;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const int g = c / cg;
    const float mean = stats[((size_t)n * groups + g) * 2 + 0];
    const float rstd = stats[((size_t)n * groups + g) * 2 + 1];
    float a = rstd * gamma[c];
    float b = beta[c] - mean * a;
    if (scale_shift != nullptr) {
      const float sc = 1.0f + scale_shift[(size_t)n * ss_stride + c];
      const float sh = scale_shift[(size_t)n * ss_stride + C + c];
      a *= sc;
      b = b * sc + sh;
    }
    ab[((size_t)n * 2 + 0) * C + c] = a;
    ab[((size_t)n * 2 + 1) * C + c] = b;
  }
}

// channel-quad partial sums of the producer convs (fm_conv_params.gn_stats format), folded inside gn_apply_kernel
constexpr int kGnMaxFusedGroups = 64;
struct GnPartials {
  const float* p0;
  const float* p1;
  int rows0, nq0, rows1, nq1;
  double inv_cnt;
  float eps;
};

__global__ void __launch_bounds__(kGnThreads) gn_apply_kernel(const uint4* __restrict__ x0, int c80,
                                                             const uint4* __restrict__ x1, int c81, int64_t HW,
                                                             int groups, const float* __restrict__ stats,
                                                             const float* __restrict__ gamma,
                                                             const float* __restrict__ beta,
                                                             const float* __restrict__ scale_shift,
                                                             int64_t ss_stride, int silu, uint4* __restrict__ out,
                                                             int64_t pix_per_block, GnPartials pt) {
  pdl_enter();
  extern __shared__ float sm[];  // a[C], b[C]
  __shared__ float s_mr[2 * kGnMaxFusedGroups];  // (mean, rstd) per group when folded here from the conv partials
  const int tpp = c80 + c81;
  const int C = tpp * 8;
  // Block order: the producer conv wrote the tensor row chunk by row chunk over all images, so its LAST rows are what
  // the L2 still holds - walk the pixel ranges from the end, images innermost, and read them before this kernel's own
  // stream evicts them (the consumer conv then starts at the rows written last here)
  const int lin = blockIdx.y * gridDim.x + blockIdx.x;
  const int n = lin % (int)gridDim.y, bx = (int)gridDim.x - 1 - lin / (int)gridDim.y;
  float* sa = sm;
  float* sb = sm + C;
  const int cg = C / groups;
  if (pt.p0 != nullptr) {
    // Small partial tables (a few KB per sample): every block folds its sample's channel-quad partials itself, in fp64
    // and a fixed order, instead of reading statistics a finalize launch wrote - one kernel node less per GroupNorm
    // where a kernel node is what a GroupNorm costs (MNIST-sized problems, the 16x16 level).
    const int qpg = (pt.nq0 + pt.nq1) / groups;
    for (int g = threadIdx.x; g < groups; g += kGnThreads) {
      double s = 0.0, q = 0.0;
      for (int k = 0; k < qpg; ++k) {
        const int qi = g * qpg + k;
        const bool first = qi < pt.nq0;
        const int rows = first ? pt.rows0 : pt.rows1, nq = first ? pt.nq0 : pt.nq1, ql = first ? qi : qi - pt.nq0;
        const float* src = (first ? pt.p0 : pt.p1) + ((size_t)n * rows * nq + ql) * 2;
        for (int r = 0; r < rows; ++r) {
          const float2 v = __ldcg(reinterpret_cast<const float2*>(src + (size_t)r * nq * 2));
          s += (double)v.x;
          q += (double)v.y;
        }
      }
      const double mean = s * pt.inv_cnt;
      double var = q * pt.inv_cnt - mean * mean;
      if (var < 0.0) var = 0.0;
      s_mr[2 * g] = (float)mean;
      s_mr[2 * g + 1] = (float)(1.0 / sqrt(var + (double)pt.eps));
    }
    __syncthreads();
  }
  for (int c = threadIdx.x; c < C; c += kGnThreads) {
    const int g = c / cg;
    const float mean = pt.p0 != nullptr ? s_mr[2 * g] : stats[((size_t)n * groups + g) * 2 + 0];
    const float rstd = pt.p0 != nullptr ? s_mr[2 * g + 1] : stats[((size_t)n * groups + g) * 2 + 1];
    float a = rstd * gamma[c];
    float b = beta[c] - mean * a;
    if (scale_shift != nullptr) {
      const float sc = 1.0f + scale_shift[(size_t)n * ss_stride + c];
      const float sh = scale_shift[(size_t)n * ss_stride + C + c];
      a *= sc;
      b = b * sc + sh;
    }
    sa[c] = a;
    sb[c] = b;
  }
  __syncthreads();

  const int ppi = kGnThreads / tpp;
  const int chunk = threadIdx.x % tpp;
  const int prow = threadIdx.x / tpp;
  if (prow >= ppi) return;
  float a[8], b[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { a[j] = sa[chunk * 8 + j]; b[j] = sb[chunk * 8 + j]; }
  const int64_t p_begin = (int64_t)bx * pix_per_block;
  int64_t p_end = p_begin + pix_per_block;
  if (p_end > HW) p_end = HW;
  const int64_t base = (int64_t)n * HW;
  constexpr int kUnroll = 4;  // independent 16-byte loads in flight per thread
  for (int64_t p = p_begin + prow; p < p_end; p += (int64_t)ppi * kUnroll) {
    uint4 u[kUnroll];
#pragma unroll
    for (int k = 0; k < kUnroll; ++k) {
      const int64_t pk = p + (int64_t)k * ppi;
      u[k] = make_uint4(0, 0, 0, 0);
      if (pk < p_end) u[k] = ld_chunk(x0, c80, x1, c81, base + pk, chunk);
    }
#pragma unroll
    for (int k = 0; k < kUnroll; ++k) {
      const int64_t pk = p + (int64_t)k * ppi;
      if (pk >= p_end) break;
      const float2 f0 = unpack_bf16x2(u[k].x), f1 = unpack_bf16x2(u[k].y), f2 = unpack_bf16x2(u[k].z),
                   f3 = unpack_bf16x2(u[k].w);
      float v[8] = {f0.x, f0.y, f1.x, f1.y, f2.x, f2.y, f3.x, f3.y};
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        v[j] = fmaf(v[j], a[j], b[j]);
        if (silu) v[j] = silu_f(v[j]);
      }
      uint4 o;
      o.x = pack_bf16x2(v[0], v[1]); o.y = pack_bf16x2(v[2], v[3]);
      o.z = pack_bf16x2(v[4], v[5]); o.w = pack_bf16x2(v[6], v[7]);
      out[(base + pk) * tpp + chunk] = o;
    }
  }
}

static int gn_grid(int B, int64_t HW, int tpp, int64_t* pix_per_block) {
  const int ppi = kGnThreads / tpp;
  int64_t target_blocks = ((int64_t)sm_count() * 8 + B - 1) / B;  // ~8 CTAs per SM across the batch
  if (target_blocks < 1) target_blocks = 1;
  int64_t ppb = (HW + target_blocks - 1) / target_blocks;
  const int64_t min_ppb = (int64_t)ppi * 4;
  if (ppb < min_ppb) ppb = min_ppb;
  ppb = (ppb + ppi - 1) / ppi * ppi;
  *pix_per_block = ppb;
  return (int)((HW + ppb - 1) / ppb);
}

static int gn_check(const void* x0, int C0, const void* x1, int C1, int B, int64_t HW, int groups) {
  FM_REQUIRE(x0 != nullptr && C0 > 0 && C0 % 8 == 0, "groupnorm: source 0 needs C %% 8 == 0 (C0=%d)", C0);
  FM_REQUIRE((x1 == nullptr) == (C1 == 0) && C1 % 8 == 0, "groupnorm: source 1 inconsistent (C1=%d)", C1);
  FM_REQUIRE(((uintptr_t)x0 & 15) == 0 && ((uintptr_t)x1 & 15) == 0, "groupnorm: sources must be 16B aligned");
  const int C = C0 + C1;
  FM_REQUIRE(C / 8 <= kGnThreads, "groupnorm: C=%d too large (max %d)", C, kGnThreads * 8);
  FM_REQUIRE(groups > 0 && C % groups == 0, "groupnorm: groups=%d must divide C=%d", groups, C);
  FM_REQUIRE(B > 0 && HW > 0, "groupnorm: empty input");
  return 0;
}

}  // namespace fm

using namespace fm;

extern "C" int64_t fm_groupnorm_workspace_elems(int32_t B, int64_t HW, int32_t C, int32_t groups) {
  if (B <= 0 || HW <= 0 || C <= 0 || C % 8 || groups <= 0 || C / 8 > kGnThreads) return 0;
  int64_t ppb;
  const int gx = gn_grid(B, HW, C / 8, &ppb);
  return (int64_t)B * gx * groups * 2;
}

extern "C" int fm_groupnorm_stats_bf16(const void* x0, int32_t C0, const void* x1, int32_t C1, int32_t B, int64_t HW,
                                       int32_t groups, float eps, float* workspace, float* stats,
                                       fm_stream_t stream) {
  if (int e = ensure_device()) return e;
  if (int e = gn_check(x0, C0, x1, C1, B, HW, groups)) return e;
  FM_REQUIRE(stats != nullptr && workspace != nullptr, "groupnorm_stats: null stats/workspace");
  const int tpp = (C0 + C1) / 8;
  int64_t ppb;
  const int gx = gn_grid(B, HW, tpp, &ppb);
  const int ppi = kGnThreads / tpp;
  const size_t smem = (size_t)ppi * (C0 + C1) * 2 * sizeof(float);
  launch_pdl(gn_stats_kernel, dim3(gx, B), dim3(kGnThreads), smem, (cudaStream_t)stream, 
      reinterpret_cast<const uint4*>(x0), C0 / 8, reinterpret_cast<const uint4*>(x1), C1 / 8, HW, groups, workspace,
      ppb);
  FM_LAUNCH_CHECK("gn_stats_kernel");
  const double inv_cnt = 1.0 / ((double)HW * (double)((C0 + C1) / groups));
  launch_pdl(gn_finalize_kernel, dim3(B), dim3(64), 0, (cudaStream_t)stream, workspace, gx, groups, inv_cnt, eps, stats);
  FM_LAUNCH_CHECK("gn_finalize_kernel");
  return 0;
}

extern "C" int fm_groupnorm_finalize_partials(const float* p0, int32_t rows0, int32_t C0, const float* p1,
                                              int32_t rows1, int32_t C1, int32_t B, int64_t HW, int32_t groups,
                                              float eps, float* stats, fm_stream_t stream) {
  if (int e = ensure_device()) return e;
  FM_REQUIRE(p0 != nullptr && rows0 > 0 && C0 > 0 && C0 % 4 == 0, "gn_finalize_partials: bad source 0");
  FM_REQUIRE((p1 == nullptr) == (C1 == 0) && C1 % 4 == 0 && (p1 == nullptr || rows1 > 0),
             "gn_finalize_partials: bad source 1");
  const int C = C0 + C1;
  FM_REQUIRE(groups > 0 && C % groups == 0 && (C / groups) % 4 == 0,
             "gn_finalize_partials: channels per group (%d/%d) must be a multiple of 4", C, groups);
  FM_REQUIRE(B > 0 && B <= 65535 && HW > 0 && stats != nullptr, "gn_finalize_partials: bad argument");
  const double inv_cnt = 1.0 / ((double)HW * (double)(C / groups));
  launch_pdl(gn_finalize_partials_kernel, dim3(groups, B), dim3(128), 0, (cudaStream_t)stream, 
      p0, rows0, C0 / 4, p1, rows1, C1 / 4, groups, inv_cnt, eps, stats, nullptr, nullptr, nullptr, 0, nullptr);
  FM_LAUNCH_CHECK("gn_finalize_partials_kernel");
  return 0;
}

extern "C" int fm_groupnorm_finalize_partials_affine(const float* p0, int32_t rows0, int32_t C0, const float* p1,
                                                     int32_t rows1, int32_t C1, int32_t B, int64_t HW, int32_t groups,
                                                     float eps, const float* gamma, const float* beta,
                                                     const float* scale_shift, int64_t ss_stride, float* stats,
                                                     float* ab, fm_stream_t stream) {
  if (int e = ensure_device()) return e;
  FM_REQUIRE(p0 != nullptr && rows0 > 0 && C0 > 0 && C0 % 4 == 0, "gn_finalize_partials_affine: bad source 0");
  FM_REQUIRE((p1 == nullptr) == (C1 == 0) && C1 % 4 == 0 && (p1 == nullptr || rows1 > 0),
             "gn_finalize_partials_affine: bad source 1");
  const int C = C0 + C1;
  FM_REQUIRE(groups > 0 && C % groups == 0 && (C / groups) % 4 == 0,
             "gn_finalize_partials_affine: channels per group (%d/%d) must be a multiple of 4", C, groups);
  FM_REQUIRE(B > 0 && B <= 65535 && HW > 0 && gamma && beta && ab, "gn_finalize_partials_affine: bad argument");
  const double inv_cnt = 1.0 / ((double)HW * (double)(C / groups));
  launch_pdl(gn_finalize_partials_kernel, dim3(groups, B), dim3(128), 0, (cudaStream_t)stream, 
      p0, rows0, C0 / 4, p1, rows1, C1 / 4, groups, inv_cnt, eps, stats, gamma, beta, scale_shift, ss_stride, ab);
  FM_LAUNCH_CHECK("gn_finalize_partials_kernel");
  return 0;
}

extern "C" int fm_groupnorm_affine_f32(const float* stats, const float* gamma, const float* beta,
                                       const float* scale_shift, int64_t ss_stride, int32_t B, int32_t C,
                                       int32_t groups, float* ab, fm_stream_t stream) {
  if (int e = ensure_device()) return e;
  FM_REQUIRE(stats && gamma && beta && ab, "groupnorm_affine: null pointer");
  FM_REQUIRE(B > 0 && C > 0 && groups > 0 && C % groups == 0, "groupnorm_affine: bad shape B=%d C=%d groups=%d", B, C,
             groups);
  launch_pdl(gn_affine_kernel, dim3(B), dim3(256), 0, (cudaStream_t)stream, stats, gamma, beta, scale_shift, ss_stride, C, groups, ab);
  FM_LAUNCH_CHECK("gn_affine_kernel");
  return 0;
}

extern "C" int fm_groupnorm_apply_bf16(const void* x0, int32_t C0, const void* x1, int32_t C1, int32_t B, int64_t HW,
                                       int32_t groups, const float* stats, const float* gamma,
                                       const float* beta, const float* scale_shift, int64_t ss_stride, int32_t silu,
                                       void* out, fm_stream_t stream) {
  if (int e = ensure_device()) return e;
  if (int e = gn_check(x0, C0, x1, C1, B, HW, groups)) return e;
  FM_REQUIRE(stats && gamma && beta && out, "groupnorm_apply: null pointer");
  FM_REQUIRE(((uintptr_t)out & 15) == 0, "groupnorm_apply: out must be 16B aligned");
  const int tpp = (C0 + C1) / 8;
  int64_t ppb;
  const int gx = gn_grid(B, HW, tpp, &ppb);
  const size_t smem = (size_t)(C0 + C1) * 2 * sizeof(float);
  GnPartials none;
  memset(&none, 0, sizeof(none));
  launch_pdl(gn_apply_kernel, dim3(gx, B), dim3(kGnThreads), smem, (cudaStream_t)stream, 
      reinterpret_cast<const uint4*>(x0), C0 / 8, reinterpret_cast<const uint4*>(x1), C1 / 8, HW, groups, stats,
      gamma, beta, scale_shift, ss_stride, silu, reinterpret_cast<uint4*>(out), ppb, none);
  FM_LAUNCH_CHECK("gn_apply_kernel");
  return 0;
}

extern "C" int fm_groupnorm_apply_partials_supported(int32_t rows0, int32_t C0, int32_t rows1, int32_t C1,
                                                     int32_t groups) {
  const int C = C0 + C1;
  if (rows0 <= 0 || C0 <= 0 || C0 % 8 || C1 % 8 || (C1 > 0 && rows1 <= 0) || groups <= 0 || groups > kGnMaxFusedGroups ||
      C % groups || (C / groups) % 4)
    return 0;
  // every block folds the whole per-sample table: worth it only while that is a few KB
  return ((int64_t)rows0 * C0 + (int64_t)rows1 * C1) * 2 <= 16 * 1024 ? 1 : 0;
}

extern "C" int fm_groupnorm_apply_partials_bf16(const void* x0, int32_t C0, const void* x1, int32_t C1, int32_t B,
                                                int64_t HW, int32_t groups, const float* p0, int32_t rows0,
                                                const float* p1, int32_t rows1, float eps, const float* gamma,
                                                const float* beta, const float* scale_shift, int64_t ss_stride,
                                                int32_t silu, void* out, fm_stream_t stream) {
  if (int e = ensure_device()) return e;
  if (int e = gn_check(x0, C0, x1, C1, B, HW, groups)) return e;
  FM_REQUIRE(p0 && gamma && beta && out && (p1 != nullptr) == (C1 > 0), "groupnorm_apply_partials: null pointer");
  FM_REQUIRE(((uintptr_t)out & 15) == 0, "groupnorm_apply_partials: out must be 16B aligned");
  if (!fm_groupnorm_apply_partials_supported(rows0, C0, rows1, C1, groups)) {
    set_error("groupnorm_apply_partials: table too large or groups not made of channel quads "
              "(fm_groupnorm_apply_partials_supported)");
    return FM_ERR_UNSUPPORTED;
  }
  const int tpp = (C0 + C1) / 8;
  int64_t ppb;
  const int gx = gn_grid(B, HW, tpp, &ppb);
  const size_t smem = (size_t)(C0 + C1) * 2 * sizeof(float);
  GnPartials pt;
  pt.p0 = p0, pt.p1 = p1, pt.rows0 = rows0, pt.nq0 = C0 / 4, pt.rows1 = rows1, pt.nq1 = C1 / 4;
  pt.inv_cnt = 1.0 / ((double)HW * (double)((C0 + C1) / groups));
  pt.eps = eps;
  launch_pdl(gn_apply_kernel, dim3(gx, B), dim3(kGnThreads), smem, (cudaStream_t)stream,
             reinterpret_cast<const uint4*>(x0), C0 / 8, reinterpret_cast<const uint4*>(x1), C1 / 8, HW, groups,
             (const float*)nullptr, gamma, beta, scale_shift, ss_stride, silu, reinterpret_cast<uint4*>(out), ppb, pt);
  FM_LAUNCH_CHECK("gn_apply_kernel");
  return 0;
}
