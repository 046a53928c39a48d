// Backward / optimiser kernels of the training step (reference: the autograd graph PyTorch builds for
// `flow_matching_lib.py:138-182`, i.e. F.conv2d / F.group_norm / SiLU / SDPA / F.linear / F.mse_loss backward + AdamW).
//
//   conv wgrad    dW[co][ci][kh][kw] = sum_p dY[p][co] * X[p (+) tap][ci]  -- a GEMM whose K dimension is the pixel axis.
//                 Both operands are pixel-major in HBM (NHWC), so they are "MN-major" for the tensor cores; this first
//                 version runs them through mma.sync.m16n8k16 with ldmatrix.trans (split-K over pixel ranges, fp32
//                 partials in a workspace, fixed-order second-stage reduction: deterministic, no atomics).
//   conv dgrad    is the forward implicit-GEMM kernel (conv_igemm.cu) run on dY with flipped/transposed weights; the
//                 helpers here are the zero-insertion that turns a stride-2 dgrad into a stride-1 conv and the 2x2
//                 sum-pool that is the backward of the nearest-2x upsample.
//   GroupNorm+SiLU backward, attention backward (fp32 CUDA cores, one CTA per (sample, head)), stem/head conv
//   backward, column sums (bias / time-embedding gradients), fused MSE loss + gradient, flat AdamW.
#include <stdlib.h>

#include "common.cuh"

namespace fm {

// self-resetting tickets for "the last CTA finishes the job" (caller-owned int32 buffer of FM_TICKET_INTS, zero on entry
// and on exit): [0, kTicketGlobal) per-sample counters, one global counter, then per-column-block counters (colsum)
constexpr int kTicketGlobal = 2048;
constexpr int kTicketCols = 2049;
constexpr int kTicketInts = 4096;

// =============================================================================================================
// generic fixed-order reduction of partials: out[o][i] = sum_p in[o][p][i]
// =============================================================================================================
// out2 (optional): results with index >= split go to out2[idx - split] (two destinations for one reduction, e.g. the
// dgamma / dbeta rows written straight into two parameter-gradient slices)
// grid (ceil(inner/32), outer), block (32 columns, 8 lanes): lane y folds parts y, y+8, ... in order, then the eight
// lane sums are folded in order (a thread-per-output loop over hundreds of partial rows is a chain of dependent L2
// round trips: 230 us for the stem's 592 partial rows)
__global__ void __launch_bounds__(256) reduce_partials_kernel(const float* __restrict__ in, float* __restrict__ out,
                                                               int parts, int64_t inner, float* __restrict__ out2,
                                                               int64_t split) {
  pdl_enter();
  __shared__ float sh[8][33];
  const int64_t i = (int64_t)blockIdx.x * 32 + threadIdx.x;
  const int64_t o = blockIdx.y;
  float acc = 0.f;
  if (i < inner) {
    const float* src = in + (o * parts) * inner + i;
#pragma unroll 4
    for (int p = threadIdx.y; p < parts; p += 8) acc += src[(int64_t)p * inner];
  }
  sh[threadIdx.y][threadIdx.x] = acc;
  __syncthreads();
  if (threadIdx.y == 0 && i < inner) {
    float t = 0.f;
    for (int y = 0; y < 8; ++y) t += sh[y][threadIdx.x];
    const int64_t idx = o * inner + i;
    if (out2 != nullptr && idx >= split) out2[idx - split] = t;
    else out[idx] = t;
  }
}

static int launch_reduce(const float* in, float* out, int64_t outer, int parts, int64_t inner, cudaStream_t st,
                         float* out2 = nullptr, int64_t split = 0) {
  if (outer * inner == 0) return 0;
  launch_pdl(reduce_partials_kernel, dim3((unsigned)((inner + 31) / 32), (unsigned)outer), dim3(32, 8), 0, st, 
      in, out, parts, inner, out2, split);
  FM_LAUNCH_CHECK("reduce_partials_kernel");
  return 0;
}

// =============================================================================================================
// conv wgrad (mma.sync m16n8k16, bf16 x bf16 -> fp32)
// =============================================================================================================
namespace wg {
constexpr int kCo = 128;      // M tile (output channels)
constexpr int kCi = 64;       // N tile (input channels)
constexpr int kPx = 64;       // K chunk (output pixels) per pipeline stage
constexpr int kStages = 3;
constexpr int kDyRow = (kCo + 8) * 2;  // padded smem row pitch in bytes (conflict-free ldmatrix)
constexpr int kXRow = (kCi + 8) * 2;
constexpr int kDyBytes = kPx * kDyRow;
constexpr int kXBytes = kPx * kXRow;   // per kw tap
__host__ __device__ constexpr int stage_bytes(int ntaps) { return kDyBytes + ntaps * kXBytes; }
}  // namespace wg

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, bool valid) {
  const int sz = valid ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
               : "r"(addr));
}
__device__ __forceinline__ void mma_bf16_16816(float* c, const uint32_t* a, uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

struct WgradParams {
  const __nv_bfloat16* dy;  // [B][Ho][Wo][Cout]
  const __nv_bfloat16* x;   // [B][H][W][Cin]
  float* part;              // [splits][KS*KS][Cout][Cin]
  int B, H, W, Ho, Wo, Cin, Cout, stride, pad;
  int n_ci, n_co, chunks, chunks_per_split;
  int64_t total_px;
};

// grid.x = n_ci * n_co * KS (kh), grid.y = splits.  NT = taps along kw handled by one CTA (3 for 3x3, 1 for 1x1).
template <int NT>
__global__ void __launch_bounds__(256, 1) conv_wgrad_mma_kernel(const WgradParams p) {
  extern __shared__ __align__(128) uint8_t smem[];
  const uint32_t smem_base = smem_u32(smem);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int wm = warp & 3, wn = warp >> 2;  // 4 x 2 warps: warp tile 32 (co) x 32 (ci) x NT taps
  int tile = blockIdx.x;
  const int ci_t = tile % p.n_ci;
  tile /= p.n_ci;
  const int co_t = tile % p.n_co;
  const int kh = tile / p.n_co;
  const int co0 = co_t * wg::kCo, ci0 = ci_t * wg::kCi;
  const int c_begin = blockIdx.y * p.chunks_per_split;
  const int c_end = min(p.chunks, c_begin + p.chunks_per_split);
  const int hw_o = p.Ho * p.Wo;
  constexpr int kStage = wg::stage_bytes(NT);

  float acc[NT][2][4][4];
#pragma unroll
  for (int t = 0; t < NT; ++t)
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int k = 0; k < 4; ++k) acc[t][i][j][k] = 0.f;

  auto load_stage = [&](int chunk, int slot) {
    const uint32_t sdy = smem_base + slot * kStage;
    const uint32_t sx = sdy + wg::kDyBytes;
    const int64_t px0 = (int64_t)chunk * wg::kPx;  // flattened output pixel (b, yo, xo)
    // dY tile: 64 pixels x 128 co = 64 x 16 chunks of 16 B
#pragma unroll
    for (int it = 0; it < (wg::kPx * (wg::kCo / 8)) / 256; ++it) {
      const int idx = it * 256 + tid;
      const int r = idx >> 4, cc = idx & 15;
      const int co = co0 + cc * 8;
      const bool ok = co < p.Cout && px0 + r < p.total_px;
      const __nv_bfloat16* src = ok ? p.dy + (px0 + r) * p.Cout + co : p.dy;
      cp_async16(sdy + r * wg::kDyRow + cc * 16, src, ok);
    }
    // X tiles, one per kw tap: 64 pixels x 64 ci = 64 x 8 chunks of 16 B
#pragma unroll
    for (int t = 0; t < NT; ++t) {
#pragma unroll
      for (int it = 0; it < (wg::kPx * (wg::kCi / 8)) / 256; ++it) {
        const int idx = it * 256 + tid;
        const int r = idx >> 3, cc = idx & 7;
        const int64_t pp = px0 + r;
        const int b = (int)(pp / hw_o);
        const int rem = (int)(pp - (int64_t)b * hw_o);
        const int yo = rem / p.Wo, xo = rem - yo * p.Wo;
        const int yi = yo * p.stride + kh - p.pad, xi = xo * p.stride + t - p.pad;
        const int ci = ci0 + cc * 8;
        const bool ok = ci < p.Cin && yi >= 0 && yi < p.H && xi >= 0 && xi < p.W && pp < p.total_px;
        const __nv_bfloat16* src = ok ? p.x + (((int64_t)b * p.H + yi) * p.W + xi) * p.Cin + ci : p.x;
        cp_async16(sx + t * wg::kXBytes + r * wg::kXRow + cc * 16, src, ok);
      }
    }
  };

  // prologue
#pragma unroll
  for (int s = 0; s < wg::kStages - 1; ++s) {
    if (c_begin + s < c_end) load_stage(c_begin + s, s);
    cp_async_commit();
  }
  const int lj = lane >> 3, lr = lane & 7;
  for (int c = c_begin; c < c_end; ++c) {
    const int slot = (c - c_begin) % wg::kStages;
    cp_async_wait<wg::kStages - 2>();
    __syncthreads();
    {
      const int nc = c + wg::kStages - 1;
      if (nc < c_end) load_stage(nc, (nc - c_begin) % wg::kStages);
      cp_async_commit();
    }
    const uint32_t sdy = smem_base + slot * kStage;
    const uint32_t sx = sdy + wg::kDyBytes;
#pragma unroll
    for (int ks = 0; ks < wg::kPx / 16; ++ks) {
      uint32_t a[2][4];
#pragma unroll
      for (int mi = 0; mi < 2; ++mi) {
        // matrix j: m-half = j & 1, k-half = j >> 1; stored rows are pixels (k), columns are channels (m)
        const int px = ks * 16 + (lj >> 1) * 8 + lr;
        const int co = wm * 32 + mi * 16 + (lj & 1) * 8;
        ldsm_x4_t(sdy + px * wg::kDyRow + co * 2, a[mi][0], a[mi][1], a[mi][2], a[mi][3]);
      }
#pragma unroll
      for (int t = 0; t < NT; ++t) {
#pragma unroll
        for (int nb = 0; nb < 2; ++nb) {
          // matrix j: k-half = j & 1, n-block = j >> 1
          const int px = ks * 16 + (lj & 1) * 8 + lr;
          const int ci = wn * 32 + nb * 16 + (lj >> 1) * 8;
          uint32_t b0, b1, b2, b3;
          ldsm_x4_t(sx + t * wg::kXBytes + px * wg::kXRow + ci * 2, b0, b1, b2, b3);
#pragma unroll
          for (int mi = 0; mi < 2; ++mi) {
            mma_bf16_16816(acc[t][mi][nb * 2 + 0], a[mi], b0, b1);
            mma_bf16_16816(acc[t][mi][nb * 2 + 1], a[mi], b2, b3);
          }
        }
      }
    }
  }
  cp_async_wait<0>();

  // partial tile -> workspace [split][tap][Cout][Cin]
  const int g = lane >> 2, tq = lane & 3;
  const int ks_total = (NT == 3) ? 9 : 1;
#pragma unroll
  for (int t = 0; t < NT; ++t) {
    const int tap = (NT == 3) ? kh * 3 + t : 0;
    float* base = p.part + ((int64_t)blockIdx.y * ks_total + tap) * p.Cout * p.Cin;
#pragma unroll
    for (int mi = 0; mi < 2; ++mi)
#pragma unroll
      for (int nj = 0; nj < 4; ++nj) {
        const int ci = ci0 + wn * 32 + nj * 8 + tq * 2;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int co = co0 + wm * 32 + mi * 16 + g + h * 8;
          if (co < p.Cout && ci < p.Cin)
            *reinterpret_cast<float2*>(base + (int64_t)co * p.Cin + ci) =
                make_float2(acc[t][mi][nj][h * 2 + 0], acc[t][mi][nj][h * 2 + 1]);
        }
      }
  }
}

// dw[co][c_begin + ci][tap] = sum_split part[split][tap][co][ci]
// One block = one output channel x 64 input channels: the partials are read tap row by tap row (64 consecutive floats
// each), transposed through shared memory, and written as ONE contiguous run of 64*taps floats (the OIHW slice) -
// a thread-per-element version wrote with a stride of `taps` floats and ran at ~1 TB/s.
constexpr int kWrCi = 64;
__global__ void __launch_bounds__(192) wgrad_reduce_kernel(const float* __restrict__ part, float* __restrict__ dw,
                                                            int splits, int taps, int Cout, int Cin, int cin_total,
                                                            int c_begin) {
  pdl_enter();
  __shared__ float tile[kWrCi * 9];
  const int co = blockIdx.y, ci0 = blockIdx.x * kWrCi;
  const int nci = Cin - ci0 < kWrCi ? Cin - ci0 : kWrCi;
  const int64_t per = (int64_t)taps * Cout * Cin;
  for (int item = threadIdx.x; item < taps * kWrCi; item += blockDim.x) {
    const int tap = item / kWrCi, i = item - tap * kWrCi;
    if (i >= nci) continue;
    const float* src = part + ((int64_t)tap * Cout + co) * Cin + ci0 + i;
    float acc = 0.f;
    int sp = 0;
    for (; sp + 8 <= splits; sp += 8) {  // eight independent loads in flight, summed in the same fixed order
      float v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) v[u] = __ldcs(src + (int64_t)(sp + u) * per);
#pragma unroll
      for (int u = 0; u < 8; ++u) acc += v[u];
    }
    for (; sp < splits; ++sp) acc += __ldcs(src + (int64_t)sp * per);
    tile[i * taps + tap] = acc;
  }
  __syncthreads();
  float* dst = dw + ((int64_t)co * cin_total + c_begin + ci0) * taps;
  for (int j = threadIdx.x; j < nci * taps; j += blockDim.x) dst[j] = tile[j];
}

static void launch_wgrad_reduce(const float* part, float* dw, int splits, int taps, int Cout, int Cin, int cin_total,
                                int c_begin, cudaStream_t st) {
  launch_pdl(wgrad_reduce_kernel, dim3((Cin + kWrCi - 1) / kWrCi, Cout), dim3(192), 0, st, part, dw, splits, taps, Cout, Cin,
                                                                             cin_total, c_begin);
}

static int wgrad_plan(int B, int Ho, int Wo, int Cin, int Cout, int ksize, int* splits, int* chunks_per_split,
                      int* chunks, int* base) {
  const int64_t px = (int64_t)B * Ho * Wo;
  if (px <= 0) return FM_ERR_UNSUPPORTED;
  *chunks = (int)((px + wg::kPx - 1) / wg::kPx);  // a ragged last chunk is zero-filled
  const int n_ci = (Cin + wg::kCi - 1) / wg::kCi, n_co = (Cout + wg::kCo - 1) / wg::kCo;
  *base = n_ci * n_co * ksize;
  int s = (2 * sm_count() + *base - 1) / *base;  // about two waves of CTAs
  if (s > *chunks) s = *chunks;
  if (s < 1) s = 1;
  *chunks_per_split = (*chunks + s - 1) / s;
  *splits = (*chunks + *chunks_per_split - 1) / *chunks_per_split;
  return 0;
}

// =============================================================================================================
// column sums: dY [B][HW][C] bf16 -> out[B][C] fp32 (per-sample), two stages
// =============================================================================================================
// thread = (8-channel group, row lane) as in the GroupNorm passes: 16-byte loads, four in flight per thread
__global__ void __launch_bounds__(256) colsum_partial_kernel(const uint4* __restrict__ dy, float* __restrict__ part,
                                                              int64_t HW, int C8, int lanes, int rows_per_blk) {
  pdl_enter();
  extern __shared__ float cred[];  // [lanes][C]
  const int b = blockIdx.y, blk = blockIdx.x, nblk = gridDim.x;
  const int c8 = threadIdx.x % C8, lane = threadIdx.x / C8;
  const int C = C8 * 8;
  if (lane < lanes) {
    float cs[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    const int64_t r0 = (int64_t)blk * rows_per_blk;
    const int64_t r1 = r0 + rows_per_blk < HW ? r0 + rows_per_blk : HW;
    const uint4* src = dy + ((int64_t)b * HW) * C8 + c8;
    constexpr int kU = 4;
    for (int64_t r = r0 + lane; r < r1; r += (int64_t)lanes * kU) {
      uint4 v[kU];
#pragma unroll
      for (int u = 0; u < kU; ++u) {
        const int64_t ru = r + (int64_t)u * lanes;
        v[u] = ru < r1 ? __ldg(src + ru * C8) : make_uint4(0, 0, 0, 0);
      }
#pragma unroll
      for (int u = 0; u < kU; ++u) {
        const uint32_t w[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float2 f = unpack_bf16x2(w[k]);
          cs[2 * k] += f.x;
          cs[2 * k + 1] += f.y;
        }
      }
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) cred[(lane * C8 + c8) * 8 + k] = cs[k];
  }
  __syncthreads();
  for (int idx = threadIdx.x; idx < C; idx += blockDim.x) {
    float acc = 0.f;
    for (int l = 0; l < lanes; ++l) acc += cred[l * C + idx];
    part[((int64_t)b * nblk + blk) * C + idx] = acc;
  }
}

// stage 2: out[b][c] = sum_blk part[b][blk][c] (row length ld; fixed order); total[c] = sum_b out[b][c] (fixed order,
// folded by whichever block of the column group finishes last).  grid (ceil(C/32), B), block (32 channels, 8 lanes).
__global__ void __launch_bounds__(256) colsum_final_kernel(const float* __restrict__ part, float* out,
                                                            float* __restrict__ total, int B, int nblk, int C, int ld,
                                                            int* tickets) {
  pdl_enter();
  __shared__ float sh[8][33];
  __shared__ int s_last;
  const int c = blockIdx.x * 32 + threadIdx.x, b = blockIdx.y;
  float acc = 0.f;
  if (c < C)
    for (int k = threadIdx.y; k < nblk; k += 8) acc += __ldg(part + ((int64_t)b * nblk + k) * ld + c);
  sh[threadIdx.y][threadIdx.x] = acc;
  __syncthreads();
  if (threadIdx.y == 0 && c < C) {
    float t = 0.f;
    for (int y = 0; y < 8; ++y) t += sh[y][threadIdx.x];
    out[(int64_t)b * C + c] = t;
  }
  if (total == nullptr) return;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0 && threadIdx.y == 0) s_last = atomicAdd(tickets + kTicketCols + blockIdx.x, 1) == B - 1;
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  if (threadIdx.y == 0 && c < C) {
    float t = 0.f;
    for (int s = 0; s < B; ++s) t += __ldcg(out + (int64_t)s * C + c);
    total[c] = t;
  }
  if (threadIdx.x == 0 && threadIdx.y == 0) tickets[kTicketCols + blockIdx.x] = 0;
}

// =============================================================================================================
// zero insertion (stride-2 dgrad) and 2x2 sum-pool (nearest-2x upsample backward), bf16 NHWC, 8 channels / thread
// =============================================================================================================
__global__ void __launch_bounds__(256) zero_insert2x_kernel(const uint4* __restrict__ in, uint4* __restrict__ out,
                                                             int B, int H, int W, int C8) {
  pdl_enter();
  const int64_t total = (int64_t)B * 2 * H * 2 * W * C8;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(idx % C8);
    int64_t r = idx / C8;
    const int x = (int)(r % (2 * W));
    r /= 2 * W;
    const int y = (int)(r % (2 * H));
    const int b = (int)(r / (2 * H));
    uint4 v = make_uint4(0, 0, 0, 0);
    if (!(x & 1) && !(y & 1)) v = in[(((int64_t)b * H + (y >> 1)) * W + (x >> 1)) * C8 + c];
    out[idx] = v;
  }
}

__global__ void __launch_bounds__(256) sumpool2x2_kernel(const uint4* __restrict__ in, uint4* __restrict__ out,
                                                          int B, int H, int W, int C8) {
  pdl_enter();  // H, W: output size
  const int64_t total = (int64_t)B * H * W * C8;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(idx % C8);
    int64_t r = idx / C8;
    const int x = (int)(r % W);
    r /= W;
    const int y = (int)(r % H);
    const int b = (int)(r / H);
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll
    for (int dy = 0; dy < 2; ++dy)
#pragma unroll
      for (int dx = 0; dx < 2; ++dx) {
        const uint4 v = in[(((int64_t)b * 2 * H + 2 * y + dy) * (2 * W) + 2 * x + dx) * C8 + c];
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float2 f = unpack_bf16x2(w[k]);
          acc[2 * k] += f.x;
          acc[2 * k + 1] += f.y;
        }
      }
    out[idx] = make_uint4(pack_bf16x2(acc[0], acc[1]), pack_bf16x2(acc[2], acc[3]), pack_bf16x2(acc[4], acc[5]),
                          pack_bf16x2(acc[6], acc[7]));
  }
}

// =============================================================================================================
// GroupNorm (+ scale-shift) (+ SiLU) backward
// =============================================================================================================
__device__ __forceinline__ float silu_grad_f(float n) {
  const float s = 1.f / (1.f + __expf(-n));
  return s * (1.f + n * (1.f - s));
}

// coefficient table per (sample, channel), written by the partial pass's finalize, read by the apply pass:
// a, b (forward affine n = a*x + b), P, Q (dx = a*dn + P + Q*x)
constexpr int kGnTab = 4;

// SiLU'(n) with one special-function op: sigma(n) = 0.5 + 0.5 * tanh(n / 2)
__device__ __forceinline__ float silu_grad_fast(float n) {
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(0.5f * n));
  const float s = fmaf(0.5f, t, 0.5f);
  return s * fmaf(n, 1.f - s, 1.f);
}

// SiLU'(2h) for a channel pair from h = n/2: t = tanh(h), sigma = (1 + t)/2, SiLU' = sigma * (1 + h*(1 - t))
__device__ __forceinline__ float2 silu_grad2(float2 h) {
  float2 t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t.x) : "f"(h.x));
  asm("tanh.approx.f32 %0, %1;" : "=f"(t.y) : "f"(h.y));
  const float2 one = make_float2(1.f, 1.f), half = make_float2(0.5f, 0.5f);
  const float2 w = ffma2(t, make_float2(-1.f, -1.f), one);  // 1 - t
  const float2 u = ffma2(h, w, one);
  return fmul2(ffma2(t, half, half), u);
}

// forward affine of channel c of sample b: n = a*x + b (GroupNorm, optional scale-shift folded in)
__device__ __forceinline__ void gn_affine_coef(const float* __restrict__ stats, const float* __restrict__ gamma,
                                               const float* __restrict__ beta, const float* __restrict__ ss,
                                               int64_t ss_stride, int b, int c, int C, int cpg, int groups, float& a,
                                               float& bb, float& mean, float& rstd) {
  const int g = c / cpg;
  mean = stats[((int64_t)b * groups + g) * 2 + 0];
  rstd = stats[((int64_t)b * groups + g) * 2 + 1];
  float ga = gamma[c], be = beta[c];
  if (ss != nullptr) {
    const float sc = 1.f + ss[b * ss_stride + c];
    ga *= sc;
    be = fmaf(be, sc, ss[b * ss_stride + C + c]);
  }
  a = rstd * ga;
  bb = fmaf(-mean, a, be);
}

struct GnBwdTail {  // what the last CTA of a sample / of the launch needs
  const float* stats;
  const float* gamma;
  const float* beta;
  const float* ss;
  int64_t ss_stride;
  int groups;
  float inv_n;
  float* tab;       // [B][C][kGnTab]
  float* dgb_part;  // [B][2][C] per-sample (dgamma, dbeta)
  float* dss;       // [B][2C] or NULL
  float* dgamma;    // [2][C], or [C] when dbeta is given
  float* dbeta;     // [C] or NULL
  int* tickets;
  int B;
};

// Thread layout of the two streaming passes: a block covers all C channels of a run of rows; thread = (8-channel
// group c8 = tid % C8, row lane = tid / C8), so its per-channel coefficients stay in registers for the whole run and
// every global access is a 16-byte vector.
// pass 1: S1[b][c] = sum_p dn, S2[b][c] = sum_p dn * x (partials per row block; xhat is applied to the folded sums).
// The LAST block of a sample to finish folds that sample's partials in a fixed order, forms the group sums and
// writes the apply pass's coefficient table and the per-sample parameter gradients; the last block of the launch
// folds those over the samples into dgamma / dbeta.  (Tickets only decide WHO folds; the order is fixed.)
__global__ void __launch_bounds__(256, 2) gn_bwd_partial_kernel(const uint4* __restrict__ x0, int C0_8,
                                                              const uint4* __restrict__ x1, const uint4* __restrict__ da,
                                                              float* part, int64_t HW, int C8, int lanes,
                                                              int rows_per_blk, int silu, GnBwdTail tl) {
  pdl_enter();
  extern __shared__ float red[];  // [lanes][C8*16], reused by the tail as [4][C]
  __shared__ int s_last;
  // Block order: dOut was just written by the data-gradient conv row chunk by row chunk over all images, so its last
  // rows are what the L2 still holds: walk the row blocks from the end, images innermost (the apply pass then walks
  // them from the start and finds the rows this pass read last)
  const int nblk = gridDim.x, lin = blockIdx.y * nblk + blockIdx.x;
  const int b = lin % (int)gridDim.y, blk = nblk - 1 - lin / (int)gridDim.y;
  const int c8 = threadIdx.x % C8, lane = threadIdx.x / C8;
  const int C = C8 * 8;
  const int cpg = C / tl.groups;
  float s1[8], s2[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) s1[k] = s2[k] = 0.f;
  if (lane < lanes) {
    // forward affine pre-halved: h = (a*x + b) / 2, so that SiLU'(2h) = sigma*(1 + h*(1 - t)), t = tanh(h),
    // sigma = (1 + t)/2 - all of it on packed fp32 pairs (FFMA2 / FMUL2 / FADD2), two channels per instruction
    float2 cah[4], cbh[4];
    if (silu) {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        float a0, b0, a1, b1, mean, rstd;
        gn_affine_coef(tl.stats, tl.gamma, tl.beta, tl.ss, tl.ss_stride, b, c8 * 8 + 2 * k, C, cpg, tl.groups, a0, b0,
                       mean, rstd);
        gn_affine_coef(tl.stats, tl.gamma, tl.beta, tl.ss, tl.ss_stride, b, c8 * 8 + 2 * k + 1, C, cpg, tl.groups, a1,
                       b1, mean, rstd);
        cah[k] = make_float2(0.5f * a0, 0.5f * a1);
        cbh[k] = make_float2(0.5f * b0, 0.5f * b1);
      }
    }
    float2 p1[4], p2[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) p1[k] = p2[k] = make_float2(0.f, 0.f);
    const int64_t r0 = (int64_t)blk * rows_per_blk;
    const int64_t r1 = r0 + rows_per_blk < HW ? r0 + rows_per_blk : HW;
    // the input is the virtual channel concat of (x0 [C0], x1 [C - C0]); this thread's 8 channels lie in one of them
    const int xC8 = c8 < C0_8 ? C0_8 : C8 - C0_8;
    const uint4* xs = (c8 < C0_8 ? x0 + c8 : x1 + (c8 - C0_8)) + ((int64_t)b * HW) * xC8;
    const uint4* ds = da + ((int64_t)b * HW) * C8 + c8;
    constexpr int kU = 4;  // independent 16-byte loads in flight per stream and thread
    for (int64_t r = r0 + lane; r < r1; r += (int64_t)lanes * kU) {
      uint4 xv[kU], dv[kU];
#pragma unroll
      for (int u = 0; u < kU; ++u) {
        const int64_t ru = r + (int64_t)u * lanes;
        xv[u] = dv[u] = make_uint4(0, 0, 0, 0);  // zero gradient rows contribute nothing
        if (ru < r1) {
          xv[u] = __ldg(xs + ru * xC8);
          dv[u] = __ldg(ds + ru * C8);
        }
      }
#pragma unroll
      for (int u = 0; u < kU; ++u) {
        const uint32_t xw[4] = {xv[u].x, xv[u].y, xv[u].z, xv[u].w}, dw[4] = {dv[u].x, dv[u].y, dv[u].z, dv[u].w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float2 xf = bf16x2_as_f32x2(xw[k]);
          float2 dn = bf16x2_as_f32x2(dw[k]);
          if (silu) dn = fmul2(dn, silu_grad2(ffma2(cah[k], xf, cbh[k])));
          p1[k] = fadd2(p1[k], dn);
          p2[k] = ffma2(dn, xf, p2[k]);
        }
      }
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      s1[2 * k] = p1[k].x, s1[2 * k + 1] = p1[k].y;
      s2[2 * k] = p2[k].x, s2[2 * k + 1] = p2[k].y;
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      red[(lane * C8 + c8) * 16 + k] = s1[k];
      red[(lane * C8 + c8) * 16 + 8 + k] = s2[k];
    }
  }
  __syncthreads();
  // fixed-order sum over the row lanes; thread -> (c8, j in 0..15)
  for (int idx = threadIdx.x; idx < C8 * 16; idx += blockDim.x) {
    const int cc = idx >> 4, j = idx & 15;
    float acc = 0.f;
    for (int l = 0; l < lanes; ++l) acc += red[(l * C8 + cc) * 16 + j];
    float* dst = part + (((int64_t)b * nblk + blk) * 2) * C;
    dst[(j >> 3) * C + cc * 8 + (j & 7)] = acc;
  }
  // ---- ticket: is this the last block of sample b? -----------------------------------------------------------
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_last = atomicAdd(tl.tickets + b, 1) == nblk - 1;
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  float* S = red;          // [2][C] folded sums
  float* G = red + 2 * C;  // [2][C] gamma' * sums, for the group totals
  {
    // fold the nblk partial rows of this sample: float4 columns x up to 8 row groups (rows k = grp, grp + ng, ...), then
    // the groups in order - the summation order depends only on (nblk, C), never on which block got here last
    const int n4 = (2 * C) >> 2;
    int ng = (int)blockDim.x / n4;
    ng = ng < 1 ? 1 : (ng > 8 ? 8 : ng);
    float4* F = reinterpret_cast<float4*>(red);  // [ng][n4]
    const float4* src0 = reinterpret_cast<const float4*>(part + ((int64_t)b * nblk * 2) * C);
    for (int item = threadIdx.x; item < ng * n4; item += blockDim.x) {
      const int grp = item / n4, col = item - grp * n4;
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
      int k = grp;
      for (; k + 7 * ng < nblk; k += 8 * ng) {  // eight independent L2 loads in flight, added in row order
        float4 v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) v[u] = __ldcg(src0 + (int64_t)(k + u * ng) * n4 + col);
#pragma unroll
        for (int u = 0; u < 8; ++u) acc.x += v[u].x, acc.y += v[u].y, acc.z += v[u].z, acc.w += v[u].w;
      }
      for (; k < nblk; k += ng) {
        const float4 v = __ldcg(src0 + (int64_t)k * n4 + col);
        acc.x += v.x, acc.y += v.y, acc.z += v.z, acc.w += v.w;
      }
      F[grp * n4 + col] = acc;
    }
    __syncthreads();
    for (int idx = threadIdx.x; idx < 2 * C; idx += blockDim.x) {
      float acc = red[idx];
      for (int gq = 1; gq < ng; ++gq) acc += red[gq * 2 * C + idx];
      S[idx] = acc;  // in place over group 0's row: column idx is touched by this thread only
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const int g = c / cpg;
    const float mean = tl.stats[((int64_t)b * tl.groups + g) * 2 + 0];
    const float rstd = tl.stats[((int64_t)b * tl.groups + g) * 2 + 1];
    const float t1 = S[c];
    const float t2 = rstd * fmaf(-mean, t1, S[C + c]);  // sum dn * xhat
    const float ga = tl.gamma[c];
    const float sc = tl.ss != nullptr ? 1.f + tl.ss[b * tl.ss_stride + c] : 1.f;
    const float gp = ga * sc;
    G[c] = gp * t1;
    G[C + c] = gp * t2;
    tl.dgb_part[((int64_t)b * 2) * C + c] = t2 * sc;      // dgamma contribution
    tl.dgb_part[((int64_t)b * 2 + 1) * C + c] = t1 * sc;  // dbeta contribution
    if (tl.dss != nullptr) {
      tl.dss[(int64_t)b * 2 * C + c] = fmaf(t2, ga, t1 * tl.beta[c]);  // d scale
      tl.dss[(int64_t)b * 2 * C + C + c] = t1;                         // d shift
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const int g0 = (c / cpg) * cpg;
    float g1 = 0.f, g2 = 0.f;
    for (int k = 0; k < cpg; ++k) {
      g1 += G[g0 + k];
      g2 += G[C + g0 + k];
    }
    float a, bb, mean, rstd;
    gn_affine_coef(tl.stats, tl.gamma, tl.beta, tl.ss, tl.ss_stride, b, c, C, cpg, tl.groups, a, bb, mean, rstd);
    const float m1 = rstd * g1 * tl.inv_n, m2 = rstd * g2 * tl.inv_n;
    // dx = A*dn - m1 - m2*xhat = a*dn + P + Q*x  (P = m2*mean*rstd - m1, Q = -m2*rstd)
    *reinterpret_cast<float4*>(tl.tab + ((int64_t)b * C + c) * kGnTab) =
        make_float4(a, bb, fmaf(m2, mean * rstd, -m1), -m2 * rstd);
  }
  if (threadIdx.x == 0) tl.tickets[b] = 0;  // leave the ticket buffer zero
  // (dgamma / dbeta = the per-sample rows folded over the batch: done by one block of the apply pass, which the kernel
  // boundary orders after every sample's finalize - a second ticket here cost ~4 us of fences per launch)
}

// pass 2: dx = a*dn + P + Q*x (+ the gradients other consumers of x left: add_a / add_b for source 0, add_1 for source 1)
template <bool ADD>
__global__ void __launch_bounds__(256, 2) gn_bwd_apply_kernel(const uint4* __restrict__ x0, int C0_8,
                                                            const uint4* __restrict__ x1, const uint4* __restrict__ da,
                                                            const float* __restrict__ tab, uint4* __restrict__ dx0,
                                                            uint4* __restrict__ dx1, int64_t HW, int C8, int lanes,
                                                            int rows_per_blk, int silu, float* __restrict__ colpart,
                                                            const uint4* __restrict__ add_a,
                                                            const uint4* __restrict__ add_b,
                                                            const uint4* __restrict__ add_1,
                                                            const float* __restrict__ dgb_part,
                                                            float* __restrict__ dgamma, float* __restrict__ dbeta,
                                                            int B) {
  pdl_enter();
  // colpart (optional): per-block column sums of dx, [B][nblk][C] -- the bias / time-embedding-add gradient of the conv
  // that produced x, so that conv's backward needs no column-sum pass over dx
  extern __shared__ float cred[];  // [lanes][C8*8], only when colpart != NULL
  // row blocks from the start, images innermost: the rows the partial pass read last are still in the L2
  const int nblk = gridDim.x, lin = blockIdx.y * nblk + blockIdx.x;
  const int b = lin % (int)gridDim.y, blk = lin / (int)gridDim.y;
  const int c8 = threadIdx.x % C8, lane = threadIdx.x / C8;
  const int C = C8 * 8;
  float cs[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  if (lane < lanes) {
    float2 ca[4], cah[4], cbh[4], cP[4], cQ[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float4 t0 = *reinterpret_cast<const float4*>(tab + ((int64_t)b * C + c8 * 8 + 2 * k) * kGnTab);
      const float4 t1 = *reinterpret_cast<const float4*>(tab + ((int64_t)b * C + c8 * 8 + 2 * k + 1) * kGnTab);
      ca[k] = make_float2(t0.x, t1.x);
      cah[k] = make_float2(0.5f * t0.x, 0.5f * t1.x), cbh[k] = make_float2(0.5f * t0.y, 0.5f * t1.y);
      cP[k] = make_float2(t0.z, t1.z), cQ[k] = make_float2(t0.w, t1.w);
    }
    float2 cs2[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) cs2[k] = make_float2(0.f, 0.f);
    const int64_t r0 = (int64_t)blk * rows_per_blk;
    const int64_t r1 = r0 + rows_per_blk < HW ? r0 + rows_per_blk : HW;
    const bool first = c8 < C0_8;
    const int xC8 = first ? C0_8 : C8 - C0_8;
    const int64_t xoff = ((int64_t)b * HW) * xC8 + (first ? c8 : c8 - C0_8);
    const uint4* xs = (first ? x0 : x1) + xoff;
    const uint4* ds = da + ((int64_t)b * HW) * C8 + c8;
    uint4* os = (first ? dx0 : dx1) + xoff;
    const uint4* ea = ADD ? (first ? add_a : add_1) : nullptr;
    const uint4* eb = ADD && first ? add_b : nullptr;
    if (ea != nullptr) ea += xoff;
    if (eb != nullptr) eb += xoff;
    constexpr int kU = ADD ? 2 : 4;  // independent row loads in flight per thread (x, dout and the added tensors)
    constexpr int kA = ADD ? kU : 1;
    for (int64_t r = r0 + lane; r < r1; r += (int64_t)lanes * kU) {
      uint4 xv[kU], dv[kU], av[kA], bv[kA];
#pragma unroll
      for (int u = 0; u < kU; ++u) {
        const int64_t ru = r + (int64_t)u * lanes;
        if (ADD) av[u % kA] = bv[u % kA] = make_uint4(0, 0, 0, 0);
        if (ru < r1) {
          xv[u] = __ldg(xs + ru * xC8);
          dv[u] = __ldg(ds + ru * C8);
          if (ADD && ea != nullptr) av[u % kA] = __ldg(ea + ru * xC8);
          if (ADD && eb != nullptr) bv[u % kA] = __ldg(eb + ru * xC8);
        }
      }
#pragma unroll
      for (int u = 0; u < kU; ++u) {
        const int64_t ru = r + (int64_t)u * lanes;
        if (ru >= r1) break;
        const uint32_t xw[4] = {xv[u].x, xv[u].y, xv[u].z, xv[u].w}, dw[4] = {dv[u].x, dv[u].y, dv[u].z, dv[u].w};
        const uint32_t aw[4] = {av[u % kA].x, av[u % kA].y, av[u % kA].z, av[u % kA].w};
        const uint32_t bw[4] = {bv[u % kA].x, bv[u % kA].y, bv[u % kA].z, bv[u % kA].w};
        uint32_t ow[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float2 xf = bf16x2_as_f32x2(xw[k]);
          float2 dn = bf16x2_as_f32x2(dw[k]);
          if (silu) dn = fmul2(dn, silu_grad2(ffma2(cah[k], xf, cbh[k])));
          float2 o = ffma2(ca[k], dn, ffma2(cQ[k], xf, cP[k]));
          if (ADD) o = fadd2(o, fadd2(bf16x2_as_f32x2(aw[k]), bf16x2_as_f32x2(bw[k])));  // absent tensors load as zeros
          ow[k] = pack_bf16x2(o.x, o.y);
          cs2[k] = fadd2(cs2[k], o);
        }
        os[ru * xC8] = make_uint4(ow[0], ow[1], ow[2], ow[3]);
      }
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) cs[2 * k] = cs2[k].x, cs[2 * k + 1] = cs2[k].y;
  }
  if (colpart != nullptr) {
    if (lane < lanes) {
#pragma unroll
      for (int k = 0; k < 8; ++k) cred[(lane * C8 + c8) * 8 + k] = cs[k];
    }
    __syncthreads();
    for (int idx = threadIdx.x; idx < C; idx += blockDim.x) {
      float acc = 0.f;
      for (int l = 0; l < lanes; ++l) acc += cred[l * C + idx];
      colpart[((int64_t)b * nblk + blk) * C + idx] = acc;
    }
  }
  // one block also folds the per-sample (dgamma, dbeta) rows the partial pass left, in sample order
  if (blk == nblk - 1 && b == (int)gridDim.y - 1) {
    for (int idx = threadIdx.x; idx < 2 * C; idx += blockDim.x) {
      float acc = 0.f;
      int sm = 0;
      for (; sm + 8 <= B; sm += 8) {
        float v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) v[u] = __ldg(dgb_part + (int64_t)(sm + u) * 2 * C + idx);
#pragma unroll
        for (int u = 0; u < 8; ++u) acc += v[u];
      }
      for (; sm < B; ++sm) acc += __ldg(dgb_part + (int64_t)sm * 2 * C + idx);
      if (dbeta != nullptr && idx >= C) dbeta[idx - C] = acc;
      else dgamma[idx] = acc;
    }
  }
}

// =============================================================================================================
// attention backward: one CTA per (sample, head); Q, K, V, dO staged in shared memory as bf16, math in fp32
// =============================================================================================================
// one row of HD staged values as fp32 (16-byte vector loads); the tiles are staged as fp32 when they fit in shared
// memory (no per-use bf16 unpacking in the T^2 loops), else as bf16
template <int HD, typename ST>
__device__ __forceinline__ void att_row(const ST* row, float* out) {
  if constexpr (sizeof(ST) == 4) {
#pragma unroll
    for (int c = 0; c < HD / 4; ++c) {
      const float4 v = reinterpret_cast<const float4*>(row)[c];
      out[c * 4] = v.x, out[c * 4 + 1] = v.y, out[c * 4 + 2] = v.z, out[c * 4 + 3] = v.w;
    }
  } else {
#pragma unroll
    for (int c = 0; c < HD / 8; ++c) {
      const uint4 v = reinterpret_cast<const uint4*>(row)[c];
      const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float2 f = unpack_bf16x2(w[k]);
        out[c * 8 + 2 * k] = f.x;
        out[c * 8 + 2 * k + 1] = f.y;
      }
    }
  }
}
__device__ __forceinline__ void att_stage8(float* dst, const __nv_bfloat16* src) {
  const uint4 v = *reinterpret_cast<const uint4*>(src);
  const float2 a = unpack_bf16x2(v.x), b = unpack_bf16x2(v.y), c = unpack_bf16x2(v.z), d = unpack_bf16x2(v.w);
  reinterpret_cast<float4*>(dst)[0] = make_float4(a.x, a.y, b.x, b.y);
  reinterpret_cast<float4*>(dst)[1] = make_float4(c.x, c.y, d.x, d.y);
}
__device__ __forceinline__ void att_stage8(__nv_bfloat16* dst, const __nv_bfloat16* src) {
  *reinterpret_cast<uint4*>(dst) = *reinterpret_cast<const uint4*>(src);
}

// ---- head_dim 8 on the tensor cores (mma.sync; K = 8 is half a tcgen05 K step, as in the forward kernel) ----------
// One CTA per (sample, head), 8 warps.  Q, K, V, dO are staged once as bf16 rows (operands of the S / dP products) and
// K, Q, dO also transposed (operands of the dQ / dK / dV products).
//   phase A, warp = 16 query rows: online (max, sum) of S = Q K^T over all keys and delta = rowsum(dO * O); then, per 16
//            keys, P = 2^(S c - lse), dP = dO V^T, dS = P (dP - delta) scale, dQ += dS K.  lse / delta are parked in
//            shared memory for phase B.
//   phase B, warp = 16 key rows: per 16 queries, S^T = K Q^T, P^T, dP^T = V dO^T, dS^T; dV += P^T dO, dK += dS^T Q.
// The fragments of S / S^T are re-used in place as the A operands of the second products (same lane layout), so no
// tile passes through shared memory; every sum has a fixed order (no atomics).
constexpr int kAttB8Pad = 8;
__global__ void __launch_bounds__(256) attention_bwd_hd8_mma_kernel(
    const __nv_bfloat16* __restrict__ q, const __nv_bfloat16* __restrict__ k, const __nv_bfloat16* __restrict__ v,
    const __nv_bfloat16* __restrict__ o, const __nv_bfloat16* __restrict__ dout, __nv_bfloat16* __restrict__ dq,
    __nv_bfloat16* __restrict__ dk, __nv_bfloat16* __restrict__ dv, int heads, int Tq, int Tqp, int Tk, int Tkp,
    int64_t qs_b, int64_t qs_h, int64_t qs_t, int64_t ks_b, int64_t ks_h, int64_t ks_t, int64_t os_b, int64_t os_h,
    int64_t os_t, float scale) {
  extern __shared__ __align__(16) uint8_t smem_att[];
  // queries (Tq rows: Q, dO, lse, delta) and keys (Tk rows: K, V) have their own lengths: cross-attention has Tq != Tk
  __nv_bfloat16* sQ = reinterpret_cast<__nv_bfloat16*>(smem_att);  // [Tqp][8]
  __nv_bfloat16* sD = sQ + (size_t)Tqp * 8;                         // dO [Tqp][8]
  __nv_bfloat16* sK = sD + (size_t)Tqp * 8;                         // [Tkp][8]
  __nv_bfloat16* sV = sK + (size_t)Tkp * 8;
  const int tsq = Tqp + kAttB8Pad, tsk = Tkp + kAttB8Pad;           // transposed row strides
  __nv_bfloat16* sKt = sV + (size_t)Tkp * 8;                        // [8][tsk]
  __nv_bfloat16* sQt = sKt + (size_t)8 * tsk;                       // [8][tsq]
  __nv_bfloat16* sDt = sQt + (size_t)8 * tsq;
  float* s_lse = reinterpret_cast<float*>(sDt + (size_t)8 * tsq);   // [Tqp]
  float* s_del = s_lse + Tqp;
  const int b = blockIdx.x / heads, h = blockIdx.x % heads;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int64_t qoff = b * qs_b + h * qs_h, koff = b * ks_b + h * ks_h, ooff = b * os_b + h * os_h;
  const float c = scale * 1.4426950408889634f;

  for (int r = threadIdx.x; r < Tqp; r += blockDim.x) {
    uint4 vq = make_uint4(0, 0, 0, 0), vd = vq;
    if (r < Tq) {
      vq = *reinterpret_cast<const uint4*>(q + qoff + (int64_t)r * qs_t);
      vd = *reinterpret_cast<const uint4*>(dout + ooff + (int64_t)r * os_t);
    }
    *reinterpret_cast<uint4*>(sQ + r * 8) = vq;
    *reinterpret_cast<uint4*>(sD + r * 8) = vd;
    const __nv_bfloat16* eq = reinterpret_cast<const __nv_bfloat16*>(&vq);
    const __nv_bfloat16* ed = reinterpret_cast<const __nv_bfloat16*>(&vd);
#pragma unroll
    for (int d = 0; d < 8; ++d) {
      sQt[d * tsq + r] = eq[d];
      sDt[d * tsq + r] = ed[d];
    }
  }
  for (int r = threadIdx.x; r < Tkp; r += blockDim.x) {
    uint4 vk = make_uint4(0, 0, 0, 0), vv = vk;
    if (r < Tk) {
      vk = *reinterpret_cast<const uint4*>(k + koff + (int64_t)r * ks_t);
      vv = *reinterpret_cast<const uint4*>(v + koff + (int64_t)r * ks_t);
    }
    *reinterpret_cast<uint4*>(sK + r * 8) = vk;
    *reinterpret_cast<uint4*>(sV + r * 8) = vv;
    const __nv_bfloat16* ek = reinterpret_cast<const __nv_bfloat16*>(&vk);
#pragma unroll
    for (int d = 0; d < 8; ++d) sKt[d * tsk + r] = ek[d];
  }
  __syncthreads();

  // ================= phase A: dQ, lse, delta =================
  for (int q0 = warp * 16; q0 < Tqp; q0 += 8 * 16) {
    const uint32_t qa0 = *reinterpret_cast<const uint32_t*>(sQ + (q0 + g) * 8 + 2 * t);
    const uint32_t qa1 = *reinterpret_cast<const uint32_t*>(sQ + (q0 + g + 8) * 8 + 2 * t);
    const uint32_t da0 = *reinterpret_cast<const uint32_t*>(sD + (q0 + g) * 8 + 2 * t);
    const uint32_t da1 = *reinterpret_cast<const uint32_t*>(sD + (q0 + g + 8) * 8 + 2 * t);
    // delta = sum_d dO * O (O from global memory: it is read exactly once)
    float del0 = 0.f, del1 = 0.f;
    {
      const int r0 = q0 + g, r1 = q0 + g + 8;
      const float2 d0 = bf16x2_as_f32x2(da0), d1 = bf16x2_as_f32x2(da1);
      if (r0 < Tq) {
        const float2 o0 = bf16x2_as_f32x2(*reinterpret_cast<const uint32_t*>(o + ooff + (int64_t)r0 * os_t + 2 * t));
        del0 = fmaf(d0.x, o0.x, d0.y * o0.y);
      }
      if (r1 < Tq) {
        const float2 o1 = bf16x2_as_f32x2(*reinterpret_cast<const uint32_t*>(o + ooff + (int64_t)r1 * os_t + 2 * t));
        del1 = fmaf(d1.x, o1.x, d1.y * o1.y);
      }
      del0 += __shfl_xor_sync(0xffffffffu, del0, 1);
      del0 += __shfl_xor_sync(0xffffffffu, del0, 2);
      del1 += __shfl_xor_sync(0xffffffffu, del1, 1);
      del1 += __shfl_xor_sync(0xffffffffu, del1, 2);
    }
    // pass 1: online row max / sum (log2 domain), 64 keys at a time
    float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;
    for (int kb0 = 0; kb0 < Tkp; kb0 += 64) {
      float sc[8][4];
      const int nblk = (Tkp - kb0) >= 64 ? 8 : (Tkp - kb0) / 8;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        sc[j][0] = sc[j][1] = sc[j][2] = sc[j][3] = -INFINITY;
        if (j < nblk) {
          sc[j][0] = sc[j][1] = sc[j][2] = sc[j][3] = 0.f;
          mma_m16n8k8_bf16(sc[j], qa0, qa1, *reinterpret_cast<const uint32_t*>(sK + (kb0 + j * 8 + g) * 8 + 2 * t));
          const int key = kb0 + j * 8 + 2 * t;
          if (key >= Tk) sc[j][0] = sc[j][2] = -INFINITY;
          if (key + 1 >= Tk) sc[j][1] = sc[j][3] = -INFINITY;
        }
      }
      float mx0 = sc[0][0], mx1 = sc[0][2];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        mx0 = fmaxf(mx0, fmaxf(sc[j][0], sc[j][1]));
        mx1 = fmaxf(mx1, fmaxf(sc[j][2], sc[j][3]));
      }
      mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1));
      mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
      mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
      mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
      const float mn0 = fmaxf(m0, mx0 * c), mn1 = fmaxf(m1, mx1 * c);
      l0 *= ex2_approx(m0 - mn0);
      l1 *= ex2_approx(m1 - mn1);
      m0 = mn0, m1 = mn1;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        l0 += ex2_approx(fmaf(sc[j][0], c, -mn0)) + ex2_approx(fmaf(sc[j][1], c, -mn0));
        l1 += ex2_approx(fmaf(sc[j][2], c, -mn1)) + ex2_approx(fmaf(sc[j][3], c, -mn1));
      }
    }
    l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
    l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
    const float lse0 = m0 + __log2f(l0), lse1 = m1 + __log2f(l1);
    if (t == 0) {  // rows past T: +inf, so phase B's 2^(s c - lse) is exactly 0 for them
      s_lse[q0 + g] = (q0 + g < Tq) ? lse0 : INFINITY;
      s_lse[q0 + g + 8] = (q0 + g + 8 < Tq) ? lse1 : INFINITY;
      s_del[q0 + g] = del0;
      s_del[q0 + g + 8] = del1;
    }
    // pass 2: dQ
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    for (int kb = 0; kb < Tkp; kb += 16) {
      uint32_t af[4];
#pragma unroll
      for (int hb = 0; hb < 2; ++hb) {
        const int kk = kb + hb * 8;
        float sv[4] = {0.f, 0.f, 0.f, 0.f}, dp[4] = {0.f, 0.f, 0.f, 0.f};
        mma_m16n8k8_bf16(sv, qa0, qa1, *reinterpret_cast<const uint32_t*>(sK + (kk + g) * 8 + 2 * t));
        mma_m16n8k8_bf16(dp, da0, da1, *reinterpret_cast<const uint32_t*>(sV + (kk + g) * 8 + 2 * t));
        const int key = kk + 2 * t;
        const float p0 = key < Tk ? ex2_approx(fmaf(sv[0], c, -lse0)) : 0.f;
        const float p1 = key + 1 < Tk ? ex2_approx(fmaf(sv[1], c, -lse0)) : 0.f;
        const float p2 = key < Tk ? ex2_approx(fmaf(sv[2], c, -lse1)) : 0.f;
        const float p3 = key + 1 < Tk ? ex2_approx(fmaf(sv[3], c, -lse1)) : 0.f;
        af[hb * 2 + 0] = pack_bf16x2(p0 * (dp[0] - del0) * scale, p1 * (dp[1] - del0) * scale);
        af[hb * 2 + 1] = pack_bf16x2(p2 * (dp[2] - del1) * scale, p3 * (dp[3] - del1) * scale);
      }
      const uint32_t b0 = *reinterpret_cast<const uint32_t*>(sKt + g * tsk + kb + 2 * t);
      const uint32_t b1 = *reinterpret_cast<const uint32_t*>(sKt + g * tsk + kb + 8 + 2 * t);
      mma_m16n8k16_bf16(acc, af[0], af[1], af[2], af[3], b0, b1);
    }
    if (q0 + g < Tq)
      *reinterpret_cast<uint32_t*>(dq + qoff + (int64_t)(q0 + g) * qs_t + 2 * t) = pack_bf16x2(acc[0], acc[1]);
    if (q0 + g + 8 < Tq)
      *reinterpret_cast<uint32_t*>(dq + qoff + (int64_t)(q0 + g + 8) * qs_t + 2 * t) = pack_bf16x2(acc[2], acc[3]);
  }
  __syncthreads();

  // ================= phase B: dK, dV =================
  for (int k0 = warp * 16; k0 < Tkp; k0 += 8 * 16) {
    const uint32_t ka0 = *reinterpret_cast<const uint32_t*>(sK + (k0 + g) * 8 + 2 * t);
    const uint32_t ka1 = *reinterpret_cast<const uint32_t*>(sK + (k0 + g + 8) * 8 + 2 * t);
    const uint32_t va0 = *reinterpret_cast<const uint32_t*>(sV + (k0 + g) * 8 + 2 * t);
    const uint32_t va1 = *reinterpret_cast<const uint32_t*>(sV + (k0 + g + 8) * 8 + 2 * t);
    float ak[4] = {0.f, 0.f, 0.f, 0.f}, av[4] = {0.f, 0.f, 0.f, 0.f};
    for (int qb = 0; qb < Tqp; qb += 16) {
      uint32_t pf[4], sf[4];
#pragma unroll
      for (int hb = 0; hb < 2; ++hb) {
        const int qq = qb + hb * 8;
        float st[4] = {0.f, 0.f, 0.f, 0.f}, dp[4] = {0.f, 0.f, 0.f, 0.f};
        mma_m16n8k8_bf16(st, ka0, ka1, *reinterpret_cast<const uint32_t*>(sQ + (qq + g) * 8 + 2 * t));
        mma_m16n8k8_bf16(dp, va0, va1, *reinterpret_cast<const uint32_t*>(sD + (qq + g) * 8 + 2 * t));
        const float2 ls = *reinterpret_cast<const float2*>(s_lse + qq + 2 * t);   // columns = queries qq+2t, qq+2t+1
        const float2 dl = *reinterpret_cast<const float2*>(s_del + qq + 2 * t);
        const float p0 = ex2_approx(fmaf(st[0], c, -ls.x)), p1 = ex2_approx(fmaf(st[1], c, -ls.y));
        const float p2 = ex2_approx(fmaf(st[2], c, -ls.x)), p3 = ex2_approx(fmaf(st[3], c, -ls.y));
        pf[hb * 2 + 0] = pack_bf16x2(p0, p1);
        pf[hb * 2 + 1] = pack_bf16x2(p2, p3);
        sf[hb * 2 + 0] = pack_bf16x2(p0 * (dp[0] - dl.x) * scale, p1 * (dp[1] - dl.y) * scale);
        sf[hb * 2 + 1] = pack_bf16x2(p2 * (dp[2] - dl.x) * scale, p3 * (dp[3] - dl.y) * scale);
      }
      const uint32_t d0 = *reinterpret_cast<const uint32_t*>(sDt + g * tsq + qb + 2 * t);
      const uint32_t d1 = *reinterpret_cast<const uint32_t*>(sDt + g * tsq + qb + 8 + 2 * t);
      mma_m16n8k16_bf16(av, pf[0], pf[1], pf[2], pf[3], d0, d1);
      const uint32_t q0b = *reinterpret_cast<const uint32_t*>(sQt + g * tsq + qb + 2 * t);
      const uint32_t q1b = *reinterpret_cast<const uint32_t*>(sQt + g * tsq + qb + 8 + 2 * t);
      mma_m16n8k16_bf16(ak, sf[0], sf[1], sf[2], sf[3], q0b, q1b);
    }
    if (k0 + g < Tk) {
      *reinterpret_cast<uint32_t*>(dk + koff + (int64_t)(k0 + g) * ks_t + 2 * t) = pack_bf16x2(ak[0], ak[1]);
      *reinterpret_cast<uint32_t*>(dv + koff + (int64_t)(k0 + g) * ks_t + 2 * t) = pack_bf16x2(av[0], av[1]);
    }
    if (k0 + g + 8 < Tk) {
      *reinterpret_cast<uint32_t*>(dk + koff + (int64_t)(k0 + g + 8) * ks_t + 2 * t) = pack_bf16x2(ak[2], ak[3]);
      *reinterpret_cast<uint32_t*>(dv + koff + (int64_t)(k0 + g + 8) * ks_t + 2 * t) = pack_bf16x2(av[2], av[3]);
    }
  }
}

template <int HD, typename ST>
__global__ void __launch_bounds__(256) attention_bwd_kernel(
    const __nv_bfloat16* __restrict__ q, const __nv_bfloat16* __restrict__ k, const __nv_bfloat16* __restrict__ v,
    const __nv_bfloat16* __restrict__ o, const __nv_bfloat16* __restrict__ dout, __nv_bfloat16* __restrict__ dq,
    __nv_bfloat16* __restrict__ dk, __nv_bfloat16* __restrict__ dv, int heads, int Tq, int Tk, int64_t qs_b,
    int64_t qs_h, int64_t qs_t, int64_t ks_b, int64_t ks_h, int64_t ks_t, int64_t os_b, int64_t os_h, int64_t os_t,
    float scale) {
  // Tq query rows (q, dq, o, dout) and Tk key rows (k, v, dk, dv): equal for self-attention
  extern __shared__ __align__(16) uint8_t smem_att[];
  ST* sq = reinterpret_cast<ST*>(smem_att);
  ST* sdo = sq + (size_t)Tq * HD;
  ST* sk = sdo + (size_t)Tq * HD;
  ST* sv = sk + (size_t)Tk * HD;
  float* lse = reinterpret_cast<float*>(sv + (size_t)Tk * HD);
  float* dsum = lse + Tq;
  const int b = blockIdx.x / heads, h = blockIdx.x % heads;
  const int64_t qoff = b * qs_b + h * qs_h, koff = b * ks_b + h * ks_h, ooff = b * os_b + h * os_h;
  // rows of HD contiguous bf16 (16-byte aligned by the caller's strides)
  for (int idx = threadIdx.x; idx < Tq * (HD / 8); idx += blockDim.x) {
    const int t = idx / (HD / 8), c = idx - t * (HD / 8);
    att_stage8(sq + idx * 8, q + qoff + t * qs_t + c * 8);
    att_stage8(sdo + idx * 8, dout + ooff + t * os_t + c * 8);
  }
  for (int idx = threadIdx.x; idx < Tk * (HD / 8); idx += blockDim.x) {
    const int t = idx / (HD / 8), c = idx - t * (HD / 8);
    att_stage8(sk + idx * 8, k + koff + t * ks_t + c * 8);
    att_stage8(sv + idx * 8, v + koff + t * ks_t + c * 8);
  }
  __syncthreads();
  // phase 1+2: per query row i: log-sum-exp, D_i = dO_i . O_i, then dQ_i
  for (int i = threadIdx.x; i < Tq; i += blockDim.x) {
    float qi[HD], doi[HD], tmp[HD];
    att_row<HD>(sq + i * HD, qi);
    att_row<HD>(sdo + i * HD, doi);
    float D = 0.f;
#pragma unroll
    for (int c = 0; c < HD / 8; ++c) {
      const uint4 ov = *reinterpret_cast<const uint4*>(o + ooff + i * os_t + c * 8);
      const uint32_t w[4] = {ov.x, ov.y, ov.z, ov.w};
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        const float2 f = unpack_bf16x2(w[kk]);
        D = fmaf(doi[c * 8 + 2 * kk], f.x, D);
        D = fmaf(doi[c * 8 + 2 * kk + 1], f.y, D);
      }
    }
#pragma unroll
    for (int d = 0; d < HD; ++d) qi[d] *= scale;
    float m = -INFINITY, l = 0.f;
    for (int j = 0; j < Tk; ++j) {
      att_row<HD>(sk + j * HD, tmp);
      float s = 0.f;
#pragma unroll
      for (int d = 0; d < HD; ++d) s = fmaf(qi[d], tmp[d], s);
      const float mn = fmaxf(m, s);
      l = l * __expf(m - mn) + __expf(s - mn);
      m = mn;
    }
    const float L = m + __logf(l);
    lse[i] = L;
    dsum[i] = D;
    float acc[HD];
#pragma unroll
    for (int d = 0; d < HD; ++d) acc[d] = 0.f;
    for (int j = 0; j < Tk; ++j) {
      float kj[HD];
      att_row<HD>(sk + j * HD, kj);
      att_row<HD>(sv + j * HD, tmp);
      float s = 0.f, dp = 0.f;
#pragma unroll
      for (int d = 0; d < HD; ++d) {
        s = fmaf(qi[d], kj[d], s);
        dp = fmaf(doi[d], tmp[d], dp);
      }
      const float ds = __expf(s - L) * (dp - D) * scale;
#pragma unroll
      for (int d = 0; d < HD; ++d) acc[d] = fmaf(ds, kj[d], acc[d]);
    }
#pragma unroll
    for (int c = 0; c < HD / 8; ++c)
      *reinterpret_cast<uint4*>(dq + qoff + i * qs_t + c * 8) =
          make_uint4(pack_bf16x2(acc[c * 8], acc[c * 8 + 1]), pack_bf16x2(acc[c * 8 + 2], acc[c * 8 + 3]),
                     pack_bf16x2(acc[c * 8 + 4], acc[c * 8 + 5]), pack_bf16x2(acc[c * 8 + 6], acc[c * 8 + 7]));
  }
  __syncthreads();
  // phase 3: per key row j: dK_j, dV_j
  for (int j = threadIdx.x; j < Tk; j += blockDim.x) {
    float kj[HD], vj[HD], ak[HD], av[HD];
    att_row<HD>(sk + j * HD, kj);
    att_row<HD>(sv + j * HD, vj);
#pragma unroll
    for (int d = 0; d < HD; ++d) {
      kj[d] *= scale;
      ak[d] = 0.f;
      av[d] = 0.f;
    }
    for (int i = 0; i < Tq; ++i) {
      float qi[HD], doi[HD];
      att_row<HD>(sq + i * HD, qi);
      att_row<HD>(sdo + i * HD, doi);
      float s = 0.f, dp = 0.f;
#pragma unroll
      for (int d = 0; d < HD; ++d) {
        s = fmaf(kj[d], qi[d], s);
        dp = fmaf(vj[d], doi[d], dp);
      }
      const float pr = __expf(s - lse[i]);
      const float ds = pr * (dp - dsum[i]) * scale;
#pragma unroll
      for (int d = 0; d < HD; ++d) {
        av[d] = fmaf(pr, doi[d], av[d]);
        ak[d] = fmaf(ds, qi[d], ak[d]);
      }
    }
#pragma unroll
    for (int c = 0; c < HD / 8; ++c) {
      *reinterpret_cast<uint4*>(dk + koff + j * ks_t + c * 8) =
          make_uint4(pack_bf16x2(ak[c * 8], ak[c * 8 + 1]), pack_bf16x2(ak[c * 8 + 2], ak[c * 8 + 3]),
                     pack_bf16x2(ak[c * 8 + 4], ak[c * 8 + 5]), pack_bf16x2(ak[c * 8 + 6], ak[c * 8 + 7]));
      *reinterpret_cast<uint4*>(dv + koff + j * ks_t + c * 8) =
          make_uint4(pack_bf16x2(av[c * 8], av[c * 8 + 1]), pack_bf16x2(av[c * 8 + 2], av[c * 8 + 3]),
                     pack_bf16x2(av[c * 8 + 4], av[c * 8 + 5]), pack_bf16x2(av[c * 8 + 6], av[c * 8 + 7]));
    }
  }
}

// =============================================================================================================
// small fp32 elementwise: dx = dy * silu'(x)
// =============================================================================================================
__global__ void __launch_bounds__(256) silu_bwd_kernel(const float* __restrict__ x, const float* __restrict__ dy,
                                                        float* __restrict__ dx, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float xv = x[i];
    const float s = 1.f / (1.f + expf(-xv));
    dx[i] = dy[i] * (s * (1.f + xv * (1.f - s)));
  }
}

// =============================================================================================================
// stem conv wgrad: x fp32 NCHW (two concatenated sources), dY bf16 NHWC -> partial dW [blk][Cout][CIN*9]
// =============================================================================================================
template <int CIN>
__global__ void __launch_bounds__(256) stem_wgrad_kernel(const float* __restrict__ x0, int C0,
                                                          const float* __restrict__ x1, float in_scale,
                                                          float in_shift, const __nv_bfloat16* __restrict__ dy,
                                                          float* __restrict__ part, int B, int H, int W, int Cout,
                                                          int segs_per_blk) {
  constexpr int SEG = 64;  // pixels of one image row per step
  __shared__ float patch[CIN][3][SEG + 2];
  const int segs_per_row = (W + SEG - 1) / SEG;
  const int64_t total = (int64_t)B * H * segs_per_row;
  const int64_t s0 = (int64_t)blockIdx.x * segs_per_blk;
  const int64_t s1 = s0 + segs_per_blk < total ? s0 + segs_per_blk : total;
  const int co = threadIdx.x;
  float acc[CIN * 9];
#pragma unroll
  for (int i = 0; i < CIN * 9; ++i) acc[i] = 0.f;
  for (int64_t s = s0; s < s1; ++s) {
    const int xs = (int)(s % segs_per_row) * SEG;
    const int64_t r = s / segs_per_row;
    const int y = (int)(r % H), b = (int)(r / H);
    __syncthreads();
    for (int idx = threadIdx.x; idx < CIN * 3 * (SEG + 2); idx += blockDim.x) {
      const int xx = idx % (SEG + 2);
      const int kh = (idx / (SEG + 2)) % 3;
      const int ci = idx / (3 * (SEG + 2));
      const int yi = y + kh - 1, xi = xs + xx - 1;
      float v = 0.f;
      if (yi >= 0 && yi < H && xi >= 0 && xi < W) {
        const float* src = ci < C0 ? x0 + ((int64_t)b * C0 + ci) * H * W : x1 + ((int64_t)b * (CIN - C0) + ci - C0) * H * W;
        v = fmaf(src[(int64_t)yi * W + xi], in_scale, in_shift);
      }
      patch[ci][kh][xx] = v;
    }
    __syncthreads();
    if (co < Cout) {
      const int npx = min(SEG, W - xs);
      const __nv_bfloat16* drow = dy + (((int64_t)b * H + y) * W + xs) * Cout + co;
      for (int px = 0; px < npx; ++px) {
        const float d = __bfloat162float(drow[(int64_t)px * Cout]);
#pragma unroll
        for (int ci = 0; ci < CIN; ++ci)
#pragma unroll
          for (int kh = 0; kh < 3; ++kh)
#pragma unroll
            for (int kw = 0; kw < 3; ++kw) acc[ci * 9 + kh * 3 + kw] = fmaf(d, patch[ci][kh][px + kw], acc[ci * 9 + kh * 3 + kw]);
      }
    }
  }
  if (co < Cout) {
    float* dst = part + ((int64_t)blockIdx.x * Cout + co) * (CIN * 9);
#pragma unroll
    for (int i = 0; i < CIN * 9; ++i) dst[i] = acc[i];
  }
}

// =============================================================================================================
// head conv (Cout = 1) backward: a bf16 NHWC, dy fp32 [B][1][H][W]
// =============================================================================================================
// dA[p][ci] = sum_taps dy[y-kh+1][x-kw+1] * w[ci][kh][kw]
__global__ void __launch_bounds__(256) head_dgrad_kernel(const float* __restrict__ dy, const float* __restrict__ w,
                                                          uint4* __restrict__ da, int B, int H, int W, int C) {
  extern __shared__ float shw[];  // [9][C]
  for (int i = threadIdx.x; i < 9 * C; i += blockDim.x) shw[i] = w[(i % C) * 9 + i / C];
  __syncthreads();
  const int C8 = C / 8;
  const int64_t total = (int64_t)B * H * W * C8;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int c8 = (int)(idx % C8);
    int64_t r = idx / C8;
    const int x = (int)(r % W);
    r /= W;
    const int y = (int)(r % H);
    const int b = (int)(r / H);
    float o[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll
    for (int kh = 0; kh < 3; ++kh)
#pragma unroll
      for (int kw = 0; kw < 3; ++kw) {
        const int yo = y - kh + 1, xo = x - kw + 1;
        if (yo < 0 || yo >= H || xo < 0 || xo >= W) continue;
        const float d = __ldg(dy + ((int64_t)b * H + yo) * W + xo);
        const float* wr = shw + (kh * 3 + kw) * C + c8 * 8;
#pragma unroll
        for (int k = 0; k < 8; ++k) o[k] = fmaf(d, wr[k], o[k]);
      }
    da[idx] = make_uint4(pack_bf16x2(o[0], o[1]), pack_bf16x2(o[2], o[3]), pack_bf16x2(o[4], o[5]),
                         pack_bf16x2(o[6], o[7]));
  }
}

// dW[ci][kh][kw] = sum_p a[p][ci] * dy[y-kh+1][x-kw+1]; block = (C/2 channel pairs) x (256/(C/2) pixel lanes)
__global__ void __launch_bounds__(256) head_wgrad_kernel(const __nv_bfloat16* __restrict__ a,
                                                          const float* __restrict__ dy, float* __restrict__ part,
                                                          int B, int H, int W, int C, int px_per_blk) {
  __shared__ float red[256 * 2];
  const int cp_n = C / 2;
  const int lanes = blockDim.x / cp_n;
  const int cp = threadIdx.x % cp_n, pl = threadIdx.x / cp_n;
  const int64_t total = (int64_t)B * H * W;
  const int64_t p0 = (int64_t)blockIdx.x * px_per_blk;
  const int64_t p1 = p0 + px_per_blk < total ? p0 + px_per_blk : total;
  float acc[9][2];
#pragma unroll
  for (int t = 0; t < 9; ++t) acc[t][0] = acc[t][1] = 0.f;
  if (pl < lanes) {
    for (int64_t p = p0 + pl; p < p1; p += lanes) {
      const int x = (int)(p % W);
      const int64_t r = p / W;
      const int y = (int)(r % H);
      const int b = (int)(r / H);
      const float2 av = unpack_bf16x2(reinterpret_cast<const uint32_t*>(a + p * C)[cp]);
#pragma unroll
      for (int kh = 0; kh < 3; ++kh)
#pragma unroll
        for (int kw = 0; kw < 3; ++kw) {
          const int yo = y - kh + 1, xo = x - kw + 1;
          if (yo < 0 || yo >= H || xo < 0 || xo >= W) continue;
          const float d = __ldg(dy + ((int64_t)b * H + yo) * W + xo);
          acc[kh * 3 + kw][0] = fmaf(av.x, d, acc[kh * 3 + kw][0]);
          acc[kh * 3 + kw][1] = fmaf(av.y, d, acc[kh * 3 + kw][1]);
        }
    }
  }
  // fixed-order reduction over the pixel lanes, one tap at a time
  for (int t = 0; t < 9; ++t) {
    __syncthreads();
    red[threadIdx.x * 2 + 0] = acc[t][0];
    red[threadIdx.x * 2 + 1] = acc[t][1];
    __syncthreads();
    if (pl == 0) {
      float s0 = 0.f, s1 = 0.f;
      for (int l = 0; l < lanes; ++l) {
        s0 += red[(l * cp_n + cp) * 2 + 0];
        s1 += red[(l * cp_n + cp) * 2 + 1];
      }
      float* dst = part + (int64_t)blockIdx.x * C * 9;
      dst[(cp * 2 + 0) * 9 + t] = s0;
      dst[(cp * 2 + 1) * 9 + t] = s1;
    }
  }
}

// =============================================================================================================
// fp32 sums / MSE
// =============================================================================================================
// part[blk] = sum over the block's range of f(i); mode 0: x[i]; mode 1: (x[i] - (t1[i] - t2[i]))^2 (t2 may be NULL)
__global__ void __launch_bounds__(256) sum_partial_kernel(const float* __restrict__ x, const float* __restrict__ t1,
                                                           const float* __restrict__ t2, double* __restrict__ part,
                                                           int64_t n, int mode) {
  __shared__ double red[256];
  double acc = 0.0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float v = x[i];
    if (mode == 1) {
      const float t = t2 != nullptr ? t1[i] - t2[i] : t1[i];
      v = (v - t) * (v - t);
    }
    acc += (double)v;
  }
  red[threadIdx.x] = acc;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) part[blockIdx.x] = red[0];
}
__global__ void sum_final_kernel(const double* __restrict__ part, int parts, double scale, float* __restrict__ out) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    double acc = 0.0;
    for (int i = 0; i < parts; ++i) acc += part[i];
    out[0] = (float)(acc * scale);
  }
}
// dpred = (2/n) * g * (pred - target)
__global__ void __launch_bounds__(256) mse_bwd_kernel(const float* __restrict__ pred, const float* __restrict__ t1,
                                                       const float* __restrict__ t2, const float* __restrict__ g,
                                                       float* __restrict__ dpred, float two_over_n, int64_t n) {
  const float k = two_over_n * g[0];
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float t = t2 != nullptr ? t1[i] - t2[i] : t1[i];
    dpred[i] = k * (pred[i] - t);
  }
}

// =============================================================================================================
// flat AdamW (torch.optim.AdamW arithmetic, decoupled weight decay, bias-corrected): one launch for every parameter
// =============================================================================================================
__global__ void __launch_bounds__(256) adamw_kernel(float* __restrict__ p, const float* __restrict__ g,
                                                     float* __restrict__ m, float* __restrict__ v, int64_t n,
                                                     float decay, float one_m_beta1, float beta2, float one_m_beta2,
                                                     float step_size, float bc2_sqrt, float eps, float grad_scale) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float gi = __fmul_rn(g[i], grad_scale);
    float pi = __fmul_rn(p[i], decay);                                                // p *= 1 - lr*wd
    const float mi = fmaf(one_m_beta1, __fsub_rn(gi, m[i]), m[i]);                    // m.lerp_(g, 1-beta1)
    const float vi = fmaf(one_m_beta2, __fmul_rn(gi, gi), __fmul_rn(v[i], beta2));    // v.mul_(b2).addcmul_(g, g, 1-b2)
    const float denom = __fadd_rn(__fdiv_rn(__fsqrt_rn(vi), bc2_sqrt), eps);
    pi = fmaf(-step_size, __fdiv_rn(mi, denom), pi);                                  // p.addcdiv_(m, denom, -step)
    p[i] = pi;
    m[i] = mi;
    v[i] = vi;
  }
}

// =============================================================================================================
// backward of the small-batch fp32 Linear (time MLP, the batched per-block embedding projections): B <= 32 rows,
// up to ~20k output features.  y = f(x) W^T + b, f = SiLU or identity.
// =============================================================================================================
__device__ __forceinline__ float silu_exact(float x) { return x / (1.f + expf(-x)); }

// dW[o][i] = sum_b dy[b][o] * f(x[b][i]);  db[o] = sum_b dy[b][o].   block tile: 32 o x 128 i
__global__ void __launch_bounds__(256) linear_wgrad_kernel(const float* __restrict__ x, const float* __restrict__ dy,
                                                            float* __restrict__ dw, float* __restrict__ db, int B,
                                                            int I, int O, int silu_in) {
  __shared__ float sdy[32][32 + 1];
  __shared__ float sfx[32][128 + 1];
  const int i0 = blockIdx.x * 128, o0 = blockIdx.y * 32;
  for (int idx = threadIdx.x; idx < B * 32; idx += 256) {
    const int b = idx / 32, oo = idx % 32;
    sdy[b][oo] = o0 + oo < O ? dy[(int64_t)b * O + o0 + oo] : 0.f;
  }
  for (int idx = threadIdx.x; idx < B * 128; idx += 256) {
    const int b = idx / 128, ii = idx % 128;
    float v = i0 + ii < I ? x[(int64_t)b * I + i0 + ii] : 0.f;
    if (silu_in) v = silu_exact(v);
    sfx[b][ii] = v;
  }
  __syncthreads();
  const int ti = threadIdx.x % 32, to = threadIdx.x / 32;
  float acc[4][4];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int c = 0; c < 4; ++c) acc[a][c] = 0.f;
  for (int b = 0; b < B; ++b) {
    float d[4], f[4];
#pragma unroll
    for (int a = 0; a < 4; ++a) d[a] = sdy[b][to + 8 * a];
#pragma unroll
    for (int c = 0; c < 4; ++c) f[c] = sfx[b][ti + 32 * c];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int c = 0; c < 4; ++c) acc[a][c] = fmaf(d[a], f[c], acc[a][c]);
  }
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    const int o = o0 + to + 8 * a;
    if (o >= O) continue;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const int i = i0 + ti + 32 * c;
      if (i < I) dw[(int64_t)o * I + i] = acc[a][c];
    }
  }
  if (db != nullptr && blockIdx.x == 0 && threadIdx.x < 32 && o0 + threadIdx.x < O) {
    float sum = 0.f;
    for (int b = 0; b < B; ++b) sum += sdy[b][threadIdx.x];
    db[o0 + threadIdx.x] = sum;
  }
}

// part[chunk][b][i] = sum_{o in chunk} dy[b][o] * W[o][i];  chunk = 256 output features, block tile: 128 i x all b
constexpr int kLinChunk = 256;
__global__ void __launch_bounds__(256) linear_dgrad_partial_kernel(const float* __restrict__ dy,
                                                                    const float* __restrict__ W,
                                                                    float* __restrict__ part, int B, int I, int O) {
  __shared__ float sdy[32][kLinChunk];
  const int i = blockIdx.x * 128 + threadIdx.x % 128, half = threadIdx.x / 128;
  const int oc0 = blockIdx.y * kLinChunk;
  const int no = min(kLinChunk, O - oc0);
  for (int idx = threadIdx.x; idx < B * kLinChunk; idx += 256) {
    const int b = idx / kLinChunk, oo = idx % kLinChunk;
    sdy[b][oo] = oo < no ? dy[(int64_t)b * O + oc0 + oo] : 0.f;
  }
  __syncthreads();
  float acc[16];  // rows b = half, half + 2, ...
#pragma unroll
  for (int k = 0; k < 16; ++k) acc[k] = 0.f;
  if (i < I) {
    for (int oo = 0; oo < no; ++oo) {
      const float w = __ldg(W + (int64_t)(oc0 + oo) * I + i);
#pragma unroll
      for (int k = 0; k < 16; ++k) {
        const int b = half + 2 * k;
        if (b < B) acc[k] = fmaf(sdy[b][oo], w, acc[k]);
      }
    }
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      const int b = half + 2 * k;
      if (b < B) part[((int64_t)blockIdx.y * B + b) * I + i] = acc[k];
    }
  }
}
// dx[b][i] = f'(x[b][i]) * sum_chunk part[chunk][b][i]
__global__ void __launch_bounds__(256) linear_dgrad_finish_kernel(const float* __restrict__ part,
                                                                   const float* __restrict__ x, float* __restrict__ dx,
                                                                   int nchunk, int64_t n, int silu_in) {
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += (int64_t)gridDim.x * blockDim.x) {
    float acc = 0.f;
    for (int c = 0; c < nchunk; ++c) acc += part[(int64_t)c * n + idx];
    if (silu_in) {
      const float xv = x[idx];
      const float sg = 1.f / (1.f + expf(-xv));
      acc *= sg * (1.f + xv * (1.f - sg));
    }
    dx[idx] = acc;
  }
}

static int ew_grid(int64_t n) {
  int64_t b = (n + 255) / 256;
  const int64_t cap = (int64_t)sm_count() * 16;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (int)b;
}

}  // namespace fm

using namespace fm;

// ---------------------------------------------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------------------------------------------
extern "C" int64_t fm_conv_wgrad_workspace_elems(int32_t B, int32_t Ho, int32_t Wo, int32_t Cin, int32_t Cout,
                                                 int32_t ksize) {
  int splits, cps, chunks, base;
  if (ksize != 1 && ksize != 3) return 0;
  if (ensure_device()) return 0;
  if (wgrad_plan(B, Ho, Wo, Cin, Cout, ksize, &splits, &cps, &chunks, &base)) return 0;
  const int tc = wgrad_tc_max_splits(B, Ho, Wo, Cin, Cout, ksize);
  if (tc > splits) splits = tc;
  return (int64_t)splits * ksize * ksize * Cout * Cin;
}

extern "C" int fm_conv_wgrad_bf16(const void* dy, const void* x, float* dw, float* workspace, int32_t B, int32_t H,
                                  int32_t W, int32_t Cin, int32_t Cout, int32_t ksize, int32_t stride,
                                  int32_t cin_total, int32_t c_begin, fm_stream_t stream) {
  if (int e = ensure_device()) return e;
  FM_REQUIRE(dy && x && dw && workspace, "conv_wgrad: null pointer");
  FM_REQUIRE(ksize == 1 || ksize == 3, "conv_wgrad: ksize must be 1 or 3");
  FM_REQUIRE(stride == 1 || stride == 2, "conv_wgrad: stride must be 1 or 2");
  FM_REQUIRE(Cin % 8 == 0 && Cout % 8 == 0, "conv_wgrad: channel counts must be multiples of 8");
  FM_REQUIRE(c_begin >= 0 && c_begin + Cin <= cin_total, "conv_wgrad: channel slice out of range");
  const int Ho = (H + stride - 1) / stride, Wo = (W + stride - 1) / stride;
  cudaStream_t st = (cudaStream_t)stream;
  const int taps = ksize * ksize;
  const int64_t per = (int64_t)taps * Cout * Cin;
  static const bool force_mma = getenv("FMDM_WGRAD_MMA") != nullptr;  // A/B switch: the mma.sync kernel everywhere
  if (!force_mma) {
    int tc_splits = 0;
    const int64_t ws_elems = fm_conv_wgrad_workspace_elems(B, Ho, Wo, Cin, Cout, ksize);
    const int rc = wgrad_tc_launch(dy, x, workspace, ws_elems, B, H, W, Cin, Cout, ksize, stride, &tc_splits, st);
    if (rc == 0) {
      launch_wgrad_reduce(workspace, dw, tc_splits, taps, Cout, Cin, cin_total, c_begin, st);
      FM_LAUNCH_CHECK("wgrad_reduce_kernel");
      return 0;
    }
    if (rc != FM_ERR_UNSUPPORTED) return rc;
  }
  WgradParams p;
  int splits, base;
  if (wgrad_plan(B, Ho, Wo, Cin, Cout, ksize, &splits, &p.chunks_per_split, &p.chunks, &base)) {
    set_error("conv_wgrad: empty problem (B*Ho*Wo = %lld)", (long long)B * Ho * Wo);
    return FM_ERR_UNSUPPORTED;
  }
  p.dy = reinterpret_cast<const __nv_bfloat16*>(dy);
  p.x = reinterpret_cast<const __nv_bfloat16*>(x);
  p.part = workspace;
  p.B = B, p.H = H, p.W = W, p.Ho = Ho, p.Wo = Wo, p.Cin = Cin, p.Cout = Cout, p.stride = stride;
  p.pad = ksize / 2;
  p.total_px = (int64_t)B * Ho * Wo;
  p.n_ci = (Cin + wg::kCi - 1) / wg::kCi;
  p.n_co = (Cout + wg::kCo - 1) / wg::kCo;
  dim3 grid(base, splits);
  if (ksize == 3) {
    constexpr int smem = wg::kStages * wg::stage_bytes(3);
    static bool attr = false;
    if (!attr) {
      if (int e = check_cuda(cudaFuncSetAttribute(conv_wgrad_mma_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem), "wgrad attr")) return e;
      attr = true;
    }
    conv_wgrad_mma_kernel<3><<<grid, 256, smem, st>>>(p);
  } else {
    constexpr int smem = wg::kStages * wg::stage_bytes(1);
    static bool attr = false;
    if (!attr) {
      if (int e = check_cuda(cudaFuncSetAttribute(conv_wgrad_mma_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem), "wgrad attr")) return e;
      attr = true;
    }
    conv_wgrad_mma_kernel<1><<<grid, 256, smem, st>>>(p);
  }
  FM_LAUNCH_CHECK("conv_wgrad_mma_kernel");
  launch_wgrad_reduce(workspace, dw, splits, taps, Cout, Cin, cin_total, c_begin, st);
  FM_LAUNCH_CHECK("wgrad_reduce_kernel");
  return 0;
}

static int colsum_blocks(int64_t HW, int B) {
  int nblk = (int)((4LL * sm_count() + B - 1) / B);
  if (nblk > HW) nblk = (int)HW;
  if (nblk < 1) nblk = 1;
  return nblk;
}

extern "C" int64_t fm_colsum_workspace_elems(int32_t B, int64_t HW, int32_t C) {
  if (ensure_device()) return 0;
  return (int64_t)B * colsum_blocks(HW, B) * C;
}

static int launch_colsum_final(const float* part, float* out, float* total, int B, int nblk, int C, int ld, int* tickets,
                               cudaStream_t st) {
  FM_REQUIRE(total == nullptr || tickets != nullptr, "colsum: the batch total needs the ticket buffer");
  FM_REQUIRE((C + 31) / 32 <= kTicketInts - kTicketCols, "colsum: too many channels");
  launch_pdl(colsum_final_kernel, dim3((C + 31) / 32, B), dim3(32, 8), 0, st, part, out, total, B, nblk, C, ld, tickets);
  FM_LAUNCH_CHECK("colsum_final_kernel");
  return 0;
}

/* out[b][c] = sum_p dy[b][p][c]; total (or NULL)[c] = sum_b out[b][c] */
extern "C" int fm_colsum_bf16(const void* dy, float* workspace, float* out, float* total, int32_t B, int64_t HW,
                              int32_t C, int32_t* tickets, fm_stream_t stream) {
  if (int e = ensure_device()) return e;
  FM_REQUIRE(dy && workspace && out, "colsum: null pointer");
  FM_REQUIRE(C % 8 == 0 && C <= 2048, "colsum: C must be a multiple of 8, at most 2048");
  cudaStream_t st = (cudaStream_t)stream;
  const int nblk = colsum_blocks(HW, B);
  const int rows = (int)((HW + nblk - 1) / nblk);
  const int C8 = C / 8;
  const int lanes = C8 >= 256 ? 1 : 256 / C8;
  const int threads = ((C8 * lanes + 31) / 32) * 32;
  launch_pdl(colsum_partial_kernel, dim3(nblk, B), dim3(threads), (size_t)lanes * C * sizeof(float), st, 
      reinterpret_cast<const uint4*>(dy), workspace, HW, C8, lanes, rows);
  FM_LAUNCH_CHECK("colsum_partial_kernel");
  return launch_colsum_final(workspace, out, total, B, nblk, C, C, tickets, st);
}

extern "C" int fm_zero_insert2x_bf16(const void* x, void* out, int32_t B, int32_t H, int32_t W, int32_t C,
                                     fm_stream_t stream) {
  if (int e = ensure_device()) return e;
  FM_REQUIRE(x && out && C % 8 == 0, "zero_insert2x: null pointer or C not a multiple of 8");
  const int64_t total = (int64_t)B * 4 * H * W * (C / 8);
  if (total == 0) return 0;
  launch_pdl(zero_insert2x_kernel, dim3(ew_grid(total)), dim3(256), 0, (cudaStream_t)stream, reinterpret_cast<const uint4*>(x),
                                                                       reinterpret_cast<uint4*>(out), B, H, W, C / 8);
  FM_LAUNCH_CHECK("zero_insert2x_kernel");
  return 0;
}

/* x: [B][2H][2W][C] -> out [B][H][W][C] */
extern "C" int fm_sumpool2x2_bf16(const void* x, void* out, int32_t B, int32_t H, int32_t W, int32_t C,
                                  fm_stream_t stream) {
  if (int e = ensure_device()) return e;
  FM_REQUIRE(x && out && C % 8 == 0, "sumpool2x2: null pointer or C not a multiple of 8");
  const int64_t total = (int64_t)B * H * W * (C / 8);
  if (total == 0) return 0;
  launch_pdl(sumpool2x2_kernel, dim3(ew_grid(total)), dim3(256), 0, (cudaStream_t)stream, reinterpret_cast<const uint4*>(x),
                                                                    reinterpret_cast<uint4*>(out), B, H, W, C / 8);
  FM_LAUNCH_CHECK("sumpool2x2_kernel");
  return 0;
}

static int gn_bwd_blocks(int64_t HW, int B) {
  static const int per_sm = getenv("FMDM_GN_BWD_CTAS_PER_SM") ? atoi(getenv("FMDM_GN_BWD_CTAS_PER_SM")) : 8;
  int nblk = (int)(((int64_t)per_sm * sm_count() + B - 1) / B);  // CTAs per SM across the batch (A/B: env)
  if (nblk > HW / 16) nblk = (int)(HW / 16);
  if (nblk < 1) nblk = 1;
  return nblk;
}

extern "C" int32_t fm_groupnorm_bwd_blocks(int32_t B, int64_t HW) {
  if (ensure_device()) return 0;
  return gn_bwd_blocks(HW, B);
}

/* out[b][c] = sum_blk partials[(b*nblk + blk)*ld + c]; total (or NULL)[c] = sum_b out[b][c]  (second stage of
 * fm_colsum_bf16, also fed by fm_groupnorm_bwd_bf16's dx_colsum_partials, whose rows hold all C0+C1 channels) */
extern "C" int fm_colsum_finish_f32(const float* partials, float* out, float* total, int32_t B, int32_t nblk, int32_t C,
                                    int32_t ld, int32_t* tickets, fm_stream_t stream) {
  if (int e = ensure_device()) return e;
  FM_REQUIRE(partials && out && B > 0 && nblk > 0 && C > 0 && ld >= C, "colsum_finish: bad argument");
  return launch_colsum_final(partials, out, total, B, nblk, C, ld, tickets, (cudaStream_t)stream);
}

extern "C" int32_t fm_ticket_ints(void) { return kTicketInts; }

// =============================================================================================================
// Backward of the cross-attention context path (fm_context_kv_bf16, token-major): kv[b][t][o] = bias[o] + sum_c W[o][c] n,
// n = (ctx[b][c][t] - mean) rstd gamma[c] + beta[c].  Tiny (Cc <= 16): two CUDA-core kernels with fixed-order folds.
//   dW[o][c] = sum_{b,t} dkv n,  dbias[o] = sum dkv          (thread = output feature, block = 64 tokens of a sample)
//   dgamma[c] = sum_{b,t} (sum_o dkv W[o][c]) xhat,  dbeta[c] = sum (sum_o dkv W[o][c])       (thread = token)
// =============================================================================================================
constexpr int kCtxBwdMaxC = 16;
constexpr int kCtxBwdTok = 64;

__global__ void __launch_bounds__(256) context_kv_bwd_w_kernel(const float* __restrict__ ctx,
                                                                const float* __restrict__ stats,
                                                                const float* __restrict__ gamma,
                                                                const float* __restrict__ beta,
                                                                const __nv_bfloat16* __restrict__ dkv,
                                                                float* __restrict__ part, int Cc, int Tc, int O,
                                                                int groups, int chunks_per_sample) {
  __shared__ float sn[kCtxBwdMaxC][kCtxBwdTok];
  const int chunk = blockIdx.x, b = chunk / chunks_per_sample, t0 = (chunk % chunks_per_sample) * kCtxBwdTok;
  const int cpg = Cc / groups;
  for (int i = threadIdx.x; i < Cc * kCtxBwdTok; i += blockDim.x) {
    const int c = i / kCtxBwdTok, tt = i - c * kCtxBwdTok;
    float v = 0.f;
    if (t0 + tt < Tc) {
      const int g = c / cpg;
      const float mean = stats[((int64_t)b * groups + g) * 2 + 0], rstd = stats[((int64_t)b * groups + g) * 2 + 1];
      v = fmaf((ctx[((int64_t)b * Cc + c) * Tc + t0 + tt] - mean) * rstd, gamma[c], beta[c]);
    }
    sn[c][tt] = v;
  }
  __syncthreads();
  const int ntok = min(kCtxBwdTok, Tc - t0);
  // part[chunk][o][Cc + 1]: the weight-gradient row of feature o, then its bias gradient
  for (int o = threadIdx.x; o < O; o += blockDim.x) {
    float acc[kCtxBwdMaxC + 1];
#pragma unroll
    for (int c = 0; c <= kCtxBwdMaxC; ++c) acc[c] = 0.f;
    for (int tt = 0; tt < ntok; ++tt) {
      const float d = __bfloat162float(dkv[((int64_t)b * Tc + t0 + tt) * O + o]);
#pragma unroll
      for (int c = 0; c < kCtxBwdMaxC; ++c)
        if (c < Cc) acc[c] = fmaf(d, sn[c][tt], acc[c]);
      acc[kCtxBwdMaxC] += d;
    }
    float* dst = part + ((int64_t)chunk * O + o) * (Cc + 1);
#pragma unroll
    for (int c = 0; c < kCtxBwdMaxC; ++c)
      if (c < Cc) dst[c] = acc[c];
    dst[Cc] = acc[kCtxBwdMaxC];
  }
}

// dwb[o][Cc + 1] -> dW [O][Cc], dbias [O]
__global__ void context_kv_bwd_split_kernel(const float* __restrict__ dwb, float* __restrict__ dW,
                                            float* __restrict__ dbias, int O, int Cc) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= O * (Cc + 1)) return;
  const int o = i / (Cc + 1), c = i - o * (Cc + 1);
  if (c < Cc) dW[o * Cc + c] = dwb[i];
  else if (dbias != nullptr) dbias[o] = dwb[i];
}

__global__ void __launch_bounds__(128) context_kv_bwd_norm_kernel(const float* __restrict__ ctx,
                                                                   const float* __restrict__ stats,
                                                                   const float* __restrict__ W,
                                                                   const __nv_bfloat16* __restrict__ dkv,
                                                                   float* __restrict__ part, int Cc, int Tc, int O,
                                                                   int groups, int64_t tokens) {
  extern __shared__ float sW[];                 // [O][Cc]
  __shared__ float red[4][2 * kCtxBwdMaxC];
  for (int i = threadIdx.x; i < O * Cc; i += blockDim.x) sW[i] = W[i];
  __syncthreads();
  const int64_t tok = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  float dn[kCtxBwdMaxC], dg[kCtxBwdMaxC];
#pragma unroll
  for (int c = 0; c < kCtxBwdMaxC; ++c) dn[c] = dg[c] = 0.f;
  if (tok < tokens) {
    const int b = (int)(tok / Tc), t = (int)(tok - (int64_t)b * Tc);
    const __nv_bfloat16* row = dkv + tok * O;
    for (int o = 0; o < O; ++o) {
      const float d = __bfloat162float(row[o]);
#pragma unroll
      for (int c = 0; c < kCtxBwdMaxC; ++c)
        if (c < Cc) dn[c] = fmaf(d, sW[o * Cc + c], dn[c]);
    }
    const int cpg = Cc / groups;
#pragma unroll
    for (int c = 0; c < kCtxBwdMaxC; ++c) {
      if (c < Cc) {
        const int g = c / cpg;
        const float mean = stats[((int64_t)b * groups + g) * 2 + 0], rstd = stats[((int64_t)b * groups + g) * 2 + 1];
        dg[c] = dn[c] * (ctx[((int64_t)b * Cc + c) * Tc + t] - mean) * rstd;
      }
    }
  }
  // fixed-order fold over the block's tokens: lanes by shuffle, the four warps through shared memory
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int c = 0; c < kCtxBwdMaxC; ++c) {
    const float a = warp_sum(dg[c]), bsum = warp_sum(dn[c]);
    if (lane == 0) {
      red[warp][c] = a;
      red[warp][kCtxBwdMaxC + c] = bsum;
    }
  }
  __syncthreads();
  if (threadIdx.x < 2 * kCtxBwdMaxC) {
    const float v = (red[0][threadIdx.x] + red[1][threadIdx.x]) + (red[2][threadIdx.x] + red[3][threadIdx.x]);
    part[(int64_t)blockIdx.x * 2 * kCtxBwdMaxC + threadIdx.x] = v;   // [block][dgamma(16) | dbeta(16)]
  }
}

__global__ void context_kv_bwd_gb_kernel(const float* __restrict__ folded, float* __restrict__ dgamma,
                                         float* __restrict__ dbeta, int Cc) {
  const int c = threadIdx.x;
  if (c < Cc) {
    dgamma[c] = folded[c];
    dbeta[c] = folded[kCtxBwdMaxC + c];
  }
}

extern "C" int64_t fm_context_kv_bwd_workspace_elems(int32_t B, int32_t Cc, int32_t Tc, int32_t O) {
  const int64_t chunks = (int64_t)B * ((Tc + kCtxBwdTok - 1) / kCtxBwdTok);
  const int64_t nb = ((int64_t)B * Tc + 127) / 128;
  return chunks * O * (Cc + 1) + (int64_t)O * (Cc + 1) + nb * 2 * kCtxBwdMaxC + 2 * kCtxBwdMaxC;
}

extern "C" int fm_context_kv_bwd_f32(const float* ctx, const float* stats, const float* gamma, const float* beta,
                                     const float* W, const void* dkv, float* workspace, float* dW, float* dbias,
                                     float* dgamma, float* dbeta, int32_t B, int32_t Cc, int32_t Tc, int32_t O,
                                     int32_t groups, fm_stream_t stream) {
  if (int e = ensure_device()) return e;
  FM_REQUIRE(ctx && stats && gamma && beta && W && dkv && workspace && dW && dgamma && dbeta,
             "context_kv_bwd: null pointer");
  FM_REQUIRE(Cc >= 1 && Cc <= kCtxBwdMaxC && groups > 0 && Cc % groups == 0 && B > 0 && Tc > 0 && O > 0,
             "context_kv_bwd: bad shape (context_dim <= 16)");
  FM_REQUIRE((size_t)O * Cc * sizeof(float) <= 160 * 1024, "context_kv_bwd: O * Cc too large");
  cudaStream_t st = (cudaStream_t)stream;
  const int cps = (Tc + kCtxBwdTok - 1) / kCtxBwdTok;
  const int chunks = B * cps;
  float* part = workspace;
  float* dwb = part + (int64_t)chunks * O * (Cc + 1);
  const int64_t tokens = (int64_t)B * Tc;
  const int nb = (int)((tokens + 127) / 128);
  float* npart = dwb + (int64_t)O * (Cc + 1);
  float* nfold = npart + (int64_t)nb * 2 * kCtxBwdMaxC;
  context_kv_bwd_w_kernel<<<chunks, 256, 0, st>>>(ctx, stats, gamma, beta, reinterpret_cast<const __nv_bfloat16*>(dkv),
                                                  part, Cc, Tc, O, groups, cps);
  FM_LAUNCH_CHECK("context_kv_bwd_w_kernel");
  if (int e = launch_reduce(part, dwb, 1, chunks, (int64_t)O * (Cc + 1), st)) return e;
  context_kv_bwd_split_kernel<<<(O * (Cc + 1) + 255) / 256, 256, 0, st>>>(dwb, dW, dbias, O, Cc);
  FM_LAUNCH_CHECK("context_kv_bwd_split_kernel");
  const size_t smem = (size_t)O * Cc * sizeof(float);
  static size_t attr = 48 * 1024;
  if (smem > attr) {
    if (int e = check_cuda(cudaFuncSetAttribute(context_kv_bwd_norm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                (int)smem), "context_kv_bwd attr")) return e;
    attr = smem;
  }
  context_kv_bwd_norm_kernel<<<nb, 128, smem, st>>>(ctx, stats, W, reinterpret_cast<const __nv_bfloat16*>(dkv), npart,
                                                    Cc, Tc, O, groups, tokens);
  FM_LAUNCH_CHECK("context_kv_bwd_norm_kernel");
  if (int e = launch_reduce(npart, nfold, 1, nb, 2 * kCtxBwdMaxC, st)) return e;
  context_kv_bwd_gb_kernel<<<1, 32, 0, st>>>(nfold, dgamma, dbeta, Cc);
  FM_LAUNCH_CHECK("context_kv_bwd_gb_kernel");
  return 0;
}



extern "C" int64_t fm_groupnorm_bwd_workspace_elems(int32_t B, int64_t HW, int32_t C) {
  if (ensure_device()) return 0;
  /* table [B][C][4] + partials [B][nblk][2][C] + dgb_part [B][2][C] */
  return (int64_t)B * C * kGnTab + (int64_t)B * gn_bwd_blocks(HW, B) * 2 * C + 2LL * B * C;
}

extern "C" int fm_groupnorm_bwd_bf16(const void* x0, int32_t C0, const void* x1, int32_t C1, const void* dout,
                                     const float* stats, const float* gamma, const float* beta,
                                     const float* scale_shift, int64_t ss_stride, int32_t silu, int32_t B, int64_t HW,
                                     int32_t groups, float* workspace, void* dx0, void* dx1, float* dgamma_dbeta,
                                     float* dscale_shift, float* dx_colsum_partials, float* dbeta, const void* add0_a,
                                     const void* add0_b, const void* add1, int32_t* tickets, fm_stream_t stream) {
  const int32_t C = C0 + C1;
  if (int e = ensure_device()) return e;
  FM_REQUIRE(x0 && dout && stats && gamma && beta && workspace && dx0 && dgamma_dbeta && tickets,
             "groupnorm_bwd: null pointer");
  FM_REQUIRE(C0 % 8 == 0 && C1 % 8 == 0 && groups > 0 && C % groups == 0,
             "groupnorm_bwd: channel counts must be multiples of 8 and C of groups");
  FM_REQUIRE((C1 == 0) == (x1 == nullptr) && (C1 == 0) == (dx1 == nullptr), "groupnorm_bwd: second source mismatch");
  FM_REQUIRE((scale_shift == nullptr) == (dscale_shift == nullptr), "groupnorm_bwd: scale_shift / dscale_shift mismatch");
  FM_REQUIRE(add1 == nullptr || x1 != nullptr, "groupnorm_bwd: add1 without a second source");
  FM_REQUIRE(add0_b == nullptr || add0_a != nullptr, "groupnorm_bwd: add0_b without add0_a");
  FM_REQUIRE(B < kTicketGlobal, "groupnorm_bwd: batch too large for the ticket buffer");
  cudaStream_t st = (cudaStream_t)stream;
  const int nblk = gn_bwd_blocks(HW, B);
  const int rows = (int)((HW + nblk - 1) / nblk);
  float* tab = workspace;
  float* part = tab + (int64_t)B * C * kGnTab;
  float* dgb = part + (int64_t)B * nblk * 2 * C;  // per-sample (dgamma, dbeta) contributions [B][2][C]
  const int C8 = C / 8;
  const int lanes = C8 >= 256 ? 1 : 256 / C8;
  const int sthreads = ((C8 * lanes + 31) / 32) * 32;
  FM_REQUIRE(C8 <= 256, "groupnorm_bwd: C must be <= 2048");
  GnBwdTail tl;
  tl.stats = stats, tl.gamma = gamma, tl.beta = beta, tl.ss = scale_shift, tl.ss_stride = ss_stride;
  tl.groups = groups;
  tl.inv_n = 1.f / ((float)HW * (float)(C / groups));
  tl.tab = tab, tl.dgb_part = dgb, tl.dss = dscale_shift, tl.dgamma = dgamma_dbeta, tl.dbeta = dbeta;
  tl.tickets = tickets, tl.B = B;
  size_t smem1 = (size_t)lanes * C8 * 16 * sizeof(float);
  if (smem1 < (size_t)4 * C * sizeof(float)) smem1 = (size_t)4 * C * sizeof(float);
  launch_pdl(gn_bwd_partial_kernel, dim3(nblk, B), dim3(sthreads), smem1, st, 
      reinterpret_cast<const uint4*>(x0), C0 / 8, reinterpret_cast<const uint4*>(x1),
      reinterpret_cast<const uint4*>(dout), part, HW, C8, lanes, rows, silu, tl);
  FM_LAUNCH_CHECK("gn_bwd_partial_kernel");
  auto apply = (add0_a || add1) ? gn_bwd_apply_kernel<true> : gn_bwd_apply_kernel<false>;
  launch_pdl(apply, dim3(nblk, B), dim3(sthreads), dx_colsum_partials ? (size_t)lanes * C * sizeof(float) : 0, st,
      reinterpret_cast<const uint4*>(x0), C0 / 8, reinterpret_cast<const uint4*>(x1),
      reinterpret_cast<const uint4*>(dout), tab, reinterpret_cast<uint4*>(dx0), reinterpret_cast<uint4*>(dx1), HW, C8,
      lanes, rows, silu, dx_colsum_partials, reinterpret_cast<const uint4*>(add0_a),
      reinterpret_cast<const uint4*>(add0_b), reinterpret_cast<const uint4*>(add1), dgb, dgamma_dbeta, dbeta, B);
  FM_LAUNCH_CHECK("gn_bwd_apply_kernel");
  return 0;
}

/* Attention backward with queries and keys of their own length / layout (cross-attention; self-attention is the
 * special case fm_attention_bwd_bf16 forwards here): q, dq use the q strides; k, v, dk, dv the kv strides; o, dout the
 * o strides.  head_dim 8: the mma.sync kernel; 16 / 32 / 64: the CUDA-core kernel. */
static int attention_bwd_scalar(const void* q, const void* k, const void* v, const void* o, const void* dout, void* dq,
                                void* dk, void* dv, int B, int heads, int Tq, int Tk, int head_dim, int64_t qs_b,
                                int64_t qs_h, int64_t qs_t, int64_t ks_b, int64_t ks_h, int64_t ks_t, int64_t os_b,
                                int64_t os_h, int64_t os_t, float scale, cudaStream_t st) {
  // fp32 staging (no unpacking in the T^2 loops) when the four tiles fit in 96 KB, else bf16 staging
  const size_t tiles = (size_t)2 * (Tq + Tk) * head_dim;
  const bool f32 = tiles * 4 + (size_t)2 * Tq * 4 <= 96 * 1024;
  const size_t smem = tiles * (f32 ? 4 : 2) + (size_t)2 * Tq * 4;
  if (smem > 200 * 1024) {
    set_error("attention_bwd: Tq=%d Tk=%d head_dim=%d exceed the shared-memory staging budget", Tq, Tk, head_dim);
    return FM_ERR_UNSUPPORTED;
  }
#define FM_ATT_BWD_ST(HD, ST)                                                                                       \
  {                                                                                                                 \
    static size_t attr_smem = 0;                                                                                    \
    if (smem > attr_smem) {                                                                                         \
      if (int e = check_cuda(cudaFuncSetAttribute(attention_bwd_kernel<HD, ST>,                                     \
                                                  cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem),          \
                             "attention_bwd attr")) return e;                                                       \
      attr_smem = smem;                                                                                             \
    }                                                                                                               \
    attention_bwd_kernel<HD, ST><<<B * heads, 256, smem, st>>>(                                                     \
        (const __nv_bfloat16*)q, (const __nv_bfloat16*)k, (const __nv_bfloat16*)v, (const __nv_bfloat16*)o,         \
        (const __nv_bfloat16*)dout, (__nv_bfloat16*)dq, (__nv_bfloat16*)dk, (__nv_bfloat16*)dv, heads, Tq, Tk, qs_b, \
        qs_h, qs_t, ks_b, ks_h, ks_t, os_b, os_h, os_t, scale);                                                     \
  }
#define FM_ATT_BWD(HD)                                  \
  case HD:                                              \
    if (f32) FM_ATT_BWD_ST(HD, float)                   \
    else FM_ATT_BWD_ST(HD, __nv_bfloat16)               \
    break;
  switch (head_dim) {
    FM_ATT_BWD(8)
    FM_ATT_BWD(16)
    FM_ATT_BWD(32)
    FM_ATT_BWD(64)
    default:
      set_error("attention_bwd: head_dim %d unsupported (8, 16, 32, 64)", head_dim);
      return FM_ERR_UNSUPPORTED;
  }
#undef FM_ATT_BWD_ST
#undef FM_ATT_BWD
  FM_LAUNCH_CHECK("attention_bwd_kernel");
  return 0;
}

extern "C" int fm_attention_bwd_cross_bf16(const void* q, const void* k, const void* v, const void* o, const void* dout,
                                           void* dq, void* dk, void* dv, int32_t B, int32_t heads, int32_t Tq,
                                           int32_t Tk, int32_t head_dim, int64_t qs_b, int64_t qs_h, int64_t qs_t,
                                           int64_t ks_b, int64_t ks_h, int64_t ks_t, int64_t os_b, int64_t os_h,
                                           int64_t os_t, float scale, fm_stream_t stream) {
  if (int e = ensure_device()) return e;
  FM_REQUIRE(q && k && v && o && dout && dq && dk && dv && Tq > 0 && Tk > 0, "attention_bwd: bad argument");
  FM_REQUIRE(((qs_b | qs_h | qs_t | ks_b | ks_h | ks_t | os_b | os_h | os_t) % 8) == 0 &&
                 (((uintptr_t)q | (uintptr_t)k | (uintptr_t)v | (uintptr_t)o | (uintptr_t)dout | (uintptr_t)dq |
                   (uintptr_t)dk | (uintptr_t)dv) & 15) == 0,
             "attention_bwd: rows must be 16-byte aligned (strides multiples of 8 elements)");
  cudaStream_t st = (cudaStream_t)stream;
  static const bool scalar8 = getenv("FMDM_ATTENTION_BWD_SCALAR") != nullptr;  // A/B: the CUDA-core kernel at head_dim 8
  if (head_dim == 8 && !scalar8) {
    const int Tqp = (Tq + 15) & ~15, Tkp = (Tk + 15) & ~15;
    const size_t need = (size_t)(Tqp + Tkp) * 8 * 2 * 2 + (size_t)8 * (Tkp + kAttB8Pad) * 2 +
                        (size_t)2 * 8 * (Tqp + kAttB8Pad) * 2 + (size_t)2 * Tqp * 4;
    if (need <= 200 * 1024) {
      static size_t attr8 = 48 * 1024;
      if (need > attr8) {
        if (int e = check_cuda(cudaFuncSetAttribute(attention_bwd_hd8_mma_kernel,
                                                    cudaFuncAttributeMaxDynamicSharedMemorySize, (int)need),
                               "attention_bwd_hd8 attr")) return e;
        attr8 = need;
      }
      attention_bwd_hd8_mma_kernel<<<B * heads, 256, need, st>>>(
          (const __nv_bfloat16*)q, (const __nv_bfloat16*)k, (const __nv_bfloat16*)v, (const __nv_bfloat16*)o,
          (const __nv_bfloat16*)dout, (__nv_bfloat16*)dq, (__nv_bfloat16*)dk, (__nv_bfloat16*)dv, heads, Tq, Tqp, Tk, Tkp,
          qs_b, qs_h, qs_t, ks_b, ks_h, ks_t, os_b, os_h, os_t, scale);
      FM_LAUNCH_CHECK("attention_bwd_hd8_mma_kernel");
      return 0;
    }
  }
  return attention_bwd_scalar(q, k, v, o, dout, dq, dk, dv, B, heads, Tq, Tk, head_dim, qs_b, qs_h, qs_t, ks_b, ks_h,
                              ks_t, os_b, os_h, os_t, scale, st);
}

// =============================================================================================================
// Linear attention backward (`attention.py:53-70` LinearQKVAttention; forward: fm_linear_attention_bf16):
//   ks = softmax_tokens(k), qs = softmax_features(q), ctx = ks^T v / (sum_n ks + eps), out = qs ctx.
//   dctx = qs^T dout;  dqs = dout ctx^T;  dA = dctx / den;  dks = v dA^T;  dv = ks dA;
//   dq = qs * (dqs - rowsum(qs * dqs));  dk = ks * (dks - colsum(ks * dks))
// (the denominator's own gradient is a per-column constant added to dks, which the column softmax's backward cancels
// exactly because sum_n ks = 1).  O(T d^2) fp32 CUDA-core math, one CTA per (sample, head), d x d matrices in shared
// memory, every sum in a fixed order.
// =============================================================================================================
template <int HD>
__global__ void __launch_bounds__(256) linear_attention_bwd_kernel(
    const __nv_bfloat16* __restrict__ q, const __nv_bfloat16* __restrict__ k, const __nv_bfloat16* __restrict__ v,
    const __nv_bfloat16* __restrict__ dout, __nv_bfloat16* __restrict__ dq, __nv_bfloat16* __restrict__ dk,
    __nv_bfloat16* __restrict__ dv, int heads, int Tq, int Tk, int64_t q_sb, int64_t q_sh, int64_t q_st,
    int64_t kv_sb, int64_t kv_sh, int64_t kv_st, int64_t o_sb, int64_t o_sh, int64_t o_st, float eps) {
  constexpr int kLanes = 256 / HD;   // token lanes per feature column
  constexpr int kChunk = HD == 64 ? 16 : 32;   // tokens staged per step (static shared memory stays under 48 KB)
  constexpr int kPairs = HD * HD / 256 > 0 ? HD * HD / 256 : 1;
  __shared__ float ctx[HD][HD + 1], dA[HD][HD + 1];
  __shared__ float red[256];
  __shared__ float cmax[HD], csum[HD], cden[HD], cS[HD];
  __shared__ float sa[kChunk][HD + 1], sb[kChunk][HD + 1], sc[kChunk][HD + 1];
  __shared__ float rstat[kChunk];
  const int b = blockIdx.x / heads, h = blockIdx.x % heads;
  const __nv_bfloat16* kb = k + b * kv_sb + h * kv_sh;
  const __nv_bfloat16* vb = v + b * kv_sb + h * kv_sh;
  const __nv_bfloat16* qb = q + b * q_sb + h * q_sh;
  const __nv_bfloat16* gb = dout + b * o_sb + h * o_sh;
  __nv_bfloat16* dqb = dq + b * q_sb + h * q_sh;
  __nv_bfloat16* dkb = dk + b * kv_sb + h * kv_sh;
  __nv_bfloat16* dvb = dv + b * kv_sb + h * kv_sh;
  const int d = threadIdx.x % HD, lane = threadIdx.x / HD;

  // ---- column softmax statistics of k ----
  float m = -INFINITY;
  for (int n = lane; n < Tk; n += kLanes) m = fmaxf(m, __bfloat162float(kb[n * kv_st + d]));
  red[threadIdx.x] = m;
  __syncthreads();
  if (lane == 0) {
    for (int l = 1; l < kLanes; ++l) m = fmaxf(m, red[l * HD + d]);
    cmax[d] = m;
  }
  __syncthreads();
  float s = 0.f;
  const float cm = cmax[d];
  for (int n = lane; n < Tk; n += kLanes) s += __expf(__bfloat162float(kb[n * kv_st + d]) - cm);
  red[threadIdx.x] = s;
  __syncthreads();
  if (lane == 0) {
    for (int l = 1; l < kLanes; ++l) s += red[l * HD + d];
    csum[d] = s;
  }
  __syncthreads();

  // ---- forward context: ctx[d][e] = sum_n ks[n][d] v[n][e] / (sum_n ks[n][d] + eps) ----
  float acc[kPairs];
#pragma unroll
  for (int j = 0; j < kPairs; ++j) acc[j] = 0.f;
  float den = 0.f;
  for (int n0 = 0; n0 < Tk; n0 += kChunk) {
    __syncthreads();
    for (int i = threadIdx.x; i < kChunk * HD; i += 256) {
      const int nn = i / HD, dd = i - nn * HD;
      float kv_ = 0.f, vv = 0.f;
      if (n0 + nn < Tk) {
        kv_ = __expf(__bfloat162float(kb[(n0 + nn) * kv_st + dd]) - cmax[dd]) / csum[dd];
        vv = __bfloat162float(vb[(n0 + nn) * kv_st + dd]);
      }
      sa[nn][dd] = kv_;
      sb[nn][dd] = vv;
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < kPairs; ++j) {
      const int p = threadIdx.x + 256 * j;
      if (p < HD * HD) {
        const int dd = p / HD, ee = p - dd * HD;
        float a = acc[j];
#pragma unroll 8
        for (int nn = 0; nn < kChunk; ++nn) a = fmaf(sa[nn][dd], sb[nn][ee], a);
        acc[j] = a;
      }
    }
    if (threadIdx.x < HD)
      for (int nn = 0; nn < kChunk; ++nn) den += sa[nn][threadIdx.x];
  }
  if (threadIdx.x < HD) cden[threadIdx.x] = den + eps;
  __syncthreads();
#pragma unroll
  for (int j = 0; j < kPairs; ++j) {
    const int p = threadIdx.x + 256 * j;
    if (p < HD * HD) {
      const int dd = p / HD, ee = p - dd * HD;
      ctx[dd][ee] = acc[j] / cden[dd];
      acc[j] = 0.f;  // re-used for dctx
    }
  }
  __syncthreads();

  // ---- queries: dctx += qs^T dout;  dq = qs * (dout ctx^T - rowsum) ----
  for (int n0 = 0; n0 < Tq; n0 += kChunk) {
    __syncthreads();
    for (int i = threadIdx.x; i < kChunk * HD; i += 256) {
      const int nn = i / HD, dd = i - nn * HD;
      const bool in = n0 + nn < Tq;
      sa[nn][dd] = in ? __bfloat162float(qb[(n0 + nn) * q_st + dd]) : 0.f;   // raw q, normalised below
      sb[nn][dd] = in ? __bfloat162float(gb[(n0 + nn) * o_st + dd]) : 0.f;   // dout
    }
    __syncthreads();
    if (threadIdx.x < kChunk) {  // row softmax of q over the HD features (fixed order)
      const int nn = threadIdx.x;
      float mx = -INFINITY, sm = 0.f;
      for (int dd = 0; dd < HD; ++dd) mx = fmaxf(mx, sa[nn][dd]);
      for (int dd = 0; dd < HD; ++dd) sm += __expf(sa[nn][dd] - mx);
      const float inv = (n0 + nn < Tq) ? 1.f / sm : 0.f;
      for (int dd = 0; dd < HD; ++dd) sa[nn][dd] = __expf(sa[nn][dd] - mx) * inv;   // qs (zero rows past Tq)
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < kPairs; ++j) {
      const int p = threadIdx.x + 256 * j;
      if (p < HD * HD) {
        const int dd = p / HD, ee = p - dd * HD;
        float a = acc[j];
#pragma unroll 8
        for (int nn = 0; nn < kChunk; ++nn) a = fmaf(sa[nn][dd], sb[nn][ee], a);
        acc[j] = a;
      }
    }
    // dqs[nn][d] = sum_e dout[nn][e] ctx[d][e]
    for (int nn = lane; nn < kChunk; nn += kLanes) {
      float a = 0.f;
#pragma unroll 8
      for (int ee = 0; ee < HD; ++ee) a = fmaf(sb[nn][ee], ctx[d][ee], a);
      sc[nn][d] = a;
    }
    __syncthreads();
    if (threadIdx.x < kChunk) {
      const int nn = threadIdx.x;
      float dot = 0.f;
      for (int dd = 0; dd < HD; ++dd) dot = fmaf(sa[nn][dd], sc[nn][dd], dot);
      rstat[nn] = dot;
    }
    __syncthreads();
    for (int nn = lane; nn < kChunk; nn += kLanes)
      if (n0 + nn < Tq) dqb[(n0 + nn) * q_st + d] = __float2bfloat16_rn(sa[nn][d] * (sc[nn][d] - rstat[nn]));
  }
  __syncthreads();
#pragma unroll
  for (int j = 0; j < kPairs; ++j) {
    const int p = threadIdx.x + 256 * j;
    if (p < HD * HD) {
      const int dd = p / HD, ee = p - dd * HD;
      dA[dd][ee] = acc[j] / cden[dd];
    }
  }
  __syncthreads();

  // ---- keys, pass a: S[d] = sum_n ks[n][d] dks[n][d],  dks[n][d] = sum_e v[n][e] dA[d][e];  dv = ks dA ----
  float sp = 0.f;
  for (int n0 = 0; n0 < Tk; n0 += kChunk) {
    __syncthreads();
    for (int i = threadIdx.x; i < kChunk * HD; i += 256) {
      const int nn = i / HD, dd = i - nn * HD;
      float kv_ = 0.f, vv = 0.f;
      if (n0 + nn < Tk) {
        kv_ = __expf(__bfloat162float(kb[(n0 + nn) * kv_st + dd]) - cmax[dd]) / csum[dd];
        vv = __bfloat162float(vb[(n0 + nn) * kv_st + dd]);
      }
      sa[nn][dd] = kv_;
      sb[nn][dd] = vv;
    }
    __syncthreads();
    for (int nn = lane; nn < kChunk; nn += kLanes) {
      float a = 0.f, dvv = 0.f;
#pragma unroll 8
      for (int ee = 0; ee < HD; ++ee) {
        a = fmaf(sb[nn][ee], dA[d][ee], a);       // dks[nn][d]
        dvv = fmaf(sa[nn][ee], dA[ee][d], dvv);   // dv[nn][d] = sum_e ks[nn][e] dA[e][d]
      }
      sc[nn][d] = a;
      sp = fmaf(sa[nn][d], a, sp);
      if (n0 + nn < Tk) dvb[(n0 + nn) * kv_st + d] = __float2bfloat16_rn(dvv);
    }
  }
  red[threadIdx.x] = sp;
  __syncthreads();
  if (lane == 0) {
    float t = 0.f;
    for (int l = 0; l < kLanes; ++l) t += red[l * HD + d];
    cS[d] = t;
  }
  __syncthreads();
  // ---- keys, pass b: dk = ks * (dks - S) ----
  for (int n0 = 0; n0 < Tk; n0 += kChunk) {
    __syncthreads();
    for (int i = threadIdx.x; i < kChunk * HD; i += 256) {
      const int nn = i / HD, dd = i - nn * HD;
      float kv_ = 0.f, vv = 0.f;
      if (n0 + nn < Tk) {
        kv_ = __expf(__bfloat162float(kb[(n0 + nn) * kv_st + dd]) - cmax[dd]) / csum[dd];
        vv = __bfloat162float(vb[(n0 + nn) * kv_st + dd]);
      }
      sa[nn][dd] = kv_;
      sb[nn][dd] = vv;
    }
    __syncthreads();
    for (int nn = lane; nn < kChunk; nn += kLanes) {
      float a = 0.f;
#pragma unroll 8
      for (int ee = 0; ee < HD; ++ee) a = fmaf(sb[nn][ee], dA[d][ee], a);
      if (n0 + nn < Tk) dkb[(n0 + nn) * kv_st + d] = __float2bfloat16_rn(sa[nn][d] * (a - cS[d]));
    }
  }
}

extern "C" int fm_linear_attention_bwd_bf16(const void* q, const void* k, const void* v, const void* dout, void* dq,
                                            void* dk, void* dv, int32_t B, int32_t heads, int32_t Tq, int32_t Tk,
                                            int32_t head_dim, int64_t q_sb, int64_t q_sh, int64_t q_st, int64_t kv_sb,
                                            int64_t kv_sh, int64_t kv_st, int64_t o_sb, int64_t o_sh, int64_t o_st,
                                            float eps, fm_stream_t stream) {
  if (int e = ensure_device()) return e;
  FM_REQUIRE(q && k && v && dout && dq && dk && dv && B > 0 && heads > 0 && Tq > 0 && Tk > 0,
             "linear_attention_bwd: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
#define FM_LINB(HD)                                                                                                  \
  case HD:                                                                                                          \
    linear_attention_bwd_kernel<HD><<<B * heads, 256, 0, st>>>(                                                     \
        (const __nv_bfloat16*)q, (const __nv_bfloat16*)k, (const __nv_bfloat16*)v, (const __nv_bfloat16*)dout,      \
        (__nv_bfloat16*)dq, (__nv_bfloat16*)dk, (__nv_bfloat16*)dv, heads, Tq, Tk, q_sb, q_sh, q_st, kv_sb, kv_sh,  \
        kv_st, o_sb, o_sh, o_st, eps);                                                                              \
    break;
  switch (head_dim) {
    FM_LINB(8)
    FM_LINB(16)
    FM_LINB(32)
    FM_LINB(64)
    default:
      set_error("linear_attention_bwd: head_dim=%d unsupported (8, 16, 32, 64)", head_dim);
      return FM_ERR_UNSUPPORTED;
  }
#undef FM_LINB
  FM_LAUNCH_CHECK("linear_attention_bwd_kernel");
  return 0;
}

extern "C" int fm_attention_bwd_bf16(const void* q, const void* k, const void* v, const void* o, const void* dout,
                                     void* dq, void* dk, void* dv, int32_t B, int32_t heads, int32_t T,
                                     int32_t head_dim, int64_t qs_b, int64_t qs_h, int64_t qs_t, int64_t os_b,
                                     int64_t os_h, int64_t os_t, float scale, fm_stream_t stream) {
  return fm_attention_bwd_cross_bf16(q, k, v, o, dout, dq, dk, dv, B, heads, T, T, head_dim, qs_b, qs_h, qs_t, qs_b,
                                     qs_h, qs_t, os_b, os_h, os_t, scale, stream);
}

extern "C" int fm_silu_bwd_f32(const float* x, const float* dy, float* dx, int64_t n, fm_stream_t stream) {
  if (int e = ensure_device()) return e;
  FM_REQUIRE(x && dy && dx && n >= 0, "silu_bwd: bad argument");
  if (n == 0) return 0;
  silu_bwd_kernel<<<ew_grid(n), 256, 0, (cudaStream_t)stream>>>(x, dy, dx, n);
  FM_LAUNCH_CHECK("silu_bwd_kernel");
  return 0;
}

static int stem_blocks() { return 4 * sm_count(); }

extern "C" int64_t fm_conv_stem_wgrad_workspace_elems(int32_t Cin, int32_t Cout) {
  if (ensure_device()) return 0;
  return (int64_t)stem_blocks() * Cout * Cin * 9;
}

extern "C" int fm_conv_stem_wgrad_f32(const float* x0, int32_t C0, const float* x1, int32_t C1, float in_scale,
                                      float in_shift, const void* dy, float* workspace, float* dw, int32_t B, int32_t H,
                                      int32_t W, int32_t Cout, fm_stream_t stream) {
  if (int e = ensure_device()) return e;
  FM_REQUIRE(x0 && dy && workspace && dw, "stem_wgrad: null pointer");
  FM_REQUIRE((C1 == 0) == (x1 == nullptr), "stem_wgrad: x1 / C1 mismatch");
  const int cin = C0 + C1;
  FM_REQUIRE(cin >= 1 && cin <= 4, "stem_wgrad: 1..4 input channels");
  FM_REQUIRE(Cout >= 1 && Cout <= 256, "stem_wgrad: Cout must be <= 256");
  cudaStream_t st = (cudaStream_t)stream;
  const int nblk = stem_blocks();
  const int64_t segs = (int64_t)B * H * ((W + 63) / 64);
  const int spb = (int)((segs + nblk - 1) / nblk);
  switch (cin) {
    case 1: stem_wgrad_kernel<1><<<nblk, 256, 0, st>>>(x0, C0, x1, in_scale, in_shift, (const __nv_bfloat16*)dy, workspace, B, H, W, Cout, spb); break;
    case 2: stem_wgrad_kernel<2><<<nblk, 256, 0, st>>>(x0, C0, x1, in_scale, in_shift, (const __nv_bfloat16*)dy, workspace, B, H, W, Cout, spb); break;
    case 3: stem_wgrad_kernel<3><<<nblk, 256, 0, st>>>(x0, C0, x1, in_scale, in_shift, (const __nv_bfloat16*)dy, workspace, B, H, W, Cout, spb); break;
    default: stem_wgrad_kernel<4><<<nblk, 256, 0, st>>>(x0, C0, x1, in_scale, in_shift, (const __nv_bfloat16*)dy, workspace, B, H, W, Cout, spb); break;
  }
  FM_LAUNCH_CHECK("stem_wgrad_kernel");
  return launch_reduce(workspace, dw, 1, nblk, (int64_t)Cout * cin * 9, st);
}

static int head_blocks() { return 4 * sm_count(); }

extern "C" int64_t fm_conv_head_bwd_workspace_elems(int32_t Cin) {
  if (ensure_device()) return 0;
  return (int64_t)head_blocks() * Cin * 9;
}

/* Cout == 1: da (bf16 NHWC, or NULL), dw fp32 [1][Cin][3][3] */
extern "C" int fm_conv_head_bwd_f32(const void* a, const float* dy, const float* weight, float* workspace, void* da,
                                    float* dw, int32_t B, int32_t H, int32_t W, int32_t Cin, fm_stream_t stream) {
  if (int e = ensure_device()) return e;
  FM_REQUIRE(a && dy && weight && workspace && dw, "head_bwd: null pointer");
  FM_REQUIRE(Cin % 8 == 0 && Cin <= 512 && 512 % Cin == 0, "head_bwd: Cin must divide 512 and be a multiple of 8");
  cudaStream_t st = (cudaStream_t)stream;
  if (da != nullptr) {
    const int64_t total = (int64_t)B * H * W * (Cin / 8);
    head_dgrad_kernel<<<ew_grid(total), 256, 9 * Cin * sizeof(float), st>>>(dy, weight, reinterpret_cast<uint4*>(da), B,
                                                                          H, W, Cin);
    FM_LAUNCH_CHECK("head_dgrad_kernel");
  }
  const int nblk = head_blocks();
  const int64_t px = (int64_t)B * H * W;
  const int ppb = (int)((px + nblk - 1) / nblk);
  head_wgrad_kernel<<<nblk, 256, 0, st>>>(reinterpret_cast<const __nv_bfloat16*>(a), dy, workspace, B, H, W, Cin, ppb);
  FM_LAUNCH_CHECK("head_wgrad_kernel");
  return launch_reduce(workspace, dw, 1, nblk, (int64_t)Cin * 9, st);
}

/* out[0] = scale * sum(x)  (mode 0)  or  scale * sum((x - (t1 - t2))^2)  (mode 1); workspace: 1024 doubles */
extern "C" int fm_sum_f32(const float* x, const float* t1, const float* t2, double* workspace, float* out, int64_t n,
                          int32_t mode, double scale, fm_stream_t stream) {
  if (int e = ensure_device()) return e;
  FM_REQUIRE(x && workspace && out && n >= 0, "sum: bad argument");
  FM_REQUIRE(mode == 0 || (mode == 1 && t1 != nullptr), "sum: mode 1 needs a target");
  int blocks = ew_grid(n);
  if (blocks > 1024) blocks = 1024;
  cudaStream_t st = (cudaStream_t)stream;
  sum_partial_kernel<<<blocks, 256, 0, st>>>(x, t1, t2, workspace, n, mode);
  FM_LAUNCH_CHECK("sum_partial_kernel");
  sum_final_kernel<<<1, 32, 0, st>>>(workspace, blocks, scale, out);
  FM_LAUNCH_CHECK("sum_final_kernel");
  return 0;
}

extern "C" int fm_mse_bwd_f32(const float* pred, const float* t1, const float* t2, const float* gscalar, float* dpred,
                              int64_t n, fm_stream_t stream) {
  if (int e = ensure_device()) return e;
  FM_REQUIRE(pred && t1 && gscalar && dpred && n > 0, "mse_bwd: bad argument");
  mse_bwd_kernel<<<ew_grid(n), 256, 0, (cudaStream_t)stream>>>(pred, t1, t2, gscalar, dpred, (float)(2.0 / (double)n), n);
  FM_LAUNCH_CHECK("mse_bwd_kernel");
  return 0;
}

extern "C" int fm_adamw_f32(float* p, const float* g, float* m, float* v, int64_t n, float lr, float beta1, float beta2,
                            float eps, float weight_decay, int64_t step, float grad_scale, fm_stream_t stream) {
  if (int e = ensure_device()) return e;
  FM_REQUIRE(p && g && m && v && n >= 0 && step >= 1, "adamw: bad argument");
  if (n == 0) return 0;
  /* scalar prefactors in double, as torch.optim.adamw._single_tensor_adamw computes them on the host */
  const double bc1 = 1.0 - pow((double)beta1, (double)step);
  const double bc2 = 1.0 - pow((double)beta2, (double)step);
  const float decay = (float)(1.0 - (double)lr * (double)weight_decay);
  const float step_size = (float)((double)lr / bc1);
  const float bc2_sqrt = (float)sqrt(bc2);
  adamw_kernel<<<ew_grid(n), 256, 0, (cudaStream_t)stream>>>(p, g, m, v, n, decay, (float)(1.0 - (double)beta1), beta2, (float)(1.0 - (double)beta2),
                                                            step_size, bc2_sqrt, eps, grad_scale);
  FM_LAUNCH_CHECK("adamw_kernel");
  return 0;
}

extern "C" int64_t fm_linear_bwd_workspace_elems(int32_t B, int32_t I, int32_t O) {
  return (int64_t)((O + kLinChunk - 1) / kLinChunk) * B * I;
}

/* Backward of fm_linear_f32 for B <= 32 rows: dx (or NULL) [B][I], dw [O][I], db (or NULL) [O]; workspace:
 * fm_linear_bwd_workspace_elems floats (only for dx). */
extern "C" int fm_linear_bwd_f32(const float* x, const float* W, const float* dy, float* workspace, float* dx, float* dw,
                                 float* db, int32_t B, int32_t I, int32_t O, int32_t silu_in, fm_stream_t stream) {
  if (int e = ensure_device()) return e;
  FM_REQUIRE(x && W && dy && B > 0 && I > 0 && O > 0, "linear_bwd: bad argument");
  FM_REQUIRE(B <= 32, "linear_bwd: at most 32 rows (got %d)", B);
  FM_REQUIRE(dx == nullptr || workspace != nullptr, "linear_bwd: dx needs the workspace");
  cudaStream_t st = (cudaStream_t)stream;
  if (dw != nullptr) {
    linear_wgrad_kernel<<<dim3((I + 127) / 128, (O + 31) / 32), 256, 0, st>>>(x, dy, dw, db, B, I, O, silu_in);
    FM_LAUNCH_CHECK("linear_wgrad_kernel");
  }
  if (dx != nullptr) {
    const int nchunk = (O + kLinChunk - 1) / kLinChunk;
    linear_dgrad_partial_kernel<<<dim3((I + 127) / 128, nchunk), 256, 0, st>>>(dy, W, workspace, B, I, O);
    FM_LAUNCH_CHECK("linear_dgrad_partial_kernel");
    const int64_t n = (int64_t)B * I;
    linear_dgrad_finish_kernel<<<ew_grid(n), 256, 0, st>>>(workspace, x, dx, nchunk, n, silu_in);
    FM_LAUNCH_CHECK("linear_dgrad_finish_kernel");
  }
  return 0;
}
