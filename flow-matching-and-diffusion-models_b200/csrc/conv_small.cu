// The two degenerate convolutions of the UNet: the stem (Cin = 1..8, K = 9*Cin too small for a tensor-core tile)
// and the head (Cout = 1..4, N too small).  Both touch the largest tensors of the network and are treated as
// bandwidth kernels on CUDA cores.  The stem also absorbs the torch.cat([x, cond], 1) of the sampling loop
// (src/pipelines/utils.py:204-205), the optional 2x-1 centering (unet_diffusers_nd.py:155-157), the fp32->bf16
// cast and the NCHW->NHWC layout change; the head emits the fp32 NCHW prediction the scheduler step consumes.
#include <cstdlib>

#include "common.cuh"

namespace fm {

constexpr int kStemMaxCin = 8;
constexpr int kStemThreads = 256;
constexpr int kStemPix = 4;

// One thread = 8 output channels x kStemPix consecutive pixels of a row; weights live in shared memory as
// [ci*9+tap][co]; a block works on ONE image (grid.y), so the optional GroupNorm partial statistics of the output
// (sum, sum of squares per channel quad, one row per block: stats[(n*gridDim.x + block)][Cout/4][2]) need no atomics.
__global__ void __launch_bounds__(kStemThreads) conv_stem_kernel(const float* __restrict__ x0, int C0,
                                                                const float* __restrict__ x1, int C1,
                                                                float in_scale, float in_shift,
                                                                const float* __restrict__ w_oihw,
                                                                const float* __restrict__ bias,
                                                                uint4* __restrict__ out, int H, int W, int Cout,
                                                                float* __restrict__ stats) {
  extern __shared__ float sw[];  // [Cin*9][Cout] + bias[Cout] (+ statistics scratch, reusing the weights at the end)
  const int Cin = C0 + C1;
  const int K = Cin * 9;
  for (int i = threadIdx.x; i < K * Cout; i += blockDim.x) {
    const int co = i % Cout, k = i / Cout;  // k = ci*9 + tap matches OIHW inner order
    sw[i] = w_oihw[(size_t)co * K + k];
  }
  float* sbias = sw + K * Cout;
  for (int i = threadIdx.x; i < Cout; i += blockDim.x) sbias[i] = bias ? bias[i] : 0.f;
  __syncthreads();

  const int chunks = Cout >> 3;
  const int gpb = kStemThreads / chunks;  // pixel groups per block iteration
  const int c8 = threadIdx.x % chunks;
  const int gl = threadIdx.x / chunks;
  const int n = blockIdx.y;
  const int HW = H * W;
  const int WG = (W + kStemPix - 1) / kStemPix;
  const int groups_per_img = H * WG;
  float bi[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) bi[j] = sbias[c8 * 8 + j];
  float st_s[2] = {0.f, 0.f}, st_q[2] = {0.f, 0.f};  // this thread's two channel quads
  if (gl < gpb) {
    for (int grp = blockIdx.x * gpb + gl; grp < groups_per_img; grp += gridDim.x * gpb) {
      const int h = grp / WG;
      const int wbase = (grp - h * WG) * kStemPix;
      float acc[kStemPix][8];
#pragma unroll
      for (int p = 0; p < kStemPix; ++p)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[p][j] = bi[j];
      for (int ci = 0; ci < Cin; ++ci) {
        const float* src = (ci < C0) ? x0 + ((size_t)n * C0 + ci) * HW : x1 + ((size_t)n * C1 + (ci - C0)) * HW;
        // all 3 x (kStemPix+2) inputs of this channel first (one exposed memory latency per channel, not per row)
        float xin[3][kStemPix + 2];
#pragma unroll
        for (int kh = 0; kh < 3; ++kh) {
          const int ih = h + kh - 1;
#pragma unroll
          for (int q = 0; q < kStemPix + 2; ++q) {
            const int iw = wbase + q - 1;
            // zero padding applies AFTER the optional 2x-1 centering (the reference centres, then convolves)
            xin[kh][q] = ((unsigned)ih < (unsigned)H && (unsigned)iw < (unsigned)W)
                             ? fmaf(__ldg(src + ih * W + iw), in_scale, in_shift) : 0.f;
          }
        }
#pragma unroll
        for (int kh = 0; kh < 3; ++kh) {
#pragma unroll
          for (int kw = 0; kw < 3; ++kw) {
            const float* wp = sw + (ci * 9 + kh * 3 + kw) * Cout + c8 * 8;
            const float4 w0 = *reinterpret_cast<const float4*>(wp);
            const float4 w1 = *reinterpret_cast<const float4*>(wp + 4);
#pragma unroll
            for (int p = 0; p < kStemPix; ++p) {
              const float2 xv = make_float2(xin[kh][p + kw], xin[kh][p + kw]);
              float2* ap = reinterpret_cast<float2*>(acc[p]);  // packed FMAs: two output channels per instruction
              ap[0] = ffma2(xv, make_float2(w0.x, w0.y), ap[0]);
              ap[1] = ffma2(xv, make_float2(w0.z, w0.w), ap[1]);
              ap[2] = ffma2(xv, make_float2(w1.x, w1.y), ap[2]);
              ap[3] = ffma2(xv, make_float2(w1.z, w1.w), ap[3]);
            }
          }
        }
      }
      const size_t pix0 = ((size_t)n * H + h) * W + wbase;
#pragma unroll
      for (int p = 0; p < kStemPix; ++p) {
        if (wbase + p < W) {
          uint4 o;
          o.x = pack_bf16x2(acc[p][0], acc[p][1]); o.y = pack_bf16x2(acc[p][2], acc[p][3]);
          o.z = pack_bf16x2(acc[p][4], acc[p][5]); o.w = pack_bf16x2(acc[p][6], acc[p][7]);
          out[(pix0 + p) * chunks + c8] = o;
          if (stats != nullptr) {
#pragma unroll
            for (int hq = 0; hq < 2; ++hq) {
              const float a0 = acc[p][4 * hq], a1 = acc[p][4 * hq + 1], a2 = acc[p][4 * hq + 2], a3 = acc[p][4 * hq + 3];
              st_s[hq] += (a0 + a1) + (a2 + a3);
              st_q[hq] += fmaf(a0, a0, a1 * a1) + fmaf(a2, a2, a3 * a3);
            }
          }
        }
      }
    }
  }
  if (stats == nullptr) return;
  // fold the block's threads per channel quad in a fixed order (deterministic), one row of partials per block
  __syncthreads();           // the weights in shared memory are no longer needed
  float* red = sw;           // [gpb][chunks*2][2]
  if (gl < gpb) {
#pragma unroll
    for (int hq = 0; hq < 2; ++hq) {
      red[((gl * chunks + c8) * 2 + hq) * 2 + 0] = st_s[hq];
      red[((gl * chunks + c8) * 2 + hq) * 2 + 1] = st_q[hq];
    }
  }
  __syncthreads();
  const int nq = Cout >> 2;
  for (int i = threadIdx.x; i < nq * 2; i += blockDim.x) {
    float t = 0.f;
    for (int g = 0; g < gpb; ++g) t += red[g * nq * 2 + i];
    stats[((size_t)n * gridDim.x + blockIdx.x) * nq * 2 + i] = t;
  }
}

// Stem on the tensor cores, step 1: the 3x3 neighbourhoods of the (tiny-Cin) fp32 NCHW inputs as a bf16 NHWC tensor
// [B][H][W][Kp], column k = ci*9 + kh*3 + kw (the OIHW inner order, so the weight matrix is w.reshape(Cout, Cin*9)),
// columns >= 9*Cin zero.  The 1x1 implicit-GEMM kernel then does the contraction (K padded to one 64-deep block by
// TMA zero fill) and writes the first big activation tensor at store bandwidth; the CUDA-core stem above spends 19
// GFLOP of fp32 FMAs on it (0.97 ms at B=16, 512x512 against a 0.17 ms write floor).  One thread per pixel: its 9*Cin
// loads are coalesced along the row and hit L1 for the 3x3 overlap.
template <int KP>
__global__ void __launch_bounds__(256) stem_im2col_kernel(const float* __restrict__ x0, int C0,
                                                         const float* __restrict__ x1, int C1, float in_scale,
                                                         float in_shift, uint4* __restrict__ out, int H, int W,
                                                         int64_t total) {
  const int64_t pix = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (pix >= total) return;
  const int HW = H * W;
  const int n = (int)(pix / HW);
  const int rem = (int)(pix - (int64_t)n * HW);
  const int h = rem / W, w = rem - h * W;
  const int Cin = C0 + C1;
  float v[KP];
#pragma unroll
  for (int k = 0; k < KP; ++k) v[k] = 0.f;
#pragma unroll
  for (int ci = 0; ci < KP / 9; ++ci) {
    if (ci < Cin) {
      const float* src = (ci < C0) ? x0 + ((size_t)n * C0 + ci) * HW : x1 + ((size_t)n * C1 + (ci - C0)) * HW;
#pragma unroll
      for (int kh = 0; kh < 3; ++kh) {
        const int ih = h + kh - 1;
#pragma unroll
        for (int kw = 0; kw < 3; ++kw) {
          const int iw = w + kw - 1;
          // zero padding applies AFTER the optional 2x-1 centering
          if ((unsigned)ih < (unsigned)H && (unsigned)iw < (unsigned)W)
            v[ci * 9 + kh * 3 + kw] = fmaf(__ldg(src + ih * W + iw), in_scale, in_shift);
        }
      }
    }
  }
  uint4* dst = out + pix * (KP / 8);
#pragma unroll
  for (int c = 0; c < KP / 8; ++c) {
    uint4 o;
    o.x = pack_bf16x2(v[c * 8 + 0], v[c * 8 + 1]); o.y = pack_bf16x2(v[c * 8 + 2], v[c * 8 + 3]);
    o.z = pack_bf16x2(v[c * 8 + 4], v[c * 8 + 5]); o.w = pack_bf16x2(v[c * 8 + 6], v[c * 8 + 7]);
    dst[c] = o;
  }
}

// Head: a block computes a 32 x 4 tile of output pixels from a (34 x 6)-pixel halo tile staged in shared memory
// (row pitch Cin*2+16 bytes => conflict-free 16-byte reads with one thread per pixel); weights fp32 in smem as
// [tap][ci][co].  Global reads are fully coalesced (NHWC rows are contiguous), every input byte is read once per tile.
constexpr int kHeadTW = 32, kHeadTH = 4, kHeadThreads = kHeadTW * kHeadTH;

template <int COUT>
__global__ void __launch_bounds__(kHeadThreads) conv_head_kernel(const uint4* __restrict__ x,
                                                                const float* __restrict__ w_oihw,
                                                                const float* __restrict__ bias,
                                                                float* __restrict__ out, int B, int H, int W,
                                                                int Cin, const float* __restrict__ norm_ab,
                                                                int norm_act) {
  extern __shared__ __align__(16) uint8_t smem[];
  const int c8n = Cin >> 3;
  const int pitch = Cin * 2 + 16;                                   // bytes per staged pixel
  float* sw = reinterpret_cast<float*>(smem);                       // [9][Cin][COUT]
  float* sab = reinterpret_cast<float*>(smem + ((9 * Cin * COUT * 4 + 15) & ~15));  // [2][Cin] a, b (halved for SiLU)
  uint8_t* st = reinterpret_cast<uint8_t*>(sab + 2 * Cin);          // [(TH+2)*(TW+2)][pitch]
  for (int i = threadIdx.x; i < 9 * Cin * COUT; i += kHeadThreads) {
    const int co = i % COUT;
    const int ci = (i / COUT) % Cin;
    const int tap = i / (COUT * Cin);
    sw[i] = w_oihw[((size_t)co * Cin + ci) * 9 + tap];
  }
  const int w0 = blockIdx.x * kHeadTW, h0 = blockIdx.y * kHeadTH, n = blockIdx.z;
  constexpr int HP = kHeadTH + 2, WP = kHeadTW + 2;
  if (norm_ab != nullptr) {
    const float k = norm_act ? 0.5f : 1.0f;
    for (int i = threadIdx.x; i < 2 * Cin; i += kHeadThreads) sab[i] = k * norm_ab[(size_t)n * 2 * Cin + i];
    __syncthreads();
  }
  for (int i = threadIdx.x; i < HP * WP * c8n; i += kHeadThreads) {
    const int c = i % c8n;
    const int p = i / c8n;
    const int pw = p % WP, ph = p / WP;
    const int ih = h0 + ph - 1, iw = w0 + pw - 1;
    uint4 v = make_uint4(0, 0, 0, 0);
    if ((unsigned)ih < (unsigned)H && (unsigned)iw < (unsigned)W) {
      v = __ldg(x + (((size_t)n * H + ih) * W + iw) * c8n + c);
      if (norm_ab != nullptr) {  // conv_norm_out + SiLU folded into the load; padding stays zero
        const float4 a0 = *reinterpret_cast<const float4*>(sab + c * 8);
        const float4 a1 = *reinterpret_cast<const float4*>(sab + c * 8 + 4);
        const float4 b0 = *reinterpret_cast<const float4*>(sab + Cin + c * 8);
        const float4 b1 = *reinterpret_cast<const float4*>(sab + Cin + c * 8 + 4);
        if (norm_act) {
          v.x = xf_word<true>(v.x, a0.x, b0.x, a0.y, b0.y); v.y = xf_word<true>(v.y, a0.z, b0.z, a0.w, b0.w);
          v.z = xf_word<true>(v.z, a1.x, b1.x, a1.y, b1.y); v.w = xf_word<true>(v.w, a1.z, b1.z, a1.w, b1.w);
        } else {
          v.x = xf_word<false>(v.x, a0.x, b0.x, a0.y, b0.y); v.y = xf_word<false>(v.y, a0.z, b0.z, a0.w, b0.w);
          v.z = xf_word<false>(v.z, a1.x, b1.x, a1.y, b1.y); v.w = xf_word<false>(v.w, a1.z, b1.z, a1.w, b1.w);
        }
      }
    }
    *reinterpret_cast<uint4*>(st + p * pitch + c * 16) = v;
  }
  __syncthreads();
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int ow = w0 + tx, oh = h0 + ty;
  float acc[COUT];
#pragma unroll
  for (int c = 0; c < COUT; ++c) acc[c] = bias ? bias[c] : 0.f;
#pragma unroll
  for (int kh = 0; kh < 3; ++kh) {
#pragma unroll
    for (int kw = 0; kw < 3; ++kw) {
      const uint8_t* px = st + ((ty + kh) * WP + tx + kw) * pitch;
      const float* wt = sw + (kh * 3 + kw) * Cin * COUT;
#pragma unroll 2
      for (int c8 = 0; c8 < c8n; ++c8) {
        const uint4 u = *reinterpret_cast<const uint4*>(px + c8 * 16);
        const float2 f0 = unpack_bf16x2(u.x), f1 = unpack_bf16x2(u.y), f2 = unpack_bf16x2(u.z),
                     f3 = unpack_bf16x2(u.w);
        const float v[8] = {f0.x, f0.y, f1.x, f1.y, f2.x, f2.y, f3.x, f3.y};
        if (COUT == 1) {
          const float4 wa = *reinterpret_cast<const float4*>(wt + c8 * 8);
          const float4 wb = *reinterpret_cast<const float4*>(wt + c8 * 8 + 4);
          // four independent partial sums keep the FMA pipe busy (the chain is 1152 FMAs long otherwise)
          const float p0 = fmaf(v[1], wa.y, v[0] * wa.x), p1 = fmaf(v[3], wa.w, v[2] * wa.z);
          const float p2 = fmaf(v[5], wb.y, v[4] * wb.x), p3 = fmaf(v[7], wb.w, v[6] * wb.z);
          acc[0] += (p0 + p1) + (p2 + p3);
        } else {
#pragma unroll
          for (int j = 0; j < 8; ++j)
#pragma unroll
            for (int c = 0; c < COUT; ++c) acc[c] = fmaf(v[j], wt[(c8 * 8 + j) * COUT + c], acc[c]);
        }
      }
    }
  }
  if (ow < W && oh < H) {
#pragma unroll
    for (int c = 0; c < COUT; ++c) out[(((size_t)n * COUT + c) * H + oh) * W + ow] = acc[c];
  }
}


// Head, Cout = 1, Cin in {64, 128}: "dot-then-gather".  Phase 1: every input pixel of the (16+2) x (64+2) halo tile is
// read ONCE, straight from global memory into registers (LPP lanes per pixel, 16 bytes each, coalesced), passed through
// the fused output norm (act(a*x+b)) and dotted with the nine per-tap weight vectors, which live in registers for the
// whole kernel (72 floats per lane); the nine partial sums are reduced across the LPP lanes with a recursive-halving
// butterfly and parked in shared memory (36 bytes per pixel instead of 2*Cin).  Phase 2: out(h, w) = bias +
// sum over taps of P[h+kh][w+kw][tap], nine conflict-free shared loads per output pixel.  Against the tile-staging
// kernel above this removes the 9x re-read of the activations from shared memory.
constexpr int kHd2TH = 16, kHd2TW = 64, kHd2Threads = 256;
constexpr int kHd2Pix = (kHd2TH + 2) * (kHd2TW + 2);

template <int LPP>
__global__ void __launch_bounds__(kHd2Threads, 2) conv_head_dot_kernel(const uint4* __restrict__ x,
                                                                      const float* __restrict__ w_oihw,
                                                                      const float* __restrict__ bias,
                                                                      float* __restrict__ out, int H, int W,
                                                                      const float* __restrict__ norm_ab,
                                                                      int norm_act) {
  constexpr int C = LPP * 8;
  constexpr int PPW = 32 / LPP;  // pixels per warp iteration
  __shared__ float P[kHd2Pix * 9];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int c8 = lane % LPP, sub = lane / LPP;
  const int w0 = blockIdx.x * kHd2TW, h0 = blockIdx.y * kHd2TH, n = blockIdx.z;

  // taps paired for packed FMAs: wr[tp][j] = (w[tap 2tp][j], w[tap 2tp+1][j]); the tenth slot is zero
  float2 wr[5][8];
#pragma unroll
  for (int tp = 0; tp < 5; ++tp)
#pragma unroll
    for (int j = 0; j < 8; ++j)
      wr[tp][j] = make_float2(__ldg(w_oihw + (c8 * 8 + j) * 9 + 2 * tp),
                              (2 * tp + 1 < 9) ? __ldg(w_oihw + (c8 * 8 + j) * 9 + 2 * tp + 1) : 0.f);
  float a[8], b[8];
  const bool has_norm = norm_ab != nullptr;
  const bool silu = has_norm && norm_act != 0;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float k = silu ? 0.5f : 1.0f;
    a[j] = has_norm ? k * __ldg(norm_ab + (size_t)n * 2 * C + c8 * 8 + j) : 1.0f;
    b[j] = has_norm ? k * __ldg(norm_ab + (size_t)n * 2 * C + C + c8 * 8 + j) : 0.0f;
  }

  constexpr int kWP = kHd2TW + 2;
  constexpr int kPixPerIter = (kHd2Threads / 32) * PPW;
  constexpr int kIters = (kHd2Pix + kPixPerIter - 1) / kPixPerIter;
  constexpr int kUnroll = 4;  // independent 16-byte loads in flight per lane
  for (int it0 = 0; it0 < kIters; it0 += kUnroll) {
    uint4 v[kUnroll];
    bool inb[kUnroll];
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      const int pi = (it0 + u) * kPixPerIter + warp * PPW + sub;
      const int ph = pi / kWP, pw = pi - ph * kWP;
      const int ih = h0 + ph - 1, iw = w0 + pw - 1;
      inb[u] = pi < kHd2Pix && (unsigned)ih < (unsigned)H && (unsigned)iw < (unsigned)W;
      v[u] = make_uint4(0, 0, 0, 0);
      if (inb[u]) v[u] = __ldg(x + (((size_t)n * H + ih) * W + iw) * LPP + c8);
    }
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      const int pi = (it0 + u) * kPixPerIter + warp * PPW + sub;
      if ((it0 + u) * kPixPerIter + warp * PPW >= kHd2Pix) break;  // warp-uniform
      const uint32_t wd[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
      float xv[8];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float h0v = fmaf(__uint_as_float(wd[j] << 16), a[2 * j], b[2 * j]);
        float h1v = fmaf(__uint_as_float(wd[j] & 0xffff0000u), a[2 * j + 1], b[2 * j + 1]);
        if (silu) {
          float t0, t1;
          asm("tanh.approx.f32 %0, %1;" : "=f"(t0) : "f"(h0v));
          asm("tanh.approx.f32 %0, %1;" : "=f"(t1) : "f"(h1v));
          h0v = fmaf(h0v, t0, h0v);
          h1v = fmaf(h1v, t1, h1v);
        }
        // zero padding applies AFTER the activation: an out-of-image pixel contributes nothing
        xv[2 * j] = inb[u] ? h0v : 0.f;
        xv[2 * j + 1] = inb[u] ? h1v : 0.f;
      }
      float red[10];
#pragma unroll
      for (int tp = 0; tp < 5; ++tp) {
        float2 acc = make_float2(0.f, 0.f);
#pragma unroll
        for (int j = 0; j < 8; ++j) acc = ffma2(make_float2(xv[j], xv[j]), wr[tp][j], acc);
        red[2 * tp] = acc.x;
        red[2 * tp + 1] = acc.y;
      }
      // recursive halving over the top three lane bits of the pixel's lane group: 8 values -> 1 per lane
#pragma unroll
      for (int width = 4, mask = LPP / 2; width >= 1; width >>= 1, mask >>= 1) {
        const bool upper = (lane & mask) != 0;
#pragma unroll
        for (int i = 0; i < width; ++i) {
          const float keep = upper ? red[i + width] : red[i];
          const float give = upper ? red[i] : red[i + width];
          red[i] = keep + __shfl_xor_sync(0xffffffffu, give, mask);
        }
      }
#pragma unroll
      for (int mask = LPP / 16; mask >= 1; mask >>= 1) red[0] += __shfl_xor_sync(0xffffffffu, red[0], mask);
#pragma unroll
      for (int mask = LPP / 2; mask >= 1; mask >>= 1) red[8] += __shfl_xor_sync(0xffffffffu, red[8], mask);
      if (pi < kHd2Pix) {
        const int hi = c8 / (LPP / 8);  // the three halving bits: tap index 4*b2 + 2*b1 + b0
        const int tap = ((hi >> 2) & 1) * 4 + ((hi >> 1) & 1) * 2 + (hi & 1);
        if (c8 % (LPP / 8) == 0) P[pi * 9 + tap] = red[0];
        if (c8 == 0) P[pi * 9 + 8] = red[8];
      }
    }
  }
  __syncthreads();
  const float bv = bias ? __ldg(bias) : 0.f;
  for (int o = threadIdx.x; o < kHd2TH * kHd2TW; o += kHd2Threads) {
    const int ty = o / kHd2TW, tx = o - ty * kHd2TW;
    const int oh = h0 + ty, ow = w0 + tx;
    if (oh >= H || ow >= W) continue;
    float acc = bv;
#pragma unroll
    for (int kh = 0; kh < 3; ++kh)
#pragma unroll
      for (int kw = 0; kw < 3; ++kw) acc += P[((ty + kh) * kWP + tx + kw) * 9 + kh * 3 + kw];
    out[((size_t)n * H + oh) * W + ow] = acc;
  }
}

}  // namespace fm

using namespace fm;

static int stem_blocks_per_image(int B, int H, int W, int Cout) {
  const int gpb = kStemThreads / (Cout / 8);
  const int64_t groups = (int64_t)H * ((W + kStemPix - 1) / kStemPix);
  int64_t bpi = (groups + gpb - 1) / gpb;
  int64_t cap = ((int64_t)sm_count() * 8 + B - 1) / B;  // ~8 resident-CTA rounds' worth across the batch
  if (cap < 1) cap = 1;
  if (bpi > cap) bpi = cap;
  return (int)bpi;
}

extern "C" int fm_conv_stem_stats_rows(int32_t B, int32_t H, int32_t W, int32_t Cout) {
  if (B <= 0 || H <= 0 || W <= 0 || Cout <= 0 || Cout % 8 || Cout > 512) return 0;
  return stem_blocks_per_image(B, H, W, Cout);
}

extern "C" int fm_conv_stem_f32_bf16(const float* x0, int32_t C0, const float* x1, int32_t C1, float in_scale,
                                     float in_shift, const float* weight_oihw, const float* bias, void* out, int32_t B,
                                     int32_t H, int32_t W, int32_t Cout, float* gn_stats, fm_stream_t stream) {
  if (int e = ensure_device()) return e;
  FM_REQUIRE(x0 && C0 > 0 && (x1 != nullptr) == (C1 > 0), "conv_stem: inconsistent sources");
  FM_REQUIRE(C0 + C1 <= kStemMaxCin, "conv_stem: Cin=%d exceeds %d", C0 + C1, kStemMaxCin);
  FM_REQUIRE(Cout > 0 && Cout % 8 == 0 && Cout <= 512, "conv_stem: Cout=%d must be a multiple of 8 (<=512)", Cout);
  FM_REQUIRE(weight_oihw && out && B > 0 && H > 0 && W > 0, "conv_stem: bad argument");
  FM_REQUIRE((int64_t)H * W < (1ll << 31) && B <= 65535, "conv_stem: image too large for 32-bit indexing");
  const int gpb = kStemThreads / (Cout / 8);
  size_t smem = ((size_t)(C0 + C1) * 9 * Cout + Cout) * sizeof(float);
  const size_t red_bytes = (size_t)gpb * (Cout / 4) * 2 * sizeof(float);
  if (gn_stats != nullptr && red_bytes > smem) smem = red_bytes;
  static size_t attr = 0;
  if (smem > 48 * 1024 && smem > attr) {
    cudaError_t e = cudaFuncSetAttribute(conv_stem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return check_cuda(e, "cudaFuncSetAttribute(conv_stem)");
    attr = smem;
  }
  const int bpi = stem_blocks_per_image(B, H, W, Cout);
  conv_stem_kernel<<<dim3(bpi, B), kStemThreads, smem, (cudaStream_t)stream>>>(
      x0, C0, x1, C1, in_scale, in_shift, weight_oihw, bias, reinterpret_cast<uint4*>(out), H, W, Cout, gn_stats);
  FM_LAUNCH_CHECK("conv_stem_kernel");
  return 0;
}

extern "C" int fm_stem_im2col_bf16(const float* x0, int32_t C0, const float* x1, int32_t C1, float in_scale,
                                   float in_shift, void* out, int32_t B, int32_t H, int32_t W, int32_t Kp,
                                   fm_stream_t stream) {
  if (int e = ensure_device()) return e;
  FM_REQUIRE(x0 && C0 > 0 && (x1 != nullptr) == (C1 > 0), "stem_im2col: inconsistent sources");
  FM_REQUIRE(out && B > 0 && H > 0 && W > 0, "stem_im2col: bad argument");
  FM_REQUIRE(Kp % 8 == 0 && Kp >= 9 * (C0 + C1) && Kp <= 72, "stem_im2col: Kp=%d must be a multiple of 8 covering 9*Cin "
             "(<= 72)", Kp);
  FM_REQUIRE((int64_t)H * W < (1ll << 31), "stem_im2col: image too large for 32-bit indexing");
  const int64_t total = (int64_t)B * H * W;
  const unsigned blocks = (unsigned)((total + 255) / 256);
  cudaStream_t st = (cudaStream_t)stream;
  uint4* o = reinterpret_cast<uint4*>(out);
#define FM_IM2COL_CASE(K)                                                                                            \
  case K: stem_im2col_kernel<K><<<blocks, 256, 0, st>>>(x0, C0, x1, C1, in_scale, in_shift, o, H, W, total); break;
  switch (Kp) {
    FM_IM2COL_CASE(16) FM_IM2COL_CASE(24) FM_IM2COL_CASE(32) FM_IM2COL_CASE(40) FM_IM2COL_CASE(48) FM_IM2COL_CASE(56)
    FM_IM2COL_CASE(64) FM_IM2COL_CASE(72)
    default: set_error("stem_im2col: unsupported Kp=%d", Kp); return FM_ERR_UNSUPPORTED;
  }
#undef FM_IM2COL_CASE
  FM_LAUNCH_CHECK("stem_im2col_kernel");
  return 0;
}

extern "C" int fm_conv_head_bf16_f32(const void* x, const float* weight_oihw, const float* bias, float* out, int32_t B,
                                     int32_t H, int32_t W, int32_t Cin, int32_t Cout, const float* norm_ab,
                                     int32_t norm_act, fm_stream_t stream) {
  if (int e = ensure_device()) return e;
  FM_REQUIRE(x && weight_oihw && out && B > 0 && H > 0 && W > 0, "conv_head: bad argument");
  FM_REQUIRE(Cin > 0 && Cin % 8 == 0, "conv_head: Cin=%d must be a multiple of 8", Cin);
  FM_REQUIRE(Cout >= 1 && Cout <= 4, "conv_head: Cout=%d must be in 1..4", Cout);
  FM_REQUIRE(B <= 65535, "conv_head: batch too large for the grid");
  if (Cout == 1 && (Cin == 64 || Cin == 128) && getenv("FMDM_HEAD_TILE") == nullptr) {
    dim3 grid2((W + kHd2TW - 1) / kHd2TW, (H + kHd2TH - 1) / kHd2TH, B);
    if (Cin == 64)
      conv_head_dot_kernel<8><<<grid2, kHd2Threads, 0, (cudaStream_t)stream>>>(
          reinterpret_cast<const uint4*>(x), weight_oihw, bias, out, H, W, norm_ab, norm_act);
    else
      conv_head_dot_kernel<16><<<grid2, kHd2Threads, 0, (cudaStream_t)stream>>>(
          reinterpret_cast<const uint4*>(x), weight_oihw, bias, out, H, W, norm_ab, norm_act);
    FM_LAUNCH_CHECK("conv_head_dot_kernel");
    return 0;
  }
  const size_t wbytes = ((size_t)9 * Cin * Cout * sizeof(float) + 15) & ~(size_t)15;
  const size_t smem =
      wbytes + (size_t)2 * Cin * sizeof(float) + (size_t)(kHeadTH + 2) * (kHeadTW + 2) * (Cin * 2 + 16);
  FM_REQUIRE(norm_ab == nullptr || ((uintptr_t)norm_ab & 15) == 0, "conv_head: norm_ab must be 16B aligned");
  FM_REQUIRE(smem <= 200 * 1024, "conv_head: Cin=%d does not fit shared memory", Cin);
  dim3 grid((W + kHeadTW - 1) / kHeadTW, (H + kHeadTH - 1) / kHeadTH, B);
  cudaStream_t st = (cudaStream_t)stream;
  const uint4* xp = reinterpret_cast<const uint4*>(x);
#define FM_HEAD_CASE(N)                                                                                              \
  case N: {                                                                                                          \
    if (smem > 48 * 1024) {                                                                                          \
      cudaError_t e = cudaFuncSetAttribute(conv_head_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize,         \
                                           (int)smem);                                                               \
      if (e != cudaSuccess) return check_cuda(e, "cudaFuncSetAttribute(conv_head)");                                 \
    }                                                                                                                \
    conv_head_kernel<N><<<grid, kHeadThreads, smem, st>>>(xp, weight_oihw, bias, out, B, H, W, Cin, norm_ab,        \
                                                          norm_act);                                               \
    break;                                                                                                           \
  }
  switch (Cout) {
    FM_HEAD_CASE(1)
    FM_HEAD_CASE(2)
    FM_HEAD_CASE(3)
    FM_HEAD_CASE(4)
  }
#undef FM_HEAD_CASE
  FM_LAUNCH_CHECK("conv_head_kernel");
  return 0;
}
