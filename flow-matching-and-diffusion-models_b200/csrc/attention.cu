// K3 (first version): softmax(Q K^T * scale) V with an online (flash-style) softmax, bf16 in / bf16 out, fp32 math.
// One thread owns one query row; the K/V rows of the (batch, head) are staged through shared memory in chunks and
// read by all threads as broadcasts.  With the reference's head_dim = 8 (64 heads at C = 512,
// src/nn/blocks/legacy_unet.py:32) the op is exp-bound, not MMA-bound, so this CUDA-core
// formulation is the baseline; a tensor-core variant is future work.
// Replaces F.scaled_dot_product_attention (src/nn/blocks/attention.py:41-44).
#include <cstdlib>

#include "common.cuh"

namespace fm {

constexpr int kAttThreads = 128;

template <int HD>
__global__ void __launch_bounds__(kAttThreads) attention_kernel(const __nv_bfloat16* __restrict__ q,
                                                               const __nv_bfloat16* __restrict__ k,
                                                               const __nv_bfloat16* __restrict__ v,
                                                               __nv_bfloat16* __restrict__ out, int Tq, int Tk,
                                                               int64_t q_sb, int64_t q_sh, int64_t q_st, int64_t kv_sb,
                                                               int64_t kv_sh, int64_t kv_st, int64_t o_sb, int64_t o_sh,
                                                               int64_t o_st, float scale_log2e, int kchunk) {
  extern __shared__ uint4 smem_u4[];
  constexpr int V8 = HD / 8;                  // 16-byte vectors per row
  uint4* sk = smem_u4;                        // [kchunk][V8]
  uint4* sv = smem_u4 + (size_t)kchunk * V8;  // [kchunk][V8]
  const int b = blockIdx.z, h = blockIdx.y;
  const int t = blockIdx.x * kAttThreads + threadIdx.x;
  const bool active = t < Tq;

  float qr[HD], acc[HD];
#pragma unroll
  for (int d = 0; d < HD; ++d) { qr[d] = 0.f; acc[d] = 0.f; }
  if (active) {
    const uint4* qp = reinterpret_cast<const uint4*>(q + b * q_sb + h * q_sh + (int64_t)t * q_st);
#pragma unroll
    for (int i = 0; i < V8; ++i) {
      const uint4 u = qp[i];
      const float2 f0 = unpack_bf16x2(u.x), f1 = unpack_bf16x2(u.y), f2 = unpack_bf16x2(u.z), f3 = unpack_bf16x2(u.w);
      qr[i * 8 + 0] = f0.x * scale_log2e; qr[i * 8 + 1] = f0.y * scale_log2e;
      qr[i * 8 + 2] = f1.x * scale_log2e; qr[i * 8 + 3] = f1.y * scale_log2e;
      qr[i * 8 + 4] = f2.x * scale_log2e; qr[i * 8 + 5] = f2.y * scale_log2e;
      qr[i * 8 + 6] = f3.x * scale_log2e; qr[i * 8 + 7] = f3.y * scale_log2e;
    }
  }
  float m = -INFINITY, l = 0.f;
  const __nv_bfloat16* kb = k + b * kv_sb + h * kv_sh;
  const __nv_bfloat16* vb = v + b * kv_sb + h * kv_sh;

  for (int j0 = 0; j0 < Tk; j0 += kchunk) {
    const int nj = min(kchunk, Tk - j0);
    __syncthreads();
    for (int i = threadIdx.x; i < nj * V8; i += kAttThreads) {
      const int j = i / V8, c = i - j * V8;
      sk[i] = reinterpret_cast<const uint4*>(kb + (int64_t)(j0 + j) * kv_st)[c];
      sv[i] = reinterpret_cast<const uint4*>(vb + (int64_t)(j0 + j) * kv_st)[c];
    }
    __syncthreads();
    for (int j = 0; j < nj; j += 4) {
      float s[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        float sacc = 0.f;
        if (j + u < nj) {
#pragma unroll
          for (int i = 0; i < V8; ++i) {
            const uint4 kk = sk[(j + u) * V8 + i];
            const float2 f0 = unpack_bf16x2(kk.x), f1 = unpack_bf16x2(kk.y), f2 = unpack_bf16x2(kk.z),
                         f3 = unpack_bf16x2(kk.w);
            sacc = fmaf(qr[i * 8 + 0], f0.x, sacc); sacc = fmaf(qr[i * 8 + 1], f0.y, sacc);
            sacc = fmaf(qr[i * 8 + 2], f1.x, sacc); sacc = fmaf(qr[i * 8 + 3], f1.y, sacc);
            sacc = fmaf(qr[i * 8 + 4], f2.x, sacc); sacc = fmaf(qr[i * 8 + 5], f2.y, sacc);
            sacc = fmaf(qr[i * 8 + 6], f3.x, sacc); sacc = fmaf(qr[i * 8 + 7], f3.y, sacc);
          }
        } else {
          sacc = -INFINITY;
        }
        s[u] = sacc;
      }
      const float mx = fmaxf(fmaxf(fmaxf(s[0], s[1]), fmaxf(s[2], s[3])), m);
      const float corr = exp2f(m - mx);
      float pr[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) pr[u] = exp2f(s[u] - mx);
      l = l * corr + (pr[0] + pr[1]) + (pr[2] + pr[3]);
      m = mx;
#pragma unroll
      for (int d = 0; d < HD; ++d) acc[d] *= corr;
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        if (j + u < nj) {
#pragma unroll
          for (int i = 0; i < V8; ++i) {
            const uint4 vv = sv[(j + u) * V8 + i];
            const float2 f0 = unpack_bf16x2(vv.x), f1 = unpack_bf16x2(vv.y), f2 = unpack_bf16x2(vv.z),
                         f3 = unpack_bf16x2(vv.w);
            acc[i * 8 + 0] = fmaf(pr[u], f0.x, acc[i * 8 + 0]); acc[i * 8 + 1] = fmaf(pr[u], f0.y, acc[i * 8 + 1]);
            acc[i * 8 + 2] = fmaf(pr[u], f1.x, acc[i * 8 + 2]); acc[i * 8 + 3] = fmaf(pr[u], f1.y, acc[i * 8 + 3]);
            acc[i * 8 + 4] = fmaf(pr[u], f2.x, acc[i * 8 + 4]); acc[i * 8 + 5] = fmaf(pr[u], f2.y, acc[i * 8 + 5]);
            acc[i * 8 + 6] = fmaf(pr[u], f3.x, acc[i * 8 + 6]); acc[i * 8 + 7] = fmaf(pr[u], f3.y, acc[i * 8 + 7]);
          }
        }
      }
    }
  }
  if (active) {
    const float inv = 1.0f / l;
    uint4* op = reinterpret_cast<uint4*>(out + b * o_sb + h * o_sh + (int64_t)t * o_st);
#pragma unroll
    for (int i = 0; i < V8; ++i) {
      uint4 o;
      o.x = pack_bf16x2(acc[i * 8 + 0] * inv, acc[i * 8 + 1] * inv);
      o.y = pack_bf16x2(acc[i * 8 + 2] * inv, acc[i * 8 + 3] * inv);
      o.z = pack_bf16x2(acc[i * 8 + 4] * inv, acc[i * 8 + 5] * inv);
      o.w = pack_bf16x2(acc[i * 8 + 6] * inv, acc[i * 8 + 7] * inv);
      op[i] = o;
    }
  }
}


// ---------------------------------------------------------------------------------------------------------------
// head_dim == 8 (the reference's default: 64 heads at C = 512): tensor-core version on legacy mma.sync m16n8k8
// (bf16 in, fp32 accumulate).  tcgen05 is the wrong tool here: K = 8 is half of one UMMA K step and the op is
// exp-bound (64*T^2 exponentials per sample per layer), so the win is in removing the 16 FMAs per score from the
// CUDA cores, not in MMA peak.  One warp owns 16 query rows; S = Q K^T comes from one m16n8k8 per 8 keys, the
// accumulator fragment is re-used in place as the A operand of P V (same lane layout), V is staged transposed.
// ---------------------------------------------------------------------------------------------------------------
constexpr int kAtt8Warps = 8;
constexpr int kAtt8Threads = kAtt8Warps * 32;
constexpr int kAtt8KeyChunk = 1024;            // keys staged per pass
constexpr int kAtt8VtStride = kAtt8KeyChunk + 8;  // +8 bf16 (16 B) padding: conflict-free transposed reads

__device__ __forceinline__ void mma_m16n8k16_f16(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3,
                                                 uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

// a pair of probabilities as packed halves: 2^x on the special-function unit in fp32 (`ex2.approx.f16x2` is issued as
// two MUFU.EX2.F16 plus byte permutes on sm_100a - measured in SASS - so the packed form saves nothing), one pack
__device__ __forceinline__ uint32_t ex2_f16x2(float lo, float hi) {
  uint32_t y;
  asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(y) : "f"(ex2_approx(hi)), "f"(ex2_approx(lo)));
  return y;
}

// 2^x for a pair of scores WITHOUT the special-function unit (the FlashAttention-4 trick): x <= 0 is clamped at -30
// (2^-30 is zero in fp16; also maps the -inf of masked keys to zero), split as n + f with the 1.5*2^23 magic-number
// rounding (n integer, |f| <= 0.5), 2^f by a minimax cubic (relative error 7.5e-5, fp16 keeps 4.9e-4) on packed fp32
// FMAs, and 2^n applied by adding n to the exponent field.  The hd-8 loop does one MUFU per score and the MUFU pipe
// (16 / clk / SM) is its ceiling; moving a share of the exponentials to the FMA pipe trades MUFU cycles for issue slots.
__device__ __forceinline__ uint32_t ex2_poly_f16x2(float lo, float hi) {
  const float2 x = make_float2(fmaxf(lo, -30.f), fmaxf(hi, -30.f));
  const float2 t = fadd2(x, make_float2(12582912.f, 12582912.f));
  const float2 n = fadd2(t, make_float2(-12582912.f, -12582912.f));
  const float2 f = ffma2(n, make_float2(-1.f, -1.f), x);
  float2 p = ffma2(f, make_float2(0.0551716685f, 0.0551716685f), make_float2(0.2426111251f, 0.2426111251f));
  p = ffma2(p, f, make_float2(0.6932609677f, 0.6932609677f));
  p = ffma2(p, f, make_float2(0.9999280572f, 0.9999280572f));
  const float rl = __uint_as_float(__float_as_uint(p.x) + (__float_as_uint(t.x) << 23));
  const float rh = __uint_as_float(__float_as_uint(p.y) + (__float_as_uint(t.y) << 23));
  uint32_t y;
  asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(y) : "f"(rh), "f"(rl));
  return y;
}

// F16P = true: probabilities as packed halves (11 significant bits instead of bf16's 8).  Per 16 keys and thread: no
// row-sum adds (the row sums come from one extra MMA against a ones matrix, which also folds the four lanes of a
// row) and P V as ONE m16n8k16 (V staged as fp16).  The loop is issue-bound, so the instruction count is what matters.
// POLY (F16P only): how many of the four probability pairs per 16 keys take the FMA-pipe exponential (0, 1 or 2).
constexpr int kAtt8PolyDefault = 0;
template <bool F16P, int POLY>
__global__ void __launch_bounds__(kAtt8Threads) attention_hd8_mma_kernel(
    const __nv_bfloat16* __restrict__ q, const __nv_bfloat16* __restrict__ k, const __nv_bfloat16* __restrict__ v,
    __nv_bfloat16* __restrict__ out, int Tq, int Tk, int64_t q_sb, int64_t q_sh, int64_t q_st, int64_t kv_sb,
    int64_t kv_sh, int64_t kv_st, int64_t o_sb, int64_t o_sh, int64_t o_st, float scale_log2e) {
  pdl_enter();
  __shared__ __align__(16) __nv_bfloat16 sK[kAtt8KeyChunk * 8];      // [key][8]
  __shared__ __align__(16) __nv_bfloat16 sVt[8 * kAtt8VtStride];     // [dim][key]
  const int b = blockIdx.z, h = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  const int q0 = blockIdx.x * (kAtt8Warps * 16) + warp * 16;
  const __nv_bfloat16* qb = q + b * q_sb + h * q_sh;
  const __nv_bfloat16* kb = k + b * kv_sb + h * kv_sh;
  const __nv_bfloat16* vb = v + b * kv_sb + h * kv_sh;

  // Q fragment: a0 = Q[q0+g][2t,2t+1], a1 = Q[q0+g+8][2t,2t+1]  (rows clamped; stores are masked)
  const int r0 = min(q0 + g, Tq - 1), r1 = min(q0 + g + 8, Tq - 1);
  const uint32_t qa0 = *reinterpret_cast<const uint32_t*>(qb + (int64_t)r0 * q_st + 2 * t);
  const uint32_t qa1 = *reinterpret_cast<const uint32_t*>(qb + (int64_t)r1 * q_st + 2 * t);

  float o[4] = {0.f, 0.f, 0.f, 0.f};
  float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;   // running max (log2 domain) / partial sums per row

  for (int c0 = 0; c0 < Tk; c0 += kAtt8KeyChunk) {
    const int nk = min(kAtt8KeyChunk, Tk - c0);
    const int nk_pad = (nk + 63) & ~63;
    __syncthreads();
    for (int j = threadIdx.x; j < nk_pad; j += kAtt8Threads) {
      uint4 kk = make_uint4(0, 0, 0, 0), vv = make_uint4(0, 0, 0, 0);
      if (j < nk) {
        kk = *reinterpret_cast<const uint4*>(kb + (int64_t)(c0 + j) * kv_st);
        vv = *reinterpret_cast<const uint4*>(vb + (int64_t)(c0 + j) * kv_st);
      }
      *reinterpret_cast<uint4*>(sK + j * 8) = kk;
      const __nv_bfloat16* ve = reinterpret_cast<const __nv_bfloat16*>(&vv);
      if constexpr (F16P) {
        __half* sVh = reinterpret_cast<__half*>(sVt);
#pragma unroll
        for (int d = 0; d < 8; ++d) sVh[d * kAtt8VtStride + j] = __float2half_rn(__bfloat162float(ve[d]));
      } else {
#pragma unroll
        for (int d = 0; d < 8; ++d) sVt[d * kAtt8VtStride + j] = ve[d];
      }
    }
    __syncthreads();

    for (int kb0 = 0; kb0 < nk_pad; kb0 += 64) {
      float s[8][4];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        s[j][0] = s[j][1] = s[j][2] = s[j][3] = 0.f;
        const uint32_t bk = *reinterpret_cast<const uint32_t*>(sK + (kb0 + j * 8 + g) * 8 + 2 * t);
        mma_m16n8k8_bf16(s[j], qa0, qa1, bk);
      }
      if (kb0 + 64 > nk) {  // mask the padded keys of the last block
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int key = kb0 + j * 8 + 2 * t;
          if (key >= nk) { s[j][0] = -INFINITY; s[j][2] = -INFINITY; }
          if (key + 1 >= nk) { s[j][1] = -INFINITY; s[j][3] = -INFINITY; }
        }
      }
      float mx0 = s[0][0], mx1 = s[0][2];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        mx0 = fmaxf(mx0, fmaxf(s[j][0], s[j][1]));
        mx1 = fmaxf(mx1, fmaxf(s[j][2], s[j][3]));
      }
      mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1));
      mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
      mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
      mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
      const float mn0 = fmaxf(m0, mx0 * scale_log2e), mn1 = fmaxf(m1, mx1 * scale_log2e);
      const float corr0 = ex2_approx(m0 - mn0), corr1 = ex2_approx(m1 - mn1);
      m0 = mn0; m1 = mn1;
      l0 *= corr0; l1 *= corr1;
      o[0] *= corr0; o[1] *= corr0; o[2] *= corr1; o[3] *= corr1;
      if constexpr (F16P) {
        float ls[4] = {l0, 0.f, l1, 0.f};   // row sums ride in an accumulator fragment: c0 = row g, c2 = row g + 8
        const float2 sc = make_float2(scale_log2e, scale_log2e), n0 = make_float2(-mn0, -mn0), n1 = make_float2(-mn1, -mn1);
#pragma unroll
        for (int j = 0; j < 8; j += 2) {
          const float2 e0 = ffma2(make_float2(s[j][0], s[j][1]), sc, n0);
          const float2 e1 = ffma2(make_float2(s[j][2], s[j][3]), sc, n1);
          const float2 e2 = ffma2(make_float2(s[j + 1][0], s[j + 1][1]), sc, n0);
          const float2 e3 = ffma2(make_float2(s[j + 1][2], s[j + 1][3]), sc, n1);
          // A fragment of m16n8k16: (row g, keys 2t..), (row g+8, keys 2t..), (row g, keys 8+2t..), (row g+8, keys 8+2t..)
          const uint32_t pa0 = ex2_f16x2(e0.x, e0.y);
          const uint32_t pa1 = POLY >= 2 ? ex2_poly_f16x2(e1.x, e1.y) : ex2_f16x2(e1.x, e1.y);
          const uint32_t pa2 = ex2_f16x2(e2.x, e2.y);
          const uint32_t pa3 = POLY >= 1 ? ex2_poly_f16x2(e3.x, e3.y) : ex2_f16x2(e3.x, e3.y);
          const __half* vt = reinterpret_cast<const __half*>(sVt) + g * kAtt8VtStride + kb0 + j * 8 + 2 * t;
          const uint32_t b0 = *reinterpret_cast<const uint32_t*>(vt), b1 = *reinterpret_cast<const uint32_t*>(vt + 8);
          mma_m16n8k16_f16(o, pa0, pa1, pa2, pa3, b0, b1);
          mma_m16n8k16_f16(ls, pa0, pa1, pa2, pa3, 0x3C003C00u, 0x3C003C00u);   // times ones: the row sums
        }
        l0 = ls[0], l1 = ls[2];
        continue;
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float2 e01 = ffma2(make_float2(s[j][0], s[j][1]), make_float2(scale_log2e, scale_log2e),
                                 make_float2(-mn0, -mn0));
        const float2 e23 = ffma2(make_float2(s[j][2], s[j][3]), make_float2(scale_log2e, scale_log2e),
                                 make_float2(-mn1, -mn1));
        // (evaluating every fourth exponential with an FMA-pipe cubic - the FlashAttention-4 trick - was measured 3.6 %
        // SLOWER here, 0.350 vs 0.338 ms at B=16, 64 heads, T=1024: the loop is issue-bound, not MUFU-bound)
        const float p0 = ex2_approx(e01.x), p1 = ex2_approx(e01.y), p2 = ex2_approx(e23.x), p3 = ex2_approx(e23.y);
        l0 += p0 + p1;
        l1 += p2 + p3;
        const uint32_t pa0 = pack_bf16x2(p0, p1), pa1 = pack_bf16x2(p2, p3);
        // B fragment of P*V: (k = keys 2t,2t+1 ; n = dim g) from the transposed V tile
        const uint32_t bv = *reinterpret_cast<const uint32_t*>(sVt + g * kAtt8VtStride + kb0 + j * 8 + 2 * t);
        mma_m16n8k8_bf16(o, pa0, pa1, bv);
      }
    }
  }
  if constexpr (!F16P) {   // (the ones-matrix MMA already summed over the four lanes of a row)
    l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
    l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
  }
  const float i0 = 1.0f / l0, i1 = 1.0f / l1;
  __nv_bfloat16* ob = out + b * o_sb + h * o_sh;
  if (q0 + g < Tq)
    *reinterpret_cast<uint32_t*>(ob + (int64_t)(q0 + g) * o_st + 2 * t) = pack_bf16x2(o[0] * i0, o[1] * i0);
  if (q0 + g + 8 < Tq)
    *reinterpret_cast<uint32_t*>(ob + (int64_t)(q0 + g + 8) * o_st + 2 * t) = pack_bf16x2(o[2] * i1, o[3] * i1);
}

// ---------------------------------------------------------------------------------------------------------------
// head_dim 16 / 32 / 64 (SpatialSelfAttention: 4 heads x 64 in the EfficientUNetND mid block and in the KL decoder,
// where T = 4096 makes the scalar kernel above the largest item of the decode): flash attention on mma.sync m16n8k16
// (bf16 in, fp32 accumulate).  One warp owns 16 query rows; per block of 64 keys S = Q K^T is HD/16 k-steps into
// eight m16n8 accumulators, which - converted to bf16 - are exactly the A fragments of P V (two adjacent n-tiles form
// one k = 16 step); K is staged row-major with 16 bytes of row padding, V transposed ([dim][key]), so every fragment
// load is one conflict-free 32-bit shared-memory read.
// ---------------------------------------------------------------------------------------------------------------
constexpr int kAttMWarps = 8;
constexpr int kAttMThreads = kAttMWarps * 32;
constexpr int kAttMKeyChunk = 128;  // keys staged per pass

template <int HD>
__global__ void __launch_bounds__(kAttMThreads) attention_mma_kernel(
    const __nv_bfloat16* __restrict__ q, const __nv_bfloat16* __restrict__ k, const __nv_bfloat16* __restrict__ v,
    __nv_bfloat16* __restrict__ out, int Tq, int Tk, int64_t q_sb, int64_t q_sh, int64_t q_st, int64_t kv_sb,
    int64_t kv_sh, int64_t kv_st, int64_t o_sb, int64_t o_sh, int64_t o_st, float scale_log2e) {
  pdl_enter();
  constexpr int kKS = HD + 8;                  // K row stride (elements): +16 B => the 8 keys of a fragment hit 8 bank groups
  constexpr int kVS = kAttMKeyChunk + 8;       // V^T row stride
  constexpr int kSteps = HD / 16;              // k-steps of Q K^T
  constexpr int kDimTiles = HD / 8;            // n-tiles of P V
  __shared__ __align__(16) __nv_bfloat16 sK[kAttMKeyChunk * kKS];
  __shared__ __align__(16) __nv_bfloat16 sVt[HD * kVS];
  const int b = blockIdx.z, h = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  const int q0 = blockIdx.x * (kAttMWarps * 16) + warp * 16;
  const __nv_bfloat16* qb = q + b * q_sb + h * q_sh;
  const __nv_bfloat16* kb = k + b * kv_sb + h * kv_sh;
  const __nv_bfloat16* vb = v + b * kv_sb + h * kv_sh;

  // Q fragments for every k-step (rows clamped; stores are masked)
  const int r0 = min(q0 + g, Tq - 1), r1 = min(q0 + g + 8, Tq - 1);
  uint32_t qa[kSteps][4];
#pragma unroll
  for (int kk = 0; kk < kSteps; ++kk) {
    qa[kk][0] = *reinterpret_cast<const uint32_t*>(qb + (int64_t)r0 * q_st + kk * 16 + 2 * t);
    qa[kk][1] = *reinterpret_cast<const uint32_t*>(qb + (int64_t)r1 * q_st + kk * 16 + 2 * t);
    qa[kk][2] = *reinterpret_cast<const uint32_t*>(qb + (int64_t)r0 * q_st + kk * 16 + 8 + 2 * t);
    qa[kk][3] = *reinterpret_cast<const uint32_t*>(qb + (int64_t)r1 * q_st + kk * 16 + 8 + 2 * t);
  }
  float o[kDimTiles][4];
#pragma unroll
  for (int n = 0; n < kDimTiles; ++n) o[n][0] = o[n][1] = o[n][2] = o[n][3] = 0.f;
  float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;

  for (int c0 = 0; c0 < Tk; c0 += kAttMKeyChunk) {
    const int nk = min(kAttMKeyChunk, Tk - c0);
    const int nk_pad = (nk + 63) & ~63;
    __syncthreads();
    // stage K rows (16-byte chunks) and V transposed
    for (int i = threadIdx.x; i < nk_pad * (HD / 8); i += kAttMThreads) {
      const int j = i / (HD / 8), ch = i % (HD / 8);
      uint4 kk4 = make_uint4(0, 0, 0, 0), vv4 = make_uint4(0, 0, 0, 0);
      if (j < nk) {
        kk4 = *reinterpret_cast<const uint4*>(kb + (int64_t)(c0 + j) * kv_st + ch * 8);
        vv4 = *reinterpret_cast<const uint4*>(vb + (int64_t)(c0 + j) * kv_st + ch * 8);
      }
      *reinterpret_cast<uint4*>(sK + j * kKS + ch * 8) = kk4;
      const __nv_bfloat16* ve = reinterpret_cast<const __nv_bfloat16*>(&vv4);
#pragma unroll
      for (int d = 0; d < 8; ++d) sVt[(ch * 8 + d) * kVS + j] = ve[d];
    }
    __syncthreads();

    for (int kb0 = 0; kb0 < nk_pad; kb0 += 64) {
      float s[8][4];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        s[j][0] = s[j][1] = s[j][2] = s[j][3] = 0.f;
        const __nv_bfloat16* krow = sK + (kb0 + j * 8 + g) * kKS + 2 * t;
#pragma unroll
        for (int kk = 0; kk < kSteps; ++kk) {
          const uint32_t b0 = *reinterpret_cast<const uint32_t*>(krow + kk * 16);
          const uint32_t b1 = *reinterpret_cast<const uint32_t*>(krow + kk * 16 + 8);
          mma_m16n8k16_bf16(s[j], qa[kk][0], qa[kk][1], qa[kk][2], qa[kk][3], b0, b1);
        }
      }
      if (kb0 + 64 > nk) {  // mask the padded keys of the last block
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int key = kb0 + j * 8 + 2 * t;
          if (key >= nk) { s[j][0] = -INFINITY; s[j][2] = -INFINITY; }
          if (key + 1 >= nk) { s[j][1] = -INFINITY; s[j][3] = -INFINITY; }
        }
      }
      float mx0 = s[0][0], mx1 = s[0][2];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        mx0 = fmaxf(mx0, fmaxf(s[j][0], s[j][1]));
        mx1 = fmaxf(mx1, fmaxf(s[j][2], s[j][3]));
      }
      mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1));
      mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
      mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
      mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
      const float mn0 = fmaxf(m0, mx0 * scale_log2e), mn1 = fmaxf(m1, mx1 * scale_log2e);
      const float corr0 = ex2_approx(m0 - mn0), corr1 = ex2_approx(m1 - mn1);
      m0 = mn0; m1 = mn1;
      l0 *= corr0; l1 *= corr1;
#pragma unroll
      for (int n = 0; n < kDimTiles; ++n) { o[n][0] *= corr0; o[n][1] *= corr0; o[n][2] *= corr1; o[n][3] *= corr1; }
      uint32_t pa[8][2];  // P as bf16x2: [n-tile][row g | row g+8]
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float2 e01 = ffma2(make_float2(s[j][0], s[j][1]), make_float2(scale_log2e, scale_log2e),
                                 make_float2(-mn0, -mn0));
        const float2 e23 = ffma2(make_float2(s[j][2], s[j][3]), make_float2(scale_log2e, scale_log2e),
                                 make_float2(-mn1, -mn1));
        const float p0 = ex2_approx(e01.x), p1 = ex2_approx(e01.y), p2 = ex2_approx(e23.x), p3 = ex2_approx(e23.y);
        l0 += p0 + p1;
        l1 += p2 + p3;
        pa[j][0] = pack_bf16x2(p0, p1);
        pa[j][1] = pack_bf16x2(p2, p3);
      }
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {  // 16 keys per step = n-tiles 2ks, 2ks+1 of S
        const int key0 = kb0 + ks * 16 + 2 * t;
#pragma unroll
        for (int n = 0; n < kDimTiles; ++n) {
          const __nv_bfloat16* vrow = sVt + (n * 8 + g) * kVS + key0;
          const uint32_t b0 = *reinterpret_cast<const uint32_t*>(vrow);
          const uint32_t b1 = *reinterpret_cast<const uint32_t*>(vrow + 8);
          mma_m16n8k16_bf16(o[n], pa[2 * ks][0], pa[2 * ks][1], pa[2 * ks + 1][0], pa[2 * ks + 1][1], b0, b1);
        }
      }
    }
  }
  l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
  l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
  const float i0 = 1.0f / l0, i1 = 1.0f / l1;
  __nv_bfloat16* ob = out + b * o_sb + h * o_sh;
#pragma unroll
  for (int n = 0; n < kDimTiles; ++n) {
    if (q0 + g < Tq)
      *reinterpret_cast<uint32_t*>(ob + (int64_t)(q0 + g) * o_st + n * 8 + 2 * t) = pack_bf16x2(o[n][0] * i0, o[n][1] * i0);
    if (q0 + g + 8 < Tq)
      *reinterpret_cast<uint32_t*>(ob + (int64_t)(q0 + g + 8) * o_st + n * 8 + 2 * t) =
          pack_bf16x2(o[n][2] * i1, o[n][3] * i1);
  }
}

template <int HD>
static int launch_attention_mma(const void* q, const void* k, const void* v, void* out, int B, int heads, int Tq,
                                int Tk, int64_t q_sb, int64_t q_sh, int64_t q_st, int64_t kv_sb, int64_t kv_sh,
                                int64_t kv_st, int64_t o_sb, int64_t o_sh, int64_t o_st, float scale, cudaStream_t st) {
  dim3 grid((Tq + kAttMWarps * 16 - 1) / (kAttMWarps * 16), heads, B);
  launch_pdl(attention_mma_kernel<HD>, dim3(grid), dim3(kAttMThreads), 0, st, 
      reinterpret_cast<const __nv_bfloat16*>(q), reinterpret_cast<const __nv_bfloat16*>(k),
      reinterpret_cast<const __nv_bfloat16*>(v), reinterpret_cast<__nv_bfloat16*>(out), Tq, Tk, q_sb, q_sh, q_st,
      kv_sb, kv_sh, kv_st, o_sb, o_sh, o_st, scale * 1.4426950408889634f);
  FM_LAUNCH_CHECK("attention_mma_kernel");
  return 0;
}

template <int HD>
static int launch_attention(const void* q, const void* k, const void* v, void* out, int B, int heads, int Tq, int Tk,
                            int64_t q_sb, int64_t q_sh, int64_t q_st, int64_t kv_sb, int64_t kv_sh, int64_t kv_st,
                            int64_t o_sb, int64_t o_sh, int64_t o_st, float scale, cudaStream_t st) {
  int kchunk = 8192 / HD;  // 32 KB of K+V per chunk
  if (kchunk > Tk) kchunk = (Tk + 3) / 4 * 4;
  const size_t smem = (size_t)kchunk * (HD / 8) * 16 * 2;
  dim3 grid((Tq + kAttThreads - 1) / kAttThreads, heads, B);
  attention_kernel<HD><<<grid, kAttThreads, smem, st>>>(
      reinterpret_cast<const __nv_bfloat16*>(q), reinterpret_cast<const __nv_bfloat16*>(k),
      reinterpret_cast<const __nv_bfloat16*>(v), reinterpret_cast<__nv_bfloat16*>(out), Tq, Tk, q_sb, q_sh, q_st,
      kv_sb, kv_sh, kv_st, o_sb, o_sh, o_st, scale * 1.4426950408889634f, kchunk);
  FM_LAUNCH_CHECK("attention_kernel");
  return 0;
}

// =============================================================================================================
// Cross-attention context path: GroupNorm over the context tokens + the tiny-K key/value
// projection (`attention.py:149-165` SpatialCrossAttention.kv_proj, `attention.py:232-262` DiffusersAttentionND
// to_k/to_v with context_dim): ctx fp32 [B][Cc][Tc] -> bf16 K|V.  Cc (the latent channel count, 4 in the LDCT
// configs) is far below a tensor-core tile, so this is two small CUDA-core kernels; K/V do not depend on the
// sampling step, so the host computes them once per run.
// =============================================================================================================
constexpr int kCtxMaxC = 16;

// stats[b][g] = (mean, rstd) over the group's (Cc/groups) x Tc values; one block per (group, sample)
__global__ void __launch_bounds__(256) context_stats_kernel(const float* __restrict__ ctx, float* __restrict__ stats,
                                                             int Cc, int Tc, int groups, float eps) {
  __shared__ double red[2][256];
  const int g = blockIdx.x, b = blockIdx.y;
  const int cpg = Cc / groups;
  const int64_t n = (int64_t)cpg * Tc;
  const float* src = ctx + ((int64_t)b * Cc + (int64_t)g * cpg) * Tc;  // the group's channels are contiguous
  double s = 0.0, ss = 0.0;
  for (int64_t i = threadIdx.x; i < n; i += blockDim.x) {
    const double v = src[i];
    s += v;
    ss += v * v;
  }
  red[0][threadIdx.x] = s;
  red[1][threadIdx.x] = ss;
  __syncthreads();
  for (int k = 128; k > 0; k >>= 1) {
    if (threadIdx.x < k) {
      red[0][threadIdx.x] += red[0][threadIdx.x + k];
      red[1][threadIdx.x] += red[1][threadIdx.x + k];
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const double mean = red[0][0] / (double)n;
    double var = red[1][0] / (double)n - mean * mean;
    if (var < 0.0) var = 0.0;
    stats[((int64_t)b * groups + g) * 2 + 0] = (float)mean;
    stats[((int64_t)b * groups + g) * 2 + 1] = (float)(1.0 / sqrt(var + (double)eps));
  }
}

// out[b][t][o] (token-major, LAYOUT 0) or out[b][o][t] (channel-major, LAYOUT 1)
//   = bias[o] + sum_c W[o][c] * ((ctx[b][c][t] - mean) * rstd * gamma[c] + beta[c])
template <int LAYOUT>
__global__ void __launch_bounds__(256) context_proj_kernel(const float* __restrict__ ctx,
                                                            const float* __restrict__ stats,
                                                            const float* __restrict__ gamma,
                                                            const float* __restrict__ beta,
                                                            const float* __restrict__ W, const float* __restrict__ bias,
                                                            __nv_bfloat16* __restrict__ out, int Cc, int Tc, int O,
                                                            int groups) {
  constexpr int kTok = 64;
  __shared__ float sn[kCtxMaxC][kTok];
  const int b = blockIdx.y, t0 = blockIdx.x * kTok;
  const int cpg = Cc / groups;
  for (int i = threadIdx.x; i < Cc * kTok; i += blockDim.x) {
    const int c = i / kTok, tt = i - c * kTok;
    float v = 0.f;
    if (t0 + tt < Tc) {
      const int g = c / cpg;
      const float mean = stats[((int64_t)b * groups + g) * 2 + 0], rstd = stats[((int64_t)b * groups + g) * 2 + 1];
      v = fmaf((ctx[((int64_t)b * Cc + c) * Tc + t0 + tt] - mean) * rstd, gamma[c], beta[c]);
    }
    sn[c][tt] = v;
  }
  __syncthreads();
  const int ntok = min(kTok, Tc - t0);
  if (LAYOUT == 0) {
    // thread = output feature (its weight row in registers); tokens in the inner loop: stores coalesce over o
    for (int o = threadIdx.x; o < O; o += blockDim.x) {
      float w[kCtxMaxC];
#pragma unroll
      for (int c = 0; c < kCtxMaxC; ++c) w[c] = c < Cc ? W[(int64_t)o * Cc + c] : 0.f;
      const float bo = bias != nullptr ? bias[o] : 0.f;
      for (int tt = 0; tt < ntok; ++tt) {
        float acc = bo;
#pragma unroll
        for (int c = 0; c < kCtxMaxC; ++c)
          if (c < Cc) acc = fmaf(w[c], sn[c][tt], acc);
        out[((int64_t)b * Tc + t0 + tt) * O + o] = __float2bfloat16_rn(acc);
      }
    }
  } else {
    // thread = (token, feature lane); the token's normalised context vector in registers: stores coalesce over tokens
    const int tt = threadIdx.x % kTok, lane = threadIdx.x / kTok;
    float n[kCtxMaxC];
#pragma unroll
    for (int c = 0; c < kCtxMaxC; ++c) n[c] = c < Cc ? sn[c][tt] : 0.f;
    if (tt < ntok) {
      for (int o = lane; o < O; o += blockDim.x / kTok) {
        float acc = bias != nullptr ? bias[o] : 0.f;
#pragma unroll
        for (int c = 0; c < kCtxMaxC; ++c)
          if (c < Cc) acc = fmaf(__ldg(W + (int64_t)o * Cc + c), n[c], acc);
        out[((int64_t)b * O + o) * Tc + t0 + tt] = __float2bfloat16_rn(acc);
      }
    }
  }
}


// =============================================================================================================
// Linear attention (`attention.py:53-70` LinearQKVAttention; EfficientUNetND's default inside its levels):
//   ks = softmax(k over tokens), qs = softmax(q over features), ctx = ks^T v / (sum_n ks + eps), out = qs ctx.
// O(T * d^2): one CTA per (sample, head); the d x d context matrix lives in shared memory.  fp32 CUDA-core math.
// =============================================================================================================
template <int HD>
__global__ void __launch_bounds__(256) linear_attention_kernel(
    const __nv_bfloat16* __restrict__ q, const __nv_bfloat16* __restrict__ k, const __nv_bfloat16* __restrict__ v,
    __nv_bfloat16* __restrict__ out, int heads, int Tq, int Tk, int64_t q_sb, int64_t q_sh, int64_t q_st, int64_t kv_sb,
    int64_t kv_sh, int64_t kv_st, int64_t o_sb, int64_t o_sh, int64_t o_st, float eps) {
  pdl_enter();
  constexpr int kLanes = 256 / HD;   // token lanes per feature column
  constexpr int kChunk = 32;         // tokens staged per step of the context accumulation
  constexpr int kPairs = HD * HD / 256 > 0 ? HD * HD / 256 : 1;
  __shared__ float ctx[HD][HD + 1];
  __shared__ float red[256];
  __shared__ float cmax[HD], csum[HD], cden[HD];
  __shared__ float sk[kChunk][HD], sv[kChunk][HD];
  const int b = blockIdx.x / heads, h = blockIdx.x % heads;
  const __nv_bfloat16* kb = k + b * kv_sb + h * kv_sh;
  const __nv_bfloat16* vb = v + b * kv_sb + h * kv_sh;
  const __nv_bfloat16* qb = q + b * q_sb + h * q_sh;
  __nv_bfloat16* ob = out + b * o_sb + h * o_sh;
  const int d = threadIdx.x % HD, lane = threadIdx.x / HD;
  // column (per-feature) max of k over the tokens
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
  // ctx[d][e] = sum_n ks[n][d] v[n][e]; den[d] = sum_n ks[n][d]
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
      sk[nn][dd] = kv_;
      sv[nn][dd] = vv;
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < kPairs; ++j) {
      const int p = threadIdx.x + 256 * j;
      if (p < HD * HD) {
        const int dd = p / HD, ee = p - dd * HD;
        float a = acc[j];
#pragma unroll 8
        for (int nn = 0; nn < kChunk; ++nn) a = fmaf(sk[nn][dd], sv[nn][ee], a);
        acc[j] = a;
      }
    }
    if (threadIdx.x < HD)
      for (int nn = 0; nn < kChunk; ++nn) den += sk[nn][threadIdx.x];
  }
  if (threadIdx.x < HD) cden[threadIdx.x] = den + eps;
  __syncthreads();
#pragma unroll
  for (int j = 0; j < kPairs; ++j) {
    const int p = threadIdx.x + 256 * j;
    if (p < HD * HD) {
      const int dd = p / HD, ee = p - dd * HD;
      ctx[dd][ee] = acc[j] / cden[dd];
    }
  }
  __syncthreads();
  // out[n][e] = sum_d softmax_d(q[n])[d] ctx[d][e]
  for (int n = threadIdx.x; n < Tq; n += 256) {
    float qr[HD];
    float qm = -INFINITY;
#pragma unroll
    for (int dd = 0; dd < HD; ++dd) {
      qr[dd] = __bfloat162float(qb[n * q_st + dd]);
      qm = fmaxf(qm, qr[dd]);
    }
    float qs = 0.f;
#pragma unroll
    for (int dd = 0; dd < HD; ++dd) {
      qr[dd] = __expf(qr[dd] - qm);
      qs += qr[dd];
    }
    const float inv = 1.f / qs;
    for (int e0 = 0; e0 < HD; e0 += 8) {
      float o8[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll
      for (int dd = 0; dd < HD; ++dd) {
        const float w = qr[dd] * inv;
#pragma unroll
        for (int e = 0; e < 8; ++e) o8[e] = fmaf(w, ctx[dd][e0 + e], o8[e]);
      }
      *reinterpret_cast<uint4*>(ob + n * o_st + e0) =
          make_uint4(pack_bf16x2(o8[0], o8[1]), pack_bf16x2(o8[2], o8[3]), pack_bf16x2(o8[4], o8[5]),
                     pack_bf16x2(o8[6], o8[7]));
    }
  }
}

}  // namespace fm

using namespace fm;

extern "C" int fm_attention_bf16(const void* q, const void* k, const void* v, void* out, int32_t B, int32_t heads,
                                 int32_t Tq, int32_t Tk, int32_t head_dim, int64_t q_sb, int64_t q_sh, int64_t q_st,
                                 int64_t kv_sb, int64_t kv_sh, int64_t kv_st, int64_t o_sb, int64_t o_sh, int64_t o_st,
                                 float scale, fm_stream_t stream) {
  if (int e = ensure_device()) return e;
  FM_REQUIRE(q && k && v && out, "attention: null pointer");
  FM_REQUIRE(B > 0 && heads > 0 && Tq > 0 && Tk > 0, "attention: empty problem");
  FM_REQUIRE(heads <= 65535 && B <= 65535, "attention: grid too large");
  FM_REQUIRE((((uintptr_t)q | (uintptr_t)k | (uintptr_t)v | (uintptr_t)out) & 15) == 0,
             "attention: pointers must be 16B aligned");
  FM_REQUIRE(((q_sb | q_sh | q_st | kv_sb | kv_sh | kv_st | o_sb | o_sh | o_st) & 7) == 0,
             "attention: strides must be multiples of 8 elements");
  cudaStream_t st = (cudaStream_t)stream;
  const bool scalar = getenv("FMDM_ATTENTION_SCALAR") != nullptr;  // A/B switch: the CUDA-core kernel for head_dim >= 16
#define FM_ATT_ARGS q, k, v, out, B, heads, Tq, Tk, q_sb, q_sh, q_st, kv_sb, kv_sh, kv_st, o_sb, o_sh, o_st, scale, st
  switch (head_dim) {
    case 8: {
      dim3 grid((Tq + kAtt8Warps * 16 - 1) / (kAtt8Warps * 16), heads, B);
      // FMDM_ATTENTION_BF16P=1: the first-generation inner loop (scalar fp32 exponentials, bf16 probabilities), A/B only
      static const bool bf16p = getenv("FMDM_ATTENTION_BF16P") != nullptr;
      // FMDM_ATTENTION_POLY=0|1|2: share of the exponentials on the FMA pipe (0, 1/4, 1/2); A/B switch
      static const int poly = getenv("FMDM_ATTENTION_POLY") ? atoi(getenv("FMDM_ATTENTION_POLY")) : kAtt8PolyDefault;
      auto kern = bf16p ? attention_hd8_mma_kernel<false, 0>
                        : (poly >= 2 ? attention_hd8_mma_kernel<true, 2>
                                     : (poly == 1 ? attention_hd8_mma_kernel<true, 1> : attention_hd8_mma_kernel<true, 0>));
      launch_pdl(kern, grid, dim3(kAtt8Threads), 0, st,
          reinterpret_cast<const __nv_bfloat16*>(q), reinterpret_cast<const __nv_bfloat16*>(k),
          reinterpret_cast<const __nv_bfloat16*>(v), reinterpret_cast<__nv_bfloat16*>(out), Tq, Tk, q_sb, q_sh, q_st,
          kv_sb, kv_sh, kv_st, o_sb, o_sh, o_st, scale * 1.4426950408889634f);
      FM_LAUNCH_CHECK("attention_hd8_mma_kernel");
      return 0;
    }
    case 16: return scalar ? launch_attention<16>(FM_ATT_ARGS) : launch_attention_mma<16>(FM_ATT_ARGS);
    case 32: return scalar ? launch_attention<32>(FM_ATT_ARGS) : launch_attention_mma<32>(FM_ATT_ARGS);
    case 64: return scalar ? launch_attention<64>(FM_ATT_ARGS) : launch_attention_mma<64>(FM_ATT_ARGS);
    default:
      set_error("attention: head_dim=%d unsupported (8, 16, 32, 64)", head_dim);
      return FM_ERR_UNSUPPORTED;
  }
#undef FM_ATT_ARGS
}

extern "C" int fm_context_kv_bf16(const float* ctx, const float* gamma, const float* beta, const float* W,
                                  const float* bias, float* stats_ws, void* out, int32_t B, int32_t Cc, int32_t Tc,
                                  int32_t O, int32_t groups, float eps, int32_t channel_major, fm_stream_t stream) {
  if (int e = ensure_device()) return e;
  FM_REQUIRE(ctx && gamma && beta && W && stats_ws && out, "context_kv: null pointer");
  FM_REQUIRE(B > 0 && Tc > 0 && O > 0 && Cc > 0 && Cc <= kCtxMaxC, "context_kv: context_dim must be 1..%d (got %d)",
             kCtxMaxC, Cc);
  FM_REQUIRE(groups > 0 && Cc % groups == 0, "context_kv: groups must divide context_dim");
  cudaStream_t st = (cudaStream_t)stream;
  context_stats_kernel<<<dim3(groups, B), 256, 0, st>>>(ctx, stats_ws, Cc, Tc, groups, eps);
  FM_LAUNCH_CHECK("context_stats_kernel");
  const dim3 grid((Tc + 63) / 64, B);
  if (channel_major)
    context_proj_kernel<1><<<grid, 256, 0, st>>>(ctx, stats_ws, gamma, beta, W, bias,
                                                 reinterpret_cast<__nv_bfloat16*>(out), Cc, Tc, O, groups);
  else
    context_proj_kernel<0><<<grid, 256, 0, st>>>(ctx, stats_ws, gamma, beta, W, bias,
                                                 reinterpret_cast<__nv_bfloat16*>(out), Cc, Tc, O, groups);
  FM_LAUNCH_CHECK("context_proj_kernel");
  return 0;
}

extern "C" int fm_linear_attention_bf16(const void* q, const void* k, const void* v, void* out, int32_t B,
                                        int32_t heads, int32_t Tq, int32_t Tk, int32_t head_dim, int64_t q_sb,
                                        int64_t q_sh, int64_t q_st, int64_t kv_sb, int64_t kv_sh, int64_t kv_st,
                                        int64_t o_sb, int64_t o_sh, int64_t o_st, float eps, fm_stream_t stream) {
  if (int e = ensure_device()) return e;
  FM_REQUIRE(q && k && v && out, "linear_attention: null pointer");
  FM_REQUIRE(B > 0 && heads > 0 && Tq > 0 && Tk > 0, "linear_attention: empty problem");
  FM_REQUIRE((((uintptr_t)out) & 15) == 0 && ((o_sb | o_sh | o_st) & 7) == 0,
             "linear_attention: output rows must be 16B aligned");
  cudaStream_t st = (cudaStream_t)stream;
#define FM_LIN(HD)                                                                                                  \
  case HD:                                                                                                          \
    launch_pdl(linear_attention_kernel<HD>, dim3(B * heads), dim3(256), 0, st,                                                          \
        reinterpret_cast<const __nv_bfloat16*>(q), reinterpret_cast<const __nv_bfloat16*>(k),                       \
        reinterpret_cast<const __nv_bfloat16*>(v), reinterpret_cast<__nv_bfloat16*>(out), heads, Tq, Tk, q_sb, q_sh, \
        q_st, kv_sb, kv_sh, kv_st, o_sb, o_sh, o_st, eps);                                                          \
    break;
  switch (head_dim) {
    FM_LIN(8)
    FM_LIN(16)
    FM_LIN(32)
    FM_LIN(64)
    default:
      set_error("linear_attention: head_dim=%d unsupported (8, 16, 32, 64)", head_dim);
      return FM_ERR_UNSUPPORTED;
  }
#undef FM_LIN
  FM_LAUNCH_CHECK("linear_attention_kernel");
  return 0;
}
