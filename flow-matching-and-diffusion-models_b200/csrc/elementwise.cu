// K4 scheduler-step kernels and the small HBM-bound helpers (time embedding, tiny fp32 Linear, weight pre-pack,
// nearest upsample, transpose, clamp).  All fp32 scheduler arithmetic uses explicit round-to-nearest intrinsics in
// the same operation order as diffusers' eager PyTorch code so results are bit-identical to the oracle
// (oracle/schedulers.py), i.e. no FMA contraction.
#include "common.cuh"

namespace fm {

__device__ __forceinline__ int pick_step(const int32_t* step_dev, int step_host) {
  return step_dev != nullptr ? *step_dev : step_host;
}

// x_out = x + dt * v                                   (FlowMatchEulerDiscreteScheduler.step)
__global__ void __launch_bounds__(256) sched_flowmatch_kernel(float* __restrict__ xo, const float* __restrict__ x,
                                                             const float* __restrict__ v,
                                                             const float* __restrict__ coef,
                                                             const int32_t* __restrict__ step_dev, int step_host,
                                                             int64_t n) {
  pdl_enter();
  const float dt = coef[pick_step(step_dev, step_host) * FM_FLOWMATCH_NCOEF];
  const int64_t n4 = n >> 2;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    const float4 a = reinterpret_cast<const float4*>(x)[i];
    const float4 b = __ldcs(reinterpret_cast<const float4*>(v) + i);
    float4 o;
    o.x = __fadd_rn(a.x, __fmul_rn(dt, b.x));
    o.y = __fadd_rn(a.y, __fmul_rn(dt, b.y));
    o.z = __fadd_rn(a.z, __fmul_rn(dt, b.z));
    o.w = __fadd_rn(a.w, __fmul_rn(dt, b.w));
    reinterpret_cast<float4*>(xo)[i] = o;
  }
  for (int64_t i = (n4 << 2) + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    xo[i] = __fadd_rn(x[i], __fmul_rn(dt, v[i]));
}

// DDPM ancestral step (epsilon prediction, variance_type "fixed_small"):
//   x0 = clamp((x - sqrt(1-a_t) e) / sqrt(a_t)); x = (c_x0 * x0 + c_xt * x) + sigma * noise       (DDPMScheduler.step)
__device__ __forceinline__ float ddpm_one(float x, float e, float z, float sb, float sa, float c0, float c1,
                                          float sigma, int clip, float cr) {
  float x0 = __fdiv_rn(__fsub_rn(x, __fmul_rn(sb, e)), sa);
  if (clip) x0 = fminf(fmaxf(x0, -cr), cr);
  const float mean = __fadd_rn(__fmul_rn(c0, x0), __fmul_rn(c1, x));
  return __fadd_rn(mean, __fmul_rn(sigma, z));
}
__global__ void __launch_bounds__(256) sched_ddpm_kernel(float* __restrict__ xo, const float* __restrict__ x,
                                                        const float* __restrict__ eps,
                                                        const float* __restrict__ noise,
                                                        const float* __restrict__ coef,
                                                        const int32_t* __restrict__ step_dev, int step_host,
                                                        int clip, float cr, int64_t n) {
  pdl_enter();
  const float* c = coef + (size_t)pick_step(step_dev, step_host) * FM_DDPM_NCOEF;
  const float sb = c[0], sa = c[1], c0 = c[2], c1 = c[3], sg = c[4];
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    xo[i] = ddpm_one(x[i], eps[i], noise[i], sb, sa, c0, c1, sg, clip, cr);
}

// DDIM (eta = 0, epsilon prediction):  x0 = (x - sqrt(1-a_t) e) / sqrt(a_t); clamp; x = sqrt(a_p) x0 + dir * e
__device__ __forceinline__ float ddim_one(float x, float e, float sb, float sa, float sp, float dc, int clip,
                                          float cr) {
  float x0 = __fdiv_rn(__fsub_rn(x, __fmul_rn(sb, e)), sa);
  if (clip) x0 = fminf(fmaxf(x0, -cr), cr);
  return __fadd_rn(__fmul_rn(sp, x0), __fmul_rn(dc, e));
}
__global__ void __launch_bounds__(256) sched_ddim_kernel(float* __restrict__ xo, const float* __restrict__ x,
                                                        const float* __restrict__ eps,
                                                        const float* __restrict__ coef,
                                                        const int32_t* __restrict__ step_dev, int step_host,
                                                        int clip, float cr, int64_t n) {
  pdl_enter();
  const float* c = coef + (size_t)pick_step(step_dev, step_host) * FM_DDIM_NCOEF;
  const float sb = c[0], sa = c[1], sp = c[2], dc = c[3];
  const int64_t n4 = n >> 2;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    const float4 a = reinterpret_cast<const float4*>(x)[i];
    const float4 b = __ldcs(reinterpret_cast<const float4*>(eps) + i);
    float4 o;
    o.x = ddim_one(a.x, b.x, sb, sa, sp, dc, clip, cr);
    o.y = ddim_one(a.y, b.y, sb, sa, sp, dc, clip, cr);
    o.z = ddim_one(a.z, b.z, sb, sa, sp, dc, clip, cr);
    o.w = ddim_one(a.w, b.w, sb, sa, sp, dc, clip, cr);
    reinterpret_cast<float4*>(xo)[i] = o;
  }
  for (int64_t i = (n4 << 2) + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    xo[i] = ddim_one(x[i], eps[i], sb, sa, sp, dc, clip, cr);
}

// DPM-Solver++(2M), midpoint, data prediction.  m = (x - sigma_s e)/alpha_s;
//   first order : x = c1 x - c2 m ;  second order: x = c1 x - c2 m - c3 * (inv_r0 * (m - m_prev))
// raw (algorithm_type "dpmsolver", noise prediction): the solver integrates epsilon itself, m = e.
__device__ __forceinline__ float dpmpp_one(float x, float e, float mp, float ss, float as, float c1, float c2,
                                           float c3, float ir, bool second, bool raw, float* m_out) {
  const float m = raw ? e : __fdiv_rn(__fsub_rn(x, __fmul_rn(ss, e)), as);
  *m_out = m;
  float r = __fsub_rn(__fmul_rn(c1, x), __fmul_rn(c2, m));
  if (second) r = __fsub_rn(r, __fmul_rn(c3, __fmul_rn(ir, __fsub_rn(m, mp))));
  return r;
}
__global__ void __launch_bounds__(256) sched_dpmpp_kernel(float* __restrict__ xo, float* __restrict__ m_cur,
                                                         const float* __restrict__ x,
                                                         const float* __restrict__ eps,
                                                         const float* __restrict__ m_prev,
                                                         const float* __restrict__ coef,
                                                         const int32_t* __restrict__ step_dev, int step_host,
                                                         int64_t n) {
  pdl_enter();
  const float* c = coef + (size_t)pick_step(step_dev, step_host) * FM_DPMPP_NCOEF;
  const float ss = c[0], as = c[1], c1 = c[2], c2 = c[3], c3 = c[4], ir = c[5];
  const bool second = c[6] != 0.0f, raw = c[7] != 0.0f;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t n4 = n >> 2;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    const float4 a = reinterpret_cast<const float4*>(x)[i];
    const float4 b = __ldcs(reinterpret_cast<const float4*>(eps) + i);
    float4 mp = make_float4(0.f, 0.f, 0.f, 0.f);
    if (second) mp = reinterpret_cast<const float4*>(m_prev)[i];
    float4 o, m;
    o.x = dpmpp_one(a.x, b.x, mp.x, ss, as, c1, c2, c3, ir, second, raw, &m.x);
    o.y = dpmpp_one(a.y, b.y, mp.y, ss, as, c1, c2, c3, ir, second, raw, &m.y);
    o.z = dpmpp_one(a.z, b.z, mp.z, ss, as, c1, c2, c3, ir, second, raw, &m.z);
    o.w = dpmpp_one(a.w, b.w, mp.w, ss, as, c1, c2, c3, ir, second, raw, &m.w);
    reinterpret_cast<float4*>(xo)[i] = o;
    reinterpret_cast<float4*>(m_cur)[i] = m;
  }
  for (int64_t i = (n4 << 2) + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    float m;
    xo[i] = dpmpp_one(x[i], eps[i], second ? m_prev[i] : 0.f, ss, as, c1, c2, c3, ir, second, raw, &m);
    m_cur[i] = m;
  }
}

// UniPC (bh2, data prediction, order <= 2): corrector on the incoming sample with the new model output, then predictor.
//   m_t = (x - sigma e)/alpha
//   corrector (c[2]): xc = (ca*last - cb*m1) - cc*((c[6] ? rho0*((m2 - m1)/rkc) : 0) + rhoL*(m_t - m1))   else xc = x
//   predictor:        xo = (pa*xc - pb*m_t) - pc*(c[13] ? half*((m1 - m_t)/rkp) : 0)
//   state update:     last <- xc, m2 <- m1, m1 <- m_t
struct UniPCCoef {
  float sg, al, ca, cb, cc, rkc, rho0, rhoL, pa, pb, pc, rkp, half;
  bool corr, corr2, pred2;
};
__device__ __forceinline__ float unipc_one(float x, float e, float last, float m1, float m2, const UniPCCoef& k,
                                           float* xc_out, float* mt_out) {
  const float mt = __fdiv_rn(__fsub_rn(x, __fmul_rn(k.sg, e)), k.al);
  float xc = x;
  if (k.corr) {
    const float xt_ = __fsub_rn(__fmul_rn(k.ca, last), __fmul_rn(k.cb, m1));
    const float d1t = __fmul_rn(k.rhoL, __fsub_rn(mt, m1));
    const float inner = k.corr2 ? __fadd_rn(__fmul_rn(k.rho0, __fdiv_rn(__fsub_rn(m2, m1), k.rkc)), d1t) : d1t;
    xc = __fsub_rn(xt_, __fmul_rn(k.cc, inner));
  }
  *xc_out = xc;
  *mt_out = mt;
  const float pt_ = __fsub_rn(__fmul_rn(k.pa, xc), __fmul_rn(k.pb, mt));
  if (!k.pred2) return pt_;
  return __fsub_rn(pt_, __fmul_rn(k.pc, __fmul_rn(k.half, __fdiv_rn(__fsub_rn(m1, mt), k.rkp))));
}
__global__ void __launch_bounds__(256) sched_unipc_kernel(float* __restrict__ xo, float* __restrict__ last,
                                                         float* __restrict__ m1, float* __restrict__ m2,
                                                         const float* __restrict__ x, const float* __restrict__ eps,
                                                         const float* __restrict__ coef,
                                                         const int32_t* __restrict__ step_dev, int step_host,
                                                         int64_t n) {
  pdl_enter();
  const float* c = coef + (size_t)pick_step(step_dev, step_host) * FM_UNIPC_NCOEF;
  UniPCCoef k;
  k.sg = c[0]; k.al = c[1]; k.corr = c[2] != 0.0f; k.ca = c[3]; k.cb = c[4]; k.cc = c[5]; k.corr2 = c[6] != 0.0f;
  k.rkc = c[7]; k.rho0 = c[8]; k.rhoL = c[9]; k.pa = c[10]; k.pb = c[11]; k.pc = c[12]; k.pred2 = c[13] != 0.0f;
  k.rkp = c[14]; k.half = c[15];
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    float xc, mt;
    const float a1 = m1[i];
    const float r = unipc_one(x[i], __ldcs(eps + i), k.corr ? last[i] : 0.f, a1, k.corr2 ? m2[i] : 0.f, k, &xc, &mt);
    xo[i] = r;
    last[i] = xc;
    m2[i] = a1;
    m1[i] = mt;
  }
}

__global__ void __launch_bounds__(256) add_noise_kernel(float* __restrict__ xo, const float* __restrict__ x0,
                                                       const float* __restrict__ noise,
                                                       const float* __restrict__ a, const float* __restrict__ b,
                                                       int64_t per_sample, int64_t n) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const int s = (int)(i / per_sample);
    xo[i] = __fadd_rn(__fmul_rn(a[s], x0[i]), __fmul_rn(b[s], noise[i]));
  }
}

__global__ void counter_add_kernel(int32_t* ctr, int32_t delta) {
  pdl_enter();
  *ctr += delta;
}

__global__ void __launch_bounds__(256) clamp_kernel(float* __restrict__ y, const float* __restrict__ x, float lo,
                                                   float hi, int64_t n) {
  pdl_enter();
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    y[i] = fminf(fmaxf(x[i], lo), hi);
}

// ---- sinusoidal timestep embedding (src/nn/ops/time_embedding.py:23-31), accurate sin/cos ---------------------
__global__ void timestep_embedding_kernel(const float* __restrict__ t, const float* __restrict__ t_table,
                                          const int32_t* __restrict__ step_dev, float* __restrict__ out, int B,
                                          int dim, float neg_log_period, int flip, float denom) {
  pdl_enter();
  const int half = dim / 2;
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= B * dim) return;
  const int b = idx / dim, j = idx - b * dim;
  const float tv = (t_table != nullptr) ? t_table[*step_dev] : t[b];
  if (j >= 2 * half) { out[idx] = 0.f; return; }
  // reference layout [sin | cos]; flip_sin_to_cos swaps the halves
  const bool second_half = j >= half;
  const int i = second_half ? j - half : j;
  const bool want_cos = flip ? !second_half : second_half;
  const float e = __fdiv_rn(__fmul_rn(neg_log_period, (float)i), denom);
  const float arg = __fmul_rn(tv, expf(e));
  out[idx] = want_cos ? cosf(arg) : sinf(arg);
}

// ---- tiny fp32 Linear: one warp per output feature; the (activated) input rows are staged once in smem ---------
constexpr int kLinRows = 16;
__global__ void __launch_bounds__(128) linear_f32_kernel(const float* __restrict__ x, const float* __restrict__ W,
                                                        const float* __restrict__ bias,
                                                        const float* __restrict__ bias2, float* __restrict__ y,
                                                        int B, int I, int O, int silu_in, int silu_out, int rows) {
  pdl_enter();
  extern __shared__ float sx[];  // [rows][I]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int o = blockIdx.x * 4 + warp;
  const float* w = W + (size_t)(o < O ? o : 0) * I;
  float bsum = 0.f;
  if (o < O) {
    if (bias) bsum += bias[o];
    if (bias2) bsum += bias2[o];
  }
  for (int b0 = 0; b0 < B; b0 += rows) {
    const int nb = min(rows, B - b0);
    __syncthreads();
    for (int i = threadIdx.x; i < nb * I; i += blockDim.x) {
      float xv = x[(size_t)b0 * I + i];
      if (silu_in) xv = xv / (1.0f + expf(-xv));
      sx[i] = xv;
    }
    __syncthreads();
    if (o >= O) continue;
    float acc[kLinRows];
#pragma unroll
    for (int r = 0; r < kLinRows; ++r) acc[r] = 0.f;
    for (int i = lane; i < I; i += 32) {
      const float wv = __ldg(w + i);
#pragma unroll
      for (int r = 0; r < kLinRows; ++r)
        if (r < nb) acc[r] = fmaf(sx[r * I + i], wv, acc[r]);
    }
#pragma unroll
    for (int r = 0; r < kLinRows; ++r) {
      if (r < nb) {
        float s = warp_sum(acc[r]);
        if (lane == 0) {
          s += bsum;
          if (silu_out) s = s / (1.0f + expf(-s));
          y[(size_t)(b0 + r) * O + o] = s;
        }
      }
    }
  }
}

// ---- weight pre-pack: OIHW fp32 -> [Cout][K] bf16, K = (tap, channel) ------------------------------------------
template <bool LO>
__global__ void weight_prepack_kernel(__nv_bfloat16* __restrict__ dst, int64_t dst_row_stride, int64_t koff,
                                      const float* __restrict__ src, int Cout, int Cin_total, int c_begin, int Cseg,
                                      int ks) {
  const int taps = ks * ks;
  const int64_t total = (int64_t)Cout * taps * Cseg;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int c = (int)(i % Cseg);
  const int tap = (int)((i / Cseg) % taps);
  const int co = (int)(i / ((int64_t)Cseg * taps));
  const float v = src[((int64_t)co * Cin_total + c_begin + c) * taps + tap];
  const __nv_bfloat16 hi = __float2bfloat16_rn(v);
  // LO: the rounding residual of the bf16 weight (split-bf16 weights: x*w ~= x*hi + x*lo)
  dst[(int64_t)co * dst_row_stride + koff + (int64_t)tap * Cseg + c] =
      LO ? __float2bfloat16_rn(v - __bfloat162float(hi)) : hi;
}

// Packed weights of the data-gradient conv of a channel slice: dst[ci][tap' * Cout + co] = w[co][c_begin + ci][taps-1-tap']
// (input/output channels swapped, taps mirrored), K-major bf16 [Cseg][taps * Cout].
__global__ void weight_prepack_dgrad_kernel(__nv_bfloat16* __restrict__ dst, const float* __restrict__ src, int Cout,
                                            int Cin_total, int c_begin, int Cseg, int ks) {
  const int taps = ks * ks;
  const int64_t total = (int64_t)Cseg * taps * Cout;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int co = (int)(i % Cout);
  const int tap = (int)((i / Cout) % taps);
  const int ci = (int)(i / ((int64_t)Cout * taps));
  const float v = src[((int64_t)co * Cin_total + c_begin + ci) * taps + (taps - 1 - tap)];
  dst[i] = __float2bfloat16_rn(v);
}

// Batched form of the two pack kernels above: every conv weight of a training step (forward matrices and data-gradient
// matrices) re-packed from the fp32 master parameters in ONE launch.  Blocks are pre-assigned to (entry, offset) pairs
// on the host; an entry is one K segment of one packed matrix and the offset counts (matrix row, channel) PAIRS: a
// thread owns one pair and walks its ksize^2 taps, which lie contiguously in the OIHW source - the nine loads of a lane
// hit the same one or two sectors (a thread per destination element fetched every source sector nine times).
constexpr int kPackPerBlock = 256 * 4;
template <int TAPS>
__device__ __forceinline__ void pack_pair(const fm_pack_entry& e, uint32_t j) {
  const float* __restrict__ src = reinterpret_cast<const float*>(e.src);
  __nv_bfloat16* __restrict__ dst = reinterpret_cast<__nv_bfloat16*>(e.dst);
  const uint32_t cin_total = (uint32_t)e.Cin_total, c_begin = (uint32_t)e.c_begin;
  if (e.mode == 0) {  // forward: dst[co][koff + tap * Cseg + c] = w[co][c_begin + c][tap]
    const uint32_t cseg = (uint32_t)e.Cseg, co = j / cseg, c = j - co * cseg;
    const float* sp = src + ((size_t)co * cin_total + c_begin + c) * TAPS;
    __nv_bfloat16* dp = dst + (int64_t)co * e.dst_row_stride + e.koff + c;
    float v[TAPS];
#pragma unroll
    for (int t = 0; t < TAPS; ++t) v[t] = sp[t];
#pragma unroll
    for (int t = 0; t < TAPS; ++t) dp[(size_t)t * cseg] = __float2bfloat16_rn(v[t]);
  } else {            // dgrad: dst[ci][tap' * Cout + co] = w[co][c_begin + ci][taps - 1 - tap']
    const uint32_t cout = (uint32_t)e.Cout, ci = j / cout, co = j - ci * cout;
    const float* sp = src + ((size_t)co * cin_total + c_begin + ci) * TAPS;
    __nv_bfloat16* dp = dst + (size_t)ci * TAPS * cout + co;
    float v[TAPS];
#pragma unroll
    for (int t = 0; t < TAPS; ++t) v[t] = sp[t];
#pragma unroll
    for (int t = 0; t < TAPS; ++t) dp[(size_t)t * cout] = __float2bfloat16_rn(v[TAPS - 1 - t]);
  }
}

__global__ void __launch_bounds__(256) weight_prepack_batch_kernel(const fm_pack_entry* __restrict__ entries,
                                                                   const int32_t* __restrict__ block_entry,
                                                                   const int64_t* __restrict__ block_offset) {
  const fm_pack_entry e = entries[block_entry[blockIdx.x]];
  // 32-bit index arithmetic (a conv weight has < 2^31 elements; 64-bit divisions were ~90 % of this kernel's time)
  const uint32_t base = (uint32_t)block_offset[blockIdx.x], pairs = (uint32_t)e.Cout * (uint32_t)e.Cseg;
#pragma unroll
  for (int k = 0; k < kPackPerBlock / 256; ++k) {
    const uint32_t j = base + k * 256 + threadIdx.x;
    if (j >= pairs) break;
    if (e.ksize == 3) pack_pair<9>(e, j); else pack_pair<1>(e, j);
  }
}

// ---- nearest 2x upsample, NHWC bf16, 16 B per thread -----------------------------------------------------------
__global__ void __launch_bounds__(256) upsample2x_kernel(const uint4* __restrict__ x, uint4* __restrict__ out, int B,
                                                        int H, int W, int C8) {
  pdl_enter();
  const int64_t total = (int64_t)B * (2 * H) * (2 * W) * C8;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int c = (int)(i % C8);
    int64_t r = i / C8;
    const int ow = (int)(r % (2 * W)); r /= (2 * W);
    const int oh = (int)(r % (2 * H));
    const int n = (int)(r / (2 * H));
    out[i] = x[(((int64_t)n * H + (oh >> 1)) * W + (ow >> 1)) * C8 + c];
  }
}

// ---- [B][R][C] -> [B][C][R] bf16 ------------------------------------------------------------------------------
__global__ void transpose_bf16_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ out, int R,
                                      int C) {
  pdl_enter();
  __shared__ __nv_bfloat16 tile[32][33];
  const int b = blockIdx.z;
  const int r0 = blockIdx.y * 32, c0 = blockIdx.x * 32;
  const __nv_bfloat16* xb = x + (size_t)b * R * C;
  __nv_bfloat16* ob = out + (size_t)b * R * C;
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const int r = r0 + j, c = c0 + threadIdx.x;
    if (r < R && c < C) tile[j][threadIdx.x] = xb[(size_t)r * C + c];
  }
  __syncthreads();
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const int c = c0 + j, r = r0 + threadIdx.x;
    if (r < R && c < C) ob[(size_t)c * R + r] = tile[threadIdx.x][j];
  }
}

static inline int ew_blocks(int64_t work_items) {
  int64_t b = (work_items + 255) / 256;
  const int64_t cap = (int64_t)sm_count() * 16;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (int)b;
}

}  // namespace fm

using namespace fm;

extern "C" int fm_sched_flowmatch_f32(float* x_out, const float* x, const float* v, const float* coef,
                                      const int32_t* step_dev, int32_t step_host, int64_t n, fm_stream_t stream) {
  if (int e = ensure_device()) return e;
  FM_REQUIRE(x_out && x && v && coef && n >= 0, "flowmatch: null pointer or negative n");
  FM_REQUIRE((((uintptr_t)x_out | (uintptr_t)x | (uintptr_t)v) & 15) == 0, "flowmatch: pointers must be 16B aligned");
  FM_REQUIRE(step_dev != nullptr || step_host >= 0, "flowmatch: negative step");
  if (n == 0) return 0;
  launch_pdl(sched_flowmatch_kernel, dim3(ew_blocks(n / 4 + 1)), dim3(256), 0, (cudaStream_t)stream, x_out, x, v, coef, step_dev,
                                                                                 step_host, n);
  FM_LAUNCH_CHECK("sched_flowmatch_kernel");
  return 0;
}

extern "C" int fm_sched_ddim_f32(float* x_out, const float* x, const float* eps, const float* coef,
                                 const int32_t* step_dev, int32_t step_host, int32_t clip, float clip_range,
                                 int64_t n, fm_stream_t stream) {
  if (int e = ensure_device()) return e;
  FM_REQUIRE(x_out && x && eps && coef && n >= 0, "ddim: null pointer or negative n");
  FM_REQUIRE((((uintptr_t)x_out | (uintptr_t)x | (uintptr_t)eps) & 15) == 0, "ddim: pointers must be 16B aligned");
  FM_REQUIRE(step_dev != nullptr || step_host >= 0, "ddim: negative step");
  if (n == 0) return 0;
  launch_pdl(sched_ddim_kernel, dim3(ew_blocks(n / 4 + 1)), dim3(256), 0, (cudaStream_t)stream, x_out, x, eps, coef, step_dev, step_host,
                                                                            clip, clip_range, n);
  FM_LAUNCH_CHECK("sched_ddim_kernel");
  return 0;
}

extern "C" int fm_sched_ddpm_f32(float* x_out, const float* x, const float* eps, const float* noise, const float* coef,
                                 const int32_t* step_dev, int32_t step_host, int32_t clip, float clip_range,
                                 int64_t n, fm_stream_t stream) {
  if (int e = ensure_device()) return e;
  FM_REQUIRE(x_out && x && eps && noise && coef && n >= 0, "ddpm: null pointer or negative n");
  FM_REQUIRE(step_dev != nullptr || step_host >= 0, "ddpm: negative step");
  if (n == 0) return 0;
  launch_pdl(sched_ddpm_kernel, dim3(ew_blocks(n)), dim3(256), 0, (cudaStream_t)stream, x_out, x, eps, noise, coef, step_dev, step_host,
                                                                    clip, clip_range, n);
  FM_LAUNCH_CHECK("sched_ddpm_kernel");
  return 0;
}

extern "C" int fm_sched_dpmpp2m_f32(float* x_out, float* m_cur, const float* x, const float* eps,
                                    const float* m_prev, const float* coef, const int32_t* step_dev,
                                    int32_t step_host, int64_t n, fm_stream_t stream) {
  if (int e = ensure_device()) return e;
  FM_REQUIRE(x_out && m_cur && x && eps && m_prev && coef && n >= 0, "dpmpp2m: null pointer or negative n");
  FM_REQUIRE((((uintptr_t)x_out | (uintptr_t)x | (uintptr_t)eps | (uintptr_t)m_cur | (uintptr_t)m_prev) & 15) == 0,
             "dpmpp2m: pointers must be 16B aligned");
  FM_REQUIRE(step_dev != nullptr || step_host >= 0, "dpmpp2m: negative step");
  if (n == 0) return 0;
  launch_pdl(sched_dpmpp_kernel, dim3(ew_blocks(n / 4 + 1)), dim3(256), 0, (cudaStream_t)stream, x_out, m_cur, x, eps, m_prev, coef,
                                                                             step_dev, step_host, n);
  FM_LAUNCH_CHECK("sched_dpmpp_kernel");
  return 0;
}

extern "C" int fm_sched_unipc_f32(float* x_out, float* last, float* m1, float* m2, const float* x, const float* eps,
                                  const float* coef, const int32_t* step_dev, int32_t step_host, int64_t n,
                                  fm_stream_t stream) {
  if (int e = ensure_device()) return e;
  FM_REQUIRE(x_out && last && m1 && m2 && x && eps && coef && n >= 0, "unipc: null pointer or negative n");
  FM_REQUIRE(step_dev != nullptr || step_host >= 0, "unipc: negative step");
  if (n == 0) return 0;
  launch_pdl(sched_unipc_kernel, dim3(ew_blocks(n)), dim3(256), 0, (cudaStream_t)stream, x_out, last, m1, m2, x, eps, coef, step_dev,
                                                                     step_host, n);
  FM_LAUNCH_CHECK("sched_unipc_kernel");
  return 0;
}

extern "C" int fm_sched_add_noise_f32(float* x_out, const float* x0, const float* noise, const float* a,
                                      const float* b, int32_t B, int64_t per_sample, fm_stream_t stream) {
  if (int e = ensure_device()) return e;
  FM_REQUIRE(x_out && x0 && noise && a && b && B >= 0 && per_sample >= 0, "add_noise: bad argument");
  const int64_t n = (int64_t)B * per_sample;
  if (n == 0) return 0;
  add_noise_kernel<<<ew_blocks(n), 256, 0, (cudaStream_t)stream>>>(x_out, x0, noise, a, b, per_sample, n);
  FM_LAUNCH_CHECK("add_noise_kernel");
  return 0;
}

extern "C" int fm_counter_add(int32_t* ctr, int32_t delta, fm_stream_t stream) {
  if (int e = ensure_device()) return e;
  FM_REQUIRE(ctr != nullptr, "counter_add: null counter");
  launch_pdl(counter_add_kernel, dim3(1), dim3(1), 0, (cudaStream_t)stream, ctr, delta);
  FM_LAUNCH_CHECK("counter_add_kernel");
  return 0;
}

extern "C" int fm_clamp_f32(float* y, const float* x, float lo, float hi, int64_t n, fm_stream_t stream) {
  if (int e = ensure_device()) return e;
  FM_REQUIRE(y && x && n >= 0, "clamp: bad argument");
  if (n == 0) return 0;
  launch_pdl(clamp_kernel, dim3(ew_blocks(n)), dim3(256), 0, (cudaStream_t)stream, y, x, lo, hi, n);
  FM_LAUNCH_CHECK("clamp_kernel");
  return 0;
}

extern "C" int fm_timestep_embedding_f32(const float* t, const float* t_table, const int32_t* step_dev, float* out,
                                         int32_t B, int32_t dim, float max_period, int32_t flip_sin_to_cos,
                                         float freq_shift, fm_stream_t stream) {
  if (int e = ensure_device()) return e;
  FM_REQUIRE(out && B > 0 && dim > 0, "timestep_embedding: bad shape B=%d dim=%d", B, dim);
  FM_REQUIRE((t != nullptr) != (t_table != nullptr), "timestep_embedding: exactly one of t / t_table must be given");
  FM_REQUIRE(t_table == nullptr || step_dev != nullptr, "timestep_embedding: t_table needs step_dev");
  const int half = dim / 2;
  float denom = (float)half - freq_shift;
  if (denom < 1.0f) denom = 1.0f;
  const float nlp = (float)(-log((double)max_period));
  const int total = B * dim;
  launch_pdl(timestep_embedding_kernel, dim3((total + 127) / 128), dim3(128), 0, (cudaStream_t)stream, t, t_table, step_dev, out, B, dim,
                                                                                   nlp, flip_sin_to_cos, denom);
  FM_LAUNCH_CHECK("timestep_embedding_kernel");
  return 0;
}

extern "C" int fm_linear_f32(const float* x, const float* W, const float* bias, const float* bias2, float* y,
                             int32_t B, int32_t I, int32_t O, int32_t silu_in, int32_t silu_out,
                             fm_stream_t stream) {
  if (int e = ensure_device()) return e;
  FM_REQUIRE(x && W && y && B > 0 && I > 0 && O > 0, "linear: bad argument B=%d I=%d O=%d", B, I, O);
  // input rows staged in shared memory: up to 192 KB (the training step's batched projections reach I ~ 20k)
  constexpr int kLinSmemMax = 192 * 1024;
  int rows = kLinSmemMax / (I * (int)sizeof(float));
  if (rows > kLinRows) rows = kLinRows;
  FM_REQUIRE(rows >= 1, "linear: in_features=%d too large", I);
  const size_t smem = (size_t)rows * I * sizeof(float);
  static bool attr = false;
  if (!attr) {
    if (int e = check_cuda(cudaFuncSetAttribute(linear_f32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                kLinSmemMax), "linear attr")) return e;
    attr = true;
  }
  launch_pdl(linear_f32_kernel, dim3((O + 3) / 4), dim3(128), smem, (cudaStream_t)stream, x, W, bias, bias2, y, B, I, O, silu_in,
                                                                      silu_out, rows);
  FM_LAUNCH_CHECK("linear_f32_kernel");
  return 0;
}

extern "C" int fm_weight_prepack_bf16(void* dst, int64_t dst_row_stride, int64_t koff, const float* src_oihw,
                                      int32_t Cout, int32_t Cin_total, int32_t c_begin, int32_t Cseg, int32_t ksize,
                                      fm_stream_t stream) {
  if (int e = ensure_device()) return e;
  FM_REQUIRE(dst && src_oihw && Cout > 0 && Cseg > 0 && c_begin >= 0 && c_begin + Cseg <= Cin_total,
             "weight_prepack: bad channel range");
  FM_REQUIRE(ksize == 1 || ksize == 3, "weight_prepack: ksize must be 1 or 3");
  const int64_t total = (int64_t)Cout * ksize * ksize * Cseg;
  weight_prepack_kernel<false><<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<__nv_bfloat16*>(dst), dst_row_stride, koff, src_oihw, Cout, Cin_total, c_begin, Cseg, ksize);
  FM_LAUNCH_CHECK("weight_prepack_kernel");
  return 0;
}

extern "C" int fm_weight_prepack_lo_bf16(void* dst, int64_t dst_row_stride, int64_t koff, const float* src_oihw,
                                         int32_t Cout, int32_t Cin_total, int32_t c_begin, int32_t Cseg, int32_t ksize,
                                         fm_stream_t stream) {
  if (int e = ensure_device()) return e;
  FM_REQUIRE(dst && src_oihw && Cout > 0 && Cseg > 0 && c_begin >= 0 && c_begin + Cseg <= Cin_total,
             "weight_prepack_lo: bad channel range");
  FM_REQUIRE(ksize == 1 || ksize == 3, "weight_prepack_lo: ksize must be 1 or 3");
  const int64_t total = (int64_t)Cout * ksize * ksize * Cseg;
  weight_prepack_kernel<true><<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<__nv_bfloat16*>(dst), dst_row_stride, koff, src_oihw, Cout, Cin_total, c_begin, Cseg, ksize);
  FM_LAUNCH_CHECK("weight_prepack_lo_kernel");
  return 0;
}

extern "C" int fm_weight_prepack_dgrad_bf16(void* dst, const float* src_oihw, int32_t Cout, int32_t Cin_total,
                                            int32_t c_begin, int32_t Cseg, int32_t ksize, fm_stream_t stream) {
  if (int e = ensure_device()) return e;
  FM_REQUIRE(dst && src_oihw && Cout > 0 && Cseg > 0 && c_begin >= 0 && c_begin + Cseg <= Cin_total,
             "weight_prepack_dgrad: bad channel range");
  FM_REQUIRE(ksize == 1 || ksize == 3, "weight_prepack_dgrad: ksize must be 1 or 3");
  const int64_t total = (int64_t)Cout * ksize * ksize * Cseg;
  weight_prepack_dgrad_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<__nv_bfloat16*>(dst), src_oihw, Cout, Cin_total, c_begin, Cseg, ksize);
  FM_LAUNCH_CHECK("weight_prepack_dgrad_kernel");
  return 0;
}

extern "C" int32_t fm_weight_prepack_batch_block_elems(void) { return kPackPerBlock; }

extern "C" int fm_weight_prepack_batch_bf16(const fm_pack_entry* entries_dev, const int32_t* block_entry_dev,
                                            const int64_t* block_offset_dev, int32_t n_blocks, fm_stream_t stream) {
  if (int e = ensure_device()) return e;
  FM_REQUIRE(entries_dev && block_entry_dev && block_offset_dev && n_blocks >= 0, "weight_prepack_batch: bad argument");
  if (n_blocks == 0) return 0;
  weight_prepack_batch_kernel<<<n_blocks, 256, 0, (cudaStream_t)stream>>>(entries_dev, block_entry_dev,
                                                                          block_offset_dev);
  FM_LAUNCH_CHECK("weight_prepack_batch_kernel");
  return 0;
}

extern "C" int fm_upsample_nearest2x_bf16(const void* x, void* out, int32_t B, int32_t H, int32_t W, int32_t C,
                                          fm_stream_t stream) {
  if (int e = ensure_device()) return e;
  FM_REQUIRE(x && out && B > 0 && H > 0 && W > 0 && C > 0 && C % 8 == 0, "upsample: bad shape (C %% 8 != 0?)");
  const int64_t total = (int64_t)B * 4 * H * W * (C / 8);
  launch_pdl(upsample2x_kernel, dim3(ew_blocks(total)), dim3(256), 0, (cudaStream_t)stream, reinterpret_cast<const uint4*>(x),
                                                                        reinterpret_cast<uint4*>(out), B, H, W, C / 8);
  FM_LAUNCH_CHECK("upsample2x_kernel");
  return 0;
}

extern "C" int fm_transpose_bf16(const void* x, void* out, int32_t B, int32_t R, int32_t C, fm_stream_t stream) {
  if (int e = ensure_device()) return e;
  FM_REQUIRE(x && out && B > 0 && R > 0 && C > 0, "transpose: bad shape");
  dim3 grid((C + 31) / 32, (R + 31) / 32, B), block(32, 8);
  launch_pdl(transpose_bf16_kernel, dim3(grid), dim3(block), 0, (cudaStream_t)stream, reinterpret_cast<const __nv_bfloat16*>(x),
                                                                 reinterpret_cast<__nv_bfloat16*>(out), R, C);
  FM_LAUNCH_CHECK("transpose_bf16_kernel");
  return 0;
}

extern "C" int fm_memset_f32(float* p, int64_t n, fm_stream_t stream) {
  if (int e = ensure_device()) return e;
  FM_REQUIRE(p != nullptr && n >= 0, "memset: bad argument");
  return check_cuda(cudaMemsetAsync(p, 0, (size_t)n * sizeof(float), (cudaStream_t)stream), "cudaMemsetAsync");
}
