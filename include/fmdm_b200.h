/*
 * fmdm_b200 — C-ABI of the B200-native sampling hot path.
 *
 * The reference (tomn681/Flow-Matching-and-Diffusion-Models) has no FFI: its "plugin API" for this path is
 * Python nn.Module classes calling ATen ops, plus the diffusers scheduler duck-type.  Every entry point below
 * replaces the ATen/diffusers call(s) made at the cited reference line(s).  The host side
 * (`flow-matching-and-diffusion-models_b200/`) mirrors the reference modules and calls these through ctypes.
 *
 * Conventions
 *   - plain pointers and sizes only; all pointers are DEVICE pointers unless the name ends in `_host`;
 *   - activations are NHWC bf16 ("channels_last" storage of a logical NCHW tensor); the sampler state x and the
 *     model prediction are fp32 NCHW;
 *   - the caller owns every buffer (PyTorch's caching allocator); the library never allocates device memory,
 *     keeps no pointer after the call returns and is safe under CUDA-graph capture;
 *   - every call is asynchronous on `stream` (a cudaStream_t passed as void*);
 *   - return value: 0 ok, <0 bad argument (see fm_last_error()), >0 a cudaError_t / CUresult.
 *   - there is no CPU path: on a machine without an sm_100 device every compute entry point returns an error.
 */
#ifndef FMDM_B200_H
#define FMDM_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* fm_stream_t;

#define FM_OK 0
#define FM_ERR_BAD_ARG (-1)
#define FM_ERR_UNSUPPORTED (-2)
#define FM_ERR_NO_DEVICE (-3)

/* library identity / diagnostics */
int fm_version(void);
const char* fm_last_error(void);
/* number of kernel launches issued by this library in this process (bench.py's gpu_launches) */
long long fm_launch_count(void);
/* Programmatic dependent launch (PDL) for the launches that follow: a kernel may become resident while its predecessor
 * in the stream still runs and waits (griddepcontrol.wait) before its first global access.  Pays on graphs of many
 * microsecond kernels (MNIST-sized problems: +6 %), costs on the large ones, so it is off unless the caller switches it
 * on - the sampling loop does while it captures the graph of a small problem.  Returns the previous setting; the
 * environment variable FMDM_PDL=0 / 1 overrides it.  Process-wide, not thread-safe (one process per GPU). */
int fm_set_pdl(int on);

/* ------------------------------------------------------------------------------------------------------------
 * K1: conv2d as implicit GEMM on tcgen05/TMEM, TMA-fed, bf16 in / fp32 accumulate / bf16 out.
 * Replaces nn.Conv2d reached through ConvND.forward (src/nn/ops/convolution.py:53-54), the 1x1 skip
 * (src/nn/blocks/residual.py:82,120), DownsampleND.op (src/nn/ops/upsampling.py:49-56), the q/k/v/out nn.Linear
 * of DiffusersAttentionND (src/nn/blocks/attention.py:216-229) seen as 1x1 convs on NHWC, the skip torch.cat
 * (src/nn/blocks/legacy_unet.py:150) as a multi-segment K loop, the temb broadcast add
 * (src/nn/blocks/residual.py:111-112) and the residual add (residual.py:120) as epilogues.
 *
 * out[n,ho,wo,co] = sum_seg sum_tap sum_c  src_seg[n, ho*stride+kh-pad, wo*stride+kw-pad, c] * W[co, k(seg,tap,c)]
 *                   + bias[co] + addvec[n,co] + residual[n,ho,wo,co]
 * W is the pre-packed K-major weight matrix [Cout][Ktot] bf16, K ordered (segment, tap=kh*3+kw, channel).
 * ---------------------------------------------------------------------------------------------------------- */
#define FM_CONV_MAX_SEG 8

typedef struct fm_conv_seg {
  const void* src;   /* bf16 NHWC [B][H][W][C]                                        */
  int32_t C;         /* channels of this segment (multiple of 8)                      */
  int32_t ksize;     /* 1 or 3 (3 => pad 1, 1 => pad 0)                               */
  int32_t upsample;  /* 1: the segment is read through a nearest-2x upsample (src is [B][H/2][W/2][C]) */
  int32_t norm_act;  /* activation of the fused operand transform: 0 none, 1 SiLU                */
  /* Fused operand transform (GroupNorm apply + activation folded into the conv's A-operand path, so the normalised
   * tensor of residual.py:95-96,113-116 is never written): when norm_a != NULL the conv reads
   *   act(norm_a[n*norm_stride + c] * src[n,h,w,c] + norm_b[n*norm_stride + c])   (zero padding applies AFTER it)
   * instead of src.  The per-(sample, channel) table comes from fm_groupnorm_affine_f32 /
   * fm_groupnorm_finalize_partials_affine.  Only for launches fm_conv_operand_norm_supported() accepts; C % 64 == 0. */
  const float* norm_a;
  const float* norm_b;
  int32_t norm_stride; /* floats between consecutive samples' rows of norm_a / norm_b */
  int32_t _pad;
} fm_conv_seg;

typedef struct fm_conv_params {
  fm_conv_seg seg[FM_CONV_MAX_SEG];
  int32_t nseg;
  int32_t B, H, W;         /* input spatial size (after the optional upsample)          */
  int32_t stride;          /* 1 or 2; output is ceil(H/stride) x ceil(W/stride)          */
  int32_t Cout;            /* multiple of 8                                              */
  int32_t _pad0;
  const void* weight;      /* bf16 [Cout][Ktot], Ktot = sum_seg ksize^2 * C               */
  const float* bias;       /* fp32 [Cout] or NULL                                        */
  const float* addvec;     /* fp32 [B][addvec_stride] per-sample channel add, or NULL    */
  int32_t addvec_stride;
  int32_t _pad1;
  const void* residual;    /* bf16 NHWC [B][Ho][Wo][Cout] or NULL                        */
  void* out;               /* bf16 NHWC [B][Ho][Wo][Cout]                                */
  float* gn_stats;         /* fp32 [B*rows][Cout/4][2] workspace (fm_conv_stats_rows) receiving, per 32-row
                              group of every M tile, the channel-quad (sum, sumsq) of `out` for the consumer
                              GroupNorm (no atomics; folded by fm_groupnorm_finalize_partials), or NULL */
  int32_t out_upsample;    /* 1: `out` is [B][2*Ho][2*Wo][Cout] and every output pixel is stored to its 2x2 block, i.e.
                              the nearest-neighbour F.interpolate(scale_factor=2) that UpsampleND applies to this
                              tensor (src/nn/ops/upsampling.py:27) is folded into the producer's store; residual
                              and gn_stats still refer to the [Ho][Wo] result (its statistics equal the upsampled
                              tensor's) */
  int32_t _pad2;
} fm_conv_params;

int fm_conv2d_igemm_bf16(const fm_conv_params* p, fm_stream_t stream);
/* Rows of fused GroupNorm partial statistics PER IMAGE the conv described by `p` will write (pointers in `p` other
 * than the norm_a/norm_b markers are ignored): gn_stats must hold B * rows * (Cout/4) * 2 floats.  The count follows
 * the kernel the launcher picks for these shapes (one row per M tile and TMEM lane quadrant, or per strip and
 * quadrant for the rolling-row kernel; images of <= 64 / <= 32 padded pixels share an M tile two / four at a time and
 * get 2 / 1 rows).  Returns FM_ERR_UNSUPPORTED when an M tile would span more than four images or Cout % 4 != 0. */
int fm_conv_stats_rows(const fm_conv_params* p, int32_t* rows_per_image);
/* Which kernel the launcher picks for `p`: 0 persistent per-tile kernel, 1 rolling-row kernel, 2 rolling-row kernel with
 * the fused operand transform; negative = error.  (Diagnostics: bench.py tags its per-kernel timings with it.) */
int fm_conv_kernel_kind(const fm_conv_params* p);
/* 1 if a conv with this input size / stride / kernel mix can take fused operand transforms (fm_conv_seg.norm_a):
 * stride 1, rows of >= 65 pixels (the M tile is 128 consecutive pixels of one image row) and a 3x3 segment. */
int fm_conv_operand_norm_supported(int32_t H, int32_t W, int32_t stride, int32_t has_3x3);

/* Re-order an OIHW fp32 conv weight (or [O][I] linear weight with ksize=1) into the K-major bf16 matrix the
 * conv kernel reads: dst[co][koff + tap*Cseg + c] = src[co][c_begin + c][kh][kw]. */
int fm_weight_prepack_bf16(void* dst, int64_t dst_row_stride, int64_t koff, const float* src_oihw, int32_t Cout,
                           int32_t Cin_total, int32_t c_begin, int32_t Cseg, int32_t ksize, fm_stream_t stream);
/* Same layout, but stores the bf16 ROUNDING RESIDUAL of each weight: dst = bf16(w - float(bf16(w))).  A conv whose K
 * axis carries the ordinary pack followed by this one over the same sources (two segments per source) computes
 * x * (w_hi + w_lo), i.e. uses ~16 mantissa bits of the fp32 master weight ("split-bf16 weights"): the precision mode
 * of the narrow (<= 128 channel) denoisers, whose per-step error otherwise sits at the bf16 noise floor. */
int fm_weight_prepack_lo_bf16(void* dst, int64_t dst_row_stride, int64_t koff, const float* src_oihw, int32_t Cout,
                              int32_t Cin_total, int32_t c_begin, int32_t Cseg, int32_t ksize, fm_stream_t stream);

/* Stem conv (tiny Cin): fp32 NCHW inputs (x and optional concatenated conditioning, src/pipelines/utils.py:204-205,
 * unet_diffusers_nd.py:148-158,173) -> bf16 NHWC.  3x3, stride 1, pad 1.  weight fp32 OIHW, bias fp32.
 * gn_stats (or NULL): fp32 [B * fm_conv_stem_stats_rows(...)][Cout/4][2] receiving the channel-quad (sum, sumsq)
 * partials of the output for the consumer GroupNorm (same format as fm_conv_params.gn_stats). */
int fm_conv_stem_f32_bf16(const float* x0, int32_t C0, const float* x1, int32_t C1, float in_scale, float in_shift,
                          const float* weight_oihw, const float* bias, void* out_nhwc_bf16, int32_t B, int32_t H,
                          int32_t W, int32_t Cout, float* gn_stats, fm_stream_t stream);
/* rows of statistics partials per image the stem kernel writes for this problem size (0 = unsupported) */
int fm_conv_stem_stats_rows(int32_t B, int32_t H, int32_t W, int32_t Cout);
/* Stem on the tensor cores in ONE launch (Cout 64 | 128, 9*Cin + 2 <= 32, i.e. Cin <= 3): same arguments and result
 * layout as fm_conv_stem_f32_bf16, inputs and weights rounded to bf16 (fp32 accumulation, bias exact to 2^-17 as two
 * bf16 k-columns), `mma.sync` fragments gathered from a shared-memory halo tile, no im2col tensor.  gn_stats (or NULL):
 * fp32 [B * fm_conv_stem_tc_stats_rows(...)][Cout/4][2].  FM_ERR_UNSUPPORTED outside the covered shapes.  Replaces
 * conv_in of src/models/unet/unet_diffusers_nd.py:148-158,173 on large inputs. */
int fm_conv_stem_tc_f32_bf16(const float* x0, int32_t C0, const float* x1, int32_t C1, float in_scale, float in_shift,
                             const float* weight_oihw, const float* bias, void* out_nhwc_bf16, int32_t B, int32_t H,
                             int32_t W, int32_t Cout, float* gn_stats, fm_stream_t stream);
/* rows of statistics partials per image fm_conv_stem_tc_f32_bf16 writes (0 = shape not covered by that kernel) */
int fm_conv_stem_tc_stats_rows(int32_t B, int32_t H, int32_t W, int32_t Cin, int32_t Cout);
/* Stem on the tensor cores, step 1 (large inputs): the 3x3 neighbourhoods of the fp32 NCHW inputs (same x0 / x1 /
 * in_scale / in_shift meaning as fm_conv_stem_f32_bf16) as a bf16 NHWC tensor out[B][H][W][Kp], column ci*9 + kh*3 + kw
 * (zero beyond 9*Cin; Kp a multiple of 8, <= 72).  Step 2 is fm_conv2d_igemm_bf16 as a 1x1 conv over it with the
 * weight matrix w.reshape(Cout, 9*Cin) zero-padded to Kp columns: conv_in runs at store bandwidth. */
int fm_stem_im2col_bf16(const float* x0, int32_t C0, const float* x1, int32_t C1, float in_scale, float in_shift,
                        void* out, int32_t B, int32_t H, int32_t W, int32_t Kp, fm_stream_t stream);

/* Head conv (tiny Cout): bf16 NHWC -> fp32 NCHW (unet_diffusers_nd.py:190, unet.py:288-292). 3x3 s1 p1.
 * norm_ab != NULL ([B][2][Cin], fm_groupnorm_affine_f32): the input is read as SiLU(a*x+b) (norm_act=1) or a*x+b,
 * i.e. conv_norm_out + SiLU (unet_diffusers_nd.py:188-189) folded into the load. */
int fm_conv_head_bf16_f32(const void* x_nhwc_bf16, const float* weight_oihw, const float* bias, float* out_nchw,
                          int32_t B, int32_t H, int32_t W, int32_t Cin, int32_t Cout, const float* norm_ab,
                          int32_t norm_act, fm_stream_t stream);

/* ------------------------------------------------------------------------------------------------------------
 * K2: GroupNorm (+SiLU, +scale-shift) on NHWC bf16, fp32 statistics.
 * Replaces nn.GroupNorm + nn.SiLU (src/nn/ops/normalization.py:11-19; residual.py:95-96,113-116;
 * unet_diffusers_nd.py:188-189; attention.py:235-236).  Up to two sources are read as a virtual channel concat.
 * ---------------------------------------------------------------------------------------------------------- */
/* stats[n][g] = (mean, rstd) of the group.  Deterministic two-stage reduction (no atomics): per-block partial sums go
 * to `workspace` (fm_groupnorm_workspace_elems(B, HW, C0+C1, groups) floats), a second tiny kernel folds them in
 * fp64 in a fixed order. */
int64_t fm_groupnorm_workspace_elems(int32_t B, int64_t HW, int32_t C, int32_t groups);
int fm_groupnorm_stats_bf16(const void* x0, int32_t C0, const void* x1, int32_t C1, int32_t B, int64_t HW,
                            int32_t groups, float eps, float* workspace, float* stats, fm_stream_t stream);
/* Same statistics from the channel-quad partial sums written by the producing convs' epilogues (virtual concat of
 * two producers): p_s is [B*rows_s][C_s/4][2]. */
int fm_groupnorm_finalize_partials(const float* p0, int32_t rows0, int32_t C0, const float* p1, int32_t rows1,
                                   int32_t C1, int32_t B, int64_t HW, int32_t groups, float eps, float* stats,
                                   fm_stream_t stream);
/* Per-(sample, channel) affine form of the normalisation, ab[n][0][c] = a, ab[n][1][c] = b with
 *   a*x + b == ((x-mean)*rstd*gamma[c]+beta[c]) * (1+scale[n,c]) + shift[n,c]
 * (fp32 [B][2][C]); consumed by the conv operand transform (fm_conv_seg.norm_a/norm_b) and by fm_conv_head_bf16_f32. */
int fm_groupnorm_affine_f32(const float* stats, const float* gamma, const float* beta, const float* scale_shift,
                            int64_t ss_stride, int32_t B, int32_t C, int32_t groups, float* ab, fm_stream_t stream);
/* fm_groupnorm_finalize_partials + fm_groupnorm_affine_f32 in one launch (stats may be NULL). */
int fm_groupnorm_finalize_partials_affine(const float* p0, int32_t rows0, int32_t C0, const float* p1, int32_t rows1,
                                          int32_t C1, int32_t B, int64_t HW, int32_t groups, float eps,
                                          const float* gamma, const float* beta, const float* scale_shift,
                                          int64_t ss_stride, float* stats, float* ab, fm_stream_t stream);
/* y = act( ((x-mean)*rstd*gamma+beta) * (1+scale[n,c]) + shift[n,c] ), act = SiLU if silu!=0.
 * scale_shift: fp32 rows of 2*C (scale first, then shift; residual.py:109,115), row n at scale_shift + n*ss_stride,
 * or NULL. */
int fm_groupnorm_apply_bf16(const void* x0, int32_t C0, const void* x1, int32_t C1, int32_t B, int64_t HW,
                            int32_t groups, const float* stats, const float* gamma, const float* beta,
                            const float* scale_shift, int64_t ss_stride, int32_t silu, void* out, fm_stream_t stream);
/* The same apply with the statistics folded inside the kernel from the producer convs' channel-quad partials (p0 / p1,
 * rows per image rows0 / rows1: the fm_conv_params.gn_stats format) - no finalize launch, no statistics tensor.  Only
 * while the per-sample table is a few KB and the groups are made of whole quads
 * (fm_groupnorm_apply_partials_supported != 0): the one-kernel form of nn.GroupNorm (+SiLU) for the small problems
 * whose cost is the number of kernel nodes (src/nn/blocks/residual.py:95-96,113-116 on 14x14 / 7x7 feature maps). */
int fm_groupnorm_apply_partials_supported(int32_t rows0, int32_t C0, int32_t rows1, int32_t C1, int32_t groups);
int fm_groupnorm_apply_partials_bf16(const void* x0, int32_t C0, const void* x1, int32_t C1, int32_t B, int64_t HW,
                                     int32_t groups, const float* p0, int32_t rows0, const float* p1, int32_t rows1,
                                     float eps, const float* gamma, const float* beta, const float* scale_shift,
                                     int64_t ss_stride, int32_t silu, void* out, fm_stream_t stream);
int fm_memset_f32(float* p, int64_t n, fm_stream_t stream);

/* nearest-neighbour 2x upsample on NHWC bf16 (F.interpolate, src/nn/ops/upsampling.py:27) */
int fm_upsample_nearest2x_bf16(const void* x, void* out, int32_t B, int32_t H, int32_t W, int32_t C,
                               fm_stream_t stream);
/* [B][R][C] -> [B][C][R] bf16 transpose (SpatialSelfAttention raw-reshape support, attention.py:111-115) */
int fm_transpose_bf16(const void* x, void* out, int32_t B, int32_t R, int32_t C, fm_stream_t stream);

/* ------------------------------------------------------------------------------------------------------------
 * K3: softmax(QK^T/sqrt(d))V, bf16 in/out, fp32 softmax.  Replaces F.scaled_dot_product_attention
 * (src/nn/blocks/attention.py:41-44).  Generic strides (in elements) so both the DiffusersAttentionND layout
 * ([B][T][3C], head-major channels) and SpatialSelfAttention's raw reshape ([b][heads][T][3*dh]) are served.
 * head_dim in {8,16,32,64}.
 * ---------------------------------------------------------------------------------------------------------- */
int fm_attention_bf16(const void* q, const void* k, const void* v, void* out, int32_t B, int32_t heads, int32_t Tq,
                      int32_t Tk, int32_t head_dim, int64_t q_sb, int64_t q_sh, int64_t q_st, int64_t kv_sb,
                      int64_t kv_sh, int64_t kv_st, int64_t o_sb, int64_t o_sh, int64_t o_st, float scale,
                      fm_stream_t stream);

/* ------------------------------------------------------------------------------------------------------------
 * Time embedding path (fp32).  timestep_embedding (src/nn/ops/time_embedding.py:4-32) and the small Linear layers
 * (models/unet/utils.py:9-24, residual.py:99-108).
 * ---------------------------------------------------------------------------------------------------------- */
/* t: fp32 [B] (or, if t_table != NULL, every row uses t_table[*step_dev]); out fp32 [B][dim] */
int fm_timestep_embedding_f32(const float* t, const float* t_table, const int32_t* step_dev, float* out, int32_t B,
                              int32_t dim, float max_period, int32_t flip_sin_to_cos, float freq_shift,
                              fm_stream_t stream);
/* y[b][o] = bias[o] + bias2[o] + sum_i f(x[b][i]) * W[o][i]; f = SiLU if silu_in; y = SiLU(y) if silu_out */
int fm_linear_f32(const float* x, const float* W, const float* bias, const float* bias2, float* y, int32_t B,
                  int32_t I, int32_t O, int32_t silu_in, int32_t silu_out, fm_stream_t stream);

/* Linear attention (attention.py:53-70 LinearQKVAttention): out = softmax_features(q) (softmax_tokens(k)^T v / (sum_tokens
 * softmax_tokens(k) + eps)); same strided bf16 operands as fm_attention_bf16, no 1/sqrt(d) scaling. */
int fm_linear_attention_bf16(const void* q, const void* k, const void* v, void* out, int32_t B, int32_t heads, int32_t Tq,
                             int32_t Tk, int32_t head_dim, int64_t q_sb, int64_t q_sh, int64_t q_st, int64_t kv_sb,
                             int64_t kv_sh, int64_t kv_st, int64_t o_sb, int64_t o_sh, int64_t o_st, float eps,
                             fm_stream_t stream);
/* Cross-attention context path (attention.py:149-165, 232-262): GroupNorm over the context tokens and
 * the key/value projection in one pass.  ctx fp32 [B][Cc][Tc] (Cc <= 16), gamma/beta fp32 [Cc], W fp32 [O][Cc] (the
 * concatenated to_k|to_v or kv_proj weight), bias fp32 [O] or NULL, stats_ws fp32 [B][groups][2].
 * out bf16: [B][Tc][O] (channel_major = 0) or [B][O][Tc] (channel_major = 1, SpatialCrossAttention's raw reshape). */
int fm_context_kv_bf16(const float* ctx, const float* gamma, const float* beta, const float* W, const float* bias,
                       float* stats_ws, void* out, int32_t B, int32_t Cc, int32_t Tc, int32_t O, int32_t groups,
                       float eps, int32_t channel_major, fm_stream_t stream);

/* ------------------------------------------------------------------------------------------------------------
 * K4: one fused elementwise kernel per sampler step (fp32 state, round-to-nearest mul/add, no FMA contraction).
 * Replaces diffusers' Scheduler.step called at src/pipelines/utils.py:218.  Coefficients are precomputed on the
 * host exactly as diffusers computes its 0-dim fp32 scalars and live in a device table [nsteps][ncoef];
 * the row is `step_host`, or `*step_dev` when step_dev != NULL (CUDA-graph replay).
 * ---------------------------------------------------------------------------------------------------------- */
#define FM_FLOWMATCH_NCOEF 1 /* {dt = sigma[i+1]-sigma[i]} */
int fm_sched_flowmatch_f32(float* x_out, const float* x, const float* v, const float* coef, const int32_t* step_dev,
                           int32_t step_host, int64_t n, fm_stream_t stream);
#define FM_DDIM_NCOEF 4 /* {sqrt(1-a_t), sqrt(a_t), sqrt(a_prev), sqrt(1-a_prev-std^2)} */
int fm_sched_ddim_f32(float* x_out, const float* x, const float* eps, const float* coef, const int32_t* step_dev,
                      int32_t step_host, int32_t clip, float clip_range, int64_t n, fm_stream_t stream);
#define FM_DDPM_NCOEF 8 /* {sqrt(1-a_t), sqrt(a_t), c_x0, c_xt, sigma_t (0 at t = 0), unused x3} */
/* DDPM ancestral step (the reference's default `ddpm` scheduler, diffusers DDPMScheduler.step, epsilon prediction,
 * "fixed_small" variance): x_out = c_x0*clamp(x0) + c_xt*x + sigma*noise; noise: fp32 standard normal, same shape */
int fm_sched_ddpm_f32(float* x_out, const float* x, const float* eps, const float* noise, const float* coef,
                      const int32_t* step_dev, int32_t step_host, int32_t clip, float clip_range, int64_t n,
                      fm_stream_t stream);
#define FM_DPMPP_NCOEF 8 /* {sigma_s, alpha_s, c1, c2, c3, inv_r0, second_order, raw_epsilon} */
/* DPMSolverMultistepScheduler.step, order <= 2, midpoint.  raw_epsilon = 0: algorithm_type "dpmsolver++" (data
 * prediction m = (x - sigma_s e)/alpha_s); raw_epsilon != 0: algorithm_type "dpmsolver" (m = e; c1 = alpha_t/alpha_s,
 * c2 = sigma_t (exp(h) - 1)).  x_out = c1 x - c2 m [- c3 (inv_r0 (m - m_prev))]; m_cur <- m. */
int fm_sched_dpmpp2m_f32(float* x_out, float* m_cur, const float* x, const float* eps, const float* m_prev,
                         const float* coef, const int32_t* step_dev, int32_t step_host, int64_t n,
                         fm_stream_t stream);
#define FM_UNIPC_NCOEF 16
/* UniPCMultistepScheduler.step (bh2, predict_x0, order <= 2): ONE kernel = data-prediction conversion + UniC corrector
 * of the incoming sample + UniP predictor + history shift.  Row: {sigma, alpha, use_corrector, ca, cb, cc,
 * corrector_order2, rk_c, rho_0, rho_last, pa, pb, pc, predictor_order2, rk_p, rho_p (0.5)}; state tensors (same shape
 * as x, updated in place): last = corrected sample of the previous step, m1 / m2 = the last two data predictions. */
int fm_sched_unipc_f32(float* x_out, float* last, float* m1, float* m2, const float* x, const float* eps,
                       const float* coef, const int32_t* step_dev, int32_t step_host, int64_t n, fm_stream_t stream);
/* x_out = a[n]*x0 + b[n]*noise (scheduler.add_noise, src/utils/model_utils/diffusion_utils.py:222) */
int fm_sched_add_noise_f32(float* x_out, const float* x0, const float* noise, const float* a, const float* b,
                           int32_t B, int64_t per_sample, fm_stream_t stream);
/* *ctr += delta (device-side step counter for graph replay) */
int fm_counter_add(int32_t* ctr, int32_t delta, fm_stream_t stream);
/* y = clamp(x, lo, hi) fp32 (samplers/diffusion_like.py:135) */
int fm_clamp_f32(float* y, const float* x, float lo, float hi, int64_t n, fm_stream_t stream);

/* ------------------------------------------------------------------------------------------------------------
 * Training step: the backward of the ops above plus the loss and optimiser kernels, i.e. what
 * torch.autograd executes for src/pipelines/train/flow_matching_lib.py:138-182 (F.conv2d / F.group_norm / SiLU /
 * SDPA / F.linear / F.mse_loss backward, torch.optim.AdamW.step).  All reductions are two-stage with a fixed
 * summation order (deterministic, no atomics).
 * ---------------------------------------------------------------------------------------------------------- */
/* K-major bf16 weights of the data-gradient conv of the channel slice [c_begin, c_begin+Cseg) of an OIHW fp32
 * weight: dst [Cseg][k*k*Cout], in/out channels swapped and taps mirrored (dX = conv(dY, dst) for stride 1) */
int fm_weight_prepack_dgrad_bf16(void* dst, const float* src_oihw, int32_t Cout, int32_t Cin_total, int32_t c_begin,
                                 int32_t Cseg, int32_t ksize, fm_stream_t stream);
/* One launch for every weight pack of a training step.  An entry is one K segment of one packed matrix: mode 0 =
 * fm_weight_prepack_bf16 arguments, mode 1 = fm_weight_prepack_dgrad_bf16 arguments (dst_row_stride / koff unused).
 * The host pre-assigns blocks: block b packs the (matrix row, channel) pairs [block_offset[b], block_offset[b] +
 * fm_weight_prepack_batch_block_elems()) of entry block_entry[b] (pair count = Cout * Cseg; a pair = its ksize^2
 * taps, contiguous in the OIHW source). */
typedef struct fm_pack_entry {
  const void* src;          /* fp32 OIHW (or [O][I]) master weight */
  void* dst;                /* bf16 packed matrix base */
  int64_t dst_row_stride;
  int64_t koff;
  int32_t Cout, Cin_total, c_begin, Cseg, ksize, mode;
} fm_pack_entry;
int32_t fm_weight_prepack_batch_block_elems(void);
int fm_weight_prepack_batch_bf16(const fm_pack_entry* entries_dev, const int32_t* block_entry_dev,
                                 const int64_t* block_offset_dev, int32_t n_blocks, fm_stream_t stream);
/* conv wgrad: dw[co][c_begin+ci][kh][kw] = sum_{b,yo,xo} dy[b][yo][xo][co] * x[b][yo*stride+kh-pad][xo*stride+kw-pad][ci]
 * dy bf16 NHWC [B][Ho][Wo][Cout], x bf16 NHWC [B][H][W][Cin], dw fp32 OIHW [Cout][cin_total][k][k] (the slice
 * [c_begin, c_begin+Cin) is written).  ksize 1 or 3 (pad = ksize/2), stride 1 or 2.
 * workspace: fm_conv_wgrad_workspace_elems(...) floats (split-K partials). */
int64_t fm_conv_wgrad_workspace_elems(int32_t B, int32_t Ho, int32_t Wo, int32_t Cin, int32_t Cout, int32_t ksize);
int fm_conv_wgrad_bf16(const void* dy, const void* x, float* dw, float* workspace, int32_t B, int32_t H, int32_t W,
                       int32_t Cin, int32_t Cout, int32_t ksize, int32_t stride, int32_t cin_total, int32_t c_begin,
                       fm_stream_t stream);
/* out[b][c] = sum_p dy[b][p][c] (bf16 [B][HW][C] -> fp32 [B][C]; the time-embedding add gradient); total (or NULL)
 * [c] = sum_b out[b][c] (the bias gradient).  workspace: fm_colsum_workspace_elems floats. */
int64_t fm_colsum_workspace_elems(int32_t B, int64_t HW, int32_t C);
int fm_colsum_bf16(const void* dy, float* workspace, float* out, float* total, int32_t B, int64_t HW, int32_t C,
                   int32_t* tickets, fm_stream_t stream);
/* tickets: a caller-owned device buffer of fm_ticket_ints() int32, ZERO before its first use and used by one stream at
 * a time; kernels that end with a "the last block finishes the job" step count their blocks in it and leave it zero
 * again (the counters only decide which block runs the fixed-order fold, never the order of a sum). */
int32_t fm_ticket_ints(void);
/* out[b][2y][2x][c] = x[b][y][x][c], zero elsewhere: turns the dgrad of a stride-2 conv into a stride-1 conv */
int fm_zero_insert2x_bf16(const void* x, void* out, int32_t B, int32_t H, int32_t W, int32_t C, fm_stream_t stream);
/* out[b][y][x][c] = sum of the 2x2 block of x [B][2H][2W][C]: backward of fm_upsample_nearest2x_bf16 */
int fm_sumpool2x2_bf16(const void* x, void* out, int32_t B, int32_t H, int32_t W, int32_t C, fm_stream_t stream);
/* Backward of fm_groupnorm_apply_bf16 over the virtual channel concat of (x0 [C0], x1 [C1] or NULL/0), bf16
 * [B][HW][C_s]; dout bf16 [B][HW][C0+C1]; stats as the forward computed them.  dx0 / dx1 bf16 like x0 / x1;
 * dgamma_dbeta fp32 [2][C]; dscale_shift fp32 [B][2C] (NULL iff scale_shift is NULL).
 * Two launches: the partial-sum pass (whose last block per sample folds the sums, forms the group totals, the
 * coefficient table and the per-sample parameter gradients) and the apply pass (one block of which folds dgamma /
 * dbeta over the batch).
 * add0_a, add0_b (or NULL): bf16 tensors shaped like x0 that are ADDED to dx0 - the gradients other consumers of x0
 * (a residual connection, a skip connection) produced, so no separate accumulation pass runs; add1: same for dx1.
 * dbeta (or NULL): when given, dgamma_dbeta receives only the [C] dgamma row and dbeta the [C] dbeta row (two separate
 * destinations, e.g. the parameters' slices of a flat gradient buffer).
 * dx_colsum_partials (or NULL): fp32 [B][fm_groupnorm_bwd_blocks(B, HW)][C0+C1] per-block column sums of (dx0 | dx1)
 * incl. the added tensors, i.e. the first stage of the bias / time-embedding-add gradient of the conv that produced
 * x; fold with fm_colsum_finish_f32 (ld = C0+C1).
 * workspace: fm_groupnorm_bwd_workspace_elems(B, HW, C0+C1) floats; tickets: see fm_ticket_ints. */
int64_t fm_groupnorm_bwd_workspace_elems(int32_t B, int64_t HW, int32_t C);
int fm_groupnorm_bwd_bf16(const void* x0, int32_t C0, const void* x1, int32_t C1, const void* dout, const float* stats,
                          const float* gamma, const float* beta, const float* scale_shift, int64_t ss_stride,
                          int32_t silu, int32_t B, int64_t HW, int32_t groups, float* workspace, void* dx0, void* dx1,
                          float* dgamma_dbeta, float* dscale_shift, float* dx_colsum_partials, float* dbeta,
                          const void* add0_a, const void* add0_b, const void* add1, int32_t* tickets,
                          fm_stream_t stream);
int32_t fm_groupnorm_bwd_blocks(int32_t B, int64_t HW);
/* out[b][c] = sum_blk partials[(b*nblk + blk)*ld + c] (fp32 [B][C]); total (or NULL) [c] = sum_b out[b][c] */
int fm_colsum_finish_f32(const float* partials, float* out, float* total, int32_t B, int32_t nblk, int32_t C,
                         int32_t ld, int32_t* tickets, fm_stream_t stream);
/* Backward of fm_attention_bf16 (self-attention, tq == tk == T; q/k/v share the strides qs_*, o/dout share os_*;
 * dq/dk/dv are written with the q strides).  Strides in elements. */
int fm_attention_bwd_bf16(const void* q, const void* k, const void* v, const void* o, const void* dout, void* dq,
                          void* dk, void* dv, int32_t B, int32_t heads, int32_t T, int32_t head_dim, int64_t qs_b,
                          int64_t qs_h, int64_t qs_t, int64_t os_b, int64_t os_h, int64_t os_t, float scale,
                          fm_stream_t stream);
/* Attention backward with queries and keys of their own length / layout (cross-attention: attention.py:149-189,
 * 232-274; fm_attention_bwd_bf16 is the Tq == Tk case): q, dq use the q strides; k, v, dk, dv the kv strides; o, dout
 * the o strides (elements).  head_dim 8 runs on mma.sync, 16 / 32 / 64 on the CUDA cores. */
int fm_attention_bwd_cross_bf16(const void* q, const void* k, const void* v, const void* o, const void* dout, void* dq,
                                void* dk, void* dv, int32_t B, int32_t heads, int32_t Tq, int32_t Tk, int32_t head_dim,
                                int64_t qs_b, int64_t qs_h, int64_t qs_t, int64_t ks_b, int64_t ks_h, int64_t ks_t,
                                int64_t os_b, int64_t os_h, int64_t os_t, float scale, fm_stream_t stream);
/* Backward of fm_linear_attention_bf16 (attention.py:53-70): same strided bf16 operands; dq with the q strides, dk / dv
 * with the kv strides, dout with the o strides. */
int fm_linear_attention_bwd_bf16(const void* q, const void* k, const void* v, const void* dout, void* dq, void* dk,
                                 void* dv, int32_t B, int32_t heads, int32_t Tq, int32_t Tk, int32_t head_dim,
                                 int64_t q_sb, int64_t q_sh, int64_t q_st, int64_t kv_sb, int64_t kv_sh, int64_t kv_st,
                                 int64_t o_sb, int64_t o_sh, int64_t o_st, float eps, fm_stream_t stream);
/* Backward of fm_context_kv_bf16 (token-major output): dkv bf16 [B][Tc][O]; stats = the (mean, rstd) pairs the forward
 * wrote; dW fp32 [O][Cc], dbias fp32 [O] (or NULL), dgamma / dbeta fp32 [Cc] of the context GroupNorm.
 * workspace: fm_context_kv_bwd_workspace_elems floats. */
int64_t fm_context_kv_bwd_workspace_elems(int32_t B, int32_t Cc, int32_t Tc, int32_t O);
int fm_context_kv_bwd_f32(const float* ctx, const float* stats, const float* gamma, const float* beta, const float* W,
                          const void* dkv, float* workspace, float* dW, float* dbias, float* dgamma, float* dbeta,
                          int32_t B, int32_t Cc, int32_t Tc, int32_t O, int32_t groups, fm_stream_t stream);
/* Backward of fm_linear_f32 (y = f(x) W^T + b, f = SiLU if silu_in) for B <= 32 rows: dx fp32 [B][I] (or NULL), dw fp32
 * [O][I], db fp32 [O] (or NULL).  workspace: fm_linear_bwd_workspace_elems floats (needed for dx only). */
int64_t fm_linear_bwd_workspace_elems(int32_t B, int32_t I, int32_t O);
int fm_linear_bwd_f32(const float* x, const float* W, const float* dy, float* workspace, float* dx, float* dw, float* db,
                      int32_t B, int32_t I, int32_t O, int32_t silu_in, fm_stream_t stream);
/* dx = dy * SiLU'(x), fp32 */
int fm_silu_bwd_f32(const float* x, const float* dy, float* dx, int64_t n, fm_stream_t stream);
/* Stem conv wgrad (inputs as fm_conv_stem_f32_bf16; 1..4 input channels): dw fp32 [Cout][C0+C1][3][3] */
int64_t fm_conv_stem_wgrad_workspace_elems(int32_t Cin, int32_t Cout);
int fm_conv_stem_wgrad_f32(const float* x0, int32_t C0, const float* x1, int32_t C1, float in_scale, float in_shift,
                           const void* dy_nhwc_bf16, float* workspace, float* dw, int32_t B, int32_t H, int32_t W,
                           int32_t Cout, fm_stream_t stream);
/* Head conv (Cout == 1) backward: a bf16 NHWC [B][H][W][Cin] (the conv input), dy fp32 [B][1][H][W];
 * da bf16 NHWC (or NULL), dw fp32 [1][Cin][3][3] */
int64_t fm_conv_head_bwd_workspace_elems(int32_t Cin);
int fm_conv_head_bwd_f32(const void* a, const float* dy, const float* weight, float* workspace, void* da, float* dw,
                         int32_t B, int32_t H, int32_t W, int32_t Cin, fm_stream_t stream);
/* out[0] = scale * sum_i x[i] (mode 0) or scale * sum_i (x[i] - (t1[i] - t2[i]))^2 (mode 1, t2 may be NULL): the
 * fused velocity-target MSE of flow_matching_lib.py:163-164.  workspace: 1024 doubles. */
int fm_sum_f32(const float* x, const float* t1, const float* t2, double* workspace, float* out, int64_t n, int32_t mode,
               double scale, fm_stream_t stream);
/* dpred = (2/n) * gscalar[0] * (pred - (t1 - t2)) */
int fm_mse_bwd_f32(const float* pred, const float* t1, const float* t2, const float* gscalar, float* dpred, int64_t n,
                   fm_stream_t stream);
/* torch.optim.AdamW.step over one flat fp32 parameter buffer (g is multiplied by grad_scale first) */
int fm_adamw_f32(float* p, const float* g, float* m, float* v, int64_t n, float lr, float beta1, float beta2, float eps,
                 float weight_decay, int64_t step, float grad_scale, fm_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* FMDM_B200_H */
