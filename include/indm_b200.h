/* indm_b200 — C ABI of the B200-native (sm_100a) INDM hot-path library `libindm_b200.so`.
 *
 * Every entry point: plain pointers and sizes, DEVICE pointers unless stated otherwise, `stream` is a cudaStream_t
 * passed as void*, the call enqueues work on that stream and returns immediately.  Return value 0 = success,
 * non-zero = error (message via indm_last_error()).  Nothing allocates device memory: the caller owns inputs,
 * outputs and workspaces.  The library is re-entrant and device-local (the reference's ops are invoked from one
 * Python thread per GPU under nn.DataParallel; here there is one process per GPU).  There is no CPU path.
 *
 * What each entry point stands in for in the reference (byeonghu-na/INDM) is cited as file:line.
 */
#ifndef INDM_B200_H_
#define INDM_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define INDM_DTYPE_BF16 0 /* bf16 operands, fp32 accumulate (production mode)      */
#define INDM_DTYPE_TF32 1 /* fp32 storage, error-compensated tf32 tensor-core math, 3 MMAs per product (validation mode) */
#define INDM_DTYPE_F32 2  /* plain fp32 (only where stated)                        */

const char* indm_version(void);
/* last error message of the calling thread ("" if none) */
const char* indm_last_error(void);

/* ------------------------------------------------------------------------------------------------------------------
 * The reference's two native operators (pybind modules `upfirdn2d`, `fused`)
 * ---------------------------------------------------------------------------------------------------------------- */

/* upfirdn2d(Tensor input[major,in_h,in_w,minor=1], Tensor kernel[kh,kw], up_x, up_y, down_x, down_y,
 *           pad_x0, pad_x1, pad_y0, pad_y1) -> [major,out_h,out_w]        (op/upfirdn2d.cpp:12-19,
 * op/upfirdn2d_kernel.cu:209-369).  FP32.  out_h = (in_h*up_y + pad_y0 + pad_y1 - kh)/down_y + 1 (same for w);
 * the caller allocates y with that extent.  The backward pass is this same call with up<->down, the flipped
 * kernel and the g_pad values of op/upfirdn2d.py:111-114. */
int indm_upfirdn2d_f32(const float* x, const float* k, float* y, int64_t major, int in_h, int in_w, int kh, int kw,
                       int up_x, int up_y, int down_x, int down_y, int pad_x0, int pad_x1, int pad_y0, int pad_y1,
                       void* stream);

/* fused_bias_act(input, bias, refer, act, grad, alpha, scale) (op/fused_bias_act.cpp:11-17,
 * op/fused_bias_act_kernel.cu:19-99).  y[i] = f(x[i] + bias[(i / step_b) % size_b]) * scale with
 * act: 1 linear, 3 leaky-ReLU(alpha); grad: 0 forward, 1 first derivative (sign taken from `ref`), 2 -> 0.
 * bias / ref may be NULL (= the reference's empty tensors). */
int indm_bias_act_f32(const float* x, const float* bias, const float* ref, float* y, int64_t n, int size_b, int64_t step_b,
                      int act, int grad, float alpha, float scale, void* stream);

/* ------------------------------------------------------------------------------------------------------------------
 * Tensor-core implicit GEMM (tcgen05 + TMEM + TMA).  Stands in for the cuDNN / cuBLAS calls behind
 * nn.Conv2d 3x3 / 1x1 (models/layers.py:100-124), NIN (models/layers.py:546-555), the attention einsums
 * (models/layerspp.py:95,99) and the flow's 1x1 conv (flows/resflow/layers/base/lipschitz.py:434).
 * ---------------------------------------------------------------------------------------------------------------- */
typedef struct indm_igemm {
  int32_t dtype; /* INDM_DTYPE_BF16 or INDM_DTYPE_TF32: element type of a/b/a2/b2 (bf16 or fp32) */
  /* A: activations, NHWC: element (n,y,x,c) at a[n*a_img_stride + (y*W + x)*a_ld + c]; 0 strides = dense.
   * Plain GEMM [M,K]: N=1, H=1, W=M, Cin=K. */
  const void* a;
  int64_t a_ld, a_img_stride;
  int32_t N, H, W, Cin;
  /* B: weights [taps][Cout][Cin] (Cin contiguous): element (t,o,c) at b[t*b_tap_stride + o*b_ld + c].
   * taps = 9: 3x3 / stride 1 / zero pad 1 cross-correlation, t = ky*3 + kx;  taps = 1: 1x1.
   * batched_b = 1 (taps must be 1): one B matrix per image n, at b[n*b_tap_stride + ...] (attention). */
  const void* b;
  int64_t b_ld, b_tap_stride;
  int32_t Cout, taps, batched_b;
  /* optional second K segment accumulated into the same output: 1x1 conv of a2 (NHWC, Cin2 channels) with b2 [Cout][Cin2]
   * (fused res-block skip conv, models/layerspp.py:281-282).  a2 = NULL disables it. */
  const void* a2;
  int64_t a2_ld;
  int32_t Cin2;
  const void* b2;
  int64_t b2_ld;
  /* epilogue: v = act((acc + bias[c] + rowbias[n*rowbias_ld + c]) * scale * rowscale[n] + residual[pixel*res_ld + c] * res_scale) */
  const float* bias;     /* [Cout] or NULL */
  const float* rowbias;  /* per-image bias (time-embedding Dense_0 output, models/layerspp.py:276) or NULL */
  int64_t rowbias_ld;
  const float* residual; /* fp32 NHWC (NCHW when out_mode == 1) or NULL */
  int64_t res_ld;
  float res_scale;
  int32_t act;           /* 0 none; 1 Sin(x) = sin(2 pi x)/(2 pi), the resflow activation (flows/resflow/layers/base/activations.py:7-12);
                            2 ELU (posterior encoder, nnet/resnets/resnet_batchnorm.py:26-27) */
  const float* rowscale; /* [N] or NULL */
  float scale;
  int32_t out_mode;      /* 0: NHWC rows out_*[pixel*out_ld + c];  1: NCHW fp32 (out_f32);
                            2: as 0 for c < tcol0, and c >= tcol0 transposed per image into out_t[n][c - tcol0][y*W + x] (bf16) */
  float* out_f32;        /* either / both may be given in mode 0 */
  void* out_bf16;
  int64_t out_ld;
  int32_t tcol0;
  void* out_t;
  int32_t round_tf32_out; /* ignored (kept for ABI stability): TF32-mode operands are plain fp32, split hi/lo in-kernel */
  /* optional fused GroupNorm statistics of the stored output: gn_partial[n][g][2] += (sum, sum of squares) over
   * the tile (atomicAdd; caller zeroes it).  gn_cpg = channels per group. */
  float* gn_partial;
  int32_t gn_cpg, gn_groups;
  int32_t block_n;       /* 0 = choose automatically; else 32 / 64 / 128 / 256 */
  int32_t stride;        /* 0 / 1: unit stride.  2 (taps == 9): 3x3 stride-2 VALID convolution — A is the [N, 2H+1, 2W+1, Cin]
                            FIR-padded image and H x W the output grid (conv_downsample_2d, models/up_or_down_sampling.py:173-178) */
  int32_t a_H, a_W, pad; /* stride 2 only: extent of the A image grid (0 = the default 2H+1 x 2W+1) and zero padding 0 / 1, i.e.
                            input pixel (2y + ky - pad, 2x + kx - pad); taps == 1 gives a strided 1x1 conv
                            (nnet/resnets/resnet_batchnorm.py:33-36) */
  const void* mul;       /* optional elementwise multiplier applied last: NHWC in the operand dtype with row stride mul_ld
                            (0 = Cout), or NCHW fp32 when out_mode == 1 (the cos factors of the iResBlock VJP chain) */
  int64_t mul_ld;
  void* aux_cos;         /* optional second output (out_mode 0, layout of out_*, operand dtype): cos(2 pi v) of the value v that
                            enters `act` — the derivative of Sin, kept for the log-det estimators (iresblock.py:253-273) */
  int32_t gn_goff;       /* group index that output channel 0 maps to in gn_partial (the second tensor of a channel concat starts at
                            group Ca / cpg of the consumer's GroupNorm, models/ncsnpp.py:350) */
  float* gn2_partial;    /* optional second statistics target with its own (cpg, groups, goff): a skip-connection tensor feeds the
                            next block's GroupNorm and, later, the concatenated GroupNorm of the up path */
  int32_t gn2_cpg, gn2_groups, gn2_goff;
  void* splitk_ws;       /* optional fp32 workspace (splitk_ws_bytes): lets launches with too few output tiles for the chip (4x4 / 8x8
                            feature maps) split K across CTAs; a finish kernel reduces the partial sums and applies the epilogue.
                            Bytes needed: up to (#SMs / #tiles) * N*H*W*Cout*4; smaller workspaces just split less. */
  int64_t splitk_ws_bytes;
  /* padded-pixel ("PP") operand layout (round 2, small feature maps): a (and a2) point to a buffer of
   * ((N (H + 1) + 1) (W + 2)) rows of Cin channels in which pixel (n, y, x) lives at row (n (H + 1) + y + 1)(W + 2) + x + 1 and
   * every other row is zero (written by indm_gn_apply_pp).  Only for taps == 9, stride 1, BF16, plain epilogues; outputs are dense. */
  int32_t a_pp;
} indm_igemm_t;

int indm_igemm(const indm_igemm_t* desc, void* stream);

/* ------------------------------------------------------------------------------------------------------------------
 * Bandwidth-bound kernels of the score network (NHWC)
 * ---------------------------------------------------------------------------------------------------------------- */

/* GroupNorm statistics over the channel-concatenation of xa [N,P,Ca] and xb [N,P,Cb] (xb may be NULL):
 * partial[n][g][2] += (sum, sumsq), G groups over C = Ca + Cb channels (nn.GroupNorm, models/layerspp.py:232;
 * the concat is models/ncsnpp.py:350).  in_dtype: INDM_DTYPE_F32 or INDM_DTYPE_BF16. */
int indm_gn_stats(const void* xa, int Ca, const void* xb, int Cb, int in_dtype, int64_t N, int64_t P, int G, float* partial,
                  void* stream);

/* GroupNorm apply (+ optional SiLU, + optional x2 nearest upsample or 2x2 mean downsample of the result, i.e. the
 * h-branch of ResnetBlockBigGANpp.forward, models/layerspp.py:256-271 with models/up_or_down_sampling.py:59-69),
 * reading statistics from partial[n][g][2] (sum, sumsq over count = P * C / G values).
 * out: [N, P', C] in out_dtype (BF16, or TF32 = fp32 rounded to tf32); raw (optional): the un-normalised input,
 * same resampling, same dtype — the operand of the fused skip 1x1 conv.  resample: 0 none, 1 up x2, 2 down x2. */
int indm_gn_apply(const void* xa, int Ca, const void* xb, int Cb, int in_dtype, int64_t N, int H, int W, int G,
                  const float* partial, const float* gamma, const float* beta, float eps, int act_silu, int resample,
                  void* out, void* raw, int out_dtype, void* stream);

/* indm_gn_apply (no resampling, no raw copy) followed by training-mode dropout of the activated result
 * (nn.Dropout(p) in ResnetBlockBigGANpp, models/layerspp.py:278): kept elements are scaled by 1/(1-p).  The mask is the Philox
 * stream (seed = drop_ctl[0], drop_stream, element quad); drop_ctl is a DEVICE buffer {seed, enabled}: enabled = 0 makes the
 * call identical to indm_gn_apply, so one launch plan serves eval and train.  The backward kernels recompute the mask. */
int indm_gn_apply_dropout(const void* xa, int Ca, const void* xb, int Cb, int in_dtype, int64_t N, int H, int W, int G,
                          const float* partial, const float* gamma, const float* beta, float eps, int act_silu, void* out,
                          int out_dtype, float drop_p, const uint64_t* drop_ctl, uint32_t drop_stream, void* stream);

/* indm_gn_apply (no resampling; optional dropout as in indm_gn_apply_dropout, drop_ctl may be NULL) writing `out` (and the optional
 * un-normalised copy `raw`) in the padded-pixel layout of indm_igemm_t.a_pp: pixel (n, y, x) -> row (n (H + 1) + y + 1)(W + 2) + x + 1
 * of a buffer of (N (H + 1) + 1)(W + 2) rows of C channels.  Only the interior rows are written: the caller zeroes the buffer
 * once and the border rows stay zero (the 'same' padding of the 3x3 convolutions of ResnetBlockBigGANpp, models/layerspp.py:262-281). */
int indm_gn_apply_pp(const void* xa, int Ca, const void* xb, int Cb, int in_dtype, int64_t N, int H, int W, int G,
                     const float* partial, const float* gamma, const float* beta, float eps, int act_silu, void* out, void* raw,
                     int out_dtype, float drop_p, const uint64_t* drop_ctl, uint32_t drop_stream, void* stream);

/* Fused single-head self-attention forward of AttnBlockpp (models/layerspp.py:94-99): out[n] = softmax(q k^T * scale) v for
 * qkv = [N, L, 3C] (q | k | v along the channel axis, as the fused q|k|v projection writes it) -> out [N, L, C].  Scores live in
 * TMEM and probabilities in shared memory: neither reaches HBM.  Covers dtype BF16, L = 256, C = 256 (the 16x16 attention of the
 * CIFAR-10 / CelebA networks); anything else returns INDM_ERR_UNSUPPORTED and callers use the GEMM + indm_softmax_rows path. */
int indm_attention_fwd(const void* qkv, void* out, int64_t N, int L, int C, float scale, int dtype, void* stream);

/* Row softmax of fp32 scores s[rows][cols] -> probabilities in out_dtype (models/layerspp.py:96-97). */
int indm_softmax_rows(const float* s, void* out, int64_t rows, int cols, int out_dtype, void* stream);

/* Network input: x NCHW fp32 [N,C,H,W] -> NHWC with cpad >= C channels (extra channels zero), v = x*mul + add
 * (the `2x - 1` of models/ncsnpp.py:278-280 when data is not centred), then act (0 none, 1 Sin — the leading
 * activation of an iResBlock branch, flows/resflow/resflow_.py:442-444), in out_dtype. */
int indm_prep_input(const float* x, void* out, int64_t N, int C, int H, int W, int cpad, float mul, float add, int act,
                    int out_dtype, void* stream);

/* Time embedding (models/layers.py:515-529 positional: kind 0, time_cond = 999 t;
 * models/layerspp.py:45-54 Gaussian Fourier: kind 1, time_cond = sigma, freqs = W[dim/2]).
 * If sched != NULL the conditioning value is sched[(*step) * sched_ld + sched_col] for every sample (sampler loop
 * replayed from a CUDA graph); otherwise time_cond[n].  out [N, dim] fp32. */
int indm_time_embedding(const float* time_cond, const float* sched, const int32_t* step, int sched_ld, int sched_col,
                        const float* freqs, int kind, int64_t N, int dim, float* out, void* stream);

/* out[n][o] = g(bias[o] + sum_k f(in[n][k]) * w[o][k]), f / g = SiLU if act_in / act_out else identity (the nn.Linear
 * layers of the temb MLP, models/ncsnpp.py:270-274).  FP32 math; out stored as out_dtype (F32, BF16, or TF32-rounded
 * fp32) so it can feed the tensor-core GEMM that evaluates all Dense_0 layers at once (models/layerspp.py:276). */
int indm_linear_f32(const float* in, const float* w, const float* bias, void* out, int64_t N, int K, int O, int act_in,
                    int act_out, int out_dtype, void* stream);

/* FIR resampling of an NHWC tensor with the separable kernel outer(k1,k1)/sum^2*gain (models/up_or_down_sampling.py:195-257):
 * mode 1: upsample_2d (up 2, pad (2,1), gain 4); mode 2: downsample_2d (down 2, pad (1,1));
 * mode 3: the FIR stage of conv_downsample_2d (up = down = 1, pad (2,2)) -> [H+1, W+1];
 * mode 4: its transpose (up = down = 1, pad (1,1)) -> [H-1, W-1] (op/upfirdn2d.py:111-114 g_pad); the transposes of modes 1 / 2
 * are modes 2 / 1 with the same taps.
 * k1: HOST pointer to the 4 separable taps, already normalised (k/sum(k), times 2 per axis for mode 1). */
int indm_fir_nhwc(const void* x, void* y, int dtype_in, int dtype_out, int64_t N, int H, int W, int C, const float* k1, int mode,
                  void* stream);

/* ------------------------------------------------------------------------------------------------------------------
 * Backward (vector-Jacobian) kernels of the score network.  The reference gets these from the autograd engine:
 * likelihood.py:27-38 (`torch.autograd.grad(fn_eps, x)`, Hutchinson divergence) and losses.py:250,304 (`.backward()`).
 * Convolution / attention-product gradients are indm_igemm calls with transposed, tap-flipped weight packs.
 * ---------------------------------------------------------------- */

/* Backward of y = resample(act(GroupNorm(concat(xa, xb)))) (what indm_gn_apply computes), in two passes.
 * stats: partial_bwd[n][g][2] += (sum dxhat, sum dxhat * xhat) with dxhat = gamma * act'(u) * resample^T(dy); optional
 *        dgamma[c] += sum act'(u) dy xhat, dbeta[c] += sum act'(u) dy (atomics, training).  Caller zeroes partial_bwd.
 * apply: dx = rstd * (dxhat - mean_g(dxhat) - xhat * mean_g(dxhat * xhat)) + resample^T(extra_post) + extra_scale * extra_pre,
 *        written to dxa [N,H,W,Ca] / dxb [N,H,W,Cb] in out_dtype (F32: optionally accumulated; BF16: overwritten).
 * dy: [N,H',W',C] in dy_dtype; extra_post: optional fp32 [N,H',W',C] (gradient w.r.t. the raw resampled input that feeds the
 * skip 1x1 conv); extra_pre: optional fp32 [N,H,W,C] (identity skip).  Supported (dy, x, out) dtypes: (BF16,F32,F32),
 * (BF16,BF16,BF16), (F32,F32,F32).  drop_p / drop_ctl / drop_stream: the dropout of indm_gn_apply_dropout (drop_ctl NULL = none). */
int indm_gn_bwd_stats(const void* dy, int dy_dtype, const void* xa, int Ca, const void* xb, int Cb, int x_dtype, int64_t N, int H,
                      int W, int G, const float* partial_fwd, const float* gamma, const float* beta, float eps, int act_silu,
                      int resample, float* partial_bwd, float* dgamma, float* dbeta, int out_dtype, float drop_p,
                      const uint64_t* drop_ctl, uint32_t drop_stream, void* stream);
int indm_gn_bwd_apply(const void* dy, int dy_dtype, const void* xa, int Ca, const void* xb, int Cb, int x_dtype, int64_t N, int H,
                      int W, int G, const float* partial_fwd, const float* gamma, const float* beta, float eps, int act_silu,
                      int resample, const float* partial_bwd, const float* extra_post, const float* extra_pre, float extra_scale,
                      void* dxa, int acc_a, void* dxb, int acc_b, int out_dtype, float drop_p, const uint64_t* drop_ctl,
                      uint32_t drop_stream, void* stream);

/* Convolution / linear weight gradient on the tensor cores (the cuDNN wgrad / cuBLAS calls behind losses.py:250,304):
 *   dw[o*stride_o + c*stride_c + t*stride_t] += scale * sum_{n,y,x} dy[n,y,x,o] * x[n, y+ky-1, x+kx-1, c]     (taps = 9, t = ky*3+kx)
 *   dw[o*stride_o + c*stride_c]              += scale * sum_p dy[p,o] * x[p,c]                                   (taps = 1)
 * dy [N,H,W,Cout] (row stride dy_ld, 0 = Cout) and x [N,H,W,Cin] (x_ld) are NHWC in `dtype` (BF16 -> tensor cores; fp32 ->
 * exact fp32 validation path on the CUDA cores);
 * dw is fp32 and is ACCUMULATED into (atomics; split-K over CTAs), so the caller zeroes it once per step.  The strides let
 * the result land directly in the parameter's own layout: nn.Conv2d [Cout,Cin,3,3] -> (Cin*9, 9, 1); NIN W[in,out] -> (1, out, 0). */
int indm_conv_wgrad(const void* dy, int64_t dy_ld, const void* x, int64_t x_ld, int dtype, int N, int H, int W, int Cout, int Cin,
                    int taps, float* dw, int64_t stride_o, int64_t stride_c, int64_t stride_t, float scale, void* stream);

/* out[i] = in[i] * scale, fp32 -> out_dtype (BF16 or fp32), n % 4 == 0: the operand copy of a gradient tensor */
int indm_cast_scale(const float* in, void* out, int64_t n, float scale, int out_dtype, void* stream);

/* ds[i][j] = p[i][j] * (dp[i][j] - sum_k p[i][k] dp[i][k]) * scale; p, ds in dtype (BF16 / fp32), dp fp32 */
int indm_softmax_bwd_rows(const float* dp, const void* p, void* ds, int64_t rows, int cols, float scale, int dtype, void* stream);

/* out[b][c][r] = in[b*in_batch_stride + r*in_ld + c]  (BF16 or fp32 elements; in_ld = 0 -> C, in_batch_stride = 0 -> R*in_ld;
 * out is dense [B][C][R]) */
int indm_transpose_batched(const void* in, void* out, int64_t B, int R, int C, int64_t in_ld, int64_t in_batch_stride, int dtype,
                           void* stream);

/* out NHWC [N,H,W,cpad] (channels >= C zero) = x NCHW fp32 [N,C,H,W] * mul * rowscale[n] (rowscale may be NULL) */
int indm_nchw_to_nhwc(const float* x, const float* rowscale, void* out, int64_t N, int C, int H, int W, int cpad, float mul,
                      int out_dtype, void* stream);

/* out NCHW fp32 [N,C,H,W] = scale * x[n,y,x,c] for x NHWC fp32 with row stride ld >= C */
int indm_nhwc_to_nchw_f32(const float* x, int64_t ld, float* out, int64_t N, int C, int H, int W, float scale, void* stream);

/* Backward of the stride-2 VALID 3x3 convolution of the VE input pyramid (models/up_or_down_sampling.py:173-178; forward =
 * indm_igemm with stride 2): dy [N,H,W,Cout] in `dtype` (BF16 / fp32), w = the nn.Conv2d parameter [Cout,Cin,3,3] fp32.
 * dgrad: dx fp32 [N,2H+1,2W+1,cin_ld] (channels >= Cin zero) is overwritten; wgrad: dw [Cout,Cin,3,3] fp32 is accumulated into,
 * x = the forward input [N,2H+1,2W+1,x_ld] in `dtype`.  CUDA-core kernels: these three layers are < 0.1 % of the network. */
int indm_conv_s2_dgrad(const void* dy, const float* w, float* dx, int dtype, int64_t N, int H, int W, int Cout, int Cin, int cin_ld,
                       void* stream);
int indm_conv_s2_wgrad(const void* dy, const void* x, float* dw, int dtype, int64_t N, int H, int W, int Cout, int Cin, int x_ld,
                       void* stream);

/* out[n] (+)= scale * sum_i a[n][i] * b[n][i], fp32 (the eps^T (J eps) contraction of likelihood.py:36-37) */
int indm_rowdot_f32(const float* a, const float* b, float* out, int64_t N, int64_t D, float scale, int accumulate, void* stream);

/* ------------------------------------------------------------------------------------------------------------------
 * Training step pieces (losses.py:48-62 optimize_fn, :99-118 loss, models/ema.py:32-51)
 * ---------------------------------------------------------------- */

/* Column sums of an NHWC tensor x [N,P,C] (row stride x_ld, 0 = C) in `dtype` (BF16 / fp32):
 * out_img[n*out_ld + c] += scale * sum_p x[n,p,c] (per-image: gradient of the time-embedding row bias, models/layerspp.py:276)
 * and / or out_tot[c] += scale * sum_{n,p} x[n,p,c] (bias gradients).  Either output may be NULL.  Accumulates (atomics). */
int indm_colsum(const void* x, int dtype, int64_t N, int64_t P, int C, int64_t x_ld, float* out_img, int64_t out_ld, float* out_tot,
                float scale, void* stream);

/* C = alpha * op(A) op(B) + beta * C, fp32 row-major, op = transpose if trans_* (small matrices of the time-embedding MLP
 * and Dense_0 layers: nn.Linear forward / backward, models/ncsnpp.py:270-274) */
int indm_sgemm_f32(int trans_a, int trans_b, int M, int N, int K, float alpha, const float* A, int64_t lda, const float* B, int64_t ldb,
                   float beta, float* C, int64_t ldc, void* stream);

/* The same GEMM over a batch: operand z of A / B / C at base + z * stride (elements; stride 0 = shared by the whole batch).
 * Used to run the per-iResBlock conditioning-path products of the flow backward (32 blocks) as one launch each. */
int indm_sgemm_batched_f32(int trans_a, int trans_b, int M, int N, int K, float alpha, const float* A, int64_t lda, int64_t stride_a,
                           const float* B, int64_t ldb, int64_t stride_b, float beta, float* C, int64_t ldc, int64_t stride_c, int batch,
                           void* stream);

/* dx = dy * SiLU'(pre) */
int indm_silu_bwd_f32(const float* dy, const float* pre, float* dx, int64_t n, void* stream);

/* out[n] = a[n] * x[n] + b[n] * z[n] over D elements per sample: x_t = mean + std z (losses.py:104-105) */
int indm_perturb_f32(const float* x, const float* z, const float* a, const float* b, float* out, int64_t N, int64_t D, void* stream);

/* Denoising score matching (losses.py:108-118): r = score*std[n] + z; loss[n] = 0.5 * w[n] * norm * sum(r^2);
 * dscore (optional) = gscale * w[n] * norm * std[n] * r = gscale * d loss[n] / d score.  w may be NULL (= 1). */
int indm_dsm_loss_f32(const float* score, const float* z, const float* std_, const float* w, float* loss, float* dscore, int64_t N,
                      int64_t D, float norm, float gscale, void* stream);

/* out[0] += sum x^2 (global gradient norm for clip_grad_norm_, losses.py:58-59; caller zeroes out) */
int indm_sumsq_f32(const float* x, int64_t n, float* out, void* stream);

/* One pass over flat parameter storage: g *= min(1, max_norm / (sqrt(*grad_sumsq) + 1e-6)) (skipped when grad_sumsq == NULL or
 * max_norm < 0), torch.optim.AdamW update with bias correction for `step` (1-based), then, if ema != NULL,
 * ema -= (1 - ema_decay) * (ema - p)  (models/ema.py:43-51). */
int indm_adamw_ema_f32(float* p, const float* g, float* m, float* v, float* ema, int64_t n, float lr, float beta1, float beta2,
                       float eps, float weight_decay, int64_t step, const float* grad_sumsq, float max_norm, float ema_decay,
                       void* stream);
int indm_ema_f32(float* ema, const float* p, int64_t n, float decay, void* stream);

/* ------------------------------------------------------------------------------------------------------------------
 * Predictor-corrector update (sampling.py:205-210 ReverseDiffusionPredictor, :272-292 LangevinCorrector with
 * sde_lib.py:105-118,171-184,310-323), state NCHW fp32 [N, D].
 * ---------------------------------------------------------------------------------------------------------------- */

/* x_mean = a*x + c*s ;  x = x_mean + d*z.   (a, c, d) = coef[0..2] of row (*step) of a [steps][coef_ld] device table
 * (step may be NULL = row 0).  z = NULL -> standard normal noise generated in-kernel (Philox4x32-10, counter =
 * (seed, *step, rng_offset, element); seed_dev != NULL: the seed is read from device memory instead, so a captured
 * CUDA graph can be replayed with a new seed).  x is updated in place; x_mean (optional) receives the noise-free state. */
int indm_pc_predictor_update(float* x, const float* s, const float* z, float* x_mean, const float* coef, int coef_ld,
                             const int32_t* step, int64_t N, int64_t D, uint64_t seed, const uint64_t* seed_dev,
                             uint64_t rng_offset, void* stream);

/* per-sample sums of squares: out[n][0] = |s_n|^2, out[n][1] = |z_n|^2 (z NULL -> the Philox noise the following
 * indm_langevin_update call with the same seed/offset will draw). */
int indm_langevin_norms(const float* s, const float* z, float* out, const int32_t* step, int64_t N, int64_t D, uint64_t seed,
                        const uint64_t* seed_dev, uint64_t rng_offset, void* stream);

/* step = 2*alpha*(snr * mean_n|z_n| / mean_n|s_n|)^2 ; x_mean = x + step*s ; x = x_mean + sqrt(2 step) z.
 * (alpha, snr) = coef[0..1] of row (*step). */
int indm_langevin_update(float* x, const float* s, const float* z, float* x_mean, const float* norms, const float* coef,
                         int coef_ld, const int32_t* step, int64_t N, int64_t D, uint64_t seed, const uint64_t* seed_dev,
                         uint64_t rng_offset, void* stream);

/* Batch-sharded sampling with GLOBAL Langevin statistics (the reference takes .mean() over the whole batch, sampling.py:286-287):
 * sums[0] = sum_n |s_n|, sums[1] = sum_n |z_n|, sums[2] = N from the per-sample sums of squares of indm_langevin_norms; the caller
 * all-reduces the three floats over ranks (NCCL, on the same stream) and hands them to indm_langevin_update_global, which is
 * indm_langevin_update with the two means taken as gsums[0] / gsums[2], gsums[1] / gsums[2]. */
int indm_langevin_norm_sums(const float* norms, float* sums, int64_t N, void* stream);
int indm_langevin_update_global(float* x, const float* s, const float* z, float* x_mean, const float* gsums, const float* coef,
                                int coef_ld, const int32_t* step, int64_t N, int64_t D, uint64_t seed, const uint64_t* seed_dev,
                                uint64_t rng_offset, void* stream);

/* *step += 1 (device-side step counter advanced inside the captured graph) */
int indm_advance_step(int32_t* step, void* stream);

/* out[i] = sched[*step][col_num] (/ sched[*step][col_den] if col_den >= 0), i < n: per-step scalar broadcast to a
 * per-sample vector (the network's output scale -1/std(t) of models/utils.py:176-177, or 1/sigma of
 * models/ncsnpp.py:410-412) without leaving the device. */
int indm_sched_broadcast(float* out, int64_t n, const float* sched, int ld, int col_num, int col_den, const int32_t* step,
                         void* stream);

/* fill with standard normal noise (same Philox stream as above), fp32 */
int indm_randn_f32(float* out, int64_t n, uint64_t seed, uint64_t rng_offset, void* stream);

/* ------------------------------------------------------------------------------------------------------------------
 * Flow side (wolf / residual flow)
 * ---------------------------------------------------------------------------------------------------------------- */

/* One op of the 64-d latent prior flow program (flow_models/wolf/modules/discriminators/priors/flow.py:16-201).
 * off[] are offsets (in floats) into one packed fp32 parameter buffer:
 *   ACTNORM  : off[0] log_scale[64], off[1] bias[64]                       (flows/normalization.py:26-70)
 *   LINEAR   : off[0] matrix[64][64] applied as y = x W^T                   (flows/permutation.py:91-134)
 *   COUPLING : off[0..5] fc1.w[256][32], fc1.b, fc2.w[256][256], fc2.b, fc3.w_eff[64][256] (weight-norm folded), fc3.b
 *              split_skip: 0 halves / 1 even-odd; up: conditioner = part 1     (flows/couplings/coupling.py:48-145)
 * backward = 0: the op's forward(); 1: its backward().  */
#define INDM_FLOW_OP_ACTNORM 0
#define INDM_FLOW_OP_LINEAR 1
#define INDM_FLOW_OP_COUPLING 2
typedef struct indm_flow_op {
  int32_t kind, backward, split_skip, up;
  int64_t off[6];
} indm_flow_op_t;

/* out[n] = program(in[n]) for n < N, in/out [N,64]; logdet[n] (optional) = sum of the ops' log-determinants +
 * logdet_const (the invertible-linear slogdet terms, constants of the weights, computed by the caller).
 * kl_base != NULL: logdet[n] instead receives kl_base[n] - (log N(out[n]; 0, I) + log-determinant), i.e. the KL term of
 * FlowPrior.calcKL (priors/flow.py:233-253) when kl_base = log q(h|x).  params / ops are DEVICE pointers. */
int indm_prior_flow(const float* in, float* out, float* logdet, const float* params, const indm_flow_op_t* ops, int n_ops,
                    float logdet_const, const float* kl_base, int64_t N, void* stream);

/* Reparameterised posterior sample (modules/discriminators/gaussian.py:29-38): c [N,128] = (mu | logvar), eps [N,64] ->
 * h = mu + exp(logvar/2) eps and log q(h|x) = -(sum(logvar + eps^2) + 64 log 2pi)/2 (priors/flow.py:236-241). */
int indm_posterior_sample(const float* c, const float* eps, float* h, float* logq, int64_t N, void* stream);

/* Tap packing for the few-channel 3x3 convolutions of the iResBlock branch (c in {3, 12, 48}; resflow_.py:441-470), so that
 * they run as ONE 1x1 tensor-core GEMM over 9c packed values instead of nine 64-wide zero-padded taps:
 *   im2col: out[n,y,x, t*c + ch] = f(x[n, ch, y + s*dy_t, x + s*dx_t]), zero outside the image and for columns >= 9c;
 *           x NCHW fp32, out NHWC [N,H,W,Kp] in out_dtype, f = Sin if act == 1, s = -1 if flip (transposed convolution).
 *   col2im: out[n,ch,y,x] = (residual) + scale * (bias[ch] + sum_t in[n, y + s*dy_t, x + s*dx_t, t*c + ch]), then * mul;
 *           in fp32 NHWC with row stride ld >= 9c; bias / residual / mul (NCHW fp32) optional. */
int indm_im2col3x3_nchw(const float* x, void* out, int64_t N, int c, int H, int W, int Kp, int flip, int act, int out_dtype,
                        void* stream);
int indm_col2im3x3_nchw(const float* in, int64_t ld, const float* bias, const float* residual, const float* mul, float scale, float* out,
                        int64_t N, int c, int H, int W, int flip, void* stream);

/* y += alpha * x (fp32): the Neumann-series accumulation of iresblock.py:264-270 */
int indm_axpy_f32(float* y, const float* x, float alpha, int64_t n, void* stream);

/* One term of the Neumann log-det series (iresblock.py:264-273) with its coefficient read from device memory:
 * acc += coef[*k] * v ; cur = v.  `k` is a device counter the caller advances with indm_advance_step after the launch, so one
 * captured launch sequence serves every term of every series length. */
int indm_series_step_f32(float* acc, float* cur, const float* v, const float* coef, int32_t* k, int64_t n, void* stream);

/* out = cos(2 pi x) (fp32): derivative of the leading Sin of an iResBlock branch w.r.t. the block input */
int indm_cos2pi_f32(const float* x, float* out, int64_t n, void* stream);

/* flag[0] = max_i (x[i]-x_prev[i])^2 / (atol + |y[i]|*rtol)  — the stop rule of iResBlock._inverse_fixed_point
 * (flows/resflow/layers/iresblock.py:78-88): converged iff flag[0] < 1.  flag is a device float. */
int indm_fixed_point_check(const float* x, const float* x_prev, const float* y, int64_t n, float atol, float rtol, float* flag,
                           void* stream);

/* ------------------------------------------------------------------------------------------------------------------
 * Wolf-flow TRAINING backward (what torch.autograd derives in losses.py:300-304 from iresblock.py:264-273, lipschitz.py:350-359):
 * HBM-bound helpers; the contractions themselves run on indm_igemm / indm_conv_wgrad.  `dtype` = operand dtype of the branch's
 * hidden activations (INDM_DTYPE_BF16, or INDM_DTYPE_TF32 / _F32 = fp32 storage).  n % 8 == 0.
 * ---------------------------------------------------------------- */

/* out = a * b */
int indm_mul_op(const void* a, const void* b, void* out, int64_t n, int dtype, void* stream);

/* out = a (+ a2, may be NULL) + coef * b * c * d: the second-order term  r . u . Sin''(pre)  of the Neumann estimator's gradient
 * (Sin'' = -4 pi^2 Sin, so d is the stored post-activation) added to the first-order gradient a */
int indm_fma3_op(const void* a, const void* a2, const void* b, const void* c, const void* d, void* out, int64_t n, float coef, int dtype,
                 void* stream);

/* out = sin(2 pi x) / (2 pi) (fp32): Sin of layers/base/activations.py:7-12 on the block input */
int indm_sin2pi_f32(const float* x, float* out, int64_t n, void* stream);

/* out[n][i] = x[n][i] * s[n] (fp32, D values per sample): per-sample loss weight folded into the Neumann vector */
int indm_rowscale_f32(const float* x, const float* s, float* out, int64_t N, int64_t D, void* stream);

/* Backward of LopConv2d.compute_weight, domain = codomain = inf (lipschitz.py:350-359): raw, dwn (gradient w.r.t. the normalised
 * weight), out [rows][cols] fp32, rows = output channels: out (+)= dwn / f - [l1 > coeff] <dwn, raw>_row / (f^2 coeff) sign(raw),
 * f = max(1, l1 / coeff), l1 = ||raw_row||_1 */
int indm_lop_bwd_f32(const float* raw, const float* dwn, float* out, int rows, int cols, float coeff, int accumulate, void* stream);

/* Backward of ck[n] * KL[n], KL = log q(h|x) - log p(h) (priors/flow.py:233-253), through the latent prior flow's 'forward' op
 * program (the one indm_prior_flow evaluates with kl_base; every op.backward == 0).  gh [N,64] = d / d h of the -log p part.
 * Per-sample parameter-gradient factors are written to workspaces as [N, dim] fp32 matrices, slot = index of the op among the
 * ops of its kind, so the caller forms the weight gradients with small GEMMs over the batch:
 *   ws_c [n_couplings][ zin N x 32 | ha N x 256 | hb N x 256 | d1 N x 256 | d2 N x 256 | d3 N x 64 ]   (N * 1120 floats per coupling)
 *        fc1: dW = d1^T zin, db = colsum d1;  fc2: dW = d2^T ha;  fc3 (weight-normed, folded): dW = d3^T hb
 *   ws_a [n_actnorms][ d log_scale N x 64 | d bias N x 64 ]
 *   ws_l [n_linears ][ x N x 64 | gy N x 64 ]      dW = gy^T x  (+ the log|det W| term, -sum(ck) W^-T, added by the caller) */
int indm_prior_flow_bwd(const float* h, const float* params, const indm_flow_op_t* ops, int n_ops, const float* ck, float* ws_c,
                        float* ws_a, float* ws_l, float* gh, int64_t N, void* stream);

/* Backward of indm_posterior_sample + the log q term of the KL: c [N,128] = (mu | logvar), gh = total d / d h,
 * gc [N,128] = (gh | gh eps exp(logvar/2)/2 - ck/2) */
int indm_posterior_bwd(const float* c, const float* eps, const float* gh, const float* ck, float* gc, int64_t N, void* stream);

/* nn.BatchNorm2d in batch-statistics (training) mode for the posterior encoder (nnet/resnets/resnet_batchnorm.py:18-76), NHWC:
 * y [P, ld] fp32 raw convolution output with C real channels (channels >= C are padding and are written as zero),
 * sums [2C] = (sum_p y | sum_p y^2), accumulated (caller zeroes), eps = 1e-5, biased variance for the normalisation.
 *   apply:     out = act(gamma (y - mean) rstd + beta (+ residual)), act 0 none / 2 ELU -> operand dtype and / or fp32
 *   bwd_stats: bsum [2C] = (sum gs | sum gs xhat), gs = g * ELU'(o) with o the post-activation value in `dtype` (NULL: no act)
 *   bwd_apply: dy = gamma rstd (gs - bsum0 / P - xhat bsum1 / P) in `dtype`; gs_out (fp32, optional) = gs
 * d gamma = bsum1, d beta = bsum0. */
int indm_bn_stats(const float* y, int64_t P, int C, int ld, float* sums, void* stream);
int indm_bn_apply(const float* y, const float* sums, const float* gamma, const float* beta, int64_t P, int C, int ld,
                  const float* residual, int act, void* out_op, float* out_f32, int dtype, void* stream);
int indm_bn_bwd_stats(const float* g, const void* o, const float* y, const float* sums, int64_t P, int C, int ld, float* bsum,
                      int dtype, void* stream);
int indm_bn_bwd_apply(const float* g, const void* o, const float* y, const float* sums, const float* bsum, const float* gamma,
                      int64_t P, int C, int ld, void* dy, float* gs_out, int dtype, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* INDM_B200_H_ */
