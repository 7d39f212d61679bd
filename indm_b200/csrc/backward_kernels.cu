// Backward (vector-Jacobian) kernels of the score network that are not GEMM-shaped.  The reference obtains all of these
// from the PyTorch autograd engine (likelihood.py:27-38 `torch.autograd.grad(fn_eps, x)` for the Hutchinson divergence,
// losses.py:250,304 `.backward()` for training): GroupNorm(+SiLU)(+resample) backward, softmax backward, layout helpers.
// The GEMM-shaped parts (conv dgrad, attention products) reuse igemm.cu with transposed / tap-flipped weight packs.
#include "../../include/indm_b200.h"
#include "common.cuh"
#include "nhwc.cuh"
#include "philox.cuh"

namespace {

__device__ __forceinline__ float4 f4_add(float4 a, float4 b) { return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
__device__ __forceinline__ float4 f4_scale(float4 a, float s) { return make_float4(a.x * s, a.y * s, a.z * s, a.w * s); }

struct GnBwdP {
  const void* dy;        // gradient w.r.t. y = resample(act(GN(x))), [N, H', W', C]
  const void* xa;        // forward input, channels [0, Ca)
  const void* xb;        // forward input, channels [Ca, Ca + Cb) or null
  int Ca, Cb, H, W, G, R;
  const float* partial_fwd;  // [N][G][2] (sum, sumsq) of the forward pass
  const float* gamma;
  const float* beta;
  float eps;
  int act;
  float* partial_bwd;        // [N][G][2]: (sum dxhat, sum dxhat * xhat)
  // apply only
  const float* extra_post;   // optional fp32 [N, H', W', C]: gradient w.r.t. resample(x) (skip 1x1 conv input), added outside the norm
  const float* extra_pre;    // optional fp32 [N, H, W, C]: added as extra_scale * extra_pre (identity skip path)
  float extra_scale;
  void* dxa;
  void* dxb;
  int acc_a, acc_b;          // fp32 outputs only: accumulate instead of overwrite
  float* dgamma;             // optional [C] (atomicAdd), training only
  float* dbeta;
  float drop_p;              // dropout applied after the activation in the forward (resample == 0 only); mask recomputed here
  const unsigned long long* drop_ctl;
  unsigned drop_stream;
};

// gradient arriving at pre-resample pixel (y, x): RES 0 same pixel, 1 (forward nearest-up) sum of the 2x2 children,
// 2 (forward 2x2 mean) a quarter of the parent
template <typename T, int RES>
__device__ __forceinline__ float4 load_dy(const T* __restrict__ dy, long long n, int y, int x, int H, int W, int C, int c) {
  if (RES == 0) return Vec4<T>::load(dy + ((n * H + y) * W + x) * C + c);
  if (RES == 1) {
    const long long Wo = 2LL * W;
    const T* b = dy + ((n * 2 * H + 2 * y) * Wo + 2 * x) * C + c;
    return f4_add(f4_add(Vec4<T>::load(b), Vec4<T>::load(b + C)), f4_add(Vec4<T>::load(b + Wo * C), Vec4<T>::load(b + (Wo + 1) * C)));
  }
  const int Ho = H >> 1, Wo = W >> 1;
  return f4_scale(Vec4<T>::load(dy + ((n * Ho + (y >> 1)) * Wo + (x >> 1)) * C + c), 0.25f);
}

// SiLU'(u) = sg (1 + u (1 - sg)).  FAST (BF16 production path): sigmoid through ONE SFU op, sg = (1 + tanh(u / 2)) / 2
// (tanh.approx, ~2^-11 relative: below the BF16 gradients it multiplies); the validation path keeps ex2 + rcp.
template <bool FAST>
__device__ __forceinline__ float dsilu(float u) {
  float sg;
  if (FAST) {
    float t;
    asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(0.5f * u));
    sg = fmaf(0.5f, t, 0.5f);
  } else {
    sg = 1.0f / (1.0f + __expf(-u));
  }
  return sg * (1.0f + u * (1.0f - sg));
}

// MODE 0: statistics pass; MODE 1: apply pass.  grid (splits, N), block = Q * R threads (Q = C/4 channel quads).
template <typename TDy, typename TX, typename TOut, int RES, int MODE>
__global__ void gn_bwd_kernel(const GnBwdP p) {
  __shared__ float s_1[32], s_2[32];
  const int C = p.Ca + p.Cb;
  const int Q = C >> 2;
  const int q = threadIdx.x % Q;
  const int rr = threadIdx.x / Q;
  const long long n = blockIdx.y;
  const int c = q * 4;
  const int cpg = C / p.G;
  const int g = c / cpg;
  const long long P = (long long)p.H * p.W;
  const bool dropping = RES == 0 && p.drop_ctl != nullptr && p.drop_p > 0.f && p.drop_ctl[1] != 0ull;
  const unsigned long long drop_seed = dropping ? p.drop_ctl[0] : 0ull;
  if (MODE == 0) {
    for (int i = threadIdx.x; i < 32; i += blockDim.x) {
      s_1[i] = 0.f;
      s_2[i] = 0.f;
    }
    __syncthreads();
  }
  const float cnt = (float)((double)P * cpg);
  const float su = p.partial_fwd[(n * p.G + g) * 2 + 0];
  const float sq = p.partial_fwd[(n * p.G + g) * 2 + 1];
  const float mean = su / cnt;
  const float rstd = rsqrtf(fmaxf(sq / cnt - mean * mean, 0.f) + p.eps);
  const float4 ga = *reinterpret_cast<const float4*>(p.gamma + c);
  const float4 be = *reinterpret_cast<const float4*>(p.beta + c);
  float m1 = 0.f, m2 = 0.f;
  if (MODE == 1) {
    m1 = p.partial_bwd[(n * p.G + g) * 2 + 0] / cnt;
    m2 = p.partial_bwd[(n * p.G + g) * 2 + 1] / cnt;
  }
  const TX* src;
  int ld;
  TOut* dst;
  int acc;
  if (c < p.Ca) {
    src = (const TX*)p.xa + n * P * p.Ca + c;
    ld = p.Ca;
    dst = MODE == 1 ? (TOut*)p.dxa + n * P * p.Ca + c : nullptr;
    acc = p.acc_a;
  } else {
    src = (const TX*)p.xb + n * P * p.Cb + (c - p.Ca);
    ld = p.Cb;
    dst = MODE == 1 ? (TOut*)p.dxb + n * P * p.Cb + (c - p.Ca) : nullptr;
    acc = p.acc_b;
  }
  const long long per = (P + gridDim.x - 1) / gridDim.x;
  const long long p0 = (long long)blockIdx.x * per, p1 = min(P, p0 + per);
  float s1 = 0.f, s2 = 0.f;
  float4 dgam = make_float4(0.f, 0.f, 0.f, 0.f), dbet = make_float4(0.f, 0.f, 0.f, 0.f);
  if (rr < p.R) {
    for (long long pp = p0 + rr; pp < p1; pp += p.R) {
      const int y = (int)(pp / p.W), x = (int)(pp % p.W);
      const float4 v = Vec4<TX>::load(src + pp * ld);
      const float4 d = load_dy<TDy, RES>((const TDy*)p.dy, n, y, x, p.H, p.W, C, c);
      const float4 xh = make_float4((v.x - mean) * rstd, (v.y - mean) * rstd, (v.z - mean) * rstd, (v.w - mean) * rstd);
      float4 du = d;
      if (RES == 0 && dropping) {
        const float4 k = dropout_scale4(drop_seed, p.drop_stream, (unsigned long long)((n * P + pp) * Q + q), p.drop_p);
        du = make_float4(du.x * k.x, du.y * k.y, du.z * k.z, du.w * k.w);
      }
      if (p.act) {
        du.x *= dsilu<std::is_same<TDy, __nv_bfloat16>::value>(xh.x * ga.x + be.x);
        du.y *= dsilu<std::is_same<TDy, __nv_bfloat16>::value>(xh.y * ga.y + be.y);
        du.z *= dsilu<std::is_same<TDy, __nv_bfloat16>::value>(xh.z * ga.z + be.z);
        du.w *= dsilu<std::is_same<TDy, __nv_bfloat16>::value>(xh.w * ga.w + be.w);
      }
      const float4 dxh = make_float4(du.x * ga.x, du.y * ga.y, du.z * ga.z, du.w * ga.w);
      if (MODE == 0) {
        s1 += (dxh.x + dxh.y) + (dxh.z + dxh.w);
        s2 += (dxh.x * xh.x + dxh.y * xh.y) + (dxh.z * xh.z + dxh.w * xh.w);
        if (p.dgamma) {
          dgam.x += du.x * xh.x; dgam.y += du.y * xh.y; dgam.z += du.z * xh.z; dgam.w += du.w * xh.w;
          dbet.x += du.x; dbet.y += du.y; dbet.z += du.z; dbet.w += du.w;
        }
      } else {
        float4 o = make_float4(rstd * (dxh.x - m1 - xh.x * m2), rstd * (dxh.y - m1 - xh.y * m2), rstd * (dxh.z - m1 - xh.z * m2),
                               rstd * (dxh.w - m1 - xh.w * m2));
        if (p.extra_post) o = f4_add(o, load_dy<float, RES>(p.extra_post, n, y, x, p.H, p.W, C, c));
        if (p.extra_pre) o = f4_add(o, f4_scale(Vec4<float>::load(p.extra_pre + (n * P + pp) * C + c), p.extra_scale));
        if (std::is_same<TOut, float>::value && acc) o = f4_add(o, Vec4<float>::load((const float*)dst + pp * ld));
        Vec4<TOut>::store(dst + pp * ld, o);
      }
    }
  }
  if (MODE == 0) {
    atomicAdd(&s_1[g], s1);
    atomicAdd(&s_2[g], s2);
    __syncthreads();
    for (int i = threadIdx.x; i < p.G; i += blockDim.x) {
      atomicAdd(&p.partial_bwd[(n * p.G + i) * 2 + 0], s_1[i]);
      atomicAdd(&p.partial_bwd[(n * p.G + i) * 2 + 1], s_2[i]);
    }
    if (p.dgamma && rr < p.R) {
      atomicAdd(p.dgamma + c + 0, dgam.x); atomicAdd(p.dgamma + c + 1, dgam.y);
      atomicAdd(p.dgamma + c + 2, dgam.z); atomicAdd(p.dgamma + c + 3, dgam.w);
      atomicAdd(p.dbeta + c + 0, dbet.x); atomicAdd(p.dbeta + c + 1, dbet.y);
      atomicAdd(p.dbeta + c + 2, dbet.z); atomicAdd(p.dbeta + c + 3, dbet.w);
    }
  }
}

template <typename TDy, typename TX, typename TOut, int MODE>
int gn_bwd_launch(const GnBwdP& p, long long N, int resample, cudaStream_t stream) {
  const int C = p.Ca + p.Cb;
  const GnGeom g = gn_geom(C, (long long)p.H * p.W, N);
  GnBwdP q = p;
  q.R = g.R;
  dim3 grid(g.splits, (unsigned)N);
  if (resample == 0) gn_bwd_kernel<TDy, TX, TOut, 0, MODE><<<grid, g.threads, 0, stream>>>(q);
  else if (resample == 1) gn_bwd_kernel<TDy, TX, TOut, 1, MODE><<<grid, g.threads, 0, stream>>>(q);
  else gn_bwd_kernel<TDy, TX, TOut, 2, MODE><<<grid, g.threads, 0, stream>>>(q);
  INDM_CHECK_LAUNCH(MODE == 0 ? "gn_bwd_stats" : "gn_bwd_apply");
  return INDM_OK;
}

template <int MODE>
int gn_bwd_dispatch(const GnBwdP& p, long long N, int resample, int dy_dtype, int x_dtype, int out_dtype, cudaStream_t stream) {
  const bool dy_b = dy_dtype == INDM_DTYPE_BF16, x_b = x_dtype == INDM_DTYPE_BF16, o_b = out_dtype == INDM_DTYPE_BF16;
  if (dy_b && !x_b && !o_b) return gn_bwd_launch<__nv_bfloat16, float, float, MODE>(p, N, resample, stream);
  if (dy_b && x_b && o_b) return gn_bwd_launch<__nv_bfloat16, __nv_bfloat16, __nv_bfloat16, MODE>(p, N, resample, stream);
  if (!dy_b && !x_b && !o_b) return gn_bwd_launch<float, float, float, MODE>(p, N, resample, stream);
  indm_set_error("gn_bwd: unsupported dtype combination dy=%d x=%d out=%d", dy_dtype, x_dtype, out_dtype);
  return INDM_ERR_UNSUPPORTED;
}

// ---------------------------------------------------------------- elementwise cast with scale (fp32 -> operand dtype)
template <typename TOut>
__global__ void cast_scale_kernel(const float* __restrict__ in, TOut* __restrict__ out, long long n4, float scale) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x)
    Vec4<TOut>::store(out + i * 4, f4_scale(Vec4<float>::load(in + i * 4), scale));
}

// ---------------------------------------------------------------- softmax backward, one warp per row
// ds[i][j] = p[i][j] * (dp[i][j] - sum_k p[i][k] dp[i][k]) * scale
template <typename TP>
__global__ void softmax_bwd_kernel(const float* __restrict__ dp, const TP* __restrict__ p, TP* __restrict__ ds, long long rows,
                                   int cols, float scale) {
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  const float* d = dp + row * cols;
  const TP* pr = p + row * cols;
  float dot = 0.f;
  for (int i = lane; i < cols; i += 32) dot += (float)pr[i] * d[i];
  dot = warp_sum(dot);
  for (int i = lane; i < cols; i += 32) ds[row * cols + i] = (TP)((float)pr[i] * (d[i] - dot) * scale);
}

// ---------------------------------------------------------------- batched 2-D transpose  in [B][R][C] -> out [B][C][R]
template <typename T>
__global__ void transpose_kernel(const T* __restrict__ in, T* __restrict__ out, int R, int C, long long in_ld, long long in_bs) {
  pdl_trigger();
  pdl_wait();
  __shared__ T tile[32][33];
  const long long b = blockIdx.z;
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  const T* src = in + b * in_bs;
  T* dst = out + b * (long long)R * C;
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const int r = r0 + j, c = c0 + threadIdx.x;
    if (r < R && c < C) tile[j][threadIdx.x] = src[(long long)r * in_ld + c];
  }
  __syncthreads();
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const int c = c0 + j, r = r0 + threadIdx.x;
    if (r < R && c < C) dst[(long long)c * R + r] = tile[threadIdx.x][j];
  }
}

// ---------------------------------------------------------------- NCHW fp32 [N,C,H,W] * rowscale[n] -> NHWC operand (C padded)
template <typename TOut>
__global__ void nchw_to_nhwc_kernel(const float* __restrict__ x, const float* __restrict__ rowscale, TOut* __restrict__ out,
                                    long long N, int C, int HW, int cpad, float mul) {
  const long long total = N * HW * cpad;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % cpad);
    const long long pix = i / cpad;
    const long long n = pix / HW;
    const int p = (int)(pix % HW);
    float v = 0.f;
    if (c < C) v = x[(n * C + c) * HW + p] * mul * (rowscale ? rowscale[n] : 1.0f);
    out[i] = (TOut)v;
  }
}

// ---------------------------------------------------------------- NHWC fp32 [N,H,W,ld] (first C channels) -> NCHW fp32 * scale
__global__ void nhwc_to_nchw_kernel(const float* __restrict__ x, long long ld, float* __restrict__ out, long long N, int C, int HW,
                                    float scale) {
  const long long total = N * C * HW;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int p = (int)(i % HW);
    const long long r = i / HW;
    const int c = (int)(r % C);
    const long long n = r / C;
    out[i] = x[(n * HW + p) * ld + c] * scale;
  }
}

// ---------------------------------------------------------------- stride-2 VALID 3x3 convolution (input pyramid) backward
// The three pyramid convolutions (models/layerspp.py:142-176 -> up_or_down_sampling.py:173-178) are < 0.1 % of the network's
// FLOPs; their gradients run on the CUDA cores.  Forward: Y[n,y,x,co] = sum X[n,2y+ky,2x+kx,ci] W[co][ci][ky][kx].
// dgrad: one CTA per input pixel (n,u,v), threads over ci:  dX[n,u,v,ci] = sum_{ky,kx: (u-ky),(v-kx) even, in range} sum_co dY W
template <typename T>
__global__ void conv_s2_dgrad_kernel(const T* __restrict__ dy, const float* __restrict__ w, float* __restrict__ dx, int H, int W,
                                     int Cout, int Cin, int cin_ld) {
  extern __shared__ float sdy[];                       // up to 4 contributing output pixels x Cout
  const int AW = 2 * W + 1, AH = 2 * H + 1;
  long long pix = blockIdx.x;
  const int v = (int)(pix % AW);
  pix /= AW;
  const int u = (int)(pix % AH);
  const long long n = pix / AH;
  int taps[4], ntap = 0;
  for (int ky = 0; ky < 3; ++ky) {
    const int yy = u - ky;
    if (yy < 0 || (yy & 1) || (yy >> 1) >= H) continue;
    for (int kx = 0; kx < 3; ++kx) {
      const int xx = v - kx;
      if (xx < 0 || (xx & 1) || (xx >> 1) >= W) continue;
      const long long src = ((n * H + (yy >> 1)) * W + (xx >> 1)) * Cout;
      for (int c = threadIdx.x; c < Cout; c += blockDim.x) sdy[ntap * Cout + c] = (float)dy[src + c];
      taps[ntap++] = ky * 3 + kx;
    }
  }
  __syncthreads();
  for (int ci = threadIdx.x; ci < cin_ld; ci += blockDim.x) {
    float acc = 0.f;
    if (ci < Cin) {
      for (int t = 0; t < ntap; ++t) {
        const float* wr = w + (long long)ci * 9 + taps[t];
        const float* d = sdy + t * Cout;
        for (int co = 0; co < Cout; ++co) acc += d[co] * wr[(long long)co * Cin * 9];
      }
    }
    dx[(((n * AH) + u) * AW + v) * cin_ld + ci] = acc;
  }
}

// wgrad: grid (Cout, 9), threads over ci:  dW[co][ci][t] += sum_{n,y,x} dY[n,y,x,co] X[n,2y+ky,2x+kx,ci]
template <typename T>
__global__ void conv_s2_wgrad_kernel(const T* __restrict__ dy, const T* __restrict__ x, float* __restrict__ dw, long long N, int H, int W,
                                     int Cout, int Cin, int x_ld) {
  const int co = blockIdx.x, t = blockIdx.y;
  const int ky = t / 3, kx = t % 3;
  const int AW = 2 * W + 1, AH = 2 * H + 1;
  for (int ci = threadIdx.x; ci < Cin; ci += blockDim.x) {
    float acc = 0.f;
    for (long long n = 0; n < N; ++n)
      for (int y = 0; y < H; ++y)
        for (int xx = 0; xx < W; ++xx) {
          const float g = (float)dy[((n * H + y) * W + xx) * Cout + co];
          acc += g * (float)x[((n * AH + 2 * y + ky) * AW + 2 * xx + kx) * x_ld + ci];
        }
    atomicAdd(dw + ((long long)co * Cin + ci) * 9 + t, acc);
  }
}

// ---------------------------------------------------------------- per-sample dot products  out[n] (+)= sum_i a[n][i] * b[n][i]
__global__ void rowdot_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ out, long long D,
                              float scale, int accumulate) {
  __shared__ float red[32];
  const long long n = blockIdx.x;
  float s = 0.f;
  for (long long i = threadIdx.x; i < D; i += blockDim.x) s += a[n * D + i] * b[n * D + i];
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 32) {
    s = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.f;
    s = warp_sum(s);
    if (threadIdx.x == 0) out[n] = (accumulate ? out[n] : 0.f) + s * scale;
  }
}

}  // namespace

static int gn_bwd_common(int mode, const void* dy, int dy_dtype, const void* xa, int Ca, const void* xb, int Cb, int x_dtype, int64_t N,
                         int H, int W, int G, const float* partial_fwd, const float* gamma, const float* beta, float eps, int act_silu,
                         int resample, float* partial_bwd, const float* extra_post, const float* extra_pre, float extra_scale,
                         void* dxa, int acc_a, void* dxb, int acc_b, int out_dtype, float* dgamma, float* dbeta, float drop_p,
                         const uint64_t* drop_ctl, uint32_t drop_stream, cudaStream_t stream) {
  if (!xb) Cb = 0;
  const int C = Ca + Cb;
  INDM_CHECK_ARG(dy && xa && partial_fwd && partial_bwd && gamma && beta && N > 0 && H > 0 && W > 0, "gn_bwd: bad arguments");
  INDM_CHECK_ARG(G >= 1 && G <= 32 && C % G == 0 && (C / G) % 4 == 0 && Ca % 4 == 0 && C / 4 <= 1024,
                 "gn_bwd: need G <= 32, (C/G) %% 4 == 0 (C=%d G=%d)", C, G);
  INDM_CHECK_ARG(resample >= 0 && resample <= 2, "gn_bwd: resample must be 0, 1 or 2");
  INDM_CHECK_ARG(resample != 2 || (H % 2 == 0 && W % 2 == 0), "gn_bwd: down x2 needs even H, W");
  INDM_CHECK_ARG(N <= 65535, "gn_bwd: N too large for grid.y");
  INDM_CHECK_ARG((dgamma == nullptr) == (dbeta == nullptr), "gn_bwd: dgamma and dbeta go together");
  if (mode == 1) INDM_CHECK_ARG(dxa && (Cb == 0 || dxb), "gn_bwd_apply: missing output");
  GnBwdP p{};
  p.dy = dy; p.xa = xa; p.xb = xb; p.Ca = Ca; p.Cb = Cb; p.H = H; p.W = W; p.G = G;
  p.partial_fwd = partial_fwd; p.gamma = gamma; p.beta = beta; p.eps = eps; p.act = act_silu; p.partial_bwd = partial_bwd;
  p.extra_post = extra_post; p.extra_pre = extra_pre; p.extra_scale = extra_scale;
  p.dxa = dxa; p.dxb = dxb; p.acc_a = acc_a; p.acc_b = acc_b; p.dgamma = dgamma; p.dbeta = dbeta;
  p.drop_p = drop_p; p.drop_ctl = (const unsigned long long*)drop_ctl; p.drop_stream = drop_stream;
  return mode == 0 ? gn_bwd_dispatch<0>(p, N, resample, dy_dtype, x_dtype, out_dtype, stream)
                   : gn_bwd_dispatch<1>(p, N, resample, dy_dtype, x_dtype, out_dtype, stream);
}

extern "C" int indm_gn_bwd_stats(const void* dy, int dy_dtype, const void* xa, int Ca, const void* xb, int Cb, int x_dtype, int64_t N,
                                 int H, int W, int G, const float* partial_fwd, const float* gamma, const float* beta, float eps,
                                 int act_silu, int resample, float* partial_bwd, float* dgamma, float* dbeta, int out_dtype,
                                 float drop_p, const uint64_t* drop_ctl, uint32_t drop_stream, void* stream) {
  return gn_bwd_common(0, dy, dy_dtype, xa, Ca, xb, Cb, x_dtype, N, H, W, G, partial_fwd, gamma, beta, eps, act_silu, resample,
                       partial_bwd, nullptr, nullptr, 0.f, nullptr, 0, nullptr, 0, out_dtype, dgamma, dbeta, drop_p, drop_ctl, drop_stream,
                       (cudaStream_t)stream);
}

extern "C" int indm_gn_bwd_apply(const void* dy, int dy_dtype, const void* xa, int Ca, const void* xb, int Cb, int x_dtype, int64_t N,
                                 int H, int W, int G, const float* partial_fwd, const float* gamma, const float* beta, float eps,
                                 int act_silu, int resample, const float* partial_bwd, const float* extra_post, const float* extra_pre,
                                 float extra_scale, void* dxa, int acc_a, void* dxb, int acc_b, int out_dtype, float drop_p,
                                 const uint64_t* drop_ctl, uint32_t drop_stream, void* stream) {
  return gn_bwd_common(1, dy, dy_dtype, xa, Ca, xb, Cb, x_dtype, N, H, W, G, partial_fwd, gamma, beta, eps, act_silu, resample,
                       const_cast<float*>(partial_bwd), extra_post, extra_pre, extra_scale, dxa, acc_a, dxb, acc_b, out_dtype, nullptr,
                       nullptr, drop_p, drop_ctl, drop_stream, (cudaStream_t)stream);
}

extern "C" int indm_cast_scale(const float* in, void* out, int64_t n, float scale, int out_dtype, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  INDM_CHECK_ARG(in && out && n > 0 && n % 4 == 0, "cast_scale: n must be a positive multiple of 4");
  const int grid = grid_for(n / 4, 256);
  if (out_dtype == INDM_DTYPE_BF16) cast_scale_kernel<__nv_bfloat16><<<grid, 256, 0, stream>>>(in, (__nv_bfloat16*)out, n / 4, scale);
  else cast_scale_kernel<float><<<grid, 256, 0, stream>>>(in, (float*)out, n / 4, scale);
  INDM_CHECK_LAUNCH("cast_scale");
  return INDM_OK;
}

extern "C" int indm_softmax_bwd_rows(const float* dp, const void* p, void* ds, int64_t rows, int cols, float scale, int dtype,
                                     void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  INDM_CHECK_ARG(dp && p && ds && rows > 0 && cols > 0, "softmax_bwd_rows: bad arguments");
  const int wpb = 8;
  const long long blocks = (rows + wpb - 1) / wpb;
  if (dtype == INDM_DTYPE_BF16)
    softmax_bwd_kernel<__nv_bfloat16><<<(unsigned)blocks, wpb * 32, 0, stream>>>(dp, (const __nv_bfloat16*)p, (__nv_bfloat16*)ds, rows, cols, scale);
  else
    softmax_bwd_kernel<float><<<(unsigned)blocks, wpb * 32, 0, stream>>>(dp, (const float*)p, (float*)ds, rows, cols, scale);
  INDM_CHECK_LAUNCH("softmax_bwd_rows");
  return INDM_OK;
}

extern "C" int indm_transpose_batched(const void* in, void* out, int64_t B, int R, int C, int64_t in_ld, int64_t in_batch_stride,
                                      int dtype, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  INDM_CHECK_ARG(in && out && B > 0 && R > 0 && C > 0 && B <= 65535, "transpose_batched: bad arguments");
  if (in_ld == 0) in_ld = C;
  if (in_batch_stride == 0) in_batch_stride = (int64_t)R * in_ld;
  dim3 grid((C + 31) / 32, (R + 31) / 32, (unsigned)B), block(32, 8);
  if (dtype == INDM_DTYPE_BF16)
    indm_launch_pdl(transpose_kernel<__nv_bfloat16>, grid, block, 0, stream, (const __nv_bfloat16*)in, (__nv_bfloat16*)out, R, C, in_ld, in_batch_stride);
  else indm_launch_pdl(transpose_kernel<float>, grid, block, 0, stream, (const float*)in, (float*)out, R, C, in_ld, in_batch_stride);
  INDM_CHECK_LAUNCH("transpose_batched");
  return INDM_OK;
}

extern "C" int indm_nchw_to_nhwc(const float* x, const float* rowscale, void* out, int64_t N, int C, int H, int W, int cpad, float mul,
                                 int out_dtype, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  INDM_CHECK_ARG(x && out && N > 0 && C > 0 && cpad >= C, "nchw_to_nhwc: bad arguments");
  const long long total = (long long)N * H * W * cpad;
  const int grid = grid_for(total, 256);
  if (out_dtype == INDM_DTYPE_BF16)
    nchw_to_nhwc_kernel<__nv_bfloat16><<<grid, 256, 0, stream>>>(x, rowscale, (__nv_bfloat16*)out, N, C, H * W, cpad, mul);
  else
    nchw_to_nhwc_kernel<float><<<grid, 256, 0, stream>>>(x, rowscale, (float*)out, N, C, H * W, cpad, mul);
  INDM_CHECK_LAUNCH("nchw_to_nhwc");
  return INDM_OK;
}

extern "C" int indm_nhwc_to_nchw_f32(const float* x, int64_t ld, float* out, int64_t N, int C, int H, int W, float scale, void* stream_) {
  INDM_CHECK_ARG(x && out && N > 0 && C > 0 && ld >= C, "nhwc_to_nchw: bad arguments");
  const long long total = (long long)N * C * H * W;
  nhwc_to_nchw_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream_>>>(x, ld, out, N, C, H * W, scale);
  INDM_CHECK_LAUNCH("nhwc_to_nchw");
  return INDM_OK;
}

extern "C" int indm_conv_s2_dgrad(const void* dy, const float* w, float* dx, int dtype, int64_t N, int H, int W, int Cout, int Cin,
                                  int cin_ld, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  INDM_CHECK_ARG(dy && w && dx && N > 0 && H > 0 && W > 0 && Cout > 0 && Cin > 0 && cin_ld >= Cin, "conv_s2_dgrad: bad arguments");
  const long long blocks = (long long)N * (2 * H + 1) * (2 * W + 1);
  INDM_CHECK_ARG(blocks < (1LL << 31) && 4 * Cout * 4 <= 48 * 1024, "conv_s2_dgrad: problem too large");
  const int threads = cin_ld >= 256 ? 256 : (cin_ld >= 128 ? 128 : 64);
  if (dtype == INDM_DTYPE_BF16)
    conv_s2_dgrad_kernel<__nv_bfloat16><<<(unsigned)blocks, threads, 4 * Cout * 4, stream>>>((const __nv_bfloat16*)dy, w, dx, H, W, Cout, Cin, cin_ld);
  else
    conv_s2_dgrad_kernel<float><<<(unsigned)blocks, threads, 4 * Cout * 4, stream>>>((const float*)dy, w, dx, H, W, Cout, Cin, cin_ld);
  INDM_CHECK_LAUNCH("conv_s2_dgrad");
  return INDM_OK;
}

extern "C" int indm_conv_s2_wgrad(const void* dy, const void* x, float* dw, int dtype, int64_t N, int H, int W, int Cout, int Cin, int x_ld,
                                  void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  INDM_CHECK_ARG(dy && x && dw && N > 0 && H > 0 && W > 0 && Cout > 0 && Cin > 0 && x_ld >= Cin, "conv_s2_wgrad: bad arguments");
  dim3 grid((unsigned)Cout, 9);
  const int threads = Cin >= 256 ? 256 : (Cin >= 128 ? 128 : 64);
  if (dtype == INDM_DTYPE_BF16)
    conv_s2_wgrad_kernel<__nv_bfloat16><<<grid, threads, 0, stream>>>((const __nv_bfloat16*)dy, (const __nv_bfloat16*)x, dw, N, H, W, Cout, Cin, x_ld);
  else
    conv_s2_wgrad_kernel<float><<<grid, threads, 0, stream>>>((const float*)dy, (const float*)x, dw, N, H, W, Cout, Cin, x_ld);
  INDM_CHECK_LAUNCH("conv_s2_wgrad");
  return INDM_OK;
}

extern "C" int indm_rowdot_f32(const float* a, const float* b, float* out, int64_t N, int64_t D, float scale, int accumulate,
                               void* stream_) {
  INDM_CHECK_ARG(a && b && out && N > 0 && D > 0, "rowdot: bad arguments");
  rowdot_kernel<<<(unsigned)N, 256, 0, (cudaStream_t)stream_>>>(a, b, out, D, scale, accumulate);
  INDM_CHECK_LAUNCH("rowdot");
  return INDM_OK;
}
