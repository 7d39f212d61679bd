// Backward-pass helpers of the wolf flow's TRAINING step (joint flow + score training, losses.py:258-320): the pieces
// torch.autograd derives from flow_models/wolf/flows/resflow/layers/iresblock.py:264-273 (Neumann log-det estimator, whose
// parameter gradient is second order in g), layers/base/lipschitz.py:350-359 (soft L-inf weight normalisation),
// modules/encoders/global_encoder.py (BatchNorm in batch-statistics mode) and modules/discriminators/priors/flow.py.
// The dense contractions of those passes run on indm_igemm / indm_conv_wgrad; everything here is HBM-bound elementwise /
// reduction work over the branch's hidden activations (operand dtype: BF16 in production, FP32 in validation mode).
#include <cuda_bf16.h>

#include "../../include/indm_b200.h"
#include "common.cuh"

namespace {

// 8 elements per thread per step: one 16-byte (BF16) or two 16-byte (FP32) accesses per tensor
struct V8 {
  float v[8];
};

template <bool BF16>
__device__ __forceinline__ V8 ld8(const void* p, long long i) {
  V8 r;
  if (BF16) {
    const uint4 u = *reinterpret_cast<const uint4*>((const __nv_bfloat16*)p + i);
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w[k]));
      r.v[2 * k] = f.x;
      r.v[2 * k + 1] = f.y;
    }
  } else {
    const float4 a = *reinterpret_cast<const float4*>((const float*)p + i);
    const float4 b = *reinterpret_cast<const float4*>((const float*)p + i + 4);
    r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w;
    r.v[4] = b.x; r.v[5] = b.y; r.v[6] = b.z; r.v[7] = b.w;
  }
  return r;
}

template <bool BF16>
__device__ __forceinline__ void st8(void* p, long long i, const V8& r) {
  if (BF16) {
    uint4 u;
    u.x = pack_bf16x2(r.v[0], r.v[1]); u.y = pack_bf16x2(r.v[2], r.v[3]);
    u.z = pack_bf16x2(r.v[4], r.v[5]); u.w = pack_bf16x2(r.v[6], r.v[7]);
    *reinterpret_cast<uint4*>((__nv_bfloat16*)p + i) = u;
  } else {
    *reinterpret_cast<float4*>((float*)p + i) = make_float4(r.v[0], r.v[1], r.v[2], r.v[3]);
    *reinterpret_cast<float4*>((float*)p + i + 4) = make_float4(r.v[4], r.v[5], r.v[6], r.v[7]);
  }
}

template <bool BF16>
__global__ void mul_kernel(const void* __restrict__ a, const void* __restrict__ b, void* __restrict__ out, long long n8) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (long long)gridDim.x * blockDim.x) {
    const V8 x = ld8<BF16>(a, i * 8), y = ld8<BF16>(b, i * 8);
    V8 r;
#pragma unroll
    for (int k = 0; k < 8; ++k) r.v[k] = x.v[k] * y.v[k];
    st8<BF16>(out, i * 8, r);
  }
}

// out = a (+ a2) + coef * b * c * d
template <bool BF16>
__global__ void fma3_kernel(const void* __restrict__ a, const void* __restrict__ a2, const void* __restrict__ b, const void* __restrict__ c,
                            const void* __restrict__ d, void* __restrict__ out, long long n8, float coef) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (long long)gridDim.x * blockDim.x) {
    V8 r = ld8<BF16>(a, i * 8);
    if (a2 != nullptr) {
      const V8 t = ld8<BF16>(a2, i * 8);
#pragma unroll
      for (int k = 0; k < 8; ++k) r.v[k] += t.v[k];
    }
    const V8 x = ld8<BF16>(b, i * 8), y = ld8<BF16>(c, i * 8), z = ld8<BF16>(d, i * 8);
#pragma unroll
    for (int k = 0; k < 8; ++k) r.v[k] = fmaf(coef * x.v[k], y.v[k] * z.v[k], r.v[k]);
    st8<BF16>(out, i * 8, r);
  }
}

__global__ void sin2pi_kernel(const float* __restrict__ x, float* __restrict__ out, long long n) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    out[i] = sinpif(2.0f * x[i]) * 0.15915494309189535f;
}

// out[n][i] = x[n][i] * s[n]
__global__ void rowscale_kernel(const float* __restrict__ x, const float* __restrict__ s, float* __restrict__ out, long long D, long long n) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    out[i] = x[i] * s[i / D];
}

// Backward of LopConv2d.compute_weight (lipschitz.py:350-359): W = raw / f, f = max(1, ||raw_row||_1 / coeff), per output row:
//   d raw = dW / f - [||raw_row||_1 > coeff] * <dW, raw>_row / (f^2 coeff) * sign(raw)
// One CTA per row.
__global__ void lop_bwd_kernel(const float* __restrict__ raw, const float* __restrict__ dwn, float* __restrict__ out, int cols, float coeff,
                               int accumulate) {
  __shared__ float red[2][32];
  const long long base = (long long)blockIdx.x * cols;
  float l1 = 0.f, dot = 0.f;
  for (int i = threadIdx.x; i < cols; i += blockDim.x) {
    const float r = raw[base + i];
    l1 += fabsf(r);
    dot += dwn[base + i] * r;
  }
  l1 = warp_sum(l1);
  dot = warp_sum(dot);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0) { red[0][w] = l1; red[1][w] = dot; }
  __syncthreads();
  if (w == 0) {
    float a = l < (int)(blockDim.x >> 5) ? red[0][l] : 0.f;
    float b = l < (int)(blockDim.x >> 5) ? red[1][l] : 0.f;
    a = warp_sum(a);
    b = warp_sum(b);
    if (l == 0) { red[0][0] = a; red[1][0] = b; }
  }
  __syncthreads();
  l1 = red[0][0];
  dot = red[1][0];
  const float s = l1 / coeff;
  const float f = fmaxf(s, 1.0f);
  const float k = s > 1.0f ? dot / (f * f * coeff) : 0.f;
  for (int i = threadIdx.x; i < cols; i += blockDim.x) {
    const float r = raw[base + i];
    const float sg = r > 0.f ? 1.f : (r < 0.f ? -1.f : 0.f);
    const float g = dwn[base + i] / f - k * sg;
    out[base + i] = accumulate ? out[base + i] + g : g;
  }
}

inline unsigned grid_for(long long work) {
  long long blocks = (work + 255) / 256;
  const long long cap = (long long)indm_num_sms() * 16;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (unsigned)blocks;
}

}  // namespace

extern "C" int indm_mul_op(const void* a, const void* b, void* out, int64_t n, int dtype, void* stream_) {
  INDM_CHECK_ARG(a && b && out && n > 0 && n % 8 == 0, "mul_op: bad arguments (n %% 8 == 0)");
  INDM_CHECK_ARG(dtype == INDM_DTYPE_BF16 || dtype == INDM_DTYPE_TF32 || dtype == INDM_DTYPE_F32, "mul_op: dtype");
  if (dtype == INDM_DTYPE_BF16) mul_kernel<true><<<grid_for(n / 8), 256, 0, (cudaStream_t)stream_>>>(a, b, out, n / 8);
  else mul_kernel<false><<<grid_for(n / 8), 256, 0, (cudaStream_t)stream_>>>(a, b, out, n / 8);
  INDM_CHECK_LAUNCH("mul_op");
  return INDM_OK;
}

extern "C" int indm_fma3_op(const void* a, const void* a2, const void* b, const void* c, const void* d, void* out, int64_t n, float coef,
                            int dtype, void* stream_) {
  INDM_CHECK_ARG(a && b && c && d && out && n > 0 && n % 8 == 0, "fma3_op: bad arguments (n %% 8 == 0)");
  INDM_CHECK_ARG(dtype == INDM_DTYPE_BF16 || dtype == INDM_DTYPE_TF32 || dtype == INDM_DTYPE_F32, "fma3_op: dtype");
  if (dtype == INDM_DTYPE_BF16) fma3_kernel<true><<<grid_for(n / 8), 256, 0, (cudaStream_t)stream_>>>(a, a2, b, c, d, out, n / 8, coef);
  else fma3_kernel<false><<<grid_for(n / 8), 256, 0, (cudaStream_t)stream_>>>(a, a2, b, c, d, out, n / 8, coef);
  INDM_CHECK_LAUNCH("fma3_op");
  return INDM_OK;
}

extern "C" int indm_sin2pi_f32(const float* x, float* out, int64_t n, void* stream_) {
  INDM_CHECK_ARG(x && out && n > 0, "sin2pi: bad arguments");
  sin2pi_kernel<<<grid_for(n), 256, 0, (cudaStream_t)stream_>>>(x, out, n);
  INDM_CHECK_LAUNCH("sin2pi");
  return INDM_OK;
}

extern "C" int indm_rowscale_f32(const float* x, const float* s, float* out, int64_t N, int64_t D, void* stream_) {
  INDM_CHECK_ARG(x && s && out && N > 0 && D > 0, "rowscale: bad arguments");
  rowscale_kernel<<<grid_for(N * D), 256, 0, (cudaStream_t)stream_>>>(x, s, out, D, N * D);
  INDM_CHECK_LAUNCH("rowscale");
  return INDM_OK;
}

extern "C" int indm_lop_bwd_f32(const float* raw, const float* dwn, float* out, int rows, int cols, float coeff, int accumulate,
                                void* stream_) {
  INDM_CHECK_ARG(raw && dwn && out && rows > 0 && cols > 0 && coeff > 0.f, "lop_bwd: bad arguments");
  lop_bwd_kernel<<<(unsigned)rows, 256, 0, (cudaStream_t)stream_>>>(raw, dwn, out, cols, coeff, accumulate);
  INDM_CHECK_LAUNCH("lop_bwd");
  return INDM_OK;
}

// ---------------------------------------------------------------------------------------------------------------------
// Backward of the latent prior flow's KL term (priors/flow.py:233-253: KL = log q(h|x) - log p(h), log p(h) = log N(f(h)) + logdet)
// for the 'forward' op program of indm_prior_flow (every op with backward == 0).  One CTA per sample, like the forward kernel:
// a forward sweep keeps every op's input in shared memory, the reverse sweep re-evaluates each coupling MLP from its input and
// back-propagates.  Per-sample parameter-gradient FACTORS (layer inputs and pre-activation deltas) are written to a workspace
// as [N, dim] matrices; the host turns them into weight gradients with small GEMMs (sum over the batch), so there are no
// atomics on the 90 K weights of each coupling.
namespace {

constexpr int PF_DIM = 64, PF_HALF = 32, PF_HID = 256, PF_MAXOPS = 32;
constexpr int PF_CPL_FLOATS = PF_HALF + 4 * PF_HID + PF_DIM;     // zin | ha | hb | d1 | d2 | d3 per sample per coupling

__device__ __forceinline__ float pf_elu(float x) { return x > 0.f ? x : expm1f(x); }

__global__ void __launch_bounds__(PF_HID) prior_flow_bwd_kernel(const float* __restrict__ h, const float* __restrict__ params,
                                                               const indm_flow_op_t* __restrict__ ops, int n_ops,
                                                               const float* __restrict__ ck, float* __restrict__ ws_c,
                                                               float* __restrict__ ws_a, float* __restrict__ ws_l,
                                                               float* __restrict__ gh, long long N) {
  __shared__ float z[PF_DIM], zh[PF_MAXOPS][PF_DIM], zin[PF_HALF], ha[PF_HID], hb[PF_HID], prm[PF_DIM], d3[PF_DIM], d2[PF_HID], d1[PF_HID],
      gz[PF_DIM], gnew[PF_DIM];
  __shared__ int slot[PF_MAXOPS];
  const int t = threadIdx.x;
  const long long n = blockIdx.x;
  if (t < PF_DIM) z[t] = h[n * PF_DIM + t];
  if (t == 0) {
    int cnt[3] = {0, 0, 0};
    for (int oi = 0; oi < n_ops; ++oi) slot[oi] = cnt[ops[oi].kind]++;
  }
  __syncthreads();

  // coupling MLP on zin -> ha, hb, prm (same arithmetic as prior_flow_kernel)
  auto mlp = [&](const indm_flow_op_t& op) {
    {
      const float* w1 = params + op.off[0];
      float acc = params[op.off[1] + t];
      for (int k = 0; k < PF_HALF; ++k) acc += w1[t * PF_HALF + k] * zin[k];
      ha[t] = pf_elu(acc);
    }
    __syncthreads();
    {
      const float* w2 = params + op.off[2];
      float acc = params[op.off[3] + t];
      for (int k = 0; k < PF_HID; ++k) acc += w2[t * PF_HID + k] * ha[k];
      hb[t] = pf_elu(acc);
    }
    __syncthreads();
    if (t < PF_DIM) {
      const float* w3 = params + op.off[4];
      float acc = params[op.off[5] + t];
      for (int k = 0; k < PF_HID; ++k) acc += w3[t * PF_HID + k] * hb[k];
      prm[t] = acc;
    }
    __syncthreads();
  };

  // ---------------- forward sweep
  for (int oi = 0; oi < n_ops; ++oi) {
    const indm_flow_op_t op = ops[oi];
    if (t < PF_DIM) zh[oi][t] = z[t];
    __syncthreads();
    if (op.kind == INDM_FLOW_OP_ACTNORM) {
      if (t < PF_DIM) z[t] = z[t] * expf(params[op.off[0] + t]) + params[op.off[1] + t];
      __syncthreads();
    } else if (op.kind == INDM_FLOW_OP_LINEAR) {
      const float* W = params + op.off[0];
      float acc = 0.f;
      if (t < PF_DIM)
        for (int k = 0; k < PF_DIM; ++k) acc += W[t * PF_DIM + k] * z[k];
      __syncthreads();
      if (t < PF_DIM) z[t] = acc;
      __syncthreads();
    } else {
      const bool skip = op.split_skip != 0, up = op.up != 0;
      if (t < PF_HALF) zin[t] = up ? z[skip ? 2 * t : t] : z[skip ? 2 * t + 1 : PF_HALF + t];
      __syncthreads();
      mlp(op);
      if (t < PF_HALF) {
        const float scale = 1.f / (1.f + expf(-(prm[PF_HALF + t] + 2.0f))) + 1e-3f;
        const int j = up ? (skip ? 2 * t + 1 : PF_HALF + t) : (skip ? 2 * t : t);
        z[j] = scale * z[j] + prm[t];
      }
      __syncthreads();
    }
  }
  // ---------------- seed: d (ck KL) / d z_out = ck z_out ; d / d logdet = -ck
  const float c = ck[n];
  const float g_ld = -c;
  if (t < PF_DIM) gz[t] = c * z[t];
  __syncthreads();

  // ---------------- reverse sweep
  for (int oi = n_ops - 1; oi >= 0; --oi) {
    const indm_flow_op_t op = ops[oi];
    const float* x = zh[oi];
    const int sl = slot[oi];
    if (op.kind == INDM_FLOW_OP_ACTNORM) {
      if (t < PF_DIM) {
        const float e = expf(params[op.off[0] + t]);
        float* a = ws_a + (long long)sl * 2 * N * PF_DIM;
        a[n * PF_DIM + t] = gz[t] * x[t] * e + g_ld;            // d log_scale
        a[N * PF_DIM + n * PF_DIM + t] = gz[t];                   // d bias
        gz[t] *= e;
      }
      __syncthreads();
    } else if (op.kind == INDM_FLOW_OP_LINEAR) {
      const float* W = params + op.off[0];
      if (t < PF_DIM) {
        float* l = ws_l + (long long)sl * 2 * N * PF_DIM;
        l[n * PF_DIM + t] = x[t];
        l[N * PF_DIM + n * PF_DIM + t] = gz[t];
        float acc = 0.f;
        for (int k = 0; k < PF_DIM; ++k) acc += W[k * PF_DIM + t] * gz[k];
        gnew[t] = acc;
      }
      __syncthreads();
      if (t < PF_DIM) gz[t] = gnew[t];
      __syncthreads();
    } else {
      const bool skip = op.split_skip != 0, up = op.up != 0;
      auto ic = [&](int j) { return up ? (skip ? 2 * j : j) : (skip ? 2 * j + 1 : PF_HALF + j); };          // conditioning half
      auto it = [&](int j) { return up ? (skip ? 2 * j + 1 : PF_HALF + j) : (skip ? 2 * j : j); };          // transformed half
      if (t < PF_HALF) zin[t] = x[ic(t)];
      __syncthreads();
      mlp(op);
      if (t < PF_HALF) {
        const float sg = 1.f / (1.f + expf(-(prm[PF_HALF + t] + 2.0f)));
        const float scale = sg + 1e-3f;
        const int j = it(t);
        const float gy = gz[j];
        d3[t] = gy;                                                        // d mu
        d3[PF_HALF + t] = (gy * x[j] + g_ld / scale) * sg * (1.f - sg);    // d r
        gz[j] = gy * scale;
      }
      __syncthreads();
      {
        const float* w3 = params + op.off[4];
        float acc = 0.f;
        for (int k = 0; k < PF_DIM; ++k) acc += w3[k * PF_HID + t] * d3[k];
        d2[t] = acc * (hb[t] > 0.f ? 1.f : hb[t] + 1.f);                   // ELU' = exp(pre) = ELU + 1 on the negative side
      }
      __syncthreads();
      {
        const float* w2 = params + op.off[2];
        float acc = 0.f;
        for (int k = 0; k < PF_HID; ++k) acc += w2[k * PF_HID + t] * d2[k];
        d1[t] = acc * (ha[t] > 0.f ? 1.f : ha[t] + 1.f);
      }
      __syncthreads();
      if (t < PF_HALF) {
        const float* w1 = params + op.off[0];
        float acc = 0.f;
        for (int k = 0; k < PF_HID; ++k) acc += w1[k * PF_HALF + t] * d1[k];
        gz[ic(t)] += acc;
      }
      // factors for the host-side weight-gradient GEMMs
      float* base = ws_c + (long long)sl * N * PF_CPL_FLOATS;
      float* p_zin = base;
      float* p_ha = p_zin + N * PF_HALF;
      float* p_hb = p_ha + N * PF_HID;
      float* p_d1 = p_hb + N * PF_HID;
      float* p_d2 = p_d1 + N * PF_HID;
      float* p_d3 = p_d2 + N * PF_HID;
      if (t < PF_HALF) p_zin[n * PF_HALF + t] = zin[t];
      p_ha[n * PF_HID + t] = ha[t];
      p_hb[n * PF_HID + t] = hb[t];
      p_d1[n * PF_HID + t] = d1[t];
      p_d2[n * PF_HID + t] = d2[t];
      if (t < PF_DIM) p_d3[n * PF_DIM + t] = d3[t];
      __syncthreads();
    }
  }
  if (t < PF_DIM) gh[n * PF_DIM + t] = gz[t];
}

// Backward of the reparameterised posterior sample + log q (gaussian.py:29-38, priors/flow.py:236-241):
//   h = mu + exp(logvar / 2) eps,  log q = -(sum(logvar + eps^2) + 64 log 2 pi) / 2,  loss term ck * (log q - log p(h))
//   g_mu = gh,  g_logvar = gh * eps * exp(logvar / 2) / 2 - ck / 2;   c = (mu | logvar) [N, 128]
__global__ void posterior_bwd_kernel(const float* __restrict__ cc, const float* __restrict__ eps, const float* __restrict__ gh,
                                     const float* __restrict__ ck, float* __restrict__ gc) {
  const int t = threadIdx.x;   // 64
  const long long n = blockIdx.x;
  const float lv = cc[n * 2 * PF_DIM + PF_DIM + t], e = eps[n * PF_DIM + t], g = gh[n * PF_DIM + t];
  gc[n * 2 * PF_DIM + t] = g;
  gc[n * 2 * PF_DIM + PF_DIM + t] = 0.5f * g * e * expf(0.5f * lv) - 0.5f * ck[n];
}

}  // namespace

extern "C" int indm_prior_flow_bwd(const float* h, const float* params, const indm_flow_op_t* ops, int n_ops, const float* ck, float* ws_c,
                                   float* ws_a, float* ws_l, float* gh, int64_t N, void* stream_) {
  INDM_CHECK_ARG(h && params && ops && ck && ws_c && ws_a && ws_l && gh && N > 0 && n_ops > 0 && n_ops <= PF_MAXOPS,
                 "prior_flow_bwd: bad arguments (n_ops <= 32)");
  prior_flow_bwd_kernel<<<(unsigned)N, PF_HID, 0, (cudaStream_t)stream_>>>(h, params, ops, n_ops, ck, ws_c, ws_a, ws_l, gh, (long long)N);
  INDM_CHECK_LAUNCH("prior_flow_bwd");
  return INDM_OK;
}

extern "C" int indm_posterior_bwd(const float* c, const float* eps, const float* gh, const float* ck, float* gc, int64_t N, void* stream_) {
  INDM_CHECK_ARG(c && eps && gh && ck && gc && N > 0, "posterior_bwd: bad arguments");
  posterior_bwd_kernel<<<(unsigned)N, PF_DIM, 0, (cudaStream_t)stream_>>>(c, eps, gh, ck, gc);
  INDM_CHECK_LAUNCH("posterior_bwd");
  return INDM_OK;
}

// ---------------------------------------------------------------------------------------------------------------------
// BatchNorm2d in batch-statistics (training) mode for the posterior encoder (modules/encoders/global_encoder.py:12-44,
// nnet/resnets/resnet_batchnorm.py:18-76): NHWC fp32 conv outputs y [P, ld] with C <= 512 real channels (ld >= C, channels
// >= C are padding and stay zero).  sums = (sum_p y | sum_p y^2) per channel.
namespace {

constexpr float BN_EPS = 1e-5f;

template <bool BF16>
__device__ __forceinline__ float ld_op(const void* p, long long i) {
  return BF16 ? __bfloat162float(((const __nv_bfloat16*)p)[i]) : ((const float*)p)[i];
}
template <bool BF16>
__device__ __forceinline__ void st_op(void* p, long long i, float v) {
  if (BF16) ((__nv_bfloat16*)p)[i] = __float2bfloat16_rn(v);
  else ((float*)p)[i] = v;
}

// blockDim = (32 channels, 8 row lanes); grid = (ceil(C/32), row chunks)
__global__ void bn_stats_kernel(const float* __restrict__ y, long long P, int C, int ld, float* __restrict__ sums) {
  __shared__ float sh[2][8][32];
  const int c = blockIdx.x * 32 + threadIdx.x;
  float s = 0.f, q = 0.f;
  if (c < C) {
    for (long long p = (long long)blockIdx.y * 8 + threadIdx.y; p < P; p += (long long)gridDim.y * 8) {
      const float v = y[p * ld + c];
      s += v;
      q += v * v;
    }
  }
  sh[0][threadIdx.y][threadIdx.x] = s;
  sh[1][threadIdx.y][threadIdx.x] = q;
  __syncthreads();
  if (threadIdx.y == 0 && c < C) {
#pragma unroll
    for (int r = 1; r < 8; ++r) { s += sh[0][r][threadIdx.x]; q += sh[1][r][threadIdx.x]; }
    atomicAdd(sums + c, s);
    atomicAdd(sums + C + c, q);
  }
}

// out = act(gamma * (y - mean) * rstd + beta (+ residual)); one thread per element of [P, ld]
template <bool BF16>
__global__ void bn_apply_kernel(const float* __restrict__ y, const float* __restrict__ sums, const float* __restrict__ gamma,
                                const float* __restrict__ beta, long long P, int C, int ld, const float* __restrict__ residual, int act,
                                void* __restrict__ out_op, float* __restrict__ out_f32) {
  const long long total = P * ld;
  const float invP = 1.0f / (float)P;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % ld);
    float v = 0.f;
    if (c < C) {
      const float mean = sums[c] * invP;
      const float var = fmaxf(sums[C + c] * invP - mean * mean, 0.f);
      v = (y[i] - mean) * rsqrtf(var + BN_EPS) * gamma[c] + beta[c];
      if (residual) v += residual[i];
      if (act == 2) v = v > 0.f ? v : expm1f(v);
    }
    if (out_op) st_op<BF16>(out_op, i, v);
    if (out_f32) out_f32[i] = v;
  }
}

// bsum = (sum_p gs | sum_p gs * xhat), gs = g * ELU'(o) (o = post-activation value, operand dtype; NULL = no activation)
template <bool BF16>
__global__ void bn_bwd_stats_kernel(const float* __restrict__ g, const void* __restrict__ o, const float* __restrict__ y,
                                    const float* __restrict__ sums, long long P, int C, int ld, float* __restrict__ bsum) {
  __shared__ float sh[2][8][32];
  const int c = blockIdx.x * 32 + threadIdx.x;
  float s = 0.f, q = 0.f;
  if (c < C) {
    const float invP = 1.0f / (float)P;
    const float mean = sums[c] * invP;
    const float rstd = rsqrtf(fmaxf(sums[C + c] * invP - mean * mean, 0.f) + BN_EPS);
    for (long long p = (long long)blockIdx.y * 8 + threadIdx.y; p < P; p += (long long)gridDim.y * 8) {
      const long long i = p * ld + c;
      float gs = g[i];
      if (o) {
        const float ov = ld_op<BF16>(o, i);
        gs *= ov > 0.f ? 1.f : ov + 1.f;
      }
      s += gs;
      q += gs * (y[i] - mean) * rstd;
    }
  }
  sh[0][threadIdx.y][threadIdx.x] = s;
  sh[1][threadIdx.y][threadIdx.x] = q;
  __syncthreads();
  if (threadIdx.y == 0 && c < C) {
#pragma unroll
    for (int r = 1; r < 8; ++r) { s += sh[0][r][threadIdx.x]; q += sh[1][r][threadIdx.x]; }
    atomicAdd(bsum + c, s);
    atomicAdd(bsum + C + c, q);
  }
}

// dy = gamma rstd (gs - mean(gs) - xhat mean(gs xhat)) -> operand dtype; optional gs_out (fp32) = gradient w.r.t. the pre-activation sum
template <bool BF16>
__global__ void bn_bwd_apply_kernel(const float* __restrict__ g, const void* __restrict__ o, const float* __restrict__ y,
                                    const float* __restrict__ sums, const float* __restrict__ bsum, const float* __restrict__ gamma,
                                    long long P, int C, int ld, void* __restrict__ dy, float* __restrict__ gs_out) {
  const long long total = P * ld;
  const float invP = 1.0f / (float)P;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % ld);
    float d = 0.f, gs = 0.f;
    if (c < C) {
      const float mean = sums[c] * invP;
      const float rstd = rsqrtf(fmaxf(sums[C + c] * invP - mean * mean, 0.f) + BN_EPS);
      gs = g[i];
      if (o) {
        const float ov = ld_op<BF16>(o, i);
        gs *= ov > 0.f ? 1.f : ov + 1.f;
      }
      const float xh = (y[i] - mean) * rstd;
      d = gamma[c] * rstd * (gs - bsum[c] * invP - xh * bsum[C + c] * invP);
    }
    st_op<BF16>(dy, i, d);
    if (gs_out) gs_out[i] = gs;
  }
}

}  // namespace

extern "C" int indm_bn_stats(const float* y, int64_t P, int C, int ld, float* sums, void* stream_) {
  INDM_CHECK_ARG(y && sums && P > 0 && C > 0 && ld >= C, "bn_stats: bad arguments");
  long long chunks = (P + 63) / 64;
  if (chunks > 1024) chunks = 1024;
  dim3 grid((C + 31) / 32, (unsigned)chunks), block(32, 8);
  bn_stats_kernel<<<grid, block, 0, (cudaStream_t)stream_>>>(y, P, C, ld, sums);
  INDM_CHECK_LAUNCH("bn_stats");
  return INDM_OK;
}

extern "C" int indm_bn_apply(const float* y, const float* sums, const float* gamma, const float* beta, int64_t P, int C, int ld,
                             const float* residual, int act, void* out_op, float* out_f32, int dtype, void* stream_) {
  INDM_CHECK_ARG(y && sums && gamma && beta && (out_op || out_f32) && P > 0 && C > 0 && ld >= C, "bn_apply: bad arguments");
  if (dtype == INDM_DTYPE_BF16)
    bn_apply_kernel<true><<<grid_for(P * ld), 256, 0, (cudaStream_t)stream_>>>(y, sums, gamma, beta, P, C, ld, residual, act, out_op, out_f32);
  else
    bn_apply_kernel<false><<<grid_for(P * ld), 256, 0, (cudaStream_t)stream_>>>(y, sums, gamma, beta, P, C, ld, residual, act, out_op, out_f32);
  INDM_CHECK_LAUNCH("bn_apply");
  return INDM_OK;
}

extern "C" int indm_bn_bwd_stats(const float* g, const void* o, const float* y, const float* sums, int64_t P, int C, int ld, float* bsum,
                                 int dtype, void* stream_) {
  INDM_CHECK_ARG(g && y && sums && bsum && P > 0 && C > 0 && ld >= C, "bn_bwd_stats: bad arguments");
  long long chunks = (P + 63) / 64;
  if (chunks > 1024) chunks = 1024;
  dim3 grid((C + 31) / 32, (unsigned)chunks), block(32, 8);
  if (dtype == INDM_DTYPE_BF16) bn_bwd_stats_kernel<true><<<grid, block, 0, (cudaStream_t)stream_>>>(g, o, y, sums, P, C, ld, bsum);
  else bn_bwd_stats_kernel<false><<<grid, block, 0, (cudaStream_t)stream_>>>(g, o, y, sums, P, C, ld, bsum);
  INDM_CHECK_LAUNCH("bn_bwd_stats");
  return INDM_OK;
}

extern "C" int indm_bn_bwd_apply(const float* g, const void* o, const float* y, const float* sums, const float* bsum, const float* gamma,
                                 int64_t P, int C, int ld, void* dy, float* gs_out, int dtype, void* stream_) {
  INDM_CHECK_ARG(g && y && sums && bsum && gamma && dy && P > 0 && C > 0 && ld >= C, "bn_bwd_apply: bad arguments");
  if (dtype == INDM_DTYPE_BF16)
    bn_bwd_apply_kernel<true><<<grid_for(P * ld), 256, 0, (cudaStream_t)stream_>>>(g, o, y, sums, bsum, gamma, P, C, ld, dy, gs_out);
  else
    bn_bwd_apply_kernel<false><<<grid_for(P * ld), 256, 0, (cudaStream_t)stream_>>>(g, o, y, sums, bsum, gamma, P, C, ld, dy, gs_out);
  INDM_CHECK_LAUNCH("bn_bwd_apply");
  return INDM_OK;
}
