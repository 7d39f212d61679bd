// Training-side kernels that are not convolution-shaped: reductions for bias / time-embedding gradients, a small FP32 GEMM for
// the time-embedding MLP (a few hundred rows), the denoising-score-matching loss with its gradient, and the optimizer tail
// (global gradient-norm clip + AdamW + EMA in one pass over flat parameter storage).
// Reference: losses.py:48-62 (optimize_fn: clip_grad_norm_, optimizer.step), :99-118 (loss), models/ema.py:32-51 (EMA),
// torch.optim.AdamW (decoupled weight decay) and the autograd-generated gradients of nn.Linear / bias terms.
#include <math.h>

#include "../../include/indm_b200.h"
#include "common.cuh"
#include "nhwc.cuh"

namespace {

// ---------------------------------------------------------------- column sums of an NHWC tensor
// out_img[n][c] += scale * sum_p x[n][p][c]   and / or   out_tot[c] += scale * sum_{n,p} x[n][p][c].   grid (splits, N)
template <typename T>
__global__ void colsum_kernel(const T* __restrict__ x, long long P, int C, long long x_ld, float* __restrict__ out_img,
                              long long out_ld, float* __restrict__ out_tot, float scale) {
  const int Q = C >> 2;
  const int q = threadIdx.x % Q, rr = threadIdx.x / Q, R = blockDim.x / Q;
  const long long n = blockIdx.y;
  const long long per = (P + gridDim.x - 1) / gridDim.x;
  const long long p0 = (long long)blockIdx.x * per, p1 = min(P, p0 + per);
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  // 8 rows per trip, all loads issued before the first add (one dependent 8-byte load per trip left the kernel latency-bound:
  // 77 us for a 134 MB tensor)
  const T* base = x + n * P * x_ld + q * 4;
  long long p = p0 + rr;
  for (; p + 7LL * R < p1; p += 8LL * R) {
    float4 v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) v[u] = Vec4<T>::load(base + (p + (long long)u * R) * x_ld);
#pragma unroll
    for (int u = 0; u < 8; ++u) { acc.x += v[u].x; acc.y += v[u].y; acc.z += v[u].z; acc.w += v[u].w; }
  }
  for (; p < p1; p += R) {
    const float4 v = Vec4<T>::load(base + p * x_ld);
    acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
  }
  const int c = q * 4;
  // the R row groups of the CTA hold partial sums of the same columns: fold them through shared memory so that one thread per
  // column quad issues the atomics (the totals of a 512-channel tensor are 512 heavily contended addresses)
  __shared__ float4 fold[256];
  if (R > 1) {
    fold[threadIdx.x] = acc;
    __syncthreads();
    if (rr != 0) return;
    for (int r2 = 1; r2 < R; ++r2) {
      const float4 o = fold[r2 * Q + q];
      acc.x += o.x; acc.y += o.y; acc.z += o.z; acc.w += o.w;
    }
  }
  if (out_img) {
    float* d = out_img + n * out_ld + c;
    atomicAdd(d, acc.x * scale); atomicAdd(d + 1, acc.y * scale); atomicAdd(d + 2, acc.z * scale); atomicAdd(d + 3, acc.w * scale);
  }
  if (out_tot) {
    float* d = out_tot + c;
    atomicAdd(d, acc.x * scale); atomicAdd(d + 1, acc.y * scale); atomicAdd(d + 2, acc.z * scale); atomicAdd(d + 3, acc.w * scale);
  }
}

// ---------------------------------------------------------------- small FP32 GEMM  C = alpha * op(A) op(B) + beta * C
// 64 x 64 tiles, 256 threads, 4 x 4 per thread.  op(A) is M x K, op(B) is K x N; row-major storage with leading dimensions.
// blockIdx.z = batch index (operand z lives at base + z * stride; stride 0 shares an operand across the batch).
__global__ void __launch_bounds__(256) sgemm_kernel(int ta, int tb, int M, int N, int K, float alpha, const float* __restrict__ A,
                                                    long long lda, const float* __restrict__ B, long long ldb, float beta,
                                                    float* __restrict__ C, long long ldc, long long sA = 0, long long sB = 0,
                                                    long long sC = 0) {
  __shared__ float sa[16][64 + 4], sb[16][64 + 4];
  A += (long long)blockIdx.z * sA;
  B += (long long)blockIdx.z * sB;
  C += (long long)blockIdx.z * sC;
  const int m0 = blockIdx.y * 64, n0 = blockIdx.x * 64;
  const int tx = threadIdx.x % 16, ty = threadIdx.x / 16;
  float acc[4][4] = {};
  for (int k0 = 0; k0 < K; k0 += 16) {
    for (int i = threadIdx.x; i < 16 * 64; i += 256) {
      const int kk = i / 64, mm = i % 64;
      const int k = k0 + kk;
      float a = 0.f, b = 0.f;
      if (k < K) {
        const int m = m0 + mm, n = n0 + mm;
        if (m < M) a = ta ? A[(long long)k * lda + m] : A[(long long)m * lda + k];
        if (n < N) b = tb ? B[(long long)n * ldb + k] : B[(long long)k * ldb + n];
      }
      sa[kk][mm] = a;
      sb[kk][mm] = b;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < 16; ++kk) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        a[i] = sa[kk][ty * 4 + i];
        b[i] = sb[kk][tx * 4 + i];
      }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] += a[i] * b[j];
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int m = m0 + ty * 4 + i, n = n0 + tx * 4 + j;
      if (m < M && n < N) {
        float* c = C + (long long)m * ldc + n;
        *c = alpha * acc[i][j] + (beta != 0.f ? beta * *c : 0.f);
      }
    }
}

// dx = dy * silu'(pre)
__global__ void silu_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ pre, float* __restrict__ dx, long long n) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float u = pre[i];
    const float sg = 1.0f / (1.0f + __expf(-u));
    dx[i] = dy[i] * sg * (1.0f + u * (1.0f - sg));
  }
}

// ---------------------------------------------------------------- x_t = a[n] x + b[n] z   (sde.marginal_prob perturbation)
__global__ void perturb_kernel(const float* __restrict__ x, const float* __restrict__ z, const float* __restrict__ a,
                               const float* __restrict__ b, float* __restrict__ out, long long D4) {
  const long long n = blockIdx.y;
  const float an = a[n], bn = b[n];
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < D4; i += (long long)gridDim.x * blockDim.x) {
    const float4 xv = reinterpret_cast<const float4*>(x)[n * D4 + i], zv = reinterpret_cast<const float4*>(z)[n * D4 + i];
    reinterpret_cast<float4*>(out)[n * D4 + i] = make_float4(an * xv.x + bn * zv.x, an * xv.y + bn * zv.y, an * xv.z + bn * zv.z, an * xv.w + bn * zv.w);
  }
}

// ---------------------------------------------------------------- DSM loss and its gradient, one block per sample
// r = score * std[n] + z ;  loss[n] = 0.5 * w[n] * sum(r^2) * norm ;  dscore = gscale * w[n] * norm * std[n] * r
__global__ void dsm_loss_kernel(const float* __restrict__ score, const float* __restrict__ z, const float* __restrict__ std_,
                                const float* __restrict__ w, float* __restrict__ loss, float* __restrict__ dscore, long long D,
                                float norm, float gscale) {
  __shared__ float red[32];
  const long long n = blockIdx.x;
  const float sd = std_[n], wn = w ? w[n] : 1.0f;
  const float gs = gscale * wn * norm * sd;
  float acc = 0.f;
  for (long long i = threadIdx.x; i < D; i += blockDim.x) {
    const float r = score[n * D + i] * sd + z[n * D + i];
    acc += r * r;
    if (dscore) dscore[n * D + i] = gs * r;
  }
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    acc = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.f;
    acc = warp_sum(acc);
    if (threadIdx.x == 0) loss[n] = 0.5f * wn * acc * norm;
  }
}

// ---------------------------------------------------------------- optimizer tail
__global__ void sumsq_kernel(const float* __restrict__ x, long long n, float* __restrict__ out) {
  float s = 0.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) s += x[i] * x[i];
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) atomicAdd(out, s);
}

// AdamW (torch.optim.AdamW semantics, amsgrad off) with the gradient scaled by the clip coefficient
// min(1, max_norm / (sqrt(*sumsq) + 1e-6)) (torch.nn.utils.clip_grad_norm_), then the EMA update
// shadow -= (1 - decay) * (shadow - param) (models/ema.py:43-51).  sumsq == NULL or max_norm < 0: no clipping; ema == NULL: no EMA.
__global__ void adamw_ema_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                                 float* __restrict__ ema, long long n, float lr, float b1, float b2, float eps, float wd, float bc1,
                                 float bc2_sqrt, const float* __restrict__ sumsq, float max_norm, float ema_decay) {
  float clip = 1.0f;
  if (sumsq && max_norm >= 0.f) clip = fminf(1.0f, max_norm / (sqrtf(*sumsq) + 1e-6f));
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float gi = g[i] * clip;
    float pi = p[i] * (1.0f - lr * wd);
    const float mi = b1 * m[i] + (1.0f - b1) * gi;
    const float vi = b2 * v[i] + (1.0f - b2) * gi * gi;
    m[i] = mi;
    v[i] = vi;
    pi -= (lr / bc1) * mi / (sqrtf(vi) / bc2_sqrt + eps);
    p[i] = pi;
    if (ema) ema[i] -= (1.0f - ema_decay) * (ema[i] - pi);
  }
}

__global__ void ema_kernel(float* __restrict__ ema, const float* __restrict__ p, long long n, float decay) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    ema[i] -= (1.0f - decay) * (ema[i] - p[i]);
}

inline int blocks_for(long long n, int threads, int per_sm = 8) {
  long long b = (n + threads - 1) / threads;
  const long long cap = (long long)indm_num_sms() * per_sm;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (int)b;
}

}  // namespace

extern "C" int indm_colsum(const void* x, int dtype, int64_t N, int64_t P, int C, int64_t x_ld, float* out_img, int64_t out_ld,
                           float* out_tot, float scale, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  INDM_CHECK_ARG(x && (out_img || out_tot) && N > 0 && P > 0 && C > 0 && C % 4 == 0 && C <= 4096 && N <= 65535, "colsum: bad arguments");
  if (x_ld == 0) x_ld = C;
  if (out_ld == 0) out_ld = C;
  if (!out_img) {
    // totals only: the rows of all images form one [N * P, C] matrix; few, long column walks instead of (splits x N) CTAs that
    // all add into the same C addresses (1280-way contended atomics made the bias-gradient sums slower than the GEMMs they follow)
    P *= N;
    N = 1;
  }
  const int Q = C / 4;
  int R = 256 / Q;
  if (R < 1) R = 1;
  if (R > P) R = (int)P;
  const int threads = Q * R;
  INDM_CHECK_ARG(threads <= 256 || R == 1, "colsum: C too large");
  long long splits = ((out_img ? 8LL : 4LL) * indm_num_sms() + N - 1) / N;
  const long long maxs = (P + R * 8LL - 1) / (R * 8LL);
  if (splits > maxs) splits = maxs;
  if (splits < 1) splits = 1;
  dim3 grid((unsigned)splits, (unsigned)N);
  if (dtype == INDM_DTYPE_BF16)
    colsum_kernel<__nv_bfloat16><<<grid, threads, 0, stream>>>((const __nv_bfloat16*)x, P, C, x_ld, out_img, out_ld, out_tot, scale);
  else
    colsum_kernel<float><<<grid, threads, 0, stream>>>((const float*)x, P, C, x_ld, out_img, out_ld, out_tot, scale);
  INDM_CHECK_LAUNCH("colsum");
  return INDM_OK;
}

extern "C" int indm_sgemm_f32(int trans_a, int trans_b, int M, int N, int K, float alpha, const float* A, int64_t lda, const float* B,
                              int64_t ldb, float beta, float* C, int64_t ldc, void* stream_) {
  INDM_CHECK_ARG(A && B && C && M > 0 && N > 0 && K > 0, "sgemm: bad arguments");
  dim3 grid((N + 63) / 64, (M + 63) / 64);
  sgemm_kernel<<<grid, 256, 0, (cudaStream_t)stream_>>>(trans_a, trans_b, M, N, K, alpha, A, lda, B, ldb, beta, C, ldc);
  INDM_CHECK_LAUNCH("sgemm");
  return INDM_OK;
}

extern "C" int indm_sgemm_batched_f32(int trans_a, int trans_b, int M, int N, int K, float alpha, const float* A, int64_t lda,
                                      int64_t stride_a, const float* B, int64_t ldb, int64_t stride_b, float beta, float* C, int64_t ldc,
                                      int64_t stride_c, int batch, void* stream_) {
  INDM_CHECK_ARG(A && B && C && M > 0 && N > 0 && K > 0 && batch > 0 && batch <= 65535, "sgemm_batched: bad arguments");
  dim3 grid((N + 63) / 64, (M + 63) / 64, (unsigned)batch);
  sgemm_kernel<<<grid, 256, 0, (cudaStream_t)stream_>>>(trans_a, trans_b, M, N, K, alpha, A, lda, B, ldb, beta, C, ldc, stride_a, stride_b,
                                                        stride_c);
  INDM_CHECK_LAUNCH("sgemm_batched");
  return INDM_OK;
}

extern "C" int indm_silu_bwd_f32(const float* dy, const float* pre, float* dx, int64_t n, void* stream_) {
  INDM_CHECK_ARG(dy && pre && dx && n > 0, "silu_bwd: bad arguments");
  silu_bwd_kernel<<<blocks_for(n, 256), 256, 0, (cudaStream_t)stream_>>>(dy, pre, dx, n);
  INDM_CHECK_LAUNCH("silu_bwd");
  return INDM_OK;
}

extern "C" int indm_perturb_f32(const float* x, const float* z, const float* a, const float* b, float* out, int64_t N, int64_t D,
                                void* stream_) {
  INDM_CHECK_ARG(x && z && a && b && out && N > 0 && D > 0 && D % 4 == 0 && N <= 65535, "perturb: bad arguments");
  dim3 grid((unsigned)blocks_for(D / 4, 256, 2), (unsigned)N);
  perturb_kernel<<<grid, 256, 0, (cudaStream_t)stream_>>>(x, z, a, b, out, D / 4);
  INDM_CHECK_LAUNCH("perturb");
  return INDM_OK;
}

extern "C" int indm_dsm_loss_f32(const float* score, const float* z, const float* std_, const float* w, float* loss, float* dscore,
                                 int64_t N, int64_t D, float norm, float gscale, void* stream_) {
  INDM_CHECK_ARG(score && z && std_ && loss && N > 0 && D > 0, "dsm_loss: bad arguments");
  dsm_loss_kernel<<<(unsigned)N, 256, 0, (cudaStream_t)stream_>>>(score, z, std_, w, loss, dscore, D, norm, gscale);
  INDM_CHECK_LAUNCH("dsm_loss");
  return INDM_OK;
}

extern "C" int indm_sumsq_f32(const float* x, int64_t n, float* out, void* stream_) {
  INDM_CHECK_ARG(x && out && n > 0, "sumsq: bad arguments");
  sumsq_kernel<<<blocks_for(n, 256), 256, 0, (cudaStream_t)stream_>>>(x, n, out);
  INDM_CHECK_LAUNCH("sumsq");
  return INDM_OK;
}

extern "C" int indm_adamw_ema_f32(float* p, const float* g, float* m, float* v, float* ema, int64_t n, float lr, float beta1,
                                  float beta2, float eps, float weight_decay, int64_t step, const float* grad_sumsq, float max_norm,
                                  float ema_decay, void* stream_) {
  INDM_CHECK_ARG(p && g && m && v && n > 0 && step >= 1, "adamw_ema: bad arguments");
  const float bc1 = (float)(1.0 - pow((double)beta1, (double)step));
  const float bc2s = (float)sqrt(1.0 - pow((double)beta2, (double)step));
  adamw_ema_kernel<<<blocks_for(n, 256), 256, 0, (cudaStream_t)stream_>>>(p, g, m, v, ema, n, lr, beta1, beta2, eps, weight_decay, bc1,
                                                                          bc2s, grad_sumsq, max_norm, ema_decay);
  INDM_CHECK_LAUNCH("adamw_ema");
  return INDM_OK;
}

extern "C" int indm_ema_f32(float* ema, const float* p, int64_t n, float decay, void* stream_) {
  INDM_CHECK_ARG(ema && p && n > 0, "ema: bad arguments");
  ema_kernel<<<blocks_for(n, 256), 256, 0, (cudaStream_t)stream_>>>(ema, p, n, decay);
  INDM_CHECK_LAUNCH("ema");
  return INDM_OK;
}
