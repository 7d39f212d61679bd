// Bandwidth-bound kernels of the NCSN++ score network on NHWC activations: GroupNorm statistics / apply (+SiLU,
// +nearest-up / mean-down resampling, +raw copy for the fused skip conv), row softmax, input layout conversion,
// time embedding, small dense layers.  All are HBM/L2-streaming kernels: 128-bit accesses, one fixed channel quad
// per thread so per-channel parameters live in registers, grids sized from the SM count.
#include <cstring>
#include <type_traits>

#include "../../include/indm_b200.h"
#include "common.cuh"
#include "nhwc.cuh"
#include "philox.cuh"

namespace {

// ---------------------------------------------------------------- GroupNorm statistics
// grid (splits, N); block = Q * R threads, Q = C/4 channel quads, R pixel rows in flight.
template <typename TIn>
__global__ void gn_stats_kernel(const TIn* __restrict__ xa, int Ca, const TIn* __restrict__ xb, int Cb, long long P, int G,
                                int R, float* __restrict__ partial) {
  pdl_trigger();
  pdl_wait();

  __shared__ float s_sum[32], s_sq[32];
  const int C = Ca + Cb;
  const int Q = C >> 2;
  const int q = threadIdx.x % Q;
  const int rr = threadIdx.x / Q;
  const long long n = blockIdx.y;
  for (int i = threadIdx.x; i < 32; i += blockDim.x) {
    s_sum[i] = 0.f;
    s_sq[i] = 0.f;
  }
  __syncthreads();
  const long long per = (P + gridDim.x - 1) / gridDim.x;
  const long long p0 = (long long)blockIdx.x * per;
  const long long p1 = min(P, p0 + per);
  const int c = q * 4;
  const TIn* src;
  int ld;
  if (c < Ca) {
    src = xa + n * P * Ca + c;
    ld = Ca;
  } else {
    src = xb + n * P * Cb + (c - Ca);
    ld = Cb;
  }
  float s = 0.f, ss = 0.f;
  if (rr < R) {
    // 4 independent 16-byte loads in flight per thread (HBM latency x bandwidth needs ~40 KB in flight per SM)
    for (long long p = p0 + rr; p < p1; p += 4LL * R) {
      float4 v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const long long pp = p + (long long)u * R;
        v[u] = pp < p1 ? Vec4<TIn>::load(src + pp * ld) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        s += (v[u].x + v[u].y) + (v[u].z + v[u].w);
        ss += (v[u].x * v[u].x + v[u].y * v[u].y) + (v[u].z * v[u].z + v[u].w * v[u].w);
      }
    }
    const int g = c / (C / G);
    atomicAdd(&s_sum[g], s);
    atomicAdd(&s_sq[g], ss);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < G; i += blockDim.x) {
    atomicAdd(&partial[(n * G + i) * 2 + 0], s_sum[i]);
    atomicAdd(&partial[(n * G + i) * 2 + 1], s_sq[i]);
  }
}

// ---------------------------------------------------------------- GroupNorm apply (+SiLU, +resample, +raw copy)
// RES: 0 none, 1 nearest up x2, 2 mean down x2.  grid (splits, N), block = Q * R.
// DROP: the training-mode dropout path is a separate instantiation, so the eval / sampling kernel carries no Philox state
// (71-80 -> 56 registers: 4 instead of 3 resident 256-thread CTAs per SM; measured 3.74 -> 3.82 TB/s over the forward's 95 launches).
// UNR: pixels (independent 8- or 16-byte loads) in flight per thread in the RES == 0 loop (8 measured within 1 % of 4: the
// kernel is bound by per-launch latency on the small feature maps, not by bytes in flight).
template <typename TIn, typename TOut, int RES, bool DROP, int UNR = 4>
__global__ void gn_apply_kernel(const TIn* __restrict__ xa, int Ca, const TIn* __restrict__ xb, int Cb, int H, int W, int G,
                                int R, const float* __restrict__ partial, const float* __restrict__ gamma,
                                const float* __restrict__ beta, float eps, int act,
                                typename std::conditional<std::is_same<TOut, Tf32Out>::value, float, TOut>::type* __restrict__ out,
                                typename std::conditional<std::is_same<TOut, Tf32Out>::value, float, TOut>::type* __restrict__ raw,
                                float drop_p, const unsigned long long* __restrict__ drop_ctl, unsigned drop_stream) {
  pdl_trigger();
  pdl_wait();

  // training-mode dropout after the activation (models/layerspp.py:278): drop_ctl = {seed, enabled} in device memory, so the
  // same launch plan serves eval (enabled = 0) and train forward passes
  const bool dropping = DROP && RES == 0 && drop_ctl != nullptr && drop_p > 0.f && drop_ctl[1] != 0ull;
  const unsigned long long drop_seed = dropping ? drop_ctl[0] : 0ull;
  const int C = Ca + Cb;
  const int Q = C >> 2;
  const int q = threadIdx.x % Q;
  const int rr = threadIdx.x / Q;
  if (rr >= R) return;
  const long long n = blockIdx.y;
  const int c = q * 4;
  const int cpg = C / G;
  const int g = c / cpg;
  const long long P = (long long)H * W;
  const float cnt = (float)((double)P * cpg);
  const float su = partial[(n * G + g) * 2 + 0];
  const float sq = partial[(n * G + g) * 2 + 1];
  const float mean = su / cnt;
  const float var = fmaxf(sq / cnt - mean * mean, 0.f);
  const float rstd = rsqrtf(var + eps);
  const float4 ga = *reinterpret_cast<const float4*>(gamma + c);
  const float4 be = *reinterpret_cast<const float4*>(beta + c);
  const float4 sc = make_float4(ga.x * rstd, ga.y * rstd, ga.z * rstd, ga.w * rstd);
  const float4 sh = make_float4(be.x - mean * sc.x, be.y - mean * sc.y, be.z - mean * sc.z, be.w - mean * sc.w);
  const TIn* src;
  int ld;
  if (c < Ca) {
    src = xa + n * P * Ca + c;
    ld = Ca;
  } else {
    src = xb + n * P * Cb + (c - Ca);
    ld = Cb;
  }
  auto norm = [&](float4 v) {
    float4 y = make_float4(v.x * sc.x + sh.x, v.y * sc.y + sh.y, v.z * sc.z + sh.z, v.w * sc.w + sh.w);
    if (act) {
      if (std::is_same<TOut, __nv_bfloat16>::value) y = make_float4(silu_fast(y.x), silu_fast(y.y), silu_fast(y.z), silu_fast(y.w));
      else y = make_float4(silu_f(y.x), silu_f(y.y), silu_f(y.z), silu_f(y.w));
    }
    return y;
  };
  if (RES == 2) {
    const int Ho = H >> 1, Wo = W >> 1;
    const long long Po = (long long)Ho * Wo;
    const long long per = (Po + gridDim.x - 1) / gridDim.x;
    const long long p0 = (long long)blockIdx.x * per, p1 = min(Po, p0 + per);
    for (long long po = p0 + rr; po < p1; po += R) {
      const int yo = (int)(po / Wo), xo = (int)(po % Wo);
      const long long pi = (long long)(2 * yo) * W + 2 * xo;
      const float4 v00 = Vec4<TIn>::load(src + pi * ld), v01 = Vec4<TIn>::load(src + (pi + 1) * ld);
      const float4 v10 = Vec4<TIn>::load(src + (pi + W) * ld), v11 = Vec4<TIn>::load(src + (pi + W + 1) * ld);
      const float4 a = norm(v00), b = norm(v01), cc = norm(v10), d = norm(v11);
      const float4 y = make_float4(0.25f * ((a.x + b.x) + (cc.x + d.x)), 0.25f * ((a.y + b.y) + (cc.y + d.y)),
                                   0.25f * ((a.z + b.z) + (cc.z + d.z)), 0.25f * ((a.w + b.w) + (cc.w + d.w)));
      Vec4<TOut>::store(out + (n * Po + po) * C + c, y);
      if (raw) {
        const float4 r = make_float4(0.25f * ((v00.x + v01.x) + (v10.x + v11.x)), 0.25f * ((v00.y + v01.y) + (v10.y + v11.y)),
                                     0.25f * ((v00.z + v01.z) + (v10.z + v11.z)), 0.25f * ((v00.w + v01.w) + (v10.w + v11.w)));
        Vec4<TOut>::store(raw + (n * Po + po) * C + c, r);
      }
    }
  } else {
    const long long per = (P + gridDim.x - 1) / gridDim.x;
    const long long p0 = (long long)blockIdx.x * per, p1 = min(P, p0 + per);
    if (RES == 0) {
      // UNR pixels per trip: all loads issued before the first dependent use
      for (long long p = p0 + rr; p < p1; p += (long long)UNR * R) {
        float4 v[UNR];
#pragma unroll
        for (int u = 0; u < UNR; ++u) {
          const long long pp = p + (long long)u * R;
          v[u] = pp < p1 ? Vec4<TIn>::load(src + pp * ld) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int u = 0; u < UNR; ++u) {
          const long long pp = p + (long long)u * R;
          if (pp >= p1) break;
          float4 y = norm(v[u]);
          if (DROP && dropping) {
            const float4 k = dropout_scale4(drop_seed, drop_stream, (unsigned long long)((n * P + pp) * Q + q), drop_p);
            y = make_float4(y.x * k.x, y.y * k.y, y.z * k.z, y.w * k.w);
          }
          Vec4<TOut>::store(out + (n * P + pp) * C + c, y);
          if (raw) Vec4<TOut>::store(raw + (n * P + pp) * C + c, v[u]);
        }
      }
    } else {
      for (long long p = p0 + rr; p < p1; p += R) {
        const float4 v = Vec4<TIn>::load(src + p * ld);
        const float4 y = norm(v);
        const int yi = (int)(p / W), xi = (int)(p % W);
        const long long Wo = 2LL * W;
        const long long po = (long long)(2 * yi) * Wo + 2 * xi;
        const long long base = n * P * 4;
        Vec4<TOut>::store(out + (base + po) * C + c, y);
        Vec4<TOut>::store(out + (base + po + 1) * C + c, y);
        Vec4<TOut>::store(out + (base + po + Wo) * C + c, y);
        Vec4<TOut>::store(out + (base + po + Wo + 1) * C + c, y);
        if (raw) {
          Vec4<TOut>::store(raw + (base + po) * C + c, v);
          Vec4<TOut>::store(raw + (base + po + 1) * C + c, v);
          Vec4<TOut>::store(raw + (base + po + Wo) * C + c, v);
          Vec4<TOut>::store(raw + (base + po + Wo + 1) * C + c, v);
        }
      }
    }
  }
}

// Equal-work variant of gn_apply for RES == 0 (round 2): the (image, pixel) space is cut into gridDim.x contiguous ranges of the
// same length — a grid of exactly (resident CTAs per SM x SMs) CTAs with no tail wave — instead of (splits, N) blocks whose count
// (1280 for 128 images) ran 2.16 waves on 592 resident slots, the last one at 16 % occupancy.  A range may cross image boundaries:
// the per-(image, group) scale / shift are rebuilt when it does.
template <typename TIn, typename TOut, bool DROP, int UNR = 4>
__global__ void __launch_bounds__(256) gn_apply_flat_kernel(const TIn* __restrict__ xa, int Ca, const TIn* __restrict__ xb, int Cb, long long N, long long P,
                                     int G, int R, const float* __restrict__ partial, const float* __restrict__ gamma,
                                     const float* __restrict__ beta, float eps, int act,
                                     typename std::conditional<std::is_same<TOut, Tf32Out>::value, float, TOut>::type* __restrict__ out,
                                     typename std::conditional<std::is_same<TOut, Tf32Out>::value, float, TOut>::type* __restrict__ raw,
                                     float drop_p, const unsigned long long* __restrict__ drop_ctl, unsigned drop_stream, int ppH, int ppW) {
  // ppW > 0: out / raw are padded-pixel buffers (indm_igemm_t.a_pp): pixel (n, y, x) -> row (n (H + 1) + y + 1)(W + 2) + x + 1
  pdl_trigger();
  pdl_wait();
  const bool dropping = DROP && drop_ctl != nullptr && drop_p > 0.f && drop_ctl[1] != 0ull;
  const unsigned long long drop_seed = dropping ? drop_ctl[0] : 0ull;
  const int C = Ca + Cb;
  const int Q = C >> 2;
  const int q = threadIdx.x % Q;
  const int rr = threadIdx.x / Q;
  if (rr >= R) return;
  const int c = q * 4;
  const int cpg = C / G;
  const int g = c / cpg;
  const float cnt = (float)((double)P * cpg);
  const float4 ga = *reinterpret_cast<const float4*>(gamma + c);
  const float4 be = *reinterpret_cast<const float4*>(beta + c);
  const long long total = N * P;
  long long per = (total + gridDim.x - 1) / gridDim.x;
  per = (per + R - 1) / R * R;
  const long long g0 = (long long)blockIdx.x * per, g1 = min(total, g0 + per);
  const bool in_a = c < Ca;
  const int ld = in_a ? Ca : Cb;
  const TIn* base = in_a ? xa + c : xb + (c - Ca);
  for (long long n = g0 / P; n * P < g1; ++n) {
    const float su = partial[(n * G + g) * 2 + 0];
    const float sq = partial[(n * G + g) * 2 + 1];
    const float mean = su / cnt;
    const float var = fmaxf(sq / cnt - mean * mean, 0.f);
    const float rstd = rsqrtf(var + eps);
    const float4 sc = make_float4(ga.x * rstd, ga.y * rstd, ga.z * rstd, ga.w * rstd);
    const float4 sh = make_float4(be.x - mean * sc.x, be.y - mean * sc.y, be.z - mean * sc.z, be.w - mean * sc.w);
    const long long p0 = max(g0, n * P) - n * P, p1 = min(g1, (n + 1) * P) - n * P;      // this image's pixels inside the range
    const TIn* src = base + n * P * ld;
    for (long long p = p0 + rr; p < p1; p += (long long)UNR * R) {
      float4 v[UNR];
#pragma unroll
      for (int u = 0; u < UNR; ++u) {
        const long long pp = p + (long long)u * R;
        v[u] = pp < p1 ? Vec4<TIn>::load(src + pp * ld) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int u = 0; u < UNR; ++u) {
        const long long pp = p + (long long)u * R;
        if (pp >= p1) break;
        float4 y = make_float4(v[u].x * sc.x + sh.x, v[u].y * sc.y + sh.y, v[u].z * sc.z + sh.z, v[u].w * sc.w + sh.w);
        if (act) {
          if (std::is_same<TOut, __nv_bfloat16>::value) y = make_float4(silu_fast(y.x), silu_fast(y.y), silu_fast(y.z), silu_fast(y.w));
          else y = make_float4(silu_f(y.x), silu_f(y.y), silu_f(y.z), silu_f(y.w));
        }
        if (DROP && dropping) {
          const float4 k = dropout_scale4(drop_seed, drop_stream, (unsigned long long)((n * P + pp) * Q + q), drop_p);
          y = make_float4(y.x * k.x, y.y * k.y, y.z * k.z, y.w * k.w);
        }
        const long long orow = ppW ? (n * (ppH + 1) + pp / ppW + 1) * (ppW + 2) + pp % ppW + 1 : n * P + pp;
        Vec4<TOut>::store(out + orow * C + c, y);
        if (raw) Vec4<TOut>::store(raw + orow * C + c, v[u]);
      }
    }
  }
}

// Streaming variant of gn_apply for a single dense source with BF16 output (round 2).  The register kernels above were bound by
// instruction issue (~90 instructions per 4 channels: 64-bit index arithmetic, range checks, 8-byte stores), not by HBM: their
// fp32-input and bf16-input forms took the same time.  Here (1) one producer lane per CTA keeps a ring of STAGES bulk copies
// (cp.async.bulk, ~16 KB each) in flight into shared memory, so no thread waits on a global load; (2) eight consumer warps handle
// 8 channels per thread (one 16-byte store), with all 64-bit arithmetic hoisted to once per chunk; an image boundary inside a chunk
// splits it into segments with constant scale / shift.  Same equal-work pixel ranges as the flat kernel.
constexpr int kGnStreamMaxStages = 8;
template <typename TIn>
__device__ __forceinline__ void gn_load8(const unsigned char* p, float (&v)[8]);
template <>
__device__ __forceinline__ void gn_load8<float>(const unsigned char* p, float (&v)[8]) {
  const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 16);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
template <>
__device__ __forceinline__ void gn_load8<__nv_bfloat16>(const unsigned char* p, float (&v)[8]) {
  const uint4 r = *reinterpret_cast<const uint4*>(p);
  const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    v[2 * i] = __uint_as_float(w[i] << 16);
    v[2 * i + 1] = __uint_as_float(w[i] & 0xFFFF0000u);
  }
}

template <typename TIn, bool DROP, bool PP, bool RAW>
__global__ void __launch_bounds__(288) gn_apply_stream_kernel(const TIn* __restrict__ x, int C, long long N, int P, int G, int R, int chunk_px,
                                       int stages, const float* __restrict__ partial, const float* __restrict__ gamma,
                                       const float* __restrict__ beta, float eps, int act, __nv_bfloat16* __restrict__ out,
                                       __nv_bfloat16* __restrict__ raw, float drop_p, const unsigned long long* __restrict__ drop_ctl,
                                       unsigned drop_stream, int ppH, int ppW, int interleave) {
  // interleave: chunk j goes to CTA j % gridDim.x, so at any moment the grid reads (and writes) one contiguous window of
  // gridDim.x chunks instead of gridDim.x far-apart streams; N * P < 2^31 (host check)
  extern __shared__ __align__(128) unsigned char gn_smem[];
  uint64_t* full = reinterpret_cast<uint64_t*>(gn_smem);
  uint64_t* empty = full + kGnStreamMaxStages;
  unsigned char* ring = gn_smem + 128;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t px_bytes = (uint32_t)C * (uint32_t)sizeof(TIn);
  const uint32_t chunk_bytes = (uint32_t)chunk_px * px_bytes;
  const long long total = N * P;
  long long per = (total + gridDim.x - 1) / gridDim.x;
  per = (per + chunk_px - 1) / chunk_px * chunk_px;
  const long long g0 = interleave ? (long long)blockIdx.x * chunk_px : (long long)blockIdx.x * per;
  const long long g1 = interleave ? total : min(total, g0 + per);
  const long long gstep = interleave ? (long long)gridDim.x * chunk_px : (long long)chunk_px;
  const int nchunks = g1 > g0 ? (int)((g1 - g0 + gstep - 1) / gstep) : 0;
  if (tid == 0) {
    for (int s = 0; s < stages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 8);
    }
    fence_mbar_init();
  }
  __syncthreads();
  pdl_trigger();
  pdl_wait();
  if (warp == 8) {
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      for (int k = 0; k < nchunks; ++k) {
        mbar_wait(&empty[s], ph ^ 1u);
        const long long gp0 = g0 + (long long)k * gstep;
        const uint32_t bytes = (uint32_t)min((long long)chunk_px, g1 - gp0) * px_bytes;
        mbar_arrive_expect_tx(&full[s], bytes);
        bulk_load_1d(ring + (size_t)s * chunk_bytes, reinterpret_cast<const unsigned char*>(x) + gp0 * px_bytes, bytes, &full[s]);
        if (++s == stages) { s = 0; ph ^= 1u; }
      }
    }
    return;
  }
  const bool dropping = DROP && drop_ctl != nullptr && drop_p > 0.f && drop_ctl[1] != 0ull;
  const unsigned long long drop_seed = dropping ? drop_ctl[0] : 0ull;
  const int Q8 = C >> 3;
  const int q = tid % Q8;
  const int rr = tid / Q8;            // < R = 256 / Q8 by construction
  const int c = q * 8;
  const int cpg = C / G;
  const int ga_ = c / cpg, gb_ = (c + 4) / cpg;
  const float cnt = (float)((double)P * cpg);
  float gam[8], bet[8], sc[8], sh[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    gam[i] = gamma[c + i];
    bet[i] = beta[c + i];
    sc[i] = sh[i] = 0.f;
  }
  long long n = nchunks ? g0 / P : 0;            // image of the next pixel; n_end: first pixel of image n + 1
  long long n_end = (n + 1) * (long long)P;
  long long n_have = -1;
  int s = 0;
  uint32_t ph = 0;
  for (int k = 0; k < nchunks; ++k) {
    mbar_wait(&full[s], ph);
    const long long gp0 = g0 + (long long)k * gstep;
    const int npx = (int)min((long long)chunk_px, g1 - gp0);
    if (interleave && gp0 >= n_end) {
      n = (long long)((int)gp0 / P);
      n_end = (n + 1) * (long long)P;
    }
    const unsigned char* src = ring + (size_t)s * chunk_bytes + (uint32_t)c * (uint32_t)sizeof(TIn);
    __nv_bfloat16* outc = out + gp0 * C + c;
    __nv_bfloat16* rawc = RAW ? raw + gp0 * C + c : nullptr;
    int r0 = 0;
    while (r0 < npx) {
      const int r1 = (int)min((long long)npx, n_end - gp0);       // pixels [r0, r1) of this chunk belong to image n
      if (n_have != n) {
        n_have = n;
        const float2 pa = *reinterpret_cast<const float2*>(partial + (n * G + ga_) * 2);
        const float2 pb = *reinterpret_cast<const float2*>(partial + (n * G + gb_) * 2);
        const float mean_a = pa.x / cnt, mean_b = pb.x / cnt;
        const float rstd_a = rsqrtf(fmaxf(pa.y / cnt - mean_a * mean_a, 0.f) + eps);
        const float rstd_b = rsqrtf(fmaxf(pb.y / cnt - mean_b * mean_b, 0.f) + eps);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float mean = i < 4 ? mean_a : mean_b, rstd = i < 4 ? rstd_a : rstd_b;
          sc[i] = gam[i] * rstd;
          sh[i] = bet[i] - mean * sc[i];
        }
      }
      const int pp0 = PP ? (int)(gp0 - n * (long long)P) : 0;      // pixel index inside image n of chunk row 0 (may be negative)
      const long long pprow0 = PP ? n * (ppH + 1) + 1 : 0;
      int r = r0 + (rr - r0 % R + R) % R;
#pragma unroll 2
      for (; r < r1; r += R) {
        float v[8], y[8];
        gn_load8<TIn>(src + (uint32_t)r * px_bytes, v);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          y[i] = fmaf(v[i], sc[i], sh[i]);
          if (act) y[i] = silu_fast(y[i]);
        }
        if (DROP && dropping) {
          const unsigned long long qi = (unsigned long long)(gp0 + r) * (unsigned long long)(2 * Q8) + (unsigned long long)(2 * q);
          const float4 k0 = dropout_scale4(drop_seed, drop_stream, qi, drop_p), k1 = dropout_scale4(drop_seed, drop_stream, qi + 1, drop_p);
          y[0] *= k0.x; y[1] *= k0.y; y[2] *= k0.z; y[3] *= k0.w; y[4] *= k1.x; y[5] *= k1.y; y[6] *= k1.z; y[7] *= k1.w;
        }
        uint4 o;
        o.x = pack_bf16x2(y[0], y[1]); o.y = pack_bf16x2(y[2], y[3]); o.z = pack_bf16x2(y[4], y[5]); o.w = pack_bf16x2(y[6], y[7]);
        size_t off;
        if (PP) {
          const int pp = pp0 + r, yy = pp / ppW, xx = pp - yy * ppW;
          off = (size_t)((pprow0 + yy) * (ppW + 2) + xx + 1) * C - (size_t)gp0 * C;
        } else {
          off = (size_t)((uint32_t)r * (uint32_t)C);
        }
        *reinterpret_cast<uint4*>(outc + off) = o;
        if (RAW) {
          uint4 w;
          w.x = pack_bf16x2(v[0], v[1]); w.y = pack_bf16x2(v[2], v[3]); w.z = pack_bf16x2(v[4], v[5]); w.w = pack_bf16x2(v[6], v[7]);
          *reinterpret_cast<uint4*>(rawc + off) = w;
        }
      }
      r0 = r1;
      if (gp0 + r0 >= n_end) {
        ++n;
        n_end += P;
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty[s]);
    if (++s == stages) { s = 0; ph ^= 1u; }
  }
}

// ---------------------------------------------------------------- softmax over rows, one warp per row
template <typename TOut>
__global__ void softmax_rows_kernel(const float* __restrict__ s, TOut* __restrict__ out, long long rows, int cols, int tf32) {
  pdl_trigger();
  pdl_wait();

  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  const float* src = s + row * cols;
  float m = -INFINITY;
  for (int i = lane; i < cols; i += 32) m = fmaxf(m, src[i]);
  m = warp_max(m);
  float sum = 0.f;
  for (int i = lane; i < cols; i += 32) sum += __expf(src[i] - m);
  sum = warp_sum(sum);
  const float inv = 1.0f / sum;
  for (int i = lane; i < cols; i += 32) {
    float v = __expf(src[i] - m) * inv;
    if (tf32) v = tf32_operand(v);
    out[row * cols + i] = (TOut)v;
  }
}

// ---------------------------------------------------------------- network input NCHW fp32 -> NHWC (padded channels)
template <typename TOut>
__global__ void prep_input_kernel(const float* __restrict__ x, TOut* __restrict__ out, long long N, int C, int HW, int cpad,
                                  float mul, float add, int act, int tf32) {
  pdl_trigger();
  pdl_wait();

  const long long total = N * HW * cpad;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % cpad);
    const long long pix = i / cpad;
    const long long n = pix / HW;
    const int p = (int)(pix % HW);
    float v = 0.f;
    if (c < C) {
      v = x[(n * C + c) * HW + p] * mul + add;
      if (act == 1) v = sinf(6.283185307179586f * v) * 0.15915494309189535f;
      if (tf32) v = tf32_operand(v);
    }
    out[i] = (TOut)v;
  }
}

// ---------------------------------------------------------------- time embedding
__global__ void time_embedding_kernel(const float* __restrict__ time_cond, const float* __restrict__ sched,
                                      const int32_t* __restrict__ step, int sched_ld, int sched_col,
                                      const float* __restrict__ freqs, int kind, long long N, int dim, float* __restrict__ out) {
  pdl_trigger();
  pdl_wait();

  const int half = dim >> 1;
  const long long total = N * half;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long n = i / half;
    const int j = (int)(i % half);
    const float tc = sched ? sched[(long long)(step ? *step : 0) * sched_ld + sched_col] : time_cond[n];
    float arg;
    if (kind == 0) {
      // models/layers.py:515-529: emb_j = exp(-j * ln(10000) / (half - 1)), arg = timesteps * emb_j
      const float e = logf(10000.0f) / (float)(half - 1);
      arg = tc * expf((float)j * -e);
    } else {
      // models/layerspp.py:52-54 with models/ncsnpp.py:258: x = log(sigma); x * W * 2 * pi
      arg = logf(tc) * freqs[j] * 2.0f * 3.14159265358979323846f;
    }
    out[n * dim + j] = sinf(arg);
    out[n * dim + half + j] = cosf(arg);
  }
}

// ---------------------------------------------------------------- small dense layer, one warp per output column
// out[n][o] = bias[o] + sum_k f(in[n][k]) w[o][k].  K <= 1024, K % 128 == 0 handled by the 8-register path.
template <int KV>  // float4 per lane = K / 128
__global__ void linear_kernel(const float* __restrict__ in, const float* __restrict__ w, const float* __restrict__ bias,
                              void* __restrict__ out, long long N, int K, int O, int act_in, int act_out, int out_dtype) {
  pdl_trigger();
  pdl_wait();

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int o = blockIdx.x * (blockDim.x >> 5) + warp;
  if (o >= O) return;
  float4 wr[KV];
#pragma unroll
  for (int i = 0; i < KV; ++i) wr[i] = *reinterpret_cast<const float4*>(w + (long long)o * K + (i * 32 + lane) * 4);
  const float b = bias ? bias[o] : 0.f;
  for (long long n = blockIdx.y; n < N; n += gridDim.y) {
    float acc = 0.f;
#pragma unroll
    for (int i = 0; i < KV; ++i) {
      float4 v = *reinterpret_cast<const float4*>(in + n * K + (i * 32 + lane) * 4);
      if (act_in) v = make_float4(silu_f(v.x), silu_f(v.y), silu_f(v.z), silu_f(v.w));
      acc += (v.x * wr[i].x + v.y * wr[i].y) + (v.z * wr[i].z + v.w * wr[i].w);
    }
    acc = warp_sum(acc);
    if (lane == 0) {
      float v = acc + b;
      if (act_out) v = silu_f(v);
      if (out_dtype == INDM_DTYPE_BF16) reinterpret_cast<__nv_bfloat16*>(out)[n * O + o] = __float2bfloat16_rn(v);
      else reinterpret_cast<float*>(out)[n * O + o] = (out_dtype == INDM_DTYPE_TF32) ? tf32_operand(v) : v;
    }
  }
}

// any K: lanes stride over k (used for the K = 64 conditioning projections of the flow)
__global__ void linear_generic_kernel(const float* __restrict__ in, const float* __restrict__ w, const float* __restrict__ bias,
                                      void* __restrict__ out, long long N, int K, int O, int act_in, int act_out, int out_dtype) {
  pdl_trigger();
  pdl_wait();

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int o = blockIdx.x * (blockDim.x >> 5) + warp;
  if (o >= O) return;
  const float b = bias ? bias[o] : 0.f;
  for (long long n = blockIdx.y; n < N; n += gridDim.y) {
    float acc = 0.f;
    for (int k = lane; k < K; k += 32) {
      float v = in[n * K + k];
      if (act_in) v = silu_f(v);
      acc += v * w[(long long)o * K + k];
    }
    acc = warp_sum(acc);
    if (lane == 0) {
      float v = acc + b;
      if (act_out) v = silu_f(v);
      if (out_dtype == INDM_DTYPE_BF16) reinterpret_cast<__nv_bfloat16*>(out)[n * O + o] = __float2bfloat16_rn(v);
      else reinterpret_cast<float*>(out)[n * O + o] = (out_dtype == INDM_DTYPE_TF32) ? tf32_operand(v) : v;
    }
  }
}

// ---------------------------------------------------------------- FIR resampling on NHWC (4-tap separable kernel)
// MODE 1: up x2 pad (2,1); MODE 2: down x2 pad (1,1); MODE 3: up=down=1 pad (2,2); MODE 4: up=down=1 pad (1,1) (transpose of 3).
// out[oy,ox] = sum_{i,j} xp[oy*down + i, ox*down + j] * kf[i][j], kf = flipped k, xp = zero-inserted + padded input.
// Separable evaluation along a strip of output rows: thread = (output column ox, channel quad q); it walks RS consecutive output rows
// keeping the four horizontally filtered rows h(uy) = sum_j kf[j] x[iy(uy)][ix(ox, j)] the vertical taps need in registers, so an
// input pixel is loaded once per output COLUMN tap, not once per 2-D tap: 4 (plain), 8 (down x2) or ~1 (up x2) 16-byte loads per
// output quad instead of 16 / 16 / 4, no runtime division, channel-contiguous (fully coalesced) accesses.  grid = (column tiles x
// channel tiles, row strips, N).
// four channels per thread: kept for the down x2 mode (few output pixels: the eight-channel variant halves the thread count
// and ran 24 -> 33 us there)
template <typename TIn, typename TOut, int MODE>
__global__ void __launch_bounds__(256) fir_nhwc4_kernel(const TIn* __restrict__ x, typename std::conditional<std::is_same<TOut, Tf32Out>::value, float, TOut>::type* __restrict__ y,
                                long long N, int H, int W, int C, float k0, float k1, float k2, float k3, int qb_log2, int RS, int ctiles) {
  pdl_trigger();
  pdl_wait();

  constexpr int UP = MODE == 1 ? 2 : 1;
  constexpr int DOWN = MODE == 2 ? 2 : 1;
  constexpr int PAD0 = MODE == 1 ? 2 : ((MODE == 2 || MODE == 4) ? 1 : 2);
  const int Ho = MODE == 1 ? 2 * H : (MODE == 2 ? H / 2 : (MODE == 4 ? H - 1 : H + 1));
  const int Wo = MODE == 1 ? 2 * W : (MODE == 2 ? W / 2 : (MODE == 4 ? W - 1 : W + 1));
  const int Q = C >> 2;
  const float kf[4] = {k3, k2, k1, k0};  // flipped
  const int qb = 1 << qb_log2;                                   // channel quads per CTA (power of two)
  const int ct = blockIdx.x % ctiles, xt = blockIdx.x / ctiles;
  const int q = ct * qb + (threadIdx.x & (qb - 1));
  const int ox = xt * (256 >> qb_log2) + (threadIdx.x >> qb_log2);
  if (q >= Q || ox >= Wo) return;
  const long long n = blockIdx.z;
  const int oy_begin = blockIdx.y * RS;
  const int oy_end = oy_begin + RS < Ho ? oy_begin + RS : Ho;
  const TIn* xin = x + n * (long long)H * W * C + q * 4;
  // column taps of this thread: input column and weight per j (invalid taps get weight 0 and a safe column)
  int ixs[4];
  float wj[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int ux = ox * DOWN + j - PAD0;
    const bool ok = ux >= 0 && (ux % UP) == 0 && (ux / UP) < W;
    ixs[j] = ok ? ux / UP : 0;
    wj[j] = ok ? kf[j] : 0.f;
  }
  auto hrow = [&](int uy) -> float4 {
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
    if (uy < 0 || (uy % UP) != 0) return a;
    const int iy = uy / UP;
    if (iy >= H) return a;
    const TIn* row = xin + (long long)iy * W * C;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (wj[j] != 0.f) {
        const float4 v = Vec4<TIn>::load(row + (long long)ixs[j] * C);
        a.x += v.x * wj[j]; a.y += v.y * wj[j]; a.z += v.z * wj[j]; a.w += v.w * wj[j];
      }
    }
    return a;
  };
  int u0 = oy_begin * DOWN - PAD0;
  float4 h0 = hrow(u0), h1 = hrow(u0 + 1), h2 = hrow(u0 + 2), h3 = hrow(u0 + 3);
  auto* yout = y + n * (long long)Ho * Wo * C + (long long)ox * C + q * 4;
  for (int oy = oy_begin; oy < oy_end; ++oy) {
    float4 acc;
    acc.x = h0.x * kf[0] + h1.x * kf[1] + h2.x * kf[2] + h3.x * kf[3];
    acc.y = h0.y * kf[0] + h1.y * kf[1] + h2.y * kf[2] + h3.y * kf[3];
    acc.z = h0.z * kf[0] + h1.z * kf[1] + h2.z * kf[2] + h3.z * kf[3];
    acc.w = h0.w * kf[0] + h1.w * kf[1] + h2.w * kf[2] + h3.w * kf[3];
    Vec4<TOut>::store(yout + (long long)oy * Wo * C, acc);
    if (oy + 1 < oy_end) {
      if (DOWN == 1) {
        h0 = h1; h1 = h2; h2 = h3; h3 = hrow(u0 + 4);
        u0 += 1;
      } else {
        h0 = h2; h1 = h3; h2 = hrow(u0 + 4); h3 = hrow(u0 + 5);
        u0 += 2;
      }
    }
  }
}

template <typename TIn, typename TOut, int MODE>
__global__ void __launch_bounds__(256) fir_nhwc_kernel(const TIn* __restrict__ x, typename std::conditional<std::is_same<TOut, Tf32Out>::value, float, TOut>::type* __restrict__ y,
                                long long N, int H, int W, int C, float k0, float k1, float k2, float k3, int qb_log2, int RS, int ctiles) {
  pdl_trigger();
  pdl_wait();

  // a thread owns EIGHT channels (one 16-byte load of a bf16 input, two of an fp32 input; one 16-byte store of a bf16 output): with
  // four, the bf16 variants moved 8 bytes per load / store instruction and ran slower than the fp32 ones on half the bytes
  constexpr int UP = MODE == 1 ? 2 : 1;
  constexpr int DOWN = MODE == 2 ? 2 : 1;
  constexpr int PAD0 = MODE == 1 ? 2 : ((MODE == 2 || MODE == 4) ? 1 : 2);
  const int Ho = MODE == 1 ? 2 * H : (MODE == 2 ? H / 2 : (MODE == 4 ? H - 1 : H + 1));
  const int Wo = MODE == 1 ? 2 * W : (MODE == 2 ? W / 2 : (MODE == 4 ? W - 1 : W + 1));
  const int Q = C >> 3;                                          // channel octets
  const float kf[4] = {k3, k2, k1, k0};  // flipped
  const int qb = 1 << qb_log2;                                   // channel octets per CTA (power of two)
  const int ct = blockIdx.x % ctiles, xt = blockIdx.x / ctiles;
  const int q = ct * qb + (threadIdx.x & (qb - 1));
  const int ox = xt * (256 >> qb_log2) + (threadIdx.x >> qb_log2);
  if (q >= Q || ox >= Wo) return;
  const long long n = blockIdx.z;
  const int oy_begin = blockIdx.y * RS;
  const int oy_end = oy_begin + RS < Ho ? oy_begin + RS : Ho;
  const TIn* xin = x + n * (long long)H * W * C + q * 8;
  int ixs[4];
  float wj[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int ux = ox * DOWN + j - PAD0;
    const bool ok = ux >= 0 && (ux % UP) == 0 && (ux / UP) < W;
    ixs[j] = ok ? ux / UP : 0;
    wj[j] = ok ? kf[j] : 0.f;
  }
  struct F8 { float4 a, b; };
  auto hrow = [&](int uy) -> F8 {
    F8 r;
    r.a = make_float4(0.f, 0.f, 0.f, 0.f);
    r.b = r.a;
    if (uy < 0 || (uy % UP) != 0) return r;
    const int iy = uy / UP;
    if (iy >= H) return r;
    const TIn* row = xin + (long long)iy * W * C;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (wj[j] != 0.f) {
        const TIn* p = row + (long long)ixs[j] * C;
        const float4 va = Vec4<TIn>::load(p), vb = Vec4<TIn>::load(p + 4);
        r.a.x += va.x * wj[j]; r.a.y += va.y * wj[j]; r.a.z += va.z * wj[j]; r.a.w += va.w * wj[j];
        r.b.x += vb.x * wj[j]; r.b.y += vb.y * wj[j]; r.b.z += vb.z * wj[j]; r.b.w += vb.w * wj[j];
      }
    }
    return r;
  };
  int u0 = oy_begin * DOWN - PAD0;
  F8 h0 = hrow(u0), h1 = hrow(u0 + 1), h2 = hrow(u0 + 2), h3 = hrow(u0 + 3);
  auto* yout = y + n * (long long)Ho * Wo * C + (long long)ox * C + q * 8;
  for (int oy = oy_begin; oy < oy_end; ++oy) {
    float4 ya, yb;
    ya.x = h0.a.x * kf[0] + h1.a.x * kf[1] + h2.a.x * kf[2] + h3.a.x * kf[3];
    ya.y = h0.a.y * kf[0] + h1.a.y * kf[1] + h2.a.y * kf[2] + h3.a.y * kf[3];
    ya.z = h0.a.z * kf[0] + h1.a.z * kf[1] + h2.a.z * kf[2] + h3.a.z * kf[3];
    ya.w = h0.a.w * kf[0] + h1.a.w * kf[1] + h2.a.w * kf[2] + h3.a.w * kf[3];
    yb.x = h0.b.x * kf[0] + h1.b.x * kf[1] + h2.b.x * kf[2] + h3.b.x * kf[3];
    yb.y = h0.b.y * kf[0] + h1.b.y * kf[1] + h2.b.y * kf[2] + h3.b.y * kf[3];
    yb.z = h0.b.z * kf[0] + h1.b.z * kf[1] + h2.b.z * kf[2] + h3.b.z * kf[3];
    yb.w = h0.b.w * kf[0] + h1.b.w * kf[1] + h2.b.w * kf[2] + h3.b.w * kf[3];
    auto* dst = yout + (long long)oy * Wo * C;
    Vec4<TOut>::store(dst, ya);
    Vec4<TOut>::store(dst + 4, yb);
    if (oy + 1 < oy_end) {
      if (DOWN == 1) {
        h0 = h1; h1 = h2; h2 = h3; h3 = hrow(u0 + 4);
        u0 += 1;
      } else {
        h0 = h2; h1 = h3; h2 = hrow(u0 + 4); h3 = hrow(u0 + 5);
        u0 += 2;
      }
    }
  }
}

}  // namespace

extern "C" int indm_gn_stats(const void* xa, int Ca, const void* xb, int Cb, int in_dtype, int64_t N, int64_t P, int G,
                             float* partial, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (!xb) Cb = 0;
  const int C = Ca + Cb;
  INDM_CHECK_ARG(xa && partial && N > 0 && P > 0, "gn_stats: bad arguments");
  INDM_CHECK_ARG(G >= 1 && G <= 32 && C % G == 0 && (C / G) % 4 == 0 && Ca % 4 == 0 && C / 4 <= 1024,
                 "gn_stats: need G <= 32, (C/G) %% 4 == 0 (C=%d G=%d)", C, G);
  INDM_CHECK_ARG(N <= 65535, "gn_stats: N too large for grid.y");
  const GnGeom g = gn_geom(C, P, N);
  dim3 grid(g.splits, (unsigned)N);
  if (in_dtype == INDM_DTYPE_F32)
    indm_launch_pdl(gn_stats_kernel<float>, grid, dim3(g.threads), 0, stream, (const float*)xa, Ca, (const float*)xb, Cb, P, G, g.R, partial);
  else if (in_dtype == INDM_DTYPE_BF16)
    indm_launch_pdl(gn_stats_kernel<__nv_bfloat16>, grid, dim3(g.threads), 0, stream, (const __nv_bfloat16*)xa, Ca, (const __nv_bfloat16*)xb, Cb, P,
                                                                   G, g.R, partial);
  else
    INDM_CHECK_ARG(false, "gn_stats: in_dtype must be F32 or BF16");
  INDM_CHECK_LAUNCH("gn_stats");
  return INDM_OK;
}

// Register-path sibling of the stream kernel: same 8-channels-per-thread inner loop with the index arithmetic hoisted, loads
// straight from global (two rows = 2 x 32 B per thread in flight), no shared memory — so its CTAs can start under the tail of the
// preceding convolution (programmatic dependent launch), which the 64 KB ring of the stream kernel cannot.
template <typename TIn, bool DROP, bool PP, bool RAW>
__global__ void __launch_bounds__(256) gn_apply_flat8_kernel(const TIn* __restrict__ xa, int Ca, const TIn* __restrict__ xb, int Cb, long long N, int P, int G, int R,
                                      const float* __restrict__ partial, const float* __restrict__ gamma, const float* __restrict__ beta,
                                      float eps, int act, __nv_bfloat16* __restrict__ out, __nv_bfloat16* __restrict__ raw, float drop_p,
                                      const unsigned long long* __restrict__ drop_ctl, unsigned drop_stream, int ppH, int ppW) {
  pdl_trigger();
  pdl_wait();
  const bool dropping = DROP && drop_ctl != nullptr && drop_p > 0.f && drop_ctl[1] != 0ull;
  const unsigned long long drop_seed = dropping ? drop_ctl[0] : 0ull;
  const int C = Ca + Cb;
  const int Q8 = C >> 3;
  const int q = threadIdx.x % Q8;
  const int rr = threadIdx.x / Q8;
  if (rr >= R) return;
  const int c = q * 8;
  const int cpg = C / G;
  const int ga_ = c / cpg, gb_ = (c + 4) / cpg;
  const float cnt = (float)((double)P * cpg);
  const bool in_a = c < Ca;                       // channel concat of two sources: this thread's 8 channels live in one of them
  const uint32_t px_bytes = (uint32_t)(in_a ? Ca : Cb) * (uint32_t)sizeof(TIn);
  const unsigned char* xsrc = reinterpret_cast<const unsigned char*>(in_a ? xa + c : xb + (c - Ca));
  const long long total = N * P;
  long long per = (total + gridDim.x - 1) / gridDim.x;
  per = (per + R - 1) / R * R;
  const long long g0 = (long long)blockIdx.x * per, g1 = min(total, g0 + per);
  for (long long n = g0 / P; n * P < g1; ++n) {
    float sc[8], sh[8];
    {
      const float2 pa = *reinterpret_cast<const float2*>(partial + (n * G + ga_) * 2);
      const float2 pb = *reinterpret_cast<const float2*>(partial + (n * G + gb_) * 2);
      const float mean_a = pa.x / cnt, mean_b = pb.x / cnt;
      const float rstd_a = rsqrtf(fmaxf(pa.y / cnt - mean_a * mean_a, 0.f) + eps);
      const float rstd_b = rsqrtf(fmaxf(pb.y / cnt - mean_b * mean_b, 0.f) + eps);
      const float4 g0v = *reinterpret_cast<const float4*>(gamma + c), g1v = *reinterpret_cast<const float4*>(gamma + c + 4);
      const float4 b0v = *reinterpret_cast<const float4*>(beta + c), b1v = *reinterpret_cast<const float4*>(beta + c + 4);
      const float gam[8] = {g0v.x, g0v.y, g0v.z, g0v.w, g1v.x, g1v.y, g1v.z, g1v.w};
      const float bet[8] = {b0v.x, b0v.y, b0v.z, b0v.w, b1v.x, b1v.y, b1v.z, b1v.w};
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float mean = i < 4 ? mean_a : mean_b, rstd = i < 4 ? rstd_a : rstd_b;
        sc[i] = gam[i] * rstd;
        sh[i] = bet[i] - mean * sc[i];
      }
    }
    const long long p0 = max(g0, n * P), p1 = min(g1, (n + 1) * P);       // this image's pixels inside the range (global indices)
    const int npx = (int)(p1 - p0);
    const unsigned char* src = xsrc + p0 * px_bytes;
    __nv_bfloat16* outc = out + p0 * C + c;
    __nv_bfloat16* rawc = RAW ? raw + p0 * C + c : nullptr;
    const int pp0 = PP ? (int)(p0 - n * P) : 0;
    const long long pprow0 = PP ? n * (ppH + 1) + 1 : 0;
    for (int r = rr; r < npx; r += 2 * R) {
      float v[2][8];
      const bool two = r + R < npx;
      gn_load8<TIn>(src + (size_t)r * px_bytes, v[0]);
      if (two) gn_load8<TIn>(src + (size_t)(r + R) * px_bytes, v[1]);
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        if (u == 1 && !two) break;
        const int ru = r + u * R;
        float y[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          y[i] = fmaf(v[u][i], sc[i], sh[i]);
          if (act) y[i] = silu_fast(y[i]);
        }
        if (DROP && dropping) {
          const unsigned long long qi = (unsigned long long)(p0 + ru) * (unsigned long long)(2 * Q8) + (unsigned long long)(2 * q);
          const float4 k0 = dropout_scale4(drop_seed, drop_stream, qi, drop_p), k1 = dropout_scale4(drop_seed, drop_stream, qi + 1, drop_p);
          y[0] *= k0.x; y[1] *= k0.y; y[2] *= k0.z; y[3] *= k0.w; y[4] *= k1.x; y[5] *= k1.y; y[6] *= k1.z; y[7] *= k1.w;
        }
        uint4 o;
        o.x = pack_bf16x2(y[0], y[1]); o.y = pack_bf16x2(y[2], y[3]); o.z = pack_bf16x2(y[4], y[5]); o.w = pack_bf16x2(y[6], y[7]);
        size_t off;
        if (PP) {
          const int pp = pp0 + ru, yy = pp / ppW, xx = pp - yy * ppW;
          off = (size_t)((pprow0 + yy) * (ppW + 2) + xx + 1) * C - (size_t)p0 * C;
        } else {
          off = (size_t)((uint32_t)ru * (uint32_t)C);
        }
        *reinterpret_cast<uint4*>(outc + off) = o;
        if (RAW) {
          uint4 w;
          w.x = pack_bf16x2(v[u][0], v[u][1]); w.y = pack_bf16x2(v[u][2], v[u][3]); w.z = pack_bf16x2(v[u][4], v[u][5]);
          w.w = pack_bf16x2(v[u][6], v[u][7]);
          *reinterpret_cast<uint4*>(rawc + off) = w;
        }
      }
    }
  }
}

template <typename TIn, bool DROP, bool PP, bool RAW>
static int gn_apply_flat8_go(dim3 grid, int threads, cudaStream_t stream, const void* xa, int Ca, const void* xb, int Cb, int64_t N, int P, int G, int R,
                             const float* partial, const float* gamma, const float* beta, float eps, int act, void* out, void* raw,
                             float drop_p, const unsigned long long* drop_ctl, unsigned drop_stream, int ppH, int ppW) {
  indm_launch_pdl(gn_apply_flat8_kernel<TIn, DROP, PP, RAW>, grid, dim3(threads), 0, stream, (const TIn*)xa, Ca, (const TIn*)xb, Cb, (long long)N, P, G, R, partial,
                  gamma, beta, eps, act, (__nv_bfloat16*)out, (__nv_bfloat16*)raw, drop_p, drop_ctl, drop_stream, ppH, ppW);
  INDM_CHECK_LAUNCH("gn_apply (flat8)");
  return INDM_OK;
}

template <typename TIn>
static int gn_apply_flat8_launch(const void* xa, int Ca, const void* xb, int Cb, int64_t N, int H, int W, int G, const float* partial, const float* gamma,
                                 const float* beta, float eps, int act, void* out, void* raw, float drop_p,
                                 const unsigned long long* drop_ctl, unsigned drop_stream, cudaStream_t stream, bool pp) {
  static const int per_sm = []() { const char* e = getenv("INDM_GN_CTAS_PER_SM"); const int v = e ? atoi(e) : 0; return v > 0 ? v : 3; }();   // 72 - 120 registers: 3 resident CTAs; best of 2 / 3 / 4 in the forward
  const int C = Ca + Cb;
  const int Q8 = C / 8;
  int R = 256 / Q8;
  const long long total = (long long)N * H * W;
  if (R > total) R = (int)total;
  const int threads = Q8 * R;
  long long ctas = (long long)indm_num_sms() * per_sm;
  const long long max_ctas = (total + (long long)R * 2 - 1) / ((long long)R * 2);
  if (ctas > max_ctas) ctas = max_ctas;
  if (ctas < 1) ctas = 1;
  const bool drop = drop_ctl != nullptr && drop_p > 0.f;
  const dim3 grid((unsigned)ctas);
#define GN_F(D, P_, R_) return gn_apply_flat8_go<TIn, D, P_, R_>(grid, threads, stream, xa, Ca, xb, Cb, N, H * W, G, R, partial, gamma, beta, eps, act, out, raw, drop_p, drop_ctl, drop_stream, pp ? H : 0, pp ? W : 0)
  if (drop) {
    if (pp) { if (raw) GN_F(true, true, true); GN_F(true, true, false); }
    if (raw) GN_F(true, false, true);
    GN_F(true, false, false);
  }
  if (pp) { if (raw) GN_F(false, true, true); GN_F(false, true, false); }
  if (raw) GN_F(false, false, true);
  GN_F(false, false, false);
#undef GN_F
}

static int gn_stream_interleave() {
  static const int v = []() { const char* e = getenv("INDM_GN_INTERLEAVE"); return (e && e[0] == '1') ? 1 : 0; }();      // measured slower than contiguous ranges (26.9 vs 19.5 us on 32x32x128): off
  return v;
}

template <typename TIn, bool DROP, bool PP, bool RAW>
static int gn_apply_stream_go(dim3 grid, size_t smem, cudaStream_t stream, const void* xa, int C, int64_t N, int P, int G, int R, int chunk_px,
                              int stages, const float* partial, const float* gamma, const float* beta, float eps, int act, void* out, void* raw,
                              float drop_p, const unsigned long long* drop_ctl, unsigned drop_stream, int ppH, int ppW) {
  auto kern = gn_apply_stream_kernel<TIn, DROP, PP, RAW>;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e != cudaSuccess) {
      indm_set_error("gn_apply (stream): cudaFuncSetAttribute failed: %s", cudaGetErrorString(e));
      return INDM_ERR_CUDA;
    }
    configured = true;
  }
  indm_launch_pdl(kern, grid, dim3(288), smem, stream, (const TIn*)xa, C, (long long)N, P, G, R, chunk_px, stages, partial, gamma, beta, eps, act,
                  (__nv_bfloat16*)out, (__nv_bfloat16*)raw, drop_p, drop_ctl, drop_stream, ppH, ppW, gn_stream_interleave());
  INDM_CHECK_LAUNCH("gn_apply (stream)");
  return INDM_OK;
}

template <typename TIn>
static int gn_apply_stream_launch(const void* xa, int C, int64_t N, int H, int W, int G, const float* partial, const float* gamma,
                                  const float* beta, float eps, int act, void* out, void* raw, float drop_p,
                                  const unsigned long long* drop_ctl, unsigned drop_stream, cudaStream_t stream, bool pp) {
  static const int chunk_kb = []() { const char* e = getenv("INDM_GN_CHUNK_KB"); const int v = e ? atoi(e) : 0; return v > 0 ? v : 16; }();
  static const int stages = []() { const char* e = getenv("INDM_GN_STAGES"); const int v = e ? atoi(e) : 0; return v > 0 && v <= kGnStreamMaxStages ? v : 4; }();
  static const int per_sm = []() { const char* e = getenv("INDM_GN_CTAS_PER_SM"); const int v = e ? atoi(e) : 0; return v > 0 ? v : 2; }();
  const int R = 256 / (C / 8);
  int chunk_px = chunk_kb * 1024 / (C * (int)sizeof(TIn));
  chunk_px = chunk_px / R * R;
  if (chunk_px < R) chunk_px = R;
  const long long total = (long long)N * H * W;
  const long long nchunks = (total + chunk_px - 1) / chunk_px;
  long long ctas = (long long)indm_num_sms() * per_sm;
  if (ctas > (nchunks + 1) / 2) ctas = (nchunks + 1) / 2;
  if (ctas < 1) ctas = 1;
  const size_t smem = 128 + (size_t)stages * chunk_px * C * sizeof(TIn);
  const bool drop = drop_ctl != nullptr && drop_p > 0.f;
  const dim3 grid((unsigned)ctas);
#define GN_S(D, P_, R_) return gn_apply_stream_go<TIn, D, P_, R_>(grid, smem, stream, xa, C, N, H * W, G, R, chunk_px, stages, partial, gamma, beta, eps, act, out, raw, drop_p, drop_ctl, drop_stream, pp ? H : 0, pp ? W : 0)
  if (drop) {
    if (pp) { if (raw) GN_S(true, true, true); GN_S(true, true, false); }
    if (raw) GN_S(true, false, true);
    GN_S(true, false, false);
  }
  if (pp) { if (raw) GN_S(false, true, true); GN_S(false, true, false); }
  if (raw) GN_S(false, false, true);
  GN_S(false, false, false);
#undef GN_S
}

template <typename TIn, typename TOut>
static int gn_apply_launch(const void* xa, int Ca, const void* xb, int Cb, int64_t N, int H, int W, int G, const float* partial,
                           const float* gamma, const float* beta, float eps, int act, int resample, void* out, void* raw,
                           float drop_p, const unsigned long long* drop_ctl, unsigned drop_stream, cudaStream_t stream, bool pp = false) {
  using TO = typename std::conditional<std::is_same<TOut, Tf32Out>::value, float, TOut>::type;
  const int C = Ca + Cb;
  const long long Piter = resample == 2 ? (long long)(H / 2) * (W / 2) : (long long)H * W;
  const GnGeom g = gn_geom(C, Piter, N);
  dim3 grid(g.splits, (unsigned)N);
  static const bool flat = []() { const char* e = getenv("INDM_GN_FLAT"); return !(e && e[0] == '0'); }();
  INDM_CHECK_ARG(!pp || (resample == 0 && g.threads <= 256), "gn_apply_pp: no resampling, C <= 1024");
  // single dense source, BF16 out: 8 channels per thread.  INDM_GN_KERNEL = flat8 (default) | stream | flat
  static const int which = []() { const char* e = getenv("INDM_GN_KERNEL"); return !e ? 2 : (!strcmp(e, "stream") ? 1 : (!strcmp(e, "flat") ? 0 : 2)); }();
  const bool stream_on = which == 1;
  if (which == 2 && std::is_same<TOut, __nv_bfloat16>::value && resample == 0 && Ca % 8 == 0 && Cb % 8 == 0 && C / 8 <= 256 &&
      (((uintptr_t)xa | (uintptr_t)xb | (uintptr_t)out | (uintptr_t)raw) & 15) == 0 && (long long)N * H * W < (1ll << 31)) {
    return gn_apply_flat8_launch<TIn>(xa, Ca, xb, Cb, N, H, W, G, partial, gamma, beta, eps, act, out, raw, drop_p, drop_ctl, drop_stream, stream, pp);
  }
  if (stream_on && std::is_same<TOut, __nv_bfloat16>::value && resample == 0 && Cb == 0 && C % 8 == 0 && C / 8 <= 256 &&
      (256 % (C / 8) == 0) && (((uintptr_t)xa | (uintptr_t)out | (uintptr_t)raw) & 15) == 0 && (long long)N * H * W < (1ll << 31)) {
    return gn_apply_stream_launch<TIn>(xa, C, N, H, W, G, partial, gamma, beta, eps, act, out, raw, drop_p, drop_ctl, drop_stream, stream, pp);
  }
  if ((flat || pp) && resample == 0 && g.threads <= 256) {
    // equal-work ranges over the (image, pixel) space: grid = resident slots of the chip, or fewer when there is less than one
    // UNR-trip of work per CTA
    const long long total = (long long)N * H * W;
    long long ctas = (long long)indm_num_sms() * (g.threads > 128 ? 4 : 8);
    const long long max_ctas = (total + (long long)g.R * 4 - 1) / ((long long)g.R * 4);
    if (ctas > max_ctas) ctas = max_ctas;
    if (ctas < 1) ctas = 1;
    const long long P = (long long)H * W;
    if (drop_ctl != nullptr && drop_p > 0.f)
      indm_launch_pdl(gn_apply_flat_kernel<TIn, TOut, true>, dim3((unsigned)ctas), dim3(g.threads), 0, stream, (const TIn*)xa, Ca, (const TIn*)xb, Cb,
                      (long long)N, P, G, g.R, partial, gamma, beta, eps, act, (TO*)out, (TO*)raw, drop_p, drop_ctl, drop_stream, pp ? H : 0, pp ? W : 0);
    else
      indm_launch_pdl(gn_apply_flat_kernel<TIn, TOut, false>, dim3((unsigned)ctas), dim3(g.threads), 0, stream, (const TIn*)xa, Ca, (const TIn*)xb, Cb,
                      (long long)N, P, G, g.R, partial, gamma, beta, eps, act, (TO*)out, (TO*)raw, drop_p, drop_ctl, drop_stream, pp ? H : 0, pp ? W : 0);
    INDM_CHECK_LAUNCH("gn_apply (flat)");
    return INDM_OK;
  }
#define GN_ARGS (const TIn*)xa, Ca, (const TIn*)xb, Cb, H, W, G, g.R, partial, gamma, beta, eps, act, (TO*)out, (TO*)raw, drop_p, drop_ctl, drop_stream
  if (resample == 0 && drop_ctl != nullptr && drop_p > 0.f)
    indm_launch_pdl(gn_apply_kernel<TIn, TOut, 0, true>, grid, dim3(g.threads), 0, stream, GN_ARGS);
  else if (resample == 0)
    indm_launch_pdl(gn_apply_kernel<TIn, TOut, 0, false>, grid, dim3(g.threads), 0, stream, GN_ARGS);
  else if (resample == 1)
    indm_launch_pdl(gn_apply_kernel<TIn, TOut, 1, false>, grid, dim3(g.threads), 0, stream, GN_ARGS);
  else
    indm_launch_pdl(gn_apply_kernel<TIn, TOut, 2, false>, grid, dim3(g.threads), 0, stream, GN_ARGS);
#undef GN_ARGS
  INDM_CHECK_LAUNCH("gn_apply");
  return INDM_OK;
}

static int gn_apply_impl(const void* xa, int Ca, const void* xb, int Cb, int in_dtype, int64_t N, int H, int W, int G,
                         const float* partial, const float* gamma, const float* beta, float eps, int act_silu, int resample,
                         void* out, void* raw, int out_dtype, float drop_p, const unsigned long long* drop_ctl, unsigned drop_stream,
                         void* stream_, bool pp = false);

extern "C" int indm_gn_apply(const void* xa, int Ca, const void* xb, int Cb, int in_dtype, int64_t N, int H, int W, int G,
                             const float* partial, const float* gamma, const float* beta, float eps, int act_silu, int resample,
                             void* out, void* raw, int out_dtype, void* stream_) {
  return gn_apply_impl(xa, Ca, xb, Cb, in_dtype, N, H, W, G, partial, gamma, beta, eps, act_silu, resample, out, raw, out_dtype, 0.f,
                       nullptr, 0u, stream_);
}

extern "C" int indm_gn_apply_dropout(const void* xa, int Ca, const void* xb, int Cb, int in_dtype, int64_t N, int H, int W, int G,
                                     const float* partial, const float* gamma, const float* beta, float eps, int act_silu,
                                     void* out, int out_dtype, float drop_p, const uint64_t* drop_ctl, uint32_t drop_stream,
                                     void* stream_) {
  INDM_CHECK_ARG(drop_p >= 0.f && drop_p < 1.f && drop_ctl != nullptr, "gn_apply_dropout: need 0 <= p < 1 and a control buffer");
  return gn_apply_impl(xa, Ca, xb, Cb, in_dtype, N, H, W, G, partial, gamma, beta, eps, act_silu, 0, out, nullptr, out_dtype, drop_p,
                       (const unsigned long long*)drop_ctl, drop_stream, stream_);
}

extern "C" int indm_gn_apply_pp(const void* xa, int Ca, const void* xb, int Cb, int in_dtype, int64_t N, int H, int W, int G,
                                const float* partial, const float* gamma, const float* beta, float eps, int act_silu, void* out,
                                void* raw, int out_dtype, float drop_p, const uint64_t* drop_ctl, uint32_t drop_stream, void* stream_) {
  INDM_CHECK_ARG(drop_p >= 0.f && drop_p < 1.f, "gn_apply_pp: need 0 <= p < 1");
  return gn_apply_impl(xa, Ca, xb, Cb, in_dtype, N, H, W, G, partial, gamma, beta, eps, act_silu, 0, out, raw, out_dtype, drop_p,
                       (const unsigned long long*)drop_ctl, drop_stream, stream_, true);
}

static int gn_apply_impl(const void* xa, int Ca, const void* xb, int Cb, int in_dtype, int64_t N, int H, int W, int G,
                         const float* partial, const float* gamma, const float* beta, float eps, int act_silu, int resample,
                         void* out, void* raw, int out_dtype, float drop_p, const unsigned long long* drop_ctl, unsigned drop_stream,
                         void* stream_, bool pp) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (!xb) Cb = 0;
  const int C = Ca + Cb;
  INDM_CHECK_ARG(xa && partial && gamma && beta && out && N > 0 && H > 0 && W > 0, "gn_apply: bad arguments");
  INDM_CHECK_ARG(G >= 1 && G <= 32 && C % G == 0 && (C / G) % 4 == 0 && Ca % 4 == 0 && C / 4 <= 1024,
                 "gn_apply: need G <= 32, (C/G) %% 4 == 0 (C=%d G=%d)", C, G);
  INDM_CHECK_ARG(resample >= 0 && resample <= 2, "gn_apply: resample must be 0, 1 or 2");
  INDM_CHECK_ARG(resample != 2 || (H % 2 == 0 && W % 2 == 0), "gn_apply: down x2 needs even H, W");
  INDM_CHECK_ARG(N <= 65535, "gn_apply: N too large for grid.y");
  const bool in_f32 = in_dtype == INDM_DTYPE_F32, in_bf = in_dtype == INDM_DTYPE_BF16;
  const bool out_bf = out_dtype == INDM_DTYPE_BF16, out_tf = out_dtype == INDM_DTYPE_TF32, out_f = out_dtype == INDM_DTYPE_F32;
  INDM_CHECK_ARG((in_f32 || in_bf) && (out_bf || out_tf || out_f), "gn_apply: unsupported dtypes %d -> %d", in_dtype, out_dtype);
#define GN_GO(TI, TO_) return gn_apply_launch<TI, TO_>(xa, Ca, xb, Cb, N, H, W, G, partial, gamma, beta, eps, act_silu, resample, out, raw, drop_p, drop_ctl, drop_stream, stream, pp)
  if (in_f32 && out_bf) GN_GO(float, __nv_bfloat16);
  if (in_f32 && out_tf) GN_GO(float, Tf32Out);
  if (in_f32 && out_f) GN_GO(float, float);
  if (in_bf && out_bf) GN_GO(__nv_bfloat16, __nv_bfloat16);
  if (in_bf && out_tf) GN_GO(__nv_bfloat16, Tf32Out);
  GN_GO(__nv_bfloat16, float);
#undef GN_GO
}

extern "C" int indm_softmax_rows(const float* s, void* out, int64_t rows, int cols, int out_dtype, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  INDM_CHECK_ARG(s && out && rows > 0 && cols > 0, "softmax_rows: bad arguments");
  const int wpb = 8;
  const long long blocks = (rows + wpb - 1) / wpb;
  INDM_CHECK_ARG(blocks < (1LL << 31), "softmax_rows: too many rows");
  if (out_dtype == INDM_DTYPE_BF16)
    indm_launch_pdl(softmax_rows_kernel<__nv_bfloat16>, dim3((unsigned)blocks), dim3(wpb * 32), 0, stream, s, (__nv_bfloat16*)out, rows, cols, 0);
  else if (out_dtype == INDM_DTYPE_TF32 || out_dtype == INDM_DTYPE_F32)
    indm_launch_pdl(softmax_rows_kernel<float>, dim3((unsigned)blocks), dim3(wpb * 32), 0, stream, s, (float*)out, rows, cols, out_dtype == INDM_DTYPE_TF32);
  else
    INDM_CHECK_ARG(false, "softmax_rows: bad out_dtype");
  INDM_CHECK_LAUNCH("softmax_rows");
  return INDM_OK;
}

extern "C" int indm_prep_input(const float* x, void* out, int64_t N, int C, int H, int W, int cpad, float mul, float add, int act,
                               int out_dtype, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  INDM_CHECK_ARG(x && out && N > 0 && C > 0 && cpad >= C, "prep_input: bad arguments");
  const long long total = (long long)N * H * W * cpad;
  const int grid = grid_for(total, 256);
  if (out_dtype == INDM_DTYPE_BF16)
    indm_launch_pdl(prep_input_kernel<__nv_bfloat16>, dim3(grid), dim3(256), 0, stream, x, (__nv_bfloat16*)out, N, C, H * W, cpad, mul, add, act, 0);
  else if (out_dtype == INDM_DTYPE_TF32 || out_dtype == INDM_DTYPE_F32)
    indm_launch_pdl(prep_input_kernel<float>, dim3(grid), dim3(256), 0, stream, x, (float*)out, N, C, H * W, cpad, mul, add, act, out_dtype == INDM_DTYPE_TF32);
  else
    INDM_CHECK_ARG(false, "prep_input: bad out_dtype");
  INDM_CHECK_LAUNCH("prep_input");
  return INDM_OK;
}

extern "C" int indm_time_embedding(const float* time_cond, const float* sched, const int32_t* step, int sched_ld, int sched_col,
                                   const float* freqs, int kind, int64_t N, int dim, float* out, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  INDM_CHECK_ARG((time_cond || sched) && out && N > 0 && dim >= 4 && dim % 2 == 0, "time_embedding: bad arguments");
  INDM_CHECK_ARG(kind == 0 || (kind == 1 && freqs), "time_embedding: kind 1 needs freqs");
  indm_launch_pdl(time_embedding_kernel, dim3(grid_for(N * (dim / 2), 128)), dim3(128), 0, stream, time_cond, sched, step, sched_ld, sched_col, freqs, kind,
                                                                         N, dim, out);
  INDM_CHECK_LAUNCH("time_embedding");
  return INDM_OK;
}

extern "C" int indm_linear_f32(const float* in, const float* w, const float* bias, void* out, int64_t N, int K, int O,
                               int act_in, int act_out, int out_dtype, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  INDM_CHECK_ARG(in && w && out && N > 0 && O > 0, "linear: bad arguments");
  INDM_CHECK_ARG(K > 0, "linear: K must be positive");
  const int wpb = 4;
  const int gx = (O + wpb - 1) / wpb;
  // split the batch over grid.y only when there are too few output columns to fill the chip
  int gy = 1;
  while ((long long)gx * gy < 2LL * indm_num_sms() && gy < N) gy *= 2;
  dim3 grid(gx, gy);
  if (K % 128 != 0 || K > 1024) {
    indm_launch_pdl(linear_generic_kernel, grid, dim3(wpb * 32), 0, stream, in, w, bias, out, N, K, O, act_in, act_out, out_dtype);
    INDM_CHECK_LAUNCH("linear");
    return INDM_OK;
  }
  switch (K / 128) {
#define LIN(KV) case KV: indm_launch_pdl(linear_kernel<KV>, grid, dim3(wpb * 32), 0, stream, in, w, bias, out, N, K, O, act_in, act_out, out_dtype); break;
    LIN(1) LIN(2) LIN(3) LIN(4) LIN(5) LIN(6) LIN(7) LIN(8)
#undef LIN
  }
  INDM_CHECK_LAUNCH("linear");
  return INDM_OK;
}

template <typename TIn, typename TOut>
static int fir_launch(const void* x, void* y, int64_t N, int H, int W, int C, const float* k, int mode, cudaStream_t stream) {
  using TO = typename std::conditional<std::is_same<TOut, Tf32Out>::value, float, TOut>::type;
  const int Ho = mode == 1 ? 2 * H : (mode == 2 ? H / 2 : (mode == 4 ? H - 1 : H + 1));
  const int Wo = mode == 1 ? 2 * W : (mode == 2 ? W / 2 : (mode == 4 ? W - 1 : W + 1));
  // CTA = 256 threads = (qb channel groups) x (256 / qb output columns); a thread owns 8 channels (4 in the down x2 mode)
  const int Q = mode == 2 ? C / 4 : C / 8;
  int qb_log2 = 0;
  while ((2 << qb_log2) <= Q && qb_log2 < 6) ++qb_log2;
  const int qb = 1 << qb_log2, cols = 256 >> qb_log2;
  const int ctiles = (Q + qb - 1) / qb, xtiles = (Wo + cols - 1) / cols;
  // strip length: long strips amortise the 3 warm-up rows; keep enough CTAs in flight
  int RS = Ho;
  static const int fir_waves = []() { const char* e = getenv("INDM_FIR_WAVES"); const int v = e ? atoi(e) : 0; return v > 0 ? v : 8; }();
  const int waves = mode == 1 ? (fir_waves < 4 ? fir_waves : 4) : fir_waves;
  while (RS > 4 && (long long)N * ((Ho + RS - 1) / RS) * ctiles * xtiles < (long long)waves * indm_num_sms()) RS = (RS + 1) / 2;
  const dim3 grid((unsigned)(ctiles * xtiles), (unsigned)((Ho + RS - 1) / RS), (unsigned)N);
  INDM_CHECK_ARG(N <= 65535, "fir_nhwc: N > 65535");
  if (mode == 1)
    indm_launch_pdl(fir_nhwc_kernel<TIn, TOut, 1>, grid, dim3(256), 0, stream, (const TIn*)x, (TO*)y, N, H, W, C, k[0], k[1], k[2], k[3], qb_log2, RS, ctiles);
  else if (mode == 2)
    indm_launch_pdl(fir_nhwc4_kernel<TIn, TOut, 2>, grid, dim3(256), 0, stream, (const TIn*)x, (TO*)y, N, H, W, C, k[0], k[1], k[2], k[3], qb_log2, RS, ctiles);
  else if (mode == 3)
    indm_launch_pdl(fir_nhwc_kernel<TIn, TOut, 3>, grid, dim3(256), 0, stream, (const TIn*)x, (TO*)y, N, H, W, C, k[0], k[1], k[2], k[3], qb_log2, RS, ctiles);
  else
    indm_launch_pdl(fir_nhwc_kernel<TIn, TOut, 4>, grid, dim3(256), 0, stream, (const TIn*)x, (TO*)y, N, H, W, C, k[0], k[1], k[2], k[3], qb_log2, RS, ctiles);
  INDM_CHECK_LAUNCH("fir_nhwc");
  return INDM_OK;
}

extern "C" int indm_fir_nhwc(const void* x, void* y, int dtype_in, int dtype_out, int64_t N, int H, int W, int C, const float* k1,
                             int mode, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  INDM_CHECK_ARG(x && y && k1 && N > 0 && H > 0 && W > 0 && C > 0 && C % 8 == 0, "fir_nhwc: bad arguments (C must be a multiple of 8)");
  INDM_CHECK_ARG(mode >= 1 && mode <= 4, "fir_nhwc: mode must be 1 (up), 2 (down), 3 (pad (2,2)) or 4 (pad (1,1))");
  INDM_CHECK_ARG(mode != 4 || (H > 1 && W > 1), "fir_nhwc: mode 4 needs H, W > 1");
  INDM_CHECK_ARG(mode != 2 || (H % 2 == 0 && W % 2 == 0), "fir_nhwc: down needs even H, W");
  // k1 is a HOST pointer to the 4 separable taps, already normalised (and gain-scaled per axis)
  const bool in_f = dtype_in == INDM_DTYPE_F32 || dtype_in == INDM_DTYPE_TF32, in_b = dtype_in == INDM_DTYPE_BF16;
  INDM_CHECK_ARG(in_f || in_b, "fir_nhwc: bad dtype_in");
  if (dtype_out == INDM_DTYPE_BF16) return in_f ? fir_launch<float, __nv_bfloat16>(x, y, N, H, W, C, k1, mode, stream)
                                                : fir_launch<__nv_bfloat16, __nv_bfloat16>(x, y, N, H, W, C, k1, mode, stream);
  if (dtype_out == INDM_DTYPE_TF32) return in_f ? fir_launch<float, Tf32Out>(x, y, N, H, W, C, k1, mode, stream)
                                                : fir_launch<__nv_bfloat16, Tf32Out>(x, y, N, H, W, C, k1, mode, stream);
  if (dtype_out == INDM_DTYPE_F32) return in_f ? fir_launch<float, float>(x, y, N, H, W, C, k1, mode, stream)
                                               : fir_launch<__nv_bfloat16, float>(x, y, N, H, W, C, k1, mode, stream);
  INDM_CHECK_ARG(false, "fir_nhwc: bad dtype_out");
}
