// Predictor-corrector state update of the INDM sampler (sampling.py:205-210, :272-292; sde_lib.py:105-118,171-184,
// 310-323), one fused HBM pass per update instead of the reference's ~12 elementwise kernels + 2 norm reductions.
// The Gaussian noise can be passed in (parity with a recorded reference trajectory) or generated in registers from a
// counter-based Philox4x32-10 stream, in which case no noise tensor ever touches HBM.  Per-step scalars come from a
// device-resident schedule table indexed by a device step counter, so the whole sampling loop replays one CUDA
// graph with no host involvement.
#include "../../include/indm_b200.h"
#include "common.cuh"
#include "philox.cuh"

namespace {

// four standard normals for element quad `q` of stream (seed, a, b)
__device__ __forceinline__ float4 normal4(uint64_t seed, uint32_t a, uint32_t b, uint64_t q) {
  const uint4 r = Philox::round10(make_uint4((uint32_t)q, (uint32_t)(q >> 32), a, b), make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
  const float s = 2.3283064365386963e-10f;  // 2^-32
  const float u0 = ((float)r.x + 0.5f) * s, u1 = ((float)r.y + 0.5f) * s;
  const float u2 = ((float)r.z + 0.5f) * s, u3 = ((float)r.w + 0.5f) * s;
  const float r0 = sqrtf(-2.0f * logf(fmaxf(u0, 1e-12f))), r1 = sqrtf(-2.0f * logf(fmaxf(u2, 1e-12f)));
  float s0, c0, s1, c1;
  sincospif(2.0f * u1, &s0, &c0);
  sincospif(2.0f * u3, &s1, &c1);
  return make_float4(r0 * c0, r0 * s0, r1 * c1, r1 * s1);
}

__device__ __forceinline__ float4 load_or_draw(const float* z, long long i4, uint64_t seed, uint32_t a, uint32_t b) {
  return z ? reinterpret_cast<const float4*>(z)[i4] : normal4(seed, a, b, (uint64_t)i4);
}

__global__ void predictor_update_kernel(float* __restrict__ x, const float* __restrict__ s, const float* __restrict__ z,
                                        float* __restrict__ x_mean, const float* __restrict__ coef, int coef_ld,
                                        const int32_t* __restrict__ step, long long total4, uint64_t seed, const uint64_t* __restrict__ seed_dev,
                                        uint32_t off) {
  pdl_trigger();
  pdl_wait();
  const int st = step ? *step : 0;
  if (seed_dev) seed = *seed_dev;
  const float a = coef[(long long)st * coef_ld + 0], c = coef[(long long)st * coef_ld + 1], d = coef[(long long)st * coef_ld + 2];
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total4; i += (long long)gridDim.x * blockDim.x) {
    const float4 xv = reinterpret_cast<const float4*>(x)[i];
    const float4 sv = reinterpret_cast<const float4*>(s)[i];
    const float4 zv = load_or_draw(z, i, seed, (uint32_t)st, off);
    const float4 m = make_float4(a * xv.x + c * sv.x, a * xv.y + c * sv.y, a * xv.z + c * sv.z, a * xv.w + c * sv.w);
    if (x_mean) reinterpret_cast<float4*>(x_mean)[i] = m;
    reinterpret_cast<float4*>(x)[i] = make_float4(m.x + d * zv.x, m.y + d * zv.y, m.z + d * zv.z, m.w + d * zv.w);
  }
}

// grid (chunks, N): per-sample partial sums of squares, warp-shuffle + one atomic per warp
__global__ void langevin_norms_kernel(const float* __restrict__ s, const float* __restrict__ z, float* __restrict__ out,
                                      const int32_t* __restrict__ step, long long D4, uint64_t seed,
                                      const uint64_t* __restrict__ seed_dev, uint32_t off) {
  const int st = step ? *step : 0;
  if (seed_dev) seed = *seed_dev;
  const long long n = blockIdx.y;
  float ss = 0.f, zz = 0.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < D4; i += (long long)gridDim.x * blockDim.x) {
    const long long g = n * D4 + i;
    const float4 sv = reinterpret_cast<const float4*>(s)[g];
    const float4 zv = load_or_draw(z, g, seed, (uint32_t)st, off);
    ss += (sv.x * sv.x + sv.y * sv.y) + (sv.z * sv.z + sv.w * sv.w);
    zz += (zv.x * zv.x + zv.y * zv.y) + (zv.z * zv.z + zv.w * zv.w);
  }
  ss = warp_sum(ss);
  zz = warp_sum(zz);
  if ((threadIdx.x & 31) == 0) {
    atomicAdd(&out[n * 2 + 0], ss);
    atomicAdd(&out[n * 2 + 1], zz);
  }
}

// local sums of the per-sample L2 norms and the local batch size: sums[0] = sum_n |s_n|, sums[1] = sum_n |z_n|, sums[2] = N.
// Summed over ranks (one 12-byte all-reduce) they give the GLOBAL batch means of sampling.py:286-287 for a batch-sharded sampler.
__global__ void langevin_norm_sums_kernel(const float* __restrict__ norms, float* __restrict__ sums, int N) {
  float gs = 0.f, gz = 0.f;
  for (int n = threadIdx.x; n < N; n += 32) {
    gs += sqrtf(norms[n * 2 + 0]);
    gz += sqrtf(norms[n * 2 + 1]);
  }
  gs = warp_sum(gs);
  gz = warp_sum(gz);
  if (threadIdx.x == 0) {
    sums[0] = gs;
    sums[1] = gz;
    sums[2] = (float)N;
  }
}

__global__ void langevin_update_kernel(float* __restrict__ x, const float* __restrict__ s, const float* __restrict__ z,
                                       float* __restrict__ x_mean, const float* __restrict__ norms, const float* __restrict__ gsums,
                                       const float* __restrict__ coef,
                                       int coef_ld, const int32_t* __restrict__ step, int N, long long total4, uint64_t seed,
                                       const uint64_t* __restrict__ seed_dev, uint32_t off) {
  __shared__ float sh[2];
  const int st = step ? *step : 0;
  if (seed_dev) seed = *seed_dev;
  if (gsums) {
    // global statistics: sums over all ranks' samples (already all-reduced), divided by the global batch size
    if (threadIdx.x == 0) {
      sh[0] = gsums[0] / gsums[2];
      sh[1] = gsums[1] / gsums[2];
    }
  } else if (threadIdx.x < 32) {
    // batch means of the per-sample L2 norms (sampling.py:286-287): N is small, every CTA recomputes them
    float gs = 0.f, gz = 0.f;
    for (int n = threadIdx.x; n < N; n += 32) {
      gs += sqrtf(norms[n * 2 + 0]);
      gz += sqrtf(norms[n * 2 + 1]);
    }
    gs = warp_sum(gs);
    gz = warp_sum(gz);
    if (threadIdx.x == 0) {
      sh[0] = gs / (float)N;
      sh[1] = gz / (float)N;
    }
  }
  __syncthreads();
  const float alpha = coef[(long long)st * coef_ld + 0], snr = coef[(long long)st * coef_ld + 1];
  const float r = snr * sh[1] / sh[0];
  const float eps = r * r * 2.0f * alpha;         // sampling.py:288
  const float nz = sqrtf(eps * 2.0f);             // :290
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total4; i += (long long)gridDim.x * blockDim.x) {
    const float4 xv = reinterpret_cast<const float4*>(x)[i];
    const float4 sv = reinterpret_cast<const float4*>(s)[i];
    const float4 zv = load_or_draw(z, i, seed, (uint32_t)st, off);
    const float4 m = make_float4(xv.x + eps * sv.x, xv.y + eps * sv.y, xv.z + eps * sv.z, xv.w + eps * sv.w);
    if (x_mean) reinterpret_cast<float4*>(x_mean)[i] = m;
    reinterpret_cast<float4*>(x)[i] = make_float4(m.x + nz * zv.x, m.y + nz * zv.y, m.z + nz * zv.z, m.w + nz * zv.w);
  }
}

__global__ void advance_step_kernel(int32_t* step) {
  pdl_trigger();
  pdl_wait();
  *step += 1;
}

__global__ void randn_kernel(float* __restrict__ out, long long n4, long long n, uint64_t seed, uint32_t a, uint32_t b) {
  pdl_trigger();
  pdl_wait();
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const float4 v = normal4(seed, a, b, (uint64_t)i);
    if (i * 4 + 3 < n) reinterpret_cast<float4*>(out)[i] = v;
    else {
      const float t[4] = {v.x, v.y, v.z, v.w};
      for (int e = 0; e < 4 && i * 4 + e < n; ++e) out[i * 4 + e] = t[e];
    }
  }
}

inline int ew_grid(long long work) {
  long long b = (work + 255) / 256;
  const long long cap = (long long)indm_num_sms() * 8;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (int)b;
}

}  // namespace

extern "C" int indm_pc_predictor_update(float* x, const float* s, const float* z, float* x_mean, const float* coef, int coef_ld,
                                        const int32_t* step, int64_t N, int64_t D, uint64_t seed, const uint64_t* seed_dev,
                                        uint64_t rng_offset, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  INDM_CHECK_ARG(x && s && coef && N > 0 && D > 0 && coef_ld >= 3, "pc_predictor_update: bad arguments");
  INDM_CHECK_ARG((N * D) % 4 == 0, "pc_predictor_update: N*D must be a multiple of 4");
  const long long total4 = N * D / 4;
  indm_launch_pdl(predictor_update_kernel, dim3(ew_grid(total4)), dim3(256), 0, stream, x, s, z, x_mean, coef, coef_ld, step, total4, seed, seed_dev,
                                                              (uint32_t)rng_offset);
  INDM_CHECK_LAUNCH("pc_predictor_update");
  return INDM_OK;
}

extern "C" int indm_langevin_norms(const float* s, const float* z, float* out, const int32_t* step, int64_t N, int64_t D,
                                   uint64_t seed, const uint64_t* seed_dev, uint64_t rng_offset, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  INDM_CHECK_ARG(s && out && N > 0 && N <= 65535 && D > 0 && D % 4 == 0, "langevin_norms: bad arguments");
  cudaError_t e = cudaMemsetAsync(out, 0, sizeof(float) * 2 * N, stream);
  if (e != cudaSuccess) {
    indm_set_error("langevin_norms: memset: %s", cudaGetErrorString(e));
    return INDM_ERR_CUDA;
  }
  const long long D4 = D / 4;
  int chunks = (int)((D4 + 1023) / 1024);
  if (chunks < 1) chunks = 1;
  langevin_norms_kernel<<<dim3(chunks, (unsigned)N), 256, 0, stream>>>(s, z, out, step, D4, seed, seed_dev, (uint32_t)rng_offset);
  INDM_CHECK_LAUNCH("langevin_norms");
  return INDM_OK;
}

extern "C" int indm_langevin_update(float* x, const float* s, const float* z, float* x_mean, const float* norms, const float* coef,
                                    int coef_ld, const int32_t* step, int64_t N, int64_t D, uint64_t seed, const uint64_t* seed_dev,
                                    uint64_t rng_offset, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  INDM_CHECK_ARG(x && s && norms && coef && N > 0 && D > 0 && coef_ld >= 2, "langevin_update: bad arguments");
  INDM_CHECK_ARG((N * D) % 4 == 0, "langevin_update: N*D must be a multiple of 4");
  const long long total4 = N * D / 4;
  langevin_update_kernel<<<ew_grid(total4), 256, 0, stream>>>(x, s, z, x_mean, norms, nullptr, coef, coef_ld, step, (int)N, total4, seed,
                                                             seed_dev, (uint32_t)rng_offset);
  INDM_CHECK_LAUNCH("langevin_update");
  return INDM_OK;
}

extern "C" int indm_langevin_norm_sums(const float* norms, float* sums, int64_t N, void* stream_) {
  INDM_CHECK_ARG(norms && sums && N > 0, "langevin_norm_sums: bad arguments");
  langevin_norm_sums_kernel<<<1, 32, 0, (cudaStream_t)stream_>>>(norms, sums, (int)N);
  INDM_CHECK_LAUNCH("langevin_norm_sums");
  return INDM_OK;
}

extern "C" int indm_langevin_update_global(float* x, const float* s, const float* z, float* x_mean, const float* gsums, const float* coef,
                                           int coef_ld, const int32_t* step, int64_t N, int64_t D, uint64_t seed, const uint64_t* seed_dev,
                                           uint64_t rng_offset, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  INDM_CHECK_ARG(x && s && gsums && coef && N > 0 && D > 0 && coef_ld >= 2, "langevin_update_global: bad arguments");
  INDM_CHECK_ARG((N * D) % 4 == 0, "langevin_update_global: N*D must be a multiple of 4");
  const long long total4 = N * D / 4;
  langevin_update_kernel<<<ew_grid(total4), 256, 0, stream>>>(x, s, z, x_mean, nullptr, gsums, coef, coef_ld, step, (int)N, total4, seed,
                                                             seed_dev, (uint32_t)rng_offset);
  INDM_CHECK_LAUNCH("langevin_update_global");
  return INDM_OK;
}

extern "C" int indm_advance_step(int32_t* step, void* stream_) {
  INDM_CHECK_ARG(step, "advance_step: null");
  indm_launch_pdl(advance_step_kernel, dim3(1), dim3(1), 0, (cudaStream_t)stream_, step);
  INDM_CHECK_LAUNCH("advance_step");
  return INDM_OK;
}

extern "C" int indm_randn_f32(float* out, int64_t n, uint64_t seed, uint64_t rng_offset, void* stream_) {
  INDM_CHECK_ARG(out && n > 0, "randn: bad arguments");
  const long long n4 = (n + 3) / 4;
  randn_kernel<<<ew_grid(n4), 256, 0, (cudaStream_t)stream_>>>(out, n4, n, seed, (uint32_t)rng_offset, (uint32_t)(rng_offset >> 32));
  INDM_CHECK_LAUNCH("randn");
  return INDM_OK;
}

namespace {
__global__ void sched_broadcast_kernel(float* __restrict__ out, long long n, const float* __restrict__ sched, int ld, int col_num,
                                       int col_den, const int32_t* __restrict__ step) {
  pdl_trigger();
  pdl_wait();
  const int st = step ? *step : 0;
  float v = sched[(long long)st * ld + col_num];
  if (col_den >= 0) v /= sched[(long long)st * ld + col_den];
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) out[i] = v;
}
}  // namespace

extern "C" int indm_sched_broadcast(float* out, int64_t n, const float* sched, int ld, int col_num, int col_den, const int32_t* step,
                                    void* stream_) {
  INDM_CHECK_ARG(out && sched && n > 0 && ld > 0 && col_num >= 0, "sched_broadcast: bad arguments");
  indm_launch_pdl(sched_broadcast_kernel, dim3(ew_grid(n)), dim3(256), 0, (cudaStream_t)stream_, out, n, sched, ld, col_num, col_den, step);
  INDM_CHECK_LAUNCH("sched_broadcast");
  return INDM_OK;
}
