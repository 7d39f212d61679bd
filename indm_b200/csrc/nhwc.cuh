#include <cstdlib>
// Shared pieces of the NHWC streaming kernels (score_kernels.cu, backward_kernels.cu): typed 4-channel vector access and
// launch geometry.  Everything here has internal linkage.
#pragma once
#include <type_traits>

#include "common.cuh"

namespace {

// ---------------------------------------------------------------- typed 4-channel vector access
template <typename T>
struct Vec4;
template <>
struct Vec4<float> {
  static __device__ __forceinline__ float4 load(const float* p) { return *reinterpret_cast<const float4*>(p); }
  static __device__ __forceinline__ void store(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
};
template <>
struct Vec4<__nv_bfloat16> {
  static __device__ __forceinline__ float4 load(const __nv_bfloat16* p) {
    const uint2 r = *reinterpret_cast<const uint2*>(p);
    const __nv_bfloat162 a = *reinterpret_cast<const __nv_bfloat162*>(&r.x);
    const __nv_bfloat162 b = *reinterpret_cast<const __nv_bfloat162*>(&r.y);
    const float2 fa = __bfloat1622float2(a), fb = __bfloat1622float2(b);
    return make_float4(fa.x, fa.y, fb.x, fb.y);
  }
  static __device__ __forceinline__ void store(__nv_bfloat16* p, float4 v) {
    uint2 r;
    r.x = pack_bf16x2(v.x, v.y);
    r.y = pack_bf16x2(v.z, v.w);
    *reinterpret_cast<uint2*>(p) = r;
  }
};
// TF32-mode operands are stored as plain fp32: the GEMM kernel performs the tf32 hi/lo split itself (3xTF32, igemm.cu),
// so producers must NOT pre-round.
__device__ __forceinline__ float tf32_operand(float v) { return v; }
struct Tf32Out {};
template <>
struct Vec4<Tf32Out> {
  static __device__ __forceinline__ void store(float* p, float4 v) {
    *reinterpret_cast<float4*>(p) = make_float4(tf32_operand(v.x), tf32_operand(v.y), tf32_operand(v.z), tf32_operand(v.w));
  }
};

inline int grid_for(long long work_items, int threads) {
  long long b = (work_items + threads - 1) / threads;
  const long long cap = (long long)indm_num_sms() * 16;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (int)b;
}

struct GnGeom {
  int Q, R, threads, splits;
};
inline GnGeom gn_geom(int C, long long P_iter, long long N) {
  GnGeom g;
  g.Q = C / 4;
  g.R = 256 / g.Q;
  if (g.R < 1) g.R = 1;
  if (g.R > P_iter) g.R = (int)P_iter;
  g.threads = g.Q * g.R;
  // ~8 CTAs per SM over the chip (measured best of 4 / 8 / 12 / 16 / 24), but at least ~8 pixel-rows of work per thread row
  static const int waves = []() { const char* e = getenv("INDM_GN_WAVES"); const int v = e ? atoi(e) : 0; return v > 0 ? v : 8; }();
  long long want = ((long long)waves * indm_num_sms() + N - 1) / N;
  long long maxs = P_iter / (g.R * 4LL);
  if (maxs < 1) maxs = 1;
  if (want > maxs) want = maxs;
  if (want < 1) want = 1;
  g.splits = (int)want;
  return g;
}


}  // namespace
