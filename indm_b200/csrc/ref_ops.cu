// The reference's two native operators, rewritten for sm_100a (no code shared with op/*.cu):
//   upfirdn2d      op/upfirdn2d.cpp:12-19 -> op/upfirdn2d_kernel.cu:209-369
//   fused_bias_act op/fused_bias_act.cpp:11-17 -> op/fused_bias_act_kernel.cu:19-99
// Both are pure HBM-streaming kernels.  upfirdn2d stages the input footprint of an output tile in shared memory
// (coalesced 128-byte rows, zero-filled halo) and each thread produces 4 horizontally adjacent outputs so that the
// store is a 16-byte vector; only taps that land on real (non zero-inserted) samples are visited, which for the
// up=2 / 4-tap call (models/up_or_down_sampling.py:223) is 2x2 of the 4x4 taps.
#include "../../include/indm_b200.h"
#include "common.cuh"

namespace {

constexpr int TILE_H = 16;
constexpr int TILE_W = 64;   // 256 threads x 4 outputs along x
constexpr int MAX_K = 8;

__host__ __device__ inline int floor_div(int a, int b) { return (a >= 0) ? a / b : -((-a + b - 1) / b); }

struct UpfirdnParams {
  int in_h, in_w, out_h, out_w, kh, kw, up_x, up_y, down_x, down_y, pad_x0, pad_y0;
  int sm_h, sm_w;  // staged input footprint (rows, cols)
  int tiles_x, tiles_y;
};

// grid: (tiles_x * tiles_y, planes); dynamic smem: sm_h*sm_w floats + kh*kw floats
__global__ void __launch_bounds__(256) upfirdn2d_kernel(const float* __restrict__ x, const float* __restrict__ k,
                                                        float* __restrict__ y, long long planes, UpfirdnParams p) {
  extern __shared__ float smem[];
  float* sx = smem;
  float* sk = smem + p.sm_h * p.sm_w;
  const int tile = blockIdx.x;
  const int tx = tile % p.tiles_x, ty = tile / p.tiles_x;
  const int oy0 = ty * TILE_H, ox0 = tx * TILE_W;
  // first / last upsampled-grid coordinate touched by this tile, and the input rows/cols that cover them
  const int uy_min = oy0 * p.down_y - p.pad_y0;
  const int ux_min = ox0 * p.down_x - p.pad_x0;
  const int iy0 = floor_div(uy_min + p.up_y - 1, p.up_y);  // ceil(uy_min / up)
  const int ix0 = floor_div(ux_min + p.up_x - 1, p.up_x);
  // flipped kernel in smem: kf[i][j] = k[kh-1-i][kw-1-j]
  for (int i = threadIdx.x; i < p.kh * p.kw; i += blockDim.x) {
    const int a = i / p.kw, b = i % p.kw;
    sk[i] = k[(p.kh - 1 - a) * p.kw + (p.kw - 1 - b)];
  }
  for (long long plane = blockIdx.y; plane < planes; plane += gridDim.y) {
    const float* src = x + plane * (long long)p.in_h * p.in_w;
    __syncthreads();
    for (int i = threadIdx.x; i < p.sm_h * p.sm_w; i += blockDim.x) {
      const int r = i / p.sm_w, c = i % p.sm_w;
      const int iy = iy0 + r, ix = ix0 + c;
      float v = 0.f;
      if (iy >= 0 && iy < p.in_h && ix >= 0 && ix < p.in_w) v = src[(long long)iy * p.in_w + ix];
      sx[i] = v;
    }
    __syncthreads();
    const int lx = (threadIdx.x % (TILE_W / 4)) * 4;
    const int ly = threadIdx.x / (TILE_W / 4);
    const int oy = oy0 + ly;
    if (oy < p.out_h) {
      float acc[4] = {0.f, 0.f, 0.f, 0.f};
      for (int i = 0; i < p.kh; ++i) {
        const int uy = oy * p.down_y + i - p.pad_y0;
        if (uy < 0) continue;
        if (uy % p.up_y) continue;
        const int iy = uy / p.up_y;
        if (iy >= p.in_h) continue;
        const float* row = sx + (iy - iy0) * p.sm_w;
#pragma unroll
        for (int o = 0; o < 4; ++o) {
          const int ox = ox0 + lx + o;
          for (int j = 0; j < p.kw; ++j) {
            const int ux = ox * p.down_x + j - p.pad_x0;
            if (ux < 0 || (ux % p.up_x)) continue;
            const int ix = ux / p.up_x;
            if (ix >= p.in_w) continue;
            acc[o] += row[ix - ix0] * sk[i * p.kw + j];
          }
        }
      }
      float* dst = y + plane * (long long)p.out_h * p.out_w + (long long)oy * p.out_w + ox0 + lx;
      if (ox0 + lx + 3 < p.out_w && (p.out_w & 3) == 0) {
        *reinterpret_cast<float4*>(dst) = make_float4(acc[0], acc[1], acc[2], acc[3]);
      } else {
#pragma unroll
        for (int o = 0; o < 4; ++o)
          if (ox0 + lx + o < p.out_w) dst[o] = acc[o];
      }
    }
  }
}


// Compile-time specialisation for the call patterns the network makes (models/up_or_down_sampling.py:195-257: the 4-tap FIR with
// up x2 / down x2 / plain filtering): UP, DOWN and the tap count are template constants, so the tap loops unroll into predicated
// FMAs with shift / mask index arithmetic (the generic kernel below spends ~50 instructions of runtime % and / per tap and two
// block-wide barriers per plane: 0.02-0.03 of the HBM roofline on 16 x 256 x 32 x 32).  One thread = 4 horizontally adjacent
// outputs; the input span its taps touch is read straight from global memory through L1 (a 32 x 32 plane is 4 KB: every reuse is
// an L1 hit), so there is no staging, no barrier and every thread is independent; the store is one 16-byte vector.  Same tap
// order (i outer, j inner) and the same fp32 FMAs as the generic kernel: bit-identical results.
template <int UP, int DOWN, int K>
__global__ void __launch_bounds__(256) upfirdn2d_spec_kernel(const float* __restrict__ x, const float* __restrict__ k,
                                                             float* __restrict__ y, long long planes, UpfirdnParams p) {
  constexpr int UPM = UP - 1;                      // UP is 1 or 2
  constexpr int UPS = UP == 2 ? 1 : 0;
  constexpr int SPAN = 3 * DOWN + K;               // upsampled-grid columns the 4 outputs of a thread touch
  float kf[K * K];                                 // flipped kernel: kf[i][j] = k[K-1-i][K-1-j]
#pragma unroll
  for (int i = 0; i < K * K; ++i) kf[i] = __ldg(k + (K * K - 1 - i));
  constexpr int RG = 4;                            // output rows per thread: 4 independent row computations in flight
  const int qw = (p.out_w + 3) >> 2;               // output quads per row
  const int rgs = (p.out_h + RG - 1) / RG;
  const long long total = planes * (long long)rgs * qw;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int qx = (int)(idx % qw);
    const long long t = idx / qw;
    const int rg = (int)(t % rgs);
    const long long plane = t / rgs;
    const int ox0 = qx * 4;
    const float* src = x + plane * (long long)p.in_h * p.in_w;
    const int ux0 = ox0 * DOWN - p.pad_x0;         // first upsampled-grid column of the span
    float acc[RG][4];
#pragma unroll
    for (int r = 0; r < RG; ++r) {
      acc[r][0] = acc[r][1] = acc[r][2] = acc[r][3] = 0.f;
      const int oy = rg * RG + r;
#pragma unroll
      for (int i = 0; i < K; ++i) {
        const int uy = oy * DOWN + i - p.pad_y0;
        const int iy = uy >> UPS;
        if (oy >= p.out_h || uy < 0 || (uy & UPM) || iy >= p.in_h) continue;
        const float* row = src + (long long)iy * p.in_w;
        float v[SPAN];                             // the span of this input row on the upsampled grid (0 where no sample sits)
#pragma unroll
        for (int s_ = 0; s_ < SPAN; ++s_) {
          const int ux = ux0 + s_;
          const int ix = ux >> UPS;
          v[s_] = (ux >= 0 && !(ux & UPM) && ix < p.in_w) ? __ldg(row + ix) : 0.f;
        }
#pragma unroll
        for (int o = 0; o < 4; ++o) {
#pragma unroll
          for (int j = 0; j < K; ++j) acc[r][o] += v[o * DOWN + j] * kf[i * K + j];
        }
      }
    }
#pragma unroll
    for (int r = 0; r < RG; ++r) {
      const int oy = rg * RG + r;
      if (oy >= p.out_h) break;
      float* dst = y + plane * (long long)p.out_h * p.out_w + (long long)oy * p.out_w + ox0;
      if (ox0 + 3 < p.out_w && (p.out_w & 3) == 0) {
        *reinterpret_cast<float4*>(dst) = make_float4(acc[r][0], acc[r][1], acc[r][2], acc[r][3]);
      } else {
#pragma unroll
        for (int o = 0; o < 4; ++o)
          if (ox0 + o < p.out_w) dst[o] = acc[r][o];
      }
    }
  }
}

template <int UP, int DOWN, int K>
int launch_upfirdn_spec(const float* x, const float* k, float* y, long long major, const UpfirdnParams& p, cudaStream_t stream) {
  const long long total = major * (long long)((p.out_h + 3) / 4) * ((p.out_w + 3) / 4);
  long long blocks = (total + 255) / 256;
  const long long cap = (long long)indm_num_sms() * 32;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  upfirdn2d_spec_kernel<UP, DOWN, K><<<(unsigned)blocks, 256, 0, stream>>>(x, k, y, major, p);
  INDM_CHECK_LAUNCH("upfirdn2d (specialised)");
  return INDM_OK;
}

// y = act(x + b) * scale, 4 elements per thread when the bias index is constant over the 4 (step_b % 4 == 0)
template <bool VEC>
__global__ void bias_act_kernel(const float* __restrict__ x, const float* __restrict__ bias, const float* __restrict__ ref,
                                float* __restrict__ y, long long n, int size_b, long long step_b, int act, int grad, float alpha,
                                float scale) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  if (VEC) {
    const long long n4 = n >> 2;
    // four independent 16-byte loads in flight per thread; the bias index is a shift / mask when step_b and size_b are powers of two
    // (every call the network makes: step_b = H W, size_b = C), else the 64-bit division / modulo
    const bool pow2 = bias && (step_b & (step_b - 1)) == 0 && (size_b & (size_b - 1)) == 0;
    const int sh = pow2 ? (63 - __clzll(step_b)) : 0;
    auto bidx = [&](long long i) -> int { return pow2 ? (int)(((i * 4) >> sh) & (size_b - 1)) : (int)(((i * 4) / step_b) % size_b); };
    for (long long i0 = (long long)blockIdx.x * blockDim.x + threadIdx.x; i0 < n4; i0 += 4 * stride) {
      float4 v[4], r[4];
      float b[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const long long i = i0 + u * stride;
        if (i < n4) {
          v[u] = reinterpret_cast<const float4*>(x)[i];
          r[u] = ref ? reinterpret_cast<const float4*>(ref)[i] : make_float4(0.f, 0.f, 0.f, 0.f);
          b[u] = bias ? __ldg(bias + bidx(i)) : 0.f;
        }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const long long i = i0 + u * stride;
        if (i >= n4) break;
        float in[4] = {v[u].x + b[u], v[u].y + b[u], v[u].z + b[u], v[u].w + b[u]};
        const float rf[4] = {r[u].x, r[u].y, r[u].z, r[u].w};
        float o[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float xx = in[e];
          float yy;
          if (act == 1) yy = (grad == 2) ? 0.f : xx;
          else {
            if (grad == 0) yy = xx > 0.f ? xx : xx * alpha;
            else if (grad == 1) yy = rf[e] > 0.f ? xx : xx * alpha;
            else yy = 0.f;
          }
          o[e] = yy * scale;
        }
        reinterpret_cast<float4*>(y)[i] = make_float4(o[0], o[1], o[2], o[3]);
      }
    }
  } else {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
      const float xx = x[i] + (bias ? bias[(i / step_b) % size_b] : 0.f);
      const float rf = ref ? ref[i] : 0.f;
      float yy;
      if (act == 1) yy = (grad == 2) ? 0.f : xx;
      else {
        if (grad == 0) yy = xx > 0.f ? xx : xx * alpha;
        else if (grad == 1) yy = rf > 0.f ? xx : xx * alpha;
        else yy = 0.f;
      }
      y[i] = yy * scale;
    }
  }
}

}  // namespace

extern "C" int indm_upfirdn2d_f32(const float* x, const float* k, float* y, int64_t major, int in_h, int in_w, int kh, int kw,
                                  int up_x, int up_y, int down_x, int down_y, int pad_x0, int pad_x1, int pad_y0, int pad_y1,
                                  void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  INDM_CHECK_ARG(x && k && y, "upfirdn2d: null pointer");
  INDM_CHECK_ARG(major >= 0 && in_h > 0 && in_w > 0 && kh > 0 && kw > 0, "upfirdn2d: bad extents");
  INDM_CHECK_ARG(up_x >= 1 && up_y >= 1 && down_x >= 1 && down_y >= 1, "upfirdn2d: up/down must be >= 1");
  INDM_CHECK_ARG(kh <= 64 && kw <= 64, "upfirdn2d: kernel larger than 64 taps per axis");
  UpfirdnParams p;
  p.in_h = in_h; p.in_w = in_w; p.kh = kh; p.kw = kw;
  p.up_x = up_x; p.up_y = up_y; p.down_x = down_x; p.down_y = down_y; p.pad_x0 = pad_x0; p.pad_y0 = pad_y0;
  p.out_h = (in_h * up_y + pad_y0 + pad_y1 - kh) / down_y + 1;
  p.out_w = (in_w * up_x + pad_x0 + pad_x1 - kw) / down_x + 1;
  INDM_CHECK_ARG(in_h * up_y + pad_y0 + pad_y1 - kh >= 0 && in_w * up_x + pad_x0 + pad_x1 - kw >= 0 && p.out_h > 0 && p.out_w > 0,
                 "upfirdn2d: empty output (%d x %d)", p.out_h, p.out_w);
  if (major == 0) return INDM_OK;
  // footprint of a TILE_H x TILE_W output tile in input coordinates (+2 slack for the ceil/floor at both ends)
  p.sm_h = ((TILE_H - 1) * down_y + kh - 1) / up_y + 2;
  p.sm_w = ((TILE_W - 1) * down_x + kw - 1) / up_x + 2;
  p.tiles_x = (p.out_w + TILE_W - 1) / TILE_W;
  p.tiles_y = (p.out_h + TILE_H - 1) / TILE_H;
  const size_t smem = ((size_t)p.sm_h * p.sm_w + (size_t)kh * kw) * sizeof(float);
  INDM_CHECK_ARG(smem <= 200 * 1024, "upfirdn2d: tile footprint %zu bytes exceeds shared memory", smem);
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(upfirdn2d_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) {
      indm_set_error("upfirdn2d: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
      return INDM_ERR_CUDA;
    }
  }
  if (kh == 4 && kw == 4 && up_x == up_y && down_x == down_y) {
    if (up_x == 2 && down_x == 1) return launch_upfirdn_spec<2, 1, 4>(x, k, y, major, p, stream);
    if (up_x == 1 && down_x == 2) return launch_upfirdn_spec<1, 2, 4>(x, k, y, major, p, stream);
    if (up_x == 1 && down_x == 1) return launch_upfirdn_spec<1, 1, 4>(x, k, y, major, p, stream);
  }
  long long gy = major;
  if (gy > 65535) gy = 65535;
  dim3 grid(p.tiles_x * p.tiles_y, (unsigned)gy);
  upfirdn2d_kernel<<<grid, 256, smem, stream>>>(x, k, y, major, p);
  INDM_CHECK_LAUNCH("upfirdn2d");
  return INDM_OK;
}

extern "C" int indm_bias_act_f32(const float* x, const float* bias, const float* ref, float* y, int64_t n, int size_b,
                                 int64_t step_b, int act, int grad, float alpha, float scale, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  INDM_CHECK_ARG(x && y && n >= 0, "bias_act: bad arguments");
  INDM_CHECK_ARG(act == 1 || act == 3, "bias_act: act must be 1 (linear) or 3 (lrelu)");
  INDM_CHECK_ARG(grad >= 0 && grad <= 2, "bias_act: grad must be 0, 1 or 2");
  INDM_CHECK_ARG(!bias || (size_b > 0 && step_b > 0), "bias_act: bias needs size_b, step_b");
  INDM_CHECK_ARG(grad != 1 || ref, "bias_act: grad=1 needs ref");
  if (n == 0) return INDM_OK;
  if (!bias) { size_b = 1; step_b = 1; }
  const bool vec = (n % 4 == 0) && (!bias || step_b % 4 == 0) && (((uintptr_t)x | (uintptr_t)y | (uintptr_t)ref) % 16 == 0);
  const long long work = vec ? n / 4 : n;
  long long blocks = (work + 255) / 256;
  const long long cap = (long long)indm_num_sms() * 16;
  if (blocks > cap) blocks = cap;
  if (vec)
    bias_act_kernel<true><<<(unsigned)blocks, 256, 0, stream>>>(x, bias, ref, y, n, size_b, step_b, act, grad, alpha, scale);
  else
    bias_act_kernel<false><<<(unsigned)blocks, 256, 0, stream>>>(x, bias, ref, y, n, size_b, step_b, act, grad, alpha, scale);
  INDM_CHECK_LAUNCH("bias_act");
  return INDM_OK;
}
