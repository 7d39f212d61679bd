// Implicit-GEMM convolution / GEMM on the 5th-gen tensor cores (tcgen05.mma, accumulators in TMEM, operands staged by
// TMA with the 128-byte swizzle).  Replaces, on the INDM hot path, every cuDNN / cuBLAS call the reference makes for
//   * 3x3 stride-1 pad-1 convolutions   (models/layers.py:118-124 ddpm_conv3x3, used by models/layerspp.py:238,246)
//   * 1x1 convolutions / NIN            (models/layers.py:100-105, :546-555; models/layerspp.py:82-85,248)
//   * the attention contractions        (models/layerspp.py:95,99) as per-image batched GEMMs
//   * the flow's 512x512 1x1 conv       (flow_models/wolf/flows/resflow/layers/base/lipschitz.py:434)
//
// Formulation: activations NHWC.  Output tile = 128 pixels (a BN x BH x BW box of the image grid) x BLOCK_N output
// channels.  K loop = taps x (Cin / KCHUNK): for every tap the A operand is the SAME 4-D TMA box shifted by (dy,dx) —
// out-of-bounds rows/cols are zero-filled by TMA, which is exactly the conv's zero padding — so no im2col buffer
// exists anywhere.  An optional second K segment (a2/b2) accumulates a 1x1 convolution of another tensor into the
// same accumulator (the res-block skip path Conv_2, models/layerspp.py:281-282), and the epilogue fuses bias,
// per-image (time-embedding) bias, FP32 residual add, the 1/sqrt(2) skip rescale and the output cast.
//
// One CTA per output tile, 128 threads: warp 0 lane 0 = TMA producer, warp 1 lane 0 = MMA issuer, then all four warps
// drain TMEM (warp w owns TMEM lanes 32w..32w+31 = tile rows).  Several CTAs are co-resident per SM so one CTA's
// epilogue overlaps another's main loop.
#include <cuda.h>

#include "../../include/indm_b200.h"
#include "common.cuh"
#include "tmap.cuh"

namespace {

constexpr int kTileM = 128;
constexpr int kMaxStages = 8;

struct IgemmParams {
  // geometry
  int N, H, W;           // image grid of the A operand (plain GEMM: N=1,H=1,W=M)
  int BW, BH, BN;        // box (pixels) loaded per tile; BW*BH*BN <= 128
  int tiles_x, tiles_y;  // tiles along W and H
  int Cout;
  int taps;              // 1 or 9
  int chunks1, chunks2;  // K chunks in segment 1 (per tap) and segment 2
  int batched_b;         // B third coordinate = image index instead of tap
  int stride;            // 1: 3x3 pad 1 (or 1x1);  2: 3x3 stride 2 pad 0 over an A grid of (2H+1) x (2W+1)
  int stages;
  // epilogue
  const float* bias;
  const float* rowbias;
  long long rowbias_ld;
  const float* residual;
  long long res_ld;
  const float* rowscale;  // per-image multiplier (head conv: -1/std or 1/sigma)
  float scale;
  float res_scale;        // multiplier of the residual term
  int act;                // 0 none, 1 Sin(x) = sin(2 pi x) / (2 pi)  (resflow activation), 2 ELU
  int pad;                // stride 2 only: zero padding (0 or 1) of the strided window
  const void* mul;        // optional elementwise multiplier of the result: NHWC operand dtype (mode 0) / NCHW fp32 (mode 1)
  long long mul_ld;
  void* aux_cos;          // optional: cos(2 pi v) of the pre-activation value v, NHWC operand dtype (the Sin derivative)
  float* out_f32;
  __nv_bfloat16* out_bf16;
  long long out_ld;
  int out_mode;   // 0 NHWC rows, 1 NCHW fp32, 2 NHWC rows + columns >= tcol0 written transposed per image (bf16)
  int tcol0;
  __nv_bfloat16* out_t;  // mode 2: [N][Cout - tcol0][H*W]
  int round_tf32_out;    // round f32 outputs to tf32 (they feed a tf32 MMA next)
  float* gn_partial;     // optional [N][gn_groups][2] atomically accumulated (sum, sumsq) of the stored values
  int gn_cpg;            // channels per group
  int gn_groups;
};

// per-(image, group) sum / sum-of-squares of one 32-column slab held in registers (static indexing only: a dynamic
// index would push the whole slab to local memory)
template <int CPG>
__device__ __forceinline__ void gn_accumulate(const float (&f)[32], bool row_ok, bool as_bf16, int lane, int c0, float* dst) {
#pragma unroll
  for (int g0 = 0; g0 < 32; g0 += CPG) {
    float s = 0.f, q = 0.f;
#pragma unroll
    for (int i = 0; i < CPG; ++i) {
      float t = f[g0 + i];
      if (as_bf16) t = __bfloat162float(__float2bfloat16_rn(t));
      s += t;
      q += t * t;
    }
    if (!row_ok) s = q = 0.f;
    s = warp_sum(s);
    q = warp_sum(q);
    if (lane == 0 && dst) {
      atomicAdd(dst + (c0 + g0) / CPG * 2, s);
      atomicAdd(dst + (c0 + g0) / CPG * 2 + 1, q);
    }
  }
}

template <int BLOCK_N, bool TF32>
__global__ void __launch_bounds__(128, 1)
igemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
             const __grid_constant__ CUtensorMap tmA2, const __grid_constant__ CUtensorMap tmB2, const IgemmParams p) {
  constexpr int KCHUNK = TF32 ? 32 : 64;           // elements per 128-byte swizzle row
  constexpr int A_BYTES = kTileM * 128;            // 16 KB
  constexpr int B_BYTES = BLOCK_N * 128;
  constexpr uint32_t IDESC = umma_idesc(TF32 ? 2u : 1u, 128u, (uint32_t)BLOCK_N);
  constexpr uint32_t TMEM_COLS = BLOCK_N < 32 ? 32 : BLOCK_N;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = smem;
  uint8_t* sB = smem + (size_t)p.stages * A_BYTES;
  // TF32 mode is error-compensated (3xTF32): every landed fp32 stage is split by two helper warps into hi = rna(v) (in place)
  // and lo = v - hi (second buffer), and the issuer accumulates a_hi*b_hi + a_lo*b_hi + a_hi*b_lo.
  uint8_t* sAlo = sB + (size_t)p.stages * B_BYTES;
  uint8_t* sBlo = sAlo + (TF32 ? (size_t)p.stages * A_BYTES : 0);
  uint64_t* bars = (uint64_t*)(sBlo + (TF32 ? (size_t)p.stages * B_BYTES : 0));
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + kMaxStages;
  uint64_t* acc_bar = bars + 2 * kMaxStages;
  uint32_t* tmem_slot = (uint32_t*)(bars + 2 * kMaxStages + 1);
  uint64_t* split_bar = bars + 2 * kMaxStages + 2;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  // ---- tile coordinates
  const int mt = blockIdx.x;
  const int tx = mt % p.tiles_x;
  const int ty = (mt / p.tiles_x) % p.tiles_y;
  const int tn = mt / (p.tiles_x * p.tiles_y);
  const int x0 = tx * p.BW, y0 = ty * p.BH, n0 = tn * p.BN;
  const int ncol0 = blockIdx.y * BLOCK_N;

  const int iters1 = p.taps * p.chunks1;
  const int iters = iters1 + p.chunks2;
  const uint32_t a_box_bytes = (uint32_t)(p.BW * p.BH * p.BN) * 128u;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    if (p.chunks2 > 0) {
      tma_prefetch_desc(&tmA2);
      tma_prefetch_desc(&tmB2);
    }
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
      if (TF32) mbar_init(&split_bar[s], 64);
    }
    mbar_init(acc_bar, 1);
    fence_mbar_init();
  }
  if (warp == 0) {
    tmem_alloc(tmem_slot, TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      // ================= TMA producer
      int stage = 0;
      uint32_t phase = 0;
      for (int it = 0; it < iters; ++it) {
        mbar_wait(&empty_bar[stage], phase ^ 1u);
        uint8_t* a_dst = sA + (size_t)stage * A_BYTES;
        uint8_t* b_dst = sB + (size_t)stage * B_BYTES;
        mbar_arrive_expect_tx(&full_bar[stage], a_box_bytes + (uint32_t)B_BYTES);
        if (it < iters1) {
          const int tap = it / p.chunks1;
          const int ch = it - tap * p.chunks1;
          int ay = y0, ax = x0;
          if (p.stride == 2) {
            // valid (pad 0) stride-2 window: input pixel (2y + ky, 2x + kx); the tensor map traverses with element stride 2
            ay = 2 * y0 - p.pad + (p.taps == 9 ? tap / 3 : 0);
            ax = 2 * x0 - p.pad + (p.taps == 9 ? tap % 3 : 0);
          } else if (p.taps == 9) {
            ay += tap / 3 - 1;
            ax += tap % 3 - 1;
          }
          tma_load_4d(a_dst, &tmA, &full_bar[stage], ch * KCHUNK, ax, ay, n0);
          tma_load_3d(b_dst, &tmB, &full_bar[stage], ch * KCHUNK, ncol0, p.batched_b ? n0 : tap);
        } else {
          const int ch = it - iters1;
          tma_load_4d(a_dst, &tmA2, &full_bar[stage], ch * KCHUNK, x0, y0, n0);
          tma_load_3d(b_dst, &tmB2, &full_bar[stage], ch * KCHUNK, ncol0, 0);
        }
        if (++stage == p.stages) {
          stage = 0;
          phase ^= 1u;
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (lane == 0) {
      // ================= MMA issuer
      int stage = 0;
      uint32_t phase = 0;
      for (int it = 0; it < iters; ++it) {
        mbar_wait(TF32 ? &split_bar[stage] : &full_bar[stage], phase);
        tc_fence_after();
        const uint64_t adesc = umma_desc_sw128(smem_u32(sA + (size_t)stage * A_BYTES));
        const uint64_t bdesc = umma_desc_sw128(smem_u32(sB + (size_t)stage * B_BYTES));
        if (TF32) {
          const uint64_t alo = umma_desc_sw128(smem_u32(sAlo + (size_t)stage * A_BYTES));
          const uint64_t blo = umma_desc_sw128(smem_u32(sBlo + (size_t)stage * B_BYTES));
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint64_t o = (uint64_t)(2 * k);
            umma_tf32(tmem_base, alo + o, bdesc + o, IDESC, (it | k) != 0);   // small terms first
            umma_tf32(tmem_base, adesc + o, blo + o, IDESC, 1u);
            umma_tf32(tmem_base, adesc + o, bdesc + o, IDESC, 1u);
          }
        } else {
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            // advance 32 bytes along K inside the 128-byte swizzle row: +2 in the (addr >> 4) field
            umma_f16(tmem_base, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), IDESC, (it | k) != 0);
          }
        }
        umma_commit(&empty_bar[stage]);  // frees this smem stage when the MMAs above have read it
        if (++stage == p.stages) {
          stage = 0;
          phase ^= 1u;
        }
      }
      umma_commit(acc_bar);  // accumulator complete
    }
    __syncwarp();
  } else if (TF32) {
    // ================= operand splitter (warps 2, 3): hi/lo decomposition of each landed stage, elementwise, so the
    // swizzled placement is irrelevant: lo lives at the same offset of the twin buffer
    const int t = threadIdx.x - 64;
    int stage = 0;
    uint32_t phase = 0;
    for (int it = 0; it < iters; ++it) {
      mbar_wait(&full_bar[stage], phase);
      float4* a = reinterpret_cast<float4*>(sA + (size_t)stage * A_BYTES);
      float4* al = reinterpret_cast<float4*>(sAlo + (size_t)stage * A_BYTES);
      for (int i = t; i < A_BYTES / 16; i += 64) {
        const float4 v = a[i];
        const float4 h = make_float4(round_tf32(v.x), round_tf32(v.y), round_tf32(v.z), round_tf32(v.w));
        a[i] = h;
        al[i] = make_float4(v.x - h.x, v.y - h.y, v.z - h.z, v.w - h.w);
      }
      float4* b = reinterpret_cast<float4*>(sB + (size_t)stage * B_BYTES);
      float4* bl = reinterpret_cast<float4*>(sBlo + (size_t)stage * B_BYTES);
      for (int i = t; i < B_BYTES / 16; i += 64) {
        const float4 v = b[i];
        const float4 h = make_float4(round_tf32(v.x), round_tf32(v.y), round_tf32(v.z), round_tf32(v.w));
        b[i] = h;
        bl[i] = make_float4(v.x - h.x, v.y - h.y, v.z - h.z, v.w - h.w);
      }
      fence_proxy_async_smem();   // generic-proxy writes -> visible to the tensor-core (async) proxy
      mbar_arrive(&split_bar[stage]);
      if (++stage == p.stages) {
        stage = 0;
        phase ^= 1u;
      }
    }
  }

  // ================= epilogue: all 4 warps
  mbar_wait(acc_bar, 0);
  tc_fence_after();

  const int r = threadIdx.x;  // tile row == TMEM lane
  const int bw = r % p.BW;
  const int bh = (r / p.BW) % p.BH;
  const int bn = r / (p.BW * p.BH);
  const int n = n0 + bn, y = y0 + bh, x = x0 + bw;
  const bool row_ok = (bn < p.BN) && (n < p.N) && (x < p.W) && (y < p.H);
  const long long pix = ((long long)n * p.H + y) * p.W + x;
  const float rs = (p.rowscale != nullptr && row_ok) ? p.rowscale[n] : 1.0f;
  const float scale = p.scale * rs;

#pragma unroll 1
  for (int j = 0; j < BLOCK_N / 32 + (BLOCK_N < 32 ? 1 : 0); ++j) {
    uint32_t v[32];
    tmem_ld_32x32(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(j * 32), v);
    tmem_ld_wait();
    const int c0 = ncol0 + j * 32;
    if (c0 >= p.Cout) continue;  // uniform across the CTA
    float f[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) f[i] = __uint_as_float(v[i]);
    const bool full = (c0 + 32 <= p.Cout);
    if (p.bias) {
      if (full) {
#pragma unroll
        for (int i = 0; i < 32; i += 4) {
          const float4 b = __ldg(reinterpret_cast<const float4*>(p.bias + c0 + i));
          f[i] += b.x; f[i + 1] += b.y; f[i + 2] += b.z; f[i + 3] += b.w;
        }
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i)
          if (c0 + i < p.Cout) f[i] += __ldg(p.bias + c0 + i);
      }
    }
    if (row_ok) {
      if (p.rowbias) {
        const float* rb = p.rowbias + (long long)n * p.rowbias_ld + c0;
        if (full) {
#pragma unroll
          for (int i = 0; i < 32; i += 4) {
            const float4 b = __ldg(reinterpret_cast<const float4*>(rb + i));
            f[i] += b.x; f[i + 1] += b.y; f[i + 2] += b.z; f[i + 3] += b.w;
          }
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (c0 + i < p.Cout) f[i] += __ldg(rb + i);
        }
      }
#pragma unroll
      for (int i = 0; i < 32; ++i) f[i] *= scale;
      if (p.residual) {
        if (p.out_mode == 1) {
          // NCHW residual (flow fixed-point update x <- y - g(x), flow_models/.../iresblock.py:78-88)
          const long long hw = (long long)p.H * p.W;
          const float* rr = p.residual + (long long)n * p.Cout * hw + (long long)y * p.W + x;
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (c0 + i < p.Cout) f[i] += p.res_scale * rr[(long long)(c0 + i) * hw];
        } else {
          const float* rr = p.residual + pix * p.res_ld + c0;
          if (full) {
#pragma unroll
            for (int i = 0; i < 32; i += 4) {
              const float4 b = *reinterpret_cast<const float4*>(rr + i);
              f[i] += p.res_scale * b.x; f[i + 1] += p.res_scale * b.y; f[i + 2] += p.res_scale * b.z; f[i + 3] += p.res_scale * b.w;
            }
          } else {
#pragma unroll
            for (int i = 0; i < 32; ++i)
              if (c0 + i < p.Cout) f[i] += p.res_scale * rr[i];
          }
        }
      }
      if (p.aux_cos) {
        // derivative of the Sin activation at the pre-activation value, kept for the VJP chain of the log-det estimators
        if (TF32) {
          float* dst = (float*)p.aux_cos + pix * p.out_ld + c0;
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (c0 + i < p.Cout) dst[i] = cosf(6.283185307179586f * f[i]);
        } else {
          __nv_bfloat16* dst = (__nv_bfloat16*)p.aux_cos + pix * p.out_ld + c0;
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (c0 + i < p.Cout) dst[i] = __float2bfloat16_rn(cosf(6.283185307179586f * f[i]));
        }
      }
      if (p.act == 1) {
#pragma unroll
        for (int i = 0; i < 32; ++i) f[i] = sinf(6.283185307179586f * f[i]) * 0.15915494309189535f;
      } else if (p.act == 2) {
#pragma unroll
        for (int i = 0; i < 32; ++i) f[i] = f[i] > 0.f ? f[i] : expm1f(f[i]);
      }
      if (p.mul) {
        if (p.out_mode == 1) {
          const long long hw = (long long)p.H * p.W;
          const float* mm = (const float*)p.mul + (long long)n * p.Cout * hw + (long long)y * p.W + x;
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (c0 + i < p.Cout) f[i] *= mm[(long long)(c0 + i) * hw];
        } else if (TF32) {
          const float* mm = (const float*)p.mul + pix * p.mul_ld + c0;
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (c0 + i < p.Cout) f[i] *= mm[i];
        } else {
          const __nv_bfloat16* mm = (const __nv_bfloat16*)p.mul + pix * p.mul_ld + c0;
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (c0 + i < p.Cout) f[i] *= __bfloat162float(mm[i]);
        }
      }
      if (p.round_tf32_out) {
#pragma unroll
        for (int i = 0; i < 32; ++i) f[i] = round_tf32(f[i]);
      }
      if (p.out_mode == 1) {
        // NCHW fp32 (network head): few channels, strided store
        const long long hw = (long long)p.H * p.W;
        const long long base = (long long)n * p.Cout * hw + (long long)y * p.W + x;
#pragma unroll
        for (int i = 0; i < 32; ++i)
          if (c0 + i < p.Cout) p.out_f32[base + (long long)(c0 + i) * hw] = f[i];
      } else {
        const bool transposed = (p.out_mode == 2) && (c0 >= p.tcol0);
        if (transposed) {
          const long long hw = (long long)p.H * p.W;
          const long long li = (long long)y * p.W + x;
          __nv_bfloat16* dst = p.out_t + ((long long)n * (p.Cout - p.tcol0) + (c0 - p.tcol0)) * hw + li;
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (c0 + i < p.Cout) dst[(long long)i * hw] = __float2bfloat16_rn(f[i]);
        } else {
          if (p.out_f32) {
            float* dst = p.out_f32 + pix * p.out_ld + c0;
            if (full) {
#pragma unroll
              for (int i = 0; i < 32; i += 4)
                *reinterpret_cast<float4*>(dst + i) = make_float4(f[i], f[i + 1], f[i + 2], f[i + 3]);
            } else {
#pragma unroll
              for (int i = 0; i < 32; ++i)
                if (c0 + i < p.Cout) dst[i] = f[i];
            }
          }
          if (p.out_bf16) {
            __nv_bfloat16* dst = p.out_bf16 + pix * p.out_ld + c0;
            if (full) {
#pragma unroll
              for (int i = 0; i < 32; i += 8) {
                uint4 q;
                q.x = pack_bf16x2(f[i], f[i + 1]);
                q.y = pack_bf16x2(f[i + 2], f[i + 3]);
                q.z = pack_bf16x2(f[i + 4], f[i + 5]);
                q.w = pack_bf16x2(f[i + 6], f[i + 7]);
                *reinterpret_cast<uint4*>(dst + i) = q;
              }
            } else {
#pragma unroll
              for (int i = 0; i < 32; ++i)
                if (c0 + i < p.Cout) dst[i] = __float2bfloat16_rn(f[i]);
            }
          }
        }
      }
    }
    if (p.gn_partial) {
      // GroupNorm statistics of the tensor just produced, accumulated per (image, group): removes the separate
      // statistics pass over HBM for the GroupNorm that consumes this output (models/layerspp.py:244,277).
      // Requires all 32 rows of a warp to belong to one image (BW*BH >= 32 or BN == 1) — checked on the host.
      const bool as_bf16 = p.out_bf16 && !p.out_f32;
      const int nn = n0 + (warp * 32) / (p.BW * p.BH);
      float* dst = (nn < p.N) ? p.gn_partial + (long long)nn * p.gn_groups * 2 : nullptr;
      switch (p.gn_cpg) {
        case 4: gn_accumulate<4>(f, row_ok, as_bf16, lane, c0, dst); break;
        case 8: gn_accumulate<8>(f, row_ok, as_bf16, lane, c0, dst); break;
        case 16: gn_accumulate<16>(f, row_ok, as_bf16, lane, c0, dst); break;
        default: gn_accumulate<32>(f, row_ok, as_bf16, lane, c0, dst); break;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, TMEM_COLS);
}

template <int BLOCK_N, bool TF32>
int launch_igemm(const CUtensorMap& a, const CUtensorMap& b, const CUtensorMap& a2, const CUtensorMap& b2, IgemmParams p,
                 int m_tiles, int n_tiles, cudaStream_t stream) {
  constexpr int A_BYTES = kTileM * 128;
  constexpr int B_BYTES = BLOCK_N * 128;
  const int stage_bytes = (A_BYTES + B_BYTES) * (TF32 ? 2 : 1);
  const int overhead = 1024 + (3 * kMaxStages + 2) * 8;
  const int iters = p.taps * p.chunks1 + p.chunks2;
  // aim for >= 2 co-resident CTAs per SM when the tile is small enough; never more stages than K iterations
  int budget = (BLOCK_N >= 256 || TF32) ? 200 * 1024 : 100 * 1024;
  int stages = (budget - overhead) / stage_bytes;
  if (stages > kMaxStages) stages = kMaxStages;
  if (stages > iters) stages = iters;
  if (stages < 1) stages = 1;
  p.stages = stages;
  const int smem = stages * stage_bytes + overhead;
  static int configured = -1;
  auto kern = igemm_kernel<BLOCK_N, TF32>;
  if (configured < smem) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) {
      indm_set_error("igemm: cudaFuncSetAttribute failed: %s", cudaGetErrorString(e));
      return INDM_ERR_CUDA;
    }
    configured = 227 * 1024;
  }
  kern<<<dim3(m_tiles, n_tiles), 128, smem, stream>>>(a, b, a2, b2, p);
  INDM_CHECK_LAUNCH("igemm");
  return INDM_OK;
}

int pow2_floor(int v) {
  int p = 1;
  while (p * 2 <= v) p *= 2;
  return p;
}
bool is_pow2(int v) { return v > 0 && (v & (v - 1)) == 0; }

}  // namespace

extern "C" int indm_igemm(const indm_igemm_t* d, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  INDM_CHECK_ARG(d != nullptr, "igemm: null descriptor");
  const bool tf32 = d->dtype == INDM_DTYPE_TF32;
  INDM_CHECK_ARG(d->dtype == INDM_DTYPE_BF16 || tf32, "igemm: dtype must be BF16 or TF32");
  const int esz = tf32 ? 4 : 2;
  const int kchunk = tf32 ? 32 : 64;
  INDM_CHECK_ARG(d->a && d->b, "igemm: null operand");
  INDM_CHECK_ARG(d->N > 0 && d->H > 0 && d->W > 0 && d->Cin > 0 && d->Cout > 0, "igemm: bad extents");
  INDM_CHECK_ARG(d->taps == 1 || d->taps == 9, "igemm: taps must be 1 or 9 (got %d)", d->taps);
  INDM_CHECK_ARG(!(d->batched_b && d->taps != 1), "igemm: batched B requires taps == 1");
  INDM_CHECK_ARG(d->out_f32 || d->out_bf16 || d->out_t, "igemm: no output");
  INDM_CHECK_ARG(d->act >= 0 && d->act <= 2, "igemm: act must be 0 (none), 1 (Sin) or 2 (ELU)");

  IgemmParams p{};
  p.N = d->N; p.H = d->H; p.W = d->W;
  // ---- spatial box of 128 pixels
  if (d->W >= 128 || (d->H == 1 && d->N == 1)) {
    // rows of one image line (or a plain [M,K] GEMM): 128 consecutive pixels, the ragged last tile is zero-filled / masked
    INDM_CHECK_ARG(d->H == 1 || d->W % 128 == 0, "igemm: W >= 128 needs H == 1 or W %% 128 == 0");
    p.BW = 128; p.BH = 1; p.BN = 1;
  } else {
    INDM_CHECK_ARG(is_pow2(d->W), "igemm: W < 128 must be a power of two (got %d)", d->W);
    p.BW = d->W;
    int rest = 128 / p.BW;
    if (d->H >= rest) {
      INDM_CHECK_ARG(d->H % rest == 0, "igemm: H=%d not divisible by tile height %d", d->H, rest);
      p.BH = rest; p.BN = 1;
    } else {
      INDM_CHECK_ARG(is_pow2(d->H), "igemm: small H must be a power of two (got %d)", d->H);
      p.BH = d->H;
      p.BN = rest / p.BH;
    }
  }
  if (d->batched_b) p.BN = 1;  // every tile row must belong to the image whose B matrix is loaded
  p.tiles_x = (d->W + p.BW - 1) / p.BW;
  p.tiles_y = d->H / p.BH;
  const int tiles_n = (d->N + p.BN - 1) / p.BN;
  const int m_tiles = p.tiles_x * p.tiles_y * tiles_n;
  p.Cout = d->Cout;
  p.taps = d->taps;
  p.chunks1 = (d->Cin + kchunk - 1) / kchunk;
  p.chunks2 = d->a2 ? (d->Cin2 + kchunk - 1) / kchunk : 0;
  p.batched_b = d->batched_b;
  p.stride = d->stride == 2 ? 2 : 1;
  INDM_CHECK_ARG(d->stride == 0 || d->stride == 1 || d->stride == 2, "igemm: stride must be 1 or 2");
  INDM_CHECK_ARG(p.stride == 1 || (!d->a2 && !d->batched_b && d->W < 128), "igemm: stride 2 excludes a second segment / batched B");
  INDM_CHECK_ARG(d->pad == 0 || (d->pad == 1 && p.stride == 2), "igemm: pad is 0, or 1 with stride 2");
  p.pad = d->pad;
  p.mul = d->mul;
  p.mul_ld = d->mul_ld ? d->mul_ld : d->Cout;
  p.aux_cos = d->aux_cos;
  INDM_CHECK_ARG(!d->aux_cos || d->out_mode == 0, "igemm: aux_cos needs out_mode 0");
  p.bias = d->bias;
  p.rowbias = d->rowbias; p.rowbias_ld = d->rowbias_ld;
  p.residual = d->residual; p.res_ld = d->res_ld;
  p.rowscale = d->rowscale;
  p.scale = d->scale;
  p.res_scale = d->res_scale;
  p.act = d->act;
  p.out_f32 = d->out_f32;
  p.out_bf16 = (__nv_bfloat16*)d->out_bf16;
  p.out_ld = d->out_ld;
  p.out_mode = d->out_mode;
  p.tcol0 = d->tcol0;
  p.out_t = (__nv_bfloat16*)d->out_t;
  p.round_tf32_out = 0;  // legacy flag, ignored: TF32 operands stay full fp32 (the kernel splits hi/lo itself)
  p.gn_partial = d->gn_partial;
  p.gn_cpg = d->gn_cpg;
  p.gn_groups = d->gn_groups;
  if (d->out_mode == 1) INDM_CHECK_ARG(d->out_f32 != nullptr, "igemm: out_mode 1 needs out_f32");
  if (d->out_mode == 2) INDM_CHECK_ARG(d->out_t != nullptr && d->tcol0 % 32 == 0, "igemm: out_mode 2 needs out_t, tcol0 %% 32 == 0");
  if (d->out_mode != 1 && (d->out_f32 || d->out_bf16))
    INDM_CHECK_ARG(d->out_ld >= 1, "igemm: out_ld missing");
  if (d->gn_partial) {
    INDM_CHECK_ARG((d->gn_cpg == 4 || d->gn_cpg == 8 || d->gn_cpg == 16 || d->gn_cpg == 32) && d->Cout % 32 == 0,
                   "igemm: fused GroupNorm statistics need cpg | 32 and Cout %% 32 == 0 (cpg=%d Cout=%d)", d->gn_cpg, d->Cout);
    INDM_CHECK_ARG(p.BN == 1 || p.BW * p.BH >= 32, "igemm: fused GroupNorm statistics need >= 32 pixels per image per tile");
  }

  // ---- tensor maps
  const CUtensorMapDataType dt = tf32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
  CUtensorMap tmA, tmB, tmA2, tmB2;
  {
    const long long ld = d->a_ld ? d->a_ld : d->Cin;
    // stride 2: the A grid is the (2H+1) x (2W+1) FIR-padded image (models/up_or_down_sampling.py:173-178), H x W the output grid
    const int aH = p.stride == 2 ? (d->a_H ? d->a_H : 2 * d->H + 1) : d->H, aW = p.stride == 2 ? (d->a_W ? d->a_W : 2 * d->W + 1) : d->W;
    const long long img = d->a_img_stride ? d->a_img_stride : (long long)aH * aW * ld;
    uint64_t dims[4] = {(uint64_t)d->Cin, (uint64_t)aW, (uint64_t)aH, (uint64_t)d->N};
    uint64_t str[3] = {(uint64_t)ld * esz, (uint64_t)aW * ld * esz, (uint64_t)img * esz};
    uint32_t box[4] = {(uint32_t)kchunk, (uint32_t)(p.BW * p.stride), (uint32_t)(p.BH * p.stride), (uint32_t)p.BN};
    uint32_t es[4] = {1u, (uint32_t)p.stride, (uint32_t)p.stride, 1u};
    int rc = indm_make_tmap(&tmA, dt, 4, d->a, dims, str, box, "igemm A", es);
    if (rc) return rc;
  }
  // pick BLOCK_N
  int block_n = d->block_n;
  if (block_n == 0) {
    if (d->Cout <= 32) block_n = 32;
    else if (d->Cout <= 64) block_n = 64;
    else if (d->Cout % 256 == 0 && (long long)m_tiles * (d->Cout / 256) >= 2LL * indm_num_sms()) block_n = 256;
    else block_n = 128;
  }
  INDM_CHECK_ARG(block_n == 32 || block_n == 64 || block_n == 128 || block_n == 256, "igemm: block_n %d unsupported", block_n);
  const int n_tiles = (d->Cout + block_n - 1) / block_n;
  {
    const long long ld = d->b_ld ? d->b_ld : d->Cin;
    const long long ts = d->b_tap_stride ? d->b_tap_stride : (long long)d->Cout * ld;
    const int third = d->batched_b ? d->N : d->taps;
    uint64_t dims[3] = {(uint64_t)d->Cin, (uint64_t)d->Cout, (uint64_t)third};
    uint64_t str[2] = {(uint64_t)ld * esz, (uint64_t)ts * esz};
    uint32_t box[3] = {(uint32_t)kchunk, (uint32_t)block_n, 1u};
    int rc = indm_make_tmap(&tmB, dt, 3, d->b, dims, str, box, "igemm B");
    if (rc) return rc;
  }
  if (d->a2) {
    INDM_CHECK_ARG(d->b2 && d->Cin2 > 0, "igemm: second K segment needs b2 and Cin2");
    const long long ld = d->a2_ld ? d->a2_ld : d->Cin2;
    uint64_t dims[4] = {(uint64_t)d->Cin2, (uint64_t)d->W, (uint64_t)d->H, (uint64_t)d->N};
    uint64_t str[3] = {(uint64_t)ld * esz, (uint64_t)d->W * ld * esz, (uint64_t)d->H * d->W * ld * esz};
    uint32_t box[4] = {(uint32_t)kchunk, (uint32_t)p.BW, (uint32_t)p.BH, (uint32_t)p.BN};
    int rc = indm_make_tmap(&tmA2, dt, 4, d->a2, dims, str, box, "igemm A2");
    if (rc) return rc;
    const long long ldb = d->b2_ld ? d->b2_ld : d->Cin2;
    uint64_t bdims[3] = {(uint64_t)d->Cin2, (uint64_t)d->Cout, 1};
    uint64_t bstr[2] = {(uint64_t)ldb * esz, (uint64_t)d->Cout * ldb * esz};
    uint32_t bbox[3] = {(uint32_t)kchunk, (uint32_t)block_n, 1u};
    rc = indm_make_tmap(&tmB2, dt, 3, d->b2, bdims, bstr, bbox, "igemm B2");
    if (rc) return rc;
  } else {
    tmA2 = tmA;
    tmB2 = tmB;
  }

#define INDM_LAUNCH(BN_)                                                                                \
  return tf32 ? launch_igemm<BN_, true>(tmA, tmB, tmA2, tmB2, p, m_tiles, n_tiles, stream)               \
              : launch_igemm<BN_, false>(tmA, tmB, tmA2, tmB2, p, m_tiles, n_tiles, stream)
  switch (block_n) {
    case 32: INDM_LAUNCH(32);
    case 64: INDM_LAUNCH(64);
    case 128: INDM_LAUNCH(128);
    default: INDM_LAUNCH(256);
  }
#undef INDM_LAUNCH
}
