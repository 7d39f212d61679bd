// Fused single-head self-attention forward of AttnBlockpp (models/layerspp.py:88-104) for the shape the score networks use
// (16x16 maps: L = 256 positions, C = 256 channels, BF16):
//
//   w = softmax_k( q.k^T * C^-1/2 ),   h = w.v           (the two einsums and the softmax of layerspp.py:94-99)
//
// One CTA per (image, 128-query half).  Q.K^T accumulates in TMEM (128 lanes x 256 fp32 columns), the softmax is computed by the
// thread that owns the query row straight out of TMEM, the un-normalised probabilities go to shared memory as the BF16 A operand
// of the second product (over the space Q occupied), V is read in place from the q|k|v tensor as an MN-major B operand (no
// transposed copy), and the row sum divides the result in the epilogue.  Neither the fp32 score tensor nor the probabilities
// ever reach HBM.  Replaces four launches of the generic path (Q.K^T GEMM, row softmax, V transpose, P.V GEMM).
#include <cuda.h>

#include "../../include/indm_b200.h"
#include "common.cuh"
#include "tmap.cuh"

namespace {

constexpr int kL = 256, kC = 256;
constexpr int BOX = 128 * 128;                      // one TMA box: 128 rows x 64 bf16 (128 B), 128-byte swizzle
constexpr int Q_BYTES = 4 * BOX;                    // 128 queries x 256 channels; later the probabilities (128 x 256 keys)
constexpr int KV_BYTES = 8 * BOX;                   // K: 4 channel chunks x [256 keys x 128 B]; later V: 2 key halves x 4 channel boxes
constexpr uint32_t IDESC_S = umma_idesc(1u, 128u, 256u);                 // both operands K-major
constexpr uint32_t IDESC_O = umma_idesc(1u, 128u, 256u) | (1u << 16);    // B (= V) MN-major

__device__ __forceinline__ void tma_load_2d(void* smem_dst, const void* tmap, uint64_t* bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
                   smem_u32(smem_dst)),
               "l"(tmap), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
               : "memory");
}

// MN-major operand, 128-byte swizzle (same form as wgrad.cu): 64-element runs along N, 8 K-rows per 1024-byte atom, next 64 N
// elements lbo_bytes away
__device__ __forceinline__ uint64_t desc_mn_sw128(uint32_t smem_addr, uint32_t lbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)(lbo_bytes >> 4) << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

__global__ void __launch_bounds__(192, 1)
attention_fwd_kernel(const __grid_constant__ CUtensorMap tmQKV, __nv_bfloat16* __restrict__ out, float scale_log2e) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* sQ = smem;                 // Q, then P
  uint8_t* sKV = smem + Q_BYTES;      // K, then V
  uint64_t* bars = (uint64_t*)(smem + Q_BYTES + KV_BYTES);
  uint64_t* bar_qk = bars;            // Q and K landed
  uint64_t* bar_s = bars + 1;         // S = Q.K^T complete in TMEM (K's shared memory is free)
  uint64_t* bar_v = bars + 2;         // V landed
  uint64_t* bar_p = bars + 3;         // probabilities written (128 arrivals)
  uint64_t* bar_o = bars + 4;         // O = P.V complete
  uint32_t* tmem_slot = (uint32_t*)(bars + 5);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n = blockIdx.x >> 1, qh = blockIdx.x & 1;
  const int row0 = n * kL;            // first row of this image in the [N L, 3C] tensor

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmQKV);
    mbar_init(bar_qk, 1);
    mbar_init(bar_s, 1);
    mbar_init(bar_v, 1);
    mbar_init(bar_p, 128);
    mbar_init(bar_o, 1);
    fence_mbar_init();
  }
  if (warp == 0) {
    tmem_alloc(tmem_slot, 256);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  pdl_trigger();
  pdl_wait();

  if (warp == 0) {
    if (elect_one()) {
      mbar_arrive_expect_tx(bar_qk, Q_BYTES + KV_BYTES);
#pragma unroll
      for (int cc = 0; cc < 4; ++cc) {
        tma_load_2d(sQ + cc * BOX, &tmQKV, bar_qk, cc * 64, row0 + qh * 128);
        tma_load_2d(sKV + cc * 2 * BOX, &tmQKV, bar_qk, kC + cc * 64, row0);
        tma_load_2d(sKV + cc * 2 * BOX + BOX, &tmQKV, bar_qk, kC + cc * 64, row0 + 128);
      }
    }
    __syncwarp();
    mbar_wait(bar_s, 0);              // the first product has consumed K: its space takes V
    if (elect_one()) {
      mbar_arrive_expect_tx(bar_v, KV_BYTES);
#pragma unroll
      for (int kh = 0; kh < 2; ++kh)
#pragma unroll
        for (int nc = 0; nc < 4; ++nc) tma_load_2d(sKV + (kh * 4 + nc) * BOX, &tmQKV, bar_v, 2 * kC + nc * 64, row0 + kh * 128);
    }
    __syncwarp();
  } else if (warp == 1) {
    mbar_wait(bar_qk, 0);
    tc_fence_after();
    if (elect_one()) {
      const uint32_t a = smem_u32(sQ), b = smem_u32(sKV);
#pragma unroll
      for (int cc = 0; cc < 4; ++cc)
#pragma unroll
        for (int kk = 0; kk < 4; ++kk)
          umma_f16(tmem, umma_desc_sw128(a + cc * BOX + kk * 32), umma_desc_sw128(b + cc * 2 * BOX + kk * 32), IDESC_S, (cc | kk) != 0);
      umma_commit(bar_s);
    }
    __syncwarp();
    mbar_wait(bar_v, 0);
    mbar_wait(bar_p, 0);
    tc_fence_after();
    if (elect_one()) {
      const uint32_t a = smem_u32(sQ), b = smem_u32(sKV);
#pragma unroll
      for (int k16 = 0; k16 < 16; ++k16) {          // 16 keys per MMA: P chunk k16 / 4, V key half k16 / 8
        const uint64_t ad = umma_desc_sw128(a + (k16 >> 2) * BOX + (k16 & 3) * 32);
        const uint64_t bd = desc_mn_sw128(b + (k16 >> 3) * 4 * BOX + (k16 & 7) * 16 * 128, BOX);
        umma_f16(tmem, ad, bd, IDESC_O, k16 != 0);
      }
      umma_commit(bar_o);
    }
    __syncwarp();
  } else {
    // softmax + epilogue: the thread that owns TMEM lane r = query row r of this half
    const int q4 = warp & 3;                         // the lane quarter this warp may address
    const int r = q4 * 32 + lane;
    const uint32_t t_row = tmem + ((uint32_t)(q4 * 32) << 16);
    mbar_wait(bar_s, 0);
    tc_fence_after();
    float m = -INFINITY;
#pragma unroll 1
    for (int j = 0; j < 8; ++j) {
      uint32_t v[32];
      tmem_ld_32x32(t_row + j * 32, v);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 32; ++i) m = fmaxf(m, __uint_as_float(v[i]));
    }
    const float mb = m * scale_log2e;
    float sum = 0.f;
    uint8_t* prow = sQ + r * 128;
#pragma unroll 1
    for (int j = 0; j < 8; ++j) {
      uint32_t v[32];
      tmem_ld_32x32(t_row + j * 32, v);
      tmem_ld_wait();
      uint32_t pk[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const float p0 = exp2f(fmaf(__uint_as_float(v[2 * i]), scale_log2e, -mb));
        const float p1 = exp2f(fmaf(__uint_as_float(v[2 * i + 1]), scale_log2e, -mb));
        sum += p0 + p1;
        pk[i] = pack_bf16x2(p0, p1);
      }
      // keys j*32 .. +32 = chunk j / 2, 16-byte units (j & 1) * 4 .. + 4 of this row, swizzled by the row
      uint8_t* base = prow + (j >> 1) * BOX;
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int unit = ((j & 1) * 4 + u) ^ (r & 7);
        *reinterpret_cast<uint4*>(base + unit * 16) = make_uint4(pk[4 * u], pk[4 * u + 1], pk[4 * u + 2], pk[4 * u + 3]);
      }
    }
    tc_fence_before();
    fence_proxy_async_smem();                        // generic-proxy writes of P -> visible to the tensor core's reads
    mbar_arrive(bar_p);
    mbar_wait(bar_o, 0);
    tc_fence_after();
    const float inv = 1.0f / sum;
    __nv_bfloat16* orow = out + ((long long)row0 + qh * 128 + r) * kC;
#pragma unroll 1
    for (int j = 0; j < 8; ++j) {
      uint32_t v[32];
      tmem_ld_32x32(t_row + j * 32, v);
      tmem_ld_wait();
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        uint4 o;
        o.x = pack_bf16x2(__uint_as_float(v[8 * u]) * inv, __uint_as_float(v[8 * u + 1]) * inv);
        o.y = pack_bf16x2(__uint_as_float(v[8 * u + 2]) * inv, __uint_as_float(v[8 * u + 3]) * inv);
        o.z = pack_bf16x2(__uint_as_float(v[8 * u + 4]) * inv, __uint_as_float(v[8 * u + 5]) * inv);
        o.w = pack_bf16x2(__uint_as_float(v[8 * u + 6]) * inv, __uint_as_float(v[8 * u + 7]) * inv);
        *reinterpret_cast<uint4*>(orow + j * 32 + u * 8) = o;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 256);
}

}  // namespace

extern "C" int indm_attention_fwd(const void* qkv, void* out, int64_t N, int L, int C, float scale, int dtype, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  INDM_CHECK_ARG(qkv && out && N > 0, "attention_fwd: bad arguments");
  if (dtype != INDM_DTYPE_BF16 || L != kL || C != kC) {
    indm_set_error("attention_fwd: the fused kernel covers BF16, L = 256, C = 256 (got dtype %d, L = %d, C = %d)", dtype, L, C);
    return INDM_ERR_UNSUPPORTED;
  }
  INDM_CHECK_ARG(((uintptr_t)out & 15) == 0 && N * 2 < (1ll << 31), "attention_fwd: out must be 16-byte aligned");
  CUtensorMap tm;
  const uint64_t dims[2] = {(uint64_t)(3 * C), (uint64_t)(N * L)};
  const uint64_t str[1] = {(uint64_t)(3 * C) * 2};
  const uint32_t box[2] = {64u, 128u};
  int rc = indm_make_tmap(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, qkv, dims, str, box, "attention_fwd q|k|v");
  if (rc) return rc;
  const size_t smem = 1024 + Q_BYTES + KV_BYTES + 64;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(attention_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) {
      indm_set_error("attention_fwd: cudaFuncSetAttribute failed: %s", cudaGetErrorString(e));
      return INDM_ERR_CUDA;
    }
    configured = true;
  }
  indm_launch_pdl(attention_fwd_kernel, dim3((unsigned)(2 * N)), dim3(192), smem, stream, tm, (__nv_bfloat16*)out,
                  scale * 1.4426950408889634f);
  INDM_CHECK_LAUNCH("attention_fwd");
  return INDM_OK;
}
