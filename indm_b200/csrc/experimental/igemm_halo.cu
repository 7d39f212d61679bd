// EXPERIMENT OF RECORD — not part of libindm_b200.so, not on any product path.  Build: `make -C indm_b200/csrc experimental`,
// run: `python tools/experimental/halo_conv_check.py` (INDM_HALO_VARIANT = 0 / 1).  The production kernel that grew out of it is
// csrc/igemm_halo.cu.
//
// Padded-pixel implicit GEMM for 3x3 stride-1 pad-1 convolutions in its simplest form (one CTA per tile, one accumulator, 4
// epilogue warps storing bf16 rows straight from registers).  Per 64-channel K chunk ONE TMA box [1][R + 2][W + 2][64] at
// (y0 - 1, -1) lands in shared memory as (R + 2)(W + 2) consecutive 128-byte rows; tap (ty, tx) of the 3x3 window is the same tile
// read from a start address advanced by (ty (W + 2) + tx) rows.  GEMM row m <-> output position (y0 + m / (W + 2), m % (W + 2)).
//
// What it established on a B200 (round 2):
//   * variant 1 (descriptor base-offset field left 0): rel-L2 1.7e-3 against torch's convolution on every shape = BF16 output
//     rounding.  tcgen05 reads a 128B-swizzled K-major tile correctly from a start address that is a multiple of 128 B but not of
//     1024 B with NO base offset: the swizzle is a function of the absolute shared-memory address, like TMA's.
//   * variant 0 (base-offset = (start >> 7) & 7, the reading of the PTX ISA text round 1 had assumed): rel-L2 0.9 — wrong.
//   * rows past the loaded box only ever contaminate junk rows (every MMA row is independent).
//   * even this form ran the 128 -> 128 32x32 layer at 439 TFLOP/s (production tap-shifted kernel: 623).
#include <cuda.h>
#include <cuda_bf16.h>
#include <cstdlib>

#include "../common.cuh"
#include "../tmap.cuh"

namespace {

constexpr int kTileM = 128;
constexpr int kBStages = 6;
constexpr int kBlockN = 128;
constexpr int kBBytes = kBlockN * 128;

struct HaloParams {
  int N, H, W, Wp, R, tiles_per_img, n_tiles, Cout, chunks;
  uint32_t halo_bytes;   // bytes reserved per A buffer (multiple of 1024): (2 Wp + 2 + 128) rows of 128 B
  uint32_t box_bytes;    // bytes one TMA box delivers: (R + 2) Wp rows of 128 B
  const float* bias;
  __nv_bfloat16* out;
  int variant;   // 0: descriptor base-offset field = (start >> 7) & 7;  1: base-offset 0 (swizzle taken from the absolute address)
};

__device__ __forceinline__ uint64_t desc_sw128_rowoff(uint32_t smem_addr) {
  // 128B-swizzle K-major descriptor whose start may sit at any 128-byte row of the 1024-byte swizzle atom
  return umma_desc_sw128(smem_addr) | ((uint64_t)((smem_addr >> 7) & 7u) << 49);
}

__global__ void __launch_bounds__(192, 1)
igemm_halo_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const HaloParams p) {
  constexpr uint32_t IDESC = umma_idesc(1u, 128u, (uint32_t)kBlockN);
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = smem;                                      // [2][halo_bytes]
  uint8_t* sB = smem + 2 * (size_t)p.halo_bytes;           // [kBStages][kBBytes]
  uint64_t* bars = (uint64_t*)(sB + (size_t)kBStages * kBBytes);
  uint64_t* a_full = bars;                 // [2]
  uint64_t* a_empty = bars + 2;            // [2]
  uint64_t* b_full = bars + 4;             // [kBStages]
  uint64_t* b_empty = bars + 4 + kBStages; // [kBStages]
  uint64_t* t_full = bars + 4 + 2 * kBStages;
  uint32_t* tmem_slot = (uint32_t*)(bars + 5 + 2 * kBStages);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int t = blockIdx.x;
  const int nt = t % p.n_tiles, mt = t / p.n_tiles;
  const int n = mt / p.tiles_per_img, y0 = (mt % p.tiles_per_img) * p.R;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&a_full[i], 1);
      mbar_init(&a_empty[i], 1);
    }
    for (int i = 0; i < kBStages; ++i) {
      mbar_init(&b_full[i], 1);
      mbar_init(&b_empty[i], 1);
    }
    mbar_init(t_full, 1);
    fence_mbar_init();
  }
  if (warp == 0) {
    tmem_alloc(tmem_slot, kBlockN);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ---- TMA producer: one halo box per chunk, nine weight taps per chunk
    int it = 0;
    for (int c = 0; c < p.chunks; ++c) {
      const int a = c & 1;
      mbar_wait(&a_empty[a], (((uint32_t)c >> 1) & 1u) ^ 1u);
      if (elect_one()) {
        mbar_arrive_expect_tx(&a_full[a], p.box_bytes);
        tma_load_4d(sA + (size_t)a * p.halo_bytes, &tmA, &a_full[a], c * 64, -1, y0 - 1, n);
      }
      __syncwarp();
      for (int tap = 0; tap < 9; ++tap, ++it) {
        const int s = it % kBStages;
        mbar_wait(&b_empty[s], (((uint32_t)(it / kBStages)) & 1u) ^ 1u);
        if (elect_one()) {
          mbar_arrive_expect_tx(&b_full[s], (uint32_t)kBBytes);
          tma_load_3d(sB + (size_t)s * kBBytes, &tmB, &b_full[s], c * 64, nt * kBlockN, tap);
        }
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    // ---- MMA issuer: per tap the SAME halo tile at a row offset
    int it = 0;
    for (int c = 0; c < p.chunks; ++c) {
      const int a = c & 1;
      mbar_wait(&a_full[a], ((uint32_t)c >> 1) & 1u);
      tc_fence_after();
      const uint32_t a_base = smem_u32(sA + (size_t)a * p.halo_bytes);
      for (int tap = 0; tap < 9; ++tap, ++it) {
        const int s = it % kBStages;
        mbar_wait(&b_full[s], ((uint32_t)(it / kBStages)) & 1u);
        tc_fence_after();
        if (elect_one()) {
          const int ty = tap / 3, tx = tap - 3 * ty;
          const uint32_t a_addr = a_base + (uint32_t)(ty * p.Wp + tx) * 128u;
          const uint64_t adesc = p.variant == 0 ? desc_sw128_rowoff(a_addr) : umma_desc_sw128(a_addr);
          const uint64_t bdesc = umma_desc_sw128(smem_u32(sB + (size_t)s * kBBytes));
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_f16(tmem_base, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), IDESC, (uint32_t)((it | k) != 0));
          umma_commit(&b_empty[s]);
          if (tap == 8) umma_commit(&a_empty[a]);
        }
        __syncwarp();
      }
    }
    if (elect_one()) umma_commit(t_full);
    __syncwarp();
  } else {
    // ---- epilogue: warp w drains TMEM lanes 32 (w % 4) .. + 31; lane = GEMM row
    const int q = warp & 3;
    const int m = q * 32 + lane;
    const int yy = m / p.Wp, xx = m - yy * p.Wp;
    const bool ok = xx < p.W && yy < p.R && (y0 + yy) < p.H;
    mbar_wait(t_full, 0u);
    tc_fence_after();
    __nv_bfloat16* dst = p.out + (((size_t)n * p.H + (size_t)(y0 + yy)) * p.W + xx) * p.Cout + (size_t)nt * kBlockN;
#pragma unroll 1
    for (int slab = 0; slab < kBlockN / 32; ++slab) {
      uint32_t v[32];
      tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(slab * 32), v);
      tmem_ld_wait();
      if (ok) {
        uint4 o[4];
        uint32_t* ow = reinterpret_cast<uint32_t*>(o);
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const int col = nt * kBlockN + slab * 32 + 2 * j;
          const float b0 = p.bias ? p.bias[col] : 0.f, b1 = p.bias ? p.bias[col + 1] : 0.f;
          ow[j] = pack_bf16x2(__uint_as_float(v[2 * j]) + b0, __uint_as_float(v[2 * j + 1]) + b1);
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) reinterpret_cast<uint4*>(dst + slab * 32)[j] = o[j];
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kBlockN);
  }
}

}  // namespace

// x [N][H][W][Cin] bf16, wpack [9][Cout][Cin] bf16 (the engine's tap-major weight pack), bias [Cout] fp32 or NULL,
// out [N][H][W][Cout] bf16.  Cin % 64 == 0, Cout % 128 == 0, W + 2 <= 128.
extern "C" int indm_exp_conv3x3_halo_bf16(const void* x, const void* wpack, const float* bias, void* out, int N, int H, int W,
                                          int Cin, int Cout, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  INDM_CHECK_ARG(x && wpack && out && N > 0 && H > 0 && W > 0, "exp_conv3x3_halo: bad arguments");
  INDM_CHECK_ARG(Cin % 64 == 0 && Cout % kBlockN == 0 && W + 2 <= kTileM, "exp_conv3x3_halo: need Cin %% 64 == 0, Cout %% 128 == 0, W <= 126");
  HaloParams p;
  p.N = N; p.H = H; p.W = W; p.Wp = W + 2;
  p.R = kTileM / p.Wp;
  p.tiles_per_img = (H + p.R - 1) / p.R;
  p.n_tiles = Cout / kBlockN;
  p.Cout = Cout;
  p.chunks = Cin / 64;
  p.box_bytes = (uint32_t)((p.R + 2) * p.Wp) * 128u;
  p.halo_bytes = (((uint32_t)(2 * p.Wp + 2 + kTileM) * 128u) + 1023u) & ~1023u;
  p.bias = bias;
  p.out = (__nv_bfloat16*)out;
  { const char* e = getenv("INDM_HALO_VARIANT"); p.variant = e ? atoi(e) : 1; }
  INDM_CHECK_ARG(p.R + 2 <= 256 && p.box_bytes <= p.halo_bytes, "exp_conv3x3_halo: box does not fit the halo buffer");
  CUtensorMap tmA, tmB;
  {
    const uint64_t dims[4] = {(uint64_t)Cin, (uint64_t)W, (uint64_t)H, (uint64_t)N};
    const uint64_t str[3] = {(uint64_t)Cin * 2, (uint64_t)W * Cin * 2, (uint64_t)H * W * Cin * 2};
    const uint32_t box[4] = {64u, (uint32_t)p.Wp, (uint32_t)(p.R + 2), 1u};
    int rc = indm_make_tmap(&tmA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, x, dims, str, box, "exp_conv3x3_halo A");
    if (rc != INDM_OK) return rc;
  }
  {
    const uint64_t dims[3] = {(uint64_t)Cin, (uint64_t)Cout, 9ull};
    const uint64_t str[2] = {(uint64_t)Cin * 2, (uint64_t)Cout * Cin * 2};
    const uint32_t box[3] = {64u, (uint32_t)kBlockN, 1u};
    int rc = indm_make_tmap(&tmB, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, wpack, dims, str, box, "exp_conv3x3_halo B");
    if (rc != INDM_OK) return rc;
  }
  const size_t smem = 1024 + 2 * (size_t)p.halo_bytes + (size_t)kBStages * kBBytes + 256;
  cudaError_t e = cudaFuncSetAttribute(igemm_halo_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) {
    indm_set_error("exp_conv3x3_halo: cudaFuncSetAttribute(%zu B): %s", smem, cudaGetErrorString(e));
    return INDM_ERR_CUDA;
  }
  const int grid = N * p.tiles_per_img * p.n_tiles;
  igemm_halo_kernel<<<grid, 192, smem, stream>>>(tmA, tmB, p);
  INDM_CHECK_LAUNCH("exp_conv3x3_halo");
  return INDM_OK;
}
