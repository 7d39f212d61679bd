// Convolution weight gradient on the 5th-gen tensor cores.  Stands in for the cuDNN wgrad / cuBLAS calls autograd issues for
// nn.Conv2d, NIN and nn.Linear weights in `losses.py:250,304` (`torch.mean(losses).backward()`).
//
//   dW[o][c][t] = sum over pixels p of  dY[p][o] * X[p + delta_t][c]          (3x3 / pad 1: delta_t = (ky-1, kx-1); 1x1: t = 0)
//
// GEMM view: D[M = o][N = c] = sum_k A[k][m] B[k][n] with K = pixels.  Both operands are read from the SAME NHWC tensors and the
// SAME 4-D TMA boxes the forward pass uses (128 pixels x 64 channels, 128-byte swizzle): a box is a [K = 128 pixels][64 channels]
// tile with the channel (M / N) index contiguous, i.e. an "MN-major" UMMA operand — the instruction descriptor's a_major /
// b_major bits select it, so no transposed copy of activations or gradients is ever made.  Shifted boxes are zero-filled out of
// bounds by TMA = the convolution's zero padding.  K is split across CTAs (grid.y); partial products are accumulated into the
// FP32 gradient with atomics (red.global.add.f32).
#include <cuda.h>

#include "../../include/indm_b200.h"
#include "common.cuh"
#include "tmap.cuh"

namespace {

constexpr int kMaxStagesW = 4;

struct WgradParams {
  int N, H, W;
  int BW, BH, BN;
  int tiles_x, tiles_y, tiles_n;
  int Cout, Cin, taps;
  int m_tiles, n_tiles;
  int stages;
  float* dw;
  long long so, sc, st;   // element strides of dw along o, c, tap
  float scale;
};

// MN-major, 128-byte swizzle: 64-element (128 B) runs along M/N, 8 K-rows per 1024-byte atom (SBO), next 64 M/N elements LBO away
__device__ __forceinline__ uint64_t umma_desc_mn_sw128(uint32_t smem_addr, uint32_t lbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)(lbo_bytes >> 4) << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

template <int BLOCK_N, bool TF32>
__global__ void __launch_bounds__(128, 1)
wgrad_kernel(const __grid_constant__ CUtensorMap tmDY, const __grid_constant__ CUtensorMap tmX, const WgradParams p) {
  constexpr int KC = TF32 ? 32 : 64;          // channels per 128-byte row
  constexpr int UK = TF32 ? 8 : 16;           // K (pixels) per MMA
  constexpr int BOX_BYTES = 128 * 128;        // one TMA box: 128 pixels x 128 B
  constexpr int A_BOXES = 128 / KC, B_BOXES = BLOCK_N / KC;
  constexpr int STAGE_BYTES = (A_BOXES + B_BOXES) * BOX_BYTES;
  constexpr uint32_t IDESC = umma_idesc(TF32 ? 2u : 1u, 128u, (uint32_t)BLOCK_N) | (1u << 15) | (1u << 16);
  constexpr uint32_t TMEM_COLS = BLOCK_N < 32 ? 32 : BLOCK_N;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = (uint64_t*)(smem + (size_t)p.stages * STAGE_BYTES);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + kMaxStagesW;
  uint64_t* acc_bar = bars + 2 * kMaxStagesW;
  uint32_t* tmem_slot = (uint32_t*)(bars + 2 * kMaxStagesW + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int job = blockIdx.x;
  const int nt = job % p.n_tiles; job /= p.n_tiles;
  const int mt = job % p.m_tiles; job /= p.m_tiles;
  const int tap = job;
  const int o0 = mt * 128, c0 = nt * BLOCK_N;
  int dy = 0, dx = 0;
  if (p.taps == 9) {
    dy = tap / 3 - 1;
    dx = tap % 3 - 1;
  }
  const int ptiles = p.tiles_x * p.tiles_y * p.tiles_n;
  const int my_tiles = (ptiles - (int)blockIdx.y + (int)gridDim.y - 1) / (int)gridDim.y;   // pixel tiles blockIdx.y, += gridDim.y

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmDY);
    tma_prefetch_desc(&tmX);
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
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

  // Warp-uniform loops with elect.sync-predicated issue: the TMA / tcgen05.mma operands stay in uniform registers (inside a
  // lane-0 branch every issue went through an ELECT + R2UR waterfall, which bounded the K loop of igemm.cu the same way).
  if (warp == 0) {
    int stage = 0;
    uint32_t phase = 0;
    for (int i = 0; i < my_tiles; ++i) {
      const int pt = blockIdx.y + i * gridDim.y;
      const int x0 = (pt % p.tiles_x) * p.BW, y0 = ((pt / p.tiles_x) % p.tiles_y) * p.BH, n0 = (pt / (p.tiles_x * p.tiles_y)) * p.BN;
      mbar_wait(&empty_bar[stage], phase ^ 1u);
      if (elect_one()) {
        uint8_t* base = smem + (size_t)stage * STAGE_BYTES;
        mbar_arrive_expect_tx(&full_bar[stage], (uint32_t)((A_BOXES + B_BOXES) * p.BW * p.BH * p.BN * 128));
#pragma unroll
        for (int b = 0; b < A_BOXES; ++b) tma_load_4d(base + b * BOX_BYTES, &tmDY, &full_bar[stage], o0 + b * KC, x0, y0, n0);
#pragma unroll
        for (int b = 0; b < B_BOXES; ++b)
          tma_load_4d(base + (A_BOXES + b) * BOX_BYTES, &tmX, &full_bar[stage], c0 + b * KC, x0 + dx, y0 + dy, n0);
      }
      __syncwarp();
      if (++stage == p.stages) {
        stage = 0;
        phase ^= 1u;
      }
    }
  } else if (warp == 1) {
    int stage = 0;
    uint32_t phase = 0;
    const int ksteps = (p.BW * p.BH * p.BN) / UK;   // pixels actually present in a box (rows beyond are stale smem: skip them)
    for (int i = 0; i < my_tiles; ++i) {
      mbar_wait(&full_bar[stage], phase);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t a_addr = smem_u32(smem + (size_t)stage * STAGE_BYTES);
        const uint32_t b_addr = a_addr + A_BOXES * BOX_BYTES;
        for (int k = 0; k < ksteps; ++k) {
          const uint64_t adesc = umma_desc_mn_sw128(a_addr + (uint32_t)(k * UK * 128), BOX_BYTES);
          const uint64_t bdesc = umma_desc_mn_sw128(b_addr + (uint32_t)(k * UK * 128), BOX_BYTES);
          if (TF32) umma_tf32(tmem_base, adesc, bdesc, IDESC, (i | k) != 0);
          else umma_f16(tmem_base, adesc, bdesc, IDESC, (i | k) != 0);
        }
        umma_commit(&empty_bar[stage]);
        if (i + 1 == my_tiles) umma_commit(acc_bar);
      }
      __syncwarp();
      if (++stage == p.stages) {
        stage = 0;
        phase ^= 1u;
      }
    }
  }

  if (my_tiles > 0) {
    mbar_wait(acc_bar, 0);
    tc_fence_after();
    const int o = o0 + threadIdx.x;   // tile row == TMEM lane == output channel
#pragma unroll 1
    for (int j = 0; j < BLOCK_N / 32; ++j) {
      uint32_t v[32];
      tmem_ld_32x32(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(j * 32), v);
      tmem_ld_wait();
      if (o < p.Cout) {
        float* dst = p.dw + (long long)o * p.so + (long long)tap * p.st;
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const int c = c0 + j * 32 + i;
          if (c < p.Cin) atomicAdd(dst + (long long)c * p.sc, __uint_as_float(v[i]) * p.scale);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, TMEM_COLS);
}

template <int BLOCK_N, bool TF32>
int launch_wgrad(const CUtensorMap& dy, const CUtensorMap& x, WgradParams p, int ksplit, cudaStream_t stream) {
  constexpr int KC = TF32 ? 32 : 64;
  constexpr int STAGE_BYTES = (128 / KC + BLOCK_N / KC) * 128 * 128;
  const int overhead = 1024 + (2 * kMaxStagesW + 2) * 8;
  int stages = (220 * 1024 - overhead) / STAGE_BYTES;
  if (stages > kMaxStagesW) stages = kMaxStagesW;
  if (stages < 1) {
    indm_set_error("wgrad: tile does not fit shared memory");
    return INDM_ERR_UNSUPPORTED;
  }
  p.stages = stages;
  const int smem = stages * STAGE_BYTES + overhead;
  static bool configured = false;
  auto kern = wgrad_kernel<BLOCK_N, TF32>;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) {
      indm_set_error("wgrad: cudaFuncSetAttribute failed: %s", cudaGetErrorString(e));
      return INDM_ERR_CUDA;
    }
    configured = true;
  }
  dim3 grid((unsigned)(p.taps * p.m_tiles * p.n_tiles), (unsigned)ksplit);
  kern<<<grid, 128, smem, stream>>>(dy, x, p);
  INDM_CHECK_LAUNCH("wgrad");
  return INDM_OK;
}

bool is_pow2w(int v) { return v > 0 && (v & (v - 1)) == 0; }

// FP32 validation path (CUDA cores): the same sum with fp32 operands and fp32 FMAs.  MN-major TF32 tensor-core operands need
// the 32-byte-atom swizzle, which the forward TMA boxes do not use; the validation mode trades speed for exactness instead.
// grid (o tiles of 64, c tiles of 64, taps * ksplit), block 256: each thread owns a 4 x 4 patch of the 64 x 64 tile.
__global__ void __launch_bounds__(256) wgrad_f32_kernel(const float* __restrict__ dy, long long dy_ld, const float* __restrict__ x,
                                                        long long x_ld, int N, int H, int W, int Cout, int Cin, int taps, int ksplit,
                                                        float* __restrict__ dw, long long so, long long sc, long long st, float scale) {
  __shared__ float sdy[16][64 + 4], sx[16][64 + 4];
  const int tap = blockIdx.z / ksplit, ks = blockIdx.z % ksplit;
  const int o0 = blockIdx.x * 64, c0 = blockIdx.y * 64;
  int dyy = 0, dxx = 0;
  if (taps == 9) {
    dyy = tap / 3 - 1;
    dxx = tap % 3 - 1;
  }
  const long long P = (long long)N * H * W;
  const long long per = (P + ksplit - 1) / ksplit;
  const long long p0 = ks * per, p1 = min(P, p0 + per);
  const int tx = threadIdx.x % 16, ty = threadIdx.x / 16;   // tx -> c patch, ty -> o patch
  float acc[4][4] = {};
  for (long long pb = p0; pb < p1; pb += 16) {
    for (int i = threadIdx.x; i < 16 * 64; i += 256) {
      const int r = i / 64, cc = i % 64;
      const long long pp = pb + r;
      float a = 0.f, b = 0.f;
      if (pp < p1) {
        if (o0 + cc < Cout) a = dy[pp * dy_ld + o0 + cc];
        const int xx = (int)(pp % W), yy = (int)((pp / W) % H);
        const int sx_ = xx + dxx, sy_ = yy + dyy;
        if (c0 + cc < Cin && sx_ >= 0 && sx_ < W && sy_ >= 0 && sy_ < H) b = x[(pp + (long long)dyy * W + dxx) * x_ld + c0 + cc];
      }
      sdy[r][cc] = a;
      sx[r][cc] = b;
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < 16; ++r) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        a[i] = sdy[r][ty * 4 + i];
        b[i] = sx[r][tx * 4 + i];
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
      const int o = o0 + ty * 4 + i, c = c0 + tx * 4 + j;
      if (o < Cout && c < Cin) atomicAdd(dw + (long long)o * so + (long long)c * sc + (long long)tap * st, acc[i][j] * scale);
    }
}

}  // namespace

extern "C" int indm_conv_wgrad(const void* dy, int64_t dy_ld, const void* x, int64_t x_ld, int dtype, int N, int H, int W, int Cout,
                               int Cin, int taps, float* dw, int64_t stride_o, int64_t stride_c, int64_t stride_t, float scale,
                               void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  const bool tf32 = dtype == INDM_DTYPE_TF32 || dtype == INDM_DTYPE_F32;
  INDM_CHECK_ARG(dtype == INDM_DTYPE_BF16 || tf32, "wgrad: dtype must be BF16 or TF32");
  INDM_CHECK_ARG(dy && x && dw && N > 0 && H > 0 && W > 0 && Cout > 0 && Cin > 0, "wgrad: bad arguments");
  INDM_CHECK_ARG(taps == 1 || taps == 9, "wgrad: taps must be 1 or 9");
  if (tf32) {
    const long long P = (long long)N * H * W;
    const int jobs = ((Cout + 63) / 64) * ((Cin + 63) / 64) * taps;
    long long ksplit = (4LL * indm_num_sms() + jobs - 1) / jobs;
    if (ksplit > (P + 255) / 256) ksplit = (P + 255) / 256;
    if (ksplit < 1) ksplit = 1;
    INDM_CHECK_ARG((long long)taps * ksplit <= 65535, "wgrad: grid.z overflow");
    dim3 grid((Cout + 63) / 64, (Cin + 63) / 64, (unsigned)(taps * ksplit));
    wgrad_f32_kernel<<<grid, 256, 0, stream>>>((const float*)dy, dy_ld ? dy_ld : Cout, (const float*)x, x_ld ? x_ld : Cin, N, H, W, Cout,
                                               Cin, taps, (int)ksplit, dw, stride_o, stride_c, stride_t, scale);
    INDM_CHECK_LAUNCH("wgrad_f32");
    return INDM_OK;
  }
  const int esz = tf32 ? 4 : 2;
  const int kc = tf32 ? 32 : 64;
  WgradParams p{};
  p.N = N; p.H = H; p.W = W;
  if (W >= 128 || (H == 1 && N == 1)) {
    INDM_CHECK_ARG(H == 1 || W % 128 == 0, "wgrad: W >= 128 needs H == 1 or W %% 128 == 0");
    INDM_CHECK_ARG(W % 128 == 0, "wgrad: a plain [M,K] operand needs M %% 128 == 0 (stale shared-memory rows would enter the sum)");
    p.BW = 128; p.BH = 1; p.BN = 1;
  } else {
    INDM_CHECK_ARG(is_pow2w(W), "wgrad: W < 128 must be a power of two");
    p.BW = W;
    int rest = 128 / p.BW;
    if (H >= rest) {
      INDM_CHECK_ARG(H % rest == 0, "wgrad: H not divisible by the tile height");
      p.BH = rest; p.BN = 1;
    } else {
      INDM_CHECK_ARG(is_pow2w(H), "wgrad: small H must be a power of two");
      p.BH = H;
      p.BN = rest / p.BH;
    }
  }
  p.tiles_x = (W + p.BW - 1) / p.BW;
  p.tiles_y = H / p.BH;
  p.tiles_n = (N + p.BN - 1) / p.BN;
  p.Cout = Cout; p.Cin = Cin; p.taps = taps;
  p.dw = dw; p.so = stride_o; p.sc = stride_c; p.st = stride_t; p.scale = scale;
  int block_n = Cin > 128 ? 256 : (Cin > 64 ? 128 : 64);
  if (tf32 && block_n > 128) block_n = 128;   // fp32 operands: 4 + 8 boxes per stage would not fit twice
  p.m_tiles = (Cout + 127) / 128;
  p.n_tiles = (Cin + block_n - 1) / block_n;
  const int ptiles = p.tiles_x * p.tiles_y * p.tiles_n;
  const int jobs = taps * p.m_tiles * p.n_tiles;
  int ksplit = (2 * indm_num_sms() + jobs - 1) / jobs;
  if (ksplit > ptiles) ksplit = ptiles;
  if (ksplit < 1) ksplit = 1;
  if (ksplit > 65535) ksplit = 65535;

  const CUtensorMapDataType dt = tf32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
  CUtensorMap tmDY, tmX;
  {
    const long long ld = dy_ld ? dy_ld : Cout;
    uint64_t dims[4] = {(uint64_t)Cout, (uint64_t)W, (uint64_t)H, (uint64_t)N};
    uint64_t str[3] = {(uint64_t)ld * esz, (uint64_t)W * ld * esz, (uint64_t)H * W * ld * esz};
    uint32_t box[4] = {(uint32_t)kc, (uint32_t)p.BW, (uint32_t)p.BH, (uint32_t)p.BN};
    int rc = indm_make_tmap(&tmDY, dt, 4, dy, dims, str, box, "wgrad dY");
    if (rc) return rc;
  }
  {
    const long long ld = x_ld ? x_ld : Cin;
    uint64_t dims[4] = {(uint64_t)Cin, (uint64_t)W, (uint64_t)H, (uint64_t)N};
    uint64_t str[3] = {(uint64_t)ld * esz, (uint64_t)W * ld * esz, (uint64_t)H * W * ld * esz};
    uint32_t box[4] = {(uint32_t)kc, (uint32_t)p.BW, (uint32_t)p.BH, (uint32_t)p.BN};
    int rc = indm_make_tmap(&tmX, dt, 4, x, dims, str, box, "wgrad X");
    if (rc) return rc;
  }
#define INDM_WG(BN_) return tf32 ? launch_wgrad<BN_, true>(tmDY, tmX, p, ksplit, stream) : launch_wgrad<BN_, false>(tmDY, tmX, p, ksplit, stream)
  switch (block_n) {
    case 64: INDM_WG(64);
    case 128: INDM_WG(128);
    default: INDM_WG(256);
  }
#undef INDM_WG
}
