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
// Execution: a PERSISTENT, warp-specialised kernel — one CTA per SM loops over output tiles; warp 0 = TMA producer, warp 1 =
// MMA issuer (elect.sync-predicated, warp-uniform loops), 8 epilogue warps drain a DOUBLE-BUFFERED TMEM accumulator so the
// epilogue of tile i overlaps the main loop of tile i+1; the launches that fill the chip run as CTA pairs (cta_group::2,
// M = 256, B split across the pair).  The role table and the compile-time epilogue kinds are documented at igemm_kernel.
#include <cuda.h>
#include <cstdlib>

#include "../../include/indm_b200.h"
#include "common.cuh"
#include "tmap.cuh"
#include "igemm_halo.cuh"

namespace {

constexpr int kTileM = 128;
constexpr int kMaxStages = 8;

struct IgemmParams {
  // geometry
  int N, H, W;           // image grid of the A operand (plain GEMM: N=1,H=1,W=M)
  int BW, BH, BN;        // box (pixels) loaded per tile; BW*BH*BN <= 128; each a power of two
  int bw_shift, bh_shift;
  int tiles_x, tiles_y;  // tiles along W and H
  int tiles_n, n_tiles;  // tiles along the image dimension; tiles along Cout
  int Cout;
  int taps;              // 1 or 9
  int chunks1, chunks2;  // K chunks in segment 1 (per tap) and segment 2
  int batched_b;         // B third coordinate = image index instead of tap
  int ksplit;            // >1: split-K — work item t covers K-iteration slice (t % ksplit) of tile t / ksplit and writes raw fp32
                         // partial sums to out_f32 + slice * split_stride; indm_splitk_finish reduces them and applies the epilogue
  long long split_stride;
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
  int gn_goff;           // group index of output channel 0 inside the consumer's GroupNorm (non-zero for the second half of a concat)
  float* gn2_partial;    // second consumer of the same tensor (skip connections feed two GroupNorms with different group sizes)
  int gn2_cpg, gn2_groups, gn2_goff;
  int dbg;                // development probes (INDM_IGEMM_DBG): 1 = epilogue without global loads / stores
};

// ---------------------------------------------------------------- epilogue helpers
struct EpiRow {       // one output row (pixel) of the tile
  long long pix;      // flat NHWC pixel index
  int n, y, x;
  bool ok;
};

__device__ __forceinline__ EpiRow epi_row(const IgemmParams& p, int r, int n0, int y0, int x0) {
  EpiRow e;
  const int bw = r & (p.BW - 1);                       // box extents are powers of two
  const int bh = (r >> p.bw_shift) & (p.BH - 1);
  const int bn = r >> (p.bw_shift + p.bh_shift);
  e.n = n0 + bn; e.y = y0 + bh; e.x = x0 + bw;
  e.ok = (bn < p.BN) && (e.n < p.N) && (e.x < p.W) && (e.y < p.H);
  e.pix = ((long long)e.n * p.H + e.y) * p.W + e.x;
  return e;
}

// sin / cos of 2 pi v.  Production (BF16) path: exact range reduction to [-1/2, 1/2] then the SFU (abs error ~1e-6, far below
// the BF16 output resolution) — the precise sinf / cosf cost ~20 instructions each and made the flow's 512x512 GEMMs
// epilogue-bound.  Validation (TF32) path: sinpif / cospif (exact reduction, full precision).
template <bool PRECISE>
__device__ __forceinline__ float sin2pi(float v) {
  if (PRECISE) return sinpif(2.0f * v);
  return __sinf(6.283185307179586f * (v - rintf(v)));
}
template <bool PRECISE>
__device__ __forceinline__ float cos2pi(float v) {
  if (PRECISE) return cospif(2.0f * v);
  return __cosf(6.283185307179586f * (v - rintf(v)));
}
template <bool PRECISE>
__device__ __forceinline__ float act_apply(int act, float v) {
  if (act == 1) return sin2pi<PRECISE>(v) * 0.15915494309189535f;
  if (act == 2) return v > 0.f ? v : expm1f(v);
  return v;
}

// Persistent, warp-specialised kernel.  grid = min(#tiles, #SMs); every CTA walks tiles t = blockIdx.x, += gridDim.x, ...
//   warp 0      : TMA producer (one elected lane), smem ring of `stages` (A, B) tiles shared by all tiles of the CTA
//   warp 1      : tcgen05.mma issuer (one elected lane); accumulators double-buffered in TMEM (2 x BLOCK_N columns) so the
//                 main loop of tile j+1 overlaps the epilogue of tile j
//   warps 2..9  : epilogue, 8 warps (4 in TF32 mode): warp w owns TMEM lanes 32*(w%4).. and, of the two warps sharing a lane
//                 quarter, every other 32-column slab.  tcgen05.ld -> smem transpose (per-warp 4 KB, XOR-swizzled) so that every
//                 global access is a run of full 128-byte lines -> fused epilogue math -> stores.  (With one epilogue warp per
//                 scheduler the ~1000-instruction slab body ran at one dependent issue per ~5 cycles and bounded the kernel.)
//   last 2 warps: (TF32 only) hi/lo operand splitter for the error-compensated 3xTF32 product
constexpr int epi_warps(bool tf32) { return tf32 ? 4 : 8; }
constexpr int igemm_threads(bool tf32) { return 64 + 32 * epi_warps(tf32) + (tf32 ? 64 : 0); }

// KIND specialises the epilogue at compile time for the two shapes that carry >90% of the network's output bytes, so their
// slab body has no feature tests, no dead paths and ~half the instructions:
//   KIND 1: BF16 NHWC output ([+ bias] [+ per-image row bias]) * scale [+ GroupNorm statistics]   (Conv_0, q|k|v, P.V, dgrads)
//   KIND 2: FP32 NHWC output ([+ bias]) * scale [+ residual * res_scale]                         (Conv_1 [+ Conv_2], NIN_3, Q.K^T)
//   (no activation / cos side output / multiplier / per-image scale / NCHW / ragged-column paths in either)
//   KIND 3: BF16 NHWC output, ([+ bias] [+ row bias]) * scale -> [cos side output] -> [Sin / ELU]   (iResBlock branch GEMMs)
//   KIND 4: BF16 NHWC output, * scale * multiplier                                                 (iResBlock VJP chain)
//   (the generic epilogue keeps row bias + residual + multiplier registers live at once and spills: its reloads were 28 % of
//    the stall samples of the flow's 512 x 512 GEMMs, which ran 2.4 - 2.9x slower than the same GEMM with a plain epilogue)
//   KIND 0: every feature tested at run time.
//   CTA2: the CTAs of a 2-cluster (a CTA pair on one TPC) own two consecutive M tiles of the same N tile and issue ONE
//   tcgen05.mma.cta_group::2 of M = 256: each CTA stages its own A tile and half of the B tile (TMA credits the LEADER's full
//   barrier), the leader's MMA warp issues for both, tcgen05.commit multicasts stage-free / accumulator-ready to both CTAs, and
//   both CTAs' epilogue warps release the accumulator buffer on the leader's barrier.  Per-CTA operand bytes per MAC drop from
//   (128 + N) to (128 + N/2) per 128 x N x 64 chunk: the L2 -> SM operand stream was the measured bound of the one-CTA kernel.
template <int BLOCK_N, bool TF32, int KIND, bool CTA2>
__global__ void __launch_bounds__(igemm_threads(TF32), 1)
igemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
             const __grid_constant__ CUtensorMap tmA2, const __grid_constant__ CUtensorMap tmB2,
             const __grid_constant__ CUtensorMap tmOut, const IgemmParams p) {
  constexpr int KCHUNK = TF32 ? 32 : 64;           // elements per 128-byte swizzle row
  constexpr int A_BYTES = kTileM * 128;            // 16 KB
  constexpr int B_BYTES = (CTA2 ? BLOCK_N / 2 : BLOCK_N) * 128;   // rows of B staged by THIS CTA
  constexpr uint32_t IDESC = umma_idesc(TF32 ? 2u : 1u, CTA2 ? 256u : 128u, (uint32_t)BLOCK_N);
  static_assert(!(CTA2 && TF32), "the CTA-pair path is BF16 only");
  constexpr uint32_t ACC_COLS = BLOCK_N < 32 ? 32 : BLOCK_N;
  constexpr uint32_t TMEM_COLS = 2 * ACC_COLS;
  constexpr int NSLAB = BLOCK_N / 32 + (BLOCK_N < 32 ? 1 : 0);
  constexpr int EPI = epi_warps(TF32);
  constexpr int SLAB_STEP = EPI / 4;               // warps per TMEM lane quarter
  // KIND 5 / 6 = KIND 1 / 2 with the slab leaving through a TMA store: the warp keeps the accumulator's own layout (row = lane),
  // writes its 32 x 32 slab into a swizzled staging tile and one elected lane issues cp.async.bulk.tensor — no transpose, no
  // per-thread global addressing, no store instructions (the transposed epilogue spent ~1000 instructions per slab there)
  constexpr bool TSTORE = KIND == 5 || KIND == 6;
  constexpr int EK = KIND == 5 ? 1 : (KIND == 6 ? 2 : KIND);      // the epilogue feature set
  constexpr int STG_BYTES = TSTORE ? 8192 : 4096;                 // per epilogue warp (TSTORE: two buffers)

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = smem;
  uint8_t* sB = smem + (size_t)p.stages * A_BYTES;
  uint8_t* sAlo = sB + (size_t)p.stages * B_BYTES;
  uint8_t* sBlo = sAlo + (TF32 ? (size_t)p.stages * A_BYTES : 0);
  uint8_t* sStage = sBlo + (TF32 ? (size_t)p.stages * B_BYTES : 0);   // one 4 KB transpose buffer per epilogue warp
  uint64_t* bars = (uint64_t*)(sStage + EPI * STG_BYTES);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + kMaxStages;
  uint64_t* split_bar = bars + 2 * kMaxStages;
  uint64_t* tfull_bar = bars + 3 * kMaxStages;       // [2] accumulator buffer complete
  uint64_t* tempty_bar = bars + 3 * kMaxStages + 2;  // [2] accumulator buffer drained by the epilogue warps
  uint32_t* tmem_slot = (uint32_t*)(bars + 3 * kMaxStages + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  pdl_trigger();   // let the next kernel of the chain get scheduled; it waits in its own pdl_wait()

  const int iters1 = p.taps * p.chunks1;
  const int iters = iters1 + p.chunks2;
  const uint32_t a_box_bytes = (uint32_t)(p.BW * p.BH * p.BN) * 128u;
  const int m_tiles = p.tiles_x * p.tiles_y * p.tiles_n;
  // CTA2: work items are PAIRS of M tiles (2 mp, 2 mp + 1) x one N tile; CTA rank r of the cluster takes M tile 2 mp + r (an M tile
  // past the end is all out-of-bounds: TMA zero-fills it, the epilogue masks every row)
  const int cta_rank = CTA2 ? (int)cluster_ctarank() : 0;
  const int total_tiles = CTA2 ? ((m_tiles + 1) / 2) * p.n_tiles : m_tiles * p.n_tiles * p.ksplit;
  const int t_first = CTA2 ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int t_step = CTA2 ? (int)(gridDim.x >> 1) : (int)gridDim.x;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    if (TSTORE) tma_prefetch_desc(&tmOut);
    if (p.chunks2 > 0) {
      tma_prefetch_desc(&tmA2);
      tma_prefetch_desc(&tmB2);
    }
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
      if (TF32) mbar_init(&split_bar[s], 64);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&tfull_bar[b], 1);
      mbar_init(&tempty_bar[b], (CTA2 ? 64 : 32) * EPI);     // CTA2: the leader's barrier collects both CTAs' epilogue warps
    }
    fence_mbar_init();
  }
  if (warp == 0) {
    if (CTA2) {
      tmem_alloc_2sm(tmem_slot, TMEM_COLS);
      tmem_relinquish_2sm();
    } else {
      tmem_alloc(tmem_slot, TMEM_COLS);
      tmem_relinquish();
    }
  }
  tc_fence_before();
  if (CTA2) cluster_sync_all();      // the peer's barriers must be initialised before TMA / arrives target them
  else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();      // everything above overlapped the previous kernel's tail; its results are visible from here on

  if (warp == 0) {
    // ================= TMA producer.  The WHOLE warp runs the loop (warp-uniform control flow: coordinates, stage and phase live
    // in uniform registers) and one elected lane issues; inside a lane-0 branch every operand of UTMALDG / UTCHMMA went through
    // an ELECT + R2UR.BROADCAST waterfall (~90 dependent instructions per K iteration on the issuing thread, the measured
    // ~0.45 us fixed cost per iteration that bounded the main loop).
    int stage = 0;
    uint32_t phase = 0;
    for (int t = t_first; t < total_tiles; t += t_step) {
      const int ks = CTA2 ? 0 : t % p.ksplit, tt = CTA2 ? t : t / p.ksplit;
      const int nt = tt % p.n_tiles, mt = CTA2 ? 2 * (tt / p.n_tiles) + cta_rank : tt / p.n_tiles;
      const int x0 = (mt % p.tiles_x) * p.BW, y0 = ((mt / p.tiles_x) % p.tiles_y) * p.BH, n0 = (mt / (p.tiles_x * p.tiles_y)) * p.BN;
      const int ncol0 = nt * BLOCK_N + (CTA2 ? cta_rank * (BLOCK_N / 2) : 0);   // CTA2: this CTA's half of the B tile
      const int it_begin = CTA2 ? 0 : (int)((long long)iters * ks / p.ksplit), it_end = CTA2 ? iters : (int)((long long)iters * (ks + 1) / p.ksplit);
      // (tap, chunk) of the first iteration, then advanced incrementally: no division in the loop
      int tap = it_begin < iters1 ? it_begin / p.chunks1 : 0;
      int ch = it_begin < iters1 ? it_begin - tap * p.chunks1 : it_begin - iters1;
      int ty = p.taps == 9 ? tap / 3 : 0, tx = p.taps == 9 ? tap - 3 * ty : 0;
      for (int it = it_begin; it < it_end; ++it) {
        mbar_wait(&empty_bar[stage], phase ^ 1u);
        if (elect_one()) {
          uint8_t* a_dst = sA + (size_t)stage * A_BYTES;
          uint8_t* b_dst = sB + (size_t)stage * B_BYTES;
          const bool seg1 = it < iters1;
          int ay = y0, ax = x0;
          if (seg1) {
            if (p.stride == 2) {
              // strided window: input pixel (2y + ky - pad, 2x + kx - pad); the tensor map traverses with element stride 2
              ay = 2 * y0 - p.pad + ty;
              ax = 2 * x0 - p.pad + tx;
            } else if (p.taps == 9) {
              ay += ty - 1;
              ax += tx - 1;
            }
          }
          if (CTA2) {
            // one arrival (the leader's) per phase; both CTAs' four loads complete_tx on the leader's barrier
            if (cta_rank == 0) mbar_arrive_expect_tx(&full_bar[stage], 2u * (a_box_bytes + (uint32_t)B_BYTES));
            if (seg1) {
              tma_load_4d_2sm(a_dst, &tmA, &full_bar[stage], ch * KCHUNK, ax, ay, n0);
              tma_load_3d_2sm(b_dst, &tmB, &full_bar[stage], ch * KCHUNK, ncol0, tap);
            } else {
              tma_load_4d_2sm(a_dst, &tmA2, &full_bar[stage], ch * KCHUNK, x0, y0, n0);
              tma_load_3d_2sm(b_dst, &tmB2, &full_bar[stage], ch * KCHUNK, ncol0, 0);
            }
          } else {
            mbar_arrive_expect_tx(&full_bar[stage], a_box_bytes + (uint32_t)B_BYTES);
            if (seg1) {
              tma_load_4d(a_dst, &tmA, &full_bar[stage], ch * KCHUNK, ax, ay, n0);
              tma_load_3d(b_dst, &tmB, &full_bar[stage], ch * KCHUNK, ncol0, p.batched_b ? n0 : tap);
            } else {
              tma_load_4d(a_dst, &tmA2, &full_bar[stage], ch * KCHUNK, x0, y0, n0);
              tma_load_3d(b_dst, &tmB2, &full_bar[stage], ch * KCHUNK, ncol0, 0);
            }
          }
        }
        __syncwarp();
        // advance (tap, chunk): segment 1 walks taps x chunks1, then segment 2 walks chunks2
        if (++ch == (it < iters1 ? p.chunks1 : p.chunks2)) {
          ch = 0;
          if (it < iters1) {
            ++tap;
            if (++tx == 3) {
              tx = 0;
              ++ty;
            }
          }
        }
        if (++stage == p.stages) {
          stage = 0;
          phase ^= 1u;
        }
      }
    }
  } else if (warp == 1) {
    if (cta_rank == 0) {
      // ================= MMA issuer (CTA2: the pair's leader issues for both CTAs); warp-uniform loop, elected lane issues
      int stage = 0;
      uint32_t phase = 0;
      int j = 0;
      for (int t = t_first; t < total_tiles; t += t_step, ++j) {
        const int buf = j & 1;
        mbar_wait(&tempty_bar[buf], (((uint32_t)j >> 1) & 1u) ^ 1u);   // epilogue has drained this accumulator buffer
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)buf * ACC_COLS;
        const int ks = CTA2 ? 0 : t % p.ksplit;
        const int it_begin = CTA2 ? 0 : (int)((long long)iters * ks / p.ksplit), it_end = CTA2 ? iters : (int)((long long)iters * (ks + 1) / p.ksplit);
        for (int it = it_begin; it < it_end; ++it) {
          mbar_wait(TF32 ? &split_bar[stage] : &full_bar[stage], phase);
          tc_fence_after();
          if (elect_one()) {
            const uint64_t adesc = umma_desc_sw128(smem_u32(sA + (size_t)stage * A_BYTES));
            const uint64_t bdesc = umma_desc_sw128(smem_u32(sB + (size_t)stage * B_BYTES));
            if (TF32) {
              const uint64_t alo = umma_desc_sw128(smem_u32(sAlo + (size_t)stage * A_BYTES));
              const uint64_t blo = umma_desc_sw128(smem_u32(sBlo + (size_t)stage * B_BYTES));
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                const uint64_t o = (uint64_t)(2 * k);
                umma_tf32(d_tmem, alo + o, bdesc + o, IDESC, ((it - it_begin) | k) != 0);   // small terms first
                umma_tf32(d_tmem, adesc + o, blo + o, IDESC, 1u);
                umma_tf32(d_tmem, adesc + o, bdesc + o, IDESC, 1u);
              }
            } else if (CTA2) {
#pragma unroll
              for (int k = 0; k < 4; ++k)
                umma_f16_2sm(d_tmem, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), IDESC, (it | k) != 0);
            } else {
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                // advance 32 bytes along K inside the 128-byte swizzle row: +2 in the (addr >> 4) field
                umma_f16(d_tmem, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), IDESC, ((it - it_begin) | k) != 0);
              }
            }
            if (CTA2) umma_commit_2sm(&empty_bar[stage]);   // frees this smem stage in both CTAs
            else umma_commit(&empty_bar[stage]);            // frees this smem stage when the MMAs above have read it
            if (it + 1 == it_end) {
              if (CTA2) umma_commit_2sm(&tfull_bar[buf]);   // accumulator complete, both CTAs' epilogues
              else umma_commit(&tfull_bar[buf]);
            }
          }
          __syncwarp();
          if (++stage == p.stages) {
            stage = 0;
            phase ^= 1u;
          }
        }
      }
    }
  } else if (warp >= 2 + EPI) {
    if (TF32) {
      // ================= operand splitter: hi/lo decomposition of each landed stage, elementwise, so the swizzled
      // placement is irrelevant: lo lives at the same offset of the twin buffer
      const int tt = threadIdx.x - (64 + 32 * EPI);
      int stage = 0;
      uint32_t phase = 0;
      for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
        const int ks = t % p.ksplit;
        const int it_begin = (int)((long long)iters * ks / p.ksplit), it_end = (int)((long long)iters * (ks + 1) / p.ksplit);
        for (int it = it_begin; it < it_end; ++it) {
          mbar_wait(&full_bar[stage], phase);
          float4* a = reinterpret_cast<float4*>(sA + (size_t)stage * A_BYTES);
          float4* al = reinterpret_cast<float4*>(sAlo + (size_t)stage * A_BYTES);
          for (int i = tt; i < A_BYTES / 16; i += 64) {
            const float4 v = a[i];
            const float4 h = make_float4(round_tf32(v.x), round_tf32(v.y), round_tf32(v.z), round_tf32(v.w));
            a[i] = h;
            al[i] = make_float4(v.x - h.x, v.y - h.y, v.z - h.z, v.w - h.w);
          }
          float4* b = reinterpret_cast<float4*>(sB + (size_t)stage * B_BYTES);
          float4* bl = reinterpret_cast<float4*>(sBlo + (size_t)stage * B_BYTES);
          for (int i = tt; i < B_BYTES / 16; i += 64) {
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
    }
  } else {
    // ================= epilogue warps
    const bool has_bias = p.bias != nullptr;
    const bool has_rowbias = (EK == 2 || EK == 4) ? false : (p.rowbias != nullptr);
    const bool has_res = (EK == 1 || EK == 3 || EK == 4) ? false : (p.residual != nullptr);
    const bool has_rowscale = EK ? false : (p.rowscale != nullptr);
    const bool has_aux = (EK == 0 || EK == 3) ? (p.aux_cos != nullptr) : false;
    const int act = (EK == 0 || EK == 3) ? p.act : 0;
    const bool has_mul = EK == 4 ? true : (EK == 0 ? (p.mul != nullptr) : false);
    const bool st_f32 = EK == 2 ? true : (EK == 0 ? p.out_f32 != nullptr : false);
    const bool st_bf16 = EK == 2 ? false : (EK == 0 ? p.out_bf16 != nullptr : true);
    const bool has_gn = p.gn_partial != nullptr;
    const int q = warp & 3;                       // TMEM lane quarter this warp may read
    const int half = (warp - 2) >> 2;             // which of the SLAB_STEP warps sharing the quarter: takes slabs half, half + SLAB_STEP, ...
    float4* stg = reinterpret_cast<float4*>(sStage + (size_t)(warp - 2) * STG_BYTES);
    int tbuf = 0;                                 // TSTORE: staging buffer of the next slab
    const int chunk = lane & 7, rsub = lane >> 3;  // transposed domain: 4 columns (chunk), rows it*4 + rsub
    const long long hw = (long long)p.H * p.W;
    int j = 0;
    for (int t = t_first; t < total_tiles; t += t_step, ++j) {
      const int buf = j & 1;
      const int ks = CTA2 ? 0 : t % p.ksplit, tt = CTA2 ? t : t / p.ksplit;
      const int nt = tt % p.n_tiles, mt = CTA2 ? 2 * (tt / p.n_tiles) + cta_rank : tt / p.n_tiles;
      const int x0 = (mt % p.tiles_x) * p.BW, y0 = ((mt / p.tiles_x) % p.tiles_y) * p.BH, n0 = (mt / (p.tiles_x * p.tiles_y)) * p.BN;
      const int ncol0 = nt * BLOCK_N;
      float* const out32 = p.out_f32 + (long long)ks * p.split_stride;     // split-K: this slice's partial-sum plane
      // row bookkeeping, once per tile.  Direct domain (NCHW / transposed outputs, coalesced along pixels): row = lane.
      // Transposed domain (NHWC outputs): rows it*4 + rsub, it = 0..7.  Invalid rows load from pixel 0 / image 0 (always
      // valid addresses) so the loads stay branch-free and can all be in flight together; their stores are predicated off.
      const EpiRow ed = epi_row(p, q * 32 + lane, n0, y0, x0);
      long long pixs[8];
      int nsafe[8];
      unsigned okmask = 0;
#pragma unroll
      for (int it = 0; it < 8; ++it) {
        const EpiRow e = epi_row(p, q * 32 + it * 4 + rsub, n0, y0, x0);
        pixs[it] = e.ok ? e.pix : 0;
        nsafe[it] = e.ok ? e.n : 0;
        okmask |= (e.ok ? 1u : 0u) << it;
      }
      mbar_wait(&tfull_bar[buf], ((uint32_t)j >> 1) & 1u);
      tc_fence_after();
      const uint32_t t_addr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)buf * ACC_COLS;

      if (half >= NSLAB) {                       // narrow tile: this warp has no slab, it only releases the buffer
        tc_fence_before();
        if (CTA2) mbar_arrive_leader(&tempty_bar[buf]);
        else mbar_arrive(&tempty_bar[buf]);
      }
#pragma unroll 1
      for (int sl = half; sl < NSLAB; sl += SLAB_STEP) {
        const int c0 = ncol0 + sl * 32;
        if (TSTORE) {
          // ======== row = lane epilogue with a TMA store (KIND 5: bf16 NHWC; KIND 6: fp32 NHWC + residual)
          if (c0 >= p.Cout) {                    // uniform: a slab past the last channel only releases the buffer
            if (sl + SLAB_STEP >= NSLAB) {
              tc_fence_before();
              if (CTA2) mbar_arrive_leader(&tempty_bar[buf]);
              else mbar_arrive(&tempty_bar[buf]);
            }
            continue;
          }
          // the residual row of this thread (32 consecutive floats) is requested before the accumulator leaves TMEM
          float4 rsd[8];
          if (EK == 2 && has_res && !(p.dbg & 1)) {
            const float* rp = p.residual + (ed.ok ? ed.pix : 0) * p.res_ld + c0;
#pragma unroll
            for (int j = 0; j < 8; ++j) rsd[j] = *reinterpret_cast<const float4*>(rp + 4 * j);
          }
          uint32_t v[32];
          tmem_ld_32x32(t_addr + (uint32_t)(sl * 32), v);
          tmem_ld_wait();
          if (sl + SLAB_STEP >= NSLAB) {
            tc_fence_before();
            if (CTA2) mbar_arrive_leader(&tempty_bar[buf]);
            else mbar_arrive(&tempty_bar[buf]);
          }
          if (p.dbg & 1) continue;               // development probe: accumulator drained, nothing computed or stored
          // the staging buffer about to be written was handed to a TMA store two slabs ago: wait until that store has read it
          if (lane == 0) tma_store_wait_read<1>();
          __syncwarp();
          uint8_t* sbuf = reinterpret_cast<uint8_t*>(stg) + tbuf * 4096;
          const float* rbp = (EK == 1 && has_rowbias) ? p.rowbias + (long long)(ed.ok ? ed.n : 0) * p.rowbias_ld + c0 : nullptr;
          const float okf = ed.ok ? 1.0f : 0.0f;
          float s4[8], q4[8];                    // per 4-channel chunk: sum / sum of squares of this row's stored values
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            float4 x = make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]), __uint_as_float(v[4 * j + 2]),
                                   __uint_as_float(v[4 * j + 3]));
            if (has_bias) {
              const float4 b = __ldg(reinterpret_cast<const float4*>(p.bias + c0 + 4 * j));
              x.x += b.x; x.y += b.y; x.z += b.z; x.w += b.w;
            }
            if (EK == 1 && has_rowbias) {
              const float4 b = __ldg(reinterpret_cast<const float4*>(rbp + 4 * j));
              x.x += b.x; x.y += b.y; x.z += b.z; x.w += b.w;
            }
            x.x *= p.scale; x.y *= p.scale; x.z *= p.scale; x.w *= p.scale;
            if (EK == 2 && has_res) {
              x.x += p.res_scale * rsd[j].x; x.y += p.res_scale * rsd[j].y; x.z += p.res_scale * rsd[j].z; x.w += p.res_scale * rsd[j].w;
            }
            if (has_gn) {
              s4[j] = okf * ((x.x + x.y) + (x.z + x.w));
              q4[j] = okf * ((x.x * x.x + x.y * x.y) + (x.z * x.z + x.w * x.w));
            }
            if (EK == 2) {
              // fp32 rows of 128 B, SWIZZLE_128B: 16-byte chunk j of row `lane` lives at chunk j ^ (lane & 7)
              *reinterpret_cast<float4*>(sbuf + lane * 128 + ((j ^ (lane & 7)) << 4)) = x;
            } else if (j & 1) {
              // bf16 rows of 64 B, SWIZZLE_64B: chunk (j >> 1) of row `lane` lives at chunk (j >> 1) ^ ((lane >> 1) & 3)
              uint4 o;
              o.x = pack_bf16x2(__uint_as_float(v[4 * j - 4]), __uint_as_float(v[4 * j - 3]));
              o.y = pack_bf16x2(__uint_as_float(v[4 * j - 2]), __uint_as_float(v[4 * j - 1]));
              o.z = pack_bf16x2(x.x, x.y);
              o.w = pack_bf16x2(x.z, x.w);
              *reinterpret_cast<uint4*>(sbuf + lane * 64 + ((((j >> 1) ^ (lane >> 1)) & 3) << 4)) = o;
            } else {
              // even chunk: park the finished values in v[] for the pack of the following odd chunk
              v[4 * j] = __float_as_uint(x.x); v[4 * j + 1] = __float_as_uint(x.y);
              v[4 * j + 2] = __float_as_uint(x.z); v[4 * j + 3] = __float_as_uint(x.w);
            }
          }
          fence_proxy_async_smem();              // generic-proxy smem writes -> visible to the TMA (async proxy)
          __syncwarp();
          if (lane == 0) {
            const int r0 = q * 32;
            tma_store_4d(&tmOut, sbuf, c0, x0 + (r0 & (p.BW - 1)), y0 + ((r0 >> p.bw_shift) & (p.BH - 1)),
                         n0 + (r0 >> (p.bw_shift + p.bh_shift)));
            tma_store_commit();
          }
          tbuf ^= 1;
          if (has_gn) {
            // 16 per-row partial sums -> sums over the warp's 32 rows by recursive halving: after the step with lane mask m a lane
            // keeps the half of the values selected by its bit m, so 8 + 4 + 2 + 1 shuffles leave ONE value per lane (lane bits
            // 4..1 = its index: bit 4 = sum / sum of squares, bits 3..1 = chunk), and a last exchange with lane ^ 1 completes it
            float a8[8];
            {
              const bool hi = (lane & 16) != 0;
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                const float send = hi ? s4[i] : q4[i];
                const float recv = __shfl_xor_sync(0xffffffffu, send, 16);
                a8[i] = (hi ? q4[i] : s4[i]) + recv;
              }
            }
            float a4[4];
            {
              const bool hi = (lane & 8) != 0;
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const float send = hi ? a8[i] : a8[i + 4];
                const float recv = __shfl_xor_sync(0xffffffffu, send, 8);
                a4[i] = (hi ? a8[i + 4] : a8[i]) + recv;
              }
            }
            float a2[2];
            {
              const bool hi = (lane & 4) != 0;
#pragma unroll
              for (int i = 0; i < 2; ++i) {
                const float send = hi ? a4[i] : a4[i + 2];
                const float recv = __shfl_xor_sync(0xffffffffu, send, 4);
                a2[i] = (hi ? a4[i + 2] : a4[i]) + recv;
              }
            }
            float a1;
            {
              const bool hi = (lane & 2) != 0;
              const float send = hi ? a2[0] : a2[1];
              const float recv = __shfl_xor_sync(0xffffffffu, send, 2);
              a1 = (hi ? a2[1] : a2[0]) + recv;
            }
            a1 += __shfl_xor_sync(0xffffffffu, a1, 1);
            // this lane now holds: (lane bit 4 ? sum of squares : sum) of chunk ((lane >> 1) & 7) over the warp's 32 rows
            const int chunk_id = (lane >> 1) & 7;
            const int isq = (lane >> 4) & 1;
            const int nn = n0 + (q * 32) / (p.BW * p.BH);
            const int cc = c0 + chunk_id * 4;
            {
              float t1 = a1;
              const int cq = p.gn_cpg >> 2;             // chunks per group: 1, 2, 4 or 8 (adjacent chunks = lane bits 1..3)
              for (int o = 1; o < cq; o <<= 1) t1 += __shfl_xor_sync(0xffffffffu, t1, o << 1);
              if ((lane & 1) == 0 && (chunk_id & (cq - 1)) == 0 && nn < p.N)
                atomicAdd(p.gn_partial + ((long long)nn * p.gn_groups + p.gn_goff + cc / p.gn_cpg) * 2 + isq, t1);
            }
            if (p.gn2_partial) {
              float t2 = a1;
              const int cq = p.gn2_cpg >> 2;
              for (int o = 1; o < cq; o <<= 1) t2 += __shfl_xor_sync(0xffffffffu, t2, o << 1);
              if ((lane & 1) == 0 && (chunk_id & (cq - 1)) == 0 && nn < p.N)
                atomicAdd(p.gn2_partial + ((long long)nn * p.gn2_groups + p.gn2_goff + cc / p.gn2_cpg) * 2 + isq, t2);
            }
          }
          continue;
        }
        const bool direct = KIND ? false : ((p.out_mode == 1) || (p.out_mode == 2 && c0 >= p.tcol0));
        const int c = c0 + chunk * 4;           // first of this thread's 4 columns in the transposed domain
        const bool fast = !direct && c0 < p.Cout && (KIND != 0 || c + 4 <= p.Cout);
        // Prefetch: every global operand of this slab's epilogue (row bias, residual, multiplier) is requested BEFORE the
        // accumulator leaves TMEM, so the L2 / HBM round trip overlaps tcgen05.ld and the smem transpose.
        float4 rb[8], rsd[8], mulf[8];
        uint2 mulh[8];
        if (fast && !(p.dbg & 1)) {
          if (has_rowbias) {
#pragma unroll
            for (int it = 0; it < 8; ++it) rb[it] = __ldg(reinterpret_cast<const float4*>(p.rowbias + (long long)nsafe[it] * p.rowbias_ld + c));
          }
          if (has_res) {
#pragma unroll
            for (int it = 0; it < 8; ++it) rsd[it] = *reinterpret_cast<const float4*>(p.residual + pixs[it] * p.res_ld + c);
          }
          if (has_mul) {
            if (TF32) {
#pragma unroll
              for (int it = 0; it < 8; ++it) mulf[it] = *reinterpret_cast<const float4*>((const float*)p.mul + pixs[it] * p.mul_ld + c);
            } else {
#pragma unroll
              for (int it = 0; it < 8; ++it) mulh[it] = *reinterpret_cast<const uint2*>((const __nv_bfloat16*)p.mul + pixs[it] * p.mul_ld + c);
            }
          }
        }
        uint32_t v[32];
        tmem_ld_32x32(t_addr + (uint32_t)(sl * 32), v);
        tmem_ld_wait();
        if (sl + SLAB_STEP >= NSLAB) {
          tc_fence_before();
          if (CTA2) mbar_arrive_leader(&tempty_bar[buf]);
          else mbar_arrive(&tempty_bar[buf]);   // this warp's share of the accumulator buffer has left TMEM
        }
        if (c0 >= p.Cout) continue;             // uniform across the CTA
        if (p.dbg & 1) continue;
        if (direct) {
          if (!ed.ok) continue;
          const float scale = p.scale * (p.rowscale != nullptr ? p.rowscale[ed.n] : 1.0f);
          const long long li = (long long)ed.y * p.W + ed.x;
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const int c = c0 + i;
            if (c >= p.Cout) break;
            float f = __uint_as_float(v[i]);
            if (p.bias) f += __ldg(p.bias + c);
            if (p.rowbias) f += __ldg(p.rowbias + (long long)ed.n * p.rowbias_ld + c);
            f *= scale;
            if (p.out_mode == 1) {
              // NCHW fp32 (network head; flow fixed-point update x <- y - g(x), iresblock.py:78-88)
              const long long o = ((long long)ed.n * p.Cout + c) * hw + li;
              if (p.residual) f += p.res_scale * p.residual[o];
              f = act_apply<TF32>(p.act, f);
              if (p.mul) f *= ((const float*)p.mul)[o];
              p.out_f32[o] = f;
            } else {
              // mode 2, columns >= tcol0: transposed per image (bf16)
              f = act_apply<TF32>(p.act, f);
              p.out_t[((long long)ed.n * (p.Cout - p.tcol0) + (c - p.tcol0)) * hw + li] = __float2bfloat16_rn(f);
            }
          }
          continue;
        }

        // ---- NHWC outputs.  row-major -> smem (row = lane), XOR swizzle on the 16-byte chunk index
#pragma unroll
        for (int ch = 0; ch < 8; ++ch)
          stg[lane * 8 + (ch ^ (lane & 7))] = make_float4(__uint_as_float(v[4 * ch]), __uint_as_float(v[4 * ch + 1]),
                                                          __uint_as_float(v[4 * ch + 2]), __uint_as_float(v[4 * ch + 3]));
        __syncwarp();
        float gs = 0.f, gq = 0.f;               // GroupNorm partial sums of this thread's 8 rows x 4 columns
        if (fast) {
          // ---- fast path: whole 4-column chunk valid (global operands were prefetched above)
          float4 f[8];
#pragma unroll
          for (int it = 0; it < 8; ++it) {
            const int r = it * 4 + rsub;
            f[it] = stg[r * 8 + (chunk ^ (r & 7))];
          }
          float4 bia = make_float4(0.f, 0.f, 0.f, 0.f);
          if (has_bias) bia = __ldg(reinterpret_cast<const float4*>(p.bias + c));
          // Phase 2: math
#pragma unroll
          for (int it = 0; it < 8; ++it) {
            float4 x = f[it];
            x.x += bia.x; x.y += bia.y; x.z += bia.z; x.w += bia.w;
            if (has_rowbias) { x.x += rb[it].x; x.y += rb[it].y; x.z += rb[it].z; x.w += rb[it].w; }
            const float sc = has_rowscale ? p.scale * p.rowscale[nsafe[it]] : p.scale;
            x.x *= sc; x.y *= sc; x.z *= sc; x.w *= sc;
            if (has_res) {
              x.x += p.res_scale * rsd[it].x; x.y += p.res_scale * rsd[it].y; x.z += p.res_scale * rsd[it].z; x.w += p.res_scale * rsd[it].w;
            }
            f[it] = x;
          }
          if (has_aux) {
            // derivative of the Sin activation at the pre-activation value, kept for the VJP chain of the log-det estimators
#pragma unroll
            for (int it = 0; it < 8; ++it) {
              if (!((okmask >> it) & 1u)) continue;
              const float4 cc = make_float4(cos2pi<TF32>(f[it].x), cos2pi<TF32>(f[it].y),
                                            cos2pi<TF32>(f[it].z), cos2pi<TF32>(f[it].w));
              if (TF32) *reinterpret_cast<float4*>((float*)p.aux_cos + pixs[it] * p.out_ld + c) = cc;
              else *reinterpret_cast<uint2*>((__nv_bfloat16*)p.aux_cos + pixs[it] * p.out_ld + c) = make_uint2(pack_bf16x2(cc.x, cc.y), pack_bf16x2(cc.z, cc.w));
            }
          }
          if (act) {
#pragma unroll
            for (int it = 0; it < 8; ++it) {
              f[it].x = act_apply<TF32>(act, f[it].x); f[it].y = act_apply<TF32>(act, f[it].y);
              f[it].z = act_apply<TF32>(act, f[it].z); f[it].w = act_apply<TF32>(act, f[it].w);
            }
          }
          if (has_mul) {
            if (TF32) {
#pragma unroll
              for (int it = 0; it < 8; ++it) { f[it].x *= mulf[it].x; f[it].y *= mulf[it].y; f[it].z *= mulf[it].z; f[it].w *= mulf[it].w; }
            } else {
#pragma unroll
              for (int it = 0; it < 8; ++it) {
                const float2 m01 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&mulh[it].x));
                const float2 m23 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&mulh[it].y));
                f[it].x *= m01.x; f[it].y *= m01.y; f[it].z *= m23.x; f[it].w *= m23.y;
              }
            }
          }
          // Phase 3: stores (predicated on row validity) — 8 lanes cover one 128-byte (fp32) / 64-byte (bf16) run of a row
#pragma unroll
          for (int it = 0; it < 8; ++it) {
            if (!((okmask >> it) & 1u)) continue;
            if (st_f32) *reinterpret_cast<float4*>(out32 + pixs[it] * p.out_ld + c) = f[it];
            if (st_bf16)
              *reinterpret_cast<uint2*>(p.out_bf16 + pixs[it] * p.out_ld + c) = make_uint2(pack_bf16x2(f[it].x, f[it].y), pack_bf16x2(f[it].z, f[it].w));
            if (has_gn) {
              // statistics of the fp32 values: the BF16 rounding of the stored copy is zero-mean noise of relative size 2^-9,
              // i.e. ~1e-6 relative on a group's mean / variance — not worth 64 conversions per slab
              const float4 tq = f[it];
              gs += (tq.x + tq.y) + (tq.z + tq.w);
              gq += (tq.x * tq.x + tq.y * tq.y) + (tq.z * tq.z + tq.w * tq.w);
            }
          }
        } else if (KIND == 0 && c < p.Cout) {
          // ---- ragged last chunk (Cout % 4 != 0): scalar, guarded
#pragma unroll
          for (int it = 0; it < 8; ++it) {
            if (!((okmask >> it) & 1u)) continue;
            const int r = it * 4 + rsub;
            const float4 f4 = stg[r * 8 + (chunk ^ (r & 7))];
            const float sc = p.scale * (p.rowscale != nullptr ? p.rowscale[nsafe[it]] : 1.0f);
#pragma unroll
            for (int k = 0; k < 3; ++k) {       // a ragged chunk has 1..3 valid columns
              if (c + k >= p.Cout) break;
              float x = k == 0 ? f4.x : (k == 1 ? f4.y : f4.z);
              if (p.bias) x += __ldg(p.bias + c + k);
              if (p.rowbias) x += __ldg(p.rowbias + (long long)nsafe[it] * p.rowbias_ld + c + k);
              x *= sc;
              if (p.residual) x += p.res_scale * p.residual[pixs[it] * p.res_ld + c + k];
              if (p.aux_cos) {
                if (TF32) ((float*)p.aux_cos)[pixs[it] * p.out_ld + c + k] = cos2pi<TF32>(x);
                else ((__nv_bfloat16*)p.aux_cos)[pixs[it] * p.out_ld + c + k] = __float2bfloat16_rn(cos2pi<TF32>(x));
              }
              x = act_apply<TF32>(p.act, x);
              if (p.mul) x *= TF32 ? ((const float*)p.mul)[pixs[it] * p.mul_ld + c + k] : __bfloat162float(((const __nv_bfloat16*)p.mul)[pixs[it] * p.mul_ld + c + k]);
              if (p.out_f32) out32[pixs[it] * p.out_ld + c + k] = x;
              if (p.out_bf16) p.out_bf16[pixs[it] * p.out_ld + c + k] = __float2bfloat16_rn(x);
            }
          }
        }
        if (has_gn) {
          // GroupNorm statistics of the tensor just produced, per (image, group): removes the separate statistics pass over
          // HBM for the GroupNorm that consumes this output (models/layerspp.py:244,277).  All 32 rows of a warp belong to
          // one image (checked on the host); Cout % 32 == 0, so whole slabs only.
          gs += __shfl_xor_sync(0xffffffffu, gs, 8);  gq += __shfl_xor_sync(0xffffffffu, gq, 8);
          gs += __shfl_xor_sync(0xffffffffu, gs, 16); gq += __shfl_xor_sync(0xffffffffu, gq, 16);
          // gs / gq now hold this warp's 32-row sums of one 4-channel chunk; fold chunks into the consumer's groups
          const int nn = n0 + (q * 32) / (p.BW * p.BH);
          {
            float s1 = gs, q1 = gq;
            const int cq = p.gn_cpg >> 2;             // chunks per group: 1, 2, 4 or 8
            for (int o = 1; o < cq; o <<= 1) {
              s1 += __shfl_xor_sync(0xffffffffu, s1, o);
              q1 += __shfl_xor_sync(0xffffffffu, q1, o);
            }
            if (rsub == 0 && (chunk & (cq - 1)) == 0 && nn < p.N) {
              float* dst = p.gn_partial + ((long long)nn * p.gn_groups + p.gn_goff + c / p.gn_cpg) * 2;
              atomicAdd(dst, s1);
              atomicAdd(dst + 1, q1);
            }
          }
          if (p.gn2_partial) {
            float s2 = gs, q2 = gq;
            const int cq = p.gn2_cpg >> 2;
            for (int o = 1; o < cq; o <<= 1) {
              s2 += __shfl_xor_sync(0xffffffffu, s2, o);
              q2 += __shfl_xor_sync(0xffffffffu, q2, o);
            }
            if (rsub == 0 && (chunk & (cq - 1)) == 0 && nn < p.N) {
              float* dst = p.gn2_partial + ((long long)nn * p.gn2_groups + p.gn2_goff + c / p.gn2_cpg) * 2;
              atomicAdd(dst, s2);
              atomicAdd(dst + 1, q2);
            }
          }
        }
        __syncwarp();   // staging buffer is rewritten by the next slab
      }
    }
  }

  if (TSTORE && warp >= 2 && warp < 2 + EPI && lane == 0) tma_store_wait_all<0>();   // this lane's bulk stores are complete
  tc_fence_before();
  if (CTA2) {
    cluster_sync_all();          // neither CTA may retire (or free TMEM) while the peer still reads its smem / signals its barriers
    if (warp == 0) tmem_dealloc_2sm(tmem_base, TMEM_COLS);
  } else {
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_base, TMEM_COLS);
  }
}


// ---------------------------------------------------------------- split-K finish: sum the partial planes, apply the epilogue
// One thread per (pixel, 4-channel quad).  Same arithmetic and the same order of epilogue operations as the GEMM kernel:
// v = (sum + bias + rowbias[n]) * scale * rowscale[n] + residual * res_scale -> fp32 / bf16 NHWC (+ GroupNorm partial sums).
__global__ void splitk_finish_kernel(const float* __restrict__ ws, int S, long long M, int Cout, long long hw, const float* __restrict__ bias,
                                     const float* __restrict__ rowbias, long long rowbias_ld, const float* __restrict__ rowscale,
                                     float scale, const float* __restrict__ residual, long long res_ld, float res_scale,
                                     float* __restrict__ out_f32, __nv_bfloat16* __restrict__ out_bf16, long long out_ld,
                                     float* __restrict__ gn_partial, int gn_cpg, int gn_groups) {
  const int Q = Cout >> 2;
  const long long total = M * Q;
  const long long plane = M * (long long)Cout;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long pix = i / Q;
    const int c = (int)(i % Q) * 4;
    const long long n = pix / hw;
    float4 acc = *reinterpret_cast<const float4*>(ws + pix * Cout + c);
    for (int s = 1; s < S; ++s) {
      const float4 v = *reinterpret_cast<const float4*>(ws + s * plane + pix * Cout + c);
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    if (bias) {
      const float4 b = __ldg(reinterpret_cast<const float4*>(bias + c));
      acc.x += b.x; acc.y += b.y; acc.z += b.z; acc.w += b.w;
    }
    if (rowbias) {
      const float4 b = __ldg(reinterpret_cast<const float4*>(rowbias + n * rowbias_ld + c));
      acc.x += b.x; acc.y += b.y; acc.z += b.z; acc.w += b.w;
    }
    const float sc = scale * (rowscale ? rowscale[n] : 1.0f);
    acc.x *= sc; acc.y *= sc; acc.z *= sc; acc.w *= sc;
    if (residual) {
      const float4 r = *reinterpret_cast<const float4*>(residual + pix * res_ld + c);
      acc.x += res_scale * r.x; acc.y += res_scale * r.y; acc.z += res_scale * r.z; acc.w += res_scale * r.w;
    }
    if (out_f32) *reinterpret_cast<float4*>(out_f32 + pix * out_ld + c) = acc;
    if (out_bf16) *reinterpret_cast<uint2*>(out_bf16 + pix * out_ld + c) = make_uint2(pack_bf16x2(acc.x, acc.y), pack_bf16x2(acc.z, acc.w));
    if (gn_partial) {
      // lanes of one group are adjacent (cpg / 4 of them, a power of two <= 8, never straddling a pixel since Cout % 32 == 0)
      float gs = (acc.x + acc.y) + (acc.z + acc.w);
      float gq = (acc.x * acc.x + acc.y * acc.y) + (acc.z * acc.z + acc.w * acc.w);
      const int cq = gn_cpg >> 2;
      for (int o = 1; o < cq; o <<= 1) {
        gs += __shfl_xor_sync(0xffffffffu, gs, o);
        gq += __shfl_xor_sync(0xffffffffu, gq, o);
      }
      if (((c >> 2) & (cq - 1)) == 0) {
        float* dst = gn_partial + (n * gn_groups + c / gn_cpg) * 2;
        atomicAdd(dst, gs);
        atomicAdd(dst + 1, gq);
      }
    }
  }
}

template <int BLOCK_N, bool TF32, int KIND, bool CTA2 = false>
int launch_igemm(const CUtensorMap& a, const CUtensorMap& b, const CUtensorMap& a2, const CUtensorMap& b2, const CUtensorMap& tout,
                 IgemmParams p, int m_tiles, int n_tiles, cudaStream_t stream) {
  constexpr int A_BYTES = kTileM * 128;
  constexpr int B_BYTES = (CTA2 ? BLOCK_N / 2 : BLOCK_N) * 128;
  const int stage_bytes = (A_BYTES + B_BYTES) * (TF32 ? 2 : 1);
  const int overhead = 1024 + epi_warps(TF32) * ((KIND == 5 || KIND == 6) ? 8192 : 4096) + (3 * kMaxStages + 6) * 8;
  // one persistent CTA per SM: the smem ring takes what the SM has
  int stages = (220 * 1024 - overhead) / stage_bytes;
  if (stages > kMaxStages) stages = kMaxStages;
  if (stages < 2) stages = 2;
  p.stages = stages;
  p.n_tiles = n_tiles;
  const int smem = stages * stage_bytes + overhead;
  static bool configured = false;
  auto kern = igemm_kernel<BLOCK_N, TF32, KIND, CTA2>;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) {
      indm_set_error("igemm: cudaFuncSetAttribute failed: %s", cudaGetErrorString(e));
      return INDM_ERR_CUDA;
    }
    configured = true;
  }
  if (CTA2) {
    const long long pairs = (long long)((m_tiles + 1) / 2) * n_tiles;
    const int half_sms = indm_num_sms() / 2;
    const int grid = 2 * (int)(pairs < half_sms ? pairs : half_sms);
    indm_launch_pdl_cluster2(kern, dim3(grid), dim3(igemm_threads(TF32)), (size_t)smem, stream, a, b, a2, b2, tout, p);
    INDM_CHECK_LAUNCH("igemm (CTA pair)");
    return INDM_OK;
  }
  const long long total = (long long)m_tiles * n_tiles * p.ksplit;
  const int grid = (int)(total < indm_num_sms() ? total : indm_num_sms());
  indm_launch_pdl(kern, dim3(grid), dim3(igemm_threads(TF32)), (size_t)smem, stream, a, b, a2, b2, tout, p);
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
  INDM_CHECK_ARG(is_pow2(p.BW) && is_pow2(p.BH), "igemm: internal: box extents must be powers of two");
  for (p.bw_shift = 0; (1 << p.bw_shift) < p.BW; ++p.bw_shift) {}
  for (p.bh_shift = 0; (1 << p.bh_shift) < p.BH; ++p.bh_shift) {}
  p.tiles_x = (d->W + p.BW - 1) / p.BW;
  p.tiles_y = d->H / p.BH;
  p.tiles_n = (d->N + p.BN - 1) / p.BN;
  const int m_tiles = p.tiles_x * p.tiles_y * p.tiles_n;
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
  p.gn_goff = d->gn_goff;
  p.gn2_partial = d->gn2_partial;
  p.gn2_cpg = d->gn2_cpg; p.gn2_groups = d->gn2_groups; p.gn2_goff = d->gn2_goff;
  static const int dbg_flags = []() { const char* e = getenv("INDM_IGEMM_DBG"); return e ? atoi(e) : 0; }();
  p.dbg = dbg_flags;
  INDM_CHECK_ARG(!d->gn2_partial || (d->gn_partial && (d->gn2_cpg == 4 || d->gn2_cpg == 8 || d->gn2_cpg == 16 || d->gn2_cpg == 32)),
                 "igemm: the second GroupNorm target needs the first one and cpg in {4, 8, 16, 32}");
  if (d->out_mode == 1) INDM_CHECK_ARG(d->out_f32 != nullptr, "igemm: out_mode 1 needs out_f32");
  if (d->out_mode == 2) INDM_CHECK_ARG(d->out_t != nullptr && d->tcol0 % 32 == 0, "igemm: out_mode 2 needs out_t, tcol0 %% 32 == 0");
  if (d->out_mode != 1 && (d->out_f32 || d->out_bf16))
    INDM_CHECK_ARG(d->out_ld >= 1, "igemm: out_ld missing");
  if (d->gn_partial) {
    INDM_CHECK_ARG((d->gn_cpg == 4 || d->gn_cpg == 8 || d->gn_cpg == 16 || d->gn_cpg == 32) && d->Cout % 32 == 0,
                   "igemm: fused GroupNorm statistics need cpg | 32 and Cout %% 32 == 0 (cpg=%d Cout=%d)", d->gn_cpg, d->Cout);
    INDM_CHECK_ARG(d->a_pp || p.BN == 1 || p.BW * p.BH >= 32, "igemm: fused GroupNorm statistics need >= 32 pixels per image per tile");
  }

  // ---- tensor maps
  const CUtensorMapDataType dt = tf32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
  CUtensorMap tmA, tmB, tmA2, tmB2, tmOut;
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
    // widest tile that still gives every SM work: the kernel is persistent (one CTA per SM), so launches with few tiles
    // trade MMA width for parallelism (4x4 / 8x8 feature maps: M = N*16 / N*64 rows only)
    const int sms = indm_num_sms();
    if (d->Cout <= 32) block_n = 32;
    else if (d->Cout <= 64) block_n = 64;
    // operand traffic, not the tensor pipe, bounds a 128 x 128 tile (64 FLOP per byte fetched from L2): prefer 128 x 256
    else if (d->Cout % 256 == 0 && (long long)m_tiles * (d->Cout / 256) * 2 >= sms) block_n = 256;
    else {
      block_n = 32;
      for (int bn = 128; bn >= 32; bn >>= 1) {
        if ((long long)m_tiles * ((d->Cout + bn - 1) / bn) * 4 >= 3LL * sms) {
          block_n = bn;
          break;
        }
      }
    }
  }
  // split-K for launches with too few output tiles to fill the chip (4x4 / 8x8 feature maps): K-iteration slices of each tile go
  // to different CTAs with the widest tile, raw partial sums land in the caller's workspace, a finish kernel reduces them
  p.ksplit = 1;
  p.split_stride = 0;
  const long long Mpix = (long long)d->N * d->H * d->W;
  if (d->splitk_ws && d->block_n == 0 && d->out_mode == 0 && !d->batched_b && !d->aux_cos && !d->mul && d->act == 0 &&
      d->Cout % 32 == 0 && d->Cout >= 128) {
    const int sms = indm_num_sms();
    const int bn = d->Cout % 256 == 0 ? 256 : 128;
    const long long tiles = (long long)m_tiles * ((d->Cout + bn - 1) / bn);
    const int iters = p.taps * p.chunks1 + p.chunks2;
    int ks = (int)(sms / tiles);
    if (ks > iters / 4) ks = iters / 4;
    while (ks > 1 && (long long)ks * Mpix * d->Cout * 4 > d->splitk_ws_bytes) --ks;
    if (tiles * 2 <= sms && ks >= 2) {
      p.ksplit = ks;
      p.split_stride = Mpix * d->Cout;
      block_n = bn;
      // the GEMM kernel only stores raw partial sums
      p.bias = p.rowbias = p.residual = p.rowscale = nullptr;
      p.scale = 1.0f;
      p.out_f32 = (float*)d->splitk_ws;
      p.out_bf16 = nullptr;
      p.out_ld = d->Cout;
      p.gn_partial = nullptr;
      p.gn2_partial = nullptr;
    }
  }
  INDM_CHECK_ARG(block_n == 32 || block_n == 64 || block_n == 128 || block_n == 256, "igemm: block_n %d unsupported", block_n);
  const int n_tiles = (d->Cout + block_n - 1) / block_n;
  // compile-time epilogue specialisation (see igemm_kernel)
  int kind = 0;
  const bool plain = !tf32 && d->out_mode == 0 && d->Cout % 32 == 0 && !d->rowscale && !d->aux_cos && d->act == 0 && !d->mul;
  if (plain && !d->residual && d->out_bf16 && !d->out_f32) kind = 1;
  if (plain && !d->rowbias && d->out_f32 && !d->out_bf16) kind = 2;
  // the two epilogues of the iResBlock branch / VJP chain (wide tiles only: idim 128 / 512)
  const bool flowish = !tf32 && d->out_mode == 0 && d->Cout % 32 == 0 && !d->rowscale && !d->residual && d->out_bf16 && !d->out_f32 &&
                       (block_n == 128 || block_n == 256);
  if (flowish && !d->mul && (d->act != 0 || d->aux_cos)) kind = 3;
  if (flowish && d->mul && !d->bias && !d->rowbias && d->act == 0 && !d->aux_cos) kind = 4;
  // 3x3 convolutions of wide feature maps: the padded-pixel kernel loads every activation box once per K chunk instead of once
  // per tap (igemm_halo.cu)
  if (d->a_pp) {
    INDM_CHECK_ARG(d->out_mode == 0 && !d->splitk_ws, "igemm: a_pp operands need out_mode 0 and no split-K");
    return indm_igemm_halo_flat(d, kind, stream_);
  }
  if (p.ksplit == 1 && d->out_mode == 0 && indm_halo_eligible(d, kind)) return indm_igemm_halo(d, kind, stream_);
  // TMA-store epilogue (KIND 5 / 6) for the plain kinds when every 32-row slab of the tile is a rectangular box of the output
  // grid: the tile decomposes exactly (BW BH BN = 128, all powers of two) and the output rows are 16-byte aligned
  static const bool tstore_enabled = []() { const char* e = getenv("INDM_IGEMM_TSTORE"); return !(e && e[0] == '0'); }();
  bool tstore = false;
  if (tstore_enabled && (kind == 1 || kind == 2) && p.ksplit == 1 && p.BW * p.BH * p.BN == 128) {
    const bool f32o = kind == 2;
    const void* obase = f32o ? (const void*)d->out_f32 : (const void*)d->out_bf16;
    const long long old_ = d->out_ld ? d->out_ld : d->Cout;
    const int oes = f32o ? 4 : 2;
    if (((uintptr_t)obase & 15) == 0 && (old_ * oes) % 16 == 0 &&
        (!d->residual || ((((uintptr_t)d->residual) & 15) == 0 && ((d->res_ld ? d->res_ld : d->Cout) % 4) == 0)) &&
        (!d->rowbias || ((((uintptr_t)d->rowbias) & 15) == 0 && (d->rowbias_ld % 4) == 0)) && (!d->bias || (((uintptr_t)d->bias) & 15) == 0)) {
      const int sbw = p.BW < 32 ? p.BW : 32;
      const int sbh = p.BH < 32 / sbw ? p.BH : 32 / sbw;
      const int sbn = 32 / (sbw * sbh);
      const uint64_t dims[4] = {(uint64_t)d->Cout, (uint64_t)d->W, (uint64_t)d->H, (uint64_t)d->N};
      const uint64_t str[3] = {(uint64_t)old_ * oes, (uint64_t)d->W * old_ * oes, (uint64_t)d->H * d->W * old_ * oes};
      const uint32_t box[4] = {32u, (uint32_t)sbw, (uint32_t)sbh, (uint32_t)sbn};
      int rc = indm_make_tmap(&tmOut, f32o ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, obase, dims, str, box,
                              "igemm out", nullptr, f32o ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B);
      if (rc) return rc;
      tstore = true;
      kind = kind == 1 ? 5 : 6;
    }
  }
  // CTA pairs (cta_group::2) for the launches that fill the chip: wide tiles of the plain BF16 convolutions
  static const bool cta2_enabled = []() { const char* e = getenv("INDM_IGEMM_CTA2"); return !(e && e[0] == '0'); }();
  static const int cta2_min_tiles = []() { const char* e = getenv("INDM_IGEMM_CTA2_MIN_TILES"); return e ? atoi(e) : 0; }();
  const bool cta2 = cta2_enabled && !tf32 && p.ksplit == 1 && !d->batched_b && (block_n == 128 || block_n == 256) &&
                    (long long)m_tiles * n_tiles >= (cta2_min_tiles > 0 ? cta2_min_tiles : indm_num_sms());
  const int b_rows = cta2 ? block_n / 2 : block_n;
  {
    const long long ld = d->b_ld ? d->b_ld : d->Cin;
    const long long ts = d->b_tap_stride ? d->b_tap_stride : (long long)d->Cout * ld;
    const int third = d->batched_b ? d->N : d->taps;
    uint64_t dims[3] = {(uint64_t)d->Cin, (uint64_t)d->Cout, (uint64_t)third};
    uint64_t str[2] = {(uint64_t)ld * esz, (uint64_t)ts * esz};
    uint32_t box[3] = {(uint32_t)kchunk, (uint32_t)b_rows, 1u};
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
    uint32_t bbox[3] = {(uint32_t)kchunk, (uint32_t)b_rows, 1u};
    rc = indm_make_tmap(&tmB2, dt, 3, d->b2, bdims, bstr, bbox, "igemm B2");
    if (rc) return rc;
  } else {
    tmA2 = tmA;
    tmB2 = tmB;
  }
  if (!tstore) tmOut = tmA;     // unused by the other kinds: any valid map

  if (p.ksplit > 1) {
    int rc = tf32 ? (block_n == 256 ? launch_igemm<256, true, 0>(tmA, tmB, tmA2, tmB2, tmOut, p, m_tiles, n_tiles, stream)
                                    : launch_igemm<128, true, 0>(tmA, tmB, tmA2, tmB2, tmOut, p, m_tiles, n_tiles, stream))
                  : (block_n == 256 ? launch_igemm<256, false, 0>(tmA, tmB, tmA2, tmB2, tmOut, p, m_tiles, n_tiles, stream)
                                    : launch_igemm<128, false, 0>(tmA, tmB, tmA2, tmB2, tmOut, p, m_tiles, n_tiles, stream));
    if (rc) return rc;
    const long long work = Mpix * (d->Cout / 4);
    long long blocks = (work + 255) / 256;
    const long long cap = (long long)indm_num_sms() * 8;
    if (blocks > cap) blocks = cap;
    splitk_finish_kernel<<<(unsigned)blocks, 256, 0, stream>>>((const float*)d->splitk_ws, p.ksplit, Mpix, d->Cout, (long long)d->H * d->W,
                                                              d->bias, d->rowbias, d->rowbias_ld, d->rowscale, d->scale, d->residual,
                                                              d->res_ld, d->res_scale, d->out_f32, (__nv_bfloat16*)d->out_bf16, d->out_ld,
                                                              d->gn_partial, d->gn_cpg, d->gn_groups);
    INDM_CHECK_LAUNCH("splitk_finish");
    return INDM_OK;
  }
  if (cta2) {
#define INDM_LAUNCH2(BN_)                                                                                       \
  if (kind == 5) return launch_igemm<BN_, false, 5, true>(tmA, tmB, tmA2, tmB2, tmOut, p, m_tiles, n_tiles, stream);    \
  if (kind == 6) return launch_igemm<BN_, false, 6, true>(tmA, tmB, tmA2, tmB2, tmOut, p, m_tiles, n_tiles, stream);    \
  if (kind == 1) return launch_igemm<BN_, false, 1, true>(tmA, tmB, tmA2, tmB2, tmOut, p, m_tiles, n_tiles, stream);    \
  if (kind == 2) return launch_igemm<BN_, false, 2, true>(tmA, tmB, tmA2, tmB2, tmOut, p, m_tiles, n_tiles, stream);    \
  if (kind == 3) return launch_igemm<BN_, false, 3, true>(tmA, tmB, tmA2, tmB2, tmOut, p, m_tiles, n_tiles, stream);    \
  if (kind == 4) return launch_igemm<BN_, false, 4, true>(tmA, tmB, tmA2, tmB2, tmOut, p, m_tiles, n_tiles, stream);    \
  return launch_igemm<BN_, false, 0, true>(tmA, tmB, tmA2, tmB2, tmOut, p, m_tiles, n_tiles, stream)
    if (block_n == 256) { INDM_LAUNCH2(256); }
    INDM_LAUNCH2(128);
#undef INDM_LAUNCH2
  }
#define INDM_LAUNCH(BN_)                                                                                  \
  if (tf32) return launch_igemm<BN_, true, 0>(tmA, tmB, tmA2, tmB2, tmOut, p, m_tiles, n_tiles, stream);          \
  if (kind == 5) return launch_igemm<BN_, false, 5>(tmA, tmB, tmA2, tmB2, tmOut, p, m_tiles, n_tiles, stream);    \
  if (kind == 6) return launch_igemm<BN_, false, 6>(tmA, tmB, tmA2, tmB2, tmOut, p, m_tiles, n_tiles, stream);    \
  if (kind == 1) return launch_igemm<BN_, false, 1>(tmA, tmB, tmA2, tmB2, tmOut, p, m_tiles, n_tiles, stream);    \
  if (kind == 2) return launch_igemm<BN_, false, 2>(tmA, tmB, tmA2, tmB2, tmOut, p, m_tiles, n_tiles, stream);    \
  return launch_igemm<BN_, false, 0>(tmA, tmB, tmA2, tmB2, tmOut, p, m_tiles, n_tiles, stream)
  switch (block_n) {
    case 32: INDM_LAUNCH(32);
    case 64: INDM_LAUNCH(64);
    case 128:
      if (kind == 3) return launch_igemm<128, false, 3>(tmA, tmB, tmA2, tmB2, tmOut, p, m_tiles, n_tiles, stream);
      if (kind == 4) return launch_igemm<128, false, 4>(tmA, tmB, tmA2, tmB2, tmOut, p, m_tiles, n_tiles, stream);
      INDM_LAUNCH(128);
    default:
      if (kind == 3) return launch_igemm<256, false, 3>(tmA, tmB, tmA2, tmB2, tmOut, p, m_tiles, n_tiles, stream);
      if (kind == 4) return launch_igemm<256, false, 4>(tmA, tmB, tmA2, tmB2, tmOut, p, m_tiles, n_tiles, stream);
      INDM_LAUNCH(256);
  }
#undef INDM_LAUNCH
}
