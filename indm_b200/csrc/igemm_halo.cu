// Padded-pixel ("halo") implicit GEMM for the 3x3 stride-1 pad-1 convolutions of wide feature maps (W = 32 / 64): the layers of
// models/layerspp.py:238,246 whose 128 x 128 tiles were bound by the L2 -> SM operand stream in the tap-shifted kernel (igemm.cu):
// there every (tap, 64-channel chunk) iteration re-loads a 16 KB A tile, 9 loads of nearly the same pixels per chunk.
//
// Formulation.  View an image as H rows of Wp = W + 2 padded pixels; GEMM row m of tile k is padded pixel p = 128 k + m of ONE image
// (y = p / Wp, x = p % Wp; x >= W and y >= H are junk rows: computed, never stored).  Per 64-channel chunk ONE TMA box
// [64 ch][Wp][rows][1] starting at (x = -1, y = y_first - 1) lands in shared memory as consecutive 128-byte rows in exactly that
// padded order (out-of-bounds pixels zero-filled = the convolution's padding), and tap (ty, tx) of the window is the SAME smem tile
// read from a start address advanced by (s0 + ty Wp + tx) rows, s0 = 128 k - y_first Wp: nine MMA groups per A load.  The
// 128-byte swizzle of TMA and of the tcgen05 matrix descriptor are both functions of the absolute shared-memory address, so a
// row-granular start offset needs NO descriptor base-offset (verified on B200: csrc/experimental/igemm_halo.cu).
// Operand bytes from L2 per 128 x 128 x 64 MMA group drop from 24 KB to ~11 KB; rows used: 1024 / (9 x 128) = 89 % at 32 x 32.
//
// Execution mirrors igemm.cu: persistent, warp-specialised (warp 0 TMA producer, warp 1 MMA issuer, 8 epilogue warps), TMEM
// accumulator double-buffered, CTA pairs (cta_group::2, M = 256): the two CTAs of a pair take the SAME tile index of two
// consecutive images, so both see the same geometry (s0, y_first) and one descriptor serves both; each stages its own halo box
// and half of the weight tile.  A ring (per chunk) and B ring (per tap) are separate.  An optional second K segment accumulates
// the res-block's 1x1 skip convolution (Conv_2) from another tensor through the same box with the centre offset.
// Epilogue: tcgen05.ld -> per-warp smem transpose -> bias / per-image bias / scale / fp32 residual / GroupNorm partial statistics
// with every global access a run of full 128-byte (fp32) / 64-byte (bf16) lines (row -> pixel by one division per thread per tile).
#include <cuda.h>
#include <cstdlib>

#include "../../include/indm_b200.h"
#include "common.cuh"
#include "tmap.cuh"
#include "igemm_halo.cuh"

namespace {

constexpr int kMaxA = 4, kMaxB = 12;

// FLAT (round 2, small feature maps): the A operand lives in the padded-pixel ("PP") layout — pixel (n, y, x) at row
// (n (H + 1) + y + 1) Wp + x + 1 of a zero-bordered buffer, written by indm_gn_apply_pp — so the WHOLE BATCH is one sequence of padded
// pixels: tile k = padded pixels 128 k .. + 127 regardless of image boundaries, one 2-D TMA box of 128 + 2 Wp + 2 rows per chunk,
// tap offset ty Wp + tx, both CTAs of a pair take consecutive tiles.  4x4 / 8x8 maps, whose tap-shifted tiles re-read a 16 KB A tile
// for every tap and sat on the per-SM L2 -> SM ingest limit (~0.5 us per K iteration, 0.07 - 0.35 of the tensor peak), load A once
// per chunk.  Border rows are computed and dropped (8x8: 64 of 90 padded pixels are real).
template <int BLOCK_N, int EK, bool FLAT>     // EK 1: bf16 NHWC out (+ bias + per-image bias); EK 2: fp32 NHWC out (+ bias + residual)
__global__ void __launch_bounds__(320, 1)
igemm_halo_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                  const __grid_constant__ CUtensorMap tmA2, const __grid_constant__ CUtensorMap tmB2, const HaloParams p) {
  constexpr int B_BYTES = (BLOCK_N / 2) * 128;
  constexpr uint32_t IDESC = umma_idesc(1u, 256u, (uint32_t)BLOCK_N);
  constexpr uint32_t TMEM_COLS = 2 * BLOCK_N;
  constexpr int NSLAB = BLOCK_N / 32;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = smem;
  uint8_t* sB = smem + (size_t)p.sa_stages * p.a_buf_bytes;
  uint8_t* sStage = sB + (size_t)p.sb_stages * B_BYTES;        // one 4 KB transpose buffer per epilogue warp
  uint64_t* bars = (uint64_t*)(sStage + 8 * 4096);
  uint64_t* a_full = bars;
  uint64_t* a_empty = bars + kMaxA;
  uint64_t* b_full = bars + 2 * kMaxA;
  uint64_t* b_empty = bars + 2 * kMaxA + kMaxB;
  uint64_t* tfull = bars + 2 * kMaxA + 2 * kMaxB;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_slot = (uint32_t*)(tempty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  pdl_trigger();
  const int rank = (int)cluster_ctarank();
  const int img_pairs = (p.N + 1) / 2;
  // FLAT: tiles_per_img holds the number of 128-row tiles of the whole padded batch; work items are pairs of consecutive tiles
  const int total = FLAT ? ((p.tiles_per_img + 1) / 2) * p.n_tiles : img_pairs * p.tiles_per_img * p.n_tiles;
  const int t_first = (int)(blockIdx.x >> 1), t_step = (int)(gridDim.x >> 1);

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    if (p.chunks2 > 0) {
      tma_prefetch_desc(&tmA2);
      tma_prefetch_desc(&tmB2);
    }
    for (int s = 0; s < p.sa_stages; ++s) {
      mbar_init(&a_full[s], 1);
      mbar_init(&a_empty[s], 1);
    }
    for (int s = 0; s < p.sb_stages; ++s) {
      mbar_init(&b_full[s], 1);
      mbar_init(&b_empty[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&tfull[b], 1);
      mbar_init(&tempty[b], 64 * 8);      // the leader's barrier collects both CTAs' 8 epilogue warps
    }
    fence_mbar_init();
  }
  if (warp == 0) {
    tmem_alloc_2sm(tmem_slot, TMEM_COLS);
    tmem_relinquish_2sm();
  }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();

  // work item t -> (n-tile, tile index inside the image, image pair); this CTA's image is 2 * pair + rank
  auto decode = [&](int t, int& nt, int& k, int& n) {
    nt = t % p.n_tiles;
    const int r = t / p.n_tiles;
    if (FLAT) {
      k = 2 * r + rank;      // a tile past the end loads only out-of-bounds rows (zero fill) and stores nothing
      n = 0;
    } else {
      k = r % p.tiles_per_img;
      n = 2 * (r / p.tiles_per_img) + rank;
    }
  };

  if (warp == 0) {
    // ================= TMA producer: one halo box per chunk (A ring), nine weight tiles per chunk (B ring)
    int sa = 0, sb = 0;
    uint32_t pa = 0, pb = 0;
    for (int t = t_first; t < total; t += t_step) {
      int nt, k, n;
      decode(t, nt, k, n);
      const int y_first = (k * 128) / p.Wp;
      const int ncol0 = nt * BLOCK_N + rank * (BLOCK_N / 2);
      for (int seg = 0; seg < 2; ++seg) {
        const int chunks = seg == 0 ? p.chunks1 : p.chunks2;
        const int taps = seg == 0 ? 9 : 1;
        const CUtensorMap* ta = seg == 0 ? &tmA : &tmA2;
        const CUtensorMap* tb = seg == 0 ? &tmB : &tmB2;
        for (int c = 0; c < chunks; ++c) {
          mbar_wait(&a_empty[sa], pa ^ 1u);
          if (elect_one()) {
            if (rank == 0) mbar_arrive_expect_tx(&a_full[sa], 2u * p.box_bytes);
            if (FLAT) tma_load_2d_2sm(sA + (size_t)sa * p.a_buf_bytes, ta, &a_full[sa], c * 64, k * 128 - p.Wp - 1);
            else tma_load_4d_2sm(sA + (size_t)sa * p.a_buf_bytes, ta, &a_full[sa], c * 64, -1, y_first - 1, n);
          }
          __syncwarp();
          if (++sa == p.sa_stages) {
            sa = 0;
            pa ^= 1u;
          }
          for (int tap = 0; tap < taps; ++tap) {
            mbar_wait(&b_empty[sb], pb ^ 1u);
            if (elect_one()) {
              if (rank == 0) mbar_arrive_expect_tx(&b_full[sb], 2u * (uint32_t)B_BYTES);
              tma_load_3d_2sm(sB + (size_t)sb * B_BYTES, tb, &b_full[sb], c * 64, ncol0, tap);
            }
            __syncwarp();
            if (++sb == p.sb_stages) {
              sb = 0;
              pb ^= 1u;
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    if (rank == 0) {
      // ================= MMA issuer (the pair's leader issues for both CTAs)
      int sa = 0, sb = 0;
      uint32_t pa = 0, pb = 0;
      int j = 0;
      for (int t = t_first; t < total; t += t_step, ++j) {
        int nt, k, n;
        decode(t, nt, k, n);
        const int buf = j & 1;
        mbar_wait(&tempty[buf], (((uint32_t)j >> 1) & 1u) ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)buf * BLOCK_N;
        const int s0 = FLAT ? 0 : k * 128 - ((k * 128) / p.Wp) * p.Wp;     // first padded pixel of the tile inside its first image row
        int it = 0;
        for (int seg = 0; seg < 2; ++seg) {
          const int chunks = seg == 0 ? p.chunks1 : p.chunks2;
          const int taps = seg == 0 ? 9 : 1;
          for (int c = 0; c < chunks; ++c) {
            mbar_wait(&a_full[sa], pa);
            tc_fence_after();
            const uint32_t a_base = smem_u32(sA + (size_t)sa * p.a_buf_bytes) + (uint32_t)s0 * 128u;
            int ty = 0, tx = 0;
            for (int tap = 0; tap < taps; ++tap, ++it) {
              mbar_wait(&b_full[sb], pb);
              tc_fence_after();
              if (elect_one()) {
                const uint32_t off = seg == 0 ? (uint32_t)(ty * p.Wp + tx) : (uint32_t)(p.Wp + 1);
                const uint64_t adesc = umma_desc_sw128(a_base + off * 128u);
                const uint64_t bdesc = umma_desc_sw128(smem_u32(sB + (size_t)sb * B_BYTES));
#pragma unroll
                for (int kk = 0; kk < 4; ++kk)
                  umma_f16_2sm(d_tmem, adesc + (uint64_t)(2 * kk), bdesc + (uint64_t)(2 * kk), IDESC, (uint32_t)((it | kk) != 0));
                umma_commit_2sm(&b_empty[sb]);                   // frees this weight stage in both CTAs
                if (tap + 1 == taps) umma_commit_2sm(&a_empty[sa]);   // ... and the halo box once its last tap has been issued
              }
              __syncwarp();
              if (++tx == 3) {
                tx = 0;
                ++ty;
              }
              if (++sb == p.sb_stages) {
                sb = 0;
                pb ^= 1u;
              }
            }
            if (++sa == p.sa_stages) {
              sa = 0;
              pa ^= 1u;
            }
          }
        }
        if (elect_one()) umma_commit_2sm(&tfull[buf]);
        __syncwarp();
      }
    }
  } else {
    // ================= epilogue warps: warp w drains TMEM lanes 32 (w & 3) .., the two warps sharing a lane quarter alternate slabs.
    // tcgen05.ld (row = lane) -> per-warp 4 KB XOR-swizzled smem transpose -> "transposed domain": 8 lanes cover one row's 32
    // channels, so every global access (residual, output) is a run of full 128-byte (fp32) / 64-byte (bf16) lines.
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;
    const bool has_gn = p.gn_partial != nullptr;
    float4* stg = reinterpret_cast<float4*>(sStage + (size_t)(warp - 2) * 4096);
    const int chunk = lane & 7, rsub = lane >> 3;
    int j = 0;
    for (int t = t_first; t < total; t += t_step, ++j) {
      int nt, k, n;
      decode(t, nt, k, n);
      const int buf = j & 1;
      // this lane's own row -> pixel index inside the image (or -1: junk column / past the image / image past the batch); the rows
      // a thread handles in the transposed domain (it * 4 + rsub) get theirs by shuffle: one division per thread per tile
      const int pp = k * 128 + q * 32 + lane;
      int mine, my_n;
      if (FLAT) {
        // padded pixel of the whole batch -> (image, row, column); row 0 of an image's block and columns 0 / W + 1 are borders
        const int R = pp / p.Wp, col = pp - R * p.Wp;
        my_n = R / (p.H + 1);
        const int yr = R - my_n * (p.H + 1);
        mine = (yr >= 1 && col >= 1 && col <= p.W && my_n < p.N) ? (my_n * p.H + (yr - 1)) * p.W + (col - 1) : -1;
      } else {
        const int y = pp / p.Wp, x = pp - y * p.Wp;
        my_n = n;
        mine = (x < p.W && y < p.H && n < p.N) ? y * p.W + x : -1;
      }
      int poff[8], nrow[8];
      unsigned okmask = 0;
#pragma unroll
      for (int it = 0; it < 8; ++it) {
        poff[it] = __shfl_sync(0xffffffffu, mine, it * 4 + rsub);
        nrow[it] = FLAT ? __shfl_sync(0xffffffffu, my_n, it * 4 + rsub) : n;
        okmask |= (poff[it] >= 0 ? 1u : 0u) << it;
        if (poff[it] < 0) {
          poff[it] = 0;
          nrow[it] = 0;
        }
      }
      const int nA = FLAT ? __shfl_sync(0xffffffffu, my_n, 0) : n;        // image of the warp's first row (statistics: nA and nA + 1)
      const long long img = FLAT ? 0ll : (long long)(n < p.N ? n : 0) * p.H * p.W;
      mbar_wait(&tfull[buf], ((uint32_t)j >> 1) & 1u);
      tc_fence_after();
      const uint32_t t_addr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)buf * BLOCK_N;
#pragma unroll 1
      for (int sl = half; sl < NSLAB; sl += 2) {
        const int c0 = nt * BLOCK_N + sl * 32;
        const int c = c0 + chunk * 4;
        float4 rsd[8];
        if (EK == 2 && p.residual && !(p.dbg & 1)) {
#pragma unroll
          for (int it = 0; it < 8; ++it) rsd[it] = *reinterpret_cast<const float4*>(p.residual + (img + poff[it]) * p.res_ld + c);
        }
        uint32_t v[32];
        tmem_ld_32x32(t_addr + (uint32_t)(sl * 32), v);
        tmem_ld_wait();
        if (sl + 2 >= NSLAB) {
          tc_fence_before();
          mbar_arrive_leader(&tempty[buf]);
        }
        if (p.dbg & 1) continue;
#pragma unroll
        for (int ch = 0; ch < 8; ++ch)
          stg[lane * 8 + (ch ^ (lane & 7))] = make_float4(__uint_as_float(v[4 * ch]), __uint_as_float(v[4 * ch + 1]),
                                                          __uint_as_float(v[4 * ch + 2]), __uint_as_float(v[4 * ch + 3]));
        __syncwarp();
        float4 bia = make_float4(0.f, 0.f, 0.f, 0.f);
        if (p.bias) bia = __ldg(reinterpret_cast<const float4*>(p.bias + c));
        if (EK == 1 && p.rowbias && !FLAT) {
          const float4 rb = __ldg(reinterpret_cast<const float4*>(p.rowbias + (long long)(n < p.N ? n : 0) * p.rowbias_ld + c));
          bia.x += rb.x; bia.y += rb.y; bia.z += rb.z; bia.w += rb.w;
        }
        float gs = 0.f, gq = 0.f, gsB = 0.f, gqB = 0.f, gsC = 0.f, gqC = 0.f;     // FLAT: rows of image nA / nA + 1 / nA + 2
#pragma unroll
        for (int it = 0; it < 8; ++it) {
          const int r = it * 4 + rsub;
          float4 xv = stg[r * 8 + (chunk ^ (r & 7))];
          float4 bi = bia;
          if (FLAT && EK == 1 && p.rowbias) {             // a warp's rows may belong to two images: the per-image bias goes per row
            const float4 rb = __ldg(reinterpret_cast<const float4*>(p.rowbias + (long long)nrow[it] * p.rowbias_ld + c));
            bi.x += rb.x; bi.y += rb.y; bi.z += rb.z; bi.w += rb.w;
          }
          xv.x = (xv.x + bi.x) * p.scale; xv.y = (xv.y + bi.y) * p.scale; xv.z = (xv.z + bi.z) * p.scale; xv.w = (xv.w + bi.w) * p.scale;
          if (EK == 2 && p.residual) {
            xv.x += p.res_scale * rsd[it].x; xv.y += p.res_scale * rsd[it].y; xv.z += p.res_scale * rsd[it].z; xv.w += p.res_scale * rsd[it].w;
          }
          if ((okmask >> it) & 1u) {
            // junk rows read past the halo box (their values may be anything, NaN included): they are skipped, never masked by a product
            if (!(p.dbg & 2)) {
              if (EK == 2) *reinterpret_cast<float4*>(p.out_f32 + (img + poff[it]) * p.out_ld + c) = xv;
              else *reinterpret_cast<uint2*>(p.out_bf16 + (img + poff[it]) * p.out_ld + c) = make_uint2(pack_bf16x2(xv.x, xv.y), pack_bf16x2(xv.z, xv.w));
            }
            const float s_ = (xv.x + xv.y) + (xv.z + xv.w);
            const float q_ = (xv.x * xv.x + xv.y * xv.y) + (xv.z * xv.z + xv.w * xv.w);
            if (!FLAT || nrow[it] == nA) {
              gs += s_;
              gq += q_;
            } else if (nrow[it] == nA + 1) {
              gsB += s_;
              gqB += q_;
            } else {
              gsC += s_;
              gqC += q_;
            }
          }
        }
        if (has_gn && !(p.dbg & 4)) {
          // this thread holds 8 rows x 4 channels; lanes with the same chunk (lane ^ 8, ^ 16) hold the other rows
          auto flush = [&](float a, float b, int nn) {
            a += __shfl_xor_sync(0xffffffffu, a, 8);  b += __shfl_xor_sync(0xffffffffu, b, 8);
            a += __shfl_xor_sync(0xffffffffu, a, 16); b += __shfl_xor_sync(0xffffffffu, b, 16);
            {
              float s1 = a, q1 = b;
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
              float s2 = a, q2 = b;
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
          };
          flush(gs, gq, FLAT ? nA : n);
          if (FLAT) flush(gsB, gqB, nA + 1);      // zero when the warp's rows all belong to one image (adds nothing)
          if (FLAT && (p.H + 1) * p.Wp < 32) flush(gsC, gqC, nA + 2);      // 4x4 maps: 30 padded pixels per image, 32 rows can touch three
        }
        __syncwarp();   // the staging buffer is rewritten by the next slab
      }
    }
  }
  tc_fence_before();
  cluster_sync_all();          // neither CTA may retire (or free TMEM) while the peer still reads its smem / signals its barriers
  if (warp == 0) tmem_dealloc_2sm(tmem_base, TMEM_COLS);
}

template <int BLOCK_N, int EK, bool FLAT = false>
int launch_halo(const CUtensorMap& a, const CUtensorMap& b, const CUtensorMap& a2, const CUtensorMap& b2, HaloParams p, cudaStream_t stream) {
  constexpr int B_BYTES = (BLOCK_N / 2) * 128;
  const int overhead = 1024 + 8 * 4096 + (2 * kMaxA + 2 * kMaxB + 4) * 8 + 64;
  p.sa_stages = 3;
  int sb = (int)((220 * 1024 - overhead - (long long)p.sa_stages * p.a_buf_bytes) / B_BYTES);
  if (sb < 4) {
    p.sa_stages = 2;
    sb = (int)((220 * 1024 - overhead - (long long)p.sa_stages * p.a_buf_bytes) / B_BYTES);
  }
  if (sb > kMaxB) sb = kMaxB;
  if (sb < 3) {
    indm_set_error("igemm (halo): not enough shared memory for the weight ring (%d stages)", sb);
    return INDM_ERR_ARG;
  }
  p.sb_stages = sb;
  const size_t smem = (size_t)overhead + (size_t)p.sa_stages * p.a_buf_bytes + (size_t)sb * B_BYTES;
  auto kern = igemm_halo_kernel<BLOCK_N, EK, FLAT>;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) {
      indm_set_error("igemm (halo): cudaFuncSetAttribute failed: %s", cudaGetErrorString(e));
      return INDM_ERR_CUDA;
    }
    configured = true;
  }
  const long long items = FLAT ? (long long)((p.tiles_per_img + 1) / 2) * p.n_tiles : (long long)((p.N + 1) / 2) * p.tiles_per_img * p.n_tiles;
  const int half_sms = indm_num_sms() / 2;
  const int grid = 2 * (int)(items < half_sms ? items : half_sms);
  indm_launch_pdl_cluster2(kern, dim3(grid), dim3(320), smem, stream, a, b, a2, b2, p);
  INDM_CHECK_LAUNCH("igemm (halo)");
  return INDM_OK;
}

}  // namespace

bool indm_halo_eligible(const indm_igemm_t* d, int kind) {
  static const bool enabled = []() { const char* e = getenv("INDM_IGEMM_HALO"); return !(e && e[0] == '0'); }();
  if (!enabled || d->dtype != INDM_DTYPE_BF16 || d->taps != 9 || (d->stride != 0 && d->stride != 1) || d->batched_b) return false;
  if (kind != 1 && kind != 2) return false;
  if (d->W != 32 && d->W != 64) return false;                         // where whole padded tiles use >= 85 % of their rows
  if (d->H * (d->W + 2) < 4 * 128 || d->N < 2) return false;
  if (d->Cin % 64 != 0 || (d->a2 && (d->Cin2 % 64 != 0 || d->a2_ld != 0)) || d->a_ld != 0 || d->a_img_stride != 0) return false;
  if (d->Cout % 128 != 0 || d->block_n == 32 || d->block_n == 64) return false;
  if (d->b_ld != 0 || d->b_tap_stride != 0 || d->b2_ld != 0) return false;
  const long long old_ = d->out_ld ? d->out_ld : d->Cout;
  if (kind == 1 && (((uintptr_t)d->out_bf16 & 7) || old_ % 4)) return false;
  if (kind == 2 && (((uintptr_t)d->out_f32 & 15) || old_ % 4)) return false;
  if (d->residual && ((((uintptr_t)d->residual) & 15) || ((d->res_ld ? d->res_ld : d->Cout) % 4))) return false;
  if (d->rowbias && ((((uintptr_t)d->rowbias) & 15) || (d->rowbias_ld % 4))) return false;
  if (d->bias && (((uintptr_t)d->bias) & 15)) return false;
  // enough work items to fill the chip's CTA pairs
  const int Wp = d->W + 2;
  const long long items = (long long)((d->N + 1) / 2) * ((d->H * Wp + 127) / 128) * (d->Cout / ((d->Cout % 256 == 0) ? 256 : 128));
  return items >= indm_num_sms() / 2;
}

int indm_igemm_halo(const indm_igemm_t* d, int kind, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  HaloParams p{};
  p.N = d->N; p.H = d->H; p.W = d->W; p.Wp = d->W + 2;
  p.tiles_per_img = (d->H * p.Wp + 127) / 128;
  const int block_n = (d->Cout % 256 == 0) ? 256 : 128;
  p.n_tiles = d->Cout / block_n;
  p.Cout = d->Cout;
  p.chunks1 = d->Cin / 64;
  p.chunks2 = d->a2 ? d->Cin2 / 64 : 0;
  // box: rows y_first - 1 .. y_last + 1 of the padded grid; a tile starting s0 <= Wp - 1 pixels into its first row spans
  // (Wp - 1 + 127) / Wp + 1 image rows at most
  const int box_rows = (p.Wp - 1 + 127) / p.Wp + 1 + 2;
  p.box_bytes = (uint32_t)(box_rows * p.Wp) * 128u;
  const uint32_t need_rows = (uint32_t)(3 * p.Wp + 129);      // s0 + 127 + 2 Wp + 2 + 1 rows may be addressed by the last tap
  const uint32_t rows = need_rows > (uint32_t)(box_rows * p.Wp) ? need_rows : (uint32_t)(box_rows * p.Wp);
  p.a_buf_bytes = (rows * 128u + 1023u) & ~1023u;
  p.bias = d->bias; p.rowbias = d->rowbias; p.rowbias_ld = d->rowbias_ld;
  p.residual = d->residual; p.res_ld = d->res_ld ? d->res_ld : d->Cout;
  p.scale = d->scale; p.res_scale = d->res_scale;
  p.out_f32 = d->out_f32; p.out_bf16 = (__nv_bfloat16*)d->out_bf16; p.out_ld = d->out_ld ? d->out_ld : d->Cout;
  p.gn_partial = d->gn_partial; p.gn_cpg = d->gn_cpg; p.gn_groups = d->gn_groups; p.gn_goff = d->gn_goff;
  p.gn2_partial = d->gn2_partial; p.gn2_cpg = d->gn2_cpg; p.gn2_groups = d->gn2_groups; p.gn2_goff = d->gn2_goff;
  { static const int dbg_flags = []() { const char* e = getenv("INDM_IGEMM_DBG"); return e ? atoi(e) : 0; }(); p.dbg = dbg_flags; }
  if (p.gn_partial) {
    INDM_CHECK_ARG(p.gn_cpg >= 4 && 32 % p.gn_cpg == 0, "igemm (halo): fused GroupNorm statistics need cpg | 32 (cpg=%d)", p.gn_cpg);
    INDM_CHECK_ARG(!p.gn2_partial || (p.gn2_cpg >= 4 && 32 % p.gn2_cpg == 0), "igemm (halo): second GroupNorm consumer needs cpg | 32");
  }
  CUtensorMap tmA, tmB, tmA2, tmB2;
  const CUtensorMapDataType dt = CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
  {
    const uint64_t dims[4] = {(uint64_t)d->Cin, (uint64_t)d->W, (uint64_t)d->H, (uint64_t)d->N};
    const uint64_t str[3] = {(uint64_t)d->Cin * 2, (uint64_t)d->W * d->Cin * 2, (uint64_t)d->H * d->W * d->Cin * 2};
    const uint32_t box[4] = {64u, (uint32_t)p.Wp, (uint32_t)box_rows, 1u};
    int rc = indm_make_tmap(&tmA, dt, 4, d->a, dims, str, box, "igemm (halo) A");
    if (rc) return rc;
  }
  {
    const uint64_t dims[3] = {(uint64_t)d->Cin, (uint64_t)d->Cout, 9ull};
    const uint64_t str[2] = {(uint64_t)d->Cin * 2, (uint64_t)d->Cout * d->Cin * 2};
    const uint32_t box[3] = {64u, (uint32_t)(block_n / 2), 1u};
    int rc = indm_make_tmap(&tmB, dt, 3, d->b, dims, str, box, "igemm (halo) B");
    if (rc) return rc;
  }
  if (d->a2) {
    const uint64_t dims[4] = {(uint64_t)d->Cin2, (uint64_t)d->W, (uint64_t)d->H, (uint64_t)d->N};
    const uint64_t str[3] = {(uint64_t)d->Cin2 * 2, (uint64_t)d->W * d->Cin2 * 2, (uint64_t)d->H * d->W * d->Cin2 * 2};
    const uint32_t box[4] = {64u, (uint32_t)p.Wp, (uint32_t)box_rows, 1u};
    int rc = indm_make_tmap(&tmA2, dt, 4, d->a2, dims, str, box, "igemm (halo) A2");
    if (rc) return rc;
    const uint64_t bdims[3] = {(uint64_t)d->Cin2, (uint64_t)d->Cout, 1ull};
    const uint64_t bstr[2] = {(uint64_t)d->Cin2 * 2, (uint64_t)d->Cout * d->Cin2 * 2};
    const uint32_t bbox[3] = {64u, (uint32_t)(block_n / 2), 1u};
    rc = indm_make_tmap(&tmB2, dt, 3, d->b2, bdims, bstr, bbox, "igemm (halo) B2");
    if (rc) return rc;
  } else {
    tmA2 = tmA;
    tmB2 = tmB;
  }
  if (block_n == 256) return kind == 1 ? launch_halo<256, 1>(tmA, tmB, tmA2, tmB2, p, stream) : launch_halo<256, 2>(tmA, tmB, tmA2, tmB2, p, stream);
  return kind == 1 ? launch_halo<128, 1>(tmA, tmB, tmA2, tmB2, p, stream) : launch_halo<128, 2>(tmA, tmB, tmA2, tmB2, p, stream);
}

// ---------------------------------------------------------------- padded-pixel ("PP") operand layout, whole batch as one pixel sequence
int indm_igemm_halo_flat(const indm_igemm_t* d, int kind, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  INDM_CHECK_ARG(d->dtype == INDM_DTYPE_BF16 && d->taps == 9 && (d->stride == 0 || d->stride == 1) && !d->batched_b,
                 "igemm (a_pp): the padded-pixel operand layout is for BF16 3x3 stride-1 convolutions");
  INDM_CHECK_ARG(kind == 1 || kind == 2, "igemm (a_pp): only the plain epilogues (bf16 out [+ bias + row bias] / fp32 out [+ bias + residual])");
  INDM_CHECK_ARG(d->Cin % 64 == 0 && (!d->a2 || d->Cin2 % 64 == 0) && d->Cout % 64 == 0, "igemm (a_pp): Cin %% 64, Cout %% 64");
  INDM_CHECK_ARG(d->a_ld == 0 && d->a_img_stride == 0 && d->a2_ld == 0 && d->b_ld == 0 && d->b_tap_stride == 0 && d->b2_ld == 0,
                 "igemm (a_pp): dense operands only");
  HaloParams p{};
  p.N = d->N; p.H = d->H; p.W = d->W; p.Wp = d->W + 2;
  INDM_CHECK_ARG(128 + 2 * p.Wp + 2 <= 256, "igemm (a_pp): W too large for one box (W <= 61)");
  const long long T = ((long long)d->N * (d->H + 1) + 1) * p.Wp;          // rows of the padded buffer
  p.tiles_per_img = (int)((T + 127) / 128);
  // tile width: the widest whose pair count still fills most of the chip (A costs one box per pixel tile here, so narrow N tiles
  // re-read little; B is what a CTA mostly ingests, in proportion to the tile width)
  static const int forced_bn = []() { const char* e = getenv("INDM_PP_BLOCKN"); return e ? atoi(e) : 0; }();
  const long long pairs = (p.tiles_per_img + 1) / 2;
  int block_n = 64;
  if (d->Cout % 256 == 0 && pairs * (d->Cout / 256) >= indm_num_sms() / 4) block_n = 256;
  else if (pairs * (d->Cout / 128) >= indm_num_sms() / 4) block_n = 128;
  if (forced_bn && d->Cout % forced_bn == 0) block_n = forced_bn;
  p.n_tiles = d->Cout / block_n;
  p.Cout = d->Cout;
  p.chunks1 = d->Cin / 64;
  p.chunks2 = d->a2 ? d->Cin2 / 64 : 0;
  const int box_rows = 128 + 2 * p.Wp + 2;
  p.box_bytes = (uint32_t)box_rows * 128u;
  p.a_buf_bytes = (p.box_bytes + 1023u) & ~1023u;
  p.bias = d->bias; p.rowbias = d->rowbias; p.rowbias_ld = d->rowbias_ld;
  p.residual = d->residual; p.res_ld = d->res_ld ? d->res_ld : d->Cout;
  p.scale = d->scale; p.res_scale = d->res_scale;
  p.out_f32 = d->out_f32; p.out_bf16 = (__nv_bfloat16*)d->out_bf16; p.out_ld = d->out_ld ? d->out_ld : d->Cout;
  p.gn_partial = d->gn_partial; p.gn_cpg = d->gn_cpg; p.gn_groups = d->gn_groups; p.gn_goff = d->gn_goff;
  p.gn2_partial = d->gn2_partial; p.gn2_cpg = d->gn2_cpg; p.gn2_groups = d->gn2_groups; p.gn2_goff = d->gn2_goff;
  { static const int dbg_flags = []() { const char* e = getenv("INDM_IGEMM_DBG"); return e ? atoi(e) : 0; }(); p.dbg = dbg_flags; }
  if (p.gn_partial) {
    INDM_CHECK_ARG(p.gn_cpg >= 4 && 32 % p.gn_cpg == 0 && (!p.gn2_partial || (p.gn2_cpg >= 4 && 32 % p.gn2_cpg == 0)),
                   "igemm (a_pp): fused GroupNorm statistics need cpg | 32");
    INDM_CHECK_ARG((d->H + 1) * p.Wp >= 16, "igemm (a_pp): fused GroupNorm statistics need >= 16 padded pixels per image (a warp's 32 rows may span three images, not more)");
  }
  INDM_CHECK_ARG((((uintptr_t)d->out_bf16 | (uintptr_t)d->out_f32 | (uintptr_t)d->residual | (uintptr_t)d->rowbias | (uintptr_t)d->bias) & 15) == 0 &&
                 p.out_ld % 4 == 0 && p.res_ld % 4 == 0 && p.rowbias_ld % 4 == 0, "igemm (a_pp): 16-byte aligned epilogue operands");
  CUtensorMap tmA, tmB, tmA2, tmB2;
  const CUtensorMapDataType dt = CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
  {
    const uint64_t dims[2] = {(uint64_t)d->Cin, (uint64_t)T};
    const uint64_t str[1] = {(uint64_t)d->Cin * 2};
    const uint32_t box[2] = {64u, (uint32_t)box_rows};
    int rc = indm_make_tmap(&tmA, dt, 2, d->a, dims, str, box, "igemm (a_pp) A");
    if (rc) return rc;
  }
  {
    const uint64_t dims[3] = {(uint64_t)d->Cin, (uint64_t)d->Cout, 9ull};
    const uint64_t str[2] = {(uint64_t)d->Cin * 2, (uint64_t)d->Cout * d->Cin * 2};
    const uint32_t box[3] = {64u, (uint32_t)(block_n / 2), 1u};
    int rc = indm_make_tmap(&tmB, dt, 3, d->b, dims, str, box, "igemm (a_pp) B");
    if (rc) return rc;
  }
  if (d->a2) {
    const uint64_t dims[2] = {(uint64_t)d->Cin2, (uint64_t)T};
    const uint64_t str[1] = {(uint64_t)d->Cin2 * 2};
    const uint32_t box[2] = {64u, (uint32_t)box_rows};
    int rc = indm_make_tmap(&tmA2, dt, 2, d->a2, dims, str, box, "igemm (a_pp) A2");
    if (rc) return rc;
    const uint64_t bdims[3] = {(uint64_t)d->Cin2, (uint64_t)d->Cout, 1ull};
    const uint64_t bstr[2] = {(uint64_t)d->Cin2 * 2, (uint64_t)d->Cout * d->Cin2 * 2};
    const uint32_t bbox[3] = {64u, (uint32_t)(block_n / 2), 1u};
    rc = indm_make_tmap(&tmB2, dt, 3, d->b2, bdims, bstr, bbox, "igemm (a_pp) B2");
    if (rc) return rc;
  } else {
    tmA2 = tmA;
    tmB2 = tmB;
  }
  if (block_n == 256)
    return kind == 1 ? launch_halo<256, 1, true>(tmA, tmB, tmA2, tmB2, p, stream) : launch_halo<256, 2, true>(tmA, tmB, tmA2, tmB2, p, stream);
  if (block_n == 64)
    return kind == 1 ? launch_halo<64, 1, true>(tmA, tmB, tmA2, tmB2, p, stream) : launch_halo<64, 2, true>(tmA, tmB, tmA2, tmB2, p, stream);
  return kind == 1 ? launch_halo<128, 1, true>(tmA, tmB, tmA2, tmB2, p, stream) : launch_halo<128, 2, true>(tmA, tmB, tmA2, tmB2, p, stream);
}
