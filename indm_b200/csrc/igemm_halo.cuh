// Interface between igemm.cu (dispatch) and igemm_halo.cu (the padded-pixel 3x3 convolution kernel).
#pragma once
#include <cuda_bf16.h>
#include <stdint.h>

#include "../../include/indm_b200.h"

struct HaloParams {
  int N, H, W, Wp;
  int tiles_per_img, n_tiles, Cout;
  int chunks1, chunks2;          // 64-channel chunks of the 3x3 segment / of the optional 1x1 skip segment
  int sa_stages, sb_stages;
  uint32_t a_buf_bytes;          // bytes reserved per halo buffer (multiple of 1024)
  uint32_t box_bytes;            // bytes one halo TMA box delivers
  const float* bias;
  const float* rowbias;
  long long rowbias_ld;
  const float* residual;
  long long res_ld;
  float scale, res_scale;
  float* out_f32;
  __nv_bfloat16* out_bf16;
  long long out_ld;
  float* gn_partial;
  int gn_cpg, gn_groups, gn_goff;
  float* gn2_partial;
  int gn2_cpg, gn2_groups, gn2_goff;
  int dbg;                       // development probes (INDM_IGEMM_DBG): 1 = accumulators drained, nothing computed or stored
};

// kind: the plain epilogue kind igemm.cu derived (1 = bf16 out [+ bias + per-image bias], 2 = fp32 out [+ bias + residual])
bool indm_halo_eligible(const indm_igemm_t* d, int kind);
int indm_igemm_halo(const indm_igemm_t* d, int kind, void* stream);
// a / a2 in the padded-pixel layout (indm_igemm_t.a_pp): the whole batch as one padded-pixel sequence
int indm_igemm_halo_flat(const indm_igemm_t* d, int kind, void* stream);
