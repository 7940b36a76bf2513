// Shared by the attention translation units: kernel parameter block, constants, the probability-dropout stream.
#pragma once
#include "common.cuh"
#include "host_util.h"
#include "../../include/mmfm_b200.h"

namespace mmfm {

constexpr int kAttnThreads = 128;
constexpr int kTile = 64;
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;

MMFM_DEVINL float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

struct AttnParams {
  const bf16* q; long long ldq;
  const bf16* k; long long ldk;
  const bf16* v; long long ldv;
  bf16* o; long long ldo;
  float* lse;
  const unsigned char* key_valid;
  const short* mod_q;
  const short* mod_k;
  int B, nh, Sq, Sk;
  int mask_mode;
  float scale;
  DropCfg drop_p, drop_o;
  unsigned short* p_keep;
  // backward
  bf16* d_o; long long lddo;
  float* delta;
  bf16* dq; long long lddq;
  bf16* dk; long long lddk;
  bf16* dv; long long lddv;
};

// 16 random bytes of the probability-dropout field: row = (b*nh+h)*Sq + i, 64-column block blk, quad lane ql
MMFM_DEVINL uint4 pdrop_bytes(unsigned long long seed, uint32_t site, unsigned long long row, uint32_t nblk,
                              uint32_t blk, uint32_t ql) {
  const unsigned long long g = (row * nblk + blk) * 4ull + ql;
  return philox4x32((uint32_t)g, (uint32_t)(g >> 32), site, 1u, (uint32_t)seed, (uint32_t)(seed >> 32));
}

// attention_pipe.cu: persistent warp-specialised tcgen05 forward (any Sk; no modality-separation mask)
int launch_attn_fwd_pipe(const mmfm_attn_args* a, const AttnParams& p, cudaStream_t st);
// attention_bwd_stream.cu: tcgen05 backward pair for long sequences (runs after the prep kernel)
int launch_attn_bwd_stream(const mmfm_attn_args* a, const AttnParams& p, cudaStream_t st);

}  // namespace mmfm
