// Parameter blocks of the tcgen05 GEMM kernels.
#pragma once
#include "common.cuh"

namespace mmfm {

enum GemmAct : int {
  ACT_NONE = 0,
  ACT_GELU = 1,        // fwd: D2 <- pre-activation u (bf16), D <- gelu_erf(u)
  ACT_SOFTSIGN = 2,    // fwd: D <- softsign(v) * act_scale
  ACT_DGELU = 3,       // bwd: D <- v * gelu'(aux)           (aux = saved pre-activation u)
  ACT_DSOFTSIGN = 4,   // bwd: D <- v * act_scale * (1-|aux/act_scale|)^2  (aux = saved softsign*scale output)
};

// D[M,N] = epilogue(A[M,K] . B[N,K]^T); A and B bf16, K-major (row-major with K contiguous).
struct GemmTnParams {
  int M, N, K;
  void* D;            // bf16 or fp32, row pitch ldd elements
  long long ldd;
  int d_fp32;
  bf16* D2;           // optional second output (ACT_GELU pre-activation), pitch ldd
  const float* bias;  // [N] or null
  const float* res;   // fp32 residual [M, ldr] or null (added after dropout)
  long long ldr;
  const bf16* aux;    // saved tensor for ACT_DGELU / ACT_DSOFTSIGN, pitch ldaux
  long long ldaux;
  int act;
  float act_scale;
  DropCfg drop;       // dropout over the (M, N) field, row = GEMM row
  // token-embedding epilogue (encoder/decoder_embeddings.py:54-59 + mm.py:149,289 fused):
  //   out_row = (r / remap_T) * remap_S + remap_off + r % remap_T ;  zero token if zero_flags[remap_off + r % T]
  //   x = tok + (emb_mod[col] + emb_pos[ts[r] * N + col]) ; emb_out (optional) receives the embedding itself
  int remap_T, remap_S, remap_off;
  const unsigned char* zero_flags;
  const float* emb_mod;
  const float* emb_pos;
  const long long* ts;
  float* emb_out;
};

// dW[NO, KI] += sum_r dY[r, NO]^T X[r, KI]   (both operands MN-major: the reduction runs over rows)
struct GemmWgradParams {
  int R, NO, KI;
  float* dW;          // fp32 [NO, ldw] accumulated with red.global.add
  long long ldw;
  int rows_per_split; // multiple of 64
  float* dbias;       // optional [NO]: column sums of dY (computed with a ones-column MMA); may be null
};

}  // namespace mmfm
