// Masked multi-head attention backward for long sequences, tcgen05 / TMEM (autograd of F.scaled_dot_product_attention,
// reference src/multi_modal/mm_utils.py:105-112, :143-150).  Two kernels, one per orientation, so there are no atomics
// and no transposed fragments; both stream 128-wide blocks of the other sequence through a two-stage TMA ring and keep
// their output accumulators in tensor memory across the whole loop:
//   dq  : CTA = (b, h, 128 queries).  Per key block j: S = Q K_j^T and dP = dO V_j^T land side by side in TMEM; 512
//         threads (row = query = TMEM lane, 4 column groups of 32 keys) turn them into dS = P * (keep * dP - delta),
//         written back IN PLACE as bf16, which the next tcgen05.mma reads as its A operand: dQ += dS K_j (K_j re-used
//         as MN-major B operand).  The scores of block j+1 are issued right behind that product.
//   dkv : CTA = (b, h, 128 keys).  Per query block i: S^T = K Q_i^T, dP^T = V dO_i^T; threads (row = key) produce
//         dS^T and P_drop^T in place; dK += dS^T Q_i, dV += P_drop^T dO_i re-use the Q_i / dO_i tiles as MN-major B
//         operands; lse / delta / keep bits of the next query block are staged in shared memory one block ahead.
// lse / delta / keep bits come from the forward and the prep kernel exactly as in attention.cu's kernels.
#include "attn_common.cuh"

namespace mmfm {

constexpr int kStreamThreads = 512;
constexpr int kStreamColWords = 512;   // key-validity bits: Sk <= 16384

template <int D>
struct StreamCfg {
  static constexpr uint32_t kRowBytes = D * 2;
  static constexpr uint32_t kLayout = (D == 32) ? 4u : 2u;
  static constexpr uint32_t kSbo = 8 * kRowBytes;
  static constexpr uint32_t kTile = 128 * kRowBytes;            // one 128-row operand tile
  static constexpr uint32_t kSmem = 1024 + 6 * kTile;           // 2 resident tiles + 2 stages x 2 streamed tiles
};

// 16 keep bits of the forward's layout -> per-column tests: column jj (0..63) of a 64-key block lives in the 16-bit word
// of quad lane ql = (jj%8)/2 at bit 2*(jj/8) + jj%2
MMFM_DEVINL void keep_words(const uint2 w2, int c, uint32_t (&kw)[4]) {
  const int sh = 8 * (c & 1);   // second 32-column chunk of the block: n-tiles 4..7 -> bits 8..15
  kw[0] = (w2.x & 0xFFFFu) >> sh;
  kw[1] = (w2.x >> 16) >> sh;
  kw[2] = (w2.y & 0xFFFFu) >> sh;
  kw[3] = (w2.y >> 16) >> sh;
}

template <int D, bool DROP>
__global__ void __launch_bounds__(kStreamThreads, 1) attn_bwd_dq_stream_kernel(
    const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmdO,
    const __grid_constant__ CUtensorMap tmK, const __grid_constant__ CUtensorMap tmV, const AttnParams p, int nb_all) {
  using Cfg = StreamCfg<D>;
  constexpr uint32_t kRowBytes = Cfg::kRowBytes, kLayout = Cfg::kLayout, kSbo = Cfg::kSbo;
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t ld_q, kv_full[2], m1_bar, m2_bar;
  __shared__ uint32_t tmem_slot;
  __shared__ uint32_t s_colbits[kStreamColWords];

  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sQ = smem_base, sdO = sQ + Cfg::kTile, sKV = sdO + Cfg::kTile;   // stage s: K at sKV + 2s tiles, V next
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int quad = warp & 3, grp = warp >> 2;
  const int q0 = blockIdx.x * 128, h = blockIdx.y, b = blockIdx.z;
  const int mode = p.mask_mode;
  const long long bh = (long long)(b * p.nh + h);
  // mask-aware block skipping: under the causal mask the keys after this query tile's last row are never attended, so the
  // key blocks past the diagonal are neither loaded nor multiplied (half of all blocks for a long sequence)
  const int nb = (mode == MMFM_MASK_CAUSAL) ? min(nb_all, q0 / 128 + 1) : nb_all;
  const int bl = (nb == nb_all) ? (((p.Sk - (nb - 1) * 128) + 15) & ~15) : 128;   // width of the last key block

  if (tid == 0) {
    tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmdO); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmV);
    mbar_init(&ld_q, 1);
    mbar_init(&kv_full[0], 1);
    mbar_init(&kv_full[1], 1);
    mbar_init(&m1_bar, 1);
    mbar_init(&m2_bar, 1);
    fence_mbar_init();
    mbar_arrive_expect_tx(&ld_q, 2 * Cfg::kTile);
    tma_load_2d_addr(sQ, &tmQ, &ld_q, h * D, b * p.Sq + q0);
    tma_load_2d_addr(sdO, &tmdO, &ld_q, h * D, b * p.Sq + q0);
    for (int s = 0; s < 2 && s < nb; ++s) {
      mbar_arrive_expect_tx(&kv_full[s], 2 * Cfg::kTile);
      tma_load_2d_addr(sKV + 2 * s * Cfg::kTile, &tmK, &kv_full[s], h * D, b * p.Sk + s * 128);
      tma_load_2d_addr(sKV + (2 * s + 1) * Cfg::kTile, &tmV, &kv_full[s], h * D, b * p.Sk + s * 128);
    }
  }
  if (warp == 1) {
    tmem_alloc(&tmem_slot, 512u);
    tmem_relinquish();
  }
  {
    const unsigned char* kvg = p.key_valid + (long long)b * p.Sk;
    const int ncw = min((p.Sk + 31) >> 5, kStreamColWords);
    for (int w = warp; w < ncw; w += kStreamThreads / 32) {
      const int j = w * 32 + lane;
      const bool v = (j < p.Sk) && (mode == MMFM_MASK_CAUSAL || kvg[j] != 0);
      const uint32_t m = __ballot_sync(0xffffffffu, v);
      if (lane == 0) s_colbits[w] = m;
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;
  constexpr uint32_t dp_col = 128u, acc_col = 256u;

  auto issue_scores = [&](int j) {   // S = Q K_j^T, dP = dO V_j^T  (one elected thread)
    const int s = j & 1;
    mbar_wait(&kv_full[s], (uint32_t)((j >> 1) & 1));
    tc_fence_after();
    const uint32_t n = (uint32_t)(j == nb - 1 ? bl : 128);
    const uint32_t idesc = make_idesc_bf16(128, n, 0, 0);
    const uint32_t ak = sKV + 2 * s * Cfg::kTile, av = ak + Cfg::kTile;
#pragma unroll
    for (int k = 0; k < D / 16; ++k)
      umma_bf16(tmem_base, make_smem_desc(sQ + k * 32, 16, kSbo, kLayout), make_smem_desc(ak + k * 32, 16, kSbo, kLayout),
                idesc, k > 0 ? 1u : 0u);
#pragma unroll
    for (int k = 0; k < D / 16; ++k)
      umma_bf16(tmem_base + dp_col, make_smem_desc(sdO + k * 32, 16, kSbo, kLayout),
                make_smem_desc(av + k * 32, 16, kSbo, kLayout), idesc, k > 0 ? 1u : 0u);
    umma_commit(&m1_bar);
  };

  if (warp == 0) {
    if (elect_one()) {
      mbar_wait(&ld_q, 0);
      issue_scores(0);
    }
    __syncwarp();
  }

  const int row = quad * 32 + lane;
  const int i = q0 + row;
  const float sl2 = p.scale * kLog2e;
  const float dsc = DROP ? p.drop_p.scale : 1.0f;
  const float lse2 = (i < p.Sq) ? p.lse[bh * p.Sq + i] * kLog2e : INFINITY;
  const float dl = ((i < p.Sq) ? p.delta[bh * p.Sq + i] : 0.f) / dsc;
  const uint32_t t_row = tmem_base + ((uint32_t)(quad * 32) << 16);
  const int nkb = (p.Sk + kTile - 1) / kTile;
  const int ncw = min((p.Sk + 31) >> 5, kStreamColWords);
  const uint32_t idesc_dq = make_idesc_bf16(128, D, 0, 1);

#pragma unroll 1
  for (int j = 0; j < nb; ++j) {
    const int width = (j == nb - 1) ? bl : 128;
    const int c = grp;                       // this thread's 32-column chunk of the block
    const int cg = 4 * j + c;                // global chunk index
    const bool mine = 32 * c < width;
    uint2 kpre = make_uint2(0xFFFFFFFFu, 0xFFFFFFFFu);
    if (DROP && mine && i < p.Sq) kpre = *reinterpret_cast<const uint2*>(p.p_keep + ((bh * p.Sq + i) * nkb + 2 * j + (c >> 1)) * 4);
    mbar_wait(&m1_bar, (uint32_t)(j & 1));
    tc_fence_after();
    // block j-1's dQ product finished before these scores: its K/V stage is free -> fetch block j+1 into it
    if (tid == 0 && j >= 1 && j + 1 < nb) {
      const int s = (j + 1) & 1;
      mbar_arrive_expect_tx(&kv_full[s], 2 * Cfg::kTile);
      tma_load_2d_addr(sKV + 2 * s * Cfg::kTile, &tmK, &kv_full[s], h * D, b * p.Sk + (j + 1) * 128);
      tma_load_2d_addr(sKV + (2 * s + 1) * Cfg::kTile, &tmV, &kv_full[s], h * D, b * p.Sk + (j + 1) * 128);
    }
    if (mine) {
      uint32_t aw = cg < ncw ? s_colbits[cg] : 0u;
      const int rel = i - 32 * cg;
      if (mode == MMFM_MASK_KEY_OR_DIAG) {
        if (rel >= 0 && rel < 32 && i < p.Sk) aw |= 1u << rel;
      } else if (mode == MMFM_MASK_CAUSAL) {
        aw &= (rel >= 31) ? 0xFFFFFFFFu : (rel < 0 ? 0u : ((2u << rel) - 1u));
      }
      uint32_t kw[4] = {0xFFFFu, 0xFFFFu, 0xFFFFu, 0xFFFFu};
      if (DROP) keep_words(kpre, c, kw);
      uint32_t outp[16];
      uint32_t rs[2][16], rd[2][16];
      tmem_ld16(t_row + 32u * c, rs[0]);
      tmem_ld16(t_row + dp_col + 32u * c, rd[0]);
      tmem_ld16(t_row + 32u * c + 16u, rs[1]);
      tmem_ld16(t_row + dp_col + 32u * c + 16u, rd[1]);
      tmem_ld_wait();
#pragma unroll
      for (int hf = 0; hf < 2; ++hf) {
        float ds[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) {
          const int kk = 16 * hf + k;                 // column inside the chunk
          const bool ok = (aw >> kk) & 1u;
          const float pe = fast_exp2(fmaf(__uint_as_float(rs[hf][k]), sl2, -lse2));
          float dpe = __uint_as_float(rd[hf][k]);
          if (DROP) {
            // column jj = 32*(c&1) + kk of the 64-block: n = jj/8, ql = (jj%8)/2, e = jj%2 -> bit 2n+e (kw pre-shifted)
            if (!((kw[(kk & 7) >> 1] >> (2 * (kk >> 3) + (kk & 1))) & 1u)) dpe = 0.f;
          }
          ds[k] = ok ? pe * (dpe - dl) : 0.f;   // masked columns may hold uninitialised TMEM bits: never multiply them
        }
#pragma unroll
        for (int t = 0; t < 8; ++t) outp[8 * hf + t] = pack_bf16x2(ds[2 * t], ds[2 * t + 1]);
      }
      tmem_st16(t_row + 32u * c, outp);   // in place: bf16 chunk c over the first half of fp32 chunk c
      tmem_st_wait();
    }
    tc_fence_before();
    __syncthreads();

    if (warp == 0) {
      if (elect_one()) {
        tc_fence_after();
        const uint32_t ak = sKV + 2 * (j & 1) * Cfg::kTile;
        const int nks = width >> 4;
        for (int kk = 0; kk < nks; ++kk)
          umma_bf16_ts(tmem_base + acc_col, tmem_base + 32u * (kk >> 1) + 8u * (kk & 1),
                       make_smem_desc(ak + (uint32_t)kk * 16u * kRowBytes, kSbo, kSbo, kLayout), idesc_dq,
                       (j > 0 || kk > 0) ? 1u : 0u);
        if (j + 1 < nb) issue_scores(j + 1);
        else umma_commit(&m2_bar);
      }
      __syncwarp();
    }
  }
  mbar_wait(&m2_bar, 0);
  tc_fence_after();
  uint32_t r[16];
  if (16 * grp < D) {
    tmem_ld16(t_row + acc_col + 16u * grp, r);
    tmem_ld_wait();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512u);   // columns go back before the global stores
  if (16 * grp < D && i < p.Sq) {
    const float fs = p.scale * dsc;
    bf16* dst = p.dq + ((long long)b * p.Sq + i) * p.lddq + h * D + 16 * grp;
#pragma unroll
    for (int k = 0; k < 16; k += 8)
      *reinterpret_cast<uint4*>(dst + k) =
          make_uint4(pack_bf16x2(__uint_as_float(r[k]) * fs, __uint_as_float(r[k + 1]) * fs),
                     pack_bf16x2(__uint_as_float(r[k + 2]) * fs, __uint_as_float(r[k + 3]) * fs),
                     pack_bf16x2(__uint_as_float(r[k + 4]) * fs, __uint_as_float(r[k + 5]) * fs),
                     pack_bf16x2(__uint_as_float(r[k + 6]) * fs, __uint_as_float(r[k + 7]) * fs));
  }
}

template <int D, bool DROP>
__global__ void __launch_bounds__(kStreamThreads, 1) attn_bwd_dkv_stream_kernel(
    const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmdO,
    const __grid_constant__ CUtensorMap tmK, const __grid_constant__ CUtensorMap tmV, const AttnParams p, int nbq) {
  using Cfg = StreamCfg<D>;
  constexpr uint32_t kRowBytes = Cfg::kRowBytes, kLayout = Cfg::kLayout, kSbo = Cfg::kSbo;
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t ld_kv, q_full[2], m1_bar, m2_bar;
  __shared__ uint32_t tmem_slot;
  __shared__ __align__(16) float s_lse[2][128];
  __shared__ __align__(16) float s_dl[2][128];
  __shared__ __align__(16) unsigned short s_keep[2][128][8];   // [stage][query][2 key blocks of this tile][4 quad lanes]

  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sK = smem_base, sV = sK + Cfg::kTile, sQd = sV + Cfg::kTile;   // stage s: Q at sQd + 2s tiles, dO next
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int quad = warp & 3, grp = warp >> 2;
  const int k0 = blockIdx.x * 128, h = blockIdx.y, b = blockIdx.z;
  const int mode = p.mask_mode;
  const long long bh = (long long)(b * p.nh + h);
  const int nkb = (p.Sk + kTile - 1) / kTile;
  const float dsc = DROP ? p.drop_p.scale : 1.0f;
  const int bl = ((p.Sq - (nbq - 1) * 128) + 15) & ~15;   // width of the last query block
  // mask-aware block skipping: under the causal mask the queries before this key tile never attend it, so the loop starts
  // at the diagonal block (stages / barrier phases count from there)
  const int blk0 = (mode == MMFM_MASK_CAUSAL) ? min(k0 / 128, nbq - 1) : 0;

  if (tid == 0) {
    tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmdO); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmV);
    mbar_init(&ld_kv, 1);
    mbar_init(&q_full[0], 1);
    mbar_init(&q_full[1], 1);
    mbar_init(&m1_bar, 1);
    mbar_init(&m2_bar, 1);
    fence_mbar_init();
    mbar_arrive_expect_tx(&ld_kv, 2 * Cfg::kTile);
    tma_load_2d_addr(sK, &tmK, &ld_kv, h * D, b * p.Sk + k0);
    tma_load_2d_addr(sV, &tmV, &ld_kv, h * D, b * p.Sk + k0);
    for (int s = 0; s < 2 && blk0 + s < nbq; ++s) {
      mbar_arrive_expect_tx(&q_full[s], 2 * Cfg::kTile);
      tma_load_2d_addr(sQd + 2 * s * Cfg::kTile, &tmQ, &q_full[s], h * D, b * p.Sq + (blk0 + s) * 128);
      tma_load_2d_addr(sQd + (2 * s + 1) * Cfg::kTile, &tmdO, &q_full[s], h * D, b * p.Sq + (blk0 + s) * 128);
    }
  }
  if (warp == 1) {
    tmem_alloc(&tmem_slot, 512u);
    tmem_relinquish();
  }
  // per-query side data (lse, delta, keep words of this key tile) of query block `blk` -> stage st
  auto stage_rows = [&](int blk, int st, int qi) {
    const int qg = blk * 128 + qi;
    const bool ok = qg < p.Sq;
    s_lse[st][qi] = ok ? p.lse[bh * p.Sq + qg] * kLog2e : INFINITY;
    s_dl[st][qi] = ok ? p.delta[bh * p.Sq + qg] / dsc : 0.f;
    if (DROP) {
      const int kb0 = k0 / kTile;
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        uint2 w2 = make_uint2(0u, 0u);
        if (ok && kb0 + u < nkb) w2 = *reinterpret_cast<const uint2*>(p.p_keep + ((bh * p.Sq + qg) * nkb + kb0 + u) * 4);
        *reinterpret_cast<uint2*>(&s_keep[st][qi][4 * u]) = w2;
      }
    }
  };
  if (tid < 128) stage_rows(blk0, 0, tid);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;
  constexpr uint32_t dp_col = 128u, acc_col = 256u;

  auto issue_scores = [&](int blk) {   // S^T = K Q_i^T, dP^T = V dO_i^T  (one elected thread)
    const int s = (blk - blk0) & 1;
    mbar_wait(&q_full[s], (uint32_t)(((blk - blk0) >> 1) & 1));
    tc_fence_after();
    const uint32_t n = (uint32_t)(blk == nbq - 1 ? bl : 128);
    const uint32_t idesc = make_idesc_bf16(128, n, 0, 0);
    const uint32_t aq = sQd + 2 * s * Cfg::kTile, ad = aq + Cfg::kTile;
#pragma unroll
    for (int k = 0; k < D / 16; ++k)
      umma_bf16(tmem_base, make_smem_desc(sK + k * 32, 16, kSbo, kLayout), make_smem_desc(aq + k * 32, 16, kSbo, kLayout),
                idesc, k > 0 ? 1u : 0u);
#pragma unroll
    for (int k = 0; k < D / 16; ++k)
      umma_bf16(tmem_base + dp_col, make_smem_desc(sV + k * 32, 16, kSbo, kLayout),
                make_smem_desc(ad + k * 32, 16, kSbo, kLayout), idesc, k > 0 ? 1u : 0u);
    umma_commit(&m1_bar);
  };

  if (warp == 0) {
    if (elect_one()) {
      mbar_wait(&ld_kv, 0);
      issue_scores(blk0);
    }
    __syncwarp();
  }

  const int row = quad * 32 + lane;   // key row of the tile
  const int j = k0 + row;
  const bool rowvalid = (j < p.Sk) && (mode == MMFM_MASK_CAUSAL || p.key_valid[(long long)b * p.Sk + j] != 0);
  const float sl2 = p.scale * kLog2e;
  const uint32_t t_row = tmem_base + ((uint32_t)(quad * 32) << 16);
  // keep-bit address of this key inside a query's 8-word row: word 4*(row/64) + (row%8)/2, bit 2*((row%64)/8) + row%2
  const int kword = 4 * (row >> 6) + ((row & 7) >> 1);
  const int kbit = 2 * ((row & 63) >> 3) + (row & 1);
  const uint32_t idesc_acc = make_idesc_bf16(128, D, 0, 1);

#pragma unroll 1
  for (int blk = blk0; blk < nbq; ++blk) {
    const int width = (blk == nbq - 1) ? bl : 128;
    const int st = (blk - blk0) & 1;
    const int c = grp;                        // this thread's 32-query chunk of the block
    const bool mine = 32 * c < width;
    mbar_wait(&m1_bar, (uint32_t)((blk - blk0) & 1));
    tc_fence_after();
    // block blk-1's dK / dV products finished before these scores: its Q / dO stage is free -> fetch block blk+1
    if (tid == 0 && blk > blk0 && blk + 1 < nbq) {
      const int s = (blk + 1 - blk0) & 1;
      mbar_arrive_expect_tx(&q_full[s], 2 * Cfg::kTile);
      tma_load_2d_addr(sQd + 2 * s * Cfg::kTile, &tmQ, &q_full[s], h * D, b * p.Sq + (blk + 1) * 128);
      tma_load_2d_addr(sQd + (2 * s + 1) * Cfg::kTile, &tmdO, &q_full[s], h * D, b * p.Sq + (blk + 1) * 128);
    }
    if (tid >= 384 && blk + 1 < nbq) stage_rows(blk + 1, st ^ 1, tid - 384);   // read by the next iteration only
    if (mine) {
      // allowed(query i = 128 blk + 32c + k, key j)
      const int qbase = 128 * blk + 32 * c;
      const int ncol = p.Sq - qbase;
      uint32_t aw = ncol >= 32 ? 0xFFFFFFFFu : (ncol <= 0 ? 0u : ((1u << ncol) - 1u));   // queries in range
      const int rel = j - qbase;                                                         // column where i == j
      if (mode == MMFM_MASK_CAUSAL) {
        aw &= (rel <= 0) ? 0xFFFFFFFFu : (rel >= 32 ? 0u : ~((1u << rel) - 1u));          // i >= j
        if (j >= p.Sk) aw = 0u;
      } else {
        const uint32_t inr = aw;
        if (!rowvalid) aw = 0u;
        if (mode == MMFM_MASK_KEY_OR_DIAG && rel >= 0 && rel < 32 && j < p.Sk) aw |= (1u << rel) & inr;
      }
      uint32_t outs[16], outp[16];
      uint32_t rs[2][16], rd[2][16];
      tmem_ld16(t_row + 32u * c, rs[0]);
      tmem_ld16(t_row + dp_col + 32u * c, rd[0]);
      tmem_ld16(t_row + 32u * c + 16u, rs[1]);
      tmem_ld16(t_row + dp_col + 32u * c + 16u, rd[1]);
      tmem_ld_wait();
#pragma unroll
      for (int hf = 0; hf < 2; ++hf) {
        float ds[16], pd[16];
#pragma unroll
        for (int k4 = 0; k4 < 16; k4 += 4) {
          const int qi = 32 * c + 16 * hf + k4;
          const float4 l4 = *reinterpret_cast<const float4*>(&s_lse[st][qi]);
          const float4 d4 = *reinterpret_cast<const float4*>(&s_dl[st][qi]);
          const float lv[4] = {l4.x, l4.y, l4.z, l4.w}, dv[4] = {d4.x, d4.y, d4.z, d4.w};
#pragma unroll
          for (int uu = 0; uu < 4; ++uu) {
            const int k = k4 + uu, kk = 16 * hf + k;
            const bool ok = (aw >> kk) & 1u;
            const float pe = fast_exp2(fmaf(__uint_as_float(rs[hf][k]), sl2, -lv[uu]));
            float dpe = __uint_as_float(rd[hf][k]);
            float pde = pe;
            if (DROP) {
              const uint32_t w = s_keep[st][qi + uu][kword];
              if (!((w >> kbit) & 1u)) { dpe = 0.f; pde = 0.f; }
            }
            ds[k] = ok ? pe * (dpe - dv[uu]) : 0.f;   // masked columns may hold uninitialised TMEM bits
            pd[k] = ok ? pde : 0.f;
          }
        }
#pragma unroll
        for (int t = 0; t < 8; ++t) {
          outs[8 * hf + t] = pack_bf16x2(ds[2 * t], ds[2 * t + 1]);
          outp[8 * hf + t] = pack_bf16x2(pd[2 * t], pd[2 * t + 1]);
        }
      }
      tmem_st16(t_row + 32u * c, outs);
      tmem_st16(t_row + dp_col + 32u * c, outp);
      tmem_st_wait();
    }
    tc_fence_before();
    __syncthreads();

    if (warp == 0) {
      if (elect_one()) {
        tc_fence_after();
        const uint32_t aq = sQd + 2 * st * Cfg::kTile, ad = aq + Cfg::kTile;
        const int nks = width >> 4;
        for (int kk = 0; kk < nks; ++kk) {
          const uint32_t a_off = 32u * (kk >> 1) + 8u * (kk & 1);
          const uint32_t acc = (blk > blk0 || kk > 0) ? 1u : 0u;
          umma_bf16_ts(tmem_base + acc_col, tmem_base + a_off,
                       make_smem_desc(aq + (uint32_t)kk * 16u * kRowBytes, kSbo, kSbo, kLayout), idesc_acc, acc);
          umma_bf16_ts(tmem_base + acc_col + D, tmem_base + dp_col + a_off,
                       make_smem_desc(ad + (uint32_t)kk * 16u * kRowBytes, kSbo, kSbo, kLayout), idesc_acc, acc);
        }
        if (blk + 1 < nbq) issue_scores(blk + 1);
        else umma_commit(&m2_bar);
      }
      __syncwarp();
    }
  }
  mbar_wait(&m2_bar, 0);
  tc_fence_after();
  // 2*D accumulator columns (dK | dV) in 16-column pieces over the 4 thread groups; tensor memory is released before
  // the global stores
  constexpr int kPieces = (2 * D) / 16, kPer = kPieces / 4;
  uint32_t r[kPer][16];
#pragma unroll
  for (int u = 0; u < kPer; ++u) tmem_ld16(t_row + acc_col + 16u * (grp + 4 * u), r[u]);
  tmem_ld_wait();
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512u);
  if (j < p.Sk) {
#pragma unroll
    for (int u = 0; u < kPer; ++u) {
      const int piece = grp + 4 * u;
      const bool is_dv = 16 * piece >= D;
      const int col = 16 * piece - (is_dv ? D : 0);
      const float fs = is_dv ? dsc : p.scale * dsc;
      bf16* dst = (is_dv ? p.dv + ((long long)b * p.Sk + j) * p.lddv : p.dk + ((long long)b * p.Sk + j) * p.lddk) + h * D + col;
#pragma unroll
      for (int k = 0; k < 16; k += 8)
        *reinterpret_cast<uint4*>(dst + k) =
            make_uint4(pack_bf16x2(__uint_as_float(r[u][k]) * fs, __uint_as_float(r[u][k + 1]) * fs),
                       pack_bf16x2(__uint_as_float(r[u][k + 2]) * fs, __uint_as_float(r[u][k + 3]) * fs),
                       pack_bf16x2(__uint_as_float(r[u][k + 4]) * fs, __uint_as_float(r[u][k + 5]) * fs),
                       pack_bf16x2(__uint_as_float(r[u][k + 6]) * fs, __uint_as_float(r[u][k + 7]) * fs));
    }
  }
}

}  // namespace mmfm

using namespace mmfm;

template <int D>
static int launch_bwd_stream_d(const mmfm_attn_args* a, const AttnParams& p, cudaStream_t st) {
  using Cfg = StreamCfg<D>;
  const TmaSwizzle sw = (D == 32) ? TMA_SW_64 : TMA_SW_128;
  const uint64_t width = (uint64_t)a->n_heads * D;
  const bool drop = a->drop_p.thresh != 0u;
  CUtensorMap tq, tdo, tk, tv;   // every operand moves in 128-row boxes
  if (int rc = make_tmap_bf16_2d(&tq, a->q, (uint64_t)a->B * a->Sq, width, (uint64_t)a->ldq, D, 128, sw)) return rc;
  if (int rc = make_tmap_bf16_2d(&tdo, a->d_o, (uint64_t)a->B * a->Sq, width, (uint64_t)a->lddo, D, 128, sw)) return rc;
  if (int rc = make_tmap_bf16_2d(&tk, a->k, (uint64_t)a->B * a->Sk, width, (uint64_t)a->ldk, D, 128, sw)) return rc;
  if (int rc = make_tmap_bf16_2d(&tv, a->v, (uint64_t)a->B * a->Sk, width, (uint64_t)a->ldv, D, 128, sw)) return rc;
  static bool attr_set = false;
  if (!attr_set) {
    MMFM_CHECK_CUDA(cudaFuncSetAttribute(attn_bwd_dq_stream_kernel<D, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::kSmem));
    MMFM_CHECK_CUDA(cudaFuncSetAttribute(attn_bwd_dq_stream_kernel<D, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::kSmem));
    MMFM_CHECK_CUDA(cudaFuncSetAttribute(attn_bwd_dkv_stream_kernel<D, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::kSmem));
    MMFM_CHECK_CUDA(cudaFuncSetAttribute(attn_bwd_dkv_stream_kernel<D, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::kSmem));
    attr_set = true;
  }
  const int nbk = (a->Sk + 127) / 128, nbq = (a->Sq + 127) / 128;
  dim3 gq(nbq, a->n_heads, a->B);
  if (drop) attn_bwd_dq_stream_kernel<D, true><<<gq, kStreamThreads, Cfg::kSmem, st>>>(tq, tdo, tk, tv, p, nbk);
  else attn_bwd_dq_stream_kernel<D, false><<<gq, kStreamThreads, Cfg::kSmem, st>>>(tq, tdo, tk, tv, p, nbk);
  MMFM_CHECK_CUDA(cudaGetLastError());
  dim3 gk(nbk, a->n_heads, a->B);
  if (drop) attn_bwd_dkv_stream_kernel<D, true><<<gk, kStreamThreads, Cfg::kSmem, st>>>(tq, tdo, tk, tv, p, nbq);
  else attn_bwd_dkv_stream_kernel<D, false><<<gk, kStreamThreads, Cfg::kSmem, st>>>(tq, tdo, tk, tv, p, nbq);
  MMFM_CHECK_CUDA(cudaGetLastError());
  return 0;
}

namespace mmfm {
// entry used by attention.cu's dispatcher after the prep kernel (delta, output-dropout mask on dO); the caller has
// validated the arguments (no modality-separation mask, 16-byte aligned operands, Sq, Sk <= 16384)
int launch_attn_bwd_stream(const mmfm_attn_args* a, const AttnParams& p, cudaStream_t st) {
  return a->d_head == 32 ? launch_bwd_stream_d<32>(a, p, st) : launch_bwd_stream_d<64>(a, p, st);
}
}  // namespace mmfm
