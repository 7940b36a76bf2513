// Masked attention backward for the default model shape (d_head 32, Sq, Sk <= 256): persistent, software-pipelined,
// fused tcgen05 kernel (autograd of F.scaled_dot_product_attention, reference src/multi_modal/mm_utils.py:105-112,
// :143-150).  One CTA per SM loops over (batch, head) items and computes dQ, dK and dV of an item in one go.
//
// Data flow of an item (as in attention.cu's attn_bwd_fused2_tc_kernel): the key range is split into two 128-key
// halves, each with its own S / dP buffer in TMEM and its own barriers; per 128-query tile the passes run in the order
// A(h0) A(h1) B(h0) B(h1) and every MMA batch is issued right after the pass that produces its operands and waited for
// one pass later:
//   pass A(h)  p = exp2(s*scale - lse) under the mask, P_drop -> swizzled smem slab ; then  dP_h = dO V_h^T (over the
//              dead S_h), dV_h += P_drop_h^T dO
//   pass B(h)  dS = p_drop * dP - p * delta -> the SAME slab (P_drop is dead by then) ; then  dQ += dS_h K_h,
//              dK_h += dS_h^T Q, and the S_h of the next query tile -- or of the NEXT ITEM
// What the persistent form adds (per-phase clock64 traces showed 21 % of a CTA's life in its prologue and 16 % in its
// epilogue): TMEM and barriers are set up once; the operands (K, V, Q, dO) and the per-row side data (lse, delta, keep
// words -- bulk-copied to shared memory) of item n+1 are loaded while item n computes; the first scores of item n+1
// are issued before item n's accumulators are read out; outputs leave through a staged, row-contiguous copy.
// TMEM: S/dP 2 x 128 | dQ 2 x 32 | dK 2 x 32 | dV 2 x 32 = 448 columns.
#include "attn_common.cuh"
#include <stdlib.h>

#ifdef MMFM_DBG_TIMING
// globaltimer trace of one item of one CTA (tools/micro/bwd_persist_timing.py)
__device__ long long g_dbg_bp[64];
#define DBG_BP(slot) do { if (blockIdx.x == 7 && it == 3 && threadIdx.x == 32) { long long t_; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_) :: "memory"); g_dbg_bp[slot] = t_; } } while (0)
extern "C" int mmfm_debug_read_bwd_persist(long long* out) {
  return (int)cudaMemcpyFromSymbol(out, g_dbg_bp, sizeof(long long) * 64);
}
#else
#define DBG_BP(slot) do { } while (0)
#endif

namespace mmfm {

constexpr uint32_t kPRow = 64;                       // bytes per operand row (32 bf16), 64-byte swizzle
constexpr uint32_t kPOp = 256 * kPRow;               // one operand buffer: 256 rows
constexpr uint32_t kPStage = 4 * kPOp;               // K, V, Q, dO
constexpr uint32_t kPSide = 1024 + 1024 + 8192;      // lse, delta (256 floats each), keep words (256 rows x 4 x 8 B)
constexpr uint32_t kPersistSmem = 1024 + 2 * kPStage + 4 * kSlabBytes + 2 * kPSide;

MMFM_DEVINL void bulk_load(uint32_t dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

template <bool DROP>
__global__ void __launch_bounds__(kFusedThreads, 1) attn_bwd_persist_kernel(
    const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmdO,
    const __grid_constant__ CUtensorMap tmK, const __grid_constant__ CUtensorMap tmV, const AttnParams p, int npad,
    int n_items) {
  constexpr int D = 32;
  constexpr uint32_t kSbo64 = 512;
  pdl_enter();   // the prologue already issues this CTA's first operand loads: wait before anything else
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t ld_bar[2], s_bar[2], dp_bar[2], done_bar;
  __shared__ uint32_t tmem_slot;
  __shared__ uint32_t s_colbits[8];

  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_al = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t sSlab = smem_base + 2 * kPStage;                 // 4 slabs: P_drop, then dS, then the output staging
  const uint32_t side_off = 2 * kPStage + 4 * kSlabBytes;         // byte offset of the side-data stages
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int quad = warp & 3, grp = warp >> 2;
  const int mode = p.mask_mode;
  const int nqt = (p.Sq + 127) >> 7;
  const int nkh = (npad + 127) >> 7;                 // 128-key halves
  const int wlast = npad - 128 * (nkh - 1);          // width of the last half (multiple of 16)
  const int nch = (npad + 31) >> 5;
  const int nkb = (p.Sk + kTile - 1) / kTile;
  const uint32_t lse_bytes = (uint32_t)p.Sq * 4u, keep_bytes = DROP ? (uint32_t)p.Sq * (uint32_t)nkb * 8u : 0u;
  const uint32_t tx_bytes = (uint32_t)(2 * npad) * kPRow + 2u * kPOp + 2u * lse_bytes + keep_bytes;

  auto issue_loads = [&](int item, int st) {   // one thread: operands + side data of `item` into stage st
    const int b = item / p.nh, h = item - b * p.nh;
    const long long bh = (long long)item;
    const uint32_t base = smem_base + st * kPStage;
    mbar_arrive_expect_tx(&ld_bar[st], tx_bytes);
    tma_load_2d_addr(base, &tmK, &ld_bar[st], h * D, b * p.Sk);
    tma_load_2d_addr(base + kPOp, &tmV, &ld_bar[st], h * D, b * p.Sk);
    tma_load_2d_addr(base + 2 * kPOp, &tmQ, &ld_bar[st], h * D, b * p.Sq);
    tma_load_2d_addr(base + 2 * kPOp + 128 * kPRow, &tmQ, &ld_bar[st], h * D, b * p.Sq + 128);
    tma_load_2d_addr(base + 3 * kPOp, &tmdO, &ld_bar[st], h * D, b * p.Sq);
    tma_load_2d_addr(base + 3 * kPOp + 128 * kPRow, &tmdO, &ld_bar[st], h * D, b * p.Sq + 128);
    const uint32_t side = smem_base + side_off + st * kPSide;
    bulk_load(side, p.lse + bh * p.Sq, lse_bytes, &ld_bar[st]);
    bulk_load(side + 1024, p.delta + bh * p.Sq, lse_bytes, &ld_bar[st]);
    if (DROP) bulk_load(side + 2048, p.p_keep + bh * p.Sq * nkb * 4, keep_bytes, &ld_bar[st]);
  };

  if (tid == 0) {
    tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmdO); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmV);
    mbar_init(&ld_bar[0], 1); mbar_init(&ld_bar[1], 1);
    mbar_init(&s_bar[0], 1); mbar_init(&s_bar[1], 1);
    mbar_init(&dp_bar[0], 1); mbar_init(&dp_bar[1], 1);
    mbar_init(&done_bar, 1);
    fence_mbar_init();
    issue_loads(blockIdx.x, 0);
  }
  if (warp == 1) {
    tmem_alloc(&tmem_slot, 512u);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;
  constexpr uint32_t dq_col = 256u, dk_col = 320u, dv_col = 384u;

  const int row = quad * 32 + lane;
  const float sl2 = p.scale * kLog2e;
  const float dsc = DROP ? p.drop_p.scale : 1.0f;
  const float inv_dsc = 1.0f / dsc;
  const uint32_t t_row = tmem_base + ((uint32_t)(quad * 32) << 16);
  const uint32_t idesc_q = make_idesc_bf16(128, D, 0, 1);   // dQ: A K-major (slabs), B MN-major (K tile)
  const uint32_t idesc_t = make_idesc_bf16(128, D, 1, 1);   // dK / dV: A MN-major (slabs), B MN-major

  // ---- MMA batches (one elected thread of warp 0); st = operand stage of the item ----
  auto issue_s = [&](int st, int qt, int kh) {     // S_h = Q_qt K_h^T -> buffer kh
    const uint32_t base = smem_base + st * kPStage;
    const uint32_t n = (uint32_t)(kh == nkh - 1 ? wlast : 128);
    const uint32_t idesc = make_idesc_bf16(128, n, 0, 0);
    const uint32_t aq = base + 2 * kPOp + (uint32_t)qt * 128u * kPRow, bk = base + (uint32_t)kh * 128u * kPRow;
#pragma unroll
    for (int k = 0; k < D / 16; ++k)
      umma_bf16(tmem_base + 128u * kh, make_smem_desc(aq + k * 32, 16, kSbo64, 4), make_smem_desc(bk + k * 32, 16, kSbo64, 4),
                idesc, k > 0 ? 1u : 0u);
    umma_commit(&s_bar[kh]);
  };
  auto issue_dp_dv = [&](int st, int qt, int kh) {  // dP_h = dO_qt V_h^T over S_h ; dV_h += P_drop_h^T dO_qt
    const uint32_t base = smem_base + st * kPStage;
    const uint32_t n = (uint32_t)(kh == nkh - 1 ? wlast : 128);
    const uint32_t idesc = make_idesc_bf16(128, n, 0, 0);
    const uint32_t ad = base + 3 * kPOp + (uint32_t)qt * 128u * kPRow, bv = base + kPOp + (uint32_t)kh * 128u * kPRow;
#pragma unroll
    for (int k = 0; k < D / 16; ++k)
      umma_bf16(tmem_base + 128u * kh, make_smem_desc(ad + k * 32, 16, kSbo64, 4), make_smem_desc(bv + k * 32, 16, kSbo64, 4),
                idesc, k > 0 ? 1u : 0u);
    for (int kk = 0; kk < 8; ++kk)
      umma_bf16(tmem_base + dv_col + 32u * kh,
                make_smem_desc(sSlab + (uint32_t)(2 * kh) * kSlabBytes + (uint32_t)kk * 2048u, kSlabBytes, 1024, 2),
                make_smem_desc(ad + (uint32_t)kk * 16u * kPRow, kSbo64, kSbo64, 4), idesc_t, (qt > 0 || kk > 0) ? 1u : 0u);
    umma_commit(&dp_bar[kh]);
  };
  auto issue_dq_dk = [&](int st, int qt, int kh) {  // dQ_qt += dS_h K_h ; dK_h += dS_h^T Q_qt
    const uint32_t base = smem_base + st * kPStage;
    const int nks = (kh == nkh - 1 ? wlast : 128) >> 4;
    const uint32_t aq = base + 2 * kPOp + (uint32_t)qt * 128u * kPRow;
    for (int k2 = 0; k2 < nks; ++k2) {
      const int kk = 8 * kh + k2;           // 16-key step inside the whole key range
      umma_bf16(tmem_base + dq_col + 32u * qt,
                make_smem_desc(sSlab + (uint32_t)(kk >> 2) * kSlabBytes + (uint32_t)(kk & 3) * 32u, 16, 1024, 2),
                make_smem_desc(base + (uint32_t)kk * 16u * kPRow, kSbo64, kSbo64, 4), idesc_q, (kh > 0 || k2 > 0) ? 1u : 0u);
    }
    for (int kk = 0; kk < 8; ++kk)
      umma_bf16(tmem_base + dk_col + 32u * kh,
                make_smem_desc(sSlab + (uint32_t)(2 * kh) * kSlabBytes + (uint32_t)kk * 2048u, kSlabBytes, 1024, 2),
                make_smem_desc(aq + (uint32_t)kk * 16u * kPRow, kSbo64, kSbo64, 4), idesc_t, (qt > 0 || kk > 0) ? 1u : 0u);
  };

  uint32_t ph = 0;   // query tiles processed so far by this CTA: phase of s_bar / dp_bar
  int it = 0;
#pragma unroll 1
  for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
    const int st = it & 1;
    const int b = item / p.nh, h = item - b * p.nh;
    const int next = item + (int)gridDim.x;
    // the previous item is completely retired here (its copy-out ended with a block barrier): its operand stage and
    // side data may be overwritten, the slabs are free
    if (tid == 0 && next < n_items) issue_loads(next, st ^ 1);
    if (warp >= 8) {
      const unsigned char* kvg = p.key_valid + (long long)b * p.Sk;
      const int w = warp - 8;
      const int j = w * 32 + lane;
      const bool v = (j < p.Sk) && (mode == MMFM_MASK_CAUSAL || kvg[j] != 0);
      const uint32_t m = __ballot_sync(0xffffffffu, v);
      if (lane == 0) s_colbits[w] = m;
    }
    DBG_BP(0);
    mbar_wait(&ld_bar[st], (uint32_t)((it >> 1) & 1));   // operands + side data of this item have landed
    __syncthreads();
    DBG_BP(1);
    if (it == 0 && warp == 0) {
      if (elect_one()) {
        tc_fence_after();
        for (int kh = 0; kh < nkh; ++kh) issue_s(st, 0, kh);
      }
      __syncwarp();
    }
    const float* s_lse = reinterpret_cast<const float*>(smem_al + side_off + st * kPSide);
    const float* s_dl = s_lse + 256;
    const uint8_t* s_keep = smem_al + side_off + st * kPSide + 2048;

#pragma unroll 1
    for (int qt = 0; qt < nqt; ++qt, ++ph) {
      const uint32_t par = ph & 1u;
      const int i = qt * 128 + row;
      const bool rok = i < p.Sq;
      const float lse2 = rok ? s_lse[i] * kLog2e : INFINITY;
      const float dl = rok ? s_dl[i] * inv_dsc : 0.f;
      uint32_t aws[2] = {0u, 0u};
      uint2 kpre[2] = {make_uint2(0xFFFFFFFFu, 0xFFFFFFFFu), make_uint2(0xFFFFFFFFu, 0xFFFFFFFFu)};
#pragma unroll
      for (int kh = 0; kh < 2; ++kh) {
        const int c = grp + 4 * kh;
        if (c < nch) {
          if (DROP && rok) kpre[kh] = *reinterpret_cast<const uint2*>(s_keep + ((size_t)i * nkb + (c >> 1)) * 8);
          uint32_t aw = s_colbits[c];
          const int rel = i - 32 * c;
          if (mode == MMFM_MASK_KEY_OR_DIAG) {
            if (rel >= 0 && rel < 32 && i < p.Sk) aw |= 1u << rel;
          } else if (mode == MMFM_MASK_CAUSAL) {
            aw &= (rel >= 31) ? 0xFFFFFFFFu : (rel < 0 ? 0u : ((2u << rel) - 1u));
          }
          aws[kh] = aw;
        }
      }
      uint32_t pk[2][16];    // p as packed bf16, kept for pass B
      uint32_t pdq[2][16];   // p * keep as packed bf16, kept for pass B

      // ---------------- pass A (both halves): probabilities ----------------
#pragma unroll
      for (int kh = 0; kh < 2; ++kh) {
        if (kh >= nkh) break;
        const int c = grp + 4 * kh;
        DBG_BP(4 + 16 * qt + 4 * kh);
        mbar_wait(&s_bar[kh], par);
        tc_fence_after();
        DBG_BP(5 + 16 * qt + 4 * kh);
        if (c < nch) {
          const uint32_t aw = aws[kh];
          uint32_t km[4][2] = {{0u, 0u}, {0u, 0u}, {0u, 0u}, {0u, 0u}};
          if (DROP) {
            const uint2 w2 = kpre[kh];
            const int sh = 8 * (c & 1);   // second 32-column chunk of the 64-key block: n-tiles 4..7 -> bits 8..15
            keep_msb_words((w2.x & 0xFFFFu) >> sh, km[0]);
            keep_msb_words((w2.x >> 16) >> sh, km[1]);
            keep_msb_words((w2.y & 0xFFFFu) >> sh, km[2]);
            keep_msb_words((w2.y >> 16) >> sh, km[3]);
          }
          const bool masked = __any_sync(0xffffffffu, aw != 0xFFFFFFFFu);
          uint32_t rs[2][16];
          tmem_ld16(t_row + 32u * c, rs[0]);
          tmem_ld16(t_row + 32u * c + 16u, rs[1]);
          tmem_ld_wait();
          uint32_t pdk[2][8];
          if (masked) {
            bwd_prob_half<true, DROP, 0>(rs[0], aw, sl2, lse2, km, &pk[kh][0], pdk[0]);
            bwd_prob_half<true, DROP, 1>(rs[1], aw, sl2, lse2, km, &pk[kh][8], pdk[1]);
          } else {
            bwd_prob_half<false, DROP, 0>(rs[0], aw, sl2, lse2, km, &pk[kh][0], pdk[0]);
            bwd_prob_half<false, DROP, 1>(rs[1], aw, sl2, lse2, km, &pk[kh][8], pdk[1]);
          }
#pragma unroll
          for (int hf = 0; hf < 2; ++hf) {
#pragma unroll
            for (int t = 0; t < 8; ++t) pdq[kh][8 * hf + t] = pdk[hf][t];
#pragma unroll
            for (int q4 = 0; q4 < 2; ++q4) {
              const int j16 = (c & 1) * 4 + hf * 2 + q4;
              const uint32_t addr = sSlab + (uint32_t)(c >> 1) * kSlabBytes + (uint32_t)row * 128u + (uint32_t)((j16 ^ (row & 7)) * 16);
              st_shared_v4(addr, pdk[hf][4 * q4], pdk[hf][4 * q4 + 1], pdk[hf][4 * q4 + 2], pdk[hf][4 * q4 + 3]);
            }
          }
        }
        DBG_BP(6 + 16 * qt + 4 * kh);
        tc_fence_before();
        fence_proxy_async();
        __syncthreads();
        DBG_BP(7 + 16 * qt + 4 * kh);
        if (warp == 0) {
          if (elect_one()) {
            tc_fence_after();
            issue_dp_dv(st, qt, kh);
          }
          __syncwarp();
        }
      }

      // ---------------- pass B (both halves): dS into the slab P_drop just left ----------------
#pragma unroll
      for (int kh = 0; kh < 2; ++kh) {
        if (kh >= nkh) break;
        const int c = grp + 4 * kh;
        DBG_BP(12 + 16 * qt + 4 * kh);
        mbar_wait(&dp_bar[kh], par);   // dP_h is there, and the dV product has finished reading this half's slabs
        tc_fence_after();
        DBG_BP(13 + 16 * qt + 4 * kh);
        if (c < nch) {
          const uint32_t aw = aws[kh];
          const bool masked = __any_sync(0xffffffffu, aw != 0xFFFFFFFFu);
          uint32_t rd[2][16];
          tmem_ld16(t_row + 32u * c, rd[0]);
          tmem_ld16(t_row + 32u * c + 16u, rd[1]);
          tmem_ld_wait();
          uint32_t dsk[2][8];
          if (masked) {
            bwd_ds_half<true>(rd[0], aw & 0xFFFFu, dl, &pk[kh][0], &pdq[kh][0], dsk[0]);
            bwd_ds_half<true>(rd[1], aw >> 16, dl, &pk[kh][8], &pdq[kh][8], dsk[1]);
          } else {
            bwd_ds_half<false>(rd[0], 0xFFFFu, dl, &pk[kh][0], &pdq[kh][0], dsk[0]);
            bwd_ds_half<false>(rd[1], 0xFFFFu, dl, &pk[kh][8], &pdq[kh][8], dsk[1]);
          }
#pragma unroll
          for (int hf = 0; hf < 2; ++hf) {
#pragma unroll
            for (int q4 = 0; q4 < 2; ++q4) {
              const int j16 = (c & 1) * 4 + hf * 2 + q4;
              const uint32_t addr = sSlab + (uint32_t)(c >> 1) * kSlabBytes + (uint32_t)row * 128u + (uint32_t)((j16 ^ (row & 7)) * 16);
              st_shared_v4(addr, dsk[hf][4 * q4], dsk[hf][4 * q4 + 1], dsk[hf][4 * q4 + 2], dsk[hf][4 * q4 + 3]);
            }
          }
        }
        DBG_BP(14 + 16 * qt + 4 * kh);
        tc_fence_before();
        fence_proxy_async();
        __syncthreads();
        DBG_BP(15 + 16 * qt + 4 * kh);
        if (warp == 0) {
          if (elect_one()) {
            tc_fence_after();
            issue_dq_dk(st, qt, kh);
            const bool last_q = (qt + 1 == nqt);
            if (last_q && kh == nkh - 1) umma_commit(&done_bar);       // every product of this item is issued
            if (!last_q) {
              issue_s(st, qt + 1, kh);                                  // its commit also covers the batch above
            } else if (next < n_items) {
              // first scores of the NEXT item, so its first pass finds them ready
              if (kh == 0) { mbar_wait(&ld_bar[st ^ 1], (uint32_t)(((it + 1) >> 1) & 1)); tc_fence_after(); }
              issue_s(st ^ 1, 0, kh);
            }
          }
          __syncwarp();
        }
      }
    }

    // ---------------- read-out: 16-column pieces over the 4 thread groups ----------------
    //   piece 0..3   : dQ of query tile piece/2, column half piece&1          (TMEM lane = query row)
    //   piece 4..11  : (kh, which, half) = ((piece-4)/4, ((piece-4)/2)&1, (piece-4)&1); which 0 dK, 1 dV (lane = key row)
    DBG_BP(40);
    mbar_wait(&done_bar, (uint32_t)(it & 1));
    tc_fence_after();
    DBG_BP(41);
    uint32_t r[3][16];
#pragma unroll
    for (int u = 0; u < 3; ++u) {
      const int piece = grp + 4 * u;
      uint32_t col;
      if (piece < 4) col = dq_col + 32u * (piece >> 1) + 16u * (piece & 1);
      else {
        const int q = piece - 4;
        col = ((q >> 1) & 1 ? dv_col : dk_col) + 32u * (q >> 2) + 16u * (q & 1);
      }
      tmem_ld16(t_row + col, r[u]);
    }
    tmem_ld_wait();
    tc_fence_before();
    // stage the 6 output tiles ([128 rows][32 bf16]) in the slab area (every product that read it is complete), then
    // copy them out with 4 lanes per 64-byte row: a warp store covers 8 full rows instead of 32 half-filled sectors
    constexpr int kOutPitch = 80;   // bytes per staged row (64 + 16): 16-byte accesses of a quarter-warp hit distinct banks
    uint8_t* stage_o = smem_al + 2 * kPStage;
#pragma unroll
    for (int u = 0; u < 3; ++u) {
      const int piece = grp + 4 * u;
      int tile, half;
      float fs;
      if (piece < 4) { tile = piece >> 1; half = piece & 1; fs = p.scale * dsc; }
      else {
        const int q = piece - 4, kh = q >> 2, which = (q >> 1) & 1;
        tile = 2 + 2 * kh + which; half = q & 1; fs = which ? dsc : p.scale * dsc;
      }
      uint8_t* dst = stage_o + (tile * 128 + row) * kOutPitch + half * 32;
#pragma unroll
      for (int k = 0; k < 16; k += 8)
        *reinterpret_cast<uint4*>(dst + 2 * k) =
            make_uint4(pack_bf16x2(__uint_as_float(r[u][k]) * fs, __uint_as_float(r[u][k + 1]) * fs),
                       pack_bf16x2(__uint_as_float(r[u][k + 2]) * fs, __uint_as_float(r[u][k + 3]) * fs),
                       pack_bf16x2(__uint_as_float(r[u][k + 4]) * fs, __uint_as_float(r[u][k + 5]) * fs),
                       pack_bf16x2(__uint_as_float(r[u][k + 6]) * fs, __uint_as_float(r[u][k + 7]) * fs));
    }
    __syncthreads();   // staged; also: every thread has its accumulators out of TMEM before the next item's products
    DBG_BP(42);
#pragma unroll 1
    for (int k6 = 0; k6 < 6; ++k6) {
      const int idx = k6 * kFusedThreads + tid;      // (tile, row, 16-byte piece)
      const int tile = idx >> 9, rr = (idx >> 2) & 127, q4 = idx & 3;
      bf16* dst = nullptr;
      if (tile < 2) {
        const int i = tile * 128 + rr;
        if (tile < nqt && i < p.Sq) dst = p.dq + ((long long)b * p.Sq + i) * p.lddq + h * D + 8 * q4;
      } else {
        const int kh = (tile - 2) >> 1, which = (tile - 2) & 1, j = kh * 128 + rr;
        if (kh < nkh && j < p.Sk)
          dst = (which ? p.dv + ((long long)b * p.Sk + j) * p.lddv : p.dk + ((long long)b * p.Sk + j) * p.lddk) + h * D + 8 * q4;
      }
      if (dst != nullptr)
        *reinterpret_cast<uint4*>(dst) = *reinterpret_cast<const uint4*>(stage_o + (tile * 128 + rr) * kOutPitch + q4 * 16);
    }
    __syncthreads();   // the staging area becomes the next item's P_drop slabs; side data / stage st are retired
    DBG_BP(43);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512u);
}

}  // namespace mmfm

using namespace mmfm;

namespace mmfm {
int launch_attn_bwd_persist(const mmfm_attn_args* a, const AttnParams& p, cudaStream_t st) {
  constexpr int D = 32;
  const int npk = (a->Sk + 15) / 16 * 16;
  const uint64_t width = (uint64_t)a->n_heads * D;
  const bool drop = a->drop_p.thresh != 0u;
  CUtensorMap tq, tdo, tk, tv;
  if (int rc = make_tmap_bf16_2d(&tq, a->q, (uint64_t)a->B * a->Sq, width, (uint64_t)a->ldq, D, 128, TMA_SW_64)) return rc;
  if (int rc = make_tmap_bf16_2d(&tdo, a->d_o, (uint64_t)a->B * a->Sq, width, (uint64_t)a->lddo, D, 128, TMA_SW_64)) return rc;
  if (int rc = make_tmap_bf16_2d(&tk, a->k, (uint64_t)a->B * a->Sk, width, (uint64_t)a->ldk, D, npk, TMA_SW_64)) return rc;
  if (int rc = make_tmap_bf16_2d(&tv, a->v, (uint64_t)a->B * a->Sk, width, (uint64_t)a->ldv, D, npk, TMA_SW_64)) return rc;
  static bool attr_set = false;
  if (!attr_set) {
    MMFM_CHECK_CUDA(cudaFuncSetAttribute(attn_bwd_persist_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kPersistSmem));
    MMFM_CHECK_CUDA(cudaFuncSetAttribute(attn_bwd_persist_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kPersistSmem));
    attr_set = true;
  }
  const int n_items = a->B * a->n_heads;
  int grid = device_sm_count();
  if (grid > n_items) grid = n_items;
  if (drop) MMFM_CHECK_CUDA(launch_pdl(attn_bwd_persist_kernel<true>, dim3(grid), dim3(kFusedThreads), kPersistSmem, st, tq, tdo, tk, tv, p, npk, n_items));
  else MMFM_CHECK_CUDA(launch_pdl(attn_bwd_persist_kernel<false>, dim3(grid), dim3(kFusedThreads), kPersistSmem, st, tq, tdo, tk, tv, p, npk, n_items));
  return 0;
}
}  // namespace mmfm
