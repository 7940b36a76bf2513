// Masked multi-head attention, forward and backward (flash-style: no (B,h,S,S) score / mask tensors in HBM).
//
// Replaces F.scaled_dot_product_attention(q, k, v, attn_mask=bool (B,h,S,S), dropout_p) and its autograd backward
// (reference src/multi_modal/mm_utils.py:105-112 self-attention, :143-150 cross-attention).  The (B,S,S) int64 masks
// the reference materialises (mm.py:152-158 encoder / cross, :178-194 decoder) are evaluated as predicates from
// compact descriptors: per-key validity bytes (B,Sk), a diagonal flag, a causal flag and optional modality ids.
//
// Tiling: one CTA = 64 rows (4 warps x 16) of one (batch, head); it streams 64-wide column blocks through a
// double-buffered cp.async pipeline; scores live in mma.sync accumulators, the online-softmax state in registers.
// Column blocks that the mask rules out entirely are skipped.  d_head 32 and 64.
//   fwd       rows = queries, cols = keys : S = Q K^T -> P -> O += P V ; writes O (after output dropout), LSE, keep bits
//   bwd prep  delta = rowsum(dO * O), dO <- dO * output-dropout mask
//   bwd dq    rows = queries, cols = keys : dQ += (P * (dP - delta)) K
//   bwd dkv   rows = keys, cols = queries : dV += P_drop^T dO ; dK += dS^T Q
// Dropout on the probabilities uses the interleaved byte layout documented in oracle/philox_ref.py (one Philox
// call = the 16 elements one thread owns in a 64-column block); the forward stores the 16 keep bits so the
// transposed backward kernel does not have to regenerate them element by element.
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
MMFM_DEVINL void cp_async16(uint32_t dst, const void* src, bool valid) {
  const int sz = valid ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(sz) : "memory");
}
MMFM_DEVINL void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
MMFM_DEVINL void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
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

template <int D>
struct TileCfg {
  static constexpr int kPitch = D + 8;              // elements; keeps ldmatrix rows on distinct banks
  static constexpr int kBytes = kTile * kPitch * 2;  // one 64 x D tile
};

// 64 x D bf16 tile: rows row0 .. row0+63 of a matrix with `nrows` valid rows (others zero-filled)
template <int D>
MMFM_DEVINL void load_tile(uint32_t sdst, const bf16* g, long long ld, int row0, int nrows) {
  constexpr int kChunks = D / 8;
#pragma unroll
  for (int c = threadIdx.x; c < kTile * kChunks; c += kAttnThreads) {
    const int r = c / kChunks, cc = c - r * kChunks;
    const bool ok = (row0 + r) < nrows;
    const bf16* src = g + (long long)(ok ? row0 + r : 0) * ld + cc * 8;
    cp_async16(sdst + (uint32_t)(r * TileCfg<D>::kPitch + cc * 8) * 2, src, ok);
  }
}

// A fragments (16 rows of this warp x D) from a row-major tile
template <int D>
MMFM_DEVINL void load_a_frags(uint32_t stile, int warp, int lane, uint32_t (&f)[D / 16][4]) {
  const int r = warp * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
#pragma unroll
  for (int ks = 0; ks < D / 16; ++ks) {
    const int c = ks * 16 + (lane >> 4) * 8;
    ldsm_x4(f[ks], stile + (uint32_t)(r * TileCfg<D>::kPitch + c) * 2);
  }
}

// acc[8][4] (16 x 64) = A(16 x D) . T^T where T is a 64 x D row-major tile (n = tile row, k = tile column)
template <int D>
MMFM_DEVINL void mma_rowtile_nt(float (&acc)[8][4], const uint32_t (&a)[D / 16][4], uint32_t stile, int lane) {
#pragma unroll
  for (int ks = 0; ks < D / 16; ++ks) {
#pragma unroll
    for (int np = 0; np < 4; ++np) {
      uint32_t b[4];
      const int r = np * 16 + (lane >> 4) * 8 + (lane & 7);
      const int c = ks * 16 + ((lane >> 3) & 1) * 8;
      ldsm_x4(b, stile + (uint32_t)(r * TileCfg<D>::kPitch + c) * 2);
      const uint32_t b0[2] = {b[0], b[1]}, b1[2] = {b[2], b[3]};
      mma_16816(acc[2 * np], a[ks], b0);
      mma_16816(acc[2 * np + 1], a[ks], b1);
    }
  }
}

// out[D/8][4] (16 x D) += P(16 x 64, packed bf16 A fragments) . T where T is a 64 x D row-major tile (k = tile row)
template <int D>
MMFM_DEVINL void mma_rowtile_nn(float (&out)[D / 8][4], const uint32_t (&p)[4][4], uint32_t stile, int lane) {
#pragma unroll
  for (int t = 0; t < 4; ++t) {
#pragma unroll
    for (int dp = 0; dp < D / 16; ++dp) {
      uint32_t b[4];
      const int r = t * 16 + ((lane >> 3) & 1) * 8 + (lane & 7);
      const int c = dp * 16 + (lane >> 4) * 8;
      ldsm_x4_t(b, stile + (uint32_t)(r * TileCfg<D>::kPitch + c) * 2);
      const uint32_t b0[2] = {b[0], b[1]}, b1[2] = {b[2], b[3]};
      mma_16816(out[2 * dp], p[t], b0);
      mma_16816(out[2 * dp + 1], p[t], b1);
    }
  }
}

MMFM_DEVINL void pack_p(const float (&s)[8][4], uint32_t (&p)[4][4]) {
#pragma unroll
  for (int t = 0; t < 4; ++t) {
    p[t][0] = pack_bf16x2(s[2 * t][0], s[2 * t][1]);
    p[t][1] = pack_bf16x2(s[2 * t][2], s[2 * t][3]);
    p[t][2] = pack_bf16x2(s[2 * t + 1][0], s[2 * t + 1][1]);
    p[t][3] = pack_bf16x2(s[2 * t + 1][2], s[2 * t + 1][3]);
  }
}

// mask predicate for (query i, key j); kv = key_valid[b, j]
struct MaskCtx {
  int mode;
  const short* mod_q;
  const short* mod_k;
  int Sq, Sk;
};
MMFM_DEVINL bool allowed(const MaskCtx& mc, int i, int j, bool kv) {
  if (i >= mc.Sq || j >= mc.Sk) return false;
  bool a = (mc.mode == MMFM_MASK_CAUSAL) ? (j <= i) : kv;
  if (mc.mode == MMFM_MASK_KEY_OR_DIAG) a = a || (i == j);
  if (mc.mod_q) a = a || (mc.mod_q[i] != mc.mod_k[j]);
  return a;
}
// can a whole (query block, key block) pair be skipped?  CTA-uniform.  any_valid = some key of the block is valid.
MMFM_DEVINL bool skip_block(const MaskCtx& mc, int q0, int k0, bool any_valid) {
  if (mc.mod_q) return false;
  if (mc.mode == MMFM_MASK_CAUSAL) return k0 > q0 + kTile - 1;
  if (any_valid) return false;
  if (mc.mode == MMFM_MASK_KEY) return true;
  return (k0 > q0 + kTile - 1) || (k0 + kTile - 1 < q0);  // KEY_OR_DIAG: keep blocks crossing the diagonal
}

// 16 random bytes of the probability-dropout field: row = (b*nh+h)*Sq + i, 64-column block blk, quad lane ql
MMFM_DEVINL uint4 pdrop_bytes(unsigned long long seed, uint32_t site, unsigned long long row, uint32_t nblk,
                              uint32_t blk, uint32_t ql) {
  const unsigned long long g = (row * nblk + blk) * 4ull + ql;
  return philox4x32((uint32_t)g, (uint32_t)(g >> 32), site, 1u, (uint32_t)seed, (uint32_t)(seed >> 32));
}
MMFM_DEVINL uint32_t keep_bits16(const uint4& w, uint32_t thresh) {
  uint32_t bits = 0;
#pragma unroll
  for (int b = 0; b < 16; ++b) bits |= (drop_byte(w, b) >= thresh ? 1u : 0u) << b;
  return bits;
}

// ------------------------------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------------------------------
template <int D, bool DROP>
__global__ void __launch_bounds__(kAttnThreads) attn_fwd_kernel(const AttnParams p) {
  using TC = TileCfg<D>;
  extern __shared__ __align__(16) uint8_t smem_dyn[];
  uint8_t* sQ = smem_dyn;
  uint8_t* sK[2] = {smem_dyn + TC::kBytes, smem_dyn + 2 * TC::kBytes};
  uint8_t* sV[2] = {smem_dyn + 3 * TC::kBytes, smem_dyn + 4 * TC::kBytes};
  unsigned char(*sValid)[kTile] = reinterpret_cast<unsigned char(*)[kTile]>(smem_dyn + 5 * TC::kBytes);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, ql = lane & 3;
  const int q0 = blockIdx.x * kTile, h = blockIdx.y, b = blockIdx.z;
  const int nkb = (p.Sk + kTile - 1) / kTile;
  const MaskCtx mc{p.mask_mode, p.mod_q, p.mod_k, p.Sq, p.Sk};
  const bf16* qg = p.q + (long long)b * p.Sq * p.ldq + h * D;
  const bf16* kg = p.k + (long long)b * p.Sk * p.ldk + h * D;
  const bf16* vg = p.v + (long long)b * p.Sk * p.ldv + h * D;
  const unsigned char* kvg = p.key_valid + (long long)b * p.Sk;

  auto load_kv = [&](int kb, int st) {
    load_tile<D>(smem_u32(sK[st]), kg, p.ldk, kb * kTile, p.Sk);
    load_tile<D>(smem_u32(sV[st]), vg, p.ldv, kb * kTile, p.Sk);
    if (threadIdx.x < kTile) {
      const int j = kb * kTile + threadIdx.x;
      sValid[st][threadIdx.x] = (j < p.Sk) ? kvg[j] : 0;
    }
  };

  load_tile<D>(smem_u32(sQ), qg, p.ldq, q0, p.Sq);
  load_kv(0, 0);
  cp_async_commit();

  const int i0 = q0 + warp * 16 + g, i1 = i0 + 8;
  const float sl2 = p.scale * kLog2e;
  float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;
  float o[D / 8][4];
#pragma unroll
  for (int n = 0; n < D / 8; ++n) o[n][0] = o[n][1] = o[n][2] = o[n][3] = 0.f;
  uint32_t qf[D / 16][4];

  unsigned long long seed_p = 0ull;
  if (DROP) seed_p = *p.drop_p.seed;
  const unsigned long long prow0 = ((unsigned long long)(b * p.nh + h)) * p.Sq + i0;

  for (int kb = 0; kb < nkb; ++kb) {
    const int st = kb & 1;
    if (kb + 1 < nkb) load_kv(kb + 1, st ^ 1);
    cp_async_commit();
    cp_async_wait<1>();
    __syncthreads();
    if (kb == 0) load_a_frags<D>(smem_u32(sQ), warp, lane, qf);
    const bool any_valid = __syncthreads_or(threadIdx.x < kTile ? (int)sValid[st][threadIdx.x] : 0) != 0;
    if (!skip_block(mc, q0, kb * kTile, any_valid)) {
      float s[8][4];
#pragma unroll
      for (int n = 0; n < 8; ++n) s[n][0] = s[n][1] = s[n][2] = s[n][3] = 0.f;
      mma_rowtile_nt<D>(s, qf, smem_u32(sK[st]), lane);
      float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
      for (int n = 0; n < 8; ++n) {
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int jl = 8 * n + 2 * ql + e, j = kb * kTile + jl;
          const bool kv = sValid[st][jl] != 0;
          s[n][e] = allowed(mc, i0, j, kv) ? s[n][e] * sl2 : -INFINITY;
          s[n][2 + e] = allowed(mc, i1, j, kv) ? s[n][2 + e] * sl2 : -INFINITY;
          mx0 = fmaxf(mx0, s[n][e]);
          mx1 = fmaxf(mx1, s[n][2 + e]);
        }
      }
      mx0 = quad_max(mx0);
      mx1 = quad_max(mx1);
      const float mn0 = fmaxf(m0, mx0), mn1 = fmaxf(m1, mx1);
      const float base0 = (mn0 == -INFINITY) ? 0.f : mn0, base1 = (mn1 == -INFINITY) ? 0.f : mn1;
      const float al0 = fast_exp2(m0 - base0), al1 = fast_exp2(m1 - base1);
      m0 = mn0;
      m1 = mn1;
      float rs0 = 0.f, rs1 = 0.f;
#pragma unroll
      for (int n = 0; n < 8; ++n) {
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          s[n][e] = fast_exp2(s[n][e] - base0);
          s[n][2 + e] = fast_exp2(s[n][2 + e] - base1);
          rs0 += s[n][e];
          rs1 += s[n][2 + e];
        }
      }
      l0 = l0 * al0 + rs0;
      l1 = l1 * al1 + rs1;
#pragma unroll
      for (int n = 0; n < D / 8; ++n) {
        o[n][0] *= al0; o[n][1] *= al0; o[n][2] *= al1; o[n][3] *= al1;
      }
      if (DROP) {
        const uint4 w0 = pdrop_bytes(seed_p, p.drop_p.site, prow0, (uint32_t)nkb, (uint32_t)kb, (uint32_t)ql);
        const uint4 w1 = pdrop_bytes(seed_p, p.drop_p.site, prow0 + 8, (uint32_t)nkb, (uint32_t)kb, (uint32_t)ql);
        const uint32_t k0 = keep_bits16(w0, p.drop_p.thresh), k1 = keep_bits16(w1, p.drop_p.thresh);
#pragma unroll
        for (int n = 0; n < 8; ++n) {
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            s[n][e] = ((k0 >> (2 * n + e)) & 1u) ? s[n][e] * p.drop_p.scale : 0.f;
            s[n][2 + e] = ((k1 >> (2 * n + e)) & 1u) ? s[n][2 + e] * p.drop_p.scale : 0.f;
          }
        }
        if (p.p_keep) {
          const long long bh = (long long)(b * p.nh + h);
          if (i0 < p.Sq) p.p_keep[((bh * p.Sq + i0) * nkb + kb) * 4 + ql] = (unsigned short)k0;
          if (i1 < p.Sq) p.p_keep[((bh * p.Sq + i1) * nkb + kb) * 4 + ql] = (unsigned short)k1;
        }
      }
      uint32_t pf[4][4];
      pack_p(s, pf);
      mma_rowtile_nn<D>(o, pf, smem_u32(sV[st]), lane);
    }
    __syncthreads();
  }

  l0 = quad_sum(l0);
  l1 = quad_sum(l1);
  const float inv0 = l0 > 0.f ? 1.0f / l0 : 0.f, inv1 = l1 > 0.f ? 1.0f / l1 : 0.f;
  if (ql == 0) {
    float* lse = p.lse + ((long long)(b * p.nh + h)) * p.Sq;
    if (i0 < p.Sq) lse[i0] = (l0 > 0.f) ? (m0 + log2f(l0)) * kLn2 : -INFINITY;
    if (i1 < p.Sq) lse[i1] = (l1 > 0.f) ? (m1 + log2f(l1)) * kLn2 : -INFINITY;
  }
  unsigned long long seed_o = 0ull;
  const bool drop_o = p.drop_o.thresh != 0u;
  if (drop_o) seed_o = *p.drop_o.seed;
  const uint32_t gpr_o = (uint32_t)((p.nh * D + 15) >> 4);
#pragma unroll
  for (int n = 0; n < D / 8; ++n) {
    float v00 = o[n][0] * inv0, v01 = o[n][1] * inv0, v10 = o[n][2] * inv1, v11 = o[n][3] * inv1;
    const int col = h * D + 8 * n + 2 * ql;
    if (drop_o) {
      const uint4 w0 = drop_bytes16(seed_o, p.drop_o.site, (uint64_t)((long long)b * p.Sq + i0), gpr_o, (uint32_t)(col >> 4));
      const uint4 w1 = drop_bytes16(seed_o, p.drop_o.site, (uint64_t)((long long)b * p.Sq + i1), gpr_o, (uint32_t)(col >> 4));
      const int bb = col & 15;
      v00 = drop_byte(w0, bb) < p.drop_o.thresh ? 0.f : v00 * p.drop_o.scale;
      v01 = drop_byte(w0, bb + 1) < p.drop_o.thresh ? 0.f : v01 * p.drop_o.scale;
      v10 = drop_byte(w1, bb) < p.drop_o.thresh ? 0.f : v10 * p.drop_o.scale;
      v11 = drop_byte(w1, bb + 1) < p.drop_o.thresh ? 0.f : v11 * p.drop_o.scale;
    }
    if (i0 < p.Sq)
      *reinterpret_cast<uint32_t*>(p.o + ((long long)b * p.Sq + i0) * p.ldo + col) = pack_bf16x2(v00, v01);
    if (i1 < p.Sq)
      *reinterpret_cast<uint32_t*>(p.o + ((long long)b * p.Sq + i1) * p.ldo + col) = pack_bf16x2(v10, v11);
  }
}

// ------------------------------------------------------------------------------------------------------------
// backward prep: delta[b,h,i] = sum_c dO[b,i,hD+c] * O[b,i,hD+c]; dO <- dO * (output dropout mask * scale)
// ------------------------------------------------------------------------------------------------------------
template <int D>
__global__ void __launch_bounds__(256) attn_bwd_prep_kernel(const AttnParams p) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int H = p.nh * D;
  const long long R = (long long)p.B * p.Sq;
  const bool drop_o = p.drop_o.thresh != 0u;
  unsigned long long seed_o = 0ull;
  if (drop_o) seed_o = *p.drop_o.seed;
  const uint32_t gpr_o = (uint32_t)((H + 15) >> 4);
  for (long long r = (long long)blockIdx.x * 8 + warp; r < R; r += (long long)gridDim.x * 8) {
    const long long b = r / p.Sq;
    const int i = (int)(r - b * p.Sq);
    for (int c = lane * 8; c < H; c += 256) {
      uint4 dv = *reinterpret_cast<const uint4*>(p.d_o + r * p.lddo + c);
      const uint4 ov = *reinterpret_cast<const uint4*>(p.o + r * p.ldo + c);
      float2 d[4] = {unpack_bf16x2(dv.x), unpack_bf16x2(dv.y), unpack_bf16x2(dv.z), unpack_bf16x2(dv.w)};
      const float2 oo[4] = {unpack_bf16x2(ov.x), unpack_bf16x2(ov.y), unpack_bf16x2(ov.z), unpack_bf16x2(ov.w)};
      float s = 0.f;
#pragma unroll
      for (int j = 0; j < 4; ++j) s += d[j].x * oo[j].x + d[j].y * oo[j].y;
#pragma unroll
      for (int off = 1; off < D / 8; off <<= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
      if ((lane & (D / 8 - 1)) == 0) p.delta[(b * p.nh + c / D) * p.Sq + i] = s;
      if (drop_o) {
        const uint4 w = drop_bytes16(seed_o, p.drop_o.site, (uint64_t)r, gpr_o, (uint32_t)(c >> 4));
        const int bb = c & 15;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          d[j].x = drop_byte(w, bb + 2 * j) < p.drop_o.thresh ? 0.f : d[j].x * p.drop_o.scale;
          d[j].y = drop_byte(w, bb + 2 * j + 1) < p.drop_o.thresh ? 0.f : d[j].y * p.drop_o.scale;
        }
        dv.x = pack_bf16x2(d[0].x, d[0].y);
        dv.y = pack_bf16x2(d[1].x, d[1].y);
        dv.z = pack_bf16x2(d[2].x, d[2].y);
        dv.w = pack_bf16x2(d[3].x, d[3].y);
        *reinterpret_cast<uint4*>(p.d_o + r * p.lddo + c) = dv;
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------------------
// backward dQ: rows = queries, cols = keys
// ------------------------------------------------------------------------------------------------------------
template <int D, bool DROP>
__global__ void __launch_bounds__(kAttnThreads) attn_bwd_dq_kernel(const AttnParams p) {
  using TC = TileCfg<D>;
  extern __shared__ __align__(16) uint8_t smem_dyn[];
  uint8_t* sQ = smem_dyn;
  uint8_t* sdO = smem_dyn + TC::kBytes;
  uint8_t* sK[2] = {smem_dyn + 2 * TC::kBytes, smem_dyn + 3 * TC::kBytes};
  uint8_t* sV[2] = {smem_dyn + 4 * TC::kBytes, smem_dyn + 5 * TC::kBytes};
  unsigned char(*sValid)[kTile] = reinterpret_cast<unsigned char(*)[kTile]>(smem_dyn + 6 * TC::kBytes);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, ql = lane & 3;
  const int q0 = blockIdx.x * kTile, h = blockIdx.y, b = blockIdx.z;
  const int nkb = (p.Sk + kTile - 1) / kTile;
  const MaskCtx mc{p.mask_mode, p.mod_q, p.mod_k, p.Sq, p.Sk};
  const bf16* qg = p.q + (long long)b * p.Sq * p.ldq + h * D;
  const bf16* dog = p.d_o + (long long)b * p.Sq * p.lddo + h * D;
  const bf16* kg = p.k + (long long)b * p.Sk * p.ldk + h * D;
  const bf16* vg = p.v + (long long)b * p.Sk * p.ldv + h * D;
  const unsigned char* kvg = p.key_valid + (long long)b * p.Sk;

  auto load_kv = [&](int kb, int st) {
    load_tile<D>(smem_u32(sK[st]), kg, p.ldk, kb * kTile, p.Sk);
    load_tile<D>(smem_u32(sV[st]), vg, p.ldv, kb * kTile, p.Sk);
    if (threadIdx.x < kTile) {
      const int j = kb * kTile + threadIdx.x;
      sValid[st][threadIdx.x] = (j < p.Sk) ? kvg[j] : 0;
    }
  };
  load_tile<D>(smem_u32(sQ), qg, p.ldq, q0, p.Sq);
  load_tile<D>(smem_u32(sdO), dog, p.lddo, q0, p.Sq);
  load_kv(0, 0);
  cp_async_commit();

  const int i0 = q0 + warp * 16 + g, i1 = i0 + 8;
  const float sl2 = p.scale * kLog2e;
  const long long bh = (long long)(b * p.nh + h);
  const float lse0 = (i0 < p.Sq) ? p.lse[bh * p.Sq + i0] * kLog2e : INFINITY;
  const float lse1 = (i1 < p.Sq) ? p.lse[bh * p.Sq + i1] * kLog2e : INFINITY;
  const float dl0 = (i0 < p.Sq) ? p.delta[bh * p.Sq + i0] : 0.f;
  const float dl1 = (i1 < p.Sq) ? p.delta[bh * p.Sq + i1] : 0.f;
  float dq[D / 8][4];
#pragma unroll
  for (int n = 0; n < D / 8; ++n) dq[n][0] = dq[n][1] = dq[n][2] = dq[n][3] = 0.f;
  uint32_t qf[D / 16][4], dof[D / 16][4];
  unsigned long long seed_p = 0ull;
  if (DROP) seed_p = *p.drop_p.seed;
  const unsigned long long prow0 = (unsigned long long)bh * p.Sq + i0;

  for (int kb = 0; kb < nkb; ++kb) {
    const int st = kb & 1;
    if (kb + 1 < nkb) load_kv(kb + 1, st ^ 1);
    cp_async_commit();
    cp_async_wait<1>();
    __syncthreads();
    if (kb == 0) {
      load_a_frags<D>(smem_u32(sQ), warp, lane, qf);
      load_a_frags<D>(smem_u32(sdO), warp, lane, dof);
    }
    const bool any_valid = __syncthreads_or(threadIdx.x < kTile ? (int)sValid[st][threadIdx.x] : 0) != 0;
    if (!skip_block(mc, q0, kb * kTile, any_valid)) {
      float s[8][4], dp[8][4];
#pragma unroll
      for (int n = 0; n < 8; ++n) {
        s[n][0] = s[n][1] = s[n][2] = s[n][3] = 0.f;
        dp[n][0] = dp[n][1] = dp[n][2] = dp[n][3] = 0.f;
      }
      mma_rowtile_nt<D>(s, qf, smem_u32(sK[st]), lane);
      mma_rowtile_nt<D>(dp, dof, smem_u32(sV[st]), lane);
      uint32_t k0 = 0xFFFFu, k1 = 0xFFFFu;
      float dsc = 1.0f;
      if (DROP) {
        const uint4 w0 = pdrop_bytes(seed_p, p.drop_p.site, prow0, (uint32_t)nkb, (uint32_t)kb, (uint32_t)ql);
        const uint4 w1 = pdrop_bytes(seed_p, p.drop_p.site, prow0 + 8, (uint32_t)nkb, (uint32_t)kb, (uint32_t)ql);
        k0 = keep_bits16(w0, p.drop_p.thresh);
        k1 = keep_bits16(w1, p.drop_p.thresh);
        dsc = p.drop_p.scale;
      }
#pragma unroll
      for (int n = 0; n < 8; ++n) {
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int jl = 8 * n + 2 * ql + e, j = kb * kTile + jl;
          const bool kv = sValid[st][jl] != 0;
          const float p0 = allowed(mc, i0, j, kv) ? fast_exp2(s[n][e] * sl2 - lse0) : 0.f;
          const float p1 = allowed(mc, i1, j, kv) ? fast_exp2(s[n][2 + e] * sl2 - lse1) : 0.f;
          const float dp0 = ((k0 >> (2 * n + e)) & 1u) ? dp[n][e] * dsc : 0.f;
          const float dp1 = ((k1 >> (2 * n + e)) & 1u) ? dp[n][2 + e] * dsc : 0.f;
          s[n][e] = p0 * (dp0 - dl0);
          s[n][2 + e] = p1 * (dp1 - dl1);
        }
      }
      uint32_t pf[4][4];
      pack_p(s, pf);
      mma_rowtile_nn<D>(dq, pf, smem_u32(sK[st]), lane);
    }
    __syncthreads();
  }
#pragma unroll
  for (int n = 0; n < D / 8; ++n) {
    const int col = h * D + 8 * n + 2 * ql;
    if (i0 < p.Sq)
      *reinterpret_cast<uint32_t*>(p.dq + ((long long)b * p.Sq + i0) * p.lddq + col) =
          pack_bf16x2(dq[n][0] * p.scale, dq[n][1] * p.scale);
    if (i1 < p.Sq)
      *reinterpret_cast<uint32_t*>(p.dq + ((long long)b * p.Sq + i1) * p.lddq + col) =
          pack_bf16x2(dq[n][2] * p.scale, dq[n][3] * p.scale);
  }
}

// ------------------------------------------------------------------------------------------------------------
// backward dK, dV: rows = keys, cols = queries
// ------------------------------------------------------------------------------------------------------------
template <int D, bool DROP>
__global__ void __launch_bounds__(kAttnThreads) attn_bwd_dkv_kernel(const AttnParams p) {
  using TC = TileCfg<D>;
  extern __shared__ __align__(16) uint8_t smem_dyn[];
  uint8_t* sK = smem_dyn;
  uint8_t* sV = smem_dyn + TC::kBytes;
  uint8_t* sQ[2] = {smem_dyn + 2 * TC::kBytes, smem_dyn + 3 * TC::kBytes};
  uint8_t* sdO[2] = {smem_dyn + 4 * TC::kBytes, smem_dyn + 5 * TC::kBytes};
  float(*sLse)[kTile] = reinterpret_cast<float(*)[kTile]>(smem_dyn + 6 * TC::kBytes);
  float(*sDelta)[kTile] = reinterpret_cast<float(*)[kTile]>(smem_dyn + 6 * TC::kBytes + 2 * kTile * 4);
  __shared__ int sAny;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, ql = lane & 3;
  const int kblk = blockIdx.x, k0 = kblk * kTile, h = blockIdx.y, b = blockIdx.z;
  const int nqb = (p.Sq + kTile - 1) / kTile;
  const int nkb = (p.Sk + kTile - 1) / kTile;
  const MaskCtx mc{p.mask_mode, p.mod_q, p.mod_k, p.Sq, p.Sk};
  const long long bh = (long long)(b * p.nh + h);
  const bf16* qg = p.q + (long long)b * p.Sq * p.ldq + h * D;
  const bf16* dog = p.d_o + (long long)b * p.Sq * p.lddo + h * D;
  const bf16* kg = p.k + (long long)b * p.Sk * p.ldk + h * D;
  const bf16* vg = p.v + (long long)b * p.Sk * p.ldv + h * D;
  const unsigned char* kvg = p.key_valid + (long long)b * p.Sk;

  auto load_q = [&](int qb, int st) {
    load_tile<D>(smem_u32(sQ[st]), qg, p.ldq, qb * kTile, p.Sq);
    load_tile<D>(smem_u32(sdO[st]), dog, p.lddo, qb * kTile, p.Sq);
    if (threadIdx.x < kTile) {
      const int i = qb * kTile + threadIdx.x;
      sLse[st][threadIdx.x] = (i < p.Sq) ? p.lse[bh * p.Sq + i] * kLog2e : INFINITY;
      sDelta[st][threadIdx.x] = (i < p.Sq) ? p.delta[bh * p.Sq + i] : 0.f;
    }
  };
  if (threadIdx.x == 0) sAny = 0;
  load_tile<D>(smem_u32(sK), kg, p.ldk, k0, p.Sk);
  load_tile<D>(smem_u32(sV), vg, p.ldv, k0, p.Sk);
  load_q(0, 0);
  cp_async_commit();
  __syncthreads();
  if (threadIdx.x < kTile && k0 + threadIdx.x < p.Sk && kvg[k0 + threadIdx.x]) sAny = 1;

  const int j0 = k0 + warp * 16 + g, j1 = j0 + 8;  // this thread's key rows
  const bool kv0 = (j0 < p.Sk) ? kvg[j0] != 0 : false;
  const bool kv1 = (j1 < p.Sk) ? kvg[j1] != 0 : false;
  const float sl2 = p.scale * kLog2e;
  float dk[D / 8][4], dv[D / 8][4];
#pragma unroll
  for (int n = 0; n < D / 8; ++n) {
    dk[n][0] = dk[n][1] = dk[n][2] = dk[n][3] = 0.f;
    dv[n][0] = dv[n][1] = dv[n][2] = dv[n][3] = 0.f;
  }
  uint32_t kf[D / 16][4], vf[D / 16][4];
  // keep-bit addressing: word index ((bh*Sq + i)*nkb + kblk)*4 + (key%8)/2, bit ((key%64)/8)*2 + key%2
  const int kq = g >> 1;
  const int bit0 = (warp * 2) * 2 + (g & 1), bit1 = (warp * 2 + 1) * 2 + (g & 1);
  const float dsc = DROP ? p.drop_p.scale : 1.0f;

  for (int qb = 0; qb < nqb; ++qb) {
    const int st = qb & 1;
    if (qb + 1 < nqb) load_q(qb + 1, st ^ 1);
    cp_async_commit();
    cp_async_wait<1>();
    __syncthreads();
    if (qb == 0) {
      load_a_frags<D>(smem_u32(sK), warp, lane, kf);
      load_a_frags<D>(smem_u32(sV), warp, lane, vf);
    }
    const bool any_valid = sAny != 0;
    if (!skip_block(mc, qb * kTile, k0, any_valid)) {
      float s[8][4], dp[8][4];
#pragma unroll
      for (int n = 0; n < 8; ++n) {
        s[n][0] = s[n][1] = s[n][2] = s[n][3] = 0.f;
        dp[n][0] = dp[n][1] = dp[n][2] = dp[n][3] = 0.f;
      }
      mma_rowtile_nt<D>(s, kf, smem_u32(sQ[st]), lane);     // S^T[key, query]
      mma_rowtile_nt<D>(dp, vf, smem_u32(sdO[st]), lane);   // dP^T[key, query]
      float pd[8][4];                                        // dropped probabilities (for dV)
#pragma unroll
      for (int n = 0; n < 8; ++n) {
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int il = 8 * n + 2 * ql + e, i = qb * kTile + il;
          const float lse = sLse[st][il], dl = sDelta[st][il];
          bool keep0 = true, keep1 = true;
          if (DROP) {
            if (i < p.Sq) {
              const uint32_t wbits = p.p_keep[((bh * p.Sq + i) * nkb + kblk) * 4 + kq];
              keep0 = (wbits >> bit0) & 1u;
              keep1 = (wbits >> bit1) & 1u;
            }
          }
          const float p0 = allowed(mc, i, j0, kv0) ? fast_exp2(s[n][e] * sl2 - lse) : 0.f;
          const float p1 = allowed(mc, i, j1, kv1) ? fast_exp2(s[n][2 + e] * sl2 - lse) : 0.f;
          const float dp0 = keep0 ? dp[n][e] * dsc : 0.f;
          const float dp1 = keep1 ? dp[n][2 + e] * dsc : 0.f;
          pd[n][e] = keep0 ? p0 * dsc : 0.f;
          pd[n][2 + e] = keep1 ? p1 * dsc : 0.f;
          s[n][e] = p0 * (dp0 - dl);
          s[n][2 + e] = p1 * (dp1 - dl);
        }
      }
      uint32_t pf[4][4];
      pack_p(pd, pf);
      mma_rowtile_nn<D>(dv, pf, smem_u32(sdO[st]), lane);
      pack_p(s, pf);
      mma_rowtile_nn<D>(dk, pf, smem_u32(sQ[st]), lane);
    }
    __syncthreads();
  }
#pragma unroll
  for (int n = 0; n < D / 8; ++n) {
    const int col = h * D + 8 * n + 2 * ql;
    if (j0 < p.Sk) {
      *reinterpret_cast<uint32_t*>(p.dk + ((long long)b * p.Sk + j0) * p.lddk + col) =
          pack_bf16x2(dk[n][0] * p.scale, dk[n][1] * p.scale);
      *reinterpret_cast<uint32_t*>(p.dv + ((long long)b * p.Sk + j0) * p.lddv + col) = pack_bf16x2(dv[n][0], dv[n][1]);
    }
    if (j1 < p.Sk) {
      *reinterpret_cast<uint32_t*>(p.dk + ((long long)b * p.Sk + j1) * p.lddk + col) =
          pack_bf16x2(dk[n][2] * p.scale, dk[n][3] * p.scale);
      *reinterpret_cast<uint32_t*>(p.dv + ((long long)b * p.Sk + j1) * p.lddv + col) = pack_bf16x2(dv[n][2], dv[n][3]);
    }
  }
}

}  // namespace mmfm

// ------------------------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------------------------
using namespace mmfm;

template <void (*KERNEL)(const AttnParams)>
static int launch_k(const AttnParams& p, dim3 grid, int smem, cudaStream_t st) {
  static bool attr_set = false;
  if (!attr_set) {
    if (smem > 48 * 1024)
      MMFM_CHECK_CUDA(cudaFuncSetAttribute(KERNEL, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attr_set = true;
  }
  KERNEL<<<grid, kAttnThreads, smem, st>>>(p);
  MMFM_CHECK_CUDA(cudaGetLastError());
  return 0;
}

static int check_common(const mmfm_attn_args* a, const char* who) {
  MMFM_REQUIRE(a != nullptr, "%s: null args", who);
  MMFM_REQUIRE(a->q && a->k && a->v && a->o && a->lse && a->key_valid, "%s: null operand", who);
  MMFM_REQUIRE(a->B > 0 && a->n_heads > 0 && a->Sq > 0 && a->Sk > 0, "%s: bad shape", who);
  MMFM_REQUIRE(a->d_head == 32 || a->d_head == 64, "%s: d_head %d not supported (32 or 64)", who, a->d_head);
  MMFM_REQUIRE(a->mask_mode >= MMFM_MASK_KEY && a->mask_mode <= MMFM_MASK_CAUSAL, "%s: bad mask mode %d", who,
               a->mask_mode);
  MMFM_REQUIRE((a->mod_q == nullptr) == (a->mod_k == nullptr), "%s: mod_q and mod_k must be given together", who);
  MMFM_REQUIRE(a->ldq % 8 == 0 && a->ldk % 8 == 0 && a->ldv % 8 == 0 && a->ldo % 8 == 0,
               "%s: row pitches must be multiples of 8 elements", who);
  MMFM_REQUIRE(a->drop_p.thresh < 256u && a->drop_o.thresh < 256u, "%s: dropout threshold out of range", who);
  MMFM_REQUIRE(a->drop_p.thresh == 0u || a->drop_p.seed, "%s: probability dropout without seed", who);
  MMFM_REQUIRE(a->drop_o.thresh == 0u || a->drop_o.seed, "%s: output dropout without seed", who);
  MMFM_REQUIRE(a->B <= 65535 && a->n_heads <= 65535, "%s: grid too large", who);
  return 0;
}

static AttnParams to_params(const mmfm_attn_args* a) {
  AttnParams p;
  p.q = (const bf16*)a->q; p.ldq = a->ldq;
  p.k = (const bf16*)a->k; p.ldk = a->ldk;
  p.v = (const bf16*)a->v; p.ldv = a->ldv;
  p.o = (bf16*)a->o; p.ldo = a->ldo;
  p.lse = a->lse;
  p.key_valid = a->key_valid;
  p.mod_q = a->mod_q; p.mod_k = a->mod_k;
  p.B = a->B; p.nh = a->n_heads; p.Sq = a->Sq; p.Sk = a->Sk;
  p.mask_mode = a->mask_mode;
  p.scale = a->scale;
  p.drop_p = DropCfg{a->drop_p.seed, a->drop_p.site, a->drop_p.thresh, a->drop_p.scale};
  p.drop_o = DropCfg{a->drop_o.seed, a->drop_o.site, a->drop_o.thresh, a->drop_o.scale};
  p.p_keep = a->p_keep;
  p.d_o = (bf16*)a->d_o; p.lddo = a->lddo;
  p.delta = a->delta;
  p.dq = (bf16*)a->dq; p.lddq = a->lddq;
  p.dk = (bf16*)a->dk; p.lddk = a->lddk;
  p.dv = (bf16*)a->dv; p.lddv = a->lddv;
  return p;
}

extern "C" int mmfm_attention_fwd(const mmfm_attn_args* a, void* stream) {
  if (int rc = check_common(a, "mmfm_attention_fwd")) return rc;
  const bool drop = a->drop_p.thresh != 0u;
  MMFM_REQUIRE(!drop || a->p_keep, "mmfm_attention_fwd: probability dropout needs the p_keep buffer");
  const AttnParams p = to_params(a);
  dim3 grid((a->Sq + kTile - 1) / kTile, a->n_heads, a->B);
  cudaStream_t st = (cudaStream_t)stream;
  if (a->d_head == 32) {
    if (drop) return launch_k<attn_fwd_kernel<32, true>>(p, grid, 5 * TileCfg<32>::kBytes + 2 * kTile, st);
    return launch_k<attn_fwd_kernel<32, false>>(p, grid, 5 * TileCfg<32>::kBytes + 2 * kTile, st);
  }
  if (drop) return launch_k<attn_fwd_kernel<64, true>>(p, grid, 5 * TileCfg<64>::kBytes + 2 * kTile, st);
  return launch_k<attn_fwd_kernel<64, false>>(p, grid, 5 * TileCfg<64>::kBytes + 2 * kTile, st);
}

template <int D>
static int launch_bwd(const mmfm_attn_args* a, const AttnParams& p, cudaStream_t st) {
  const bool drop = a->drop_p.thresh != 0u;
  const long long R = (long long)a->B * a->Sq;
  int pgrid = (int)((R + 7) / 8);
  const int cap = device_sm_count() * 8;
  if (pgrid > cap) pgrid = cap;
  attn_bwd_prep_kernel<D><<<pgrid, 256, 0, st>>>(p);
  MMFM_CHECK_CUDA(cudaGetLastError());
  dim3 gq((a->Sq + kTile - 1) / kTile, a->n_heads, a->B);
  dim3 gk((a->Sk + kTile - 1) / kTile, a->n_heads, a->B);
  // six 64 x D tiles resident (+ flags / row statistics); D = 64 exceeds the 48 KB default -> opt-in dynamic smem
  constexpr int smem_dq = 6 * TileCfg<D>::kBytes + 2 * kTile;
  constexpr int smem_dkv = 6 * TileCfg<D>::kBytes + 4 * kTile * 4;
  if (drop) {
    if (int rc = launch_k<attn_bwd_dq_kernel<D, true>>(p, gq, smem_dq, st)) return rc;
    return launch_k<attn_bwd_dkv_kernel<D, true>>(p, gk, smem_dkv, st);
  }
  if (int rc = launch_k<attn_bwd_dq_kernel<D, false>>(p, gq, smem_dq, st)) return rc;
  return launch_k<attn_bwd_dkv_kernel<D, false>>(p, gk, smem_dkv, st);
}

extern "C" int mmfm_attention_bwd(const mmfm_attn_args* a, void* stream) {
  if (int rc = check_common(a, "mmfm_attention_bwd")) return rc;
  MMFM_REQUIRE(a->d_o && a->delta && a->dq && a->dk && a->dv, "mmfm_attention_bwd: null gradient buffer");
  MMFM_REQUIRE(a->lddo % 8 == 0 && a->lddq % 8 == 0 && a->lddk % 8 == 0 && a->lddv % 8 == 0,
               "mmfm_attention_bwd: row pitches must be multiples of 8 elements");
  MMFM_REQUIRE(a->drop_p.thresh == 0u || a->p_keep, "mmfm_attention_bwd: probability dropout needs p_keep");
  const AttnParams p = to_params(a);
  cudaStream_t st = (cudaStream_t)stream;
  return a->d_head == 32 ? launch_bwd<32>(a, p, st) : launch_bwd<64>(a, p, st);
}
