// tcgen05 / TMEM / TMA GEMM kernels of the multi-modal encoder/decoder path.
//
//   gemm_tn_kernel    D[M,N] = epilogue(A[M,K] . B[N,K]^T)     both operands K-major (nn.Linear forward + dgrad)
//   gemm_wgrad_kernel dW[NO,KI] += dY[R,NO]^T . X[R,KI]        both operands MN-major, split over R (wgrad)
//
// Replaces the cuBLAS calls under every nn.Linear of the reference path (SURVEY.md 2.2 K1,K3,K10,K12,K14-K17 and
// their autograd backward, K21).  One CTA computes one 128 x BN output tile:
//   warp 0   : TMA producer (one elected lane), STAGES-deep mbarrier ring of 128B-swizzled smem tiles
//   warp 1   : TMEM allocation + tcgen05.mma issue (one elected lane), fp32 accumulator in TMEM
//   warps 2-5: epilogue -- tcgen05.ld of their TMEM lane quadrant, fused bias / activation / dropout /
//              token-zeroing / residual, direct vectorised global stores
// Several CTAs are resident per SM (smem- and TMEM-limited), so one CTA's epilogue overlaps another's main loop.
#include "common.cuh"
#include "host_util.h"
#include "../../include/mmfm_b200.h"

namespace mmfm {

constexpr int kBM = 128;
constexpr int kBK = 64;
constexpr int kGemmThreads = 192;

// ------------------------------------------------------------------------------------------------
// epilogue math on 16 consecutive columns of one row
// ------------------------------------------------------------------------------------------------
struct EpiRow {
  bool valid;         // row < M
  bool zero;          // token zeroing flag
  long long out_row;  // output row after remap
  long long row;      // GEMM row (dropout field row, aux row)
};

MMFM_DEVINL void load16_f32(const float* p, bool vec, int nvalid, float (&o)[16]) {
  if (vec && nvalid == 16) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float4 t = __ldg(reinterpret_cast<const float4*>(p) + j);
      o[4 * j] = t.x; o[4 * j + 1] = t.y; o[4 * j + 2] = t.z; o[4 * j + 3] = t.w;
    }
  } else {
#pragma unroll
    for (int j = 0; j < 16; ++j) o[j] = (j < nvalid) ? __ldg(p + j) : 0.f;
  }
}
MMFM_DEVINL void load16_bf16(const bf16* p, bool vec, int nvalid, float (&o)[16]) {
  if (vec && nvalid == 16) {
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      uint4 t = __ldg(reinterpret_cast<const uint4*>(p) + j);
      float2 a = unpack_bf16x2(t.x), b = unpack_bf16x2(t.y), c = unpack_bf16x2(t.z), d = unpack_bf16x2(t.w);
      o[8 * j] = a.x; o[8 * j + 1] = a.y; o[8 * j + 2] = b.x; o[8 * j + 3] = b.y;
      o[8 * j + 4] = c.x; o[8 * j + 5] = c.y; o[8 * j + 6] = d.x; o[8 * j + 7] = d.y;
    }
  } else {
#pragma unroll
    for (int j = 0; j < 16; ++j) o[j] = (j < nvalid) ? __bfloat162float(p[j]) : 0.f;
  }
}
MMFM_DEVINL void store16_f32(float* p, bool vec, int nvalid, const float (&v)[16]) {
  if (vec && nvalid == 16) {
#pragma unroll
    for (int j = 0; j < 4; ++j)
      reinterpret_cast<float4*>(p)[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
  } else {
#pragma unroll
    for (int j = 0; j < 16; ++j)
      if (j < nvalid) p[j] = v[j];
  }
}
MMFM_DEVINL void store16_bf16(bf16* p, bool vec, int nvalid, const float (&v)[16]) {
  if (vec && nvalid == 16) {
#pragma unroll
    for (int j = 0; j < 2; ++j)
      reinterpret_cast<uint4*>(p)[j] =
          make_uint4(pack_bf16x2(v[8 * j], v[8 * j + 1]), pack_bf16x2(v[8 * j + 2], v[8 * j + 3]),
                     pack_bf16x2(v[8 * j + 4], v[8 * j + 5]), pack_bf16x2(v[8 * j + 6], v[8 * j + 7]));
  } else {
#pragma unroll
    for (int j = 0; j < 16; ++j)
      if (j < nvalid) p[j] = __float2bfloat16_rn(v[j]);
  }
}

MMFM_DEVINL void epilogue16(const mmfm_gemm_args& p, const EpiRow& er, int n, float (&v)[16], unsigned long long seed,
                            uint32_t drop_gpr) {
  const int nvalid = min(16, p.N - n);
  if (p.bias) {
    float b[16];
    load16_f32(p.bias + n, (reinterpret_cast<uintptr_t>(p.bias + n) & 15) == 0, nvalid, b);
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] += b[j];
  }
  if (p.act == MMFM_ACT_GELU) {
    if (p.D2) {
      bf16* d2 = reinterpret_cast<bf16*>(p.D2) + er.out_row * p.ldd + n;
      store16_bf16(d2, (reinterpret_cast<uintptr_t>(d2) & 15) == 0, nvalid, v);
    }
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] = gelu_erf(v[j]);
  } else if (p.act == MMFM_ACT_SOFTSIGN) {
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] = softsign(v[j]) * p.act_scale;
  } else if (p.act == MMFM_ACT_DGELU || p.act == MMFM_ACT_DSOFTSIGN) {
    float a[16];
    const bf16* ap = reinterpret_cast<const bf16*>(p.aux) + er.row * p.ldaux + n;
    load16_bf16(ap, (reinterpret_cast<uintptr_t>(ap) & 15) == 0, nvalid, a);
    if (p.act == MMFM_ACT_DGELU) {
#pragma unroll
      for (int j = 0; j < 16; ++j) v[j] *= gelu_erf_grad(a[j]);
    } else {
      const float inv = 1.0f / p.act_scale;
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        float t = 1.0f - fabsf(a[j] * inv);
        v[j] *= p.act_scale * t * t;
      }
    }
  }
  if (p.drop.thresh != 0u) {
    uint4 w = drop_bytes16(seed, p.drop.site, (uint64_t)er.row, drop_gpr, (uint32_t)(n >> 4));
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] = (drop_byte(w, j) < p.drop.thresh) ? 0.f : v[j] * p.drop.scale;
  }
  if (er.zero) {
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] = 0.f;
  }
  if (p.res) {
    float r[16];
    const float* rp = p.res + er.out_row * p.ldr + n;
    load16_f32(rp, (reinterpret_cast<uintptr_t>(rp) & 15) == 0, nvalid, r);
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] += r[j];
  }
  if (p.d_fp32) {
    float* dp = reinterpret_cast<float*>(p.D) + er.out_row * p.ldd + n;
    store16_f32(dp, (reinterpret_cast<uintptr_t>(dp) & 15) == 0, nvalid, v);
  } else {
    bf16* dp = reinterpret_cast<bf16*>(p.D) + er.out_row * p.ldd + n;
    store16_bf16(dp, (reinterpret_cast<uintptr_t>(dp) & 15) == 0, nvalid, v);
  }
}

// ------------------------------------------------------------------------------------------------
// TN kernel
// ------------------------------------------------------------------------------------------------
template <int BN, int STAGES>
__global__ void __launch_bounds__(kGemmThreads) gemm_tn_kernel(const __grid_constant__ CUtensorMap tmA,
                                                                const __grid_constant__ CUtensorMap tmB,
                                                                const mmfm_gemm_args p) {
  constexpr uint32_t kABytes = kBM * kBK * 2;  // 16 KB
  constexpr uint32_t kBBytes = BN * kBK * 2;
  constexpr uint32_t kStageBytes = kABytes + kBBytes;
  constexpr uint32_t kTmemCols = BN < 32 ? 32 : BN;

  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[STAGES];
  __shared__ __align__(8) uint64_t empty_bar[STAGES];
  __shared__ __align__(8) uint64_t accum_bar;
  __shared__ uint32_t tmem_slot;

  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int n0 = blockIdx.x * BN;
  const int m0 = blockIdx.y * kBM;
  const int nkb = (p.K + kBK - 1) / kBK;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
#pragma unroll
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(&accum_bar, 1);
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(&tmem_slot, kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;

  if (warp == 0) {
    if (elect_one()) {
      for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb % STAGES;
        if (kb >= STAGES) mbar_wait(&empty_bar[s], ((kb / STAGES) - 1) & 1);
        mbar_arrive_expect_tx(&full_bar[s], kStageBytes);
        const uint32_t a_dst = smem_base + s * kStageBytes;
        tma_load_2d_addr(a_dst, &tmA, &full_bar[s], kb * kBK, m0);
        tma_load_2d_addr(a_dst + kABytes, &tmB, &full_bar[s], kb * kBK, n0);
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      const uint32_t idesc = make_idesc_bf16(kBM, BN, 0, 0);
      for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb % STAGES;
        mbar_wait(&full_bar[s], (kb / STAGES) & 1);
        tc_fence_after();
        const uint32_t a_addr = smem_base + s * kStageBytes;
        const uint32_t b_addr = a_addr + kABytes;
#pragma unroll
        for (int k = 0; k < kBK / 16; ++k) {
          const uint64_t da = make_smem_desc(a_addr + k * 32, 16, 1024, 2);
          const uint64_t db = make_smem_desc(b_addr + k * 32, 16, 1024, 2);
          umma_bf16(tmem_base, da, db, idesc, (kb > 0 || k > 0) ? 1u : 0u);
        }
        umma_commit(&empty_bar[s]);
      }
      umma_commit(&accum_bar);
    }
  } else {
    // epilogue warps 2..5 -> TMEM lane quadrants (warp % 4)
    const int quad = warp & 3;
    const long long r = (long long)m0 + quad * 32 + lane;
    EpiRow er;
    er.valid = r < p.M;
    er.row = r;
    er.out_row = r;
    er.zero = false;
    if (er.valid) {
      if (p.remap_T > 0) {
        const long long b = r / p.remap_T;
        const int t = (int)(r - b * p.remap_T);
        er.out_row = b * p.remap_S + p.remap_off + t;
        if (p.row_zero) er.zero = p.row_zero[p.remap_off + t] != 0;
      } else if (p.row_zero) {
        er.zero = p.row_zero[(int)(r % p.remap_S)] != 0;
      }
    }
    unsigned long long seed = 0ull;
    if (p.drop.thresh != 0u) seed = *p.drop.seed;
    const uint32_t drop_gpr = (uint32_t)((p.N + 15) >> 4);

    mbar_wait(&accum_bar, 0);
    tc_fence_after();
    const uint32_t t_row = tmem_base + ((uint32_t)(quad * 32) << 16);
#pragma unroll 1
    for (int c0 = 0; c0 < BN; c0 += 32) {
      if (n0 + c0 >= p.N) break;  // warp-uniform
      uint32_t acc[32];
      tmem_ld32(t_row + (uint32_t)c0, acc);
      tmem_ld_wait();
      if (er.valid) {
#pragma unroll
        for (int g = 0; g < 2; ++g) {
          const int n = n0 + c0 + g * 16;
          if (n < p.N) {
            float v[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(acc[g * 16 + j]);
            epilogue16(p, er, n, v, seed, drop_gpr);
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, kTmemCols);
}

// ------------------------------------------------------------------------------------------------
// wgrad kernel: dW[NO,KI] += sum over a slice of rows of dY[r,NO]^T X[r,KI]
// ------------------------------------------------------------------------------------------------
template <int STAGES>
__global__ void __launch_bounds__(kGemmThreads) gemm_wgrad_kernel(const __grid_constant__ CUtensorMap tmY,
                                                                   const __grid_constant__ CUtensorMap tmX, int R,
                                                                   int NO, int KI, float* __restrict__ dW,
                                                                   long long ldw, int rows_per_split) {
  constexpr int BN = 128;
  constexpr uint32_t kBoxBytes = 64 * kBK * 2;  // [64 rows(k) x 64 cols(mn)] bf16 = 8 KB
  constexpr uint32_t kABytes = 2 * kBoxBytes;
  constexpr uint32_t kBBytes = 2 * kBoxBytes;
  constexpr uint32_t kStageBytes = kABytes + kBBytes;
  constexpr uint32_t kTmemCols = BN;

  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[STAGES];
  __shared__ __align__(8) uint64_t empty_bar[STAGES];
  __shared__ __align__(8) uint64_t accum_bar;
  __shared__ uint32_t tmem_slot;

  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int ki0 = blockIdx.x * BN;
  const int no0 = blockIdx.y * kBM;
  const int r_begin = blockIdx.z * rows_per_split;
  const int r_end = min(R, r_begin + rows_per_split);
  const int nkb = (r_end - r_begin + kBK - 1) / kBK;  // >= 1 by construction of the grid

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmY);
    tma_prefetch_desc(&tmX);
#pragma unroll
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(&accum_bar, 1);
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(&tmem_slot, kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;

  if (warp == 0) {
    if (elect_one()) {
      for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb % STAGES;
        if (kb >= STAGES) mbar_wait(&empty_bar[s], ((kb / STAGES) - 1) & 1);
        mbar_arrive_expect_tx(&full_bar[s], kStageBytes);
        const uint32_t a_dst = smem_base + s * kStageBytes;
        const int r0 = r_begin + kb * kBK;
        // NOTE: rows beyond r_end (but < R) of the last k-block belong to the next split; they are excluded by
        // making rows_per_split a multiple of kBK on the host, so only the global tail (>= R) is zero-filled.
        tma_load_2d_addr(a_dst, &tmY, &full_bar[s], no0, r0);
        tma_load_2d_addr(a_dst + kBoxBytes, &tmY, &full_bar[s], no0 + 64, r0);
        tma_load_2d_addr(a_dst + kABytes, &tmX, &full_bar[s], ki0, r0);
        tma_load_2d_addr(a_dst + kABytes + kBoxBytes, &tmX, &full_bar[s], ki0 + 64, r0);
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      const uint32_t idesc = make_idesc_bf16(kBM, BN, 1, 1);
      for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb % STAGES;
        mbar_wait(&full_bar[s], (kb / STAGES) & 1);
        tc_fence_after();
        const uint32_t a_addr = smem_base + s * kStageBytes;
        const uint32_t b_addr = a_addr + kABytes;
#pragma unroll
        for (int k = 0; k < kBK / 16; ++k) {
          // MN-major, 128B swizzle: 64 mn-elements per 128B row, 8 k-rows per 1024B atom (SBO), next 64
          // mn-elements in the neighbouring TMA box (LBO = box size)
          const uint64_t da = make_smem_desc(a_addr + k * 2048, kBoxBytes, 1024, 2);
          const uint64_t db = make_smem_desc(b_addr + k * 2048, kBoxBytes, 1024, 2);
          umma_bf16(tmem_base, da, db, idesc, (kb > 0 || k > 0) ? 1u : 0u);
        }
        umma_commit(&empty_bar[s]);
      }
      umma_commit(&accum_bar);
    }
  } else {
    const int quad = warp & 3;
    const int no = no0 + quad * 32 + lane;
    mbar_wait(&accum_bar, 0);
    tc_fence_after();
    const uint32_t t_row = tmem_base + ((uint32_t)(quad * 32) << 16);
#pragma unroll 1
    for (int c0 = 0; c0 < BN; c0 += 32) {
      if (ki0 + c0 >= KI) break;
      uint32_t acc[32];
      tmem_ld32(t_row + (uint32_t)c0, acc);
      tmem_ld_wait();
      if (no < NO) {
        float* dst = dW + (long long)no * ldw + ki0 + c0;
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (ki0 + c0 + j < KI) atomicAdd(dst + j, __uint_as_float(acc[j]));
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, kTmemCols);
}

// ------------------------------------------------------------------------------------------------
// bias gradient: out[c] += sum_r dY[r,c]
// ------------------------------------------------------------------------------------------------
constexpr int kColsumRows = 256;
__global__ void __launch_bounds__(256) colsum_bf16_kernel(const bf16* __restrict__ dY, long long ld, int R, int NO,
                                                           float* __restrict__ out) {
  const int c = (blockIdx.x * 256 + threadIdx.x) * 2;
  if (c >= NO) return;
  const int r0 = blockIdx.y * kColsumRows;
  const int r1 = min(R, r0 + kColsumRows);
  float s0 = 0.f, s1 = 0.f;
  if (c + 1 < NO && (ld & 1) == 0) {
    for (int r = r0; r < r1; ++r) {
      uint32_t u = __ldg(reinterpret_cast<const uint32_t*>(dY + (long long)r * ld + c));
      float2 f = unpack_bf16x2(u);
      s0 += f.x;
      s1 += f.y;
    }
  } else {
    for (int r = r0; r < r1; ++r) {
      s0 += __bfloat162float(dY[(long long)r * ld + c]);
      if (c + 1 < NO) s1 += __bfloat162float(dY[(long long)r * ld + c + 1]);
    }
  }
  atomicAdd(out + c, s0);
  if (c + 1 < NO) atomicAdd(out + c + 1, s1);
}

// ------------------------------------------------------------------------------------------------
// fp32 -> bf16 cast with optional transposed copy (weight shadows, input staging)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) cast_bf16_kernel(const float* __restrict__ x, long long ldx, bf16* __restrict__ y,
                                                         long long ldy, bf16* __restrict__ yt, long long ldyt, int R,
                                                         int C) {
  __shared__ float tile[32][33];
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = r0 + ty + i * 8, c = c0 + tx;
    float v = 0.f;
    if (r < R && c < C) {
      v = x[(long long)r * ldx + c];
      if (y) y[(long long)r * ldy + c] = __float2bfloat16_rn(v);
    }
    tile[ty + i * 8][tx] = v;
  }
  if (yt == nullptr) return;
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int c = c0 + ty + i * 8, r = r0 + tx;
    if (r < R && c < C) yt[(long long)c * ldyt + r] = __float2bfloat16_rn(tile[tx][ty + i * 8]);
  }
}

}  // namespace mmfm

// ------------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------------
using namespace mmfm;

template <int BN, int STAGES>
static int launch_tn(const mmfm_gemm_args* a, cudaStream_t st) {
  CUtensorMap tmA, tmB;
  int rc = make_tmap_bf16_2d(&tmA, a->A, (uint64_t)a->M, (uint64_t)a->K, (uint64_t)a->lda, kBK, kBM, TMA_SW_128);
  if (rc) return rc;
  rc = make_tmap_bf16_2d(&tmB, a->B, (uint64_t)a->N, (uint64_t)a->K, (uint64_t)a->ldb, kBK, BN, TMA_SW_128);
  if (rc) return rc;
  constexpr size_t smem = (size_t)STAGES * (kBM * kBK * 2 + BN * kBK * 2) + 1024;
  static bool attr_set = false;
  if (!attr_set) {
    MMFM_CHECK_CUDA(cudaFuncSetAttribute(gemm_tn_kernel<BN, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)smem));
    attr_set = true;
  }
  dim3 grid((a->N + BN - 1) / BN, (a->M + kBM - 1) / kBM, 1);
  gemm_tn_kernel<BN, STAGES><<<grid, kGemmThreads, smem, st>>>(tmA, tmB, *a);
  MMFM_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int mmfm_gemm_tn(const mmfm_gemm_args* a, void* stream) {
  MMFM_REQUIRE(a != nullptr, "mmfm_gemm_tn: null args");
  MMFM_REQUIRE(a->M > 0 && a->N > 0 && a->K > 0, "mmfm_gemm_tn: bad shape M=%d N=%d K=%d", a->M, a->N, a->K);
  MMFM_REQUIRE(a->A && a->B && a->D, "mmfm_gemm_tn: null operand");
  MMFM_REQUIRE(a->act >= MMFM_ACT_NONE && a->act <= MMFM_ACT_DSOFTSIGN, "mmfm_gemm_tn: bad act %d", a->act);
  MMFM_REQUIRE(!(a->act >= MMFM_ACT_DGELU) || a->aux, "mmfm_gemm_tn: act %d needs aux", a->act);
  MMFM_REQUIRE(a->drop.thresh == 0 || a->drop.seed, "mmfm_gemm_tn: dropout without seed pointer");
  MMFM_REQUIRE(a->drop.thresh < 256, "mmfm_gemm_tn: dropout threshold out of range");
  MMFM_REQUIRE(!(a->row_zero && a->remap_T == 0) || a->remap_S > 0, "mmfm_gemm_tn: row_zero needs remap_S");
  cudaStream_t st = (cudaStream_t)stream;
  if (a->N <= 64) return launch_tn<64, 4>(a, st);
  return launch_tn<128, 3>(a, st);
}

extern "C" int mmfm_gemm_wgrad(const void* dY, long long lddy, const void* X, long long ldx, int R, int NO, int KI,
                               float* dW, long long ldw, void* stream) {
  MMFM_REQUIRE(dY && X && dW, "mmfm_gemm_wgrad: null operand");
  MMFM_REQUIRE(R > 0 && NO > 0 && KI > 0, "mmfm_gemm_wgrad: bad shape R=%d NO=%d KI=%d", R, NO, KI);
  constexpr int STAGES = 4;
  CUtensorMap tmY, tmX;
  int rc = make_tmap_bf16_2d(&tmY, dY, (uint64_t)R, (uint64_t)NO, (uint64_t)lddy, 64, kBK, TMA_SW_128);
  if (rc) return rc;
  rc = make_tmap_bf16_2d(&tmX, X, (uint64_t)R, (uint64_t)KI, (uint64_t)ldx, 64, kBK, TMA_SW_128);
  if (rc) return rc;
  constexpr size_t smem = (size_t)STAGES * (4 * 64 * kBK * 2) + 1024;
  static bool attr_set = false;
  if (!attr_set) {
    MMFM_CHECK_CUDA(cudaFuncSetAttribute(gemm_wgrad_kernel<STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)smem));
    attr_set = true;
  }
  const int tiles = ((NO + kBM - 1) / kBM) * ((KI + 127) / 128);
  const int kblocks = (R + kBK - 1) / kBK;
  int splits = (2 * device_sm_count() + tiles - 1) / tiles;
  if (splits > kblocks) splits = kblocks;
  if (splits < 1) splits = 1;
  int rows_per_split = ((kblocks + splits - 1) / splits) * kBK;
  splits = (R + rows_per_split - 1) / rows_per_split;
  dim3 grid((KI + 127) / 128, (NO + kBM - 1) / kBM, splits);
  gemm_wgrad_kernel<STAGES><<<grid, kGemmThreads, smem, (cudaStream_t)stream>>>(tmY, tmX, R, NO, KI, dW, ldw,
                                                                                 rows_per_split);
  MMFM_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int mmfm_colsum_bf16(const void* dY, long long ld, int R, int NO, float* out, void* stream) {
  MMFM_REQUIRE(dY && out && R > 0 && NO > 0, "mmfm_colsum_bf16: bad arguments");
  dim3 grid((NO + 511) / 512, (R + kColsumRows - 1) / kColsumRows);
  colsum_bf16_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((const bf16*)dY, ld, R, NO, out);
  MMFM_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int mmfm_cast_bf16(const float* x, long long ldx, void* y, long long ldy, void* yt, long long ldyt, int R,
                              int C, void* stream) {
  MMFM_REQUIRE(x && (y || yt) && R > 0 && C > 0, "mmfm_cast_bf16: bad arguments");
  dim3 grid((C + 31) / 32, (R + 31) / 32);
  MMFM_REQUIRE(grid.y <= 65535, "mmfm_cast_bf16: too many rows (%d)", R);
  cast_bf16_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(x, ldx, (bf16*)y, ldy, (bf16*)yt, ldyt, R, C);
  MMFM_CHECK_CUDA(cudaGetLastError());
  return 0;
}
