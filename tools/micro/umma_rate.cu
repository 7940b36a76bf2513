// Micro-benchmark: cycles per tcgen05.mma (cta_group::1, kind::f16, M = 128) for the operand shapes the attention
// backward issues -- how much the N = 32 / MN-major products cost next to the N = 128 score products.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma_rate umma_rate.cu && ./umma_rate
#include "../../multi_modal_foundation_model_b200/csrc/common.cuh"
#include <cstdio>
using namespace mmfm;

// kind: 0 = S (A K-major 64B swizzle, B K-major 64B swizzle, N = n, K-steps of 32 B)
//       1 = dV / dK (A MN-major 128B-swizzle slab, B MN-major 64B swizzle, N = 32)
//       2 = dQ (A K-major 128B-swizzle slab, B MN-major 64B swizzle, N = 32)
//       3 = P.V forward-style (A from TMEM, B MN-major 64B swizzle, N = 32)
__global__ void k(int kind, int n, int reps, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t slot;
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 40960; i += blockDim.x) reinterpret_cast<uint32_t*>(smem_raw + (base - smem_u32(smem_raw)))[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  if (warp == 0) { tmem_alloc(&slot, 512u); tmem_relinquish(); }
  fence_proxy_async();
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tm = slot;
  if (warp == 0 && elect_one()) {
    const uint32_t slab = base, opnd = base + 65536;
    long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
      if (kind == 0) {
        const uint32_t idesc = make_idesc_bf16(128, n, 0, 0);
        for (int kk = 0; kk < 2; ++kk)
          umma_bf16(tm, make_smem_desc(opnd + kk * 32, 16, 512, 4), make_smem_desc(opnd + 16384 + kk * 32, 16, 512, 4), idesc, 1u);
      } else if (kind == 1) {
        const uint32_t idesc = make_idesc_bf16(128, n, 1, 1);
        for (int kk = 0; kk < 8; ++kk)
          umma_bf16(tm + 256, make_smem_desc(slab + kk * 2048, 16384, 1024, 2), make_smem_desc(opnd + kk * 1024, 512, 512, 4), idesc, 1u);
      } else if (kind == 2) {
        const uint32_t idesc = make_idesc_bf16(128, n, 0, 1);
        for (int kk = 0; kk < 8; ++kk)
          umma_bf16(tm + 256, make_smem_desc(slab + (kk >> 2) * 16384 + (kk & 3) * 32, 16, 1024, 2), make_smem_desc(opnd + kk * 1024, 512, 512, 4), idesc, 1u);
      } else {
        const uint32_t idesc = make_idesc_bf16(128, n, 0, 1);
        for (int kk = 0; kk < 8; ++kk)
          umma_bf16_ts(tm + 256, tm + 8 * kk, make_smem_desc(opnd + kk * 1024, 512, 512, 4), idesc, 1u);
      }
    }
    umma_commit(&bar);
    mbar_wait(&bar, 0);
    long long t1 = clock64();
    out[0] = t1 - t0;
  }
  __syncthreads();
  tc_fence_before(); __syncthreads();
  if (warp == 0) tmem_dealloc(tm, 512u);
}

int main() {
  long long* out;
  cudaMalloc(&out, 64);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  const char* names[4] = {"S  (K-major x K-major)", "dV (MN-major slab A, MN-major B)", "dQ (K-major slab A, MN-major B)", "PV (TMEM A, MN-major B)"};
  for (int kind = 0; kind < 4; ++kind)
    for (int n : {32, 64, 128, 208}) {
      if (kind == 0 && n < 64) continue;
      if (kind != 0 && n > 64) continue;
      const int reps = 64, per = kind == 0 ? 2 : 8;
      k<<<1, 128, 200 * 1024>>>(kind, n, reps, out);
      cudaError_t e = cudaDeviceSynchronize();
      long long c; cudaMemcpy(&c, out, 8, cudaMemcpyDeviceToHost);
      printf("%-36s N %3d: %6.1f cycles per MMA (%s)\n", names[kind], n, (double)c / (reps * per), cudaGetErrorString(e));
    }
  return 0;
}
