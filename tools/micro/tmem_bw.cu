// Micro-benchmark: tcgen05.ld throughput per SM as a function of the number of reading warps.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tmem_bw tmem_bw.cu && ./tmem_bw
#include "../../multi_modal_foundation_model_b200/csrc/common.cuh"
#include <cstdio>
using namespace mmfm;

template <int X>
__global__ void k(int iters, long long* out, float* sink) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) { tmem_alloc(&slot, 512u); tmem_relinquish(); }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t base = slot + ((uint32_t)((warp & 3) * 32) << 16);
  float acc = 0.f;
  __syncthreads();
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      if (X == 32) { uint32_t r[32]; tmem_ld32(base + 32u * ((c + warp) & 7), r); tmem_ld_wait(); acc += __uint_as_float(r[0]) + __uint_as_float(r[31]); }
      else { uint32_t r[16]; tmem_ld16(base + 32u * ((c + warp) & 7), r); tmem_ld_wait(); acc += __uint_as_float(r[0]) + __uint_as_float(r[15]); }
    }
  }
  __syncthreads();
  long long t1 = clock64();
  if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
  sink[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  __syncthreads();
  if (warp == 0) tmem_dealloc(slot, 512u);
}

int main() {
  long long* out; float* sink;
  cudaMalloc(&out, 8 * 148); cudaMalloc(&sink, 4 * 148 * 1024);
  const int iters = 2000;
  for (int x : {32, 16}) for (int nw : {1, 2, 4, 8, 16}) {
    if (x == 32) k<32><<<1, nw * 32>>>(iters, out, sink); else k<16><<<1, nw * 32>>>(iters, out, sink);
    cudaError_t e = cudaDeviceSynchronize();
    long long c; cudaMemcpy(&c, out, 8, cudaMemcpyDeviceToHost);
    double bytes = (double)iters * 8 * nw * 32 * x * 4;
    printf("x%d warps %2d: %lld cycles, %.1f B/clk/SM, %.1f cyc per ld (%s)\n", x, nw, c, bytes / c, (double)c / (iters * 8), cudaGetErrorString(e));
  }
  return 0;
}
