// Micro-benchmark: tcgen05.mma (kind::f16, M = 128, K = 16) issue throughput from resident shared-memory operands.
// Varies N, the operand layout (SW128 vs no-swizzle), whether consecutive MMAs accumulate into the same TMEM columns,
// and whether the A operand is re-read at a shifted row (the conv-tap pattern).  One CTA per SM, one issuing thread.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I audiolm_pytorch_b200/csrc tools/mma_bench.cu -o /tmp/mma_bench
#include <cstdio>
#include <cuda_runtime.h>
#include "ptx_sm100.cuh"

using namespace alm;

template <int N, int NACC, int SW, int TS>
__global__ void __launch_bounds__(128, 1) bench(int iters, long long* cycles) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  for (int i = threadIdx.x; i < 96 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    fence_mbar_init();
  }
  if (threadIdx.x < 32) tmem_alloc(&slot, 512);
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = slot;
  if (threadIdx.x == 0) {
    constexpr uint32_t idesc = umma_idesc_bf16_f32(128, N, false, false);
    const uint32_t a = smem_u32(smem), b = smem_u32(smem + 48 * 1024);
    // descriptors prepared once; only the start-address field (low word) changes between MMAs
    const uint64_t da0 = SW ? umma_smem_desc_sw128(a, 1024, 0) : umma_smem_desc_nosw(a, 128, 184 * 16);
    const uint64_t db0 = SW ? umma_smem_desc_sw128(b, 1024, 0) : umma_smem_desc_nosw(b, 128, N * 16);
    long long t0 = clock64();
    for (int it = 0; it < iters; it += 8) {
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const uint32_t d = tmem + (k % NACC) * N;
        const uint64_t da = da0 + (uint64_t)(SW ? ((k % 7) * 8 + (k & 3) * 2) : (k % 7));  // (>>4 units): row / k shifts
        const uint64_t db = db0 + (uint64_t)(SW ? (k & 3) * 2 : 0);
        if (TS) umma_bf16_ts(d, tmem + 256 + (k & 7) * 8, db, idesc, 1u);
        else umma_bf16_ss(d, da, db, idesc, 1u);
      }
    }
    umma_commit(&bar);
    mbar_wait(&bar, 0);
    long long t1 = clock64();
    if (blockIdx.x == 0) *cycles = t1 - t0;
  }
  tc_fence_before_sync();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tmem, 512);
}

template <int N, int NACC, int SW, int TS>
void run() {
  long long* d;
  cudaMalloc(&d, 8);
  const int iters = 40000;
  auto k = bench<N, NACC, SW, TS>;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  k<<<148, 128, 100 * 1024>>>(iters, d);
  k<<<148, 128, 100 * 1024>>>(iters, d);
  long long c = 0;
  cudaError_t e = cudaDeviceSynchronize();
  cudaMemcpy(&c, d, 8, cudaMemcpyDeviceToHost);
  printf("N=%3d %-9s acc=%d %s : %7.1f clk/MMA (floor %d)%s\n", N, SW ? "sw128" : "noswizzle", NACC,
         TS ? "A=tmem" : "A=smem", (double)c / iters, N / 2, e == cudaSuccess ? "" : cudaGetErrorString(e));
  cudaFree(d);
}

template <int SW, int TS>
void sweep() {
  run<32, 1, SW, TS>();  run<64, 1, SW, TS>();  run<128, 1, SW, TS>();  run<256, 1, SW, TS>();
  run<32, 2, SW, TS>();  run<64, 2, SW, TS>();  run<128, 2, SW, TS>();
}

int main() {
  sweep<0, 0>();
  sweep<1, 0>();
  sweep<1, 1>();
  return 0;
}
