// Microbenchmark: cycles per tcgen05.mma (cta_group::1, kind::f16, M=128, K=16, SS mode, SWIZZLE_128B K-major operands)
// as a function of N, issued back to back by one thread into one TMEM accumulator, operands resident in shared memory.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I ccdm_b200/csrc -I include -o gpurun_out/mma_rate tools/ubench/mma_rate.cu
#include <cstdio>
#include <cuda_runtime.h>
#include <cuda.h>
#include "ptx.cuh"
using namespace ccdm;

__global__ void __launch_bounds__(128, 1) mma_rate_kernel(int N, int iters, int a_shared_rows, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < (16384 + 256 * 128) / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3C003C00u;
  if (warp == 0) tmem_alloc(&tmem_slot, 256);
  if (tid == 32) { mbar_init(&bar, 1); fence_mbar_init(); }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  if (warp == 1) {
    const uint32_t idesc = umma_idesc_bf16(128, N);
    const uint32_t a16 = (smem_u32(smem) & 0x3FFFF) >> 4, b16 = a16 + (16384 >> 4);
    long long t0 = 0, t1 = 0;
    if (elect_one()) {
      t0 = clock64();
      for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_bf16_ss(tmem, umma_desc_sw128_a16(a16 + 2 * k), umma_desc_sw128_a16(b16 + 2 * k), idesc, 1u);
      }
      umma_commit(&bar);
    }
    __syncwarp();
    mbar_wait(&bar, 0);
    t1 = clock64();
    if (elect_one() && blockIdx.x == 0) out[0] = t1 - t0;
    __syncwarp();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem, 256); }
}

int main() {
  long long* d; cudaMalloc(&d, 8);
  const size_t smem = 16384 + 256 * 128 + 1024;
  cudaFuncSetAttribute(mma_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  const int iters = 2000;
  for (int all = 0; all < 2; ++all)
    for (int N : {16, 32, 64, 128, 256}) {
      const int grid = all ? 148 : 1;
      mma_rate_kernel<<<grid, 128, smem>>>(N, iters, 0, d);
      cudaDeviceSynchronize();
      mma_rate_kernel<<<grid, 128, smem>>>(N, iters, 0, d);
      cudaError_t e = cudaDeviceSynchronize();
      long long c = 0; cudaMemcpy(&c, d, 8, cudaMemcpyDeviceToHost);
      const double per = (double)c / (iters * 4.0);
      printf("{\"grid\": %d, \"M\": 128, \"N\": %d, \"K\": 16, \"cycles_per_mma\": %.1f, \"tensor_floor_cycles\": %.1f, \"smem_operand_bytes\": %d, \"bytes_per_cycle\": %.1f, \"err\": \"%s\"}\n",
             grid, N, per, 128.0 * N / 256.0, 4096 + N * 32, (4096 + N * 32) / per, cudaGetErrorString(e));
    }
  return 0;
}
