// Microbenchmark: cycles per tcgen05.mma (cta_group::1, kind::f16, M=128, K=16, SS mode, SWIZZLE_128B K-major operands)
// as a function of N, issued back to back by one thread into one TMEM accumulator, operands resident in shared memory.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I ccdm_b200/csrc -I include -o gpurun_out/mma_rate tools/ubench/mma_rate.cu
#include <cstdio>
#include <cuda_runtime.h>
#include <cuda.h>
#include "ptx.cuh"
using namespace ccdm;

// mode bit 0: A is MN-major (the weight-gradient kernels: positions are the K axis, channels contiguous), bit 1: B is MN-major.
// MN-major operands: 64-element blocks LBO = 16 KiB apart, 8-row K groups SBO = 1024 B apart, a K step = 2048 B.
__global__ void __launch_bounds__(128, 1) mma_rate_kernel(int N, int iters, int mode, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < (32768 + 65536) / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3C003C00u;
  if (warp == 0) tmem_alloc(&tmem_slot, 256);
  if (tid == 32) { mbar_init(&bar, 1); fence_mbar_init(); }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  if (warp == 1) {
    const uint32_t idesc = umma_idesc_bf16(128, N) | ((mode & 1) << 15) | (((mode >> 1) & 1) << 16);
    const uint32_t a16 = (smem_u32(smem) & 0x3FFFF) >> 4, b16 = a16 + (32768 >> 4);
    const uint64_t mn_hi = (static_cast<uint64_t>(1024 >> 4) << 32) | (static_cast<uint64_t>(1) << 46) |
                           (static_cast<uint64_t>(2) << 61);
    const uint32_t mn_lbo = static_cast<uint32_t>((16384 >> 4) & 0x3FFF) << 16;
    long long t0 = 0, t1 = 0;
    if (elect_one()) {
      t0 = clock64();
      for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_bf16_ss(tmem, (mode & 1) ? (mn_hi | (a16 + 128 * k) | mn_lbo) : umma_desc_sw128_a16(a16 + 2 * k),
                       (mode & 2) ? (mn_hi | (b16 + 128 * k) | mn_lbo) : umma_desc_sw128_a16(b16 + 2 * k), idesc, 1u);
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
  const size_t smem = 32768 + 65536 + 1024;
  cudaFuncSetAttribute(mma_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  const int iters = 2000;
  for (int mode = 0; mode < 4; ++mode)
    for (int N : {64, 80, 96, 128, 256}) {
      const int grid = 148;
      mma_rate_kernel<<<grid, 128, smem>>>(N, iters, mode, d);
      cudaDeviceSynchronize();
      mma_rate_kernel<<<grid, 128, smem>>>(N, iters, mode, d);
      cudaError_t e = cudaDeviceSynchronize();
      long long c = 0; cudaMemcpy(&c, d, 8, cudaMemcpyDeviceToHost);
      const double per = (double)c / (iters * 4.0);
      printf("{\"a_major\": \"%s\", \"b_major\": \"%s\", \"grid\": %d, \"M\": 128, \"N\": %d, \"K\": 16, \"cycles_per_mma\": %.1f, \"tensor_floor_cycles\": %.1f, \"smem_operand_bytes\": %d, \"bytes_per_cycle\": %.1f, \"err\": \"%s\"}\n",
             (mode & 1) ? "MN" : "K", (mode & 2) ? "MN" : "K", grid, N, per, 128.0 * N / 256.0, 4096 + N * 32, (4096 + N * 32) / per, cudaGetErrorString(e));
    }
  return 0;
}
