// Experiment: can a SWIZZLE_128B K-major A operand start at a 128-byte row that is NOT a multiple of 8 rows (1024 B), with a
// stride of 10 rows (1280 B) between its 8-row groups?  That is what a 3x3 convolution needs to read all nine taps out of ONE
// shared-memory box {64 ch, tw+2 = 10, th+2} instead of three horizontally shifted boxes.
// X: 190 rows x 64 bf16 laid out as TMA SWIZZLE_128B would (16-byte chunk c of row R at R*128 + ((c ^ (R & 7)) << 4)).
// B = identity (64 x 64), so D[m][n] should equal X[rowoff + 10*(m/8) + m%8][n].
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I ccdm_b200/csrc -I include -o gpurun_out/shift_desc tools/ubench/shift_desc.cu
#include <cstdio>
#include <cuda_runtime.h>
#include <cuda.h>
#include "ptx.cuh"
using namespace ccdm;

constexpr int kRows = 192;

__global__ void __launch_bounds__(128, 1) shift_kernel(int rowoff, int sbo_bytes, int base_off_mode, float* out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  uint8_t* xa = smem;                      // kRows x 128 B
  uint8_t* xb = smem + kRows * 128;        // 64 x 128 B identity (24576 is 1024-aligned)
  for (int i = tid; i < kRows * 64; i += 128) {
    const int R = i / 64, ch = i % 64, c = ch / 8, e = ch % 8;
    const float v = (float)(R * 2 + (ch % 2)) + (ch / 2) * 0.0f;      // value identifies the row (exact in bf16 for R < 128*...)
    __nv_bfloat16 h = __float2bfloat16((float)(R) + (ch == 0 ? 0.f : 0.f) + (float)ch * 256.f);
    (void)v;
    *reinterpret_cast<__nv_bfloat16*>(xa + R * 128 + ((c ^ (R & 7)) << 4) + e * 2) = h;
  }
  for (int i = tid; i < 64 * 64; i += 128) {
    const int R = i / 64, ch = i % 64, c = ch / 8, e = ch % 8;
    *reinterpret_cast<__nv_bfloat16*>(xb + R * 128 + ((c ^ (R & 7)) << 4) + e * 2) = __float2bfloat16(R == ch ? 1.f : 0.f);
  }
  if (warp == 0) tmem_alloc(&tmem_slot, 64);
  if (tid == 32) { mbar_init(&bar, 1); fence_mbar_init(); }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  if (warp == 1) {
    const uint32_t idesc = umma_idesc_bf16(128, 64);
    const uint32_t a_addr = smem_u32(xa) + rowoff * 128;
    const uint32_t b16 = (smem_u32(xb) & 0x3FFFF) >> 4;
    uint64_t hi = (static_cast<uint64_t>(sbo_bytes >> 4) << 32) | (static_cast<uint64_t>(1) << 46) | (static_cast<uint64_t>(2) << 61);
    if (base_off_mode == 1) hi |= static_cast<uint64_t>((a_addr >> 7) & 7) << 49;
    if (elect_one()) {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const uint64_t adesc = hi | (1ull << 16) | (((a_addr & 0x3FFFF) >> 4) + 2 * k);
        umma_bf16_ss(tmem, adesc, umma_desc_sw128_a16(b16 + 2 * k), idesc, k != 0);
      }
      umma_commit(&bar);
    }
    __syncwarp();
  }
  mbar_wait(&bar, 0);
  tc_fence_after();
  uint32_t r[32];
  for (int c = 0; c < 2; ++c) {
    tmem_ld32(tmem + (static_cast<uint32_t>(warp * 32) << 16) + c * 32, r);
    tmem_ld_wait();
    for (int i = 0; i < 32; ++i) out[(warp * 32 + lane) * 64 + c * 32 + i] = __uint_as_float(r[i]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem, 64); }
}

int main() {
  float* d; cudaMalloc(&d, 128 * 64 * 4);
  static float h[128 * 64];
  const size_t smem = kRows * 128 + 64 * 128 + 2048;
  cudaFuncSetAttribute(shift_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  const int offs[] = {0, 1, 2, 7, 8, 10, 11, 21, 22};
  for (int mode = 0; mode < 2; ++mode)
    for (int sbo : {1024, 1280})
      for (int off : offs) {
        shift_kernel<<<1, 128, smem>>>(off, sbo, mode, d);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("mode %d sbo %d off %d: CUDA error %s\n", mode, sbo, off, cudaGetErrorString(e)); return 1; }
        cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
        int bad = 0, first = -1;
        const int stride_rows = sbo / 128;
        for (int m = 0; m < 128; ++m)
          for (int n = 0; n < 64; ++n) {
            const int R = off + stride_rows * (m / 8) + (m % 8);
            const float want = (float)R + (float)n * 256.f;
            // bf16 rounding of want: emulate
            __nv_bfloat16 wb = __float2bfloat16(want);
            if (h[m * 64 + n] != __bfloat162float(wb)) { if (first < 0) first = m * 64 + n; ++bad; }
          }
        printf("base_offset_mode %d  SBO %4d  rowoff %2d : %s (%d mismatches%s)\n", mode, sbo, off, bad ? "WRONG" : "ok", bad,
               bad ? "" : "");
        if (bad && first >= 0) printf("    first mismatch m=%d n=%d got %.1f\n", first / 64, first % 64, h[first]);
      }
  return 0;
}
