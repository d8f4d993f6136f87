// Microbenchmark: how fast can ONE SM (with all 148 running) move NHWC bf16 activation boxes between global and shared
// memory with TMA, with nothing else going on?  The 3x3 tap-GEMM layers of the 64-channel levels deliver 13-16 bytes per
// clock and SM of box + output bytes whatever their instruction count (profiles/r2_notes.md); this separates "that is the
// TMA / L2 -> SM delivery rate for 128-byte box rows" from "the kernel leaves TMA idle".
//
// A persistent CTA per SM walks its share of the 8 x 16 pixel tiles of a [B][H][W][64] tensor: one elected lane issues the
// box loads into a ring of stages (mbarrier complete_tx), a second warp "consumes" a stage the moment it lands, and --
// optionally -- one 16 KB TMA store per tile leaves from a staging buffer, as the tap-GEMM epilogue's output does.
// modes: box = 0: {64, 16, 10} three-row-reuse box of the R = 3 plan (x3 per tile: dw = -1, 0, 1), 1: {64, 10, 18} halo box.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I ccdm_b200/csrc -I include -o /tmp/tma_rate tools/ubench/tma_rate.cu -lcuda
#include <cstdio>
#include <cstring>
#include <cuda.h>
#include <cuda_runtime.h>
#include "ptx.cuh"
using namespace ccdm;

constexpr int kStages = 6;

struct Maps {
  CUtensorMap in, out;
};

__global__ void __launch_bounds__(96, 1) tma_rate_kernel(const __grid_constant__ Maps maps, int tiles_w, int tiles_h, int B,
                                                         int boxes_per_tile, uint32_t box_bytes, uint32_t slot_bytes,
                                                         int with_store, long long* cycles) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t full[kStages], empty[kStages];
  uint8_t* ring = smem;
  uint8_t* stg = smem + (size_t)kStages * slot_bytes;
  const int tid = threadIdx.x, warp = tid >> 5;
  if (tid == 0) {
    for (int s = 0; s < kStages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    fence_mbar_init();
    tma_prefetch_desc(&maps.in);
    tma_prefetch_desc(&maps.out);
  }
  for (int i = tid; i < 16384 / 4; i += 96) reinterpret_cast<uint32_t*>(stg)[i] = 0x3C003C00u;
  fence_proxy_async_smem();
  __syncthreads();
  const int tiles = tiles_w * tiles_h * B;
  const int per = (tiles + gridDim.x - 1) / gridDim.x;
  const int t0 = blockIdx.x * per, t1 = min(tiles, t0 + per);
  const long long c0 = clock64();
  if (warp == 0) {
    int s = 0; uint32_t ph = 0;
    for (int t = t0; t < t1; ++t) {
      const int tx = t % tiles_w, ty = (t / tiles_w) % tiles_h, tz = t / (tiles_w * tiles_h);
      for (int bx = 0; bx < boxes_per_tile; ++bx) {
        mbar_wait(&empty[s], ph ^ 1u);
        if (elect_one()) {
          mbar_arrive_expect_tx(&full[s], box_bytes);
          const int dw = boxes_per_tile == 3 ? bx - 1 : -1;
          if (boxes_per_tile == 3) tma_load_4d(&maps.in, &full[s], ring + (size_t)s * slot_bytes, 0, tx * 16 + dw, ty * 8 - 1, tz);
          else tma_load_4d(&maps.in, &full[s], ring + (size_t)s * slot_bytes, 0, tx * 8 - 1, ty * 16 - 1, tz);
        }
        __syncwarp();
        if (++s == kStages) { s = 0; ph ^= 1u; }
      }
    }
  } else if (warp == 1) {
    int s = 0; uint32_t ph = 0;
    for (int t = t0; t < t1; ++t) {
      for (int bx = 0; bx < boxes_per_tile; ++bx) {
        mbar_wait(&full[s], ph);
        if (elect_one()) mbar_arrive(&empty[s]);
        __syncwarp();
        if (++s == kStages) { s = 0; ph ^= 1u; }
      }
      if (with_store && elect_one()) {
        const int tx = t % tiles_w, ty = (t / tiles_w) % tiles_h, tz = t / (tiles_w * tiles_h);
        tma_store_wait_read0();
        if (boxes_per_tile == 3) tma_store_4d(&maps.out, stg, 0, tx * 16, ty * 8, tz);
        else tma_store_4d(&maps.out, stg, 0, tx * 8, ty * 16, tz);
        tma_store_commit();
      }
      __syncwarp();
    }
    if (elect_one()) tma_store_wait_all();
    __syncwarp();
  }
  __syncthreads();
  if (tid == 0) cycles[blockIdx.x] = clock64() - c0;
}

static int encode(CUtensorMap* m, void* base, int C, int W, int H, int B, int bw, int bh) {
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
  cuuint64_t str[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
  cuuint32_t box[4] = {64, (cuuint32_t)bw, (cuuint32_t)bh, 1};
  cuuint32_t es[4] = {1, 1, 1, 1};
  return (int)cuTensorMapEncodeTiled(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, base, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
}

int main() {
  cudaFree(0);
  const int B = 400, H = 64, W = 64;
  long long* d_cyc; cudaMalloc(&d_cyc, 148 * 8);
  for (int C : {64, 72}) {
    void *x, *y;
    const size_t bytes = (size_t)B * H * W * C * 2;
    cudaMalloc(&x, bytes); cudaMalloc(&y, bytes);
    cudaMemset(x, 0, bytes);
    for (int mode = 0; mode < 2; ++mode)
      for (int with_store = 0; with_store < 2; ++with_store) {
        Maps maps;
        std::memset(&maps, 0, sizeof(maps));
        const int bw = mode ? 10 : 16, bh = mode ? 18 : 10, ow = mode ? 8 : 16, oh = mode ? 16 : 8;
        if (encode(&maps.in, x, C, W, H, B, bw, bh) || encode(&maps.out, y, C, W, H, B, ow, oh)) { printf("encode failed\n"); return 1; }
        const int tiles_w = W / ow, tiles_h = H / oh, boxes = mode ? 1 : 3;
        const uint32_t box_bytes = (uint32_t)bw * bh * 128, slot = (box_bytes + 1023u) & ~1023u;
        const size_t smem = (size_t)kStages * slot + 16384 + 1024;
        cudaFuncSetAttribute(tma_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        tma_rate_kernel<<<148, 96, smem>>>(maps, tiles_w, tiles_h, B, boxes, box_bytes, slot, with_store, d_cyc);
        cudaDeviceSynchronize();
        cudaEventRecord(e0);
        tma_rate_kernel<<<148, 96, smem>>>(maps, tiles_w, tiles_h, B, boxes, box_bytes, slot, with_store, d_cyc);
        cudaEventRecord(e1);
        cudaError_t err = cudaDeviceSynchronize();
        float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
        long long cyc[148]; cudaMemcpy(cyc, d_cyc, sizeof(cyc), cudaMemcpyDeviceToHost);
        long long mx = 0; for (long long c : cyc) mx = c > mx ? c : mx;
        const int tiles = tiles_w * tiles_h * B, per = (tiles + 147) / 148;
        const double smem_bytes_per_tile = (double)boxes * box_bytes + (with_store ? 16384.0 : 0.0);
        printf("{\"C\": %d, \"box\": \"%s\", \"boxes_per_tile\": %d, \"tma_store\": %d, \"us\": %.1f, \"clk_per_tile\": %.0f, "
               "\"box_rows_per_tile\": %d, \"clk_per_128B_row\": %.2f, \"smem_bytes_per_clk_per_sm\": %.1f, \"global_TBps\": %.2f, \"err\": \"%s\"}\n",
               C, mode ? "{64,10,18} halo" : "{64,16,10} x3", boxes, with_store, ms * 1e3, (double)mx / per,
               boxes * bw * bh + (with_store ? 128 : 0), (double)mx / per / (boxes * bw * bh + (with_store ? 128 : 0)),
               smem_bytes_per_tile / ((double)mx / per),
               ((double)tiles * (boxes * bw * bh * (double)C * 2 + (with_store ? 128.0 * C * 2 : 0.0))) / (ms * 1e-3) / 1e12,
               cudaGetErrorString(err));
      }
    cudaFree(x); cudaFree(y);
  }
  return 0;
}
