// Multi-tensor weight packing for the training step: every convolution weight is re-packed to K-blocked bf16 once per
// optimizer step, in the forward layout (ccdm_pack_weights) and in the transposed layout of the data gradient
// (ccdm_pack_weights_t) -- ~400 launches of a few microseconds each (7 % of a UK64 step).  One launch walks a device-resident
// job table instead (grid.y = job).  Same arithmetic as the two single-tensor kernels (tapgemm.cu / wgrad.cu).
#include "common.cuh"
#include "ptx.cuh"

namespace ccdm {

__global__ void __launch_bounds__(256) pack_multi_kernel(const ccdm_pack_job* __restrict__ jobs) {
  const ccdm_pack_job jb = jobs[blockIdx.y];
  const float* __restrict__ w = jb.w;
  const int4* __restrict__ psched = reinterpret_cast<const int4*>(jb.psched);
  __nv_bfloat16* __restrict__ out = reinterpret_cast<__nv_bfloat16*>(jb.out);
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < jb.total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int j = (int)(idx & 63);
    const long long t = idx >> 6;
    const int kb = (int)(t % jb.nkb);
    const long long zn = t / jb.nkb;
    const int n = (int)(zn % jb.n_rows);
    const int z = (int)(zn / jb.n_rows);
    const int4 e = psched[z * jb.nkb + kb];
    float acc = 0.f;
    const unsigned mask = (unsigned)e.z;
    if (jb.mode == 0) {                                    // forward layout: W[n][cin0 + j][taps]
      if (n < jb.cout && j < e.y) {
        const float* wp = w + ((long long)n * jb.cin_total + (e.x + j)) * jb.ntaps;
        for (int tp = 0; tp < jb.ntaps; ++tp)
          if (mask & (1u << tp)) acc += wp[tp];
      }
    } else {                                               // transposed: W[cin0 + j][n_off + n][taps]
      if (n < jb.n_count && j < e.y) {
        const float* wp = w + ((long long)(e.x + j) * jb.cin_total + (jb.n_off + n)) * jb.ntaps;
        for (int tp = 0; tp < jb.ntaps; ++tp)
          if (mask & (1u << tp)) acc += wp[tp];
      }
    }
    out[idx] = __float2bfloat16(acc);
  }
}

}  // namespace ccdm

using namespace ccdm;

extern "C" int ccdm_pack_multi(const ccdm_pack_job* jobs, int32_t njobs, void* stream) {
  CCDM_REQUIRE(jobs && njobs > 0 && njobs <= 65535, CCDM_ERR_BAD_ARG, "pack_multi: %d jobs", njobs);
  dim3 grid(48, (unsigned)njobs);
  pack_multi_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(jobs);
  return after_launch("pack_multi_kernel");
}
