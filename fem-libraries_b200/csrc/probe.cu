// probe.cu -- FP64 FMA throughput probe: the measured denominator of the FP64 roofline that bench.py reports
// for the element kernels (north_star: "with the FP64 roofline also reported for the element kernel";
// MEASURED_PEAKS.json carries no FP64 figure).
#include "common.cuh"

namespace femb {

constexpr int kProbeChains = 8;

// every thread runs kProbeChains independent dependent-FMA chains: 2 * kProbeChains * iters flops per thread
__global__ void __launch_bounds__(256) fp64_fma_kernel(double *__restrict__ out, int iters, double b, double c)
{
   double a[kProbeChains];
#pragma unroll
   for (int k = 0; k < kProbeChains; ++k) a[k] = 1e-3 * (threadIdx.x + 1) + k;
   for (int i = 0; i < iters; ++i)
   {
#pragma unroll
      for (int k = 0; k < kProbeChains; ++k) a[k] = fma(a[k], b, c);
   }
   double s = 0.;
#pragma unroll
   for (int k = 0; k < kProbeChains; ++k) s += a[k];
   out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = s;
}

}  // namespace femb

using namespace femb;

// Launches the probe once on `stream` (time it with events around the call): blocks_per_sm * SM count blocks of
// 256 threads, `iters` rounds of 8 FMAs per thread.  *flops = floating-point operations of the launch (FMA = 2).
// d_out: blocks * 256 doubles.
extern "C" int femb200_fp64_probe(int blocks_per_sm, int iters, double *d_out, int64_t *blocks_out, double *flops, void *stream)
{
   FEMB_CHECK(blocks_per_sm >= 1 && blocks_per_sm <= 8 && iters >= 1, "fp64_probe: bad argument");
   const int64_t blocks = (int64_t)blocks_per_sm * devinfo().sm_count;
   if (blocks_out) *blocks_out = blocks;
   if (flops) *flops = 2. * kProbeChains * (double)iters * 256. * (double)blocks;
   if (!d_out) return 0;  // size query
   fp64_fma_kernel<<<(unsigned)blocks, 256, 0, as_stream(stream)>>>(d_out, iters, 0.999999, 1e-7);
   FEMB_LAUNCH_CHECK();
   return 0;
}
