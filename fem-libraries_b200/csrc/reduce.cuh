// reduce.cuh -- deterministic grid-wide sum fused into the producing kernel.
//
// Every CTA reduces its contribution with shuffles (fixed tree), stores one
// partial, and takes a ticket; the CTA that draws the last ticket sums all
// partials in a fixed order and writes the result.  No floating-point atomics:
// the value depends only on the grid shape, never on scheduling, so CG runs are
// reproducible bit for bit (the reference's MPI_Allreduce of the CG dot products
// has the same property for a fixed rank count).
#pragma once
#include "common.cuh"

namespace femb {

struct ReduceScratch
{
   double *partials;      // [>= gridDim.x * nvals]
   unsigned int *ticket;  // zero between kernels
};

// per-(device, stream) scratch, grown on demand; never freed before process exit
int reduce_scratch(size_t nblocks, cudaStream_t st, ReduceScratch *out, int nvals = 1);

template <int THREADS>
__device__ __forceinline__ double block_sum(double v, double *sh /* [THREADS/32] */)
{
   v = warp_sum(v);
   const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
   __syncthreads();  // protect sh from the previous use
   if (lane == 0) sh[w] = v;
   __syncthreads();
   double t = 0.;
   if (w == 0)
   {
      t = (lane < THREADS / 32) ? sh[lane] : 0.;
      t = warp_sum(t);
   }
   return t;  // valid in warp 0
}

// Sum `part` over the whole grid into out[0].  Must be reached by every thread of
// every CTA of the launch.
template <int THREADS, int NV = 1>
__device__ __forceinline__ void block_reduce_finish_n(const double (&part)[NV], ReduceScratch red, double *out,
                                                      bool accumulate = false)
{
   __shared__ double sh[THREADS / 32];
   __shared__ bool last;
#pragma unroll
   for (int k = 0; k < NV; ++k)
   {
      const double t = block_sum<THREADS>(part[k], sh);
      if (threadIdx.x == 0) red.partials[(size_t)k * gridDim.x + blockIdx.x] = t;
   }
   if (threadIdx.x == 0)
   {
      __threadfence();
      const unsigned int n = atomicAdd(red.ticket, 1u);
      last = (n == gridDim.x - 1);
   }
   __syncthreads();
   if (!last) return;
   __threadfence();
#pragma unroll
   for (int k = 0; k < NV; ++k)
   {
      double s = 0.;
      const volatile double *p = red.partials + (size_t)k * gridDim.x;
      for (unsigned int i = threadIdx.x; i < gridDim.x; i += THREADS) s += p[i];
      const double t = block_sum<THREADS>(s, sh);
      if (threadIdx.x == 0) out[k] = accumulate ? out[k] + t : t;
   }
   if (threadIdx.x == 0) *red.ticket = 0u;
}

template <int THREADS>
__device__ __forceinline__ void block_reduce_finish(double part, ReduceScratch red, double *out, bool accumulate = false)
{
   const double p[1] = {part};
   block_reduce_finish_n<THREADS, 1>(p, red, out, accumulate);
}

}  // namespace femb
