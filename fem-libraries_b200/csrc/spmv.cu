// spmv.cu -- assembled operator apply y = A x on the node-block pattern.
//
// Role in the reference: HypreParMatrix::Mult inside mfem::CGSolver
// (M.cc:1502,1525-1528) / PETSc MatMult inside KSP cg (F.cc:718-722).
//
// The matrix values live in the scalar-CSR layout handed across the boundary
// (row 2I: 2 deg doubles, row 2I+1: 2 deg doubles, contiguous), but the kernel
// walks the node-block pattern (brp/bcol): one int32 column index per 2x2 block
// instead of four, i.e. 36 B per block instead of 48 B.  Eight lanes own one
// node row pair; lane s handles blocks s, s+8, ...: 16-byte loads of both
// scalar rows and of x[2J..2J+1], shuffle reduction over the eight lanes.
// <x, y> is optionally fused (CG needs <d, A d>): deterministic two-stage
// reduction (block partials, last block sums them in a fixed order).
#include "plan.cuh"
#include "reduce.cuh"

namespace femb {

constexpr int kSpmvLanes = 8;
constexpr int kSpmvThreads = 256;

template <bool DOT>
__global__ void __launch_bounds__(kSpmvThreads)
spmv_kernel(int64_t row_lo, int64_t nnodes, const int64_t *__restrict__ brp, const int32_t *__restrict__ bcol,
            const double *__restrict__ values, const double *__restrict__ x, double *__restrict__ y,
            const double *__restrict__ flag, ReduceScratch red, double *__restrict__ out)
{
   if (flag && *flag != 0.) return;  // converged CG: become a no-op (uniform over the grid)
   const int lane = threadIdx.x & (kSpmvLanes - 1);
   const int64_t I = row_lo + ((int64_t)blockIdx.x * kSpmvThreads + threadIdx.x) / kSpmvLanes;
   double y0 = 0., y1 = 0.;
   if (I < nnodes)
   {
      const int64_t bi = brp[I];
      const int deg = (int)(brp[I + 1] - bi);
      const double2 *row0 = reinterpret_cast<const double2 *>(values + 4 * bi);
      const double2 *row1 = row0 + deg;
      const int32_t *cols = bcol + bi;
      const double2 *x2 = reinterpret_cast<const double2 *>(x);
      // three blocks per lane in flight (deg <= 24 covers P1/P2/Q2 interior rows)
      int s = lane;
      for (; s + 2 * kSpmvLanes < deg; s += 3 * kSpmvLanes)
      {
         const int32_t j0 = cols[s], j1 = cols[s + kSpmvLanes], j2 = cols[s + 2 * kSpmvLanes];
         const double2 a0 = row0[s], a1 = row0[s + kSpmvLanes], a2 = row0[s + 2 * kSpmvLanes];
         const double2 c0 = row1[s], c1 = row1[s + kSpmvLanes], c2 = row1[s + 2 * kSpmvLanes];
         const double2 v0 = x2[j0], v1 = x2[j1], v2 = x2[j2];
         y0 += a0.x * v0.x + a0.y * v0.y + a1.x * v1.x + a1.y * v1.y + a2.x * v2.x + a2.y * v2.y;
         y1 += c0.x * v0.x + c0.y * v0.y + c1.x * v1.x + c1.y * v1.y + c2.x * v2.x + c2.y * v2.y;
      }
      for (; s < deg; s += kSpmvLanes)
      {
         const int32_t j0 = cols[s];
         const double2 a0 = row0[s], c0 = row1[s], v0 = x2[j0];
         y0 += a0.x * v0.x + a0.y * v0.y;
         y1 += c0.x * v0.x + c0.y * v0.y;
      }
   }
#pragma unroll
   for (int o = kSpmvLanes / 2; o > 0; o >>= 1)
   {
      y0 += __shfl_xor_sync(0xffffffffu, y0, o);
      y1 += __shfl_xor_sync(0xffffffffu, y1, o);
   }
   double part = 0.;
   if (I < nnodes && lane == 0)
   {
      reinterpret_cast<double2 *>(y)[I] = make_double2(y0, y1);
      if (DOT)
      {
         const double2 xi = reinterpret_cast<const double2 *>(x)[I];
         part = xi.x * y0 + xi.y * y1;
      }
   }
   if (DOT) block_reduce_finish<kSpmvThreads>(part, red, out);
}

__global__ void diag_kernel(int64_t nnodes, const int64_t *__restrict__ brp, const int32_t *__restrict__ bcol,
                            const double *__restrict__ values, double *__restrict__ diag)
{
   const int64_t I = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
   if (I >= nnodes) return;
   const int64_t bi = brp[I];
   const int deg = (int)(brp[I + 1] - bi);
   int lo = 0, hi = deg;
   while (lo < hi)
   {
      const int mid = (lo + hi) >> 1;
      if (bcol[bi + mid] < I)
         lo = mid + 1;
      else
         hi = mid;
   }
   double d0 = 0., d1 = 0.;
   if (lo < deg && bcol[bi + lo] == I)
   {
      d0 = values[4 * bi + 2 * lo];
      d1 = values[4 * bi + 2 * deg + 2 * lo + 1];
   }
   diag[2 * I] = d0;
   diag[2 * I + 1] = d1;
}

int spmv_launch(const femb200_plan *p, const double *d_values, const double *d_x, double *d_y, const double *d_flag,
                double *d_dot_out, cudaStream_t st)
{
   const int64_t nrows = p->row_hi - p->row_lo;
   if (nrows <= 0) return 0;
   const unsigned grid = (unsigned)cdiv(nrows * kSpmvLanes, kSpmvThreads);
   if (d_dot_out)
   {
      ReduceScratch red;
      if (int rc = reduce_scratch(grid, st, &red)) return rc;
      spmv_kernel<true><<<grid, kSpmvThreads, 0, st>>>(p->row_lo, p->row_hi, p->brp, p->bcol, d_values, d_x, d_y, d_flag, red,
                                                        d_dot_out);
   }
   else
      spmv_kernel<false><<<grid, kSpmvThreads, 0, st>>>(p->row_lo, p->row_hi, p->brp, p->bcol, d_values, d_x, d_y, d_flag,
                                                         ReduceScratch{nullptr, nullptr}, nullptr);
   FEMB_LAUNCH_CHECK();
   return 0;
}

}  // namespace femb

using namespace femb;

extern "C" int femb200_spmv(const femb200_plan *p, const double *d_values, const double *d_x, double *d_y, void *stream)
{
   FEMB_CHECK(p && d_values && d_x && d_y, "spmv: null argument");
   FEMB_CHECK(d_x != d_y, "spmv: x and y must not alias");
   return spmv_launch(p, d_values, d_x, d_y, nullptr, nullptr, as_stream(stream));
}

extern "C" int femb200_spmv_dot(const femb200_plan *p, const double *d_values, const double *d_x, double *d_y,
                                double *d_dot, void *stream)
{
   FEMB_CHECK(p && d_values && d_x && d_y && d_dot, "spmv_dot: null argument");
   FEMB_CHECK(d_x != d_y, "spmv_dot: x and y must not alias");
   return spmv_launch(p, d_values, d_x, d_y, nullptr, d_dot, as_stream(stream));
}

extern "C" int femb200_extract_diagonal(const femb200_plan *p, const double *d_values, double *d_diag, void *stream)
{
   FEMB_CHECK(p && d_values && d_diag, "extract_diagonal: null argument");
   const int T = 256;
   diag_kernel<<<(unsigned)cdiv(p->nnodes, T), T, 0, as_stream(stream)>>>(p->nnodes, p->brp, p->bcol, d_values, d_diag);
   FEMB_LAUNCH_CHECK();
   return 0;
}
