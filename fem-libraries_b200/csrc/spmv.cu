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
#include <algorithm>
#include <type_traits>

#include "plan.cuh"
#include "reduce.cuh"
#include "tma.cuh"

namespace femb {

constexpr int kSpmvLanes = 8;
constexpr int kSpmvThreads = 256;

template <bool DOT>
__global__ void __launch_bounds__(kSpmvThreads)
spmv_kernel(int64_t row_lo, int64_t nnodes, const int64_t *__restrict__ brp, const int32_t *__restrict__ bcol,
            const double *__restrict__ values, const double *__restrict__ x, double *__restrict__ y,
            const double *__restrict__ flag, ReduceScratch red, double *__restrict__ out, bool accumulate = false)
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
   if (DOT) block_reduce_finish<kSpmvThreads>(part, red, out, accumulate);
}

// ---------------------------------------------------------------------------
// TMA-staged persistent variant (default).  The node rows [n0, n0 + R) of a tile
// own one contiguous byte range of the value array, of bcol and of brp: a producer
// warp streams those three ranges into shared memory with 1-D bulk copies
// (cp.async.bulk -> UBLKCP) completing on an mbarrier, S stages deep, while eight
// consumer warps (L lanes per node row pair) reduce the previous tiles out of shared memory; the only
// per-thread global loads left are the gathers of x (L1/L2 hits: the lattice
// numbering keeps the reuse window at a few node rows).  Grid = resident CTAs only,
// so the fused <x, y> needs ~300 tickets instead of one per 32 rows.
// ---------------------------------------------------------------------------
constexpr int kTmaConsumers = 256;
constexpr int kTmaThreads = kTmaConsumers + 32;

struct SpmvTile
{
   int vbytes, cbytes, pbytes;  // stage capacities: values, column indices, row pointers
};

// C16: the column indices are the plan's 16-bit offsets from the row's own node (bcol16: every |J - I| of the pattern fits,
// which any numbering with locality gives: the lattice numbering of config 4 has |J - I| <= 2 (2 n + 1) + 2 = 23 172):
// 34 instead of 36 bytes per node block, 5 % of the kernel's DRAM traffic.
template <bool DOT, int R, int kTmaStages, int L, bool C16>
__global__ void __launch_bounds__(kTmaThreads)
spmv_tma_kernel(int64_t row_lo, int64_t row_hi, const int64_t *__restrict__ brp, const void *__restrict__ bcol_any,
                const double *__restrict__ values, const double *__restrict__ x, double *__restrict__ y,
                const double *__restrict__ flag, ReduceScratch red, double *__restrict__ out, SpmvTile cap, int ntiles,
                bool accumulate)
{
   extern __shared__ __align__(128) unsigned char smem[];
   __shared__ uint64_t full[kTmaStages], empty[kTmaStages];
   if (flag && *flag != 0.) return;  // converged CG: become a no-op (uniform over the grid)
   const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
   const int stage_bytes = cap.vbytes + cap.cbytes + cap.pbytes;
   if (tid == 0)
   {
#pragma unroll
      for (int s = 0; s < kTmaStages; ++s)
      {
         mbar_init(&full[s], 1);
         mbar_init(&empty[s], kTmaConsumers / 32);
      }
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
   }
   __syncthreads();
   const int first = blockIdx.x, stride = gridDim.x;
   const int nmine = first < ntiles ? (ntiles - first + stride - 1) / stride : 0;
   double part = 0.;
   if (warp == kTmaConsumers / 32)
   {  // ---- producer warp: one lane issues the bulk copies ----
      if (lane == 0)
         for (int it = 0; it < nmine; ++it)
         {
            const int s = it % kTmaStages;
            if (it >= kTmaStages) mbar_wait(&empty[s], ((it / kTmaStages) - 1) & 1);
            const int64_t n0 = row_lo + (int64_t)(first + it * stride) * R;
            const int nloc = (int)min((int64_t)R, row_hi - n0);
            const int64_t b0 = brp[n0], b1 = brp[n0 + nloc];
            constexpr int CW = C16 ? 2 : 4;                                    // bytes per column index
            const int64_t c0 = b0 & ~(int64_t)(16 / CW - 1), p0 = n0 & ~(int64_t)1;
            const uint32_t vb = (uint32_t)(32 * (b1 - b0));
            const uint32_t cb = (uint32_t)((CW * (b1 - c0) + 15) & ~(int64_t)15);
            const uint32_t pb = (uint32_t)((8 * (n0 + nloc + 1 - p0) + 15) & ~(int64_t)15);
            unsigned char *st = smem + (size_t)s * stage_bytes;
            mbar_expect_tx(&full[s], vb + cb + pb);
            if (vb) bulk_g2s(st, values + 4 * b0, vb, &full[s]);
            if (cb) bulk_g2s(st + cap.vbytes, static_cast<const unsigned char *>(bcol_any) + CW * c0, cb, &full[s]);
            bulk_g2s(st + cap.vbytes + cap.cbytes, brp + p0, pb, &full[s]);
         }
   }
   else
   {  // ---- consumer warps: L lanes per node row pair ----
      constexpr int kSpmvLanes = L;
      const int sub = lane & (kSpmvLanes - 1);
      const double2 *x2 = reinterpret_cast<const double2 *>(x);
      for (int it = 0; it < nmine; ++it)
      {
         const int s = it % kTmaStages;
         const int64_t n0 = row_lo + (int64_t)(first + it * stride) * R;
         const int nloc = (int)min((int64_t)R, row_hi - n0);
         const unsigned char *st = smem + (size_t)s * stage_bytes;
         const double2 *sval = reinterpret_cast<const double2 *>(st);
         using col_t = typename std::conditional<C16, int16_t, int32_t>::type;
         const col_t *scol = reinterpret_cast<const col_t *>(st + cap.vbytes);
         const int64_t *sbrp = reinterpret_cast<const int64_t *>(st + cap.vbytes + cap.cbytes) + (n0 & 1);
         mbar_wait(&full[s], (it / kTmaStages) & 1);
         const int64_t b0 = sbrp[0];
         const int coff = (int)(b0 & (C16 ? 7 : 3));
#pragma unroll
         for (int pass = 0; pass < R * kSpmvLanes / kTmaConsumers; ++pass)
         {
            const int i = pass * (kTmaConsumers / kSpmvLanes) + tid / kSpmvLanes;
            double y0 = 0., y1 = 0.;
            if (i < nloc)
            {
               const int pb = (int)(sbrp[i] - b0);
               const int deg = (int)(sbrp[i + 1] - b0) - pb;
               const double2 *row0 = sval + 2 * pb, *row1 = row0 + deg;
               const col_t *cols = scol + coff + pb;
               const double2 *xr = C16 ? x2 + (n0 + i) : x2;  // 16-bit indices are offsets from the row's node
               int t = sub;
               for (; t + 2 * kSpmvLanes < deg; t += 3 * kSpmvLanes)
               {
                  const double2 v0 = xr[cols[t]], v1 = xr[cols[t + kSpmvLanes]], v2 = xr[cols[t + 2 * kSpmvLanes]];
                  const double2 a0 = row0[t], a1 = row0[t + kSpmvLanes], a2 = row0[t + 2 * kSpmvLanes];
                  const double2 c0 = row1[t], c1 = row1[t + kSpmvLanes], c2 = row1[t + 2 * kSpmvLanes];
                  y0 += a0.x * v0.x + a0.y * v0.y + a1.x * v1.x + a1.y * v1.y + a2.x * v2.x + a2.y * v2.y;
                  y1 += c0.x * v0.x + c0.y * v0.y + c1.x * v1.x + c1.y * v1.y + c2.x * v2.x + c2.y * v2.y;
               }
               if (t + kSpmvLanes < deg)
               {
                  const double2 v0 = xr[cols[t]], v1 = xr[cols[t + kSpmvLanes]];
                  const double2 a0 = row0[t], a1 = row0[t + kSpmvLanes];
                  const double2 c0 = row1[t], c1 = row1[t + kSpmvLanes];
                  y0 += a0.x * v0.x + a0.y * v0.y + a1.x * v1.x + a1.y * v1.y;
                  y1 += c0.x * v0.x + c0.y * v0.y + c1.x * v1.x + c1.y * v1.y;
               }
               else if (t < deg)
               {
                  const double2 v0 = xr[cols[t]];
                  const double2 a0 = row0[t], c0 = row1[t];
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
            if (i < nloc && sub == 0)
            {
               reinterpret_cast<double2 *>(y)[n0 + i] = make_double2(y0, y1);
               if (DOT)
               {
                  const double2 xi = x2[n0 + i];
                  part += xi.x * y0 + xi.y * y1;
               }
            }
         }
         __syncwarp();
         if (lane == 0) mbar_arrive(&empty[s]);
      }
   }
   if (DOT) block_reduce_finish<kTmaThreads>(part, red, out, accumulate);
}

__global__ void diag_kernel(int64_t nnodes, const int64_t *__restrict__ brp, const uint8_t *__restrict__ dslot,
                            const double *__restrict__ values, double *__restrict__ diag)
{
   const int64_t I = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
   if (I >= nnodes) return;
   const int64_t bi = brp[I];
   const int deg = (int)(brp[I + 1] - bi), s = dslot[I];
   double2 d = make_double2(0., 0.);
   if (s < deg) d = make_double2(values[4 * bi + 2 * s], values[4 * bi + 2 * deg + 2 * s + 1]);
   reinterpret_cast<double2 *>(diag)[I] = d;
}

template <bool DOT, int R, int S, int L, bool C16>
static int spmv_tma_launch(const femb200_plan *p, const RowRange &rr, const double *d_values, const double *d_x,
                           double *d_y, const double *d_flag, double *d_dot_out, bool accumulate, cudaStream_t st)
{
   static_assert(R == 32 || R == 64, "RowRange::tile_max is measured for 32- and 64-row tiles");
   static_assert(R * L % kTmaConsumers == 0, "a tile must be a whole number of consumer passes");
   const int64_t nrows = rr.hi - rr.lo;
   const int ntiles = (int)cdiv(nrows, R);
   const int maxb = rr.tile_max[R == 32 ? 0 : 1];
   SpmvTile cap;
   cap.vbytes = 32 * maxb;
   cap.cbytes = C16 ? ((2 * (maxb + 8) + 15) & ~15) : ((4 * (maxb + 4) + 15) & ~15);
   cap.pbytes = ((8 * (R + 3) + 15) & ~15);
   const size_t smem = (size_t)S * (cap.vbytes + cap.cbytes + cap.pbytes);
   const size_t budget = devinfo().smem_optin ? devinfo().smem_optin : 227 * 1024;
   if (smem > budget) return -1;  // caller falls back to the direct kernel
   if (int rc = ensure_dynamic_smem<spmv_tma_kernel<DOT, R, S, L, C16>>(smem)) return rc;
   const int per_sm = (int)std::max<size_t>(1, std::min<size_t>(7, (budget + 1024) / (smem + 1024)));
   const unsigned grid = (unsigned)std::min<int64_t>(ntiles, (int64_t)devinfo().sm_count * per_sm);
   ReduceScratch red{nullptr, nullptr};
   if (DOT)
      if (int rc = reduce_scratch(grid, st, &red)) return rc;
   spmv_tma_kernel<DOT, R, S, L, C16><<<grid, kTmaThreads, smem, st>>>(
       rr.lo, rr.hi, p->brp, C16 ? static_cast<const void *>(p->bcol16) : static_cast<const void *>(p->bcol), d_values, d_x, d_y,
       d_flag, red, d_dot_out, cap, ntiles, accumulate);
   FEMB_LAUNCH_CHECK();
   return 0;
}

// y = A x on the node rows of rr (+ d_dot_out (+)= <x, y> over those rows; no-op when *d_flag != 0).
// Default: 64-row tiles, 2 stages (3 CTAs per SM), 4 lanes per row: 0.59 ms on the n = 1448 P2 matrix against
// 0.99 ms with 3 stages and 8 lanes (profiles/r1_summary.md).  Tiles too large for shared memory, and plans
// with spmv_path = 1, take the direct kernel.
int spmv_launch(const femb200_plan *p, const RowRange &rr, const double *d_values, const double *d_x, double *d_y,
                const double *d_flag, double *d_dot_out, bool accumulate, cudaStream_t st)
{
   const int64_t nrows = rr.hi - rr.lo;
   if (nrows <= 0)
   {
      if (d_dot_out && !accumulate) FEMB_CUDA(cudaMemsetAsync(d_dot_out, 0, sizeof(double), st));
      return 0;
   }
   if (p->opt_spmv_path == 0)
   {
      const bool c16 = p->bcol16 != nullptr && p->opt_spmv_cols != 1;
      int rc;
      if (c16)
         rc = d_dot_out ? spmv_tma_launch<true, 64, 2, 4, true>(p, rr, d_values, d_x, d_y, d_flag, d_dot_out, accumulate, st)
                        : spmv_tma_launch<false, 64, 2, 4, true>(p, rr, d_values, d_x, d_y, d_flag, d_dot_out, accumulate, st);
      else
         rc = d_dot_out ? spmv_tma_launch<true, 64, 2, 4, false>(p, rr, d_values, d_x, d_y, d_flag, d_dot_out, accumulate, st)
                        : spmv_tma_launch<false, 64, 2, 4, false>(p, rr, d_values, d_x, d_y, d_flag, d_dot_out, accumulate, st);
      if (rc >= 0) return rc;
   }
   const unsigned grid = (unsigned)cdiv(nrows * kSpmvLanes, kSpmvThreads);
   if (d_dot_out)
   {
      ReduceScratch red;
      if (int rc = reduce_scratch(grid, st, &red)) return rc;
      spmv_kernel<true><<<grid, kSpmvThreads, 0, st>>>(rr.lo, rr.hi, p->brp, p->bcol, d_values, d_x, d_y, d_flag, red,
                                                        d_dot_out, accumulate);
   }
   else
      spmv_kernel<false><<<grid, kSpmvThreads, 0, st>>>(rr.lo, rr.hi, p->brp, p->bcol, d_values, d_x, d_y, d_flag,
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
   RowRange rr;
   if (int rc = plan_row_range(p, 0, p->nnodes, &rr)) return rc;
   return spmv_launch(p, rr, d_values, d_x, d_y, nullptr, nullptr, false, as_stream(stream));
}

extern "C" int femb200_spmv_dot(const femb200_plan *p, const double *d_values, const double *d_x, double *d_y,
                                double *d_dot, void *stream)
{
   FEMB_CHECK(p && d_values && d_x && d_y && d_dot, "spmv_dot: null argument");
   FEMB_CHECK(d_x != d_y, "spmv_dot: x and y must not alias");
   RowRange rr;
   if (int rc = plan_row_range(p, 0, p->nnodes, &rr)) return rc;
   return spmv_launch(p, rr, d_values, d_x, d_y, nullptr, d_dot, false, as_stream(stream));
}

// y = A x on the node rows [row_lo, row_hi) (the rows a rank owns, or the rows next to its ghosts); rows
// outside are left untouched.  d_dot (or NULL) receives <x, y> over those rows, added to its content when
// `accumulate`; d_flag (or NULL): no-op when *d_flag != 0 (converged CG).  The first call with a range that
// does not start on a 64-row boundary measures its tiling (synchronises the device once).
extern "C" int femb200_spmv_rows(const femb200_plan *p, const double *d_values, const double *d_x, double *d_y,
                                 int64_t row_lo, int64_t row_hi, double *d_dot, int accumulate, const double *d_flag,
                                 void *stream)
{
   FEMB_CHECK(p && d_values && d_x && d_y, "spmv_rows: null argument");
   FEMB_CHECK(d_x != d_y, "spmv_rows: x and y must not alias");
   RowRange rr;
   if (int rc = plan_row_range(p, row_lo, row_hi, &rr)) return rc;
   return spmv_launch(p, rr, d_values, d_x, d_y, d_flag, d_dot, accumulate != 0, as_stream(stream));
}

extern "C" int femb200_extract_diagonal(const femb200_plan *p, const double *d_values, double *d_diag, void *stream)
{
   FEMB_CHECK(p && d_values && d_diag, "extract_diagonal: null argument");
   const int T = 256;
   diag_kernel<<<(unsigned)cdiv(p->nnodes, T), T, 0, as_stream(stream)>>>(p->nnodes, p->brp, p->dslot, d_values, d_diag);
   FEMB_LAUNCH_CHECK();
   return 0;
}
