// cg.cu -- (Jacobi-)preconditioned conjugate gradients with device-resident
// scalars.
//
// Role in the reference: mfem::CGSolver lin_solv (SetRelTol 1e-12, SetMaxIter
// 2000, M.cc:1502,1525-1528) / PETSc KSP cg (F.cc:718-722) called by the Newton
// solver once per non-linear iteration.  The reference preconditions with HYPRE
// BoomerAMG (third party, out of scope: SURVEY.md 2 #6); here B = diag(A)^-1 or I.
//
// Semantics are mfem::CGSolver::Mult's: zero initial guess, r = b, d = B r,
// nom = <d, r>, r0 = max(nom rtol^2, atol^2); per iteration
//    alpha = nom / den;  x += alpha d;  r -= alpha A d;  betanom = <B r, r>;
//    stop if betanom <= r0;  beta = betanom / nom;  d = B r + beta d;
//    den = <d, A d>;  nom = betanom.
// One iteration is 3 vector kernels (update_xr, update_dir, spmv+dot) and 2
// one-thread scalar kernels; the host never reads a scalar inside the loop except
// for the convergence poll every `check_every` iterations.  After convergence the
// remaining queued kernels see flag != 0 and return immediately.
//
// The multi-GPU driver (femb200/dist.py) calls the same kernels one by one and
// all-reduces the freshly written partial (scal[SC_RED_*]) between a vector
// kernel and its scalar kernel: the analogue of the MPI_Allreduce inside
// CGSolver / KSP.
#include <algorithm>

#include "plan.cuh"
#include "reduce.cuh"

namespace femb {

int spmv_launch(const femb200_plan *p, const double *d_values, const double *d_x, double *d_y, const double *d_flag,
                double *d_dot_out, cudaStream_t st);
int pa_apply_launch(const femb200_pa *pa, const double *d_x, double *d_y, const double *d_flag, double *d_dot_out,
                    cudaStream_t st);

enum
{
   SC_NOM = 0,
   SC_DEN = 1,
   SC_BETANOM = 2,
   SC_R0 = 3,
   SC_FLAG = 4,
   SC_ITERS = 5,
   SC_FINAL = 6,
   SC_RTOL2 = 7,
   SC_ATOL2 = 8,
   SC_RED_NOM = 9,   // local (per-rank) sums written by the vector kernels
   SC_RED_DEN = 10,
   SC_RED_BETA = 11,
   SC_BETA = 12,
   SC_COUNT = 16
};

constexpr int kVecThreads = 256;

// x = 0, r = b, d = B b, partial <d, r>
__global__ void __launch_bounds__(kVecThreads)
cg_init_kernel(int64_t n, const double *__restrict__ b, const double *__restrict__ dinv, double *__restrict__ x,
               double *__restrict__ r, double *__restrict__ d, ReduceScratch red, double *__restrict__ out)
{
   double part = 0.;
   const int64_t stride = (int64_t)gridDim.x * kVecThreads;
   for (int64_t i = (int64_t)blockIdx.x * kVecThreads + threadIdx.x; i < n; i += stride)
   {
      const double bi = b[i];
      const double di = dinv ? dinv[i] * bi : bi;
      x[i] = 0.;
      r[i] = bi;
      d[i] = di;
      part += di * bi;
   }
   block_reduce_finish<kVecThreads>(part, red, out);
}

// x += alpha d;  r -= alpha z;  partial <B r, r>
__global__ void __launch_bounds__(kVecThreads)
cg_update_xr_kernel(int64_t n2, const double *__restrict__ scal, const double2 *__restrict__ d,
                    const double2 *__restrict__ z, const double2 *__restrict__ dinv, double2 *__restrict__ x,
                    double2 *__restrict__ r, ReduceScratch red, double *__restrict__ out)
{
   if (scal[SC_FLAG] != 0.) return;
   const double alpha = scal[SC_NOM] / scal[SC_DEN];
   double part = 0.;
   const int64_t stride = (int64_t)gridDim.x * kVecThreads;
   for (int64_t i = (int64_t)blockIdx.x * kVecThreads + threadIdx.x; i < n2; i += stride)
   {
      const double2 di = d[i], zi = z[i];
      double2 xi = x[i], ri = r[i];
      xi.x += alpha * di.x, xi.y += alpha * di.y;
      ri.x -= alpha * zi.x, ri.y -= alpha * zi.y;
      x[i] = xi;
      r[i] = ri;
      if (dinv)
      {
         const double2 pi = dinv[i];
         part += (pi.x * ri.x) * ri.x + (pi.y * ri.y) * ri.y;
      }
      else
         part += ri.x * ri.x + ri.y * ri.y;
   }
   block_reduce_finish<kVecThreads>(part, red, out);
}

// d = B r + beta d
__global__ void __launch_bounds__(kVecThreads)
cg_update_dir_kernel(int64_t n2, const double *__restrict__ scal, const double2 *__restrict__ r,
                     const double2 *__restrict__ dinv, double2 *__restrict__ d)
{
   if (scal[SC_FLAG] != 0.) return;
   const double beta = scal[SC_BETA];
   const int64_t stride = (int64_t)gridDim.x * kVecThreads;
   for (int64_t i = (int64_t)blockIdx.x * kVecThreads + threadIdx.x; i < n2; i += stride)
   {
      double2 ri = r[i];
      if (dinv)
      {
         const double2 pi = dinv[i];
         ri.x *= pi.x, ri.y *= pi.y;
      }
      double2 di = d[i];
      di.x = ri.x + beta * di.x;
      di.y = ri.y + beta * di.y;
      d[i] = di;
   }
}

__global__ void dinv_kernel(int64_t n, const double *__restrict__ diag, double *__restrict__ dinv)
{
   const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
   if (i < n) dinv[i] = 1. / diag[i];
}

__global__ void __launch_bounds__(kVecThreads)
dot_kernel(int64_t n, const double *__restrict__ a, const double *__restrict__ b, ReduceScratch red,
           double *__restrict__ out)
{
   double part = 0.;
   const int64_t stride = (int64_t)gridDim.x * kVecThreads;
   for (int64_t i = (int64_t)blockIdx.x * kVecThreads + threadIdx.x; i < n; i += stride) part += a[i] * b[i];
   block_reduce_finish<kVecThreads>(part, red, out);
}

// phase 0: after init (nom)   phase 1: after spmv+dot (den)   phase 2: after update_xr (betanom)
__global__ void cg_scalar_kernel(double *__restrict__ s, int phase)
{
   if (phase == 0)
   {
      const double nom = s[SC_RED_NOM];
      s[SC_NOM] = nom;
      s[SC_R0] = fmax(nom * s[SC_RTOL2], s[SC_ATOL2]);
      s[SC_FINAL] = nom;
      s[SC_ITERS] = 0.;
      s[SC_FLAG] = (nom <= s[SC_R0]) ? 1. : 0.;
      return;
   }
   if (s[SC_FLAG] != 0.) return;
   if (phase == 1)
   {
      const double den = s[SC_RED_DEN];
      s[SC_DEN] = den;
      if (!(den > 0.)) s[SC_FLAG] = 2.;  // not positive definite: mfem leaves the loop
   }
   else
   {
      const double betanom = s[SC_RED_BETA];
      s[SC_BETANOM] = betanom;
      s[SC_ITERS] += 1.;
      s[SC_FINAL] = betanom;
      if (betanom <= s[SC_R0])
         s[SC_FLAG] = 1.;
      else
      {
         s[SC_BETA] = betanom / s[SC_NOM];
         s[SC_NOM] = betanom;
      }
   }
}

static unsigned vec_grid(int64_t n)
{
   const int64_t want = cdiv(n, kVecThreads);
   const int64_t cap = (int64_t)devinfo().sm_count * 8;  // 8 resident CTAs of 256 threads per SM
   return (unsigned)std::max<int64_t>(1, std::min(want, cap));
}

}  // namespace femb

using namespace femb;

extern "C" int femb200_dot(int64_t n, const double *d_a, const double *d_b, double *d_out, void *stream)
{
   FEMB_CHECK(d_a && d_b && d_out && n >= 0, "dot: bad argument");
   cudaStream_t st = as_stream(stream);
   const unsigned grid = vec_grid(n);
   ReduceScratch red;
   if (int rc = reduce_scratch(grid, st, &red)) return rc;
   dot_kernel<<<grid, kVecThreads, 0, st>>>(n, d_a, d_b, red, d_out);
   FEMB_LAUNCH_CHECK();
   return 0;
}

extern "C" int femb200_jacobi_setup(int64_t n, const double *d_diag, double *d_dinv, void *stream)
{
   FEMB_CHECK(d_diag && d_dinv && n >= 0, "jacobi_setup: bad argument");
   dinv_kernel<<<(unsigned)cdiv(n, 256), 256, 0, as_stream(stream)>>>(n, d_diag, d_dinv);
   FEMB_LAUNCH_CHECK();
   return 0;
}

extern "C" int femb200_cg_set_tolerances(double *d_scal, double rtol, double atol, void *stream)
{
   FEMB_CHECK(d_scal, "cg_set_tolerances: null scalars");
   double h[SC_COUNT];
   for (int i = 0; i < SC_COUNT; ++i) h[i] = 0.;
   h[SC_RTOL2] = rtol * rtol;
   h[SC_ATOL2] = atol * atol;
   // pageable source: the copy is staged before the call returns
   FEMB_CUDA(cudaMemcpyAsync(d_scal, h, sizeof(h), cudaMemcpyHostToDevice, as_stream(stream)));
   return 0;
}

extern "C" int femb200_cg_init(int64_t n, const double *d_b, const double *d_dinv, double *d_x, double *d_r,
                               double *d_dir, double *d_scal, void *stream)
{
   FEMB_CHECK(d_b && d_x && d_r && d_dir && d_scal && n >= 0, "cg_init: bad argument");
   cudaStream_t st = as_stream(stream);
   const unsigned grid = vec_grid(n);
   ReduceScratch red;
   if (int rc = reduce_scratch(grid, st, &red)) return rc;
   cg_init_kernel<<<grid, kVecThreads, 0, st>>>(n, d_b, d_dinv, d_x, d_r, d_dir, red, d_scal + SC_RED_NOM);
   FEMB_LAUNCH_CHECK();
   return 0;
}

extern "C" int femb200_cg_scalar_step(double *d_scal, int phase, void *stream)
{
   FEMB_CHECK(d_scal && phase >= 0 && phase <= 2, "cg_scalar_step: bad argument");
   cg_scalar_kernel<<<1, 1, 0, as_stream(stream)>>>(d_scal, phase);
   FEMB_LAUNCH_CHECK();
   return 0;
}

extern "C" int femb200_cg_update_xr(int64_t n, double *d_scal, const double *d_dir, const double *d_Ad,
                                    const double *d_dinv, double *d_x, double *d_r, void *stream)
{
   FEMB_CHECK(d_scal && d_dir && d_Ad && d_x && d_r && n >= 0 && (n & 1) == 0, "cg_update_xr: bad argument");
   cudaStream_t st = as_stream(stream);
   const unsigned grid = vec_grid(n / 2);
   ReduceScratch red;
   if (int rc = reduce_scratch(grid, st, &red)) return rc;
   cg_update_xr_kernel<<<grid, kVecThreads, 0, st>>>(
       n / 2, d_scal, reinterpret_cast<const double2 *>(d_dir), reinterpret_cast<const double2 *>(d_Ad),
       reinterpret_cast<const double2 *>(d_dinv), reinterpret_cast<double2 *>(d_x), reinterpret_cast<double2 *>(d_r),
       red, d_scal + SC_RED_BETA);
   FEMB_LAUNCH_CHECK();
   return 0;
}

extern "C" int femb200_cg_update_dir(int64_t n, const double *d_scal, const double *d_r, const double *d_dinv,
                                     double *d_dir, void *stream)
{
   FEMB_CHECK(d_scal && d_r && d_dir && n >= 0 && (n & 1) == 0, "cg_update_dir: bad argument");
   cg_update_dir_kernel<<<vec_grid(n / 2), kVecThreads, 0, as_stream(stream)>>>(
       n / 2, d_scal, reinterpret_cast<const double2 *>(d_r), reinterpret_cast<const double2 *>(d_dinv),
       reinterpret_cast<double2 *>(d_dir));
   FEMB_LAUNCH_CHECK();
   return 0;
}

// z = A d with the fused partial <d, A d> into scal[SC_RED_DEN]; no-op once the flag is set
extern "C" int femb200_cg_apply(const femb200_plan *plan, int op_kind, const void *op, const double *d_values,
                                const double *d_dir, double *d_Ad, double *d_scal, void *stream)
{
   FEMB_CHECK(d_dir && d_Ad && d_scal, "cg_apply: null argument");
   cudaStream_t st = as_stream(stream);
   if (op_kind == FEMB200_OP_CSR)
   {
      FEMB_CHECK(plan && d_values, "cg_apply: CSR operator needs a plan and values");
      return spmv_launch(plan, d_values, d_dir, d_Ad, d_scal + SC_FLAG, d_scal + SC_RED_DEN, st);
   }
   FEMB_CHECK(op_kind == FEMB200_OP_PA && op, "cg_apply: unknown operator kind %d", op_kind);
   return pa_apply_launch(static_cast<const femb200_pa *>(op), d_dir, d_Ad, d_scal + SC_FLAG, d_scal + SC_RED_DEN, st);
}

extern "C" int femb200_pcg(const femb200_plan *plan, int op_kind, const void *op, const double *d_values,
                           const double *d_b, double *d_x, int64_t n, double rtol, double atol, int maxit,
                           const double *d_dinv, int check_every, int fixed_iters, double *d_work, int *iters,
                           double *final_norm, int *converged, void *stream)
{
   FEMB_CHECK(d_b && d_x && d_work && n > 0 && (n & 1) == 0, "pcg: bad argument");
   FEMB_CHECK(maxit >= 0, "pcg: negative maxit");
   cudaStream_t st = as_stream(stream);
   double *r = d_work, *dir = d_work + n, *z = d_work + 2 * n, *scal = d_work + 3 * n;
   if (check_every <= 0) check_every = 25;
   int rc;
   if ((rc = femb200_cg_set_tolerances(scal, rtol, atol, stream))) return rc;
   if ((rc = femb200_cg_init(n, d_b, d_dinv, d_x, r, dir, scal, stream))) return rc;
   if ((rc = femb200_cg_scalar_step(scal, 0, stream))) return rc;
   if ((rc = femb200_cg_apply(plan, op_kind, op, d_values, dir, z, scal, stream))) return rc;
   if ((rc = femb200_cg_scalar_step(scal, 1, stream))) return rc;
   const int nit = fixed_iters > 0 ? fixed_iters : maxit;
   double hs[SC_COUNT];
   bool stopped = false;
   for (int i = 1; i <= nit && !stopped; ++i)
   {
      if ((rc = femb200_cg_update_xr(n, scal, dir, z, d_dinv, d_x, r, stream))) return rc;
      if ((rc = femb200_cg_scalar_step(scal, 2, stream))) return rc;
      if (i < nit)
      {
         if ((rc = femb200_cg_update_dir(n, scal, r, d_dinv, dir, stream))) return rc;
         if ((rc = femb200_cg_apply(plan, op_kind, op, d_values, dir, z, scal, stream))) return rc;
         if ((rc = femb200_cg_scalar_step(scal, 1, stream))) return rc;
      }
      if (fixed_iters <= 0 && (i % check_every == 0) && i < nit)
      {
         FEMB_CUDA(cudaMemcpyAsync(hs, scal, sizeof(hs), cudaMemcpyDeviceToHost, st));
         FEMB_CUDA(cudaStreamSynchronize(st));
         stopped = hs[SC_FLAG] != 0.;
      }
   }
   FEMB_CUDA(cudaMemcpyAsync(hs, scal, sizeof(hs), cudaMemcpyDeviceToHost, st));
   FEMB_CUDA(cudaStreamSynchronize(st));
   const bool conv = hs[SC_FLAG] == 1.;
   if (converged) *converged = conv ? 1 : 0;
   if (iters) *iters = conv ? (int)hs[SC_ITERS] : (fixed_iters > 0 ? (int)hs[SC_ITERS] : maxit);
   if (final_norm) *final_norm = sqrt(hs[SC_FINAL] > 0. ? hs[SC_FINAL] : 0.);
   return 0;
}

// dst[k] = src[idx[k]] over node pairs (16-byte items): packs the interface dofs
// of a halo message (role of the dolfinx Scatterer pack step behind
// VecGhostUpdate(INSERT, FORWARD), F.cc:865-866)
namespace femb {
__global__ void gather_kernel(int64_t n, const int32_t *__restrict__ idx, const double2 *__restrict__ src,
                              double2 *__restrict__ dst)
{
   const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
   if (k < n) dst[k] = src[idx[k]];
}
}  // namespace femb

extern "C" int femb200_gather(int64_t nnodes_out, const int32_t *d_node_idx, const double *d_src, double *d_dst,
                              void *stream)
{
   FEMB_CHECK(nnodes_out >= 0 && (nnodes_out == 0 || (d_node_idx && d_src && d_dst)), "gather: bad argument");
   if (nnodes_out == 0) return 0;
   femb::gather_kernel<<<(unsigned)cdiv(nnodes_out, 256), 256, 0, as_stream(stream)>>>(
       nnodes_out, d_node_idx, reinterpret_cast<const double2 *>(d_src), reinterpret_cast<double2 *>(d_dst));
   FEMB_LAUNCH_CHECK();
   return 0;
}

// dst[idx[k]] = src[k] over nodes of `width` doubles (2 or 3): uploads the coordinates of the
// geometry vertices only (dolfinx keeps the P1 geometry apart from the P2 space: mesh.geometry.x
// holds the vertices, F.cc:213), the edge nodes of a straight-sided P2 mesh are never read
namespace femb {
__global__ void scatter_rows_kernel(int64_t n, int width, const int32_t *__restrict__ idx, const double *__restrict__ src,
                                    double *__restrict__ dst)
{
   const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
   if (k >= n) return;
   const int64_t j = idx[k];
   for (int c = 0; c < width; ++c) dst[j * width + c] = src[k * width + c];
}
}  // namespace femb

extern "C" int femb200_scatter_rows(int64_t n, int width, const int32_t *d_idx, const double *d_src, double *d_dst,
                                    void *stream)
{
   FEMB_CHECK(n >= 0 && width >= 1 && width <= 3 && (n == 0 || (d_idx && d_src && d_dst)), "scatter_rows: bad argument");
   if (n == 0) return 0;
   femb::scatter_rows_kernel<<<(unsigned)cdiv(n, 256), 256, 0, as_stream(stream)>>>(n, width, d_idx, d_src, d_dst);
   FEMB_LAUNCH_CHECK();
   return 0;
}
