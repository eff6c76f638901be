// pa.cu -- partial assembly: matrix-free apply of the elasticity operator.
//
// Role in the reference: mfem BilinearFormIntegrator::AssemblePA / AddMultPA.  The
// reference does not exercise them (its integrator is a NonlinearFormIntegrator,
// M.cc:559,639; partial assembly is only discussed at doc.tex:1445-1449), so the
// contract is SURVEY.md A.9:  y_e = sum_q w_q |det J_q| B_q (D (B_q^t x_e)).
//
// "AssemblePA" stores per cell the straight-sided geometry (nv vertices) and the
// Lame pair: 10 doubles (Q2) / 8 doubles (triangles), read contiguously; J_q is
// rebuilt per quadrature point (bilinear / affine, a handful of FMAs), which is
// fewer bytes than storing (J^-1, w) per point.  "AddMultPA": one thread per
// cell, E-vector gathered with 16-byte loads, sum-factorised contractions for Q2
// (1-D 3x3 tables), results added with fp64 red.global.add.  Dirichlet dofs are
// handled like the assembled operator (rows/cols zeroed, diag on the diagonal)
// through a per-cell bit mask of constrained local dofs.
#include <algorithm>

#include "constitutive.cuh"
#include "element.cuh"
#include "plan.cuh"
#include "reduce.cuh"

struct femb200_pa
{
   int etype = 0, nd = 0, nv = 0;
   int64_t nnodes = 0, ncells = 0;
   const int32_t *dofmap = nullptr;  // borrowed
   double *geo = nullptr;            // [ncells][2 nv + 2]: vertices, lambda, mu
   uint32_t *cmask = nullptr;        // [ncells] constrained local dofs (bit 2a+i) or null
   uint8_t *bc = nullptr;            // [2 nnodes] or null
   int32_t *bc_dofs = nullptr;       // compact list of constrained dofs
   int32_t nbc = 0;
   double diag = 1.0;
};

namespace femb {

constexpr int kPaThreads = 128;

template <int ET>
__global__ void pa_setup_kernel(int64_t ncells, const int32_t *__restrict__ xdofmap, const double *__restrict__ x,
                                int xs, const double *__restrict__ E, LameCoef lc, double *__restrict__ geo)
{
   constexpr int nv = Elem<ET>::nv, W = 2 * nv + 2;
   const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
   if (e >= ncells) return;
   double *g = geo + e * W;
#pragma unroll
   for (int v = 0; v < nv; ++v)
   {
      const int64_t n = xdofmap[e * nv + v];
      g[2 * v] = x[n * xs];
      g[2 * v + 1] = x[n * xs + 1];
   }
   g[2 * nv] = E[e] * lc.c2;      // lambda, M.cc:1093-1098
   g[2 * nv + 1] = E[e] * lc.c3;  // mu
}

__global__ void pa_cmask_kernel(int64_t ncells, int nd, const int32_t *__restrict__ dofmap,
                                const uint8_t *__restrict__ bc, uint32_t *__restrict__ cmask)
{
   const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
   if (e >= ncells) return;
   uint32_t m = 0;
   for (int a = 0; a < nd; ++a)
   {
      const int64_t n = dofmap[e * nd + a];
      if (bc[2 * n]) m |= 1u << (2 * a);
      if (bc[2 * n + 1]) m |= 1u << (2 * a + 1);
   }
   cmask[e] = m;
}

__global__ void pa_bc_list_kernel(int64_t ndofs, const uint8_t *__restrict__ bc, int32_t *__restrict__ list,
                                  int32_t *__restrict__ count)
{
   const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
   if (i >= ndofs || !bc[i]) return;
   const int32_t p = atomicAdd(count, 1);
   if (list) list[p] = (int32_t)i;
}

// stress at one point: sg = w D (exx, eyy, gxy), Hooke (M.cc:873-881)
__device__ __forceinline__ void hooke_stress(double lam, double mu, double w, double exx, double eyy, double gxy,
                                             double &sxx, double &syy, double &sxy)
{
   const double tr = lam * (exx + eyy);
   sxx = w * (tr + 2. * mu * exx);
   syy = w * (tr + 2. * mu * eyy);
   sxy = w * (mu * gxy);
}

// element-local product ye = K_e xe for triangles: generic loop over the points
template <int ET>
__device__ __forceinline__ void local_apply_tri(const double *g, const double *ux, const double *uy, double *yx,
                                                double *yy)
{
   constexpr int nd = Elem<ET>::nd, nv = Elem<ET>::nv, nq = Elem<ET>::nq;
   double xv[nv][2];
#pragma unroll
   for (int v = 0; v < nv; ++v) xv[v][0] = g[2 * v], xv[v][1] = g[2 * v + 1];
   const double lam = g[2 * nv], mu = g[2 * nv + 1];
#pragma unroll
   for (int a = 0; a < nd; ++a) yx[a] = yy[a] = 0.;
#pragma unroll
   for (int q = 0; q < nq; ++q)
   {
      double G[nd][2], phi[nv];
      const double w = qp_geometry<ET>(xv, q, G, phi);
      double exx = 0., eyy = 0., gxy = 0.;
#pragma unroll
      for (int a = 0; a < nd; ++a)
      {
         exx += G[a][0] * ux[a];
         eyy += G[a][1] * uy[a];
         gxy += G[a][1] * ux[a] + G[a][0] * uy[a];
      }
      double sxx, syy, sxy;
      hooke_stress(lam, mu, w, exx, eyy, gxy, sxx, syy, sxy);
#pragma unroll
      for (int a = 0; a < nd; ++a)
      {
         yx[a] += G[a][0] * sxx + G[a][1] * sxy;
         yy[a] += G[a][1] * syy + G[a][0] * sxy;
      }
   }
}

// Q2: sum factorisation with the 1-D tables of the three Lagrange (GLL) basis
// functions at the three Gauss points
__device__ __forceinline__ void q2_tables(double (*B)[3], double (*dB)[3])
{
   const double s = 0.7745966692414834;
   const double xq[3] = {0.5 * (1. - s), 0.5, 0.5 * (1. + s)};
#pragma unroll
   for (int q = 0; q < 3; ++q)
   {
      const double t = xq[q];
      B[q][0] = 2. * (t - 0.5) * (t - 1.), B[q][1] = 4. * t * (1. - t), B[q][2] = 2. * t * (t - 0.5);
      dB[q][0] = 4. * t - 3., dB[q][1] = 4. - 8. * t, dB[q][2] = 4. * t - 1.;
   }
}

__device__ __forceinline__ void local_apply_q2(const double *g, const double *ux, const double *uy, double *yx,
                                               double *yy)
{
   double B[3][3], dB[3][3];
   q2_tables(B, dB);
   const double s = 0.7745966692414834;
   const double xq[3] = {0.5 * (1. - s), 0.5, 0.5 * (1. + s)};
   const double wq[3] = {5. / 18., 8. / 18., 5. / 18.};
   const double lam = g[8], mu = g[9];
   // reference gradients at the 9 points: a[q][0..3] = (dux/dxi, dux/deta, duy/dxi, duy/deta)
   double a[9][4];
   {
      double t0[3][3], t1[3][3];  // [qx][j]
#pragma unroll
      for (int c = 0; c < 2; ++c)
      {
         const double *u = c ? uy : ux;
#pragma unroll
         for (int qx = 0; qx < 3; ++qx)
#pragma unroll
            for (int j = 0; j < 3; ++j)
            {
               t0[qx][j] = B[qx][0] * u[3 * j] + B[qx][1] * u[3 * j + 1] + B[qx][2] * u[3 * j + 2];
               t1[qx][j] = dB[qx][0] * u[3 * j] + dB[qx][1] * u[3 * j + 1] + dB[qx][2] * u[3 * j + 2];
            }
#pragma unroll
         for (int qy = 0; qy < 3; ++qy)
#pragma unroll
            for (int qx = 0; qx < 3; ++qx)
            {
               a[3 * qy + qx][2 * c] = t1[qx][0] * B[qy][0] + t1[qx][1] * B[qy][1] + t1[qx][2] * B[qy][2];
               a[3 * qy + qx][2 * c + 1] = t0[qx][0] * dB[qy][0] + t0[qx][1] * dB[qy][1] + t0[qx][2] * dB[qy][2];
            }
      }
   }
   // point work: a[q] <- (px0, px1, py0, py1) = J^-1 (w sigma) rows
#pragma unroll
   for (int qy = 0; qy < 3; ++qy)
#pragma unroll
      for (int qx = 0; qx < 3; ++qx)
      {
         const double xi = xq[qx], eta = xq[qy];
         // bilinear geometry, vertices (0,0),(1,0),(0,1),(1,1)
         const double J00 = (1. - eta) * (g[2] - g[0]) + eta * (g[6] - g[4]);
         const double J01 = (1. - xi) * (g[4] - g[0]) + xi * (g[6] - g[2]);
         const double J10 = (1. - eta) * (g[3] - g[1]) + eta * (g[7] - g[5]);
         const double J11 = (1. - xi) * (g[5] - g[1]) + xi * (g[7] - g[3]);
         const double det = J00 * J11 - J01 * J10;
         const double id = 1. / det;
         const double i00 = J11 * id, i01 = -J01 * id, i10 = -J10 * id, i11 = J00 * id;  // J^-1[m][k]
         double *aq = a[3 * qy + qx];
         // physical gradient: d u / d x_k = sum_m (du/dxi_m) Jinv[m][k]
         const double uxx = aq[0] * i00 + aq[1] * i10, uxy = aq[0] * i01 + aq[1] * i11;
         const double uyx = aq[2] * i00 + aq[3] * i10, uyy = aq[2] * i01 + aq[3] * i11;
         double sxx, syy, sxy;
         hooke_stress(lam, mu, wq[qx] * wq[qy] * fabs(det), uxx, uyy, uxy + uyx, sxx, syy, sxy);
         // back to reference directions: p_m = sum_k Jinv[m][k] sigma_{c k}
         aq[0] = i00 * sxx + i01 * sxy;
         aq[1] = i10 * sxx + i11 * sxy;
         aq[2] = i00 * sxy + i01 * syy;
         aq[3] = i10 * sxy + i11 * syy;
      }
   // transpose contractions
#pragma unroll
   for (int c = 0; c < 2; ++c)
   {
      double *y = c ? yy : yx;
      double t0[3][3], t1[3][3];  // [qx][j]: sum over qy
#pragma unroll
      for (int qx = 0; qx < 3; ++qx)
#pragma unroll
         for (int j = 0; j < 3; ++j)
         {
            t1[qx][j] = B[0][j] * a[qx][2 * c] + B[1][j] * a[3 + qx][2 * c] + B[2][j] * a[6 + qx][2 * c];
            t0[qx][j] = dB[0][j] * a[qx][2 * c + 1] + dB[1][j] * a[3 + qx][2 * c + 1] + dB[2][j] * a[6 + qx][2 * c + 1];
         }
#pragma unroll
      for (int j = 0; j < 3; ++j)
#pragma unroll
         for (int i = 0; i < 3; ++i)
            y[3 * j + i] = dB[0][i] * t1[0][j] + dB[1][i] * t1[1][j] + dB[2][i] * t1[2][j] + B[0][i] * t0[0][j] +
                           B[1][i] * t0[1][j] + B[2][i] * t0[2][j];
   }
}

__device__ __forceinline__ void red_add_f64(double *p, double v)
{
   asm volatile("red.global.add.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory");
}

template <int ET, bool DOT>
__global__ void __launch_bounds__(kPaThreads)
pa_apply_kernel(int64_t ncells, const int32_t *__restrict__ dofmap, const double *__restrict__ geo,
                const uint32_t *__restrict__ cmask, const double *__restrict__ x, double *__restrict__ y,
                const double *__restrict__ flag, ReduceScratch red, double *__restrict__ out)
{
   constexpr int nd = Elem<ET>::nd, nv = Elem<ET>::nv, W = 2 * nv + 2;
   if (flag && *flag != 0.) return;
   const int64_t e = (int64_t)blockIdx.x * kPaThreads + threadIdx.x;
   double part = 0.;
   if (e < ncells)
   {
      int32_t dof[nd];
#pragma unroll
      for (int a = 0; a < nd; ++a) dof[a] = dofmap[e * nd + a];
      const uint32_t m = cmask ? cmask[e] : 0u;
      double g[W];
      const double2 *g2 = reinterpret_cast<const double2 *>(geo + e * W);
#pragma unroll
      for (int k = 0; k < W / 2; ++k)
      {
         const double2 v = g2[k];
         g[2 * k] = v.x, g[2 * k + 1] = v.y;
      }
      double ux[nd], uy[nd], yx[nd], yy[nd];
      const double2 *x2 = reinterpret_cast<const double2 *>(x);
#pragma unroll
      for (int a = 0; a < nd; ++a)
      {
         const double2 v = x2[dof[a]];
         ux[a] = ((m >> (2 * a)) & 1u) ? 0. : v.x;
         uy[a] = ((m >> (2 * a + 1)) & 1u) ? 0. : v.y;
      }
      if (ET == FEMB200_Q2)
         local_apply_q2(g, ux, uy, yx, yy);
      else
         local_apply_tri<ET>(g, ux, uy, yx, yy);
#pragma unroll
      for (int a = 0; a < nd; ++a)
      {
         double *yp = y + 2 * (int64_t)dof[a];
         if (!((m >> (2 * a)) & 1u))
         {
            red_add_f64(yp, yx[a]);
            if (DOT) part += ux[a] * yx[a];
         }
         if (!((m >> (2 * a + 1)) & 1u))
         {
            red_add_f64(yp + 1, yy[a]);
            if (DOT) part += uy[a] * yy[a];
         }
      }
   }
   if (DOT) block_reduce_finish<kPaThreads>(part, red, out);
}

// y[bc] = diag x[bc]; adds diag x[bc]^2 to the fused dot.  One CTA.
__global__ void __launch_bounds__(256)
pa_bc_fix_kernel(int nbc, const int32_t *__restrict__ list, double diag, const double *__restrict__ x,
                 double *__restrict__ y, const double *__restrict__ flag, double *__restrict__ out)
{
   __shared__ double sh[8];
   if (flag && *flag != 0.) return;
   double part = 0.;
   for (int k = threadIdx.x; k < nbc; k += 256)
   {
      const int64_t i = list[k];
      const double xi = x[i];
      y[i] = diag * xi;
      part += diag * xi * xi;
   }
   if (out)
   {
      const double t = block_sum<256>(part, sh);
      if (threadIdx.x == 0) *out += t;
   }
}

template <int ET>
__global__ void __launch_bounds__(kPaThreads)
pa_diag_kernel(int64_t ncells, const int32_t *__restrict__ dofmap, const double *__restrict__ geo,
               const uint32_t *__restrict__ cmask, double *__restrict__ diag)
{
   constexpr int nd = Elem<ET>::nd, nv = Elem<ET>::nv, nq = Elem<ET>::nq, W = 2 * nv + 2;
   const int64_t e = (int64_t)blockIdx.x * kPaThreads + threadIdx.x;
   if (e >= ncells) return;
   const double *g = geo + e * W;
   double xv[nv][2];
#pragma unroll
   for (int v = 0; v < nv; ++v) xv[v][0] = g[2 * v], xv[v][1] = g[2 * v + 1];
   double D[9];
   hooke_scaled(g[2 * nv], g[2 * nv + 1], 1., D);
   double kd[nd][2];
#pragma unroll
   for (int a = 0; a < nd; ++a) kd[a][0] = kd[a][1] = 0.;
#pragma unroll 1
   for (int q = 0; q < nq; ++q)
   {
      double G[nd][2], phi[nv];
      const double w = qp_geometry<ET>(xv, q, G, phi);
#pragma unroll
      for (int a = 0; a < nd; ++a)
      {
         double k[4] = {0., 0., 0., 0.};
         bdb_block(G[a], G[a], D, w, k);
         kd[a][0] += k[0];
         kd[a][1] += k[3];
      }
   }
   const uint32_t m = cmask ? cmask[e] : 0u;
#pragma unroll
   for (int a = 0; a < nd; ++a)
   {
      double *dp = diag + 2 * (int64_t)dofmap[e * nd + a];
      if (!((m >> (2 * a)) & 1u)) red_add_f64(dp, kd[a][0]);
      if (!((m >> (2 * a + 1)) & 1u)) red_add_f64(dp + 1, kd[a][1]);
   }
}

__global__ void pa_diag_bc_kernel(int nbc, const int32_t *__restrict__ list, double diag, double *__restrict__ d)
{
   const int k = blockIdx.x * blockDim.x + threadIdx.x;
   if (k < nbc) d[list[k]] = diag;
}

template <int ET>
static int pa_apply_t(const femb200_pa *pa, const double *d_x, double *d_y, const double *d_flag, double *d_dot_out,
                      cudaStream_t st)
{
   const unsigned grid = (unsigned)cdiv(pa->ncells, kPaThreads);
   if (d_dot_out)
   {
      ReduceScratch red;
      if (int rc = reduce_scratch(grid, st, &red)) return rc;
      pa_apply_kernel<ET, true><<<grid, kPaThreads, 0, st>>>(pa->ncells, pa->dofmap, pa->geo, pa->cmask, d_x, d_y,
                                                              d_flag, red, d_dot_out);
   }
   else
      pa_apply_kernel<ET, false><<<grid, kPaThreads, 0, st>>>(pa->ncells, pa->dofmap, pa->geo, pa->cmask, d_x, d_y,
                                                               d_flag, ReduceScratch{nullptr, nullptr}, nullptr);
   FEMB_LAUNCH_CHECK();
   return 0;
}

int pa_apply_launch(const femb200_pa *pa, const double *d_x, double *d_y, const double *d_flag, double *d_dot_out,
                    cudaStream_t st)
{
   // y = 0 (a converged CG leaves y stale: harmless, nothing reads it any more)
   FEMB_CUDA(cudaMemsetAsync(d_y, 0, sizeof(double) * 2 * (size_t)pa->nnodes, st));
   int rc;
   switch (pa->etype)
   {
      case FEMB200_P1: rc = pa_apply_t<FEMB200_P1>(pa, d_x, d_y, d_flag, d_dot_out, st); break;
      case FEMB200_P2: rc = pa_apply_t<FEMB200_P2>(pa, d_x, d_y, d_flag, d_dot_out, st); break;
      default: rc = pa_apply_t<FEMB200_Q2>(pa, d_x, d_y, d_flag, d_dot_out, st);
   }
   if (rc) return rc;
   if (pa->nbc > 0)
   {
      pa_bc_fix_kernel<<<1, 256, 0, st>>>(pa->nbc, pa->bc_dofs, pa->diag, d_x, d_y, d_flag, d_dot_out);
      FEMB_LAUNCH_CHECK();
   }
   return 0;
}

}  // namespace femb

using namespace femb;

extern "C" void femb200_pa_destroy(femb200_pa *pa)
{
   if (!pa) return;
   cudaFree(pa->geo);
   cudaFree(pa->cmask);
   cudaFree(pa->bc);
   cudaFree(pa->bc_dofs);
   delete pa;
}

extern "C" int femb200_pa_create(int etype, int64_t nnodes, int64_t ncells, const int32_t *d_dofmap,
                                 const int32_t *d_xdofmap, const double *d_x, int x_stride, const double *d_E, double nu,
                                 void *stream, femb200_pa **out)
{
   FEMB_CHECK(out != nullptr, "pa_create: out is null");
   *out = nullptr;
   FEMB_CHECK(etype >= FEMB200_P1 && etype <= FEMB200_Q2, "pa_create: unknown element family %d", etype);
   FEMB_CHECK(nnodes > 0 && ncells > 0, "pa_create: empty mesh");
   FEMB_CHECK(d_dofmap && d_xdofmap && d_x && d_E, "pa_create: null argument");
   FEMB_CHECK(x_stride == 2 || x_stride == 3, "pa_create: x_stride must be 2 or 3, got %d", x_stride);
   femb200_pa *pa = new femb200_pa();
   pa->etype = etype, pa->nd = elem_nd(etype), pa->nv = elem_nv(etype);
   pa->nnodes = nnodes, pa->ncells = ncells, pa->dofmap = d_dofmap;
   const size_t W = 2 * (size_t)pa->nv + 2;
   if (cudaMalloc(&pa->geo, sizeof(double) * W * (size_t)ncells) != cudaSuccess)
   {
      femb200_pa_destroy(pa);
      return set_error("pa_create: cudaMalloc of %zu bytes failed", sizeof(double) * W * (size_t)ncells);
   }
   const LameCoef lc = lame_coef(nu);
   const unsigned grid = (unsigned)cdiv(ncells, 256);
   cudaStream_t st = as_stream(stream);
   if (etype == FEMB200_Q2)
      pa_setup_kernel<FEMB200_Q2><<<grid, 256, 0, st>>>(ncells, d_xdofmap, d_x, x_stride, d_E, lc, pa->geo);
   else
      pa_setup_kernel<FEMB200_P1><<<grid, 256, 0, st>>>(ncells, d_xdofmap, d_x, x_stride, d_E, lc, pa->geo);
   if (cudaGetLastError() != cudaSuccess)
   {
      femb200_pa_destroy(pa);
      return set_error("pa_create: setup launch failed");
   }
   *out = pa;
   return 0;
}

extern "C" int femb200_pa_set_dirichlet(femb200_pa *pa, const uint8_t *d_bc, double diag, void *stream)
{
   FEMB_CHECK(pa != nullptr, "pa_set_dirichlet: null operator");
   cudaStream_t st = as_stream(stream);
   cudaFree(pa->cmask), cudaFree(pa->bc), cudaFree(pa->bc_dofs);
   pa->cmask = nullptr, pa->bc = nullptr, pa->bc_dofs = nullptr, pa->nbc = 0;
   pa->diag = diag;
   if (!d_bc) return 0;
   const int64_t nd = 2 * pa->nnodes;
   FEMB_CUDA(cudaMalloc(&pa->bc, (size_t)nd));
   FEMB_CUDA(cudaMalloc(&pa->cmask, sizeof(uint32_t) * (size_t)pa->ncells));
   FEMB_CUDA(cudaMemcpyAsync(pa->bc, d_bc, (size_t)nd, cudaMemcpyDeviceToDevice, st));
   pa_cmask_kernel<<<(unsigned)cdiv(pa->ncells, 256), 256, 0, st>>>(pa->ncells, pa->nd, pa->dofmap, pa->bc, pa->cmask);
   int32_t *count = nullptr;
   FEMB_CUDA(cudaMalloc(&count, sizeof(int32_t)));
   FEMB_CUDA(cudaMemsetAsync(count, 0, sizeof(int32_t), st));
   pa_bc_list_kernel<<<(unsigned)cdiv(nd, 256), 256, 0, st>>>(nd, pa->bc, nullptr, count);
   int32_t n = 0;
   cudaMemcpyAsync(&n, count, sizeof(int32_t), cudaMemcpyDeviceToHost, st);
   if (cudaStreamSynchronize(st) != cudaSuccess)
   {
      cudaFree(count);
      return set_error("pa_set_dirichlet: %s", cudaGetErrorString(cudaGetLastError()));
   }
   cudaMalloc(&pa->bc_dofs, sizeof(int32_t) * (size_t)(n ? n : 1));
   cudaMemsetAsync(count, 0, sizeof(int32_t), st);
   pa_bc_list_kernel<<<(unsigned)cdiv(nd, 256), 256, 0, st>>>(nd, pa->bc, pa->bc_dofs, count);
   cudaStreamSynchronize(st);
   cudaFree(count);
   pa->nbc = n;
   FEMB_LAUNCH_CHECK();
   return 0;
}

extern "C" int femb200_pa_apply(const femb200_pa *pa, const double *d_x, double *d_y, void *stream)
{
   FEMB_CHECK(pa && d_x && d_y, "pa_apply: null argument");
   FEMB_CHECK(d_x != d_y, "pa_apply: x and y must not alias");
   return pa_apply_launch(pa, d_x, d_y, nullptr, nullptr, as_stream(stream));
}

extern "C" int femb200_pa_diagonal(const femb200_pa *pa, double *d_diag, void *stream)
{
   FEMB_CHECK(pa && d_diag, "pa_diagonal: null argument");
   cudaStream_t st = as_stream(stream);
   FEMB_CUDA(cudaMemsetAsync(d_diag, 0, sizeof(double) * 2 * (size_t)pa->nnodes, st));
   const unsigned grid = (unsigned)cdiv(pa->ncells, kPaThreads);
   switch (pa->etype)
   {
      case FEMB200_P1:
         pa_diag_kernel<FEMB200_P1><<<grid, kPaThreads, 0, st>>>(pa->ncells, pa->dofmap, pa->geo, pa->cmask, d_diag);
         break;
      case FEMB200_P2:
         pa_diag_kernel<FEMB200_P2><<<grid, kPaThreads, 0, st>>>(pa->ncells, pa->dofmap, pa->geo, pa->cmask, d_diag);
         break;
      default:
         pa_diag_kernel<FEMB200_Q2><<<grid, kPaThreads, 0, st>>>(pa->ncells, pa->dofmap, pa->geo, pa->cmask, d_diag);
   }
   FEMB_LAUNCH_CHECK();
   if (pa->nbc > 0)
   {
      pa_diag_bc_kernel<<<(unsigned)cdiv(pa->nbc, 256), 256, 0, st>>>(pa->nbc, pa->bc_dofs, pa->diag, d_diag);
      FEMB_LAUNCH_CHECK();
   }
   return 0;
}
