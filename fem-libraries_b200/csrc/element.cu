// element.cu -- batched element tangent kernel.
//
// Drop-in for the per-cell ufcx `tabulate_tensor` of form J (manual.py:102) and
// for damIntegrator::AssembleElementGrad (M.cc:639-916): one launch tabulates
// every cell.  K_e = sum_q w_q |det J_q| B_q D_q B_q^t  (SURVEY.md A.1, A.9).
#include "constitutive.cuh"
#include "element.cuh"

namespace femb {

template <int ET>
__global__ void __launch_bounds__(128)
tabulate_kernel(int64_t ncells, double *__restrict__ A, const double *__restrict__ x, int xs,
                const int32_t *__restrict__ xdofmap, const int32_t *__restrict__ dofmap, const double *__restrict__ E,
                LameCoef lc, const double *__restrict__ dnod, const double *__restrict__ u, int variant, int layout)
{
   constexpr int nd = Elem<ET>::nd, nv = Elem<ET>::nv, nq = Elem<ET>::nq, n = 2 * nd;
   const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
   if (e >= ncells) return;

   double xv[nv][2], dv[nv];
#pragma unroll
   for (int v = 0; v < nv; ++v)
   {
      const int64_t g = xdofmap[e * nv + v];
      xv[v][0] = x[g * xs];
      xv[v][1] = x[g * xs + 1];
      dv[v] = dnod ? dnod[g] : 0.;
   }
   const double Ee = E[e];
   const double lam = Ee * lc.c2, mu = Ee * lc.c3;  // M.cc:1093-1098

   double G[nq][nd][2], w[nq], D[nq][9];
#pragma unroll
   for (int q = 0; q < nq; ++q)
   {
      double phi[nv];
      w[q] = qp_geometry<ET>(xv, q, G[q], phi);
      double d = 0.;
#pragma unroll
      for (int v = 0; v < nv; ++v) d += phi[v] * dv[v];
      if (d > 0.)
      {
         double g00 = 0., g01 = 0., g10 = 0., g11 = 0.;  // grad u (M.cc:742)
         if (u)
#pragma unroll
            for (int a = 0; a < nd; ++a)
            {
               const int64_t gd = 2 * (int64_t)dofmap[e * nd + a];
               const double ux = u[gd], uy = u[gd + 1];
               g00 += ux * G[q][a][0];
               g01 += ux * G[q][a][1];
               g10 += uy * G[q][a][0];
               g11 += uy * G[q][a][1];
            }
         const double s = 0.5 * (g01 + g10);
         const double eps[4] = {g00, s, s, g11};
         tangent(variant, lam, mu, d, eps, D[q]);
      }
      else
         hooke_scaled(lam, mu, 1., D[q]);
   }

   double *Ae = A + e * (int64_t)(n * n);
#pragma unroll 1
   for (int a = 0; a < nd; ++a)
   {
#pragma unroll 1
      for (int b = 0; b < nd; ++b)
      {
         double k[4] = {0., 0., 0., 0.};
#pragma unroll
         for (int q = 0; q < nq; ++q) bdb_block(G[q][a], G[q][b], D[q], w[q], k);
         if (layout == FEMB200_ROWMAJOR_INTERLEAVED)
         {
            Ae[(2 * a) * n + 2 * b] = k[0];
            Ae[(2 * a) * n + 2 * b + 1] = k[1];
            Ae[(2 * a + 1) * n + 2 * b] = k[2];
            Ae[(2 * a + 1) * n + 2 * b + 1] = k[3];
         }
         else
         {
            Ae[a + b * n] = k[0];
            Ae[a + (nd + b) * n] = k[1];
            Ae[(nd + a) + b * n] = k[2];
            Ae[(nd + a) + (nd + b) * n] = k[3];
         }
      }
   }
}

}  // namespace femb

using namespace femb;

extern "C" int femb200_tabulate_tensor_batched(int etype, int64_t ncells, double *d_A, const double *d_x, int x_stride,
                                               const int32_t *d_xdofmap, const int32_t *d_dofmap, const double *d_E,
                                               double nu, const double *d_dnod, const double *d_u, int variant,
                                               int layout, void *stream)
{
   FEMB_CHECK(etype >= FEMB200_P1 && etype <= FEMB200_Q2, "tabulate: unknown element family %d", etype);
   FEMB_CHECK(x_stride == 2 || x_stride == 3, "tabulate: x_stride must be 2 or 3, got %d", x_stride);
   FEMB_CHECK(d_A && d_x && d_xdofmap && d_dofmap && d_E, "tabulate: null pointer argument");
   FEMB_CHECK(layout == FEMB200_ROWMAJOR_INTERLEAVED || layout == FEMB200_COLMAJOR_BYNODES, "tabulate: bad layout %d",
              layout);
   if (ncells <= 0) return 0;
   const LameCoef lc = lame_coef(nu);
   const unsigned grid = (unsigned)cdiv(ncells, 128);
   cudaStream_t st = as_stream(stream);
   switch (etype)
   {
      case FEMB200_P1:
         tabulate_kernel<FEMB200_P1><<<grid, 128, 0, st>>>(ncells, d_A, d_x, x_stride, d_xdofmap, d_dofmap, d_E, lc,
                                                         d_dnod, d_u, variant, layout);
         break;
      case FEMB200_P2:
         tabulate_kernel<FEMB200_P2><<<grid, 128, 0, st>>>(ncells, d_A, d_x, x_stride, d_xdofmap, d_dofmap, d_E, lc,
                                                         d_dnod, d_u, variant, layout);
         break;
      default:
         tabulate_kernel<FEMB200_Q2><<<grid, 128, 0, st>>>(ncells, d_A, d_x, x_stride, d_xdofmap, d_dofmap, d_E, lc,
                                                         d_dnod, d_u, variant, layout);
   }
   FEMB_LAUNCH_CHECK();
   return 0;
}
