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

// ---- the ufcx kernel signature -----------------------------------------------------------------
// `ufcx_tabulate_tensor_float64` (ufcx.h of ffcx 0.8, un-vendored; corroborated by the generated-kernel patch
// FEniCSx/mechanic2d/addprofile:6-9 and by the form look-ups F.cc:31-67): what a dolfinx Form holds as its cell
// kernel for the bilinear form J of manual.py:102 and calls once per cell.  Batch of one with the device staging
// inside: A (6 x 6 row-major, interleaved dofs, caller-owned, pre-zeroed by the caller) is ACCUMULATED into, as
// ffcx kernels do; no return value (ufcx has no error channel: on a CUDA failure A is left untouched and the
// message is in femb200_last_error()).  Coefficient packing as the form file creates them (manual.py:19,22,30):
//   w = [d0, d1, d2,  E,  u0x, u0y, u1x, u1y, u2x, u2y],  c = [nu],  coordinate_dofs = 3 x (x, y, z).
// Re-entrant: the device scratch is per host thread.  This is the parity surface of the FEniCSx side; the fast path
// is femb200_tabulate_tensor_batched / femb200_assemble_matrix.
namespace femb {
struct UfcxScratch
{
   double *d = nullptr;   // x[9] | E[1] | dnod[3] | u[6] | A[36]
   int32_t *map = nullptr;
};
static UfcxScratch *ufcx_scratch()
{
   static thread_local UfcxScratch s;
   if (!s.d)
   {
      const int32_t id[3] = {0, 1, 2};
      if (cudaMalloc(&s.d, sizeof(double) * 55) != cudaSuccess || cudaMalloc(&s.map, sizeof(id)) != cudaSuccess ||
          cudaMemcpy(s.map, id, sizeof(id), cudaMemcpyHostToDevice) != cudaSuccess)
      {
         set_error("tabulate_tensor_ufcx: device scratch: %s", cudaGetErrorString(cudaGetLastError()));
         cudaFree(s.d), cudaFree(s.map);
         s.d = nullptr, s.map = nullptr;
         return nullptr;
      }
   }
   return &s;
}
}  // namespace femb

extern "C" void femb200_tabulate_tensor_ufcx(double *A, const double *w, const double *c, const double *coordinate_dofs,
                                             const int *entity_local_index, const uint8_t *quadrature_permutation)
{
   (void)entity_local_index, (void)quadrature_permutation;  // cell integrals use neither
   if (!A || !w || !c || !coordinate_dofs)
   {
      set_error("tabulate_tensor_ufcx: null argument");
      return;
   }
   UfcxScratch *s = ufcx_scratch();
   if (!s) return;
   double h[19];
   for (int i = 0; i < 9; ++i) h[i] = coordinate_dofs[i];
   h[9] = w[3];
   for (int i = 0; i < 3; ++i) h[10 + i] = w[i];
   for (int i = 0; i < 6; ++i) h[13 + i] = w[4 + i];
   const bool damaged = w[0] != 0. || w[1] != 0. || w[2] != 0.;
   if (cudaMemcpy(s->d, h, sizeof(h), cudaMemcpyHostToDevice) != cudaSuccess)
   {
      set_error("tabulate_tensor_ufcx: H2D: %s", cudaGetErrorString(cudaGetLastError()));
      return;
   }
   if (femb200_tabulate_tensor_batched(FEMB200_P1, 1, s->d + 19, s->d, 3, s->map, s->map, s->d + 9, c[0],
                                       damaged ? s->d + 10 : nullptr, s->d + 13, FEMB200_TANGENT_CLOSED,
                                       FEMB200_ROWMAJOR_INTERLEAVED, nullptr))
      return;
   double out[36];
   if (cudaMemcpy(out, s->d + 19, sizeof(out), cudaMemcpyDeviceToHost) != cudaSuccess)
   {
      set_error("tabulate_tensor_ufcx: D2H: %s", cudaGetErrorString(cudaGetLastError()));
      return;
   }
   for (int i = 0; i < 36; ++i) A[i] += out[i];
}
