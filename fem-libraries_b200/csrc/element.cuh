// element.cuh -- reference elements, quadrature and per-point geometry for the
// P1 / P2 triangle and the Q2 quadrilateral (conventions: SURVEY.md 8c).
//
//   P1  vertices (0,0),(1,0),(0,1); gradients [[-1,-1],[1,0],[0,1]] == CalcDShape
//       of the reference (M.cc:692); one point (1/3,1/3), weight 1/2
//       (M.cc:1151-1152, manual.py:97).
//   P2  basix order: 3 vertices then the midpoint of the edge opposite vertex i;
//       3-point degree-2 rule, weights 1/6.
//   Q2  tensor Gauss-Lobatto nodes 0,1/2,1, local index ix + 3 iy; 3x3 Gauss.
// Physical gradients G = dN . J^-1 (M.cc:696), weight w_q |det J| (M.cc:684-685,
// |.| because square.msh is clockwise).
#pragma once
#include "common.cuh"

namespace femb {

template <int ET>
struct Elem
{
   static constexpr int nd = elem_nd(ET), nv = elem_nv(ET), nq = elem_nq(ET);
};

// quadrature point q of element family ET on the reference cell
template <int ET>
__device__ __forceinline__ void quad_point(int q, double &xi, double &eta, double &w)
{
   if (ET == FEMB200_P1)
   {
      xi = eta = 1. / 3.;
      w = 0.5;
   }
   else if (ET == FEMB200_P2)
   {
      xi = (q == 1) ? 2. / 3. : 1. / 6.;
      eta = (q == 2) ? 2. / 3. : 1. / 6.;
      w = 1. / 6.;
   }
   else
   {
      const double s = 0.7745966692414834;  // sqrt(3/5)
      const int i = q % 3, j = q / 3;
      xi = (i == 0) ? 0.5 * (1. - s) : (i == 1 ? 0.5 : 0.5 * (1. + s));
      eta = (j == 0) ? 0.5 * (1. - s) : (j == 1 ? 0.5 : 0.5 * (1. + s));
      const double wi = (i == 1) ? 8. / 18. : 5. / 18., wj = (j == 1) ? 8. / 18. : 5. / 18.;
      w = wi * wj;
   }
}

// reference gradients of the nd scalar shape functions
template <int ET>
__device__ __forceinline__ void ref_grads(double xi, double eta, double (*dN)[2])
{
   if (ET == FEMB200_P1)
   {
      dN[0][0] = -1., dN[0][1] = -1.;
      dN[1][0] = 1., dN[1][1] = 0.;
      dN[2][0] = 0., dN[2][1] = 1.;
   }
   else if (ET == FEMB200_P2)
   {
      const double L0 = 1. - xi - eta, L1 = xi, L2 = eta;
      dN[0][0] = -(4. * L0 - 1.), dN[0][1] = -(4. * L0 - 1.);
      dN[1][0] = (4. * L1 - 1.), dN[1][1] = 0.;
      dN[2][0] = 0., dN[2][1] = (4. * L2 - 1.);
      dN[3][0] = 4. * L2, dN[3][1] = 4. * L1;          // 4 L1 L2
      dN[4][0] = -4. * L2, dN[4][1] = 4. * (L0 - L2);  // 4 L0 L2
      dN[5][0] = 4. * (L0 - L1), dN[5][1] = -4. * L1;  // 4 L0 L1
   }
   else
   {
      const double lx[3] = {2. * (xi - 0.5) * (xi - 1.), 4. * xi * (1. - xi), 2. * xi * (xi - 0.5)};
      const double ly[3] = {2. * (eta - 0.5) * (eta - 1.), 4. * eta * (1. - eta), 2. * eta * (eta - 0.5)};
      const double dx[3] = {4. * xi - 3., 4. - 8. * xi, 4. * xi - 1.};
      const double dy[3] = {4. * eta - 3., 4. - 8. * eta, 4. * eta - 1.};
#pragma unroll
      for (int j = 0; j < 3; ++j)
#pragma unroll
         for (int i = 0; i < 3; ++i)
         {
            dN[3 * j + i][0] = dx[i] * ly[j];
            dN[3 * j + i][1] = lx[i] * dy[j];
         }
   }
}

// values of the nd scalar shape functions
template <int ET>
__device__ __forceinline__ void basis_values(double xi, double eta, double *N)
{
   if (ET == FEMB200_P1)
   {
      N[0] = 1. - xi - eta, N[1] = xi, N[2] = eta;
   }
   else if (ET == FEMB200_P2)
   {
      const double L0 = 1. - xi - eta, L1 = xi, L2 = eta;
      N[0] = L0 * (2. * L0 - 1.), N[1] = L1 * (2. * L1 - 1.), N[2] = L2 * (2. * L2 - 1.);
      N[3] = 4. * L1 * L2, N[4] = 4. * L0 * L2, N[5] = 4. * L0 * L1;
   }
   else
   {
      const double lx[3] = {2. * (xi - 0.5) * (xi - 1.), 4. * xi * (1. - xi), 2. * xi * (xi - 0.5)};
      const double ly[3] = {2. * (eta - 0.5) * (eta - 1.), 4. * eta * (1. - eta), 2. * eta * (eta - 0.5)};
#pragma unroll
      for (int j = 0; j < 3; ++j)
#pragma unroll
         for (int i = 0; i < 3; ++i) N[3 * j + i] = lx[i] * ly[j];
   }
}

// geometry (vertex) basis: values and reference gradients
template <int ET>
__device__ __forceinline__ void geom_basis(double xi, double eta, double *phi, double (*dphi)[2])
{
   if (ET == FEMB200_Q2)
   {
      phi[0] = (1. - xi) * (1. - eta), phi[1] = xi * (1. - eta), phi[2] = (1. - xi) * eta, phi[3] = xi * eta;
      dphi[0][0] = -(1. - eta), dphi[0][1] = -(1. - xi);
      dphi[1][0] = (1. - eta), dphi[1][1] = -xi;
      dphi[2][0] = -eta, dphi[2][1] = (1. - xi);
      dphi[3][0] = eta, dphi[3][1] = xi;
   }
   else
   {
      phi[0] = 1. - xi - eta, phi[1] = xi, phi[2] = eta;
      dphi[0][0] = -1., dphi[0][1] = -1.;
      dphi[1][0] = 1., dphi[1][1] = 0.;
      dphi[2][0] = 0., dphi[2][1] = 1.;
   }
}

// physical gradients G (nd x 2) and vertex basis phi (nv) at the reference point (xi, eta);
// returns |det J|
template <int ET>
__device__ __forceinline__ double point_geometry(const double (*xv)[2], double xi, double eta, double (*G)[2],
                                                 double *phi)
{
   constexpr int nd = Elem<ET>::nd, nv = Elem<ET>::nv;
   double dN[nd][2], dphi[nv][2];
   ref_grads<ET>(xi, eta, dN);
   geom_basis<ET>(xi, eta, phi, dphi);
   double J00 = 0., J01 = 0., J10 = 0., J11 = 0.;  // J[i][m] = d x_i / d xi_m
#pragma unroll
   for (int v = 0; v < nv; ++v)
   {
      J00 += xv[v][0] * dphi[v][0];
      J01 += xv[v][0] * dphi[v][1];
      J10 += xv[v][1] * dphi[v][0];
      J11 += xv[v][1] * dphi[v][1];
   }
   const double det = J00 * J11 - J01 * J10;
   const double i00 = J11 / det, i01 = -J01 / det, i10 = -J10 / det, i11 = J00 / det;
#pragma unroll
   for (int a = 0; a < nd; ++a)
   {
      G[a][0] = dN[a][0] * i00 + dN[a][1] * i10;
      G[a][1] = dN[a][0] * i01 + dN[a][1] * i11;
   }
   return fabs(det);
}

// everything needed at quadrature point q: G (nd x 2), phi (nv), weight
template <int ET>
__device__ __forceinline__ double qp_geometry(const double (*xv)[2], int q, double (*G)[2], double *phi)
{
   double xi, eta, wq;
   quad_point<ET>(q, xi, eta, wq);
   return wq * point_geometry<ET>(xv, xi, eta, G, phi);
}

// Straight-sided triangles: J is constant, so the physical gradients at point q of the rule are fixed combinations of
// the barycentric gradients gl[v] = grad lambda_v (one reciprocal per cell instead of the per-point inverse of
// point_geometry): P1 G_v = gl[v]; P2 G_v = (4 L_v - 1) gl[v], G_(3+i) = 4 (L_j gl[k] + L_k gl[j]) for the edge (j, k)
// opposite vertex i, with L = (2/3 at the point's own vertex, 1/6 elsewhere).  Returns the vertex basis in phi.
template <int ET>
__device__ __forceinline__ void tri_point_grads(int q, const double (*gl)[2], double (*G)[2], double *phi)
{
   static_assert(ET == FEMB200_P1 || ET == FEMB200_P2, "triangles only");
   if (ET == FEMB200_P1)
   {
      phi[0] = phi[1] = phi[2] = 1. / 3.;
#pragma unroll
      for (int v = 0; v < 3; ++v) G[v][0] = gl[v][0], G[v][1] = gl[v][1];
      return;
   }
   const double L0 = q == 0 ? 2. / 3. : 1. / 6., L1 = q == 1 ? 2. / 3. : 1. / 6., L2 = q == 2 ? 2. / 3. : 1. / 6.;
   phi[0] = L0, phi[1] = L1, phi[2] = L2;
#pragma unroll
   for (int c = 0; c < 2; ++c)
   {
      G[0][c] = (4. * L0 - 1.) * gl[0][c];
      G[1][c] = (4. * L1 - 1.) * gl[1][c];
      G[2][c] = (4. * L2 - 1.) * gl[2][c];
      if (ET == FEMB200_P2)
      {
         G[Elem<ET>::nd > 3 ? 3 : 0][c] = 4. * (L1 * gl[2][c] + L2 * gl[1][c]);
         G[Elem<ET>::nd > 3 ? 4 : 0][c] = 4. * (L0 * gl[2][c] + L2 * gl[0][c]);
         G[Elem<ET>::nd > 3 ? 5 : 0][c] = 4. * (L0 * gl[1][c] + L1 * gl[0][c]);
      }
   }
}

// 2x2 block  w * B_a D B_b^t  with B rows (a,0) = [Gx, 0, Gy], (a,1) = [0, Gy, Gx]
// (M.cc:699-704, 885-887); D row-major 3x3.
__device__ __forceinline__ void bdb_block(const double *ga, const double *gb, const double *D, double w, double *k)
{
   const double c00 = ga[0] * D[0] + ga[1] * D[6], c01 = ga[0] * D[1] + ga[1] * D[7], c02 = ga[0] * D[2] + ga[1] * D[8];
   const double c10 = ga[1] * D[3] + ga[0] * D[6], c11 = ga[1] * D[4] + ga[0] * D[7], c12 = ga[1] * D[5] + ga[0] * D[8];
   k[0] += w * (c00 * gb[0] + c02 * gb[1]);
   k[1] += w * (c01 * gb[1] + c02 * gb[0]);
   k[2] += w * (c10 * gb[0] + c12 * gb[1]);
   k[3] += w * (c11 * gb[1] + c12 * gb[0]);
}

}  // namespace femb
