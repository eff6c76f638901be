/*
 * fem_oracle.c -- CPU ORACLE.  TEST INFRASTRUCTURE ONLY.
 *
 * A plain-C restatement of the arithmetic of the mechanic2d hot path of
 * SalzmanA/fem-libraries (element tangent integration -> CSR scatter-add ->
 * operator apply inside CG).  Nothing under fem-libraries_b200/ may import,
 * link or call this file: it is the checker used by tests/, by
 * __graft_entry__.smoke() and by bench.py's cpu_baseline / --impl reference
 * legs, never the product.
 *
 * PARITY STATUS.  The reference ships no golden vectors, no tests and no fixtures
 * for this path, and MFEM / dolfinx / ffcx are absent from this
 * image, so the reference binaries cannot be run.  What pins this file:
 *   - PINNED (P1 element kernel, the only element the reference has): oracle/_ref
 *     = the reference's own damIntegrator + asym_stress (M.cc:207-329, 490-953)
 *     compiled in place from /root/reference against a minimal MFEM stand-in
 *     (oracle/ref_shim/); tests/golden/ref_p1_vectors.json holds its outputs on
 *     every triangle of common/data/square.msh (d = 0, both reference code paths)
 *     and on 40 damaged cases (tangent, residual, load; special branches);
 *     tests/test_reference_goldens.py checks this file against them (<= 1e-14
 *     linear, <= 1e-12 damaged);
 *   - the known answers derived from the reference formulas on the reference's
 *     only mesh (common/data/square.msh): tests/golden/square_kat.json;
 *   - three independent code paths for the P1 tangent that must agree, and the
 *     closed-form damaged tangent against the dual-number Hessian of the potential.
 *   - UNPINNED (no reference code or vectors exist): P2 / Q2 elements, the CSR
 *     ordering convention (dolfinx: un-vendored), the Jacobi-PCG iteration (the
 *     reference preconditions with HYPRE BoomerAMG).  These follow SURVEY.md 8c /
 *     A.9 and are checked by properties only (patch test, rigid-body null space,
 *     symmetry, assembled == matrix-free, direct solve).
 *
 * All file:line citations are relative to /root/reference/, with
 *   M.cc  = MFEM/mechanic2d/asym_elasto_damage_model.cc
 *   F.cc  = FEniCSx/mechanic2d/asym_elasto_damage_model.cc
 *   manual.py = FEniCSx/mechanic2d/asym_manual.py
 *
 * P2 triangles and Q2 quads do not exist in the reference:
 * they are the same formulas with the element conventions fixed in
 * SURVEY.md 8c / Appendix A.9.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_P1 0
#define ORC_P2 1
#define ORC_Q2 2

#define ORC_LAYOUT_ROWMAJOR_INTERLEAVED 0 /* ufcx: A[(2a+i)*ndof + 2b+k]         */
#define ORC_LAYOUT_COLMAJOR_BYNODES 1     /* MFEM: elmat(i*nd+a, k*nd+b), col-major */

#define ORC_TANGENT_CLOSED 0 /* M.cc:766-871 */
#define ORC_TANGENT_AD 1     /* M.cc:752-765 */

static const double ORC_LIMIT = 1.e-12; /* M.cc:513-514 */

int orc_num_threads(void)
{
#ifdef _OPENMP
   return omp_get_max_threads();
#else
   return 1;
#endif
}

/* ------------------------------------------------------------------------ */
/* Materials                                                                 */
/* ------------------------------------------------------------------------ */

/* The 200 Young moduli of the reference: glibc srand(6575)/rand()%200
 * (M.cc:1076-1085, F.cc:533-541).  */
void orc_E_table(double *E200)
{
   srand(6575);
   for (int i = 0; i < 200; ++i)
   {
      const double a = (1.e8 - 5.e6) / 199.;
      E200[i] = a * (rand() % 200) + 5.e6;
   }
}

/* Lame coefficients, MFEM form (M.cc:1087-1098): lambda = E*c2, mu = E*c3. */
void orc_lame(double E, double nu, double *lam, double *mu)
{
   const double c1 = 1. + nu;
   const double c2 = nu / (c1 * (1. - 2. * nu));
   const double c3 = 1. / (2 * c1);
   *lam = E * c2;
   *mu = E * c3;
}

/* ------------------------------------------------------------------------ */
/* Constitutive tangent D (Voigt xx, yy, engineering xy), row-major 3x3      */
/* ------------------------------------------------------------------------ */

/* M.cc:873-881 (d <= 0) and M.cc:861-870 (null strain with damage: scale). */
static void hooke_scaled(double l, double m, double s, double D[9])
{
   const double md = s * m, ld = s * l;
   for (int i = 0; i < 9; ++i) D[i] = 0.;
   D[0] = D[4] = 2 * md + ld;
   D[1] = D[3] = ld;
   D[8] = md;
}

/* Closed-form damaged tangent, M.cc:766-859.  eps = [e00, e01, e10, e11]
 * (symmetrised displacement gradient, M.cc:742-748).  d already clamped. */
static void tangent_closed(double l, double m, double d, const double eps[4], double D[9])
{
   const double I1 = eps[0] + eps[3];
   const double I2 = eps[1] * eps[1] - eps[0] * eps[3];
   if (I1 > ORC_LIMIT || I2 > ORC_LIMIT || I1 < -ORC_LIMIT || I2 < -ORC_LIMIT)
   {
      const double delta = I1 * I1 + 4 * I2;
      const double r = sqrt(delta > 0. ? delta : 0.);
      const double e1 = (I1 + r) / 2.;
      const double e2 = (I1 - r) / 2.;
      double coss, sinn;
      if (r < ORC_LIMIT)
      {
         const double signe = (2 * eps[1] / (eps[0] - eps[3])) > 0. ? 1 : -1;
         coss = signe * sqrt(2.) / 2.;
         sinn = coss;
      }
      else
      {
         coss = (eps[0] - eps[3]) / r;
         sinn = 2 * eps[1] / r;
      }
      const double alpha1 = (e1 >= 0) ? 1. : 0.;
      const double alpha2 = (e2 >= 0) ? 1. : 0.;
      const double alpha = (I1 >= 0) ? 1. : 0.;
      const double factor = 2. * m;
      const double gamma = 0.5 * l / m;
      const double c1 = 1. - alpha1 * d;
      const double c2 = 1. - alpha2 * d;
      const double c3 = 1. - alpha * d;
      double P[2][2];
      P[0][0] = factor * (c1 + gamma * c3);
      P[0][1] = factor * gamma * c3;
      P[1][1] = factor * (c2 + gamma * c3);
      P[1][0] = P[0][1];
      double de[2][3];
      de[0][0] = 0.5 * (1 + coss);
      de[0][1] = 0.5 * (1 - coss);
      de[0][2] = 0.5 * sinn;
      de[1][0] = 0.5 * (1 - coss);
      de[1][1] = 0.5 * (1 + coss);
      de[1][2] = 0.5 * -sinn;
      const double cos2 = coss * coss, sin2 = sinn * sinn, sc = sinn * coss;
      double M[3][3];
      M[0][0] = 1. - cos2;
      M[0][1] = -1. + cos2;
      M[1][1] = 1. - cos2;
      M[0][2] = -1. * sc;
      M[1][2] = 1. * sc;
      M[2][2] = (1 - sin2);
      M[1][0] = M[0][1];
      M[2][0] = M[0][2];
      M[2][1] = M[1][2];
      const double q = (r >= ORC_LIMIT) ? (I1 / r * (c1 - c2) + (c1 + c2)) : (c1 + c2);
      /* hook = dedeps^t . P . dedeps + q * (0.5 m) M   (M.cc:855-858) */
      double T[3][2];
      for (int i = 0; i < 3; ++i)
         for (int j = 0; j < 2; ++j) T[i][j] = de[0][i] * P[0][j] + de[1][i] * P[1][j];
      for (int i = 0; i < 3; ++i)
         for (int j = 0; j < 3; ++j)
            D[3 * i + j] = (T[i][0] * de[0][j] + T[i][1] * de[1][j]) + q * (0.5 * m * M[i][j]);
   }
   else
      hooke_scaled(l, m, 1. - d, D); /* M.cc:861-870 */
}

/* ---- second-order forward jets: the nested dual<dual<double>> of
 * admfem.hpp:619-631; v = value.value, a = value.gradient,
 * b = gradient.value, ab = gradient.gradient. ---- */
typedef struct
{
   double v, a, b, ab;
} jet2;
static jet2 j_c(double c)
{
   jet2 r = {c, 0., 0., 0.};
   return r;
}
static jet2 j_add(jet2 x, jet2 y)
{
   jet2 r = {x.v + y.v, x.a + y.a, x.b + y.b, x.ab + y.ab};
   return r;
}
static jet2 j_sub(jet2 x, jet2 y)
{
   jet2 r = {x.v - y.v, x.a - y.a, x.b - y.b, x.ab - y.ab};
   return r;
}
static jet2 j_mul(jet2 x, jet2 y)
{
   jet2 r;
   r.v = x.v * y.v;
   r.a = x.a * y.v + x.v * y.a;
   r.b = x.b * y.v + x.v * y.b;
   r.ab = x.ab * y.v + x.b * y.a + x.a * y.b + x.v * y.ab;
   return r;
}
static jet2 j_scale(jet2 x, double s)
{
   jet2 r = {x.v * s, x.a * s, x.b * s, x.ab * s};
   return r;
}
static jet2 j_sqrt(jet2 x)
{
   const double s = sqrt(x.v);
   const double f1 = 0.5 / s;             /* f'  */
   const double f2 = -0.25 / (s * x.v);   /* f'' */
   jet2 r = {s, f1 * x.a, f1 * x.b, f2 * x.a * x.b + f1 * x.ab};
   return r;
}

/* psi(strain; l, m, d), strain = (e11, e21, e12, e22): M.cc:100-155.  The
 * alpha switches are plain reals (not differentiated), M.cc:128-141. */
static jet2 potential_jet(double l, double m, double d, const jet2 s[4])
{
   const jet2 I1 = j_add(s[0], s[3]);
   const jet2 I2 = j_sub(j_mul(s[1], s[2]), j_mul(s[0], s[3]));
   if (I1.v > ORC_LIMIT || I2.v > ORC_LIMIT || I1.v < -ORC_LIMIT || I2.v < -ORC_LIMIT)
   {
      const jet2 delta = j_add(j_mul(I1, I1), j_scale(I2, 4.));
      const jet2 r = j_sqrt(delta);
      const jet2 ev1 = j_scale(j_add(I1, r), 0.5);
      const jet2 ev2 = j_scale(j_sub(I1, r), 0.5);
      const double alpha1 = (ev1.v >= 0) ? 1. : 0.;
      const double alpha2 = (ev2.v >= 0) ? 1. : 0.;
      const double alpha = ((ev1.v + ev2.v) >= 0) ? 1. : 0.;
      const jet2 t1 = j_scale(j_mul(I1, I1), (1. - alpha * d) * l / 2.);
      const jet2 t2 = j_add(j_scale(j_mul(ev1, ev1), (1 - alpha1 * d)), j_scale(j_mul(ev2, ev2), (1. - alpha2 * d)));
      return j_add(t1, j_scale(t2, m));
   }
   else
   {
      jet2 q = j_add(j_add(j_mul(s[0], s[0]), j_mul(s[3], s[3])), j_add(j_mul(s[1], s[1]), j_mul(s[2], s[2])));
      return j_scale(j_add(j_scale(j_mul(I1, I1), l / 2.), j_scale(q, m)), (1 - d));
   }
}

/* AD tangent: 4x4 Hessian by 10 jet evaluations (admfem.hpp:672-699), then
 * the Voigt re-indexing of M.cc:761-763. */
static void tangent_ad(double l, double m, double d, const double eps[4], double D[9])
{
   double H[4][4];
   /* MFEM DenseMatrix 'strain' is column-major: data = (e00, e10, e01, e11)
    * = (eps11, eps21, eps12, eps22) (M.cc:96-97, vstrain at M.cc:681). */
   const double u[4] = {eps[0], eps[2], eps[1], eps[3]};
   for (int ii = 0; ii < 4; ++ii)
      for (int jj = 0; jj <= ii; ++jj)
      {
         jet2 s[4];
         for (int k = 0; k < 4; ++k) s[k] = j_c(u[k]);
         s[ii].a = 1.0;
         s[jj].b = 1.0;
         const jet2 rez = potential_jet(l, m, d, s);
         H[ii][jj] = H[jj][ii] = rez.ab;
      }
   for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) D[3 * i + j] = H[i + 2 * (i % 2)][j + 2 * (j % 2)];
   D[8] = 0.5 * (D[8] + H[1][2]);
}

/* Tangent at one quadrature point: branch structure of M.cc:732-882. */
void orc_tangent(int variant, double l, double m, double d, const double eps[4], double D[9])
{
   if (d > 0.)
   {
      d = fmin(d, 1. - ORC_LIMIT); /* M.cc:739 */
      if (variant == ORC_TANGENT_AD)
         tangent_ad(l, m, d, eps, D);
      else
         tangent_closed(l, m, d, eps, D);
   }
   else
      hooke_scaled(l, m, 1., D);
}

/* Stress (already multiplied by the weight w), non-AD path M.cc:207-329.
 * eps = [e00,e01,e10,e11]; sig = [s00,s01,s10,s11].  Reproduces the
 * identity-eigenvector branch quirk as the reference does. */
void orc_stress(double l, double m, double d, double w, const double eps[4], double sig[4])
{
   if (d > 0.)
   {
      const double I1 = eps[0] + eps[3];
      const double I2 = eps[1] * eps[1] - eps[0] * eps[3];
      if (I1 > ORC_LIMIT || I2 > ORC_LIMIT || I1 < -ORC_LIMIT || I2 < -ORC_LIMIT)
      {
         const double delta = I1 * I1 + 4 * I2;
         const double r = sqrt(delta > 0. ? delta : 0.);
         const double ev[2] = {(I1 + r) / 2., (I1 - r) / 2.};
         const double alpha1 = (ev[0] >= 0) ? 1. : 0.;
         const double alpha2 = (ev[1] >= 0) ? 1. : 0.;
         const double alpha = ((ev[0] + ev[1]) >= 0) ? 1. : 0.;
         if (!((d == 1.) && (alpha == 1) && (alpha1 == 1) && (alpha2 == 1)))
         {
            double V[2][2];
            if (fabs(eps[2]) > ORC_LIMIT)
            {
               V[0][0] = ev[0] - eps[3];
               V[0][1] = ev[1] - eps[3];
               V[1][0] = V[1][1] = eps[2];
               const double n0 = sqrt(V[0][0] * V[0][0] + V[1][0] * V[1][0]);
               const double n1 = sqrt(V[0][1] * V[0][1] + V[1][1] * V[1][1]);
               V[0][0] /= n0;
               V[1][0] /= n0;
               V[0][1] /= n1;
               V[1][1] /= n1;
            }
            else
            {
               V[0][0] = V[1][1] = 1.;
               V[1][0] = V[0][1] = 0.;
            }
            const double temp = 2. * m * w;
            const double gamma = 0.5 * l / m;
            const double c = 1 - alpha * d, c1 = 1 - alpha1 * d, c2 = 1 - alpha2 * d;
            const double D0 = temp * (c1 + gamma * c), D1 = temp * gamma * c, D2 = temp * (c2 + gamma * c);
            const double s0 = D0 * ev[0] + D1 * ev[1];
            const double s1 = D1 * ev[0] + D2 * ev[1];
            /* MultADAt: sig = V diag(s) V^t */
            sig[0] = V[0][0] * s0 * V[0][0] + V[0][1] * s1 * V[0][1];
            sig[1] = V[0][0] * s0 * V[1][0] + V[0][1] * s1 * V[1][1];
            sig[2] = sig[1];
            sig[3] = V[1][0] * s0 * V[1][0] + V[1][1] * s1 * V[1][1];
         }
         else
            sig[0] = sig[1] = sig[2] = sig[3] = 0.;
      }
      else
         sig[0] = sig[1] = sig[2] = sig[3] = 0.;
   }
   else
   {
      const double m2plw = w * (2 * m + l), lw = l * w;
      sig[0] = m2plw * eps[0] + lw * eps[3];
      sig[3] = m2plw * eps[3] + lw * eps[0];
      sig[1] = sig[2] = w * m * (eps[1] + eps[2]);
   }
}

/* ------------------------------------------------------------------------ */
/* Reference elements and quadrature (SURVEY.md 8c conventions)              */
/* ------------------------------------------------------------------------ */

static int elem_nd(int etype) { return etype == ORC_P1 ? 3 : (etype == ORC_P2 ? 6 : 9); }
static int elem_nv(int etype) { return etype == ORC_Q2 ? 4 : 3; }
static int elem_nq(int etype) { return etype == ORC_P1 ? 1 : (etype == ORC_P2 ? 3 : 9); }
int orc_elem_nd(int etype) { return elem_nd(etype); }
int orc_elem_nv(int etype) { return elem_nv(etype); }
int orc_elem_nq(int etype) { return elem_nq(etype); }

static void quad_rule(int etype, double pts[][2], double *wts)
{
   if (etype == ORC_P1)
   { /* 1 point, centroid, weight 1/2 (manual.py:97, M.cc:1151-1152) */
      pts[0][0] = pts[0][1] = 1. / 3.;
      wts[0] = 0.5;
   }
   else if (etype == ORC_P2)
   { /* degree-2, 3 points, weights 1/6 */
      const double a = 1. / 6., b = 2. / 3.;
      pts[0][0] = a, pts[0][1] = a;
      pts[1][0] = b, pts[1][1] = a;
      pts[2][0] = a, pts[2][1] = b;
      wts[0] = wts[1] = wts[2] = 1. / 6.;
   }
   else
   { /* 3x3 Gauss-Legendre on [0,1]^2 */
      const double s = sqrt(3. / 5.);
      const double t[3] = {0.5 * (1. - s), 0.5, 0.5 * (1. + s)};
      const double w[3] = {5. / 18., 8. / 18., 5. / 18.};
      for (int j = 0; j < 3; ++j)
         for (int i = 0; i < 3; ++i)
         {
            pts[3 * j + i][0] = t[i];
            pts[3 * j + i][1] = t[j];
            wts[3 * j + i] = w[i] * w[j];
         }
   }
}

/* Scalar basis values N[a] and reference gradients dN[a][2] at (xi, eta). */
static void basis(int etype, double xi, double eta, double *N, double dN[][2])
{
   if (etype == ORC_P1)
   { /* gradients [[-1,-1],[1,0],[0,1]] = CalcDShape of M.cc:692 */
      N[0] = 1. - xi - eta, N[1] = xi, N[2] = eta;
      dN[0][0] = -1., dN[0][1] = -1.;
      dN[1][0] = 1., dN[1][1] = 0.;
      dN[2][0] = 0., dN[2][1] = 1.;
   }
   else if (etype == ORC_P2)
   { /* basix order: 3 vertices, then edge i opposite vertex i */
      const double L[3] = {1. - xi - eta, xi, eta};
      const double dL[3][2] = {{-1., -1.}, {1., 0.}, {0., 1.}};
      for (int i = 0; i < 3; ++i)
      {
         N[i] = L[i] * (2. * L[i] - 1.);
         dN[i][0] = (4. * L[i] - 1.) * dL[i][0];
         dN[i][1] = (4. * L[i] - 1.) * dL[i][1];
      }
      const int ep[3][2] = {{1, 2}, {0, 2}, {0, 1}};
      for (int e = 0; e < 3; ++e)
      {
         const int p = ep[e][0], q = ep[e][1];
         N[3 + e] = 4. * L[p] * L[q];
         dN[3 + e][0] = 4. * (L[q] * dL[p][0] + L[p] * dL[q][0]);
         dN[3 + e][1] = 4. * (L[q] * dL[p][1] + L[p] * dL[q][1]);
      }
   }
   else
   { /* Q2: tensor Gauss-Lobatto nodes 0, 1/2, 1; local index ix + 3*iy */
      double lx[3], ly[3], dx[3], dy[3];
      lx[0] = 2. * (xi - 0.5) * (xi - 1.), lx[1] = 4. * xi * (1. - xi), lx[2] = 2. * xi * (xi - 0.5);
      ly[0] = 2. * (eta - 0.5) * (eta - 1.), ly[1] = 4. * eta * (1. - eta), ly[2] = 2. * eta * (eta - 0.5);
      dx[0] = 4. * xi - 3., dx[1] = 4. - 8. * xi, dx[2] = 4. * xi - 1.;
      dy[0] = 4. * eta - 3., dy[1] = 4. - 8. * eta, dy[2] = 4. * eta - 1.;
      for (int j = 0; j < 3; ++j)
         for (int i = 0; i < 3; ++i)
         {
            N[3 * j + i] = lx[i] * ly[j];
            dN[3 * j + i][0] = dx[i] * ly[j];
            dN[3 * j + i][1] = lx[i] * dy[j];
         }
   }
}

/* Geometry basis on the nv vertices (affine triangle / bilinear quad with
 * vertices in tensor order (0,0),(1,0),(0,1),(1,1)). */
static void geom_basis(int etype, double xi, double eta, double *phi, double dphi[][2])
{
   if (etype == ORC_Q2)
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

/* Per-quadrature-point geometry: physical gradients G = dN . J^-1 (M.cc:696)
 * and weight w = w_q * |det J| (M.cc:684-685; |.| per SURVEY.md B6). */
static double qp_geometry(int etype, const double *xv, double xi, double eta, double wq, double *N, double G[][2],
                          double *phi)
{
   const int nd = elem_nd(etype), nv = elem_nv(etype);
   double dN[9][2], dphi[4][2];
   basis(etype, xi, eta, N, dN);
   geom_basis(etype, xi, eta, phi, dphi);
   double J[2][2] = {{0., 0.}, {0., 0.}}; /* J[i][m] = d x_i / d xi_m */
   for (int v = 0; v < nv; ++v)
      for (int i = 0; i < 2; ++i)
         for (int mm = 0; mm < 2; ++mm) J[i][mm] += xv[2 * v + i] * dphi[v][mm];
   const double det = J[0][0] * J[1][1] - J[0][1] * J[1][0];
   const double Ji[2][2] = {{J[1][1] / det, -J[0][1] / det}, {-J[1][0] / det, J[0][0] / det}};
   for (int a = 0; a < nd; ++a)
   {
      G[a][0] = dN[a][0] * Ji[0][0] + dN[a][1] * Ji[1][0];
      G[a][1] = dN[a][0] * Ji[0][1] + dN[a][1] * Ji[1][1];
   }
   return wq * fabs(det);
}

/* ------------------------------------------------------------------------ */
/* Element tangent matrix, generic quadrature loop                           */
/*   K_e = sum_q w_q |det J_q| B_q D_q B_q^t   (M.cc:639-916; SURVEY A.1,A.9) */
/* xv: nv x 2 vertex coordinates.  dnod: nodal damage at the nv vertices or  */
/* NULL (d = 0).  u: nd x 2 interleaved element displacement or NULL.        */
/* ------------------------------------------------------------------------ */
void orc_element_grad(int etype, const double *xv, double lam, double mu, const double *dnod, const double *u,
                      int variant, int layout, double *A)
{
   const int nd = elem_nd(etype), nv = elem_nv(etype), nq = elem_nq(etype), n = 2 * nd;
   double pts[9][2], wts[9];
   quad_rule(etype, pts, wts);
   double K[18 * 18];
   for (int i = 0; i < n * n; ++i) K[i] = 0.;
   for (int q = 0; q < nq; ++q)
   {
      double N[9], G[9][2], phi[4];
      const double w = qp_geometry(etype, xv, pts[q][0], pts[q][1], wts[q], N, G, phi);
      double d = 0.;
      if (dnod)
         for (int v = 0; v < nv; ++v) d += phi[v] * dnod[v];
      double D[9];
      if (d > 0.)
      {
         double gu[2][2] = {{0., 0.}, {0., 0.}}; /* gu[i][j] = d u_i / d x_j  (M.cc:742) */
         if (u)
            for (int a = 0; a < nd; ++a)
               for (int i = 0; i < 2; ++i)
                  for (int j = 0; j < 2; ++j) gu[i][j] += u[2 * a + i] * G[a][j];
         const double eps[4] = {gu[0][0], 0.5 * (gu[0][1] + gu[1][0]), 0.5 * (gu[0][1] + gu[1][0]), gu[1][1]};
         orc_tangent(variant, lam, mu, d, eps, D);
      }
      else
         orc_tangent(variant, lam, mu, 0., NULL, D);
      /* B row for (a, comp 0) = [Gx, 0, Gy]; (a, comp 1) = [0, Gy, Gx]  (M.cc:699-704) */
      for (int a = 0; a < nd; ++a)
      {
         const double Ba[2][3] = {{G[a][0], 0., G[a][1]}, {0., G[a][1], G[a][0]}};
         double C[2][3];
         for (int i = 0; i < 2; ++i)
            for (int c = 0; c < 3; ++c) C[i][c] = Ba[i][0] * D[c] + Ba[i][1] * D[3 + c] + Ba[i][2] * D[6 + c];
         for (int b = 0; b < nd; ++b)
         {
            const double Bb[2][3] = {{G[b][0], 0., G[b][1]}, {0., G[b][1], G[b][0]}};
            for (int i = 0; i < 2; ++i)
               for (int k = 0; k < 2; ++k)
                  K[(2 * a + i) * n + 2 * b + k] += w * (C[i][0] * Bb[k][0] + C[i][1] * Bb[k][1] + C[i][2] * Bb[k][2]);
         }
      }
   }
   if (layout == ORC_LAYOUT_ROWMAJOR_INTERLEAVED)
      for (int i = 0; i < n * n; ++i) A[i] += K[i]; /* ufcx: caller pre-zeroes, kernel accumulates */
   else
      for (int a = 0; a < nd; ++a)
         for (int i = 0; i < 2; ++i)
            for (int b = 0; b < nd; ++b)
               for (int k = 0; k < 2; ++k) A[(i * nd + a) + (k * nd + b) * n] = K[(2 * a + i) * n + 2 * b + k];
}

/* ---- P1, the reference's own two formulations, in MFEM layout
 * (6x6 column-major, dofs byNODES ux0,ux1,ux2,uy0,uy1,uy2). ---- */

static double p1_geometry(const double *xv, double G[3][2])
{ /* M.cc:684-696 */
   const double J[2][2] = {{xv[2] - xv[0], xv[4] - xv[0]}, {xv[3] - xv[1], xv[5] - xv[1]}};
   const double det = J[0][0] * J[1][1] - J[0][1] * J[1][0];
   const double Ji[2][2] = {{J[1][1] / det, -J[0][1] / det}, {-J[1][0] / det, J[0][0] / det}};
   const double dsh[3][2] = {{-1., -1.}, {1., 0.}, {0., 1.}};
   for (int a = 0; a < 3; ++a)
   {
      G[a][0] = dsh[a][0] * Ji[0][0] + dsh[a][1] * Ji[1][0];
      G[a][1] = dsh[a][0] * Ji[0][1] + dsh[a][1] * Ji[1][1];
   }
   return 0.5 * fabs(det);
}

static void p1_tangent_at_centroid(const double G[3][2], double l, double m, double d, const double *elfun,
                                   int variant, double D[9])
{ /* M.cc:732-882; elfun byNODES as a 3x2 column-major matrix (M.cc:673) */
   if (d > 0. && elfun)
   {
      double gu[2][2] = {{0., 0.}, {0., 0.}};
      for (int a = 0; a < 3; ++a)
         for (int i = 0; i < 2; ++i)
            for (int j = 0; j < 2; ++j) gu[i][j] += elfun[i * 3 + a] * G[a][j];
      const double s = 0.5 * (gu[0][1] + gu[1][0]);
      const double eps[4] = {gu[0][0], s, s, gu[1][1]};
      orc_tangent(variant, l, m, d, eps, D);
   }
   else if (d > 0.)
   {
      const double eps[4] = {0., 0., 0., 0.};
      orc_tangent(variant, l, m, d, eps, D);
   }
   else
      orc_tangent(variant, l, m, 0., NULL, D);
}

/* USE_B path: elmat = w * (B.hook) . B^t   (M.cc:699-704, 885-887). */
void orc_p1_grad_mfem_B(const double *xv, double l, double m, double d, const double *elfun, int variant, double *elmat)
{
   double G[3][2], D[9];
   const double w = p1_geometry(xv, G);
   p1_tangent_at_centroid(G, l, m, d, elfun, variant, D);
   double B[6][3], C[6][3];
   memset(B, 0, sizeof(B));
   for (int a = 0; a < 3; ++a)
   {
      B[a][0] = G[a][0];     /* B.SetSubMatrix(0, 0, gdshapex)  */
      B[3 + a][1] = G[a][1]; /* B.SetSubMatrix(nd, 1, gdshapey) */
      B[a][2] = G[a][1];     /* B.SetSubMatrix(0, 2, gdshapey)  */
      B[3 + a][2] = G[a][0]; /* B.SetSubMatrix(nd, 2, gdshapex) */
   }
   for (int r = 0; r < 6; ++r)
      for (int c = 0; c < 3; ++c) C[r][c] = B[r][0] * D[c] + B[r][1] * D[3 + c] + B[r][2] * D[6 + c];
   for (int c = 0; c < 6; ++c)
      for (int r = 0; r < 6; ++r) elmat[r + 6 * c] = w * (C[r][0] * B[c][0] + C[r][1] * B[c][1] + C[r][2] * B[c][2]);
}

/* Tensor-product block path (no USE_B): M.cc:705-717, 893-911. */
void orc_p1_grad_mfem_blocks(const double *xv, double l, double m, double d, const double *elfun, int variant,
                             double *elmat)
{
   double G[3][2], D[9];
   const double w = p1_geometry(xv, G);
   p1_tangent_at_centroid(G, l, m, d, elfun, variant, D);
   double gxgx[3][3], gygy[3][3], gxgy[3][3], gygx[3][3], gxy[3][3];
   for (int a = 0; a < 3; ++a)
      for (int b = 0; b < 3; ++b)
      {
         gxgx[a][b] = w * G[a][0] * G[b][0];
         gygy[a][b] = w * G[a][1] * G[b][1];
         gxgy[a][b] = w * G[a][0] * G[b][1];
      }
   for (int a = 0; a < 3; ++a)
      for (int b = 0; b < 3; ++b)
      {
         gygx[a][b] = gxgy[b][a];
         gxy[a][b] = gxgy[a][b] + gxgy[b][a];
      }
   const int b20 = fabs(D[6]) > ORC_LIMIT, b21 = fabs(D[7]) > ORC_LIMIT;
   for (int i = 0; i < 36; ++i) elmat[i] = 0.;
   for (int a = 0; a < 3; ++a)
      for (int b = 0; b < 3; ++b)
      {
         double xx = D[0] * gxgx[a][b] + D[8] * gygy[a][b];
         if (b20) xx += D[6] * gxy[a][b];
         double yy = D[4] * gygy[a][b] + D[8] * gxgx[a][b];
         if (b21) yy += D[7] * gxy[a][b];
         double xy = D[3] * gxgy[a][b] + D[8] * gygx[a][b];
         if (b20) xy += D[6] * gxgx[a][b];
         if (b21) xy += D[5] * gygy[a][b];
         elmat[a + 6 * b] += xx;
         elmat[(3 + a) + 6 * (3 + b)] += yy;
         elmat[a + 6 * (3 + b)] += xy;
         elmat[(3 + b) + 6 * a] += xy; /* block.Transpose() */
      }
}

/* ufcx-signature shim for the J form of manual.py:102 (P1, degree-1 rule).
 * w packs the coefficients in creation order d[3], E[1], u[6] (manual.py:19,
 * 22,30); c = [nu] (manual.py:23); coordinate_dofs is 3 x (x,y,z).  A is
 * 6x6 row-major, dofs interleaved, accumulated into (caller pre-zeroes). */
void orc_tabulate_tensor_J_p1(double *A, const double *w, const double *c, const double *coordinate_dofs,
                              const int *entity_local_index, const uint8_t *quadrature_permutation)
{
   (void)entity_local_index;
   (void)quadrature_permutation;
   const double xv[6] = {coordinate_dofs[0], coordinate_dofs[1], coordinate_dofs[3],
                         coordinate_dofs[4], coordinate_dofs[6], coordinate_dofs[7]};
   double lam, mu;
   /* manual.py:25-26 */
   mu = w[3] / (2.0 * (1.0 + c[0]));
   lam = w[3] * c[0] / ((1.0 + c[0]) * (1.0 - 2.0 * c[0]));
   orc_element_grad(ORC_P1, xv, lam, mu, w, w + 4, ORC_TANGENT_CLOSED, ORC_LAYOUT_ROWMAJOR_INTERLEAVED, A);
}

/* ------------------------------------------------------------------------ */
/* Element residual  r_e = G.sigma_w - sum_q w_q N f   (M.cc:559-637)         */
/* interleaved output (ux0,uy0,...), P1 only as in the reference.            */
/* ------------------------------------------------------------------------ */
void orc_p1_element_vector(const double *xv, double l, double m, double d, const double *u /*3x2 interleaved*/,
                           const double *fnod /*3x2 interleaved nodal load or NULL*/, double *r)
{
   double G[3][2];
   const double w = p1_geometry(xv, G);
   double gu[2][2] = {{0., 0.}, {0., 0.}};
   for (int a = 0; a < 3; ++a)
      for (int i = 0; i < 2; ++i)
         for (int j = 0; j < 2; ++j) gu[i][j] += u[2 * a + i] * G[a][j];
   const double s = 0.5 * (gu[0][1] + gu[1][0]);
   const double eps[4] = {gu[0][0], s, s, gu[1][1]};
   double sig[4];
   orc_stress(l, m, d, w, eps, sig);
   for (int a = 0; a < 3; ++a)
   { /* AddMult(gdshape, sig, res): res(a,i) = sum_j G(a,j) sig(j,i)  (M.cc:601) */
      r[2 * a + 0] = G[a][0] * sig[0] + G[a][1] * sig[2];
      r[2 * a + 1] = G[a][0] * sig[1] + G[a][1] * sig[3];
   }
   if (fnod)
   { /* 3-point degree-2 rule, load interpolated from nodes (M.cc:613-632) */
      double pts[9][2], wts[9];
      quad_rule(ORC_P2, pts, wts);
      for (int q = 0; q < 3; ++q)
      {
         const double N[3] = {1. - pts[q][0] - pts[q][1], pts[q][0], pts[q][1]};
         const double wl = wts[q] * (2. * w); /* ipl.weight * Tr.Weight() */
         double f[2] = {0., 0.};
         for (int a = 0; a < 3; ++a)
         {
            f[0] += N[a] * fnod[2 * a];
            f[1] += N[a] * fnod[2 * a + 1];
         }
         for (int a = 0; a < 3; ++a)
         {
            r[2 * a + 0] -= wl * N[a] * f[0];
            r[2 * a + 1] -= wl * N[a] * f[1];
         }
      }
   }
}

/* Generic element residual (any family): stress term with the element's own rule
 * (P1: the 1-point rule the reference forces, M.cc:1112,1151-1152), load term with the
 * degree-2 3-point rule on triangles (M.cc:613-632) / 3x3 Gauss on quads, the nodal
 * load interpolated with the element's own basis.  P2 / Q2 are extrapolations of the
 * reference (P1 only); for P1 this equals orc_p1_element_vector.
 * u: nd x 2 interleaved; dnod: nv nodal damage or NULL; fnod: nd x 2 or NULL. */
void orc_element_vector(int etype, const double *xv, double lam, double mu, const double *dnod, const double *u,
                        const double *fnod, double *r)
{
   const int nd = elem_nd(etype), nv = elem_nv(etype), nq = elem_nq(etype);
   double pts[9][2], wts[9];
   quad_rule(etype, pts, wts);
   for (int i = 0; i < 2 * nd; ++i) r[i] = 0.;
   for (int q = 0; q < nq; ++q)
   {
      double N[9], G[9][2], phi[4];
      const double w = qp_geometry(etype, xv, pts[q][0], pts[q][1], wts[q], N, G, phi);
      double d = 0.;
      if (dnod)
         for (int v = 0; v < nv; ++v) d += phi[v] * dnod[v];
      double gu[2][2] = {{0., 0.}, {0., 0.}};
      for (int a = 0; a < nd; ++a)
         for (int i = 0; i < 2; ++i)
            for (int j = 0; j < 2; ++j) gu[i][j] += u[2 * a + i] * G[a][j];
      const double s = 0.5 * (gu[0][1] + gu[1][0]);
      const double eps[4] = {gu[0][0], s, s, gu[1][1]};
      double sig[4];
      orc_stress(lam, mu, d, w, eps, sig);
      for (int a = 0; a < nd; ++a)
      {
         r[2 * a + 0] += G[a][0] * sig[0] + G[a][1] * sig[2];
         r[2 * a + 1] += G[a][0] * sig[1] + G[a][1] * sig[3];
      }
   }
   if (fnod)
   {
      const int lrule = (etype == ORC_Q2) ? ORC_Q2 : ORC_P2;
      const int lq = elem_nq(lrule);
      quad_rule(lrule, pts, wts);
      for (int q = 0; q < lq; ++q)
      {
         double N[9], G[9][2], phi[4];
         const double w = qp_geometry(etype, xv, pts[q][0], pts[q][1], wts[q], N, G, phi);
         double f[2] = {0., 0.};
         for (int a = 0; a < nd; ++a)
         {
            f[0] += N[a] * fnod[2 * a];
            f[1] += N[a] * fnod[2 * a + 1];
         }
         for (int a = 0; a < nd; ++a)
         {
            r[2 * a + 0] -= w * N[a] * f[0];
            r[2 * a + 1] -= w * N[a] * f[1];
         }
      }
   }
}

/* Global residual vector b = sum_e P_e^t r_e (role of assemble_vector(F), F.cc:825, and of
 * ParNonlinearForm::Mult -> AssembleElementVector, M.cc:559-637), no boundary treatment.
 * fnod: nnodes x 2 nodal load or NULL. */
void orc_assemble_vector(int etype, int64_t ncells, int64_t nnodes, const double *x, const int32_t *xdofmap,
                         const int32_t *dofmap, const double *E, double nu, const double *dnod, const double *u,
                         const double *fnod, double *b)
{
   const int nd = elem_nd(etype), nv = elem_nv(etype);
   for (int64_t i = 0; i < 2 * nnodes; ++i) b[i] = 0.;
   for (int64_t e = 0; e < ncells; ++e)
   {
      double xv[8], dv[4], ue[18], fe[18], r[18];
      for (int v = 0; v < nv; ++v)
      {
         const int32_t g = xdofmap[e * nv + v];
         xv[2 * v] = x[2 * (int64_t)g];
         xv[2 * v + 1] = x[2 * (int64_t)g + 1];
         if (dnod) dv[v] = dnod[g];
      }
      for (int a = 0; a < nd; ++a)
      {
         const int64_t g = dofmap[e * nd + a];
         ue[2 * a] = u[2 * g], ue[2 * a + 1] = u[2 * g + 1];
         if (fnod) fe[2 * a] = fnod[2 * g], fe[2 * a + 1] = fnod[2 * g + 1];
      }
      double lam, mu;
      orc_lame(E[e], nu, &lam, &mu);
      orc_element_vector(etype, xv, lam, mu, dnod ? dv : NULL, ue, fnod ? fe : NULL, r);
      for (int a = 0; a < nd; ++a)
      {
         const int64_t g = dofmap[e * nd + a];
         b[2 * g] += r[2 * a];
         b[2 * g + 1] += r[2 * a + 1];
      }
   }
}

/* ------------------------------------------------------------------------ */
/* Field operators either side of the hot path      */
/* ------------------------------------------------------------------------ */

/* Damage-field smoothing, serial semantics of M.cc:1258-1315 (same algorithm with a SciPy
 * adjacency product at F.py:160-199).  adjptr/adj: the vertex graph (edge neighbours of every
 * vertex, ascending, no self).  niter double sweeps, each reading the previous field in full
 * (the reference fills the scratch vector vv before updating, M.cc:1270-1291):
 *   sweep A: s_l = [d_l < thr] sum_n d_n ; d_l = max(s_l * (1/deg_l), d_l)   (thr = 0.01, M.cc:1274)
 *   sweep B: s_l = sum_n d_n             ; d_l = max(s_l * (1/deg_l), d_l)
 * The reference multiplies by the precomputed reciprocal of the edge count (M.cc:1247,1292).
 * Its summation order is MFEM's vertex-to-edge table order (not available here); the oracle sums
 * in ascending neighbour order ("parity unpinned" for the last bits). */
void orc_smooth_damage(int64_t nverts, const int64_t *adjptr, const int32_t *adj, double *d, int niter, double thr)
{
   double *t = (double *)malloc(sizeof(double) * (size_t)(nverts > 0 ? nverts : 1));
   for (int it = 0; it < niter; ++it)
      for (int sweep = 0; sweep < 2; ++sweep)
      {
         const double *in = sweep ? t : d;
         double *out = sweep ? d : t;
         for (int64_t l = 0; l < nverts; ++l)
         {
            const int64_t deg = adjptr[l + 1] - adjptr[l];
            double s = 0.;
            if (sweep == 1 || in[l] < thr)
               for (int64_t k = adjptr[l]; k < adjptr[l + 1]; ++k) s += in[adj[k]];
            const double inv = deg > 0 ? 1. / (double)deg : 0.;
            const double v = s * inv;
            out[l] = v > in[l] ? v : in[l];
         }
      }
   free(t);
}

/* DG0 strain and stress output fields (strainTensor / stressTensor, M.cc:333-430 projected on a
 * DG0 space at M.cc:1551-1563; F.cc:909-942): per cell the symmetrised gradient of u
 * (M.cc:343-349) and asym_stress with weight one (M.cc:415-427) at the DG0 node = cell centroid,
 * stored (xx, xy, yy).  dnod: nodal damage (indexed like x) or NULL.  strain/stress may be NULL. */
void orc_cell_strain_stress(int etype, int64_t ncells, const double *x, const int32_t *xdofmap, const int32_t *dofmap,
                            const double *E, double nu, const double *dnod, const double *u, double *strain,
                            double *stress)
{
   const int nd = elem_nd(etype), nv = elem_nv(etype);
   const double xi = etype == ORC_Q2 ? 0.5 : 1. / 3., eta = xi;
   for (int64_t e = 0; e < ncells; ++e)
   {
      double xv[8], N[9], G[9][2], phi[4];
      for (int v = 0; v < nv; ++v)
      {
         const int64_t g = xdofmap[e * nv + v];
         xv[2 * v] = x[2 * g], xv[2 * v + 1] = x[2 * g + 1];
      }
      qp_geometry(etype, xv, xi, eta, 1., N, G, phi);
      double d = 0.;
      if (dnod)
         for (int v = 0; v < nv; ++v) d += phi[v] * dnod[xdofmap[e * nv + v]];
      double gu[2][2] = {{0., 0.}, {0., 0.}};
      for (int a = 0; a < nd; ++a)
      {
         const int64_t g = dofmap[e * nd + a];
         for (int i = 0; i < 2; ++i)
            for (int j = 0; j < 2; ++j) gu[i][j] += u[2 * g + i] * G[a][j];
      }
      const double sh = 0.5 * (gu[0][1] + gu[1][0]);
      if (strain) strain[3 * e] = gu[0][0], strain[3 * e + 1] = sh, strain[3 * e + 2] = gu[1][1];
      if (stress)
      {
         const double eps[4] = {gu[0][0], sh, sh, gu[1][1]};
         double lam, mu, sig[4];
         orc_lame(E[e], nu, &lam, &mu);
         orc_stress(lam, mu, d, 1., eps, sig);
         stress[3 * e] = sig[0], stress[3 * e + 1] = sig[1], stress[3 * e + 2] = sig[3];
      }
   }
}

/* ------------------------------------------------------------------------ */
/* Sparsity pattern (role of dolfinx create_matrix, F.cc:688)                */
/* rows in dof order, columns ascending and unique, structural, bs=2 dofs    */
/* (2*node + comp).  Call with colidx == NULL to get rowptr and nnz.          */
/* ------------------------------------------------------------------------ */
static int cmp_i32(const void *a, const void *b)
{
   const int32_t x = *(const int32_t *)a, y = *(const int32_t *)b;
   return (x > y) - (x < y);
}

int64_t orc_build_pattern(int64_t nnodes, int64_t ncells, int nd, const int32_t *dofmap, int64_t *rowptr,
                          int32_t *colidx)
{
   /* node -> cells */
   int64_t *nptr = (int64_t *)calloc((size_t)nnodes + 1, sizeof(int64_t));
   for (int64_t e = 0; e < ncells; ++e)
      for (int a = 0; a < nd; ++a) nptr[dofmap[e * nd + a] + 1]++;
   for (int64_t i = 0; i < nnodes; ++i) nptr[i + 1] += nptr[i];
   int64_t *cur = (int64_t *)malloc((size_t)nnodes * sizeof(int64_t));
   memcpy(cur, nptr, (size_t)nnodes * sizeof(int64_t));
   int32_t *ncell = (int32_t *)malloc((size_t)nptr[nnodes] * sizeof(int32_t));
   for (int64_t e = 0; e < ncells; ++e)
      for (int a = 0; a < nd; ++a) ncell[cur[dofmap[e * nd + a]]++] = (int32_t)e;
   free(cur);
   int64_t maxc = 0;
   for (int64_t i = 0; i < nnodes; ++i)
      if (nptr[i + 1] - nptr[i] > maxc) maxc = nptr[i + 1] - nptr[i];
   int32_t *tmp = (int32_t *)malloc((size_t)(maxc * nd + 1) * sizeof(int32_t));
   int64_t nnzb = 0;
   rowptr[0] = 0;
   for (int64_t I = 0; I < nnodes; ++I)
   {
      int cnt = 0;
      for (int64_t k = nptr[I]; k < nptr[I + 1]; ++k)
         for (int b = 0; b < nd; ++b) tmp[cnt++] = dofmap[(int64_t)ncell[k] * nd + b];
      qsort(tmp, (size_t)cnt, sizeof(int32_t), cmp_i32);
      int deg = 0;
      for (int k = 0; k < cnt; ++k)
         if (k == 0 || tmp[k] != tmp[k - 1]) tmp[deg++] = tmp[k];
      const int64_t base = 4 * nnzb;
      if (colidx)
         for (int r = 0; r < 2; ++r)
            for (int s = 0; s < deg; ++s)
            {
               colidx[base + (int64_t)r * 2 * deg + 2 * s] = 2 * tmp[s];
               colidx[base + (int64_t)r * 2 * deg + 2 * s + 1] = 2 * tmp[s] + 1;
            }
      rowptr[2 * I + 1] = base + 2 * deg;
      rowptr[2 * I + 2] = base + 4 * deg;
      nnzb += deg;
   }
   free(tmp);
   free(ncell);
   free(nptr);
   return 4 * nnzb;
}

/* ------------------------------------------------------------------------ */
/* Assembly (role of the setJ lambda, F.cc:847-862)                          */
/*   zero -> per cell: tabulate, zero Dirichlet rows/cols of A_e, add into   */
/*   the CSR by column search -> flush -> diag on Dirichlet dofs -> final.   */
/* x: nnodes x 2; xdofmap: ncells x nv; dofmap: ncells x nd; E: ncells;      */
/* dnod: damage per node (P1 field on the geometry vertices) or NULL;        */
/* u: 2*nnodes dof vector or NULL; bc: per-dof marker or NULL.               */
/* nthreads > 1: row-block ownership per thread (domain decomposition).       */
/* ------------------------------------------------------------------------ */
static inline int64_t find_col(const int32_t *colidx, int64_t lo, int64_t hi, int32_t c)
{
   while (lo < hi)
   {
      const int64_t mid = (lo + hi) >> 1;
      if (colidx[mid] < c)
         lo = mid + 1;
      else
         hi = mid;
   }
   return lo;
}

void orc_assemble_matrix(int etype, int64_t ncells, int64_t nnodes, const double *x, const int32_t *xdofmap,
                         const int32_t *dofmap, const double *E, double nu, const double *dnod, const double *u,
                         int variant, const uint8_t *bc, double diag, const int64_t *rowptr, const int32_t *colidx,
                         double *values, int nthreads)
{
   const int nd = elem_nd(etype), nv = elem_nv(etype), n = 2 * nd;
   const int64_t nnz = rowptr[2 * nnodes];
   memset(values, 0, (size_t)nnz * sizeof(double)); /* MatZeroEntries, F.cc:850 */
   if (nthreads < 1) nthreads = 1;
   /* nthreads > 1: the reference parallelises by domain decomposition (one MPI rank
    * per mesh partition, doc.tex:393-464): thread t owns a contiguous block of
    * node rows and integrates every cell touching them (interface cells are
    * integrated by both neighbours, as ghost cells would be); no atomics. */
#pragma omp parallel num_threads(nthreads)
   {
#ifdef _OPENMP
      const int tid = omp_get_thread_num(), nt = omp_get_num_threads();
#else
      const int tid = 0, nt = 1;
#endif
      const int64_t node_lo = nnodes * tid / nt, node_hi = nnodes * (tid + 1) / nt;
      for (int64_t e = 0; e < ncells; ++e)
      {
         int mine = 0;
         for (int a = 0; a < nd; ++a)
         {
            const int64_t g = dofmap[e * nd + a];
            if (g >= node_lo && g < node_hi) mine = 1;
         }
         if (!mine) continue;
         double xv[8], dv[4], ue[18], A[18 * 18];
         for (int v = 0; v < nv; ++v)
         {
            const int32_t g = xdofmap[e * nv + v];
            xv[2 * v] = x[2 * (int64_t)g];
            xv[2 * v + 1] = x[2 * (int64_t)g + 1];
            if (dnod) dv[v] = dnod[g];
         }
         if (u)
            for (int a = 0; a < nd; ++a)
            {
               const int64_t g = dofmap[e * nd + a];
               ue[2 * a] = u[2 * g];
               ue[2 * a + 1] = u[2 * g + 1];
            }
         double lam, mu;
         orc_lame(E[e], nu, &lam, &mu);
         for (int i = 0; i < n * n; ++i) A[i] = 0.;
         orc_element_grad(etype, xv, lam, mu, dnod ? dv : NULL, u ? ue : NULL, variant,
                          ORC_LAYOUT_ROWMAJOR_INTERLEAVED, A);
         if (bc) /* dolfinx assemble_matrix zeroes BC rows and columns of A_e, F.cc:852 */
            for (int a = 0; a < nd; ++a)
               for (int i = 0; i < 2; ++i)
                  if (bc[2 * (int64_t)dofmap[e * nd + a] + i])
                  {
                     const int r = 2 * a + i;
                     for (int c = 0; c < n; ++c) A[r * n + c] = A[c * n + r] = 0.;
                  }
         for (int a = 0; a < nd; ++a)
         {
            const int64_t ga = dofmap[e * nd + a];
            if (ga < node_lo || ga >= node_hi) continue;
            for (int i = 0; i < 2; ++i)
            {
               const int64_t row = 2 * ga + i;
               const int64_t lo = rowptr[row], hi = rowptr[row + 1];
               for (int b = 0; b < nd; ++b)
               {
                  const int64_t p = find_col(colidx, lo, hi, 2 * dofmap[e * nd + b]);
                  values[p] += A[(2 * a + i) * n + 2 * b];
                  values[p + 1] += A[(2 * a + i) * n + 2 * b + 1];
               }
            }
         }
      }
   }
   if (bc) /* set_diagonal(..., 1.) with INSERT_VALUES, F.cc:857 */
      for (int64_t row = 0; row < 2 * nnodes; ++row)
         if (bc[row]) values[find_col(colidx, rowptr[row], rowptr[row + 1], (int32_t)row)] = diag;
}

/* Dense element matrices for a batch of cells (parity target of the batched
 * tabulate kernel).  out: ncells x (2nd)^2. */
void orc_tabulate_batch(int etype, int64_t ncells, const double *x, const int32_t *xdofmap, const int32_t *dofmap,
                        const double *E, double nu, const double *dnod, const double *u, int variant, int layout,
                        double *out)
{
   const int nd = elem_nd(etype), nv = elem_nv(etype), n = 2 * nd;
#pragma omp parallel for schedule(static)
   for (int64_t e = 0; e < ncells; ++e)
   {
      double xv[8], dv[4], ue[18];
      for (int v = 0; v < nv; ++v)
      {
         const int32_t g = xdofmap[e * nv + v];
         xv[2 * v] = x[2 * (int64_t)g];
         xv[2 * v + 1] = x[2 * (int64_t)g + 1];
         if (dnod) dv[v] = dnod[g];
      }
      if (u)
         for (int a = 0; a < nd; ++a)
         {
            const int64_t g = dofmap[e * nd + a];
            ue[2 * a] = u[2 * g];
            ue[2 * a + 1] = u[2 * g + 1];
         }
      double lam, mu;
      orc_lame(E[e], nu, &lam, &mu);
      double *A = out + e * (int64_t)(n * n);
      for (int i = 0; i < n * n; ++i) A[i] = 0.;
      orc_element_grad(etype, xv, lam, mu, dnod ? dv : NULL, u ? ue : NULL, variant, layout, A);
   }
}

/* ------------------------------------------------------------------------ */
/* Operator apply                                                            */
/* ------------------------------------------------------------------------ */
void orc_spmv(int64_t nrows, const int64_t *rowptr, const int32_t *colidx, const double *values, const double *x,
              double *y, int nthreads)
{
   if (nthreads < 1) nthreads = 1;
#pragma omp parallel for schedule(static) num_threads(nthreads)
   for (int64_t r = 0; r < nrows; ++r)
   {
      double s = 0.;
      for (int64_t p = rowptr[r]; p < rowptr[r + 1]; ++p) s += values[p] * x[colidx[p]];
      y[r] = s;
   }
}

/* Matrix-free apply of the same operator:
 *   y = sum_e P_e^t ( sum_q w_q B_q (D_q (B_q^t x_e)) ),
 * Dirichlet treated as in the assembled operator: rows/cols zeroed, diag on
 * the diagonal. Serial (used as a checker only). */
void orc_apply_matrix_free(int etype, int64_t ncells, int64_t nnodes, const double *x, const int32_t *xdofmap,
                           const int32_t *dofmap, const double *E, double nu, const uint8_t *bc, double diag,
                           const double *xin, double *y)
{
   const int nd = elem_nd(etype), nv = elem_nv(etype), nq = elem_nq(etype);
   double pts[9][2], wts[9];
   quad_rule(etype, pts, wts);
   for (int64_t i = 0; i < 2 * nnodes; ++i) y[i] = 0.;
   for (int64_t e = 0; e < ncells; ++e)
   {
      double xv[8], xe[18], ye[18];
      for (int v = 0; v < nv; ++v)
      {
         const int32_t g = xdofmap[e * nv + v];
         xv[2 * v] = x[2 * (int64_t)g];
         xv[2 * v + 1] = x[2 * (int64_t)g + 1];
      }
      for (int a = 0; a < nd; ++a)
         for (int i = 0; i < 2; ++i)
         {
            const int64_t g = 2 * (int64_t)dofmap[e * nd + a] + i;
            xe[2 * a + i] = (bc && bc[g]) ? 0. : xin[g];
            ye[2 * a + i] = 0.;
         }
      double lam, mu;
      orc_lame(E[e], nu, &lam, &mu);
      double D[9];
      orc_tangent(0, lam, mu, 0., NULL, D);
      for (int q = 0; q < nq; ++q)
      {
         double N[9], G[9][2], phi[4];
         const double w = qp_geometry(etype, xv, pts[q][0], pts[q][1], wts[q], N, G, phi);
         double st[3] = {0., 0., 0.}; /* B^t x_e: (exx, eyy, gxy) */
         for (int a = 0; a < nd; ++a)
         {
            st[0] += G[a][0] * xe[2 * a];
            st[1] += G[a][1] * xe[2 * a + 1];
            st[2] += G[a][1] * xe[2 * a] + G[a][0] * xe[2 * a + 1];
         }
         double sg[3];
         for (int c = 0; c < 3; ++c) sg[c] = w * (D[3 * c] * st[0] + D[3 * c + 1] * st[1] + D[3 * c + 2] * st[2]);
         for (int a = 0; a < nd; ++a)
         {
            ye[2 * a] += G[a][0] * sg[0] + G[a][1] * sg[2];
            ye[2 * a + 1] += G[a][1] * sg[1] + G[a][0] * sg[2];
         }
      }
      for (int a = 0; a < nd; ++a)
         for (int i = 0; i < 2; ++i)
         {
            const int64_t g = 2 * (int64_t)dofmap[e * nd + a] + i;
            if (!(bc && bc[g])) y[g] += ye[2 * a + i];
         }
   }
   if (bc)
      for (int64_t g = 0; g < 2 * nnodes; ++g)
         if (bc[g]) y[g] = diag * xin[g];
}

/* ------------------------------------------------------------------------ */
/* (Jacobi-)preconditioned CG with mfem::CGSolver semantics (M.cc:1502,      */
/* 1525-1528): zero initial guess, stop when <B r, r> <= max(rtol^2 <B r0,   */
/* r0>, atol^2).  Returns 1 if converged.                                    */
/* ------------------------------------------------------------------------ */
static double dot(int64_t n, const double *a, const double *b, int nthreads)
{
   double s = 0.;
#pragma omp parallel for reduction(+ : s) schedule(static) num_threads(nthreads)
   for (int64_t i = 0; i < n; ++i) s += a[i] * b[i];
   return s;
}

int orc_pcg(int64_t n, const int64_t *rowptr, const int32_t *colidx, const double *values, const double *b, double *x,
            double rtol, double atol, int maxit, int jacobi, int *iters, double *final_norm, int nthreads)
{
   if (nthreads < 1) nthreads = 1;
   double *r = (double *)malloc((size_t)n * sizeof(double));
   double *z = (double *)malloc((size_t)n * sizeof(double));
   double *d = (double *)malloc((size_t)n * sizeof(double));
   double *dinv = NULL;
   if (jacobi)
   {
      dinv = (double *)malloc((size_t)n * sizeof(double));
      for (int64_t i = 0; i < n; ++i)
         dinv[i] = 1. / values[find_col(colidx, rowptr[i], rowptr[i + 1], (int32_t)i)];
   }
   int converged = 0;
   for (int64_t i = 0; i < n; ++i)
   {
      x[i] = 0.;
      r[i] = b[i];
      d[i] = jacobi ? dinv[i] * r[i] : r[i];
   }
   double nom = dot(n, d, r, nthreads);
   const double r0 = fmax(nom * rtol * rtol, atol * atol);
   int it = 0;
   if (nom <= r0)
      converged = 1;
   else
   {
      orc_spmv(n, rowptr, colidx, values, d, z, nthreads);
      double den = dot(n, z, d, nthreads);
      it = maxit;
      for (int i = 1; den > 0.;)
      {
         const double alpha = nom / den;
#pragma omp parallel for schedule(static) num_threads(nthreads)
         for (int64_t k = 0; k < n; ++k)
         {
            x[k] += alpha * d[k];
            r[k] -= alpha * z[k];
         }
         double betanom;
         if (jacobi)
         {
#pragma omp parallel for schedule(static) num_threads(nthreads)
            for (int64_t k = 0; k < n; ++k) z[k] = dinv[k] * r[k];
            betanom = dot(n, r, z, nthreads);
         }
         else
            betanom = dot(n, r, r, nthreads);
         if (betanom <= r0)
         {
            converged = 1;
            it = i;
            nom = betanom;
            break;
         }
         if (++i > maxit)
         {
            nom = betanom;
            break;
         }
         const double beta = betanom / nom;
         const double *zz = jacobi ? z : r;
#pragma omp parallel for schedule(static) num_threads(nthreads)
         for (int64_t k = 0; k < n; ++k) d[k] = zz[k] + beta * d[k];
         orc_spmv(n, rowptr, colidx, values, d, z, nthreads);
         den = dot(n, d, z, nthreads);
         nom = betanom;
      }
   }
   *iters = it;
   *final_norm = sqrt(nom > 0. ? nom : 0.);
   free(r);
   free(z);
   free(d);
   free(dinv);
   return converged;
}
