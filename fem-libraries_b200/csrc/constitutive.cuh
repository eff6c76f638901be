// constitutive.cuh -- tangent operator D (Voigt xx, yy, engineering xy) of the
// asymmetric traction/compression elasto-damage law, plane strain.
//
// Behaviour follows damIntegrator::AssembleElementGrad of the reference
// (MFEM/mechanic2d/asym_elasto_damage_model.cc, "M.cc"):
//   d <= 0            Hooke                                  M.cc:873-881
//   d > 0             d <- min(d, 1 - 1e-12)                 M.cc:739
//     closed form     E^t P E + q M                          M.cc:766-859
//     null strain     (1 - d) Hooke                          M.cc:861-870
//     AD variant      Hessian of psi by forward-mode duals   M.cc:100-155,752-765;
//                                                            autodiff/admfem.hpp:672-699
#pragma once
#include "common.cuh"

namespace femb {

constexpr double kLimit = 1.e-12;  // M.cc:513-514

__device__ __forceinline__ void hooke_scaled(double l, double m, double s, double *D)
{
   const double md = s * m, ld = s * l;
   D[0] = D[4] = 2 * md + ld;
   D[1] = D[3] = ld;
   D[8] = md;
   D[2] = D[5] = D[6] = D[7] = 0.;
}

// eps = [e00, e01, e10, e11], the symmetrised displacement gradient (M.cc:742-748)
__device__ inline void tangent_closed(double l, double m, double d, const double *eps, double *D)
{
   const double I1 = eps[0] + eps[3];
   const double I2 = eps[1] * eps[1] - eps[0] * eps[3];
   if (I1 > kLimit || I2 > kLimit || I1 < -kLimit || I2 < -kLimit)
   {
      const double delta = I1 * I1 + 4 * I2;
      const double r = sqrt(fmax(0., delta));
      const double e1 = (I1 + r) / 2., e2 = (I1 - r) / 2.;
      double cs, sn;
      if (r < kLimit)
      {  // singular second derivative: M.cc:784-789
         const double sg = (2 * eps[1] / (eps[0] - eps[3])) > 0. ? 1. : -1.;
         cs = sg * sqrt(2.) / 2.;
         sn = cs;
      }
      else
      {
         cs = (eps[0] - eps[3]) / r;
         sn = 2 * eps[1] / r;
      }
      const double a1 = (e1 >= 0) ? 1. : 0., a2 = (e2 >= 0) ? 1. : 0., a = (I1 >= 0) ? 1. : 0.;
      const double factor = 2. * m, gamma = 0.5 * l / m;
      const double c1 = 1. - a1 * d, c2 = 1. - a2 * d, c3 = 1. - a * d;
      const double P00 = factor * (c1 + gamma * c3), P01 = factor * gamma * c3, P11 = factor * (c2 + gamma * c3);
      const double de0[3] = {0.5 * (1 + cs), 0.5 * (1 - cs), 0.5 * sn};
      const double de1[3] = {0.5 * (1 - cs), 0.5 * (1 + cs), -0.5 * sn};
      const double cos2 = cs * cs, sin2 = sn * sn, sc = sn * cs;
      const double M[9] = {1. - cos2, -1. + cos2, -sc, -1. + cos2, 1. - cos2, sc, -sc, sc, 1. - sin2};
      const double q = (r >= kLimit) ? (I1 / r * (c1 - c2) + (c1 + c2)) : (c1 + c2);
#pragma unroll
      for (int i = 0; i < 3; ++i)
      {
         const double t0 = de0[i] * P00 + de1[i] * P01, t1 = de0[i] * P01 + de1[i] * P11;
#pragma unroll
         for (int j = 0; j < 3; ++j) D[3 * i + j] = (t0 * de0[j] + t1 * de1[j]) + q * (0.5 * m * M[3 * i + j]);
      }
   }
   else
      hooke_scaled(l, m, 1. - d, D);
}

// Second-order forward-mode jets in the four strain components: value, gradient g[4] and the lower triangle of the
// Hessian h[i (i + 1) / 2 + j], j <= i.  The reference obtains the 4 x 4 Hessian of psi from ten evaluations of the
// functor on nested duals internal::dual<dual<real_t>>, one per (i, j <= i) (admfem.hpp:619-631, 672-699); one pass of
// this jet carries all ten second derivatives through the same operations (product and square-root rules), so the
// result differs from the ten-pass form by rounding only, at about a quarter of the arithmetic (the per-cell pre-pass of
// the damaged reassembly is FP64-bound with the ten-pass form: 0.72 against 0.43 ms for the closed form at n = 1448).
struct Jet4
{
   double v, g[4], h[10];
};
__device__ __forceinline__ constexpr int jh(int i, int j) { return i * (i + 1) / 2 + j; }  // j <= i
__device__ __forceinline__ Jet4 jvar(double x, int k)
{
   Jet4 r;
   r.v = x;
#pragma unroll
   for (int i = 0; i < 4; ++i) r.g[i] = (i == k) ? 1. : 0.;
#pragma unroll
   for (int i = 0; i < 10; ++i) r.h[i] = 0.;
   return r;
}
__device__ __forceinline__ Jet4 operator+(const Jet4 &x, const Jet4 &y)
{
   Jet4 r;
   r.v = x.v + y.v;
#pragma unroll
   for (int i = 0; i < 4; ++i) r.g[i] = x.g[i] + y.g[i];
#pragma unroll
   for (int i = 0; i < 10; ++i) r.h[i] = x.h[i] + y.h[i];
   return r;
}
__device__ __forceinline__ Jet4 operator-(const Jet4 &x, const Jet4 &y)
{
   Jet4 r;
   r.v = x.v - y.v;
#pragma unroll
   for (int i = 0; i < 4; ++i) r.g[i] = x.g[i] - y.g[i];
#pragma unroll
   for (int i = 0; i < 10; ++i) r.h[i] = x.h[i] - y.h[i];
   return r;
}
__device__ __forceinline__ Jet4 operator*(const Jet4 &x, const Jet4 &y)
{
   Jet4 r;
   r.v = x.v * y.v;
#pragma unroll
   for (int i = 0; i < 4; ++i) r.g[i] = x.g[i] * y.v + x.v * y.g[i];
#pragma unroll
   for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j <= i; ++j)
         r.h[jh(i, j)] = x.h[jh(i, j)] * y.v + x.g[i] * y.g[j] + x.g[j] * y.g[i] + x.v * y.h[jh(i, j)];
   return r;
}
__device__ __forceinline__ Jet4 operator*(const Jet4 &x, double s)
{
   Jet4 r;
   r.v = x.v * s;
#pragma unroll
   for (int i = 0; i < 4; ++i) r.g[i] = x.g[i] * s;
#pragma unroll
   for (int i = 0; i < 10; ++i) r.h[i] = x.h[i] * s;
   return r;
}
__device__ __forceinline__ Jet4 jsqrt(const Jet4 &x)
{
   const double s = sqrt(x.v), f1 = 0.5 / s, f2 = -0.25 / (s * x.v);
   Jet4 r;
   r.v = s;
#pragma unroll
   for (int i = 0; i < 4; ++i) r.g[i] = f1 * x.g[i];
#pragma unroll
   for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j <= i; ++j) r.h[jh(i, j)] = f2 * x.g[i] * x.g[j] + f1 * x.h[jh(i, j)];
   return r;
}

// psi(strain; l, m, d), strain = (e11, e21, e12, e22): M.cc:100-155
__device__ inline Jet4 potential(double l, double m, double d, const Jet4 *s)
{
   const Jet4 I1 = s[0] + s[3];
   const Jet4 I2 = s[1] * s[2] - s[0] * s[3];
   if (I1.v > kLimit || I2.v > kLimit || I1.v < -kLimit || I2.v < -kLimit)
   {
      const Jet4 r = jsqrt(I1 * I1 + I2 * 4.);
      const Jet4 ev1 = (I1 + r) * 0.5, ev2 = (I1 - r) * 0.5;
      const double a1 = (ev1.v >= 0) ? 1. : 0., a2 = (ev2.v >= 0) ? 1. : 0.;
      const double a = ((ev1.v + ev2.v) >= 0) ? 1. : 0.;
      return (I1 * I1) * ((1. - a * d) * l / 2.) + ((ev1 * ev1) * (1 - a1 * d) + (ev2 * ev2) * (1. - a2 * d)) * m;
   }
   const Jet4 q = (s[0] * s[0] + s[3] * s[3]) + (s[1] * s[1] + s[2] * s[2]);
   return ((I1 * I1) * (l / 2.) + q * m) * (1 - d);
}

__device__ inline void tangent_ad(double l, double m, double d, const double *eps, double *D)
{
   // column-major DenseMatrix strain -> (eps11, eps21, eps12, eps22), M.cc:96-97,681
   const Jet4 s[4] = {jvar(eps[0], 0), jvar(eps[2], 1), jvar(eps[1], 2), jvar(eps[3], 3)};
   const Jet4 psi = potential(l, m, d, s);
   auto H = [&](int i, int j) { return i >= j ? psi.h[jh(i, j)] : psi.h[jh(j, i)]; };
#pragma unroll
   for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int j = 0; j < 3; ++j) D[3 * i + j] = H(i + 2 * (i % 2), j + 2 * (j % 2));  // M.cc:761-762
   D[8] = 0.5 * (D[8] + H(1, 2));                                                       // M.cc:763
}

// branch structure of M.cc:732-882
__device__ inline void tangent(int variant, double l, double m, double d, const double *eps, double *D)
{
   if (d > 0.)
   {
      d = fmin(d, 1. - kLimit);
      if (variant == FEMB200_TANGENT_AD)
         tangent_ad(l, m, d, eps, D);
      else
         tangent_closed(l, m, d, eps, D);
   }
   else
      hooke_scaled(l, m, 1., D);
}

// Stress (already multiplied by the weight w): asym_stress of the reference, non-AD path
// (M.cc:207-329), including its identity-eigenvector branch (when |eps_xy| <= 1e-12
// the larger eigenvalue is paired with the x axis whatever the strain).  eps = [e00,e01,e10,e11]
// symmetrised; sig = [s00,s01,s10,s11].
__device__ inline void asym_stress(double l, double m, double d, double w, const double *eps, double *sig)
{
   if (d > 0.)
   {
      const double I1 = eps[0] + eps[3];
      const double I2 = eps[1] * eps[1] - eps[0] * eps[3];
      sig[0] = sig[1] = sig[2] = sig[3] = 0.;
      if (I1 > kLimit || I2 > kLimit || I1 < -kLimit || I2 < -kLimit)
      {
         const double delta = I1 * I1 + 4 * I2;
         const double r = sqrt(fmax(0., delta));
         const double e0 = (I1 + r) / 2., e1 = (I1 - r) / 2.;
         const double a1 = (e0 >= 0) ? 1. : 0., a2 = (e1 >= 0) ? 1. : 0., a = ((e0 + e1) >= 0) ? 1. : 0.;
         if (!((d == 1.) && (a == 1) && (a1 == 1) && (a2 == 1)))
         {
            double v00 = 1., v01 = 0., v10 = 0., v11 = 1.;
            if (fabs(eps[2]) > kLimit)
            {
               v00 = e0 - eps[3], v01 = e1 - eps[3], v10 = v11 = eps[2];
               const double n0 = sqrt(v00 * v00 + v10 * v10), n1 = sqrt(v01 * v01 + v11 * v11);
               v00 /= n0, v10 /= n0, v01 /= n1, v11 /= n1;
            }
            const double temp = 2. * m * w, gamma = 0.5 * l / m;
            const double c = 1 - a * d, c1 = 1 - a1 * d, c2 = 1 - a2 * d;
            const double D0 = temp * (c1 + gamma * c), D1 = temp * gamma * c, D2 = temp * (c2 + gamma * c);
            const double s0 = D0 * e0 + D1 * e1, s1 = D1 * e0 + D2 * e1;
            sig[0] = v00 * s0 * v00 + v01 * s1 * v01;
            sig[1] = v00 * s0 * v10 + v01 * s1 * v11;
            sig[2] = sig[1];
            sig[3] = v10 * s0 * v10 + v11 * s1 * v11;
         }
      }
   }
   else
   {
      const double m2plw = w * (2 * m + l), lw = l * w;
      sig[0] = m2plw * eps[0] + lw * eps[3];
      sig[3] = m2plw * eps[3] + lw * eps[0];
      sig[1] = sig[2] = w * m * (eps[1] + eps[2]);
   }
}

// Lame coefficients, the MFEM form (M.cc:1087-1098): lambda = E*c2, mu = E*c3
struct LameCoef
{
   double c2, c3;
};
inline LameCoef lame_coef(double nu)
{
   const double c1 = 1. + nu;
   return LameCoef{nu / (c1 * (1. - 2. * nu)), 1. / (2 * c1)};
}

}  // namespace femb
