// mfem_adaptor_driver.cc -- compiles include/femb200_mfem.hpp against the MFEM stand-in of oracle/ref_shim
// (the real mfem.hpp is not in this image) and exposes C entry points with the argument lists of
// oracle/ref_shim/ref_driver.cc, so that tests/test_mfem_adaptor.py can run the adaptor and the reference's
// own damIntegrator side by side.  TEST INFRASTRUCTURE.
#include "mfem.hpp"  // oracle/ref_shim (stand-in)

#include "femb200_mfem.hpp"

extern "C" {

static mfem::IntegrationPoint centroid_point()
{
   mfem::IntegrationPoint ip;  // IntRules.Get(TRIANGLE, 1), M.cc:1112,1151-1152
   ip.x = 1. / 3., ip.y = 1. / 3., ip.weight = 0.5, ip.index = 0;
   return ip;
}

static void load_rule(mfem::IntegrationRule &ir)
{  // IntRules.Get(TRIANGLE, 2) in the order of ref_driver.cc
   const double p[3][2] = {{1. / 6., 1. / 6.}, {1. / 6., 2. / 3.}, {2. / 3., 1. / 6.}};
   for (int i = 0; i < 3; ++i) ir.IntPoint(i).x = p[i][0], ir.IntPoint(i).y = p[i][1], ir.IntPoint(i).weight = 1. / 6., ir.IntPoint(i).index = i;
}

// returns 0 on success; message of a failure through adaptor_last_error()
static thread_local char g_msg[512] = "";
const char *adaptor_last_error() { return g_msg; }

int adaptor_element_grad(const double *xv, double lam, double mu, double d, const double *elfun, int variant, double *elmat)
{
   try
   {
      mfem::ConstantCoefficient l(lam), m(mu);
      mfem::QuadratureFunctionCoefficient dam(d);
      mfem::VectorQuadratureFunctionCoefficient load;
      load.values.assign(6, 0.);
      mfem::IntegrationPoint ip = centroid_point();
      mfem::IntegrationRule ir(3);
      load_rule(ir);
      // ownership as M.cc:1485-1489: a raw `new`, released by whoever the form would be
      mfem::NonlinearFormIntegrator *pdi = new femb200::DamIntegrator(l, m, dam, ip, &ir, load, variant);
      femb200::DamIntegrator *integ = static_cast<femb200::DamIntegrator *>(pdi);
      mfem::FiniteElement el;
      mfem::ElementTransformation Tr(xv);
      mfem::Vector u(6);
      for (int i = 0; i < 6; ++i) u[i] = elfun ? elfun[i] : 0.;
      mfem::DenseMatrix K;
      integ->AssembleElementGrad(el, Tr, u, K);
      for (int i = 0; i < 36; ++i) elmat[i] = K.GetData()[i];
      delete pdi;
      return 0;
   }
   catch (const std::exception &e)
   {
      snprintf(g_msg, sizeof(g_msg), "%s", e.what());
      return 1;
   }
}

int adaptor_element_vector(const double *xv, double lam, double mu, double d, const double *elfun, const double *fq, double *elvect)
{
   try
   {
      mfem::ConstantCoefficient l(lam), m(mu);
      mfem::QuadratureFunctionCoefficient dam(d);
      mfem::VectorQuadratureFunctionCoefficient load;
      load.values.assign(6, 0.);
      if (fq)
         for (int i = 0; i < 6; ++i) load.values[i] = fq[i];
      mfem::IntegrationPoint ip = centroid_point();
      mfem::IntegrationRule ir(3);
      load_rule(ir);
      femb200::DamIntegrator integ(l, m, dam, ip, &ir, load);
      mfem::FiniteElement el;
      mfem::ElementTransformation Tr(xv);
      mfem::Vector u(6), r;
      for (int i = 0; i < 6; ++i) u[i] = elfun[i];
      integ.AssembleElementVector(el, Tr, u, r);
      for (int i = 0; i < 6; ++i) elvect[i] = r[i];
      return 0;
   }
   catch (const std::exception &e)
   {
      snprintf(g_msg, sizeof(g_msg), "%s", e.what());
      return 1;
   }
}

// GradientOperator: assemble on a small mesh and apply; y = K(u) x
int adaptor_gradient_mult(int etype, long nnodes, long ncells, const double *xy, const int *dofmap, const int *xdofmap,
                          const double *E, double nu, const unsigned char *ess, const double *u, const double *x, double *y,
                          long *nnz)
{
   try
   {
      femb200::GradientOperator G(etype, nnodes, ncells, xy, dofmap, xdofmap, E, nu, nullptr, ess);
      mfem::Vector U(const_cast<double *>(u), (int)(2 * nnodes)), X(const_cast<double *>(x), (int)(2 * nnodes)), Y((int)(2 * nnodes));
      G.Assemble(U);
      G.Mult(X, Y);
      for (long i = 0; i < 2 * nnodes; ++i) y[i] = Y[(int)i];
      *nnz = (long)G.NumNonZeros();
      return 0;
   }
   catch (const std::exception &e)
   {
      snprintf(g_msg, sizeof(g_msg), "%s", e.what());
      return 1;
   }
}
}
