// ref_driver.cc -- C entry points around the reference's own damIntegrator (appended
// to the reference line ranges by build_ref.sh, one translation unit).  TEST
// INFRASTRUCTURE ONLY.
//
//   xv      3 x 2 vertex coordinates, MFEM (counter-clockwise) orientation
//   elfun   6 dofs byNODES (ux0,ux1,ux2,uy0,uy1,uy2), M.cc:673
//   d       damage at the single quadrature point (M.cc:1319-1321)
//   elmat   6 x 6 column-major DenseMatrix; elvect 6 byNODES
//   fq      body force at the 3 points of the degree-2 rule, [3][2], or NULL
extern "C" {

static mfem::IntegrationPoint centroid_point()
{
   mfem::IntegrationPoint ip;  // IntRules.Get(TRIANGLE, 1): (1/3, 1/3), weight 1/2 (M.cc:1112,1151-1152)
   ip.x = 1. / 3., ip.y = 1. / 3., ip.weight = 0.5, ip.index = 0;
   return ip;
}

static void load_rule(mfem::IntegrationRule &ir)
{  // IntRules.Get(TRIANGLE, 2): 3 points, weights 1/6 (M.cc:1431-1454)
   const double p[3][2] = {{1. / 6., 1. / 6.}, {1. / 6., 2. / 3.}, {2. / 3., 1. / 6.}};
   for (int i = 0; i < 3; ++i)
   {
      ir.IntPoint(i).x = p[i][0], ir.IntPoint(i).y = p[i][1];
      ir.IntPoint(i).weight = 1. / 6., ir.IntPoint(i).index = i;
   }
}

void ref_load_points(double *pts /*[3][2]*/)
{
   mfem::IntegrationRule ir(3);
   load_rule(ir);
   for (int i = 0; i < 3; ++i) pts[2 * i] = ir.IntPoint(i).x, pts[2 * i + 1] = ir.IntPoint(i).y;
}

void ref_element_grad(const double *xv, double lam, double mu, double d, const double *elfun, double *elmat)
{
   mfem::ConstantCoefficient l(lam), m(mu);
   mfem::QuadratureFunctionCoefficient dam(d);
   mfem::VectorQuadratureFunctionCoefficient load;
   load.values.assign(6, 0.);
   mfem::IntegrationPoint ip = centroid_point();
   mfem::IntegrationRule ir(3);
   load_rule(ir);
   damIntegrator integ(l, m, dam, ip, &ir, load);
   mfem::FiniteElement el;
   mfem::ElementTransformation Tr(xv);
   mfem::Vector u(6);
   for (int i = 0; i < 6; ++i) u[i] = elfun ? elfun[i] : 0.;
   mfem::DenseMatrix K;
   integ.AssembleElementGrad(el, Tr, u, K);
   for (int i = 0; i < 36; ++i) elmat[i] = K.GetData()[i];
}

// n elements through ONE integrator object, as ParNonlinearForm::GetGradient drives it (M.cc:1546):
// xv [n][6], lam/mu/d [n], elmat [n][36] (d = 0 -> elfun is not read).  For timing the reference's
// element kernel on the host cores.
void ref_element_grad_batch(long n, const double *xv, const double *lam, const double *mu, const double *d,
                            double *elmat)
{
   mfem::ConstantCoefficient l(0.), m(0.);
   mfem::QuadratureFunctionCoefficient dam(0.);
   mfem::VectorQuadratureFunctionCoefficient load;
   load.values.assign(6, 0.);
   mfem::IntegrationPoint ip = centroid_point();
   mfem::IntegrationRule ir(3);
   load_rule(ir);
   damIntegrator integ(l, m, dam, ip, &ir, load);
   mfem::FiniteElement el;
   mfem::Vector u(6);
   mfem::DenseMatrix K;
   for (long e = 0; e < n; ++e)
   {
      l.constant = lam[e], m.constant = mu[e], dam.constant = d[e];
      mfem::ElementTransformation Tr(xv + 6 * e);
      integ.AssembleElementGrad(el, Tr, u, K);
      for (int i = 0; i < 36; ++i) elmat[36 * e + i] = K.GetData()[i];
   }
}

void ref_element_vector(const double *xv, double lam, double mu, double d, const double *elfun, const double *fq,
                        double *elvect)
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
   damIntegrator integ(l, m, dam, ip, &ir, load);
   mfem::FiniteElement el;
   mfem::ElementTransformation Tr(xv);
   mfem::Vector u(6), r;
   for (int i = 0; i < 6; ++i) u[i] = elfun[i];
   integ.AssembleElementVector(el, Tr, u, r);
   for (int i = 0; i < 6; ++i) elvect[i] = r[i];
}
}
