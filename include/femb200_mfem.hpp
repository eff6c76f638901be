// femb200_mfem.hpp -- header-only MFEM adaptor over the C ABI of femb200.h.
//
// Include AFTER <mfem.hpp> (MFEM 4.7, README.md:7 of the reference) in a translation unit that links
// libfemb200.so and the CUDA runtime.  It gives the reference's MFEM driver
// (MFEM/mechanic2d/asym_elasto_damage_model.cc = M.cc) two drop-ins:
//
//   femb200::DamIntegrator   a mfem::NonlinearFormIntegrator with the constructor and the virtuals of the
//                            reference's damIntegrator (M.cc:490-546, 559, 639): AssembleElementVector /
//                            AssembleElementGrad evaluate ONE element on the device (batch of one through
//                            femb200_assemble_vector / femb200_element_grad_batched), so that
//                            `F.AddDomainIntegrator(new femb200::DamIntegrator(...))` (M.cc:1485-1489) works
//                            unchanged: the form takes ownership of the raw pointer, the destructor releases
//                            the device buffers.  Layouts are MFEM's: elfun / elvect byNODES (M.cc:673), elmat
//                            column-major DenseMatrix.  This is the element-level parity surface; it pays two
//                            PCIe round trips per element and is not the fast path.
//   femb200::GradientOperator  the fast path: the whole of ParNonlinearForm::GetGradient + HypreParMatrix::Mult
//                            (M.cc:1546, 1502-1528) on the device: Assemble(u) = femb200_assemble_matrix on a
//                            plan built once from the mesh arrays, Mult(x, y) = femb200_spmv.  vdofs byVDIM
//                            (M.cc:1107) are 2 * node + component, the library's global numbering.
//
// Only the dozen MFEM calls the reference's own integrator uses are needed (Vector, DenseMatrix,
// ElementTransformation::InverseJacobian / Weight / SetIntPoint, Coefficient::Eval,
// VectorQuadratureFunctionCoefficient::Eval, IntegrationRule), so the header also compiles against the MFEM
// stand-in of oracle/ref_shim/mfem.hpp: tests/test_mfem_adaptor.py builds it that way and compares it with
// the reference's own damIntegrator compiled from M.cc.
#ifndef FEMB200_MFEM_HPP
#define FEMB200_MFEM_HPP

#include <cuda_runtime_api.h>

#include <stdexcept>
#include <string>

#include "femb200.h"

namespace femb200
{

inline void check(int rc, const char *what)
{
   if (rc != 0) throw std::runtime_error(std::string(what) + ": " + femb200_last_error());
}
inline void cuda_check(cudaError_t e, const char *what)
{
   if (e != cudaSuccess) throw std::runtime_error(std::string(what) + ": " + cudaGetErrorString(e));
}

class DamIntegrator : public mfem::NonlinearFormIntegrator
{
   mfem::Coefficient &lambda, &mu;
   mfem::QuadratureFunctionCoefficient &dam;
   mfem::IntegrationPoint ip;  // the single stiffness point (M.cc:1112,1151-1152)
   mfem::VectorQuadratureFunctionCoefficient *load;
   int variant;
   // device: one P1 cell {0, 1, 2}
   femb200_plan *plan = nullptr;
   int32_t *d_map = nullptr;
   double *d_buf = nullptr;  // xv[6] | E[1] | dnod[3] | u[6] | f[6] | out[36]
   enum
   {
      O_X = 0,
      O_E = 6,
      O_D = 7,
      O_U = 10,
      O_F = 16,
      O_OUT = 22,
      N_BUF = 58
   };

   // vertex coordinates up to a translation (the element matrices do not see it) from the transformation:
   // J = (InverseJacobian)^-1, vertices (0,0), J e_1, J e_2
   static void vertices(mfem::ElementTransformation &Tr, double *xv)
   {
      const mfem::DenseMatrix &Ji = Tr.InverseJacobian();
      const double a = Ji(0, 0), b = Ji(0, 1), c = Ji(1, 0), d = Ji(1, 1), det = a * d - b * c;
      const double J00 = d / det, J01 = -b / det, J10 = -c / det, J11 = a / det;
      xv[0] = 0., xv[1] = 0., xv[2] = J00, xv[3] = J10, xv[4] = J01, xv[5] = J11;
   }
   // the C ABI takes (E, nu): invert lambda = E nu / ((1 + nu)(1 - 2 nu)), mu = E / (2 (1 + nu))
   static void young_poisson(double l, double m, double &E, double &nu)
   {
      nu = l / (2. * (l + m));
      E = m * (3. * l + 2. * m) / (l + m);
   }
   void stage(mfem::ElementTransformation &Tr, const mfem::Vector &elfun, double &nu)
   {
      double h[N_BUF] = {0.};
      vertices(Tr, h + O_X);
      Tr.SetIntPoint(&ip);
      const double l = lambda.Eval(Tr, ip), m = mu.Eval(Tr, ip), d = dam.Eval(Tr, ip);
      young_poisson(l, m, h[O_E], nu);
      h[O_D] = h[O_D + 1] = h[O_D + 2] = d;  // P1 damage field with the point value at the centroid
      for (int a = 0; a < 3; ++a) h[O_U + 2 * a] = elfun[a], h[O_U + 2 * a + 1] = elfun[3 + a];  // byNODES -> interleaved
      if (load && IntRule)
      {  // nodal force whose P1 interpolant reproduces the coefficient at the three points of the load rule:
         // f(q) = sum_a N_a(q) f_a  (N = barycentric coordinates), solved for f_a
         double N[3][3], fq[3][2];
         for (int q = 0; q < 3; ++q)
         {
            const mfem::IntegrationPoint &p = IntRule->IntPoint(q);
            N[q][0] = 1. - p.x - p.y, N[q][1] = p.x, N[q][2] = p.y;
            mfem::Vector f(2);
            load->Eval(f, Tr, p);
            fq[q][0] = f[0], fq[q][1] = f[1];
         }
         const double det = N[0][0] * (N[1][1] * N[2][2] - N[1][2] * N[2][1]) - N[0][1] * (N[1][0] * N[2][2] - N[1][2] * N[2][0]) +
                            N[0][2] * (N[1][0] * N[2][1] - N[1][1] * N[2][0]);
         for (int c = 0; c < 2; ++c)
            for (int a = 0; a < 3; ++a)
            {  // Cramer
               double M[3][3];
               for (int q = 0; q < 3; ++q)
                  for (int k = 0; k < 3; ++k) M[q][k] = (k == a) ? fq[q][c] : N[q][k];
               const double da = M[0][0] * (M[1][1] * M[2][2] - M[1][2] * M[2][1]) - M[0][1] * (M[1][0] * M[2][2] - M[1][2] * M[2][0]) +
                                 M[0][2] * (M[1][0] * M[2][1] - M[1][1] * M[2][0]);
               h[O_F + 2 * a + c] = da / det;
            }
      }
      cuda_check(cudaMemcpy(d_buf, h, sizeof(double) * O_OUT, cudaMemcpyHostToDevice), "femb200::DamIntegrator H2D");
   }

  public:
   // the reference's constructor (M.cc:493-498); variant: FEMB200_TANGENT_CLOSED, or _AD for a USE_AD build
   DamIntegrator(mfem::Coefficient &l, mfem::Coefficient &m, mfem::QuadratureFunctionCoefficient &d,
                 const mfem::IntegrationPoint &ip_, const mfem::IntegrationRule *ir,
                 mfem::VectorQuadratureFunctionCoefficient &load_, int variant_ = FEMB200_TANGENT_CLOSED)
       : mfem::NonlinearFormIntegrator(ir), lambda(l), mu(m), dam(d), ip(ip_), load(&load_), variant(variant_)
   {
      const int32_t id[3] = {0, 1, 2};
      cuda_check(cudaMalloc(reinterpret_cast<void **>(&d_map), sizeof(id)), "femb200::DamIntegrator");
      cuda_check(cudaMalloc(reinterpret_cast<void **>(&d_buf), sizeof(double) * N_BUF), "femb200::DamIntegrator");
      cuda_check(cudaMemcpy(d_map, id, sizeof(id), cudaMemcpyHostToDevice), "femb200::DamIntegrator");
      check(femb200_plan_create(FEMB200_P1, 3, 1, d_map, d_map, nullptr, &plan), "femb200_plan_create");
   }
   DamIntegrator(const DamIntegrator &) = delete;
   DamIntegrator &operator=(const DamIntegrator &) = delete;
   virtual ~DamIntegrator()
   {
      femb200_plan_destroy(plan);
      cudaFree(d_map);
      cudaFree(d_buf);
   }

   virtual double GetElementEnergy(const mfem::FiniteElement &, mfem::ElementTransformation &, const mfem::Vector &)
   {
      throw std::runtime_error("femb200::DamIntegrator::GetElementEnergy: not provided (the reference throws too, M.cc:552-557)");
   }

   // r_e = w gdshape sigma(u) - sum_q w_q N f  (M.cc:559-637), elvect byNODES
   virtual void AssembleElementVector(const mfem::FiniteElement &, mfem::ElementTransformation &Tr, const mfem::Vector &elfun,
                                      mfem::Vector &elvect)
   {
      double nu;
      stage(Tr, elfun, nu);
      check(femb200_assemble_vector(plan, d_buf + O_X, 2, d_buf + O_E, nu, d_buf + O_D, d_buf + O_U,
                                    (load && IntRule) ? d_buf + O_F : nullptr, d_buf + O_OUT, nullptr),
            "femb200_assemble_vector");
      double r[6];
      cuda_check(cudaMemcpy(r, d_buf + O_OUT, sizeof(r), cudaMemcpyDeviceToHost), "femb200::DamIntegrator D2H");
      elvect.SetSize(6);
      for (int a = 0; a < 3; ++a) elvect[a] = r[2 * a], elvect[3 + a] = r[2 * a + 1];
   }

   // K_e = w B D B^t  (M.cc:639-916), elmat column-major, dofs byNODES
   virtual void AssembleElementGrad(const mfem::FiniteElement &, mfem::ElementTransformation &Tr, const mfem::Vector &elfun,
                                    mfem::DenseMatrix &elmat)
   {
      double nu;
      stage(Tr, elfun, nu);
      check(femb200_element_grad_batched(FEMB200_P1, 1, d_buf + O_OUT, d_buf + O_X, 2, d_map, d_map, d_buf + O_E, nu,
                                         d_buf + O_D, d_buf + O_U, variant, nullptr),
            "femb200_element_grad_batched");
      elmat.SetSize(6);
      cuda_check(cudaMemcpy(elmat.GetData(), d_buf + O_OUT, sizeof(double) * 36, cudaMemcpyDeviceToHost),
                 "femb200::DamIntegrator D2H");
   }
};

// The assembled tangent as an operator: role of ParNonlinearForm::GetGradient (M.cc:1546) + HypreParMatrix::Mult
// inside CGSolver (M.cc:1502,1525-1528).  Host vectors in and out (mfem::Vector data), device-resident matrix.
class GradientOperator
{
   femb200_plan *plan = nullptr;
   int64_t nnodes, ncells, nnz = 0;
   int32_t *d_dofmap = nullptr, *d_xdofmap = nullptr;
   double *d_x = nullptr, *d_E = nullptr, *d_dnod = nullptr, *d_u = nullptr, *d_values = nullptr, *d_in = nullptr, *d_out = nullptr;
   uint8_t *d_bc = nullptr;
   double nu;
   int variant;

   template <typename T>
   static T *upload(const T *h, size_t n)
   {
      void *d = nullptr;
      cuda_check(cudaMalloc(&d, sizeof(T) * (n ? n : 1)), "femb200::GradientOperator");
      if (h) cuda_check(cudaMemcpy(d, h, sizeof(T) * n, cudaMemcpyHostToDevice), "femb200::GradientOperator H2D");
      return static_cast<T *>(d);
   }

  public:
   // etype FEMB200_P1 / _P2; coordinates nnodes x 2; dofmap ncells x nd, xdofmap ncells x 3 (node ids);
   // E per cell; damage per node or null; ess_dof_marker per vdof (byVDIM) or null
   GradientOperator(int etype, int64_t nnodes_, int64_t ncells_, const double *xy, const int32_t *dofmap, const int32_t *xdofmap,
                    const double *E, double nu_, const double *dnod, const uint8_t *ess_dof_marker,
                    int variant_ = FEMB200_TANGENT_CLOSED)
       : nnodes(nnodes_), ncells(ncells_), nu(nu_), variant(variant_)
   {
      const int nd = etype == FEMB200_P1 ? 3 : 6;
      d_dofmap = upload(dofmap, (size_t)ncells * nd);
      d_xdofmap = upload(xdofmap, (size_t)ncells * 3);
      d_x = upload(xy, (size_t)nnodes * 2);
      d_E = upload(E, (size_t)ncells);
      if (dnod) d_dnod = upload(dnod, (size_t)nnodes);
      d_u = upload<double>(nullptr, (size_t)nnodes * 2);
      d_in = upload<double>(nullptr, (size_t)nnodes * 2);
      d_out = upload<double>(nullptr, (size_t)nnodes * 2);
      check(femb200_plan_create(etype, nnodes, ncells, d_dofmap, d_xdofmap, nullptr, &plan), "femb200_plan_create");
      check(femb200_plan_sizes(plan, nullptr, nullptr, nullptr, &nnz, nullptr, nullptr), "femb200_plan_sizes");
      d_values = upload<double>(nullptr, (size_t)nnz);
      if (ess_dof_marker)
      {
         d_bc = upload(ess_dof_marker, (size_t)nnodes * 2);
         check(femb200_plan_set_dirichlet(plan, d_bc, nullptr), "femb200_plan_set_dirichlet");
      }
   }
   GradientOperator(const GradientOperator &) = delete;
   GradientOperator &operator=(const GradientOperator &) = delete;
   ~GradientOperator()
   {
      femb200_plan_destroy(plan);
      cudaFree(d_dofmap), cudaFree(d_xdofmap), cudaFree(d_x), cudaFree(d_E), cudaFree(d_dnod), cudaFree(d_u);
      cudaFree(d_values), cudaFree(d_in), cudaFree(d_out), cudaFree(d_bc);
   }
   int64_t Height() const { return 2 * nnodes; }
   int64_t NumNonZeros() const { return nnz; }
   // GetGradient(u): tangent at u with the essential rows / columns eliminated (unit diagonal)
   void Assemble(const mfem::Vector &u)
   {
      cuda_check(cudaMemcpy(d_u, u.GetData(), sizeof(double) * 2 * nnodes, cudaMemcpyHostToDevice), "femb200::GradientOperator H2D");
      check(femb200_assemble_matrix(plan, d_x, 2, d_E, nu, d_dnod, d_u, variant, d_values, nullptr), "femb200_assemble_matrix");
   }
   void Mult(const mfem::Vector &x, mfem::Vector &y) const
   {
      cuda_check(cudaMemcpy(d_in, x.GetData(), sizeof(double) * 2 * nnodes, cudaMemcpyHostToDevice), "femb200::GradientOperator H2D");
      check(femb200_spmv(plan, d_values, d_in, d_out, nullptr), "femb200_spmv");
      y.SetSize((int)(2 * nnodes));
      cuda_check(cudaMemcpy(y.GetData(), d_out, sizeof(double) * 2 * nnodes, cudaMemcpyDeviceToHost), "femb200::GradientOperator D2H");
   }
   // CGSolver::Mult with the reference's tolerances (M.cc:1525-1528), Jacobi instead of BoomerAMG
   int Solve(const mfem::Vector &b, mfem::Vector &x, double rel_tol = 1e-12, int max_iter = 2000) const
   {
      cuda_check(cudaMemcpy(d_in, b.GetData(), sizeof(double) * 2 * nnodes, cudaMemcpyHostToDevice), "femb200::GradientOperator H2D");
      int it = 0, conv = 0;
      double fin = 0.;
      check(femb200_cg(plan, FEMB200_OP_CSR, nullptr, d_values, d_in, d_out, 2 * nnodes, rel_tol, 0., max_iter,
                       FEMB200_PRECOND_JACOBI, &it, &fin, &conv, nullptr),
            "femb200_cg");
      x.SetSize((int)(2 * nnodes));
      cuda_check(cudaMemcpy(x.GetData(), d_out, sizeof(double) * 2 * nnodes, cudaMemcpyDeviceToHost), "femb200::GradientOperator D2H");
      return conv ? it : -it;
   }
};

}  // namespace femb200
#endif  // FEMB200_MFEM_HPP
