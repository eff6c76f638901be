// aliases.cu -- the entry points under the names SURVEY.md 8b gives the drop-in boundary
// (create_pattern, element_grad_batched, assemble_pa / add_mult_pa, cg): thin forms of the
// plan / pa / pcg functions for callers that bind the reference-side names directly.
#include "plan.cuh"

#include <map>
#include <mutex>

namespace femb {

// scratch of femb200_cg: one buffer per device, grown on demand, kept for the life of the process (a solver
// called once per Newton iteration must not pay a cudaMalloc / cudaFree pair every time)
static int cg_scratch(size_t doubles, double **out)
{
   struct Slot
   {
      double *p = nullptr;
      size_t cap = 0;
   };
   static std::map<int, Slot> slots;
   static std::mutex mtx;
   std::lock_guard<std::mutex> lock(mtx);
   int dev = 0;
   FEMB_CUDA(cudaGetDevice(&dev));
   Slot &s = slots[dev];
   if (doubles > s.cap)
   {
      FEMB_CUDA(cudaDeviceSynchronize());
      cudaFree(s.p);
      s.p = nullptr, s.cap = 0;
      FEMB_CUDA(cudaMalloc(&s.p, sizeof(double) * doubles));
      s.cap = doubles;
   }
   *out = s.p;
   return 0;
}

}  // namespace femb

using namespace femb;

// dolfinx::fem::petsc::create_matrix(*J_form), F.cc:688
extern "C" int femb200_create_pattern(int etype, int64_t nnodes, int64_t ncells, const int32_t *d_dofmap,
                                      const int32_t *d_xdofmap, void *stream, femb200_plan **out)
{
   return femb200_plan_create(etype, nnodes, ncells, d_dofmap, d_xdofmap, stream, out);
}

// damIntegrator::AssembleElementGrad for a batch of elements (M.cc:639-916): elmat column-major, byNODES
extern "C" int femb200_element_grad_batched(int etype, int64_t ncells, double *d_elmat, const double *d_x, int x_stride,
                                            const int32_t *d_xdofmap, const int32_t *d_dofmap, const double *d_E,
                                            double nu, const double *d_dnod, const double *d_u, int variant, void *stream)
{
   return femb200_tabulate_tensor_batched(etype, ncells, d_elmat, d_x, x_stride, d_xdofmap, d_dofmap, d_E, nu, d_dnod,
                                          d_u, variant, FEMB200_COLMAJOR_BYNODES, stream);
}

// BilinearFormIntegrator::AssemblePA(fes)
extern "C" int femb200_assemble_pa(int etype, int64_t nnodes, int64_t ncells, const int32_t *d_dofmap,
                                   const int32_t *d_xdofmap, const double *d_x, int x_stride, const double *d_E, double nu,
                                   void *stream, femb200_pa **out)
{
   return femb200_pa_create(etype, nnodes, ncells, d_dofmap, d_xdofmap, d_x, x_stride, d_E, nu, stream, out);
}

// CGSolver::Mult / KSPSolve with the tolerances of M.cc:1525-1528, F.cc:718-722; precond NONE or JACOBI.
// Its scratch (4 n + 64 doubles) comes from a per-device buffer that is grown on demand and kept; synchronises the
// stream.  Not re-entrant on one device (one solve at a time, as the reference's Newton loop).
extern "C" int femb200_cg(const femb200_plan *plan, int op_kind, const void *op, const double *d_values, const double *d_b,
                          double *d_x, int64_t n, double rtol, double atol, int maxit, int precond, int *iters,
                          double *final_res, int *converged, void *stream)
{
   FEMB_CHECK(precond == FEMB200_PRECOND_NONE || precond == FEMB200_PRECOND_JACOBI, "cg: unknown preconditioner %d", precond);
   FEMB_CHECK(n > 0 && d_b && d_x, "cg: null argument");
   cudaStream_t st = as_stream(stream);
   double *work = nullptr;
   if (int rc0 = cg_scratch(4 * (size_t)n + 64, &work)) return rc0;
   double *dinv = nullptr;
   int rc = 0;
   if (precond == FEMB200_PRECOND_JACOBI)
   {
      dinv = work + 3 * (size_t)n + 64;
      rc = op_kind == FEMB200_OP_PA ? femb200_pa_diagonal(static_cast<const femb200_pa *>(op), dinv, stream)
                                    : femb200_extract_diagonal(plan, d_values, dinv, stream);
      if (!rc) rc = femb200_jacobi_setup(n, dinv, dinv, stream);
   }
   int it = 0, conv = 0;
   double fin = 0.;
   if (!rc)
      rc = femb200_pcg(plan, op_kind, op, d_values, d_b, d_x, n, rtol, atol, maxit, dinv, 25, 0, work, &it, &fin, &conv, stream);
   cudaStreamSynchronize(st);
   if (iters) *iters = it;
   if (final_res) *final_res = fin;
   if (converged) *converged = conv;
   return rc;
}
