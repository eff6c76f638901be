// aliases.cu -- the entry points under the names SURVEY.md 8b gives the drop-in boundary
// (create_pattern, element_grad_batched, assemble_pa / add_mult_pa, cg): thin forms of the
// plan / pa / pcg functions for callers that bind the reference-side names directly.
#include "plan.cuh"

namespace femb {

__global__ void axpy1_kernel(int64_t n, const double *__restrict__ t, double *__restrict__ y)
{
   const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
   if (i < n) y[i] += t[i];
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

// BilinearFormIntegrator::AddMultPA(x, y): y += A x.  d_work: 2 * nnodes doubles of scratch.
extern "C" int femb200_add_mult_pa(const femb200_pa *pa, int64_t ndofs, const double *d_x, double *d_y, double *d_work,
                                   void *stream)
{
   FEMB_CHECK(pa && d_x && d_y && d_work, "add_mult_pa: null argument");
   FEMB_CHECK(d_work != d_x && d_work != d_y, "add_mult_pa: the scratch vector must not alias x or y");
   if (int rc = femb200_pa_apply(pa, d_x, d_work, stream)) return rc;
   axpy1_kernel<<<(unsigned)cdiv(ndofs, 256), 256, 0, as_stream(stream)>>>(ndofs, d_work, d_y);
   FEMB_LAUNCH_CHECK();
   return 0;
}

// CGSolver::Mult / KSPSolve with the tolerances of M.cc:1525-1528, F.cc:718-722; precond NONE or JACOBI.
// Allocates and frees its own scratch (4 n + 64 doubles); synchronises the stream.
extern "C" int femb200_cg(const femb200_plan *plan, int op_kind, const void *op, const double *d_values, const double *d_b,
                          double *d_x, int64_t n, double rtol, double atol, int maxit, int precond, int *iters,
                          double *final_res, int *converged, void *stream)
{
   FEMB_CHECK(precond == FEMB200_PRECOND_NONE || precond == FEMB200_PRECOND_JACOBI, "cg: unknown preconditioner %d", precond);
   FEMB_CHECK(n > 0 && d_b && d_x, "cg: null argument");
   cudaStream_t st = as_stream(stream);
   double *work = nullptr;
   FEMB_CUDA(cudaMalloc(&work, sizeof(double) * (4 * (size_t)n + 64)));
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
   cudaFree(work);
   if (iters) *iters = it;
   if (final_res) *final_res = fin;
   if (converged) *converged = conv;
   return rc;
}
