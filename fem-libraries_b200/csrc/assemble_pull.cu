// assemble_pull.cu -- output-centric ("pull") write-once assembly, fast path.
//
// Same role as assemble.cu (the setJ lambda, F.cc:847-862; ParNonlinearForm::
// GetGradient -> damIntegrator::AssembleElementGrad, M.cc:639-916), for straight-sided
// P1 / P2 triangles with a constant tangent per cell (d = 0: Hooke, M.cc:873-881).
//
// One thread owns one 2x2 node block K_IJ of the matrix and computes it completely:
//     K_IJ = sum over the (at most two) cells e containing both I and J of K_e[a, b]
// so nothing is staged, nothing is added twice into memory and there are no atomics:
// each thread issues two coalesced 16-byte stores (its slice of scalar row 2I and of
// row 2I+1).  Diagonal blocks (up to ~8 cells) are summed by one thread per node in
// a second phase of the same CTA.  Sums run in ascending cell order, i.e. in the
// order of the reference's serial cell loop: results are bit-reproducible.
//
// The element arithmetic is table driven and branch free.  With g_1, g_2 the
// gradients of the barycentric coordinates 1, 2 of the cell and
//     W(u, v) = |T| (lam u (x) v + mu v (x) u + mu (u . v) I)         (M.cc:699-704,885-887)
// the block of local dofs (a, b) is  K_e[a,b] = sum_q 2 w_q W(grad N_a(q), grad N_b(q)),
// and grad N_a(q) = dN_a/dxi(q) g_1 + dN_a/deta(q) g_2, hence
//     K_e[a,b] = lam M + mu M^t + mu tr(M) I,   M = sum_{c,d} C_ab[c][d] g_c (x) g_d,
//     C_ab[c][d] = sum_q 2 w_q dN_a/dxi_c(q) dN_b/dxi_d(q)
// a 2x2 table per (a, b) that depends on the element family only (36 entries for
// P2, 9 for P1; built on the host from the quadrature rule of element.cuh).
//
// A CTA owns kPullR consecutive node rows.  It first copies what its blocks need into
// shared memory with coalesced loads -- the 32-byte records sqrt(|T| E) (g_1, g_2)
// of the ~130 distinct incident cells (cell_setup_kernel pre-pass), the 16-bit visit
// entries, the row offsets -- so the per-block work touches global memory only for
// its 4-byte pull record and its two stores.
#include <algorithm>

#include "constitutive.cuh"
#include "element.cuh"
#include "plan.cuh"

namespace femb {

constexpr int kPullThreads = 256;

struct PullArgs
{
   int64_t nnodes;
   const int64_t *brp;
   const int32_t *nptr, *tile_cptr, *tile_cells;
   const uint16_t *vis16;
   const uint32_t *pull;
   const double *cellrec;
   const double *table;  // [nd*nd][4] device copy of C_ab
   double *values;
   int max_cells, max_visits;
   double c2, c3;
};

__device__ __forceinline__ void add_block(const double2 *s_cell, const double2 *s_tab, double2 lm, int cl, int ab,
                                          double &k00, double &k01, double &k10, double &k11)
{
   const double2 g1 = s_cell[2 * cl], g2 = s_cell[2 * cl + 1];
   const double2 c0 = s_tab[2 * ab], c1 = s_tab[2 * ab + 1];  // (C11, C12), (C21, C22)
   const double p1x = c0.x * g1.x + c1.x * g2.x, p1y = c0.x * g1.y + c1.x * g2.y;
   const double p2x = c0.y * g1.x + c1.y * g2.x, p2y = c0.y * g1.y + c1.y * g2.y;
   const double m00 = p1x * g1.x + p2x * g2.x, m01 = p1x * g1.y + p2x * g2.y;
   const double m10 = p1y * g1.x + p2y * g2.x, m11 = p1y * g1.y + p2y * g2.y;
   const double tr = m00 + m11;
   k00 += lm.x * m00 + lm.y * (m00 + tr);
   k01 += lm.x * m01 + lm.y * m10;
   k10 += lm.x * m10 + lm.y * m01;
   k11 += lm.x * m11 + lm.y * (m11 + tr);
}

template <int ET>
__global__ void __launch_bounds__(kPullThreads) assemble_pull_kernel(PullArgs P)
{
   constexpr int nd = Elem<ET>::nd, R = kPullR;
   extern __shared__ __align__(16) unsigned char smem[];
   double2 *s_cell = reinterpret_cast<double2 *>(smem);
   double2 *s_tab = s_cell + 2 * P.max_cells;
   const double2 lm = make_double2(P.c2, P.c3);  // lambda / E, mu / E (M.cc:1087-1098)
   int32_t *s_brp = reinterpret_cast<int32_t *>(s_tab + 2 * nd * nd);  // [R + 1] block offsets in the tile
   int32_t *s_k0 = s_brp + (R + 1);                                    // [R + 1] visit offsets in the tile
   int32_t *s_diag = s_k0 + (R + 1);                                   // [R] diagonal block of every node
   uint16_t *s_vis = reinterpret_cast<uint16_t *>(s_diag + R);         // [max_visits]
   const int tid = threadIdx.x;
   const int64_t n0 = (int64_t)blockIdx.x * R;
   const int nloc = (int)min((int64_t)R, P.nnodes - n0);
   // level 1: tile extents (three independent loads)
   const int64_t b0 = P.brp[n0];
   const int32_t k0 = P.nptr[n0];
   const int32_t c0 = P.tile_cptr[blockIdx.x];
   const int nb = (int)(P.brp[n0 + nloc] - b0);
   const int nv = P.nptr[n0 + nloc] - k0;
   const int nc = P.tile_cptr[blockIdx.x + 1] - c0;
   // level 2: everything addressed by the extents is requested before anything is consumed:
   // the pull records of this thread's blocks (registers), the cell ids, then the tile tables
   constexpr int PB = 4;  // blocks per thread and pass
   uint32_t recs[PB];
#pragma unroll
   for (int j = 0; j < PB; ++j)
   {
      const int blk = tid + j * kPullThreads;
      recs[j] = blk < nb ? P.pull[b0 + blk] : 0xffffffffu;
   }
   int32_t mycell[2];
#pragma unroll
   for (int j = 0; j < 2; ++j)
   {
      const int c = tid + j * kPullThreads;
      mycell[j] = c < nc ? P.tile_cells[c0 + c] : -1;
   }
   for (int i = tid; i <= nloc; i += kPullThreads)
   {
      s_brp[i] = (int32_t)(P.brp[n0 + i] - b0);
      s_k0[i] = P.nptr[n0 + i] - k0;
   }
   for (int i = tid; i < nv; i += kPullThreads) s_vis[i] = P.vis16[k0 + i];
   for (int i = tid; i < 2 * nd * nd; i += kPullThreads) s_tab[i] = reinterpret_cast<const double2 *>(P.table)[i];
   {  // level 3: the 32-byte records of the tile's cells
      const double2 *rec = reinterpret_cast<const double2 *>(P.cellrec);
#pragma unroll
      for (int j = 0; j < 2; ++j)
         if (mycell[j] >= 0)
         {
            const int c = tid + j * kPullThreads;
            const double2 r0 = rec[2 * (int64_t)mycell[j]], r1 = rec[2 * (int64_t)mycell[j] + 1];
            s_cell[2 * c] = r0, s_cell[2 * c + 1] = r1;
         }
      for (int c = tid + 2 * kPullThreads; c < nc; c += kPullThreads)
      {
         const int64_t g = P.tile_cells[c0 + c];
         s_cell[2 * c] = rec[2 * g], s_cell[2 * c + 1] = rec[2 * g + 1];
      }
   }
   __syncthreads();
   double *dst = P.values + 4 * b0;
   // phase A: off-diagonal blocks, one thread per block
   for (int base = 0; base < nb; base += PB * kPullThreads)
   {
      if (base > 0)
      {
#pragma unroll
         for (int j = 0; j < PB; ++j)
         {
            const int blk = base + tid + j * kPullThreads;
            recs[j] = blk < nb ? P.pull[b0 + blk] : 0xffffffffu;
         }
      }
#pragma unroll
      for (int j = 0; j < PB; ++j)
      {
         const int blk = base + tid + j * kPullThreads;
         const uint32_t rec = recs[j];
         if (rec == 0xffffffffu) continue;
         const int il = rec & 0xffu;
         const uint32_t ca = (rec >> 8) & 0xffu, cb = (rec >> 16) & 0xffu;
         if (ca == 0xfeu)
         {
            s_diag[il] = blk;
            continue;
         }
         const int kb = s_k0[il];
         double k00 = 0., k01 = 0., k10 = 0., k11 = 0.;
         if (ca != 0xffu)
         {
            const uint32_t v = s_vis[kb + (ca >> 4)];
            add_block(s_cell, s_tab, lm, v >> 4, (v & 15u) * nd + (ca & 15u), k00, k01, k10, k11);
         }
         if (cb != 0xffu)
         {
            const uint32_t v = s_vis[kb + (cb >> 4)];
            add_block(s_cell, s_tab, lm, v >> 4, (v & 15u) * nd + (cb & 15u), k00, k01, k10, k11);
         }
         st_stream_d2(dst + 2 * (int64_t)(s_brp[il] + blk), make_double2(k00, k01));
         st_stream_d2(dst + 2 * (int64_t)(s_brp[il + 1] + blk), make_double2(k10, k11));
      }
   }
   __syncthreads();
   // phase B: diagonal blocks, one thread per node
   for (int il = tid; il < nloc; il += kPullThreads)
   {
      const int ke = s_k0[il + 1];
      if (ke == s_k0[il]) continue;  // a node without cells has no row entries
      double k00 = 0., k01 = 0., k10 = 0., k11 = 0.;
      for (int k = s_k0[il]; k < ke; ++k)
      {
         const uint32_t v = s_vis[k];
         const int a = v & 15u;
         add_block(s_cell, s_tab, lm, v >> 4, a * nd + a, k00, k01, k10, k11);
      }
      const int blk = s_diag[il];
      st_stream_d2(dst + 2 * (int64_t)(s_brp[il] + blk), make_double2(k00, k01));
      st_stream_d2(dst + 2 * (int64_t)(s_brp[il + 1] + blk), make_double2(k10, k11));
   }
}

// C_ab[c][d] = sum_q 2 w_q dN_a/dxi_c(q) dN_b/dxi_d(q) for the P1 / P2 rules of element.cuh
static void host_table(int etype, double *tab /*[nd*nd][4]*/)
{
   const int nd = elem_nd(etype), nq = elem_nq(etype);
   for (int i = 0; i < nd * nd * 4; ++i) tab[i] = 0.;
   for (int q = 0; q < nq; ++q)
   {
      double xi, eta, w, dN[6][2];
      if (etype == FEMB200_P1)
      {
         xi = eta = 1. / 3., w = 0.5;
         dN[0][0] = -1., dN[0][1] = -1., dN[1][0] = 1., dN[1][1] = 0., dN[2][0] = 0., dN[2][1] = 1.;
      }
      else
      {
         xi = (q == 1) ? 2. / 3. : 1. / 6., eta = (q == 2) ? 2. / 3. : 1. / 6., w = 1. / 6.;
         const double L0 = 1. - xi - eta, L1 = xi, L2 = eta;
         dN[0][0] = -(4. * L0 - 1.), dN[0][1] = -(4. * L0 - 1.);
         dN[1][0] = (4. * L1 - 1.), dN[1][1] = 0.;
         dN[2][0] = 0., dN[2][1] = (4. * L2 - 1.);
         dN[3][0] = 4. * L2, dN[3][1] = 4. * L1;
         dN[4][0] = -4. * L2, dN[4][1] = 4. * (L0 - L2);
         dN[5][0] = 4. * (L0 - L1), dN[5][1] = -4. * L1;
      }
      for (int a = 0; a < nd; ++a)
         for (int b = 0; b < nd; ++b)
            for (int c = 0; c < 2; ++c)
               for (int d = 0; d < 2; ++d) tab[(a * nd + b) * 4 + c * 2 + d] += 2. * w * dN[a][c] * dN[b][d];
   }
}

static double *device_table(int etype)
{
   static double *d_tab[2] = {nullptr, nullptr};
   const int k = etype == FEMB200_P1 ? 0 : 1;
   if (!d_tab[k])
   {
      double h[36 * 4];
      host_table(etype, h);
      const int nd = elem_nd(etype);
      if (cudaMalloc(&d_tab[k], sizeof(double) * 4 * nd * nd) != cudaSuccess) return nullptr;
      cudaMemcpy(d_tab[k], h, sizeof(double) * 4 * nd * nd, cudaMemcpyHostToDevice);
   }
   return d_tab[k];
}

// returns -1 when the plan has no pull data (caller falls back to the staged kernel)
int launch_assemble_pull(const femb200_plan *p, const double *cellrec, LameCoef lc, double *d_values, cudaStream_t st)
{
   if (!p->pull_ok || p->etype == FEMB200_Q2) return -1;
   PullArgs P;
   P.nnodes = p->nnodes, P.brp = p->brp, P.nptr = p->nptr, P.tile_cptr = p->tile_cptr, P.tile_cells = p->tile_cells;
   P.vis16 = p->vis16, P.pull = p->pull, P.cellrec = cellrec, P.values = d_values;
   P.max_cells = p->pull_max_cells, P.max_visits = (p->pull_max_visits + 7) & ~7;
   P.table = device_table(p->etype);
   P.c2 = lc.c2, P.c3 = lc.c3;
   FEMB_CHECK(P.table != nullptr, "assemble: cannot allocate the coefficient table");
   const int nd = p->nd;
   const size_t smem = 32 * (size_t)P.max_cells + 32 * (size_t)nd * nd + 4 * (size_t)(3 * kPullR + 2) +
                       2 * (size_t)P.max_visits + 16;
   const size_t budget = devinfo().smem_optin ? devinfo().smem_optin : 227 * 1024;
   if (smem > budget) return -1;
   const unsigned grid = (unsigned)p->ntiles;
   if (p->etype == FEMB200_P1)
   {
      FEMB_CUDA(cudaFuncSetAttribute(assemble_pull_kernel<FEMB200_P1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      assemble_pull_kernel<FEMB200_P1><<<grid, kPullThreads, smem, st>>>(P);
   }
   else
   {
      FEMB_CUDA(cudaFuncSetAttribute(assemble_pull_kernel<FEMB200_P2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      assemble_pull_kernel<FEMB200_P2><<<grid, kPullThreads, smem, st>>>(P);
   }
   FEMB_LAUNCH_CHECK();
   return 0;
}

}  // namespace femb
