// assemble.cu -- write-once gather assembly of the global tangent matrix.
//
// Role in the reference: the setJ lambda (F.cc:847-862) = MatZeroEntries +
// dolfinx assemble_matrix(set_block_fn(A, ADD_VALUES), J, bcs) + set_diagonal +
// MatAssembly, i.e. per cell: tabulate_tensor -> zero Dirichlet rows/cols ->
// MatSetValuesBlockedLocal (binary searches, read-modify-write of the CSR); on the
// MFEM side ParNonlinearForm::GetGradient -> AssembleElementGrad (M.cc:639-916) ->
// SparseMatrix::AddSubMatrix.
//
// B200 design: no scatter, no atomics, no zero-fill.  One thread owns one block
// row (node I) of the matrix.  It walks the cells incident to I (VisitRec list of
// the plan), computes only the 2 x 2nd row slice of each element matrix that
// belongs to I, and accumulates it into a shared-memory image of its CSR rows
// (first contribution stores, later ones add: the plan pre-computes which is
// which).  A CTA owns R consecutive nodes, whose CSR values are one contiguous
// byte range: after the tile is complete it is streamed out with 16-byte coalesced
// stores, every value written exactly once.  HBM traffic = nnz*8 B written +
// ~120 B/cell of maps and geometry read (kernel K1+K3).
//
// Fast path (P1/P2, d = 0): on a straight-sided triangle with constant D the P2
// element matrix is an exact linear combination of the nine 2x2 blocks
//   W^{cd} = |T| g_c^t C g_d ,  g_c = grad lambda_c   (c,d = vertices)
// (these are the P1 stiffness blocks, M.cc:699-704,885-887 with w = |T|):
//   vertex a / vertex b :  W^{aa}  or  -W^{ab}/3
//   vertex p / edge (p,q):  4/3 W^{pq};  vertex / opposite edge: 0
//   edge (p,q) / edge (r,s): 4/3 [(1+d_qs) W^{pr} + (1+d_qr) W^{ps} + (1+d_ps) W^{qr} + (1+d_pr) W^{qs}]
// which is what the 3-point rule integrates exactly.  The row
// owner relabels the triangle cyclically so that its own local index is vertex 0
// / edge 0: only 3 W blocks are needed per visit.  Damaged tangents (d > 0, D
// varying per point) and Q2 take the generic per-quadrature-point path.
#include <algorithm>

#include "constitutive.cuh"
#include "element.cuh"
#include "plan.cuh"
#include "reduce.cuh"

namespace femb {

// staging swizzle on 16-byte units: spreads the systematically aligned row starts
// of neighbouring threads over different bank groups; an involution inside
// aligned groups of 8 units, so the stream-out stays coalesced.
__device__ __forceinline__ int swz(int u) { return u ^ ((u >> 3) & 7); }

__device__ __forceinline__ void stage_put(double2 *sv, int unit, double a, double b, bool first)
{
   double2 *p = sv + swz(unit);
   if (first)
      *p = make_double2(a, b);
   else
   {
      double2 v = *p;
      v.x += a;
      v.y += b;
      *p = v;
   }
}

// one 2x2 block into (row0 unit base r0, row1 unit base r1), slot s
__device__ __forceinline__ void stage_block(double2 *sv, int r0, int r1, int s, const double *k, bool first)
{
   stage_put(sv, r0 + s, k[0], k[1], first);
   stage_put(sv, r1 + s, k[2], k[3], first);
}

// W^{cd} for Hooke: |T| (lam g_c (x) g_d + mu g_d (x) g_c + mu (g_c . g_d) I)
__device__ __forceinline__ void w_block(const double *gc, const double *gd, double tl, double tm, double *w)
{
   const double xx = gc[0] * gd[0], yy = gc[1] * gd[1], xy = gc[0] * gd[1], yx = gc[1] * gd[0];
   w[0] = (tl + 2. * tm) * xx + tm * yy;
   w[1] = tl * xy + tm * yx;
   w[2] = tl * yx + tm * xy;
   w[3] = (tl + 2. * tm) * yy + tm * xx;
}

struct AsmArgs
{
   int64_t nnodes;
   const int32_t *nptr;
   const VisitRec *vrec;
   const int64_t *brp;
   const int32_t *xdofmap, *dofmap;
   const double *x;
   int xs;
   const double *E;
   LameCoef lc;
   const double *dnod, *u;
   int variant;
   double *values;
   int stage_units;  // capacity of the staging image in 16-byte units (multiple of 8)
};

// ---- fast path: straight-sided P1 / P2 triangle, linear elasticity --------------
// kr[t] = 2x2 block (row node = local dof a of the visit, column = ROTATED local dof t:
// vertices (m, m1, m2) then edges (3+m, 3+m1, 3+m2), see rotated_index)
__device__ __forceinline__ int rotated_index(int t, int a)
{
   const int m = (a >= 3) ? a - 3 : a;
   const int tt = (t >= 3) ? t - 3 : t;
   int b = m + tt;
   b = (b >= 3) ? b - 3 : b;
   return (t >= 3) ? b + 3 : b;
}

template <int ET>
__device__ __forceinline__ void visit_fast(const AsmArgs &A, const Visit &r, double (*kr)[4])
{
   const int a = r.a;
   const int m = (a >= 3) ? a - 3 : a;  // rotation: own vertex / own edge becomes number 0
   const int64_t e = r.e;
   const int32_t *xd = A.xdofmap + e * 3;
   const int m1 = (m + 1 >= 3) ? m - 2 : m + 1, m2 = (m + 2 >= 3) ? m - 1 : m + 2;
   const int32_t xd0 = xd[0], xd1 = xd[1], xd2 = xd[2];
   const int64_t v0 = (m == 0) ? xd0 : (m == 1 ? xd1 : xd2);
   const int64_t v1 = (m1 == 0) ? xd0 : (m1 == 1 ? xd1 : xd2);
   const int64_t v2 = (m2 == 0) ? xd0 : (m2 == 1 ? xd1 : xd2);
   const double x0 = A.x[v0 * A.xs], y0 = A.x[v0 * A.xs + 1];
   const double x1 = A.x[v1 * A.xs], y1 = A.x[v1 * A.xs + 1];
   const double x2 = A.x[v2 * A.xs], y2 = A.x[v2 * A.xs + 1];
   const double Ee = A.E[e];
   const double det = (x1 - x0) * (y2 - y0) - (x2 - x0) * (y1 - y0);
   const double id = 1. / det;
   const double g1[2] = {(y2 - y0) * id, -(x2 - x0) * id};
   const double g2[2] = {-(y1 - y0) * id, (x1 - x0) * id};
   const double T = 0.5 * fabs(det);
   const double tl = T * Ee * A.lc.c2, tm = T * Ee * A.lc.c3;
   if (a < 3)
   {  // row = (rotated) vertex 0
      const double g0[2] = {-g1[0] - g2[0], -g1[1] - g2[1]};
      double w00[4], w01[4], w02[4];
      w_block(g0, g0, tl, tm, w00);
      w_block(g0, g1, tl, tm, w01);
      w_block(g0, g2, tl, tm, w02);
      if (ET == FEMB200_P1)
      {  // P1: K_ab = W^{ab}  (M.cc:885-887)
#pragma unroll
         for (int i = 0; i < 4; ++i) kr[0][i] = w00[i], kr[1][i] = w01[i], kr[2][i] = w02[i];
      }
      else
      {
         const double c3 = -1. / 3., c43 = 4. / 3.;
#pragma unroll
         for (int i = 0; i < 4; ++i)
         {
            kr[0][i] = w00[i];
            kr[1][i] = c3 * w01[i];
            kr[2][i] = c3 * w02[i];
            kr[3][i] = 0.;  // rotated edges: 0' = (1,2) opposite (structural zero), 1' = (2,0), 2' = (0,1)
            kr[4][i] = c43 * w02[i];
            kr[5][i] = c43 * w01[i];
         }
      }
   }
   else
   {  // row = (rotated) edge 0 = (1,2); uses sum_d W^{cd} = 0 to stay within W11, W12, W22
      double w11[4], w12[4], w22[4];
      w_block(g1, g1, tl, tm, w11);
      w_block(g1, g2, tl, tm, w12);
      w_block(g2, g2, tl, tm, w22);
      const double c43 = 4. / 3.;
      // vertices: opposite 0' -> 0; 1' (= p) -> 4/3 W^{21} = 4/3 (W^{12})^t; 2' (= q) -> 4/3 W^{12}
      kr[0][0] = kr[0][1] = kr[0][2] = kr[0][3] = 0.;
      kr[1][0] = c43 * w12[0], kr[1][1] = c43 * w12[2], kr[1][2] = c43 * w12[1], kr[1][3] = c43 * w12[3];
      // S = W12 + W21 (symmetric)
      const double sy[4] = {2. * w12[0], w12[1] + w12[2], w12[1] + w12[2], 2. * w12[3]};
#pragma unroll
      for (int i = 0; i < 4; ++i)
      {
         kr[2][i] = c43 * w12[i];
         kr[3][i] = c43 * (2. * w11[i] + sy[i] + 2. * w22[i]);
         kr[4][i] = -c43 * (2. * w11[i] + sy[i]);
         kr[5][i] = -c43 * (sy[i] + 2. * w22[i]);
      }
   }
}

// ---- generic path: per-quadrature-point loop, damaged tangent, any family -------
template <int ET>
__device__ inline void visit_generic(const AsmArgs &A, const Visit &r, double (*kb)[4])
{
   constexpr int nd = Elem<ET>::nd, nv = Elem<ET>::nv, nq = Elem<ET>::nq;
   const int a = r.a;
   const int64_t e = r.e;
   double xv[nv][2], dv[nv];
#pragma unroll
   for (int v = 0; v < nv; ++v)
   {
      const int64_t g = A.xdofmap[e * nv + v];
      xv[v][0] = A.x[g * A.xs];
      xv[v][1] = A.x[g * A.xs + 1];
      dv[v] = A.dnod ? A.dnod[g] : 0.;
   }
   const double Ee = A.E[e];
   const double lam = Ee * A.lc.c2, mu = Ee * A.lc.c3;
#pragma unroll
   for (int b = 0; b < nd; ++b) kb[b][0] = kb[b][1] = kb[b][2] = kb[b][3] = 0.;
#pragma unroll 1
   for (int q = 0; q < nq; ++q)
   {
      double G[nd][2], phi[nv], D[9];
      const double w = qp_geometry<ET>(xv, q, G, phi);
      double d = 0.;
#pragma unroll
      for (int v = 0; v < nv; ++v) d += phi[v] * dv[v];
      if (d > 0.)
      {
         double g00 = 0., g01 = 0., g10 = 0., g11 = 0.;
         if (A.u)
#pragma unroll
            for (int b = 0; b < nd; ++b)
            {
               const int64_t gd = 2 * (int64_t)A.dofmap[e * nd + b];
               const double ux = A.u[gd], uy = A.u[gd + 1];
               g00 += ux * G[b][0];
               g01 += ux * G[b][1];
               g10 += uy * G[b][0];
               g11 += uy * G[b][1];
            }
         const double s = 0.5 * (g01 + g10);
         const double eps[4] = {g00, s, s, g11};
         tangent(A.variant, lam, mu, d, eps, D);
      }
      else
         hooke_scaled(lam, mu, 1., D);
      double ga[2] = {0., 0.};
#pragma unroll
      for (int b = 0; b < nd; ++b)
         if (b == a) ga[0] = G[b][0], ga[1] = G[b][1];
#pragma unroll
      for (int b = 0; b < nd; ++b) bdb_block(ga, G[b], D, w, kb[b]);
   }
}

// One thread per VISIT (row node I, incident cell e): all visits of a tile of R
// consecutive node rows are computed concurrently, then merged into the staging
// image in `maxcnt` rounds: round q stages the q-th visit of every node, so a
// slot is never touched by two threads at once and every sum runs in ascending
// cell order (the order of the reference's serial cell loop).
template <int ET, bool FAST, int THREADS>
__global__ void __launch_bounds__(THREADS) assemble_kernel(AsmArgs A, int R)
{
   constexpr int nd = Elem<ET>::nd;
   extern __shared__ double2 sv[];
   __shared__ int s_maxcnt;
   const int tid = threadIdx.x;
   const int64_t n0 = (int64_t)blockIdx.x * R;
   const int nloc = (int)min((int64_t)R, A.nnodes - n0);
   const int64_t b0 = A.brp[n0];
   const int units = 2 * (int)(A.brp[n0 + nloc] - b0);  // 16-byte units of this tile
   // tile metadata behind the staging area: visit offsets and row unit offsets
   int32_t *s_nptr = reinterpret_cast<int32_t *>(sv + A.stage_units);
   int32_t *s_roff = s_nptr + (R + 1);
   if (tid == 0) s_maxcnt = 0;
   for (int i = tid; i <= nloc; i += THREADS)
   {
      s_nptr[i] = A.nptr[n0 + i];
      s_roff[i] = 2 * (int)(A.brp[n0 + i] - b0);
   }
   __syncthreads();
   for (int i = tid; i < nloc; i += THREADS) atomicMax(&s_maxcnt, s_nptr[i + 1] - s_nptr[i]);
   const int32_t k0 = s_nptr[0];
   const int nvis = s_nptr[nloc] - k0;
   __syncthreads();
   const int maxcnt = s_maxcnt;
   for (int base = 0; base < nvis; base += THREADS)
   {
      const int v = base + tid;
      const bool active = v < nvis;
      double kb[nd][4];
      int rank = -1, r0 = 0, r1 = 0;
      uint4 raw = make_uint4(0u, 0u, 0u, 0u);
      if (active)
      {
         const int32_t k = k0 + v;
         raw = *reinterpret_cast<const uint4 *>(A.vrec + k);
         int lo = 0, hi = nloc;  // largest i with s_nptr[i] <= k
         while (hi - lo > 1)
         {
            const int mid = (lo + hi) >> 1;
            if (s_nptr[mid] <= k)
               lo = mid;
            else
               hi = mid;
         }
         rank = k - s_nptr[lo];
         r0 = s_roff[lo];
         r1 = r0 + ((s_roff[lo + 1] - r0) >> 1);
      }
      const Visit r(raw);
      if (active)
      {
         if (FAST)
            visit_fast<ET>(A, r, kb);
         else
            visit_generic<ET>(A, r, kb);
      }
      for (int q = 0; q < maxcnt; ++q)
      {
         if (rank == q)
         {
#pragma unroll
            for (int t = 0; t < nd; ++t)
            {
               const int b = FAST ? rotated_index(t, (int)r.a) : t;
               stage_block(sv, r0, r1, r.slot(b), kb[t], r.is_first(b));
            }
         }
         __syncthreads();
      }
   }
   // stream the finished tile out: one contiguous byte range of the CSR values
   double *dst = A.values + 4 * b0;
   const int padded = (units + 7) & ~7;
   for (int i = tid; i < padded; i += THREADS)
   {
      const int u = swz(i);
      if (u < units) st_stream_d2(dst + 2 * (int64_t)u, sv[i]);
   }
}

// ---- Dirichlet rows / columns / diagonal (F.cc:852-857) --------------------------
// one warp per constrained node I; lanes over the blocks of row I.  Zeroing summed
// values equals summing zeroed element contributions exactly, so this matches the
// element-level treatment of dolfinx bit for bit.
__global__ void dirichlet_kernel(int nbc, const int32_t *__restrict__ bc_nodes, const uint8_t *__restrict__ bc,
                                 const int64_t *__restrict__ brp, const int32_t *__restrict__ bcol,
                                 double *__restrict__ values, double diag)
{
   const int w = (int)(((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
   if (w >= nbc) return;
   const int64_t I = bc_nodes[w];
   const bool m0 = bc[2 * I], m1 = bc[2 * I + 1];
   const int64_t bi = brp[I];
   const int deg = (int)(brp[I + 1] - bi);
   double *row0 = values + 4 * bi, *row1 = row0 + 2 * deg;
   int self = -1;
   for (int s = lane; s < deg; s += 32)
   {
      const int64_t J = bcol[bi + s];
      if (J == I) self = s;
      if (m0) row0[2 * s] = row0[2 * s + 1] = 0.;
      if (m1) row1[2 * s] = row1[2 * s + 1] = 0.;
      // column I of row J (the pattern is symmetric)
      const int64_t bj = brp[J];
      const int degj = (int)(brp[J + 1] - bj);
      int lo = 0, hi = degj;
      while (lo < hi)
      {
         const int mid = (lo + hi) >> 1;
         if (bcol[bj + mid] < I)
            lo = mid + 1;
         else
            hi = mid;
      }
      if (lo < degj && bcol[bj + lo] == I)
      {
         double *c0 = values + 4 * bj, *c1 = c0 + 2 * degj;
         if (m0) c0[2 * lo] = c1[2 * lo] = 0.;
         if (m1) c0[2 * lo + 1] = c1[2 * lo + 1] = 0.;
      }
   }
   __syncwarp();
   if (self >= 0)
   {  // set_diagonal(..., diag) with INSERT_VALUES (F.cc:857)
      if (m0) row0[2 * self] = diag;
      if (m1) row1[2 * self + 1] = diag;
   }
}

// ---- Frobenius norm^2 and trace --------------------------------------------------
__global__ void __launch_bounds__(256)
norms_kernel(int64_t nnodes, const int64_t *__restrict__ brp, const int32_t *__restrict__ bcol,
             const double *__restrict__ values, ReduceScratch red, double *__restrict__ out)
{
   double acc[2] = {0., 0.};  // fro^2, trace
   const int64_t nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
   const int lane = threadIdx.x & 31;
   for (int64_t I = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; I < nnodes; I += nw)
   {
      const int64_t bi = brp[I];
      const int deg = (int)(brp[I + 1] - bi);
      const double2 *row = reinterpret_cast<const double2 *>(values + 4 * bi);
      for (int t = lane; t < 2 * deg; t += 32)
      {
         const double2 v = row[t];
         acc[0] += v.x * v.x + v.y * v.y;
      }
      for (int s = lane; s < deg; s += 32)
         if (bcol[bi + s] == I) acc[1] += values[4 * bi + 2 * s] + values[4 * bi + 2 * deg + 2 * s + 1];
   }
   block_reduce_finish_n<256, 2>(acc, red, out);
}

template <int ET, bool FAST>
static int launch_assemble(const femb200_plan *p, AsmArgs A, cudaStream_t st)
{
   constexpr int THREADS = FAST ? 256 : 128;
   // tile height R (node rows per CTA): about one visit per thread
   const size_t budget = devinfo().smem_optin ? devinfo().smem_optin : 227 * 1024;
   const double vis_per_node = (double)p->nvisits / (double)p->nnodes;
   int best = 0;
   for (int r = 0; r < kNumTileR; ++r)
   {
      const size_t bytes = 32 * (size_t)p->tile_max_blocks[r] + 8 * (tile_r(r) + 1) + 128;
      if (bytes <= budget / 3 && tile_r(r) * vis_per_node <= 1.25 * THREADS) best = r;
   }
   const char *env = getenv("FEMB200_TILE_R");
   if (env)
      for (int r = 0; r < kNumTileR; ++r)
         if (atoi(env) == tile_r(r)) best = r;
   const int R = tile_r(best);
   A.stage_units = (2 * p->tile_max_blocks[best] + 7) & ~7;
   const size_t smem = 16 * (size_t)A.stage_units + 8 * (size_t)(R + 1) + 16;
   FEMB_CHECK(smem <= budget, "assemble: a %d-node tile needs %zu B of shared memory (> %zu)", R, smem, budget);
   FEMB_CUDA(cudaFuncSetAttribute(assemble_kernel<ET, FAST, THREADS>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)smem));
   const unsigned grid = (unsigned)cdiv(p->nnodes, R);
   assemble_kernel<ET, FAST, THREADS><<<grid, THREADS, smem, st>>>(A, R);
   FEMB_LAUNCH_CHECK();
   return 0;
}

}  // namespace femb

using namespace femb;

extern "C" int femb200_assemble_matrix(const femb200_plan *p, const double *d_x, int x_stride, const double *d_E,
                                       double nu, const double *d_dnod, const double *d_u, int variant,
                                       double *d_values, void *stream)
{
   FEMB_CHECK(p && d_x && d_E && d_values, "assemble_matrix: null argument");
   FEMB_CHECK(x_stride == 2 || x_stride == 3, "assemble_matrix: x_stride must be 2 or 3, got %d", x_stride);
   AsmArgs A;
   A.nnodes = p->nnodes, A.nptr = p->nptr, A.vrec = p->vrec, A.brp = p->brp;
   A.xdofmap = p->xdofmap, A.dofmap = p->dofmap, A.x = d_x, A.xs = x_stride, A.E = d_E, A.lc = lame_coef(nu);
   A.dnod = d_dnod, A.u = d_u, A.variant = variant, A.values = d_values;
   cudaStream_t st = as_stream(stream);
   const bool linear = (d_dnod == nullptr) && !getenv("FEMB200_FORCE_GENERIC");
   int rc;
   switch (p->etype)
   {
      case FEMB200_P1:
         rc = linear ? launch_assemble<FEMB200_P1, true>(p, A, st) : launch_assemble<FEMB200_P1, false>(p, A, st);
         break;
      case FEMB200_P2:
         rc = linear ? launch_assemble<FEMB200_P2, true>(p, A, st) : launch_assemble<FEMB200_P2, false>(p, A, st);
         break;
      default:
         rc = launch_assemble<FEMB200_Q2, false>(p, A, st);
   }
   if (rc) return rc;
   if (p->bc && p->nbc > 0) return femb200_apply_dirichlet(p, d_values, 1.0, stream);
   return 0;
}

extern "C" int femb200_apply_dirichlet(const femb200_plan *p, double *d_values, double diag, void *stream)
{
   FEMB_CHECK(p && d_values, "apply_dirichlet: null argument");
   if (!p->bc || p->nbc == 0) return 0;
   const int T = 128;
   dirichlet_kernel<<<(unsigned)cdiv((int64_t)p->nbc * 32, T), T, 0, as_stream(stream)>>>(
       p->nbc, p->bc_nodes, p->bc, p->brp, p->bcol, d_values, diag);
   FEMB_LAUNCH_CHECK();
   return 0;
}

extern "C" int femb200_matrix_norms(const femb200_plan *p, const double *d_values, double *d_out, void *stream)
{
   FEMB_CHECK(p && d_values && d_out, "matrix_norms: null argument");
   cudaStream_t st = as_stream(stream);
   const int T = 256;
   const unsigned grid = (unsigned)std::min<int64_t>(cdiv(p->nnodes * 32, T), (int64_t)devinfo().sm_count * 8);
   ReduceScratch red;
   if (int rc = reduce_scratch(grid, st, &red, 2)) return rc;
   norms_kernel<<<grid, T, 0, st>>>(p->nnodes, p->brp, p->bcol, d_values, red, d_out);
   FEMB_LAUNCH_CHECK();
   return 0;
}
