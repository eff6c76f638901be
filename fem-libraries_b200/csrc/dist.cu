// dist.cu -- multi-GPU operator apply and PCG: one mesh partition per rank (process), halo of the search
// direction and all-reduce of the CG dot products over NVLink.  Design and roles in the reference: dist.cuh.
//
// Partition model (any element-block partition, the structured strips of femb200/dist.py being one): local
// node numbering with the owned nodes in ONE contiguous range [own_lo, own_hi) and ghost nodes around it;
// the rank integrates every cell that touches an owned node, so owned matrix rows are complete without
// communication; per neighbour one contiguous range of owned nodes to send and one range of
// ghost nodes to receive.
#include <dlfcn.h>
#include <nccl.h>  // types and enums only: the functions are resolved with dlsym (no link-time dependency)

#include <algorithm>
#include <vector>

#include "dist.cuh"
#include "reduce.cuh"

namespace femb {

// ---- NCCL, resolved at run time ------------------------------------------------------------------
// RTLD_NOLOAD first: when the host process already carries an NCCL (PyTorch bundles one) that instance
// is used, so a communicator created by the host library can be handed over.
struct NcclApi
{
   void *lib = nullptr;
   ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
   ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
   ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
   ncclResult_t (*CommCount)(const ncclComm_t, int *) = nullptr;
   ncclResult_t (*CommUserRank)(const ncclComm_t, int *) = nullptr;
   ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
   ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
   ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
   ncclResult_t (*GroupStart)() = nullptr;
   ncclResult_t (*GroupEnd)() = nullptr;
   const char *(*GetErrorString)(ncclResult_t) = nullptr;
};

static const NcclApi *nccl_api()
{
   static NcclApi api;
   static bool tried = false;
   if (!tried)
   {
      tried = true;
      void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);
      if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
      if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
      if (h)
      {
         api.lib = h;
#define FEMB_NCCL_SYM(name) *reinterpret_cast<void **>(&api.name) = dlsym(h, "nccl" #name)
         FEMB_NCCL_SYM(GetUniqueId);
         FEMB_NCCL_SYM(CommInitRank);
         FEMB_NCCL_SYM(CommDestroy);
         FEMB_NCCL_SYM(CommCount);
         FEMB_NCCL_SYM(CommUserRank);
         FEMB_NCCL_SYM(AllReduce);
         FEMB_NCCL_SYM(Send);
         FEMB_NCCL_SYM(Recv);
         FEMB_NCCL_SYM(GroupStart);
         FEMB_NCCL_SYM(GroupEnd);
         FEMB_NCCL_SYM(GetErrorString);
#undef FEMB_NCCL_SYM
         if (!api.GetUniqueId || !api.CommInitRank || !api.AllReduce || !api.Send || !api.Recv || !api.GroupStart ||
             !api.GroupEnd || !api.CommCount || !api.CommUserRank)
            api.lib = nullptr;
      }
   }
   return api.lib ? &api : nullptr;
}

#define FEMB_NCCL(call)                                                                                    \
   do                                                                                                      \
   {                                                                                                       \
      ncclResult_t r__ = (call);                                                                           \
      if (r__ != ncclSuccess)                                                                              \
         return femb::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call,                              \
                                nccl_api()->GetErrorString ? nccl_api()->GetErrorString(r__) : "NCCL error"); \
   } while (0)

}  // namespace femb

enum
{
   FEMB_TRANSPORT_NONE = 0,
   FEMB_TRANSPORT_NCCL = FEMB200_DIST_NCCL,
   FEMB_TRANSPORT_P2P = FEMB200_DIST_P2P
};

struct femb200_dist
{
   const femb200_plan *plan = nullptr;
   int rank = 0, world = 1;
   int64_t own_lo = 0, own_hi = 0, nnodes = 0;
   int nneigh = 0;
   int peer[femb::kMaxNeigh] = {};
   int64_t send_lo[femb::kMaxNeigh] = {}, send_cnt[femb::kMaxNeigh] = {};
   int64_t recv_lo[femb::kMaxNeigh] = {}, recv_cnt[femb::kMaxNeigh] = {};
   // arena: header, then the CG work vectors (local length, ghosts included)
   unsigned char *arena = nullptr;
   size_t arena_bytes = 0, vec_bytes = 0;
   femb::ArenaHdr *hdr = nullptr;
   double *d = nullptr, *r = nullptr, *z = nullptr, *scal = nullptr;
   int transport = FEMB_TRANSPORT_NONE;
   // NCCL
   ncclComm_t comm = nullptr;
   // P2P
   bool p2p_ready = false;
   unsigned char *peer_arena[femb::kMaxWorld] = {};   // IPC mappings (self: arena)
   int64_t peer_recv_lo[femb::kMaxNeigh] = {};        // first node (neighbour's numbering) of my message there
   femb::IterGraph graph;
   cudaStream_t capture_stream = nullptr;  // iteration graphs are captured here and launched on the caller's stream
};

namespace femb {

// exported with the IPC handle: where the neighbours' messages land in this rank's numbering
struct P2pBlob
{
   cudaIpcMemHandle_t handle;  // 64 bytes
   int32_t rank, nneigh;
   struct
   {
      int32_t peer, pad;
      int64_t recv_lo, recv_cnt;
   } nb[kMaxNeigh - 1];
};
static_assert(sizeof(P2pBlob) <= FEMB200_DIST_BLOB_BYTES, "P2P blob");

// ---- halo over peer memory -----------------------------------------------------------------------
struct HaloArgs
{
   int nneigh;
   const double2 *src;              // local vector (node units)
   int64_t send_lo[kMaxNeigh], send_cnt[kMaxNeigh];
   double2 *dst[kMaxNeigh];         // neighbour's ghost rows of ITS arena vector (peer memory)
   unsigned long long *peer_flag[kMaxNeigh];  // neighbour's halo_flag[my rank]
   int peer_rank[kMaxNeigh];
   ArenaHdr *hdr;
};

// Stores the interface rows into the neighbours' ghost rows, then the block that finishes last publishes
// the sequence number to every neighbour and waits for theirs: when the kernel ends, this rank's ghost rows
// hold the neighbours' current values.  Both sides push before they wait: no circular wait.
__global__ void __launch_bounds__(256) halo_exchange_kernel(HaloArgs H)
{
   __shared__ bool last;
   const unsigned long long seq = H.hdr->seq_halo + 1;
   int64_t total = 0;
   for (int k = 0; k < H.nneigh; ++k) total += H.send_cnt[k];
   for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x)
   {
      int64_t j = i;
      int k = 0;
      while (j >= H.send_cnt[k]) j -= H.send_cnt[k], ++k;
      H.dst[k][j] = H.src[H.send_lo[k] + j];
   }
   __threadfence_system();
   __syncthreads();
   if (threadIdx.x == 0)
   {
      const unsigned int t = atomicAdd(&H.hdr->halo_ticket, 1u);
      last = (t == gridDim.x - 1);
   }
   __syncthreads();
   if (!last) return;
   __threadfence_system();
   if (threadIdx.x < H.nneigh)
   {
      st_release_sys_u64(H.peer_flag[threadIdx.x], seq);
      const unsigned long long *in = &H.hdr->halo_flag[H.peer_rank[threadIdx.x]];
      const unsigned long long t0 = global_timer_ns();
      while (ld_acquire_sys_u64(in) < seq)
         if (global_timer_ns() - t0 > kWaitTimeoutNs)
         {
            H.hdr->error = 2;
            break;
         }
   }
   __syncthreads();
   if (threadIdx.x == 0)
   {
      H.hdr->seq_halo = seq;
      H.hdr->halo_ticket = 0u;
   }
}

// dst[lo + j] = src[lo + j] over node ranges (ghost rows of the arena vector -> the caller's vector, or the
// caller's owned rows -> the arena vector)
struct RangeCopyArgs
{
   int n;
   int64_t lo[kMaxNeigh + 1], cnt[kMaxNeigh + 1];
   const double2 *src;
   double2 *dst;
};
__global__ void range_copy_kernel(RangeCopyArgs A)
{
   int64_t total = 0;
   for (int k = 0; k < A.n; ++k) total += A.cnt[k];
   for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x)
   {
      int64_t j = i;
      int k = 0;
      while (j >= A.cnt[k]) j -= A.cnt[k], ++k;
      A.dst[A.lo[k] + j] = A.src[A.lo[k] + j];
   }
}

// stand-alone all-reduce of up to kRedVals doubles (norms, checksums): one warp
__global__ void __launch_bounds__(32) allreduce_kernel(double *__restrict__ v, int count, RedArgs ra)
{
   double m[kRedVals];
#pragma unroll
   for (int k = 0; k < kRedVals; ++k) m[k] = k < count ? v[k] : 0.;
   __syncwarp();
   mailbox_allreduce<kRedVals>(ra, m);
   if (threadIdx.x == 0)
      for (int k = 0; k < count; ++k) v[k] = m[k];
}

RedArgs dist_red_args(femb200_dist *D)
{
   RedArgs ra;
   if (!D || D->world == 1 || D->transport != FEMB_TRANSPORT_P2P) return ra;
   ra.world = D->world, ra.rank = D->rank, ra.hdr = D->hdr;
   for (int p = 0; p < D->world; ++p) ra.peer[p] = reinterpret_cast<ArenaHdr *>(D->peer_arena[p]);
   return ra;
}

int dist_allreduce_pre(femb200_dist *D, double *d_val, int count, cudaStream_t st)
{
   if (!D || D->world == 1 || D->transport != FEMB_TRANSPORT_NCCL) return 0;
   FEMB_NCCL(nccl_api()->AllReduce(d_val, d_val, (size_t)count, ncclDouble, ncclSum, D->comm, st));
   return 0;
}

// ghost update of a vector: `v` must be the arena's search direction on the P2P transport (the neighbours
// store into it); any device vector on the NCCL transport
static int dist_halo_vec(femb200_dist *D, double *v, cudaStream_t st)
{
   if (!D || D->world == 1 || D->nneigh == 0) return 0;
   if (D->transport == FEMB_TRANSPORT_NCCL)
   {
      const NcclApi *N = nccl_api();
      FEMB_NCCL(N->GroupStart());
      for (int k = 0; k < D->nneigh; ++k)
      {
         FEMB_NCCL(N->Send(v + 2 * D->send_lo[k], (size_t)(2 * D->send_cnt[k]), ncclDouble, D->peer[k], D->comm, st));
         FEMB_NCCL(N->Recv(v + 2 * D->recv_lo[k], (size_t)(2 * D->recv_cnt[k]), ncclDouble, D->peer[k], D->comm, st));
      }
      FEMB_NCCL(N->GroupEnd());
      return 0;
   }
   FEMB_CHECK(D->transport == FEMB_TRANSPORT_P2P && D->p2p_ready, "dist: no transport attached (world = %d)", D->world);
   FEMB_CHECK(v == D->d, "dist: the P2P halo works on the communicator's own search-direction vector");
   HaloArgs H;
   H.nneigh = D->nneigh, H.src = reinterpret_cast<const double2 *>(D->d), H.hdr = D->hdr;
   int64_t total = 0;
   for (int k = 0; k < D->nneigh; ++k)
   {
      const int p = D->peer[k];
      H.send_lo[k] = D->send_lo[k], H.send_cnt[k] = D->send_cnt[k], H.peer_rank[k] = p;
      H.dst[k] = reinterpret_cast<double2 *>(D->peer_arena[p] + kArenaHdrBytes) + D->peer_recv_lo[k];
      H.peer_flag[k] = &reinterpret_cast<ArenaHdr *>(D->peer_arena[p])->halo_flag[D->rank];
      total += D->send_cnt[k];
   }
   const unsigned grid = (unsigned)std::max<int64_t>(1, std::min<int64_t>(cdiv(total, 256 * 4), 32));
   halo_exchange_kernel<<<grid, 256, 0, st>>>(H);
   FEMB_LAUNCH_CHECK();
   return 0;
}

int dist_halo_arena(femb200_dist *D, cudaStream_t st) { return D ? dist_halo_vec(D, D->d, st) : 0; }

int dist_check_error(femb200_dist *D, cudaStream_t st)
{
   if (!D || D->world == 1 || D->transport != FEMB_TRANSPORT_P2P) return 0;
   int e = 0;
   FEMB_CUDA(cudaMemcpyAsync(&e, &D->hdr->error, sizeof(int), cudaMemcpyDeviceToHost, st));
   FEMB_CUDA(cudaStreamSynchronize(st));
   FEMB_CHECK(e == 0, "dist: a peer-memory wait timed out (%s): a rank died or the ranks issued different call sequences",
              e == 1 ? "all-reduce" : "halo");
   return 0;
}

IterGraph *dist_iter_graph(femb200_dist *D) { return D ? &D->graph : nullptr; }

cudaStream_t dist_capture_stream(femb200_dist *D)
{
   if (!D) return nullptr;
   if (!D->capture_stream && cudaStreamCreateWithFlags(&D->capture_stream, cudaStreamNonBlocking) != cudaSuccess) return nullptr;
   return D->capture_stream;
}

}  // namespace femb

using namespace femb;

extern "C" void femb200_dist_destroy(femb200_dist *D)
{
   if (!D) return;
   if (D->graph.exec) cudaGraphExecDestroy(D->graph.exec);
   if (D->capture_stream) cudaStreamDestroy(D->capture_stream);
   for (int p = 0; p < D->world && p < kMaxWorld; ++p)
      if (D->peer_arena[p] && D->peer_arena[p] != D->arena) cudaIpcCloseMemHandle(D->peer_arena[p]);
   cudaFree(D->arena);
   cudaFree(D->scal);
   delete D;
}

extern "C" int femb200_dist_create(const femb200_plan *plan, int rank, int world, int64_t own_lo, int64_t own_hi,
                                   int nneigh, const int32_t *peers, const int64_t *send_lo, const int64_t *send_hi,
                                   const int64_t *recv_lo, const int64_t *recv_hi, void *stream, femb200_dist **out)
{
   FEMB_CHECK(out != nullptr, "dist_create: out is null");
   *out = nullptr;
   FEMB_CHECK(plan != nullptr, "dist_create: null plan");
   FEMB_CHECK(world >= 1 && world <= kMaxWorld && rank >= 0 && rank < world, "dist_create: rank %d of %d (at most %d ranks)",
              rank, world, kMaxWorld);
   FEMB_CHECK(0 <= own_lo && own_lo <= own_hi && own_hi <= plan->nnodes, "dist_create: bad owned range [%lld, %lld)",
              (long long)own_lo, (long long)own_hi);
   FEMB_CHECK(nneigh >= 0 && nneigh < kMaxNeigh, "dist_create: at most %d neighbours", kMaxNeigh - 1);
   FEMB_CHECK(nneigh == 0 || (peers && send_lo && send_hi && recv_lo && recv_hi), "dist_create: null neighbour table");
   femb200_dist *D = new femb200_dist();
   D->plan = plan, D->rank = rank, D->world = world, D->own_lo = own_lo, D->own_hi = own_hi, D->nnodes = plan->nnodes;
   D->nneigh = nneigh;
   for (int k = 0; k < nneigh; ++k)
   {
      const bool ok = peers[k] >= 0 && peers[k] < world && peers[k] != rank && own_lo <= send_lo[k] && send_lo[k] <= send_hi[k] &&
                      send_hi[k] <= own_hi && 0 <= recv_lo[k] && recv_lo[k] <= recv_hi[k] && recv_hi[k] <= plan->nnodes &&
                      (recv_hi[k] <= own_lo || recv_lo[k] >= own_hi);
      if (!ok)
      {
         delete D;
         return set_error("dist_create: neighbour %d: sends must be owned rows, receives ghost rows", k);
      }
      D->peer[k] = peers[k];
      D->send_lo[k] = send_lo[k], D->send_cnt[k] = send_hi[k] - send_lo[k];
      D->recv_lo[k] = recv_lo[k], D->recv_cnt[k] = recv_hi[k] - recv_lo[k];
   }
   D->vec_bytes = ((size_t)(2 * plan->nnodes) * sizeof(double) + 255) & ~(size_t)255;
   D->arena_bytes = kArenaHdrBytes + 3 * D->vec_bytes;
   cudaStream_t st = as_stream(stream);
   if (cudaMalloc(&D->arena, D->arena_bytes) != cudaSuccess || cudaMalloc(&D->scal, sizeof(double) * 16) != cudaSuccess)
   {
      femb200_dist_destroy(D);
      return set_error("dist_create: cudaMalloc of %zu bytes failed", D->arena_bytes);
   }
   // ghost rows that no neighbour fills must not hold NaN patterns: zero everything once
   cudaMemsetAsync(D->arena, 0, D->arena_bytes, st);
   cudaMemsetAsync(D->scal, 0, sizeof(double) * 16, st);
   if (cudaStreamSynchronize(st) != cudaSuccess)
   {
      femb200_dist_destroy(D);
      return set_error("dist_create: %s", cudaGetErrorString(cudaGetLastError()));
   }
   D->hdr = reinterpret_cast<ArenaHdr *>(D->arena);
   D->d = reinterpret_cast<double *>(D->arena + kArenaHdrBytes);
   D->r = reinterpret_cast<double *>(D->arena + kArenaHdrBytes + D->vec_bytes);
   D->z = reinterpret_cast<double *>(D->arena + kArenaHdrBytes + 2 * D->vec_bytes);
   D->peer_arena[rank] = D->arena;
   *out = D;
   return 0;
}

// ---- NCCL transport ------------------------------------------------------------------------------
extern "C" int femb200_dist_nccl_unique_id(unsigned char *id128)
{
   FEMB_CHECK(id128 != nullptr, "dist_nccl_unique_id: null argument");
   const NcclApi *N = nccl_api();
   FEMB_CHECK(N != nullptr, "dist: libnccl.so.2 not found");
   ncclUniqueId id;
   FEMB_NCCL(N->GetUniqueId(&id));
   static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId");
   memcpy(id128, &id, 128);
   return 0;
}

extern "C" int femb200_dist_nccl_comm_create(const unsigned char *id128, int rank, int world, void **comm_out)
{
   FEMB_CHECK(id128 && comm_out, "dist_nccl_comm_create: null argument");
   const NcclApi *N = nccl_api();
   FEMB_CHECK(N != nullptr, "dist: libnccl.so.2 not found");
   ncclUniqueId id;
   memcpy(&id, id128, 128);
   ncclComm_t c = nullptr;
   FEMB_NCCL(N->CommInitRank(&c, world, id, rank));
   *comm_out = c;
   return 0;
}

extern "C" int femb200_dist_nccl_comm_destroy(void *comm)
{
   const NcclApi *N = nccl_api();
   if (N && N->CommDestroy && comm) N->CommDestroy(static_cast<ncclComm_t>(comm));
   return 0;
}

extern "C" int femb200_dist_attach_nccl(femb200_dist *D, void *nccl_comm)
{
   FEMB_CHECK(D && nccl_comm, "dist_attach_nccl: null argument");
   const NcclApi *N = nccl_api();
   FEMB_CHECK(N != nullptr, "dist: libnccl.so.2 not found");
   int n = 0, r = -1;
   FEMB_NCCL(N->CommCount(static_cast<ncclComm_t>(nccl_comm), &n));
   FEMB_NCCL(N->CommUserRank(static_cast<ncclComm_t>(nccl_comm), &r));
   FEMB_CHECK(n == D->world && r == D->rank, "dist_attach_nccl: communicator is rank %d of %d, the partition rank %d of %d", r,
              n, D->rank, D->world);
   D->comm = static_cast<ncclComm_t>(nccl_comm);
   if (D->transport == FEMB_TRANSPORT_NONE) D->transport = FEMB_TRANSPORT_NCCL;
   return 0;
}

// ---- P2P transport -------------------------------------------------------------------------------
extern "C" int femb200_dist_p2p_export(femb200_dist *D, unsigned char *blob)
{
   FEMB_CHECK(D && blob, "dist_p2p_export: null argument");
   P2pBlob b;
   memset(&b, 0, sizeof(b));
   FEMB_CUDA(cudaIpcGetMemHandle(&b.handle, D->arena));
   b.rank = D->rank, b.nneigh = D->nneigh;
   for (int k = 0; k < D->nneigh; ++k) b.nb[k].peer = D->peer[k], b.nb[k].recv_lo = D->recv_lo[k], b.nb[k].recv_cnt = D->recv_cnt[k];
   memset(blob, 0, FEMB200_DIST_BLOB_BYTES);
   memcpy(blob, &b, sizeof(b));
   return 0;
}

extern "C" int femb200_dist_p2p_attach(femb200_dist *D, const unsigned char *blobs)
{
   FEMB_CHECK(D && blobs, "dist_p2p_attach: null argument");
   int dev = 0;
   FEMB_CUDA(cudaGetDevice(&dev));
   for (int p = 0; p < D->world; ++p)
   {
      if (p == D->rank) continue;
      P2pBlob b;
      memcpy(&b, blobs + (size_t)p * FEMB200_DIST_BLOB_BYTES, sizeof(b));
      FEMB_CHECK(b.rank == p, "dist_p2p_attach: blob %d carries rank %d", p, b.rank);
      void *ptr = nullptr;
      const cudaError_t e = cudaIpcOpenMemHandle(&ptr, b.handle, cudaIpcMemLazyEnablePeerAccess);
      if (e != cudaSuccess)
      {
         cudaGetLastError();
         return set_error("dist_p2p_attach: cudaIpcOpenMemHandle(rank %d) failed: %s", p, cudaGetErrorString(e));
      }
      D->peer_arena[p] = static_cast<unsigned char *>(ptr);
   }
   for (int k = 0; k < D->nneigh; ++k)
   {  // where does neighbour k expect my message?
      P2pBlob b;
      memcpy(&b, blobs + (size_t)D->peer[k] * FEMB200_DIST_BLOB_BYTES, sizeof(b));
      bool found = false;
      for (int j = 0; j < b.nneigh; ++j)
         if (b.nb[j].peer == D->rank)
         {
            FEMB_CHECK(b.nb[j].recv_cnt == D->send_cnt[k], "dist_p2p_attach: rank %d sends %lld nodes to rank %d, which expects %lld",
                       D->rank, (long long)D->send_cnt[k], D->peer[k], (long long)b.nb[j].recv_cnt);
            D->peer_recv_lo[k] = b.nb[j].recv_lo;
            found = true;
         }
      FEMB_CHECK(found, "dist_p2p_attach: rank %d does not list rank %d as a neighbour", D->peer[k], D->rank);
   }
   D->p2p_ready = true;
   D->transport = FEMB_TRANSPORT_P2P;
   return 0;
}

extern "C" int femb200_dist_set_transport(femb200_dist *D, int transport)
{
   FEMB_CHECK(D != nullptr, "dist_set_transport: null argument");
   if (transport == FEMB200_DIST_NCCL)
      FEMB_CHECK(D->comm != nullptr, "dist_set_transport: no NCCL communicator attached");
   else if (transport == FEMB200_DIST_P2P)
      FEMB_CHECK(D->p2p_ready, "dist_set_transport: peer memory not attached");
   else
      return set_error("dist_set_transport: unknown transport %d", transport);
   if (D->transport != transport && D->graph.exec) cudaGraphExecDestroy(D->graph.exec), D->graph.exec = nullptr;
   D->transport = transport;
   return 0;
}

extern "C" int femb200_dist_transport(const femb200_dist *D) { return D ? D->transport : 0; }

// ---- collectives on caller vectors -----------------------------------------------------------------
extern "C" int femb200_dist_allreduce_sum(femb200_dist *D, double *d_vals, int count, void *stream)
{
   FEMB_CHECK(D && d_vals && count >= 1 && count <= kRedVals, "dist_allreduce_sum: 1..%d doubles", kRedVals);
   if (D->world == 1) return 0;
   cudaStream_t st = as_stream(stream);
   if (D->transport == FEMB_TRANSPORT_NCCL) return dist_allreduce_pre(D, d_vals, count, st);
   FEMB_CHECK(D->transport == FEMB_TRANSPORT_P2P, "dist: no transport attached");
   allreduce_kernel<<<1, 32, 0, st>>>(d_vals, count, dist_red_args(D));
   FEMB_LAUNCH_CHECK();
   return 0;
}

// forward ghost update of a caller vector (VecGhostUpdate(INSERT, FORWARD), F.cc:865-866)
extern "C" int femb200_dist_halo(femb200_dist *D, double *d_v, void *stream)
{
   FEMB_CHECK(D && d_v, "dist_halo: null argument");
   if (D->world == 1 || D->nneigh == 0) return 0;
   cudaStream_t st = as_stream(stream);
   if (D->transport == FEMB_TRANSPORT_NCCL) return dist_halo_vec(D, d_v, st);
   // P2P: the interface rows travel through the arena vector: caller's send rows -> arena, exchange, arena
   // ghost rows -> caller
   RangeCopyArgs A;
   A.n = D->nneigh;
   int64_t total = 0;
   for (int k = 0; k < D->nneigh; ++k) A.lo[k] = D->send_lo[k], A.cnt[k] = D->send_cnt[k], total += D->send_cnt[k];
   A.src = reinterpret_cast<const double2 *>(d_v), A.dst = reinterpret_cast<double2 *>(D->d);
   const unsigned grid = (unsigned)std::max<int64_t>(1, std::min<int64_t>(cdiv(total, 256), 64));
   range_copy_kernel<<<grid, 256, 0, st>>>(A);
   FEMB_LAUNCH_CHECK();
   if (int rc = dist_halo_vec(D, D->d, st)) return rc;
   total = 0;
   for (int k = 0; k < D->nneigh; ++k) A.lo[k] = D->recv_lo[k], A.cnt[k] = D->recv_cnt[k], total += D->recv_cnt[k];
   A.src = reinterpret_cast<const double2 *>(D->d), A.dst = reinterpret_cast<double2 *>(d_v);
   range_copy_kernel<<<(unsigned)std::max<int64_t>(1, std::min<int64_t>(cdiv(total, 256), 64)), 256, 0, st>>>(A);
   FEMB_LAUNCH_CHECK();
   // Barrier: a neighbour that runs ahead must not store the NEXT message into the arena's ghost rows while this
   // rank is still copying the current one out (inside the PCG the all-reduce of <d, A d> plays this role).
   allreduce_kernel<<<1, 32, 0, st>>>(D->scal + 13, 1, dist_red_args(D));
   FEMB_LAUNCH_CHECK();
   return 0;
}

// y[owned] = (A v)[owned] after the ghost update of v (the operator apply of one CG iteration, stand-alone)
extern "C" int femb200_dist_mult(femb200_dist *D, const double *d_values, double *d_v, double *d_y, void *stream)
{
   FEMB_CHECK(D && d_values && d_v && d_y, "dist_mult: null argument");
   if (int rc = femb200_dist_halo(D, d_v, stream)) return rc;
   RowRange rr;
   if (int rc = plan_row_range(D->plan, D->own_lo, D->own_hi, &rr)) return rc;
   return spmv_launch(D->plan, rr, d_values, d_v, d_y, nullptr, nullptr, false, as_stream(stream));
}

// (Jacobi-)PCG over the ranks, mfem::CGSolver semantics (femb200_pcg).  d_b, d_x, d_dinv are local vectors
// (2 * plan nodes, ghosts included; only the owned entries are read / written).  Host scalars out are
// identical on every rank.  Synchronises the stream.
extern "C" int femb200_dist_pcg(femb200_dist *D, int op_kind, const void *op, const double *d_values, const double *d_b,
                                double *d_x, double rtol, double atol, int maxit, const double *d_dinv, int check_every,
                                int fixed_iters, int use_graph, int *iters, double *final_norm, int *converged, void *stream)
{
   FEMB_CHECK(D && d_b && d_x, "dist_pcg: null argument");
   FEMB_CHECK(maxit >= 0, "dist_pcg: negative maxit");
   FEMB_CHECK(D->world == 1 || D->transport != FEMB_TRANSPORT_NONE, "dist_pcg: no transport attached (world = %d)", D->world);
   CgProblem P;
   P.plan = D->plan, P.op_kind = op_kind, P.op = op, P.values = d_values;
   P.own_lo = D->own_lo, P.own_hi = D->own_hi;
   P.b = d_b, P.dinv = d_dinv, P.x = d_x;
   P.r = D->r, P.d = D->d, P.z = D->z, P.scal = D->scal;
   P.rtol = rtol, P.atol = atol, P.maxit = maxit, P.check_every = check_every, P.fixed_iters = fixed_iters;
   P.comm = D, P.use_graph = use_graph != 0;
   return cg_core(P, iters, final_norm, converged, as_stream(stream));
}

// device pointers of the communicator's work vectors after a solve: r (recurrence residual), d, z
extern "C" int femb200_dist_vectors(femb200_dist *D, double **d_r, double **d_dir, double **d_z, double **d_scal)
{
   FEMB_CHECK(D != nullptr, "dist_vectors: null argument");
   if (d_r) *d_r = D->r;
   if (d_dir) *d_dir = D->d;
   if (d_z) *d_z = D->z;
   if (d_scal) *d_scal = D->scal;
   return 0;
}
