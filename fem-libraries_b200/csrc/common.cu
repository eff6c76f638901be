// common.cu -- error plumbing, device info, device-wide exclusive scan.
#include <stdarg.h>

#include <map>
#include <mutex>

#include "common.cuh"
#include "reduce.cuh"

namespace femb {

thread_local char g_err[512] = "";

int set_error(const char *fmt, ...)
{
   va_list ap;
   va_start(ap, fmt);
   vsnprintf(g_err, sizeof(g_err), fmt, ap);
   va_end(ap);
   return 1;
}

const DevInfo &devinfo()
{
   static DevInfo info;
   static bool init = false;
   if (!init)
   {
      int dev = 0;
      cudaGetDevice(&dev);
      cudaDeviceProp p;
      if (cudaGetDeviceProperties(&p, dev) == cudaSuccess)
      {
         info.sm_count = p.multiProcessorCount;
         info.cc_major = p.major;
         info.cc_minor = p.minor;
         info.smem_optin = p.sharedMemPerBlockOptin;
      }
      init = true;
   }
   return info;
}

// ---------------------------------------------------------------------------
// Exclusive scan: 256 threads x 8 items per block; block sums scanned
// recursively.  Setup-time only (pattern build), not on the hot path.
// ---------------------------------------------------------------------------
constexpr int SCAN_T = 256, SCAN_I = 8, SCAN_B = SCAN_T * SCAN_I;

template <typename TOut>
__device__ TOut block_exclusive_scan(TOut v, TOut *total)
{
   __shared__ TOut warp_tot[SCAN_T / 32];
   const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
   TOut inc = v;
#pragma unroll
   for (int o = 1; o < 32; o <<= 1)
   {
      TOut t = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += t;
   }
   if (lane == 31) warp_tot[w] = inc;
   __syncthreads();
   if (w == 0)
   {
      TOut t = (lane < SCAN_T / 32) ? warp_tot[lane] : TOut(0);
#pragma unroll
      for (int o = 1; o < SCAN_T / 32; o <<= 1)
      {
         TOut u = __shfl_up_sync(0xffffffffu, t, o);
         if (lane >= o) t += u;
      }
      if (lane < SCAN_T / 32) warp_tot[lane] = t;
   }
   __syncthreads();
   const TOut base = (w == 0) ? TOut(0) : warp_tot[w - 1];
   *total = warp_tot[SCAN_T / 32 - 1];
   return base + inc - v;
}

template <typename TIn, typename TOut>
__global__ void scan_block_sums(const TIn *__restrict__ in, TOut *__restrict__ sums, int64_t n)
{
   const int64_t base = (int64_t)blockIdx.x * SCAN_B;
   TOut s = 0;
#pragma unroll
   for (int k = 0; k < SCAN_I; ++k)
   {
      const int64_t i = base + (int64_t)k * SCAN_T + threadIdx.x;
      if (i < n) s += (TOut)in[i];
   }
   TOut tot;
   block_exclusive_scan<TOut>(s, &tot);
   if (threadIdx.x == 0) sums[blockIdx.x] = tot;
}

template <typename TIn, typename TOut>
__global__ void scan_apply(const TIn *__restrict__ in, const TOut *__restrict__ offs, TOut *__restrict__ out, int64_t n)
{
   const int64_t base = (int64_t)blockIdx.x * SCAN_B + (int64_t)threadIdx.x * SCAN_I;
   TOut v[SCAN_I];
   TOut s = 0;
#pragma unroll
   for (int k = 0; k < SCAN_I; ++k)
   {
      const int64_t i = base + k;
      v[k] = (i < n) ? (TOut)in[i] : TOut(0);
      s += v[k];
   }
   TOut tot;
   TOut ex = block_exclusive_scan<TOut>(s, &tot) + (offs ? offs[blockIdx.x] : TOut(0));
#pragma unroll
   for (int k = 0; k < SCAN_I; ++k)
   {
      const int64_t i = base + k;
      if (i < n) out[i] = ex;
      ex += v[k];
      if (i == n - 1) out[n] = ex;
   }
}

template <typename TIn, typename TOut>
static int scan_rec(const TIn *d_in, TOut *d_out, int64_t n, cudaStream_t st)
{
   if (n <= 0)
   {
      FEMB_CUDA(cudaMemsetAsync(d_out, 0, sizeof(TOut), st));
      return 0;
   }
   const int64_t nb = cdiv(n, SCAN_B);
   if (nb == 1)
   {
      scan_apply<TIn, TOut><<<1, SCAN_T, 0, st>>>(d_in, nullptr, d_out, n);
      FEMB_LAUNCH_CHECK();
      return 0;
   }
   TOut *sums = nullptr, *offs = nullptr;
   FEMB_CUDA(cudaMalloc(&sums, sizeof(TOut) * (size_t)nb));
   FEMB_CUDA(cudaMalloc(&offs, sizeof(TOut) * (size_t)(nb + 1)));
   scan_block_sums<TIn, TOut><<<(unsigned)nb, SCAN_T, 0, st>>>(d_in, sums, n);
   FEMB_LAUNCH_CHECK();
   int rc = scan_rec<TOut, TOut>(sums, offs, nb, st);
   if (rc == 0)
   {
      scan_apply<TIn, TOut><<<(unsigned)nb, SCAN_T, 0, st>>>(d_in, offs, d_out, n);
      if (cudaGetLastError() != cudaSuccess) rc = set_error("scan_apply launch failed");
   }
   cudaStreamSynchronize(st);
   cudaFree(sums);
   cudaFree(offs);
   return rc;
}

// ---------------------------------------------------------------------------
// scratch of the fused grid reductions: one buffer per stream of the current
// device, grown on demand (growth synchronises that stream once).
// ---------------------------------------------------------------------------
int reduce_scratch(size_t nblocks, cudaStream_t st, ReduceScratch *out, int nvals)
{
   struct Slot
   {
      double *partials = nullptr;
      unsigned int *ticket = nullptr;
      size_t cap = 0;
   };
   static std::map<std::pair<int, cudaStream_t>, Slot> slots;
   static std::mutex mtx;
   std::lock_guard<std::mutex> lock(mtx);
   int dev = 0;
   FEMB_CUDA(cudaGetDevice(&dev));
   Slot &s = slots[std::make_pair(dev, st)];
   const size_t need = nblocks * (size_t)(nvals > 0 ? nvals : 1);
   if (need > s.cap)
   {
      FEMB_CUDA(cudaStreamSynchronize(st));
      if (s.partials) cudaFree(s.partials);
      const size_t cap = need < 4096 ? 4096 : need + need / 2;
      FEMB_CUDA(cudaMalloc(&s.partials, sizeof(double) * cap));
      s.cap = cap;
   }
   if (!s.ticket)
   {
      FEMB_CUDA(cudaMalloc(&s.ticket, sizeof(unsigned int)));
      FEMB_CUDA(cudaMemsetAsync(s.ticket, 0, sizeof(unsigned int), st));
   }
   out->partials = s.partials;
   out->ticket = s.ticket;
   return 0;
}

int exclusive_scan_i32_i64(const int32_t *d_in, int64_t *d_out, int64_t n, cudaStream_t st)
{
   return scan_rec<int32_t, int64_t>(d_in, d_out, n, st);
}
int exclusive_scan_i32_i32(const int32_t *d_in, int32_t *d_out, int64_t n, cudaStream_t st)
{
   return scan_rec<int32_t, int32_t>(d_in, d_out, n, st);
}

}  // namespace femb

extern "C" int femb200_version(void) { return FEMB200_VERSION; }
extern "C" const char *femb200_last_error(void) { return femb::g_err; }
extern "C" int femb200_device_info(int *sm_count, int *cc_major, int *cc_minor)
{
   const femb::DevInfo &d = femb::devinfo();
   if (sm_count) *sm_count = d.sm_count;
   if (cc_major) *cc_major = d.cc_major;
   if (cc_minor) *cc_minor = d.cc_minor;
   return d.sm_count > 0 ? 0 : femb::set_error("no CUDA device");
}
