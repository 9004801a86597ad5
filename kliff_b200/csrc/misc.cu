// Version / error strings / device info and the two measurement probes used by bench.py.
#include "common.cuh"

#include <atomic>
#include <stdlib.h>

static std::atomic<long long> g_launches{0};
void kb_note_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

bool kb_pdl_on(cudaStream_t st)
{
  static const bool on = [] {
    const char* e = getenv("KB_PDL");
    return !(e && e[0] == '0');
  }();
  if (!on) return false;
  cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
  return cudaStreamIsCapturing(st, &cs) == cudaSuccess && cs == cudaStreamCaptureStatusActive;
}

void kb_pool_setup()
{
  static std::atomic<unsigned long long> done_mask{0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return;
  const unsigned long long bit = 1ull << dev;
  if (done_mask.load(std::memory_order_acquire) & bit) return;
  cudaMemPool_t pool;
  if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess)
  {
    unsigned long long keep = ~0ull;
    cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
  }
  done_mask.fetch_or(bit, std::memory_order_release);
}

// Give cached scratch back: the stream-ordered pool keeps everything it ever held (release threshold
// above), e.g. 4.4 GB of atom records after a 1 M-atom descriptor pass, where torch's caching
// allocator cannot see it.  Synchronises the device first (blocks still in flight are not trimmed).
extern "C" int kb_pool_trim(size_t keep_bytes)
{
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return -(int) e;
  cudaMemPool_t pool;
  e = cudaDeviceGetDefaultMemPool(&pool, dev);
  if (e != cudaSuccess) return -(int) e;
  e = cudaDeviceSynchronize();
  if (e != cudaSuccess) return -(int) e;
  e = cudaMemPoolTrimTo(pool, keep_bytes);
  return e == cudaSuccess ? KB_OK : -(int) e;
}

// ---- optional kernel timing ------------------------------------------------------------------
#include <mutex>
#include <vector>
namespace {
std::atomic<int> g_timing{0};
std::mutex g_timing_mu;
struct TimedLaunch { int tag; cudaEvent_t e0, e1; };
std::vector<TimedLaunch> g_timed;
thread_local cudaEvent_t t_open[KB_T_COUNT];  // begin/end pair up inside one host thread
}  // namespace
bool kb_timing_on() { return g_timing.load(std::memory_order_relaxed) != 0; }
void kb_timing_begin(int tag, cudaStream_t st)
{
  if (!kb_timing_on()) return;
  cudaEventCreate(&t_open[tag]);
  cudaEventRecord(t_open[tag], st);
}
void kb_timing_end(int tag, cudaStream_t st)
{
  if (!kb_timing_on()) return;
  std::lock_guard<std::mutex> lk(g_timing_mu);
  TimedLaunch t{tag, t_open[tag], nullptr};
  cudaEventCreate(&t.e1);
  cudaEventRecord(t.e1, st);
  g_timed.push_back(t);
}

extern "C" int kb_kernel_timing(int32_t enable, double* h_ms, int32_t* h_launches)
{
  // collect what was recorded so far (synchronises on the recorded events), then set the switch
  std::lock_guard<std::mutex> lk(g_timing_mu);
  for (int q = 0; q < KB_T_COUNT; q++)
  {
    if (h_ms) h_ms[q] = 0.0;
    if (h_launches) h_launches[q] = 0;
  }
  int rc = KB_OK;
  for (TimedLaunch& t : g_timed)
  {
    float ms = 0.f;
    cudaError_t e = cudaEventSynchronize(t.e1);
    if (e == cudaSuccess) e = cudaEventElapsedTime(&ms, t.e0, t.e1);
    if (e != cudaSuccess) rc = -(int) e;
    else
    {
      if (h_ms) h_ms[t.tag] += ms;
      if (h_launches) h_launches[t.tag] += 1;
    }
    cudaEventDestroy(t.e0);
    cudaEventDestroy(t.e1);
  }
  g_timed.clear();
  g_timing.store(enable ? 1 : 0, std::memory_order_relaxed);
  return rc;
}

extern "C" int64_t kb_launch_count(int32_t reset)
{
  return reset ? g_launches.exchange(0) : g_launches.load();
}

namespace {

// 8 independent DFMA chains per thread, fully unrolled: measures the FP64 FMA pipe, not memory.
__global__ void __launch_bounds__(256)
dfma_probe_kernel(int iters, double seed, double* __restrict__ sink)
{
  double a0 = seed + threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5,
         a6 = a0 + 6, a7 = a0 + 7;
  const double m = 0.999999, c = 1e-9;
  for (int it = 0; it < iters; it++)
  {
#pragma unroll
    for (int u = 0; u < 16; u++)
    {
      a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
      a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
    }
  }
  const double s = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
  if (s == 12345.6789) sink[0] = s;  // never true; keeps the chains alive
}

__global__ void __launch_bounds__(256)
copy_probe_kernel(const double2* __restrict__ in, double2* __restrict__ out, size_t n)
{
  for (size_t q = (size_t) blockIdx.x * blockDim.x + threadIdx.x; q < n;
       q += (size_t) gridDim.x * blockDim.x)
    out[q] = in[q];
}

}  // namespace

extern "C" const char* kb_version(void) { return "kliff_b200 0.1 (sm_100a)"; }

extern "C" const char* kb_error_string(int code)
{
  switch (code)
  {
    case KB_OK: return "ok";
    case KB_ERR_REFERENCE:
      return "reference error class: singular cell, atom collision, or cell grid too large";
    case KB_ERR_INVALID: return "invalid argument";
    case KB_ERR_LIMIT: return "compiled-in limit exceeded";
    case KB_ERR_UNSUPPORTED: return "device is not sm_100 (B200)";
    default:
      if (code < 0) return cudaGetErrorString((cudaError_t) (-code));
      return "unknown error";
  }
}

extern "C" int kb_device_info(int32_t* sm_count, int32_t* cc_major, int32_t* cc_minor)
{
  int dev = 0;
  KB_CUDA(cudaGetDevice(&dev));
  cudaDeviceProp prop;
  KB_CUDA(cudaGetDeviceProperties(&prop, dev));
  if (sm_count) *sm_count = prop.multiProcessorCount;
  if (cc_major) *cc_major = prop.major;
  if (cc_minor) *cc_minor = prop.minor;
  return prop.major == 10 ? KB_OK : KB_ERR_UNSUPPORTED;
}

extern "C" int kb_probe_fp64_peak(double* h_tflops, void* stream)
{
  if (!h_tflops) return KB_ERR_INVALID;
  cudaStream_t st = kb_stream(stream);
  int32_t sms = 0;
  int rc = kb_device_info(&sms, nullptr, nullptr);
  if (rc) return rc;
  double* sink = nullptr;
  KB_CUDA(cudaMallocAsync((void**) &sink, sizeof(double), st));
  cudaEvent_t e0, e1;
  KB_CUDA(cudaEventCreate(&e0));
  KB_CUDA(cudaEventCreate(&e1));
  const int blocks = sms * 8, iters = 4096;
  double best = 0.0;
  for (int rep = 0; rep < 5; rep++)
  {
    KB_CUDA(cudaEventRecord(e0, st));
    dfma_probe_kernel<<<blocks, 256, 0, st>>>(iters, 1.0 + rep, sink);
    kb_note_launch();
    KB_CUDA(cudaEventRecord(e1, st));
    KB_CUDA(cudaEventSynchronize(e1));
    float ms = 0.f;
    KB_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    const double flops = 2.0 * 8.0 * 16.0 * (double) iters * 256.0 * (double) blocks;
    const double tf = flops / (ms * 1e-3) / 1e12;
    if (rep > 0 && tf > best) best = tf;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  KB_CUDA(cudaFreeAsync(sink, st));
  *h_tflops = best;
  return KB_OK;
}

extern "C" int kb_probe_hbm_copy(size_t bytes, double* h_gbs, void* stream)
{
  if (!h_gbs || bytes < 32) return KB_ERR_INVALID;
  cudaStream_t st = kb_stream(stream);
  const size_t n = bytes / 16;
  double2 *a = nullptr, *b = nullptr;
  KB_CUDA(cudaMallocAsync((void**) &a, n * 16, st));
  KB_CUDA(cudaMallocAsync((void**) &b, n * 16, st));
  KB_CUDA(cudaMemsetAsync(a, 0, n * 16, st));
  cudaEvent_t e0, e1;
  KB_CUDA(cudaEventCreate(&e0));
  KB_CUDA(cudaEventCreate(&e1));
  double best = 0.0;
  for (int rep = 0; rep < 6; rep++)
  {
    KB_CUDA(cudaEventRecord(e0, st));
    copy_probe_kernel<<<148 * 16, 256, 0, st>>>(a, b, n);
    kb_note_launch();
    KB_CUDA(cudaEventRecord(e1, st));
    KB_CUDA(cudaEventSynchronize(e1));
    float ms = 0.f;
    KB_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    const double gbs = 2.0 * (double) (n * 16) / (ms * 1e-3) / 1e9;
    if (rep > 0 && gbs > best) best = gbs;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  KB_CUDA(cudaFreeAsync(a, st));
  KB_CUDA(cudaFreeAsync(b, st));
  *h_gbs = best;
  return KB_OK;
}
