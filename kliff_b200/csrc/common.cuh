// Shared device/host helpers of libkliff_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "kliff_b200.h"

#define KB_PI 3.1415926535897932  // the reference's literal (sym_fn.cpp:9)
#define KB_TOL 1.0e-10            // neighbor_list.cpp:14

#define KB_CUDA(call)                                                          \
  do {                                                                         \
    cudaError_t e_ = (call);                                                   \
    if (e_ != cudaSuccess) return -(int) e_;                                   \
  } while (0)

// process-wide count of kernels launched by this library (bench.py reports it as gpu_launches)
void kb_note_launch(int n = 1);

#define KB_LAUNCH_CHECK()                                                      \
  do {                                                                         \
    cudaError_t e_ = cudaGetLastError();                                       \
    if (e_ != cudaSuccess) return -(int) e_;                                   \
    kb_note_launch();                                                          \
  } while (0)

// Optional per-kernel timing (kb_kernel_timing): when switched on, launch sites bracket their kernel
// with CUDA events on the launching stream under a small integer tag.  Off by default (no events).
enum { KB_T_PREP = 0, KB_T_SWEEP = 1, KB_T_FUSED = 2, KB_T_SCATTER = 3, KB_T_STAGE = 4, KB_T_COUNT = 5 };
bool kb_timing_on();
void kb_timing_begin(int tag, cudaStream_t st);
void kb_timing_end(int tag, cudaStream_t st);

static inline cudaStream_t kb_stream(void* s) { return static_cast<cudaStream_t>(s); }

// ---- programmatic dependent launch (sm_90+) ---------------------------------------------------
// A kernel launched through kb_launch_pdl may be scheduled while its predecessor in the stream is still
// draining: everything it does before kb_grid_dep_wait() — index loads, loads of data no kernel of the
// chain writes, shared-memory set-up — overlaps the predecessor's tail and the launch latency.  The wait
// returns when the predecessor grid has completed and its writes are visible; it is a no-op for a kernel
// launched the ordinary way.  The attribute is set only while the stream is being captured (the fused
// trainer's epoch graph, where every neighbour in the chain is one of these kernels and the data read ahead
// of the wait was written before the graph launch); eager calls keep ordinary stream order.  KB_PDL=0
// switches it off altogether.
#ifdef __CUDACC__
__device__ __forceinline__ void kb_grid_dep_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void kb_grid_dep_launch() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
#endif
bool kb_pdl_on(cudaStream_t st);
template <typename... KArgs, typename... Args>
inline cudaError_t kb_launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                                 Args&&... args)
{
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = kb_pdl_on(st) ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

// Scratch comes from the stream-ordered allocator.  Its default pool gives memory back to the
// driver at every synchronisation (release threshold 0), which turns each *_begin call into
// fresh cudaMalloc-class allocations (measured: 23 ms per 1M-atom neighbor build instead of 2 ms).
// Keep freed scratch cached in the pool instead.  Idempotent and thread-safe.
void kb_pool_setup();

// ---- IEEE-exact, never-contracted double arithmetic --------------------------------------
// The integer outputs (padding selection, neighbor sets) must equal the reference's bit for
// bit, and the reference is built without FMA (x86-64 -O2).  nvcc contracts a*b+c into DFMA by
// default, so every expression that feeds a predicate uses these.
__device__ __forceinline__ double xmul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double xadd(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double xsub(double a, double b) { return __dsub_rn(a, b); }
// a0*b0 + a1*b1 + a2*b2 evaluated left to right (kliff/neighbor/helper.hpp:26-29)
__device__ __forceinline__ double xdot3(double a0, double a1, double a2, double b0, double b1,
                                        double b2)
{
  return xadd(xadd(xmul(a0, b0), xmul(a1, b1)), xmul(a2, b2));
}
// dx*dx + dy*dy + dz*dz (neighbor_list.cpp:197, sym_fn.cpp:447)
__device__ __forceinline__ double xnorm2(double dx, double dy, double dz)
{
  return xadd(xadd(xmul(dx, dx), xmul(dy, dy)), xmul(dz, dz));
}

// ---- device-wide exclusive scan (hand-written; used for pad counts, cell starts, CSR) -----
template <typename Tin, typename Tout>
int kb_exclusive_scan(const Tin* d_in, Tout* d_out, int64_t n, Tout* d_total /*nullable*/,
                      cudaStream_t stream);

// 64-bit shuffle helpers
__device__ __forceinline__ double shfl_xor_d(double v, int m)
{
  return __shfl_xor_sync(0xffffffffu, v, m);
}
__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
  for (int m = 16; m > 0; m >>= 1) v += shfl_xor_d(v, m);
  return v;
}

// Sum 32 per-lane values across the warp in 31 exchange steps (instead of 32 x 5): on return lane l
// holds the warp-wide total of v[l].  v is clobbered.
__device__ __forceinline__ double warp_reduce32(double (&v)[32], int lane)
{
#pragma unroll
  for (int half = 16; half >= 1; half >>= 1)
  {
    const bool upper = (lane & half) != 0;
#pragma unroll
    for (int i = 0; i < half; i++)
    {
      const double keep = upper ? v[i + half] : v[i];
      const double send = upper ? v[i] : v[i + half];
      v[i] = keep + shfl_xor_d(send, half);
    }
  }
  return v[0];
}

// 16-value variant (fewer live registers): on return lane l holds the warp-wide total of v[l & 15].
__device__ __forceinline__ double warp_reduce16(double (&v)[16], int lane)
{
#pragma unroll
  for (int i = 0; i < 16; i++) v[i] += shfl_xor_d(v[i], 16);
#pragma unroll
  for (int half = 8; half >= 1; half >>= 1)
  {
    const bool upper = (lane & half) != 0;
#pragma unroll
    for (int i = 0; i < half; i++)
    {
      const double keep = upper ? v[i + half] : v[i];
      const double send = upper ? v[i] : v[i + half];
      v[i] = keep + shfl_xor_d(send, half);
    }
  }
  return v[0];
}
