// Warp-specialised single-kernel form of the hot path (mode KB_SYMFUN_WS): table warps and sweep warps
// live in the SAME persistent CTA, so the latency-bound table phases fill the issue slots the FP64-bound
// triple sweep leaves free instead of running as a separate kernel before it (symfun_split.cu spends
// ~19 % of the stage there).
//
//   CTA = 256 threads = two warpgroups, 2 CTAs per SM:
//     warps 0-3  producers: `setmaxnreg.dec` to 40 registers, run prep_atom() (symfun_atom.cuh) for atoms
//                taken from an atomic counter and write each atom record into a two-slot ring in
//                global memory (40 MB for the whole grid: it lives in L2);
//     warps 4-7  consumers: `setmaxnreg.inc` to 216 registers, copy the record of their producer
//                (warp w-4) into shared memory with cp.async and run sweep_atom().
//   Hand-off per producer/consumer pair: a state word per ring slot in shared memory (0 = free, 1 = full);
//   every producer lane fences its record writes before lane 0 publishes the slot, the consumer frees the
//   slot as soon as the record is in shared memory, so the producer runs up to two atoms ahead.
//
// Arithmetic, record format and results are those of the two-stage kernels.
#include "symfun_atom.cuh"

#include <cstdlib>

namespace {

constexpr int WS_PAIRS = 4;
constexpr int WS_SYNC_BYTES = 64;  // per CTA: WS_PAIRS x {state[2], tt[2]}

__device__ __forceinline__ int ld_acquire_smem(const int* p)
{
  int v;
  asm volatile("ld.acquire.cta.shared.s32 %0, [%1];" : "=r"(v) : "r"((unsigned) __cvta_generic_to_shared(p)) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_smem(int* p, int v)
{
  asm volatile("st.release.cta.shared.s32 [%0], %1;" ::"r"((unsigned) __cvta_generic_to_shared(p)), "r"(v) : "memory");
}

template <bool GRAD, typename TOut>
__global__ void __launch_bounds__(WS_PAIRS * 64, 2)
symfun_ws_kernel(const __grid_constant__ kb_symfun_params P, const SplitArgs A, const int sweep_bytes,
                 const int prep_bytes)
{
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int pair = warp & (WS_PAIRS - 1);
  const unsigned FULL = 0xffffffffu;
  const RecLayout L = rec_layout(A.NP, A.RC);
  int* sync = reinterpret_cast<int*>(smem_raw + WS_PAIRS * (sweep_bytes + prep_bytes)) + 4 * pair;
  int* state = sync;      // [2]
  int* token = sync + 2;  // [2] centre ordinal handed over with the slot (-1 = no more atoms)
  if (threadIdx.x < 4 * WS_PAIRS)
    reinterpret_cast<int*>(smem_raw + WS_PAIRS * (sweep_bytes + prep_bytes))[threadIdx.x] = 0;
  __syncthreads();
  unsigned char* ring = A.records + (size_t) (blockIdx.x * WS_PAIRS + pair) * 2 * L.bytes;

  if (warp < WS_PAIRS)
  {
    // ---------------------------------------------------------------- producers (tables)
    asm volatile("setmaxnreg.dec.sync.aligned.u32 40;\n");
    const PrepSmem W = prep_carve(smem_raw + WS_PAIRS * sweep_bytes + pair * prep_bytes, P, A.NP, A.RC);
    prep_tables(W, P, lane);
    int slot = 0;
    for (;;)
    {
      int tt = 0;
      if (lane == 0) tt = A.lo + atomicAdd(A.work, 1);
      tt = __shfl_sync(FULL, tt, 0);
      if (lane == 0)
        while (ld_acquire_smem(state + slot) != 0) __nanosleep(2000);  // a producer is ~4x faster than its consumer
      __syncwarp();
      if (tt >= A.hi)
      {
        if (lane == 0) { token[slot] = -1; st_release_smem(state + slot, 1); }
        break;
      }
      const int t = A.subset ? A.subset[tt] : tt;
      prep_atom<GRAD, TOut>(P, A, L, W, t, ring + (size_t) slot * L.bytes, lane);
      __threadfence();  // this lane's part of the record is visible before the slot is published
      __syncwarp();
      if (lane == 0) { token[slot] = t; st_release_smem(state + slot, 1); }
      slot ^= 1;
    }
  }
  else
  {
    // ---------------------------------------------------------------- consumers (triple sweep)
    asm volatile("setmaxnreg.inc.sync.aligned.u32 216;\n");
    unsigned char* image = smem_raw + pair * sweep_bytes;
    const SweepSmem W = sweep_carve(image, L);
    const SweepLane Q = sweep_lane(P, A.NGP, lane);
    sweep_tables(W, lane);
    int slot = 0;
    for (;;)
    {
      if (lane == 0)
        while (ld_acquire_smem(state + slot) != 1) __nanosleep(200);
      __syncwarp();
      const int t = *reinterpret_cast<volatile int*>(token + slot);
      if (t < 0) break;
      const unsigned char* rec = ring + (size_t) slot * L.bytes;
      const int4 hdr = __ldcg(reinterpret_cast<const int4*>(rec));
      const int n = hdr.x, nrec = hdr.y;
      if (n >= 0)
      {
        const int words = (L.rk + 32 * nrec) >> 4;
        const unsigned sbase = (unsigned) __cvta_generic_to_shared(image);
        for (int q = lane; q < words; q += 32)
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sbase + 16u * q), "l"(rec + 16 * (size_t) q) : "memory");
        asm volatile("cp.async.commit_group;" ::: "memory");
        asm volatile("cp.async.wait_group 0;" ::: "memory");
      }
      __syncwarp();
      if (lane == 0) st_release_smem(state + slot, 0);  // the record is in shared memory: slot free again
      if (n >= 0) sweep_atom<GRAD, TOut>(P, A, W, Q, t, n, lane);
      slot ^= 1;
    }
  }
}

}  // namespace

// Same contract as kb_symfun_split_launch.
int kb_symfun_ws_launch(const kb_symfun_params& P, int n_centers, const int32_t* d_centers,
                        const int32_t* d_subset, const double* d_coords, const int32_t* d_species,
                        const int64_t* d_nl_off, const int32_t* d_nl_idx, int maxn, double* d_zeta,
                        void* d_dz, int dz_dtype, const int64_t* d_dz_ptr, int32_t* d_overflow,
                        bool* launched, bool* may_overflow, cudaStream_t st)
{
  *launched = false;
  *may_overflow = false;
  if (n_centers <= 0) return KB_OK;
  if (!P.canonical_zl || P.n_groups < 1) return KB_OK;
  for (int g = 0; g < P.n_groups; g++)
    if (P.grp_kind[g] != KB_G4) return KB_OK;
  int NGP = 1;
  while (NGP < P.n_groups) NGP <<= 1;
  if (NGP > 32) return KB_OK;
  int NP = maxn < 2 ? 2 : maxn;
  if (NP > 64) { NP = 64; *may_overflow = true; }
  NP = (NP + 1) & ~1;
  const int worst = NP * (NP - 1) / 2;
  const bool grad = d_dz != nullptr;
  const size_t limit = 227 * 1024, per_block_reserve = 1024;
  const size_t extra = (size_t) SC_DOUBLES * 8 + 64 * 8;
  // shared memory of one CTA (two per SM): WS_PAIRS x (record image + scratch + tables of the pair)
  int RC = worst;
  const size_t per_cta = limit / 2 - per_block_reserve - WS_SYNC_BYTES;
  auto cta_bytes = [&](int rc) {
    return WS_PAIRS * (((size_t) rec_layout(NP, rc).bytes + extra + 15) / 16 * 16 + prep_smem_bytes(P, NP, rc));
  };
  while (RC > 8 && cta_bytes(RC) > per_cta) RC = RC * 15 / 16;
  if (cta_bytes(RC) > per_cta || RC < (worst * 9 + 19) / 20) return KB_OK;  // the two-stage kernels take it
  if (RC < worst) *may_overflow = true;
  const RecLayout L = rec_layout(NP, RC);
  const int sweep_bytes = (int) (((size_t) L.bytes + extra + 15) / 16 * 16);
  const int prep_bytes = (int) prep_smem_bytes(P, NP, RC);
  const size_t smem = (size_t) WS_PAIRS * (sweep_bytes + prep_bytes) + WS_SYNC_BYTES;

  int dev = 0, sms = 148;
  KB_CUDA(cudaGetDevice(&dev));
  KB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  int grid = sms * 2;
  if (grid * WS_PAIRS > n_centers) grid = (n_centers + WS_PAIRS - 1) / WS_PAIRS;

  unsigned char* ring = nullptr;
  int32_t* work = nullptr;
  KB_CUDA(cudaMallocAsync((void**) &ring, (size_t) grid * WS_PAIRS * 2 * L.bytes, st));
  KB_CUDA(cudaMallocAsync((void**) &work, sizeof(int32_t), st));
  KB_CUDA(cudaMemsetAsync(work, 0, sizeof(int32_t), st));

  SplitArgs A;
  A.lo = 0; A.hi = n_centers;
  A.centers = d_centers; A.subset = d_subset; A.coords = d_coords; A.species = d_species;
  A.nl_off = d_nl_off; A.nl_idx = d_nl_idx; A.zeta = d_zeta; A.dz = d_dz; A.dz_ptr = d_dz_ptr;
  A.overflow = d_overflow; A.work = work; A.records = ring; A.NP = NP; A.RC = RC; A.NGP = NGP;
  A.ang_lo = A.ang_hi = 0;

  int rc = KB_OK;
#define KB_WS_LAUNCH(G, T)                                                                         \
  do {                                                                                             \
    cudaError_t e_ = cudaFuncSetAttribute(symfun_ws_kernel<G, T>,                                  \
                                          cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem); \
    if (e_ != cudaSuccess) { rc = -(int) e_; break; }                                              \
    kb_timing_begin(KB_T_WS, st);                                                                  \
    symfun_ws_kernel<G, T><<<grid, WS_PAIRS * 64, smem, st>>>(P, A, sweep_bytes, prep_bytes);      \
    kb_timing_end(KB_T_WS, st);                                                                    \
  } while (0)
  if (!grad) KB_WS_LAUNCH(false, double);
  else if (dz_dtype == KB_F32) KB_WS_LAUNCH(true, float);
  else KB_WS_LAUNCH(true, double);
#undef KB_WS_LAUNCH
  if (rc == KB_OK)
  {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) rc = -(int) e;
    else kb_note_launch(1);
  }
  cudaFreeAsync(ring, st);
  cudaFreeAsync(work, st);
  if (rc == KB_OK) *launched = true;
  return rc;
}
