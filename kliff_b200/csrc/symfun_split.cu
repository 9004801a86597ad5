// Two-stage form of the warp-per-atom symmetry-function kernel (symfun_warp.cu): the hot path, v3.
//
// ncu on the fused kernel (profiles/r01_ncu_symfun_warp_v3.txt) shows where its time goes: the FMA
// sweep over neighbor triples keeps the FP64 pipe ~78 % busy, but it is only a third of the run;
// the table phases before it (pair-ij table, radial sets, adjacency rows, pair-jk records, slot
// sort) are chains of dependent shared-memory round trips that issue one instruction every ~7
// cycles, and with 242 registers per thread only 8 warps fit on an SM to hide that.  So the two
// halves want opposite launch shapes, and here they get them:
//
//   symfun_prep_kernel   80 registers, 4.6 KB of shared memory per warp -> 24 resident warps per SM.
//                        Per atom: phases A, D, B, C, S of symfun_warp.cu (same arithmetic, same
//                        order), radial rows of dG/dr written to their final place, and the tables
//                        the sweep needs packed into one "atom record" in global memory.
//   symfun_sweep_kernel  persistent, 9 warps per SM (216 registers), one atom per warp at a time: copies the record
//                        (~10 KB, one coalesced pass) into shared memory and runs only the triple
//                        loop (phases E, F) — the part that is FP64-pipe bound.
//
// Atoms are processed in chunks so that the records of a chunk (stride ~17 KB per atom for n <= 28)
// stay a bounded scratch allocation; the extra HBM traffic is ~2 x 10 KB per atom on top of the 37 KB
// block itself, on kernels that keep DRAM ~29 % busy.  The per-atom code of both stages lives in
// symfun_atom.cuh.
//
// Reference: Descriptor::generate_one_atom (sym_fn.cpp:416-602), sym_d_g2 / sym_d_g4 (:627-809).
#include "symfun_atom.cuh"
#include "symfun_qsweep.cuh"
#include "symfun_sweep3.cuh"

#include <cstdlib>

namespace {

// =============================================================================================
// stage 1: tables
// =============================================================================================
template <bool GRAD, typename TOut, bool SOA = false>
__global__ void __launch_bounds__(PREP_WARPS * 32, KB_PREP_MIN_BLOCKS)
symfun_prep_kernel(const __grid_constant__ kb_symfun_params P, const SplitArgs A, const int warp_bytes)
{
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const RecLayout L = rec_layout(A.NP, A.RC);
  const PrepSmem W = prep_carve(smem_raw + (size_t) warp * warp_bytes, P, A.NP, A.RC);
  prep_tables(W, P, lane);

  const int stride = gridDim.x * PREP_WARPS;
  for (int tt = A.lo + blockIdx.x * PREP_WARPS + warp; tt < A.hi; tt += stride)
  {
    const int t = A.subset ? A.subset[tt] : tt;
    prep_atom<GRAD, TOut, SOA>(P, A, L, W, t, A.records + (size_t) (tt - A.lo) * L.bytes, lane);
  }
}

// =============================================================================================
// stage 2: triple sweep
// =============================================================================================
template <bool GRAD, typename TOut>
__global__ void __maxnreg__(GRAD ? KB_SWEEP_REGS : KB_SWEEP_REGS_NOGRAD)
symfun_sweep_kernel(const __grid_constant__ kb_symfun_params P, const SplitArgs A)
{
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x;
  const unsigned FULL = 0xffffffffu;
  const RecLayout L = rec_layout(A.NP, A.RC);
  const SweepSmem W = sweep_carve(smem_raw, L);
  const SweepLane Q = sweep_lane(P, A.NGP, lane);
  sweep_tables(W, lane);

  // work distribution: one atom per grab, taken one atom AHEAD so that the next record can be pulled
  // into L2 while the current atom is swept
  int tt_next = 0;
  if (lane == 0) tt_next = A.lo + atomicAdd(A.work, 1);
  tt_next = __shfl_sync(FULL, tt_next, 0);
  int4 hdr_next = make_int4(-1, 0, 0, 0);
  if (tt_next < A.hi)
    hdr_next = __ldcg(reinterpret_cast<const int4*>(A.records + (size_t) (tt_next - A.lo) * L.bytes));
  for (;;)
  {
    const int tt = tt_next;
    if (tt >= A.hi) break;
    const int4 hdr = hdr_next;
    if (lane == 0) tt_next = A.lo + atomicAdd(A.work, 1);
    tt_next = __shfl_sync(FULL, tt_next, 0);
    const int t = A.subset ? A.subset[tt] : tt;
    const unsigned char* rec = A.records + (size_t) (tt - A.lo) * L.bytes;
    const int n = hdr.x, nrec = hdr.y;
    if (tt_next < A.hi)
    {
      // next atom: its header now (consumed one atom later, so the load latency is never waited for)
      // and its lists / tables into L2 (the pair records follow at L.rk; their count is not known
      // yet, fetch the typical half of the capacity)
      const unsigned char* nx = A.records + (size_t) (tt_next - A.lo) * L.bytes;
      hdr_next = __ldcg(reinterpret_cast<const int4*>(nx));
      const int lines = (L.rk + 16 * A.RC + 127) >> 7;
      for (int q = lane; q < lines; q += 32)
        asm volatile("prefetch.global.L2 [%0];" ::"l"(nx + ((size_t) q << 7)));
    }
    if (n < 0) continue;  // declined by stage 1
    __syncwarp();
    {
      // record -> shared memory with asynchronous 16-byte copies: all of them in flight at once
      const int words = (L.rk + 32 * nrec) >> 4;
      const unsigned sbase = (unsigned) __cvta_generic_to_shared(smem_raw);
      for (int q = lane; q < words; q += 32)
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sbase + 16u * q), "l"(rec + 16 * (size_t) q) : "memory");
      asm volatile("cp.async.commit_group;" ::: "memory");
      asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    sweep_atom<GRAD, TOut>(P, A, W, Q, t, n, lane);
  }
}


// =============================================================================================
// stage 2, queue form (symfun_qsweep.cuh): fp64 blocks with gradient, parameter shapes SH
// =============================================================================================
template <class SH>
__global__ void __maxnreg__(KB_QSWEEP_REGS)
symfun_qsweep_kernel(const __grid_constant__ kb_symfun_params P, const SplitArgs A)
{
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const RecLayout L = rec_layout(A.NP, A.RC);
  const QSmem W = qsweep_carve(smem_raw, L, SH::NI);
  qsweep_tables<SH>(P, W, smem_raw, L.bytes, threadIdx.x);
  __syncthreads();
  // Roles: the two warps of a CTA normally sit in adjacent hardware warp slots, i.e. on neighbouring
  // schedulers; alternating the assignment with bit 2 of the slot number puts one helper and one sweep
  // warp on every scheduler when four CTAs share an SM.  (Any assignment is correct; this one balances.)
  if (threadIdx.x == 0)
  {
    unsigned slot;
    asm volatile("mov.u32 %0, %%warpid;" : "=r"(slot));
    W.SYNC[3] = (int) (((slot >> 2) ^ slot) & 1u);
  }
  __syncthreads();
  const int role = warp == 0 ? W.SYNC[3] : 1 - W.SYNC[3];
  if (role == 0) q_sweeper<SH>(P, A, W, lane);
  else q_helper<SH>(P, A, L, W, smem_raw, lane);
}

using QShape51 = QShape<7, 0, 0, 3, 3>;  // set51: zeta 1, 2 on all 7 eta, zeta 4, 16 on the last 4
using QShape30 = QShape<6, 0, 0, 3, 3>;  // set30

// 0: the queue sweep does not cover this parameter set; 1: QShape51; 2: QShape30
int qsweep_shape(const kb_symfun_params& P)
{
  const int NE = P.n_groups;
  if (NE != 7 && NE != 6) return 0;
  const int lo_need[4] = {0, 0, 3, 3};
  for (int z = 0; z < 4; z++)
    for (int g = 0; g < lo_need[z]; g++)
      if (P.grp_out[g][2 * z] >= 0 || P.grp_out[g][2 * z + 1] >= 0) return 0;
  return NE == 7 ? 1 : 2;
}


}  // namespace

// Returns KB_OK and *launched = true when the parameter set is one these kernels handle.  Atoms
// with more than 64 neighbors (or more pair records than fit) are appended to d_overflow and left
// to the caller's second pass.
int kb_symfun_split_launch(const kb_symfun_params& P, int n_centers, const int32_t* d_centers,
                           const int32_t* d_subset, const double* d_coords, const int32_t* d_species,
                           const int64_t* d_nl_off, const int32_t* d_nl_idx, int maxn, double* d_zeta,
                           void* d_dz, int dz_dtype, const int64_t* d_dz_ptr, int32_t* d_overflow,
                           int variant, bool* launched, bool* may_overflow, cudaStream_t st)
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
  // variant 0: the queue sweep where it applies (fp64 blocks with gradient, shapes of set51 / set30);
  // variant 1: always the v3 sweep
  const int qshape = (variant == 0 && grad && dz_dtype == KB_F64) ? qsweep_shape(P) : 0;
  // variant 2: the round-2 sweep (symfun_sweep3.cuh: 12 warps per SM) where it applies — gradient blocks,
  // eight lanes per slot; KB_SWEEP3=0 keeps the round-1 sweep (A/B timing)
  bool s3 = variant == 2 && grad && NGP == 8 && S3_CTAS >= 1;
  if (const char* v = getenv("KB_SWEEP3")) { if (v[0] == '0') s3 = false; }
  const size_t limit = 227 * 1024, per_block_reserve = 1024;
  const size_t extra = qshape ? qsweep_extra_bytes(qshape == 1 ? QShape51::NI : QShape30::NI)
                       : s3   ? (size_t) SC_DOUBLES * 8
                              : (size_t) SC_DOUBLES * 8 + 64 * 8;
  // record capacity: worst case when that still leaves the wanted residency, else what fits (never
  // below 45 % of the worst case: a filled sphere has ~44 % of its neighbor pairs within the cutoff)
  const int want = qshape ? KB_QSWEEP_MIN_BLOCKS : grad ? KB_SWEEP_MIN_BLOCKS : KB_SWEEP_MIN_BLOCKS_NOGRAD;
  int RC = worst;
  {
    const size_t budget = s3 ? (limit / S3_CTAS - per_block_reserve - 512) / S3_WARPS
                             : limit / want - per_block_reserve;
    const size_t fixed = rec_layout(NP, 0).bytes + extra;
    if (fixed + 32 * (size_t) RC > budget)
    {
      const int floor_rc = (worst * 9 + 19) / 20;
      long long cap = budget > fixed ? (long long) ((budget - fixed) / 32) : 0;
      RC = cap > floor_rc ? (int) cap : floor_rc;
      if (RC > worst) RC = worst;
    }
  }
  while (RC > 8 && rec_layout(NP, RC).bytes + extra > limit) RC = RC * 3 / 4;
  if (RC < worst) *may_overflow = true;
  const RecLayout L = rec_layout(NP, RC);
  const size_t sweep_smem = L.bytes + extra;
  if (sweep_smem > limit) return KB_OK;
  const size_t prep_warp = prep_smem_bytes(P, NP, RC);
  if (prep_warp * PREP_WARPS > limit) return KB_OK;

  int dev = 0, sms = 148;
  KB_CUDA(cudaGetDevice(&dev));
  KB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));

  // chunk of atoms whose records are alive at the same time
  int chunk = 1 << 18;
  if (const char* c = getenv("KB_SPLIT_CHUNK")) { const int v = atoi(c); if (v >= 1024) chunk = v; }
  {
    // records of a chunk stay below ~6 GB of scratch whatever the record stride (82 KB at n = 64)
    const long long cap = (6LL << 30) / L.bytes;
    if (chunk > cap) chunk = (int) (cap < 8192 ? 8192 : cap);
  }
  if (chunk > n_centers) chunk = n_centers;
  const int n_chunks = (n_centers + chunk - 1) / chunk;
  unsigned char* records = nullptr;
  int32_t* work = nullptr;
  {
    cudaError_t e = cudaMallocAsync((void**) &records, (size_t) chunk * L.bytes, st);
    if (e == cudaSuccess) e = cudaMallocAsync((void**) &work, sizeof(int32_t) * (size_t) n_chunks, st);
    if (e == cudaSuccess) e = cudaMemsetAsync(work, 0, sizeof(int32_t) * (size_t) n_chunks, st);
    if (e != cudaSuccess)
    {
      if (records) cudaFreeAsync(records, st);
      if (work) cudaFreeAsync(work, st);
      return -(int) e;
    }
  }

  SplitArgs A;
  A.centers = d_centers; A.subset = d_subset; A.coords = d_coords; A.species = d_species;
  A.nl_off = d_nl_off; A.nl_idx = d_nl_idx; A.zeta = d_zeta; A.dz = d_dz; A.dz_ptr = d_dz_ptr;
  A.overflow = d_overflow; A.records = records; A.NP = NP; A.RC = RC; A.NGP = NGP;
  {
    // angular rows as one contiguous range with no radial row inside (true for set51 / set30)
    int lo = P.n_desc, hi = 0, cnt = 0;
    for (int g = 0; g < P.n_groups; g++)
      for (int e = 0; e < GW; e++)
        if (P.grp_out[g][e] >= 0)
        {
          lo = P.grp_out[g][e] < lo ? P.grp_out[g][e] : lo;
          hi = P.grp_out[g][e] + 1 > hi ? P.grp_out[g][e] + 1 : hi;
          cnt++;
        }
    const bool contiguous = cnt > 0 && hi - lo == cnt;
    A.ang_lo = contiguous ? lo : 0;
    A.ang_hi = contiguous ? hi : 0;
    // Off by default: it removes the stage's excess DRAM traffic (82.2 -> 60.3 KB per atom at 1M atoms,
    // profiles/r01_ncu_symfun_split_1M_traffic.txt) but costs 2.8 % of sweep time while DRAM is only
    // 29 % busy.  The cost is not the store path (the same zeros sent as cp.async.bulk copies measure
    // identically); the row pieces themselves get slower once they land on resident dirty lines.
    // KB_PREZERO=1 switches it on.
    // The queue sweep always zeroes first (its centre-slot read-back wants the rows resident in L2).
    const char* pz = getenv("KB_PREZERO");
    if (!qshape && !(pz && pz[0] == '1')) A.ang_lo = A.ang_hi = 0;
  }

  const size_t s3_smem = 512 + S3_WARPS * sweep_smem;
  int per_sm = s3 ? (int) (limit / (s3_smem + per_block_reserve)) : (int) (limit / (sweep_smem + per_block_reserve));
  if (per_sm > (s3 ? S3_CTAS : want)) per_sm = s3 ? S3_CTAS : want;
  if (const char* cap = getenv("KB_WARPS_PER_SM"))  // tuning experiments only
  {
    const int c = atoi(cap);
    if (c >= 1 && c < per_sm) per_sm = c;
  }
  if (per_sm < 1) per_sm = 1;

  int rc = KB_OK;
#define KB_SPLIT_LAUNCH(G, T)                                                                          \
  do {                                                                                                 \
    cudaError_t e_ = cudaFuncSetAttribute(symfun_prep_kernel<G, T>,                                    \
                                          cudaFuncAttributeMaxDynamicSharedMemorySize,                 \
                                          (int) (prep_warp * PREP_WARPS));                             \
    if (e_ == cudaSuccess)                                                                             \
      e_ = cudaFuncSetAttribute(symfun_sweep_kernel<G, T>,                                             \
                                cudaFuncAttributeMaxDynamicSharedMemorySize, (int) sweep_smem);        \
    if (e_ != cudaSuccess) { rc = -(int) e_; break; }                                                  \
    kb_timing_begin(KB_T_PREP, st);                                                                    \
    symfun_prep_kernel<G, T><<<pgrid, PREP_WARPS * 32, prep_warp * PREP_WARPS, st>>>(P, A,             \
                                                                                     (int) prep_warp); \
    kb_timing_end(KB_T_PREP, st);                                                                      \
    kb_timing_begin(KB_T_SWEEP, st);                                                                   \
    symfun_sweep_kernel<G, T><<<sgrid, 32, sweep_smem, st>>>(P, A);                                    \
    kb_timing_end(KB_T_SWEEP, st);                                                                     \
  } while (0)
#define KB_S3_LAUNCH(T)                                                                                \
  do {                                                                                                 \
    cudaError_t e_ = cudaFuncSetAttribute(symfun_prep_kernel<true, T, true>,                           \
                                          cudaFuncAttributeMaxDynamicSharedMemorySize,                 \
                                          (int) (prep_warp * PREP_WARPS));                             \
    if (e_ == cudaSuccess)                                                                             \
      e_ = cudaFuncSetAttribute(symfun_sweep3_kernel<T>,                                               \
                                cudaFuncAttributeMaxDynamicSharedMemorySize, (int) s3_smem);           \
    if (e_ != cudaSuccess) { rc = -(int) e_; break; }                                                  \
    kb_timing_begin(KB_T_PREP, st);                                                                    \
    symfun_prep_kernel<true, T, true><<<pgrid, PREP_WARPS * 32, prep_warp * PREP_WARPS, st>>>(         \
        P, A, (int) prep_warp);                                                                        \
    kb_timing_end(KB_T_PREP, st);                                                                      \
    kb_timing_begin(KB_T_SWEEP, st);                                                                   \
    symfun_sweep3_kernel<T><<<sgrid, S3_WARPS * 32, s3_smem, st>>>(P, A, (int) sweep_smem);            \
    kb_timing_end(KB_T_SWEEP, st);                                                                     \
  } while (0)
#define KB_QSPLIT_LAUNCH(SH)                                                                           \
  do {                                                                                                 \
    cudaError_t e_ = cudaFuncSetAttribute(symfun_prep_kernel<true, double>,                            \
                                          cudaFuncAttributeMaxDynamicSharedMemorySize,                 \
                                          (int) (prep_warp * PREP_WARPS));                             \
    if (e_ == cudaSuccess)                                                                             \
      e_ = cudaFuncSetAttribute(symfun_qsweep_kernel<SH>,                                              \
                                cudaFuncAttributeMaxDynamicSharedMemorySize, (int) sweep_smem);        \
    if (e_ != cudaSuccess) { rc = -(int) e_; break; }                                                  \
    kb_timing_begin(KB_T_PREP, st);                                                                    \
    symfun_prep_kernel<true, double><<<pgrid, PREP_WARPS * 32, prep_warp * PREP_WARPS, st>>>(          \
        P, A, (int) prep_warp);                                                                        \
    kb_timing_end(KB_T_PREP, st);                                                                      \
    kb_timing_begin(KB_T_SWEEP, st);                                                                   \
    symfun_qsweep_kernel<SH><<<sgrid, 64, sweep_smem, st>>>(P, A);                                     \
    kb_timing_end(KB_T_SWEEP, st);                                                                     \
  } while (0)
  // (A pipelined form — table kernel of chunk c + 1 on a low-priority side stream under the sweep of chunk
  // c, two sweep CTAs + two table CTAs per SM — was measured and removed: co-resident, the table kernel
  // runs 5x slower (1.9 instead of 0.37 ms per 64k atoms) and the sweep 30 % slower, 37.6 instead of
  // 32.0 ms per 1M atoms; both kernels live on issue slots and the shared-memory pipe.)
  kb_timing_begin(KB_T_STAGE, st);
  for (int c = 0; c < n_chunks && rc == KB_OK; c++)
  {
    A.lo = c * chunk;
    A.hi = A.lo + chunk < n_centers ? A.lo + chunk : n_centers;
    A.work = work + c;
    const int cnt = A.hi - A.lo;
    int pgrid = (cnt + PREP_WARPS - 1) / PREP_WARPS;
    if (pgrid > sms * 16) pgrid = sms * 16;
    int sgrid = sms * per_sm;
    if (s3) { if (sgrid > (cnt + S3_WARPS - 1) / S3_WARPS) sgrid = (cnt + S3_WARPS - 1) / S3_WARPS; }
    else if (sgrid > cnt) sgrid = cnt;
    if (s3 && dz_dtype == KB_F32) KB_S3_LAUNCH(float);
    else if (s3) KB_S3_LAUNCH(double);
    else if (qshape == 1) KB_QSPLIT_LAUNCH(QShape51);
    else if (qshape == 2) KB_QSPLIT_LAUNCH(QShape30);
    else if (!grad) KB_SPLIT_LAUNCH(false, double);
    else if (dz_dtype == KB_F32) KB_SPLIT_LAUNCH(true, float);
    else KB_SPLIT_LAUNCH(true, double);
    if (rc == KB_OK)
    {
      cudaError_t e = cudaGetLastError();
      if (e != cudaSuccess) rc = -(int) e;
      else kb_note_launch(2);
    }
  }
#undef KB_SPLIT_LAUNCH
#undef KB_QSPLIT_LAUNCH
#undef KB_S3_LAUNCH
  kb_timing_end(KB_T_STAGE, st);
  cudaFreeAsync(records, st);
  cudaFreeAsync(work, st);
  if (rc == KB_OK) *launched = true;
  return rc;
}
