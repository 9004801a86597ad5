// Stage 2 of the two-stage symmetry-function path, round-2 form: the "queue sweep", run by warp
// PAIRS (one CTA = a helper warp + a sweep warp working on the same stream of atoms).
//
// Why a second sweep (ncu on the v3 sweep, profiles/r01_ncu_symfun_sweep_v3.txt): lane = (slot,
// eta-group) computes 8 (zeta, lambda) entries per lane whether the group has them or not (43 used of
// 64 for set51), re-derives C / C' for every entry in every lane, and re-reads a 22-double scratch
// entry per lane and step, so the FP64 pipe (53 %) and the shared-memory data pipe (70 %) saturate
// together with a third of the arithmetic wasted.  Here the register tile is turned the other way:
//
//   lane = (position p of 16, lambda class s)        16 slots in flight per warp, two lanes each
//   a lane owns EVERY (zeta, eta) item of its lambda: NI = sum_z (NE - LO_z) items (22 for set51),
//   66 gradient + 22 value accumulators in registers; no lane carries an absent item (43 of 44).
//
// Per step (one neighbor j of the lane's slot k) the lane reads one 20-double entry that both lanes of
// the position share {E_eta[NE], cos, F, V1[3], U[3], W[3]} and spends 2 DFMA per item-component:
//   T1 = A' V1 + A U,  T2 = A W,  T0 = A F         (per zeta: 10 instructions, shared by its etas)
//   acc[item][c] += E_eta (T1[c] + eta T2[c]),   phi[item] += E_eta T0
// i.e. ~205 FP64 instructions per 22 items and 10 shared loads, against 80 per 5.4 items and 11 loads.
// The exponentials move to the geometry phase (lanes = entries, one evaluation per (pair, eta)).
//
// Work distribution inside the warp: slots are handed to the 16 positions in descending order of
// their adjacency degree, a position taking the next slot the moment its current one is finished
// (longest-processing-time list scheduling: diamond Si, degrees 4 x 18 / 12 x 12 / 12 x 10, needs 22
// steps instead of the 28 of a two-pass split).  Entries are produced in windows of QW steps per
// position; a slot that ends inside a window is flushed at that step (its 22 x 3 row pieces) and the
// position continues with its next slot in the same window.
//
// Why warp pairs (ncu on the single-warp form of this sweep, profiles/r02_ncu_qsweep_single.txt): with
// 254 registers only two warps fit per scheduler, and a warp that does everything spends three
// quarters of its time in the latency-bound parts (stream bookkeeping + geometry, flush, centre-slot
// read-back) while the FP64 pipe idles (36 %).  So the roles are split:
//   helper warp   record copy, row pre-zeroing, stream bookkeeping + geometry (entries of one window
//                 into a two-deep ring in shared memory), centre-slot read-back of finished atoms
//   sweep warp    the steps and the flushes, zeta at the end of an atom
// The two run one window (and across atoms: up to two windows) apart, synchronised by three
// release/acquire counters in shared memory.  Each scheduler hosts one warp of each role.
//
// The centre slot (translation invariance: minus the sum over slots) is formed by reading the
// finished rows back through L2 — keeping 66 more running sums in registers is not affordable; the
// rows are zeroed with full-sector writes first, so that the 24-byte row pieces land on L2-resident
// lines (no sector fills from DRAM, and the read-back hits L2).
//
// Reference: Descriptor::generate_one_atom (sym_fn.cpp:416-602), sym_d_g4 (:713-809).
#pragma once
#include "symfun_atom.cuh"

namespace {

constexpr int QW = 4;                    // steps per window and position
constexpr int QE = 20;                   // doubles per entry
constexpr int QSTRIDE = QW * QE + 2;     // doubles per position: (QSTRIDE / 2) odd -> the 16 positions start on
                                         // different 16-byte bank groups (two per group)
constexpr int QSC_DOUBLES = 16 * QSTRIDE;
constexpr int QCS = 33;                  // row stride of the centre-slot sums (odd: rows start on different banks)

// Compile-time shape of a parameter set: NE distinct eta values (groups), and for each of the four
// exponents zeta = {1,2,4,16}[z] the first group that carries it; groups [LO_z, NE) are computed for
// both lambda classes, absent (group, entry) combinations are computed and dropped (grp_out < 0).
template <int NE_, int LA, int LB, int LC, int LD>
struct QShape
{
  static constexpr int NE = NE_;
  __host__ __device__ static constexpr int lo(int z) { return z == 0 ? LA : z == 1 ? LB : z == 2 ? LC : LD; }
  __host__ __device__ static constexpr int off(int z)
  {
    return z == 0 ? 0 : z == 1 ? NE - LA : z == 2 ? 2 * NE - LA - LB : 3 * NE - LA - LB - LC;
  }
  static constexpr int NI = 4 * NE - LA - LB - LC - LD;
};

struct QInfo  // one per ring slot: which atom the window belongs to
{
  long long dzoff;  // start of the atom's block in dz (doubles)
  int t, n, flags, pad;
};
enum { QF_LAST = 1, QF_EXIT = 2 };  // flags; bits 8.. = number of steps of the window that carry work

struct QSmem
{
  const unsigned char *KORD, *DEG, *JL;
  const unsigned short* JR;
  const double *XYZ, *TJ, *RK;
  double *SC, *TAB;   // SC: [2][QSC_DOUBLES] entry ring
  int* GOUT;          // [2][NI] descriptor row of (lambda class, item), -1 = absent
  unsigned* META;     // [2][32] per ring slot and lane: QW x (slot | 0x80 on the slot's last step)
  QInfo* INFO;        // [2]
  int* SYNC;          // [0] windows produced, [1] windows consumed, [2] atoms finished, [3] role of warp 0
  double* CS;         // [3 NI][QCS] per sweep lane: sum over its flushed slots (for the centre slot)
};

__host__ __device__ inline size_t qsweep_extra_bytes(int NI)
{
  return (size_t) 2 * QSC_DOUBLES * 8 + 64 * 8 + (size_t) 3 * NI * QCS * 8
         + (((size_t) 2 * NI * 4 + 15) & ~(size_t) 15) + 2 * 32 * 4 + 2 * sizeof(QInfo) + 16;
}

__device__ __forceinline__ QSmem qsweep_carve(unsigned char* base, const RecLayout& L, int NI)
{
  QSmem W;
  W.KORD = base + L.kord;
  W.DEG = base + L.deg;
  W.XYZ = reinterpret_cast<const double*>(base + L.xyz);
  W.TJ = reinterpret_cast<const double*>(base + L.tj);
  W.JL = base + L.jl;
  W.JR = reinterpret_cast<const unsigned short*>(base + L.jr);
  W.RK = reinterpret_cast<const double*>(base + L.rk);
  W.SC = reinterpret_cast<double*>(base + L.bytes);
  W.TAB = W.SC + 2 * QSC_DOUBLES;
  W.CS = W.TAB + 64;
  W.GOUT = reinterpret_cast<int*>(W.CS + 3 * NI * QCS);
  unsigned char* q = reinterpret_cast<unsigned char*>(W.GOUT) + ((2 * NI * 4 + 15) & ~15);
  W.META = reinterpret_cast<unsigned*>(q);
  W.INFO = reinterpret_cast<QInfo*>(q + 2 * 32 * 4);
  W.SYNC = reinterpret_cast<int*>(q + 2 * 32 * 4 + 2 * sizeof(QInfo));
  return W;
}

// ---- release / acquire counters in shared memory ----------------------------------------------
__device__ __forceinline__ void q_signal(int* p, int v, int lane)
{
  __syncwarp();
  if (lane == 0)
    asm volatile("st.release.cta.shared.s32 [%0], %1;" ::"r"((unsigned) __cvta_generic_to_shared(p)), "r"(v) : "memory");
}
__device__ __forceinline__ void q_wait(const int* p, int target)
{
  const unsigned a = (unsigned) __cvta_generic_to_shared(p);
  int v;
  for (;;)
  {
    asm volatile("ld.acquire.cta.shared.s32 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
    if (v >= target) break;
    __nanosleep(64);  // the partner warp shares this scheduler's issue slots and the shared-memory pipe
  }
  __syncwarp();
}

// once per kernel (both warps, 64 threads): exponential table, finite scratch, the output-row table
template <class SH>
__device__ __forceinline__ void qsweep_tables(const kb_symfun_params& P, const QSmem& W, unsigned char* base,
                                              int rec_bytes, int tid)
{
  // the record area must hold finite numbers before the first atom arrives: idle positions read
  // row 0 of its tables and multiply by zero
  for (int q = tid; q < rec_bytes / 8; q += 64) reinterpret_cast<double*>(base)[q] = 0.0;
  for (int j = tid; j < 64; j += 64) W.TAB[j] = exp2((double) j * (1.0 / 64.0));
  for (int q = tid; q < 2 * QSC_DOUBLES; q += 64) W.SC[q] = 0.0;
  for (int q = tid; q < 3 * SH::NI * QCS; q += 64) W.CS[q] = 0.0;
  for (int q = tid; q < 2 * SH::NI; q += 64)
  {
    const int s = q / SH::NI, item = q - s * SH::NI;
    int z = 3;
    while (z > 0 && item < SH::off(z)) z--;
    const int g = SH::lo(z) + item - SH::off(z);
    W.GOUT[q] = g < P.n_groups ? P.grp_out[g][2 * z + s] : -1;
  }
  if (tid < 4) W.SYNC[tid] = 0;
}

// ---- helper side -------------------------------------------------------------------------------
struct QStream
{
  int cursor;            // next slot (in degree order) to hand out; warp-uniform
  int ck, cdeg, cidx;    // this position's slot, its step count, steps generated so far
  bool creal;            // false: the single empty step of a slot without triples
  int nsteps;            // steps of the last produced window that carry work (trailing idle steps dropped)
};

// Zero the angular rows of one atom with full, coalesced sector writes.
template <class SH>
__device__ __forceinline__ void q_prezero(const SplitArgs& A, const QSmem& W, double* blk, int n, int lane)
{
  const int row = 3 * (n + 1);
  if (A.ang_hi > A.ang_lo)
  {
    double* z0 = blk + (unsigned) (A.ang_lo * row);
    const int len = (A.ang_hi - A.ang_lo) * row;
    for (int q = lane; q < len; q += 32) z0[q] = 0.0;
  }
  else
    for (int i = 0; i < 2 * SH::NI; i++)
    {
      const int d = W.GOUT[i];
      if (d >= 0)
        for (int q = lane; q < row; q += 32) blk[(unsigned) (d * row) + q] = 0.0;
    }
}

// One window: QW (slot, step) pairs per position, their entries into `scb`, the sweep lanes' step
// descriptors into `metab`.  Returns true when the atom has no work left after this window.
template <class SH>
__device__ __forceinline__ bool q_produce(const kb_symfun_params& P, const SplitArgs& A, const QSmem& W,
                                          QStream& S, int n, double* scb, unsigned* metab,
                                          double (&phiH)[2][SH::NI], int lane)
{
  constexpr int NE = SH::NE;
  const unsigned FULL = 0xffffffffu;
  const int NP = A.NP;
  const unsigned char *KORD = W.KORD, *DEG = W.DEG, *JL = W.JL;
  const unsigned short* JR = W.JR;
  const double *XYZ = W.XYZ, *TJ = W.TJ, *RK = W.RK, *TAB = W.TAB;
  const int s = lane & 1, p = lane >> 1;
  const unsigned below = (1u << (lane & ~1)) - 1u;  // lanes of the positions before this one
  double* scp = scb + p * QSTRIDE;

  // ---- stream bookkeeping: which (slot, step) each of the QW entries of this position is --------
  unsigned meta = 0;
  int gk[QW / 2], gi[QW / 2];
  bool gv[QW / 2];
#pragma unroll
  for (int it = 0; it < QW; it++)
  {
    const bool need = S.cidx >= S.cdeg;
    const unsigned bal = __ballot_sync(FULL, need && s == 0);
    if (need)
    {
      const int q = S.cursor + __popc(bal & below);
      if (q < n)
      {
        S.ck = KORD[q];
        const int d = DEG[S.ck];
        S.creal = d > 0;
        S.cdeg = S.creal ? d : 1;
      }
      else { S.ck = -1; S.cdeg = 0; S.creal = false; }
      S.cidx = 0;
    }
    S.cursor += __popc(bal);
    const bool busy = S.cidx < S.cdeg;
    if (busy) meta |= (unsigned) (S.ck | 0x40 | (S.cidx == S.cdeg - 1 ? 0x80 : 0)) << (8 * it);  // slot < 64
    if ((it & 1) == s) { gk[it >> 1] = S.ck; gi[it >> 1] = S.cidx; gv[it >> 1] = busy && S.creal; }
    if (busy) S.cidx++;
  }
  metab[lane] = meta;
  S.nsteps = 0;
#pragma unroll
  for (int it = 0; it < QW; it++)
    if (__any_sync(FULL, (meta >> (8 * it)) & 0x40u)) S.nsteps = it + 1;

  // ---- geometry: this lane's QW/2 entries (steps it = s, s + 2, ...), staged side by side so that the
  //      dependent shared-memory round trips of the entries overlap (no branches: an empty entry reads
  //      row 0 of every table and is switched off through fij = 0)
  constexpr int EN = QW / 2;
  int k[EN], j[EN], rcd[EN];
#pragma unroll
  for (int e = 0; e < EN; e++)
  {
    k[e] = gv[e] ? gk[e] : 0;
    const int li = k[e] * NP + (gv[e] ? gi[e] : 0);
    const int jl = JL[li], jr = JR[li];
    j[e] = gv[e] ? jl : 0;
    rcd[e] = gv[e] ? jr : 0;
  }
  double rik[EN], rik2[EN], irik[EN], fik[EN], dfik[EN], ukx[EN], uky[EN], ukz[EN], rij2[EN], irij[EN], fij[EN];
  double rjk[EN], irjk[EN], fjk[EN], dfjk[EN], dx[EN], dy[EN], dz_[EN];
#pragma unroll
  for (int e = 0; e < EN; e++)
  {
    const double2* tk = reinterpret_cast<const double2*>(TJ + k[e] * SJ);
    const double2* tj = reinterpret_cast<const double2*>(TJ + j[e] * SJ);
    const double2* rk = reinterpret_cast<const double2*>(RK + rcd[e] * SR);
    const double2 k0 = tk[0], k1 = tk[1], k2 = tk[2], k3 = tk[3], j0 = tj[0], j1 = tj[1], r0 = rk[0], r1 = rk[1];
    rik[e] = k0.x; rik2[e] = k0.y; irik[e] = k1.x; fik[e] = k1.y; dfik[e] = k2.x; ukx[e] = k2.y; uky[e] = k3.x; ukz[e] = k3.y;
    rij2[e] = j0.y; irij[e] = j1.x; fij[e] = gv[e] ? j1.y : 0.0;
    rjk[e] = r0.x; irjk[e] = r0.y; fjk[e] = r1.x; dfjk[e] = r1.y;
    dx[e] = XYZ[k[e]] - XYZ[j[e]]; dy[e] = XYZ[NP + k[e]] - XYZ[NP + j[e]]; dz_[e] = XYZ[2 * NP + k[e]] - XYZ[2 * NP + j[e]];
  }
#pragma unroll
  for (int e = 0; e < EN; e++)
  {
    double* sc = scp + (2 * e + s) * QE;
    const double wx = dx[e] * irjk[e], wy = dy[e] * irjk[e], wz = dz_[e] * irjk[e];  // unit vector j -> k
    const double rjk2 = rjk[e] * rjk[e];
    const double h = 0.5 * irij[e] * irik[e];
    const double ct = gv[e] ? (rij2[e] + rik2[e] - rjk2) * h : 0.0;   // cos(theta_jik), :744
    const double fijk = fij[e] * fjk[e];
    const double F = fijk * fik[e];
    const double dct_ik = (rik2[e] - rij2[e] + rjk2) * h * irik[e] * F;  // :746, times F
    const double dct_jk = -2.0 * rjk[e] * h * F;                         // :747, times F
    const double dF_ik = dfik[e] * fijk, dF_jk = dfjk[e] * (fij[e] * fik[e]);
    const double a = -2.0 * F * rik[e], b = -2.0 * F * rjk[e];
    const double R2 = rij2[e] + rik2[e] + rjk2;
    double E[8];
#pragma unroll
    for (int g = 0; g < NE; g++) E[g] = kb_exp_tab(-P.grp_eta[g] * R2, TAB);  // :770
#pragma unroll
    for (int g = NE; g < 8; g++) E[g] = 0.0;
    double2* o = reinterpret_cast<double2*>(sc);
    o[0] = make_double2(E[0], E[1]); o[1] = make_double2(E[2], E[3]);
    o[2] = make_double2(E[4], E[5]); o[3] = make_double2(E[6], E[7]);
    o[4] = make_double2(ct, F);
    o[5] = make_double2(dct_ik * ukx[e] + dct_jk * wx, dct_ik * uky[e] + dct_jk * wy);
    o[6] = make_double2(dct_ik * ukz[e] + dct_jk * wz, dF_ik * ukx[e] + dF_jk * wx);
    o[7] = make_double2(dF_ik * uky[e] + dF_jk * wy, dF_ik * ukz[e] + dF_jk * wz);
    o[8] = make_double2(a * ukx[e] + b * wx, a * uky[e] + b * wy);
    o[9] = make_double2(a * ukz[e] + b * wz, 0.0);
    // values (sym_fn.cpp:769-771): every (zeta, lambda, eta) item from this entry; lanes = entries, so
    // this costs the sweep warp nothing (an empty entry has F = 0)
#pragma unroll
    for (int sg = 0; sg < 2; sg++)
    {
      double bs = fma(sg ? 1.0 : -1.0, ct, 1.0);
      bs = bs > 0.0 ? bs : 0.0;                    // guard of :755-767
      const double b2 = bs * bs, b4 = b2 * b2, b8 = b4 * b4, b16 = b8 * b8;
      const double AF[4] = {bs * F, b2 * F, b4 * F, b16 * F};
#pragma unroll
      for (int z = 0; z < 4; z++)
#pragma unroll
        for (int g = SH::lo(z); g < NE; g++)
        {
          const int i = SH::off(z) + g - SH::lo(z);
          phiH[sg][i] = fma(E[g], AF[z], phiH[sg][i]);
        }
    }
  }
  return !__any_sync(FULL, S.cursor < n || S.cidx < S.cdeg);
}

// The helper warp: produces the windows of every atom it grabs and finishes the atoms the sweep warp
// has completed.
template <class SH>
__device__ __forceinline__ void q_helper(const kb_symfun_params& P, const SplitArgs& A, const RecLayout& L,
                                         const QSmem& W, unsigned char* smem_raw, int lane)
{
  const unsigned FULL = 0xffffffffu;
  const unsigned sbase = (unsigned) __cvta_generic_to_shared(smem_raw);
  const int D = P.n_desc;
  int* s_prod = W.SYNC;
  const int* s_cons = W.SYNC + 1;

  auto copy_record = [&](int tt, int nrec) {
    const unsigned char* rec = A.records + (size_t) (tt - A.lo) * L.bytes;
    const int words = (L.rk + 32 * nrec) >> 4;
    for (int q = lane; q < words; q += 32)
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sbase + 16u * q), "l"(rec + 16 * (size_t) q) : "memory");
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  auto grab = [&]() {
    int v = 0;
    if (lane == 0) v = A.lo + atomicAdd(A.work, 1);
    return __shfl_sync(FULL, v, 0);
  };
  auto header = [&](int tt) {
    return tt < A.hi ? __ldcg(reinterpret_cast<const int4*>(A.records + (size_t) (tt - A.lo) * L.bytes))
                     : make_int4(-1, 0, 0, 0);
  };

  int produced = 0;                       // windows handed over
  int cur = grab();
  int4 hcur = header(cur);
  int nxt = grab();
  int4 hnxt = header(nxt);
  while (cur < A.hi)
  {
    const int nx2 = grab();
    const int4 hnx2 = header(nx2);
    if (nx2 < A.hi)
    {
      const unsigned char* nx = A.records + (size_t) (nx2 - A.lo) * L.bytes;
      const int lines = (L.rk + 16 * A.RC + 127) >> 7;
      for (int q = lane; q < lines; q += 32)
        asm volatile("prefetch.global.L2 [%0];" ::"l"(nx + ((size_t) q << 7)));
    }
    const int t = A.subset ? A.subset[cur] : cur;
    const int n = hcur.x;
    if (n >= 0)  // else: declined by stage 1
    {
      __syncwarp();
      copy_record(cur, hcur.y);
      const long long dzoff = A.dz_ptr[t];
      double* blk = static_cast<double*>(A.dz) + dzoff;
      q_prezero<SH>(A, W, blk, n, lane);
      {
        const int len = (int) (A.dz_ptr[t + 1] - dzoff);
        for (int q = D * 3 * (n + 1) + lane; q < len; q += 32) blk[q] = 0.0;  // alignment padding
      }
      asm volatile("cp.async.wait_group 0;" ::: "memory");
      __syncwarp();
      if (n == 0)
      {
        // no neighbors: angular values are zero, the rows (centre slot only) were zeroed above
        double* zrow = A.zeta + (int64_t) t * D;
        for (int i = lane; i < 2 * SH::NI; i += 32)
          if (W.GOUT[i] >= 0) zrow[W.GOUT[i]] = 0.0;
      }
      else
      {
        QStream S;
        S.cursor = 0; S.ck = -1; S.cdeg = 0; S.cidx = 0; S.creal = false; S.nsteps = 0;
        double phiH[2][SH::NI];
#pragma unroll
        for (int i = 0; i < SH::NI; i++) phiH[0][i] = phiH[1][i] = 0.0;
        for (;;)
        {
          q_wait(s_cons, produced - 1);     // ring slot (produced & 1) is free again
          const int b = produced & 1;
          const bool last = q_produce<SH>(P, A, W, S, n, W.SC + b * QSC_DOUBLES, W.META + b * 32, phiH, lane);
          if (lane == 0)
          {
            QInfo I;
            I.dzoff = dzoff; I.t = t; I.n = n; I.flags = (last ? QF_LAST : 0) | (S.nsteps << 8); I.pad = 0;
            W.INFO[b] = I;
          }
          produced++;
          q_signal(s_prod, produced, lane);
          if (last) break;
        }
        // zeta: the 2 NI values summed over the lanes (each unordered triple was visited twice)
        {
          double* zrow = A.zeta + (int64_t) t * D;
#pragma unroll
          for (int sg = 0; sg < 2; sg++)
          {
            static_assert(SH::NI <= 32, "one transposing reduction per lambda class");
            double v[32];
#pragma unroll
            for (int i = 0; i < 32; i++) v[i] = i < SH::NI ? phiH[sg][i] : 0.0;
            const double tot = warp_reduce32(v, lane);   // lane l: total of value l
            if (lane < SH::NI)
            {
              const int z = lane >= SH::off(3) ? 3 : lane >= SH::off(2) ? 2 : lane >= SH::off(1) ? 1 : 0;
              const double p2 = z == 0 ? 0.5 : z == 1 ? 0.25 : z == 2 ? 0.0625 : 1.0 / 65536.0;
              const int d = W.GOUT[sg * SH::NI + lane];
              if (d >= 0) zrow[d] = p2 * tot;
            }
          }
        }
      }
    }
    cur = nxt; hcur = hnxt;
    nxt = nx2; hnxt = hnx2;
  }
  // closing window, then the atoms still waiting for their centre slots
  q_wait(s_cons, produced - 1);
  if (lane == 0) W.INFO[produced & 1].flags = QF_EXIT;
  produced++;
  q_signal(s_prod, produced, lane);
}

// ---- sweep side ----------------------------------------------------------------------------------
template <class SH>
__device__ __forceinline__ void q_sweeper(const kb_symfun_params& P, const SplitArgs& A, const QSmem& W, int lane)
{
  constexpr int NE = SH::NE, NI = SH::NI;
  const unsigned FULL = 0xffffffffu;
  const int D = P.n_desc;
  const int s = lane & 1, p = lane >> 1;
  const double lam = s ? 1.0 : -1.0;
  const int* gout = W.GOUT + s * NI;
  const int* s_prod = W.SYNC;
  int* s_cons = W.SYNC + 1;

  double* cs = W.CS + lane;   // this lane's column of the centre-slot sums

  double acc[NI][3];
#pragma unroll
  for (int i = 0; i < NI; i++) acc[i][0] = acc[i][1] = acc[i][2] = 0.0;

  int consumed = 0;
  for (;;)
  {
    q_wait(s_prod, consumed + 1);
    const int b = consumed & 1;
    const QInfo I = W.INFO[b];
    if (I.flags & QF_EXIT) break;
    const unsigned meta = W.META[b * 32 + lane];
    const int row = 3 * (I.n + 1);
    double* blk = static_cast<double*>(A.dz) + I.dzoff;
    const double* scp = W.SC + b * QSC_DOUBLES + p * QSTRIDE;

    // ---- QW steps; a slot that ends at a step leaves for its rows right there ---------------------
    const int nsteps = I.flags >> 8;
#pragma unroll 1
    for (int it = 0; it < nsteps; it++)
    {
      const double2* sc = reinterpret_cast<const double2*>(scp + it * QE);
      double E[8], Ee[8];
      {
        const double2 e0 = sc[0], e1 = sc[1], e2 = sc[2], e3 = sc[3];
        E[0] = e0.x; E[1] = e0.y; E[2] = e1.x; E[3] = e1.y; E[4] = e2.x; E[5] = e2.y; E[6] = e3.x; E[7] = e3.y;
      }
#pragma unroll
      for (int g = 0; g < NE; g++) Ee[g] = E[g] * P.grp_eta[g];
      const double2 cF = sc[4], q5 = sc[5], q6 = sc[6], q7 = sc[7], q8 = sc[8], q9 = sc[9];
      const double v1x = q5.x, v1y = q5.y, v1z = q6.x, ux = q6.y, uy = q7.x, uz = q7.y;
      const double wx = q8.x, wy = q8.y, wz = q9.x;
      double bs = fma(lam, cF.x, 1.0);
      bs = bs > 0.0 ? bs : 0.0;                       // guard of :755-767
      const double b2 = bs * bs, b3 = b2 * bs, b4 = b2 * b2, b8 = b4 * b4;
      const double b16 = b8 * b8, b15 = (b8 * b4) * b3;
      const double Az[4] = {bs, b2, b4, b16};       // (1 + lambda cos)^zeta; 2^(1-zeta) applied at the end
      const double dAz[4] = {bs > 0.0 ? lam : 0.0, 2.0 * lam * bs, 4.0 * lam * b3, 16.0 * lam * b15};
      // Two passes over the accumulators of a zeta, each a run of independent DFMAs (a single warp
      // sustains ~0.41 DFMA per cycle on such runs and about half that on back-to-back dependent
      // pairs: scripts/probes/dfma_issue_probe.cu)
#pragma unroll
      for (int z = 0; z < 4; z++)
      {
        const double Aq = Az[z], dA = dAz[z];
        const double t1x = fma(dA, v1x, Aq * ux), t1y = fma(dA, v1y, Aq * uy), t1z = fma(dA, v1z, Aq * uz);
        const double t2x = Aq * wx, t2y = Aq * wy, t2z = Aq * wz;
#pragma unroll
        for (int g = SH::lo(z); g < NE; g++)
        {
          const int i = SH::off(z) + g - SH::lo(z);
          acc[i][0] = fma(E[g], t1x, acc[i][0]);
          acc[i][1] = fma(E[g], t1y, acc[i][1]);
          acc[i][2] = fma(E[g], t1z, acc[i][2]);
        }
#pragma unroll
        for (int g = SH::lo(z); g < NE; g++)
        {
          const int i = SH::off(z) + g - SH::lo(z);
          acc[i][0] = fma(Ee[g], t2x, acc[i][0]);
          acc[i][1] = fma(Ee[g], t2y, acc[i][1]);
          acc[i][2] = fma(Ee[g], t2z, acc[i][2]);
        }
      }
      const unsigned m = meta >> (8 * it);
      if (__any_sync(FULL, m & 0x80u))
      {
        if (m & 0x80u)
        {
          const int k = (int) (m & 0x3fu);
          double* ok = blk + 3 * k;
          // running sums of the centre slot first: independent load / add / store triples
#pragma unroll
          for (int z = 0; z < 4; z++)
          {
            double c0[NE], c1[NE], c2[NE];
#pragma unroll
            for (int g = SH::lo(z); g < NE; g++)
            {
              const double* c3 = cs + 3 * (SH::off(z) + g - SH::lo(z)) * QCS;
              c0[g] = c3[0]; c1[g] = c3[QCS]; c2[g] = c3[2 * QCS];
            }
#pragma unroll
            for (int g = SH::lo(z); g < NE; g++)
            {
              const int i = SH::off(z) + g - SH::lo(z);
              double* c3 = cs + 3 * i * QCS;
              c3[0] = c0[g] + acc[i][0]; c3[QCS] = c1[g] + acc[i][1]; c3[2 * QCS] = c2[g] + acc[i][2];
            }
          }
          // row pieces: all addresses of a zeta first (distinct registers), then the stores back to back —
          // with the address formed next to each store the compiler recycles one register pair and every
          // store waits for the previous one to leave (ncu: long-scoreboard stalls on the address)
#pragma unroll
          for (int z = 0; z < 4; z++)
          {
            const double p2 = z == 0 ? 1.0 : z == 1 ? 0.5 : z == 2 ? 0.125 : 1.0 / 32768.0;
            double* o[NE];
            int dd[NE];
#pragma unroll
            for (int g = SH::lo(z); g < NE; g++)
            {
              dd[g] = gout[SH::off(z) + g - SH::lo(z)];
              o[g] = ok + (unsigned) ((dd[g] < 0 ? 0 : dd[g]) * row);
              asm volatile("" : "+l"(o[g]));
              __builtin_assume(__isGlobal(o[g]));
            }
#pragma unroll
            for (int g = SH::lo(z); g < NE; g++)
            {
              const int i = SH::off(z) + g - SH::lo(z);
              if (dd[g] >= 0)
              {
                o[g][0] = p2 * acc[i][0]; o[g][1] = p2 * acc[i][1]; o[g][2] = p2 * acc[i][2];
              }
              acc[i][0] = acc[i][1] = acc[i][2] = 0.0;
            }
          }
        }
      }
    }
    consumed++;
    q_signal(s_cons, consumed, lane);

    if (I.flags & QF_LAST)
    {
      // ---- centre slot (translation invariance, sym_fn.cpp:587-593): minus the sum over the slots.
      //      Lane = row (item, component) of the per-lane sums; even columns are lambda = -1 lanes.
      __syncwarp();
      for (int r = lane; r < 3 * NI; r += 32)
      {
        double* c3 = W.CS + r * QCS;
        double s0 = 0.0, s1 = 0.0;
#pragma unroll
        for (int q = 0; q < 32; q += 2) { s0 += c3[q]; s1 += c3[q + 1]; c3[q] = 0.0; c3[q + 1] = 0.0; }
        const int i = r / 3, c = r - 3 * i;
        const int z = i >= SH::off(3) ? 3 : i >= SH::off(2) ? 2 : i >= SH::off(1) ? 1 : 0;
        const double p2 = z == 0 ? 1.0 : z == 1 ? 0.5 : z == 2 ? 0.125 : 1.0 / 32768.0;
        const int d0 = W.GOUT[i], d1 = W.GOUT[NI + i];
        if (d0 >= 0) blk[(unsigned) (d0 * row) + 3 * I.n + c] = -p2 * s0;
        if (d1 >= 0) blk[(unsigned) (d1 * row) + 3 * I.n + c] = -p2 * s1;
      }
      __syncwarp();
    }
  }
}

}  // namespace
