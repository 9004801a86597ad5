// Stage 2 of the two-stage symmetry-function path, round-2 form: the "queue sweep".
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
// The centre slot (translation invariance: minus the sum over slots) is formed at the end by reading
// the finished rows back through L2 — keeping 66 more running sums in registers is not affordable.
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

struct QSmem
{
  const unsigned char *KORD, *DEG, *JL;
  const unsigned short* JR;
  const double *XYZ, *TJ, *RK;
  double *SC, *TAB;
  int* GOUT;  // [2][NI] descriptor row of (lambda class, item), -1 = absent
};

__host__ __device__ inline size_t qsweep_extra_bytes(int NI)
{
  return (size_t) QSC_DOUBLES * 8 + 64 * 8 + (((size_t) 2 * NI * 4 + 15) & ~(size_t) 15);
}

__device__ __forceinline__ QSmem qsweep_carve(unsigned char* base, const RecLayout& L)
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
  W.TAB = W.SC + QSC_DOUBLES;
  W.GOUT = reinterpret_cast<int*>(W.TAB + 64);
  return W;
}

// once per kernel: exponential table, finite scratch, the output-row table
template <class SH>
__device__ __forceinline__ void qsweep_tables(const kb_symfun_params& P, const QSmem& W, unsigned char* base,
                                              int rec_bytes, int lane)
{
  // the record area must hold finite numbers before the first atom arrives: idle positions read
  // row 0 of its tables and multiply by zero
  for (int q = lane; q < rec_bytes / 8; q += 32) reinterpret_cast<double*>(base)[q] = 0.0;
  kb_exp_table_fill(W.TAB, lane);
  for (int q = lane; q < QSC_DOUBLES; q += 32) W.SC[q] = 0.0;
  for (int q = lane; q < 2 * SH::NI; q += 32)
  {
    const int s = q / SH::NI, item = q - s * SH::NI;
    int z = 3;
    while (z > 0 && item < SH::off(z)) z--;
    const int g = SH::lo(z) + item - SH::off(z);
    W.GOUT[q] = g < P.n_groups ? P.grp_out[g][2 * z + s] : -1;
  }
  __syncwarp();
}

// =============================================================================================
// stage 2 for one atom whose record (n neighbors) has been copied to shared memory
// =============================================================================================
template <class SH>
__device__ __forceinline__ void qsweep_atom(const kb_symfun_params& P, const SplitArgs& A, const QSmem& W,
                                            int t, int n, int lane)
{
  constexpr int NE = SH::NE, NI = SH::NI;
  const unsigned FULL = 0xffffffffu;
  const int NP = A.NP, D = P.n_desc;
  const unsigned char *KORD = W.KORD, *DEG = W.DEG, *JL = W.JL;
  const unsigned short* JR = W.JR;
  const double *XYZ = W.XYZ, *TJ = W.TJ, *RK = W.RK, *TAB = W.TAB;
  const int s = lane & 1, p = lane >> 1;
  const double lam = s ? 1.0 : -1.0;
  const unsigned below = (1u << (lane & ~1)) - 1u;  // lanes of the positions before this one
  const int row = 3 * (n + 1);
  double* blk = static_cast<double*>(A.dz) + A.dz_ptr[t];
  double* zrow = A.zeta + (int64_t) t * D;
  double* scp = W.SC + p * QSTRIDE;
  const int* gout = W.GOUT + s * NI;

  // Zero the angular rows with full, coalesced sector writes first.  The row pieces written at a flush
  // are 24 bytes at 8-byte alignment, i.e. partial 32-byte sectors: a partial write to a sector that is
  // not in L2 makes L2 fetch it from DRAM (ncu on the round-1 sweep: 13.5 KB of reads per atom that the
  // algorithm never asks for, and the centre-slot read-back below would wait for those fills).  After
  // this pass every piece lands on a resident dirty line.
  if (A.ang_hi > A.ang_lo)
  {
    double* z0 = blk + (unsigned) (A.ang_lo * row);
    const int len = (A.ang_hi - A.ang_lo) * row;
    for (int q = lane; q < len; q += 32) z0[q] = 0.0;
  }
  else
    for (int i = 0; i < 2 * NI; i++)
    {
      const int d = W.GOUT[i];
      if (d >= 0)
        for (int q = lane; q < row; q += 32) blk[(unsigned) (d * row) + q] = 0.0;
    }

  double acc[NI][3], phi[NI];
#pragma unroll
  for (int i = 0; i < NI; i++) { acc[i][0] = acc[i][1] = acc[i][2] = 0.0; phi[i] = 0.0; }

  int cursor = 0;                     // next slot (in degree order) to hand out; warp-uniform
  int ck = -1, cdeg = 0, cidx = 0;    // this position's slot, its step count, steps generated so far
  bool creal = false;                 // false: the single empty step of a slot without triples

  for (;;)
  {
    if (!__any_sync(FULL, cursor < n || cidx < cdeg)) break;

    // ---- stream bookkeeping: which (slot, step) each of the QW entries of this position is --------
    unsigned meta = 0;                // per step: slot | 0x80 when it is the slot's last step
    int gk[QW / 2], gi[QW / 2];
    bool gv[QW / 2];
#pragma unroll
    for (int it = 0; it < QW; it++)
    {
      const bool need = cidx >= cdeg;
      const unsigned bal = __ballot_sync(FULL, need && s == 0);
      if (need)
      {
        const int q = cursor + __popc(bal & below);
        if (q < n)
        {
          ck = KORD[q];
          const int d = DEG[ck];
          creal = d > 0;
          cdeg = creal ? d : 1;
        }
        else { ck = -1; cdeg = 0; creal = false; }
        cidx = 0;
      }
      cursor += __popc(bal);
      const bool busy = cidx < cdeg;
      if (busy) meta |= (unsigned) (ck | (cidx == cdeg - 1 ? 0x80 : 0)) << (8 * it);
      if ((it & 1) == s) { gk[it >> 1] = ck; gi[it >> 1] = cidx; gv[it >> 1] = busy && creal; }
      if (busy) cidx++;
    }

    // ---- geometry: this lane's QW/2 entries (steps it = s, s + 2, ...) ------------------------------
#pragma unroll
    for (int e = 0; e < QW / 2; e++)
    {
      double* sc = scp + (2 * e + s) * QE;
      const bool v = gv[e];
      const int k = v ? gk[e] : 0;
      const int li = k * NP + (v ? gi[e] : 0);
      const int j = v ? JL[li] : 0;
      const int rcd = v ? JR[li] : 0;
      const double* tk = TJ + k * SJ;
      const double* tj = TJ + j * SJ;
      const double* rk = RK + rcd * SR;
      const double rik = tk[0], rik2 = tk[1], irik = tk[2], fik = tk[3], dfik = tk[4];
      const double ukx = tk[5], uky = tk[6], ukz = tk[7];
      const double rij2 = tj[1], irij = tj[2];
      const double fij = v ? tj[3] : 0.0;   // an empty entry contributes exactly zero
      const double rjk = rk[0], irjk = rk[1], fjk = rk[2], dfjk = rk[3];
      const double wx = (XYZ[k] - XYZ[j]) * irjk, wy = (XYZ[NP + k] - XYZ[NP + j]) * irjk,
                   wz = (XYZ[2 * NP + k] - XYZ[2 * NP + j]) * irjk;      // unit vector j -> k
      const double rjk2 = rjk * rjk;
      const double h = 0.5 * irij * irik;
      const double ct = v ? (rij2 + rik2 - rjk2) * h : 0.0;   // cos(theta_jik), :744
      const double fijk = fij * fjk;
      const double F = fijk * fik;
      const double dct_ik = (rik2 - rij2 + rjk2) * h * irik * F;  // :746, times F
      const double dct_jk = -2.0 * rjk * h * F;                   // :747, times F
      const double dF_ik = dfik * fijk, dF_jk = dfjk * (fij * fik);
      const double a = -2.0 * F * rik, b = -2.0 * F * rjk;
      const double R2 = rij2 + rik2 + rjk2;
      double E[8];
#pragma unroll
      for (int g = 0; g < NE; g++) E[g] = kb_exp_tab(-P.grp_eta[g] * R2, TAB);  // :770
#pragma unroll
      for (int g = NE; g < 8; g++) E[g] = 0.0;
      double2* o = reinterpret_cast<double2*>(sc);
      o[0] = make_double2(E[0], E[1]); o[1] = make_double2(E[2], E[3]);
      o[2] = make_double2(E[4], E[5]); o[3] = make_double2(E[6], E[7]);
      o[4] = make_double2(ct, F);
      o[5] = make_double2(dct_ik * ukx + dct_jk * wx, dct_ik * uky + dct_jk * wy);
      o[6] = make_double2(dct_ik * ukz + dct_jk * wz, dF_ik * ukx + dF_jk * wx);
      o[7] = make_double2(dF_ik * uky + dF_jk * wy, dF_ik * ukz + dF_jk * wz);
      o[8] = make_double2(a * ukx + b * wx, a * uky + b * wy);
      o[9] = make_double2(a * ukz + b * wz, 0.0);
    }
    __syncwarp();

    // ---- sweep: QW steps; a slot that ends at a step leaves for its rows right there --------------
#pragma unroll 1
    for (int it = 0; it < QW; it++)
    {
      const double2* sc = reinterpret_cast<const double2*>(scp + it * QE);
      double E[8];
      {
        const double2 e0 = sc[0], e1 = sc[1], e2 = sc[2], e3 = sc[3];
        E[0] = e0.x; E[1] = e0.y; E[2] = e1.x; E[3] = e1.y; E[4] = e2.x; E[5] = e2.y; E[6] = e3.x; E[7] = e3.y;
      }
      const double2 cF = sc[4], q5 = sc[5], q6 = sc[6], q7 = sc[7], q8 = sc[8], q9 = sc[9];
      const double v1x = q5.x, v1y = q5.y, v1z = q6.x, ux = q6.y, uy = q7.x, uz = q7.y;
      const double wx = q8.x, wy = q8.y, wz = q9.x, F = cF.y;
      double bs = fma(lam, cF.x, 1.0);
      bs = bs > 0.0 ? bs : 0.0;                       // guard of :755-767
      const double b2 = bs * bs, b3 = b2 * bs, b4 = b2 * b2, b8 = b4 * b4;
      const double b16 = b8 * b8, b15 = (b8 * b4) * b3;
      const double Az[4] = {bs, b2, b4, b16};       // (1 + lambda cos)^zeta; 2^(1-zeta) applied at the end
      const double dAz[4] = {bs > 0.0 ? lam : 0.0, 2.0 * lam * bs, 4.0 * lam * b3, 16.0 * lam * b15};
#pragma unroll
      for (int z = 0; z < 4; z++)
      {
        const double Aq = Az[z], dA = dAz[z];
        const double t1x = fma(dA, v1x, Aq * ux), t1y = fma(dA, v1y, Aq * uy), t1z = fma(dA, v1z, Aq * uz);
        const double t2x = Aq * wx, t2y = Aq * wy, t2z = Aq * wz, t0 = Aq * F;
#pragma unroll
        for (int g = SH::lo(z); g < NE; g++)
        {
          const int i = SH::off(z) + g - SH::lo(z);
          const double eta = P.grp_eta[g], Eg = E[g];
          acc[i][0] = fma(Eg, fma(eta, t2x, t1x), acc[i][0]);
          acc[i][1] = fma(Eg, fma(eta, t2y, t1y), acc[i][1]);
          acc[i][2] = fma(Eg, fma(eta, t2z, t1z), acc[i][2]);
          phi[i] = fma(Eg, t0, phi[i]);
        }
      }
      const unsigned m = meta >> (8 * it);
      if (__any_sync(FULL, m & 0x80u))
      {
        if (m & 0x80u)
        {
          const int k = (int) (m & 0x3fu);
          double* ok = blk + 3 * k;
#pragma unroll
          for (int z = 0; z < 4; z++)
          {
            const double p2 = z == 0 ? 1.0 : z == 1 ? 0.5 : z == 2 ? 0.125 : 1.0 / 32768.0;
#pragma unroll
            for (int g = SH::lo(z); g < NE; g++)
            {
              const int i = SH::off(z) + g - SH::lo(z);
              const int d = gout[i];
              if (d >= 0)
              {
                double* o = ok + (unsigned) (d * row);
                o[0] = p2 * acc[i][0]; o[1] = p2 * acc[i][1]; o[2] = p2 * acc[i][2];
              }
              acc[i][0] = acc[i][1] = acc[i][2] = 0.0;
            }
          }
        }
      }
    }
    __syncwarp();
  }

  // ---- zeta: values summed over the 16 positions (each unordered triple was visited twice) ------
#pragma unroll
  for (int i = 0; i < NI; i++)
  {
#pragma unroll
    for (int msk = 2; msk < 32; msk <<= 1) phi[i] += shfl_xor_d(phi[i], msk);
  }
  if (p == 0)
  {
#pragma unroll
    for (int z = 0; z < 4; z++)
    {
      const double p2 = z == 0 ? 0.5 : z == 1 ? 0.25 : z == 2 ? 0.0625 : 1.0 / 65536.0;
#pragma unroll
      for (int g = SH::lo(z); g < NE; g++)
      {
        const int i = SH::off(z) + g - SH::lo(z);
        const int d = gout[i];
        if (d >= 0) zrow[d] = p2 * phi[i];
      }
    }
  }

}

// ---- centre slot of the angular rows: minus the sum over the slots (translation invariance), the
// finished rows read back through L2.  Runs after the next atom's record copy has been issued, so the
// two latencies overlap; all loads of a lambda class (NI rows x 4 slot residues) are in flight at once.
template <class SH>
__device__ __forceinline__ void qsweep_centre(const kb_symfun_params& P, const SplitArgs& A, const QSmem& W,
                                              int t, int n, int lane)
{
  constexpr int NI = SH::NI, RB = NI;          // rows per trip: 4 RB loads in flight per lane
  const int D = P.n_desc, row = 3 * (n + 1);
  double* blk = static_cast<double*>(A.dz) + A.dz_ptr[t];
  const int sub = lane & 7, cc = lane >> 3;    // lanes 0..23: (slot residue mod 8, component)
  const bool on = cc < 3;
  const double* lbase = blk + (on ? 3 * sub + cc : 0);
  double* obase = blk + 3 * n + (on ? cc : 0);
  for (int m0 = 0; m0 < n; m0 += 32)           // one trip unless n > 32
  {
    const bool p0 = on && m0 + sub < n, p1 = on && m0 + sub + 8 < n, p2 = on && m0 + sub + 16 < n,
               p3 = on && m0 + sub + 24 < n;
    const bool first = m0 == 0;
#pragma unroll 1
    for (int i0 = 0; i0 < 2 * NI; i0 += RB)
    {
      double sum[RB];
      int dd[RB];
#pragma unroll
      for (int u = 0; u < RB; u++)
      {
        dd[u] = i0 + u < 2 * NI ? W.GOUT[i0 + u] : -1;
        const double* r = lbase + (unsigned) ((dd[u] < 0 ? 0 : dd[u]) * row + 3 * m0);
        double v0 = 0.0, v1 = 0.0, v2 = 0.0, v3 = 0.0;
        if (p0) v0 = __ldcg(r);
        if (p1) v1 = __ldcg(r + 24);
        if (p2) v2 = __ldcg(r + 48);
        if (p3) v3 = __ldcg(r + 72);
        sum[u] = (v0 + v1) + (v2 + v3);
      }
#pragma unroll
      for (int u = 0; u < RB; u++)
      {
        sum[u] += shfl_xor_d(sum[u], 1);
        sum[u] += shfl_xor_d(sum[u], 2);
        sum[u] += shfl_xor_d(sum[u], 4);
        if (on && sub == 0 && dd[u] >= 0)
        {
          double* o = obase + (unsigned) (dd[u] * row);
          *o = first ? -sum[u] : *o - sum[u];
        }
      }
    }
  }
  if (n == 0 && lane < 3)                       // no neighbors: the centre slot is the whole row
    for (int i = 0; i < 2 * NI; i++)
    {
      const int d = W.GOUT[i];
      if (d >= 0) blk[(unsigned) (d * row) + lane] = 0.0;
    }
  {
    const int len = (int) (A.dz_ptr[t + 1] - A.dz_ptr[t]);
    for (int q = D * row + lane; q < len; q += 32) blk[q] = 0.0;  // alignment padding
  }
}

}  // namespace
