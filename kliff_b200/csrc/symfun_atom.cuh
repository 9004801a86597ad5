// Per-atom building blocks of the warp-per-atom symmetry-function path, shared by the two-stage
// kernels (symfun_split.cu) and the warp-specialised single kernel (symfun_ws.cu):
//   prep_atom   pair-ij table, radial sets, adjacency rows, pair-jk records, slot order -> atom record
//   sweep_atom  the triple loop over one atom whose record sits in shared memory
// See symfun_split.cu for the design notes.  Reference: Descriptor::generate_one_atom
// (sym_fn.cpp:416-602), sym_d_g2 / sym_d_g4 (:627-809).
#pragma once
#include "common.cuh"
#include "symfun_math.cuh"

namespace {


constexpr int GW = KB_GROUP_WIDTH;
// doubles per scratch entry: 12 geometry + 8 C (the powers b, b^2, b^4, b^16 of b = 1 -+ cos) + 2 (the
// derivative factors 16 lambda b^15; those of zeta = 1, 2, 4 follow from the stored powers with 6
// multiplications, cheaper than loading them: the kernel's busiest unit is the shared-memory data
// pipe).  22 doubles = 44 words = 12 mod 32, so the eight lanes of a slot, which write eight consecutive
// entries at once, start on eight different 16-byte bank groups (0,12,24,4,16,28,8,20).
constexpr int SCW = 22;
constexpr int SC_DOUBLES = 32 * SCW + 2 * 32;
constexpr int SJ = 8, SR = 4;
constexpr int PREP_WARPS = 4;

// Register caps of the sweep kernel (one warp per CTA) and the resident warps per SM they allow.
// The gradient variant compiles to 230 registers uncapped (8 warps per SM); 224 gives 9 warps with a
// handful of spilled values outside the inner loop and measures 1.3 % faster (33.9 -> 33.5 ms at 1M
// atoms).  Going further does not pay (200 registers / 10 warps: 7.7 -> 8.2 ms on 216k atoms): the
// kernel is bound by the FP64 pipe and the shared-memory data pipe, not by latency.
#ifndef KB_SWEEP_REGS
#define KB_SWEEP_REGS 224
#endif
#ifndef KB_SWEEP_REGS_NOGRAD
#define KB_SWEEP_REGS_NOGRAD 128
#endif
#ifndef KB_QSWEEP_REGS
#define KB_QSWEEP_REGS 255  // queue sweep (symfun_qsweep.cuh): 88 accumulators per sweep lane; 4 warp pairs per SM
#endif
#define KB_QSWEEP_MIN_BLOCKS (65536 / (64 * ((KB_QSWEEP_REGS + 7) & ~7)))  // CTAs of two warps
#define KB_SWEEP_MIN_BLOCKS (65536 / (32 * KB_SWEEP_REGS))
#define KB_SWEEP_MIN_BLOCKS_NOGRAD (65536 / (32 * KB_SWEEP_REGS_NOGRAD))
#ifndef KB_PREP_MIN_BLOCKS
#define KB_PREP_MIN_BLOCKS 6  // 80 registers: 24 resident warps per SM, no spills
#endif

// ---- the atom record (global memory and, byte for byte, the head of the sweep warp's shared memory)
struct RecLayout
{
  int kord, deg, xyz, tj, jl, jr, rk, bytes;  // byte offsets; header int32 {n, nrec, -, -} at 0
};
__host__ __device__ inline int al16(int x) { return (x + 15) & ~15; }
__host__ __device__ inline RecLayout rec_layout(int NP, int RC)
{
  RecLayout L;
  L.kord = 16;
  L.deg = 16 + NP;
  int o = al16(16 + 2 * NP);
  L.xyz = o; o += 24 * NP;        // [3][NP] doubles
  L.tj = o;  o += 64 * NP;        // [NP][8]: r r2 1/r fc dfc ux uy uz
  L.jl = o;  o = al16(o + NP * NP);      // [NP][NP] u8: idx-th adjacent neighbor of slot k
  L.jr = o;  o = al16(o + 2 * NP * NP);  // [NP][NP] u16: its pair-jk record
  L.rk = o;  o += 32 * RC;        // [nrec][4]: r 1/r fc dfc
  L.bytes = al16(o);
  return L;
}

struct SplitArgs
{
  int lo, hi;  // centre ordinals [lo, hi) of this chunk
  const int32_t* centers;
  const int32_t* subset;
  const double* coords;
  const int32_t* species;
  const int64_t* nl_off;
  const int32_t* nl_idx;
  double* zeta;
  void* dz;
  const int64_t* dz_ptr;
  int32_t* overflow;      // [0] count, [2..] declined ordinals
  int32_t* work;          // sweep: dynamic work counter of this chunk
  unsigned char* records; // [hi - lo] records
  int NP, RC, NGP;
  int ang_lo, ang_hi;     // descriptor rows [ang_lo, ang_hi) are exactly the angular sets (else 0, 0)
};

__host__ __device__ inline size_t prep_smem_bytes(const kb_symfun_params& P, int NP, int RC)
{
  size_t b = (size_t) 24 * NP + 32 * (size_t) P.n_two + 24 * (size_t) P.n_species * P.n_species;
  b += (size_t) 8 * NP + 4 * NP + 4 * NP;           // ADJ, SPC, BASE
  b = (b + 15) & ~(size_t) 15;
  b += al16(2 * NP * NP) + al16(NP * NP);           // JR, JL (copied out as 16-byte words)
  b += (size_t) 2 * RC + 2 * (size_t) NP;           // PAIR, DEG + KORD
  return (b + 15) & ~(size_t) 15;
}

// ---- adjacency among the neighbors of one atom ------------------------------------------------
__device__ __forceinline__ int popc_m(unsigned m) { return __popc(m); }
__device__ __forceinline__ int popc_m(unsigned long long m) { return __popcll(m); }
__device__ __forceinline__ int ffs_m(unsigned m) { return __ffs((int) m); }
__device__ __forceinline__ int ffs_m(unsigned long long m) { return __ffsll((long long) m); }

// Lane = row j computes its adjacency row in registers with the reference predicate
// !(sqrt(r2) > rc) (sym_fn.cpp:729), decided on r2 against pre-squared thresholds whenever that
// is unambiguous; a warp scan numbers the (j < k) pairs ("records"); then every slot gets its
// ascending neighbor list JL with the record id JR of each entry, and PAIR lists the records.
// Returns the number of records (nothing is written beyond RC).
template <typename M>
__device__ __forceinline__ int build_lists(int n, int NP, int S, M act, const double* XYZ, const int* SPC,
                                           const double* rcut, const double* R2LO, const double* R2HI,
                                           M* ADJ, int* BASE, unsigned char* DEG, unsigned char* JL,
                                           unsigned short* JR, unsigned short* PAIR, int RC, int lane)
{
  const unsigned FULL = 0xffffffffu;
  const M one = 1;
  int nrec = 0;
  for (int j0 = 0; j0 < n; j0 += 32)
  {
    const int j = j0 + lane;
    M mask = 0;
    if (sizeof(M) == 4)
    {
      // n <= 32, one row per lane: every unordered pair once.  In round m lane j tests the pair
      // (j, j + m mod n) and hands the verdict to lane j + m through a shuffle, so n/2 rounds cover
      // the n(n-1)/2 pairs (the predicate is symmetric bit for bit: (xk - xj)^2 == (xj - xk)^2).
      const bool rowok = j < n;
      const int jc = rowok ? j : 0;
      const double xj = XYZ[jc], yj = XYZ[NP + jc], zj = XYZ[2 * NP + jc];
      const int prow = SPC[jc] * S;
      for (int m = 1; m <= n / 2; m++)
      {
        int k = jc + m;
        k = k >= n ? k - n : k;
        const double r2 = xnorm2(xsub(XYZ[k], xj), xsub(XYZ[NP + k], yj), xsub(XYZ[2 * NP + k], zj));
        const int pr = prow + SPC[k];
        bool hit = r2 < R2LO[pr];
        if (!hit && r2 <= R2HI[pr]) hit = !(sqrt(r2) > rcut[pr]);
        hit = hit && rowok;
        int src = lane - m;               // the lane whose partner in this round is this lane
        src = src < 0 ? src + n : src;
        const bool hin = __shfl_sync(FULL, (int) hit, src & 31) != 0;
        mask |= (M) hit << k;
        if (rowok) mask |= (M) hin << src;
      }
      mask &= ~(one << jc) & act;
      if (!rowok || !((act >> jc) & one)) mask = 0;
    }
    else if (j < n)
    {
      const double xj = XYZ[j], yj = XYZ[NP + j], zj = XYZ[2 * NP + j];
      const int prow = SPC[j] * S;
      for (int k = 0; k < n; k++)
      {
        const double r2 = xnorm2(xsub(XYZ[k], xj), xsub(XYZ[NP + k], yj), xsub(XYZ[2 * NP + k], zj));
        const int pr = prow + SPC[k];
        bool hit = r2 < R2LO[pr];
        if (!hit && r2 <= R2HI[pr]) hit = !(sqrt(r2) > rcut[pr]);
        mask |= (M) hit << k;
      }
      mask &= ~(one << j) & act;
      if (!((act >> j) & one)) mask = 0;
    }
    const int upper = (j + 1 < (int) (8 * sizeof(M))) ? popc_m((M) (mask >> (j + 1))) : 0;
    int inc = upper;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1)
    {
      const int y = __shfl_up_sync(FULL, inc, d);
      if (lane >= d) inc += y;
    }
    if (j < n)
    {
      ADJ[j] = mask;
      DEG[j] = (unsigned char) popc_m(mask);
      BASE[j] = nrec + inc - upper;
    }
    nrec += __shfl_sync(FULL, inc, 31);
  }
  __syncwarp();
  if (nrec > RC) return nrec;
  for (int k = lane; k < n; k += 32)
  {
    M m = ADJ[k];
    int idx = 0;
    int up = BASE[k];  // records of pairs (k, j > k) are numbered consecutively from BASE[k]
    while (m)
    {
      const int j = ffs_m(m) - 1;
      m &= m - one;
      int slot;
      if (j > k) { slot = up++; PAIR[slot] = (unsigned short) (k | (j << 8)); }
      else
      {
        const M between = (M) (ADJ[j] >> (j + 1)) & ((one << (k - j - 1)) - one);
        slot = BASE[j] + popc_m(between);
      }
      JL[k * NP + idx] = (unsigned char) j;
      JR[k * NP + idx] = (unsigned short) slot;
      idx++;
    }
  }
  __syncwarp();
  return nrec;
}

// ---- shared-memory views -----------------------------------------------------------------------
struct PrepSmem
{
  double *XYZ, *Z2, *RCI, *R2LO, *R2HI;
  unsigned long long* ADJ;
  int *SPC, *BASE;
  unsigned short *JR, *PAIR;
  unsigned char *JL, *DEG, *KORD;
};

__device__ __forceinline__ PrepSmem prep_carve(unsigned char* base, const kb_symfun_params& P, int NP, int RC)
{
  const int NT = P.n_two, S = P.n_species;
  PrepSmem W;
  W.XYZ = reinterpret_cast<double*>(base);      // [3][NP]
  W.Z2 = W.XYZ + 3 * NP;                        // [NT] radial zeta, [3 NT] radial centre slot
  W.RCI = W.Z2 + 4 * NT;                        // [S][S] 1 / rc
  W.R2LO = W.RCI + S * S;                       // [S][S] rc^2 (1 - 1e-12): surely inside
  W.R2HI = W.R2LO + S * S;                      // [S][S] rc^2 (1 + 1e-12): surely outside beyond
  W.ADJ = reinterpret_cast<unsigned long long*>(W.R2HI + S * S);  // [NP]
  W.SPC = reinterpret_cast<int*>(W.ADJ + NP);   // [NP]
  W.BASE = W.SPC + NP;                          // [NP]
  unsigned char* p = reinterpret_cast<unsigned char*>(W.BASE + NP);
  p = base + al16((int) (p - base));
  W.JR = reinterpret_cast<unsigned short*>(p);  p += al16(2 * NP * NP);
  W.JL = p;                                     p += al16(NP * NP);
  W.PAIR = reinterpret_cast<unsigned short*>(p);                 // [RC]
  W.DEG = reinterpret_cast<unsigned char*>(W.PAIR + RC);         // [NP]
  W.KORD = W.DEG + NP;                                           // [NP]
  return W;
}

// cutoff tables of one warp (once per kernel)
__device__ __forceinline__ void prep_tables(const PrepSmem& W, const kb_symfun_params& P, int lane)
{
  const int S = P.n_species;
  for (int q = lane; q < S * S; q += 32)
  {
    const double rc = P.rcut[q], rc2 = rc * rc;
    W.RCI[q] = 1.0 / rc;
    W.R2LO[q] = rc2 * (1.0 - 1e-12);
    W.R2HI[q] = rc2 * (1.0 + 1e-12);
  }
  __syncwarp();
}

// ---- phases A + D of stage 1: pair-ij table (record) and the radial sets (final rows) ----------
// Lanes = neighbors.  Returns the mask of neighbors within their species cutoff (sym_fn.cpp:451).
// SOA selects the layout of the pair-ij table in the record: [NP][8] (two-stage sweep) or [8][NP]
// (lane sweep, symfun_lane.cu).
template <bool GRAD, typename TOut, bool SOA>
__device__ __forceinline__ unsigned long long prep_pairs(const kb_symfun_params& P, const SplitArgs& A,
                                                         const PrepSmem& W, int i, int si, int64_t off, int n,
                                                         TOut* blk, double* zrow, double* gXYZ, double* gTJ,
                                                         int lane)
{
  const unsigned FULL = 0xffffffffu;
  const int NP = A.NP, NT = P.n_two, S = P.n_species;
  double *XYZ = W.XYZ, *Z2 = W.Z2, *RCI = W.RCI;
  int* SPC = W.SPC;
  const int row = 3 * (n + 1);
  const double xi = A.coords[3 * i], yi = A.coords[3 * i + 1], zi = A.coords[3 * i + 2];
  __syncwarp();
  for (int q = lane; q < 4 * NT; q += 32) Z2[q] = 0.0;
  __syncwarp();

  // ---- phases A + D: pair-ij table and radial sets; lanes = neighbors ----------------------
  unsigned long long ACT = 0ull;
  for (int j0 = 0; j0 < n; j0 += 32)
  {
    const int jj = j0 + lane;
    bool act = false;
    double r = 1.0, fc = 0.0, dfc = 0.0, ux = 0.0, uy = 0.0, uz = 0.0;
    if (jj < n)
    {
      const int j = A.nl_idx[off + jj];
      const int sj = A.species[j];
      const double xj = A.coords[3 * j], yj = A.coords[3 * j + 1], zj = A.coords[3 * j + 2];
      const double vx = xsub(xj, xi), vy = xsub(yj, yi), vz = xsub(zj, zi);
      const double r2x = xnorm2(vx, vy, vz);
      r = sqrt(r2x);                  // sym_fn.cpp:447-448
      const double rc = P.rcut[si * S + sj];
      act = !(r > rc);                // :451
      const double ir = rsqrt(r2x);
      kb_cutoff_fast(r, rc, RCI[si * S + sj], fc, dfc);
      if (!act) { fc = 0.0; dfc = 0.0; }
      ux = vx * ir; uy = vy * ir; uz = vz * ir;
      XYZ[jj] = xj; XYZ[NP + jj] = yj; XYZ[2 * NP + jj] = zj;
      gXYZ[jj] = xj; gXYZ[NP + jj] = yj; gXYZ[2 * NP + jj] = zj;
      SPC[jj] = sj;
      if (SOA)
      {
        // field-major table [8][NP]: lanes = consecutive neighbors, so each field is one coalesced store
        gTJ[jj] = r; gTJ[NP + jj] = r2x; gTJ[2 * NP + jj] = ir; gTJ[3 * NP + jj] = fc;
        gTJ[4 * NP + jj] = dfc; gTJ[5 * NP + jj] = ux; gTJ[6 * NP + jj] = uy; gTJ[7 * NP + jj] = uz;
      }
      else
      {
        double2* tj = reinterpret_cast<double2*>(gTJ + jj * SJ);
        tj[0] = make_double2(r, r2x); tj[1] = make_double2(ir, fc);
        tj[2] = make_double2(dfc, ux); tj[3] = make_double2(uy, uz);
      }
    }
    ACT |= (unsigned long long) __ballot_sync(FULL, act) << j0;
    for (int q0 = 0; q0 < NT; q0 += 4)
    {
      double v[16];
#pragma unroll
      for (int u = 0; u < 4; u++)  // 4 independent radial sets per trip
      {
        const int q = q0 + u;
        double phi = 0.0, dphi = 0.0;
        if (q < NT && act) kb_radial(P.two_kind[q], P.two_p0[q], P.two_p1[q], r, fc, dfc, phi, dphi);
        const double px = dphi * ux, py = dphi * uy, pz = dphi * uz;  // :497
        if (GRAD && q < NT && jj < n)
        {
          TOut* o = blk + (unsigned) (P.two_out[q] * row + 3 * jj);
          o[0] = (TOut) px; o[1] = (TOut) py; o[2] = (TOut) pz;
        }
        v[4 * u] = phi; v[4 * u + 1] = px; v[4 * u + 2] = py; v[4 * u + 3] = pz;
      }
      const double tot = warp_reduce16(v, lane);  // lanes 0-15: total of value (set l/4, component l%4)
      const int q = q0 + ((lane & 15) >> 2), comp = lane & 3;
      if (q < NT && lane < 16)
      {
        if (comp == 0) Z2[q] += tot;
        else Z2[NT + 3 * q + comp - 1] += tot;
      }
    }
  }
  __syncwarp();
  for (int q = lane; q < NT; q += 32)
  {
    zrow[P.two_out[q]] = Z2[q];
    if (GRAD)
    {
      TOut* o = blk + (unsigned) (P.two_out[q] * row + 3 * n);  // centre slot, :499-501
      o[0] = (TOut) -Z2[NT + 3 * q]; o[1] = (TOut) -Z2[NT + 3 * q + 1]; o[2] = (TOut) -Z2[NT + 3 * q + 2];
    }
  }

  return ACT;
}

// ---- phase C of stage 1: pair-jk records {r, 1/r, fc, dfc}; lanes = records, two per trip -------
// SOA: field-major [4][stride] (lane sweep) instead of [nrec][4].
template <bool SOA>
__device__ __forceinline__ void prep_pair_records(const kb_symfun_params& P, const PrepSmem& W, int NP, int nrec,
                                                  int stride, double* gRK, int lane)
{
  const int S = P.n_species;
  const double *XYZ = W.XYZ, *RCI = W.RCI;
  const int* SPC = W.SPC;
  const unsigned short* PAIR = W.PAIR;
  for (int r0 = lane; r0 < nrec; r0 += 64)
  {
    const int r1 = r0 + 32;
    const bool two = r1 < nrec;
    const int ja = PAIR[r0] & 0xff, ka = PAIR[r0] >> 8;
    const int pb = two ? PAIR[r1] : PAIR[r0];
    const int jb = pb & 0xff, kb = pb >> 8;
    const double r2a = xnorm2(xsub(XYZ[ka], XYZ[ja]), xsub(XYZ[NP + ka], XYZ[NP + ja]),
                              xsub(XYZ[2 * NP + ka], XYZ[2 * NP + ja]));
    const double r2b = xnorm2(xsub(XYZ[kb], XYZ[jb]), xsub(XYZ[NP + kb], XYZ[NP + jb]),
                              xsub(XYZ[2 * NP + kb], XYZ[2 * NP + jb]));
    const double ra = sqrt(r2a), rb = sqrt(r2b);
    const double ira = rsqrt(r2a), irb = rsqrt(r2b);
    double fca, dfca, fcb, dfcb;
    const int pa = SPC[ja] * S + SPC[ka], pb2 = SPC[jb] * S + SPC[kb];
    kb_cutoff_fast(ra, P.rcut[pa], RCI[pa], fca, dfca);
    kb_cutoff_fast(rb, P.rcut[pb2], RCI[pb2], fcb, dfcb);
    if (SOA)
    {
      gRK[r0] = ra; gRK[stride + r0] = ira; gRK[2 * stride + r0] = fca; gRK[3 * stride + r0] = dfca;
      if (two) { gRK[r1] = rb; gRK[stride + r1] = irb; gRK[2 * stride + r1] = fcb; gRK[3 * stride + r1] = dfcb; }
    }
    else
    {
      double2* rka = reinterpret_cast<double2*>(gRK + r0 * SR);
      rka[0] = make_double2(ra, ira); rka[1] = make_double2(fca, dfca);
      if (two)
      {
        double2* rkb = reinterpret_cast<double2*>(gRK + r1 * SR);
        rkb[0] = make_double2(rb, irb); rkb[1] = make_double2(fcb, dfcb);
      }
    }
  }

}

// ---- phase S of stage 1: slots ordered by adjacency degree (descending, index as tie-break) ----
__device__ __forceinline__ void prep_sort_slots(int n, const unsigned char* DEG, unsigned char* KORD, int lane)
{
  const unsigned FULL = 0xffffffffu;
  if (n <= 32)
  {
    // one slot per lane, degrees < 32: bit-serial ballots from the top bit down leave in `tied` the
    // lanes with the same degree and count in `above` the slots with a larger one
    const int dk = lane < n ? DEG[lane] : 0;
    unsigned tied = n == 32 ? FULL : ((1u << n) - 1u);
    int above = 0;
#pragma unroll
    for (int b = 4; b >= 0; b--)
    {
      const unsigned bal = __ballot_sync(FULL, (dk >> b) & 1);
      if ((dk >> b) & 1) tied &= bal;
      else { above += __popc(tied & bal); tied &= ~bal; }
    }
    if (lane < n) KORD[above + __popc(tied & ((1u << lane) - 1u))] = (unsigned char) lane;
  }
  else
    for (int k = lane; k < n; k += 32)
    {
      const int dk = DEG[k];
      int rank = 0;
      for (int m = 0; m < n; m++) rank += (DEG[m] > dk) || (DEG[m] == dk && m < k);
      KORD[rank] = (unsigned char) k;
    }
  __syncwarp();

}

// =============================================================================================
// stage 1 for one atom: centre ordinal t -> record `rec` (header n = -1 when the atom is declined)
// =============================================================================================
// SOA: pair-ij table [8][NP] and pair-jk records [4][stride], stride = nrec rounded up to even (the
// round-2 sweep gathers single fields of different neighbors: field-major tables spread them over the banks)
template <bool GRAD, typename TOut, bool SOA = false>
__device__ __forceinline__ void prep_atom(const kb_symfun_params& P, const SplitArgs& A, const RecLayout& L,
                                          const PrepSmem& W, int t, unsigned char* rec, int lane)
{
  const unsigned FULL = 0xffffffffu;
  const int NP = A.NP, RC = A.RC, D = P.n_desc, NT = P.n_two, S = P.n_species;
  double *XYZ = W.XYZ, *Z2 = W.Z2, *RCI = W.RCI, *R2LO = W.R2LO, *R2HI = W.R2HI;
  unsigned long long* ADJ = W.ADJ;
  int *SPC = W.SPC, *BASE = W.BASE;
  unsigned short *JR = W.JR, *PAIR = W.PAIR;
  unsigned char *JL = W.JL, *DEG = W.DEG, *KORD = W.KORD;
  int* hdr = reinterpret_cast<int*>(rec);
  const int i = A.centers ? A.centers[t] : t;
  const int64_t off = A.nl_off[i];
  const int n = (int) (A.nl_off[i + 1] - off);
  if (n > NP)
  {
    if (lane == 0) { A.overflow[2 + atomicAdd(A.overflow, 1)] = t; hdr[0] = -1; hdr[1] = 0; }
    return;
  }
  const int si = A.species[i];
  TOut* blk = GRAD ? static_cast<TOut*>(A.dz) + A.dz_ptr[t] : nullptr;
  double* zrow = A.zeta + (int64_t) t * D;
  double* gXYZ = reinterpret_cast<double*>(rec + L.xyz);
  double* gTJ = reinterpret_cast<double*>(rec + L.tj);
  double* gRK = reinterpret_cast<double*>(rec + L.rk);

  const unsigned long long ACT =
      prep_pairs<GRAD, TOut, SOA>(P, A, W, i, si, off, n, blk, zrow, gXYZ, gTJ, lane);

  // ---- phase B: adjacency rows + per-slot lists (32-bit masks when every list fits in 32) ----
  const int nrec = NP <= 32
      ? build_lists<unsigned>(n, NP, S, (unsigned) ACT, XYZ, SPC, P.rcut, R2LO, R2HI,
                              reinterpret_cast<unsigned*>(ADJ), BASE, DEG, JL, JR, PAIR, RC, lane)
      : build_lists<unsigned long long>(n, NP, S, ACT, XYZ, SPC, P.rcut, R2LO, R2HI, ADJ, BASE, DEG, JL,
                                        JR, PAIR, RC, lane);
  if (nrec > RC)
  {
    if (lane == 0) { A.overflow[2 + atomicAdd(A.overflow, 1)] = t; hdr[0] = -1; hdr[1] = 0; }
    return;
  }

  prep_pair_records<SOA>(P, W, NP, nrec, (nrec + 1) & ~1, gRK, lane);

  prep_sort_slots(n, DEG, KORD, lane);

  // ---- record: lists, order, header ----------------------------------------------------------
  {
    const uint4* s4 = reinterpret_cast<const uint4*>(JL);
    uint4* g4 = reinterpret_cast<uint4*>(rec + L.jl);
    for (int q = lane; q < al16(NP * NP) / 16; q += 32) g4[q] = s4[q];
    s4 = reinterpret_cast<const uint4*>(JR);
    g4 = reinterpret_cast<uint4*>(rec + L.jr);
    for (int q = lane; q < al16(2 * NP * NP) / 16; q += 32) g4[q] = s4[q];
    for (int q = lane; q < n; q += 32) { rec[L.kord + q] = KORD[q]; rec[L.deg + q] = DEG[q]; }
    if (lane == 0) { hdr[0] = n; hdr[1] = nrec; }
  }
}

// ---- sweep side ----------------------------------------------------------------------------------
struct SweepSmem
{
  const unsigned char *KORD, *DEG, *JL;
  const unsigned short* JR;
  const double *XYZ, *TJ, *RK;
  double *SC, *TAB;
};

__device__ __forceinline__ SweepSmem sweep_carve(unsigned char* base, const RecLayout& L)
{
  SweepSmem W;
  W.KORD = base + L.kord;
  W.DEG = base + L.deg;
  W.XYZ = reinterpret_cast<const double*>(base + L.xyz);
  W.TJ = reinterpret_cast<const double*>(base + L.tj);
  W.JL = base + L.jl;
  W.JR = reinterpret_cast<const unsigned short*>(base + L.jr);
  W.RK = reinterpret_cast<const double*>(base + L.rk);
  W.SC = reinterpret_cast<double*>(base + L.bytes);  // [32][SCW] (+ padding)
  W.TAB = W.SC + SC_DOUBLES;                         // [64] 2^(j/64)
  return W;
}

// per-lane constants of the triple loop: lane = (slot k_sub, eta-group g)
struct SweepLane
{
  int KPW, ksub, g;
  bool valid_g;
  double eta;
  int gout[GW];
};

__device__ __forceinline__ SweepLane sweep_lane(const kb_symfun_params& P, int NGP, int lane)
{
  SweepLane Q;
  Q.KPW = 32 / NGP;
  Q.ksub = lane / NGP;
  Q.g = lane & (NGP - 1);
  Q.valid_g = Q.g < P.n_groups;
  Q.eta = 0.0;
#pragma unroll
  for (int e = 0; e < GW; e++) Q.gout[e] = -1;
  if (Q.valid_g)
  {
    Q.eta = P.grp_eta[Q.g];
#pragma unroll
    for (int e = 0; e < GW; e++) Q.gout[e] = P.grp_out[Q.g][e];
  }
  return Q;
}

// scratch and exponential table of one sweep warp (once per kernel)
__device__ __forceinline__ void sweep_tables(const SweepSmem& W, int lane)
{
  kb_exp_table_fill(W.TAB, lane);
  for (int q = lane; q < SC_DOUBLES; q += 32) W.SC[q] = 0.0;  // empty scratch entries must be finite
  __syncwarp();
}

// =============================================================================================
// stage 2 for one atom whose record (n neighbors) has been copied to shared memory
// =============================================================================================
template <bool GRAD, typename TOut>
__device__ __forceinline__ void sweep_atom(const kb_symfun_params& P, const SplitArgs& A, const SweepSmem& W,
                                           const SweepLane& Q, int t, int n, int lane)
{
  const unsigned FULL = 0xffffffffu;
  const int NP = A.NP, NGP = A.NGP, D = P.n_desc;
  const unsigned char *KORD = W.KORD, *DEG = W.DEG, *JL = W.JL;
  const unsigned short* JR = W.JR;
  const double *XYZ = W.XYZ, *TJ = W.TJ, *RK = W.RK;
  double *SC = W.SC, *TAB = W.TAB;
  const int KPW = Q.KPW, ksub = Q.ksub, g = Q.g;
  const bool valid_g = Q.valid_g;
  const double eta = Q.eta;
  int gout[GW];
#pragma unroll
  for (int e = 0; e < GW; e++) gout[e] = Q.gout[e];
  const double p2tab[GW] = {1.0, 1.0, 0.5, 0.5, 0.125, 0.125, 1.0 / 32768.0, 1.0 / 32768.0};
  const int row = 3 * (n + 1);
  TOut* blk = GRAD ? static_cast<TOut*>(A.dz) + A.dz_ptr[t] : nullptr;
  double* zrow = A.zeta + (int64_t) t * D;
  if (GRAD && A.ang_hi > A.ang_lo)
  {
    // (optional, KB_PREZERO=1) Zero the angular rows with full, coalesced sector writes first.  The row pieces written
    // below are 24 bytes at 8-byte alignment, i.e. partial 32-byte sectors: a partial write to a
    // sector that is not in L2 makes L2 fetch it from DRAM (ncu: 13.5 KB of reads per atom that the
    // algorithm never asks for); after this pass every later piece lands on a resident dirty line.
    TOut* z0 = blk + (unsigned) (A.ang_lo * row);
    const int len = (A.ang_hi - A.ang_lo) * row;
    for (int q = lane; q < len; q += 32) z0[q] = (TOut) 0;
  }
  __syncwarp();

  // ---- phase E: triple loop (see symfun_warp.cu for the derivation) ---------------------------
  double csum[GW][3], phi[GW];
#pragma unroll
  for (int e = 0; e < GW; e++) { csum[e][0] = csum[e][1] = csum[e][2] = 0.0; phi[e] = 0.0; }

  double acc[GW][3];
  auto sweep_step = [&](const double* sc) {
    const double2 sF = *reinterpret_cast<const double2*>(sc);  // {rij^2 + rik^2 + rjk^2, F}
    const double F = sF.y;
    const double E = kb_exp_tab(-eta * sF.x, TAB);  // e^{-eta (rij^2 + rik^2 + rjk^2)}, :770
    const double eF = E * F, Ee = E * eta;
    const double v1x = eF * sc[2], v1y = eF * sc[3], v1z = eF * sc[4];
    const double v2x = fma(Ee, sc[5], E * sc[8]), v2y = fma(Ee, sc[6], E * sc[9]),
                 v2z = fma(Ee, sc[7], E * sc[10]);
    // entries e = 2 zi + si, zeta = {1,2,4,16}[zi], lambda = {-1,+1}[si]:  C = b^zeta,
    // C' = zeta lambda b^(zeta-1) (0 when b = 0: the guard of sym_fn.cpp:755-767)
    double2 c2[4], d2[4];
#pragma unroll
    for (int zi = 0; zi < 4; zi++) c2[zi] = *reinterpret_cast<const double2*>(sc + 12 + 2 * zi);
    d2[3] = *reinterpret_cast<const double2*>(sc + 20);
    d2[0].x = __hiloint2double(__double2hiint(c2[0].x) > 0 ? 0xbff00000 : 0, 0);  // b > 0 ? -1 : 0
    d2[0].y = __hiloint2double(__double2hiint(c2[0].y) > 0 ? 0x3ff00000 : 0, 0);  // b > 0 ? +1 : 0
    d2[1].x = -2.0 * c2[0].x;            d2[1].y = 2.0 * c2[0].y;
    d2[2].x = -4.0 * (c2[0].x * c2[1].x); d2[2].y = 4.0 * (c2[0].y * c2[1].y);
#pragma unroll
    for (int zi = 0; zi < 4; zi++)
    {
      const int e = 2 * zi;
      phi[e] = fma(c2[zi].x, eF, phi[e]);
      phi[e + 1] = fma(c2[zi].y, eF, phi[e + 1]);
      if (GRAD)
      {
        acc[e][0] = fma(c2[zi].x, v2x, fma(d2[zi].x, v1x, acc[e][0]));
        acc[e][1] = fma(c2[zi].x, v2y, fma(d2[zi].x, v1y, acc[e][1]));
        acc[e][2] = fma(c2[zi].x, v2z, fma(d2[zi].x, v1z, acc[e][2]));
        acc[e + 1][0] = fma(c2[zi].y, v2x, fma(d2[zi].y, v1x, acc[e + 1][0]));
        acc[e + 1][1] = fma(c2[zi].y, v2y, fma(d2[zi].y, v1y, acc[e + 1][1]));
        acc[e + 1][2] = fma(c2[zi].y, v2z, fma(d2[zi].y, v1z, acc[e + 1][2]));
      }
    }
  };

  const int nwi = (n + KPW - 1) / KPW;
  for (int wi = 0; wi < nwi; wi++)
  {
    const int kk = wi * KPW + ksub;
    const bool valid_k = kk < n;
    const int k = valid_k ? KORD[kk] : 0;
    const int degk = valid_k ? DEG[k] : 0;
    const int maxdeg = __reduce_max_sync(FULL, degk);
    double* scg = SC + ksub * (NGP * SCW + 2);  // this slot's scratch group

#pragma unroll
    for (int e = 0; e < GW; e++) acc[e][0] = acc[e][1] = acc[e][2] = 0.0;

    for (int s0 = 0; s0 < maxdeg; s0 += NGP)
    {
      double* sc = scg + g * SCW;
      if (s0 + g < degk)
      {
        const double* tk = TJ + k * SJ;
        const double rik = tk[0], rik2 = tk[1], irik = tk[2], fik = tk[3], dfik = tk[4];
        const double ukx = tk[5], uky = tk[6], ukz = tk[7];
        const int j = JL[k * NP + s0 + g];
        const int rcd = JR[k * NP + s0 + g];
        const double* tj = TJ + j * SJ;
        const double* rk = RK + rcd * SR;
        const double rij2 = tj[1], irij = tj[2], fij = tj[3];
        const double rjk = rk[0], irjk = rk[1], fjk = rk[2], dfjk = rk[3];
        const double wx = (XYZ[k] - XYZ[j]) * irjk, wy = (XYZ[NP + k] - XYZ[NP + j]) * irjk,
                     wz = (XYZ[2 * NP + k] - XYZ[2 * NP + j]) * irjk;      // unit vector j -> k
        const double rjk2 = rjk * rjk;
        const double h = 0.5 * irij * irik;
        const double ct = (rij2 + rik2 - rjk2) * h;            // cos(theta_jik), :744
        const double dct_ik = (rik2 - rij2 + rjk2) * h * irik;  // :746
        const double dct_jk = -2.0 * rjk * h;                   // :747
        const double fijk = fij * fjk;
        const double F = fijk * fik;
        const double dF_ik = dfik * fijk, dF_jk = dfjk * (fij * fik);
        const double a = -2.0 * F * rik, b = -2.0 * F * rjk;
        sc[0] = rij2 + rik2 + rjk2;
        sc[1] = F;
        sc[2] = dct_ik * ukx + dct_jk * wx; sc[3] = dct_ik * uky + dct_jk * wy; sc[4] = dct_ik * ukz + dct_jk * wz;
        sc[5] = a * ukx + b * wx;           sc[6] = a * uky + b * wy;           sc[7] = a * ukz + b * wz;
        sc[8] = dF_ik * ukx + dF_jk * wx;   sc[9] = dF_ik * uky + dF_jk * wy;   sc[10] = dF_ik * ukz + dF_jk * wz;
#pragma unroll
        for (int si2 = 0; si2 < 2; si2++)  // (1 +- cos)^{1,2,4,16} and 16 lambda (1 +- cos)^15, :749-767
        {
          const double lam = si2 ? 1.0 : -1.0;
          double bse = fma(lam, ct, 1.0);
          bse = bse > 0.0 ? bse : 0.0;  // guard of :755-767
          const double b2 = bse * bse, b3 = b2 * bse, b4 = b2 * b2, b8 = b4 * b4;
          const double b16 = b8 * b8, b15 = b8 * b4 * b3;
          sc[12 + 0 + si2] = bse;
          sc[12 + 2 + si2] = b2;
          sc[12 + 4 + si2] = b4;
          sc[12 + 6 + si2] = b16;
          sc[20 + si2] = 16.0 * lam * b15;
        }
      }
      else
      {
#pragma unroll
        for (int q = 0; q < 11; q++) sc[q] = 0.0;  // empty entry: contributes exactly zero
      }
      const int left = maxdeg - s0;
      const int maxcnt = left < NGP ? left : NGP;  // warp-uniform
      __syncwarp();

      int ej = 0;
#pragma unroll 1
      for (; ej + 1 < maxcnt; ej += 2)
      {
        sweep_step(scg + ej * SCW);
        sweep_step(scg + (ej + 1) * SCW);
      }
      if (ej < maxcnt) sweep_step(scg + ej * SCW);
      __syncwarp();
    }

    if (GRAD && valid_k && valid_g)
    {
      // all row addresses first (distinct registers), then the stores back to back: with the
      // address formed next to each store the compiler recycles one register pair and every
      // store waits for the previous one to leave (ncu: long-scoreboard stalls on the address)
      TOut* o[GW];
#pragma unroll
      for (int e = 0; e < GW; e++)
      {
        o[e] = blk + (unsigned) ((gout[e] < 0 ? 0 : gout[e]) * row + 3 * k);
        asm volatile("" : "+l"(o[e]));  // keep the eight addresses in eight register pairs
        __builtin_assume(__isGlobal(o[e]));
      }
#pragma unroll
      for (int e = 0; e < GW; e++)
      {
        csum[e][0] += acc[e][0]; csum[e][1] += acc[e][1]; csum[e][2] += acc[e][2];
        if (gout[e] >= 0)
        {
          o[e][0] = (TOut) (p2tab[e] * acc[e][0]); o[e][1] = (TOut) (p2tab[e] * acc[e][1]);
          o[e][2] = (TOut) (p2tab[e] * acc[e][2]);
        }
      }
    }
  }

  // ---- phase F: centre slot (translation invariance) and zeta, reduced over the slot lanes ----
#pragma unroll
  for (int e = 0; e < GW; e++)
    for (int m = NGP; m < 32; m <<= 1)
    {
      phi[e] += shfl_xor_d(phi[e], m);
      if (GRAD)
      {
        csum[e][0] += shfl_xor_d(csum[e][0], m);
        csum[e][1] += shfl_xor_d(csum[e][1], m);
        csum[e][2] += shfl_xor_d(csum[e][2], m);
      }
    }
  if (ksub == 0)
  {
    TOut* o[GW];
    double* z[GW];
#pragma unroll
    for (int e = 0; e < GW; e++)
    {
      const int d = gout[e] < 0 ? 0 : gout[e];
      z[e] = zrow + d;
      o[e] = GRAD ? blk + (unsigned) (d * row + 3 * n) : nullptr;
      asm volatile("" : "+l"(o[e]), "+l"(z[e]));  // distinct address registers: stores back to back
      __builtin_assume(__isGlobal(o[e]));
      __builtin_assume(__isGlobal(z[e]));
    }
#pragma unroll
    for (int e = 0; e < GW; e++)
      if (gout[e] >= 0)
      {
        *z[e] = 0.5 * p2tab[e] * phi[e];  // each unordered triple was visited twice
        if (GRAD)
        {
          o[e][0] = (TOut) (-p2tab[e] * csum[e][0]); o[e][1] = (TOut) (-p2tab[e] * csum[e][1]);
          o[e][2] = (TOut) (-p2tab[e] * csum[e][2]);
        }
      }
  }
  if (GRAD)
  {
    const int len = (int) (A.dz_ptr[t + 1] - A.dz_ptr[t]);
    for (int q = D * row + lane; q < len; q += 32) blk[q] = (TOut) 0;  // alignment padding
  }
}

}  // namespace
