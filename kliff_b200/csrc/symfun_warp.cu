// Warp-per-atom symmetry-function + dG/dr kernel (fp64) for sm_100a — the hot kernel, v2.
//
// Replaces the per-atom call chain sym_fn.py:155-160 -> Descriptor::generate_one_atom
// (sym_fn.cpp:416-602) -> sym_d_g2 / sym_d_g4 (:627-809) for radial g1/g2/g3 and angular g4
// sets with (zeta, lambda) in {1,2,4,16} x {-1,+1} (set51, set30 and every Behler-style set).
// Anything else (g5, other exponents, n > 64) is declined and taken by symfun_tiled.cu /
// symfun_general.cu.
//
// Why this shape (ncu evidence of the CTA-per-atom predecessor, profiles/r01_ncu_symfun_tiled_v1.txt):
// 40 % of warp time sat in CTA barriers between the table / record / triple / copy phases, lanes were
// 74 % utilised, and every eta-lane recomputed the triple geometry.  Here ONE WARP owns an atom end
// to end — no __syncthreads, only __syncwarp — and 8 such warps are resident per SM, each with its
// own ~22 KB of shared memory, so one warp's latency-bound table phases overlap another warp's
// FMA-bound triple loop on the same FP64 pipe.
//
// Per atom:
//   A  pair-ij table   lanes = neighbors: r, r^2, 1/r, unit vector, fc, fc'                         [smem]
//   D  radial sets     lanes = neighbors, 8 sets at a time: phi, dphi -> rows written straight to
//                      global; zeta and centre slot by a 31-step transpose reduction
//   B  pair-jk test    lane = row j computes its whole 64-bit adjacency row in registers with the
//                      reference predicate r_jk <= rc_jk (sym_fn.cpp:729); record slots from a warp scan;
//                      per-slot neighbor / record lists (no ballots, no atomics)
//   C  pair-jk records lanes = records (two per trip): r, 1/r, fc, fc'                               [smem]
//   S  slots sorted by adjacency degree, so that the 32/NGP slots sharing a warp-item have the same
//      trip count (diamond Si: degrees 18/12/10 -> zero divergence)
//   E  triple loop     item = (slot k, eta-group g), lane = (k_sub, g).  Each step, the NGP lanes of a
//      slot first compute, ONCE per ordered pair (j -> k): cos, F, sum of squared distances, the three
//      3-vectors W1, P, R and the power chain (1 +- cos)^{1,2,4,16} with its derivative factors, into a
//      bank-padded scratch (28 doubles per pair); then every eta-lane sweeps those entries with one
//      exponential E_g = e^{-eta_g (rij^2+rik^2+rjk^2)} (Estrin polynomial, dependency depth 5) and a
//      rank-2 update of its 8x3 register accumulators:
//            acc[e][c] += C'_e * (E_g F W1[c]) + C_e * (E_g (eta_g P[c] + R[c]))
//      Only the derivative w.r.t. slot k is accumulated, so there are no write conflicts; every
//      unordered triple is visited from both ends like the reference's slots jj/kk (:587-593).
//      (An earlier revision tabulated e^{-eta_g r^2} per pair in shared memory; dropping those tables
//      removed 1 372 exponentials per atom from the table phases and 10 KB of shared memory per warp
//      for one Estrin exponential per sweep step — measured 10.1 -> 9.4 ms on 216 k atoms.)
//   F  centre slot = -sum_k slot_k (translation invariance) and zeta = 1/2 sum over ordered pairs are
//      carried in registers across items and reduced once per atom with xor-shuffles.
#include "common.cuh"
#include "symfun_math.cuh"

#include <cstdlib>

namespace {

constexpr int GW = KB_GROUP_WIDTH;
// resident warps per SM the kernel is compiled for: the gradient variant needs ~232 registers (8x3
// accumulators + 8x3 centre partial sums per lane) and spilling them costs more than the extra
// warps give (measured: 10 -> 11.4 ms vs 8 -> 9.4 ms on 216k atoms); the value-only variant fits 12.
#ifndef KB_WARP_MIN_BLOCKS
#define KB_WARP_MIN_BLOCKS 8
#endif
#ifndef KB_WARP_MIN_BLOCKS_NOGRAD
#define KB_WARP_MIN_BLOCKS_NOGRAD 12
#endif
constexpr int SCW = 28;  // doubles per scratch entry: 12 geometry + 8 C + 8 C'
// scratch: 32 entries; each group of NGP entries is followed by 2 doubles of padding so that the
// 32/NGP groups, which are read simultaneously (broadcast inside a group), fall on different banks
constexpr int SC_DOUBLES = 32 * SCW + 2 * 32;

struct WarpArgs
{
  int n_centers;
  const int32_t* centers;
  const int32_t* subset;  // optional list of centre ordinals (second pass over overflow atoms)
  const double* coords;
  const int32_t* species;
  const int64_t* nl_off;
  const int32_t* nl_idx;
  double* zeta;
  void* dz;           // packed blocks, element type chosen by the kernel template
  const int64_t* dz_ptr;
  int32_t* overflow;  // [0] count, [1] work counter, [2..] ordinals this kernel declined
  int NP;             // neighbor capacity (even, <= 64)
  int RC;             // pair-jk record capacity
  int NGP;            // power of two >= n_groups, <= 32
};

__host__ __device__ inline size_t warp_smem_bytes(const kb_symfun_params& P, int NP, int RC)
{
  const int NT = P.n_two;
  size_t d = (size_t) 3 * NP + (size_t) NP * 8 + (size_t) RC * 4 + SC_DOUBLES + 4 * (size_t) NT
             + (size_t) P.n_species * P.n_species + 64;
  size_t bytes = d * 8 + (size_t) NP * 8;        // + adjacency masks
  bytes += (size_t) NP * 8;                      // species, record base
  bytes += (size_t) (NP * NP + RC) * 2;          // per-slot record lists, pair list
  bytes += (size_t) NP * NP + 2 * (size_t) NP;   // per-slot neighbor lists, degree, slot order
  return (bytes + 15) & ~(size_t) 15;
}

template <bool GRAD, typename TOut>
__global__ void __launch_bounds__(32, GRAD ? KB_WARP_MIN_BLOCKS : KB_WARP_MIN_BLOCKS_NOGRAD)
symfun_warp_kernel(const __grid_constant__ kb_symfun_params P, const WarpArgs A)
{
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x;
  const unsigned FULL = 0xffffffffu;
  const int NP = A.NP, RC = A.RC, NG = P.n_groups, NGP = A.NGP, D = P.n_desc, NT = P.n_two;
  constexpr int SJ = 8, SR = 4;
  const int S = P.n_species;

  double* XYZ = reinterpret_cast<double*>(smem_raw);  // [3][NP]
  double* TJ = XYZ + 3 * NP;                          // [NP][SJ]: r r2 1/r fc dfc ux uy uz x_g..
  double* RK = TJ + NP * SJ;                          // [RC][SR]: r 1/r fc dfc x_g..
  double* SC = RK + RC * SR;                          // [32][SCW]
  double* Z2 = SC + SC_DOUBLES;                       // [NT] radial zeta, [3 NT] radial centre slot
  double* RCI = Z2 + 4 * NT;                          // [S][S] 1 / rc
  double* TAB = RCI + S * S;                          // [64] 2^(j/64) for kb_exp_tab
  unsigned long long* ADJ = reinterpret_cast<unsigned long long*>(TAB + 64);  // [NP]
  int* SPC = reinterpret_cast<int*>(ADJ + NP);        // [NP]
  int* BASE = SPC + NP;                               // [NP] first record slot of row j (pairs j<k)
  unsigned short* JR = reinterpret_cast<unsigned short*>(BASE + NP);   // [NP][NP] record of (slot, idx)
  unsigned short* PAIR = JR + NP * NP;                // [RC]
  unsigned char* JL = reinterpret_cast<unsigned char*>(PAIR + RC);     // [NP][NP] idx-th neighbor of slot
  unsigned char* DEG = JL + NP * NP;                  // [NP]
  unsigned char* KORD = DEG + NP;                     // [NP]

  // ---- per-lane constants of the triple loop (the warp is persistent)
  const int KPW = 32 / NGP;
  const int ksub = lane / NGP, g = lane & (NGP - 1);
  const bool valid_g = g < NG;
  double eta = 0.0;
  int gout[GW];
#pragma unroll
  for (int e = 0; e < GW; e++) gout[e] = -1;
  if (valid_g)
  {
    eta = P.grp_eta[g];
#pragma unroll
    for (int e = 0; e < GW; e++) gout[e] = P.grp_out[g][e];
  }
  const double p2tab[GW] = {1.0, 1.0, 0.5, 0.5, 0.125, 0.125, 1.0 / 32768.0, 1.0 / 32768.0};
  for (int q = lane; q < S * S; q += 32) RCI[q] = 1.0 / P.rcut[q];
  kb_exp_table_fill(TAB, lane);
  for (int q = lane; q < SC_DOUBLES; q += 32) SC[q] = 0.0;  // empty scratch entries must be finite
  __syncwarp();

  for (;;)
  {
    // ---- dynamic work distribution: one atom per grab
    int t = 0;
    if (lane == 0) t = atomicAdd(A.overflow + 1, 1);
    t = __shfl_sync(FULL, t, 0);
    if (t >= A.n_centers) break;
    if (A.subset) t = A.subset[t];
    const int i = A.centers ? A.centers[t] : t;
    const int64_t off = A.nl_off[i];
    const int n = (int) (A.nl_off[i + 1] - off);
    if (n > NP)
    {
      if (lane == 0) A.overflow[2 + atomicAdd(A.overflow, 1)] = t;
      continue;
    }
    const int si = A.species[i];
    const double xi = A.coords[3 * i], yi = A.coords[3 * i + 1], zi = A.coords[3 * i + 2];
    const int row = 3 * (n + 1);
    TOut* blk = GRAD ? static_cast<TOut*>(A.dz) + A.dz_ptr[t] : nullptr;
    double* zrow = A.zeta + (int64_t) t * D;

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
        const double ir = rsqrt(r2x);   // independent of the sqrt chain
        kb_cutoff_fast(r, rc, RCI[si * S + sj], fc, dfc);
        if (!act) { fc = 0.0; dfc = 0.0; }
        ux = vx * ir; uy = vy * ir; uz = vz * ir;
        XYZ[jj] = xj; XYZ[NP + jj] = yj; XYZ[2 * NP + jj] = zj;
        SPC[jj] = sj;
        double* tj = TJ + jj * SJ;
        const double r2 = r2x;
        tj[0] = r; tj[1] = r2; tj[2] = ir; tj[3] = fc; tj[4] = dfc; tj[5] = ux; tj[6] = uy; tj[7] = uz;
      }
      ACT |= (unsigned long long) __ballot_sync(FULL, act) << j0;
      for (int q0 = 0; q0 < NT; q0 += 8)
      {
        double v[32];
#pragma unroll
        for (int u = 0; u < 8; u++)  // 8 independent radial sets per trip
        {
          const int q = q0 + u;
          double phi = 0.0, dphi = 0.0;
          if (q < NT && act) kb_radial(P.two_kind[q], P.two_p0[q], P.two_p1[q], r, fc, dfc, phi, dphi);
          const double px = dphi * ux, py = dphi * uy, pz = dphi * uz;  // :497
          if (GRAD && q < NT && jj < n)
          {
            TOut* o = blk + (int64_t) P.two_out[q] * row + 3 * jj;
            o[0] = (TOut) px; o[1] = (TOut) py; o[2] = (TOut) pz;
          }
          v[4 * u] = phi; v[4 * u + 1] = px; v[4 * u + 2] = py; v[4 * u + 3] = pz;
        }
        const double tot = warp_reduce32(v, lane);  // lane l: total of value (set l/4, component l%4)
        const int q = q0 + (lane >> 2), comp = lane & 3;
        if (q < NT)
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
        TOut* o = blk + (int64_t) P.two_out[q] * row + 3 * n;  // centre slot, :499-501
        o[0] = (TOut) -Z2[NT + 3 * q]; o[1] = (TOut) -Z2[NT + 3 * q + 1]; o[2] = (TOut) -Z2[NT + 3 * q + 2];
      }
    }

    // ---- phase B: contributing neighbor pairs.  Lane = row j computes its whole adjacency row in
    //      registers (no ballots, no atomics); the reference predicate !(sqrt(r2) > rc) (:729) is
    //      decided on r2 whenever that is unambiguous and by the square root otherwise.
    int nrec = 0;
    for (int j0 = 0; j0 < n; j0 += 32)
    {
      const int j = j0 + lane;
      unsigned long long mask = 0ull;
      if (j < n && ((ACT >> j) & 1ull))
      {
        const double xj = XYZ[j], yj = XYZ[NP + j], zj = XYZ[2 * NP + j];
        const double* rcrow = P.rcut + SPC[j] * S;
        for (int k = 0; k < n; k++)
        {
          const double r2 = xnorm2(xsub(XYZ[k], xj), xsub(XYZ[NP + k], yj), xsub(XYZ[2 * NP + k], zj));
          const double rc = rcrow[SPC[k]];
          const double rc2 = rc * rc;
          bool hit = r2 < rc2 * (1.0 - 1e-12);
          if (!hit && r2 <= rc2 * (1.0 + 1e-12)) hit = !(sqrt(r2) > rc);
          if (hit && k != j && ((ACT >> k) & 1ull)) mask |= 1ull << k;
        }
      }
      const int upper = (j < 63) ? __popcll(mask >> (j + 1)) : 0;
      int inc = upper;  // inclusive warp scan of the per-row record counts
#pragma unroll
      for (int d = 1; d < 32; d <<= 1)
      {
        const int y = __shfl_up_sync(FULL, inc, d);
        if (lane >= d) inc += y;
      }
      if (j < n)
      {
        ADJ[j] = mask;
        DEG[j] = (unsigned char) __popcll(mask);
        BASE[j] = nrec + inc - upper;
      }
      nrec += __shfl_sync(FULL, inc, 31);
    }
    if (nrec > RC)
    {
      if (lane == 0) A.overflow[2 + atomicAdd(A.overflow, 1)] = t;
      continue;
    }
    __syncwarp();
    // per-slot neighbor lists (ascending) with their record ids, and the (j<k) pair list
    for (int k = lane; k < n; k += 32)
    {
      unsigned long long m = ADJ[k];
      int idx = 0;
      while (m)
      {
        const int j = __ffsll((long long) m) - 1;
        m &= m - 1ull;
        const int a = j < k ? j : k, b = j < k ? k : j;
        const unsigned long long between = (ADJ[a] >> (a + 1)) & ((1ull << (b - a - 1)) - 1ull);
        const int slot = BASE[a] + __popcll(between);
        JL[k * NP + idx] = (unsigned char) j;
        JR[k * NP + idx] = (unsigned short) slot;
        if (j > k) PAIR[slot] = (unsigned short) (k | (j << 8));
        idx++;
      }
    }
    __syncwarp();

    // ---- phase C: pair-jk records; lanes = records, two per trip so that the two dependent chains
    //      (sqrt -> 1/r -> sincos -> 7 exp) overlap --------------------------------------------------
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
      double* rka = RK + r0 * SR;
      double* rkb = RK + (two ? r1 : r0) * SR;
      rka[0] = ra; rka[1] = ira; rka[2] = fca; rka[3] = dfca;
      if (two) { rkb[0] = rb; rkb[1] = irb; rkb[2] = fcb; rkb[3] = dfcb; }
    }

    // ---- phase S: slots ordered by adjacency degree (descending, index as tie-break) ---------
    for (int k = lane; k < n; k += 32)
    {
      const int dk = DEG[k];
      int rank = 0;
      for (int m = 0; m < n; m++) rank += (DEG[m] > dk) || (DEG[m] == dk && m < k);
      KORD[rank] = (unsigned char) k;
    }
    __syncwarp();

    // ---- phase E: triple loop -------------------------------------------------------------------
    // csum: centre-slot partial sums; phi: zeta partial sums (both carried across the atom's items)
    double csum[GW][3], phi[GW];
#pragma unroll
    for (int e = 0; e < GW; e++) { csum[e][0] = csum[e][1] = csum[e][2] = 0.0; phi[e] = 0.0; }

    // one sweep step: every eta-lane folds scratch entry `sc` of its slot into its accumulators.
    // Entries beyond a slot's list are all-zero (F = 0, vectors = 0), so the step needs no predicate.
    double acc[GW][3];
    auto sweep_step = [&](const double* sc) {
      const double F = sc[1];
      const double E = kb_exp_tab(-eta * sc[11], TAB);  // e^{-eta (rij^2 + rik^2 + rjk^2)}, :770
      const double eF = E * F, Ee = E * eta;
      const double v1x = eF * sc[2], v1y = eF * sc[3], v1z = eF * sc[4];
      const double v2x = fma(Ee, sc[5], E * sc[8]), v2y = fma(Ee, sc[6], E * sc[9]),
                   v2z = fma(Ee, sc[7], E * sc[10]);
#pragma unroll
      for (int e = 0; e < GW; e += 2)
      {
        const double2 c2 = *reinterpret_cast<const double2*>(sc + 12 + e);
        const double2 d2 = *reinterpret_cast<const double2*>(sc + 20 + e);
        phi[e] = fma(c2.x, eF, phi[e]);
        phi[e + 1] = fma(c2.y, eF, phi[e + 1]);
        if (GRAD)
        {
          acc[e][0] = fma(c2.x, v2x, fma(d2.x, v1x, acc[e][0]));
          acc[e][1] = fma(c2.x, v2y, fma(d2.x, v1y, acc[e][1]));
          acc[e][2] = fma(c2.x, v2z, fma(d2.x, v1z, acc[e][2]));
          acc[e + 1][0] = fma(c2.y, v2x, fma(d2.y, v1x, acc[e + 1][0]));
          acc[e + 1][1] = fma(c2.y, v2y, fma(d2.y, v1y, acc[e + 1][1]));
          acc[e + 1][2] = fma(c2.y, v2z, fma(d2.y, v1z, acc[e + 1][2]));
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
        // geometry of up to NGP ordered pairs (j -> k) per slot: lane g takes list entry s0 + g.
        // The slot's own constants are re-read from shared memory here (not kept in registers
        // across the sweep below, where every register is needed for accumulators).
        double* sc = scg + g * SCW;
        if (s0 + g < degk)
        {
          const double* tk = TJ + k * SJ;
          const double rik = tk[0], rik2 = tk[1], irik = tk[2], fik = tk[3], dfik = tk[4];
          const double ukx = tk[5], uky = tk[6], ukz = tk[7];
          const int j = JL[k * NP + s0 + g];
          const int rec = JR[k * NP + s0 + g];
          const double* tj = TJ + j * SJ;
          const double* rk = RK + rec * SR;
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
          sc[1] = F;
          sc[2] = dct_ik * ukx + dct_jk * wx; sc[3] = dct_ik * uky + dct_jk * wy; sc[4] = dct_ik * ukz + dct_jk * wz;
          sc[5] = a * ukx + b * wx;           sc[6] = a * uky + b * wy;           sc[7] = a * ukz + b * wz;
          sc[8] = dF_ik * ukx + dF_jk * wx;   sc[9] = dF_ik * uky + dF_jk * wy;   sc[10] = dF_ik * ukz + dF_jk * wz;
          sc[11] = rij2 + rik2 + rjk2;
          // (1 + lambda cos)^zeta and its derivative factor for zeta in {1,2,4,16}, lambda = -1, +1:
          // entries e = 2*zi + si (sym_fn.cpp:749-767, multiplication chain instead of pow)
#pragma unroll
          for (int si2 = 0; si2 < 2; si2++)
          {
            const double lam = si2 ? 1.0 : -1.0;
            double bse = fma(lam, ct, 1.0);
            const bool ok = bse > 0.0;  // guard of :755-767
            bse = ok ? bse : 0.0;
            const double b2 = bse * bse, b3 = b2 * bse, b4 = b2 * b2, b8 = b4 * b4;
            const double b16 = b8 * b8, b15 = b8 * b4 * b3;
            sc[12 + 0 + si2] = bse; sc[20 + 0 + si2] = ok ? lam : 0.0;
            sc[12 + 2 + si2] = b2;  sc[20 + 2 + si2] = 2.0 * lam * bse;
            sc[12 + 4 + si2] = b4;  sc[20 + 4 + si2] = 4.0 * lam * b3;
            sc[12 + 6 + si2] = b16; sc[20 + 6 + si2] = 16.0 * lam * b15;
          }
        }
        else
        {
          // empty entry: contributes exactly zero (its C / C' words are finite leftovers or the
          // zeros the scratch was initialised with)
#pragma unroll
          for (int q = 1; q < 12; q++) sc[q] = 0.0;
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

      // item done: write slot k's rows, fold into the centre-slot partial sums
      if (GRAD && valid_k && valid_g)
      {
#pragma unroll
        for (int e = 0; e < GW; e++)
        {
          csum[e][0] += acc[e][0]; csum[e][1] += acc[e][1]; csum[e][2] += acc[e][2];
          if (gout[e] >= 0)
          {
            TOut* o = blk + (unsigned) (gout[e] * row + 3 * k);
            o[0] = (TOut) (p2tab[e] * acc[e][0]); o[1] = (TOut) (p2tab[e] * acc[e][1]);
            o[2] = (TOut) (p2tab[e] * acc[e][2]);
          }
        }
      }
    }

    // ---- phase F: centre slot (translation invariance) and zeta, reduced over the slot lanes ----
#pragma unroll
    for (int e = 0; e < GW; e++)
    {
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
      if (ksub == 0 && gout[e] >= 0)
      {
        zrow[gout[e]] = 0.5 * p2tab[e] * phi[e];  // each unordered triple was visited twice
        if (GRAD)
        {
          TOut* o = blk + (unsigned) (gout[e] * row + 3 * n);
          o[0] = (TOut) (-p2tab[e] * csum[e][0]); o[1] = (TOut) (-p2tab[e] * csum[e][1]);
          o[2] = (TOut) (-p2tab[e] * csum[e][2]);
        }
      }
    }
    if (GRAD)
    {
      const int len = (int) (A.dz_ptr[t + 1] - A.dz_ptr[t]);
      for (int q = D * row + lane; q < len; q += 32) blk[q] = (TOut) 0;  // alignment padding
    }
    __syncwarp();
  }
}

}  // namespace

// Returns KB_OK and *launched = true when the parameter set is one this kernel handles.
int kb_symfun_warp_launch(const kb_symfun_params& P, int n_centers, const int32_t* d_centers,
                          const int32_t* d_subset, const double* d_coords, const int32_t* d_species,
                          const int64_t* d_nl_off, const int32_t* d_nl_idx, int maxn, double* d_zeta,
                          void* d_dz, int dz_dtype, const int64_t* d_dz_ptr, int32_t* d_overflow,
                          int full_capacity, bool* launched, bool* may_overflow, cudaStream_t st)
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
  const int worst = NP > 1 ? NP * (NP - 1) / 2 : 1;
  const size_t limit = 227 * 1024, per_block_reserve = 1024;
  // Record capacity.  Pairs of neighbors that are within the cutoff of each other are ~45 % of all
  // pairs for a filled sphere (diamond Si at 5 A: 168 of 378).  Take the largest capacity that still
  // lets the most warps be resident per SM, but never below 45 % of the worst case unless asked for
  // the worst case itself; atoms that exceed it are redone by the caller with full capacity.
  int RC = worst;
  if (!full_capacity)
  {
    const size_t fixed = warp_smem_bytes(P, NP, 0);
    const size_t per_rec = (size_t) 4 * 8 + 2;
    const int floor_rc = (worst * 9 + 19) / 20;
    RC = 0;
    const int want_blocks = d_dz != nullptr ? KB_WARP_MIN_BLOCKS : KB_WARP_MIN_BLOCKS_NOGRAD;
    for (int target = want_blocks; target >= 1 && RC == 0; target--)
    {
      const size_t budget = limit / target;
      if (budget <= fixed + per_block_reserve + 32) continue;
      long long cap = (long long) ((budget - fixed - per_block_reserve - 32) / per_rec);
      if (cap >= floor_rc) RC = cap > worst ? worst : (int) cap;
    }
    if (RC == 0) RC = floor_rc;
  }
  while (RC > 8 && warp_smem_bytes(P, NP, RC) > limit) RC = RC * 3 / 4;
  if (RC < worst) *may_overflow = true;
  const size_t smem = warp_smem_bytes(P, NP, RC);
  if (smem > limit) return KB_OK;

  WarpArgs A;
  A.n_centers = n_centers; A.centers = d_centers; A.subset = d_subset; A.coords = d_coords;
  A.species = d_species; A.nl_off = d_nl_off; A.nl_idx = d_nl_idx; A.zeta = d_zeta; A.dz = d_dz;
  A.dz_ptr = d_dz_ptr; A.overflow = d_overflow; A.NP = NP; A.RC = RC; A.NGP = NGP;

  int dev = 0, sms = 148;
  KB_CUDA(cudaGetDevice(&dev));
  KB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  int per_sm = (int) (limit / (smem + per_block_reserve));
  const int max_blocks = d_dz != nullptr ? KB_WARP_MIN_BLOCKS : KB_WARP_MIN_BLOCKS_NOGRAD;
  if (per_sm > max_blocks) per_sm = max_blocks;
  if (const char* cap = getenv("KB_WARPS_PER_SM"))  // tuning experiments only
  {
    const int c = atoi(cap);
    if (c >= 1 && c < per_sm) per_sm = c;
  }
  if (per_sm < 1) per_sm = 1;
  int grid = sms * per_sm;
  if (grid > n_centers) grid = n_centers;
  const bool grad = d_dz != nullptr;
#define KB_WARP_LAUNCH(G, T)                                                                        \
  do {                                                                                              \
    KB_CUDA(cudaFuncSetAttribute(symfun_warp_kernel<G, T>,                                          \
                                 cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));         \
    kb_timing_begin(KB_T_FUSED, st);                                                                \
    symfun_warp_kernel<G, T><<<grid, 32, smem, st>>>(P, A);                                         \
    kb_timing_end(KB_T_FUSED, st);                                                                  \
  } while (0)
  if (!grad) KB_WARP_LAUNCH(false, double);
  else if (dz_dtype == KB_F32) KB_WARP_LAUNCH(true, float);
  else KB_WARP_LAUNCH(true, double);
#undef KB_WARP_LAUNCH
  KB_LAUNCH_CHECK();
  *launched = true;
  return KB_OK;
}
