// Round-2 triple sweep (stage 2 of the two-stage path): the arithmetic of sweep_atom (symfun_atom.cuh)
// with a register diet, so that 12 warps instead of 9 are resident per SM.
//
// ncu on the round-1 sweep (profiles/r01_ncu_symfun_sweep_v3.txt and the source view behind it): the FMA
// loop holds 70 % of the instructions and 57 % of the warp time and keeps the FP64 pipe ~84 % busy while
// a warp is inside it; the other 43 % of the warp time (geometry rounds, row stores, window setup, the
// closing reductions) are dependent chains that use the pipe at ~15 %, and with 2.25 warps per scheduler
// nothing covers them — 53 % pipe activity overall.  The limiter of residency was the register file:
// 216 registers, 112 of them accumulators (24 row sums, 24 centre-slot sums, 8 zeta sums, all doubles).
//
// Here a lane keeps only the 32 per-window accumulators (8 entries x {x, y, z, phi}).  At every window
// flush the four slot-lanes of an eta-group reduce-scatter them with 48 shuffles: each lane ends up owning
// two entries of its group and adds the window's total to 8 running sums — the centre slot (translation
// invariance, slot_i = -sum_k slot_k) and zeta — instead of 32.
// 168 registers -> CTAs of four warps (own atom each, one shared exponential table), three CTAs per SM.
//
// Reference: Descriptor::generate_one_atom (sym_fn.cpp:416-602), sym_d_g4 (:744-807).
#pragma once
#include <type_traits>

#include "symfun_atom.cuh"

namespace {

#ifndef KB_SWEEP3_REGS
#define KB_SWEEP3_REGS 168
#endif
constexpr int S3_WARPS = 4;                                 // warps per CTA
constexpr int S3_CTAS = 65536 / (S3_WARPS * 32 * KB_SWEEP3_REGS);  // CTAs per SM the register file allows

template <typename TOut>
__device__ __forceinline__ void sweep3_atom(const kb_symfun_params& P, const SplitArgs& A, const SweepSmem& W,
                                            int t, int n, int rstride, int64_t dz_lo, int64_t dz_hi, int lane)
{
  const unsigned FULL = 0xffffffffu;
  constexpr int NGP = 8, KPW = 4;
  const int NP = A.NP, D = P.n_desc;
  const unsigned char *KORD = W.KORD, *DEG = W.DEG, *JL = W.JL;
  const unsigned short* JR = W.JR;
  const double *XYZ = W.XYZ, *TJ = W.TJ, *RK = W.RK;
  double *SC = W.SC;
  const double* TAB = W.TAB;
  const int ksub = lane >> 3, g = lane & 7;
  const bool valid_g = g < P.n_groups;
  const int gc = valid_g ? g : 0;
  const double eta = valid_g ? P.grp_eta[gc] : 0.0;
  const int row = 3 * (n + 1);
  TOut* blk = static_cast<TOut*>(A.dz) + dz_lo;
  double* zrow = A.zeta + (int64_t) t * D;

  if (A.ang_hi > A.ang_lo)
  {
    // (KB_PREZERO=1) Zero the angular rows with full, coalesced sector writes first: the row pieces written at
    // the flushes are 24 bytes at 8-byte alignment, and a partial write to a sector that is not in L2
    // makes L2 fetch it from DRAM; after this pass every piece lands on a resident dirty line.
    // (L2 cache hints — evict_last on the row stores, evict_first on the record copy — change nothing:
    // 45.4 instead of 45.6 KB written, 29.7 instead of 31.5 KB read per atom, same time.)
    TOut* z0 = blk + (unsigned) (A.ang_lo * row);
    const int len = (A.ang_hi - A.ang_lo) * row;
    for (int q = lane; q < len; q += 32) z0[q] = (TOut) 0;
  }

  double csum[2][4];  // running sums over all slots of entry (4 h + ksub): x, y, z, phi
#pragma unroll
  for (int h = 0; h < 2; h++) csum[h][0] = csum[h][1] = csum[h][2] = csum[h][3] = 0.0;

  double acc[GW][4];  // per window: x, y, z of the slot's row pieces; [3] = zeta sums, running over the atom
#pragma unroll
  for (int e = 0; e < GW; e++) acc[e][3] = 0.0;
#ifdef KB_S3_LEGACY
  auto sweep_step = [&](const double* sc) {
    const double2 sF = *reinterpret_cast<const double2*>(sc);  // {rij^2 + rik^2 + rjk^2, F}
#ifdef KB_S3_EXP64
    const double E = kb_exp_tab(-eta * sF.x, TAB);
#else
    const double E = kb_exp_tab16(-eta * sF.x, TAB);
#endif
    // e^{-eta (rij^2 + rik^2 + rjk^2)}, :770
    const double eF = E * sF.y, Ee = E * eta;
    const double v1x = eF * sc[2], v1y = eF * sc[3], v1z = eF * sc[4];
    const double v2x = fma(Ee, sc[5], E * sc[8]), v2y = fma(Ee, sc[6], E * sc[9]),
                 v2z = fma(Ee, sc[7], E * sc[10]);
    // entries e = 2 zi + si, zeta = {1,2,4,16}[zi], lambda = {-1,+1}[si]:  C = b^zeta,
    // C' = zeta lambda b^(zeta-1) (0 when b = 0: the guard of sym_fn.cpp:755-767)
    const double2 c0 = *reinterpret_cast<const double2*>(sc + 12);
    const double2 c1 = *reinterpret_cast<const double2*>(sc + 14);
    const double2 c2 = *reinterpret_cast<const double2*>(sc + 16);
    const double2 c3 = *reinterpret_cast<const double2*>(sc + 18);
    const double2 d3 = *reinterpret_cast<const double2*>(sc + 20);
    // the scratch holds the powers already scaled by 2^(1-zeta) (exact), so C' of zeta = 2 is -+C of
    // zeta = 1 and C' of zeta = 4 is -+ C(1) C(2): two multiplications, and the flush needs no scaling
    double2 d0, d1, d2;
    d0.x = __hiloint2double(__double2hiint(c0.x) > 0 ? 0xbff00000 : 0, 0);  // b > 0 ? -1 : 0
    d0.y = __hiloint2double(__double2hiint(c0.y) > 0 ? 0x3ff00000 : 0, 0);  // b > 0 ? +1 : 0
    d1.x = -c0.x;           d1.y = c0.y;
    d2.x = -(c0.x * c1.x);  d2.y = c0.y * c1.y;
    const double cc[GW] = {c0.x, c0.y, c1.x, c1.y, c2.x, c2.y, c3.x, c3.y};
    const double dd[GW] = {d0.x, d0.y, d1.x, d1.y, d2.x, d2.y, d3.x, d3.y};
#pragma unroll
    for (int e = 0; e < GW; e++)
    {
      acc[e][0] = fma(cc[e], v2x, fma(dd[e], v1x, acc[e][0]));
      acc[e][1] = fma(cc[e], v2y, fma(dd[e], v1y, acc[e][1]));
      acc[e][2] = fma(cc[e], v2z, fma(dd[e], v1z, acc[e][2]));
      acc[e][3] = fma(cc[e], eF, acc[e][3]);
    }
  };

#else
  // Round-2 (late) form of the step.  ncu source view of the form above (gpurun_out/s6d_sweep.ncu-rep): 28 % of the
  // loop's warp time are fixed-latency waits on the exponential chain of the trip's FIRST step — nothing of
  // that trip is ready to run beside it — so the exponential is taken out of the step and software-pipelined:
  // step_E of step i + 1 is issued before the FMAs of step i, across trips too (one carried double).
  // Fewer FP64 instructions on the way (84 -> 77 per step; alone worth < 1 %: the loop is latency-, not
  // throughput-bound): exponential in 11 (kb_exp_tab16s); v2 = E (eta A + B) with the eta-combination formed
  // beside the chain; the two zeta = 1 entries have C' = -+1, so their C' v1 terms are ONE running sum sv
  // (3 additions instead of 6 FMAs) that the flush subtracts from entry 0 and adds to entry 1.  The guard of
  // sym_fn.cpp:755-767 (C' = 0 where 1 +- cos = 0: exactly collinear triples) is decided per scratch chunk
  // in the geometry round (warp vote): a chunk holding such an entry runs the exact form of the step.
  double sv[3] = {0.0, 0.0, 0.0};
  auto step_E = [&](const double* sc) { return kb_exp_tab16s(-eta * sc[0], TAB); };  // e^{-eta (rij^2+rik^2+rjk^2)}, :770
  auto step_fma = [&](const double* sc, const double E, auto exact_tag) {
    constexpr bool EXACT = decltype(exact_tag)::value;
    const double eF = E * sc[1];
    const double v1x = eF * sc[2], v1y = eF * sc[3], v1z = eF * sc[4];
    const double v2x = E * fma(eta, sc[5], sc[8]), v2y = E * fma(eta, sc[6], sc[9]),
                 v2z = E * fma(eta, sc[7], sc[10]);
    const double2 c0 = *reinterpret_cast<const double2*>(sc + 12);
    const double2 c1 = *reinterpret_cast<const double2*>(sc + 14);
    const double2 c2 = *reinterpret_cast<const double2*>(sc + 16);
    const double2 c3 = *reinterpret_cast<const double2*>(sc + 18);
    const double2 d3 = *reinterpret_cast<const double2*>(sc + 20);
    double2 d0, d1, d2;
    d1.x = -c0.x;           d1.y = c0.y;
    d2.x = -(c0.x * c1.x);  d2.y = c0.y * c1.y;
    if (EXACT)
    {
      // C' of zeta = 1 as data: -+1, or 0 where the base vanished; sv gets nothing from this step's entries 0, 1
      d0.x = __hiloint2double(__double2hiint(c0.x) > 0 ? 0xbff00000 : 0, 0);
      d0.y = __hiloint2double(__double2hiint(c0.y) > 0 ? 0x3ff00000 : 0, 0);
    }
    else
    {
      d0.x = d0.y = 0.0;
      sv[0] += v1x; sv[1] += v1y; sv[2] += v1z;
    }
    const double cc[GW] = {c0.x, c0.y, c1.x, c1.y, c2.x, c2.y, c3.x, c3.y};
    const double dd[GW] = {d0.x, d0.y, d1.x, d1.y, d2.x, d2.y, d3.x, d3.y};
#pragma unroll
    for (int e = 0; e < GW; e++)
    {
      if (EXACT || e >= 2)
      {
        acc[e][0] = fma(cc[e], v2x, fma(dd[e], v1x, acc[e][0]));
        acc[e][1] = fma(cc[e], v2y, fma(dd[e], v1y, acc[e][1]));
        acc[e][2] = fma(cc[e], v2z, fma(dd[e], v1z, acc[e][2]));
      }
      else
      {
        acc[e][0] = fma(cc[e], v2x, acc[e][0]);
        acc[e][1] = fma(cc[e], v2y, acc[e][1]);
        acc[e][2] = fma(cc[e], v2z, acc[e][2]);
      }
      acc[e][3] = fma(cc[e], eF, acc[e][3]);
    }
  };
  auto sweep_chunk = [&](const double* scg, const int maxcnt, auto exact_tag) {
    int ej = 0;
#ifdef KB_S3_PIPE
    // exponential of step i + 1 issued before the FMAs of step i, across trips (one carried double): measured
    // 25.52 ms against 25.24 ms for the plain two-step trip below (1 M atoms) — the scheduler already overlaps
    // the two steps of a trip, the carried value only lengthens live ranges
    double E0 = step_E(scg);
#pragma unroll 1
    for (; ej + 1 < maxcnt; ej += 2)
    {
      const double E1 = step_E(scg + (ej + 1) * SCW);
      step_fma(scg + ej * SCW, E0, exact_tag);
      const int nx = ej + 2 < maxcnt ? ej + 2 : ej + 1;  // past the end: any valid entry, the value is not used
      E0 = step_E(scg + nx * SCW);
      step_fma(scg + (ej + 1) * SCW, E1, exact_tag);
    }
    if (ej < maxcnt) step_fma(scg + ej * SCW, E0, exact_tag);
#else
#pragma unroll 1
    for (; ej + 1 < maxcnt; ej += 2)
    {
      step_fma(scg + ej * SCW, step_E(scg + ej * SCW), exact_tag);
      step_fma(scg + (ej + 1) * SCW, step_E(scg + (ej + 1) * SCW), exact_tag);
    }
    if (ej < maxcnt) step_fma(scg + ej * SCW, step_E(scg + ej * SCW), exact_tag);
#endif
  };
#endif

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
      bool base_zero = false;
      if (s0 + g < degk)
      {
        // field-major tables (prep_atom<.., SOA = true>): lanes gather one field of different neighbors
        const double rik = TJ[k], rik2 = TJ[NP + k], irik = TJ[2 * NP + k], fik = TJ[3 * NP + k];
        const double dfik = TJ[4 * NP + k], ukx = TJ[5 * NP + k], uky = TJ[6 * NP + k], ukz = TJ[7 * NP + k];
        const int j = JL[k * NP + s0 + g];
        const int rcd = JR[k * NP + s0 + g];
        const double rij2 = TJ[NP + j], irij = TJ[2 * NP + j], fij = TJ[3 * NP + j];
        const double rjk = RK[rcd], irjk = RK[rstride + rcd], fjk = RK[2 * rstride + rcd],
                     dfjk = RK[3 * rstride + rcd];
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
        // 16-byte stores: the eight lanes of a slot start on eight different 16-byte bank groups (SCW = 22)
        double2* s2 = reinterpret_cast<double2*>(sc);
        s2[0] = make_double2(rij2 + rik2 + rjk2, F);
        s2[1] = make_double2(dct_ik * ukx + dct_jk * wx, dct_ik * uky + dct_jk * wy);
        s2[2] = make_double2(dct_ik * ukz + dct_jk * wz, a * ukx + b * wx);
        s2[3] = make_double2(a * uky + b * wy, a * ukz + b * wz);
        s2[4] = make_double2(dF_ik * ukx + dF_jk * wx, dF_ik * uky + dF_jk * wy);
        s2[5] = make_double2(dF_ik * ukz + dF_jk * wz, 0.0);
        double pw[2][5];
#pragma unroll
        for (int si2 = 0; si2 < 2; si2++)  // (1 +- cos)^{1,2,4,16} and 16 lambda (1 +- cos)^15, :749-767
        {
          const double lam = si2 ? 1.0 : -1.0;
          double bse = fma(lam, ct, 1.0);
          base_zero |= !(bse > 0.0);
          bse = bse > 0.0 ? bse : 0.0;  // guard of :755-767
          const double b2 = bse * bse, b3 = b2 * bse, b4 = b2 * b2, b8 = b4 * b4;
          const double b16 = b8 * b8, b15 = b8 * b4 * b3;
          pw[si2][0] = bse; pw[si2][1] = 0.5 * b2; pw[si2][2] = 0.125 * b4; pw[si2][3] = (1.0 / 32768.0) * b16;
          pw[si2][4] = (16.0 / 32768.0) * lam * b15;  // all scaled by 2^(1-zeta), sym_fn.cpp:772
        }
#pragma unroll
        for (int q = 0; q < 5; q++) s2[6 + q] = make_double2(pw[0][q], pw[1][q]);
      }
      else
      {
        double2* s2 = reinterpret_cast<double2*>(sc);
#pragma unroll
        for (int q = 0; q < 6; q++) s2[q] = make_double2(0.0, 0.0);  // empty entry: contributes exactly zero
      }
      const int left = maxdeg - s0;
      const int maxcnt = left < NGP ? left : NGP;  // warp-uniform
      const bool chunk_guard = __any_sync(FULL, base_zero);
      __syncwarp();
      (void) chunk_guard;

#ifdef KB_S3_LEGACY
      int ej = 0;
#pragma unroll 1
      for (; ej + 1 < maxcnt; ej += 2)
      {
        sweep_step(scg + ej * SCW);
        sweep_step(scg + (ej + 1) * SCW);
      }
      if (ej < maxcnt) sweep_step(scg + ej * SCW);
#else
      if (chunk_guard) sweep_chunk(scg, maxcnt, std::true_type());
      else sweep_chunk(scg, maxcnt, std::false_type());
#endif
      __syncwarp();
    }

    // ---- flush: rows of slot k, then the transposed reduction over the four slots of the window
#ifndef KB_S3_LEGACY
#pragma unroll
    for (int c = 0; c < 3; c++)
    {
      acc[0][c] -= sv[c];
      acc[1][c] += sv[c];
      sv[c] = 0.0;
    }
#endif
    if (valid_k && valid_g)
    {
      TOut* o[GW];
      int gout[GW];
#pragma unroll
      for (int e = 0; e < GW; e++)
      {
        gout[e] = P.grp_out[gc][e];
        o[e] = blk + (unsigned) ((gout[e] < 0 ? 0 : gout[e]) * row + 3 * k);
        asm volatile("" : "+l"(o[e]));  // distinct address registers: the stores issue back to back
        __builtin_assume(__isGlobal(o[e]));
      }
#pragma unroll
      for (int e = 0; e < GW; e++)
        if (gout[e] >= 0)
        {
          o[e][0] = (TOut) acc[e][0]; o[e][1] = (TOut) acc[e][1]; o[e][2] = (TOut) acc[e][2];
        }
    }
    {
      // reduce-scatter over the four slot-lanes (lane bits 3, 4): after two exchange rounds lane
      // (ksub, g) holds, summed over the window's slots, entries 4 b0 + 2 b1 + {0, 1} (b0 = bit 3, b1 = bit 4)
      const bool b0 = (lane & 8) != 0, b1 = (lane & 16) != 0;
      double r[4][3];
#pragma unroll
      for (int i = 0; i < 4; i++)
#pragma unroll
        for (int c = 0; c < 3; c++)
        {
          const double keep = b0 ? acc[4 + i][c] : acc[i][c];
          const double send = b0 ? acc[i][c] : acc[4 + i][c];
          r[i][c] = keep + shfl_xor_d(send, 8);
        }
#pragma unroll
      for (int m = 0; m < 2; m++)
#pragma unroll
        for (int c = 0; c < 3; c++)
        {
          const double keep = b1 ? r[2 + m][c] : r[m][c];
          const double send = b1 ? r[m][c] : r[2 + m][c];
          csum[m][c] += keep + shfl_xor_d(send, 16);
        }
    }
  }
  {
    // zeta sums: the same reduce-scatter, once per atom
    const bool b0 = (lane & 8) != 0, b1 = (lane & 16) != 0;
    double r[4];
#pragma unroll
    for (int i = 0; i < 4; i++)
    {
      const double keep = b0 ? acc[4 + i][3] : acc[i][3];
      const double send = b0 ? acc[i][3] : acc[4 + i][3];
      r[i] = keep + shfl_xor_d(send, 8);
    }
#pragma unroll
    for (int m = 0; m < 2; m++)
    {
      const double keep = b1 ? r[2 + m] : r[m];
      const double send = b1 ? r[m] : r[2 + m];
      csum[m][3] = keep + shfl_xor_d(send, 16);
    }
  }

  // ---- centre slot and zeta: lane (ksub, g) owns entries 4 b0 + 2 b1 + {0, 1} of its group ----------------
  if (valid_g)
  {
    const int zi = ((lane >> 3) & 1) * 2 + ((lane >> 4) & 1);
#pragma unroll
    for (int m = 0; m < 2; m++)
    {
      const int out = P.grp_out[gc][2 * zi + m];
      if (out >= 0)
      {
        zrow[out] = 0.5 * csum[m][3];  // each unordered triple was visited twice
        TOut* o = blk + (unsigned) (out * row + 3 * n);
        o[0] = (TOut) -csum[m][0]; o[1] = (TOut) -csum[m][1]; o[2] = (TOut) -csum[m][2];
      }
    }
  }
  {
    const int len = (int) (dz_hi - dz_lo);
    for (int q = D * row + lane; q < len; q += 32) blk[q] = (TOut) 0;  // alignment padding
  }
}

template <typename TOut>
__global__ void __launch_bounds__(S3_WARPS * 32, S3_CTAS)
symfun_sweep3_kernel(const __grid_constant__ kb_symfun_params P, const SplitArgs A, const int warp_bytes)
{
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const unsigned FULL = 0xffffffffu;
  const RecLayout L = rec_layout(A.NP, A.RC);
  double* TAB = reinterpret_cast<double*>(smem_raw);  // [16] 2^(j/16), shared by the CTA's warps
  unsigned char* mine = smem_raw + 512 + (size_t) warp * warp_bytes;
  SweepSmem W = sweep_carve(mine, L);
  W.TAB = TAB;
#ifdef KB_S3_EXP64
  if (warp == 0) kb_exp_table_fill(TAB, lane);
#else
  if (warp == 0) kb_exp_table16_fill(TAB, lane);
#endif
  for (int q = lane; q < SC_DOUBLES; q += 32) W.SC[q] = 0.0;  // empty scratch entries must be finite
  __syncthreads();

  // Work distribution: one atom per grab from an atomic counter, taken TWO atoms ahead — the counter's
  // round trip, the next record's header and block offsets, and the L2 prefetch of the next record all
  // complete under the sweep of the current atom instead of being waited for.
  int grab = 0;
  if (lane == 0) grab = A.lo + atomicAdd(A.work, 1);
  int tt_next = __shfl_sync(FULL, grab, 0);
  if (lane == 0) grab = A.lo + atomicAdd(A.work, 1);
  int4 hdr_next = make_int4(-1, 0, 0, 0);
  int t_next = 0;
  int64_t lo_next = 0, hi_next = 0;
  if (tt_next < A.hi)
  {
    hdr_next = __ldcg(reinterpret_cast<const int4*>(A.records + (size_t) (tt_next - A.lo) * L.bytes));
    t_next = A.subset ? A.subset[tt_next] : tt_next;
    lo_next = A.dz_ptr[t_next]; hi_next = A.dz_ptr[t_next + 1];
  }
  for (;;)
  {
    const int tt = tt_next;
    if (tt >= A.hi) break;
    const int4 hdr = hdr_next;
    const int t = t_next;
    const int64_t dz_lo = lo_next, dz_hi = hi_next;
    tt_next = __shfl_sync(FULL, grab, 0);
    if (lane == 0) grab = A.lo + atomicAdd(A.work, 1);
    const unsigned char* rec = A.records + (size_t) (tt - A.lo) * L.bytes;
    const int n = hdr.x, rstride = (hdr.y + 1) & ~1;  // pair-jk records: [4][rstride]
    if (tt_next < A.hi)
    {
      const unsigned char* nx = A.records + (size_t) (tt_next - A.lo) * L.bytes;
      hdr_next = __ldcg(reinterpret_cast<const int4*>(nx));
      t_next = A.subset ? A.subset[tt_next] : tt_next;
      lo_next = A.dz_ptr[t_next]; hi_next = A.dz_ptr[t_next + 1];
      const int lines = (L.rk + 24 * A.RC + 127) >> 7;  // its pair-record count is not known yet: 3/4 of the capacity
      for (int q = lane; q < lines; q += 32)
        asm volatile("prefetch.global.L2 [%0];" ::"l"(nx + ((size_t) q << 7)));
    }
    if (n < 0) continue;  // declined by stage 1
    __syncwarp();
    {
      const int words = (L.rk + 32 * rstride) >> 4;
      const unsigned sbase = (unsigned) __cvta_generic_to_shared(mine);
      for (int q = lane; q < words; q += 32)
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sbase + 16u * q), "l"(rec + 16 * (size_t) q) : "memory");
      asm volatile("cp.async.commit_group;" ::: "memory");
      asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    __syncwarp();
    sweep3_atom<TOut>(P, A, W, t, n, rstride, dz_lo, dz_hi, lane);
  }
}

}  // namespace
