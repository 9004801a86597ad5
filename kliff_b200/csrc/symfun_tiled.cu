// Tiled symmetry-function + dG/dr kernel (fp64) for sm_100a — the hot kernel of the path.
//
// Replaces the per-atom call chain sym_fn.py:155-160 -> Descriptor::generate_one_atom
// (sym_fn.cpp:416-602) -> sym_d_g2/g4/g5 (:627-936).
//
// One CTA (8 warps) per centre atom i with n <= 64 neighbors.  Everything the O(n^2 * D) triple
// loop needs more than once is computed ONCE into shared memory:
//   pair-ij table  : r, 1/r, unit vector, fc, fc', exp(-eta_g r^2) for every angular group g
//   pair-jk records: the same for every pair of neighbors that can contribute (r_jk <= rc_jk, or
//                    all pairs when a g5 group is present), found with a ballot compaction, plus
//                    a 64-bit adjacency mask per neighbor.
// so the reference's per-(triple, parameter set) exp / pow / cos / sin / sqrt (sym_fn.cpp:744-792)
// become ~n*G + T4*G exponentials per ATOM, and the angular power (1 + lambda cos)^zeta for
// zeta in {1,2,4,16} is a multiplication chain.
//
// The triple loop is organised so that NO atomics and NO cross-thread reductions are needed
// for the gradient: work item = (neighbor slot k, angular group g).  The thread owning the item
// walks the neighbors j adjacent to k and accumulates ONLY the derivative w.r.t. slot k,
//     dG[d][k][:] += dphi/dr_ik * u_ik + dphi/dr_jk * u_jk ,
// for the <= 8 (zeta, lambda) entries of its group in registers (8x3 accumulators, a rank-2
// update per visited pair: acc[e][c] += C'_e * V1[c] + C_e * V2[c]).  Every unordered triple is
// therefore visited from both ends, which is exactly the work the reference does per triple
// (slots jj and kk, sym_fn.cpp:587-593) without any write conflict.  The centre-atom slot is
// obtained from translation invariance, slot(i) = -sum_k slot(k) (identical to the reference's
// separately accumulated value up to rounding, ~1e-16 of the row norm).
//
// The per-atom block [D][n+1][3] is assembled in shared memory in the reference's layout and
// leaves the SM as one contiguous, 16-byte vectorised, fully coalesced copy.
#include "common.cuh"
#include "symfun_math.cuh"

namespace {

constexpr int TT = 256;  // threads per CTA
constexpr int TW = TT / 32;
constexpr int GW = KB_GROUP_WIDTH;

struct TiledArgs
{
  int n_centers;
  const int32_t* centers;
  const double* coords;
  const int32_t* species;
  const int64_t* nl_off;
  const int32_t* nl_idx;
  double* zeta;
  void* dz;
  const int64_t* dz_ptr;
  int32_t* overflow;  // [0] = count, [1..] = centre ordinals to redo with the general kernel
  int NP;             // neighbor capacity (even, <= 64)
  int RC;             // jk-record capacity
  int NGP;            // power of two >= n_groups (lanes per neighbor slot)
  int NPW;            // warp-items capacity = ceil(NP / (32 / NGP))
  int stage_elems;    // doubles reserved for the output block (0 when dz == NULL)
  int has_g5;
};

__host__ __device__ inline size_t tiled_smem_bytes(const kb_symfun_params& P, int NP, int RC,
                                                   int NGP, int NPW, int stage_elems)
{
  const int NG = P.n_groups;
  size_t d = (size_t) stage_elems + (size_t) NP * (11 + NG + P.n_two) + (size_t) RC * (6 + NG)
             + (size_t) NPW * NGP * GW + (size_t) NG * (1 + 2 * GW);
  size_t bytes = d * 8;
  bytes += (size_t) 2 * NP * 8;                  // adjacency masks
  bytes += (size_t) (2 * NP + NG + NG * GW + 8) * 4;
  bytes += (size_t) (NP * NP + RC) * 2;
  return (bytes + 15) & ~(size_t) 15;
}

template <int ZLK, bool GRAD, typename TOut>
__global__ void __launch_bounds__(TT, 2)
symfun_tiled_kernel(const __grid_constant__ kb_symfun_params P, const TiledArgs A)
{
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int NP = A.NP, RC = A.RC, NG = P.n_groups, NGP = A.NGP, D = P.n_desc, NT = P.n_two;

  // ---- carve shared memory
  double* stage = reinterpret_cast<double*>(smem_raw);
  double* X = stage + A.stage_elems;  // [3][NP]
  double* Rr = X + 3 * NP;
  double* R2 = Rr + NP;
  double* IR = R2 + NP;
  double* U = IR + NP;                // [3][NP]
  double* FC = U + 3 * NP;
  double* DFC = FC + NP;
  double* XG = DFC + NP;              // [NG][NP]
  double* PHI2 = XG + NG * NP;        // [NT][NP]
  double* RR = PHI2 + NT * NP;        // records
  double* RU = RR + RC;               // [3][RC]
  double* RFC = RU + 3 * RC;
  double* RDFC = RFC + RC;
  double* RX = RDFC + RC;             // [NG][RC]
  double* PART = RX + NG * RC;        // [NPW][NGP*GW]
  double* GETA = PART + A.NPW * NGP * GW;
  double* GZETA = GETA + NG;          // [NG][GW]
  double* GLAM = GZETA + NG * GW;
  unsigned long long* ADJ4 = reinterpret_cast<unsigned long long*>(GLAM + NG * GW);
  unsigned long long* ADJA = ADJ4 + NP;
  int* NBR = reinterpret_cast<int*>(ADJA + NP);
  int* SPC = NBR + NP;
  int* GKIND = SPC + NP;
  int* GOUT = GKIND + NG;             // [NG][GW]
  int* CNT = GOUT + NG * GW;          // [0] nrec, [1..2] active mask words
  unsigned short* PIDX = reinterpret_cast<unsigned short*>(CNT + 8);
  unsigned short* PAIR = PIDX + NP * NP;

  const int t = blockIdx.x;
  const int i = A.centers ? A.centers[t] : t;
  const int64_t off = A.nl_off[i];
  const int n = (int) (A.nl_off[i + 1] - off);
  if (n > NP)
  {
    if (tid == 0) A.overflow[1 + atomicAdd(A.overflow, 1)] = t;
    return;
  }
  const int S = P.n_species;
  const int si = A.species[i];
  const double xi = A.coords[3 * i], yi = A.coords[3 * i + 1], zi = A.coords[3 * i + 2];
  const int row = 3 * (n + 1);

  // ---- phase A: pair-ij table -------------------------------------------------------------
  if (tid < 8) CNT[tid] = 0;
  for (int q = tid; q < NG; q += TT)
  {
    GETA[q] = P.grp_eta[q];
    GKIND[q] = P.grp_kind[q];
  }
  for (int q = tid; q < NG * GW; q += TT)
  {
    GOUT[q] = (q % GW) < P.grp_n[q / GW] ? P.grp_out[q / GW][q % GW] : -1;
    GZETA[q] = P.grp_zeta[q / GW][q % GW];
    GLAM[q] = P.grp_lambda[q / GW][q % GW];
  }
  __syncthreads();
  bool act = false;
  if (tid < n)
  {
    const int j = A.nl_idx[off + tid];
    const int sj = A.species[j];
    const double xj = A.coords[3 * j], yj = A.coords[3 * j + 1], zj = A.coords[3 * j + 2];
    const double vx = xsub(xj, xi), vy = xsub(yj, yi), vz = xsub(zj, zi);
    const double r = sqrt(xnorm2(vx, vy, vz));  // sym_fn.cpp:447-448
    const double rc = P.rcut[si * S + sj];
    act = !(r > rc);                            // :451
    const double ir = 1.0 / r;
    double fc = 0.0, dfc = 0.0;
    if (act) kb_cutoff(r, rc, fc, dfc);
    NBR[tid] = j; SPC[tid] = sj;
    X[tid] = xj; X[NP + tid] = yj; X[2 * NP + tid] = zj;
    Rr[tid] = r; R2[tid] = r * r; IR[tid] = ir;
    U[tid] = vx * ir; U[NP + tid] = vy * ir; U[2 * NP + tid] = vz * ir;
    FC[tid] = fc; DFC[tid] = dfc;
    for (int g = 0; g < NG; g++) XG[g * NP + tid] = act ? exp(-GETA[g] * (r * r)) : 0.0;
    ADJ4[tid] = 0ull;
  }
  if (tid < 64)
  {
    const unsigned bal = __ballot_sync(0xffffffffu, act);
    if (lane == 0) CNT[1 + warp] = (int) bal;
  }
  __syncthreads();
  const unsigned long long ACT =
      (unsigned long long) (unsigned) CNT[1] | ((unsigned long long) (unsigned) CNT[2] << 32);

  // ---- phase B: which neighbor pairs (j<k) can contribute; record slots -------------------
  // (nothing to do for radial-only parameter sets: no records, no triple loop)
  for (int p0 = 0; p0 < (NG > 0 ? n * n : 0); p0 += TT)
  {
    const int p = p0 + tid;
    bool rec = false, hit4 = false;
    int j = 0, k = 0;
    if (p < n * n)
    {
      j = p / n;
      k = p - j * n;
      if (j < k && ((ACT >> j) & 1ull) && ((ACT >> k) & 1ull))
      {
        const double r = sqrt(xnorm2(xsub(X[k], X[j]), xsub(X[NP + k], X[NP + j]),
                                     xsub(X[2 * NP + k], X[2 * NP + j])));  // :533-537
        hit4 = !(r > P.rcut[SPC[j] * S + SPC[k]]);                           // :729
        rec = hit4 || A.has_g5;
      }
    }
    const unsigned bal = __ballot_sync(0xffffffffu, rec);
    int base = 0;
    if (lane == 0 && bal) base = atomicAdd(&CNT[0], __popc(bal));
    base = __shfl_sync(0xffffffffu, base, 0);
    if (rec)
    {
      const int slot = base + __popc(bal & ((1u << lane) - 1u));
      if (slot < RC)
      {
        PAIR[slot] = (unsigned short) (j | (k << 8));
        PIDX[j * NP + k] = (unsigned short) slot;
        PIDX[k * NP + j] = (unsigned short) slot;
      }
      if (hit4)
      {
        unsigned* a32 = reinterpret_cast<unsigned*>(ADJ4);  // native 32-bit ATOMS.OR
        atomicOr(a32 + 2 * j + (k >> 5), 1u << (k & 31));
        atomicOr(a32 + 2 * k + (j >> 5), 1u << (j & 31));
      }
    }
  }
  if (tid < n) ADJA[tid] = ((ACT >> tid) & 1ull) ? (ACT & ~(1ull << tid)) : 0ull;
  __syncthreads();
  const int nrec = CNT[0];
  if (nrec > RC)
  {
    if (tid == 0) A.overflow[1 + atomicAdd(A.overflow, 1)] = t;
    return;
  }

  // ---- phase C: pair-jk records ------------------------------------------------------------
  for (int r0 = tid; r0 < nrec; r0 += TT)
  {
    const int j = PAIR[r0] & 0xff, k = PAIR[r0] >> 8;
    const double vx = xsub(X[k], X[j]), vy = xsub(X[NP + k], X[NP + j]),
                 vz = xsub(X[2 * NP + k], X[2 * NP + j]);
    const double r = sqrt(xnorm2(vx, vy, vz));
    const double ir = 1.0 / r;
    double fc, dfc;
    kb_cutoff(r, P.rcut[SPC[j] * S + SPC[k]], fc, dfc);  // 0 beyond rc_jk (g4 then vanishes)
    RR[r0] = r;
    RU[r0] = vx * ir; RU[RC + r0] = vy * ir; RU[2 * RC + r0] = vz * ir;
    RFC[r0] = fc; RDFC[r0] = dfc;
    for (int g = 0; g < NG; g++)
      RX[g * RC + r0] = (GKIND[g] == KB_G4) ? exp(-GETA[g] * (r * r)) : 1.0;
  }

  // ---- phase D: radial families (one contribution per (set, neighbor): plain stores) ------
  if (tid < n)
  {
    const bool on = (ACT >> tid) & 1ull;
    const double r = Rr[tid], fc = FC[tid], dfc = DFC[tid];
    const double ux = U[tid], uy = U[NP + tid], uz = U[2 * NP + tid];
    for (int q = 0; q < NT; q++)
    {
      double phi = 0.0, dphi = 0.0;
      if (on) kb_radial(P.two_kind[q], P.two_p0[q], P.two_p1[q], r, fc, dfc, phi, dphi);
      PHI2[q * NP + tid] = phi;
      if (GRAD)
      {
        double* o = stage + (P.two_out[q] * (n + 1) + tid) * 3;
        o[0] = dphi * ux; o[1] = dphi * uy; o[2] = dphi * uz;  // :497-502
      }
    }
  }
  __syncthreads();

  // ---- phase E: angular families; item = (slot k, group g) ---------------------------------
  const int KPW = 32 / NGP;
  const int nwi = (n + KPW - 1) / KPW;
  for (int wi = warp; wi < nwi; wi += TW)
  {
    const int k = wi * KPW + lane / NGP;
    const int g = lane & (NGP - 1);
    const bool valid = (k < n) && (g < NG);
    double acc[GW][3], phi[GW];
#pragma unroll
    for (int e = 0; e < GW; e++) { acc[e][0] = acc[e][1] = acc[e][2] = 0.0; phi[e] = 0.0; }
    int kind = KB_G4;
    if (valid)
    {
      kind = GKIND[g];
      const bool is5 = kind == KB_G5;
      unsigned long long mask = is5 ? ADJA[k] : ADJ4[k];
      const double eta = GETA[g];
      const double m2eta = -2.0 * eta;
      const double rik = Rr[k], rik2 = R2[k], irik = IR[k], fik = FC[k], dfik = DFC[k];
      const double uikx = U[k], uiky = U[NP + k], uikz = U[2 * NP + k];
      const double xik = XG[g * NP + k];
      const unsigned short* prow = PIDX + k * NP;
      const double* xgj = XG + g * NP;
      const double* rxg = RX + g * RC;
      while (mask)
      {
        const int j = __ffsll((long long) mask) - 1;
        mask &= mask - 1;
        const int rec = prow[j];
        const double rij2 = R2[j], irij = IR[j], fij = FC[j];
        const double rjk = RR[rec], fjk = is5 ? 1.0 : RFC[rec], dfjk = is5 ? 0.0 : RDFC[rec];
        const double sg = (k > j) ? 1.0 : -1.0;  // record orientation is low slot -> high slot
        const double ujkx = sg * RU[rec], ujky = sg * RU[RC + rec], ujkz = sg * RU[2 * RC + rec];
        const double rjk2 = rjk * rjk;
        const double h = 0.5 * irij * irik;
        const double ct = (rij2 + rik2 - rjk2) * h;            // cos(theta_jik), :744
        const double dct_ik = (rik2 - rij2 + rjk2) * h * irik;  // :746
        const double dct_jk = -2.0 * rjk * h;                   // :747
        const double fijk = fij * fjk;
        const double F = fijk * fik;
        const double dF_ik = dfik * fijk;
        const double dF_jk = dfjk * (fij * fik);
        const double E = xgj[j] * xik * rxg[rec];
        const double eF = E * F;
        const double a_ik = E * (m2eta * rik * F + dF_ik);
        const double a_jk = is5 ? 0.0 : E * (m2eta * rjk * F + dF_jk);
        const double w_ik = eF * dct_ik, w_jk = eF * dct_jk;
        const double v1x = w_ik * uikx + w_jk * ujkx, v1y = w_ik * uiky + w_jk * ujky,
                     v1z = w_ik * uikz + w_jk * ujkz;
        const double v2x = a_ik * uikx + a_jk * ujkx, v2y = a_ik * uiky + a_jk * ujky,
                     v2z = a_ik * uikz + a_jk * ujkz;
        double C[GW], Cp[GW];
        if (ZLK == 1)
        {
          // entries: e = 2*zi + si, zeta = {1,2,4,16}[zi], lambda = {-1,+1}[si]
#pragma unroll
          for (int s = 0; s < 2; s++)
          {
            const double lam = s ? 1.0 : -1.0;
            double b = 1.0 + lam * ct;
            const bool ok = b > 0.0;  // guard of :755-767
            b = ok ? b : 0.0;
            const double b2 = b * b, b3 = b2 * b, b4 = b2 * b2, b8 = b4 * b4;
            const double b16 = b8 * b8, b15 = b8 * b4 * b3;
            C[0 + s] = b;   Cp[0 + s] = ok ? lam : 0.0;
            C[2 + s] = b2;  Cp[2 + s] = 2.0 * lam * b;
            C[4 + s] = b4;  Cp[4 + s] = 4.0 * lam * b3;
            C[6 + s] = b16; Cp[6 + s] = 16.0 * lam * b15;
          }
        }
        else
        {
#pragma unroll
          for (int e = 0; e < GW; e++)
          {
            const double zt = GZETA[g * GW + e], lam = GLAM[g * GW + e];
            const double b = 1.0 + lam * ct;
            C[e] = 0.0; Cp[e] = 0.0;
            if (b > 0.0 && GOUT[g * GW + e] >= 0)
            {
              C[e] = pow(b, zt);
              Cp[e] = zt * lam * (C[e] / b);
            }
          }
        }
#pragma unroll
        for (int e = 0; e < GW; e++)
        {
          phi[e] = fma(C[e], eF, phi[e]);
          if (GRAD)
          {
            acc[e][0] = fma(C[e], v2x, fma(Cp[e], v1x, acc[e][0]));
            acc[e][1] = fma(C[e], v2y, fma(Cp[e], v1y, acc[e][1]));
            acc[e][2] = fma(C[e], v2z, fma(Cp[e], v1z, acc[e][2]));
          }
        }
      }
    }
    __syncwarp();
    // prefactor 2^(1-zeta) applied once per item
    double p2[GW];
#pragma unroll
    for (int e = 0; e < GW; e++)
    {
      if (ZLK == 1) p2[e] = (e < 2) ? 1.0 : (e < 4) ? 0.5 : (e < 6) ? 0.125 : 1.0 / 32768.0;
      else p2[e] = valid ? exp2(1.0 - GZETA[g * GW + e]) : 0.0;
    }
    if (GRAD && valid)
    {
#pragma unroll
      for (int e = 0; e < GW; e++)
      {
        const int d = GOUT[g * GW + e];
        if (d >= 0)
        {
          double* o = stage + (d * (n + 1) + k) * 3;
          o[0] = p2[e] * acc[e][0]; o[1] = p2[e] * acc[e][1]; o[2] = p2[e] * acc[e][2];
        }
      }
    }
    // G[d] = sum over unordered triples = 1/2 sum over ordered pairs
#pragma unroll
    for (int e = 0; e < GW; e++)
    {
      double v = 0.5 * p2[e] * phi[e];
      for (int m = NGP; m < 32; m <<= 1) v += shfl_xor_d(v, m);
      if (lane < NGP) PART[wi * (NGP * GW) + lane * GW + e] = v;
    }
  }
  __syncthreads();

  // ---- phase F: centre slot by translation invariance; zeta ---------------------------------
  if (GRAD)
    for (int q = tid; q < 3 * D; q += TT)
    {
      const int d = q / 3, c = q - 3 * d;
      const double* rp = stage + d * row + c;
      double s = 0.0;
      for (int k = 0; k < n; k++) s += rp[3 * k];
      stage[d * row + 3 * n + c] = -s;
    }
  double* zrow = A.zeta + (int64_t) t * D;
  for (int q = tid; q < NT; q += TT)
  {
    double s = 0.0;
    for (int j = 0; j < n; j++) s += PHI2[q * NP + j];
    zrow[P.two_out[q]] = s;
  }
  for (int q = tid; q < NG * GW; q += TT)
  {
    const int d = GOUT[q];
    if (d >= 0)
    {
      const int g = q / GW, e = q - g * GW;
      double s = 0.0;
      for (int wi = 0; wi < nwi; wi++) s += PART[wi * (NGP * GW) + g * GW + e];
      zrow[d] = s;
    }
  }
  if (!GRAD) return;
  __syncthreads();

  // ---- phase G: one contiguous, vectorised, coalesced copy of the block --------------------
  {
    const int64_t o0 = A.dz_ptr[t];
    const int len = (int) (A.dz_ptr[t + 1] - o0);  // padded to a multiple of 16 doubles
    const int exact = D * row;
    if (sizeof(TOut) == 8)
    {
      double2* out = reinterpret_cast<double2*>(static_cast<double*>(A.dz) + o0);
      const double2* in = reinterpret_cast<const double2*>(stage);
      for (int q = tid; q < len / 2; q += TT)
      {
        double2 v = in[q];
        if (2 * q + 1 >= exact)
        {
          if (2 * q >= exact) v.x = 0.0;
          v.y = 0.0;
        }
        out[q] = v;
      }
    }
    else
    {
      TOut* out = static_cast<TOut*>(A.dz) + o0;
      for (int q = tid; q < len; q += TT) out[q] = q < exact ? (TOut) stage[q] : (TOut) 0;
    }
  }
}

}  // namespace

int kb_symfun_tiled_launch(const kb_symfun_params& P, int n_centers, const int32_t* d_centers,
                           const double* d_coords, const int32_t* d_species,
                           const int64_t* d_nl_off, const int32_t* d_nl_idx, int maxn,
                           double* d_zeta, void* d_dz, int dz_dtype, const int64_t* d_dz_ptr,
                           int32_t* d_overflow, bool* may_overflow, cudaStream_t st)
{
  *may_overflow = false;
  if (n_centers <= 0) return KB_OK;
  int NP = maxn < 2 ? 2 : maxn;
  if (NP > 64) { NP = 64; *may_overflow = true; }
  NP = (NP + 1) & ~1;
  int NGP = 1;
  while (NGP < P.n_groups) NGP <<= 1;
  if (NGP > 32) return KB_ERR_LIMIT;
  const int KPW = 32 / NGP;
  const int NPW = (NP + KPW - 1) / KPW;
  const bool grad = d_dz != nullptr;
  const int stage_elems = grad ? ((P.n_desc * 3 * (NP + 1) + 15) & ~15) : 0;
  int has_g5 = 0;
  for (int g = 0; g < P.n_groups; g++) has_g5 |= (P.grp_kind[g] == KB_G5);
  int RC = P.n_groups > 0 ? NP * (NP - 1) / 2 : 1;
  const size_t limit = 227 * 1024;
  // prefer two CTAs per SM: shrink the record capacity before giving up residency
  while (RC > 64 && tiled_smem_bytes(P, NP, RC, NGP, NPW, stage_elems) > limit)
  {
    RC = RC * 3 / 4;
    *may_overflow = true;
  }
  const size_t smem = tiled_smem_bytes(P, NP, RC, NGP, NPW, stage_elems);
  if (smem > limit) return KB_ERR_LIMIT;

  TiledArgs A;
  A.n_centers = n_centers; A.centers = d_centers; A.coords = d_coords; A.species = d_species;
  A.nl_off = d_nl_off; A.nl_idx = d_nl_idx; A.zeta = d_zeta; A.dz = d_dz; A.dz_ptr = d_dz_ptr;
  A.overflow = d_overflow; A.NP = NP; A.RC = RC; A.NGP = NGP; A.NPW = NPW;
  A.stage_elems = stage_elems; A.has_g5 = has_g5;

#define KB_TILED_LAUNCH(Z, G, T)                                                                \
  do {                                                                                          \
    KB_CUDA(cudaFuncSetAttribute(symfun_tiled_kernel<Z, G, T>,                                  \
                                 cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));     \
    symfun_tiled_kernel<Z, G, T><<<n_centers, TT, smem, st>>>(P, A);                            \
  } while (0)
  const bool f32 = dz_dtype == KB_F32;
  if (P.canonical_zl)
  {
    if (!grad) KB_TILED_LAUNCH(1, false, double);
    else if (f32) KB_TILED_LAUNCH(1, true, float);
    else KB_TILED_LAUNCH(1, true, double);
  }
  else
  {
    if (!grad) KB_TILED_LAUNCH(0, false, double);
    else if (f32) KB_TILED_LAUNCH(0, true, float);
    else KB_TILED_LAUNCH(0, true, double);
  }
#undef KB_TILED_LAUNCH
  KB_LAUNCH_CHECK();
  return KB_OK;
}
