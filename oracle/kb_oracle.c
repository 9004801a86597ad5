/*
 * kb_oracle.c — TEST INFRASTRUCTURE ONLY.
 *
 * A plain-C, single-threaded CPU restatement of the algorithm on the hot path of
 * openkim/kliff (neighbor list -> symmetry functions + dG/dr).  It exists so that the CUDA
 * path can be checked against something that runs on the GPU box (where /root/reference does
 * not exist).  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs may load it.  The product (kliff_b200/) never does: it fails loudly when the
 * CUDA library is missing.
 *
 * Parity status: PINNED.  tests/test_oracle.py checks every function here against
 *   (a) the reference's own golden vectors (tests/golden/ref_kat_*.npz, extracted from
 *       /root/reference/tests/descriptors/test_symmetry_function.py and tests/test_neighbor.py
 *       by tests/golden/make_golden.py), and
 *   (b) the unmodified reference C++ compiled into oracle/_ref/libkliff_ref.so (bitwise on the
 *       integer outputs and paddings, <= 4 ulp-ish on floating point; observed: identical).
 *
 * Every function cites the reference lines whose arithmetic (including operation ORDER, which
 * matters for the bit-exact integer outputs) it restates.  Paths are relative to
 * /root/reference/.  Build: gcc -O2 -ffp-contract=off (no FMA contraction, like the reference's
 * x86-64 -O2 build).
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#define KBO_TOL 1.0e-10 /* kliff/neighbor/neighbor_list.cpp:14 (TOL), helper.hpp:10 (SMALL) */
#define KBO_PI 3.1415926535897932 /* symmetry_function/sym_fn.cpp:9 */

/* ------------------------------------------------------------------------------------ */
/* small 3x3 helpers: kliff/neighbor/helper.hpp:18-93                                  */
/* ------------------------------------------------------------------------------------ */
static double dot3(const double* a, const double* b)
{
  return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; /* helper.hpp:26-29 */
}
static void cross3(const double* a, const double* b, double* c)
{
  c[0] = a[1] * b[2] - a[2] * b[1]; /* helper.hpp:33-38 */
  c[1] = a[2] * b[0] - a[0] * b[2];
  c[2] = a[0] * b[1] - a[1] * b[0];
}
static double minor2(double p, double q, double r, double s) { return (p * s) - (q * r); }

/* inverse via adjugate / determinant, helper.hpp:42-93 (term order of det kept) */
static int inv3(const double* m, double* out)
{
  out[0] = minor2(m[4], m[5], m[7], m[8]);
  out[1] = minor2(m[2], m[1], m[8], m[7]);
  out[2] = minor2(m[1], m[2], m[4], m[5]);
  out[3] = minor2(m[5], m[3], m[8], m[6]);
  out[4] = minor2(m[0], m[2], m[6], m[8]);
  out[5] = minor2(m[2], m[0], m[5], m[3]);
  out[6] = minor2(m[3], m[4], m[6], m[7]);
  out[7] = minor2(m[1], m[0], m[7], m[6]);
  out[8] = minor2(m[0], m[1], m[3], m[4]);
  double d = m[0] * m[4] * m[8] - m[0] * m[5] * m[7] - m[1] * m[3] * m[8]
             + m[1] * m[5] * m[6] + m[2] * m[3] * m[7] - m[2] * m[4] * m[6];
  if (fabs(d) < KBO_TOL) return 1;
  for (int q = 0; q < 9; q++) out[q] /= d;
  return 0;
}

/* ------------------------------------------------------------------------------------ */
/* Padding (periodic image) atoms: kliff/neighbor/neighbor_list.cpp:283-418             */
/*   cell: rows are lattice vectors.  Output order: shift (a,b,c) lexicographic, atom    */
/*   index innermost (:369-383).  Two-call protocol: pass pad_coords == NULL to count.   */
/*   Returns 0, or 1 when the cell is singular (:300-301).                              */
/* ------------------------------------------------------------------------------------ */
int kbo_create_paddings(int natoms, double cutoff, const double* cell, const int* pbc,
                        const double* coords, const int* species, long capacity,
                        double* pad_coords, int* pad_species, int* pad_image, long* npad)
{
  double tc[9], fc[9];
  /* transpose, :296 / helper.hpp:58-69 */
  for (int r = 0; r < 3; r++)
    for (int c = 0; c < 3; c++) tc[3 * r + c] = cell[3 * c + r];
  if (inv3(tc, fc)) return 1;

  double* frac = (double*) malloc(sizeof(double) * 3 * (size_t) (natoms > 0 ? natoms : 1));
  double lo[3] = {1e10, 1e10, 1e10}, hi[3] = {-1e10, -1e10, -1e10}; /* :305-306 */
  for (int a = 0; a < natoms; a++)
  {
    for (int c = 0; c < 3; c++)
    {
      double f = dot3(fc + 3 * c, coords + 3 * a); /* :312-314 */
      frac[3 * a + c] = f;
      if (f < lo[c]) lo[c] = f;
      if (f > hi[c]) hi[c] = f;
    }
  }
  for (int c = 0; c < 3; c++) { lo[c] -= KBO_TOL; hi[c] += KBO_TOL; } /* :329-334 */

  /* face distances, :337-353 */
  double x[3], dist[3];
  cross3(cell + 3, cell + 6, x);
  double vol = fabs(dot3(cell, x));
  dist[0] = vol / sqrt(x[0] * x[0] + x[1] * x[1] + x[2] * x[2]);
  cross3(cell + 6, cell + 0, x);
  dist[1] = vol / sqrt(x[0] * x[0] + x[1] * x[1] + x[2] * x[2]);
  cross3(cell, cell + 3, x);
  dist[2] = vol / sqrt(x[0] * x[0] + x[1] * x[1] + x[2] * x[2]);

  double ratio[3], slack[3];
  int nrep[3];
  for (int c = 0; c < 3; c++)
  {
    ratio[c] = cutoff / dist[c];                   /* :356-357 */
    nrep[c] = (int) ceil(ratio[c]);                /* :359-361 */
    slack[c] = (double) nrep[c] - ratio[c];        /* :363-365 */
  }

  long count = 0;
  for (int sa = -nrep[0]; sa <= nrep[0]; sa++)
    for (int sb = -nrep[1]; sb <= nrep[1]; sb++)
      for (int sc = -nrep[2]; sc <= nrep[2]; sc++)
      {
        if (sa == 0 && sb == 0 && sc == 0) continue;           /* :376 */
        if ((!pbc[0] && sa) || (!pbc[1] && sb) || (!pbc[2] && sc)) continue; /* :379-381 */
        const int sh[3] = {sa, sb, sc};
        for (int a = 0; a < natoms; a++)
        {
          const double* f = frac + 3 * a;
          int keep = 1;
          for (int c = 0; c < 3 && keep; c++) /* outermost shells only, :392-397 */
          {
            if (sh[c] == -nrep[c] && f[c] - lo[c] < slack[c]) keep = 0;
            if (sh[c] == nrep[c] && hi[c] - f[c] < slack[c]) keep = 0;
          }
          if (!keep) continue;
          if (pad_coords)
          {
            if (count >= capacity) { free(frac); return 2; }
            double g[3] = {sa + f[0], sb + f[1], sc + f[2]}; /* :400 */
            pad_coords[3 * count + 0] = dot3(tc, g);          /* :403-405 */
            pad_coords[3 * count + 1] = dot3(tc + 3, g);
            pad_coords[3 * count + 2] = dot3(tc + 6, g);
            pad_species[count] = species[a];
            pad_image[count] = a;
          }
          count++;
        }
      }
  free(frac);
  *npad = count;
  return 0;
}

/* ------------------------------------------------------------------------------------ */
/* Cell-list full neighbor list: kliff/neighbor/neighbor_list.cpp:78-243 and             */
/* coords_to_index helper.hpp:96-108.  One cutoff (the Python layer only ever passes one, */
/* kliff/neighbor/neighbor.py:107).  Two-call protocol: indices == NULL fills counts and  */
/* offsets[ntot+1] only.  Neighbor order = reference cell-scan order.  Returns 0; 1 on    */
/* "cell size too large" (:124-128) or atom collision rsq < 1e-10 (:199-209).            */
/* ------------------------------------------------------------------------------------ */
static void bin_of(const double* p, const int* nb, const double* hi, const double* lo, int* b)
{
  for (int c = 0; c < 3; c++)
  {
    b[c] = (int) (((p[c] - lo[c]) / (hi[c] - lo[c])) * nb[c]);
    if (b[c] > nb[c] - 1) b[c] = nb[c] - 1;
  }
}

int kbo_nl_build(int ntot, const double* coords, double infl, double cutoff, const int* need,
                 int* counts, long long* offsets, int* indices)
{
  double lo[3], hi[3];
  for (int c = 0; c < 3; c++) { lo[c] = coords[c]; hi[c] = coords[c] + 1.0; } /* :90-96 */
  for (int a = 0; a < ntot; a++)
    for (int c = 0; c < 3; c++)
    {
      double v = coords[3 * a + c];
      if (hi[c] < v) hi[c] = v;
      if (lo[c] > v) lo[c] = v;
    }
  int nb[3];
  for (int c = 0; c < 3; c++)
  {
    nb[c] = (int) ((hi[c] - lo[c]) / infl); /* :115-117 */
    if (nb[c] <= 0) nb[c] = 1;
  }
  long long ncell = (long long) nb[0] * nb[1] * nb[2];
  if (ncell > 1000000000LL) return 1;

  /* counting sort into bins, atom index ascending inside a bin (== push_back order, :131-142) */
  int* start = (int*) calloc((size_t) ncell + 1, sizeof(int));
  int* bin = (int*) malloc(sizeof(int) * (size_t) (ntot > 0 ? ntot : 1));
  int* order = (int*) malloc(sizeof(int) * (size_t) (ntot > 0 ? ntot : 1));
  for (int a = 0; a < ntot; a++)
  {
    int b[3];
    bin_of(coords + 3 * a, nb, hi, lo, b);
    bin[a] = b[0] + b[1] * nb[0] + b[2] * nb[0] * nb[1];
    start[bin[a] + 1]++;
  }
  for (long long q = 0; q < ncell; q++) start[q + 1] += start[q];
  {
    int* fill = (int*) malloc(sizeof(int) * (size_t) ncell);
    memcpy(fill, start, sizeof(int) * (size_t) ncell);
    for (int a = 0; a < ntot; a++) order[fill[bin[a]]++] = a;
    free(fill);
  }

  const double cutsq = cutoff * cutoff; /* :152-156 */
  long long total = 0;
  int err = 0;
  for (int a = 0; a < ntot && !err; a++)
  {
    int cnt = 0;
    if (need[a])
    {
      const double px = coords[3 * a], py = coords[3 * a + 1], pz = coords[3 * a + 2];
      int b[3];
      bin_of(coords + 3 * a, nb, hi, lo, b);
      int x0 = b[0] > 0 ? b[0] - 1 : 0, x1 = b[0] + 1 < nb[0] - 1 ? b[0] + 1 : nb[0] - 1;
      int y0 = b[1] > 0 ? b[1] - 1 : 0, y1 = b[1] + 1 < nb[1] - 1 ? b[1] + 1 : nb[1] - 1;
      int z0 = b[2] > 0 ? b[2] - 1 : 0, z1 = b[2] + 1 < nb[2] - 1 ? b[2] + 1 : nb[2] - 1;
      for (int bx = x0; bx <= x1 && !err; bx++)   /* x outer ... z inner, :175-185 */
        for (int by = y0; by <= y1 && !err; by++)
          for (int bz = z0; bz <= z1 && !err; bz++)
          {
            int q = bx + by * nb[0] + bz * nb[0] * nb[1];
            for (int s = start[q]; s < start[q + 1]; s++)
            {
              int n = order[s];
              if (n == a) continue;
              double dx = coords[3 * n] - px, dy = coords[3 * n + 1] - py,
                     dz = coords[3 * n + 2] - pz;
              double rsq = dx * dx + dy * dy + dz * dz; /* :197 */
              if (rsq < KBO_TOL) { err = 1; break; }     /* :199 */
              if (rsq < cutsq)                           /* :212 strict */
              {
                if (indices) indices[total + cnt] = n;
                cnt++;
              }
            }
          }
    }
    counts[a] = cnt;
    offsets[a] = total;
    total += cnt;
  }
  offsets[ntot] = total;
  free(start); free(bin); free(order);
  return err;
}

/* ------------------------------------------------------------------------------------ */
/* Symmetry functions: kliff/legacy/descriptors/symmetry_function/sym_fn.cpp              */
/* ------------------------------------------------------------------------------------ */
static double fcut(double r, double rc) /* cut_cos :37-40 */
{
  return (r < rc) ? 0.5 * (cos(KBO_PI * r / rc) + 1.0) : 0.0;
}
static double dfcut(double r, double rc) /* d_cut_cos :42-45 */
{
  return (r < rc) ? -0.5 * KBO_PI / rc * sin(KBO_PI * r / rc) : 0.0;
}

/* radial families.  kind 1/2/3 = g1/g2/g3; p[0],p[1] as stored by sym_fn.py:257-262 */
static void radial(int kind, const double* p, double r, double rc, int grad, double* phi,
                   double* dphi)
{
  *dphi = 0.0;
  if (kind == 1) /* sym_g1/sym_d_g1 :604-616 */
  {
    *phi = fcut(r, rc);
    if (grad) *dphi = dfcut(r, rc);
  }
  else if (kind == 2) /* sym_g2 :618-625, sym_d_g2 :627-650 */
  {
    double eta = p[0], Rs = p[1];
    if (!grad) { *phi = exp(-eta * (r - Rs) * (r - Rs)) * fcut(r, rc); return; }
    if (r > rc) { *phi = 0.0; return; }
    double e = exp(-eta * (r - Rs) * (r - Rs));
    double de = -2 * eta * (r - Rs) * e;
    double f = fcut(r, rc), df = dfcut(r, rc);
    *phi = e * f;
    *dphi = de * f + e * df;
  }
  else /* sym_g3 :652-658, sym_d_g3 :660-674 */
  {
    double kappa = p[0];
    double c = cos(kappa * r);
    double f = fcut(r, rc);
    *phi = c * f;
    if (grad) *dphi = -kappa * sin(kappa * r) * f + c * dfcut(r, rc);
  }
}

/* angular families, kind 4/5 = g4/g5; p = (zeta, lambda, eta) (sym_fn.py:264-271).
 * r = (rij, rik, rjk), rc likewise.  sym_g4 :676-711, sym_d_g4 :713-809, sym_g5 :811-847,
 * sym_d_g5 :849-936.  The value-only and value+gradient branches of the reference differ in
 * how base^(zeta-1) is formed (g4: power/base :762; g5: pow(base, zeta-1) :897); kept. */
static void angular(int kind, const double* p, const double* r, const double* rc, int grad,
                    double* phi, double* dphi)
{
  const double zeta = p[0], lambda = p[1], eta = p[2];
  const double rij = r[0], rik = r[1], rjk = r[2];
  dphi[0] = dphi[1] = dphi[2] = 0.0;
  *phi = 0.0;
  if (rij > rc[0] || rik > rc[1]) return;
  if (kind == 4 && rjk > rc[2]) return;

  const double a2 = rij * rij, b2 = rik * rik, c2 = rjk * rjk;
  const double ct = (a2 + b2 - c2) / (2 * rij * rik);
  const double base = 1 + lambda * ct;
  const double esum = (kind == 4) ? (a2 + b2 + c2) : (a2 + b2);
  const double et = exp(-eta * esum);
  const double p2 = pow(2, 1 - zeta);
  const double fij = fcut(rij, rc[0]), fik = fcut(rik, rc[1]);
  const double fjk = (kind == 4) ? fcut(rjk, rc[2]) : 1.0;

  if (!grad)
  {
    double cterm = (base <= 0) ? 0.0 : pow(base, zeta);
    if (kind == 4) *phi = p2 * cterm * et * fij * fik * fjk;
    else *phi = p2 * cterm * et * fij * fik;
    return;
  }

  const double dct_ij = (a2 - b2 + c2) / (2 * a2 * rik);
  const double dct_ik = (b2 - a2 + c2) / (2 * rij * b2);
  const double dct_jk = -rjk / (rij * rik);
  double cterm, dcterm;
  if (base <= 0) { cterm = 0.0; dcterm = 0.0; }
  else if (kind == 4)
  {
    double pw = pow(base, zeta);
    cterm = pw;
    dcterm = zeta * (pw / base) * lambda; /* :761-765 */
  }
  else
  {
    cterm = pow(base, zeta);
    dcterm = zeta * pow(base, zeta - 1) * lambda; /* :896-897 */
  }
  const double dc_ij = dcterm * dct_ij, dc_ik = dcterm * dct_ik, dc_jk = dcterm * dct_jk;
  const double de_ij = -2 * et * eta * rij, de_ik = -2 * et * eta * rik;

  if (kind == 4)
  {
    const double de_jk = -2 * et * eta * rjk;
    const double fprod = fij * fik * fjk;
    const double df_ij = dfcut(rij, rc[0]) * fik * fjk;
    const double df_ik = dfcut(rik, rc[1]) * fij * fjk;
    const double df_jk = dfcut(rjk, rc[2]) * fij * fik;
    *phi = p2 * cterm * et * fprod;
    dphi[0] = p2 * (dc_ij * et * fprod + cterm * de_ij * fprod + cterm * et * df_ij);
    dphi[1] = p2 * (dc_ik * et * fprod + cterm * de_ik * fprod + cterm * et * df_ik);
    dphi[2] = p2 * (dc_jk * et * fprod + cterm * de_jk * fprod + cterm * et * df_jk);
  }
  else
  {
    const double fprod = fij * fik;
    const double df_ij = dfcut(rij, rc[0]) * fik;
    const double df_ik = dfcut(rik, rc[1]) * fij;
    *phi = p2 * cterm * et * fprod;
    dphi[0] = p2 * (dc_ij * et * fprod + cterm * de_ij * fprod + cterm * et * df_ij);
    dphi[1] = p2 * (dc_ik * et * fprod + cterm * de_ik * fprod + cterm * et * df_ik);
    dphi[2] = p2 * dc_jk * et * fprod; /* :934 */
  }
}

/*
 * One centre atom: Descriptor::generate_one_atom, sym_fn.cpp:416-602.
 *   fam_kind[nfam] in 1..5, fam_rows[nfam], fam_vals = all families' row-major parameter
 *   matrices back to back with cols = {1,2,1,3,3}[kind-1] (g1 has one dummy row, sym_fn.py:250-252).
 *   Descriptor index = running sum of rows in family order (:96-100).
 *   desc[D] and grad[D][numnei+1][3] must be zeroed by the caller (+= semantics, :482, :497-502);
 *   slot numnei is the centre atom (:495).
 */
void kbo_symfun_atom(int nspecies, const double* rcut, int nfam, const int* fam_kind,
                     const int* fam_rows, const double* fam_vals, int i, const double* coords,
                     const int* species, const int* neigh, int numnei, double* desc,
                     double* grad, int want_grad)
{
  static const int ncols[5] = {1, 2, 1, 3, 3};
  const int si = species[i];
  int has3 = 0;
  for (int f = 0; f < nfam; f++)
    if (fam_kind[f] >= 4) has3 = 1;
  const long stride = 3L * (numnei + 1);

  for (int jj = 0; jj < numnei; jj++)
  {
    const int j = neigh[jj];
    const int sj = species[j];
    const double rcij = rcut[si * nspecies + sj];
    double vij[3];
    for (int c = 0; c < 3; c++) vij[c] = coords[3 * j + c] - coords[3 * i + c];
    const double rij = sqrt(vij[0] * vij[0] + vij[1] * vij[1] + vij[2] * vij[2]);
    if (rij > rcij) continue; /* :451 */

    /* radial part, :455-508 */
    {
      int d = 0;
      const double* pv = fam_vals;
      for (int f = 0; f < nfam; f++)
      {
        const int kind = fam_kind[f], rows = fam_rows[f], nc = ncols[kind - 1];
        if (kind <= 3)
          for (int q = 0; q < rows; q++)
          {
            double phi, dphi;
            radial(kind, pv + q * nc, rij, rcij, want_grad, &phi, &dphi);
            desc[d + q] += phi;
            if (want_grad)
              for (int c = 0; c < 3; c++)
              {
                double pr = dphi * vij[c] / rij; /* :497 */
                grad[(d + q) * stride + 3 * numnei + c] -= pr;
                grad[(d + q) * stride + 3 * jj + c] += pr;
              }
          }
        d += rows;
        pv += rows * nc;
      }
    }
    if (!has3) continue; /* :511 */

    /* angular part, :514-600; kk > jj in list order */
    for (int kk = jj + 1; kk < numnei; kk++)
    {
      const int k = neigh[kk];
      const int sk = species[k];
      const double rcik = rcut[si * nspecies + sk], rcjk = rcut[sj * nspecies + sk];
      double vik[3], vjk[3];
      for (int c = 0; c < 3; c++)
      {
        vik[c] = coords[3 * k + c] - coords[3 * i + c];
        vjk[c] = coords[3 * k + c] - coords[3 * j + c];
      }
      const double rik = sqrt(vik[0] * vik[0] + vik[1] * vik[1] + vik[2] * vik[2]);
      const double rjk = sqrt(vjk[0] * vjk[0] + vjk[1] * vjk[1] + vjk[2] * vjk[2]);
      if (rik > rcik) continue; /* :541 */
      const double rv[3] = {rij, rik, rjk}, rcv[3] = {rcij, rcik, rcjk};

      int d = 0;
      const double* pv = fam_vals;
      for (int f = 0; f < nfam; f++)
      {
        const int kind = fam_kind[f], rows = fam_rows[f], nc = ncols[kind - 1];
        if (kind >= 4)
          for (int q = 0; q < rows; q++)
          {
            double phi, dphi[3];
            angular(kind, pv + q * nc, rv, rcv, want_grad, &phi, dphi);
            desc[d + q] += phi;
            if (want_grad)
              for (int c = 0; c < 3; c++)
              {
                double pij = dphi[0] * vij[c] / rij; /* :583-585 */
                double pik = dphi[1] * vik[c] / rik;
                double pjk = dphi[2] * vjk[c] / rjk;
                grad[(d + q) * stride + 3 * numnei + c] += -pij - pik; /* :587 */
                grad[(d + q) * stride + 3 * jj + c] += pij - pjk;
                grad[(d + q) * stride + 3 * kk + c] += pik + pjk;
              }
          }
        d += rows;
        pv += rows * nc;
      }
    }
  }
}

/* Range form for timing / bulk checks: atoms [first,last) from a CSR list; zeta kept, the
 * gradient block optionally stored packed at grad_out + 3*D*(offsets[i]-offsets[first] + (i-first)). */
void kbo_symfun_range(int nspecies, const double* rcut, int nfam, const int* fam_kind,
                      const int* fam_rows, const double* fam_vals, int D, int first, int last,
                      const double* coords, const int* species, const long long* offsets,
                      const int* indices, double* zeta_out, double* grad_out, double* scratch,
                      int want_grad)
{
  for (int i = first; i < last; i++)
  {
    const int n = (int) (offsets[i + 1] - offsets[i]);
    double* z = zeta_out + (size_t) (i - first) * D;
    double* g = scratch;
    if (grad_out) g = grad_out + 3 * (size_t) D * (size_t) (offsets[i] - offsets[first] + (i - first));
    for (int a = 0; a < D; a++) z[a] = 0.0;
    if (want_grad) memset(g, 0, sizeof(double) * 3 * (size_t) D * (size_t) (n + 1));
    kbo_symfun_atom(nspecies, rcut, nfam, fam_kind, fam_rows, fam_vals, i, coords, species,
                    indices + offsets[i], n, z, g, want_grad);
  }
}
