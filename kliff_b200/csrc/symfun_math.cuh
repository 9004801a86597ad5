// Per-pair / per-triple mathematics of the Behler-Parrinello symmetry functions
// (what sym_fn.cpp:37-45 and :604-936 of the reference compute), in fp64.
#pragma once
#include "common.cuh"

// cos cutoff and derivative, sym_fn.cpp:37-45: fc = (r < rc) ? 0.5 (cos(pi r/rc) + 1) : 0
__device__ __forceinline__ void kb_cutoff(double r, double rc, double& fc, double& dfc)
{
  if (r < rc)
  {
    double s, c;
    sincos(KB_PI * r / rc, &s, &c);
    fc = 0.5 * (c + 1.0);
    dfc = -0.5 * KB_PI / rc * s;
  }
  else
  {
    fc = 0.0;
    dfc = 0.0;
  }
}

// ---- branch-free elementary functions for the table phases ---------------------------------
// The library exp / sincos carry special-case branches (huge arguments, NaN, Payne-Hanek), which
// split the instruction stream into basic blocks and keep the compiler from overlapping the 7
// independent exponentials of a pair.  On this path the arguments are bounded (x <= 0 for the
// Gaussians, 0 <= r/rc <= 1 for the cutoff), so straight-line versions suffice; both are accurate
// to ~2 ulp (tests compare against the reference at 1e-10).

// exp(x) for x <= ~1; tends to 0 smoothly for very negative x.
__device__ __forceinline__ double kb_exp_neg(double x)
{
  const double SHIFT = 6755399441055744.0;  // 1.5 * 2^52: rounds to nearest integer in the low word
  const double t = fma(x, 1.4426950408889634, SHIFT);
  int n = __double2loint(t);
  const double nf = t - SHIFT;
  double f = fma(nf, -6.93147180369123816490e-01, x);  // Cody-Waite, ln2 = hi + lo
  f = fma(nf, -1.90821492927058770002e-10, f);
  // degree-12 Taylor polynomial in Estrin form: dependency depth 5 instead of 12 (this chain sits
  // in the triple loop, where a serial Horner evaluation was the single largest stall source)
  const double f2 = f * f;
  const double a0 = fma(f, 1.0, 1.0);
  const double a1 = fma(f, 1.66666666666666666667e-01, 0.5);
  const double a2 = fma(f, 8.33333333333333333333e-03, 4.16666666666666666667e-02);
  const double a3 = fma(f, 1.98412698412698412698e-04, 1.38888888888888888889e-03);
  const double a4 = fma(f, 2.75573192239858906526e-06, 2.48015873015873015873e-05);
  const double a5 = fma(f, 2.50521083854417187751e-08, 2.75573192239858906526e-07);
  const double f4 = f2 * f2;
  const double b0 = fma(a1, f2, a0);
  const double b1 = fma(a3, f2, a2);
  const double b2 = fma(a5, f2, a4);
  const double f8 = f4 * f4;
  const double d0 = fma(b1, f4, b0);
  const double d1 = fma(2.08767569878680989792e-09, f4, b2);
  const double p = fma(d1, f8, d0);
  n = n < -1021 ? -1021 : n;  // exp underflows to ~1e-308 instead of wrapping the exponent field
  return p * __hiloint2double((n + 1023) << 20, 0);
}

// exp(x) for x <= ~1 from a 64-entry table of 2^(j/64) (shared memory, filled by kb_exp_table_fill)
// and a degree-5 polynomial on |f| <= ln2/128 (truncation error f^6/720 < 4e-17): 12 FP64
// instructions with dependency depth 8 instead of 22 / 10 — the triple loop runs one per sweep step.
__device__ __forceinline__ void kb_exp_table_fill(double* tab, int lane)
{
  for (int j = lane; j < 64; j += 32) tab[j] = exp2((double) j * (1.0 / 64.0));
}
__device__ __forceinline__ double kb_exp_tab(double x, const double* __restrict__ tab)
{
  const double SHIFT = 6755399441055744.0;  // 1.5 * 2^52
  const double t = fma(x, 92.33248261689366, SHIFT);  // 64 / ln 2
  int n = __double2loint(t);
  const double nf = t - SHIFT;
  double f = fma(nf, -1.083042469326756e-02, x);  // ln2/64 = hi + lo (hi: 32 significant bits)
  f = fma(nf, -2.9815858269852933e-12, f);
  n = n < -64000 ? -64000 : n;  // below e^-693 the result is ~1e-301 instead of a wrapped exponent
  const double tj = tab[n & 63];
  const double f2 = f * f;
  const double a0 = 1.0 + f;
  const double a1 = fma(f, 1.66666666666666666667e-01, 0.5);
  const double a2 = fma(f, 8.33333333333333333333e-03, 4.16666666666666666667e-02);
  const double f4 = f2 * f2;
  const double b = fma(a1, f2, a0);
  const double p = fma(a2, f4, b);
  // 2^(n >> 6) applied to the table value through its exponent field
  const double ts = __hiloint2double(__double2hiint(tj) + ((n >> 6) << 20), __double2loint(tj));
  return p * ts;
}

// 16-entry variant: the table is 128 bytes, one entry per 8-byte bank, so a warp's 32 lookups are never
// in conflict (the 64-entry table costs 4.6 shared-memory wavefronts per lookup instead of 2, ncu on the
// round-2 sweep, whose binding unit is the shared-memory pipe).  |f| <= ln2/32, degree-6 polynomial,
// truncation error f^7/5040 < 5e-16; 14 FP64 instructions.
__device__ __forceinline__ void kb_exp_table16_fill(double* tab, int lane)
{
  if (lane < 16) tab[lane] = exp2((double) lane * (1.0 / 16.0));
}
__device__ __forceinline__ double kb_exp_tab16(double x, const double* __restrict__ tab)
{
  const double SHIFT = 6755399441055744.0;  // 1.5 * 2^52
  const double t = fma(x, 23.083120654223414, SHIFT);  // 16 / ln 2
  int n = __double2loint(t);
  const double nf = t - SHIFT;
  double f = fma(nf, -0.04332169877307024, x);  // ln2/16 = hi + lo (hi: 32 significant bits)
  f = fma(nf, -1.1926343307941173e-11, f);
  n = n < -16000 ? -16000 : n;  // below e^-693 the result is ~1e-301 instead of a wrapped exponent
  const double tj = tab[n & 15];
  const double f2 = f * f;
  const double a0 = 1.0 + f;
  const double a1 = fma(f, 1.66666666666666666667e-01, 0.5);
  const double a2 = fma(f, 8.33333333333333333333e-03, 4.16666666666666666667e-02);
  const double f4 = f2 * f2;
  const double b = fma(a1, f2, a0);
  const double c = fma(f2, 1.38888888888888888889e-03, a2);
  const double p = fma(c, f4, b);
  const double ts = __hiloint2double(__double2hiint(tj) + ((n >> 4) << 20), __double2loint(tj));
  return p * ts;
}

// The same with three FP64 instructions fewer (11): ln2/16 as ONE constant (its rounding error, 4.8e-18,
// times |n| moves f by 7e-16 at x = -6 and 4e-15 at x = -40, i.e. the result by that relative amount: the
// sweep's arguments are -eta (rij^2 + rik^2 + rjk^2) >= -25) and the degree-6 polynomial as a Horner chain
// over f^2 with paired coefficients.  Checked against exp() on the host grid of tests/test_abi.py's twin.
__device__ __forceinline__ double kb_exp_tab16s(double x, const double* __restrict__ tab)
{
  const double SHIFT = 6755399441055744.0;  // 1.5 * 2^52
  const double t = fma(x, 23.083120654223414, SHIFT);  // 16 / ln 2
  int n = __double2loint(t);
  const double nf = t - SHIFT;
  const double f = fma(nf, -0.043321698784996581, x);  // ln2 / 16
  n = n < -16000 ? -16000 : n;
  const double tj = tab[n & 15];
  const double f2 = f * f;
  const double a0 = 1.0 + f;
  const double a1 = fma(f, 1.66666666666666666667e-01, 0.5);
  const double a2 = fma(f, 8.33333333333333333333e-03, 4.16666666666666666667e-02);
  double q = fma(f2, 1.38888888888888888889e-03, a2);
  q = fma(f2, q, a1);
  const double p = fma(f2, q, a0);
  const double ts = __hiloint2double(__double2hiint(tj) + ((n >> 4) << 20), __double2loint(tj));
  return p * ts;
}

// (A 16-entry table with a degree-6 polynomial makes the lookup conflict-free but measured the same:
// 7.37 ms either way on 216 k atoms.)

// sin(pi y), cos(pi y) for 0 <= y <= 1 (also fine slightly outside).
__device__ __forceinline__ void kb_sincospi01(double y, double& s, double& c)
{
  const double q = rint(y + y);              // quadrant 0, 1, 2
  const double w = (y - 0.5 * q) * KB_PI;    // |w| <= pi/4
  const double w2 = w * w;
  double ps = -7.64716373181981647590e-13;   // sin: w - w^3/3! + ... - w^15/15!
  ps = fma(ps, w2, 1.60590438368216145994e-10);
  ps = fma(ps, w2, -2.50521083854417187751e-08);
  ps = fma(ps, w2, 2.75573192239858906526e-06);
  ps = fma(ps, w2, -1.98412698412698412698e-04);
  ps = fma(ps, w2, 8.33333333333333333333e-03);
  ps = fma(ps, w2, -1.66666666666666666667e-01);
  ps = fma(ps * w2, w, w);
  double pc = 4.77947733238738529744e-14;    // cos: 1 - w^2/2! + ... + w^16/16!
  pc = fma(pc, w2, -1.14707455977297247139e-11);
  pc = fma(pc, w2, 2.08767569878680989792e-09);
  pc = fma(pc, w2, -2.75573192239858906526e-07);
  pc = fma(pc, w2, 2.48015873015873015873e-05);
  pc = fma(pc, w2, -1.38888888888888888889e-03);
  pc = fma(pc, w2, 4.16666666666666666667e-02);
  pc = fma(pc, w2, -0.5);
  pc = fma(pc, w2, 1.0);
  const bool odd = (q == 1.0);
  const double sg = (q == 2.0) ? -1.0 : 1.0;
  s = odd ? pc : sg * ps;          // sin(w + q pi/2)
  c = odd ? -ps : sg * pc;         // cos(w + q pi/2)
}

// cos cutoff from a precomputed 1/rc (table phases); same definition as kb_cutoff.
__device__ __forceinline__ void kb_cutoff_fast(double r, double rc, double inv_rc, double& fc, double& dfc)
{
  double s, c;
  kb_sincospi01(r * inv_rc, s, c);
  const bool in = r < rc;
  fc = in ? 0.5 * (c + 1.0) : 0.0;
  dfc = in ? (-0.5 * KB_PI) * inv_rc * s : 0.0;
}

// radial families g1/g2/g3, sym_fn.cpp:604-674.  Caller guarantees r <= rc (sym_fn.cpp:451).
__device__ __forceinline__ void kb_radial(int kind, double p0, double p1, double r, double fc,
                                          double dfc, double& phi, double& dphi)
{
  if (kind == KB_G2)
  {
    const double dr = r - p1;
    const double arg = -p0 * dr * dr;
    const double e = (p0 >= 0.0) ? kb_exp_neg(arg) : exp(arg);  // uniform branch (p0 is a parameter)
    phi = e * fc;
    dphi = e * (dfc - 2.0 * p0 * dr * fc);
  }
  else if (kind == KB_G1)
  {
    phi = fc;
    dphi = dfc;
  }
  else
  {
    double s, c;
    sincos(p0 * r, &s, &c);
    phi = c * fc;
    dphi = c * dfc - p0 * s * fc;
  }
}

// One angular term g4/g5 for explicit (zeta, lambda, eta), generic exponent (pow), as in
// sym_fn.cpp:713-809 / :849-936.  r = (rij, rik, rjk); f/df = cutoff values of the three pairs
// (for g5 the jk pair does not enter).  Outputs phi and d(phi)/d(rij, rik, rjk).
__device__ __forceinline__ void kb_angular_generic(int kind, double zeta, double lambda, double eta,
                                                   const double* r, const double* f,
                                                   const double* df, double& phi, double* dphi)
{
  const double a2 = r[0] * r[0], b2 = r[1] * r[1], c2 = r[2] * r[2];
  const double inv = 1.0 / (r[0] * r[1]);
  const double ct = 0.5 * (a2 + b2 - c2) * inv;
  const double dct0 = 0.5 * (a2 - b2 + c2) * inv / r[0];
  const double dct1 = 0.5 * (b2 - a2 + c2) * inv / r[1];
  const double dct2 = -r[2] * inv;
  const double base = 1.0 + lambda * ct;
  double C = 0.0, dC = 0.0;
  if (base > 0.0)
  {
    C = pow(base, zeta);
    dC = zeta * lambda * (C / base);
  }
  const double p2 = exp2(1.0 - zeta);
  if (kind == KB_G4)
  {
    const double E = exp(-eta * (a2 + b2 + c2));
    const double F = f[0] * f[1] * f[2];
    phi = p2 * C * E * F;
    const double CE = p2 * C * E, dCEF = p2 * dC * E * F;
    dphi[0] = dCEF * dct0 + CE * (df[0] * f[1] * f[2] - 2.0 * eta * r[0] * F);
    dphi[1] = dCEF * dct1 + CE * (df[1] * f[0] * f[2] - 2.0 * eta * r[1] * F);
    dphi[2] = dCEF * dct2 + CE * (df[2] * f[0] * f[1] - 2.0 * eta * r[2] * F);
  }
  else
  {
    const double E = exp(-eta * (a2 + b2));
    const double F = f[0] * f[1];
    phi = p2 * C * E * F;
    const double CE = p2 * C * E, dCEF = p2 * dC * E * F;
    dphi[0] = dCEF * dct0 + CE * (df[0] * f[1] - 2.0 * eta * r[0] * F);
    dphi[1] = dCEF * dct1 + CE * (df[1] * f[0] - 2.0 * eta * r[1] * F);
    dphi[2] = dCEF * dct2;
  }
}
