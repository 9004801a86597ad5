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

// radial families g1/g2/g3, sym_fn.cpp:604-674.  Caller guarantees r <= rc (sym_fn.cpp:451).
__device__ __forceinline__ void kb_radial(int kind, double p0, double p1, double r, double fc,
                                          double dfc, double& phi, double& dphi)
{
  if (kind == KB_G2)
  {
    const double dr = r - p1;
    const double e = exp(-p0 * dr * dr);
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
