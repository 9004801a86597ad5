// General symmetry-function kernel: one warp per centre atom, any neighbor count, any
// hyper-parameters; gradient contributions are accumulated straight into the packed global
// block with fp64 red.global (the block is L2 resident while the warp works on it).
//
// This is the fallback behind the tiled shared-memory kernel (symfun_tiled.cu): it takes atoms
// with more than 64 neighbors or whose pair tables do not fit in shared memory, and it doubles
// as an independent GPU cross-check in the tests.  Loop structure follows
// Descriptor::generate_one_atom (sym_fn.cpp:416-602): pairs jj, then kk > jj.
#include "common.cuh"
#include "symfun_math.cuh"

namespace {

constexpr int GEN_WARPS = 4;

template <typename TOut>
__global__ void __launch_bounds__(GEN_WARPS * 32)
symfun_general_kernel(const __grid_constant__ kb_symfun_params P, int n_centers,
                      const int32_t* __restrict__ centers, const int32_t* __restrict__ subset,
                      const double* __restrict__ coords, const int32_t* __restrict__ species,
                      const int64_t* __restrict__ nl_off, const int32_t* __restrict__ nl_idx,
                      double* __restrict__ zeta, TOut* __restrict__ dz,
                      const int64_t* __restrict__ dz_ptr)
{
  extern __shared__ double zacc_all[];  // [GEN_WARPS][D]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int t = blockIdx.x * GEN_WARPS + warp;
  if (t >= n_centers) return;
  if (subset) t = subset[t];
  const int D = P.n_desc;
  double* zacc = zacc_all + warp * D;
  const int i = centers ? centers[t] : t;
  const int64_t off = nl_off[i];
  const int n = (int) (nl_off[i + 1] - off);
  const int32_t* nb = nl_idx + off;
  const bool grad = dz != nullptr;
  TOut* blk = grad ? dz + dz_ptr[t] : nullptr;
  const int64_t row = 3LL * (n + 1);

  for (int d = lane; d < D; d += 32) zacc[d] = 0.0;
  if (grad)
  {
    const int64_t len = dz_ptr[t + 1] - dz_ptr[t];
    for (int64_t e = lane; e < len; e += 32) blk[e] = (TOut) 0;
  }
  __syncwarp();

  const double xi = coords[3 * i], yi = coords[3 * i + 1], zi = coords[3 * i + 2];
  const int si = species[i];

  // ---- radial part: lanes over jj (sym_fn.cpp:431-508)
  for (int j0 = 0; j0 < n; j0 += 32)
  {
    const int jj = j0 + lane;
    bool on = jj < n;
    double v[3] = {0, 0, 0}, r = 1.0, fc = 0, dfc = 0;
    if (on)
    {
      const int j = nb[jj];
      v[0] = xsub(coords[3 * j], xi);
      v[1] = xsub(coords[3 * j + 1], yi);
      v[2] = xsub(coords[3 * j + 2], zi);
      r = sqrt(xnorm2(v[0], v[1], v[2]));
      const double rc = P.rcut[si * P.n_species + species[j]];
      on = !(r > rc);  // :451
      if (on) kb_cutoff(r, rc, fc, dfc);
    }
    for (int q = 0; q < P.n_two; q++)
    {
      double phi = 0.0, dphi = 0.0;
      if (on) kb_radial(P.two_kind[q], P.two_p0[q], P.two_p1[q], r, fc, dfc, phi, dphi);
      const double tot = warp_sum(phi);
      if (lane == 0) zacc[P.two_out[q]] += tot;
      if (grad && on)
      {
        TOut* rowp = blk + P.two_out[q] * row;
#pragma unroll
        for (int c = 0; c < 3; c++)
        {
          const double pr = dphi * v[c] / r;
          atomicAdd(rowp + 3 * jj + c, (TOut) pr);
          atomicAdd(rowp + 3 * n + c, (TOut) -pr);
        }
      }
    }
  }

  // ---- angular part: jj serial, lanes over kk > jj (sym_fn.cpp:514-600)
  if (P.n_groups > 0)
    for (int jj = 0; jj < n; jj++)
    {
      const int j = nb[jj];
      const int sj = species[j];
      const double xj = coords[3 * j], yj = coords[3 * j + 1], zj = coords[3 * j + 2];
      const double vij[3] = {xsub(xj, xi), xsub(yj, yi), xsub(zj, zi)};
      const double rij = sqrt(xnorm2(vij[0], vij[1], vij[2]));
      const double rcij = P.rcut[si * P.n_species + sj];
      if (rij > rcij) continue;  // warp-uniform
      double fij, dfij;
      kb_cutoff(rij, rcij, fij, dfij);
      for (int k0 = jj + 1; k0 < n; k0 += 32)
      {
        const int kk = k0 + lane;
        bool on = kk < n;
        double vik[3] = {0, 0, 0}, vjk[3] = {0, 0, 0}, r[3] = {rij, 1.0, 1.0};
        double f[3] = {fij, 0, 0}, df[3] = {dfij, 0, 0};
        bool in_jk = false;
        if (on)
        {
          const int k = nb[kk];
          const int sk = species[k];
          const double xk = coords[3 * k], yk = coords[3 * k + 1], zk = coords[3 * k + 2];
          vik[0] = xsub(xk, xi); vik[1] = xsub(yk, yi); vik[2] = xsub(zk, zi);
          vjk[0] = xsub(xk, xj); vjk[1] = xsub(yk, yj); vjk[2] = xsub(zk, zj);
          r[1] = sqrt(xnorm2(vik[0], vik[1], vik[2]));
          r[2] = sqrt(xnorm2(vjk[0], vjk[1], vjk[2]));
          const double rcik = P.rcut[si * P.n_species + sk], rcjk = P.rcut[sj * P.n_species + sk];
          on = !(r[1] > rcik);  // :541
          if (on)
          {
            kb_cutoff(r[1], rcik, f[1], df[1]);
            in_jk = !(r[2] > rcjk);  // g4 only, :729
            if (in_jk) kb_cutoff(r[2], rcjk, f[2], df[2]);
          }
        }
        for (int g = 0; g < P.n_groups; g++)
        {
          const int kind = P.grp_kind[g];
          const bool live = on && (kind == KB_G5 || in_jk);
          for (int e = 0; e < P.grp_n[g]; e++)
          {
            const int d = P.grp_out[g][e];
            if (d < 0) continue;
            double phi = 0.0, dphi[3] = {0, 0, 0};
            if (live)
              kb_angular_generic(kind, P.grp_zeta[g][e], P.grp_lambda[g][e], P.grp_eta[g], r, f,
                                 df, phi, dphi);
            const double tot = warp_sum(phi);
            if (lane == 0) zacc[d] += tot;
            if (grad && live)
            {
              TOut* rowp = blk + d * row;
#pragma unroll
              for (int c = 0; c < 3; c++)
              {
                const double pij = dphi[0] * vij[c] / r[0];  // :583-593
                const double pik = dphi[1] * vik[c] / r[1];
                const double pjk = dphi[2] * vjk[c] / r[2];
                atomicAdd(rowp + 3 * n + c, (TOut) (-pij - pik));
                atomicAdd(rowp + 3 * jj + c, (TOut) (pij - pjk));
                atomicAdd(rowp + 3 * kk + c, (TOut) (pik + pjk));
              }
            }
          }
        }
      }
    }
  __syncwarp();
  for (int d = lane; d < D; d += 32) zeta[(int64_t) t * D + d] = zacc[d];
}

}  // namespace

// subset (nullable): device list of centre ordinals to process (n_centers = its length)
int kb_symfun_general_launch(const kb_symfun_params& P, int n_centers, const int32_t* d_centers,
                             const int32_t* d_subset, const double* d_coords,
                             const int32_t* d_species, const int64_t* d_nl_off,
                             const int32_t* d_nl_idx, double* d_zeta, void* d_dz, int dz_dtype,
                             const int64_t* d_dz_ptr, cudaStream_t st)
{
  if (n_centers <= 0) return KB_OK;
  const int nblk = (n_centers + GEN_WARPS - 1) / GEN_WARPS;
  const size_t smem = sizeof(double) * GEN_WARPS * (size_t) P.n_desc;
  if (dz_dtype == KB_F32)
    symfun_general_kernel<float><<<nblk, GEN_WARPS * 32, smem, st>>>(
        P, n_centers, d_centers, d_subset, d_coords, d_species, d_nl_off, d_nl_idx, d_zeta,
        static_cast<float*>(d_dz), d_dz_ptr);
  else
    symfun_general_kernel<double><<<nblk, GEN_WARPS * 32, smem, st>>>(
        P, n_centers, d_centers, d_subset, d_coords, d_species, d_nl_off, d_nl_idx, d_zeta,
        static_cast<double*>(d_dz), d_dz_ptr);
  KB_LAUNCH_CHECK();
  return KB_OK;
}
