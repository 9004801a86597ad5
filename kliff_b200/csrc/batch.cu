// Batched padding + neighbor-list build for MANY small configurations in one call.
//
// A training set is thousands of 8-64-atom cells (Si_training_set: 400, the synthetic epoch set:
// 100 000).  Driving pad.cu / neighbor.cu once per configuration costs ~12 launches and 3 host
// synchronisations each (measured 0.6-0.9 ms per configuration, all of it host overhead).  Here
// one CTA owns one configuration and the whole batch takes 4 launches + 2 scans + 2 synchronisations:
//
//   K1  per configuration: lattice analysis ON THE DEVICE in the reference's operation order
//       (neighbor_list.cpp:291-365, helper.hpp:42-93; every operation is an explicit IEEE
//       __dmul_rn/__dadd_rn/__ddiv_rn/__dsqrt_rn so no FMA contraction can change a bit),
//       fractional coordinates, exact extrema, number of surviving image atoms
//   K2  order-preserving fill of the image atoms (shift-major, atom-minor, :369-383), straight
//       into the packed batch layout: per configuration its contributing atoms, then its images
//   K3  neighbor counts: every contributing atom against all atoms of its configuration
//       (all-pairs inside a configuration: <= a few hundred atoms, so no cell grid is needed; the
//       pair test is the reference's rsq < cutoff^2 on the exactly evaluated rsq, :197, :212)
//   K4  neighbor fill, ascending (the canonical order) by construction
//
// Results are bit-identical to the per-configuration path (tests/test_gpu_parity.py).
#include "common.cuh"

#include <new>

struct kb_batch_plan
{
  int32_t n_cfg;
  int64_t n_contrib;
  double cutoff;
  const double* d_coords;    // [n_contrib][3]
  int32_t* d_cfg_ptr;        // [n_cfg+1] contributing ranges
  double* d_cells;           // [n_cfg][9]
  int32_t* d_pbc;            // [n_cfg][3]
  double* d_frac;            // [n_contrib][3]
  double* d_geom;            // [n_cfg][GEOM_DOUBLES]
  int32_t* d_npad;           // [n_cfg]
  int64_t* d_atom_ptr;       // [n_cfg+1] atom ranges (contributing + images)
  int32_t* d_err;            // [0] singular cell, [1] collision, [2] max neighbor count
  int64_t total_atoms;
  int64_t total_pairs;
};

namespace {

constexpr int BT = 128;              // threads per CTA
constexpr int GEOM_DOUBLES = 9 + 3 + 6 + 6;  // tc[9], slack[3], lohi[6], (nrep[3], mdig[3]) as doubles

__device__ __forceinline__ double minor2(double p, double q, double r, double s)
{
  return xsub(xmul(p, s), xmul(q, r));
}

// transpose(cell), its inverse, face distances, shells and slack: host code of kb_pad_begin, on device
// returns 0, or 1 = singular cell (KB_ERR_REFERENCE), 2 = too many shells (KB_ERR_LIMIT)
__device__ int lattice(const double* cell, const int* pbc, double cutoff, double* tc, double* fc,
                       double* slack, int* nrep, int* mdig)
{
  for (int r = 0; r < 3; r++)
    for (int c = 0; c < 3; c++) tc[3 * r + c] = cell[3 * c + r];
  const double* m = tc;
  fc[0] = minor2(m[4], m[5], m[7], m[8]);
  fc[1] = minor2(m[2], m[1], m[8], m[7]);
  fc[2] = minor2(m[1], m[2], m[4], m[5]);
  fc[3] = minor2(m[5], m[3], m[8], m[6]);
  fc[4] = minor2(m[0], m[2], m[6], m[8]);
  fc[5] = minor2(m[2], m[0], m[5], m[3]);
  fc[6] = minor2(m[3], m[4], m[6], m[7]);
  fc[7] = minor2(m[1], m[0], m[7], m[6]);
  fc[8] = minor2(m[0], m[1], m[3], m[4]);
  // det: ((((t0 - t1) - t2) + t3) + t4) - t5, helper.hpp:42-48
  double d = xmul(xmul(m[0], m[4]), m[8]);
  d = xsub(d, xmul(xmul(m[0], m[5]), m[7]));
  d = xsub(d, xmul(xmul(m[1], m[3]), m[8]));
  d = xadd(d, xmul(xmul(m[1], m[5]), m[6]));
  d = xadd(d, xmul(xmul(m[2], m[3]), m[7]));
  d = xsub(d, xmul(xmul(m[2], m[4]), m[6]));
  if (fabs(d) < KB_TOL) return 1;
  for (int q = 0; q < 9; q++) fc[q] = __ddiv_rn(fc[q], d);
  // volume and face distances, neighbor_list.cpp:337-353
  const double* a = cell;
  const double* b = cell + 3;
  const double* c = cell + 6;
  double x[3];
  auto cross = [](const double* u, const double* v, double* o) {
    o[0] = xsub(xmul(u[1], v[2]), xmul(u[2], v[1]));
    o[1] = xsub(xmul(u[2], v[0]), xmul(u[0], v[2]));
    o[2] = xsub(xmul(u[0], v[1]), xmul(u[1], v[0]));
  };
  cross(b, c, x);
  const double vol = fabs(xdot3(a[0], a[1], a[2], x[0], x[1], x[2]));
  double dist[3];
  dist[0] = __ddiv_rn(vol, __dsqrt_rn(xdot3(x[0], x[1], x[2], x[0], x[1], x[2])));
  cross(c, a, x);
  dist[1] = __ddiv_rn(vol, __dsqrt_rn(xdot3(x[0], x[1], x[2], x[0], x[1], x[2])));
  cross(a, b, x);
  dist[2] = __ddiv_rn(vol, __dsqrt_rn(xdot3(x[0], x[1], x[2], x[0], x[1], x[2])));
  for (int q = 0; q < 3; q++)
  {
    const double ratio = __ddiv_rn(cutoff, dist[q]);
    const double up = ceil(ratio);
    if (!(up < 1.0e6)) return 2;
    nrep[q] = (int) up;
    slack[q] = xsub((double) nrep[q], ratio);
    mdig[q] = pbc[q] ? 2 * nrep[q] + 1 : 1;
  }
  if ((int64_t) mdig[0] * mdig[1] * mdig[2] > 65535) return 2;  // same limit as kb_pad_begin
  return 0;
}

__device__ __forceinline__ bool keep_image(const int* s, const int* nrep, const double* slack,
                                           const double* f, const double* lohi)
{
  bool keep = true;
#pragma unroll
  for (int c = 0; c < 3; c++)
  {
    if (s[c] == -nrep[c] && xsub(f[c], lohi[c]) < slack[c]) keep = false;   // :392-397
    if (s[c] == nrep[c] && xsub(lohi[3 + c], f[c]) < slack[c]) keep = false;
  }
  return keep;
}

__device__ __forceinline__ void shift_of(int q, const int* nrep, const int* mdig, int* s)
{
  const int c2 = q % mdig[2];
  q /= mdig[2];
  const int c1 = q % mdig[1];
  const int c0 = q / mdig[1];
  s[0] = mdig[0] == 1 ? 0 : c0 - nrep[0];
  s[1] = mdig[1] == 1 ? 0 : c1 - nrep[1];
  s[2] = mdig[2] == 1 ? 0 : c2 - nrep[2];
}

// block-wide helpers (BT threads)
__device__ double block_minmax(double v, bool want_min, double* sh)
{
#pragma unroll
  for (int m = 16; m > 0; m >>= 1)
  {
    const double o = __shfl_xor_sync(0xffffffffu, v, m);
    v = want_min ? fmin(v, o) : fmax(v, o);
  }
  __syncthreads();
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  double r = sh[0];
  for (int w = 1; w < BT / 32; w++) r = want_min ? fmin(r, sh[w]) : fmax(r, sh[w]);
  return r;
}

template <bool FILL>
__global__ void __launch_bounds__(BT)
batch_pad_kernel(int n_cfg, const int32_t* __restrict__ cfg_ptr, const double* __restrict__ cells,
                 const int32_t* __restrict__ pbc_all, double cutoff, const double* __restrict__ coords,
                 double* __restrict__ frac, double* __restrict__ geom, int32_t* __restrict__ npad,
                 const int64_t* __restrict__ atom_ptr, double* __restrict__ coords_all,
                 int32_t* __restrict__ image, int32_t* __restrict__ centers, int32_t* __restrict__ err)
{
  __shared__ double sh[BT / 32];
  __shared__ double s_geom[GEOM_DOUBLES + 9];
  __shared__ int s_int[8];
  __shared__ int s_wtot[BT / 32];
  const int c = blockIdx.x;
  const int a0 = cfg_ptr[c], n = cfg_ptr[c + 1] - a0;
  double* tc = s_geom;          // 9
  double* slack = s_geom + 9;   // 3
  double* lohi = s_geom + 12;   // 6
  double* fcm = s_geom + GEOM_DOUBLES;  // 9 (inverse), K1 only
  int* nrep = s_int;
  int* mdig = s_int + 3;

  if (!FILL)
  {
    if (threadIdx.x == 0)
    {
      int pbc[3] = {pbc_all[3 * c], pbc_all[3 * c + 1], pbc_all[3 * c + 2]};
      const int bad = lattice(cells + 9 * c, pbc, cutoff, tc, fcm, slack, nrep, mdig);
      s_int[6] = bad ? 0 : 1;
      if (bad) atomicMax(err, bad);
    }
    __syncthreads();
    if (!s_int[6]) { if (threadIdx.x == 0) npad[c] = 0; return; }
    double lo[3] = {1e10, 1e10, 1e10}, hi[3] = {-1e10, -1e10, -1e10};  // :305-306
    for (int a = threadIdx.x; a < n; a += BT)
    {
      const double x = coords[3 * (a0 + a)], y = coords[3 * (a0 + a) + 1], z = coords[3 * (a0 + a) + 2];
#pragma unroll
      for (int q = 0; q < 3; q++)
      {
        const double f = xdot3(fcm[3 * q], fcm[3 * q + 1], fcm[3 * q + 2], x, y, z);  // :312
        frac[3 * (a0 + a) + q] = f;
        lo[q] = fmin(lo[q], f);
        hi[q] = fmax(hi[q], f);
      }
    }
    for (int q = 0; q < 3; q++)
    {
      const double l = block_minmax(lo[q], true, sh);
      const double h = block_minmax(hi[q], false, sh);
      if (threadIdx.x == 0) { lohi[q] = xsub(l, KB_TOL); lohi[3 + q] = xadd(h, KB_TOL); }  // :329-334
    }
    __syncthreads();
    if (threadIdx.x < GEOM_DOUBLES - 6) geom[c * GEOM_DOUBLES + threadIdx.x] = s_geom[threadIdx.x];
    if (threadIdx.x < 6) geom[c * GEOM_DOUBLES + 18 + threadIdx.x] = (double) s_int[threadIdx.x];
  }
  else
  {
    if (threadIdx.x < 18) s_geom[threadIdx.x] = geom[c * GEOM_DOUBLES + threadIdx.x];
    if (threadIdx.x < 6) s_int[threadIdx.x] = (int) geom[c * GEOM_DOUBLES + 18 + threadIdx.x];
    __syncthreads();
    // contributing atoms first
    const int64_t base = atom_ptr[c];
    for (int a = threadIdx.x; a < n; a += BT)
    {
      coords_all[3 * (base + a)] = coords[3 * (a0 + a)];
      coords_all[3 * (base + a) + 1] = coords[3 * (a0 + a) + 1];
      coords_all[3 * (base + a) + 2] = coords[3 * (a0 + a) + 2];
      image[base + a] = a0 + a;
      centers[a0 + a] = (int32_t) (base + a);
    }
  }
  __syncthreads();

  const int nshift = mdig[0] * mdig[1] * mdig[2];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int running = 0;  // images emitted so far (identical in every thread)
  for (int q = 0; q < nshift; q++)
  {
    int s[3];
    shift_of(q, nrep, mdig, s);
    if (s[0] == 0 && s[1] == 0 && s[2] == 0) continue;  // :376 (uniform)
    for (int ab = 0; ab < n; ab += BT)
    {
      const int a = ab + threadIdx.x;
      bool keep = false;
      double f[3] = {0, 0, 0};
      if (a < n)
      {
        f[0] = frac[3 * (a0 + a)]; f[1] = frac[3 * (a0 + a) + 1]; f[2] = frac[3 * (a0 + a) + 2];
        keep = keep_image(s, nrep, slack, f, lohi);
      }
      const unsigned bal = __ballot_sync(0xffffffffu, keep);
      __syncthreads();
      if (lane == 0) s_wtot[warp] = __popc(bal);
      __syncthreads();
      int before = 0, total = 0;
#pragma unroll
      for (int w = 0; w < BT / 32; w++)
      {
        if (w < warp) before += s_wtot[w];
        total += s_wtot[w];
      }
      if (FILL && keep)
      {
        const int64_t o = atom_ptr[c] + n + running + before + __popc(bal & ((1u << lane) - 1u));
        const double g0 = xadd((double) s[0], f[0]), g1 = xadd((double) s[1], f[1]),
                     g2 = xadd((double) s[2], f[2]);  // :400
        coords_all[3 * o] = xdot3(tc[0], tc[1], tc[2], g0, g1, g2);  // :403-405
        coords_all[3 * o + 1] = xdot3(tc[3], tc[4], tc[5], g0, g1, g2);
        coords_all[3 * o + 2] = xdot3(tc[6], tc[7], tc[8], g0, g1, g2);
        image[o] = a0 + a;
      }
      running += total;
    }
  }
  if (!FILL && threadIdx.x == 0) npad[c] = running;
}

__global__ void atom_count_kernel(int n_cfg, const int32_t* __restrict__ cfg_ptr,
                                  const int32_t* __restrict__ npad, int64_t* __restrict__ natoms)
{
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < n_cfg) natoms[c] = (int64_t) (cfg_ptr[c + 1] - cfg_ptr[c]) + npad[c];
}

// neighbors of contributing atoms: all pairs inside the configuration, ascending index order
template <bool FILL>
__global__ void __launch_bounds__(BT)
batch_nl_kernel(int n_cfg, const int32_t* __restrict__ cfg_ptr, const int64_t* __restrict__ atom_ptr,
                const double* __restrict__ coords_all, double cutsq, int32_t* __restrict__ counts,
                const int64_t* __restrict__ offsets, int32_t* __restrict__ indices,
                int32_t* __restrict__ err)
{
  const int c = blockIdx.x;
  const int n = cfg_ptr[c + 1] - cfg_ptr[c];
  const int64_t base = atom_ptr[c];
  const int na = (int) (atom_ptr[c + 1] - base);
  if (!FILL)  // image atoms have empty lists
    for (int m = n + threadIdx.x; m < na; m += BT) counts[base + m] = 0;
  for (int i = threadIdx.x; i < n; i += BT)
  {
    const double px = coords_all[3 * (base + i)], py = coords_all[3 * (base + i) + 1],
                 pz = coords_all[3 * (base + i) + 2];
    int cnt = 0;
    bool clash = false;
    int32_t* out = FILL ? indices + offsets[base + i] : nullptr;
    for (int m = 0; m < na; m++)
    {
      if (m == i) continue;
      const double rsq = xnorm2(xsub(coords_all[3 * (base + m)], px), xsub(coords_all[3 * (base + m) + 1], py),
                                xsub(coords_all[3 * (base + m) + 2], pz));
      if (rsq < KB_TOL) clash = true;  // neighbor_list.cpp:199
      if (rsq < cutsq)                 // :212
      {
        if (FILL) out[cnt] = (int32_t) (base + m);
        cnt++;
      }
    }
    if (!FILL)
    {
      counts[base + i] = cnt;
      if (clash) atomicExch(err + 1, 1);
      atomicMax(err + 2, cnt);
    }
  }
}

}  // namespace

extern "C" void kb_batch_destroy(kb_batch_plan* p, void* stream)
{
  if (!p) return;
  cudaStream_t st = kb_stream(stream);
  void* bufs[] = {p->d_cfg_ptr, p->d_cells, p->d_pbc, p->d_frac, p->d_geom, p->d_npad, p->d_atom_ptr, p->d_err};
  for (void* b : bufs)
    if (b) cudaFreeAsync(b, st);
  delete p;
}

extern "C" int kb_batch_begin(int32_t n_cfg, const int32_t* h_cfg_ptr, const double* h_cells,
                              const int32_t* h_pbc, double cutoff, const double* d_coords,
                              int64_t* h_total_atoms, kb_batch_plan** out_plan, void* stream)
{
  if (n_cfg < 0 || !h_cfg_ptr || !h_cells || !h_pbc || !h_total_atoms || !out_plan) return KB_ERR_INVALID;
  *out_plan = nullptr;
  *h_total_atoms = 0;
  cudaStream_t st = kb_stream(stream);
  kb_pool_setup();
  kb_batch_plan* p = new (std::nothrow) kb_batch_plan();
  if (!p) return KB_ERR_INVALID;
  p->n_cfg = n_cfg; p->n_contrib = n_cfg ? h_cfg_ptr[n_cfg] : 0; p->cutoff = cutoff; p->d_coords = d_coords;
  p->d_cfg_ptr = nullptr; p->d_cells = nullptr; p->d_pbc = nullptr; p->d_frac = nullptr;
  p->d_geom = nullptr; p->d_npad = nullptr; p->d_atom_ptr = nullptr; p->d_err = nullptr;
  p->total_atoms = 0; p->total_pairs = 0;
  if (n_cfg == 0) { *out_plan = p; return KB_OK; }
  int rc = KB_OK;
  int64_t* d_natoms = nullptr;
  int32_t h_err[3] = {0, 0, 0};
#define B_TRY(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { rc = -(int) e_; goto fail; } } while (0)
  B_TRY(cudaMallocAsync((void**) &p->d_cfg_ptr, sizeof(int32_t) * ((size_t) n_cfg + 1), st));
  B_TRY(cudaMallocAsync((void**) &p->d_cells, sizeof(double) * 9 * (size_t) n_cfg, st));
  B_TRY(cudaMallocAsync((void**) &p->d_pbc, sizeof(int32_t) * 3 * (size_t) n_cfg, st));
  B_TRY(cudaMallocAsync((void**) &p->d_frac, sizeof(double) * 3 * (size_t) (p->n_contrib + 1), st));
  B_TRY(cudaMallocAsync((void**) &p->d_geom, sizeof(double) * GEOM_DOUBLES * (size_t) n_cfg, st));
  B_TRY(cudaMallocAsync((void**) &p->d_npad, sizeof(int32_t) * (size_t) n_cfg, st));
  B_TRY(cudaMallocAsync((void**) &p->d_atom_ptr, sizeof(int64_t) * ((size_t) n_cfg + 1), st));
  B_TRY(cudaMallocAsync((void**) &p->d_err, sizeof(int32_t) * 4, st));
  B_TRY(cudaMallocAsync((void**) &d_natoms, sizeof(int64_t) * (size_t) n_cfg, st));
  B_TRY(cudaMemsetAsync(p->d_err, 0, sizeof(int32_t) * 4, st));
  B_TRY(cudaMemcpyAsync(p->d_cfg_ptr, h_cfg_ptr, sizeof(int32_t) * ((size_t) n_cfg + 1), cudaMemcpyHostToDevice, st));
  B_TRY(cudaMemcpyAsync(p->d_cells, h_cells, sizeof(double) * 9 * (size_t) n_cfg, cudaMemcpyHostToDevice, st));
  B_TRY(cudaMemcpyAsync(p->d_pbc, h_pbc, sizeof(int32_t) * 3 * (size_t) n_cfg, cudaMemcpyHostToDevice, st));
  batch_pad_kernel<false><<<n_cfg, BT, 0, st>>>(n_cfg, p->d_cfg_ptr, p->d_cells, p->d_pbc, cutoff, d_coords,
                                                p->d_frac, p->d_geom, p->d_npad, nullptr, nullptr, nullptr,
                                                nullptr, p->d_err);
  B_TRY(cudaGetLastError()); kb_note_launch();
  atom_count_kernel<<<(n_cfg + 255) / 256, 256, 0, st>>>(n_cfg, p->d_cfg_ptr, p->d_npad, d_natoms);
  B_TRY(cudaGetLastError()); kb_note_launch();
  rc = kb_exclusive_scan<int64_t, int64_t>(d_natoms, p->d_atom_ptr, n_cfg, p->d_atom_ptr + n_cfg, st);
  if (rc) goto fail;
  B_TRY(cudaMemcpyAsync(&p->total_atoms, p->d_atom_ptr + n_cfg, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
  B_TRY(cudaMemcpyAsync(h_err, p->d_err, sizeof(int32_t) * 3, cudaMemcpyDeviceToHost, st));
  B_TRY(cudaFreeAsync(d_natoms, st)); d_natoms = nullptr;
  B_TRY(cudaStreamSynchronize(st));
#undef B_TRY
  if (h_err[0]) { rc = h_err[0] == 1 ? KB_ERR_REFERENCE : KB_ERR_LIMIT; goto fail; }
  *h_total_atoms = p->total_atoms;
  *out_plan = p;
  return KB_OK;
fail:
  if (d_natoms) cudaFreeAsync(d_natoms, st);
  kb_batch_destroy(p, stream);
  return rc;
}

extern "C" int kb_batch_neighbors(kb_batch_plan* p, double* d_coords_all, int32_t* d_image,
                                  int32_t* d_centers, int64_t* d_atom_ptr_out, int32_t* d_counts,
                                  int64_t* d_offsets, int64_t* h_total_pairs, int32_t* h_maxn, void* stream)
{
  if (!p || !h_total_pairs || !h_maxn) return KB_ERR_INVALID;
  cudaStream_t st = kb_stream(stream);
  *h_total_pairs = 0;
  *h_maxn = 0;
  if (p->n_cfg == 0) return KB_OK;
  if (p->total_atoms == 0)
  {
    KB_CUDA(cudaMemcpyAsync(d_atom_ptr_out, p->d_atom_ptr, sizeof(int64_t) * ((size_t) p->n_cfg + 1),
                            cudaMemcpyDeviceToDevice, st));
    return KB_OK;
  }
  int32_t h_err[3] = {0, 0, 0};
  batch_pad_kernel<true><<<p->n_cfg, BT, 0, st>>>(p->n_cfg, p->d_cfg_ptr, p->d_cells, p->d_pbc, p->cutoff,
                                                  p->d_coords, p->d_frac, p->d_geom, p->d_npad, p->d_atom_ptr,
                                                  d_coords_all, d_image, d_centers, p->d_err);
  KB_LAUNCH_CHECK();
  const double cutsq = p->cutoff * p->cutoff;
  batch_nl_kernel<false><<<p->n_cfg, BT, 0, st>>>(p->n_cfg, p->d_cfg_ptr, p->d_atom_ptr, d_coords_all, cutsq,
                                                  d_counts, nullptr, nullptr, p->d_err);
  KB_LAUNCH_CHECK();
  int rc = kb_exclusive_scan<int32_t, int64_t>(d_counts, d_offsets, p->total_atoms, d_offsets + p->total_atoms, st);
  if (rc) return rc;
  KB_CUDA(cudaMemcpyAsync(d_atom_ptr_out, p->d_atom_ptr, sizeof(int64_t) * ((size_t) p->n_cfg + 1),
                          cudaMemcpyDeviceToDevice, st));
  KB_CUDA(cudaMemcpyAsync(&p->total_pairs, d_offsets + p->total_atoms, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
  KB_CUDA(cudaMemcpyAsync(h_err, p->d_err, sizeof(int32_t) * 3, cudaMemcpyDeviceToHost, st));
  KB_CUDA(cudaStreamSynchronize(st));
  if (h_err[1]) return KB_ERR_REFERENCE;  // collision
  *h_total_pairs = p->total_pairs;
  *h_maxn = h_err[2];
  return KB_OK;
}

extern "C" int kb_batch_finish(kb_batch_plan* p, const double* d_coords_all, const int64_t* d_offsets,
                               int32_t* d_indices, void* stream)
{
  if (!p) return KB_ERR_INVALID;
  cudaStream_t st = kb_stream(stream);
  int rc = KB_OK;
  if (p->n_cfg > 0 && p->total_pairs > 0)
  {
    if (!d_indices) rc = KB_ERR_INVALID;
    else
    {
      batch_nl_kernel<true><<<p->n_cfg, BT, 0, st>>>(p->n_cfg, p->d_cfg_ptr, p->d_atom_ptr, d_coords_all,
                                                     p->cutoff * p->cutoff, nullptr, d_offsets, d_indices,
                                                     p->d_err);
      cudaError_t e = cudaGetLastError();
      if (e != cudaSuccess) rc = -(int) e;
      else kb_note_launch();
    }
  }
  kb_batch_destroy(p, stream);
  return rc;
}
