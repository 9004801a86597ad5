// Periodic padding ("image") atoms on the GPU.
//
// Replaces nbl_create_paddings (kliff/neighbor/neighbor_list.cpp:283-418).  The selection
// predicates and the image coordinates are evaluated with the reference's operation order and
// without FMA contraction, and the output is an ORDER-PRESERVING compaction (shift-major,
// atom-minor, neighbor_list.cpp:369-383), so coordinates, species, image map and the padding
// numbering are bit-identical to the reference.
//
// Work decomposition: the candidate space is (shift q) x (atom a).  One CTA takes 256
// consecutive atoms of one shift; pass 1 counts survivors per CTA, a device scan turns the
// counts into output offsets, pass 2 re-evaluates the predicate and writes through a ballot
// scan.  Fractional coordinates are computed once (24 B/atom, L2 resident for the re-reads).
#include "common.cuh"

#include <cmath>
#include <new>

struct kb_pad_plan
{
  int32_t natoms;
  int32_t nrep[3];    // shells per axis (ceil(rc / face distance))
  int32_t mdig[3];    // number of shift values per axis (2*nrep+1, or 1 when not periodic)
  double slack[3];    // nrep - rc/dist
  double tc[9];       // transpose(cell): absolute = tc * fractional
  int64_t nshift;     // mdig[0]*mdig[1]*mdig[2], includes the zero shift
  int32_t chunks;     // CTAs per shift
  double* d_frac;     // [3*natoms]
  double* d_lohi;     // lo[3], hi[3] (already widened by TOL)
  int32_t* d_cnt;     // [nshift*chunks] survivors per CTA
  int64_t* d_off;     // exclusive scan of d_cnt, + total at the end
  int64_t npad;
};

namespace {

constexpr int PAD_THREADS = 256;

struct PadGeom
{
  double fc[9];
  double tc[9];
  double slack[3];
  int32_t nrep[3];
  int32_t mdig[3];
};

// ---- host: 3x3 algebra in the reference's operation order (helper.hpp:42-93) --------------
static double h_minor(double p, double q, double r, double s) { return (p * s) - (q * r); }
static int h_inverse(const double* m, double* inv)
{
  inv[0] = h_minor(m[4], m[5], m[7], m[8]);
  inv[1] = h_minor(m[2], m[1], m[8], m[7]);
  inv[2] = h_minor(m[1], m[2], m[4], m[5]);
  inv[3] = h_minor(m[5], m[3], m[8], m[6]);
  inv[4] = h_minor(m[0], m[2], m[6], m[8]);
  inv[5] = h_minor(m[2], m[0], m[5], m[3]);
  inv[6] = h_minor(m[3], m[4], m[6], m[7]);
  inv[7] = h_minor(m[1], m[0], m[7], m[6]);
  inv[8] = h_minor(m[0], m[1], m[3], m[4]);
  volatile double d = m[0] * m[4] * m[8] - m[0] * m[5] * m[7] - m[1] * m[3] * m[8]
                      + m[1] * m[5] * m[6] + m[2] * m[3] * m[7] - m[2] * m[4] * m[6];
  if (std::fabs(d) < KB_TOL) return 1;
  for (int q = 0; q < 9; q++) inv[q] /= d;
  return 0;
}
static void h_cross(const double* a, const double* b, double* c)
{
  c[0] = a[1] * b[2] - a[2] * b[1];
  c[1] = a[2] * b[0] - a[0] * b[2];
  c[2] = a[0] * b[1] - a[1] * b[0];
}
static double h_dot(const double* a, const double* b)
{
  return a[0] * b[0] + a[1] * b[1] + a[2] * b[2];
}

// ---- pass 0: fractional coordinates + per-CTA extrema -------------------------------------
__global__ void __launch_bounds__(PAD_THREADS)
pad_frac_kernel(int natoms, const double* __restrict__ coords, PadGeom g,
                double* __restrict__ frac, double* __restrict__ part)
{
  const int a = blockIdx.x * PAD_THREADS + threadIdx.x;
  double lo[3] = {1e10, 1e10, 1e10}, hi[3] = {-1e10, -1e10, -1e10};  // :305-306
  if (a < natoms)
  {
    const double x = coords[3 * a], y = coords[3 * a + 1], z = coords[3 * a + 2];
#pragma unroll
    for (int c = 0; c < 3; c++)
    {
      const double f = xdot3(g.fc[3 * c], g.fc[3 * c + 1], g.fc[3 * c + 2], x, y, z);  // :312
      frac[3 * a + c] = f;
      lo[c] = f;
      hi[c] = f;
    }
  }
  __shared__ double s[PAD_THREADS / 32][6];
#pragma unroll
  for (int c = 0; c < 3; c++)
  {
#pragma unroll
    for (int m = 16; m > 0; m >>= 1)
    {
      lo[c] = fmin(lo[c], __shfl_xor_sync(0xffffffffu, lo[c], m));
      hi[c] = fmax(hi[c], __shfl_xor_sync(0xffffffffu, hi[c], m));
    }
  }
  if ((threadIdx.x & 31) == 0)
    for (int c = 0; c < 3; c++)
    {
      s[threadIdx.x >> 5][c] = lo[c];
      s[threadIdx.x >> 5][3 + c] = hi[c];
    }
  __syncthreads();
  if (threadIdx.x < 6)
  {
    double v = s[0][threadIdx.x];
    for (int w = 1; w < PAD_THREADS / 32; w++)
      v = (threadIdx.x < 3) ? fmin(v, s[w][threadIdx.x]) : fmax(v, s[w][threadIdx.x]);
    part[6 * blockIdx.x + threadIdx.x] = v;
  }
}

__global__ void __launch_bounds__(256)
pad_lohi_kernel(int nparts, const double* __restrict__ part, double* __restrict__ lohi)
{
  __shared__ double s[256][6];
  double v[6] = {1e10, 1e10, 1e10, -1e10, -1e10, -1e10};
  for (int p = threadIdx.x; p < nparts; p += 256)
    for (int c = 0; c < 6; c++)
      v[c] = (c < 3) ? fmin(v[c], part[6 * p + c]) : fmax(v[c], part[6 * p + c]);
  for (int c = 0; c < 6; c++) s[threadIdx.x][c] = v[c];
  __syncthreads();
  if (threadIdx.x < 6)
  {
    const int c = threadIdx.x;
    double r = s[0][c];
    for (int t = 1; t < 256; t++) r = (c < 3) ? fmin(r, s[t][c]) : fmax(r, s[t][c]);
    lohi[c] = (c < 3) ? xsub(r, KB_TOL) : xadd(r, KB_TOL);  // :329-334
  }
}

// shift index q (lexicographic over the allowed values of each axis) -> (sa, sb, sc)
__device__ __forceinline__ void decode_shift(int64_t q, const PadGeom& g, int* s)
{
  const int c2 = (int) (q % g.mdig[2]);
  q /= g.mdig[2];
  const int c1 = (int) (q % g.mdig[1]);
  const int c0 = (int) (q / g.mdig[1]);
  s[0] = (g.mdig[0] == 1) ? 0 : c0 - g.nrep[0];
  s[1] = (g.mdig[1] == 1) ? 0 : c1 - g.nrep[1];
  s[2] = (g.mdig[2] == 1) ? 0 : c2 - g.nrep[2];
}

// keep-predicate of neighbor_list.cpp:392-397 for one atom and one shift
__device__ __forceinline__ bool pad_keep(const int* s, const PadGeom& g, const double* f,
                                         const double* lohi)
{
  bool keep = true;
#pragma unroll
  for (int c = 0; c < 3; c++)
  {
    if (s[c] == -g.nrep[c] && xsub(f[c], lohi[c]) < g.slack[c]) keep = false;
    if (s[c] == g.nrep[c] && xsub(lohi[3 + c], f[c]) < g.slack[c]) keep = false;
  }
  return keep;
}

template <bool FILL>
__global__ void __launch_bounds__(PAD_THREADS)
pad_select_kernel(int natoms, PadGeom g, const double* __restrict__ frac,
                  const double* __restrict__ lohi, int32_t* __restrict__ cnt,
                  const int64_t* __restrict__ off, const int32_t* __restrict__ species,
                  double* __restrict__ pad_coords, int32_t* __restrict__ pad_species,
                  int32_t* __restrict__ pad_image)
{
  const int64_t q = blockIdx.y;
  int s[3];
  decode_shift(q, g, s);
  const bool zero = (s[0] == 0 && s[1] == 0 && s[2] == 0);  // :376
  const int a = blockIdx.x * PAD_THREADS + threadIdx.x;
  double f[3] = {0, 0, 0};
  bool keep = false;
  if (!zero && a < natoms)
  {
    f[0] = frac[3 * a];
    f[1] = frac[3 * a + 1];
    f[2] = frac[3 * a + 2];
    double lh[6];
#pragma unroll
    for (int c = 0; c < 6; c++) lh[c] = lohi[c];
    keep = pad_keep(s, g, f, lh);
  }
  const unsigned bal = __ballot_sync(0xffffffffu, keep);
  __shared__ int wtot[PAD_THREADS / 32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) wtot[warp] = __popc(bal);
  __syncthreads();
  int before = 0, total = 0;
#pragma unroll
  for (int w = 0; w < PAD_THREADS / 32; w++)
  {
    if (w < warp) before += wtot[w];
    total += wtot[w];
  }
  const int64_t blk = q * gridDim.x + blockIdx.x;
  if (!FILL)
  {
    if (threadIdx.x == 0) cnt[blk] = total;
    return;
  }
  if (keep)
  {
    const int64_t o = off[blk] + before + __popc(bal & ((1u << lane) - 1u));
    const double g0 = xadd((double) s[0], f[0]);  // :400  {i + x, j + y, k + z}
    const double g1 = xadd((double) s[1], f[1]);
    const double g2 = xadd((double) s[2], f[2]);
    pad_coords[3 * o + 0] = xdot3(g.tc[0], g.tc[1], g.tc[2], g0, g1, g2);  // :403-405
    pad_coords[3 * o + 1] = xdot3(g.tc[3], g.tc[4], g.tc[5], g0, g1, g2);
    pad_coords[3 * o + 2] = xdot3(g.tc[6], g.tc[7], g.tc[8], g0, g1, g2);
    if (pad_species) pad_species[o] = species[a];
    pad_image[o] = a;
  }
}

static PadGeom geom_of(const kb_pad_plan* p, const double* fc)
{
  PadGeom g;
  for (int q = 0; q < 9; q++)
  {
    g.tc[q] = p->tc[q];
    g.fc[q] = fc ? fc[q] : 0.0;
  }
  for (int c = 0; c < 3; c++)
  {
    g.slack[c] = p->slack[c];
    g.nrep[c] = p->nrep[c];
    g.mdig[c] = p->mdig[c];
  }
  return g;
}

}  // namespace

extern "C" void kb_pad_destroy(kb_pad_plan* plan, void* stream)
{
  if (!plan) return;
  cudaStream_t st = kb_stream(stream);
  if (plan->d_frac) cudaFreeAsync(plan->d_frac, st);
  if (plan->d_lohi) cudaFreeAsync(plan->d_lohi, st);
  if (plan->d_cnt) cudaFreeAsync(plan->d_cnt, st);
  if (plan->d_off) cudaFreeAsync(plan->d_off, st);
  delete plan;
}

extern "C" int kb_pad_begin(int32_t natoms, double cutoff, const double* h_cell,
                            const int32_t* h_pbc, const double* d_coords, int64_t* h_npad,
                            kb_pad_plan** out_plan, void* stream)
{
  if (!h_cell || !h_pbc || !h_npad || !out_plan || natoms < 0) return KB_ERR_INVALID;
  *out_plan = nullptr;
  *h_npad = 0;
  cudaStream_t st = kb_stream(stream);
  kb_pool_setup();

  // lattice analysis on the host, same operation order as neighbor_list.cpp:291-365
  double tc[9], fc[9];
  for (int r = 0; r < 3; r++)
    for (int c = 0; c < 3; c++) tc[3 * r + c] = h_cell[3 * c + r];
  if (h_inverse(tc, fc)) return KB_ERR_REFERENCE;
  double x[3], dist[3];
  h_cross(h_cell + 3, h_cell + 6, x);
  const double vol = std::fabs(h_dot(h_cell, x));
  dist[0] = vol / std::sqrt(x[0] * x[0] + x[1] * x[1] + x[2] * x[2]);
  h_cross(h_cell + 6, h_cell + 0, x);
  dist[1] = vol / std::sqrt(x[0] * x[0] + x[1] * x[1] + x[2] * x[2]);
  h_cross(h_cell, h_cell + 3, x);
  dist[2] = vol / std::sqrt(x[0] * x[0] + x[1] * x[1] + x[2] * x[2]);

  kb_pad_plan* p = new (std::nothrow) kb_pad_plan();
  if (!p) return KB_ERR_INVALID;
  p->natoms = natoms;
  p->d_frac = nullptr; p->d_lohi = nullptr; p->d_cnt = nullptr; p->d_off = nullptr;
  p->npad = 0;
  p->nshift = 1;
  for (int c = 0; c < 3; c++)
  {
    const double ratio = cutoff / dist[c];
    const double up = std::ceil(ratio);
    if (!(up < 1.0e6)) { delete p; return KB_ERR_LIMIT; }
    p->nrep[c] = (int) up;
    p->slack[c] = (double) p->nrep[c] - ratio;
    p->mdig[c] = h_pbc[c] ? 2 * p->nrep[c] + 1 : 1;  // non-periodic axis: shift 0 only (:379-381)
    p->nshift *= p->mdig[c];
  }
  for (int q = 0; q < 9; q++) p->tc[q] = tc[q];
  if (natoms == 0 || p->nshift == 1) { *out_plan = p; return KB_OK; }
  if (p->nshift > 65535) { delete p; return KB_ERR_LIMIT; }

  p->chunks = (natoms + PAD_THREADS - 1) / PAD_THREADS;
  const int64_t nblk = p->nshift * p->chunks;
  int rc = KB_OK;
  double* d_part = nullptr;
#define PAD_TRY(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { rc = -(int) e_; goto fail; } } while (0)
  PAD_TRY(cudaMallocAsync((void**) &p->d_frac, sizeof(double) * 3 * (size_t) natoms, st));
  PAD_TRY(cudaMallocAsync((void**) &p->d_lohi, sizeof(double) * 6, st));
  PAD_TRY(cudaMallocAsync((void**) &p->d_cnt, sizeof(int32_t) * (size_t) nblk, st));
  PAD_TRY(cudaMallocAsync((void**) &p->d_off, sizeof(int64_t) * ((size_t) nblk + 1), st));
  PAD_TRY(cudaMallocAsync((void**) &d_part, sizeof(double) * 6 * (size_t) p->chunks, st));
  {
    PadGeom g = geom_of(p, fc);
    pad_frac_kernel<<<p->chunks, PAD_THREADS, 0, st>>>(natoms, d_coords, g, p->d_frac, d_part);
    PAD_TRY(cudaGetLastError()); kb_note_launch();
    pad_lohi_kernel<<<1, 256, 0, st>>>(p->chunks, d_part, p->d_lohi);
    PAD_TRY(cudaGetLastError()); kb_note_launch();
    dim3 grid(p->chunks, (unsigned) p->nshift);
    pad_select_kernel<false><<<grid, PAD_THREADS, 0, st>>>(natoms, g, p->d_frac, p->d_lohi,
                                                           p->d_cnt, nullptr, nullptr, nullptr,
                                                           nullptr, nullptr);
    PAD_TRY(cudaGetLastError()); kb_note_launch();
    rc = kb_exclusive_scan<int32_t, int64_t>(p->d_cnt, p->d_off, nblk, p->d_off + nblk, st);
    if (rc) goto fail;
    PAD_TRY(cudaMemcpyAsync(&p->npad, p->d_off + nblk, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
    PAD_TRY(cudaFreeAsync(d_part, st));
    d_part = nullptr;
    PAD_TRY(cudaStreamSynchronize(st));
  }
#undef PAD_TRY
  *h_npad = p->npad;
  *out_plan = p;
  return KB_OK;
fail:
  if (d_part) cudaFreeAsync(d_part, st);
  kb_pad_destroy(p, stream);
  return rc;
}

extern "C" int kb_pad_finish(kb_pad_plan* p, const int32_t* d_species, double* d_pad_coords,
                             int32_t* d_pad_species, int32_t* d_pad_image, void* stream)
{
  if (!p) return KB_ERR_INVALID;
  cudaStream_t st = kb_stream(stream);
  int rc = KB_OK;
  if (p->npad > 0)
  {
    if (!d_pad_coords || !d_pad_image || (d_pad_species && !d_species)) rc = KB_ERR_INVALID;
    else
    {
      PadGeom g = geom_of(p, nullptr);
      dim3 grid(p->chunks, (unsigned) p->nshift);
      pad_select_kernel<true><<<grid, PAD_THREADS, 0, st>>>(p->natoms, g, p->d_frac, p->d_lohi,
                                                            nullptr, p->d_off, d_species,
                                                            d_pad_coords, d_pad_species,
                                                            d_pad_image);
      cudaError_t e = cudaGetLastError();
      if (e != cudaSuccess) rc = -(int) e;
      else kb_note_launch();
    }
  }
  kb_pad_destroy(p, stream);
  return rc;
}
