// Cell-list neighbor build on the GPU.
//
// Replaces nbl_build / nbl_get_neigh (kliff/neighbor/neighbor_list.cpp:78-280).  Binning uses
// the reference's own formula (helper.hpp:96-108, including its `max = first atom + 1` start,
// neighbor_list.cpp:90-96) and the pair test is the reference's `rsq < cutoff^2` on
// dx*dx+dy*dy+dz*dz evaluated without FMA (neighbor_list.cpp:197, :212), so the neighbor SET of
// every atom is bit-identical.  Lists are emitted sorted ascending ("canonical sorting"): the
// reference's cell-scan order is an artefact of its bins and nothing downstream depends on it.
//
// Pipeline (all on `stream`): bbox reduce -> bin + histogram -> scan -> cell-ordered copy of
// the coordinates (coalesced candidate reads) -> count pass -> scan (CSR offsets, int64) ->
// fill pass, one warp per atom: ballot compaction of the candidates into shared memory, rank
// sort, coalesced write.
#include "common.cuh"

#include <cmath>
#include <new>

struct kb_nl_plan
{
  int32_t ntot;
  int32_t n_need;
  int32_t half;
  int32_t nb[3];
  double lo[3], hi[3];
  double cutsq;
  const double* d_coords;
  const int32_t* d_need;
  int32_t* d_counts;
  int64_t* d_offsets;
  int32_t* d_cellstart;  // [ncell+1]
  int32_t* d_sorted_id;  // [ntot] atom index by cell-major position
  double* d_sorted_xyz;  // [3*ntot]
  int32_t* d_err;        // collision flag
  int64_t total;
  int32_t maxn;
};

namespace {

constexpr int NL_THREADS = 256;

struct Grid3
{
  double lo[3], hi[3];
  int32_t nb[3];
};

// helper.hpp:96-108 (division and multiplication are IEEE; no contraction possible here)
__device__ __forceinline__ void bin_of(const double* p, const Grid3& g, int* b)
{
#pragma unroll
  for (int c = 0; c < 3; c++)
  {
    const double t = xmul(xsub(p[c], g.lo[c]) / xsub(g.hi[c], g.lo[c]), (double) g.nb[c]);
    int v = (int) t;
    b[c] = v < g.nb[c] - 1 ? v : g.nb[c] - 1;
  }
}

__global__ void __launch_bounds__(NL_THREADS)
bbox_kernel(int ntot, const double* __restrict__ coords, double* __restrict__ part)
{
  double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
  for (int a = blockIdx.x * NL_THREADS + threadIdx.x; a < ntot; a += gridDim.x * NL_THREADS)
#pragma unroll
    for (int c = 0; c < 3; c++)
    {
      const double v = coords[3 * a + c];
      lo[c] = fmin(lo[c], v);
      hi[c] = fmax(hi[c], v);
    }
  __shared__ double s[NL_THREADS / 32][6];
#pragma unroll
  for (int c = 0; c < 3; c++)
#pragma unroll
    for (int m = 16; m > 0; m >>= 1)
    {
      lo[c] = fmin(lo[c], __shfl_xor_sync(0xffffffffu, lo[c], m));
      hi[c] = fmax(hi[c], __shfl_xor_sync(0xffffffffu, hi[c], m));
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
    for (int w = 1; w < NL_THREADS / 32; w++)
      v = threadIdx.x < 3 ? fmin(v, s[w][threadIdx.x]) : fmax(v, s[w][threadIdx.x]);
    part[6 * blockIdx.x + threadIdx.x] = v;
  }
}

__global__ void __launch_bounds__(NL_THREADS)
bin_hist_kernel(int ntot, const double* __restrict__ coords, Grid3 g, int32_t* __restrict__ bin,
                int32_t* __restrict__ hist)
{
  const int a = blockIdx.x * NL_THREADS + threadIdx.x;
  if (a >= ntot) return;
  const double p[3] = {coords[3 * a], coords[3 * a + 1], coords[3 * a + 2]};
  int b[3];
  bin_of(p, g, b);
  const int q = b[0] + b[1] * g.nb[0] + b[2] * g.nb[0] * g.nb[1];
  bin[a] = q;
  atomicAdd(&hist[q], 1);
}

__global__ void __launch_bounds__(NL_THREADS)
cell_fill_kernel(int ntot, const double* __restrict__ coords, const int32_t* __restrict__ bin,
                 const int32_t* __restrict__ cellstart, int32_t* __restrict__ cursor,
                 int32_t* __restrict__ sorted_id, double* __restrict__ sorted_xyz)
{
  const int a = blockIdx.x * NL_THREADS + threadIdx.x;
  if (a >= ntot) return;
  const int q = bin[a];
  const int pos = cellstart[q] + atomicAdd(&cursor[q], 1);
  sorted_id[pos] = a;
  sorted_xyz[3 * pos] = coords[3 * a];
  sorted_xyz[3 * pos + 1] = coords[3 * a + 1];
  sorted_xyz[3 * pos + 2] = coords[3 * a + 2];
}

// Count pass: one thread per cell-ordered position (neighbouring threads share candidates).
__global__ void __launch_bounds__(NL_THREADS)
nl_count_kernel(int ntot, int n_need, const int32_t* __restrict__ need, int half, Grid3 g,
                double cutsq, const int32_t* __restrict__ cellstart,
                const int32_t* __restrict__ sorted_id, const double* __restrict__ sorted_xyz,
                int32_t* __restrict__ counts, int32_t* __restrict__ err)
{
  const int pos = blockIdx.x * NL_THREADS + threadIdx.x;
  if (pos >= ntot) return;
  const int a = sorted_id[pos];
  const bool wants = need ? (need[a] != 0) : (a < n_need);
  if (!wants) { counts[a] = 0; return; }
  const double p[3] = {sorted_xyz[3 * pos], sorted_xyz[3 * pos + 1], sorted_xyz[3 * pos + 2]};
  int b[3];
  bin_of(p, g, b);
  int cnt = 0;
  bool clash = false;
  for (int bz = max(0, b[2] - 1); bz <= min(b[2] + 1, g.nb[2] - 1); bz++)
    for (int by = max(0, b[1] - 1); by <= min(b[1] + 1, g.nb[1] - 1); by++)
    {
      // the three x-cells of a row are contiguous in the cell-major order
      const int q0 = max(0, b[0] - 1) + by * g.nb[0] + bz * g.nb[0] * g.nb[1];
      const int q1 = min(b[0] + 1, g.nb[0] - 1) + by * g.nb[0] + bz * g.nb[0] * g.nb[1];
      for (int s = cellstart[q0]; s < cellstart[q1 + 1]; s++)
      {
        if (s == pos) continue;
        const double rsq = xnorm2(xsub(sorted_xyz[3 * s], p[0]), xsub(sorted_xyz[3 * s + 1], p[1]),
                                  xsub(sorted_xyz[3 * s + 2], p[2]));
        if (rsq < KB_TOL) clash = true;  // neighbor_list.cpp:199
        if (rsq < cutsq && (!half || sorted_id[s] > a)) cnt++;
      }
    }
  counts[a] = cnt;
  if (clash) atomicExch(err, 1);
}

// Fill pass: one warp per atom.  Candidates are compacted in scan order into shared memory,
// rank-sorted (indices are distinct), and written coalesced.
constexpr int FILL_WARPS = 8;
constexpr int FILL_CAP = 512;  // neighbors staged per warp; longer lists sort in global memory

__global__ void __launch_bounds__(FILL_WARPS * 32)
nl_fill_kernel(int ntot, int n_need, const int32_t* __restrict__ need, int half, Grid3 g,
               double cutsq, const int32_t* __restrict__ cellstart,
               const int32_t* __restrict__ sorted_id, const double* __restrict__ sorted_xyz,
               const int64_t* __restrict__ offsets, int32_t* __restrict__ indices)
{
  __shared__ int32_t stage[FILL_WARPS][FILL_CAP];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int pos = blockIdx.x * FILL_WARPS + warp;
  if (pos >= ntot) return;
  const int a = sorted_id[pos];
  const bool wants = need ? (need[a] != 0) : (a < n_need);
  if (!wants) return;
  const int64_t off = offsets[a];
  const int n = (int) (offsets[a + 1] - off);
  if (n == 0) return;
  const bool in_smem = n <= FILL_CAP;
  int32_t* dst = in_smem ? stage[warp] : indices + off;
  const double p[3] = {sorted_xyz[3 * pos], sorted_xyz[3 * pos + 1], sorted_xyz[3 * pos + 2]};
  int b[3];
  bin_of(p, g, b);
  int filled = 0;
  for (int bz = max(0, b[2] - 1); bz <= min(b[2] + 1, g.nb[2] - 1); bz++)
    for (int by = max(0, b[1] - 1); by <= min(b[1] + 1, g.nb[1] - 1); by++)
    {
      const int q0 = max(0, b[0] - 1) + by * g.nb[0] + bz * g.nb[0] * g.nb[1];
      const int q1 = min(b[0] + 1, g.nb[0] - 1) + by * g.nb[0] + bz * g.nb[0] * g.nb[1];
      const int s1 = cellstart[q1 + 1];
      for (int s0 = cellstart[q0]; s0 < s1; s0 += 32)
      {
        const int s = s0 + lane;
        bool hit = false;
        int id = 0;
        if (s < s1 && s != pos)
        {
          const double rsq = xnorm2(xsub(sorted_xyz[3 * s], p[0]), xsub(sorted_xyz[3 * s + 1], p[1]),
                                    xsub(sorted_xyz[3 * s + 2], p[2]));
          id = sorted_id[s];
          hit = rsq < cutsq && (!half || id > a);
        }
        const unsigned bal = __ballot_sync(0xffffffffu, hit);
        if (hit) dst[filled + __popc(bal & ((1u << lane) - 1u))] = id;
        filled += __popc(bal);
      }
    }
  __syncwarp();
  if (in_smem)
  {
    for (int e = lane; e < n; e += 32)
    {
      const int v = dst[e];
      int rank = 0;
      for (int t = 0; t < n; t++) rank += (dst[t] < v);
      indices[off + rank] = v;
    }
  }
  else
  {
    // rare: very long list; odd-even transposition sort in place in global memory
    for (int round = 0; round < n; round++)
    {
      for (int e = (round & 1) + 2 * lane; e + 1 < n; e += 64)
      {
        const int x = dst[e], y = dst[e + 1];
        if (x > y) { dst[e] = y; dst[e + 1] = x; }
      }
      __syncwarp();
    }
  }
}

__global__ void max_count_kernel(int ntot, const int32_t* __restrict__ counts, int32_t* __restrict__ out)
{
  int m = 0;
  for (int a = blockIdx.x * blockDim.x + threadIdx.x; a < ntot; a += gridDim.x * blockDim.x)
    m = max(m, counts[a]);
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, d));
  if ((threadIdx.x & 31) == 0 && m > 0) atomicMax(out, m);
}

}  // namespace

extern "C" void kb_nl_destroy(kb_nl_plan* p, void* stream)
{
  if (!p) return;
  cudaStream_t st = kb_stream(stream);
  if (p->d_cellstart) cudaFreeAsync(p->d_cellstart, st);
  if (p->d_sorted_id) cudaFreeAsync(p->d_sorted_id, st);
  if (p->d_sorted_xyz) cudaFreeAsync(p->d_sorted_xyz, st);
  if (p->d_err) cudaFreeAsync(p->d_err, st);
  delete p;
}

extern "C" int kb_nl_begin(int32_t ntot, int32_t n_need, const double* d_coords,
                           const int32_t* d_need, double infl, double cutoff, int32_t half,
                           int32_t* d_counts, int64_t* d_offsets, int64_t* h_total,
                           int32_t* h_maxn, kb_nl_plan** out_plan, void* stream)
{
  if (!out_plan || !h_total || !h_maxn || ntot < 0 || !(infl > 0.0)) return KB_ERR_INVALID;
  *out_plan = nullptr;
  *h_total = 0;
  *h_maxn = 0;
  cudaStream_t st = kb_stream(stream);
  kb_pool_setup();
  kb_nl_plan* p = new (std::nothrow) kb_nl_plan();
  if (!p) return KB_ERR_INVALID;
  p->ntot = ntot; p->n_need = n_need; p->half = half; p->d_coords = d_coords; p->d_need = d_need;
  p->d_counts = d_counts; p->d_offsets = d_offsets;
  p->d_cellstart = nullptr; p->d_sorted_id = nullptr; p->d_sorted_xyz = nullptr; p->d_err = nullptr;
  p->total = 0; p->maxn = 0;
  p->cutsq = cutoff * cutoff;  // neighbor_list.cpp:155
  if (ntot == 0)
  {
    cudaMemsetAsync(d_offsets, 0, sizeof(int64_t), st);
    *out_plan = p;
    return KB_OK;
  }

  int rc = KB_OK;
  double* d_part = nullptr;
  int32_t *d_bin = nullptr, *d_hist = nullptr;
  int32_t h_tail[2] = {0, 0};
#define NL_TRY(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { rc = -(int) e_; goto fail; } } while (0)
  {
    // ---- bounding box; the reference starts max at (first atom + 1), neighbor_list.cpp:90-96
    const int nblk_bbox = 296;
    double h_part[6 * 296 + 3];
    NL_TRY(cudaMallocAsync((void**) &d_part, sizeof(double) * (6 * nblk_bbox + 3), st));
    bbox_kernel<<<nblk_bbox, NL_THREADS, 0, st>>>(ntot, d_coords, d_part);
    NL_TRY(cudaGetLastError()); kb_note_launch();
    NL_TRY(cudaMemcpyAsync(d_part + 6 * nblk_bbox, d_coords, sizeof(double) * 3, cudaMemcpyDeviceToDevice, st));
    NL_TRY(cudaMemcpyAsync(h_part, d_part, sizeof(double) * (6 * nblk_bbox + 3), cudaMemcpyDeviceToHost, st));
    NL_TRY(cudaStreamSynchronize(st));
    for (int c = 0; c < 3; c++)
    {
      const double first = h_part[6 * nblk_bbox + c];
      double lo = first, hi = first + 1.0;
      for (int b = 0; b < nblk_bbox; b++)
      {
        if (h_part[6 * b + c] < lo) lo = h_part[6 * b + c];
        if (h_part[6 * b + 3 + c] > hi) hi = h_part[6 * b + 3 + c];
      }
      p->lo[c] = lo;
      p->hi[c] = hi;
      const double cells = (hi - lo) / infl;  // :115-117
      if (!(cells < 2.0e9)) { rc = KB_ERR_REFERENCE; goto fail; }
      p->nb[c] = (int) cells;
      if (p->nb[c] <= 0) p->nb[c] = 1;
    }
    const double ncell_d = (double) p->nb[0] * (double) p->nb[1] * (double) p->nb[2];
    if (ncell_d > 1.0e9) { rc = KB_ERR_REFERENCE; goto fail; }  // :123-128
    // Cells wider than the influence distance are equally valid (27-cell scan still covers the
    // cutoff sphere); coarsen when the reference's grid would be mostly empty bookkeeping.
    while ((double) p->nb[0] * p->nb[1] * p->nb[2] > 8.0 * ntot + 4096.0)
    {
      int c = 0;
      if (p->nb[1] > p->nb[c]) c = 1;
      if (p->nb[2] > p->nb[c]) c = 2;
      p->nb[c] = (p->nb[c] + 1) / 2;
    }
  }
  {
    const int64_t ncell = (int64_t) p->nb[0] * p->nb[1] * p->nb[2];
    Grid3 g;
    for (int c = 0; c < 3; c++) { g.lo[c] = p->lo[c]; g.hi[c] = p->hi[c]; g.nb[c] = p->nb[c]; }
    const int nblk = (ntot + NL_THREADS - 1) / NL_THREADS;
    NL_TRY(cudaMallocAsync((void**) &d_bin, sizeof(int32_t) * (size_t) ntot, st));
    NL_TRY(cudaMallocAsync((void**) &d_hist, sizeof(int32_t) * (size_t) ncell, st));
    NL_TRY(cudaMallocAsync((void**) &p->d_cellstart, sizeof(int32_t) * ((size_t) ncell + 1), st));
    NL_TRY(cudaMallocAsync((void**) &p->d_sorted_id, sizeof(int32_t) * (size_t) ntot, st));
    NL_TRY(cudaMallocAsync((void**) &p->d_sorted_xyz, sizeof(double) * 3 * (size_t) ntot, st));
    NL_TRY(cudaMallocAsync((void**) &p->d_err, sizeof(int32_t) * 2, st));
    NL_TRY(cudaMemsetAsync(d_hist, 0, sizeof(int32_t) * (size_t) ncell, st));
    NL_TRY(cudaMemsetAsync(p->d_err, 0, sizeof(int32_t) * 2, st));
    bin_hist_kernel<<<nblk, NL_THREADS, 0, st>>>(ntot, d_coords, g, d_bin, d_hist);
    NL_TRY(cudaGetLastError()); kb_note_launch();
    rc = kb_exclusive_scan<int32_t, int32_t>(d_hist, p->d_cellstart, ncell, p->d_cellstart + ncell, st);
    if (rc) goto fail;
    NL_TRY(cudaMemsetAsync(d_hist, 0, sizeof(int32_t) * (size_t) ncell, st));
    cell_fill_kernel<<<nblk, NL_THREADS, 0, st>>>(ntot, d_coords, d_bin, p->d_cellstart, d_hist,
                                                  p->d_sorted_id, p->d_sorted_xyz);
    NL_TRY(cudaGetLastError()); kb_note_launch();
    nl_count_kernel<<<nblk, NL_THREADS, 0, st>>>(ntot, n_need, d_need, half, g, p->cutsq,
                                                 p->d_cellstart, p->d_sorted_id, p->d_sorted_xyz,
                                                 d_counts, p->d_err);
    NL_TRY(cudaGetLastError()); kb_note_launch();
    rc = kb_exclusive_scan<int32_t, int64_t>(d_counts, d_offsets, ntot, d_offsets + ntot, st);
    if (rc) goto fail;
    max_count_kernel<<<296, 256, 0, st>>>(ntot, d_counts, p->d_err + 1);
    NL_TRY(cudaGetLastError()); kb_note_launch();
    NL_TRY(cudaMemcpyAsync(&p->total, d_offsets + ntot, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
    NL_TRY(cudaMemcpyAsync(h_tail, p->d_err, sizeof(int32_t) * 2, cudaMemcpyDeviceToHost, st));
    NL_TRY(cudaFreeAsync(d_part, st)); d_part = nullptr;
    NL_TRY(cudaFreeAsync(d_bin, st)); d_bin = nullptr;
    NL_TRY(cudaFreeAsync(d_hist, st)); d_hist = nullptr;
    NL_TRY(cudaStreamSynchronize(st));
    if (h_tail[0]) { rc = KB_ERR_REFERENCE; goto fail; }  // collision, :199-209
    p->maxn = h_tail[1];
  }
#undef NL_TRY
  *h_total = p->total;
  *h_maxn = p->maxn;
  *out_plan = p;
  return KB_OK;
fail:
  if (d_part) cudaFreeAsync(d_part, st);
  if (d_bin) cudaFreeAsync(d_bin, st);
  if (d_hist) cudaFreeAsync(d_hist, st);
  kb_nl_destroy(p, stream);
  return rc;
}

extern "C" int kb_nl_finish(kb_nl_plan* p, int32_t* d_indices, void* stream)
{
  if (!p) return KB_ERR_INVALID;
  cudaStream_t st = kb_stream(stream);
  int rc = KB_OK;
  if (p->total > 0)
  {
    if (!d_indices) rc = KB_ERR_INVALID;
    else
    {
      Grid3 g;
      for (int c = 0; c < 3; c++) { g.lo[c] = p->lo[c]; g.hi[c] = p->hi[c]; g.nb[c] = p->nb[c]; }
      const int nblk = (p->ntot + FILL_WARPS - 1) / FILL_WARPS;
      nl_fill_kernel<<<nblk, FILL_WARPS * 32, 0, st>>>(p->ntot, p->n_need, p->d_need, p->half, g,
                                                       p->cutsq, p->d_cellstart, p->d_sorted_id,
                                                       p->d_sorted_xyz, p->d_offsets, d_indices);
      cudaError_t e = cudaGetLastError();
      if (e != cudaSuccess) rc = -(int) e;
      else kb_note_launch();
    }
  }
  kb_nl_destroy(p, stream);
  return rc;
}
