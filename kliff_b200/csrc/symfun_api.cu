// C-ABI entry points of the descriptor stage (see include/kliff_b200.h).
#include "common.cuh"

int kb_symfun_general_launch(const kb_symfun_params& P, int n_centers, const int32_t* d_centers,
                             const int32_t* d_subset, const double* d_coords,
                             const int32_t* d_species, const int64_t* d_nl_off,
                             const int32_t* d_nl_idx, double* d_zeta, void* d_dz, int dz_dtype,
                             const int64_t* d_dz_ptr, cudaStream_t st);
int kb_symfun_tiled_launch(const kb_symfun_params& P, int n_centers, const int32_t* d_centers,
                           const double* d_coords, const int32_t* d_species,
                           const int64_t* d_nl_off, const int32_t* d_nl_idx, int maxn,
                           double* d_zeta, void* d_dz, int dz_dtype, const int64_t* d_dz_ptr,
                           int32_t* d_overflow, bool* may_overflow, cudaStream_t st);

int kb_symfun_warp_launch(const kb_symfun_params& P, int n_centers, const int32_t* d_centers,
                          const int32_t* d_subset, const double* d_coords, const int32_t* d_species,
                          const int64_t* d_nl_off, const int32_t* d_nl_idx, int maxn, double* d_zeta,
                          void* d_dz, int dz_dtype, const int64_t* d_dz_ptr, int32_t* d_overflow,
                          int full_capacity, bool* launched, bool* may_overflow, cudaStream_t st);

int kb_symfun_split_launch(const kb_symfun_params& P, int n_centers, const int32_t* d_centers,
                           const int32_t* d_subset, const double* d_coords, const int32_t* d_species,
                           const int64_t* d_nl_off, const int32_t* d_nl_idx, int maxn, double* d_zeta,
                           void* d_dz, int dz_dtype, const int64_t* d_dz_ptr, int32_t* d_overflow,
                           int variant, bool* launched, bool* may_overflow, cudaStream_t st);

namespace {

__global__ void block_size_kernel(int n_centers, const int32_t* __restrict__ centers,
                                  const int64_t* __restrict__ nl_off, int n_desc,
                                  int64_t* __restrict__ sizes)
{
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_centers) return;
  const int a = centers ? centers[t] : t;
  const int64_t n = nl_off[a + 1] - nl_off[a];
  sizes[t] = (3LL * n_desc * (n + 1) + 15) & ~15LL;  // 128-byte aligned blocks
}

int validate(const kb_symfun_params* P)
{
  if (!P) return KB_ERR_INVALID;
  if (P->n_species < 1 || P->n_species > KB_MAX_SPECIES) return KB_ERR_LIMIT;
  if (P->n_two < 0 || P->n_two > KB_MAX_TWO) return KB_ERR_LIMIT;
  if (P->n_groups < 0 || P->n_groups > KB_MAX_GROUPS) return KB_ERR_LIMIT;
  if (P->n_desc < 1) return KB_ERR_INVALID;
  for (int q = 0; q < P->n_two; q++)
    if (P->two_kind[q] < KB_G1 || P->two_kind[q] > KB_G3 || P->two_out[q] < 0
        || P->two_out[q] >= P->n_desc)
      return KB_ERR_INVALID;
  for (int g = 0; g < P->n_groups; g++)
  {
    if (P->grp_kind[g] != KB_G4 && P->grp_kind[g] != KB_G5) return KB_ERR_INVALID;
    if (P->grp_n[g] < 1 || P->grp_n[g] > KB_GROUP_WIDTH) return KB_ERR_INVALID;
    for (int e = 0; e < P->grp_n[g]; e++)
      if (P->grp_out[g][e] >= P->n_desc) return KB_ERR_INVALID;
  }
  return KB_OK;
}

}  // namespace

extern "C" int kb_symfun_block_offsets(int32_t n_centers, const int32_t* d_centers,
                                       const int64_t* d_nl_offsets, int32_t n_desc,
                                       int64_t* d_dz_ptr, int64_t* h_total, void* stream)
{
  if (n_centers < 0 || !d_dz_ptr || !h_total || n_desc < 1) return KB_ERR_INVALID;
  cudaStream_t st = kb_stream(stream);
  kb_pool_setup();
  *h_total = 0;
  if (n_centers == 0)
  {
    KB_CUDA(cudaMemsetAsync(d_dz_ptr, 0, sizeof(int64_t), st));
    return KB_OK;
  }
  int64_t* sizes = nullptr;
  KB_CUDA(cudaMallocAsync((void**) &sizes, sizeof(int64_t) * (size_t) n_centers, st));
  int rc = KB_OK;
  block_size_kernel<<<(n_centers + 255) / 256, 256, 0, st>>>(n_centers, d_centers, d_nl_offsets,
                                                             n_desc, sizes);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) rc = -(int) e;
  else kb_note_launch();
  if (rc == KB_OK) rc = kb_exclusive_scan<int64_t, int64_t>(sizes, d_dz_ptr, n_centers, d_dz_ptr + n_centers, st);
  if (rc == KB_OK)
  {
    e = cudaMemcpyAsync(h_total, d_dz_ptr + n_centers, sizeof(int64_t), cudaMemcpyDeviceToHost, st);
    if (e != cudaSuccess) rc = -(int) e;
  }
  cudaFreeAsync(sizes, st);  // on every path
  e = cudaStreamSynchronize(st);
  if (rc == KB_OK && e != cudaSuccess) rc = -(int) e;
  return rc;
}

extern "C" int kb_symfun_forward(const kb_symfun_params* h_params, int32_t n_centers,
                                 const int32_t* d_centers, const double* d_coords,
                                 const int32_t* d_species, const int64_t* d_nl_offsets,
                                 const int32_t* d_nl_indices, int32_t maxn, double* d_zeta,
                                 void* d_dz, int32_t dz_dtype, const int64_t* d_dz_ptr,
                                 int32_t mode, void* stream)
{
  int rc = validate(h_params);
  if (rc) return rc;
  if (n_centers == 0) return KB_OK;
  if (n_centers < 0 || !d_zeta || (d_dz && !d_dz_ptr) || maxn < 0) return KB_ERR_INVALID;
  if (dz_dtype != KB_F64 && dz_dtype != KB_F32) return KB_ERR_INVALID;
  cudaStream_t st = kb_stream(stream);
  kb_pool_setup();
  const kb_symfun_params& P = *h_params;

  if (mode == KB_SYMFUN_GENERAL)
    return kb_symfun_general_launch(P, n_centers, d_centers, nullptr, d_coords, d_species,
                                    d_nl_offsets, d_nl_indices, d_zeta, d_dz, dz_dtype, d_dz_ptr, st);

  // ---- preferred: warp-per-atom kernels (g4 sets with canonical exponents, n <= 64): the two-stage
  //      form first (table kernel + triple-sweep kernel), the fused kernel on request and for the
  //      atoms the first pass declines
  if (mode == KB_SYMFUN_AUTO || mode == KB_SYMFUN_WARP || mode == KB_SYMFUN_SPLIT || mode == KB_SYMFUN_QUEUE)
  {
    int32_t* ov = nullptr;
    KB_CUDA(cudaMallocAsync((void**) &ov, sizeof(int32_t) * ((size_t) n_centers + 2), st));
    bool launched = false, may_overflow = false;
    {
      cudaError_t e0 = cudaMemsetAsync(ov, 0, 2 * sizeof(int32_t), st);
      if (e0 != cudaSuccess) { cudaFreeAsync(ov, st); return -(int) e0; }
    }
    if (mode == KB_SYMFUN_WARP)
      rc = kb_symfun_warp_launch(P, n_centers, d_centers, nullptr, d_coords, d_species, d_nl_offsets,
                                 d_nl_indices, maxn, d_zeta, d_dz, dz_dtype, d_dz_ptr, ov, 0, &launched,
                                 &may_overflow, st);
    else  // QUEUE: queue sweep where it applies, else the v3 sweep; AUTO / SPLIT: v3 (until the queue form wins)
      rc = kb_symfun_split_launch(P, n_centers, d_centers, nullptr, d_coords, d_species, d_nl_offsets,
                                  d_nl_indices, maxn, d_zeta, d_dz, dz_dtype, d_dz_ptr, ov,
                                  mode == KB_SYMFUN_QUEUE ? 0 : 2, &launched, &may_overflow, st);
    if (rc == KB_OK && launched && may_overflow)
    {
      int32_t n_over = 0;
      cudaError_t e = cudaMemcpyAsync(&n_over, ov, sizeof(int32_t), cudaMemcpyDeviceToHost, st);
      if (e == cudaSuccess) e = cudaStreamSynchronize(st);
      if (e != cudaSuccess) rc = -(int) e;
      if (rc == KB_OK && n_over > 0)
      {
        // second pass with worst-case record capacity over the declined atoms only
        int32_t* ov2 = nullptr;
        e = cudaMallocAsync((void**) &ov2, sizeof(int32_t) * ((size_t) n_over + 2), st);
        if (e == cudaSuccess) e = cudaMemsetAsync(ov2, 0, 2 * sizeof(int32_t), st);
        if (e != cudaSuccess)
        {
          if (ov2) cudaFreeAsync(ov2, st);
          cudaFreeAsync(ov, st);
          return -(int) e;
        }
        bool l2 = false, m2 = false;
        rc = kb_symfun_warp_launch(P, n_over, d_centers, ov + 2, d_coords, d_species, d_nl_offsets,
                                   d_nl_indices, maxn, d_zeta, d_dz, dz_dtype, d_dz_ptr, ov2, 1, &l2, &m2, st);
        int32_t n_over2 = l2 ? 0 : n_over;
        if (rc == KB_OK && l2 && m2)
        {
          e = cudaMemcpyAsync(&n_over2, ov2, sizeof(int32_t), cudaMemcpyDeviceToHost, st);
          if (e == cudaSuccess) e = cudaStreamSynchronize(st);
          if (e != cudaSuccess) rc = -(int) e;
        }
        if (rc == KB_OK && n_over2 > 0)  // n > 64 (or nothing fits): global-atomics kernel
          rc = kb_symfun_general_launch(P, n_over2, d_centers, l2 ? ov2 + 2 : ov + 2, d_coords,
                                        d_species, d_nl_offsets, d_nl_indices, d_zeta, d_dz, dz_dtype, d_dz_ptr, st);
        cudaFreeAsync(ov2, st);
      }
    }
    cudaFreeAsync(ov, st);
    if (rc != KB_OK || launched) return rc;
    if (mode != KB_SYMFUN_AUTO) return KB_ERR_LIMIT;  // parameter set outside the warp kernels' domain
  }

  // ---- CTA-per-atom kernel: g5 families and generic exponents
  int32_t* d_overflow = nullptr;
  KB_CUDA(cudaMallocAsync((void**) &d_overflow, sizeof(int32_t) * ((size_t) n_centers + 1), st));
  {
    cudaError_t e0 = cudaMemsetAsync(d_overflow, 0, sizeof(int32_t), st);
    if (e0 != cudaSuccess) { cudaFreeAsync(d_overflow, st); return -(int) e0; }
  }
  bool may_overflow = false;
  rc = kb_symfun_tiled_launch(P, n_centers, d_centers, d_coords, d_species, d_nl_offsets,
                              d_nl_indices, maxn, d_zeta, d_dz, dz_dtype, d_dz_ptr, d_overflow,
                              &may_overflow, st);
  if (rc == KB_OK && may_overflow)
  {
    // atoms the tiled kernel declined (n > 64 or pair records beyond shared memory)
    int32_t n_over = 0;
    cudaError_t e = cudaMemcpyAsync(&n_over, d_overflow, sizeof(int32_t), cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) rc = -(int) e;
    else if (n_over > 0)
    {
      if (mode == KB_SYMFUN_TILED) rc = KB_ERR_LIMIT;
      else
        rc = kb_symfun_general_launch(P, n_over, d_centers, d_overflow + 1, d_coords, d_species,
                                      d_nl_offsets, d_nl_indices, d_zeta, d_dz, dz_dtype, d_dz_ptr, st);
    }
  }
  cudaFreeAsync(d_overflow, st);
  return rc;
}
