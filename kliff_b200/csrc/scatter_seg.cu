// Owner-sorted (segmented) form of the chain-rule force assembly: bit-reproducible forces.
//
// kb_force_scatter_fwd (scatter.cu) adds every centre's slot contributions to their owner atoms with
// fp64 atomics: correct, but the order of the additions — and with it the last bits of the forces —
// changes from run to run.  Here the assembly F[image[a], c] -= sum_d dEdG[t, d] dz_t[d, slot(a), c]
// (CalculatorTorch._compute_forces, calculator_torch.py:207-213, 266-269; NeighborList image folding,
// neighbor.py:260-296) is split in two launches:
//   1. one warp per centre atom contracts its [D][3(n+1)] block with dEdG[t, :] exactly as before and
//      writes the 3(n+1) partial forces of its slots, coalesced, to a scratch array ("slot contributions");
//   2. one warp per owner atom sums the contributions of the slots it owns in a FIXED order: the plan
//      lists, per owner, the slot ids in ascending order (kb_scatter_plan_*: count, scan, fill, per-
//      segment rank sort — integer work only, built once per neighbor list and reused every step).
// HBM traffic: the blocks once (36 KB per atom) + 2 x 24 B per slot (1.4 KB per atom) + 4 B per slot of plan.
#include "common.cuh"

namespace {

constexpr int SG_WARPS = 8;
constexpr int SG_COLS = 4;

__global__ void seg_slot_count_kernel(int n_centers, const int32_t* __restrict__ centers,
                                      const int64_t* __restrict__ nl_off, int64_t* __restrict__ sizes)
{
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_centers) return;
  const int a = centers ? centers[t] : t;
  sizes[t] = nl_off[a + 1] - nl_off[a] + 1;
}

// pass 0: owners' slot counts (integer atomics: the counts do not depend on the order);
// pass 1: slot ids appended to their owner's segment in arrival order (sorted afterwards)
template <int PASS>
__global__ void __launch_bounds__(SG_WARPS * 32)
seg_fill_kernel(int n_centers, const int32_t* __restrict__ centers, const int64_t* __restrict__ nl_off,
                const int32_t* __restrict__ nl_idx, const int32_t* __restrict__ image, int n_out,
                const int64_t* __restrict__ slot_ptr, const int64_t* __restrict__ seg_ptr,
                int32_t* __restrict__ cursor, int32_t* __restrict__ seg_tmp, int32_t* __restrict__ bad)
{
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int t = blockIdx.x * SG_WARPS + warp;
  if (t >= n_centers) return;
  const int a = centers ? centers[t] : t;
  const int64_t off = nl_off[a];
  const int n = (int) (nl_off[a + 1] - off);
  for (int slot = lane; slot <= n; slot += 32)
  {
    const int atom = slot < n ? nl_idx[off + slot] : a;
    const int owner = image[atom];
    if (owner < 0 || owner >= n_out) { *bad = 1; continue; }
    if (PASS == 0) atomicAdd(cursor + owner, 1);
    else
    {
      const int pos = atomicAdd(cursor + owner, 1);
      seg_tmp[seg_ptr[owner] + pos] = (int32_t) (slot_ptr[t] - slot_ptr[0] + slot);
    }
  }
}

// ascending slot ids inside every owner's segment: rank sort (ids are distinct)
__global__ void __launch_bounds__(SG_WARPS * 32)
seg_sort_kernel(int n_out, const int64_t* __restrict__ seg_ptr, const int32_t* __restrict__ seg_tmp,
                int32_t* __restrict__ seg_src)
{
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int o = blockIdx.x * SG_WARPS + warp;
  if (o >= n_out) return;
  const int64_t s0 = seg_ptr[o];
  const int len = (int) (seg_ptr[o + 1] - s0);
  if (len <= 32)
  {
    const int key = lane < len ? seg_tmp[s0 + lane] : 0x7fffffff;
    int rank = 0;
#pragma unroll
    for (int m = 0; m < 32; m++) rank += __shfl_sync(0xffffffffu, key, m) < key;
    if (lane < len) seg_src[s0 + rank] = key;
  }
  else
    for (int i = lane; i < len; i += 32)
    {
      const int key = seg_tmp[s0 + i];
      int rank = 0;
      for (int m = 0; m < len; m++) rank += seg_tmp[s0 + m] < key;
      seg_src[s0 + rank] = key;
    }
}

// launch 1: slot contributions contrib[slot id][c] = sum_d dEdG[t, d] scale[d] dz_t[d, slot, c]
template <typename TDz>
__global__ void __launch_bounds__(SG_WARPS * 32)
seg_contract_kernel(int n_centers, const int32_t* __restrict__ centers, const int64_t* __restrict__ nl_off,
                    int D, const double* __restrict__ dEdG, const double* __restrict__ scale,
                    const TDz* __restrict__ dz, const int64_t* __restrict__ dz_ptr,
                    const int64_t* __restrict__ slot_ptr, double* __restrict__ contrib)
{
  extern __shared__ double wsm[];  // [SG_WARPS][D] weights
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int t = blockIdx.x * SG_WARPS + warp;
  if (t >= n_centers) return;
  const int a = centers ? centers[t] : t;
  const int n = (int) (nl_off[a + 1] - nl_off[a]);
  const int row = 3 * (n + 1);
  const TDz* blk = dz + dz_ptr[t];
  double* w = wsm + warp * D;
  for (int d = lane; d < D; d += 32) w[d] = dEdG[(int64_t) t * D + d] * (scale ? scale[d] : 1.0);
  __syncwarp();
  double* out = contrib + 3 * (slot_ptr[t] - slot_ptr[0]);
  for (int c0 = 0; c0 < row; c0 += 32 * SG_COLS)
  {
    double acc[SG_COLS];
    bool on[SG_COLS];
#pragma unroll
    for (int u = 0; u < SG_COLS; u++) { acc[u] = 0.0; on[u] = c0 + u * 32 + lane < row; }
    for (int d = 0; d < D; d++)
    {
      const double wd = w[d];
      const TDz* rp = blk + (int64_t) d * row + c0 + lane;
#pragma unroll
      for (int u = 0; u < SG_COLS; u++)
        if (on[u]) acc[u] += wd * (double) rp[u * 32];
    }
#pragma unroll
    for (int u = 0; u < SG_COLS; u++)
      if (on[u]) out[c0 + u * 32 + lane] = acc[u];
  }
}

// launch 2: F[owner, c] -= sum over the owner's slots, ascending slot id per lane, fixed shuffle tree
__global__ void __launch_bounds__(SG_WARPS * 32)
seg_reduce_kernel(int n_out, const int64_t* __restrict__ seg_ptr, const int32_t* __restrict__ seg_src,
                  int64_t slot_base, const double* __restrict__ contrib, double* __restrict__ forces)
{
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int o = blockIdx.x * SG_WARPS + warp;
  if (o >= n_out) return;
  const int64_t s0 = seg_ptr[o], s1 = seg_ptr[o + 1];
  double fx = 0.0, fy = 0.0, fz = 0.0;
  for (int64_t i = s0 + lane; i < s1; i += 32)
  {
    const double* p = contrib + 3 * ((int64_t) seg_src[i] - slot_base);
    fx += p[0]; fy += p[1]; fz += p[2];
  }
  fx = warp_sum(fx); fy = warp_sum(fy); fz = warp_sum(fz);
  if (lane == 0)
  {
    double* f = forces + 3 * (int64_t) o;
    f[0] -= fx; f[1] -= fy; f[2] -= fz;
  }
}

}  // namespace

extern "C" int kb_scatter_plan_begin(int32_t n_centers, const int32_t* d_centers,
                                     const int64_t* d_nl_offsets, int64_t* d_slot_ptr,
                                     int64_t* h_total_slots, void* stream)
{
  if (n_centers < 0 || !d_nl_offsets || !d_slot_ptr || !h_total_slots) return KB_ERR_INVALID;
  cudaStream_t st = kb_stream(stream);
  kb_pool_setup();
  if (n_centers == 0)
  {
    *h_total_slots = 0;
    KB_CUDA(cudaMemsetAsync(d_slot_ptr, 0, sizeof(int64_t), st));
    KB_CUDA(cudaStreamSynchronize(st));
    return KB_OK;
  }
  int64_t* sizes = nullptr;
  KB_CUDA(cudaMallocAsync((void**) &sizes, sizeof(int64_t) * (size_t) n_centers, st));
  seg_slot_count_kernel<<<(n_centers + 255) / 256, 256, 0, st>>>(n_centers, d_centers, d_nl_offsets, sizes);
  int rc = KB_OK;
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) rc = -(int) e;
  else kb_note_launch();
  if (rc == KB_OK)
    rc = kb_exclusive_scan<int64_t, int64_t>(sizes, d_slot_ptr, n_centers, d_slot_ptr + n_centers, st);
  if (rc == KB_OK)
  {
    e = cudaMemcpyAsync(h_total_slots, d_slot_ptr + n_centers, sizeof(int64_t), cudaMemcpyDeviceToHost, st);
    if (e != cudaSuccess) rc = -(int) e;
  }
  cudaFreeAsync(sizes, st);
  e = cudaStreamSynchronize(st);
  if (rc == KB_OK && e != cudaSuccess) rc = -(int) e;
  if (rc == KB_OK && *h_total_slots >= 0x7fffffffLL) rc = KB_ERR_INVALID;  // slot ids are 32-bit
  return rc;
}

extern "C" int kb_scatter_plan_finish(int32_t n_centers, const int32_t* d_centers,
                                      const int64_t* d_nl_offsets, const int32_t* d_nl_indices,
                                      const int32_t* d_image, int32_t n_out, const int64_t* d_slot_ptr,
                                      int64_t total_slots, int64_t* d_seg_ptr, int32_t* d_seg_src,
                                      void* stream)
{
  if (n_centers < 0 || n_out < 0 || total_slots < 0 || !d_seg_ptr) return KB_ERR_INVALID;
  cudaStream_t st = kb_stream(stream);
  kb_pool_setup();
  if (n_out == 0 || n_centers == 0)
  {
    KB_CUDA(cudaMemsetAsync(d_seg_ptr, 0, sizeof(int64_t) * ((size_t) n_out + 1), st));
    return KB_OK;
  }
  int32_t *cursor = nullptr, *tmp = nullptr, *bad = nullptr;
  int rc = KB_OK;
  cudaError_t e = cudaMallocAsync((void**) &cursor, sizeof(int32_t) * ((size_t) n_out + 1), st);
  if (e == cudaSuccess) e = cudaMallocAsync((void**) &tmp, sizeof(int32_t) * (size_t) (total_slots ? total_slots : 1), st);
  if (e == cudaSuccess) e = cudaMemsetAsync(cursor, 0, sizeof(int32_t) * ((size_t) n_out + 1), st);
  bad = cursor ? cursor + n_out : nullptr;
  const int nblk = (n_centers + SG_WARPS - 1) / SG_WARPS, oblk = (n_out + SG_WARPS - 1) / SG_WARPS;
  int32_t h_bad = 0;
  if (e == cudaSuccess)
  {
    seg_fill_kernel<0><<<nblk, SG_WARPS * 32, 0, st>>>(n_centers, d_centers, d_nl_offsets, d_nl_indices, d_image,
                                                       n_out, d_slot_ptr, nullptr, cursor, nullptr, bad);
    rc = kb_exclusive_scan<int32_t, int64_t>(cursor, d_seg_ptr, n_out, d_seg_ptr + n_out, st);
    if (rc == KB_OK) e = cudaMemsetAsync(cursor, 0, sizeof(int32_t) * (size_t) n_out, st);
    if (rc == KB_OK && e == cudaSuccess)
    {
      seg_fill_kernel<1><<<nblk, SG_WARPS * 32, 0, st>>>(n_centers, d_centers, d_nl_offsets, d_nl_indices,
                                                         d_image, n_out, d_slot_ptr, d_seg_ptr, cursor, tmp, bad);
      seg_sort_kernel<<<oblk, SG_WARPS * 32, 0, st>>>(n_out, d_seg_ptr, tmp, d_seg_src);
      e = cudaGetLastError();
      if (e == cudaSuccess) kb_note_launch(3);
      if (e == cudaSuccess) e = cudaMemcpyAsync(&h_bad, bad, sizeof(int32_t), cudaMemcpyDeviceToHost, st);
    }
  }
  if (cursor) cudaFreeAsync(cursor, st);
  if (tmp) cudaFreeAsync(tmp, st);
  if (rc == KB_OK && e != cudaSuccess) rc = -(int) e;
  if (rc == KB_OK)
  {
    e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) rc = -(int) e;
    else if (h_bad) rc = KB_ERR_INVALID;  // an image index outside [0, n_out)
  }
  return rc;
}

extern "C" int kb_force_scatter_fwd_seg(int32_t n_centers, const int32_t* d_centers,
                                        const int64_t* d_nl_offsets, int32_t n_desc, const double* d_dEdG,
                                        const double* d_scale, const void* d_dz, int32_t dz_dtype,
                                        const int64_t* d_dz_ptr, const int64_t* d_slot_ptr,
                                        int64_t n_slots, int64_t slot_base, int32_t n_out,
                                        const int64_t* d_seg_ptr, const int32_t* d_seg_src,
                                        double* d_contrib, double* d_forces, void* stream)
{
  if (n_centers < 0 || n_out < 0 || n_desc < 1 || n_slots < 0 || (dz_dtype != KB_F64 && dz_dtype != KB_F32))
    return KB_ERR_INVALID;
  if (n_centers == 0 || n_out == 0) return KB_OK;
  cudaStream_t st = kb_stream(stream);
  kb_pool_setup();
  double* contrib = d_contrib;  // caller's scratch [3 n_slots] (e.g. inside a captured graph), else from the pool
  if (!contrib) KB_CUDA(cudaMallocAsync((void**) &contrib, sizeof(double) * 3 * (size_t) (n_slots ? n_slots : 1), st));
  const int nblk = (n_centers + SG_WARPS - 1) / SG_WARPS, oblk = (n_out + SG_WARPS - 1) / SG_WARPS;
  const size_t smem = sizeof(double) * SG_WARPS * n_desc;
  kb_timing_begin(KB_T_SCATTER, st);
  if (dz_dtype == KB_F64)
    seg_contract_kernel<double><<<nblk, SG_WARPS * 32, smem, st>>>(
        n_centers, d_centers, d_nl_offsets, n_desc, d_dEdG, d_scale, static_cast<const double*>(d_dz),
        d_dz_ptr, d_slot_ptr, contrib);
  else
    seg_contract_kernel<float><<<nblk, SG_WARPS * 32, smem, st>>>(
        n_centers, d_centers, d_nl_offsets, n_desc, d_dEdG, d_scale, static_cast<const float*>(d_dz),
        d_dz_ptr, d_slot_ptr, contrib);
  seg_reduce_kernel<<<oblk, SG_WARPS * 32, 0, st>>>(n_out, d_seg_ptr, d_seg_src, slot_base, contrib, d_forces);
  kb_timing_end(KB_T_SCATTER, st);
  cudaError_t e = cudaGetLastError();
  if (!d_contrib) cudaFreeAsync(contrib, st);
  if (e != cudaSuccess) return -(int) e;
  kb_note_launch(2);
  return KB_OK;
}
