// Column statistics and normalisation of the descriptor matrix — what Descriptor.generate_fingerprints
// does with numpy after the per-configuration loop (kliff/legacy/descriptors/descriptor.py:116-118:
// mean = np.mean(zeta, axis=0), stdev = np.std(zeta, axis=0); :187-194: zeta = (zeta - mean) / stdev).
// Two-pass definition as numpy's: sum -> mean, then sum of squared deviations; both sums are
// deterministic (per-block partials in a fixed order, no atomics).  HBM-bound: 8 D bytes per atom and pass.
#include "common.cuh"

namespace {

constexpr int MO_ROWS = 512;  // rows per block
constexpr int MO_TX = 64;     // columns per tile (threads along a row: coalesced 512-byte segments)
constexpr int MO_TY = 4;      // row groups per block

// partial[b][d] = sum over the block's rows of z[r][d]            (mean == nullptr)
//               = sum of (z[r][d] - mean[d])^2                    (mean != nullptr)
__global__ void __launch_bounds__(MO_TX * MO_TY)
moments_partial_kernel(int64_t n, int D, const double* __restrict__ z, const double* __restrict__ mean,
                       double* __restrict__ partial)
{
  __shared__ double red[MO_TY][MO_TX];
  const int tx = threadIdx.x % MO_TX, ty = threadIdx.x / MO_TX;
  const int64_t r0 = (int64_t) blockIdx.x * MO_ROWS, r1 = min(n, r0 + MO_ROWS);
  for (int d0 = 0; d0 < D; d0 += MO_TX)
  {
    const int d = d0 + tx;
    double s = 0.0;
    if (d < D)
    {
      const double m = mean ? mean[d] : 0.0;
      for (int64_t r = r0 + ty; r < r1; r += MO_TY)
      {
        const double v = z[r * D + d] - m;
        s += mean ? v * v : v;
      }
    }
    red[ty][tx] = s;
    __syncthreads();
    if (ty == 0 && d < D) partial[(int64_t) blockIdx.x * D + d] = (red[0][tx] + red[1][tx]) + (red[2][tx] + red[3][tx]);
    __syncthreads();
  }
}

__global__ void moments_final_kernel(int nblk, int D, const double* __restrict__ partial, double* __restrict__ out)
{
  const int d = blockIdx.x * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (d >= D) return;
  double s = 0.0;
  for (int b = lane; b < nblk; b += 32) s += partial[(int64_t) b * D + d];
  s = warp_sum(s);
  if (lane == 0) out[d] = s;
}

__global__ void normalize_kernel(int64_t total, int D, const double* __restrict__ z, const double* __restrict__ mean,
                                 const double* __restrict__ inv_std, double* __restrict__ out)
{
  const int64_t q = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= total) return;
  const int d = (int) (q % D);
  out[q] = (z[q] - mean[d]) * inv_std[d];
}

}  // namespace

extern "C" int kb_moments(int64_t n_rows, int32_t n_desc, const double* d_zeta, const double* d_mean,
                          double* d_out, void* stream)
{
  if (n_rows < 0 || n_desc < 1 || !d_out || (n_rows > 0 && !d_zeta)) return KB_ERR_INVALID;
  cudaStream_t st = kb_stream(stream);
  kb_pool_setup();
  if (n_rows == 0)
  {
    KB_CUDA(cudaMemsetAsync(d_out, 0, sizeof(double) * n_desc, st));
    return KB_OK;
  }
  const int nblk = (int) ((n_rows + MO_ROWS - 1) / MO_ROWS);
  double* partial = nullptr;
  KB_CUDA(cudaMallocAsync((void**) &partial, sizeof(double) * (size_t) nblk * n_desc, st));
  moments_partial_kernel<<<nblk, MO_TX * MO_TY, 0, st>>>(n_rows, n_desc, d_zeta, d_mean, partial);
  moments_final_kernel<<<(n_desc + 3) / 4, 128, 0, st>>>(nblk, n_desc, partial, d_out);
  cudaError_t e = cudaGetLastError();
  cudaFreeAsync(partial, st);
  if (e != cudaSuccess) return -(int) e;
  kb_note_launch(2);
  return KB_OK;
}

extern "C" int kb_normalize(int64_t n_rows, int32_t n_desc, const double* d_zeta, const double* d_mean,
                            const double* d_inv_stdev, double* d_out, void* stream)
{
  if (n_rows < 0 || n_desc < 1 || !d_mean || !d_inv_stdev || (n_rows > 0 && (!d_zeta || !d_out))) return KB_ERR_INVALID;
  if (n_rows == 0) return KB_OK;
  const int64_t total = n_rows * n_desc;
  normalize_kernel<<<(unsigned) ((total + 255) / 256), 256, 0, kb_stream(stream)>>>(total, n_desc, d_zeta, d_mean,
                                                                                 d_inv_stdev, d_out);
  KB_LAUNCH_CHECK();
  return KB_OK;
}
