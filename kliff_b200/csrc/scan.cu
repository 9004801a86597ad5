// Device-wide exclusive prefix sum: tile scan -> recursive scan of tile sums -> add-back.
// Small and simple on purpose: the arrays scanned on this path (pad block counts, cell
// occupancies, per-atom neighbor counts) are O(N) ints and the scan is < 1 % of a build.
#include "common.cuh"

namespace {

constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 8;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;  // 2048 elements per CTA

template <typename Tin, typename Tout>
__global__ void __launch_bounds__(SCAN_THREADS)
scan_tiles(const Tin* __restrict__ in, Tout* __restrict__ out, int64_t n, Tout* __restrict__ sums)
{
  __shared__ Tout warp_tot[SCAN_THREADS / 32];
  const int64_t base = (int64_t) blockIdx.x * SCAN_TILE + (int64_t) threadIdx.x * SCAN_ITEMS;
  Tout v[SCAN_ITEMS];
  Tout run = 0;
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; k++)
  {
    Tout x = (base + k < n) ? (Tout) in[base + k] : (Tout) 0;
    v[k] = run;  // exclusive within the thread
    run += x;
  }
  // inclusive warp scan of per-thread totals
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  Tout inc = run;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1)
  {
    Tout y = __shfl_up_sync(0xffffffffu, inc, d);
    if (lane >= d) inc += y;
  }
  if (lane == 31) warp_tot[warp] = inc;
  __syncthreads();
  Tout woff = 0, total = 0;
#pragma unroll
  for (int w = 0; w < SCAN_THREADS / 32; w++)
  {
    Tout t = warp_tot[w];
    if (w < warp) woff += t;
    total += t;
  }
  const Tout off = woff + inc - run;
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; k++)
    if (base + k < n) out[base + k] = v[k] + off;
  if (threadIdx.x == 0 && sums) sums[blockIdx.x] = total;
}

template <typename Tout>
__global__ void __launch_bounds__(SCAN_THREADS)
scan_addback(Tout* __restrict__ out, int64_t n, const Tout* __restrict__ tile_off)
{
  const Tout off = tile_off[blockIdx.x];
  const int64_t base = (int64_t) blockIdx.x * SCAN_TILE + (int64_t) threadIdx.x * SCAN_ITEMS;
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; k++)
    if (base + k < n) out[base + k] += off;
}

template <typename T>
__global__ void scan_write_total(const T* __restrict__ tile_off, const T* __restrict__ tile_sum,
                                 int64_t last_tile, T* __restrict__ total)
{
  *total = tile_off[last_tile] + tile_sum[last_tile];
}

}  // namespace

template <typename Tin, typename Tout>
int kb_exclusive_scan(const Tin* d_in, Tout* d_out, int64_t n, Tout* d_total, cudaStream_t stream)
{
  if (n <= 0)
  {
    if (d_total) KB_CUDA(cudaMemsetAsync(d_total, 0, sizeof(Tout), stream));
    return KB_OK;
  }
  const int64_t tiles = (n + SCAN_TILE - 1) / SCAN_TILE;
  Tout* sums = nullptr;
  KB_CUDA(cudaMallocAsync((void**) &sums, sizeof(Tout) * 2 * (size_t) tiles, stream));
  Tout* offs = sums + tiles;
  scan_tiles<Tin, Tout><<<(unsigned) tiles, SCAN_THREADS, 0, stream>>>(d_in, d_out, n, sums);
  KB_LAUNCH_CHECK();
  if (tiles > 1)
  {
    int rc = kb_exclusive_scan<Tout, Tout>(sums, offs, tiles, (Tout*) nullptr, stream);
    if (rc) return rc;
    scan_addback<Tout><<<(unsigned) tiles, SCAN_THREADS, 0, stream>>>(d_out, n, offs);
    KB_LAUNCH_CHECK();
    if (d_total)
    {
      scan_write_total<Tout><<<1, 1, 0, stream>>>(offs, sums, tiles - 1, d_total);
      KB_LAUNCH_CHECK();
    }
  }
  else if (d_total)
  {
    KB_CUDA(cudaMemcpyAsync(d_total, sums, sizeof(Tout), cudaMemcpyDeviceToDevice, stream));
  }
  KB_CUDA(cudaFreeAsync(sums, stream));
  return KB_OK;
}

template int kb_exclusive_scan<int32_t, int32_t>(const int32_t*, int32_t*, int64_t, int32_t*,
                                                 cudaStream_t);
template int kb_exclusive_scan<int32_t, int64_t>(const int32_t*, int64_t*, int64_t, int64_t*,
                                                 cudaStream_t);
template int kb_exclusive_scan<int64_t, int64_t>(const int64_t*, int64_t*, int64_t, int64_t*,
                                                 cudaStream_t);
