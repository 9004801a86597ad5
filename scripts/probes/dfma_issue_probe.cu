// How many FP64 FMA instructions per cycle does one scheduler sustain from 1, 2, 3 ... resident warps?
// Pattern A: 64 independent accumulators, acc_i = fma(x_i, y, acc_i)   (the shape of the sweep step)
// Pattern B: dependent pairs t = fma(e, t2, t1); acc = fma(E, t, acc)
// Pattern C: A with 10 LDS.128 per 200 DFMA
// One CTA per SM (148 CTAs) of nw*4 warps; cycles by clock64 in warp 0.
#include <cstdio>
#include <cuda_runtime.h>
template <int PAT>
__global__ void k(int iters, double* sink, long long* cyc, double seed)
{
  __shared__ double sh[2048];
  for (int i = threadIdx.x; i < 2048; i += blockDim.x) sh[i] = seed + i;
  __syncthreads();
  double acc[64];
#pragma unroll
  for (int i = 0; i < 64; i++) acc[i] = seed * i + threadIdx.x;
  double x[8];
#pragma unroll
  for (int i = 0; i < 8; i++) x[i] = seed + i + threadIdx.x * 1e-3;
  const double y = seed * 0.999;
  const long long t0 = clock64();
  for (int it = 0; it < iters; it++)
  {
    if (PAT == 2)
    {
      const double2* p = reinterpret_cast<const double2*>(sh + ((it * 20 + (threadIdx.x >> 1) * 82) & 1023));
#pragma unroll
      for (int i = 0; i < 4; i++) { double2 v = p[i]; x[2 * i] += v.x; x[2 * i + 1] += v.y; }
    }
#pragma unroll
    for (int r = 0; r < 3; r++)
#pragma unroll
      for (int i = 0; i < 64; i++)
      {
        if (PAT == 1) acc[i] = fma(x[i & 7], fma(y, x[(i + r) & 7], acc[(i + 32) & 63]), acc[i]);
        else acc[i] = fma(x[(i + r) & 7], y, acc[i]);
      }
  }
  const long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int i = 0; i < 64; i++) s += acc[i];
  if (s == 1234.5678) sink[0] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}
template <int PAT> void run(const char* name)
{
  double* sink; long long* cyc; cudaMalloc(&sink, 8); cudaMalloc(&cyc, 8);
  for (int nw = 1; nw <= 4; nw++)
  {
    const int iters = 2000;
    k<PAT><<<148, nw * 128>>>(iters, sink, cyc, 1.000001);
    k<PAT><<<148, nw * 128>>>(iters, sink, cyc, 1.000001);
    cudaDeviceSynchronize();
    long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    const double nfma = (double) iters * 192 * (PAT == 1 ? 2 : 1);
    printf("%s warps/scheduler %d: %.2f cycles per DFMA per warp, %.3f DFMA/cycle/scheduler\n", name, nw,
           c / nfma, nfma * nw / c);
  }
}
int main()
{
  run<0>("A independent     ");
  run<1>("B dependent pairs ");
  run<2>("C independent+LDS ");
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
