// Training step of small tanh MLPs as six launches (round 2, second form): what
// CalculatorTorch.compute + Loss._get_loss_batch + loss.backward() + Adam.step() do for one batch
// (kliff/legacy/calculators/calculator_torch.py:161-228, kliff/legacy/loss.py:843-920).
//
// The first fused form (train.cu) was thirteen launches per step, thread-per-atom MLP kernels with their
// per-atom vectors in local memory, and an adjoint force kernel reading fp32 blocks one dependent row at a
// time: 205 us per 100-configuration step (ncu launch list: adjoint 80 us, backward 41 + 20, contraction 36,
// forward 26 ...) for 2 x 117 MB of HBM traffic = 36 us.  A fully fused pair of kernels (MLP + block stream
// per warp) was tried first and measured no faster (forward 80 us, backward 101 us): one warp then carries a
// long serial chain (metadata loads, MLP shuffles, chunk waits) and only 8 warps fit beside the chunk
// buffers.  So the MLP and the block passes stay separate launches, each shaped for its own limiter:
//
//   mlp_forward_kernel      warp per atom, 32 warps per SM: MLP forward + dE/dx with lanes = units (hidden
//                           vectors live one element per lane, mat-vecs by shuffles, the padded network in
//                           shared memory) -> e_atom, contraction weights
//   contract_kernel         warp per atom, lanes = block columns, 8 x CG independent loads in flight per
//                           lane -> slot contributions (HBM bound)
//   train_residual_seg_kernel  block per configuration: owner-sorted sum of the slot contributions (fixed
//                           order: bit-reproducible forces), energy, residuals, dL/dF, dL/de, loss term
//   gather_kernel           the adjoint pass over the blocks, 32 x CG independent loads per lane, one
//                           transposing reduction per 32 rows -> G = dL/d(dE/dx) (HBM bound)
//   mlp_backward_kernel     warp per atom, persistent: second-order backward, parameter gradients
//                           accumulated in registers over the warp's atoms, one fixed-order reduction per
//                           block -> partials[block][P]
//   train_update_kernel     warp per parameter: sum of the partials (+ the batch loss), Adam in the same
//                           launch when there is no all-reduce in between
//
// Domain: fan-in <= 64, hidden widths <= 16, <= 4 hidden layers, <= 84 neighbors (3 (n + 1) <= 256 columns);
// anything else stays on train.cu.  Notation as in train.cu:
//   a_l = tanh(W_l a_{l-1} + b_l), d_l = 1 - a_l^2, delta_l = u_l d_l, u_L = w_out, u_{l-1} = W_l^T delta_l
//   P_1 = W_1 G; R_l = P_l d_l; S_l = -2 a_l P_l u_l; P_{l+1} = W_{l+1} R_l
//   da_L = a w_out + S_L; Q_l = da_l d_l; da_{l-1} = W_l^T Q_l + S_{l-1}
//   dW_1 += Q_1 x^T + delta_1 G^T; dW_l += Q_l a_{l-1}^T + delta_l R_{l-1}^T; db_l += Q_l;
//   dw_out += a a_L + R_L; db_out += a
#include "common.cuh"

namespace {

constexpr int TF_H = 16;                      // hidden width limit
constexpr int TF_DP = 64;                     // fan-in limit (two inputs per lane)
constexpr int TF_MAXL = KB_MLP_MAX_HIDDEN;    // hidden layers limit
constexpr int TF_WARPS = 8;
constexpr int TF_MAXCOLS = 256;               // 3 (n + 1) <= 256
constexpr unsigned FULL = 0xffffffffu;

struct FNet
{
  int D, L, H[TF_MAXL];
  int w_off[TF_MAXL + 1], b_off[TF_MAXL + 1];  // offsets into the flat (torch-ordered) parameters
  int P;
};

__host__ __device__ inline FNet make_fnet(const kb_mlp_desc& M)
{
  FNet N;
  N.D = M.n_in;
  N.L = M.n_hidden;
  int pos = 0, fan_in = M.n_in;
  for (int l = 0; l < TF_MAXL; l++) N.H[l] = 0;
  for (int l = 0; l < N.L; l++)
  {
    N.H[l] = M.width[l];
    N.w_off[l] = pos; pos += N.H[l] * fan_in;
    N.b_off[l] = pos; pos += N.H[l];
    fan_in = N.H[l];
  }
  N.w_off[N.L] = pos; pos += fan_in;
  N.b_off[N.L] = pos; pos += 1;
  for (int l = N.L + 1; l <= TF_MAXL; l++) N.w_off[l] = N.b_off[l] = pos;
  N.P = pos;
  return N;
}

// Padded shared-memory image of the network (elements of T); everything outside the real shapes is zero,
// so lanes / loop trips beyond a layer's width contribute exactly 0 and no guard is needed:
constexpr int SN_W0 = 0;                                      // [TF_H][TF_DP]        W_1[h][d]
constexpr int SN_B = SN_W0 + TF_H * TF_DP;                    // [TF_MAXL][TF_H]      hidden biases
constexpr int SN_WR = SN_B + TF_MAXL * TF_H;                  // [TF_MAXL-1][i][j]    W_{l+1}[i][j], l >= 1
constexpr int SN_WC = SN_WR + (TF_MAXL - 1) * TF_H * TF_H;    // [TF_MAXL-1][j][i]    the same, transposed
constexpr int SN_WO = SN_WC + (TF_MAXL - 1) * TF_H * TF_H;    // [TF_H]               w_out
constexpr int SN_BO = SN_WO + TF_H;                           // [1] (+ padding)
constexpr int SN_SIZE = SN_BO + 16;

// thread per element of the padded image: the source index follows from shifts (padded extents are powers
// of two), one independent global load each.  (The first form walked the flat parameters with a division
// per element: as many instructions per block as the eight atoms' worth of MLP work it serves.)
template <typename T>
__device__ void load_net(const FNet& N, const T* __restrict__ theta, T* sn)
{
  for (int q = threadIdx.x; q < SN_SIZE; q += blockDim.x)
  {
    int src = -1;
    if (q < SN_B)
    {
      const int i = q >> 6, j = q & 63;
      if (i < N.H[0] && j < N.D) src = N.w_off[0] + i * N.D + j;
    }
    else if (q < SN_WR)
    {
      const int r = q - SN_B, l = r >> 4, i = r & 15;
      if (l < N.L && i < N.H[l]) src = N.b_off[l] + i;
    }
    else if (q < SN_WO)
    {
      const bool tr = q >= SN_WC;  // transposed copy: [j][i]
      const int r = q - (tr ? SN_WC : SN_WR), l = (r >> 8) + 1, hi = (r >> 4) & 15, lo = r & 15;
      const int i = tr ? lo : hi, j = tr ? hi : lo;
      if (l < N.L && i < N.H[l] && j < N.H[l - 1]) src = N.w_off[l] + i * N.H[l - 1] + j;
    }
    else if (q < SN_BO)
    {
      const int j = q - SN_WO;
      if (j < N.H[N.L - 1]) src = N.w_off[N.L] + j;
    }
    else if (q == SN_BO) src = N.b_off[N.L];
    sn[q] = src >= 0 ? theta[src] : T(0);
  }
  __syncthreads();
}

template <typename T> __device__ __forceinline__ T tanh_f(T x);
template <> __device__ __forceinline__ float tanh_f<float>(float x) { return tanhf(x); }
template <> __device__ __forceinline__ double tanh_f<double>(double x) { return tanh(x); }

template <typename T>
__device__ __forceinline__ T warp_sum_t(T v)
{
#pragma unroll
  for (int m = 16; m > 0; m >>= 1) v += __shfl_xor_sync(FULL, v, m);
  return v;
}

// 16 per-lane values -> lane l holds the warp-wide total of v[l & 15] (31 exchanges instead of 80)
template <typename T>
__device__ __forceinline__ T warp_reduce16_t(T (&v)[16], int lane)
{
#pragma unroll
  for (int i = 0; i < 16; i++) v[i] += __shfl_xor_sync(FULL, v[i], 16);
#pragma unroll
  for (int half = 8; half >= 1; half >>= 1)
  {
    const bool upper = (lane & half) != 0;
#pragma unroll
    for (int i = 0; i < half; i++)
    {
      const T keep = upper ? v[i + half] : v[i];
      const T send = upper ? v[i] : v[i + half];
      v[i] = keep + __shfl_xor_sync(FULL, send, half);
    }
  }
  return v[0];
}

// y[h] = sum_d W_1[h][d] v[d], v distributed two per lane (d = lane, lane + 32); result in lane h < 16
template <typename T>
__device__ __forceinline__ T first_layer_matvec(const FNet& N, const T* sn, T v0, T v1, int lane)
{
  T p[16];
#pragma unroll
  for (int h = 0; h < 16; h++)
  {
    p[h] = T(0);
    if (h < N.H[0]) p[h] = fma(sn[SN_W0 + h * TF_DP + lane], v0, sn[SN_W0 + h * TF_DP + 32 + lane] * v1);
  }
  const T y = warp_reduce16_t(p, lane);
  return lane < 16 ? y : T(0);
}

// y[i] = sum_j M[j][i] v[j] for a [TF_H][TF_H] table M in shared memory, v one element per lane (< 16)
template <typename T>
__device__ __forceinline__ T hidden_matvec(const T* M, T v, T y0, int lane)
{
  const int li = lane & 15;
  T y = y0;
#pragma unroll
  for (int j = 0; j < 16; j++) y = fma(M[j * TF_H + li], __shfl_sync(FULL, v, j), y);
  return lane < 16 ? y : T(0);
}

// forward pass of one atom by one warp; a[l] = activation of hidden layer l in lane h < 16; returns e
template <typename T, int LT>
__device__ __forceinline__ T mlp_forward(const FNet& N, const T* sn, T x0, T x1, int lane, T (&a)[LT])
{
  const int li = lane & 15;
  const bool lo16 = lane < 16;
  T z = first_layer_matvec<T>(N, sn, x0, x1, lane) + sn[SN_B + li];
  a[0] = lo16 ? tanh_f<T>(z) : T(0);
#pragma unroll
  for (int l = 1; l < LT; l++)
  {
    z = hidden_matvec<T>(sn + SN_WC + (l - 1) * TF_H * TF_H, a[l - 1], sn[SN_B + l * TF_H + li], lane);
    a[l] = lo16 ? tanh_f<T>(z) : T(0);
  }
  return warp_sum_t<T>(sn[SN_WO + li] * a[LT - 1]) + sn[SN_BO];
}

// u_l and delta_l of every hidden layer
template <typename T, int LT>
__device__ __forceinline__ void mlp_deltas(const T* sn, const T (&a)[LT], int lane, T (&u)[LT], T (&delta)[LT])
{
  const int li = lane & 15;
  u[LT - 1] = lane < 16 ? sn[SN_WO + li] : T(0);
#pragma unroll
  for (int l = LT - 1; l >= 0; l--)
  {
    delta[l] = u[l] * (T(1) - a[l] * a[l]);
    if (l > 0) u[l - 1] = hidden_matvec<T>(sn + SN_WR + (l - 1) * TF_H * TF_H, delta[l], T(0), lane);  // W_l^T delta_l
  }
}

template <typename T> __host__ __device__ constexpr size_t net_smem_bytes() { return (sizeof(T) * SN_SIZE + 15) & ~(size_t) 15; }

// =====================================================================================================
// launch 1: e_atom and the contraction weights w[t][d] = dE/dx[d] * scale[d]; warp per atom
// =====================================================================================================
template <typename T, int LT>
__global__ void __launch_bounds__(TF_WARPS * 32)
mlp_forward_kernel(const __grid_constant__ FNet N, int n, const T* __restrict__ x, const T* __restrict__ theta,
                   const double* __restrict__ scale, double* __restrict__ e_atom, double* __restrict__ w)
{
  extern __shared__ __align__(16) unsigned char sm_raw[];
  T* sn = reinterpret_cast<T*>(sm_raw);
  kb_grid_dep_launch();
  kb_grid_dep_wait();  // theta comes from the previous step's update launch
  load_net<T>(N, theta, sn);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int D = N.D;
  const double sc0 = (scale && lane < D) ? scale[lane] : 1.0, sc1 = (scale && 32 + lane < D) ? scale[32 + lane] : 1.0;
  for (int t = blockIdx.x * TF_WARPS + warp; t < n; t += gridDim.x * TF_WARPS)
  {
    const T x0 = lane < D ? x[(int64_t) t * D + lane] : T(0);
    const T x1 = 32 + lane < D ? x[(int64_t) t * D + 32 + lane] : T(0);
    T a[LT];
    const T e = mlp_forward<T, LT>(N, sn, x0, x1, lane, a);
    if (lane == 0) e_atom[t] = (double) e;
    if (!w) continue;
    T u[LT], delta[LT];
    mlp_deltas<T, LT>(sn, a, lane, u, delta);
    T g0 = T(0), g1 = T(0);  // dE/dx[d] = sum_h W_1[h][d] delta_1[h]
#pragma unroll
    for (int h = 0; h < 16; h++)
      if (h < N.H[0])
      {
        const T dh = __shfl_sync(FULL, delta[0], h);
        g0 = fma(sn[SN_W0 + h * TF_DP + lane], dh, g0);
        g1 = fma(sn[SN_W0 + h * TF_DP + 32 + lane], dh, g1);
      }
    if (lane < D) w[(int64_t) t * D + lane] = (double) g0 * sc0;
    if (32 + lane < D) w[(int64_t) t * D + 32 + lane] = (double) g1 * sc1;
  }
}

// ---- the two passes over the dG/dr blocks: warp per atom, lanes = columns (slot, c) of the [D][3 (n + 1)]
// block, every load unconditional and independent of the others (columns past the row re-read column 0,
// rows past D re-read row D - 1), so a lane keeps 8 x CG (contraction) or 32 x CG (adjoint) loads in
// flight.  fp32 blocks are multiplied and summed in fp32 over at most 8 terms before they enter the fp64
// accumulators: one float -> double conversion per 8 elements instead of one per element.
template <typename TDz> struct PartT { typedef double type; };
template <> struct PartT<float> { typedef float type; };

template <typename TDz, int CG, int RB>
__device__ __forceinline__ void contract_rows(const TDz* __restrict__ blk, int row, int D,
                                              const double* __restrict__ wglob, double* wrow, int lane,
                                              double* __restrict__ out)
{
  typedef typename PartT<TDz>::type A;
  int cofs[CG];
#pragma unroll
  for (int q = 0; q < CG; q++) cofs[q] = q * 32 + lane < row ? q * 32 + lane : 0;
  double acc[CG];
#pragma unroll
  for (int q = 0; q < CG; q++) acc[q] = 0.0;
  for (int d0 = 0; d0 < D; d0 += RB)
  {
    TDz v[RB][CG];
#pragma unroll
    for (int r = 0; r < RB; r++)
    {
      const TDz* rp = blk + (d0 + r < D ? d0 + r : D - 1) * row;
#pragma unroll
      for (int q = 0; q < CG; q++) v[r][q] = rp[cofs[q]];
    }
    if (d0 == 0)
    {
      // the first batch of block rows (written by no kernel of the step) is in flight before this kernel
      // waits for its predecessor, the MLP forward launch, whose weights are fetched only now
      kb_grid_dep_wait();
      if (lane < D) wrow[lane] = wglob[lane];
      if (32 + lane < D) wrow[32 + lane] = wglob[32 + lane];
      __syncwarp();
    }
#pragma unroll
    for (int r0 = 0; r0 < RB; r0 += 8)
    {
      A part[CG];
#pragma unroll
      for (int q = 0; q < CG; q++) part[q] = A(0);
#pragma unroll
      for (int r = r0; r < r0 + 8; r++)
      {
        const A wr = d0 + r < D ? (A) wrow[d0 + r] : A(0);
#pragma unroll
        for (int q = 0; q < CG; q++) part[q] = fma(wr, (A) v[r][q], part[q]);
      }
#pragma unroll
      for (int q = 0; q < CG; q++) acc[q] += (double) part[q];
    }
  }
#pragma unroll
  for (int q = 0; q < CG; q++)
    if (q * 32 + lane < row) out[q * 32 + lane] = acc[q];
}

// launch 2: slot contributions contrib[slot][c] = sum_d w[t][d] dz_t[d][slot][c]
// (a centre with n neighbors owns n + 1 consecutive slots: its row length comes from slot_ptr alone)
template <typename TDz>
__global__ void __launch_bounds__(TF_WARPS * 32)
contract_kernel(int n, int D, const double* __restrict__ w, const TDz* __restrict__ dz,
                const int64_t* __restrict__ dz_ptr, const int64_t* __restrict__ slot_ptr,
                double* __restrict__ contrib)
{
  __shared__ double wsm[TF_WARPS][TF_DP];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int t = blockIdx.x * TF_WARPS + warp;
  kb_grid_dep_launch();
  if (t >= n) return;
  const int64_t sl = slot_ptr[t];
  const int row = 3 * (int) (slot_ptr[t + 1] - sl);
  const TDz* blk = dz + dz_ptr[t];
  double* out = contrib + 3 * (sl - slot_ptr[0]);
  const double* wg = w + (int64_t) t * D;
  const int ncg = (row + 31) >> 5;
  if (ncg <= 2) contract_rows<TDz, 2, 16>(blk, row, D, wg, wsm[warp], lane, out);
  else if (ncg == 3) contract_rows<TDz, 3, 16>(blk, row, D, wg, wsm[warp], lane, out);
  else if (ncg == 4) contract_rows<TDz, 4, 8>(blk, row, D, wg, wsm[warp], lane, out);
  else if (ncg <= 6) contract_rows<TDz, 6, 8>(blk, row, D, wg, wsm[warp], lane, out);
  else contract_rows<TDz, 8, 8>(blk, row, D, wg, wsm[warp], lane, out);
}

// =====================================================================================================
// launch 3: one block per configuration — forces (owner-sorted sums), energy, residuals, loss term
// =====================================================================================================
constexpr int RS_WARPS = 32;
__global__ void __launch_bounds__(RS_WARPS * 32)
train_residual_seg_kernel(int n_cfg, const int64_t* __restrict__ cfg_ptr, int64_t atom0,
                          const double* __restrict__ e_atom, const int64_t* __restrict__ seg_ptr,
                          const int32_t* __restrict__ seg_src, int64_t slot_base,
                          const double* contrib, double* gslot, const double* __restrict__ e_ref,
                          const double* __restrict__ f_ref, const double* __restrict__ w_e,
                          const double* __restrict__ w_f, double inv_nglobal, int use_energy, int use_forces,
                          double* __restrict__ a_out, double* __restrict__ gF, double* __restrict__ forces,
                          double* __restrict__ loss_cfg, double* __restrict__ step)
{
  __shared__ double red[RS_WARPS];
  const int c = blockIdx.x;
  kb_grid_dep_launch();
  if (c >= n_cfg)
  {
    kb_grid_dep_wait();
    if (c == 0 && threadIdx.x == 0 && step) step[0] += 1.0;
    return;
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int lo = (int) (cfg_ptr[c] - atom0), hi = (int) (cfg_ptr[c + 1] - atom0);
  // everything addressable from (c, lo, hi) is requested here in one batch: this kernel is a chain of
  // dependent memory round trips (configuration range -> segment bounds -> slot ids -> contributions)
  const int o_first = lo + warp, o_second = o_first + RS_WARPS;
  const bool f1 = use_forces && o_first < hi, f2 = use_forces && o_second < hi;
  const int64_t pre_s0 = f1 ? seg_ptr[o_first] : 0, pre_s1 = f1 ? seg_ptr[o_first + 1] : 0;
  const int64_t pre_t0 = f2 ? seg_ptr[o_second] : 0, pre_t1 = f2 ? seg_ptr[o_second + 1] : 0;
  const double we = use_energy ? w_e[c] : 0.0, eref = e_ref[c];
  const double wf = use_forces ? w_f[c] : 0.0;
  // slot ids of the first trip: the plan is constant, so these loads too run ahead of the predecessor's end
  int64_t pre_ia = 0, pre_ib = 0;
  if (f1 && pre_s0 + lane < pre_s1) pre_ia = (int64_t) seg_src[pre_s0 + lane] - slot_base;
  if (f2 && pre_t0 + lane < pre_t1) pre_ib = (int64_t) seg_src[pre_t0 + lane] - slot_base;
  kb_grid_dep_wait();  // e_atom and the slot contributions come from the two launches before
  if (c == 0 && threadIdx.x == 0 && step) step[0] += 1.0;  // Adam's step counter, read by the update launch
  double s = 0.0;
  for (int q = lo + threadIdx.x; q < hi; q += blockDim.x) s += e_atom[q];
  s = warp_sum(s);
  if (lane == 0) red[warp] = s;
  __syncthreads();
  double E = 0.0;
#pragma unroll
  for (int w = 0; w < RS_WARPS; w++) E += red[w];
  __syncthreads();
  const double re = we * (E - eref);  // residual[0] (loss.py:100-109)
  const double a = 2.0 * we * re * inv_nglobal;
  for (int q = lo + threadIdx.x; q < hi; q += blockDim.x) a_out[q] = a;
  double part = 0.0;  // lanes 0-2 of every warp: rf^2 of their component over the warp's owners, in owner order
  if (use_forces)
  {
    // two owners per trip: their segment bounds, slot ids and contributions travel together
    for (int o = lo + warp; o < hi; o += 2 * RS_WARPS)
    {
      const int o2 = o + RS_WARPS;
      const bool two = o2 < hi;
      const bool first = o == o_first;
      const int64_t s0 = first ? pre_s0 : seg_ptr[o], s1 = first ? pre_s1 : seg_ptr[o + 1];
      const int64_t t0 = first ? pre_t0 : (two ? seg_ptr[o2] : 0), t1 = first ? pre_t1 : (two ? seg_ptr[o2 + 1] : 0);
      const double fra = lane < 3 ? f_ref[3 * (int64_t) o + lane] : 0.0;
      const double frb = (lane < 3 && two) ? f_ref[3 * (int64_t) o2 + lane] : 0.0;
      double f[2][3] = {{0.0, 0.0, 0.0}, {0.0, 0.0, 0.0}};
      for (int64_t i = lane; i < max(s1 - s0, t1 - t0); i += 32)
      {
        const bool ina = s0 + i < s1, inb = t0 + i < t1;
        const bool pre = first && i == lane;
        const int64_t ia = pre ? pre_ia : (ina ? (int64_t) seg_src[s0 + i] - slot_base : 0);
        const int64_t ib = pre ? pre_ib : (inb ? (int64_t) seg_src[t0 + i] - slot_base : 0);
        const double* pa = contrib + 3 * ia;
        const double* pb = contrib + 3 * ib;
        const double ax = pa[0], ay = pa[1], az = pa[2], bx = pb[0], by = pb[1], bz = pb[2];
        if (ina) { f[0][0] += ax; f[0][1] += ay; f[0][2] += az; }
        if (inb) { f[1][0] += bx; f[1][1] += by; f[1][2] += bz; }
      }
      double g[2][3];
#pragma unroll
      for (int k = 0; k < 2; k++)
      {
        const int ow = k ? o2 : o;
        const double fx = warp_sum(f[k][0]), fy = warp_sum(f[k][1]), fz = warp_sum(f[k][2]);
        double gl = 0.0;
        if (lane < 3 && (k == 0 || two))
        {
          const double F = -(lane == 0 ? fx : lane == 1 ? fy : fz);
          const double rf = wf * (F - (k ? frb : fra));
          forces[3 * (int64_t) ow + lane] = F;
          gl = 2.0 * wf * rf * inv_nglobal;
          gF[3 * (int64_t) ow + lane] = gl;
          part += rf * rf;
        }
        g[k][0] = __shfl_sync(FULL, gl, 0); g[k][1] = __shfl_sync(FULL, gl, 1); g[k][2] = __shfl_sync(FULL, gl, 2);
      }
      // dL/dF of the owner back into its slots: the adjoint pass then reads them as one coalesced row
      for (int64_t i = lane; i < max(s1 - s0, t1 - t0); i += 32)
      {
        if (s0 + i < s1)
        {
          double* pa = gslot + 3 * ((int64_t) seg_src[s0 + i] - slot_base);
          pa[0] = g[0][0]; pa[1] = g[0][1]; pa[2] = g[0][2];
        }
        if (t0 + i < t1)
        {
          double* pb = gslot + 3 * ((int64_t) seg_src[t0 + i] - slot_base);
          pb[0] = g[1][0]; pb[1] = g[1][1]; pb[2] = g[1][2];
        }
      }
    }
    part += __shfl_down_sync(FULL, part, 1) + __shfl_down_sync(FULL, part, 2);  // lane 0: x + y + z
  }
  if (lane == 0) red[warp] = part;
  __syncthreads();
  if (threadIdx.x == 0)
  {
    double tot = re * re;
#pragma unroll
    for (int w = 0; w < RS_WARPS; w++) tot += red[w];
    loss_cfg[c] = tot * inv_nglobal;
  }
}

// =====================================================================================================
// launch 4: G[t][d] = dL/d(dE/dx) = -scale[d] sum_{slot, c} dz_t[d][slot][c] dL/dF[owner(slot)][c]
// =====================================================================================================
template <typename TDz, int CG>
__device__ __forceinline__ void gather_rows(const TDz* __restrict__ blk, int row, int D, const double* gfs,
                                            const double* __restrict__ scale, int lane, double* __restrict__ Gout)
{
  typedef typename PartT<TDz>::type A;
  int cofs[CG];
  A g[CG];
#pragma unroll
  for (int q = 0; q < CG; q++)
  {
    const bool on = q * 32 + lane < row;
    cofs[q] = on ? q * 32 + lane : 0;
    g[q] = on ? (A) gfs[q * 32 + lane] : A(0);
  }
  for (int d0 = 0; d0 < D; d0 += 32)
  {
    double s[32];
#pragma unroll
    for (int r = 0; r < 32; r++)
    {
      const TDz* rp = blk + (d0 + r < D ? d0 + r : D - 1) * row;
      A acc = A(0);
#pragma unroll
      for (int q = 0; q < CG; q++) acc = fma(g[q], (A) rp[cofs[q]], acc);
      s[r] = (double) acc;
    }
    const double tot = warp_reduce32(s, lane);
    const int d = d0 + lane;
    if (d < D) Gout[d] = -tot * (scale ? scale[d] : 1.0);
  }
}

template <typename TDz>
__global__ void __launch_bounds__(TF_WARPS * 32)
gather_kernel(int n, int D, const double* __restrict__ gslot, const double* __restrict__ scale,
              const TDz* __restrict__ dz, const int64_t* __restrict__ dz_ptr,
              const int64_t* __restrict__ slot_ptr, double* __restrict__ G)
{
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int t = blockIdx.x * TF_WARPS + warp;
  if (t >= n) return;
  const int64_t sl = slot_ptr[t];
  const int row = 3 * (int) (slot_ptr[t + 1] - sl);
  const TDz* blk = dz + dz_ptr[t];
  const double* gfs = gslot + 3 * (sl - slot_ptr[0]);  // dL/dF of this centre's slots, written by launch 3
  double* Gout = G + (int64_t) t * D;
  const int ncg = (row + 31) >> 5;
  if (ncg <= 2) gather_rows<TDz, 2>(blk, row, D, gfs, scale, lane, Gout);
  else if (ncg == 3) gather_rows<TDz, 3>(blk, row, D, gfs, scale, lane, Gout);
  else if (ncg == 4) gather_rows<TDz, 4>(blk, row, D, gfs, scale, lane, Gout);
  else if (ncg <= 6) gather_rows<TDz, 6>(blk, row, D, gfs, scale, lane, Gout);
  else gather_rows<TDz, 8>(blk, row, D, gfs, scale, lane, Gout);
}

// =====================================================================================================
// launch 5: second-order backward per atom (warp per atom, persistent).  Parameter gradients accumulate over
// the warp's atoms in a per-warp shared-memory image of the flat parameter vector (every element is owned by
// one lane, so plain read-modify-writes: no atomics); keeping them out of the registers lets 24 warps
// instead of 8 share an SM (ncu on the register form: 157 registers, 8 warps, 38 us of pure latency).
// One fixed-order sum over the block's warps -> partials[block][P].
// =====================================================================================================
constexpr int BW_BLOCKS = 3;  // resident blocks per SM
template <typename T, int LT>
__global__ void __launch_bounds__(TF_WARPS * 32, BW_BLOCKS)
mlp_backward_kernel(const __grid_constant__ FNet N, int n, const T* __restrict__ x, const T* __restrict__ theta,
                    const double* __restrict__ G, const double* __restrict__ a_in, double* __restrict__ partials)
{
  extern __shared__ __align__(16) unsigned char sm_raw[];
  T* sn = reinterpret_cast<T*>(sm_raw);
  load_net<T>(N, theta, sn);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int li = lane & 15;
  const bool lo16 = lane < 16;
  const int D = N.D, P = N.P;
  T* accAll = reinterpret_cast<T*>(sm_raw + net_smem_bytes<T>());  // [TF_WARPS][P]
  T* accS = accAll + warp * P;
  for (int q = lane; q < P; q += 32) accS[q] = T(0);
  __syncwarp();
  const int H0 = N.H[0];
  const bool in0 = lane < D, in1 = 32 + lane < D;

  for (int t = blockIdx.x * TF_WARPS + warp; t < n; t += gridDim.x * TF_WARPS)
  {
    const T x0 = in0 ? x[(int64_t) t * D + lane] : T(0);
    const T x1 = in1 ? x[(int64_t) t * D + 32 + lane] : T(0);
    const T G0 = (G && in0) ? (T) G[(int64_t) t * D + lane] : T(0);
    const T G1 = (G && in1) ? (T) G[(int64_t) t * D + 32 + lane] : T(0);
    const T aT = (T) a_in[t];
    T a[LT], u[LT], delta[LT];
    mlp_forward<T, LT>(N, sn, x0, x1, lane, a);
    mlp_deltas<T, LT>(sn, a, lane, u, delta);
    // second-order part, up the layers
    T R[LT], Sv[LT];
    T Pl = G ? first_layer_matvec<T>(N, sn, G0, G1, lane) : T(0);
#pragma unroll
    for (int l = 0; l < LT; l++)
    {
      R[l] = Pl * (T(1) - a[l] * a[l]);
      Sv[l] = T(-2) * a[l] * Pl * u[l];
      if (l + 1 < LT) Pl = hidden_matvec<T>(sn + SN_WC + l * TF_H * TF_H, R[l], T(0), lane);  // W_{l+1} R_l
    }
    // first-order backward with the extra sources S_l
    T Q[LT];
    T da = lo16 ? fma(aT, sn[SN_WO + li], Sv[LT - 1]) : T(0);
#pragma unroll
    for (int l = LT - 1; l >= 0; l--)
    {
      Q[l] = da * (T(1) - a[l] * a[l]);
      if (l > 0) da = hidden_matvec<T>(sn + SN_WR + (l - 1) * TF_H * TF_H, Q[l], Sv[l - 1], lane);  // W_l^T Q_l + S_{l-1}
    }
    // parameter gradients of this atom into the warp's running sums
    {
      T* w0 = accS + N.w_off[0] + lane;
#pragma unroll
      for (int h = 0; h < 16; h++)
        if (h < H0)
        {
          const T qh = __shfl_sync(FULL, Q[0], h), dh = __shfl_sync(FULL, delta[0], h);
          if (in0) w0[h * D] += fma(qh, x0, dh * G0);
          if (in1) w0[h * D + 32] += fma(qh, x1, dh * G1);
        }
    }
#pragma unroll
    for (int l = 0; l < LT; l++)
      if (lo16 && li < N.H[l]) accS[N.b_off[l] + li] += Q[l];
#pragma unroll
    for (int l = 1; l < LT; l++)
    {
      const int Hp = N.H[l - 1];
      const bool mine = lo16 && li < N.H[l];
      T* wl = accS + N.w_off[l] + li * Hp;
#pragma unroll
      for (int j = 0; j < 16; j++)
        if (j < Hp)
        {
          const T aj = __shfl_sync(FULL, a[l - 1], j), rj = __shfl_sync(FULL, R[l - 1], j);
          if (mine) wl[j] += fma(Q[l], aj, delta[l] * rj);
        }
    }
    if (lo16 && li < N.H[LT - 1]) accS[N.w_off[LT] + li] += fma(aT, a[LT - 1], R[LT - 1]);
    if (lane == 0) accS[N.b_off[LT]] += aT;
  }
  __syncthreads();
  for (int q = threadIdx.x; q < P; q += blockDim.x)
  {
    double sum = 0.0;
#pragma unroll
    for (int w = 0; w < TF_WARPS; w++) sum += (double) accAll[w * P + q];
    partials[(int64_t) blockIdx.x * P + q] = sum;
  }
}

// torch.optim.Adam (no amsgrad, no weight decay), torch/optim/adam.py `_single_tensor_adam`, for ONE
// parameter; t = the step number of this update (already counted by launch 2)
template <typename T>
__device__ __forceinline__ void adam_one(int p, double gd, T* theta, T* m, T* v, double t, double lr, double b1,
                                         double b2, double eps)
{
  const T g = (T) gd;
  const T mm = m[p] + (T) (1.0 - b1) * (g - m[p]);           // exp_avg.lerp_(grad, 1 - beta1)
  const T vv = (T) b2 * v[p] + (T) (1.0 - b2) * g * g;       // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, 1 - beta2)
  m[p] = mm;
  v[p] = vv;
  const double bc1 = 1.0 - pow(b1, t), bc2 = 1.0 - pow(b2, t);
  const T denom = (T) sqrt((double) vv) / (T) sqrt(bc2) + (T) eps;
  theta[p] = theta[p] + (T) (-lr / bc1) * (mm / denom);      // param.addcdiv_(exp_avg, denom, value=-step_size)
}

// launch 6: grad[p] = sum of the partial rows in a fixed order, grad[P] = sum of the loss terms; a block
// takes 32 consecutive parameters, its 32 warps every 32nd row each (coalesced 256-byte reads), then warp 0 adds
// the 32 sub-sums.  With `adam` the update follows in the same launch (single rank: no all-reduce in between).
constexpr int UP_WARPS = 32;
template <typename T>
__global__ void __launch_bounds__(UP_WARPS * 32)
train_update_kernel(int P, int rows, const double* __restrict__ partials, int n_cfg,
                    const double* __restrict__ loss_cfg, double* __restrict__ grad, int adam, T* theta, T* m, T* v,
                    const double* __restrict__ step, double lr, double b1, double b2, double eps,
                    double* __restrict__ epoch_loss)
{
  __shared__ double sub[UP_WARPS][32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int p = blockIdx.x * 32 + lane;
  double s = 0.0;
  if (p < P)
  {
    // rows warp, warp + 32, ...: four independent loads per trip (a plain loop is one memory round trip per row)
    int r = warp;
    for (; r + 3 * UP_WARPS < rows; r += 4 * UP_WARPS)
    {
      const double v0 = partials[(int64_t) r * P + p], v1 = partials[(int64_t) (r + UP_WARPS) * P + p];
      const double v2 = partials[(int64_t) (r + 2 * UP_WARPS) * P + p], v3 = partials[(int64_t) (r + 3 * UP_WARPS) * P + p];
      s += (v0 + v1) + (v2 + v3);
    }
    for (; r < rows; r += UP_WARPS) s += partials[(int64_t) r * P + p];
  }
  else if (p == P)
    for (int c = warp; c < n_cfg; c += UP_WARPS) s += loss_cfg[c];
  sub[warp][lane] = s;
  __syncthreads();
  if (warp != 0 || p > P) return;
  s = 0.0;
#pragma unroll
  for (int w = 0; w < UP_WARPS; w++) s += sub[w][lane];
  grad[p] = s;
  if (p == P) { if (epoch_loss) epoch_loss[0] += s; }
  else if (adam) adam_one<T>(p, s, theta, m, v, step[0], lr, b1, b2, eps);
}

// after an all-reduce of grad[P + 1]: Adam on every parameter, epoch loss
template <typename T>
__global__ void train_adam_kernel(int P, const double* __restrict__ grad, T* theta, T* m, T* v,
                                  const double* __restrict__ step, double lr, double b1, double b2, double eps,
                                  double* __restrict__ epoch_loss)
{
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p < P) adam_one<T>(p, grad[p], theta, m, v, step[0], lr, b1, b2, eps);
  else if (p == P && epoch_loss) epoch_loss[0] += grad[P];
}

int fast_check(const kb_mlp_desc* M)
{
  if (!M) return KB_ERR_INVALID;
  if (M->dtype != KB_F32 && M->dtype != KB_F64) return KB_ERR_INVALID;
  if (M->n_in < 1 || M->n_in > TF_DP) return KB_ERR_LIMIT;
  if (M->n_hidden < 1 || M->n_hidden > TF_MAXL) return KB_ERR_LIMIT;
  for (int l = 0; l < M->n_hidden; l++)
    if (M->width[l] < 1 || M->width[l] > TF_H) return KB_ERR_LIMIT;
  return KB_OK;
}

int sm_count()
{
  static int n = 0;
  if (n == 0)
  {
    int dev = 0, v = 0;
    if (cudaGetDevice(&dev) == cudaSuccess &&
        cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && v > 0)
      n = v;
    else return 148;
  }
  return n;
}

template <typename T, int LT>
int launch_mlp_forward(const kb_mlp_desc& M, int n, const void* x, const void* theta, const double* scale,
                       double* e_atom, double* w, cudaStream_t st)
{
  const size_t smem = net_smem_bytes<T>();
  const int nb = min(8 * sm_count(), (n + TF_WARPS - 1) / TF_WARPS);
  cudaError_t e = kb_launch_pdl(mlp_forward_kernel<T, LT>, dim3(nb), dim3(TF_WARPS * 32), smem, st, make_fnet(M), n,
                                static_cast<const T*>(x), static_cast<const T*>(theta), scale, e_atom, w);
  return e == cudaSuccess ? KB_OK : -(int) e;
}

template <typename T, int LT>
int launch_mlp_backward(const kb_mlp_desc& M, int n, int nb, const void* x, const void* theta, const double* G,
                        const double* a, double* partials, cudaStream_t st)
{
  const FNet N = make_fnet(M);
  const size_t smem = net_smem_bytes<T>() + sizeof(T) * TF_WARPS * (size_t) N.P;
  if (smem > 48 * 1024)
  {
    cudaError_t e = cudaFuncSetAttribute(mlp_backward_kernel<T, LT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);
    if (e != cudaSuccess) return -(int) e;
  }
  mlp_backward_kernel<T, LT><<<nb, TF_WARPS * 32, smem, st>>>(N, n, static_cast<const T*>(x),
                                                             static_cast<const T*>(theta), G, a, partials);
  return KB_OK;
}

#define KB_DISPATCH_LT(FN, T, L, ...)                      \
  ((L) == 1 ? FN<T, 1>(__VA_ARGS__) : (L) == 2 ? FN<T, 2>(__VA_ARGS__) \
   : (L) == 3 ? FN<T, 3>(__VA_ARGS__) : FN<T, 4>(__VA_ARGS__))

}  // namespace

extern "C" int kb_train_fast_ok(const kb_mlp_desc* M, int32_t maxn)
{
  int rc = fast_check(M);
  if (rc) return rc;
  if (maxn < 0 || 3 * (maxn + 1) > TF_MAXCOLS) return KB_ERR_LIMIT;
  return KB_OK;
}

extern "C" int kb_train_forward(const kb_mlp_desc* M, int32_t n_atoms, const void* d_x, const void* d_theta,
                                const double* d_scale, const void* d_dz, int32_t dz_dtype,
                                const int64_t* d_dz_ptr, const int64_t* d_slot_ptr, double* d_e_atom,
                                double* d_w, double* d_contrib, void* stream)
{
  int rc = fast_check(M);
  if (rc) return rc;
  if (n_atoms < 0 || !d_x || !d_theta || !d_e_atom) return KB_ERR_INVALID;
  if (d_contrib && (!d_w || !d_dz || !d_dz_ptr || !d_slot_ptr)) return KB_ERR_INVALID;
  if (d_contrib && dz_dtype != KB_F32 && dz_dtype != KB_F64) return KB_ERR_INVALID;
  if (n_atoms == 0) return KB_OK;
  cudaStream_t st = kb_stream(stream);
  double* w = d_contrib ? d_w : nullptr;
  if (M->dtype == KB_F32) rc = KB_DISPATCH_LT(launch_mlp_forward, float, M->n_hidden, *M, n_atoms, d_x, d_theta, d_scale, d_e_atom, w, st);
  else rc = KB_DISPATCH_LT(launch_mlp_forward, double, M->n_hidden, *M, n_atoms, d_x, d_theta, d_scale, d_e_atom, w, st);
  if (rc) return rc;
  KB_LAUNCH_CHECK();
  if (!d_contrib) return KB_OK;
  const int nb = (n_atoms + TF_WARPS - 1) / TF_WARPS;
  if (dz_dtype == KB_F32)
    KB_CUDA(kb_launch_pdl(contract_kernel<float>, dim3(nb), dim3(TF_WARPS * 32), 0, st, n_atoms, M->n_in, d_w,
                          static_cast<const float*>(d_dz), d_dz_ptr, d_slot_ptr, d_contrib));
  else
    KB_CUDA(kb_launch_pdl(contract_kernel<double>, dim3(nb), dim3(TF_WARPS * 32), 0, st, n_atoms, M->n_in, d_w,
                          static_cast<const double*>(d_dz), d_dz_ptr, d_slot_ptr, d_contrib));
  KB_LAUNCH_CHECK();
  return KB_OK;
}

extern "C" int kb_train_residual_seg(int32_t n_cfg, const int64_t* d_cfg_ptr, int64_t atom0,
                                     const double* d_e_atom, const int64_t* d_seg_ptr, const int32_t* d_seg_src,
                                     int64_t slot_base, double* d_contrib, const double* d_e_ref,
                                     const double* d_f_ref, const double* d_w_e, const double* d_w_f,
                                     double inv_nglobal, int32_t use_energy, int32_t use_forces, double* d_a,
                                     double* d_gF, double* d_forces, double* d_loss_cfg, double* d_step,
                                     void* stream)
{
  if (n_cfg < 0) return KB_ERR_INVALID;
  if (n_cfg > 0 && (!d_cfg_ptr || !d_e_atom || !d_e_ref || !d_w_e || !d_a || !d_loss_cfg)) return KB_ERR_INVALID;
  if (n_cfg > 0 && use_forces && (!d_seg_ptr || !d_seg_src || !d_contrib || !d_f_ref || !d_w_f || !d_gF || !d_forces))
    return KB_ERR_INVALID;
  if (n_cfg == 0 && !d_step) return KB_OK;
  KB_CUDA(kb_launch_pdl(train_residual_seg_kernel, dim3(n_cfg > 0 ? n_cfg : 1), dim3(RS_WARPS * 32), 0,
                        kb_stream(stream), n_cfg, d_cfg_ptr, atom0, d_e_atom, d_seg_ptr, d_seg_src, slot_base, d_contrib,
                        d_contrib, d_e_ref, d_f_ref, d_w_e, d_w_f, inv_nglobal, use_energy, use_forces, d_a, d_gF,
                        d_forces, d_loss_cfg, d_step));
  KB_LAUNCH_CHECK();
  return KB_OK;
}

extern "C" int kb_train_backward(const kb_mlp_desc* M, int32_t n_atoms, const void* d_x, const void* d_theta,
                                 const double* d_scale, const void* d_dz, int32_t dz_dtype,
                                 const int64_t* d_dz_ptr, const int64_t* d_slot_ptr, const double* d_gslot,
                                 const double* d_a, double* d_G, double* d_partials, int32_t partial_rows,
                                 int32_t* h_rows_used, void* stream)
{
  int rc = fast_check(M);
  if (rc) return rc;
  if (n_atoms < 0 || !d_x || !d_theta || !d_a || !d_partials || partial_rows < 1 || !h_rows_used)
    return KB_ERR_INVALID;
  if (d_gslot && (!d_G || !d_dz || !d_dz_ptr || !d_slot_ptr)) return KB_ERR_INVALID;
  if (d_gslot && dz_dtype != KB_F32 && dz_dtype != KB_F64) return KB_ERR_INVALID;
  *h_rows_used = 0;
  if (n_atoms == 0) return KB_OK;
  cudaStream_t st = kb_stream(stream);
  if (d_gslot)
  {
    const int nbg = (n_atoms + TF_WARPS - 1) / TF_WARPS;
    if (dz_dtype == KB_F32)
      gather_kernel<float><<<nbg, TF_WARPS * 32, 0, st>>>(n_atoms, M->n_in, d_gslot, d_scale,
                                                         static_cast<const float*>(d_dz), d_dz_ptr, d_slot_ptr, d_G);
    else
      gather_kernel<double><<<nbg, TF_WARPS * 32, 0, st>>>(n_atoms, M->n_in, d_gslot, d_scale,
                                                          static_cast<const double*>(d_dz), d_dz_ptr, d_slot_ptr, d_G);
    KB_LAUNCH_CHECK();
  }
  const int nb = min(min(BW_BLOCKS * sm_count(), partial_rows), (n_atoms + TF_WARPS - 1) / TF_WARPS);
  *h_rows_used = nb;
  const double* G = d_gslot ? d_G : nullptr;
  if (M->dtype == KB_F32) rc = KB_DISPATCH_LT(launch_mlp_backward, float, M->n_hidden, *M, n_atoms, nb, d_x, d_theta, G, d_a, d_partials, st);
  else rc = KB_DISPATCH_LT(launch_mlp_backward, double, M->n_hidden, *M, n_atoms, nb, d_x, d_theta, G, d_a, d_partials, st);
  if (rc) return rc;
  KB_LAUNCH_CHECK();
  return KB_OK;
}

extern "C" int kb_train_update(const kb_mlp_desc* M, int32_t rows_used, const double* d_partials, int32_t n_cfg,
                               const double* d_loss_cfg, double* d_grad, int32_t apply_adam, void* d_theta,
                               void* d_m, void* d_v, const double* d_step, double lr, double beta1, double beta2,
                               double eps, double* d_epoch_loss, void* stream)
{
  int rc = fast_check(M);
  if (rc) return rc;
  if (rows_used < 0 || n_cfg < 0 || !d_grad || (rows_used > 0 && !d_partials) || (n_cfg > 0 && !d_loss_cfg))
    return KB_ERR_INVALID;
  if (apply_adam && (!d_theta || !d_m || !d_v || !d_step)) return KB_ERR_INVALID;
  const FNet N = make_fnet(*M);
  const int nb = (N.P + 1 + 31) / 32;
  cudaStream_t st = kb_stream(stream);
  if (M->dtype == KB_F32)
    train_update_kernel<float><<<nb, UP_WARPS * 32, 0, st>>>(
        N.P, rows_used, d_partials, n_cfg, d_loss_cfg, d_grad, apply_adam, static_cast<float*>(d_theta),
        static_cast<float*>(d_m), static_cast<float*>(d_v), d_step, lr, beta1, beta2, eps, d_epoch_loss);
  else
    train_update_kernel<double><<<nb, UP_WARPS * 32, 0, st>>>(
        N.P, rows_used, d_partials, n_cfg, d_loss_cfg, d_grad, apply_adam, static_cast<double*>(d_theta),
        static_cast<double*>(d_m), static_cast<double*>(d_v), d_step, lr, beta1, beta2, eps, d_epoch_loss);
  KB_LAUNCH_CHECK();
  return KB_OK;
}

extern "C" int kb_train_adam(int32_t n_params, const double* d_grad, void* d_theta, void* d_m, void* d_v,
                             const double* d_step, double lr, double beta1, double beta2, double eps,
                             int32_t dtype, double* d_epoch_loss, void* stream)
{
  if (n_params < 1 || !d_grad || !d_theta || !d_m || !d_v || !d_step) return KB_ERR_INVALID;
  if (dtype != KB_F32 && dtype != KB_F64) return KB_ERR_INVALID;
  cudaStream_t st = kb_stream(stream);
  const int nb = (n_params + 1 + 255) / 256;
  if (dtype == KB_F32)
    train_adam_kernel<float><<<nb, 256, 0, st>>>(n_params, d_grad, static_cast<float*>(d_theta),
                                                 static_cast<float*>(d_m), static_cast<float*>(d_v), d_step, lr,
                                                 beta1, beta2, eps, d_epoch_loss);
  else
    train_adam_kernel<double><<<nb, 256, 0, st>>>(n_params, d_grad, static_cast<double*>(d_theta),
                                                  static_cast<double*>(d_m), static_cast<double*>(d_v), d_step, lr,
                                                  beta1, beta2, eps, d_epoch_loss);
  KB_LAUNCH_CHECK();
  return KB_OK;
}
