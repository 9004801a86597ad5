// Fused training step for small tanh MLPs on top of the packed fingerprints: what
// CalculatorTorch.compute + Loss._get_loss_batch + loss.backward() + optimizer.step() do for one batch
// (kliff/legacy/calculators/calculator_torch.py:161-228, kliff/legacy/loss.py:843-920), as a handful
// of launches instead of ~150 torch micro-kernels:
//
//   kb_mlp_energy_grad     per atom: forward pass, e_atom, dE/dzeta                       (replaces model(zeta) +
//                                                                                        autograd.grad(create_graph))
//   kb_force_scatter_fwd   F = -sum dE/dzeta * dzeta/dr                                   (scatter.cu)
//   kb_train_residual      per configuration: E = sum e_atom, weighted residuals, loss,
//                          dL/de_atom and dL/dF                                           (loss.py:37-110, 879-920)
//   kb_force_scatter_bwd   G = dL/d(dE/dzeta)                                             (scatter.cu)
//   kb_mlp_param_grad      per atom: second-order backward through dE/dzeta and first-order
//                          backward through e_atom; parameter gradients as per-chunk partial
//                          sums, then one deterministic reduction into the flat gradient
//   (all-reduce of the flat gradient + loss: torch.distributed / NCCL, by the caller)
//   kb_adam_step           torch.optim.Adam's update on the flat parameter buffer
//
// Network: x -> [Linear(H_l) -> tanh] x L -> Linear(1); parameters flat in torch's order
// (W_1 [H_1][D], b_1, ..., W_out [1][H_L], b_out).  With a_0 = x, a_l = tanh(W_l a_{l-1} + b_l),
// d_l = 1 - a_l^2:
//   dE/dx = W_1^T delta_1,  delta_l = u_l * d_l,  u_L = w_out,  u_{l-1} = W_l^T delta_l
// and for a loss L with a = dL/de, G = dL/d(dE/dx):
//   P_1 = W_1 G;  R_l = P_l * d_l;  S_l = -2 a_l * P_l * u_l;  P_{l+1} = W_{l+1} R_l
//   dW_1 += delta_1 G^T;  dW_{l+1} += delta_{l+1} R_l^T;  dw_out += R_L
//   da_L = a w_out + S_L;  Q_l = da_l * d_l;  dW_l += Q_l a_{l-1}^T;  db_l += Q_l;  da_{l-1} = W_l^T Q_l + S_{l-1}
//   dw_out += a a_L;  db_out += a
#include "common.cuh"

namespace {

constexpr int MAXH = KB_MLP_MAX_WIDTH;    // hidden width limit
constexpr int MAXL = KB_MLP_MAX_HIDDEN;   // hidden layers limit
constexpr int MAXD = KB_MLP_MAX_IN;       // input width limit

struct Net
{
  int D, L, H[MAXL];
  int w_off[MAXL + 1], b_off[MAXL + 1];  // offsets into theta; index L = output layer
  int P;
};

__host__ __device__ inline Net make_net(const kb_mlp_desc& M)
{
  Net N;
  N.D = M.n_in;
  N.L = M.n_hidden;
  int pos = 0, fan_in = M.n_in;
  for (int l = 0; l < N.L; l++)
  {
    N.H[l] = M.width[l];
    N.w_off[l] = pos; pos += N.H[l] * fan_in;
    N.b_off[l] = pos; pos += N.H[l];
    fan_in = N.H[l];
  }
  N.w_off[N.L] = pos; pos += fan_in;
  N.b_off[N.L] = pos; pos += 1;
  N.P = pos;
  return N;
}

template <typename T> __device__ __forceinline__ T tanh_t(T x);
template <> __device__ __forceinline__ float tanh_t<float>(float x) { return tanhf(x); }
template <> __device__ __forceinline__ double tanh_t<double>(double x) { return tanh(x); }

// forward pass of one atom; act[l][h] = a_{l+1}[h]; returns e
template <typename T>
__device__ __forceinline__ T forward_atom(const Net& N, const T* __restrict__ th, const T* xr, T (&act)[MAXL][MAXH])
{
  const T* in = xr;
  int fan_in = N.D;
  for (int l = 0; l < N.L; l++)
  {
    const T* W = th + N.w_off[l];
    const T* b = th + N.b_off[l];
    const int H = N.H[l];
    for (int h = 0; h < H; h += 4)  // four units at a time: four independent chains per input load
    {
      const int h1 = min(h + 1, H - 1), h2 = min(h + 2, H - 1), h3 = min(h + 3, H - 1);
      T z0 = b[h], z1 = b[h1], z2 = b[h2], z3 = b[h3];
      const T *w0 = W + h * fan_in, *w1 = W + h1 * fan_in, *w2 = W + h2 * fan_in, *w3 = W + h3 * fan_in;
      for (int d = 0; d < fan_in; d++)
      {
        const T xd = in[d];
        z0 = fma(w0[d], xd, z0); z1 = fma(w1[d], xd, z1); z2 = fma(w2[d], xd, z2); z3 = fma(w3[d], xd, z3);
      }
      act[l][h] = tanh_t<T>(z0);
      if (h + 1 < H) act[l][h + 1] = tanh_t<T>(z1);
      if (h + 2 < H) act[l][h + 2] = tanh_t<T>(z2);
      if (h + 3 < H) act[l][h + 3] = tanh_t<T>(z3);
    }
    in = act[l];
    fan_in = N.H[l];
  }
  const T* wo = th + N.w_off[N.L];
  T e = th[N.b_off[N.L]];
  for (int j = 0; j < fan_in; j++) e = fma(wo[j], in[j], e);
  return e;
}

// delta_l (all layers) from the activations
template <typename T>
__device__ __forceinline__ void deltas_atom(const Net& N, const T* __restrict__ th, const T (&act)[MAXL][MAXH],
                                            T (&u)[MAXL][MAXH], T (&delta)[MAXL][MAXH])
{
  const int L = N.L;
  const T* wo = th + N.w_off[L];
  for (int j = 0; j < N.H[L - 1]; j++) u[L - 1][j] = wo[j];
  for (int l = L - 1; l >= 0; l--)
  {
    for (int j = 0; j < N.H[l]; j++) delta[l][j] = u[l][j] * (T(1) - act[l][j] * act[l][j]);
    if (l > 0)
    {
      const T* W = th + N.w_off[l];  // [H_l][H_{l-1}]
      for (int i = 0; i < N.H[l - 1]; i++)
      {
        T s = T(0);
        for (int j = 0; j < N.H[l]; j++) s = fma(W[j * N.H[l - 1] + i], delta[l][j], s);
        u[l - 1][i] = s;
      }
    }
  }
}

constexpr int MLP_BLOCK = 64;  // small blocks: a 6 400-atom batch still covers 100 SMs
template <typename T>
__global__ void __launch_bounds__(MLP_BLOCK)
mlp_energy_grad_kernel(const kb_mlp_desc M, int n, const T* __restrict__ x, const T* __restrict__ theta,
                       double* __restrict__ e_atom, double* __restrict__ dEdx)
{
  extern __shared__ __align__(16) unsigned char sm_raw[];
  T* th = reinterpret_cast<T*>(sm_raw);
  const Net N = make_net(M);
  for (int q = threadIdx.x; q < N.P; q += blockDim.x) th[q] = theta[q];
  __syncthreads();
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  T xr[MAXD];
  for (int d = 0; d < N.D; d++) xr[d] = x[(int64_t) t * N.D + d];
  T act[MAXL][MAXH], u[MAXL][MAXH], delta[MAXL][MAXH];
  const T e = forward_atom<T>(N, th, xr, act);
  e_atom[t] = (double) e;
  if (dEdx)
  {
    deltas_atom<T>(N, th, act, u, delta);
    const T* W1 = th + N.w_off[0];
    double* go = dEdx + (int64_t) t * N.D;
    for (int d = 0; d < N.D; d += 4)
    {
      const int d1 = min(d + 1, N.D - 1), d2 = min(d + 2, N.D - 1), d3 = min(d + 3, N.D - 1);
      T s0 = T(0), s1 = T(0), s2 = T(0), s3 = T(0);
      for (int h = 0; h < N.H[0]; h++)
      {
        const T dh = delta[0][h];
        const T* w = W1 + h * N.D;
        s0 = fma(w[d], dh, s0); s1 = fma(w[d1], dh, s1); s2 = fma(w[d2], dh, s2); s3 = fma(w[d3], dh, s3);
      }
      go[d] = (double) s0;
      if (d + 1 < N.D) go[d + 1] = (double) s1;
      if (d + 2 < N.D) go[d + 2] = (double) s2;
      if (d + 3 < N.D) go[d + 3] = (double) s3;
    }
  }
}

// One block per configuration: E = sum e_atom; residuals; loss; dL/de_atom, dL/dF.
__global__ void __launch_bounds__(128)
train_residual_kernel(int n_cfg, const int64_t* __restrict__ cfg_ptr, int64_t atom0,
                      const double* __restrict__ e_atom, const double* __restrict__ forces,
                      const double* __restrict__ e_ref, const double* __restrict__ f_ref,
                      const double* __restrict__ w_e, const double* __restrict__ w_f, double inv_nglobal,
                      int use_energy, int use_forces, double* __restrict__ a_out, double* __restrict__ gF,
                      double* __restrict__ loss)
{
  __shared__ double red[4];
  const int c = blockIdx.x;
  if (c >= n_cfg) return;
  const int64_t lo = cfg_ptr[c] - atom0, hi = cfg_ptr[c + 1] - atom0;  // atoms of this configuration in the batch
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double s = 0.0;
  for (int64_t q = lo + threadIdx.x; q < hi; q += blockDim.x) s += e_atom[q];
  s = warp_sum(s);
  if (lane == 0) red[warp] = s;
  __syncthreads();
  const double E = red[0] + red[1] + red[2] + red[3];
  __syncthreads();
  double part = 0.0;
  const double we = use_energy ? w_e[c] : 0.0;
  const double re = we * (E - e_ref[c]);                 // residual[0] (loss.py:100-109)
  const double a = 2.0 * we * re * inv_nglobal;
  for (int64_t q = lo + threadIdx.x; q < hi; q += blockDim.x) a_out[q] = a;
  if (threadIdx.x == 0) part = re * re;
  if (use_forces)
  {
    const double wf = w_f[c];
    for (int64_t q = 3 * lo + threadIdx.x; q < 3 * hi; q += blockDim.x)
    {
      const double rf = wf * (forces[q] - f_ref[q]);
      gF[q] = 2.0 * wf * rf * inv_nglobal;
      part += rf * rf;
    }
  }
  part = warp_sum(part);
  if (lane == 0) red[warp] = part;
  __syncthreads();
  if (threadIdx.x == 0) atomicAdd(loss, (red[0] + red[1] + red[2] + red[3]) * inv_nglobal);
}

// Per-atom vectors of the backward pass, S values per atom:
//   per layer l: Q_l [H_l], delta_l [H_l], R_l [H_l], a_l [H_l]   (R_l = 0 without forces)
__host__ __device__ inline int atom_vec_len(const Net& N)
{
  int s = 0;
  for (int l = 0; l < N.L; l++) s += 4 * N.H[l];
  return s;
}

template <typename T>
__global__ void __launch_bounds__(MLP_BLOCK)
mlp_backward_atoms_kernel(const kb_mlp_desc M, int n, const T* __restrict__ x, const T* __restrict__ theta,
                          const double* __restrict__ a_in, const double* __restrict__ G, T* __restrict__ V)
{
  extern __shared__ __align__(16) unsigned char sm_raw[];
  T* th = reinterpret_cast<T*>(sm_raw);
  const Net N = make_net(M);
  for (int q = threadIdx.x; q < N.P; q += blockDim.x) th[q] = theta[q];
  __syncthreads();
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  const int L = N.L;
  T xr[MAXD];
  for (int d = 0; d < N.D; d++) xr[d] = x[(int64_t) t * N.D + d];
  T act[MAXL][MAXH], u[MAXL][MAXH], delta[MAXL][MAXH], R[MAXL][MAXH], S[MAXL][MAXH], Q[MAXL][MAXH];
  forward_atom<T>(N, th, xr, act);
  deltas_atom<T>(N, th, act, u, delta);
  // second-order part: P_1 = W_1 G, then up the layers
  T Pv[MAXH];
  if (G)
  {
    const T* W1 = th + N.w_off[0];
    const double* gr = G + (int64_t) t * N.D;
    const int H = N.H[0];
    for (int h = 0; h < H; h += 4)
    {
      const int h1 = min(h + 1, H - 1), h2 = min(h + 2, H - 1), h3 = min(h + 3, H - 1);
      T s0 = T(0), s1 = T(0), s2 = T(0), s3 = T(0);
      const T *w0 = W1 + h * N.D, *w1 = W1 + h1 * N.D, *w2 = W1 + h2 * N.D, *w3 = W1 + h3 * N.D;
      for (int d = 0; d < N.D; d++)
      {
        const T gd = (T) gr[d];
        s0 = fma(w0[d], gd, s0); s1 = fma(w1[d], gd, s1); s2 = fma(w2[d], gd, s2); s3 = fma(w3[d], gd, s3);
      }
      Pv[h] = s0;
      if (h + 1 < H) Pv[h + 1] = s1;
      if (h + 2 < H) Pv[h + 2] = s2;
      if (h + 3 < H) Pv[h + 3] = s3;
    }
  }
  for (int l = 0; l < L; l++)
  {
    T Pn[MAXH];
    for (int j = 0; j < N.H[l]; j++)
    {
      const T dl = T(1) - act[l][j] * act[l][j];
      const T p = G ? Pv[j] : T(0);
      R[l][j] = p * dl;
      S[l][j] = T(-2) * act[l][j] * p * u[l][j];
    }
    if (G && l + 1 < L)
    {
      const T* W = th + N.w_off[l + 1];  // [H_{l+1}][H_l]
      for (int i = 0; i < N.H[l + 1]; i++)
      {
        T s = T(0);
        for (int j = 0; j < N.H[l]; j++) s = fma(W[i * N.H[l] + j], R[l][j], s);
        Pn[i] = s;
      }
      for (int i = 0; i < N.H[l + 1]; i++) Pv[i] = Pn[i];
    }
  }
  // first-order backward with the extra sources S_l
  const T a = (T) a_in[t];
  T da[MAXH];
  const T* wo = th + N.w_off[L];
  for (int j = 0; j < N.H[L - 1]; j++) da[j] = a * wo[j] + S[L - 1][j];
  for (int l = L - 1; l >= 0; l--)
  {
    for (int j = 0; j < N.H[l]; j++) Q[l][j] = da[j] * (T(1) - act[l][j] * act[l][j]);
    if (l > 0)
    {
      const T* W = th + N.w_off[l];  // [H_l][H_{l-1}]
      T dn[MAXH];
      for (int i = 0; i < N.H[l - 1]; i++)
      {
        T s = S[l - 1][i];
        for (int j = 0; j < N.H[l]; j++) s = fma(W[j * N.H[l - 1] + i], Q[l][j], s);
        dn[i] = s;
      }
      for (int i = 0; i < N.H[l - 1]; i++) da[i] = dn[i];
    }
  }
  T* v = V + (int64_t) t * atom_vec_len(N);
  for (int l = 0; l < L; l++)
  {
    for (int j = 0; j < N.H[l]; j++)
    {
      v[j] = Q[l][j]; v[N.H[l] + j] = delta[l][j]; v[2 * N.H[l] + j] = R[l][j]; v[3 * N.H[l] + j] = act[l][j];
    }
    v += 4 * N.H[l];
  }
}

// Partial parameter gradients of a chunk of atoms: the chunk's per-atom vectors, inputs and G are staged
// in shared memory (contiguous, coalesced copies), then thread = parameter sums over the chunk.
constexpr int PG_CHUNK = 32;
template <typename T>
__global__ void __launch_bounds__(256)
mlp_param_partials_kernel(const kb_mlp_desc M, int n, const T* __restrict__ x, const double* __restrict__ a_in,
                          const double* __restrict__ G, const T* __restrict__ V, double* __restrict__ partials)
{
  extern __shared__ __align__(16) unsigned char sm_raw[];
  const Net N = make_net(M);
  const int S = atom_vec_len(N), D = N.D;
  const int lo = blockIdx.x * PG_CHUNK, cnt = min(n, lo + PG_CHUNK) - lo;
  T* sV = reinterpret_cast<T*>(sm_raw);        // [cnt][S]
  T* sX = sV + PG_CHUNK * S;                   // [cnt][D]
  T* sG = sX + PG_CHUNK * D;                   // [cnt][D]
  T* sA = sG + PG_CHUNK * D;                   // [cnt]
  T* sC = sA + PG_CHUNK;                       // {1, 0}
  if (threadIdx.x == 0) { sC[0] = T(1); sC[1] = T(0); }
  for (int q = threadIdx.x; q < cnt * S; q += blockDim.x) sV[q] = V[(int64_t) lo * S + q];
  for (int q = threadIdx.x; q < cnt * D; q += blockDim.x)
  {
    sX[q] = x[(int64_t) lo * D + q];
    sG[q] = G ? (T) G[(int64_t) lo * D + q] : T(0);
  }
  for (int q = threadIdx.x; q < cnt; q += blockDim.x) sA[q] = (T) a_in[lo + q];
  __syncthreads();
  for (int p = threadIdx.x; p < N.P; p += blockDim.x)
  {
    // which parameter: layer (0..L = output), weight (i, j) or bias i
    int layer = 0;
    while (layer < N.L && p >= N.w_off[layer + 1]) layer++;
    const bool is_bias = p >= N.b_off[layer];
    const int fan_in = layer == 0 ? D : N.H[layer - 1];
    const int q = is_bias ? p - N.b_off[layer] : p - N.w_off[layer];
    const int i = is_bias ? q : q / fan_in, j = is_bias ? 0 : q - i * fan_in;
    int voff = 0;                                    // start of layer `layer`'s vectors inside an atom's row
    for (int l = 0; l < layer && l < N.L; l++) voff += 4 * N.H[l];
    const int vprev = layer > 0 ? voff - 4 * N.H[layer - 1] : 0;  // previous hidden layer's vectors
    // every case is sum_t A[t] * B[t] + C[t] * E[t] over rows of the staged arrays
    const T *pa, *pb, *pc, *pe;
    int sa, sb, sc, se;
    const T *one = sC, *zero = sC + 1;
    if (layer < N.L)
    {
      const int H = N.H[layer];
      pa = sV + voff + i; sa = S;                                        // Q_l[i]
      if (is_bias) { pb = one; sb = 0; pc = zero; sc = 0; pe = zero; se = 0; }
      else if (layer == 0)
      {
        pb = sX + j; sb = D;                                             // x[j]
        pc = sV + voff + H + i; sc = S; pe = sG + j; se = D;             // delta_1[i] G[j]
      }
      else
      {
        const int Hp = N.H[layer - 1];
        pb = sV + vprev + 3 * Hp + j; sb = S;                            // a_{l-1}[j]
        pc = sV + voff + H + i; sc = S; pe = sV + vprev + 2 * Hp + j; se = S;  // delta_l[i] R_{l-1}[j]
      }
    }
    else
    {
      pa = sA; sa = 1;                                                   // a
      if (is_bias) { pb = one; sb = 0; pc = zero; sc = 0; pe = zero; se = 0; }
      else
      {
        const int Hp = N.H[N.L - 1];
        pb = sV + vprev + 3 * Hp + j; sb = S;                            // a_L[j]
        pc = one; sc = 0; pe = sV + vprev + 2 * Hp + j; se = S;          // + R_L[j]
      }
    }
    double acc = 0.0;
    for (int t = 0; t < cnt; t++)
      acc += (double) fma(pa[t * sa], pb[t * sb], pc[t * sc] * pe[t * se]);
    partials[(int64_t) blockIdx.x * N.P + p] = acc;
  }
}

// grad[p] = sum over chunks, one warp per parameter, fixed order (deterministic)
__global__ void __launch_bounds__(128)
reduce_partials_kernel(int n_chunks, int P, const double* __restrict__ partials, double* __restrict__ grad)
{
  const int p = blockIdx.x * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (p >= P) return;
  double s = 0.0;
  for (int c = lane; c < n_chunks; c += 32) s += partials[(int64_t) c * P + p];
  s = warp_sum(s);
  if (lane == 0) grad[p] = s;
}

// torch.optim.Adam (no amsgrad, no weight decay), torch/optim/adam.py `_single_tensor_adam`:
//   m = b1 m + (1 - b1) g;  v = b2 v + (1 - b2) g^2
//   p -= lr / (1 - b1^t) * m / (sqrt(v) / sqrt(1 - b2^t) + eps)
template <typename T>
__global__ void adam_kernel(int P, T* __restrict__ theta, const double* __restrict__ grad, double grad_scale,
                            T* __restrict__ m, T* __restrict__ v, double* __restrict__ step, double lr,
                            double b1, double b2, double eps)
{
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  const double t = step[0] + 1.0;
  if (p < P)
  {
    const T g = (T) (grad[p] * grad_scale);
    const T mm = m[p] + (T) (1.0 - b1) * (g - m[p]);           // exp_avg.lerp_(grad, 1 - beta1)
    const T vv = (T) b2 * v[p] + (T) (1.0 - b2) * g * g;       // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, 1 - beta2)
    m[p] = mm;
    v[p] = vv;
    const double bc1 = 1.0 - pow(b1, t), bc2 = 1.0 - pow(b2, t);
    const T denom = (T) sqrt((double) vv) / (T) sqrt(bc2) + (T) eps;   // sqrt of a T value, rounded back to T
    theta[p] = theta[p] + (T) (-lr / bc1) * (mm / denom);      // param.addcdiv_(exp_avg, denom, value=-step_size)
  }
  __syncthreads();
  if (gridDim.x == 1 && threadIdx.x == 0) step[0] = t;
}
__global__ void adam_bump_kernel(double* step) { step[0] += 1.0; }

int check_desc(const kb_mlp_desc* M)
{
  if (!M) return KB_ERR_INVALID;
  if (M->n_in < 1 || M->n_in > MAXD) return KB_ERR_LIMIT;
  if (M->n_hidden < 1 || M->n_hidden > MAXL) return KB_ERR_LIMIT;
  for (int l = 0; l < M->n_hidden; l++)
    if (M->width[l] < 1 || M->width[l] > MAXH) return KB_ERR_LIMIT;
  if (M->dtype != KB_F32 && M->dtype != KB_F64) return KB_ERR_INVALID;
  return KB_OK;
}

}  // namespace

extern "C" int kb_mlp_num_params(const kb_mlp_desc* M, int32_t* n_params, int32_t* atom_vec)
{
  int rc = check_desc(M);
  if (rc) return rc;
  const Net N = make_net(*M);
  if (n_params) *n_params = N.P;
  if (atom_vec) *atom_vec = atom_vec_len(N);
  return KB_OK;
}

extern "C" int kb_mlp_energy_grad(const kb_mlp_desc* M, int32_t n_atoms, const void* d_x, const void* d_theta,
                                  double* d_e_atom, double* d_dEdx, void* stream)
{
  int rc = check_desc(M);
  if (rc) return rc;
  if (n_atoms < 0 || !d_x || !d_theta || !d_e_atom) return KB_ERR_INVALID;
  if (n_atoms == 0) return KB_OK;
  const Net N = make_net(*M);
  const int nb = (n_atoms + MLP_BLOCK - 1) / MLP_BLOCK;
  if (M->dtype == KB_F32)
    mlp_energy_grad_kernel<float><<<nb, MLP_BLOCK, sizeof(float) * N.P, kb_stream(stream)>>>(
        *M, n_atoms, static_cast<const float*>(d_x), static_cast<const float*>(d_theta), d_e_atom, d_dEdx);
  else
    mlp_energy_grad_kernel<double><<<nb, MLP_BLOCK, sizeof(double) * N.P, kb_stream(stream)>>>(
        *M, n_atoms, static_cast<const double*>(d_x), static_cast<const double*>(d_theta), d_e_atom, d_dEdx);
  KB_LAUNCH_CHECK();
  return KB_OK;
}

extern "C" int kb_train_residual(int32_t n_cfg, const int64_t* d_cfg_ptr, int64_t atom0, const double* d_e_atom,
                                 const double* d_forces, const double* d_e_ref, const double* d_f_ref,
                                 const double* d_w_e, const double* d_w_f, double inv_nglobal,
                                 int32_t use_energy, int32_t use_forces, double* d_a, double* d_gF,
                                 double* d_loss, void* stream)
{
  if (n_cfg < 0 || !d_cfg_ptr || !d_e_atom || !d_e_ref || !d_w_e || !d_a || !d_loss) return KB_ERR_INVALID;
  if (use_forces && (!d_forces || !d_f_ref || !d_w_f || !d_gF)) return KB_ERR_INVALID;
  if (n_cfg == 0) return KB_OK;
  train_residual_kernel<<<n_cfg, 128, 0, kb_stream(stream)>>>(n_cfg, d_cfg_ptr, atom0, d_e_atom, d_forces, d_e_ref,
                                                              d_f_ref, d_w_e, d_w_f, inv_nglobal, use_energy,
                                                              use_forces, d_a, d_gF, d_loss);
  KB_LAUNCH_CHECK();
  return KB_OK;
}

extern "C" int kb_mlp_param_grad(const kb_mlp_desc* M, int32_t n_atoms, const void* d_x, const void* d_theta,
                                 const double* d_a, const double* d_G, void* d_atom_vecs, double* d_partials,
                                 double* d_grad, void* stream)
{
  int rc = check_desc(M);
  if (rc) return rc;
  if (n_atoms < 0 || !d_x || !d_theta || !d_a || !d_atom_vecs || !d_partials || !d_grad) return KB_ERR_INVALID;
  const Net N = make_net(*M);
  cudaStream_t st = kb_stream(stream);
  if (n_atoms == 0)
  {
    KB_CUDA(cudaMemsetAsync(d_grad, 0, sizeof(double) * N.P, st));
    return KB_OK;
  }
  const int nb = (n_atoms + MLP_BLOCK - 1) / MLP_BLOCK;
  const int nc = (n_atoms + PG_CHUNK - 1) / PG_CHUNK;
  const size_t pg_smem = (size_t) PG_CHUNK * (atom_vec_len(N) + 2 * N.D + 1) + 2;  // elements
  if (M->dtype == KB_F32)
  {
    mlp_backward_atoms_kernel<float><<<nb, MLP_BLOCK, sizeof(float) * N.P, st>>>(
        *M, n_atoms, static_cast<const float*>(d_x), static_cast<const float*>(d_theta), d_a, d_G,
        static_cast<float*>(d_atom_vecs));
    mlp_param_partials_kernel<float><<<nc, 256, pg_smem * sizeof(float), st>>>(
        *M, n_atoms, static_cast<const float*>(d_x), d_a, d_G, static_cast<const float*>(d_atom_vecs), d_partials);
  }
  else
  {
    mlp_backward_atoms_kernel<double><<<nb, MLP_BLOCK, sizeof(double) * N.P, st>>>(
        *M, n_atoms, static_cast<const double*>(d_x), static_cast<const double*>(d_theta), d_a, d_G,
        static_cast<double*>(d_atom_vecs));
    cudaFuncSetAttribute(mlp_param_partials_kernel<double>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                         (int) (pg_smem * sizeof(double)));
    mlp_param_partials_kernel<double><<<nc, 256, pg_smem * sizeof(double), st>>>(
        *M, n_atoms, static_cast<const double*>(d_x), d_a, d_G, static_cast<const double*>(d_atom_vecs), d_partials);
  }
  reduce_partials_kernel<<<(N.P + 3) / 4, 128, 0, st>>>(nc, N.P, d_partials, d_grad);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return -(int) e;
  kb_note_launch(3);
  return KB_OK;
}

extern "C" int kb_adam_step(int32_t n_params, void* d_theta, const double* d_grad, double grad_scale, void* d_m,
                            void* d_v, double* d_step, double lr, double beta1, double beta2, double eps,
                            int32_t dtype, void* stream)
{
  if (n_params < 1 || !d_theta || !d_grad || !d_m || !d_v || !d_step) return KB_ERR_INVALID;
  if (dtype != KB_F32 && dtype != KB_F64) return KB_ERR_INVALID;
  cudaStream_t st = kb_stream(stream);
  const int nb = (n_params + 255) / 256;
  if (dtype == KB_F32)
    adam_kernel<float><<<nb, 256, 0, st>>>(n_params, static_cast<float*>(d_theta), d_grad, grad_scale,
                                           static_cast<float*>(d_m), static_cast<float*>(d_v), d_step, lr, beta1,
                                           beta2, eps);
  else
    adam_kernel<double><<<nb, 256, 0, st>>>(n_params, static_cast<double*>(d_theta), d_grad, grad_scale,
                                            static_cast<double*>(d_m), static_cast<double*>(d_v), d_step, lr, beta1,
                                            beta2, eps);
  if (nb > 1) adam_bump_kernel<<<1, 1, 0, st>>>(d_step);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return -(int) e;
  kb_note_launch(nb > 1 ? 2 : 1);
  return KB_OK;
}
