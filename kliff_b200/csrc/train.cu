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
    for (int h = 0; h < N.H[l]; h++)
    {
      T z = b[h];
      for (int d = 0; d < fan_in; d++) z = fma(W[h * fan_in + d], in[d], z);
      act[l][h] = tanh_t<T>(z);
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

template <typename T>
__global__ void __launch_bounds__(128)
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
    for (int d = 0; d < N.D; d++)
    {
      T s = T(0);
      for (int h = 0; h < N.H[0]; h++) s = fma(W1[h * N.D + d], delta[0][h], s);
      dEdx[(int64_t) t * N.D + d] = (double) s;
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
__global__ void __launch_bounds__(128)
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
    for (int h = 0; h < N.H[0]; h++)
    {
      T s = T(0);
      for (int d = 0; d < N.D; d++) s = fma(W1[h * N.D + d], (T) G[(int64_t) t * N.D + d], s);
      Pv[h] = s;
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

// Partial parameter gradients of a chunk of atoms: thread = parameter.
constexpr int PG_CHUNK = 64;
template <typename T>
__global__ void __launch_bounds__(256)
mlp_param_partials_kernel(const kb_mlp_desc M, int n, const T* __restrict__ x, const double* __restrict__ a_in,
                          const double* __restrict__ G, const T* __restrict__ V, double* __restrict__ partials)
{
  const Net N = make_net(M);
  const int S = atom_vec_len(N);
  const int lo = blockIdx.x * PG_CHUNK, hi = min(n, lo + PG_CHUNK);
  for (int p = threadIdx.x; p < N.P; p += blockDim.x)
  {
    // which parameter: layer (0..L = output), weight (i, j) or bias i
    int layer = 0;
    while (layer < N.L && p >= N.w_off[layer + 1]) layer++;
    const bool is_bias = p >= N.b_off[layer];
    const int fan_in = layer == 0 ? N.D : N.H[layer - 1];
    const int q = is_bias ? p - N.b_off[layer] : p - N.w_off[layer];
    const int i = is_bias ? q : q / fan_in, j = is_bias ? 0 : q - i * fan_in;
    int voff = 0;                                    // start of layer `layer`'s vectors inside an atom's V row
    for (int l = 0; l < layer && l < N.L; l++) voff += 4 * N.H[l];
    const int vprev = layer > 0 ? voff - 4 * N.H[layer - 1] : 0;  // previous hidden layer's vectors
    double acc = 0.0;
    for (int t = lo; t < hi; t++)
    {
      const T* v = V + (int64_t) t * S;
      if (layer < N.L)
      {
        const int H = N.H[layer];
        const T Qi = v[voff + i];
        if (is_bias) acc += (double) Qi;
        else if (layer == 0)
        {
          T c = Qi * x[(int64_t) t * N.D + j];
          if (G) c = fma(v[voff + H + i], (T) G[(int64_t) t * N.D + j], c);   // delta_1 G^T
          acc += (double) c;
        }
        else
        {
          const int Hp = N.H[layer - 1];
          T c = Qi * v[vprev + 3 * Hp + j];                                    // Q_l a_{l-1}^T
          c = fma(v[voff + H + i], v[vprev + 2 * Hp + j], c);                  // delta_l R_{l-1}^T
          acc += (double) c;
        }
      }
      else
      {
        const T a = (T) a_in[t];
        if (is_bias) acc += (double) a;
        else
        {
          const int Hp = N.H[N.L - 1];
          acc += (double) fma(a, v[vprev + 3 * Hp + j], v[vprev + 2 * Hp + j]);  // a a_L + R_L
        }
      }
    }
    partials[(int64_t) blockIdx.x * N.P + p] = acc;
  }
}

__global__ void reduce_partials_kernel(int n_chunks, int P, const double* __restrict__ partials,
                                       double* __restrict__ grad)
{
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= P) return;
  double s = 0.0;
  for (int c = 0; c < n_chunks; c++) s += partials[(int64_t) c * P + p];
  grad[p] = s;
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
  const int nb = (n_atoms + 127) / 128;
  if (M->dtype == KB_F32)
    mlp_energy_grad_kernel<float><<<nb, 128, sizeof(float) * N.P, kb_stream(stream)>>>(
        *M, n_atoms, static_cast<const float*>(d_x), static_cast<const float*>(d_theta), d_e_atom, d_dEdx);
  else
    mlp_energy_grad_kernel<double><<<nb, 128, sizeof(double) * N.P, kb_stream(stream)>>>(
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
  const int nb = (n_atoms + 127) / 128;
  const int nc = (n_atoms + PG_CHUNK - 1) / PG_CHUNK;
  if (M->dtype == KB_F32)
  {
    mlp_backward_atoms_kernel<float><<<nb, 128, sizeof(float) * N.P, st>>>(
        *M, n_atoms, static_cast<const float*>(d_x), static_cast<const float*>(d_theta), d_a, d_G,
        static_cast<float*>(d_atom_vecs));
    mlp_param_partials_kernel<float><<<nc, 256, 0, st>>>(*M, n_atoms, static_cast<const float*>(d_x), d_a, d_G,
                                                         static_cast<const float*>(d_atom_vecs), d_partials);
  }
  else
  {
    mlp_backward_atoms_kernel<double><<<nb, 128, sizeof(double) * N.P, st>>>(
        *M, n_atoms, static_cast<const double*>(d_x), static_cast<const double*>(d_theta), d_a, d_G,
        static_cast<double*>(d_atom_vecs));
    mlp_param_partials_kernel<double><<<nc, 256, 0, st>>>(*M, n_atoms, static_cast<const double*>(d_x), d_a, d_G,
                                                          static_cast<const double*>(d_atom_vecs), d_partials);
  }
  reduce_partials_kernel<<<(N.P + 127) / 128, 128, 0, st>>>(nc, N.P, d_partials, d_grad);
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
