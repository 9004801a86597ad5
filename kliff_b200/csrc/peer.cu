// Gradient all-reduce over NVLink peer memory, fused with the partial reduction and Adam
// (the multi-GPU form of kb_train_update; loss.py:843-920 + torch.optim.Adam per batch, one process per GPU).
//
// With NCCL a data-parallel step is  reduce partials -> ncclAllReduce(642 doubles) -> Adam : three launches
// whose middle one costs ~20 us of latency for 5 KB — as much as all the arithmetic of a 100-configuration
// step on 8 GPUs.  Here every rank owns a small workspace in device memory that its peers map through CUDA
// IPC; one kernel per step
//   1. sums its own gradient partials (fixed order),
//   2. stores the 32 values of each block straight into every peer's inbox (one 256-byte P2P store per
//      peer), fences (system scope) and raises a sequence flag in every peer,
//   3. waits for the W flags of its own inbox, adds the W rows in rank order — every rank forms the same sum
//      bit for bit — and applies Adam.
// Inbox and flags are double-buffered on the parity of the step number: a rank can only be one exchange
// ahead of the slowest peer (it needs that peer's row of step s + 1, which the peer sends after it has read
// step s), so two buffers are enough.  Flags carry the step number itself (monotonic), nothing is ever reset.
// A rank that waits longer than KB_PEER_TIMEOUT_NS gives up, records it in the workspace's error word and
// carries on without waiting again in later steps (the host checks the word after the pass), so a lost
// peer costs one timeout, not one per step of the epoch graph, and cannot hang the GPU.
#include "common.cuh"

namespace {

constexpr int PR_WARPS = 32;
constexpr int PR_MAX_WORLD = 16;
constexpr unsigned long long KB_PEER_TIMEOUT_NS = 4000000000ull;

// workspace layout (bytes), PP = padded vector length (multiple of 32), NB = PP / 32 blocks:
//   [0, 64)                          error word (int) + padding
//   flags  [2][W][NB] u64            sequence numbers
//   inbox  [2][W][PP] f64
__host__ __device__ inline size_t ws_flags_off() { return 64; }
__host__ __device__ inline size_t ws_inbox_off(int W, int NB) { return 64 + (size_t) 2 * W * NB * 8; }
__host__ __device__ inline size_t ws_bytes(int W, int PP) { return ws_inbox_off(W, PP / 32) + (size_t) 2 * W * PP * 8; }

struct PeerPtrs
{
  unsigned char* p[PR_MAX_WORLD];
};

template <typename T>
__device__ __forceinline__ void adam_apply(int p, double gd, T* theta, T* m, T* v, double t, double lr, double b1,
                                           double b2, double eps)
{
  // torch.optim.Adam `_single_tensor_adam` (no amsgrad, no weight decay), as in train_fast.cu
  const T g = (T) gd;
  const T mm = m[p] + (T) (1.0 - b1) * (g - m[p]);
  const T vv = (T) b2 * v[p] + (T) (1.0 - b2) * g * g;
  m[p] = mm;
  v[p] = vv;
  const double bc1 = 1.0 - pow(b1, t), bc2 = 1.0 - pow(b2, t);
  const T denom = (T) sqrt((double) vv) / (T) sqrt(bc2) + (T) eps;
  theta[p] = theta[p] + (T) (-lr / bc1) * (mm / denom);
}

template <typename T>
__global__ void __launch_bounds__(PR_WARPS * 32)
train_update_peer_kernel(int P, int rows, const double* __restrict__ partials, int n_cfg,
                         const double* __restrict__ loss_cfg, double* __restrict__ grad, int adam, T* theta, T* m,
                         T* v, const double* __restrict__ step, double lr, double b1, double b2, double eps,
                         double* __restrict__ epoch_loss, const PeerPtrs peers, int W, int me, int PP)
{
  __shared__ double sub[PR_WARPS][32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int p = blockIdx.x * 32 + lane;
  const int NB = PP >> 5;
  double s = 0.0;
  if (p < P)
  {
    int r = warp;
    for (; r + 3 * PR_WARPS < rows; r += 4 * PR_WARPS)
    {
      const double v0 = partials[(int64_t) r * P + p], v1 = partials[(int64_t) (r + PR_WARPS) * P + p];
      const double v2 = partials[(int64_t) (r + 2 * PR_WARPS) * P + p], v3 = partials[(int64_t) (r + 3 * PR_WARPS) * P + p];
      s += (v0 + v1) + (v2 + v3);
    }
    for (; r < rows; r += PR_WARPS) s += partials[(int64_t) r * P + p];
  }
  else if (p == P)
    for (int c = warp; c < n_cfg; c += PR_WARPS) s += loss_cfg[c];
  sub[warp][lane] = s;
  __syncthreads();
  if (warp != 0) return;
  s = 0.0;
#pragma unroll
  for (int w = 0; w < PR_WARPS; w++) s += sub[w][lane];

  // ---- exchange: this block's 32 values to every rank's inbox, then the flag
  const unsigned long long seq = (unsigned long long) step[0];  // counted by the residual launch of this step
  const int par = (int) (seq & 1ull);
  const size_t f_off = ws_flags_off() + ((size_t) (par * W + me) * NB + blockIdx.x) * 8;
  const size_t d_off = ws_inbox_off(W, NB) + ((size_t) (par * W + me) * PP + p) * 8;
  for (int r = 0; r < W; r++)
    *reinterpret_cast<volatile double*>(peers.p[r] + d_off) = s;
  __threadfence_system();
  __syncwarp();
  if (lane < W)
  {
    unsigned long long* f = reinterpret_cast<unsigned long long*>(peers.p[lane] + f_off);
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(f), "l"(seq) : "memory");
  }
  // ---- wait for the W rows of my own inbox
  unsigned char* mine = peers.p[me];
  // once a wait has timed out the run is lost (the host raises after the pass): later steps of the same
  // graph must not each wait another 4 s
  const bool dead = *reinterpret_cast<const volatile int*>(mine) != 0;
  bool ok = !dead;
  if (lane < W && !dead)
  {
    const unsigned long long* f = reinterpret_cast<const unsigned long long*>(
        mine + ws_flags_off() + ((size_t) (par * W + lane) * NB + blockIdx.x) * 8);
    unsigned long long t0 = 0, seen = 0;
    for (int spin = 0;; spin++)
    {
      asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(seen) : "l"(f) : "memory");
      if (seen >= seq) break;
      if ((spin & 1023) == 1023)
      {
        unsigned long long now;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
        if (t0 == 0) t0 = now;
        else if (now - t0 > KB_PEER_TIMEOUT_NS) { ok = false; break; }
      }
    }
  }
  ok = __all_sync(0xffffffffu, ok);
  if (!ok && lane == 0) *reinterpret_cast<volatile int*>(mine) = 1;
  __threadfence_system();
  double tot = 0.0;
  if (p <= P)
    for (int r = 0; r < W; r++)
      tot += *reinterpret_cast<const volatile double*>(mine + ws_inbox_off(W, NB) + ((size_t) (par * W + r) * PP + p) * 8);
  if (p > P) return;
  grad[p] = tot;
  if (p == P) { if (epoch_loss) epoch_loss[0] += tot; }
  else if (adam) adam_apply<T>(p, tot, theta, m, v, step[0], lr, b1, b2, eps);
}

}  // namespace

extern "C" int kb_peer_ws_bytes(int32_t world, int32_t n_values, int64_t* bytes, int32_t* padded)
{
  if (world < 1 || world > PR_MAX_WORLD || n_values < 1 || !bytes) return KB_ERR_INVALID;
  const int PP = (n_values + 31) & ~31;
  *bytes = (int64_t) ws_bytes(world, PP);
  if (padded) *padded = PP;
  return KB_OK;
}

// cudaMalloc'ed (IPC-exportable), zeroed workspace + its 64-byte IPC handle
extern "C" int kb_peer_alloc(int64_t bytes, void** d_ptr, void* h_handle64)
{
  if (bytes < 64 || !d_ptr || !h_handle64) return KB_ERR_INVALID;
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  void* p = nullptr;
  KB_CUDA(cudaMalloc(&p, (size_t) bytes));
  cudaError_t e = cudaMemset(p, 0, (size_t) bytes);
  if (e == cudaSuccess) e = cudaDeviceSynchronize();
  cudaIpcMemHandle_t h;
  if (e == cudaSuccess) e = cudaIpcGetMemHandle(&h, p);
  if (e != cudaSuccess) { cudaFree(p); return -(int) e; }
  memcpy(h_handle64, &h, 64);
  *d_ptr = p;
  return KB_OK;
}

extern "C" int kb_peer_open(const void* h_handle64, void** d_ptr)
{
  if (!h_handle64 || !d_ptr) return KB_ERR_INVALID;
  cudaIpcMemHandle_t h;
  memcpy(&h, h_handle64, 64);
  KB_CUDA(cudaIpcOpenMemHandle(d_ptr, h, cudaIpcMemLazyEnablePeerAccess));
  return KB_OK;
}

extern "C" int kb_peer_close(void* d_ptr)
{
  if (!d_ptr) return KB_OK;
  KB_CUDA(cudaIpcCloseMemHandle(d_ptr));
  return KB_OK;
}

extern "C" int kb_peer_free(void* d_ptr)
{
  if (!d_ptr) return KB_OK;
  KB_CUDA(cudaFree(d_ptr));
  return KB_OK;
}

// host read of the workspace's error word (1: a peer did not arrive within the timeout)
extern "C" int kb_peer_error(const void* d_ws, int32_t* h_err, void* stream)
{
  if (!d_ws || !h_err) return KB_ERR_INVALID;
  KB_CUDA(cudaMemcpyAsync(h_err, d_ws, sizeof(int32_t), cudaMemcpyDeviceToHost, kb_stream(stream)));
  KB_CUDA(cudaStreamSynchronize(kb_stream(stream)));
  return KB_OK;
}

extern "C" int kb_train_update_peer(int32_t n_params, int32_t rows_used, const double* d_partials, int32_t n_cfg,
                                    const double* d_loss_cfg, double* d_grad, int32_t apply_adam, void* d_theta,
                                    void* d_m, void* d_v, const double* d_step, double lr, double beta1,
                                    double beta2, double eps, int32_t dtype, double* d_epoch_loss,
                                    void* const* h_peer_ws, int32_t world, int32_t rank, void* stream)
{
  if (n_params < 1 || rows_used < 0 || n_cfg < 0 || !d_grad || !d_step || !h_peer_ws) return KB_ERR_INVALID;
  if ((rows_used > 0 && !d_partials) || (n_cfg > 0 && !d_loss_cfg)) return KB_ERR_INVALID;
  if (world < 2 || world > PR_MAX_WORLD || rank < 0 || rank >= world) return KB_ERR_INVALID;
  if (apply_adam && (!d_theta || !d_m || !d_v)) return KB_ERR_INVALID;
  if (dtype != KB_F32 && dtype != KB_F64) return KB_ERR_INVALID;
  PeerPtrs peers;
  for (int r = 0; r < PR_MAX_WORLD; r++) peers.p[r] = r < world ? static_cast<unsigned char*>(h_peer_ws[r]) : nullptr;
  for (int r = 0; r < world; r++)
    if (!peers.p[r]) return KB_ERR_INVALID;
  const int PP = (n_params + 1 + 31) & ~31;
  const int nb = PP / 32;
  cudaStream_t st = kb_stream(stream);
  if (dtype == KB_F32)
    train_update_peer_kernel<float><<<nb, PR_WARPS * 32, 0, st>>>(
        n_params, rows_used, d_partials, n_cfg, d_loss_cfg, d_grad, apply_adam, static_cast<float*>(d_theta),
        static_cast<float*>(d_m), static_cast<float*>(d_v), d_step, lr, beta1, beta2, eps, d_epoch_loss, peers, world,
        rank, PP);
  else
    train_update_peer_kernel<double><<<nb, PR_WARPS * 32, 0, st>>>(
        n_params, rows_used, d_partials, n_cfg, d_loss_cfg, d_grad, apply_adam, static_cast<double*>(d_theta),
        static_cast<double*>(d_m), static_cast<double*>(d_v), d_step, lr, beta1, beta2, eps, d_epoch_loss, peers, world,
        rank, PP);
  KB_LAUNCH_CHECK();
  return KB_OK;
}
