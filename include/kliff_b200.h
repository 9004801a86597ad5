/*
 * kliff_b200.h — C-ABI of libkliff_b200.so: the B200-native (sm_100a) replacement for the two
 * native extensions that KLIFF's descriptor/NN hot path sits on:
 *
 *   kliff.neighbor.neighlist                 (kliff/neighbor/neighbor_list_bind.cpp:31-253 over
 *                                             neighbor_list.h:24-53: nbl_create_paddings,
 *                                             nbl_build, nbl_get_neigh)
 *   kliff.legacy.descriptors.symmetry_function.sf
 *                                            (sym_fn_bind.cpp:8-95 over sym_fn.hpp: Descriptor::
 *                                             set_cutoff, add_descriptor, generate_one_atom)
 * plus the force contraction of kliff/legacy/calculators/calculator_torch.py:266-269 and the
 * dense fold of kliff/legacy/descriptors/symmetry_function/sym_fn.py:163-185.
 *
 * Where the reference FFI is called once per atom from a Python loop (sym_fn.py:155-160),
 * these entry points are batch-level: one call covers all centre atoms of a configuration (or of
 * many concatenated configurations).
 *
 * Conventions
 *   - plain C: pointers and sizes only; no torch / C++ types cross the boundary.
 *   - pointers named d_* are DEVICE pointers owned by the caller; h_* are HOST pointers.
 *   - `stream` is a cudaStream_t passed as void*; all device work is enqueued on it.  Calls that
 *     must return a size to the host (…_begin, …_count) synchronise that stream once; all others
 *     are asynchronous.
 *   - scratch memory is taken from the CUDA stream-ordered allocator (cudaMallocAsync) on
 *     `stream` and released before return or by the matching *_finish / *_destroy call.
 *   - return value: 0 = ok; 1 = the reference's own error class for that call (singular cell,
 *     atom collision / cell grid too large — what neighbor_list_bind.cpp:68-72,205-209 turn into
 *     RuntimeError); other positive codes are KB_ERR_*; negative = -(cudaError_t).
 *   - global mutable state: a relaxed-atomic launch counter (kb_launch_count), the one-time memory
 *     pool setup, and — only while kb_kernel_timing is switched on — a mutex-guarded list of
 *     recorded event pairs (open events are per host thread).  Every function may be called from
 *     any host thread; one CUDA context (device) per process/rank is assumed.
 */
#ifndef KLIFF_B200_H_
#define KLIFF_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define KB_OK 0
#define KB_ERR_REFERENCE 1      /* singular cell / collision / cell grid too large */
#define KB_ERR_INVALID 2        /* bad argument */
#define KB_ERR_LIMIT 3          /* a compiled-in limit (KB_MAX_*) exceeded */
#define KB_ERR_UNSUPPORTED 4    /* device is not sm_100 */

#define KB_MAX_SPECIES 8
#define KB_MAX_TWO 64           /* radial (g1,g2,g3) parameter sets */
#define KB_MAX_GROUPS 32        /* angular groups: one (family, eta) each */
#define KB_GROUP_WIDTH 8        /* (zeta, lambda) entries per angular group */

/* Kinds, as in the reference's hyper-parameter dictionary keys (sym_fn.py:244-247). */
#define KB_G1 1
#define KB_G2 2
#define KB_G3 3
#define KB_G4 4
#define KB_G5 5

/*
 * Descriptor definition = Descriptor::set_cutoff + add_descriptor state (sym_fn.cpp:71-106),
 * regrouped for the GPU: radial sets are listed one by one; angular sets are grouped by
 * (family, eta) so that the exponential is shared by up to KB_GROUP_WIDTH (zeta, lambda)
 * entries.  `*_out` is the descriptor index (running sum of rows in family insertion order,
 * sym_fn.cpp:96-100).  kliff_b200.legacy.descriptors builds this from the hyper-parameter dict.
 */
typedef struct kb_symfun_params
{
  int32_t n_species;
  int32_t n_desc;                                  /* D */
  int32_t n_two;
  int32_t n_groups;
  double rcut[KB_MAX_SPECIES * KB_MAX_SPECIES];    /* row-major [si][sj] */
  int32_t two_kind[KB_MAX_TWO];                    /* KB_G1 / KB_G2 / KB_G3 */
  int32_t two_out[KB_MAX_TWO];
  double two_p0[KB_MAX_TWO];                       /* g2: eta,  g3: kappa */
  double two_p1[KB_MAX_TWO];                       /* g2: Rs */
  int32_t grp_kind[KB_MAX_GROUPS];                 /* KB_G4 / KB_G5 */
  int32_t grp_n[KB_MAX_GROUPS];                    /* entries used, 1..KB_GROUP_WIDTH */
  double grp_eta[KB_MAX_GROUPS];
  double grp_zeta[KB_MAX_GROUPS][KB_GROUP_WIDTH];
  double grp_lambda[KB_MAX_GROUPS][KB_GROUP_WIDTH];
  int32_t grp_out[KB_MAX_GROUPS][KB_GROUP_WIDTH];
  /* 1 when every group's entry e is (zeta, lambda) = ({1,2,4,16}[e/2], {-1,+1}[e%2]) with unused
   * entries marked grp_out < 0 (set51/set30 and every Behler-style set): enables the
   * multiplication-only power chain.  0 = generic pow(). */
  int32_t canonical_zl;
  int32_t reserved;
} kb_symfun_params;

/* ------------------------------------------------------------------------------------------ */
const char* kb_version(void);
const char* kb_error_string(int code);
/* sm count / compute capability of the current device; KB_ERR_UNSUPPORTED unless cc 10.x */
int kb_device_info(int32_t* sm_count, int32_t* cc_major, int32_t* cc_minor);

/* ------------------------------------------------------------------------------------------ */
/* Padding (periodic image) atoms — replaces neighlist.create_paddings                        */
/* (neighbor_list_bind.cpp:168-252 -> nbl_create_paddings neighbor_list.cpp:283-418).          */
/* Output order and every coordinate bit match the reference: shift-major, atom-minor.         */
/* ------------------------------------------------------------------------------------------ */
typedef struct kb_pad_plan kb_pad_plan;
/* Counts paddings (synchronises once).  h_cell: 9 doubles, rows = lattice vectors; h_pbc: 3. */
int kb_pad_begin(int32_t natoms, double cutoff, const double* h_cell, const int32_t* h_pbc,
                 const double* d_coords, int64_t* h_npad, kb_pad_plan** plan, void* stream);
/* Writes the paddings (asynchronous) and frees the plan. d_species may be NULL with
 * d_pad_species NULL (species then follow from d_pad_image). */
int kb_pad_finish(kb_pad_plan* plan, const int32_t* d_species, double* d_pad_coords,
                  int32_t* d_pad_species, int32_t* d_pad_image, void* stream);
void kb_pad_destroy(kb_pad_plan* plan, void* stream);

/* ------------------------------------------------------------------------------------------ */
/* Cell-list neighbor build — replaces NeighList.build + get_neigh                             */
/* (neighbor_list_bind.cpp:41-146 -> nbl_build neighbor_list.cpp:78-243).                      */
/* CSR over ALL ntot atoms (atoms with need == 0 get empty lists), each list sorted ascending  */
/* ("canonical sorting"); as a set it equals the reference's list bit for bit.                 */
/* half != 0 keeps only neighbors j > i (no reference counterpart; SURVEY.md §8a note i).      */
/* ------------------------------------------------------------------------------------------ */
typedef struct kb_nl_plan kb_nl_plan;
/* d_need may be NULL: then atoms [0, n_need) need neighbors (contributing atoms first,
 * kliff/neighbor/neighbor.py:101-104).  Writes d_counts[ntot], d_offsets[ntot+1]; returns the
 * total number of indices and the longest list (synchronises once). */
int kb_nl_begin(int32_t ntot, int32_t n_need, const double* d_coords, const int32_t* d_need,
                double influence_distance, double cutoff, int32_t half, int32_t* d_counts,
                int64_t* d_offsets, int64_t* h_total, int32_t* h_maxn, kb_nl_plan** plan,
                void* stream);
int kb_nl_finish(kb_nl_plan* plan, int32_t* d_indices, void* stream);
void kb_nl_destroy(kb_nl_plan* plan, void* stream);

/* ------------------------------------------------------------------------------------------ */
/* Batched paddings + neighbor lists for MANY configurations in one call — replaces the        */
/* per-configuration loop `for conf in configs: NeighborList(conf, infl_dist, padding_need_    */
/* neigh=False)` of Descriptor.generate_fingerprints / SymmetryFunction.transform              */
/* (descriptor.py:186-238, sym_fn.py:118-133 -> neighbor.py:88-126).                           */
/* One CTA per configuration; per configuration the result equals kb_pad_* followed by         */
/* kb_nl_* bit for bit.  Output is the packed batch layout: per configuration its              */
/* contributing atoms then its image atoms, atom ids global to the batch.                      */
/* ------------------------------------------------------------------------------------------ */
typedef struct kb_batch_plan kb_batch_plan;
/* h_cfg_ptr[n_cfg+1]: contributing-atom range of every configuration in d_coords[.,3];
 * h_cells[n_cfg*9] (rows = lattice vectors); h_pbc[n_cfg*3].  Counts image atoms
 * (synchronises once) and returns the total number of atoms (contributing + images). */
int kb_batch_begin(int32_t n_cfg, const int32_t* h_cfg_ptr, const double* h_cells,
                   const int32_t* h_pbc, double cutoff, const double* d_coords,
                   int64_t* h_total_atoms, kb_batch_plan** plan, void* stream);
/* Writes d_coords_all[A,3], d_image[A] (global ordinal of the owning contributing atom),
 * d_centers[n_contrib] (atom id of every contributing atom), d_atom_ptr[n_cfg+1], d_counts[A],
 * d_offsets[A+1]; returns the number of neighbor indices and the longest list (synchronises
 * once).  KB_ERR_REFERENCE on atom collisions (neighbor_list.cpp:199). */
int kb_batch_neighbors(kb_batch_plan* plan, double* d_coords_all, int32_t* d_image,
                       int32_t* d_centers, int64_t* d_atom_ptr, int32_t* d_counts,
                       int64_t* d_offsets, int64_t* h_total_pairs, int32_t* h_maxn, void* stream);
/* Writes the neighbor indices (ascending per atom, asynchronous) and frees the plan. */
int kb_batch_finish(kb_batch_plan* plan, const double* d_coords_all, const int64_t* d_offsets,
                    int32_t* d_indices, void* stream);
void kb_batch_destroy(kb_batch_plan* plan, void* stream);

/* ------------------------------------------------------------------------------------------ */
/* Symmetry functions + dG/dr — replaces the per-atom loop over sf.Descriptor.generate_one_atom */
/* (sym_fn.py:155-160 -> sym_fn_bind.cpp:38-94 -> sym_fn.cpp:416-602).                         */
/*                                                                                            */
/* Centre atom t (0 <= t < n_centers) is atom a = d_centers[t] (or t when d_centers == NULL);  */
/* its neighbors are d_nl_indices[d_nl_offsets[a] .. d_nl_offsets[a+1]).                       */
/* Outputs: d_zeta[t*D + d]; and, when d_dz != NULL, the reference's per-atom gradient block   */
/* [D][n+1][3] (slot n = centre atom, sym_fn.hpp:188-191) stored at d_dz + d_dz_ptr[t].        */
/* d_dz_ptr comes from kb_symfun_block_offsets: blocks are padded to 16 doubles (128 B).       */
/* ------------------------------------------------------------------------------------------ */
/* Storage type of the packed dG/dr blocks.  Arithmetic is always fp64; KB_F32 rounds the finished
 * block to float on the way out (what the reference does with its default dtype=np.float32,
 * kliff/legacy/descriptors/descriptor.py:197-205) and halves HBM footprint and contraction traffic. */
#define KB_F64 0
#define KB_F32 1

#define KB_SYMFUN_AUTO 0
#define KB_SYMFUN_GENERAL 1     /* global-atomics kernel: any neighbor count */
#define KB_SYMFUN_TILED 2       /* CTA-per-atom shared-memory kernel: n <= 64, any families */
#define KB_SYMFUN_WARP 3        /* warp-per-atom kernel: n <= 64, g1-g4 with zeta in {1,2,4,16}, lambda = +-1 */
#define KB_SYMFUN_SPLIT 4       /* same domain as WARP, two launches: table kernel + triple-sweep kernel (round-1 sweep) */
/* 5 was the warp-specialised single kernel of round 1 (measured 1.7x slower; removed from the build) */
#define KB_SYMFUN_QUEUE 6       /* two launches, table kernel + queue sweep (fp64 blocks, shapes of set51 / set30:
                                   AUTO's first choice); falls back to SPLIT's sweep elsewhere */

int kb_symfun_block_offsets(int32_t n_centers, const int32_t* d_centers,
                            const int64_t* d_nl_offsets, int32_t n_desc, int64_t* d_dz_ptr,
                            int64_t* h_total, void* stream);
int kb_symfun_forward(const kb_symfun_params* h_params, int32_t n_centers,
                      const int32_t* d_centers, const double* d_coords, const int32_t* d_species,
                      const int64_t* d_nl_offsets, const int32_t* d_nl_indices, int32_t maxn,
                      double* d_zeta, void* d_dz, int32_t dz_dtype, const int64_t* d_dz_ptr,
                      int32_t mode, void* stream);

/* ------------------------------------------------------------------------------------------ */
/* Chain-rule force assembly — replaces CalculatorTorch._compute_forces on the dense           */
/* (N, D, 3N) tensor (calculator_torch.py:207-213, 266-269) by a contraction over the packed   */
/* blocks:  F[image[a], c] -= sum_d dEdG[t, d] * dz_t[d, slot(a), c].                           */
/* fwd accumulates into d_forces[n_atoms_out*3] (caller zeroes); bwd is its adjoint w.r.t.     */
/* dEdG (needed because the loss differentiates through dE/dG, calculator_torch.py:200-202):   */
/*   gdEdG[t, d] = - sum_{slot,c} gF[image[a], c] * dz_t[d, slot, c].                           */
/* scale (nullable, [D]) multiplies dEdG / gdEdG per descriptor: the 1/stdev normalisation of  */
/* kliff/legacy/descriptors/descriptor.py:187-194 folded into the contraction.                 */
/* ------------------------------------------------------------------------------------------ */
int kb_force_scatter_fwd(int32_t n_centers, const int32_t* d_centers, const int64_t* d_nl_offsets,
                         const int32_t* d_nl_indices, const int32_t* d_image, int32_t n_desc,
                         const double* d_dEdG, const double* d_scale, const void* d_dz,
                         int32_t dz_dtype, const int64_t* d_dz_ptr, double* d_forces, void* stream);
int kb_force_scatter_bwd(int32_t n_centers, const int32_t* d_centers, const int64_t* d_nl_offsets,
                         const int32_t* d_nl_indices, const int32_t* d_image, int32_t n_desc,
                         const double* d_gF, const double* d_scale, const void* d_dz,
                         int32_t dz_dtype, const int64_t* d_dz_ptr, double* d_gdEdG, void* stream);

/* ------------------------------------------------------------------------------------------ */
/* Dense views for small configurations — what SymmetryFunction.transform returns             */
/* (sym_fn.py:163-185).  d_dense_forces: [n_centers][D][3*n_contrib] (caller zeroes);          */
/* d_dense_stress: [n_centers][D][6], Voigt 11,22,33,23,31,12.  Either may be NULL.            */
/* ------------------------------------------------------------------------------------------ */
int kb_dense_fold(int32_t n_centers, const int32_t* d_centers, const int64_t* d_nl_offsets,
                  const int32_t* d_nl_indices, const int32_t* d_image, const double* d_coords,
                  int32_t n_desc, int32_t n_contrib, const void* d_dz, int32_t dz_dtype,
                  const int64_t* d_dz_ptr, double* d_dense_forces, double* d_dense_stress,
                  void* stream);

/* ------------------------------------------------------------------------------------------ */
/* Measurement helpers (bench.py): FP64 FMA peak of this device, TFLOP/s (2 flops per FMA),    */
/* measured with an unrolled register-resident DFMA chain; and an HBM copy probe, GB/s.        */
/* ------------------------------------------------------------------------------------------ */
int kb_probe_fp64_peak(double* h_tflops, void* stream);
int kb_probe_hbm_copy(size_t bytes, double* h_gbs, void* stream);
/* Returns the scratch the library's stream-ordered allocations have cached (cudaMallocAsync on the
 * device's default pool, kept across calls so that repeated passes do not re-allocate) to the driver,
 * down to keep_bytes; synchronises the device.  Optional: call it after a large descriptor pass when
 * another allocator in the process (torch) needs the memory. */
int kb_pool_trim(size_t keep_bytes);
/* Number of kernels this library has launched in this process (reset != 0 zeroes it). */
int64_t kb_launch_count(int32_t reset);
/* Per-kernel device times of the descriptor stage, measured with CUDA events on the launching
 * stream.  Returns in h_ms[5] / h_launches[5] (either may be NULL) the summed duration and number
 * of launches recorded since the previous call, by kernel: [0] symfun_prep_kernel, [1]
 * symfun_sweep3_kernel / symfun_sweep_kernel / symfun_qsweep_kernel, [2] fused symfun_warp_kernel,
 * [3] force_scatter / force_gather kernels, [4] the whole two-stage descriptor path of a call (all its
 * chunks, table kernel + sweep); then switches recording on (enable != 0) or off.  Off by default: no events
 * are created unless a caller (bench.py) asks. */
int kb_kernel_timing(int32_t enable, double* h_ms, int32_t* h_launches);

/* ------------------------------------------------------------------------------------------ */
/* Owner-sorted (segmented) force assembly: the same contraction as kb_force_scatter_fwd without  */
/* floating-point atomics — forces are bit-identical from run to run (calculator_torch.py:207-213, */
/* 266-269; image folding of neighbor.py:260-296).                                                 */
/* Plan (integer work, once per neighbor list): slot id of (centre t, slot s) = d_slot_ptr[t] + s   */
/* with slot n_t = the centre itself; d_seg_src[d_seg_ptr[o] .. d_seg_ptr[o+1]) = the slot ids      */
/* whose atom is owned by contributing atom o (image[atom] == o), ascending.                       */
/*   kb_scatter_plan_begin  writes d_slot_ptr[n_centers+1], returns the number of slots (syncs)    */
/*   kb_scatter_plan_finish writes d_seg_ptr[n_out+1], d_seg_src[total_slots] (syncs; KB_ERR_INVALID */
/*                          when an image index lies outside [0, n_out))                           */
/* kb_force_scatter_fwd_seg accumulates into d_forces[3 n_out] (caller zeroes) for a CLOSED range  */
/* of centres (whole configurations: every slot owned by one of the n_out atoms comes from these   */
/* centres).  For a sub-range of a larger plan pass d_slot_ptr + lo, d_seg_ptr + lo, the plan's    */
/* d_seg_src, slot_base = slot_ptr[lo] - slot_ptr[0] and n_slots = slot_ptr[hi] - slot_ptr[lo].    */
/* ------------------------------------------------------------------------------------------ */
int kb_scatter_plan_begin(int32_t n_centers, const int32_t* d_centers, const int64_t* d_nl_offsets,
                          int64_t* d_slot_ptr, int64_t* h_total_slots, void* stream);
int kb_scatter_plan_finish(int32_t n_centers, const int32_t* d_centers, const int64_t* d_nl_offsets,
                           const int32_t* d_nl_indices, const int32_t* d_image, int32_t n_out,
                           const int64_t* d_slot_ptr, int64_t total_slots, int64_t* d_seg_ptr,
                           int32_t* d_seg_src, void* stream);
int kb_force_scatter_fwd_seg(int32_t n_centers, const int32_t* d_centers, const int64_t* d_nl_offsets,
                             int32_t n_desc, const double* d_dEdG, const double* d_scale,
                             const void* d_dz, int32_t dz_dtype, const int64_t* d_dz_ptr,
                             const int64_t* d_slot_ptr, int64_t n_slots, int64_t slot_base,
                             int32_t n_out, const int64_t* d_seg_ptr, const int32_t* d_seg_src,
                             double* d_contrib /* scratch [3 n_slots] or NULL */, double* d_forces,
                             void* stream);

/* ------------------------------------------------------------------------------------------ */
/* Column statistics and normalisation of the descriptor matrix — replaces np.mean / np.std and  */
/* (zeta - mean) / stdev of Descriptor.generate_fingerprints                                     */
/* (kliff/legacy/descriptors/descriptor.py:116-118, 187-194).  Deterministic two-stage sums.      */
/* kb_moments: d_out[d] = sum_r zeta[r][d] when d_mean == NULL, else sum_r (zeta[r][d] - mean[d])^2 */
/* (numpy's two-pass definition: the caller divides by the row count, all-reduces across ranks).  */
/* kb_normalize: d_out[r][d] = (zeta[r][d] - mean[d]) * inv_stdev[d]  (d_out may alias d_zeta).    */
/* ------------------------------------------------------------------------------------------ */
int kb_moments(int64_t n_rows, int32_t n_desc, const double* d_zeta, const double* d_mean,
               double* d_out, void* stream);
int kb_normalize(int64_t n_rows, int32_t n_desc, const double* d_zeta, const double* d_mean,
                 const double* d_inv_stdev, double* d_out, void* stream);

/* ------------------------------------------------------------------------------------------ */
/* Fused training step for small tanh MLPs over the packed fingerprints — replaces, for one     */
/* batch, CalculatorTorch.compute (calculator_torch.py:161-228: model(zeta), autograd.grad with  */
/* create_graph, the force contraction), Loss._get_loss_batch / energy_forces_residual            */
/* (loss.py:37-110, 843-920), loss.backward() and torch.optim.Adam.step().                        */
/* Network: x[n_in] -> (Linear(width[l]) -> tanh) x n_hidden -> Linear(1); parameters flat in     */
/* torch order (W_1 [width_0][n_in] row-major, b_1, ..., W_out [1][width_last], b_out) in `dtype`. */
/* ------------------------------------------------------------------------------------------ */
#define KB_MLP_MAX_IN 128
#define KB_MLP_MAX_HIDDEN 4
#define KB_MLP_MAX_WIDTH 32
typedef struct kb_mlp_desc
{
  int32_t n_in;
  int32_t n_hidden;
  int32_t width[KB_MLP_MAX_HIDDEN];
  int32_t dtype;      /* KB_F32 / KB_F64: type of x, theta and of the optimizer moments */
  int32_t reserved;
} kb_mlp_desc;
/* number of parameters and length (in `dtype` elements) of one atom's row of d_atom_vecs */
int kb_mlp_num_params(const kb_mlp_desc* M, int32_t* n_params, int32_t* atom_vec);
/* e_atom[n] (f64) and, when d_dEdx != NULL, dE/dx [n][n_in] (f64) */
int kb_mlp_energy_grad(const kb_mlp_desc* M, int32_t n_atoms, const void* d_x, const void* d_theta,
                       double* d_e_atom, double* d_dEdx, void* stream);
/* Per configuration c of the batch (atoms [d_cfg_ptr[c] - atom0, d_cfg_ptr[c+1] - atom0) of the batch
 * arrays): E_c = sum e_atom; r_e = w_e[c] (E_c - e_ref[c]); r_f = w_f[c] (F - f_ref) per component;
 * *d_loss += (r_e^2 + sum r_f^2) * inv_nglobal; d_a[atom] = dL/de_atom; d_gF = dL/dF.
 * w_e / w_f already hold config_weight * energy_weight (forces_weight) / natoms (loss.py:100-109). */
int kb_train_residual(int32_t n_cfg, const int64_t* d_cfg_ptr, int64_t atom0, const double* d_e_atom,
                      const double* d_forces, const double* d_e_ref, const double* d_f_ref,
                      const double* d_w_e, const double* d_w_f, double inv_nglobal, int32_t use_energy,
                      int32_t use_forces, double* d_a, double* d_gF, double* d_loss, void* stream);
/* Parameter gradients d_grad[n_params] (f64) of L for a = dL/de_atom (d_a) and G = dL/d(dE/dx) (d_G, may
 * be NULL: energy-only).  d_atom_vecs: scratch [n_atoms][atom_vec] in `dtype`; d_partials: scratch
 * f64 [ceil(n_atoms / 32)][n_params].  Deterministic (no atomics). */
int kb_mlp_param_grad(const kb_mlp_desc* M, int32_t n_atoms, const void* d_x, const void* d_theta,
                      const double* d_a, const double* d_G, void* d_atom_vecs, double* d_partials,
                      double* d_grad, void* stream);
/* torch.optim.Adam's update (no amsgrad / weight decay) on the flat buffer, gradient = d_grad * grad_scale;
 * *d_step (f64, device) is incremented: capturable in a CUDA graph. */
int kb_adam_step(int32_t n_params, void* d_theta, const double* d_grad, double grad_scale, void* d_m,
                 void* d_v, double* d_step, double lr, double beta1, double beta2, double eps,
                 int32_t dtype, void* stream);

/* ------------------------------------------------------------------------------------------ */
/* The same training step as six launches (csrc/train_fast.cu), for networks with n_in <= 64,   */
/* widths <= 16 and lists of <= 84 neighbors (kb_train_fast_ok): MLP forward, force contraction,  */
/* owner-sorted forces + residuals per configuration, adjoint gather, second-order backward +     */
/* parameter-gradient partials, reduction (+ Adam).  Same references as above; the owner-sorted   */
/* plan arguments are those of kb_force_scatter_fwd_seg.                                          */
/* ------------------------------------------------------------------------------------------ */
int kb_train_fast_ok(const kb_mlp_desc* M, int32_t maxn);
/* e_atom[n] and, when d_contrib != NULL, the slot contributions [slots of the batch][3] (f64);
 * d_w: scratch f64 [n_atoms][n_in] (dE/dx * scale); d_dz_ptr / d_slot_ptr: the batch's [n_atoms + 1] slices */
int kb_train_forward(const kb_mlp_desc* M, int32_t n_atoms, const void* d_x, const void* d_theta,
                     const double* d_scale, const void* d_dz, int32_t dz_dtype, const int64_t* d_dz_ptr,
                     const int64_t* d_slot_ptr, double* d_e_atom, double* d_w, double* d_contrib, void* stream);
/* kb_train_residual with the owner-sorted force sums folded in; one loss term per configuration in
 * d_loss_cfg (no atomics); *d_step (Adam's step counter, may be NULL) is incremented.  d_contrib is read
 * (slot contributions) and then overwritten, slot by slot, with dL/dF of the slot's owner atom: the row
 * kb_train_backward reads. */
int kb_train_residual_seg(int32_t n_cfg, const int64_t* d_cfg_ptr, int64_t atom0, const double* d_e_atom,
                          const int64_t* d_seg_ptr, const int32_t* d_seg_src, int64_t slot_base,
                          double* d_contrib, const double* d_e_ref, const double* d_f_ref,
                          const double* d_w_e, const double* d_w_f, double inv_nglobal, int32_t use_energy,
                          int32_t use_forces, double* d_a, double* d_gF, double* d_forces,
                          double* d_loss_cfg, double* d_step, void* stream);
/* d_partials [partial_rows][n_params] (f64): per-block partial parameter gradients, *h_rows_used of them
 * written; d_gslot (dL/dF per slot, from kb_train_residual_seg) == NULL: energy-only; d_G: scratch f64
 * [n_atoms][n_in] (dL/d(dE/dx)). */
int kb_train_backward(const kb_mlp_desc* M, int32_t n_atoms, const void* d_x, const void* d_theta,
                      const double* d_scale, const void* d_dz, int32_t dz_dtype, const int64_t* d_dz_ptr,
                      const int64_t* d_slot_ptr, const double* d_gslot, const double* d_a, double* d_G,
                      double* d_partials, int32_t partial_rows, int32_t* h_rows_used, void* stream);
/* d_grad[0..P) = sum of the partial rows, d_grad[P] = sum of d_loss_cfg; apply_adam: Adam's update in the
 * same launch with step number *d_step; *d_epoch_loss += d_grad[P] when not NULL. */
int kb_train_update(const kb_mlp_desc* M, int32_t rows_used, const double* d_partials, int32_t n_cfg,
                    const double* d_loss_cfg, double* d_grad, int32_t apply_adam, void* d_theta, void* d_m,
                    void* d_v, const double* d_step, double lr, double beta1, double beta2, double eps,
                    double* d_epoch_loss, void* stream);
/* Adam on d_grad[0..P) after an all-reduce (step number *d_step, already counted); *d_epoch_loss += d_grad[P]. */
int kb_train_adam(int32_t n_params, const double* d_grad, void* d_theta, void* d_m, void* d_v,
                  const double* d_step, double lr, double beta1, double beta2, double eps, int32_t dtype,
                  double* d_epoch_loss, void* stream);

/* ------------------------------------------------------------------------------------------ */
/* Data-parallel form of kb_train_update (csrc/peer.cu): the gradient all-reduce over NVLink    */
/* peer memory inside the same launch as the partial reduction and Adam — replaces               */
/* ncclAllReduce + optimizer.step() of a sharded Loss.minimize step (loss.py:843-920).           */
/* Every rank (one process per GPU of one node, world <= 16) allocates a workspace with          */
/* kb_peer_alloc (cudaMalloc + IPC handle, zeroed), the 64-byte handles are exchanged by the     */
/* host (torch.distributed), kb_peer_open maps a peer's workspace.  h_peer_ws[r] = rank r's       */
/* workspace as mapped in this process (own pointer at [rank]).  The sum over ranks is formed in  */
/* rank order on every rank: identical bits everywhere.  kb_peer_error reads the workspace's      */
/* error word (non-zero: a peer did not arrive within 4 s; the step then used what was there).    */
/* ------------------------------------------------------------------------------------------ */
int kb_peer_ws_bytes(int32_t world, int32_t n_values, int64_t* bytes, int32_t* padded);
int kb_peer_alloc(int64_t bytes, void** d_ptr, void* h_handle64);
int kb_peer_open(const void* h_handle64, void** d_ptr);
int kb_peer_close(void* d_ptr);
int kb_peer_free(void* d_ptr);
int kb_peer_error(const void* d_ws, int32_t* h_err, void* stream);
int kb_train_update_peer(int32_t n_params, int32_t rows_used, const double* d_partials, int32_t n_cfg,
                         const double* d_loss_cfg, double* d_grad, int32_t apply_adam, void* d_theta,
                         void* d_m, void* d_v, const double* d_step, double lr, double beta1, double beta2,
                         double eps, int32_t dtype, double* d_epoch_loss, void* const* h_peer_ws,
                         int32_t world, int32_t rank, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* KLIFF_B200_H_ */
