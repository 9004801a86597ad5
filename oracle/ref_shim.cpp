// TEST INFRASTRUCTURE ONLY — never linked or loaded by the product path.
//
// extern "C" doorway onto the UNMODIFIED reference C++ (openkim/kliff), compiled from the
// sources where they lie under /root/reference by oracle/Makefile into
// oracle/_ref/libkliff_ref.so.  Nothing of the reference is copied: this file only
// includes the reference headers and forwards calls.
//
//   kref_create_paddings -> nbl_create_paddings  kliff/neighbor/neighbor_list.cpp:283-418
//   kref_build/get_neigh -> nbl_build / nbl_get_neigh  neighbor_list.cpp:78-243, 246-280
//   kref_desc_*          -> class Descriptor  kliff/legacy/descriptors/symmetry_function/sym_fn.cpp
//                           (set_cutoff :71-79, add_descriptor :86-106, generate_one_atom :416-602)
#include <cstring>
#include <vector>

#include "neighbor_list.h"
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <iostream>
#include <map>
#include <numeric>
#include <sstream>
#include <string>
// the parsed centring data has only `inline` accessors defined inside sym_fn.cpp (not linkable from here):
// read the two private vectors directly.  The reference sources themselves stay untouched.
#define private public
#include "sym_fn.hpp"
#undef private

#include <cstdio>

extern "C" {

// ---- paddings: two-call protocol (first call with out pointers NULL returns the count) ----
int kref_create_paddings(int natoms, double cutoff, const double* cell, const int* pbc,
                         const double* coords, const int* species, int capacity,
                         double* pad_coords, int* pad_species, int* pad_image, int* npad)
{
  int n = 0;
  std::vector<double> c;
  std::vector<int> s, m;
  int err = nbl_create_paddings(natoms, cutoff, cell, pbc, coords, species, n, c, s, m);
  if (err) return err;
  *npad = n;
  if (pad_coords && capacity >= n)
  {
    if (n) {
      std::memcpy(pad_coords, c.data(), sizeof(double) * 3 * n);
      std::memcpy(pad_species, s.data(), sizeof(int) * n);
      std::memcpy(pad_image, m.data(), sizeof(int) * n);
    }
  }
  return 0;
}

// ---- neighbor list handle ----
void* kref_nl_create()
{
  NeighList* nl = nullptr;
  nbl_initialize(&nl);
  return nl;
}

void kref_nl_free(void* h)
{
  NeighList* nl = static_cast<NeighList*>(h);
  nbl_clean(&nl);
}

int kref_nl_build(void* h, int natoms, const double* coords, double infl, int ncut,
                  const double* cutoffs, const int* need)
{
  return nbl_build(static_cast<NeighList*>(h), natoms, coords, infl, ncut, cutoffs, need);
}

// copies the whole CSR of list `which` out: counts[natoms], then indices (raw reference order)
long long kref_nl_total(void* h, int which)
{
  NeighList* nl = static_cast<NeighList*>(h);
  NeighListOne* one = &nl->lists[which];
  long long tot = 0;
  for (int i = 0; i < one->numberOfParticles; i++) tot += one->Nneighbors[i];
  return tot;
}

int kref_nl_export(void* h, int which, int* counts, long long* offsets, int* indices)
{
  NeighList* nl = static_cast<NeighList*>(h);
  NeighListOne* one = &nl->lists[which];
  long long tot = 0;
  for (int i = 0; i < one->numberOfParticles; i++)
  {
    counts[i] = one->Nneighbors[i];
    offsets[i] = one->beginIndex[i];
    tot += one->Nneighbors[i];
  }
  offsets[one->numberOfParticles] = tot;
  if (tot) std::memcpy(indices, one->neighborList, sizeof(int) * tot);
  return 0;
}

int kref_nl_get_neigh(void* h, int ncut, const double* cutoffs, int which, int particle,
                      int* numnei, const int** neigh)
{
  return nbl_get_neigh(h, ncut, cutoffs, which, particle, numnei, neigh);
}

// ---- symmetry-function descriptor handle ----
void* kref_desc_create() { return new Descriptor(); }
void kref_desc_free(void* d) { delete static_cast<Descriptor*>(d); }

void kref_desc_set_cutoff(void* d, const char* name, int nspecies, const double* rcut)
{
  static_cast<Descriptor*>(d)->set_cutoff(name, nspecies, rcut);
}

void kref_desc_add(void* d, const char* name, const double* values, int rows, int cols)
{
  static_cast<Descriptor*>(d)->add_descriptor(name, values, rows, cols);
}

int kref_desc_size(void* d) { return static_cast<Descriptor*>(d)->get_num_descriptors(); }

// Descriptor::read_parameter_file (sym_fn.cpp:108-387) on a `descriptor.params` text file: cutoffs,
// species, hyper-parameter tables, and the centring / normalising data.  Returns 0 on success (the
// reference returns `true` on a parse error), -1 when the file cannot be opened.
int kref_desc_read_file(void* d, const char* path)
{
  FILE* f = std::fopen(path, "r");
  if (!f) return -1;
  int err = static_cast<Descriptor*>(d)->read_parameter_file(f);
  std::fclose(f);
  return err;
}

// mean / standard deviation of feature i as parsed (sym_fn.cpp:394-412); returns need_normalize()
int kref_desc_mean_std(void* d, int i, double* mean, double* stdev)
{
  Descriptor* D = static_cast<Descriptor*>(d);
  const bool has = D->normalize_ && i >= 0 && (size_t) i < D->feature_mean_.size();
  if (has) { *mean = D->feature_mean_[i]; *stdev = D->feature_std_[i]; }
  return has ? 1 : 0;
}

// zeta[D] and grad[D*3*(numnei+1)] are zeroed here, as the pybind wrapper does
// (sym_fn_bind.cpp:49-52) before calling generate_one_atom.
void kref_desc_one_atom(void* d, int i, const double* coords, const int* species,
                        const int* neigh, int numnei, double* zeta, double* grad, int want_grad)
{
  Descriptor* D = static_cast<Descriptor*>(d);
  int nd = D->get_num_descriptors();
  for (int a = 0; a < nd; a++) zeta[a] = 0.0;
  if (grad)
    for (long a = 0; a < (long) nd * 3 * (numnei + 1); a++) grad[a] = 0.0;
  D->generate_one_atom(i, coords, species, neigh, numnei, zeta, grad, want_grad != 0);
}

// loop form used by the CPU-baseline timing: atoms [first,last) of a CSR neighbor list;
// grad blocks are written to a scratch buffer of size D*3*(maxnei+1) and discarded, zeta kept.
void kref_desc_range(void* d, int first, int last, const double* coords, const int* species,
                     const long long* offsets, const int* indices, double* zeta_out,
                     double* scratch, int want_grad)
{
  Descriptor* D = static_cast<Descriptor*>(d);
  int nd = D->get_num_descriptors();
  for (int i = first; i < last; i++)
  {
    int n = (int) (offsets[i + 1] - offsets[i]);
    double* z = zeta_out + (size_t) (i - first) * nd;
    for (int a = 0; a < nd; a++) z[a] = 0.0;
    if (want_grad)
      for (long a = 0; a < (long) nd * 3 * (n + 1); a++) scratch[a] = 0.0;
    D->generate_one_atom(i, coords, species, indices + offsets[i], n, z, scratch, want_grad != 0);
  }
}

}  // extern "C"
