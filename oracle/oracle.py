"""
TEST INFRASTRUCTURE ONLY — numpy/ctypes front end of the CPU oracle.

Two back ends with one interface:

* ``Oracle("port")``      -> oracle/libkb_oracle.so, the plain-C restatement (kb_oracle.c);
* ``Oracle("reference")`` -> oracle/_ref/libkliff_ref.so, the UNMODIFIED reference C++
  (kliff/neighbor/neighbor_list.cpp, kliff/legacy/descriptors/symmetry_function/sym_fn.cpp)
  behind oracle/ref_shim.cpp.  Built in the build container by ``make -C oracle ref``; it
  travels to the GPU box as a prebuilt file (``/root/reference`` does not exist there).

Only tests/, __graft_entry__.smoke() and bench.py (cpu_baseline / --impl reference) may import
this module.  The product package kliff_b200 never does.

The numpy helpers at the bottom restate the Python-side steps of the reference path:
``dense_fold`` = kliff/legacy/descriptors/symmetry_function/sym_fn.py:163-174,
``stress_fold`` = sym_fn.py:176-185, ``forces_from_dense`` =
kliff/legacy/calculators/calculator_torch.py:266-269.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_PORT = os.path.join(_HERE, "libkb_oracle.so")
_REF = os.path.join(_HERE, "_ref", "libkliff_ref.so")

KIND = {"g1": 1, "g2": 2, "g3": 3, "g4": 4, "g5": 5}
NCOLS = {1: 1, 2: 2, 3: 1, 4: 3, 5: 3}

_dp = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
_ip = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
_lp = np.ctypeslib.ndpointer(dtype=np.int64, flags="C_CONTIGUOUS")


def build(force=False):
    """Compile the C restatement (always possible: gcc only) and, when the reference tree is
    present, the reference shim."""
    if force or not os.path.exists(_PORT):
        subprocess.check_call(["make", "-C", _HERE, "oracle"], stdout=subprocess.DEVNULL)
    if os.path.isdir("/root/reference/kliff") and (force or not os.path.exists(_REF)):
        subprocess.check_call(["make", "-C", _HERE, "ref"], stdout=subprocess.DEVNULL)


def have_reference():
    return os.path.exists(_REF)


def fams_from_hyperparams(hyperparams):
    """OrderedDict {'g2': [{'eta','Rs'}...], 'g4': [{'zeta','lambda','eta'}...]} -> list of
    (kind, float64[rows, cols]) in insertion order; same packing as sym_fn.py:237-283."""
    fams = []
    for key, rows in hyperparams.items():
        kind = KIND[key.lower()]
        if kind == 1:
            vals = np.zeros((1, 1))
        elif kind == 2:
            vals = np.array([[r["eta"], r["Rs"]] for r in rows], dtype=np.float64)
        elif kind == 3:
            vals = np.array([[r["kappa"]] for r in rows], dtype=np.float64)
        else:
            vals = np.array([[r["zeta"], r["lambda"], r["eta"]] for r in rows], dtype=np.float64)
        fams.append((kind, np.ascontiguousarray(vals)))
    return fams


class Oracle:
    def __init__(self, kind="port"):
        self.kind = kind
        if kind == "port":
            build()
            self.lib = C.CDLL(_PORT)
            self._bind_port()
        elif kind == "reference":
            if not os.path.exists(_REF):
                build()
            if not os.path.exists(_REF):
                raise FileNotFoundError(_REF)
            self.lib = C.CDLL(_REF)
            self._bind_ref()
        else:
            raise ValueError(kind)

    # ------------------------------------------------------------------ bindings
    def _bind_port(self):
        L = self.lib
        L.kbo_create_paddings.restype = C.c_int
        L.kbo_create_paddings.argtypes = [C.c_int, C.c_double, _dp, _ip, _dp, _ip, C.c_long,
                                          C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_long)]
        L.kbo_nl_build.restype = C.c_int
        L.kbo_nl_build.argtypes = [C.c_int, _dp, C.c_double, C.c_double, _ip, _ip, _lp, C.c_void_p]
        L.kbo_symfun_atom.restype = None
        L.kbo_symfun_atom.argtypes = [C.c_int, _dp, C.c_int, _ip, _ip, _dp, C.c_int, _dp, _ip,
                                      _ip, C.c_int, _dp, C.c_void_p, C.c_int]
        L.kbo_symfun_range.restype = None
        L.kbo_symfun_range.argtypes = [C.c_int, _dp, C.c_int, _ip, _ip, _dp, C.c_int, C.c_int,
                                       C.c_int, _dp, _ip, _lp, _ip, _dp, C.c_void_p, _dp, C.c_int]

    def _bind_ref(self):
        L = self.lib
        L.kref_create_paddings.restype = C.c_int
        L.kref_create_paddings.argtypes = [C.c_int, C.c_double, _dp, _ip, _dp, _ip, C.c_int,
                                           C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_int)]
        L.kref_nl_create.restype = C.c_void_p
        L.kref_nl_free.argtypes = [C.c_void_p]
        L.kref_nl_build.restype = C.c_int
        L.kref_nl_build.argtypes = [C.c_void_p, C.c_int, _dp, C.c_double, C.c_int, _dp, _ip]
        L.kref_nl_total.restype = C.c_longlong
        L.kref_nl_total.argtypes = [C.c_void_p, C.c_int]
        L.kref_nl_export.argtypes = [C.c_void_p, C.c_int, _ip, _lp, _ip]
        L.kref_desc_create.restype = C.c_void_p
        L.kref_desc_free.argtypes = [C.c_void_p]
        L.kref_desc_set_cutoff.argtypes = [C.c_void_p, C.c_char_p, C.c_int, _dp]
        L.kref_desc_add.argtypes = [C.c_void_p, C.c_char_p, _dp, C.c_int, C.c_int]
        L.kref_desc_size.restype = C.c_int
        L.kref_desc_size.argtypes = [C.c_void_p]
        L.kref_desc_read_file.restype = C.c_int
        L.kref_desc_read_file.argtypes = [C.c_void_p, C.c_char_p]
        L.kref_desc_mean_std.restype = C.c_int
        L.kref_desc_mean_std.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double)]
        L.kref_desc_one_atom.argtypes = [C.c_void_p, C.c_int, _dp, _ip, _ip, C.c_int, _dp,
                                         C.c_void_p, C.c_int]
        L.kref_desc_range.argtypes = [C.c_void_p, C.c_int, C.c_int, _dp, _ip, _lp, _ip, _dp, _dp,
                                      C.c_int]

    # ------------------------------------------------------------------ paddings
    def create_paddings(self, cutoff, cell, pbc, coords, species):
        """-> (pad_coords f64[P,3], pad_species i32[P], pad_image i32[P]); raises RuntimeError
        on a singular cell like neighbor_list_bind.cpp:205-209."""
        cell = np.ascontiguousarray(cell, dtype=np.float64).reshape(9)
        pbc = np.ascontiguousarray(pbc, dtype=np.int32)
        coords = np.ascontiguousarray(coords, dtype=np.float64).reshape(-1, 3)
        species = np.ascontiguousarray(species, dtype=np.int32)
        n = coords.shape[0]
        if self.kind == "port":
            cnt = C.c_long(0)
            err = self.lib.kbo_create_paddings(n, cutoff, cell, pbc, coords, species, 0, None,
                                               None, None, C.byref(cnt))
            if err:
                raise RuntimeError("singular cell")
            P = cnt.value
            pc = np.empty((P, 3)); ps = np.empty(P, np.int32); pi = np.empty(P, np.int32)
            if P:
                self.lib.kbo_create_paddings(n, cutoff, cell, pbc, coords, species, P,
                                             pc.ctypes.data, ps.ctypes.data, pi.ctypes.data,
                                             C.byref(cnt))
            return pc, ps, pi
        cnt = C.c_int(0)
        err = self.lib.kref_create_paddings(n, cutoff, cell, pbc, coords, species, 0, None, None,
                                            None, C.byref(cnt))
        if err:
            raise RuntimeError("singular cell")
        P = cnt.value
        pc = np.empty((P, 3)); ps = np.empty(P, np.int32); pi = np.empty(P, np.int32)
        if P:
            self.lib.kref_create_paddings(n, cutoff, cell, pbc, coords, species, P,
                                          pc.ctypes.data, ps.ctypes.data, pi.ctypes.data,
                                          C.byref(cnt))
        return pc, ps, pi

    # ------------------------------------------------------------------ neighbor list
    def build_neighbors(self, coords, infl, need, cutoff=None):
        """-> (counts i32[Ntot], offsets i64[Ntot+1], indices i32[sum]) in the reference's
        raw cell-scan order; raises RuntimeError on collision / oversize like
        neighbor_list_bind.cpp:68-72."""
        coords = np.ascontiguousarray(coords, dtype=np.float64).reshape(-1, 3)
        need = np.ascontiguousarray(need, dtype=np.int32)
        n = coords.shape[0]
        cutoff = infl if cutoff is None else cutoff
        counts = np.zeros(n, np.int32)
        offsets = np.zeros(n + 1, np.int64)
        if self.kind == "port":
            if self.lib.kbo_nl_build(n, coords, infl, cutoff, need, counts, offsets, None):
                raise RuntimeError("neighbor build failed")
            idx = np.empty(int(offsets[-1]), np.int32)
            self.lib.kbo_nl_build(n, coords, infl, cutoff, need, counts, offsets,
                                  idx.ctypes.data if idx.size else None)
            return counts, offsets, idx
        h = self.lib.kref_nl_create()
        try:
            cuts = np.array([cutoff], dtype=np.float64)
            if self.lib.kref_nl_build(h, n, coords, infl, 1, cuts, need):
                raise RuntimeError("neighbor build failed")
            tot = self.lib.kref_nl_total(h, 0)
            idx = np.empty(max(int(tot), 1), np.int32)
            self.lib.kref_nl_export(h, 0, counts, offsets, idx)
            return counts, offsets, idx[: int(tot)]
        finally:
            self.lib.kref_nl_free(h)

    # ------------------------------------------------------------------ descriptor
    def _pack(self, rcut, fams):
        rcut = np.ascontiguousarray(rcut, dtype=np.float64)
        S = rcut.shape[0]
        kinds = np.array([k for k, _ in fams], np.int32)
        rows = np.array([v.shape[0] for _, v in fams], np.int32)
        vals = np.concatenate([np.ascontiguousarray(v, np.float64).reshape(-1) for _, v in fams])
        D = int(rows.sum())
        return rcut, S, kinds, rows, vals, D

    def _ref_desc(self, rcut, fams):
        rcut, S, kinds, rows, vals, D = self._pack(rcut, fams)
        d = self.lib.kref_desc_create()
        self.lib.kref_desc_set_cutoff(d, b"cos", S, rcut.reshape(-1))
        names = {1: b"g1", 2: b"g2", 3: b"g3", 4: b"g4", 5: b"g5"}
        for k, v in fams:
            v = np.ascontiguousarray(v, np.float64)
            self.lib.kref_desc_add(d, names[k], v.reshape(-1), v.shape[0], v.shape[1])
        return d, D

    def symfun_atoms_from_file(self, path, atoms, coords, species, offsets, indices):
        """The reference's own parser + kernel: Descriptor::read_parameter_file (sym_fn.cpp:108-387) on
        a `descriptor.params` file, then generate_one_atom for every atom in `atoms`.
        -> (zeta [m, D], list of [D, n+1, 3] blocks, mean [D] | None, stdev [D] | None)."""
        assert self.kind == "reference"
        coords = np.ascontiguousarray(coords, np.float64).reshape(-1, 3)
        species = np.ascontiguousarray(species, np.int32)
        d = self.lib.kref_desc_create()
        try:
            err = self.lib.kref_desc_read_file(d, os.fsencode(path))
            if err:
                raise RuntimeError(f"read_parameter_file failed ({err}) on {path}")
            D = self.lib.kref_desc_size(d)
            mean = np.zeros(D); stdev = np.zeros(D)
            has = 0
            for q in range(D):
                m, sd = C.c_double(0.0), C.c_double(0.0)
                has = self.lib.kref_desc_mean_std(d, q, C.byref(m), C.byref(sd))
                mean[q], stdev[q] = m.value, sd.value
            zs, blocks = [], []
            for i in atoms:
                nb = np.ascontiguousarray(indices[offsets[i]:offsets[i + 1]], np.int32)
                n = nb.shape[0]
                z = np.zeros(D); g = np.zeros((D, n + 1, 3))
                self.lib.kref_desc_one_atom(d, int(i), coords, species,
                                            nb if n else np.zeros(1, np.int32), n, z, g.ctypes.data, 1)
                zs.append(z); blocks.append(g)
        finally:
            self.lib.kref_desc_free(d)
        return np.array(zs), blocks, (mean if has else None), (stdev if has else None)

    def symfun_atom(self, rcut, fams, i, coords, species, neigh, grad=True):
        """generate_one_atom: -> (zeta f64[D], dzeta f64[D, n+1, 3] | None)."""
        coords = np.ascontiguousarray(coords, np.float64).reshape(-1, 3)
        species = np.ascontiguousarray(species, np.int32)
        neigh = np.ascontiguousarray(neigh, np.int32)
        n = neigh.shape[0]
        rcut, S, kinds, rows, vals, D = self._pack(rcut, fams)
        z = np.zeros(D)
        g = np.zeros((D, n + 1, 3)) if grad else None
        if self.kind == "port":
            self.lib.kbo_symfun_atom(S, rcut.reshape(-1), len(fams), kinds, rows, vals, int(i),
                                     coords, species, neigh, n, z,
                                     g.ctypes.data if grad else None, int(grad))
        else:
            d, _ = self._ref_desc(rcut, fams)
            try:
                self.lib.kref_desc_one_atom(d, int(i), coords, species, neigh, n, z,
                                            g.ctypes.data if grad else None, int(grad))
            finally:
                self.lib.kref_desc_free(d)
        return z, g

    def symfun_range(self, rcut, fams, first, last, coords, species, offsets, indices,
                     grad=True, keep_grad=True):
        """Atoms [first,last) of a CSR list -> (zeta f64[m, D], packed dzeta f64 | None).
        Packed layout: atom i's [D][n_i+1][3] block starts at
        3*D*(offsets[i]-offsets[first] + (i-first))."""
        coords = np.ascontiguousarray(coords, np.float64).reshape(-1, 3)
        species = np.ascontiguousarray(species, np.int32)
        offsets = np.ascontiguousarray(offsets, np.int64)
        indices = np.ascontiguousarray(indices, np.int32)
        if indices.size == 0:
            indices = np.zeros(1, np.int32)
        rcut, S, kinds, rows, vals, D = self._pack(rcut, fams)
        m = last - first
        z = np.zeros((m, D))
        counts = np.diff(offsets)
        maxn = int(counts[first:last].max()) if m else 0
        scratch = np.zeros(3 * D * (maxn + 1) + 1)
        g = None
        if grad and keep_grad:
            g = np.zeros(3 * D * int(offsets[last] - offsets[first] + m))
        if self.kind == "port":
            self.lib.kbo_symfun_range(S, rcut.reshape(-1), len(fams), kinds, rows, vals, D,
                                      first, last, coords, species, offsets, indices, z,
                                      g.ctypes.data if g is not None else None, scratch, int(grad))
        else:
            d, _ = self._ref_desc(rcut, fams)
            try:
                if g is None:
                    self.lib.kref_desc_range(d, first, last, coords, species, offsets, indices,
                                             z, scratch, int(grad))
                else:
                    for i in range(first, last):
                        n = int(counts[i])
                        o = 3 * D * int(offsets[i] - offsets[first] + (i - first))
                        blk = np.zeros(3 * D * (n + 1))
                        zi = np.zeros(D)
                        nb = np.ascontiguousarray(indices[offsets[i]:offsets[i + 1]])
                        if nb.size == 0:
                            nb = np.zeros(1, np.int32)
                        self.lib.kref_desc_one_atom(d, i, coords, species, nb, n, zi,
                                                    blk.ctypes.data, 1)
                        z[i - first] = zi
                        g[o:o + blk.size] = blk
            finally:
                self.lib.kref_desc_free(d)
        return z, g


# ---------------------------------------------------------------------- numpy restatements
def canonical(offsets, indices):
    """Sort every atom's neighbor slice ascending ("canonical sorting")."""
    out = indices.copy()
    for a in range(len(offsets) - 1):
        out[offsets[a]:offsets[a + 1]].sort()
    return out


def dense_fold(dz_blocks, neigh_lists, image, ncontrib):
    """sym_fn.py:163-174: per-atom [D][n+1][3] blocks -> dense (N, D, 3N) through `image`."""
    D = dz_blocks[0].shape[0]
    out = np.zeros((ncontrib, D, ncontrib, 3))
    for i in range(ncontrib):
        ids = np.concatenate((neigh_lists[i], [i]))
        for slot, a in enumerate(ids):
            out[i, :, image[a], :] += dz_blocks[i][:, slot, :]
    return out.reshape(ncontrib, D, 3 * ncontrib)


def stress_fold(dz_blocks, neigh_lists, coords, ncontrib):
    """sym_fn.py:176-185: Voigt order 11,22,33,23,31,12."""
    D = dz_blocks[0].shape[0]
    out = np.zeros((ncontrib, D, 6))
    for i in range(ncontrib):
        ids = np.concatenate((neigh_lists[i], [i]))
        for slot, a in enumerate(ids):
            g = dz_blocks[i][:, slot, :]
            x = coords[a]
            out[i, :, 0] += g[:, 0] * x[0]
            out[i, :, 1] += g[:, 1] * x[1]
            out[i, :, 2] += g[:, 2] * x[2]
            out[i, :, 3] += g[:, 1] * x[2]
            out[i, :, 4] += g[:, 2] * x[0]
            out[i, :, 5] += g[:, 0] * x[1]
    return out


def forces_from_dense(dedz, dense):
    """calculator_torch.py:266-269: F = -tensordot(dE/dzeta[N,D], dzetadr[N,D,3N])."""
    return -np.tensordot(dedz, dense, axes=([0, 1], [0, 1]))


def forces_sparse(dedz, dz_blocks, neigh_lists, image, ncontrib):
    """Same contraction without the dense intermediate (used at sizes where (N,D,3N) is
    impossible): F[image[a]] -= sum_d dE/dzeta[i,d] * dz_i[d, slot(a), :]."""
    F = np.zeros((ncontrib, 3))
    for i in range(len(dz_blocks)):
        ids = np.concatenate((neigh_lists[i], [i]))
        contrib = np.tensordot(dedz[i], dz_blocks[i], axes=([0], [0]))  # (n+1, 3)
        np.subtract.at(F, image[ids], contrib)
    return F.reshape(-1)
