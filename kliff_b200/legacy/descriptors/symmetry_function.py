"""`SymmetryFunction` with the interface of
kliff/legacy/descriptors/symmetry_function/sym_fn.py:17-396, computed by the sm_100a kernels.

`transform(conf, fit_forces, fit_stress)` returns the reference's dense numpy arrays (zeta
(N, D); dzetadr_forces (N, D, 3N); dzetadr_stress (N, D, 6)) — only sensible for small cells, as
in the reference.  `transform_packed(configs)` is the scalable form: any number of
configurations (or one huge supercell) in one batched GPU pass, result left on the device in the
packed layout of kliff_b200/packed.py.
"""
import os
from collections import OrderedDict

import numpy as np
import torch

from ... import ops
from ...neighbor import NeighborList
from ...packed import PackedFingerprints
from ...utils import default_device, map_species
from .descriptor import (Descriptor, generate_full_cutoff, generate_species_code,
                         generate_unique_cutoff_pairs)
from .hyperparams import get_set30, get_set51


class SymmetryFunctionError(Exception):
    def __init__(self, msg):
        super().__init__(msg)
        self.msg = msg


_DENSE_LIMIT_BYTES = 8 << 30
# configurations up to this many atoms go through the batched all-pairs neighbor kernels (one CTA
# per configuration); larger ones through the cell-list build
_BATCH_MAX_ATOMS = 2048


class SymmetryFunction(Descriptor):
    def __init__(self, cut_dists, cut_name, hyperparams, normalize=True, dtype=np.float32):
        super().__init__(cut_dists, cut_name, hyperparams, normalize, dtype)
        self._desc = OrderedDict()
        self._set_cutoff()
        self._set_hyperparams()
        self.size = self.get_size()
        self._params = ops.build_params(self.hyperparams, self._rcut)
        self.kernel_mode = 0  # KB_SYMFUN_AUTO
        # storage type of the packed dG/dr blocks kept by generate_fingerprints: None = follow
        # `dtype` (the reference casts its fingerprints to `dtype`, float32 by default, after
        # computing them in float64, descriptor.py:197-205); the arithmetic is fp64 either way
        self.gradient_dtype = None

    # ------------------------------------------------------------------ setup (sym_fn.py:214-283)
    def _set_cutoff(self):
        supported = ["cos"]
        if self.cut_name is None:
            self.cut_name = supported[0]
        if self.cut_name not in supported:
            spd = ['"{}", '.format(s) for s in supported]
            raise SymmetryFunctionError(
                'Cutoff "{}" not supported by this descriptor. Use {}.'.format(self.cut_name, spd))
        self.cutoff = generate_full_cutoff(self.cut_dists)
        self.species_code = generate_species_code(self.cut_dists)
        ns = len(self.species_code)
        rcut = np.zeros([ns, ns], dtype=np.double)
        for si, i in self.species_code.items():
            for sj, j in self.species_code.items():
                rcut[i][j] = self.cutoff[si + "-" + sj]
        self._rcut = rcut

    def _set_hyperparams(self):
        if isinstance(self.hyperparams, str):
            name = self.hyperparams.lower()
            if name == "set51":
                self.hyperparams = get_set51()
            elif name == "set30":
                self.hyperparams = get_set30()
            else:
                raise SymmetryFunctionError('hyperparams "{}" unrecognized.'.format(name))
        if not isinstance(self.hyperparams, OrderedDict):
            self.hyperparams = OrderedDict(self.hyperparams)
        ncols = {"g2": ("eta", "Rs"), "g3": ("kappa",), "g4": ("zeta", "lambda", "eta"),
                 "g5": ("zeta", "lambda", "eta")}
        for key, values in self.hyperparams.items():
            name = key.lower()
            if name not in ["g1", "g2", "g3", "g4", "g5"]:
                raise SymmetryFunctionError('Symmetry function "{}" unrecognized.'.format(key))
            if name == "g1":
                params = np.zeros([1, 1], dtype=np.double)
            else:
                params = np.array([[line[c] for c in ncols[name]] for line in values], dtype=np.double)
            self._desc[name] = params

    # ------------------------------------------------------------------ GPU passes
    def _config_part(self, conf, device):
        infl = max(self.cutoff.values())
        nei = NeighborList(conf, infl, padding_need_neigh=False, device=device)
        d = nei.device_data
        n = d["n_contrib"]
        try:
            code_cb = map_species(conf.species, self.species_code)
        except KeyError as e:
            raise SymmetryFunctionError(f"species {e} not in the cutoff dictionary")
        code_cb = torch.as_tensor(code_cb, device=device)
        code = code_cb[d["image"].long()]
        return dict(coords=d["coords"], code=code, image_local=d["image"], counts=d["counts"],
                    indices_local=d["indices"], n_contrib=n, maxn=d["maxn"]), nei

    def transform_packed(self, configs, grad=True, device=None, gradient_dtype=np.float64, zeta_host=None):
        """Descriptors (+ packed dG/dr) of a list of configurations in one batched pass.
        `gradient_dtype` is the storage type of the blocks (fp64 unless asked otherwise;
        `generate_fingerprints` passes the descriptor's `dtype`, as the reference casts there).
        `zeta_host`: optional pinned host array [N, D] the raw descriptors are streamed into while
        the kernels run (`packed.host_ready` is the event to wait for)."""
        device = device if device is not None else default_device()
        if not isinstance(configs, (list, tuple)):
            configs = [configs]
        if len(configs) == 0:  # a rank whose share of a sharded data set is empty
            packed = PackedFingerprints(len(self), device)
            packed.zeta = torch.zeros((0, len(self)), dtype=torch.float64, device=device)
            packed.host_ready = None
            return packed
        if max(c.get_num_atoms() for c in configs) <= _BATCH_MAX_ATOMS:
            # many small cells: one CTA per configuration, 2 synchronisations for the whole list
            try:
                packed = PackedFingerprints.from_configs(len(self), device, configs, self.species_code,
                                                         max(self.cutoff.values()))
            except KeyError as e:
                raise SymmetryFunctionError(f"species {e} not in the cutoff dictionary")
        else:
            # large cells: cell-list neighbor build per configuration
            parts = [self._config_part(c, device)[0] for c in configs]
            packed = PackedFingerprints.from_parts(len(self), device, parts)
        packed.compute(self._params, grad=grad, mode=self.kernel_mode,
                       dz_dtype=torch.float32 if gradient_dtype in (np.float32, torch.float32) else torch.float64,
                       zeta_host=zeta_host)
        return packed

    def transform(self, conf, fit_forces=False, fit_stress=False):
        """sym_fn.py:98-212."""
        n = conf.get_num_atoms()
        D = len(self)
        grad = fit_forces or fit_stress
        if fit_forces and n * D * 3 * n * 8 > _DENSE_LIMIT_BYTES:
            raise SymmetryFunctionError(
                f"the dense (N, D, 3N) gradient of a {n}-atom configuration needs "
                f"{n * D * 3 * n * 8 / 2**30:.0f} GiB; use transform_packed()")
        packed = self.transform_packed([conf], grad=grad)
        zeta = packed.zeta.cpu().numpy()
        dzf = dzs = None
        if grad:
            dense, stress = packed.dense_config(0, fit_forces, fit_stress)
            if fit_forces:
                dzf = dense.cpu().numpy()
            if fit_stress:
                dzs = stress.cpu().numpy()
        return zeta, dzf, dzs

    # ------------------------------------------------------------------ misc API
    def get_size(self):
        return len(self)

    def get_hyperparams(self):
        return self._desc

    def __len__(self):
        return sum(len(v) for v in self._desc.values())

    def __getstate__(self):
        state = self.__dict__.copy()
        state["_params"] = None              # ctypes struct: rebuilt on load
        state["packed_fingerprints"] = None  # device memory does not pickle
        state["last_examples"] = None
        return state

    def __setstate__(self, state):
        self.__dict__.update(state)
        self._params = ops.build_params(self.hyperparams, self._rcut)

    def write_kim_params(self, path, fname="descriptor.params"):
        """Text parameter file of sym_fn.py:285-384 (same layout, read back by the reference's
        sym_fn.cpp:108-387)."""
        fmt = "{:.15e} " if self.dtype == np.float64 else "{:.7e} "
        bar = "#" + "=" * 80 + "\n"
        with open(os.path.join(path, fname), "w") as fout:
            fout.write(bar + "# Descriptor parameters file generated by KLIFF.\n" + bar + "\n")
            cutname, rcut = self.get_cutoff()
            pairs = generate_unique_cutoff_pairs(rcut)
            species = generate_species_code(rcut)
            fout.write("{}  # cutoff type\n\n".format(cutname))
            fout.write("{}  # number of species\n\n".format(len(species)))
            fout.write("# species 1    species 2    cutoff\n")
            for key, value in pairs.items():
                s1, s2 = key.split("-")
                fout.write(("{}  {}  " + fmt + "\n").format(s1, s2, value))
            fout.write("\n" + bar + "# symmetry functions\n" + bar + "\n")
            desc = self.get_hyperparams()
            fout.write("{}  # number of symmetry functions types\n\n".format(len(desc)))
            fout.write("# sym_function    rows    cols\n")
            labels = {"g2": "    # eta  Rs\n", "g3": "    # kappa\n",
                      "g4": "    # zeta  lambda  eta\n", "g5": "    # zeta  lambda  eta\n"}
            for name, values in desc.items():
                if name == "g1":
                    fout.write("g1\n\n")
                    continue
                fout.write("{}    {}    {}\n".format(name, len(values), len(values[0])))
                for val in values:
                    fout.write((fmt * len(val)).format(*val))
                    fout.write(labels[name])
                fout.write("\n")
            fout.write(bar + "# Preprocessing data to center and normalize\n" + bar)
            mean, stdev = self.get_mean(), self.get_stdev()
            if mean is None and stdev is None:
                fout.write("center_and_normalize  False\n")
            else:
                fout.write("center_and_normalize  True\n\n")
                fout.write("{}   # descriptor size\n".format(self.get_size()))
                fout.write("# mean\n")
                for v in mean:
                    fout.write((fmt + "\n").format(v))
                fout.write("\n# standard deviation\n")
                for v in stdev:
                    fout.write((fmt + "\n").format(v))
                fout.write("\n")
