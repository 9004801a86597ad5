"""`CalculatorTorch` / `CalculatorTorchSeparateSpecies` with the interface of
kliff/legacy/calculators/calculator_torch.py:16-525.

What changes underneath:
  * `create()` runs the descriptor kernels once for all configurations and keeps the packed
    result on the GPU (kliff_b200/packed.py); the fingerprints pickle is still written in the
    reference's per-example format (zeta, energy, forces, ...), so `reuse=True` works.
  * `compute(batch)` evaluates the MLP with torch (cuBLAS) on the resident, normalised zeta and
    replaces the per-configuration dense `-tensordot(dE/dzeta, dzetadr_forces)`
    (calculator_torch.py:207-213, 266-269) by ONE launch of the force-scatter kernel per batch,
    differentiable through dE/dzeta (create_graph=True, :200-202).  The return value keeps the
    reference's type: dict of lists of per-configuration tensors.
  * `gpu=None` (the README call `CalculatorTorch(model)`) selects the process's CUDA device:
    there is no CPU path.
"""
import numpy as np
import torch
from torch.utils.data import DataLoader

from ...dataset.dataset_torch import FingerprintsDataset, fingerprints_collate_fn
from ...packed import PackedFingerprints, PackedGradient
from ...utils import default_device, pickle_load, to_path


class CalculatorTorchError(Exception):
    def __init__(self, msg):
        super().__init__(msg)
        self.msg = msg


def _dist_active():
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        return dist
    return None


def _rank_suffix(filename, rank):
    """fingerprints.pkl -> fingerprints.rank3.pkl (ranks of one node share the working directory)."""
    path = to_path(filename)
    return path.with_name(f"{path.stem}.rank{rank}{path.suffix}")


def _members_file(rank_file):
    """fingerprints.rank3.pkl -> fingerprints.rank3.members.npy: the global indices a size-balanced
    create() gave this rank, read back by `reuse=True`."""
    path = to_path(rank_file)
    return path.with_name(path.stem + ".members.npy")


def _get_device(gpu):
    """calculator_torch.py:582-594, minus the CPU branch: `gpu=False` / `None` in the reference mean
    "run on the CPU"; this package has no CPU path, so they select the current CUDA device (said once)."""
    if gpu is None or isinstance(gpu, bool):
        if gpu is False and not _get_device.warned:
            _get_device.warned = True
            import warnings
            warnings.warn("kliff_b200 has no CPU path: CalculatorTorch(gpu=False) runs on the current CUDA "
                          "device", stacklevel=3)
        return default_device()
    return torch.device("cuda", int(gpu))


_get_device.warned = False


def _on_model_device(method):
    """Run a calculator method with the model's device as the current CUDA device, so that every
    library call (streams, scratch, default_device()) lands on the device `gpu=<index>` selected."""
    import functools

    @functools.wraps(method)
    def wrapper(self, *args, **kwargs):
        dev = self.model.device
        if dev.type != "cuda" or dev.index is None or dev.index == torch.cuda.current_device():
            return method(self, *args, **kwargs)
        with torch.cuda.device(dev):
            return method(self, *args, **kwargs)
    return wrapper


def _runs(indices):
    """Split a list of configuration indices into maximal runs of consecutive integers."""
    runs, start, prev = [], None, None
    for pos, i in enumerate(indices):
        if start is None:
            start, prev, p0 = i, i, pos
        elif i == prev + 1:
            prev = i
        else:
            runs.append((start, prev + 1, p0))
            start, prev, p0 = i, i, pos
    if start is not None:
        runs.append((start, prev + 1, p0))
    return runs


class CalculatorTorch:
    implemented_property = ["energy", "forces", "stress"]

    def __init__(self, model, gpu=None):
        device = _get_device(gpu)
        self._model = model.to(device)
        self.dtype = self.model.descriptor.dtype
        self.fingerprints_path = None
        self.use_energy = None
        self.use_forces = None
        self.use_stress = None
        self.results = dict([(i, None) for i in self.implemented_property])
        self._packed = None
        self._ref = None
        self._shard = None
        self._shard_members = None

    # ------------------------------------------------------------------ create (:41-126)
    @_on_model_device
    def create(self, configs, use_energy=True, use_forces=True, use_stress=False,
               fingerprints_filename="fingerprints.pkl", fingerprints_mean_stdev_filename=None,
               reuse=False, use_welford_method=False, nprocs=1, shard_configs=None):
        """`shard_configs` (default: on whenever torch.distributed runs with more than one rank and
        fingerprints are generated, not reused): every rank generates and keeps only its share of the
        configurations — the multi-GPU form of the reference's `parmap1` over configurations
        (descriptor.py:234-236).  Mean / stdev are all-reduced, so every rank normalises with the
        global statistics; `Loss` then takes each rank's members of every batch.

        Which configurations a rank keeps: `True` / `None` / "auto" — r, r + world, ... for a uniform
        set, and the size-balanced assignment of `kliff_b200.parallel.shard_members` (cost = atoms · n²,
        one member per rank in every run of `world` configurations) when the estimated costs differ by
        more than 2×; "round_robin" / "balanced" force one of the two; `False` keeps everything on
        every rank."""
        self.configs = configs
        self.use_energy = use_energy
        self.use_forces = use_forces
        self.use_stress = use_stress
        if not isinstance(configs, (list, tuple)):
            configs = [configs]
        desc = self.model.descriptor
        self._shard = None
        dist = _dist_active()
        self._shard_members = None
        if dist is not None and not reuse and (shard_configs is None or shard_configs):
            r, w = dist.get_rank(), dist.get_world_size()
            self._shard = (r, w, len(configs))
            mine = self._choose_members(configs, r, w, shard_configs)
            configs = [configs[i] for i in mine]
            fingerprints_filename = _rank_suffix(fingerprints_filename, r)
            sidecar = _members_file(fingerprints_filename)
            if self._shard_members is not None:
                np.save(sidecar, self._shard_members)
            elif sidecar.exists():
                sidecar.unlink()  # a balanced create() of an earlier run must not be read back with these files
            if fingerprints_mean_stdev_filename is None:
                fingerprints_mean_stdev_filename = "fingerprints_mean_and_stdev.pkl"
            if r != 0:  # rank 0 writes the file the user asked for; the others keep private copies
                fingerprints_mean_stdev_filename = _rank_suffix(fingerprints_mean_stdev_filename, r)
        if reuse:
            self.fingerprints_path = to_path(fingerprints_filename)
            if dist is not None and (shard_configs is None or shard_configs):
                # files of an earlier sharded create(): every rank reads its own share back
                r, w = dist.get_rank(), dist.get_world_size()
                mine = _rank_suffix(fingerprints_filename, r)
                if mine.exists():
                    self.fingerprints_path = mine
                    self._shard = (r, w, len(configs))
                    if _members_file(mine).exists():  # written by a size-balanced create()
                        self._shard_members = np.load(_members_file(mine)).astype(np.int64)
                    ms = fingerprints_mean_stdev_filename or "fingerprints_mean_and_stdev.pkl"
                    if r != 0 and _rank_suffix(ms, r).exists():
                        fingerprints_mean_stdev_filename = _rank_suffix(ms, r)
                    elif fingerprints_mean_stdev_filename is None and to_path(ms).exists():
                        fingerprints_mean_stdev_filename = ms
            if not self.fingerprints_path.exists():
                raise CalculatorTorchError(
                    f"You specified `reuse=True` to reuse the fingerprints stored in "
                    f"`{self.fingerprints_path}` This file does not exists.")
            if desc.normalize:
                path = None if fingerprints_mean_stdev_filename is None else to_path(fingerprints_mean_stdev_filename)
                if path is None or not path.exists():
                    raise CalculatorTorchError(
                        f"You specified `reuse=True` to reuse the fingerprints. The "
                        f"mean and stdev file of the fingerprints `{path}` does not exists.")
                desc.load_state_dict(pickle_load(path))
            self._packed = None
        else:
            self.fingerprints_path = desc.generate_fingerprints(
                configs, use_forces, use_stress, fingerprints_filename,
                fingerprints_mean_stdev_filename, use_welford_method, nprocs,
                **({"reduce_over": dist} if self._shard is not None else {}))
            self._packed = desc.packed_fingerprints
        examples = None if reuse else getattr(desc, "last_examples", None)
        self.fingerprints_dataset = FingerprintsDataset(examples if examples is not None
                                                        else self.fingerprints_path)
        desc.last_examples = None
        if self._packed is None:
            self._packed = self._packed_from_samples(self.fingerprints_dataset.fp)
        self._upload_references()

    def _choose_members(self, configs, r, w, mode):
        """Global indices of this rank's configurations; leaves `_shard_members` at None for the
        strided assignment (index arithmetic) and sets it to the ascending index array otherwise."""
        from ...parallel import config_costs, shard_members
        mode = "auto" if mode is None or mode is True else mode
        if mode == "round_robin":
            return range(r, len(configs), w)
        cut = getattr(self.model.descriptor, "cutoff", None)
        rc = max(cut.values()) if isinstance(cut, dict) and cut else 5.0
        if mode == "auto" and len(configs) > 2048:
            # decide on a strided sample first (0.5 ms): a uniform set pays nothing more
            probe = config_costs(configs[::len(configs) // 1024], rc)
            if probe.min() > 0.0 and probe.max() / probe.min() <= 2.0:
                return range(r, len(configs), w)
        parts = shard_members(config_costs(configs, rc), w, mode)
        if all(np.array_equal(p, np.arange(k, len(configs), w)) for k, p in enumerate(parts)):
            return range(r, len(configs), w)
        self._shard_members = parts[r]
        return [int(i) for i in parts[r]]

    def _shard_below(self, lo):
        """How many of this rank's configurations have a global index below `lo` (= the position in
        the local list where global batch boundary `lo` falls)."""
        r, w, _ = self._shard
        members = getattr(self, "_shard_members", None)
        if members is None:
            return max(0, -((r - lo) // w))  # ceil((lo - r) / w) members of r, r + w, ...
        return int(np.searchsorted(members, lo))

    def _packed_from_samples(self, samples):
        """Rebuild the device store from a loaded pickle when its examples carry PackedGradient
        objects; dense (reference-made) examples are served by the dense branch of compute()."""
        if not samples or not self.use_forces:
            return None
        if not all(isinstance(s.get("dzetadr_forces"), PackedGradient) for s in samples):
            return None
        dev = self.model.device
        packed = PackedFingerprints.from_packed_gradients(
            self.model.descriptor.get_size(), dev, [s["dzetadr_forces"] for s in samples],
            dz_dtype=torch.float32 if self.dtype == np.float32 else torch.float64)
        zeta = np.concatenate([np.asarray(s["zeta"], np.float64) for s in samples])
        packed.zeta_normalized = torch.as_tensor(zeta, device=dev)
        packed.scale = None  # gradients on disk are already divided by stdev
        return packed

    def _upload_references(self):
        """Reference energies / forces / weights of every configuration, resident on the GPU for
        the vectorised loss (kliff_b200/legacy/loss.py)."""
        self._ref = None
        if self._packed is None:
            return
        fp = self.fingerprints_dataset.fp
        dev, dt = self.model.device, self.model.dtype
        try:
            e = np.array([float(s["energy"]) for s in fp]) if self.use_energy else np.zeros(len(fp))
            f = (np.concatenate([np.asarray(s["forces"], np.float64).reshape(-1) for s in fp])
                 if (self.use_forces and len(fp)) else np.zeros(0))
        except Exception:
            return
        self._ref = dict(energy=torch.as_tensor(e, device=dev).to(dt),
                         forces=torch.as_tensor(f, device=dev).to(dt))

    # ------------------------------------------------------------------ small API (:128-159)
    def get_fingerprints(self):
        return self.fingerprints_dataset.fp

    def get_compute_arguments(self, batch_size=1):
        return DataLoader(dataset=self.fingerprints_dataset, batch_size=batch_size,
                          collate_fn=fingerprints_collate_fn)

    def set_fingerprints(self, fingerprints):
        self.fingerprints_dataset.fp = fingerprints

    def fit(self):
        self.model.fit(self.fingerprints_path)

    # ------------------------------------------------------------------ compute (:161-228)
    def _model_energy(self, zeta, cfg_indices):
        """Per-atom energies [N, 1] for stacked zeta; overridden for per-species models."""
        return self.model(zeta)

    def _packed_usable(self, batch):
        if self._packed is None:
            return False
        P = self._packed
        for s in batch:
            i = s.get("index")
            if i is None or i < 0 or i >= P.n_configs:
                return False
            lo, hi = P.config_slice(i)
            if hi - lo != len(s["zeta"]):
                return False
        return True

    @_on_model_device
    def compute_packed(self, cfg_indices, want_forces=None):
        """Vectorised core: energies [B] and flat forces [3 * atoms in batch] (both in the model
        dtype, differentiable w.r.t. the model parameters) for configurations `cfg_indices`,
        plus the per-configuration atom counts."""
        P = self._packed
        dev, dt = self.model.device, self.model.dtype
        want_forces = self.use_forces if want_forces is None else want_forces
        runs = _runs(list(cfg_indices))
        spans = [(int(P.cfg_ptr[a]), int(P.cfg_ptr[b])) for a, b, _ in runs]
        zparts = [P.zeta_normalized[lo:hi] for lo, hi in spans]
        zeta = (zparts[0] if len(zparts) == 1 else torch.cat(zparts)).to(dt)
        if want_forces:
            zeta = zeta.detach().requires_grad_(True)
        cfg_indices = list(cfg_indices)
        natoms = (P.cfg_ptr[1:] - P.cfg_ptr[:-1])[np.asarray(cfg_indices, dtype=np.int64)]
        if getattr(P, "_natoms_dev", None) is None:
            P._natoms_dev = torch.as_tensor(np.diff(P.cfg_ptr), device=dev)
        if len(runs) == 1:
            reps = P._natoms_dev[runs[0][0]:runs[0][1]]
        else:
            reps = P._natoms_dev.index_select(0, torch.as_tensor(cfg_indices, device=dev))
        cfg_of_atom = torch.repeat_interleave(torch.arange(len(natoms), device=dev), reps,
                                              output_size=zeta.shape[0])
        e_atom = self._model_energy(zeta, cfg_indices).reshape(-1)
        energy = torch.zeros(len(natoms), dtype=dt, device=dev).index_add(0, cfg_of_atom, e_atom)
        forces = None
        if want_forces:
            dedz = torch.autograd.grad(e_atom.sum(), zeta, create_graph=True)[0]
            pieces, pos = [], 0
            for lo, hi in spans:
                f = P.forces(dedz[pos:pos + (hi - lo)].to(torch.float64), lo, hi)
                pieces.append(f)
                pos += hi - lo
            forces = (pieces[0] if len(pieces) == 1 else torch.cat(pieces)).to(dt)
        return energy, forces, natoms

    @_on_model_device
    def compute(self, batch):
        grad = self.use_forces or self.use_stress
        if self._packed_usable(batch) and not self.use_stress:
            idx = [s["index"] for s in batch]
            energy, forces, natoms = self.compute_packed(idx)
            energy_config = list(energy.unbind(0))
            forces_config = None
            if self.use_forces:
                forces_config = list(torch.split(forces, [3 * int(n) for n in natoms]))
            stress_config = None
        else:
            energy_config, forces_config, stress_config = self._compute_dense(batch, grad)
        self.results["energy"] = energy_config
        self.results["forces"] = forces_config
        self.results["stress"] = stress_config
        return {"energy": energy_config, "forces": forces_config, "stress": stress_config}

    def _compute_dense(self, batch, grad):
        """Compatibility branch for fingerprints that carry dense (N, D, 3N) arrays (a cache
        written by the reference, or `fingerprints_format = "dense"`): the reference's own
        formulation, evaluated by torch on the GPU."""
        device = self.model.device
        zeta_config = [s["zeta"] for s in batch]
        zeta = torch.cat(zeta_config, dim=0).to(device)
        if grad:
            zeta.requires_grad_(True)
        e_atom = self._model_energy(zeta, [s.get("index", -1) for s in batch])
        natoms = [len(z) for z in zeta_config]
        energy_config = [e.sum() for e in torch.split(e_atom, natoms)]
        forces_config = [] if self.use_forces else None
        stress_config = [] if self.use_stress else None
        if grad:
            dedz_all = torch.autograd.grad(e_atom.sum(), zeta, create_graph=True)[0]
            for s, dedz in zip(batch, torch.split(dedz_all, natoms)):
                if self.use_forces:
                    dzf = s.get("dzetadr_forces")
                    if dzf is None:
                        raise CalculatorTorchError(
                            "the fingerprints hold no `dzetadr_forces` (they were written with "
                            "`store_gradients_on_disk = False`): forces cannot be computed from this file; "
                            "re-create the fingerprints (reuse=False) or store the gradients")
                    if isinstance(dzf, PackedGradient):
                        dzf = torch.from_numpy(dzf.to_dense())
                    forces_config.append(self._compute_forces(dedz, dzf.to(device)))
                if self.use_stress:
                    vol = s["volume"].to(device) if torch.is_tensor(s["volume"]) else s["volume"]
                    stress_config.append(self._compute_stress(dedz, s["dzetadr_stress"].to(device), vol))
        return energy_config, forces_config, stress_config

    @staticmethod
    def _compute_forces(denergy_dzeta, dzetadr):
        return -torch.tensordot(denergy_dzeta, dzetadr, dims=([0, 1], [0, 1]))

    @staticmethod
    def _compute_stress(denergy_dzeta, dzetadr, volume):
        return torch.tensordot(denergy_dzeta, dzetadr, dims=([0, 1], [0, 1])) / volume

    # ------------------------------------------------------------------ model access (:230-353)
    @property
    def model(self):
        return self._model

    def save_model(self, epoch, force_save=False):
        m = self.model
        path = to_path(m.save_prefix).joinpath(f"model_epoch{epoch}.pkl")
        if force_save or (epoch >= m.save_start and (epoch - m.save_start) % m.save_frequency == 0):
            m.save(path)

    def get_energy(self, batch):
        return self.results["energy"]

    def get_forces(self, batch):
        return self.results["forces"]

    def get_stress(self, batch):
        return self.results["stress"]

    def get_size_opt_params(self):
        sizes, nelements = [], []
        for p in self.model.parameters():
            sizes.append(p.size())
            nelements.append(int(np.prod(p.size())))
        return sizes, nelements, sum(nelements)

    def get_num_opt_params(self):
        return self.get_size_opt_params()[-1]

    def get_opt_params(self, flat=True):
        params = []
        for p in self.model.parameters():
            if flat:
                params = np.append(params, p.data.cpu().numpy().flatten())
            else:
                params.append(p)
        return params

    def update_model_params(self, parameters):
        sizes, nelems, _ = self.get_size_opt_params()
        idx = np.append(0, np.cumsum(nelems)).astype(int)
        for ii, p in enumerate(self.model.parameters()):
            chunk = np.asarray(parameters[idx[ii]:idx[ii + 1]]).reshape(tuple(sizes[ii]))
            p.data = torch.as_tensor(chunk, dtype=p.dtype, device=p.device)


class _ModelWrapper(torch.nn.Module):
    """calculator_torch.py:499-525."""

    def __init__(self, models):
        super().__init__()
        self._models = torch.nn.ModuleDict(models)
        self._descriptor = list(models.values())[0].descriptor
        self._dtype = list(models.values())[0].dtype

    device = property(lambda self: next(self.parameters()).device)
    descriptor = property(lambda self: self._descriptor)
    dtype = property(lambda self: self._dtype)


class CalculatorTorchSeparateSpecies(CalculatorTorch):
    """One MLP per species (calculator_torch.py:356-496).  The reference buckets atoms in a
    Python loop and differentiates once per configuration (:404-449); here the species of every
    resident atom is a device tensor and one autograd pass covers the batch."""

    def __init__(self, models, gpu=None):
        device = _get_device(gpu)
        self.models = models
        self.dtype = None
        for s, m in self.models.items():
            m.to(device)
            if self.dtype is None:
                self.dtype = m.descriptor.dtype
            elif self.dtype != m.descriptor.dtype:
                raise CalculatorTorchError("inconsistent `dtype` from descriptors.")
        self._model = _ModelWrapper(models)
        self.fingerprints_path = None
        self.use_energy = self.use_forces = self.use_stress = None
        self.results = dict([(i, None) for i in self.implemented_property])
        self._packed = None
        self._ref = None
        self._species_of_cfg = None

    def create(self, configs, *args, **kwargs):
        super().create(configs, *args, **kwargs)
        names = list(self.models.keys())
        self._species_names = names
        per_cfg = []
        for s in self.fingerprints_dataset.fp:
            sp = s["configuration"].species
            for x in sp:
                if x not in self.models:
                    raise CalculatorTorchError(f"No model for species: {x}")
            per_cfg.append(torch.as_tensor([names.index(x) for x in sp], device=self.model.device))
        self._species_of_cfg = per_cfg

    def _model_energy(self, zeta, cfg_indices):
        if self._species_of_cfg is None or any(i is None or i < 0 for i in cfg_indices):
            raise CalculatorTorchError("per-species evaluation needs samples created by this calculator")
        sp = torch.cat([self._species_of_cfg[i] for i in cfg_indices])
        e_atom = torch.zeros((zeta.shape[0], 1), dtype=zeta.dtype, device=zeta.device)
        for code, name in enumerate(self._species_names):
            rows = torch.nonzero(sp == code, as_tuple=False).reshape(-1)
            if rows.numel():
                e_atom = e_atom.index_add(0, rows, self.models[name](zeta.index_select(0, rows)))
        return e_atom

    def save_model(self, epoch, force_save=False):
        for name, m in self.models.items():
            path = to_path(m.save_prefix).joinpath(f"model_{name}_epoch{epoch}.pkl")
            if force_save or (epoch >= m.save_start and (epoch - m.save_start) % m.save_frequency == 0):
                m.save(path)
