"""Base `Descriptor` with the interface of kliff/legacy/descriptors/descriptor.py:14-387.

`generate_fingerprints` keeps the reference's contract — mean/stdev over all atoms of all
configurations (np.mean / np.std, ddof 0, descriptor.py:116-118), normalisation, cast to
`dtype`, one pickled example per configuration appended to `fingerprints_filename`, the
mean/stdev state dict in a second pickle — but computes everything with the GPU kernels in one
batched pass and additionally keeps the packed, device-resident result in
`self.packed_fingerprints` for `CalculatorTorch` (no per-batch re-upload of dense tensors).
"""
import pickle
import threading
from pathlib import Path

import numpy as np
import torch

from ... import ops
from ...utils import create_directory, pickle_dump, to_path


class DescriptorError(Exception):
    def __init__(self, msg):
        super().__init__(msg)
        self.msg = msg


class Descriptor:
    def __init__(self, cut_dists, cut_name, hyperparams, normalize=True, dtype=np.float32):
        self.cut_dists = cut_dists
        self.cut_name = cut_name
        self.hyperparams = hyperparams
        self.normalize = normalize
        self.dtype = dtype
        self.size = None
        self.mean = None
        self.stdev = None
        # "packed": examples carry a PackedGradient (sparse, 36 KB/atom); "dense": the reference's
        # (N, D, 3N) arrays (O(N^2), only sensible for small cells)
        self.fingerprints_format = "packed"
        self.store_gradients_on_disk = True
        self.packed_fingerprints = None

    # ------------------------------------------------------------------ to be provided by subclasses
    def transform(self, conf, fit_forces=False, fit_stress=False):
        raise NotImplementedError

    def transform_packed(self, configs, grad=True, device=None, gradient_dtype=np.float64):
        raise NotImplementedError

    def write_kim_params(self, path, fname="descriptor.params"):
        raise NotImplementedError

    # ------------------------------------------------------------------ descriptor.py:65-147
    def generate_fingerprints(self, configs, fit_forces=False, fit_stress=False,
                              fingerprints_filename="fingerprints.pkl",
                              fingerprints_mean_stdev_filename=None, use_welford_method=False,
                              nprocs=1, reduce_over=None):
        """`reduce_over`: a torch.distributed module when `configs` is this rank's share of a
        larger set — mean and stdev are then the statistics of the whole set (two all-reduces).
        `nprocs` is accepted for interface compatibility; the work is one GPU pass.
        `use_welford_method` selects the same statistics computed in a streaming fashion in the
        reference (descriptor.py:243-282); both give the population mean/stdev, so it is a no-op."""
        if fit_stress and self.fingerprints_format != "dense":
            dense_stress = True
        else:
            dense_stress = fit_stress
        grad = fit_forces or fit_stress
        gdt = getattr(self, "gradient_dtype", None)
        packed = self.transform_packed(configs, grad=grad, gradient_dtype=gdt if gdt is not None else self.dtype)

        has_mean_stdev = self.mean is not None and self.stdev is not None
        if self.normalize and not has_mean_stdev:
            # np.mean / np.std (ddof = 0) in their two-pass form, as device reductions (kb_moments);
            # over all ranks' atoms when the set is sharded
            z = packed.zeta
            first = torch.cat([ops.moments(z), torch.tensor([float(z.shape[0])], dtype=z.dtype, device=z.device)])
            if reduce_over is not None:
                reduce_over.all_reduce(first, op=reduce_over.ReduceOp.SUM)
            mean = first[:-1] / first[-1]
            second = ops.moments(z, mean)
            if reduce_over is not None:
                reduce_over.all_reduce(second, op=reduce_over.ReduceOp.SUM)
            stdev = torch.sqrt(second / first[-1])
            self.mean = mean.cpu().numpy()
            self.stdev = stdev.cpu().numpy()
            if fingerprints_mean_stdev_filename is None:
                fingerprints_mean_stdev_filename = "fingerprints_mean_and_stdev.pkl"
            pickle_dump(self.state_dict(), fingerprints_mean_stdev_filename)

        if self.normalize:
            mean_d = torch.as_tensor(self.mean, device=packed.device, dtype=torch.float64)
            inv_d = 1.0 / torch.as_tensor(self.stdev, device=packed.device, dtype=torch.float64)
            packed.zeta_normalized = ops.normalize(packed.zeta, mean_d, inv_d)
            packed.scale = inv_d.contiguous()
        else:
            packed.zeta_normalized = packed.zeta
            packed.scale = None
        self.packed_fingerprints = packed
        # the examples just written, for callers that would otherwise read the file straight back
        self.last_examples = self._dump_fingerprints(configs, fingerprints_filename, packed, fit_forces,
                                                     dense_stress)
        return fingerprints_filename

    def _dump_fingerprints(self, configs, fname, packed, fit_forces, fit_stress):
        """descriptor.py:149-218 — same keys per example."""
        create_directory(fname, is_directory=False)
        fname = to_path(fname)
        if fname.exists():
            fname.unlink()
        # cast on the device (fp64 -> `dtype` rounds as np.asarray(zeta, dtype) does, descriptor.py:197), so that
        # half the bytes cross PCIe for the default float32 and the per-configuration entries are plain views
        np_dtype = np.dtype(self.dtype)
        zn = packed.zeta_normalized
        if np_dtype == np.float32 and zn.dtype == torch.float64:
            zn = zn.to(torch.float32)
        zeta_all = zn.cpu().numpy()
        if zeta_all.dtype != np_dtype:
            zeta_all = zeta_all.astype(np_dtype)
        inv = packed.scale
        ptr = np.asarray(packed.cfg_ptr).tolist()
        examples = []
        for i, conf in enumerate(configs):
            example = {"configuration": conf, "zeta": zeta_all[ptr[i]:ptr[i + 1]],
                       "energy": np.asarray(conf.energy, self.dtype)}
            if fit_forces:
                if self.fingerprints_format == "dense":
                    dense, _ = packed.dense_config(i, True, False)
                    if inv is not None:
                        dense = dense * inv[None, :, None]
                    example["dzetadr_forces"] = np.asarray(dense.cpu().numpy(), self.dtype)
                elif self.store_gradients_on_disk:
                    example["dzetadr_forces"] = packed.packed_gradient_host(i, self.dtype, inv)
                example["forces"] = np.asarray(conf.forces, self.dtype)
            if fit_stress:
                _, stress = packed.dense_config(i, False, True)
                if inv is not None:
                    stress = stress * inv[None, :, None]
                example["dzetadr_stress"] = np.asarray(stress.cpu().numpy(), self.dtype)
                example["stress"] = np.asarray(conf.stress, self.dtype)
                example["volume"] = np.asarray(conf.get_volume(), self.dtype)
            examples.append(example)

        # the writer works on its own shallow copies: callers may add keys to the returned examples while
        # the thread is still pickling ("dictionary changed size during iteration" otherwise)
        snapshot = [dict(e) for e in examples]

        def write():
            with open(fname, "ab") as f:
                # one Pickler for the whole stream, memo cleared per example: the file is the same sequence
                # of independent pickles that repeated pickle.dump() calls write (load_fingerprints reads it)
                pickler = pickle.Pickler(f, protocol=pickle.HIGHEST_PROTOCOL)
                for example in snapshot:
                    pickler.dump(example)
                    pickler.clear_memo()

        # Serialising the examples is 40 % of create() on large sets (20 000 configurations: 0.50 of 1.16 s,
        # scripts/profile_create.py) and nothing in this process needs the file before a later
        # `reuse=True`: write it on a background thread.  load_fingerprints() / process exit wait for it.
        if len(examples) >= getattr(self, "background_write_min_examples", 256):
            t = threading.Thread(target=write, name="kliff_b200-fingerprints-writer", daemon=False)
            _PENDING_WRITES[str(fname.resolve())] = t
            t.start()
        else:
            write()
        return examples

    # ------------------------------------------------------------------ getters (descriptor.py:315-353)
    def get_size(self):
        return len(self.mean) if self.mean is not None else self.size

    def get_mean(self):
        return self.mean

    def get_stdev(self):
        return self.stdev

    def get_dtype(self):
        return self.dtype

    def get_cutoff(self):
        return self.cut_name, self.cut_dists

    def get_hyperparams(self):
        return self.hyperparams

    def state_dict(self):
        return {"mean": self.mean, "stdev": self.stdev, "size": self.size}

    def load_state_dict(self, data):
        try:
            mean, stdev, size = data["mean"], data["stdev"], data["size"]
        except Exception as e:
            raise DescriptorError(f"Corrupted state dict for descriptor: {str(e)}")
        if mean is not None and stdev is not None and size is not None:
            if len(mean.shape) != 1 or mean.shape[0] != size:
                raise DescriptorError("Corrupted descriptor mean.")
            if len(stdev.shape) != 1 or stdev.shape[0] != size:
                raise DescriptorError("Corrupted descriptor standard deviation.")
        self.mean, self.stdev, self.size = mean, stdev, size


_PENDING_WRITES = {}  # resolved path -> writer thread of a fingerprint file still being written


def wait_for_fingerprints(path=None):
    """Block until the background writer of `path` (default: of every file) has finished."""
    keys = list(_PENDING_WRITES) if path is None else [str(to_path(path).resolve())]
    for k in keys:
        t = _PENDING_WRITES.pop(k, None)
        if t is not None:
            t.join()


def load_fingerprints(path):
    """Inverse of `_dump_fingerprints` (descriptor.py:390-413)."""
    wait_for_fingerprints(path)
    data = []
    with open(path, "rb") as f:
        try:
            while True:
                data.append(pickle.load(f))
        except EOFError:
            pass
        except Exception as e:
            raise DescriptorError(f"Cannot load fingerprints from `{path}`. {str(e)}")
    return data


def generate_full_cutoff(cutoff):
    """{'C-C': 4, 'C-H': 3.5} -> adds 'H-C': 3.5 (descriptor.py:416-449)."""
    full = {}
    for key, val in cutoff.items():
        s1, s2 = key.split("-")
        if s1 != s2:
            rev = s2 + "-" + s1
            if rev in cutoff and cutoff[rev] != val:
                raise Exception(
                    'Corrupted cutoff dictionary. cutoff["{0}-{1}"] != cutoff["{1}-{0}"].'.format(s1, s2))
            full[rev] = val
    full.update(cutoff)
    return full


def generate_unique_cutoff_pairs(cutoff):
    """Drop the mirrored keys again (descriptor.py:452-479)."""
    out = {}
    for key, val in cutoff.items():
        s1, s2 = key.split("-")
        if key not in out and (s2 + "-" + s1) not in out:
            out[key] = val
    return out


def generate_species_code(cutoff):
    """Species -> integer code.  The reference enumerates a `set` (descriptor.py:482-505), i.e. in
    hash order; codes are only ever used consistently with the cutoff matrix built from the same
    map, so a deterministic first-appearance order is equivalent."""
    seen = []
    for key in cutoff:
        for s in key.split("-"):
            if s not in seen:
                seen.append(s)
    return {s: i for i, s in enumerate(seen)}
