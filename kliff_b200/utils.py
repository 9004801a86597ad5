"""Small helpers shared by the plugin-API mirror (same roles as kliff/utils.py)."""
import contextlib
import gc
import os
import pickle
import random
from pathlib import Path

import numpy as np
import torch


def seed_all(seed=35, cudnn_benchmark=False, cudnn_deterministic=False):
    """kliff/utils.py:75-86 — the MLP initialisation of the tutorials depends on it."""
    random.seed(seed)
    os.environ["PYTHONHASHSEED"] = str(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)
    if torch.cuda.is_available():
        torch.cuda.manual_seed(seed)
        torch.cuda.manual_seed_all(seed)
    torch.backends.cudnn.benchmark = cudnn_benchmark
    torch.backends.cudnn.deterministic = cudnn_deterministic


def to_path(path):
    return Path(path).expanduser().resolve()


def create_directory(path, is_directory=False):
    p = to_path(path)
    d = p if is_directory else p.parent
    if not d.exists():
        os.makedirs(d)


def pickle_dump(data, filename):
    create_directory(filename)
    with open(to_path(filename), "wb") as f:
        pickle.dump(data, f)


def pickle_load(filename):
    with open(to_path(filename), "rb") as f:
        return pickle.load(f)


@contextlib.contextmanager
def capture_without_gc():
    """Wrap a CUDA-graph capture: Python's cyclic collector stays off until the capture has ended.  A dead
    reference cycle that owns CUDA graphs (an earlier Loss / trainer of this process) may otherwise be
    collected in the middle of the capture; destroying a graph while a stream captures in global mode
    invalidates the capture (cudaErrorStreamCaptureInvalidated — seen on one rank of a 2-GPU run that built
    its seventh trainer; torch.cuda.graph itself no longer collects on entry).  Nothing is collected here
    either (a full collection costs 50-100 ms beside a 20 000-configuration set): the garbage simply waits
    for the first automatic collection after the capture."""
    was_enabled = gc.isenabled()
    gc.disable()
    try:
        yield
    finally:
        if was_enabled:
            gc.enable()


def default_device():
    """The CUDA device of this process (one process per GPU; honours torch.cuda.set_device)."""
    if not torch.cuda.is_available():
        raise RuntimeError(
            "kliff_b200 needs a CUDA device (B200, sm_100a): the descriptor path has no CPU fallback"
        )
    return torch.device("cuda", torch.cuda.current_device())


def map_species(species, mapping, what="species"):
    """Vectorised `[mapping[s] for s in species]` -> int32 array (a 1M-atom configuration must
    not cost a million dict lookups, nor a million string comparisons: fixed-width unicode arrays of
    up to two characters are compared as integers)."""
    arr = np.asarray(species)
    out = np.empty(arr.shape[0], dtype=np.int32)
    out.fill(-1)
    if arr.shape[0] == 0:
        return out
    if arr.dtype.kind == "U" and arr.dtype.itemsize in (4, 8) and arr.ndim == 1:
        keys = np.ascontiguousarray(arr).view(np.uint32 if arr.dtype.itemsize == 4 else np.uint64)
        if not (keys != keys[0]).any():  # one species (the common large case): two passes over the array
            name = str(arr[0])
            if name not in mapping:
                raise KeyError(name)
            out.fill(mapping[name])
            return out
        pos, seen = 0, 0
        while seen <= 128:
            k = keys[pos]
            name = str(np.array([k], dtype=keys.dtype).view(arr.dtype)[0])
            if name not in mapping:
                raise KeyError(name)
            np.putmask(out, keys == k, mapping[name])
            seen += 1
            left = out < 0
            if not left.any():
                return out
            pos = int(left.argmax())
        out[:] = -1
    for name in np.unique(arr) if arr.shape[0] < 4096 else _unique_large(arr):
        if str(name) not in mapping:
            raise KeyError(str(name))
        out[arr == name] = mapping[str(name)]
    return out


def _unique_large(arr):
    # np.unique sorts strings; on large uniform arrays peel distinct values off one at a time
    found, rest = [], arr
    while rest.shape[0]:
        found.append(rest[0])
        rest = rest[rest != rest[0]]
        if len(found) > 128:
            return np.unique(arr)
    return found
