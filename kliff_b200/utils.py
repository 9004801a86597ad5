"""Small helpers shared by the plugin-API mirror (same roles as kliff/utils.py)."""
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


def default_device():
    """The CUDA device of this process (one process per GPU; honours torch.cuda.set_device)."""
    if not torch.cuda.is_available():
        raise RuntimeError(
            "kliff_b200 needs a CUDA device (B200, sm_100a): the descriptor path has no CPU fallback"
        )
    return torch.device("cuda", torch.cuda.current_device())
