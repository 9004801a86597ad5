"""Shared builders for the API-level tests."""
import os

import numpy as np

from conftest import GOLDEN
from kliff_b200.dataset import Configuration
from kliff_b200.dataset.weight import Weight

SYMBOL = {14: "Si", 6: "C", 8: "O", 42: "Mo", 16: "S"}


def configs_from_npz(name, weight=None, limit=None):
    """tests/golden/{si_varying_alat,sic_training_set}.npz -> list of Configuration in the
    reference's dataset order (sorted file names, kliff/dataset/dataset.py:233-237)."""
    g = np.load(os.path.join(GOLDEN, name), allow_pickle=False)
    out, pos = [], 0
    n_cfg = len(g["natoms"]) if limit is None else min(limit, len(g["natoms"]))
    for c in range(n_cfg):
        n = int(g["natoms"][c])
        sl = slice(pos, pos + n)
        out.append(Configuration(cell=g["cell"][c], species=[SYMBOL[int(z)] for z in g["z"][sl]],
                                 coords=g["coords"][sl], PBC=[bool(p) for p in g["pbc"][c]],
                                 energy=float(g["energy"][c]), forces=g["forces"][sl],
                                 weight=weight, identifier=str(g["names"][c])))
        pos += n
    return out


def config_from_golden(g, symbols=None):
    z = g["z"]
    sp = [SYMBOL[int(v)] for v in z] if symbols is None else symbols
    return Configuration(cell=g["cell"], species=sp, coords=g["coords"],
                         PBC=[bool(p) for p in g["pbc"]])
