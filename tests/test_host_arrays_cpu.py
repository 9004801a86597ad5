"""Host-side packing of configuration lists (kliff_b200/packed.py: host_arrays) against the plain
per-configuration conversion it replaces — the batched form of the reference's `for conf in configs`
loop (kliff/legacy/descriptors/descriptor.py:186-238)."""
import numpy as np
import pytest

from kliff_b200.dataset import Configuration
from kliff_b200.packed import host_arrays


def _plain(configs, species_code):
    sizes = np.asarray([c.get_num_atoms() for c in configs], np.int64)
    cfg_ptr = np.zeros(len(configs) + 1, np.int64)
    np.cumsum(sizes, out=cfg_ptr[1:])
    coords = np.concatenate([np.asarray(c.coords, np.float64).reshape(-1, 3) for c in configs])
    code = np.concatenate([np.asarray([species_code[s] for s in c.species], np.int32) for c in configs])
    cells = np.stack([np.asarray(c.cell, np.float64).reshape(3, 3) for c in configs])
    pbcs = np.stack([np.asarray(c.PBC).astype(bool) for c in configs])
    return cfg_ptr, coords, code, cells, pbcs


def _configs(rng, n, flat=False, f32=False):
    out = []
    for i in range(n):
        na = int(rng.integers(1, 9))
        pos = rng.normal(size=(na, 3))
        if f32:
            pos = pos.astype(np.float32)
        species = [("Si", "C")[int(b)] for b in rng.integers(0, 2, na)]
        cell = np.eye(3) * (4.0 + i) + rng.normal(0, 0.1, (3, 3))
        out.append(Configuration(cell=cell, species=species, coords=pos.reshape(-1) if flat and i % 2 else pos,
                                 PBC=[bool(b) for b in rng.integers(0, 2, 3)]))
    return out


@pytest.mark.parametrize("flat,f32", [(False, False), (True, False), (False, True)])
def test_host_arrays_matches_plain(flat, f32):
    rng = np.random.default_rng(7)
    configs = _configs(rng, 40, flat, f32)
    code = {"Si": 0, "C": 1}
    got, ref = host_arrays(configs, code), _plain(configs, code)
    for g, r in zip(got, ref):
        assert g.dtype == r.dtype and g.shape == r.shape and np.array_equal(g, r)


def test_host_arrays_unknown_species_and_empty():
    rng = np.random.default_rng(1)
    with pytest.raises(KeyError):
        host_arrays(_configs(rng, 3), {"Si": 0})
    ptr, coords, code, cells, pbcs = host_arrays([], {"Si": 0})
    assert ptr.tolist() == [0] and coords.shape == (0, 3) and code.shape == (0,)
