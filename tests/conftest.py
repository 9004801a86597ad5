import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name), allow_pickle=False)


def fams_from_npz(g):
    """(kind, values[rows, cols]) list from the packed fam_* arrays of a golden file."""
    ncols = {1: 1, 2: 2, 3: 1, 4: 3, 5: 3}
    fams, pos = [], 0
    for k, r in zip(g["fam_kinds"], g["fam_rows"]):
        c = ncols[int(k)]
        fams.append((int(k), np.ascontiguousarray(g["fam_vals"][pos:pos + r * c].reshape(r, c))))
        pos += r * c
    return fams


def rowwise_rel_err(a, b):
    """max over rows of max|a-b| / max|b| — the max-norm-relative metric of SURVEY.md §7/§8d
    (rows = (atom, descriptor))."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    a2 = a.reshape(a.shape[0], -1) if a.ndim > 1 else a.reshape(1, -1)
    b2 = b.reshape(a2.shape)
    scale = np.max(np.abs(b2), axis=1)
    err = np.max(np.abs(a2 - b2), axis=1)
    ok = scale > 0
    out = np.zeros_like(err)
    out[ok] = err[ok] / scale[ok]
    out[~ok] = err[~ok]
    return float(out.max()) if out.size else 0.0
