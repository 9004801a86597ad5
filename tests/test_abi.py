"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports every symbol
include/kliff_b200.h declares; the parameter packing is what the kernels expect.  No compute."""
import ctypes
import os
import re

import numpy as np
import pytest

from conftest import ROOT
from kliff_b200 import _lib, ops
from kliff_b200.legacy.descriptors.hyperparams import get_set30, get_set51


@pytest.fixture(scope="module")
def built():
    if not os.path.exists(_lib.LIB_PATH):
        import __graft_entry__ as g
        g.build()
    return _lib.lib()


def test_library_exports_every_declared_symbol(built):
    header = open(os.path.join(ROOT, "include", "kliff_b200.h")).read()
    declared = set(re.findall(r"\b(kb_[a-z0-9_]+)\s*\(", header))
    assert declared, "no declarations parsed"
    assert declared == set(_lib.EXPORTS)
    for name in declared:
        assert hasattr(built, name), name
    assert b"sm_100a" in built.kb_version()
    assert built.kb_error_string(1).decode().startswith("reference error")


def test_struct_layout_matches_header(built):
    """sizeof(kb_symfun_params) computed from the header's array bounds."""
    S, T, G, W = 8, 64, 32, 8
    expect = 4 * 4 + 8 * S * S + 4 * T * 2 + 8 * T * 2 + 4 * G * 2 + 8 * G + 8 * G * W * 2 + 4 * G * W + 8
    assert ctypes.sizeof(_lib.kb_symfun_params) == expect


def test_params_set51_grouping():
    P = ops.build_params(get_set51(), np.array([[5.0]]))
    assert (P.n_desc, P.n_two, P.n_groups, P.canonical_zl) == (51, 8, 7, 1)
    outs = sorted(P.grp_out[g][e] for g in range(7) for e in range(8) if P.grp_out[g][e] >= 0)
    assert outs == list(range(8, 51))
    # (zeta 16, lambda -1, eta 0.08) is absent from set51 (sym_fn.py:462)
    assert P.grp_out[6][6] == -1 and P.grp_out[6][7] == 50
    assert [P.two_out[q] for q in range(8)] == list(range(8))
    P30 = ops.build_params(get_set30(), np.array([[5.0]]))
    assert (P30.n_desc, P30.n_groups, P30.canonical_zl) == (30, 6, 1)


def test_params_generic_and_limits():
    from collections import OrderedDict
    hp = OrderedDict()
    hp["g1"] = None
    hp["g4"] = [{"zeta": 3, "lambda": -1, "eta": 0.01}, {"zeta": 3, "lambda": -1, "eta": 0.01}]
    hp["g5"] = [{"zeta": 1, "lambda": 1, "eta": 0.01}]
    P = ops.build_params(hp, np.array([[4.0, 3.5], [3.5, 3.0]]))
    assert P.canonical_zl == 0 and P.n_desc == 4 and P.n_two == 1
    assert P.n_groups == 3  # duplicate (zeta, lambda, eta) rows cannot share a group
    assert P.rcut[1] == 3.5 and P.rcut[3] == 3.0
    with pytest.raises(ValueError):
        ops.build_params(hp, np.ones((9, 9)))


def test_no_cpu_fallback():
    import torch
    with pytest.raises(RuntimeError, match="no CPU implementation"):
        ops.create_paddings(3.0, np.eye(3), [1, 1, 1], torch.zeros((2, 3), dtype=torch.float64))
