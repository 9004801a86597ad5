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


def test_integration_notes_name_every_entry_point():
    """INTEGRATION.md maps each C-ABI entry point to the reference interface it replaces."""
    text = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    missing = [e for e in _lib.EXPORTS if e not in text and not (e.startswith("kb_probe_") and "kb_probe_*" in text)]
    assert not missing, missing


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


def test_write_kim_model_files_round_trip(tmp_path):
    """KIM/DUNN export (neural_network.py:83-330 layouts): NN.params parses back to the layers'
    weights (transposed, row per input) and keep probabilities; descriptor.params and the CMake
    stub name the three parameter files.  Host-only: no kernel is involved."""
    import numpy as np
    import torch

    from kliff_b200.legacy import nn
    from kliff_b200.legacy.descriptors import SymmetryFunction
    from kliff_b200.models import NeuralNetwork
    from kliff_b200.models.neural_network import NeuralNetworkError

    desc = SymmetryFunction({"Si-Si": 5.0}, "cos", "set30", normalize=False, dtype=np.float64)
    model = NeuralNetwork(desc)
    model.add_layers(nn.Linear(30, 6), nn.Tanh(), nn.Dropout(0.25), nn.Linear(6, 4), nn.Tanh(), nn.Linear(4, 1))
    out = model.write_kim_model(tmp_path / "NN_roundtrip__MO_000000111111_000", dropout_ensemble_size=2)
    assert sorted(p.name for p in out.iterdir()) == ["CMakeLists.txt", "NN.params", "descriptor.params",
                                                     "dropout_binary.params"]
    lines = [ln.split("#")[0].split() for ln in open(out / "NN.params")]
    rows = [ln for ln in lines if ln]
    assert rows[0] == ["3"] and rows[1] == ["6", "4", "1"] and rows[2] == ["tanh"]
    assert [float(x) for x in rows[3]] == [1.0, 0.75, 1.0]
    pos = 4
    for layer in (model.layers[0], model.layers[3], model.layers[5]):
        w = layer.weight.detach().numpy().T
        got = np.array([[float(x) for x in r] for r in rows[pos:pos + w.shape[0]]])
        assert np.allclose(got, w, rtol=1e-14, atol=0)
        pos += w.shape[0]
        assert np.allclose([float(x) for x in rows[pos]], layer.bias.detach().numpy(), rtol=1e-14, atol=0)
        pos += 1
    assert pos == len(rows)
    drop = [ln.split("#")[0].split() for ln in open(out / "dropout_binary.params")]
    drop = [d for d in drop if d]
    assert drop[0] == ["2"] and [len(d) for d in drop[1:]] == [30, 6, 4, 30, 6, 4]
    assert 'PARAMETER_FILES "descriptor.params" "NN.params" "dropout_binary.params"' in open(out / "CMakeLists.txt").read()
    bad = NeuralNetwork(desc)
    bad.add_layers(nn.Linear(30, 4), nn.Linear(4, 1))
    try:
        bad.write_kim_model(tmp_path / "bad")
        assert False, "two Linear layers in a row must be refused"
    except NeuralNetworkError:
        pass
