"""Loss weights (kliff/dataset/weight.py) and how the training loss consumes them — the reference's own
tests/dataset/test_weight.py restated for the in-scope part (fixed and magnitude-inverse weights; the
weight-file readers belong to the ase-based dataset loaders, out of scope), plus the residual functions of
kliff/legacy/loss.py:37-166 with a weight per force component."""
import warnings

import numpy as np
import pytest
import torch

from kliff_b200.dataset import Configuration, MagnitudeInverseWeight, Weight, WeightError
from kliff_b200.legacy.loss import (LossNeuralNetworkModel, energy_forces_residual, energy_residual,
                                    forces_residual)


def _config(weight, seed=2022, natoms=8, with_stress=True):
    rng = np.random.default_rng(seed)
    return Configuration(cell=np.eye(3) * 5.43, species=["Si"] * natoms, coords=rng.random((natoms, 3)) * 5.43,
                         PBC=[True] * 3, energy=float(rng.normal() * 10), forces=rng.normal(size=(natoms, 3)),
                         stress=list(rng.normal(size=6)) if with_stress else None, weight=weight)


def test_base_weight():
    cw, ew, fw, sw = np.random.default_rng(2022).uniform(0.0, 1.1, size=4)
    weight = Weight(config_weight=cw, energy_weight=ew, forces_weight=fw, stress_weight=sw)
    assert (weight.config_weight, weight.energy_weight, weight.forces_weight, weight.stress_weight) == (cw, ew, fw, sw)
    weight.forces_weight = 0.3
    assert weight.forces_weight == 0.3 and weight.to_dict()["forces"] == 0.3
    assert Weight().to_dict() == {"config": 1.0, "energy": 1.0, "forces": 1.0, "stress": 1.0}


def _manual(c1, c2, norm):
    return 1 / np.sqrt(c1 ** 2 + (c2 * norm) ** 2)


def test_magnitude_inverse_weight():
    """tests/dataset/test_weight.py:29-82: against the formula written out by hand."""
    rng = np.random.default_rng(7)
    (c1_e, c2_e), (c1_f, c2_f), (c1_s, c2_s) = rng.uniform(0.0, 1.0, size=(3, 2))
    weight = MagnitudeInverseWeight(weight_params={"energy_weight_params": [c1_e, c2_e],
                                                   "forces_weight_params": [c1_f, c2_f],
                                                   "stress_weight_params": [c1_s, c2_s]})
    conf = _config(weight)
    ew = conf.weight.energy_weight
    assert isinstance(float(ew), float) and np.ndim(ew) == 0
    assert np.allclose(ew, _manual(c1_e, c2_e, np.abs(conf.energy)), rtol=1e-12, atol=1e-12)
    fw = conf.weight.forces_weight
    assert isinstance(fw, np.ndarray) and len(fw) == conf.forces.size
    assert np.allclose(fw, _manual(c1_f, c2_f, np.repeat(np.linalg.norm(conf.forces, axis=1), 3)), rtol=1e-12, atol=1e-12)
    stress = np.asarray(conf.stress)
    norm = np.sqrt(np.linalg.norm(stress) ** 2 + np.linalg.norm(stress[3:]) ** 2)
    assert np.allclose(conf.weight.stress_weight, _manual(c1_s, c2_s, norm), rtol=1e-12, atol=1e-12)
    assert conf.weight.config_weight == 1.0


def test_magnitude_inverse_weight_parameter_forms_and_errors():
    w = MagnitudeInverseWeight(config_weight=2.0, weight_params={"forces_weight_params": 0.5})
    conf = _config(w, with_stress=False)
    assert np.allclose(conf.weight.forces_weight, 2.0) and conf.weight.energy_weight == 1.0      # c2 = 0: 1 / c1
    assert conf.weight.stress_weight == 0.0 and conf.weight.config_weight == 2.0                 # no stress data
    assert MagnitudeInverseWeight()._weight_params == MagnitudeInverseWeight.default_weight_params
    with pytest.raises(WeightError):
        MagnitudeInverseWeight(weight_params={"force_weight_params": [1.0, 0.0]})
    with pytest.raises(WeightError):
        MagnitudeInverseWeight(weight_params={"forces_weight_params": [1.0, 0.0, 3.0]})
    with warnings.catch_warnings(record=True) as got:
        warnings.simplefilter("always")
        _config(Weight(forces_weight=1e-15))
    assert any("forces_weight" in str(g.message) for g in got)


def test_residuals_with_a_weight_per_force_component():
    """energy_forces_residual / forces_residual (loss.py:37-166) with MagnitudeInverseWeight: same numbers as the
    reference's CPU arithmetic (tensor *= ndarray), on whatever device the prediction lives."""
    conf = _config(MagnitudeInverseWeight(config_weight=0.7, weight_params={"energy_weight_params": [0.5, 0.1],
                                                                           "forces_weight_params": [0.3, 2.0]}),
                   with_stress=False)
    n = conf.get_num_atoms()
    rng = np.random.default_rng(1)
    for dtype in (torch.float64, torch.float32):
        pred = torch.tensor(rng.normal(size=1 + 3 * n), dtype=dtype)
        ref = torch.tensor(np.r_[conf.energy, conf.forces.reshape(-1)], dtype=dtype)
        w = conf.weight
        res = energy_forces_residual("id", n, w, pred, ref, {"normalize_by_natoms": True})
        expect = 0.7 * (pred - ref).numpy().astype(np.float64)
        expect[0] *= w.energy_weight
        expect[1:] *= w.forces_weight
        expect /= n
        tol = 1e-14 if dtype == torch.float64 else 1e-6
        assert res.dtype == dtype and np.allclose(res.numpy(), expect, rtol=tol, atol=tol)
        res_f = forces_residual("id", n, w, pred[1:], ref[1:], {"normalize_by_natoms": False})
        assert np.allclose(res_f.numpy(), 0.7 * w.forces_weight * (pred[1:] - ref[1:]).numpy(), rtol=tol, atol=tol)
        res_e = energy_residual("id", n, w, pred[:1], ref[:1], {"normalize_by_natoms": True})
        assert np.allclose(res_e.numpy(), 0.7 * w.energy_weight * (pred[:1] - ref[:1]).numpy() / n, rtol=tol, atol=tol)
    # fixed weights: unchanged path (numbers stay numbers)
    plain = energy_forces_residual("id", n, Weight(forces_weight=0.3), pred, ref, {"normalize_by_natoms": False})
    assert np.allclose(plain.numpy()[1:], 0.3 * (pred - ref).numpy()[1:])


class _DS:
    def __init__(self, fp):
        self.fp = fp


class _Calc:
    def __init__(self, fp):
        self.fingerprints_dataset = _DS(fp)


def test_array_weights_leave_the_vectorised_route():
    loss = LossNeuralNetworkModel.__new__(LossNeuralNetworkModel)
    loss.calculator = _Calc([{"configuration": _config(Weight(forces_weight=0.3), seed=s)} for s in range(3)])
    assert loss._scalar_weights()
    loss.calculator = _Calc([{"configuration": _config(Weight(), seed=0)},
                             {"configuration": _config(MagnitudeInverseWeight(), seed=1)}])
    assert not loss._scalar_weights()


# ------------------------------------------------------------------ new-API Descriptor: host-side writers
def test_descriptor_state_and_kim_param_files_equal_the_reference_writers(tmp_path):
    """`Descriptor.save_descriptor_state` / `export_kim_model`
    (kliff/transforms/configuration_transforms/descriptors.py:358-512) against files written by the reference's
    own functions (tests/golden/make_descriptor_state_golden.py)."""
    import os
    from collections import OrderedDict

    from conftest import GOLDEN
    from kliff_b200.transforms.configuration_transforms.descriptors import Descriptor

    d30 = Descriptor(cutoff=5.0, species=["Si"], descriptor="SymmetryFunctions", hyperparameters="set30")
    d30.save_descriptor_state(tmp_path, "a.dat")
    assert (tmp_path / "a.dat").read_text() == open(os.path.join(GOLDEN, "descriptor_state_set30.dat")).read()
    custom = OrderedDict([
        ("g1", None),
        ("g2", [{"eta": 0.0035710676725828126, "Rs": 0.0}, {"eta": 0.03571067672582813, "Rs": 1.5}]),
        ("g3", [{"kappa": 0.3}, {"kappa": 1.25}]),
        ("g4", [{"zeta": 1, "lambda": -1, "eta": 0.00035710676725828126}, {"zeta": 2.5, "lambda": 0.5, "eta": 0.02}]),
        ("g5", [{"zeta": 4, "lambda": 1, "eta": 0.001}]),
    ])
    dc = Descriptor(cutoff=4.5, species=["Si", "C"], descriptor="SymmetryFunctions", hyperparameters=custom)
    assert dc.width == 8
    dc.save_descriptor_state(tmp_path)
    assert (tmp_path / "descriptor.dat").read_text() == open(os.path.join(GOLDEN, "descriptor_state_custom.dat")).read()
    dc.export_kim_model(str(tmp_path), "model.pt")
    assert (tmp_path / "kim_model.param").read_text() == open(os.path.join(GOLDEN, "kim_model_param.txt")).read()
    assert dc.inverse() is None and dc.copy_to_config is False
    conf = _config(Weight())
    conf.fingerprint = "x"
    assert conf.fingerprint == "x"
