"""Loss weights of a configuration (kliff/dataset/weight.py): the attributes `energy_forces_residual`
reads (kliff/legacy/loss.py:96-99).  `Weight` holds four fixed numbers; `MagnitudeInverseWeight`
computes them from the configuration's own data, and its `forces_weight` is an array (one value per
force component), which the training loss handles on its per-configuration route."""
import warnings

import numpy as np


class WeightError(Exception):
    def __init__(self, msg):
        super().__init__(msg)
        self.msg = msg


class Weight:
    """Fixed weights (weight.py:8-107)."""

    def __init__(self, config_weight=1.0, energy_weight=1.0, forces_weight=1.0, stress_weight=1.0):
        self._config_weight = config_weight
        self._energy_weight = energy_weight
        self._forces_weight = forces_weight
        self._stress_weight = stress_weight

    def compute_weight(self, config):
        self._check_compute_flag(config)

    config_weight = property(lambda self: self._config_weight,
                             lambda self, value: setattr(self, "_config_weight", value))
    energy_weight = property(lambda self: self._energy_weight,
                             lambda self, value: setattr(self, "_energy_weight", value))
    forces_weight = property(lambda self: self._forces_weight,
                             lambda self, value: setattr(self, "_forces_weight", value))
    stress_weight = property(lambda self: self._stress_weight,
                             lambda self, value: setattr(self, "_stress_weight", value))

    def _check_compute_flag(self, config):
        """weight.py:78-98: a tiny but non-zero weight usually means the property should not be fitted."""
        msg = ('"{0}_weight" are near zero. Seems you do not want to use {0} in the fitting. You can set '
               '"use_{0}" in "calculator.create()" to "False" to speed up the fitting.')
        for name, data in (("energy", config._energy), ("forces", config._forces), ("stress", config._stress)):
            w = getattr(self, name + "_weight")
            if data is not None and w is not None and np.all(np.asarray(w) < 1e-12):
                warnings.warn(msg.format(name), stacklevel=3)

    def to_dict(self):
        return {
            "config": self.config_weight,
            "energy": self.energy_weight,
            "forces": self.forces_weight,
            "stress": self.stress_weight,
        }

    def __repr__(self):
        return (
            f"Weights: config={self.config_weight}, energy={self.energy_weight}, "
            f"forces={self.forces_weight}, stress={self.stress_weight}"
        )


class MagnitudeInverseWeight(Weight):
    """Weights computed from the data (weight.py:109-232; Lenosky et al. 1997): 1 / w² = c1² + c2² |f|², with
    |f| the absolute energy, the norm of each atom's force vector (repeated for its three components) or the
    Frobenius norm of the stress tensor.  `weight_params`: `energy_weight_params`, `forces_weight_params`,
    `stress_weight_params`, each a number (c1; c2 = 0) or a pair [c1, c2]; default [1, 0]."""

    default_weight_params = {
        "energy_weight_params": [1.0, 0.0],
        "forces_weight_params": [1.0, 0.0],
        "stress_weight_params": [1.0, 0.0],
    }

    def __init__(self, config_weight=1.0, weight_params=None):
        super().__init__(config_weight, 0.0, 0.0, 0.0)
        self._weight_params = self._check_weight_params(
            weight_params, {k: list(v) for k, v in self.default_weight_params.items()})

    def compute_weight(self, config):
        energy, forces, stress = config._energy, config._forces, config._stress
        if energy is not None:
            self._energy_weight = self._compute_weight_one_property(
                [np.abs(energy)], self._weight_params["energy_weight_params"], "energy")[0]
        if forces is not None:
            norms = np.linalg.norm(np.asarray(forces, dtype=np.float64).reshape(-1, 3), axis=1)
            self._forces_weight = np.repeat(self._compute_weight_one_property(
                norms, self._weight_params["forces_weight_params"], "forces"), 3)
        if stress is not None:
            stress = np.asarray(stress, dtype=np.float64)
            norm = np.sqrt(np.linalg.norm(stress[:3]) ** 2 + 2 * np.linalg.norm(stress[3:]) ** 2)
            self._stress_weight = self._compute_weight_one_property(
                [norm], self._weight_params["stress_weight_params"], "stress")[0]
        self._check_compute_flag(config)

    @staticmethod
    def _compute_weight_one_property(data_norm, property_weight_params, property_type):
        c1, c2 = property_weight_params
        sigma = np.sqrt(c1 ** 2 + (c2 * np.asarray(data_norm, dtype=np.float64)) ** 2)
        if np.any(sigma < 1e-12):
            warnings.warn(f"Found near zero inverse {property_type} weight. Be aware that some "
                          f"{property_type} data might be overweight.", stacklevel=3)
        with np.errstate(divide="ignore"):
            return 1 / sigma

    @staticmethod
    def _check_weight_params(weight_params, default):
        if weight_params is not None:
            for key, value in weight_params.items():
                if key not in default:
                    raise WeightError(f"Expect the keys of `weight_params` to be one or combinations "
                                      f"of {', '.join(default.keys())}; got {key}. ")
                if np.ndim(value) == 0:
                    default[key][0] = value
                elif np.ndim(value) == 1 and len(value) == 2:
                    default[key] = value
                else:
                    raise WeightError("Expect a float or a list of floats with format [c1, c2]")
        return default
