class Weight:
    """Loss weights of one configuration; mirrors kliff/dataset/weight.py:8-108 (the
    attributes `energy_forces_residual` reads, kliff/legacy/loss.py:96-99)."""

    def __init__(self, config_weight=1.0, energy_weight=1.0, forces_weight=1.0, stress_weight=1.0):
        self.config_weight = config_weight
        self.energy_weight = energy_weight
        self.forces_weight = forces_weight
        self.stress_weight = stress_weight

    def compute_weight(self, config):
        pass

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
