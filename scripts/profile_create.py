"""cProfile of CalculatorTorch.create on the synthetic 64-atom-configuration set (host-side costs)."""
import cProfile
import io
import os
import pstats
import sys
import tempfile

import numpy as np
import torch

sys.path.insert(0, ".")
from kliff_b200.dataset import Configuration
from kliff_b200.dataset.weight import Weight
from kliff_b200.legacy import nn
from kliff_b200.legacy.calculators import CalculatorTorch
from kliff_b200.legacy.descriptors import SymmetryFunction
from kliff_b200.models import NeuralNetwork


def main():
    ncfg = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
    basis = np.array([[0, 0, 0], [0, .5, .5], [.5, 0, .5], [.5, .5, 0],
                      [.25, .25, .25], [.25, .75, .75], [.75, .25, .75], [.75, .75, .25]])
    g = np.stack(np.meshgrid(np.arange(2), np.arange(2), np.arange(2), indexing="ij"), -1).reshape(-1, 3)
    frac = (g[:, None, :] + basis[None]).reshape(-1, 3)
    weight = Weight(forces_weight=0.3)
    confs = []
    for c in range(ncfg):
        rng = np.random.default_rng(c)
        a = rng.uniform(5.2, 5.7)
        confs.append(Configuration(cell=np.eye(3) * 2 * a, species=["Si"] * 64,
                                   coords=frac * a + rng.normal(0, 0.1, frac.shape), PBC=[True] * 3,
                                   energy=float(rng.normal()), forces=rng.normal(size=(64, 3)),
                                   weight=weight, identifier=f"syn{c}"))
    tmp = tempfile.mkdtemp()

    def create():
        desc = SymmetryFunction(cut_name="cos", cut_dists={"Si-Si": 5.0}, hyperparams="set51", normalize=True)
        desc.store_gradients_on_disk = False
        model = NeuralNetwork(desc)
        model.add_layers(nn.Linear(51, 10), nn.Tanh(), nn.Linear(10, 10), nn.Tanh(), nn.Linear(10, 1))
        calc = CalculatorTorch(model)
        calc.create(confs, fingerprints_filename=os.path.join(tmp, "fp.pkl"),
                    fingerprints_mean_stdev_filename=os.path.join(tmp, "ms.pkl"))
        torch.cuda.synchronize()

    create()  # warm-up
    pr = cProfile.Profile()
    pr.enable()
    create()
    pr.disable()
    out = io.StringIO()
    pstats.Stats(pr, stream=out).sort_stats("cumulative").print_stats(22)
    print(out.getvalue())


if __name__ == "__main__":
    main()
