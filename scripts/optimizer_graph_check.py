"""Graphed vs eager `Loss.minimize` for several optimizers on 12 Si configurations (fp64 model).
Last run: SGD 2e-16, Adam 0.0 (graphs used); SGD+momentum and LBFGS stay eager by design."""
import sys, os, tempfile
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
import numpy as np, torch
from helpers import configs_from_npz
from kliff_b200.dataset.weight import Weight
from kliff_b200.legacy import nn
from kliff_b200.legacy.calculators import CalculatorTorch
from kliff_b200.legacy.descriptors import SymmetryFunction
from kliff_b200.legacy.loss import Loss
from kliff_b200.models import NeuralNetwork
os.chdir(tempfile.mkdtemp())
confs = configs_from_npz("si_varying_alat.npz", weight=Weight(forces_weight=0.3), limit=12)
def run(method, graphs, **kw):
    d = SymmetryFunction({"Si-Si": 5.0}, "cos", "set30", normalize=True, dtype=np.float64)
    m = NeuralNetwork(d); m.add_layers(nn.Linear(30, 8), nn.Tanh(), nn.Linear(8, 1))
    m.set_save_metadata(prefix="m", start=10**9, frequency=10**9)
    c = CalculatorTorch(m); c.create(confs)
    L = Loss(c); L.use_cuda_graphs = graphs
    L.minimize(method=method, num_epochs=3, batch_size=5, **kw)
    return np.array(L.epoch_losses), L._graphs_live()
for method, kw in (("SGD", dict(lr=1e-3)), ("Adam", dict(lr=1e-2)), ("SGD", dict(lr=1e-3, momentum=0.9)), ("LBFGS", dict(lr=0.1))):
    a, ga = run(method, True, **kw); b, gb = run(method, False, **kw)
    print(method, kw, "graphs used:", ga, gb, "max rel diff", float(np.max(np.abs(a / b - 1))))
