"""Wall-clock breakdown of CalculatorTorch.create on the synthetic 64-atom-configuration set: each phase is
wrapped with perf_counter + a device synchronisation, so host and device time of a phase land on that phase.
Usage: python scripts/create_breakdown.py [n_configs]"""
import os
import sys
import tempfile
import time

import numpy as np
import torch

sys.path.insert(0, ".")
import bench  # noqa: E402  (synthetic set builder)
from kliff_b200 import packed as packed_mod  # noqa: E402
from kliff_b200.dataset import dataset_torch  # noqa: E402
from kliff_b200.legacy import nn  # noqa: E402
from kliff_b200.legacy.calculators import CalculatorTorch  # noqa: E402
from kliff_b200.legacy.descriptors import SymmetryFunction  # noqa: E402
from kliff_b200.legacy.descriptors.descriptor import Descriptor, wait_for_fingerprints  # noqa: E402
from kliff_b200 import ops  # noqa: E402
from kliff_b200.models import NeuralNetwork  # noqa: E402

T = {}


def timed(owner, name, label=None, static=False):
    fn = getattr(owner, name)
    raw = fn.__func__ if hasattr(fn, "__func__") else fn
    label = label or f"{getattr(owner, '__name__', owner)}.{name}"

    def wrapper(*a, **k):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        r = raw(*a, **k)
        torch.cuda.synchronize()
        T[label] = T.get(label, 0.0) + time.perf_counter() - t0
        return r
    if isinstance(owner.__dict__.get(name), classmethod):
        setattr(owner, name, classmethod(wrapper))
    else:
        setattr(owner, name, wrapper)


def main():
    ncfg = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
    confs = bench.synthetic_configs(ncfg)
    timed(packed_mod.PackedFingerprints, "from_configs")
    timed(packed_mod.PackedFingerprints, "compute")
    timed(Descriptor, "_dump_fingerprints")
    timed(Descriptor, "generate_fingerprints")
    timed(dataset_torch.FingerprintsDataset, "__init__", "FingerprintsDataset.__init__")
    timed(CalculatorTorch, "_upload_references")
    timed(ops, "batch_neighbors", "ops.batch_neighbors")
    timed(ops, "symfun_forward", "ops.symfun_forward")
    timed(ops, "moments", "ops.moments")
    timed(ops, "normalize", "ops.normalize")
    tmp = tempfile.mkdtemp()
    for rep in range(2):
        T.clear()
        desc = SymmetryFunction(cut_name="cos", cut_dists={"Si-Si": 5.0}, hyperparams="set51", normalize=True)
        desc.store_gradients_on_disk = False
        model = NeuralNetwork(desc)
        model.add_layers(nn.Linear(51, 10), nn.Tanh(), nn.Linear(10, 10), nn.Tanh(), nn.Linear(10, 1))
        calc = CalculatorTorch(model)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        calc.create(confs, fingerprints_filename=os.path.join(tmp, f"fp{rep}.pkl"),
                    fingerprints_mean_stdev_filename=os.path.join(tmp, f"ms{rep}.pkl"))
        torch.cuda.synchronize()
        total = time.perf_counter() - t0
        t1 = time.perf_counter()
        wait_for_fingerprints()
        print(f"rep {rep}: create {total:.3f} s for {ncfg} configurations; writer thread done {time.perf_counter() - t1:.3f} s later")
        for k, v in sorted(T.items(), key=lambda kv: -kv[1]):
            print(f"   {k:45s} {v:.3f} s")


if __name__ == "__main__":
    main()
