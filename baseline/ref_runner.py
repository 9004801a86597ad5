"""Runs the UNMODIFIED reference (baseline/_ref/kliff_src + the import stubs in baseline/stubs) on the
host cores and prints one JSON line.  Executed by bench.py as a subprocess with
PYTHONPATH=baseline/_ref/kliff_src:baseline/stubs so that the reference's `kliff` package never
shares a process with kliff_b200.

    python baseline/ref_runner.py descriptor --nx 3 --configs 32 --nprocs 16
        kliff.parallel.parmap1(SymmetryFunction.transform, configs, fit_forces=True, fit_stress=False)
        — the call Descriptor.generate_fingerprints makes (kliff/legacy/descriptors/descriptor.py:234-236) —
        over `configs` jittered diamond-Si cells of nx^3 conventional cells, set51, cos 5.0 A.
    python baseline/ref_runner.py epoch --configs 200 --nprocs 16
        CalculatorTorch.create (descriptors through the same parmap1) + Loss.minimize(Adam, batch 100) on
        CPU torch (kliff/legacy/loss.py:740-828), MLP 51-10-10-1, the synthetic 64-atom set of bench.py.
"""
import argparse
import contextlib
import io
import json
import os
import sys
import tempfile
import time

import numpy as np

A_SI, SIGMA, RC = 5.431, 0.05, 5.0
BASIS = np.array([[0, 0, 0], [0, .5, .5], [.5, 0, .5], [.5, .5, 0],
                  [.25, .25, .25], [.25, .75, .75], [.75, .25, .75], [.75, .75, .25]])


def diamond(nx, seed, a=A_SI, sigma=SIGMA):
    g = np.stack(np.meshgrid(np.arange(nx), np.arange(nx), np.arange(nx), indexing="ij"), -1).reshape(-1, 3)
    pos = (g[:, None, :] + BASIS[None]).reshape(-1, 3) * a
    pos += np.random.default_rng(seed).normal(0.0, sigma, pos.shape)
    return np.eye(3) * a * nx, np.ascontiguousarray(pos)


def run_descriptor(args):
    from kliff import parallel
    from kliff.dataset import Configuration
    from kliff.legacy.descriptors import SymmetryFunction

    desc = SymmetryFunction(cut_name="cos", cut_dists={"Si-Si": RC}, hyperparams="set51", normalize=False)
    confs = []
    for c in range(args.configs):
        cell, pos = diamond(args.nx, c)
        confs.append(Configuration(cell=cell, species=["Si"] * len(pos), coords=pos, PBC=[True] * 3,
                                   energy=0.0, forces=np.zeros_like(pos)))
    n = len(pos)
    secs = []
    for _ in range(args.repeat):
        t0 = time.time()
        if args.nprocs == 1:
            out = [desc.transform(c, True, False) for c in confs]
        else:
            out = parallel.parmap1(desc.transform, confs, True, False, nprocs=args.nprocs)
        secs.append(time.time() - t0)
    z = out[0][0]
    return {"what": "kliff.parallel.parmap1(SymmetryFunction.transform, configs, fit_forces=True, fit_stress=False)",
            "configs": args.configs, "atoms_per_config": n, "atoms": n * args.configs, "nprocs": args.nprocs,
            "seconds": secs, "atoms_per_s": n * args.configs * len(secs) / sum(secs), "zeta_shape": list(z.shape),
            "dzetadr_shape": list(out[0][1].shape)}


def run_epoch(args):
    import torch
    from kliff.dataset import Configuration
    from kliff.dataset.weight import Weight
    from kliff.legacy import nn
    from kliff.legacy.calculators import CalculatorTorch
    from kliff.legacy.descriptors import SymmetryFunction
    from kliff.legacy.loss import Loss
    from kliff.models import NeuralNetwork

    torch.set_num_threads(args.nprocs)
    g = np.stack(np.meshgrid(np.arange(2), np.arange(2), np.arange(2), indexing="ij"), -1).reshape(-1, 3)
    frac = (g[:, None, :] + BASIS[None]).reshape(-1, 3)
    weight = Weight(forces_weight=0.3)
    confs = []
    for c in range(args.configs):
        rng = np.random.default_rng(c)
        a = rng.uniform(5.2, 5.7)
        pos = frac * a + rng.normal(0, 0.1, frac.shape)
        confs.append(Configuration(cell=np.eye(3) * 2 * a, species=["Si"] * 64, coords=pos, PBC=[True] * 3,
                                   energy=float(rng.normal()), forces=rng.normal(size=(64, 3)),
                                   weight=weight, identifier=f"syn{c}"))
    desc = SymmetryFunction(cut_name="cos", cut_dists={"Si-Si": RC}, hyperparams="set51", normalize=True)
    model = NeuralNetwork(desc)
    model.add_layers(nn.Linear(51, 10), nn.Tanh(), nn.Linear(10, 10), nn.Tanh(), nn.Linear(10, 1))
    tmp = tempfile.mkdtemp()
    model.set_save_metadata(prefix=os.path.join(tmp, "m"), start=10 ** 9, frequency=10 ** 9)
    calc = CalculatorTorch(model, gpu=False)
    t0 = time.time()
    calc.create(confs, nprocs=args.nprocs, fingerprints_filename=os.path.join(tmp, "fp.pkl"),
                fingerprints_mean_stdev_filename=os.path.join(tmp, "ms.pkl"))
    t_create = time.time() - t0
    loss = Loss(calc)
    with contextlib.redirect_stdout(io.StringIO()):
        t0 = time.time()
        loss.minimize(method="Adam", num_epochs=args.epochs, batch_size=100, lr=1e-3)
        t_run = time.time() - t0
    # minimize() runs num_epochs optimisation passes + one final loss pass (loss.py:771-828)
    return {"what": "CalculatorTorch(gpu=False).create(nprocs) + Loss.minimize(Adam, batch 100), CPU torch",
            "configs": args.configs, "atoms": 64 * args.configs, "nprocs": args.nprocs, "create_s": t_create,
            "epochs": args.epochs, "minimize_s": t_run, "epoch_s": t_run / (args.epochs + 1)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("what", choices=["descriptor", "epoch"])
    ap.add_argument("--nx", type=int, default=3)
    ap.add_argument("--configs", type=int, default=16)
    ap.add_argument("--nprocs", type=int, default=len(os.sched_getaffinity(0)))
    ap.add_argument("--epochs", type=int, default=1)
    ap.add_argument("--repeat", type=int, default=1, help="descriptor: timed repetitions over the same set")
    args = ap.parse_args()
    import warnings
    warnings.filterwarnings("ignore")
    import logging
    logging.disable(logging.CRITICAL)
    try:
        from loguru import logger
        logger.remove()
    except Exception:
        pass
    out = run_descriptor(args) if args.what == "descriptor" else run_epoch(args)
    sys.stdout.flush()
    print("REF_JSON " + json.dumps(out))


if __name__ == "__main__":
    main()
