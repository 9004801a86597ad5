"""world_size-2 gloo test (CPU) of the data-parallel host logic in kliff_b200.legacy.loss:
every rank takes a contiguous share of each batch, gradients and loss are all-reduced, and the
2-rank trajectory equals the single-process one.  The calculator here is a CPU test double that
serves dense fingerprints through the reference's own tensordot formulation
(CalculatorTorch._compute_dense) — the product kernels are GPU-only and are not exercised."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from kliff_b200.dataset import Configuration
from kliff_b200.dataset.weight import Weight
from kliff_b200.legacy.calculators.calculator_torch import CalculatorTorch
from kliff_b200.legacy.loss import Loss


class _Desc:
    dtype = np.float64
    normalize = False

    def get_size(self):
        return 6

    def get_dtype(self):
        return self.dtype

    def state_dict(self):
        return {}


class _Dataset:
    def __init__(self, fp):
        self.fp = fp

    def __len__(self):
        return len(self.fp)

    def __getitem__(self, i):
        return self.fp[i]


class _StubCalculator(CalculatorTorch):
    def __init__(self, model, samples):
        self._model = model
        self.dtype = np.float64
        self.use_energy, self.use_forces, self.use_stress = True, True, False
        self.results = {}
        self._packed = None
        self._ref = None
        self.fingerprints_dataset = _Dataset(samples)

    def save_model(self, epoch, force_save=False):
        pass


def _samples(seed=0, ncfg=10):
    rng = np.random.default_rng(seed)
    out = []
    for i in range(ncfg):
        n = int(rng.integers(2, 5))
        conf = Configuration(cell=np.eye(3), species=["Si"] * n, coords=rng.random((n, 3)),
                             PBC=[True] * 3, energy=float(rng.normal()), forces=rng.normal(size=(n, 3)),
                             weight=Weight(forces_weight=0.3), identifier=f"c{i}")
        out.append({"configuration": conf, "index": i,
                    "zeta": rng.normal(size=(n, 6)), "energy": np.asarray(conf.energy),
                    "dzetadr_forces": rng.normal(size=(n, 6, 3 * n)), "forces": conf.forces})
    return out


class _Model(torch.nn.Module):
    def __init__(self):
        super().__init__()
        torch.manual_seed(3)
        self.net = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.Tanh(), torch.nn.Linear(5, 1)).double()
        self.descriptor = _Desc()
        self.dtype = torch.float64

    @property
    def device(self):
        return next(self.parameters()).device

    def forward(self, x):
        return self.net(x)


def _train():
    calc = _StubCalculator(_Model(), _samples())
    loss = Loss(calc)
    loss.minimize(method="Adam", num_epochs=3, batch_size=4, lr=0.01)
    params = torch.cat([p.detach().reshape(-1) for p in calc.model.parameters()])
    return np.array(loss.epoch_losses), params.numpy()


def _rank_main(rank, world, port, queue):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        losses, params = _train()
        queue.put((rank, losses, params))
    finally:
        dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.timeout(300)
def test_two_rank_trajectory_equals_single_process():
    ref_losses, ref_params = _train()
    ctx = mp.get_context("spawn")
    queue = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_rank_main, args=(r, 2, port, queue)) for r in range(2)]
    for p in procs:
        p.start()
    got = [queue.get(timeout=240) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, losses, params in got:
        assert np.allclose(losses, ref_losses, rtol=1e-10), (rank, losses, ref_losses)
        assert np.allclose(params, ref_params, rtol=1e-9, atol=1e-12)
    assert np.array_equal(got[0][2], got[1][2])  # ranks stay bit-identical to each other


def test_runs_helper_and_share():
    from kliff_b200.legacy.calculators.calculator_torch import _runs
    assert _runs([0, 1, 2, 5, 6, 9]) == [(0, 3, 0), (5, 7, 3), (9, 10, 5)]
    assert _runs([]) == []


def test_sharded_batches_cover_every_configuration_once():
    """Host logic of sharded fingerprints: rank r keeps configurations r, r + w, ...; for every global
    batch the ranks' shares are disjoint, contiguous in the local lists, and together the batch."""
    from kliff_b200.legacy.loss import LossNeuralNetworkModel, _Share

    n, bs = 37, 10
    for world in (1, 2, 3, 8):
        seen = []
        for rank in range(world):
            local = [{"index": k, "global": g} for k, g in enumerate(range(rank, n, world))]
            calc = _StubCalculator(_Model(), local)
            calc._shard = (rank, world, n)
            loss = LossNeuralNetworkModel.__new__(LossNeuralNetworkModel)
            loss.calculator, loss.batch_size = calc, bs
            loss._fast_path = lambda: True
            shares = list(loss._batches(None))
            assert len(shares) == (n + bs - 1) // bs          # every rank walks every global batch
            for b, share in enumerate(shares):
                assert isinstance(share, _Share) and share.nglobal == min(bs, n - b * bs)
                assert all(b * bs <= s["global"] < b * bs + bs for s in share)
                idx = [s["index"] for s in share]
                assert idx == list(range(idx[0], idx[0] + len(idx))) if idx else True
                seen += [s["global"] for s in share]
        assert sorted(seen) == list(range(n))
