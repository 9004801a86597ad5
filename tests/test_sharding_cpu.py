"""Host logic of sharding by configuration on CPU: the size-balanced assignment,
the batch-share arithmetic on top of it, and `parmap1` over ranks (gloo, world_size 2) against the
serial evaluation — the reference's contract for it is "the list of f(x, *args) in the order of X"
(kliff/parallel.py:8-101)."""
import os
import socket

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

from kliff_b200.dataset import Configuration
from kliff_b200.parallel import config_cost, parmap1, parmap2, shard_members


def _conf(natoms, a, pbc=True):
    rng = np.random.default_rng(natoms)
    return Configuration(cell=np.eye(3) * a, species=["Si"] * natoms, coords=rng.random((natoms, 3)) * a,
                         PBC=[pbc] * 3)


def test_config_cost_follows_atoms_times_neighbors_squared():
    # same density, 8x the atoms: 8x the cost; same atoms, half the volume: 4x
    assert np.isclose(config_cost(_conf(64, 10.86), 5.0) / config_cost(_conf(8, 5.43), 5.0), 8.0)
    assert np.isclose(config_cost(_conf(8, 5.43), 5.0) / config_cost(_conf(8, 5.43 * 2 ** (1 / 3)), 5.0), 4.0)
    # a non-periodic cluster cannot have more than natoms - 1 neighbors
    assert config_cost(_conf(8, 5.43, pbc=False), 5.0) == 8 * 7.0 ** 2


@pytest.mark.parametrize("world", [1, 2, 3, 8])
def test_shard_members_partition_and_one_member_per_run(world):
    rng = np.random.default_rng(world)
    costs = rng.choice([1.0, 8.0, 64.0], size=37) * rng.uniform(0.9, 1.1, size=37)
    parts = shard_members(costs, world, "balanced")
    assert len(parts) == world
    assert sorted(np.concatenate(parts).tolist()) == list(range(37))
    for p in parts:
        assert np.all(np.diff(p) > 0)
        runs = p // world
        assert len(set(runs.tolist())) == len(p)              # never two members of one run
        assert set(range(37 // world)) <= set(runs.tolist())  # and one of every complete run
    rr = shard_members(costs, world, "round_robin")
    assert all(np.array_equal(p, np.arange(r, 37, world)) for r, p in enumerate(rr))
    if world > 1:
        def worst(ps):
            return max(costs[p].sum() for p in ps)
        assert worst(parts) <= worst(rr)


def test_shard_members_auto_keeps_uniform_sets_strided():
    uniform = np.random.default_rng(0).uniform(1.0, 1.7, size=50)     # the bench's a ~ U(5.2, 5.7) spread
    assert all(np.array_equal(p, np.arange(r, 50, 4)) for r, p in enumerate(shard_members(uniform, 4)))
    mixed = np.r_[np.full(30, 512.0), np.full(10, 1.0)]               # 64-atom and 8-atom cells
    auto = shard_members(mixed, 4)
    assert any(not np.array_equal(p, np.arange(r, 40, 4)) for r, p in enumerate(auto))
    assert all(np.array_equal(a, b) for a, b in zip(auto, shard_members(mixed, 4, "balanced")))
    with pytest.raises(ValueError):
        shard_members(mixed, 4, "bogus")


def test_balanced_assignment_evens_a_skewed_set():
    # heavy configurations all on the stride of rank 0: round-robin gives rank 0 every one of them
    costs = np.ones(64)
    costs[::4] = 50.0
    loads_rr = [costs[p].sum() for p in shard_members(costs, 4, "round_robin")]
    loads_b = [costs[p].sum() for p in shard_members(costs, 4, "balanced")]
    assert max(loads_rr) / min(loads_rr) == 50.0
    assert max(loads_b) / min(loads_b) < 1.3


def test_batches_over_balanced_members_cover_every_configuration_once():
    """The batch-share arithmetic of Loss (`_batches`, `_fused_batches`) on a balanced assignment: for every
    global batch the ranks' shares are disjoint, contiguous in the local lists, together the batch, and
    differ in size by at most one configuration."""
    from kliff_b200.legacy.loss import LossNeuralNetworkModel, _Share
    from test_dist_cpu import _Model, _StubCalculator

    n = 41
    costs = np.random.default_rng(5).choice([1.0, 30.0], size=n)
    for world in (2, 3, 8):
        parts = shard_members(costs, world, "balanced")
        for bs in (10, 8, 5):
            seen, sizes = [], {}
            for rank in range(world):
                local = [{"index": k, "global": int(g)} for k, g in enumerate(parts[rank])]
                calc = _StubCalculator(_Model(), local)
                calc._shard, calc._shard_members = (rank, world, n), parts[rank]
                loss = LossNeuralNetworkModel.__new__(LossNeuralNetworkModel)
                loss.calculator, loss.batch_size = calc, bs
                loss._fast_path = lambda: True
                shares = list(loss._batches(None))
                runs = loss._fused_batches()
                assert len(shares) == len(runs) == (n + bs - 1) // bs
                for b, (share, (c0, c1, nglobal)) in enumerate(zip(shares, runs)):
                    assert isinstance(share, _Share) and share.nglobal == nglobal == min(bs, n - b * bs)
                    assert [s["index"] for s in share] == list(range(c0, c1))
                    assert all(b * bs <= s["global"] < b * bs + bs for s in share)
                    seen += [s["global"] for s in share]
                    sizes.setdefault(b, []).append(len(share))
            assert sorted(seen) == list(range(n))
            if bs % world == 0:   # batch boundaries on run boundaries: equal shares of every full batch
                assert all(max(v) - min(v) == 0 for b, v in sizes.items() if (b + 1) * bs <= n - n % world)
            assert all(max(v) - min(v) <= 2 for v in sizes.values())


class _Desc:
    cutoff = {"Si-Si": 5.0}


class _M:
    descriptor = _Desc()


def test_calculator_chooses_members_and_reuse_reads_them_back(tmp_path):
    from kliff_b200.legacy.calculators.calculator_torch import CalculatorTorch, _members_file, _rank_suffix

    calc = CalculatorTorch.__new__(CalculatorTorch)
    calc._model, calc._shard_members = _M(), None
    uniform = [_conf(64, 10.86) for _ in range(12)]
    assert list(calc._choose_members(uniform, 1, 4, None)) == [1, 5, 9] and calc._shard_members is None
    mixed = [_conf(64, 10.86) if i % 4 == 0 else _conf(8, 5.43) for i in range(12)]
    mine = calc._choose_members(mixed, 1, 4, True)
    assert calc._shard_members is not None and list(calc._shard_members) == list(mine)
    calc._shard = (1, 4, 12)
    assert [calc._shard_below(lo) for lo in (0, 4, 8, 12)] == [0, 1, 2, 3]
    calc._shard_members = None
    assert [calc._shard_below(lo) for lo in (0, 1, 2, 5, 6, 12)] == [0, 0, 1, 1, 2, 3]
    assert list(calc._choose_members(mixed, 1, 4, "round_robin")) == [1, 5, 9] and calc._shard_members is None
    big = [_conf(8, 5.43 + 0.0001 * (i % 7)) for i in range(3000)]   # decided on a strided sample
    assert list(calc._choose_members(big, 2, 8, None)) == list(range(2, 3000, 8)) and calc._shard_members is None
    big[5::3] = [_conf(64, 10.86)] * len(big[5::3])
    mine = calc._choose_members(big, 2, 8, None)
    assert calc._shard_members is not None and len(mine) == 3000 // 8
    f = _rank_suffix(tmp_path / "fingerprints.pkl", 3)
    assert f.name == "fingerprints.rank3.pkl" and _members_file(f).name == "fingerprints.rank3.members.npy"


# ---------------------------------------------------------------------------------------------- parmap1
def _f(x, y, z=1):
    return x * 10 + y + z


def test_parmap1_single_process_matches_the_reference_examples():
    # the docstring examples of kliff/parallel.py:51-60 (func = x + y + z)
    def func(x, y, z=1):
        return x + y + z
    assert parmap1(func, range(3), 1, nprocs=2) == [2, 3, 4]
    assert parmap1(func, range(3), 1, 1, nprocs=2) == [2, 3, 4]
    assert parmap1(func, zip(range(3), range(3)), tuple_X=True, nprocs=2) == [1, 3, 5]
    assert parmap1(func, zip(range(3), range(3)), 1, tuple_X=True, nprocs=2) == [1, 3, 5]
    assert parmap2(func, range(3), 1) == [2, 3, 4]
    assert parmap1(func, [], 1) == []


def _rank_main(rank, world, port, queue):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        calls = []

        def traced(x, y):
            calls.append(x)
            return _f(x, y)
        out = parmap1(traced, range(7), 3)
        out_t = parmap1(_f, [(i, i + 1) for i in range(5)], tuple_X=True)
        queue.put((rank, out, out_t, calls))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_parmap1_two_ranks_split_the_items_and_return_the_full_list():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    queue = ctx.Queue()
    procs = [ctx.Process(target=_rank_main, args=(r, 2, port, queue)) for r in range(2)]
    for p in procs:
        p.start()
    got = sorted(queue.get(timeout=240) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, out, out_t, calls in got:
        assert out == [_f(x, 3) for x in range(7)]
        assert out_t == [_f(i, i + 1) for i in range(5)]
        assert calls == list(range(rank, 7, 2))      # each rank evaluated only its own items
