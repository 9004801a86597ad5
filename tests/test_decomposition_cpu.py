"""Spatial slab decomposition with ghost halos (kliff_b200.parallel) checked on the CPU against
the oracle: for W emulated ranks (in-process transport) and for 2 real gloo ranks, a centre-based
pair energy E = sum_i sum_{j in N(i)} phi(r_ij)/2 evaluated per slab — owned centres, ghosts from
the neighbours, periodic images along a2/a3 from the oracle's padding routine, partial forces on
ghosts sent back to their owners — reproduces the forces of the undivided periodic cell."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from kliff_b200.parallel import DistTransport, InProcessTransport, SlabDecomposition
from oracle import oracle as orc

RC = 3.2


def system(seed=4, n=90):
    rng = np.random.default_rng(seed)
    cell = np.array([[14.0, 0.4, 0.0], [0.9, 6.1, 0.3], [0.2, 0.5, 6.6]])
    g = np.stack(np.meshgrid(np.arange(9), np.arange(4), np.arange(4), indexing="ij"), -1).reshape(-1, 3)
    frac = (g + 0.5) / np.array([9, 4, 4]) + rng.normal(0, 0.02, (len(g), 3))
    coords = frac @ cell + 3.0 * cell[0] * (rng.random(len(g)) < 0.2)[:, None]  # some atoms unwrapped
    return cell, coords, (np.arange(len(g)) % 2).astype(np.int32)


def pair_forces(port, cell, pbc, local, n_centres):
    """Forces on the `local` atoms (first n_centres are centres) of E = sum_i sum_j (r_ij - RC)^2 / 2,
    periodic images per `pbc`, image atoms folded onto their local owners."""
    z = np.ones(len(local), np.int32)
    pc, _, pi = port.create_paddings(RC, cell, pbc, local, z)
    allc = np.concatenate([local, pc]) if len(pc) else local
    image = np.concatenate([np.arange(len(local)), pi]).astype(np.int64)
    need = np.zeros(len(allc), np.int32)
    need[:n_centres] = 1
    counts, off, idx = port.build_neighbors(allc, RC, need)
    F = np.zeros((len(local), 3))
    energy = 0.0
    for i in range(n_centres):
        nb = idx[off[i]:off[i + 1]]
        d = allc[nb] - allc[i]
        r = np.linalg.norm(d, axis=1)
        energy += 0.5 * np.sum((r - RC) ** 2)
        g = (r - RC)[:, None] * d / r[:, None]  # dE/dx_j ; dE/dx_i = -sum
        np.subtract.at(F, image[nb], g)
        F[i] += g.sum(axis=0)
    return F, energy


def global_reference(port):
    cell, coords, code = system()
    dd = SlabDecomposition(cell, [1, 1, 1], RC, rank=0, world=1, transport=InProcessTransport())
    wrapped = dd.wrap(torch.tensor(coords)).numpy()
    F, E = pair_forces(port, cell, [1, 1, 1], wrapped, len(wrapped))
    return cell, wrapped, code, F, E


def run_rank(port, dd, cell, wrapped, code, owner):
    mine = np.nonzero(owner == dd.rank)[0]
    x = torch.tensor(wrapped[mine])
    c = torch.tensor(code[mine])
    return mine, x, c


@pytest.mark.parametrize("world", [1, 2, 3])
def test_slab_decomposition_in_process(world):
    port = orc.Oracle("port")
    cell, wrapped, code, F_ref, E_ref = global_reference(port)
    tr = InProcessTransport()
    dds = [SlabDecomposition(cell, [1, 1, 1], RC, rank=r, world=world, transport=tr) for r in range(world)]
    owner = dds[0].owner(torch.tensor(wrapped)).numpy()
    state = []
    for dd in dds:
        mine, x, c = run_rank(port, dd, cell, wrapped, code, owner)
        dd.post_halo(x, c)
        state.append((mine, x, c))
    partial, energy = [], 0.0
    for dd, (mine, x, c) in zip(dds, state):
        gx, gc = dd.collect_halo()
        # ghosts carry the species of their owners
        local = np.concatenate([x.numpy(), gx.numpy()])
        F_loc, e = pair_forces(port, cell, [0, 1, 1], local, len(mine))
        energy += e
        partial.append(F_loc)
        dd.post_ghost_forces(torch.tensor(F_loc), len(mine))
    F = np.zeros_like(F_ref)
    for dd, (mine, x, c) in zip(dds, state):
        F[mine] = dd.collect_ghost_forces().numpy()
    assert abs(energy - E_ref) <= 1e-10 * abs(E_ref)
    assert np.max(np.abs(F - F_ref)) <= 1e-10 * np.max(np.abs(F_ref))
    assert not tr.box  # every message was consumed


def test_slab_too_thin_is_rejected():
    cell, _, _ = system()
    with pytest.raises(ValueError, match="thinner than the halo"):
        SlabDecomposition(cell, [1, 1, 1], RC, rank=0, world=8, transport=InProcessTransport())


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _gloo_rank(rank, world, port_no, queue):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port_no)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        port = orc.Oracle("port")
        cell, wrapped, code, F_ref, E_ref = global_reference(port)
        dd = SlabDecomposition(cell, [1, 1, 1], RC)  # rank / world / transport from torch.distributed
        assert isinstance(dd.transport, DistTransport) and dd.world == world
        owner = dd.owner(torch.tensor(wrapped)).numpy()
        mine, x, c = run_rank(port, dd, cell, wrapped, code, owner)
        gx, gc = dd.exchange_halo(x, c)
        local = np.concatenate([x.numpy(), gx.numpy()])
        F_loc, e = pair_forces(port, cell, [0, 1, 1], local, len(mine))
        F_own = dd.reduce_ghost_forces(torch.tensor(F_loc), len(mine)).numpy()
        et = torch.tensor([e], dtype=torch.float64)
        dist.all_reduce(et)
        err = float(np.max(np.abs(F_own - F_ref[mine])) / np.max(np.abs(F_ref)))
        queue.put((rank, err, abs(float(et) - E_ref) / abs(E_ref), int(gx.shape[0])))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_slab_decomposition_two_gloo_ranks():
    ctx = mp.get_context("spawn")
    queue = ctx.Queue()
    port_no = _free_port()
    procs = [ctx.Process(target=_gloo_rank, args=(r, 2, port_no, queue)) for r in range(2)]
    for p in procs:
        p.start()
    got = [queue.get(timeout=240) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, ferr, eerr, nghost in got:
        assert ferr <= 1e-10 and eerr <= 1e-10 and nghost > 0, (rank, ferr, eerr, nghost)
