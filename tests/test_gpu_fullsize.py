"""Parity at BASELINE.json's full size (configs[1]: 1 000 000-atom diamond-Si supercell, set51):
bit-exact paddings and neighbor lists against the oracle for the whole system, descriptor + dG/dr
blocks against the oracle on a fixed random sample plus atoms next to the cell faces
(SURVEY.md §8d), and size-independent properties over ALL atoms: translation invariance of every
(atom, descriptor) row, Newton's third law of the assembled forces, adjointness of the two
scatter kernels."""
import numpy as np
import pytest
import torch

from kliff_b200 import ops
from kliff_b200.legacy.descriptors.hyperparams import get_set51
from oracle import oracle as orc

pytestmark = pytest.mark.gpu

A_SI, RC, NX = 5.431, 5.0, 50


def diamond(nx, seed=0):
    basis = np.array([[0, 0, 0], [0, .5, .5], [.5, 0, .5], [.5, .5, 0],
                      [.25, .25, .25], [.25, .75, .75], [.75, .25, .75], [.75, .75, .25]])
    g = np.stack(np.meshgrid(np.arange(nx), np.arange(nx), np.arange(nx), indexing="ij"), -1).reshape(-1, 3)
    pos = (g[:, None, :] + basis[None]).reshape(-1, 3) * A_SI
    pos += np.random.default_rng(seed).normal(0.0, 0.05, pos.shape)
    return np.eye(3) * A_SI * nx, np.ascontiguousarray(pos)


@pytest.mark.timeout(900)
def test_one_million_atoms():
    dev = torch.device("cuda:0")
    free, _ = torch.cuda.mem_get_info()
    if free < 90 << 30:
        pytest.skip("needs ~80 GB of device memory")
    cell, pos = diamond(NX)
    n = len(pos)
    assert n == 1_000_000
    port = orc.Oracle("port")
    z = np.full(n, 14, np.int32)

    # ---- paddings + neighbor list: the whole system, bit for bit
    coords = torch.tensor(pos, device=dev)
    pc, ps, pi = ops.create_paddings(RC, cell, [1, 1, 1], coords, torch.tensor(z, device=dev))
    opc, ops_, opi = port.create_paddings(RC, cell, [1, 1, 1], pos, z)
    assert pc.shape[0] == 92727
    assert np.array_equal(pc.cpu().numpy(), opc) and np.array_equal(pi.cpu().numpy(), opi)
    allc_h = np.concatenate([pos, opc])
    allc = torch.cat([coords, pc])
    counts, offsets, indices, maxn = ops.build_neighbors(allc, n, RC)
    need = np.zeros(len(allc_h), np.int32)
    need[:n] = 1
    oc, oo, oi = port.build_neighbors(allc_h, RC, need)
    assert np.array_equal(counts.cpu().numpy(), oc) and np.array_equal(offsets.cpu().numpy(), oo)
    assert maxn == 28 and int(oc[:n].min()) == 28 and indices.numel() == 28 * n
    oi_sorted = np.sort(oi.reshape(n, 28), axis=1).reshape(-1)  # canonical (ascending) order
    assert np.array_equal(indices.cpu().numpy(), oi_sorted)

    # ---- descriptor + dG/dr for all atoms
    hp = get_set51()
    P = ops.build_params(hp, np.array([[RC]]))
    D = 51
    code = torch.zeros(allc.shape[0], dtype=torch.int32, device=dev)
    zeta, dz, dz_ptr = ops.symfun_forward(P, allc, code, offsets, indices, maxn, n_centers=n)
    stride = int(dz_ptr[1] - dz_ptr[0])
    assert stride == (3 * D * 29 + 15) // 16 * 16 and dz.numel() == stride * n

    # sample: 1 500 random atoms + 500 atoms closest to a cell face
    rng = np.random.default_rng(0)
    edge = np.minimum(pos, A_SI * NX - pos).min(axis=1)
    sample = np.unique(np.concatenate([rng.choice(n, 1500, replace=False), np.argsort(edge)[:500]]))
    fams = orc.fams_from_hyperparams(hp)
    zsel = zeta[torch.as_tensor(sample, device=dev)].cpu().numpy()
    ocode = np.zeros(len(allc_h), np.int32)
    worst_z = worst_g = 0.0
    for row, a in enumerate(sample):
        nb = oi_sorted[28 * a:28 * a + 28]
        oz, og = port.symfun_atom(np.array([[RC]]), fams, int(a), allc_h, ocode, nb)
        worst_z = max(worst_z, float(np.max(np.abs(zsel[row] - oz) / np.abs(oz))))
        blk = dz[a * stride:a * stride + 3 * D * 29].cpu().numpy().reshape(D, 29, 3)
        scale = np.abs(og).reshape(D, -1).max(axis=1)
        worst_g = max(worst_g, float(np.max(np.abs(blk - og).reshape(D, -1).max(axis=1) / scale)))
    assert worst_z <= 1e-10 and worst_g <= 1e-10, (worst_z, worst_g)

    # ---- properties over ALL atoms
    view = dz.view(n, stride)
    worst_inv = 0.0
    for lo in range(0, n, 125_000):  # translation invariance: slots of every row sum to zero
        blk = view[lo:lo + 125_000, :3 * D * 29].reshape(-1, D, 29, 3)
        resid = blk.sum(dim=2).abs().amax(dim=2)
        scale = blk.abs().amax(dim=(2, 3))
        worst_inv = max(worst_inv, float((resid / scale).max()))
        assert float(view[lo:lo + 125_000, 3 * D * 29:].abs().max()) == 0.0  # alignment padding is zero
    assert worst_inv <= 1e-12, worst_inv

    image = torch.cat([torch.arange(n, dtype=torch.int32, device=dev), pi])
    pack = dict(offsets=offsets, indices=indices, image=image, dz=dz, dz_ptr=dz_ptr, D=D, n_out=n)
    w = torch.randn(n, D, dtype=torch.float64, device=dev, generator=torch.Generator(dev).manual_seed(1))
    F = ops._scatter_fwd(w, pack)
    # Newton's third law: total force vanishes (relative to the sum of magnitudes)
    tot = F.view(n, 3).sum(dim=0).abs().max()
    assert float(tot) <= 1e-9 * float(F.abs().sum())
    gF = torch.randn(3 * n, dtype=torch.float64, device=dev, generator=torch.Generator(dev).manual_seed(2))
    gw = ops._scatter_bwd(gF, pack)
    lhs, rhs = float((gF * F).sum()), float((gw * w).sum())
    assert abs(lhs - rhs) <= 1e-10 * abs(lhs)
