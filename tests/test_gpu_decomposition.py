"""Single large supercell split by spatial domain with ghost halos (BASELINE.json config 4),
emulating W ranks on one GPU with the in-process transport: descriptors of every owned atom and
the total forces after the reverse halo exchange equal the undivided computation (fp64, 1e-10)."""
import numpy as np
import pytest
import torch

from kliff_b200.dataset import Configuration
from kliff_b200.legacy.descriptors import SymmetryFunction
from kliff_b200.parallel import InProcessTransport, SlabDecomposition

pytestmark = pytest.mark.gpu


def supercell(nx, ny, nz, a=5.431, sigma=0.05, seed=3):
    basis = np.array([[0, 0, 0], [0, .5, .5], [.5, 0, .5], [.5, .5, 0],
                      [.25, .25, .25], [.25, .75, .75], [.75, .25, .75], [.75, .75, .25]])
    g = np.stack(np.meshgrid(np.arange(nx), np.arange(ny), np.arange(nz), indexing="ij"), -1).reshape(-1, 3)
    pos = (g[:, None, :] + basis[None]).reshape(-1, 3) * a
    pos += np.random.default_rng(seed).normal(0, sigma, pos.shape)
    return np.diag([a * nx, a * ny, a * nz]), pos


@pytest.mark.parametrize("world", [1, 2, 4])
def test_slab_decomposition_matches_global(world):
    dev = torch.device("cuda:0")
    cell, pos = supercell(8, 2, 2)  # 256 atoms, 43.4 A along a1 -> slabs of >= 10.9 A at W = 4
    n = len(pos)
    desc = SymmetryFunction({"Si-Si": 5.0}, "cos", "set51", normalize=False, dtype=np.float64)
    conf = Configuration(cell=cell, species=["Si"] * n, coords=pos, PBC=[True] * 3)
    glob = desc.transform_packed([conf], device=dev)
    rng = np.random.default_rng(0)
    w = torch.tensor(rng.normal(size=(n, 51)), device=dev)
    F_ref = glob.forces(w, 0, n).reshape(n, 3)

    tr = InProcessTransport()
    dds = [SlabDecomposition(cell, [1, 1, 1], 5.0, rank=r, world=world, transport=tr) for r in range(world)]
    x_all = dds[0].wrap(torch.tensor(pos, device=dev))
    owner = dds[0].owner(x_all)
    code_all = torch.zeros(n, dtype=torch.int32, device=dev)
    mine = [torch.nonzero(owner == r, as_tuple=False).reshape(-1) for r in range(world)]
    assert sum(len(m) for m in mine) == n
    for dd, m in zip(dds, mine):
        dd.post_halo(x_all[m], code_all[m])
    fps = []
    for dd, m in zip(dds, mine):
        gx, gc = dd.collect_halo()
        fp = dd.local_fingerprints(desc, x_all[m], code_all[m], gx, gc)
        assert torch.allclose(fp["zeta"], glob.zeta[m], rtol=1e-10, atol=0)
        f_local = torch.zeros(0)
        from kliff_b200 import ops
        f_local = ops.force_scatter(w[m], fp["pack"]).reshape(fp["n_local"], 3)
        dd.post_ghost_forces(f_local, len(m))
        fps.append(fp)
    F = torch.zeros_like(F_ref)
    for dd, m in zip(dds, mine):
        F[m] = dd.collect_ghost_forces()
    assert float((F - F_ref).abs().max()) <= 1e-10 * float(F_ref.abs().max())
    # every slab holds only its share (+ halo), not the whole cell
    if world > 1:
        assert max(fp["n_local"] for fp in fps) < n
