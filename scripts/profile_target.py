"""Small fixed workload for ncu: one pass of the hot path on a jittered diamond-Si supercell
(nx^3 conventional cells), set51, rc 5.  Usage: python scripts/profile_target.py [nx] [reps]"""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from kliff_b200 import ops
from kliff_b200.legacy.descriptors.hyperparams import get_set51
from scripts.first_light import diamond


def main():
    nx = int(sys.argv[1]) if len(sys.argv) > 1 else 16
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
    dev = torch.device("cuda:0")
    cell, pos = diamond(nx)
    n = len(pos)
    coords = torch.tensor(pos, device=dev)
    z = torch.full((n,), 14, dtype=torch.int32, device=dev)
    P = ops.build_params(get_set51(), np.array([[5.0]]))
    for _ in range(reps):
        pc, ps, pi = ops.create_paddings(5.0, cell, [1, 1, 1], coords, z)
        allc = torch.cat([coords, pc])
        counts, offsets, indices, maxn = ops.build_neighbors(allc, n, 5.0)
        species = torch.zeros(allc.shape[0], dtype=torch.int32, device=dev)
        zeta, dz, dz_ptr = ops.symfun_forward(P, allc, species, offsets, indices, maxn, n_centers=n, mode=int(__import__("os").environ.get("KB_MODE", "0")))
        image = torch.cat([torch.arange(n, dtype=torch.int32, device=dev), pi])
        pack = dict(offsets=offsets, indices=indices, image=image, dz=dz, dz_ptr=dz_ptr, D=51, n_out=n)
        w = torch.randn(n, 51, dtype=torch.float64, device=dev)
        F = ops._scatter_fwd(w, pack)
        g = ops._scatter_bwd(F, pack)
    torch.cuda.synchronize()
    print("ok", n, float(zeta.sum()), float(F.abs().max()), float(g.abs().max()))


if __name__ == "__main__":
    main()
