"""One descriptor pass over the bench workload (1 000 000-atom jittered diamond Si) for ncu traffic
captures.  Usage: python scripts/profile_1m.py [nx]"""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from kliff_b200 import ops
from kliff_b200.legacy.descriptors.hyperparams import get_set51
from scripts.first_light import diamond


def main():
    nx = int(sys.argv[1]) if len(sys.argv) > 1 else 50
    dev = torch.device("cuda:0")
    cell, pos = diamond(nx)
    n = len(pos)
    coords = torch.tensor(pos, device=dev)
    z = torch.full((n,), 14, dtype=torch.int32, device=dev)
    P = ops.build_params(get_set51(), np.array([[5.0]]))
    pc, ps, pi = ops.create_paddings(5.0, cell, [1, 1, 1], coords, z)
    allc = torch.cat([coords, pc])
    counts, offsets, indices, maxn = ops.build_neighbors(allc, n, 5.0)
    species = torch.zeros(allc.shape[0], dtype=torch.int32, device=dev)
    zeta, dz, dz_ptr = ops.symfun_forward(P, allc, species, offsets, indices, maxn, n_centers=n)
    torch.cuda.synchronize()
    print("ok", n, float(zeta.sum()))


if __name__ == "__main__":
    main()
