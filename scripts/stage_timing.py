"""Per-kernel device times of the descriptor stage (kb_kernel_timing) on a jittered diamond-Si supercell.
Usage: python scripts/stage_timing.py [nx=50] [reps=3]; environment switches as in symfun_split.cu."""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from kliff_b200 import ops
from kliff_b200.legacy.descriptors.hyperparams import get_set51
from scripts.first_light import diamond


def main():
    nx = int(sys.argv[1]) if len(sys.argv) > 1 else 50
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
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
    dz_ptr, total = ops.block_offsets(n, None, offsets, 51)
    dz = torch.empty(total, dtype=torch.float64, device=dev)
    run = lambda: ops.symfun_forward(P, allc, species, offsets, indices, maxn, n_centers=n, dz_ptr=dz_ptr, dz=dz)
    run()
    torch.cuda.synchronize()
    ops.kernel_timing(True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        run()
    e1.record()
    torch.cuda.synchronize()
    t = ops.kernel_timing(False)
    print("atoms", n, "ms/call", e0.elapsed_time(e1) / reps,
          {k: (round(v[0] / reps, 3), v[1] // reps) for k, v in t.items()})


if __name__ == "__main__":
    main()
