"""Where the end-to-end time of bench.py's e2e leg goes (1M-atom supercell): synchronised phases."""
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from kliff_b200.dataset import Configuration
from kliff_b200.legacy.descriptors.symmetry_function import SymmetryFunction
from kliff_b200.neighbor import NeighborList
from kliff_b200.packed import PackedFingerprints
from scripts.first_light import diamond


def main():
    nx = int(sys.argv[1]) if len(sys.argv) > 1 else 50
    dev = torch.device("cuda:0")
    cell, pos = diamond(nx)
    n = len(pos)
    desc = SymmetryFunction({"Si-Si": 5.0}, "cos", "set51")
    D = len(desc)
    host_pos = torch.from_numpy(pos).pin_memory()
    conf = Configuration(cell=cell, species=np.full(n, "Si"), coords=host_pos.numpy(), PBC=[True] * 3)
    w = torch.randn(n, D, dtype=torch.float64, device=dev)
    zeta_host = torch.empty((n, D), dtype=torch.float64).pin_memory()
    f_host = torch.empty(3 * n, dtype=torch.float64).pin_memory()

    def tick(label, t0, acc):
        torch.cuda.synchronize()
        t1 = time.time()
        acc[label] = acc.get(label, 0.0) + (t1 - t0) * 1e3
        return t1

    for rep in range(3):
        acc = {}
        torch.cuda.synchronize()
        t = time.time()
        part, nei = desc._config_part(conf, dev)
        t = tick("neighbor list + species (H2D incl.)", t, acc)
        packed = PackedFingerprints.from_parts(D, dev, [part])
        t = tick("from_parts", t, acc)
        packed.compute(desc._params, grad=True, zeta_host=zeta_host)
        t = tick("compute (+ streamed zeta D2H)", t, acc)
        F = packed.forces(w, 0, n)
        t = tick("force contraction", t, acc)
        f_host.copy_(F, non_blocking=True)
        t = tick("forces D2H", t, acc)
        del packed, F, part, nei
        t = tick("free", t, acc)
        print(rep, {k: round(v, 2) for k, v in acc.items()}, "total", round(sum(acc.values()), 2))


if __name__ == "__main__":
    main()
