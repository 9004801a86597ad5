"""First-light timing of the hot path on one B200 (not the bench contract; see bench.py)."""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from kliff_b200 import ops
from kliff_b200.legacy.descriptors.hyperparams import get_set51


def diamond(nx, a=5.431, sigma=0.05, seed=0):
    basis = np.array([[0, 0, 0], [0, .5, .5], [.5, 0, .5], [.5, .5, 0],
                      [.25, .25, .25], [.25, .75, .75], [.75, .25, .75], [.75, .75, .25]])
    g = np.stack(np.meshgrid(np.arange(nx), np.arange(nx), np.arange(nx), indexing="ij"), -1).reshape(-1, 3)
    pos = (g[:, None, :] + basis[None]).reshape(-1, 3) * a
    pos += np.random.default_rng(seed).normal(0, sigma, pos.shape)
    return np.eye(3) * a * nx, pos


def timed(fn, reps=3):
    torch.cuda.synchronize()
    best = 1e9
    out = None
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = fn()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best, out


def main():
    dev = torch.device("cuda:0")
    print("device", torch.cuda.get_device_name(0), ops.device_info())
    print("fp64 peak TFLOP/s", ops.probe_fp64_peak())
    print("hbm copy GB/s", ops.probe_hbm_copy(1 << 30))
    two_species = os.environ.get("KB_SPECIES", "1") == "2"  # zinc-blende decoration, all cutoffs 5.0 (M1-SiC)
    P = ops.build_params(get_set51(), np.full((2, 2), 5.0) if two_species else np.array([[5.0]]))
    mode = int(os.environ.get("KB_MODE", "0"))  # include/kliff_b200.h KB_SYMFUN_*
    dz_dtype = torch.float32 if os.environ.get("KB_DZ", "f64") == "f32" else torch.float64
    for nx in [int(a) for a in sys.argv[1:]] or [10, 30, 50]:
        cell, pos = diamond(nx)
        n = len(pos)
        coords = torch.tensor(pos, device=dev)
        z = torch.full((n,), 14, dtype=torch.int32, device=dev)
        t_pad, (pc, ps, pi) = timed(lambda: ops.create_paddings(5.0, cell, [1, 1, 1], coords, z))
        allc = torch.cat([coords, pc])
        t_nl, (counts, offsets, indices, maxn) = timed(lambda: ops.build_neighbors(allc, n, 5.0))
        species = torch.zeros(allc.shape[0], dtype=torch.int32, device=dev)
        if two_species:  # basis atoms 0-3 Si, 4-7 C; images inherit their owner's species
            own = torch.cat([torch.arange(n, device=dev), pi.long()])
            species = ((own % 8) >= 4).to(torch.int32)
        dz_ptr, total = ops.block_offsets(n, None, offsets, 51)
        dz = torch.empty(total, dtype=dz_dtype, device=dev)
        t_sf, (zeta, _, _) = timed(lambda: ops.symfun_forward(P, allc, species, offsets, indices, maxn,
                                                              n_centers=n, dz_ptr=dz_ptr, dz=dz, mode=mode))
        t_z, _ = timed(lambda: ops.symfun_forward(P, allc, species, offsets, indices, maxn,
                                                  n_centers=n, grad=False, mode=mode))
        image = torch.cat([torch.arange(n, dtype=torch.int32, device=dev), pi])
        pack = dict(offsets=offsets, indices=indices, image=image, dz=dz, dz_ptr=dz_ptr, D=51, n_out=n)
        w = torch.randn(n, 51, dtype=torch.float64, device=dev)
        t_f, F = timed(lambda: ops._scatter_fwd(w, pack))
        t_b, _ = timed(lambda: ops._scatter_bwd(F, pack))
        print(json.dumps(dict(atoms=n, paddings=int(pc.shape[0]), pairs=int(indices.numel()), maxn=maxn,
                              ms_pad=t_pad, ms_nl=t_nl, ms_symfun_grad=t_sf, ms_symfun_zeta=t_z,
                              ms_scatter_fwd=t_f, ms_scatter_bwd=t_b,
                              atoms_per_s_symfun=n / t_sf * 1e3,
                              dz_GB=total * dz.element_size() / 1e9, symfun_GBps=total * dz.element_size() / t_sf / 1e6,
                              scatter_fwd_GBps=total * dz.element_size() / t_f / 1e6,
                              scatter_bwd_GBps=total * dz.element_size() / t_b / 1e6,
                              species=2 if two_species else 1, dz_dtype=str(dz_dtype))))


if __name__ == "__main__":
    main()
