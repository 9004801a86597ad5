"""Small fixed workload for ncu: a few fused training steps (1 000 synthetic 64-atom configurations,
batch 100, MLP 51-10-10-1, fp32), CUDA graphs off so that every launch is visible.
Usage: KB_NO_GRAPHS=1 python scripts/profile_train.py [n_configs]"""
import contextlib
import io
import os
import sys
import tempfile

import numpy as np
import torch

sys.path.insert(0, ".")
import bench  # noqa: E402


def breakdown():
    """KB_TRAIN_BREAKDOWN=1: CUDA events around every library call of the eager steps -> mean us per entry point."""
    from kliff_b200 import _lib
    L = _lib.lib()
    names = ["kb_train_forward", "kb_train_residual_seg", "kb_train_backward", "kb_train_update", "kb_train_adam",
             "kb_mlp_energy_grad", "kb_force_scatter_fwd_seg", "kb_train_residual", "kb_force_scatter_bwd",
             "kb_mlp_param_grad", "kb_adam_step"]
    rec = {n: [] for n in names}

    class Timed:
        def __init__(self, name, fn):
            self.name, self.fn = name, fn

        def __call__(self, *a):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            r = self.fn(*a)
            e1.record()
            rec[self.name].append((e0, e1))
            return r

    class Proxy:
        def __getattr__(self, k):
            f = getattr(L, k)
            return Timed(k, f) if k in rec else f

    proxy = Proxy()
    _lib.lib = lambda: proxy
    return rec


def main():
    ncfg = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
    os.environ.setdefault("KB_NO_GRAPHS", "1")
    dev = torch.device("cuda:0")
    torch.cuda.set_device(0)
    rec = breakdown() if os.environ.get("KB_TRAIN_BREAKDOWN") == "1" else None
    out = bench.train_epoch_probe(ncfg, dev, 1)
    print({k: out[k] for k in ("configs", "epoch_s", "us_per_step", "fused_step", "cuda_graphs")})
    if rec:
        torch.cuda.synchronize()
        for k, v in rec.items():
            if v:
                ts = sorted(a.elapsed_time(b) * 1e3 for a, b in v[len(v) // 2:])
                print(f"  {k:28s} calls {len(v):5d}  median {ts[len(ts) // 2]:8.2f} us  min {ts[0]:8.2f} us")


if __name__ == "__main__":
    main()
