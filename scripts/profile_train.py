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


def main():
    ncfg = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
    os.environ.setdefault("KB_NO_GRAPHS", "1")
    dev = torch.device("cuda:0")
    torch.cuda.set_device(0)
    out = bench.train_epoch_probe(ncfg, dev, 1)
    print({k: out[k] for k in ("configs", "epoch_s", "us_per_step", "fused_step", "cuda_graphs")})


if __name__ == "__main__":
    main()
