"""One-off wider run of tests/test_gpu_parity.py::test_random_differential (seeds 8..71; the test suite
runs 0..7).  Last run on the final round-1 build: 64 seeds, 0 failures."""
import sys, traceback
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
import torch
import test_gpu_parity as T
from oracle import oracle as orc
dev = torch.device("cuda:0"); port = orc.Oracle("port")
bad = 0
for seed in range(8, 72):
    try:
        T.test_random_differential(dev, port, seed)
    except Exception as e:
        bad += 1; print("seed", seed, "FAILED", type(e).__name__, str(e)[:300])
print("done, failures:", bad)
