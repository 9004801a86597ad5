#!/usr/bin/env python
"""bench.py — headline benchmark of the hot path (contract: see the task statement / DESIGN.md §5).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference's C++ on host cores

Workload (BASELINE.json configs[1]): synthetic jittered diamond-Si supercell, nx^3 conventional
cells (default 50 -> 1 000 000 atoms), a = 5.431 A, sigma = 0.05 A, set51, cos cutoff 5.0 A, fp64.
One step = one pass of the hot path over that supercell: padding atoms -> cell-list neighbor
list -> symmetry functions + dG/dr (packed, materialised in HBM).
metric = descriptor+dG/dr atoms/sec.  For N > 1 every rank runs the pass on its own supercell
(weak scaling, no data-path collective); timing is per-step CUDA events, max over ranks.

The same JSON line also carries the two multi-GPU paths of the metric's second half, timed in the
same run at every N:
  "slab"         ONE supercell of 48^3 cells (884 736 atoms, fits one GPU) split over the N ranks by
                 spatial domain with ghost halos: NCCL halo exchange + descriptors + forces + reverse
                 force exchange inside the step (strong scaling); at N = 8 also BASELINE.json's
                 configs[3]-sized cell (10 M atoms).
  "train_epoch"  synthetic 20 000-configuration Si set (configs[4] scaled to minutes): sharded create(),
                 then optimisation epochs with one gradient all-reduce per step; graph capture / warm-up
                 outside the timed region.

--impl reference runs the UNMODIFIED reference package (baseline/_ref, built by baseline/build_ref.sh)
through its own driver: kliff.parallel.parmap1(SymmetryFunction.transform, ...) for the descriptors
and CalculatorTorch(gpu=False) + Loss.minimize for the epoch (baseline/ref_runner.py).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

A_SI, SIGMA, RC = 5.431, 0.05, 5.0
FLOP_CANON = 380618.0   # per atom, set51 / n=28 / T=378 / T4=168
BYTES_CANON = 36724.0   # per atom, fp64 zeta + dG/dr written + inputs read once
BYTES_SCATTER = 36044.0  # per atom and direction of the force contraction: 8 D 3 (n+1) block + weights + indices
# dram__bytes_read.sum + dram__bytes_write.sum of the descriptor stage on this very workload come from
# an ncu --set full capture of the same command (a number taken under a profiler cannot be a bench value,
# and DRAM counters are not readable without one): scripts/ncu_traffic.py parses the capture and writes
# profiles/traffic_1M.json, which is read here.
TRAFFIC_FILE = os.path.join(ROOT, "profiles", "traffic_1M.json")


def ncu_traffic(nx):
    """(bytes per step, source) from the committed ncu capture of this workload, or (None, reason)."""
    try:
        d = json.load(open(TRAFFIC_FILE))
        if int(d.get("nx", -1)) != nx:
            return None, f"no capture for nx={nx}"
        return float(d["dram_bytes_per_step"]), d.get("source", TRAFFIC_FILE)
    except Exception as e:
        return None, f"no capture ({e})"


def diamond(nx, seed):
    basis = np.array([[0, 0, 0], [0, .5, .5], [.5, 0, .5], [.5, .5, 0],
                      [.25, .25, .25], [.25, .75, .75], [.75, .25, .75], [.75, .75, .25]])
    g = np.stack(np.meshgrid(np.arange(nx), np.arange(nx), np.arange(nx), indexing="ij"), -1).reshape(-1, 3)
    pos = (g[:, None, :] + basis[None]).reshape(-1, 3) * A_SI
    pos += np.random.default_rng(seed).normal(0.0, SIGMA, pos.shape)
    return np.eye(3) * A_SI * nx, np.ascontiguousarray(pos)


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return float(d.get("hbm_gbs", 6650.0)), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [x.strip() for x in line.split(",")]))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        rows = [r for t, r in self.rows if t0 <= t <= t1 + 0.2 and len(r) >= 7] or \
               [r for _, r in self.rows if len(r) >= 7]
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        sm = sorted(float(r[0]) for r in rows)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for k, n in enumerate(names) if any(r[3 + k].lower().startswith("active") for r in rows)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(rows[0][1]),
                "power_w_max": max(float(r[2]) for r in rows), "samples": len(rows), "reasons": reasons}


# ------------------------------------------------------------------------------------------------
# CPU baseline: the reference's own C++ (oracle/_ref) — or the C port when it is absent — on the
# host cores: serial paddings + neighbor list, then generate_one_atom(grad=True) over forked workers
# ------------------------------------------------------------------------------------------------
_W = {}


def _worker_init(kind, rc, kinds, rows, vals, coords, species, offsets, indices):
    from oracle import oracle as orc
    _W["o"] = orc.Oracle(kind)
    fams, pos = [], 0
    ncols = {1: 1, 2: 2, 3: 1, 4: 3, 5: 3}
    for k, r in zip(kinds, rows):
        c = ncols[int(k)]
        fams.append((int(k), np.ascontiguousarray(vals[pos:pos + r * c].reshape(r, c))))
        pos += r * c
    _W.update(rc=rc, fams=fams, coords=coords, species=species, offsets=offsets, indices=indices)


def _worker_run(span):
    lo, hi = span
    z, _ = _W["o"].symfun_range(_W["rc"], _W["fams"], lo, hi, _W["coords"], _W["species"],
                                _W["offsets"], _W["indices"], grad=True, keep_grad=False)
    return float(z.sum())


class CpuBaseline:
    def __init__(self, nx=12, cores=None):
        import multiprocessing as mp
        from kliff_b200.legacy.descriptors.hyperparams import get_set51
        from oracle import oracle as orc
        self.kind = "reference" if orc.have_reference() else "port"
        self.cores = cores or len(os.sched_getaffinity(0))
        o = orc.Oracle(self.kind)
        cell, pos = diamond(nx, 0)
        n = len(pos)
        z = np.full(n, 14, np.int32)
        t0 = time.time()
        pc, ps, pi = o.create_paddings(RC, cell, [1, 1, 1], pos, z)
        allc = np.concatenate([pos, pc])
        need = np.zeros(len(allc), np.int32)
        need[:n] = 1
        counts, offsets, indices = o.build_neighbors(allc, RC, need)
        self.t_neigh = time.time() - t0
        self.n = n
        fams = orc.fams_from_hyperparams(get_set51())
        kinds = np.array([k for k, _ in fams], np.int32)
        rows = np.array([v.shape[0] for _, v in fams], np.int32)
        vals = np.concatenate([v.reshape(-1) for _, v in fams])
        ctx = mp.get_context("fork")
        self.pool = ctx.Pool(self.cores, initializer=_worker_init,
                             initargs=(self.kind, np.array([[RC]]), kinds, rows, vals, allc,
                                       np.zeros(len(allc), np.int32), offsets, indices))

    def run(self, natoms):
        """Descriptor + dG/dr of `natoms` atoms (cyclic over the sample cell) on all cores.
        Returns atoms/s."""
        chunk = max(8, min(256, natoms // (self.cores * 4) or 8))
        spans, done = [], 0
        while done < natoms:
            lo = done % self.n
            hi = min(lo + chunk, self.n, lo + (natoms - done))
            spans.append((lo, hi))
            done += hi - lo
        t0 = time.time()
        self.pool.map(_worker_run, spans)
        return natoms / (time.time() - t0)

    def close(self):
        self.pool.close()
        self.pool.join()


def ref_run(what, timeout=1500, **kw):
    """Runs baseline/ref_runner.py (the unmodified reference, own process) and returns its JSON."""
    src = os.path.join(ROOT, "baseline", "_ref", "kliff_src")
    if not os.path.isdir(os.path.join(src, "kliff")):
        raise RuntimeError("baseline/_ref is not built (baseline/build_ref.sh needs /root/reference)")
    env = dict(os.environ)
    env["PYTHONPATH"] = src + os.pathsep + os.path.join(ROOT, "baseline", "stubs")
    env["CUDA_VISIBLE_DEVICES"] = ""
    cmd = [sys.executable, "-W", "ignore", os.path.join(ROOT, "baseline", "ref_runner.py"), what]
    for k, v in kw.items():
        cmd += [f"--{k}", str(v)]
    r = subprocess.run(cmd, env=env, cwd="/tmp", capture_output=True, text=True, timeout=timeout)
    for line in r.stdout.splitlines():
        if line.startswith("REF_JSON "):
            return json.loads(line[9:])
    raise RuntimeError(f"reference run failed (rc {r.returncode}): {r.stderr[-400:]}")


REF_NX = 3  # sample configurations of the reference arm: 3^3 cells = 216 atoms (its dense dG/dr is 57 MB each)


def reference_descriptor_sample(cores, repeat):
    """The reference's own driver (kliff/legacy/descriptors/descriptor.py:234-236) on `cores` processes
    over 2 x cores sample configurations; returns (atoms/s, description, raw)."""
    out = ref_run("descriptor", nx=REF_NX, configs=2 * cores, nprocs=cores, repeat=repeat)
    text = (f"{out['configs']} jittered diamond-Si cells of {out['atoms_per_config']} atoms (same lattice, cutoff, "
            f"set51 as the job; the reference's dense (N, D, 3N) output rules out larger cells) through "
            f"kliff.parallel.parmap1(SymmetryFunction.transform, fit_forces=True), nprocs={out['nprocs']}: "
            f"paddings, neighbor list, C++ generate_one_atom and the Python scatter into the dense arrays included")
    return out["atoms_per_s"], text, out


def run_reference_arm(args, rank, world):
    """--impl reference: the unmodified reference through its own driver on the host cores (rank 0 only)."""
    if rank != 0:
        return
    cores = len(os.sched_getaffinity(0))
    natoms_job = 8 * args.nx ** 3
    line = {"impl": "reference", "metric": "descriptor+dG/dr atoms/sec", "unit": "atoms/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "gpu_launches": 0}
    try:
        t0 = time.time()
        rate, text, out = reference_descriptor_sample(cores, repeat=args.steps + min(args.warmup, 1))
        secs = out["seconds"][min(args.warmup, 1):]
        value = out["atoms"] * len(secs) / sum(secs)
        line.update({
            "value": value, "ms_per_step": 1e3 * sum(secs) / len(secs),
            "config": {"workload": f"diamond-Si {natoms_job}-atom supercell, set51, cos 5.0 A, neighbor list + "
                                   f"descriptor + dG/dr (configs[1]) — SAMPLED: each step = {text}",
                       "nx": args.nx},
            "cpu_baseline": {"value": value, "unit": "atoms/s", "cores": cores, "kind": "reference", "sample": text},
            "e2e": {"value": value, "unit": "atoms/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}})
        try:  # second half of the metric: the reference's training epoch on CPU torch
            ep = ref_run("epoch", configs=200, nprocs=cores, epochs=1)
            line["train_epoch"] = {
                "configs": ep["configs"], "create_s": ep["create_s"], "epoch_s": ep["epoch_s"],
                "per_config_ms": {"create": 1e3 * ep["create_s"] / ep["configs"],
                                  "epoch": 1e3 * ep["epoch_s"] / ep["configs"]}, "what": ep["what"]}
        except Exception as e:
            line["train_epoch"] = {"error": str(e)}
    except Exception as e:
        line = {"impl": "reference", "unavailable": str(e).splitlines()[0][:200]}
    print(json.dumps(line))


def slab_block(nxs, ny, x0_cells, seed):
    """Jittered diamond block of nxs x ny x ny conventional cells starting x0_cells along a1."""
    basis = np.array([[0, 0, 0], [0, .5, .5], [.5, 0, .5], [.5, .5, 0],
                      [.25, .25, .25], [.25, .75, .75], [.75, .25, .75], [.75, .75, .25]])
    g = np.stack(np.meshgrid(np.arange(nxs), np.arange(ny), np.arange(ny), indexing="ij"), -1).reshape(-1, 3)
    pos = (g[:, None, :] + basis[None]).reshape(-1, 3) * A_SI
    pos += np.random.default_rng(seed).normal(0.0, SIGMA, pos.shape)
    pos[:, 0] += A_SI * x0_cells + A_SI / 8.0   # the 1/8-cell shift keeps jittered atoms off the slab faces
    return np.ascontiguousarray(pos)


def slab_probe(nx_total, ny, rank, world, dev, steps, warmup=3):
    """ONE periodic supercell of nx_total x ny x ny cells, slab r (nx_total / world cells thick) owned by
    rank r.  Step = halo exchange (NCCL p2p) -> paddings along a2/a3 -> neighbor list -> descriptor +
    dG/dr of the owned atoms -> force contraction with a fixed dE/dG -> reverse halo exchange of the
    ghost forces.  All ranks call this; times are CUDA events, max over ranks."""
    import torch
    import torch.distributed as dist

    from kliff_b200 import ops
    from kliff_b200.legacy.descriptors import SymmetryFunction
    from kliff_b200.parallel import SlabDecomposition

    assert nx_total % world == 0
    nxs = nx_total // world
    cell = np.diag([A_SI * nx_total, A_SI * ny, A_SI * ny])
    pos = slab_block(nxs, ny, rank * nxs, seed=rank)
    desc = SymmetryFunction(cut_name="cos", cut_dists={"Si-Si": RC}, hyperparams="set51", normalize=False,
                            dtype=np.float64)
    dd = SlabDecomposition(cell, [1, 1, 1], RC, rank=rank, world=world)
    x_own = dd.wrap(torch.tensor(pos, device=dev))
    keep = dd.owner(x_own) == rank             # jitter can push a face atom across the slab boundary:
    moved = int((~keep).sum())                 # such atoms are simply not simulated (counted below)
    x_own = x_own[keep].contiguous()
    n_own = x_own.shape[0]
    code = torch.zeros(n_own, dtype=torch.int32, device=dev)
    w = torch.randn(n_own, 51, dtype=torch.float64, device=dev, generator=torch.Generator(dev).manual_seed(rank))
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def step():
        gx, gc = dd.exchange_halo(x_own, code)
        fp = dd.local_fingerprints(desc, x_own, code, gx, gc)
        F = dd.forces(w, fp)
        return fp, F, gx.shape[0]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(warmup):
        fp, F, ng = step()
        del fp
        flush.fill_(1)
    barrier()
    launches0 = ops.launch_count()
    ms = []
    for _ in range(steps):
        flush.fill_(0)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fp, F, ng = step()
        e1.record()
        torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
        npad, nloc = fp["n_pad"], fp["n_local"]
        del fp
    barrier()
    launches = ops.launch_count() - launches0
    net = F.sum(dim=0)                          # total force on the cell: zero by translation invariance
    tot = torch.tensor([sum(ms), float(n_own), 0.0, 0.0, 0.0], dtype=torch.float64, device=dev)
    tot[2:5] = net
    tmax = tot.clone()
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    total_ms, atoms = float(tmax[0]), float(tot[1])
    tot[2] = tot[2:5].abs().max()
    del flush, w, x_own
    torch.cuda.empty_cache()
    return {"workload": f"ONE diamond-Si supercell of {nx_total}x{ny}x{ny} cells, slab of {nxs} cells per GPU, "
                        f"halo {RC} A, NCCL halo + reverse force exchange inside the step",
            "atoms": int(atoms), "n_gpus": world, "ms_per_step": total_ms / steps,
            "atoms_per_s": atoms * steps / (total_ms * 1e-3), "steps": steps,
            "atoms_rank0": n_own, "ghosts_rank0": int(ng), "paddings_rank0": int(npad),
            "local_atoms_rank0": int(nloc), "atoms_dropped_at_faces_rank0": moved,
            "net_force_max_abs": float(tot[2]), "gpu_launches": int(launches),
            "l2": "256 MiB write between timed steps"}


def run_slab_workload(args, rank, world, dev):
    """--workload slab: weak-scaling form (nx cells per GPU along a1), one JSON line of its own."""
    r = slab_probe(args.nx * world, args.nx, rank, world, dev, args.steps, max(args.warmup, 3))
    if rank == 0:
        print(json.dumps({
            "metric": "descriptor+dG/dr+forces atoms/sec (one supercell, spatial domains + ghost halos)",
            "value": r["atoms_per_s"], "unit": "atoms/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": r["ms_per_step"], "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {k: r[k] for k in r if k not in ("atoms_per_s", "ms_per_step", "gpu_launches")},
            "gpu_launches": r["gpu_launches"]}))


def synthetic_configs(ncfg):
    """The synthetic training set of the train_epoch probe: `ncfg` 64-atom diamond-Si configurations
    (2x2x2 cells, a ~ U(5.2, 5.7), jitter 0.1 A, seed = index; reference energies/forces = N(0,1) noise)."""
    from kliff_b200.dataset import Configuration
    from kliff_b200.dataset.weight import Weight

    basis = np.array([[0, 0, 0], [0, .5, .5], [.5, 0, .5], [.5, .5, 0],
                      [.25, .25, .25], [.25, .75, .75], [.75, .25, .75], [.75, .75, .25]])
    g = np.stack(np.meshgrid(np.arange(2), np.arange(2), np.arange(2), indexing="ij"), -1).reshape(-1, 3)
    frac = (g[:, None, :] + basis[None]).reshape(-1, 3)
    weight = Weight(forces_weight=0.3)
    confs = []
    for c in range(ncfg):
        rng = np.random.default_rng(c)
        a = rng.uniform(5.2, 5.7)
        pos = frac * a + rng.normal(0, 0.1, frac.shape)
        confs.append(Configuration(cell=np.eye(3) * 2 * a, species=["Si"] * 64, coords=pos, PBC=[True] * 3,
                                   energy=float(rng.normal()), forces=rng.normal(size=(64, 3)),
                                   weight=weight, identifier=f"syn{c}"))
    return confs


def train_epoch_probe(ncfg, dev, world=1):
    """Second half of BASELINE.json's metric ("train epoch time"): synthetic set of `ncfg` 64-atom
    diamond-Si configurations (2x2x2 cells, a ~ U(5.2, 5.7), jitter 0.1 A, seed = index; reference
    energies/forces = N(0,1) noise), set51, MLP 51-10-10-1, Adam lr 1e-3, batch 100, fp32 model as in
    the reference's tutorials.  Reports create() (descriptors + dG/dr + normalisation) and the time
    of one optimisation pass over the set through Loss.minimize.  With world > 1 every rank calls
    this: configurations are sharded round-robin over the ranks (create) and every batch is
    evaluated data-parallel with a gradient all-reduce per step; times are the max over ranks."""
    import contextlib
    import io
    import tempfile

    import torch

    from kliff_b200.dataset import Configuration
    from kliff_b200.dataset.weight import Weight
    from kliff_b200.legacy import nn
    from kliff_b200.legacy.calculators import CalculatorTorch
    from kliff_b200.legacy.descriptors import SymmetryFunction
    from kliff_b200.legacy.loss import Loss
    from kliff_b200.models import NeuralNetwork

    confs = synthetic_configs(ncfg)
    desc = SymmetryFunction(cut_name="cos", cut_dists={"Si-Si": RC}, hyperparams="set51", normalize=True)
    desc.store_gradients_on_disk = False  # GPU-resident gradients only; the pickle keeps zeta/energy/forces
    model = NeuralNetwork(desc)
    model.add_layers(nn.Linear(51, 10), nn.Tanh(), nn.Linear(10, 10), nn.Tanh(), nn.Linear(10, 1))
    tmp = tempfile.mkdtemp()
    model.set_save_metadata(prefix=os.path.join(tmp, "m"), start=10 ** 9, frequency=10 ** 9)
    calc = CalculatorTorch(model)

    def fence():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    def over_ranks(seconds):
        t = torch.tensor([seconds], dtype=torch.float64, device=dev)
        if world > 1:
            torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        return float(t)

    # warm-up create() on a small set, outside the timed region: kernel modules, scratch pools, the NCCL
    # communicator of the mean/stdev all-reduce (0.25 s on the first call of a process)
    wdesc = SymmetryFunction(cut_name="cos", cut_dists={"Si-Si": RC}, hyperparams="set51", normalize=True)
    wdesc.store_gradients_on_disk = False
    wmodel = NeuralNetwork(wdesc)
    wmodel.add_layers(nn.Linear(51, 10), nn.Tanh(), nn.Linear(10, 10), nn.Tanh(), nn.Linear(10, 1))
    CalculatorTorch(wmodel).create(confs[:min(ncfg, 64 * world)], fingerprints_filename=os.path.join(tmp, "wfp.pkl"),
                                   fingerprints_mean_stdev_filename=os.path.join(tmp, "wms.pkl"))
    del wmodel, wdesc
    fence()
    t0 = time.time()
    calc.create(confs, fingerprints_filename=os.path.join(tmp, "fp.pkl"),
                fingerprints_mean_stdev_filename=os.path.join(tmp, "ms.pkl"))
    torch.cuda.synchronize()
    t_create = over_ranks(time.time() - t0)
    loss = Loss(calc)
    epochs = 5
    with contextlib.redirect_stdout(io.StringIO()):
        # warm-up OUTSIDE the timed region: library handles, NCCL communicator, CUDA-graph capture of the
        # training epoch and of the loss pass
        loss.minimize(method="Adam", num_epochs=1, batch_size=100, lr=1e-3)
        fused = loss._fused is not None
        fence()
        if fused:
            batches = loss._fused_batches()
            loss._fused.run_pass(batches, train=True)
            fence()
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
            t0 = time.time()
            ev[0].record()
            for _ in range(epochs):
                loss._fused.run_pass(batches, train=True)   # one graph replay + one host read of the epoch loss
            ev[1].record()
            torch.cuda.synchronize()
            t_epoch = over_ranks(ev[0].elapsed_time(ev[1]) * 1e-3 / epochs)
            t_epoch_wall = over_ranks((time.time() - t0) / epochs)
            t0 = time.time()
            for _ in range(3):
                loss._fused.run_pass(batches, train=False)
            torch.cuda.synchronize()
            t_eval = over_ranks((time.time() - t0) / 3.0)
        else:
            t0 = time.time()
            loss.minimize(method="Adam", num_epochs=epochs, batch_size=100, lr=1e-3)
            torch.cuda.synchronize()
            t_epoch = t_epoch_wall = over_ranks((time.time() - t0) / (epochs + 1))
            t_eval = None
        # the reference tutorial's call, as a user makes it (optimizer + graphs rebuilt inside): wall clock
        fence()
        t0 = time.time()
        loss.minimize(method="Adam", num_epochs=10, batch_size=100, lr=1e-3)
        torch.cuda.synchronize()
        t_run = over_ranks(time.time() - t0)
        # the usual data-parallel alternative: batch 100 PER GPU (global batch 100 N, N times fewer steps) — the
        # per-GPU work of a step is then what one GPU does at batch 100 instead of 1/N of it
        per_gpu = None
        if world > 1 and fused:
            try:
                loss.minimize(method="Adam", num_epochs=1, batch_size=100 * world, lr=1e-3)
                batches = loss._fused_batches()
                loss._fused.run_pass(batches, train=True)
                fence()
                ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
                ev[0].record()
                for _ in range(epochs):
                    loss._fused.run_pass(batches, train=True)
                ev[1].record()
                torch.cuda.synchronize()
                tw = over_ranks(ev[0].elapsed_time(ev[1]) * 1e-3 / epochs)
                nst = (ncfg + 100 * world - 1) // (100 * world)
                per_gpu = {"batch_size": 100 * world, "epoch_s": tw, "steps_per_epoch": nst,
                           "us_per_step": 1e6 * tw / nst,
                           "note": "same set, batch 100 per GPU: a different optimisation trajectory than the "
                                   "reference's batch_size=100, reported beside it, not instead of it"}
            except Exception as e:  # the extra line must never cost the main one
                per_gpu = {"error": f"{type(e).__name__}: {e}"[:200]}
    steps = (ncfg + 99) // 100
    return {"configs": ncfg, "atoms": 64 * ncfg, "n_gpus": world,
            "configs_on_rank0": len(calc.fingerprints_dataset), "create_s": t_create,
            "epoch_s": t_epoch, "epoch_wall_s": t_epoch_wall, "full_epoch_s": t_create + t_epoch,
            "steps_per_epoch": steps, "batch100_per_gpu": per_gpu,
            "us_per_step": 1e6 * t_epoch / steps, "loss_pass_s": t_eval,
            "minimize_10_epochs_s": t_run, "fused_step": fused, "cuda_graphs": bool(loss.use_cuda_graphs),
            "six_launch_step": bool(fused and loss._fused.fast),
            "grad_allreduce": (None if world == 1 else
                               "nvlink-peer (one launch with the partial reduction and Adam)"
                               if fused and loss._fused.peer is not None else "nccl"),
            "note": "epoch_s = device time (CUDA events, max over ranks) of one optimisation epoch = one replay of "
                    "the epoch graph (batch 100: forward, forces, loss, second-order backward, gradient all-reduce, "
                    "Adam per step), graph capture and warm-up outside the timed region; minimize_10_epochs_s = "
                    "wall clock of Loss.minimize(num_epochs=10) as a user calls it (capture + 10 epochs + final "
                    "loss pass); fp32 MLP 51-10-10-1, dG/dr computed in fp64 and stored in fp32 (descriptor dtype, "
                    "as the reference); create_s = sharded descriptors + dG/dr + normalisation (wall clock, max over ranks, a "
                    "64-configuration-per-rank warm-up create() before it); full_epoch_s = create_s + epoch_s = config 5's "
                    "'full training epoch (descriptors + forces + gradient all-reduce)'"}


# ------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--nx", type=int, default=50, help="conventional cells per edge (50 -> 1M atoms)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--workload", default="supercell", choices=["supercell", "slab"],
                    help="supercell: configs[1], one independent 1M-atom cell per GPU (default, the contract "
                         "line); slab: configs[3]-style, ONE supercell of gpus x nx^3 cells split by spatial "
                         "domain with ghost halos (NCCL halo + reverse force exchange inside the step)")
    ap.add_argument("--epoch-configs", type=int, default=20000,
                    help="size of the synthetic 64-atom-configuration set for the train_epoch probe (0 = off)")
    ap.add_argument("--slab-cells", type=int, default=48,
                    help="edge (conventional cells) of the ONE supercell of the slab probe, split over the ranks "
                         "along a1 (0 = off); must be divisible by the number of GPUs")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference_arm(args, rank, world)
        return

    # cpu_baseline and the train_epoch probe are N=1 extras (rank 0 of a single-process run).  The
    # baseline's worker pool is forked HERE, before CUDA / NCCL exist in this process: forking a process
    # that already holds a CUDA context and NCCL watchdog threads can deadlock the children.
    cpu_base = None
    if world == 1 and not args.no_cpu_baseline and args.workload == "supercell":
        try:
            cpu_base = CpuBaseline(nx=12)
        except Exception as e:
            cpu_base = e

    import torch
    import torch.distributed as dist

    from kliff_b200 import ops
    from kliff_b200.dataset import Configuration
    from kliff_b200.legacy.descriptors import SymmetryFunction

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: kliff_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        import datetime
        # a rank that fails leaves the others in a collective: bound that wait to minutes, not the default
        dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=240))
    ops.device_info()  # raises unless sm_100
    if args.workload == "slab":
        run_slab_workload(args, rank, world, dev)
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()
        leave(world)
        return

    # ---- synthetic input, resident in HBM before the timed region
    cell, pos = diamond(args.nx, seed=rank)
    n = len(pos)
    coords = torch.tensor(pos, device=dev)
    z = torch.full((n,), 14, dtype=torch.int32, device=dev)
    desc = SymmetryFunction(cut_name="cos", cut_dists={"Si-Si": RC}, hyperparams="set51", normalize=True,
                            dtype=np.float64)
    P = desc._params
    D = P.n_desc
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    state = {}

    def step(timers=None):
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)] if timers is not None else None
        if ev:
            ev[0].record()
        pc, ps, pi = ops.create_paddings(RC, cell, [1, 1, 1], coords, z)
        allc = torch.cat([coords, pc])
        counts, offsets, indices, maxn = ops.build_neighbors(allc, n, RC)
        if ev:
            ev[1].record()
        code = torch.zeros(allc.shape[0], dtype=torch.int32, device=dev)
        if "dz" not in state:  # 36 GB output buffer allocated once and reused
            dz_ptr, total = ops.block_offsets(n, None, offsets, D)
            state.update(dz=torch.empty(total, dtype=torch.float64, device=dev), total=total)
        dz_ptr, total = ops.block_offsets(n, None, offsets, D)
        assert total == state["total"]
        if ev:
            ev[2].record()
        zeta, dz, _ = ops.symfun_forward(P, allc, code, offsets, indices, maxn, n_centers=n,
                                         dz_ptr=dz_ptr, dz=state["dz"])
        if ev:
            ev[3].record()
            timers.append(ev)
        state.update(zeta=zeta, offsets=offsets, indices=indices, pi=pi, dz_ptr=dz_ptr, npad=pc.shape[0],
                     pairs=indices.numel())
        return zeta

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step()
        flush.fill_(1)
    barrier()

    sampler = ClockSampler(local_rank) if rank == 0 else None
    ops.kernel_timing(True)   # CUDA events around the descriptor kernels, on their stream
    launches0 = ops.launch_count()
    timers, t_wall0 = [], time.time()
    for _ in range(args.steps):
        flush.fill_(0)            # L2 flush between timed steps (outside the step's events)
        torch.cuda.synchronize()
        step(timers)
    barrier()
    t_wall1 = time.time()
    launches = ops.launch_count() - launches0
    ktimes = ops.kernel_timing(False)
    clocks = sampler.stop(t_wall0, t_wall1) if sampler else None

    step_ms = [ev[0].elapsed_time(ev[3]) for ev in timers]
    sym_ms = [ev[2].elapsed_time(ev[3]) for ev in timers]
    nl_ms = [ev[0].elapsed_time(ev[1]) for ev in timers]
    total_ms = torch.tensor([sum(step_ms)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(total_ms, op=dist.ReduceOp.MAX)
    total_ms = float(total_ms)
    value = world * n * args.steps / (total_ms * 1e-3)

    # ---- end to end through the plugin API: host buffers in, zeta + forces out (rank-local)
    host_pos = torch.from_numpy(pos).pin_memory()
    conf = Configuration(cell=cell, species=np.full(n, "Si"), coords=host_pos.numpy(), PBC=[True] * 3)
    w = torch.randn(n, D, dtype=torch.float64, device=dev)
    zeta_host = torch.empty((n, D), dtype=torch.float64).pin_memory()
    f_host = torch.empty(3 * n, dtype=torch.float64).pin_memory()

    def e2e_step():
        # descriptors stream to the pinned host array piece by piece while the kernels run
        packed = desc.transform_packed([conf], grad=True, device=dev, zeta_host=zeta_host)
        F = packed.forces(w, 0, n)
        f_host.copy_(F, non_blocking=True)
        packed.host_ready.synchronize()
        torch.cuda.synchronize()
        return packed

    # ---- force contraction on the blocks of the last step (both directions), library events
    scat = None
    try:
        image = torch.cat([torch.arange(n, dtype=torch.int32, device=dev), state["pi"]])
        sc_pack = dict(offsets=state["offsets"], indices=state["indices"], image=image, dz=state["dz"],
                    dz_ptr=state["dz_ptr"], D=D, n_out=n)
        sc_w = torch.randn(n, D, dtype=torch.float64, device=dev)
        sc_F = ops._scatter_fwd(sc_w, sc_pack)
        ops._scatter_bwd(sc_F, sc_pack)
        torch.cuda.synchronize()
        reps = 3
        ops.kernel_timing(True)
        for _ in range(reps):
            flush.fill_(0)
            sc_F = ops._scatter_fwd(sc_w, sc_pack)
        t_fwd = ops.kernel_timing(True)["scatter"][0] / reps
        for _ in range(reps):
            flush.fill_(0)
            ops._scatter_bwd(sc_F, sc_pack)
        t_bwd = ops.kernel_timing(False)["scatter"][0] / reps
        scat = (t_fwd, t_bwd)
        del sc_w, sc_F, sc_pack, image
    except Exception as e:
        scat = e
    npad, pairs = int(state["npad"]), int(state["pairs"])
    state_total = int(state["total"])
    state.clear()
    del timers
    torch.cuda.empty_cache()
    e2e_step()
    barrier()
    t0 = time.time()
    e2e_steps = max(1, min(args.steps, 3))
    for _ in range(e2e_steps):
        e2e_step()
    barrier()
    e2e_s = torch.tensor([time.time() - t0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    e2e_value = world * n * e2e_steps / float(e2e_s)

    # second half of the metric: the training-epoch probe (all ranks: sharded create + data-parallel steps)
    train_epoch = None
    if args.epoch_configs > 0:
        try:
            train_epoch = train_epoch_probe(args.epoch_configs, dev, world)
        except Exception as e:  # reported, never required (single-process case only: see NCCL timeout)
            train_epoch = {"error": str(e)}

    # one supercell split by spatial domain (strong scaling at a size that fits one GPU), all ranks
    slab = None
    if args.slab_cells > 0:
        try:
            del flush
            torch.cuda.empty_cache()
            slab = slab_probe(args.slab_cells, args.slab_cells, rank, world, dev, steps=max(3, min(args.steps, 5)))
            if world == 8:  # BASELINE.json configs[3]: 10 M atoms over 8 GPUs
                slab["config4_10M"] = slab_probe(8 * 54, 54, rank, world, dev, steps=3)
        except Exception as e:
            slab = {"error": f"{type(e).__name__}: {e}"}

    if world > 1:
        dist.barrier()
        torch.cuda.synchronize()
    if rank != 0:
        leave(world)
        return

    hbm_peak, peak_src = measured_peaks()
    fp64_peak = ops.probe_fp64_peak()
    # the descriptor stage = symfun_prep_kernel + symfun_sweep_kernel launches (one pair per chunk of
    # 262 144 atoms); their device time per step, from the library's own events
    kern_ms = {k: v[0] / args.steps for k, v in ktimes.items()}
    kern_n = {k: v[1] // args.steps for k, v in ktimes.items()}
    sym_s = sum(kern_ms[k] for k in ("prep", "sweep", "fused")) * 1e-3
    if sym_s <= 0.0:
        sym_s = float(np.mean(sym_ms)) * 1e-3
    ach_gbs = BYTES_CANON * n / sym_s / 1e9
    ach_tf = FLOP_CANON * n / sym_s / 1e12
    kname = "symfun_prep_kernel<true, double, true> + symfun_sweep3_kernel<double>"
    kinfo = {"launches_per_step": kern_n, "ms_per_step": kern_ms,
             "dominant": "symfun_sweep3_kernel<double>",
             "note": "achieved = canonical work of one step / summed device time of these launches"}
    traffic, traffic_src = ncu_traffic(args.nx)
    fp64 = {"kernel": kname, "kernels": kinfo, "bound": "fp64", "achieved": ach_tf, "peak": fp64_peak,
            "unit": "TFLOP/s", "frac": ach_tf / fp64_peak, "traffic": traffic, "traffic_source": traffic_src,
            "peak_source": "DFMA-chain probe in this run (kb_probe_fp64_peak); B200 has no other FP64 unit: "
                           "the FP64 mma.sync path shares this pipe (profiles/r01_probe_dmma_vs_dfma.txt)",
            "note": "the BINDING roofline of the stage (SURVEY.md §8d: arithmetic intensity 10 flop/B > machine "
                    "balance 5.6): canonical 380 618 flop/atom / summed device time of the stage's launches; "
                    "traffic = DRAM bytes per step of those launches from the ncu capture named in traffic_source"}
    line = {
        "metric": "descriptor+dG/dr atoms/sec", "value": value, "unit": "atoms/s", "n_gpus": world,
        "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": total_ms / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": f"diamond-Si {n}-atom supercell per GPU (jitter 0.05 A, seed=rank), set51, cos 5.0 A: "
                               "paddings + cell-list neighbor list + descriptor + dG/dr (configs[1])",
                   "nx": args.nx, "atoms_per_gpu": n, "paddings": npad, "neighbor_pairs": pairs,
                   "l2": "256 MiB write between timed steps (outside the step's CUDA events)",
                   "step_breakdown_ms": {"paddings+neighbors": float(np.mean(nl_ms)),
                                         "symfun+dG/dr": float(np.mean(sym_ms))}},
        "roofline": fp64,
        "roofline_fp64": {k: fp64[k] for k in ("kernel", "bound", "achieved", "peak", "unit", "frac", "peak_source")},
        "roofline_hbm": {"kernel": kname, "bound": "hbm", "achieved": ach_gbs, "peak": hbm_peak, "unit": "GB/s",
                         "frac": ach_gbs / hbm_peak, "peak_source": peak_src,
                         "note": "algorithmic bytes 36 724 B/atom (fp64 zeta + dG/dr written, inputs read once); "
                                 "the non-binding term"},
        "e2e": {"value": e2e_value, "unit": "atoms/s", "h2d_bytes_per_step": int(24 * n),
                "d2h_bytes_per_step": int(8 * n * D + 24 * n),
                "what": "SymmetryFunction.transform_packed(Configuration with pinned host coords, "
                        "zeta_host=pinned array) + force contraction; zeta (streamed per 262 144-atom piece "
                        "on a side stream) and forces copied back to pinned host memory; the dG/dr blocks "
                        f"({8 * state_total / 1e9:.1f} GB) stay on the device: they are consumed there by the force "
                        "contraction, the reference's dense equivalent would not fit any host"},
        "gpu_launches": int(launches),
        "clocks": clocks,
    }
    if isinstance(scat, tuple):
        line["roofline_scatter"] = {
            "kernel": "force_scatter_kernel<false,double> (F = -sum dE/dG dG/dr) / force_gather_kernel<double> (adjoint)",
            "bound": "hbm", "unit": "GB/s", "peak": hbm_peak, "peak_source": peak_src,
            "bytes_per_atom": BYTES_SCATTER,
            "fwd": {"ms": scat[0], "achieved": BYTES_SCATTER * n / scat[0] / 1e6,
                    "frac": BYTES_SCATTER * n / scat[0] / 1e6 / hbm_peak},
            "bwd": {"ms": scat[1], "achieved": BYTES_SCATTER * n / scat[1] / 1e6,
                    "frac": BYTES_SCATTER * n / scat[1] / 1e6 / hbm_peak},
            "note": "library CUDA events around each launch, 256 MiB L2 flush between launches, on this step's blocks"}
    elif scat is not None:
        line["roofline_scatter"] = {"error": str(scat)}
    if slab is not None:
        line["slab"] = slab
    if train_epoch is not None:
        line["train_epoch"] = train_epoch
    if world == 1 and not args.no_cpu_baseline:
        # (1) the reference package through its own driver (same call as --impl reference)
        cores = len(os.sched_getaffinity(0))
        try:
            rate, text, _ = reference_descriptor_sample(cores, repeat=2)
            line["cpu_baseline"] = {"value": rate, "unit": "atoms/s", "cores": cores, "kind": "reference",
                                    "sample": text}
        except Exception as e:  # the baseline is reported, never required
            line["cpu_baseline"] = {"value": None, "unit": "atoms/s", "cores": 0, "kind": "reference",
                                    "sample": f"failed: {str(e)[:200]}"}
        # (2) its C++ alone (oracle/_ref shim, no Python scatter, no dense arrays): the faster CPU figure
        try:
            if isinstance(cpu_base, Exception):
                raise cpu_base
            base = cpu_base
            sample = int(max(2000, min(200000, 470.0 * base.cores * 8.0)))
            base.run(base.cores * 8)
            rate = base.run(sample)
            base.close()
            line["cpu_baseline_native"] = {
                "value": rate, "unit": "atoms/s", "cores": base.cores, "kind": base.kind,
                "sample": f"{sample} atoms of a {base.n}-atom jittered diamond cell (same lattice, cutoff, set51), "
                          f"the reference's C++ generate_one_atom(grad=True) alone over {base.cores} forked workers "
                          f"(serial paddings + neighbor list of the sample cell: {base.t_neigh:.2f} s, not included)"}
        except Exception as e:
            line["cpu_baseline_native"] = {"value": None, "sample": f"failed: {str(e)[:200]}"}
        # (3) the reference's training epoch on CPU torch, beside train_epoch
        if isinstance(train_epoch, dict) and "error" not in train_epoch:
            try:
                ep = ref_run("epoch", configs=200, nprocs=cores, epochs=1)
                train_epoch["cpu_baseline"] = {
                    "kind": "reference", "cores": cores, "configs": ep["configs"], "create_s": ep["create_s"],
                    "epoch_s": ep["epoch_s"],
                    "epoch_s_at_this_size": ep["epoch_s"] * train_epoch["configs"] / ep["configs"],
                    "create_s_at_this_size": ep["create_s"] * train_epoch["configs"] / ep["configs"],
                    "sample": f"{ep['configs']} of the same synthetic configurations: {ep['what']}; scaled linearly "
                              f"in the number of configurations (both costs are per configuration)"}
            except Exception as e:
                train_epoch["cpu_baseline"] = {"error": str(e)[:200]}
    print(json.dumps(line))
    leave(world)


def leave(world):
    """End of a rank.  Multi-rank runs leave without dist.destroy_process_group(): CUDA graphs that
    captured NCCL collectives are still alive at this point and tearing the communicator down under them
    was seen to block; every collective has completed (barrier + synchronize above), so the process
    simply exits."""
    sys.stdout.flush()
    sys.stderr.flush()
    if world > 1:
        from kliff_b200.legacy.descriptors.descriptor import wait_for_fingerprints
        wait_for_fingerprints()  # background writers of fingerprint files (create()) finish first
        os._exit(0)


if __name__ == "__main__":
    try:
        main()
    except Exception:
        if int(os.environ.get("WORLD_SIZE", "1")) > 1:
            # a failing rank leaves at once: interpreter shutdown under live graphs / a pending collective can
            # block until the job's time limit, and torchrun only ends the other ranks once this one is gone
            import traceback
            traceback.print_exc()
            sys.stderr.flush()
            os._exit(1)
        raise
