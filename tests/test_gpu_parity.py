"""Parity tests proper: the CUDA path (through the C-ABI) against the CPU oracle, the recorded
outputs of the unmodified reference C++ (tests/golden/ref_run_*.npz) and the reference's own
known-answer vectors.  Bars: integer / coordinate outputs bit-exact; zeta and
dG/dr <= 1e-10 max-norm-relative per (atom, descriptor) row; forces <= 1e-10 of max|F|."""
import glob
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN, fams_from_npz, load_golden, rowwise_rel_err
from kliff_b200 import _lib, ops
from kliff_b200.legacy.descriptors.hyperparams import get_set30, get_set51
from oracle import oracle as orc

pytestmark = pytest.mark.gpu

TOL = 1e-10
RUNS = sorted(os.path.basename(p) for p in glob.glob(os.path.join(GOLDEN, "ref_run_*.npz"))
              if "mos2" not in p)
KNAME = {1: "g1", 2: "g2", 3: "g3", 4: "g4", 5: "g5"}


def hp_from_fams(fams):
    from collections import OrderedDict
    hp = OrderedDict()
    for kind, v in fams:
        if kind == 1:
            hp["g1"] = None
        elif kind == 2:
            hp["g2"] = [{"eta": r[0], "Rs": r[1]} for r in v]
        elif kind == 3:
            hp["g3"] = [{"kappa": r[0]} for r in v]
        else:
            hp[KNAME[kind]] = [{"zeta": r[0], "lambda": r[1], "eta": r[2]} for r in v]
    return hp


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available()
    sm, major, minor = ops.device_info()
    assert major == 10, "these kernels are built for sm_100a only"
    return torch.device("cuda:0")


@pytest.fixture(scope="module")
def port():
    return orc.Oracle("port")


def gpu_neighbors(dev, cell, pbc, coords, z, infl, padding_need_neigh=False, half=False):
    c = torch.tensor(np.ascontiguousarray(coords), dtype=torch.float64, device=dev)
    zt = torch.tensor(np.ascontiguousarray(z), dtype=torch.int32, device=dev)
    pc, ps, pi = ops.create_paddings(infl, cell, pbc, c, zt)
    allc = torch.cat([c, pc]) if pc.shape[0] else c
    n = c.shape[0]
    need = None
    if padding_need_neigh:
        need = torch.ones(allc.shape[0], dtype=torch.int32, device=dev)
    counts, offsets, indices, maxn = ops.build_neighbors(allc, n, infl, need=need, half=half)
    image = torch.cat([torch.arange(n, dtype=torch.int32, device=dev), pi])
    return dict(pc=pc, ps=ps, pi=pi, allc=allc, counts=counts, offsets=offsets, indices=indices,
                maxn=maxn, image=image, n=n)


def blocks_from_packed(dz, dz_ptr, counts, D):
    dz = dz.cpu().numpy()
    ptr = dz_ptr.cpu().numpy()
    out = []
    for t in range(len(ptr) - 1):
        n = int(counts[t])
        out.append(dz[ptr[t]:ptr[t] + 3 * D * (n + 1)].reshape(D, n + 1, 3))
    return out


# --------------------------------------------------------------------------------------------
def test_kat_neighbor(dev):
    """Reference tests/test_neighbor.py:37-64 through the GPU path."""
    g = load_golden("ref_kat_neighbor.npz")
    r = gpu_neighbors(dev, g["cell"], g["pbc"], g["coords"], g["z"], float(g["infl_dist"]))
    assert np.allclose(r["allc"].cpu().numpy(), g["target_coords"])
    assert np.array_equal(np.concatenate([g["z"], r["ps"].cpu().numpy()]), g["target_z"])
    off = r["offsets"].cpu().numpy()
    idx = r["indices"].cpu().numpy()
    for i in range(4):
        assert np.array_equal(idx[off[i]:off[i + 1]], np.sort(g["all_indices"][i]))
    assert np.all(r["counts"].cpu().numpy()[4:] == 0)


def test_kat_dimer(dev):
    cell = np.eye(3) * 200.0
    coords = np.array([[0.0, 0, 0], [2.0, 0, 0]])
    for P in (1, 0):
        r = gpu_neighbors(dev, cell, [P] * 3, coords, np.array([6, 8], np.int32), 5.0)
        off = r["offsets"].cpu().numpy()
        idx = r["indices"].cpu().numpy()
        assert list(idx[off[0]:off[1]]) == [1] and list(idx[off[1]:off[2]]) == [0]
        assert r["pc"].shape[0] == 0


@pytest.mark.parametrize("name", RUNS + ["ref_run_mos2_neigh.npz"])
def test_neighbors_bit_exact_vs_reference(dev, name):
    g = load_golden(name)
    infl = float(np.max(g["rcut"])) if "rcut" in g.files else float(g["infl_dist"])
    r = gpu_neighbors(dev, g["cell"], g["pbc"], g["coords"], g["z"], infl)
    assert np.array_equal(r["pc"].cpu().numpy(), g["pad_coords"])  # every coordinate bit
    assert np.array_equal(r["ps"].cpu().numpy(), g["pad_z"])
    assert np.array_equal(r["pi"].cpu().numpy(), g["pad_image"])
    assert np.array_equal(r["counts"].cpu().numpy(), g["nl_counts"])
    assert np.array_equal(r["offsets"].cpu().numpy(), g["nl_offsets"])
    assert np.array_equal(r["indices"].cpu().numpy(), orc.canonical(g["nl_offsets"], g["nl_indices"]))
    assert r["maxn"] == int(g["nl_counts"].max())


def test_neighbors_padding_need_neigh_and_half(dev, port):
    g = load_golden("ref_run_tri30_allfam.npz")
    infl = float(np.max(g["rcut"]))
    full = gpu_neighbors(dev, g["cell"], g["pbc"], g["coords"], g["z"], infl, padding_need_neigh=True)
    allc = full["allc"].cpu().numpy()
    oc, oo, oi = port.build_neighbors(allc, infl, np.ones(len(allc), np.int32))
    assert np.array_equal(full["counts"].cpu().numpy(), oc)
    assert np.array_equal(full["indices"].cpu().numpy(), orc.canonical(oo, oi))
    half = gpu_neighbors(dev, g["cell"], g["pbc"], g["coords"], g["z"], infl, half=True)
    hoff, hidx = half["offsets"].cpu().numpy(), half["indices"].cpu().numpy()
    can = orc.canonical(g["nl_offsets"], g["nl_indices"])
    for a in range(30):
        ref = can[g["nl_offsets"][a]:g["nl_offsets"][a + 1]]
        assert np.array_equal(hidx[hoff[a]:hoff[a + 1]], ref[ref > a])


def test_neighbors_several_cutoffs_per_build(dev, port):
    """nbl_build with k cutoffs (neighbor_list.cpp:148-160, 210-217): one binning by the influence
    distance, list k keeps rsq < cutoff_k^2.  Checked against the oracle's build for every cutoff."""
    g = load_golden("ref_run_tri30_allfam.npz")
    infl = float(np.max(g["rcut"]))
    full = gpu_neighbors(dev, g["cell"], g["pbc"], g["coords"], g["z"], infl)
    allc = full["allc"]
    need = np.zeros(allc.shape[0], np.int32)
    need[:30] = 1
    cutoffs = [infl, 0.8 * infl, 0.5 * infl, 0.1]
    lists = ops.build_neighbors_multi(allc, 30, infl, cutoffs)
    assert len(lists) == len(cutoffs)
    for c, (counts, offsets, indices, maxn) in zip(cutoffs, lists):
        oc, oo, oi = port.build_neighbors(allc.cpu().numpy(), infl, need, cutoff=c)
        assert np.array_equal(counts.cpu().numpy(), oc)
        assert np.array_equal(offsets.cpu().numpy(), oo)
        assert np.array_equal(indices.cpu().numpy(), orc.canonical(oo, oi))
        assert maxn == int(oc.max())
    assert np.array_equal(lists[0][2].cpu().numpy(), full["indices"].cpu().numpy())
    assert lists[-1][2].numel() == 0                      # nothing within 0.1 A
    with pytest.raises(ValueError):
        ops.build_neighbors_multi(allc, 30, infl, [2 * infl])


def test_neighbor_errors(dev):
    bad = torch.tensor([[0.0, 0, 0], [1e-6, 0, 0], [1.0, 1, 1]], dtype=torch.float64, device=dev)
    with pytest.raises(_lib.KliffB200Error) as e:
        ops.build_neighbors(bad, 3, 3.0)
    assert e.value.code == _lib.KB_ERR_REFERENCE
    with pytest.raises(_lib.KliffB200Error) as e:
        ops.create_paddings(3.0, np.zeros((3, 3)), [1, 1, 1], bad)
    assert e.value.code == _lib.KB_ERR_REFERENCE
    # empty input
    counts, offsets, indices, maxn = ops.build_neighbors(bad[:0], 0, 3.0)
    assert counts.numel() == 0 and indices.numel() == 0 and maxn == 0


# --------------------------------------------------------------------------------------------
@pytest.mark.parametrize("mode", [_lib.KB_SYMFUN_AUTO, _lib.KB_SYMFUN_TILED, _lib.KB_SYMFUN_GENERAL])
@pytest.mark.parametrize("name", RUNS)
def test_symfun_vs_recorded_reference(dev, port, name, mode):
    g = load_golden(name)
    fams = fams_from_npz(g)
    infl = float(np.max(g["rcut"]))
    r = gpu_neighbors(dev, g["cell"], g["pbc"], g["coords"], g["z"], infl)
    n = r["n"]
    code = torch.tensor(g["code"][r["image"].cpu().numpy()].astype(np.int32), device=dev)
    P = ops.build_params(hp_from_fams(fams), g["rcut"])
    zeta, dz, dz_ptr = ops.symfun_forward(P, r["allc"], code, r["offsets"], r["indices"], r["maxn"],
                                          n_centers=n, mode=mode)
    zeta0, _, _ = ops.symfun_forward(P, r["allc"], code, r["offsets"], r["indices"], r["maxn"],
                                     n_centers=n, grad=False, mode=mode)
    zh = zeta.cpu().numpy()
    assert np.allclose(zh, g["zeta"], rtol=TOL, atol=TOL * np.abs(g["zeta"]).max(axis=0))
    assert np.allclose(zeta0.cpu().numpy(), zh, rtol=1e-13, atol=1e-13 * np.abs(zh).max())
    D = P.n_desc
    counts = r["counts"].cpu().numpy()
    blocks = blocks_from_packed(dz, dz_ptr, counts, D)
    # recorded blocks use the reference's raw neighbor order; ours are ascending: permute slots
    off = g["nl_offsets"]
    for a in g["keep_atoms"]:
        raw = g["nl_indices"][off[a]:off[a + 1]]
        order = np.argsort(raw, kind="stable")
        ref = np.concatenate([g[f"dz_{a}"][:, order, :], g[f"dz_{a}"][:, -1:, :]], axis=1)
        assert rowwise_rel_err(blocks[a], ref) <= TOL, (name, a)
    # all blocks at once through the force contraction with the recorded weights
    w = torch.tensor(g["w"], device=dev)
    pack = dict(offsets=r["offsets"], indices=r["indices"], image=r["image"], dz=dz, dz_ptr=dz_ptr,
                D=D, n_out=n)
    F = ops.force_scatter(w, pack).cpu().numpy().reshape(n, 3)
    assert np.max(np.abs(F - g["forces_w"])) <= TOL * np.max(np.abs(g["forces_w"]))
    # translation invariance: the slots of every (atom, descriptor) row sum to zero
    for b in blocks:
        s = np.abs(b.sum(axis=1)).max(axis=1)
        assert np.all(s <= 1e-12 * np.maximum(np.abs(b).max(axis=(1, 2)), 1e-300))


def test_kat_symmetry_function(dev):
    """Reference tests/descriptors/test_symmetry_function.py:294-307 (np.allclose defaults):
    zeta 8x9, dzetadr_forces[0] 9x24, dzetadr_stress[0] 9x6, through the dense-fold kernel."""
    g = load_golden("ref_kat_symfun.npz")
    fams = fams_from_npz(g)
    rc = float(g["cutoff"])
    r = gpu_neighbors(dev, g["cell"], g["pbc"], g["coords"], g["z"], rc)
    n = r["n"]
    code = torch.zeros(r["allc"].shape[0], dtype=torch.int32, device=dev)
    P = ops.build_params(hp_from_fams(fams), np.array([[rc]]))
    zeta, dz, dz_ptr = ops.symfun_forward(P, r["allc"], code, r["offsets"], r["indices"], r["maxn"],
                                          n_centers=n)
    assert np.allclose(zeta.cpu().numpy(), g["zeta_ref"])
    pack = dict(offsets=r["offsets"], indices=r["indices"], image=r["image"], dz=dz, dz_ptr=dz_ptr,
                D=P.n_desc, n_out=n)
    dense, stress = ops.dense_fold(pack, r["allc"], n, True, True)
    assert np.allclose(dense[0].cpu().numpy(), g["dzetadr_forces_ref"])
    assert np.allclose(stress[0].cpu().numpy(), g["dzetadr_stress_ref"])


@pytest.mark.parametrize("hpname", ["set51", "set30"])
@pytest.mark.parametrize("mode", [_lib.KB_SYMFUN_SPLIT, _lib.KB_SYMFUN_QUEUE, _lib.KB_SYMFUN_WARP, _lib.KB_SYMFUN_TILED])
def test_symfun_512_atoms_vs_oracle(dev, port, hpname, mode):
    """4x4x4 jittered diamond (512 atoms, 1 600+ paddings): every block against the oracle."""
    hp = get_set51() if hpname == "set51" else get_set30()
    rng = np.random.default_rng(5)
    a, nx = 5.431, 4
    basis = np.array([[0, 0, 0], [0, .5, .5], [.5, 0, .5], [.5, .5, 0],
                      [.25, .25, .25], [.25, .75, .75], [.75, .25, .75], [.75, .75, .25]])
    cells = np.array([[i, j, k] for i in range(nx) for j in range(nx) for k in range(nx)])
    pos = (cells[:, None, :] + basis[None]).reshape(-1, 3) * a + rng.normal(0, 0.05, (512, 3))
    cell = np.eye(3) * a * nx
    z = np.full(512, 14, np.int32)
    r = gpu_neighbors(dev, cell, [1, 1, 1], pos, z, 5.0)
    opc, ops_, opi = port.create_paddings(5.0, cell, [1, 1, 1], pos, z)
    assert np.array_equal(opc, r["pc"].cpu().numpy())
    allc = np.concatenate([pos, opc])
    need = np.zeros(len(allc), np.int32); need[:512] = 1
    oc, oo, oi = port.build_neighbors(allc, 5.0, need)
    oi = orc.canonical(oo, oi)
    assert np.array_equal(oi, r["indices"].cpu().numpy())
    rc = np.array([[5.0]])
    P = ops.build_params(hp, rc)
    code = torch.zeros(len(allc), dtype=torch.int32, device=dev)
    zeta, dz, dz_ptr = ops.symfun_forward(P, r["allc"], code, r["offsets"], r["indices"], r["maxn"],
                                          n_centers=512, mode=mode)
    zeta0, _, _ = ops.symfun_forward(P, r["allc"], code, r["offsets"], r["indices"], r["maxn"],
                                     n_centers=512, grad=False, mode=mode)
    assert torch.allclose(zeta0, zeta, rtol=1e-13, atol=0)
    oz, og = port.symfun_range(rc, orc.fams_from_hyperparams(hp), 0, 512, allc,
                               np.zeros(len(allc), np.int32), oo, oi)
    assert np.max(np.abs(zeta.cpu().numpy() - oz) / np.abs(oz)) <= TOL
    D = P.n_desc
    blocks = blocks_from_packed(dz, dz_ptr, oc, D)
    worst = 0.0
    for t in range(512):
        ref = og[3 * D * int(oo[t] + t):3 * D * int(oo[t + 1] + t + 1)].reshape(D, -1, 3)
        worst = max(worst, rowwise_rel_err(blocks[t], ref))
    assert worst <= TOL, worst
    # padding of the packed layout is zero-filled and 128-byte aligned
    ptr = dz_ptr.cpu().numpy()
    assert np.all(ptr % 16 == 0)
    dzh = dz.cpu().numpy()
    for t in (0, 100, 511):
        assert np.all(dzh[ptr[t] + 3 * D * (oc[t] + 1):ptr[t + 1]] == 0.0)


def test_symfun_centers_subset_and_many_neighbors(dev, port):
    """(a) an explicit centre list; (b) a dense lattice with n > 64 neighbors, which the tiled
    kernel hands to the general kernel."""
    rng = np.random.default_rng(9)
    cell = np.eye(3) * 9.0
    nx = 5
    pos = (np.array([[i, j, k] for i in range(nx) for j in range(nx) for k in range(nx)]) + 0.5) * (9.0 / nx)
    pos = pos + rng.normal(0, 0.05, pos.shape)
    z = np.full(len(pos), 14, np.int32)
    hp = get_set30()
    rc = np.array([[5.0]])
    r = gpu_neighbors(dev, cell, [1, 1, 1], pos, z, 5.0)
    assert r["maxn"] > 64
    allc = r["allc"].cpu().numpy()
    off, idx = r["offsets"].cpu().numpy(), r["indices"].cpu().numpy()
    P = ops.build_params(hp, rc)
    code = torch.zeros(len(allc), dtype=torch.int32, device=dev)
    centers = torch.tensor([3, 77, 124, 0], dtype=torch.int32, device=dev)
    zeta, dz, dz_ptr = ops.symfun_forward(P, r["allc"], code, r["offsets"], r["indices"], r["maxn"],
                                          centers=centers)
    D = P.n_desc
    ptr = dz_ptr.cpu().numpy()
    dzh = dz.cpu().numpy()
    for t, a in enumerate(centers.cpu().numpy()):
        nb = idx[off[a]:off[a + 1]]
        oz, og = port.symfun_atom(rc, orc.fams_from_hyperparams(hp), a, allc,
                                  np.zeros(len(allc), np.int32), nb)
        assert np.max(np.abs(zeta[t].cpu().numpy() - oz) / np.abs(oz)) <= TOL
        blk = dzh[ptr[t]:ptr[t] + 3 * D * (len(nb) + 1)].reshape(D, -1, 3)
        assert rowwise_rel_err(blk, og) <= TOL


def test_force_scatter_forward_backward(dev):
    """fwd against numpy; bwd is the exact adjoint (<gF, fwd(w)> == <bwd(gF), w>), scale folds in;
    autograd double-backward works (needed for create_graph=True training)."""
    g = load_golden("ref_run_sic64_set51.npz")
    fams = fams_from_npz(g)
    r = gpu_neighbors(dev, g["cell"], g["pbc"], g["coords"], g["z"], float(np.max(g["rcut"])))
    n = r["n"]
    code = torch.tensor(g["code"][r["image"].cpu().numpy()].astype(np.int32), device=dev)
    P = ops.build_params(hp_from_fams(fams), g["rcut"])
    zeta, dz, dz_ptr = ops.symfun_forward(P, r["allc"], code, r["offsets"], r["indices"], r["maxn"],
                                          n_centers=n)
    D = P.n_desc
    rng = np.random.default_rng(2)
    scale = torch.tensor(rng.uniform(0.5, 2.0, D), device=dev)
    pack = dict(offsets=r["offsets"], indices=r["indices"], image=r["image"], dz=dz, dz_ptr=dz_ptr,
                D=D, n_out=n, scale=scale)
    w = torch.tensor(rng.normal(size=(n, D)), device=dev, requires_grad=True)
    F = ops.force_scatter(w, pack)
    counts = r["counts"].cpu().numpy()
    blocks = blocks_from_packed(dz, dz_ptr, counts, D)
    off, idx = r["offsets"].cpu().numpy(), r["indices"].cpu().numpy()
    lists = [idx[off[t]:off[t + 1]] for t in range(n)]
    Fo = orc.forces_sparse(w.detach().cpu().numpy() * scale.cpu().numpy(), blocks, lists,
                           r["image"].cpu().numpy(), n)
    assert np.max(np.abs(F.detach().cpu().numpy() - Fo)) <= 1e-12 * np.max(np.abs(Fo))
    gF = torch.tensor(rng.normal(size=3 * n), device=dev)
    (gw,) = torch.autograd.grad(F, w, gF, create_graph=True)
    lhs = float((gF * F.detach()).sum())
    rhs = float((gw.detach() * w.detach()).sum())
    assert abs(lhs - rhs) <= 1e-11 * abs(lhs)
    # double backward: d/d(gF) of <gw, v> = fwd(v)
    gF2 = gF.clone().requires_grad_(True)
    gw2 = ops._ForceGather.apply(gF2, pack)
    v = torch.tensor(rng.normal(size=(n, D)), device=dev)
    (back,) = torch.autograd.grad(gw2, gF2, v)
    assert torch.allclose(back, ops.force_scatter(v, pack), rtol=1e-12, atol=1e-12)


def test_force_assembly_is_bit_reproducible(dev):
    """The owner-sorted force assembly (kb_scatter_plan_* + kb_force_scatter_fwd_seg, the default) gives
    bit-identical forces run after run and after rebuilding the plan, equals the atomic-add form
    (kb_force_scatter_fwd) to rounding, and a sub-range of whole configurations through the parent's plan
    equals the full result restricted to it.  Plan invariants: every slot appears exactly once, ascending
    inside its owner's segment."""
    g = load_golden("ref_run_si64_set51.npz")
    fams = fams_from_npz(g)
    r = gpu_neighbors(dev, g["cell"], g["pbc"], g["coords"], g["z"], float(np.max(g["rcut"])))
    n = r["n"]
    code = torch.tensor(g["code"][r["image"].cpu().numpy()].astype(np.int32), device=dev)
    P = ops.build_params(hp_from_fams(fams), g["rcut"])
    zeta, dz, dz_ptr = ops.symfun_forward(P, r["allc"], code, r["offsets"], r["indices"], r["maxn"], n_centers=n)
    D = P.n_desc
    base = dict(offsets=r["offsets"], indices=r["indices"], image=r["image"], dz=dz, dz_ptr=dz_ptr, D=D, n_out=n)
    w = torch.tensor(np.random.default_rng(5).normal(size=(n, D)), device=dev)
    runs = []
    for _ in range(4):
        pack = dict(base)  # fresh plan every time
        runs.append(ops._scatter_fwd(w, pack))
        runs.append(ops._scatter_fwd(w, pack))  # cached plan
    for F in runs[1:]:
        assert torch.equal(F, runs[0])
    atomic = ops._scatter_fwd(w, dict(base, deterministic=False))
    assert float((atomic - runs[0]).abs().max()) <= 1e-13 * float(runs[0].abs().max())
    plan = pack["plan"]
    counts = r["counts"].cpu().numpy()[:n]
    assert plan["n_slots"] == int(counts.sum()) + n
    seg_ptr, seg_src = plan["seg_ptr"].cpu().numpy(), plan["seg_src"].cpu().numpy()[:plan["n_slots"]]
    assert seg_ptr[0] == 0 and seg_ptr[-1] == plan["n_slots"]
    assert np.array_equal(np.sort(seg_src), np.arange(plan["n_slots"]))
    assert all(np.all(np.diff(seg_src[seg_ptr[o]:seg_ptr[o + 1]]) > 0) for o in range(n))
    # owners of the slots in segment o are o
    off, idx, image = r["offsets"].cpu().numpy(), r["indices"].cpu().numpy(), r["image"].cpu().numpy()
    slot_ptr = plan["slot_ptr"].cpu().numpy()
    owner_of = np.concatenate([np.concatenate([image[idx[off[t]:off[t + 1]]], [t]]) for t in range(n)])
    for o in (0, n // 2, n - 1):
        assert np.all(owner_of[seg_src[seg_ptr[o]:seg_ptr[o + 1]]] == o)
    assert slot_ptr[-1] == plan["n_slots"]


def test_float32_block_storage(dev, port):
    """KB_F32: arithmetic stays fp64, the finished blocks are rounded to float on the way out
    (the reference's default dtype=np.float32 cast, descriptor.py:197-205).  Every kernel that reads
    or writes blocks is checked against its fp64 twin at float precision."""
    g = load_golden("ref_run_si64_set51.npz")
    fams = fams_from_npz(g)
    r = gpu_neighbors(dev, g["cell"], g["pbc"], g["coords"], g["z"], 5.0)
    n = r["n"]
    code = torch.zeros(r["allc"].shape[0], dtype=torch.int32, device=dev)
    P = ops.build_params(hp_from_fams(fams), g["rcut"])
    ref_z, ref_dz, ptr = ops.symfun_forward(P, r["allc"], code, r["offsets"], r["indices"], r["maxn"], n_centers=n)
    for mode in (_lib.KB_SYMFUN_SPLIT, _lib.KB_SYMFUN_QUEUE, _lib.KB_SYMFUN_WARP, _lib.KB_SYMFUN_TILED, _lib.KB_SYMFUN_GENERAL):
        z, dz, ptr32 = ops.symfun_forward(P, r["allc"], code, r["offsets"], r["indices"], r["maxn"],
                                          n_centers=n, mode=mode, dz_dtype=torch.float32)
        assert dz.dtype == torch.float32 and torch.equal(ptr32, ptr)
        assert torch.allclose(z, ref_z, rtol=1e-12, atol=0)  # zeta is fp64 either way
        scale = float(ref_dz.abs().max())
        if mode == _lib.KB_SYMFUN_GENERAL:  # accumulates in float atomics
            assert float((dz.double() - ref_dz).abs().max()) <= 2e-6 * scale
        else:                                # one rounding per element
            assert torch.equal(dz, ref_dz.float())
    pack64 = dict(offsets=r["offsets"], indices=r["indices"], image=r["image"], dz=ref_dz, dz_ptr=ptr, D=51, n_out=n)
    pack32 = dict(pack64, dz=ref_dz.float())
    w = torch.tensor(g["w"], device=dev)
    F64, F32 = ops._scatter_fwd(w, pack64), ops._scatter_fwd(w, pack32)
    assert float((F64 - F32).abs().max()) <= 1e-6 * float(F64.abs().max())
    gF = torch.randn(3 * n, dtype=torch.float64, device=dev)
    b64, b32 = ops._scatter_bwd(gF, pack64), ops._scatter_bwd(gF, pack32)
    assert float((b64 - b32).abs().max()) <= 1e-6 * float(b64.abs().max())
    d64, s64 = ops.dense_fold(pack64, r["allc"], n, True, True)
    d32, s32 = ops.dense_fold(pack32, r["allc"], n, True, True)
    assert float((d64 - d32).abs().max()) <= 1e-6 * float(d64.abs().max())
    assert float((s64 - s32).abs().max()) <= 1e-6 * float(s64.abs().max())


def test_edge_cases_radial_only_isolated_atom_nonperiodic(dev, port):
    """Radial-only hyper-parameters (no angular group), an isolated atom (empty neighbor list, its
    block is just the centre slot), a non-periodic cluster and an empty centre list."""
    from collections import OrderedDict
    hp = OrderedDict()
    hp["g2"] = [{"eta": 0.3, "Rs": 0.0}, {"eta": 1.0, "Rs": 1.0}]
    hp["g1"] = None
    rng = np.random.default_rng(12)
    pos = np.concatenate([rng.random((20, 3)) * 6.0 + 2.0, [[40.0, 40.0, 40.0]]])  # last atom isolated
    cell = np.eye(3) * 60.0
    z = np.full(len(pos), 14, np.int32)
    rc = np.array([[4.0]])
    for hyper in (hp, get_set30()):
        r = gpu_neighbors(dev, cell, [0, 0, 0], pos, z, 4.0)
        assert r["pc"].shape[0] == 0 and int(r["counts"][-1]) == 0
        P = ops.build_params(hyper, rc)
        code = torch.zeros(len(pos), dtype=torch.int32, device=dev)
        fams = orc.fams_from_hyperparams(hyper)
        off, idx = r["offsets"].cpu().numpy(), r["indices"].cpu().numpy()
        for mode in (_lib.KB_SYMFUN_AUTO, _lib.KB_SYMFUN_TILED, _lib.KB_SYMFUN_GENERAL):
            zeta, dz, ptr = ops.symfun_forward(P, r["allc"], code, r["offsets"], r["indices"], r["maxn"],
                                               n_centers=len(pos), mode=mode)
            ptr_h, dz_h = ptr.cpu().numpy(), dz.cpu().numpy()
            for a in range(len(pos)):
                nb = idx[off[a]:off[a + 1]]
                oz, og = port.symfun_atom(rc, fams, a, pos, np.zeros(len(pos), np.int32), nb)
                assert np.allclose(zeta[a].cpu().numpy(), oz, rtol=TOL, atol=TOL * max(1.0, np.abs(oz).max()))
                blk = dz_h[ptr_h[a]:ptr_h[a] + 3 * P.n_desc * (len(nb) + 1)].reshape(P.n_desc, -1, 3)
                assert rowwise_rel_err(blk, og) <= TOL
            assert np.all(zeta[-1].cpu().numpy() == 0.0)  # the isolated atom
        # no centres at all
        zeta, dz, ptr = ops.symfun_forward(P, r["allc"], code, r["offsets"], r["indices"], r["maxn"],
                                           centers=torch.zeros(0, dtype=torch.int32, device=dev))
        assert zeta.shape == (0, P.n_desc) and dz.numel() == 0


def test_unknown_species_and_bad_hyperparams(dev):
    from kliff_b200.dataset import Configuration
    from kliff_b200.legacy.descriptors import SymmetryFunction, SymmetryFunctionError
    with pytest.raises(SymmetryFunctionError):
        SymmetryFunction({"Si-Si": 5.0}, "poly", "set51")
    with pytest.raises(SymmetryFunctionError):
        SymmetryFunction({"Si-Si": 5.0}, "cos", "set99")
    with pytest.raises(SymmetryFunctionError):
        SymmetryFunction({"Si-Si": 5.0}, "cos", {"g7": [{"eta": 1.0}]})
    desc = SymmetryFunction({"Si-Si": 5.0}, "cos", "set30")
    conf = Configuration(cell=np.eye(3) * 10, species=["Si", "Ge"], coords=np.array([[0.0, 0, 0], [2.0, 0, 0]]),
                         PBC=[True] * 3)
    with pytest.raises(SymmetryFunctionError):
        desc.transform(conf)


@pytest.mark.parametrize("seed", range(8))
def test_random_differential(dev, port, seed):
    """Randomised differential test of the whole path against the oracle: random triclinic cells,
    1-3 species with different pair cutoffs, random PBC, random subsets of the five families with
    random (also non-canonical) hyper-parameters.  Neighbor system bit-exact, zeta / dG/dr 1e-10."""
    from collections import OrderedDict
    rng = np.random.default_rng(100 + seed)
    cell = np.diag(rng.uniform(5.0, 9.0, 3)) + rng.normal(0, 0.6, (3, 3))
    n = int(rng.integers(6, 48))
    frac = rng.random((n, 3))
    pos = frac @ cell
    # push apart atoms that are too close (the reference rejects r < 1e-5, and r ~ 0.3 A is unphysical)
    for _ in range(50):
        d = np.linalg.norm(pos[:, None] - pos[None], axis=-1) + np.eye(n) * 10
        i, j = np.unravel_index(np.argmin(d), d.shape)
        if d[i, j] > 0.9:
            break
        pos[j] += rng.normal(0, 0.8, 3)
    nsp = int(rng.integers(1, 4))
    code = rng.integers(0, nsp, n).astype(np.int32)
    zmap = np.array([14, 6, 8], np.int32)
    rcut = rng.uniform(2.6, 4.4, (nsp, nsp))
    rcut = (rcut + rcut.T) / 2
    pbc = rng.integers(0, 2, 3)
    hp = OrderedDict()
    fam_order = list(rng.permutation(["g1", "g2", "g3", "g4", "g5"]))[: int(rng.integers(1, 6))]
    canonical = bool(rng.integers(0, 2))
    for f in fam_order:
        k = int(rng.integers(1, 5))
        if f == "g1":
            hp[f] = None
        elif f == "g2":
            hp[f] = [{"eta": float(rng.uniform(0.01, 1.5)), "Rs": float(rng.uniform(0, 2))} for _ in range(k)]
        elif f == "g3":
            hp[f] = [{"kappa": float(rng.uniform(0.1, 2.0))} for _ in range(k)]
        else:
            rows = []
            for _ in range(k):
                if canonical and f == "g4":
                    rows.append({"zeta": float(rng.choice([1, 2, 4, 16])), "lambda": float(rng.choice([-1, 1])),
                                 "eta": float(rng.choice([0.01, 0.05, 0.2]))})
                else:
                    rows.append({"zeta": float(rng.uniform(0.5, 6.0)), "lambda": float(rng.uniform(-1, 1)),
                                 "eta": float(rng.uniform(0.0, 0.3))})
            hp[f] = rows
    infl = float(rcut.max())
    r = gpu_neighbors(dev, cell, pbc, pos, zmap[code], infl)
    opc, ops_, opi = port.create_paddings(infl, cell, pbc, pos, zmap[code])
    assert np.array_equal(opc, r["pc"].cpu().numpy()) and np.array_equal(opi, r["pi"].cpu().numpy())
    allc = np.concatenate([pos, opc]) if len(opc) else pos
    need = np.zeros(len(allc), np.int32)
    need[:n] = 1
    oc, oo, oi = port.build_neighbors(allc, infl, need)
    oi = orc.canonical(oo, oi)
    assert np.array_equal(oi, r["indices"].cpu().numpy()) and np.array_equal(oc, r["counts"].cpu().numpy())
    image = np.concatenate([np.arange(n), opi]).astype(np.int64)
    allcode = code[image].astype(np.int32)
    P = ops.build_params(hp, rcut)
    zeta, dz, ptr = ops.symfun_forward(P, r["allc"], torch.tensor(allcode, device=dev), r["offsets"],
                                       r["indices"], r["maxn"], n_centers=n)
    fams = orc.fams_from_hyperparams(hp)
    oz, og = port.symfun_range(rcut, fams, 0, n, allc, allcode, oo, oi)
    D = P.n_desc
    scale = np.maximum(np.abs(oz).max(axis=0), 1e-300)
    assert np.max(np.abs(zeta.cpu().numpy() - oz) / scale) <= TOL
    ptr_h, dz_h = ptr.cpu().numpy(), dz.cpu().numpy()
    for t in range(n):
        ref = og[3 * D * int(oo[t] + t):3 * D * int(oo[t + 1] + t + 1)].reshape(D, -1, 3)
        blk = dz_h[ptr_h[t]:ptr_h[t] + ref.size].reshape(ref.shape)
        # rows whose gradient vanishes identically are compared absolutely
        rs = np.abs(ref).reshape(D, -1).max(axis=1)
        err = np.abs(blk - ref).reshape(D, -1).max(axis=1)
        assert np.all(err <= TOL * np.maximum(rs, 1e-12)), (seed, t, float((err / np.maximum(rs, 1e-12)).max()))


def _random_cell_config(rng, n_lo=1, n_hi=70):
    cell = np.diag(rng.uniform(4.0, 11.0, 3)) + rng.normal(0, 0.7, (3, 3))
    n = int(rng.integers(n_lo, n_hi))
    pos = rng.random((n, 3)) @ cell
    for _ in range(80):
        if n < 2:
            break
        d = np.linalg.norm(pos[:, None] - pos[None], axis=-1) + np.eye(n) * 10
        i, j = np.unravel_index(np.argmin(d), d.shape)
        if d[i, j] > 0.8:
            break
        pos[j] += rng.normal(0, 0.8, 3)
    return cell, pos, rng.integers(0, 2, 3)


def test_batched_neighbors_bit_exact(dev, port):
    """kb_batch_* (one CTA per configuration) against the oracle, configuration by configuration:
    image-atom coordinates, owners, counts and neighbor indices all bit-exact.  Mix of random
    triclinic cells / PBC flags, the reference's Si and SiC training cells, and an empty cell."""
    from helpers import configs_from_npz
    rng = np.random.default_rng(77)
    items = [_random_cell_config(rng) for _ in range(40)]
    for c in configs_from_npz("si_varying_alat.npz", limit=6) + configs_from_npz("sic_training_set.npz", limit=4):
        items.append((np.asarray(c.cell, np.float64), np.asarray(c.coords, np.float64), np.asarray(c.PBC, int)))
    items.insert(7, (np.eye(3) * 5.0, np.zeros((0, 3)), np.array([1, 1, 1])))  # no atoms at all
    infl = 4.3
    cfg_ptr = np.zeros(len(items) + 1, np.int64)
    np.cumsum([len(p) for _, p, _ in items], out=cfg_ptr[1:])
    coords = torch.tensor(np.concatenate([p for _, p, _ in items]), dtype=torch.float64, device=dev)
    out = ops.batch_neighbors(cfg_ptr, np.stack([c for c, _, _ in items]),
                              np.stack([b for _, _, b in items]), coords, infl)
    ap = out["atom_ptr"]
    gc, gi = out["coords"].cpu().numpy(), out["image"].cpu().numpy()
    goff, gidx, gcen = out["offsets"].cpu().numpy(), out["indices"].cpu().numpy(), out["centers"].cpu().numpy()
    maxn = 0
    for c, (cell, pos, pbc) in enumerate(items):
        n = len(pos)
        a0, a1 = int(ap[c]), int(ap[c + 1])
        if n == 0:
            assert a0 == a1
            continue
        opc, _, opi = port.create_paddings(infl, cell, pbc, pos, np.full(n, 14, np.int32))
        assert a1 - a0 == n + len(opc), f"config {c}: image count"
        assert np.array_equal(gc[a0:a0 + n], pos)
        assert np.array_equal(gc[a0 + n:a1], opc.reshape(-1, 3)), f"config {c}: image coordinates"
        assert np.array_equal(gi[a0:a0 + n], cfg_ptr[c] + np.arange(n))
        assert np.array_equal(gi[a0 + n:a1], cfg_ptr[c] + opi), f"config {c}: owners"
        assert np.array_equal(gcen[cfg_ptr[c]:cfg_ptr[c + 1]], a0 + np.arange(n))
        allc = np.concatenate([pos, opc.reshape(-1, 3)])
        need = np.zeros(len(allc), np.int32)
        need[:n] = 1
        oc, oo, oi = port.build_neighbors(allc, infl, need)
        oi = orc.canonical(oo, oi)
        cnt = np.diff(goff[a0:a1 + 1])
        assert np.array_equal(cnt, oc), f"config {c}: counts"
        assert np.array_equal(gidx[goff[a0]:goff[a1]] - a0, oi), f"config {c}: indices"
        maxn = max(maxn, int(oc.max()))
    assert out["maxn"] == maxn
    # errors: singular cell and colliding atoms are the reference's own failures
    with pytest.raises(_lib.KliffB200Error):
        ops.batch_neighbors([0, 1], np.zeros((1, 3, 3)), np.ones((1, 3)), coords[:1], infl)
    two = torch.tensor([[0.0, 0, 0], [0, 0, 1e-7]], dtype=torch.float64, device=dev)
    with pytest.raises(_lib.KliffB200Error):
        ops.batch_neighbors([0, 2], np.eye(3)[None] * 6.0, np.ones((1, 3)), two, infl)


def test_batched_descriptor_equals_per_configuration(dev):
    """SymmetryFunction.transform_packed over a list (batched kernels) gives exactly what the
    per-configuration cell-list path gives: same atoms, lists, zeta and gradient blocks."""
    from helpers import configs_from_npz
    from kliff_b200.legacy.descriptors import symmetry_function as sfm
    from kliff_b200.legacy.descriptors.symmetry_function import SymmetryFunction
    desc = SymmetryFunction({"Si-Si": 5.0, "Si-C": 4.5, "C-C": 4.0}, "cos", "set30")
    confs = configs_from_npz("sic_training_set.npz", limit=5) + configs_from_npz("si_varying_alat.npz", limit=3)
    a = desc.transform_packed(confs, grad=True, device=dev)
    old = sfm._BATCH_MAX_ATOMS
    sfm._BATCH_MAX_ATOMS = 0
    try:
        b = desc.transform_packed(confs, grad=True, device=dev)
    finally:
        sfm._BATCH_MAX_ATOMS = old
    for name in ("coords", "code", "image", "offsets", "indices", "centers", "zeta", "dz", "dz_ptr"):
        assert torch.equal(getattr(a, name), getattr(b, name)), name
    assert np.array_equal(a.atom_ptr, b.atom_ptr) and np.array_equal(a.cfg_ptr, b.cfg_ptr)
    assert a.maxn == b.maxn


def test_kernel_timing_records_stage_launches(dev):
    """kb_kernel_timing brackets the descriptor kernels with CUDA events on their stream (what
    bench.py's roofline uses): AUTO = one table launch + one sweep launch for a small system."""
    g = load_golden("ref_run_si64_set51.npz")
    rc = float(np.max(g["rcut"]))
    r = gpu_neighbors(dev, g["cell"], g["pbc"], g["coords"], g["z"], rc)
    P = ops.build_params(get_set51(), np.array([[rc]]))
    code = torch.zeros(r["allc"].shape[0], dtype=torch.int32, device=dev)
    ops.kernel_timing(True)
    ops.symfun_forward(P, r["allc"], code, r["offsets"], r["indices"], r["maxn"], n_centers=r["n"])
    ops.symfun_forward(P, r["allc"], code, r["offsets"], r["indices"], r["maxn"], n_centers=r["n"],
                       mode=_lib.KB_SYMFUN_WARP)
    t = ops.kernel_timing(False)
    assert t["prep"][1] == 1 and t["sweep"][1] == 1 and t["fused"][1] == 1
    assert all(0.0 < t[k][0] < 1000.0 for k in ("prep", "sweep", "fused")) and t["scatter"][1] == 0
    ops.symfun_forward(P, r["allc"], code, r["offsets"], r["indices"], r["maxn"], n_centers=r["n"])
    assert ops.kernel_timing(False)["sweep"][1] == 0  # recording is off again


def test_streamed_host_descriptors_equal_single_pass(dev):
    """PackedFingerprints.compute(zeta_host=...) runs the centre list in pieces and streams each
    piece's descriptors to pinned host memory on a side stream: same zeta, same blocks."""
    from helpers import configs_from_npz
    from kliff_b200.legacy.descriptors.symmetry_function import SymmetryFunction
    from kliff_b200.packed import PackedFingerprints
    desc = SymmetryFunction({"Si-Si": 5.0, "Si-C": 4.5, "C-C": 4.0}, "cos", "set51")
    confs = configs_from_npz("sic_training_set.npz", limit=4) + configs_from_npz("si_varying_alat.npz", limit=3)
    ref = desc.transform_packed(confs, grad=True, device=dev)
    n = ref.n_centers
    host = torch.empty((n, len(desc)), dtype=torch.float64).pin_memory()
    host.fill_(-1.0)
    p = PackedFingerprints.from_configs(len(desc), dev, confs, desc.species_code, 5.0)
    p.compute(desc._params, grad=True, zeta_host=host, piece=50)
    p.host_ready.synchronize()
    assert torch.equal(p.zeta, ref.zeta) and torch.equal(p.dz_ptr, ref.dz_ptr)
    assert torch.equal(host, ref.zeta.cpu())
    # blocks: compare the used part of every block (alignment padding of the last block of a piece
    # is zero-filled by whichever call owns it)
    assert torch.equal(p.dz, ref.dz)
    # the one-piece path fills the host array too
    host.fill_(-1.0)
    q = desc.transform_packed(confs, grad=False, device=dev, zeta_host=host)
    q.host_ready.synchronize()
    # (the value-only pass runs the round-1 sweep: same sums in another order)
    assert torch.allclose(host, ref.zeta.cpu(), rtol=1e-12, atol=1e-300)


@pytest.mark.parametrize("n,D", [(0, 51), (1, 51), (513, 30), (70001, 51)])
def test_moments_and_normalize_vs_numpy(dev, n, D):
    """kb_moments / kb_normalize against np.mean / np.std / (zeta - mean) / stdev
    (kliff/legacy/descriptors/descriptor.py:116-118, 187-194): <= 1e-13 relative, and two runs
    bit-identical (deterministic two-stage sums)."""
    rng = np.random.default_rng(n + D)
    z = rng.normal(3.0, 2.0, size=(n, D)) * rng.uniform(0.1, 50.0, size=D)
    zd = torch.as_tensor(z, device=dev)
    s1 = ops.moments(zd)
    assert torch.equal(s1, ops.moments(zd))
    if n == 0:
        assert float(s1.abs().max()) == 0.0
        return
    mean = s1 / n
    assert np.allclose(mean.cpu().numpy(), z.mean(axis=0), rtol=1e-13, atol=0)
    s2 = ops.moments(zd, mean)
    assert torch.equal(s2, ops.moments(zd, mean))
    std = torch.sqrt(s2 / n)
    assert np.allclose(std.cpu().numpy(), z.std(axis=0), rtol=1e-12, atol=1e-300)
    if n > 1:
        out = ops.normalize(zd, mean, 1.0 / std)
        want = (z - z.mean(axis=0)) / z.std(axis=0)
        assert np.allclose(out.cpu().numpy(), want, rtol=1e-11, atol=1e-11)


@pytest.mark.skipif(not orc.have_reference(), reason="oracle/_ref not built")
def test_kim_params_file_reference_parser_vs_gpu(dev, tmp_path):
    """write_kim_params -> the reference's Descriptor::read_parameter_file (sym_fn.cpp:108-387) ->
    generate_one_atom on every atom, folded into the dense (N, D, 3N) array (sym_fn.py:163-174), against
    SymmetryFunction.transform on the GPU for the same two-species configuration: <= 1e-10 row-relative."""
    from helpers import config_from_golden
    from kliff_b200.legacy.descriptors import SymmetryFunction
    from test_oracle import _full, _species_codes_as_parsed
    g = load_golden("ref_run_sic64_set51.npz")
    desc = SymmetryFunction({"Si-Si": 5.0, "Si-C": 4.5, "C-C": 4.0}, "cos", "set51", normalize=False,
                            dtype=np.float64)
    desc.write_kim_params(str(tmp_path))
    path = str(tmp_path / "descriptor.params")
    codes = _species_codes_as_parsed(path)
    n, allc, image = _full(g)
    code = np.array([codes["Si" if z == 14 else "C"] for z in g["z"][image]], np.int32)
    idx = orc.canonical(g["nl_offsets"], g["nl_indices"])
    ref = orc.Oracle("reference")
    zeta_r, blocks, _, _ = ref.symfun_atoms_from_file(path, range(n), allc, code, g["nl_offsets"], idx)
    lists = [idx[g["nl_offsets"][a]:g["nl_offsets"][a + 1]] for a in range(n)]
    dense_r = orc.dense_fold(blocks, lists, image, n)
    zeta, dzetadr, _ = desc.transform(config_from_golden(g), fit_forces=True, fit_stress=False)
    assert rowwise_rel_err(zeta, zeta_r) <= TOL
    assert rowwise_rel_err(dzetadr.reshape(n * 51, -1), dense_r.reshape(n * 51, -1)) <= TOL
