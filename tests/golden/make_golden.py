"""
Regenerates every fixture under tests/golden/ from /root/reference (build container only).

  python tests/golden/make_golden.py

1. ref_kat_*.npz   — the reference's OWN known-answer vectors, lifted verbatim (as data) from
                     /root/reference/tests/descriptors/test_symmetry_function.py:9-291 and
                     /root/reference/tests/test_neighbor.py:6-34, together with the parsed input
                     configurations they are quoted on.
2. ref_run_*.npz   — outputs of the UNMODIFIED reference C++ (oracle/_ref/libkliff_ref.so, built
                     by `make -C oracle ref`) run HERE on seeded inputs: paddings, raw neighbor
                     lists, zeta, per-atom dG/dr blocks.
3. published_losses.json — the two tutorial loss sequences (docs/source/tutorials/nn_Si.rst:173-183,
                     nn_SiC.rst:104-114).
4. si_varying_alat.npz / sic_training_set.npz — the two example data sets of the reference
                     (examples/*.tar.gz) parsed into arrays, for the loss-sequence parity tests.

Nothing at test time reads /root/reference; only this script does.
"""
import ast
import io
import json
import os
import sys
import tarfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = "/root/reference"

from kliff_b200.dataset.extxyz import read_extxyz  # noqa: E402
from kliff_b200.legacy.descriptors.hyperparams import get_set30, get_set51  # noqa: E402
from oracle.oracle import Oracle, fams_from_hyperparams  # noqa: E402


def literal_assignments(path, names):
    """Pull `name = <literal>` module-level assignments out of a python file without importing
    it; understands `np.asarray(<list>)` and `np.asarray(<list>).reshape(a, b)` wrappers."""
    tree = ast.parse(open(path).read())
    out = {}
    for node in tree.body:
        if not (isinstance(node, ast.Assign) and len(node.targets) == 1):
            continue
        t = node.targets[0]
        if not (isinstance(t, ast.Name) and t.id in names):
            continue
        v, shape = node.value, None
        if isinstance(v, ast.Call) and isinstance(v.func, ast.Attribute) and v.func.attr == "reshape":
            shape = tuple(ast.literal_eval(a) for a in v.args)
            v = v.func.value
        if isinstance(v, ast.Call):  # np.asarray([...])
            v = v.args[0]
        arr = np.array(ast.literal_eval(v))
        out[t.id] = arr.reshape(shape) if shape else arr
    return out


def pack_fams(fams):
    kinds = np.array([k for k, _ in fams], np.int32)
    rows = np.array([v.shape[0] for _, v in fams], np.int32)
    vals = np.concatenate([v.reshape(-1) for _, v in fams])
    return kinds, rows, vals


Z = {"Si": 14, "C": 6, "O": 8, "Mo": 42, "S": 16}


def cfg_arrays(path):
    cell, species, coords, pbc, energy, forces, stress = read_extxyz(path)
    return dict(cell=cell, coords=coords, pbc=np.array(pbc, np.int32),
                z=np.array([Z[s] for s in species], np.int32))


# --------------------------------------------------------------------------- 1. KATs
def make_kats():
    t = literal_assignments(f"{REF}/tests/descriptors/test_symmetry_function.py",
                            {"zeta_ref", "dzetadr_forces_ref", "dzetadr_stress_ref"})
    from collections import OrderedDict
    hp = OrderedDict()  # test_symmetry_function.py:272-291
    hp["g1"] = None
    hp["g2"] = [{"eta": 0.0009, "Rs": 0.0}, {"eta": 0.01, "Rs": 0.0}]
    hp["g3"] = [{"kappa": 0.03214}, {"kappa": 0.13123}]
    hp["g4"] = [{"zeta": 1, "lambda": -1, "eta": 0.0001}, {"zeta": 2, "lambda": 1, "eta": 0.003}]
    hp["g5"] = [{"zeta": 1, "lambda": -1, "eta": 0.0001}, {"zeta": 2, "lambda": 1, "eta": 0.003}]
    kinds, rows, vals = pack_fams(fams_from_hyperparams(hp))
    c = cfg_arrays(f"{REF}/tests/test_data/configs/Si.xyz")
    np.savez(f"{HERE}/ref_kat_symfun.npz", zeta_ref=np.array(t["zeta_ref"]),
             dzetadr_forces_ref=np.array(t["dzetadr_forces_ref"]),
             dzetadr_stress_ref=np.array(t["dzetadr_stress_ref"]),
             fam_kinds=kinds, fam_rows=rows, fam_vals=vals, cutoff=4.5, **c)

    t = literal_assignments(f"{REF}/tests/test_neighbor.py", {"target_coords", "all_indices"})
    c = cfg_arrays(f"{REF}/tests/test_data/configs/bilayer_graphene/bilayer_sep3.36_i0_j0.xyz")
    c["z"][0] = Z["O"]  # test_neighbor.py:41
    tz = np.full(16, Z["C"], np.int32)
    tz[[0, 10, 12, 14]] = Z["O"]  # test_neighbor.py:27-31
    np.savez(f"{HERE}/ref_kat_neighbor.npz", target_coords=np.array(t["target_coords"]),
             target_z=tz, all_indices=np.array(t["all_indices"], np.int32), infl_dist=2.0, **c)


# --------------------------------------------------------------------------- 2. reference runs
def diamond(nx, a, sigma, seed, zincblende=False):
    basis = np.array([[0, 0, 0], [0, .5, .5], [.5, 0, .5], [.5, .5, 0],
                      [.25, .25, .25], [.25, .75, .75], [.75, .25, .75], [.75, .75, .25]])
    cells = np.array([[i, j, k] for i in range(nx) for j in range(nx) for k in range(nx)])
    pos = (cells[:, None, :] + basis[None, :, :]).reshape(-1, 3) * a
    rng = np.random.default_rng(seed)
    pos = pos + rng.normal(0.0, sigma, pos.shape)
    sp = np.zeros(len(pos), np.int32)
    if zincblende:
        sp = np.tile(np.array([0, 0, 0, 0, 1, 1, 1, 1], np.int32), len(cells))
    return np.eye(3) * a * nx, pos, sp


def reference_run(name, cell, pbc, coords, z, code, rcut, hp, keep_atoms):
    """z: atomic numbers (padding species), code: species code per contributing atom."""
    ref = Oracle("reference")
    fams = fams_from_hyperparams(hp)
    infl = float(np.max(rcut))
    pc, ps, pi = ref.create_paddings(infl, cell, pbc, coords, z)
    allc = np.concatenate([coords, pc]) if len(pc) else coords.copy()
    n = len(coords)
    image = np.concatenate([np.arange(n, dtype=np.int32), pi])
    allcode = code[image].astype(np.int32)
    need = np.zeros(len(allc), np.int32)
    need[:n] = 1
    counts, offsets, idx = ref.build_neighbors(allc, infl, need)
    zeta, packed = ref.symfun_range(rcut, fams, 0, n, allc, allcode, offsets, idx)
    D = zeta.shape[1]
    blocks = {}
    for a in keep_atoms:
        o = 3 * D * int(offsets[a] + a)
        blocks[f"dz_{a}"] = packed[o:o + 3 * D * (counts[a] + 1)].reshape(D, counts[a] + 1, 3)
    # size-independent digest of ALL blocks: contraction with a fixed pseudo-random dE/dzeta
    w = np.random.default_rng(7).normal(size=(n, D))
    F = np.zeros((n, 3))
    for a in range(n):
        o = 3 * D * int(offsets[a] + a)
        blk = packed[o:o + 3 * D * (counts[a] + 1)].reshape(D, counts[a] + 1, 3)
        ids = np.concatenate([idx[offsets[a]:offsets[a + 1]], [a]])
        np.subtract.at(F, image[ids], np.tensordot(w[a], blk, axes=([0], [0])))
    kinds, rows, vals = pack_fams(fams)
    np.savez_compressed(f"{HERE}/ref_run_{name}.npz", cell=cell, pbc=np.asarray(pbc, np.int32),
                        coords=coords, z=z, code=code, rcut=rcut, fam_kinds=kinds, fam_rows=rows,
                        fam_vals=vals, pad_coords=pc, pad_z=ps, pad_image=pi, nl_counts=counts,
                        nl_offsets=offsets, nl_indices=idx, zeta=zeta, w=w, forces_w=F,
                        keep_atoms=np.array(keep_atoms, np.int32), **blocks)
    print(name, "N", n, "P", len(pc), "nbar", counts[:n].mean(), "D", D)


def make_runs():
    from collections import OrderedDict
    one = np.array([[5.0]])
    # (a) the bench lattice in miniature: jittered diamond Si, set51, rc 5
    cell, pos, sp = diamond(2, 5.431, 0.05, 0)
    reference_run("si64_set51", cell, [1, 1, 1], pos, np.full(64, 14, np.int32), sp, one,
                  get_set51(), [0, 17, 42, 63])
    # (b) set30 on the same lattice
    reference_run("si64_set30", cell, [1, 1, 1], pos, np.full(64, 14, np.int32), sp, one,
                  get_set30(), [5, 60])
    # (c) two species with three different cutoffs (exercises rc_ij / rc_ik / rc_jk tables)
    cell, pos, sp = diamond(2, 4.36, 0.08, 1, zincblende=True)
    rc2 = np.array([[5.0, 4.5], [4.5, 4.0]])
    reference_run("sic64_set51", cell, [1, 1, 1], pos, np.where(sp == 0, 14, 6).astype(np.int32),
                  sp, rc2, get_set51(), [0, 4, 31])
    # (d) compressed 8-atom cell: self-images, n = 34+ (Si_training_set/varying_alat style)
    cell, pos, sp = diamond(1, 4.9, 0.03, 2)
    reference_run("si8_small", cell, [1, 1, 1], pos, np.full(8, 14, np.int32), sp, one,
                  get_set51(), [0, 3, 7])
    # (e) triclinic cell, mixed PBC, all five families incl. non-power-of-two zeta and lambda != +-1
    rng = np.random.default_rng(3)
    cell = np.array([[7.1, 0.3, 0.0], [2.2, 6.4, 0.4], [0.5, 1.1, 9.0]])
    frac = rng.random((30, 3))
    pos = frac @ cell
    # push atoms apart a little so that no pair collides
    hp = OrderedDict()
    hp["g1"] = None
    hp["g2"] = [{"eta": 0.05, "Rs": 0.5}, {"eta": 0.3, "Rs": 1.5}]
    hp["g3"] = [{"kappa": 0.7}]
    hp["g4"] = [{"zeta": 3, "lambda": -1, "eta": 0.01}, {"zeta": 1.5, "lambda": 0.5, "eta": 0.02},
                {"zeta": 8, "lambda": 1, "eta": 0.0}]
    hp["g5"] = [{"zeta": 2, "lambda": -1, "eta": 0.03}, {"zeta": 1, "lambda": 1, "eta": 0.001}]
    sp = (np.arange(30) % 2).astype(np.int32)
    rc2 = np.array([[4.2, 3.7], [3.7, 3.1]])
    reference_run("tri30_allfam", cell, [1, 1, 0], pos, np.where(sp == 0, 42, 16).astype(np.int32),
                  sp, rc2, hp, [0, 1, 15, 29])
    # (f) the reference's own triclinic 288-atom MoS2 fixture (tests/models/test_lennard_jones.py
    #     exercises paddings + neighbor list on it): neighbor data only
    ref = Oracle("reference")
    path = f"{REF}/tests/test_data/configs/MoS2"
    f = sorted(os.listdir(path))[0]
    c = cfg_arrays(os.path.join(path, f))
    pc, ps, pi = ref.create_paddings(5.0, c["cell"], c["pbc"], c["coords"], c["z"])
    allc = np.concatenate([c["coords"], pc])
    need = np.zeros(len(allc), np.int32)
    need[:len(c["coords"])] = 1
    counts, offsets, idx = ref.build_neighbors(allc, 5.0, need)
    np.savez_compressed(f"{HERE}/ref_run_mos2_neigh.npz", pad_coords=pc, pad_z=ps, pad_image=pi,
                        nl_counts=counts, nl_offsets=offsets, nl_indices=idx, infl_dist=5.0, **c)
    print("mos2", len(c["coords"]), len(pc), counts[:288].mean())


# --------------------------------------------------------------------------- 3. published losses
def make_losses():
    si = [7.3307514191e+01, 7.2090658188e+01, 7.1389844894e+01, 7.0744287491e+01,
          7.0117309570e+01, 6.9499519348e+01, 6.8886822701e+01, 6.8277156830e+01,
          6.7668612480e+01, 6.7058616638e+01, 6.6683933258e+01]
    sic = [5.7103012085e+01, 5.6625793457e+01, 5.6168552399e+01, 5.5734367371e+01,
           5.5320205688e+01, 5.4922414780e+01, 5.4534166336e+01, 5.4145248413e+01,
           5.3747819901e+01, 5.3339361191e+01, 5.3061624527e+01]
    # cross-check against the docs text so a typo cannot slip in
    for path, vals in ((f"{REF}/docs/source/tutorials/nn_Si.rst", si),
                       (f"{REF}/docs/source/tutorials/nn_SiC.rst", sic)):
        txt = open(path).read()
        for v in vals:
            assert f"{v:.10e}" in txt, (path, v)
    json.dump({"nn_Si.rst:173-183": si, "nn_SiC.rst:104-114": sic,
               "setup": "set51, cos 5.0 A, 51-10-10-1 tanh (Dropout keep... see tests), Adam lr 1e-3, "
                        "fp32, seed 35, Weight(forces_weight=0.3) for Si; batch 100 (Si) / 4 (SiC)"},
              open(f"{HERE}/published_losses.json", "w"), indent=1)


# --------------------------------------------------------------------------- 4. data sets
def tar_to_npz(tar, prefix, out):
    cells, pbcs, energies, nat, coords, forces, zs, names = [], [], [], [], [], [], [], []
    with tarfile.open(tar) as tf:
        members = sorted((m for m in tf.getmembers() if m.name.endswith(".xyz") and prefix in m.name),
                         key=lambda m: m.name)
        import tempfile
        for m in members:
            data = tf.extractfile(m).read().decode()
            with tempfile.NamedTemporaryFile("w", suffix=".xyz", delete=False) as tmp:
                tmp.write(data)
            cell, species, xyz, pbc, energy, frc, _ = read_extxyz(tmp.name)
            os.unlink(tmp.name)
            cells.append(cell); pbcs.append(pbc); energies.append(energy); nat.append(len(species))
            coords.append(xyz); forces.append(frc); zs.append([Z[s] for s in species])
            names.append(m.name)
    np.savez_compressed(out, cell=np.array(cells), pbc=np.array(pbcs, np.int32),
                        energy=np.array(energies), natoms=np.array(nat, np.int32),
                        coords=np.concatenate(coords), forces=np.concatenate(forces),
                        z=np.concatenate(zs).astype(np.int32), names=np.array(names))
    print(out, len(names), sum(nat))


def make_datasets():
    tar_to_npz(f"{REF}/examples/Si_training_set.tar.gz", "varying_alat", f"{HERE}/si_varying_alat.npz")
    tar_to_npz(f"{REF}/examples/SiC_training_set.tar.gz", "", f"{HERE}/sic_training_set.npz")


if __name__ == "__main__":
    make_kats()
    make_runs()
    make_losses()
    make_datasets()
