"""Pins the CPU oracle (oracle/kb_oracle.c) against (a) the reference's own known-answer
vectors and (b) outputs of the unmodified reference C++ recorded in tests/golden/ref_run_*.npz
(and, when oracle/_ref/libkliff_ref.so is present, against the live reference)."""
import glob
import os

import numpy as np
import pytest

from conftest import GOLDEN, fams_from_npz, load_golden, rowwise_rel_err
from oracle import oracle as orc

RUNS = sorted(os.path.basename(p) for p in glob.glob(os.path.join(GOLDEN, "ref_run_*.npz"))
              if "mos2" not in p)


@pytest.fixture(scope="module")
def port():
    return orc.Oracle("port")


def _full(g):
    n = len(g["coords"])
    allc = np.concatenate([g["coords"], g["pad_coords"]])
    image = np.concatenate([np.arange(n, dtype=np.int32), g["pad_image"]])
    return n, allc, image


def test_kat_neighbor(port):
    """tests/test_neighbor.py:37-64 of the reference: 12 paddings in order, neighbor indices."""
    g = load_golden("ref_kat_neighbor.npz")
    pc, ps, pi = port.create_paddings(float(g["infl_dist"]), g["cell"], g["pbc"], g["coords"], g["z"])
    allc = np.concatenate([g["coords"], pc])
    assert np.allclose(allc, g["target_coords"])
    assert np.array_equal(np.concatenate([g["z"], ps]), g["target_z"])
    need = np.zeros(len(allc), np.int32)
    need[:4] = 1
    counts, offsets, idx = port.build_neighbors(allc, float(g["infl_dist"]), need)
    for i in range(4):
        assert np.array_equal(idx[offsets[i]:offsets[i + 1]], g["all_indices"][i])
    assert np.all(counts[4:] == 0)


def test_kat_dimer(port):
    """tests/test_neighbor.py:67-100: dimer, PBC on/off."""
    cell = np.eye(3) * 200.0
    coords = np.array([[0.0, 0, 0], [2.0, 0, 0]])
    for P in (1, 0):
        pc, ps, pi = port.create_paddings(5.0, cell, [P] * 3, coords, np.array([6, 8], np.int32))
        allc = np.concatenate([coords, pc]) if len(pc) else coords
        need = np.zeros(len(allc), np.int32)
        need[:2] = 1
        counts, offsets, idx = port.build_neighbors(allc, 5.0, need)
        assert list(idx[offsets[0]:offsets[1]]) == [1]
        assert list(idx[offsets[1]:offsets[2]]) == [0]


def test_kat_symmetry_function(port):
    """tests/descriptors/test_symmetry_function.py:294-307: zeta 8x9, dzetadr_forces[0] 9x24,
    dzetadr_stress[0] 9x6 with np.allclose defaults."""
    g = load_golden("ref_kat_symfun.npz")
    fams = fams_from_npz(g)
    rc = np.array([[float(g["cutoff"])]])
    pc, ps, pi = port.create_paddings(float(g["cutoff"]), g["cell"], g["pbc"], g["coords"], g["z"])
    n = len(g["coords"])
    allc = np.concatenate([g["coords"], pc]) if len(pc) else g["coords"]
    image = np.concatenate([np.arange(n, dtype=np.int32), pi])
    need = np.zeros(len(allc), np.int32)
    need[:n] = 1
    counts, offsets, idx = port.build_neighbors(allc, float(g["cutoff"]), need)
    code = np.zeros(len(allc), np.int32)
    zeta, blocks, lists = [], [], []
    for i in range(n):
        nb = idx[offsets[i]:offsets[i + 1]]
        z, dz = port.symfun_atom(rc, fams, i, allc, code, nb)
        z0, _ = port.symfun_atom(rc, fams, i, allc, code, nb, grad=False)
        assert np.allclose(z, z0, rtol=1e-14, atol=0)
        zeta.append(z); blocks.append(dz); lists.append(nb)
    assert np.allclose(np.array(zeta), g["zeta_ref"])
    dense = orc.dense_fold(blocks, lists, image, n)
    assert np.allclose(dense[0], g["dzetadr_forces_ref"])
    stress = orc.stress_fold(blocks, lists, allc, n)
    assert np.allclose(stress[0], g["dzetadr_stress_ref"])


@pytest.mark.parametrize("name", RUNS)
def test_port_vs_recorded_reference(port, name):
    g = load_golden(name)
    fams = fams_from_npz(g)
    infl = float(np.max(g["rcut"]))
    pc, ps, pi = port.create_paddings(infl, g["cell"], g["pbc"], g["coords"], g["z"])
    # integer / byte work: bit-exact, in the reference's own order
    assert np.array_equal(pc, g["pad_coords"])
    assert np.array_equal(ps, g["pad_z"]) and np.array_equal(pi, g["pad_image"])
    n, allc, image = _full(g)
    need = np.zeros(len(allc), np.int32)
    need[:n] = 1
    counts, offsets, idx = port.build_neighbors(allc, infl, need)
    assert np.array_equal(counts, g["nl_counts"])
    assert np.array_equal(offsets, g["nl_offsets"])
    assert np.array_equal(idx, g["nl_indices"])
    code = g["code"][image].astype(np.int32)
    zeta, packed = port.symfun_range(g["rcut"], fams, 0, n, allc, code, offsets, idx)
    assert rowwise_rel_err(zeta, g["zeta"]) <= 1e-13
    D = zeta.shape[1]
    for a in g["keep_atoms"]:
        o = 3 * D * int(offsets[a] + a)
        blk = packed[o:o + 3 * D * (counts[a] + 1)].reshape(D, counts[a] + 1, 3)
        assert rowwise_rel_err(blk, g[f"dz_{a}"]) <= 1e-13
    # digest over ALL blocks
    blocks = [packed[3 * D * int(offsets[a] + a):3 * D * int(offsets[a + 1] + a + 1)]
              .reshape(D, counts[a] + 1, 3) for a in range(n)]
    lists = [idx[offsets[a]:offsets[a + 1]] for a in range(n)]
    F = orc.forces_sparse(g["w"], blocks, lists, image, n).reshape(n, 3)
    assert np.max(np.abs(F - g["forces_w"])) <= 1e-12 * np.max(np.abs(g["forces_w"]))


def test_port_vs_recorded_reference_mos2(port):
    """Triclinic 288-atom MoS2 cell of the reference's LJ test: paddings + neighbor list."""
    g = load_golden("ref_run_mos2_neigh.npz")
    pc, ps, pi = port.create_paddings(float(g["infl_dist"]), g["cell"], g["pbc"], g["coords"], g["z"])
    assert np.array_equal(pc, g["pad_coords"]) and np.array_equal(pi, g["pad_image"])
    allc = np.concatenate([g["coords"], pc])
    need = np.zeros(len(allc), np.int32)
    need[:288] = 1
    counts, offsets, idx = port.build_neighbors(allc, float(g["infl_dist"]), need)
    assert np.array_equal(idx, g["nl_indices"]) and np.array_equal(counts, g["nl_counts"])


@pytest.mark.skipif(not orc.have_reference(), reason="oracle/_ref not built")
def test_port_vs_live_reference_random():
    """Live differential test against the compiled reference on fresh random inputs,
    including padding_need_neigh=True style lists and a collision error."""
    port, ref = orc.Oracle("port"), orc.Oracle("reference")
    rng = np.random.default_rng(11)
    for trial in range(4):
        cell = np.eye(3) * rng.uniform(5, 9) + rng.normal(0, 0.4, (3, 3))
        n = int(rng.integers(5, 40))
        coords = rng.random((n, 3)) @ cell + rng.normal(0, 2.0, 3)
        z = rng.choice([14, 6], n).astype(np.int32)
        pbc = rng.integers(0, 2, 3).astype(np.int32)
        a = port.create_paddings(3.9, cell, pbc, coords, z)
        b = ref.create_paddings(3.9, cell, pbc, coords, z)
        for x, y in zip(a, b):
            assert np.array_equal(x, y)
        allc = np.concatenate([coords, a[0]]) if len(a[0]) else coords
        need = np.ones(len(allc), np.int32)
        try:
            ra = port.build_neighbors(allc, 3.9, need)
        except RuntimeError:
            with pytest.raises(RuntimeError):
                ref.build_neighbors(allc, 3.9, need)
            continue
        rb = ref.build_neighbors(allc, 3.9, need)
        for x, y in zip(ra, rb):
            assert np.array_equal(x, y)
    # collision -> error from both
    bad = np.array([[0.0, 0, 0], [1e-6, 0, 0], [1.0, 1, 1]])
    for o in (port, ref):
        with pytest.raises(RuntimeError):
            o.build_neighbors(bad, 3.0, np.ones(3, np.int32))
    # singular cell -> error from both
    for o in (port, ref):
        with pytest.raises(RuntimeError):
            o.create_paddings(3.0, np.zeros((3, 3)), [1, 1, 1], bad, np.ones(3, np.int32))


def _species_codes_as_parsed(path):
    """Species -> code in first-appearance order of the cutoff lines, as Descriptor::read_parameter_file
    assigns them (sym_fn.cpp:150-200)."""
    codes = {}
    lines = [l.split("#")[0].split() for l in open(path)]
    lines = [l for l in lines if l]
    nsp = int(lines[1][0])
    for l in lines[2:2 + nsp * (nsp + 1) // 2]:
        for s in l[:2]:
            codes.setdefault(s, len(codes))
    return codes


@pytest.mark.skipif(not orc.have_reference(), reason="oracle/_ref not built (no /root/reference)")
def test_write_kim_params_read_back_by_the_reference_parser(tmp_path):
    """`descriptor.params` written by SymmetryFunction.write_kim_params (sym_fn.py:285-384) is parsed by
    the reference's own Descriptor::read_parameter_file (sym_fn.cpp:108-387), and generate_one_atom on the
    parsed object reproduces the recorded reference run (two species, three cutoffs, set51) — i.e. the file
    carries the cutoffs, species order, every hyper-parameter and the mean / stdev without loss."""
    from kliff_b200.legacy.descriptors import SymmetryFunction
    g = load_golden("ref_run_sic64_set51.npz")
    desc = SymmetryFunction({"Si-Si": 5.0, "Si-C": 4.5, "C-C": 4.0}, "cos", "set51", normalize=True,
                            dtype=np.float64)
    rng = np.random.default_rng(3)
    desc.mean, desc.stdev = rng.normal(size=51), rng.uniform(0.5, 2.0, size=51)
    desc.write_kim_params(str(tmp_path))
    path = tmp_path / "descriptor.params"
    codes = _species_codes_as_parsed(path)
    assert set(codes) == {"Si", "C"}
    n, allc, image = _full(g)
    sym = np.where(g["z"][image] == 14, "Si", "C")
    code = np.array([codes[s] for s in sym], np.int32)
    ref = orc.Oracle("reference")
    atoms = [int(a) for a in g["keep_atoms"]]
    zeta, blocks, mean, stdev = ref.symfun_atoms_from_file(str(path), atoms, allc, code, g["nl_offsets"],
                                                           g["nl_indices"])
    assert zeta.shape == (len(atoms), 51)
    # 16 significant digits in the file: parameters round-trip to <= 1 ulp, outputs to ~1e-13
    assert rowwise_rel_err(zeta, g["zeta"][atoms]) <= 1e-12
    for a, blk in zip(atoms, blocks):
        assert rowwise_rel_err(blk, g[f"dz_{a}"]) <= 1e-12
    assert np.allclose(mean, desc.mean, rtol=1e-15, atol=0) and np.allclose(stdev, desc.stdev, rtol=1e-15, atol=0)
    # fp32 descriptors write 8 significant digits (sym_fn.py:296-299): still parsed, to that precision
    desc32 = SymmetryFunction({"Si-Si": 5.0, "Si-C": 4.5, "C-C": 4.0}, "cos", "set51", normalize=False)
    desc32.write_kim_params(str(tmp_path), fname="d32.params")
    z32, _, m32, _ = ref.symfun_atoms_from_file(str(tmp_path / "d32.params"), atoms, allc, code,
                                                g["nl_offsets"], g["nl_indices"])
    assert m32 is None
    assert rowwise_rel_err(z32, g["zeta"][atoms]) <= 1e-5
