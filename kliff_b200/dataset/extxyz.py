"""Extended-xyz reader/writer for the format of kliff/dataset/extxyz.py:9-130:
line 1 = atom count; line 2 = Lattice="9 floats" PBC="1 1 1" [Energy=..] [Stress="6 floats"]
Properties=...; then `symbol x y z [fx fy fz]` per atom."""
import re

import numpy as np


class InputError(Exception):
    pass


_KV = re.compile(r'(\w+)\s*=\s*(?:"([^"]*)"|(\S+))')


def read_extxyz(filename):
    with open(filename, "r") as fh:
        lines = fh.readlines()
    try:
        natoms = int(lines[0].split()[0])
    except (ValueError, IndexError) as e:
        raise InputError(f"{e}.\nCorrupted data at line 1 of file {filename}.")
    # the regex yields the quoted value (group 2) or the bare token (group 3)
    header = {}
    for m in _KV.finditer(lines[1].replace("'", '"')):
        header[m.group(1).lower()] = m.group(2) if m.group(2) is not None else m.group(3)
    if "lattice" not in header or "pbc" not in header:
        raise InputError(f"Lattice/PBC missing at line 2 of file {filename}.")
    cell = np.array([float(x) for x in header["lattice"].split()], dtype=np.float64)
    if cell.size != 9:
        raise InputError(f"Lattice needs 9 numbers at line 2 of file {filename}.")
    cell = cell.reshape(3, 3)
    pbc = []
    for tok in header["pbc"].split():
        if tok.lower() in ("t", "true"):
            pbc.append(True)
        elif tok.lower() in ("f", "false"):
            pbc.append(False)
        else:
            try:
                pbc.append(bool(int(tok)))
            except ValueError:
                raise InputError(f"Invalid PBC value {tok} at line 2 of file {filename}.")
    energy = float(header["energy"].split()[0]) if "energy" in header else None
    stress = [float(x) for x in header["stress"].split()] if "stress" in header else None

    body = [ln.split() for ln in lines[2:] if ln.strip()]
    if len(body) != natoms:
        raise InputError(
            f"Corrupted data file {filename}. Number of atoms is {natoms}, "
            f"whereas number of data lines is {len(body)}."
        )
    width = len(body[0]) if body else 4
    if width not in (4, 7):
        raise InputError(f"Corrupted data at line 3 of file {filename}.")
    species, coords, forces = [], [], []
    for n, tok in enumerate(body):
        if len(tok) != width:
            raise InputError(f'Corrupted data at line {n + 3} of file "{filename}".')
        species.append(tok[0].lower().capitalize())
        try:
            coords.append([float(v) for v in tok[1:4]])
            if width == 7:
                forces.append([float(v) for v in tok[4:7]])
        except ValueError as e:
            raise InputError(f"{e}.\nCorrupted data at line {n + 3} of file {filename}.")
    coords = np.asarray(coords, dtype=np.float64).reshape(-1, 3)
    forces = np.asarray(forces, dtype=np.float64).reshape(-1, 3) if width == 7 else None
    return cell, species, coords, pbc, energy, forces, stress


def write_extxyz(filename, cell, species, coords, PBC, energy=None, forces=None, stress=None):
    cell = np.asarray(cell, dtype=np.float64).reshape(9)
    with open(filename, "w") as fh:
        fh.write(f"{len(species)}\n")
        fh.write('Lattice="' + " ".join(f"{v:.15g}" for v in cell) + '" ')
        fh.write('PBC="' + " ".join(str(int(bool(p))) for p in PBC) + '" ')
        if energy is not None:
            fh.write(f"Energy={energy:.15g} ")
        if stress is not None:
            fh.write('Stress="' + " ".join(f"{v:.15g}" for v in stress) + '" ')
        fh.write("Properties=species:S:1:pos:R:3" + (":for:R:3" if forces is not None else "") + "\n")
        for a, s in enumerate(species):
            row = f"{s:2s} " + " ".join(f"{v:23.15e}" for v in coords[a])
            if forces is not None:
                row += " " + " ".join(f"{v:23.15e}" for v in forces[a])
            fh.write(row + "\n")
