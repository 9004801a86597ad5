"""Pre-defined symmetry-function hyper-parameter sets.

Same data (and the same ordering, which fixes the descriptor index) as `get_set51` /
`get_set30` in kliff/legacy/descriptors/symmetry_function/sym_fn.py:399-532 — Artrith & Behler,
PRB 85, 045439 (2012) and Artrith, Hiller & Behler, pss(b) 250, 1191 (2013).  eta is tabulated
in bohr^-2 and converted with the reference's own constant 0.529177 (sym_fn.py:467-473,
524-530); all Rs = 0.
"""
from collections import OrderedDict

BOHR_TO_ANGSTROM = 0.529177


def _assemble(g2_eta, g4_blocks):
    scale = BOHR_TO_ANGSTROM**2
    params = OrderedDict()
    params["g2"] = [{"eta": eta / scale, "Rs": 0.0} for eta in g2_eta]
    params["g4"] = [
        {"zeta": zeta, "lambda": lam, "eta": eta / scale}
        for eta, pairs in g4_blocks
        for zeta, lam in pairs
    ]
    return params


def _pm(*zetas):
    return [(z, s) for z in zetas for s in (-1, 1)]


def _plus(*zetas):
    return [(z, 1) for z in zetas]


def get_set51():
    """8 radial G2 + 43 angular G4 = 51 (note: (zeta=16, lambda=-1, eta=0.08) is absent,
    sym_fn.py:462)."""
    g2 = [0.001, 0.01, 0.02, 0.035, 0.06, 0.1, 0.2, 0.4]
    g4 = [(eta, _pm(1, 2)) for eta in (0.0001, 0.003, 0.008)]
    g4 += [(eta, _pm(1, 2, 4, 16)) for eta in (0.015, 0.025, 0.045)]
    g4 += [(0.08, _pm(1, 2, 4) + [(16, 1)])]
    return _assemble(g2, g4)


def get_set30():
    """8 radial G2 + 22 angular G4 = 30."""
    g2 = [0.0009, 0.01, 0.02, 0.035, 0.06, 0.1, 0.2, 0.4]
    g4 = [(eta, _pm(1, 2)) for eta in (0.0001, 0.003)]
    g4 += [(0.008, _plus(1, 2))]
    g4 += [(eta, _plus(1, 2, 4, 16)) for eta in (0.015, 0.025, 0.045)]
    return _assemble(g2, g4)
