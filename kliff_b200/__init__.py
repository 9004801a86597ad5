"""kliff_b200 — B200-native (sm_100a) drop-in for the descriptor / neural-network hot path of
openkim/kliff: neighbor list -> Behler-Parrinello symmetry functions + dG/dr -> chain-rule
force assembly, behind KLIFF's own plugin API.

    from kliff_b200.dataset import Dataset
    from kliff_b200.dataset.weight import Weight
    from kliff_b200.legacy import nn
    from kliff_b200.legacy.calculators import CalculatorTorch
    from kliff_b200.legacy.descriptors import SymmetryFunction
    from kliff_b200.legacy.loss import Loss
    from kliff_b200.models import NeuralNetwork
    from kliff_b200.neighbor import NeighborList

mirrors the imports of the reference's README example (README.md:45-92).  All numerics on the
path run in hand-written CUDA kernels (kliff_b200/csrc) reached through the C-ABI of
include/kliff_b200.h; there is no CPU fallback.
"""
__version__ = "0.1.0"
