"""
Minimal input containers for the hot path.

The reference's kliff.dataset hard-imports `ase` and `monty` (kliff/dataset/configuration.py:20-27,
dataset.py:14), neither of which exists in this image or on the GPU box, so the drop-in ships a
small duck-type-compatible subset: exactly the attributes the hot path reads (SURVEY.md §2 row
14, App. D).  Anything else of kliff.dataset is out of scope.
"""
from .configuration import Configuration, ConfigurationError, Dataset, DatasetError
from .extxyz import read_extxyz, write_extxyz
from .weight import MagnitudeInverseWeight, Weight, WeightError

__all__ = [
    "Configuration",
    "ConfigurationError",
    "Dataset",
    "DatasetError",
    "Weight",
    "read_extxyz",
    "write_extxyz",
]
