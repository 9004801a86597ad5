"""`symmetry_functions_set51/set30` of kliff/transforms/configuration_transforms/
default_hyperparams.py:4-111: the same sets as the legacy `get_set51/get_set30`, already converted
to Angstrom^-2."""
from ...legacy.descriptors.hyperparams import get_set30, get_set51


def symmetry_functions_set51():
    return dict(get_set51())


def symmetry_functions_set30():
    return dict(get_set30())
