from .descriptor import Descriptor, DescriptorError, load_fingerprints
from .hyperparams import get_set30, get_set51
from .symmetry_function import SymmetryFunction, SymmetryFunctionError

__all__ = ["Descriptor", "DescriptorError", "SymmetryFunction", "SymmetryFunctionError",
           "load_fingerprints", "get_set51", "get_set30"]
