from .default_hyperparams import symmetry_functions_set30, symmetry_functions_set51
from .descriptors import Descriptor, DescriptorsError, show_available_descriptors

__all__ = ["Descriptor", "DescriptorsError", "show_available_descriptors",
           "symmetry_functions_set30", "symmetry_functions_set51"]
