"""New-API `Descriptor` (kliff/transforms/configuration_transforms/descriptors.py:69-356) for
`descriptor="SymmetryFunctions"`.  The reference delegates the arithmetic to the un-vendored
`libdescriptor` (Enzyme-AD VJP); its contract is pinned by the reference's own differential test
against the legacy C++ (tests/transformations/test_descriptors.py:14-72): `forward(conf)` equals
the legacy zeta, and `backward(conf, dE_dZeta)` equals tensordot(dE_dZeta, dzetadr) folded onto
the contributing atoms.  Here both are the same sm_100a kernels as the legacy path: forward =
descriptor kernel, backward = the force-scatter contraction (sign flipped: it returns dE/dr)."""
from collections import OrderedDict
from pathlib import Path

import numpy as np
import torch

from ...legacy.descriptors.symmetry_function import SymmetryFunction
from ...neighbor import NeighborList
from ...utils import default_device
from .default_hyperparams import symmetry_functions_set30, symmetry_functions_set51


class DescriptorsError(Exception):
    def __init__(self, msg):
        super().__init__(msg)
        self.msg = msg


_AVAILABLE = ["SymmetryFunctions"]


def show_available_descriptors():
    print("-" * 80)
    print("Descriptors below are currently available, select them by `descriptor: str` attribute:")
    print("-" * 80)
    for key in _AVAILABLE:
        print(key)


class Descriptor:
    def __init__(self, cutoff, species, descriptor, hyperparameters, cutoff_function="cos",
                 copy_to_config=False):
        if descriptor not in _AVAILABLE:
            raise DescriptorsError(
                f"Descriptor {descriptor} is outside the accelerated path; available: {_AVAILABLE}")
        if cutoff_function != "cos":
            raise DescriptorsError("only the `cos` cutoff function is supported")
        self.copy_to_config = copy_to_config
        self.cutoff = cutoff
        self.species = list(species)
        self.descriptor_name = descriptor
        self.cutoff_function = cutoff_function
        self.hyperparameters = self.get_default_hyperparams(hyperparameters)
        cut_dists = {}
        for a in self.species:
            for b in self.species:
                if f"{b}-{a}" not in cut_dists:
                    cut_dists[f"{a}-{b}"] = float(cutoff)
        self._sf = SymmetryFunction(cut_dists, "cos", OrderedDict(self.hyperparameters),
                                    normalize=False, dtype=np.float64)
        self.width = len(self._sf)

    @staticmethod
    def get_default_hyperparams(hyperparameters):
        if isinstance(hyperparameters, str):
            if hyperparameters == "set51":
                return symmetry_functions_set51()
            if hyperparameters == "set30":
                return symmetry_functions_set30()
            raise DescriptorsError("Hyperparameter set not found")
        if isinstance(hyperparameters, (dict, OrderedDict)):
            return hyperparameters
        raise DescriptorsError("Hyperparameters must be either a string or an OrderedDict")

    def _map_species_to_int(self, species):
        return [self.species.index(s) for s in species]

    def forward(self, configuration):
        """(n_atoms, width) descriptors."""
        packed = self._sf.transform_packed([configuration], grad=False)
        return packed.zeta.cpu().numpy()

    def backward(self, configuration, dE_dZeta):
        """Vector-Jacobian product: (n_atoms, 3) = sum_{i,d} dE_dZeta[i,d] * d zeta[i,d] / d r."""
        packed = self._sf.transform_packed([configuration], grad=True)
        n = configuration.get_num_atoms()
        w = torch.as_tensor(np.ascontiguousarray(dE_dZeta, dtype=np.float64), device=packed.device)
        return (-packed.forces(w, 0, n)).reshape(n, 3).cpu().numpy()

    def __call__(self, configuration, return_extended_state=False):
        nl_ctx = NeighborList(configuration, self.cutoff, device=default_device())
        n_atoms = configuration.get_num_atoms()
        descriptors = self.forward(configuration)
        if return_extended_state:
            num_neigh, neigh_list = nl_ctx.get_numneigh_and_neighlist_1D()
            output = {
                "n_atoms": n_atoms,
                "species": np.array(self._map_species_to_int(nl_ctx.species), np.intc),
                "neigh_list": neigh_list,
                "num_neigh": num_neigh,
                "image": nl_ctx.get_image(),
                "coords": nl_ctx.get_coords(),
                "descriptor": descriptors,
                "index": configuration.metadata.get("index", -1),
                "weight": configuration.weight.to_dict(),
            }
        else:
            output = descriptors
        if self.copy_to_config:
            configuration.fingerprint = output
        return output

    # ------------------------------------------------------------------ ConfigurationTransform interface
    # (kliff/transforms/configuration_transforms/configuration_transform.py:58-96)
    def transform(self, configuration):
        return self(configuration)

    def inverse(self, *args, **kargs):
        """The reference builds a NotImplementedError without raising it (configuration_transform.py:58-67):
        descriptors have no inverse mapping; `backward` is the vector-Jacobian product."""
        return None

    def collate_fn(self, config_list):
        return [self(conf) for conf in config_list]

    # ------------------------------------------------------------------ parameter files (descriptors.py:358-512)
    _COLUMNS = {"g2": ("eta", "Rs"), "g3": ("kappa",), "g4": ("zeta", "lambda", "eta"),
                "g5": ("zeta", "lambda", "eta")}

    def save_descriptor_state(self, path, fname="descriptor.dat"):
        """The text file libdescriptor re-initialises a symmetry-function descriptor from: cutoff type, the
        species-pair cutoff table, every family with its rows of parameters, and the (identity) centring
        block — same layout, line for line, as the reference writes (checked against a file written by the
        reference's own function, tests/golden/descriptor_state_*.dat)."""
        rule = "#" + "=" * 80
        out = [rule, "# Descriptor parameters file generated by KLIFF.", rule, "",
               f"{self.cutoff_function}  # cutoff type", "",
               f"{len(self.species)}  # number of species", "",
               "# species 1    species 2    cutoff"]
        out += [f"{a}  {b}  {self.cutoff}" for a in self.species for b in self.species]
        out += ["", rule, "# symmetry functions", rule, "",
                f"{len(self.hyperparameters)}  # number of symmetry functions types", "",
                "# sym_function    rows    cols"]
        for name, rows in self.hyperparameters.items():
            if name == "g1":
                out += ["g1", ""]
                continue
            out.append(f"{name}    {len(rows)}    {len(rows[0])}")
            cols = self._COLUMNS.get(name)
            if cols is not None:
                out += ["    ".join(str(row[c]) for c in cols) + "    # " + "  ".join(cols) for row in rows]
                out.append("")
        out += [rule, "# Preprocessing data to center and normalize", rule,
                "center_and_normalize  True", "",
                f"{self.width}   # descriptor size", "# mean", "0.0 ", "", "# standard deviation", "1.0 ", ""]
        Path(path, fname).write_text("\n".join(out) + "\n")

    def export_kim_model(self, path, model):
        """`kim_model.param` of the reference's TorchML driver export (descriptors.py:484-512)."""
        blocks = [("Number of elements", [str(len(self.species)), " ".join(self.species)]),
                  ("Preprocessing kind", ["Descriptor"]),
                  ("Cutoff distance", [str(self.cutoff)]),
                  ("Model", [str(model)]),
                  ("Returns Forces", ["False"]),
                  ("Number of inputs", ["1"])]
        text = "".join(f"# {title}\n" + "".join(v + "\n" for v in values) + "\n" for title, values in blocks)
        text += f"# Any descriptors?\n{self.descriptor_name}\n"
        Path(path, "kim_model.param").write_text(text)
