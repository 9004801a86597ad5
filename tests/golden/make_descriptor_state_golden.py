"""Writes tests/golden/descriptor_state_{set30,custom}.dat and kim_model_param.txt with the REFERENCE's own
`Descriptor.save_descriptor_state` / `export_kim_model`
(kliff/transforms/configuration_transforms/descriptors.py:358-512), run on plain attribute holders — the
reference class itself cannot be instantiated here (un-vendored libdescriptor), but these two writers only
read attributes.  `lds.AvailableDescriptors(k)` is stubbed by an object that compares equal by k.
Run in the build container (needs /root/repo/baseline/_ref and baseline/stubs); the outputs are committed."""
import os
import sys
import types
from collections import OrderedDict

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "baseline", "stubs"))
sys.path.insert(0, os.path.join(ROOT, "baseline", "_ref", "kliff_src"))

import kliff.transforms.configuration_transforms.descriptors as ref  # noqa: E402
from kliff.transforms.configuration_transforms.default_hyperparams import symmetry_functions_set30  # noqa: E402


class _Kind(int):
    pass


ref.lds = types.SimpleNamespace(AvailableDescriptors=_Kind)


def holder(species, cutoff, hyper):
    h = types.SimpleNamespace()
    h.cutoff_function, h.cutoff, h.species = "cos", cutoff, species
    h.descriptor_kind, h.descriptor_name = _Kind(0), "SymmetryFunctions"
    h.hyperparameters = hyper
    h.width = sum(1 if k == "g1" else len(v) for k, v in hyper.items())
    return h


custom = OrderedDict([
    ("g1", None),
    ("g2", [{"eta": 0.0035710676725828126, "Rs": 0.0}, {"eta": 0.03571067672582813, "Rs": 1.5}]),
    ("g3", [{"kappa": 0.3}, {"kappa": 1.25}]),
    ("g4", [{"zeta": 1, "lambda": -1, "eta": 0.00035710676725828126}, {"zeta": 2.5, "lambda": 0.5, "eta": 0.02}]),
    ("g5", [{"zeta": 4, "lambda": 1, "eta": 0.001}]),
])
for tag, h in (("set30", holder(["Si"], 5.0, symmetry_functions_set30())),
               ("custom", holder(["Si", "C"], 4.5, custom))):
    ref.Descriptor.save_descriptor_state(h, HERE, f"descriptor_state_{tag}.dat")
h = holder(["Si", "C"], 4.5, custom)
ref.Descriptor.export_kim_model(h, HERE, "model.pt")
os.replace(os.path.join(HERE, "kim_model.param"), os.path.join(HERE, "kim_model_param.txt"))
print("written")
