"""`ModelTorch` / `NeuralNetwork` (kliff/models/model_torch.py:13-163, neural_network.py:13-81).
The tiny MLP stays on torch/cuBLAS (north_star); this is the container the calculator expects:
`.descriptor`, `.device`, `.dtype`, `add_layers`, save/load + save metadata."""
from pathlib import Path

import numpy as np
import torch
import torch.nn as nn

from ..utils import create_directory, seed_all, to_path


class ModelTorchError(Exception):
    def __init__(self, msg):
        super().__init__(msg)
        self.msg = msg


class NeuralNetworkError(Exception):
    def __init__(self, msg):
        super().__init__(msg)
        self.msg = msg


class ModelTorch(nn.Module):
    def __init__(self, descriptor, seed=35):
        super().__init__()
        self._descriptor = descriptor
        self.seed = seed
        seed_all(seed)  # model_torch.py:26-31: fixes the layer initialisation of the tutorials
        dtype = self._descriptor.get_dtype()
        if dtype == np.float32:
            self._dtype = torch.float32
        elif dtype == np.float64:
            self._dtype = torch.float64
        else:
            raise ModelTorchError(f"Not support dtype {dtype}.")
        self._save_prefix = Path.cwd() / "kliff_saved_model"
        self._save_start = 1
        self._save_frequency = 10

    def forward(self, x):
        raise NotImplementedError("`forward` not implemented.")

    def fit(self, path):
        raise ModelTorchError(
            "Analytic fitting not supported for this model. Minimize a loss function "
            "to train the model instead.")

    def save(self, filename):
        state = {"model_state_dict": self.state_dict(),
                 "descriptor_state_dict": self.descriptor.state_dict()}
        filename = to_path(filename)
        create_directory(filename)
        torch.save(state, str(filename))

    def load(self, filename, mode="train"):
        state = torch.load(str(to_path(filename)), weights_only=False)
        self.load_state_dict(state["model_state_dict"])
        if mode == "train":
            self.train()
        elif mode == "eval":
            self.eval()
        else:
            raise ModelTorchError(f"Unrecognized mode `{mode}`.")
        self.descriptor.load_state_dict(state["descriptor_state_dict"])

    def set_save_metadata(self, prefix, start, frequency=1):
        self._save_prefix = to_path(prefix)
        self._save_start = int(start)
        self._save_frequency = int(frequency)

    descriptor = property(lambda self: self._descriptor)
    device = property(lambda self: next(self.parameters()).device)
    dtype = property(lambda self: self._dtype)
    save_prefix = property(lambda self: self._save_prefix)
    save_start = property(lambda self: self._save_start)
    save_frequency = property(lambda self: self._save_frequency)


class NeuralNetwork(ModelTorch):
    def __init__(self, descriptor, seed=35):
        super().__init__(descriptor, seed)
        self.layers = None

    def add_layers(self, *layers):
        if self.layers is not None:
            raise NeuralNetworkError("`add_layers()` called multiple times. It should be called only once.")
        self.layers = []
        for la in layers:
            self.layers.append(la)
            setattr(self, "layer_{}".format(len(self.layers)), la)
        first, last = self.layers[0], self.layers[-1]
        if first.in_features != self.descriptor.get_size():
            raise NeuralNetworkError(
                f"Expect `in_features` of first layer ({first.in_features}) be equal "
                f"to descriptor size ({self.descriptor.get_size()}).")
        if last.out_features != 1:
            raise NeuralNetworkError("`out_features` of last layer should be 1.")
        self.type(self.dtype)

    def forward(self, x):
        for layer in self.layers:
            x = layer(x)
        return x

    # ------------------------------------------------------------------ KIM / DUNN export
    # File formats of neural_network.py:83-330 (read by the DUNN model driver).  Layers must come
    # as [Dropout]? (Linear Activation [Dropout]?)* Linear, as there.
    _ACTIVATIONS = {"Sigmoid": "sigmoid", "Tanh": "tanh", "ReLU": "relu", "ELU": "elu"}

    def _kim_groups(self):
        """-> (linears, activation name, keep probabilities per Linear input)."""
        if not self.layers:
            raise NeuralNetworkError("no layers to write; call `add_layers()` first")
        linears, acts, keep, pending_drop, prev = [], [], [], 0.0, None
        for i, layer in enumerate(self.layers):
            name = layer.__class__.__name__
            if name == "Linear":
                linears.append(layer)
                keep.append(1.0 - pending_drop)
                pending_drop = 0.0
            elif name in self._ACTIVATIONS:
                if prev != "Linear":
                    raise NeuralNetworkError(
                        f"Cannot convert to KIM model. a `{name}` layer must follow a \"Linear\" layer.")
                acts.append(self._ACTIVATIONS[name])
            elif name.startswith("Dropout"):
                if i != 0 and prev not in self._ACTIVATIONS:
                    raise NeuralNetworkError(
                        f"Cannot convert to KIM model. a `{name}` layer must follow an activation layer.")
                pending_drop = float(layer.p)
            else:
                raise NeuralNetworkError(f"Layer `{name}` not supported by KIM model. Cannot proceed to write.")
            prev = name
        if prev != "Linear" or len(acts) != len(linears) - 1:
            raise NeuralNetworkError("Cannot convert to KIM model: the last layer must be `Linear` and every "
                                     "other `Linear` layer needs an activation.")
        if len(set(acts)) > 1:
            raise NeuralNetworkError("KIM model supports a single activation function type")
        return linears, (acts[0] if acts else "tanh"), keep

    def write_kim_model(self, path=None, driver_name="DUNN__MD_292677547454_000", dropout_ensemble_size=None):
        if path is None:
            path = Path.cwd() / "NeuralNetwork_KLIFF__MO_000000111111_000"
        path = Path(path).expanduser().resolve()
        path.mkdir(parents=True, exist_ok=True)
        files = ["descriptor.params", "NN.params", "dropout_binary.params"]
        linears, activation, keep = self._kim_groups()
        bar = "#" + "=" * 80 + "\n"
        num = "{:23.15e}" if self.dtype == torch.float64 else "{:15.7e}"

        with open(path / "CMakeLists.txt", "w") as f:
            f.write("#\n# Contributors:\n#    KLIFF (https://kliff.readthedocs.io)\n#\n\n"
                    "cmake_minimum_required(VERSION 3.4)\n\n"
                    "list(APPEND CMAKE_PREFIX_PATH $ENV{KIM_API_CMAKE_PREFIX_DIR})\n"
                    "find_package(KIM-API 2.0 REQUIRED CONFIG)\n"
                    "if(NOT TARGET kim-api)\n  enable_testing()\n"
                    '  project("${KIM_API_PROJECT_NAME}" VERSION "${KIM_API_VERSION}"\n'
                    "    LANGUAGES CXX C Fortran)\nendif()\n\n"
                    "add_kim_api_model_library(\n"
                    f'  NAME            "{path.name}"\n  DRIVER_NAME     "{driver_name}"\n'
                    "  PARAMETER_FILES" + "".join(f' "{s}"' for s in files) + "\n  )\n")

        with open(path / files[1], "w") as f:
            f.write(bar + "# NN structure and parameters file generated by KLIFF\n# \n"
                    '# Note that the NN assumes each row of the input "X" is an \n'
                    "# observation, i.e. the layer is implemented as\n# Y = activation(XW + b).\n"
                    '# You need to transpose your weight matrix if each column of "X" is \n'
                    "# an observation.\n" + bar + "\n")
            f.write(f"{len(linears)}    # number of layers (excluding input layer,including output layer)\n")
            f.write("".join(f"{la.out_features}  " for la in linears) + "  # size of each layer (last must be 1)\n")
            f.write(f"{activation}    # activation function\n")
            f.write("".join("{:.15g}  ".format(k) for k in keep) + "  # keep probability of input for each layer\n\n")
            for i, la in enumerate(linears):
                w = la.weight.detach().cpu().t()  # torch stores W^T (y = x W^T + b)
                b = la.bias.detach().cpu()
                last = i == len(linears) - 1
                f.write("# weight of output layer, shape({}, {})\n".format(*w.shape) if last else
                        "# weight of hidden layer {},  shape({}, {})\n".format(i + 1, *w.shape))
                for row in w:
                    f.write("".join(num.format(float(x)) for x in row) + "\n")
                f.write("# bias of output layer, shape({}, )\n".format(w.shape[1]) if last else
                        "# bias of hidden layer {}, shape({}, )\n".format(i + 1, w.shape[1]))
                f.write("".join(num.format(float(x)) for x in b) + "\n\n")

        self.descriptor.write_kim_params(path, files[0])

        drop = [1.0 - k for k in keep]
        size = 0 if all(d < 1e-10 for d in drop) else (100 if dropout_ensemble_size is None else dropout_ensemble_size)
        units = [self.descriptor.get_size()] + [la.out_features for la in linears]
        with open(path / files[2], "w") as f:
            f.write(bar + "# Dropout binary parameters file generated by KLIFF.\n#\n"
                    '# Note, "ensemble size = 0", means that no dropout needs to be\n'
                    "# applied at all." + bar + "\n")
            f.write(f"{size}  # ensemble size\n")
            for rep in range(size):
                f.write(bar + f"# instance {rep}\n")
                for i, k in enumerate(keep):
                    mask = np.clip(np.floor(np.random.uniform(k, k + 1, units[i])).astype(np.intc), 0, 1)
                    f.write(f"# layer {i}\n" + "".join(f"{d} " for d in mask) + "\n")
        return path
