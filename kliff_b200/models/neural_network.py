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
