"""torch.nn re-export plus the per-descriptor-component Dropout of kliff/legacy/nn.py:6-45
(the tutorials build their MLPs from `kliff.legacy.nn`)."""
import torch
from torch.nn import *  # noqa: F401,F403


class Dropout(torch.nn.modules.dropout._DropoutNd):
    """Zeroes the same descriptor component for all atoms of the batch."""

    def forward(self, input):
        dim, shape = input.dim(), input.shape
        if dim == 2:
            shape4 = (1, *shape, 1)
        elif dim == 3:
            if shape[0] != 1:
                raise Exception("Shape[0] needs to be 1 for a 3D tensor.")
            shape4 = (*shape, 1)
        else:
            raise Exception("Input need to be 2D or 3D tensor, but got a {}D tensor.".format(dim))
        x = torch.transpose(torch.reshape(input, shape4), 1, 2)
        y = torch.nn.functional.dropout2d(x, self.p, self.training, self.inplace)
        return torch.reshape(torch.transpose(y, 1, 2), shape)
