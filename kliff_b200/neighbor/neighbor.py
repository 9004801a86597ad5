"""`NeighborList` with the interface of kliff/neighbor/neighbor.py:10-257, built by the sm_100a
padding and cell-list kernels (kliff_b200/csrc/pad.cu, neighbor.cu) instead of the `neighlist`
C++ extension.

Differences a caller can observe: each atom's neighbors come back sorted ascending (the
reference returns its internal cell-scan order; as sets they are identical, and padding atoms
keep the reference's numbering), and the host-side numpy attributes are materialised lazily —
the device-resident tensors (`.device_data`) are what the descriptor kernels consume.
"""
import numpy as np
import torch

from .. import ops
from .._lib import KB_ERR_REFERENCE, KliffB200Error
from ..atomic_data import atomic_number, atomic_species
from ..utils import default_device, map_species


class NeighborListError(Exception):
    def __init__(self, msg):
        super().__init__(msg)
        self.msg = msg


class NeighborList:
    def __init__(self, conf, infl_dist, padding_need_neigh=False, device=None):
        self.conf = conf
        self.infl_dist = infl_dist
        self.padding_need_neigh = padding_need_neigh
        self._device = device if device is not None else default_device()
        self._host = {}
        self.device_data = None
        self.create_neigh()

    # ------------------------------------------------------------------ build (neighbor.py:68-111)
    def create_neigh(self):
        dev = self._device
        coords_cb = np.ascontiguousarray(self.conf.coords, dtype=np.double).reshape(-1, 3)
        cell = np.asarray(self.conf.cell, dtype=np.double)
        pbc = np.asarray(self.conf.PBC, dtype=np.intc)
        # species play no role in the neighbor build: paddings inherit the species of the atom they
        # are an image of (neighbor_list.cpp:407-408), looked up lazily through `pad_image`
        d_coords = torch.as_tensor(coords_cb, device=dev)
        try:
            pc, _, pi = ops.create_paddings(self.infl_dist, cell, pbc, d_coords, None)
        except KliffB200Error as e:
            if e.code == KB_ERR_REFERENCE:
                raise NeighborListError("Calling `neighlist.create_paddings` failed.")
            raise
        num_cb = coords_cb.shape[0]
        allc = torch.cat([d_coords, pc]) if pc.shape[0] else d_coords
        image = torch.cat([torch.arange(num_cb, dtype=torch.int32, device=dev), pi])
        need = None
        if self.padding_need_neigh:
            need = torch.ones(allc.shape[0], dtype=torch.int32, device=dev)
        try:
            counts, offsets, indices, maxn = ops.build_neighbors(allc, num_cb, self.infl_dist, need=need)
        except KliffB200Error as e:
            if e.code == KB_ERR_REFERENCE:
                raise NeighborListError("Calling `neighlist.build` failed.")
            raise
        self.device_data = dict(coords=allc, image=image, counts=counts,
                                offsets=offsets, indices=indices, maxn=maxn, n_contrib=num_cb,
                                pad_coords=pc, pad_image=pi)
        self._host = {}

    # ------------------------------------------------------------------ lazy host mirrors
    def _h(self, key):
        if key not in self._host:
            self._host[key] = self.device_data[key].cpu().numpy()
        return self._host[key]

    @property
    def coords(self):
        return self._h("coords")

    @property
    def image(self):
        return self._h("image").astype(np.intc, copy=False)

    @property
    def species(self):
        if "species" not in self._host:
            self._host["species"] = np.concatenate((self.conf.species, self.padding_species))
        return self._host["species"]

    @property
    def padding_coords(self):
        return self._h("pad_coords")

    @property
    def padding_species(self):
        sp = np.asarray(self.conf.species)
        for name in np.unique(sp):
            if str(name) not in atomic_number:
                raise NeighborListError(f"unknown chemical species {name}")
        return [str(x) for x in sp[self._h("pad_image")]]

    @property
    def padding_image(self):
        return self._h("pad_image").astype(np.intc, copy=False)

    # ------------------------------------------------------------------ accessors
    def get_neigh(self, index):
        """(indices, coords, species) of the neighbors of atom `index` (neighbor.py:113-139)."""
        off = self._h("offsets")
        if index < 0 or index >= len(off) - 1:
            raise NeighborListError("Calling `neighlist.get_neigh` failed.")
        neigh_indices = self._h("indices")[off[index]:off[index + 1]].astype(np.intc, copy=True)
        return neigh_indices, self.coords[neigh_indices], self.species[neigh_indices]

    def get_numneigh_and_neighlist_1D(self, request_padding=False):
        """neighbor.py:141-187."""
        if request_padding:
            if not self.padding_need_neigh:
                raise NeighborListError(
                    "Request to get neighbors of padding atoms, but "
                    '"padding_need_neigh" is set to "False" at initialization.'
                )
            n = len(self._h("counts"))
        else:
            n = self.conf.get_num_atoms()
        off = self._h("offsets")
        numneigh = self._h("counts")[:n].astype(np.intc, copy=True)
        neighlist = self._h("indices")[: off[n]].astype(np.intc, copy=True)
        return numneigh, neighlist

    def get_coords(self):
        return self.coords.copy()

    def get_species(self):
        return self.species[:]

    def get_species_code(self, mapping):
        return np.asarray([mapping[s] for s in self.species], dtype=np.intc)

    def get_image(self):
        return self.image.copy()

    def get_padding_coords(self):
        return self.padding_coords.copy()

    def get_padding_species(self):
        return self.padding_species[:]

    def get_padding_species_code(self, mapping):
        return np.asarray([mapping[s] for s in self.padding_species], dtype=np.intc)

    def get_padding_image(self):
        return self.padding_image.copy()


def assemble_forces(forces, n, padding_image):
    """Fold forces on padding atoms back onto their owners (neighbor.py:260-296)."""
    forces = np.asarray(forces)
    total = np.array(forces[:n])
    padding_image = np.asarray(padding_image)
    if padding_image.size:
        np.add.at(total, padding_image, forces[n:])
    return total


def assemble_stress(coords, forces, volume):
    """Virial stress, Voigt xx,yy,zz,yz,xz,xy (neighbor.py:301-328)."""
    coords, forces = np.asarray(coords), np.asarray(forces)
    pairs = ((0, 0), (1, 1), (2, 2), (1, 2), (2, 0), (0, 1))
    return np.array([-np.sum(coords[:, a] * forces[:, b]) / volume for a, b in pairs])
