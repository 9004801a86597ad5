"""`Configuration` / `Dataset` subset (kliff/dataset/configuration.py:34-147, 350-520;
dataset.py:175-246, 1022): the hot path reads cell, species, coords, PBC (+ energy, forces,
stress, weight, identifier for the loss)."""
import os
from pathlib import Path

import numpy as np

from .extxyz import read_extxyz, write_extxyz
from .weight import Weight


class ConfigurationError(Exception):
    pass


class DatasetError(Exception):
    pass


class Configuration:
    def __init__(self, cell, species, coords, PBC, energy=None, forces=None, stress=None,
                 weight=None, identifier=None):
        self._cell = cell
        self._species = species
        self._coords = coords
        self._PBC = PBC
        self._energy = energy
        self._forces = forces
        self._stress = stress
        self._identifier = identifier
        self._path = None
        self._metadata = {}
        self._fingerprint = None
        self._weight = Weight() if weight is None else weight
        self._weight.compute_weight(self)

    @classmethod
    def from_file(cls, filename, weight=None, file_format="xyz"):
        if file_format != "xyz":
            raise ConfigurationError(f"Expect data file_format to be one of ['xyz'], got: {file_format}.")
        cell, species, coords, PBC, energy, forces, stress = read_extxyz(filename)
        self = cls(np.asarray(cell), [str(s) for s in species], np.asarray(coords),
                   [bool(p) for p in PBC], float(energy) if energy is not None else None,
                   np.asarray(forces) if forces is not None else None,
                   [float(s) for s in stress] if stress is not None else None,
                   weight, identifier=str(filename))
        self._path = Path(filename)
        return self

    def to_file(self, filename, file_format="xyz"):
        if file_format != "xyz":
            raise ConfigurationError(f"Expect data file_format to be one of ['xyz'], got: {file_format}.")
        write_extxyz(filename, self._cell, self._species, self._coords, self._PBC, self._energy,
                     self._forces, self._stress)

    cell = property(lambda self: self._cell)
    PBC = property(lambda self: self._PBC)
    species = property(lambda self: self._species)
    coords = property(lambda self: self._coords)
    path = property(lambda self: self._path)

    @property
    def energy(self):
        if self._energy is None:
            raise ConfigurationError("Configuration does not contain energy.")
        return self._energy

    @energy.setter
    def energy(self, value):
        self._energy = value

    @property
    def forces(self):
        if self._forces is None:
            raise ConfigurationError("Configuration does not contain forces.")
        return self._forces

    @forces.setter
    def forces(self, value):
        self._forces = value

    @property
    def stress(self):
        if self._stress is None:
            raise ConfigurationError("Configuration does not contain stress.")
        return self._stress

    @stress.setter
    def stress(self, value):
        self._stress = value

    @property
    def weight(self):
        return self._weight

    @weight.setter
    def weight(self, value):
        self._weight = value
        self._weight.compute_weight(self)

    @property
    def identifier(self):
        return self._identifier

    @identifier.setter
    def identifier(self, value):
        self._identifier = value

    @property
    def metadata(self):
        return self._metadata

    @property
    def fingerprint(self):
        """What a configuration transform left here with `copy_to_config=True` (configuration.py:455-469)."""
        return self._fingerprint

    @fingerprint.setter
    def fingerprint(self, value):
        self._fingerprint = value

    def get_num_atoms(self):
        return len(self._species)

    def get_num_atoms_by_species(self):
        out = {}
        for s in self._species:
            out[s] = out.get(s, 0) + 1
        return out

    def get_volume(self):
        c = np.asarray(self._cell, dtype=np.float64)
        return abs(float(np.dot(c[0], np.cross(c[1], c[2]))))


class Dataset:
    def __init__(self, configs=None):
        self.configs = list(configs) if configs is not None else []

    @classmethod
    def from_path(cls, path, weight=None, file_format="xyz"):
        inst = cls()
        inst.add_from_path(path, weight, file_format)
        return inst

    def add_from_path(self, path, weight=None, file_format="xyz"):
        if file_format != "xyz":
            raise DatasetError(f"Expect data file_format to be one of ['xyz'], got: {file_format}.")
        path = Path(path)
        if path.is_dir():
            files = []
            for root, _, names in os.walk(path, followlinks=True):
                files += [Path(root) / n for n in names if n.endswith(".xyz")]
            files = sorted(files)
        else:
            files = [path]
        configs = [Configuration.from_file(f, weight, file_format) for f in files]
        if not configs:
            raise DatasetError(f"No dataset file with file format `{file_format}` found at {path}.")
        self.configs.extend(configs)

    def get_configs(self):
        return self.configs

    def get_num_configs(self):
        return len(self.configs)

    def __len__(self):
        return len(self.configs)

    def __getitem__(self, idx):
        if isinstance(idx, slice):
            return Dataset(self.configs[idx])
        return self.configs[idx]
