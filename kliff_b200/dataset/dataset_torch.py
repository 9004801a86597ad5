"""`FingerprintsDataset` / `fingerprints_collate_fn` with the interface of
kliff/dataset/dataset_torch.py:10-58.

Besides a pickle-stream path, the dataset also wraps fingerprints that are already in memory
(`CalculatorTorch.create` has just produced them; reading the file back would only repeat work)."""
import os

import torch
from torch.utils.data import Dataset

from ..legacy.descriptors.descriptor import load_fingerprints


class FingerprintsDataset(Dataset):
    def __init__(self, source, transform=None):
        in_memory = isinstance(source, (list, tuple))
        self.fp = list(source) if in_memory else load_fingerprints(os.fspath(source))
        self.transform = transform
        for position, sample in enumerate(self.fp):
            sample["index"] = position  # configuration ordinal: the key into the packed GPU store

    def __getitem__(self, index):
        sample = self.fp[index]
        return self.transform(sample) if self.transform else sample

    def __len__(self):
        return len(self.fp)


def _as_tensor(value):
    return torch.from_numpy(value) if type(value).__module__ == "numpy" else value


def fingerprints_collate_fn(batch):
    """Samples stay separate (their atom counts differ); numpy values become tensors."""
    return [{key: _as_tensor(value) for key, value in sample.items()} for sample in batch]
