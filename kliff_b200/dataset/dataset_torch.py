"""`FingerprintsDataset` / `fingerprints_collate_fn` (kliff/dataset/dataset_torch.py:10-58)."""
import torch
from torch.utils.data import Dataset

from ..legacy.descriptors.descriptor import load_fingerprints


class FingerprintsDataset(Dataset):
    def __init__(self, filename, transform=None):
        self.fp = load_fingerprints(filename)
        for i, f in enumerate(self.fp):
            f["index"] = i
        self.transform = transform

    def __len__(self):
        return len(self.fp)

    def __getitem__(self, index):
        sample = self.fp[index]
        if self.transform:
            sample = self.transform(sample)
        return sample


def fingerprints_collate_fn(batch):
    """List of samples with numpy values turned into tensors; no stacking."""
    out = []
    for sample in batch:
        ts = {}
        for key, value in sample.items():
            if type(value).__module__ == "numpy":
                value = torch.from_numpy(value)
            ts[key] = value
        out.append(ts)
    return out
