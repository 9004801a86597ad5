"""Stand-in for `ase.build.bulk` (kliff/dataset/configuration.py:21, :657)."""


def bulk(*a, **k):
    raise NotImplementedError("ase stub")
