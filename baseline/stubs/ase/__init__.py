"""Stand-in for the `ase` package: the reference imports it at module level
(kliff/dataset/configuration.py:20-27, dataset.py:14) but the timed paths never call into it."""


class Atoms:  # noqa: D101
    pass
