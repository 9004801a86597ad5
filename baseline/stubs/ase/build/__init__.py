"""Stand-in for `ase.build`: imported at module level by the reference
(kliff/dataset/configuration.py:21), never called on the timed paths."""
