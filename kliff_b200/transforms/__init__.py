"""Mirror of `kliff.transforms` for the one piece that shares the accelerated path: the new-API
symmetry-function `Descriptor` (forward / backward VJP)."""
