from .neighbor import NeighborList, NeighborListError, assemble_forces, assemble_stress

__all__ = ["NeighborList", "NeighborListError", "assemble_forces", "assemble_stress"]
