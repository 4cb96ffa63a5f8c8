"""cellmapper_b200 -- B200-native (sm_100a) implementation of CellMapper's k-NN mapping hot path.

Drop-in for the path ``compute_neighbors -> compute_mapping_matrix -> map_obs / map_obsm / map_layers``
of quadbio/cellmapper with ``method="b200"``; see DESIGN.md and INTEGRATION.md.
"""

from .logging import logger

__version__ = "0.1.0"
__all__ = ["CellMapper", "Neighbors", "NeighborsResults", "logger", "__version__"]


def __getattr__(name):  # lazy: importing the package must not need torch / the native library
    if name == "CellMapper":
        from .cellmapper import CellMapper

        return CellMapper
    if name in ("Neighbors", "NeighborsResults"):
        from . import knn

        return getattr(knn, name)
    raise AttributeError(f"module 'cellmapper_b200' has no attribute {name!r}")
