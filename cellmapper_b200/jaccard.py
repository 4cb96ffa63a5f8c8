"""jaccard / hnoca mapping matrix on the device (reference: cellmapper.py:287-301)."""

from __future__ import annotations

from . import device


def jaccard_mapping_device(knn, hnoca: bool = False):
    """Row-normalised float32 mapping matrix (indptr int32, cols int32, vals float32) from the four
    neighbour graphs of a ``Neighbors`` object: shared-neighbour counts J = yx @ xx.T + yy @ xy.T,
    J/(4k-J) (jaccard) or (J/(2k-J))^2 (hnoca), then the normalisation of cellmapper.py:99-137."""
    indptr, cols, vals64 = device.jaccard(
        knn.yx.indices_device, knn.yy.indices_device, knn.xx.indices_device, knn.xy.indices_device, hnoca=hnoca
    )
    vals32, _zero = device.csr_row_normalize(indptr, vals64)
    return indptr, cols, vals32
