"""Import the UNMODIFIED reference (quadbio/cellmapper) from /root/reference in the build container.

TEST INFRASTRUCTURE ONLY. ``/root/reference`` does not exist on the GPU box, so nothing that runs
there may import this module; it is used by ``tests/golden/make_golden.py`` (which writes the
committed fixtures) and by the container-only test that re-checks the oracle against the live
reference.  The recipe is the one validated in SURVEY.md Appendix A: ``import cellmapper`` fails
here (scanpy / anndata / matplotlib are not installed and ``cellmapper/__init__.py:9`` asks for
package metadata), so a package shell plus three stub modules are registered first.
"""

from __future__ import annotations

import os
import sys
import types

import pandas as pd

REFERENCE_SRC = "/root/reference/src"


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_SRC, "cellmapper"))


class AnnData:
    """The minimal AnnData surface cellmapper.py / utils.py touch (constructor used at utils.py:109-118)."""

    def __init__(self, X=None, obs=None, var=None, uns=None, obsm=None, varm=None, layers=None, obsp=None):
        self.X = X
        self.obs = obs if obs is not None else pd.DataFrame(index=[str(i) for i in range(X.shape[0])])
        self.var = var if var is not None else pd.DataFrame(index=[f"g{i}" for i in range(X.shape[1])])
        self.uns = uns if uns is not None else {}
        self.obsm = obsm if obsm is not None else {}
        self.varm = varm if varm is not None else {}
        self.layers = layers if layers is not None else {}
        self.obsp = obsp if obsp is not None else {}

    def __getitem__(self, key):
        """``adata[:, gene_names]`` -- the one slicing form the path's consumers use (evaluate.py:345-350)."""
        rows, genes = key
        if not (isinstance(rows, slice) and rows == slice(None)):
            raise NotImplementedError("the stand-in AnnData only slices variables")
        pos = self.var_names.get_indexer(list(genes))
        assert (pos >= 0).all()
        take = lambda m: m[:, pos]  # noqa: E731
        return AnnData(X=take(self.X), obs=self.obs, var=self.var.iloc[pos], uns=self.uns, obsm=self.obsm,
                       layers={k: take(v) for k, v in self.layers.items()}, obsp=self.obsp)

    n_obs = property(lambda s: s.X.shape[0])
    n_vars = property(lambda s: s.X.shape[1])
    obs_names = property(lambda s: s.obs.index)
    var_names = property(lambda s: s.var.index)


def load():
    """Return (CellMapper, Neighbors, NeighborsResults, AnnData) of the reference itself."""
    if not available():
        raise RuntimeError("/root/reference is not present (GPU box?) -- the live reference cannot be imported")
    if "cellmapper" not in sys.modules or not hasattr(sys.modules["cellmapper"], "_graft_shim"):
        if REFERENCE_SRC not in sys.path:
            sys.path.insert(0, REFERENCE_SRC)
        pkg = types.ModuleType("cellmapper")
        pkg.__path__ = [os.path.join(REFERENCE_SRC, "cellmapper")]
        pkg.version = lambda name: "0.0.0"  # check.py:8 does `from . import version`
        pkg._graft_shim = True
        sys.modules["cellmapper"] = pkg
        for name in ("anndata", "scanpy", "scanpy.get", "matplotlib", "matplotlib.pyplot"):
            if name not in sys.modules:
                sys.modules[name] = types.ModuleType(name)
        sys.modules["anndata"].AnnData = AnnData
        sys.modules["scanpy.get"]._check_mask = None  # embedding.py:4, never called on the hot path
        sys.modules["scanpy.get"]._get_obs_rep = None
    from cellmapper.model.cellmapper import CellMapper  # noqa: E402
    from cellmapper.model.knn import Neighbors, NeighborsResults  # noqa: E402

    return CellMapper, Neighbors, NeighborsResults, AnnData
