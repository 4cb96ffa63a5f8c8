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

#: where the reference's package lives: the read-only checkout in the build container, or the copy that
#: ``__graft_entry__.build()`` places under the git-ignored ``oracle/_ref/`` so that it travels to the GPU box
#: (the reference is pure Python: "building" it is copying ``src/cellmapper``; nothing of it is committed).
_CANDIDATES = ("/root/reference/src", os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref"))


def reference_src() -> str | None:
    for root in _CANDIDATES:
        if os.path.isfile(os.path.join(root, "cellmapper", "model", "cellmapper.py")):
            return root
    return None


REFERENCE_SRC = reference_src() or _CANDIDATES[0]


def available() -> bool:
    return reference_src() is not None


def vendor_reference() -> str | None:
    """Copy the reference's package (pure Python, unmodified) from the read-only checkout into ``oracle/_ref/``
    (git-ignored, travels with the repo snapshot): the recipe ``__graft_entry__.build()`` runs.  Returns the
    destination, or None when there is no checkout to copy from."""
    import shutil

    src = os.path.join(_CANDIDATES[0], "cellmapper")
    if not os.path.isdir(src):
        return None
    dst = os.path.join(_CANDIDATES[1], "cellmapper")
    if os.path.isdir(dst):
        shutil.rmtree(dst)
    shutil.copytree(src, dst, ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
    return dst


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
    src = reference_src()
    if src is None:
        raise RuntimeError("neither /root/reference nor oracle/_ref is present -- the live reference cannot be imported")
    if "cellmapper" not in sys.modules or not hasattr(sys.modules["cellmapper"], "_graft_shim"):
        if src not in sys.path:
            sys.path.insert(0, src)
        pkg = types.ModuleType("cellmapper")
        pkg.__path__ = [os.path.join(src, "cellmapper")]
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


# ------------------------------------------------------------------------------------------------
# the path driven through the reference's own public API (bench.py's reference arm / cpu_baseline)
# ------------------------------------------------------------------------------------------------
def _quiet():
    import logging

    logging.getLogger("cellmapper.logging").setLevel(logging.ERROR)  # the reference's logger is named after its module


def run_map(xr, xq, labels=None, obsm=None, layer=None, n_neighbors=30, kernel="gaussian") -> dict:
    """compute_neighbors(method="sklearn", only_yx=True) -> compute_mapping_matrix -> map_obs / map_obsm / map_layers
    of the UNMODIFIED reference, the calls ``CellMapper.map`` makes in its order (cellmapper.py:465-484), timed per
    phase.  Returns results + ``seconds``."""
    import time

    import numpy as np
    from scipy.sparse import csr_matrix

    CellMapper, _, _, AD = load()
    _quiet()
    n_r, n_q = xr.shape[0], xq.shape[0]
    ref = AD(
        X=layer if layer is not None else csr_matrix((n_r, 1), dtype=np.float32),
        obs=pd.DataFrame({"celltype": pd.Categorical(labels)} if labels is not None else {}, index=pd.RangeIndex(n_r).astype(str)),
        obsm={"X_joint": xr, **({"X_umap": obsm} if obsm is not None else {})},
    )
    qry = AD(X=csr_matrix((n_q, 1), dtype=np.float32), obs=pd.DataFrame(index=pd.RangeIndex(n_q).astype(str)), obsm={"X_joint": xq})
    cm = CellMapper(qry, ref)
    t, out = {}, {}
    t0 = time.perf_counter()
    cm.compute_neighbors(n_neighbors=n_neighbors, use_rep="X_joint", method="sklearn", only_yx=True)
    t["search"] = time.perf_counter() - t0
    t0 = time.perf_counter()
    cm.compute_mapping_matrix(method=kernel)
    t["kernel"] = time.perf_counter() - t0
    out.update(indices=cm.knn.yx.indices, distances=cm.knn.yx.distances, mapping_matrix=cm.mapping_matrix)
    if labels is not None:
        t0 = time.perf_counter()
        cm.map_obs(key="celltype")
        t["map_obs"] = time.perf_counter() - t0
        out.update(pred=qry.obs["celltype_pred"].to_numpy().astype(str), conf=qry.obs["celltype_conf"].to_numpy())
    if obsm is not None:
        t0 = time.perf_counter()
        cm.map_obsm(key="X_umap")
        t["map_obsm"] = time.perf_counter() - t0
        out["obsm_pred"] = qry.obsm["X_umap_pred"]
    if layer is not None:
        t0 = time.perf_counter()
        cm.map_layers(key="X")
        t["map_layers"] = time.perf_counter() - t0
        out["layer_pred"] = cm.query_imputed.X
    out["seconds"] = t
    return out


def run_presence(xr, xq, n_neighbors=30) -> dict:
    """compute_neighbors(sklearn, only_yx=True) -> estimate_presence_score of the unmodified reference
    (evaluate.py:426-480): BASELINE config 5's path."""
    import time

    import numpy as np
    from scipy.sparse import csr_matrix

    CellMapper, _, _, AD = load()
    _quiet()
    n_r, n_q = xr.shape[0], xq.shape[0]
    ref = AD(X=csr_matrix((n_r, 1), dtype=np.float32), obs=pd.DataFrame(index=pd.RangeIndex(n_r).astype(str)), obsm={"X_joint": xr})
    qry = AD(X=csr_matrix((n_q, 1), dtype=np.float32), obs=pd.DataFrame(index=pd.RangeIndex(n_q).astype(str)), obsm={"X_joint": xq})
    cm = CellMapper(qry, ref)
    t = {}
    t0 = time.perf_counter()
    cm.compute_neighbors(n_neighbors=n_neighbors, use_rep="X_joint", method="sklearn", only_yx=True)
    t["search"] = time.perf_counter() - t0
    t0 = time.perf_counter()
    cm.estimate_presence_score()
    t["presence"] = time.perf_counter() - t0
    return dict(indices=cm.knn.yx.indices, distances=cm.knn.yx.distances, presence=ref.obs["presence_score"].to_numpy(), seconds=t)
