"""AnnData access.  The real ``anndata`` package is used when it is importable; otherwise a minimal
stand-in with the attributes this path touches (X, obs, var, uns, obsm, varm, layers, obsp, n_obs,
n_vars, obs_names, var_names) so the package, its tests and bench.py run in images without anndata.
"""

from __future__ import annotations

import pandas as pd

try:  # pragma: no cover - depends on the image
    from anndata import AnnData  # type: ignore
except Exception:  # anndata not installed

    class AnnData:  # type: ignore[no-redef]
        def __init__(self, X=None, obs=None, var=None, uns=None, obsm=None, varm=None, layers=None, obsp=None):
            self.X = X
            n_obs = X.shape[0] if X is not None else (len(obs) if obs is not None else 0)
            n_vars = X.shape[1] if X is not None else (len(var) if var is not None else 0)
            self.obs = obs if obs is not None else pd.DataFrame(index=[str(i) for i in range(n_obs)])
            self.var = var if var is not None else pd.DataFrame(index=[f"g{i}" for i in range(n_vars)])
            self.uns = uns if uns is not None else {}
            self.obsm = obsm if obsm is not None else {}
            self.varm = varm if varm is not None else {}
            self.layers = layers if layers is not None else {}
            self.obsp = obsp if obsp is not None else {}

        @property
        def n_obs(self) -> int:
            return self.X.shape[0] if self.X is not None else len(self.obs)

        @property
        def n_vars(self) -> int:
            return self.X.shape[1] if self.X is not None else len(self.var)

        @property
        def obs_names(self):
            return self.obs.index

        @property
        def var_names(self):
            return self.var.index

        def __repr__(self):
            return f"AnnData(n_obs={self.n_obs}, n_vars={self.n_vars})"
