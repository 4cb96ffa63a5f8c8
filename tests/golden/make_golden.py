"""Generate the committed golden vectors by running the UNMODIFIED reference code.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py

Writes ``tests/golden/*.npz``. Each file holds the inputs and what the reference's own
``CellMapper`` / ``Neighbors`` produced for them (sklearn 1.9.0, scipy 1.18.1, numpy 2.3.5,
pandas 3.0.2), so the oracle and the CUDA path can be checked on a box where the reference
does not exist.
"""

from __future__ import annotations

import os
import sys

import numpy as np
import pandas as pd
from scipy.sparse import csr_matrix

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from cellmapper_b200 import synth  # noqa: E402
from oracle import reference_shim  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
KERNELS = ["gaussian", "scarches", "inverse_distance", "equal"]


def csr_parts(prefix: str, m) -> dict:
    m = m.tocsr().copy()
    m.sort_indices()
    return {
        f"{prefix}_indptr": m.indptr.astype(np.int64),
        f"{prefix}_indices": m.indices.astype(np.int64),
        f"{prefix}_data": m.data,
        f"{prefix}_shape": np.array(m.shape, dtype=np.int64),
    }


def build_pair(n_q, n_r, d, n_comp, n_genes, padded=True):
    centres = synth.mixture_centres(n_comp, d)
    xr, cr = synth.mixture_embedding(n_r, centres, seed=1)
    xq, cq = synth.mixture_embedding(n_q, centres, seed=2)
    labels = synth.celltype_names(cr, padded=padded)
    umap = synth.umap_like(n_r)
    score = synth.numeric_obs(n_r)
    expr = synth.sparse_expression(cr, n_genes=n_genes, mean_nnz=n_genes / 8, nnz_clip=(4, n_genes // 2), n_top=n_genes // 4)
    return xr, xq, cr, cq, labels, umap, score, expr


def case_query_to_reference(name, n_q, n_r, d, k, n_comp, n_genes, padded=True):
    CellMapper, Neighbors, NeighborsResults, AnnData = reference_shim.load()
    xr, xq, cr, cq, labels, umap, score, expr = build_pair(n_q, n_r, d, n_comp, n_genes, padded)
    out = dict(xr=xr, xq=xq, labels=labels.astype(str), umap=umap, score=score, k=np.int64(k))
    out.update(csr_parts("expr", expr))
    for kernel in KERNELS:
        ref = AnnData(
            X=expr,
            obs=pd.DataFrame(
                {
                    "celltype": pd.Categorical(labels),
                    "score": score,
                    "score64": score.astype(np.float64),
                    "count": (score * 100).astype(np.int64),
                },
                index=[f"r{i}" for i in range(n_r)],
            ),
            obsm={"X_joint": xr, "X_umap": umap, "X_umap64": umap.astype(np.float64)},
            layers={"dense": np.asarray(expr.todense())},
        )
        qry = AnnData(
            X=csr_matrix((n_q, 5), dtype=np.float32),
            obs=pd.DataFrame(index=[f"q{i}" for i in range(n_q)]),
            obsm={"X_joint": xq},
        )
        cm = CellMapper(qry, ref)
        cm.map(
            use_rep="X_joint",
            obs_keys=["celltype", "score", "score64", "count"],
            obsm_keys=["X_umap", "X_umap64"],
            layer_key="X",
            n_neighbors=k,
            knn_method="sklearn",
            only_yx=True,
            mapping_method=kernel,
        )
        if kernel == KERNELS[0]:
            out["distances"] = cm.knn.yx.distances
            out["indices"] = cm.knn.yx.indices
            conn = cm.knn.yx.knn_graph_connectivities(kernel="gaussian")
            out.update(csr_parts("conn_gaussian", conn))
            cm.estimate_presence_score()
            out["presence_score"] = ref.obs["presence_score"].to_numpy()
            cm.estimate_presence_score(key_added="presence_log", log=True, percentile=(5, 95))
            out["presence_log"] = ref.obs["presence_log"].to_numpy()
        out.update(csr_parts(f"mm_{kernel}", cm.mapping_matrix))
        out[f"pred_{kernel}"] = qry.obs["celltype_pred"].to_numpy().astype(str)
        out[f"conf_{kernel}"] = qry.obs["celltype_conf"].to_numpy()
        out[f"score_{kernel}"] = qry.obs["score_pred"].to_numpy()
        out[f"score64_{kernel}"] = qry.obs["score64_pred"].to_numpy()
        out[f"count_{kernel}"] = qry.obs["count_pred"].to_numpy()
        out[f"umap_{kernel}"] = qry.obsm["X_umap_pred"]
        out[f"umap64_{kernel}"] = qry.obsm["X_umap64_pred"]
        out.update(csr_parts(f"imputed_{kernel}", cm.query_imputed.X))
        if kernel == KERNELS[0]:
            cm.map_layers("dense")
            out["imputed_dense"] = np.asarray(cm.query_imputed.X)
    np.savez_compressed(os.path.join(OUT, f"{name}.npz"), **out)
    print(name, {k_: getattr(v, "shape", v) for k_, v in list(out.items())[:6]})


def case_four_graphs(name, n_q, n_r, d, k, n_comp):
    """only_yx=False + jaccard / hnoca (cellmapper.py:287-301); also self-mapping identity."""
    CellMapper, Neighbors, NeighborsResults, AnnData = reference_shim.load()
    centres = synth.mixture_centres(n_comp, d)
    xr, cr = synth.mixture_embedding(n_r, centres, seed=1)
    xq, cq = synth.mixture_embedding(n_q, centres, seed=2)
    labels = synth.celltype_names(cr)
    out = dict(xr=xr, xq=xq, labels=labels.astype(str), k=np.int64(k))
    for method in ["jaccard", "hnoca"]:
        ref = AnnData(
            X=csr_matrix((n_r, 3), dtype=np.float32),
            obs=pd.DataFrame({"celltype": pd.Categorical(labels)}, index=[f"r{i}" for i in range(n_r)]),
            obsm={"X_joint": xr},
        )
        qry = AnnData(X=csr_matrix((n_q, 3), dtype=np.float32), obsm={"X_joint": xq})
        cm = CellMapper(qry, ref)
        cm.map(use_rep="X_joint", obs_keys="celltype", n_neighbors=k, only_yx=False, mapping_method=method)
        if method == "jaccard":
            for key in ("xx", "yy", "xy", "yx"):
                nr = getattr(cm.knn, key)
                out[f"{key}_indices"] = nr.indices
                out[f"{key}_distances"] = nr.distances
        out.update(csr_parts(f"mm_{method}", cm.mapping_matrix))
        out[f"pred_{method}"] = qry.obs["celltype_pred"].to_numpy().astype(str)
        out[f"conf_{method}"] = qry.obs["celltype_conf"].to_numpy()
    # self-mapping identity (tests/model/test_self_mapping.py:18-37): k=1 jaccard reproduces labels
    selfq = AnnData(
        X=csr_matrix((n_r, 3), dtype=np.float32),
        obs=pd.DataFrame({"celltype": pd.Categorical(labels)}, index=[f"r{i}" for i in range(n_r)]),
        obsm={"X_joint": xr},
    )
    cm = CellMapper(selfq)
    cm.map(use_rep="X_joint", obs_keys="celltype", n_neighbors=1, only_yx=False, mapping_method="jaccard")
    out["self_pred"] = selfq.obs["celltype_pred"].to_numpy().astype(str)
    np.savez_compressed(os.path.join(OUT, f"{name}.npz"), **out)
    print(name, "ok")


def case_ragged(name, n, d, k):
    """Precomputed ragged graph: load_precomputed_distances -> gaussian -> map_obs
    (cellmapper.py:493-532, knn.py:296-337, utils.py:129-219)."""
    CellMapper, Neighbors, NeighborsResults, AnnData = reference_shim.load()
    import sklearn.neighbors

    centres = synth.mixture_centres(4, d)
    x, c = synth.mixture_embedding(n, centres, seed=1)
    labels = synth.celltype_names(c)
    nn = sklearn.neighbors.NearestNeighbors(n_neighbors=k).fit(x)
    g = nn.kneighbors_graph(x, mode="distance").tolil()
    rng = np.random.default_rng(7)
    for i in range(n):  # drop a random number of edges per row -> ragged rows
        drop = rng.integers(0, k // 2)
        cols = [cc for cc in g.rows[i] if cc != i]
        for cc in rng.choice(cols, size=min(drop, len(cols)), replace=False):
            g[i, cc] = 0
    g = g.tocsr()
    g.eliminate_zeros()
    out = dict(labels=labels.astype(str))
    out.update(csr_parts("graph", g))
    for include_self in (None, True, False):
        ad = AnnData(
            X=csr_matrix((n, 3), dtype=np.float32),
            obs=pd.DataFrame({"celltype": pd.Categorical(labels)}, index=[f"c{i}" for i in range(n)]),
            obsp={"distances": g},
        )
        cm = CellMapper(ad)
        cm.load_precomputed_distances("distances", include_self=include_self)
        tag = {None: "none", True: "true", False: "false"}[include_self]
        out[f"indices_{tag}"] = cm.knn.yx.indices
        out[f"distances_{tag}"] = cm.knn.yx.distances
        cm.compute_mapping_matrix("gaussian")
        out.update(csr_parts(f"mm_{tag}", cm.mapping_matrix))
        cm.map_obs("celltype")
        out[f"pred_{tag}"] = ad.obs["celltype_pred"].to_numpy().astype(str)
        out[f"conf_{tag}"] = ad.obs["celltype_conf"].to_numpy()
    np.savez_compressed(os.path.join(OUT, f"{name}.npz"), **out)
    print(name, "ok")


def case_reference_unit_fixtures(name):
    """The reference's own tiny fixtures (tests/conftest.py:12-28) through its NeighborsResults."""
    CellMapper, Neighbors, NeighborsResults, AnnData = reference_shim.load()
    sd = np.array([[0.0, 1.0], [0.0, 2.0], [0.0, 3.0]])
    si = np.array([[0, 1], [1, 2], [2, 0]])
    nr = NeighborsResults(distances=sd, indices=si)
    out = dict(sample_distances=sd, sample_indices=si)
    for kernel in KERNELS:
        out.update(csr_parts(f"conn_{kernel}", nr.knn_graph_connectivities(kernel=kernel)))
    out.update(csr_parts("dist_graph", nr.knn_graph_distances))
    out.update(csr_parts("bool_adj", nr.boolean_adjacency()))
    x = np.array([[0, 0], [1, 0], [0, 1], [1, 1], [0.5, 0.5]], dtype=np.float64)
    y = x + 0.1
    nb = Neighbors(x, y)
    nb.compute_neighbors(n_neighbors=3, method="sklearn")
    out["small_x"], out["small_y"] = x, y
    for key in ("xx", "yy", "xy", "yx"):
        out[f"small_{key}_indices"] = getattr(nb, key).indices
        out[f"small_{key}_distances"] = getattr(nb, key).distances
    np.savez_compressed(os.path.join(OUT, f"{name}.npz"), **out)
    print(name, "ok")


def case_evaluate(name, n_q, n_r, d, k, n_comp, n_genes):
    """Consumers of the path: evaluate_expression_transfer (evaluate.py:236-424: pearson / rmse / js, groupby,
    test_var_key; round 2 also spearman), estimate_presence_score with groupby / log / percentile (evaluate.py:426-521), and an
    integer-valued layer (scipy promotes M @ X to float64)."""
    CellMapper, Neighbors, NeighborsResults, AnnData = reference_shim.load()
    xr, xq, cr, cq, labels, umap, score, expr = build_pair(n_q, n_r, d, n_comp, n_genes)
    rng = np.random.default_rng(21)
    ref_genes = np.array([f"g{i}" for i in range(n_genes)])
    # query genes: a shuffled subset of the reference's plus genes the reference does not have
    n_sh = (2 * n_genes) // 3
    q_genes = np.concatenate([rng.permutation(ref_genes)[:n_sh], np.array([f"q_only{i}" for i in range(17)])])
    q_genes = rng.permutation(q_genes)
    # original query expression: the (noisy) expression of a same-type reference cell, so correlations are non-trivial
    donor = np.array([rng.choice(np.flatnonzero(cr == c)) for c in cq])
    dense = np.asarray(expr[donor].todense()).astype(np.float32)
    dense = dense * (rng.random(dense.shape) < 0.8) + (rng.random(dense.shape) < 0.02) * rng.random(dense.shape).astype(np.float32)
    col_of = {g: i for i, g in enumerate(ref_genes)}
    qx = np.zeros((n_q, len(q_genes)), dtype=np.float32)
    for j, g in enumerate(q_genes):
        if g in col_of:
            qx[:, j] = dense[:, col_of[g]]
        else:
            qx[:, j] = rng.random(n_q) < 0.1
    qx[:, 3] = 0.0  # a gene without any mass in the query: NaN metrics
    qx_csr = csr_matrix(qx)
    batch = np.array(["b%d" % b for b in rng.integers(0, 3, n_q)], dtype=object)
    is_test = rng.random(len(q_genes)) < 0.5
    counts = csr_matrix((np.rint(np.expm1(expr.data) * 3).astype(np.int64), expr.indices, expr.indptr), shape=expr.shape)
    ref = AnnData(
        X=expr,
        obs=pd.DataFrame({"celltype": pd.Categorical(labels)}, index=[f"r{i}" for i in range(n_r)]),
        var=pd.DataFrame(index=ref_genes),
        obsm={"X_joint": xr},
        layers={"counts": counts},
    )
    qry = AnnData(
        X=qx_csr,
        obs=pd.DataFrame({"batch": pd.Categorical(batch)}, index=[f"q{i}" for i in range(n_q)]),
        var=pd.DataFrame({"is_test": is_test}, index=q_genes),
        obsm={"X_joint": xq},
    )
    cm = CellMapper(qry, ref)
    cm.map(use_rep="X_joint", layer_key="X", n_neighbors=k, knn_method="sklearn", only_yx=True, mapping_method="gaussian")
    out = dict(xr=xr, xq=xq, labels=labels.astype(str), k=np.int64(k), ref_genes=ref_genes.astype(str), q_genes=q_genes.astype(str),
               batch=batch.astype(str), is_test=is_test, indices=cm.knn.yx.indices, distances=cm.knn.yx.distances)
    out.update(csr_parts("expr", expr))
    out.update(csr_parts("counts", counts))
    out.update(csr_parts("qx", qx_csr))
    out.update(csr_parts("imputed", cm.query_imputed.X))
    import warnings

    for method in ("pearson", "rmse", "js", "spearman"):
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            cm.evaluate_expression_transfer(layer_key="X", method=method, groupby="batch", test_var_key="is_test")
        out[f"metric_{method}"] = qry.var[f"metric_{method}"].to_numpy().astype(np.float64)
        out[f"valid_{method}"] = qry.var[f"_is_valid_test_gene_{method}"].to_numpy().astype(bool)
        out[f"groups_{method}"] = qry.varm[f"metric_{method}"].to_numpy().astype(np.float64)
        out[f"group_names_{method}"] = np.array([str(c) for c in qry.varm[f"metric_{method}"].columns])
        out[f"average_{method}"] = np.float64(cm.expression_transfer_metrics["average"])
        out[f"n_test_{method}"] = np.int64(cm.expression_transfer_metrics["n_test_genes"])
    cm.estimate_presence_score(groupby="batch")
    out["presence_all"] = ref.obs["presence_score"].to_numpy()
    out["presence_groups"] = ref.obsm["presence_score"].to_numpy()
    out["presence_group_names"] = np.array([str(c) for c in ref.obsm["presence_score"].columns])
    out["presence_groups_dtype"] = np.array(str(ref.obsm["presence_score"].dtypes.iloc[0]))
    cm.estimate_presence_score(groupby="batch", key_added="p_log", log=True, percentile=(5, 90))
    out["presence_all_log"] = ref.obs["p_log"].to_numpy()
    out["presence_groups_log"] = ref.obsm["p_log"].to_numpy()
    cm.estimate_presence_score(groupby="batch", key_added="p_raw", percentile=(0, 100))
    out["presence_all_raw"] = ref.obs["p_raw"].to_numpy()
    out["presence_groups_raw"] = ref.obsm["p_raw"].to_numpy()
    cm.map_layers("counts")
    imp = cm.query_imputed.X
    out["imputed_counts_dtype"] = np.array(str(imp.dtype))
    out.update(csr_parts("imputed_counts", imp))
    np.savez_compressed(os.path.join(OUT, f"{name}.npz"), **out)
    print(name, "ok", {m: float(out[f"average_{m}"]) for m in ("pearson", "rmse", "js")}, out["presence_groups_dtype"], out["imputed_counts_dtype"])


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "evaluate":  # only the newest fixture
        case_evaluate("evaluate", n_q=400, n_r=900, d=30, k=15, n_comp=6, n_genes=240)
        sys.exit(0)
    case_query_to_reference("q2r_d30", n_q=300, n_r=700, d=30, k=30, n_comp=7, n_genes=200)
    case_query_to_reference("q2r_d10_kdtree", n_q=200, n_r=500, d=10, k=15, n_comp=5, n_genes=64, padded=False)
    case_query_to_reference("q2r_d50", n_q=257, n_r=1031, d=50, k=30, n_comp=12, n_genes=300, padded=False)
    case_four_graphs("four_graphs", n_q=300, n_r=400, d=20, k=10, n_comp=6)
    case_ragged("ragged_selfmap", n=300, d=16, k=12)
    case_reference_unit_fixtures("reference_unit_fixtures")
    case_evaluate("evaluate", n_q=400, n_r=900, d=30, k=15, n_comp=6, n_genes=240)
