"""The oracle against the golden vectors produced by the unmodified reference (CPU, no GPU)."""

from __future__ import annotations

import numpy as np
import pytest
from conftest import assert_csr_equal, golden_csr, load_golden, neighbours_match

from oracle import cellmapper_oracle as orc
from oracle import reference_shim

KERNELS = ["gaussian", "scarches", "inverse_distance", "equal"]
Q2R = ["q2r_d30", "q2r_d10_kdtree", "q2r_d50"]


@pytest.mark.parametrize("name", Q2R)
def test_search_matches_reference(name):
    g = load_golden(name)
    d, i = orc.search_sklearn(g["xr"], g["xq"], int(g["k"]))
    np.testing.assert_array_equal(i, g["indices"])
    np.testing.assert_array_equal(d, g["distances"])
    assert d.dtype == np.float64 and i.dtype == np.int64


@pytest.mark.parametrize("name", Q2R)
def test_independent_bruteforce_agrees_with_sklearn(name):
    g = load_golden(name)
    d, i = orc.bruteforce_knn_f64(g["xr"], g["xq"], int(g["k"]))
    assert neighbours_match(i, d, g["indices"], g["distances"]) == 0
    # with sklearn's rounding rule (float32 sqrt on the brute path, float64 on the KD-tree path)
    # the independent direct-difference distances reproduce sklearn's bit for bit
    d, i = orc.bruteforce_knn_f64(g["xr"], g["xq"], int(g["k"]), sklearn_rounding=True)
    same = i == g["indices"]
    brute = orc.sklearn_uses_brute(g["xr"].shape[1], int(g["k"]), g["xr"].shape[0])
    if brute:
        np.testing.assert_array_equal(d[same], g["distances"][same])
        assert np.array_equal(g["distances"], g["distances"].astype(np.float32).astype(np.float64))
    else:
        np.testing.assert_allclose(d[same], g["distances"][same], rtol=1e-15)


@pytest.mark.parametrize("name", Q2R)
@pytest.mark.parametrize("kernel", KERNELS)
def test_mapping_matrix_bit_exact(name, kernel):
    g = load_golden(name)
    m = orc.mapping_matrix_from_neighbors(g["distances"], g["indices"], g["xr"].shape[0], kernel)
    ref = golden_csr(g, f"mm_{kernel}")
    assert m.dtype == np.float32 and m.indices.dtype == np.int32
    assert_csr_equal(m, ref)


@pytest.mark.parametrize("name", Q2R)
def test_connectivities_and_presence(name):
    g = load_golden(name)
    conn = orc.connectivities_csr(g["distances"], g["indices"], g["xr"].shape[0], "gaussian")
    assert_csr_equal(conn, golden_csr(g, "conn_gaussian"))
    np.testing.assert_array_equal(
        orc.presence_scores(g["distances"], g["indices"], g["xr"].shape[0]), g["presence_score"]
    )
    np.testing.assert_array_equal(
        orc.presence_scores(g["distances"], g["indices"], g["xr"].shape[0], log=True, percentile=(5, 95)),
        g["presence_log"],
    )


@pytest.mark.parametrize("name", Q2R)
@pytest.mark.parametrize("kernel", KERNELS)
def test_transfers_bit_exact(name, kernel):
    g = load_golden(name)
    m = golden_csr(g, f"mm_{kernel}").astype(np.float32)
    pred, conf, cats, codes = orc.map_obs_categorical(m, g["labels"])
    np.testing.assert_array_equal(pred.astype(str), g[f"pred_{kernel}"])
    np.testing.assert_array_equal(conf, g[f"conf_{kernel}"])
    assert conf.dtype == np.float32
    # independent loop restatement: same codes, same float32 confidences
    cats2, ref_codes = orc.onehot_sorted(g["labels"])
    np.testing.assert_array_equal(cats2.astype(str), cats.astype(str))
    c2, f2 = orc.vote_argmax_loops(m.indptr, m.indices, m.data, ref_codes, len(cats2))
    np.testing.assert_array_equal(c2, codes)
    np.testing.assert_array_equal(f2, conf)
    for key, src in (("score", g["score"]), ("score64", g["score"].astype(np.float64)), ("count", (g["score"] * 100).astype(np.int64))):
        out = orc.map_obs_numerical(m, src)
        np.testing.assert_array_equal(out, g[f"{key}_{kernel}"])
        assert out.dtype == g[f"{key}_{kernel}"].dtype
    np.testing.assert_array_equal(orc.map_obsm(m, g["umap"]), g[f"umap_{kernel}"])
    np.testing.assert_array_equal(orc.map_obsm(m, g["umap"].astype(np.float64)), g[f"umap64_{kernel}"])
    np.testing.assert_array_equal(orc.spmm_loops(m.indptr, m.indices, m.data, g["umap"]), g[f"umap_{kernel}"])
    expr = golden_csr(g, "expr")
    assert_csr_equal(orc.map_layers(m, expr), golden_csr(g, f"imputed_{kernel}"))
    if kernel == "gaussian":
        np.testing.assert_array_equal(orc.map_layers(m, np.asarray(expr.todense())), g["imputed_dense"])


def test_unpadded_names_sort_lexicographically():
    g = load_golden("q2r_d50")
    cats, _ = orc.onehot_sorted(g["labels"])
    assert list(cats[:3]) == ["0", "1", "10"]


@pytest.mark.parametrize("method", ["jaccard", "hnoca"])
def test_jaccard_hnoca(method):
    g = load_golden("four_graphs")
    res = orc.search_sklearn_all(g["xr"], g["xq"], int(g["k"]))
    for key in ("xx", "yy", "xy", "yx"):
        np.testing.assert_array_equal(res[key][1], g[f"{key}_indices"])
    m = orc.jaccard_mapping(g["xx_indices"], g["yy_indices"], g["xy_indices"], g["yx_indices"], method)
    assert_csr_equal(m, golden_csr(g, f"mm_{method}"))
    pred, conf, _, _ = orc.map_obs_categorical(m, g["labels"])
    np.testing.assert_array_equal(pred.astype(str), g[f"pred_{method}"])
    np.testing.assert_array_equal(conf, g[f"conf_{method}"])


def test_self_mapping_identity():
    """k=1 jaccard self-mapping reproduces the labels exactly (reference tests/model/test_self_mapping.py:18-37)."""
    g = load_golden("four_graphs")
    np.testing.assert_array_equal(g["self_pred"], g["labels"])
    x = g["xr"]
    res = orc.search_sklearn_all(x, x, 1)
    m = orc.jaccard_mapping(res["xx"][1], res["yy"][1], res["xy"][1], res["yx"][1], "jaccard")
    pred, _, _, _ = orc.map_obs_categorical(m, g["labels"])
    np.testing.assert_array_equal(pred.astype(str), g["labels"])


@pytest.mark.parametrize("tag,include_self", [("none", None), ("true", True), ("false", False)])
def test_ragged_precomputed(tag, include_self):
    g = load_golden("ragged_selfmap")
    graph = golden_csr(g, "graph")
    idx, dist = orc.extract_neighbors_from_distances(graph, include_self=include_self)
    np.testing.assert_array_equal(idx, g[f"indices_{tag}"])
    np.testing.assert_array_equal(dist, g[f"distances_{tag}"])
    assert (idx == -1).any() and np.isinf(dist).any()
    m = orc.mapping_matrix_from_neighbors(dist, idx, idx.shape[0], "gaussian")
    assert_csr_equal(m, golden_csr(g, f"mm_{tag}"))
    pred, conf, _, _ = orc.map_obs_categorical(m, g["labels"])
    np.testing.assert_array_equal(pred.astype(str), g[f"pred_{tag}"])
    np.testing.assert_array_equal(conf, g[f"conf_{tag}"])


def test_reference_unit_fixtures():
    """The reference's own tiny fixtures (tests/conftest.py:12-28 there)."""
    g = load_golden("reference_unit_fixtures")
    sd, si = g["sample_distances"], g["sample_indices"]
    for kernel in KERNELS:
        assert_csr_equal(orc.connectivities_csr(sd, si, 3, kernel), golden_csr(g, f"conn_{kernel}"))
    assert_csr_equal(orc.boolean_adjacency(si, 3), golden_csr(g, "bool_adj"))
    res = orc.search_sklearn_all(g["small_x"], g["small_y"], 3)
    for key in ("xx", "yy", "xy", "yx"):
        np.testing.assert_array_equal(res[key][1], g[f"small_{key}_indices"])
        np.testing.assert_array_equal(res[key][0], g[f"small_{key}_distances"])


def test_no_finite_distance_raises():
    with pytest.raises(ValueError):
        orc.connectivities_csr(np.full((2, 2), np.inf), np.full((2, 2), -1), 2, "gaussian")


@pytest.mark.skipif(not reference_shim.available(), reason="/root/reference only exists in the build container")
def test_oracle_against_live_reference():
    """Differential check on a fresh seed against the reference imported live."""
    import pandas as pd
    from scipy.sparse import csr_matrix

    from cellmapper_b200 import synth

    CellMapper, _, _, AnnData = reference_shim.load()
    centres = synth.mixture_centres(5, 24, seed=11)
    xr, cr = synth.mixture_embedding(900, centres, seed=12)
    xq, _ = synth.mixture_embedding(400, centres, seed=13)
    labels = synth.celltype_names(cr)
    ref = AnnData(
        X=csr_matrix((900, 3), dtype=np.float32),
        obs=pd.DataFrame({"celltype": pd.Categorical(labels)}, index=[f"r{i}" for i in range(900)]),
        obsm={"X_joint": xr, "X_umap": synth.umap_like(900)},
    )
    qry = AnnData(X=csr_matrix((400, 3), dtype=np.float32), obsm={"X_joint": xq})
    cm = CellMapper(qry, ref).map(use_rep="X_joint", obs_keys="celltype", obsm_keys="X_umap", only_yx=True, mapping_method="scarches")
    out = orc.run_path(xr, xq, labels=labels, obsm=ref.obsm["X_umap"], kernel="scarches")
    np.testing.assert_array_equal(out["indices"], cm.knn.yx.indices)
    assert_csr_equal(out["mapping_matrix"], cm.mapping_matrix)
    np.testing.assert_array_equal(out["pred"].astype(str), qry.obs["celltype_pred"].to_numpy().astype(str))
    np.testing.assert_array_equal(out["conf"], qry.obs["celltype_conf"].to_numpy())
    np.testing.assert_array_equal(out["obsm_pred"], qry.obsm["X_umap_pred"])


# --------------------------------------------------------------------------------------------
# consumers of the path (evaluate.npz: evaluate_expression_transfer + presence score with groups)
# --------------------------------------------------------------------------------------------
def _aligned_evaluate_matrices(g):
    """(imputed, original) restricted to the shared genes in the order the reference uses (evaluate.py:343-350)."""
    import pandas as pd

    ref_genes, q_genes = pd.Index(g["ref_genes"]), pd.Index(g["q_genes"])
    shared = ref_genes.intersection(q_genes)
    imp = golden_csr(g, "imputed")[:, ref_genes.get_indexer(shared)]
    orig = golden_csr(g, "qx")[:, q_genes.get_indexer(shared)]
    return imp, orig, shared, q_genes


@pytest.mark.parametrize("method", ["pearson", "rmse", "js", "spearman"])
def test_expression_transfer_metrics_match_reference(method):
    g = load_golden("evaluate")
    imp, orig, shared, q_genes = _aligned_evaluate_matrices(g)
    pos = q_genes.get_indexer(shared)
    got = orc.expression_transfer_metrics(imp, orig, method)
    np.testing.assert_array_equal(got.astype(np.float64), g[f"metric_{method}"][pos])
    assert np.isnan(g[f"metric_{method}"][np.setdiff1d(np.arange(len(q_genes)), pos)]).all()
    for gi, name in enumerate(g[f"group_names_{method}"]):
        got = orc.expression_transfer_metrics(imp, orig, method, mask=g["batch"] == name)
        np.testing.assert_array_equal(got.astype(np.float64), g[f"groups_{method}"][pos, gi])
    valid = g[f"valid_{method}"]
    assert int(valid.sum()) == int(g[f"n_test_{method}"])
    np.testing.assert_allclose(np.mean(g[f"metric_{method}"][valid]), float(g[f"average_{method}"]), rtol=1e-12)


@pytest.mark.parametrize("tag,log,pct", [("", False, (1, 99)), ("_log", True, (5, 90)), ("_raw", False, (0, 100))])
def test_presence_scores_with_groups_match_reference(tag, log, pct):
    g = load_golden("evaluate")
    overall, groups, names = orc.presence_scores_grouped(g["distances"], g["indices"], g["xr"].shape[0], g["batch"], log=log, percentile=pct)
    assert [str(n) for n in names] == [str(n) for n in g["presence_group_names"]]
    np.testing.assert_array_equal(overall, g[f"presence_all{tag}"])
    assert str(g["presence_groups_dtype"]) == "float32" and groups.dtype == np.float32
    np.testing.assert_array_equal(groups, g[f"presence_groups{tag}"])


def test_integer_layer_is_transferred_in_float64():
    """scipy promotes the float32 mapping matrix with an integer layer: float64 result (cellmapper.py:372-373)."""
    g = load_golden("evaluate")
    m = orc.mapping_matrix_from_neighbors(g["distances"], g["indices"], g["xr"].shape[0], "gaussian")
    out = orc.map_layers(m, golden_csr(g, "counts"))
    assert str(g["imputed_counts_dtype"]) == "float64" and out.dtype == np.float64
    assert_csr_equal(out, golden_csr(g, "imputed_counts"))
