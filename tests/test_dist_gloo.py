"""world_size-2 gloo tests (CPU) of the multi-GPU plumbing in cellmapper_b200.dist: shard bounds,
the bandwidth all-reduce of the query-sharded mode and the all-gather + merge of the
reference-sharded mode.  The CUDA kernels are replaced by oracle stand-ins (tests may use oracle/)."""

from __future__ import annotations

import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank: int, world: int, port: int, out_dir: str):
    sys.path.insert(0, ROOT)
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    from cellmapper_b200 import dist as cmd
    from cellmapper_b200 import synth
    from oracle import cellmapper_oracle as orc

    r, w, _ = cmd.init_from_env(backend="gloo")
    assert (r, w) == (rank, world) and cmd.world() == (rank, world)

    centres = synth.mixture_centres(5, 20)
    xr, cr = synth.mixture_embedding(1203, centres, seed=1)
    xq, _ = synth.mixture_embedding(701, centres, seed=2)
    labels = synth.celltype_names(cr)
    k = 15

    # ---- query-sharded: local search + local stats, all-reduce, kernel with the GLOBAL bandwidth
    lo, hi = cmd.shard_bounds(xq.shape[0], w, r)
    d_loc, i_loc = orc.search_sklearn(xr, xq[lo:hi], k)
    first = torch.tensor([d_loc.sum(), 0.0, float(d_loc.size)], dtype=torch.float64)
    cmd.allreduce_sum(first)
    mean = float(first[0] / first[2])
    second = torch.tensor([0.0, float(((d_loc - mean) ** 2).sum()), 0.0], dtype=torch.float64)
    cmd.allreduce_sum(second)
    sigma = mean
    std = float(np.sqrt(second[1] / first[2]))
    w_loc = np.exp(-(d_loc**2) / (2 * sigma**2))
    w_loc = (w_loc / w_loc.sum(1, keepdims=True)).astype(np.float32)
    gathered_w = cmd.gather_rows(torch.from_numpy(w_loc))
    gathered_i = cmd.gather_rows(torch.from_numpy(i_loc))

    # ---- reference-sharded: every rank searches its block for ALL queries, all-gather, merge
    rlo, rhi = cmd.shard_bounds(xr.shape[0], w, r)

    def search(q, rl, kk, off):
        dd, ii = orc.bruteforce_knn_f64(rl.numpy(), q.numpy(), kk)
        return torch.from_numpy(dd), torch.from_numpy(ii + off)

    def merge(cd, ci, kk):
        cd2 = cd.permute(1, 0, 2).reshape(cd.shape[1], -1).numpy()
        ci2 = ci.permute(1, 0, 2).reshape(ci.shape[1], -1).numpy()
        key = np.where(ci2 < 0, np.inf, cd2)
        order = np.lexsort((ci2, key), axis=1)[:, :kk]
        rows = np.arange(cd2.shape[0])[:, None]
        return torch.from_numpy(key[rows, order]), torch.from_numpy(ci2[rows, order])

    md, mi = cmd.reference_sharded_search(torch.from_numpy(xq), torch.from_numpy(xr[rlo:rhi]), rlo, k, search, merge)
    # ---- reference-sharded presence score (BASELINE config 5): merged graph replicated, every rank sums the columns
    # of ITS block of reference cells, all-gather of the blocks (uneven: 1203 rows over 2 ranks)
    def block_sums(dd, ii, st, n_loc, lo_):
        conn = orc.connectivities_csr(dd.numpy(), ii.numpy(), xr.shape[0], "gaussian")
        return torch.from_numpy(np.asarray(conn.sum(axis=0)).ravel()[lo_ : lo_ + n_loc].copy())

    pd_, pi_, presence = cmd.presence_reference_sharded(
        torch.from_numpy(xq), torch.from_numpy(xr[rlo:rhi]), rlo, k, 0,
        search_merge=lambda q, rl, off, kk, mode: cmd.reference_sharded_search(q, rl, off, kk, search, merge),
        edge_stats=lambda dd, ii: None, block_sums=block_sums)
    assert torch.equal(pi_, mi) and presence.shape == (xr.shape[0],)

    # ---- reference-sharded expression transfer: partial CSR x CSR products, all-gather, per-block sum
    import scipy.sparse as sp

    def spgemm_cpu(ip, cc, vv, xip, xcc, xvv, n_genes):
        a = sp.csr_matrix((vv.numpy(), cc.numpy(), ip.numpy()), shape=(ip.numel() - 1, xip.numel() - 1))
        b = sp.csr_matrix((xvv.numpy(), xcc.numpy(), xip.numpy()), shape=(xip.numel() - 1, n_genes))
        c = (a @ b).tocsr()
        c.sort_indices()
        return torch.from_numpy(c.indptr.astype(np.int64)), torch.from_numpy(c.indices.astype(np.int32)), torch.from_numpy(c.data.astype(np.float32))

    rng = np.random.default_rng(7)
    n_genes = 97
    xmat = sp.random(xr.shape[0], n_genes, density=0.08, format="csr", dtype=np.float32, random_state=5)
    xmat.sort_indices()
    d_all, i_all = orc.search_sklearn(xr, xq, k)
    wts = np.exp(-(d_all**2) / (2 * d_all.mean() ** 2))
    wts = (wts / wts.sum(1, keepdims=True)).astype(np.float32)
    order = np.argsort(i_all, axis=1)
    mm = sp.csr_matrix(
        (np.take_along_axis(wts, order, 1).ravel(), np.take_along_axis(i_all, order, 1).ravel().astype(np.int32),
         np.arange(0, xq.shape[0] * k + 1, k, dtype=np.int32)), shape=(xq.shape[0], xr.shape[0]))
    xs = xmat[rlo:rhi]
    oip, ocols, ovals, (qlo_b, qhi_b) = cmd.spgemm_reference_sharded(
        torch.from_numpy(mm.indptr), torch.from_numpy(mm.indices), torch.from_numpy(mm.data),
        torch.from_numpy(xs.indptr.astype(np.int64)), torch.from_numpy(xs.indices), torch.from_numpy(xs.data),
        rlo, rhi, n_genes, spgemm_cpu)
    full = (mm @ xmat).tocsr()
    full.sort_indices()
    blk = full[qlo_b:qhi_b]
    assert (qlo_b, qhi_b) == cmd.shard_bounds(xq.shape[0], w, r)
    assert np.array_equal(oip.numpy(), blk.indptr) and np.array_equal(ocols.numpy(), blk.indices)
    np.testing.assert_allclose(ovals.numpy(), blk.data, rtol=2e-6)
    sub_ip, sub_c, sub_v = cmd.csr_column_block(torch.from_numpy(mm.indptr), torch.from_numpy(mm.indices), torch.from_numpy(mm.data), rlo, rhi)
    ref_sub = mm[:, rlo:rhi].tocsr()
    assert np.array_equal(sub_ip.numpy(), ref_sub.indptr) and np.array_equal(sub_c.numpy(), ref_sub.indices) and np.array_equal(sub_v.numpy(), ref_sub.data)

    # ---- reference side of the coarse cells: block per rank, all-gather of the cell numbers, max of the radii
    def fake_assign(rt, kk, lo, hi):
        cell = (torch.arange(lo, hi) % 251).to(torch.uint8)
        rad = torch.zeros(256, dtype=torch.int32)
        rad[r] = 1000 + r  # every rank contributes its own maximum somewhere
        rad[200] = 7 * (r + 1)
        return cell, rad

    got_cells = cmd.assign_reference_sharded(torch.from_numpy(xr), k, assign=fake_assign)
    if w == 1:
        assert got_cells is None
    else:
        cell_all, rad_all = got_cells
        assert torch.equal(cell_all, (torch.arange(xr.shape[0]) % 251).to(torch.uint8))
        assert int(rad_all[200]) == 7 * w and all(int(rad_all[j]) == 1000 + j for j in range(w))
        assert cmd.assign_reference_sharded(torch.from_numpy(xr), k, assign=lambda *a: None) is None

    # ---- replicated reference arrays: every rank uploads its row block, all-gather (odd row count, 1-D and 2-D)
    up2 = cmd.upload_replicated(xr, device=torch.device("cpu"), min_bytes=0)
    up1 = cmd.upload_replicated(cr.astype(np.int32), device=torch.device("cpu"), min_bytes=0)
    small = cmd.upload_replicated(xr[:3], device=torch.device("cpu"))  # below min_bytes: direct upload
    assert up2.shape == xr.shape and torch.equal(up2, torch.from_numpy(xr))
    assert torch.equal(up1, torch.from_numpy(cr.astype(np.int32))) and torch.equal(small, torch.from_numpy(xr[:3]))
    if rank == 0:
        np.savez(
            os.path.join(out_dir, "out.npz"), w=gathered_w.numpy(), i=gathered_i.numpy(), mean=mean, std=std,
            md=md.numpy(), mi=mi.numpy(), presence=presence.numpy(),
        )  # fmt: skip
    torch.distributed.barrier()
    torch.distributed.destroy_process_group()


def test_shard_bounds_cover_everything():
    from cellmapper_b200.dist import shard_bounds

    for n in (0, 1, 7, 100, 1_500_000):
        for world in (1, 2, 3, 8):
            blocks = [shard_bounds(n, world, r) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(blocks[:-1], blocks[1:]))
            sizes = [b[1] - b[0] for b in blocks]
            assert max(sizes) - min(sizes) <= 1


@pytest.mark.timeout(300)
def test_two_rank_gloo_matches_single_process(tmp_path):
    from cellmapper_b200 import synth
    from oracle import cellmapper_oracle as orc

    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    out = np.load(tmp_path / "out.npz")

    centres = synth.mixture_centres(5, 20)
    xr, _ = synth.mixture_embedding(1203, centres, seed=1)
    xq, _ = synth.mixture_embedding(701, centres, seed=2)
    d, i = orc.search_sklearn(xr, xq, 15)
    # global bandwidth statistics from the two all-reduces == single-process numpy
    np.testing.assert_allclose(out["mean"], d.mean(), rtol=1e-14)
    np.testing.assert_allclose(out["std"], d.std(), rtol=1e-12)
    # sharded mapping-matrix rows == unsharded ones (same global sigma)
    w = np.exp(-(d**2) / (2 * d.mean() ** 2))
    w = (w / w.sum(1, keepdims=True)).astype(np.float32)
    np.testing.assert_array_equal(out["i"], i)
    np.testing.assert_allclose(out["w"], w, rtol=1e-6)
    # reference-sharded search + merge == global exact search
    gd, gi = orc.bruteforce_knn_f64(xr, xq, 15)
    np.testing.assert_array_equal(out["mi"], gi)
    np.testing.assert_allclose(out["md"], gd, rtol=1e-15)
    # sharded presence sums == column sums of the global graph
    conn = orc.connectivities_csr(gd, gi, xr.shape[0], "gaussian")
    np.testing.assert_array_equal(out["presence"], np.asarray(conn.sum(axis=0)).ravel())
