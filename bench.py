#!/usr/bin/env python
"""Benchmark of the k-NN mapping hot path (BASELINE.json metric: query cells mapped / second).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --steps K --warmup W    # the reference's CPU path (oracle port)

A *step* is one pass of the hot path over one batch of synthetic input:
search (k=30) -> gaussian kernel -> row-normalised mapping matrix -> celltype vote + X_umap SpMM.
Default workload: the configuration BASELINE.json's metric is quoted on, C3 = 1.5M query -> 1.5M
reference, d=50 (it fits one B200).  For N>1 the 1.5M queries are sharded over the ranks and the
reference is replicated (north_star's default partitioning): the job is the same at every N, i.e.
STRONG scaling; the only collective on the data path is the all-reduce of the three kernel-bandwidth
statistics.  `--workload C2` runs BASELINE config 2 (100k -> 100k) the same way.

`value`  : whole-job cells/s with inputs resident in HBM (CUDA events, L2 flushed between steps).
`e2e`    : the same metric through the public `CellMapper.map()` API with HOST buffers (pinned),
           host->device copies of the embeddings/labels and device->host reads of the results inside
           the timed region.
Rank 0 prints ONE JSON line.
"""

from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (n_query per GPU, n_reference, d, n_components)
    "C1": (5_000, 5_000, 30, 8),
    "C2": (100_000, 100_000, 50, 32),
    "C3": (1_500_000, 1_500_000, 50, 32),
}
K = 30
UMAP_DIMS = 2


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return {"hbm_gbs": p["hbm_gbs"], "bf16_tflops": p["bf16_tflops"], "bf16_tflops_sustained": p.get("bf16_tflops_sustained"), "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


def make_inputs(name: str, rank: int = 0, world: int = 1):
    """Synthetic Gaussian-mixture embeddings of the named shape; every rank draws the
    same arrays and keeps its contiguous block of query rows."""
    from cellmapper_b200 import synth
    from cellmapper_b200.dist import shard_bounds

    n_q, n_r, d, n_comp = WORKLOADS[name]
    centres = synth.mixture_centres(n_comp, d)
    xr, cr = synth.mixture_embedding(n_r, centres, seed=1)
    xq, _ = synth.mixture_embedding(n_q, centres, seed=2)
    lo, hi = shard_bounds(n_q, world, rank)
    xq = np.ascontiguousarray(xq[lo:hi])
    labels = synth.celltype_names(cr)
    umap = synth.umap_like(n_r, UMAP_DIMS)
    return xr, xq, cr, labels, umap


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons of one GPU during the timed region (NVML)."""

    def __init__(self, index: int, period: float = 0.02):
        super().__init__(daemon=True)
        self.index, self.period = index, period
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop_evt = threading.Event()
        self.ok = False
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = int(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            # the first query of each kind initialises driver state for several milliseconds while holding a lock the
            # kernel launches need (it showed up as one slow step in every short-step run): pay that before timing
            pynvml.nvmlDeviceGetClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            try:
                pynvml.nvmlDeviceGetCurrentClocksEventReasons(self.h)
            except Exception:
                pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
            self.ok = True
        except Exception:  # pragma: no cover
            self.ok = False

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
            getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonHwPowerBrakeSlowdown", 0x80): "hw_power_brake",
        }
        while not self._stop_evt.is_set():
            try:
                self.samples.append(int(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                try:
                    mask = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
                except Exception:
                    mask = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                for bit, nm in names.items():
                    if mask & bit:
                        self.reasons.add(nm)
            except Exception:
                pass
            time.sleep(self.period)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=2)
        med = float(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}


def physical_gpu_index(local_rank: int) -> int:
    vis = os.environ.get("CUDA_VISIBLE_DEVICES")
    if vis:
        try:
            return int(vis.split(",")[local_rank])
        except Exception:
            return local_rank
    return local_rank


# ------------------------------------------------------------------------------------------------
# reference arm: the reference's own CPU implementation of the path (oracle port: the same sklearn /
# numpy / scipy calls at the reference's call sites), on a bounded sample of the workload
# ------------------------------------------------------------------------------------------------
def cpu_sample_queries(name: str) -> int:
    n_q, n_r, d, _ = WORKLOADS[name]
    # brute force is linear in n_q: keep one step at ~3-10 s of host time (about 4e9 pairs/s at d=50 on 16 cores measured)
    budget_pairs = 3.0e10
    return int(max(1000, min(n_q, budget_pairs / n_r)))


def run_cpu_path(xr, xq_sample, labels, umap):
    from oracle import cellmapper_oracle as orc

    t0 = time.perf_counter()
    out = orc.run_path(xr, xq_sample, labels=labels, obsm=umap, n_neighbors=K, kernel="gaussian")
    return time.perf_counter() - t0, out


def cpu_threads() -> int:
    try:
        from threadpoolctl import threadpool_info

        return max([p.get("num_threads", 1) for p in threadpool_info()] + [1])
    except Exception:
        return os.cpu_count() or 1


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    name = args.workload
    xr, xq, cr, labels, umap = make_inputs(name)
    ns = cpu_sample_queries(name)
    xs = xq[:ns]
    times = []
    for it in range(args.warmup + args.steps):
        dt, _ = run_cpu_path(xr, xs, labels, umap)
        if it >= args.warmup:
            times.append(dt)
    total = float(np.sum(times))
    value = ns * args.steps / total
    n_q, n_r, d, _ = WORKLOADS[name]
    line = {
        "impl": "reference",
        "metric": "query cells mapped/sec",
        "value": value,
        "unit": "cells/s",
        "n_gpus": args.gpus,
        "steps": args.steps,
        "warmup": args.warmup,
        "ms_per_step": 1e3 * total / args.steps,
        "higher_is_better": True,
        "scaling": "strong",
        "vs_baseline": None,
        "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": f"{name}: {n_q} query -> {n_r} reference, d={d}, k={K}, gaussian kernel, celltype + X_umap transfer", "sample": f"first {ns} queries against the full reference"},
        "cpu_baseline": {"value": value, "unit": "cells/s", "cores": cpu_threads(), "kind": "port", "sample": f"{ns} of {n_q} queries x full {n_r} reference per step"},
        "e2e": {"value": value, "unit": "cells/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# this repo's arm
# ------------------------------------------------------------------------------------------------
def b200_arm(args):
    import pandas as pd
    import torch
    import torch.distributed as tdist
    from scipy.sparse import csr_matrix

    from cellmapper_b200 import CellMapper, _lib, device
    from cellmapper_b200 import dist as cmd
    from cellmapper_b200._anndata import AnnData
    from cellmapper_b200.cellmapper import sorted_category_codes
    from cellmapper_b200.knn import sklearn_like_dist_mode

    rank, world, local_rank = cmd.init_from_env()
    if world != args.gpus and rank == 0:
        print(f"warning: --gpus {args.gpus} but WORLD_SIZE={world}; using {world}", file=sys.stderr)
    _lib.require_device(local_rank)
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    lib = _lib.load()
    name = args.workload
    n_q_total, n_r, d, _ = WORKLOADS[name]
    xr, xq, cr, labels, umap = make_inputs(name, rank, world)
    n_q = xq.shape[0]  # this rank's block of query rows

    # pinned host buffers (the e2e arm copies from these every step)
    def pin(a):
        t = torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
        return t, t.numpy()

    xr_t, xr_p = pin(xr)
    xq_t, xq_p = pin(xq)
    umap_t, umap_p = pin(umap)
    label_series = pd.Series(pd.Categorical(labels))
    cats, codes = sorted_category_codes(label_series)
    codes_t, _ = pin(codes)

    def barrier():
        if world > 1:
            tdist.barrier()

    allreduce = cmd.allreduce_sum if world > 1 else None
    mode = sklearn_like_dist_mode(np.float32, d, K, n_r)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)  # > 126 MB L2

    # ---------------- device-resident arm ----------------
    xr_d, xq_d = xr_t.to(dev), xq_t.to(dev)
    umap_d, codes_d = umap_t.to(dev), codes_t.to(dev)

    PHASES = ["search", "edge_stats", "kernel_to_csr", "vote", "spmm"]
    search_stats = []  # device int64[4] per search: [fallback rows, -, candidates re-ranked, (query tile, reference tile) pairs evaluated]

    def device_step(marks=None):
        def mark():
            if marks is not None:
                e = torch.cuda.Event(enable_timing=True)
                e.record()
                marks.append(e)

        mark()
        # N > 1: the reference side of the coarse cells is computed block by block on the ranks and all-gathered (NCCL)
        ref_cells = cmd.assign_reference_sharded(xr_d, K) if world > 1 else None
        dd, ii, st_search = device.knn_search(xq_d, xr_d, K, dist_mode=mode, return_stats=True, ref_cells=ref_cells)
        search_stats.append(st_search)
        mark()
        st = device.edge_stats(dd, ii, allreduce=allreduce, need_std=False)
        mark()
        ip, cols, vals = device.edge_kernel_to_csr(dd, ii, "gaussian", st, normalize=True)
        mark()
        code, conf = device.vote_argmax(ip, cols, vals, codes_d, len(cats))
        mark()
        emb = device.spmm(ip, cols, vals, umap_d)
        mark()
        return dd, ii, code, conf, emb

    res = None
    for _ in range(args.warmup):
        # keep the previous step's results alive while the next one runs, exactly as the timed loop does: the caching
        # allocator then owns both sets of output blocks before timing starts (otherwise the second timed step paid
        # the cudaMalloc of the second set: +1..6 ms on short steps)
        res = device_step()
    torch.cuda.synchronize()
    lib.cm_profile_enable(1)
    # NVML queries take a driver lock that can hold up kernel launches for milliseconds: only the rank that
    # reports the clocks samples them
    sampler = ClockSampler(physical_gpu_index(local_rank)) if (rank == 0 and not os.environ.get("CM_BENCH_NO_CLOCKS")) else None
    barrier()
    torch.cuda.synchronize()
    if sampler is not None:
        sampler.start()
    launches0 = lib.cm_launch_count()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    phase_ms = np.zeros(4)
    import ctypes

    buf4 = (ctypes.c_float * 4)()
    step_marks = []
    search_stats.clear()
    for s in range(args.steps):
        flush.zero_()
        ev[s][0].record()
        marks = []
        res = device_step(marks)
        ev[s][1].record()
        step_marks.append(marks)
        if lib.cm_profile_last_knn_ms(buf4) == 0:  # synchronises on the search's own events only
            phase_ms += np.array(list(buf4))
    torch.cuda.synchronize()
    barrier()
    launches = lib.cm_launch_count() - launches0
    clocks = sampler.stop() if sampler is not None else None
    lib.cm_profile_enable(0)
    t_dev = sum(a.elapsed_time(b) for a, b in ev) / 1e3
    step_ms = [a.elapsed_time(b) for a, b in ev]
    st_host = torch.stack(search_stats).double().mean(0).cpu().numpy()
    tiles_scanned, fallback_rows = float(st_host[3]), float(st_host[0])
    path_ms = {n: float(np.mean([m[i].elapsed_time(m[i + 1]) for m in step_marks])) for i, n in enumerate(PHASES)}
    phase_ms /= args.steps
    tt = torch.tensor([t_dev], dtype=torch.float64, device=dev)
    if world > 1:
        tdist.all_reduce(tt, op=tdist.ReduceOp.MAX)
    t_dev = float(tt.item())
    value = n_q_total * args.steps / t_dev
    dd, ii, code, conf, emb = res


    # ---------------- end-to-end arm: public API, host buffers ----------------
    ref_ad = AnnData(
        X=csr_matrix((n_r, 1), dtype=np.float32),
        obs=pd.DataFrame({"celltype": label_series.values}, index=pd.RangeIndex(n_r).astype(str)),
        obsm={"X_joint": xr_p, "X_umap": umap_p},
    )
    qry_index = pd.RangeIndex(n_q).astype(str)

    def e2e_step():
        qry_ad = AnnData(X=csr_matrix((n_q, 1), dtype=np.float32), obs=pd.DataFrame(index=qry_index), obsm={"X_joint": xq_p})
        cm = CellMapper(qry_ad, ref_ad, allreduce=allreduce, upload_replicated=cmd.upload_replicated if world > 1 else None,
                        reference_cells=cmd.assign_reference_sharded if world > 1 else None)
        cm.map(use_rep="X_joint", obs_keys="celltype", obsm_keys="X_umap", n_neighbors=K, only_yx=True, mapping_method="gaussian")
        return qry_ad

    for _ in range(args.warmup):
        e2e_step()
    torch.cuda.synchronize()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        out_ad = e2e_step()
    torch.cuda.synchronize()
    barrier()
    t_e2e = time.perf_counter() - t0
    tt = torch.tensor([t_e2e], dtype=torch.float64, device=dev)
    if world > 1:
        tdist.all_reduce(tt, op=tdist.ReduceOp.MAX)
    t_e2e = float(tt.item())
    e2e_value = n_q_total * args.steps / t_e2e
    # whole job, all ranks: every rank uploads the replicated reference side and its own query block
    # the replicated reference side crosses PCIe once in total (each rank uploads 1/world of it, NCCL all-gather)
    h2d = (xr.nbytes + codes.nbytes + umap.nbytes) + n_q_total * d * 4
    d2h = n_q_total * (4 + 4 + UMAP_DIMS * 4) + 8 * world

    if rank != 0:
        return

    # ---------------- roofline of the dominant kernel (mma_topk: tensor pipe) ----------------
    # The search is exact but pruned: whole reference cells whose triangle-inequality lower bound exceeds
    # every threshold of a query tile are never multiplied.  `achieved` counts the
    # algorithmic flops of the pairs the kernel actually evaluated (2*d per pair, un-padded d, one
    # fp32-equivalent product per pair-dimension); `brute_force_equivalent` is 2*n_q*n_r*d over the same
    # time, the figure comparable with an exhaustive scan (it may exceed the peak: work not done).
    peaks = measured_peaks()
    peak = peaks["bf16_tflops"] / 3.0  # three fp16 passes per fp32-accurate product
    t_mma = phase_ms[1] / 1e3
    pairs = tiles_scanned * 128.0 * 128.0
    achieved = 2.0 * pairs * d / t_mma / 1e12 if t_mma > 0 else None
    n_pairs_all = float(-(-n_q // 128)) * float(-(-n_r // 128))
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if os.path.exists(tpath):
        traffic = json.load(open(tpath)).get(name)
    roofline = {
        "bound": "tensor",
        "kernel": "mma_topk_kernel",
        "achieved": achieved,
        "peak": peak,
        "unit": "TFLOP/s",
        "frac": (achieved / peak) if achieved else None,
        "traffic": traffic,
        "peak_source": f"{peaks['source']} cuBLAS bf16 burst {peaks['bf16_tflops']} TFLOP/s / 3 split-precision passes",
        "avg_launch_ms": phase_ms[1],
        "pairs_evaluated_frac": tiles_scanned / n_pairs_all,
        "brute_force_equivalent": {"tflops": 2.0 * n_q * n_r * d / t_mma / 1e12, "frac_of_peak": 2.0 * n_q * n_r * d / t_mma / 1e12 / peak},
        "phases_ms": {"prep": phase_ms[0], "mma_topk": phase_ms[1], "rerank": phase_ms[2], "exact_fallback": phase_ms[3]},
        "fallback_rows": fallback_rows,
    }
    # the same kernel with pruning switched off (exhaustive scan) on a slice of whole waves of query tiles:
    # the tensor-pipe figure of the kernel itself, outside the timed region
    if not args.no_exhaustive_probe:
        n_slice = min(n_q, 2 * 148 * 128)
        lib.cm_debug_probe_flags(32)
        lib.cm_profile_enable(1)
        ex_ms = []
        for i in range(3):
            device.knn_search(xq_d[:n_slice], xr_d, K, dist_mode=mode)
            if lib.cm_profile_last_knn_ms(buf4) == 0 and i:
                ex_ms.append(buf4[1])
        lib.cm_profile_enable(0)
        lib.cm_debug_probe_flags(0)
        if ex_ms:
            t_ex = float(np.mean(ex_ms)) / 1e3
            ach_ex = 2.0 * n_slice * n_r * d / t_ex / 1e12
            roofline["exhaustive_scan_probe"] = {"queries": n_slice, "ms": t_ex * 1e3, "achieved": ach_ex, "frac": ach_ex / peak}
    # HBM-side phases: algorithmic bytes per query (SURVEY.md 8d / DESIGN.md 4) over the CUDA-event time of the call
    hbm_bytes = {"edge_stats": K * 16.0, "kernel_to_csr": 484.0, "vote": 368.0, "spmm": K * 8.0 + K * UMAP_DIMS * 4.0 + UMAP_DIMS * 4.0}
    hbm_phases = {
        n: {"ms": path_ms[n], "achieved_gbs": hbm_bytes[n] * n_q / (path_ms[n] * 1e-3) / 1e9, "frac_of_hbm_peak": hbm_bytes[n] * n_q / (path_ms[n] * 1e-3) / 1e9 / peaks["hbm_gbs"]}
        for n in hbm_bytes
    }

    # ---------------- CPU baseline (oracle port) + recall against it ----------------
    cpu = None
    recall = None
    if world == 1 and not args.no_cpu_baseline:
        ns = cpu_sample_queries(name)
        run_cpu_path(xr, xq[: min(ns, 2000)], labels, umap)  # warm-up, discarded
        dt, out = run_cpu_path(xr, xq[:ns], labels, umap)
        cpu = {"value": ns / dt, "unit": "cells/s", "cores": cpu_threads(), "kind": "port", "sample": f"first {ns} of {n_q_total} queries x full {n_r} reference, 1 run after warm-up", "phases_s": out["seconds"]}
        got = ii[:ns].cpu().numpy()
        want = out["indices"]
        hits = sum(len(set(a.tolist()) & set(b.tolist())) for a, b in zip(got, want))
        recall = hits / want.size
        pred = np.asarray(cats)[code[:ns].cpu().numpy()].astype(str)
        label_match = float((pred == out["pred"].astype(str)).mean())
    line = {
        "metric": "query cells mapped/sec",
        "value": value,
        "unit": "cells/s",
        "n_gpus": world,
        "steps": args.steps,
        "warmup": args.warmup,
        "ms_per_step": 1e3 * t_dev / args.steps,
        "higher_is_better": True,
        "scaling": "strong",
        "vs_baseline": None,
        "dtype": "f32 (fp16x3 split tensor-core candidates, f64 exact re-rank)",
        "data": "synthetic",
        "config": {
            "workload": f"{name}: {n_q_total} query -> {n_r} reference, d={d}, k={K}, gaussian kernel, celltype + X_umap transfer; queries sharded over the ranks ({n_q} on rank 0), reference replicated",
            "l2": "flushed between timed steps (256 MB write)",
            "parallelism": f"query-sharded x{world}",
        },
        "e2e": {"value": e2e_value, "unit": "cells/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h), "ms_per_step": 1e3 * t_e2e / args.steps},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": roofline,
        "path_phases_ms": path_ms,
        "hbm_phases": hbm_phases,
        "step_ms": step_ms,
        "cpu_baseline": cpu,
        "recall_at_30": recall,
    }
    if cpu is not None:
        line["label_agreement_vs_cpu"] = label_match
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="C3", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-exhaustive-probe", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.impl == "reference":
        reference_arm(args)
    else:
        b200_arm(args)
        try:
            import torch.distributed as tdist

            if tdist.is_initialized():
                tdist.destroy_process_group()
        except Exception:
            pass


if __name__ == "__main__":
    main()
