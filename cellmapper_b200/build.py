"""Build libcellmapper_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python -m cellmapper_b200.build [--force] [--verbose]

The shared library exports the C ABI declared in include/cellmapper_b200.h and is loaded with
ctypes by cellmapper_b200._lib.  It lives at cellmapper_b200/lib/libcellmapper_b200.so (git-ignored,
but it travels to the GPU box with the repo snapshot).
"""

from __future__ import annotations

import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
LIBPATH = os.path.join(LIBDIR, "libcellmapper_b200.so")
STAMP = os.path.join(LIBDIR, "build.stamp")

SOURCES = ["cabi.cu", "scan.cu", "knn_exact.cu", "knn_mma.cu", "graph_kernel.cu", "transfer.cu", "jaccard.cu", "evaluate.cu"]
NVCC_FLAGS = [
    "-gencode",
    "arch=compute_100a,code=sm_100a",
    "-O3",
    "-lineinfo",
    "-std=c++17",
    "--expt-relaxed-constexpr",
    "--extended-lambda",
    "-Xcompiler",
    "-fPIC",
    "-Xptxas",
    "-v",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    raise RuntimeError("nvcc not found")


def _fingerprint() -> str:
    h = hashlib.sha256()
    for root in (CSRC, os.path.join(os.path.dirname(HERE), "include")):
        for name in sorted(os.listdir(root)):
            if name.endswith((".cu", ".cuh", ".h")):
                with open(os.path.join(root, name), "rb") as f:
                    h.update(name.encode())
                    h.update(f.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def sources() -> list[str]:
    return [os.path.join(CSRC, s) for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]


def build(force: bool = False, verbose: bool = False, dev_probes: bool = False, variant: str = "", defines: tuple = ()) -> str:
    """``dev_probes``: a separate library (lib/libcellmapper_b200_probes.so, -DCM_DEV_PROBES) that also exports
    the development probes cm_debug_probe_flags / cm_debug_probe_prof; tools/ load it through CM_LIBPATH.  The
    shipping library never contains them."""
    from concurrent.futures import ThreadPoolExecutor

    os.makedirs(LIBDIR, exist_ok=True)
    # variant / defines: tuning builds (tools/ab_search.py): lib/libcellmapper_b200_<variant>.so with extra -D flags
    tag = ("_probes" if dev_probes else "") + (f"_{variant}" if variant else "")
    libpath = LIBPATH.replace(".so", f"{tag}.so")
    stamp = STAMP + tag
    flags = NVCC_FLAGS + (["-DCM_DEV_PROBES"] if dev_probes else []) + [f"-D{d}" for d in defines]
    fp = _fingerprint() + tag + " ".join(defines)
    if not force and os.path.exists(libpath) and os.path.exists(stamp) and open(stamp).read().strip() == fp:
        return libpath

    def compile_one(src):
        obj = os.path.join(LIBDIR, os.path.basename(src).replace(".cu", f"{tag}.o"))
        cmd = [_nvcc(), *flags, "-c", src, "-o", obj]
        res = subprocess.run(cmd, capture_output=True, text=True)
        return obj, cmd, res

    objs = []
    logs = []
    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as pool:
        for obj, cmd, res in pool.map(compile_one, sources()):
            logs.append(f"$ {' '.join(cmd)}\n{res.stdout}{res.stderr}")
            if res.returncode != 0:
                raise RuntimeError(f"nvcc failed for {cmd[-3]}:\n{res.stdout}\n{res.stderr}")
            objs.append(obj)
    cmd = [_nvcc(), "-shared", "-o", libpath, *objs, "-gencode", "arch=compute_100a,code=sm_100a", "-lcudart"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    logs.append(f"$ {' '.join(cmd)}\n{res.stdout}{res.stderr}")
    if res.returncode != 0:
        raise RuntimeError(f"link failed:\n{res.stdout}\n{res.stderr}")
    with open(os.path.join(LIBDIR, f"build{tag}.log"), "w") as f:
        f.write("\n".join(logs))
    with open(stamp, "w") as f:
        f.write(fp)
    if verbose:
        print("\n".join(logs))
    return libpath


if __name__ == "__main__":
    path = build(force="--force" in sys.argv, verbose="--verbose" in sys.argv, dev_probes="--dev-probes" in sys.argv)
    print(path)
