"""The C-ABI library loads and exports every symbol include/cellmapper_b200.h declares (no GPU needed)."""

from __future__ import annotations

import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "cellmapper_b200.h")


def declared_functions():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    # development-build-only probes (-DCM_DEV_PROBES) are not part of the shipping ABI
    text = re.sub(r"#ifdef CM_DEV_PROBES.*?#endif", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(cm_[a-z0-9_]+)\s*\(", text)))


@pytest.fixture(scope="module")
def lib():
    from cellmapper_b200 import _lib, build

    build.build()
    return _lib.load()


def test_header_declares_expected_surface():
    names = declared_functions()
    for must in ("cm_knn_search", "cm_edge_kernel_to_csr", "cm_vote_argmax", "cm_spmm_csr_dense", "cm_spgemm_fill"):
        assert must in names


def test_library_exports_every_declared_symbol(lib):
    missing = [n for n in declared_functions() if not hasattr(lib, n)]
    assert not missing, f"declared in the header but not exported: {missing}"


def test_binding_covers_every_declared_symbol():
    from cellmapper_b200 import _lib

    missing = [n for n in declared_functions() if n not in _lib.SIGNATURES]
    assert not missing, f"declared in the header but not bound in _lib.SIGNATURES: {missing}"


def test_abi_version_and_error_string(lib):
    assert lib.cm_abi_version() == 1
    # argument validation happens before any CUDA call, so it works without a GPU
    rc = lib.cm_knn_search(None, 1, 1, None, 1, 1, 1, 0, 1, 0, 0, 0, None, None, None, 0, None, None)
    assert rc == 1
    assert b"null pointer" in lib.cm_last_error()


def test_workspace_query_needs_no_gpu(lib):
    small = lib.cm_knn_workspace_bytes(1000, 1000, 30, 30, 0)
    big = lib.cm_knn_workspace_bytes(100000, 100000, 50, 30, 0)
    assert 0 < small < big < (1 << 31)
    assert lib.cm_knn_workspace_bytes(1000, 1000, 300, 30, 0) == 256  # d too large for the MMA path -> exact kernel


def test_sass_has_blackwell_instructions():
    """tcgen05.mma / tcgen05.ld / bulk-async copy must be in the binary (UTCHMMA / LDTM / UBLKCP)."""
    import shutil
    import subprocess

    from cellmapper_b200 import _lib

    if shutil.which("cuobjdump") is None:
        pytest.skip("cuobjdump not on PATH")
    sass = subprocess.run(["cuobjdump", "-sass", _lib.LIBPATH], capture_output=True, text=True).stdout
    for mnemonic in ("UTCHMMA", "LDTM", "UBLKCP"):
        assert mnemonic in sass, mnemonic


def test_no_cpu_fallback_without_gpu():
    import numpy as np
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from cellmapper_b200 import Neighbors

    with pytest.raises(RuntimeError, match="no CPU fallback"):
        Neighbors(np.zeros((8, 4), dtype=np.float32)).compute_neighbors(n_neighbors=2)


def test_product_package_never_imports_the_oracle():
    """The product path may not import oracle/ or scikit-learn (no CPU route to the results)."""
    pkg = os.path.join(ROOT, "cellmapper_b200")
    bad = re.compile(r"^\s*(from|import)\s+(oracle|sklearn)\b", re.M)
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert not bad.search(src), f"{f} imports the oracle or scikit-learn"


def test_shipping_library_has_no_debug_switches(lib):
    """The process-global development probes exist only in -DCM_DEV_PROBES builds (tools/), never in the product."""
    for name in ("cm_debug_probe_flags", "cm_debug_probe_prof"):
        assert not hasattr(lib, name), f"{name} is exported by the shipping library"
