"""The C-ABI library loads and exports every symbol include/wmf_b200.h declares, with the
signatures the ctypes binding expects. No compute calls (CPU only)."""
import ctypes
import os
import re

import pytest

from conftest import ROOT
from recmodel_b200 import _lib, build


@pytest.fixture(scope="module")
def lib():
    build.build()
    return _lib.load()


def declared_functions():
    text = open(os.path.join(ROOT, "include", "wmf_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(wmf_[a-z0-9_]+)\s*\(", text)))


def test_header_and_binding_agree(lib):
    names = declared_functions()
    assert len(names) >= 15
    assert set(names) == set(_lib.SIGNATURES), "include/wmf_b200.h and recmodel_b200/_lib.py drifted apart"
    for n in names:
        assert hasattr(lib, n), f"{n} declared in the header but not exported by libwmf_b200.so"


def header_arity(name):
    """Number of parameters of `name` in include/wmf_b200.h (comments stripped)."""
    text = open(os.path.join(ROOT, "include", "wmf_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    m = re.search(r"\b" + name + r"\s*\(([^;]*?)\)\s*;", text, flags=re.S)
    assert m, name
    args = m.group(1).strip()
    return 0 if args in ("", "void") else args.count(",") + 1


def test_binding_arity_matches_the_header():
    """Every ctypes signature has as many arguments as the declaration in the header (a drifted binding would pass
    garbage to a kernel), and the algorithm selectors of the binding are the header's."""
    for name, (_, argtypes) in _lib.SIGNATURES.items():
        assert header_arity(name) == len(argtypes), f"{name}: header and recmodel_b200/_lib.py disagree on the argument count"
    text = open(os.path.join(ROOT, "include", "wmf_b200.h")).read()
    for macro, value in (("WMF_ALGO_AUTO", _lib.ALGO_AUTO), ("WMF_ALGO_SIMT", _lib.ALGO_SIMT),
                         ("WMF_ALGO_TCGEN05", _lib.ALGO_TCGEN05), ("WMF_ALGO_TCGEN05_DIRECT", _lib.ALGO_TCGEN05_DIRECT)):
        m = re.search(r"#define\s+" + macro + r"\s+(\d+)", text)
        assert m and int(m.group(1)) == value, macro


def test_integration_example_matches_the_header():
    """INTEGRATION.md's reference-side ctypes stub declares wmf_als_half_step and its workspace query with the header's
    argument counts."""
    doc = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    for name in ("wmf_als_half_step", "wmf_als_half_step_workspace_bytes", "wmf_gram", "wmf_gram_workspace_bytes"):
        m = re.search(r"_lib\." + name + r"\.argtypes\s*=\s*(.*?)\n(?=_lib\.|\n|def )", doc, flags=re.S)
        assert m, f"{name}: no argtypes in INTEGRATION.md"
        expr = m.group(1)
        ns = {"ctypes": ctypes}
        assert len(eval(expr, ns)) == header_arity(name), f"{name}: INTEGRATION.md and the header disagree"


def test_version_and_error_string(lib):
    assert lib.wmf_version() >= 100
    assert isinstance(_lib.last_error(), str)


def test_argument_validation_needs_no_device(lib):
    # invalid arguments are rejected before any CUDA call
    assert lib.wmf_preprocess(None, 10, 7, 1.0, 1.0, None) == 1
    assert "mode" in _lib.last_error()
    assert lib.wmf_gram(None, 10, 0, 0, 0.1, 0, None, None, 0, None) == 1
    assert lib.wmf_als_half_step(None, None, None, 5, 7, None, 0, None, 4, 100000, None, 0, None, 4, 0, None, 0, None) == 1
    assert lib.wmf_gram_workspace_bytes(1000, 64) > 0
    assert lib.wmf_als_half_step_workspace_bytes(1000, 500, 256, 0) > 0
    # the tcgen05 pipeline takes every width up to 256, with or without biases (whitened factors, dual + primal kernels)
    for f, bias in ((64, 0), (65, 1), (128, 0), (129, 1), (256, 0)):
        assert lib.wmf_als_half_step_supports(_lib.ALGO_TCGEN05, f, bias) == 1
    assert lib.wmf_als_half_step_supports(_lib.ALGO_TCGEN05, 257, 1) == 0
    assert 16 <= lib.wmf_als_dual_max_entries() <= 128
    assert lib.wmf_sddmm_loss_workspace_bytes(10**6) >= 3 * 8


def test_no_cpu_fallback_without_device(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a device is present")
    with pytest.raises(_lib.WMFLibraryError):
        _lib.require_device()
    from recmodel_b200 import WMF
    import numpy as np
    import scipy.sparse
    m = WMF(num_items=5, num_users=4, dim=3, gamma=0.1, weighted=True)
    assert m.items.shape == (5, 3) and m.users is None
    R = scipy.sparse.random(4, 5, density=0.5, format="csr", dtype=np.float32, random_state=0)
    with pytest.raises(_lib.WMFLibraryError):
        m.train(R, 1, eval_mat=R, count_mat=R, cores=1)


def test_product_never_imports_oracle():
    """The product path must not import, call or link anything under oracle/."""
    pkg = os.path.join(ROOT, "recmodel_b200")
    pat = re.compile(r"^\s*(from|import)\s+oracle\b|importlib\.import_module\(.oracle|oracle/|wmf_oracle", re.M)
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, fn)).read()
                assert not pat.search(src), f"{os.path.join(dirpath, fn)} references the oracle"
