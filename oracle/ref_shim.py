"""Import the UNMODIFIED reference WMF from /root/reference (this container only).

TEST INFRASTRUCTURE - never imported by the product path (recmodel_b200/*). Used by
tests/golden/make_golden.py to produce the committed golden vectors, and by the optional
`-m "not gpu"` test that re-checks the oracle against the live reference when
/root/reference is mounted. /root/reference does not exist on the GPU box.

Stubs: a `sharedmem` module (imported at base_model.py:8, never used on the
WMF path), an empty `RecModel` package object so RecModel/__init__.py:2 (compiled Cython
import) is skipped, and MKLThreads (base_model.py:181-214, needs libmkl_rt.so) replaced
by a threadpoolctl wrapper so `cores` keeps its meaning. No reference file is edited.
"""
import importlib
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("WMF_REFERENCE_ROOT", "/root/reference")


def available():
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "RecModel", "wmf_model.py"))


class _BlasThreads:
    """Drop-in for the reference's MKLThreads context manager."""

    def __init__(self, num_threads):
        self._n = int(num_threads)
        self._ctx = None

    def __enter__(self):
        try:
            from threadpoolctl import threadpool_limits
            self._ctx = threadpool_limits(limits=self._n)
            self._ctx.__enter__()
        except Exception:  # threadpoolctl missing: run with whatever BLAS does
            self._ctx = None
        return self

    def __exit__(self, *exc):
        if self._ctx is not None:
            self._ctx.__exit__(*exc)
        return False


def load_reference_wmf():
    """Returns the reference's WMF class (RecModel/wmf_model.py:8)."""
    if not available():
        raise RuntimeError(f"reference not mounted at {REFERENCE_ROOT}")
    if "sharedmem" not in sys.modules:
        sys.modules["sharedmem"] = types.ModuleType("sharedmem")
    if "RecModel" not in sys.modules or not hasattr(sys.modules["RecModel"], "__path__"):
        pkg = types.ModuleType("RecModel")
        pkg.__path__ = [os.path.join(REFERENCE_ROOT, "RecModel")]
        sys.modules["RecModel"] = pkg
    base = importlib.import_module("RecModel.base_model")
    base.MKLThreads = _BlasThreads
    wmf = importlib.import_module("RecModel.wmf_model")
    wmf.MKLThreads = _BlasThreads
    return wmf.WMF
