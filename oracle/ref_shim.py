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


def load_reference_ease(build_dir="/tmp/wmf_ref_build"):
    """Returns the reference's Ease class (RecModel/ease_model.py:45) with its Cython predictor
    (RecModel/fast_utils/ease_utils.pyx) compiled from the sources where they lie, into ``build_dir``
    (outside the repository; nothing of the reference is copied into the tree)."""
    import shutil
    import subprocess
    load_reference_wmf()   # package stubs
    os.makedirs(build_dir, exist_ok=True)
    so = [f for f in os.listdir(build_dir) if f.startswith("ease_utils") and f.endswith(".so")]
    if not so:
        shutil.copy(os.path.join(REFERENCE_ROOT, "RecModel", "fast_utils", "ease_utils.pyx"), build_dir)
        setup = ("from setuptools import setup, Extension\nfrom Cython.Build import cythonize\nimport numpy\n"
                 "setup(ext_modules=cythonize([Extension('ease_utils', ['ease_utils.pyx'], include_dirs=[numpy.get_include()])], "
                 "compiler_directives={'legacy_implicit_noexcept': True}))\n")
        with open(os.path.join(build_dir, "setup.py"), "w") as fh:
            fh.write(setup)
        env = dict(os.environ, CC="/usr/bin/gcc")
        subprocess.run([sys.executable, "setup.py", "build_ext", "--inplace"], cwd=build_dir, env=env, check=True,
                       capture_output=True)
    if build_dir not in sys.path:
        sys.path.insert(0, build_dir)
    ease_utils = importlib.import_module("ease_utils")
    fast = types.ModuleType("RecModel.fast_utils")
    fast.__path__ = []
    fast.ease_utils = ease_utils
    sys.modules["RecModel.fast_utils"] = fast
    sys.modules["RecModel.fast_utils.ease_utils"] = ease_utils
    sm = sys.modules["sharedmem"]
    if not hasattr(sm, "empty"):
        import numpy as _np
        sm.empty = lambda shape, dtype=_np.float32: _np.empty(shape, dtype=dtype)   # ease_model.py:110
    ease = importlib.import_module("RecModel.ease_model")
    return ease.Ease
