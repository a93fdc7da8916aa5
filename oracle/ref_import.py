"""Import the UNMODIFIED reference package from /root/reference in this container.

TEST INFRASTRUCTURE ONLY (used by tests/golden/make_golden.py, tests that pin the oracle, and
bench.py's reference arm / cpu_baseline leg).  On the GPU box /root/reference does not exist; the
unmodified copy that oracle/build_ref.sh installed under oracle/_ref/pkg/ is imported instead.

Two runtime workarounds, neither touching reference files:
  1. the Cython extension is loaded from oracle/_ref/ (built by oracle/build_ref.sh) and
     registered as tt_sketch.drm.fast_lazy_gaussian before tt_sketch.drm imports it;
  2. NumPy >= 2 rejects np.mod(int64, 2**63, dtype=uint64) at
     tt_sketch/drm/sparse_gaussian_drm.py:34-36; making DRM.seed a Python int
     (tt_sketch/drm_base.py:62) is value-identical and avoids it.
"""
import glob
import importlib.machinery
import importlib.util
import os
import sys

import numpy as np

_REF_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")
# /root/reference in the authoring container; on the GPU box the unmodified copy build_ref.sh installed
REFERENCE_ROOT = os.environ.get("TTSK_REFERENCE", "/root/reference")
if not os.path.isdir(os.path.join(REFERENCE_ROOT, "tt_sketch")):
    REFERENCE_ROOT = os.path.join(_REF_DIR, "pkg")


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "tt_sketch")) and bool(
        glob.glob(os.path.join(_REF_DIR, "fast_lazy_gaussian*.so"))
    )


def load_ref_extension():
    """Load oracle/_ref/fast_lazy_gaussian*.so on its own (works on the GPU box too)."""
    hits = glob.glob(os.path.join(_REF_DIR, "fast_lazy_gaussian*.so"))
    if not hits:
        raise ImportError("oracle/_ref extension not built (run oracle/build_ref.sh)")
    name = "fast_lazy_gaussian"
    loader = importlib.machinery.ExtensionFileLoader(name, hits[0])
    spec = importlib.util.spec_from_file_location(name, hits[0], loader=loader)
    mod = importlib.util.module_from_spec(spec)
    loader.exec_module(mod)
    return mod


def import_reference():
    """Return the reference's `tt_sketch` package (must not be mixed with the product's
    package of the same name in one process)."""
    if "tt_sketch" in sys.modules and not getattr(sys.modules["tt_sketch"], "__file__", "").startswith(REFERENCE_ROOT):
        raise RuntimeError("a different tt_sketch is already imported in this process")
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import tt_sketch  # noqa: F401  (the reference's namespace package root)
    import tt_sketch.drm_base as drm_base

    drm_base.mod = lambda a, m: int(np.mod(a, m))
    hits = glob.glob(os.path.join(_REF_DIR, "fast_lazy_gaussian*.so"))
    full = "tt_sketch.drm.fast_lazy_gaussian"
    loader = importlib.machinery.ExtensionFileLoader(full, hits[0])
    spec = importlib.util.spec_from_file_location(full, hits[0], loader=loader)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[full] = mod
    loader.exec_module(mod)
    import tt_sketch.drm  # noqa: F401
    import tt_sketch.sketch  # noqa: F401

    return sys.modules["tt_sketch"]
