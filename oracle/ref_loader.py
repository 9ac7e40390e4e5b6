"""Loader for the live, unmodified reference -- TEST INFRASTRUCTURE ONLY.

Works only where the reference checkout exists (the build container,
``/root/reference``); on the GPU box it does not, and everything that needs the live
reference is skipped there (the committed fixtures under ``tests/golden/`` carry its
outputs instead).  ``Bio`` and ``matplotlib`` are not installed and are not on the hot
path (aligners.py:3 is only used by local_alignment_biopython, aligners.py:225), so they
are stubbed.
"""
from __future__ import annotations

import os
import sys
from unittest.mock import MagicMock

REFERENCE_DIR = os.environ.get("OVL_REFERENCE_DIR", "/root/reference")
_mods = None


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_DIR, "aligners.py"))


def load():
    """Return (aligners, overlapGraphs) modules of the unmodified reference, JIT warmed."""
    global _mods
    if _mods is None:
        if not available():
            raise RuntimeError(f"reference checkout not found at {REFERENCE_DIR}")
        for m in ("Bio", "Bio.Align", "matplotlib", "matplotlib.pyplot"):
            sys.modules.setdefault(m, MagicMock())
        saved = {k: sys.modules.pop(k) for k in ("aligners", "overlapGraphs") if k in sys.modules}
        sys.path.insert(0, REFERENCE_DIR)
        try:
            import aligners as ref_aligners          # noqa: E402
            import overlapGraphs as ref_graphs       # noqa: E402
        finally:
            sys.path.remove(REFERENCE_DIR)
            # do not leave the reference registered under the drop-in's module names
            sys.modules.pop("aligners", None)
            sys.modules.pop("overlapGraphs", None)
            sys.modules.update(saved)
        assert os.path.dirname(os.path.abspath(ref_aligners.__file__)) == os.path.abspath(REFERENCE_DIR)
        ref_aligners.overlap_alignment("ACGT", "ACGT")   # warm the Numba JIT (~8 s)
        _mods = (ref_aligners, ref_graphs)
    return _mods
