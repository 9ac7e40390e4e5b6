"""B200-native overlap detection: drop-in for the reference's
``aligners.overlap_alignment`` (aligners.py:6-82) and
``overlapGraphs.construct_overlap_graph_nx_k`` (overlapGraphs.py:5-61).

Host code is Python; the work runs in hand-written sm_100a CUDA kernels reached through the
C ABI of ``libovl_b200.so`` (``include/ovl.h``) with ctypes.  PyTorch only owns device
buffers and streams.  There is no CPU fallback.

Typical use::

    import importlib
    ovl = importlib.import_module("genome-assembly-using-overlap-graphs_b200")
    G, read_copies = ovl.construct_overlap_graph_nx_k(reads, k=5)

or put this directory first on ``sys.path`` and ``import overlapGraphs`` / ``import aligners``
exactly as the reference's callers do (testAssembly.py:3, overlapGraphs.py:2).
"""


def __getattr__(name):
    # lazy: importing the package must not need a GPU or the built library
    if name in ("overlap_alignment",):
        from .aligners import overlap_alignment
        return overlap_alignment
    if name in ("construct_overlap_graph_nx_k", "construct_overlap_graph_string", "construct_string_graph",
                "construct_overlap_graphs_batch"):
        from . import overlapGraphs
        return getattr(overlapGraphs, name)
    if name in ("OverlapEngine", "get_engine"):
        from . import engine
        return getattr(engine, name)
    raise AttributeError(name)
