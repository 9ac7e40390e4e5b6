"""ctypes binding of libovl_b200.so (C ABI declared in include/ovl.h).

There is no CPU fallback: if the shared library has not been built, importing this module
raises, and if no CUDA device is present creating a context raises.
"""
from __future__ import annotations

import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("OVL_B200_LIB", os.path.join(_HERE, "libovl_b200.so"))

OVL_OK = 0
OVL_E_CUDA = -1
OVL_E_ARG = -2
OVL_E_UNSUPPORTED = -3
OVL_MAX_K = 32
OVL_MAX_READ_LEN = 2432
OVL_MAX_LONG_READ_LEN = 16384
OVL_LOCAL_BATCH_MAX_QUERY = 1024


class OvlError(RuntimeError):
    """A call into libovl_b200.so failed (message from ovl_last_error())."""


class OvlUnsupported(OvlError, NotImplementedError):
    """Valid input for the reference that the CUDA kernels do not cover."""


class OvlBadAlphabet(OvlUnsupported):
    """The 2-bit packed kernels were handed reads with symbols other than A, C, G, T (the graph
    builder catches this and re-letters or byte-codes the read set)."""


if not os.path.isfile(LIB_PATH):
    raise ImportError(
        f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
        "(nvcc, sm_100a). The overlap path has no CPU fallback.")

lib = ctypes.CDLL(LIB_PATH)

_vp = ctypes.c_void_p
_i32 = ctypes.c_int32
_i64 = ctypes.c_int64
_sz = ctypes.c_size_t

class CandLayout(ctypes.Structure):
    """ovl_cand_layout (include/ovl.h): byte offsets of every array of the K0-K3 job inside its arena."""
    _fields_ = [(n, ctypes.c_size_t) for n in (
        "packed", "len", "bad", "n_indexed", "prefix_key", "suffix_key", "sorted_key", "sorted_uid", "table", "pos_of",
        "bucket_lo", "self_rank", "pair_off", "edge_base", "cum", "sorted_copies", "scratch", "scratch_bytes", "total_bytes")] + \
        [(n, ctypes.c_int32) for n in ("row_words", "key_bits", "table_bits", "has_copies")]


_lay_p = ctypes.POINTER(CandLayout)

_SIGS = {
    "ovl_last_error": (ctypes.c_char_p, []),
    "ovl_version": (ctypes.c_int, []),
    "ovl_ctx_create": (ctypes.c_int, [ctypes.c_int, ctypes.POINTER(_vp)]),
    "ovl_ctx_destroy": (ctypes.c_int, [_vp]),
    "ovl_ctx_sm_count": (ctypes.c_int, [_vp]),
    "ovl_ctx_launch_count": (_i64, [_vp]),
    "ovl_row_words": (_i32, [_i32]),
    "ovl_pack_reads": (ctypes.c_int, [_vp, _vp, _vp, _i64, _i32, _vp, _vp, _vp, _vp]),
    "ovl_kmer_keys": (ctypes.c_int, [_vp, _vp, _i32, _vp, _i64, _i32, _vp, _vp, _vp, _vp]),
    "ovl_kmer_hashes": (ctypes.c_int, [_vp, _vp, _i32, _vp, _i64, _i32, _vp, _vp, _vp]),
    "ovl_join_count_verify": (ctypes.c_int, [_vp, _vp, _i32, _vp, _i32, _vp, _i64, _i64, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "ovl_join_fill_verify": (ctypes.c_int, [_vp, _vp, _i32, _vp, _i32, _vp, _i64, _i64, _vp, _vp, _vp, _vp, _i64, _i64,
                                            _vp, _vp, _vp]),
    "ovl_index_workspace_bytes": (_sz, [_i64]),
    "ovl_pack_reads_keys": (ctypes.c_int, [_vp, _vp, _vp, _i64, _i32, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "ovl_index_table_bits": (_i32, [_i64, _i32]),
    "ovl_index_build": (ctypes.c_int, [_vp, _vp, _vp, _i64, _i32, _i32, _vp, _vp, _vp, _vp, _i32, _vp, _vp, _vp, _vp, _sz, _vp]),
    "ovl_join_workspace_bytes": (_sz, [_i64]),
    "ovl_join_count": (ctypes.c_int, [_vp, _vp, _vp, _vp, _i32, _i64, _vp, _vp, _vp, _vp, _i32, _i32, _vp, _vp, _vp, _vp, _vp, _vp,
                                      _vp, _vp, _vp, _sz, _vp]),
    "ovl_totals_len": (_i32, []),
    "ovl_join_finalize": (ctypes.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _vp, _vp, _i32, _i32, _vp, _vp]),
    "ovl_candidates_layout": (ctypes.c_int, [_i64, _i32, _i32, _i32, _i32, _lay_p]),
    "ovl_candidates_build": (ctypes.c_int, [_vp, _vp, _vp, _i64, _i32, _vp, _vp, _i32, _i32, _vp, _lay_p, _vp, _vp]),
    "ovl_overlap_dp_edges_join": (ctypes.c_int, [_vp, _vp, _i32, _vp, _vp, _vp, _i64, _i32, _i64, _i64, _i64, _vp, _vp, _vp,
                                                 _vp, _vp, _vp, _vp, _i64, _i64, _vp, _vp]),
    "ovl_join_fill": (ctypes.c_int, [_vp, _vp, _i64, _i64, _vp, _vp, _vp, _i64, _i64, _i64, _vp, _vp, _vp]),
    "ovl_all_pairs_fill": (ctypes.c_int, [_vp, _i64, _i64, _i64, _i64, _vp, _vp, _vp]),
    "ovl_overlap_dp": (ctypes.c_int, [_vp, _vp, _i32, _vp, _vp, _vp, _i64, _i32, _i64, _i64, _i64, _vp, _vp,
                                      _i32, _i32, _i32, _vp]),
    "ovl_overlap_dp_edges": (ctypes.c_int, [_vp, _vp, _i32, _vp, _vp, _vp, _i64, _i32, _i64, _i64, _i64, _vp, _vp, _vp,
                                            _vp, _vp]),
    "ovl_overlap_dp_plan": (ctypes.c_int, [_i32, _i64, _i64, _i64, _i32, ctypes.POINTER(_i32 * 3)]),
    "ovl_pack_bytes": (ctypes.c_int, [_vp, _vp, _vp, _i64, _i32, _vp, _vp, _vp]),
    "ovl_kmer_hashes8": (ctypes.c_int, [_vp, _vp, _i32, _vp, _i64, _i32, _vp, _vp, _vp]),
    "ovl_join_count_verify8": (ctypes.c_int, [_vp, _vp, _i32, _vp, _i32, _vp, _i64, _i64, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "ovl_join_fill_verify8": (ctypes.c_int, [_vp, _vp, _i32, _vp, _i32, _vp, _i64, _i64, _vp, _vp, _vp, _vp, _i64, _i64,
                                             _vp, _vp, _vp]),
    "ovl_overlap_dp8": (ctypes.c_int, [_vp, _vp, _i32, _vp, _vp, _vp, _i64, _i32, _i64, _i64, _i64, _vp, _vp, _vp, _vp,
                                       _vp, _vp, _vp]),
    "ovl_expand_workspace_bytes": (_sz, [_i64]),
    "ovl_expand_count": (ctypes.c_int, [_vp, _vp, _vp, _vp, _i64, _vp, _vp, _sz, _vp]),
    "ovl_expand_fill": (ctypes.c_int, [_vp, _vp, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _i64, _vp, _vp]),
    "ovl_expand_unit": (ctypes.c_int, [_vp, _vp, _vp, _vp, _vp, _i64, _vp, _vp]),
    "ovl_filter_workspace_bytes": (_sz, [_i64]),
    "ovl_filter_count": (ctypes.c_int, [_vp, _vp, _i64, _i32, _vp, _vp, _sz, _vp]),
    "ovl_filter_fill": (ctypes.c_int, [_vp, _vp, _vp, _i64, _i32, _vp, _vp]),
    "ovl_align_pair_workspace_bytes": (_sz, [_i32, _i32]),
    "ovl_align_pair": (ctypes.c_int, [_vp, _vp, _i32, _vp, _i32, _i64, _i64, _i64, _vp, _sz, _vp, _vp, _vp]),
    "ovl_local_align_workspace_bytes": (_sz, [_i32, _i32]),
    "ovl_local_align": (ctypes.c_int, [_vp, _vp, _i32, _vp, _i32, _i64, _i64, _i64, _vp, _sz, _vp, _vp, _vp]),
    "ovl_local_align_batch": (ctypes.c_int, [_vp, _vp, _vp, _i32, _i32, _vp, _vp, _vp, _i64, _i64, _i64, _vp, _vp, _vp, _vp,
                                             _vp, _vp]),
    "ovl_simulate_workspace_bytes": (_sz, [_i64]),
    "ovl_simulate_reads": (ctypes.c_int, [_vp, _vp, _i64, _i64, _i32, ctypes.c_uint32, ctypes.c_uint64, _vp, _vp, _vp, _sz, _vp]),
    "ovl_trim_workspace_bytes": (_sz, [_i64]),
    "ovl_trim_sinks": (ctypes.c_int, [_vp, _vp, _vp, _i64, _i64, _vp, _vp, _sz, ctypes.POINTER(_i32), _vp]),
    "ovl_edge_list_hash": (ctypes.c_int, [_vp, _vp, _i64, _i64, _vp, _vp]),
    "ovl_int_peak_probe": (ctypes.c_int, [_vp, _i32, _i32, ctypes.POINTER(ctypes.c_double),
                                          ctypes.POINTER(ctypes.c_double)]),
}

EXPORTS = tuple(_SIGS)

for _name, (_res, _args) in _SIGS.items():
    _fn = getattr(lib, _name)      # AttributeError here == the .so does not match include/ovl.h
    _fn.restype = _res
    _fn.argtypes = _args


def last_error() -> str:
    return lib.ovl_last_error().decode("utf-8", "replace")


def check(rc: int) -> None:
    if rc == OVL_OK:
        return
    msg = last_error()
    if rc == OVL_E_UNSUPPORTED:
        raise OvlUnsupported(msg)
    if rc == OVL_E_ARG:
        raise ValueError(msg)
    raise OvlError(msg)
