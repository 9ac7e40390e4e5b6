"""Drop-in for the reference's ``aligners.overlap_alignment`` (aligners.py:6-82).

Same name, same keyword arguments, same 5-tuple; the DP, the last-row arg-max and the
traceback run on the GPU (kernel K7, csrc/align.cuh) -- there is no CPU fallback.
Every other symbol of the reference's ``aligners`` module (local_alignment, ...) is outside
the accelerated path and is forwarded to the reference module when ``OVL_REFERENCE_DIR`` names
its checkout (opt-in).
"""
from __future__ import annotations

import numpy as np

try:
    from . import engine as _engine
except ImportError:      # imported as a top-level module with the package directory on sys.path
    import importlib as _il
    import os as _os
    import sys as _sys
    _pkg_dir = _os.path.dirname(_os.path.abspath(__file__))
    if _os.path.dirname(_pkg_dir) not in _sys.path:
        _sys.path.append(_os.path.dirname(_pkg_dir))
    _engine = _il.import_module(_os.path.basename(_pkg_dir) + ".engine")

INDEL_DEFAULT = -2 ** 31


def _codes(x: str) -> np.ndarray:
    return np.frombuffer(x.encode("utf-32-le"), dtype=np.int32) if x else np.zeros(0, np.int32)


def overlap_alignment(s, t, match_score=10, mismatch=-1, indel=-2 ** 31):
    """Best suffix(s)/prefix(t) overlap alignment -- see aligners.py:6-26 for the contract.

    Returns (alignment_to_print, align_s, align_t, best_score, alignment_end_position) with
    ``best_score`` and ``alignment_end_position`` as Python ints.
    """
    if not isinstance(s, str) or not isinstance(t, str):
        # the reference is @njit-compiled for unicode arguments and raises a TypingError here
        raise TypeError("overlap_alignment expects str arguments (the reference raises a Numba TypingError)")
    for name, v in (("match_score", match_score), ("mismatch", mismatch), ("indel", indel)):
        if isinstance(v, bool) or not isinstance(v, (int, np.integer)):
            raise TypeError(f"{name} must be an integer")
        if not -2 ** 63 <= int(v) < 2 ** 63:
            raise OverflowError(f"{name} does not fit int64")
    eng = _engine.get_engine()
    score, end, ops = eng.align_pair(_codes(s), _codes(t), int(match_score), int(mismatch), int(indel))
    # rebuild the aligned strings from the device traceback (aligners.py:63-76)
    i, j = len(s), end
    a_s, a_t = [], []
    for op in ops.tolist():
        if op == 0:
            a_s.append(s[i - 1]); a_t.append(t[j - 1]); i -= 1; j -= 1
        elif op == 1:
            a_s.append(s[i - 1]); a_t.append("-"); i -= 1
        else:
            a_s.append("-"); a_t.append(t[j - 1]); j -= 1
    align_s = "".join(reversed(a_s))
    align_t = "".join(reversed(a_t))
    alignment_to_print = f"\nTarget:   {align_t}\n          {'|' * len(align_t)}\nQuery:    {align_s}"   # aligners.py:78
    return alignment_to_print, align_s, align_t, int(score), int(end)


def _local_tuple(query, reference, score, start, end, best_i, ops):
    """Device result (scores + op list from the best cell backwards) -> the reference's 6-tuple."""
    i, j = best_i, end
    a_q, a_r = [], []
    for op in ops.tolist():                                           # aligners.py:143-160
        if op == 1:
            a_q.append(query[i - 1]); a_r.append(reference[j - 1]); i -= 1; j -= 1
        elif op == 2:
            a_q.append(query[i - 1]); a_r.append("-"); i -= 1
        else:
            a_q.append("-"); a_r.append(reference[j - 1]); j -= 1
    aligned_query = "".join(reversed(a_q))
    aligned_reference = "".join(reversed(a_r))
    alignment_to_print = (f"\nTarget:   {aligned_reference}\n          {'|' * len(aligned_reference)}\nQuery:    "
                          f"{aligned_query}")                          # aligners.py:166-167
    return alignment_to_print, aligned_reference, aligned_query, int(score), int(start), int(end)


def _check_local_args(query, reference, match_score, mismatch, indel):
    if not isinstance(query, str) or not isinstance(reference, str):
        raise TypeError("local_alignment expects str arguments (the reference raises a Numba TypingError)")
    for name, v in (("match_score", match_score), ("mismatch", mismatch), ("indel", indel)):
        if isinstance(v, bool) or not isinstance(v, (int, np.integer)):
            raise TypeError(f"{name} must be an integer")


def local_alignment(query, reference, match_score=10, mismatch=-1, indel=-1):
    """Best local alignment of `query` in `reference` -- drop-in for aligners.py:85-167.

    Returns (alignment_to_print, aligned_reference, aligned_query, best_score, start_pos, end_pos).
    """
    _check_local_args(query, reference, match_score, mismatch, indel)
    eng = _engine.get_engine()
    score, start, end, best_i, ops = eng.local_align(_codes(query), _codes(reference), int(match_score),
                                                     int(mismatch), int(indel))
    return _local_tuple(query, reference, score, start, end, best_i, ops)


def align_read_or_contig_to_reference(read_or_contig, reference_genome, read_length, match_score=10, mismatch=-1,
                                      indel=-1):
    """Drop-in for aligners.py:170-202: local alignment against the genome, or -- for a sequence
    shorter than a read -- against the genome's last len(sequence) bases.  A result computed ahead by
    prefetch_alignments() for exactly these arguments is returned without touching the GPU again."""
    hit = _PREFETCHED.get((read_or_contig, id(reference_genome), len(reference_genome), read_length, match_score, mismatch, indel))
    if hit is not None:
        return hit
    n = len(read_or_contig)
    if n < read_length:
        to_print, aligned_ref, aligned, score, start, end = local_alignment(
            read_or_contig, reference_genome[-n:], match_score, mismatch, indel)
        start = len(reference_genome) - n + start
        end = len(reference_genome) - n + end
    else:
        to_print, aligned_ref, aligned, score, start, end = local_alignment(
            read_or_contig, reference_genome, match_score, mismatch, indel)
    return to_print, aligned_ref, aligned, score, start, end


def align_reads_or_contigs_to_reference(sequences, reference_genome, read_length, match_score=10, mismatch=-1, indel=-1):
    """[align_read_or_contig_to_reference(s, reference_genome, read_length, ...) for s in sequences] from ONE
    kernel launch: one CTA per sequence, the genome uploaded once (not in the reference, which aligns the
    contigs of an assembly one call at a time, performanceMeasures.py:219-221).  Sequences longer than
    1,024 symbols take the per-call path."""
    sequences = list(sequences)
    for q in sequences:
        _check_local_args(q, reference_genome, match_score, mismatch, indel)
    eng = _engine.get_engine()
    G = len(reference_genome)
    out = [None] * len(sequences)
    batch = [x for x, q in enumerate(sequences) if len(q) <= _engine.nat.OVL_LOCAL_BATCH_MAX_QUERY]
    windows = []
    for x in batch:
        n = len(sequences[x])
        windows.append((G - n, n) if n < read_length else (0, G))      # aligners.py:191-199
    if batch:
        res = eng.local_align_batch([_codes(sequences[x]) for x in batch], _codes(reference_genome), windows,
                                    int(match_score), int(mismatch), int(indel))
        for x, (w0, wl), (score, start, end, best_i, ops) in zip(batch, windows, res):
            q = sequences[x]
            ref = reference_genome if w0 == 0 and wl == G else reference_genome[w0:w0 + wl]
            to_print, aligned_ref, aligned, score, start, end = _local_tuple(q, ref, score, start, end, best_i, ops)
            out[x] = (to_print, aligned_ref, aligned, score, w0 + start, w0 + end)
    for x, q in enumerate(sequences):
        if out[x] is None:
            out[x] = align_read_or_contig_to_reference(q, reference_genome, read_length, match_score, mismatch, indel)
    return out


def local_alignment_batch(queries, reference, match_score=10, mismatch=-1, indel=-1):
    """[local_alignment(q, reference, ...) for q in queries] from one kernel launch."""
    return align_reads_or_contigs_to_reference(queries, reference, 0, match_score, mismatch, indel)


_PREFETCHED = {}


def prefetch_alignments(sequences, reference_genome, read_length, match_score=10, mismatch=-1, indel=-1):
    """Compute align_read_or_contig_to_reference for all `sequences` in one launch and keep the results, so that
    the reference's own per-contig loop (performanceMeasures.py:219-221) is served from them: call this once
    before the loop with the same genome object.  The cache holds the results of the latest call only."""
    _PREFETCHED.clear()
    seqs = list(dict.fromkeys(sequences))
    res = align_reads_or_contigs_to_reference(seqs, reference_genome, read_length, match_score, mismatch, indel)
    for q, r in zip(seqs, res):
        _PREFETCHED[(q, id(reference_genome), len(reference_genome), read_length, match_score, mismatch, indel)] = r
    return len(seqs)


def __getattr__(name):
    if name.startswith("__") and name.endswith("__"):
        raise AttributeError(name)
    ref = _engine.reference_module("aligners")
    if ref is not None and hasattr(ref, name):
        return getattr(ref, name)
    raise AttributeError(f"module 'aligners' (B200 drop-in) has no attribute {name!r}")
