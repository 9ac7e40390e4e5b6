"""Seeded synthetic read generator (bench / test input only, NumPy on the host).

Distribution-equivalent to the reference's read simulator:
  * generateErrorFreeReads.py:38-50 -- uniform start on a LINEAR genome, reads truncated
    at the genome end (no wrap), so lengths are 1..l;
  * generateErrorProneReads.py:17-28 -- every base independently replaced, with
    probability p, by one of the three other bases.
The reference's own generators are un-seeded (Python ``random`` / Numba RNG); this one is
seeded (PCG64) so the GPU path, the oracle and the fixtures all see the same reads.
"""
from __future__ import annotations

import os
from typing import List, Tuple

import numpy as np

_ALPHABET = np.frombuffer(b"ACGT", dtype=np.uint8)

# PhiX174 NC_001422.1 is 5,386 bp with composition A:1291 C:1157 G:1254 T:1684
# (sequence.fasta in the reference).  The GPU box has no copy of the reference, so the
# "PhiX-like" genome below is a seeded random genome of the same length and composition.
PHIX_LEN = 5386
PHIX_COMPOSITION = (1291, 1157, 1254, 1684)


def random_genome(length: int, seed: int) -> np.ndarray:
    rng = np.random.Generator(np.random.PCG64(seed))
    return _ALPHABET[rng.integers(0, 4, size=length)]


def phix_like_genome(seed: int = 174) -> np.ndarray:
    rng = np.random.Generator(np.random.PCG64(seed))
    g = np.repeat(_ALPHABET, PHIX_COMPOSITION)
    rng.shuffle(g)
    return g


def read_fasta(path: str) -> np.ndarray:
    """Same parsing rule as generateErrorFreeReads.py:4-19 (skip '>' lines, join the rest)."""
    seq = []
    with open(path) as fh:
        for line in fh:
            if not line.startswith(">"):
                seq.append(line.strip())
    return np.frombuffer("".join(seq).encode("ascii"), dtype=np.uint8)


def simulate_reads(genome: np.ndarray, n_reads: int, read_len: int, error_prob: float,
                   seed: int) -> Tuple[np.ndarray, np.ndarray]:
    """Return (bases uint8[sum len], offsets int64[n_reads+1]) of error-prone reads."""
    rng = np.random.Generator(np.random.PCG64(seed))
    G = int(genome.shape[0])
    starts = rng.integers(0, G, size=n_reads)
    lens = np.minimum(read_len, G - starts).astype(np.int64)
    offsets = np.zeros(n_reads + 1, dtype=np.int64)
    np.cumsum(lens, out=offsets[1:])
    total = int(offsets[-1])
    # gather genome[start + i] for every base of every read
    read_id = np.repeat(np.arange(n_reads, dtype=np.int64), lens)
    within = np.arange(total, dtype=np.int64) - offsets[read_id]
    bases = genome[starts[read_id] + within].copy()
    if error_prob > 0:
        hit = rng.random(total) <= error_prob
        n_hit = int(hit.sum())
        if n_hit:
            code = np.zeros(256, dtype=np.uint8)
            code[_ALPHABET] = np.arange(4, dtype=np.uint8)
            old = code[bases[hit]]
            new = (old + rng.integers(1, 4, size=n_hit).astype(np.uint8)) & 3   # one of the 3 others
            bases[hit] = _ALPHABET[new]
    return bases, offsets


_M64 = (1 << 64) - 1


def _sim_hash(seed: int, stream: int, index: np.ndarray) -> np.ndarray:
    """splitmix64 finaliser over (seed, stream, index) -- the mirror of csrc/simulate.cuh:sim_hash."""
    with np.errstate(over="ignore"):
        z = (np.uint64((seed + 0x9E3779B97F4A7C15 * (stream + 1)) & _M64)
             + np.uint64(0xD1B54A32D192ED03) * index.astype(np.uint64))
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))


def error_threshold(error_prob: float) -> int:
    """P(error) as the 32-bit integer threshold both the device kernel and the NumPy mirror compare with."""
    return max(0, min(0xFFFFFFFF, int(error_prob * 4294967296.0)))


def simulate_reads_counter(genome: np.ndarray, n_reads: int, read_len: int, error_prob: float,
                           seed: int) -> Tuple[np.ndarray, np.ndarray]:
    """The same read model as simulate_reads() on a counter-based stream: every start and every base is an
    independent function of (seed, read, position), so the device kernel ovl_simulate_reads produces these
    exact bytes (OverlapEngine.simulate_reads).  Returns (bases uint8[sum len], offsets int64[n_reads+1])."""
    G = int(genome.shape[0])
    idx = np.arange(n_reads, dtype=np.uint64)
    h = _sim_hash(seed, 0, idx)
    starts = (((h >> np.uint64(32)) * np.uint64(G)) >> np.uint64(32)).astype(np.int64)
    lens = np.minimum(read_len, G - starts).astype(np.int64)
    offsets = np.zeros(n_reads + 1, dtype=np.int64)
    np.cumsum(lens, out=offsets[1:])
    total = int(offsets[-1])
    read_id = np.repeat(np.arange(n_reads, dtype=np.int64), lens)
    within = np.arange(total, dtype=np.int64) - offsets[read_id]
    bases = genome[starts[read_id] + within].copy()
    hb = _sim_hash(seed, 1, (read_id * read_len + within).astype(np.uint64))
    hit = (hb >> np.uint64(32)) < np.uint64(error_threshold(error_prob))
    if hit.any():
        code = np.zeros(256, dtype=np.uint32)
        code[_ALPHABET] = np.arange(4, dtype=np.uint32)
        shift = 1 + (((hb[hit] & np.uint64(0xFFFFFFFF)) * np.uint64(3)) >> np.uint64(32)).astype(np.uint32)
        bases[hit] = _ALPHABET[(code[bases[hit]] + shift) & 3]
    return bases, offsets


def to_strings(bases: np.ndarray, offsets: np.ndarray) -> List[str]:
    buf = bases.tobytes().decode("ascii")
    off = offsets.tolist()
    return [buf[off[i]:off[i + 1]] for i in range(len(off) - 1)]


def dedup(bases: np.ndarray, offsets: np.ndarray):
    """Host-side read de-duplication in first-appearance order (overlapGraphs.py:18-20)
    on the flat representation.  Returns (uniq_bases, uniq_offsets, counts int32[U],
    read_to_uid int32[N])."""
    n = len(offsets) - 1
    buf = bases.tobytes()
    off = offsets.tolist()
    first = {}
    read_to_uid = np.empty(n, dtype=np.int32)
    counts: List[int] = []
    keep: List[int] = []
    for i in range(n):
        r = buf[off[i]:off[i + 1]]
        u = first.get(r)
        if u is None:
            u = len(keep)
            first[r] = u
            keep.append(i)
            counts.append(1)
        else:
            counts[u] += 1
        read_to_uid[i] = u
    keep_a = np.asarray(keep, dtype=np.int64)
    lens = (offsets[1:] - offsets[:-1])[keep_a]
    uo = np.zeros(len(keep) + 1, dtype=np.int64)
    np.cumsum(lens, out=uo[1:])
    rid = np.repeat(np.arange(len(keep), dtype=np.int64), lens)
    within = np.arange(int(uo[-1]), dtype=np.int64) - uo[rid]
    ub = bases[offsets[keep_a][rid] + within] if len(keep) else np.zeros(0, np.uint8)
    return ub, uo, np.asarray(counts, dtype=np.int32), read_to_uid


WORKLOADS = {
    # name: (genome kind, genome length, N reads, read length l, error prob p)
    "phix_n1000_l100": ("phix", PHIX_LEN, 1000, 100, 0.01),         # BASELINE.json configs[0]
    "phix_n50000_l150": ("phix", PHIX_LEN, 50000, 150, 0.01),       # configs[1]
    "ecoli_n1m_l150": ("random", 4_600_000, 1_000_000, 150, 0.005),  # configs[2]
    "ecoli_n200k_l1000": ("random", 4_600_000, 200_000, 1000, 0.02),  # configs[3]
}


def make_workload(name: str, seed: int = 12345, fasta: str | None = None):
    kind, glen, n, l, p = WORKLOADS[name]
    if kind == "phix":
        if fasta and os.path.isfile(fasta):
            genome = read_fasta(fasta)
        else:
            genome = phix_like_genome()
    else:
        genome = random_genome(glen, seed ^ 0x5EED)
    return simulate_reads(genome, n, l, p, seed)
