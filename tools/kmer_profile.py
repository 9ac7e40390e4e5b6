"""Per-stage HBM figures of the k-mer stages (K0+K1 pack/keys, K2 index, K3 count, K3 fill) on synthetic reads
generated on the device:   python tools/kmer_profile.py N_READS K [READ_LEN]
Each stage is one granular library call timed alone with CUDA events (L2 flushed before every call), so the
numbers are per-stage, not per-step; algorithmic bytes follow SURVEY 8(d).  GPU box only."""
import ctypes
import importlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
PKG = "genome-assembly-using-overlap-graphs_b200"
import torch  # noqa: E402

engine = importlib.import_module(PKG + ".engine")
nat = importlib.import_module(PKG + "._native")
eng = engine.get_engine()
dev = eng.device
N = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
k = int(sys.argv[2]) if len(sys.argv) > 2 else 5
L = int(sys.argv[3]) if len(sys.argv) > 3 else 150
G = 4_600_000
HBM = 6548.8
try:
    HBM = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass

g = torch.Generator(device=dev).manual_seed(1234)
genome = torch.randint(0, 4, (G,), device=dev, generator=g, dtype=torch.int64)
starts = torch.randint(0, G, (N,), device=dev, generator=g)
lens = torch.minimum(torch.full_like(starts, L), G - starts)
offsets = torch.zeros(N + 1, dtype=torch.int64, device=dev)
offsets[1:] = torch.cumsum(lens, 0)
total = int(offsets[-1].item())
ascii_dev = torch.empty(total + 64, dtype=torch.uint8, device=dev)
letters = torch.tensor(list(b"ACGT"), dtype=torch.uint8, device=dev)
CH = 1 << 20
for r0 in range(0, N, CH):                      # chunked gather: bounded temporary memory
    r1 = min(N, r0 + CH)
    ln = lens[r0:r1]
    rid = torch.repeat_interleave(torch.arange(r1 - r0, device=dev), ln)
    within = torch.arange(int(ln.sum().item()), device=dev) - (offsets[r0:r1] - offsets[r0])[rid]
    codes = genome[starts[r0:r1][rid] + within]
    err = torch.rand(codes.shape[0], device=dev, generator=g) <= 0.005
    codes = torch.where(err, (codes + torch.randint(1, 4, codes.shape, device=dev, generator=g)) & 3, codes)
    ascii_dev[int(offsets[r0].item()):int(offsets[r1].item())] = letters[codes]
del genome, starts, rid, within, codes, err
torch.cuda.synchronize()

flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
U = N
lay = nat.CandLayout()
nat.check(nat.lib.ovl_candidates_layout(U, L, k, 1, 0, ctypes.byref(lay)))
arena = torch.empty(int(lay.total_bytes), dtype=torch.uint8, device=dev)
base = arena.data_ptr()
P = lambda off: ctypes.c_void_p(base + off)
st = lambda: ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
totals = torch.zeros(int(nat.lib.ovl_totals_len()), dtype=torch.int64).pin_memory()
ctx = eng._ctx
ev = lambda: torch.cuda.Event(enable_timing=True)


def timed(fn, reps=5):
    ms = []
    for _ in range(reps + 1):
        flush.zero_()
        torch.cuda.synchronize()
        e0, e1 = ev(), ev()
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    return min(ms[1:])


def k0():
    arena[lay.bad:lay.bad + 8].zero_()
    nat.check(nat.lib.ovl_pack_reads_keys(ctx, ctypes.c_void_p(ascii_dev.data_ptr()), ctypes.c_void_p(offsets.data_ptr()), U,
                                          lay.row_words, k, None, P(lay.packed), P(lay.len), P(lay.bad), P(lay.prefix_key),
                                          P(lay.suffix_key), st()))


def k2():
    nat.check(nat.lib.ovl_index_build(ctx, P(lay.prefix_key), P(lay.len), U, k, lay.key_bits, P(lay.sorted_key), P(lay.sorted_uid),
                                      P(lay.n_indexed), P(lay.table), lay.table_bits, None, None, None, P(lay.scratch),
                                      lay.scratch_bytes, st()))


def k3c():
    nat.check(nat.lib.ovl_join_count(ctx, P(lay.suffix_key), P(lay.prefix_key), P(lay.len), k, U, P(lay.sorted_key), P(lay.sorted_uid),
                                     P(lay.n_indexed), P(lay.table), lay.table_bits, lay.key_bits, None, None, None, None,
                                     P(lay.bucket_lo), P(lay.self_rank), P(lay.pair_off), None, P(lay.scratch), lay.scratch_bytes, st()))
    nat.check(nat.lib.ovl_join_finalize(ctx, P(lay.pair_off), None, None, None, None, None, U, P(lay.bad), P(lay.n_indexed), 0, 1,
                                        ctypes.c_void_p(totals.data_ptr()), st()))


t_k0 = timed(k0)
t_k2 = timed(k2)
t_k3c = timed(k3c)
torch.cuda.synchronize()
pairs = int(totals[0])
assert int(totals[2]) == 0
max_pairs = int(60e9 // 8)
fill_pairs = min(pairs, max_pairs)
pa = torch.empty(max(fill_pairs, 1), dtype=torch.int32, device=dev)
pb = torch.empty(max(fill_pairs, 1), dtype=torch.int32, device=dev)


def k3f():
    nat.check(nat.lib.ovl_join_fill(ctx, P(lay.pair_off), 0, U, P(lay.bucket_lo), P(lay.self_rank), P(lay.sorted_uid), 0, fill_pairs,
                                    pairs, ctypes.c_void_p(pa.data_ptr()), ctypes.c_void_p(pb.data_ptr()), st()))


t_k3f = timed(k3f, reps=3)
passes = (lay.key_bits + 9) // 10
rows = {
    "K0+K1 pack_reads_keys": (total + total / 4 + 16 * U + 12 * U, t_k0),          # ASCII + offsets in, rows + len + keys out
    f"K2 index ({passes} pass{'es' if passes > 1 else ''} + bucket table)": (12 * (1 + 2 * passes) * U + 4 * U, t_k2),
    "K3 join count + scan + totals": (16 * U + 8 * U + 16 * U, t_k3c),        # keys in, lo/self_rank out, count out+in, pair_off out
    "K3 join fill": (12 * fill_pairs, t_k3f),
}
out = {"reads": N, "read_len": L, "k": k, "pairs": pairs, "pairs_filled": fill_pairs, "hbm_peak_gbs": HBM, "stages": {}}
tot_b = tot_t = 0
for name, (b, ms) in rows.items():
    out["stages"][name] = {"algorithmic_bytes": int(b), "ms": ms, "gbs": b / ms / 1e6, "frac_of_hbm_peak": b / ms / 1e6 / HBM}
    tot_b += b
    tot_t += ms
out["sum"] = {"algorithmic_bytes": int(tot_b), "ms": tot_t, "gbs": tot_b / tot_t / 1e6, "frac_of_hbm_peak": tot_b / tot_t / 1e6 / HBM}
print(json.dumps(out))
