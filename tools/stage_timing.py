"""Where does the time of the non-DP stages go?  CPU issue time vs GPU time (GPU box only)."""
import importlib, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
PKG = "genome-assembly-using-overlap-graphs_b200"
import torch
synth = importlib.import_module(PKG + ".synth")
engine = importlib.import_module(PKG + ".engine")
eng = engine.get_engine()
wl = sys.argv[1] if len(sys.argv) > 1 else "phix_n50000_l150"
k = int(sys.argv[2]) if len(sys.argv) > 2 else 5
bases, offsets = synth.make_workload(wl)
ub, uo, counts, _ = synth.dedup(bases, offsets)
U = len(counts); total = int(uo[-1]); max_len = int((uo[1:] - uo[:-1]).max())
d_ascii = torch.empty(total + 64, dtype=torch.uint8, device=eng.device); d_ascii[:total].copy_(torch.from_numpy(ub[:total]))
d_off = torch.from_numpy(uo).to(eng.device)
ev = lambda: torch.cuda.Event(enable_timing=True)
for it in range(6):
    torch.cuda.synchronize()
    e = [ev() for _ in range(6)]
    t0 = time.perf_counter(); e[0].record()
    rs = eng.pack_reads(d_ascii, d_off, U, max_len); e[1].record()
    t1 = time.perf_counter()
    idx = eng.kmer_index(rs, k); e[2].record()
    t2 = time.perf_counter()
    pa, pb, _ = eng.candidate_pairs(rs, idx, k); e[3].record()
    t3 = time.perf_counter()
    torch.cuda.synchronize(); t4 = time.perf_counter()
    if it >= 3:
        print(f"cpu us: pack {1e6*(t1-t0):.0f} index {1e6*(t2-t1):.0f} join(incl sync) {1e6*(t3-t2):.0f} tail-sync {1e6*(t4-t3):.0f} | "
              f"gpu us: pack {1e3*e[0].elapsed_time(e[1]):.0f} index {1e3*e[1].elapsed_time(e[2]):.0f} join {1e3*e[2].elapsed_time(e[3]):.0f} | pairs {pa.shape[0]}")
