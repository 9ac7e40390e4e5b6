"""Time the overlap DP kernel alone for every (lanes, columns) instantiation that fits a
workload, plus the integer-pipe probes.  GPU box only:

    OVL_B200_LIB=build/variants/libovl_minb3.so python tools/dp_sweep.py --workload phix_n50000_l150
"""
import argparse
import ctypes
import importlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
PKG = "genome-assembly-using-overlap-graphs_b200"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="phix_n50000_l150")
    ap.add_argument("--k", type=int, default=5)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--max-pairs", type=int, default=4_000_000)
    ap.add_argument("--indel", type=int, default=-2 ** 31)
    ap.add_argument("--modes", default="1,2")
    ap.add_argument("--only", default="", help="LANESxCOLS, e.g. 4x38")
    ap.add_argument("--no-probe", action="store_true")
    ap.add_argument("--probe-only", action="store_true")
    args = ap.parse_args()
    import torch
    synth = importlib.import_module(PKG + ".synth")
    engine = importlib.import_module(PKG + ".engine")
    nat = importlib.import_module(PKG + "._native")
    eng = engine.get_engine()
    bases, offsets = synth.make_workload(args.workload)
    ub, uo, counts, _ = synth.dedup(bases, offsets)
    rs = eng.upload_reads(ub, uo)
    idx = eng.kmer_index(rs, args.k)
    pa, pb, _ = eng.candidate_pairs(rs, idx, args.k)
    if pa.shape[0] > args.max_pairs:
        pa, pb = pa[:args.max_pairs].contiguous(), pb[:args.max_pairs].contiguous()
    P = int(pa.shape[0])
    lens = rs.length[:rs.n_reads].to(torch.int64)
    cells = int((lens[pa.long()] * lens[pb.long()]).sum().item())
    print(json.dumps({"lib": nat.LIB_PATH, "workload": args.workload, "pairs": P, "cells": cells,
                      "max_len": rs.max_len, "plan": eng.dp_plan(rs.max_len, 10, -1, args.indel)}), flush=True)
    ref = None
    for mode in ([] if args.probe_only else [int(x) for x in args.modes.split(",")]):
        for lanes in (1, 2, 4, 8, 16, 32):
            for cols in ((19, 25, 32, 38, 76) if mode == 1 else (32,)):
                if lanes * cols < rs.max_len or lanes * cols > 4 * max(rs.max_len, 38):
                    continue
                if args.only and f"{lanes}x{cols}" not in args.only.split(","):
                    continue
                try:
                    s, e = eng.overlap_scores(rs, pa, pb, 10, -1, args.indel, mode=mode, lanes=lanes, cols=cols)
                except nat.OvlError as exc:
                    print(json.dumps({"mode": mode, "lanes": lanes, "cols": cols, "error": str(exc)}))
                    continue
                torch.cuda.synchronize()
                if ref is None:
                    ref = (s.clone(), e.clone())
                ok = bool(torch.equal(s, ref[0]) and torch.equal(e, ref[1]))
                best = 1e30
                for _ in range(args.reps):
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    eng.overlap_scores(rs, pa, pb, 10, -1, args.indel, mode=mode, lanes=lanes, cols=cols, out=(s, e))
                    e1.record()
                    torch.cuda.synchronize()
                    best = min(best, e0.elapsed_time(e1))
                print(json.dumps({"mode": mode, "lanes": lanes, "cols": cols, "ms": round(best, 3),
                                  "gcups": round(cells / best / 1e6, 1), "same_as_first": ok}), flush=True)
    if args.no_probe:
        return
    names = {0: "iadd3", 1: "imad", 2: "vimnmx_s32", 3: "viaddmnmx_s16x2", 4: "dp_mix", 5: "prmt", 6: "lop3",
                 7: "lop3_imad_pair", 8: "vimnmx3_imad_distinct_regs", 9: "dp_form1_column", 10: "dp_form2_column", 11: "alu2_fma2_column", 12: "prmt+viaddmnmx", 13: "viaddmnmx+imad", 14: "prmt+imad",
                 15: "2viaddmnmx+imad", 16: "form1_iadd_on_alu", 17: "vminu2_as_vimnmx3_u16x2", 18: "vimnmx3_u16x2",
                 20: "viaddmnmx_u16x2", 21: "viaddmnmx_u16x2_imm", 22: "dp_form1_column_imm_gaps", 23: "imad_imm",
                 24: "prmt_2regs", 25: "lop3_imm", 26: "hmnmx2", 27: "hmnmx2+lop3", 28: "hmnmx2+imad",
                 29: "hmnmx2+viaddmnmx", 30: "hmnmx2+prmt"}
    probe = {}
    for kind, nm in names.items():
        g, ms = ctypes.c_double(), ctypes.c_double()
        nat.check(nat.lib.ovl_int_peak_probe(eng._ctx, kind, 4000, ctypes.byref(g), ctypes.byref(ms)))
        probe[nm] = round(g.value, 1)
    print(json.dumps({"int_probe_gops": probe}), flush=True)


if __name__ == "__main__":
    main()
