"""Build DP-kernel variants into build/variants/libovl_<name>.so (git-ignored; they travel to the GPU box).
    python tools/build_variants.py name=-DOVL_DP_PATTERN=1,1,3 other=-DOVL_DP_MINB=4,-DX=1
Prints ptxas' register / spill line of the 4x38 immediate-gap instantiation for each."""
import os
import re
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge


def one(spec):
    name, _, flags = spec.partition("=")
    flags = [f for f in flags.split(";") if f]
    out = os.path.join(ROOT, "build", "variants", f"libovl_{name}.so")
    cmd = ["nvcc", *ge.NVCC_FLAGS, "-Xptxas", "-v", *flags, "-o", out, os.path.join(ge.CSRC, "ovl.cu"), "-lcudart"]
    r = subprocess.run(cmd, cwd=ge.CSRC, capture_output=True, text=True)
    if r.returncode:
        return name, "BUILD FAILED\n" + r.stderr[-2000:]
    lines = r.stderr.splitlines()
    info = []
    for i, ln in enumerate(lines):
        if "overlap_dp_kernel" in ln and ("ILi4ELi38ELb1ELi2ELb1" in ln or "ILi32ELi32ELb1ELi2ELb1" in ln) and "Compiling" in ln:
            tag = "4x38imm" if "ILi4ELi38" in ln else "32x32imm"
            blk = " ".join(lines[i + 1:i + 5])
            m = re.search(r"Used (\d+) registers", blk)
            sp = re.search(r"(\d+) bytes spill stores", blk)
            info.append(f"{tag}: regs={m.group(1) if m else '?'} spill={sp.group(1) if sp else '?'}")
    return name, "; ".join(info)


if __name__ == "__main__":
    os.makedirs(os.path.join(ROOT, "build", "variants"), exist_ok=True)
    with ThreadPoolExecutor(max_workers=8) as ex:
        for name, info in ex.map(one, sys.argv[1:]):
            print(name, "->", info, flush=True)
