"""Digest of eddy_currents_3d_b200/csrc/ptxas.log: registers / stack / spills per kernel (for profiles/)."""
import os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
t = open(os.path.join(ROOT, "eddy_currents_3d_b200", "csrc", "ptxas.log")).read()
blocks = re.split(r"ptxas info\s+: Compiling entry function '", t)[1:]
flt = sys.argv[1] if len(sys.argv) > 1 else ""
for b in blocks:
    name = b.split("'")[0]
    dem = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip() or name
    m = re.search(r"Used (\d+) registers", b)
    sp = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", b)
    short = re.sub(r"\(.*", "", dem).replace("void ", "")
    if flt and flt not in short:
        continue
    print(f"{short:64s} regs={m.group(1) if m else '?':>3s} stack={sp.group(1):>4s} spill_st={sp.group(2):>4s} spill_ld={sp.group(3):>4s}")
