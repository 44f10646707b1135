"""Per-kernel timings (CUDA events inside the library) on plate(N)."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from eddy_currents_3d_b200 import lib, plate
N = int(sys.argv[1]) if len(sys.argv) > 1 else 256
only = [int(a) for a in sys.argv[2].split(",")] if len(sys.argv) > 2 else [0, 1, 2, 3, 4, 5]
warm = int(sys.argv[3]) if len(sys.argv) > 3 else 3
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 20
p = plate(N, "A")
h = lib.Handle(p, device=0)
n, nC = p.nCellsGlob, p.nCells
fused = os.environ.get("EC3D_FUSE_S", "1") != "0"
names = ["spmv_Ap", "spmv_SAs" if fused else "spmv_As", "s_update", "xr_update", "p_update", "iteration"]
byt = [24.0 * n + 5 * nC, (32.0 if fused else 16.0) * n + 5 * nC, 24.0 * n, 56.0 * n, 32.0 * n, 152.0 * n + 10 * nC]
out = {}
for w in only:
    ms = h.bench_kernel(w, warm, reps)
    out[names[w]] = (round(ms, 4), round(byt[w] / ms / 1e6, 1))
line = {"N": N, "env": {k: v for k, v in os.environ.items() if k.startswith("EC3D_")}, "kernels(ms,GB/s)": out}
print(json.dumps(line))
h.close()
