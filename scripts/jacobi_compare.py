"""iterations per timestep with the optional Jacobi scaling next to the reference's
unpreconditioned BiCGSTABwr (default), same plate(N) timesteps, one GPU."""
import json, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from eddy_currents_3d_b200 import lib, plate
N = int(sys.argv[1]) if len(sys.argv) > 1 else 256
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 6
out = {"N": N}
for kind, name in ((0, "none"), (1, "jacobi")):
    p = plate(N, "A")
    h = lib.Handle(p, device=0)
    h.set_preconditioner(kind)
    its, ms = [], []
    for s in range(steps):
        its.append(h.step())
        ms.append(round(h.counters()["last_step_ms"], 1))
    out[name] = {"iters_per_step": its, "ms_per_step": ms}
    h.close()
print(json.dumps(out))
