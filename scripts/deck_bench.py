"""ms per timestep of the three shipped decks (tests/golden/*.npz = what vxc2data hands to the hot
path) on one GPU: resident stepping (ec3d_step) and the strict drop-in (sprsbcgstabwr_ on the
reference's CSR with host buffers every call), first `nsteps` timesteps."""
import os, sys, json, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from eddy_currents_3d_b200 import lib
from eddy_currents_3d_b200.problem import load_problem_npz
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
nsteps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
for deck in ("compare_to_Elmer", "ec_src_move_hole", "LIM"):
    p = load_problem_npz(os.path.join(ROOT, "tests", "golden", deck + ".npz"))
    h = lib.Handle(p, device=0)
    T, its, ms = 0.0, [], []
    for s in range(nsteps):
        f, v = p.source_scalars(T); T += p.dt
        t0 = time.perf_counter(); it = h.step(f, v); dt = time.perf_counter() - t0
        its.append(it); ms.append(1e3 * dt)
    c = h.counters()
    kern = {name: round(1e3 * h.bench_kernel(which, 5, 50), 2) for which, name, *_ in lib.KERNELS}   # us per launch
    kern["whole_iteration_enqueued"] = round(1e3 * h.bench_kernel(5, 5, 50), 2)
    print(json.dumps({"deck": deck, "n": p.nCellsGlob, "steps": nsteps, "iters": its,
                      "ms_per_step_mean_after_first": round(float(np.mean(ms[1:])), 3),
                      "us_per_iteration": round(1e3 * float(np.sum(ms[1:])) / max(sum(its[1:]), 1), 2),
                      "us_per_kernel_launch": kern}))
    h.close()
