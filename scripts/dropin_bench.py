"""Strict drop-in (sprsbcgstabwr_, the reference's own CSR arrays) on plate(N): ms per BiCGSTABwr
iteration through the C ABI with host buffers, next to the matrix-free resident path."""
import os, sys, json, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from eddy_currents_3d_b200 import lib, plate
N = int(sys.argv[1]) if len(sys.argv) > 1 else 128
iters = 50
p = plate(N, "A")
h = lib.Handle(p, device=0)
A = h.assemble_csr()
n, nnz = p.nCellsGlob, int(A["num_nz"][4])
rng = np.random.default_rng(1)
b = rng.standard_normal(n)
out = {"N": N, "n": n, "nnz": nnz}
for rep in range(3):
    x = np.zeros(n)
    t0 = time.perf_counter()
    it = lib.sprsBCGstabWR(A["valA"], A["irow"], A["jcol"], n, b, x, 0.0, iters - 1)
    dt = time.perf_counter() - t0
    out[f"call{rep}_ms"] = round(1e3 * dt, 2)
ms_it = 1e3 * dt / it
bytes_it = 2 * (12.0 * nnz + 4.0 * n + 16.0 * n) + 8.0 * n * 15
out.update({"iters": it, "ms_per_iter_incl_copies": round(ms_it, 4), "csr_iter_bytes": bytes_it,
            "GBps": round(bytes_it / ms_it / 1e6, 1)})
x2 = np.zeros(n)
t0 = time.perf_counter(); it2 = h.solve(b, x2); dt2 = time.perf_counter() - t0
out["matrix_free_same_rhs_iters"] = it2
print(json.dumps(out))
