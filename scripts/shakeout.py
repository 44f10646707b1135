"""Scratch GPU shake-out: assembly / operator / solver / step parity on small cases."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from eddy_currents_3d_b200 import lib, plate
from oracle import oracle

def rel(a, b):
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)

for variant in ("A", "B", "M"):
    p = plate(32, variant)
    t = time.time(); h = lib.Handle(p, device=0); print(variant, "create", time.time() - t)
    ref = oracle.OracleRun(p)
    A = h.assemble_csr()
    print(" asm irow", np.array_equal(A["irow"], ref.A.irow), "jcol", np.array_equal(A["jcol"], ref.A.jcol),
          "valA", np.array_equal(A["valA"], ref.A.valA), "nz", A["num_nz"], ref.A.num_nz)
    for nm in ("cel_bndX", "cel_bndY", "cel_bndZ", "cel_bndUx", "cel_bndUy", "cel_bndUz"):
        print("  ", nm, np.array_equal(A[nm], getattr(ref.A, nm)), len(A[nm]))
    rng = np.random.default_rng(1)
    x = rng.uniform(-1, 1, p.nCellsGlob)
    y = h.apply_operator(x)
    yo = oracle.spmv(ref.A.valA, ref.A.irow, ref.A.jcol, x)
    print(" spmv bitwise", np.array_equal(y, yo), "rel", rel(y, yo), "nbad", int((y != yo).sum()))
    # drop-in
    b = rng.uniform(-1, 1, p.nCellsGlob); x0 = np.zeros(p.nCellsGlob); x1 = x0.copy()
    it_o = oracle.bicgstabwr(ref.A.valA, ref.A.irow, ref.A.jcol, b, x0, 1e-6, 500)
    it_g = lib.sprsBCGstabWR(ref.A.valA, ref.A.irow, ref.A.jcol, p.nCellsGlob, b, x1, 1e-6, 500)
    print(" dropin iter", it_g, it_o, "rel", rel(x1, x0))
    x2 = np.zeros(p.nCellsGlob)
    h2 = lib.Handle(p.__class__(**{**p.__dict__, "tolerance": 1e-6, "itmax": 500}), device=0)
    it_m = h2.solve(b, x2)
    print(" matrix-free solve iter", it_m, "rel", rel(x2, x0), "bitwise vs dropin", np.array_equal(x2, x1))
    h2.close()
    for s in range(4):
        f, v = p.source_scalars(ref.T)
        it_o = ref.step(f, v); it_g = h.step(f, v)
        U, J = h.get_fields()
        print(" step", s, "iter", it_g, it_o, "relU", rel(U, ref.Uaf), "relJ", rel(J, ref.Jaf), h.counters())
        if p.numfun and ref.flag_move:
            print("   cells equal", np.array_equal(h.source_cells(), ref.new_nodes[:len(h.source_cells())]))
    h.close()
