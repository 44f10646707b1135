import os, sys, ctypes as C
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from eddy_currents_3d_b200 import lib, plate
from oracle import oracle
def rel(a, b): return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)
p = plate(32, "A")
h = lib.Handle(p, device=0)
ref = oracle.OracleRun(p)
nC = p.nCells
L = oracle.lib(); P = oracle._p
def parts(tag, a, b):
    d = np.abs(a - b)
    print("   ", tag, "maxerr Ax,Ay,Az,U:", [float(d[c*nC:(c+1)*nC].max()) for c in range(3)], float(d[3*nC:].max()), " nbad", int((d > 0).sum()))
for s in range(2):
    f, v = p.source_scalars(ref.T)
    fv = np.ascontiguousarray(f); vv = np.zeros(1)
    L.orc_scatter_sources(C.byref(ref.A.grid), C.byref(ref.src), C.byref(ref.cond), P(fv), P(vv), P(ref.Jaf), P(ref.Jafbuf), P(ref.new_nodes))
    L.orc_rhs_pre(C.byref(ref.A.grid), C.byref(ref.cond), C.byref(ref.A.csr), P(ref.Uaf), P(ref.Jaf))
    h.stage(0, f, v); h.stage(1)
    U, J = h.get_fields()
    parts("after pre  J", J, ref.Jaf); parts("after pre  U", U, ref.Uaf)
    it_o = oracle.bicgstabwr(ref.A.valA, ref.A.irow, ref.A.jcol, ref.Jaf, ref.Uaf, p.tolerance, p.itmax)
    it_g = h.stage(2)
    U, J = h.get_fields()
    print("  iter", it_g, it_o)
    parts("after solve U", U, ref.Uaf); parts("after solve J", J, ref.Jaf)
    L.orc_rhs_post(C.byref(ref.A.grid), C.byref(ref.cond), C.byref(ref.A.csr), P(ref.Uaf), P(ref.Jaf))
    h.stage(3)
    U, J = h.get_fields()
    parts("after post U", U, ref.Uaf); parts("after post J", J, ref.Jaf)
    ref.T += p.dt
