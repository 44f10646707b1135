"""A second, independent CPU assembly of the reference's CSR (test infrastructure, numpy only).

The oracle (oracle/ec3d_oracle.c) transcribes gen_sparse_matrix (reference src/EC3D.f90:465-1049) case
by case -- the 26 boundary cases of the A rows and the 8 corner / 12 edge / 6 face / interior cascade
of the U rows.  This module builds the same matrix from the RULES those cases follow (SURVEY.md
Appendix A), vectorised over the whole grid and written without looking at the cascade:

  * per axis a cell is in one of three states (both neighbours exist / only '+' exists / only '-'
    exists); a boundary A row is the product of three per-axis states, a U row likewise with
    "exists" = "is a conductor cell";
  * entries are generated as COO triplets and sorted by (row, column).

Agreement with the oracle (tests/test_oracle.py::test_numpy_assembly_*) pins the oracle's cascade:
a transcription slip in any of its ~40 hand-written cases would show as a differing column or value.
"""
from __future__ import annotations

import numpy as np


def assemble(p):
    """Returns (irow, jcol, valA, lists) with 1-based index values like the reference arrays;
    lists = dict of cel_bndX/Y/Z (global A-row numbers) and cel_bndUx/Uy/Uz (global U numbers)."""
    sdx, sdy, sdz = p.sdx, p.sdy, p.sdz
    nC = sdx * sdy * sdz
    d = np.asarray(p.delta, np.float64)
    dt = float(p.dt)
    s = np.array([1.0 / (d[0] * d[0]), 1.0 / (d[1] * d[1]), 1.0 / (d[2] * d[2])])
    ds = 0.5 / d
    BND = np.asarray(p.BND, np.float64)               # BND[axis, 0]: used at the HIGH face, [axis, 1]: LOW face
    geo = np.asarray(p.geoPHYS, np.int64).reshape(sdz, sdy, sdx)
    gC = np.asarray(p.geoPHYS_C, np.int64).reshape(sdz, sdy, sdx)
    valPHYS = np.asarray(p.valPHYS, np.float64)
    kk, jj, ii = np.meshgrid(np.arange(sdz), np.arange(sdy), np.arange(sdx), indexing="ij")
    nn = (ii + sdx * jj + sdx * sdy * kk + 1)           # 1-based cell numbers
    idx = (ii, jj, kk)
    dims = (sdx, sdy, sdz)
    stride = (1, sdx, sdx * sdy)
    axis_np = (2, 1, 0)                                 # numpy axis of x, y, z

    rows, cols, vals = [], [], []

    def emit(mask, r, c, v):
        m = np.asarray(mask)
        if m.any():
            rows.append(np.broadcast_to(r, m.shape)[m].astype(np.int64))
            cols.append(np.broadcast_to(c, m.shape)[m].astype(np.int64))
            vals.append(np.broadcast_to(v, m.shape)[m].astype(np.float64))

    lo = [idx[a] == 0 for a in range(3)]
    hi = [idx[a] == dims[a] - 1 for a in range(3)]
    on_face = lo[0] | hi[0] | lo[1] | hi[1] | lo[2] | hi[2]
    cond = (gC != 0)
    assert not (cond & on_face).any(), "conductor cells on a domain face (reference would STOP)"

    # ---------------- A rows (three identical blocks, column offset comp*nC) ----------------
    # diagonal: per axis s on a face of that axis, 2 s otherwise, summed x, y, z
    t = [np.where(lo[a] | hi[a], s[a], 2.0 * s[a]) for a in range(3)]
    diag_face = (t[0] + t[1]) + t[2]
    diag_int = 2.0 * (s[0] + s[1] + s[2])
    mat = geo - 1
    C = np.where(cond, valPHYS[np.clip(mat, 0, valPHYS.shape[0] - 1), 1], 0.0)
    V = [np.where(cond, valPHYS[np.clip(mat, 0, valPHYS.shape[0] - 1), 2 + a], 0.0) for a in range(3)]
    diag_cond = diag_int + 2.0 * C / dt
    diag = np.where(on_face, diag_face, np.where(cond, diag_cond, diag_int))

    def shifted(arr, a, k):
        """arr at the cell k steps along axis a (0 outside the grid)."""
        out = np.zeros_like(arr)
        ax = axis_np[a]
        src = [slice(None)] * 3
        dst = [slice(None)] * 3
        if k > 0:
            src[ax] = slice(k, None); dst[ax] = slice(0, -k)
        else:
            src[ax] = slice(0, k); dst[ax] = slice(-k, None)
        out[tuple(dst)] = arr[tuple(src)]
        return out

    g_sh = {(a, k): shifted(gC, a, k) for a in range(3) for k in (-2, -1, 1, 2)}

    for comp in range(3):
        r = nn + comp * nC
        emit(np.ones_like(cond), r, r, diag)
        for a in range(3):
            # '-' neighbour: absent on the low face; BND(a,1)*s on the high face; -s (- V/(2 d) in a conductor) otherwise
            cm = np.where(hi[a], BND[a, 0] * s[a], np.where(cond, -s[a] - V[a] / (2.0 * d[a]), -s[a]))
            emit(~lo[a], r, r - stride[a], cm)
            cp = np.where(lo[a], BND[a, 1] * s[a], np.where(cond, -s[a] + V[a] / (2.0 * d[a]), -s[a]))
            emit(~hi[a], r, r + stride[a], cp)
        # grad U along this component's own axis (conductor cells only)
        a = comp
        gm1, gp1, gm2, gp2 = g_sh[(a, -1)], g_sh[(a, 1)], g_sh[(a, -2)], g_sh[(a, 2)]
        back = cond & (gp1 == 0)                        # '+' neighbour is not a conductor: backward one-sided
        fwd = cond & ~back & (gm1 == 0)                 # only the '-' neighbour is missing: forward one-sided
        cen = cond & ~back & ~fwd
        Cd1, Cd3, Cd4 = C * ds[a], (3.0 * C) * ds[a], (4.0 * C) * ds[a]
        emit(back, r, gC, -Cd3); emit(back, r, gm1, Cd4); emit(back, r, gm2, -Cd1)
        emit(fwd, r, gC, Cd3); emit(fwd, r, gp1, -Cd4); emit(fwd, r, gp2, Cd1)
        emit(cen, r, gp1, -Cd1); emit(cen, r, gm1, Cd1)
        if comp == 0:
            onesided = [None, None, None]
        onesided[comp] = back | fwd

    # ---------------- U rows ----------------
    miss_m = [cond & (g_sh[(a, -1)] == 0) for a in range(3)]     # '-' neighbour is not a conductor
    miss_p = [cond & (g_sh[(a, 1)] == 0) for a in range(3)]
    assert not any((miss_m[a] & miss_p[a]).any() for a in range(3)), "conductor thinner than 3 cells"
    any_missing = miss_m[0] | miss_p[0] | miss_m[1] | miss_p[1] | miss_m[2] | miss_p[2]
    interior = cond & ~any_missing
    rU = gC
    emit(cond, rU, gC, diag_int)
    anomaly = miss_m[0] & miss_p[1] & miss_p[2]                  # EC3D.f90:803-807: a, b signs swapped
    for a in range(3):
        both = cond & ~miss_m[a] & ~miss_p[a]
        emit(both, rU, g_sh[(a, -1)], -s[a]); emit(both, rU, g_sh[(a, 1)], -s[a])
        emit(miss_m[a], rU, g_sh[(a, 1)], -2.0 * s[a])
        emit(miss_p[a], rU, g_sh[(a, -1)], -2.0 * s[a])
        # A couplings: interior rows look at the neighbours' A, surface rows at the same cell's A
        emit(interior, rU, a * nC + nn + stride[a], 0.5 / dt * (-1.0 / d[a]))
        emit(interior, rU, a * nC + nn - stride[a], 0.5 / dt * (1.0 / d[a]))
        cm_, cp_ = -2.0 / (dt * d[a]), 2.0 / (dt * d[a])
        sign_m, sign_p = cm_, cp_
        if a == 0:
            emit(miss_m[a] & ~anomaly, rU, a * nC + nn, sign_m); emit(miss_m[a] & anomaly, rU, a * nC + nn, cp_)
            emit(miss_p[a], rU, a * nC + nn, sign_p)
        elif a == 1:
            emit(miss_m[a], rU, a * nC + nn, sign_m)
            emit(miss_p[a] & ~anomaly, rU, a * nC + nn, sign_p); emit(miss_p[a] & anomaly, rU, a * nC + nn, cm_)
        else:
            emit(miss_m[a], rU, a * nC + nn, sign_m); emit(miss_p[a], rU, a * nC + nn, sign_p)

    R = np.concatenate(rows); Cc = np.concatenate(cols); Vv = np.concatenate(vals)
    order = np.lexsort((Cc, R))
    R, Cc, Vv = R[order], Cc[order], Vv[order]
    n = 3 * nC + int(cond.sum())
    counts = np.bincount(R - 1, minlength=n)
    irow = np.concatenate(([1], 1 + np.cumsum(counts))).astype(np.int64)
    flat = lambda m: m.reshape(-1)
    lists = {
        "cel_bndX": (nn.reshape(-1)[flat(onesided[0])]),
        "cel_bndY": (nn.reshape(-1)[flat(onesided[1])] + nC),
        "cel_bndZ": (nn.reshape(-1)[flat(onesided[2])] + 2 * nC),
        "cel_bndUx": gC.reshape(-1)[flat(miss_m[0] | miss_p[0])],
        "cel_bndUy": gC.reshape(-1)[flat(miss_m[1] | miss_p[1])],
        "cel_bndUz": gC.reshape(-1)[flat(miss_m[2] | miss_p[2])],
    }
    return irow, Cc, Vv, lists
