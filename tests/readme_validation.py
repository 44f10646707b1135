"""The only output the reference PUBLISHES for the hot path: README.md:89-129, Fig. 5
(img/valid_Line_Xa.jpg, img/valid_Line_Ya.jpg) -- eddy-current density on the plate surface of the
TEAM-7-like deck compare_to_Elmer.vxc along "Line X" and "Line Y" at t = 0.017 s, EC3D (solid) next
to Elmer FEM (dashed).  Values read off the EC3D curves (reading accuracy about 5 %):

  Line X (y = 160 mm from the plate edge, x from the plate's left edge):
      |J| ~ 1.2e5 A/m^2 at x ~ 0.005 m;  Jx plateau ~ 1.0e5 at x ~ 0.08-0.10 m;
      Jy minimum ~ -0.9e5 at x ~ 0.19 m (under the coil's right bar);  |J| ~ 0.95e5 at x ~ 0.19 m
  Line Y (x = 140 mm, y from the plate's lower edge):
      Jx ~ -1.6e5 at y ~ 0.005 m; Jx changes sign at y ~ 0.10-0.13 m; Jx maximum ~ +0.9e5 at y ~ 0.185 m;
      Jy ~ -0.35e5 around y ~ 0.10 m;  |J| minimum ~ 0.35e5 at y ~ 0.10 m

Geometry (img/domain_size.jpg): plate cells i, j = 7..96 (1-based), top layer
k = 8; cell size 3.333 mm, so Line X is the row j = 55 and Line Y the column i = 49; field_17.vtk is
written by the step whose source time is T = 0.017, i.e. after 18 steps (EC3D.f90:436-455).
Vector_field_eddy = s * Jaf on conductor cells, s = -1/mu0 literal (utilites.f90:237-250).
"""
import numpy as np

S_EDDY = -0.07957747154594766788444e7
NSTEPS = 18


def line_features(p, Jaf):
    nC, sdx, sdy, sdz = p.nCells, p.sdx, p.sdy, p.sdz
    J = np.asarray(Jaf)[:3 * nC].reshape(3, sdz, sdy, sdx) * S_EDDY
    k, j, i = 7, 54, 48                                   # 0-based: top conductor layer, Line X row, Line Y column
    cells = slice(6, 96)
    x = (np.arange(90) + 0.5) * p.delta[0]
    Jx, Jy, Jz = J[0, k, j, cells], J[1, k, j, cells], J[2, k, j, cells]
    Jm = np.sqrt(Jx ** 2 + Jy ** 2 + Jz ** 2)
    f = {"X_Jm_start": float(Jm[1]), "X_Jx_max": float(Jx.max()), "X_Jx_max_at": float(x[Jx.argmax()]),
         "X_Jy_min": float(Jy.min()), "X_Jy_min_at": float(x[Jy.argmin()]), "X_Jm_at_0.19": float(Jm[57])}
    y = (np.arange(90) + 0.5) * p.delta[1]
    Jx, Jy, Jz = J[0, k, cells, i], J[1, k, cells, i], J[2, k, cells, i]
    Jm = np.sqrt(Jx ** 2 + Jy ** 2 + Jz ** 2)
    f.update({"Y_Jx_start": float(Jx[1]), "Y_Jx_max": float(Jx.max()), "Y_Jx_max_at": float(y[Jx.argmax()]),
              "Y_Jx_zero_at": float(y[np.argmax(Jx > 0)]), "Y_Jy_at_0.10": float(Jy[30]),
              "Y_Jm_min": float(Jm[5:60].min()), "Y_Jm_min_at": float(y[5 + Jm[5:60].argmin()])})
    return f


# (published value, relative tolerance) / (published position, absolute tolerance in m)
PUBLISHED = {
    "X_Jm_start": (1.2e5, 0.20), "X_Jx_max": (1.0e5, 0.15), "X_Jy_min": (-0.9e5, 0.15), "X_Jm_at_0.19": (0.95e5, 0.15),
    "Y_Jx_start": (-1.6e5, 0.15), "Y_Jx_max": (0.9e5, 0.15), "Y_Jy_at_0.10": (-0.35e5, 0.30), "Y_Jm_min": (0.35e5, 0.30),
}
POSITIONS = {"X_Jx_max_at": (0.09, 0.02), "X_Jy_min_at": (0.19, 0.015), "Y_Jx_max_at": (0.185, 0.015),
             "Y_Jx_zero_at": (0.115, 0.025), "Y_Jm_min_at": (0.10, 0.02)}


def check_features(f):
    for k, (v, tol) in PUBLISHED.items():
        assert abs(f[k] - v) <= tol * abs(v), (k, f[k], v)
    for k, (v, tol) in POSITIONS.items():
        assert abs(f[k] - v) <= tol, (k, f[k], v)
