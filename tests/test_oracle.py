"""CPU tests of the oracle (oracle/ec3d_oracle.c) against everything that can pin it without a
Fortran compiler: the structural counts the reference prints (derived independently in SURVEY.md
section 8), README cell counts, analytic identities of the stencil, the golden json files, and the
documented semantics of sprsBCGstabWR (solvers.f90:3-50)."""
import json
import os

import numpy as np
import pytest

from conftest import DECKS, GOLDEN

# independent derivation (SURVEY.md section 8 size table): what EC3D.f90:113,968-971,993 print
EXPECTED = {
    "compare_to_Elmer": dict(grid=(102, 102, 24), nC=249696, Nc=43200, n=792288,
                             nz=(1805112, 1805112, 1818072, 463776), nnz=5892072, bnd=(1440, 1440, 14400), steps=100),
    "ec_src_move_hole": dict(grid=(102, 102, 24), nC=249696, Nc=43200, n=792288,
                             nz=(1805112, 1805112, 1818072, 463776), nnz=5892072, bnd=(1440, 1440, 14400), steps=101),
    "LIM": dict(grid=(176, 32, 22), nC=123904, Nc=19152, n=390864,
                nz=(885440, 886584, 894792, 186792), nnz=2853608, bnd=(224, 1368, 9576), steps=200),
}


@pytest.fixture(scope="module")
def assembled(oracle_mod, deck_problems):
    return {d: oracle_mod.Assembled(deck_problems[d]) for d in DECKS}


@pytest.mark.parametrize("deck", DECKS)
def test_structural_counts(deck, deck_problems, assembled):
    p, A, e = deck_problems[deck], assembled[deck], EXPECTED[deck]
    assert (p.sdx, p.sdy, p.sdz) == e["grid"]
    assert p.nCells == e["nC"] and p.nCells0 == e["Nc"] and p.nCellsGlob == e["n"]   # README.md:109,235
    assert A.rc == 0
    assert (A.num_nzX, A.num_nzY, A.num_nzZ, A.num_nzU) == e["nz"] and A.num_nz == e["nnz"]
    assert (len(A.cel_bndX), len(A.cel_bndY), len(A.cel_bndZ)) == e["bnd"]
    assert (len(A.cel_bndUx), len(A.cel_bndUy), len(A.cel_bndUz)) == e["bnd"]
    assert p.n_steps() == e["steps"]                                                    # EC3D.f90:452-455
    assert A.irow[0] == 1 and A.irow[-1] == e["nnz"] + 1
    assert np.all(np.diff(A.irow) >= 4) and np.all(np.diff(A.irow) <= 13)


@pytest.mark.parametrize("deck", DECKS)
def test_golden_json(deck, assembled):
    A = assembled[deck]
    g = json.load(open(os.path.join(GOLDEN, deck + "_oracle.json")))
    assert g["num_nz"] == [A.num_nzX, A.num_nzY, A.num_nzZ, A.num_nzU, A.num_nz]
    assert g["irow_sum"] == int(A.irow.astype(np.int64).sum())
    assert g["jcol_sum"] == int(A.jcol.astype(np.int64).sum())
    assert g["valA_sum"] == float(A.valA.sum()) and g["valA_abs_sum"] == float(np.abs(A.valA).sum())


@pytest.mark.parametrize("deck", ["LIM"])
def test_golden_steps(deck, oracle_mod, deck_problems, assembled):
    g = json.load(open(os.path.join(GOLDEN, deck + "_oracle.json")))
    run = oracle_mod.OracleRun(deck_problems[deck], assembled[deck])
    idx = np.array(g["sample_idx"])
    for st in g["steps"]:
        it = run.step()
        assert it == st["iter"]
        assert np.array_equal(run.Uaf[idx], np.array(st["U_sample"]))
        assert np.array_equal(run.Jaf[idx], np.array(st["J_sample"]))
        assert float(np.linalg.norm(run.Uaf)) == st["Unorm"]


@pytest.mark.parametrize("deck", DECKS)
def test_stencil_identities(deck, deck_problems, assembled):
    """Rows sorted by column, unique columns; interior air rows sum to zero; the Laplacian part of
    interior rows is symmetric; interior U rows: U part sums to 0 and A-coupling part sums to 0."""
    p, A = deck_problems[deck], assembled[deck]
    n, nC = p.nCellsGlob, p.nCells
    irow0 = A.irow.astype(np.int64) - 1
    lens = np.diff(irow0)
    rows = np.repeat(np.arange(n), lens)
    # ascending, duplicate-free columns inside each row
    d = np.diff(A.jcol.astype(np.int64))
    same_row = rows[1:] == rows[:-1]
    assert np.all(d[same_row] > 0)
    assert A.jcol.min() >= 1 and A.jcol.max() <= n
    rowsum = np.zeros(n)
    np.add.at(rowsum, rows, A.valA)
    g3 = p.geoPHYS_C.reshape(p.sdz, p.sdy, p.sdx)
    interior = np.zeros((p.sdz, p.sdy, p.sdx), bool)
    interior[1:-1, 1:-1, 1:-1] = True
    air_int = (interior & (g3 == 0)).reshape(-1)
    scale = np.abs(A.valA).max()
    for c in range(3):
        assert np.all(lens[c * nC:(c + 1) * nC][air_int] == 7)
        assert np.max(np.abs(rowsum[c * nC:(c + 1) * nC][air_int])) <= 1e-12 * scale
    # U rows
    ulen = lens[3 * nC:]
    assert set(np.unique(ulen)) <= {7, 13}
    ucols = A.jcol.astype(np.int64)
    urows = rows >= 3 * nC
    int13 = np.isin(rows, 3 * nC + np.flatnonzero(ulen == 13))
    upart = urows & int13 & (ucols > 3 * nC)
    apart = urows & int13 & (ucols <= 3 * nC)
    su = np.zeros(n); np.add.at(su, rows[upart], A.valA[upart])
    sa = np.zeros(n); np.add.at(sa, rows[apart], A.valA[apart])
    assert np.max(np.abs(su)) <= 1e-12 * scale
    assert np.max(np.abs(sa)) <= 1e-9 * np.abs(A.valA[apart]).max()
    # symmetry of the A-A block on interior air cells (pure Laplacian there)
    import scipy.sparse as sp
    M = sp.csr_matrix((A.valA, A.jcol - 1, irow0), shape=(n, n))
    sel = np.flatnonzero(air_int)
    # restrict to cells whose six neighbours are interior air as well
    S = M[:nC, :nC]
    D = (S - S.T).tocsr()
    nb_ok = air_int.copy()
    for sh in (1, p.sdx, p.sdx * p.sdy):
        nb_ok[sh:] &= air_int[:-sh]
        nb_ok[:-sh] &= air_int[sh:]
    sel = np.flatnonzero(nb_ok)
    assert abs(D[sel][:, sel]).max() == 0.0


def test_boundary_rows_small(oracle_mod):
    """Every one of the 26 domain-boundary cases (EC3D.f90:528-646) on a 4x5x6 air box with distinct
    BND values and spacings: entry count 4/5/6, neighbour coefficient BND(axis,side)*s on the single
    neighbour of a face axis, diagonal = sum of s (on-face axes) + 2 s (others)."""
    from eddy_currents_3d_b200.problem import Problem
    sdx, sdy, sdz = 4, 5, 6
    nC = sdx * sdy * sdz
    delta = np.array([0.1, 0.2, 0.3])
    BND = np.array([[-0.91, -0.92], [-0.93, -0.94], [-0.95, -0.96]])
    p = Problem(sdx=sdx, sdy=sdy, sdz=sdz, delta=delta, dt=1e-3, Time=1e-3, BND=BND, tolerance=1e-3, itmax=10,
                geoPHYS=np.ones(nC, np.int8), geoPHYS_C=np.zeros(nC, np.int32), valPHYS=np.array([[1.0, 0, 0, 0, 0]]),
                cond_numdom=[], cond_nod=[], cond_valdom=np.zeros(0), sources=[], numMech=0,
                evaluate_functions=lambda t: (np.zeros(0), np.zeros(0)))
    A = oracle_mod.Assembled(p)
    assert A.rc == 0 and A.num_nzU == 0
    s = 1.0 / delta ** 2
    dims = (sdx, sdy, sdz)
    strides = (1, sdx, sdx * sdy)
    for k in range(sdz):
        for j in range(sdy):
            for i in range(sdx):
                nn = i + sdx * j + sdx * sdy * k
                ijk = (i, j, k)
                want = {}
                diag_terms = []
                for a in range(3):
                    lo, hi = ijk[a] == 0, ijk[a] == dims[a] - 1
                    if lo:
                        want[nn + strides[a]] = BND[a, 1] * s[a]
                    elif hi:
                        want[nn - strides[a]] = BND[a, 0] * s[a]
                    else:
                        want[nn - strides[a]] = -s[a]
                        want[nn + strides[a]] = -s[a]
                    diag_terms.append(s[a] if (lo or hi) else 2.0 * s[a])
                onb = any(ijk[a] in (0, dims[a] - 1) for a in range(3))
                want[nn] = (diag_terms[0] + diag_terms[1]) + diag_terms[2] if onb else 2.0 * ((s[0] + s[1]) + s[2])
                for c in range(3):
                    r = c * nC + nn
                    cols = A.jcol[A.irow[r] - 1:A.irow[r + 1] - 1] - 1
                    vals = A.valA[A.irow[r] - 1:A.irow[r + 1] - 1]
                    assert list(cols) == sorted(c * nC + q for q in want)
                    assert all(vals[t] == want[cols[t] - c * nC] for t in range(len(cols)))


def test_solver_semantics(oracle_mod):
    """solvers.f90:23-29: Bnorm == 0 returns iter = 0 and leaves x; at most itmax+1 iterations;
    converged solves satisfy the stop test; warm start at the solution exits in one iteration."""
    from eddy_currents_3d_b200 import plate
    p = plate(32, "A")
    A = oracle_mod.Assembled(p)
    n = p.nCellsGlob
    x = np.full(n, 3.0)
    assert oracle_mod.bicgstabwr(A.valA, A.irow, A.jcol, np.zeros(n), x, 1e-3, 100) == 0
    assert np.all(x == 3.0)
    rng = np.random.default_rng(0)
    xt = rng.uniform(-1, 1, n)
    b = oracle_mod.spmv(A.valA, A.irow, A.jcol, xt)
    x = np.zeros(n)
    it = oracle_mod.bicgstabwr(A.valA, A.irow, A.jcol, b, x, 1e-8, 4)
    assert it == 5                                   # iter > itmax test happens before the increment
    run = oracle_mod.OracleRun(p, A)
    run.step(solve=False)
    b = run.rhs.copy()
    x = np.zeros(n)
    # (the unpreconditioned solver stagnates near 2e-3 on this singular-ish coupled system, which
    #  is why every shipped deck uses tol = 5m; tight tolerances are not reachable)
    it = oracle_mod.bicgstabwr(A.valA, A.irow, A.jcol, b, x, 5e-3, 10000)
    r = b - oracle_mod.spmv(A.valA, A.irow, A.jcol, x)
    assert 1 < it < 10000 and np.linalg.norm(r) / np.linalg.norm(b) < 5e-3
    x2 = x.copy()
    assert oracle_mod.bicgstabwr(A.valA, A.irow, A.jcol, b, x2, 2e-2, 10000) == 1


def test_norm2_and_dot(oracle_mod):
    rng = np.random.default_rng(3)
    x = rng.normal(size=10001) * 1e150
    assert np.isfinite(oracle_mod.norm2(x))           # scaled algorithm does not overflow
    assert abs(oracle_mod.norm2(x) / np.linalg.norm(x / 1e150) / 1e150 - 1) < 1e-12
    a, b = rng.normal(size=1000), rng.normal(size=1000)
    s = 0.0
    for u, v in zip(a, b):
        s += u * v
    assert oracle_mod.dot(a, b) == s                  # one sequential accumulator


def test_invalid_geometry_is_reported(oracle_mod):
    """A conductor thinner than 3 cells at a free face makes the reference hit column <= 0 and
    STOP (EC3D.f90:717-720); the oracle reports rc = 1."""
    from eddy_currents_3d_b200 import plate
    from eddy_currents_3d_b200.problem import number_conductor
    p = plate(32, "A")
    v = p.geoPHYS.astype(np.int64).reshape(32, 32, 32).copy()
    v[v == 1] = 6
    v[10:12, 8:20, 8:20] = 1                           # 2 cells thick along z
    p.geoPHYS = v.reshape(-1).astype(np.int8)
    p.geoPHYS_C, p.cond_nod = number_conductor(v.reshape(-1), [1], p.nCells)
    A = oracle_mod.Assembled(p)
    assert A.rc == 1 and A.err_col <= 0


def test_vtk_fields_restatement():
    """orc_vtk_fields (utilites.f90:222-290): curl of a linear potential is exact in the interior and
    halved at the faces (clamped indices), eddy / source fields are masked by geoPHYS_C."""
    from eddy_currents_3d_b200 import plate
    from oracle import oracle
    p = plate(16, "A")
    nC, n = p.nCells, p.nCellsGlob
    i = np.arange(nC) % p.sdx
    j = (np.arange(nC) // p.sdx) % p.sdy
    U = np.zeros(n)
    U[nC:2 * nC] = 3.0 * i * p.delta[0]            # Ay = 3 x  ->  Bz = dAy/dx = 3
    U[0:nC] = -2.0 * j * p.delta[1]                # Ax = -2 y ->  Bz -= dAx/dy = +2
    J = np.arange(n, dtype=np.float64)
    A, E, S, B = oracle.vtk_fields(p, U, J)
    g = p.geoPHYS_C.reshape(-1)
    inner = (i > 0) & (i < p.sdx - 1) & (j > 0) & (j < p.sdy - 1)
    assert np.allclose(B[inner, 2], 5.0, rtol=1e-6) and np.allclose(B[:, :2], 0.0)
    corner = (i == 0) & (j == 0)
    assert np.allclose(B[corner, 2], 2.5, rtol=1e-6)               # one-sided halves at the faces
    assert np.array_equal(A[:, 0], U[:nC].astype(np.float32))
    assert np.all(E[g == 0] == 0) and np.all(S[g != 0] == 0)
    s = -0.07957747154594766788444e7
    assert np.array_equal(E[g != 0, 1], (s * J[nC:2 * nC][g != 0]).astype(np.float32))
    assert np.array_equal(S[g == 0, 2], J[2 * nC:3 * nC][g == 0].astype(np.float32))


# ---- pinning the oracle: independent assembly, the reference's published validation curves ----------

@pytest.mark.parametrize("case", ["plate32A", "plate32B", "LIM", "compare_to_Elmer"])
def test_numpy_assembly_equals_oracle(case, oracle_mod, deck_problems):
    """a second, independent (numpy, rule-based, vectorised) assembly of the
    reference's CSR agrees with the oracle's case-by-case transcription of EC3D.f90:465-1049 bit for
    bit -- index arrays, values, and the six boundary-cell lists."""
    import numpy_assembly
    from eddy_currents_3d_b200 import plate
    p = plate(32, case[-1]) if case.startswith("plate") else deck_problems[case]
    O = oracle_mod.Assembled(p)
    irow, jcol, valA, lists = numpy_assembly.assemble(p)
    assert np.array_equal(irow, O.irow) and np.array_equal(jcol, O.jcol)
    assert np.array_equal(valA.view(np.int64), O.valA.view(np.int64))
    for nm, v in lists.items():
        assert np.array_equal(v, getattr(O, nm)), nm


def test_readme_validation_curves(oracle_mod, deck_problems):
    """The reference's only published output for this path (README.md:89-129, Fig. 5): eddy-current
    density along Line X / Line Y on the plate surface of compare_to_Elmer.vxc at t = 0.017 s.  The
    oracle, run for the 18 timesteps that lead to field_17.vtk, reproduces the EC3D curves' peaks,
    sign changes and their positions within the reading accuracy of the plots (tests/readme_validation.py)."""
    import readme_validation as rv
    p = deck_problems["compare_to_Elmer"]
    run = oracle_mod.OracleRun(p)
    for _ in range(rv.NSTEPS):
        run.step()
    f = rv.line_features(p, run.Jaf)
    print(f)
    rv.check_features(f)


def test_exact_dot_bridge_is_the_same_algorithm(oracle_mod):
    """orc_sprsBCGstabWR_exact_dots differs from orc_sprsBCGstabWR only in the rounding of the inner
    products: same iteration counts and fields to ~1e-11 on plate(32), where the reference's own
    reassociation sensitivity is that small; bit-identical operator / vector updates by construction."""
    from eddy_currents_3d_b200 import plate
    p = plate(32, "A")
    a = oracle_mod.OracleRun(p)
    b = oracle_mod.OracleRun(p, a.A, exact_dots=True)
    for s in range(3):
        f, v = p.source_scalars(a.T)
        assert a.step(f, v) == b.step(f, v)
        assert np.linalg.norm(a.Uaf - b.Uaf) <= 1e-9 * np.linalg.norm(a.Uaf)
    # odd sdx has no cell pairs: the bridge refuses instead of guessing a grouping
    with pytest.raises(RuntimeError):
        q = plate(32, "A"); q.sdx = 31
        oracle_mod.bicgstabwr_exact_dots(a.A.valA, a.A.irow, a.A.jcol, a.Jaf, a.Uaf.copy(), 5e-3, 10, q)


def test_committed_bridge_record_reproduces(oracle_mod):
    """tests/golden/parity_bridge.json (the record the GPU hash tests compare with) is what the bridge
    solver produces: re-run plate(32) and compare iteration counts and SHA-256 of the fields."""
    import hashlib
    from eddy_currents_3d_b200 import plate
    rec = json.load(open(os.path.join(GOLDEN, "parity_bridge.json")))["32"]["free"]
    p = plate(32, "A")
    run = oracle_mod.OracleRun(p, exact_dots=True)
    for r in rec:
        assert run.step() == r["it_bridge"]
        assert hashlib.sha256(run.Uaf.tobytes()).hexdigest() == r["sha256_U_bridge"]
        assert hashlib.sha256(run.Jaf.tobytes()).hexdigest() == r["sha256_J_bridge"]
