"""CPU tests of the host side: expression evaluator / numeric prefixes (harness row N1), the .vxc
loader against the golden problems, the slab partition, and the C-ABI library (loads, exports every
symbol of include/ec3d_gpu.h; no compute calls without a GPU)."""
import ctypes
import os
import re

import numpy as np
import pytest

from conftest import DECKS, GOLDEN, ROOT


def test_numeric_prefixes():
    from eddy_currents_3d_b200.fparser import numeric
    assert numeric("100m") == 1e-3 * 100.0          # utilites.f90:343-475
    assert numeric("0.4m") == 1e-3 * 0.4
    assert numeric("5m") == 1e-3 * 5.0
    assert numeric("10000") == 10000.0
    assert numeric("1k3") == 1e3 * 1.3
    assert numeric("2meg") == 1e6 * 2.0
    assert numeric("50") == 50.0
    assert numeric("1e-3") == 1e-3


def test_fparser_operator_split():
    """m_fparser.f90:633-657: operators are split in the order + - * / ^ from the right, so
    a*b/c = a*(b/c) and a+b-c = a+(b-c)."""
    from eddy_currents_3d_b200.fparser import evalf
    v = {"A": 3.0, "B": 7.0, "C": 0.3, "T": 0.01}
    assert evalf("A*B/C", v) == 3.0 * (7.0 / 0.3)
    assert evalf("A+B-C", v) == 3.0 + (7.0 - 0.3)
    assert evalf("A-B+C", v) == (3.0 - 7.0) + 0.3
    assert evalf("-A*B", v) == -(3.0 * 7.0)
    assert evalf("A*COS(B*C*T)", v) == 3.0 * np.cos((7.0 * 0.3) * 0.01)
    assert evalf("A*IMPL2(SIND(360*C*T))", v) == 3.0
    assert evalf("-A*COSD(360*B*T+120)", v) == -(3.0 * np.cos(np.radians((360 * 7.0) * 0.01 + 120)))
    assert evalf("2^3^2", v) == (2.0 ** 3.0) ** 2.0
    assert evalf("1.5E+2+A", v) == 153.0


def test_fparser_evaluation_errors_zero_the_whole_expression():
    """m_fparser.f90:182,187,210-214: on division by zero, lg(<=0), asin/acos outside [-1,1] evalf sets
    EvalErrType and returns ZERO for the whole expression (`res=zero; RETURN`); sqrt / ln are unchecked."""
    import math
    from eddy_currents_3d_b200.fparser import evalf
    assert evalf("1+1/0", {}) == 0.0                 # not 1
    assert evalf("5*(2+lg(0))", {}) == 0.0
    assert evalf("3-asin(2)", {}) == 0.0 and evalf("3-acos(-1.5)", {}) == 0.0
    assert evalf("1+A/B", {"A": 1.0, "B": 0.0}) == 0.0
    assert evalf("1+1/4", {}) == 1.25
    assert math.isnan(evalf("sqrt(0-1)", {})) and evalf("ln(0)", {}) == -math.inf


@pytest.mark.skipif(not os.path.isdir("/root/reference/src"), reason="reference decks not present")
@pytest.mark.parametrize("deck", DECKS)
def test_vxc_loader_matches_golden(deck, deck_problems):
    from eddy_currents_3d_b200 import load_vxc
    p = load_vxc(f"/root/reference/src/{deck}.vxc")
    q = deck_problems[deck]
    assert (p.sdx, p.sdy, p.sdz) == (q.sdx, q.sdy, q.sdz)
    assert np.array_equal(p.delta, q.delta) and p.dt == q.dt and p.Time == q.Time
    assert p.tolerance == q.tolerance and p.itmax == q.itmax and np.array_equal(p.BND, q.BND)
    assert np.array_equal(p.geoPHYS, q.geoPHYS) and np.array_equal(p.geoPHYS_C, q.geoPHYS_C)
    assert np.array_equal(p.valPHYS, q.valPHYS) and np.array_equal(p.cond_valdom, q.cond_valdom)
    assert len(p.sources) == len(q.sources) and p.numMech == q.numMech
    for a, b in zip(p.sources, q.sources):
        assert a.ex == b.ex and np.array_equal(a.nods, b.nods) and np.array_equal(a.num_Vmech, b.num_Vmech)
    for s in (0, 1, 7):
        T = 0.0
        for _ in range(s):
            T = T + p.dt
        fa, va = p.source_scalars(T)
        fb, vb = q.source_scalars(T)
        assert np.array_equal(fa, fb) and np.array_equal(va, vb)


def test_deck_parameters(deck_problems):
    p = deck_problems["compare_to_Elmer"]
    assert p.tolerance == 1e-3 * 5.0 and p.itmax == 10000 and p.dt == 1e-3 * 1.0
    assert np.all(p.BND == -0.95)
    # C = mu0*35.26e6 with the reference's MU0 literal (vxc2data.f90:402), valdom = 2C/dt (:461)
    C = 0.12566370964050292e-5 * 35.26e6
    assert p.valPHYS[0, 1] == C and p.cond_valdom[0] == 2.0 * C / p.dt
    assert [len(s.nods) for s in p.sources] == [1944] * 4
    q = deck_problems["LIM"]
    assert len(q.sources) == 12 and q.numMech == 12 and q.cond_numdom == [13]
    # node lists are in descending cell order (vxc2data.f90:656-752 pops a linked list)
    assert all(np.all(np.diff(s.nods) < 0) for s in q.sources)


def test_plate_generator():
    from eddy_currents_3d_b200 import plate
    p = plate(32, "A")
    assert p.nCells0 == 32 ** 3 // 8 and p.nCellsGlob == 3 * 32 ** 3 + 32 ** 3 // 8     # n = 3.125 N^3
    g = p.geoPHYS_C.reshape(32, 32, 32)
    assert g[:, :, 0].max() == 0 and g[0].max() == 0 and g[-1].max() == 0             # >= 1 cell from faces
    nz = g[g != 0]
    assert np.array_equal(nz, 3 * 32 ** 3 + 1 + np.arange(nz.size))                    # k,j,i numbering
    assert p.n_steps() == 11
    assert plate(32, "B").valPHYS[0, 2] != 0.0 and plate(32, "M").sources[0].move[0] == 1


def test_library_exports_every_header_symbol():
    """The C-ABI library loads without a GPU and exports every function include/ec3d_gpu.h declares."""
    from eddy_currents_3d_b200 import lib
    L = lib.load()
    hdr = open(os.path.join(ROOT, "include", "ec3d_gpu.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    names = set(re.findall(r"\b(ec3d_[a-z0-9_]+|sprsbcgstabwr_|SPRSBCGSTABWR)\s*\(", hdr))
    assert len(names) >= 20
    for nm in sorted(names):
        assert hasattr(L, nm), nm
    assert set(lib.EXPORTS) == names
    assert b"sm_100a" in L.ec3d_version()


def test_no_cpu_fallback_in_product_path():
    """The product package never imports, loads or executes anything under oracle/."""
    pkg = os.path.join(ROOT, "eddy_currents_3d_b200")
    for base, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h")) or fn == "Makefile":
                src = open(os.path.join(base, fn)).read()
                assert not re.search(r"(import|from)\s+oracle|libec3d_oracle|ec3d_oracle\.h|orc_[a-z]", src), fn


def test_partition_planes():
    from eddy_currents_3d_b200 import lib
    sdz = 64
    cpp = np.zeros(sdz, np.int64)
    cpp[8:24] = 64 * 64 // 2                               # conductor planes are heavier
    for nr in (1, 2, 4, 8):
        ks = lib.partition_planes(64, 64, sdz, cpp, nr)
        assert ks[0] == 0 and ks[-1] == sdz and np.all(np.diff(ks) >= 2)
        w = 152.0 * (3.0 * 64 * 64 + cpp) + 8.0 * 64 * 64 + 120.0 * cpp
        loads = np.array([w[ks[r]:ks[r + 1]].sum() for r in range(nr)])
        assert loads.max() / loads.mean() < 1.12, (nr, loads)
    ks8 = lib.partition_planes(64, 64, sdz, cpp, 8)
    counts = np.diff(ks8)
    assert counts[1:3].max() <= counts[4:].min()           # heavier planes -> fewer planes per rank
    with pytest.raises(lib.Ec3dError):
        lib.partition_planes(64, 64, 6, np.zeros(6, np.int64), 8)


@pytest.mark.parametrize("case", [
    # sdx, sdy, k0, k1, conductor box (i0,i1,j0,j1,k0,k1), zc
    (256, 256, 0, 256, (32, 224, 32, 224, 32, 96), 0),       # plate(256), one GPU
    (512, 512, 64, 128, (64, 448, 64, 448, 64, 192), 0),     # a slab of plate(512) inside the conductor range
    (102, 102, 0, 24, (6, 96, 6, 96, 2, 8), 0),              # compare_to_Elmer.vxc: partial tiles in x and y
    (176, 32, 5, 22, (2, 174, 2, 30, 9, 13), 3),             # LIM.vxc, upper slab, odd item length
    (64, 40, 0, 16, (0, 0, 0, 0, 0, 0), 0),                  # no conductor
])
def test_spmv_work_list_covers_every_tile_plane_once(case):
    """Host logic of the TMA SpMV (ec3d_plan_spmv_items): the (tile column, z range) items cover every
    64 x 8 tile of every owned plane exactly once, items flagged has_u are exactly those that lie in
    the conductor's z range and touch its footprint, unflagged items contain no conductor plane of a
    touching column, and the launch order is plane-major."""
    from eddy_currents_3d_b200 import lib
    sdx, sdy, k0, k1, box, zc = case
    it = lib.plan_spmv_items(sdx, sdy, k0, k1, box, zc)
    tx, ty = (sdx + 63) // 64, (sdy + 7) // 8
    cover = np.zeros((ty, tx, k1 - k0), np.int32)
    i0, i1, j0, j1, b0, b1 = box
    c0, c1 = max(k0, b0), min(k1, b1)
    for x0, y0, ka, kb, has_u in it:
        assert x0 % 64 == 0 and y0 % 8 == 0 and k0 <= ka < kb <= k1
        cover[y0 // 8, x0 // 64, ka - k0:kb - k0] += 1
        touch = b1 > b0 and c1 > c0 and x0 < i1 and x0 + 64 > i0 and y0 < j1 and y0 + 8 > j0
        if has_u:
            assert touch and c0 <= ka and kb <= c1
        elif touch:
            assert kb <= c0 or ka >= c1                       # lean items hold no conductor plane
    assert np.all(cover == 1)
    assert np.all(np.diff(it[:, 2]) >= 0)                       # plane-major launch order
    # column-major order keeps the same set of items
    it2 = lib.plan_spmv_items(sdx, sdy, k0, k1, box, zc, plane_major=False)
    assert sorted(map(tuple, it)) == sorted(map(tuple, it2))


def _synthetic_deck(tmp_path, compression: str):
    """A small VoxCad deck written from scratch (no reference file needed): 16 x 12 x 10 grid, conductor
    block (material 1), one x-directed and one y-directed coil bar (materials 2, 3), palette strings in
    the reference's keyword syntax (vxc2data.f90:127-222, 443-548)."""
    import base64
    import zlib
    sdx, sdy, sdz = 16, 12, 10
    v = np.zeros((sdz, sdy, sdx), np.uint8)
    v[2:6, 2:10, 3:13] = 1                       # conductor: k 2..5, j 2..9, i 3..12 (>= 3 thick, off the faces)
    v[7, 3:5, 4:12] = 2                          # coil bar, SRCx
    v[7, 5:9, 4:6] = 3                           # coil bar, SRCy
    names = ["plast D=1 C='mu0*35.26e6'", "axp D=1 SRCx=Fp", "ayp D=1 SRCy=Fm", "param  tran stop=10m step=1m",
             "p2 solver tol=5m itmax=500  dir=vec", "f1 func Fp=a*cos(p2*f*t) a='183/(2*dx*1*dz)' p2='2*pi' f=50 t=t ",
             "f2 func Fm=a*cos(p2*f*t) a='-183/(2*dx*1*dz)' p2='2*pi' f=50 t=t "]
    letters = "0123456789:;<=>?@ABCDEFGHIJKLMNOPQRSTUVWXYZ[\\]^_`abcdefghijklmnopqrstuvwxyz"
    out = ["<?xml version=\"1.0\" encoding=\"ISO-8859-1\"?>", "<VXC Version=\"0.94\">", "<Lattice>",
           "    <Lattice_Dim>0.004</Lattice_Dim>", "    <X_Dim_Adj>1</X_Dim_Adj>", "    <Y_Dim_Adj>1.25</Y_Dim_Adj>",
           "    <Z_Dim_Adj>1</Z_Dim_Adj>", "</Lattice>", "<Palette>"]
    for m, nm in enumerate(names, 1):
        out += [f"    <Material ID=\"{m}\">", f"      <Name>{nm}</Name>", "    </Material>"]
    out += ["</Palette>", f"  <Structure Compression=\"{compression}\">", f"    <X_Voxels>{sdx}</X_Voxels>",
            f"    <Y_Voxels>{sdy}</Y_Voxels>", f"    <Z_Voxels>{sdz}</Z_Voxels>", "    <Data>"]
    for k in range(sdz):
        raw = v[k].reshape(-1)
        if compression == "ZLIB":
            payload = base64.b64encode(zlib.compress(raw.tobytes())).decode()
        else:
            payload = "".join(letters[int(b)] for b in raw)
        out.append(f"      <Layer><![CDATA[{payload}]]></Layer>")
    out += ["    </Data>", "  </Structure>", "</VXC>"]
    path = tmp_path / f"deck_{compression}.vxc"
    path.write_text("\n".join(out) + "\n")
    return str(path), v


def test_vxc_loader_synthetic_deck_zlib_and_ascii(tmp_path):
    """Harness row N1: both voxel encodings of the reference (ZLIB+base64 layers, vxc2data.f90:253-296 and
    uncompress_zlib.py; ASCII_READABLE letters '1'..'z', :298-311) give the same Problem, with the
    numbering rules of vxc2data (geoPHYS_C = 3*nCells + m in k,j,i order, :609-652)."""
    from eddy_currents_3d_b200 import load_vxc
    pz, v = _synthetic_deck(tmp_path, "ZLIB")
    pa, _ = _synthetic_deck(tmp_path, "ASCII_READABLE")
    A, B = load_vxc(pz), load_vxc(pa)
    assert (A.sdx, A.sdy, A.sdz) == (16, 12, 10)
    assert np.allclose(A.delta, [0.004, 0.005, 0.004])
    assert A.tolerance == 5e-3 and A.itmax == 500 and abs(A.dt - 1e-3) < 1e-18
    for name in ("geoPHYS", "geoPHYS_C", "valPHYS", "delta", "BND"):
        assert np.array_equal(getattr(A, name), getattr(B, name)), name
    assert len(A.sources) == len(B.sources) == 2
    for sa, sb in zip(A.sources, B.sources):
        assert np.array_equal(sa.nods, sb.nods)
    # conductor numbering: consecutive in k,j,i order, starting at 3*nCells + 1
    g = A.geoPHYS_C.reshape(-1)
    cond = np.flatnonzero(v.reshape(-1) == 1)
    assert np.array_equal(np.flatnonzero(g), cond)
    assert np.array_equal(g[cond], 3 * A.nCells + 1 + np.arange(cond.size))
    assert A.nCellsGlob == 3 * A.nCells + cond.size
    # coil cells: SRCx nodes are plain cell numbers, SRCy nodes are offset by nCells (EC3D.f90:203-230)
    assert set(A.sources[0].nods) == set(np.flatnonzero(v.reshape(-1) == 2) + 1)
    assert set(A.sources[1].nods) == set(np.flatnonzero(v.reshape(-1) == 3) + 1 + A.nCells)
    f0, _ = A.source_scalars(0.0)
    assert f0[0] > 0 > f0[1] and abs(f0[0] + f0[1]) < 1e-12 * abs(f0[0])


def test_synthetic_deck_runs_through_the_oracle(tmp_path):
    """A user deck that is not one of the three shipped ones goes from .vxc text to converged timesteps
    on the CPU oracle (loader + expression evaluator + assembly + BiCGSTABwr)."""
    from eddy_currents_3d_b200 import load_vxc
    from oracle import oracle
    path, _ = _synthetic_deck(tmp_path, "ASCII_READABLE")
    p = load_vxc(path)
    run = oracle.OracleRun(p)
    assert run.A.rc == 0 and run.A.num_nz == run.A.irow[-1] - 1
    its = []
    for _ in range(2):
        f, v = p.source_scalars(run.T)
        its.append(run.step(f, v))
    assert all(0 < it <= p.itmax for it in its), its          # converged (the reference runs itmax+1 otherwise)
    assert np.isfinite(run.Uaf).all() and np.isfinite(run.Jaf).all() and np.abs(run.Uaf).max() > 0
