"""Minimal VoxCad ``.vxc`` importer for the harness.

Restates what the reference's ``vxc2data`` (src/vxc2data.f90:7-893) extracts from a deck -- grid,
spacing, materials, conductor numbering, source node lists, source / motion functions, solver and
transient parameters -- into a :class:`Problem`.  Voxel layers are decoded directly from the
base64+zlib payload (the reference shells out to src/uncompress_zlib.py for that).

Host-side text parsing only; nothing here is on the hot path.
"""
from __future__ import annotations

import base64
import re
import zlib
from typing import Dict, List

import numpy as np

from . import fparser
from .problem import Problem, Source, number_conductor, renumber_air, source_nodes

_LETTER = "123456789:;<=>?@ABCDEFGHIJKLMNOPQRSTUVWXYZ[\\]^_`abcdefghijklmnopqrstuvwxyz"


def _tag(text: str, name: str):
    m = re.search(rf"<{name}>(.*?)</{name}>", text)
    return m.group(1) if m else None


def _words(name_string: str) -> List[str]:
    # vxc2data.f90:127-142: every '=' becomes a blank, upper-case, split on blanks/tabs
    return name_string.replace("=", " ").upper().split()


class _Fn:
    def __init__(self, name: str, nomsch: int, ex: str):
        self.name, self.nomsch, self.ex = name, nomsch, ex
        self.eqn = " "
        self.namex: List[str] = []
        self.velx: List[float] = []

    def value(self, t: float) -> float:
        # EC3D.f90:247-254: arguments literally named 'T' follow the simulation time
        var: Dict[str, float] = {}
        for nm, vv in zip(self.namex, self.velx):
            var[nm] = t if nm == "T" else vv
        return fparser.evalf(self.eqn, var)


def load_vxc(path: str) -> Problem:
    with open(path, "r", errors="replace") as fh:
        text = fh.read()
    lines = text.splitlines()

    delta0 = fparser.numeric(_tag(text, "Lattice_Dim").strip())
    delta = np.array([fparser.numeric(_tag(text, f"{a}_Dim_Adj").strip()) * delta0 for a in "XYZ"],
                     np.float64)                                            # vxc2data.f90:94-121
    sdx, sdy, sdz = (int(_tag(text, f"{a}_Voxels")) for a in "XYZ")         # :229-248
    nC = sdx * sdy * sdz
    zlib_mode = 'Structure Compression="ZLIB"' in text
    layers = re.findall(r"<Layer><!\[CDATA\[(.*?)\]\]></Layer>", text)
    if len(layers) < sdz:
        raise ValueError("not enough <Layer> records")
    v = np.zeros(nC, np.int64)
    pos = 0
    for lay in layers[:sdz]:
        if zlib_mode:
            raw = np.frombuffer(zlib.decompress(base64.b64decode(lay)), np.uint8).astype(np.int64)
        else:                                                               # :298-311
            raw = np.array([_LETTER.find(ch) + 1 for ch in lay.strip()], np.int64)
        v[pos:pos + raw.size] = raw
        pos += raw.size

    names = [ln[ln.index("<Name>") + 6: ln.index("</", ln.index("<Name>"))] for ln in lines
             if "<Name>" in ln]

    # ---- pass 1 (vxc2data.f90:127-222): TRAN / SOLVER keywords ----
    Time = dt = dtt = 0.0
    solv, tolerance, itmax, bound, files = "BCG", 1e-3, 10000, "DDDDDD", "out"  # :74
    for nm in names:
        w = _words(nm)
        for i in range(1, len(w)):
            if w[i] == "TRAN":
                for j in range(i + 1, len(w) - 1, 2):
                    if "STOP" in w[j]:
                        Time = fparser.numeric(w[j + 1])
                    elif "STEP" in w[j]:
                        dt = fparser.numeric(w[j + 1])
                    elif "JUMP" in w[j]:
                        dtt = fparser.numeric(w[j + 1])
            if w[i] == "SOLVER":
                for j in range(i + 1, len(w) - 1):
                    if "TOL" in w[j]:
                        tolerance = fparser.numeric(w[j + 1])
                    elif "ITMAX" in w[j]:
                        itmax = int(round(fparser.numeric(w[j + 1])))
                    elif "SOLV" in w[j]:
                        solv = w[j + 1][:3]
                    elif "DIR" in w[j]:
                        files = w[j + 1]
                    elif "BOUND" in w[j]:
                        bound = w[j + 1][:6]

    nsub = int(v.max())
    v, nsub_air = renumber_air(v, nsub)                                     # :320-338
    nsub_glob = nsub + nsub_air
    valPHYS = np.zeros((nsub_glob, 5), np.float64)
    valPHYS[nsub:, 0] = 1.0
    consts = fparser.constants(dt, delta, Time, sdx, sdy, sdz)
    BND = np.full((3, 2), -0.95, np.float64)

    cond_numdom: List[int] = []
    cond_valdom: List[float] = []
    funs: List[_Fn] = []
    mechs: List[_Fn] = []
    fun_move, fun_numv, fun_velv = [], [], []

    def calc_vmech(ch: str, w: List[str], j: int, kp: int):                 # :836-891
        f = _Fn(w[j + 1], kp, ch)
        funs.append(f)
        move, numv, velv = np.zeros(3, np.int32), np.zeros(3, np.int32), np.zeros(3, np.float64)
        for n in range(1, 7):
            if j + 1 + n + 1 <= len(w) - 1:
                key, val = w[j + 1 + n], w[j + 1 + n + 1]
                for ax, tagname, exm in ((0, "VSX", "X"), (1, "VSY", "Y"), (2, "VSZ", "D")):
                    if tagname in key:
                        move[ax] = 1
                        if "A" <= val[:1] <= "Z":
                            mechs.append(_Fn(val, kp, exm))
                            numv[ax] = len(mechs)
                            velv[ax] = 0.0
                        else:
                            numv[ax] = 0
                            velv[ax] = fparser.evaluate(val, consts)
                        break
        fun_move.append(move); fun_numv.append(numv); fun_velv.append(velv)

    def fill_args(fn: _Fn, w: List[str], i: int):                           # :497-548
        fn.eqn = w[i + 2]
        fn.namex, fn.velx = [], []
        j = i + 3
        while j + 1 <= len(w) - 1:
            fn.namex.append(w[j][:8])
            fn.velx.append(fparser.evaluate(w[j + 1], consts))
            j += 2

    # ---- pass 2 (vxc2data.f90:420-600) ----
    for kp, nm in enumerate(names, start=1):
        w = [""] + _words(nm)          # 1-based like the Fortran words(:)
        neww = len(w) - 1
        for i in range(2, neww + 1):
            if w[i][:1] == "D" and kp <= nsub:
                valPHYS[kp - 1, 0] = fparser.evaluate(w[i + 1], consts)
                if i + 1 == neww:
                    continue
                for j in range(i + 2, neww):
                    if w[j][:1] == "C":
                        valPHYS[kp - 1, 1] = fparser.evaluate(w[j + 1], consts)
                        if valPHYS[kp - 1, 1] != 0.0:
                            cond_numdom.append(kp)
                            cond_valdom.append(2.0 * valPHYS[kp - 1, 1] / dt)   # :461
                    elif "VEX" in w[j]:
                        valPHYS[kp - 1, 2] = fparser.evaluate(w[j + 1], consts)
                    elif "VEY" in w[j]:
                        valPHYS[kp - 1, 3] = fparser.evaluate(w[j + 1], consts)
                    elif "VEZ" in w[j]:
                        valPHYS[kp - 1, 4] = fparser.evaluate(w[j + 1], consts)
                if i + 2 <= neww and "SRC" in w[i + 2]:
                    for j in range(i + 2, neww):
                        if "SRCY" in w[j]:
                            calc_vmech("Y", w, j, kp)
                        elif "SRCX" in w[j]:
                            calc_vmech("X", w, j, kp)
                        elif "SRCZ" in w[j]:
                            calc_vmech("D", w, j, kp)
            elif "FUNC" in w[i]:
                for fn in funs:
                    if w[i + 1] == fn.name:
                        fill_args(fn, w, i)
                for fn in mechs:
                    if w[i + 1] == fn.name:
                        fill_args(fn, w, i)
            elif "BOUNDARY" in w[i]:
                for j in range(i + 1, neww, 2):
                    key = w[j][:3]
                    val = fparser.evaluate(w[j + 1], consts)
                    slot = {"BXM": (0, 0), "BXP": (0, 1), "BYM": (1, 0), "BYP": (1, 1),
                            "BZM": (2, 0), "BZP": (2, 1)}
                    if key in slot:
                        BND[slot[key]] = val
                    elif key == "ALL":
                        BND[:, :] = val
                    else:
                        raise ValueError("not recognized BOUNDARY " + w[j])
            elif "ENVIRON" in w[i]:
                for j in range(i + 1, neww):
                    if w[j][:1] == "D":
                        valPHYS[nsub_glob - 1, 0] = fparser.evaluate(w[j + 1], consts)
                    elif w[j][:1] == "C":
                        valPHYS[nsub_glob - 1, 1] = fparser.evaluate(w[j + 1], consts)
                        if valPHYS[nsub_glob - 1, 1] != 0.0:
                            cond_numdom.append(nsub_glob)
                            cond_valdom.append(2.0 * valPHYS[nsub_glob - 1, 1] / dt)
                    elif "VEX" in w[j]:
                        valPHYS[nsub_glob - 1, 2] = fparser.evaluate(w[j + 1], consts)
                    elif "VEY" in w[j]:
                        valPHYS[nsub_glob - 1, 3] = fparser.evaluate(w[j + 1], consts)
                    elif "VEZ" in w[j]:
                        valPHYS[nsub_glob - 1, 4] = fparser.evaluate(w[j + 1], consts)

    geoPHYS = v.astype(np.int8)                                             # :604-606
    if cond_numdom and solv == "BCG" and ("A" in bound or "N" in bound):    # :611-622
        g3 = v.reshape(sdz, sdy, sdx)
        face = np.zeros_like(g3, bool)
        face[0], face[-1], face[:, 0], face[:, -1], face[:, :, 0], face[:, :, -1] = (True,) * 6
        for np_ in cond_numdom:
            g3[face & (g3 == np_)] = nsub_glob
        v = g3.reshape(-1)
    geoC, nod = number_conductor(v, cond_numdom, nC)                        # :624-652

    sources = []
    for f, mv, nv, vv in zip(funs, fun_move, fun_numv, fun_velv):           # :656-752
        sources.append(Source(name=f.name, ex=f.ex, nomsch=f.nomsch,
                              nods=source_nodes(geoPHYS.astype(np.int64), f.nomsch, f.ex, nC),
                              move=mv, num_Vmech=nv, vel_Vmech=vv))

    def evaluate_functions(t: float):
        return (np.array([f.value(t) for f in funs], np.float64),
                np.array([m.value(t) for m in mechs], np.float64))

    p = Problem(sdx=sdx, sdy=sdy, sdz=sdz, delta=delta, dt=dt, Time=Time, BND=BND,
                tolerance=tolerance, itmax=itmax, geoPHYS=geoPHYS, geoPHYS_C=geoC, valPHYS=valPHYS,
                cond_numdom=cond_numdom, cond_nod=nod,
                cond_valdom=np.array(cond_valdom, np.float64), sources=sources, numMech=len(mechs),
                evaluate_functions=evaluate_functions, name=path.split("/")[-1])
    p.solv, p.files, p.bound, p.dtt = solv, files, bound, dtt
    return p
