"""Host-side evaluator for the reference's runtime expressions (harness row N1).

Restates the semantics of the reference's function parser (src/m_fparser.f90): the recursive
operator split of ``CompileSubstr`` (m_fparser.f90:570-672) -- operators are searched in the order
``+ - * / ^`` and, for each, from the RIGHT at parenthesis level 0, so ``a*b/c`` evaluates as
``a*(b/c)`` and ``a+b-c`` as ``a+(b-c)`` -- and the function table of ``evalf``
(m_fparser.f90:158-241).  The per-step source scalars it produces are INPUTS to the hot path (the
same doubles are fed to the oracle and to the GPU path), so last-bit differences between this
evaluator and gfortran's ``cosd``/``sind`` cannot leak into parity results.

Also restates ``numeric`` (src/utilites.f90:343-475): SPICE-style decimal prefixes.
"""
from __future__ import annotations

import math
import re
from typing import Dict

_OPS = "+-*/^"


def _completely_enclosed(F: str, b: int, e: int) -> bool:
    """m_fparser.f90:543-566 -- is F[b..e] of the form '(...)' with matching outer parens."""
    if b > e or F[b] != "(" or F[e] != ")":
        return False
    k = 0
    for j in range(b + 1, e):
        if F[j] == "(":
            k += 1
        elif F[j] == ")":
            k -= 1
        if k < 0:
            return False
    return k == 0


def _is_binary_op(j: int, F: str) -> bool:
    """m_fparser.f90:674-722."""
    if F[j] in "+-":
        if j == 0:
            return False
        if F[j - 1] in "+-*/^(":
            return False
        if j + 1 < len(F) and F[j + 1] in "0123456789" and F[j - 1] in "eEdD":
            dflag = pflag = False
            k = j - 1
            while k > 0:
                k -= 1
                if F[k] in "0123456789":
                    dflag = True
                elif F[k] == ".":
                    if pflag:
                        break
                    pflag = True
                else:
                    break
            # exponent sign iff digits were seen and the mantissa starts the string / follows an op
            if dflag and (k == 0 or F[k] in "+-*/^("):
                return False
    return True


def _sind(x: float) -> float:
    return math.sin(math.radians(math.fmod(x, 360.0)))


def _cosd(x: float) -> float:
    return math.cos(math.radians(math.fmod(x, 360.0)))


def _tand(x: float) -> float:
    return math.tan(math.radians(math.fmod(x, 360.0)))


class _EvalError(Exception):
    """An evaluation error of the reference's evalf (EvalErrType > 0): the WHOLE expression evaluates to 0
    (m_fparser.f90:182,187,210-214: `res=zero; RETURN`), not just the failing sub-expression."""


def _lg(y):
    if y <= 0.0:                       # m_fparser.f90:187  EvalErrType=3
        raise _EvalError("lg of a non-positive number")
    return math.log10(y)


def _ln(y):                            # m_fparser.f90:189: plain LOG, no check
    return math.log(y) if y > 0.0 else (-math.inf if y == 0.0 else math.nan)


def _sqrt(y):                          # m_fparser.f90:190: plain DSQRT, no check
    return math.sqrt(y) if y >= 0.0 else math.nan


def _asin(y):
    if y < -1.0 or y > 1.0:            # m_fparser.f90:210-211  EvalErrType=4
        raise _EvalError("asin argument outside [-1, 1]")
    return math.asin(y)


def _acos(y):
    if y < -1.0 or y > 1.0:            # m_fparser.f90:213-214  EvalErrType=4
        raise _EvalError("acos argument outside [-1, 1]")
    return math.acos(y)


_FUNCS = {
    "abs": abs,
    "exp": math.exp,
    "lg": _lg,
    "ln": _ln,
    "sqrt": _sqrt,
    "sh": math.sinh,
    "ch": math.cosh,
    "th": lambda y: math.sinh(y) / math.cosh(y),
    "cth": lambda y: math.cosh(y) / math.sinh(y),
    "sind": _sind,
    "cosd": _cosd,
    "tgd": _tand,
    "sin": math.sin,
    "cos": math.cos,
    "tg": math.tan,
    "asin": _asin,
    "acos": _acos,
    "impls": lambda y: 1.0 if y > 0.0 else 0.0,
    "impl2": lambda y: 1.0 if y >= 0.0 else -1.0,
    "pos": lambda y: y if y > 0.0 else 0.0,
    "int": lambda y: float(math.trunc(y)),
    "nint": lambda y: float(math.floor(abs(y) + 0.5)) * (1.0 if y >= 0 else -1.0),
    "floor": lambda y: float(math.floor(y)),
    "ceil": lambda y: float(math.ceil(y)),
    "atg": math.atan,
}
# longest names first, as MathFunctionIndex matches a prefix of the substring
_FUNC_NAMES = sorted(_FUNCS, key=len, reverse=True)


def _math_function(sub: str):
    low = sub.lower()
    # m_fparser.f90:386-406 walks the table in declaration order and takes the first prefix match;
    # 'sind'/'cosd' precede 'sin'/'cos' there, and no other name is a prefix of an earlier one
    # except 'sh'..; longest-first is equivalent for this table.
    for name in _FUNC_NAMES:
        if low.startswith(name + "("):
            return name
    return None


def _eval(F: str, b: int, e: int, var: Dict[str, float]) -> float:
    if F[b] == "+":
        return _eval(F, b + 1, e, var)
    if _completely_enclosed(F, b, e):
        return _eval(F, b + 1, e - 1, var)
    if F[b].isalpha():
        name = _math_function(F[b:e + 1])
        if name is not None:
            b2 = b + F[b:e + 1].index("(")
            if _completely_enclosed(F, b2, e):
                return float(_FUNCS[name](_eval(F, b2 + 1, e - 1, var)))
    elif F[b] == "-":
        if _completely_enclosed(F, b + 1, e):
            return -_eval(F, b + 2, e - 1, var)
        if F[b + 1].isalpha():
            name = _math_function(F[b + 1:e + 1])
            if name is not None:
                b2 = b + 1 + F[b + 1:e + 1].index("(")
                if _completely_enclosed(F, b2, e):
                    return -float(_FUNCS[name](_eval(F, b2 + 1, e - 1, var)))
    for op in _OPS:
        k = 0
        for j in range(e, b - 1, -1):
            c = F[j]
            if c == ")":
                k += 1
            elif c == "(":
                k -= 1
            if k == 0 and c == op and _is_binary_op(j, F):
                if op in "*/^" and F[b] == "-":
                    return -_eval(F, b + 1, e, var)
                lhs = _eval(F, b, j - 1, var)
                rhs = _eval(F, j + 1, e, var)
                if op == "+":
                    return lhs + rhs
                if op == "-":
                    return lhs - rhs
                if op == "*":
                    return lhs * rhs
                if op == "/":
                    if rhs == 0.0:
                        raise _EvalError("division by zero")   # EvalErrType=1, res=zero; RETURN (m_fparser.f90:182)
                    return lhs / rhs
                return lhs ** rhs
    b2 = b + 1 if F[b] == "-" else b
    item = F[b2:e + 1]
    if item in var:
        val = float(var[item])
    else:
        val = float(item.replace("D", "E").replace("d", "e"))
    return -val if b2 > b else val


def evalf(expr: str, var: Dict[str, float]) -> float:
    """parsef + evalf for one expression; variable names are matched exactly (the reference
    upper-cases the whole <Name> string before parsing, vxc2data.f90:439)."""
    F = expr.replace("**", "^").replace(" ", "").replace("\t", "")
    if not F:
        raise ValueError("empty expression")
    try:
        return _eval(F, 0, len(F) - 1, var)
    except _EvalError:
        return 0.0


_PREFIXES = [("M", None), ("K", 1e3), ("U", 1e-6), ("N", 1e-9), ("P", None), ("G", 1e9), ("T", 1e12),
             ("F", 1e-15), ("C", 1e-2)]  # 'H' is declared but the loop stops at 9 (utilites.f90:390)


def numeric(sa: str) -> float:
    """utilites.f90:343-475 -- string to number with decimal prefixes ('100M' -> 1e-3*100.)."""
    s = sa.upper().replace(",", ".", 1)
    mult, lm, l, hit = 1.0, -1, -1, False
    for sym, m in _PREFIXES:
        l = s.find(sym)
        if l >= 0:
            hit = True
            if sym == "M":
                lm = s.find("MEG")
                mult = 1e6 if lm >= 0 else 1e-3
            elif sym == "P":
                lm = s.find("PET")
                mult = 1e15 if lm >= 0 else 1e-12
            else:
                mult = m
            break
    if hit and "." not in s:
        s = s[:l] + "." + s[l + 1:]
        if lm >= 0:
            s = s[:lm + 1] + s[lm + 3:]
    elif lm >= 0:
        s = s[:lm] + s[lm + 3:]
    if "E" not in s:
        s = "".join(ch if (ch.isdigit() or ch in ".-") else " " for ch in s)
    s = s.replace(" ", "")  # READ(...,'(BN,G20.0)'): blanks ignored
    if s in ("", ".", "-", "-."):
        return mult * 0.0
    m = re.match(r"^[-+]?(\d+\.?\d*|\.\d+)([ED][-+]?\d+)?", s)
    if not m:
        raise ValueError(f"numeric(): cannot read {sa!r}")
    return mult * float(m.group(0).replace("D", "E"))


# constants usable inside quoted expressions (vxc2data.f90:398-411); literal values as written there
def constants(dt: float, delta, time: float, sdx: int, sdy: int, sdz: int) -> Dict[str, float]:
    return {
        "PI": 3.1415926535897932384626433832795,
        "E": 0.27182818284590451e+001,
        "MU0": 0.12566370964050292e-005,
        "E0": 0.88541878176203908e-011,
        "DT": dt, "DX": float(delta[0]), "DY": float(delta[1]), "DZ": float(delta[2]),
        "TIME": time, "NX": float(sdx), "NY": float(sdy), "NZ": float(sdz),
    }


def evaluate(word: str, consts: Dict[str, float]) -> float:
    """vxc2data.f90:822-834 -- quoted => expression over the constants, else numeric()."""
    if word[:1] in "\"'`":
        return evalf(word[1:len(word.rstrip()) - 1], consts)
    return numeric(word)
