"""Test infrastructure: per-value Python restatements of the Fortran edit descriptors the reference's files use (Ew.d, list-directed
real(8)), and ctypes calls of the library's bulk formatters (gimic_b200_format_e / _f) that the tests hold against them."""
import ctypes as C
import numpy as np


def fortran_e(x, w, d):
    """Fortran Ew.d: 0.dddddE+ee (gfortran drops the 'E' when the exponent needs three digits; asterisks on overflow)"""
    x = float(x)
    if x != x or x in (float("inf"), -float("inf")):          # gfortran: NaN / Infinity / -Infinity, right-justified
        s = "NaN" if x != x else ("-Infinity" if x < 0 else "Infinity")
    elif x == 0.0:
        s = ("-" if str(x)[0] == "-" else "") + "0." + "0" * d + "E+00"      # gfortran keeps the sign of a negative zero
    else:
        m, e = f"{abs(x):.{d - 1}E}".split("E")
        digits = m.replace(".", "")
        e = int(e) + 1
        s = ("-" if x < 0 else "") + "0." + digits + (f"E{e:+03d}" if abs(e) < 100 else f"{e:+04d}")
    return "*" * w if len(s) > w else s.rjust(w)


def ld_real(x):
    """gfortran list-directed real(8): 17 significant digits, F form for 1e-1 <= |x| < 1e16"""
    x = float(x)
    ax = abs(x)
    if x != x or ax == float("inf"):
        return ("NaN" if x != x else ("-Infinity" if x < 0 else "Infinity")).rjust(26)
    if ax != 0.0 and not (0.1 <= ax < 1e16):
        m, e = f"{ax:.16E}".split("E")
        return "  " + ("-" if x < 0 else "") + m + f"E{int(e):+04d}" + " "
    nint = len(str(int(ax))) if ax >= 1.0 else 0
    dec = 17 - nint if ax >= 1.0 else 17
    if ax == 0.0:
        dec = 16                      # gfortran prints zero as 0.0000000000000000 (test/benzene/2d/reference/jvec.vti)
    return (("-" if x < 0 else "") + f"{ax:.{dec}f}").rjust(21) + "     "


def _format(kind, values, w, d, per_line, first=0, prefix=""):
    from gimic_b200 import _lib
    v = np.ascontiguousarray(values, dtype=np.float64).ravel()
    n = v.size
    if n == 0:
        return b""
    cap = n * w + (n // max(per_line, 1) + 2) * (len(prefix) + 1) + 16
    buf = np.empty(cap, dtype=np.uint8)
    fn = _lib.lib().gimic_b200_format_e if kind == "E" else _lib.lib().gimic_b200_format_f
    got = fn(n, v.ctypes.data_as(_lib.dp), w, d, per_line, first, prefix.encode(), C.c_void_p(buf.ctypes.data), cap)
    if got < 0:
        raise _lib.GimicB200Error(got, "number formatting failed")
    return buf[:got].tobytes()


def format_e(values, w, d, per_line, first=0, prefix=""):
    """gimic_b200_format_e: many values with Ew.d, `per_line` per line (`first` on the first line if > 0), every line starting with `prefix`"""
    return _format("E", values, w, d, per_line, first, prefix)


def format_f(values, w, d, per_line, first=0, prefix=""):
    """gimic_b200_format_f: the same for Fw.d"""
    return _format("F", values, w, d, per_line, first, prefix)
