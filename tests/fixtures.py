"""Shared test inputs: reference test cases (from tests/golden) and seeded synthetic molecules."""
import json
import os
import numpy as np

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def write_xdens(path, vals):
    """One number per line like the reference's XDENS (dens.f90:129-135); %.14E round-trips the file's digits."""
    with open(path, "w") as f:
        f.write("\n".join("%.14E" % v for v in vals))
        f.write("\n")


def materialize(tmp):
    out = {}
    for name in ("c4h4", "open_shell"):
        d = os.path.join(str(tmp), name)
        os.makedirs(d, exist_ok=True)
        mol = os.path.join(d, "MOL")
        with open(os.path.join(GOLD, f"{name}_MOL")) as f, open(mol, "w") as g:
            g.write(f.read())
        xd = os.path.join(d, "XDENS")
        write_xdens(xd, np.load(os.path.join(GOLD, f"{name}_xdens.npz"))["xdens"])
        out[name] = dict(mol=mol, xdens=xd, dir=d)
    out["benzene_mol"] = os.path.join(GOLD, "benzene_MOL")
    return out


_CLOCK = None


def strip_clock(text):
    """a report with the run-dependent part of its trailer (timer.f90:13-43: wall / user / sys times and the date) blanked, so that
    two runs can be compared byte for byte"""
    global _CLOCK
    import re
    if _CLOCK is None:
        _CLOCK = (re.compile(r"^( {4}wall time:| {9}user:| {10}sys:).*$", re.M),
                  re.compile(r"^ \w{3} \w{3} [ \d]\d \d\d:\d\d:\d\d \d{4}$", re.M))
    return _CLOCK[1].sub(" <date>", _CLOCK[0].sub(r"\1", text))


def golden_npz(name):
    return np.load(os.path.join(GOLD, name))


def golden_json(name):
    return json.load(open(os.path.join(GOLD, name)))


from gimic_b200.synthetic import (C_TZVP, NFUNC_C, hex_flake, ring, synthetic_shells, synthetic_density,  # noqa: E402,F401
                                  synthetic_case, dens_to_colmajor, box_grid)
