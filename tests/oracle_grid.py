"""Test infrastructure: the ORACLE's grid (oracle/grid code behind tests/oracle_lib.py) for a parsed gimic.inp (tests/inp_reader.py)."""
import numpy as np

import oracle_lib as O


def oracle_grid(I, coords):
    """I: inp_reader.Input; coords: (natoms, 3) geometry of the MOL file"""
    G = lambda k: I.get("Grid." + k)
    S = lambda k: I.is_set("Grid." + k)
    kw = dict(type=G("type"), gauss_order=G("gauss_order"), grid_points=G("grid_points") if S("grid_points") else None,
              spacing=G("spacing") if S("spacing") else None, rotation=G("rotation") if S("rotation") else None,
              rotation_origin=G("rotation_origin") if S("rotation_origin") else None)
    if I.grid_arg == "bond":
        if S("bond"):
            b = G("bond"); c1, c2 = coords[b[0] - 1], coords[b[1] - 1]
        else:
            c1, c2 = np.array(G("coord1")), np.array(G("coord2"))
        fix = coords[G("fixpoint") - 1] if S("fixpoint") else np.array(G("fixcoord"))
        return O.grid_bond(c1, c2, fix, G("distance"), height=G("height"), width=G("width"), radius=G("radius") if S("radius") else None,
                           magnet=I.get("magnet") if I.is_set("magnet") else None, **kw)
    return O.grid_std(G("origin"), G("ivec"), G("jvec"), G("lengths"), **kw)


def same_grid(g, og, atol=1e-11):
    """product Grid (gimic_b200.driver.input_grid) against an oracle grid: point counts, points, weights"""
    if list(g.npts) != list(og.npts):
        return False
    ok = np.allclose(g.points().reshape(-1, 3), og.points(), rtol=0, atol=atol, equal_nan=True)      # one point along an axis: l/0 in the reference too
    for d in range(3):
        p, w = og.axis(d)
        ok = ok and np.allclose(g.pts[d], p, rtol=0, atol=atol, equal_nan=True) and np.allclose(g.wgt[d], w, rtol=1e-12, atol=1e-14, equal_nan=True)
    return bool(ok)
