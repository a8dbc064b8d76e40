#!/usr/bin/env python
"""Randomized check of the compiled driver's input / grid layer on dry runs (no device needed): random gimic.inp files over the grid
keywords (bond / base grids, even / gauss / lobatto, grid_points / spacing, rotation, rotation_origin, radius, magnet / magnet_axis,
coord1/coord2/fixcoord) on the benzene geometry.  `gimic-b200 -y` must accept what an independent reader of the grammar
(tests/inp_reader.py) accepts, and the grid it lays out (gimic_b200_input_grid: points, weights, field direction) must be the
oracle's (oracle grid code, tests/oracle_lib.py).

    python tools/fuzz_dryrun_drivers.py [seed] [cases]

Round 1 ran this as a byte comparison of two drivers (a Python one and the compiled one); its findings are regression-tested:
non-finite coordinates printed as printf's '-nan'; check_field's x > 0 decided by rounding noise for a field in the grid plane."""
import io, os, shutil, subprocess, sys, filecmp, pathlib
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import fixtures
from gimic_b200 import driver
EXE = os.path.join(ROOT, "gimic_b200", "gimic-b200"); GOLD = fixtures.GOLD
rng = None
NATOMS, MAXPTS = 12, 40          # benzene; points per axis (the complete-run variant of this fuzz uses smaller grids)
def num(x): return f"{x:.6g}"
def gen():
    calc = rng.choice(["integral","cdens","integral","cdens","edens","divj"])
    typ = rng.choice(["gauss","even","lobatto"]) if calc=="integral" else rng.choice(["even","gauss"])
    lines = [f"calc={calc}", 'title=""','basis="MOL"','xdens="XDENS"',"debug=1","openshell=false"]
    if rng.random()<0.5: lines.append("magnet_axis="+rng.choice(["X","Y","Z","-x","i","j","k","-k","T"]))
    else: lines.append("magnet=[%s]"%", ".join(num(v) for v in rng.normal(size=3)))
    bond = rng.random()<0.6
    g=[]
    if bond:
        g.append("Grid(bond) {"); g.append(f" type={typ}")
        if rng.random()<0.7:
            a,b = rng.choice(np.arange(1,NATOMS+1),2,replace=False); g.append(f" bond=[{a},{b}]")
        else:
            g.append(" coord1=[%s]"%", ".join(num(v) for v in rng.normal(size=3)*2)); g.append(" coord2=[%s]"%", ".join(num(v) for v in rng.normal(size=3)*2+1))
        if rng.random()<0.7: g.append(f" fixpoint={rng.integers(1,NATOMS+1)}")
        else: g.append(" fixcoord=[%s]"%", ".join(num(v) for v in rng.normal(size=3)*3))
        g.append(f" distance={num(rng.uniform(0.1,2.5))}")
        g.append(f" height=[{num(-rng.uniform(0.5,6))}, {num(rng.uniform(0.5,6))}]"); g.append(f" width=[{num(-rng.uniform(0.5,6))}, {num(rng.uniform(0.5,6))}]")
        if rng.random()<0.3: g.append(f" radius={num(rng.uniform(1,5))}")
    else:
        g.append("Grid(base) {"); g.append(f" type={typ}")
        g.append(" origin=[%s]"%", ".join(num(v) for v in rng.uniform(-8,0,size=3)))
        i = rng.normal(size=3); j = np.cross(i, rng.normal(size=3))
        if rng.random()<0.5: i=np.array([1.,0,0]); j=np.array([0,1.,0])
        g.append(" ivec=[%s]"%", ".join(num(v) for v in i)); g.append(" jvec=[%s]"%", ".join(num(v) for v in j))
        L = rng.uniform(1,10,size=3)
        if rng.random()<0.4: L[2]=0.0
        g.append(" lengths=[%s]"%", ".join(num(v) for v in L))
    if typ!="even": g.append(f" gauss_order={rng.integers(2,12)}")
    if rng.random()<0.6 or typ!="even":
        n3 = rng.integers(0,min(12,MAXPTS)) if not bond else 0
        g.append(f" grid_points=[{rng.integers(1,MAXPTS)}, {rng.integers(1,MAXPTS)}, {n3}]")
    else:
        g.append(" spacing=[%s]"%", ".join(num(v) for v in rng.uniform(0.1,2,size=3)))
    if rng.random()<0.4:
        g.append(" rotation=[%s]"%", ".join(num(v) for v in rng.uniform(-90,90,size=3)))
        if rng.random()<0.5: g.append(" rotation_origin=[%s]"%", ".join(num(v) for v in rng.normal(size=3)))
    g.append("}")
    onoff = lambda: rng.choice(["on", "on", "off"])
    adv = ["Advanced {"," lip_order=5"," spherical=off",f" diamag={onoff()}",f" paramag={onoff()}",f" GIAO={onoff()}",f" screening={onoff()}",
           " screening_thrs=" + rng.choice(["1.d-8", "1e-6", "0.5D-10", "1.0E-12"]),"}"]
    ess = ["Essential {", f" acid={rng.choice(['on','off'])}", f" jmod={rng.choice(['on','off'])}", "}"]
    return "\n".join(lines+g+adv+ess)+"\n"

def run(seed, N, workdir):
    """number of inputs on which the driver disagrees with the independent reader / the oracle grid (directories kept under workdir)"""
    global rng
    import inp_reader, oracle_grid
    rng = np.random.default_rng(seed)
    _, coords = driver.mol_geometry(os.path.join(GOLD, "benzene_MOL"))
    bad = 0
    for k in range(N):
        bad0 = bad
        text = gen()
        d = pathlib.Path(workdir) / f"run{k}"
        shutil.rmtree(d, ignore_errors=True)
        d.mkdir(parents=True); shutil.copy(os.path.join(GOLD, "benzene_MOL"), d / "MOL")
        (d / "gimic.inp").write_text(text)
        p = subprocess.run([EXE, "-y", str(d / "gimic.inp")], capture_output=True, text=True, timeout=60)
        try:
            I = inp_reader.parse_text(text); parsed = True
        except (inp_reader.InputError, ValueError):
            parsed = False
        if not parsed:
            if p.returncode == 0: bad += 1; print("ACCEPTED WHAT THE READER REFUSES", k)
            continue
        try:
            og = oracle_grid.oracle_grid(I, coords)
            axis = I.get("magnet_axis") if I.is_set("magnet_axis") else ""
            ob = og.magnet(axis, I.get("magnet") if I.is_set("magnet") else (0.0, 0.0, 0.0))
        except Exception as e:
            og = None; oerr = repr(e)
        if p.returncode != 0:
            # refusals that are not about the grammar or the geometry: nothing to calculate (gimic.F90:124-130), an axis without points
            nothing = not I.get("Advanced.diamag") and not I.get("Advanced.paramag")
            empty = og is not None and 0 in list(og.npts)
            if og is not None and not nothing and not empty: bad += 1; print("REFUSED", k, p.stderr[-200:])
            continue
        if og is None:
            print("note: oracle refuses what the driver accepts", k, oerr[:120]); continue
        g, b, info = driver.input_grid(str(d / "gimic.inp"))
        if not oracle_grid.same_grid(g, og):
            bad += 1; print("GRID MISMATCH", k, list(g.npts), list(og.npts))
        # a field along an in-plane basis vector of a grid that is not axis-aligned: check_field tests the sign of a dot product of two
        # orthogonal vectors, i.e. of rounding noise (magnet.f90:75), in the reference as well -- either sign is the reference's behaviour
        inplane = axis in ("i", "j")
        if not (np.allclose(b, ob, atol=1e-12) or (inplane and np.allclose(b, -ob, atol=1e-12))):
            bad += 1; print("FIELD MISMATCH", k, axis, b, ob)
        if "NaN" not in p.stdout and ("nan" in p.stdout or "nan" in open(d / "grid.xyz").read()):
            bad += 1; print("PRINTF NAN", k)
        if bad == bad0: shutil.rmtree(d, ignore_errors=True)
    print("cases", N, "mismatches", bad)
    return bad


if __name__ == "__main__":
    import tempfile
    with tempfile.TemporaryDirectory() as tmp:
        sys.exit(1 if run(int(sys.argv[1]) if len(sys.argv) > 1 else 0, int(sys.argv[2]) if len(sys.argv) > 2 else 60, tmp) else 0)
