#!/usr/bin/env python
"""Randomized comparison of the two drivers above the C ABI on dry runs (no device needed): random gimic.inp files over the grid
keywords (bond / base grids, even / gauss / lobatto, grid_points / spacing, rotation, rotation_origin, radius, magnet / magnet_axis,
coord1/coord2/fixcoord) on the benzene geometry; `gimic-b200 -y` and the Python driver must accept the same inputs and then write
the same report and the same grid.xyz byte for byte.

    python tools/fuzz_dryrun_drivers.py [seed] [cases]

Findings so far (fixed, regression-tested): non-finite coordinates printed as printf's '-nan' by one driver; check_field's x > 0
decided differently for a field lying in the grid plane (BLAS vs plain summation order)."""
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
    """number of inputs on which the drivers disagree (their directories are kept under workdir)"""
    global rng
    rng = np.random.default_rng(seed)
    bad = 0
    for k in range(N):
        bad0=bad
        text=gen()
        base=pathlib.Path(workdir) / f"run{k}"
        shutil.rmtree(base, ignore_errors=True)
        ds=[]
        for s in ("nat","py"):
            d=base/s; d.mkdir(parents=True); shutil.copy(os.path.join(GOLD,"benzene_MOL"), d/"MOL") if os.path.exists(os.path.join(GOLD,"benzene_MOL")) else None
            (d/"gimic.inp").write_text(text); ds.append(d)
        p=subprocess.run([EXE,"-y",str(ds[0]/"gimic.inp")],capture_output=True,text=True,timeout=60)
        out=io.StringIO()
        try:
            driver.Driver(str(ds[1]/"gimic.inp"),out=out,dryrun=True).run(); ok=True; perr=""
        except Exception as e:
            ok=False; perr=repr(e)
        if (p.returncode==0)!=ok:
            bad+=1; print("ACCEPT MISMATCH",k,p.returncode,p.stderr[-200:],perr); continue
        if ok:
            a,b=fixtures.strip_clock(p.stdout),fixtures.strip_clock(out.getvalue())
            if a!=b:
                bad+=1; print("REPORT MISMATCH",k)
                al,bl=a.split("\n"),b.split("\n")
                for x,y in zip(al,bl):
                    if x!=y: print("  nat:",x); print("  py :",y); break
                continue
            fa=sorted(os.listdir(ds[0])); fb=sorted(os.listdir(ds[1]))
            if fa!=fb: bad+=1; print("FILES MISMATCH",k,fa,fb); continue
            for f in fa:
                if not filecmp.cmp(ds[0]/f,ds[1]/f,shallow=False): bad+=1; print("FILE DIFF",k,f)
        if bad==bad0: shutil.rmtree(base, ignore_errors=True)
    print("cases", N, "mismatches", bad)
    return bad


if __name__ == "__main__":
    import tempfile
    with tempfile.TemporaryDirectory() as tmp:
        sys.exit(1 if run(int(sys.argv[1]) if len(sys.argv) > 1 else 0, int(sys.argv[2]) if len(sys.argv) > 2 else 60, tmp) else 0)
