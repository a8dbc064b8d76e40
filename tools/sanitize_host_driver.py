"""Runs every mode of the native driver through sanitizer builds linked with the test double of the C ABI; see sanitize_host_driver.sh."""
import os, re, shutil, subprocess, sys, tempfile
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, ROOT)
import fixtures

A, T = sys.argv[1], sys.argv[2]
G = fixtures.GOLD; INP = os.path.join(G, "inputs")
base = tempfile.mkdtemp(prefix="gimic_san_")
cases = fixtures.materialize(os.path.join(base, "cases"))
findings = 0


def wd(name, case, edit=None, extra=None):
    d = os.path.join(base, name); os.makedirs(d)
    shutil.copy(cases[case]["mol"] if case != "benzene" else cases["benzene_mol"], d + "/MOL")
    if case == "benzene":
        fixtures.write_xdens(d + "/XDENS", fixtures.dens_to_colmajor(fixtures.synthetic_density(252, seed=21)))
    else:
        shutil.copy(cases[case]["xdens"], d + "/XDENS")
    t = open(os.path.join(INP, name + ".inp")).read()
    open(d + "/gimic.inp", "w").write(edit(t) if edit else t)
    if extra:
        extra(d)
    return d


def run(exe, args):
    global findings
    p = subprocess.run([exe, *args], capture_output=True, text=True, timeout=1800)
    bad = p.returncode != 0 or "ERROR" in p.stderr or "runtime error" in p.stderr or "WARNING: ThreadSanitizer" in p.stderr
    findings += bad
    print(("FINDING " if bad else "ok      ") + os.path.basename(exe), " ".join(a.replace(base, "") for a in args), "rc", p.returncode)
    if bad:
        print(p.stderr[:3000])


gold = fixtures.golden_npz("c4h4_readgrid.npz")
ELE = "2  4  0\n 1 1 2 3 4\n 2 5 6 7 8\n"


def rg(d):
    np.savetxt(d + "/gridfile.grd", gold["grid"][:200], fmt="%.6f"); open(d + "/grid.1.ele", "w").write(ELE)


def prop(d):
    from gimic_b200.driver import mol_geometry as read_mol_geometry
    _, coords = read_mol_geometry(d + "/MOL")
    rng = np.random.default_rng(5); counts = rng.integers(4, 8, size=coords.shape[0])
    pts = np.vstack([coords[a] + rng.normal(scale=1.5, size=(c, 3)) for a, c in enumerate(counts)])
    np.savetxt(d + "/gridfile.grd", pts, fmt="%.10f"); np.savetxt(d + "/grid_w.grd", rng.uniform(0, 0.1, size=pts.shape[0]), fmt="%.12e")
    shutil.copy(os.path.join(G, "benzene_coord.au"), d + "/coord.au")
    np.savetxt(d + "/nelpts.info", np.column_stack([np.arange(1, len(counts) + 1), counts]), fmt="%d")
    open(d + "/grid.1.ele", "w").write(ELE)


small2 = lambda t: t.replace("grid_points=[30, 30, 0]", "grid_points=[9, 9, 0]")
small3 = lambda t: t.replace("spacing=[0.5, 0.5, 0.5]", "spacing=[4.0, 4.0, 8.0]")
d = wd("c4h4_read-grid", "c4h4", extra=rg)
run(A, [d + "/gimic.inp"]); run(A, ["--devices", "0,1", d + "/gimic.inp"]); run(T, ["--devices", "0,1", d + "/gimic.inp"])
d = wd("c4h4_integration", "c4h4", edit=small2)
run(A, [d + "/gimic.inp"]); run(T, ["--devices", "all", d + "/gimic.inp"])
d = wd("open-shell_3d", "open_shell", edit=small3)
run(A, [d + "/gimic.inp"]); run(A, ["--vtk", "appended", d + "/gimic.inp"]); run(T, ["--devices", "0,1", d + "/gimic.inp"])
d = wd("open-shell_integration", "open_shell", edit=small2); run(A, [d + "/gimic.inp"])
d = wd("benzene_magnetizability", "benzene", extra=prop); run(A, [d + "/gimic.inp"])
d = wd("benzene_3d", "benzene", edit=lambda t: re.sub(r"grid_points=\[\s*\d+\s*,\s*\d+\s*,\s*\d+\s*\]", "grid_points=[4,3,3]", t)); run(A, [d + "/gimic.inp"])
d = wd("benzene_int-cdens", "benzene", edit=small2); run(A, [d + "/gimic.inp"])
d = os.path.join(base, "c4h4_integration"); t = open(d + "/gimic.inp").read(); names = []
for k in range(3):
    open(f"{d}/gimic.{k}.inp", "w").write(t.replace("width=[-1.25614, 6.0]", f"width=[{-1.25614 + 2 * k:.6f}, {-1.25614 + 2 * k + 2:.6f}]"))
    names.append(f"{d}/gimic.{k}.inp")
run(A, names)
print("findings:", findings)
sys.exit(1 if findings else 0)
