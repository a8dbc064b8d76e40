"""Wall clock of the drop-in CLI path on a synthetic coronene-size case (BASELINE.json configs[3]): real MOL / XDENS /
gimic.inp files on disk -> `Driver(gimic.inp).run()` -> jvec.vti, jmod.vti, acid.vti in the reference's ASCII formats.
Phases: parse MOL + XDENS text and upload; tensors + fields on the GPU; writing the files.
    python tools/driver_e2e.py [--natoms 42] [--grid 128] [--vtk ascii|appended]"""
import argparse, io, json, os, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from gimic_b200 import synthetic, writers
from gimic_b200.driver import Driver

ap = argparse.ArgumentParser()
ap.add_argument("--natoms", type=int, default=42)
ap.add_argument("--grid", type=int, default=128)
ap.add_argument("--vtk", default="ascii")
ap.add_argument("--outdir", default=None)
a = ap.parse_args()

sh, dens, nbf = synthetic.synthetic_case(a.natoms, "flake", seed=1234)
d = a.outdir or tempfile.mkdtemp()
os.makedirs(d, exist_ok=True)
# MOL (INTGRL format, intgrl.f90): one block per contraction, all atoms carbon-like
with open(os.path.join(d, "MOL"), "w") as f:
    f.write("INTGRL        1    0    1    0    0    0    0    0    0\nCFOUR\n      synthetic flake\n")
    f.write(f"{a.natoms:5d}    0            0.10E-08              0    0\n9999.00      3.00\n")
    for i, xyz in enumerate(sh["coords"]):
        byl = {}
        for l, xp, cc in synthetic.C_TZVP:
            byl.setdefault(l, []).append((xp, cc))
        f.write(f"6.0    1 {len(byl)} " + " ".join(str(len(byl[l])) for l in sorted(byl)) + "\n")
        f.write(f"C {1:1d}" + "".join(f"{v:20.12f}" for v in xyz) + "\n")
        for l in sorted(byl):
            for xp, cc in byl[l]:
                f.write(f"{len(xp):6d}{1:4d}\n")
                for x, c in zip(xp, cc):
                    f.write(f"{x:20.10f}{c:16.10f}\n")
t0 = time.perf_counter()
with open(os.path.join(d, "XDENS"), "wb") as f:
    f.write(writers.format_e(synthetic.dens_to_colmajor(dens), 22, 14, 1))
t_x = time.perf_counter() - t0
origin, basv, pts = synthetic.box_grid(sh["coords"], (a.grid,) * 3)
origin = [float(v) for v in origin]
L = [float(p[-1]) for p in pts]
with open(os.path.join(d, "gimic.inp"), "w") as f:
    f.write(f"""calc=cdens
basis="MOL"
xdens="XDENS"
openshell=false
magnet=[0.0, 0.0, 1.0]
Grid(base) {{
    type=even
    origin=[{origin[0]!r}, {origin[1]!r}, {origin[2]!r}]
    ivec=[1.0, 0.0, 0.0]
    jvec=[0.0, 1.0, 0.0]
    lengths=[{L[0]!r}, {L[1]!r}, {L[2]!r}]
    grid_points=[{a.grid}, {a.grid}, {a.grid}]
}}
Advanced {{
    spherical=off
    diamag=on
    paramag=on
    GIAO=on
    screening=on
    screening_thrs=1.d-8
}}
Essential {{
    acid=on
    jmod=on
}}
""")
res = {"natoms": a.natoms, "nbf": nbf, "grid": a.grid, "vtk": a.vtk, "xdens_text_MB": os.path.getsize(os.path.join(d, "XDENS")) / 1e6,
       "write_xdens_s": t_x}
for rep in range(2):
    t0 = time.perf_counter()
    Driver(os.path.join(d, "gimic.inp"), out=io.StringIO(), vtk_appended=(a.vtk == "appended")).run()
    t1 = time.perf_counter()
    res[f"rep{rep}"] = {"run_parse_upload_compute_write_s": t1 - t0}
res["files_MB"] = {n: os.path.getsize(os.path.join(d, n)) / 1e6 for n in sorted(os.listdir(d)) if n.endswith(".vti")}
print(json.dumps(res, indent=1))
