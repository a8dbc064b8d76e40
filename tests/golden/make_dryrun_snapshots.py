"""Writes tests/golden/native_dryrun/<input>.{out,mol.xyz,grid.xyz}: what `gimic-b200 -y` prints and writes for every reference input under
tests/golden/inputs.  A characterisation snapshot of this repo's own driver, taken when the round-1 Python driver (which the GPU tests had
pinned to the reference's goldens, and which the compiled driver matched byte for byte on all these inputs) was retired in round 2; the
report preambles are separately held to the reference's stdout goldens (tests/test_native_driver_cpu.py).  Re-run only on purpose:
    python tests/golden/make_dryrun_snapshots.py
"""
import os
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(HERE)); sys.path.insert(0, ROOT)
import fixtures  # noqa: E402
import numpy as np  # noqa: E402

EXE = os.path.join(ROOT, "gimic_b200", "gimic-b200")
OUT = os.path.join(HERE, "native_dryrun")
os.makedirs(OUT, exist_ok=True)
for f in sorted(os.listdir(os.path.join(HERE, "inputs"))):
    name = f[:-4]
    mol = "c4h4_MOL" if name.startswith("c4h4") else "open_shell_MOL" if name.startswith("open") else "benzene_MOL"
    with tempfile.TemporaryDirectory() as d:
        shutil.copy(os.path.join(HERE, mol), os.path.join(d, "MOL"))
        shutil.copy(os.path.join(HERE, "inputs", f), os.path.join(d, "gimic.inp"))
        if "read-grid" in name or "magnetizability" in name:
            np.savetxt(os.path.join(d, "gridfile.grd"), fixtures.golden_npz("c4h4_readgrid.npz")["grid"][:64], fmt="%.6f")
        p = subprocess.run([EXE, "-y", os.path.join(d, "gimic.inp")], capture_output=True, text=True, timeout=60, check=True)
        open(os.path.join(OUT, name + ".out"), "w").write(fixtures.strip_clock(p.stdout))
        for x in ("mol.xyz", "grid.xyz"):
            shutil.copy(os.path.join(d, x), os.path.join(OUT, name + "." + x))
    print(name)
