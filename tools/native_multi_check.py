"""gimic-b200 --devices all vs a single device on the reference's runnable cases (needs >= 1 GPU; with one GPU the partition runs as
--devices 0,0): reports and files must agree at print precision; wall clocks of both are printed."""
import filecmp, os, shutil, subprocess, sys, tempfile, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import fixtures
import test_native_driver_gpu as T      # the text comparison helpers

import torch
ndev = torch.cuda.device_count()
devs = "all" if ndev > 1 else "0,0"
exe = os.path.join(ROOT, "gimic_b200", "gimic-b200")
base = tempfile.mkdtemp(prefix="gimic_multi_")
cases = fixtures.materialize(os.path.join(base, "cases"))
gold = fixtures.golden_npz("c4h4_readgrid.npz")
bad = 0
for name, case in (("c4h4_integration", "c4h4"), ("open-shell_3d", "open_shell"), ("c4h4_read-grid", "c4h4"), ("open-shell_integration", "open_shell")):
    dirs = []
    for k in ("one", "many"):
        d = os.path.join(base, k, name); os.makedirs(d)
        shutil.copy(cases[case]["mol"], d + "/MOL"); shutil.copy(cases[case]["xdens"], d + "/XDENS")
        shutil.copy(os.path.join(fixtures.GOLD, "inputs", name + ".inp"), d + "/gimic.inp")
        if name == "c4h4_read-grid":
            np.savetxt(d + "/gridfile.grd", gold["grid"], fmt="%.6f")
        dirs.append(d)
    t0 = time.perf_counter(); one = subprocess.run([exe, dirs[0] + "/gimic.inp"], capture_output=True, text=True, timeout=600); t1 = time.perf_counter()
    many = subprocess.run([exe, "--devices", devs, dirs[1] + "/gimic.inp"], capture_output=True, text=True, timeout=600); t2 = time.perf_counter()
    ok = one.returncode == 0 and many.returncode == 0
    try:
        if ok:
            T._same_text(fixtures.strip_clock(one.stdout), fixtures.strip_clock(many.stdout), "report")
            assert sorted(os.listdir(dirs[0])) == sorted(os.listdir(dirs[1]))
            for f in sorted(os.listdir(dirs[0])):
                a, b = os.path.join(dirs[0], f), os.path.join(dirs[1], f)
                if not filecmp.cmp(a, b, shallow=False):
                    T._same_text(open(a, errors="replace").read(), open(b, errors="replace").read(), f, floor_rel=1e-9)
    except AssertionError as e:
        ok = False; print("MISMATCH", name, str(e)[:400])
    bad += not ok
    print(f"{name:26s} devices={devs} ({ndev} GPU): {'ok' if ok else 'FAILED'}  one device {t1 - t0:6.2f} s, --devices {t2 - t1:6.2f} s  {one.stderr[:200]} {many.stderr[:200]}")
sys.exit(1 if bad else 0)
