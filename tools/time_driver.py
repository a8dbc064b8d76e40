"""Wall-clock of the driver (library entry through the Python launcher; --native: the gimic-b200 program) on the reference's open-shell 3d case (33^3 points, UHF, 8 output files) and c4h4 integration."""
import io, os, shutil, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import fixtures
from gimic_b200.driver import Driver
tmp = tempfile.mkdtemp(); cases = fixtures.materialize(tmp)
for case, inp in (("open_shell", "open-shell_3d"), ("c4h4", "c4h4_integration"), ("open_shell", "open-shell_integration")):
    d = cases[case]["dir"]
    shutil.copy(os.path.join(fixtures.GOLD, "inputs", inp + ".inp"), os.path.join(d, "gimic.inp"))
    for rep in range(2):
        t0 = time.perf_counter(); Driver(os.path.join(d, "gimic.inp"), out=io.StringIO()).run(); t1 = time.perf_counter()
        print(f"{inp:26s} rep {rep}: gimic_b200_run in a warm process (parse MOL/XDENS, upload, compute, write files) {t1 - t0:6.3f} s")

if "--native" in sys.argv:
    # the gimic-b200 program: whole-process wall clock (CUDA context creation + MOL/XDENS parse + compute + files), the number to put
    # beside the reference's `gimic gimic.inp` timings (c4h4/integration 2.13 s, open-shell/integration 10.04 s in its golden stdout)
    import subprocess
    exe = os.path.join(ROOT, "gimic_b200", "gimic-b200")
    for case, inp in (("open_shell", "open-shell_3d"), ("c4h4", "c4h4_integration"), ("open_shell", "open-shell_integration")):
        d = cases[case]["dir"]
        shutil.copy(os.path.join(fixtures.GOLD, "inputs", inp + ".inp"), os.path.join(d, "gimic.inp"))
        for rep in range(2):
            t0 = time.perf_counter()
            p = subprocess.run([exe, os.path.join(d, "gimic.inp")], capture_output=True, text=True)
            t1 = time.perf_counter()
            print(f"{inp:26s} rep {rep}: gimic-b200 process wall clock {t1 - t0:6.3f} s (rc {p.returncode}, {len(p.stdout)} bytes of report)")
