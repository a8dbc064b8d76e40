"""Wall-clock of the native driver on the reference's open-shell 3d case (33^3 points, UHF, 8 output files) and c4h4 integration."""
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
        t0 = time.perf_counter(); drv = Driver(os.path.join(d, "gimic.inp"), out=io.StringIO()); t1 = time.perf_counter()
        drv.run(); t2 = time.perf_counter()
        print(f"{inp:26s} rep {rep}: setup (parse MOL/XDENS, upload) {t1 - t0:6.3f} s, run (compute + write files) {t2 - t1:6.3f} s")
        drv.g.close()
