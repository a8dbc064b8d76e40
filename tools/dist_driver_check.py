"""torchrun --nproc-per-node N tools/dist_driver_check.py : the compiled driver in rank mode under torch.distributed (NCCL).
cdens (open-shell 3d: an equal-cost share of the tiles per rank, rows gathered on rank 0, files written by rank 0) and integral
(c4h4: plane rows per rank, one all-reduce) compared with the reference goldens on rank 0."""
import io, json, os, re, shutil, sys, tempfile
import numpy as np, torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import fixtures
from make_golden import read_vti
from gimic_b200.driver import Driver
rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1)); lr = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(lr)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
base = os.path.join(tempfile.gettempdir(), "gimic_b200_dist_check")
if rank == 0:
    shutil.rmtree(base, ignore_errors=True); os.makedirs(base)
    fixtures.materialize(base)
if world > 1:
    dist.barrier()
ok = {}
# ---- cdens, open shell
d = os.path.join(base, "open_shell")
if rank == 0:
    shutil.copy(os.path.join(fixtures.GOLD, "inputs", "open-shell_3d.inp"), os.path.join(d, "gimic.inp"))
if world > 1:
    dist.barrier()
Driver(os.path.join(d, "gimic.inp"), out=io.StringIO(), device=lr).run()
if rank == 0:
    gold = fixtures.golden_npz("open_shell_3d.npz"); idx = gold["index"]
    good = True
    for tag in ("", "alpha", "beta", "spindens"):
        jv = read_vti(os.path.join(d, f"jvec{tag}.vti")); gj = gold["jvec" + tag]
        good &= bool((np.abs(jv[idx] - gj) <= 2e-6 * np.abs(gj) + 1e-9 * np.abs(gj).max()).all())
    ok["cdens_open_shell_files_match_golden"] = good
# ---- integral, c4h4
d = os.path.join(base, "c4h4")
if rank == 0:
    shutil.copy(os.path.join(fixtures.GOLD, "inputs", "c4h4_integration.inp"), os.path.join(d, "gimic.inp"))
if world > 1:
    dist.barrier()
out = io.StringIO()
drv = Driver(os.path.join(d, "gimic.inp"), out=out, device=lr); drv.run()
if rank == 0:
    blk = fixtures.golden_json("c4h4_integration.json")["blocks"][1]
    txt = out.getvalue()
    txt = txt[txt.index("*** Integrating current"):]
    r = [float(re.search(k + r"\s*([-\d.]+)", txt).group(1)) for k in (r"Induced current \(au\)\s+:", "Positive contribution:", "Negative contribution:")]
    ok["integral_matches_golden"] = bool(abs(r[0] - blk["au"]) < 1.01e-6 and abs(r[1] - blk["pos"]) < 1.01e-6 and abs(r[2] - blk["neg"]) < 1.01e-6)
    ok["world"] = world
    print(json.dumps(ok))
    assert all(v for k, v in ok.items() if k != "world")
if world > 1:
    dist.destroy_process_group()
