"""torchrun --nproc-per-node N tools/dist_integral_check.py : integral mode sharded by plane rows over N GPUs,
ONE NCCL all-reduce of the 7 partial sums; rank 0 compares with the single-GPU result and the reference golden."""
import json, os, sys, tempfile
import numpy as np
import torch
import torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import fixtures
import gimic_b200

rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1)); lr = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(lr)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
tmp = tempfile.mkdtemp()
cases = fixtures.materialize(tmp)
g = gimic_b200.Gimic(cases["c4h4"]["mol"], cases["c4h4"]["xdens"], screening_thrs=1e-8, device=lr)
xyz = g.atom_coords()
# bond grid of test/c4h4/integration (geometry from the golden stdout, Gauss points from the product)
gold = fixtures.golden_json("c4h4_integration.json")
geo = gold["geometry"]
p1 = np.zeros(36); w1 = np.zeros(36); p2 = np.zeros(36); w2 = np.zeros(36)
gimic_b200.gausspoints(0.0, 10.0, 9, p1, w1); gimic_b200.gausspoints(0.0, 7.25614, 9, p2, w2)
# exact geometry as grid.f90:232-251 builds it
v1c, v2c, fix = xyz[1], xyz[0], xyz[3]
a = v1c - fix; b = v2c - fix
ortho = np.cross(a, b); ortho /= np.linalg.norm(ortho)
v3 = (b - a) / np.linalg.norm(b - a); v1 = -ortho; v2 = np.cross(v3, v1); v2 /= np.linalg.norm(v2)
oo = v1c + 1.48794 * v3
origin = oo - 6.0 * v2 - 5.0 * v1
grid = gimic_b200.Grid(origin, [v1, v2, v3], [p1, p2, np.zeros(1)], [w1, w2, np.ones(1)], radius=1e10)
B = np.array([0.0, 0.0, 1.0])
res = gimic_b200.integrate_distributed(g, grid, B, "total", 7)
if rank == 0:
    full = g.integrate(grid, B, "total", 7)
    ok = np.allclose(res, full, rtol=1e-12, atol=1e-14)
    blk = gold["blocks"][1]
    ok_gold = abs(res[0] - blk["au"]) < 1.01e-6 and abs(res[1] - blk["pos"]) < 1.01e-6 and abs(res[2] - blk["neg"]) < 1.01e-6
    print(json.dumps({"world": world, "sharded": res.tolist(), "single": full.tolist(), "match_single": bool(ok), "match_golden": bool(ok_gold)}))
    assert ok and ok_gold
    print("OK")
if world > 1:
    dist.destroy_process_group()
