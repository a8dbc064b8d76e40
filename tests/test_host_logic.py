"""CPU tests of the host-side logic: grid mirror, slab partitioning, and the world_size-2 (gloo) reduction
path of integral mode.  No CUDA calls: the per-rank partial sums are produced by the oracle here."""
import os
import sys
import numpy as np
import pytest

import fixtures
import oracle_lib as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_slab_partition_covers_range():
    from gimic_b200 import slab
    for n in (1, 7, 36, 256, 1000):
        for world in (1, 2, 3, 4, 8):
            parts = [slab(n, r, world) for r in range(world)]
            assert parts[0][0] == 0 and parts[-1][1] == n
            assert all(parts[i][1] == parts[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in parts]
            assert max(sizes) - min(sizes) <= 1


def test_grid_mirror_matches_oracle_gridpoints():
    from gimic_b200 import Grid
    og = O.grid_std([-1.0, 2.0, 0.5], [1.0, 0.2, 0], [0, 1.0, 0.1], [3.0, 2.0, 1.0], type="even", spacing=[0.5, 0.25, 0.5])
    pts, wgt = zip(*[og.axis(d) for d in range(3)])
    g = Grid(og.origin, og.basv, pts, wgt)
    assert g.npts == og.npts
    assert np.allclose(g.points(), og.points(), rtol=0, atol=1e-14)


def test_bench_slab_points_cover_grid_once():
    sys.path.insert(0, ROOT)
    import bench
    origin = np.array([-1.0, -2.0, -3.0]); pts = [np.arange(16) * 0.1, np.arange(16) * 0.2, np.arange(16) * 0.3]
    allp = np.vstack([bench.slab_points(origin, np.eye(3), pts, s) for s in range(bench.NSLAB)])
    assert allp.shape == (16 ** 3, 3)
    assert len({tuple(np.round(p, 9)) for p in allp}) == 16 ** 3


def _worker(rank, world, port, mol, xdens, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist
    import oracle_lib as OO
    from gimic_b200 import Grid, integrate_distributed
    dist.init_process_group("gloo", rank=rank, world_size=world)
    o = OO.Oracle.from_files(mol, xdens, screening_thrs=1e-8)
    xyz = o.atom_coords()
    og = OO.grid_bond(xyz[1], xyz[0], xyz[3], 1.48794, height=[-5.0, 5.0], width=[-1.25614, 6.0], type="gauss", gauss_order=9,
                      grid_points=[12, 12, 0])
    bb = og.magnet("z")
    pts, wgt = zip(*[og.axis(d) for d in range(3)])
    grid = Grid(og.origin, og.basv, pts, wgt, radius=og.radius)

    class FakeGimic:
        """stands in for the CUDA context: partial sums over rows [jlo, jhi) from the oracle"""
        def integrate(self, grid, B, spincase, what, jlo, jhi):
            r = grid.points().reshape(grid.npts[2], grid.npts[1], grid.npts[0], 3)
            out = np.zeros(7)
            for k in range(grid.npts[2]):
                for j in range(jlo, jhi):
                    tens = o.ctensor(r[k, j], spincase)
                    jv = OO.jvectors(tens, B)
                    nj = jv @ og.basv[2]
                    w = grid.wgt[0] * grid.wgt[1][j] * grid.wgt[2][k]
                    jp = nj * w
                    out[0] += jp.sum(); out[1] += jp[jp > 0].sum(); out[2] += jp[jp <= 0].sum()
            return out

    res = integrate_distributed(FakeGimic(), grid, bb, "total", 1)
    if rank == 0:
        q.put((res, o.integrate(og, bb, "total", 0)))
    dist.destroy_process_group()


def test_integral_allreduce_world2_gloo(cases):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, cases["c4h4"]["mol"], cases["c4h4"]["xdens"], q)) for r in range(2)]
    for p in procs:
        p.start()
    res, ref = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert np.allclose(res[:3], ref, rtol=1e-11, atol=1e-13)


def test_bench_reference_arm_json_contract():
    """`bench.py --impl reference` (the reference algorithm on the host CPU, oracle port): one JSON line with the keys the driver
    reads; tiny workload so that it runs in seconds.  The CUDA arm refuses to run without a GPU (no CPU fallback)."""
    import json, subprocess, sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.check_output([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "1",
                                   "--natoms", "6", "--grid", "16", "--cpu-points", "32"], text=True)
    lines = [l for l in out.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype",
              "data", "config", "impl", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["unit"] == "points/s" and d["dtype"] == "f64" and d["higher_is_better"] is True
    assert d["steps"] == 2 and d["value"] > 0 and "workload" in d["config"] and d["vs_baseline"] is None
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "sample" in cb
    assert d["e2e"] == {"value": d["value"], "unit": "points/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    import torch
    if not torch.cuda.is_available():
        p = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--steps", "1", "--natoms", "6", "--grid", "16"], capture_output=True, text=True)
        assert p.returncode != 0 and "no CPU path" in p.stdout


def test_gimic_is_a_backend_of_the_generic_connector():
    """src/pygimic/connector.pyx: the generic backend interface (jvector / jtensor / set_property); Gimic implements it"""
    import gimic_b200
    c = gimic_b200.GimicConnector()
    for call in (lambda: c.jvector((0, 0, 0)), lambda: c.jtensor((0, 0, 0))):
        try:
            call(); raise AssertionError("the generic connector must refuse")
        except gimic_b200.NotAvailable:
            pass
    assert c.set_property("magnet", [0, 0, 1]) is None
    assert issubclass(gimic_b200.Gimic, gimic_b200.GimicConnector)
    for m in ("jvector", "jtensor", "set_property"):
        assert getattr(gimic_b200.Gimic, m) is not getattr(gimic_b200.GimicConnector, m)
