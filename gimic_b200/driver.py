"""`python -m gimic_b200 gimic.inp`: launcher of the compiled run-mode driver (libgimic_b200_driver.so, include/gimic_b200_driver.h --
the code behind the `gimic-b200` program: gimic.inp -> grid + magnetic field -> cdens | integral | edens | divj, the reference's files
and stdout report; src/fgimic/gimic.F90:60-261, jfield.f90:250-443, integral.f90:167-183).

There is ONE driver, the compiled one.  This module adds what a Python process has and a C++ library has not: the torchrun /
torch.distributed glue.  Under torchrun (one process per GPU) every rank calls gimic_b200_run with its rank and two callbacks --
an all-gather of result rows and an all-reduce of the <= 7 integral sums, over NCCL (gloo in the CPU tests); the partition itself
(equal-cost tile shares, row slabs in integral mode) is decided by the compiled driver and the library.

    python -m gimic_b200 gimic.inp [--workdir DIR] [-y] [-t TITLE] [--vtk appended]
    python -m gimic_b200 gimic.1.inp gimic.2.inp ...        # a current-profile scan (one context, batched integrals)
    torchrun --nproc-per-node 8 -m gimic_b200 gimic.inp     # one rank per GPU
"""
import ctypes as C
import os
import sys
import tempfile
import numpy as np

from . import _lib

_GATHER = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_long, C.c_int, C.c_long, C.POINTER(C.c_long), C.POINTER(C.c_double), C.POINTER(C.c_double))
_REDUCE = C.CFUNCTYPE(C.c_int, C.c_void_p, C.POINTER(C.c_double), C.c_int)


class RunOpts(C.Structure):
    """gimic_b200_run_opts, include/gimic_b200_driver.h"""
    _fields_ = [("flags", C.c_int), ("device", C.c_int), ("ndevices", C.c_int), ("devices", C.POINTER(C.c_int)), ("workdir", C.c_char_p),
                ("title", C.c_char_p), ("report_path", C.c_char_p), ("rank", C.c_int), ("nranks", C.c_int), ("allgather_rows", _GATHER),
                ("allreduce_sum", _REDUCE), ("user", C.c_void_p)]


_D = None


def driver_lib():
    global _D
    if _D is None:
        _lib.lib()                                                   # libgimic_b200.so first (RTLD_GLOBAL): the driver links against it
        D = C.CDLL(os.path.join(os.path.dirname(os.path.abspath(__file__)), "libgimic_b200_driver.so"))
        D.gimic_b200_driver_last_error.restype = C.c_char_p
        D.gimic_b200_run.argtypes = [C.c_char_p, C.POINTER(RunOpts)]
        D.gimic_b200_run_scan.argtypes = [C.c_int, C.POINTER(C.c_char_p), C.c_int, C.c_int]
        _D = D
    return _D


def _dist():
    try:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            return dist, dist.get_rank(), dist.get_world_size()
    except Exception:
        pass
    return None, 0, 1


def collectives(dist):
    """the two callbacks of gimic_b200_run_opts over an initialised torch.distributed group: NCCL moves device tensors, any other
    backend (gloo in the CPU tests) host tensors.  Rows are gathered on rank 0 -- the only rank that writes anything."""
    import torch
    rank, world = dist.get_rank(), dist.get_world_size()
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")

    def allgather_rows(_user, n_total, ncols, count, index, rows, full):
        try:
            cnt = torch.tensor([count], dtype=torch.int64, device=dev)
            counts = [torch.zeros_like(cnt) for _ in range(world)]
            dist.all_gather(counts, cnt)
            counts = [int(c[0]) for c in counts]
            mx = max(max(counts), 1)
            buf = torch.zeros((mx, ncols + 1), dtype=torch.float64, device=dev)      # column 0: the row number (exact in a double below 2^53)
            if count:
                buf[:count, 0] = torch.from_numpy(np.ctypeslib.as_array(index, (count,)).astype(np.float64)).to(dev)
                buf[:count, 1:] = torch.from_numpy(np.ctypeslib.as_array(rows, (count, ncols))).to(dev)
            gathered = [torch.empty_like(buf) for _ in range(world)] if rank == 0 else None
            dist.gather(buf, gathered, dst=0)
            if rank == 0:
                out = np.ctypeslib.as_array(full, (n_total, ncols))
                for r, c in enumerate(counts):
                    if c:
                        blk = gathered[r][:c].cpu().numpy()
                        out[blk[:, 0].astype(np.int64)] = blk[:, 1:]
            return 0
        except Exception as e:                                         # an exception must not unwind through the C frames
            sys.stderr.write(f"gimic_b200: allgather_rows failed: {e}\n")
            return -1

    def allreduce_sum(_user, v, n):
        try:
            a = np.ctypeslib.as_array(v, (n,))
            t = torch.from_numpy(a.copy()).to(dev)
            dist.all_reduce(t)
            a[:] = t.cpu().numpy()
            return 0
        except Exception as e:
            sys.stderr.write(f"gimic_b200: allreduce_sum failed: {e}\n")
            return -1

    return _GATHER(allgather_rows), _REDUCE(allreduce_sum)


def run(infiles, workdir=None, dryrun=False, vtk_appended=False, title=None, device=-1, devices=None, report=None):
    """gimic_b200_run for one input, gimic_b200_run_scan for several.  With an initialised torch.distributed group of more than one
    rank the run is spread over the ranks (rank 0 writes the report and the files).  Returns 0; RuntimeError with the driver's
    message otherwise."""
    D = driver_lib()
    flags = (1 if dryrun else 0) | (2 if vtk_appended else 0)
    infiles = [infiles] if isinstance(infiles, (str, os.PathLike)) else list(infiles)
    if len(infiles) > 1:
        arr = (C.c_char_p * len(infiles))(*[os.fsencode(f) for f in infiles])
        rc = D.gimic_b200_run_scan(len(infiles), arr, int(device), flags)
    else:
        devs = (C.c_int * len(devices))(*devices) if devices else None
        o = RunOpts(flags=flags, device=int(device), ndevices=len(devices) if devices else 0, devices=devs,
                    workdir=os.fsencode(workdir) if workdir else None, title=title.encode() if title else None,
                    report_path=os.fsencode(report) if report else None)
        dist, rank, world = _dist()
        keep = None
        if world > 1:
            keep = collectives(dist)                                   # referenced until the call returns
            o.rank, o.nranks, o.allgather_rows, o.allreduce_sum = rank, world, keep[0], keep[1]
        sys.stdout.flush()
        rc = D.gimic_b200_run(os.fsencode(infiles[0]), C.byref(o))
        del keep
    if rc != 0:
        raise RuntimeError(f"gimic_b200 driver error {rc}: " + D.gimic_b200_driver_last_error().decode(errors="replace"))
    return 0


run_native = run          # the name of the round-1 switch


class Driver:
    """One gimic.inp through the compiled driver, with the report captured as text: Driver(inp, out=stream).run() writes the report to
    `out` (default sys.stdout) and the files into the work directory.  Under torch.distributed only rank 0 receives a report."""

    def __init__(self, inpfile, workdir=None, out=None, device=-1, vtk_appended=False, dryrun=False, title=None, devices=None):
        self.inpfile, self.workdir, self.out, self.device = str(inpfile), workdir, out if out is not None else sys.stdout, device
        self.vtk_appended, self.dryrun, self.title, self.devices = vtk_appended, dryrun, title, devices
        self.report = ""

    def run(self):
        with tempfile.NamedTemporaryFile(prefix="gimic_b200_report_", suffix=".txt", delete=False) as tf:
            path = tf.name
        try:
            run(self.inpfile, workdir=self.workdir, dryrun=self.dryrun, vtk_appended=self.vtk_appended, title=self.title, device=self.device,
                devices=self.devices, report=path)
            with open(path) as f:
                self.report = f.read()
        finally:
            os.unlink(path)
        self.out.write(self.report)
        return self


def run_scan(infiles, device=-1):
    """a current-profile scan: reports go to <input stem>.out, the table to current_profile.dat (gimic_b200_run_scan)"""
    return run(list(infiles), device=device)


class GridInfo(C.Structure):
    """gimic_b200_grid_info, include/gimic_b200_driver.h"""
    _fields_ = [("is_file", C.c_int), ("npts", C.c_int * 3), ("npoints", C.c_long), ("origin", C.c_double * 3), ("basv", C.c_double * 9),
                ("lengths", C.c_double * 3), ("magnet", C.c_double * 3), ("radius", C.c_double), ("has_center_bond", C.c_int),
                ("center_bond", C.c_double * 3)]


def input_grid(inpfile, workdir=None):
    """The grid and field direction a gimic.inp describes (gimic_b200_input_grid; host only): (grid, magnet, info) with grid a
    gimic_b200.Grid for std / base / bond grids or an (n, 3) array of points for Grid(file), magnet the unit field vector and info
    the remaining GridInfo fields as a dict."""
    from .gimic import Grid
    D = driver_lib()
    D.gimic_b200_input_grid.argtypes = [C.c_char_p, C.c_char_p, C.POINTER(GridInfo), C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_long]
    gi = GridInfo()
    wd = os.fsencode(workdir) if workdir else None

    def call(p, w, cap):
        if D.gimic_b200_input_grid(os.fsencode(str(inpfile)), wd, C.byref(gi), p, w, cap) != 0:
            raise RuntimeError("gimic_b200 driver error: " + D.gimic_b200_driver_last_error().decode(errors="replace"))
    call(None, None, 0)
    npts = tuple(gi.npts)
    need = 3 * gi.npoints if gi.is_file else sum(npts)
    pts, wgt = np.zeros(need), np.zeros(need)
    dp = C.POINTER(C.c_double)
    call(pts.ctypes.data_as(dp), wgt.ctypes.data_as(dp), need)
    info = dict(is_file=bool(gi.is_file), npts=npts, npoints=int(gi.npoints), lengths=np.array(gi.lengths), radius=float(gi.radius),
                center_bond=np.array(gi.center_bond) if gi.has_center_bond else None)
    magnet = np.array(gi.magnet)
    if gi.is_file:
        return pts.reshape(-1, 3), magnet, info
    cuts = np.cumsum((0,) + npts)
    axes = [pts[cuts[k]:cuts[k + 1]] for k in range(3)]
    weights = [wgt[cuts[k]:cuts[k + 1]] for k in range(3)]
    return Grid(np.array(gi.origin), np.array(gi.basv).reshape(3, 3), axes, weights, radius=float(gi.radius)), magnet, info


def mol_geometry(mol):
    """(symbols, coords) from the library's own MOL reader (gimic_b200_mol_geometry; host only, no device context)"""
    L = _lib.lib()
    n = L.gimic_b200_mol_geometry(os.fsencode(mol), 0, None, None)
    if n < 0:
        raise RuntimeError(L.gimic_b200_last_error().decode())
    xyz = np.zeros((n, 3)); sym = C.create_string_buffer(2 * n)
    L.gimic_b200_mol_geometry(os.fsencode(mol), n, xyz.ctypes.data_as(C.POINTER(C.c_double)), sym)
    raw = sym.raw[: 2 * n].decode()
    return [raw[2 * a: 2 * a + 2] for a in range(n)], xyz


def mol_summary(mol):
    """(natoms, primitive GTOs, contracted cartesian GTOs, is_turbomole) as new_basis prints them (gimic_b200_mol_summary)"""
    L = _lib.lib()
    info = (C.c_int * 5)()
    if L.gimic_b200_mol_summary(os.fsencode(mol), info) < 0:
        raise RuntimeError(L.gimic_b200_last_error().decode())
    return int(info[0]), int(info[1]), int(info[2]), bool(info[3])


def main(argv=None):
    import argparse
    ap = argparse.ArgumentParser(prog="gimic_b200", description="GIMIC grid hot path on B200 (cdens / integral / edens / divj)")
    ap.add_argument("infile", nargs="*", default=["gimic.inp"],
                    help="one gimic.inp, or several (a current-profile scan): they share one device context, integrals are "
                         "batched into one tensor pass, and each report is written to <input stem>.out")
    ap.add_argument("--workdir", default=None)
    ap.add_argument("--vtk", default="ascii", choices=["ascii", "appended"],
                    help="ascii: the reference's .vti files (e14.6); appended: same files with raw Float64 blocks (extra, not a reference format)")
    ap.add_argument("-y", "--dryrun", action="store_true",
                    help="lay out the grid and write mol.xyz / grid.xyz without calculating anything (src/gimic.in:53-54); needs no GPU")
    ap.add_argument("--devices", default=None, help="comma-separated CUDA ordinals, or 'all': several GPUs from this one process")
    # switches of the reference front end (src/gimic.in:36-57) that do not touch the hot path; accepted so that existing command
    # lines keep working
    ap.add_argument("-t", "--title", default=None, help="title of job (label only)")
    ap.add_argument("-d", "--debug", type=int, default=None, help="debug level (label only)")
    ap.add_argument("-o", "--output", default=None, help="base name for output file(s) (unused, like in the reference's fgimic backend)")
    ap.add_argument("-b", "--backend", default="fgimic", choices=["fgimic", "gimic"], help="only the fgimic path is provided")
    ap.add_argument("--native", action="store_true", help="accepted for round-1 command lines: the compiled driver is the only one")
    a = ap.parse_args(argv)
    device, devices = -1, None
    if a.devices:
        if a.devices == "all":
            n = _lib.lib().gimic_b200_device_count()
            devices = list(range(max(n, 0)))
        else:
            devices = [int(x) for x in a.devices.split(",")]
    if len(a.infile) == 1 and not a.dryrun and "LOCAL_RANK" in os.environ and int(os.environ.get("WORLD_SIZE", "1")) > 1:
        import torch
        import torch.distributed as dist
        device = int(os.environ["LOCAL_RANK"])
        if torch.cuda.is_available():
            torch.cuda.set_device(device)
            dist.init_process_group("nccl", device_id=torch.device("cuda", device))
        else:
            dist.init_process_group("gloo")
    try:
        return run(a.infile, workdir=a.workdir, dryrun=a.dryrun, vtk_appended=(a.vtk == "appended"), title=a.title, device=device, devices=devices)
    finally:
        dist, _, _ = _dist()
        if dist is not None:
            dist.destroy_process_group()
