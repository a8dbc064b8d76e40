"""Native driver: gimic.inp -> basis/densities on the GPU -> grid + magnetic field -> cdens | integral | edens | divj,
writing the reference's files and stdout report.  Replaces, for the hot path, `program gimic`
(src/fgimic/gimic.F90:60-261: initialize, driver, run_cdens, run_integral), jvector_plots
(src/fgimic/jfield.f90:250-443) and the report printing of src/fgimic/integral.f90:167-183,306-322,502-510.

    python -m gimic_b200 gimic.inp [--workdir DIR]

Under torch.distributed (torchrun, one process per GPU) cdens splits the flat point index into contiguous slabs and
gathers the tensors on rank 0 (which writes the files, like the reference's MPI path, jfield.f90:90-137); integral mode
splits plane rows and all-reduces the partial sums.
"""
import os
import re
import sys
import time
import numpy as np

from . import inp as _inp
from . import grids, writers
from .gimic import Gimic, integrate_distributed, slab

SPIN_LABEL = {"total": "total", "alpha": "alpha", "beta": "beta", "spindens": "spin"}


def au2si(au):
    """au2si, globals.f90:309-332 (nA/T per atomic unit of dJ/dB)"""
    aulength, auspeedoflight, speedoflight = 0.52917726e-10, 137.03599e0, 299792458.0
    aucharge, hbar = 1.60217733e-19, 1.05457267e-34
    autime = aulength * auspeedoflight / speedoflight
    autesla = hbar / aucharge / aulength / aulength
    return au * (aucharge / autime / autesla) * 1.0e9


def read_mol_geometry(mol):
    """atom symbols and coordinates (bohr) of an INTGRL/MOL file (intgrl.f90:91-115); host-only helper"""
    with open(mol) as f:
        lines = f.read().split("\n")
    natoms = int(lines[3].split()[0])
    syms, xyz, i = [], [], 5
    for _ in range(natoms):
        hdr = lines[i].split()
        nsh = int(hdr[2]); nblk = [int(x) for x in hdr[3:3 + nsh]]
        syms.append(lines[i + 1][:2]); xyz.append([float(v.replace("D", "E").replace("d", "e")) for v in lines[i + 1][4:].split()[:3]])
        i += 2
        for nb in nblk:
            for _b in range(nb):
                npf, ncf = (int(x) for x in lines[i].split()[:2])
                i += 1
                for _p in range(npf):          # a primitive's 1+ncf values may wrap over several lines
                    got = 0
                    while got < 1 + ncf:
                        got += len(lines[i].split()); i += 1
    return syms, np.array(xyz)


def mol_geometry(mol):
    """(symbols, coords) from the library's own MOL reader (gimic_b200_mol_geometry; host only, no device context)"""
    import ctypes as C
    from . import _lib
    L = _lib.lib()
    n = L.gimic_b200_mol_geometry(os.fsencode(mol), 0, None, None)
    if n < 0:
        raise RuntimeError(L.gimic_b200_last_error().decode())
    xyz = np.zeros((n, 3)); sym = C.create_string_buffer(2 * n)
    L.gimic_b200_mol_geometry(os.fsencode(mol), n, xyz.ctypes.data_as(C.POINTER(C.c_double)), sym)
    raw = sym.raw[: 2 * n].decode()
    return [raw[2 * a: 2 * a + 2] for a in range(n)], xyz


def mol_summary(mol):
    """(natoms, primitive GTOs, contracted cartesian GTOs, is_turbomole) as new_basis prints them (gimic_b200_mol_summary; its fifth
    entry, the spherical count, is only needed to size an XDENS over spherical components)"""
    import ctypes as C
    from . import _lib
    L = _lib.lib()
    info = (C.c_int * 5)()
    if L.gimic_b200_mol_summary(os.fsencode(mol), info) < 0:
        raise RuntimeError(L.gimic_b200_last_error().decode())
    return int(info[0]), int(info[1]), int(info[2]), bool(info[3])


def _dist():
    try:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            return dist, dist.get_rank(), dist.get_world_size()
    except Exception:
        pass
    return None, 0, 1



class _GfortranNaN:
    """report sink: Python formats a NaN as 'nan'; gfortran's F and E edits (and the native driver) write 'NaN'"""
    _pat = re.compile(r"(?<![A-Za-z])nan(?![A-Za-z])")

    def __init__(self, out):
        self._out = out

    def write(self, s):
        return self._out.write(self._pat.sub("NaN", s) if "nan" in s else s)

    def __getattr__(self, name):
        return getattr(self._out, name)


class Driver:
    def __init__(self, inpfile, workdir=None, out=None, device=-1, gimic=None, vtk_appended=False, dryrun=False, title=None):
        self._t0, self._cpu0 = time.perf_counter(), os.times()    # stockas_klocka reports the times of the whole run
        self.workdir = workdir or os.path.dirname(os.path.abspath(inpfile))
        self.inp = _inp.parse_file(inpfile)
        if dryrun:                               # the -y switch overrides the keyword (src/gimic.in:139-140)
            self.inp.values[""]["dryrun"] = True
        if title:                                # -t (src/gimic.in:135-136)
            self.inp.values[""]["title"] = str(title)
        self.vtk_appended = bool(vtk_appended)   # extra: .vti files with raw appended Float64 data instead of ASCII e14.6
        self.dist, self.rank, self.world = _dist()
        self.out = _GfortranNaN(out if out is not None else sys.stdout)
        I = self.inp
        self.uhf = bool(I.get("openshell"))
        path = lambda n: n if os.path.isabs(n) else os.path.join(self.workdir, n)
        # what decides the contents of the device context; inputs that agree on it can share one (run_scan)
        self.context_key = (os.path.realpath(path(I.get("basis"))), os.path.realpath(path(I.get("xdens"))), self.uhf,
                            bool(I.get("Advanced.GIAO")), bool(I.get("Advanced.diamag")), bool(I.get("Advanced.paramag")),
                            bool(I.get("Advanced.screening")), float(I.get("Advanced.screening_thrs")),
                            bool(I.get("Advanced.spherical")))
        if I.get("dryrun") and gimic is None:
            # driver (gimic.F90:142-159): a dry run builds the basis but neither the c2s operator nor the densities;
            # here that means no device context at all, only the MOL geometry (parsed by the library's own reader)
            self.g = None
            self.symbols, self.xyz = mol_geometry(path(I.get("basis")))
        else:
            self.g = gimic if gimic is not None else Gimic(
                path(I.get("basis")), path(I.get("xdens")), uhf=self.uhf, giao=I.get("Advanced.GIAO"),
                diamag=I.get("Advanced.diamag"), paramag=I.get("Advanced.paramag"),
                screening=I.get("Advanced.screening"), screening_thrs=I.get("Advanced.screening_thrs"), device=device,
                spherical=bool(I.get("Advanced.spherical")))
            self.xyz = self.g.atom_coords()
            self.symbols = self._symbols(path(I.get("basis")))
        self.summary = mol_summary(path(I.get("basis")))
        self.grid = grids.from_input(I, self.xyz, self.workdir)
        self.magnet_log = []        # what get_magnet prints every time it is called (magnet.f90:66-86)
        self.magnet = grids.get_magnet(self.grid, I.get("magnet_axis"), I.get("magnet"), self.magnet_log)

    @staticmethod
    def _symbols(mol):
        return read_mol_geometry(mol)[0]

    def say(self, s=""):
        if self.rank == 0:
            self.out.write(" " + s + "\n" if s else "\n")

    # -------------------------------------------------------------------------------------------------
    def run(self, integral_results=None):
        self._run(integral_results)
        # finalize() + stockas_klocka (gimic.F90:43-52,134-138; grid.f90:453; basis.f90:348-349; timer.f90:13-43).  The jobscripts
        # recognise a finished slice by the word "wall" in gimic.N.out (jobscripts/src/current-profile-local-submit:52).
        cpu = os.times()
        self.say("*** Deallocated grid data")
        self.say("INFO: Deallocated basis set and atom data")
        self.say()
        self.say("-" * 70)
        for label, t in (("   wall time:", time.perf_counter() - self._t0), ("        user:", cpu.user - self._cpu0.user),
                         ("         sys:", cpu.system - self._cpu0.system)):
            self.say(f"{label}{t:9.2f}sec ({t / 3600.0:6.1f} h )")
        self.say("-" * 70)
        self.say(time.strftime("%a %b %e %H:%M:%S %Y"))
        self.say("Hello World! (tm)")
        self.say()
        self.say("done.")
        self.say()

    def _run(self, integral_results=None):
        I = self.inp
        # initialize(), gimic.F90:107-131
        self.say()
        self.say(time.strftime("%a %b %e %H:%M:%S %Y"))
        self.say((" TITLE: " + str(I.get("title")).strip()).rstrip())       # msg_out trims trailing blanks
        self.say()
        if not I.get("Advanced.GIAO"):
            self.say("INFO: GIAOs not used!"); self.say()
        if not I.get("Advanced.diamag"):
            self.say("INFO: Diamagnetic contributions not calculated!"); self.say()
        if not I.get("Advanced.paramag"):
            self.say("INFO: Paramagnetic contributions not calculated!"); self.say()
        if not I.get("Advanced.diamag") and not I.get("Advanced.paramag"):
            self.say("    ...this does not make sense..."); self.say()
            raise ValueError("neither diamagnetic nor paramagnetic contributions requested: nothing to calculate (gimic.F90:124-130)")
        # driver(), gimic.F90:141-165: what new_basis (intgrl.f90:48-60, basis.f90:44-80), read_dens (dens.f90:94-103), new_grid and
        # plot_grid_xyz (grid.f90:608-609) print on the way
        natoms, ngto, ncgto, turbomole = self.summary
        if turbomole:
            self.say("INFO: Detected TURBOMOLE input"); self.say()
        self.say(f"Number of atoms ={natoms:4d}"); self.say()
        self.say("Normalizing basis"); self.say()
        self.say(f"  Total number of primitive  GTO's {ngto:6d}")
        self.say(f"  Total number of contracted GTO's {ncgto:6d}"); self.say()
        if I.get("Advanced.screening") and float(I.get("Advanced.screening_thrs")) > 0.0:
            self.say("*** Calculating screening coefficients")
            self.say("INFO: Screening threshold: " + writers.fortran_e(float(I.get("Advanced.screening_thrs")), 12, 4)); self.say()
        else:
            self.say("INFO: Screening is not used")
        if not I.get("dryrun"):
            if self.uhf:
                self.say("INFO: scaling perturbed densities by 0.d5")
            if turbomole:
                self.say("INFO: Reordering densities [TURBOMOLE]")
        if self.rank == 0:
            for line in self.grid.log:
                self.out.write(line + "\n")
            writers.write_mol_xyz(os.path.join(self.workdir, "mol.xyz"), self.symbols, self.xyz)
            writers.write_grid_xyz(os.path.join(self.workdir, "grid.xyz"), self.grid, self.symbols, self.xyz)
        self.say("*** Grid plot in grid.xyz")
        self._field_lines()
        self.say("INFO: " + ("Open-shell calculation" if self.uhf else "Closed-shell calculation"))
        self.say()
        calc = I.get("calc")
        if I.get("dryrun"):
            # gimic.F90:174-185,196-204,222-230: the note, then the run mode's banner, then return before any arithmetic
            self.say("*** Dry run, not calculating ...")
            self.say()
            if calc == "cdens":
                self.say("Calculating current density")
                self.say("*****************************************")
            elif calc == "integral":
                self.say("Integrating current density")
                self.say("*****************************************")
            return
        if calc == "cdens":
            self.run_cdens()
        elif calc == "integral":
            self.run_integral(integral_results)
        elif calc in ("edens", "divj"):
            self.run_scalar(calc)

    def _gather_rows(self, part, n):
        """rows of this rank's slab -> the full (n, width) array on rank 0 (None elsewhere); identity for a single process"""
        if self.world == 1:
            return part
        import torch
        # NCCL moves device tensors; any other backend (gloo in the CPU tests) moves host tensors
        dev = torch.device("cuda", torch.cuda.current_device()) if self.dist.get_backend() == "nccl" else torch.device("cpu")
        part = np.ascontiguousarray(part, dtype=np.float64).reshape(part.shape[0], -1)
        sizes = [slab(n, r, self.world) for r in range(self.world)]
        mx = max(b - a for a, b in sizes)
        buf = torch.zeros((mx, part.shape[1]), dtype=torch.float64, device=dev)
        buf[: part.shape[0]] = torch.from_numpy(part).to(dev)
        gathered = [torch.empty_like(buf) for _ in range(self.world)] if self.rank == 0 else None
        self.dist.gather(buf, gathered, dst=0)
        if self.rank != 0:
            return None
        return np.concatenate([gathered[r][: b - a].cpu().numpy() for r, (a, b) in enumerate(sizes)])

    def _field_lines(self):
        if self.rank == 0:
            for line in self.magnet_log:
                self.out.write(line + "\n")

    def _gather_indexed(self, index, part, n):
        """rows `part` of the points `index` (this rank's share of a cost-balanced partition) -> the full (n, width) array on rank 0"""
        import torch
        dev = torch.device("cuda", torch.cuda.current_device()) if self.dist.get_backend() == "nccl" else torch.device("cpu")
        part = np.ascontiguousarray(part, dtype=np.float64).reshape(index.shape[0], -1)
        cnt = torch.tensor([index.shape[0]], dtype=torch.int64, device=dev)
        counts = [torch.zeros_like(cnt) for _ in range(self.world)]
        self.dist.all_gather(counts, cnt)
        counts = [int(c[0]) for c in counts]
        mx = max(max(counts), 1)
        buf = torch.zeros((mx, part.shape[1] + 1), dtype=torch.float64, device=dev)     # column 0: the point index (exact in a double below 2^53)
        buf[: index.shape[0], 0] = torch.from_numpy(index.astype(np.float64)).to(dev)
        buf[: index.shape[0], 1:] = torch.from_numpy(part).to(dev)
        gathered = [torch.empty_like(buf) for _ in range(self.world)] if self.rank == 0 else None
        self.dist.gather(buf, gathered, dst=0)
        if self.rank != 0:
            return None
        full = np.zeros((n, part.shape[1]))
        for r, c in enumerate(counts):
            rows = gathered[r][:c].cpu().numpy()
            full[rows[:, 0].astype(np.int64)] = rows[:, 1:]
        return full

    def _partition(self):
        """this rank's equal-COST share of the grid (gimic_b200_partition_*): replaces the equal-count slabs of schedule(), parallel.F90:66-84"""
        grid = self.grid
        self.g.partition(grid.points() if grid.mode == "file" else grid, self.rank, self.world)

    def _tensors(self, spincase):
        """calc_jtensors (jfield.f90:62-138): one call on a single device; a cost-balanced share per rank, gathered on rank 0"""
        grid, n = self.grid, self.grid.n
        if self.world == 1:
            return self.g.jtensors(grid.points(), spincase) if grid.mode == "file" else self.g.jtensors_grid(grid, 0, n, spincase)
        self._partition()
        res = self.g.partition_calc(None, spincase, tens=True)
        return self._gather_indexed(res["index"], res["tens"], n)

    def _jvectors(self, spincase, want_jmod):
        """J = T.B (and the signed modulus) straight from the contraction on this rank's share of the points, gathered on rank 0:
        3 (+1) doubles per point cross the wire instead of 9, and the contraction runs with 2 operand planes instead of 4"""
        n = self.grid.n
        if self.world == 1:
            f = self.g.fields(self.grid.points(), self.magnet, spincase, jvec=True, jmod=want_jmod)
            return f["jvec"], (f["jmod"] if want_jmod else None)
        self._partition()
        res = self.g.partition_calc(self.magnet, spincase, jvec=True, jmod=want_jmod)
        rows = np.concatenate([res["jvec"], res["jmod"].reshape(-1, 1)], axis=1) if want_jmod else res["jvec"]
        full = self._gather_indexed(res["index"], rows, n)
        if full is None:
            return None, None
        return np.ascontiguousarray(full[:, :3]), (np.ascontiguousarray(full[:, 3]) if want_jmod else None)

    def run_cdens(self):
        """run_cdens (gimic.F90:196-220) + jvector_plots (jfield.f90:250-443)"""
        self.say("Calculating current density")
        self.say("*****************************************")
        cases = [("total", "")] + ([("alpha", "alpha"), ("beta", "beta"), ("spindens", "spindens")] if self.uhf else [])
        grid, wd, I = self.grid, self.workdir, self.inp
        want_jmod = bool(I.get("Essential.jmod")) and grid.is_3d()
        want_acid = bool(I.get("Essential.acid")) and grid.is_3d()
        # Only J (and |J|) is written when neither ACID nor the property quadrature is asked for: the library then contracts with B
        # inside the GEMM (2 operand planes instead of 4) and never forms the tensors.
        j_only = not want_acid and not I.get("Essential.prop")
        cache, jcache = {}, {}
        if self.uhf:
            # everything is linear in the densities: alpha and beta are evaluated once, total = alpha + beta and
            # spindens = alpha - beta exactly as ctensor combines them (jtensor.F90:86-99); the reference re-evaluates
            # everything for each of the four spin cases (6 tensor passes per point, gimic.F90:206-217)
            store = jcache if j_only else cache
            for sc in ("alpha", "beta"):
                store[sc] = self._jvectors(sc, False)[0] if j_only else self._tensors(sc)
            if self.rank == 0:
                store["total"] = store["alpha"] + store["beta"]
                store["spindens"] = store["alpha"] - store["beta"]
        for sc, tag in cases:
            if j_only:
                tens = None
                if self.uhf:
                    if self.rank != 0:
                        continue
                    r = grid.points()
                    f = {"jvec": jcache[sc]}
                    if want_jmod:
                        f["jmod"] = self.g.jmod_from_jvec(r, jcache[sc], self.magnet)
                else:
                    jv, jm = self._jvectors(sc, want_jmod)
                    if self.rank != 0:
                        continue
                    r = grid.points()
                    f = {"jvec": jv, "jmod": jm}
            else:
                tens = cache[sc] if self.uhf and (self.rank == 0) else (None if self.uhf else self._tensors(sc))
                if self.rank != 0:
                    continue
                r = grid.points()
                f = self.g.fields_from_tensors(r, tens, self.magnet, jvec=True, jmod=want_jmod, acid=want_acid)
            self.out.write(" magnetic field\n" + "".join(writers._ld_real(b) for b in self.magnet) + "\n \n")   # print *, magnet
            jv = f["jvec"]
            regular = grid.mode in ("std", "base", "bond")
            if grid.gauss and grid.mode != "file":
                writers.write_jmod_txt(os.path.join(wd, f"jmod{tag}.txt"), grid, jv, regular=regular and
                                       (grid.mode == "bond" or grid.gtype == "even"))
            if grid.is_3d():
                if I.get("Essential.acid"):
                    writers.write_vti_scalar(os.path.join(wd, "acid.vti"), grid, f["acid"], self.vtk_appended)
                if I.get("Essential.jmod"):
                    writers.write_vti_scalar(os.path.join(wd, f"jmod{tag}.vti"), grid, f["jmod"], self.vtk_appended)
            if I.get("Essential.prop"):
                self.run_property(tens)
            if grid.mode in ("std", "base", "bond") and grid.gtype == "even":
                writers.write_vti_vector(os.path.join(wd, f"jvec{tag}.vti"), grid, writers.radius_masked_vectors(grid, jv), self.vtk_appended)
            elif (grid.mode in ("std", "base") and grid.gauss) or grid.mode == "file":
                ele = os.path.join(wd, "grid.1.ele")
                if os.path.exists(ele):
                    writers.write_vtu_vector(os.path.join(wd, "jvec.vtu"), r, jv, writers.read_ele(ele))
                else:
                    self.out.write(" not writing a vtu file, because the file grid.1.ele was not found.\n")

    def run_property(self, tens):
        """get_property (jfield.f90:584-929): needs coord.au, gridfile.grd, grid_w.grd (and nelpts.info) in the work dir;
        the tensor field must have been computed on the points of gridfile.grd (Grid(file))."""
        wd, w = self.workdir, self.out.write
        need = [os.path.join(wd, f) for f in ("coord.au", "gridfile.grd", "grid_w.grd")]
        if not all(os.path.exists(f) for f in need):
            w(" at least one of the files coord.au, gridfile.grd, and grid_w.grd is missing.Therefore any property calculation is skipped.\n")
            return
        coord = np.loadtxt(need[0]).reshape(-1, 3)
        grd = np.loadtxt(need[1]).reshape(-1, 3)
        wg = np.loadtxt(need[2]).ravel()
        nel = os.path.join(wd, "nelpts.info")
        counts = np.loadtxt(nel, dtype=np.int64).reshape(-1, 2)[:, 1] if os.path.exists(nel) else np.array([grd.shape[0]])
        if counts.sum() != grd.shape[0]:
            counts = np.array([grd.shape[0]])
        res = self.g.property(grd, wg, tens, coord, counts)
        self.property_results = res
        # integrand plots (only when the TetGen cell file is there, jfield.f90:677-686, 786-808, 911-919)
        ele = os.path.join(wd, "grid.1.ele")
        cells = writers.read_ele(ele) if os.path.exists(ele) else None
        def plot_integrands(centre, names):
            f4 = self.g.property_integrand(grd, tens, centre)
            for col, name in zip((3, 0, 1, 2), names):
                writers.write_vtu_scalar(os.path.join(wd, name), grd, f4[:, col], cells)
        w(f" npts{grd.shape[0]:12d}\n")
        def table(contrib, lead="  "):       # write(*,*) " " before the shielding tables, write(*,*) "" before the chi table (jfield.f90:762,890)
            w(lead + "\n atom contributions, total, positive, negative\n")
            for l, c in enumerate(contrib):
                w(f"atom {l + 1:5d}{c[0]:14.6f}{c[1]:14.6f}{c[2]:14.6f}\n")
            cs = contrib.sum(0)
            w(f"{'sum ':>10s}{cs[0]:14.6f}{cs[1]:14.6f}{cs[2]:14.6f}\n")
            w(" ****************************************************\n")
        for k in range(coord.shape[0]):
            sg = res["sigma"][k]
            w(f" atom {k + 1:12d}\n in ppm\n")
            for lbl, v in zip(("sigma_xx ", "sigma_yy ", "sigma_zz "), sg):
                w(f" {lbl:>10s}  {v:14.6f}\n")
            w(f"{'shielding constant    = ':>30s}  {res['sigma_iso'][k]:14.6f}\n")
            w(f"{'positive contribution = ':>30s}  {res['sigma_pos'][k]:14.6f}\n")
            w(f"{'negative contribution = ':>30s}  {res['sigma_neg'][k]:14.6f}\n")
            w(f"{'sum = ':>30s}  {res['sigma_pos'][k] + res['sigma_neg'][k]:14.6f}\n")
            table(res["sigma_atoms"][k])
            if cells is not None:
                names = [f"sigma{k + 1}.vtu"] + [f"sigma_{c}{k + 1}.vtu" for c in ("xx", "yy", "zz")]
                for nm in names[1:]:
                    w(f" {nm:<70s}\n")                      # print *, filename  (character(len=70), jfield.f90:606,801)
                plot_integrands(coord[k], names)
        w(" \n \n")
        for lbl, v in zip(("chi_xx ", "chi_yy ", "chi_zz "), res["chi"]):
            w(f" {lbl:>7s}  {v:14.8f}\n")
        w(" in au\n")
        # (X,A30,2X,F14.6) with 32-character labels, jfield.f90:876-878: the A30 edit descriptor keeps the leftmost 30 characters, so the
        # reference prints these three lines without their '= ' (test/benzene/magnetizability/reference/stdout)
        w(f" {'isotropic magnetizability chi = '[:30]:>30s}  {res['chi_iso']:14.6f}\n")
        w(f" {'positive contribution         = '[:30]:>30s}  {res['chi_pos']:14.6f}\n")
        w(f" {'negative contribution         = '[:30]:>30s}  {res['chi_neg']:14.6f}\n")
        w(f" {'sum ':>30s}  {res['chi_pos'] + res['chi_neg']:14.6f}\n \n")
        fac = 7.89104e-29                                   # fac_au2simag, jfield.f90:606
        w(" in SI units J/T^2 \n conversion factor: 7.89104*10^-29 J/T^2 \n \n")
        for lbl, v in (("isotropic magnetizability = ", res["chi_iso"]), ("positive contribution     = ", res["chi_pos"]),
                       ("negative contribution     = ", res["chi_neg"]), ("sum ", res["chi_pos"] + res["chi_neg"])):
            w(f"{lbl:>30s}  {writers.fortran_e(v * fac, 14, 6)}\n")
        w(" ****************************************************\n")
        table(res["chi_atoms"], " ")
        if cells is not None:
            plot_integrands(None, ["intchi.vtu", "intchi_xx.vtu", "intchi_yy.vtu", "intchi_zz.vtu"])

    def _note_spin(self, sc):
        if self.uhf:
            self.say(f"*** Integrating {SPIN_LABEL[sc]} density")

    def integral_cases(self):
        I = self.inp
        cases = ["total"] + (["alpha", "beta", "spindens"] if self.uhf else [])
        what = 1 | (2 if I.get("Essential.jmod") else 0) | (4 if I.get("Essential.acid") else 0)
        return cases, what

    def run_integral(self, res=None):
        """run_integral (gimic.F90:222-261) with the report formats of integral.f90:167-183,306-322,502-510.
        res: precomputed {spincase: 7 sums} (run_scan evaluates many inputs in one tensor pass)"""
        I = self.inp
        self.say("Integrating current density")
        self.say("*****************************************")
        cases, what = self.integral_cases()
        if res is None:
            res = {sc: integrate_distributed(self.g, self.grid, self.magnet, sc, what if sc == "total" else (what & 3)) for sc in cases}
        self.results = res
        bar = "*" * 60
        bound = self.grid.radius
        field_line = self._field_lines            # integrate_* call get_magnet again (integral.f90:85,225)
        def block(lbl_au, lbl_si, x, p, n):
            self.say()
            self.say(bar)
            self.say(f"{lbl_au}{x:13.6f}")
            self.say(f"      Positive contribution:{p:13.6f}  ({au2si(p):11.6f} )")
            self.say(f"      Negative contribution:{n:13.6f}  ({au2si(n):11.6f} )")
            self.say()
            self.say(f"{lbl_si}{au2si(x):13.6f}")
            self.say(f"      (conversion factor)  :{au2si(1.0):13.6f}")
            self.say(bar)
            self.say()
        if I.get("Essential.jmod"):
            self.say("*** Integrating |J|")
            for sc in cases:
                self._note_spin(sc)
                field_line()
                if bound < 1.0e10:
                    self.say(" Integration bound set to radius " + writers._ld_real(bound).rstrip())   # write(str_g, *) ..., bound (integral.f90:93,231)
                block("Induced mod current (au)   :", "Induced mod current (nA/T) :", *res[sc][3:6])
            self.say()
        else:
            self.out.write(" Jmod integration skipped.\n")          # write(*,*) "...", gimic.F90:242
        self.say("*** Integrating current")
        for sc in cases:
            self._note_spin(sc)
            field_line()
            if bound < 1.0e10:
                self.say(" Integration bound set to radius " + writers._ld_real(bound).rstrip())   # write(str_g, *) ..., bound (integral.f90:93,231)
            block("   Induced current (au)    :", "   Induced current (nA/T)  :", *res[sc][0:3])
        self.say()
        if I.get("Essential.acid"):
            self.say("*** Integrating ACID density")
            acid = float(np.sqrt(res["total"][6]))
            self.say()
            self.say(bar)
            self.say(f"   ACID (au) sqrt(delta J^2):{acid:13.6f}")
            self.say(f"   ACID (nA/T)              :{au2si(acid):13.6f}")
            self.say()
            self.say(bar)
            self.say()
            self.say()                          # call nl after integrate_acid, gimic.F90:257

    def run_scalar(self, calc):
        """edens / divj: whitelisted by the reference front-end (src/gimic.in:267) but not implemented at this commit.
        Defined here as rho = Phi^T D Phi and div(T.B) (central differences); written as <calc>.vti on 3-D even grids and
        <calc>.txt ('x y z value', bohr) otherwise.  No reference output exists: parity unpinned."""
        grid = self.grid
        r = grid.points()
        f = self.g.fields(r, self.magnet, "total", edens=(calc == "edens"), divj=(calc == "divj"))
        if self.rank != 0:
            return
        if grid.mode != "file" and grid.gtype == "even" and grid.npts[0] > 1 and grid.npts[1] > 1:
            writers.write_vti_scalar(os.path.join(self.workdir, f"{calc}.vti"), grid, f[calc], self.vtk_appended)
        else:
            np.savetxt(os.path.join(self.workdir, f"{calc}.txt"), np.column_stack([r, f[calc]]), fmt="%20.12e")


def run_scan(infiles, device=-1, outs=None):
    """A current-profile scan (jobscripts/src/current-profile-local-submit: `gimic gimic.N.inp > gimic.N.out` for hundreds of
    thin slices, one process and one MOL/XDENS read each) as ONE context and ONE tensor pass per spin case: all inputs that
    share basis, densities and Advanced settings are integrated by gimic_b200_integrate_batch.  Reports go to
    <input stem>.out next to each input (or to the streams in `outs`).  Inputs with calc != integral run one by one on the
    shared context.  Returns the drivers (results in .results)."""
    import io
    drivers = []
    for k, f in enumerate(infiles):
        # a report file is only open while it is written (a scan can have more slices than the process may hold open files)
        out = outs[k] if outs is not None else io.StringIO()
        share = next((d.g for d in drivers if d.g is not None and d.context_key == _context_key_of(f)), None)
        drivers.append(Driver(f, out=out, device=device, gimic=share))
    batch = [d for d in drivers if d.inp.get("calc") == "integral" and not d.inp.get("dryrun") and d.world == 1]
    pre = {id(d): {} for d in batch}
    by_ctx = {}
    for d in batch:
        by_ctx.setdefault(id(d.g), []).append(d)
    for ds in by_ctx.values():
        cases = ds[0].integral_cases()[0]
        for sc in cases:
            what = 0
            for d in ds:
                what |= d.integral_cases()[1] if sc == "total" else (d.integral_cases()[1] & 3)
            sums = ds[0].g.integrate_batch([d.grid for d in ds], np.array([d.magnet for d in ds]), sc, what)
            for d, row in zip(ds, sums):
                pre[id(d)][sc] = row
    for d, f in zip(drivers, infiles):
        if outs is None:
            with open(os.path.splitext(f)[0] + ".out", "w") as fh:
                d.out = fh
                d.run(pre.get(id(d)))
            d.out = None
        else:
            d.run(pre.get(id(d)))
    write_current_profile(drivers)
    return drivers


def _profile_delta(workdir):
    """slice width from the jobscripts' calculation.dat ('delta=0.02 nsteps=400', jobscripts/src/current-profile-header:38), or None"""
    import re
    try:
        m = re.search(r"delta=([-+.\deEdD]+)", open(os.path.join(workdir, "calculation.dat")).read())
        return float(m.group(1).replace("d", "e").replace("D", "e")) if m else None
    except (OSError, ValueError):
        return None


def write_current_profile(drivers):
    """current_profile.dat next to the first input: slice position (index x delta when calculation.dat is there, else the index),
    net / diatropic / paratropic current strength in nA/T -- what jobscripts/src/gradient.sh.in:38-47 assembles by grepping the
    'Induced current' blocks of every gimic.N.out, here from the unrounded sums of the batched pass (8 decimals)."""
    rows = [d for d in drivers if d.rank == 0 and d.inp.get("calc") == "integral" and not d.inp.get("dryrun") and getattr(d, "results", None)]
    if len(rows) < 2:
        return None
    wd = rows[0].workdir
    delta = _profile_delta(wd)
    path = os.path.join(wd, "current_profile.dat")
    with open(path, "w") as f:
        for k, d in enumerate(rows):
            tot, pos, neg = (au2si(v) for v in d.results["total"][0:3])
            x = f"{k * delta:5.2f}" if delta is not None else f"{k:5d}"
            f.write(f"{x}\t{tot: .8f}\t{pos: .8f}\t{neg: .8f}\n")
    return path


def _context_key_of(inpfile):
    I = _inp.parse_file(inpfile)
    wd = os.path.dirname(os.path.abspath(inpfile))
    path = lambda n: n if os.path.isabs(n) else os.path.join(wd, n)
    return (os.path.realpath(path(I.get("basis"))), os.path.realpath(path(I.get("xdens"))), bool(I.get("openshell")),
            bool(I.get("Advanced.GIAO")), bool(I.get("Advanced.diamag")), bool(I.get("Advanced.paramag")),
            bool(I.get("Advanced.screening")), float(I.get("Advanced.screening_thrs")), bool(I.get("Advanced.spherical")))


def run_native(infiles, workdir=None, dryrun=False, vtk_appended=False, title=None, device=-1, devices=None, report=None):
    """The same run through the compiled driver (include/gimic_b200_driver.h): gimic_b200_run for one input, gimic_b200_run_scan for several.
    Returns 0; raises RuntimeError with the driver's message otherwise."""
    import ctypes as C
    from . import _lib
    _lib.lib()                                                   # libgimic_b200.so first (RTLD_GLOBAL): the driver links against it
    D = C.CDLL(os.path.join(os.path.dirname(_lib.SO_PATH), "libgimic_b200_driver.so"))
    D.gimic_b200_driver_last_error.restype = C.c_char_p

    class RunOpts(C.Structure):
        _fields_ = [("flags", C.c_int), ("device", C.c_int), ("ndevices", C.c_int), ("devices", C.POINTER(C.c_int)), ("workdir", C.c_char_p),
                    ("title", C.c_char_p), ("report_path", C.c_char_p)]
    flags = (1 if dryrun else 0) | (2 if vtk_appended else 0)
    infiles = [infiles] if isinstance(infiles, str) else list(infiles)
    if len(infiles) > 1:
        arr = (C.c_char_p * len(infiles))(*[os.fsencode(f) for f in infiles])
        D.gimic_b200_run_scan.argtypes = [C.c_int, C.POINTER(C.c_char_p), C.c_int, C.c_int]
        rc = D.gimic_b200_run_scan(len(infiles), arr, int(device), flags)
    else:
        devs = (C.c_int * len(devices))(*devices) if devices else None
        o = RunOpts(flags=flags, device=int(device), ndevices=len(devices) if devices else 0, devices=devs,
                    workdir=os.fsencode(workdir) if workdir else None, title=title.encode() if title else None,
                    report_path=os.fsencode(report) if report else None)
        D.gimic_b200_run.argtypes = [C.c_char_p, C.POINTER(RunOpts)]
        sys.stdout.flush()
        rc = D.gimic_b200_run(os.fsencode(infiles[0]), C.byref(o))
    if rc != 0:
        raise RuntimeError(f"gimic_b200 driver error {rc}: " + D.gimic_b200_driver_last_error().decode(errors="replace"))
    return 0


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
    # switches of the reference front end (src/gimic.in:36-57) that do not touch the hot path; accepted so that existing command
    # lines keep working
    ap.add_argument("-t", "--title", default=None, help="title of job (label only)")
    ap.add_argument("-d", "--debug", type=int, default=None, help="debug level (label only)")
    ap.add_argument("-o", "--output", default=None, help="base name for output file(s) (unused, like in the reference's fgimic backend)")
    ap.add_argument("-b", "--backend", default="fgimic", choices=["fgimic", "gimic"], help="only the fgimic path is provided")
    ap.add_argument("--native", action="store_true", help="hand the run to the compiled driver (libgimic_b200_driver.so, the code behind gimic-b200)")
    a = ap.parse_args(argv)
    if a.native:
        return run_native(a.infile, workdir=a.workdir, dryrun=a.dryrun, vtk_appended=(a.vtk == "appended"), title=a.title)
    if len(a.infile) > 1:
        run_scan(a.infile)
        return 0
    a.infile = a.infile[0]
    if a.dryrun:
        Driver(a.infile, a.workdir, dryrun=True, title=a.title).run()
        return 0
    device = -1
    if "LOCAL_RANK" in os.environ and int(os.environ.get("WORLD_SIZE", "1")) > 1:
        import torch
        import torch.distributed as dist
        device = int(os.environ["LOCAL_RANK"])
        torch.cuda.set_device(device)
        dist.init_process_group("nccl", device_id=torch.device("cuda", device))
    Driver(a.infile, a.workdir, device=device, vtk_appended=(a.vtk == "appended"), title=a.title).run()
    return 0
