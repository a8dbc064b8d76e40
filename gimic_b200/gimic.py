"""Host-side mirror of the reference's Python binding, on top of the batched C ABI.

`Gimic(mol, xdens)` keeps the interface of the Cython class in src/libgimic/gimic.pyx:6-57
(jtensor / jvector / set_property / set_uhf / set_magnet / set_spin / set_screening), which the
reference's pygimic driver calls one point at a time (src/pygimic/field.py:82-93).  The batched
methods (jtensors, fields, jtensors_grid, integrate) are what a driver should use instead.

Arrays may be numpy arrays (host; copied through the C ABI's host path) or CUDA torch tensors
(float64, contiguous; passed zero-copy as device pointers).
"""
import ctypes as C
import numpy as np

from . import _lib
from ._lib import SPINCASES, GimicB200Error  # noqa: F401


def _is_torch_cuda(x):
    return hasattr(x, "data_ptr") and hasattr(x, "is_cuda") and x.is_cuda


def _host(a, shape=None):
    a = np.ascontiguousarray(np.asarray(a, dtype=np.float64))
    return a if shape is None else a.reshape(shape)


def _dptr(a):
    return a.ctypes.data_as(_lib.dp)


class Grid:
    """What gridpoint()/get_weight() of src/fgimic/grid.f90:456-511 need: origin, basis vectors, axis points, weights."""

    def __init__(self, origin, basv, pts, wgt=None, radius=-1.0):
        self.origin = _host(origin, (3,))
        self.basv = _host(basv, (3, 3))          # basv[v] = v-th basis vector
        self.pts = [_host(p).ravel() for p in pts]
        self.wgt = [np.ones_like(p) for p in self.pts] if wgt is None else [_host(w).ravel() for w in wgt]
        self.radius = float(radius)
        self.npts = tuple(int(p.size) for p in self.pts)

    @property
    def n(self):
        return self.npts[0] * self.npts[1] * self.npts[2]

    def struct(self):
        g = _lib.GridStruct()
        for i in range(3):
            g.origin[i] = self.origin[i]
            g.npts[i] = self.npts[i]
            g.pts[i] = _dptr(self.pts[i])
            g.wgt[i] = _dptr(self.wgt[i])
        for v in range(3):
            for c in range(3):
                g.basv[c + 3 * v] = self.basv[v, c]
        g.radius = self.radius
        return g

    def points(self):
        i, j, k = np.meshgrid(self.pts[0], self.pts[1], self.pts[2], indexing="ij")
        r = (self.origin[None, None, None, :] + i[..., None] * self.basv[0] + j[..., None] * self.basv[1]
             + k[..., None] * self.basv[2])
        return np.ascontiguousarray(np.transpose(r, (2, 1, 0, 3)).reshape(-1, 3))  # i fastest


class NotAvailable(NotImplementedError):
    """pygimic.gimic_exceptions.NotAvailable: what the generic connector raises for an operation a backend does not provide"""


class GimicConnector:
    """The backend interface pygimic programs against (src/pygimic/connector.pyx:9-17, connector.pxd:6-9): jvector(r), jtensor(r),
    set_property(name, value).  `Gimic` below is the B200 backend of it."""

    def jvector(self, r):
        raise NotAvailable("jvector()")

    def jtensor(self, r):
        raise NotAvailable("jtensor()")

    def set_property(self, prop, val):
        pass


class Gimic(GimicConnector):
    def __init__(self, mol=None, xdens=None, *, uhf=False, giao=True, diamag=True, paramag=True, screening=True,
                 screening_thrs=1e-6, device=-1, spherical=False, _handle=None):
        L = _lib.lib()
        self._h = C.c_void_p()
        self._magnet = np.zeros(3)
        self._spin = "total"
        self._screening = screening_thrs
        if _handle is not None:
            self._h = _handle
        else:
            o = self._opts(uhf, giao, diamag, paramag, screening, screening_thrs, device, spherical)
            _lib.check(L.gimic_b200_create(C.byref(self._h), str(mol).encode(), str(xdens).encode(), C.byref(o)))
        self.nbf = L.gimic_b200_nbf(self._h)
        self.natoms = L.gimic_b200_natoms(self._h)
        self.uhf = bool(L.gimic_b200_is_uhf(self._h))

    @staticmethod
    def _opts(uhf, giao, diamag, paramag, screening, screening_thrs, device, spherical=False):
        o = _lib.Opts()
        o.uhf, o.giao, o.diamag, o.paramag, o.screening = int(uhf), int(giao), int(diamag), int(paramag), int(screening)
        o.screening_thrs = float(screening_thrs)
        o.device = int(device)
        o.spherical = int(bool(spherical))   # Advanced.spherical (cao2sao.f90): densities over 2l+1 components per shell
        return o

    @classmethod
    def from_arrays(cls, coords, nctr_per_atom, ctr_l, ctr_npf, xp, cc, dens_alpha, dens_beta=None, *,
                    turbomole_order=False, giao=True, diamag=True, paramag=True, screening=True, screening_thrs=1e-6,
                    device=-1, spherical=False):
        """dens_*: 4 matrices in the XDENS layout (flat, element (a,b) at a + nbf*b), numpy or CUDA torch tensor."""
        L = _lib.lib()
        coords = _host(coords)
        nca = np.ascontiguousarray(nctr_per_atom, dtype=np.int32); cl = np.ascontiguousarray(ctr_l, dtype=np.int32)
        cn = np.ascontiguousarray(ctr_npf, dtype=np.int32); xp = _host(xp); cc = _host(cc)
        flags = 0
        if _is_torch_cuda(dens_alpha):
            flags = _lib.DEVICE_PTR
            pa = C.c_void_p(dens_alpha.data_ptr()); pb = C.c_void_p(dens_beta.data_ptr()) if dens_beta is not None else None
            keep = (dens_alpha, dens_beta)
        else:
            da = _host(dens_alpha).ravel(); db = None if dens_beta is None else _host(dens_beta).ravel()
            pa = C.c_void_p(da.ctypes.data); pb = None if db is None else C.c_void_p(db.ctypes.data)
            keep = (da, db)
        o = cls._opts(dens_beta is not None, giao, diamag, paramag, screening, screening_thrs, device, spherical)
        h = C.c_void_p()
        ip = _lib.ip
        _lib.check(L.gimic_b200_create_from_arrays(C.byref(h), coords.shape[0], _dptr(coords), nca.ctypes.data_as(ip),
                                                   cl.ctypes.data_as(ip), cn.ctypes.data_as(ip), _dptr(xp), _dptr(cc),
                                                   int(turbomole_order), pa, pb, flags, C.byref(o)))
        del keep
        return cls(_handle=h)

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            _lib.lib().gimic_b200_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- interface of src/libgimic/gimic.pyx ---------------------------------------------------
    def jtensor(self, r):
        """9 values, jt[m + 3*b] = dJ_m/dB_b (gimic_interface.f90:121-139)."""
        return self.jtensors(np.asarray(r, dtype=np.float64).reshape(1, 3), self._spin)[0]

    def jvector(self, r):
        """J = T.B for the stored magnet, as a list like gimic.pyx:25-34."""
        out = self.fields(np.asarray(r, dtype=np.float64).reshape(1, 3), self._magnet, self._spin, jvec=True)
        return [float(v) for v in out["jvec"][0]]

    def set_property(self, prop, val):
        getattr(self, "set_" + prop)(val)        # gimic.pyx:36-37 (eval-dispatch there)

    def set_uhf(self, onoff):
        if not isinstance(onoff, int):
            raise TypeError
        if bool(onoff) != self.uhf:
            raise GimicB200Error(-1, "open-shell state is fixed at construction: "
                                 "pass uhf=True to Gimic(...) (the reference's set_uhf after init leaves the beta "
                                 "densities unallocated, gimic_interface.f90:80-86)")

    def set_magnet(self, b):
        self._magnet = np.array([b[0], b[1], b[2]], dtype=np.float64)

    def set_spin(self, spin):
        if not isinstance(spin, str):
            raise TypeError
        if spin not in SPINCASES:
            raise ValueError("Invalid spin case.")
        self._spin = spin

    def set_screening(self, thrs):
        if not isinstance(thrs, float):
            raise TypeError
        self._screening = thrs                    # like the reference: recorded, radii are fixed at construction

    # ---- batched API -----------------------------------------------------------------------------
    def atom_coords(self):
        out = np.zeros((self.natoms, 3))
        _lib.check(_lib.lib().gimic_b200_atom_coords(self._h, _dptr(out)))
        return out

    def basis(self, r, want_dr=True):
        """Phi[i, f] and dPhi[i, m, f] in the reference AO order (calc_basis, bfeval.f90:61-122,295-338)."""
        r = _host(r).reshape(-1, 3)
        n = r.shape[0]
        bf = np.empty((n, self.nbf)); dr = np.empty((n, 3, self.nbf)) if want_dr else None
        _lib.check(_lib.lib().gimic_b200_calc_basis(self._h, n, C.c_void_p(r.ctypes.data), C.c_void_p(bf.ctypes.data),
                                                    C.c_void_p(dr.ctypes.data) if want_dr else None, 0))
        return (bf, dr) if want_dr else bf

    def basis_tiles(self, r, want_dr=True):
        """The same vectors as basis(), but evaluated by the hot-path kernels (sort, tiles, k_basis panels) -- diagnostic."""
        r = _host(r).reshape(-1, 3)
        n = r.shape[0]
        bf = np.empty((n, self.nbf)); dr = np.empty((n, 3, self.nbf)) if want_dr else None
        info = (C.c_int * 3)()
        _lib.check(_lib.lib().gimic_b200_calc_basis_tiles(self._h, n, C.c_void_p(r.ctypes.data), C.c_void_p(bf.ctypes.data),
                                                          C.c_void_p(dr.ctypes.data) if want_dr else None, info))
        self.last_tile_info = tuple(info)
        return (bf, dr) if want_dr else bf

    def jtensors(self, r, spincase="total", out=None):
        """tens[i, m + 3*b] for points r[i, :]  (calc_jtensors, jfield.f90:62-138)."""
        L = _lib.lib()
        sc = SPINCASES[spincase]
        if _is_torch_cuda(r):
            import torch
            n = r.shape[0]
            if out is None:
                out = torch.empty((n, 9), dtype=torch.float64, device=r.device)
            assert r.is_contiguous() and out.is_contiguous() and r.dtype == torch.float64
            _lib.check(L.gimic_b200_calc_jtensors(self._h, n, C.c_void_p(r.data_ptr()), sc, C.c_void_p(out.data_ptr()),
                                                  _lib.DEVICE_PTR))
            return out
        r = _host(r).reshape(-1, 3)
        n = r.shape[0]
        if out is None:
            out = np.empty((n, 9))
        _lib.check(L.gimic_b200_calc_jtensors(self._h, n, C.c_void_p(r.ctypes.data), sc, C.c_void_p(out.ctypes.data), 0))
        return out

    def fields(self, r, B, spincase="total", tens=False, jvec=False, jmod=False, acid=False, edens=False, divj=False,
               divj_h=1e-3):
        """Tensors plus derived fields in one pass; returns a dict of the requested arrays (host numpy)."""
        L = _lib.lib()
        r = _host(r).reshape(-1, 3)
        n = r.shape[0]
        B = _host(B, (3,))
        res = {}
        def buf(flag, name, width):
            if not flag:
                return None
            res[name] = np.empty((n, width)) if width > 1 else np.empty(n)
            return C.c_void_p(res[name].ctypes.data)
        args = [buf(tens, "tens", 9), buf(jvec, "jvec", 3), buf(jmod, "jmod", 1), buf(acid, "acid", 1),
                buf(edens, "edens", 1), buf(divj, "divj", 1)]
        _lib.check(L.gimic_b200_calc_fields(self._h, n, C.c_void_p(r.ctypes.data), _dptr(B), SPINCASES[spincase], *args,
                                            float(divj_h), 0))
        return res

    def fields_from_tensors(self, r, tens, B, jvec=True, jmod=False, acid=False):
        L = _lib.lib()
        r = _host(r).reshape(-1, 3); tens = _host(tens).reshape(-1, 9); B = _host(B, (3,))
        n = r.shape[0]
        res = {}
        def buf(flag, name, width):
            if not flag:
                return None
            res[name] = np.empty((n, width)) if width > 1 else np.empty(n)
            return C.c_void_p(res[name].ctypes.data)
        args = [buf(jvec, "jvec", 3), buf(jmod, "jmod", 1), buf(acid, "acid", 1)]
        _lib.check(L.gimic_b200_fields_from_tensors(self._h, n, C.c_void_p(r.ctypes.data), C.c_void_p(tens.ctypes.data),
                                                    _dptr(B), *args, 0))
        return res

    def jmod_from_jvec(self, r, jvec, B):
        """signed |J| (jfield.f90:446-489) of given J vectors"""
        r = _host(r).reshape(-1, 3); jvec = _host(jvec).reshape(-1, 3)
        out = np.empty(r.shape[0])
        _lib.check(_lib.lib().gimic_b200_jmod_from_jvec(self._h, r.shape[0], C.c_void_p(r.ctypes.data), C.c_void_p(jvec.ctypes.data),
                                                        _dptr(_host(B, (3,))), C.c_void_p(out.ctypes.data), 0))
        return out

    def jtensors_grid(self, grid, lo=0, hi=None, spincase="total", out=None):
        """Tensors on the flat index range [lo, hi) of a regular grid; points are generated on the device."""
        L = _lib.lib()
        hi = grid.n if hi is None else hi
        g = grid.struct()
        if out is not None and _is_torch_cuda(out):
            _lib.check(L.gimic_b200_calc_jtensors_grid(self._h, C.byref(g), lo, hi, SPINCASES[spincase],
                                                       C.c_void_p(out.data_ptr()), _lib.DEVICE_PTR))
            return out
        if out is None:
            out = np.empty((hi - lo, 9))
        _lib.check(L.gimic_b200_calc_jtensors_grid(self._h, C.byref(g), lo, hi, SPINCASES[spincase],
                                                   C.c_void_p(out.ctypes.data), 0))
        return out

    # ---- cost-balanced multi-GPU partition (replaces schedule(), parallel.F90:66-84) ----------------------
    def partition(self, points_or_grid, rank=0, nranks=1):
        """Sort and tile the COMPLETE point set (a Grid, an (n, 3) numpy array or a CUDA torch tensor) and take this rank's equal-cost
        share of the tiles.  Returns the number of points this rank owns; evaluate them with partition_calc()."""
        L = _lib.lib()
        cnt = C.c_long(0)
        if isinstance(points_or_grid, Grid):
            g = points_or_grid.struct()
            _lib.check(L.gimic_b200_partition_grid(self._h, C.byref(g), int(rank), int(nranks), C.byref(cnt)))
        elif _is_torch_cuda(points_or_grid):
            r = points_or_grid
            assert r.is_contiguous() and r.shape[-1] == 3
            _lib.check(L.gimic_b200_partition_points(self._h, r.shape[0], C.c_void_p(r.data_ptr()), _lib.DEVICE_PTR, int(rank), int(nranks), C.byref(cnt)))
        else:
            r = _host(points_or_grid).reshape(-1, 3)
            _lib.check(L.gimic_b200_partition_points(self._h, r.shape[0], C.c_void_p(r.ctypes.data), 0, int(rank), int(nranks), C.byref(cnt)))
        self._part_count = int(cnt.value)
        return self._part_count

    def partition_info(self):
        v = (C.c_long * 8)()
        _lib.check(_lib.lib().gimic_b200_partition_info(self._h, v))
        keys = ("points", "owned_points", "tiles", "owned_tiles", "cost_total", "cost_owned", "batches", "first_tile")
        return dict(zip(keys, [int(x) for x in v]))

    def partition_calc(self, B=None, spincase="total", tens=False, jvec=False, jmod=False, acid=False, edens=False, device=None):
        """Evaluate the owned points of the last partition().  Returns a dict with `index` (caller point index / flat grid index of each
        output row) and the requested arrays; numpy by default, CUDA torch tensors on `device` (no host copy) when given."""
        L = _lib.lib()
        m = self._part_count
        Bp = None if B is None else _dptr(_host(B, (3,)))
        res = {}
        if device is not None:
            import torch
            def buf(flag, name, width, dtype=torch.float64):
                if not flag:
                    return None
                res[name] = torch.empty((m, width) if width > 1 else (m,), dtype=dtype, device=device)
                return C.c_void_p(res[name].data_ptr())
            idx = buf(True, "index", 1, torch.int64)
            flags = _lib.DEVICE_PTR
        else:
            def buf(flag, name, width, dtype=np.float64):
                if not flag:
                    return None
                res[name] = np.empty((m, width) if width > 1 else (m,), dtype=dtype)
                return C.c_void_p(res[name].ctypes.data)
            idx = buf(True, "index", 1, np.int64)
            flags = 0
        args = [buf(tens, "tens", 9), buf(jvec, "jvec", 3), buf(jmod, "jmod", 1), buf(acid, "acid", 1), buf(edens, "edens", 1)]
        _lib.check(L.gimic_b200_partition_calc(self._h, Bp, SPINCASES[spincase], idx, *args, flags))
        return res

    def integrate(self, grid, B, spincase="total", what=3, jlo=0, jhi=None):
        """Partial quadrature sums over rows [jlo, jhi): [J, J+, J-, |J|, |J|+, |J|-, sum w*ACID]  (integral.f90)."""
        L = _lib.lib()
        jhi = grid.npts[1] if jhi is None else jhi
        g = grid.struct()
        out = np.zeros(7)
        _lib.check(L.gimic_b200_integrate(self._h, C.byref(g), _dptr(_host(B, (3,))), SPINCASES[spincase], int(what),
                                          int(jlo), int(jhi), _dptr(out)))
        return out

    def integrate_batch(self, grids, Bs, spincase="total", what=3):
        """The quadrature sums of integrate() for many grids in ONE tensor pass (a current-profile scan: the reference runs one
        gimic process per slice, jobscripts/src/current-profile-local-submit).  Bs: one field direction per grid or a single
        one for all.  Returns an (ngrids, 7) array."""
        ng = len(grids)
        out = np.zeros((ng, 7))
        if ng == 0:
            return out
        Bs = _host(Bs)
        Bs = np.ascontiguousarray(np.broadcast_to(Bs.reshape(-1, 3), (ng, 3)))
        arr = (_lib.GridStruct * ng)()
        keep = []
        for i, g in enumerate(grids):
            st = g.struct(); keep.append(st)
            arr[i] = st
        _lib.check(_lib.lib().gimic_b200_integrate_batch(self._h, ng, arr, _dptr(Bs), SPINCASES[spincase], int(what), _dptr(out)))
        return out

    def property(self, r, w, tens, coords, seg_counts=None):
        """Shieldings and magnetizability from a tensor field on a weighted point set (get_property, jfield.f90:584-929).
        seg_counts: points per atom block (nelpts.info, 2nd column); default one segment.  Returns a dict with
        sigma[k] = (xx, yy, zz) in ppm, sigma_pos/neg[k] (already /3 like the reference prints), sigma_atoms[k][s] =
        (total, positive, negative) contribution of point block s, and the same for chi (au)."""
        r = _host(r).reshape(-1, 3); w = _host(w).ravel(); tens = _host(tens).reshape(-1, 9); coords = _host(coords).reshape(-1, 3)
        n, nat = r.shape[0], coords.shape[0]
        seg = np.array([n] if seg_counts is None else np.cumsum(seg_counts), dtype=np.int64)
        nseg = seg.size
        part = np.zeros((nat + 1, nseg, 5))
        _lib.check(_lib.lib().gimic_b200_property(self._h, n, C.c_void_p(r.ctypes.data), C.c_void_p(w.ctypes.data),
                                                  C.c_void_p(tens.ctypes.data), nat, _dptr(coords), nseg,
                                                  seg.ctypes.data_as(C.POINTER(C.c_long)), _dptr(part), 0))
        cum = np.cumsum(part, axis=1)                                   # running sums at the segment ends (scont)
        tot = cum[:, -1, :]
        scont = np.stack([cum[:, :, 0:3].sum(-1) / 3.0, cum[:, :, 3] / 3.0, cum[:, :, 4] / 3.0], -1)
        contrib = np.diff(scont, axis=1, prepend=0.0)                   # what the reference prints per atom block
        return dict(sigma=tot[:nat, 0:3], sigma_iso=tot[:nat, 0:3].sum(1) / 3.0, sigma_pos=tot[:nat, 3] / 3.0,
                    sigma_neg=tot[:nat, 4] / 3.0, sigma_atoms=contrib[:nat], chi=tot[nat, 0:3], chi_iso=tot[nat, 0:3].sum() / 3.0,
                    chi_pos=tot[nat, 3] / 3.0, chi_neg=tot[nat, 4] / 3.0, chi_atoms=contrib[nat])

    def property_integrand(self, r, tens, centre=None):
        """(n, 4) per-point integrands xx, yy, zz and their sum for one nucleus (shielding, ppm) or centre=None (magnetizability):
        the fields the reference plots as sigma<k>.vtu / sigma_xx<k>.vtu ... / intchi*.vtu (jfield.f90:786-808, 915-918)"""
        r = _host(r).reshape(-1, 3); tens = _host(tens).reshape(-1, 9)
        out = np.zeros((r.shape[0], 4))
        c3 = None if centre is None else _dptr(_host(centre, (3,)))
        _lib.check(_lib.lib().gimic_b200_property_integrand(self._h, r.shape[0], C.c_void_p(r.ctypes.data), C.c_void_p(tens.ctypes.data),
                                                            c3, C.c_void_p(out.ctypes.data), 0))
        return out

    def set_profiling(self, on=True):
        _lib.check(_lib.lib().gimic_b200_set_profiling(self._h, int(on)))

    def stats(self):
        s = _lib.Stats()
        _lib.check(_lib.lib().gimic_b200_get_stats(self._h, C.byref(s)))
        return {k: getattr(s, k) for k, _ in s._fields_}


def integrate_distributed(gimic, grid, B, spincase="total", what=3):
    """integral mode over all ranks of torch.distributed: rows j are split in contiguous slabs like
    schedule() of src/fgimic/parallel.F90:66-84, and the partial sums are combined by ONE all-reduce
    (the collect_sum calls of integral.f90:157-161 are commented out in the reference)."""
    import torch
    import torch.distributed as dist
    p2 = grid.npts[1]
    if dist.is_available() and dist.is_initialized():
        rank, world = dist.get_rank(), dist.get_world_size()
    else:
        rank, world = 0, 1
    lo, hi = slab(p2, rank, world)
    part = gimic.integrate(grid, B, spincase, what, lo, hi)
    if world > 1:
        dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
        t = torch.from_numpy(part).to(dev)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        part = t.cpu().numpy()
    return part


def slab(n, rank, world):
    """contiguous block partition of range(n); remainder spread over the first ranks"""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def c2s_rows(l, turbomole_order=False):
    """(2l+1) x ncart(l) cartesian -> spherical projection of cao2sao.f90 (rows m = -l..l), the convention of spherical=on"""
    out = np.zeros((2 * l + 1, (l + 1) * (l + 2) // 2))
    _lib.check(_lib.lib().gimic_b200_c2s_rows(int(l), int(turbomole_order), _dptr(out)))
    return out


def convert_xdens(xdens_text, nbf, xdens_binary, uhf=False):
    """Parse a text XDENS once (all host threads) and write the binary cache Gimic(mol, xdens_binary) loads directly."""
    _lib.check(_lib.lib().gimic_b200_convert_xdens(str(xdens_text).encode(), int(nbf), 8 if uhf else 4, str(xdens_binary).encode()))
