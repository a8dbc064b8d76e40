"""ctypes loader for libgimic_b200.so (the C ABI declared in include/gimic_b200.h).

The library is built in-tree by `__graft_entry__.build()` / `make -C gimic_b200/csrc`.  There is no
Python or CPU fallback: if the shared object is missing, or no CUDA device is present when a compute
entry point is called, the call fails loudly.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.environ.get("GIMIC_B200_LIB") or os.path.join(_HERE, "libgimic_b200.so")     # GIMIC_B200_LIB: an alternative build (A/B measurements)

dp = C.POINTER(C.c_double)
ip = C.POINTER(C.c_int)

# every symbol include/gimic_b200.h declares (checked by tests/test_abi.py)
LEGACY_SYMBOLS = ["gimic_init", "gimic_finalize", "gimic_set_uhf", "gimic_set_magnet", "gimic_set_spin",
                  "gimic_set_screening", "gimic_calc_jtensor", "gimic_calc_jvector", "gimic_calc_modj",
                  "gimic_get_gauss_points", "mkgausspoints"]
API_SYMBOLS = ["gimic_b200_default_opts", "gimic_b200_create", "gimic_b200_create_from_arrays", "gimic_b200_destroy", "gimic_b200_device_count",
               "gimic_b200_nbf", "gimic_b200_natoms", "gimic_b200_atom_coords", "gimic_b200_is_uhf",
               "gimic_b200_calc_jtensors", "gimic_b200_calc_basis", "gimic_b200_calc_basis_tiles", "gimic_b200_calc_fields", "gimic_b200_fields_from_tensors", "gimic_b200_jmod_from_jvec",
               "gimic_b200_calc_jtensors_grid", "gimic_b200_partition_points", "gimic_b200_partition_grid", "gimic_b200_partition_calc", "gimic_b200_partition_info", "gimic_b200_integrate", "gimic_b200_integrate_batch", "gimic_b200_property", "gimic_b200_property_integrand", "gimic_b200_gauss_points",
               "gimic_b200_mol_geometry", "gimic_b200_mol_summary", "gimic_b200_c2s_rows", "gimic_b200_convert_xdens", "gimic_b200_format_e", "gimic_b200_format_f",
               "gimic_b200_get_stats", "gimic_b200_set_profiling", "gimic_b200_last_error", "gimic_b200_version"]

ALPHA, BETA, TOTAL, SPINDENS = 0, 1, 2, 3
SPINCASES = {"alpha": ALPHA, "beta": BETA, "total": TOTAL, "spindens": SPINDENS}
DEVICE_PTR = 1


class Opts(C.Structure):
    _fields_ = [("uhf", C.c_int), ("giao", C.c_int), ("diamag", C.c_int), ("paramag", C.c_int), ("screening", C.c_int),
                ("screening_thrs", C.c_double), ("device", C.c_int), ("spherical", C.c_int)]


class GridStruct(C.Structure):
    _fields_ = [("origin", C.c_double * 3), ("basv", C.c_double * 9), ("npts", C.c_int * 3), ("pts", dp * 3),
                ("wgt", dp * 3), ("radius", C.c_double)]


class Stats(C.Structure):
    _fields_ = [("n_points", C.c_long), ("n_tiles", C.c_long), ("sum_nact", C.c_double), ("executed_flops", C.c_double),
                ("dense_flops", C.c_double), ("ms_sort", C.c_float), ("ms_tiles", C.c_float), ("ms_basis", C.c_float),
                ("ms_contract", C.c_float), ("ms_fields", C.c_float), ("ms_total", C.c_float), ("launches", C.c_long),
                ("contract_launches", C.c_long), ("useful_flops", C.c_double), ("ms_plan", C.c_float), ("ms_span", C.c_float),
                ("panel_bytes", C.c_double)]


class GimicB200Error(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"gimic_b200 error {code}: {msg}")
        self.code = code


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(SO_PATH):
        raise ImportError(f"{SO_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                          "or `make -C gimic_b200/csrc` (there is no non-CUDA fallback)")
    L = C.CDLL(SO_PATH, mode=C.RTLD_GLOBAL)
    vp = C.c_void_p
    L.gimic_b200_default_opts.argtypes = [C.POINTER(Opts)]
    L.gimic_b200_create.argtypes = [C.POINTER(vp), C.c_char_p, C.c_char_p, C.POINTER(Opts)]
    L.gimic_b200_create_from_arrays.argtypes = [C.POINTER(vp), C.c_int, dp, ip, ip, ip, dp, dp, C.c_int, vp, vp, C.c_int,
                                                C.POINTER(Opts)]
    L.gimic_b200_destroy.argtypes = [vp]
    for f in ("gimic_b200_nbf", "gimic_b200_natoms", "gimic_b200_is_uhf"):
        getattr(L, f).argtypes = [vp]
    L.gimic_b200_atom_coords.argtypes = [vp, dp]
    L.gimic_b200_calc_jtensors.argtypes = [vp, C.c_long, vp, C.c_int, vp, C.c_int]
    L.gimic_b200_calc_basis.argtypes = [vp, C.c_long, vp, vp, vp, C.c_int]
    L.gimic_b200_calc_basis_tiles.argtypes = [vp, C.c_long, vp, vp, vp, ip]
    L.gimic_b200_calc_fields.argtypes = [vp, C.c_long, vp, dp, C.c_int, vp, vp, vp, vp, vp, vp, C.c_double, C.c_int]
    L.gimic_b200_fields_from_tensors.argtypes = [vp, C.c_long, vp, vp, dp, vp, vp, vp, C.c_int]
    L.gimic_b200_jmod_from_jvec.argtypes = [vp, C.c_long, vp, vp, dp, vp, C.c_int]
    L.gimic_b200_calc_jtensors_grid.argtypes = [vp, C.POINTER(GridStruct), C.c_long, C.c_long, C.c_int, vp, C.c_int]
    lp = C.POINTER(C.c_long)
    L.gimic_b200_partition_points.argtypes = [vp, C.c_long, vp, C.c_int, C.c_int, C.c_int, lp]
    L.gimic_b200_partition_grid.argtypes = [vp, C.POINTER(GridStruct), C.c_int, C.c_int, lp]
    L.gimic_b200_partition_calc.argtypes = [vp, dp, C.c_int, vp, vp, vp, vp, vp, vp, C.c_int]
    L.gimic_b200_partition_info.argtypes = [vp, lp]
    L.gimic_b200_integrate.argtypes = [vp, C.POINTER(GridStruct), dp, C.c_int, C.c_int, C.c_int, C.c_int, dp]
    L.gimic_b200_integrate_batch.argtypes = [vp, C.c_int, C.POINTER(GridStruct), dp, C.c_int, C.c_int, dp]
    L.gimic_b200_property.argtypes = [vp, C.c_long, vp, vp, vp, C.c_int, dp, C.c_int, C.POINTER(C.c_long), dp, C.c_int]
    L.gimic_b200_property_integrand.argtypes = [vp, C.c_long, vp, vp, dp, vp, C.c_int]
    L.gimic_b200_gauss_points.argtypes = [C.c_double, C.c_double, C.c_int, C.c_int, C.c_int, dp, dp]
    L.gimic_b200_c2s_rows.argtypes = [C.c_int, C.c_int, dp]
    L.gimic_b200_mol_geometry.argtypes = [C.c_char_p, C.c_int, dp, C.c_char_p]
    L.gimic_b200_mol_summary.argtypes = [C.c_char_p, ip]
    L.gimic_b200_format_e.restype = C.c_long
    L.gimic_b200_format_e.argtypes = [C.c_long, dp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_char_p, C.c_void_p, C.c_long]
    L.gimic_b200_format_f.restype = C.c_long
    L.gimic_b200_format_f.argtypes = [C.c_long, dp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_char_p, C.c_void_p, C.c_long]
    L.gimic_b200_convert_xdens.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_char_p]
    L.gimic_b200_get_stats.argtypes = [vp, C.POINTER(Stats)]
    L.gimic_b200_set_profiling.argtypes = [vp, C.c_int]
    L.gimic_b200_last_error.restype = C.c_char_p
    L.gimic_b200_version.restype = C.c_char_p
    # legacy boundary (gimic_interface.h / gausspoints.h)
    L.gimic_init.argtypes = [C.c_char_p, C.c_char_p]
    L.gimic_set_uhf.argtypes = [ip]
    L.gimic_set_magnet.argtypes = [dp]
    L.gimic_set_spin.argtypes = [C.c_char_p]
    L.gimic_set_screening.argtypes = [dp]
    L.gimic_calc_jtensor.argtypes = [dp, dp]
    L.gimic_calc_jvector.argtypes = [dp, dp]
    L.gimic_calc_modj.argtypes = [dp, dp]
    L.gimic_get_gauss_points.argtypes = [dp, dp, ip, ip, dp, dp]
    L.mkgausspoints.argtypes = [dp, dp, ip, ip, dp, dp]
    for f in LEGACY_SYMBOLS:
        getattr(L, f).restype = None
    _lib = L
    return L


def check(rc):
    if rc < 0:
        raise GimicB200Error(rc, lib().gimic_b200_last_error().decode(errors="replace"))
    return rc
