"""Output files in the reference's formats (src/fgimic/vtkplot.f90, jfield.f90:356-364,540, basis.f90 write_xyz,
grid.f90:586-672).  Numbers use Fortran edit descriptors (e14.6, 3e20.10, 6f11.7, f16.10); headers that the
reference writes list-directed are emitted with gfortran-like spacing (the reference's own tests compare numeric
tokens only, test/benzene/3d/test:15-23)."""
import numpy as np

AU2A = float(np.float32(0.52917726))     # globals.f90:51 is a single-precision literal


def _f16(x):
    """F16.10; a coordinate that is not finite (grid_points = 1 along a non-zero length: step = l/0, grid.f90:157) prints like gfortran"""
    x = float(x)
    if x != x:
        return "NaN".rjust(16)
    if x in (float("inf"), -float("inf")):
        return ("-Infinity" if x < 0 else "Infinity").rjust(16)
    return f"{x:16.10f}"


def fortran_e(x, w, d):
    """Fortran Ew.d: 0.dddddE+ee (gfortran drops the 'E' when the exponent needs three digits; asterisks on overflow)"""
    x = float(x)
    if x != x or x in (float("inf"), -float("inf")):          # gfortran: NaN / Infinity / -Infinity, right-justified
        s = "NaN" if x != x else ("-Infinity" if x < 0 else "Infinity")
    elif x == 0.0:
        s = ("-" if str(x)[0] == "-" else "") + "0." + "0" * d + "E+00"      # gfortran keeps the sign of a negative zero
    else:
        m, e = f"{abs(x):.{d - 1}E}".split("E")
        digits = m.replace(".", "")
        e = int(e) + 1
        s = ("-" if x < 0 else "") + "0." + digits + (f"E{e:+03d}" if abs(e) < 100 else f"{e:+04d}")
    return "*" * w if len(s) > w else s.rjust(w)


def _format(kind, values, w, d, per_line, first=0, prefix=""):
    import ctypes as C
    from . import _lib
    v = np.ascontiguousarray(values, dtype=np.float64).ravel()
    n = v.size
    if n == 0:
        return b""
    nlines = n // max(per_line, 1) + 2
    cap = n * w + nlines * (len(prefix) + 1) + 16
    buf = np.empty(cap, dtype=np.uint8)
    fn = _lib.lib().gimic_b200_format_e if kind == "E" else _lib.lib().gimic_b200_format_f
    got = fn(n, v.ctypes.data_as(_lib.dp), w, d, per_line, first, prefix.encode(), C.c_void_p(buf.ctypes.data), cap)
    if got < 0:
        raise _lib.GimicB200Error(got, "number formatting failed")
    return buf[:got].tobytes()


def format_e(values, w, d, per_line, first=0, prefix=""):
    """Many values with Ew.d in one native, threaded call (gimic_b200_format_e): `per_line` values per line (`first` on the
    first line if > 0), every line starting with `prefix`, complete lines ending in a newline.  Returns bytes."""
    return _format("E", values, w, d, per_line, first, prefix)


def format_f(values, w, d, per_line, first=0, prefix=""):
    """Same for the fixed-point descriptor Fw.d (gimic_b200_format_f)."""
    return _format("F", values, w, d, per_line, first, prefix)


def _ld_real(x):
    """gfortran list-directed real(8): 17 significant digits, F form for 1e-1 <= |x| < 1e16"""
    x = float(x)
    ax = abs(x)
    if x != x or ax == float("inf"):
        # e.g. the z spacing of a tilted 2-D grid: vtkplot.f90:33-38 divides a non-zero extent by npts-1 = 0; gfortran prints
        # non-finite values right-justified in the same 25-column field
        return ("NaN" if x != x else ("-Infinity" if x < 0 else "Infinity")).rjust(26)
    if ax != 0.0 and not (0.1 <= ax < 1e16):
        m, e = f"{ax:.16E}".split("E")
        s = ("-" if x < 0 else "") + m + f"E{int(e):+04d}"
        return "  " + s + " "
    nint = len(str(int(ax))) if ax >= 1.0 else 0
    dec = 17 - max(nint, 0) if ax >= 1.0 else 17
    if ax == 0.0:
        dec = 16                      # gfortran prints zero as 0.0000000000000000 (test/benzene/2d/reference/jvec.vti)
    s = f"{ax:.{dec}f}"
    if ax < 1.0:
        s = s  # 0.ddddddddddddddddd
    s = ("-" if x < 0 else "") + s
    return s.rjust(21) + "     "


def _ld_int(i):
    return f"{int(i):12d}"


class _Text:
    """text-mode write() on a binary file (the bulk number blocks are written as bytes)"""

    def __init__(self, fb):
        self.fb = fb

    def write(self, s):
        self.fb.write(s.encode("ascii"))


def _vti_header(f, npts, qmin, step, name, ncomp):
    ext = "".join(_ld_int(v) for v in (0, npts[0] - 1, 0, npts[1] - 1, 0, npts[2] - 1))
    f.write('<?xml version="1.0"?>\n')
    f.write(' <VTKFile type="ImageData" version="0.1" byte_order="LittleEndian">\n')
    # gfortran separates a numeric list item from a following character item by one blank (test/*/reference/*.vti)
    f.write('   <ImageData WholeExtent="' + ext + ' " Origin="' + "".join(_ld_real(v) for v in qmin) + ' " Spacing="'
            + "".join(_ld_real(v) for v in step) + ' ">\n')
    f.write('   <Piece Extent="' + ext + ' ">\n')
    f.write('   <PointData Scalars="scalars">\n')
    f.write(f'   <DataArray Name="{name}" type="Float64" NumberOfComponents="{ncomp}" Format="ascii">\n')


def _vti_geometry(grid):
    npts = grid.npts
    qmin = grid.gridpoint(0, 0, 0)
    qmax = grid.gridpoint(npts[0] - 1, npts[1] - 1, npts[2] - 1)
    step = qmax - qmin                                          # vtkplot.f90:33-38
    for i in range(3):
        if step[i] > 1e-8:
            step[i] = step[i] / (npts[i] - 1) if npts[i] > 1 else float("inf")     # IEEE x/0 like the Fortran
    return qmin, step


def _cell_average_norm(v, p1, p2, p3):
    """|J| averaged over the corners of each cell (vtkplot.f90:132-226); None for grids without cells"""
    sl = [slice(0, -1), slice(1, None)]
    if p1 > 1 and p2 > 1 and p3 > 1:
        avg = sum(v[a, b, c] for a in sl for b in sl for c in sl) / 8.0
    elif p1 > 1 and p2 > 1 and p3 == 1:
        avg = sum(v[0:1, b, c] for b in sl for c in sl) / 4.0
    elif p1 > 1 and p2 == 1 and p3 > 1:
        avg = sum(v[a, 0:1, c] for a in sl for c in sl) / 4.0
    elif p1 == 1 and p2 > 1 and p3 > 1:
        avg = sum(v[a, b, 0:1] for a in sl for b in sl) / 4.0
    else:
        return None
    return np.sqrt((avg ** 2).sum(-1)).ravel()


def _vti_appended(path, grid, name, ncomp, data, cell=None):
    """EXTRA (not a reference format): the same ImageData file with raw appended Float64 blocks instead of ASCII numbers --
    8 bytes per value instead of 14 characters, no formatting cost; ParaView/VTK read both."""
    qmin, step = _vti_geometry(grid)
    npts = grid.npts
    ext = " ".join(str(v) for v in (0, npts[0] - 1, 0, npts[1] - 1, 0, npts[2] - 1))
    blocks = [np.ascontiguousarray(data, dtype="<f8").ravel()]
    head = ['<?xml version="1.0"?>', '<VTKFile type="ImageData" version="0.1" byte_order="LittleEndian" header_type="UInt64">',
            f'  <ImageData WholeExtent="{ext}" Origin="{float(qmin[0])!r} {float(qmin[1])!r} {float(qmin[2])!r}" Spacing="{float(step[0])!r} {float(step[1])!r} {float(step[2])!r}">',
            f'    <Piece Extent="{ext}">', f'      <PointData {"Vectors" if ncomp == 3 else "Scalars"}="{name}">',
            f'        <DataArray Name="{name}" type="Float64" NumberOfComponents="{ncomp}" format="appended" offset="0"/>',
            '      </PointData>']
    if cell is not None:
        off = 8 + blocks[0].nbytes
        blocks.append(np.ascontiguousarray(cell, dtype="<f8").ravel())
        head += ['      <CellData Scalars="cell_norm">',
                 f'        <DataArray Name="cell_norm" type="Float64" NumberOfComponents="1" format="appended" offset="{off}"/>', '      </CellData>']
    head += ['    </Piece>', '  </ImageData>', '  <AppendedData encoding="raw">']
    with open(path, "wb") as fb:
        fb.write(("\n".join(head) + "\n_").encode("ascii"))
        for b in blocks:
            fb.write(np.uint64(b.nbytes).tobytes()); fb.write(b.tobytes())
        fb.write(b"\n  </AppendedData>\n</VTKFile>\n")


def write_vti_scalar(path, grid, values, appended=False):
    """write_vtk_imagedata, vtkplot.f90:14-86: values[i + p1*(j + p2*k)], e14.6, line break when mod(l,4)==0"""
    v = np.asarray(values, dtype=np.float64).ravel()
    if appended:
        return _vti_appended(path, grid, "scalars", 1, v)
    qmin, step = _vti_geometry(grid)
    with open(path, "wb") as fb:
        f = _Text(fb)
        _vti_header(f, grid.npts, qmin, step, "scalars", 1)
        fb.write(format_e(v, 14, 6, 4, first=1))              # a line break after value l when mod(l,4)==0 (0-based)
        f.write("\n    </DataArray>\n    </PointData>\n    </Piece>\n    </ImageData>\n </VTKFile>\n")


def write_vti_vector(path, grid, vec, appended=False):
    """write_vtk_vector_imagedata, vtkplot.f90:88-234: 3e14.6 per point + CellData of cell-averaged |J|"""
    p1, p2, p3 = grid.npts
    v = np.asarray(vec, dtype=np.float64).reshape(p3, p2, p1, 3)
    nrm = _cell_average_norm(v, p1, p2, p3)
    if appended:
        return _vti_appended(path, grid, "vectors", 3, v, nrm)
    qmin, step = _vti_geometry(grid)
    with open(path, "wb") as fb:
        f = _Text(fb)
        _vti_header(f, grid.npts, qmin, step, "vectors", 3)
        fb.write(format_e(v, 14, 6, 3))
        f.write("\n    </DataArray>\n    </PointData>\n    <CellData Scalars=\"foo\">\n")
        if nrm is not None:
            fb.write(format_e(nrm, 14, 6, 1))
        f.write("    </CellData>\n    </Piece>\n    </ImageData>\n </VTKFile>\n")


def radius_masked_vectors(grid, vec):
    """cdens visualisation with the `radius` keyword (jfield.f90:310-346): on 2-D bond grids the vectors written to jvec.vti are
    zeroed where |coord*AU2A - center| > radius.  The reference compares Angstrom coordinates with the bohr centre and radius
    (unit slip); replicated as is.  Returns vec itself when the rule does not apply."""
    if grid.mode != "bond" or grid.is_3d() or not (grid.radius > 0.1) or grid.radius >= 1.0e10:
        return vec
    coord = grid.points() * AU2A
    out = np.array(vec, dtype=np.float64, copy=True).reshape(-1, 3)
    out[np.sqrt(((coord - grid.center()) ** 2).sum(1)) > grid.radius] = 0.0
    return out


def read_ele(path):
    """TetGen .ele: first line 'ncells 4 0', then 'idx n1 n2 n3 n4' (jfield.f90:421-431)"""
    with open(path) as f:
        ncells = int(f.readline().split()[0])
        cells = np.loadtxt(f, dtype=np.int64, max_rows=ncells).reshape(-1, 5)[:, 1:5]
    return cells


def _vtu_write(path, points, name, ncomp, data, cells):
    pts = np.asarray(points, dtype=np.float64).reshape(-1, 3)
    v = np.asarray(data, dtype=np.float64).reshape(pts.shape[0], ncomp)
    nc = cells.shape[0]
    with open(path, "wb") as fb:
        f = _Text(fb)
        f.write('<?xml version="1.0"?>\n<VTKFile type="UnstructuredGrid" version="0.1" byte_order="LittleEndian">\n  <UnstructuredGrid>\n')
        f.write(f'    <Piece NumberOfPoints="{pts.shape[0]:10d}" NumberOfCells="{nc:10d}">\n      <Points>\n')
        f.write('        <DataArray type="Float32" NumberOfComponents="3" Format="ascii">\n')
        fb.write(format_e(pts, 20, 10, 3, prefix="        "))
        f.write('        </DataArray>\n      </Points>\n      <PointData Scalars="scalars">\n')
        f.write(f'        <DataArray Name="{name}" type="Float64" NumberOfComponents="{ncomp}" Format="ascii">\n')
        fb.write(format_e(v, 20, 10, ncomp, prefix="        "))
        f.write('        </DataArray>\n      </PointData>\n      <Cells>\n        <DataArray type="Int32" Name="connectivity" Format="ascii">\n')
        f.write("".join("        " + "".join(f"{int(i) - 1:10d}" for i in c) + "\n" for c in cells))
        f.write('        </DataArray>\n        <DataArray type="Int32" Name="offsets" Format="ascii">\n        ')
        f.write("".join(f"{4 * (c + 1):10d}" for c in range(nc)) + "\n")
        f.write('        </DataArray>\n        <DataArray type="Int32" Name="types" Format="ascii">\n        ')
        f.write("".join(f"{10:5d}" for _ in range(nc)) + "\n")
        f.write('        </DataArray>\n      </Cells>\n      <CellData Scalars="foo">\n        ')
        f.write("".join(" 0.0" for _ in range(nc)) + "\n")
        f.write("      </CellData>\n    </Piece>\n  </UnstructuredGrid>\n</VTKFile>\n")


def write_vtu_vector(path, points, vec, cells):
    """write_vtk_vector_unstructuredgrid, vtkplot.f90:241-312"""
    _vtu_write(path, points, "vectors", 3, vec, cells)


def write_vtu_scalar(path, points, values, cells):
    """write_vtk_scalar_unstructuredgrid, vtkplot.f90:318-391"""
    _vtu_write(path, points, "scalars", 1, values, cells)


def write_jmod_txt(path, grid, vec, regular=True):
    """jmod<tag>.txt on Gauss grids (jfield.f90:294-301,356-376,531-541): '(6f11.7)' of coord*AU2A and |J|;
    a blank line after each i-row on regular grids"""
    v = np.asarray(vec, dtype=np.float64).reshape(-1, 3)
    r = grid.points() * AU2A
    jm = np.sqrt((v ** 2).sum(1))
    p1 = grid.npts[0]
    n = v.shape[0]
    rows = np.frombuffer(format_f(np.column_stack([r, jm]), 11, 7, 4), dtype=np.uint8).reshape(n, 45)   # 4 x f11.7 + newline
    with open(path, "wb") as f:
        if regular and n % p1 == 0:
            blocks = np.concatenate([rows.reshape(n // p1, p1 * 45), np.full((n // p1, 1), 10, np.uint8)], axis=1)   # blank line per i-row
            f.write(blocks.tobytes())
        elif regular:
            for k in range(0, n, p1):
                f.write(rows[k:k + p1].tobytes())
                if k + p1 <= n:
                    f.write(b"\n")
        else:
            f.write(rows.tobytes())


def write_mol_xyz(path, symbols, coords):
    """write_xyz (basis.f90): natoms, blank, 'sym x y z' in Angstrom"""
    with open(path, "w") as f:
        f.write(f"{len(symbols):12d}\n\n")
        for s, c in zip(symbols, coords):
            f.write(f"{s}" + "".join(_f16(x * AU2A) for x in c) + "\n")


def write_grid_xyz(path, grid, symbols, coords):
    """plot_grid_xyz, grid.f90:586-672: atoms + grid corners ('X') + field-direction marker ('Be')"""
    p1, p2, p3 = grid.npts
    # file grids take the p3 == 1 branch like every grid with npts = (n, 1, 1): gridpoint(i,j,k) = xdata(:,i), so the four 'X' lines are the
    # first and the last point twice, and the count line says natoms+5 although no 'Be' line follows (no case for 'file', grid.f90:660-667)
    if p3 > 1:
        idx = [(0, 0, 0), (p1 - 1, 0, 0), (0, p2 - 1, 0), (0, 0, p3 - 1), (p1 - 1, p2 - 1, 0), (p1 - 1, 0, p3 - 1), (0, p2 - 1, p3 - 1),
               (p1 - 1, p2 - 1, p3 - 1)]
        corners = [grid.gridpoint(*i) for i in idx]
    else:
        corners = [grid.gridpoint(*i) for i in [(0, 0, 0), (p1 - 1, 0, 0), (0, p2 - 1, 0), (p1 - 1, p2 - 1, 0)]]
    marker = None
    if grid.mode in ("std", "base"):
        marker = grid.origin + grid.basv[2] * 2.0
    elif grid.mode == "bond":
        marker = grid.origin + grid.ortho * 2.0
    with open(path, "w") as f:
        f.write(f"{len(symbols) + len(corners) + 1:12d}\n\n")
        for s, c in zip(symbols, coords):
            f.write(f"{s}" + "".join(_f16(x * AU2A) for x in c) + "\n")
        for c in corners:
            f.write("X " + "".join(_f16(x * AU2A) for x in c) + "\n")
        if marker is not None:
            f.write("Be " + "".join(_f16(x * AU2A) for x in marker) + "\n")
