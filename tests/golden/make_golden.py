#!/usr/bin/env python3
"""Generate tests/golden/* from the reference's own test inputs and golden outputs.

Run once in the dev container (where /root/reference is mounted):
    python tests/golden/make_golden.py
The GPU box has no /root/reference, so everything the tests need is committed here:
  * MOL files (text, as the reference ships them)
  * XDENS contents as compressed float64 arrays (the text parse is exact: %.14E -> double)
  * the reference's golden numbers, parsed from its .vtu/.vti/stdout files
The open-shell .vti goldens are sub-sampled (every 7th grid point) to keep the fixture small;
test_oracle_golden.py uses the complete files when /root/reference is present.
"""
import json
import os
import re
import shutil
import numpy as np

REF = os.environ.get("GIMIC_REFERENCE", "/root/reference")
OUT = os.path.dirname(os.path.abspath(__file__))


def read_xdens(path):
    with open(path) as f:
        vals = [float(t.replace("D", "E").replace("d", "e")) for t in f.read().split()]
    return np.array(vals)


def read_vtu_vectors(path):
    lines = open(path).read().split("\n")
    out, grab, pts, grab_pts = [], False, [], False
    for l in lines:
        if "<Points>" in l:
            grab_pts = True; continue
        if 'Name="vectors"' in l:
            grab = True; continue
        if "</DataArray>" in l:
            grab = False; grab_pts = False; continue
        if l.lstrip().startswith("<"):
            continue
        if grab:
            out.append([float(t) for t in l.split()])
        elif grab_pts:
            pts.append([float(t) for t in l.split()])
    return np.array(pts), np.array(out)


def read_vti(path):
    """first DataArray of PointData (vectors: n x 3, scalars: flat n)"""
    vals, grab = [], False
    ncomp = 1
    for l in open(path):
        if "<DataArray" in l and not vals and not grab:
            grab = True
            ncomp = int(re.search(r'NumberOfComponents="(\d+)"', l).group(1))
            continue
        if "</DataArray>" in l and grab:
            break
        if grab:
            vals.extend(float(t) for t in l.split())
    a = np.array(vals)
    return a.reshape(-1, 3) if ncomp == 3 else a


def parse_integral_stdout(path):
    """all 'Induced ... (au)', Positive, Negative blocks in order of appearance"""
    txt = open(path, encoding="utf-8", errors="replace").read().split("\n")
    blocks, cur, section, spin = [], None, None, "total"
    for l in txt:
        if "*** Integrating |J|" in l:
            section = "modulus"
        elif "*** Integrating current" in l:
            section = "current"
        elif "*** Integrating ACID" in l:
            section = "acid"
        m = re.search(r"Integrating (total|alpha|beta|spin) density", l)
        if m and "current density" not in l:
            spin = {"spin": "spindens"}.get(m.group(1), m.group(1))
        m = re.search(r"Induced (mod )?current \(au\)\s*:\s*([-\d.]+)", l)
        if m:
            cur = dict(section=section, spin=spin, au=float(m.group(2)))
        m = re.search(r"Positive contribution:\s*([-\d.]+)\s*\(\s*([-\d.]+)", l)
        if m and cur is not None:
            cur["pos"] = float(m.group(1)); cur["pos_si"] = float(m.group(2))
        m = re.search(r"Negative contribution:\s*([-\d.]+)\s*\(\s*([-\d.]+)", l)
        if m and cur is not None:
            cur["neg"] = float(m.group(1)); cur["neg_si"] = float(m.group(2))
        m = re.search(r"Induced (mod )?current \(nA/T\)\s*:\s*([-\d.]+)", l)
        if m and cur is not None:
            cur["si"] = float(m.group(2)); blocks.append(cur); cur = None; spin = "total"
    geo = {}
    for l in txt:
        m = re.match(r"\s*(center|origin|basv1|basv2|basv3|lenghts|magnet)\s+([-\d.]+)\s+([-\d.]+)\s+([-\d.]+)\s*$", l)
        if m:
            geo[m.group(1)] = [float(m.group(i)) for i in (2, 3, 4)]
    return dict(blocks=blocks, geometry=geo)


_NUMLINE = re.compile(r"^\s*(?:[-+]?(?:\d+\.\d*|\.\d+|\d+)(?:[EeDd]?[-+]\d+)?\s*)+$")


def file_frame(path, coord_chars=0):
    """The non-numeric skeleton of an output file: text lines verbatim ("T:<line>"), lines that hold only numbers as "D:<tokens>:<length>",
    run-length encoded.  coord_chars > 0: also a SHA-256 over the first coord_chars characters of every numeric line (the coordinate
    columns of jmod.txt, which do not depend on the densities)."""
    import hashlib
    kinds, h = [], hashlib.sha256()
    for line in open(path, encoding="utf-8", errors="replace").read().split("\n"):
        if line.strip() and _NUMLINE.match(line):
            kinds.append(f"D:{len(line.split())}:{len(line)}")
            if coord_chars:
                h.update(line[:coord_chars].encode() + b"\n")
        else:
            kinds.append("T:" + line)
    rle = []
    for k in kinds:
        if rle and rle[-1][0] == k:
            rle[-1][1] += 1
        else:
            rle.append([k, 1])
    out = {"rle": rle}
    if coord_chars:
        out["coord_sha256"] = h.hexdigest()
    return out


def main():
    t = os.path.join(REF, "test")
    shutil.copyfile(os.path.join(t, "c4h4/MOL"), os.path.join(OUT, "c4h4_MOL"))
    shutil.copyfile(os.path.join(t, "open-shell/MOL"), os.path.join(OUT, "open_shell_MOL"))
    shutil.copyfile(os.path.join(t, "benzene/MOL"), os.path.join(OUT, "benzene_MOL"))
    np.savez_compressed(os.path.join(OUT, "c4h4_xdens.npz"), xdens=read_xdens(os.path.join(t, "c4h4/XDENS")))
    np.savez_compressed(os.path.join(OUT, "open_shell_xdens.npz"), xdens=read_xdens(os.path.join(t, "open-shell/XDENS")))

    grid = np.loadtxt(os.path.join(t, "c4h4/read-grid/gridfile.grd"))
    pts, jvec = read_vtu_vectors(os.path.join(t, "c4h4/read-grid/reference/jvec.vtu"))
    assert grid.shape == pts.shape == jvec.shape == (4110, 3)
    np.savez_compressed(os.path.join(OUT, "c4h4_readgrid.npz"), grid=grid, jvec=jvec, magnet=np.array([0.0, 0.0, -1.0]))

    json.dump(parse_integral_stdout(os.path.join(t, "c4h4/integration/reference/stdout")),
              open(os.path.join(OUT, "c4h4_integration.json"), "w"), indent=1)
    json.dump(parse_integral_stdout(os.path.join(t, "open-shell/integration/reference/stdout")),
              open(os.path.join(OUT, "open_shell_integration.json"), "w"), indent=1)

    # the reports themselves from two lines above 'TITLE:' on (i.e. without the program banner), for line-by-line comparison
    os.makedirs(os.path.join(OUT, "stdout"), exist_ok=True)
    reports = [("c4h4/integration", "c4h4_integration"), ("open-shell/integration", "open-shell_integration")]
    reports += [("benzene/" + n, "benzene_" + n) for n in sorted(os.listdir(os.path.join(t, "benzene")))
                if os.path.exists(os.path.join(t, "benzene", n, "reference", "stdout"))]
    for src, dst in reports:
        lines = open(os.path.join(t, src, "reference", "stdout"), encoding="utf-8", errors="replace").read().split("\n")
        k = next(i for i, l in enumerate(lines) if l.strip().startswith("TITLE:"))
        open(os.path.join(OUT, "stdout", dst + ".txt"), "w").write("\n".join(lines[k - 2:]))

    d = {}
    n = 33 ** 3
    idx = np.arange(0, n, 7)
    d["index"] = idx
    for tag in ("", "alpha", "beta", "spindens"):
        jv = read_vti(os.path.join(t, f"open-shell/3d/reference/jvec{tag}.vti"))
        jm = read_vti(os.path.join(t, f"open-shell/3d/reference/jmod{tag}.vti"))
        assert jv.shape == (n, 3) and jm.shape == (n,), (jv.shape, jm.shape)
        d[f"jvec{tag}"] = jv[idx]; d[f"jmod{tag}"] = jm[idx]
    np.savez_compressed(os.path.join(OUT, "open_shell_3d.npz"), **d)
    # skeletons of the reference's output files (every byte that is not a data value: XML boilerplate, list-directed header numbers,
    # block sizes, tokens per line, line lengths); benzene files included -- their skeleton only depends on MOL + gimic.inp
    frames = {}
    for rel, cc in (("benzene/2d/reference/jvec.vti", 0), ("benzene/2d-keyword-magnet/reference/jvec.vti", 0), ("benzene/vectors/reference/jvec.vti", 0),
                    ("benzene/3d/reference/jvec.vti", 0), ("benzene/3d/reference/jmod.vti", 0), ("benzene/3d/reference/acid.vti", 0),
                    ("benzene/3d-keyword-magnet/reference/jvec.vti", 0), ("benzene/int-cdens/reference/jmod.txt", 33),
                    ("open-shell/3d/reference/jvec.vti", 0), ("open-shell/3d/reference/jmodspindens.vti", 0)):
        frames[rel] = file_frame(os.path.join(t, rel), cc)
    # jvec.vtu: the Points block holds the grid coordinates (independent of the densities): hash its lines as well
    import hashlib
    vtu = os.path.join(t, "c4h4/read-grid/reference/jvec.vtu")
    frames["c4h4/read-grid/reference/jvec.vtu"] = file_frame(vtu)
    hp = hashlib.sha256()
    for line in open(vtu).read().split("\n")[6:6 + 4110]:
        hp.update(line.encode() + b"\n")
    frames["c4h4/read-grid/reference/jvec.vtu"]["points_sha256"] = hp.hexdigest()
    json.dump(frames, open(os.path.join(OUT, "file_frames.json"), "w"), indent=0)
    # grid geometry / point counts / field direction printed by the reference for the benzene keyword tests
    # (their XDENS is a stripped blob, but these lines only depend on MOL + gimic.inp)
    grids = {}
    for name in ("int-grid-bond-even", "keyword-rotation", "keyword-rotation_origin", "keyword-radius", "keyword-spacing",
                 "keyword-magnet", "integration-lobatto", "integration-gauss", "int-cdens", "3d", "2d", "3d-keyword-magnet",
                 "2d-keyword-magnet", "vectors"):
        path = os.path.join(t, "benzene", name, "reference", "stdout")
        if not os.path.exists(path):
            continue
        txt = open(path, encoding="utf-8", errors="replace").read()
        d = parse_integral_stdout(path)
        m = re.search(r"Number of grid points <v1,v2>:\s*(\d+)\s+(\d+)\s+(\d+)", txt)
        g = dict(geometry=d["geometry"], blocks=d["blocks"])
        if m:
            g["npts"] = [int(m.group(i)) for i in (1, 2, 3)]
        m = re.findall(r"Magnetic field <x,y,z> =\s*([-\d.]+)\s+([-\d.]+)\s+([-\d.]+)", txt)
        if m:
            g["magnet"] = [float(v) for v in m[0]]
        grids[name] = g
        shutil.copyfile(os.path.join(t, "benzene", name, "gimic.inp"), os.path.join(OUT, "inputs", f"benzene_{name}.inp"))
    # inputs of the remaining benzene tests (no stdout geometry block is parsed for them)
    for name, fname in (("diamag-off", "gimic.inp"), ("giao-test", "gimic.inp"), ("paramag-off", "gimic.inp"),
                        ("skip-jmod-integration", "gimic.inp"), ("vectors", "vectors.inp"), ("magnetizability", "gimic.inp")):
        shutil.copyfile(os.path.join(t, "benzene", name, fname), os.path.join(OUT, "inputs", f"benzene_{name}.inp"))
    shutil.copyfile(os.path.join(t, "benzene", "magnetizability", "coord.au"), os.path.join(OUT, "benzene_coord.au"))
    json.dump(grids, open(os.path.join(OUT, "benzene_grids.json"), "w"), indent=1)
    print("golden fixtures written to", OUT)
    for f in sorted(os.listdir(OUT)):
        print(f"  {f:32s} {os.path.getsize(os.path.join(OUT, f)):9d} B")


if __name__ == "__main__":
    main()
