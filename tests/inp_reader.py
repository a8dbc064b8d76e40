"""gimic.inp reader: the surface syntax of the reference's getkw grammar (doc/input.rst:4-19, src/fgimic/getkw.py)
and the schema / defaults / cross-checks of src/gimic.in:59-147,161-283.

    key=value          key=[a, b, c]          Section(arg) { ... }          # comment          '|' continues a line

Booleans accept on/off/true/false/yes/no/0/1 (getkw.py:32,280-281); reals accept Fortran exponents (1.d-8).
`Input.is_set(path)` mirrors keyword_is_set() of the Fortran reader, which the grid code depends on
(e.g. grid.f90:93,199-202,253).
"""
import re

_TRUE = {"on", "true", "yes", "1", "t", "y"}
_FALSE = {"off", "false", "no", "0", "f", "n"}

# name -> (type, default)  ; type in STR INT DBL BOOL INT_ARRAY DBL_ARRAY  (src/gimic.in:60-112)
SCHEMA = {
    "": {"title": ("STR", ""), "debug": ("INT", 0), "calc": ("STR", None), "backend": ("STR", "gimic"),
         "basis": ("STR", "mol"), "density": ("STR", ""), "mofile": ("STR", ""), "mos": ("INT_ARRAY", [0, 0]),
         "xdens": ("STR", "XDENS"), "magnet_axis": ("STR", ""), "magnet": ("DBL_ARRAY", [0.0, 0.0, 0.0]),
         "openshell": ("BOOL", False), "dryrun": ("BOOL", False)},
    "Advanced": {"screening": ("BOOL", False), "screening_thrs": ("DBL", 1.0e-8), "spherical": ("BOOL", True),
                 "GIAO": ("BOOL", True), "diamag": ("BOOL", True), "paramag": ("BOOL", True), "lip_order": ("INT", 3)},
    "Essential": {"acid": ("BOOL", False), "jmod": ("BOOL", False), "prop": ("BOOL", False)},
    "Grid": {"type": ("STR", "even"), "file": ("STR", None), "origin": ("DBL_ARRAY", None), "ivec": ("DBL_ARRAY", None),
             "jvec": ("DBL_ARRAY", None), "lengths": ("DBL_ARRAY", None), "bond": ("INT_ARRAY", None),
             "fixpoint": ("INT", None), "coord1": ("DBL_ARRAY", None), "coord2": ("DBL_ARRAY", None),
             "fixcoord": ("DBL_ARRAY", None), "distance": ("DBL", None), "rotation": ("DBL_ARRAY", [0.0, 0.0, 0.0]),
             "rotation_origin": ("DBL_ARRAY", [0.0, 0.0, 0.0]), "spacing": ("DBL_ARRAY", None),
             "height": ("DBL_ARRAY", None), "width": ("DBL_ARRAY", None), "radius": ("DBL", -1.0),
             "gridplot": ("INT", None), "grid_points": ("INT_ARRAY", None), "gauss_order": ("INT", 7)},
}


class InputError(ValueError):
    pass


def _num(tok):
    return float(re.sub(r"[dD]", "e", tok))


def _convert(typ, raw, name):
    def one(t, v):
        v = v.strip().strip('"').strip("'")
        if t == "STR":
            return v
        if t == "INT":
            return int(_num(v))
        if t == "DBL":
            return _num(v)
        if t == "BOOL":
            lv = v.lower()
            if lv in _TRUE:
                return True
            if lv in _FALSE:
                return False
            raise InputError(f"invalid boolean '{v}' for {name}")
        raise InputError(f"unknown type {t}")
    if typ.endswith("_ARRAY"):
        vals = raw if isinstance(raw, list) else [raw]
        return [one(typ[:-6], v) for v in vals]
    if isinstance(raw, list):
        if len(raw) != 1:
            raise InputError(f"{name} expects a scalar")
        raw = raw[0]
    return one(typ, raw)


class Input:
    def __init__(self):
        self.values = {s: {k: d for k, (t, d) in kws.items()} for s, kws in SCHEMA.items()}
        self._set = set()
        self.grid_arg = "std"          # grid.set_arg('STR', ('std',)), src/gimic.in:92
        self.grid_present = False

    def get(self, path):
        sect, _, key = path.rpartition(".")
        return self.values[sect][key]

    def is_set(self, path):
        return path in self._set

    def _assign(self, sect, key, raw):
        if sect not in SCHEMA or key not in SCHEMA[sect]:
            if sect == "Gimlet" or sect.startswith("Gimlet"):
                return
            raise InputError(f"unknown keyword '{key}' in section '{sect or 'top'}'")
        typ = SCHEMA[sect][key][0]
        self.values[sect][key] = _convert(typ, raw, key)
        self._set.add(f"{sect}.{key}" if sect else key)


def _strip_comments(text):
    out = []
    for line in text.split("\n"):
        q = None
        buf = []
        for ch in line:
            if q:
                buf.append(ch)
                if ch == q:
                    q = None
            elif ch in "\"'":
                q = ch
                buf.append(ch)
            elif ch == "#":
                break
            else:
                buf.append(ch)
        out.append("".join(buf))
    text = "\n".join(out)
    return re.sub(r"\|\s*\n", " ", text)      # '|' line continuation


_TOKEN = re.compile(r"""\s*(?:(?P<sect>[A-Za-z_]\w*)\s*(?:\(\s*(?P<arg>[^)]*)\))?\s*\{|(?P<close>\})|(?P<key>[A-Za-z_]\w*)\s*=\s*(?P<val>\[[^\]]*\]|"[^"]*"|'[^']*'|[^\s{}]+))""")


def parse_text(text):
    inp = Input()
    text = _strip_comments(text)
    pos, stack = 0, [""]
    while True:
        m = _TOKEN.match(text, pos)
        if not m:
            if text[pos:].strip():
                raise InputError(f"cannot parse input near: {text[pos:pos + 40]!r}")
            break
        pos = m.end()
        if m.group("sect"):
            name = m.group("sect")
            if name not in SCHEMA and name != "Gimlet":
                raise InputError(f"unknown section '{name}'")
            if name == "Grid":
                inp.grid_present = True
                if m.group("arg") is not None and m.group("arg").strip():
                    inp.grid_arg = m.group("arg").strip().strip('"').strip("'")
            stack.append(name)
        elif m.group("close"):
            if len(stack) == 1:
                raise InputError("unbalanced '}'")
            stack.pop()
        else:
            val = m.group("val")
            if val.startswith("["):
                raw = [v for v in re.split(r"[,\s]+", val[1:-1].strip()) if v]
            else:
                raw = val
            inp._assign(stack[-1], m.group("key"), raw)
    if len(stack) != 1:
        raise InputError("unbalanced '{'")
    _check(inp)
    return inp


def parse_file(path):
    with open(path) as f:
        return parse_text(f.read())


def _check(inp):
    """check_top / check_grid of src/gimic.in:161-283 (errors raise instead of sys.exit)."""
    calc = inp.get("calc")
    if calc not in ("cdens", "integral", "divj", "edens"):
        raise InputError(f"Error: unknown option calc = {calc}")
    om, pm = inp.is_set("magnet"), inp.is_set("magnet_axis")
    if om and pm:
        raise InputError("Error: Both magnet vector and axis set simultaneously!")
    if not om and not pm:
        raise InputError("Error: Direction of magnetic field must be set!")
    if pm and inp.get("magnet_axis") in ("T", "X"):
        inp.values[""]["magnet_axis"] = "X"
    if not inp.grid_present:
        raise InputError("Error: no Grid section")
    arg = inp.grid_arg
    S = lambda k: inp.is_set("Grid." + k)
    if arg in ("std", "base"):
        for k in ("origin", "ivec", "jvec", "lengths"):
            if not S(k):
                raise InputError(f"Error: Required option '{k}' not set for grid({arg})!")
        if S("spacing") == S("grid_points"):
            raise InputError("Error: Either spacing or grid_points must be set" if not S("spacing")
                             else "Error: Both spacing and grid_points set!")
    elif arg == "file":
        if not S("file"):
            raise InputError("Error: Required option 'file' not set for grid(file)!")
        return
    elif arg == "bond":
        if not S("distance"):
            raise InputError("Error: Required option 'distance' not set for grid(bond)!")
        if S("origin"):
            raise InputError("Error: Keyword 'origin' incompatible with 'bond' grids")
        if not (S("fixpoint") or S("fixcoord")):
            raise InputError("Error: Either fixpoint or fixcoord must be specified")
        if not (S("width") and S("height")):
            raise InputError("Error: missing or incomplete specification for width and height")
        if S("bond"):
            if S("coord1") or S("coord2"):
                raise InputError("Error: Both bond and coord(s) have been specified")
        elif not (S("coord1") and S("coord2")):
            raise InputError("Error: Invalid bond specification")
    else:
        raise InputError(f"Error: unknown grid type '{arg}'")
    if inp.get("Grid.type") == "even":
        if S("gauss_order"):
            raise InputError("Error: 'gauss_order' incompatible with type=even grids")
        if S("spacing") and S("grid_points"):
            raise InputError("Error: both spacing and grid_points cannot be specified")
        if not S("spacing") and not S("grid_points"):
            raise InputError("Error: either spacing or grid_points must be specified")
