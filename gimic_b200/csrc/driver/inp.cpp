// gimic.inp reader: the surface syntax of doc/input.rst:4-19 (key=value, key=[a,b,c], Section(arg){...}, '#' comments,
// '|' continuation), the keyword schema and defaults of src/gimic.in:60-112 and the cross-checks of check_top / check_grid
// (src/gimic.in:161-283).  Errors throw InputError with the front end's message instead of sys.exit.
#include <algorithm>
#include <cctype>
#include <cstdlib>
#include <fstream>
#include <sstream>

#include "native_driver.hpp"

namespace gbd {

namespace {

struct KW { const char *sect, *key; Type type; const char *def; };   // def == nullptr: no default
const KW SCHEMA[] = {
    {"", "title", Type::STR, ""}, {"", "debug", Type::INT, "0"}, {"", "calc", Type::STR, nullptr}, {"", "backend", Type::STR, "gimic"},
    {"", "basis", Type::STR, "mol"}, {"", "density", Type::STR, ""}, {"", "mofile", Type::STR, ""}, {"", "mos", Type::INT_ARRAY, "0 0"},
    {"", "xdens", Type::STR, "XDENS"}, {"", "magnet_axis", Type::STR, ""}, {"", "magnet", Type::DBL_ARRAY, "0 0 0"},
    {"", "openshell", Type::BOOL, "false"}, {"", "dryrun", Type::BOOL, "false"},
    {"Advanced", "screening", Type::BOOL, "false"}, {"Advanced", "screening_thrs", Type::DBL, "1.0e-8"},
    {"Advanced", "spherical", Type::BOOL, "true"}, {"Advanced", "GIAO", Type::BOOL, "true"}, {"Advanced", "diamag", Type::BOOL, "true"},
    {"Advanced", "paramag", Type::BOOL, "true"}, {"Advanced", "lip_order", Type::INT, "3"},
    {"Essential", "acid", Type::BOOL, "false"}, {"Essential", "jmod", Type::BOOL, "false"}, {"Essential", "prop", Type::BOOL, "false"},
    {"Grid", "type", Type::STR, "even"}, {"Grid", "file", Type::STR, nullptr}, {"Grid", "origin", Type::DBL_ARRAY, nullptr},
    {"Grid", "ivec", Type::DBL_ARRAY, nullptr}, {"Grid", "jvec", Type::DBL_ARRAY, nullptr}, {"Grid", "lengths", Type::DBL_ARRAY, nullptr},
    {"Grid", "bond", Type::INT_ARRAY, nullptr}, {"Grid", "fixpoint", Type::INT, nullptr}, {"Grid", "coord1", Type::DBL_ARRAY, nullptr},
    {"Grid", "coord2", Type::DBL_ARRAY, nullptr}, {"Grid", "fixcoord", Type::DBL_ARRAY, nullptr}, {"Grid", "distance", Type::DBL, nullptr},
    {"Grid", "rotation", Type::DBL_ARRAY, "0 0 0"}, {"Grid", "rotation_origin", Type::DBL_ARRAY, "0 0 0"},
    {"Grid", "spacing", Type::DBL_ARRAY, nullptr}, {"Grid", "height", Type::DBL_ARRAY, nullptr}, {"Grid", "width", Type::DBL_ARRAY, nullptr},
    {"Grid", "radius", Type::DBL, "-1.0"}, {"Grid", "gridplot", Type::INT, nullptr}, {"Grid", "grid_points", Type::INT_ARRAY, nullptr},
    {"Grid", "gauss_order", Type::INT, "7"},
};

std::string path_of(const std::string &sect, const std::string &key) { return sect.empty() ? key : sect + "." + key; }

bool known_section(const std::string &s) {
    for (const KW &k : SCHEMA) if (s == k.sect) return true;
    return false;
}

std::string strip(const std::string &s) {
    size_t a = 0, b = s.size();
    while (a < b && std::isspace((unsigned char)s[a])) ++a;
    while (b > a && std::isspace((unsigned char)s[b - 1])) --b;
    return s.substr(a, b - a);
}

std::string unquote(std::string v) {
    v = strip(v);
    auto strip_ch = [&](char q) { while (!v.empty() && v.front() == q) v.erase(v.begin()); while (!v.empty() && v.back() == q) v.pop_back(); };
    strip_ch('"'); strip_ch('\'');
    return v;
}

double number(const std::string &tok, const std::string &name) {          // reals accept 1.d-8 (getkw.py)
    std::string t = tok;
    for (char &c : t) if (c == 'd' || c == 'D') c = 'e';
    char *end = nullptr;
    double x = std::strtod(t.c_str(), &end);
    if (t.empty() || end == t.c_str() || *end != '\0') throw InputError("invalid number '" + tok + "' for " + name);
    return x;
}

bool boolean(const std::string &tok, const std::string &name) {           // getkw.py:32,280-281
    std::string lv = tok;
    for (char &c : lv) c = (char)std::tolower((unsigned char)c);
    for (const char *t : {"on", "true", "yes", "1", "t", "y"}) if (lv == t) return true;
    for (const char *f : {"off", "false", "no", "0", "f", "n"}) if (lv == f) return false;
    throw InputError("invalid boolean '" + tok + "' for " + name);
}

void convert(Value &v, const std::vector<std::string> &raw, bool is_list, const std::string &name) {
    const bool array = v.type == Type::INT_ARRAY || v.type == Type::DBL_ARRAY;
    if (!array && is_list && raw.size() != 1) throw InputError(name + " expects a scalar");
    v.none = false;
    v.iv.clear(); v.dv.clear();
    switch (v.type) {
        case Type::STR: v.s = unquote(raw[0]); break;
        case Type::INT: v.i = (long)number(unquote(raw[0]), name); break;
        case Type::DBL: v.x = number(unquote(raw[0]), name); break;
        case Type::BOOL: v.b = boolean(unquote(raw[0]), name); break;
        case Type::INT_ARRAY: for (const auto &t : raw) v.iv.push_back((long)number(unquote(t), name)); break;
        case Type::DBL_ARRAY: for (const auto &t : raw) v.dv.push_back(number(unquote(t), name)); break;
    }
}

std::vector<std::string> split_list(const std::string &body) {           // re.split(r"[,\s]+", ...)
    std::vector<std::string> out;
    std::string cur;
    for (char c : body) {
        if (c == ',' || std::isspace((unsigned char)c)) { if (!cur.empty()) { out.push_back(cur); cur.clear(); } }
        else cur.push_back(c);
    }
    if (!cur.empty()) out.push_back(cur);
    return out;
}

std::string strip_comments(const std::string &text) {
    std::string out;
    out.reserve(text.size());
    size_t pos = 0;
    while (pos <= text.size()) {
        size_t nl = text.find('\n', pos);
        std::string line = text.substr(pos, nl == std::string::npos ? std::string::npos : nl - pos);
        char q = 0;
        std::string buf;
        for (char ch : line) {
            if (q) { buf.push_back(ch); if (ch == q) q = 0; }
            else if (ch == '"' || ch == '\'') { q = ch; buf.push_back(ch); }
            else if (ch == '#') break;
            else buf.push_back(ch);
        }
        out += buf;
        if (nl == std::string::npos) break;
        out.push_back('\n');
        pos = nl + 1;
    }
    // '|' line continuation: "|" + optional blanks + newline -> one blank
    std::string res;
    for (size_t i = 0; i < out.size(); ++i) {
        if (out[i] == '|') {
            size_t j = i + 1;
            while (j < out.size() && out[j] != '\n' && std::isspace((unsigned char)out[j])) ++j;
            if (j < out.size() && out[j] == '\n') {
                // \s* is greedy and also swallows blank lines that follow; the last newline it can reach ends the match
                size_t k = j;
                size_t last_nl = j;
                while (k < out.size() && std::isspace((unsigned char)out[k])) { if (out[k] == '\n') last_nl = k; ++k; }
                res.push_back(' ');
                i = last_nl;
                continue;
            }
        }
        res.push_back(out[i]);
    }
    return res;
}

bool ident_start(char c) { return std::isalpha((unsigned char)c) || c == '_'; }
bool ident_char(char c) { return std::isalnum((unsigned char)c) || c == '_'; }

void check(Input &inp);

}  // namespace

Input::Input() {
    for (const KW &k : SCHEMA) {
        Value v;
        v.type = k.type;
        if (k.def) {
            if (k.type == Type::STR) { v.none = false; v.s = k.def; }
            else convert(v, split_list(k.def), true, k.key);
        }
        values_[path_of(k.sect, k.key)] = v;
    }
}

const Value &Input::get(const std::string &path) const {
    auto it = values_.find(path);
    if (it == values_.end()) throw InputError("unknown keyword '" + path + "'");
    return it->second;
}

Vec3 Input::vec3(const std::string &path) const {
    const Value &v = get(path);
    if (v.dv.size() < 3) throw InputError("'" + path + "' needs three values");
    return Vec3{{v.dv[0], v.dv[1], v.dv[2]}};
}

void Input::assign(const std::string &sect, const std::string &key, const std::vector<std::string> &raw, bool is_list) {
    auto it = values_.find(path_of(sect, key));
    if (!known_section(sect) || it == values_.end()) {
        if (sect.rfind("Gimlet", 0) == 0) return;
        throw InputError("unknown keyword '" + key + "' in section '" + (sect.empty() ? std::string("top") : sect) + "'");
    }
    if (raw.empty()) throw InputError("empty value for " + key);
    convert(it->second, raw, is_list, key);
    set_.insert(path_of(sect, key));
}

void Input::force_flag(const std::string &path, bool v) { Value &x = values_.at(path); x.none = false; x.b = v; }
void Input::force_str(const std::string &path, const std::string &v) { Value &x = values_.at(path); x.none = false; x.s = v; }

Input parse_text(const std::string &raw_text) {
    Input inp;
    const std::string text = strip_comments(raw_text);
    std::vector<std::string> stack{""};
    size_t pos = 0;
    const size_t n = text.size();
    auto skip_ws = [&](size_t p) { while (p < n && std::isspace((unsigned char)text[p])) ++p; return p; };
    auto fail_near = [&](size_t p) { throw InputError("cannot parse input near: '" + text.substr(p, 40) + "'"); };
    while (true) {
        size_t p = skip_ws(pos);
        if (p >= n) break;
        if (text[p] == '}') {
            if (stack.size() == 1) throw InputError("unbalanced '}'");
            stack.pop_back();
            pos = p + 1;
            continue;
        }
        if (!ident_start(text[p])) fail_near(pos);
        size_t e = p;
        while (e < n && ident_char(text[e])) ++e;
        const std::string name = text.substr(p, e - p);
        // Section [ (arg) ] {
        size_t q = skip_ws(e);
        bool has_arg = false;
        std::string arg;
        size_t after = q;
        if (q < n && text[q] == '(') {
            size_t close = text.find(')', q);
            if (close != std::string::npos) { has_arg = true; arg = strip(text.substr(q + 1, close - q - 1)); after = skip_ws(close + 1); }
        }
        if (after < n && text[after] == '{') {
            if (!known_section(name) && name != "Gimlet") throw InputError("unknown section '" + name + "'");
            if (name == "Grid") {
                inp.grid_present = true;
                if (has_arg && !arg.empty()) inp.grid_arg = unquote(arg);
            }
            stack.push_back(name);
            pos = after + 1;
            continue;
        }
        // key = value
        if (q >= n || text[q] != '=') fail_near(pos);
        size_t v = skip_ws(q + 1);
        if (v >= n) fail_near(pos);
        std::vector<std::string> raw;
        bool is_list = false;
        if (text[v] == '[') {
            size_t close = text.find(']', v);
            if (close == std::string::npos) fail_near(pos);
            raw = split_list(text.substr(v + 1, close - v - 1));
            is_list = true;
            pos = close + 1;
        } else if (text[v] == '"' || text[v] == '\'') {
            size_t close = text.find(text[v], v + 1);
            if (close == std::string::npos) fail_near(pos);
            raw.push_back(text.substr(v, close - v + 1));
            pos = close + 1;
        } else {
            size_t w = v;
            while (w < n && !std::isspace((unsigned char)text[w]) && text[w] != '{' && text[w] != '}') ++w;
            if (w == v) fail_near(pos);
            raw.push_back(text.substr(v, w - v));
            pos = w;
        }
        inp.assign(stack.back(), name, raw, is_list);
    }
    if (stack.size() != 1) throw InputError("unbalanced '{'");
    check(inp);
    return inp;
}

Input parse_file(const std::string &path) {
    std::ifstream f(path);
    if (!f) throw InputError("cannot open input file " + path);
    std::stringstream ss;
    ss << f.rdbuf();
    return parse_text(ss.str());
}

namespace {

// check_top / check_grid, src/gimic.in:161-283
void check(Input &inp) {
    const Value &calc = inp.get("calc");
    const std::string c = calc.none ? "None" : calc.s;
    if (c != "cdens" && c != "integral" && c != "divj" && c != "edens") throw InputError("Error: unknown option calc = " + c);
    const bool om = inp.is_set("magnet"), pm = inp.is_set("magnet_axis");
    if (om && pm) throw InputError("Error: Both magnet vector and axis set simultaneously!");
    if (!om && !pm) throw InputError("Error: Direction of magnetic field must be set!");
    if (pm && (inp.str("magnet_axis") == "T" || inp.str("magnet_axis") == "X")) inp.force_str("magnet_axis", "X");
    if (!inp.grid_present) throw InputError("Error: no Grid section");
    const std::string arg = inp.grid_arg;
    auto S = [&](const char *k) { return inp.is_set(std::string("Grid.") + k); };
    if (arg == "std" || arg == "base") {
        for (const char *k : {"origin", "ivec", "jvec", "lengths"})
            if (!S(k)) throw InputError(std::string("Error: Required option '") + k + "' not set for grid(" + arg + ")!");
        if (S("spacing") == S("grid_points"))
            throw InputError(!S("spacing") ? "Error: Either spacing or grid_points must be set" : "Error: Both spacing and grid_points set!");
    } else if (arg == "file") {
        if (!S("file")) throw InputError("Error: Required option 'file' not set for grid(file)!");
        return;
    } else if (arg == "bond") {
        if (!S("distance")) throw InputError("Error: Required option 'distance' not set for grid(bond)!");
        if (S("origin")) throw InputError("Error: Keyword 'origin' incompatible with 'bond' grids");
        if (!(S("fixpoint") || S("fixcoord"))) throw InputError("Error: Either fixpoint or fixcoord must be specified");
        if (!(S("width") && S("height"))) throw InputError("Error: missing or incomplete specification for width and height");
        if (S("bond")) {
            if (S("coord1") || S("coord2")) throw InputError("Error: Both bond and coord(s) have been specified");
        } else if (!(S("coord1") && S("coord2"))) {
            throw InputError("Error: Invalid bond specification");
        }
    } else {
        throw InputError("Error: unknown grid type '" + arg + "'");
    }
    if (inp.str("Grid.type") == "even") {
        if (S("gauss_order")) throw InputError("Error: 'gauss_order' incompatible with type=even grids");
        if (S("spacing") && S("grid_points")) throw InputError("Error: both spacing and grid_points cannot be specified");
        if (!S("spacing") && !S("grid_points")) throw InputError("Error: either spacing or grid_points must be specified");
    }
}

}  // namespace

}  // namespace gbd
