#include "host_basis.hpp"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <charconv>
#include <fstream>
#include <thread>
#include <sstream>

namespace gb {

namespace {

// ---- tiny Fortran-list-directed tokenizer ------------------------------------------------------
struct Records {
    std::vector<std::string> lines;
    size_t cur = 0;
    bool open(const std::string &path) {
        std::ifstream f(path);
        if (!f) return false;
        std::string s;
        while (std::getline(f, s)) lines.push_back(s);
        return true;
    }
    bool line(std::string &s) {
        if (cur >= lines.size()) return false;
        s = lines[cur++];
        return true;
    }
};

void tokens_of(const std::string &s, std::vector<std::string> &out) {
    size_t i = 0, n = s.size();
    while (i < n) {
        while (i < n && (std::isspace((unsigned char)s[i]) || s[i] == ',')) ++i;
        size_t j = i;
        while (j < n && !std::isspace((unsigned char)s[j]) && s[j] != ',') ++j;
        if (j > i) out.push_back(s.substr(i, j - i));
        i = j;
    }
}

double to_double(const std::string &t) {
    char buf[64];
    size_t n = std::min(t.size(), sizeof(buf) - 1);
    for (size_t i = 0; i < n; ++i) buf[i] = (t[i] == 'D' || t[i] == 'd') ? 'e' : t[i];
    buf[n] = 0;
    return std::strtod(buf, nullptr);
}

// A list-directed READ of `want` items: begins on a fresh record and keeps consuming records until
// satisfied; whatever is left on the last record is dropped.
bool read_items(Records &r, size_t want, std::vector<std::string> &got) {
    got.clear();
    std::string s;
    while (got.size() < want) {
        if (!r.line(s)) return false;
        tokens_of(s, got);
    }
    got.resize(want);
    return true;
}

// component tables, generated rather than tabulated:
// standard order = lexicographically descending (lx, ly) ; Turbomole order is tabulated.
const int TM_D[6][3] = {{2,0,0},{0,2,0},{0,0,2},{1,1,0},{1,0,1},{0,1,1}};
const int TM_F[10][3] = {{3,0,0},{0,3,0},{0,0,3},{2,1,0},{2,0,1},{1,2,0},{0,2,1},{1,0,2},{0,1,2},{1,1,1}};
const int TM_G[15][3] = {{4,0,0},{0,4,0},{0,0,4},{3,1,0},{3,0,1},{1,3,0},{0,3,1},{1,0,3},{0,1,3},{2,2,0},
                         {2,0,2},{0,2,2},{2,1,1},{1,2,1},{1,1,2}};
const int TM_H[21][3] = {{5,0,0},{0,5,0},{0,0,5},{4,1,0},{4,0,1},{1,4,0},{0,4,1},{1,0,4},{0,1,4},{3,2,0},
                         {3,0,2},{2,3,0},{0,3,2},{2,0,3},{0,2,3},{3,1,1},{1,3,1},{1,1,3},{2,2,1},{2,1,2},{1,2,2}};

void append_shell(HostBasis &b, int atom, int l, const std::vector<double> &xp, const std::vector<double> &c) {
    Shell s;
    s.atom = atom; s.l = l; s.nprim = (int)xp.size(); s.prim_off = (int)b.alpha.size();
    s.ncomp = (l + 1) * (l + 2) / 2; s.user_off = b.nbf; s.thr = 1e10;
    s.nsph = 2 * l + 1; s.sph_off = b.nbf_sph;
    b.nbf_sph += s.nsph;
    b.alpha.insert(b.alpha.end(), xp.begin(), xp.end());
    b.cc.insert(b.cc.end(), c.begin(), c.end());
    b.shells.push_back(s);
    b.nbf += s.ncomp;
    b.ngto += s.nprim * s.ncomp;
}

}  // namespace

void component_exponents(int l, bool turbomole, int c, int lmn[3]) {
    if (turbomole && l >= 2) {
        const int(*t)[3] = l == 2 ? TM_D : l == 3 ? TM_F : l == 4 ? TM_G : TM_H;
        lmn[0] = t[c][0]; lmn[1] = t[c][1]; lmn[2] = t[c][2];
        return;
    }
    // standard GIMIC order (gtodefs.f90:86-106): lx descending, then ly descending
    int k = 0;
    for (int lx = l; lx >= 0; --lx)
        for (int ly = l - lx; ly >= 0; --ly, ++k)
            if (k == c) { lmn[0] = lx; lmn[1] = ly; lmn[2] = l - lx - ly; return; }
}

bool parse_mol(const std::string &path, HostBasis &b, std::string &err) {
    Records r;
    if (!r.open(path)) { err = "read_intgrl(): open failed: " + path; return false; }
    std::string s;
    std::vector<std::string> t;
    if (!r.line(s) || s.compare(0, 6, "INTGRL") != 0) { err = "this doesn't look like an 'INTGRL' file: " + path; return false; }
    if (!r.line(s)) { err = "MOL file truncated"; return false; }
    b.turbomole = s.compare(0, 9, "TURBOMOLE") == 0;
    r.line(s);
    if (!read_items(r, 1, t)) { err = "MOL file truncated (natoms)"; return false; }
    b.natoms = std::atoi(t[0].c_str());
    if (b.natoms <= 0 || b.natoms > 10000000) { err = "MOL: bad atom count"; return false; }
    r.line(s);
    b.atom_shell_off.assign(1, 0);
    b.atom_func_off.assign(1, 0);
    for (int a = 0; a < b.natoms; ++a) {
        if (!r.line(s)) { err = "MOL file truncated (atom header)"; return false; }
        t.clear(); tokens_of(s, t);
        if (t.size() < 3) { err = "MOL: malformed atom header"; return false; }
        int nsh = std::atoi(t[2].c_str());
        if (nsh - 1 > MAX_L) { err = "Largest allowed l-quantum number in basis exceeded"; return false; }
        if (nsh < 0) { err = "MOL: malformed atom header (shell count)"; return false; }
        if ((int)t.size() < 3 + nsh) { err = "MOL: malformed atom header (block counts)"; return false; }
        b.charge.push_back(to_double(t[0]));
        std::vector<int> nblk(nsh);
        for (int i = 0; i < nsh; ++i) {
            nblk[i] = std::atoi(t[3 + i].c_str());
            if (nblk[i] < 0 || nblk[i] > MAX_SHELLS_PER_ATOM) { err = "MOL: malformed atom header (block counts)"; return false; }   // posvec(99), basis.f90:118-136
        }
        if (!r.line(s)) { err = "MOL file truncated (atom position)"; return false; }
        if (s.size() < 4) s.resize(4, ' ');
        b.symbol.push_back(s.substr(0, 2));
        t.clear(); tokens_of(s.substr(4), t);
        if (t.size() < 3) { err = "MOL: malformed atom position"; return false; }
        for (int k = 0; k < 3; ++k) b.xyz.push_back(to_double(t[k]));
        for (int l = 0; l < nsh; ++l) {
            for (int blk = 0; blk < nblk[l]; ++blk) {
                if (!read_items(r, 2, t)) { err = "MOL file truncated (block header)"; return false; }
                int npf = std::atoi(t[0].c_str()), ncf = std::atoi(t[1].c_str());
                if (npf <= 0 || ncf <= 0 || npf > 10000 || ncf > 10000) { err = "MOL: bad contraction block"; return false; }
                std::vector<double> xp(npf);
                std::vector<std::vector<double>> co(ncf, std::vector<double>(npf));
                for (int p = 0; p < npf; ++p) {
                    if (!read_items(r, 1 + (size_t)ncf, t)) { err = "MOL file truncated (primitives)"; return false; }
                    xp[p] = to_double(t[0]);
                    for (int c = 0; c < ncf; ++c) co[c][p] = to_double(t[1 + c]);
                }
                // a generally contracted block becomes ncf segmented contractions (intgrl.f90:172-216)
                for (int c = 0; c < ncf; ++c) append_shell(b, a, l, xp, co[c]);
            }
        }
        b.atom_shell_off.push_back((int)b.shells.size());
        b.atom_func_off.push_back(b.nbf);
        if (b.atom_shell_off[a + 1] - b.atom_shell_off[a] > MAX_SHELLS_PER_ATOM) { err = "more than 99 contractions on one atom"; return false; }
    }
    b.nprim_total = (int)b.alpha.size();
    return true;
}

bool basis_from_arrays(int natoms, const double *coords, const int *nctr_per_atom, const int *ctr_l,
                       const int *ctr_npf, const double *xp, const double *cc, int turbomole_order,
                       HostBasis &b, std::string &err) {
    if (natoms <= 0) { err = "natoms <= 0"; return false; }
    b.turbomole = turbomole_order != 0;
    b.natoms = natoms;
    b.xyz.assign(coords, coords + 3 * (size_t)natoms);
    b.charge.assign(natoms, 0.0);
    b.symbol.assign(natoms, "X ");
    b.atom_shell_off.assign(1, 0);
    b.atom_func_off.assign(1, 0);
    size_t ic = 0, ip = 0;
    for (int a = 0; a < natoms; ++a) {
        if (nctr_per_atom[a] > MAX_SHELLS_PER_ATOM) { err = "more than 99 contractions on one atom"; return false; }
        for (int j = 0; j < nctr_per_atom[a]; ++j, ++ic) {
            int l = ctr_l[ic], npf = ctr_npf[ic];
            if (l < 0 || l > MAX_L || npf <= 0) { err = "bad shell description"; return false; }
            append_shell(b, a, l, std::vector<double>(xp + ip, xp + ip + npf), std::vector<double>(cc + ip, cc + ip + npf));
            ip += npf;
        }
        b.atom_shell_off.push_back((int)b.shells.size());
        b.atom_func_off.push_back(b.nbf);
    }
    b.nprim_total = (int)b.alpha.size();
    return true;
}

void finalize_basis(HostBasis &b, bool use_screening, double screening_thrs) {
    const double pi = 3.141592653589793;  // PII, globals.f90:41
    b.ncc.assign(b.alpha.size(), 0.0);
    for (Shell &s : b.shells) {
        const double *a = &b.alpha[s.prim_off], *c = &b.cc[s.prim_off];
        // contraction self-overlap of the normalised primitives (basis.f90:171-184)
        double j = 1.0 * (s.l + 1), n = 0.0;
        for (int p = 0; p < s.nprim; ++p)
            for (int q = 0; q <= p; ++q) {
                double t = 2.0 * std::sqrt(a[p] * a[q]) / (a[p] + a[q]);
                t = c[p] * c[q] * std::pow(t, j + 0.5);
                n += (p == q) ? t : 2.0 * t;
            }
        n = 1.0 / std::sqrt(n);
        for (int p = 0; p < s.nprim; ++p)
            b.ncc[s.prim_off + p] = c[p] * n * std::pow(4.0 * a[p], 0.5 * j + 0.25) * std::pow(0.5 / pi, 0.75);
        // screening radius: first multiple of 0.25 bohr where d^l exp(-a_min d^2) <= thrs (basis.f90:90-112)
        if (!use_screening || screening_thrs <= 0.0) { s.thr = 1e10; continue; }
        double amin = a[0];
        for (int p = 1; p < s.nprim; ++p) amin = std::min(amin, a[p]);
        double d = 0.0, x = 1e15;
        while (x > screening_thrs) {
            d += 0.25;
            double dl = 1.0;
            for (int k = 0; k < s.l; ++k) dl *= d;
            x = dl * std::exp(-amin * d * d);
        }
        s.thr = d;
    }
}

// XDENS text: list-directed reals, one per line (dens.f90:129-135).  A 10^4-function molecule has 4e8 of them (3+ GB), so
// the buffer is cut at line boundaries into one piece per host thread; each piece is counted, then parsed into its slot.
// Binary cache (this project's extension, written by write_xdens_binary): "GB2XDENS" | int64 nbf | int64 nmat | doubles.
namespace {
const char XD_MAGIC[8] = {'G', 'B', '2', 'X', 'D', 'E', 'N', 'S'};

inline bool is_sep(char ch) { return ch == ' ' || ch == '\n' || ch == '\r' || ch == '\t' || ch == ','; }

size_t count_tokens(const char *p, const char *e) {
    size_t n = 0;
    while (p < e) {
        while (p < e && is_sep(*p)) ++p;
        if (p < e) ++n;
        while (p < e && !is_sep(*p)) ++p;
    }
    return n;
}

// parses at most `limit` tokens of [p, e) into out; returns the number parsed, or (size_t)-1 on a malformed token
size_t parse_tokens(const char *p, const char *e, double *out, size_t limit) {
    size_t n = 0;
    char tok[80];
    while (p < e && n < limit) {
        while (p < e && is_sep(*p)) ++p;
        if (p >= e) break;
        const char *q = p;
        while (q < e && !is_sep(*q)) ++q;
        size_t len = std::min<size_t>((size_t)(q - p), sizeof(tok) - 1);
        const char *b = p;
        if (*b == '+') { ++b; --len; }                       // from_chars takes no leading '+'
        bool fortran_d = false;
        for (size_t i = 0; i < len; ++i) { char ch = b[i]; if (ch == 'D' || ch == 'd') { ch = 'e'; fortran_d = true; } tok[i] = ch; }
        const char *tb = fortran_d ? tok : b;
        double v;
        auto res = std::from_chars(tb, tb + len, v);
        if (res.ec != std::errc() || res.ptr != tb + len) {
            tok[len] = 0;
            if (!fortran_d) std::memcpy(tok, b, len);
            char *endp = nullptr;                               // forms from_chars rejects ("1.-5", ".5+3"): fall back to strtod
            v = std::strtod(tok, &endp);
            if (endp == tok) return (size_t)-1;
        }
        out[n++] = v;
        p = q;
    }
    return n;
}
}  // namespace

bool read_xdens(const std::string &path, int nbf, int nmat, std::vector<double> &out, std::string &err) {
    FILE *f = std::fopen(path.c_str(), "rb");
    if (!f) { err = "Density file not found: " + path; return false; }
    std::fseek(f, 0, SEEK_END);
    long sz = std::ftell(f);
    std::fseek(f, 0, SEEK_SET);
    const size_t want = (size_t)nmat * nbf * nbf;
    char head[24] = {0};
    size_t hgot = std::fread(head, 1, sizeof head, f);
    if (hgot == sizeof head && std::memcmp(head, XD_MAGIC, 8) == 0) {
        long long hn, hm;
        std::memcpy(&hn, head + 8, 8); std::memcpy(&hm, head + 16, 8);
        if (hn != nbf || hm != nmat) { std::fclose(f); err = "binary XDENS cache is for nbf=" + std::to_string(hn) + ", " + std::to_string(hm) + " matrices; expected nbf=" + std::to_string(nbf) + ", " + std::to_string(nmat); return false; }
        out.resize(want);
        size_t got = std::fread(out.data(), sizeof(double), want, f);
        std::fclose(f);
        if (got != want) { err = "binary XDENS cache truncated: " + path; return false; }
        return true;
    }
    std::fseek(f, 0, SEEK_SET);
    std::vector<char> buf((size_t)sz + 1);
    size_t got = std::fread(buf.data(), 1, (size_t)sz, f);
    std::fclose(f);
    buf[got] = '\n';
    const char *base = buf.data(), *end = base + got;
    unsigned nthr = std::max(1u, std::min(std::thread::hardware_concurrency(), 64u));
    if (got < (1u << 20)) nthr = 1;
    std::vector<const char *> cut(nthr + 1);
    cut[0] = base; cut[nthr] = end;
    for (unsigned t = 1; t < nthr; ++t) {
        const char *p = base + got / nthr * t;
        while (p < end && !is_sep(*p)) ++p;   // never inside a token
        cut[t] = std::max(p, cut[t - 1]);
    }
    std::vector<size_t> cnt(nthr, 0), off(nthr + 1, 0);
    {
        std::vector<std::thread> th;
        for (unsigned t = 1; t < nthr; ++t) th.emplace_back([&, t] { cnt[t] = count_tokens(cut[t], cut[t + 1]); });
        cnt[0] = count_tokens(cut[0], cut[1]);
        for (auto &x : th) x.join();
    }
    for (unsigned t = 0; t < nthr; ++t) off[t + 1] = off[t] + cnt[t];
    if (off[nthr] < want) { err = "XDENS too short: expected " + std::to_string(want) + " values, found " + std::to_string(off[nthr]); return false; }
    // The reference reads one value per record (read(iunit,*) array(i), dens.f90:129-135) and stops after nmat*nbf*nbf of them; a
    // file with MORE numbers is one written for another basis size or an 8-matrix open-shell file opened as closed shell (or the
    // other way round) -- the most common user error.  Refuse it instead of silently loading the first `want` tokens.
    if (off[nthr] > want) {
        err = "XDENS holds " + std::to_string(off[nthr]) + " values, expected " + std::to_string(nmat) + " matrices of " + std::to_string(nbf) + " x " +
              std::to_string(nbf) + " = " + std::to_string(want) + (off[nthr] == 2 * want && nmat == 4 ? " (an open-shell file with 8 matrices? set uhf)" : " (density for another basis set?)");
        return false;
    }
    out.resize(want);
    std::vector<int> bad(nthr, 0);
    auto work = [&](unsigned t) {
        if (off[t] >= want) return;
        size_t lim = std::min(cnt[t], want - off[t]);
        if (parse_tokens(cut[t], cut[t + 1], out.data() + off[t], lim) != lim) bad[t] = 1;
    };
    {
        std::vector<std::thread> th;
        for (unsigned t = 1; t < nthr; ++t) th.emplace_back(work, t);
        work(0);
        for (auto &x : th) x.join();
    }
    for (unsigned t = 0; t < nthr; ++t) if (bad[t]) { err = "XDENS: malformed number in " + path; return false; }
    return true;
}

bool write_xdens_binary(const std::string &path, int nbf, int nmat, const double *vals, std::string &err) {
    FILE *f = std::fopen(path.c_str(), "wb");
    if (!f) { err = "cannot write " + path; return false; }
    long long hn = nbf, hm = nmat;
    size_t want = (size_t)nmat * nbf * nbf;
    bool ok = std::fwrite(XD_MAGIC, 1, 8, f) == 8 && std::fwrite(&hn, 8, 1, f) == 1 && std::fwrite(&hm, 8, 1, f) == 1 &&
              std::fwrite(vals, sizeof(double), want, f) == want;
    ok = (std::fclose(f) == 0) && ok;
    if (!ok) err = "short write to " + path;
    return ok;
}

void turbomole_permutation(const HostBasis &b, std::vector<int> &sv) {
    sv.clear();
    sv.reserve(b.nbf);
    for (int l = 0; l <= MAX_L; ++l)
        for (const Shell &s : b.shells)
            if (s.l == l)
                for (int c = 0; c < s.ncomp; ++c) sv.push_back(s.user_off + c);
}

void turbomole_permutation_sph(const HostBasis &b, std::vector<int> &sv) {
    sv.clear();
    sv.reserve(b.nbf_sph);
    for (int l = 0; l <= MAX_L; ++l)
        for (const Shell &s : b.shells)
            if (s.l == l)
                for (int c = 0; c < s.nsph; ++c) sv.push_back(s.sph_off + c);
}

// ---- spherical components (cao2sao.f90) ---------------------------------------------------------
namespace {
long long ifact(int n) { long long m = 1; for (int i = 2; i <= n; ++i) m *= i; return m; }
long long ibinom(int a, int b) { return ifact(a) / (ifact(b) * ifact(a - b)); }
long long igcd(long long a, long long b) { a = a < 0 ? -a : a; b = b < 0 ? -b : b; while (b) { long long t = a % b; a = b; b = t; } return a; }
}  // namespace

// Real solid harmonic S_lm as a polynomial (Helgaker et al., Molecular Electronic-Structure Theory, eq. 9.1.9):
//   S_lm ~ sum_{t,u,v} (-1)^(t+v) 4^-t C(l,t) C(l-t,|m|+t) C(t,u) C(|m|,2v') x^(2t+|m|-2(u+v')) y^(2(u+v')) z^(l-2t-|m|),
// v' = v (m >= 0) or v + 1/2 (m < 0).  Several (u,v) pairs hit the same monomial (same u+v); the reference STORES instead
// of accumulating (cao2sao.f90:188), so the coefficient of a monomial is the term of the last pair in its loop order,
// u = min(t, u+v).  Kept bug-compatible (affects only |m| >= 2 rows of g and h shells); the common factor N_lm drops out
// in the reference's integer renormalisation (cao2sao.f90:201-231), which for l <= 5 yields the row divided by its gcd.
void c2s_rows(int l, bool turbomole, std::vector<double> &po) {
    const int nc = (l + 1) * (l + 2) / 2;
    po.assign((size_t)(2 * l + 1) * nc, 0.0);
    int cidx[MAX_L + 1][MAX_L + 1];   // (lx, ly) -> component
    for (int c = 0; c < nc; ++c) { int e[3]; component_exponents(l, turbomole, c, e); cidx[e[0]][e[1]] = c; }
    const long long scale = 1LL << (2 * (l / 2));   // 4^tmax: makes every coefficient an integer
    for (int m = -l; m <= l; ++m) {
        const int am = m < 0 ? -m : m, odd = m < 0 ? 1 : 0;
        const int vmax = (am - odd) / 2;
        std::vector<long long> row(nc, 0);
        for (int t = 0; 2 * t <= l - am; ++t)
            for (int s = 0; s <= t + vmax; ++s) {   // s = u + v
                const int u = s < t ? s : t, v = s - u;
                long long q = ibinom(l, t) * ibinom(l - t, am + t) * ibinom(t, u) * ibinom(am, 2 * v + odd) * (scale >> (2 * t));
                if ((t + v) & 1) q = -q;
                const int ly = 2 * s + odd, lx = 2 * t + am - ly;
                row[cidx[lx][ly]] = q;
            }
        long long g = 0;
        for (long long q : row) g = igcd(g, q);
        for (int c = 0; c < nc; ++c) po[(size_t)(m + l) * nc + c] = g ? (double)(row[c] / g) : 0.0;
    }
}

void density_sph_to_cart(const HostBasis &b, const double *dsph, double *dcart) {
    const size_t ns = (size_t)b.nbf_sph, nc = (size_t)b.nbf;
    std::vector<double> po[MAX_L + 1];
    for (int l = 0; l <= MAX_L; ++l) c2s_rows(l, b.turbomole, po[l]);
    // half = dsph . po  (ns x nc), then dcart = po^T . half; po is block diagonal per shell
    std::vector<double> half(ns * nc, 0.0);
    for (const Shell &s : b.shells) {
        const std::vector<double> &p = po[s.l];
        for (int k = 0; k < s.ncomp; ++k) {
            double *out = &half[ns * (size_t)(s.user_off + k)];
            for (int q = 0; q < s.nsph; ++q) {
                const double w = p[(size_t)q * s.ncomp + k];
                if (w == 0.0) continue;
                const double *col = dsph + ns * (size_t)(s.sph_off + q);
                for (size_t a = 0; a < ns; ++a) out[a] += w * col[a];
            }
        }
    }
    for (size_t nu = 0; nu < nc; ++nu) {
        const double *hcol = &half[ns * nu];
        double *out = dcart + nc * nu;
        for (const Shell &s : b.shells) {
            const std::vector<double> &p = po[s.l];
            for (int k = 0; k < s.ncomp; ++k) {
                double acc = 0.0;
                for (int q = 0; q < s.nsph; ++q) acc += p[(size_t)q * s.ncomp + k] * hcol[s.sph_off + q];
                out[s.user_off + k] = acc;
            }
        }
    }
}

// ---- Fortran Ew.d formatting ---------------------------------------------------------------------
namespace {
// One value, right-justified in w characters: 0.ddddddE+ee (gfortran: no 'E' when the exponent needs three digits).
void put_e(double x, int w, int d, char *dst) {
    char tmp[64], body[64];
    int len;
    if (x == 0.0) {
        len = std::snprintf(body, sizeof body, "%s0.%0*dE+00", std::signbit(x) ? "-" : "", d, 0);   // gfortran keeps the sign of a negative zero
    } else if (!std::isfinite(x)) {
        len = std::snprintf(body, sizeof body, "%s", std::isnan(x) ? "NaN" : (x < 0 ? "-Infinity" : "Infinity"));
    } else {
        auto res = std::to_chars(tmp, tmp + sizeof tmp - 1, std::fabs(x), std::chars_format::scientific, d - 1);   // d.ddddde+ee, correctly rounded
        *res.ptr = 0;
        const char *ep = std::strchr(tmp, 'e');
        int e = std::atoi(ep + 1) + 1;
        char *o = body;
        if (x < 0) *o++ = '-';
        *o++ = '0'; *o++ = '.';
        for (const char *q = tmp; q < ep; ++q) if (*q != '.') *o++ = *q;
        const int ae = e < 0 ? -e : e;
        if (ae < 100) o += std::snprintf(o, 8, "E%c%02d", e < 0 ? '-' : '+', ae);
        else o += std::snprintf(o, 8, "%c%03d", e < 0 ? '-' : '+', ae);
        len = (int)(o - body);
    }
    if (len > w) { std::memset(dst, '*', (size_t)w); return; }   // field overflow: asterisks like Fortran
    std::memset(dst, ' ', (size_t)(w - len));
    std::memcpy(dst + (w - len), body, (size_t)len);
}
}  // namespace

namespace {
// One value with the fixed-point descriptor Fw.d, right-justified (asterisks on overflow like Fortran)
void put_f(double x, int w, int d, char *dst) {
    char body[400];
    int len;
    if (!std::isfinite(x)) len = std::snprintf(body, sizeof body, "%s", std::isnan(x) ? "NaN" : (x < 0 ? "-Infinity" : "Infinity"));
    else {
        auto res = std::to_chars(body, body + sizeof body - 1, x, std::chars_format::fixed, d);   // correctly rounded, like printf %.*f
        len = (int)(res.ptr - body);
    }
    if (len > w) { std::memset(dst, '*', (size_t)w); return; }
    std::memset(dst, ' ', (size_t)(w - len));
    std::memcpy(dst + (w - len), body, (size_t)len);
}
}  // namespace

long format_fortran(long n, const double *v, char kind, int w, int d, int per_line, int first_count, const char *prefix, char *out, long cap);

long format_fortran_e(long n, const double *v, int w, int d, int per_line, int first_count, const char *prefix, char *out, long cap) {
    return format_fortran(n, v, 'E', w, d, per_line, first_count, prefix, out, cap);
}

long format_fortran(long n, const double *v, char kind, int w, int d, int per_line, int first_count, const char *prefix, char *out, long cap) {
    if (n <= 0) return 0;
    if (per_line <= 0) per_line = 1;
    const long first = first_count > 0 ? std::min<long>(first_count, n) : std::min<long>(per_line, n);
    const long plen = prefix ? (long)std::strlen(prefix) : 0;
    auto line_of = [&](long l) { return l < first ? 0L : 1 + (l - first) / per_line; };
    const long nlines = line_of(n - 1) + 1;
    const long last_count = nlines == 1 ? n : (n - first) - (nlines - 2) * (long)per_line;
    const bool last_complete = nlines == 1 ? (n == (first_count > 0 ? first_count : per_line)) : last_count == per_line;
    const long total = n * w + nlines * plen + (nlines - 1) + (last_complete ? 1 : 0);
    if (total > cap) return -1;
    unsigned nthr = std::max(1u, std::min(std::thread::hardware_concurrency(), 64u));
    if (n < 4096) nthr = 1;
    auto work = [&](long lo, long hi) {
        for (long l = lo; l < hi; ++l) {
            const long ln = line_of(l);
            char *p = out + l * w + (ln + 1) * plen + ln;
            const bool starts = (l == 0) || (l == first) || (l > first && (l - first) % per_line == 0);
            if (starts && plen) std::memcpy(p - plen, prefix, (size_t)plen);
            if (kind == 'F') put_f(v[l], w, d, p); else put_e(v[l], w, d, p);
            const bool ends = (l + 1 == first) || (l + 1 > first && (l + 1 - first) % per_line == 0);
            if (ends && (l + 1 < n || last_complete)) p[w] = '\n';
        }
    };
    std::vector<std::thread> th;
    const long chunk = (n + nthr - 1) / nthr;
    for (unsigned t = 1; t < nthr; ++t) { long lo = t * chunk, hi = std::min(n, lo + chunk); if (lo < hi) th.emplace_back(work, lo, hi); }
    work(0, std::min(n, chunk));
    for (auto &x : th) x.join();
    return total;
}

// ---- quadrature nodes ---------------------------------------------------------------------------
namespace {
// P_n(x) and derivatives by the three-term recurrence (gaussint.f90:144-222)
void legendre(double x, int n, double &p, double &dp, double &d2p) {
    if (n == 0) { p = 1; dp = 0; d2p = 0; return; }
    double pm = 1, dpm = 0, d2pm = 0;
    p = x; dp = 1; d2p = 0;
    for (int i = 2; i <= n; ++i) {
        double c1 = i, c2 = 2.0 * i - 1.0, c4 = i - 1.0;
        double pn = (c2 * x * p - c4 * pm) / c1;
        double dpn = (c2 * x * dp - c4 * dpm + c2 * p) / c1;
        double d2pn = (c2 * x * d2p - c4 * d2pm + c2 * 2.0 * dp) / c1;
        pm = p; p = pn; dpm = dp; dp = dpn; d2pm = d2p; d2p = d2pn;
    }
}
const double NEWTON_EPS = 3.0e-12;  // gaussint.f90:15
const int NEWTON_MAX = 10;

int unit_nodes(int n, int quadrature, std::vector<double> &x, std::vector<double> &w) {  // on [-1, 1]
    const double pi = 3.141592653589793;
    x.assign(n, 0); w.assign(n, 0);
    if (quadrature == 0) {
        for (int i = 1; i <= (n + 1) / 2; ++i) {
            double z = std::cos(pi * (i - 0.25) / (n + 0.5)), p, dp, d2p;
            int it = 1;
            for (; it <= NEWTON_MAX; ++it) {
                legendre(z, n, p, dp, d2p);
                double z1 = z;
                z = z1 - p / dp;
                if (std::fabs(z - z1) <= NEWTON_EPS) break;
            }
            if (it >= NEWTON_MAX) return -1;
            x[i - 1] = -z; x[n - i] = z;
            w[i - 1] = w[n - i] = 2.0 / ((1.0 - z * z) * dp * dp);
        }
    } else {
        x[0] = -1; x[n - 1] = 1;
        w[0] = w[n - 1] = 2.0 / (double)(n * n - n);
        for (int i = 2; i <= n - 1; ++i) {
            double z = std::cos(pi * (i - 0.25) / (n + 0.5)), p = 0, dp, d2p;
            int it = 1;
            for (; it <= NEWTON_MAX; ++it) {
                legendre(z, n - 1, p, dp, d2p);
                double z1 = z, damp = 0.5;
                z = z1 - dp / d2p;
                while (std::fabs(z) > 1.0) { z = z1 - dp / d2p * damp; damp *= damp; }
                if (std::fabs(z - z1) <= NEWTON_EPS) break;
            }
            if (it >= NEWTON_MAX) return -1;
            x[i - 1] = -z;
            w[i - 1] = 2.0 / ((double)(n * n - n) * p * p);
        }
    }
    return 0;
}
}  // namespace

int gauss_blocks(double a, double b, int npts, int order, int quadrature, double *pts, double *wgts) {
    if (npts == 1) { pts[0] = 0.0; wgts[0] = 1.0; return 0; }  // collapsed axis (gaussint.f90:280-284)
    if (order <= 0 || npts % order != 0) return -2;
    std::vector<double> x, w;
    if (int rc = unit_nodes(order, quadrature, x, w)) return rc;
    int nblock = npts / order;
    double step = (b - a) / nblock, half = 0.5 * step;
    for (int blk = 0; blk < nblock; ++blk)
        for (int k = 0; k < order; ++k) {
            pts[blk * order + k] = (x[k] + 1.0) * half + blk * step;   // note: offset from 0, not from a (as the reference)
            wgts[blk * order + k] = w[k] * half;
        }
    return 0;
}

}  // namespace gb
