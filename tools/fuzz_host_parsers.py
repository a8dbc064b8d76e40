"""Mutation fuzzing of the host-side readers through the (sanitizer-built) gimic-b200 program in dry-run mode; see fuzz_host_parsers.sh."""
import os, random, shutil, subprocess, sys, tempfile

exe, rounds = sys.argv[1], int(sys.argv[2])
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "tests", "golden")
random.seed(int(os.environ.get("FUZZ_SEED", "11")))
base = tempfile.mkdtemp(prefix="gimic_fuzz_")
mols = [open(os.path.join(G, m)).read() for m in ("c4h4_MOL", "open_shell_MOL", "benzene_MOL")]
inps = [open(os.path.join(G, "inputs", f)).read() for f in sorted(os.listdir(os.path.join(G, "inputs")))]


def mutate(s):
    s = list(s)
    for _ in range(random.randint(1, 5)):
        if not s:
            s = ["\n"]
        k = random.randrange(len(s)); op = random.random()
        if op < 0.3:
            del s[k:k + random.randint(1, 60)]
        elif op < 0.6:
            s[k] = random.choice("0123456789.-+eEdD []{}()=,#|\"'\nxX ")
        elif op < 0.85:
            s.insert(k, random.choice(["\n", " 99999999999 ", " -1 ", " 1e400 ", "{", "}", "[", "]", " nan ", "  0  ", " -5 ", "=", "()"]))
        else:
            s = s[:k]
    return "".join(s)


bad, rcs = 0, {}
for i in range(rounds):
    which = i % 3
    open(base + "/MOL", "w").write(mutate(random.choice(mols)) if which == 0 else random.choice(mols))
    open(base + "/gimic.inp", "w").write(mutate(random.choice(inps)) if which == 1 else random.choice(inps))
    grd = "0.1 0.2 0.3\n1.0 1.1 1.2\n4 5 6\n"
    open(base + "/gridfile.grd", "w").write(mutate(grd) if which == 2 else grd)
    try:
        p = subprocess.run([exe, "-y", base + "/gimic.inp"], capture_output=True, text=True, timeout=20)
    except subprocess.TimeoutExpired:
        bad += 1; print("TIMEOUT", i); shutil.copytree(base, f"{base}_timeout_{i}"); continue
    rcs[p.returncode] = rcs.get(p.returncode, 0) + 1
    if p.returncode not in (0, 1) or "ERROR" in p.stderr or "runtime error" in p.stderr:
        bad += 1; print("FINDING", i, p.returncode, p.stderr[:1500]); shutil.copytree(base, f"{base}_finding_{i}")
print(f"{rounds} mutated runs, findings: {bad}, exit codes: {rcs}")
sys.exit(1 if bad else 0)
