"""The measurement / profiling scripts under tools/ only run on a GPU box; here they must at least parse, and every file they name must exist."""
import ast
import os
import re
import subprocess
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TOOLS = os.path.join(ROOT, "tools")


@pytest.mark.parametrize("name", sorted(f for f in os.listdir(TOOLS) if f.endswith(".py")))
def test_python_tools_parse(name):
    ast.parse(open(os.path.join(TOOLS, name)).read(), filename=name)


@pytest.mark.parametrize("name", sorted(f for f in os.listdir(TOOLS) if f.endswith(".sh")))
def test_shell_tools_parse_and_reference_existing_files(name):
    path = os.path.join(TOOLS, name)
    subprocess.check_call(["bash", "-n", path])
    text = open(path).read()
    for rel in set(re.findall(r"\b(tools/[\w./-]+\.(?:py|sh)|tests/[\w./-]+\.(?:py|cpp)|bench\.py)\b", text)):
        assert os.path.exists(os.path.join(ROOT, rel)), f"{name} names {rel}, which does not exist"


def test_bench_and_entry_parse():
    for f in ("bench.py", "__graft_entry__.py"):
        ast.parse(open(os.path.join(ROOT, f)).read(), filename=f)


def test_ncu_launch_list_summary_reads_the_committed_capture():
    """tools/ncu_summary.py launches on profiles/r01_launches_tap.csv (the committed ncu launch list of bench.py): the contraction kernel's
    share of a step is what DESIGN.md quotes"""
    import sys
    out = subprocess.check_output([sys.executable, os.path.join(TOOLS, "ncu_summary.py"), "launches", os.path.join(ROOT, "profiles", "r01_launches_tap.csv")],
                                  text=True)
    first = out.split("\n")[0]
    assert "k_jtensor" in first
    share = float(re.search(r"share=\s*([\d.]+)%", first).group(1))
    assert 90.0 < share < 97.0


@pytest.mark.skipif(subprocess.call("command -v cuobjdump >/dev/null 2>&1", shell=True) != 0, reason="cuobjdump not on PATH")
def test_static_sass_record_of_the_built_library():
    """tools/sass_static.py on the library as built: the contraction kernel is FP64 tensor-core code (DMMA) fed by bulk-TMA copies
    (UBLKCP), mbarriers (SYNCS) and cp.async gathers (LDGSTS), re-partitions registers (USETMAXREG) and does not spill"""
    import sys
    so = os.path.join(ROOT, "gimic_b200", "libgimic_b200.so")
    if not os.path.exists(so):
        import __graft_entry__ as ge
        ge.build()
    out = subprocess.check_output([sys.executable, os.path.join(TOOLS, "sass_static.py")], text=True)
    blocks = {m.group(1): m.group(2) for m in re.finditer(r"^(gb::[^\n]+)\n((?:    .*\n)+)", out, flags=re.M)}
    # the default kernels (epilogue warpgroup: tensor path and J path) and the round-1 mapping kept for A/B runs
    for variant in ("gb::k_jtensor_e<true, false>", "gb::k_jtensor_e<false, false>", "gb::k_jtensor_e<true, true>", "gb::k_jtensor_e<false, true>",
                    "gb::k_jtensor<true, true, 8>", "gb::k_jtensor<false, true, 8>", "gb::k_jtensor<true, false, 8>", "gb::k_jtensor<false, false, 8>"):
        b = blocks[variant]
        for op in ("DMMA", "UBLKCP", "SYNCS", "LDGSTS", "USETMAXREG"):
            assert re.search(rf"\b{op} [1-9]", b), (variant, op)
        assert "0 bytes spill stores, 0 bytes spill loads" in b and not re.search(r"\bSTL [1-9]", b), variant
    assert any(k.startswith("gb::k_basis") for k in blocks) and any(k.startswith("gb::k_fields") for k in blocks)
    assert "gb::k_tile_split" in blocks
