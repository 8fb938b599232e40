"""Static checks of the built library (no GPU): it holds sm_100a code only, the kernels the launchers pick by default keep
the register budget their CTAs-per-SM count needs (the occupancy the measured numbers in profiles/README.md rest on), and the
headline kernel really is the TMA bulk-copy / mbarrier one (SASS mnemonics of /opt/skills/guides/B200_PROFILING.md)."""
import re
import shutil
import subprocess

import pytest

CUOBJDUMP = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"


@pytest.fixture(scope="module")
def usage(crd):
    """{demangled-ish kernel signature: (mangled name, registers, stack bytes)}"""
    from crdmodel_b200 import build as B
    out = subprocess.run([CUOBJDUMP, "-res-usage", B.LIB], capture_output=True, text=True, check=True).stdout
    res = {}
    for m in re.finditer(r"Function (\S+):\n\s*REG:(\d+) STACK:(\d+)", out):
        mangled = m.group(1)
        dem = subprocess.run(["c++filt", mangled], capture_output=True, text=True).stdout.strip()
        dem = dem.replace("(anonymous namespace)::", "").replace("void ", "").split("(")[0]
        res[dem] = (mangled, int(m.group(2)), int(m.group(3)))
    return res


def test_library_is_sm_100a_only(crd):
    from crdmodel_b200 import build as B
    out = subprocess.run([CUOBJDUMP, "-lelf", B.LIB], capture_output=True, text=True, check=True).stdout
    archs = set(re.findall(r"\.(sm_\w+)\.cubin", out))
    assert archs == {"sm_100a"}, archs
    ptx = subprocess.run([CUOBJDUMP, "-lptx", B.LIB], capture_output=True, text=True).stdout
    assert set(re.findall(r"\.(sm_\w+)\.ptx", ptx)) <= {"sm_100a"}


# model ids: 0 FHN torus, 1 Goldbeter torus, 2 FHN flat, 3 Goldbeter flat (include/crd_b200.h)
@pytest.mark.parametrize("model", [0, 1, 2, 3])
@pytest.mark.parametrize("exact", ["true", "false"])
def test_register_budgets_of_the_default_kernels(usage, model, exact):
    def regs(sig):
        assert sig in usage, "kernel not in the library: " + sig
        return usage[sig][1]
    # headline: tiled kernel, 256 threads, 3 CTAs/SM -> 65536 / 768 = 85 -> 80
    assert regs("rhs_tile_kernel<%d, %s, 256, 16, 3, false, false, 3>" % (model, exact)) <= 80
    # streaming kernel, 288 threads: 3 CTAs/SM -> 72 registers (2- and 3-vector stages, last stage + finish)
    for nv, rb, fin in ((2, 2, 0), (3, 2, 0), (5, 1, 2)):
        assert regs("rhs_stream_kernel<%d, %s, %d, false, %d, %d, 3>" % (model, exact, nv, rb, fin)) <= 72
    # 2 CTAs/SM -> 112
    assert regs("rhs_stream_kernel<%d, %s, 5, false, 1, 0, 2>" % (model, exact)) <= 112


def test_resident_loop_fits_one_cta_of_512_threads_per_sm(usage):
    ks = {k: v for k, v in usage.items() if k.startswith("erk_resident_kernel<")}
    assert ks
    for k, (_, reg, _) in ks.items():
        assert reg <= 128, (k, reg)


def test_headline_kernel_uses_tma_bulk_copies_and_mbarriers(usage, crd):
    from crdmodel_b200 import build as B
    mangled = usage["rhs_tile_kernel<0, true, 256, 16, 3, false, false, 3>"][0]
    sass = subprocess.run([CUOBJDUMP, "-sass", "-fun", mangled, B.LIB], capture_output=True, text=True).stdout
    assert "UBLKCP" in sass, "no TMA bulk copy in the headline kernel"
    assert "SYNCS" in sass, "no mbarrier operations in the headline kernel"
    assert "DFMA" in sass and "LDL" not in sass and "STL" not in sass   # fp64 arithmetic, nothing spilled to local memory
