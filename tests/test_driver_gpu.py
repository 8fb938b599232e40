"""The host drivers (bin/<Model>_<surface>) against the reference's own main() run on the CPU
(oracle/_ref: reference sources compiled in place, this repository's RK driver behind the ARKode names):
same banner, same subdomain file, same output-file layout — and, with EXACT arithmetic (the default), the same output FILES byte
for byte for the FHN programs (the GPU run takes the steps of the CPU run and every state has the same bits; device-resident or
host-driven loop alike); the Goldbeter programs within rtol*|y| + atol (libm pow vs the device's single-rounded x^4)."""
import os
import subprocess

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

FHN_INI = """[Parameters]
diffusion = 0.12
beta = 1.25
surfaceWidth = 20
surfaceLength = 80
waveLength = 0.1
waveWidth = 0.5
waveInside = {inside}
outputTimestep = 4
tBoundary = 1.0
tFinal = 2
thetaMesh = 24
betaMin = 0.7
betaMax = 1.7

[System]
includeAllVars = 1
varyBeta = {vb}
"""

GB_INI = """[Parameters]
diffusion = 0.12
beta = 0.4
surfaceWidth = 20
surfaceLength = 80
waveLength = 0.2
waveWidth = 0.5
waveInside = 1
outputTimestep = 3
tBoundary = 0
tFinal = 0.15
xMesh = 20
betaMin = 0
betaMax = 1
Zs = 0.392
Ys = 1.6469

[System]
includeAllVars = 1
varyBeta = 0
justDiffusion = 0
icType = 0
"""


def run_driver(exe, ini_text, d):
    os.makedirs(d, exist_ok=True)
    ini = os.path.join(d, "args.ini")
    open(ini, "w").write(ini_text)
    r = subprocess.run([os.path.join(ROOT, "bin", exe), ini], cwd=d, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    return r.stdout


def run_reference(oracle, model, ini_text, d):
    os.makedirs(d, exist_ok=True)
    ini = os.path.join(d, "args.ini")
    open(ini, "w").write(ini_text)
    code = ("import sys, os; sys.path.insert(0, %r); os.chdir(%r); import oracle as O; "
            "sys.exit(O.ref_lib(%r).crd_ref_main(%r, 1))" % (ROOT, d, model, ini.encode()))
    r = subprocess.run(["python", "-c", code], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr
    return r.stdout


def banner(text):
    return text.split("   ----------------------")[0].split("\n\n    ")[0].split("   2")[0]


@pytest.mark.parametrize("vb,inside", [(0, 1), (0, 0), (1, 0)])
def test_fhn_torus_driver_matches_reference_main(crd, oracle, tmp_path, vb, inside):
    if not oracle.ref_available("fhn_torus"):
        pytest.skip("oracle/_ref not built")
    from crdmodel_b200 import build as B
    B.build_drivers()
    ini = FHN_INI.format(vb=vb, inside=inside)
    out_gpu = run_driver("FHNmodel_torus", ini, str(tmp_path / "gpu"))
    out_cpu = run_reference(oracle, "fhn_torus", ini, str(tmp_path / "cpu"))
    head = lambda s: s[:s.index("rtol")]
    assert head(out_gpu) == head(out_cpu)                      # banner, byte for byte
    g, c = tmp_path / "gpu", tmp_path / "cpu"
    assert (g / "FHNmodel_torus_subdomain.000.txt").read_text() == (c / "FHNmodel_torus_subdomain.000.txt").read_text()
    for var in ("u", "v"):
        a = (g / ("FHNmodel_torus_%s.000.txt" % var)).read_text().splitlines()
        b = (c / ("FHNmodel_torus_%s.000.txt" % var)).read_text().splitlines()
        assert len(a) == len(b) == 5
        assert a == b                                            # every output line: identical text
        assert a[1][0] == " " and len(a[1].split()[0]) >= 22     # " %.16e"
    # the same with the host-driven loop instead of the device-resident one
    out_h = run_driver("FHNmodel_torus", ini + "resident = 0\n", str(tmp_path / "gpu_host_loop"))
    for var in ("u", "v"):
        assert (tmp_path / "gpu_host_loop" / ("FHNmodel_torus_%s.000.txt" % var)).read_text() == (c / ("FHNmodel_torus_%s.000.txt" % var)).read_text()
    assert "Steps = " in out_h
    assert "Steps = " in out_gpu and "RHS evaluations = " in out_gpu


def test_goldbeter_torus_driver_matches_reference_main(crd, oracle, tmp_path):
    if not oracle.ref_available("gb_torus"):
        pytest.skip("oracle/_ref not built")
    from crdmodel_b200 import build as B
    B.build_drivers()
    # the reference obtains Zs, Ys from an external script (popen); give it one on PATH that prints them
    bindir = tmp_path / "path"
    bindir.mkdir()
    script = bindir / "SolveGoldbeterODE.py"
    script.write_text("#!/bin/sh\necho '[0.392] [1.6469]'\n")
    script.chmod(0o755)
    os.environ["PATH"] = str(bindir) + os.pathsep + os.environ["PATH"]
    out_gpu = run_driver("GoldbeterModel_torus", GB_INI, str(tmp_path / "gpu"))
    out_cpu = run_reference(oracle, "gb_torus", GB_INI, str(tmp_path / "cpu"))
    head = lambda s: s[:s.index("rtol")]
    assert head(out_gpu) == head(out_cpu)
    g, c = tmp_path / "gpu", tmp_path / "cpu"
    assert (g / "GoldbeterModel_torus_subdomain.000.txt").read_text() == (c / "GoldbeterModel_torus_subdomain.000.txt").read_text()
    for var in ("Z", "Y"):
        a = (g / ("GoldbeterModel_torus_%s.000.txt" % var)).read_text().splitlines()
        b = (c / ("GoldbeterModel_torus_%s.000.txt" % var)).read_text().splitlines()
        assert len(a) == len(b) == 4 and a[0] == b[0]
        A, Bm = np.array([l.split() for l in a], float), np.array([l.split() for l in b], float)
        assert np.all(np.abs(A - Bm) <= 1.0 * (1e-5 * np.abs(Bm) + 1e-10))


def test_driver_usage_and_missing_key(crd, tmp_path):
    from crdmodel_b200 import build as B
    B.build_drivers()
    exe = os.path.join(ROOT, "bin", "FHNmodel_flat")
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode != 0 and "Usage:" in r.stderr
    ini = tmp_path / "bad.ini"
    ini.write_text("[Parameters]\ndiffusion = 0.12\n")
    r = subprocess.run([exe, str(ini)], capture_output=True, text=True)
    assert r.returncode != 0 and "No such node" in r.stderr


def test_shipped_fhn_ini_keys_are_accepted(crd, tmp_path):
    """data/FHNmodelArgs.ini ships xMesh while the FHN programs read thetaMesh: accept both."""
    from crdmodel_b200 import build as B
    B.build_drivers()
    ini = FHN_INI.format(vb=1, inside=0).replace("thetaMesh = 24", "xMesh = 16").replace("tFinal = 2", "tFinal = 0.2")
    out = run_driver("FHNmodel_flat", ini, str(tmp_path / "flat"))
    assert "nx = 16" in out and "ny = 64" in out
    u = np.loadtxt(tmp_path / "flat" / "FHNmodel_flat_u.000.txt")
    assert u.shape == (5, 16 * 64) and np.isfinite(u).all()


FLAT_GB_INI = GB_INI.replace("waveInside = 1\n", "").replace("icType = 0", "icType = 1").replace("varyBeta = 0", "varyBeta = {vb}")


@pytest.mark.parametrize("exe,model,ini", [
    ("FHNmodel_flat", "fhn_flat", FHN_INI.format(vb=0, inside=0)),
    ("GoldbeterModel_flat", "gb_flat", FLAT_GB_INI.format(vb=0)),
    ("GoldbeterModel_flat", "gb_flat", FLAT_GB_INI.format(vb=1)),
])
def test_flat_drivers_match_reference_main(crd, oracle, tmp_path, exe, model, ini):
    if not oracle.ref_available(model):
        pytest.skip("oracle/_ref not built")
    from crdmodel_b200 import build as B
    B.build_drivers()
    bindir = tmp_path / "path"
    bindir.mkdir()
    script = bindir / "SolveGoldbeterODE.py"
    script.write_text("#!/bin/sh\necho '[0.392] [1.6469]'\n")
    script.chmod(0o755)
    os.environ["PATH"] = str(bindir) + os.pathsep + os.environ["PATH"]
    out_gpu = run_driver(exe, ini, str(tmp_path / "gpu"))
    out_cpu = run_reference(oracle, model, ini, str(tmp_path / "cpu"))
    head = lambda s: s[:s.index("rtol")]
    assert head(out_gpu) == head(out_cpu)
    g, c = tmp_path / "gpu", tmp_path / "cpu"
    names = sorted(p.name for p in g.iterdir() if p.name.endswith(".txt"))
    assert names == sorted(p.name for p in c.iterdir() if p.name.endswith(".txt")) and len(names) == 3
    for n in names:
        a, b = (g / n).read_text().splitlines(), (c / n).read_text().splitlines()
        assert len(a) == len(b)
        if "subdomain" in n:
            assert a == b
            continue
        assert a[0] == b[0]
        if exe.startswith("FHN"):
            assert a == b
        A, Bm = np.array([l.split() for l in a], float), np.array([l.split() for l in b], float)
        assert A.shape == Bm.shape and np.all(np.abs(A - Bm) <= 1.0 * (1e-5 * np.abs(Bm) + 1e-10))
