"""Goldbeter steady state of the drivers (crdmodel_b200/host/crd_steady.hpp): the closed form is a fixed point of the
reference's kinetics (GoldbeterModel_torus.cpp:694-695,715-716), and the reference's own protocol — popen of
`SolveGoldbeterODE.py <beta>` printing "[Zs] [Ys]" (:254-261) — is honoured when System.steadyStateCommand asks for it."""
import os
import stat
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def exe(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("steady") / "steady_check")
    subprocess.run(["g++", "-O2", "-std=c++17", "-I" + os.path.join(ROOT, "crdmodel_b200", "host"),
                    os.path.join(ROOT, "tests", "cpp", "steady_check.cpp"), "-o", out], check=True)
    return out


@pytest.mark.parametrize("beta", [0.0, 0.1, 0.289, 0.4, 0.774, 0.9, 1.0])
def test_closed_form_is_a_fixed_point(exe, beta):
    Z, Y, rz, ry = (float(v) for v in subprocess.run([exe, "closed", str(beta)], capture_output=True, text=True, check=True).stdout.split())
    assert Z == (1.0 + 7.3 * beta) / 10.0
    assert Y > 0.0 and abs(rz) < 1e-12 and abs(ry) < 1e-12
    if beta == 0.4:     # the shipped data/GoldbeterModelArgs.ini
        assert abs(Z - 0.392) < 1e-15 and abs(Y - 1.6456214671440605) < 1e-12


def test_script_protocol(exe, tmp_path):
    script = tmp_path / "SolveGoldbeterODE.py"
    script.write_text("#!/bin/sh\n# stands in for util/GoldbeterModel/SolveGoldbeterODE.py: same output format\necho \"[ 0.39200001] [ 1.64690002]\"\n")
    script.chmod(script.stat().st_mode | stat.S_IXUSR)
    out = subprocess.run([exe, "command", str(script), "0.4"], capture_output=True, text=True, check=True).stdout.split()
    assert out[0] == "ok" and float(out[1]) == 0.39200001 and float(out[2]) == 1.64690002
    # a command that is missing or prints something else is reported, not silently used
    assert subprocess.run([exe, "command", str(tmp_path / "missing.py"), "0.4"], capture_output=True, text=True).stdout.strip() == "failed"
    bad = tmp_path / "bad.sh"
    bad.write_text("#!/bin/sh\necho hello\n")
    bad.chmod(bad.stat().st_mode | stat.S_IXUSR)
    assert subprocess.run([exe, "command", str(bad), "0.4"], capture_output=True, text=True).stdout.strip() == "failed"
