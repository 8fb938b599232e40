"""The drivers' ini reader (crdmodel_b200/host/crd_ini.hpp) honours the lookup contract the reference gets from
Boost.PropertyTree (src/FHNmodel_torus.cpp:158-174): sections, whole-line comments, throwing lookups, fallbacks."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

INI = """top = 7
; a comment
# another one
[Parameters]
diffusion = 0.12
beta=1.25
  xMesh   =   400
tFinal = 50
comment_like = a # b ; c
word = abc
sci_int = 1e3
trailing = 0.4abc
two_numbers = 1 2
padded =    2.5   

[System]
varyBeta = 1
"""


def test_ini_reader_contract(tmp_path):
    exe = str(tmp_path / "ini_check")
    subprocess.run(["g++", "-O2", "-std=c++17", "-I" + os.path.join(ROOT, "crdmodel_b200", "host"),
                    os.path.join(ROOT, "tests", "cpp", "ini_check.cpp"), "-o", exe], check=True)
    ini = tmp_path / "a.ini"
    ini.write_text(INI)
    r = subprocess.run([exe, str(ini)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    bad = tmp_path / "bad.ini"
    bad.write_text("[Parameters\nx = 1\n")
    r = subprocess.run([exe, str(bad)], capture_output=True, text=True)
    assert r.returncode != 0          # unmatched '[' is an error, like Boost's ini_parser_error
    dup = tmp_path / "dup.ini"
    dup.write_text(INI.replace("tFinal = 50", "tFinal = 50\ndiffusion = 0.5"))
    r = subprocess.run([exe, str(dup)], capture_output=True, text=True)
    assert r.returncode != 0          # Boost: "duplicate key name"
