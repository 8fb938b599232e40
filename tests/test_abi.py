"""The C-ABI library builds, loads and exports every symbol include/*.h declares (no GPU needed)."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions(headers=("crd_b200.h", "crd_ark.h", "crd_sundials_compat.h")):
    names = set()
    for h in headers:
        src = open(os.path.join(ROOT, "include", h)).read()
        src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
        src = re.sub(r'extern\s+"C"\s*\{', "", src)
        # drop every remaining brace block (struct / enum bodies hold function-pointer members)
        out, depth = [], 0
        for ch in src:
            if ch == "{":
                depth += 1
            elif ch == "}":
                depth = max(0, depth - 1)
            elif depth == 0:
                out.append(ch)
        src = "".join(out)
        src = re.sub(r"typedef[^;]*\(\s*\*\s*\w+\s*\)[^;]*;", "", src)       # function-pointer typedefs
        src = re.sub(r"^\s*#.*$", "", src, flags=re.M)
        for m in re.finditer(r"\b([A-Za-z_]\w*)\s*\(([^;{}]*)\)\s*;", src):
            names.add(m.group(1))
    return sorted(names)


def test_header_symbols_exported(crd):
    """include/crd_b200.h -> libcrd_b200.so (the device path); include/crd_sundials_compat.h + crd_ark.h (the ARKode-legacy
    names, the generic N_VXxx dispatchers, the RK driver's extensions) -> libcrd_ark.so, and ONLY there: a program that links
    the real SUNDIALS 2.x beside libcrd_b200.so (option A) must never see ARKode or N_VLinearSum twice."""
    from crdmodel_b200 import _lib
    b200, ark = C.CDLL(_lib.LIB_PATH), C.CDLL(_lib.ARK_LIB_PATH)
    dev = declared_functions(("crd_b200.h",))
    host = sorted(set(declared_functions(("crd_ark.h", "crd_sundials_compat.h"))))
    assert len(dev) > 90 and len(host) > 40
    dev_only = [n for n in dev if n not in host]
    assert not [n for n in dev_only if not hasattr(b200, n)]
    assert not [n for n in host if not hasattr(ark, n)]
    assert not [n for n in host if hasattr(b200, n)], "libcrd_b200.so must not define SUNDIALS' names"


def test_sundials_names_bind_to_whatever_the_program_links(crd, tmp_path):
    """Option A of INTEGRATION.md in miniature: a library standing in for libsundials_arkode / libsundials_nvecparallel
    defines ARKodeCreate and N_VLinearSum; a program linked against it AND libcrd_b200.so gets the stand-in's, whatever the
    link order."""
    import subprocess
    from crdmodel_b200 import build as B
    inc = os.path.join(ROOT, "include")
    stub = tmp_path / "stub.c"
    stub.write_text('#include "crd_sundials_compat.h"\nstatic int marker = 4242;\n'
                    'void *ARKodeCreate(void) { return &marker; }\n'
                    'void N_VLinearSum(realtype a, N_Vector x, realtype b, N_Vector y, N_Vector z) { (void)a; (void)x; (void)b; (void)y; (void)z; marker = 17; }\n'
                    'int stub_marker(void) { return marker; }\n')
    subprocess.run(["gcc", "-std=c99", "-shared", "-fPIC", "-I" + inc, str(stub), "-o", str(tmp_path / "libstub.so")], check=True)
    prog = tmp_path / "prog.c"
    prog.write_text('#include "crd_b200.h"\nint stub_marker(void);\n'
                    'int main(void) { int *m = (int *)ARKodeCreate(); if (!m || *m != 4242) return 1;\n'
                    '  N_VLinearSum(1.0, 0, 1.0, 0, 0); if (stub_marker() != 17) return 2;\n'
                    '  return crd_device_count() >= 0 ? 0 : 3; }\n')
    for order in (["-lstub", "-lcrd_b200"], ["-lcrd_b200", "-lstub"]):
        exe = tmp_path / "prog"
        subprocess.run(["gcc", "-std=c99", "-I" + inc, str(prog), "-o", str(exe), "-L" + str(tmp_path), "-L" + B.LIB_DIR] + order +
                       ["-Wl,-rpath," + str(tmp_path), "-Wl,-rpath," + B.LIB_DIR], check=True)
        r = subprocess.run([str(exe)], capture_output=True, text=True)
        assert r.returncode == 0, (order, r.returncode, r.stderr)


def test_binding_covers_header(crd):
    from crdmodel_b200 import _lib
    assert set(declared_functions()) <= set(_lib.SIGNATURES), set(declared_functions()) - set(_lib.SIGNATURES)


def test_no_cpu_fallback(crd):
    """Without a GPU the product must fail loudly, not compute on the CPU."""
    if crd.lib().crd_device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(crd.CrdError):
        crd.Context(0)


def test_product_does_not_import_oracle():
    """Only tests/, bench.py and __graft_entry__.py may touch oracle/."""
    pkg = os.path.join(ROOT, "crdmodel_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".c", ".h", ".hpp")):
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "import oracle" not in txt and "from oracle" not in txt and "crd_oracle" not in txt \
                    and "liboracle" not in txt and "_ref" not in txt.replace("halo_ref", ""), os.path.join(dirpath, f)


def test_decomp_phi_matches_reference_formula(crd):
    for ny in (1600, 400, 16384, 1601, 37):
        for nr in (1, 2, 3, 4, 8):
            rows = []
            for r in range(nr):
                js, je = crd.decomp_phi(ny, nr, r)
                assert js == ny * r // nr and je == ny * (r + 1) // nr - 1
                rows += list(range(js, je + 1))
            assert rows == list(range(ny))


def test_headers_are_plain_c(tmp_path):
    """The boundary is a C ABI: every public header must compile as C99 (cgo / JNI / ctypes generators include it as C) and
    as C++11, on its own, and a C program using the structs must link against the library."""
    import subprocess
    inc = os.path.join(ROOT, "include")
    for h in ("crd_b200.h", "crd_ark.h", "crd_sundials_compat.h"):
        for lang, std in (("c", "-std=c99"), ("c++", "-std=c++11")):
            r = subprocess.run(["gcc", "-x", lang, std, "-pedantic", "-Wall", "-Werror", "-I" + inc, "-fsyntax-only", "-"],
                               input='#include "%s"\n' % h, capture_output=True, text=True)
            assert r.returncode == 0, (h, lang, r.stderr)
    from crdmodel_b200 import build as B
    B.build()
    src = tmp_path / "use.c"
    src.write_text('#include <stdio.h>\n#include "crd_b200.h"\n#include "crd_ark.h"\n'
                   'int main(void) { int64_t js, je; crd_params p; p.nx = 4; (void)p;\n'
                   '  if (crd_decomp_phi(10, 2, 1, &js, &je) != 0 || js != 5 || je != 9) return 1;\n'
                   '  void *m = ARKodeCreate(); if (!m) return 2; ARKodeFree(&m);\n'
                   '  printf("%d\\n", crd_device_count()); return 0; }\n')
    exe = tmp_path / "use"
    subprocess.run(["gcc", "-std=c99", "-I" + inc, str(src), "-o", str(exe), "-L" + B.LIB_DIR, "-lcrd_ark", "-lcrd_b200",
                    "-Wl,-rpath," + B.LIB_DIR], check=True)
    r = subprocess.run([str(exe)], capture_output=True, text=True)
    assert r.returncode == 0, (r.returncode, r.stderr)
