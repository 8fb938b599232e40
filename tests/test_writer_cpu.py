"""The drivers' background output writer (crdmodel_b200/host/crd_writer.hpp) emits exactly the bytes of the
reference's fprintf(" %.16e") loops (src/FHNmodel_torus.cpp:393-410,438-455)."""
import filecmp
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_async_writer_bytes_equal_fprintf(tmp_path):
    exe = str(tmp_path / "writer_check")
    subprocess.run(["g++", "-O2", "-std=c++17", "-I" + os.path.join(ROOT, "crdmodel_b200", "host"),
                    os.path.join(ROOT, "tests", "cpp", "writer_check.cpp"), "-o", exe, "-lpthread"], check=True)
    a, b, r = (str(tmp_path / n) for n in ("a.txt", "b.txt", "ref.txt"))
    subprocess.run([exe, a, b, r], check=True)
    assert filecmp.cmp(a, r, shallow=False)
    assert os.path.getsize(b) > 0 and sum(1 for _ in open(b)) == 3
