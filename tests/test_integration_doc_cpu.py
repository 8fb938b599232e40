"""INTEGRATION.md shows the lines a maintainer of the reference changes to bind it to libcrd_b200.so.  Those snippets
must stay valid against include/*.h: they are extracted from the document and compiled (syntax and types only; MPI and
the reference's own variables are declared as stubs)."""
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

PRELUDE = '''
#include "crd_b200.h"
#include "crd_ark.h"
typedef int MPI_Op; typedef int MPI_Comm; typedef int MPI_Datatype;
enum { MPI_SUCCESS = 0, MPI_SUM, MPI_MAX, MPI_MIN, MPI_BYTE, MPI_DOUBLE, MPI_COMM_WORLD };
#define MPI_IN_PLACE ((void *)1)
int MPI_Allgather(const void *, int, MPI_Datatype, void *, int, MPI_Datatype, MPI_Comm);
int MPI_Allreduce(const void *, void *, int, MPI_Datatype, MPI_Op, MPI_Comm);
static int mpi_allreduce_hook(double *v, int n, int op, void *);
// what main() of src/FHNmodel_torus.cpp has in scope at these lines
int rank, nprocs, flag; long nx, ny;
double DIFF, BETA, BETAMIN, BETAMAX, TBOUNDARY, MAJORCIRC, MINORCIRC; int VARYBETA;
N_Vector y; realtype *ydata; void *arkode_mem; realtype T0, tout, t;
'''


def test_integration_snippets_compile_against_the_headers(tmp_path):
    md = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    blocks = re.findall(r"```cpp\n(.*?)```", md, re.S)
    assert len(blocks) >= 2
    body, hook = blocks[0].split("static int mpi_allreduce_hook", 1)
    body = body.replace('#include "crd_b200.h"', "")
    src = (PRELUDE + "void option_a() {\n" + body + "\n}\nstatic int mpi_allreduce_hook" + hook +
           "\nvoid option_b() {\n  crd_grid *grid = 0;   // created as in option A\n" + blocks[1] + "\n}\n")
    f = tmp_path / "integration.cpp"
    f.write_text(src)
    r = subprocess.run(["g++", "-std=c++17", "-fsyntax-only", "-Wno-vla", "-I" + os.path.join(ROOT, "include"), str(f)],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[:3000]
    # every crd_ / N_V..._Crd function the document names is declared by the headers
    hdr = "".join(open(os.path.join(ROOT, "include", h)).read() for h in ("crd_b200.h", "crd_ark.h", "crd_sundials_compat.h"))
    for name in set(re.findall(r"\b(crd_[A-Za-z_]+|N_V\w+_Crd|crd_ARKode\w+)\s*\(", md)):
        assert re.search(r"\b%s\s*\(" % re.escape(name), hdr), name
