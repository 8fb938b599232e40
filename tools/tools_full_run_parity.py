"""Full-run parity: the reference's own main() (oracle/_ref, CPU) vs bin/<driver> (GPU) on the shipped parameter
sets at a reduced mesh, compared at every output time in units of the integrator tolerance rtol*|y| + atol.
Prints a JSON summary (profiles/r01_full_run_parity.json)."""
import json, os, re, subprocess, sys, tempfile, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

FHN = """[Parameters]
diffusion = 0.12
beta = 1.25
surfaceWidth = 20
surfaceLength = 80
waveLength = 0.1
waveWidth = 0.5
waveInside = 0
outputTimestep = 20
tBoundary = 38
tFinal = 50
thetaMesh = {mesh}
betaMin = 0.7
betaMax = 1.7

[System]
includeAllVars = 1
varyBeta = {vb}
"""
GB = """[Parameters]
diffusion = 0.12
beta = 0.4
surfaceWidth = 20
surfaceLength = 80
waveLength = 0.2
waveWidth = 0.5
waveInside = 1
outputTimestep = 5
tBoundary = 0
tFinal = 4
xMesh = {mesh}
betaMin = 0
betaMax = 1
Zs = 0.392
Ys = 1.6469

[System]
includeAllVars = 1
varyBeta = 0
justDiffusion = 0
icType = 2
"""

def run(cmd, cwd, env=None):
    t0 = time.time()
    r = subprocess.run(cmd, cwd=cwd, capture_output=True, text=True, timeout=3000, env=env)
    assert r.returncode == 0, r.stderr[-2000:]
    return r, time.time() - t0

out = []
cases = [("FHNmodel_torus", "fhn_torus", FHN.format(mesh=64, vb=1), "u", "v"), ("FHNmodel_torus", "fhn_torus", FHN.format(mesh=64, vb=0), "u", "v"),
         ("GoldbeterModel_torus", "gb_torus", GB.format(mesh=100), "Z", "Y")]
for exe, model, ini, v0, v1 in cases:
    d = tempfile.mkdtemp()
    os.makedirs(d + "/gpu"); os.makedirs(d + "/cpu"); os.makedirs(d + "/path")
    open(d + "/path/SolveGoldbeterODE.py", "w").write("#!/bin/sh\necho '[0.392] [1.6469]'\n"); os.chmod(d + "/path/SolveGoldbeterODE.py", 0o755)
    env = dict(os.environ, PATH=d + "/path" + os.pathsep + os.environ["PATH"], CRD_ARK_STATS="1")
    for sub in ("gpu", "cpu"):
        open("%s/%s/a.ini" % (d, sub), "w").write(ini)
    rg, tg = run([os.path.join(ROOT, "bin", exe), "a.ini"], d + "/gpu", env)
    code = "import sys, os; sys.path.insert(0, %r); import oracle as O; sys.exit(O.ref_lib(%r).crd_ref_main(b'a.ini', 1))" % (ROOT, model)
    rc, tc = run([sys.executable, "-c", code], d + "/cpu", env)
    rec = {"driver": exe, "mesh": re.search(r"Mesh = (\d+)", ini).group(1), "varyBeta": re.search(r"varyBeta = (\d)", ini).group(1),
           "gpu_seconds": round(tg, 2), "cpu_reference_main_seconds": round(tc, 2),
           "gpu_stats": re.findall(r"crd_ark: (.*)", rg.stderr)[-1], "cpu_stats": re.findall(r"crd_ark: (.*)", rc.stderr)[-1]}
    for var in (v0, v1):
        stem = exe + "_" + var + ".000.txt"
        A, B = np.loadtxt(d + "/gpu/" + stem), np.loadtxt(d + "/cpu/" + stem)
        assert A.shape == B.shape
        dev = np.abs(A - B) / (1e-5 * np.abs(B) + 1e-10)
        rec["max_dev_in_tolerances_per_output_" + var] = [round(float(x), 2) for x in dev.max(axis=1)]
    out.append(rec)
    print(json.dumps(rec), flush=True)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "full_run_parity.json"), "w"), indent=1)
