"""Full-run parity and driver-level wall time: the reference's own main() (oracle/_ref: its sources compiled in place, this
repository's RK driver behind the ARKode names, host N_Vector; CPU, `np` emulated MPI ranks) against bin/<driver> (GPU) on the
parameter sets the reference ships (data/FHNmodelArgs.ini, data/GoldbeterModelArgs.ini; the FHN file with thetaMesh added,
which the FHN programs read instead of the shipped xMesh), compared at every output time in units of the integrator tolerance
rtol*|y| + atol, with nst / nfe / netf of both sides.  Prints one JSON record per case and writes
gpurun_out/full_run_parity.json (copied to profiles/).     python tools/tools_full_run_parity.py [small]"""
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
xMesh = {mesh}
thetaMesh = {mesh}
betaMin = 0.7
betaMax = 1.7

[System]
includeAllVars = {allv}
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
includeAllVars = 0
varyBeta = 0
justDiffusion = 0
icType = 2
"""


def run(cmd, cwd, env=None):
    t0 = time.time()
    r = subprocess.run(cmd, cwd=cwd, capture_output=True, text=True, timeout=6000, env=env)
    assert r.returncode == 0, r.stderr[-2000:]
    return r, time.time() - t0


def gather(d, stem, var, nranks):
    """rows of all subdomain files placed by their is..ie / js..je (the post-processing contract)"""
    full = None
    for rk in range(nranks):
        sub = open("%s/%s_subdomain.%03d.txt" % (d, stem, rk)).read().split()
        nx, ny, is_, ie, js, je = (int(x) for x in sub[:6])
        a = np.loadtxt("%s/%s_%s.%03d.txt" % (d, stem, var, rk))
        nt = a.shape[0]
        if full is None:
            full = np.zeros((nt, ny, nx))
        full[:, js:je + 1, is_:ie + 1] = a.reshape(nt, je - js + 1, ie - is_ + 1)
    return full


small = len(sys.argv) > 1 and sys.argv[1] == "small"
cases = [("FHNmodel_torus", "fhn_torus", FHN.format(mesh=128, vb=1, allv=1), "u", 1, "shipped FHN parameters at thetaMesh 128"),
         ("FHNmodel_torus", "fhn_torus", FHN.format(mesh=128, vb=0, allv=1), "u", 2, "the same with varyBeta 0, reference on 2 ranks"),
         ("GoldbeterModel_torus", "gb_torus", GB.format(mesh=100), "Z", 1, "data/GoldbeterModelArgs.ini as shipped (xMesh 100)")]
if not small:
    cases.append(("FHNmodel_torus", "fhn_torus", FHN.format(mesh=400, vb=1, allv=0), "u", 4,
                  "data/FHNmodelArgs.ini as shipped (mesh 400 x 1600, tFinal 50, 20 outputs), reference on 4 ranks (2 x 2)"))
out = []
for exe, model, ini, var, nranks, what in cases:
    d = tempfile.mkdtemp()
    os.makedirs(d + "/gpu"); os.makedirs(d + "/gpu_host"); os.makedirs(d + "/cpu"); os.makedirs(d + "/path")
    open(d + "/path/SolveGoldbeterODE.py", "w").write("#!/bin/sh\necho '[0.392] [1.6469]'\n"); os.chmod(d + "/path/SolveGoldbeterODE.py", 0o755)
    env = dict(os.environ, PATH=d + "/path" + os.pathsep + os.environ["PATH"], CRD_ARK_STATS="1")
    for sub, extra in (("gpu", ""), ("gpu_host", "resident = 0\n"), ("cpu", "")):
        open("%s/%s/a.ini" % (d, sub), "w").write(ini + extra)
    rg, tg = run([os.path.join(ROOT, "bin", exe), "a.ini"], d + "/gpu", env)
    rh, th = run([os.path.join(ROOT, "bin", exe), "a.ini"], d + "/gpu_host", env)
    code = "import sys, os; sys.path.insert(0, %r); import oracle as O; sys.exit(O.ref_lib(%r).crd_ref_main(b'a.ini', %d))" % (ROOT, model, nranks)
    rc, tc = run([sys.executable, "-c", code], d + "/cpu", env)
    A, H, B = gather(d + "/gpu", exe, var, 1), gather(d + "/gpu_host", exe, var, 1), gather(d + "/cpu", exe, var, nranks)
    assert A.shape == B.shape == H.shape
    dev = (np.abs(A - B) / (1e-5 * np.abs(B) + 1e-10)).reshape(A.shape[0], -1).max(axis=1)
    devh = (np.abs(H - B) / (1e-5 * np.abs(B) + 1e-10)).reshape(A.shape[0], -1).max(axis=1)
    rec = {"case": what, "driver": exe, "mesh": "%d x %d" % (A.shape[2], A.shape[1]), "outputs": int(A.shape[0]),
           "gpu_wall_seconds": round(tg, 2), "gpu_host_driven_loop_wall_seconds": round(th, 2),
           "cpu_reference_main_wall_seconds": round(tc, 2), "cpu_ranks": nranks, "cpu_cores": os.cpu_count(),
           "gpu_stats": re.findall(r"crd_ark: (.*)", rg.stderr)[-1], "gpu_host_driven_stats": re.findall(r"crd_ark: (.*)", rh.stderr)[-1],
           "cpu_stats": re.findall(r"crd_ark: (.*)", rc.stderr)[-1],
           "max_deviation_per_output_in_units_of_rtol_y_plus_atol": [round(float(x), 3) for x in dev],
           "same_for_the_host_driven_gpu_loop": [round(float(x), 3) for x in devh],
           "text_identical": bool(np.array_equal(A, B)), "note": "wall time of the whole program: ini -> integrate -> text files"}
    out.append(rec)
    print(json.dumps(rec), flush=True)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "r2_full_run_parity.json"), "w"), indent=1)
