"""Generates tests/golden/rhs_baseline_digests.json (+ rhs_baseline_samples.npz) from the reference ITSELF (oracle/_ref = the
reference's f() compiled in place from /root/reference/src) at the mesh sizes of BASELINE.json configs[0..4]: SHA-256 of the
ydot bytes plus a few sampled values; for the two headline meshes (configs[3] FHN torus 16384 x 16384, configs[4] Goldbeter
torus theta 8192 x phi 32768) also 4096 sampled values each (the Goldbeter kinetics go through libm pow and are compared to a
tolerance, so a hash alone cannot pin them).  Run in the build container only (needs /root/reference, ~20 GB of memory and a
few minutes):

    python tests/golden/make_baseline_digests.py
"""
import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import oracle as O  # noqa: E402

# (name, model, nx, ny, t, just_diffusion, variants the GPU test forces in addition to the automatic choice)
CASES = [
    ("cfg1 FHN flat 400x1600 (data/FHNmodelArgs.ini mesh), boundary rows frozen", "fhn_flat", 400, 1600, 10.0, 0, []),
    ("cfg1 FHN flat 400x1600, boundary released", "fhn_flat", 400, 1600, 50.0, 0, []),
    ("cfg2 FHN torus 400x1600 (default ini grid), boundary rows frozen", "fhn_torus", 400, 1600, 10.0, 0, []),
    ("cfg2 FHN torus 400x1600, boundary released", "fhn_torus", 400, 1600, 50.0, 0, []),
    ("cfg3 Goldbeter torus 100x400, diffusion only (no pow: bit-exact)", "gb_torus", 100, 400, 50.0, 1, []),
    ("cfg3 Goldbeter torus 100x400, full kinetics (libm pow: compare sampled values to 4e-16 of the terms)", "gb_torus", 100, 400, 50.0, 0, []),
    # a full-width band of the headline mesh through every kernel the automatic choice can pick for large slabs
    ("FHN torus 16384x512 (one full-width band), boundary rows frozen", "fhn_torus", 16384, 512, 10.0, 0, [13, 15, 20, 21]),
    ("FHN torus 16384x512 (one full-width band), boundary released", "fhn_torus", 16384, 512, 50.0, 0, [13, 15, 20, 21]),
    ("Goldbeter torus 8192x512, diffusion only", "gb_torus", 8192, 512, 50.0, 1, [13, 15, 20, 21]),
    # the headline meshes themselves
    ("cfg4 FHN torus 16384x16384 (BASELINE configs[3]), boundary rows frozen", "fhn_torus", 16384, 16384, 10.0, 0, []),
    ("cfg4 FHN torus 16384x16384 (BASELINE configs[3]), boundary released", "fhn_torus", 16384, 16384, 50.0, 0, []),
    ("cfg5 Goldbeter torus 8192x32768 (BASELINE configs[4]), diffusion only (bit-exact)", "gb_torus", 8192, 32768, 50.0, 1, []),
    ("cfg5 Goldbeter torus 8192x32768 (BASELINE configs[4]), full kinetics (sampled values)", "gb_torus", 8192, 32768, 50.0, 0, []),
]
SEED = 0x5EED
NSAMPLES = 4096


def main():
    O.build()
    out, samples = [], {}
    for k, (name, model, nx, ny, t, jd, variants) in enumerate(CASES):
        P = O.make_params(model, nx, ny, just_diffusion=jd, t_boundary=38.0)
        y = O.fill_state(model, 2 * nx * ny, seed=SEED)
        ydot, sec = O.ref_rhs(P, t, y)
        idx = [0, 1, 2 * nx + 3, nx * ny, 2 * nx * ny - 2, 2 * nx * ny - 1]
        row = {"name": name, "model": model, "nx": nx, "ny": ny, "t": t, "just_diffusion": jd, "seed": SEED,
               "variants": variants, "sha256": hashlib.sha256(ydot.tobytes()).hexdigest(),
               "samples": {str(i): float(ydot[i]).hex() for i in idx}}
        if nx * ny >= (1 << 26):
            # sampled elements: fixed pseudo-random points (both variables), plus the first / last rows' ends
            rng = np.random.default_rng(1000 + k)
            pts = np.unique(np.concatenate([rng.integers(0, nx * ny, NSAMPLES - 8), [0, nx - 1, nx, nx * ny - nx - 1, nx * ny - nx, nx * ny - 1, nx * (ny // 2), nx * (ny // 2) + nx - 1]]))
            el = np.sort(np.concatenate([2 * pts, 2 * pts + 1]))
            samples["idx_%02d" % k] = el.astype(np.int64)
            samples["val_%02d" % k] = ydot[el]
            row["sample_set"] = "%02d" % k
        out.append(row)
        print("%-95s %6.1f s  %s" % (name, sec, row["sha256"][:16]), flush=True)
        del y, ydot
    here = os.path.dirname(os.path.abspath(__file__))
    json.dump(out, open(os.path.join(here, "rhs_baseline_digests.json"), "w"), indent=1)
    np.savez_compressed(os.path.join(here, "rhs_baseline_samples.npz"), **samples)
    print("wrote", len(out), "cases")


if __name__ == "__main__":
    main()
