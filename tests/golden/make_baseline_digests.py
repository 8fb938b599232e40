"""Generates tests/golden/rhs_baseline_digests.json from the reference ITSELF (oracle/_ref = the reference's f() compiled in
place from /root/reference/src) at the mesh sizes of BASELINE.json configs[0..2]: SHA-256 of the ydot bytes plus a few sampled
values.  Run in the build container only (needs /root/reference):

    python tests/golden/make_baseline_digests.py
"""
import hashlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import oracle as O  # noqa: E402

# (name, model, nx, ny, t, just_diffusion)
CASES = [
    ("cfg1 FHN flat 400x1600 (data/FHNmodelArgs.ini mesh), boundary rows frozen", "fhn_flat", 400, 1600, 10.0, 0),
    ("cfg1 FHN flat 400x1600, boundary released", "fhn_flat", 400, 1600, 50.0, 0),
    ("cfg2 FHN torus 400x1600 (default ini grid), boundary rows frozen", "fhn_torus", 400, 1600, 10.0, 0),
    ("cfg2 FHN torus 400x1600, boundary released", "fhn_torus", 400, 1600, 50.0, 0),
    ("cfg3 Goldbeter torus 100x400, diffusion only (no pow: bit-exact)", "gb_torus", 100, 400, 50.0, 1),
    ("cfg3 Goldbeter torus 100x400, full kinetics (libm pow: compare sampled values to 4e-16 of the terms)", "gb_torus", 100, 400, 50.0, 0),
]
SEED = 0x5EED


def main():
    O.build()
    out = []
    for name, model, nx, ny, t, jd in CASES:
        P = O.make_params(model, nx, ny, just_diffusion=jd, t_boundary=38.0)
        y = O.fill_state(model, 2 * nx * ny, seed=SEED)
        ydot, _ = O.ref_rhs(P, t, y)
        idx = [0, 1, 2 * nx + 3, nx * ny, 2 * nx * ny - 2, 2 * nx * ny - 1]
        out.append({"name": name, "model": model, "nx": nx, "ny": ny, "t": t, "just_diffusion": jd, "seed": SEED,
                    "sha256": hashlib.sha256(ydot.tobytes()).hexdigest(),
                    "samples": {str(i): float(ydot[i]).hex() for i in idx}})
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "rhs_baseline_digests.json")
    json.dump(out, open(path, "w"), indent=1)
    print("wrote", len(out), "cases to", path)


if __name__ == "__main__":
    main()
