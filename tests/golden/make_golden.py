"""Generates tests/golden/oracle_digests.json: digests of the oracle's outputs on the scaled-down BASELINE configs.
The reference has no golden vectors and cannot be run here, so these were produced by oracle/oracle.cpp (serial
schedule, normalize() in exact-1/sqrt mode so the colour digest does not depend on the host's _mm_rsqrt_ps table).
Run from the repo root:  python tests/golden/make_golden.py"""
import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)


def digest(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()[:32]


def compute():
    import swraster_viewer_b200 as swr
    from helpers import small_configs, render_oracle
    out = {}
    for name, scene, spec, W, H in small_configs():
        cam = swr.RenderCamera.from_spec(spec, W, H)
        o = render_oracle(scene, cam, W, H, exact_rsqrt=True)
        out[name] = {
            "W": W, "H": H, "depth": digest(o["depth"]), "seq": digest(o["seq"]), "bary1": digest(o["bary1"]),
            "bary2": digest(o["bary2"]), "pixels_exact_rsqrt": digest(o["pixels"]),
            "covered": int((o["seq"] != 0xFFFFFFFF).sum()),
            "stats": {k: int(o["stats"][k]) for k in ("triangles_submitted", "vertices_submitted", "triangles_binned", "triangles_clipped", "tile_refs")},
            "sample_seq": [int(x) for x in o["seq"][:: max(1, W * H // 16)][:16]],
        }
    return out


if __name__ == "__main__":
    d = compute()
    with open(os.path.join(ROOT, "tests", "golden", "oracle_digests.json"), "w") as f:
        json.dump(d, f, indent=1, sort_keys=True)
    print("wrote", len(d), "configs")
