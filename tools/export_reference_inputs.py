"""Writes the procedural BASELINE scenes C1-C3 as glTF (scenes.export_gltf) plus, per scene, the SWR_DIGEST_CAMERA line
tools/reference_digest.rs needs, so that a maintainer with a Rust toolchain can render them with the real
swraster-viewer and drop per-tile digests into tests/golden/reference_digests/ (tests/test_reference_digest.py).
usage: python tools/export_reference_inputs.py OUTDIR [c1 c2 c3]"""
import math
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from swraster_viewer_b200 import scenes  # noqa: E402


def camera_line(spec, W, H):
    """RenderCamera.from_spec (renderer.py) in the reference's own terms: level look-at target + mouse pitch."""
    f = [spec.look_at[i] - spec.position[i] for i in range(3)]
    n = math.sqrt(sum(c * c for c in f))
    f = [c / n for c in f]
    level = (spec.position[0] + f[0], spec.position[1], spec.position[2] + f[2])
    if abs(f[0]) + abs(f[2]) < 1e-6:
        level = (spec.position[0], spec.position[1], spec.position[2] - 1.0)
    dy = -math.asin(max(-1.0, min(1.0, f[1]))) / 0.01
    vals = list(spec.position) + list(level) + [dy, spec.fov, spec.far_plane, W, H]
    return " ".join(repr(float(v)) for v in vals)


def main():
    out = sys.argv[1]
    os.makedirs(out, exist_ok=True)
    for name in sys.argv[2:] or ["c1", "c2", "c3"]:
        cfg = bench.CONFIGS[name]
        scene, spec = bench.build_scene(name)
        path = os.path.join(out, name + ".gltf")
        scenes.export_gltf(scene, path)
        print(f'RAYON_NUM_THREADS=1 SWR_DIGEST_OUT={name}.json SWR_DIGEST_CAMERA="{camera_line(spec, cfg["W"], cfg["H"])}" cargo run --release -- {path}')


if __name__ == "__main__":
    main()
