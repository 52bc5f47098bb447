"""Writes tests/golden/bench_cameras.json: the swr_camera blocks (raw bytes, hex) of the five BASELINE configs as the host
mirror builds them (RenderCamera::new + rotate_mouse, rendercamera.rs:28-150). bench.py --impl reference reads this file so
that the CPU arm never has to load the product libraries; tests/test_host_mirror.py checks the host mirror still
reproduces every entry bit for bit."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import swraster_viewer_b200 as swr  # noqa: E402
import bench  # noqa: E402

out = {}
for name, cfg in bench.CONFIGS.items():
    spec = bench.camera_spec(name)
    cam = swr.RenderCamera.from_spec(spec, cfg["W"], cfg["H"])
    out[name] = {"width": cfg["W"], "height": cfg["H"], "position": list(spec.position), "look_at": list(spec.look_at), "fov": spec.fov,
                 "far_plane": spec.far_plane, "camera_hex": bytes(cam.abi).hex()}
json.dump(out, open(os.path.join(ROOT, "tests", "golden", "bench_cameras.json"), "w"), indent=1)
print("wrote", len(out), "cameras")
