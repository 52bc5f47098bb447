"""The procedural scenes can be written as glTF 2.0 so the real swraster-viewer can load them elsewhere.
Round trip: re-read the .gltf/.bin with a minimal parser and compare every accessor with the in-memory arrays."""
import json
import os

import numpy as np

from swraster_viewer_b200 import scenes
from helpers import SMALL


def test_export_gltf_roundtrip(tmp_path):
    sc, _ = scenes.scene_translucent_test(**SMALL)
    path = str(tmp_path / "scene")
    scenes.export_gltf(sc, path)
    g = json.load(open(path + ".gltf"))
    blob = open(path + ".bin", "rb").read()
    assert g["asset"]["version"] == "2.0" and len(blob) == g["buffers"][0]["byteLength"]

    def read(acc_i):
        a = g["accessors"][acc_i]
        v = g["bufferViews"][a["bufferView"]]
        n = {"SCALAR": 1, "VEC2": 2, "VEC3": 3, "VEC4": 4}[a["type"]]
        dt = {5126: np.float32, 5125: np.uint32}[a["componentType"]]
        return np.frombuffer(blob, dt, a["count"] * n, v["byteOffset"]).reshape(a["count"], n)

    flat = [p for m in sc.meshes for p in m]
    gflat = [p for m in g["meshes"] for p in m["primitives"]]
    assert len(flat) == len(gflat)
    for p, gp in zip(flat, gflat):
        assert np.array_equal(read(gp["attributes"]["POSITION"]), p.positions[:, :3])
        assert np.array_equal(read(gp["attributes"]["NORMAL"]), p.normals[:, :3])
        assert np.array_equal(read(gp["attributes"]["TANGENT"]), p.tangents)
        assert np.array_equal(read(gp["attributes"]["TEXCOORD_0"]), p.texcoords)
        assert np.array_equal(read(gp["indices"]).reshape(-1), p.indices)
        assert gp["material"] == p.material_index
    assert len(g["nodes"]) == len(sc.nodes)
    for n, gn in zip(sc.nodes, g["nodes"]):
        assert np.allclose(gn["matrix"], n.transform) and gn["mesh"] == n.mesh_index
    tr = [m for m in g["materials"] if "extensions" in m]
    assert len(tr) == 2 and "KHR_materials_transmission" in g["extensionsUsed"]
    assert os.path.exists(str(tmp_path / "scene_tex0.png"))
