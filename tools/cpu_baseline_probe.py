"""How fast can the CPU port go on this host? Sweep thread counts (run with different OMP_* env to compare)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import swraster_viewer_b200 as swr
import bench, oracle as orc
scene, spec = bench.build_scene()
cam = swr.RenderCamera.from_spec(spec, bench.W, bench.H)
o = orc.Oracle(bench.W, bench.H)
for nt in [int(x) for x in sys.argv[1:]] or [16]:
    ts = []
    for i in range(4):
        t0 = time.perf_counter(); o.render(scene, cam.abi, nthreads=nt, shade=True, fresh=False, outputs=False); o.resolve(2.0, nt); ts.append(time.perf_counter() - t0)
    st = o.stats.as_dict()
    print(f"threads {nt}: frames {[round(t*1e3) for t in ts]} ms | clipbin {st['ms_clipbin']:.0f} raster+shade {st['ms_raster']:.0f} resolve {st['ms_resolve']:.1f}", flush=True)
