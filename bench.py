#!/usr/bin/env python
"""bench.py — frames/s and Mtriangles/s of the rasterisation hot path at 3840x2160 (BASELINE.json metric).

Workload (config.workload): C3 = BASELINE.json configs[2], the configuration the metric is quoted on:
procedural ~10 M-triangle instanced scene (3 base meshes, seeded scatter, camera inside the cloud, heavy
clipping), 3840x2160. One step = one frame = Renderer.render_scene + blit (set-up, clip, bin, raster,
vis-buffer shading, resolve).

  value     device-resident throughput: swr_render + swr_resolve(NULL) timed with CUDA events on the
            library's stream, inputs already in HBM, L2 flushed between frames (256 MiB write).
  e2e       the same frames through the public host API (Renderer.render_scene + blit_to_buffer into a pinned
            host buffer): draw-list build + H2D of camera/draw table + all kernels + D2H of the RGBA8 frame,
            wall clock, max over ranks.
  N > 1     sort-first: every rank holds the scene and owns a contiguous range of tile rows; each rank's resolve kernel
            stores its RGBA8 rows straight into rank 0's frame over NVLink peer memory (swr_peer_*), inside the timed
            region ("scaling": "strong" — total work fixed).
  --impl reference   the reference's CPU path (oracle port: C++ restatement with the reference's parallel
            structure, all host threads) on the same config, rank 0 only.
"""
import argparse
import json

import numpy as np
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

W, H = 3840, 2160
WORKLOAD = "C3: procedural 10M-triangle instanced scene (icosphere 20480 / torus 8192 / box-grid 1200 tris, seeded scatter, scales 0.05-4, camera inside the cloud, heavy clipping), 3840x2160"
KERNELS_PER_FRAME = 11  # k_cull, k_compact, k_setup, k_clip, k_scan_tiles, k_scatter, k_scatter_list, k_raster_tiles, k_shade, k_luminance, k_resolve


def build_scene():
    from swraster_viewer_b200 import scenes
    return scenes.scene_c3_instanced(voxel_dim=128, cube_size=256)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p))["hbm_gbs"], "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=lambda: self.lines.extend(self.proc.stdout), daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.06)
        self.proc.terminate()
        self.t.join(timeout=2)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons), "samples": len(sm)}


def algorithmic_bytes(st, w, h):
    """B_frame = 12 T + 16 V + 8 R + 8 W H + 4 W H."""
    return 12 * st["triangles_submitted"] + 16 * st["vertices_submitted"] + 8 * st["tile_refs"] + 12 * w * h


def cpu_frame_times(scene, cam_abi, steps, warmup, nthreads):
    """Times the oracle (CPU port of the reference path, reference's parallel structure) on full frames."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle as orc
    o = orc.Oracle(W, H)
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        o.render(scene, cam_abi, nthreads=nthreads, shade=True, fresh=False, outputs=False)
        o.resolve(2.0, nthreads)
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    return times, o.stats.as_dict()


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import swraster_viewer_b200 as swr
    scene, spec = build_scene()
    cam = swr.RenderCamera.from_spec(spec, W, H)
    ncores = os.cpu_count() or 1
    steps = max(1, args.steps)
    warmup = max(0, args.warmup)
    # bounded: keep the whole run within a few minutes whatever K/W the driver passes
    probe, st = cpu_frame_times(scene, cam.abi, 1, 0, ncores)
    budget = 150.0
    per = probe[0]
    steps_eff = max(1, min(steps, int(budget / per)))
    warm_eff = min(warmup, 1)
    times, st = cpu_frame_times(scene, cam.abi, steps_eff, warm_eff, ncores)
    ms = 1e3 * sum(times) / len(times)
    fps = 1e3 / ms
    T = st["triangles_submitted"]
    line = {
        "impl": "reference", "metric": "frames_per_sec_3840x2160", "value": fps, "unit": "frames/s", "n_gpus": args.gpus,
        "steps": steps_eff, "warmup": warm_eff, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "mtriangles_per_sec": fps * T / 1e6,
        "config": {"workload": WORKLOAD, "width": W, "height": H, "triangles_submitted": T, "scene_triangles": scene.total_triangles},
        "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": ncores, "kind": "port",
                         "sample": f"{steps_eff} full frames of the same workload (C++ restatement of swraster-viewer's rayon+glam path; Rust toolchain unavailable)",
                         "ms_per_frame": ms, "ms_clipbin": st["ms_clipbin"], "ms_raster_shade": st["ms_raster"], "ms_resolve": st["ms_resolve"]},
        "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def run_gpu(args):
    import torch
    import torch.distributed as dist
    import swraster_viewer_b200 as swr
    from swraster_viewer_b200 import abi
    from swraster_viewer_b200.multigpu import tile_row_ranges, balanced_row_ranges, PeerAssembly, device_tensor

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch with torchrun --nproc-per-node N for --gpus N > 1")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    scene, spec = build_scene()
    cam = swr.RenderCamera.from_spec(spec, W, H)
    r = swr.Renderer(W, H, device=local)
    tiles_y = r.tiles_y
    ranges = tile_row_ranges(tiles_y, world)
    if world > 1:
        # probe frame on every rank (full screen, identical everywhere) -> cost-balanced contiguous row bands
        for _ in range(3):
            r.render_scene(scene, cam, shade=False)
        cyc = torch.from_numpy(r.read_tile_costs()[1].astype(np.int64)).to(f"cuda:{local}")
        dist.broadcast(cyc, src=0)  # measured cycles differ slightly per rank: everybody uses rank 0's
        ranges = balanced_row_ranges(cyc.cpu().numpy(), world)
        r.set_tile_rows(*ranges[rank])
    buf = swr.RenderBuffer(W, H, pinned=True)
    stream = torch.cuda.ExternalStream(r.cuda_stream(), device=local)
    pix = device_tensor(r.device_pixels_ptr(), W * H * 4, torch.int32, f"cuda:{local}").view(H, W)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=f"cuda:{local}")

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # N > 1: the frame is assembled in rank 0's pixel buffer by the other ranks' resolve kernels (peer stores over NVLink
    # + a device-side handshake, include/swr.h swr_peer_*): no collective on the data path
    pa = PeerAssembly(r, dst=0) if world > 1 else None

    def step_device():
        r.render_scene(scene, cam)
        if world > 1:
            pa.frame(2.0)
            pa.release()
        else:
            r.resolve_device_only(2.0)

    def step_e2e():
        r.render_scene(scene, cam)
        if world > 1:
            pa.frame(2.0)
            if rank == 0:
                # D2H of the assembled frame on rank 0 (the caller's RenderBuffer), then hand the buffer back
                with torch.cuda.stream(stream):
                    buf._t.view(H, W).copy_(pix, non_blocking=True)
            pa.release()
            stream.synchronize()
        else:
            r.blit_to_buffer(buf)

    # ---- warm-up (also settles buffer growth) -------------------------------------------------------------
    for _ in range(max(3, args.warmup)):
        step_device()
    r.synchronize()
    st0 = r.stats()

    # ---- device-resident timing: CUDA events on the launching stream, L2 flushed between frames ------------
    sampler = ClockSampler(local)
    barrier()
    if rank == 0:
        sampler.start()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    phase = {"ms_setup_bin": 0.0, "ms_raster": 0.0, "ms_shade": 0.0}
    for i in range(args.steps):
        with torch.cuda.stream(stream):
            flush.fill_(i & 0xFF)  # evict L2 (126 MB) between timed frames
            ev[i][0].record(stream)
        step_device()
        ev[i][1].record(stream)
        r.synchronize()
        s = r.stats()
        for k in phase:
            phase[k] += s[k]
    barrier()
    dev_ms = sum(a.elapsed_time(b) for a, b in ev)
    # ---- end-to-end timing: public API, host buffers, wall clock ---------------------------------------------
    for _ in range(2):
        step_e2e()
    barrier()
    if world == 1:
        # pipelined like the reference's App (present N-1 || render N, main.rs:526-597): every step still uploads its draw
        # table and reads its own 33 MB frame back into pinned host memory; the read-back of frame N overlaps frame N+1
        bufs = [buf, swr.RenderBuffer(W, H, pinned=True)]
        prev = None
        t0 = time.perf_counter()
        for i in range(args.steps):
            r.render_scene(scene, cam)
            tk = r.blit_to_buffer_async(bufs[i & 1])
            if prev is not None:
                r.wait_blit(prev)
            prev = tk
        r.wait_blit(prev)
        e2e_s = time.perf_counter() - t0
        # and the strictly synchronous form (render, blit, wait) for reference
        t0 = time.perf_counter()
        for _ in range(args.steps):
            step_e2e()
        e2e_sync_s = time.perf_counter() - t0
    else:
        t0 = time.perf_counter()
        for _ in range(args.steps):
            step_e2e()
        barrier()
        e2e_s = time.perf_counter() - t0
        e2e_sync_s = e2e_s
    clocks = sampler.stop() if rank == 0 else None

    t = torch.tensor([dev_ms, e2e_s], dtype=torch.float64, device=f"cuda:{local}")
    cnt = torch.tensor([st0["tile_refs"], st0["triangles_binned"]], dtype=torch.float64, device=f"cuda:{local}")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
    dev_ms, e2e_s = float(t[0]), float(t[1])
    if rank == 0:
        K = args.steps
        ms = dev_ms / K
        fps = 1e3 / ms
        T = st0["triangles_submitted"]
        stats_all = dict(st0)
        stats_all["tile_refs"] = int(cnt[0])
        B = algorithmic_bytes(stats_all, W, H)
        peak, peak_src = peaks()
        # dominant kernel of the step on this rank
        dom = max(phase, key=lambda k: phase[k])
        dom_ms = phase[dom] / K
        kname = {"ms_setup_bin": "k_setup+k_scan_tiles+k_scatter", "ms_raster": "k_raster_tiles", "ms_shade": "k_shade"}[dom]
        kbytes = {"ms_setup_bin": 12 * T + 16 * st0["vertices_submitted"] + 4 * st0["tile_refs"],
                  "ms_raster": 4 * st0["tile_refs"] + 8 * W * H // world,
                  "ms_shade": 4 * W * H // world}[dom]
        achieved = kbytes / (dom_ms * 1e-3) / 1e9
        # DRAM traffic of the dominant kernel from the committed `ncu --set full` capture of this same command (N=1)
        traffic = None
        tj = os.path.join(ROOT, "profiles", "r01_traffic.json")
        if world == 1 and os.path.exists(tj):
            traffic = json.load(open(tj)).get(kname.split("+")[0])
        # What really bounds these kernels is instruction issue, not HBM: warp instructions per launch from the
        # same committed ncu capture against the SM issue rate (4 schedulers x 1 warp instruction / clock / SM).
        issue = None
        ij = os.path.join(ROOT, "profiles", "r01_warp_inst.json")
        if world == 1 and os.path.exists(ij) and clocks and clocks.get("sm_mhz"):
            winst = json.load(open(ij)).get(kname.split("+")[0])
            if winst:
                ipeak = torch.cuda.get_device_properties(local).multi_processor_count * 4 * clocks["sm_mhz"] * 1e6 / 1e9
                issue = {"warp_inst_per_launch": winst, "achieved": winst / (dom_ms * 1e-3) / 1e9, "peak": ipeak, "unit": "G warp-inst/s",
                         "frac": winst / (dom_ms * 1e-3) / 1e9 / ipeak, "source": "smsp__inst_executed.sum (profiles/r01_warp_inst.json) / live kernel time"}
        h2d = r.num_draws * (144 + 4) + 4
        line = {
            "metric": "frames_per_sec_3840x2160", "value": fps, "unit": "frames/s", "n_gpus": world, "steps": K, "warmup": max(3, args.warmup),
            "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "mtriangles_per_sec": fps * T / 1e6,
            "config": {"workload": WORKLOAD, "width": W, "height": H, "triangles_submitted": T, "scene_triangles": scene.total_triangles,
                       "vertices_submitted": st0["vertices_submitted"], "tile_refs": int(cnt[0]), "triangles_binned": int(cnt[1]),
                       "l2": "256 MiB device write between timed frames (L2 flush)", "parallelism": (f"sort-first x{world}: cost-balanced contiguous tile-row bands {ranges}, per-band draw culling, peer-store frame assembly on rank 0 (NVLink P2P, no collective)" if world > 1 else "single GPU")},
            "e2e": {"value": K / e2e_s, "unit": "frames/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": W * H * 4, "ms_per_step": 1e3 * e2e_s / K,
                    "mode": "pipelined: read-back of frame N overlaps frame N+1 (swr_resolve_async)" if world == 1 else "synchronous + peer-store frame assembly on rank 0 (NVLink P2P, no collective)",
                    "synchronous_value": K / e2e_sync_s},
            # our kernels launched inside the timed regions on this rank: K device-timed frames + K e2e frames (+ K synchronous e2e frames at N=1;
            # + k_peer_wait / k_peer_release per frame on the assembling rank at N>1)
            "gpu_launches": (KERNELS_PER_FRAME * 3 * K) if world == 1 else ((KERNELS_PER_FRAME + 2) * 2 * K),
            "clocks": clocks,
            "roofline": {"bound": "hbm", "kernel": kname, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                         "peak_source": peak_src, "kernel_ms": dom_ms, "kernel_algorithmic_bytes": kbytes,
                         "frame": {"algorithmic_bytes": B, "achieved": B / (ms * 1e-3) / 1e9, "frac": B / (ms * 1e-3) / 1e9 / peak},
                         "phase_ms": {k: v / K for k, v in phase.items()}, "issue": issue},
        }
        if world == 1 and not args.no_cpu:
            # The CPU port is timed in a fresh interpreter: inside this process (torch and its OpenMP runtime loaded, CUDA
            # context alive) the same code runs about 1.5x slower, which would flatter the GPU arm.
            env = {k: v for k, v in os.environ.items() if k not in ("RANK", "LOCAL_RANK", "WORLD_SIZE", "MASTER_ADDR", "MASTER_PORT")}
            ref = None
            try:
                out = subprocess.run([sys.executable, os.path.abspath(__file__), "--impl", "reference", "--steps", "2", "--warmup", "1"], env=env,
                                     capture_output=True, text=True, timeout=600)
                for ln in out.stdout.splitlines():
                    if ln.startswith("{"):
                        ref = json.loads(ln)
            except (subprocess.SubprocessError, OSError, ValueError):
                ref = None
            if ref is not None:
                line["cpu_baseline"] = dict(ref["cpu_baseline"])
                line["cpu_baseline"]["sample"] = ("2 full frames (+1 warm-up) of the same workload on the host CPU, separate torch-free process: C++ restatement "
                                                  "of swraster-viewer's rayon+glam path (Rust toolchain unavailable)")
            else:  # the baseline must not cost the bench line: time it here instead
                ncores = os.cpu_count() or 1
                times, ost = cpu_frame_times(scene, cam.abi, 2, 1, ncores)
                cfps = len(times) / sum(times)
                line["cpu_baseline"] = {"value": cfps, "unit": "frames/s", "cores": ncores, "kind": "port",
                                        "sample": "2 full frames (+1 warm-up) of the same workload on the host CPU, inside the bench process (the separate process failed)",
                                        "ms_per_frame": 1e3 / cfps, "ms_clipbin": ost["ms_clipbin"], "ms_raster_shade": ost["ms_raster"], "ms_resolve": ost["ms_resolve"]}
        print(json.dumps(line), flush=True)
    # orderly teardown: torch tensors that were used on the library's stream must die before the stream does
    torch.cuda.synchronize()
    del ev, flush, pix, stream
    buf = None
    import gc
    gc.collect()
    torch.cuda.empty_cache()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    r.close()


def main():
    if "--impl" in sys.argv and "reference" in sys.argv:
        # torchrun exports OMP_NUM_THREADS=1; the CPU arm must use every host core, and libgomp reads this at load time
        os.environ["OMP_NUM_THREADS"] = str(os.cpu_count() or 1)
        os.environ.pop("OMP_PROC_BIND", None)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    args = ap.parse_args()
    import __graft_entry__ as g
    if int(os.environ.get("LOCAL_RANK", "0")) == 0:
        g.build()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
