#!/usr/bin/env python
"""bench.py — frames/s and Mtriangles/s of the rasterisation hot path (BASELINE.json metric).

  --config c3 (default)  BASELINE.json configs[2], the configuration the metric is quoted on: procedural ~10 M-triangle
                         instanced scene, 3840x2160, sort-first over N GPUs.
  --config c1|c2|c4|c5   the other BASELINE configs: 100 K-triangle sphere and 1 M-triangle textured terrain
                         at 1920x1080, 50 M micro-triangles at 3840x2160 (sort-first), 8 x 25 M-triangle shards at 7680x4320
                         (sort-last: shard by draw, u64-min key composite over NCCL).
One step = one frame = Renderer.render_scene + blit (set-up, clip, bin, raster, vis-buffer shading, resolve).

  value     device-resident throughput: swr_render + swr_resolve(NULL) timed with CUDA events on the library's stream,
            inputs already in HBM, L2 flushed between frames (256 MiB write).
  e2e       the same frames through the public host API (Renderer.render_scene + blit_to_buffer into pinned host memory):
            draw-list build + H2D of camera/draw table + all kernels + D2H of the RGBA8 frame, wall clock, max over ranks.
            N = 1: pipelined like the reference's App (read-back of frame N overlaps frame N+1). N > 1, sort-first: every
            rank reads ITS rows back over its own PCIe link into one host frame shared by all ranks (POSIX shared memory,
            page-locked by every rank) — no GPU-to-GPU hop on the e2e path.
  N > 1     sort-first (c1-c4): every rank holds the scene and owns a contiguous, cost-balanced range of tile rows; for
            `value` each rank's resolve kernel stores its RGBA8 rows straight into rank 0's frame over NVLink peer memory
            (swr_peer_*), inside the timed region. sort-last (c5): see multigpu.sort_last_frame. "scaling": "strong".
  --impl reference   the reference's CPU path (oracle port: C++ restatement with the reference's parallel structure, all
            host threads) on the same config, rank 0 only. Loads neither torch nor the product libraries.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CONFIGS = {
    "c1": {"W": 1920, "H": 1080, "mode": "sort-first", "builder": "scene_c1_sphere", "args": {},
           "workload": "C1: procedural 99,904-triangle UV sphere (224 x 224), untextured, one default material, reference default camera, 1920x1080"},
    "c2": {"W": 1920, "H": 1080, "mode": "sort-first", "builder": "scene_c2_terrain", "args": {},
           "workload": "C2: procedural 999,698-triangle textured height-field (708 x 708 vertices, 2048^2 sRGB texture, mips 0-5 hit), low-oblique camera, 1920x1080"},
    "c3": {"W": 3840, "H": 2160, "mode": "sort-first", "builder": "scene_c3_instanced", "args": {},
           "workload": "C3: procedural 10M-triangle instanced scene (icosphere 20480 / torus 8192 / box-grid 1200 tris, seeded scatter, scales 0.05-4, camera inside the cloud, heavy clipping), 3840x2160"},
    "c4": {"W": 3840, "H": 2160, "mode": "sort-first", "builder": "scene_c4_micro", "args": {},
           "workload": "C4: procedural 50.0M-triangle dense micro-triangle grid (5001 x 5001 vertices, ~0.3 px^2 per triangle, binning-bound), 3840x2160"},
    "c5": {"W": 7680, "H": 4320, "mode": "sort-last", "builder": "scene_c5_shards", "args": {"nshards": 8},
           "workload": "C5: procedural 200M-triangle scene, 8 shards of 25M-triangle grids at different depths, sharded by primitive with u64-min depth composite, 7680x4320"},
}
CAMERA_FIXTURE = os.path.join(ROOT, "tests", "golden", "bench_cameras.json")


def build_scene(name):
    from swraster_viewer_b200 import scenes
    cfg = CONFIGS[name]
    return getattr(scenes, cfg["builder"])(voxel_dim=128, cube_size=256, **cfg["args"])


def camera_spec(name):
    return build_scene(name)[1]


def fixture_camera(name, spec):
    """The config's swr_camera block from the committed fixture (tests/golden/make_bench_cameras.py): lets the CPU arm run
    without the product libraries. The fixture's spec must be the scene's spec."""
    from swraster_viewer_b200 import abi
    fx = json.load(open(CAMERA_FIXTURE))[name]
    same = (np.allclose(fx["position"], spec.position, rtol=0, atol=0) and np.allclose(fx["look_at"], spec.look_at, rtol=0, atol=0)
            and fx["fov"] == spec.fov and fx["far_plane"] == spec.far_plane)
    if not same:
        raise SystemExit(f"tests/golden/bench_cameras.json is stale for {name}: rerun tests/golden/make_bench_cameras.py")
    return abi.Camera.from_buffer_copy(bytes.fromhex(fx["camera_hex"]))


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

    def mark(self):
        """The timed region starts here: only samples taken from now on are reported."""
        self.first = len(self.lines)

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.06)
        self.proc.terminate()
        self.t.join(timeout=2)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines[getattr(self, "first", 0):]:
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


def metric_name(cfg):
    return f"frames_per_sec_{cfg['W']}x{cfg['H']}"


# ---------------------------------------------------------------------------------------------------------------------
# CPU arm
# ---------------------------------------------------------------------------------------------------------------------
def cpu_frame_times(scene, cam_abi, W, H, steps, warmup, nthreads):
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
    cfg = CONFIGS[args.config]
    W, H = cfg["W"], cfg["H"]
    scene, spec = build_scene(args.config)
    cam_abi = fixture_camera(args.config, spec)
    ncores = os.cpu_count() or 1
    steps = max(1, args.steps)
    warmup = max(0, args.warmup)
    # bounded: keep the whole run within a few minutes whatever K/W the driver passes
    probe, st = cpu_frame_times(scene, cam_abi, W, H, 1, 0, ncores)
    budget = 150.0
    per = probe[0]
    steps_eff = max(1, min(steps, int(budget / per)))
    warm_eff = min(warmup, 1)
    if per * (steps_eff + warm_eff) > budget:  # huge configs (C4 / C5): the probe frame is the warm-up
        warm_eff = 0
    times, st = cpu_frame_times(scene, cam_abi, W, H, steps_eff, warm_eff, ncores)
    ms = 1e3 * sum(times) / len(times)
    fps = 1e3 / ms
    T = st["triangles_submitted"]
    line = {
        "impl": "reference", "metric": metric_name(cfg), "value": fps, "unit": "frames/s", "n_gpus": args.gpus,
        "steps": steps_eff, "warmup": warm_eff, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "mtriangles_per_sec": fps * T / 1e6,
        "config": {"workload": cfg["workload"], "width": W, "height": H, "triangles_submitted": T, "scene_triangles": scene.total_triangles},
        "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": ncores, "kind": "port",
                         "sample": f"{steps_eff} full frames of the same workload (C++ restatement of swraster-viewer's rayon+glam path; Rust toolchain unavailable)",
                         "ms_per_frame": ms, "ms_clipbin": st["ms_clipbin"], "ms_raster_shade": st["ms_raster"], "ms_resolve": st["ms_resolve"]},
        "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------------------------
# CUDA arm
# ---------------------------------------------------------------------------------------------------------------------
def run_gpu(args):
    import torch
    import torch.distributed as dist
    import swraster_viewer_b200 as swr
    from swraster_viewer_b200.multigpu import (tile_row_ranges, balanced_row_ranges, rebalance_row_ranges, PeerAssembly, SharedHostFrame, device_tensor,
                                               sort_last_frame)

    cfg = CONFIGS[args.config]
    W, H = cfg["W"], cfg["H"]
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch with torchrun --nproc-per-node N for --gpus N > 1")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = f"cuda:{local}"
    sort_last = cfg["mode"] == "sort-last" and world > 1

    scene, spec = build_scene(args.config)
    cam = swr.RenderCamera.from_spec(spec, W, H)
    r = swr.Renderer(W, H, device=local)
    tiles_y = r.tiles_y
    ranges = tile_row_ranges(tiles_y, world)
    if world > 1 and not sort_last:
        # probe frame on every rank (full screen, identical everywhere) -> cost-balanced contiguous row bands
        for _ in range(3):
            r.render_scene(scene, cam, shade=False)
        cyc = torch.from_numpy(r.read_tile_costs()[1].astype(np.int64)).to(dev)
        dist.broadcast(cyc, src=0)  # measured cycles differ slightly per rank: everybody uses rank 0's
        cost = cyc.cpu().numpy()
        ranges = balanced_row_ranges(cost, world)
        r.set_tile_rows(*ranges[rank])
        # feedback: the probe only knows raster cycles; two rounds of "render the band, measure all its phases, cut again"
        # (set-up time, untimed) level what set-up and shading add per band
        row_cost = cost.sum(axis=1).astype(np.float64) + 0.9 * cost.mean() * cost.shape[1]
        for _ in range(2):
            acc = 0.0
            for _ in range(4):
                r.render_scene(scene, cam)
                r.synchronize()
                st = r.stats()
                acc += st["ms_setup_bin"] + st["ms_raster"] + st["ms_shade"]
            t = torch.tensor([0.0] * world, dtype=torch.float64, device=dev)
            t[rank] = acc / 4
            dist.all_reduce(t)
            ranges = rebalance_row_ranges(ranges, t.cpu().numpy(), row_cost)
            r.set_tile_rows(*ranges[rank])
    stream = torch.cuda.ExternalStream(r.cuda_stream(), device=local)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # N > 1, sort-first: the frame is assembled in rank 0's pixel buffer by the other ranks' resolve kernels (peer stores over
    # NVLink + a device-side handshake, include/swr.h swr_peer_*): no collective on the data path
    pa = PeerAssembly(r, dst=0) if (world > 1 and not sort_last) else None
    # host frame(s) of the e2e path: pinned; shared between the ranks at N > 1
    if world > 1:
        hosts = [SharedHostFrame(W, H, tag=f"swr_bench_{os.environ.get('MASTER_PORT', '0')}_{k}") for k in range(4)]
        bufs = [swr.RenderBuffer(W, H, pixels=h.pixels) for h in hosts]
    else:
        hosts = None
        bufs = [swr.RenderBuffer(W, H, pinned=True) for _ in range(4)]

    def step_device():
        if sort_last:
            sort_last_frame(r, scene, cam, rank, world, stream)
            return
        r.render_scene(scene, cam)
        if world > 1:
            pa.frame(2.0)
            pa.release()
        else:
            r.resolve_device_only(2.0)

    # nvidia-smi is started BEFORE the warm-up: its start-up (NVML initialisation over every GPU of the box, 0.1-0.3 s) stalls
    # CUDA submissions on all devices for milliseconds, which used to land inside the 15-90 ms timed region (N = 8: 1.4 ms
    # per frame reported for frames whose phases add up to 0.47 ms). Only samples taken after mark() are reported.
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.5)
    # ---- warm-up (also settles buffer growth) -------------------------------------------------------------
    for _ in range(max(3, args.warmup)):
        step_device()
    r.synchronize()
    st0 = r.stats()

    # ---- device-resident timing: CUDA events on the launching stream, L2 flushed between frames ------------
    barrier()
    if rank == 0:
        sampler.mark()
    K = args.steps
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    phase = {"ms_setup_bin": 0.0, "ms_raster": 0.0, "ms_shade": 0.0}
    launches0 = r.launch_count
    for i in range(K):
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
    launches_dev = r.launch_count - launches0
    step_ms = sorted(a.elapsed_time(b) for a, b in ev)
    dev_ms = sum(step_ms)
    # ---- end-to-end timing: public API, host buffers, wall clock ---------------------------------------------
    if sort_last:
        pix_host = torch.empty(W * H, dtype=torch.int32).pin_memory()

        def e2e_loop(n):
            for _ in range(n):
                pix = sort_last_frame(r, scene, cam, rank, world, stream)
                if rank == 0:
                    with torch.cuda.stream(stream):
                        pix_host.copy_(pix, non_blocking=True)
                stream.synchronize()
            barrier()
        e2e_mode = "sort-last composite over NCCL, then D2H of the composited frame on rank 0 (synchronous)"
    else:
        # the pipelined renderer: frames alternate between two contexts of the device (two streams, two sets of per-frame
        # buffers, one scene), so frame N+1's geometry pass overlaps frame N's raster tail and shading
        re = swr.Renderer(W, H, device=local, lanes=2)
        if world > 1:
            re.set_tile_rows(*ranges[rank])

        def e2e_loop(n):
            # pipelined like the reference's App (present N-1 || render N, main.rs:526-597): every step still uploads its draw
            # table and reads its own pixels back into pinned host memory; the read-back of frame N overlaps frame N+1.
            # N > 1: each rank's rows go over its own PCIe link into the host frame all ranks share; the frame is complete
            # when every rank has waited for its copy (hosts[k].arrive / wait_all: flags in the same shared segment).
            def complete(p):
                re.wait_blit(p[0])
                if hosts:
                    hosts[p[1]].arrive(rank, p[2])
                    if rank == 0:
                        hosts[p[1]].wait_all(world, p[2])
            pend = []
            for i in range(n):
                re.render_scene(scene, cam)
                tk = re.blit_to_buffer_async(bufs[i % len(bufs)])
                pend.append((tk, i % len(bufs), i // len(bufs) + 1))
                if len(pend) > 2:  # two frames stay in flight behind the one just enqueued
                    complete(pend.pop(0))
            for p in pend:
                complete(p)
        e2e_mode = ("pipelined, Renderer(lanes=2): frames alternate between two contexts (streams) of the device, read-back of frame N overlaps frames N+1, N+2 (swr_resolve_async)" if world == 1 else
                    "pipelined, Renderer(lanes=2); every rank reads its own rows back over its own PCIe link into one page-locked host frame shared by all ranks")
    for h in hosts or []:
        h.reset()
    barrier()
    e2e_loop(2)
    barrier()  # rank 0 may still be polling the warm-up frames' arrival flags: nobody resets its flag before everybody is through
    for h in hosts or []:
        h.reset()
    barrier()
    rl = r if sort_last else re
    launches1 = rl.launch_count
    t0 = time.perf_counter()
    e2e_loop(K)
    barrier()
    e2e_s = time.perf_counter() - t0
    launches2 = rl.launch_count
    e2e_sync_s = None
    if world == 1:
        # the strictly synchronous form (render, blit, wait) for reference
        t0 = time.perf_counter()
        for _ in range(K):
            r.render_scene(scene, cam)
            r.blit_to_buffer(bufs[0])
        e2e_sync_s = time.perf_counter() - t0
    clocks = sampler.stop() if rank == 0 else None

    per_rank = torch.zeros(world, 3, dtype=torch.float64, device=dev)
    per_rank[rank] = torch.tensor([phase[k] / K for k in ("ms_setup_bin", "ms_raster", "ms_shade")], dtype=torch.float64)
    if world > 1:
        dist.all_reduce(per_rank)
    t = torch.tensor([dev_ms, e2e_s], dtype=torch.float64, device=dev)
    cnt = torch.tensor([st0["tile_refs"], st0["triangles_binned"]], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
    dev_ms, e2e_s = float(t[0]), float(t[1])
    if rank == 0:
        ms = dev_ms / K
        fps = 1e3 / ms
        T = scene_triangles_submitted(st0, sort_last, scene)
        stats_all = dict(st0)
        stats_all["tile_refs"] = int(cnt[0])
        stats_all["triangles_submitted"] = T
        B = algorithmic_bytes(stats_all, W, H)
        peak, peak_src = peaks()
        # dominant kernel of the step on this rank
        dom = max(phase, key=lambda k: phase[k])
        dom_ms = phase[dom] / K
        kname = {"ms_setup_bin": "k_setup+k_clip+k_scan_tiles+k_scatter", "ms_raster": "k_raster_tiles", "ms_shade": "k_shade"}[dom]
        kbytes = {"ms_setup_bin": 12 * st0["triangles_submitted"] + 16 * st0["vertices_submitted"] + 4 * st0["tile_refs"],
                  "ms_raster": 4 * st0["tile_refs"] + 8 * W * H // (1 if sort_last else world),
                  "ms_shade": 4 * W * H // world}[dom]
        achieved = kbytes / (dom_ms * 1e-3) / 1e9
        # DRAM traffic / warp instructions of the dominant kernel from the committed `ncu --set full` capture of this same
        # command (C3, N=1 only): profiles/r02_traffic.json, r02_warp_inst.json
        traffic, issue = None, None
        tj, ij = os.path.join(ROOT, "profiles", "r02_traffic.json"), os.path.join(ROOT, "profiles", "r02_warp_inst.json")
        if world == 1 and args.config == "c3":
            if os.path.exists(tj):
                traffic = json.load(open(tj)).get(kname.split("+")[0])
            if os.path.exists(ij) and clocks and clocks.get("sm_mhz"):
                winst = json.load(open(ij)).get(kname.split("+")[0])
                if winst:
                    ipeak = torch.cuda.get_device_properties(local).multi_processor_count * 4 * clocks["sm_mhz"] * 1e6 / 1e9
                    issue = {"warp_inst_per_launch": winst, "achieved": winst / (dom_ms * 1e-3) / 1e9, "peak": ipeak, "unit": "G warp-inst/s",
                             "frac": winst / (dom_ms * 1e-3) / 1e9 / ipeak, "source": "smsp__inst_executed.sum (profiles/r02_warp_inst.json) / live kernel time"}
        h2d = r.num_draws * (144 + 4) + 4
        if sort_last:
            par = f"sort-last x{world}: draws dealt round-robin, u64-min key all-reduce + barycentric / RGBA8 sum all-reduce over NCCL, every rank shades the pixels whose winner it owns"
        elif world > 1:
            par = f"sort-first x{world}: cost-balanced contiguous tile-row bands {ranges}, per-band draw + cluster culling, peer-store frame assembly on rank 0 (NVLink P2P, no collective)"
        else:
            par = "single GPU"
        line = {
            "metric": metric_name(cfg), "value": fps, "unit": "frames/s", "n_gpus": world, "steps": K, "warmup": max(3, args.warmup),
            "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "mtriangles_per_sec": fps * T / 1e6,
            "config": {"workload": cfg["workload"], "width": W, "height": H, "triangles_submitted": T, "scene_triangles": scene.total_triangles,
                       "vertices_submitted": st0["vertices_submitted"], "tile_refs": int(cnt[0]), "triangles_binned": int(cnt[1]),
                       "l2": "256 MiB device write between timed frames (L2 flush)", "parallelism": par},
            "e2e": {"value": K / e2e_s, "unit": "frames/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": W * H * 4, "ms_per_step": 1e3 * e2e_s / K,
                    "mode": e2e_mode, "synchronous_value": (K / e2e_sync_s) if e2e_sync_s else None},
            # kernels of ours launched on this rank inside the two timed regions, counted by the library (swr_launch_count)
            "gpu_launches": launches_dev + (launches2 - launches1),
            "step_ms": {"min": step_ms[0], "median": step_ms[len(step_ms) // 2], "max": step_ms[-1], "note": "device-resident steps on rank 0 (ms_per_step is the mean, max over ranks)"},
            "per_rank_phase_ms": [[round(float(x), 4) for x in row] for row in per_rank.cpu().tolist()],
            "clocks": clocks,
            "roofline": {"bound": "hbm", "kernel": kname, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                         "peak_source": peak_src, "kernel_ms": dom_ms, "kernel_algorithmic_bytes": kbytes,
                         "frame": {"algorithmic_bytes": B, "achieved": B / (ms * 1e-3) / 1e9, "frac": B / (ms * 1e-3) / 1e9 / peak},
                         "phase_ms": {k: v / K for k, v in phase.items()}, "issue": issue},
        }
        if world == 1 and not args.no_cpu:
            # The CPU port is timed in a fresh interpreter: inside this process (torch and its OpenMP runtime loaded, CUDA
            # context alive) the same code runs slower, which would flatter the GPU arm.
            env = {k: v for k, v in os.environ.items() if k not in ("RANK", "LOCAL_RANK", "WORLD_SIZE", "MASTER_ADDR", "MASTER_PORT")}
            ref = None
            try:
                out = subprocess.run([sys.executable, os.path.abspath(__file__), "--impl", "reference", "--config", args.config, "--steps", "3", "--warmup", "1"],
                                     env=env, capture_output=True, text=True, timeout=900)
                for ln in out.stdout.splitlines():
                    if ln.startswith("{"):
                        ref = json.loads(ln)
            except (subprocess.SubprocessError, OSError, ValueError):
                ref = None
            if ref is not None:
                line["cpu_baseline"] = dict(ref["cpu_baseline"])
                line["cpu_baseline"]["sample"] = (f"{ref['steps']} full frames (+{ref['warmup']} warm-up) of the same workload on the host CPU, separate torch-free process: C++ "
                                                  "restatement of swraster-viewer's rayon+glam path (Rust toolchain unavailable)")
            else:
                line["cpu_baseline"] = {"value": None, "unit": "frames/s", "cores": os.cpu_count() or 1, "kind": "port", "sample": "the CPU leg failed to run"}
        print(json.dumps(line), flush=True)
    # orderly teardown: torch tensors that were used on the library's stream must die before the stream does
    torch.cuda.synchronize()
    del ev, flush, stream
    e2e_loop = step_device = None  # closures hold tensors that live on the library's stream
    if not sort_last:
        re.close()
    if sort_last:
        del pix_host
    bufs = None
    for h in hosts or []:
        h.close(unlink=(rank == 0))
    import gc
    gc.collect()
    torch.cuda.empty_cache()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    r.close()


def scene_triangles_submitted(st0, sort_last, scene):
    """T of the frame: this rank's submitted triangles (sort-first ranks all submit the same draw list before culling);
    for sort-last the shards partition the scene, so the frame's T is the whole scene's."""
    return scene.total_triangles if sort_last else st0["triangles_submitted"]


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
    ap.add_argument("--config", default="c3", choices=sorted(CONFIGS))
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    args = ap.parse_args()
    import __graft_entry__ as g
    if int(os.environ.get("LOCAL_RANK", "0")) == 0:
        g.build(load=args.impl != "reference")  # the CPU arm builds the checker but never maps the product libraries
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
