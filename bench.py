#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200-native path-tracing core.

Metric (BASELINE.json): Mrays/s counted in path SEGMENTS (every closest-hit "extend" query + every any-hit shadow
query) and s/frame, on configs[1]: CBdragon 1920x1080, 256 spp, 4 light samples, depth 8, cam_dragon.info.
CBdragon.dae is not in the reference checkout (.MISSING_LARGE_BLOBS), so the workload is the documented stand-in:
the reference's Cornell box + a procedural 100 012-triangle closed mesh (dsgpuraytracing_b200/scenes.py).

  python bench.py [--gpus N --steps K --warmup W]           our arm (N>1: launched by torchrun, one rank per GPU)
  python bench.py --impl reference [...]                    the reference's own CPU path tracer on the host cores

One "step" = one full frame.  `value` is timed with CUDA events on the render stream with all inputs resident in HBM
(max over ranks, NCCL reduce of the partial framebuffers included); `e2e` repeats the measurement through the C ABI
with host buffers: scene/BVH upload (dsrt_set_scene, dsrt_set_bvh, dsrt_build_accel), render, frame read-back.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "Mrays/s (path segments), CBdragon 1080p 256spp"
UNIT = "Mrays/s"
L2_BYTES = 126e6          # B200 L2 capacity (below it the traversal roofline is the L2 read bandwidth, above it HBM)

# --workload: c2 is the configuration the metric is quoted on (BASELINE.json configs[1]); the others reproduce the figures of
# profiles/r2_configs.md through the same timed path.  (name: (label, width, height, spp, light samples, depth))
WORKLOADS = {
    "c2": ("CBdragon stand-in: Cornell box + procedural 100012-triangle closed mesh (CBdragon.dae missing from the reference "
           "checkout), cam_dragon.info", 1920, 1080, 256, 4, 8),
    "c3": ("CBlucy stand-in: Cornell box + procedural 133796-triangle GLASS mesh (CBlucy.dae missing), cam_dragon.info", 1920, 1080, 256, 4, 8),
    "c4": ("bunny.dae (hemisphere sky light), default camera", 1920, 1080, 512, 4, 8),
}
for _n in (1, 2, 4, 8, 16, 32, 64):
    WORKLOADS["soup%d" % _n] = ("synthetic triangle soup, %d Mi triangles, hemisphere light" % _n, 3840, 2160, 64, 1, 8)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--width", type=int, default=0)
    ap.add_argument("--height", type=int, default=0)
    ap.add_argument("--spp", type=int, default=0)
    ap.add_argument("--light-samples", type=int, default=0)
    ap.add_argument("--depth", type=int, default=-1)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-procs", type=int, default=0, help="host processes for the CPU arm (0 = all cores)")
    ap.add_argument("--split", default="samples", choices=["samples", "tiles"],
                    help="N > 1: each GPU takes a disjoint slice of the samples of every pixel (default) or a horizontal band of the image")
    ap.add_argument("--device-build", action="store_true", help="option device_build: LBVH + collapse on the GPU instead of the host SAH builder")
    ap.add_argument("--skip-null-shadow", action="store_true", help="option skip_null_shadow (default off: the reference traces them)")
    a = ap.parse_args()
    label, w, h, spp, nl, depth = WORKLOADS[a.workload]
    a.label = label
    a.width = a.width or w; a.height = a.height or h; a.spp = a.spp or spp
    a.light_samples = a.light_samples or nl; a.depth = depth if a.depth < 0 else a.depth
    global METRIC
    if a.workload != "c2":
        METRIC = "Mrays/s (path segments), workload %s %dx%d %dspp" % (a.workload, a.width, a.height, a.spp)
    return a


def workload_config(a, n_gpus, triangles=None):
    return {"workload": a.label, "name": a.workload,
            "width": a.width, "height": a.height, "spp": a.spp, "light_samples": a.light_samples, "max_depth": a.depth,
            "triangles": triangles, "builder": "device LBVH + collapse (option device_build)" if a.device_build else "host SAH (reference topology) + DP collapse",
            "parallelism": (f"{'tile' if a.split == 'tiles' else 'sample'}-split x{n_gpus} + 1 NCCL reduce") if n_gpus > 1 else "single GPU",
            "l2": "explicit 256 MiB L2 flush between steps; the per-batch wavefront state (GBs, larger than L2) streams through "
                  "L2; the wide BVH + primitive records are re-read within a step"}


def stage_workload(a, td, need_arrays):
    """Files / arrays of the workload.  c2 / c3 / c4 are .dae scenes: BOTH arms start from the same file (the reference through
    its ColladaParser, the product through csrc/host/scene_loader.cpp, proven bit-identical by the tests); the soups are flat arrays."""
    from dsgpuraytracing_b200 import scenes as S
    w = {"dae": None, "cam": None, "sc": None, "camera": None}
    if a.workload in ("c2", "c3"):
        w["dae"], w["cam"] = S.write_standin("cbdragon_standin" if a.workload == "c2" else "cblucy_standin", td, a.width, a.height)
    elif a.workload == "c4":
        w["dae"] = os.path.join(ROOT, "oracle", "_ref", "scenes", "bunny.dae")
    if need_arrays:
        if w["dae"]:
            import dsgpuraytracing_b200 as D
            w["sc"], w["camera"] = D.load_dae(w["dae"], a.width, a.height, w["cam"])
        else:
            w["sc"], w["camera"] = S.triangle_soup(int(a.workload[4:]) << 20, W=a.width, H=a.height)
    return w


# ---------------------------------------------------------------------------------------------- clocks sampler
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows = []
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--id={index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                       "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.p = None

    def _read(self):
        for line in self.p.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if not self.p:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.p.terminate()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ts, line in self.rows:
            f = [x.strip() for x in line.split(",")]
            if len(f) < 7 or not (t0 <= ts <= t1 + 0.2):
                continue
            try:
                sm.append(float(f[0])); mx = float(f[1])
            except ValueError:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


# ---------------------------------------------------------------------------------------------- CPU arm
def cpu_step(a, procs, seed0, w, spp_each=1, keep_frames=False):
    """One bounded CPU sample: `procs` independent single-threaded processes (the reference's own -t N mode
    anti-scales because every thread shares glibc rand()), each rendering `spp_each` spp of the full
    frame.  Returns (segments, seconds = slowest process's render time, kind, mean frame or None)."""
    from oracle import oracle as O
    if O.have_reference() and w["dae"]:
        from concurrent.futures import ThreadPoolExecutor
        def one(k):
            r = O.run_reference(w["dae"], a.width, a.height, cam=w["cam"], spp=spp_each, nl=a.light_samples, depth=a.depth,
                                seed=seed0 + k, render=True)
            return r["counters"], (r["rgb"].astype(np.float64) if keep_frames else None)
        with ThreadPoolExecutor(procs) as ex:
            rs = list(ex.map(one, range(procs)))
        cs = [r[0] for r in rs]
        frame = np.mean([r[1] for r in rs], axis=0) if keep_frames else None
        return float(sum(c[0] + c[1] for c in cs)), float(max(c[2] for c in cs)), "reference", frame
    # no .dae (soups) or no compiled reference: the plain-C port of the same algorithm (oracle/pt_oracle.c), one process per core
    from concurrent.futures import ProcessPoolExecutor
    with ProcessPoolExecutor(procs) as ex:
        cs = list(ex.map(_port_part, [(a.workload, a.width, a.height, a.light_samples, a.depth, seed0 + k, spp_each) for k in range(procs)]))
    return float(sum(c[0] for c in cs)), float(max(c[1] for c in cs)), "port", None


def _port_part(args):
    name, W, H, nl, depth, k, spp_each = args
    from oracle import oracle as O
    from dsgpuraytracing_b200 import scenes as S
    if name.startswith("soup"):
        sc, cam = S.triangle_soup(int(name[4:]) << 20, W=W, H=H)
    elif name == "c3":
        sc, cam = S.cblucy_standin(W, H)
    else:
        sc, cam = S.cbdragon_standin(W, H)
    s = O.Scene(dict(sc, camera=cam))
    s.build_bvh()
    t = time.time()
    _, cnt = s.render(W, H, 256, nl, depth, rng="philox", seed=0, spp_begin=k * spp_each, spp_count=spp_each)
    return float(cnt[0] + cnt[1]), time.time() - t


def run_reference_arm(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    procs = a.cpu_procs or os.cpu_count() or 1
    with tempfile.TemporaryDirectory() as td:
        w = stage_workload(a, td, need_arrays=False)
        for k in range(a.warmup):
            cpu_step(a, procs, 1000 + 97 * k, w)
        segs = 0.0; secs = 0.0; kind = "port"
        for k in range(a.steps):
            sg, t, kind, _ = cpu_step(a, procs, 5000 + 97 * k, w)
            segs += sg; secs += t
    value = segs / secs / 1e6
    sample = f"each step = {procs} single-threaded processes x 1 spp of the full {a.width}x{a.height} frame (different seeds)"
    seg_per_frame = segs / (a.steps * procs) * a.spp
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps,
            "warmup": a.warmup, "ms_per_step": secs / a.steps * 1e3, "frame_s_extrapolated": seg_per_frame / (value * 1e6),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(a, 1),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": procs, "kind": kind, "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------- our arm
def run_ours(a):
    import torch
    import torch.distributed as dist
    import dsgpuraytracing_b200 as D

    world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the product has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    if a.spp % world:
        raise SystemExit("--spp must be divisible by the number of GPUs")

    # ---- workload (synthetic, host side): the .dae both arms start from -> product loader -> reference-identical SAH BVH
    tdir = tempfile.TemporaryDirectory()
    w = stage_workload(a, tdir.name, need_arrays=True)
    sc, cam = w["sc"], w["camera"]
    n_tris = int(len(sc["prim_type"]))
    t_sah0 = time.perf_counter()
    bvh = None if a.device_build else D.build_bvh2(sc)
    sah_seconds = time.perf_counter() - t_sah0
    core = D.Core(local)
    core.set_params(a.spp, a.light_samples, a.depth, 0)
    core.load(sc, camera=cam, bvh=bvh, device_build=a.device_build)
    if a.skip_null_shadow:
        core.set_option("skip_null_shadow", 1)
    info = core.accel_info()
    npix = a.width * a.height
    accum = torch.zeros(npix * 3, dtype=torch.float32, device=dev)
    rgb = torch.zeros(npix * 3, dtype=torch.float32, device=dev)
    rgba = torch.zeros(npix, dtype=torch.int32, device=dev)
    flush = torch.zeros(64 << 20, dtype=torch.float32, device=dev)      # 256 MiB > L2
    host_rgb = torch.zeros(npix * 3, dtype=torch.float32).pin_memory()
    # everything (torch ops, our kernels, the NCCL reduce, the timing events) is ordered on ONE explicit stream; torch's
    # default stream has handle 0, which the C ABI reads as "use the context's own stream"
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    assert stream.cuda_stream != 0
    torch.cuda.synchronize(dev)
    spp_local = a.spp // world
    ev = lambda: torch.cuda.Event(enable_timing=True)
    # how this rank's share of the frame is expressed to dsrt_render_device: (first sample, count, stride)
    share = (rank, spp_local, world)
    if world > 1 and a.split == "tiles":          # tile partitioning: a band of rows, all samples; the same sum-reduce combines the bands
        y0 = a.height * rank // world; y1 = a.height * (rank + 1) // world
        core.set_window(0, y0, a.width, y1 - y0)
        share = (0, a.spp, 1)

    def step(i, marks=None):
        flush.fill_(float(i))
        accum.zero_()
        if marks: marks[0].record(stream)
        core.render_device(accum.data_ptr(), *share, stream=stream.cuda_stream)
        if marks: marks[1].record(stream)
        if world > 1:
            dist.reduce(accum, dst=0, op=dist.ReduceOp.SUM)
        if marks: marks[2].record(stream)
        if rank == 0:
            core.resolve_device(accum.data_ptr(), rgb.data_ptr(), rgba.data_ptr(), stream=stream.cuda_stream)
        if marks: marks[3].record(stream)

    def fence():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize(dev)

    core.set_option("stage_timing", 1)
    for i in range(a.warmup):
        step(i)
    fence()
    sampler = ClockSampler(local) if rank == 0 else None
    e0, e1 = ev(), ev()
    t0 = time.time()
    marks = [[ev() for _ in range(4)] for _ in range(a.steps)]
    e0.record(stream)
    for i in range(a.steps):
        step(i, marks[i])
    e1.record(stream)
    fence()
    t1 = time.time()
    ms = e0.elapsed_time(e1)
    step_ms = [(e0 if i == 0 else marks[i - 1][3]).elapsed_time(marks[i][3]) for i in range(a.steps)]   # this rank, per step
    render_ms = sum(m[0].elapsed_time(m[1]) for m in marks) / a.steps          # this rank's own kernels
    reduce_wait_ms = sum(m[1].elapsed_time(m[2]) for m in marks) / a.steps     # waiting for the slowest rank + the NCCL reduce
    resolve_ms = sum(m[2].elapsed_time(m[3]) for m in marks) / a.steps
    st = core.collect_stats()                      # counters / per-stage events of the last step on this rank
    frame_dev = rgb.cpu().numpy().reshape(a.height, a.width, 3).copy() if rank == 0 else None
    clocks = sampler.stop(t0, t1) if sampler else None
    traced = float(st.segments) - float(st.null_shadow_rays)
    t = torch.tensor([ms, traced, float(st.kernel_launches), float(st.null_shadow_rays)], dtype=torch.float64, device=dev)
    tmax = t.clone()
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    ms_max = float(tmax[0]); seg_step = float(t[1]); launches = int(t[2]); nulls = float(t[3])
    value = seg_step * a.steps / (ms_max * 1e-3) / 1e6
    rr = torch.tensor([render_ms], dtype=torch.float64, device=dev)
    rlist = [torch.zeros_like(rr) for _ in range(world)]
    if world > 1:
        dist.all_gather(rlist, rr)
    else:
        rlist = [rr]
    render_ms_ranks = [float(x[0]) for x in rlist]
    tail_ms = max(render_ms_ranks) - render_ms_ranks[0]        # how long rank 0 waits for the slowest rank's kernels
    multi = {"render_ms_per_rank": render_ms_ranks, "tail_ms": tail_ms, "reduce_ms": max(reduce_wait_ms - tail_ms, 0.0),
             "reduce_wait_ms_rank0": reduce_wait_ms, "resolve_ms": resolve_ms}

    # ---- e2e: the reference-facing call sequence with host buffers, every step:
    #   H2D  dsrt_upload_accel (flattened scene + BVH, what CUDAPathTracer::init cudaMemcpy's), camera, parameters
    #   run  dsrt_render
    #   D2H  the frame into a host buffer
    # Scene PREPARATION on the host (SAH build, collapse to the wide BVH, record flattening) is one-off per scene, as
    # in the reference (PathTracer::build_accel runs at set_scene time), and is reported separately below.
    t_prep0 = time.perf_counter()
    core.set_scene(sc)
    if bvh is not None:
        core.set_bvh(bvh)
    core.build_accel()
    scene_prepare_s = time.perf_counter() - t_prep0
    h2d = core.accel_bytes() + 200
    d2h = npix * 12
    def e2e_step():
        core.upload_accel(); core.set_camera(cam); core.set_params(a.spp, a.light_samples, a.depth, 0)
        if world == 1:
            core.render(out=host_rgb.numpy().reshape(a.height, a.width, 3))                   # dsrt_render: host frame out
        else:
            accum.zero_()
            if a.split == "tiles":
                core.set_window(0, a.height * rank // world, a.width, a.height * (rank + 1) // world - a.height * rank // world)
            core.render_device(accum.data_ptr(), *share, stream=stream.cuda_stream)
            dist.reduce(accum, dst=0, op=dist.ReduceOp.SUM)
            if rank == 0:
                core.resolve_device(accum.data_ptr(), rgb.data_ptr(), 0, stream=stream.cuda_stream)
                host_rgb.copy_(rgb, non_blocking=True)
            torch.cuda.synchronize(dev)
    core.set_option("stage_timing", 0)
    e2e_step(); fence()
    w0 = time.perf_counter()
    for i in range(a.steps):
        e2e_step()
    fence()
    w1 = time.perf_counter()
    tw = torch.tensor([w1 - w0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tw, op=dist.ReduceOp.MAX)
    e2e_value = seg_step * a.steps / float(tw[0]) / 1e6
    frame_check = None
    if rank == 0:      # the device-resident path and the host-buffer path render the same Philox samples
        fh = host_rgb.numpy().reshape(a.height, a.width, 3)
        denom = max(float(np.abs(fh).mean()), 1e-12)
        frame_check = {"mean_rgb": [float(x) for x in frame_dev.mean(axis=(0, 1))],
                       "device_vs_host_path_mean_abs_diff_rel": float(np.abs(frame_dev - fh).mean() / denom),
                       "finite": bool(np.isfinite(frame_dev).all())}
        if not frame_check["finite"] or frame_check["device_vs_host_path_mean_abs_diff_rel"] > 1e-3:
            raise SystemExit(f"bench.py: frame check failed: {frame_check}")

    # ---- roofline of the dominant kernel: algorithmic bytes / CUDA-event time of its launches (last timed step)
    core.set_option("count_traversal", 1)
    accum.zero_()
    stc = core.render_device(accum.data_ptr(), *share, stream=stream.cuda_stream, collect=True)
    core.set_option("count_traversal", 0)
    kinds = {"extend (closest-hit traversal, k_trace<false>)": (st.extend_seconds, stc.extend_nodes, stc.extend_prims, st.extend_rays),
             "connect (any-hit traversal, k_trace<true>)": (st.connect_seconds, stc.connect_nodes, stc.connect_prims, st.shadow_rays - st.null_shadow_rays)}
    dom = max(kinds, key=lambda k: kinds[k][0])
    sec, nn, nt, nr = kinds[dom]
    algo_bytes = nn * 80 + nt * 48 + nr * 48      # 80 B of information per node visit (the record is padded to 96 B for 256-bit loads: not counted)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    hbm_src = "MEASURED_PEAKS.json hbm_gbs (of measured)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (of fallback)"
    achieved = algo_bytes / sec / 1e9 if sec > 0 else 0.0
    # while the wide BVH + primitive records fit in L2 the traversal roofline is the L2 -> SM read bandwidth
    # (no driver-written L2 peak exists: measured live with a read-only sweep of a 32 MiB buffer by every SM); above it, HBM
    l2_gbs = core.measure_read_bandwidth(32 << 20, 40)
    hbm_read_gbs = core.measure_read_bandwidth(1 << 30, 3)
    accel_bytes = info["node_bytes"] + info["prim_bytes"]
    l2_regime = accel_bytes < L2_BYTES
    traffic = None; traffic_detail = None
    try:
        traffic_detail = json.load(open(os.path.join(ROOT, "profiles", "dominant_kernel_traffic.json")))
        if traffic_detail.get("workload", "c2") == a.workload:
            traffic = traffic_detail["traffic_bytes_per_launch"]
    except Exception:
        pass
    peak = l2_gbs if l2_regime else hbm_peak
    roofline = {"bound": "l2" if l2_regime else "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak if peak > 0 else None,
                "peak_source": ("live probe dsrt_measure_read_bandwidth: all SMs sweep one 32 MiB (L2-resident) buffer 40x with 128-bit "
                                "ld.global.cg (no driver-written L2 peak exists)") if l2_regime else hbm_src,
                "traffic": traffic, "traffic_detail": traffic_detail, "kernel": dom, "kernel_seconds_per_step": sec,
                "kernel_share_of_step": sec / (ms_max * 1e-3 / a.steps), "bytes_per_segment": algo_bytes / max(nr, 1),
                "nodes_per_segment": nn / max(nr, 1), "prims_per_segment": nt / max(nr, 1),
                "accel_bytes": accel_bytes, "regime": "wide BVH + primitive records %.1f MB %s the %.0f MB L2" % (accel_bytes / 1e6, "<" if l2_regime else ">", L2_BYTES / 1e6),
                "other_roofline": {"bound": "hbm" if l2_regime else "l2", "peak": hbm_peak if l2_regime else l2_gbs,
                                   "frac": achieved / (hbm_peak if l2_regime else l2_gbs), "peak_source": hbm_src if l2_regime else "live L2 probe",
                                   "hbm_read_gbs_same_probe": hbm_read_gbs},
                "note": "the kernel is instruction-issue / ALU-pipe bound (profiles/r2_ktrace_summary.md); DRAM only sees the ray queues; bytes_per_segment "
                        "comes from this run's exact node / primitive fetch counters, so a tree that needs fewer fetches lowers 'achieved' together with the time "
                        "(bench scene: 666 B/segment before the light-aligned grid and the coplanar-mate drop, DESIGN.md section 3 and 4.1)"}

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
                "ms_per_step": ms_max / a.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic", "config": workload_config(a, world, n_tris),
                "segments_per_step": seg_step, "s_per_frame": ms_max / a.steps * 1e-3,
                "null_shadow_rays_skipped_per_step": nulls,
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                        "s_per_frame": float(tw[0]) / a.steps, "scene_prepare_seconds_once": scene_prepare_s, "sah_build_seconds_once": sah_seconds},
                "gpu_launches": launches * a.steps, "clocks": clocks, "roofline": roofline, "frame_check": frame_check,
                "step_ms_rank0": step_ms, "multi_gpu": multi,
                "stage_seconds_per_step": {"extend": st.extend_seconds, "connect": st.connect_seconds, "generate+shade": st.shade_seconds},
                "accel": info}
        if world == 1 and not a.no_cpu_baseline and not (a.workload.startswith("soup") and int(a.workload[4:]) > 2):
            procs = a.cpu_procs or os.cpu_count() or 1
            sg, tsec, kind, cpu_frame = cpu_step(a, procs, 4242, w, keep_frames=True)
            line["cpu_baseline"] = {"value": sg / tsec / 1e6, "unit": UNIT, "cores": procs, "kind": kind,
                                    "sample": f"{procs} single-threaded processes x 1 spp of the full {a.width}x{a.height} frame"}
            if cpu_frame is not None:
                # cross-arm frame check: the GPU frame against the mean of the CPU leg's frames (the reference's own renderer on
                # the same .dae), 20x20 block means.  The CPU mean carries `procs` spp of Monte-Carlo noise, the bound allows for it.
                k = 20
                bm = lambda x: x[:a.height // k * k, :a.width // k * k].reshape(a.height // k, k, a.width // k, k, 3).mean(axis=(1, 3))
                g, c = bm(frame_dev.astype(np.float64)), bm(cpu_frame)
                rel_block = float(np.sqrt(((g - c) ** 2).mean()) / c.mean())
                rel_mean = float(abs(g.mean() - c.mean()) / c.mean())
                line["frame_check"]["vs_cpu_reference_frames"] = {"cpu_spp": procs, "block20_rel_rmse": rel_block, "mean_rel_diff": rel_mean}
                if rel_mean > 0.03 or rel_block > 0.5:
                    raise SystemExit(f"bench.py: GPU frame disagrees with the CPU reference frames: {line['frame_check']}")
        print(json.dumps(line), flush=True)
    core.close()
    tdir.cleanup()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    args = parse()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_ours(args)
