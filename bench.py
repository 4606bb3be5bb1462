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


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--width", type=int, default=1920)
    ap.add_argument("--height", type=int, default=1080)
    ap.add_argument("--spp", type=int, default=256)
    ap.add_argument("--light-samples", type=int, default=4)
    ap.add_argument("--depth", type=int, default=8)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-procs", type=int, default=0, help="host processes for the CPU arm (0 = all cores)")
    return ap.parse_args()


def workload_config(a, n_gpus):
    return {"workload": "CBdragon stand-in: Cornell box + procedural 100012-triangle closed mesh (CBdragon.dae missing "
                        "from the reference checkout), cam_dragon.info",
            "width": a.width, "height": a.height, "spp": a.spp, "light_samples": a.light_samples, "max_depth": a.depth,
            "triangles": 100024, "parallelism": f"sample-split x{n_gpus} + 1 NCCL reduce" if n_gpus > 1 else "single GPU",
            "l2": "explicit 256 MiB L2 flush between steps; per-batch wavefront state (~4.4 GB, larger than L2) streams through "
                  "L2, the 5.9 MB wide BVH + primitive records are re-read within a step (L2-resident by design)"}


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
def stage_reference_scene(a, td):
    from dsgpuraytracing_b200 import scenes as S
    V, F = S.torus_knot()
    V = V.astype(np.float32).astype(np.float64)
    dae = os.path.join(td, "cbdragon_standin.dae"); cam = os.path.join(td, "cam_dragon.info")
    S.write_cb_mesh_dae(dae, V, F)
    S.write_cam_info(cam, S.cam_dragon(a.width, a.height))
    return dae, cam


def cpu_step(a, procs, seed0, dae=None, cam=None, spp_each=1):
    """One bounded CPU sample: `procs` independent single-threaded processes (the reference's own -t N mode
    anti-scales because every thread shares glibc rand()), each rendering `spp_each` spp of the full
    frame.  Returns (segments, seconds = slowest process's render time, kind)."""
    from oracle import oracle as O
    if O.have_reference() and dae:
        from concurrent.futures import ThreadPoolExecutor
        def one(k):
            return O.run_reference(dae, a.width, a.height, cam=cam, spp=spp_each, nl=a.light_samples, depth=a.depth,
                                   seed=seed0 + k, render=True)["counters"]
        with ThreadPoolExecutor(procs) as ex:
            cs = list(ex.map(one, range(procs)))
        return float(sum(c[0] + c[1] for c in cs)), float(max(c[2] for c in cs)), "reference"
    # fallback: the plain-C port of the same algorithm (oracle/pt_oracle.c), one process per core
    from concurrent.futures import ProcessPoolExecutor
    with ProcessPoolExecutor(procs) as ex:
        cs = list(ex.map(_port_part, [(a.width, a.height, a.light_samples, a.depth, seed0 + k, spp_each) for k in range(procs)]))
    return float(sum(c[0] for c in cs)), float(max(c[1] for c in cs)), "port"


def _port_part(args):
    W, H, nl, depth, k, spp_each = args
    from oracle import oracle as O
    from dsgpuraytracing_b200 import scenes as S
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
        dae, cam = stage_reference_scene(a, td)
        for w in range(a.warmup):
            cpu_step(a, procs, 1000 + 97 * w, dae, cam)
        segs = 0.0; secs = 0.0; kind = "port"
        for k in range(a.steps):
            s, t, kind = cpu_step(a, procs, 5000 + 97 * k, dae, cam)
            segs += s; secs += t
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
    from dsgpuraytracing_b200 import scenes as S

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

    # ---- workload (synthetic, host side): scene arrays + reference-identical SAH BVH
    V, F = S.torus_knot()
    V = V.astype(np.float32).astype(np.float64)
    sc = S.cb_mesh_scene(V, F); cam = S.cam_dragon(a.width, a.height)
    bvh = D.build_bvh2(sc)
    core = D.Core(local)
    core.set_params(a.spp, a.light_samples, a.depth, 0)
    core.load(sc, camera=cam, bvh=bvh)
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

    def step(i):
        flush.fill_(float(i))
        accum.zero_()
        core.render_device(accum.data_ptr(), rank, spp_local, world, stream=stream.cuda_stream)
        if world > 1:
            dist.reduce(accum, dst=0, op=dist.ReduceOp.SUM)
        if rank == 0:
            core.resolve_device(accum.data_ptr(), rgb.data_ptr(), rgba.data_ptr(), stream=stream.cuda_stream)

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
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.time()
    marks = [torch.cuda.Event(enable_timing=True) for _ in range(a.steps)]
    e0.record(stream)
    for i in range(a.steps):
        step(i)
        marks[i].record(stream)
    e1.record(stream)
    fence()
    t1 = time.time()
    ms = e0.elapsed_time(e1)
    step_ms = [(e0 if i == 0 else marks[i - 1]).elapsed_time(marks[i]) for i in range(a.steps)]   # this rank, per step
    st = core.collect_stats()                      # counters / per-stage events of the last step on this rank
    frame_dev = rgb.cpu().numpy().reshape(a.height, a.width, 3).copy() if rank == 0 else None
    clocks = sampler.stop(t0, t1) if sampler else None
    t = torch.tensor([ms, float(st.segments), float(st.kernel_launches)], dtype=torch.float64, device=dev)
    tmax = t.clone()
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    ms_max = float(tmax[0]); seg_step = float(t[1]); launches = int(t[2])
    value = seg_step * a.steps / (ms_max * 1e-3) / 1e6

    # ---- e2e: the reference-facing call sequence with host buffers, every step:
    #   H2D  dsrt_upload_accel (flattened scene + BVH, what CUDAPathTracer::init cudaMemcpy's), camera, parameters
    #   run  dsrt_render
    #   D2H  the frame into a host buffer
    # Scene PREPARATION on the host (SAH build, collapse to the wide BVH, record flattening) is one-off per scene, as
    # in the reference (PathTracer::build_accel runs at set_scene time), and is reported separately below.
    t_prep0 = time.perf_counter()
    core.set_scene(sc); core.set_bvh(bvh); core.build_accel()
    scene_prepare_s = time.perf_counter() - t_prep0
    h2d = core.accel_bytes() + 200
    d2h = npix * 12
    def e2e_step():
        core.upload_accel(); core.set_camera(cam); core.set_params(a.spp, a.light_samples, a.depth, 0)
        if world == 1:
            core.render(out=host_rgb.numpy().reshape(a.height, a.width, 3))                   # dsrt_render: host frame out
        else:
            accum.zero_()
            core.render_device(accum.data_ptr(), rank, spp_local, world, stream=stream.cuda_stream)
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
    stc = core.render_device(accum.data_ptr(), rank, spp_local, world, stream=stream.cuda_stream, collect=True)
    core.set_option("count_traversal", 0)
    kinds = {"extend (closest-hit traversal, k_trace<false>)": (st.extend_seconds, stc.extend_nodes, stc.extend_prims, st.extend_rays),
             "connect (any-hit traversal, k_trace<true>)": (st.connect_seconds, stc.connect_nodes, stc.connect_prims, st.shadow_rays)}
    dom = max(kinds, key=lambda k: kinds[k][0])
    sec, nn, nt, nr = kinds[dom]
    algo_bytes = nn * 80 + nt * 48 + nr * 48
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    achieved = algo_bytes / sec / 1e9 if sec > 0 else 0.0
    # L2-resident regime: the denominator is the L2 -> SM read bandwidth, measured live with a read-only
    # sweep of a 32 MiB buffer by every SM; the same probe over 1 GiB gives the HBM read bandwidth for comparison
    l2_gbs = core.measure_read_bandwidth(32 << 20, 40)
    hbm_read_gbs = core.measure_read_bandwidth(1 << 30, 3)
    traffic = None; traffic_detail = None
    try:
        traffic_detail = json.load(open(os.path.join(ROOT, "profiles", "dominant_kernel_traffic.json")))
        traffic = traffic_detail["traffic_bytes_per_launch"]
    except Exception:
        pass
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (of fallback)",
                "traffic": traffic, "traffic_detail": traffic_detail, "kernel": dom, "kernel_seconds_per_step": sec,
                "kernel_share_of_step": sec / (ms_max * 1e-3 / a.steps), "bytes_per_segment": algo_bytes / max(nr, 1),
                "nodes_per_segment": nn / max(nr, 1), "prims_per_segment": nt / max(nr, 1),
                "l2": {"read_gbs_measured": l2_gbs, "frac": achieved / l2_gbs if l2_gbs > 0 else None,
                       "how": "dsrt_measure_read_bandwidth: all SMs sweep one 32 MiB buffer 40x with 128-bit ld.global.cg",
                       "hbm_read_gbs_same_probe": hbm_read_gbs},
                "note": "working set (wide BVH + primitive records = %.1f MB) is L2-resident, so the HBM roofline is an upper "
                        "bound the kernel is not expected to approach; the kernel is latency/issue bound" % ((info["node_bytes"] + info["prim_bytes"]) / 1e6)}

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
                "ms_per_step": ms_max / a.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic", "config": workload_config(a, world),
                "segments_per_step": seg_step, "s_per_frame": ms_max / a.steps * 1e-3,
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                        "s_per_frame": float(tw[0]) / a.steps, "scene_prepare_seconds_once": scene_prepare_s},
                "gpu_launches": launches * a.steps, "clocks": clocks, "roofline": roofline, "frame_check": frame_check,
                "step_ms_rank0": step_ms,
                "stage_seconds_per_step": {"extend": st.extend_seconds, "connect": st.connect_seconds, "generate+shade": st.shade_seconds},
                "accel": info}
        if world == 1 and not a.no_cpu_baseline:
            procs = a.cpu_procs or os.cpu_count() or 1
            with tempfile.TemporaryDirectory() as td:
                dae, camf = stage_reference_scene(a, td)
                s, tsec, kind = cpu_step(a, procs, 4242, dae, camf)
            line["cpu_baseline"] = {"value": s / tsec / 1e6, "unit": UNIT, "cores": procs, "kind": kind,
                                    "sample": f"{procs} single-threaded processes x 1 spp of the full {a.width}x{a.height} frame"}
        print(json.dumps(line), flush=True)
    core.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    args = parse()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_ours(args)
