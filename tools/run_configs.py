#!/usr/bin/env python
"""Runs the five BASELINE.json configurations (parity-test cases, not bench lines) on one GPU and prints one JSON line
per case: segments, s/frame, Mrays/s, node / primitive fetches per segment, algorithmic bytes per segment and GB/s.
  python tools/run_configs.py [--soup-max 16] [--quick]
Scenes: the reference's own .dae scenes come from tests/golden (flattened by the compiled reference); the missing
CBdragon / CBlucy files are replaced by the procedural stand-ins of dsgpuraytracing_b200/scenes.py; config 5 is the
synthetic triangle soup of SURVEY.md 8d."""
import argparse, json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import dsgpuraytracing_b200 as D
from dsgpuraytracing_b200 import scenes as S


def golden(name, W, H):
    z = np.load(os.path.join(ROOT, "tests", "golden", name + ".npz")); g = {k: z[k] for k in z.files}
    cam = g["camera"].copy()
    hf, vf, dist = S.configure_camera(49.13434, 37.849289, W, H) if name != "bunny" else (0, 0, cam[14] * H / cam[13])
    cam[12], cam[13], cam[14] = W, H, dist
    return g, cam


N_GPUS = 1
DEVICE_BUILD = False


def run(tag, sc, cam, spp, nl, depth, reps=2):
    t0 = time.time(); bvh = None if DEVICE_BUILD else D.build_bvh2(sc); t_sah = time.time() - t0
    # N_GPUS > 1: one context over N devices (dsrt_create_multi): sample-split render, partial framebuffers combined by
    # device 0 reading its peers over NVLink inside the resolve kernel
    core = D.Core(0) if N_GPUS == 1 else D.Core(devices=list(range(N_GPUS)))
    core.set_params(spp, nl, depth, 0)
    t0 = time.time(); core.load(sc, camera=cam, bvh=bvh, device_build=DEVICE_BUILD); t_accel = time.time() - t0
    core.set_option("stage_timing", 1)
    best = None; wall = None
    for _ in range(reps):
        w0 = time.perf_counter(); rgb, st = core.render(); w1 = time.perf_counter()
        if best is None or st.gpu_seconds < best.gpu_seconds:
            best = st
        wall = (w1 - w0) if wall is None else min(wall, w1 - w0)
    core.set_option("count_traversal", 1)
    core.set_params(max(1, min(spp, 4)), nl, depth, 0)
    _, sc2 = core.render()
    info = core.accel_info()
    seg = best.segments
    nn = sc2.nodes_visited / sc2.segments; nt = sc2.prims_tested / sc2.segments
    bps = nn * 80 + nt * 48 + 48
    out = {"case": tag, "n_gpus": N_GPUS, "builder": "device" if DEVICE_BUILD else "host_sah", "prims": int(len(sc["prim_type"])), "wide_nodes": info["wide_nodes"], "wide_depth": info["max_depth"],
           "accel_MB": (info["node_bytes"] + info["prim_bytes"]) / 1e6, "sah_build_s": round(t_sah, 3), "flatten_upload_s": round(t_accel, 3),
           "width": int(cam[12]), "height": int(cam[13]), "spp": spp, "light_samples": nl, "max_depth": depth,
           "segments": int(seg), "segments_per_sample": seg / best.camera_samples, "s_per_frame": best.gpu_seconds, "wall_s_incl_reduce_and_readback": wall,
           "Mrays_s": seg / best.gpu_seconds / 1e6, "extend_s": best.extend_seconds, "connect_s": best.connect_seconds,
           "shade_s": best.shade_seconds, "nodes_per_seg": nn, "prims_per_seg": nt, "bytes_per_seg": bps,
           "algorithmic_GB_s": seg * bps / best.gpu_seconds / 1e9, "mean_rgb": [float(x) for x in rgb.mean(axis=(0, 1))]}
    print(json.dumps(out), flush=True)
    core.close()


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--soup-max", type=int, default=8, help="largest soup in Mi triangles")
    ap.add_argument("--soup-min", type=int, default=1, help="smallest soup in Mi triangles")
    ap.add_argument("--quick", action="store_true", help="16 spp instead of the configured 256/512")
    ap.add_argument("--gpus", type=int, default=1, help="GPUs of this box driven by one context (dsrt_create_multi)")
    ap.add_argument("--device-build", action="store_true", help="option device_build (LBVH + collapse on the first GPU) instead of the host SAH builder")
    ap.add_argument("--only", default="", help="comma list of case prefixes to run, e.g. C4,C5")
    a = ap.parse_args()
    q = a.quick
    N_GPUS = a.gpus
    DEVICE_BUILD = a.device_build
    only = [x for x in a.only.split(",") if x]
    want = lambda c: not only or any(c == x for x in only)
    gtag = f"{N_GPUS} GPU" + ("s" if N_GPUS > 1 else "")
    if want("C1"):
        g, cam = golden("CBspheres_lambertian", 480, 360)
        run(f"C1 CBspheres_lambertian 480x360 16spp l4 m5, {gtag}", g, cam, 16, 4, 5)
    if want("C2"):
        sc, cam = S.cbdragon_standin(1920, 1080)
        run(f"C2 CBdragon stand-in (100012-tri mesh) 1080p 256spp l4 m8, {gtag}", sc, cam, 16 if q else 256, 4, 8)
    if want("C3"):
        sc, cam = S.cblucy_standin(1920, 1080)
        run(f"C3 CBlucy stand-in (133796-tri GLASS mesh) 1080p 256spp l4 m8, {gtag}", sc, cam, 16 if q else 256, 4, 8)
        for nm in ("CBgems", "CBcoil", "CBbunny"):
            g, cam = golden(nm, 1920, 1080)
            run(f"C3' {nm} 1080p 256spp l4 m8, {gtag}", g, cam, 16 if q else 256, 4, 8)
    if want("C4"):
        g, cam = golden("bunny", 1920, 1080)
        run(f"C4 bunny (hemisphere light) 1080p 512spp l4 m8, {gtag}", g, cam, 16 if q else 512, 4, 8)
    n = a.soup_min
    while want("C5") and n <= a.soup_max:
        sc, cam = S.triangle_soup(n << 20)
        run(f"C5 triangle soup {n}Mi tris 3840x2160 64spp l1 m8, {gtag}", sc, cam, 8 if q else 64, 1, 8, reps=2)
        del sc
        n *= 2
