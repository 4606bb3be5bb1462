#!/usr/bin/env python
"""Host SAH builder vs the device-side builder (option "device_build") on the config-5 soups: preparation time and traversal quality.
  python tools/device_build_bench.py [sizes in Mi triangles ...]       (default 1 8)
One JSON line per (size, builder)."""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import dsgpuraytracing_b200 as D
from dsgpuraytracing_b200 import scenes as S

sizes = [int(x) for x in sys.argv[1:]] or [1, 8]
for n in sizes:
    if n == 0:
        sc, cam = S.load_standin("cbdragon_standin", 1920, 1080); spp, nl, tag = 32, 4, "cbdragon_standin"
    else:
        sc, cam = S.triangle_soup(n << 20); spp, nl, tag = 8, 1, f"soup{n}Mi"
    for builder in ("host_sah", "device_lbvh"):
        core = D.Core(0)
        core.set_params(spp, nl, 8, 0)
        t0 = time.perf_counter()
        if builder == "host_sah":
            bvh = D.build_bvh2(sc); t_sah = time.perf_counter() - t0
            core.load(sc, camera=cam, bvh=bvh)
        else:
            t_sah = 0.0
            core.load(sc, camera=cam, device_build=True)
        t_total = time.perf_counter() - t0
        core.set_option("stage_timing", 1)
        core.render(); rgb, st = core.render()
        core.set_option("count_traversal", 1); core.set_params(1, nl, 8, 0)
        _, sc2 = core.render()
        info = core.accel_info()
        print(json.dumps({"scene": tag, "builder": builder, "prepare_s": round(t_total, 3), "host_sah_s": round(t_sah, 3), "wide_nodes": info["wide_nodes"],
                          "wide_depth": info["max_depth"], "Mrays_s": st.segments / st.gpu_seconds / 1e6, "s_per_frame_at_spp": [spp, st.gpu_seconds],
                          "nodes_per_seg": sc2.nodes_visited / sc2.segments, "prims_per_seg": sc2.prims_tested / sc2.segments,
                          "mean_rgb": [float(x) for x in rgb.mean(axis=(0, 1))]}), flush=True)
        core.close()
