"""Small driver for ncu captures: one wavefront batch (`spp` samples per pixel at 1080p; 4 spp = 8.3 M paths) of the bench workload, twice.
k_trace launches per render: 9 depths x (extend, connect) = 18, so
  ncu --set full --import-source on -k regex:k_trace -s 18 -c 2 ... python profiles/profile_run.py
captures the depth-0 extend and connect launches of the second (warm) render."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dsgpuraytracing_b200 as D
from dsgpuraytracing_b200 import scenes as S

V, F = S.torus_knot(); V = V.astype(np.float32).astype(np.float64)
sc = S.cb_mesh_scene(V, F); cam = S.cam_dragon(1920, 1080)
core = D.Core(0)
args = [x for x in sys.argv[1:] if "=" not in x]
spp = int(args[0]) if args else 2
core.set_params(spp, 4, 8, 0)
core.load(sc, camera=cam)
core.set_option("stage_timing", 1)
for kv in sys.argv[1:]:          # e.g. refill_busy_lanes=0 postpone_min_lanes=0 coop_min_pairs=1000000 (no compaction)
    if "=" in kv:
        k, v = kv.split("="); core.set_option(k, int(v))
for i in range(2):
    rgb, st = core.render()
print("segments %d  gpu_s %.4f  Mrays/s %.1f  extend %.4f connect %.4f shade %.4f" % (
    st.segments, st.gpu_seconds, st.segments / st.gpu_seconds / 1e6, st.extend_seconds, st.connect_seconds, st.shade_seconds))
