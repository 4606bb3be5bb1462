import sys; sys.path.insert(0,'.')
import numpy as np
import dsgpuraytracing_b200 as D
from dsgpuraytracing_b200 import scenes as S
V, F = S.torus_knot(); V = V.astype(np.float32).astype(np.float64)
sc = S.cb_mesh_scene(V, F); cam = S.cam_dragon(1920, 1080)
core = D.Core(0); core.set_params(64, 4, 8, 0); core.load(sc, camera=cam); core.set_option("stage_timing", 1)
ref = None
for cm, tm in [(1<<20, 20), (2, 20), (6, 20), (12, 20), (6, 8), (6, 12), (6, 28), (12, 28), (20, 28), (6, 32)]:
    core.set_option("coop_min_pairs", cm); core.set_option("postpone_min_lanes", tm)
    core.render(spp_count=64)
    rgb, st = core.render(spp_count=64)
    if ref is None: ref = rgb
    print("coop_min %7d tri_min %2d  Mrays/s %7.1f  gpu_s %.4f extend %.4f connect %.4f shade %.4f maxdiff %.1e" % (cm, tm, st.segments/st.gpu_seconds/1e6, st.gpu_seconds, st.extend_seconds, st.connect_seconds, st.shade_seconds, np.abs(rgb-ref).max()))
