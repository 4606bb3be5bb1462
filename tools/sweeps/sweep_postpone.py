import sys; sys.path.insert(0,'.')
import numpy as np
import dsgpuraytracing_b200 as D
from dsgpuraytracing_b200 import scenes as S
V, F = S.torus_knot(); V = V.astype(np.float32).astype(np.float64)
sc = S.cb_mesh_scene(V, F); cam = S.cam_dragon(1920, 1080)
core = D.Core(0); core.set_params(16, 4, 8, 0); core.load(sc, camera=cam); core.set_option("stage_timing", 1)
for tm in [0, 4, 8, 12, 16, 20, 24, 28]:
    core.set_option("postpone_min_lanes", tm)
    core.render()
    rgb, st = core.render()
    core.set_option("count_traversal", 1); _, sc2 = core.render(); core.set_option("count_traversal", 0)
    print("tri_min %2d  Mrays/s %7.1f  extend %.4f connect %.4f shade %.4f   nodes/seg e %.2f c %.2f  prims/seg e %.2f c %.2f" % (
        tm, st.segments/st.gpu_seconds/1e6, st.extend_seconds, st.connect_seconds, st.shade_seconds,
        sc2.extend_nodes/sc2.extend_rays, sc2.connect_nodes/sc2.shadow_rays, sc2.extend_prims/sc2.extend_rays, sc2.connect_prims/sc2.shadow_rays))
