import sys; sys.path.insert(0,'.')
import numpy as np
import dsgpuraytracing_b200 as D
from dsgpuraytracing_b200 import scenes as S
V, F = S.torus_knot(); V = V.astype(np.float32).astype(np.float64)
sc = S.cb_mesh_scene(V, F); cam = S.cam_dragon(1920, 1080)
core = D.Core(0); core.set_params(16, 4, 8, 0); core.load(sc, camera=cam); core.set_option("stage_timing", 1)
for tm in [20, 28]:
  for rf in [12, 16, 20, 24, 28, 30]:
    core.set_option("postpone_min_lanes", tm); core.set_option("refill_busy_lanes", rf)
    core.render()
    rgb, st = core.render()
    print("tri_min %2d refill %2d  Mrays/s %7.1f  extend %.4f connect %.4f shade %.4f" % (tm, rf, st.segments/st.gpu_seconds/1e6, st.extend_seconds, st.connect_seconds, st.shade_seconds))
