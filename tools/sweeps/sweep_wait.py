import sys; sys.path.insert(0,'.')
import numpy as np
import dsgpuraytracing_b200 as D
from dsgpuraytracing_b200 import scenes as S
V, F = S.torus_knot(); V = V.astype(np.float32).astype(np.float64)
sc = S.cb_mesh_scene(V, F); cam = S.cam_dragon(1920, 1080)
core = D.Core(0); core.set_params(16, 4, 8, 0); core.load(sc, camera=cam); core.set_option("stage_timing", 1)
ref = None
for wm in [0, 1]:
  for tm in [4, 8, 12, 16, 20, 24]:
    core.set_option("postpone_wait_mode", wm); core.set_option("postpone_min_lanes", tm)
    core.render()
    rgb, st = core.render()
    if ref is None: ref = rgb
    print("wait %d tri_min %2d  Mrays/s %7.1f  extend %.4f connect %.4f  maxdiff %.2e" % (wm, tm, st.segments/st.gpu_seconds/1e6, st.extend_seconds, st.connect_seconds, np.abs(rgb-ref).max()))
