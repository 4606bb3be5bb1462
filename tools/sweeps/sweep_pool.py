import sys; sys.path.insert(0,'.')
import numpy as np
import dsgpuraytracing_b200 as D
from dsgpuraytracing_b200 import scenes as S
V, F = S.torus_knot(); V = V.astype(np.float32).astype(np.float64)
sc = S.cb_mesh_scene(V, F); cam = S.cam_dragon(1920, 1080)
core = D.Core(0); core.set_params(128, 4, 8, 0); core.load(sc, camera=cam); core.set_option("stage_timing", 1)
for pb in [1, 4, 8, 16, 32, 64]:
    core.set_option("pool_batches", pb)
    core.render(spp_count=128)
    rgb, st = core.render(spp_count=128)
    print("pool_batches %2d  Mrays/s %7.1f  gpu_s %.4f extend %.4f connect %.4f shade %.4f launches %d" % (pb, st.segments/st.gpu_seconds/1e6, st.gpu_seconds, st.extend_seconds, st.connect_seconds, st.shade_seconds, st.kernel_launches))
