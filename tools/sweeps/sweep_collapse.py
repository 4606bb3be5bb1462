"""Collapse cost model: SAH cost of a primitive test relative to a wide-node visit (collapse_prim_cost_pct)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import dsgpuraytracing_b200 as D
from dsgpuraytracing_b200 import scenes as S
V, F = S.torus_knot(); V = V.astype(np.float32).astype(np.float64)
sc = S.cb_mesh_scene(V, F); cam = S.cam_dragon(1920, 1080)
bvh = D.build_bvh2(sc)
core = D.Core(0); core.set_params(64, 4, 8, 0); core.load(sc, camera=cam, bvh=bvh); core.set_option("stage_timing", 1)
for pct in (100, 25, 40, 60, 150, 250, 400):
    core.set_option("collapse_prim_cost_pct", pct); core.build_accel()
    core.render(); rgb, st = core.render()
    core.set_option("count_traversal", 1); core.set_params(4, 4, 8, 0); _, c = core.render(); core.set_option("count_traversal", 0); core.set_params(64, 4, 8, 0)
    print("prim cost %.2f  wide nodes %6d  Mrays/s %7.1f  extend %.4f connect %.4f  nodes/seg %.2f prims/seg %.2f" % (
        pct / 100, core.accel_info()["wide_nodes"], st.segments / st.gpu_seconds / 1e6, st.extend_seconds, st.connect_seconds,
        c.nodes_visited / c.segments, c.prims_tested / c.segments), flush=True)
