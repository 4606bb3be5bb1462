"""Collapse cost model on the config-5 soups: SAH cost of a primitive test relative to a wide-node visit (option
collapse_prim_cost_pct) vs frame rate and fetch counts.  python tools/sweeps/sweep_collapse_soup.py [Mi triangles] [spp]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import dsgpuraytracing_b200 as D
from dsgpuraytracing_b200 import scenes as S
n = int(sys.argv[1]) if len(sys.argv) > 1 else 8
spp = int(sys.argv[2]) if len(sys.argv) > 2 else 8
sc, cam = S.triangle_soup(n << 20)
bvh = D.build_bvh2(sc)
for pct in (10, 25, 50, 100, 200, 400):
    core = D.Core(0)
    core.set_params(spp, 1, 8, 0)
    core.set_option("collapse_prim_cost_pct", pct)
    core.load(sc, camera=cam, bvh=bvh)
    core.set_option("stage_timing", 1)
    core.render(); rgb, st = core.render()
    core.set_option("count_traversal", 1); core.set_params(1, 1, 8, 0)
    _, c = core.render()
    info = core.accel_info()
    print("soup %d Mi  prim cost %3d %%  wide nodes %8d depth %2d  Mrays/s %7.1f  extend %.4f connect %.4f  nodes/seg %.2f prims/seg %.2f" % (
        n, pct, info["wide_nodes"], info["max_depth"], st.segments / st.gpu_seconds / 1e6, st.extend_seconds, st.connect_seconds,
        c.nodes_visited / c.segments, c.prims_tested / c.segments), flush=True)
    core.close()
