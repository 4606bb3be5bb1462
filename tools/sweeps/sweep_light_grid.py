"""A/B of the light-aligned quantisation grid (option "light_aligned_grid", wide_bvh.h EndPlane) and of the regrouped top nodes
("regroup_top", wide_bvh.cpp step 2b) and of the coplanar slot mates dropped with a ray's source ("drop_coplanar_mates", layout.h
WideNode::flat) on the bench scene (c2) and the glass stand-in (c3) at 1920x1080, alternating in ONE process; per run: Mrays/s, stage seconds of a 64-spp render (second of two)
and the exact traversal counters of a 4-spp render.  The two frames are compared (must be equal up to the order of float atomics)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import dsgpuraytracing_b200 as D
from dsgpuraytracing_b200 import scenes as S
spp = int(sys.argv[1]) if len(sys.argv) > 1 else 64
for name, mk in (("c2", S.cbdragon_standin), ("c3", S.cblucy_standin)):
    sc, cam = mk(1920, 1080)
    bvh = D.build_bvh2(sc)
    core = D.Core(0); core.set_params(spp, 4, 8, 0); core.load(sc, camera=cam, bvh=bvh); core.set_option("stage_timing", 1)
    frames = {}
    for on, fl, rg in ((0, 0, 0), (1, 0, 0), (1, 1, 0), (1, 1, 1), (0, 0, 0), (1, 0, 0), (1, 1, 0), (1, 1, 1)):
        core.set_option("light_aligned_grid", on); core.set_option("drop_coplanar_mates", fl); core.set_option("regroup_top", rg); core.build_accel()
        core.render(); rgb, st = core.render()
        frames[(on, fl, rg)] = np.array(rgb, copy=True)
        core.set_option("count_traversal", 1); core.set_params(4, 4, 8, 0); _, c = core.render(); core.set_option("count_traversal", 0); core.set_params(spp, 4, 8, 0)
        print("%s light_aligned_grid %d drop_coplanar_mates %d regroup_top %d  wide levels %d  Mrays/s %7.1f  gpu_s %.4f extend %.4f connect %.4f shade %.4f  nodes/seg %.3f prims/seg %.3f" % (
            name, on, fl, rg, core.accel_info()["max_depth"], st.segments / st.gpu_seconds / 1e6, st.gpu_seconds, st.extend_seconds, st.connect_seconds, st.shade_seconds,
            c.nodes_visited / c.segments, c.prims_tested / c.segments), flush=True)
    d = np.abs(frames[(0, 0, 0)] - frames[(1, 1, 0)]); print("%s frames all off vs the defaults (grid + mates): max abs diff %.3g, mean rel diff %.3g" % (name, d.max(), d.mean() / frames[(0, 0, 0)].mean()), flush=True)
    core.close()
