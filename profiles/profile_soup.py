"""ncu driver for the HBM regime (BASELINE config 5): an N-Mi-triangle soup at 3840x2160, one 1-spp batch (8.3 M paths), twice.
  ncu --set full --clock-control none -k regex:k_trace -s 18 -c 2 -o soup python profiles/profile_soup.py 8
captures the depth-0 extend and connect launches of the second render (wide BVH + primitive records = 578 MB >> L2 at 8 Mi)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dsgpuraytracing_b200 as D
from dsgpuraytracing_b200 import scenes as S

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8
sc, cam = S.triangle_soup(n << 20)
core = D.Core(0)
core.set_params(1, 1, 8, 0)
core.load(sc, camera=cam)
core.set_option("stage_timing", 1)
for i in range(2):
    rgb, st = core.render()
print("soup %d Mi: segments %d  gpu_s %.4f  Mrays/s %.1f  extend %.4f connect %.4f shade %.4f" % (
    n, st.segments, st.gpu_seconds, st.segments / st.gpu_seconds / 1e6, st.extend_seconds, st.connect_seconds, st.shade_seconds))
core.set_option("count_traversal", 1)
_, c = core.render()
print("extend: %.2f nodes + %.2f prims per ray; connect: %.2f nodes + %.2f prims per ray; rays %d + %d" % (
    c.extend_nodes / max(c.extend_rays, 1), c.extend_prims / max(c.extend_rays, 1), c.connect_nodes / max(c.shadow_rays, 1),
    c.connect_prims / max(c.shadow_rays, 1), c.extend_rays, c.shadow_rays))
