"""Effect of the three round-2 refinements of the host-built wide BVH -- the light-aligned quantisation grid (wide_bvh.h EndPlane), the
coplanar slot mates dropped with a ray's source (layout.h WideNode::flat) and the regrouping of the top wide nodes (wide_bvh.cpp step
2b, opt-in) -- on the traversal counters, measured with the CPU walk of the product's own host-device
code (tests/cpu_walk): node visits and primitive tests per path segment, 2 spp at 192x108
(stand-ins) or the fixtures' small camera.  python tools/sweeps/light_grid_probe.py"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import dsgpuraytracing_b200 as D
from dsgpuraytracing_b200 import scenes as S
from tests.cpuwalk import Walk
from tests.scenes import CONFIGS

def run(arr, bvh, cam, depth):
    out = []
    for env in ({"CW_NO_LIGHT_GRID": "1", "CW_NO_FLAT_SLOTS": "1"}, {"CW_NO_FLAT_SLOTS": "1"}, {}, {"CW_REGROUP": "1"}):
        for k in ("CW_NO_LIGHT_GRID", "CW_REGROUP", "CW_NO_FLAT_SLOTS"):
            os.environ.pop(k, None)
        os.environ.update(env)
        w = Walk(arr, bvh, 4, camera=cam)
        rgb, c = w.render(2, depth, seed=3)
        out.append((rgb, c, w.info()))
    seg = float(out[0][1][1] + out[0][1][2])
    same = all(np.allclose(out[0][0], o[0], rtol=0, atol=1e-6) and list(out[0][1][:3]) == list(o[1][:3]) for o in out[1:])
    return same, [o[1][3] / seg for o in out], [o[1][4] / seg for o in out], [o[2][1] for o in out]

print("| scene | same image (1e-6) + segment counts | node visits / segment: off -> defaults (grid + mates) -> + regroup_top | primitive tests / segment: off -> + light-aligned grid -> + coplanar mates dropped (the defaults) -> + regroup_top | wide levels without -> with regroup_top |")
print("|---|---|---|---|---|")
rows = [(n, mk(192, 108), 8) for n, mk in (("c2 CBdragon stand-in", S.cbdragon_standin), ("c3 CBlucy (glass) stand-in", S.cblucy_standin))]
for name, (sc, cam), depth in rows:
    same, n, p, lv = run(sc, D.build_bvh2(sc), cam, depth)
    print(f"| {name} | {same} | {n[0]:.3f} -> {n[2]:.3f} -> {n[3]:.3f} ({100 * (n[3] / n[0] - 1):+.1f} %) | {p[0]:.3f} -> {p[1]:.3f} -> {p[2]:.3f} ({100 * (p[2] / p[0] - 1):+.1f} %) -> {p[3]:.3f} | {lv[0]} -> {lv[3]} |")
for name in ("CBspheres_lambertian", "CBspheres", "CBgems", "CBcoil", "CBbunny", "bunny"):
    g = dict(np.load(os.path.join(os.path.dirname(__file__), "..", "..", "tests", "golden", name + ".npz")))
    same, n, p, lv = run(g, g, g["small_camera"], CONFIGS[name]["depth"])
    print(f"| {name} | {same} | {n[0]:.3f} -> {n[2]:.3f} -> {n[3]:.3f} ({100 * (n[3] / n[0] - 1):+.1f} %) | {p[0]:.3f} -> {p[1]:.3f} -> {p[2]:.3f} ({100 * (p[2] / p[0] - 1):+.1f} %) -> {p[3]:.3f} | {lv[0]} -> {lv[3]} |")
sc, cam = S.triangle_soup(1 << 17, W=192, H=108)
same, n, p, lv = run(sc, D.build_bvh2(sc), cam, 8)
print(f"| 128 Ki-triangle soup | {same} | {n[0]:.3f} -> {n[2]:.3f} -> {n[3]:.3f} ({100 * (n[3] / n[0] - 1):+.1f} %) | {p[0]:.3f} -> {p[1]:.3f} -> {p[2]:.3f} ({100 * (p[2] / p[0] - 1):+.1f} %) -> {p[3]:.3f} | {lv[0]} -> {lv[3]} |")
