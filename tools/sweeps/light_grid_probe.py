"""Effect of the light-aligned quantisation grid (wide_bvh.h EndPlane) on the traversal counters, measured with the CPU walk of the
product's own host-device code (tests/cpu_walk): primitive tests and node visits per path segment with the option off / on, 2 spp at
192x108 (stand-ins) or the fixtures' small camera.  python tools/sweeps/light_grid_probe.py"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import dsgpuraytracing_b200 as D
from dsgpuraytracing_b200 import scenes as S
from tests.cpuwalk import Walk
from tests.scenes import CONFIGS

def run(arr, bvh, cam, depth):
    out = []
    for off in (True, False):
        if off: os.environ["CW_NO_LIGHT_GRID"] = "1"
        else: os.environ.pop("CW_NO_LIGHT_GRID", None)
        rgb, c = Walk(arr, bvh, 4, camera=cam).render(2, depth, seed=3)
        out.append((rgb, c))
    (r0, c0), (r1, c1) = out
    seg = float(c0[1] + c0[2])
    return np.array_equal(r0, r1) and list(c0[:3]) == list(c1[:3]), c0[3] / seg, c1[3] / seg, c0[4] / seg, c1[4] / seg

print("| scene | image + segment counts identical | node visits / segment off -> on | primitive tests / segment off -> on |")
print("|---|---|---|---|")
for name, mk in (("c2 CBdragon stand-in", S.cbdragon_standin), ("c3 CBlucy (glass) stand-in", S.cblucy_standin)):
    sc, cam = mk(192, 108)
    same, n0, n1, p0, p1 = run(sc, D.build_bvh2(sc), cam, 8)
    print(f"| {name} | {same} | {n0:.3f} -> {n1:.3f} | {p0:.3f} -> {p1:.3f} ({100 * (p1 / p0 - 1):+.1f} %) |")
for name in ("CBspheres_lambertian", "CBspheres", "CBgems", "CBcoil", "CBbunny", "bunny"):
    g = dict(np.load(os.path.join(os.path.dirname(__file__), "..", "..", "tests", "golden", name + ".npz")))
    same, n0, n1, p0, p1 = run(g, g, g["small_camera"], CONFIGS[name]["depth"])
    print(f"| {name} | {same} | {n0:.3f} -> {n1:.3f} | {p0:.3f} -> {p1:.3f} ({100 * (p1 / p0 - 1):+.1f} %) |")
