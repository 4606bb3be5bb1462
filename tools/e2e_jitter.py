"""Per-step wall time of bench.py's e2e sequence (upload_accel + camera + params + dsrt_render into a pinned host frame) on the bench
scene, with a breakdown: python tools/e2e_jitter.py [steps]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import dsgpuraytracing_b200 as D
from dsgpuraytracing_b200 import scenes as S
n = int(sys.argv[1]) if len(sys.argv) > 1 else 10
sc, cam = S.load_standin("cbdragon_standin", 1920, 1080)
bvh = D.build_bvh2(sc)
core = D.Core(0); core.set_params(256, 4, 8, 0); core.load(sc, camera=cam, bvh=bvh)
host = torch.zeros(1080 * 1920 * 3, dtype=torch.float32).pin_memory()
out = host.numpy().reshape(1080, 1920, 3)
for i in range(n):
    t0 = time.perf_counter(); core.upload_accel(); t1 = time.perf_counter()
    core.set_camera(cam); core.set_params(256, 4, 8, 0); t2 = time.perf_counter()
    rgb, st = core.render(out=out); t3 = time.perf_counter()
    print("step %2d  upload %.4f  params %.4f  render call %.4f  (gpu_seconds %.4f)  total %.4f" % (i, t1 - t0, t2 - t1, t3 - t2, st.gpu_seconds, t3 - t0), flush=True)
