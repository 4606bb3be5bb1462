"""Does running TWO wavefront streams on one GPU hide the bandwidth-bound stages (generate / shade, 11 % of the frame) behind
the issue-bound traversal stages of the other stream?  A context over devices [0, 0] renders half of the samples on each
of two streams of the same GPU (the multi-GPU path, unchanged); DSRT_STAGGER=1 offsets the second stream by half a batch.
  python tools/sweeps/overlap_probe.py [spp]
Each line: configuration, wall seconds of the dsrt_render call (second of two), Mrays/s by wall time."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np
import dsgpuraytracing_b200 as D
from dsgpuraytracing_b200 import scenes as S

spp = int(sys.argv[1]) if len(sys.argv) > 1 else 256
V, F = S.torus_knot(); V = V.astype(np.float32).astype(np.float64)
sc = S.cb_mesh_scene(V, F); cam = S.cam_dragon(1920, 1080)
bvh = D.build_bvh2(sc)


def run(devices, opts, label):
    core = D.Core(devices=devices) if devices else D.Core(0)
    core.set_params(spp, 4, 8, 0); core.load(sc, camera=cam, bvh=bvh)
    for k, v in opts.items():
        core.set_option(k, v)
    best = None
    for i in range(3):
        t0 = time.perf_counter(); rgb, st = core.render(); t1 = time.perf_counter()
        if i and (best is None or t1 - t0 < best[0]):
            best = (t1 - t0, st)
    w, st = best
    print("%-46s wall %.4f s  gpu_s %.4f  Mrays/s(wall) %7.1f  mean %.6f" % (label, w, st.gpu_seconds, st.segments / w / 1e6, float(rgb.mean())), flush=True)
    core.close()


run(None, {}, "one stream")
for ctas in (0, 3, 4, 5):
    for batch in (0, 4):
        run([0, 0], {"max_ctas_per_sm": ctas, "batch_spp": batch}, "two streams, max_ctas %d, batch_spp %d%s" % (ctas, batch, ", staggered" if os.environ.get("DSRT_STAGGER") else ""))
