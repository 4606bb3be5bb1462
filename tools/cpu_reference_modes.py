#!/usr/bin/env python
"""SURVEY.md 8d, 'CPU reference timing beside it': the compiled reference (oracle/_ref/ref_driver) on the bench scene in
three modes -- (i) one process, one thread; (ii) P independent single-threaded processes with different seeds; (iii) the
reference's own -t P mode (all threads share glibc rand()'s lock).  Bounded sample: 1 spp of a 960x540 frame
per process.  Prints one JSON line (Mseg/s per mode)."""
import json, os, sys, tempfile, time
from concurrent.futures import ThreadPoolExecutor
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from oracle import oracle as O
from dsgpuraytracing_b200 import scenes as S

W, H, NL, DEPTH = 960, 540, 4, 8
P = os.cpu_count() or 1
with tempfile.TemporaryDirectory() as td:
    V, F = S.torus_knot(); V = V.astype(np.float32).astype(np.float64)
    dae = os.path.join(td, "s.dae"); cam = os.path.join(td, "c.info")
    S.write_cb_mesh_dae(dae, V, F); S.write_cam_info(cam, S.cam_dragon(W, H))

    def run(seed, threads, spp=1):
        c = O.run_reference(dae, W, H, cam=cam, spp=spp, nl=NL, depth=DEPTH, seed=seed, render=True, threads=threads)["counters"]
        return float(c[0] + c[1]), float(c[2])
    s1, t1 = run(1, 1)
    with ThreadPoolExecutor(P) as ex:
        rs = list(ex.map(lambda k: run(100 + k, 1), range(P)))
    sp, tp = sum(r[0] for r in rs), max(r[1] for r in rs)
    st, tt = run(7, P, spp=2)
print(json.dumps({"scene": "bench workload (Cornell box + 100 012-triangle mesh), %dx%d, 1 spp per process, l%d m%d" % (W, H, NL, DEPTH),
                  "host_cores": P,
                  "one_process_one_thread_Mseg_s": s1 / t1 / 1e6,
                  "P_single_threaded_processes_Mseg_s": sp / tp / 1e6,
                  "reference_own_t_P_threads_Mseg_s": st / tt / 1e6,
                  "note": "the reference's -t P mode serialises on glibc rand() (SURVEY F4); bench.py uses mode (ii)"}))
