"""Generates tests/golden/<scene>_philox.npz: the ORACLE PORT (oracle/pt_oracle.c, pinned bit-exactly to the compiled
reference by tests/test_oracle_vs_reference.py) rendered with the Philox streams the CUDA path uses
(seed 0, RMSE_SPP spp, RMSE_RES).  Because both sides consume identical uniforms, the GPU image must match this one
far below Monte-Carlo noise -- this is how the "RMSE < 1 % of mean radiance at 1024 spp" gate is evaluated
(two independent 1024-spp runs of the reference itself differ by 8-110 % per-pixel RMSE).

    python tests/golden/make_golden_philox.py [scene ...]
"""
import os, sys
from concurrent.futures import ProcessPoolExecutor
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as O
from tests.scenes import CONFIGS, RMSE_RES, RMSE_SPP

OUT = os.path.dirname(os.path.abspath(__file__))
NPROC = 8


def part(args):
    name, k = args
    g = np.load(os.path.join(OUT, name + ".npz")); g = {q: g[q] for q in g.files}
    cfg = CONFIGS[name]; W, H = RMSE_RES
    sc = O.Scene(g).with_camera(g["ref_camera"])
    per = RMSE_SPP // NPROC
    rgb, cnt = sc.render(W, H, RMSE_SPP, cfg["nl"], cfg["depth"], rng="philox", seed=0, spp_begin=k * per, spp_count=per)
    return rgb.astype(np.float64), cnt


if __name__ == "__main__":
    for name in (sys.argv[1:] or list(CONFIGS)):
        with ProcessPoolExecutor(NPROC) as ex:
            parts = list(ex.map(part, [(name, k) for k in range(NPROC)]))
        rgb = np.sum([p[0] for p in parts], axis=0).astype(np.float32)
        cnt = np.sum([p[1] for p in parts], axis=0)
        np.savez_compressed(os.path.join(OUT, name + "_philox.npz"), philox_rgb=rgb, philox_cnt=cnt)
        print(name, rgb.mean(axis=(0, 1)), cnt, flush=True)
