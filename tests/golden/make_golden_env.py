"""Generates tests/golden/env_<scene>.npz: the compiled reference (oracle/_ref/ref_driver --envmap 64 32) rendering with
an EnvironmentLight built from a procedural lat-long map (no .exr ships with the reference).
    python tests/golden/make_golden_env.py
Contents: flat scene arrays (light list ends with type 4 = the environment light), env_rgb, camera, BVH dump,
small_rgb / small_cnt = 2-spp srand(1) render at SMALL_RES (bit-exact target for the oracle port)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as O
from tests.scenes import SMALL_RES

OUT = os.path.dirname(os.path.abspath(__file__))
CASES = {"env_CBspheres": ("CBspheres.dae", 5), "env_bunny": ("bunny.dae", 8), "env_CBgems": ("CBgems.dae", 8)}

if __name__ == "__main__":
    W, H = SMALL_RES
    for name, (dae, depth) in CASES.items():
        r = O.run_reference(O.ref_scene_path(dae), W, H, spp=2, nl=4, depth=depth, seed=1, dump_scene=True, render=True, envmap=(64, 32))
        out = {k: r[k] for k in O.SCENE_KEYS + O.BVH_KEYS}
        out["env_rgb"] = r["env_rgb"]; out["small_rgb"] = r["rgb"]; out["small_cnt"] = r["counters"][:2]; out["depth"] = np.array(depth)
        np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
        print(name, r["light_type"], r["rgb"].mean(axis=(0, 1)))
