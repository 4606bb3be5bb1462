"""Generates tests/golden/*.npz from the COMPILED REFERENCE (oracle/_ref/ref_driver = the reference's own CPU
path tracer, see oracle/build_ref.sh).  Run in the build container where /root/reference exists:

    python tests/golden/make_golden.py [scene ...]

Per scene it writes <name>.npz with
  flat scene arrays + camera + SAH BVH dump            (ref_driver --dump-scene, at ID_RES)
  hit_id / hit_t   primary closest-hit ids at pixel centres at ID_RES   (BVHAccel::intersect)
  hit_tie          exact-tie mask (computed with the oracle port, brute force)
  small_rgb/small_cnt   2-spp render at SMALL_RES, srand(1), 1 thread (bit-exact target for the oracle port)
  ref_rgb          RMSE_SPP-spp render at RMSE_RES: mean of 8 single-threaded runs with seeds 101..108
  ref_rgb_b        an independent second render (seeds 201..208) -> the Monte-Carlo noise floor
  ref_cnt          [closest, any] segment counts summed over the ref_rgb runs
"""
import os, sys
from concurrent.futures import ThreadPoolExecutor
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as O
from tests.scenes import CONFIGS, ID_RES, SMALL_RES, RMSE_RES, RMSE_SPP

OUT = os.path.dirname(os.path.abspath(__file__))


def camp(cfg):
    return O.ref_scene_path(cfg["cam"]) if cfg["cam"] else None


def multi_render(cfg, W, H, spp, seeds):
    per = spp // len(seeds)
    def one(seed):
        return O.run_reference(O.ref_scene_path(cfg["file"]), W, H, cam=camp(cfg), spp=per, nl=cfg["nl"],
                               depth=cfg["depth"], seed=seed, render=True)
    with ThreadPoolExecutor(len(seeds)) as ex:
        outs = list(ex.map(one, seeds))
    rgb = np.mean([o["rgb"].astype(np.float64) for o in outs], axis=0).astype(np.float32)
    cnt = np.sum([o["counters"][:2] for o in outs], axis=0)
    return rgb, cnt


def make(name):
    cfg = CONFIGS[name]
    path = O.ref_scene_path(cfg["file"])
    W, H = ID_RES
    d = O.run_reference(path, W, H, cam=camp(cfg), dump_scene=True, ids=True)
    out = {k: d[k] for k in O.SCENE_KEYS + O.BVH_KEYS}
    out["hit_id"], out["hit_t"] = d["hit_id"], d["hit_t"]
    sc = O.Scene(out)
    ids, ts, tie = sc.primary_hits(W, H, ties=True)
    assert np.array_equal(ids, d["hit_id"]) and np.array_equal(ts, d["hit_t"]), "oracle port != reference ids"
    out["hit_tie"] = tie
    sw, sh = SMALL_RES
    s = O.run_reference(path, sw, sh, cam=camp(cfg), spp=2, nl=cfg["nl"], depth=cfg["depth"], seed=1, render=True,
                        dump_scene=True)
    out["small_rgb"], out["small_cnt"], out["small_camera"] = s["rgb"], s["counters"][:2], s["camera"]
    rw, rh = RMSE_RES
    out["ref_rgb"], out["ref_cnt"] = multi_render(cfg, rw, rh, RMSE_SPP, list(range(101, 109)))
    out["ref_rgb_b"], _ = multi_render(cfg, rw, rh, RMSE_SPP, list(range(201, 209)))
    out["ref_camera"] = O.run_reference(path, rw, rh, cam=camp(cfg), dump_scene=True)["camera"]
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    print(name, "prims", sc.n_prims, "nodes", len(out["node_start"]), "ties", int(tie.sum()),
          "ref mean", out["ref_rgb"].mean(axis=(0, 1)), flush=True)


if __name__ == "__main__":
    for n in (sys.argv[1:] or list(CONFIGS)):
        make(n)
