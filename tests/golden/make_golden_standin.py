"""Generates the fixtures of the MEASURED workloads (what bench.py and profiles/r2_configs.md time), from the compiled
reference (oracle/_ref/ref_driver) and the pinned oracle port.  Run in the build container:

    python tests/golden/make_golden_standin.py [cbdragon_standin cblucy_standin soup1m C1]

standin_<name>.npz  (name = cbdragon_standin | cblucy_standin; scene = the .dae + cam_dragon.info that
                     dsgpuraytracing_b200.scenes.write_standin emits, loaded by the REFERENCE's ColladaParser)
  scene_sha           sha256 over the reference loader's dump (prim_type, prim_bsdf, tri_pos, tri_nrm, sphere, bsdf_*, light_*)
  hit_id/hit_t/hit_tie   BVHAccel::intersect primary ids at STANDIN_ID_RES (reference) + brute-force exact-tie mask (oracle)
  hit_id_full, hit_t_full_sha   the same at 1920x1080 (ids stored, t as a sha256 of the float64 bytes)
  camera, camera_full, small_camera, ref_camera
  small_rgb/small_cnt    2-spp rand()-driven render at SMALL_RES, srand(1) (bit-exact target of the oracle port)
  ref_rgb/ref_rgb_b/ref_cnt   two independent RMSE_SPP-spp reference renders at RMSE_RES (8 seeds each)
  philox_rgb/philox_cnt  oracle port, Philox streams, RMSE_SPP spp at RMSE_RES, seed 0
soup1m.npz   1 Mi-triangle soup (scenes.triangle_soup): oracle-port primary ids/t at STANDIN_ID_RES (the soup bypasses .dae)
c1_fullres.npz   BASELINE configs[0] at its own 480x360: reference ids/t + tie mask, camera
"""
import hashlib, os, sys, tempfile
from concurrent.futures import ProcessPoolExecutor, ThreadPoolExecutor
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as O
from dsgpuraytracing_b200 import scenes as S
from tests.scenes import CONFIGS, STANDIN_CONFIGS, STANDIN_ID_RES, SMALL_RES, RMSE_RES, RMSE_SPP, scene_sha

OUT = os.path.dirname(os.path.abspath(__file__))
NPROC = 8


def _philox_part(args):
    arrays, cam, nl, depth, k = args
    W, H = RMSE_RES
    sc = O.Scene(dict(arrays, camera=cam))
    per = RMSE_SPP // NPROC
    rgb, cnt = sc.render(W, H, RMSE_SPP, nl, depth, rng="philox", seed=0, spp_begin=k * per, spp_count=per)
    return rgb.astype(np.float64), cnt


def multi_render(dae, cam, cfg, W, H, spp, seeds):
    per = spp // len(seeds)
    def one(seed):
        return O.run_reference(dae, W, H, cam=cam, spp=per, nl=cfg["nl"], depth=cfg["depth"], seed=seed, render=True)
    with ThreadPoolExecutor(len(seeds)) as ex:
        outs = list(ex.map(one, seeds))
    rgb = np.mean([o["rgb"].astype(np.float64) for o in outs], axis=0).astype(np.float32)
    return rgb, np.sum([o["counters"][:2] for o in outs], axis=0)


def make_standin(name):
    cfg = STANDIN_CONFIGS[name]
    out = {}
    with tempfile.TemporaryDirectory() as td:
        dae, cam = S.write_standin(name, td)
        W, H = STANDIN_ID_RES
        d = O.run_reference(dae, W, H, cam=cam, dump_scene=True, ids=True)
        arrays = {k: d[k] for k in O.SCENE_KEYS if k != "camera"}
        out["scene_sha"] = np.frombuffer(scene_sha(arrays).encode(), np.uint8)
        out["hit_id"], out["hit_t"], out["camera"] = d["hit_id"], d["hit_t"], d["camera"]
        sc = O.Scene(dict(arrays, camera=d["camera"], **{k: d[k] for k in O.BVH_KEYS}))
        ids, ts, tie = sc.primary_hits(W, H, ties=True)
        assert np.array_equal(ids, d["hit_id"]) and np.array_equal(ts, d["hit_t"]), "oracle port != reference ids"
        out["hit_tie"] = tie
        f = O.run_reference(dae, 1920, 1080, cam=cam, dump_scene=True, ids=True)
        out["hit_id_full"] = f["hit_id"]; out["camera_full"] = f["camera"]
        out["hit_t_full_sha"] = np.frombuffer(hashlib.sha256(np.ascontiguousarray(f["hit_t"]).tobytes()).hexdigest().encode(), np.uint8)
        sw, sh = SMALL_RES
        s = O.run_reference(dae, sw, sh, cam=cam, spp=2, nl=cfg["nl"], depth=cfg["depth"], seed=1, render=True, dump_scene=True)
        out["small_rgb"], out["small_cnt"], out["small_camera"] = s["rgb"], s["counters"][:2], s["camera"]
        rw, rh = RMSE_RES
        out["ref_rgb"], out["ref_cnt"] = multi_render(dae, cam, cfg, rw, rh, RMSE_SPP, list(range(101, 109)))
        out["ref_rgb_b"], _ = multi_render(dae, cam, cfg, rw, rh, RMSE_SPP, list(range(201, 209)))
        out["ref_camera"] = O.run_reference(dae, rw, rh, cam=cam, dump_scene=True)["camera"]
        with ProcessPoolExecutor(NPROC) as ex:
            parts = list(ex.map(_philox_part, [(arrays, out["ref_camera"], cfg["nl"], cfg["depth"], k) for k in range(NPROC)]))
        out["philox_rgb"] = np.sum([p[0] for p in parts], axis=0).astype(np.float32)
        out["philox_cnt"] = np.sum([p[1] for p in parts], axis=0)
    np.savez_compressed(os.path.join(OUT, "standin_" + name + ".npz"), **out)
    print(name, "ties", int(out["hit_tie"].sum()), "full-res hits", int((out["hit_id_full"] >= 0).sum()),
          "ref mean", out["ref_rgb"].mean(axis=(0, 1)), "philox mean", out["philox_rgb"].mean(axis=(0, 1)), flush=True)


def make_soup():
    W, H = STANDIN_ID_RES
    arr, cam = S.triangle_soup(1 << 20, W=W, H=H)
    sc = O.Scene(dict(arr, camera=cam))
    sc.build_bvh()
    ids, ts, _ = sc.primary_hits(W, H, ties=False)
    np.savez_compressed(os.path.join(OUT, "soup1m.npz"), hit_id=ids, hit_t=ts, camera=cam)
    print("soup1m hits", int((ids >= 0).sum()), "of", ids.size, flush=True)


def make_c1():
    cfg = CONFIGS["CBspheres_lambertian"]
    d = O.run_reference(O.ref_scene_path(cfg["file"]), 480, 360, dump_scene=True, ids=True)
    sc = O.Scene({k: d[k] for k in O.SCENE_KEYS + O.BVH_KEYS})
    ids, ts, tie = sc.primary_hits(480, 360, ties=True)
    assert np.array_equal(ids, d["hit_id"]) and np.array_equal(ts, d["hit_t"])
    np.savez_compressed(os.path.join(OUT, "c1_fullres.npz"), hit_id=d["hit_id"], hit_t=d["hit_t"], hit_tie=tie, camera=d["camera"])
    print("C1 480x360 ties", int(tie.sum()), flush=True)


if __name__ == "__main__":
    for n in (sys.argv[1:] or list(STANDIN_CONFIGS) + ["soup1m", "C1"]):
        if n == "soup1m":
            make_soup()
        elif n == "C1":
            make_c1()
        else:
            make_standin(n)
