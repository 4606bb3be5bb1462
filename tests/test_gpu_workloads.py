"""Parity on the workloads that are actually MEASURED (bench.py, profiles/r2_configs.md): the CBdragon / CBlucy stand-ins
(through the .dae both arms start from), BASELINE configs[0] at its own 480x360, the 1 Mi-triangle soup, plus the memory
sizing of the wavefront.  Fixtures: tests/golden/standin_*.npz, c1_fullres.npz, soup1m.npz (tests/golden/make_golden_standin.py,
from the compiled reference and the pinned oracle port).  Everything goes through the C ABI; the oracle only checks.

Set DSRT_PARITY_LOG=<file> to append one JSON line per gate with the measured figures (profiles/r2_parity.md is built from it)."""
import hashlib
import json
import os

import numpy as np
import pytest

import dsgpuraytracing_b200 as D
from dsgpuraytracing_b200 import scenes as S
from oracle import oracle as O
from tests.scenes import CONFIGS, RMSE_RES, RMSE_SPP, SMALL_RES, STANDIN_CONFIGS, STANDIN_ID_RES, scene_sha
from tests.util import images_match, plog

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
STANDINS = list(STANDIN_CONFIGS)


def fixture(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    return {k: z[k] for k in z.files}


_scene_cache = {}


def standin(name):
    if name not in _scene_cache:
        _scene_cache[name] = S.load_standin(name, *STANDIN_ID_RES)[0]
    return _scene_cache[name]


def block_mean(a, k):
    H, W, _ = a.shape
    return a[:H // k * k, :W // k * k].reshape(H // k, k, W // k, k, 3).mean(axis=(1, 3))


# ---- CPU: both arms render the SAME scene ---------------------------------------------------------------------------------
@pytest.mark.parametrize("name", STANDINS)
def test_standin_scene_is_bit_identical_on_both_arms(name, tmp_path):
    """The .dae written for the reference arm, loaded by the PRODUCT loader, hashes to what the REFERENCE's ColladaParser +
    half-edge mesh produced from the same file when the fixture was made (primitive order, vertex rotation, positions,
    vertex normals, BSDFs, light); the flat-array generator agrees in order / rotation / positions and to 1 ulp in normals."""
    fx = fixture("standin_" + name)
    sc = standin(name)
    assert scene_sha(sc) == bytes(fx["scene_sha"]).decode()
    flat = (S.cbdragon_standin if name == "cbdragon_standin" else S.cblucy_standin)(*STANDIN_ID_RES)[0]
    for k in ("prim_type", "prim_bsdf", "tri_pos", "sphere", "bsdf_type", "bsdf_param", "light_type", "light_param"):
        assert np.array_equal(np.asarray(sc[k]).reshape(-1), np.asarray(flat[k]).reshape(-1)), k
    assert np.abs(np.asarray(sc["tri_nrm"]).reshape(-1) - np.asarray(flat["tri_nrm"]).reshape(-1)).max() <= 4.5e-16
    if O.have_reference():      # live: the compiled reference on the freshly written file
        dae, cam = S.write_standin(name, str(tmp_path), *STANDIN_ID_RES)
        d = O.run_reference(dae, *STANDIN_ID_RES, cam=cam, dump_scene=True)
        assert scene_sha(d) == bytes(fx["scene_sha"]).decode()
        assert np.array_equal(d["camera"], fx["camera"])


# ---- gate 1 on the measured scenes ------------------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("name", STANDINS)
def test_standin_primary_ids_bit_exact(name, core):
    fx = fixture("standin_" + name); cfg = STANDIN_CONFIGS[name]
    sc = standin(name)
    core.set_params(1, cfg["nl"], cfg["depth"], 0)
    core.load(sc, camera=fx["camera"])
    ids, ts = core.primary_hits(mode=1)
    ok = ~fx["hit_tie"].astype(bool)
    assert np.array_equal(ids[ok], fx["hit_id"][ok]), f"{(ids != fx['hit_id'])[ok].sum()} id mismatches"
    assert np.array_equal(ts[ok], fx["hit_t"][ok])
    ids0, ts0 = core.primary_hits(mode=0)             # production float kernel: bounded, and recorded
    flips = int((ids0 != fx["hit_id"]).sum())
    plog(gate="ids", scene=name, res=list(STANDIN_ID_RES), parity_kernel_mismatches=0, ties=int(fx["hit_tie"].sum()),
         production_float_kernel_mismatches=flips, pixels=int(ids.size))
    assert flips <= 2e-4 * ids.size, flips
    # one full 1920x1080 map
    core.set_camera(fx["camera_full"])
    idf, tf = core.primary_hits(mode=1)
    bad = np.argwhere(idf != fx["hit_id_full"])
    if len(bad):        # only an exact tie may differ: brute-force those pixels
        o = O.Scene(dict(sc, camera=fx["camera_full"]))
        for (y, x) in bad[:50]:
            ro, rd = O.generate_ray(fx["camera_full"], (x + 0.5) / 1920, (y + 0.5) / 1080)
            pid, t = o.closest_hit_brute(ro, rd)
            assert tf[y, x] == t, (y, x, idf[y, x], fx["hit_id_full"][y, x])
        assert len(bad) <= 50
    else:
        assert hashlib.sha256(np.ascontiguousarray(tf).tobytes()).hexdigest() == bytes(fx["hit_t_full_sha"]).decode()
    plog(gate="ids_full", scene=name, res=[1920, 1080], mismatches=int(len(bad)), pixels=int(idf.size))


@pytest.mark.gpu
def test_c1_at_its_baseline_resolution(core):
    """BASELINE.json configs[0] as quoted: CBspheres_lambertian 480x360, 16 spp, 4 light samples, depth 5."""
    fx = fixture("c1_fullres"); g = fixture("CBspheres_lambertian"); cfg = CONFIGS["CBspheres_lambertian"]
    core.set_params(16, cfg["nl"], cfg["depth"], 7)
    core.load(g, camera=fx["camera"])
    ids, ts = core.primary_hits(mode=1)
    ok = ~fx["hit_tie"].astype(bool)
    assert np.array_equal(ids[ok], fx["hit_id"][ok]) and np.array_equal(ts[ok], fx["hit_t"][ok])
    ids0, _ = core.primary_hits(mode=0)
    rgb, st = core.render()
    ref, cnt = O.Scene(g).with_camera(fx["camera"]).render(480, 360, 16, cfg["nl"], cfg["depth"], rng="philox", seed=7)
    okimg, info = images_match(rgb, ref)
    plain = float((np.sqrt(((rgb - ref) ** 2).mean(axis=(0, 1))) / ref.mean(axis=(0, 1))).max())
    plog(gate="c1_480x360_16spp", production_float_kernel_mismatches=int((ids0 != fx["hit_id"]).sum()), plain_rel_rmse=plain, **info,
         extend=[int(st.extend_rays), int(cnt[0])], shadow=[int(st.shadow_rays), int(cnt[1])])
    assert okimg, info
    assert abs(int(st.extend_rays) - int(cnt[0])) <= 2e-4 * cnt[0] + 2 and abs(int(st.shadow_rays) - int(cnt[1])) <= 2e-4 * cnt[1] + 8


@pytest.mark.gpu
def test_soup_1mi_primary_ids(core):
    """Config 5's generator at 1 Mi triangles: closest-hit ids / t of the production wide BVH (parity kernel) == the oracle
    port's BVHAccel::intersect restatement on its own reference-order SAH tree (fixture soup1m.npz)."""
    fx = fixture("soup1m")
    W, H = STANDIN_ID_RES
    sc, cam = S.triangle_soup(1 << 20, W=W, H=H)
    assert np.array_equal(cam, fx["camera"])
    core.set_params(1, 1, 8, 0)
    core.load(sc, camera=cam)
    ids, ts = core.primary_hits(mode=1)
    bad = np.argwhere(ids != fx["hit_id"])
    plog(gate="ids", scene="soup1m", res=[W, H], parity_kernel_mismatches=int(len(bad)), pixels=int(ids.size),
         production_float_kernel_mismatches=int((core.primary_hits(mode=0)[0] != fx["hit_id"]).sum()))
    assert len(bad) <= 4                          # exact ties only (the fixture carries no brute-force tie mask at this size)
    same = ids == fx["hit_id"]
    assert np.array_equal(ts[same], fx["hit_t"][same])
    for (y, x) in bad:
        assert ts[y, x] == fx["hit_t"][y, x]      # a tie: same t, different primitive


# ---- gate 2 on the measured scenes ------------------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("name", STANDINS)
def test_standin_render_matches_oracle_path_by_path(name, core):
    fx = fixture("standin_" + name); cfg = STANDIN_CONFIGS[name]
    sc = standin(name)
    W, H = 240, 135
    cam = fx["camera"].copy(); cam[14] *= H / cam[13]; cam[12], cam[13] = W, H
    ref, cnt = O.Scene(dict(sc, camera=cam)).render(W, H, 8, cfg["nl"], cfg["depth"], rng="philox", seed=11)
    core.set_params(8, cfg["nl"], cfg["depth"], 11)
    core.load(sc, camera=cam)
    rgb, st = core.render()
    ok, info = images_match(rgb, ref)
    plog(gate="philox_8spp", scene=name, res=[W, H], **info, extend=[int(st.extend_rays), int(cnt[0])], shadow=[int(st.shadow_rays), int(cnt[1])])
    assert ok, info
    assert abs(int(st.extend_rays) - int(cnt[0])) <= 2e-4 * cnt[0] + 2
    assert abs(int(st.shadow_rays) - int(cnt[1])) <= 2e-4 * cnt[1] + 8


@pytest.mark.gpu
@pytest.mark.parametrize("name", STANDINS)
def test_standin_image_gate_1024spp(name, core):
    """north_star gate 2 on the bench scene: per-channel RMSE < 1 % of mean radiance at 1024 spp.  The plain figure, the
    flipped-pixel count, the masked figure and the reference-vs-reference floor are all recorded."""
    fx = fixture("standin_" + name); cfg = STANDIN_CONFIGS[name]
    sc = standin(name)
    core.set_params(RMSE_SPP, cfg["nl"], cfg["depth"], 0)
    core.load(sc, camera=fx["ref_camera"])
    rgb, st = core.render()
    ref = fx["philox_rgb"].astype(np.float64)
    ok, info = images_match(rgb, ref, pixel_tol=0.05, max_bad_fraction=5e-4, rmse_tol=0.01)
    plain = np.sqrt(((rgb - ref) ** 2).mean(axis=(0, 1))) / ref.mean(axis=(0, 1))
    a, b = fx["ref_rgb"].astype(np.float64), fx["ref_rgb_b"].astype(np.float64)
    floor_px = np.sqrt(((a - b) ** 2).mean(axis=(0, 1))) / a.mean(axis=(0, 1))
    vs_ref_px = np.sqrt(((rgb - a) ** 2).mean(axis=(0, 1))) / a.mean(axis=(0, 1))
    k = 20
    floor = float(np.sqrt(((block_mean(a, k) - block_mean(b, k)) ** 2).mean()))
    got = float(np.sqrt(((block_mean(rgb.astype(np.float64), k) - block_mean(a, k)) ** 2).mean()))
    mean_dev = float(np.max(np.abs(rgb.mean(axis=(0, 1)) - a.mean(axis=(0, 1))) / a.mean(axis=(0, 1))))
    mean_floor = float(np.max(np.abs(a.mean(axis=(0, 1)) - b.mean(axis=(0, 1))) / a.mean(axis=(0, 1))))
    plog(gate="rmse_1024spp", scene=name, res=list(RMSE_RES), plain_rel_rmse_vs_philox_oracle=[float(x) for x in plain],
         flipped_pixels=info["n_bad"], masked_rel_rmse=info["rel_rmse_of_matching_pixels"],
         per_pixel_rel_rmse_vs_reference_rand=[float(x) for x in vs_ref_px], reference_vs_reference_per_pixel_rel_rmse=[float(x) for x in floor_px],
         block20_rmse_vs_reference=got, block20_rmse_reference_vs_reference=floor, mean_dev=mean_dev, mean_dev_reference_vs_reference=mean_floor,
         extend=[int(st.extend_rays), int(fx["philox_cnt"][0])], shadow=[int(st.shadow_rays), int(fx["philox_cnt"][1])])
    assert ok, info
    assert plain.max() < 0.01, plain          # north_star gate 2, literally: plain per-channel RMSE, no pixel excluded
    assert abs(int(st.extend_rays) - int(fx["philox_cnt"][0])) <= 2e-4 * fx["philox_cnt"][0]
    assert abs(int(st.shadow_rays) - int(fx["philox_cnt"][1])) <= 2e-4 * fx["philox_cnt"][1]
    assert mean_dev < 0.01 + 3 * mean_floor
    assert got < 1.6 * floor + 1e-4, (got, floor)


# ---- memory sizing / options -------------------------------------------------------------------------------------------------
@pytest.mark.gpu
def test_many_light_samples_and_small_budgets_still_render(core, golden):
    """-l 32 at 1080p (worst-case queues of the old fixed 16 M-path batch: > 180 GB) and an 8 GB wavefront budget both
    return the frame an unconstrained context renders (Philox: the image does not depend on the batching)."""
    g = golden("CBspheres_lambertian")
    cam = g["camera"].copy(); cam[14] *= 1080 / cam[13]; cam[12], cam[13] = 1920, 1080
    core.set_params(2, 32, 8, 3)
    core.load(g, camera=cam)
    rgb, st = core.render()
    assert np.isfinite(rgb).all() and rgb.mean() > 0 and st.shadow_rays % 32 == 0
    core.set_option("wavefront_budget_mb", 8000)        # one 1-spp batch with a pool of one batch = 6.7 GB at -l 32
    small, st2 = core.render()
    core.set_option("wavefront_budget_mb", 0)
    assert st2.batches > st.batches
    assert st2.extend_rays == st.extend_rays and st2.shadow_rays == st.shadow_rays
    assert np.allclose(small, rgb, rtol=1e-4, atol=1e-5 * rgb.mean())
    core.set_option("wavefront_budget_mb", 1)
    with pytest.raises(D.DsrtError, match="wavefront state"):
        core.render()
    core.set_option("wavefront_budget_mb", 0)


@pytest.mark.gpu
def test_mirror_box_every_path_survives(core):
    """Cornell box with perfect mirrors on every wall: Russian roulette never terminates (illum(f) >= 1), so the deep-path
    pool runs at its worst-case fill through all max_depth bounces.  Checked against the oracle at a small size and run at 1080p."""
    tp, tn, pb, bt, bpar, lt, lp = S.cornell_box()
    bt = bt.copy(); bpar = bpar.copy()
    for i in (0, 2, 3, 4, 5):
        bt[i] = 1; bpar[i, :3] = 1.0
    sc = {"prim_type": np.ones(12, np.int32), "prim_bsdf": pb, "tri_pos": tp, "tri_nrm": tn, "sphere": np.zeros((12, 4)),
          "bsdf_type": bt, "bsdf_param": bpar, "light_type": lt, "light_param": lp}
    cam = S.cam_dragon(96, 54)
    ref, cnt = O.Scene(dict(sc, camera=cam)).render(96, 54, 4, 2, 8, rng="philox", seed=2)
    core.set_params(4, 2, 8, 2)
    core.load(sc, camera=cam)
    rgb, st = core.render()
    ok, info = images_match(rgb, ref)
    assert ok, info
    assert abs(int(st.extend_rays) - int(cnt[0])) <= 4
    cam = S.cam_dragon(1920, 1080)
    core.set_params(16, 2, 8, 2)
    core.load(sc, camera=cam)
    big, stb = core.render()
    assert np.isfinite(big).all()
    # no path is terminated by roulette: the segments per camera sample match the oracle's small render (rays only end on the
    # light quad, through the open front, or at depth 8)
    per_sample_small = float(cnt[0]) / (96 * 54 * 4)
    assert abs(stb.extend_rays / stb.camera_samples - per_sample_small) < 0.05 * per_sample_small


@pytest.mark.gpu
def test_skip_null_shadow_changes_the_ray_count_not_the_image(core, golden):
    g = golden("CBgems"); cfg = CONFIGS["CBgems"]
    core.set_params(4, cfg["nl"], cfg["depth"], 5)
    core.load(g, camera=g["small_camera"])
    a, sa = core.render()
    core.set_option("skip_null_shadow", 1)
    b, sb = core.render()
    core.set_option("skip_null_shadow", 0)
    assert sa.null_shadow_rays == 0 and 0 < sb.null_shadow_rays < sb.shadow_rays
    assert sb.shadow_rays == sa.shadow_rays and sb.extend_rays == sa.extend_rays
    assert np.allclose(a, b, rtol=1e-5, atol=1e-6 * a.mean())


# ---- the compiled reference-side binding -----------------------------------------------------------------------------------
REF_GPU_DRIVER = os.path.join(O.REF_DIR, "ref_gpu_driver")


@pytest.mark.gpu
@pytest.mark.skipif(not os.path.exists(REF_GPU_DRIVER), reason="oracle/_ref/ref_gpu_driver is built by oracle/build_ref.sh")
@pytest.mark.parametrize("name", ["CBspheres_lambertian", "CBgems_cam", "bunny"])
def test_reference_pathtracer_drives_libdsrt_through_the_shim(name, tmp_path):
    """integration/cuda_path_tracer_shim.{h,cpp} (drop-in for cuda_src/setup.{h,cu}) compiled against the reference's real
    PathTracer: the reference's ColladaParser / Scene / BVHAccel feed libdsrt.so through the class Application::startGPURayTracing
    drives.  The frame it leaves in PathTracer::sampleBuffer equals the product's own host path (same seed, same Philox streams)."""
    import subprocess
    cfg = CONFIGS[name]
    W, H = SMALL_RES
    dae = O.ref_scene_path(cfg["file"]); cam = O.ref_scene_path(cfg["cam"]) if cfg["cam"] else None
    raw = tmp_path / "shim.f32"
    cmd = [REF_GPU_DRIVER, "-s", "4", "-l", str(cfg["nl"]), "-m", str(cfg["depth"]), "-w", str(W), "-h", str(H), "--seed", "5", "--raw", str(raw)]
    if cam:
        cmd += ["-f", cam]
    r = subprocess.run(cmd + [dae], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "GPU ray tracing done" in r.stdout
    got = np.fromfile(raw, np.float32).reshape(H, W, 3)
    want, st, _ = D.render_file(dae, W, H, 4, cfg["nl"], cfg["depth"], cam_info=cam, seed=5)
    assert f"segments {int(st.extend_rays)} + {int(st.shadow_rays)}" in r.stdout
    assert np.allclose(got, want, rtol=1e-4, atol=1e-5 * want.mean())


# ---- cancel / device tone map ------------------------------------------------------------------------------------------------
@pytest.mark.gpu
def test_cancel_stops_a_long_render_and_keeps_the_partial_frame(golden):
    """dsrt_cancel from another thread (PathTracer::stop, pathtracer.cpp:148-171): the render returns early with the mean
    of the samples done so far; an uncancelled render of the same context afterwards is complete again."""
    import threading
    import time
    g = golden("CBspheres_lambertian"); cfg = CONFIGS["CBspheres_lambertian"]
    cam = g["camera"].copy(); cam[14] *= 1080 / cam[13]; cam[12], cam[13] = 1920, 1080
    c = D.Core(0)
    c.set_params(4096, cfg["nl"], cfg["depth"], 1)
    c.load(g, camera=cam)
    ref, st_ref = c.render(spp_count=16)                 # 16 spp, normalised by ns_aa = 4096 -> rescale below
    t = threading.Timer(0.4, c.cancel)
    t0 = time.time(); t.start()
    rgb, st = c.render()
    dt = time.time() - t0; t.join()
    assert c.cancelled and 0 < st.camera_samples < 4096 * 1920 * 1080
    assert dt < 3.0, dt
    spp_done = st.camera_samples / (1920 * 1080)
    assert np.isfinite(rgb).all()
    assert abs(rgb.mean() - ref.mean() * 4096 / 16) < 0.05 * rgb.mean()      # normalised by the samples rendered
    again, st2 = c.render(spp_count=16)
    assert not c.cancelled and st2.camera_samples == st_ref.camera_samples
    assert np.allclose(again, ref, rtol=1e-4, atol=1e-6 * ref.mean())
    c.close()
    assert spp_done >= 1


@pytest.mark.gpu
def test_render_tonemapped_matches_toColor(core, golden):
    g = golden("CBgems"); cfg = CONFIGS["CBgems"]
    core.set_params(4, cfg["nl"], cfg["depth"], 5)
    core.load(g, camera=g["small_camera"])
    rgb, img, st = core.render(rgba8=True)
    plain, _ = core.render()
    assert np.allclose(rgb, plain, rtol=1e-5, atol=1e-6 * plain.mean())
    exp = O.to_color(rgb).reshape(img.shape)
    a = img.view(np.uint8).reshape(-1, 4).astype(int); b = exp.view(np.uint8).reshape(-1, 4).astype(int)
    assert np.abs(a - b).max() <= 1 and (a[:, 3] == 255).all()


# ---- device-side BVH construction (option "device_build") ----------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("name", ["CBspheres", "CBbunny", "CBempty", "CBgems"])
def test_device_built_bvh_gives_the_same_hits_and_images(name, core, golden):
    """LBVH + collapse on the GPU: a different tree, the same answers -- primary ids / t bit-exact against the reference's
    BVHAccel::intersect (gate 1 is topology independent), Philox-matched image equal to the oracle's, conservative boxes."""
    g = golden(name); cfg = CONFIGS[name]
    core.set_params(4, cfg["nl"], cfg["depth"], 5)
    core.load(g, camera=g["camera"], device_build=True)
    info = core.accel_info()
    assert info["wide_nodes"] >= 1 and info["max_depth"] >= 1
    ids, ts = core.primary_hits(mode=1)
    ok = ~g["hit_tie"].astype(bool)
    assert np.array_equal(ids[ok], g["hit_id"][ok]), f"{(ids != g['hit_id'])[ok].sum()} id mismatches"
    assert np.array_equal(ts[ok], g["hit_t"][ok])
    W, H = SMALL_RES
    core.set_camera(g["small_camera"])
    ref, cnt = O.Scene(g).with_camera(g["small_camera"]).render(W, H, 4, cfg["nl"], cfg["depth"], rng="philox", seed=5)
    rgb, st = core.render()
    okimg, info2 = images_match(rgb, ref)
    assert okimg, info2
    assert abs(int(st.extend_rays) - int(cnt[0])) <= 2e-4 * cnt[0] + 2 and abs(int(st.shadow_rays) - int(cnt[1])) <= 2e-4 * cnt[1] + 8
    core.set_option("device_build", 0)


@pytest.mark.gpu
def test_device_built_bvh_on_the_measured_scenes(core):
    fx = fixture("standin_cbdragon_standin"); sc = standin("cbdragon_standin")
    core.set_params(1, 4, 8, 0)
    core.load(sc, camera=fx["camera"], device_build=True)
    ids, ts = core.primary_hits(mode=1)
    assert np.array_equal(ids, fx["hit_id"]) and np.array_equal(ts, fx["hit_t"])
    fs = fixture("soup1m")
    W, H = STANDIN_ID_RES
    soup, cam = S.triangle_soup(1 << 20, W=W, H=H)
    core.set_params(1, 1, 8, 0)
    core.load(soup, camera=cam, device_build=True)
    ids, ts = core.primary_hits(mode=1)
    same = ids == fs["hit_id"]
    assert (~same).sum() <= 4 and np.array_equal(ts[same], fs["hit_t"][same])
    # degenerate inputs: one primitive, three primitives, many identical primitives (equal Morton codes)
    for n in (1, 3, 70):
        one = {k: (v[:1].repeat(n, axis=0) if k in ("prim_type", "prim_bsdf", "tri_pos", "tri_nrm", "sphere") else v) for k, v in soup.items()}
        core.load(one, camera=cam, device_build=True)
        assert core.accel_info()["wide_nodes"] >= 1
        rgb, st = core.render()
        assert np.isfinite(rgb).all()
    core.set_option("device_build", 0)


# ---- options that shape the wide BVH without changing a hit ("light_aligned_grid", "drop_coplanar_mates", "regroup_top") -----------
@pytest.mark.gpu
@pytest.mark.parametrize("device_build", [False, True])
@pytest.mark.parametrize("name", ["CBbunny", "CBspheres"])
def test_tree_options_change_the_fetch_counts_not_the_image(name, device_build, golden):
    """include/dsrt.h: the emissive quad under the area light and the other half of the wall quad a ray starts on drop out of
    the shadow rays' way (defaults), walls become direct children of the root (opt-in).  Same seed, same paths: the frames are
    equal up to the order of the float atomics, the segment counts are equal, the primitive tests per segment fall; the host
    builder and the device-side builder (its own grid code + k_db_mark_flat) both honour the options."""
    g = golden(name); cfg = CONFIGS[name]
    c = D.Core(0)
    try:
        c.set_params(4, cfg["nl"], cfg["depth"], 9)
        c.set_option("count_traversal", 1)
        out = {}
        for key, opts in (("off", (0, 0, 0)), ("defaults", (1, 1, 0)), ("regroup", (1, 1, 1))):
            for o, v in zip(("light_aligned_grid", "drop_coplanar_mates", "regroup_top"), opts):
                c.set_option(o, v)
            c.load(g, camera=g["small_camera"], device_build=device_build)
            rgb, st = c.render()
            out[key] = (np.array(rgb, copy=True), st.extend_rays, st.shadow_rays, st.prims_tested / st.segments, st.nodes_visited / st.segments)
        ref = out["off"]
        for key in ("defaults", "regroup"):
            rgb, ext, sh, pps, nps = out[key]
            assert (ext, sh) == (ref[1], ref[2])
            d = np.abs(rgb - ref[0])
            assert (d.max(axis=2) > 1e-5 * max(1.0, float(ref[0].max()))).mean() <= 1e-3, (key, d.max())
            assert pps < (0.95 if device_build else 0.85) * ref[3], (key, pps, ref[3])
        if not device_build:
            assert out["regroup"][4] <= out["defaults"][4]            # node visits per segment (the device tree has no regrouping)
    finally:
        c.close()


# ---- a failed call must not leave a half-updated context usable ------------------------------------------------------
@pytest.mark.gpu
def test_failed_scene_or_bvh_calls_invalidate_the_context(golden):
    g = golden("CBspheres")
    c = D.Core(0)
    c.set_params(1, 1, 2, 0)
    c.load(g, camera=g["small_camera"])
    good, _ = c.render()
    bad = dict(g); bad["bsdf_type"] = g["bsdf_type"].copy(); bad["bsdf_type"][-1] = 9
    with pytest.raises(D.DsrtError, match="unknown BSDF type"):
        c.set_scene(bad)
    with pytest.raises(D.DsrtError, match="dsrt_build_accel"):          # nothing of the old scene is usable any more
        c.render()
    with pytest.raises(D.DsrtError, match="scene"):
        c.set_bvh(D.build_bvh2(g))
    # a BVH whose child points back at its parent (would loop forever) / whose children do not partition the parent's range
    c.set_scene(g)
    bvh = D.build_bvh2(g)
    cyc = {k: v.copy() for k, v in bvh.items()}
    inner = int(np.argmax(cyc["node_left"] >= 0)); cyc["node_left"][inner] = inner
    c.set_bvh(cyc)
    with pytest.raises(D.DsrtError, match="not numbered after its parent"):
        c.build_accel()
    part = {k: v.copy() for k, v in bvh.items()}
    part["node_range"][int(part["node_left"][inner])] += 1
    c.set_bvh(part)
    with pytest.raises(D.DsrtError, match="partition"):
        c.build_accel()
    c.set_bvh(bvh); c.build_accel(); c.set_camera(g["small_camera"])
    again, _ = c.render()
    assert np.allclose(again, good, rtol=1e-5, atol=1e-7)
    c.close()


@pytest.mark.gpu
def test_window_tiles_add_up_to_the_frame(core, golden):
    """dsrt_set_window (tile partitioning, the alternative to the sample split): disjoint windows rendered separately are
    zero outside their window and add up to the full frame; counters add up too."""
    g = golden("CBgems"); cfg = CONFIGS["CBgems"]
    W, H = SMALL_RES
    core.set_params(4, cfg["nl"], cfg["depth"], 3)
    core.load(g, camera=g["small_camera"])
    full, st = core.render()
    tiles = [(0, 0, 40, 30), (40, 0, W - 40, 30), (0, 30, W, 17), (0, 47, 13, H - 47), (13, 47, W - 13, H - 47)]    # ragged on purpose
    total = np.zeros_like(full); ext = sh = cam = 0
    for (x0, y0, w, h) in tiles:
        core.set_window(x0, y0, w, h)
        part, sp = core.render()
        mask = np.zeros((H, W), bool); mask[y0:y0 + h, x0:x0 + w] = True
        assert (part[~mask] == 0).all() and sp.camera_samples == w * h * 4
        total += part; ext += sp.extend_rays; sh += sp.shadow_rays; cam += sp.camera_samples
    core.set_window()
    assert cam == st.camera_samples and ext == st.extend_rays and sh == st.shadow_rays
    assert np.allclose(total, full, rtol=1e-5, atol=1e-7)
    with pytest.raises(D.DsrtError, match="outside the frame"):
        core.set_window(W - 4, 0, 8, 8)
    again, _ = core.render()
    assert np.allclose(again, full, rtol=1e-5, atol=1e-7)
