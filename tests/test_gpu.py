"""GPU parity tests (run on the B200 box: python -m pytest tests -m gpu).  Everything goes through the C ABI
(include/dsrt.h) of libdsrt.so; the oracle is only the checker."""
import os

import numpy as np
import pytest

import dsgpuraytracing_b200 as D
from oracle import oracle as O
from tests.cpuwalk import Walk
from tests.scenes import CONFIGS, ID_RES, RMSE_RES, RMSE_SPP, SMALL_RES
from tests.util import images_match, plog

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
SCENES = list(CONFIGS)


def setup_core(core, g, cam, nl, depth, spp, seed=0):
    core.set_params(spp, nl, depth, seed)
    core.load(g, camera=cam)      # set_scene + host SAH build (product code) + build_accel + set_camera


def block_mean(a, k):
    H, W, _ = a.shape
    return a[:H // k * k, :W // k * k].reshape(H // k, k, W // k, k, 3).mean(axis=(1, 3))


# ---- gate 1: primary-ray closest-hit primitive ids, bit-exact vs BVHAccel::intersect (exact ties excluded)
@pytest.mark.parametrize("name", SCENES)
def test_primary_hit_ids_bit_exact(name, core, golden):
    g = golden(name)
    setup_core(core, g, g["camera"], CONFIGS[name]["nl"], CONFIGS[name]["depth"], 1)
    ids, ts = core.primary_hits(mode=1)
    ok = ~g["hit_tie"].astype(bool)
    assert np.array_equal(ids[ok], g["hit_id"][ok]), f"{(ids != g['hit_id'])[ok].sum()} id mismatches"
    assert np.array_equal(ts[ok], g["hit_t"][ok])            # t is bit-exact too (fp64 leaf tests, reference op order)
    # production float kernel (watertight test): silhouette pixels may flip; report and bound
    ids0, ts0 = core.primary_hits(mode=0)
    flips = int((ids0 != g["hit_id"]).sum())
    plog(gate="ids", scene=name, res=list(ID_RES), parity_kernel_mismatches=0, ties=int(g["hit_tie"].sum()),
         production_float_kernel_mismatches=flips, pixels=int(ids0.size))
    assert flips <= 2, flips          # measured: 0 on every fixture scene (float watertight test vs the reference's fp64 Moller-Trumbore)
    same = (ids0 == g["hit_id"]) & (g["hit_id"] >= 0)
    assert np.max(np.abs(ts0[same] - g["hit_t"][same]) / g["hit_t"][same]) < 1e-3


# ---- gate 2a: same Philox streams as the oracle -> same paths, same segment counts, same image
@pytest.mark.parametrize("name", SCENES)
def test_render_matches_oracle_small(name, core, golden):
    g = golden(name); cfg = CONFIGS[name]
    cam = g["small_camera"]; W, H = SMALL_RES
    ref, cnt = O.Scene(g).with_camera(cam).render(W, H, 4, cfg["nl"], cfg["depth"], rng="philox", seed=5)
    setup_core(core, g, cam, cfg["nl"], cfg["depth"], 4, seed=5)
    rgb, st = core.render()
    assert st.camera_samples == W * H * 4
    assert abs(int(st.extend_rays) - int(cnt[0])) <= 2e-4 * cnt[0] + 2
    assert abs(int(st.shadow_rays) - int(cnt[1])) <= 2e-4 * cnt[1] + 8
    ok, info = images_match(rgb, ref)
    assert ok, info


# ---- gate 2b: the image gate at 1024 spp.  Tolerance: per-channel RMSE < 1 % of mean radiance (north_star),
# evaluated against the oracle driven by the same Philox streams (committed fixture), plus a statistical check
# against the reference's own rand()-driven 1024-spp render (mean within 1 %, block RMSE within the noise floor).
@pytest.mark.parametrize("name", SCENES)
def test_image_gate_1024spp(name, core, golden):
    g = golden(name); cfg = CONFIGS[name]
    ph = np.load(os.path.join(GOLDEN, name + "_philox.npz"))
    setup_core(core, g, g["ref_camera"], cfg["nl"], cfg["depth"], RMSE_SPP, seed=0)
    rgb, st = core.render()
    ref = ph["philox_rgb"].astype(np.float64)
    # The 1 % gate.  A float-vs-double flip of ONE discrete path decision (a roulette draw at its threshold, a
    # silhouette hit) can add or remove a firefly worth hundreds of radiance units in one of 19.7 M samples, which alone
    # moves a plain RMSE by >1 %; such pixels are counted (at most 5 in 10 000 allowed) and excluded from the RMSE.
    ok, info = images_match(rgb, ref, pixel_tol=0.05, max_bad_fraction=5e-4, rmse_tol=0.01)
    assert ok, info
    plain = np.sqrt(((rgb - ref) ** 2).mean(axis=(0, 1))) / np.maximum(ref.mean(axis=(0, 1)), 1e-12)
    print(f"{name}: plain rel RMSE {plain.max():.2e}, flipped pixels {info['n_bad']}, RMSE of the rest {info['rel_rmse_of_matching_pixels']:.2e}")
    assert plain.max() < 0.01, plain          # north_star gate 2, literally: no pixel excluded (measured 4e-6 ... 1.5e-3, profiles/r2_parity.md)
    assert abs(int(st.extend_rays) - int(ph["philox_cnt"][0])) <= 2e-4 * ph["philox_cnt"][0]
    assert abs(int(st.shadow_rays) - int(ph["philox_cnt"][1])) <= 2e-4 * ph["philox_cnt"][1]
    # statistical agreement with the compiled reference (different random source)
    a, b = g["ref_rgb"].astype(np.float64), g["ref_rgb_b"].astype(np.float64)
    assert np.max(np.abs(rgb.mean(axis=(0, 1)) - a.mean(axis=(0, 1))) / a.mean(axis=(0, 1))) < 0.01 + 3 * np.max(
        np.abs(a.mean(axis=(0, 1)) - b.mean(axis=(0, 1))) / a.mean(axis=(0, 1)))
    k = 20
    floor = np.sqrt(((block_mean(a, k) - block_mean(b, k)) ** 2).mean())
    got = np.sqrt(((block_mean(rgb.astype(np.float64), k) - block_mean(a, k)) ** 2).mean())
    am = np.maximum(a.mean(axis=(0, 1)), 1e-12)
    plog(gate="rmse_1024spp", scene=name, res=list(RMSE_RES), plain_rel_rmse_vs_philox_oracle=[float(x) for x in plain],
         flipped_pixels=info["n_bad"], masked_rel_rmse=info["rel_rmse_of_matching_pixels"],
         per_pixel_rel_rmse_vs_reference_rand=[float(x) for x in np.sqrt(((rgb - a) ** 2).mean(axis=(0, 1))) / am],
         reference_vs_reference_per_pixel_rel_rmse=[float(x) for x in np.sqrt(((a - b) ** 2).mean(axis=(0, 1))) / am],
         block20_rmse_vs_reference=float(got), block20_rmse_reference_vs_reference=float(floor),
         extend=[int(st.extend_rays), int(ph["philox_cnt"][0])], shadow=[int(st.shadow_rays), int(ph["philox_cnt"][1])])
    assert got < 1.6 * floor + 1e-4, (got, floor)


# ---- size-independent properties -------------------------------------------------------------------------------
def test_sample_split_is_linear_and_batch_invariant(core, golden):
    """The multi-GPU decomposition: samples k = r (mod g) rendered separately add up to the full render, and the
    result does not depend on the wavefront batch size (Philox counters, not execution order, define the samples)."""
    g = golden("CBgems"); cfg = CONFIGS["CBgems"]
    setup_core(core, g, g["small_camera"], cfg["nl"], cfg["depth"], 8, seed=3)
    full, st = core.render()
    parts = [core.render(spp_begin=r, spp_count=2, spp_stride=4)[0] for r in range(4)]
    assert np.allclose(np.sum(parts, axis=0), full, rtol=1e-4, atol=1e-5 * full.mean())
    core.set_option("batch_spp", 1)
    one, st1 = core.render()
    core.set_option("batch_spp", 0)
    assert st1.batches == 8 and st1.extend_rays == st.extend_rays and st1.shadow_rays == st.shadow_rays
    assert np.allclose(one, full, rtol=1e-4, atol=1e-5 * full.mean())
    again, _ = core.render()
    assert np.allclose(again, full, rtol=1e-5, atol=1e-6 * full.mean())   # reproducible up to float atomics order


def test_full_size_frame_properties(core, golden):
    """BASELINE config size (1920x1080): sample-split linearity and segment accounting at full resolution."""
    g = golden("CBbunny"); cfg = CONFIGS["CBbunny"]
    cam = g["camera"].copy(); cam[14] *= 1080 / cam[13]; cam[12], cam[13] = 1920, 1080
    setup_core(core, g, cam, cfg["nl"], cfg["depth"], 4, seed=1)
    full, st = core.render()
    assert st.camera_samples == 1920 * 1080 * 4
    assert st.extend_rays >= st.camera_samples and st.shadow_rays % 4 == 0
    halves = core.render(0, 2, 2)[0] + core.render(1, 2, 2)[0]
    assert np.allclose(halves, full, rtol=1e-4, atol=1e-5 * full.mean())
    assert np.isfinite(full).all() and full.min() >= 0


def test_ragged_frame_and_odd_sizes(core, golden):
    g = golden("CBspheres"); cfg = CONFIGS["CBspheres"]
    for (W, H) in [(37, 23), (8, 4), (1, 1), (130, 3)]:
        cam = g["camera"].copy(); cam[14] *= H / cam[13]; cam[12], cam[13] = W, H
        ref, cnt = O.Scene(g).with_camera(cam).render(W, H, 2, cfg["nl"], cfg["depth"], rng="philox", seed=9)
        setup_core(core, g, cam, cfg["nl"], cfg["depth"], 2, seed=9)
        rgb, st = core.render()
        assert st.camera_samples == W * H * 2
        assert abs(int(st.extend_rays) - int(cnt[0])) <= 2 and abs(int(st.shadow_rays) - int(cnt[1])) <= 8
        assert np.allclose(rgb, ref, rtol=5e-2, atol=5e-3 * ref.mean() + 1e-6) or \
            np.sqrt(((rgb - ref) ** 2).mean()) < 5e-3 * ref.mean()


def test_empty_scene_and_errors(core):
    base = dict(bsdf_type=np.zeros(1, np.int32), bsdf_param=np.zeros((1, 8), np.float32), light_type=np.zeros(0, np.int32),
                light_param=np.zeros((0, 28)), prim_type=np.zeros(0, np.int32), prim_bsdf=np.zeros(0, np.int32),
                tri_pos=np.zeros((0, 9)), tri_nrm=np.zeros((0, 9)), sphere=np.zeros((0, 4)))
    cam = np.array([0, 0, 3, 1, 0, 0, 0, 1, 0, 0, 0, 1, 16, 8, 10, 0, 0], float)
    core.set_params(1, 1, 2, 0)
    core.load(base, camera=cam)
    rgb, st = core.render()
    assert rgb.sum() == 0 and st.extend_rays == 16 * 8 and st.shadow_rays == 0
    ids, _ = core.primary_hits(1)
    assert (ids == -1).all()
    # errors are status codes + messages, never exit() (the reference exits, cuda_src/setup.cu:139-143)
    c2 = D.Core(0)
    with pytest.raises(D.DsrtError, match="dsrt_set_camera|dsrt_build_accel"):
        c2.width = c2.height = 4
        c2.render()
    with pytest.raises(D.DsrtError, match="prim_bsdf"):
        bad = dict(base, prim_type=np.ones(1, np.int32), prim_bsdf=np.array([5], np.int32), tri_pos=np.zeros((1, 9)),
                   tri_nrm=np.zeros((1, 9)), sphere=np.zeros((1, 4)))
        c2.set_scene(bad)
    c2.close()


@pytest.mark.parametrize("name", ["CBcoil", "CBspheres"])
def test_arbitrary_rays_closest_any_and_fetch_counters(name, core, golden):
    """Incoherent random rays: closest / any hit vs the oracle's BVHAccel restatement; node / primitive fetch
    counters vs the CPU walk of the same flattened wide BVH (bytes per segment accounting)."""
    g = golden(name)
    rng = np.random.default_rng(1)
    n = 4000
    lo = g["node_bbox"][0, :3]; hi = g["node_bbox"][0, 3:]
    o = (lo + (hi - lo) * rng.random((n, 3))).astype(np.float32)
    d = rng.normal(size=(n, 3)); d = (d / np.linalg.norm(d, axis=1, keepdims=True)).astype(np.float32)
    setup_core(core, g, g["camera"], 4, 2, 1)
    ids, ts = core.trace_closest(o, d)
    sc = O.Scene(g)
    bad = 0
    for i in range(n):
        pid, t, _ = sc.closest_hit(o[i].astype(np.float64), d[i].astype(np.float64))
        if pid != ids[i]:
            bad += 1
        elif pid >= 0:
            assert abs(ts[i] - t) <= 1e-3 * t + 1e-5
    assert bad <= 4, bad          # float vs double on silhouettes
    tmax = (rng.random(n) * np.linalg.norm(hi - lo)).astype(np.float32)
    anyh = core.trace_any(o, d, tmax)
    badany = sum(int(sc.any_hit(o[i].astype(np.float64), d[i].astype(np.float64), float(tmax[i])) != bool(anyh[i])) for i in range(n))
    assert badany <= 6, badany
    w = Walk(g, D.build_bvh2(g), 4)
    wid, wt, wcnt = w.trace(o, d)
    assert (wid != ids).sum() <= 2
    core.set_option("count_traversal", 1)
    core.set_option("postpone_min_lanes", 0)          # test primitives at once: the order the CPU walk uses
    core.set_option("coop_min_pairs", 1 << 20)        # ... one lane per ray (no cooperative any-hit tests)
    core.set_params(1, 4, 0, 0)
    # counters of a full render == CPU walk counters of the same wavefront (same code, same rays)
    cam = g["small_camera"]; core.set_camera(cam)
    rgb, st = core.render()
    wr, wc = Walk(g, D.build_bvh2(g), 4, camera=cam).render(1, 0, seed=0)
    assert abs(int(st.nodes_visited) - int(wc[3])) <= 2e-3 * wc[3]
    assert abs(int(st.prims_tested) - int(wc[4])) <= 2e-3 * wc[4]
    # postponed primitive tests (the default) change the visiting order, never the result
    core.set_option("postpone_min_lanes", 8)
    core.set_option("coop_min_pairs", 6)
    rgb2, st2 = core.render()
    core.set_option("count_traversal", 0)
    assert st2.extend_rays == st.extend_rays and st2.shadow_rays == st.shadow_rays
    assert np.allclose(rgb2, rgb, rtol=1e-4, atol=1e-6)
    assert st2.nodes_visited >= st.nodes_visited * 0.98 and st2.nodes_visited < st.nodes_visited * 1.5


def test_tonemap_matches_toColor(core):
    rng = np.random.default_rng(0)
    rgb = (rng.random((64, 3)) ** 3 * 2).astype(np.float32)
    rgb[0] = 0; rgb[1] = 100
    got = core.tonemap(rgb)
    exp = O.to_color(rgb)
    gb = got.view(np.uint8).reshape(-1, 4).astype(int); eb = exp.view(np.uint8).reshape(-1, 4).astype(int)
    assert np.abs(gb - eb).max() <= 1      # powf rounding may move a channel by one LSB
    assert (gb[:, 3] == 255).all()


# ---- the C++ host side on the GPU: PathTracer class, CLI, multi-GPU context -------------------------------------------
needs_scenes = pytest.mark.skipif(not os.path.exists(O.ref_scene_path("CBspheres_lambertian.dae")),
                                  reason="reference .dae scenes are staged by oracle/build_ref.sh (not committed)")


@needs_scenes
@pytest.mark.parametrize("name", ["CBspheres", "CBgems_cam", "bunny"])
def test_pathtracer_class_render_file(name, golden, tmp_path):
    """.dae + cam .info -> C++ loader -> PathTracer::set_scene/build_accel/start_raytracing -> frame == oracle."""
    g = golden(name); cfg = CONFIGS[name]
    W, H = SMALL_RES
    cam = O.ref_scene_path(cfg["cam"]) if cfg["cam"] else None
    png = str(tmp_path / "o.png")
    rgb, st, secs = D.render_file(O.ref_scene_path(cfg["file"]), W, H, 4, cfg["nl"], cfg["depth"], cam_info=cam, seed=5, png=png)
    camv = g["small_camera"] if cfg["cam"] is None else D.load_dae(O.ref_scene_path(cfg["file"]), W, H, cam)[1]
    ref, cnt = O.Scene(g).with_camera(camv).render(W, H, 4, cfg["nl"], cfg["depth"], rng="philox", seed=5)
    ok, info = images_match(rgb, ref)
    assert ok, info
    assert abs(int(st.extend_rays) - int(cnt[0])) <= 2e-4 * cnt[0] + 2
    from PIL import Image
    im = np.asarray(Image.open(png))
    assert im.shape == (H, W, 4)
    exp = O.to_color(rgb)[::-1].view(np.uint8).reshape(H, W, 4)       # save_image flips rows (pathtracer.cpp:666-668)
    assert np.abs(im.astype(int) - exp.astype(int)).max() <= 1


@needs_scenes
def test_cli_binary(tmp_path):
    import subprocess
    exe = os.path.join(os.path.dirname(D.lib_path()), "pathtracer")
    raw = tmp_path / "f.raw"; png = tmp_path / "f.png"
    r = subprocess.run([exe, "-s", "4", "-l", "4", "-m", "5", "-w", "96", "-h", "72", "-S", "5", "-j", "-o", str(png), "-r", str(raw),
                        O.ref_scene_path("CBspheres_lambertian.dae")], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "GPU ray tracing done" in r.stdout
    import json
    line = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][-1])          # -j: one JSON line of run statistics
    assert line["camera_samples"] == 96 * 72 * 4 and line["extend_rays"] >= line["camera_samples"] and line["mrays_per_s"] > 0
    rgb = np.fromfile(raw, np.float32).reshape(72, 96, 3)
    ref = np.load(os.path.join(GOLDEN, "CBspheres_lambertian.npz"))
    sc = O.Scene({k: ref[k] for k in ref.files}).with_camera(ref["small_camera"])
    exp, _ = sc.render(96, 72, 4, 4, 5, rng="philox", seed=5)
    ok, info = images_match(rgb, exp)
    assert ok, info
    r = subprocess.run([exe, "-c", O.ref_scene_path("CBspheres_lambertian.dae")], capture_output=True, text=True)
    assert r.returncode != 0 and "no CPU fallback" in r.stderr
    # -O NAME=INT hands options to the GPU core: skipped null shadow rays are not counted as traced, an unknown name is an error
    base = ["-s", "2", "-l", "4", "-m", "5", "-w", "96", "-h", "72", "-S", "5", "-j", "-o", str(png), O.ref_scene_path("CBspheres_lambertian.dae")]
    r = subprocess.run([exe, "-O", "skip_null_shadow=1", "-O", "regroup_top=1"] + base, capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    r0 = subprocess.run([exe] + base, capture_output=True, text=True, timeout=120)
    j1 = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][-1]); j0 = json.loads([l for l in r0.stdout.splitlines() if l.startswith("{")][-1])
    assert j1["extend_rays"] == j0["extend_rays"] and j1["shadow_rays"] <= j0["shadow_rays"]
    r = subprocess.run([exe, "-O", "no_such_option=1"] + base, capture_output=True, text=True, timeout=120)
    assert r.returncode != 0 and "dsrt_set_option" in r.stderr


def test_multi_device_context_matches_single(golden):
    """dsrt_create_multi: samples split k = r (mod G), partial framebuffers combined by device 0 reading peer memory.
    With one GPU in the box the same code path runs with the device listed twice."""
    import torch
    n = torch.cuda.device_count()
    devs = [0, 1] if n >= 2 else [0, 0]
    g = golden("CBgems"); cfg = CONFIGS["CBgems"]
    one = D.Core(0); one.set_params(8, cfg["nl"], cfg["depth"], 4); one.load(g, camera=g["small_camera"])
    a, sa = one.render(); one.close()
    two = D.Core(devices=devs); two.set_params(8, cfg["nl"], cfg["depth"], 4); two.load(g, camera=g["small_camera"])
    b, sb = two.render(); two.close()
    assert sb.camera_samples == sa.camera_samples and sb.extend_rays == sa.extend_rays and sb.shadow_rays == sa.shadow_rays
    assert np.allclose(a, b, rtol=1e-4, atol=1e-5 * a.mean())


@pytest.mark.parametrize("name", ["env_CBspheres", "env_bunny", "env_CBgems"])
def test_environment_light_gpu(name, core, golden):
    """dsrt_set_envmap: EnvironmentLight::sample_L as a light + sample_dir on rays that leave the scene."""
    g = golden(name); depth = int(g["depth"]); W, H = SMALL_RES
    ref, cnt = O.Scene(g).render(W, H, 8, 4, depth, rng="philox", seed=21)
    keep = g["light_type"] != 4
    arr = dict(g); arr["light_type"] = g["light_type"][keep]; arr["light_param"] = g["light_param"][keep]
    core.set_params(8, 4, depth, 21)
    core.set_envmap(g["env_rgb"])
    core.load(arr, camera=g["camera"])
    rgb, st = core.render()
    core.set_envmap(None)
    assert abs(int(st.extend_rays) - int(cnt[0])) <= 2e-4 * cnt[0] + 2
    assert abs(int(st.shadow_rays) - int(cnt[1])) <= 2e-4 * cnt[1] + 8
    ok, info = images_match(rgb, ref)
    assert ok, info


@needs_scenes
def test_cli_environment_map_from_exr(tmp_path, core):
    """`pathtracer -e map.exr` (the reference's dead -e option, main.cpp:99-101): the file goes through the host OpenEXR
    reader into dsrt_set_envmap; same frame as handing the decoded array to the C ABI directly."""
    import subprocess
    exr = os.path.join(GOLDEN, "exr", "zip_half.exr")
    exe = os.path.join(os.path.dirname(D.lib_path()), "pathtracer")
    raw = tmp_path / "e.raw"; png = tmp_path / "e.png"
    dae = O.ref_scene_path("CBspheres_lambertian.dae")
    r = subprocess.run([exe, "-s", "4", "-l", "2", "-m", "3", "-w", "96", "-h", "72", "-S", "9", "-e", exr, "-o", str(png), "-r", str(raw), dae],
                       capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    got = np.fromfile(raw, np.float32).reshape(72, 96, 3)
    arr, cam = D.load_dae(dae, 96, 72)
    core.set_params(4, 2, 3, 9)
    core.set_envmap(D.load_envmap(exr))
    core.load(arr, camera=cam)
    want, _ = core.render()
    core.set_envmap(None)
    assert np.isfinite(got).all() and got.mean() > 0
    assert np.allclose(got, want, rtol=1e-4, atol=1e-5 * want.mean())


def test_read_bandwidth_probe(core):
    """dsrt_measure_read_bandwidth: the L2-resident sweep must beat the HBM sweep, and both must be plausible for a B200."""
    l2 = core.measure_read_bandwidth(32 << 20, 20)
    hbm = core.measure_read_bandwidth(1 << 30, 2)
    assert 1000.0 < hbm < 9000.0, hbm
    assert hbm < l2 < 40000.0, (hbm, l2)
    with pytest.raises(D.DsrtError):
        core.measure_read_bandwidth(8, 1)
