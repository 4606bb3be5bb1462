"""Pins the oracle (oracle/pt_oracle.c) to the reference.

The reference ships no golden vectors for this path, so the pins are outputs of the
reference's OWN CPU code compiled in the build container (oracle/_ref/ref_driver, see oracle/build_ref.sh),
committed under tests/golden/ by tests/golden/make_golden.py.  Everything here is bit-exact.
"""
import numpy as np
import pytest

from oracle import oracle as O
from tests.scenes import CONFIGS, ID_RES, SMALL_RES

SCENES = list(CONFIGS)


@pytest.mark.parametrize("name", SCENES)
def test_sah_bvh_topology_matches_reference(name, golden):
    g = golden(name)
    sc = O.Scene({k: g[k] for k in O.SCENE_KEYS})
    b = sc.build_bvh()
    for k in O.BVH_KEYS:
        assert b[k].shape == g[k].shape, k
        assert np.array_equal(b[k], g[k]), f"{name}: {k} differs from the reference's BVHAccel"


@pytest.mark.parametrize("name", SCENES)
def test_primary_hit_ids_match_reference(name, golden):
    g = golden(name)
    sc = O.Scene(g)
    W, H = ID_RES
    ids, ts, _ = sc.primary_hits(W, H, ties=False)
    assert np.array_equal(ids, g["hit_id"])
    assert np.array_equal(ts, g["hit_t"])          # bit-exact t, inf on miss


@pytest.mark.parametrize("name", SCENES)
def test_rand_driven_render_is_bit_exact(name, golden):
    """Same srand(1) seed, glibc rand() drawn in the reference's call order -> identical float image and
    identical segment counts (BVHAccel::intersect call counters)."""
    g = golden(name)
    cfg = CONFIGS[name]
    W, H = SMALL_RES
    sc = O.Scene(g).with_camera(g["small_camera"])
    rgb, cnt = sc.render(W, H, 2, cfg["nl"], cfg["depth"], rng="rand", seed=1)
    assert np.array_equal(cnt[:2], g["small_cnt"]), (cnt, g["small_cnt"])
    assert np.array_equal(rgb, g["small_rgb"])


def test_config1_mean_radiance(golden):
    """config 1 mean RGB 0.1464/0.1272/0.1462 (measured with the compiled reference)."""
    g = golden("CBspheres_lambertian")
    m = g["ref_rgb"].mean(axis=(0, 1))
    assert np.allclose(m, [0.1464, 0.1272, 0.1462], atol=2e-3)


def test_tie_mask_is_brute_force(golden):
    g = golden("CBgems")
    sc = O.Scene(g)
    W, H = 32, 24
    cam = g["camera"].copy(); cam[12], cam[13] = W, H
    cam[14] = g["camera"][14] * H / g["camera"][13]
    sc = sc.with_camera(cam)
    ids, ts, tie = sc.primary_hits(W, H, ties=True)
    for y in range(0, H, 5):
        for x in range(0, W, 5):
            o, d = O.generate_ray(cam, (x + .5) / W, (y + .5) / H)
            pid, t = sc.closest_hit_brute(o, d)
            assert pid == ids[y, x] or tie[y, x]


@pytest.mark.skipif(not O.have_reference(), reason="oracle/_ref not built (no /root/reference here)")
def test_live_reference_agrees_with_golden(golden):
    """When the compiled reference is available, re-run it and check the committed fixture is what it produces."""
    g = golden("CBspheres_lambertian")
    W, H = SMALL_RES
    out = O.run_reference(O.ref_scene_path("CBspheres_lambertian.dae"), W, H, spp=2, nl=4, depth=5, seed=1, render=True)
    assert np.array_equal(out["rgb"], g["small_rgb"])


def test_known_answers():
    # make_coord_space: z = n/|n|, orthonormal, right-handed (bsdf.cpp:13-30)
    for n in ([0, 0, 2.0], [0.3, -0.2, 0.9], [1, 0, 0], [0, -3, 0]):
        M = O.make_coord_space(n)
        assert np.allclose(M.T @ M, np.eye(3), atol=1e-12)
        assert np.allclose(M[:, 2], np.array(n) / np.linalg.norm(n))
    # Camera::generate_ray at the image centre looks along -c2w[2] from pos + c2w[2] (camera.cpp:113-129)
    cam = np.array([1, 2, 3, 1, 0, 0, 0, 1, 0, 0, 0, 1, 640, 480, 500, 0, 0], float)
    o, d = O.generate_ray(cam, 0.5, 0.5)
    assert np.allclose(o, [1, 2, 4]) and np.allclose(d, [0, 0, -1])
    # Philox4x32-10 known answers (Random123 kat_vectors, philox4x32 10 rounds: counter[4], key[2] -> output[4]); the oracle's
    # generator and the product's (rng.cuh, host build in tests/cpu_walk) must both reproduce them, and each other
    import ctypes as C
    from tests import cpuwalk
    kat = [((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
           ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
           ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0), (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]
    for ctr, key, want in kat:
        assert tuple(int(x) for x in O.philox_raw(ctr, *key)) == want
        c = np.array(ctr, np.uint32); out = np.zeros(4, np.uint32)
        cpuwalk.lib().cw_philox_raw(C.c_void_p(c.ctypes.data), C.c_uint32(key[0]), C.c_uint32(key[1]), C.c_void_p(out.ctypes.data))
        assert tuple(int(x) for x in out) == want
    assert tuple(O.philox(7, 11, 13, 2, 1)) == tuple(O.philox_raw((11, 13, 2, 1), 7, 0x5EED))      # key = (seed, 0x5EED), counter = (pixel, sample, depth, block)
    # toColor: (s*sqrt2)^(1/2.2) clamp, *255 truncating, ABGR packing (image.h:49-58,174-189)
    c = O.to_color(np.array([[0.0, 0.5, 10.0]], np.float32))
    assert c[0] == (0 | (int((0.5 * 2 ** 0.5) ** (1 / 2.2) * 255) << 8) | (255 << 16) | (255 << 24))


ENV_CASES = ["env_CBspheres", "env_bunny", "env_CBgems"]


@pytest.mark.parametrize("name", ENV_CASES)
def test_environment_light_is_bit_exact(name, golden):
    """EnvironmentLight (importance sampling tables, sample_L, sample_dir on miss; environment_light.cpp:6-201) with a
    procedural map: the oracle port reproduces the compiled reference's rand()-driven image bit for bit."""
    g = golden(name)
    W, H = SMALL_RES
    sc = O.Scene(g)
    assert sc.light_type[-1] == 4
    rgb, cnt = sc.render(W, H, 2, 4, int(g["depth"]), rng="rand", seed=1)
    assert np.array_equal(cnt[:2], g["small_cnt"])
    assert np.array_equal(rgb, g["small_rgb"])
