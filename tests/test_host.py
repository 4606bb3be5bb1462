"""CPU-side tests of the product's host code: C-ABI surface, host SAH builder, wide-BVH flattening.
No GPU is needed (the CUDA entry points are only checked for presence / clean failure)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import dsgpuraytracing_b200 as D
from oracle import oracle as O
from tests.cpuwalk import Walk
from tests.scenes import CONFIGS, ID_RES, SMALL_RES
from tests.util import images_match

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "dsrt.h")).read()
    declared = sorted(set(re.findall(r"\b(dsrt_[a-z0-9_]+)\s*\(", hdr)))
    assert declared == sorted(D.EXPORTED_SYMBOLS)
    L = D.load_library()
    for s in declared:
        assert hasattr(L, s), s
    assert b"dsrt" in L.dsrt_version()


def test_no_cpu_fallback_without_gpu():
    """On a box without a GPU dsrt_create must fail loudly; with one it must succeed."""
    import torch
    if torch.cuda.is_available():
        c = D.Core(0); c.close()
    else:
        with pytest.raises(D.DsrtError):
            D.Core(0)


def test_product_never_imports_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "dsgpuraytracing_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                assert "oracle" not in open(os.path.join(dirpath, f)).read().replace("CPU oracle (oracle/pt_oracle.c", ""), f


@pytest.mark.parametrize("name", list(CONFIGS))
def test_host_sah_builder_matches_reference_topology(name, golden):
    g = golden(name)
    b = D.build_bvh2(g)
    for k in O.BVH_KEYS:
        assert np.array_equal(b[k], g[k]), f"{name}: {k}"


def test_host_builder_edge_cases():
    base = dict(bsdf_type=[0], bsdf_param=np.zeros((1, 8), np.float32), light_type=np.zeros(0, np.int32),
                light_param=np.zeros((0, 28)))
    # empty scene
    e = dict(base, prim_type=np.zeros(0, np.int32), prim_bsdf=np.zeros(0, np.int32), tri_pos=np.zeros((0, 9)),
             tri_nrm=np.zeros((0, 9)), sphere=np.zeros((0, 4)))
    b = D.build_bvh2(e)
    assert len(b["node_start"]) == 1 and b["node_range"][0] == 0
    # many identical triangles: no finite split -> the reference leaves an oversized leaf (bvh.cpp:143-173)
    n = 9
    tri = np.tile(np.array([0, 0, 0, 1, 0, 0, 0, 1, 0], float), (n, 1))
    s = dict(base, prim_type=np.ones(n, np.int32), prim_bsdf=np.zeros(n, np.int32), tri_pos=tri,
             tri_nrm=np.tile(np.array([0, 0, 1] * 3, float), (n, 1)), sphere=np.zeros((n, 4)))
    b = D.build_bvh2(s)
    ob = O.Scene(s | {"camera": np.zeros(17)}).build_bvh()
    for k in O.BVH_KEYS:
        assert np.array_equal(b[k], ob[k]), k
    w = Walk(s, b, 1)
    assert sorted(w.slot_prim()) == list(range(n))


@pytest.mark.parametrize("name", list(CONFIGS))
def test_wide_bvh_cpu_walk_primary_hits_bit_exact(name, golden):
    """CPU walk of the product's flattened 8-wide BVH with the product's own traversal code (host build):
    parity mode must reproduce BVHAccel::intersect ids and t bit-exactly (exact ties excluded)."""
    g = golden(name)
    w = Walk(g, g, CONFIGS[name]["nl"], camera=g["camera"])
    assert sorted(w.slot_prim()) == list(range(len(g["prim_type"])))
    ids, ts = w.primary_hits(1)
    ok = ~g["hit_tie"].astype(bool)
    assert np.array_equal(ids[ok], g["hit_id"][ok])
    assert np.array_equal(ts[ok], g["hit_t"][ok])
    ids0, _ = w.primary_hits(0)                       # production float path: report, allow silhouette flips
    assert (ids0 != g["hit_id"]).mean() < 2e-3


@pytest.mark.parametrize("name", ["CBspheres_lambertian", "CBspheres", "CBgems", "CBcoil", "CBbunny", "bunny"])
def test_float_pipeline_cpu_walk_matches_oracle_paths(name, golden):
    """Same Philox streams on both sides: the float wavefront code walks the same paths as the fp64 oracle, so
    segment counts agree (almost) exactly and images agree far below Monte-Carlo noise."""
    g = golden(name); cfg = CONFIGS[name]
    cam = g["small_camera"]; W, H = SMALL_RES
    ref, cnt = O.Scene(g).with_camera(cam).render(W, H, 2, cfg["nl"], cfg["depth"], rng="philox", seed=11)
    rgb, c2 = Walk(g, g, cfg["nl"], camera=cam).render(2, cfg["depth"], seed=11)
    assert abs(int(c2[1]) - int(cnt[0])) <= 2e-4 * cnt[0] + 2
    assert abs(int(c2[2]) - int(cnt[1])) <= 2e-4 * cnt[1] + 8
    ok, info = images_match(rgb, ref)
    assert ok, info


# ---- host C++ side: COLLADA import ----------------------------------------------------------------------------------
def test_host_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "dsrt_host.h")).read()
    declared = sorted(set(re.findall(r"\b(dsrth_[a-z0-9_]+)\s*\(", hdr)))
    assert declared == sorted(D.HOST_EXPORTED_SYMBOLS)
    H = D.load_host_library()
    for s in declared:
        assert hasattr(H, s), s


needs_scenes = pytest.mark.skipif(not os.path.exists(O.ref_scene_path("CBspheres_lambertian.dae")),
                                  reason="reference .dae scenes are staged by oracle/build_ref.sh (not committed)")


@needs_scenes
@pytest.mark.parametrize("name", list(CONFIGS))
def test_collada_loader_is_bit_identical_to_reference(name, golden):
    """.dae -> flat scene + camera through the product's C++ loader == dump of the compiled reference
    (primitive order, triangle vertex rotation, world positions, half-edge vertex normals, BSDFs, lights, camera)."""
    g = golden(name); cfg = CONFIGS[name]
    W, H = ID_RES
    sc, cam = D.load_dae(O.ref_scene_path(cfg["file"]), W, H, O.ref_scene_path(cfg["cam"]) if cfg["cam"] else None)
    for k in sc:
        assert sc[k].shape == g[k].shape, k
        assert np.array_equal(sc[k], g[k]), f"{name}: {k}"
    assert np.array_equal(cam, g["camera"])


def test_collada_loader_errors_and_procedural_scene(tmp_path):
    with pytest.raises(D.DsrtError, match="cannot open"):
        D.load_dae(str(tmp_path / "missing.dae"), 8, 8)
    bad = tmp_path / "bad.dae"; bad.write_text("<COLLADA><asset><up_axis>Y_UP</up_axis></asset>")
    with pytest.raises(D.DsrtError, match="XML error"):
        D.load_dae(str(bad), 8, 8)
    # non-manifold input is an error message, not exit(1) (halfEdgeMesh.cpp:165-175)
    from dsgpuraytracing_b200 import scenes as S
    V = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0], [0, 0, 1]], float)
    F = np.array([[0, 1, 2], [0, 1, 3]])          # the directed edge 0->1 appears twice
    S.write_dae(str(tmp_path / "nm.dae"), [("m", V, F, 0)], [(0, (0.5,) * 3, (0,) * 3, 0)])
    with pytest.raises(D.DsrtError, match="halfedge"):
        D.load_dae(str(tmp_path / "nm.dae"), 8, 8)
    # ... unless the direct indexed-triangle import is switched on: same triangles, summed-cross-product normals
    D.set_loader_option("direct_triangles", 1)
    try:
        nm, _ = D.load_dae(str(tmp_path / "nm.dae"), 8, 8)
        assert len(nm["prim_type"]) == 2 and np.allclose(nm["tri_pos"].reshape(2, 3, 3), V[F])
        n0 = np.cross(V[1] - V[0], V[2] - V[0]) + np.cross(V[1] - V[0], V[3] - V[0])
        assert np.allclose(nm["tri_nrm"].reshape(2, 3, 3)[0, 0], n0 / np.linalg.norm(n0))
        # a manifold mesh still takes the reference's half-edge route with the option on
        Vk, Fk = S.torus_knot(n_around=6, n_along=40); Vk = Vk.astype(np.float32).astype(np.float64)
        S.write_cb_mesh_dae(str(tmp_path / "k2.dae"), Vk, Fk)
        on, _ = D.load_dae(str(tmp_path / "k2.dae"), 8, 8)
    finally:
        D.set_loader_option("direct_triangles", 0)
    off, _ = D.load_dae(str(tmp_path / "k2.dae"), 8, 8)
    assert all(np.array_equal(on[k], off[k]) for k in on)
    with pytest.raises(D.DsrtError):
        D.set_loader_option("no_such_option", 1)
    # a procedural closed mesh written as .dae flows through the loader: same geometry as the flat-array route
    V, F = S.torus_knot(n_around=6, n_along=40)
    V = V.astype(np.float32).astype(np.float64)
    S.write_cb_mesh_dae(str(tmp_path / "knot.dae"), V, F)
    sc, cam = D.load_dae(str(tmp_path / "knot.dae"), 64, 36)
    flat = S.cb_mesh_scene(V, F)
    assert len(sc["prim_type"]) == len(flat["prim_type"]) == 2 * 6 * 40 + 12
    a = np.sort(np.sort(sc["tri_pos"].reshape(-1, 3, 3).round(9), axis=1).reshape(-1, 9), axis=0)
    b = np.sort(np.sort(flat["tri_pos"].reshape(-1, 3, 3).round(9), axis=1).reshape(-1, 9), axis=0)
    assert np.allclose(a, b, atol=1e-7)
    assert np.array_equal(sc["light_type"], [3]) and np.allclose(sc["light_param"][0, :16], flat["light_param"][0, :16])


def test_sample_split_across_ranks_gloo(tmp_path):
    """The N>1 path of bench.py on CPU: two gloo ranks each accumulate their k = r (mod 2) samples and one
    reduce to rank 0 reproduces the single-rank frame (the CPU walk of the product's float pipeline stands in for
    the GPU render here; the collective plumbing is what is under test)."""
    import subprocess, sys, textwrap
    script = tmp_path / "w.py"
    script.write_text(textwrap.dedent(f"""
        import os, sys, numpy as np, torch, torch.distributed as dist
        sys.path.insert(0, {ROOT!r})
        from tests.cpuwalk import Walk
        from tests.scenes import CONFIGS, SMALL_RES
        dist.init_process_group("gloo")
        r, n = dist.get_rank(), dist.get_world_size()
        z = np.load(os.path.join({ROOT!r}, "tests", "golden", "CBgems.npz")); g = {{k: z[k] for k in z.files}}
        w = Walk(g, g, 4, camera=g["small_camera"])
        spp = 4
        part, _ = w.render(spp, 8, seed=2, spp_begin=r, spp_count=spp // n, spp_stride=n)
        t = torch.from_numpy(part.astype(np.float64))
        dist.reduce(t, dst=0, op=dist.ReduceOp.SUM)
        if r == 0:
            full, _ = w.render(spp, 8, seed=2)
            assert np.allclose(t.numpy(), full, rtol=1e-5, atol=1e-7), np.abs(t.numpy() - full).max()
            print("SPLIT_OK")
        dist.destroy_process_group()
    """))
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                          "--master-port", "29531", str(script)], capture_output=True, text=True, timeout=300)
    assert "SPLIT_OK" in out.stdout, out.stdout + out.stderr


@pytest.mark.parametrize("name", ["CBspheres", "CBgems", "CBcoil", "CBbunny", "bunny"])
def test_wide_bvh_never_culls_a_hit(name, golden):
    """Random incoherent rays: the production traversal of the quantised 8-wide BVH must find exactly what a brute
    force loop over all primitive records (same float primitive tests, no BVH) finds -- i.e. the quantised slabs are
    conservative (this caught a float-rounding cull on zero-thickness wall boxes)."""
    g = golden(name)
    rng = np.random.default_rng(7)
    n = 3000
    lo = g["node_bbox"][0, :3]; hi = g["node_bbox"][0, 3:]
    o = (lo + (hi - lo) * rng.random((n, 3))).astype(np.float32)
    d = rng.normal(size=(n, 3)); d = (d / np.linalg.norm(d, axis=1, keepdims=True)).astype(np.float32)
    # axis-parallel and wall-grazing directions too
    d[:300] = np.eye(3, dtype=np.float32)[rng.integers(0, 3, 300)] * rng.choice([-1, 1], (300, 1)).astype(np.float32)
    w = Walk(g, g, 4)
    ids, ts, _ = w.trace(o, d)
    bi, bt = w.trace_brute(o, d)
    assert np.array_equal(ids, bi), f"{(ids != bi).sum()} rays culled"
    assert np.array_equal(ts, bt)


def split_env(g):
    """golden env scene -> (scene arrays without the type-4 light, env map)"""
    keep = g["light_type"] != 4
    arr = dict(g); arr["light_type"] = g["light_type"][keep]; arr["light_param"] = g["light_param"][keep]
    return arr, g["env_rgb"]


@pytest.mark.parametrize("name", ["env_CBspheres", "env_bunny", "env_CBgems"])
def test_environment_light_float_pipeline_matches_oracle(name, golden):
    g = golden(name); depth = int(g["depth"]); W, H = SMALL_RES
    ref, cnt = O.Scene(g).render(W, H, 4, 4, depth, rng="philox", seed=13)
    arr, env = split_env(g)
    rgb, c2 = Walk(arr, D.build_bvh2(arr), 4, camera=g["camera"], envmap=env).render(4, depth, seed=13)
    assert abs(int(c2[1]) - int(cnt[0])) <= 2 and abs(int(c2[2]) - int(cnt[1])) <= 8
    ok, info = images_match(rgb, ref)
    assert ok, info


# ---- host C++ side: environment-map readers (-e option, reference src/main.cpp:30-67) --------------------------------
EXR_DIR = os.path.join(ROOT, "tests", "golden", "exr")


@pytest.mark.parametrize("name", ["zip_half", "zip_float", "none_float", "zips_half", "zip_half_decreasing", "rle_half", "piz_half", "piz_float", "piz_half_wide"])
def test_exr_reader_matches_reference_tinyexr(name):
    """<name>.npy = what the reference's load_exr (vendored tinyexr) reads from <name>.exr (tests/golden/exr/make_exr_golden.py;
    zip_half / zip_float were also WRITTEN by tinyexr).  rle_half (not supported by that tinyexr) is pinned by the writer's input."""
    got = D.load_envmap(os.path.join(EXR_DIR, name + ".exr"))
    want = np.load(os.path.join(EXR_DIR, name + ".npy"))
    assert got.shape == want.shape and got.dtype == np.float32
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))      # bit exact, halves included


def test_exr_reader_extra_channels_offsets_and_errors(tmp_path):
    from tests.exr_writer import write_exr
    rng = np.random.default_rng(7)
    a = rng.random((19, 23, 4)).astype(np.float32)
    # RGBA with a HALF alpha and a UINT id channel, data window not at the origin: channels are found by NAME
    p = str(tmp_path / "rgba.exr")
    write_exr(p, {"R": a[..., 0], "G": a[..., 1], "B": a[..., 2], "A": a[..., 3].astype(np.float16),
                  "id": (a[..., 3] * 100).astype(np.uint32)}, "zip", data_window_origin=(-5, 11))
    assert np.array_equal(D.load_envmap(p), a[..., :3])
    # 1x1 image, ragged last ZIP block (19 rows = 16 + 3)
    q = str(tmp_path / "one.exr")
    write_exr(q, {"R": a[:1, :1, 0], "G": a[:1, :1, 1], "B": a[:1, :1, 2]}, "zips")
    assert np.array_equal(D.load_envmap(q), a[:1, :1, :3])
    # .pfm through the same entry point (bottom row first in the file)
    f = str(tmp_path / "m.pfm")
    with open(f, "wb") as fh:
        fh.write(b"PF\n23 19\n-1.0\n" + a[::-1, :, :3].tobytes())
    assert np.array_equal(D.load_envmap(f), a[..., :3])
    # errors are returned, never exit(): missing file, not an image, truncated, zlib data labelled PIZ, no RGB
    raw = open(p, "rb").read()
    bad = {"trunc.exr": raw[: len(raw) // 2], "junk.exr": b"hello world, not an image", "piz.exr": raw.replace(b"compression\0compression\0\x01\0\0\0\x03", b"compression\0compression\0\x01\0\0\0\x04")}
    for nm, data in bad.items():
        fp = str(tmp_path / nm); open(fp, "wb").write(data)
        with pytest.raises(D.DsrtError):
            D.load_envmap(fp)
    with pytest.raises(D.DsrtError):
        D.load_envmap(str(tmp_path / "missing.exr"))
    # PIZ round trip incl. a block the encoder had to store raw (noise does not compress) and odd sizes
    n = rng.random((33, 17, 3)).astype(np.float16)
    z = str(tmp_path / "noise_piz.exr")
    write_exr(z, {"R": n[..., 0], "G": n[..., 1], "B": n[..., 2]}, "piz")
    assert np.array_equal(D.load_envmap(z), n.astype(np.float32))
    g = str(tmp_path / "grey.exr")
    write_exr(g, {"Y": a[..., 0]}, "none")
    with pytest.raises(D.DsrtError):
        D.load_envmap(g)


def test_parallel_host_builders_do_not_depend_on_the_thread_count():
    """SAH build (task parallel), DP collapse (subtree tasks) and wide-node emission (1024 independent subtrees after a
    breadth-first top) must give byte-identical results whatever the number of worker threads."""
    from dsgpuraytracing_b200 import scenes as S
    sc, _ = S.triangle_soup(300000, seed=11)
    res = []
    old = os.environ.get("DSRT_HOST_THREADS")
    try:
        for threads in ("1", "5"):
            os.environ["DSRT_HOST_THREADS"] = threads
            bvh = D.build_bvh2(sc)
            w = Walk(sc, bvh, 1)
            res.append((bvh, w.nodes(), w.slot_prim(), w.info()))
    finally:
        if old is None:
            os.environ.pop("DSRT_HOST_THREADS", None)
        else:
            os.environ["DSRT_HOST_THREADS"] = old
    (b1, n1, s1, i1), (b5, n5, s5, i5) = res
    for k in b1:
        assert np.array_equal(b1[k], b5[k]), k
    assert i1 == i5 and np.array_equal(n1, n5) and np.array_equal(s1, s5)
    assert i1[0] > 1024 * 8          # large enough for the parallel emission path
    assert np.array_equal(np.sort(s1), np.arange(300000))


def test_untrusted_file_fields_are_errors_not_crashes(tmp_path):
    """a 64-bit chunk offset near 2^64 must not wrap the EXR reader's bounds check; negative / absurd `count`
    attributes in a .dae must come back as an error string through the C boundary, never as an escaped C++ exception."""
    import struct
    from tests.exr_writer import write_exr
    from dsgpuraytracing_b200 import scenes as S
    a = np.random.default_rng(3).random((8, 6, 3)).astype(np.float32)
    p = str(tmp_path / "ok.exr")
    write_exr(p, {"R": a[..., 0], "G": a[..., 1], "B": a[..., 2]}, "none")
    raw = bytearray(open(p, "rb").read())
    assert np.array_equal(D.load_envmap(p), a)
    # the line-offset table sits right after the header's terminating NUL: find the first chunk offset by value and poison it
    first_chunk = None
    for off in range(len(raw) - 8):
        v = struct.unpack_from("<Q", raw, off)[0]
        if v == off + 8 * 8:           # 8 scan lines, NONE compression: table of 8 offsets, the first chunk follows it
            first_chunk = off; break
    assert first_chunk is not None
    for poison in (0xFFFFFFFFFFFFFFF0, 0xFFFFFFFFFFFFFFFF, len(raw) + 5, len(raw) - 3):
        bad = bytearray(raw); struct.pack_into("<Q", bad, first_chunk, poison)
        fp = str(tmp_path / "poison.exr"); open(fp, "wb").write(bad)
        with pytest.raises(D.DsrtError, match="corrupt|truncated"):
            D.load_envmap(fp)
    # .dae counts
    V, F = S.torus_knot(n_around=4, n_along=12); V = V.astype(np.float32).astype(np.float64)
    dae = str(tmp_path / "k.dae"); S.write_cb_mesh_dae(dae, V, F)
    text = open(dae).read()
    assert len(D.load_dae(dae, 8, 8)[0]["prim_type"]) == 2 * 4 * 12 + 12
    import re
    for pat, rep in ((r'<polylist count="\d+">', '<polylist count="-1">', ), (r'<polylist count="\d+">', '<polylist count="2000000000">'),
                     (r'(<float_array id="mesh-mesh-positions-array" count=)"\d+"', r'\1"-7"'), (r'(<float_array id="mesh-mesh-positions-array" count=)"\d+"', r'\1"1999999999"')):
        broken, n = re.subn(pat, rep, text, count=1)
        assert n == 1
        fp = str(tmp_path / "broken.dae"); open(fp, "w").write(broken)
        with pytest.raises(D.DsrtError, match="implausible|too short"):
            D.load_dae(fp, 8, 8)


def test_node_word_helpers_match_a_plain_restatement():
    """layout.h node words: the order in which next_child opens the hit internal children of a node (front to back for the
    ray's octant, child index = child_base + rank among the internal slots), the record index of a primitive bit, and the
    removal of a ray's source triangle -- product code (host build) against a few lines of Python."""
    import ctypes as C
    from tests.cpuwalk import lib
    L = lib()
    L.cw_drop_source.restype = C.c_uint32
    rng = np.random.default_rng(11)
    for trial in range(400):
        n_inner = int(rng.integers(0, 9)); slots = rng.permutation(8)
        inner_slots = sorted(int(x) for x in slots[:n_inner])
        inner = sum(8 << (4 * s) for s in inner_slots)
        valid = 0
        for s in slots[n_inner:]:
            c = int(rng.integers(0, 4))
            valid |= ((1 << c) - 1) << (4 * int(s))
        hit_slots = [s for s in range(8) if rng.random() < 0.6]
        hits = sum(0xF << (4 * s) for s in hit_slots)
        d = rng.normal(size=3).astype(np.float32); d[np.abs(d) < 1e-3] = 0.5
        octant = int(d[0] < 0) | (int(d[1] < 0) << 1) | (int(d[2] < 0) << 2)
        base = int(rng.integers(0, 1 << 20))
        out = (C.c_uint32 * 8)()
        for ordered in (1, 0):
            n = L.cw_open_order(d.ctypes.data_as(C.c_void_p), C.c_uint32(base), C.c_uint32(hits), C.c_uint32(inner), ordered, out)
            want = [s for s in inner_slots if s in hit_slots]
            want.sort(key=(lambda s: -(s ^ (7 - octant))) if ordered else (lambda s: -s))
            assert [out[i] for i in range(n)] == [base + inner_slots.index(s) for s in want], (trial, ordered)
        # primitive bits -> records, and the source drop
        bits = [b for b in range(32) if (valid >> b) & 1]
        for rank, b in enumerate(bits):
            assert L.cw_prim_slot(C.c_uint32(base), C.c_uint32(valid), C.c_uint32(1 << b)) == base + rank
            mask = hits & valid
            assert L.cw_drop_source(C.c_uint32(mask), C.c_uint32(base), C.c_uint32(valid), base + rank) == mask & ~(1 << b)
        for src in (-1, -5, base - 1, base + len(bits), base + 40):
            assert L.cw_drop_source(C.c_uint32(hits & valid), C.c_uint32(base), C.c_uint32(valid), src) == hits & valid


@pytest.mark.parametrize("name", ["CBspheres_lambertian", "CBbunny"])
def test_light_aligned_grid_same_image_fewer_primitive_tests(name, golden, monkeypatch):
    """wide_bvh.h EndPlane: a node holding the emissive quad under an axis-aligned area light shifts its quantisation grid so
    that the quad's light-facing plane lies 3/128 quantum past a grid line.  Shadow rays (which stop at 0.999 x distance,
    pathtracer.cpp:486-504) then miss the quad's box.  Same paths, same image, same segment counts -- fewer primitive tests;
    and every quantised box still encloses its exact box with the 1/64-quantum margin (checked on the node itself)."""
    g = golden(name); depth = CONFIGS[name]["depth"]; cam = g["small_camera"]
    w1 = Walk(g, g, 4, camera=cam)
    rgb1, c1 = w1.render(2, depth, seed=5)
    monkeypatch.setenv("CW_NO_LIGHT_GRID", "1")
    w0 = Walk(g, g, 4, camera=cam)
    rgb0, c0 = w0.render(2, depth, seed=5)
    assert np.array_equal(rgb0, rgb1) and list(c0[:3]) == list(c1[:3])
    assert int(c1[4]) < 0.9 * int(c0[4]), (c0, c1)          # primitive tests
    assert int(c1[3]) <= int(c0[3])                          # node visits
    # the node that holds the light quad: its low plane along the light's axis is within 1/16 quantum below the quad
    lt = np.asarray(g["light_type"]); lp = np.asarray(g["light_param"]).reshape(-1, 28)
    area = lp[lt == 3][0]; axis = int(np.argmax(np.abs(area[6:9]))); coord = area[3 + axis]
    assert area[6 + axis] < 0                                # faces -axis: rays arrive from the low side
    found = 0
    for raw in w1.nodes():
        b = raw.tobytes()
        org = np.frombuffer(b[:12], np.float32).astype(np.float64); scale = np.frombuffer(b[16:28], np.float32).astype(np.float64) / 2 ** 15
        valid = int(np.frombuffer(b[36:40], np.uint32)[0]); inner = int(np.frombuffer(b[44:48], np.uint32)[0])
        q = np.frombuffer(b[48:96], np.uint8).reshape(6, 8).astype(np.float64)
        for s in range(8):
            if not ((valid >> (4 * s)) & 0xF or (inner >> (4 * s + 3)) & 1):
                continue
            lo = org[axis] + q[axis, s] * scale[axis]; hi = org[axis] + q[3 + axis, s] * scale[axis]
            if hi - lo <= 1.0001 * scale[axis] and lo < coord < hi and (coord - lo) < scale[axis] / 16:
                assert (coord - lo) >= scale[axis] / 64
                found += 1
    assert found >= 1


@pytest.mark.parametrize("name", ["CBcoil", "CBbunny", "CBgems"])
def test_regrouped_top_nodes_same_image_fewer_node_visits(name, golden, monkeypatch):
    """wide_bvh.cpp step 2b: the scene-sized wall triangles of a Cornell box sit, in the reference's binary SAH tree, in a subtree
    whose box is the whole scene; regrouping the children of the top wide nodes by summed internal-node area makes the walls direct
    children of the root and gives the mesh its own node.  Same paths, same image, same segment counts, >= 10 % fewer node visits;
    every primitive still appears exactly once."""
    g = golden(name); depth = CONFIGS[name]["depth"]; cam = g["small_camera"]
    w0 = Walk(g, g, 4, camera=cam)
    rgb0, c0 = w0.render(2, depth, seed=5)
    monkeypatch.setenv("CW_REGROUP", "1")                     # the option is off by default (see include/dsrt.h)
    w1 = Walk(g, g, 4, camera=cam)
    rgb1, c1 = w1.render(2, depth, seed=5)
    assert sorted(w1.slot_prim().tolist()) == list(range(w1.n_prims))
    assert np.array_equal(rgb0, rgb1) and list(c0[:3]) == list(c1[:3])
    assert int(c1[3]) < 0.9 * int(c0[3]), (c0, c1)
    assert w1.info()[1] <= w0.info()[1] + 1                  # at most one more level of wide nodes


@pytest.mark.parametrize("name", ["CBspheres_lambertian", "CBbunny", "CBgems"])
def test_coplanar_slot_mates_are_dropped_with_the_source(name, golden, monkeypatch):
    """layout.h WideNode::flat: a ray leaving one half of a wall quad does not fetch the other half (it meets that plane at t = 0
    only).  Same image, same segment counts, fewer primitive tests; the marked slots are exactly pairs / triples of triangles whose
    vertices lie in one plane, and the mesh's curved leaves are not marked."""
    g = golden(name); depth = CONFIGS[name]["depth"]; cam = g["small_camera"]
    w1 = Walk(g, g, 4, camera=cam)
    rgb1, c1 = w1.render(2, depth, seed=5)
    monkeypatch.setenv("CW_NO_FLAT_SLOTS", "1")
    w0 = Walk(g, g, 4, camera=cam)
    rgb0, c0 = w0.render(2, depth, seed=5)
    assert list(c0[:4]) == list(c1[:4])
    assert np.allclose(rgb0, rgb1, rtol=0, atol=1e-6) and (np.abs(rgb0 - rgb1).max(axis=2) > 0).mean() < 1e-3
    assert int(c1[4]) < 0.9 * int(c0[4]), (c0, c1)
    sp = w1.slot_prim(); tri = np.asarray(g["tri_pos"]).reshape(-1, 3, 3); marked = 0; unmarked_multi = 0
    for raw, raw0 in zip(w1.nodes(), w0.nodes()):
        b = raw.tobytes()
        flat = int(np.frombuffer(b[28:32], np.uint32)[0]); base, valid = (int(x) for x in np.frombuffer(b[32:40], np.uint32))
        assert int(np.frombuffer(raw0.tobytes()[28:32], np.uint32)[0]) == 0
        rank = 0
        for s in range(8):
            c = bin((valid >> (4 * s)) & 0xF).count("1"); prims = sp[base + rank: base + rank + c]; rank += c
            f = (flat >> (4 * s)) & 0xF
            assert f in (0, 0xF)
            if c >= 2 and all(g["prim_type"][p] == 1 for p in prims):
                P = tri[prims].reshape(-1, 3); nrm = max((np.cross(t[1] - t[0], t[2] - t[0]) for t in tri[prims]), key=np.linalg.norm)
                dev = np.abs((P - P[0]) @ (nrm / np.linalg.norm(nrm))).max(); ext = np.linalg.norm(P - P[0], axis=1).max()
                if f:
                    assert dev <= 2e-6 * ext; marked += 1
                else:
                    assert dev > 0.5e-6 * ext; unmarked_multi += 1
            else:
                assert f == 0
    assert marked >= 3                                       # wall quads (and the light quad) that share a leaf slot


def test_light_aligned_grid_ignores_lights_it_cannot_use(golden, monkeypatch):
    """wide_bvh.cpp light_end_planes: only area lights whose direction is a coordinate axis and whose edges span the other two
    define a plane the grid can be aligned to; a tilted light, a non-finite one or a point light leave every node exactly as
    the option-off build has it (and nothing crashes)."""
    g = dict(golden("CBspheres_lambertian"))
    lp = np.array(g["light_param"], np.float64).reshape(-1, 28).copy()
    variants = []
    t = lp.copy(); t[0, 6:9] = [0.0, -0.8, 0.6]; variants.append(("tilted", g["light_type"], t))
    t = lp.copy(); t[0, 3:6] = [0.0, np.nan, 0.0]; variants.append(("nan", g["light_type"], t))
    t = lp.copy(); variants.append(("point", np.full_like(g["light_type"], 2), t))
    t = lp.copy(); t[0, 9:12] = [0.6, 0.1, 0.0]; variants.append(("edge leaves the plane", g["light_type"], t))
    for name, lt, p in variants:
        arr = dict(g); arr["light_type"] = lt; arr["light_param"] = p
        monkeypatch.delenv("CW_NO_LIGHT_GRID", raising=False)
        on = Walk(arr, g, 4).nodes()
        monkeypatch.setenv("CW_NO_LIGHT_GRID", "1")
        off = Walk(arr, g, 4).nodes()
        assert np.array_equal(on, off), name
    monkeypatch.delenv("CW_NO_LIGHT_GRID", raising=False)
    assert not np.array_equal(Walk(g, g, 4).nodes(), off)       # the scene's own light does move the grid
