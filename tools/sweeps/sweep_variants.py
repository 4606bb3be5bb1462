"""A/B of differently compiled builds of libdsrt.so (DSRT_LIB) x run-time knobs on the bench workload.
  python tools/sweeps/sweep_variants.py [spp]            # parent: one subprocess per library
Each line: library, option set, Mrays/s and per-stage seconds of a 64-spp render (second of two)."""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
LIBS = ["libdsrt.so"]
# edit for the experiment at hand: LIBS = builds to compare (make ../libdsrt_x.so OUT=../libdsrt_x.so EXTRA="-DDSRT_...=..."),
# OPTS = dsrt_set_option overrides per run (on top of `defaults` below)
OPTS = [{}, {"postpone_min_lanes": 8}, {"postpone_min_lanes": 16}, {"refill_busy_lanes": 16}, {"refill_busy_lanes": 20},
        {"coop_min_pairs": 2}, {"coop_min_pairs": 12}, {"pool_batches": 4}, {"pool_batches": 16}, {"batch_spp": 4}, {"batch_spp": 8}]
# or from the environment: SWEEP_LIBS="libdsrt.so,libdsrt_x.so" SWEEP_OPTS='[{}, {"refill_busy_lanes": 16}]'
if os.environ.get("SWEEP_LIBS"):
    LIBS = os.environ["SWEEP_LIBS"].split(",")
if os.environ.get("SWEEP_OPTS"):
    OPTS = json.loads(os.environ["SWEEP_OPTS"])

if len(sys.argv) > 1 and sys.argv[1] == "child":
    import numpy as np
    import dsgpuraytracing_b200 as D
    from dsgpuraytracing_b200 import scenes as S
    spp = int(sys.argv[2])
    scene = os.environ.get("SWEEP_SCENE", "c2")        # c2 (bench scene) | soupN (config 5, N Mi triangles, 3840x2160, 1 light sample)
    if scene.startswith("soup"):
        sc, cam = S.triangle_soup(int(scene[4:]) << 20); nl = 1
    elif scene == "c3":
        sc, cam = S.cblucy_standin(1920, 1080); nl = 4
    else:
        V, F = S.torus_knot(); V = V.astype(np.float32).astype(np.float64)
        sc = S.cb_mesh_scene(V, F); cam = S.cam_dragon(1920, 1080); nl = 4
    core = D.Core(0); core.set_params(spp, nl, 8, 0); core.load(sc, camera=cam, device_build=bool(int(os.environ.get("SWEEP_DEVICE_BUILD", "0")))); core.set_option("stage_timing", 1)
    tree_defaults = {"light_aligned_grid": 1, "drop_coplanar_mates": 1, "regroup_top": 0}      # take effect at dsrt_build_accel
    tree_now = dict(tree_defaults)
    defaults = {"max_ctas_per_sm": 0, "postpone_min_lanes": 8, "refill_busy_lanes": 18, "coop_min_pairs": 6, "postpone_wait_mode": 0, "pool_batches": 8, "batch_spp": 0, "smem_carveout_pct": -1, "refill_hi_lanes": 26, "refill_patience": 6}
    for o in OPTS:
        try:
            for k, v in {**defaults, **{k: v for k, v in o.items() if k not in tree_defaults}}.items():
                core.set_option(k, v)
            tree = {k: {**tree_defaults, **o}[k] for k in tree_defaults}
            if tree != tree_now:
                for k, v in tree.items():
                    core.set_option(k, v)
                core.build_accel(); tree_now = tree
        except D.DsrtError as e:
            print(os.path.basename(os.environ.get("DSRT_LIB", "libdsrt.so")), o, "unsupported:", str(e)[:60]); continue
        core.render(); rgb, st = core.render()
        print("%-18s %-60s Mrays/s %7.1f  gpu_s %.4f extend %.4f connect %.4f shade %.4f" % (
            os.path.basename(os.environ.get("DSRT_LIB", "libdsrt.so")), o, st.segments / st.gpu_seconds / 1e6, st.gpu_seconds,
            st.extend_seconds, st.connect_seconds, st.shade_seconds), flush=True)
else:
    spp = sys.argv[1] if len(sys.argv) > 1 else "64"
    for lib in LIBS:
        p = os.path.join(ROOT, "dsgpuraytracing_b200", lib)
        if not os.path.exists(p):
            continue
        subprocess.run([sys.executable, os.path.abspath(__file__), "child", spp], env={**os.environ, "DSRT_LIB": p})
