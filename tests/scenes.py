"""Scene/camera configurations shared by the golden-vector generator, the parity tests and bench.py.

Scene files are the reference's own .dae assets; they are staged (not committed) under oracle/_ref/scenes by
oracle/build_ref.sh.  Everything the tests need at run time on the GPU box is committed under tests/golden/.
"""
# name -> reference scene file, camera file (None = the default orbit camera of Application::load),
#         light samples (-l) and max depth (-m) used for the golden renders
CONFIGS = {
    # BASELINE.json configs[0]: CBspheres_lambertian 480x360, 16 spp, 4 light samples, depth 5
    "CBspheres_lambertian": dict(file="CBspheres_lambertian.dae", cam=None, nl=4, depth=5),
    "CBspheres": dict(file="CBspheres.dae", cam=None, nl=4, depth=8),        # mirror + glass SPHERES
    "CBgems": dict(file="CBgems.dae", cam=None, nl=4, depth=8),              # glass MESH (divergent secondary rays)
    "CBcoil": dict(file="CBcoil.dae", cam=None, nl=4, depth=8),              # mirror mesh, 7 884 tris
    "CBbunny": dict(file="CBbunny.dae", cam=None, nl=4, depth=8),            # Cornell box + 28 576-tri mesh
    "bunny": dict(file="bunny.dae", cam=None, nl=4, depth=8),                # hemisphere ("ambient") light
    "CBempty": dict(file="CBempty.dae", cam=None, nl=4, depth=4),
    "CBgems_cam": dict(file="CBgems.dae", cam="cam_dragon.info", nl=1, depth=8),   # -f camera override path
}
ID_RES = (240, 180)        # primary-hit id maps
SMALL_RES = (96, 72)       # bit-exact rand()-driven render, 2 spp, seed 1
RMSE_RES = (160, 120)      # 1024-spp reference renders for the image gate
RMSE_SPP = 1024

# ---- the MEASURED workloads (bench.py, profiles/r2_configs.md): stand-ins for the scene files that are missing from the
# reference checkout (dsgpuraytracing_b200/scenes.py) -> fixtures tests/golden/standin_<name>.npz (make_golden_standin.py)
STANDIN_CONFIGS = {
    "cbdragon_standin": dict(nl=4, depth=8),      # BASELINE.json configs[1]: the bench workload
    "cblucy_standin": dict(nl=4, depth=8),        # configs[2]: glass mesh
}
STANDIN_ID_RES = (480, 270)


def scene_sha(arrays):
    """sha256 over the flat scene arrays (dtype-normalised) -- pins that two routes produced the SAME scene, bit for bit."""
    import hashlib
    import numpy as np
    h = hashlib.sha256()
    for k, dt in (("prim_type", np.int32), ("prim_bsdf", np.int32), ("tri_pos", np.float64), ("tri_nrm", np.float64),
                  ("sphere", np.float64), ("bsdf_type", np.int32), ("bsdf_param", np.float32), ("light_type", np.int32),
                  ("light_param", np.float64)):
        h.update(np.ascontiguousarray(arrays[k], dtype=dt).tobytes())
    return h.hexdigest()
