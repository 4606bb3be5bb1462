"""oracle/oracle.py -- TEST INFRASTRUCTURE ONLY.

ctypes front-end for the plain-C restatement of the reference render path (oracle/pt_oracle.c ->
oracle/liboracle.so) and a helper that runs the compiled reference itself (oracle/_ref/ref_driver).
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module; the product (dsgpuraytracing_b200) never does.
"""
import ctypes as C
import os
import subprocess
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")
REF_DRIVER = os.path.join(REF_DIR, "ref_driver")
REF_SCENES = os.path.join(REF_DIR, "scenes")
LIB_PATH = os.path.join(HERE, "liboracle.so")

SCENE_KEYS = ("prim_type", "prim_bsdf", "tri_pos", "tri_nrm", "sphere", "bsdf_type", "bsdf_param",
              "light_type", "light_param", "camera")
BVH_KEYS = ("node_bbox", "node_start", "node_range", "node_left", "node_right", "prim_order")


def build():
    """(Re)build liboracle.so (and oracle/_ref when /root/reference is present)."""
    subprocess.run(["make", "-C", HERE, "liboracle.so"], check=True, stdout=subprocess.DEVNULL)
    subprocess.run([os.path.join(HERE, "build_ref.sh")], check=True, stdout=subprocess.DEVNULL,
                   stderr=subprocess.DEVNULL)


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            subprocess.run(["make", "-C", HERE, "liboracle.so"], check=True, stdout=subprocess.DEVNULL)
        _lib = C.CDLL(LIB_PATH)
    return _lib


class _Scene(C.Structure):
    _fields_ = [("n_prims", C.c_int), ("prim_type", C.c_void_p), ("prim_bsdf", C.c_void_p),
                ("tri_pos", C.c_void_p), ("tri_nrm", C.c_void_p), ("sphere", C.c_void_p),
                ("n_bsdf", C.c_int), ("bsdf_type", C.c_void_p), ("bsdf_param", C.c_void_p),
                ("n_lights", C.c_int), ("light_type", C.c_void_p), ("light_param", C.c_void_p),
                ("cam", C.c_void_p), ("env_w", C.c_int), ("env_h", C.c_int), ("env_rgb", C.c_void_p),
                ("env_pThetaPhi", C.c_void_p), ("env_pTheta", C.c_void_p), ("env_pPhiGivenTheta", C.c_void_p)]


class _Bvh(C.Structure):
    _fields_ = [("n_nodes", C.c_int), ("node_bbox", C.c_void_p), ("node_start", C.c_void_p),
                ("node_range", C.c_void_p), ("node_left", C.c_void_p), ("node_right", C.c_void_p),
                ("prim_order", C.c_void_p)]


def _c(a, dt):
    return np.ascontiguousarray(a, dtype=dt)


class Scene:
    """Flat scene arrays (same layout as include/dsrt.h) + optional BVH."""

    def __init__(self, arrays):
        a = arrays
        self.prim_type = _c(a["prim_type"], np.int32)
        self.prim_bsdf = _c(a["prim_bsdf"], np.int32)
        self.tri_pos = _c(a["tri_pos"], np.float64).reshape(-1, 9)
        self.tri_nrm = _c(a["tri_nrm"], np.float64).reshape(-1, 9)
        self.sphere = _c(a["sphere"], np.float64).reshape(-1, 4)
        self.bsdf_type = _c(a["bsdf_type"], np.int32)
        self.bsdf_param = _c(a["bsdf_param"], np.float32).reshape(-1, 8)
        self.light_type = _c(a["light_type"], np.int32)
        self.light_param = _c(a["light_param"], np.float64).reshape(-1, 28)
        self.camera = _c(a["camera"], np.float64)
        self.env_rgb = None
        if "env_rgb" in a and a["env_rgb"] is not None and np.size(a["env_rgb"]):
            self.set_envmap(a["env_rgb"])
        self.bvh = None
        if "node_bbox" in a:
            self.set_bvh({k: a[k] for k in BVH_KEYS})

    @property
    def n_prims(self):
        return len(self.prim_type)

    def arrays(self):
        d = {k: getattr(self, k) for k in SCENE_KEYS}
        if self.env_rgb is not None:
            d["env_rgb"] = self.env_rgb
        if self.bvh is not None:
            d.update(self.bvh)
        return d

    def set_envmap(self, rgb):
        """EnvironmentLight tables (environment_light.cpp:6-53); the scene must list a light of type 4."""
        self.env_rgb = _c(rgb, np.float32)
        h, w = self.env_rgb.shape[:2]
        self.env_tp = np.zeros((h, w), np.float32); self.env_t = np.zeros(h, np.float32); self.env_pgt = np.zeros((h, w), np.float32)
        lib().orc_env_build(w, h, self.env_rgb.ctypes.data_as(C.c_void_p), self.env_tp.ctypes.data_as(C.c_void_p),
                            self.env_t.ctypes.data_as(C.c_void_p), self.env_pgt.ctypes.data_as(C.c_void_p))

    def set_bvh(self, b):
        self.bvh = {
            "node_bbox": _c(b["node_bbox"], np.float64).reshape(-1, 6),
            "node_start": _c(b["node_start"], np.int32), "node_range": _c(b["node_range"], np.int32),
            "node_left": _c(b["node_left"], np.int32), "node_right": _c(b["node_right"], np.int32),
            "prim_order": _c(b["prim_order"], np.int32),
        }

    def with_camera(self, camera):
        s = Scene(self.arrays())
        s.camera = _c(camera, np.float64)
        return s

    def _cs(self):
        s = _Scene()
        s.n_prims = self.n_prims
        for k in ("prim_type", "prim_bsdf", "tri_pos", "tri_nrm", "sphere", "bsdf_type", "bsdf_param",
                  "light_type", "light_param"):
            setattr(s, k, getattr(self, k).ctypes.data)
        s.n_bsdf = len(self.bsdf_type)
        s.n_lights = len(self.light_type)
        s.cam = self.camera.ctypes.data
        if self.env_rgb is not None:
            s.env_h, s.env_w = self.env_rgb.shape[:2]
            s.env_rgb = self.env_rgb.ctypes.data; s.env_pThetaPhi = self.env_tp.ctypes.data
            s.env_pTheta = self.env_t.ctypes.data; s.env_pPhiGivenTheta = self.env_pgt.ctypes.data
        return s

    def _cb(self):
        if self.bvh is None:
            self.build_bvh()
        b = _Bvh()
        b.n_nodes = len(self.bvh["node_start"])
        for k in BVH_KEYS:
            setattr(b, k, self.bvh[k].ctypes.data)
        return b

    # ---- restated algorithms -------------------------------------------------
    def build_bvh(self):
        n = max(self.n_prims, 1)
        bbox = np.zeros((2 * n, 6)); st = np.zeros(2 * n, np.int32); rg = np.zeros(2 * n, np.int32)
        le = np.zeros(2 * n, np.int32); ri = np.zeros(2 * n, np.int32); order = np.zeros(n, np.int32)
        s = self._cs()
        m = lib().orc_build_bvh(C.byref(s), *[x.ctypes.data_as(C.c_void_p) for x in (bbox, st, rg, le, ri, order)])
        self.set_bvh({"node_bbox": bbox[:m], "node_start": st[:m], "node_range": rg[:m], "node_left": le[:m],
                      "node_right": ri[:m], "prim_order": order[:self.n_prims]})
        return self.bvh

    def primary_hits(self, W, H, ties=True):
        ids = np.zeros((H, W), np.int32); ts = np.zeros((H, W)); tie = np.zeros((H, W), np.uint8)
        s, b = self._cs(), self._cb()
        lib().orc_primary_hits(C.byref(s), C.byref(b), W, H, ids.ctypes.data_as(C.c_void_p),
                               ts.ctypes.data_as(C.c_void_p), tie.ctypes.data_as(C.c_void_p) if ties else None)
        return ids, ts, tie

    def render(self, W, H, spp, ns_area_light, max_depth, rng="philox", seed=0, spp_begin=0, spp_count=None,
               arg_order_rtl=1):
        """Returns (rgb[H,W,3] float32 row0=bottom, counters[closest, any, box_tests, prim_tests])."""
        if spp_count is None:
            spp_count = spp
        rgb = np.zeros((H, W, 3), np.float32); cnt = np.zeros(4)
        s, b = self._cs(), self._cb()
        lib().orc_render(C.byref(s), C.byref(b), W, H, spp_begin, spp_count, spp, ns_area_light, max_depth,
                         0 if rng == "rand" else 1, C.c_uint32(seed), arg_order_rtl,
                         rgb.ctypes.data_as(C.c_void_p), cnt.ctypes.data_as(C.c_void_p))
        return rgb, cnt

    def closest_hit(self, o, d, max_t=np.inf):
        o = _c(o, np.float64); d = _c(d, np.float64); t = C.c_double(); n = np.zeros(3)
        s, b = self._cs(), self._cb()
        f = lib().orc_closest_hit
        f.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_double, C.c_void_p, C.c_void_p]
        pid = f(C.addressof(s), C.addressof(b), o.ctypes.data, d.ctypes.data, max_t, C.addressof(t), n.ctypes.data)
        return pid, t.value, n

    def any_hit(self, o, d, max_t=np.inf):
        o = _c(o, np.float64); d = _c(d, np.float64)
        s, b = self._cs(), self._cb()
        f = lib().orc_any_hit
        f.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_double]
        return bool(f(C.addressof(s), C.addressof(b), o.ctypes.data, d.ctypes.data, max_t))

    def closest_hit_brute(self, o, d):
        o = _c(o, np.float64); d = _c(d, np.float64); t = C.c_double()
        s = self._cs()
        f = lib().orc_closest_hit_brute
        f.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        pid = f(C.addressof(s), o.ctypes.data, d.ctypes.data, C.addressof(t))
        return pid, t.value


def generate_ray(camera, x, y):
    cam = _c(camera, np.float64); o = np.zeros(3); d = np.zeros(3)
    f = lib().orc_generate_ray
    f.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_void_p, C.c_void_p]
    f(cam.ctypes.data, x, y, o.ctypes.data, d.ctypes.data)
    return o, d


def make_coord_space(n):
    n = _c(n, np.float64); out = np.zeros(9)
    lib().orc_make_coord_space(n.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p))
    return out.reshape(3, 3).T  # columns x,y,z


def philox(seed, pixel, sample, depth, block):
    out = np.zeros(4, np.uint32)
    lib().orc_philox(C.c_uint32(seed), C.c_uint32(pixel), C.c_uint32(sample), C.c_uint32(depth), C.c_uint32(block),
                     out.ctypes.data_as(C.c_void_p))
    return out


def philox_raw(ctr, k0, k1):
    c = np.asarray(ctr, np.uint32).copy(); out = np.zeros(4, np.uint32)
    lib().orc_philox_raw(c.ctypes.data_as(C.c_void_p), C.c_uint32(k0), C.c_uint32(k1), out.ctypes.data_as(C.c_void_p))
    return out


def to_color(rgb):
    rgb = _c(rgb, np.float32); n = rgb.size // 3
    out = np.zeros(n, np.uint32)
    lib().orc_to_color(rgb.ctypes.data_as(C.c_void_p), n, out.ctypes.data_as(C.c_void_p))
    return out.reshape(rgb.shape[:-1])


# ---- npz / directory I/O -------------------------------------------------------
def load_dir(d):
    arrays = {}
    for k in SCENE_KEYS + BVH_KEYS:
        p = os.path.join(d, k + ".npy")
        if os.path.exists(p):
            arrays[k] = np.load(p)
    return Scene(arrays)


def load_npz(path):
    z = np.load(path)
    return Scene({k: z[k] for k in z.files})


# ---- the compiled reference -----------------------------------------------------
def have_reference():
    return os.path.exists(REF_DRIVER)


def ref_scene_path(name):
    return os.path.join(REF_SCENES, name)


def run_reference(scene_path, W, H, cam=None, spp=1, nl=4, depth=1, seed=1, dump_scene=False, ids=False,
                  render=False, threads=1, timeout=3600, envmap=None):
    """Runs oracle/_ref/ref_driver (the reference's own CPU code) and returns its .npy outputs."""
    if not have_reference():
        raise RuntimeError("oracle/_ref/ref_driver is not built (run oracle/build_ref.sh where /root/reference exists)")
    with tempfile.TemporaryDirectory() as td:
        cmd = [REF_DRIVER, "-w", str(W), "-h", str(H), "-s", str(spp), "-l", str(nl), "-m", str(depth), "-t",
               str(threads), "--seed", str(seed), "--out", td]
        if cam:
            cmd += ["-f", cam]
        if envmap:
            cmd += ["--envmap", str(envmap[0]), str(envmap[1])]
        if dump_scene:
            cmd.append("--dump-scene")
        if ids:
            cmd.append("--ids")
        if render:
            cmd.append("--render")
        cmd.append(scene_path)
        subprocess.run(cmd, check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, timeout=timeout)
        out = {}
        for f in os.listdir(td):
            if f.endswith(".npy"):
                out[f[:-4]] = np.load(os.path.join(td, f))
        return out
