"""ctypes binding of include/dsrt.h (libdsrt.so).  No CPU fallback: a missing library raises."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))

EXPORTED_SYMBOLS = [
    "dsrt_create", "dsrt_create_multi", "dsrt_device_count", "dsrt_destroy", "dsrt_last_error", "dsrt_version", "dsrt_set_scene", "dsrt_set_bvh",
    "dsrt_set_camera", "dsrt_set_window", "dsrt_set_params", "dsrt_set_envmap", "dsrt_set_option", "dsrt_build_bvh2", "dsrt_build_accel",
    "dsrt_accel_info", "dsrt_upload_accel", "dsrt_accel_bytes", "dsrt_render", "dsrt_render_tonemapped", "dsrt_cancel", "dsrt_render_device", "dsrt_resolve_device", "dsrt_sync",
    "dsrt_collect_stats", "dsrt_primary_hits", "dsrt_trace_closest", "dsrt_trace_any", "dsrt_tonemap", "dsrt_measure_read_bandwidth",
]


class DsrtError(RuntimeError):
    pass


def lib_path():
    # DSRT_LIB: A/B runs of differently compiled builds of the same library (tools/sweeps/sweep_variants.py); never a fallback
    return os.environ.get("DSRT_LIB") or os.path.join(_HERE, "libdsrt.so")


_lib = None


def load_library():
    global _lib
    if _lib is None:
        p = lib_path()
        if not os.path.exists(p):
            raise DsrtError(f"{p} is missing: build it with `make -C dsgpuraytracing_b200/csrc` "
                            "(or __graft_entry__.build()); there is no CPU fallback")
        _lib = C.CDLL(p)
        _lib.dsrt_last_error.restype = C.c_char_p
        _lib.dsrt_version.restype = C.c_char_p
    return _lib


class _Scene(C.Structure):
    _fields_ = [("n_prims", C.c_int32), ("prim_type", C.c_void_p), ("prim_bsdf", C.c_void_p),
                ("tri_pos", C.c_void_p), ("tri_nrm", C.c_void_p), ("sphere", C.c_void_p),
                ("n_bsdf", C.c_int32), ("bsdf_type", C.c_void_p), ("bsdf_param", C.c_void_p),
                ("n_lights", C.c_int32), ("light_type", C.c_void_p), ("light_param", C.c_void_p)]


class _Bvh2(C.Structure):
    _fields_ = [("n_nodes", C.c_int32), ("node_bbox", C.c_void_p), ("node_start", C.c_void_p),
                ("node_range", C.c_void_p), ("node_left", C.c_void_p), ("node_right", C.c_void_p),
                ("prim_order", C.c_void_p)]


class Stats(C.Structure):
    _fields_ = [("camera_samples", C.c_uint64), ("extend_rays", C.c_uint64), ("shadow_rays", C.c_uint64),
                ("extend_nodes", C.c_uint64), ("extend_prims", C.c_uint64),
                ("connect_nodes", C.c_uint64), ("connect_prims", C.c_uint64), ("gpu_seconds", C.c_double),
                ("extend_seconds", C.c_double), ("connect_seconds", C.c_double), ("shade_seconds", C.c_double),
                ("kernel_launches", C.c_uint32), ("batches", C.c_uint32), ("null_shadow_rays", C.c_uint64)]

    @property
    def segments(self):
        return self.extend_rays + self.shadow_rays

    @property
    def nodes_visited(self):
        return self.extend_nodes + self.connect_nodes

    @property
    def prims_tested(self):
        return self.extend_prims + self.connect_prims

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


def _c(a, dt):
    return np.ascontiguousarray(a, dtype=dt)


def _scene_struct(arr, keep):
    s = _Scene()
    pt = _c(arr["prim_type"], np.int32); pb = _c(arr["prim_bsdf"], np.int32)
    tp = _c(arr["tri_pos"], np.float64); tn = _c(arr["tri_nrm"], np.float64); sp = _c(arr["sphere"], np.float64)
    bt = _c(arr["bsdf_type"], np.int32); bp = _c(arr["bsdf_param"], np.float32)
    lt = _c(arr["light_type"], np.int32); lp = _c(arr["light_param"], np.float64)
    keep.extend([pt, pb, tp, tn, sp, bt, bp, lt, lp])
    s.n_prims = len(pt); s.prim_type = pt.ctypes.data; s.prim_bsdf = pb.ctypes.data
    s.tri_pos = tp.ctypes.data; s.tri_nrm = tn.ctypes.data; s.sphere = sp.ctypes.data
    s.n_bsdf = len(bt); s.bsdf_type = bt.ctypes.data; s.bsdf_param = bp.ctypes.data
    s.n_lights = len(lt); s.light_type = lt.ctypes.data; s.light_param = lp.ctypes.data
    return s


def build_bvh2(arr):
    """Host SAH builder (dsrt_build_bvh2).  Returns the dict of BVH arrays; needs no GPU."""
    L = load_library()
    keep = []
    s = _scene_struct(arr, keep)
    n = max(s.n_prims, 1)
    bbox = np.zeros((2 * n, 6)); st = np.zeros(2 * n, np.int32); rg = np.zeros(2 * n, np.int32)
    le = np.zeros(2 * n, np.int32); ri = np.zeros(2 * n, np.int32); order = np.zeros(n, np.int32)
    m = C.c_int32(0)
    rc = L.dsrt_build_bvh2(C.byref(s), *[C.c_void_p(x.ctypes.data) for x in (bbox, st, rg, le, ri, order)], C.byref(m))
    if rc:
        raise DsrtError(f"dsrt_build_bvh2 failed ({rc})")
    m = m.value
    return {"node_bbox": bbox[:m].copy(), "node_start": st[:m].copy(), "node_range": rg[:m].copy(),
            "node_left": le[:m].copy(), "node_right": ri[:m].copy(), "prim_order": order[:s.n_prims].copy()}


class Core:
    """One dsrt context = one GPU.  Mirrors the call order of CUDAPathTracer::init (cuda_src/setup.cu:181-201)."""

    def __init__(self, device=0, devices=None):
        self.L = load_library()
        self.ctx = C.c_void_p()
        if devices is not None:
            arr = (C.c_int * len(devices))(*devices)
            rc = self.L.dsrt_create_multi(len(devices), arr, C.byref(self.ctx))
        else:
            rc = self.L.dsrt_create(int(device), C.byref(self.ctx))
        if rc:
            msg = self.L.dsrt_last_error(self.ctx).decode() if self.ctx else "no context"
            if self.ctx:
                self.L.dsrt_destroy(self.ctx)
            self.ctx = None
            raise DsrtError(f"dsrt_create failed ({rc}): {msg}")
        self.width = self.height = 0
        self.ns_aa = 1

    def close(self):
        if getattr(self, "ctx", None):
            self.L.dsrt_destroy(self.ctx)
            self.ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc, what):
        if rc:
            raise DsrtError(f"{what} failed ({rc}): {self.L.dsrt_last_error(self.ctx).decode()}")

    def set_scene(self, arr):
        keep = []
        s = _scene_struct(arr, keep)
        self._ck(self.L.dsrt_set_scene(self.ctx, C.byref(s)), "dsrt_set_scene")
        self.n_prims = s.n_prims

    def set_bvh(self, b):
        keep = [_c(b["node_bbox"], np.float64), _c(b["node_start"], np.int32), _c(b["node_range"], np.int32),
                _c(b["node_left"], np.int32), _c(b["node_right"], np.int32), _c(b["prim_order"], np.int32)]
        s = _Bvh2()
        s.n_nodes = len(keep[1])
        (s.node_bbox, s.node_start, s.node_range, s.node_left, s.node_right, s.prim_order) = [k.ctypes.data for k in keep]
        self._ck(self.L.dsrt_set_bvh(self.ctx, C.byref(s)), "dsrt_set_bvh")

    def set_camera(self, cam):
        """cam: the 17-double camera vector (pos[3], c2w[9] column-major, W, H, screenDist, hFov, vFov)."""
        cam = _c(cam, np.float64)
        pos = cam[0:3].copy(); c2w = cam[3:12].copy()
        self.width, self.height = int(cam[12]), int(cam[13])
        self._ck(self.L.dsrt_set_camera(self.ctx, C.c_void_p(pos.ctypes.data), C.c_void_p(c2w.ctypes.data),
                                        self.width, self.height, C.c_double(float(cam[14]))), "dsrt_set_camera")

    def set_window(self, x0=0, y0=0, width=0, height=0):
        """Tile partitioning: render only the pixels [x0, x0+width) x [y0, y0+height); width=0 clears the window."""
        self._ck(self.L.dsrt_set_window(self.ctx, int(x0), int(y0), int(width), int(height)), "dsrt_set_window")

    def set_params(self, ns_aa, ns_area_light, max_depth, seed=0):
        self.ns_aa = int(ns_aa)
        self._ck(self.L.dsrt_set_params(self.ctx, int(ns_aa), int(ns_area_light), int(max_depth), C.c_uint32(seed)),
                 "dsrt_set_params")

    def set_envmap(self, rgb):
        """rgb: [H, W, 3] float lat-long map, or None to remove the environment light."""
        if rgb is None:
            self._ck(self.L.dsrt_set_envmap(self.ctx, 0, 0, None), "dsrt_set_envmap")
            return
        rgb = _c(rgb, np.float32)
        self._ck(self.L.dsrt_set_envmap(self.ctx, rgb.shape[1], rgb.shape[0], C.c_void_p(rgb.ctypes.data)), "dsrt_set_envmap")

    def set_option(self, name, value):
        self._ck(self.L.dsrt_set_option(self.ctx, name.encode(), C.c_int64(int(value))), "dsrt_set_option")

    def build_accel(self):
        self._ck(self.L.dsrt_build_accel(self.ctx), "dsrt_build_accel")

    def upload_accel(self):
        self._ck(self.L.dsrt_upload_accel(self.ctx), "dsrt_upload_accel")

    def accel_bytes(self):
        b = C.c_int64()
        self._ck(self.L.dsrt_accel_bytes(self.ctx, C.byref(b)), "dsrt_accel_bytes")
        return b.value

    def accel_info(self):
        a, b, c, d = C.c_int64(), C.c_int64(), C.c_int64(), C.c_int32()
        self._ck(self.L.dsrt_accel_info(self.ctx, C.byref(a), C.byref(b), C.byref(c), C.byref(d)), "dsrt_accel_info")
        return {"wide_nodes": a.value, "node_bytes": b.value, "prim_bytes": c.value, "max_depth": d.value}

    def load(self, arr, camera=None, bvh=None, device_build=False):
        """set_scene + (host SAH build | given BVH | option "device_build": LBVH + collapse on the GPU) + build_accel (+ camera)."""
        self.set_scene(arr)
        self.set_option("device_build", 1 if device_build else 0)
        if not device_build:
            self.set_bvh(bvh if bvh is not None else build_bvh2(arr))
        self.build_accel()
        if camera is not None:
            self.set_camera(camera)

    def render(self, spp_begin=0, spp_count=None, spp_stride=1, out=None, rgba8=False):
        """Host-buffer render (H2D of nothing, D2H of the frame): returns (rgb[H,W,3], Stats), or (rgb, rgba8[H,W] uint32,
        Stats) with rgba8=True.  A render stopped by cancel() returns normally with self.cancelled = True."""
        if spp_count is None:
            spp_count = self.ns_aa
        rgb = out if out is not None else np.zeros((self.height, self.width, 3), np.float32)
        st = Stats()
        if rgba8:
            img = np.zeros((self.height, self.width), np.uint32)
            rc = self.L.dsrt_render_tonemapped(self.ctx, int(spp_begin), int(spp_count), int(spp_stride),
                                               C.c_void_p(rgb.ctypes.data), C.c_void_p(img.ctypes.data), C.byref(st))
        else:
            rc = self.L.dsrt_render(self.ctx, int(spp_begin), int(spp_count), int(spp_stride),
                                    C.c_void_p(rgb.ctypes.data), C.byref(st))
        self.cancelled = rc == 4
        if rc != 4:
            self._ck(rc, "dsrt_render")
        return (rgb, img, st) if rgba8 else (rgb, st)

    def cancel(self):
        """dsrt_cancel: stops the dsrt_render running in another thread after its current chunk of samples."""
        self._ck(self.L.dsrt_cancel(self.ctx), "dsrt_cancel")

    def render_device(self, d_accum_ptr, spp_begin, spp_count, spp_stride=1, stream=0, collect=False):
        st = Stats() if collect else None
        self._ck(self.L.dsrt_render_device(self.ctx, int(spp_begin), int(spp_count), int(spp_stride),
                                           C.c_void_p(d_accum_ptr), C.c_void_p(stream),
                                           C.byref(st) if collect else None), "dsrt_render_device")
        return st

    def resolve_device(self, d_accum_ptr, d_rgb_ptr=0, d_rgba8_ptr=0, stream=0):
        self._ck(self.L.dsrt_resolve_device(self.ctx, C.c_void_p(d_accum_ptr), C.c_void_p(d_rgb_ptr),
                                            C.c_void_p(d_rgba8_ptr), C.c_void_p(stream)), "dsrt_resolve_device")

    def collect_stats(self):
        st = Stats()
        self._ck(self.L.dsrt_collect_stats(self.ctx, C.byref(st)), "dsrt_collect_stats")
        return st

    def sync(self):
        self._ck(self.L.dsrt_sync(self.ctx), "dsrt_sync")

    def primary_hits(self, mode=1):
        ids = np.zeros((self.height, self.width), np.int32); ts = np.zeros((self.height, self.width))
        self._ck(self.L.dsrt_primary_hits(self.ctx, int(mode), C.c_void_p(ids.ctypes.data), C.c_void_p(ts.ctypes.data)),
                 "dsrt_primary_hits")
        return ids, ts

    def trace_closest(self, o, d, tmax=None):
        o = _c(o, np.float32).reshape(-1, 3); d = _c(d, np.float32).reshape(-1, 3); n = len(o)
        tm = _c(tmax, np.float32) if tmax is not None else None
        ids = np.zeros(n, np.int32); ts = np.zeros(n, np.float32)
        self._ck(self.L.dsrt_trace_closest(self.ctx, C.c_int64(n), C.c_void_p(o.ctypes.data), C.c_void_p(d.ctypes.data),
                                           C.c_void_p(tm.ctypes.data) if tm is not None else None,
                                           C.c_void_p(ids.ctypes.data), C.c_void_p(ts.ctypes.data)), "dsrt_trace_closest")
        return ids, ts

    def trace_any(self, o, d, tmax=None):
        o = _c(o, np.float32).reshape(-1, 3); d = _c(d, np.float32).reshape(-1, 3); n = len(o)
        tm = _c(tmax, np.float32) if tmax is not None else None
        hit = np.zeros(n, np.int32)
        self._ck(self.L.dsrt_trace_any(self.ctx, C.c_int64(n), C.c_void_p(o.ctypes.data), C.c_void_p(d.ctypes.data),
                                       C.c_void_p(tm.ctypes.data) if tm is not None else None,
                                       C.c_void_p(hit.ctypes.data)), "dsrt_trace_any")
        return hit

    def measure_read_bandwidth(self, nbytes, repeats=20):
        """GB/s of a read-only sweep over `nbytes` (32 MiB -> L2 read bandwidth, >> L2 -> HBM read bandwidth)."""
        g = C.c_double(0)
        self._ck(self.L.dsrt_measure_read_bandwidth(self.ctx, C.c_int64(int(nbytes)), C.c_int32(int(repeats)), C.byref(g)),
                 "dsrt_measure_read_bandwidth")
        return g.value

    def tonemap(self, rgb):
        rgb = _c(rgb, np.float32); n = rgb.size // 3
        out = np.zeros(n, np.uint32)
        self._ck(self.L.dsrt_tonemap(self.ctx, C.c_void_p(rgb.ctypes.data), C.c_int64(n), C.c_void_p(out.ctypes.data)),
                 "dsrt_tonemap")
        return out.reshape(rgb.shape[:-1])


# ---- host side (libdsrt_host.so): COLLADA import + the PathTracer mirror -------------------------------------------
HOST_EXPORTED_SYMBOLS = ["dsrth_load_dae", "dsrth_free", "dsrth_get_scene", "dsrth_get_camera", "dsrth_render_file", "dsrth_load_envmap", "dsrth_set_loader_option"]
_hostlib = None


def load_host_library():
    global _hostlib
    if _hostlib is None:
        load_library()
        p = os.path.join(_HERE, "libdsrt_host.so")
        if not os.path.exists(p):
            raise DsrtError(f"{p} is missing: build it with `make -C dsgpuraytracing_b200/csrc all`")
        _hostlib = C.CDLL(p)
    return _hostlib


def load_dae(path, width, height, cam_info=None):
    """ColladaParser::load + Application::load (+ loadCamera): returns (scene arrays dict, camera[17])."""
    H = load_host_library()
    h = C.c_void_p(); err = C.create_string_buffer(512)
    rc = H.dsrth_load_dae(path.encode(), int(width), int(height), cam_info.encode() if cam_info else None, C.byref(h), err, 512)
    if rc:
        raise DsrtError(f"dsrth_load_dae({path}) failed: {err.value.decode()}")
    try:
        s = _Scene()
        H.dsrth_get_scene(h, C.byref(s))
        n, nb, nl = s.n_prims, s.n_bsdf, s.n_lights

        def arr(ptr, dt, shape):
            cnt = int(np.prod(shape))
            if cnt == 0:
                return np.zeros(shape, dt)
            buf = (C.c_char * (cnt * np.dtype(dt).itemsize)).from_address(ptr)
            return np.frombuffer(buf, dtype=dt).reshape(shape).copy()
        out = {"prim_type": arr(s.prim_type, np.int32, (n,)), "prim_bsdf": arr(s.prim_bsdf, np.int32, (n,)),
               "tri_pos": arr(s.tri_pos, np.float64, (n, 9)), "tri_nrm": arr(s.tri_nrm, np.float64, (n, 9)),
               "sphere": arr(s.sphere, np.float64, (n, 4)), "bsdf_type": arr(s.bsdf_type, np.int32, (nb,)),
               "bsdf_param": arr(s.bsdf_param, np.float32, (nb, 8)), "light_type": arr(s.light_type, np.int32, (nl,)),
               "light_param": arr(s.light_param, np.float64, (nl, 28))}
        cam = np.zeros(17)
        H.dsrth_get_camera(h, C.c_void_p(cam.ctypes.data))
        return out, cam
    finally:
        H.dsrth_free(h)


def set_loader_option(name, value):
    """Host loader options (dsrth_set_loader_option), e.g. ("direct_triangles", 1)."""
    if load_host_library().dsrth_set_loader_option(name.encode(), int(value)):
        raise DsrtError(f"unknown loader option {name}")


def load_envmap(path):
    """-e option: scan-line OpenEXR or .pfm lat-long map -> float32 [h, w, 3], top row first (dsrt_set_envmap layout)."""
    H = load_host_library()
    w, h = C.c_int32(0), C.c_int32(0); err = C.create_string_buffer(512)
    rc = H.dsrth_load_envmap(path.encode(), C.byref(w), C.byref(h), None, C.c_int64(0), err, 512)
    if rc:
        raise DsrtError(f"dsrth_load_envmap({path}) failed: {err.value.decode()}")
    out = np.zeros((h.value, w.value, 3), np.float32)
    rc = H.dsrth_load_envmap(path.encode(), C.byref(w), C.byref(h), C.c_void_p(out.ctypes.data), C.c_int64(out.size), err, 512)
    if rc:
        raise DsrtError(f"dsrth_load_envmap({path}) failed: {err.value.decode()}")
    return out


def render_file(path, width, height, spp, ns_area_light, max_depth, cam_info=None, n_gpus=1, seed=0, png=None):
    """main.cpp's headless GPU path through the C++ PathTracer class.  Returns (rgb[H,W,3], Stats, seconds dict)."""
    H = load_host_library()
    rgb = np.zeros((height, width, 3), np.float32); st = Stats(); err = C.create_string_buffer(512)
    tb, tr = C.c_double(), C.c_double()
    rc = H.dsrth_render_file(path.encode(), cam_info.encode() if cam_info else None, int(width), int(height), int(spp),
                             int(ns_area_light), int(max_depth), int(n_gpus), C.c_uint32(seed), C.c_void_p(rgb.ctypes.data),
                             png.encode() if png else None, C.byref(st), C.byref(tb), C.byref(tr), err, 512)
    if rc:
        raise DsrtError(f"dsrth_render_file failed ({rc}): {err.value.decode()}")
    return rgb, st, {"bvh_build": tb.value, "render": tr.value}
