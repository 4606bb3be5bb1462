"""ctypes front-end of tests/cpu_walk/libcpuwalk.so (TEST INFRASTRUCTURE, see cpu_walk.cu)."""
import ctypes as C
import os
import subprocess

import numpy as np

from dsgpuraytracing_b200._lib import _scene_struct, _Bvh2, _c

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "cpu_walk", "libcpuwalk.so")
_lib = None


def lib():
    global _lib
    if _lib is None:
        subprocess.run(["make", "-C", os.path.join(HERE, "cpu_walk")], check=True, stdout=subprocess.DEVNULL)
        _lib = C.CDLL(SO)
        _lib.cw_create.restype = C.c_void_p
    return _lib


class Walk:
    def __init__(self, arr, bvh, ns_area_light=4, camera=None, envmap=None):
        L = lib()
        keep = []
        s = _scene_struct(arr, keep)
        kb = [_c(bvh["node_bbox"], np.float64), _c(bvh["node_start"], np.int32), _c(bvh["node_range"], np.int32),
              _c(bvh["node_left"], np.int32), _c(bvh["node_right"], np.int32), _c(bvh["prim_order"], np.int32)]
        b = _Bvh2(); b.n_nodes = len(kb[1])
        (b.node_bbox, b.node_start, b.node_range, b.node_left, b.node_right, b.prim_order) = [k.ctypes.data for k in kb]
        env = _c(envmap, np.float32) if envmap is not None else None
        self.h = C.c_void_p(L.cw_create(C.byref(s), C.byref(b), int(ns_area_light), env.shape[1] if env is not None else 0,
                                        env.shape[0] if env is not None else 0, C.c_void_p(env.ctypes.data) if env is not None else None))
        if not self.h:
            raise RuntimeError("cw_create failed")
        self.n_prims = s.n_prims
        if camera is not None:
            self.set_camera(camera)

    def __del__(self):
        if getattr(self, "h", None):
            lib().cw_destroy(self.h); self.h = None

    def info(self):
        n = C.c_int64(); d = C.c_int32()
        lib().cw_info(self.h, C.byref(n), C.byref(d))
        return n.value, d.value

    def nodes(self):
        n, _ = self.info()
        out = np.zeros((n, int(lib().cw_node_bytes())), np.uint8)
        lib().cw_nodes(self.h, C.c_void_p(out.ctypes.data))
        return out

    def slot_prim(self):
        out = np.zeros(self.n_prims, np.int32)
        lib().cw_slot_prim(self.h, C.c_void_p(out.ctypes.data))
        return out

    def set_camera(self, cam):
        cam = _c(cam, np.float64); pos = cam[0:3].copy(); c2w = cam[3:12].copy()
        self.W, self.H = int(cam[12]), int(cam[13])
        lib().cw_set_camera(self.h, C.c_void_p(pos.ctypes.data), C.c_void_p(c2w.ctypes.data), self.W, self.H,
                            C.c_double(float(cam[14])))

    def trace(self, o, d, tmax=None, any_hit=False):
        o = _c(o, np.float32).reshape(-1, 3); d = _c(d, np.float32).reshape(-1, 3); n = len(o)
        tm = _c(tmax, np.float32) if tmax is not None else None
        ids = np.zeros(n, np.int32); ts = np.zeros(n, np.float32); cnt = np.zeros(2, np.uint64)
        lib().cw_trace(self.h, int(any_hit), C.c_int64(n), C.c_void_p(o.ctypes.data), C.c_void_p(d.ctypes.data),
                       C.c_void_p(tm.ctypes.data) if tm is not None else None, C.c_void_p(ids.ctypes.data),
                       C.c_void_p(ts.ctypes.data), C.c_void_p(cnt.ctypes.data))
        return ids, ts, cnt

    def trace_brute(self, o, d):
        o = _c(o, np.float32).reshape(-1, 3); d = _c(d, np.float32).reshape(-1, 3); n = len(o)
        ids = np.zeros(n, np.int32); ts = np.zeros(n, np.float32)
        lib().cw_trace_brute(self.h, C.c_int64(n), C.c_void_p(o.ctypes.data), C.c_void_p(d.ctypes.data),
                             C.c_void_p(ids.ctypes.data), C.c_void_p(ts.ctypes.data))
        return ids, ts

    def primary_hits(self, mode=1):
        ids = np.zeros((self.H, self.W), np.int32); ts = np.zeros((self.H, self.W))
        lib().cw_primary_hits(self.h, int(mode), C.c_void_p(ids.ctypes.data), C.c_void_p(ts.ctypes.data))
        return ids, ts

    def render(self, spp, max_depth, seed=0, spp_begin=0, spp_count=None, spp_stride=1):
        if spp_count is None:
            spp_count = spp
        rgb = np.zeros((self.H, self.W, 3), np.float32); cnt = np.zeros(5, np.uint64)
        lib().cw_render(self.h, int(spp_begin), int(spp_count), int(spp_stride), int(spp), int(max_depth), C.c_uint32(seed),
                        C.c_void_p(rgb.ctypes.data), C.c_void_p(cnt.ctypes.data))
        return rgb, cnt
