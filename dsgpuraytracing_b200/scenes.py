"""Procedural workloads for the configurations of BASELINE.json whose scene files are not shipped with the
reference checkout (CBdragon.dae, CBlucy.dae, dragon.dae: reference .MISSING_LARGE_BLOBS), and the synthetic triangle
soups of config 5.  They produce the flat `dsrt_scene` arrays of include/dsrt.h (what the reference's loader hands to
CUDAPathTracer::loadPrimitives / loadLights) and can also be written as COLLADA (`write_dae`) so that they flow
through the same .dae import path as the reference's own scenes.

Cornell box constants are those of the reference's CB*.dae scenes (box [-1,1] x [0,1.5] x [-1,1] open towards +z,
0.8 x 0.6 emissive quad at y = 1.49 with radiance 10, area light at (0,1.49,0) facing -y with dim_x = 0.6 x,
dim_y = 0.8 z; wall albedos 0.6 grey / (0.6,0.2,0.2) / (0.2,0.2,0.6)); the camera is cam_dragon.info.
"""
import math

import numpy as np

# cam_dragon.info (reference repo root): pos, then c2w as stored row-major in the file
CAM_DRAGON_POS = (-1.918e-07, 1.30283, 2.54047)
CAM_DRAGON_C2W_ROWS = ((1, 2.65091e-08, -7.0691e-08), (-0.0, 0.936329, 0.351124), (7.5498e-08, -0.351124, 0.936329))
CB_HFOV = 49.13434          # <xfov> of the CB*.dae cameras
CB_ASPECT = 1.333333        # <aspect_ratio>


def configure_camera(hfov, vfov, W, H):
    """Camera::configure (reference src/camera.cpp:15-34): widen the narrower FOV to the frame's aspect ratio and
    return (hFov, vFov, screenDist)."""
    ar1 = math.tan(math.radians(hfov) / 2) / math.tan(math.radians(vfov) / 2)
    ar = W / H
    if ar1 < ar:
        hfov = 2 * math.degrees(math.atan(math.tan(math.radians(vfov) / 2) * ar))
    elif ar1 > ar:
        vfov = 2 * math.degrees(math.atan(math.tan(math.radians(hfov) / 2) / ar))
    return hfov, vfov, H / (2.0 * math.tan(math.radians(vfov) / 2))


def cam_dragon(W, H):
    """17-double camera vector (pos, c2w column-major, W, H, screenDist, hFov, vFov) for cam_dragon.info on a CB scene:
    parse_camera (collada.cpp:432-473) derives vFov from xfov/aspect_ratio in float, then Camera::configure."""
    aspect = float(np.float32(CB_ASPECT))
    vfov = float(np.float32(2 * math.degrees(math.atan(math.tan(math.radians(0.5 * float(np.float32(CB_HFOV)))) / aspect))))
    hf, vf, dist = configure_camera(float(np.float32(CB_HFOV)), vfov, W, H)
    rows = np.array(CAM_DRAGON_C2W_ROWS, float)
    return np.array(list(CAM_DRAGON_POS) + list(rows.T.reshape(-1)) + [W, H, dist, hf, vf], float)


def look_at_camera(pos, target, W, H, vfov=40.0):
    """Camera::compute_position conventions (camera.cpp:88-111): c2w[2] = unit(pos - target), up = +y."""
    pos = np.asarray(pos, float); target = np.asarray(target, float)
    z = pos - target; z /= np.linalg.norm(z)
    x = np.cross([0.0, 1.0, 0.0], z); x /= np.linalg.norm(x)
    y = np.cross(z, x); y /= np.linalg.norm(y)
    hf, vf, dist = configure_camera(vfov * W / H, vfov, W, H)
    return np.array(list(pos) + list(x) + list(y) + list(z) + [W, H, dist, hf, vf], float)


# the 12 box triangles of the reference's CB scenes (object order: ceiling, light quad, floor, left, right, back)
_CB_TRIS = (
    ((-1, 1.5, 1), (1, 1.5, -1), (1, 1.5, 1)), ((-1, 1.5, 1), (-1, 1.5, -1), (1, 1.5, -1)),
    ((-0.4, 1.49, 0.3), (0.4, 1.49, -0.3), (0.4, 1.49, 0.3)), ((-0.4, 1.49, 0.3), (-0.4, 1.49, -0.3), (0.4, 1.49, -0.3)),
    ((-1, 0, 1), (1, 0, -1), (-1, 0, -1)), ((-1, 0, 1), (1, 0, 1), (1, 0, -1)),
    ((-1, 0, 1), (-1, 1.5, -1), (-1, 1.5, 1)), ((-1, 0, 1), (-1, 0, -1), (-1, 1.5, -1)),
    ((1, 0, -1), (1, 1.5, 1), (1, 1.5, -1)), ((1, 0, -1), (1, 0, 1), (1, 1.5, 1)),
    ((-1, 0, -1), (1, 1.5, -1), (-1, 1.5, -1)), ((-1, 0, -1), (1, 0, -1), (1, 1.5, -1)),
)
_CB_NORMALS = ((0, 1, 0), (0, 1, 0), (0, -1, 0), (-1, 0, 0), (1, 0, 0), (0, 0, -1))


def cornell_box():
    """(tri_pos[12,9], tri_nrm[12,9], prim_bsdf[12], bsdf_type, bsdf_param, light_type, light_param)"""
    tris = [np.array(t, float).reshape(9) for t in _CB_TRIS]
    nrm = [np.tile(np.array(_CB_NORMALS[i // 2], float), 3) for i in range(12)]
    pb = [i // 2 for i in range(12)]
    bsdf_type = np.array([0, 4, 0, 0, 0, 0], np.int32)
    bsdf_param = np.zeros((6, 8), np.float32)
    bsdf_param[0, :3] = 0.6; bsdf_param[1, :3] = 10.0; bsdf_param[2, :3] = 0.6
    bsdf_param[3, :3] = (0.6, 0.2, 0.2); bsdf_param[4, :3] = (0.2, 0.2, 0.6); bsdf_param[5, :3] = 0.6
    light_param = np.zeros((1, 28))
    light_param[0, :16] = [10, 10, 10, 0, 1.49, 0, 0, -1, 0, 0.6, 0, 0, 0, 0, 0.8, float(np.float32(0.48))]
    return (np.array(tris), np.array(nrm), np.array(pb, np.int32), bsdf_type, bsdf_param,
            np.array([3], np.int32), light_param)


def torus_knot(n_around=22, n_along=2273, p=2, q=3, R=0.40, r=0.16, tube=0.075, bumps=0.35, centre=(0.0, 0.48, -0.1)):
    """Closed, consistently oriented genus-1 tube around a (p,q) torus knot with a bumpy radius: n_along*n_around
    vertices, 2*n_along*n_around triangles (22 x 2273 -> 100 012, the triangle count of CBdragon.dae per data.xlsx)."""
    t = np.linspace(0, 2 * np.pi, n_along, endpoint=False)
    def curve(t):
        c = R + r * np.cos(q * t)
        return np.stack([c * np.cos(p * t), r * np.sin(q * t) * 1.6, c * np.sin(p * t)], axis=1)
    C = curve(t)
    T = curve(t + 1e-4) - curve(t - 1e-4); T /= np.linalg.norm(T, axis=1, keepdims=True)
    # parallel-transported frame (no flips along a closed curve up to a final twist, absorbed smoothly)
    N = np.zeros_like(C); ref = np.array([0.0, 1.0, 0.0])
    n0 = ref - T[0] * (ref @ T[0]); N[0] = n0 / np.linalg.norm(n0)
    for i in range(1, n_along):
        v = N[i - 1] - T[i] * (N[i - 1] @ T[i]); N[i] = v / np.linalg.norm(v)
    # holonomy of the closed curve: transport the last frame one more step and measure its angle to the first
    v = N[-1] - T[0] * (N[-1] @ T[0]); v /= np.linalg.norm(v)
    twist = math.atan2(np.cross(N[0], v) @ T[0], N[0] @ v)
    ang = -twist * np.arange(n_along) / n_along
    B = np.cross(T, N)
    Nn = N * np.cos(ang)[:, None] + B * np.sin(ang)[:, None]
    Bn = np.cross(T, Nn)
    a = np.linspace(0, 2 * np.pi, n_around, endpoint=False)
    rad = tube * (1 + bumps * np.sin(37 * t)[:, None] * np.cos(3 * a)[None, :] + 0.5 * bumps * np.sin(11 * t)[:, None])
    V = C[:, None, :] + rad[:, :, None] * (np.cos(a)[None, :, None] * Nn[:, None, :] + np.sin(a)[None, :, None] * Bn[:, None, :])
    V = V.reshape(-1, 3) + np.asarray(centre)
    i = np.arange(n_along)[:, None]; j = np.arange(n_around)[None, :]
    v00 = (i * n_around + j); v01 = (i * n_around + (j + 1) % n_around)
    v10 = (((i + 1) % n_along) * n_around + j); v11 = (((i + 1) % n_along) * n_around + (j + 1) % n_around)
    F = np.concatenate([np.stack([v00, v11, v10], axis=-1).reshape(-1, 3), np.stack([v00, v01, v11], axis=-1).reshape(-1, 3)])   # outward orientation
    return V, F.astype(np.int64)


def vertex_normals(V, F):
    """Area-weighted vertex normals = normalised sum of cross(pj-pi, pk-pi) over incident faces
    (Vertex::computeNormal, reference halfEdgeMesh.h:492-515, interior-vertex branch)."""
    fn = np.cross(V[F[:, 1]] - V[F[:, 0]], V[F[:, 2]] - V[F[:, 0]])
    N = np.zeros_like(V)
    for k in range(3):
        np.add.at(N, F[:, k], fn)
    return N / np.linalg.norm(N, axis=1, keepdims=True)


def mesh_arrays(V, F):
    N = vertex_normals(V, F)
    return V[F].reshape(-1, 9), N[F].reshape(-1, 9)


def _as_loaded(tri):
    """What the .dae route does to a triangle list (n, 9): positions pass through a float32 text field (collada.cpp:621-636),
    and the half-edge mesh hands each face back starting from its LAST file vertex (v2, v0, v1)."""
    t = np.asarray(tri, float).reshape(-1, 3, 3)
    return np.roll(t, 1, axis=1).reshape(-1, 9)


def cb_mesh_scene(V, F, mesh_bsdf=(0, (0.5, 0.5, 0.5), (0, 0, 0), 0.0)):
    """Cornell box + one mesh object (the mesh comes first, as in CBbunny.dae / CBdragon.dae).  Same primitive order,
    vertex rotation and positions as write_cb_mesh_dae() + the .dae loader give (tests/test_host.py checks positions for
    equality and the half-edge vertex normals to 1 ulp); V must already be float32-representable."""
    mp, mn = mesh_arrays(V, F)
    bp, bn, bb, bt, bpar, lt, lp = cornell_box()
    mp, mn, bn = _as_loaded(mp), _as_loaded(mn), _as_loaded(bn)
    bp = _as_loaded(np.asarray(bp, np.float32).astype(np.float64))
    nm = len(mp)
    btype, a, b, ior = mesh_bsdf
    bsdf_type = np.concatenate([[btype], bt]).astype(np.int32)
    row = np.zeros((1, 8), np.float32); row[0, :3] = a; row[0, 3:6] = b; row[0, 6] = ior
    return {
        "prim_type": np.ones(nm + len(bp), np.int32),
        "prim_bsdf": np.concatenate([np.zeros(nm, np.int32), bb + 1]).astype(np.int32),
        "tri_pos": np.concatenate([mp, bp]), "tri_nrm": np.concatenate([mn, bn]),
        "sphere": np.zeros((nm + len(bp), 4)),
        "bsdf_type": bsdf_type, "bsdf_param": np.concatenate([row, bpar]).astype(np.float32),
        "light_type": lt, "light_param": lp,
    }


def cbdragon_standin(W=1920, H=1080):
    """Stand-in for BASELINE.json configs[1] (CBdragon.dae is missing from the checkout): Cornell box + a
    100 012-triangle closed diffuse mesh, cam_dragon.info.  Returns (scene arrays, camera)."""
    V, F = torus_knot()
    V = V.astype(np.float32).astype(np.float64)         # what survives the .dae's float text fields
    return cb_mesh_scene(V, F), cam_dragon(W, H)


def cblucy_standin(W=1920, H=1080):
    """Stand-in for configs[2] (CBlucy.dae missing): 133 796 triangles = 22 x 3041 tube... uses glass (ior 1.45)."""
    V, F = torus_knot(n_around=26, n_along=2573, tube=0.05)      # 133 796 triangles
    V = V.astype(np.float32).astype(np.float64)
    sc = cb_mesh_scene(V, F, mesh_bsdf=(3, (1, 1, 1), (1, 1, 1), 1.45))
    return sc, cam_dragon(W, H)


STANDINS = {
    # name: (mesh generator kwargs, mesh BSDF (type, a, b, ior))
    "cbdragon_standin": (dict(), (0, (0.5, 0.5, 0.5), (0, 0, 0), 0.0)),
    "cblucy_standin": (dict(n_around=26, n_along=2573, tube=0.05), (3, (1, 1, 1), (1, 1, 1), 1.45)),
}


def write_standin(name, directory, W=1920, H=1080):
    """Writes <directory>/<name>.dae + cam_dragon.info for a stand-in scene; returns (dae path, camera-file path).
    Both arms of bench.py and the parity tests start from these two files: the reference through its own ColladaParser,
    the product through csrc/host/scene_loader.cpp."""
    import os
    kw, bsdf = STANDINS[name]
    V, F = torus_knot(**kw)
    V = V.astype(np.float32).astype(np.float64)
    dae = os.path.join(directory, name + ".dae"); cam = os.path.join(directory, "cam_dragon.info")
    write_cb_mesh_dae(dae, V, F, mesh_bsdf=bsdf)
    write_cam_info(cam, cam_dragon(W, H))
    return dae, cam


def load_standin(name, W=1920, H=1080):
    """Stand-in scene through the .dae import path (product loader): (scene arrays, camera)."""
    import tempfile
    from . import load_dae
    with tempfile.TemporaryDirectory() as td:
        dae, cam = write_standin(name, td, W, H)
        return load_dae(dae, W, H, cam)


def triangle_soup(n_tris, seed=0x5EED, W=3840, H=2160):
    """configs[4]: centroids uniform in [-1,1]^3, edge vectors uniform in [-s,s]^3 with
    s = 0.5 * N^(-1/3), geometric normals, one diffuse material (albedo 0.7), one hemisphere light, camera at
    (0,0,3.5) looking at the origin, vfov 40 degrees."""
    rng = np.random.Generator(np.random.Philox(key=seed))
    c = rng.uniform(-1, 1, (n_tris, 3))
    s = 0.5 * n_tris ** (-1.0 / 3.0)
    e1 = rng.uniform(-s, s, (n_tris, 3)); e2 = rng.uniform(-s, s, (n_tris, 3))
    p1 = c - (e1 + e2) / 3; p2 = p1 + e1; p3 = p1 + e2
    n = np.cross(e1, e2); n /= np.maximum(np.linalg.norm(n, axis=1, keepdims=True), 1e-30)
    bsdf_param = np.zeros((1, 8), np.float32); bsdf_param[0, :3] = 0.7
    lp = np.zeros((1, 28)); lp[0, :3] = 1.0
    lp[0, 16:25] = [1, 0, 0, 0, 0, -1, 0, 1, 0]      # InfiniteHemisphereLight::sampleToWorld columns, light.cpp:28-32
    sc = {"prim_type": np.ones(n_tris, np.int32), "prim_bsdf": np.zeros(n_tris, np.int32),
          "tri_pos": np.concatenate([p1, p2, p3], axis=1), "tri_nrm": np.tile(n, (1, 3)), "sphere": np.zeros((n_tris, 4)),
          "bsdf_type": np.zeros(1, np.int32), "bsdf_param": bsdf_param, "light_type": np.array([1], np.int32), "light_param": lp}
    return sc, look_at_camera((0, 0, 3.5), (0, 0, 0), W, H, vfov=40.0)


# ---- COLLADA writer ---------------------------------------------------------------------------------------------
def _effect_xml(eid, btype, a, b, ior):
    phong = ("<profile_COMMON><technique sid=\"common\"><phong><diffuse><color sid=\"diffuse\">%.9g %.9g %.9g 1</color>"
             "</diffuse></phong></technique></profile_COMMON>" % tuple(a if btype == 0 else (0.5, 0.5, 0.5)))
    extra = ""
    if btype == 4:
        extra = "<emission><radiance>%.9g %.9g %.9g</radiance></emission>" % tuple(a)
    elif btype == 1:
        extra = "<mirror><reflectance>%.9g %.9g %.9g</reflectance></mirror>" % tuple(a)
    elif btype == 2:
        extra = ("<refraction><transmittance>%.9g %.9g %.9g</transmittance><roughness>0</roughness><ior>%.9g</ior>"
                 "</refraction>" % (tuple(b) + (ior,)))
    elif btype == 3:
        extra = ("<glass><reflectance>%.9g %.9g %.9g</reflectance><transmittance>%.9g %.9g %.9g</transmittance>"
                 "<roughness>0</roughness><ior>%.9g</ior></glass>" % (tuple(a) + tuple(b) + (ior,)))
    if extra:
        extra = "<extra><technique profile=\"CMU462\">%s</technique></extra>" % extra
    return "<effect id=\"%s-effect\">%s%s</effect>" % (eid, phong, extra)


def write_dae(path, meshes, bsdfs, hfov=CB_HFOV, aspect=CB_ASPECT, area_light_radiance=(10, 10, 10)):
    """Writes a Y_UP COLLADA file the reference's ColladaParser (src/collada/collada.cpp) accepts.
    meshes: list of (name, V[n,3], F[m,3], bsdf index) in world space (node matrices are identity);
    bsdfs: list of (type, a, b, ior).  One CMU462 area light is placed like the reference's CB scenes
    (position (0,1.49,0), direction -y, dim_x 0.6, dim_y 0.8).  Positions are written with 9 significant digits
    (the parser reads them as float, collada.cpp:621-636)."""
    out = ['<?xml version="1.0" encoding="utf-8"?>',
           '<COLLADA xmlns="http://www.collada.org/2005/11/COLLADASchema" version="1.4.1">',
           '<asset><unit name="meter" meter="1"/><up_axis>Y_UP</up_axis></asset>',
           '<library_lights><light id="Area-light" name="Light"><technique_common><point><color sid="color">1 1 1</color>'
           '<constant_attenuation>1</constant_attenuation><linear_attenuation>0</linear_attenuation>'
           '<quadratic_attenuation>0</quadratic_attenuation></point></technique_common><extra><technique profile="CMU462">'
           '<area><color sid="color">%.9g %.9g %.9g</color></area></technique></extra></light></library_lights>' % tuple(area_light_radiance),
           '<library_cameras><camera id="Camera-camera" name="Camera"><optics><technique_common><perspective>'
           '<xfov sid="xfov">%.9g</xfov><aspect_ratio>%.9g</aspect_ratio><znear sid="znear">0.1</znear><zfar sid="zfar">100</zfar>'
           '</perspective></technique_common></optics></camera></library_cameras>' % (hfov, aspect),
           '<library_effects>']
    for i, (bt, a, b, ior) in enumerate(bsdfs):
        out.append(_effect_xml("m%d" % i, bt, a, b, ior))
    out.append('</library_effects><library_materials>')
    for i in range(len(bsdfs)):
        out.append('<material id="m%d" name="m%d"><instance_effect url="#m%d-effect"/></material>' % (i, i, i))
    out.append('</library_materials><library_geometries>')
    for name, V, F, _ in meshes:
        V = np.asarray(V, np.float32); F = np.asarray(F)
        out.append('<geometry id="%s-mesh" name="%s"><mesh><source id="%s-mesh-positions"><float_array id="%s-mesh-positions-array" count="%d">'
                   % (name, name, name, name, V.size))
        out.append(" ".join("%.9g" % x for x in V.reshape(-1)))
        out.append('</float_array><technique_common><accessor source="#%s-mesh-positions-array" count="%d" stride="3">'
                   '<param name="X" type="float"/><param name="Y" type="float"/><param name="Z" type="float"/></accessor>'
                   '</technique_common></source><vertices id="%s-mesh-vertices"><input semantic="POSITION" source="#%s-mesh-positions"/>'
                   '</vertices><polylist count="%d"><input semantic="VERTEX" source="#%s-mesh-vertices" offset="0"/><vcount>'
                   % (name, len(V), name, name, len(F), name))
        out.append(" ".join(["3"] * len(F)))
        out.append('</vcount><p>')
        out.append(" ".join(str(int(x)) for x in F.reshape(-1)))
        out.append('</p></polylist></mesh></geometry>')
    out.append('</library_geometries><library_visual_scenes><visual_scene id="Scene" name="Scene">')
    out.append('<node id="Area" name="Area" type="NODE"><matrix sid="transform">-0.6 0 0 0 0 0 1 1.49 0 0.8 0 0 0 0 0 1</matrix>'
               '<instance_light url="#Area-light"/></node>')
    out.append('<node id="Camera" name="Camera" type="NODE"><matrix sid="transform">1 0 0 0 0 1 0 0.75 0 0 1 4.8 0 0 0 1</matrix>'
               '<instance_camera url="#Camera-camera"/></node>')
    for name, _, _, bi in meshes:
        out.append('<node id="%s" name="%s" type="NODE"><matrix sid="transform">1 0 0 0 0 1 0 0 0 0 1 0 0 0 0 1</matrix>'
                   '<instance_geometry url="#%s-mesh"><bind_material><technique_common><instance_material symbol="m%d" target="#m%d"/>'
                   '</technique_common></bind_material></instance_geometry></node>' % (name, name, name, bi, bi))
    out.append('</visual_scene></library_visual_scenes><scene><instance_visual_scene url="#Scene"/></scene></COLLADA>')
    with open(path, "w") as f:
        f.write("\n".join(out))


def write_cam_info(path, cam):
    """cam_*.info format (Application::saveCamera / loadCamera, reference src/application.cpp:801-853):
    pos / targetPos / phi / theta / minR / maxR / c2w row-major.  Only pos and c2w reach generate_ray."""
    cam = np.asarray(cam, float)
    c2w = cam[3:12].reshape(3, 3).T
    with open(path, "w") as f:
        f.write("%.17g %.17g %.17g\n0 0 0\n0\n0\n0\n0\n" % tuple(cam[:3]))
        f.write(" ".join("%.17g" % x for x in c2w.reshape(-1)) + " \n")


def write_cb_mesh_dae(path, V, F, mesh_bsdf=(0, (0.5, 0.5, 0.5), (0, 0, 0), 0.0)):
    """The Cornell-box + mesh scene of cb_mesh_scene() as a .dae (mesh object first, then the six box objects)."""
    bsdfs = [mesh_bsdf, (0, (0.6,) * 3, (0,) * 3, 0), (4, (10,) * 3, (0,) * 3, 0), (0, (0.6,) * 3, (0,) * 3, 0),
             (0, (0.6, 0.2, 0.2), (0,) * 3, 0), (0, (0.2, 0.2, 0.6), (0,) * 3, 0), (0, (0.6,) * 3, (0,) * 3, 0)]
    meshes = [("mesh", V, F, 0)]
    names = ["ceiling", "light", "floor", "leftWall", "rightWall", "backWall"]
    for q in range(6):
        t0, t1 = np.array(_CB_TRIS[2 * q], float), np.array(_CB_TRIS[2 * q + 1], float)
        verts = np.concatenate([t0, t1])
        meshes.append((names[q], verts[[0, 1, 2, 4]], np.array([[0, 1, 2], [0, 3, 1]]), q + 1))
    write_dae(path, meshes, bsdfs)
