"""ctypes binding of the CPU oracle (oracle/_build/liboracle.so).  TEST INFRASTRUCTURE ONLY.

Importers allowed: tests/, __graft_entry__.smoke(), bench.py's cpu_baseline / --impl reference legs.
Nothing under vrad_b200/ may import this module.  PARITY UNPINNED -- see oracle/oracle.h.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "liboracle.so")

TRI48_DTYPE = np.dtype([("n", "<f4", 3), ("d", "<f4"), ("id", "<i4"), ("e", "<f4", 6),
                        ("sel0", "u1"), ("sel1", "u1"), ("flags", "u1"), ("unused", "u1")])
assert TRI48_DTYPE.itemsize == 48


def build(force: bool = False) -> str:
    srcs = [os.path.join(_HERE, f) for f in os.listdir(_HERE) if f.endswith((".cpp", ".h", ".hpp", "Makefile"))]
    stale = force or not os.path.exists(_LIB_PATH) or any(
        os.path.getmtime(s) > os.path.getmtime(_LIB_PATH) for s in srcs)
    if stale:
        subprocess.run(["make", "-C", _HERE], check=True, capture_output=True)
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        _lib.orc_env_create.restype = C.c_void_p
        _lib.orc_env_build_seconds.restype = C.c_double
        _lib.orc_env_build_seconds.argtypes = [C.c_void_p]
    return _lib


def _p(a, t=None):
    if a is None:
        return None
    return a.ctypes.data_as(C.c_void_p)


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


class OracleEnv:
    """CPU oracle environment: same call surface as vrad_b200.Environment."""

    def __init__(self):
        self._l = lib()
        self._h = C.c_void_p(self._l.orc_env_create())
        self.n_patches = 0

    def close(self):
        if self._h:
            self._l.orc_env_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- geometry / build --------------------------------------------------
    def add_triangles(self, ids, verts9, flags=None):
        ids = np.ascontiguousarray(ids, np.int32); verts9 = _f32(verts9).reshape(-1, 9)
        flags = None if flags is None else np.ascontiguousarray(flags, np.uint8)
        rc = self._l.orc_env_add_triangles(self._h, C.c_int(ids.shape[0]), _p(ids), _p(verts9), _p(flags))
        assert rc == 0

    def build(self):
        assert self._l.orc_env_build(self._h) == 0
        return self._l.orc_env_build_seconds(self._h)

    def sizes(self):
        v = [C.c_int() for _ in range(5)]
        assert self._l.orc_env_sizes(self._h, *[C.byref(x) for x in v]) == 0
        return dict(zip(("n_nodes", "n_idx", "n_tris", "max_depth", "n_leaves"), [x.value for x in v]))

    def export(self):
        s = self.sizes()
        children = np.empty(s["n_nodes"], np.int32); split = np.empty(s["n_nodes"], np.float32)
        tri_index = np.empty(s["n_idx"], np.int32); tris = np.empty(s["n_tris"], TRI48_DTYPE)
        aabb = np.empty(6, np.float32)
        assert self._l.orc_env_export(self._h, _p(children), _p(split), _p(tri_index), _p(tris), _p(aabb)) == 0
        return {"children": children, "split": split, "tri_index": tri_index, "tris": tris, "aabb": aabb}

    def replace_tree(self, children, split, tri_index, aabb):
        """Walk a foreign kd tree (reference layout) over the same triangles."""
        c = np.ascontiguousarray(children, np.int32); s = np.ascontiguousarray(split, np.float32)
        t = np.ascontiguousarray(tri_index, np.int32); a = np.ascontiguousarray(aabb, np.float32)
        assert self._l.orc_env_replace_tree(self._h, C.c_int(c.shape[0]), _p(c), _p(s), C.c_int(t.shape[0]), _p(t), _p(a)) == 0

    # -- tracing -----------------------------------------------------------
    def _trace(self, fn, o, d, tmax, tmin=None, skip_id=-1, threads=1):
        o = _f32(o); d = _f32(d); tmax = _f32(tmax)
        n = o.shape[1]
        tmin = None if tmin is None else _f32(tmin)
        hit_tri = np.empty(n, np.int32); hit_sid = np.empty(n, np.int32); hit_t = np.empty(n, np.float32)
        rc = fn(self._h, C.c_int64(n), _p(o[0]), _p(o[1]), _p(o[2]), _p(d[0]), _p(d[1]), _p(d[2]),
                _p(tmin), _p(tmax), C.c_int32(skip_id), _p(hit_tri), _p(hit_sid), _p(hit_t), C.c_int(threads))
        assert rc == 0
        return hit_tri, hit_sid, hit_t

    def trace_brute(self, o, d, tmax, tmin=None, skip_id=-1, threads=1):
        return self._trace(self._l.orc_trace_brute, o, d, tmax, tmin, skip_id, threads)

    def trace1(self, o, d, tmax, tmin=None, skip_id=-1, threads=1):
        return self._trace(self._l.orc_trace1, o, d, tmax, tmin, skip_id, threads)

    def trace4(self, o, d, tmax, tmin=None, skip_id=-1, threads=1):
        return self._trace(self._l.orc_trace4, o, d, tmax, tmin, skip_id, threads)

    def trace4_packet(self, origin_xyz4, dir_xyz4, tmin, tmax, skip_id=-1):
        o = _f32(origin_xyz4).reshape(12); d = _f32(dir_xyz4).reshape(12)
        tmin = _f32(tmin); tmax = _f32(tmax)
        ids = np.empty(4, np.int32); dist = np.empty(4, np.float32); nrm = np.empty(12, np.float32)
        assert self._l.orc_trace4_packet(self._h, _p(o), _p(d), _p(tmin), _p(tmax), C.c_int32(skip_id),
                                         _p(ids), _p(dist), _p(nrm)) == 0
        return ids, dist, nrm.reshape(3, 4)

    def counters(self):
        v = [C.c_int64() for _ in range(3)]
        self._l.orc_trace_counters(self._h, *[C.byref(x) for x in v])
        return dict(zip(("nodes", "tris", "leaves"), [x.value for x in v]))

    def test_lines(self, start_soa, stop_soa, sky_mode=0, mode=0, threads=1):
        s = _f32(start_soa); e = _f32(stop_soa)
        n = s.shape[1]
        bits = np.zeros((n + 31) // 32, np.uint32)
        assert self._l.orc_test_lines(self._h, C.c_int64(n), _p(s), _p(e), C.c_int(sky_mode), _p(bits),
                                      C.c_int(mode), C.c_int(threads)) == 0
        return bits

    # -- full TestLineDoesHitSky surface (oracle/skytrace.cpp) --------------
    def set_triangle_colors(self, rgb):
        rgb = _f32(rgb).reshape(-1, 3)
        assert self._l.orc_env_set_triangle_colors(self._h, C.c_int(rgb.shape[0]), _p(rgb)) == 0

    def bsp_set(self, bsp):
        i32 = lambda a: np.ascontiguousarray(a, np.int32)
        npl, nch = i32(bsp.node_plane), i32(bsp.node_children).reshape(-1, 2)
        pn, pd, pt = _f32(bsp.plane_normal).reshape(-1, 3), _f32(bsp.plane_dist), i32(bsp.plane_type)
        lc, la = i32(bsp.leaf_cluster), i32(bsp.leaf_area)
        assert self._l.orc_bsp_set(self._h, C.c_int(npl.shape[0]), _p(npl), _p(nch), C.c_int(pd.shape[0]), _p(pn), _p(pd), _p(pt),
                                   C.c_int(lc.shape[0]), _p(lc), _p(la), C.c_int(int(bsp.n_areas))) == 0
        self._n_areas = int(bsp.n_areas)

    def point_leafnum(self, pts):
        pts = _f32(pts).reshape(-1, 3); out = np.empty(pts.shape[0], np.int32)
        assert self._l.orc_point_leafnum(self._h, C.c_int64(pts.shape[0]), _p(pts), _p(out)) == 0
        return out

    def cluster_from_point(self, pts):
        pts = _f32(pts).reshape(-1, 3); out = np.empty(pts.shape[0], np.int32)
        assert self._l.orc_cluster_from_point(self._h, C.c_int64(pts.shape[0]), _p(pts), _p(out)) == 0
        return out

    def process_sky_cameras(self, origins, scales):
        o = _f32(origins).reshape(-1, 3); s = _f32(scales).reshape(-1)
        n = self._l.orc_sky_cameras_set(self._h, C.c_int(o.shape[0]), _p(o), _p(s))
        assert n >= 0
        return n

    def sky_cameras(self):
        n = self._l.orc_sky_cameras_get(self._h, None, None, None)
        cam_area = np.empty(n, np.int32); w2s = np.empty(n, np.float32); area_cam = np.empty(self._n_areas, np.int32)
        self._l.orc_sky_cameras_get(self._h, _p(cam_area), _p(w2s), _p(area_cam))
        return cam_area, w2s, area_cam

    def test_lines_sky(self, start_soa, stop_soa, flags=1, static_prop_to_skip=-1, threads=1):
        s = _f32(start_soa); e = _f32(stop_soa)
        n = s.shape[1]
        out = np.empty(n, np.float32)
        assert self._l.orc_test_lines_sky(self._h, C.c_int64(n), _p(s), _p(e), C.c_int(flags), C.c_int32(static_prop_to_skip),
                                          _p(out), C.c_int(threads)) == 0
        return out

    def leafs_trace_to_sky(self, mins, maxs, dirs3, threads=1):
        mins = np.ascontiguousarray(mins, np.int16).reshape(-1, 3); maxs = np.ascontiguousarray(maxs, np.int16).reshape(-1, 3)
        d = _f32(dirs3).reshape(-1, 3)
        out = np.empty(mins.shape[0], np.uint8)
        assert self._l.orc_leafs_trace_to_sky(self._h, C.c_int(mins.shape[0]), _p(mins), _p(maxs), C.c_int(d.shape[0]), _p(d),
                                              _p(out), C.c_int(threads)) == 0
        return out

    # -- radiosity ---------------------------------------------------------
    def patches_upload(self, origin, normal, plane_dist, area, refl, cluster=None, flags=None):
        origin = _f32(origin); normal = _f32(normal); plane_dist = _f32(plane_dist); area = _f32(area); refl = _f32(refl)
        n = origin.shape[0]
        cluster = None if cluster is None else np.ascontiguousarray(cluster, np.int32)
        flags = None if flags is None else np.ascontiguousarray(flags, np.uint8)
        assert self._l.orc_patches_set(self._h, C.c_int(n), _p(origin), _p(normal), _p(plane_dist), _p(area),
                                       _p(refl), _p(cluster), _p(flags)) == 0
        self.n_patches = n

    def set_hierarchy(self, parent, child1, child2, face=None):
        i32 = lambda a: None if a is None else np.ascontiguousarray(a, np.int32)
        parent, child1, child2, face = i32(parent), i32(child1), i32(child2), i32(face)
        assert self._l.orc_patches_set_hierarchy(self._h, C.c_int(parent.shape[0]), _p(parent), _p(child1), _p(child2), _p(face)) == 0

    def set_windings(self, first, count, points):
        """Patch.Winding per patch (None removes them): near pairs then use the polygon-to-differential form factor."""
        if first is None:
            assert self._l.orc_patches_set_windings(self._h, C.c_int(0), None, None, C.c_int(0), None) == 0
            return
        first = np.ascontiguousarray(first, np.int32); count = np.ascontiguousarray(count, np.int32); points = _f32(points).reshape(-1, 3)
        assert self._l.orc_patches_set_windings(self._h, C.c_int(first.shape[0]), _p(first), _p(count), C.c_int(points.shape[0]), _p(points)) == 0

    def build_transfers(self, pvs=None, threads=1):
        nnz = C.c_int64()
        nc = 0
        if pvs is not None:
            pvs = np.ascontiguousarray(pvs, np.uint8); nc = pvs.shape[0]
        assert self._l.orc_build_transfers(self._h, C.c_int(nc), _p(pvs), C.byref(nnz), C.c_int(threads)) == 0
        self.nnz = nnz.value
        return nnz.value

    def transfers(self):
        rowptr = np.empty(self.n_patches + 1, np.int64); col = np.empty(self.nnz, np.int32); w = np.empty(self.nnz, np.float32)
        assert self._l.orc_transfers_get(self._h, _p(rowptr), _p(col), _p(w)) == 0
        return rowptr, col, w

    def transfer_row(self, i, pvs=None, cap=1 << 20):
        nc = 0
        if pvs is not None:
            pvs = np.ascontiguousarray(pvs, np.uint8); nc = pvs.shape[0]
        col = np.empty(cap, np.int32); w = np.empty(cap, np.float32)
        self._l.orc_transfer_row.restype = C.c_int64
        n = self._l.orc_transfer_row(self._h, C.c_int(int(i)), C.c_int(nc), _p(pvs), _p(col), _p(w), C.c_int64(cap))
        assert n >= 0
        return col[:n].copy(), w[:n].copy()

    def set_sky_dirs(self, dirs3):
        d = _f32(dirs3).reshape(-1, 3)
        assert self._l.orc_set_sky_dirs(C.c_int(d.shape[0]), _p(d)) == 0

    def set_light_trace_flags(self, flags):
        assert self._l.orc_env_set_light_trace_flags(self._h, C.c_int(flags)) == 0

    def direct_light(self, pos, normal, lights, threads=1):
        pos = _f32(pos); normal = _f32(normal)
        lights = np.ascontiguousarray(lights)
        assert lights.dtype.itemsize == 96
        out = np.empty((pos.shape[0], 3), np.float32)
        assert self._l.orc_direct_light(self._h, C.c_int64(pos.shape[0]), _p(pos), _p(normal),
                                        C.c_int(lights.shape[0]), _p(lights), _p(out), C.c_int(threads)) == 0
        return out

    def set_bump(self, needs_bump, bump_normals):
        nb = np.ascontiguousarray(needs_bump, np.uint8); bn = _f32(bump_normals).reshape(-1, 9)
        assert self._l.orc_patches_set_bump(self._h, C.c_int(nb.shape[0]), _p(nb), _p(bn)) == 0

    def bump_totals(self):
        out = np.empty((self.n_patches, 3, 3), np.float32)
        assert self._l.orc_bounce_bump_totals(self._h, _p(out)) == 0
        return out

    def bounce(self, emit0, n_bounces, early_out=False, threads=1):
        emit0 = _f32(emit0)
        total = np.empty_like(emit0); added = np.empty(3, np.float32); done = C.c_int()
        assert self._l.orc_bounce(self._h, _p(emit0), C.c_int(n_bounces), C.c_int(int(early_out)), _p(total),
                                  _p(added), C.byref(done), C.c_int(threads)) == 0
        return total, added, done.value


def gather_rows(row0, row1, rowptr, col, w, emit, refl, threads=1):
    rowptr = np.ascontiguousarray(rowptr, np.int64); col = np.ascontiguousarray(col, np.int32)
    w = _f32(w); emit = _f32(emit); refl = _f32(refl)
    out = np.empty((row1 - row0, 3), np.float32)
    lib().orc_gather_rows(C.c_int64(row0), C.c_int64(row1), _p(rowptr), _p(col), _p(w), _p(emit), _p(refl),
                          _p(out), C.c_int(threads))
    return out


def subdivide_patches(faces, points, min_chop=4.0):
    """oracle/patches.cpp; faces: structured array with the vrad_face_patch fields."""
    points = _f32(points).reshape(-1, 3)
    col = lambda k, t: np.ascontiguousarray(faces[k], t)
    args_in = [col("first_point", np.int32), col("n_points", np.int32), col("normal", np.float32), col("plane_dist", np.float32),
               col("lux_scale", np.float32), col("chop", np.float32), col("sky", np.uint8), col("no_subdivide", np.uint8),
               col("has_base_light", np.uint8)]
    n_, m_ = C.c_int(), C.c_int()
    L = lib()
    rc = L.orc_patches_subdivide(C.c_int(len(faces)), *[_p(a) for a in args_in], _p(points), C.c_float(min_chop),
                                 C.byref(n_), C.byref(m_), *([None] * 14))
    assert rc == 0
    n, m = n_.value, m_.value
    f3 = lambda: np.empty((n, 3), np.float32)
    f1 = lambda: np.empty(n, np.float32)
    i1 = lambda: np.empty(n, np.int32)
    out = dict(origin=f3(), normal=f3(), plane_dist=f1(), area=f1(), mins=f3(), maxs=f3(), chop=f1(), parent=i1(), child1=i1(),
               child2=i1(), face=i1(), wind_first=i1(), wind_count=i1(), wind_points=np.empty((m, 3), np.float32))
    keys = ("origin", "normal", "plane_dist", "area", "mins", "maxs", "chop", "parent", "child1", "child2", "face",
            "wind_first", "wind_count", "wind_points")
    rc = L.orc_patches_subdivide(C.c_int(len(faces)), *[_p(a) for a in args_in], _p(points), C.c_float(min_chop),
                                 C.byref(n_), C.byref(m_), *[_p(out[k]) for k in keys])
    assert rc == 0
    return out


def bump_normals(s_vect, t_vect, flat_normal, phong_normal):
    out = np.empty((3, 3), np.float32)
    lib().orc_bump_normals(_p(_f32(s_vect)), _p(_f32(t_vect)), _p(_f32(flat_normal)), _p(_f32(phong_normal)), _p(out))
    return out


def light_for_string(value: str):
    out = np.zeros(3, np.float32)
    rc = lib().orc_light_for_string(C.c_char_p(value.encode()), _p(out))
    return out, rc


def lights_from_entities(ents, light_dtype):
    ents = np.ascontiguousarray(ents)
    assert ents.dtype.itemsize == 124
    out = np.zeros(2 * max(1, ents.shape[0]), light_dtype)
    n = lib().orc_lights_from_entities(C.c_int(ents.shape[0]), _p(ents), C.c_int(out.shape[0]), _p(out))
    assert n >= 0
    return out[:n].copy()


def lights_from_patches(origin, normal, base_light, area, scale2, base_area, child1, light_threshold, light_dtype):
    origin = _f32(origin).reshape(-1, 3); normal = _f32(normal).reshape(-1, 3); base_light = _f32(base_light).reshape(-1, 3)
    area = _f32(area); scale2 = _f32(scale2).reshape(-1, 2); base_area = _f32(base_area)
    child1 = None if child1 is None else np.ascontiguousarray(child1, np.int32)
    out = np.zeros(max(1, origin.shape[0]), light_dtype)
    n = lib().orc_lights_from_patches(C.c_int(origin.shape[0]), _p(origin), _p(normal), _p(base_light), _p(area), _p(scale2), _p(base_area),
                                      _p(child1), C.c_float(light_threshold), C.c_int(out.shape[0]), _p(out))
    assert n >= 0
    return out[:n].copy()


def decompress_vis(data: bytes, n_clusters: int):
    buf = np.frombuffer(bytes(data), np.uint8)
    out = np.zeros((n_clusters + 7) // 8, np.uint8)
    lib().orc_decompress_vis.restype = C.c_int64
    used = lib().orc_decompress_vis(_p(buf), C.c_int64(buf.shape[0]), C.c_int(n_clusters), _p(out))
    return out, int(used)


def num_threads():
    return int(lib().orc_num_threads())


def env_from_scene(scene, with_patches=True) -> OracleEnv:
    e = OracleEnv()
    e.add_triangles(scene.tri_ids, scene.tri_verts, scene.tri_flags)
    e.build()
    if with_patches and scene.patch_origin is not None:
        e.patches_upload(scene.patch_origin, scene.patch_normal, scene.patch_plane_dist, scene.patch_area,
                         scene.patch_refl, scene.patch_cluster, scene.patch_flags)
    return e
