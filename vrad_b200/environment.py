"""Host-side mirror of the reference's ray-tracing call surface, on top of the C-ABI.

`Environment` follows raytracer.Environment (raytracer/environment.go:28-39): AddTriangle (:41),
AddTriangleWithMaterial (:45), AddQuad (:71), AddAxisAlignedRectangularSolid (:77),
SetupAccelerationStructure (:119), Trace4Rays (:140), GetTriangle (:422).  `test_line_does_hit_sky`
follows trace.TestLineDoesHitSky (raytracer/trace/testline.go:18-94).  The batched methods
(`trace_rays`, `test_lines`, `build_transfers`, `direct_light`, `bounce`) are the throughput path.

Arrays may be numpy (host) or torch CUDA tensors (device-resident, zero-copy).  Everything runs in
libvradcuda.so on the GPU; nothing here computes on the CPU and there is no fallback.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import lib as _lib
from .lib import TRI48_DTYPE, VradConfig, VradError, check, ptr

TRACE_ID_SKY = 0x01000000         # raytracer/constants.go:9-11
TRACE_ID_OPAQUE = 0x02000000
TRACE_ID_STATICPROP = 0x04000000
KDNODE_STATE_LEAF = 3             # raytracer/constants.go:16
TRI_TRANSPARENT = 0x01            # TriIntersectData.NFlags bit (upstream FCACHETRI_TRANSPARENT)
TL_CAN_RECURSE, TL_TEXTURE_SHADOWS, TL_PACKET_LEAF = 1, 2, 4   # vrad_test_lines_sky flags


def _is_torch(a):
    return hasattr(a, "data_ptr")


def _f32(a):
    if a is None or _is_torch(a):
        return a
    return np.ascontiguousarray(a, dtype=np.float32)


class Environment:
    def __init__(self, device: int = 0, rank: int = 0, world: int = 1, devices=None):
        """devices = a list of CUDA ordinals: ONE handle that drives them all from this process (vrad_env_create_multi) --
        host buffers only, calls synchronous; the rank / world arguments are for the one-process-per-GPU form."""
        self._l = _lib.load()
        self._h = C.c_void_p()
        if devices is not None:
            devices = [int(d) for d in devices]
            mc = _lib.VradMultiConfig(len(devices), (C.c_int * 8)(*(devices + [0] * (8 - len(devices)))), 0)
            check(self._l.vrad_env_create_multi(C.byref(mc), C.byref(self._h)))
            device, rank, world = devices[0], 0, 1
        else:
            cfg = VradConfig(device, rank, world, 0)
            check(self._l.vrad_env_create(C.byref(cfg), C.byref(self._h)))
        self.device, self.rank, self.world, self.devices = device, rank, world, devices
        self.n_patches = 0
        self._pending_ids, self._pending_verts, self._pending_flags = [], [], []

    # ---- lifetime ----------------------------------------------------------------------
    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            self._l.vrad_env_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_stream(self, cuda_stream: int | None):
        """None: the library's own stream.  An integer cudaStream_t handle otherwise; torch's default
        stream has handle 0, which is passed as cudaStreamLegacy (0x1) so that it is not mistaken for None."""
        if cuda_stream is None:
            handle = 0
        else:
            handle = int(cuda_stream) or 1
        check(self._l.vrad_env_set_stream(self._h, C.c_void_p(handle)))

    def set_async(self, flag: bool):
        check(self._l.vrad_env_set_async(self._h, C.c_int(int(flag))))

    def set_option(self, name: str, value: int):
        """Tuning switches (include/vrad_cuda.h: vrad_env_set_option): k1_sort, k1_top, k4_seg, k4_long_first, k4_pdl, k4_graph, k4_sim_peers."""
        check(self._l.vrad_env_set_option(self._h, C.c_char_p(name.encode()), C.c_int(int(value))))

    def last_timing(self):
        ms, nl = C.c_float(), C.c_int()
        check(self._l.vrad_env_last_timing(self._h, C.byref(ms), C.byref(nl)))
        return ms.value, nl.value

    # ---- geometry (reference names) ----------------------------------------------------
    def add_triangle(self, tri_id, v1, v2, v3, colour=None):                       # environment.go:41-43
        self.add_triangle_with_material(tri_id, v1, v2, v3, colour, 0, 0)

    def add_triangle_with_material(self, tri_id, v1, v2, v3, colour=None, flags=0, material_index=0):  # :45-69
        self._pending_ids.append(int(tri_id))
        self._pending_verts.append([*v1, *v2, *v3])
        self._pending_flags.append(int(flags) & 0xFF)

    def add_quad(self, tri_id, v1, v2, v3, v4, colour=None):                       # :71-75
        self.add_triangle(tri_id, v1, v2, v3, colour)
        self.add_triangle(tri_id + 1, v1, v3, v4, colour)

    def add_axis_aligned_rectangular_solid(self, tri_id, mn, mx, colour=None):     # :77-117
        q = self.add_quad
        q(tri_id, (mn[0], mx[1], mx[2]), (mx[0], mx[1], mx[2]), (mx[0], mn[1], mx[2]), (mn[0], mn[1], mx[2]))
        q(tri_id, (mn[0], mx[1], mn[2]), (mx[0], mx[1], mn[2]), (mx[0], mn[1], mn[2]), (mn[0], mn[1], mn[2]))
        q(tri_id, (mn[0], mx[1], mx[2]), (mn[0], mx[1], mn[2]), (mn[0], mn[1], mn[2]), (mn[0], mn[1], mx[2]))
        q(tri_id, (mx[0], mx[1], mx[2]), (mx[0], mx[1], mn[2]), (mx[0], mn[1], mn[2]), (mx[0], mn[1], mx[2]))
        q(tri_id, (mn[0], mx[1], mx[2]), (mx[0], mx[1], mx[2]), (mx[0], mx[1], mn[2]), (mn[0], mx[1], mn[2]))
        q(tri_id, (mn[0], mn[1], mx[2]), (mx[0], mn[1], mx[2]), (mx[0], mn[1], mn[2]), (mn[0], mn[1], mn[2]))

    def add_triangles(self, ids, verts9, flags=None):
        """Bulk form of AddTriangleWithMaterial."""
        self._flush_pending()
        ids = np.ascontiguousarray(ids, np.int32)
        verts9 = np.ascontiguousarray(verts9, np.float32).reshape(-1, 9)
        assert verts9.shape[0] == ids.shape[0]
        flags = None if flags is None else np.ascontiguousarray(flags, np.uint8)
        check(self._l.vrad_env_add_triangles(self._h, C.c_int(ids.shape[0]), ptr(ids), ptr(verts9), ptr(flags)))

    def _flush_pending(self):
        if self._pending_ids:
            ids = np.asarray(self._pending_ids, np.int32)
            verts = np.asarray(self._pending_verts, np.float32).reshape(-1, 9)
            flags = np.asarray(self._pending_flags, np.uint8)
            self._pending_ids, self._pending_verts, self._pending_flags = [], [], []
            check(self._l.vrad_env_add_triangles(self._h, C.c_int(ids.shape[0]), ptr(ids), ptr(verts), ptr(flags)))

    def setup_acceleration_structure(self):                                         # :119-138
        self._flush_pending()
        check(self._l.vrad_env_build(self._h))
        return self.stats()["build_seconds"]

    def build_fast(self, on_host: bool | None = False):
        """RTE_FLAGS_FAST_TREE_GENERATION (raytracer/constants.go:5): the binned-SAH builder, on the device, on the host's cores, or
        (on_host=None) wherever it is faster for the triangle count."""
        self._flush_pending()
        check(self._l.vrad_env_build_fast(self._h, C.c_int(2 if on_host is None else (1 if on_host else 0))))
        return self.stats()["build_seconds"]

    def upload_tree(self, children, split, tri_index, tris, aabb):
        children = np.ascontiguousarray(children, np.int32); split = np.ascontiguousarray(split, np.float32)
        tri_index = np.ascontiguousarray(tri_index, np.int32); tris = np.ascontiguousarray(tris)
        assert tris.dtype.itemsize == 48
        aabb = np.ascontiguousarray(aabb, np.float32)
        check(self._l.vrad_env_upload_tree(self._h, C.c_int(children.shape[0]), ptr(children), ptr(split),
                                           C.c_int(tri_index.shape[0]), ptr(tri_index), C.c_int(tris.shape[0]), ptr(tris), ptr(aabb)))

    def stats(self):
        v = [C.c_int() for _ in range(5)]
        aabb = np.empty(6, np.float32); secs = C.c_double()
        check(self._l.vrad_env_stats(self._h, *[C.byref(x) for x in v], ptr(aabb), C.byref(secs)))
        d = dict(zip(("n_nodes", "n_idx", "n_tris", "max_depth", "n_leaves"), [x.value for x in v]))
        d["aabb"] = aabb; d["build_seconds"] = secs.value
        return d

    def download_tree(self):
        s = self.stats()
        children = np.empty(s["n_nodes"], np.int32); split = np.empty(s["n_nodes"], np.float32)
        tri_index = np.empty(s["n_idx"], np.int32); tris = np.empty(s["n_tris"], TRI48_DTYPE)
        check(self._l.vrad_env_download_tree(self._h, ptr(children), ptr(split), ptr(tri_index), ptr(tris)))
        return {"children": children, "split": split, "tri_index": tri_index, "tris": tris, "aabb": s["aabb"]}

    def get_triangle(self, index: int):                                             # :422-424
        return self.download_tree()["tris"][index]

    # ---- K1 ----------------------------------------------------------------------------
    def trace4_rays(self, origin_xyz4, dir_xyz4, tmin, tmax, skip_id=-1):           # :140-145
        """One FourRays packet -> (HitIds[4], HitDistance[4], SurfaceNormal[3,4])."""
        o = np.ascontiguousarray(origin_xyz4, np.float32).reshape(12)
        d = np.ascontiguousarray(dir_xyz4, np.float32).reshape(12)
        tmin = np.ascontiguousarray(tmin, np.float32); tmax = np.ascontiguousarray(tmax, np.float32)
        ids = np.empty(4, np.int32); dist = np.empty(4, np.float32); nrm = np.empty(12, np.float32)
        check(self._l.vrad_trace4(self._h, ptr(o), ptr(d), ptr(tmin), ptr(tmax), C.c_int32(skip_id), ptr(ids), ptr(dist), ptr(nrm)))
        return ids, dist, nrm.reshape(3, 4)

    def trace_rays(self, o, d, tmax, tmin=None, skip_id=-1, out=None):
        """Batched closest hit.  o, d: [3, n] SoA.  Returns (hit_tri, hit_sid, hit_t)."""
        o = _f32(o); d = _f32(d); tmax = _f32(tmax); tmin = _f32(tmin)
        n = int(o.shape[1])
        if out is None:
            if _is_torch(o):
                import torch
                out = (torch.empty(n, dtype=torch.int32, device=o.device), torch.empty(n, dtype=torch.int32, device=o.device),
                       torch.empty(n, dtype=torch.float32, device=o.device))
            else:
                out = (np.empty(n, np.int32), np.empty(n, np.int32), np.empty(n, np.float32))
        check(self._l.vrad_trace_rays(self._h, C.c_int64(n), ptr(o[0]), ptr(o[1]), ptr(o[2]), ptr(d[0]), ptr(d[1]), ptr(d[2]),
                                      ptr(tmin), ptr(tmax), C.c_int32(skip_id), ptr(out[0]), ptr(out[1]), ptr(out[2])))
        return out

    def test_lines(self, start_soa, stop_soa, sky_mode=0, out=None):
        """Batched TestLine: [3, n] SoA endpoints -> uint32 visibility words (bit = 1: visible)."""
        s = _f32(start_soa); e = _f32(stop_soa)
        n = int(s.shape[1])
        nw = (n + 31) // 32
        if out is None:
            if _is_torch(s):
                import torch
                out = torch.empty(nw, dtype=torch.int32, device=s.device)
            else:
                out = np.empty(nw, np.uint32)
        check(self._l.vrad_test_lines(self._h, C.c_int64(n), ptr(s), ptr(e), C.c_int(sky_mode), ptr(out)))
        return out

    def points_upload(self, xyz):
        """Resident endpoint table for test_lines_indexed: [P, 3] points (patch origins, light origins, luxel samples ...)."""
        xyz = _f32(xyz).reshape(-1, 3)
        check(self._l.vrad_points_upload(self._h, C.c_int64(xyz.shape[0]), ptr(xyz)))
        self.n_points = int(xyz.shape[0])

    def test_lines_indexed(self, pairs, sky_mode=0, out=None):
        """Batched TestLine over index pairs [n, 2] (int32) into the uploaded point table -> uint32 visibility words."""
        if not _is_torch(pairs):
            pairs = np.ascontiguousarray(pairs, np.int32)
        n = int(pairs.shape[0])
        nw = (n + 31) // 32
        if out is None:
            if _is_torch(pairs):
                import torch
                out = torch.empty(nw, dtype=torch.int32, device=pairs.device)
            else:
                out = np.empty(nw, np.uint32)
        check(self._l.vrad_test_lines_indexed(self._h, C.c_int64(n), ptr(pairs), C.c_int(sky_mode), ptr(out)))
        return out

    # ---- full TestLineDoesHitSky surface: colours, BSP lumps, sky cameras (section 8 f2) -----
    def set_triangle_colors(self, rgb):                                             # environment.go:61-63, :430-432
        self._flush_pending()
        rgb = np.ascontiguousarray(rgb, np.float32).reshape(-1, 3)
        check(self._l.vrad_env_set_triangle_colors(self._h, C.c_int(rgb.shape[0]), ptr(rgb)))

    def bsp_upload(self, bsp):
        """bsp: scenes.Bsp (Nodes/Planes/Leafs lumps as flat arrays)."""
        i32 = lambda a: np.ascontiguousarray(a, np.int32)
        node_plane, node_children = i32(bsp.node_plane), i32(bsp.node_children).reshape(-1, 2)
        pn, pd, pt = _f32(bsp.plane_normal).reshape(-1, 3), _f32(bsp.plane_dist), i32(bsp.plane_type)
        lc, la = i32(bsp.leaf_cluster), i32(bsp.leaf_area)
        check(self._l.vrad_bsp_upload(self._h, C.c_int(node_plane.shape[0]), ptr(node_plane), ptr(node_children),
                                      C.c_int(pd.shape[0]), ptr(pn), ptr(pd), ptr(pt), C.c_int(lc.shape[0]), ptr(lc), ptr(la),
                                      C.c_int(int(bsp.n_areas))))
        self._n_areas = int(bsp.n_areas)

    def point_leafnum(self, pts):                                                   # pointleaf.go:8-33
        pts = _f32(pts).reshape(-1, 3)
        out = np.empty(pts.shape[0], np.int32)
        check(self._l.vrad_point_leafnum(self._h, C.c_int64(pts.shape[0]), ptr(pts), ptr(out)))
        return out

    def cluster_from_point(self, pts):                                              # clustertable/point.go:10-38
        pts = _f32(pts).reshape(-1, 3)
        out = np.empty(pts.shape[0], np.int32)
        check(self._l.vrad_cluster_from_point(self._h, C.c_int64(pts.shape[0]), ptr(pts), ptr(out)))
        return out

    def process_sky_cameras(self, origins, scales):                                 # rad/cameras/skycamera.go:10-49
        o = _f32(origins).reshape(-1, 3); s = _f32(scales).reshape(-1)
        kept = C.c_int()
        check(self._l.vrad_sky_cameras_set(self._h, C.c_int(o.shape[0]), ptr(o), ptr(s), C.byref(kept)))
        return kept.value

    def sky_cameras(self):
        n = C.c_int()
        check(self._l.vrad_sky_cameras_get(self._h, C.byref(n), None, None, None))
        cam_area = np.empty(n.value, np.int32); w2s = np.empty(n.value, np.float32)
        area_cam = np.empty(getattr(self, "_n_areas", 0), np.int32)
        check(self._l.vrad_sky_cameras_get(self._h, C.byref(n), ptr(cam_area), ptr(w2s), ptr(area_cam)))
        return cam_area, w2s, area_cam

    def test_lines_sky(self, start_soa, stop_soa, flags=TL_CAN_RECURSE, static_prop_to_skip=-1, out=None):
        """Batched, complete TestLineDoesHitSky: [3, n] SoA endpoints -> fractionVisible[n]."""
        s = _f32(start_soa); e = _f32(stop_soa)
        n = int(s.shape[1])
        if out is None:
            if _is_torch(s):
                import torch
                out = torch.empty(n, dtype=torch.float32, device=s.device)
            else:
                out = np.empty(n, np.float32)
        check(self._l.vrad_test_lines_sky(self._h, C.c_int64(n), ptr(s), ptr(e), C.c_int(flags), C.c_int32(static_prop_to_skip), ptr(out)))
        return out

    def leafs_trace_to_sky(self, mins, maxs):                                       # lightmap.go:425-451
        mins = np.ascontiguousarray(mins, np.int16).reshape(-1, 3); maxs = np.ascontiguousarray(maxs, np.int16).reshape(-1, 3)
        out = np.empty(mins.shape[0], np.uint8)
        check(self._l.vrad_leafs_trace_to_sky(self._h, C.c_int(mins.shape[0]), ptr(mins), ptr(maxs), ptr(out)))
        return out

    # ---- patches / K2 / K3 / K4 --------------------------------------------------------
    def patches_upload(self, origin, normal, plane_dist, area, refl, cluster=None, flags=None):
        origin = _f32(origin); normal = _f32(normal); plane_dist = _f32(plane_dist); area = _f32(area); refl = _f32(refl)
        n = int(origin.shape[0])
        cluster = None if cluster is None else np.ascontiguousarray(cluster, np.int32)
        flags = None if flags is None else np.ascontiguousarray(flags, np.uint8)
        check(self._l.vrad_patches_upload(self._h, C.c_int(n), ptr(origin), ptr(normal), ptr(plane_dist), ptr(area), ptr(refl), ptr(cluster), ptr(flags)))
        self.n_patches = n

    def set_hierarchy(self, parent, child1, child2, face=None):
        """Patch.Parent / Child1 / Child2 / FaceNumber for the uploaded patches (common/types/patch.go:33,49-51)."""
        i32 = lambda a: None if a is None else np.ascontiguousarray(a, np.int32)
        parent, child1, child2, face = i32(parent), i32(child1), i32(child2), i32(face)
        check(self._l.vrad_patches_set_hierarchy(self._h, C.c_int(parent.shape[0]), ptr(parent), ptr(child1), ptr(child2), ptr(face)))

    def set_windings(self, first, count, points):
        """Patch.Winding per uploaded patch (vrad_patches_subdivide's wind_first / wind_count / wind_points; None removes them):
        build_transfers then uses the polygon-to-differential form factor for emitters that are large for their distance."""
        if first is None:
            check(self._l.vrad_patches_set_windings(self._h, C.c_int(0), None, None, C.c_int(0), None))
            return
        first = np.ascontiguousarray(first, np.int32); count = np.ascontiguousarray(count, np.int32)
        points = np.ascontiguousarray(points, np.float32).reshape(-1, 3)
        check(self._l.vrad_patches_set_windings(self._h, C.c_int(first.shape[0]), ptr(first), ptr(count), C.c_int(points.shape[0]), ptr(points)))

    def set_bump(self, needs_bump, bump_normals):
        """Patch.NeedsBumpMap + the three bump normals per patch ([N, 3, 3]); bounce() then accumulates TotalLight.Light[1..3]."""
        nb = np.ascontiguousarray(needs_bump, np.uint8); bn = np.ascontiguousarray(bump_normals, np.float32).reshape(-1, 9)
        check(self._l.vrad_patches_set_bump(self._h, C.c_int(nb.shape[0]), ptr(nb), ptr(bn)))

    def bump_totals(self):
        out = np.empty((self.n_patches, 3, 3), np.float32)
        check(self._l.vrad_bounce_bump_totals(self._h, ptr(out)))
        return out

    def build_transfers(self, pvs=None):
        nnz = C.c_int64(); nc = 0
        if pvs is not None:
            pvs = np.ascontiguousarray(pvs, np.uint8); nc = int(pvs.shape[0])
        check(self._l.vrad_build_transfers(self._h, C.c_int(nc), ptr(pvs), C.byref(nnz)))
        return nnz.value

    def transfers_upload(self, row0, row1, rowptr, col, w):
        rowptr = np.ascontiguousarray(rowptr, np.int64); col = np.ascontiguousarray(col, np.int32); w = np.ascontiguousarray(w, np.float32)
        check(self._l.vrad_transfers_upload(self._h, C.c_int64(row0), C.c_int64(row1), ptr(rowptr), ptr(col), ptr(w)))

    def transfers_info(self):
        a, b, c = C.c_int64(), C.c_int64(), C.c_int64()
        check(self._l.vrad_transfers_info(self._h, C.byref(a), C.byref(b), C.byref(c)))
        return a.value, b.value, c.value

    def transfers_layout(self):
        """(entries of the {col,w} pair array, entries of the packed 6-byte streams, their segments, entries of the block-row streams,
        rows per block); 0 where a form is not in use."""
        a, b, c, d, r = C.c_int64(), C.c_int64(), C.c_int64(), C.c_int64(), C.c_int64()
        check(self._l.vrad_transfers_layout(self._h, C.byref(a), C.byref(b), C.byref(c), C.byref(d), C.byref(r)))
        return a.value, b.value, c.value, d.value, r.value

    def transfers_download(self):
        row0, row1, nnz = self.transfers_info()
        rowptr = np.empty(row1 - row0 + 1, np.int64); col = np.empty(nnz, np.int32); w = np.empty(nnz, np.float32)
        check(self._l.vrad_transfers_download(self._h, ptr(rowptr), ptr(col), ptr(w)))
        return rowptr, col, w

    def transfers_download_rows(self, row_begin, row_end, capacity=None):
        """Rows [row_begin,row_end) of the resident transfer lists (for matrices too large to download whole)."""
        if capacity is None:
            capacity = 1 << 22
        rowptr = np.empty(row_end - row_begin + 1, np.int64); col = np.empty(capacity, np.int32); w = np.empty(capacity, np.float32)
        check(self._l.vrad_transfers_download_rows(self._h, C.c_int64(row_begin), C.c_int64(row_end), ptr(rowptr), ptr(col), ptr(w), C.c_int64(capacity)))
        n = int(rowptr[-1])
        return rowptr, col[:n].copy(), w[:n].copy()

    def set_sky_dirs(self, dirs3):
        d = np.ascontiguousarray(dirs3, np.float32).reshape(-1, 3)
        check(self._l.vrad_set_sky_dirs(self._h, C.c_int(d.shape[0]), ptr(d)))

    def set_light_trace_flags(self, flags: int):
        """TL_CAN_RECURSE / TL_TEXTURE_SHADOWS for the light rays of direct_light (0 = binary TestLine)."""
        check(self._l.vrad_set_light_trace_flags(self._h, C.c_int(flags)))

    def direct_light(self, pos, normal, lights, out=None):
        pos = _f32(pos); normal = _f32(normal)
        lights = np.ascontiguousarray(lights)
        assert lights.dtype.itemsize == 96
        n = int(pos.shape[0])
        if out is None:
            if _is_torch(pos):
                import torch
                out = torch.empty((n, 3), dtype=torch.float32, device=pos.device)
            else:
                out = np.empty((n, 3), np.float32)
        check(self._l.vrad_direct_light(self._h, C.c_int64(n), ptr(pos), ptr(normal), C.c_int(lights.shape[0]), ptr(lights), ptr(out)))
        return out

    def bounce(self, emit0, n_bounces, early_out=False, out=None, want_added=True):
        emit0 = _f32(emit0)
        if out is None:
            if _is_torch(emit0):
                import torch
                out = torch.empty_like(emit0)
            else:
                out = np.empty_like(emit0)
        added = np.zeros(3, np.float32) if want_added else None
        done = C.c_int()
        check(self._l.vrad_bounce(self._h, ptr(emit0), C.c_int(n_bounces), C.c_int(int(early_out)), ptr(out), ptr(added), C.byref(done)))
        return out, added, done.value

    # ---- multi-GPU ---------------------------------------------------------------------
    @staticmethod
    def comm_unique_id() -> bytes:
        buf = C.create_string_buffer(128)
        check(_lib.load().vrad_comm_unique_id(buf))
        return buf.raw

    def comm_init(self, unique_id: bytes):
        assert len(unique_id) == 128
        check(self._l.vrad_comm_init(self._h, C.c_char_p(unique_id)))


def environment_from_scene(scene, device=0, rank=0, world=1, with_patches=True, devices=None) -> Environment:
    env = Environment(device, rank, world, devices=devices)
    env.add_triangles(scene.tri_ids, scene.tri_verts, scene.tri_flags)
    env.setup_acceleration_structure()
    if with_patches and scene.patch_origin is not None:
        env.patches_upload(scene.patch_origin, scene.patch_normal, scene.patch_plane_dist, scene.patch_area,
                           scene.patch_refl, scene.patch_cluster, scene.patch_flags)
    return env


def test_line_does_hit_sky(env: Environment, start_xyz4, stop_xyz4, can_recurse=True, static_prop_to_skip=-1, do_debug=False,
                           texture_shadows=False):
    """trace.TestLineDoesHitSky for one FourVectors pair (raytracer/trace/testline.go:18-94).

    Returns fractionVisible[4].  As in the reference the leaf that decides the 3D-skybox recursion is
    the one of lane 0 (`start.Vec(0)`, :63); `texture_shadows` is the package variable of :14.
    """
    s = np.ascontiguousarray(start_xyz4, np.float32).reshape(3, 4)
    e = np.ascontiguousarray(stop_xyz4, np.float32).reshape(3, 4)
    flags = TL_PACKET_LEAF | (TL_CAN_RECURSE if can_recurse else 0) | (TL_TEXTURE_SHADOWS if texture_shadows else 0)
    return env.test_lines_sky(s, e, flags, static_prop_to_skip)


PATCH_TREE_FIELDS = ("origin", "normal", "plane_dist", "area", "mins", "maxs", "chop", "parent", "child1", "child2", "face",
                     "wind_first", "wind_count", "wind_points")


def subdivide_patches(faces, points, min_chop=4.0):
    """patches.MakePatchForFace + SubdividePatches (rad/patches/face.go:29-197, subdivide.go:25-437) on the host.
    faces: lib.FACE_PATCH_DTYPE records; points: [P, 3] winding points.  Returns a dict of arrays (PATCH_TREE_FIELDS)."""
    L = _lib.load()
    faces = np.ascontiguousarray(faces, _lib.FACE_PATCH_DTYPE)
    points = np.ascontiguousarray(points, np.float32).reshape(-1, 3)
    np_, npt = C.c_int(), C.c_int()
    rc = L.vrad_patches_subdivide(C.c_int(faces.shape[0]), ptr(faces), ptr(points), C.c_float(min_chop), C.c_int(0), C.c_int(0),
                                  C.byref(np_), C.byref(npt), *([None] * 14))
    if rc not in (0, -3):
        check(rc)
    n, m = np_.value, npt.value
    f3 = lambda: np.empty((n, 3), np.float32)
    f1 = lambda: np.empty(n, np.float32)
    i1 = lambda: np.empty(n, np.int32)
    out = dict(origin=f3(), normal=f3(), plane_dist=f1(), area=f1(), mins=f3(), maxs=f3(), chop=f1(), parent=i1(), child1=i1(),
               child2=i1(), face=i1(), wind_first=i1(), wind_count=i1(), wind_points=np.empty((m, 3), np.float32))
    check(L.vrad_patches_subdivide(C.c_int(faces.shape[0]), ptr(faces), ptr(points), C.c_float(min_chop), C.c_int(n), C.c_int(m),
                                   C.byref(np_), C.byref(npt), *[ptr(out[k]) for k in PATCH_TREE_FIELDS]))
    return out


def bump_normals(s_vect, t_vect, flat_normal, phong_normal):
    """upstream GetBumpNormals for one face -> [3, 3] bump-basis normals (host-only)."""
    out = np.empty((3, 3), np.float32)
    check(_lib.load().vrad_bump_normals(ptr(_f32(s_vect)), ptr(_f32(t_vect)), ptr(_f32(flat_normal)), ptr(_f32(phong_normal)), ptr(out)))
    return out


def light_for_string(value: str):
    """Entity.LightForString (common/types/entity.go:103-158): "r g b [scale]" -> linear RGB intensity (raises on a bad value)."""
    out = np.zeros(3, np.float32)
    check(_lib.load().vrad_light_for_string(C.c_char_p(value.encode()), ptr(out)))
    return out


def lights_from_entities(ents):
    """CreateDirectLights, entity part (rad/lightmap/lights.go:90-426): lib.LIGHT_ENTITY_DTYPE records -> scenes.LIGHT_DTYPE records."""
    from .scenes import LIGHT_DTYPE
    ents = np.ascontiguousarray(ents, _lib.LIGHT_ENTITY_DTYPE)
    out = np.zeros(2 * max(1, ents.shape[0]), LIGHT_DTYPE)
    n = C.c_int()
    check(_lib.load().vrad_lights_from_entities(C.c_int(ents.shape[0]), ptr(ents), C.c_int(out.shape[0]), ptr(out), C.byref(n)))
    return out[: n.value].copy()


def lights_from_patches(origin, normal, base_light, area, scale2, base_area, child1=None, light_threshold=0.1):
    """CreateDirectLights, surface part (lights.go:49-82): EMIT_SURFACE lights of the emitting leaf patches."""
    from .scenes import LIGHT_DTYPE
    origin = _f32(origin).reshape(-1, 3); normal = _f32(normal).reshape(-1, 3); base_light = _f32(base_light).reshape(-1, 3)
    area = _f32(area); scale2 = _f32(scale2).reshape(-1, 2); base_area = _f32(base_area)
    child1 = None if child1 is None else np.ascontiguousarray(child1, np.int32)
    n_p = origin.shape[0]
    out = np.zeros(max(1, n_p), LIGHT_DTYPE)
    n = C.c_int()
    check(_lib.load().vrad_lights_from_patches(C.c_int(n_p), ptr(origin), ptr(normal), ptr(base_light), ptr(area), ptr(scale2), ptr(base_area),
                                               ptr(child1), C.c_float(light_threshold), C.c_int(out.shape[0]), ptr(out), C.byref(n)))
    return out[: n.value].copy()


def decompress_vis(data: bytes, n_clusters: int):
    """lightmap.DecompressVis (rad/lightmap/vis.go:54-94): one PVS row -> ((n_clusters+7)//8 bytes, input bytes used)."""
    buf = np.frombuffer(bytes(data), np.uint8)
    out = np.zeros((n_clusters + 7) // 8, np.uint8)
    used = _lib.load().vrad_decompress_vis(ptr(buf), C.c_int64(buf.shape[0]), C.c_int(n_clusters), ptr(out))
    if used < 0:
        raise VradError(int(used), _lib.load().vrad_last_error().decode("utf-8", "replace"))
    return out, int(used)


def pvs_from_vis_lump(n_clusters: int, byteofs, visdata: bytes):
    """The visibility lump -> the [C, C] byte matrix `Environment.build_transfers` takes (vis.go:9-47)."""
    ofs = np.ascontiguousarray(byteofs, np.int32).reshape(-1, 2)
    assert ofs.shape[0] == n_clusters
    buf = np.frombuffer(bytes(visdata), np.uint8)
    out = np.empty((n_clusters, n_clusters), np.uint8)
    check(_lib.load().vrad_pvs_from_vis_lump(C.c_int(n_clusters), ptr(ofs), ptr(buf), C.c_int64(buf.shape[0]), ptr(out)))
    return out


def kd_build_binned_host(verts9):
    """The binned-SAH builder run on the host's cores, without an environment: the tree in reference layout."""
    v = np.ascontiguousarray(verts9, np.float32).reshape(-1, 9)
    l = _lib.load()
    nn, ni, depth = C.c_int(), C.c_int(), C.c_int()
    aabb = np.zeros(6, np.float32)
    check(l.vrad_kd_build_binned_host(C.c_int(v.shape[0]), ptr(v), C.c_int(0), C.c_int(0), None, None, None, C.byref(nn), C.byref(ni), ptr(aabb), C.byref(depth)))
    children = np.zeros(nn.value, np.int32); split = np.zeros(nn.value, np.float32); tri_index = np.zeros(max(ni.value, 1), np.int32)
    check(l.vrad_kd_build_binned_host(C.c_int(v.shape[0]), ptr(v), C.c_int(nn.value), C.c_int(ni.value), ptr(children), ptr(split), ptr(tri_index),
                                      C.byref(nn), C.byref(ni), ptr(aabb), C.byref(depth)))
    return {"children": children, "split": split, "tri_index": tri_index[:ni.value], "aabb": aabb, "max_depth": depth.value}


def row_partition(n_rows: int, world: int):
    """Row ranges owned by each rank (equal blocks of ceil(n/world) rounded up to a multiple of 4 -- the gather's block-row streams group
    4 consecutive global rows; the same rule the library uses)."""
    rpr = ((n_rows + world - 1) // world + 3) & ~3
    return [(min(n_rows, r * rpr), min(n_rows, (r + 1) * rpr)) for r in range(world)]
