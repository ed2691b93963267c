"""Deterministic synthetic scenes for the hot path (BASELINE.json configs).

The generator plays the role of the reference's input producers -- brush -> triangle extraction
(cmd/tasks/loadbsp/main.go:241-340: world triangles carry TRACE_ID_OPAQUE, sky faces TRACE_ID_SKY),
face -> patch creation (rad/patches/face.go:30-197: origin lifted off the face along the normal,
reflectivity clamped, area) and light creation (rad/lightmap/lights.go:210-341) -- and emits the
flat arrays the C-ABI takes.  Geometry is appended exactly the way raytracer.Environment's
AddQuad / AddAxisAlignedRectangularSolid do (raytracer/environment.go:71-117), so triangle order
and ids are what the reference's own helper calls would produce.

RNG: splitmix64 -> 24-bit uniform floats; seeds are fixed per scene.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np

TRACE_ID_SKY = 0x01000000         # raytracer/constants.go:9
TRACE_ID_OPAQUE = 0x02000000      # raytracer/constants.go:10
TRACE_ID_STATICPROP = 0x04000000  # raytracer/constants.go:11
MAX_TRACE_LENGTH = np.float32(1.732050807569 * 32768.0)   # common/constants/constants.go:15-19

EMIT_SURFACE, EMIT_POINT, EMIT_SPOTLIGHT, EMIT_SKYLIGHT, EMIT_SKYAMBIENT = 0, 1, 2, 3, 5

# 96-byte light record == vrad_light in include/vrad_cuda.h
LIGHT_DTYPE = np.dtype([
    ("type", "<i4"), ("origin", "<f4", 3), ("intensity", "<f4", 3), ("normal", "<f4", 3),
    ("stopdot", "<f4"), ("stopdot2", "<f4"), ("exponent", "<f4"), ("radius", "<f4"),
    ("constant_attn", "<f4"), ("linear_attn", "<f4"), ("quadratic_attn", "<f4"),
    ("start_fade", "<f4"), ("end_fade", "<f4"), ("cap_dist", "<f4"),
    ("flags", "<i4"), ("pad", "<f4", 3)])
assert LIGHT_DTYPE.itemsize == 96

_MASK = (1 << 64) - 1


class SplitMix64:
    """Scalar + vectorised splitmix64."""

    def __init__(self, seed: int):
        self.state = seed & _MASK

    def u64(self, n: int) -> np.ndarray:
        with np.errstate(over="ignore"):
            idx = np.arange(1, n + 1, dtype=np.uint64)
            z = np.uint64(self.state) + idx * np.uint64(0x9E3779B97F4A7C15)
            self.state = int(z[-1]) if n else self.state
            z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
            z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
            return z ^ (z >> np.uint64(31))

    def uniform(self, n: int, lo: float = 0.0, hi: float = 1.0) -> np.ndarray:
        u = (self.u64(n) >> np.uint64(40)).astype(np.float64) * (1.0 / (1 << 24))
        return (lo + (hi - lo) * u).astype(np.float32)

    def integers(self, n: int, hi: int) -> np.ndarray:
        return (self.u64(n) % np.uint64(hi)).astype(np.int64)


@dataclass
class Scene:
    name: str
    tri_ids: np.ndarray            # int32 [T]
    tri_verts: np.ndarray          # float32 [T, 9]
    tri_flags: np.ndarray          # uint8 [T]
    patch_origin: np.ndarray = None      # float32 [N,3]  (lifted 1 unit off the face)
    patch_normal: np.ndarray = None
    patch_plane_dist: np.ndarray = None
    patch_area: np.ndarray = None
    patch_refl: np.ndarray = None        # float32 [N,3]
    patch_cluster: np.ndarray = None     # int32 [N]
    patch_flags: np.ndarray = None       # uint8 [N], bit0 = sky
    n_clusters: int = 1
    pvs: np.ndarray = None               # uint8 [C, C] or None (everything visible)
    luxel_pos: np.ndarray = None         # float32 [L,3]
    luxel_normal: np.ndarray = None
    lights: np.ndarray = None            # LIGHT_DTYPE [n]
    meta: dict = field(default_factory=dict)

    @property
    def n_tris(self): return int(self.tri_ids.shape[0])

    @property
    def n_patches(self): return 0 if self.patch_origin is None else int(self.patch_origin.shape[0])


class _Geom:
    """Mirror of the reference's triangle-append helpers (raytracer/environment.go:41-117)."""

    def __init__(self):
        self.ids, self.verts = [], []

    def add_triangle(self, tid, v1, v2, v3):
        self.ids.append(tid)
        self.verts.append([*v1, *v2, *v3])

    def add_quad(self, tid, v1, v2, v3, v4):           # environment.go:71-75
        self.add_triangle(tid, v1, v2, v3)
        self.add_triangle(tid + 1, v1, v3, v4)

    def add_box(self, tid, mn, mx):                    # environment.go:77-117 (same face/vertex order)
        q = self.add_quad
        q(tid, (mn[0], mx[1], mx[2]), (mx[0], mx[1], mx[2]), (mx[0], mn[1], mx[2]), (mn[0], mn[1], mx[2]))
        q(tid, (mn[0], mx[1], mn[2]), (mx[0], mx[1], mn[2]), (mx[0], mn[1], mn[2]), (mn[0], mn[1], mn[2]))
        q(tid, (mn[0], mx[1], mx[2]), (mn[0], mx[1], mn[2]), (mn[0], mn[1], mn[2]), (mn[0], mn[1], mx[2]))
        q(tid, (mx[0], mx[1], mx[2]), (mx[0], mx[1], mn[2]), (mx[0], mn[1], mn[2]), (mx[0], mn[1], mx[2]))
        q(tid, (mn[0], mx[1], mx[2]), (mx[0], mx[1], mx[2]), (mx[0], mx[1], mn[2]), (mn[0], mx[1], mn[2]))
        q(tid, (mn[0], mn[1], mx[2]), (mx[0], mn[1], mx[2]), (mx[0], mn[1], mn[2]), (mn[0], mn[1], mn[2]))

    def add_box_array(self, tid, mins, maxs):
        for mn, mx in zip(mins, maxs):
            self.add_box(tid, mn, mx)

    def arrays(self):
        ids = np.asarray(self.ids, dtype=np.int32)
        verts = np.asarray(self.verts, dtype=np.float32).reshape(-1, 9)
        return ids, verts, np.zeros(ids.shape[0], dtype=np.uint8)


def _grid_points(origin, udir, vdir, nu, nv, cell, normal, lift):
    """Centres of an nu x nv grid of square cells on a planar face, lifted off the face."""
    origin = np.asarray(origin, np.float64); udir = np.asarray(udir, np.float64)
    vdir = np.asarray(vdir, np.float64); normal = np.asarray(normal, np.float64)
    iu, iv = np.meshgrid(np.arange(nu), np.arange(nv), indexing="ij")
    c = (origin[None, :] + (iu.reshape(-1, 1) + 0.5) * cell * udir[None, :]
         + (iv.reshape(-1, 1) + 0.5) * cell * vdir[None, :])
    centres = c.astype(np.float32)
    pos = (c + lift * normal[None, :]).astype(np.float32)
    nrm = np.broadcast_to(normal.astype(np.float32), pos.shape).copy()
    return centres, pos, nrm


def _room_faces(x0, y0, z0, sx, sy, sz):
    """Six inward-facing faces of a room: (origin, udir, vdir, ulen, vlen, normal)."""
    x1, y1, z1 = x0 + sx, y0 + sy, z0 + sz
    return [
        ((x0, y0, z0), (1, 0, 0), (0, 1, 0), sx, sy, (0, 0, 1)),     # floor
        ((x0, y0, z1), (1, 0, 0), (0, 1, 0), sx, sy, (0, 0, -1)),    # ceiling
        ((x0, y0, z0), (0, 1, 0), (0, 0, 1), sy, sz, (1, 0, 0)),     # wall x = x0
        ((x1, y0, z0), (0, 1, 0), (0, 0, 1), sy, sz, (-1, 0, 0)),    # wall x = x1
        ((x0, y0, z0), (1, 0, 0), (0, 0, 1), sx, sz, (0, 1, 0)),     # wall y = y0
        ((x0, y1, z0), (1, 0, 0), (0, 0, 1), sx, sz, (0, -1, 0)),    # wall y = y1
    ]


def _face_quad(g: _Geom, tid, face):
    o, u, v, ul, vl, _ = face
    o = np.asarray(o, np.float64); u = np.asarray(u, np.float64) * ul; v = np.asarray(v, np.float64) * vl
    g.add_quad(tid, tuple(o), tuple(o + u), tuple(o + u + v), tuple(o + v))


def _place_boxes(rng: SplitMix64, count, x0, y0, x1, y1, z0, smin=16.0, smax=96.0, margin=8.0):
    """Non-overlapping axis-aligned boxes resting on the floor z0 (rejection sampling)."""
    mins, maxs = [], []
    guard = 0
    while len(mins) < count:
        guard += 1
        if guard > 100000:
            raise RuntimeError("box placement did not converge")
        s = rng.uniform(3, smin, smax).astype(np.float64)
        p = rng.uniform(2).astype(np.float64)
        bx0 = x0 + margin + p[0] * (x1 - x0 - 2 * margin - s[0])
        by0 = y0 + margin + p[1] * (y1 - y0 - 2 * margin - s[1])
        mn = (bx0, by0, z0); mx = (bx0 + s[0], by0 + s[1], z0 + s[2])
        ok = True
        for m2, x2 in zip(mins, maxs):
            if mn[0] < x2[0] + 4 and mx[0] > m2[0] - 4 and mn[1] < x2[1] + 4 and mx[1] > m2[1] - 4:
                ok = False
                break
        if ok:
            mins.append(tuple(np.float32(v) for v in mn)); maxs.append(tuple(np.float32(v) for v in mx))
    return mins, maxs


def _make_lights(rng: SplitMix64, n_point, n_spot, x0, y0, x1, y1, zlo, zhi):
    n = n_point + n_spot
    lights = np.zeros(n, dtype=LIGHT_DTYPE)
    for k in range(n):
        p = rng.uniform(3)
        lights[k]["origin"] = (x0 + 64 + p[0] * (x1 - x0 - 128), y0 + 64 + p[1] * (y1 - y0 - 128), zlo + p[2] * (zhi - zlo))
        inten = rng.uniform(3, 100.0, 400.0)
        # quadratic falloff: c=0, l=0, q=1; "scale intensity for unit 100 distance" (lights.go:335-339)
        lights[k]["constant_attn"], lights[k]["linear_attn"], lights[k]["quadratic_attn"] = 0.0, 0.0, 1.0
        ratio = np.float32(0.0 + 100 * 0.0 + 100 * 100 * 1.0)
        lights[k]["intensity"] = inten * ratio
        lights[k]["start_fade"], lights[k]["end_fade"], lights[k]["cap_dist"] = 0.0, -1.0, 1.0e22   # light.go:38-44
        if k < n_point:
            lights[k]["type"] = EMIT_POINT
        else:
            lights[k]["type"] = EMIT_SPOTLIGHT
            d = rng.uniform(2, -0.5, 0.5).astype(np.float64)
            v = np.array([d[0], d[1], -1.0]); v /= np.linalg.norm(v)
            lights[k]["normal"] = v.astype(np.float32)
            lights[k]["stopdot"] = np.float32(math.cos(30.0 / 180.0 * math.pi))    # lights.go:251-252
            lights[k]["stopdot2"] = np.float32(math.cos(45.0 / 180.0 * math.pi))
            lights[k]["exponent"] = 1.0
    return lights


def _surface_samples(faces, cell, lift, skip=None):
    cs, ps, ns, fi = [], [], [], []
    for k, (o, u, v, ul, vl, nrm) in enumerate(faces):
        c, p, n = _grid_points(o, u, v, int(round(ul / cell)), int(round(vl / cell)), cell, nrm, lift)
        if skip is not None:
            keep = ~skip(k, c)
            c, p, n = c[keep], p[keep], n[keep]
        cs.append(c); ps.append(p); ns.append(n); fi.append(np.full(p.shape[0], k, np.int32))
    return np.concatenate(cs), np.concatenate(ps), np.concatenate(ns), np.concatenate(fi)


def box_room(seed: int = 0x5EED0001, n_boxes: int = 82, size=(1024.0, 1024.0, 512.0), patch_cell=32.0,
             luxel_cell=16.0, n_point=4, n_spot=4) -> Scene:
    """S1 (configs C1/C2): one room + box occluders: 12 + 12*n_boxes triangles (996 by default),
    4096 leaf patches, 16384 luxels, 8 lights."""
    rng = SplitMix64(seed)
    sx, sy, sz = size
    x0, y0, z0 = -sx / 2, -sy / 2, 0.0
    g = _Geom()
    faces = _room_faces(x0, y0, z0, sx, sy, sz)
    for f in faces:
        _face_quad(g, TRACE_ID_OPAQUE, f)
    mins, maxs = _place_boxes(rng, n_boxes, x0, y0, x0 + sx, y0 + sy, z0)
    g.add_box_array(TRACE_ID_OPAQUE, mins, maxs)
    ids, verts, flags = g.arrays()

    centres, origin, normal, face_idx = _surface_samples(faces, patch_cell, 1.0)
    refl_face = rng.uniform(6 * 3, 0.2, 0.7).reshape(6, 3)
    refl = np.minimum(refl_face[face_idx], np.float32(0.99))         # rad/patches/face.go:225-230
    plane_dist = np.einsum("ij,ij->i", normal.astype(np.float64), centres.astype(np.float64)).astype(np.float32)
    area = np.full(origin.shape[0], patch_cell * patch_cell, np.float32)
    _, lpos, lnrm, _ = _surface_samples(faces, luxel_cell, 1.0)
    lights = _make_lights(rng, n_point, n_spot, x0, y0, x0 + sx, y0 + sy, sz * 0.5, sz - 32.0)
    return Scene("S1_box_room", ids, verts, flags, origin, normal, plane_dist, area, refl.astype(np.float32),
                 np.zeros(origin.shape[0], np.int32), np.zeros(origin.shape[0], np.uint8), 1, None, lpos, lnrm, lights,
                 {"seed": seed, "n_boxes": n_boxes})


def multi_room(seed: int = 0x5EED0002, nx: int = 12, ny: int = 11, room: float = 512.0, boxes_per_room: int = 30,
               patch_cell: float = 32.0, door_w: float = 128.0, door_h: float = 256.0, pvs_radius: int = 2) -> Scene:
    """S2 (configs C3/C4): nx x ny rooms on a grid, shared zero-thickness walls with a centred door
    between neighbours, box occluders in every room.  One cluster per room; PVS = self + rooms reached
    through <= pvs_radius collinear doors."""
    rng = SplitMix64(seed)
    g = _Geom()
    R = room

    def wall_with_door(o, u, h_axis_len, door):
        # wall spans u in [0,R], z in [0,R]; door centred: three quads (left, right, above) or one solid quad
        o = np.asarray(o, np.float64); u = np.asarray(u, np.float64)
        z = np.array([0.0, 0.0, 1.0])
        def quad(u0, u1, z0, z1):
            g.add_quad(TRACE_ID_OPAQUE, tuple(o + u * u0 + z * z0), tuple(o + u * u1 + z * z0),
                       tuple(o + u * u1 + z * z1), tuple(o + u * u0 + z * z1))
        if not door:
            quad(0.0, R, 0.0, R)
        else:
            a, b = (R - door_w) / 2, (R + door_w) / 2
            quad(0.0, a, 0.0, R); quad(b, R, 0.0, R); quad(a, b, door_h, R)

    for i in range(nx):
        for j in range(ny):
            x0, y0 = i * R, j * R
            fl = _room_faces(x0, y0, 0.0, R, R, R)
            _face_quad(g, TRACE_ID_OPAQUE, fl[0]); _face_quad(g, TRACE_ID_OPAQUE, fl[1])
    for i in range(nx + 1):          # walls on planes x = i*R
        for j in range(ny):
            wall_with_door((i * R, j * R, 0.0), (0, 1, 0), R, 0 < i < nx)
    for j in range(ny + 1):          # walls on planes y = j*R
        for i in range(nx):
            wall_with_door((i * R, j * R, 0.0), (1, 0, 0), R, 0 < j < ny)
    for i in range(nx):
        for j in range(ny):
            mins, maxs = _place_boxes(rng, boxes_per_room, i * R, j * R, (i + 1) * R, (j + 1) * R, 0.0, margin=40.0)
            g.add_box_array(TRACE_ID_OPAQUE, mins, maxs)
    ids, verts, flags = g.arrays()

    a, b = (R - door_w) / 2, (R + door_w) / 2
    origins, normals, pdists, refls, clusters = [], [], [], [], []
    for i in range(nx):
        for j in range(ny):
            x0, y0 = i * R, j * R
            faces = _room_faces(x0, y0, 0.0, R, R, R)
            has_door = {2: i > 0, 3: i < nx - 1, 4: j > 0, 5: j < ny - 1}

            def skip(k, c, x0=x0, y0=y0, has_door=has_door):
                if k < 2 or not has_door[k]:
                    return np.zeros(c.shape[0], bool)
                along = (c[:, 1] - y0) if k in (2, 3) else (c[:, 0] - x0)
                return (along > a) & (along < b) & (c[:, 2] < door_h)
            centres, origin, normal, face_idx = _surface_samples(faces, patch_cell, 1.0, skip)
            rf = rng.uniform(6 * 3, 0.2, 0.7).reshape(6, 3)
            origins.append(origin); normals.append(normal)
            pdists.append(np.einsum("ij,ij->i", normal.astype(np.float64), centres.astype(np.float64)).astype(np.float32))
            refls.append(rf[face_idx]); clusters.append(np.full(origin.shape[0], i * ny + j, np.int32))
    origin = np.concatenate(origins); normal = np.concatenate(normals)
    nC = nx * ny
    pvs = np.zeros((nC, nC), np.uint8)
    for i in range(nx):
        for j in range(ny):
            c = i * ny + j
            pvs[c, c] = 1
            for r in range(1, pvs_radius + 1):
                for (di, dj) in ((r, 0), (-r, 0), (0, r), (0, -r)):
                    ii, jj = i + di, j + dj
                    if 0 <= ii < nx and 0 <= jj < ny:
                        pvs[c, ii * ny + jj] = 1
    # one point light per room near the ceiling centre (used by the direct-light stage of C3/C4)
    lights = np.zeros(nC, dtype=LIGHT_DTYPE)
    for i in range(nx):
        for j in range(ny):
            k = i * ny + j
            p = rng.uniform(2, -64.0, 64.0)
            lights[k]["type"] = EMIT_POINT
            lights[k]["origin"] = (i * R + R / 2 + p[0], j * R + R / 2 + p[1], R - 48.0)
            lights[k]["quadratic_attn"] = 1.0
            lights[k]["intensity"] = rng.uniform(3, 100.0, 400.0) * np.float32(10000.0)
            lights[k]["start_fade"], lights[k]["end_fade"], lights[k]["cap_dist"] = 0.0, -1.0, 1.0e22
    N = origin.shape[0]
    return Scene("S2_multi_room", ids, verts, flags, origin, normal, np.concatenate(pdists),
                 np.full(N, patch_cell * patch_cell, np.float32),
                 np.minimum(np.concatenate(refls), np.float32(0.99)).astype(np.float32),
                 np.concatenate(clusters), np.zeros(N, np.uint8), nC, pvs, None, None, lights,
                 {"seed": seed, "nx": nx, "ny": ny, "boxes_per_room": boxes_per_room})


def shadow_segments(scene: Scene, n: int, seed: int = 0xC0FFEE) -> tuple[np.ndarray, np.ndarray]:
    """C1 ray workload: n shadow segments as SoA blocks [3, n] (start, stop): even lanes run
    patch -> light, odd lanes patch -> patch, both from `origin + normal` like the transfer test."""
    rng = SplitMix64(seed)
    N = scene.n_patches
    a = rng.integers(n, N)
    b = rng.integers(n, N)
    start = (scene.patch_origin[a] + scene.patch_normal[a]).astype(np.float32)
    stop = (scene.patch_origin[b] + scene.patch_normal[b]).astype(np.float32)
    if scene.lights is not None and len(scene.lights):
        li = rng.integers(n, len(scene.lights))
        to_light = (np.arange(n) % 2) == 0
        stop[to_light] = scene.lights["origin"][li[to_light]]
    return np.ascontiguousarray(start.T), np.ascontiguousarray(stop.T)


def shadow_segment_indices(scene: Scene, n: int, seed: int = 0xC0FFEE) -> tuple[np.ndarray, np.ndarray]:
    """The segments of `shadow_segments(scene, n, seed)` as the lighting stages name them: a point table
    (`origin + normal` of every patch, then the light origins) and [n, 2] int32 index pairs (start, stop) into it.
    points[pairs[:, 0]].T == start and points[pairs[:, 1]].T == stop, bit for bit."""
    rng = SplitMix64(seed)
    N = scene.n_patches
    a = rng.integers(n, N)
    b = rng.integers(n, N)
    pts = (scene.patch_origin + scene.patch_normal).astype(np.float32)
    pairs = np.stack([a, b], axis=1).astype(np.int32)
    if scene.lights is not None and len(scene.lights):
        li = rng.integers(n, len(scene.lights))
        to_light = (np.arange(n) % 2) == 0
        pairs[to_light, 1] = N + li[to_light]
        pts = np.concatenate([pts, np.asarray(scene.lights["origin"], np.float32)], axis=0)
    return np.ascontiguousarray(pts), np.ascontiguousarray(pairs)


def random_rays(scene: Scene, n: int, seed: int = 0xBEEF) -> dict:
    """General closest-hit rays: origins inside the scene AABB, uniform directions, tmax = MAX_TRACE_LENGTH."""
    rng = SplitMix64(seed)
    v = scene.tri_verts.reshape(-1, 3)
    lo, hi = v.min(axis=0), v.max(axis=0)
    o = np.stack([rng.uniform(n, float(lo[c]) + 1.0, float(hi[c]) - 1.0) for c in range(3)])
    z = rng.uniform(n, -1.0, 1.0).astype(np.float64)
    phi = rng.uniform(n, 0.0, 2.0 * math.pi).astype(np.float64)
    r = np.sqrt(np.maximum(0.0, 1.0 - z * z))
    d = np.stack([r * np.cos(phi), r * np.sin(phi), z]).astype(np.float32)
    return {"o": np.ascontiguousarray(o), "d": np.ascontiguousarray(d),
            "tmax": np.full(n, MAX_TRACE_LENGTH, np.float32)}


def _value_noise(rng: SplitMix64, n: int, octaves=((64, 600.0), (16, 160.0), (4, 30.0))) -> np.ndarray:
    """Deterministic value noise on an (n+1) x (n+1) vertex grid: bilinear interpolation of random lattices."""
    h = np.zeros((n + 1, n + 1), np.float64)
    xs = np.arange(n + 1, dtype=np.float64)
    for cell, amp in octaves:
        m = n // cell + 2
        lat = rng.uniform(m * m).astype(np.float64).reshape(m, m)
        g = xs / cell
        i0 = np.floor(g).astype(np.int64); f = g - i0
        f = f * f * (3 - 2 * f)
        a = lat[i0][:, i0]; b = lat[i0 + 1][:, i0]; c = lat[i0][:, i0 + 1]; d = lat[i0 + 1][:, i0 + 1]
        fx, fy = f[:, None], f[None, :]
        h += amp * ((a * (1 - fx) + b * fx) * (1 - fy) + (c * (1 - fx) + d * fx) * fy)
    return h


def outdoor(seed: int = 0x5EED0003, cells: int = 708, cell: float = 46.0, n_buildings: int = 2000, patch_split: int = 2) -> Scene:
    """S3 (config C5): value-noise heightfield (cells x cells x 2 triangles), box 'buildings' standing on it and a
    TRACE_ID_SKY box around everything; one leaf patch per heightfield cell quarter; sun + sky-ambient lights."""
    rng = SplitMix64(seed)
    n = cells
    ext = n * cell
    h = _value_noise(rng, n)
    h -= h.min()
    ii, jj = np.meshgrid(np.arange(n), np.arange(n), indexing="ij")
    x0 = (ii * cell - ext / 2).ravel(); y0 = (jj * cell - ext / 2).ravel()
    z00 = h[ii, jj].ravel(); z10 = h[ii + 1, jj].ravel(); z01 = h[ii, jj + 1].ravel(); z11 = h[ii + 1, jj + 1].ravel()
    x1, y1 = x0 + cell, y0 + cell
    # two triangles per cell, same fan order as AddQuad: (v1,v2,v3), (v1,v3,v4)
    t1 = np.stack([x0, y0, z00, x1, y0, z10, x1, y1, z11], axis=1)
    t2 = np.stack([x0, y0, z00, x1, y1, z11, x0, y1, z01], axis=1)
    terrain = np.empty((2 * n * n, 9), np.float32)
    terrain[0::2] = t1; terrain[1::2] = t2
    g = _Geom()
    bi = rng.integers(n_buildings, n - 8) + 4; bj = rng.integers(n_buildings, n - 8) + 4
    bs = rng.uniform(3 * n_buildings, 0.0, 1.0).reshape(n_buildings, 3)
    for k in range(n_buildings):
        cx = bi[k] * cell - ext / 2; cy = bj[k] * cell - ext / 2
        base = float(h[bi[k]:bi[k] + 4, bj[k]:bj[k] + 4].min()) - 8.0
        w, d, hh = 60 + 120 * bs[k, 0], 60 + 120 * bs[k, 1], 120 + 500 * bs[k, 2]
        g.add_box(TRACE_ID_OPAQUE, (cx, cy, base), (cx + w, cy + d, base + hh))
    zmax = float(h.max()) + 1200.0
    g.add_box(TRACE_ID_SKY, (-ext / 2 - 64, -ext / 2 - 64, -64.0), (ext / 2 + 64, ext / 2 + 64, zmax))
    bids, bverts, _ = g.arrays()
    ids = np.concatenate([np.full(terrain.shape[0], TRACE_ID_OPAQUE, np.int32), bids])
    verts = np.concatenate([terrain, bverts])
    flags = np.zeros(ids.shape[0], np.uint8)
    # leaf patches: patch_split x patch_split per cell, on the cell's bilinear surface, lifted along the cell normal
    s = patch_split
    u = (np.arange(s) + 0.5) / s
    uu, vv = np.meshgrid(u, u, indexing="ij")
    uu = uu.ravel()[None, :]; vv = vv.ravel()[None, :]
    px = x0[:, None] + uu * cell; py = y0[:, None] + vv * cell
    pz = (z00[:, None] * (1 - uu) * (1 - vv) + z10[:, None] * uu * (1 - vv) + z01[:, None] * (1 - uu) * vv + z11[:, None] * uu * vv)
    nx = -(z10 - z00 + z11 - z01) / (2 * cell); ny = -(z01 - z00 + z11 - z10) / (2 * cell)
    nrm = np.stack([nx, ny, np.ones_like(nx)], axis=1); nrm /= np.linalg.norm(nrm, axis=1, keepdims=True)
    normal = np.repeat(nrm, s * s, axis=0).astype(np.float32)
    centres = np.stack([px.ravel(), py.ravel(), pz.ravel()], axis=1)
    origin = (centres + normal.astype(np.float64)).astype(np.float32)
    plane_dist = np.einsum("ij,ij->i", normal.astype(np.float64), centres).astype(np.float32)
    N = origin.shape[0]
    area = np.full(N, (cell / s) ** 2, np.float32)
    refl = np.minimum(rng.uniform(3 * N, 0.2, 0.6).reshape(N, 3), np.float32(0.99)).astype(np.float32)
    # clusters: 32x32-cell tiles; PVS = tiles within Chebyshev distance 1 (a far clip, like a real map's vis)
    tile = 32
    nt = (n + tile - 1) // tile
    cluster_cell = (ii // tile * nt + jj // tile).ravel().astype(np.int32)
    cluster = np.repeat(cluster_cell, s * s)
    pvs = np.zeros((nt * nt, nt * nt), np.uint8)
    for a in range(nt):
        for b in range(nt):
            for da in (-1, 0, 1):
                for db in (-1, 0, 1):
                    if 0 <= a + da < nt and 0 <= b + db < nt:
                        pvs[a * nt + b, (a + da) * nt + (b + db)] = 1
    lights = np.zeros(2, dtype=LIGHT_DTYPE)
    sun = np.array([-0.4, -0.3, -0.87]); sun /= np.linalg.norm(sun)
    lights["start_fade"], lights["end_fade"], lights["cap_dist"] = 0.0, -1.0, 1.0e22
    lights[0]["type"] = EMIT_SKYLIGHT; lights[0]["normal"] = sun.astype(np.float32); lights[0]["intensity"] = (300.0, 280.0, 250.0)
    lights[1]["type"] = EMIT_SKYAMBIENT; lights[1]["intensity"] = (40.0, 50.0, 70.0)
    return Scene("S3_outdoor", ids, verts, flags, origin, normal, plane_dist, area, refl, cluster, np.zeros(N, np.uint8),
                 nt * nt, pvs, origin, normal, lights, {"seed": seed, "cells": cells, "n_buildings": n_buildings})


@dataclass
class Bsp:
    """The lumps trace.PointLeafnum (raytracer/trace/pointleaf.go:8-33) and clustertable.PointInLeaf
    (rad/clustertable/point.go:14-38) walk, as flat arrays: Nodes{PlaneNum, Children[2]} (negative child =
    -1 - leaf), Planes{Normal, Distance, AxisType}, Leafs{Cluster, Area, Mins, Maxs}, len(Areas)."""
    node_plane: np.ndarray
    node_children: np.ndarray      # int32 [n_nodes, 2]
    plane_normal: np.ndarray       # float32 [n_planes, 3]
    plane_dist: np.ndarray
    plane_type: np.ndarray         # 0,1,2 = axial x,y,z; >= 3 = general (dot-product path)
    leaf_cluster: np.ndarray
    leaf_area: np.ndarray
    leaf_mins: np.ndarray          # int16 [n_leafs, 3]
    leaf_maxs: np.ndarray
    n_areas: int


class _BspBuilder:
    def __init__(self):
        self.nodes, self.planes, self.leafs = [], [], []

    def plane(self, normal, dist, ptype):
        self.planes.append((tuple(float(v) for v in normal), float(dist), int(ptype)))
        return len(self.planes) - 1

    def leaf(self, cluster, area, mins=(0, 0, 0), maxs=(0, 0, 0)):
        self.leafs.append((int(cluster), int(area), tuple(mins), tuple(maxs)))
        return -1 - (len(self.leafs) - 1)

    def node(self, plane):
        """Reserve a node (parents before children, as in a BSP lump); fill with set_children."""
        self.nodes.append([plane, 0, 0])
        return len(self.nodes) - 1

    def set_children(self, node, front, back):
        self.nodes[node][1], self.nodes[node][2] = front, back

    def finish(self, n_areas) -> Bsp:
        nd = np.asarray(self.nodes, np.int32).reshape(-1, 3)
        return Bsp(nd[:, 0].copy(), nd[:, 1:3].copy(),
                   np.asarray([p[0] for p in self.planes], np.float32).reshape(-1, 3),
                   np.asarray([p[1] for p in self.planes], np.float32), np.asarray([p[2] for p in self.planes], np.int32),
                   np.asarray([l[0] for l in self.leafs], np.int32), np.asarray([l[1] for l in self.leafs], np.int32),
                   np.asarray([l[2] for l in self.leafs], np.int16).reshape(-1, 3),
                   np.asarray([l[3] for l in self.leafs], np.int16).reshape(-1, 3), n_areas)


def sky_room(seed: int = 0x5EED0004, n_boxes: int = 40, n_panes: int = 24) -> Scene:
    """Scene for the complete TestLineDoesHitSky surface: the S1 room with a
    TRACE_ID_SKY ceiling, a static prop (TRACE_ID_STATICPROP | 7), transparent panes with per-triangle
    coverage colours, a closed opaque bunker, and two 3D sky boxes far outside the room, each with a
    sky_camera (scales 16 and 32; a third camera entity has scale 0 and is dropped).  meta carries the BSP
    lumps (`bsp`), the camera entities and the triangle colours."""
    rng = SplitMix64(seed)
    sx, sy, sz = 1024.0, 1024.0, 512.0
    x0, y0, z0 = -sx / 2, -sy / 2, 0.0
    g = _Geom()
    faces = _room_faces(x0, y0, z0, sx, sy, sz)
    for k, f in enumerate(faces):
        _face_quad(g, TRACE_ID_SKY if k == 1 else TRACE_ID_OPAQUE, f)          # loadbsp/main.go:255 sky sides
    mins, maxs = _place_boxes(rng, n_boxes, x0, y0, x0 + sx, y0 + sy, z0)
    g.add_box_array(TRACE_ID_OPAQUE, mins, maxs)
    first = len(g.ids)
    g.add_box(TRACE_ID_STATICPROP | 7, (-60.0, -60.0, 300.0), (60.0, 60.0, 340.0))    # a floating static prop:
    g.ids[first:] = [TRACE_ID_STATICPROP | 7] * (len(g.ids) - first)                   # every triangle carries the prop id
    g.add_box(TRACE_ID_OPAQUE, (300.0, 300.0, 100.0), (428.0, 428.0, 228.0))          # bunker (hollow, closed)
    n_opaque = len(g.ids)
    # transparent panes: horizontal quads between z=360 and z=480, stacked so a ray can cross several
    colors = []
    for k in range(n_panes):
        p = rng.uniform(5)
        cx, cy = -400.0 + 800.0 * float(p[0]), -400.0 + 800.0 * float(p[1])
        w, z = 80.0 + 160.0 * float(p[2]), 360.0 + 120.0 * float(p[3])
        g.add_quad(TRACE_ID_OPAQUE, (cx - w, cy - w, z), (cx + w, cy - w, z), (cx + w, cy + w, z), (cx - w, cy + w, z))
        cov = (0.25, 0.5, 0.75, 1.0)[k % 4] if k % 5 else float(np.float32(p[4]))
        colors += [cov, cov]
    n_world = len(g.ids)
    # 3D sky boxes: sky faces outside, opaque "mountains" inside
    sky = [((8192.0, 0.0, 0.0), 16.0), ((8192.0, 4096.0, 0.0), 32.0)]
    for (c, scale) in sky:
        hx, hz = 256.0, 160.0
        g.add_box(TRACE_ID_SKY, (c[0] - hx, c[1] - hx, c[2] - 16.0), (c[0] + hx, c[1] + hx, c[2] + hz))
        for m in range(6):
            q = rng.uniform(4)
            mx_, my_ = c[0] - 200.0 + 400.0 * float(q[0]), c[1] - 200.0 + 400.0 * float(q[1])
            if abs(mx_ - c[0]) < 48.0 and abs(my_ - c[1]) < 48.0:
                mx_ += 96.0
            s_ = 20.0 + 40.0 * float(q[2]); h_ = 60.0 + 90.0 * float(q[3])
            g.add_box(TRACE_ID_OPAQUE, (mx_ - s_, my_ - s_, c[2] - 16.0), (mx_ + s_, my_ + s_, c[2] + h_))
    ids, verts, flags = g.arrays()
    flags[n_opaque:n_world] = 1                                                # transparent panes
    tri_colors = np.zeros((ids.shape[0], 3), np.float32)
    tri_colors[:, :] = 1.0
    tri_colors[n_opaque:n_world, 0] = np.asarray(colors, np.float32)

    # BSP: root splits world | sky boxes; world side: solid below the floor (cluster -1), then x/y quadrants
    # with one general (non-axial) plane; sky side: one leaf per sky box.  Areas: 0 solid, 1 world, 2/3 sky boxes.
    b = _BspBuilder()
    n_root = b.node(b.plane((1, 0, 0), 4096.0, 0))
    n_sky = b.node(b.plane((0, 1, 0), 2048.0, 1))
    n_floor = b.node(b.plane((0, 0, 1), 0.0, 2))
    b.set_children(n_root, n_sky, n_floor)
    b.set_children(n_sky, b.leaf(5, 3, (7936, 3840, -16), (8448, 4352, 160)), b.leaf(4, 2, (7936, -256, -16), (8448, 256, 160)))
    n_x = b.node(b.plane((1, 0, 0), 0.0, 0))
    b.set_children(n_floor, n_x, b.leaf(-1, 0, (-512, -512, -64), (512, 512, 0)))
    n_diag = b.node(b.plane((0.6, 0.8, 0.0), 100.0, 3))                        # general plane: dot-product path
    n_y = b.node(b.plane((0, 1, 0), 0.0, 1))
    b.set_children(n_x, n_diag, n_y)
    b.set_children(n_diag, b.leaf(0, 1, (0, -512, 0), (512, 512, 512)), b.leaf(1, 1, (0, -512, 0), (300, 200, 512)))
    b.set_children(n_y, b.leaf(2, 1, (-512, 0, 0), (0, 512, 512)), b.leaf(3, 1, (-512, -512, 0), (0, 0, 512)))
    bsp = b.finish(4)

    cams_origin = np.asarray([[8192.0, 0.0, 24.0], [0.0, 0.0, 100.0], [8192.0, 4096.0, 24.0]], np.float32)
    cams_scale = np.asarray([16.0, 0.0, 32.0], np.float32)                     # the second entity is dropped (scale <= 0)
    # "leafs" for CanLeafTraceToSky: BSP leaf bounds + probe boxes (inside the bunker, under a wide pane, ...)
    probe_mins = np.asarray([[332, 332, 132], [-400, -400, 10], [8100, -60, 0], [-64, -64, 200]], np.int16)
    probe_maxs = np.asarray([[396, 396, 196], [-300, -300, 90], [8160, 60, 60], [64, 64, 280]], np.int16)
    meta = {"seed": seed, "bsp": bsp, "cams_origin": cams_origin, "cams_scale": cams_scale, "tri_colors": tri_colors,
            "n_opaque": n_opaque, "n_world": n_world,
            "probe_mins": np.concatenate([bsp.leaf_mins, probe_mins]), "probe_maxs": np.concatenate([bsp.leaf_maxs, probe_maxs])}
    return Scene("S4_sky_room", ids, verts, flags, meta=meta)


def sky_segments(scene: Scene, n: int, seed: int = 0x5C1) -> tuple[np.ndarray, np.ndarray]:
    """Segments for the sky test: starts inside the room (some inside a sky box, some below the floor), stops far
    along mostly-upward directions; every 97th segment has zero length."""
    rng = SplitMix64(seed)
    a = np.stack([rng.uniform(n, -500, 500), rng.uniform(n, -500, 500), rng.uniform(n, 4, 500)])
    z = rng.uniform(n, 0.05, 1.0).astype(np.float64)
    side = rng.integers(n, 8)
    z[side == 0] *= -1.0                                                      # some point down / sideways
    phi = rng.uniform(n, 0.0, 2.0 * math.pi).astype(np.float64)
    r = np.sqrt(np.maximum(0.0, 1.0 - z * z))
    d = np.stack([r * np.cos(phi), r * np.sin(phi), z])
    ln = rng.uniform(n, 50.0, 30000.0).astype(np.float64)
    insky = side == 1                                                         # start inside sky box 0 (its area has a camera)
    a[0, insky] = 8192.0 + a[0, insky] * 0.4; a[1, insky] *= 0.4; a[2, insky] = 2.0 + a[2, insky] * 0.25
    below = side == 2
    a[2, below] = -10.0                                                       # solid leaf (area 0 has no camera either)
    b = (a.astype(np.float64) + d * ln).astype(np.float32)
    zero = (np.arange(n) % 97) == 5
    b[:, zero] = a[:, zero]
    return np.ascontiguousarray(a.astype(np.float32)), np.ascontiguousarray(b)


def mini_sky_scene():
    """Hand-made scene for known answers: floor, TRACE_ID_SKY ceiling at z=512, a static prop (id 7) at z=300 over
    the origin, and three transparent panes over x in [100,200] (the third only over [150,200]) with coverages
    0.25, 0.5, 0.5 at z = 400, 420, 440."""
    ids, verts, flags, cols = [], [], [], []

    def quad(tid, x0, x1, y0, y1, z, flag=0, cov=1.0):
        for tri in (((x0, y0, z), (x1, y0, z), (x1, y1, z)), ((x0, y0, z), (x1, y1, z), (x0, y1, z))):
            ids.append(tid); verts.append([c for v in tri for c in v]); flags.append(flag); cols.append([cov, 0.0, 0.0])
    quad(TRACE_ID_OPAQUE, -512, 512, -512, 512, 0)
    quad(TRACE_ID_SKY, -512, 512, -512, 512, 512)
    quad(TRACE_ID_STATICPROP | 7, -60, 60, -60, 60, 300)
    quad(TRACE_ID_OPAQUE, 100, 200, -50, 50, 400, 1, 0.25)
    quad(TRACE_ID_OPAQUE, 100, 200, -50, 50, 420, 1, 0.5)
    quad(TRACE_ID_OPAQUE, 150, 200, -50, 50, 440, 1, 0.5)
    return (np.asarray(ids, np.int32), np.asarray(verts, np.float32), np.asarray(flags, np.uint8), np.asarray(cols, np.float32))


MINI_SKY_CASES = [
    # start, stop, flags, prop to skip, expected fractionVisible
    ((0, 0, 100), (0, 0, 5000), 0, -1, 0.0),        # the prop blocks
    ((0, 0, 100), (0, 0, 5000), 0, 7, 1.0),         # prop 7 skipped (testline.go:36); the ceiling is sky (:46-48)
    ((0, 0, 100), (0, 0, 5000), 0, 8, 0.0),         # another prop id does not help
    ((0, 0, 100), (0, 0, 250), 0, -1, 1.0),         # ends below the prop
    ((300, 0, 100), (300, 0, -50), 0, -1, 0.0),     # the floor is an ordinary blocker
    ((120, 0, 100), (120, 0, 5000), 0, -1, 0.0),    # no callback: a transparent pane blocks like any triangle
    ((120, 0, 100), (120, 0, 5000), 2, -1, 0.25),   # coverage 0.25 + 0.5 (coverageCount.go:29-30)
    ((120, 0, 100), (120, 0, 410), 2, -1, 0.75),    # only the first pane lies before the segment end
    ((170, 0, 100), (170, 0, 5000), 2, -1, 0.0),    # 0.25 + 0.5 + 0.5 clamps to 1: fully covered
    ((170, 0, 100), (170, 0, 430), 2, -1, 0.25),
    ((170, 0, 450), (170, 0, 100), 2, -1, 0.0),     # downwards through all three, then the floor is beyond the end
    ((50, 50, 50), (50, 50, 50), 3, -1, 1.0),       # zero-length segment: visible
]


# ---- hierarchical patches (rad/patches/face.go + subdivide.go) ------------------------------------

FACE_PATCH_DTYPE = np.dtype([("first_point", "<i4"), ("n_points", "<i4"), ("normal", "<f4", 3), ("plane_dist", "<f4"),
                             ("lux_scale", "<f4"), ("chop", "<f4"), ("sky", "u1"), ("no_subdivide", "u1"),
                             ("has_base_light", "u1"), ("pad", "u1")])     # == vrad_face_patch (include/vrad_cuda.h)


class _FaceList:
    """Face windings the way the BSP hands them to MakePatchForFace (rad/patches/face.go:29): planar convex polygons."""

    def __init__(self, lux_scale=1.0 / 16.0, chop=4.0):
        self.points, self.faces, self.lux, self.chop = [], [], lux_scale, chop

    def quad(self, o, u, v, normal, sky=0, no_subdivide=0):
        o = np.asarray(o, np.float64); u = np.asarray(u, np.float64); v = np.asarray(v, np.float64)
        first = len(self.points)
        for p in (o, o + u, o + u + v, o + v):
            self.points.append(tuple(np.float32(c) for c in p))
        n = np.asarray(normal, np.float64)
        self.faces.append((first, 4, tuple(np.float32(c) for c in n), np.float32(float(np.dot(n, o))), np.float32(self.lux),
                           np.float32(self.chop), sky, no_subdivide, 0, 0))

    def arrays(self):
        return np.asarray(self.faces, dtype=FACE_PATCH_DTYPE), np.asarray(self.points, np.float32).reshape(-1, 3)


def room_faces(nx=3, ny=2, room=512.0, door_w=128.0, door_h=256.0):
    """The faces of the multi_room geometry (same rooms, shared walls with door openings), one winding each;
    returns (faces, points, face_room) -- face_room = the room (cluster) the face belongs to."""
    R = room
    fl = _FaceList()
    face_room = []
    a, b = (R - door_w) / 2, (R + door_w) / 2
    for i in range(nx):
        for j in range(ny):
            x0, y0, k = i * R, j * R, i * ny + j
            fl.quad((x0, y0, 0), (R, 0, 0), (0, R, 0), (0, 0, 1)); face_room.append(k)            # floor
            fl.quad((x0, y0, R), (R, 0, 0), (0, R, 0), (0, 0, -1)); face_room.append(k)           # ceiling
            walls = [((x0, y0, 0), (0, 1, 0), (1, 0, 0), i > 0), ((x0 + R, y0, 0), (0, 1, 0), (-1, 0, 0), i < nx - 1),
                     ((x0, y0, 0), (1, 0, 0), (0, 1, 0), j > 0), ((x0, y0 + R, 0), (1, 0, 0), (0, -1, 0), j < ny - 1)]
            for (o, u, nrm, door) in walls:
                o = np.asarray(o, np.float64); u = np.asarray(u, np.float64)
                if not door:
                    fl.quad(o, u * R, (0, 0, R), nrm); face_room.append(k)
                else:                                                                              # left, right, above the door
                    fl.quad(o, u * a, (0, 0, R), nrm); face_room.append(k)
                    fl.quad(o + u * b, u * (R - b), (0, 0, R), nrm); face_room.append(k)
                    fl.quad(o + u * a + np.array([0, 0, door_h]), u * (b - a), (0, 0, R - door_h), nrm); face_room.append(k)
    faces, points = fl.arrays()
    return faces, points, np.asarray(face_room, np.int32)


def multi_room_hier(seed: int = 0x5EED0002, nx: int = 3, ny: int = 2, boxes_per_room: int = 30, pvs_radius: int = 2) -> Scene:
    """The multi_room map with its patches made the reference's way: one root patch per face, subdivided by
    patches.SubdividePatches (product host code, vrad_patches_subdivide) into Parent/Child1/Child2 trees with
    4-luxel (64-unit) chop.  Patch arrays hold ALL patches (roots, interior, leaves); meta["tree"] has the links."""
    from .environment import subdivide_patches
    base = multi_room(seed, nx, ny, boxes_per_room=boxes_per_room, pvs_radius=pvs_radius)
    faces, points, face_room = room_faces(nx, ny)
    t = subdivide_patches(faces, points, min_chop=4.0)
    N = t["origin"].shape[0]
    rng = SplitMix64(seed ^ 0xABCDEF)
    refl_face = np.minimum(rng.uniform(3 * len(faces), 0.2, 0.7).reshape(-1, 3), np.float32(0.99))
    sc = Scene(base.name + "_hier", base.tri_ids, base.tri_verts, base.tri_flags, t["origin"], t["normal"], t["plane_dist"], t["area"],
               refl_face[t["face"]].astype(np.float32), face_room[t["face"]].astype(np.int32), np.zeros(N, np.uint8),
               base.n_clusters, base.pvs, None, None, base.lights, dict(base.meta))
    sc.meta["tree"] = t
    sc.meta["faces"] = faces
    sc.meta["face_points"] = points
    return sc


def room_grid_bsp(nx=3, ny=2, room=512.0) -> Bsp:
    """BSP lumps for the multi_room grid: axial splits at the room boundaries, one leaf (cluster = i*ny + j, area 1) per room."""
    b = _BspBuilder()

    def build(i0, i1, j0, j1):
        if i1 - i0 == 1 and j1 - j0 == 1:
            return b.leaf(i0 * ny + j0, 1, (int(i0 * room), int(j0 * room), 0), (int(i1 * room), int(j1 * room), int(room)))
        if i1 - i0 >= j1 - j0:
            m = (i0 + i1) // 2
            n = b.node(b.plane((1, 0, 0), m * room, 0))
            front, back = build(m, i1, j0, j1), build(i0, m, j0, j1)
        else:
            m = (j0 + j1) // 2
            n = b.node(b.plane((0, 1, 0), m * room, 1))
            front, back = build(i0, i1, m, j1), build(i0, i1, j0, m)
        b.set_children(n, front, back)
        return n
    build(0, nx, 0, ny)
    return b.finish(2)


def compress_vis_rows(pvs) -> tuple[bytes, np.ndarray]:
    """The visibility lump a BSP compiler would write for a [C, C] 0/1 matrix: each row bit-packed (cluster k = bit k&7 of
    byte k>>3) and run-length coded (zero runs -> 0, count); returns (lump bytes, ByteOffset[C][2] with PAS = -1)."""
    pvs = np.asarray(pvs)
    lump = bytearray()
    ofs = np.full((pvs.shape[0], 2), -1, np.int32)
    for c in range(pvs.shape[0]):
        row = np.packbits(pvs[c].astype(bool), bitorder="little").tobytes()
        ofs[c, 0] = len(lump)
        j = 0
        while j < len(row):
            lump.append(row[j])
            if row[j]:
                j += 1
                continue
            rep = 1
            j += 1
            while j < len(row) and row[j] == 0 and rep < 255:
                rep += 1; j += 1
            lump.append(rep)
    return bytes(lump), ofs
