"""ctypes mirror of include/vrad_bsp.h (the BSP side of the hot path: .bsp container, lumps -> kernel inputs, lighting
lump write-back) and a synthetic BSP v20 generator for the multi-room scenes, so that the whole path can be driven from a
file the reference's loader (github.com/galaco/bsp via cache.BuildLumpCache, cache/bsp.go:51-91) can also read.

Everything here calls the C-ABI in libvradcuda.so; all of it except `lightmap_finalize` runs without a GPU.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import lib as _lib
from .lib import FACE_PATCH_DTYPE

# lump records, == the structs of include/vrad_bsp.h
DPLANE = np.dtype([("normal", "<f4", 3), ("dist", "<f4"), ("type", "<i4")])
DEDGE = np.dtype([("v", "<u2", 2)])
DFACE = np.dtype([("planenum", "<u2"), ("side", "u1"), ("on_node", "u1"), ("firstedge", "<i4"), ("numedges", "<i2"),
                  ("texinfo", "<i2"), ("dispinfo", "<i2"), ("fog_volume", "<i2"), ("styles", "u1", 4), ("lightofs", "<i4"),
                  ("area", "<f4"), ("lm_mins", "<i4", 2), ("lm_size", "<i4", 2), ("orig_face", "<i4"), ("num_prims", "<u2"),
                  ("first_prim", "<u2"), ("smoothing_groups", "<u4")])
TEXINFO = np.dtype([("texture_vecs", "<f4", (2, 4)), ("lightmap_vecs", "<f4", (2, 4)), ("flags", "<i4"), ("texdata", "<i4")])
DTEXDATA = np.dtype([("reflectivity", "<f4", 3), ("name_id", "<i4"), ("width", "<i4"), ("height", "<i4"),
                     ("view_width", "<i4"), ("view_height", "<i4")])
DMODEL = np.dtype([("mins", "<f4", 3), ("maxs", "<f4", 3), ("origin", "<f4", 3), ("headnode", "<i4"), ("firstface", "<i4"),
                   ("numfaces", "<i4")])
DNODE = np.dtype([("planenum", "<i4"), ("children", "<i4", 2), ("mins", "<i2", 3), ("maxs", "<i2", 3), ("firstface", "<u2"),
                  ("numfaces", "<u2"), ("area", "<i2"), ("pad", "<i2")])
DLEAF = np.dtype([("contents", "<i4"), ("cluster", "<i2"), ("area_flags", "<i2"), ("mins", "<i2", 3), ("maxs", "<i2", 3),
                  ("firstleafface", "<u2"), ("numleaffaces", "<u2"), ("firstleafbrush", "<u2"), ("numleafbrushes", "<u2"),
                  ("leaf_water_data", "<i2"), ("pad", "<i2")])
DBRUSH = np.dtype([("firstside", "<i4"), ("numsides", "<i4"), ("contents", "<i4")])
DBRUSHSIDE = np.dtype([("planenum", "<u2"), ("texinfo", "<i2"), ("dispinfo", "<i2"), ("bevel", "<i2")])
RGBEXP32 = np.dtype([("r", "u1"), ("g", "u1"), ("b", "u1"), ("exponent", "i1")])
for _dt, _n in ((DPLANE, 20), (DEDGE, 4), (DFACE, 56), (TEXINFO, 72), (DTEXDATA, 32), (DMODEL, 48), (DNODE, 32), (DLEAF, 32),
                (DBRUSH, 12), (DBRUSHSIDE, 8), (RGBEXP32, 4)):
    assert _dt.itemsize == _n, (_dt, _n)

LUMP = dict(ENTITIES=0, PLANES=1, TEXDATA=2, VERTEXES=3, VISIBILITY=4, NODES=5, TEXINFO=6, FACES=7, LIGHTING=8, LEAFS=10,
            EDGES=12, SURFEDGES=13, MODELS=14, LEAFFACES=16, LEAFBRUSHES=17, BRUSHES=18, BRUSHSIDES=19, AREAS=20,
            AREAPORTALS=21, VERTNORMALS=30, VERTNORMALINDICES=31, TEXDATA_STRING_DATA=43, TEXDATA_STRING_TABLE=44,
            LIGHTING_HDR=53, FACES_HDR=58, MAP_FLAGS=59)
SURF_LIGHT, SURF_SKY2D, SURF_SKY, SURF_NOLIGHT, SURF_BUMPLIGHT, SURF_NOCHOP = 0x1, 0x2, 0x4, 0x400, 0x800, 0x4000
CONTENTS_SOLID, CONTENTS_WINDOW = 0x1, 0x2
LEAF_FLAGS_SKY, LEAF_FLAGS_RADIAL, LEAF_FLAGS_SKY2D = 0x1, 0x2, 0x4
TRACE_ID_SKY, TRACE_ID_OPAQUE = 0x01000000, 0x02000000

# every symbol include/vrad_bsp.h declares (tests check the library exports all of them)
SYMBOLS = [
    "vrad_bspfile_create", "vrad_bspfile_open", "vrad_bspfile_get_lump", "vrad_bspfile_set_lump", "vrad_bspfile_save",
    "vrad_bspfile_close", "vrad_bspfile_lumps", "vrad_bsp_raytrace_triangles", "vrad_env_add_bsp", "vrad_bsp_face_patches",
    "vrad_bsp_rescale_lightmap_vecs", "vrad_bsp_face_extents", "vrad_bsp_make_parents", "vrad_bsp_cluster_table",
    "vrad_bsp_vis_for_light_environment", "vrad_bsp_pair_edges", "vrad_bsp_save_vertex_normals", "vrad_bsp_phong_normals",
    "vrad_bsp_layout_lighting", "vrad_bsp_face_luxels", "vrad_color_to_rgbexp32", "vrad_color_from_rgbexp32",
    "vrad_lightmap_finalize", "vrad_bsp_pack_lighting", "vrad_luxel_nearest_patch", "vrad_lightmap_finalize_patches",
    "vrad_texlights_parse", "vrad_bsp_apply_texlights", "vrad_bspfile_set_target_faces", "vrad_bsp_validate",
    "vrad_bsp_radial_entries", "vrad_luxel_radial_light", "vrad_luxel_radial_light_host", "vrad_bsp_place_samples",
]


class _LumpsStruct(C.Structure):
    _fields_ = [f for name in ("planes", "vertexes3", "edges", "surfedges", "faces", "texinfo", "texdata", "models", "nodes", "leafs",
                               "leaffaces", "leafbrushes", "brushes", "brushsides")
                for f in (("n_" + name.rstrip("3"), C.c_int32), (name, C.c_void_p))] + \
               [("n_areas", C.c_int32), ("pad0", C.c_int32), ("vis_len", C.c_int64), ("visdata", C.c_void_p)]


assert C.sizeof(_LumpsStruct) == 14 * 16 + 8 + 16


def _check(rc, what):
    if rc != 0:
        msg = _lib.load().vrad_last_error()
        raise _lib.VradError(rc, f"{what}: {msg.decode(errors='replace') if msg else rc}")


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


_LUMP_FIELDS = (("planes", DPLANE), ("vertexes3", np.dtype("<f4")), ("edges", DEDGE), ("surfedges", np.dtype("<i4")), ("faces", DFACE),
                ("texinfo", TEXINFO), ("texdata", DTEXDATA), ("models", DMODEL), ("nodes", DNODE), ("leafs", DLEAF),
                ("leaffaces", np.dtype("<u2")), ("leafbrushes", np.dtype("<u2")), ("brushes", DBRUSH), ("brushsides", DBRUSHSIDE))


class Lumps:
    """The typed lump views the input functions take (vrad_bsp_lumps), held as numpy arrays."""

    def __init__(self, **arrays):
        self.a = {}
        for name, dt in _LUMP_FIELDS:
            v = arrays.get(name)
            v = np.zeros(0, dt) if v is None else np.ascontiguousarray(v, dtype=dt)
            self.a[name] = v.reshape(-1, 3) if name == "vertexes3" else v.reshape(-1)
        self.n_areas = int(arrays.get("n_areas", 0))
        vis = arrays.get("visdata", b"")
        self.visdata = np.frombuffer(bytes(vis), np.uint8).copy()
        self._struct()

    def _struct(self):
        s = _LumpsStruct()
        for name, _ in _LUMP_FIELDS:
            arr = self.a[name]
            setattr(s, "n_" + name.rstrip("3"), arr.shape[0])
            setattr(s, name, arr.ctypes.data if arr.size else None)
        s.n_areas = self.n_areas
        s.vis_len = self.visdata.size
        s.visdata = self.visdata.ctypes.data if self.visdata.size else None
        self.s = s
        return s

    def __getattr__(self, name):
        a = self.__dict__.get("a", {})
        if name in a:
            return a[name]
        raise AttributeError(name)

    def replace(self, **arrays):
        cur = dict(self.a); cur["n_areas"] = self.n_areas; cur["visdata"] = self.visdata.tobytes()
        cur.update(arrays)
        return Lumps(**cur)

    @property
    def n_clusters(self):
        return int(np.frombuffer(self.visdata[:4].tobytes(), "<i4")[0]) if self.visdata.size >= 4 else 0

    @property
    def ref(self):
        return C.byref(self.s)

    def validate(self):
        """vrad_bsp_validate: raises VradError when an index stored in one lump points outside another."""
        _check(_lib.load().vrad_bsp_validate(self.ref), "vrad_bsp_validate")


class BspFile:
    """vrad_bspfile: a .bsp as 64 opaque lumps (loadBSP, cmd/tasks/loadbsp/main.go:163-170; the writer of finish/main.go:15-18)."""

    def __init__(self, path: str | None = None, map_revision: int = 1):
        self._l = _lib.load()
        self._h = C.c_void_p()
        if path is None:
            _check(self._l.vrad_bspfile_create(C.c_int(map_revision), C.byref(self._h)), "vrad_bspfile_create")
        else:
            _check(self._l.vrad_bspfile_open(path.encode(), C.byref(self._h)), "vrad_bspfile_open")

    def close(self):
        if self._h:
            self._l.vrad_bspfile_close.restype = None
            self._l.vrad_bspfile_close(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def get(self, lump: int) -> tuple[bytes, int]:
        data, n, ver = C.c_void_p(), C.c_int64(), C.c_int()
        _check(self._l.vrad_bspfile_get_lump(self._h, C.c_int(lump), C.byref(data), C.byref(n), C.byref(ver)), "vrad_bspfile_get_lump")
        return (C.string_at(data, n.value) if n.value else b""), ver.value

    def set(self, lump: int, data, version: int = 0):
        b = data.tobytes() if isinstance(data, np.ndarray) else bytes(data)
        _check(self._l.vrad_bspfile_set_lump(self._h, C.c_int(lump), b, C.c_int64(len(b)), C.c_int(version)), "vrad_bspfile_set_lump")

    def save(self, path: str):
        _check(self._l.vrad_bspfile_save(self._h, path.encode()), "vrad_bspfile_save")

    def set_target_faces(self, hdr: bool):
        """cache.SetTargetFaces: returns (face lump, lighting lump) the job reads / writes."""
        a, b = C.c_int(), C.c_int()
        _check(self._l.vrad_bspfile_set_target_faces(self._h, C.c_int(int(hdr)), C.byref(a), C.byref(b)), "vrad_bspfile_set_target_faces")
        return a.value, b.value

    def lumps(self) -> Lumps:
        """Typed views (validated by the library), copied into numpy arrays."""
        s = _LumpsStruct()
        _check(self._l.vrad_bspfile_lumps(self._h, C.byref(s)), "vrad_bspfile_lumps")
        arrays = {}
        for name, dt in _LUMP_FIELDS:
            n = getattr(s, "n_" + name.rstrip("3"))
            count = n * (3 if name == "vertexes3" else 1)
            p = getattr(s, name)
            arrays[name] = np.frombuffer(C.string_at(p, count * dt.itemsize), dt).copy() if n else np.zeros(0, dt)
        arrays["n_areas"] = s.n_areas
        arrays["visdata"] = C.string_at(s.visdata, s.vis_len) if s.vis_len else b""
        return Lumps(**arrays)

    def set_lumps(self, L: Lumps, entities: str = ""):
        for name, lump in (("planes", 1), ("texdata", 2), ("vertexes3", 3), ("nodes", 5), ("texinfo", 6), ("faces", 7), ("edges", 12),
                           ("surfedges", 13), ("models", 14), ("leaffaces", 16), ("leafbrushes", 17), ("brushes", 18), ("brushsides", 19)):
            self.set(lump, L.a[name], version=1 if name == "faces" else 0)      # v20 files carry face lump version 1
        self.set(LUMP["LEAFS"], L.a["leafs"], version=1)
        self.set(LUMP["VISIBILITY"], L.visdata)
        self.set(LUMP["AREAS"], np.zeros(2 * L.n_areas, "<i4"))
        if entities:
            self.set(LUMP["ENTITIES"], entities.encode() + b"\0")


# ---- lumps -> kernel inputs ------------------------------------------------------------------------------------------

def raytrace_triangles(L: Lumps, caster_model=None, caster_origin=None, caster_angles=None):
    """addBrushesForRayTrace (+ shadow-casting brush entities): returns (ids int32 [n], verts float32 [n,3,3])."""
    l = _lib.load()
    nc = 0 if caster_model is None else len(caster_model)
    cm = np.ascontiguousarray(caster_model if nc else [], np.int32)
    co = np.ascontiguousarray(caster_origin if nc else [], np.float32).reshape(-1, 3)
    ca = np.ascontiguousarray(caster_angles if nc else [], np.float32).reshape(-1, 3)
    n = C.c_int()
    _check(l.vrad_bsp_raytrace_triangles(L.ref, C.c_int(nc), _ptr(cm), _ptr(co), _ptr(ca), C.c_int(0), None, None, C.byref(n)), "vrad_bsp_raytrace_triangles")
    ids = np.zeros(n.value, np.int32); verts = np.zeros((n.value, 3, 3), np.float32)
    _check(l.vrad_bsp_raytrace_triangles(L.ref, C.c_int(nc), _ptr(cm), _ptr(co), _ptr(ca), C.c_int(n.value), _ptr(ids), _ptr(verts), C.byref(n)), "vrad_bsp_raytrace_triangles")
    return ids, verts


def face_patches(L: Lumps, model_origins=None, max_chop: float = 4.0) -> dict:
    """MakePatches: one vrad_face_patch per non-displacement face (+ winding points and the per-face material fields)."""
    l = _lib.load()
    mo = None if model_origins is None else np.ascontiguousarray(model_origins, np.float32).reshape(-1, 3)
    nf, npnt = C.c_int(), C.c_int()
    args = (L.ref, _ptr(mo), C.c_float(max_chop))
    _check(l.vrad_bsp_face_patches(*args, C.c_int(0), C.c_int(0), C.byref(nf), C.byref(npnt), None, None, None, None, None, None, None), "vrad_bsp_face_patches")
    out = dict(faces=np.zeros(nf.value, FACE_PATCH_DTYPE), points=np.zeros((npnt.value, 3), np.float32), face_number=np.zeros(nf.value, np.int32),
               reflectivity=np.zeros((nf.value, 3), np.float32), base_area=np.zeros(nf.value, np.float32), needs_bump=np.zeros(nf.value, np.uint8),
               scale=np.zeros((nf.value, 2), np.float32))
    _check(l.vrad_bsp_face_patches(*args, C.c_int(nf.value), C.c_int(npnt.value), C.byref(nf), C.byref(npnt), _ptr(out["faces"]), _ptr(out["points"]),
                                   _ptr(out["face_number"]), _ptr(out["reflectivity"]), _ptr(out["base_area"]), _ptr(out["needs_bump"]), _ptr(out["scale"])),
           "vrad_bsp_face_patches")
    return out


TEXLIGHT = np.dtype([("name", "S116"), ("value", "<f4", 3)])
assert TEXLIGHT.itemsize == 128


def texlights_parse(text: str, hdr: bool = False):
    """lights.rad text -> (texlight table, noshadow materials, forcetextureshadow models)."""
    raw = text.encode()
    l = _lib.load()
    n, nn, nf, used = C.c_int(), C.c_int(), C.c_int(), C.c_int64()
    _check(l.vrad_texlights_parse(raw, C.c_int64(len(raw)), C.c_int(int(hdr)), C.c_int(0), None, C.byref(n), None, C.c_int64(0), C.byref(nn), C.byref(nf), C.byref(used)),
           "vrad_texlights_parse")
    table = np.zeros(max(n.value, 1), TEXLIGHT)
    names = C.create_string_buffer(max(used.value, 1))
    _check(l.vrad_texlights_parse(raw, C.c_int64(len(raw)), C.c_int(int(hdr)), C.c_int(table.shape[0]), _ptr(table), C.byref(n), names, C.c_int64(used.value),
                                  C.byref(nn), C.byref(nf), C.byref(used)), "vrad_texlights_parse")
    parts = names.raw[:used.value].split(b"\0")[:-1] if used.value else []
    parts = [p.decode() for p in parts]
    return table[:n.value].copy(), parts[:nn.value], parts[nn.value:nn.value + nf.value]


def apply_texlights(L: Lumps, string_table, string_data: bytes, map_name: str, texlights, face_number, faces=None):
    """BaseLightForFace for the faces of face_patches(): (base_light [n, 3], faces with has_base_light / no_subdivide updated)."""
    st = np.ascontiguousarray(string_table, np.int32); tl = np.ascontiguousarray(texlights, TEXLIGHT)
    fnum = np.ascontiguousarray(face_number, np.int32)
    out = np.zeros((fnum.shape[0], 3), np.float32)
    fc = None if faces is None else np.ascontiguousarray(faces, FACE_PATCH_DTYPE).copy()
    _check(_lib.load().vrad_bsp_apply_texlights(L.ref, _ptr(st), C.c_int(st.shape[0]), string_data, C.c_int64(len(string_data)), map_name.encode(),
                                                C.c_int(tl.shape[0]), _ptr(tl) if tl.shape[0] else None, C.c_int(fnum.shape[0]), _ptr(fnum), _ptr(fc), _ptr(out)),
           "vrad_bsp_apply_texlights")
    return out, fc


def rescale_lightmap_vecs(texinfo: np.ndarray, luxel_density: float) -> np.ndarray:
    t = np.ascontiguousarray(texinfo, TEXINFO).copy()
    _check(_lib.load().vrad_bsp_rescale_lightmap_vecs(C.c_int(t.shape[0]), _ptr(t), C.c_float(luxel_density)), "vrad_bsp_rescale_lightmap_vecs")
    return t


def face_extents(L: Lumps):
    n = L.faces.shape[0]
    mins, size, over = np.zeros((n, 2), np.int32), np.zeros((n, 2), np.int32), C.c_int()
    _check(_lib.load().vrad_bsp_face_extents(L.ref, _ptr(mins), _ptr(size), C.byref(over)), "vrad_bsp_face_extents")
    return mins, size, over.value


def make_parents(L: Lumps):
    npar, lpar = np.zeros(L.nodes.shape[0], np.int32), np.zeros(L.leafs.shape[0], np.int32)
    _check(_lib.load().vrad_bsp_make_parents(L.ref, _ptr(npar), _ptr(lpar)), "vrad_bsp_make_parents")
    return npar, lpar


def cluster_table(L: Lumps, n_clusters: int):
    first, leafs = np.zeros(n_clusters + 1, np.int32), np.zeros(L.leafs.shape[0], np.int32)
    _check(_lib.load().vrad_bsp_cluster_table(L.ref, C.c_int(n_clusters), _ptr(first), _ptr(leafs)), "vrad_bsp_cluster_table")
    return first, leafs[:first[-1]]


def vis_for_light_environment(L: Lumps, env=None):
    """BuildVisForLightEnvironment: (leaf flags uint8 [n_leafs], merged sky PVS bytes or None)."""
    flags = np.zeros(L.leafs.shape[0], np.uint8)
    row = (L.n_clusters + 7) // 8
    pvs, has = np.zeros(max(row, 1), np.uint8), C.c_int()
    h = env._h if env is not None else None
    _check(_lib.load().vrad_bsp_vis_for_light_environment(h, L.ref, _ptr(flags), _ptr(pvs), C.byref(has)), "vrad_bsp_vis_for_light_environment")
    return flags, (pvs[:row] if has.value else None)


def pair_edges(L: Lumps, smoothing_threshold: float = 0.7071067):
    nfv = int(L.faces["numedges"].astype(np.int64).sum())
    vn = np.zeros((nfv, 3), np.float32)
    first = np.zeros(L.faces.shape[0] + 1, np.int32)
    cap = 64 * max(1, L.faces.shape[0])
    nb = np.zeros(cap, np.int32)
    _check(_lib.load().vrad_bsp_pair_edges(L.ref, C.c_float(smoothing_threshold), _ptr(vn), _ptr(first), _ptr(nb), C.c_int(cap)), "vrad_bsp_pair_edges")
    return vn, first, nb[:first[-1]]


def save_vertex_normals(vertex_normals: np.ndarray):
    vn = np.ascontiguousarray(vertex_normals, np.float32).reshape(-1, 3)
    n = C.c_int()
    idx = np.zeros(vn.shape[0], np.uint16)
    normals = np.zeros((max(vn.shape[0], 1), 3), np.float32)
    _check(_lib.load().vrad_bsp_save_vertex_normals(C.c_int(vn.shape[0]), _ptr(vn), C.c_int(normals.shape[0]), _ptr(normals), _ptr(idx), C.byref(n)), "vrad_bsp_save_vertex_normals")
    return normals[:n.value].copy(), idx


def phong_normals(L: Lumps, vertex_normals, centroids, face, points, smoothing_threshold: float = 0.7071067):
    vn = np.ascontiguousarray(vertex_normals, np.float32); ce = np.ascontiguousarray(centroids, np.float32)
    fa = np.ascontiguousarray(face, np.int32); pt = np.ascontiguousarray(points, np.float32).reshape(-1, 3)
    out = np.zeros_like(pt)
    _check(_lib.load().vrad_bsp_phong_normals(L.ref, C.c_float(smoothing_threshold), _ptr(vn), _ptr(ce), C.c_int64(pt.shape[0]), _ptr(fa), _ptr(pt), _ptr(out)), "vrad_bsp_phong_normals")
    return out


def layout_lighting(L: Lumps, mins, size):
    n = L.faces.shape[0]
    faces = np.zeros(n, DFACE); first = np.zeros(n + 1, np.int64); nbytes = C.c_int64()
    mins = np.ascontiguousarray(mins, np.int32); size = np.ascontiguousarray(size, np.int32)
    _check(_lib.load().vrad_bsp_layout_lighting(L.ref, _ptr(mins), _ptr(size), _ptr(faces), _ptr(first), C.byref(nbytes)), "vrad_bsp_layout_lighting")
    return faces, first, nbytes.value


def face_luxels(L: Lumps, mins, size, luxel_first, face_origins=None):
    n = int(luxel_first[-1])
    pos, nrm, lf = np.zeros((n, 3), np.float32), np.zeros((n, 3), np.float32), np.zeros(n, np.int32)
    mins = np.ascontiguousarray(mins, np.int32); size = np.ascontiguousarray(size, np.int32)
    first = np.ascontiguousarray(luxel_first, np.int64)
    fo = None if face_origins is None else np.ascontiguousarray(face_origins, np.float32)
    _check(_lib.load().vrad_bsp_face_luxels(L.ref, _ptr(mins), _ptr(size), _ptr(fo), _ptr(first), _ptr(pos), _ptr(nrm), _ptr(lf)), "vrad_bsp_face_luxels")
    return pos, nrm, lf


def place_samples(L: Lumps, mins, size, luxel_first, pos, face_origins=None):
    """Samples moved onto their faces: (new positions, sample lightmap coordinates relative to the mins)."""
    mins = np.ascontiguousarray(mins, np.int32); size = np.ascontiguousarray(size, np.int32); first = np.ascontiguousarray(luxel_first, np.int64)
    p = np.ascontiguousarray(pos, np.float32).copy()
    st = np.zeros((p.shape[0], 2), np.float32)
    fo = None if face_origins is None else np.ascontiguousarray(face_origins, np.float32)
    _check(_lib.load().vrad_bsp_place_samples(L.ref, _ptr(mins), _ptr(size), _ptr(fo), _ptr(first), _ptr(p), _ptr(st)), "vrad_bsp_place_samples")
    return p, st


def color_to_rgbexp32(rgb) -> np.ndarray:
    rgb = np.ascontiguousarray(rgb, np.float32).reshape(-1, 3)
    out = np.zeros(rgb.shape[0], RGBEXP32)
    _check(_lib.load().vrad_color_to_rgbexp32(C.c_int64(rgb.shape[0]), _ptr(rgb), _ptr(out)), "vrad_color_to_rgbexp32")
    return out


def color_from_rgbexp32(colors) -> np.ndarray:
    c = np.ascontiguousarray(colors, RGBEXP32)
    out = np.zeros((c.shape[0], 3), np.float32)
    _check(_lib.load().vrad_color_from_rgbexp32(C.c_int64(c.shape[0]), _ptr(c), _ptr(out)), "vrad_color_from_rgbexp32")
    return out


def lightmap_finalize(env, direct, indirect=None) -> np.ndarray:
    """K5 on the device: RGBExp32(direct + indirect) per luxel."""
    d = np.ascontiguousarray(direct, np.float32).reshape(-1, 3)
    i = None if indirect is None else np.ascontiguousarray(indirect, np.float32).reshape(-1, 3)
    out = np.zeros(d.shape[0], RGBEXP32)
    _check(_lib.load().vrad_lightmap_finalize(env._h, C.c_int64(d.shape[0]), _ptr(d), _ptr(i), _ptr(out)), "vrad_lightmap_finalize")
    return out


def luxel_nearest_patch(luxel_face, pos, patch_face, origin, child1=None) -> np.ndarray:
    lf = np.ascontiguousarray(luxel_face, np.int32); p = np.ascontiguousarray(pos, np.float32).reshape(-1, 3)
    pf = np.ascontiguousarray(patch_face, np.int32); o = np.ascontiguousarray(origin, np.float32).reshape(-1, 3)
    c1 = None if child1 is None else np.ascontiguousarray(child1, np.int32)
    out = np.zeros(lf.shape[0], np.int32)
    _check(_lib.load().vrad_luxel_nearest_patch(C.c_int64(lf.shape[0]), _ptr(lf), _ptr(p), C.c_int(pf.shape[0]), _ptr(pf), _ptr(c1), _ptr(o), _ptr(out)),
           "vrad_luxel_nearest_patch")
    return out


def lightmap_finalize_patches(env, direct, luxel_patch, patch_total) -> np.ndarray:
    """K5 on the device with the bounced light of each luxel's patch."""
    d = np.ascontiguousarray(direct, np.float32).reshape(-1, 3); ix = np.ascontiguousarray(luxel_patch, np.int32)
    t = np.ascontiguousarray(patch_total, np.float32).reshape(-1, 3)
    out = np.zeros(d.shape[0], RGBEXP32)
    _check(_lib.load().vrad_lightmap_finalize_patches(env._h, C.c_int64(d.shape[0]), _ptr(d), _ptr(ix), C.c_int(t.shape[0]), _ptr(t), _ptr(out)),
           "vrad_lightmap_finalize_patches")
    return out


RADIAL_ENTRY = np.dtype([("patch", "<i4"), ("s", "<f4"), ("t", "<f4"), ("inv_ds", "<f4"), ("inv_dt", "<f4")])
assert RADIAL_ENTRY.itemsize == 20


def radial_entries(L: Lumps, mins, tree: dict, patch_face, face_origins=None, neighbour_first=None, neighbours=None):
    """Per lit face the leaf patches (own + smoothing neighbours) in the face's luxel space: (entry_first [n_faces + 1], entries)."""
    mins = np.ascontiguousarray(mins, np.int32)
    pf = np.ascontiguousarray(patch_face, np.int32); c1 = np.ascontiguousarray(tree["child1"], np.int32)
    org = np.ascontiguousarray(tree["origin"], np.float32)
    wf = np.ascontiguousarray(tree["wind_first"], np.int32); wc = np.ascontiguousarray(tree["wind_count"], np.int32)
    wp = np.ascontiguousarray(tree["wind_points"], np.float32)
    fo = None if face_origins is None else np.ascontiguousarray(face_origins, np.float32)
    nbf = None if neighbour_first is None else np.ascontiguousarray(neighbour_first, np.int32)
    nb = None if neighbours is None else np.ascontiguousarray(neighbours, np.int32)
    if nb is not None and nb.shape[0] == 0:
        nb = np.zeros(1, np.int32)
    first = np.zeros(L.faces.shape[0] + 1, np.int64); n = C.c_int64()
    l = _lib.load()
    args = (L.ref, _ptr(mins), _ptr(fo), C.c_int(pf.shape[0]), _ptr(pf), _ptr(c1), _ptr(org), _ptr(wf), _ptr(wc), _ptr(wp), _ptr(nbf), _ptr(nb))
    _check(l.vrad_bsp_radial_entries(*args, C.c_int64(0), _ptr(first), None, C.byref(n)), "vrad_bsp_radial_entries")
    entries = np.zeros(max(n.value, 1), RADIAL_ENTRY)
    _check(l.vrad_bsp_radial_entries(*args, C.c_int64(entries.shape[0]), _ptr(first), _ptr(entries), C.byref(n)), "vrad_bsp_radial_entries")
    return first, entries[:n.value]


def luxel_radial_light(env, luxel_face, luxel_first, size, entry_first, entries, patch_total, patch_bump=None) -> np.ndarray:
    """Bounced light per luxel by the radial filter; env = an Environment (device) or None (the same code on the host's cores)."""
    lf = np.ascontiguousarray(luxel_face, np.int32); first = np.ascontiguousarray(luxel_first, np.int64); size = np.ascontiguousarray(size, np.int32)
    ef = np.ascontiguousarray(entry_first, np.int64); en = np.ascontiguousarray(entries, RADIAL_ENTRY)
    if en.shape[0] == 0:
        en = np.zeros(1, RADIAL_ENTRY)
    tot = np.ascontiguousarray(patch_total, np.float32).reshape(-1, 3)
    bump = None if patch_bump is None else np.ascontiguousarray(patch_bump, np.float32).reshape(-1, 9)
    out = np.zeros((lf.shape[0], 3), np.float32)
    l = _lib.load()
    tail = (C.c_int64(lf.shape[0]), _ptr(lf), C.c_int(size.shape[0]), _ptr(first), _ptr(size), _ptr(ef), _ptr(en), C.c_int(tot.shape[0]), _ptr(tot), _ptr(bump), _ptr(out))
    if env is None:
        _check(l.vrad_luxel_radial_light_host(*tail), "vrad_luxel_radial_light_host")
    else:
        _check(l.vrad_luxel_radial_light(env._h, *tail), "vrad_luxel_radial_light")
    return out


def pack_lighting(L: Lumps, luxel_first, colors, lump_bytes: int) -> bytes:
    first = np.ascontiguousarray(luxel_first, np.int64); c = np.ascontiguousarray(colors, RGBEXP32)
    out = np.zeros(max(lump_bytes, 1), np.uint8)
    _check(_lib.load().vrad_bsp_pack_lighting(L.ref, _ptr(first), _ptr(c), _ptr(out), C.c_int64(lump_bytes)), "vrad_bsp_pack_lighting")
    return out[:lump_bytes].tobytes()


# ---- synthetic BSP v20 generator --------------------------------------------------------------------------------------

class _MapBuilder:
    def __init__(self):
        self.planes, self._plane_ix = [], {}
        self.verts, self._vert_ix = [], {}
        self.edges, self._edge_ix = [(0, 0)], {}          # edge 0 is never referenced (its sign could not be coded)
        self.surfedges, self.faces, self.texinfo, self.texdata = [], [], [], []
        self.brushes, self.brushsides = [], []

    def plane(self, normal, dist):
        """Index of the plane (normal, dist); planes are stored in pairs so that index ^ 1 is the opposite plane."""
        n = tuple(float(np.float32(c)) + 0.0 for c in normal); d = float(np.float32(dist)) + 0.0
        key = (n, d)
        if key in self._plane_ix:
            return self._plane_ix[key]
        neg = (tuple(-c + 0.0 for c in n), -d + 0.0)
        positive = sum(n) > 0
        first, second = (key, neg) if positive else (neg, key)
        base = len(self.planes)
        for k, (nn, dd) in enumerate((first, second)):
            axis = [i for i in range(3) if abs(nn[i]) == 1.0]
            self.planes.append((nn, dd, axis[0] if axis else 3))
            self._plane_ix[(nn, dd)] = base + k
        return self._plane_ix[key]

    def vertex(self, p):
        key = tuple(float(np.float32(c)) + 0.0 for c in p)
        if key not in self._vert_ix:
            self._vert_ix[key] = len(self.verts); self.verts.append(key)
        return self._vert_ix[key]

    def surfedge(self, a, b):
        if (a, b) in self._edge_ix:
            return self._edge_ix[(a, b)]
        if (b, a) in self._edge_ix:
            return -self._edge_ix[(b, a)]
        self._edge_ix[(a, b)] = len(self.edges); self.edges.append((a, b))
        return self._edge_ix[(a, b)]

    def face(self, pts, normal, texinfo, smoothing=0):
        pts = [np.asarray(p, np.float64) for p in pts]
        ids = [self.vertex(p) for p in pts]
        first = len(self.surfedges)
        for k in range(len(ids)):
            self.surfedges.append(self.surfedge(ids[k], ids[(k + 1) % len(ids)]))
        pn = self.plane(normal, float(np.dot(np.asarray(normal, np.float64), pts[0])))
        self.faces.append(dict(planenum=pn, side=pn & 1, firstedge=first, numedges=len(ids), texinfo=texinfo, smoothing=smoothing))
        return len(self.faces) - 1

    def box_brush(self, mn, mx, texinfo, contents=CONTENTS_SOLID, side_texinfo=None):
        """Axis-aligned brush: 6 sides with outward planes (-x +x -y +y -z +z)."""
        first = len(self.brushsides)
        for k, (axis, sign) in enumerate(((0, -1), (0, 1), (1, -1), (1, 1), (2, -1), (2, 1))):
            n = [0.0, 0.0, 0.0]; n[axis] = float(sign)
            d = sign * (mx[axis] if sign > 0 else mn[axis])
            ti = side_texinfo.get(k, texinfo) if side_texinfo else texinfo
            self.brushsides.append((self.plane(n, d), ti, 0, 0))
        self.brushes.append((first, 6, contents))
        return len(self.brushes) - 1


def _axes_for(normal):
    """Two in-plane unit axes (u, v) of an axial face."""
    a = int(np.argmax(np.abs(normal)))
    u = np.zeros(3); v = np.zeros(3)
    u[(a + 1) % 3] = 1.0; v[(a + 2) % 3] = 1.0
    return u, v


def synthetic_map(nx: int = 3, ny: int = 2, room: float = 512.0, boxes_per_room: int = 6, seed: int = 0x5EED0B5F, wall: float = 16.0,
                  door_w: float = 128.0, door_h: float = 256.0, sky_rooms=(), bump_rooms=(), radial_rooms=(), pvs_radius: int = 2,
                  with_brush_entity: bool = True, luxels_per_unit: float = 1.0 / 16.0, ramps: bool = False) -> tuple[Lumps, dict]:
    """A BSP v20 map of the multi-room grid: rooms of `room`^3 separated by `wall`-thick brushes with door openings, box occluder
    brushes, one leaf/cluster per room (+ the solid leaf 0), faces wound from shared vertices, run-length coded PVS.
    sky_rooms: rooms whose ceiling is a SURF_SKY face; bump_rooms: rooms whose floor is SURF_BUMPLIGHT; radial_rooms: leafs
    flagged LEAF_FLAGS_RADIAL; ramps: one wedge per room with a sloping top (a non-axial plane whose lightmap axes are world axes, so the
    texture normal differs from the face normal).  Returns (Lumps, meta) -- meta: entity text, per-face room, brush-entity placement."""
    from .scenes import SplitMix64, _place_boxes, compress_vis_rows
    rng = SplitMix64(seed)
    b = _MapBuilder()
    R, T = float(room), float(wall)
    # materials: texdata 0..3 plain, 4 sky
    names = ["concrete/floor01", "plaster/wall01", "metal/box01", "tile/bump01", "tools/toolsskybox"]
    refl = rng.uniform(12, 0.2, 0.7).reshape(4, 3)
    for k in range(5):
        r = refl[k] if k < 4 else (0.0, 0.0, 0.0)
        b.texdata.append((tuple(float(c) for c in r), k, 512, 512, 512, 512))
    string_data = b"".join(n.encode() + b"\0" for n in names)
    string_table = np.cumsum([0] + [len(n) + 1 for n in names[:-1]]).astype("<i4")

    tex_ix = {}

    def texinfo(normal, texdata, flags):
        key = (tuple(np.asarray(normal, np.float64)), texdata, flags)
        if key not in tex_ix:
            u, v = _axes_for(normal)
            tv = np.zeros((2, 4), np.float32); lv = np.zeros((2, 4), np.float32)
            tv[0, :3] = u * 0.25; tv[1, :3] = v * 0.25
            lv[0, :3] = u * luxels_per_unit; lv[1, :3] = v * luxels_per_unit
            tex_ix[key] = len(b.texinfo); b.texinfo.append((tv, lv, flags, texdata))
        return tex_ix[key]

    face_room, leaf_faces, leaf_brushes = [], {}, {}
    a0, a1 = (R - door_w) / 2, (R + door_w) / 2
    sky_tex = texinfo((0, 0, -1), 4, SURF_SKY | SURF_SKY2D * 0 | SURF_NOLIGHT)

    def quad(o, u, v, normal, ti, k, smoothing=0):
        o = np.asarray(o, np.float64); u = np.asarray(u, np.float64); v = np.asarray(v, np.float64)
        f = b.face([o, o + u, o + u + v, o + v], normal, ti, smoothing)
        face_room.append(k); leaf_faces.setdefault(k, []).append(f)
        return f

    box_list = {}
    for i in range(nx):
        for j in range(ny):
            k = i * ny + j
            x0 = i * R + (T / 2 if i > 0 else 0.0); x1 = (i + 1) * R - (T / 2 if i < nx - 1 else 0.0)
            y0 = j * R + (T / 2 if j > 0 else 0.0); y1 = (j + 1) * R - (T / 2 if j < ny - 1 else 0.0)
            floor_flags = SURF_BUMPLIGHT if k in bump_rooms else 0
            quad((x0, y0, 0), (x1 - x0, 0, 0), (0, y1 - y0, 0), (0, 0, 1), texinfo((0, 0, 1), 3 if floor_flags else 0, floor_flags), k)
            if k in sky_rooms:
                quad((x0, y0, R), (x1 - x0, 0, 0), (0, y1 - y0, 0), (0, 0, -1), sky_tex, k)
            else:
                quad((x0, y0, R), (x1 - x0, 0, 0), (0, y1 - y0, 0), (0, 0, -1), texinfo((0, 0, -1), 1, 0), k)
            walls = [((x0, y0, 0), (0, 1, 0), y1 - y0, (1, 0, 0), i > 0, j * R), ((x1, y0, 0), (0, 1, 0), y1 - y0, (-1, 0, 0), i < nx - 1, j * R),
                     ((x0, y0, 0), (1, 0, 0), x1 - x0, (0, 1, 0), j > 0, i * R), ((x0, y1, 0), (1, 0, 0), x1 - x0, (0, -1, 0), j < ny - 1, i * R)]
            for (o, u, length, nrm, door, cell0) in walls:
                o = np.asarray(o, np.float64); u = np.asarray(u, np.float64)
                ti = texinfo(nrm, 1, 0)
                if not door:
                    quad(o, u * length, (0, 0, R), nrm, ti, k, smoothing=1)
                else:                                                   # the door sits at [a0, a1] of the cell, whatever the inset
                    s0 = float(np.dot(u, o)) - cell0                    # where this wall starts inside its cell
                    quad(o, u * (a0 - s0), (0, 0, R), nrm, ti, k)
                    quad(o + u * (a1 - s0), u * (length - (a1 - s0)), (0, 0, R), nrm, ti, k)
                    quad(o + u * (a0 - s0) + np.array([0, 0, door_h]), u * (a1 - a0), (0, 0, R - door_h), nrm, ti, k)
            # occluder boxes: brush + 6 faces each
            bmins, bmaxs = _place_boxes(rng, boxes_per_room, x0 + 16, y0 + 16, x1 - 16, y1 - 16, 0.0)
            box_list[k] = (bmins, bmaxs)
            for (mn, mx) in zip(bmins, bmaxs):
                mn = np.asarray(mn, np.float64); mx = np.asarray(mx, np.float64)
                br = b.box_brush(mn, mx, texinfo((0, 0, 1), 2, 0))
                leaf_brushes.setdefault(k, []).append(br)
                d = mx - mn
                quad((mn[0], mn[1], mx[2]), (d[0], 0, 0), (0, d[1], 0), (0, 0, 1), texinfo((0, 0, 1), 2, 0), k)
                quad((mn[0], mn[1], mn[2]), (0, d[1], 0), (0, 0, d[2]), (-1, 0, 0), texinfo((-1, 0, 0), 2, 0), k)
                quad((mx[0], mn[1], mn[2]), (0, d[1], 0), (0, 0, d[2]), (1, 0, 0), texinfo((1, 0, 0), 2, 0), k)
                quad((mn[0], mn[1], mn[2]), (d[0], 0, 0), (0, 0, d[2]), (0, -1, 0), texinfo((0, -1, 0), 2, 0), k)
                quad((mn[0], mx[1], mn[2]), (d[0], 0, 0), (0, 0, d[2]), (0, 1, 0), texinfo((0, 1, 0), 2, 0), k)
            if ramps:                                                   # a wedge: 128 x 64 base, rising 64 units along +x
                rx0, rx1, ry0, ry1, rh = x0 + 40.0, x0 + 168.0, y0 + 40.0, y0 + 104.0, 64.0
                n_top = np.array([-rh, 0.0, rx1 - rx0]); n_top /= np.linalg.norm(n_top)
                first = len(b.brushsides)
                for (nn, dd) in (((0.0, 0.0, -1.0), 0.0), ((1.0, 0.0, 0.0), rx1), ((0.0, -1.0, 0.0), -ry0), ((0.0, 1.0, 0.0), ry1),
                                 (tuple(n_top), float(np.dot(n_top, (rx0, ry0, 0.0))))):
                    b.brushsides.append((b.plane(nn, dd), texinfo((0, 0, 1), 2, 0), 0, 0))
                b.brushes.append((first, 5, CONTENTS_SOLID))
                leaf_brushes.setdefault(k, []).append(len(b.brushes) - 1)
                for pts, nrm in (([(rx0, ry0, 0), (rx1, ry0, rh), (rx1, ry1, rh), (rx0, ry1, 0)], tuple(n_top)),
                                 ([(rx1, ry0, 0), (rx1, ry1, 0), (rx1, ry1, rh), (rx1, ry0, rh)], (1.0, 0.0, 0.0)),
                                 ([(rx0, ry0, 0), (rx1, ry0, 0), (rx1, ry0, rh)], (0.0, -1.0, 0.0)),
                                 ([(rx0, ry1, 0), (rx1, ry1, rh), (rx1, ry1, 0)], (0.0, 1.0, 0.0))):
                    f = b.face([np.asarray(q, np.float64) for q in pts], nrm, texinfo(nrm, 2, 0))
                    face_room.append(k); leaf_faces.setdefault(k, []).append(f)
    n_world_faces = len(b.faces)

    # structural brushes: outer shell slabs and the interior walls (three pieces around each door)
    X, Y = nx * R, ny * R
    wall_tex = texinfo((0, 0, 1), 1, 0)
    shell = [((-T, -T, -T), (X + T, Y + T, 0.0), None), ((-T, -T, 0.0), (0.0, Y + T, R), None), ((X, -T, 0.0), (X + T, Y + T, R), None),
             ((0.0, -T, 0.0), (X, 0.0, R), None), ((0.0, Y, 0.0), (X, Y + T, R), None)]
    structural = []
    for (mn, mx, _) in shell:
        structural.append((b.box_brush(mn, mx, wall_tex), mn, mx))
    # the roof: one slab per room so that sky rooms can carry the sky texture on the inward (-z) side
    for i in range(nx):
        for j in range(ny):
            k = i * ny + j
            side_tex = {4: sky_tex} if k in sky_rooms else None
            mn, mx = (i * R, j * R, R), ((i + 1) * R, (j + 1) * R, R + T)
            structural.append((b.box_brush(mn, mx, wall_tex, side_texinfo=side_tex), mn, mx))
    for i in range(1, nx):                                              # walls between columns i-1 and i
        for j in range(ny):
            xa, xb, yb = i * R - T / 2, i * R + T / 2, j * R
            for (mn, mx) in (((xa, yb, 0.0), (xb, yb + a0, R)), ((xa, yb + a1, 0.0), (xb, yb + R, R)), ((xa, yb + a0, door_h), (xb, yb + a1, R))):
                structural.append((b.box_brush(mn, mx, wall_tex), mn, mx))
    for j in range(1, ny):
        for i in range(nx):
            ya, yb_, xb = j * R - T / 2, j * R + T / 2, i * R
            for (mn, mx) in (((xb, ya, 0.0), (xb + a0, yb_, R)), ((xb + a1, ya, 0.0), (xb + R, yb_, R)), ((xb + a0, ya, door_h), (xb + a1, yb_, R))):
                structural.append((b.box_brush(mn, mx, wall_tex), mn, mx))
    for (br, mn, mx) in structural:                                     # a structural brush is listed by every room leaf it touches
        for i in range(nx):
            for j in range(ny):
                if mn[0] <= (i + 1) * R and mx[0] >= i * R and mn[1] <= (j + 1) * R and mx[1] >= j * R:
                    leaf_brushes.setdefault(i * ny + j, []).append(br)

    # a shadow-casting brush entity (func_brush with vrad_brush_cast_shadows): model 1, its own brush, leaf and faces
    ent = None
    if with_brush_entity:
        mn, mx = np.array([-32.0, -8.0, 0.0]), np.array([32.0, 8.0, 96.0])      # model space; the entity's origin places it
        ent_brush = b.box_brush(mn, mx, texinfo((0, 0, 1), 2, 0))
        ent_first_face = len(b.faces)
        d = mx - mn
        for (o, u, v, nrm) in (((mn[0], mn[1], mx[2]), (d[0], 0, 0), (0, d[1], 0), (0, 0, 1)), ((mn[0], mn[1], mn[2]), (0, d[1], 0), (0, 0, d[2]), (-1, 0, 0)),
                               ((mx[0], mn[1], mn[2]), (0, d[1], 0), (0, 0, d[2]), (1, 0, 0)), ((mn[0], mn[1], mn[2]), (d[0], 0, 0), (0, 0, d[2]), (0, -1, 0)),
                               ((mn[0], mx[1], mn[2]), (d[0], 0, 0), (0, 0, d[2]), (0, 1, 0))):
            b.face([np.asarray(o, np.float64), np.asarray(o) + np.asarray(u), np.asarray(o) + np.asarray(u) + np.asarray(v), np.asarray(o) + np.asarray(v)], nrm, texinfo(nrm, 2, 0))
            face_room.append(-1)
        ent = dict(model=1, brush=ent_brush, first_face=ent_first_face, n_faces=len(b.faces) - ent_first_face,
                   origin=np.float32([R / 2, R / 2 + 40.0, 0.0]), angles=np.float32([0.0, 30.0, 0.0]))

    # leafs: 0 = the solid leaf, 1 + k = room k, then the brush entity's leaf; tree = axial splits at the room boundaries
    n_rooms = nx * ny
    leafs = np.zeros(1 + n_rooms + (1 if ent else 0), DLEAF)
    leafs[0]["contents"] = CONTENTS_SOLID; leafs[0]["cluster"] = -1
    lf_list, lb_list = [], []
    for k in range(n_rooms):
        i, j = divmod(k, ny)
        lf = leafs[1 + k]
        lf["cluster"] = k
        lf["area_flags"] = 1 | ((LEAF_FLAGS_RADIAL if k in radial_rooms else 0) << 9)
        lf["mins"] = (int(i * R), int(j * R), 0); lf["maxs"] = (int((i + 1) * R), int((j + 1) * R), int(R))
        lf["firstleafface"] = len(lf_list); lf["numleaffaces"] = len(leaf_faces.get(k, [])); lf_list += leaf_faces.get(k, [])
        lf["firstleafbrush"] = len(lb_list); lf["numleafbrushes"] = len(leaf_brushes.get(k, [])); lb_list += leaf_brushes.get(k, [])
        lf["leaf_water_data"] = -1
    if ent:
        lf = leafs[1 + n_rooms]
        lf["contents"] = CONTENTS_SOLID; lf["cluster"] = -1; lf["leaf_water_data"] = -1
        lf["firstleafbrush"] = len(lb_list); lf["numleafbrushes"] = 1; lb_list.append(ent["brush"])
    nodes = []

    def build(i0, i1, j0, j1):
        if i1 - i0 == 1 and j1 - j0 == 1:
            return -1 - (1 + i0 * ny + j0)
        me = len(nodes); nodes.append(None)
        if i1 - i0 >= j1 - j0:
            m = (i0 + i1) // 2
            pn = b.plane((1, 0, 0), m * R); front, back = build(m, i1, j0, j1), build(i0, m, j0, j1)
        else:
            m = (j0 + j1) // 2
            pn = b.plane((0, 1, 0), m * R); front, back = build(i0, i1, m, j1), build(i0, i1, j0, m)
        nodes[me] = (pn, front, back, (int(i0 * R), int(j0 * R), 0), (int(i1 * R), int(j1 * R), int(R)))
        return me
    if n_rooms == 1:                                                    # a single room still needs a node: split off the solid leaf below the floor
        nodes.append((b.plane((0, 0, 1), 0.0), -1 - 1, -1 - 0, (0, 0, 0), (int(R), int(R), int(R))))
    else:
        build(0, nx, 0, ny)

    pvs = np.zeros((n_rooms, n_rooms), np.uint8)
    for ka in range(n_rooms):
        for kb in range(n_rooms):
            (ia, ja), (ib, jb) = divmod(ka, ny), divmod(kb, ny)
            pvs[ka, kb] = 1 if (ka == kb or (ia == ib and abs(ja - jb) <= pvs_radius) or (ja == jb and abs(ia - ib) <= pvs_radius)) else 0
    rows, ofs = compress_vis_rows(pvs)
    head = np.int32(n_rooms).tobytes() + (ofs + np.where(ofs >= 0, 4 + 8 * n_rooms, 0)).astype("<i4").tobytes()
    visdata = head + rows

    planes = np.zeros(len(b.planes), DPLANE)
    for k, (n, d, t) in enumerate(b.planes):
        planes[k] = (n, d, t)
    faces = np.zeros(len(b.faces), DFACE)
    for k, f in enumerate(b.faces):
        faces[k]["planenum"] = f["planenum"]; faces[k]["side"] = f["side"]; faces[k]["firstedge"] = f["firstedge"]; faces[k]["numedges"] = f["numedges"]
        faces[k]["texinfo"] = f["texinfo"]; faces[k]["dispinfo"] = -1; faces[k]["fog_volume"] = -1; faces[k]["styles"] = 255; faces[k]["lightofs"] = -1
        faces[k]["orig_face"] = k; faces[k]["smoothing_groups"] = f["smoothing"]
    tinfo = np.zeros(len(b.texinfo), TEXINFO)
    for k, (tv, lv, fl, td) in enumerate(b.texinfo):
        tinfo[k]["texture_vecs"] = tv; tinfo[k]["lightmap_vecs"] = lv; tinfo[k]["flags"] = fl; tinfo[k]["texdata"] = td
    tdata = np.zeros(len(b.texdata), DTEXDATA)
    for k, t in enumerate(b.texdata):
        tdata[k] = t
    models = np.zeros(2 if ent else 1, DMODEL)
    models[0]["mins"] = (-T, -T, -T); models[0]["maxs"] = (X + T, Y + T, R + T); models[0]["headnode"] = 0
    models[0]["firstface"] = 0; models[0]["numfaces"] = n_world_faces
    if ent:
        models[1]["mins"] = (-32, -8, 0); models[1]["maxs"] = (32, 8, 96); models[1]["headnode"] = -1 - (1 + n_rooms)
        models[1]["firstface"] = ent["first_face"]; models[1]["numfaces"] = ent["n_faces"]
    nd = np.zeros(len(nodes), DNODE)
    for k, (pn, fr, bk, mn, mx) in enumerate(nodes):
        nd[k]["planenum"] = pn; nd[k]["children"] = (fr, bk); nd[k]["mins"] = mn; nd[k]["maxs"] = mx
    brushes = np.zeros(len(b.brushes), DBRUSH)
    for k, t in enumerate(b.brushes):
        brushes[k] = t
    sides = np.zeros(len(b.brushsides), DBRUSHSIDE)
    for k, t in enumerate(b.brushsides):
        sides[k] = t
    L = Lumps(planes=planes, vertexes3=np.asarray(b.verts, np.float32), edges=np.asarray(b.edges, "<u2").view(DEDGE).reshape(-1),
              surfedges=np.asarray(b.surfedges, "<i4"), faces=faces, texinfo=tinfo, texdata=tdata, models=models, nodes=nd, leafs=leafs,
              leaffaces=np.asarray(lf_list, "<u2"), leafbrushes=np.asarray(lb_list, "<u2"), brushes=brushes, brushsides=sides,
              n_areas=2, visdata=visdata)
    ents = ['{\n"classname" "worldspawn"\n}']
    if ent:
        o, a = ent["origin"], ent["angles"]
        ents.append('{\n"classname" "func_brush"\n"model" "*1"\n"vrad_brush_cast_shadows" "1"\n"origin" "%g %g %g"\n"angles" "%g %g %g"\n}' % (*o, *a))
    for k in range(n_rooms):
        i, j = divmod(k, ny)
        ents.append('{\n"classname" "light"\n"origin" "%g %g %g"\n"_light" "255 240 220 300"\n}' % (i * R + R / 2, j * R + R / 2, R - 64))
    if sky_rooms:
        ents.append('{\n"classname" "light_environment"\n"origin" "64 64 64"\n"_light" "255 255 240 200"\n"_ambient" "120 140 180 60"\n"pitch" "-60"\n"angles" "0 40 0"\n}')
    meta = dict(entities="\n".join(ents) + "\n", face_room=np.asarray(face_room, np.int32), brush_entity=ent, pvs=pvs, boxes=box_list,
                string_data=string_data, string_table=string_table, n_rooms=n_rooms, nx=nx, ny=ny, room=R, wall=T)
    return L, meta


def write_bsp(path: str, L: Lumps, meta: dict, lighting: bytes | None = None) -> None:
    """Write the map as a BSP v20 file through the library's container."""
    f = BspFile()
    f.set_lumps(L, meta.get("entities", ""))
    if "string_data" in meta:
        f.set(LUMP["TEXDATA_STRING_DATA"], meta["string_data"]); f.set(LUMP["TEXDATA_STRING_TABLE"], meta["string_table"])
    if lighting is not None:
        f.set(LUMP["LIGHTING"], lighting, version=1)
    f.save(path)
    f.close()
