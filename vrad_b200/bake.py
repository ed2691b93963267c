"""File-driven bake: .bsp in -> lit .bsp out, every step through the C-ABI of libvradcuda.so.

This is the Python mirror of the sequence the reference's tasks run (cmd/tasks/loadbsp/main.go:38-160 ->
rad.Start, rad/start.go:21-98 -> cmd/tasks/computerad/main.go:5-10 -> cmd/tasks/finish/main.go:8-40), with the reference's stubbed
or absent stages (tracer, transfers, direct light, bounces, lightmap write-back) replaced by the library:

    load        vrad_bspfile_open / _lumps                       loadBSP, cache.BuildLumpCache
    geometry    vrad_env_add_bsp, vrad_env_build                 ExtractBrushEntityShadowCasters, addBrushesForRayTrace, SetupAccelerationStructure
    patches     vrad_bsp_face_patches, vrad_patches_subdivide    patches.MakePatches, SubdividePatches
    vis         vrad_pvs_from_vis_lump                           lightmap.GetVisCache / DecompressVis
    lights      vrad_lights_from_entities                        lightmap.CreateDirectLights
    luxels      vrad_bsp_face_extents / _layout_lighting / _face_luxels
    K3          vrad_direct_light (luxels, then patch origins -> emit0)
    K2, K4      vrad_build_transfers, vrad_bounce
    K5          vrad_luxel_nearest_patch, vrad_lightmap_finalize_patches
    write       vrad_bsp_pack_lighting, vrad_bspfile_set_lump / _save

`prepare` and `finish` are host-side and GPU-side product code; `light` only needs the Environment call surface, so the GPU
tests run it a second time on the CPU oracle's environment to check the device stages end to end.
"""
from __future__ import annotations

import numpy as np

from . import bspfile as B
from .environment import bump_normals, lights_from_entities, lights_from_patches, light_for_string, pvs_from_vis_lump, subdivide_patches
from .lib import LIGHT_ENTITY_DTYPE, VradError


def anorms() -> np.ndarray:
    """vmath.Anorms (vmath/constants.go:15,21-184): the 162 sky-ambient sample directions."""
    import os
    return np.loadtxt(os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", "anorms.txt"), dtype=np.float32)


def parse_entities(text: str) -> list[dict]:
    """The entity lump as a list of key -> value dicts (`{ "key" "value" ... }` blocks; what vmf.NewReader(...).Read() yields in
    loadbsp/main.go:176-181).  Later duplicates of a key win, as Entity.ValueForKey's linear search from the list head does."""
    ents, cur, i, n = [], None, 0, len(text)
    while i < n:
        c = text[i]
        if c == "{":
            cur = {}; i += 1
        elif c == "}":
            if cur is not None:
                ents.append(cur)
            cur = None; i += 1
        elif c == '"':
            j = text.index('"', i + 1)
            k0 = text.index('"', j + 1); k1 = text.index('"', k0 + 1)
            if cur is not None:
                cur[text[i + 1:j]] = text[k0 + 1:k1]
            i = k1 + 1
        else:
            i += 1
    return ents


def _floats(s: str, n: int):
    v = [float(x) for x in s.split()[:n]] if s else []
    return v + [0.0] * (n - len(v))


def light_entities(ents: list[dict]) -> np.ndarray:
    """light / light_spot / light_environment entities -> vrad_light_entity records (the conversions Entity.FloatForKey /
    VectorForKey / LightForKey do, common/types/entity.go:60-101)."""
    classes = {"light": 0, "light_spot": 1, "light_environment": 2}
    by_name = {e.get("targetname"): e for e in ents if e.get("targetname")}
    out = []
    for e in ents:
        if e.get("classname") not in classes:
            continue
        r = np.zeros(1, LIGHT_ENTITY_DTYPE)[0]
        r["classname"] = classes[e["classname"]]
        r["origin"] = _floats(e.get("origin", ""), 3)
        for key, ok, dst in (("_light", "light_ok", "light"), ("_ambient", "ambient_ok", "ambient")):
            try:
                r[dst] = light_for_string(e[key]); r[ok] = 1
            except (KeyError, VradError):
                r[ok] = 0
        tgt = by_name.get(e.get("target"))
        if tgt is not None:
            r["has_target"] = 1; r["target_origin"] = _floats(tgt.get("origin", ""), 3)
        r["angles"] = _floats(e.get("angles", ""), 3)
        for key, dst in (("pitch", "pitch"), ("angle", "angle"), ("_inner_cone", "inner_cone"), ("_cone", "cone"), ("_exponent", "exponent"),
                         ("_fifty_percent_distance", "fifty_percent_distance"), ("_zero_percent_distance", "zero_percent_distance"),
                         ("_constant_attn", "constant_attn"), ("_linear_attn", "linear_attn"), ("_quadratic_attn", "quadratic_attn"),
                         ("_distance", "distance")):
            r[dst] = _floats(e.get(key, ""), 1)[0]
        r["hardfalloff"] = int(_floats(e.get("_hardfalloff", ""), 1)[0])
        out.append(r)
    return np.asarray(out, LIGHT_ENTITY_DTYPE) if out else np.zeros(0, LIGHT_ENTITY_DTYPE)


def shadow_casters(ents: list[dict]):
    """Brush entities with vrad_brush_cast_shadows (ExtractBrushEntityShadowCasters, loadbsp/main.go:186-211): model "*N" -> N."""
    model, origin, angles = [], [], []
    for e in ents:
        if "vrad_brush_cast_shadows" in e and e.get("model", "").startswith("*"):
            model.append(int(e["model"][1:])); origin.append(_floats(e.get("origin", ""), 3)); angles.append(_floats(e.get("angles", ""), 3))
    return np.asarray(model, np.int32), np.asarray(origin, np.float32).reshape(-1, 3), np.asarray(angles, np.float32).reshape(-1, 3)


def model_origins(L: B.Lumps, ents: list[dict]) -> np.ndarray:
    """The "origin" key of each brush model's entity (patches.MakePatches, rad/patches/build.go:38-45); the world's is zero."""
    out = np.zeros((L.models.shape[0], 3), np.float32)
    for e in ents:
        m = e.get("model", "")
        if m.startswith("*") and m[1:].isdigit() and int(m[1:]) < out.shape[0]:
            out[int(m[1:])] = _floats(e.get("origin", ""), 3)
    return out


def _point_cluster(L: B.Lumps, p) -> int:
    """trace.PointLeafnum (raytracer/trace/pointleaf.go:8-33) on the host, for the handful of per-model lookups."""
    node = int(L.models[0]["headnode"])
    while node >= 0:
        nd = L.nodes[node]; pl = L.planes[int(nd["planenum"])]
        t = int(pl["type"])
        d = (np.float32(p[t]) - pl["dist"]) if t < 3 else (np.float32(np.dot(pl["normal"], np.asarray(p, np.float32))) - pl["dist"])
        node = int(nd["children"][1 if d < 0 else 0])
    return int(L.leafs[-1 - node]["cluster"])


def prepare(L: B.Lumps, entity_text: str, min_chop: float = 4.0, max_chop: float = 4.0, lights_rad: str | None = None,
            texdata_strings=None, map_name: str = "", lights_rad_hdr: bool = False, smoothing_threshold: float = 0.7071067,
            place_samples: bool = True, luxel_density: float = 1.0) -> dict:
    """Everything the device stages take, from the lumps (host code in the library; no GPU needed).
    lights_rad = the text of a lights.rad file, texdata_strings = (LUMP_TEXDATA_STRING_TABLE as int32, LUMP_TEXDATA_STRING_DATA bytes):
    faces whose material is a texlight get Patch.BaseLight and their leaf patches become EMIT_SURFACE lights (CreateDirectLights)."""
    if luxel_density < 1.0:                                           # rad.Start (rad/start.go:21-66): luxels no denser than -luxeldensity
        L = L.replace(texinfo=B.rescale_lightmap_vecs(L.texinfo, luxel_density))
    ents = parse_entities(entity_text)
    cm, co, ca = shadow_casters(ents)
    tri_ids, tri_verts = B.raytrace_triangles(L, cm, co, ca)
    origins = model_origins(L, ents)
    fp = B.face_patches(L, origins, max_chop)
    base_light = np.zeros((fp["faces"].shape[0], 3), np.float32)
    if lights_rad is not None and texdata_strings is not None:
        table, _, _ = B.texlights_parse(lights_rad, lights_rad_hdr)
        base_light, fp["faces"] = B.apply_texlights(L, texdata_strings[0], texdata_strings[1], map_name, table, fp["face_number"], fp["faces"])
    tree = subdivide_patches(fp["faces"], fp["points"], min_chop=min_chop)
    face_of_patch = fp["face_number"][tree["face"]]                       # face lump index of every patch
    # lightmap.PairEdges (rad/start.go:66-79) -> smoothed vertex normals; CreateChildPatch gives every child patch the phong normal at its
    # origin (lightmap.GetPhongNormal, rad/patches/subdivide.go:385); root patches keep the plane normal (rad/patches/face.go:152).
    # cache.faceCentroids[fn] = root origin - face offset (face.go:151); the lookup runs in the face's model space.
    vn, nb_first, nb = B.pair_edges(L, smoothing_threshold)
    n_faces = L.faces.shape[0]
    face_origin = np.zeros((n_faces, 3), np.float32)
    for m in range(L.models.shape[0]):
        face_origin[int(L.models[m]["firstface"]):int(L.models[m]["firstface"]) + int(L.models[m]["numfaces"])] = origins[m]
    centroids = np.zeros((n_faces, 3), np.float32)
    roots = tree["parent"] == -1
    centroids[face_of_patch[roots]] = tree["origin"][roots] - face_origin[face_of_patch[roots]]
    kids = np.nonzero(~roots)[0]
    if kids.size:
        tree["normal"][kids] = B.phong_normals(L, vn, centroids, face_of_patch[kids], tree["origin"][kids] - face_origin[face_of_patch[kids]], smoothing_threshold)
    # cluster of a face = cluster of the leaf that lists it (leaffaces); faces of brush models: the leaf their model origin is in
    face_cluster = np.full(n_faces, -1, np.int32)
    for lf in L.leafs:
        for k in range(int(lf["numleaffaces"])):
            f = int(L.leaffaces[int(lf["firstleafface"]) + k])
            if face_cluster[f] < 0:
                face_cluster[f] = lf["cluster"]
    for m in range(1, L.models.shape[0]):
        mod = L.models[m]
        centre = origins[m] + 0.5 * (mod["mins"] + mod["maxs"])
        face_cluster[int(mod["firstface"]):int(mod["firstface"]) + int(mod["numfaces"])] = _point_cluster(L, centre)
    nc = L.n_clusters
    pvs = None
    if nc:
        ofs = np.frombuffer(L.visdata[4:4 + 8 * nc].tobytes(), "<i4").reshape(nc, 2)
        pvs = pvs_from_vis_lump(nc, ofs, L.visdata.tobytes())
    has_radial = bool(np.any((L.leafs["area_flags"].astype(np.int32) >> 9) & B.LEAF_FLAGS_RADIAL))
    sky_pvs = None if has_radial else B.vis_for_light_environment(L)[1]       # radial-vis maps: BuildVisForLightEnvironment needs the device
    mins, size, oversize = B.face_extents(L)
    faces_lit, luxel_first, lump_bytes = B.layout_lighting(L, mins, size)
    Llit = L.replace(faces=faces_lit)
    lux_pos, lux_normal, lux_face = B.face_luxels(Llit, mins, size, luxel_first, face_origin)
    grid_pos = lux_pos
    if place_samples:                                                 # grid points that hang over a face's edge sit inside the next brush
        lux_pos, _ = B.place_samples(Llit, mins, size, luxel_first, lux_pos, face_origin)
    # the sample normal is the phong normal at the sample (upstream BuildFacelights; same GetPhongNormal).  The three extra blocks of a
    # bump-mapped face keep the bump basis face_luxels built around the flat normal.
    flags_of_lux = L.texinfo["flags"][L.faces["texinfo"][lux_face]]
    k_in_face = np.arange(lux_face.shape[0]) - luxel_first[lux_face]
    per_block = (size[lux_face, 0] + 1).astype(np.int64) * (size[lux_face, 1] + 1)
    flat_block = np.nonzero(((flags_of_lux & B.SURF_BUMPLIGHT) == 0) | (k_in_face < per_block))[0]
    if flat_block.size:
        on_surface = lux_pos[flat_block] - lux_normal[flat_block] - face_origin[lux_face[flat_block]]
        lux_normal[flat_block] = B.phong_normals(L, vn, centroids, lux_face[flat_block], on_surface, smoothing_threshold)
    lux_patch = B.luxel_nearest_patch(lux_face, lux_pos, face_of_patch, tree["origin"], tree["child1"])
    sky = fp["faces"]["sky"][tree["face"]].astype(np.uint8)
    # Patch.NeedsBumpMap (SURF_BUMPLIGHT, rad/patches/face.go:66-70) and the bump basis of every such patch: upstream GetBumpNormals with the
    # face's texture vectors, the face normal and the patch's (phong) normal
    needs_bump = fp["needs_bump"][tree["face"]].astype(np.uint8)
    bump_basis = np.zeros((tree["origin"].shape[0], 3, 3), np.float32)
    for p in np.nonzero(needs_bump)[0]:
        tx = L.texinfo[L.faces["texinfo"][face_of_patch[p]]]
        flat = L.planes["normal"][L.faces["planenum"][face_of_patch[p]]]
        bump_basis[p] = bump_normals(tx["texture_vecs"][0][:3], tx["texture_vecs"][1][:3], flat, tree["normal"][p])
    from .scenes import Bsp
    bsp = Bsp(L.nodes["planenum"].astype(np.int32), L.nodes["children"].astype(np.int32), L.planes["normal"].astype(np.float32),
              L.planes["dist"].astype(np.float32), L.planes["type"].astype(np.int32), L.leafs["cluster"].astype(np.int32),
              (L.leafs["area_flags"].astype(np.int32) & 0x1ff), L.leafs["mins"].copy(), L.leafs["maxs"].copy(), max(int(L.n_areas), 1))
    lights = lights_from_entities(light_entities(ents))
    if base_light.any():                                              # surface lights first, as CreateDirectLights does (lights.go:49-82, then :90-113)
        surf = lights_from_patches(tree["origin"], tree["normal"], base_light[tree["face"]], tree["area"], fp["scale"][tree["face"]],
                                   fp["base_area"][tree["face"]], tree["child1"])
        lights = np.concatenate([surf, lights])
    entry_first, entries = B.radial_entries(Llit, mins, tree, face_of_patch, face_origin, nb_first, nb)
    return dict(ents=ents, bsp=bsp, sky_pvs=sky_pvs, needs_bump=needs_bump, bump_basis=bump_basis, face_origin=face_origin, face_centroids=centroids, lux_grid_pos=grid_pos, lm_mins=mins, lm_size=size, vertex_normals=vn,
                radial_first=entry_first, radial_entries=entries, base_light=base_light[tree["face"]].astype(np.float32), tri_ids=tri_ids, tri_verts=tri_verts, tree=tree, refl=fp["reflectivity"][tree["face"]].astype(np.float32),
                cluster=face_cluster[face_of_patch].astype(np.int32), flags=sky, pvs=pvs, lights=lights,
                lumps=Llit, luxel_first=luxel_first, lump_bytes=lump_bytes, lux_pos=lux_pos, lux_normal=lux_normal, lux_face=lux_face,
                lux_patch=lux_patch, oversize=oversize, face_of_patch=face_of_patch)


def patch_clusters(env, prep) -> np.ndarray:
    """Patch.ClusterNumber, assigned AFTER subdivision as the reference does (rad/patches/subdivide.go:92-116): faces span leaves, so
    every patch -- children included -- asks ClusterFromPoint for its own origin; an origin in solid space (cluster -1: detail and
    displacement surfaces) takes the cluster of the first winding point that is not.  A patch that stays at -1 is in no cluster's
    child list: vrad_build_transfers leaves it out, as receiver and as emitter."""
    t = prep["tree"]
    n = t["origin"].shape[0]
    if prep["pvs"] is None or n == 0:
        return np.zeros(n, np.int32)
    (env.bsp_upload if hasattr(env, "bsp_upload") else env.bsp_set)(prep["bsp"])
    cl = np.asarray(env.cluster_from_point(np.ascontiguousarray(t["origin"], np.float32))).astype(np.int32)
    bad = np.nonzero(cl < 0)[0]
    if bad.size:
        first, count = t["wind_first"][bad].astype(np.int64), t["wind_count"][bad].astype(np.int64)
        owner = np.repeat(bad, count)
        idx = np.concatenate([np.arange(f, f + c) for f, c in zip(first, count)]) if owner.size else np.zeros(0, np.int64)
        if idx.size:
            wc = np.asarray(env.cluster_from_point(np.ascontiguousarray(t["wind_points"][idx], np.float32)))
            for p, c in zip(owner, wc):                # first winding point with a cluster wins
                if cl[p] < 0 and c >= 0:
                    cl[p] = c
    return cl


def all_gather_blocks(local: np.ndarray, parts, rank: int, device=None) -> np.ndarray:
    """Concatenation over ranks of the contiguous blocks `parts` (range_partition), each rank holding `local` = its own block.
    torch.distributed.all_gather wants equal shapes, so blocks travel padded to the largest one.  device: None (gloo, host tensors)
    or a CUDA device (nccl)."""
    import torch
    import torch.distributed as dist
    width = max(b - a for a, b in parts)
    mine = torch.zeros((width,) + local.shape[1:], dtype=torch.from_numpy(local[:0].copy()).dtype)
    mine[: local.shape[0]] = torch.from_numpy(np.ascontiguousarray(local))
    if device is not None:
        mine = mine.to(device)
    bufs = [torch.empty_like(mine) for _ in parts]
    dist.all_gather(bufs, mine)
    return np.concatenate([bufs[r][: b - a].cpu().numpy() for r, (a, b) in enumerate(parts)], axis=0)


def light(env, prep: dict, bounces: int = 8, early_out: bool = True, rank: int = 0, world: int = 1, device=None, use_light_pvs: bool = True,
          fast_tree: bool = False, texture_shadows: bool = False, poly_form_factor: bool = False) -> dict:
    """The device stages, on any object with the Environment call surface: geometry + kd build (K1), transfers (K2), direct
    light on the luxels and on the patches (K3; the patch value is Patch.DirectLight, what the first bounce emits), bounces (K4).
    world > 1 (one process per GPU, torch.distributed initialised): luxels and patch origins are independent work items, each
    rank lights its contiguous range (sharding.range_partition) and the blocks are all-gathered; the transfer rows and the
    per-bounce radiance exchange are sharded inside the environment itself (vrad_config.rank / world, vrad_comm_init)."""
    from .sharding import range_partition
    t = prep["tree"]
    env.add_triangles(prep["tri_ids"], prep["tri_verts"].reshape(-1, 9), np.zeros(prep["tri_ids"].shape[0], np.uint8))
    if fast_tree and hasattr(env, "build_fast"):
        env.build_fast()                                              # RTE_FLAGS_FAST_TREE_GENERATION: the binned-SAH build on the device
    else:
        env.setup_acceleration_structure() if hasattr(env, "setup_acceleration_structure") else env.build()
    if texture_shadows:
        env.set_light_trace_flags(2)                                  # VRAD_TL_TEXTURE_SHADOWS (-textureshadows, testline.go:14)
    cluster = patch_clusters(env, prep)
    env.patches_upload(t["origin"], t["normal"], t["plane_dist"], t["area"], prep["refl"], cluster, prep["flags"])
    env.set_hierarchy(t["parent"], t["child1"], t["child2"], t["face"])
    if poly_form_factor:                                              # MakeTransfer: polygon-to-differential form factor for near pairs
        env.set_windings(t["wind_first"], t["wind_count"], t["wind_points"])
    if prep["needs_bump"].any() and world == 1:                       # TotalLight.Light[1..3] of the bump-mapped leaf patches (single GPU)
        env.set_bump(prep["needs_bump"], prep["bump_basis"])
    nnz = env.build_transfers(prep["pvs"])
    if np.any(prep["lights"]["type"] == 5):                            # EMIT_SKYAMBIENT samples the sky along vmath.Anorms
        env.set_sky_dirs(anorms())
    lifted = (t["origin"] + t["normal"]).astype(np.float32)            # one unit off the surface, like the luxel samples

    # DirectLight.PVS (common/types/light.go; AllocDLight + SetDLightVis, rad/lightmap/lights.go:118-161; sky lights: MergeDLightVis,
    # lightmap.go:399-411): a light only reaches the samples whose cluster is in the PVS of the light's own cluster (PVSCheck,
    # lightmap.go:413-422: a sample with cluster -1 is lit by everything).  K3 takes one light list for all its luxels, so the samples
    # go to the device grouped by cluster, each group with the lights that pass.
    light_sees = None
    if use_light_pvs and prep["pvs"] is not None and len(prep["lights"]):
        (env.bsp_upload if hasattr(env, "bsp_upload") else env.bsp_set)(prep["bsp"])
        nc = prep["pvs"].shape[0]
        lcl = np.asarray(env.cluster_from_point(np.ascontiguousarray(prep["lights"]["origin"], np.float32)))
        light_sees = np.ones((len(prep["lights"]), nc), bool)
        for k, c in enumerate(lcl):
            if prep["lights"]["type"][k] in (3, 5):                    # sky light / sky ambient: the merged PVS of the sky leafs
                if prep["sky_pvs"] is not None:
                    light_sees[k] = np.unpackbits(prep["sky_pvs"], bitorder="little")[:nc].astype(bool)
            elif 0 <= c < nc:
                light_sees[k] = prep["pvs"][c] != 0

    def lit_block(pos, nrm):
        if light_sees is None or pos.shape[0] == 0:
            return np.asarray(env.direct_light(pos, nrm, prep["lights"])) if pos.shape[0] else np.zeros((0, 3), np.float32)
        cl = np.asarray(env.cluster_from_point(pos))
        out = np.zeros((pos.shape[0], 3), np.float32)
        # one K3 call per distinct light subset: clusters that see the same lights go to the device together
        nc = light_sees.shape[1]
        masks = np.concatenate([light_sees.T, np.ones((1, light_sees.shape[0]), bool)])      # row nc = "sees every light" (cluster -1)
        row = np.where((cl >= 0) & (cl < nc), cl, nc)
        uniq, inverse = np.unique(masks, axis=0, return_inverse=True)
        group = inverse.reshape(-1)[row]
        for g in np.unique(group):
            keep = uniq[g]
            if keep.any():
                sel = np.nonzero(group == g)[0]
                out[sel] = np.asarray(env.direct_light(np.ascontiguousarray(pos[sel]), np.ascontiguousarray(nrm[sel]), prep["lights"][keep]))
        return out

    def lit_points(pos, nrm):
        if world == 1:
            return lit_block(pos, nrm)
        parts = range_partition(pos.shape[0], world)
        a, b = parts[rank]
        return all_gather_blocks(lit_block(pos[a:b], nrm[a:b]), parts, rank, device)
    direct = lit_points(prep["lux_pos"], prep["lux_normal"])
    emit0 = lit_points(lifted, t["normal"])
    total, _, done = env.bounce(emit0, bounces, early_out)
    bump = np.asarray(env.bump_totals()) if (prep["needs_bump"].any() and world == 1) else None
    return dict(nnz=int(nnz), direct=direct, emit0=emit0, total=np.asarray(total), bump_totals=bump, bounces_done=int(done))


def finish(env, prep: dict, lit: dict, rank: int = 0, world: int = 1, device=None, indirect: str = "radial") -> tuple[bytes, np.ndarray]:
    """K5 on the device + the lighting lump: returns (lump bytes, packed luxel colours).  indirect = "radial": the bounced light of a
    luxel is upstream's radial filter over the patch lights of its face and the face's smoothing neighbours (vrad_luxel_radial_light);
    "nearest": the light of the nearest leaf patch of the face (vrad_lightmap_finalize_patches).  world > 1: each rank packs its range."""
    from .sharding import range_partition
    parts = range_partition(lit["direct"].shape[0], world)
    a, b = parts[rank]
    if b > a and indirect == "radial":
        ind = B.luxel_radial_light(env, prep["lux_face"][a:b], prep["luxel_first"] - a, prep["lm_size"], prep["radial_first"], prep["radial_entries"], lit["total"],
                                   lit.get("bump_totals"))    # the three extra blocks of a bump-mapped face take TotalLight.Light[1..3]
        mine = B.lightmap_finalize(env, lit["direct"][a:b], ind)
    elif b > a:
        mine = B.lightmap_finalize_patches(env, lit["direct"][a:b], prep["lux_patch"][a:b], lit["total"])
    else:
        mine = np.zeros(0, B.RGBEXP32)
    colors = mine if world == 1 else all_gather_blocks(mine.view(np.int32), parts, rank, device).view(B.RGBEXP32)
    return B.pack_lighting(prep["lumps"], prep["luxel_first"], colors, prep["lump_bytes"]), colors


def bake_file(path_in: str, path_out: str, device: int = 0, bounces: int = 8, rank: int = 0, world: int = 1, comm_id: bytes | None = None,
              lights_rad_path: str | None = None, hdr: bool = False, luxel_density: float = 1.0, smooth_degrees: float = 45.0, chop: float = 4.0,
              max_chop: float = 4.0, fast: bool = False, texture_shadows: bool = False) -> dict:
    """The whole job: read a .bsp, light it on the GPU, write it back with LUMP_LIGHTING and the face lump replaced.
    world > 1: one process per GPU under torch.distributed (nccl); comm_id = the 128 bytes of Environment.comm_unique_id() from rank 0;
    every rank reads the file and lights its share, rank 0 writes the result.
    The keyword arguments are the command-line switches of the reference that reach this path (cmd/args.go:53-93): -bounce, -lights, -hdr,
    -luxeldensity, -smooth (degrees), -chop / -maxchop, -fast (here: the binned kd build), -textureshadows."""
    from .environment import Environment
    f = B.BspFile(path_in)
    try:
        face_lump, lighting_lump = f.set_target_faces(hdr)            # cache.SetTargetFaces (loadbsp/main.go:79-89)
        L = f.lumps()
        text = f.get(B.LUMP["ENTITIES"])[0].rstrip(b"\0").decode("utf-8", "replace")
        rad = open(lights_rad_path, "r", errors="replace").read() if lights_rad_path else None
        strings = (np.frombuffer(f.get(B.LUMP["TEXDATA_STRING_TABLE"])[0], "<i4"), f.get(B.LUMP["TEXDATA_STRING_DATA"])[0])
        import os
        import math
        prep = prepare(L, text, min_chop=chop, max_chop=max_chop, lights_rad=rad, lights_rad_hdr=hdr, texdata_strings=strings,
                       map_name=os.path.splitext(os.path.basename(path_in))[0], luxel_density=luxel_density,
                       smoothing_threshold=0.7071067 if smooth_degrees == 45.0 else float(np.float32(math.cos(math.radians(smooth_degrees)))))
        # (45 degrees is the reference's default, spelled float32(0.7071067) at rad/lightmap/lightmap.go:29 -- one ulp below cos(45 deg))
        env = Environment(device, rank, world)
        try:
            dev = None
            if world > 1:
                import torch
                dev = torch.device("cuda", device)
                env.comm_init(comm_id)
            lit = light(env, prep, bounces, rank=rank, world=world, device=dev, fast_tree=fast, texture_shadows=texture_shadows)
            lump, colors = finish(env, prep, lit, rank=rank, world=world, device=dev)
        finally:
            env.close()
        if rank == 0:
            f.set(lighting_lump, lump, version=1)
            f.set(face_lump, prep["lumps"].faces, version=1)
            if luxel_density < 1.0:                                   # the rescaled lightmap vectors go back into the file (rad/start.go:48-50)
                f.set(B.LUMP["TEXINFO"], prep["lumps"].texinfo)
            # lightmap.SaveVertexNormals (rad/start.go:82-85): "store the vertex normals calculated in PairEdges so that they can be written
            # to the bsp file for use in the engine"
            normals, indices = B.save_vertex_normals(prep["vertex_normals"])
            f.set(B.LUMP["VERTNORMALS"], normals); f.set(B.LUMP["VERTNORMALINDICES"], indices)
            f.save(path_out)
    finally:
        f.close()
    return dict(prep=prep, lit=lit, lump=lump, colors=colors)
