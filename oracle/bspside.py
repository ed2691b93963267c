"""CPU ORACLE, BSP side of the path.  TEST INFRASTRUCTURE ONLY (see oracle/oracle.h): only tests/ may import this.

A function-by-function restatement, in plain Python with numpy float32 scalars (IEEE binary32 per operation, no
contraction), of the reference code that turns BSP lumps into the hot path's inputs.  Written the way the reference is
written (recursion, per-face Python lists, heap windings) -- deliberately NOT the flat-array formulation of
vrad_b200/csrc/bsp_input.cpp / bsp_light.cpp, so that agreement between the two is evidence.  PARITY UNPINNED: the
reference ships no vectors for these functions and cannot be executed here (no Go toolchain); the SURVEY App. A
"intent" corrections are applied and listed next to each function.

Inputs are the numpy structured arrays of vrad_b200.bspfile (same field names as include/vrad_bsp.h); small maps only.
"""
from __future__ import annotations

import math

import numpy as np

F = np.float32
SURF_LIGHT, SURF_SKY2D, SURF_SKY, SURF_NOLIGHT, SURF_BUMPLIGHT, SURF_NOCHOP = 0x1, 0x2, 0x4, 0x400, 0x800, 0x4000
MASK_OPAQUE = 0x1 | 0x4000 | 0x80
LEAF_SKY, LEAF_RADIAL, LEAF_SKY2D = 1, 2, 4
TRACE_ID_SKY, TRACE_ID_OPAQUE = 0x01000000, 0x02000000
SIDE_FRONT, SIDE_BACK, SIDE_ON = 0, 1, 2


# ---- mgl32.Vec3 ------------------------------------------------------------------------------------------------------
def vec(p):
    return [F(p[0]), F(p[1]), F(p[2])]


def vsub(a, b):
    return [a[0] - b[0], a[1] - b[1], a[2] - b[2]]


def vadd(a, b):
    return [a[0] + b[0], a[1] + b[1], a[2] + b[2]]


def vscale(a, s):
    return [a[0] * s, a[1] * s, a[2] * s]


def vdot(a, b):
    return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]


def vcross(a, b):
    return [a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]]


def vlen(a):
    return F(math.sqrt(float(a[0] * a[0] + a[1] * a[1] + a[2] * a[2])))


def vnormalize(a):
    with np.errstate(divide="ignore", invalid="ignore"):
        l = F(1.0) / vlen(a)
        return [a[0] * l, a[1] * l, a[2] * l]


# ---- vmath/polygon/winding.go ------------------------------------------------------------------------------------------
def base_winding_for_plane(normal, dist):
    """winding.go:25-83.  Intent: |normal[i]| per axis (:36 reads normal[1]); case 0 and 1 share vup = z (:46-51, Go has no
    fallthrough); the fourth point is written to Points[3] (:78)."""
    best, x = F(-1.0), -1
    for i in range(3):
        v = abs(normal[i])
        if v > best:
            x, best = i, v
    vup = [F(0), F(0), F(0)]
    if x in (0, 1):
        vup[2] = F(1)
    else:
        vup[0] = F(1)
    v = vdot(vup, normal)
    vup = [vup[i] + (-v) * normal[i] for i in range(3)]
    vup = vnormalize(vup)
    org = vscale(normal, dist)
    vright = vcross(vup, normal)
    big = F(16384 * 4)
    vup, vright = vscale(vup, big), vscale(vright, big)
    return [vadd(vsub(org, vright), vup), vadd(vadd(org, vright), vup), vsub(vadd(org, vright), vup), vsub(vsub(org, vright), vup)]


def chop_winding_in_place(w, normal, dist, epsilon):
    """winding.go:85-171; returns the front part or None.  Intent: the inner loop advances j (:149 advances i)."""
    n = len(w)
    dists, sides, counts = [], [], [0, 0, 0]
    for p in w:
        d = vdot(p, normal) - dist
        dists.append(d)
        s = SIDE_FRONT if d > epsilon else (SIDE_BACK if d < -epsilon else SIDE_ON)
        sides.append(s); counts[s] += 1
    sides.append(sides[0]); dists.append(dists[0])
    if counts[SIDE_FRONT] == 0:
        return None
    if counts[SIDE_BACK] == 0:
        return w
    f = []
    for i in range(n):
        p1 = w[i]
        if sides[i] == SIDE_ON:
            f.append(p1); continue
        if sides[i] == SIDE_FRONT:
            f.append(p1)
        if sides[i + 1] == SIDE_ON or sides[i + 1] == sides[i]:
            continue
        p2 = w[(i + 1) % n]
        t = dists[i] / (dists[i] - dists[i + 1])
        mid = [F(0)] * 3
        for j in range(3):
            if normal[j] == 1:
                mid[j] = dist
            elif normal[j] == -1:
                mid[j] = -dist
            else:
                mid[j] = p1[j] + t * (p2[j] - p1[j])
        f.append(mid)
    return f


def winding_area(w):
    """winding.go:217-229"""
    total = F(0)
    for i in range(2, len(w)):
        total = total + vlen(vcross(vsub(w[i - 1], w[0]), vsub(w[i], w[0])))
    return total * F(0.5)


# ---- cmd/tasks/loadbsp ---------------------------------------------------------------------------------------------------
def get_brush_recursive(L, node, out):
    """brush.go:7-36"""
    if node < 0:
        leaf = L.leafs[-1 - node]
        for i in range(int(leaf["numleafbrushes"])):
            b = int(L.leafbrushes[int(leaf["firstleafbrush"]) + i])
            if b not in out:
                out.append(b)
    else:
        get_brush_recursive(L, int(L.nodes[node]["children"][0]), out)
        get_brush_recursive(L, int(L.nodes[node]["children"][1]), out)


def matrix_org_angles(origin, angles):
    """mat4.go:10-49: rows 0..2 of the matrix"""
    pi32 = F(math.pi)
    pitch, yaw, roll = (float(F(a) * pi32 / F(180.0)) for a in angles)
    sy, cy, sp, cp, sr, cr = math.sin(yaw), math.cos(yaw), math.sin(pitch), math.cos(pitch), math.sin(roll), math.cos(roll)
    return [[F(cp * cy), F(sr * sp * cy + cr * -sy), F(cr * sp * cy + -sr * -sy), F(origin[0])],
            [F(cp * sy), F(sr * sp * sy + cr * cy), F(cr * sp * sy + -sr * cy), F(origin[1])],
            [F(-sp), F(sr * cp), F(cr * cp), F(origin[2])]]


def mul4x3(m, p):
    """mat4.go:63-69"""
    if m is None:
        return p
    return [m[r][0] * p[0] + m[r][1] * p[1] + m[r][2] * p[2] + m[r][3] for r in range(3)]


def add_brush_to_raytrace_environment(L, brush, xform, ids, tris):
    """main.go:237-277.  Intent (#15): sky and displacement sides are the ones skipped."""
    if not (int(brush["contents"]) & MASK_OPAQUE):
        return
    first, n = int(brush["firstside"]), int(brush["numsides"])
    for i in range(n):
        side = L.brushsides[first + i]
        plane = L.planes[int(side["planenum"])]
        w = base_winding_for_plane(vec(plane["normal"]), F(plane["dist"]))
        tflags = int(L.texinfo[int(side["texinfo"])]["flags"]) if int(side["texinfo"]) >= 0 else 0
        if (tflags & SURF_SKY) or int(side["dispinfo"]) != 0:
            continue
        for j in range(n):
            if w is None:
                break
            if i == j:
                continue
            other = L.brushsides[first + j]
            if int(other["bevel"]) != 0:
                continue
            op = L.planes[int(other["planenum"]) ^ 1]
            w = chop_winding_in_place(w, vec(op["normal"]), F(op["dist"]), F(0))
        if w is not None:
            for j in range(2, len(w)):
                ids.append(TRACE_ID_OPAQUE)
                tris.append([mul4x3(xform, w[0]), mul4x3(xform, w[j - 1]), mul4x3(xform, w[j])])


def edge_vertex(L, f, k):
    """lightmap.EdgeVertex, lightmap.go:267-282"""
    n = int(f["numedges"])
    if k < 0:
        k += n
    elif k >= n:
        k %= n
    se = int(L.surfedges[int(f["firstedge"]) + k])
    return int(L.edges[-se]["v"][1]) if se < 0 else int(L.edges[se]["v"][0])


def raytrace_triangles(L, casters=()):
    """ExtractBrushEntityShadowCasters (main.go:186-211) then addBrushesForRayTrace (:279-340).
    casters: (model index, origin, angles) per entity with vrad_brush_cast_shadows."""
    ids, tris = [], []
    for (model, origin, angles) in casters:
        if model <= 0 or model >= len(L.models):
            continue
        lst = []
        get_brush_recursive(L, int(L.models[model]["headnode"]), lst)
        m = matrix_org_angles(origin, angles)
        for b in lst:
            add_brush_to_raytrace_environment(L, L.brushes[b], m, ids, tris)
    if len(L.models):
        lst = []
        get_brush_recursive(L, int(L.models[0]["headnode"]), lst)
        for b in lst:
            add_brush_to_raytrace_environment(L, L.brushes[b], None, ids, tris)
        world = L.models[0]
        for i in range(int(world["numfaces"])):
            f = L.faces[int(world["firstface"]) + i]
            if not (int(L.texinfo[int(f["texinfo"])]["flags"]) & SURF_SKY):
                continue
            pts = [vec(L.vertexes3[edge_vertex(L, f, j)]) for j in range(int(f["numedges"]))]
            for j in range(2, len(pts)):
                ids.append(TRACE_ID_SKY); tris.append([pts[0], pts[j - 1], pts[j]])
    return np.asarray(ids, np.int32), np.asarray(tris, np.float32).reshape(-1, 3, 3)


# ---- rad/world, rad/patches ----------------------------------------------------------------------------------------------
def winding_from_face(L, f, origin):
    """world.WindingFromFace + RemoveColinearPoints (face.go:92-114, point.go:12-46)"""
    w = [vadd(vec(L.vertexes3[edge_vertex(L, f, i)]), origin) for i in range(int(f["numedges"]))]
    n = len(w)
    kept = []
    for i in range(n):
        j, k = (i + 1) % n, (i + n - 1) % n
        v1, v2 = vnormalize(vsub(w[j], w[i])), vnormalize(vsub(w[i], w[k]))
        if vdot(v1, v2) < F(0.999):
            kept.append(w[i])
    return kept if len(kept) != n else w


def face_patches(L, model_origins=None, max_chop=4.0):
    """patches.MakePatches (build.go:21-65) + the per-face part of MakePatchForFace (face.go:29-197), BaseLightForFace
    (:208-230) and PreventSubdivision (subdivide.go:151-165)."""
    out = dict(windings=[], normal=[], plane_dist=[], lux_scale=[], chop=[], sky=[], no_subdivide=[], face_number=[],
               reflectivity=[], base_area=[], needs_bump=[], scale=[])
    for m in range(len(L.models)):
        mod = L.models[m]
        origin = vec(model_origins[m]) if model_origins is not None else vec((0, 0, 0))
        for j in range(int(mod["numfaces"])):
            fn = int(mod["firstface"]) + j
            f = L.faces[fn]
            if int(f["dispinfo"]) != -1:
                continue
            tx = L.texinfo[int(f["texinfo"])]
            td = L.texdata[int(tx["texdata"])]
            plane = L.planes[int(f["planenum"])]
            normal, dist = vec(plane["normal"]), F(plane["dist"])
            if origin[0] != 0 or origin[1] != 0 or origin[2] != 0:
                dist = dist + vdot(origin, normal)
            scale, chop_scale = [F(0), F(0)], [F(0), F(0)]
            for i in range(2):
                for k in range(3):
                    scale[i] = scale[i] + F(tx["texture_vecs"][i][k]) * F(tx["texture_vecs"][i][k])
                    chop_scale[i] = chop_scale[i] + F(tx["lightmap_vecs"][i][k]) * F(tx["lightmap_vecs"][i][k])
                scale[i] = F(math.sqrt(float(scale[i]))); chop_scale[i] = F(math.sqrt(float(chop_scale[i])))
            fl = int(tx["flags"])
            out["windings"].append(winding_from_face(L, f, origin))
            out["normal"].append(normal); out["plane_dist"].append(dist)
            out["lux_scale"].append((chop_scale[0] + chop_scale[1]) / F(2)); out["chop"].append(F(max_chop))
            out["sky"].append(1 if fl & SURF_SKY else 0)
            out["no_subdivide"].append(1 if (fl & SURF_NOCHOP) or ((fl & SURF_NOLIGHT) and not (fl & SURF_LIGHT)) else 0)
            out["face_number"].append(fn)
            out["reflectivity"].append([min(F(td["reflectivity"][k]) * F(1.0), F(0.99)) for k in range(3)])
            prod = (int(td["height"]) * int(td["width"]) + 2 ** 31) % 2 ** 32 - 2 ** 31       # Go: int32 product, wrapping
            out["base_area"].append(F(prod))
            out["needs_bump"].append(1 if fl & SURF_BUMPLIGHT else 0)
            out["scale"].append(scale)
    return out


def calc_face_extents(L, f):
    """world.CalcFaceExtents (face.go:14-90): (mins[2], size[2])"""
    tx = L.texinfo[int(f["texinfo"])]
    mins, maxs = [F(1e24), F(1e24)], [F(-1e24), F(-1e24)]
    for i in range(int(f["numedges"])):
        v = vec(L.vertexes3[edge_vertex(L, f, i)])
        for j in range(2):
            lv = tx["lightmap_vecs"][j]
            val = v[0] * F(lv[0]) + v[1] * F(lv[1]) + v[2] * F(lv[2]) + F(lv[3])
            if val < mins[j]:
                mins[j] = val
            if val > maxs[j]:
                maxs[j] = val
    lo = [F(math.floor(float(mins[i]))) for i in range(2)]
    hi = [F(math.ceil(float(maxs[i]))) for i in range(2)]
    return [int(lo[0]), int(lo[1])], [int(hi[0] - lo[0]), int(hi[1] - lo[1])]


def face_extents(L):
    """rad.UpdateAllFaceLightmapExtents (start.go:100-111): unlit faces keep the stored extents"""
    mins, size = [], []
    for f in L.faces:
        if int(L.texinfo[int(f["texinfo"])]["flags"]) & (SURF_SKY | SURF_NOLIGHT):
            mins.append([int(f["lm_mins"][0]), int(f["lm_mins"][1])]); size.append([int(f["lm_size"][0]), int(f["lm_size"][1])])
        else:
            a, b = calc_face_extents(L, f)
            mins.append(a); size.append(b)
    return np.asarray(mins, np.int32).reshape(-1, 2), np.asarray(size, np.int32).reshape(-1, 2)


def rescale_lightmap_vecs(texinfo, luxel_density):
    """rad.Start (start.go:21-64).  Intent: scale = VectorNormalize(tmp) (:43-45 computes the length of the normalised vector)."""
    t = texinfo.copy()
    if not (luxel_density < 1.0):
        return t
    for i in range(len(t)):
        for j in range(2):
            tmp = vec(t[i]["lightmap_vecs"][j][:3])
            scale = vlen(tmp)
            if scale == 0:
                continue
            tmp = vscale(tmp, F(1.0) / scale)
            if abs(scale) > F(luxel_density):
                scale = F(-luxel_density) if scale < 0 else F(luxel_density)
                tmp = vscale(tmp, scale)
                for k in range(3):
                    t[i]["lightmap_vecs"][j][k] = tmp[k]
    return t


# ---- rad/clustertable ------------------------------------------------------------------------------------------------------
def make_parents(L):
    """clustertable.MakeParents(0, -1), nodes.go:21-36"""
    node_parents = [-1] * len(L.nodes); leaf_parents = [-1] * len(L.leafs)

    def rec(n, parent):
        node_parents[n] = parent
        for i in range(2):
            j = int(L.nodes[n]["children"][i])
            if j < 0:
                leaf_parents[-j - 1] = n
            else:
                rec(j, n)
    if len(L.nodes):
        rec(0, -1)
    return np.asarray(node_parents, np.int32), np.asarray(leaf_parents, np.int32)


def build_cluster_table(L, n_clusters):
    """clustertable.BuildClusterTable, build.go:9-31"""
    return [[j for j in range(len(L.leafs)) if int(L.leafs[j]["cluster"]) == i] for i in range(n_clusters)]


# ---- rad/lightmap: vis -----------------------------------------------------------------------------------------------------
def decompress_vis(data, row):
    """lightmap.DecompressVis, vis.go:54-94 (App. A #23: the standard run-length code)"""
    out, i = bytearray(), 0
    while len(out) < row:
        if data[i]:
            out.append(data[i]); i += 1
            continue
        c = data[i + 1]
        i += 2
        c = min(c, row - len(out))
        out += bytes(c)
    return bytes(out)


def get_vis_cache(L, cluster):
    """lightmap.GetVisCache, vis.go:9-47"""
    vis = L.visdata.tobytes()
    nc = int(np.frombuffer(vis[:4], "<i4")[0]) if len(vis) >= 4 else 0
    row = (nc + 7) >> 3
    if nc == 0 or cluster < 0:
        return bytes([255]) * max(row, 1)
    ofs = int(np.frombuffer(vis[4 + 8 * cluster:8 + 8 * cluster], "<i4")[0])
    return decompress_vis(vis[ofs:], row)


def pvs_check(pvs, cluster):
    """lightmap.PVSCheck, lightmap.go:413-422"""
    return (pvs[cluster >> 3] & (1 << (cluster & 7))) != 0 if cluster >= 0 else True


def build_vis_for_light_environment(L, can_leaf_trace_to_sky=None):
    """lightmap.BuildVisForLightEnvironment + MergeDLightVis (lightmap.go:284-411).  Returns (leaf flags, merged PVS or None).
    Intent: a leaf is marked when the sky leaf IS in its PVS (:337 continues when it is)."""
    nl = len(L.leafs)
    flags = [((int(L.leafs[i]["area_flags"]) & 0xffff) >> 9) & 0x7f for i in range(nl)]
    nc = L.n_clusters
    merged = None
    for i in range(nl):
        flags[i] &= ~(LEAF_SKY | LEAF_SKY2D)
        lf = L.leafs[i]
        for k in range(int(lf["numleaffaces"])):
            f = L.faces[int(L.leaffaces[int(lf["firstleafface"]) + k])]
            tf = int(L.texinfo[int(f["texinfo"])]["flags"])
            if tf & SURF_SKY:
                flags[i] |= LEAF_SKY2D if tf & SURF_SKY2D else LEAF_SKY
                pvs = get_vis_cache(L, int(lf["cluster"]))
                merged = bytearray(pvs) if merged is None else bytearray(a | b for a, b in zip(merged, pvs))
                break
    bits3d, bits2d = [False] * nl, [False] * nl
    for i in range(nl):
        if flags[i] & LEAF_SKY:
            continue
        if int(L.leafs[i]["contents"]) & 1:
            continue
        pvs = get_vis_cache(L, int(L.leafs[i]["cluster"]))
        for j in range(nl):
            if j == i:
                continue
            if not (flags[j] & (LEAF_SKY | LEAF_SKY2D)):
                continue
            if nc and not pvs_check(pvs, int(L.leafs[j]["cluster"])):
                continue
            if flags[j] & LEAF_SKY2D:
                bits2d[i] = True
            if flags[j] & LEAF_SKY:
                bits3d[i] = True
                break
    for i in range(nl):
        if flags[i] & LEAF_SKY:
            continue
        if int(L.leafs[i]["contents"]) & 1:
            continue
        if bits2d[i]:
            flags[i] |= LEAF_SKY2D
        if bits3d[i]:
            flags[i] |= LEAF_SKY
            flags[i] &= ~LEAF_SKY2D
        elif flags[i] & LEAF_RADIAL:
            if can_leaf_trace_to_sky is None:
                raise RuntimeError("radial leaf needs CanLeafTraceToSky")
            if can_leaf_trace_to_sky(i):
                flags[i] |= LEAF_SKY
    return np.asarray(flags, np.uint8), (bytes(merged) if merged is not None and nc else None)


# ---- rad/lightmap: smoothing normals -------------------------------------------------------------------------------------
def valid_disp_face(f):
    """polygon.ValidDispFace, vmath/polygon/face.go:5-17"""
    return int(f["dispinfo"]) != -1 and int(f["numedges"]) == 4


def pair_edges(L, smoothing_threshold=0.7071067):
    """lightmap.PairEdges, lightmap.go:37-216.  Intent: the face's own edges are walked (:74 walks len(Edges)).
    Returns (per-face list of per-vertex normals, per-face neighbour lists)."""
    thr = F(smoothing_threshold)
    nf = len(L.faces)
    vertex_face = {}
    for i in range(nf):
        f = L.faces[i]
        for j in range(int(f["numedges"])):
            lst = vertex_face.setdefault(edge_vertex(L, f, j), [])
            if i not in lst:
                lst.append(i)
    face_normal = [vec(L.planes[int(L.faces[i]["planenum"])]["normal"]) for i in range(nf)]
    has_disp = [valid_disp_face(L.faces[i]) for i in range(nf)]
    normals, neighbours = [], []
    for i in range(nf):
        f = L.faces[i]
        nbs = []
        fn_normal = [[F(0), F(0), F(0)] for _ in range(int(f["numedges"]))]
        for j in range(int(f["numedges"])):
            n = edge_vertex(L, f, j)
            for o in vertex_face[n]:
                if o == i:
                    continue
                if not has_disp[i] and has_disp[o]:
                    continue
                nb = face_normal[o]
                cos_angle = vdot(nb, face_normal[i])
                if has_disp[i]:
                    fn_normal[j] = vadd(fn_normal[j], nb)
                elif int(f["smoothing_groups"]) == 0 and int(L.faces[o]["smoothing_groups"]) == 0:
                    if cos_angle >= thr:
                        fn_normal[j] = vadd(fn_normal[j], nb)
                    else:
                        continue
                else:
                    g = int(f["smoothing_groups"]) & int(L.faces[o]["smoothing_groups"])
                    if g & 0xff000000:
                        continue
                    if g != 0:
                        fn_normal[j] = vadd(fn_normal[j], nb)
                    else:
                        continue
                if o not in nbs:
                    nbs.append(o)
                    assert len(nbs) <= 64, "Stack overflow in neighbors"
        for j in range(int(f["numedges"])):
            fn_normal[j] = vnormalize(vadd(fn_normal[j], face_normal[i]))
        normals.append(fn_normal); neighbours.append(nbs)
    return normals, neighbours


def save_vertex_normals(normals):
    """lightmap.SaveVertexNormals + NormalList.FindOrAddNormal (lightmap.go:218-265, normallist.go:12-49; App. A #24)"""
    grid, lst, indices = {}, [], []
    for face_normals in normals:
        for v in face_normals:
            gi = []
            for d in range(3):
                g = int((v[d] + F(1.0)) * F(0.5) * F(8) - F(0.000001))
                gi.append(max(min(g, 7), 0))
            cell = grid.setdefault(tuple(gi), [])
            found = -1
            for idx in cell:
                d = vsub(lst[idx], v)
                if vdot(d, d) < F(0.00001):
                    found = idx
                    break
            if found < 0:
                found = len(lst); cell.append(found); lst.append(v)
            indices.append(found)
    return np.asarray(lst, np.float32).reshape(-1, 3), np.asarray(indices, np.uint16)


def get_phong_normal(L, normals, centroids, face_num, spot, smoothing_threshold=0.7071067):
    """lightmap.GetPhongNormal, normallist.go:52-143"""
    f = L.faces[face_num]
    face_normal = vec(L.planes[int(f["planenum"])]["normal"])
    phong = face_normal
    if F(smoothing_threshold) != 1:
        fn = normals[face_num]
        ne = int(f["numedges"])
        centre = vec(centroids[face_num])
        for j in range(ne):
            n1, n2 = fn[j], fn[(j + 1) % ne]
            p1, p2 = vec(L.vertexes3[edge_vertex(L, f, j)]), vec(L.vertexes3[edge_vertex(L, f, j + 1)])
            v1, v2, vspot = vsub(p1, centre), vsub(p2, centre), vsub(vec(spot), centre)
            aa, bb, ab = vdot(v1, v1), vdot(v2, v2), vdot(v1, v2)
            with np.errstate(divide="ignore", invalid="ignore"):
                a1 = (bb * vdot(v1, vspot) - ab * vdot(vspot, v2)) / (aa * bb - ab * ab)
                a2 = (vdot(vspot, v2) - a1 * ab) / bb
            if a1 >= 0 and a2 >= 0:
                scale = F(1.0) - a1 - a2
                phong = vscale(face_normal, scale)
                phong = vadd(phong, vscale(n1, a1))
                phong = vadd(phong, vscale(n2, a2))
                assert vlen(phong) >= F(1.0e-20), "Phong normal length out of bounds"
                return vnormalize(phong)
    return phong


# ---- luxels + lighting lump (UNCITED upstream; see include/vrad_bsp.h) -----------------------------------------------------
def pack_rgbexp32(rgb):
    """upstream VectorToColorRGBExp32 (halve / double the largest component into [128,255]); own rule below 2^-120 -> 0"""
    r, g, b = (F(c) if c > 0 else F(0) for c in rgb)
    mx = max(r, g, b)
    if not (mx >= F(math.ldexp(1.0, -120))):
        return (0, 0, 0, 0)
    mx = min(mx, F(3.0e38))
    power, x = 0, mx
    while x > 255:
        power += 1; x = x * F(0.5)
    while x < 128:
        power -= 1; x = x * F(2.0)
    scalar = F(math.ldexp(1.0, -power))
    out = [int(min(c * scalar, F(255.0))) for c in (r, g, b)]
    return (out[0], out[1], out[2], power)


def face_luxels(L, mins, size, face_origins=None):
    """upstream InitLightinfo + CalcPoints without sample nudging: per lit face the (w+1)(h+1) sample points, one unit off the face;
    bump-mapped faces repeat the block for the three bump-basis normals (normals passed in by the caller are not modelled here --
    this returns positions, flat normals and the face of each luxel for the flat block only)."""
    pos, nrm, faces = [], [], []
    for i in range(len(L.faces)):
        f = L.faces[i]
        tx = L.texinfo[int(f["texinfo"])]
        if int(tx["flags"]) & (SURF_SKY | SURF_NOLIGHT):
            continue
        plane = L.planes[int(f["planenum"])]
        n, dist = vec(plane["normal"]), F(plane["dist"])
        lv0, lv1 = vec(tx["lightmap_vecs"][0][:3]), vec(tx["lightmap_vecs"][1][:3])
        texnormal = vnormalize(vcross(lv1, lv0))
        distscale = vdot(texnormal, n)
        if distscale < 0:
            distscale = -distscale; texnormal = [-c for c in texnormal]
        distscale = F(1.0) / distscale
        l2w0 = vcross(lv1, n); l2w0 = vscale(l2w0, F(1.0) / vdot(l2w0, lv0))
        l2w1 = vcross(lv0, n); l2w1 = vscale(l2w1, F(1.0) / vdot(l2w1, lv1))
        org = [-(F(tx["lightmap_vecs"][0][3]) * l2w0[k]) - F(tx["lightmap_vecs"][1][3]) * l2w1[k] for k in range(3)]
        d = (vdot(org, n) - dist) * distscale
        org = [org[k] + (-d) * texnormal[k] for k in range(3)]
        if face_origins is not None:
            org = vadd(org, vec(face_origins[i]))
        w, h = int(size[i][0]) + 1, int(size[i][1]) + 1
        for t in range(h):
            for s in range(w):
                us, ut = F(int(mins[i][0]) + s), F(int(mins[i][1]) + t)
                pos.append([org[k] + us * l2w0[k] + ut * l2w1[k] + n[k] for k in range(3)])
                nrm.append(n); faces.append(i)
    return np.asarray(pos, np.float32).reshape(-1, 3), np.asarray(nrm, np.float32).reshape(-1, 3), np.asarray(faces, np.int32)


# ---- common/parser/lights-rad: texture lights -----------------------------------------------------------------------------
def _scan_floats(s, limit=8):
    """fmt.Sscanf(light, "%e %e ...") into float32: leading numeric tokens, stopping at the first that does not parse"""
    out = []
    for tok in s.split()[:limit]:
        try:
            out.append(F(float(tok)))
        except ValueError:
            break
    return out


def rad_light_for_string(light):
    """lights_rad.lightForString, reader.go:122-186 (useHDR := true: the second 4-tuple wins when 8 numbers are given).  None = rejected."""
    v = _scan_floats(light)
    n = len(v)
    v = v + [F(0)] * (8 - n)
    r, g, b, scaler = v[0], v[1], v[2], v[3]
    if n == 8:
        r, g, b, scaler = v[4], v[5], v[6], v[7]
        n = 4
    if r < 0 or g < 0 or b < 0 or scaler < 0:
        return None
    lin = lambda c: F(math.pow(float(c / F(255.0)), 2.2) * 255)
    if n == 1:
        x = lin(r)
        return [x, x, x]
    if n in (3, 4):
        out = [lin(r), lin(g), lin(b)]
        if n == 4:
            out = [c * (scaler / F(255.0)) for c in out]
        return out
    return None


def read_lights_rad(text, hdr=False):
    """lights_rad.Reader.Read, reader.go:19-120 -> (list of (name, value), noshadow materials, forcetextureshadow models).
    Intent: lines end at newline, empty lines are skipped, "hdr:" / "ldr:" are prefixes."""
    table, noshadow, forced = [], [], []
    for line in text.split("\n"):
        line = line.strip(" \t\r")
        if not line:
            continue
        if line.startswith("hdr:"):
            if not hdr:
                continue
            line = line[4:]
        if line.startswith("ldr:"):
            if hdr:
                continue
            line = line[4:]
        parts = line.split(None, 1)
        if not parts:
            continue
        tok, rest = parts[0], (parts[1] if len(parts) > 1 else "")
        if tok == "noshadow":
            if rest.split():
                noshadow.append(rest.split()[0].split(".")[0])
        elif tok == "forcetextureshadow":
            if rest.split():
                name = rest.split()[0]
                if name.startswith("models/"):
                    name = name[len("models/"):]
                if name.endswith(".mdl"):
                    name = name[:-4]
                forced.append(name)
        else:
            value = rad_light_for_string(rest) if rest.strip() else None
            if value is None:
                continue                                          # "ignoring bad texlight"
            for entry in table:
                if entry[0] == tok:
                    entry[1] = value                              # overriding
                    break
            else:
                assert len(table) < 128, "Too many texlights"
                table.append([tok, value])
    return table, noshadow, forced


def light_for_texture(name, map_name, table):
    """patches.LightForTexture, rad/patches/face.go:232-280"""
    prefix = "maps/" + map_name + "/"
    if map_name and name.startswith(prefix):
        base = name[len(prefix):]
        found = True
        for _ in range(3):
            us = base.rfind("_")
            if us == -1:
                found = False
                break
            base = base[:us]
        if found:
            name = base
    for n, value in table:
        if n == name:
            return value
    return [F(0), F(0), F(0)]


# ---- bounced light per luxel: upstream radial.cpp (UNCITED, see include/vrad_bsp.h) ---------------------------------------
def world_to_luxel(tx, p, offset=None):
    """WorldToLuxelSpace"""
    q = vec(p) if offset is None else vsub(vec(p), vec(offset))
    return [q[0] * F(tx["lightmap_vecs"][k][0]) + q[1] * F(tx["lightmap_vecs"][k][1]) + q[2] * F(tx["lightmap_vecs"][k][2]) + F(tx["lightmap_vecs"][k][3]) for k in range(2)]


def build_patch_radial(L, f, mins, size, patch_lists, tree, totals, neighbours=None, face_origins=None):
    """BuildPatchRadial + AddBouncedToRadial for face f, the way upstream does it: every patch SCATTERS r * light into the luxels it
    reaches; then SampleRadial = light / weight per luxel.  patch_lists[face] = leaf patches of that face (ascending).  Returns
    [(size[1]+1) * (size[0]+1), 3] in luxel order (t major)."""
    tx = L.texinfo[int(L.faces[f]["texinfo"])]
    w, h = int(size[f][0]) + 1, int(size[f][1]) + 1
    light = [[F(0), F(0), F(0)] for _ in range(w * h)]
    weight = [F(0)] * (w * h)
    off = None if face_origins is None else face_origins[f]
    sources = [f] + ([] if neighbours is None else list(neighbours[f]))
    for src in sources:
        for p in patch_lists[src]:
            pts = tree["wind_points"][tree["wind_first"][p]:tree["wind_first"][p] + tree["wind_count"][p]]
            st = [world_to_luxel(tx, q, off) for q in pts]
            st = [[c[0] - F(int(mins[f][0])), c[1] - F(int(mins[f][1]))] for c in st]
            mn = [min(c[k] for c in st) for k in range(2)]; mx = [max(c[k] for c in st) for k in range(2)]
            c = world_to_luxel(tx, tree["origin"][p], off)
            c = [c[0] - F(int(mins[f][0])), c[1] - F(int(mins[f][1]))]
            dists, distt = max(F(1.0), mx[0] - mn[0]), max(F(1.0), mx[1] - mn[1])
            inv_s, inv_t = F(1.0) / dists, F(1.0) / distt
            for t in range(h):
                for s in range(w):
                    ds, dt = (c[0] - F(s)) * inv_s, (c[1] - F(t)) * inv_t
                    r = F(2.0) - (ds * ds + dt * dt)
                    if r > 0:
                        i = s + t * w
                        light[i] = [light[i][k] + F(totals[p][k]) * r for k in range(3)]
                        weight[i] = weight[i] + r
    out = np.zeros((w * h, 3), np.float32)
    for i in range(w * h):
        if weight[i] > F(0.00001):
            inv = F(1.0) / weight[i]
            out[i] = [light[i][k] * inv for k in range(3)]
    return out


# ---- samples inside the face (own rule, see include/vrad_bsp.h: vrad_bsp_place_samples) --------------------------------------
def place_sample(poly, s, t):
    """The sample of luxel (s, t): centroid of face-polygon ∩ cell, else the nearest outline point.  float64 geometry -- an independent
    computation, compared with a tolerance (the product works in float32)."""
    cell = [(s - 0.5, t - 0.5), (s + 0.5, t - 0.5), (s + 0.5, t + 0.5), (s - 0.5, t + 0.5)]
    pts = [(float(a), float(b)) for a, b in poly]

    def clip(points, axis, bound, keep_greater):
        out = []
        for i in range(len(points)):
            a, b = points[i], points[(i + 1) % len(points)]
            ina = a[axis] >= bound if keep_greater else a[axis] <= bound
            inb = b[axis] >= bound if keep_greater else b[axis] <= bound
            if ina:
                out.append(a)
            if ina != inb:
                u = (bound - a[axis]) / (b[axis] - a[axis])
                out.append((a[0] + u * (b[0] - a[0]), a[1] + u * (b[1] - a[1])))
        return out
    c = clip(clip(clip(clip(pts, 0, s - 0.5, True), 0, s + 0.5, False), 1, t - 0.5, True), 1, t + 0.5, False)
    if len(c) >= 3:
        a2 = cx = cy = 0.0
        for i in range(1, len(c) - 1):
            cr = (c[i][0] - c[0][0]) * (c[i + 1][1] - c[0][1]) - (c[i][1] - c[0][1]) * (c[i + 1][0] - c[0][0])
            a2 += cr; cx += cr * (c[0][0] + c[i][0] + c[i + 1][0]); cy += cr * (c[0][1] + c[i][1] + c[i + 1][1])
        if abs(a2) > 1e-6:
            return cx / (3 * a2), cy / (3 * a2), abs(a2) / 2
    best, best_d = pts[0], 1e30
    for i in range(len(pts)):
        a, b = pts[i], pts[(i + 1) % len(pts)]
        ex, ey = b[0] - a[0], b[1] - a[1]
        l2 = ex * ex + ey * ey
        u = ((s - a[0]) * ex + (t - a[1]) * ey) / l2 if l2 > 0 else 0.0
        u = min(max(u, 0.0), 1.0)
        x = (a[0] + u * ex, a[1] + u * ey)
        d = (x[0] - s) ** 2 + (x[1] - t) ** 2
        if d < best_d:
            best, best_d = x, d
    return best[0], best[1], 0.0
