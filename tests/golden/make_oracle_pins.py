"""Regenerates tests/golden/oracle_pins.json: SHA-256 digests (plus a few raw values) of the CPU oracle's outputs on
fixed seeded inputs.  They are NOT reference-derived golden vectors -- the reference ships none and cannot be run
("parity unpinned") -- they pin the oracle itself, so that a later edit of oracle/ or of the scene
generators cannot silently move the target the CUDA path is checked against.  Run from the repo root:
    python tests/golden/make_oracle_pins.py
"""
import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)


def digest(*arrays):
    h = hashlib.sha256()
    for a in arrays:
        h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()


def compute():
    from oracle import pyoracle
    from vrad_b200 import scenes
    pins = {}
    s1 = scenes.box_room()
    o1 = pyoracle.env_from_scene(s1)
    t = o1.export()
    pins["s1_kd_tree"] = digest(t["children"], t["split"], t["tri_index"], t["tris"])
    a, b = scenes.shadow_segments(s1, 8192, seed=0xC0FFEE)
    bits = o1.test_lines(a, b, threads=4)
    pins["s1_test_lines_8192"] = digest(bits)
    pins["s1_test_lines_first_words"] = [int(x) for x in bits[:4]]
    r = scenes.random_rays(s1, 4096, seed=0xBEEF)
    ht, hs, hd = o1.trace1(r["o"], r["d"], r["tmax"], threads=4)
    pins["s1_closest_hit_4096"] = digest(ht, hs, hd)
    sel = slice(0, None, 8)
    o1.patches_upload(s1.patch_origin[sel], s1.patch_normal[sel], s1.patch_plane_dist[sel], s1.patch_area[sel], s1.patch_refl[sel])
    pins["s1_transfers_every_8th_patch_nnz"] = int(o1.build_transfers(threads=4))
    pins["s1_transfers_every_8th_patch"] = digest(*o1.transfers())
    lum = o1.direct_light(s1.luxel_pos[::16], s1.luxel_normal[::16], s1.lights, threads=4)
    pins["s1_direct_light_every_16th_luxel"] = digest(lum)
    sk = scenes.sky_room(); m = sk.meta
    o = pyoracle.OracleEnv(); o.add_triangles(sk.tri_ids, sk.tri_verts, sk.tri_flags); o.build()
    o.set_triangle_colors(m["tri_colors"]); o.bsp_set(m["bsp"]); o.process_sky_cameras(m["cams_origin"], m["cams_scale"])
    sa, sb = scenes.sky_segments(sk, 8192, seed=0x5C1)
    for flags in (0, 1, 3, 7):
        pins[f"s4_sky_fraction_flags{flags}"] = digest(o.test_lines_sky(sa, sb, flags, 7, threads=4))
    pins["s4_point_leafnum"] = digest(o.point_leafnum(sa.T.copy()), o.cluster_from_point(sa.T.copy()))
    faces, pts, _ = scenes.room_faces(3, 2)
    tr = pyoracle.subdivide_patches(faces, pts)
    pins["subdivide_room_faces_3x2"] = digest(*[tr[k] for k in sorted(tr)])
    pins["subdivide_room_faces_3x2_patches"] = int(tr["parent"].shape[0])
    hs_ = scenes.multi_room_hier(nx=2, ny=1, boxes_per_room=6); ht_ = hs_.meta["tree"]
    oh = pyoracle.env_from_scene(hs_); oh.set_hierarchy(ht_["parent"], ht_["child1"], ht_["child2"], ht_["face"])
    pins["hier_2x1_nnz"] = int(oh.build_transfers(hs_.pvs, threads=4))
    pins["hier_2x1_transfers"] = digest(*oh.transfers())
    emit0 = scenes.SplitMix64(5).uniform(3 * hs_.n_patches, 0.0, 200.0).reshape(-1, 3)
    pins["hier_2x1_bounce4_single_thread"] = digest(oh.bounce(emit0, 4, threads=1)[0])
    row, used = pyoracle.decompress_vis(bytes([0xFF, 0x00, 0x03, 0x81]), 40)
    pins["decompress_vis_kat"] = [[int(x) for x in row], used]
    # BSP side (oracle/bspside.py on the synthetic BSP v20 map of vrad_b200/bspfile.py) and the binned kd builder's tree
    from oracle import bspside
    from vrad_b200 import bspfile
    from vrad_b200.environment import kd_build_binned_host
    L, meta = bspfile.synthetic_map(3, 2, boxes_per_room=5, sky_rooms=(1,), bump_rooms=(0,))
    e = meta["brush_entity"]
    pins["bsp_map_lumps"] = digest(*[L.a[k] for k in sorted(L.a)], L.visdata)
    ids, verts = bspside.raytrace_triangles(L, [(e["model"], e["origin"], e["angles"])])
    pins["bsp_raytrace_triangles"] = digest(ids, verts)
    fp = bspside.face_patches(L)
    pins["bsp_face_windings"] = digest(np.asarray([q for w in fp["windings"] for q in w], np.float32), np.asarray(fp["plane_dist"], np.float32))
    mins, size = bspside.face_extents(L)
    pins["bsp_face_extents"] = digest(mins, size)
    flags, pvs = bspside.build_vis_for_light_environment(L)
    pins["bsp_sky_vis"] = [int(x) for x in flags] + list(pvs)
    normals, nbs = bspside.pair_edges(L, -1.0)
    pins["bsp_pair_edges_all_smooth"] = digest(np.asarray([v for fn in normals for v in fn], np.float32), np.asarray([x for l in nbs for x in l], np.int32))
    un, ui = bspside.save_vertex_normals(normals)
    pins["bsp_vertex_normal_lumps"] = digest(un, ui)
    pos, nrm, lf = bspside.face_luxels(L, mins, size)
    pins["bsp_luxels"] = digest(pos, nrm, lf)
    pins["rgbexp32_kat"] = [list(bspside.pack_rgbexp32(c)) for c in ((1, 2, 3), (300, 200, 100), (255.5, 1, 1), (0, 0, 0), (1e-3, 2e-3, 5e-4))]
    kt = kd_build_binned_host(s1.tri_verts)
    pins["s1_binned_kd_tree"] = digest(kt["children"], kt["split"], kt["tri_index"])
    return pins


if __name__ == "__main__":
    out = os.path.join(ROOT, "tests", "golden", "oracle_pins.json")
    with open(out, "w") as f:
        json.dump(compute(), f, indent=1, sort_keys=True)
    print("wrote", out)
