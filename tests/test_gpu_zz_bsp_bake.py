"""GPU parity for the BSP side: K5 (final light -> ColorRGBExp32) against the host entry point running the same inline function and
the oracle's restatement; vrad_env_add_bsp; and the file-driven bake (.bsp in -> lit .bsp out) against the same sequence run on the
CPU oracle's environment: K1/K2 results exact (transfer count), K3/K4 within 1e-4 (fp32 radiance, north_star), the lighting lump
equal up to the truncation step of the 8-bit mantissas."""
import numpy as np
import pytest

from vrad_b200 import bspfile as B

pytestmark = pytest.mark.gpu
RTOL = 1e-4


@pytest.fixture(scope="module")
def env():
    from vrad_b200.environment import Environment
    e = Environment(0)
    yield e
    e.close()


def _rows(c):
    return np.stack([c["r"], c["g"], c["b"], c["exponent"]], axis=1).astype(np.int64)


@pytest.mark.parametrize("n", [1, 3, 4, 5, 1023, 100003])
def test_k5_matches_host_function(env, n):
    rng = np.random.default_rng(n)
    mags = np.float32(2.0) ** rng.integers(-20, 14, (n, 1)).astype(np.float32)
    direct = (rng.uniform(0, 1, (n, 3)).astype(np.float32) * mags).astype(np.float32)
    indirect = (rng.uniform(0, 1, (n, 3)).astype(np.float32) * mags * np.float32(0.25)).astype(np.float32)
    direct[::7] = 0; indirect[::7] = 0; direct[::11, 0] = -1.0
    got = B.lightmap_finalize(env, direct, indirect)
    want = B.color_to_rgbexp32(direct + indirect)                 # fp32 add, then the shared inline pack function on the host
    assert np.array_equal(_rows(got), _rows(want))
    got1 = B.lightmap_finalize(env, direct)                       # no bounced light
    assert np.array_equal(_rows(got1), _rows(B.color_to_rgbexp32(direct)))
    from oracle import bspside as O             # the independent restatement, every size (round-1 verdict: device vs oracle, not only vs the shared source)
    assert np.array_equal(_rows(got), np.asarray([O.pack_rgbexp32(c) for c in direct + indirect], np.int64))


def test_k5_device_pointers_and_empty(env):
    import torch
    n = 50001
    rng = np.random.default_rng(2)
    direct = rng.uniform(0, 500, (n, 3)).astype(np.float32)
    d = torch.from_numpy(direct).cuda()
    out = torch.zeros(n, dtype=torch.int32, device="cuda")
    from vrad_b200 import lib
    import ctypes as C
    lib.check(lib.load().vrad_lightmap_finalize(env._h, C.c_int64(n), lib.ptr(d), None, lib.ptr(out)))
    torch.cuda.synchronize()
    assert np.array_equal(out.cpu().numpy().view(B.RGBEXP32), B.color_to_rgbexp32(direct))
    # a device pointer at an odd (12-byte) offset takes the staged path and gives the same answer
    lib.check(lib.load().vrad_lightmap_finalize(env._h, C.c_int64(n - 1), C.c_void_p(d.data_ptr() + 12), None, lib.ptr(out)))
    torch.cuda.synchronize()
    assert np.array_equal(out.cpu().numpy()[:n - 1].view(B.RGBEXP32), B.color_to_rgbexp32(direct[1:]))
    assert B.lightmap_finalize(env, np.zeros((0, 3), np.float32)).shape[0] == 0


@pytest.mark.parametrize("n", [2, 4, 4099])
def test_k5_with_patch_lookup(env, n):
    rng = np.random.default_rng(100 + n)
    npatch = 37
    direct = rng.uniform(0, 300, (n, 3)).astype(np.float32)
    total = rng.uniform(0, 100, (npatch, 3)).astype(np.float32)
    idx = rng.integers(-1, npatch, n).astype(np.int32)
    got = B.lightmap_finalize_patches(env, direct, idx, total)
    ind = np.where(idx[:, None] >= 0, total[np.maximum(idx, 0)], np.float32(0)).astype(np.float32)
    assert np.array_equal(_rows(got), _rows(B.color_to_rgbexp32(direct + ind)))
    from vrad_b200.lib import VradError
    bad = idx.copy(); bad[0] = npatch
    with pytest.raises(VradError):
        B.lightmap_finalize_patches(env, direct, bad, total)


def test_radial_filter_device_equals_host_policy(env):
    """vrad_luxel_radial_light: the gather functor as a CUDA kernel against the same functor on the host's cores (which the CPU tests
    compare with the oracle's scatter form) -- bit for bit, flat and bump-mapped blocks, host and device inputs."""
    from vrad_b200 import bake
    L, meta = B.synthetic_map(3, 2, boxes_per_room=5, sky_rooms=(1,), bump_rooms=(0,))
    prep = bake.prepare(L, meta["entities"])
    N = prep["tree"]["origin"].shape[0]
    rng = np.random.default_rng(8)
    totals = rng.uniform(0, 300, (N, 3)).astype(np.float32)
    bump = rng.uniform(0, 300, (N, 3, 3)).astype(np.float32)
    args = (prep["lux_face"], prep["luxel_first"], prep["lm_size"], prep["radial_first"], prep["radial_entries"])
    for bp in (None, bump):
        want = B.luxel_radial_light(None, *args, totals, bp)
        got = B.luxel_radial_light(env, *args, totals, bp)
        assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
    assert (want != 0).any(axis=1).mean() > 0.95
    # a sub-range of luxels (what a rank of a sharded bake computes): offsets shifted, same values
    a, b = 1000, 1000 + 12345
    part = B.luxel_radial_light(env, prep["lux_face"][a:b], prep["luxel_first"] - a, prep["lm_size"], prep["radial_first"], prep["radial_entries"], totals, bump)
    assert np.array_equal(part, want[a:b])


def test_radial_filter_device_equals_the_oracle_scatter_form(env):
    """The device gather against the ORACLE's scatter restatement (oracle/bspside.py::build_patch_radial, upstream's BuildPatchRadial /
    AddBouncedToRadial / SampleRadial order) on every lit face of the 3x2-room map -- bit for bit, flat and bump blocks.  The test above
    holds the kernel against the same functor on the host; this one against independent code."""
    from oracle import bspside as O
    from vrad_b200 import bake
    L, meta = B.synthetic_map(3, 2, boxes_per_room=5, sky_rooms=(1,), bump_rooms=(0,))
    prep = bake.prepare(L, meta["entities"])
    t = prep["tree"]
    N = t["origin"].shape[0]
    rng = np.random.default_rng(18)
    totals = rng.uniform(0, 300, (N, 3)).astype(np.float32)
    bump = rng.uniform(0, 300, (N, 3, 3)).astype(np.float32)
    got = B.luxel_radial_light(env, prep["lux_face"], prep["luxel_first"], prep["lm_size"], prep["radial_first"], prep["radial_entries"], totals, bump)
    patch_lists = [[] for _ in range(L.faces.shape[0])]
    for p_ in range(N):
        if t["child1"][p_] == -1:
            patch_lists[prep["face_of_patch"][p_]].append(p_)
    vn, nb_first, nb = B.pair_edges(L)
    nbs = [list(nb[nb_first[f]:nb_first[f + 1]]) for f in range(L.faces.shape[0])]
    lit_faces = np.nonzero(np.diff(prep["luxel_first"]) > 0)[0]
    bump_faces = set(np.nonzero(L.texinfo["flags"][L.faces["texinfo"]] & B.SURF_BUMPLIGHT)[0].tolist())
    checked = 0
    for f in lit_faces[::2]:
        f = int(f)
        want = O.build_patch_radial(L, f, prep["lm_mins"], prep["lm_size"], patch_lists, t, totals, nbs, prep["face_origin"])
        a = int(prep["luxel_first"][f])
        assert np.array_equal(got[a:a + want.shape[0]].view(np.uint32), want.view(np.uint32)), f
        checked += want.shape[0]
        if f in bump_faces:
            for b in range(1, 4):
                wb = O.build_patch_radial(L, f, prep["lm_mins"], prep["lm_size"], patch_lists, t, bump[:, b - 1], nbs, prep["face_origin"])
                assert np.array_equal(got[a + b * want.shape[0]:a + (b + 1) * want.shape[0]].view(np.uint32), wb.view(np.uint32)), (f, b)
    assert checked > 15000


def test_env_add_bsp_traces_like_the_oracle():
    """vrad_env_add_bsp (brush entity + world brushes + sky faces) then K1: visibility bits and closest hits equal the oracle's on
    the same triangles."""
    import ctypes as C
    from oracle import pyoracle
    from vrad_b200 import bake, lib, scenes
    from vrad_b200.environment import Environment
    L, meta = B.synthetic_map(3, 2, boxes_per_room=8, sky_rooms=(1,))
    ents = bake.parse_entities(meta["entities"])
    cm, co, ca = bake.shadow_casters(ents)
    assert list(cm) == [1]
    g = Environment(0)
    n = C.c_int()
    lib.check(lib.load().vrad_env_add_bsp(g._h, L.ref, C.c_int(len(cm)), lib.ptr(cm), lib.ptr(co), lib.ptr(ca), C.byref(n)))
    ids, verts = B.raytrace_triangles(L, cm, co, ca)
    assert n.value == ids.shape[0]
    g.setup_acceleration_structure()
    o = pyoracle.OracleEnv(); o.add_triangles(ids, verts.reshape(-1, 9)); o.build()
    assert g.stats()["n_tris"] == ids.shape[0]
    rng = scenes.SplitMix64(77)
    m = 1 << 16
    a = rng.uniform(3 * m, 8.0, 500.0).reshape(3, m).astype(np.float32)
    b = rng.uniform(3 * m, 8.0, 1000.0).reshape(3, m).astype(np.float32)
    a[0] *= 3; b[0] *= 1.5                                          # spread over the 3 x 2 rooms
    assert np.array_equal(g.test_lines(a, b), o.test_lines(a, b, threads=8))
    assert np.array_equal(g.test_lines(a, b, sky_mode=1), o.test_lines(a, b, sky_mode=1, threads=8))
    g.close()


def test_bsp_file_bake(tmp_path):
    from oracle import pyoracle
    from vrad_b200 import bake
    L, meta = B.synthetic_map(2, 2, boxes_per_room=4, sky_rooms=(1,), bump_rooms=(0,), ramps=True)
    src, dst = str(tmp_path / "in.bsp"), str(tmp_path / "out.bsp")
    B.write_bsp(src, L, meta)
    res = bake.bake_file(src, dst, device=0, bounces=8)
    prep, lit = res["prep"], res["lit"]
    # the same device stages on the CPU oracle
    ref = bake.light(pyoracle.OracleEnv(), prep, bounces=8)
    assert lit["nnz"] == ref["nnz"] and lit["bounces_done"] == ref["bounces_done"]
    for key in ("direct", "emit0", "total", "bump_totals"):
        assert np.abs(lit[key] - ref[key]).max() <= RTOL * float(np.abs(ref[key]).max()), key
    assert lit["bump_totals"].any()                                # room 0's floor is SURF_BUMPLIGHT
    assert lit["direct"].max() > 1 and lit["total"].max() > 1
    # radial filter + K5 == the same functor on the host's cores + the host pack function, on the GPU's own inputs; the lump is what
    # pack_lighting makes of it
    ind = B.luxel_radial_light(None, prep["lux_face"], prep["luxel_first"], prep["lm_size"], prep["radial_first"], prep["radial_entries"], lit["total"], lit["bump_totals"])
    assert np.array_equal(_rows(res["colors"]), _rows(B.color_to_rgbexp32(lit["direct"] + ind)))
    assert res["lump"] == B.pack_lighting(prep["lumps"], prep["luxel_first"], res["colors"], prep["lump_bytes"])
    # against the oracle's light the decoded luxels agree to the 8-bit truncation step (one part in 128 of the largest component)
    ind_ref = B.luxel_radial_light(None, prep["lux_face"], prep["luxel_first"], prep["lm_size"], prep["radial_first"], prep["radial_entries"], ref["total"], ref["bump_totals"])
    want = ref["direct"] + ind_ref
    got = B.color_from_rgbexp32(res["colors"])
    step = np.maximum(want.max(axis=1, keepdims=True), 1e-3) / 64
    assert np.all(np.abs(got - want) <= step)
    # the written file: same lumps except LIGHTING (version 1) and FACES (lightofs / styles / extents)
    f = B.BspFile(dst)
    lump, ver = f.get(B.LUMP["LIGHTING"])
    assert lump == res["lump"] and ver == 1 and len(lump) == prep["lump_bytes"]
    L2 = f.lumps()
    assert np.array_equal(L2.faces, prep["lumps"].faces)
    for k in L.a:
        if k != "faces":
            assert np.array_equal(L.a[k], L2.a[k]), k
    lit_faces = L2.faces["lightofs"] >= 0
    assert lit_faces.sum() == (np.diff(prep["luxel_first"]) > 0).sum() and np.all(L2.faces["styles"][lit_faces][:, 0] == 0)
    f.close()
