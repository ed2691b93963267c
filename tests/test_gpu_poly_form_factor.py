"""GPU side of the polygon-to-differential form factor (vrad_patches_set_windings; SURVEY App. B.3 "optional", parity unpinned):
K2 with windings against the oracle on the 3x2-room map with its patch hierarchy (same columns; weights within 1e-5 -- asinf is the
one libm call in the build whose last bit CUDA and glibc do not share), against the closed form for a square, and through the
in-process multi-GPU handle."""
import math

import numpy as np
import pytest

from vrad_b200 import scenes

pytestmark = pytest.mark.gpu


def _square_env(half, dist):
    from vrad_b200.environment import Environment
    g = Environment()
    g.add_triangles(np.int32([0]), np.float32([[1e4, 1e4, 1e4, 1e4 + 1, 1e4, 1e4, 1e4, 1e4 + 1, 1e4]]))
    g.setup_acceleration_structure()
    g.patches_upload(np.float32([[0, 0, 0], [0, 0, dist]]), np.float32([[0, 0, 1], [0, 0, -1]]), np.float32([0.0, -dist]),
                     np.float32([1.0, 4 * half * half]), np.full((2, 3), 0.5, np.float32))
    sq = np.float32([[half, -half, dist], [-half, -half, dist], [-half, half, dist], [half, half, dist]])     # counter-clockwise from the front: reversed on upload
    rc = np.float32([[0.5, -0.5, 0], [-0.5, -0.5, 0], [-0.5, 0.5, 0], [0.5, 0.5, 0]])
    g.set_windings([0, 4], [4, 4], np.concatenate([rc, sq]))
    return g


def test_square_over_element_closed_form_on_the_gpu():
    from test_poly_form_factor_cpu import corner_form_factor
    for half, dist in ((64.0, 32.0), (64.0, 100.0), (16.0, 8.0)):
        g = _square_env(half, dist)
        g.build_transfers(None)
        rp, col, w = g.transfers_download()
        g.close()
        assert list(col[rp[0]:rp[1]]) == [1]
        want = 4 * corner_form_factor(half / dist, half / dist)
        assert abs(float(w[rp[0]]) - want) <= 2e-6 * max(want, 1.0)


def test_hierarchical_map_with_windings_matches_the_oracle():
    from oracle import pyoracle
    from vrad_b200.environment import environment_from_scene
    hs = scenes.multi_room_hier(nx=3, ny=2)
    t = hs.meta["tree"]
    g = environment_from_scene(hs); o = pyoracle.env_from_scene(hs)
    for e in (g, o):
        e.set_hierarchy(t["parent"], t["child1"], t["child2"], t["face"])
    nnz_plain = g.build_transfers(hs.pvs)
    rp0, c0, w0 = g.transfers_download()
    for e in (g, o):
        e.set_windings(t["wind_first"], t["wind_count"], t["wind_points"])
    nnz_g = g.build_transfers(hs.pvs)
    nnz_o = o.build_transfers(hs.pvs, threads=8)
    rg, cg, wg = g.transfers_download()
    ro, co, wo = o.transfers()
    assert nnz_g == nnz_o
    assert np.array_equal(rg, ro) and np.array_equal(cg, co)
    assert np.abs(wg - wo).max() <= 1e-5 * np.abs(wo).max()
    assert np.allclose(wg, wo, rtol=2e-5, atol=1e-9)
    # the near pairs did change, the far ones did not; rows still sum to at most 1
    if nnz_plain == nnz_g and np.array_equal(c0, cg):
        changed = wg != w0
        assert 0.001 < changed.mean() < 0.5
    sums = np.add.reduceat(wg.astype(np.float64), rg[:-1][np.diff(rg) > 0])
    assert sums.max() <= 1.0 + 1e-5
    # bounced light with the new weights: within 1e-4 of the oracle
    emit0 = scenes.SplitMix64(3).uniform(3 * hs.n_patches, 0.0, 200.0).reshape(hs.n_patches, 3)
    tg, _, _ = g.bounce(emit0, 6)
    to, _, _ = o.bounce(emit0, 6, threads=8)
    assert np.abs(tg - to).max() <= 1e-4 * np.abs(to).max()
    # removing the windings restores the differential rows bit for bit
    g.set_windings(None, None, None)
    assert g.build_transfers(hs.pvs) == nnz_plain
    rp1, c1, w1 = g.transfers_download()
    assert np.array_equal(c1, c0) and np.array_equal(w1.view(np.uint32), w0.view(np.uint32))
    g.close()


def test_windings_inside_the_multi_gpu_handle():
    import torch
    from vrad_b200.environment import Environment, environment_from_scene
    hs = scenes.multi_room_hier(nx=3, ny=2)
    t = hs.meta["tree"]
    devs = [0, 1] if torch.cuda.device_count() >= 2 else [0, 0]
    one = environment_from_scene(hs)
    many = environment_from_scene(hs, devices=devs)
    for e in (one, many):
        e.set_hierarchy(t["parent"], t["child1"], t["child2"], t["face"])
        e.set_windings(t["wind_first"], t["wind_count"], t["wind_points"])
    assert one.build_transfers(hs.pvs) == many.build_transfers(hs.pvs)
    emit0 = scenes.SplitMix64(3).uniform(3 * hs.n_patches, 0.0, 200.0).reshape(hs.n_patches, 3)
    a, _, _ = one.bounce(emit0, 4)
    b, _, _ = many.bounce(emit0, 4)
    assert np.abs(a - b).max() <= 1e-5 * np.abs(a).max()
    one.close(); many.close()
