"""Polygon-to-differential form factor of MakeTransfer (SURVEY App. B.3 "optional"; upstream vismat.cpp -- not in the reference, parity
unpinned): the oracle's contour integral against the closed form for a rectangle over a differential element, the switch-over rule
pi * 0.04 * |delta|^2 < area_j, and independence of the winding's direction."""
import math

import numpy as np

from oracle import pyoracle


def corner_form_factor(x, y):
    """differential element under one corner of an X x Y rectangle at unit height, parallel planes (Howell C-3 / Siegel-Howell)"""
    return (x / math.sqrt(1 + x * x) * math.atan(y / math.sqrt(1 + x * x)) + y / math.sqrt(1 + y * y) * math.atan(x / math.sqrt(1 + y * y))) / (2 * math.pi)


def two_patch_env(half, dist, reverse=False, with_windings=True):
    """receiver patch 0 at the origin facing +z; emitter patch 1 a square of half-side `half` at height `dist`, facing -z"""
    origin = np.float32([[0, 0, 0], [0, 0, dist]])
    normal = np.float32([[0, 0, 1], [0, 0, -1]])
    plane_dist = np.float32([0.0, -dist])
    area = np.float32([1.0, 4 * half * half])
    refl = np.full((2, 3), 0.5, np.float32)
    o = pyoracle.OracleEnv()
    o.add_triangles(np.int32([0]), np.float32([[1e4, 1e4, 1e4, 1e4 + 1, 1e4, 1e4, 1e4, 1e4 + 1, 1e4]]))   # something to build a tree from, far away
    o.build()
    o.patches_upload(origin, normal, plane_dist, area, refl)
    if with_windings:
        sq = np.float32([[half, -half, dist], [half, half, dist], [-half, half, dist], [-half, -half, dist]])     # clockwise seen from below (the front)
        if reverse: sq = sq[::-1].copy()
        rc = np.float32([[0.5, -0.5, 0], [-0.5, -0.5, 0], [-0.5, 0.5, 0], [0.5, 0.5, 0]])
        o.set_windings([0, 4], [4, 4], np.concatenate([rc, sq]))
    return o


def row0(o):
    o.build_transfers(None)
    rp, col, w = o.transfers()
    return {int(c): float(v) for c, v in zip(col[rp[0]:rp[1]], w[rp[0]:rp[1]])}


def test_square_over_element_matches_closed_form():
    for half, dist in ((64.0, 32.0), (64.0, 100.0), (16.0, 8.0), (100.0, 10.0)):
        got = row0(two_patch_env(half, dist))[1]
        want = 4 * corner_form_factor(half / dist, half / dist)
        # fp32 asin near 90 degrees per edge (the last case) costs a few 1e-5; the others agree to 1e-6
        assert abs(got - want) <= (3e-5 if half / dist > 5 else 2e-6) * max(want, 1.0), (half, dist, got, want)
        # the differential form overshoots badly this close: that is what the switch is for
        diff = row0(two_patch_env(half, dist, with_windings=False))[1]
        assert diff > want


def test_switch_over_rule_and_far_pairs_keep_the_differential_form():
    half, dist = 8.0, 200.0                      # pi * 0.04 * 200^2 = 5027 > area 256: differential form, windings or not
    a = row0(two_patch_env(half, dist))[1]
    b = row0(two_patch_env(half, dist, with_windings=False))[1]
    assert a == b
    np.testing.assert_allclose(a, 256.0 / (math.pi * dist * dist), rtol=1e-6)
    # just inside the rule the two forms differ
    half, dist = 64.0, 300.0                     # pi * 0.04 * 9e4 = 11310 < 16384
    assert row0(two_patch_env(half, dist))[1] != row0(two_patch_env(half, dist, with_windings=False))[1]


def test_winding_direction_does_not_matter_and_windings_can_be_removed():
    a = row0(two_patch_env(64.0, 32.0))
    b = row0(two_patch_env(64.0, 32.0, reverse=True))
    assert a == b
    o = two_patch_env(64.0, 32.0)
    o.set_windings(None, None, None)
    assert row0(o) == row0(two_patch_env(64.0, 32.0, with_windings=False))
