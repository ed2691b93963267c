"""The C++ host mirror (integration/cpp/vrad_environment.hpp: raytracer::Environment with the
reference's method names over the C-ABI) driven by a compiled C++ program, checked against the oracle."""
import os
import subprocess

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_cpp_driver_matches_oracle():
    from oracle import pyoracle
    from vrad_b200 import build as vbuild, scenes
    vbuild.build()
    subprocess.run(["make", "-C", os.path.join(ROOT, "integration", "cpp")], check=True, capture_output=True)
    out = subprocess.run([os.path.join(ROOT, "integration", "cpp", "drive")], check=True, capture_output=True, text=True).stdout.splitlines()
    g = scenes._Geom()
    OP = scenes.TRACE_ID_OPAQUE
    g.add_box(OP, (0, 0, 0), (256, 256, 256)); g.add_box(OP + 16, (40, 40, 0), (100, 90, 64)); g.add_box(OP + 32, (150, 120, 0), (220, 200, 128))
    ids, verts, flags = g.arrays()
    o = pyoracle.OracleEnv(); o.add_triangles(ids, verts, flags); o.build()
    dirs = np.array([[0.6, 0.0, -0.8], [-0.6, 0.48, -0.64], [0.0, 1.0, 0.0], [0.36, 0.48, 0.8]], np.float32)
    # the sky scenario: hand-derived fractions (same scene and cases as scenes.MINI_SKY_CASES), all four lanes alike
    sky = [[float(x) for x in l.split()[1:]] for l in out if l.startswith("sky ")]
    assert sky == [[v] * 4 for v in (0.0, 1.0, 0.0, 0.25, 0.75, 0.0, 0.25)]
    assert [l for l in out if l.startswith("colour")] == ["colour 0.25"]
    # the host-only subdivision scenario: 256 x 128 face -> 8 x 4 leaves of 32 x 32, numbered depth first
    assert [l for l in out if l.startswith("patches")] == ["patches 63 leaves 32 leaf_area 32768 child1 1 child2 2"]
    out = [l for l in out if not l.startswith(("sky ", "colour", "patches"))]
    lines = [l.split() for l in out if not l.startswith("vis")]
    assert len(lines) == 16
    for p in range(4):
        org = np.zeros((3, 4), np.float32); d = np.zeros((3, 4), np.float32)
        for l in range(4):
            org[:, l] = (np.float32(20.0) + np.float32(50.0) * l + np.float32(3.0) * p, np.float32(30.0) + np.float32(40.0) * p, 200.0)
            d[:, l] = dirs[(l + p) % 4]
        hi, hd, hn = o.trace4_packet(org, d, np.zeros(4, np.float32), np.full(4, 1000, np.float32))
        for l in range(4):
            rec = lines[4 * p + l]
            assert int(rec[2]) == hi[l]
            assert np.float32(rec[3]) == hd[l]
            assert np.allclose([float(x) for x in rec[4:7]], hn[:, l], atol=1e-7)
        assert np.all(hi >= 0)
    a = np.zeros((3, 4), np.float32); b = np.zeros((3, 4), np.float32)
    for l in range(4):
        a[:, l] = (10, 10 + 60.0 * l, 10); b[:, l] = (250, 240 - 50.0 * l, 20 + 60.0 * l)
    bits = int(o.test_lines(a, b, sky_mode=1)[0])
    vis = [float(x) for x in [l for l in out if l.startswith("vis")][0].split()[1:]]
    assert vis == [float((bits >> l) & 1) for l in range(4)]
