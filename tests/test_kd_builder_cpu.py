"""CPU-only: the PRODUCT host code (iterative, parallel SAH kd builder + triangle precompute in
vrad_b200/csrc/kd_builder.cpp) must produce exactly the oracle's tree (the literal recursive restatement of
raytracer/environment.go:238-387) -- same nodes, splits, leaf lists and 48-byte triangle records -- and must not
depend on the number of host threads."""
import os
import subprocess

import numpy as np
import pytest

from oracle import pyoracle
from vrad_b200 import scenes

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def kd_dump(tmp_path_factory):
    out = tmp_path_factory.mktemp("kd") / "kd_dump"
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    subprocess.run([cxx, "-O2", "-std=c++17", "-fopenmp", "-ffp-contract=off", "-I", os.path.join(ROOT, "include"), "-o", str(out),
                    os.path.join(ROOT, "tests", "helpers", "kd_dump.cpp"), os.path.join(ROOT, "vrad_b200", "csrc", "kd_builder.cpp")], check=True)
    return str(out)


def _run(kd_dump, scene, tmp_path, threads):
    fin, fout = tmp_path / "in.bin", tmp_path / f"out{threads}.bin"
    with open(fin, "wb") as f:
        np.array([scene.n_tris], np.int64).tofile(f); scene.tri_ids.astype(np.int32).tofile(f); scene.tri_verts.astype(np.float32).tofile(f)
    subprocess.run([kd_dump, str(fin), str(fout)], check=True, env=dict(os.environ, OMP_NUM_THREADS=str(threads)))
    with open(fout, "rb") as f:
        nn, ni, depth, leaves, vdepth, vbad = np.fromfile(f, np.int64, 6)
        d = {"children": np.fromfile(f, np.int32, nn), "split": np.fromfile(f, np.float32, nn), "tri_index": np.fromfile(f, np.int32, ni),
             "aabb": np.fromfile(f, np.float32, 6), "tris": np.fromfile(f, pyoracle.TRI48_DTYPE, scene.n_tris)}
    return d, int(depth), int(leaves), int(vdepth), int(vbad)


@pytest.mark.parametrize("which", ["s1", "s2_small", "s3_small", "tiny"])
def test_product_builder_equals_oracle(which, kd_dump, tmp_path):
    scene = {"s1": scenes.box_room, "s2_small": lambda: scenes.multi_room(nx=3, ny=2),
             "s3_small": lambda: scenes.outdoor(cells=96, n_buildings=40),
             "tiny": lambda: scenes.box_room(n_boxes=0)}[which]()
    orc = pyoracle.OracleEnv(); orc.add_triangles(scene.tri_ids, scene.tri_verts); orc.build()
    ref, sizes = orc.export(), orc.sizes()
    got1, depth, leaves, vdepth, vbad = _run(kd_dump, scene, tmp_path, 1)
    got4, *_ = _run(kd_dump, scene, tmp_path, 4)
    for got in (got1, got4):
        assert np.array_equal(got["children"], ref["children"])
        assert np.array_equal(got["split"].view(np.uint32), ref["split"].view(np.uint32))
        assert np.array_equal(got["tri_index"], ref["tri_index"])
        assert np.array_equal(got["aabb"], ref["aabb"])
        assert got["tris"].tobytes() == ref["tris"].tobytes()
    assert depth == sizes["max_depth"] == vdepth and leaves == sizes["n_leaves"]
    assert vbad == -1                                   # a tree with an out-of-range child is refused
    assert depth <= 31                                  # fits the kernels' per-ray stack
