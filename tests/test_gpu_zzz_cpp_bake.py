"""GPU: integration/cpp/drive --bake (bake::BakeFile, the C++ host side over the C-ABI) against vrad_b200.bake.bake_file on the same .bsp.
Kept in its own file, collected after every other GPU test."""
import numpy as np
import pytest

from vrad_b200 import bspfile as B

pytestmark = pytest.mark.gpu


def test_cpp_driver_bakes_the_same_file(tmp_path):
    """integration/cpp/drive --bake (bake::BakeFile, the C++ host side) against vrad_b200.bake.bake_file on the same .bsp: same transfer
    count, same direct / emitted / bounced light, same lighting and face lumps in the written file -- both run the same device stages
    through the C-ABI, so the comparison is exact."""
    import os
    import subprocess
    from vrad_b200 import bake
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    L, meta = B.synthetic_map(2, 2, boxes_per_room=4, sky_rooms=(1,), bump_rooms=(0,), ramps=True)
    src, out_py, out_cpp = str(tmp_path / "in.bsp"), str(tmp_path / "py.bsp"), str(tmp_path / "cpp.bsp")
    B.write_bsp(src, L, meta)
    res = bake.bake_file(src, out_py, device=0, bounces=8)
    subprocess.run(["make", "-C", os.path.join(root, "integration", "cpp")], check=True, capture_output=True)
    anorms = os.path.join(root, "vrad_b200", "data", "anorms.txt")
    r = subprocess.run([os.path.join(root, "integration", "cpp", "drive"), "--bake", src, out_cpp, anorms, "-bounce", "8"], check=True, capture_output=True, text=True)
    words = r.stdout.split()
    got = {words[i]: int(words[i + 1]) for i in range(1, len(words), 2)}

    def checksum(a):
        b = np.ascontiguousarray(a).tobytes(); b += b"\0" * (-len(b) % 4)
        w = np.frombuffer(b, "<u4").astype(np.uint64)
        with np.errstate(over="ignore"):
            return int(np.sum(w * (np.uint64(2654435761) * np.arange(w.shape[0], dtype=np.uint64) + np.uint64(1)), dtype=np.uint64))
    lit = res["lit"]
    assert got["transfers"] == lit["nnz"] and got["bounces"] == lit["bounces_done"]
    assert got["direct"] == checksum(lit["direct"]) and got["emit"] == checksum(lit["emit0"]) and got["total"] == checksum(lit["total"])
    assert got["bump"] == checksum(lit["bump_totals"])
    a, b = B.BspFile(out_py), B.BspFile(out_cpp)
    for lump in range(64):
        assert a.get(lump) == b.get(lump), lump
    a.close(); b.close()
