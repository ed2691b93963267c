"""Side numbers for the BSP side of the path, one JSON object on stdout: K5 (final light -> RGBExp32) against the HBM
peak with a host-function check, the file-driven bake (.bsp in -> lit .bsp out), and the binned-SAH kd build on the device against the
exact host builder with K1 throughput on both trees.  bench.py runs this in a child process and files the result under "bsp_side";
each part has its own guard and reports what it reached.  Timing: CUDA events on torch's current stream, which the library is told to
use (vrad_env_set_stream), after warm-up; inputs resident in HBM and larger than L2 where it matters.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def k5_part(dev, device, stream, hbm_peak):
    import torch
    from vrad_b200 import bspfile, lib as vlib
    from vrad_b200.environment import Environment
    k5 = Environment(device); k5.set_stream(stream); k5.set_async(True)
    nl = 1 << 24                                             # 16.8 M luxels: 201 MB + 201 MB in, 67 MB out > 126 MB L2
    d_dir = torch.rand((nl, 3), device=dev) * 400.0
    d_ind = torch.rand((nl, 3), device=dev) * 100.0
    d_out = torch.empty(nl, dtype=torch.int32, device=dev)
    l5 = vlib.load()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(3):
        vlib.check(l5.vrad_lightmap_finalize(k5._h, C.c_int64(nl), vlib.ptr(d_dir), vlib.ptr(d_ind), vlib.ptr(d_out)))
    e0.record()
    for _ in range(10):
        vlib.check(l5.vrad_lightmap_finalize(k5._h, C.c_int64(nl), vlib.ptr(d_dir), vlib.ptr(d_ind), vlib.ptr(d_out)))
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    sample = slice(0, 1 << 16)
    want = bspfile.color_to_rgbexp32((d_dir[sample] + d_ind[sample]).cpu().numpy())
    ok = bool(np.array_equal(d_out[sample].cpu().numpy().view(bspfile.RGBEXP32), want))
    k5.close()
    gbs = 28.0 * nl / (ms * 1e-3) / 1e9
    return {"workload": "2^24 luxels, direct + indirect resident in HBM, vrad_lightmap_finalize", "luxels": nl, "ms": ms, "luxels_per_sec": nl / (ms * 1e-3),
            "gbs": gbs, "frac_of_hbm_peak": gbs / hbm_peak, "bytes_per_luxel": 28, "matches_host_function": ok}


def bake_part(device):
    """The file-driven bake at the C3/C4 map size (12 x 11 rooms, 30 occluders each), stage by stage."""
    from vrad_b200 import bake, bspfile
    from vrad_b200.environment import Environment
    Lm, meta = bspfile.synthetic_map(12, 11, boxes_per_room=30, sky_rooms=(5, 60), bump_rooms=(0,))
    with tempfile.TemporaryDirectory() as td:
        src, dst = os.path.join(td, "in.bsp"), os.path.join(td, "out.bsp")
        bspfile.write_bsp(src, Lm, meta)
        t0 = time.perf_counter()
        f = bspfile.BspFile(src)
        face_lump, lighting_lump = f.set_target_faces(False)
        L = f.lumps()
        text = f.get(bspfile.LUMP["ENTITIES"])[0].rstrip(b"\0").decode()
        t1 = time.perf_counter()
        prep = bake.prepare(L, text)
        t2 = time.perf_counter()
        env = Environment(device)
        lit = bake.light(env, prep, bounces=8)
        t3 = time.perf_counter()
        lump, colors = bake.finish(env, prep, lit)
        env.close()
        t4 = time.perf_counter()
        f.set(lighting_lump, lump, version=1); f.set(face_lump, prep["lumps"].faces, version=1); f.save(dst); f.close()
        t5 = time.perf_counter()
        out_size = os.path.getsize(dst)
    # cpu_baseline leg: the same device stages on the CPU oracle (the checker, timed beside -- never on the product path), all host threads
    cpu = None
    try:
        from oracle import pyoracle

        class AllThreads(pyoracle.OracleEnv):
            T = pyoracle.num_threads()
            def build_transfers(self, pvs=None, threads=None): return super().build_transfers(pvs, threads=self.T)
            def direct_light(self, pos, normal, lights, threads=None): return super().direct_light(pos, normal, lights, threads=self.T)
            def bounce(self, emit0, n, early_out=False, threads=None): return super().bounce(emit0, n, early_out, threads=self.T)
        c0 = time.perf_counter(); ref = bake.light(AllThreads(), prep, bounces=8); c1 = time.perf_counter()
        worst = max(float(np.abs(lit[k] - ref[k]).max() / max(np.abs(ref[k]).max(), 1e-30)) for k in ("direct", "emit0", "total"))
        cpu = {"kind": "port", "cores": AllThreads.T, "light_k1_k2_k3_k4_seconds": c1 - c0, "transfers": ref["nnz"], "bounces": ref["bounces_done"],
               "max_relative_difference_gpu_vs_oracle": worst, "transfers_equal": bool(ref["nnz"] == lit["nnz"])}
    except Exception as exc:
        cpu = {"error": repr(exc)}
    return {"workload": "synthetic BSP v20 map at the C3/C4 size (12 x 11 rooms), .bsp in -> lit .bsp out through vrad_b200.bake; first call of every stage",
            "faces": int(Lm.faces.shape[0]), "triangles": int(prep["tri_ids"].shape[0]), "patches": int(prep["tree"]["origin"].shape[0]),
            "leaf_patches": int((prep["tree"]["child1"] == -1).sum()), "luxels": int(prep["lux_pos"].shape[0]), "lights": int(prep["lights"].shape[0]),
            "transfers": lit["nnz"], "bounces": lit["bounces_done"],
            "seconds": {"read": t1 - t0, "prepare_host": t2 - t1, "light_k1_k2_k3_k4": t3 - t2, "radial_k5_pack": t4 - t3, "write": t5 - t4, "total": t5 - t0},
            "lit_luxel_fraction": float((colors.view(np.uint8).reshape(-1, 4)[:, :3].max(axis=1) > 0).mean()),
            "lighting_lump_bytes": len(lump), "file_bytes": out_size, "cpu_baseline": cpu}


def kd_part(dev, device, stream):
    import torch
    from vrad_b200 import scenes
    from vrad_b200.environment import Environment
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    out = {}
    for name, scn, nseg in (("C1_box_room", scenes.box_room(), 1 << 24), ("C5_outdoor", scenes.outdoor(), 1 << 22)):
        sa, sb = scenes.shadow_segments(scn, nseg)
        d_sa, d_sb = torch.from_numpy(sa).to(dev), torch.from_numpy(sb).to(dev)
        d_bits = torch.empty((nseg + 31) // 32, dtype=torch.int32, device=dev)
        row = {"triangles": scn.n_tris, "segments": nseg}
        for kind in ("exact_host", "binned_device", "binned_auto"):      # auto = VRAD_BUILD_AUTO: the host's cores below 10,000 triangles
            ek = Environment(device); ek.add_triangles(scn.tri_ids, scn.tri_verts, scn.tri_flags)
            secs = ek.setup_acceleration_structure() if kind == "exact_host" else ek.build_fast(None if kind == "binned_auto" else False)
            st = ek.stats()
            ek.set_stream(stream); ek.set_async(True)
            for _ in range(2):
                ek.test_lines(d_sa, d_sb, out=d_bits)
            e0.record()
            for _ in range(5):
                ek.test_lines(d_sa, d_sb, out=d_bits)
            e1.record(); torch.cuda.synchronize()
            row[kind] = {"build_seconds": secs, "nodes": st["n_nodes"], "leaves": st["n_leaves"], "index_entries": st["n_idx"], "max_depth": st["max_depth"],
                         "triangles_per_leaf": st["n_idx"] / max(1, st["n_leaves"]),
                         "segments_per_sec": nseg / (e0.elapsed_time(e1) / 5 * 1e-3), "visible": int(np.unpackbits(d_bits.cpu().numpy().view(np.uint8)).sum())}
            ek.close()
        out[name] = row
        del d_sa, d_sb, d_bits
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--device", type=int, default=0)
    ap.add_argument("--hbm-peak", type=float, default=6546.6)
    args = ap.parse_args()
    import torch
    if not torch.cuda.is_available():
        print(json.dumps({"error": "no CUDA device (there is no CPU fallback)"}))
        return
    torch.cuda.set_device(args.device)
    dev = torch.device("cuda", args.device)
    stream = torch.cuda.current_stream().cuda_stream
    result = {}
    for key, fn in (("k5_finalize", lambda: k5_part(dev, args.device, stream, args.hbm_peak)), ("bake_file", lambda: bake_part(args.device)),
                    ("kd_build_fast", lambda: kd_part(dev, args.device, stream))):
        try:
            result[key] = fn()
        except Exception as exc:
            result[key] = {"error": repr(exc)}
    print(json.dumps(result))


if __name__ == "__main__":
    main()
