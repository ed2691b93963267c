#!/usr/bin/env python
"""Round-2 tuning sweep on one B200 (run under gpurun; writes gpurun_out/r02_tune.json).

K1: segments/s on S1 (2^24 C1 segments) and S3 (2^22 random segments, exact tree) for the coherence pre-pass
(k1_sort) x shared-memory top levels (k1_top), the index-pair entry point, and the host-buffer (e2e) forms.
K4: us/bounce on the full S2 matrix at world 1, and on the rank-3 slice of a simulated world-8 run (k4_sim_peers:
the multi-GPU kernel with peer stores, in-kernel barrier and PDL chaining, all on this device), for item size,
item order, PDL and graph replay.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--skip-s3", action="store_true")
    ap.add_argument("--skip-k4", action="store_true")
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "r02_tune.json"))
    args = ap.parse_args()
    import torch
    from vrad_b200 import scenes
    from vrad_b200.environment import environment_from_scene
    from vrad_b200.lib import PinnedArray

    dev = torch.device("cuda", 0)
    stream = torch.cuda.current_stream().cuda_stream
    res = {"gpu": torch.cuda.get_device_name(0)}
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def timed(fn, reps=3, warm=2):
        for _ in range(warm):
            fn()
        torch.cuda.synchronize()
        e0.record()
        for _ in range(reps):
            fn()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps

    def flush():
        os.makedirs(os.path.dirname(args.out), exist_ok=True)
        with open(args.out, "w") as f:
            json.dump(res, f, indent=1)

    # ---------------- K1 on S1 ----------------
    s1 = scenes.box_room()
    env = environment_from_scene(s1, with_patches=False)
    env.set_stream(stream); env.set_async(True)
    n = 1 << 24
    a, b = scenes.shadow_segments(s1, n, seed=0xC0FFEE)
    pts, pairs = scenes.shadow_segment_indices(s1, n, seed=0xC0FFEE)
    d_a, d_b = torch.from_numpy(a).to(dev), torch.from_numpy(b).to(dev)
    d_pairs = torch.from_numpy(pairs).to(dev)
    d_bits = torch.empty(n // 32, dtype=torch.int32, device=dev)
    env.points_upload(pts)
    k1 = {}
    ref = None
    for sort, key, top in ((0, 0, 0), (0, 0, -1), (1, 0, 0), (1, -1, 0)):
        env.set_option("k1_sort", sort); env.set_option("k1_top", max(top, 0)); env.set_option("k1_key", key); env.set_option("k1_stream", 0 if top < 0 else 1)
        ms = timed(lambda: env.test_lines(d_a, d_b, out=d_bits))
        got = d_bits.cpu().numpy()
        if ref is None:
            ref = got
        k1[f"coords_sort{sort}_key{key}_top{top}"] = {"ms": ms, "seg_per_s": n / ms * 1e3, "same_bits": bool(np.array_equal(got, ref))}
        print("S1", sort, key, top, ms, flush=True)
    env.set_option("k1_top", 0)
    env.set_option("k1_stream", 1)
    for sort, key in ((0, 0), (1, 0)):
        env.set_option("k1_sort", sort); env.set_option("k1_key", key)
        ms = timed(lambda: env.test_lines_indexed(d_pairs, out=d_bits))
        k1[f"indexed_sort{sort}_key{key}"] = {"ms": ms, "seg_per_s": n / ms * 1e3, "same_bits": bool(np.array_equal(d_bits.cpu().numpy(), ref))}
        print("S1 indexed", sort, key, ms, flush=True)
    env.set_option("k1_key", 0)
    # host-buffer forms (e2e): pinned coordinates (24 B/segment) vs pinned index pairs (8 B/segment)
    env.set_async(False)
    h_a, h_b = PinnedArray((3, n), np.float32), PinnedArray((3, n), np.float32)
    h_p = PinnedArray((n, 2), np.int32); h_bits = PinnedArray((n // 32,), np.uint32)
    h_a.array[...] = a; h_b.array[...] = b; h_p.array[...] = pairs
    for sort in (-1, 0, 1):
        env.set_option("k1_sort", sort)
        for name, fn in (("coords", lambda: env.test_lines(h_a.array, h_b.array, out=h_bits.array)),
                         ("indexed", lambda: env.test_lines_indexed(h_p.array, out=h_bits.array))):
            fn()
            t0 = time.perf_counter()
            for _ in range(3):
                fn()
            dt = (time.perf_counter() - t0) / 3
            k1[f"e2e_{name}_sort{sort}"] = {"ms": dt * 1e3, "seg_per_s": n / dt, "same_bits": bool(np.array_equal(h_bits.array.view(np.int32), ref))}
            print("S1 e2e", name, sort, dt * 1e3, flush=True)
    res["k1_s1"] = k1
    h_a.free(); h_b.free(); h_p.free(); h_bits.free()
    env.close(); del d_a, d_b, d_pairs, d_bits
    flush()

    # ---------------- K1 on S3 (exact tree) ----------------
    if not args.skip_s3:
        s3 = scenes.outdoor()
        env = environment_from_scene(s3, with_patches=False)
        env.set_stream(stream); env.set_async(True)
        n3 = 1 << 22
        a3, b3 = scenes.shadow_segments(s3, n3, seed=0xC5)
        d_a, d_b = torch.from_numpy(a3).to(dev), torch.from_numpy(b3).to(dev)
        d_bits = torch.empty(n3 // 32, dtype=torch.int32, device=dev)
        k3 = {"stats": {k: (v if not hasattr(v, "tolist") else v.tolist()) for k, v in env.stats().items()}}
        ref = None
        for sort, key, st in ((0, 0, 0), (0, 0, 1), (1, 0, 1), (1, 2, 1), (1, -1, 1)):
            env.set_option("k1_sort", sort); env.set_option("k1_key", key); env.set_option("k1_stream", st)
            ms = timed(lambda: env.test_lines(d_a, d_b, out=d_bits))
            got = d_bits.cpu().numpy()
            if ref is None:
                ref = got
            k3[f"sort{sort}_key{key}_stream{st}"] = {"ms": ms, "seg_per_s": n3 / ms * 1e3, "same_bits": bool(np.array_equal(got, ref))}
            print("S3", sort, key, st, ms, flush=True)
        res["k1_s3"] = k3
        env.close(); del d_a, d_b, d_bits
        flush()

    # ---------------- K1 on S2 (49,586 tris: 3 MB of nodes and triangles) ----------------
    s2k = scenes.multi_room()
    env = environment_from_scene(s2k, with_patches=False)
    env.set_stream(stream); env.set_async(True)
    n2 = 1 << 24
    a2, b2 = scenes.shadow_segments(s2k, n2, seed=0xC2)
    d_a, d_b = torch.from_numpy(a2).to(dev), torch.from_numpy(b2).to(dev)
    d_bits = torch.empty(n2 // 32, dtype=torch.int32, device=dev)
    k2r = {}
    ref = None
    for sort in (0, 1, -1):
        env.set_option("k1_sort", sort)
        ms = timed(lambda: env.test_lines(d_a, d_b, out=d_bits))
        got = d_bits.cpu().numpy()
        if ref is None:
            ref = got
        k2r[f"sort{sort}"] = {"ms": ms, "seg_per_s": n2 / ms * 1e3, "same_bits": bool(np.array_equal(got, ref))}
        print("S2 rays", sort, ms, flush=True)
    res["k1_s2"] = k2r
    env.close(); del d_a, d_b, d_bits
    flush()

    # ---------------- K4 on S2 ----------------
    if not args.skip_k4:
        s2 = scenes.multi_room()
        N = s2.n_patches
        emit0 = scenes.SplitMix64(0xE1).uniform(3 * N, 0.0, 200.0).reshape(N, 3)
        d_emit = torch.from_numpy(emit0).to(dev); d_tot = torch.empty_like(d_emit)
        k4 = {}
        for world, rank in ((1, 0), (8, 3), (2, 1), (4, 0)):
            env = environment_from_scene(s2, rank=rank, world=world)
            env.set_stream(stream)
            if world > 1:
                env.set_option("k4_sim_peers", 1)
            nnz = env.build_transfers(s2.pvs)
            row0, row1, _ = env.transfers_info()
            if world == 1:
                # row-length histogram of the C4 matrix (for DESIGN): from the row pointers of 8 sampled blocks
                lens = []
                for r0 in range(0, N, N // 8):
                    rpb, _, _ = env.transfers_download_rows(r0, min(N, r0 + 2048), 1 << 23)
                    lens.append(np.diff(rpb))
                lens = np.concatenate(lens)
                k4["row_length_sample"] = {"rows": int(lens.size), "mean": float(lens.mean()), "p50": float(np.percentile(lens, 50)), "p90": float(np.percentile(lens, 90)),
                                           "p99": float(np.percentile(lens, 99)), "max": int(lens.max()), "min": int(lens.min()),
                                           "frac_gt_2048": float((lens > 2048).mean()), "frac_gt_4096": float((lens > 4096).mean())}
                print("row lengths", k4["row_length_sample"], flush=True)
            env.set_async(True)
            tag = f"world{world}_rank{rank}"
            k4[tag] = {"nnz_local": nnz, "rows": [row0, row1], "stream_us_at_peak": 8 * nnz / 6455.3e9 * 1e6}
            # (items kernel, block, persist, pool %, pdl, sim)
            if world == 1:
                sweeps = [(0, 192, 1, 12, 1, 0), (1, 192, 1, 12, 1, 0), (1, 192, 1, 0, 1, 0), (1, 192, 1, 25, 1, 0), (1, 192, 0, 0, 1, 0)]
            else:
                sweeps = [(1, 192, 1, 12, 1, 1), (1, 192, 1, 0, 1, 1), (1, 192, 1, 6, 1, 1), (1, 192, 1, 25, 1, 1), (1, 256, 1, 12, 1, 1),
                          (1, 192, 1, 12, 0, 1), (1, 192, 0, 0, 1, 1), (1, 192, 1, 12, 1, 2)]
            for items, blk, per, pool, pdl, sim in sweeps:
                env.set_option("k4_items", items); env.set_option("k4_block", blk); env.set_option("k4_persist", per); env.set_option("k4_pool", pool)
                env.set_option("k4_pdl", pdl)
                if world > 1:
                    env.set_option("k4_sim_peers", sim)      # 2 = no barrier wait (timing diagnostic only)
                ms = timed(lambda: env.bounce(d_emit, 100, out=d_tot, want_added=False), reps=3, warm=1)
                us = ms * 10.0
                k4[tag][f"items{items}_block{blk}_persist{per}_pool{pool}_pdl{pdl}_sim{sim}"] = {"us_per_bounce": us, "gbs": (8 * nnz + 40 * (row1 - row0)) / us / 1e3}
                print("K4", tag, items, blk, per, pool, pdl, sim, us, flush=True)
            env.close()
            res["k4_s2"] = k4
            flush()
    flush()
    print(json.dumps(res)[:2000])


if __name__ == "__main__":
    main()
