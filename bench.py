#!/usr/bin/env python
"""bench.py -- headline benchmark of the vrad-b200 hot path (driver contract in the task prompt).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl graft|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One JSON line on rank 0.  Top level = BASELINE.json's first metric, shadow rays/sec:
  step     = one batch of 2^24 shadow segments (C1 workload) through K1 (vrad_test_lines)
             on the S1 box-room map, inputs resident in HBM; N>1: every rank traces its own batch (weak).
  e2e      = the same call with pinned HOST buffers (H2D of the segments + D2H of the bits inside the timing).
  roofline = K1's algorithmic HBM bytes (24.125 B/segment) / CUDA-event time vs the measured HBM peak.
`gather` = BASELINE.json's second metric, bounce-gather iters/sec (C4: S2 multi-room map, 100 bounces per
step through K4, patch rows sharded over the ranks, radiance rows exchanged by peer stores fused into the
kernel (NCCL all-gather fallback); strong scaling), with its own
roofline (8*nnz + 40*N bytes per iteration), e2e and cpu_baseline.
`cpu_baseline` = the CPU oracle (a port: the Go reference cannot be built or run, and its tracer is a stub)
timed on this box's host cores on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

N_SEGMENTS = 1 << 24
N_BOUNCES = 100
METRIC = "shadow_rays_per_sec"
UNIT = "rays/s"


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.proc = None
        self.path = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if not self.proc:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        try:
            for line in open(self.path):
                f = [x.strip() for x in line.split(",")]
                if len(f) < 9:
                    continue
                try:
                    sm.append(float(f[1])); mx.append(float(f[2]))
                except ValueError:
                    continue
                for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                    if val.lower().startswith("active"):
                        reasons.add(name)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            out.update(sm_mhz=statistics.median(sm), sm_max_mhz=max(mx), reasons=sorted(reasons), samples=len(sm))
        return out


def dist_setup(args):
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, local_rank, world


# ------------------------------------------------------------------------------------------------
# reference arm: the CPU oracle on the host cores (bounded sample)
# ------------------------------------------------------------------------------------------------
def cpu_rays(scene, orc, n_sample, threads):
    from vrad_b200 import scenes
    a, b = scenes.shadow_segments(scene, n_sample, seed=0xC0FFEE)
    best = {}
    for mode, name in ((0, "single-ray"), (1, "FourRays packet")):
        orc.test_lines(a[:, :4096].copy(), b[:, :4096].copy(), mode=mode, threads=threads)
        t0 = time.perf_counter()
        orc.test_lines(a, b, mode=mode, threads=threads)
        best[name] = n_sample / (time.perf_counter() - t0)
    return best


def run_reference(args, rank, world):
    """bench.py --impl reference: the reference's CPU path.  The Go reference cannot be built here (no Go
    toolchain, unvendored deps, Trace4Rays is a stub), so this is the oracle port on all host threads."""
    if rank != 0:
        return
    from oracle import pyoracle
    from vrad_b200 import scenes
    threads = pyoracle.num_threads()
    scene = scenes.box_room()
    orc = pyoracle.env_from_scene(scene)
    n_sample = 1 << 21
    a, b = scenes.shadow_segments(scene, n_sample, seed=0xC0FFEE)
    for _ in range(max(1, min(args.warmup, 2))):
        orc.test_lines(a[:, :65536].copy(), b[:, :65536].copy(), mode=0, threads=threads)
    times = []
    for _ in range(args.steps):
        t0 = time.perf_counter()
        orc.test_lines(a, b, mode=0, threads=threads)
        times.append(time.perf_counter() - t0)
    total = sum(times)
    value = n_sample * args.steps / total
    # gather leg on a bounded S2 cut
    s2 = scenes.multi_room(nx=3, ny=2)
    o2 = pyoracle.env_from_scene(s2)
    nnz = o2.build_transfers(s2.pvs, threads=threads)
    emit0 = np.full((s2.n_patches, 3), 100.0, np.float32)
    t0 = time.perf_counter(); o2.bounce(emit0, 20, threads=threads); dt = time.perf_counter() - t0
    sample = f"{n_sample} of the 2^24 C1 shadow segments per step, S1 box room, single-ray kd oracle, OpenMP {threads} threads"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "C1: S1 box room (996 tris, 1325 kd nodes), 2^24 shadow segments per step per GPU via vrad_test_lines (K1)",
                   "segments_per_step_per_gpu": N_SEGMENTS, "reference_arm_sample_per_step": n_sample,
                   "note": "same workload as the graft arm; each reference step traces a bounded sample of it (cpu_baseline.sample)"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gather": {"metric": "bounce_gather_iters_per_sec", "value": 20 / dt, "unit": "iters/s",
                   "config": {"workload": "3x2-room cut of S2", "patches": s2.n_patches, "nnz": nnz}, "cores": threads,
                   "gbs": (8 * nnz + 40 * s2.n_patches) * 20 / dt / 1e9},
        "note": "Go reference not runnable (no toolchain; Trace4Rays stub); this arm is the repo's CPU oracle port",
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# graft arm
# ------------------------------------------------------------------------------------------------
def run_graft(args, rank, local_rank, world):
    import torch
    import torch.distributed as dist
    from vrad_b200 import scenes
    from vrad_b200.environment import Environment, environment_from_scene, row_partition
    from vrad_b200.lib import PinnedArray

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    hbm_peak, peak_src = measured_peaks()
    stream = torch.cuda.current_stream().cuda_stream

    # ---------------- rays: C1 on S1 ----------------
    s1 = scenes.box_room()
    env1 = environment_from_scene(s1, device=local_rank, with_patches=False)
    env1.set_stream(stream)
    a, b = scenes.shadow_segments(s1, N_SEGMENTS, seed=0xC0FFEE + rank)
    nwords = N_SEGMENTS // 32
    d_a, d_b = torch.from_numpy(a).to(dev), torch.from_numpy(b).to(dev)
    d_bits = torch.empty(nwords, dtype=torch.int32, device=dev)
    env1.set_async(True)

    def ray_step():
        env1.test_lines(d_a, d_b, out=d_bits)

    for _ in range(args.warmup):
        ray_step()
    sampler = ClockSampler(local_rank)
    barrier()
    if rank == 0:
        sampler.start()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    barrier()
    ev[0].record()
    for k in range(args.steps):
        ray_step()
        ev[k + 1].record()
    barrier()
    ray_ms_total = max_over_ranks(ev[0].elapsed_time(ev[-1]))
    k1_ms = statistics.mean(ev[k].elapsed_time(ev[k + 1]) for k in range(args.steps))   # one launch per step
    rays_value = world * N_SEGMENTS * args.steps / (ray_ms_total * 1e-3)
    ray_launches = args.steps

    # parity spot check on the bench inputs (outside the timed region): first 2^16 segments vs the oracle
    parity = None
    if rank == 0:
        try:
            from oracle import pyoracle
            orc = pyoracle.env_from_scene(s1, with_patches=False)
            ns = 1 << 16
            ref = orc.test_lines(a[:, :ns].copy(), b[:, :ns].copy(), threads=pyoracle.num_threads())
            parity = bool(np.array_equal(d_bits[: ns // 32].cpu().numpy().view(np.uint32), ref))
        except Exception as exc:  # the checker must never take the bench down
            parity = f"unchecked: {exc}"

    # e2e: pinned host buffers through the same C-ABI call
    env1.set_async(False)
    h_a, h_b = PinnedArray((3, N_SEGMENTS), np.float32), PinnedArray((3, N_SEGMENTS), np.float32)
    h_bits = PinnedArray((nwords,), np.uint32)
    h_a.array[...] = a; h_b.array[...] = b
    e2e_steps = max(2, min(args.steps, 5))
    env1.test_lines(h_a.array, h_b.array, out=h_bits.array)
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        env1.test_lines(h_a.array, h_b.array, out=h_bits.array)
    torch.cuda.synchronize()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    rays_e2e = world * N_SEGMENTS * e2e_steps / e2e_s
    h2d = 2 * 3 * 4 * N_SEGMENTS; d2h = 4 * nwords
    assert np.array_equal(h_bits.array, d_bits.cpu().numpy().view(np.uint32)), "host-buffer path differs from device-buffer path"
    h_a.free(); h_b.free(); h_bits.free()
    env1.close()
    del d_a, d_b

    # ---------------- gather: C4 on S2 ----------------
    s2 = scenes.multi_room()
    env2 = environment_from_scene(s2, device=local_rank, rank=rank, world=world)
    env2.set_stream(stream)
    if world > 1:
        uid = [Environment.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        env2.comm_init(uid[0])
    t0 = time.perf_counter()
    nnz_local = env2.build_transfers(s2.pvs)
    torch.cuda.synchronize()
    k2_cold_s = max_over_ranks(time.perf_counter() - t0)       # first call: cold caches, first-touch allocations
    barrier()
    t0 = time.perf_counter()
    nnz_local = env2.build_transfers(s2.pvs)
    torch.cuda.synchronize()
    k2_s = max_over_ranks(time.perf_counter() - t0)
    k2_ms, k2_launches = env2.last_timing()
    nnz_t = torch.tensor([nnz_local], dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_reduce(nnz_t)
    nnz = int(nnz_t.item())
    N = s2.n_patches
    rng = scenes.SplitMix64(0xE1)
    emit0 = rng.uniform(3 * N, 0.0, 200.0).reshape(N, 3)
    d_emit0 = torch.from_numpy(emit0).to(dev)
    d_total = torch.empty_like(d_emit0)
    env2.set_async(True)

    def gather_step():
        env2.bounce(d_emit0, N_BOUNCES, out=d_total, want_added=False)

    g_steps = max(1, args.steps // 2)
    for _ in range(max(1, args.warmup // 2)):
        gather_step()
    barrier()
    gev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    gev[0].record()
    for _ in range(g_steps):
        gather_step()
    gev[1].record()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    gather_ms = max_over_ranks(gev[0].elapsed_time(gev[1]))
    _, bounce_launches = env2.last_timing()
    iters = N_BOUNCES * g_steps
    gather_value = iters / (gather_ms * 1e-3)
    bytes_per_iter_job = 8 * nnz + 40 * N
    bytes_per_iter_gpu = 8 * nnz_local + 40 * (N // world) + (12 * N if world > 1 else 0)
    # the step also carries init / unpack / reduce launches; per-iteration time attributes them to the gather
    gather_gbs_gpu = bytes_per_iter_gpu * iters / (gather_ms * 1e-3) / 1e9

    # e2e gather: host emit0 in, host total out, every step
    env2.set_async(False)
    h_tot = np.empty_like(emit0)
    env2.bounce(emit0, N_BOUNCES, out=h_tot, want_added=False)
    barrier()
    t0 = time.perf_counter()
    for _ in range(g_steps):
        env2.bounce(emit0, N_BOUNCES, out=h_tot, want_added=False)
    torch.cuda.synchronize()
    gather_e2e_s = max_over_ranks(time.perf_counter() - t0)
    gather_e2e = iters / gather_e2e_s
    energy_ok = bool(np.isfinite(h_tot).all() and h_tot.min() >= 0.0)
    env2.close()

    # ---------------- C5 (S3 outdoor map): informational side numbers, outside every timed region above ----------------
    # rays + direct light on rank 0 at N=1; transfer build + bounce gather at every N (rows sharded by rank)
    large = None
    if not args.no_large:
        s3 = scenes.outdoor()
        env3 = environment_from_scene(s3, device=local_rank, rank=rank, world=world)
        env3.set_stream(stream)
        if world > 1:
            uid3 = [Environment.comm_unique_id() if rank == 0 else None]
            dist.broadcast_object_list(uid3, src=0)
            env3.comm_init(uid3[0])
        st3 = env3.stats()
        large = {"workload": "C5: S3 outdoor map (1,026,540 tris, 2,005,056 luxels/patches, 32x32-cell tile PVS), informational",
                 "kd_build_seconds_host": st3["build_seconds"], "kd_nodes": st3["n_nodes"]}
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        if world == 1:
            env3.set_async(True)
            n3 = 1 << 22
            a3, b3 = scenes.shadow_segments(s3, n3, seed=0xC5)
            d_a3, d_b3 = torch.from_numpy(a3).to(dev), torch.from_numpy(b3).to(dev)
            d_bits3 = torch.empty(n3 // 32, dtype=torch.int32, device=dev)
            for _ in range(2):
                env3.test_lines(d_a3, d_b3, out=d_bits3)
            e0.record()
            for _ in range(3):
                env3.test_lines(d_a3, d_b3, out=d_bits3)
            e1.record(); torch.cuda.synchronize()
            seg_ms = e0.elapsed_time(e1) / 3
            dirs = np.loadtxt(os.path.join(ROOT, "vrad_b200", "data", "anorms.txt"), dtype=np.float32)
            env3.set_sky_dirs(dirs)
            env3.set_async(False)
            d_pos, d_nrm = torch.from_numpy(s3.luxel_pos).to(dev), torch.from_numpy(s3.luxel_normal).to(dev)
            d_rgb = torch.empty((d_pos.shape[0], 3), device=dev)
            env3.direct_light(d_pos, d_nrm, s3.lights, out=d_rgb)
            env3.direct_light(d_pos, d_nrm, s3.lights, out=d_rgb)
            k3_ms, _ = env3.last_timing()
            up = float((dirs @ s3.luxel_normal[::97].T > 0.001).sum()) / s3.luxel_normal[::97].shape[0]
            large.update({"shadow_segments_per_sec": n3 / (seg_ms * 1e-3), "segments": n3, "direct_light_ms": k3_ms,
                          "direct_light_rays_per_sec": d_pos.shape[0] * (1 + up) / (k3_ms * 1e-3),
                          "direct_light_lights": "sun (EMIT_SKYLIGHT) + sky ambient over 162 directions"})
            del d_a3, d_b3, d_pos, d_nrm, d_rgb
        env3.set_async(False)
        barrier()
        t0 = time.perf_counter(); nnz3_local = env3.build_transfers(s3.pvs); torch.cuda.synchronize()
        k2_s3 = max_over_ranks(time.perf_counter() - t0)
        nnz3_t = torch.tensor([nnz3_local], dtype=torch.int64, device=dev)
        if world > 1:
            dist.all_reduce(nnz3_t)
        nnz3 = int(nnz3_t.item())
        N3 = s3.n_patches
        e30 = torch.full((N3, 3), 100.0, device=dev); o30 = torch.empty_like(e30)
        env3.set_async(True)
        env3.bounce(e30, 2, out=o30, want_added=False)
        barrier()
        e0.record(); env3.bounce(e30, 20, out=o30, want_added=False); e1.record()
        barrier()
        k4_ms3 = max_over_ranks(e0.elapsed_time(e1)) / 20
        gpu_bytes3 = 8 * nnz3_local + 40 * (N3 // world) + (12 * N3 if world > 1 else 0)
        large.update({"transfer_build_seconds": k2_s3, "transfers": nnz3, "transfer_bytes": 8 * nnz3,
                      "gather_ms_per_bounce": k4_ms3, "gather_iters_per_sec": 1e3 / k4_ms3,
                      "gather_job_gbs": (8 * nnz3 + 40 * N3) / (k4_ms3 * 1e-3) / 1e9,
                      "gather_frac_of_hbm_peak_per_gpu": gpu_bytes3 / (k4_ms3 * 1e-3) / 1e9 / hbm_peak})
        env3.close()
        del e30, o30

    # ---------------- widened rows, N=1 only, informational, outside every timed region above ----------------
    widened = None
    if world == 1 and not args.no_large:
        try:
            widened = {}
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            # complete TestLineDoesHitSky: static-prop skip, transparent coverage, 3D-skybox recursion
            sk = scenes.sky_room(); m = sk.meta
            g = Environment(local_rank); g.add_triangles(sk.tri_ids, sk.tri_verts, sk.tri_flags); g.set_triangle_colors(m["tri_colors"])
            g.setup_acceleration_structure(); g.bsp_upload(m["bsp"]); g.process_sky_cameras(m["cams_origin"], m["cams_scale"])
            g.set_stream(stream); g.set_async(True)
            ns = 1 << 23
            sa, sb = scenes.sky_segments(sk, ns)
            d_sa, d_sb = torch.from_numpy(sa).to(dev), torch.from_numpy(sb).to(dev)
            d_fv = torch.empty(ns, dtype=torch.float32, device=dev)
            sky = {"workload": "S4 sky room (732 tris, 48 transparent, 2 sky cameras), 2^23 segments via vrad_test_lines_sky", "segments": ns}
            for flags, name in ((0, "sky_id_only"), (1, "with_skybox_recursion"), (3, "recursion_and_texture_shadows")):
                for _ in range(2):
                    g.test_lines_sky(d_sa, d_sb, flags, 7, out=d_fv)
                e0.record()
                for _ in range(3):
                    g.test_lines_sky(d_sa, d_sb, flags, 7, out=d_fv)
                e1.record(); torch.cuda.synchronize()
                sky[name + "_segments_per_sec"] = ns / (e0.elapsed_time(e1) / 3 * 1e-3)
            g.close(); del d_sa, d_sb, d_fv
            widened["test_line_does_hit_sky"] = sky
            # patch hierarchy: the C4 map with SubdividePatches trees, hierarchical transfers + CollectLight
            hs = scenes.multi_room_hier(nx=12, ny=11); tr = hs.meta["tree"]
            h = environment_from_scene(hs, device=local_rank)
            h.set_hierarchy(tr["parent"], tr["child1"], tr["child2"], tr["face"])
            h.set_stream(stream)
            h.build_transfers(hs.pvs)
            t0 = time.perf_counter(); hn = h.build_transfers(hs.pvs); torch.cuda.synchronize(); hk2 = time.perf_counter() - t0
            hk2_ms, _ = h.last_timing()
            hN = hs.n_patches
            he = torch.full((hN, 3), 100.0, device=dev); ho = torch.empty_like(he)
            h.set_async(True)
            h.bounce(he, 10, out=ho, want_added=False)
            e0.record(); h.bounce(he, 100, out=ho, want_added=False); e1.record(); torch.cuda.synchronize()
            hms = e0.elapsed_time(e1) / 100
            widened["patch_hierarchy"] = {
                "workload": "C4 map, patches from vrad_patches_subdivide (roots + interior + leaves), hierarchical vrad_build_transfers, vrad_bounce with CollectLight",
                "patches": hN, "leaf_patches": int((tr["child1"] == -1).sum()), "transfers": hn, "transfer_build_seconds": hk2, "transfer_build_kernel_ms": hk2_ms,
                "ms_per_bounce": hms, "iters_per_sec": 1e3 / hms, "gbs": (8 * hn + 40 * hN) / (hms * 1e-3) / 1e9,
                "flat_map_transfers": nnz, "flat_map_ms_per_bounce": gather_ms / iters}
            h.close(); del he, ho
        except Exception as exc:  # informational only: never take the bench line down
            widened = {"error": repr(exc)}

    # ---------------- BSP side: K5, the file-driven bake, the binned kd build; N=1, informational ----------------
    # Runs in a child process (tools/bsp_side_bench.py): these paths had not yet run on a GPU when this was written, so nothing they do
    # -- an exception, a sticky CUDA error, a crash, a hang -- may touch the numbers above.
    bsp_side = None
    if world == 1 and not args.no_large:
        try:
            r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "bsp_side_bench.py"), "--device", str(local_rank), "--hbm-peak", str(hbm_peak)],
                               capture_output=True, text=True, timeout=240)
            lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
            bsp_side = json.loads(lines[-1]) if lines else {"error": f"exit {r.returncode}", "stderr": r.stderr[-800:]}
        except Exception as exc:  # informational only: never take the bench line down
            bsp_side = {"error": repr(exc)}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---------------- cpu baseline (rank 0, N=1 only) ----------------
    cpu_rays_obj, cpu_gather_obj = None, None
    if world == 1 and not args.no_cpu:
        from oracle import pyoracle
        threads = pyoracle.num_threads()
        orc = pyoracle.env_from_scene(s1, with_patches=False)
        n_sample = 1 << 22
        allc = cpu_rays(s1, orc, n_sample, threads)
        one = cpu_rays(s1, orc, n_sample // 8, 1)
        best_name = max(allc, key=allc.get)
        cpu_rays_obj = {"value": allc[best_name], "unit": UNIT, "cores": threads, "kind": "port",
                        "sample": f"{n_sample} of the 2^24 C1 segments, oracle {best_name} tracer, OpenMP {threads} threads",
                        "single_thread": {k: v for k, v in one.items()}, "all_cores": allc}
        s2c = scenes.multi_room(nx=3, ny=2)
        o2 = pyoracle.env_from_scene(s2c)
        nnz_c = o2.build_transfers(s2c.pvs, threads=threads)
        e0 = np.full((s2c.n_patches, 3), 100.0, np.float32)
        o2.bounce(e0, 2, threads=threads)
        t0 = time.perf_counter(); o2.bounce(e0, 50, threads=threads); dt = time.perf_counter() - t0
        t1 = time.perf_counter(); o2.bounce(e0, 10, threads=1); dt1 = time.perf_counter() - t1
        scale = (8 * nnz_c + 40 * s2c.n_patches) / bytes_per_iter_job     # bytes ratio sample/full
        cpu_gather_obj = {"value": 50 / dt * scale, "unit": "iters/s", "cores": threads, "kind": "port",
                          "sample": f"3x2-room cut of S2 (N={s2c.n_patches}, nnz={nnz_c}), 50 bounces, scaled by bytes to the full map",
                          "gbs": (8 * nnz_c + 40 * s2c.n_patches) * 50 / dt / 1e9,
                          "single_thread_gbs": (8 * nnz_c + 40 * s2c.n_patches) * 10 / dt1 / 1e9}

    k1_bytes = 24.125 * N_SEGMENTS
    k1_gbs = k1_bytes / (k1_ms * 1e-3) / 1e9
    line = {
        "metric": METRIC, "value": rays_value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ray_ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": "C1: S1 box room (996 tris, 1325 kd nodes), 2^24 shadow segments per step per GPU via vrad_test_lines (K1)",
                   "segments_per_step_per_gpu": N_SEGMENTS, "l2": "inputs 403 MB per step > 126 MB L2", "parity_checked": parity},
        "e2e": {"value": rays_e2e, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "steps": e2e_steps},
        "gpu_launches": ray_launches,
        "roofline": {"bound": "hbm", "achieved": k1_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": k1_gbs / hbm_peak,
                     "traffic": 425.8e6, "traffic_source": "profiles/r01_final_k1_ncu_summary.txt (dram read+write per launch)",
                     "kernel": "k1_test_lines", "peak_source": peak_src,
                     "note": "K1 is issue/divergence-bound by design (SURVEY 8d: HBM fraction ~1%); the binding target is >=1e9 rays/s. "
                             "The HBM-bound kernel of this path is k4_gather: see gather.roofline"},
        "cpu_baseline": cpu_rays_obj,
        "clocks": clocks,
        "large_scene": large,
        "widened": widened,
        "bsp_side": bsp_side,
        "gather": {
            "metric": "bounce_gather_iters_per_sec", "value": gather_value, "unit": "iters/s", "scaling": "strong",
            "steps": g_steps, "bounces_per_step": N_BOUNCES, "ms_per_iter": gather_ms / iters,
            "config": {"workload": "C4: S2 multi-room (49,586 tris), 100 forced bounces per step via vrad_bounce (K4), rows sharded by rank",
                       "patches": N, "nnz": nnz, "l2": f"transfer stream {8 * nnz / 1e6:.0f} MB per iteration > 126 MB L2" if 8 * nnz // world > 126e6 else
                       f"per-GPU transfer stream {8 * nnz / world / 1e6:.0f} MB per iteration fits L2"},
            "e2e": {"value": gather_e2e, "unit": "iters/s", "h2d_bytes_per_step": 12 * N, "d2h_bytes_per_step": 12 * N},
            "gpu_launches": bounce_launches * g_steps,
            "roofline": {"bound": "hbm", "achieved": gather_gbs_gpu, "peak": hbm_peak, "unit": "GB/s", "frac": gather_gbs_gpu / hbm_peak,
                         "traffic": 1.5537e9 if world == 1 else None,
                         "traffic_source": "profiles/r01_final_k2_k3_k4_ncu_summary.txt (dram read+write per launch, N=1)",
                         "kernel": "k4_gather", "bytes_per_iter_per_gpu": bytes_per_iter_gpu, "peak_source": peak_src,
                         "exchange": "none" if world == 1 else "peer stores into every rank's next-bounce buffer (NVLink, CUDA IPC) + epoch barrier; NCCL all-gather fallback"},
            "job_gbs": bytes_per_iter_job * iters / (gather_ms * 1e-3) / 1e9,
            "cpu_baseline": cpu_gather_obj,
            "transfer_build": {"wall_s": k2_s, "first_call_wall_s": k2_cold_s, "kernel_ms": k2_ms, "launches": k2_launches, "nnz_local_rank0": nnz_local},
            "finite_nonnegative": energy_ok,
        },
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="graft", choices=["graft", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-large", action="store_true", help="skip the informational C5 (S3 map) side numbers")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "graft" else args.warmup
    rank, local_rank, world = dist_setup(args)
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if world != args.gpus and rank == 0 and world == 1 and args.gpus > 1:
        print(json.dumps({"error": f"--gpus {args.gpus} needs torchrun with {args.gpus} ranks (WORLD_SIZE={world})"}))
        sys.exit(2)
    run_graft(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
