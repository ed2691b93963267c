#!/usr/bin/env python
"""bench.py -- headline benchmark of the vrad-b200 hot path (driver contract in the task prompt).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl graft|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One JSON line on rank 0.  BASELINE.json's metric has two halves, "shadow rays/sec and bounce-gather iters/sec at 1/2/4/8 B200":

  top level = bounce-gather iters/sec (C4): the HBM-bound kernel of the path and the only stage with a per-iteration exchange,
      so the one whose scaling can fail.  step = one vrad_bounce call of 100 forced bounces over the resident transfer lists of
      the S2 multi-room map (K4); patch rows are sharded over the ranks, so N ranks do the SAME job (strong scaling);
      value = 100 * K / time.  e2e = the same call with HOST emit0 in and HOST total out every step.  roofline = K4's
      algorithmic HBM bytes per GPU per iteration (8 * nnz_local + 40 * N / world, + 12 * N received when world > 1) / time
      against the measured HBM peak.
  "rays" = shadow rays/sec (C1): 2^24 shadow segments per step through K1 (vrad_test_lines) on the S1 box room, inputs resident
      in HBM; with N ranks every rank traces its own batch (weak scaling, no communication).  rays.e2e = the same segments from
      pinned HOST buffers, as index pairs into the resident point table (vrad_test_lines_indexed, 8 B per segment) and as
      coordinates (vrad_test_lines, 24 B per segment).

`cpu_baseline` / `--impl reference`: the CPU oracle (a port: the Go reference cannot be built or run here -- no toolchain, unvendored
dependencies, and its tracer is a stub) timed on this box's host cores on a bounded sample of the same workloads.
Correctness is checked where the numbers are taken, at every N, outside the timed regions: gather.parity_checked, rays.parity_checked.
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
METRIC = "bounce_gather_iters_per_sec"
UNIT = "iters/s"
RAYS_METRIC = "shadow_rays_per_sec"
RAYS_UNIT = "rays/s"
# identical in both arms (the driver compares the configs of the two lines)
CONFIG = {
    "workload": "C4: S2 multi-room map (49,586 tris, 187,328 patches, 191,917,752 transfers), 100 forced bounces per step via vrad_bounce (K4), "
                "patch rows sharded by rank, radiance rows exchanged every bounce",
    "patches": 187328, "bounces_per_step": N_BOUNCES,
    "rays_workload": "C1: S1 box room (996 tris, 1325 kd nodes), 2^24 shadow segments per step per GPU via vrad_test_lines (K1)",
    "l2": "transfer stream 1.01 GB per iteration as block rows (126 MB per GPU at 8 GPUs -- the size of L2, of which the radiance and the other traffic take their share; 1535 / 192 MB as {col,w} pairs); ray inputs 403 MB per step > L2",
}
CPU_GATHER_CUT = (5, 4)          # rooms of the S2 cut the CPU arm bounces (29 M transfers, 234 MB: larger than the host's caches)
CPU_GATHER_BOUNCES = 50
CPU_RAY_SAMPLE = 1 << 21


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def host_threads() -> int:
    """Threads for the CPU arm: the cores this process may run on.  torchrun exports OMP_NUM_THREADS=1; the oracle's
    entry points take the thread count as an argument (an OpenMP num_threads clause), so the variable does not apply."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


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


def dist_setup():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))


# ------------------------------------------------------------------------------------------------
# CPU legs: the oracle on the host cores, bounded samples (used by --impl reference and by cpu_baseline)
# ------------------------------------------------------------------------------------------------
class CpuGather:
    """The bounce gather of a cut of S2 on the CPU oracle; rates are reported as iterations/s of the FULL C4 matrix (scaled by bytes)."""

    def __init__(self, threads):
        from oracle import pyoracle
        from vrad_b200 import scenes
        self.threads = threads
        nx, ny = CPU_GATHER_CUT
        self.scene = scenes.multi_room(nx=nx, ny=ny)
        self.orc = pyoracle.env_from_scene(self.scene)
        t0 = time.perf_counter()
        self.nnz = self.orc.build_transfers(self.scene.pvs, threads=threads)
        self.build_s = time.perf_counter() - t0
        self.N = self.scene.n_patches
        self.bytes_per_iter = 8 * self.nnz + 40 * self.N
        self.emit0 = np.full((self.N, 3), 100.0, np.float32)

    def step(self, threads=None):
        t0 = time.perf_counter()
        self.orc.bounce(self.emit0, CPU_GATHER_BOUNCES, threads=threads or self.threads)
        return time.perf_counter() - t0

    def describe(self, threads=None):
        nx, ny = CPU_GATHER_CUT
        return (f"{nx}x{ny}-room cut of S2 (N={self.N}, nnz={self.nnz}, {self.bytes_per_iter / 1e6:.0f} MB per iteration), {CPU_GATHER_BOUNCES} bounces per step, "
                f"oracle CSR gather, OpenMP {threads or self.threads} threads; iterations/s scaled by bytes to the full C4 matrix")


def full_bytes_per_iter():
    return 8 * 191917752 + 40 * CONFIG["patches"]


def run_reference(args, rank, world):
    """bench.py --impl reference: the reference's CPU path.  The Go reference cannot be built here (no Go toolchain, unvendored deps,
    Trace4Rays is a stub, the bounce step is commented out), so this is the oracle port on all the host threads this process may use."""
    if rank != 0:
        return
    from oracle import pyoracle
    from vrad_b200 import scenes
    threads = host_threads()
    g = CpuGather(threads)
    for _ in range(max(1, min(args.warmup, 2))):
        g.step()
    times = [g.step() for _ in range(args.steps)]
    total = sum(times)
    gbs = g.bytes_per_iter * CPU_GATHER_BOUNCES * args.steps / total / 1e9
    value = gbs * 1e9 / full_bytes_per_iter()
    one_gbs = g.bytes_per_iter * CPU_GATHER_BOUNCES / g.step(threads=1) / 1e9
    # rays leg
    s1 = scenes.box_room()
    orc = pyoracle.env_from_scene(s1, with_patches=False)
    a, b = scenes.shadow_segments(s1, CPU_RAY_SAMPLE, seed=0xC0FFEE)
    orc.test_lines(a[:, :65536].copy(), b[:, :65536].copy(), mode=0, threads=threads)
    t0 = time.perf_counter()
    for _ in range(3):
        orc.test_lines(a, b, mode=0, threads=threads)
    rays = 3 * CPU_RAY_SAMPLE / (time.perf_counter() - t0)
    t0 = time.perf_counter()
    orc.test_lines(a[:, :1 << 18].copy(), b[:, :1 << 18].copy(), mode=0, threads=1)
    rays_one = (1 << 18) / (time.perf_counter() - t0)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": CONFIG,
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": g.describe(), "gbs": gbs,
                         "single_thread": {"value": one_gbs * 1e9 / full_bytes_per_iter(), "gbs": one_gbs}, "transfer_build_seconds": g.build_s},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "rays": {"metric": RAYS_METRIC, "value": rays, "unit": RAYS_UNIT, "cores": threads, "kind": "port",
                 "sample": f"{CPU_RAY_SAMPLE} of the 2^24 C1 shadow segments, S1 box room, single-ray kd oracle, OpenMP {threads} threads",
                 "single_thread": rays_one},
        "note": "Go reference not runnable (no toolchain; Trace4Rays is a stub; the bounce step is commented out); this arm is the repo's CPU oracle port",
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# graft arm
# ------------------------------------------------------------------------------------------------
def run_graft(args, rank, local_rank, world):
    import torch
    import torch.distributed as dist
    from vrad_b200 import scenes
    from vrad_b200.environment import Environment, environment_from_scene
    from vrad_b200.lib import PinnedArray

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def reduce_scalar(x: float, op) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=op)
        return float(t.item())

    def max_over_ranks(x: float) -> float:
        return reduce_scalar(x, dist.ReduceOp.MAX) if world > 1 else x

    hbm_peak, peak_src = measured_peaks()
    stream = torch.cuda.current_stream().cuda_stream
    threads = host_threads()
    sampler = ClockSampler(local_rank)

    # =================================== gather: C4 on S2 (top level) ===================================
    s2 = scenes.multi_room()
    N = s2.n_patches
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
    nnz = int(reduce_scalar(float(nnz_local), dist.ReduceOp.SUM)) if world > 1 else nnz_local
    row0, row1, _ = env2.transfers_info()
    emit0 = scenes.SplitMix64(0xE1).uniform(3 * N, 0.0, 200.0).reshape(N, 3)
    d_emit0 = torch.from_numpy(emit0).to(dev)
    d_total = torch.empty_like(d_emit0)
    env2.set_async(True)

    def gather_step():
        env2.bounce(d_emit0, N_BOUNCES, out=d_total, want_added=False)

    for _ in range(args.warmup):
        gather_step()
    barrier()
    if rank == 0:
        sampler.start()
    gev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    barrier()
    gev[0].record()
    for _ in range(args.steps):
        gather_step()
    gev[1].record()
    barrier()
    gather_ms = max_over_ranks(gev[0].elapsed_time(gev[1]))
    _, bounce_launches = env2.last_timing()
    iters = N_BOUNCES * args.steps
    gather_value = iters / (gather_ms * 1e-3)
    bytes_per_iter_job = 8 * nnz + 40 * N
    bytes_per_iter_gpu = 8 * nnz_local + 40 * (row1 - row0) + (12 * N if world > 1 else 0)
    bytes_gpu_max = max_over_ranks(float(bytes_per_iter_gpu))
    # what the gather actually streams: 6 B per packed entry (f32 weight + u16 column offset; segments padded to 64) when the rows are packed
    pair_entries, packed_entries, packed_segments, block_entries, block_rows = env2.transfers_layout()
    moved_per_iter_gpu = ((2 + 4 * block_rows) * block_entries if block_entries else 6 * packed_entries if packed_entries else 8 * pair_entries) + 40 * (row1 - row0) + (12 * N if world > 1 else 0)
    moved_gpu_max = max_over_ranks(float(moved_per_iter_gpu))
    # the step also carries init / unpack / reduce launches; per-iteration time attributes them to the gather
    gather_gbs_gpu = bytes_gpu_max * iters / (gather_ms * 1e-3) / 1e9
    total_sharded = d_total.cpu().numpy()

    # e2e gather: host emit0 in, host total out, every step
    env2.set_async(False)
    h_tot = np.empty_like(emit0)
    env2.bounce(emit0, N_BOUNCES, out=h_tot, want_added=False)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        env2.bounce(emit0, N_BOUNCES, out=h_tot, want_added=False)
    torch.cuda.synchronize()
    gather_e2e_s = max_over_ranks(time.perf_counter() - t0)
    gather_e2e = iters / gather_e2e_s

    # ---- parity of the gather, at this N, outside the timing (the oracle is the checker, never the thing measured) ----
    # (1) 1 bounce: rows sampled from EVERY rank's block against the oracle's own single-row transfer builder + gather;
    # (2) 2 bounces: the second bounce of a sampled row reads the first bounce's rows of all ranks -> checks the exchange;
    # (3) N > 1: the sharded 100-bounce result against a single-GPU run of the same job on rank 0's device.
    gparity = {"finite_nonnegative": bool(np.isfinite(h_tot).all() and h_tot.min() >= 0.0),
               "e2e_equals_device_run": bool(np.array_equal(h_tot, total_sharded))}
    t1g, _, _ = env2.bounce(emit0, 1)
    t2g, _, _ = env2.bounce(emit0, 2)
    bounds = [None] * world
    if world > 1:
        dist.all_gather_object(bounds, (row0, row1))
    else:
        bounds = [(row0, row1)]
    if rank == 0:
        try:
            from oracle import pyoracle
            orc2 = pyoracle.env_from_scene(s2)
            rng = scenes.SplitMix64(0x51)
            rows = []
            for a0, a1 in bounds:
                if a1 > a0:
                    rows += [int(a0), int(a1 - 1)] + [int(a0 + r) for r in rng.integers(4, a1 - a0)]
            refl = s2.patch_refl.astype(np.float32)
            sky = (s2.patch_flags & 1).astype(bool) if s2.patch_flags is not None else np.zeros(N, bool)
            er0 = np.where(sky[:, None], 0.0, emit0 * refl).astype(np.float32)
            er1 = np.where(sky[:, None], 0.0, t1g * refl).astype(np.float32)
            e1m, e2m = 0.0, 0.0
            scale1, scale2 = float(np.abs(t1g).max()), float(np.abs(t2g).max())
            for i in rows:
                col, w = orc2.transfer_row(i, s2.pvs)
                want1 = (w[:, None].astype(np.float64) * er0[col]).sum(axis=0) if not sky[i] else np.zeros(3)
                want2 = want1 + ((w[:, None].astype(np.float64) * er1[col]).sum(axis=0) if not sky[i] else 0.0)
                e1m = max(e1m, float(np.abs(t1g[i] - want1).max()) / scale1)
                e2m = max(e2m, float(np.abs(t2g[i] - want2).max()) / scale2)
            gparity.update({"oracle_rows_checked": len(rows), "rows_per_rank_block": 6, "max_rel_err_1_bounce": e1m, "max_rel_err_2_bounces": e2m,
                            "tolerance": 1e-4, "ok": bool(e1m <= 1e-4 and e2m <= 1e-4)})
        except Exception as exc:  # the checker must never take the bench down
            gparity["oracle_rows_checked"] = f"unchecked: {exc!r}"
    env2.close()
    if world > 1:
        barrier()
        if rank == 0:
            try:
                solo = environment_from_scene(s2, device=local_rank)
                solo.set_stream(stream)
                assert solo.build_transfers(s2.pvs) == nnz
                t_solo, _, _ = solo.bounce(emit0, N_BOUNCES)
                solo.close()
                rel = float(np.abs(total_sharded - t_solo).max() / np.abs(t_solo).max())
                gparity.update({"vs_single_gpu_100_bounces_max_rel": rel, "nnz_equals_single_gpu": True})
                gparity["ok"] = bool(gparity.get("ok", False) and rel <= 1e-4)
            except Exception as exc:
                gparity["vs_single_gpu_100_bounces_max_rel"] = f"unchecked: {exc!r}"
        barrier()

    # =================================== rays: C1 on S1 ===================================
    s1 = scenes.box_room()
    env1 = environment_from_scene(s1, device=local_rank, with_patches=False)
    env1.set_stream(stream)
    a, b = scenes.shadow_segments(s1, N_SEGMENTS, seed=0xC0FFEE + rank)
    pts, pairs = scenes.shadow_segment_indices(s1, N_SEGMENTS, seed=0xC0FFEE + rank)
    env1.points_upload(pts)
    nwords = N_SEGMENTS // 32
    d_a, d_b = torch.from_numpy(a).to(dev), torch.from_numpy(b).to(dev)
    d_bits = torch.empty(nwords, dtype=torch.int32, device=dev)
    env1.set_async(True)

    def ray_step():
        env1.test_lines(d_a, d_b, out=d_bits)

    for _ in range(args.warmup):
        ray_step()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    barrier()
    ev[0].record()
    for k in range(args.steps):
        ray_step()
        ev[k + 1].record()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    ray_ms_total = max_over_ranks(ev[0].elapsed_time(ev[-1]))
    k1_ms = statistics.mean(ev[k].elapsed_time(ev[k + 1]) for k in range(args.steps))
    rays_value = world * N_SEGMENTS * args.steps / (ray_ms_total * 1e-3)
    _, ray_launches = env1.last_timing()
    # the unordered kernel alone (one launch per step), for the roofline line and the comparison with round 1
    env1.set_option("k1_sort", 0)
    for _ in range(2):
        ray_step()
    ev2 = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    ev2[0].record()
    for _ in range(3):
        ray_step()
    ev2[1].record(); torch.cuda.synchronize()
    k1_unsorted_ms = ev2[0].elapsed_time(ev2[1]) / 3
    env1.set_option("k1_sort", -1)

    # parity on the bench inputs (outside the timed region): EVERY rank checks its first 2^16 segments against the oracle
    rparity = {}
    try:
        from oracle import pyoracle
        orc = pyoracle.env_from_scene(s1, with_patches=False)
        ns = 1 << 16
        ref = orc.test_lines(a[:, :ns].copy(), b[:, :ns].copy(), threads=max(1, threads // world))
        ray_step(); torch.cuda.synchronize()
        mine = bool(np.array_equal(d_bits[: ns // 32].cpu().numpy().view(np.uint32), ref))
        rparity = {"segments_checked_per_rank": ns, "ranks_checked": world, "bit_exact": bool(reduce_scalar(1.0 if mine else 0.0, dist.ReduceOp.MIN) == 1.0) if world > 1 else mine}
    except Exception as exc:  # the checker must never take the bench down
        rparity = {"bit_exact": f"unchecked: {exc!r}"}
    bits_dev = d_bits.cpu().numpy().view(np.uint32)

    # e2e: pinned host buffers through the C-ABI -- index pairs (8 B per segment) and coordinates (24 B per segment)
    env1.set_async(False)
    h_bits = PinnedArray((nwords,), np.uint32)
    e2e_steps = max(2, min(args.steps, 5))

    def host_rate(fn):
        fn()
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            fn()
        torch.cuda.synchronize()
        return world * N_SEGMENTS * e2e_steps / max_over_ranks(time.perf_counter() - t0)

    h_p = PinnedArray((N_SEGMENTS, 2), np.int32); h_p.array[...] = pairs
    rays_e2e_idx = host_rate(lambda: env1.test_lines_indexed(h_p.array, out=h_bits.array))
    idx_same = bool(np.array_equal(h_bits.array, bits_dev))
    h_p.free()
    h_a, h_b = PinnedArray((3, N_SEGMENTS), np.float32), PinnedArray((3, N_SEGMENTS), np.float32)
    h_a.array[...] = a; h_b.array[...] = b
    rays_e2e_xyz = host_rate(lambda: env1.test_lines(h_a.array, h_b.array, out=h_bits.array))
    xyz_same = bool(np.array_equal(h_bits.array, bits_dev))
    rparity["host_paths_equal_device_path"] = idx_same and xyz_same
    h_a.free(); h_b.free(); h_bits.free()
    env1.close()
    del d_a, d_b

    # ---------------- C5 (S3 outdoor map): informational side numbers, outside every timed region above ----------------
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
            seg = {}
            for name, sort in (("ordered", -1), ("unordered", 0)):
                env3.set_option("k1_sort", sort)
                for _ in range(2):
                    env3.test_lines(d_a3, d_b3, out=d_bits3)
                e0.record()
                for _ in range(3):
                    env3.test_lines(d_a3, d_b3, out=d_bits3)
                e1.record(); torch.cuda.synchronize()
                seg[name] = n3 / (e0.elapsed_time(e1) / 3 * 1e-3)
            env3.set_option("k1_sort", -1)
            dirs = np.loadtxt(os.path.join(ROOT, "vrad_b200", "data", "anorms.txt"), dtype=np.float32)
            env3.set_sky_dirs(dirs)
            env3.set_async(False)
            d_pos, d_nrm = torch.from_numpy(s3.luxel_pos).to(dev), torch.from_numpy(s3.luxel_normal).to(dev)
            d_rgb = torch.empty((d_pos.shape[0], 3), device=dev)
            env3.direct_light(d_pos, d_nrm, s3.lights, out=d_rgb)
            env3.direct_light(d_pos, d_nrm, s3.lights, out=d_rgb)
            k3_ms, _ = env3.last_timing()
            up = float((dirs @ s3.luxel_normal[::97].T > 0.001).sum()) / s3.luxel_normal[::97].shape[0]
            large.update({"shadow_segments_per_sec": seg["ordered"], "shadow_segments_per_sec_unordered_kernel": seg["unordered"], "segments": n3,
                          "segments_workload": "random patch->patch / patch->sun segments across the whole map, exact (reference-recipe) kd tree",
                          "direct_light_ms": k3_ms, "direct_light_rays_per_sec": d_pos.shape[0] * (1 + up) / (k3_ms * 1e-3),
                          "direct_light_lights": "sun (EMIT_SKYLIGHT) + sky ambient over 162 directions"})
            del d_a3, d_b3, d_pos, d_nrm, d_rgb
        env3.set_async(False)
        barrier()
        t0 = time.perf_counter(); nnz3_local = env3.build_transfers(s3.pvs); torch.cuda.synchronize()
        k2_s3 = max_over_ranks(time.perf_counter() - t0)
        nnz3 = int(reduce_scalar(float(nnz3_local), dist.ReduceOp.SUM)) if world > 1 else nnz3_local
        N3 = s3.n_patches
        r30, r31, _ = env3.transfers_info()
        e30 = torch.full((N3, 3), 100.0, device=dev); o30 = torch.empty_like(e30)
        env3.set_async(True)
        env3.bounce(e30, 2, out=o30, want_added=False)
        barrier()
        e0.record(); env3.bounce(e30, 20, out=o30, want_added=False); e1.record()
        barrier()
        k4_ms3 = max_over_ranks(e0.elapsed_time(e1)) / 20
        gpu_bytes3 = max_over_ranks(float(8 * nnz3_local + 40 * (r31 - r30) + (12 * N3 if world > 1 else 0)))
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
        orc = pyoracle.env_from_scene(s1, with_patches=False)
        n_sample = 1 << 22
        ca, cb = scenes.shadow_segments(s1, n_sample, seed=0xC0FFEE)
        allc, one = {}, {}
        for mode, name in ((0, "single-ray"), (1, "FourRays packet")):
            orc.test_lines(ca[:, :4096].copy(), cb[:, :4096].copy(), mode=mode, threads=threads)
            t0 = time.perf_counter(); orc.test_lines(ca, cb, mode=mode, threads=threads); allc[name] = n_sample / (time.perf_counter() - t0)
            t0 = time.perf_counter(); orc.test_lines(ca[:, :n_sample // 8].copy(), cb[:, :n_sample // 8].copy(), mode=mode, threads=1)
            one[name] = (n_sample // 8) / (time.perf_counter() - t0)
        best_name = max(allc, key=allc.get)
        cpu_rays_obj = {"value": allc[best_name], "unit": RAYS_UNIT, "cores": threads, "kind": "port",
                        "sample": f"{n_sample} of the 2^24 C1 segments, oracle {best_name} tracer, OpenMP {threads} threads",
                        "single_thread": one, "all_cores": allc}
        cg = CpuGather(threads)
        cg.step()
        dt = cg.step()
        dt1 = cg.step(threads=1)
        gbs = cg.bytes_per_iter * CPU_GATHER_BOUNCES / dt / 1e9
        cpu_gather_obj = {"value": gbs * 1e9 / bytes_per_iter_job, "unit": UNIT, "cores": threads, "kind": "port", "sample": cg.describe(), "gbs": gbs,
                          "single_thread_gbs": cg.bytes_per_iter * CPU_GATHER_BOUNCES / dt1 / 1e9, "transfer_build_seconds": cg.build_s}

    k1_bytes = 24.125 * N_SEGMENTS
    k1_gbs = k1_bytes / (k1_unsorted_ms * 1e-3) / 1e9
    cfg = dict(CONFIG)
    line = {
        "metric": METRIC, "value": gather_value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": gather_ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "config": cfg,
        "e2e": {"value": gather_e2e, "unit": UNIT, "h2d_bytes_per_step": 12 * N, "d2h_bytes_per_step": 12 * N, "steps": args.steps,
                "note": "host emit0 (N x RGB f32) in and host total out on every 100-bounce call; transfer lists and patches stay resident, as in the bake"},
        "gpu_launches": bounce_launches * args.steps,
        "roofline": {"bound": "hbm", "achieved": gather_gbs_gpu, "peak": hbm_peak, "unit": "GB/s", "frac": gather_gbs_gpu / hbm_peak,
                     "traffic": (1.0260e9 if block_entries else 1.2047e9 if packed_entries else 1.5537e9) if world == 1 else None,
                     "traffic_source": "profiles/r02_k4_packed_ncu_summary.txt / profiles/r02_k4_block_ncu_summary.txt (dram read+write per launch, N=1; same kernels)",
                     "kernel": (f"k4_gather_blocked<{block_rows}, 4, {4 if block_rows == 4 else 5}>" if block_entries else "k4_gather_packed<4, 5>" if packed_entries else "k4_gather") if world == 1 else
                               (f"k4_gather_items_blocked<true, 8, {block_rows}>" if block_entries else "k4_gather_items<true, 8, true, PACKED>" if packed_entries else "k4_gather_items<true, 6>"),
                     "bytes_per_iter_per_gpu": bytes_gpu_max, "peak_source": peak_src,
                     "note": "achieved / frac are quoted on the ALGORITHMIC bytes of SURVEY 8(d): 8 B per transfer (the reference's Transfer struct) + 40 B per row. "
                             "The gather reads the transfers in a denser form -- block rows: 4 consecutive rows share the union of their columns, 18 B per union entry "
                             "(u16 column offset + 4 f32 weights, bit-exact), about 5.3 B per transfer on this map; or packed 6-byte entries -- so it moves fewer bytes "
                             "than that and frac can exceed what the memory system delivers: moved_* is the same rate on the bytes actually streamed",
                     "stream_format": (f"block rows: u16 column offset + {block_rows} f32 weights per entry of the union of {block_rows} rows' columns, {2 + 4 * block_rows} B" if block_entries else
                                       "packed: f32 weight + u16 column offset, 6 B per transfer" if packed_entries else "{col:int32, w:f32} pairs, 8 B per transfer"),
                     "block_entries_rank0": block_entries, "block_rows": block_rows,
                     "packed_entries_rank0": packed_entries, "packed_segments_rank0": packed_segments, "pair_entries_rank0": pair_entries,
                     "moved_bytes_per_iter_per_gpu": moved_gpu_max, "moved_achieved": moved_gpu_max * iters / (gather_ms * 1e-3) / 1e9,
                     "moved_frac": moved_gpu_max * iters / (gather_ms * 1e-3) / 1e9 / hbm_peak,
                     "exchange": "none" if world == 1 else "peer stores into every rank's next-bounce buffer (NVLink, CUDA IPC) + in-kernel epoch barrier; NCCL all-gather fallback"},
        "cpu_baseline": cpu_gather_obj,
        "clocks": clocks,
        "ms_per_iter": gather_ms / iters,
        "job_gbs": bytes_per_iter_job * iters / (gather_ms * 1e-3) / 1e9,
        "nnz": nnz, "nnz_local_rank0": nnz_local, "row_blocks": bounds,
        "parity_checked": gparity,
        "transfer_build": {"wall_s": k2_s, "first_call_wall_s": k2_cold_s, "kernel_ms": k2_ms, "launches": k2_launches},
        "rays": {
            "metric": RAYS_METRIC, "value": rays_value, "unit": RAYS_UNIT, "scaling": "weak", "steps": args.steps, "ms_per_step": ray_ms_total / args.steps,
            "segments_per_step_per_gpu": N_SEGMENTS, "gpu_launches": ray_launches * args.steps,
            "order": "segments are ordered by start cell / direction octant / end cell before the traversal (key pass + radix sort inside the timed region)",
            "unordered_kernel": {"ms": k1_unsorted_ms, "value": N_SEGMENTS / (k1_unsorted_ms * 1e-3)},
            "e2e": {"value": rays_e2e_idx, "unit": RAYS_UNIT, "h2d_bytes_per_step": 8 * N_SEGMENTS, "d2h_bytes_per_step": 4 * nwords, "steps": e2e_steps,
                    "form": "vrad_test_lines_indexed: {int32 start, int32 stop} pairs into the resident point table (patch origins + light origins)",
                    "coordinates": {"value": rays_e2e_xyz, "h2d_bytes_per_step": 24 * N_SEGMENTS, "form": "vrad_test_lines: xyz start / stop SoA"}},
            "roofline": {"bound": "hbm", "achieved": k1_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": k1_gbs / hbm_peak,
                         "traffic": 413.7e6, "traffic_source": "profiles/r02c_k1_s1_ncu.txt (dram read+write per launch)",
                         "kernel": "k1_test_lines (unordered form: one launch per step)", "peak_source": peak_src,
                         "note": "K1 is issue/divergence-bound by design (SURVEY 8d: HBM fraction ~1%); the binding target is >=1e9 rays/s"},
            "parity_checked": rparity,
            "cpu_baseline": cpu_rays_obj,
        },
        "large_scene": large,
        "widened": widened,
        "bsp_side": bsp_side,
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
    rank, local_rank, world = dist_setup()
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if world != args.gpus and rank == 0 and world == 1 and args.gpus > 1:
        print(json.dumps({"error": f"--gpus {args.gpus} needs torchrun with {args.gpus} ranks (WORLD_SIZE={world})"}))
        sys.exit(2)
    run_graft(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
