#!/usr/bin/env python
"""Benchmark of the grid particle-filter SLAM step (BASELINE.json metric: particle-beam updates/s,
plus the fraction of the measured HBM roofline reached by the dominant kernel).

    python bench.py --gpus N --steps K --warmup W              # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...    # the reference's CPU path (oracle port)

A "step" is one full GridMapSlam::update (slamrs/slam/src/grid/slam.rs:46-75) over one synthetic
simulator scan: motion sample + beam likelihood + ray update for every particle, weight
normalisation, systematic resampling, grid copies. Workload at N GPUs = N shards of configs[2]
(8,192 particles x 360 beams on a 1024^2 grid per GPU; 8 GPUs = configs[3], 65,536 particles),
i.e. weak scaling. Under torchrun each rank drives one GPU; rank 0 prints ONE JSON line.

Legs of the CUDA arm
  value  K steps on device-resident scans, no host synchronisation inside the timed region,
         CUDA events on the library's stream, max over ranks.
  e2e    K steps through the reference-facing call sequence of GridMapSlamNode::update
         (node.rs:47-60): update(host scan) -> estimated_pose() -> estimated_likelihood() into a
         pinned host grid; host<->device copies inside the timed region.
  roofline  the grid-copy kernel (k_copy): algorithmic bytes = bytes_per_grid * (grids
         written + source grids read (both counted on the device per step), divided by the
         kernel's own CUDA-event time, against MEASURED_PEAKS.json's HBM copy bandwidth.
  cpu_baseline  (N=1, rank 0) the CPU oracle on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "particle_beam_updates_per_s"
UNIT = "particle-beam updates/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=16)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["cuda", "reference"], default="cuda")
    ap.add_argument("--workload", default="c3", help="c1 | c2 | c3 (per-GPU shard; default configs[2])")
    ap.add_argument("--particles", type=int, default=0, help="override particles per GPU")
    ap.add_argument("--cpu-particles", type=int, default=0, help="particles of the bounded CPU sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    return ap.parse_args()


# ------------------------------------------------------------------------------ helpers
def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled every 200 ms while the timed region runs."""
    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
              "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu_index = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.gpu_index}", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                 "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, smax, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            p = [x.strip() for x in ln.split(",")]
            if len(p) < 9:
                continue
            try:
                sm.append(float(p[1])); smax.append(float(p[2])); power.append(float(p[3]))
            except ValueError:
                continue
            for name, v in zip(names, p[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        # "under load": samples in the upper half of the observed power range
        pw = np.array(power); smv = np.array(sm)
        load = smv[pw >= (pw.min() + pw.max()) / 2.0] if pw.max() > pw.min() else smv
        return {"sm_mhz": float(np.median(load)), "sm_max_mhz": float(max(smax)), "reasons": sorted(reasons),
                "samples": len(sm), "power_w_max": float(pw.max())}


def dist_env():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


# ------------------------------------------------------------------------------ reference arm
def run_reference(args, wl, rank, world):
    """The reference's own CPU implementation of the path: the C oracle (a port -- the Rust
    reference cannot be compiled in this image), all host threads, bounded sample per step."""
    if rank != 0:
        return  # rank 0 alone runs the CPU arm
    from oracle import oracle as O
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    cores = os.cpu_count() or 1
    n_cpu = args.cpu_particles or max(cores * 8, 64)
    res = time_oracle(O, wl, n_cpu, args.steps, args.warmup, cores)
    line = {
        "impl": "reference", "metric": METRIC, "value": res["value"], "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": res["ms_per_step"], "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32 geometry, f64 cells/weights", "data": "synthetic",
        "config": workload_config(wl, args.gpus, wl.n_particles * args.gpus),
        "cpu_baseline": {"value": res["value"], "unit": UNIT, "cores": cores, "kind": "port", "sample": res["sample"]},
        "e2e": {"value": res["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def time_oracle(O, wl, n_cpu, steps, warmup, threads, dead_likelihood=False):
    from slamrs_b200.slam import GridMapSlamConfig  # noqa: F401  (host types only)
    cfg = wl.slam_config(n_cpu)
    sim = wl.simulator()
    scans = [sim.next_scan(wl.speed_left, wl.speed_right) for _ in range(warmup + steps)]
    osl = O.OracleSlam(cfg.position, cfg.width, cfg.height, cfg.resolution, n_cpu, False)
    osl.set_threads(threads)
    osl.set_dead_likelihood(dead_likelihood)
    seed = 0x5EED5A11
    times = []
    for s, (obs, odo) in enumerate(scans):
        z = O.motion_normals(seed, s, 0, n_cpu)
        u = O.resample_uniform(seed, s)
        ang = obs.angle.astype(np.float32).astype(np.float64)
        dist = obs.distance.astype(np.float32).astype(np.float64)
        t0 = time.perf_counter()
        osl.update(ang, dist, obs.valid.astype(np.uint8), np.float32(odo.distance_left),
                   np.float32(odo.distance_right), np.float32(odo.wheel_distance), z, u)
        osl.estimated_pose(); osl.estimated_likelihood()   # what the node publishes per scan (node.rs:51-57)
        dt = time.perf_counter() - t0
        if s >= warmup:
            times.append(dt)
    osl.close()
    total = float(sum(times))
    return {"value": n_cpu * wl.n_beams * len(times) / total, "ms_per_step": 1e3 * total / len(times),
            "sample": f"{n_cpu} particles x {wl.n_beams} beams, {wl.grid}^2 f64 grid, {len(times)} steps after "
                      f"{warmup} warm-up, {threads} OpenMP threads over particles"
                      f"{', with the dead Map::likelihood() transform' if dead_likelihood else ''}"}


def workload_config(wl, n_gpus, n_total):
    return {"workload": f"{n_total} particles x {wl.n_beams} beams, {wl.grid}x{wl.grid} grid @ {wl.resolution} m "
                        f"({wl.name} shard per GPU; simulator scene x{wl.scene_scale:g}, range {wl.scanner_range:g} m)",
            "particles_per_gpu": wl.n_particles, "beams": wl.n_beams, "grid": wl.grid,
            "parallelism": f"particles sharded dp{n_gpus}",
            "l2": "working set (per-particle grids) is far larger than L2; no flush needed"}


# ------------------------------------------------------------------------------ CUDA arm
def run_cuda(args, wl, rank, world, local):
    import torch
    import torch.distributed as dist
    from slamrs_b200 import GpuPlacement, GridMapSlam, nccl_unique_id
    from slamrs_b200 import _lib

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the CUDA arm has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    nccl_id = None
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        buf = torch.zeros(_lib.NCCL_ID_BYTES, dtype=torch.uint8, device=dev)
        if rank == 0:
            buf.copy_(torch.frombuffer(bytearray(nccl_unique_id()), dtype=torch.uint8))
        dist.broadcast(buf, 0)
        nccl_id = bytes(buf.cpu().numpy().tobytes())

    n_total = wl.n_particles * world
    K, W = args.steps, args.warmup
    cfg = wl.slam_config(n_total)
    sim = wl.simulator()
    n_scans = W + K + (0 if args.no_e2e else K)
    scans = [sim.next_scan(wl.speed_left, wl.speed_right) for _ in range(n_scans)]

    slam = GridMapSlam(cfg, GpuPlacement(device=local, rank=rank, world_size=world, nccl_id=nccl_id))
    stream = torch.cuda.ExternalStream(slam.stream_ptr, device=dev)
    grid_bytes = slam.stats()["bytes_per_grid"]

    # device-resident scans for the `value` leg
    d_scans = []
    for obs, odo in scans[:W + K]:
        a = torch.from_numpy(obs.angle.astype(np.float32)).to(dev)
        d = torch.from_numpy(obs.distance.astype(np.float32)).to(dev)
        v = torch.from_numpy(obs.valid.astype(np.uint8)).to(dev)
        d_scans.append((a, d, v, float(obs.distance.max()) if len(obs) else 0.0, odo))
    torch.cuda.synchronize()

    def device_step(i):
        a, d, v, maxd, odo = d_scans[i]
        slam.set_scan_device(a.data_ptr(), d.data_ptr(), v.data_ptr(), a.numel(), maxd)
        slam.step_async(odo)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(W):
        device_step(i)
    slam.sync()

    clocks = ClockSampler(local)
    clocks.start()
    slam.set_profiling(True)
    launches0 = slam.launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record(stream)
    for i in range(W, W + K):
        device_step(i)
    e1.record(stream)
    slam.sync()
    barrier()
    ms_value = e0.elapsed_time(e1)
    launches = slam.launch_count - launches0
    phase_ms, psteps = slam.phase_ms()
    slam.set_profiling(False)
    hist = slam.step_history(W, K)
    copies = hist[:, 0].astype(np.float64)
    pulls = hist[:, 1].astype(np.float64)
    src_reads = hist[:, 3].astype(np.float64)
    st = slam.stats()

    # ---- e2e: the node's call sequence with host buffers
    e2e = None
    if not args.no_e2e:
        pinned = torch.empty(cfg_cells(slam), dtype=torch.float64).pin_memory()
        out_np = pinned.numpy()
        barrier()
        e0.record(stream)
        for obs, odo in scans[W + K:]:
            slam.update(obs, odo)
            slam.estimated_pose()
            slam.estimated_likelihood(out_np)
        e1.record(stream)
        slam.sync()
        barrier()
        ms_e2e = e0.elapsed_time(e1)
        e2e = {"ms": ms_e2e, "h2d": int(wl.n_beams * (4 + 4 + 1)), "d2h": int(12 + 8 * slam.grid_w * slam.grid_h)}
    clock_info = clocks.stop()

    # ---- max over ranks
    t = torch.tensor([ms_value, e2e["ms"] if e2e else 0.0, phase_ms["copy"], phase_ms["ray_update"],
                      phase_ms["motion_likelihood"], phase_ms["resample"], phase_ms["pull"], phase_ms["all_gather"]],
                     dtype=torch.float64, device=dev)
    agg = torch.tensor([copies.sum(), pulls.sum()], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(agg, op=dist.ReduceOp.SUM)
    t = t.cpu().numpy(); agg = agg.cpu().numpy()
    ms_value, ms_e2e = float(t[0]), float(t[1])

    slam.close()
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peak, peak_src = measured_peaks()
    pbu = n_total * wl.n_beams
    value = pbu * K / (ms_value * 1e-3)
    # roofline of the dominant kernel (grid copy) on rank 0's own launches
    copy_ms = phase_ms["copy"]
    # bytes the kernel really moves: every copied grid is written once; a source grid is read once
    # per fan-out sub-run (<= 16 destinations), not once per copy
    copy_bytes = float(grid_bytes) * (copies.sum() + src_reads.sum())
    achieved = copy_bytes / (copy_ms * 1e-3) / 1e9 if copy_ms > 0 else 0.0
    # whole-step algorithmic bytes: copies + ray RMW (8 B per cell-step, C_p ~ measured per scan) + gathers
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": ms_value / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32 geometry, f64 weights, u16x2 hit-counter cells", "data": "synthetic",
        "config": workload_config(wl, world, n_total),
        "clocks": clock_info,
        "gpu_launches": int(launches),
        "roofline": {
            "bound": "hbm", "kernel": "k_copy (resampling grid copies)", "achieved": achieved, "peak": peak,
            "unit": "GB/s", "frac": achieved / peak if peak else None, "traffic": None, "peak_source": peak_src,
            "bytes_per_launch": copy_bytes / K, "ms_per_launch": copy_ms / K,
            "grids_copied_per_step": float(copies.mean()), "source_reads_per_step": float(src_reads.mean()),
            "bytes_per_grid": int(grid_bytes),
            "algorithmic_bytes": "bytes_per_grid * (grids written + source grids read), per launch",
        },
        "phases_ms_per_step": {k: v / K for k, v in phase_ms.items()},
        "resample": {"grids_copied_per_step_all_gpus": float(agg[0] / K), "grids_pulled_per_step_all_gpus": float(agg[1] / K),
                     "distinct_sources_last_step_rank0": st["distinct_sources"], "spilled_cells_last_step": st["spilled_cells"],
                     "window_cells": st["window_cells"], "counter_saturated": st["counter_saturated"]},
    }
    if e2e:
        line["e2e"] = {"value": pbu * K / (ms_e2e * 1e-3), "unit": UNIT, "ms_per_step": ms_e2e / K,
                       "h2d_bytes_per_step": e2e["h2d"], "d2h_bytes_per_step": e2e["d2h"],
                       "path": "update(host scan) + estimated_pose() + estimated_likelihood() per step (node.rs:47-60)"}
    if world == 1 and not args.no_cpu_baseline:
        from oracle import oracle as O
        cores = os.cpu_count() or 1
        n_cpu = args.cpu_particles or max(cores * 8, 64)
        res = time_oracle(O, wl, n_cpu, steps=6, warmup=2, threads=cores)
        res1 = time_oracle(O, wl, max(16, n_cpu // 8), steps=4, warmup=1, threads=1)
        line["cpu_baseline"] = {"value": res["value"], "unit": UNIT, "cores": cores, "kind": "port",
                                "sample": res["sample"], "single_thread_value": res1["value"],
                                "single_thread_sample": res1["sample"],
                                "note": "C restatement of the Rust reference (no Rust toolchain in the image); the "
                                        "reference itself is single-threaded"}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def cfg_cells(slam):
    return slam.grid_w * slam.grid_h


def main():
    args = parse_args()
    rank, world, local = dist_env()
    if world == 1 and args.gpus > 1 and args.impl == "cuda":
        raise SystemExit("bench.py: --gpus N>1 must be launched with torch.distributed.run (one rank per GPU)")
    from slamrs_b200.workloads import WORKLOADS
    wl = WORKLOADS[args.workload]
    if args.particles:
        import dataclasses
        wl = dataclasses.replace(wl, n_particles=args.particles)
    if args.impl == "reference":
        run_reference(args, wl, rank, world)
    else:
        run_cuda(args, wl, rank, world, local)


if __name__ == "__main__":
    main()
