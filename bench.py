#!/usr/bin/env python
"""bench.py -- scans matched+integrated per second of the hsgpu engine (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one HectorSlamProcessor::update (scan match + gate + map integration) for every stream of the
batch: `streams` scans per GPU per step. Workload at every N: BASELINE.json configs[1] per GPU -- 1024
independent 360-beam streams, 3-level 1024^2 maps @0.05 m -- i.e. streams are sharded over ranks with no
data-path collective (weak scaling). The map-update gate is 0 m / 0 rad so that EVERY scan is integrated
("matched+integrated"; the reference's height_mapping.launch uses the same 0/0 gate), update factors are the
ROS node's 0.4 / 0.9.

    value  : inputs (DataContainer points of all steps) already resident in HBM, hs_update_batch_device,
             timed with CUDA events on the launching stream.
    e2e    : hs_update_batch (the reference-facing C-ABI call) with pinned HOST buffers; the H2D copy of
             each step's points/counts and the D2H read of the matched poses are inside the timed region. Each call
             returns when its poses are in host memory; its map integration finishes on the engine's stream (the next call
             is ordered behind it) and the timed region ends with a device synchronisation, so all K integrations are inside.
    roofline    : the dominant kernel's algorithmic bytes / its CUDA-event time, against MEASURED_PEAKS.json.
    cpu_baseline: the CPU reference (oracle/_ref if built, else the C restatement) on a bounded sample.

--impl reference times the reference's CPU implementation (oracle/) on the host cores for the same workload.
"""
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

import workload  # noqa: E402  synthetic scan streams (workload/libhsscangen.so): shared by both arms, loads no CUDA library

METRIC = "scans matched+integrated/sec (360-beam, 3-level 1024^2 grid)"
UNIT = "scans/s"
PROF_EVERY = 8   # per-kernel CUDA events around every 8th timed step (event pairs keep consecutive kernels from overlapping)
NODE_FACTORS = (0.4, 0.9)  # HectorMappingRos.cpp:72-73
E_EVALS = 14               # 4*(L-1)+6 evaluations per scan, L=3
FALLBACK_HBM_GBS = 6650.0  # B200_PROFILING.md fallback


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--streams", type=int, default=1024, help="streams per GPU (BASELINE configs[1])")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="target CPU work for the cpu_baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks/throttle reasons sampled DURING the timed region."""

    def __init__(self, gpu_index):
        super().__init__(daemon=True)
        self.gpu = gpu_index
        self.samples = []
        self.stop_flag = threading.Event()
        self.proc = None

    def run(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + q, "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                if self.stop_flag.is_set():
                    break
                self.samples.append(line.strip())
        except Exception:
            pass

    def finish(self):
        self.stop_flag.set()
        if self.proc:
            try:
                self.proc.terminate()
            except Exception:
                pass
        sm, mx, reasons = [], [], set()
        for s in self.samples:
            parts = [p.strip() for p in s.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0])); mx.append(float(parts[1]))
            except ValueError:
                continue
            for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], parts[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons), "samples": len(sm)}


def gen_inputs(n_streams, n_steps, seed0):
    sc = workload.scene_preset(workload.SCENE_RPLIDAR_ROOM)
    ranges, _ = workload.scangen_batch(sc, seed0, n_streams, 0, n_steps, truth=False)
    pts, cnt = workload.scan_to_points_host(ranges, sc, 20.0)  # scaleToMap = 1/0.05
    return sc, pts, cnt


def cpu_run(pts, cnt, n_streams, n_scans, threads):
    """The reference's CPU path on the host cores: one processor per stream, streams over threads."""
    from oracle import pyoracle
    kind = "ref" if pyoracle.have("ref") else "port"
    procs = [pyoracle.CpuProcessor(kind, 0.05, 1024, 1024, (0.5, 0.5), 3) for _ in range(n_streams)]
    for p in procs:
        p.set_update_factor_free(NODE_FACTORS[0]); p.set_update_factor_occupied(NODE_FACTORS[1])
        p.set_map_update_min_dist_diff(0.0); p.set_map_update_min_angle_diff(0.0)
    secs, _, _ = pyoracle.run_streams(procs, pts[:n_scans, :n_streams], cnt[:n_scans, :n_streams], threads=threads,
                                      stream_major=True, want_poses=False)
    for p in procs:
        p.close()
    return kind, secs


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    cores = os.cpu_count() or 1
    K, W, S = args.steps, max(args.warmup, 0), args.streams
    config = {"workload": "BASELINE configs[1] per GPU: %d independent 360-beam streams, 3-level 1024^2 maps @0.05 m, gate 0/0 "
                          "(every scan integrated), factors 0.4/0.9" % S,
              "streams_per_gpu": S, "beams": 360, "levels": 3, "map": "1024^2@0.05m",
              "cache": "working set %.1f GiB of grids per GPU >> 126 MB L2 (inputs larger than L2, no flush needed)" % (S * 1376256 * 8 / 2**30),
              "sharding": "streams sharded over ranks, no data-path collective"}

    # ------------------------------------------------------------------ reference arm (CPU)
    if args.impl == "reference":
        if rank != 0:
            return
        from oracle import pyoracle
        if not pyoracle.have("port") and not pyoracle.have("ref"):
            pyoracle.build(ref=False)
        sample_streams = max(cores, min(S, 16 * cores))   # a bounded sample of the S streams, every host thread busy
        # one step = `spp` consecutive scans of each sampled stream: with few steps the timed region would otherwise be a few
        # milliseconds of thread start-up and first-touch page faults, which understates the CPU
        spp = max(1, min(64, 1024 // max(W + K, 1)))
        steps = (W + K) * spp
        _, pts, cnt = gen_inputs(sample_streams, steps, 0xC0FFEE)
        kind = "ref" if pyoracle.have("ref") else "port"
        procs = [pyoracle.CpuProcessor(kind, 0.05, 1024, 1024, (0.5, 0.5), 3) for _ in range(sample_streams)]
        for p in procs:
            p.set_update_factor_free(NODE_FACTORS[0]); p.set_update_factor_occupied(NODE_FACTORS[1])
            p.set_map_update_min_dist_diff(0.0); p.set_map_update_min_angle_diff(0.0)
        if W:   # warm-up steps are part of the same stream histories, run untimed first
            pyoracle.run_streams(procs, pts[:W * spp], cnt[:W * spp], threads=cores, stream_major=True, want_poses=False)
        # stream-major order (a thread runs all timed scans of one stream before the next stream): the CPU's best case, its
        # 21 MiB of grids + caches per stream stay in cache
        secs, _, _ = pyoracle.run_streams(procs, pts[W * spp:], cnt[W * spp:], threads=cores, stream_major=True, want_poses=False)
        value = sample_streams * K * spp / secs
        sample = "%d of the %d streams x %d scans per step, one processor per stream, stream-major order, %d host threads (%s)" % (
            sample_streams, S, spp, cores, "unmodified reference headers, oracle/_ref" if kind == "ref" else "C restatement, oracle/hs_oracle.c")
        print(json.dumps({"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": K, "warmup": W,
                          "ms_per_step": 1e3 * secs / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                          "dtype": "f32", "data": "synthetic", "config": config,
                          "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "reference" if kind == "ref" else "port", "sample": sample},
                          "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                          "gpu_launches": 0}))
        return

    # ------------------------------------------------------------------ our arm (GPU)
    import torch
    import torch.distributed as dist
    import tfe_slam_ucl_b200 as hs
    from tfe_slam_ucl_b200 import sharding

    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl ours needs a CUDA device: hsgpu has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        # stdout carries exactly one JSON line: NCCL prints its version banner to fd 1 when the first communicator is
        # created, so fd 1 points at stderr until that has happened
        sys.stdout.flush()
        saved_fd = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved_fd, 1)
            os.close(saved_fd)

    steps = W + K
    t_gen = time.time()
    first_stream, n_local = sharding.shard_range(world * S, world, rank)   # this rank's block of the global streams
    assert n_local == S
    sc, pts, cnt = gen_inputs(S, steps, 0xC0FFEE + first_stream)
    t_gen = time.time() - t_gen

    eng = hs.Engine(0.05, 1024, 1024, (0.5, 0.5), 3, n_streams=S, max_points=360, device=local_rank)
    eng.set_update_factor_free(NODE_FACTORS[0]); eng.set_update_factor_occupied(NODE_FACTORS[1])
    eng.set_map_update_min_dist_diff(0.0); eng.set_map_update_min_angle_diff(0.0)
    stream = torch.cuda.Stream()
    eng.set_cuda_stream(stream.cuda_stream)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- leg 1: device-resident inputs --------------------------------------------------------
    d_pts = torch.from_numpy(pts).to("cuda", non_blocking=False)   # [steps][S][360][2]
    d_cnt = torch.from_numpy(cnt).to("cuda", non_blocking=False)   # [steps][S]
    pts_step = d_pts[0].numel() * 4
    cnt_step = d_cnt[0].numel() * 4

    def run_device(k):
        eng.update_batch_device(S, d_pts.data_ptr() + k * pts_step, d_cnt.data_ptr() + k * cnt_step)

    for k in range(W):
        run_device(k)
    barrier()
    st0 = eng.stats()
    eng.profile_enable(PROF_EVERY)   # CUDA events around the kernels of every PROF_EVERY-th timed step
    sampler = ClockSampler(local_rank)
    sampler.start()
    time.sleep(0.3)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record(stream)
    for k in range(W, W + K):
        run_device(k)
    ev1.record(stream)
    barrier()
    ms_total = ev0.elapsed_time(ev1)
    ktimes = eng.profile_collect()
    eng.profile_enable(False)
    st1 = eng.stats()
    launches = st1["kernel_launches"] - st0["kernel_launches"]
    integrations = st1["integrations"] - st0["integrations"]
    visits = st1["cell_visits"] - st0["cell_visits"]

    agg = torch.tensor([float(integrations), float(visits), float(launches)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(agg, op=dist.ReduceOp.SUM)
    total_scans, max_s = sharding.aggregate_throughput(S * K, ms_total * 1e-3, device="cuda")   # sum of units, MAX of times
    ms_max = max_s * 1e3
    value = total_scans / max_s

    # ---- leg 2: end to end through the host C ABI --------------------------------------------
    eng.reset()
    h_pts = torch.from_numpy(pts).pin_memory()
    h_cnt = torch.from_numpy(cnt).pin_memory()
    h_pose = torch.empty((S, 3), dtype=torch.float32).pin_memory()

    def run_host(k):
        eng.update_batch_raw(S, h_pts.data_ptr() + k * pts_step, h_cnt.data_ptr() + k * cnt_step, h_pose.data_ptr())

    for k in range(W):
        run_host(k)
    barrier()
    t0 = time.perf_counter()
    for k in range(W, W + K):
        run_host(k)   # returns when this step's poses are valid in host memory
    torch.cuda.synchronize()   # ... and the last step's map integration has finished
    e2e_s = time.perf_counter() - t0
    pose_checksum = float(h_pose.sum())

    # ---- leg 2b (informational): the same, but the caller hands over LaserScan ranges and the engine does
    # rosLaserScanToDataContainer on the GPU (hs_update_batch_ranges): half the host-to-device bytes
    eng.reset()
    ranges_all, _ = workload.scangen_batch(sc, 0xC0FFEE + first_stream, S, 0, steps, truth=False)
    eng.set_scan_geometry(sc.n_beams, sc.angle_min, sc.angle_increment, sc.range_min, sc.range_max)
    h_rng = torch.from_numpy(ranges_all).pin_memory()
    rng_step = h_rng[0].numel() * 4
    for k in range(W):
        eng.update_batch_ranges_raw(S, h_rng.data_ptr() + k * rng_step, h_pose.data_ptr())
    barrier()
    t0 = time.perf_counter()
    for k in range(W, W + K):
        eng.update_batch_ranges_raw(S, h_rng.data_ptr() + k * rng_step, h_pose.data_ptr())
    torch.cuda.synchronize()
    e2e_rng_s = time.perf_counter() - t0
    e2e_rng_scans, e2e_rng_max_s = sharding.aggregate_throughput(S * K, e2e_rng_s, device="cuda")
    ranges_checksum = float(h_pose.sum())
    # ---- leg 2c (informational): pipelined submission, hs_submit_batch / hs_wait -- what a server with queued scans uses.
    # Every step's inputs still cross PCIe from pinned host memory and every step's poses are written back to the host
    # inside the timed region, but step k+1's transfer overlaps step k's kernels (not the headline: the reference's update()
    # is synchronous)
    eng.reset()
    h_pose_all = torch.empty((steps, S, 3), dtype=torch.float32).pin_memory()
    for k in range(W):
        eng.submit_batch_raw(S, h_pts.data_ptr() + k * pts_step, h_cnt.data_ptr() + k * cnt_step, h_pose_all[k].data_ptr())
    eng.wait()
    barrier()
    t0 = time.perf_counter()
    for k in range(W, W + K):
        eng.submit_batch_raw(S, h_pts.data_ptr() + k * pts_step, h_cnt.data_ptr() + k * cnt_step, h_pose_all[k].data_ptr())
    eng.wait()
    e2e_pipe_s = time.perf_counter() - t0
    e2e_pipe_scans, e2e_pipe_max_s = sharding.aggregate_throughput(S * K, e2e_pipe_s, device="cuda")
    pipe_checksum = float(h_pose_all[W + K - 1].sum())
    clocks = sampler.finish()      # sampled across all timed legs (device-resident and end-to-end)
    e2e_scans, e2e_max_s = sharding.aggregate_throughput(S * K, e2e_s, device="cuda")
    e2e_value = e2e_scans / e2e_max_s
    h2d_step = pts_step + cnt_step
    d2h_step = S * 3 * 4

    # ---- roofline of the dominant kernel (rank 0's engine) ----------------------------------
    peak, peak_src = measured_peak()
    n_valid = float(cnt[W:W + K].sum()) / K                       # valid points per step
    bytes_match = E_EVALS * n_valid * 16.0                        # E * N * 4 cells * 4 B, per launch
    bytes_ray = (st1["cell_visits"] - st0["cell_visits"]) / K * 16.0   # visits * 8 B * (read + write), per launch
    km, kr = ktimes["k_match"], ktimes["k_raycast"]
    per = {"k_match": (bytes_match, km["ms"] / max(km["launches"], 1)), "k_raycast": (bytes_ray, kr["ms"] / max(kr["launches"], 1))}
    dom = max(per, key=lambda n: per[n][1])
    ach = per[dom][0] / (per[dom][1] * 1e-3) / 1e9
    traffic, traffic_src = None, None
    try:   # dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed `ncu --set full` capture
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            tj = json.load(f)
        if tj.get("streams_per_launch") == S and dom in tj.get("kernels", {}):
            traffic, traffic_src = tj["kernels"][dom]["dram_bytes_per_launch"], tj.get("source")
    except Exception:
        pass
    roofline = {"bound": "hbm", "kernel": dom, "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": traffic,
                "traffic_source": traffic_src,
                "peak_source": peak_src, "algorithmic_bytes_per_launch": per[dom][0], "avg_launch_ms": per[dom][1],
                "kernel_timing": "CUDA events on the launching stream around the kernels of every %d-th timed step (%d launches of each kernel)" % (PROF_EVERY, kr["launches"]),
                "kernels": {n: {"avg_launch_ms": per[n][1], "algorithmic_bytes_per_launch": per[n][0],
                                "achieved_gbs": per[n][0] / (per[n][1] * 1e-3) / 1e9 if per[n][1] > 0 else None,
                                "share_of_step": per[n][1] / (ms_total / K)} for n in per},
                "whole_step_achieved_gbs": (bytes_match + bytes_ray + h2d_step) / (ms_total / K * 1e-3) / 1e9}

    # ---- CPU baseline on a bounded sample (rank 0, N=1 only) ---------------------------------
    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle import pyoracle
        if not pyoracle.have("port") and not pyoracle.have("ref"):
            pyoracle.build(ref=False)
        est_rate = 2500.0 * cores
        n_cs = int(max(cores, min(S, args.cpu_seconds * est_rate / steps)))
        kind, secs = cpu_run(pts, cnt, n_cs, steps, cores)
        cpu_baseline = {"value": n_cs * steps / secs, "unit": UNIT, "cores": cores, "kind": "reference" if kind == "ref" else "port",
                        "sample": "%d of the %d streams x %d scans (same inputs), one processor per stream, %d threads, stream-major order"
                                  % (n_cs, S, steps, cores),
                        "per_core": n_cs * steps / secs / cores}

    if rank == 0:
        out = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
               "ms_per_step": ms_max / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
               "dtype": "f32", "data": "synthetic", "config": config, "impl": "ours",
               "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d_step, "d2h_bytes_per_step": d2h_step,
                       "api": "hs_update_batch (host pointers, pinned)", "pose_checksum": pose_checksum,
                       "sync": "every call returns with its poses in host memory; its map integration finishes on the engine's stream, ordered before the next call's match; the timed region ends with a device synchronisation",
                       "ranges_api": {"value": e2e_rng_scans / e2e_rng_max_s, "h2d_bytes_per_step": rng_step, "d2h_bytes_per_step": d2h_step,
                                      "api": "hs_update_batch_ranges (LaserScan ranges in, conversion on the GPU)",
                                      "pose_checksum": ranges_checksum},
                       "pipelined_api": {"value": e2e_pipe_scans / e2e_pipe_max_s, "h2d_bytes_per_step": h2d_step, "d2h_bytes_per_step": d2h_step,
                                         "api": "hs_submit_batch x K + hs_wait (inputs double-buffered on a copy stream; not the headline)",
                                         "pose_checksum": pipe_checksum}},
               "gpu_launches": int(agg[2].item()),
               "integrated_scans": int(agg[0].item()), "scans": world * S * K,
               "roofline": roofline, "cpu_baseline": cpu_baseline, "clocks": clocks,
               "input_gen_s": t_gen}
        print(json.dumps(out))
    eng.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
