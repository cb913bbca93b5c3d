#!/usr/bin/env python
"""bench.py -- scans matched+integrated per second of the hsgpu engine (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--no-configs]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one HectorSlamProcessor::update (scan match + gate + map integration) for every stream of the
batch: `streams` scans per GPU per step. Headline workload at every N: BASELINE.json configs[1] per GPU -- 1024
independent 360-beam streams, 3-level 1024^2 maps @0.05 m -- i.e. streams are sharded over ranks with no
data-path collective (weak scaling). The map-update gate is 0 m / 0 rad so that EVERY scan is integrated
("matched+integrated"; the reference's height_mapping.launch uses the same 0/0 gate), update factors are the
ROS node's 0.4 / 0.9.

    e2e    : THE HEADLINE (SURVEY 8(d) defines the metric with the transfers): hs_update_batch, the reference-facing C-ABI
             call, with pinned HOST buffers; the H2D copy of each step's points/counts and the D2H read of the matched poses
             are inside the timed region. Each call returns when its poses are in host memory; its map integration finishes
             on the engine's stream (the next call is ordered behind it) and the timed region ends with a device
             synchronisation, so all K integrations are inside. The K-step loop is repeated R times on continuing stream
             histories; e2e.value is the MEDIAN repeat, e2e.spread = (max - min) / median.
    value  : inputs (DataContainer points of all steps) already resident in HBM, hs_update_batch_device,
             timed with CUDA events on the launching stream.
    roofline    : the dominant kernel's algorithmic bytes / its CUDA-event time, against MEASURED_PEAKS.json.
    cpu_baseline: the CPU reference (oracle/_ref if built, else the C restatement) on a bounded sample.
    configs     : one sub-record per other BASELINE.json configuration (they are parity-test cases, not the headline):
                  configs[2] hi-res 1440-beam / 4 x 4096^2 / 256 streams, configs[3] 4096 pose hypotheses (single
                  getCompleteHessianDerivs evaluation and the full multi-start match), configs[4] 8192 streams (all on this
                  GPU at N=1, sharded 8192/N per GPU at N>1), each with ms_per_step, per-kernel CUDA-event times and a roofline.

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
FALLBACK_HBM_GBS = 6650.0  # B200_PROFILING.md fallback
E2E_REPEATS = 7
E2E_STEP_BUDGET = 448      # R = clamp(budget // K, 1, E2E_REPEATS): the pinned inputs of all repeats are distinct scans


def n_evals(levels):
    return 4 * (levels - 1) + 6   # 1 + maxIterations evaluations per level: 6 on level 0, 4 on the others


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--streams", type=int, default=1024, help="streams per GPU (BASELINE configs[1])")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="target CPU work for the cpu_baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="skip the sub-records of the other BASELINE configurations")
    ap.add_argument("--clock-interval-ms", type=int, default=20, help="nvidia-smi sampling period during the timed legs")
    ap.add_argument("--only-config", default=None, help="run one sub-record alone and print it (c3 | c4 | c5): for ncu captures")
    return ap.parse_args()


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks/throttle reasons sampled DURING the timed region."""

    def __init__(self, gpu_index, interval_ms=20):
        super().__init__(daemon=True)
        self.gpu = gpu_index
        self.interval_ms = max(int(interval_ms), 5)
        self.samples = []
        self.stop_flag = threading.Event()
        self.proc = None

    def run(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + q, "--format=csv,noheader,nounits", "-lms", str(self.interval_ms)],
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


def gen_inputs(n_streams, n_steps, seed0, preset=workload.SCENE_RPLIDAR_ROOM, scale_to_map=20.0, first_scan=0):
    sc = workload.scene_preset(preset)
    ranges, _ = workload.scangen_batch(sc, seed0, n_streams, first_scan, n_steps, truth=False)
    pts, cnt = workload.scan_to_points_host(ranges, sc, scale_to_map)  # scaleToMap = 1/resolution
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


def reference_arm(args, config, cores):
    """bench.py --impl reference: the reference's own CPU implementation of the path on the host cores (rank 0 only)."""
    from oracle import pyoracle
    K, W, S = args.steps, max(args.warmup, 0), args.streams
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
    sample = ("%d of the %d streams x %d scans per step, one processor per stream, stream-major order, %d host threads (%s); the timed scans "
              "are scans %d..%d of each sampled stream's history (our arm times scans %d..%d of every stream: the rates are comparable, "
              "the maturity of the maps is not)") % (
        sample_streams, S, spp, cores, "unmodified reference headers, oracle/_ref" if kind == "ref" else "C restatement, oracle/hs_oracle.c",
        W * spp, steps - 1, W, W + K - 1)
    print(json.dumps({"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": K, "warmup": W,
                      "ms_per_step": 1e3 * secs / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                      "dtype": "f32", "data": "synthetic", "config": config,
                      "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "reference" if kind == "ref" else "port", "sample": sample},
                      "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                      "gpu_launches": 0}))


# ---------------------------------------------------------------------------------------------------------------------
# sub-records of the other BASELINE.json configurations

def configure_node(eng):
    eng.set_update_factor_free(NODE_FACTORS[0]); eng.set_update_factor_occupied(NODE_FACTORS[1])
    eng.set_map_update_min_dist_diff(0.0); eng.set_map_update_min_angle_diff(0.0)


def kernel_roofline(name, algorithmic_bytes_per_launch, ms_per_launch, peak, bound="hbm"):
    ach = algorithmic_bytes_per_launch / (ms_per_launch * 1e-3) / 1e9 if ms_per_launch > 0 else None
    return {"kernel": name, "bound": bound, "algorithmic_bytes_per_launch": algorithmic_bytes_per_launch, "avg_launch_ms": ms_per_launch,
            "achieved": ach, "peak": peak, "unit": "GB/s", "frac": (ach / peak) if ach else None}


def run_stream_config(hs, torch, device, S, K, W, preset, res, size, levels, max_points, seed0, peak, label):
    """S streams x (W + K) scans through hs_update_batch_device with device-resident inputs; the step time from CUDA events
    over the K timed steps, the kernel times from event pairs around the kernels of a sample of those steps."""
    t0 = time.time()
    sc, pts, cnt = gen_inputs(S, W + K, seed0, preset, 1.0 / res)
    gen_s = time.time() - t0
    stream = torch.cuda.Stream()
    prof_every = max(2, K // 4)
    with hs.Engine(res, size, size, (0.5, 0.5), levels, n_streams=S, max_points=max_points, device=device) as eng:
        configure_node(eng)
        eng.set_cuda_stream(stream.cuda_stream)
        d_pts = torch.from_numpy(pts).to("cuda"); d_cnt = torch.from_numpy(cnt).to("cuda")
        ps, cs = d_pts[0].numel() * 4, d_cnt[0].numel() * 4
        for k in range(W):
            eng.update_batch_device(S, d_pts.data_ptr() + k * ps, d_cnt.data_ptr() + k * cs)
        torch.cuda.synchronize()
        s0 = eng.stats()
        eng.profile_enable(prof_every)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for k in range(W, W + K):
            eng.update_batch_device(S, d_pts.data_ptr() + k * ps, d_cnt.data_ptr() + k * cs)
        e1.record(stream)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        kt = eng.profile_collect()
        eng.profile_enable(False)
        s1 = eng.stats()
        del d_pts, d_cnt
    visits = (s1["cell_visits"] - s0["cell_visits"]) / K
    n_valid = float(cnt[W:W + K].sum()) / K
    km, kr = kt["k_match"], kt["k_raycast"]
    ms_m, ms_r = km["ms"] / max(km["launches"], 1), kr["ms"] / max(kr["launches"], 1)
    rec = {"workload": label, "streams": S, "steps": K, "warmup": W, "value": S * K / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms / K,
           "integrated_scans": int(s1["integrations"] - s0["integrations"]), "mean_points_per_scan": n_valid / S,
           "cell_visits_per_scan": visits / S, "grid_bytes": int(S * sum((size >> l) ** 2 for l in range(levels)) * 8),
           "timing": "CUDA events on the launching stream over the K steps; kernel times from event pairs around the kernels of every %d-th step" % prof_every,
           "kernels": {"k_match": kernel_roofline("k_match2" if max_points > 768 else "k_match", n_evals(levels) * n_valid * 16.0, ms_m, peak),
                       "k_raycast": kernel_roofline("k_raycast", visits * 16.0, ms_r, peak)},
           "input_gen_s": gen_s}
    dom = max(rec["kernels"], key=lambda n: rec["kernels"][n]["avg_launch_ms"])
    rec["roofline"] = dict(rec["kernels"][dom])
    rec["roofline"]["whole_step_achieved_gbs"] = (n_evals(levels) * n_valid * 16.0 + visits * 16.0) / (ms / K * 1e-3) / 1e9
    return rec


def measure_l2_gbs(torch):
    """L2-resident copy bandwidth (read + write bytes) of torch's copy kernel over two 24 MB buffers: the ceiling that applies
    to a kernel whose whole working set sits in the 126 MB L2 (configs[3]: one level's probability plane)."""
    n = 6 * 1024 * 1024
    a = torch.empty(n, dtype=torch.float32, device="cuda").normal_()
    b = torch.empty_like(a)
    for _ in range(5):
        b.copy_(a)
    torch.cuda.synchronize()
    best = 0.0
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            b.copy_(a)
        e1.record()
        torch.cuda.synchronize()
        best = max(best, 20 * 2 * n * 4 / (e0.elapsed_time(e1) * 1e-3) / 1e9)
    return best


def run_hypotheses_config(hs, torch, device, peak_hbm, n_hyp=4096, reps=20):
    """BASELINE configs[3]: 4096 pose hypotheses of one 360-beam scan against one stream's maps (built by 200 updates).
    (a) one getCompleteHessianDerivs evaluation per hypothesis, device-pointer entry point (CUDA events) and host entry point
    (host clock, copies inside); (b) the full coarse-to-fine multi-start match (14 evaluations per hypothesis)."""
    sc, pts, cnt = gen_inputs(1, 201, 0xC0FFEE)
    l2_peak = measure_l2_gbs(torch)
    stream = torch.cuda.Stream()
    with hs.Engine(n_streams=1, max_points=360, device=device) as eng:
        configure_node(eng)
        eng.set_cuda_stream(stream.cuda_stream)
        for k in range(200):
            pose, _ = eng.update_batch(pts[k], cnt[k], want_cov=False)
        rng = np.random.default_rng(1)
        base = np.array([20.0 * pose[0, 0] + 512.0, 20.0 * pose[0, 1] + 512.0, pose[0, 2]], np.float32)   # level-0 map coordinates
        poses = (base[None, :] + rng.uniform(-1, 1, (n_hyp, 3)).astype(np.float32) * np.array([20.0, 20.0, 0.5], np.float32)).astype(np.float32)
        p = np.ascontiguousarray(pts[200, 0, :cnt[200, 0]])
        n_pts = p.shape[0]
        alg_bytes = float(n_hyp) * n_pts * 16.0
        d_pose = torch.from_numpy(poses).cuda(); d_pts = torch.from_numpy(p).cuda()
        d_H = torch.zeros((n_hyp, 9), device="cuda"); d_d = torch.zeros((n_hyp, 3), device="cuda")
        out = {}
        for strict in (1, 0):
            eng.set_strict_summation(strict)
            for _ in range(3):
                eng.hessian_derivs_batch_device(0, 0, d_pose.data_ptr(), n_hyp, d_pts.data_ptr(), n_pts, d_H.data_ptr(), d_d.data_ptr())
            torch.cuda.synchronize()
            eng.profile_enable(1)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for _ in range(reps):
                eng.hessian_derivs_batch_device(0, 0, d_pose.data_ptr(), n_hyp, d_pts.data_ptr(), n_pts, d_H.data_ptr(), d_d.data_ptr())
            e1.record(stream)
            torch.cuda.synchronize()
            kt = eng.profile_collect(); eng.profile_enable(False)
            ms_call = e0.elapsed_time(e1) / reps
            ms_kernel = kt["k_hessian_derivs"]["ms"] / max(kt["k_hessian_derivs"]["launches"], 1)
            for _ in range(3):
                eng.hessian_derivs_batch(0, 0, poses, p)
            t0 = time.perf_counter()
            for _ in range(reps):
                eng.hessian_derivs_batch(0, 0, poses, p)
            ms_host = (time.perf_counter() - t0) / reps * 1e3
            out["strict" if strict else "tree"] = {
                "device_entry": {"api": "hs_hessian_derivs_batch_device", "ms_per_call": ms_call, "hypotheses_per_s": n_hyp / (ms_call * 1e-3),
                                 "kernel_ms": ms_kernel},
                "host_entry": {"api": "hs_hessian_derivs_batch (H2D poses+points, D2H H+dTr, sync inside)", "ms_per_call": ms_host,
                               "hypotheses_per_s": n_hyp / (ms_host * 1e-3)},
                "roofline_l2": kernel_roofline("k_hessian_derivs", alg_bytes, ms_kernel, l2_peak, bound="l2"),
                "roofline_hbm": kernel_roofline("k_hessian_derivs", alg_bytes, ms_kernel, peak_hbm)}
        # the probability block plane is built once per map state: time a cold build
        eng.set_strict_summation(1)
        eng.profile_enable(1)
        eng.reset(0)   # bumps the map epoch
        eng.hessian_derivs_batch_device(0, 0, d_pose.data_ptr(), n_hyp, d_pts.data_ptr(), n_pts, d_H.data_ptr(), d_d.data_ptr())
        kt = eng.profile_collect(); eng.profile_enable(False)
        plane_ms = kt["k_build_pblk"]["ms"] / max(kt["k_build_pblk"]["launches"], 1)
        # (b) the full multi-start match
        for k in range(200):   # rebuild the map the reset above cleared
            pose, _ = eng.update_batch(pts[k], cnt[k], want_cov=False)
        hints = (pose[0][None, :] + rng.uniform(-1, 1, (n_hyp, 3)).astype(np.float32) * np.array([1.0, 1.0, 0.5], np.float32)).astype(np.float32)
        for _ in range(3):
            eng.match_hypotheses(0, p, hints)
        eng.profile_enable(1)
        t0 = time.perf_counter()
        for _ in range(reps):
            eng.match_hypotheses(0, p, hints)
        ms_full = (time.perf_counter() - t0) / reps * 1e3
        kt = eng.profile_collect(); eng.profile_enable(False)
        ms_full_kernel = kt["k_match"]["ms"] / max(kt["k_match"]["launches"], 1)
        s = out["strict"]
        rec = {"workload": "BASELINE configs[3]: %d pose hypotheses (truth +-1 m, +-0.5 rad) x %d points of one scan against one stream's 3-level 1024^2 maps after 200 scans" % (n_hyp, n_pts),
               "hypotheses": n_hyp, "points": n_pts, "value": s["device_entry"]["hypotheses_per_s"], "unit": "hypotheses/s",
               "ms_per_step": s["device_entry"]["ms_per_call"],
               "single_evaluation": out,
               "probability_plane_build_ms": plane_ms,
               "full_match": {"api": "hs_match_hypotheses (host call incl. copies): 3 levels, %d evaluations per hypothesis" % n_evals(3),
                              "ms_per_call": ms_full, "hypotheses_per_s": n_hyp / (ms_full * 1e-3), "evaluations_per_s": n_evals(3) * n_hyp / (ms_full * 1e-3),
                              "kernel_ms": ms_full_kernel,
                              "roofline_l2": kernel_roofline("k_match (hypothesis mode)", n_evals(3) * alg_bytes, ms_full_kernel, l2_peak, bound="l2")},
               "l2_peak_gbs": l2_peak, "l2_peak_source": "measured in this run: torch copy kernel over two 24 MB buffers (L2-resident), read+write bytes, best of 5 x 20",
               "roofline": dict(s["roofline_l2"]),
               "note": "one level's probability plane (16 MB) and the scan stay in the 126 MB L2: the ceiling is L2 bandwidth, HBM fraction given for reference"}
    return rec


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

    if args.impl == "reference":
        if rank == 0:
            reference_arm(args, config, cores)
        return

    # ------------------------------------------------------------------ our arm (GPU)
    import torch
    import torch.distributed as dist
    import tfe_slam_ucl_b200 as hs
    from tfe_slam_ucl_b200 import sharding

    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl ours needs a CUDA device: hsgpu has no CPU fallback")
    torch.cuda.set_device(local_rank)
    peak, peak_src = measured_peak()

    if args.only_config:   # one sub-record alone (ncu captures of its kernels)
        if args.only_config == "c3":
            rec = run_stream_config(hs, torch, local_rank, 256, max(K, 1), W, workload.SCENE_HIRES_HALL, 0.025, 4096, 4, 1440, 0xC0FFEE, peak, "configs[2]")
        elif args.only_config == "c4":
            rec = run_hypotheses_config(hs, torch, local_rank, peak, reps=max(K, 1))
        else:
            rec = run_stream_config(hs, torch, local_rank, 8192, max(K, 1), W, workload.SCENE_RPLIDAR_ROOM, 0.05, 1024, 3, 360, 0xC0FFEE, peak, "configs[4] on one GPU")
        print(json.dumps(rec))
        return

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

    R = max(1, min(E2E_REPEATS, E2E_STEP_BUDGET // max(K, 1)))
    steps = W + R * K          # leg 1 uses the first W + K of them
    t_gen = time.time()
    first_stream, n_local = sharding.shard_range(world * S, world, rank)   # this rank's block of the global streams
    assert n_local == S
    sc, pts, cnt = gen_inputs(S, steps, 0xC0FFEE + first_stream)
    t_gen = time.time() - t_gen

    eng = hs.Engine(0.05, 1024, 1024, (0.5, 0.5), 3, n_streams=S, max_points=360, device=local_rank)
    configure_node(eng)
    stream = torch.cuda.Stream()
    eng.set_cuda_stream(stream.cuda_stream)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- leg 1: device-resident inputs --------------------------------------------------------
    d_pts = torch.from_numpy(pts[:W + K]).to("cuda", non_blocking=False)   # [W+K][S][360][2]
    d_cnt = torch.from_numpy(cnt[:W + K]).to("cuda", non_blocking=False)   # [W+K][S]
    pts_step = d_pts[0].numel() * 4
    cnt_step = d_cnt[0].numel() * 4

    def run_device(k):
        eng.update_batch_device(S, d_pts.data_ptr() + k * pts_step, d_cnt.data_ptr() + k * cnt_step)

    for k in range(W):
        run_device(k)
    barrier()
    st0 = eng.stats()
    eng.profile_enable(PROF_EVERY)   # CUDA events around the kernels of every PROF_EVERY-th timed step
    sampler = ClockSampler(local_rank, args.clock_interval_ms)
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
    del d_pts, d_cnt

    # ---- leg 2: end to end through the host C ABI (the headline) ---------------------------------
    eng.reset()
    h_pts = torch.from_numpy(pts).pin_memory()
    h_cnt = torch.from_numpy(cnt).pin_memory()
    h_pose = torch.empty((S, 3), dtype=torch.float32).pin_memory()

    def run_host(k):
        eng.update_batch_raw(S, h_pts.data_ptr() + k * pts_step, h_cnt.data_ptr() + k * cnt_step, h_pose.data_ptr())

    for k in range(W):
        run_host(k)
    e2e_times = []
    pose_checksum = 0.0
    for r in range(R):           # R repeats of the K-step loop on continuing stream histories (distinct scans)
        barrier()
        t0 = time.perf_counter()
        for k in range(W + r * K, W + (r + 1) * K):
            run_host(k)   # returns when this step's poses are valid in host memory
        torch.cuda.synchronize()   # ... and the last step's map integration has finished
        e2e_times.append(time.perf_counter() - t0)
        if r == 0:
            pose_checksum = float(h_pose.sum())
    e2e_rates = []
    for t in e2e_times:          # every repeat: units summed over ranks, time = max over ranks
        u, s_max = sharding.aggregate_throughput(S * K, t, device="cuda")
        e2e_rates.append(u / s_max)
    e2e_value = float(np.median(e2e_rates))
    e2e_spread = (max(e2e_rates) - min(e2e_rates)) / e2e_value

    # ---- leg 2b (informational): the same, but the caller hands over LaserScan ranges and the engine does
    # rosLaserScanToDataContainer on the GPU (hs_update_batch_ranges): half the host-to-device bytes
    eng.reset()
    ranges_all, _ = workload.scangen_batch(sc, 0xC0FFEE + first_stream, S, 0, W + K, truth=False)
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
    h_pose_all = torch.empty((W + K, S, 3), dtype=torch.float32).pin_memory()
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
    h2d_step = pts_step + cnt_step
    d2h_step = S * 3 * 4

    # ---- roofline of the dominant kernel (rank 0's engine) ----------------------------------
    n_valid = float(cnt[W:W + K].sum()) / K                       # valid points per step
    bytes_match = n_evals(3) * n_valid * 16.0                     # E * N * 4 cells * 4 B, per launch
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
                "dram_frac": (traffic / (per[dom][1] * 1e-3) / 1e9 / peak) if traffic else None,
                "peak_source": peak_src, "algorithmic_bytes_per_launch": per[dom][0], "avg_launch_ms": per[dom][1],
                "kernel_timing": "CUDA events on the launching stream around the kernels of every %d-th timed step (%d launches of each kernel)" % (PROF_EVERY, kr["launches"]),
                "kernels": {n: {"avg_launch_ms": per[n][1], "algorithmic_bytes_per_launch": per[n][0],
                                "achieved_gbs": per[n][0] / (per[n][1] * 1e-3) / 1e9 if per[n][1] > 0 else None,
                                "share_of_step": per[n][1] / (ms_total / K)} for n in per},
                "whole_step_achieved_gbs": (bytes_match + bytes_ray + h2d_step) / (ms_total / K * 1e-3) / 1e9}
    eng.close()
    del h_pts, h_cnt, h_rng, h_pose_all

    # ---- sub-records of the other BASELINE configurations ---------------------------------------------------------
    configs = None
    if not args.no_configs:
        configs = {}
        t_cfg = time.time()
        # configs[4]: 8192 streams in total -- all on this GPU at N=1 (84 GiB of grids), 8192/N per GPU at N>1 (strong scaling)
        n5_first, n5 = sharding.shard_range(8192, world, rank)
        K5, W5 = 10, 5
        c5 = run_stream_config(hs, torch, local_rank, n5, K5, W5, workload.SCENE_RPLIDAR_ROOM, 0.05, 1024, 3, 360, 0xC0FFEE + n5_first, peak,
                               "BASELINE configs[4]: 8192 independent 360-beam streams in total, %d per GPU on %d GPU(s), 3-level 1024^2 maps, gate 0/0" % (n5, world))
        u5, t5 = sharding.aggregate_throughput(n5 * K5, c5["ms_per_step"] * K5 * 1e-3, device="cuda")
        c5["value"] = u5 / t5
        c5["scaling"] = "strong (8192 streams over %d GPU(s)); time = max over ranks" % world
        c5["per_gpu_vs_headline"] = {"per_gpu_1024_resident_streams": value / world, "per_gpu_this": c5["value"] / world,
                                     "note": "per-GPU rate with %d resident streams against the headline's 1024 (TLB reach / L2 effects of the larger working set)" % n5}
        configs["c5_8192_streams"] = c5
        if rank == 0 and world == 1:
            configs["c3_hires"] = run_stream_config(
                hs, torch, local_rank, 256, 8, 3, workload.SCENE_HIRES_HALL, 0.025, 4096, 4, 1440, 0xC0FFEE, peak,
                "BASELINE configs[2]: 256 streams x 1440 beams (30 m), 4 levels 4096^2 @0.025 m, every scan integrated")
            configs["c4_hypotheses"] = run_hypotheses_config(hs, torch, local_rank, peak)
        if world > 1:
            from tfe_slam_ucl_b200 import hypothesis_sharding
            c4s = hypothesis_sharding.bench_record(hs, torch, dist, local_rank, rank, world)
            if rank == 0:
                configs["c4_hypotheses_sharded"] = c4s
        if rank == 0:
            configs["seconds"] = time.time() - t_cfg

    # ---- CPU baseline on a bounded sample (rank 0, N=1 only) ---------------------------------
    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle import pyoracle
        if not pyoracle.have("port") and not pyoracle.have("ref"):
            pyoracle.build(ref=False)
        est_rate = 2500.0 * cores
        n_scans_cpu = W + K
        n_cs = int(max(cores, min(S, args.cpu_seconds * est_rate / n_scans_cpu)))
        kind, secs = cpu_run(pts, cnt, n_cs, n_scans_cpu, cores)
        cpu_baseline = {"value": n_cs * n_scans_cpu / secs, "unit": UNIT, "cores": cores, "kind": "reference" if kind == "ref" else "port",
                        "sample": "%d of the %d streams x scans 0..%d of their histories (the GPU's inputs; the GPU times scans %d..%d), one processor per stream, %d threads, stream-major order"
                                  % (n_cs, S, n_scans_cpu - 1, W, W + K - 1, cores),
                        "per_core": n_cs * n_scans_cpu / secs / cores}

    if rank == 0:
        out = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
               "ms_per_step": ms_max / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
               "dtype": "f32", "data": "synthetic", "config": config, "impl": "ours",
               "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d_step, "d2h_bytes_per_step": d2h_step,
                       "api": "hs_update_batch (host pointers, pinned)", "pose_checksum": pose_checksum,
                       "repeats": R, "repeat_values": e2e_rates, "spread": e2e_spread,
                       "statistic": "median of %d repeats of the %d-step loop on continuing stream histories (distinct scans per repeat)" % (R, K),
                       "sync": "every call returns with its poses in host memory; its map integration finishes on the engine's stream, ordered before the next call's match; every repeat's timed region ends with a device synchronisation",
                       "ranges_api": {"value": e2e_rng_scans / e2e_rng_max_s, "h2d_bytes_per_step": rng_step, "d2h_bytes_per_step": d2h_step,
                                      "api": "hs_update_batch_ranges (LaserScan ranges in, conversion on the GPU)",
                                      "pose_checksum": ranges_checksum},
                       "pipelined_api": {"value": e2e_pipe_scans / e2e_pipe_max_s, "h2d_bytes_per_step": h2d_step, "d2h_bytes_per_step": d2h_step,
                                         "api": "hs_submit_batch x K + hs_wait (inputs double-buffered on a copy stream; not the headline)",
                                         "pose_checksum": pipe_checksum}},
               "gpu_launches": int(agg[2].item()),
               "integrated_scans": int(agg[0].item()), "scans": world * S * K,
               "roofline": roofline, "cpu_baseline": cpu_baseline, "clocks": clocks,
               "configs": configs, "input_gen_s": t_gen}
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
