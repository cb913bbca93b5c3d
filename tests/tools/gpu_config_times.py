"""Measurement (GPU): BASELINE.json configs[2] (hi-res 1440-beam 30 m scans, 4 levels of 4096^2 @0.025 m, 256 streams) and
configs[3] (4096 pose hypotheses of one scan through getCompleteHessianDerivs). Prints one JSON line per config.
usage: gpu_config_times.py [steps]"""
import json, os, sys, time
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import tfe_slam_ucl_b200 as hs
import workload

K = int(sys.argv[1]) if len(sys.argv) > 1 else 20
W = 3

def hires():
    S, res, size, levels, mp = 256, 0.025, 4096, 4, 1440
    sc = workload.scene_preset(workload.SCENE_HIRES_HALL)
    ranges, _ = workload.scangen_batch(sc, 0xC0FFEE, S, 0, K + W, truth=False)
    pts, cnt = hs.scan_to_points_host(ranges, sc, 1.0 / res)
    with hs.Engine(res, size, size, (0.5, 0.5), levels, n_streams=S, max_points=mp) as eng:
        eng.set_update_factor_free(0.4); eng.set_update_factor_occupied(0.9)
        eng.set_map_update_min_dist_diff(0.0); eng.set_map_update_min_angle_diff(0.0)
        st = torch.cuda.Stream(); eng.set_cuda_stream(st.cuda_stream)
        dp = torch.from_numpy(pts).cuda(); dc = torch.from_numpy(cnt).cuda()
        ps, cs = dp[0].numel() * 4, dc[0].numel() * 4
        for k in range(W):
            eng.update_batch_device(S, dp.data_ptr() + k * ps, dc.data_ptr() + k * cs)
        torch.cuda.synchronize()
        s0 = eng.stats(); eng.profile_enable(True)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        for k in range(W, W + K):
            eng.update_batch_device(S, dp.data_ptr() + k * ps, dc.data_ptr() + k * cs)
        e1.record(st); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1); kt = eng.profile_collect(); s1 = eng.stats()
        visits = (s1["cell_visits"] - s0["cell_visits"]) / (S * K)
        print(json.dumps({"config": "configs[2] hi-res: 256 streams x 1440 beams (30 m), 4 levels 4096^2 @0.025 m, every scan integrated",
                          "scans_per_s": S * K / (ms * 1e-3), "ms_per_step": ms / K,
                          "k_match_ms": kt["k_match"]["ms"] / K, "k_raycast_ms": kt["k_raycast"]["ms"] / K,
                          "cell_visits_per_scan": visits, "mean_points": float(cnt.mean())}))

def hypotheses():
    sc = workload.scene_preset(workload.SCENE_RPLIDAR_ROOM)
    ranges, truth = workload.scangen_batch(sc, 0xC0FFEE, 1, 0, 201, truth=True)
    pts, cnt = hs.scan_to_points_host(ranges, sc, 20.0)
    with hs.Engine(n_streams=1, max_points=360) as eng:
        eng.set_update_factor_free(0.4); eng.set_update_factor_occupied(0.9)
        eng.set_map_update_min_dist_diff(0.0); eng.set_map_update_min_angle_diff(0.0)
        for k in range(200):
            pose, _ = eng.update_batch(pts[k], cnt[k], want_cov=False)
        rng = np.random.default_rng(1)
        n = 4096
        # map coordinates of level 0: scale * world + scale * offset (offset = 0.05*1024*0.5)
        base = np.array([20.0 * pose[0, 0] + 512.0, 20.0 * pose[0, 1] + 512.0, pose[0, 2]], np.float32)
        poses = base[None, :] + rng.uniform(-1, 1, (n, 3)).astype(np.float32) * np.array([20.0, 20.0, 0.5], np.float32)
        p = pts[200, 0, :cnt[200, 0]]
        for strict in (1, 0):
            eng.set_strict_summation(strict)
            for _ in range(3):
                eng.hessian_derivs_batch(0, 0, poses, p)
            t0 = time.perf_counter()
            R = 20
            for _ in range(R):
                H, d = eng.hessian_derivs_batch(0, 0, poses, p)
            dt = (time.perf_counter() - t0) / R
            print(json.dumps({"config": "configs[3]: 4096 pose hypotheses x %d points, level 0, host call incl. H2D/D2H" % len(p), "strict": strict,
                              "ms_per_call": dt * 1e3, "hypotheses_per_s": n / dt, "algorithmic_GBps": n * len(p) * 16 / dt / 1e9}))
        # full coarse-to-fine multi-start match (14 evaluations per hypothesis), state untouched
        eng.set_strict_summation(1)
        hints = pose[0][None, :] + rng.uniform(-1, 1, (n, 3)).astype(np.float32) * np.array([1.0, 1.0, 0.5], np.float32)
        for _ in range(3):
            eng.match_hypotheses(0, p, hints)
        t0 = time.perf_counter()
        for _ in range(R):
            got, cov = eng.match_hypotheses(0, p, hints)
        dt = (time.perf_counter() - t0) / R
        print(json.dumps({"config": "configs[3] full match: 4096 pose hypotheses x %d points, 3 levels (14 evaluations each), host call incl. H2D/D2H" % len(p),
                          "ms_per_call": dt * 1e3, "hypotheses_per_s": n / dt, "evaluations_per_s": 14 * n / dt}))

hypotheses()
hires()
