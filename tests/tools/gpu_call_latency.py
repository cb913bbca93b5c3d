"""Experiment (GPU): latency of ONE isolated hs_update_batch call (device idle before it, 1024 streams x 360 points, pinned
buffers) -- time until the poses are in host memory -- and the back-to-back rate, per HSGPU_E2E_DEFER setting."""
import os, sys, time
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import tfe_slam_ucl_b200 as hs
import workload

S, K, W = 1024, 100, 5
sc = workload.scene_preset(0)
ranges, _ = workload.scangen_batch(sc, 0xC0FFEE, S, 0, K + W, truth=False)
pts, cnt = hs.scan_to_points_host(ranges, sc, 20.0)
h_pts = torch.from_numpy(pts).pin_memory(); h_cnt = torch.from_numpy(cnt).pin_memory()
h_pose = torch.empty((S, 3), dtype=torch.float32).pin_memory()
ps, cs = h_pts[0].numel() * 4, h_cnt[0].numel() * 4
with hs.Engine(n_streams=S, max_points=360) as eng:
    eng.set_update_factor_free(0.4); eng.set_update_factor_occupied(0.9)
    eng.set_map_update_min_dist_diff(0.0); eng.set_map_update_min_angle_diff(0.0)
    lat = []
    for k in range(W + K):
        eng.synchronize()
        t0 = time.perf_counter()
        eng.update_batch_raw(S, h_pts.data_ptr() + k * ps, h_cnt.data_ptr() + k * cs, h_pose.data_ptr())
        t1 = time.perf_counter()
        eng.synchronize()
        t2 = time.perf_counter()
        if k >= W: lat.append((t1 - t0, t2 - t0))
    lat = np.array(lat) * 1e6
    print("HSGPU_E2E_DEFER=%s isolated call: poses after %.1f us (median), all work done after %.1f us" %
          (os.environ.get("HSGPU_E2E_DEFER", "default"), np.median(lat[:, 0]), np.median(lat[:, 1])))
