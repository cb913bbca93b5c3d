"""Experiment (GPU): where the end-to-end overhead of hs_update_batch comes from. Per step, 1024 streams:
 A device pointers, no sync between steps | B device pointers + hs_synchronize per step | C = B + hs_get_poses per step
 D hs_update_batch with pinned host buffers (copies + sync inside) | E = C with the kernels reading the pinned host
 buffers directly (zero-copy over PCIe)."""
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
d_pts = torch.from_numpy(pts).cuda(); d_cnt = torch.from_numpy(cnt).cuda()
h_pts = torch.from_numpy(pts).pin_memory(); h_cnt = torch.from_numpy(cnt).pin_memory()
h_pose = torch.empty((S, 3), dtype=torch.float32).pin_memory()
ps, cs = d_pts[0].numel() * 4, d_cnt[0].numel() * 4
out = np.empty((S, 3), np.float32)

def run(mode):
    with hs.Engine(n_streams=S, max_points=360) as eng:
        eng.set_update_factor_free(0.4); eng.set_update_factor_occupied(0.9)
        eng.set_map_update_min_dist_diff(0.0); eng.set_map_update_min_angle_diff(0.0)
        def step(k):
            if mode == "D":
                eng.update_batch_raw(S, h_pts.data_ptr() + k * ps, h_cnt.data_ptr() + k * cs, h_pose.data_ptr())
                return
            if mode == "E":   # zero-copy: the kernels read the pinned host buffers over PCIe (UVA), no staging copy
                eng.update_batch_device(S, h_pts.data_ptr() + k * ps, h_cnt.data_ptr() + k * cs)
            else:
                eng.update_batch_device(S, d_pts.data_ptr() + k * ps, d_cnt.data_ptr() + k * cs)
            if mode in ("B", "C", "E"):
                eng.synchronize()
            if mode in ("C", "E"):
                hs.lib().hs_get_poses(eng._h, 0, S, h_pose.data_ptr())
        for k in range(W): step(k)
        eng.synchronize()
        t0 = time.perf_counter()
        for k in range(W, W + K): step(k)
        eng.synchronize()
        return (time.perf_counter() - t0) / K * 1e6

for mode in ("A", "C", "D", "E", "D", "E"):
    print("mode %s: %.1f us per step" % (mode, run(mode)))
