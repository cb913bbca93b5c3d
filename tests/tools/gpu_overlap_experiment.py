"""Experiment (GPU): does splitting the batch over E engines on E CUDA streams overlap the latency-bound matcher of
one part with the issue-bound ray-caster of another?  usage: gpu_overlap_experiment.py <total_streams> <steps>"""
import os, sys, time
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import tfe_slam_ucl_b200 as hs
import workload

S = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
K = int(sys.argv[2]) if len(sys.argv) > 2 else 30
W = 3
sc = workload.scene_preset(0)
ranges, _ = workload.scangen_batch(sc, 0xC0FFEE, S, 0, K + W, truth=False)
pts, cnt = hs.scan_to_points_host(ranges, sc, 20.0)
for E in (1, 2, 4):
    per = S // E
    engs, streams, dp, dc = [], [], [], []
    for e in range(E):
        eng = hs.Engine(n_streams=per, max_points=360)
        eng.set_update_factor_free(0.4); eng.set_update_factor_occupied(0.9)
        eng.set_map_update_min_dist_diff(0.0); eng.set_map_update_min_angle_diff(0.0)
        st = torch.cuda.Stream(); eng.set_cuda_stream(st.cuda_stream)
        engs.append(eng); streams.append(st)
        dp.append(torch.from_numpy(np.ascontiguousarray(pts[:, e * per:(e + 1) * per])).cuda())
        dc.append(torch.from_numpy(np.ascontiguousarray(cnt[:, e * per:(e + 1) * per])).cuda())
    ps, cs = dp[0][0].numel() * 4, dc[0][0].numel() * 4
    def step(k):
        for e in range(E):
            engs[e].update_batch_device(per, dp[e].data_ptr() + k * ps, dc[e].data_ptr() + k * cs)
    for k in range(W): step(k)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for k in range(W, W + K): step(k)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print("engines %d x %d streams: %.1f us/step, %.3f M scans/s" % (E, per, 1e6 * dt / K, S * K / dt / 1e6))
    for eng in engs: eng.close()
