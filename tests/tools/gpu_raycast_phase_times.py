"""[needs a library built with `make -C tfe-slam-ucl_b200/csrc EXPERIMENTS=1`] Experiment (GPU): per-CTA phase time stamps of k_raycast (HSGPU_DEBUG_TIMES), printed as a phase/occupancy summary.
usage: gpu_raycast_phase_times.py [streams] [scans]"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
out = os.path.join(ROOT, "gpurun_out", "raycast_times.bin")
os.environ["HSGPU_DEBUG_TIMES"] = out
import torch
import tfe_slam_ucl_b200 as hs
import workload

S = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
K = int(sys.argv[2]) if len(sys.argv) > 2 else 30
sc = workload.scene_preset(0)
ranges, _ = workload.scangen_batch(sc, 0xC0FFEE, S, 0, K, truth=False)
pts, cnt = hs.scan_to_points_host(ranges, sc, 20.0)
eng = hs.Engine(n_streams=S, max_points=360)
eng.set_update_factor_free(0.4); eng.set_update_factor_occupied(0.9)
eng.set_map_update_min_dist_diff(0.0); eng.set_map_update_min_angle_diff(0.0)
dp = torch.from_numpy(pts).cuda(); dc = torch.from_numpy(cnt).cuda()
ps, cs = dp[0].numel() * 4, dc[0].numel() * 4
for k in range(K):
    eng.update_batch_device(S, dp.data_ptr() + k * ps, dc.data_ptr() + k * cs)
torch.cuda.synchronize()
eng.close()
t = np.fromfile(out, dtype=np.uint64).reshape(-1, 8)[: S * 3, :6].astype(np.int64)
t0 = t[:, 0].min()
t = (t - t0) / 1e3   # us
names = ["setup+A1", "A2+A3", "B1", "B2", "C"]
print("kernel span %.1f us, CTAs %d" % (t[:, 5].max(), len(t)))
for lvl in range(3):
    tl = t[lvl::3]
    d = np.diff(tl, axis=1)
    print("level %d: CTA life mean %.1f us | " % (lvl, (tl[:, 5] - tl[:, 0]).mean()) + "  ".join("%s %.2f" % (n, d[:, i].mean()) for i, n in enumerate(names)))
# how many CTAs are inside each phase over time (sampled every 2 us)
ts = np.arange(0, t[:, 5].max(), 4.0)
print("time_us  " + "  ".join("%8s" % n for n in names))
for x in ts:
    row = [int(((t[:, i] <= x) & (x < t[:, i + 1])).sum()) for i in range(5)]
    print("%7.1f  " % x + "  ".join("%8d" % r for r in row))
