"""Diagnostic (run on a GPU box): closed-loop pose differences fast vs strict vs oracle, and kernel times."""
import sys, os, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import tfe_slam_ucl_b200 as hs
from oracle import pyoracle
from hs_testutil import make_streams

n_streams, n_scans = int(sys.argv[1]) if len(sys.argv) > 1 else 32, int(sys.argv[2]) if len(sys.argv) > 2 else 200
gate = (0.4, 0.9) if (len(sys.argv) > 3 and sys.argv[3] == "gate") else (0.0, 0.0)
sc, ranges, pts, cnt, truth = make_streams(hs, n_streams, n_scans, seed0=900)
procs = [pyoracle.CpuProcessor("port") for _ in range(n_streams)]
for p in procs:
    p.set_update_factor_free(0.4); p.set_update_factor_occupied(0.9); p.set_map_update_min_dist_diff(gate[0]); p.set_map_update_min_angle_diff(gate[1])
_, cpu_poses, _ = pyoracle.run_streams(procs, pts, cnt, threads=8)
res = {}
for strict in (True, False):
    with hs.Engine(n_streams=n_streams, max_points=360) as eng:
        eng.set_update_factor_free(0.4); eng.set_update_factor_occupied(0.9); eng.set_map_update_min_dist_diff(gate[0]); eng.set_map_update_min_angle_diff(gate[1])
        eng.set_strict_summation(strict)
        eng.profile_enable(True)
        out = np.zeros((n_scans, n_streams, 3), np.float32)
        for k in range(n_scans):
            out[k], _ = eng.update_batch(pts[k], cnt[k], want_cov=False)
        print("strict" if strict else "fast", eng.profile_collect(), eng.stats())
        res[strict] = out
        nd = 0
        for s in range(n_streams):
            for l in range(3):
                g, c = eng.download_level(s, l), procs[s].download_level(l)
                nd += int((g["updateIndex"] != c["updateIndex"]).sum())
        print("  cells with different marker vs oracle:", nd)
d = np.abs(res[False] - cpu_poses)
print("strict == oracle bitwise:", np.array_equal(res[True].view(np.int32), cpu_poses.view(np.int32)))
print("fast vs oracle: max per component", d.max(axis=(0, 1)))
per_scan = d.max(axis=(1, 2))
print("fast vs oracle per-scan max (first 12):", per_scan[:12])
print("quantiles of per-(scan,stream) max diff:", np.quantile(d.max(axis=2), [0.5, 0.9, 0.99, 0.999, 1.0]))
bad = np.argwhere(d.max(axis=2) > 1e-4)
print("n (scan,stream) pairs above 1e-4:", len(bad), "of", n_scans * n_streams, "first:", bad[:10].tolist())
print("truth error (oracle):", np.abs(cpu_poses - truth).max(axis=(0, 1)))
