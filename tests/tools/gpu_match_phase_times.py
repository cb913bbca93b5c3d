"""Experiment (GPU): per-phase clock64 accumulators of the matcher CTAs (HSGPU_DEBUG_TIMES; thread 0's view).
usage: gpu_match_phase_times.py [streams] [scans]   (set HSGPU_MATCH_VARIANT=1|2 selects the matcher; build with make EXPERIMENTS=1)"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
out = os.path.join(ROOT, "gpurun_out", "match_times.bin")
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
for k in range(K - 1):
    eng.update_batch(pts[k], cnt[k], want_cov=False)
eng.match_batch(pts[K - 1], cnt[K - 1])     # the matcher is the last kernel that writes the stamp buffer
eng.close()
t = np.fromfile(out, dtype=np.uint64).reshape(-1, 8)[:S, :5].astype(np.float64) / 1.965e3   # us at 1965 MHz
names = ["step1(addr/miss)", "step2(prob)", "step3/points", "sums", "solve+barrier"]
# thread 0 sits in the serial warp of the CTAs with blockIdx % 4 == 0 (there "sums" and "solve" are the serial work itself); in
# the other CTAs it is a worker, skips the sums and waits for the serial warp in "solve+barrier"
for label, sel in (("thread 0 in the serial warp", np.arange(S) % 4 == 0), ("thread 0 in a worker warp", np.arange(S) % 4 != 0)):
    print("matcher variant %s, %s: mean per-CTA us over the scan (14 evaluations): " % (os.environ.get("HSGPU_MATCH_VARIANT", "default"), label)
          + "  ".join("%s %.1f" % (n, t[sel, i].mean()) for i, n in enumerate(names)) + "  total %.1f" % t[sel].sum(1).mean())
