"""[needs a library built with `make -C tfe-slam-ucl_b200/csrc EXPERIMENTS=1`] Experiment (GPU): is k_raycast bound by HBM? Teacher-forced ray-casts (work independent of the map content) with the
grids of all CTAs aliased onto 8 streams' grids (HSGPU_DEBUG_SKIP=256: 84 MB, fits the 126 MB L2) versus the normal layout
(10.5 GiB). Prints the k_raycast device time per launch from the engine's CUDA events."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import tfe_slam_ucl_b200 as hs
import workload

S, K = 1024, 30
sc = workload.scene_preset(0)
ranges, _ = workload.scangen_batch(sc, 0xC0FFEE, S, 0, K, truth=False)
pts, cnt = hs.scan_to_points_host(ranges, sc, 20.0)
os.environ.pop("HSGPU_DEBUG_SKIP", None)
with hs.Engine(n_streams=S, max_points=360) as eng:
    eng.set_update_factor_free(0.4); eng.set_update_factor_occupied(0.9)
    eng.set_map_update_min_dist_diff(0.0); eng.set_map_update_min_angle_diff(0.0)
    for k in range(K):
        poses, _ = eng.update_batch(pts[k], cnt[k], want_cov=False)
for mode in ("0", "256"):
    os.environ["HSGPU_DEBUG_SKIP"] = mode
    with hs.Engine(n_streams=S, max_points=360) as eng:
        eng.set_update_factor_free(0.4); eng.set_update_factor_occupied(0.9)
        for r in range(5):
            eng.raycast_batch(pts[K - 1], cnt[K - 1], poses)
        eng.profile_enable(True)
        for r in range(20):
            eng.raycast_batch(pts[K - 1 - (r % 5)], cnt[K - 1 - (r % 5)], poses)
        t = eng.profile_collect()
        print("HSGPU_DEBUG_SKIP=%s: k_raycast %.1f us per launch (%d launches)" % (mode, 1e3 * t["k_raycast"]["ms"] / max(t["k_raycast"]["launches"], 1), t["k_raycast"]["launches"]))
