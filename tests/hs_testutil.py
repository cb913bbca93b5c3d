"""Helpers shared by the test modules."""
import numpy as np
import workload


def make_streams(hs, n_streams, n_scans, preset=0, seed0=1000, scale_to_map=20.0, first_scan=0, max_points=None):
    """Synthetic scan streams -> (scene, ranges[k][s][b], points[k][s][mp][2], counts[k][s], truth)."""
    sc = workload.scene_preset(preset)
    ranges, truth = workload.scangen_batch(sc, seed0, n_streams, first_scan, n_scans)
    pts, cnt = hs.scan_to_points_host(ranges, sc, scale_to_map, max_points)
    return sc, ranges, pts, cnt, truth


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.int32)
