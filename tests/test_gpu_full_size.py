"""BASELINE.json's full-size configurations on the GPU, checked through sampled streams (bit for bit against the CPU
oracle) and size-independent properties:
  configs[1]  1024 streams x 360 beams, 3 levels 1024^2
  configs[2]  256 streams x 1440 beams (30 m), 4 levels 4096^2 @0.025 m
  configs[4]  8192 streams on one GPU (84 GiB of grids)
Properties: streams are independent (a stream's result does not depend on what else is in the batch or on its
slot), untouched streams stay pristine, integrations/cell-visit counters add up, update indices advance by one per
integrated scan."""
import numpy as np
import workload
import pytest

from hs_testutil import bits, make_streams

pytestmark = pytest.mark.gpu


def oracle_stream(oracle, pts, cnt, s, n_scans, size, levels, res, gate):
    p = oracle.CpuProcessor("port", res, size, size, (0.5, 0.5), levels)
    p.set_update_factor_free(0.4); p.set_update_factor_occupied(0.9)
    p.set_map_update_min_dist_diff(gate[0]); p.set_map_update_min_angle_diff(gate[1])
    poses = []
    for k in range(n_scans):
        poses.append(p.update(pts[k, s], cnt[k, s], hint=p.pose()).copy())
    return p, np.stack(poses)


def run_config(hs, oracle, n_streams, n_scans, preset, size, levels, res, max_points, sample, gate=(0.0, 0.0)):
    sc, ranges, pts, cnt, truth = make_streams(hs, n_streams, n_scans, preset=preset, seed0=0xC0FFEE, scale_to_map=1.0 / res)
    with hs.Engine(res, size, size, (0.5, 0.5), levels, n_streams=n_streams, max_points=max_points) as eng:
        eng.set_update_factor_free(0.4); eng.set_update_factor_occupied(0.9)
        eng.set_map_update_min_dist_diff(gate[0]); eng.set_map_update_min_angle_diff(gate[1])
        out = np.zeros((n_scans, n_streams, 3), np.float32)
        for k in range(n_scans):
            out[k], _ = eng.update_batch(pts[k], cnt[k], want_cov=False)
        st = eng.stats()
        assert st["updates"] == n_streams * n_scans
        if gate == (0.0, 0.0):
            assert st["integrations"] >= n_streams * (n_scans - 1)   # scan k>0 may be gated only if the pose did not move at all
        visits = 0
        for s in sample:
            p, ref = oracle_stream(oracle, pts, cnt, s, n_scans, size, levels, res, gate)
            assert np.array_equal(bits(out[:, s]), bits(ref)), "stream %d poses" % s
            for l in range(levels):
                g, c = eng.download_level(s, l), p.download_level(l)
                assert np.array_equal(g["updateIndex"], c["updateIndex"]) and np.array_equal(bits(g["logOddsVal"]), bits(c["logOddsVal"])), (s, l)
            assert eng.update_index(s, 0) == p.update_index(0)
            visits += p.cell_visits()
            p.close()
        # independence: re-run two sampled streams alone, in swapped slots of a tiny engine
        a, b = sample[0], sample[-1]
    with hs.Engine(res, size, size, (0.5, 0.5), levels, n_streams=2, max_points=max_points) as small:
        small.set_update_factor_free(0.4); small.set_update_factor_occupied(0.9)
        small.set_map_update_min_dist_diff(gate[0]); small.set_map_update_min_angle_diff(gate[1])
        for k in range(n_scans):
            o2, _ = small.update_batch(pts[k][[b, a]], cnt[k][[b, a]], want_cov=False)
            assert np.array_equal(bits(o2[0]), bits(out[k, b])) and np.array_equal(bits(o2[1]), bits(out[k, a]))
    return st, visits


def test_config1_1024_streams(hs, oracle):
    rng = np.random.default_rng(0)
    sample = sorted(rng.choice(1024, 12, replace=False).tolist())
    st, _ = run_config(hs, oracle, 1024, 6, workload.SCENE_RPLIDAR_ROOM, 1024, 3, 0.05, 360, sample)
    assert st["cell_visits"] > 1024 * 5 * 30000


def test_config2_hires_256_streams(hs, oracle):
    st, _ = run_config(hs, oracle, 256, 3, workload.SCENE_HIRES_HALL, 4096, 4, 0.025, 1440, [0, 131, 255])
    assert st["kernel_launches"] > 0


def test_config4_8192_streams_one_gpu(hs, oracle):
    run_config(hs, oracle, 8192, 3, workload.SCENE_RPLIDAR_ROOM, 1024, 3, 0.05, 360, [0, 4097, 8191], gate=(0.4, 0.9))
