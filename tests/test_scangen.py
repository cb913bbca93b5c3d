"""Host-side input path: the synthetic scan generator is deterministic, and the product's LaserScan -> DataContainer
conversion (hs_scan_to_points_host, the host twin of k_scan_to_points) equals the oracle's restatement of
HectorMappingRos::rosLaserScanToDataContainer (HectorMappingRos.cpp:483-507), including the range window."""
import numpy as np
import workload

from hs_testutil import bits


def test_scangen_is_deterministic_and_random_access(hs):
    sc = workload.scene_preset(workload.SCENE_RPLIDAR_ROOM)
    a, ta = workload.scangen_batch(sc, 77, 3, 0, 6, threads=1)
    b, tb = workload.scangen_batch(sc, 77, 3, 0, 6, threads=4)
    assert np.array_equal(bits(a), bits(b)) and np.array_equal(bits(ta), bits(tb))
    c, _ = workload.scangen_batch(sc, 77, 3, 4, 2)          # scans 4,5 generated on their own
    assert np.array_equal(bits(a[4:6]), bits(c))
    d, _ = workload.scangen_batch(sc, 78, 2, 0, 6)           # stream seed0+1 of the first batch == stream 0 here
    assert np.array_equal(bits(a[:, 1:3]), bits(d))
    assert a.shape == (6, 3, 360) and np.isfinite(a).mean() > 0.95
    assert 0.2 < a[np.isfinite(a)].min() and a[np.isfinite(a)].max() < 12.9
    # smooth trajectory: <= 0.05 m and <= 0.03 rad per scan
    step = np.abs(np.diff(ta, axis=0))
    assert step[..., :2].max() < 0.05 and step[..., 2].max() < 0.03
    hall = workload.scene_preset(workload.SCENE_HIRES_HALL)
    r, _ = workload.scangen_batch(hall, 5, 1, 0, 1)
    assert r.shape == (1, 1, 1440)


def test_scan_to_points_matches_oracle_restatement(hs, oracle):
    sc = workload.scene_preset(workload.SCENE_RPLIDAR_ROOM)
    ranges, _ = workload.scangen_batch(sc, 3, 2, 0, 4)
    ranges = ranges.reshape(-1, 360).copy()
    rng = np.random.default_rng(0)
    ranges[0, ::7] = np.inf                 # dropped: >= range_max - 0.1
    ranges[1, ::5] = 0.1                    # dropped: <= range_min
    ranges[2, :] = 11.95                    # all dropped (max window is strict: 12 - 0.1)
    ranges[3, :] = sc.range_min             # all dropped (min window is strict)
    ranges[4, 10] = np.nan
    ranges[5] = rng.uniform(0.0, 13.0, 360)
    pts, cnt = hs.scan_to_points_host(ranges, sc, 20.0)
    for i in range(ranges.shape[0]):
        ref = oracle.scan_to_points(ranges[i], 360, sc.angle_min, sc.angle_increment, sc.range_min, sc.range_max, 20.0)
        assert cnt[i] == ref.shape[0]
        assert np.array_equal(bits(pts[i, :cnt[i]]), bits(ref))
    assert cnt[2] == 0 and cnt[3] == 0 and cnt[0] < 360


def _laser_transform(rng):
    """A laser->base tf::Transform: yaw + small roll/pitch, offset origin; row-major basis then origin (double)."""
    yaw, pitch, roll = rng.uniform(-3, 3), rng.uniform(-0.05, 0.05), rng.uniform(-0.05, 0.05)
    cy, sy, cp, sp, cr, sr = np.cos(yaw), np.sin(yaw), np.cos(pitch), np.sin(pitch), np.cos(roll), np.sin(roll)
    R = np.array([[cy * cp, cy * sp * sr - sy * cr, cy * sp * cr + sy * sr],
                  [sy * cp, sy * sp * sr + cy * cr, sy * sp * cr - cy * sr],
                  [-sp, cp * sr, cp * cr]])
    return np.concatenate([R.reshape(-1), rng.uniform(-0.3, 0.3, 3)])


def test_cloud_to_points_host_matches_oracle(hs, oracle):
    """rosPointCloudToDataContainer (HectorMappingRos.cpp:509-542): the product's host helper against the oracle's C
    restatement, bit for bit -- distance window, rear-point rejection, z window, origo, empty cloud."""
    rng = np.random.default_rng(12)
    for trial in range(20):
        n = int(rng.integers(0, 500))
        cloud = np.concatenate([rng.uniform(-6, 6, (n, 2)), rng.uniform(-1.5, 1.5, (n, 1))], 1).astype(np.float32)
        if n > 10:
            cloud[:5] = [(-0.5, 0.2, 0), (0.5, 0.2, 0), (0.39, 0.0, 0), (0.0, 0.41, 0), (-0.7, 0.1, 0.0)]
        t = _laser_transform(rng)
        args = (np.float32(20.0), np.float32(0.4 * 0.4), np.float32(30.0 * 30.0) if trial % 2 else np.float32(25.0), np.float32(-1.0), np.float32(1.0))
        a, oa = hs.cloud_to_points_host(cloud, t, *args)
        b, ob = oracle.cloud_to_points(cloud, t, *args)
        assert a.shape == b.shape and np.array_equal(a.view(np.int32), b.view(np.int32)) and np.array_equal(oa.view(np.int32), ob.view(np.int32))
        if n > 10:
            assert len(a) < n   # some points were filtered
