"""Pins the plain-C restatement (oracle/hs_oracle.c) to the reference itself (oracle/_ref = the unmodified
hector_slam_lib headers compiled here with the mini-Eigen shim): function-level known-answer cases and whole
streams, all bit for bit. Skipped where the reference build is absent (it needs /root/reference)."""
import numpy as np
import pytest

from hs_testutil import bits, make_streams

KINDS = ["port", "ref"]


@pytest.fixture()
def pair(oracle, have_ref):
    if not have_ref:
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    made = []

    def make(size=64, levels=2, res=0.05, start=(0.5, 0.5)):
        ps = [oracle.CpuProcessor(k, res, size, size, start, levels) for k in KINDS]
        made.extend(ps)
        return ps
    yield make
    for p in made:
        p.close()


def same_cells(a, b, level):
    ca, cb = a.download_level(level), b.download_level(level)
    return np.array_equal(ca["updateIndex"], cb["updateIndex"]) and np.array_equal(bits(ca["logOddsVal"]), bits(cb["logOddsVal"]))


def test_float_overload_trap(oracle, have_ref):
    """SURVEY hard part 4: abs(angleDiff) must be the float overload (0.5 rad > 0.13 rad -> True)."""
    for kind in KINDS:
        if kind == "ref" and not have_ref:
            continue
        assert oracle.pose_difference_larger_than(kind, (0, 0, .5), (0, 0, 0), 1e9, 0.13) is True
        assert oracle.pose_difference_larger_than(kind, (0, 0, .1), (0, 0, 0), 1e9, 0.13) is False


def test_pose_difference_and_normalize_edges(oracle, have_ref):
    if not have_ref:
        pytest.skip("needs oracle/_ref")
    rng = np.random.default_rng(0)
    cases = [((0, 0, 0), (3e38, 3e38, 3e38), 0.4, 0.9), ((1, 1, 3.1), (1, 1, -3.1), 0.4, 0.13), ((0.4, 0, 0), (0, 0, 0), 0.4, 0.9),
             ((0.40000004, 0, 0), (0, 0, 0), 0.4, 0.9), ((0, 0, 0.13), (0, 0, 0), 1.0, 0.13), ((0, 0, 6.5), (0, 0, 0), 1.0, 0.13)]
    for _ in range(200):
        cases.append((rng.uniform(-5, 5, 3), rng.uniform(-5, 5, 3), float(rng.uniform(0, 2)), float(rng.uniform(0, 2))))
    for p1, p2, d, a in cases:
        assert oracle.pose_difference_larger_than("port", p1, p2, d, a) == oracle.pose_difference_larger_than("ref", p1, p2, d, a)
    for a in list(rng.uniform(-50, 50, 500)) + [0.0, np.pi, -np.pi, 3.1415927, -3.1415927, 2 * np.pi, 1e-30, 100.0]:
        x, y = np.float32(oracle.normalize_angle("port", a)), np.float32(oracle.normalize_angle("ref", a))
        assert bits(x) == bits(y)


def test_transforms_roundtrip(pair):
    port, ref = pair(size=1024, levels=3)
    rng = np.random.default_rng(1)
    for level in range(3):
        for _ in range(50):
            w = rng.uniform(-20, 20, 3).astype(np.float32)
            assert np.array_equal(bits(port.world_to_map(level, w)), bits(ref.world_to_map(level, w)))
            m = rng.uniform(0, 1024 >> level, 3).astype(np.float32)
            assert np.array_equal(bits(port.map_to_world(level, m)), bits(ref.map_to_world(level, m)))
    assert port.scale_to_map() == ref.scale_to_map() == 20.0
    for level in range(3):
        assert port.level_dims(level) == ref.level_dims(level)


def test_interp_hand_built_grid(pair):
    """interpMapValueWithDerivatives on a small hand-built grid, incl. the map-limit rule 0 <= c <= size-2."""
    port, ref = pair(size=8, levels=1)
    rng = np.random.default_rng(2)
    vals = rng.uniform(-4, 4, 64).astype(np.float32)
    for p in (port, ref):
        for i, v in enumerate(vals):
            p.set_cell(0, i, v)
    coords = [(0.0, 0.0), (6.0, 6.0), (6.0000005, 3.0), (-1e-7, 2.0), (2.5, 2.5), (5.999, 0.001), (3.0, 6.0), (7.0, 7.0), (-0.0, 1.5)]
    coords += [tuple(c) for c in rng.uniform(-0.5, 6.5, (300, 2))]
    for x, y in coords:
        a, b = port.interp_with_derivs(0, x, y), ref.interp_with_derivs(0, x, y)
        assert np.array_equal(bits(a), bits(b)), (x, y, a, b)
    # out-of-bounds returns zeros
    assert not port.interp_with_derivs(0, 6.01, 1.0).any() and not port.interp_with_derivs(0, -0.01, 1.0).any()


def beams_from(origin_cell, ends):
    """points (robot frame, level-0 cells) that land exactly on the given end cells for pose (0,0,0) with start (0.5,0.5)."""
    return np.array(ends, np.float32)


def test_bresenham_octants_and_degenerate(pair):
    """updateByScan on a 64^2 grid: all octants, horizontal/vertical/diagonal, 1-cell beams, begin==end,
    beams leaving the map (dropped entirely), occupied-wins and undo-free sequences, bit-exact cells."""
    size = 64
    rng = np.random.default_rng(3)
    ends = [(10, 0), (-10, 0), (0, 10), (0, -10), (7, 7), (-7, 7), (7, -7), (-7, -7), (10, 3), (3, 10), (-10, 3), (-3, 10),
            (10, -3), (3, -10), (-10, -3), (-3, -10), (1, 0), (0, 1), (0, 0), (0.3, -0.2), (40, 5), (-40, 5), (5, 40), (5, -40),
            (31, 31), (-32, -32), (-32.4, 0), (31.4, 0), (31.6, 0)]
    ends += [tuple(e) for e in rng.uniform(-30, 30, (120, 2))]
    # occupied-wins / undo-free: beam A passes through (5,0) to end at (10,0) before/after beam B ends at (5,0)
    seqs = [ends, [(10, 0), (5, 0)], [(5, 0), (10, 0)], [(5, 0), (10, 0), (5, 0), (10.2, 0.1)]]
    for pts in seqs:
        port, ref = pair(size=size, levels=1)
        pts = np.array(pts, np.float32)
        for pose in [(0, 0, 0), (0.013, -0.02, 0.3), (0.5, 0.25, -2.0)]:
            for p in (port, ref):
                p.update_by_scan_level(0, pts, pose)
            assert same_cells(port, ref, 0)
        assert port.update_index(0) == ref.update_index(0) == 2


def test_occupied_clamp_at_50(pair):
    port, ref = pair(size=32, levels=1)
    pts = np.array([(6, 0), (0, 6), (4, 4)], np.float32)
    for p in (port, ref):
        p.set_update_factor_occupied(0.9)
    for k in range(40):   # 40 * 2.197 > 50: the v < 50 guard must kick in identically
        for p in (port, ref):
            p.update_by_scan_level(0, pts, (0, 0, 0))
    assert same_cells(port, ref, 0)
    v = port.download_level(0)["logOddsVal"]
    assert 50.0 <= v.max() < 50.0 + 2.2


def test_match_level_and_hessian_on_random_maps(pair):
    port, ref = pair(size=128, levels=2)
    rng = np.random.default_rng(4)
    for level in range(2):
        n = (128 >> level) ** 2
        cells = np.zeros(n, dtype=[("logOddsVal", np.float32), ("updateIndex", np.int32)])
        cells["logOddsVal"] = rng.normal(0, 2, n).astype(np.float32)
        cells["updateIndex"] = -1
        for p in (port, ref):
            p.upload_level(level, cells)
        pts = rng.uniform(-30 >> level, 30 >> level, (97, 2)).astype(np.float32)
        poses = np.concatenate([rng.uniform(20 >> level, 100 >> level, (40, 2)), rng.uniform(-4, 4, (40, 1))], axis=1).astype(np.float32)
        Ha, da = port.hessian_derivs_batch(level, poses, pts)
        Hb, db = ref.hessian_derivs_batch(level, poses, pts)
        assert np.array_equal(bits(Ha), bits(Hb)) and np.array_equal(bits(da), bits(db))
        for hint in [(0.1, -0.2, 0.05), (0.0, 0.0, 3.0), (-0.3, 0.4, -1.0)]:
            pa, ca = port.match_level(level, hint, pts, 5)
            pb, cb = ref.match_level(level, hint, pts, 5)
            assert np.array_equal(bits(pa), bits(pb)) and np.array_equal(bits(ca), bits(cb))
        # empty container: the hint is returned untouched
        pa, _ = port.match_level(level, (1, 2, 3), np.zeros((0, 2), np.float32), 5)
        assert np.array_equal(pa, np.array([1, 2, 3], np.float32))


def test_streams_bit_exact_incl_reset_and_flags(oracle, have_ref, hs):
    """Whole update() streams: ROS-node parameters, library defaults, gate 0/0, map_without_matching with stale
    coarse containers, empty scans, and reset() in the middle (update counters keep counting)."""
    if not have_ref:
        pytest.skip("needs oracle/_ref")
    n_scans = 45
    sc, ranges, pts, cnt, truth = make_streams(hs, 3, n_scans, seed0=321)
    ranges = ranges.copy()
    rng = np.random.default_rng(6)
    ranges[rng.random(ranges.shape) < 0.1] = np.inf
    ranges[20, 1, :] = np.inf            # an empty scan for stream 1
    pts, cnt = hs.scan_to_points_host(ranges, sc, 20.0)
    settings = [dict(free=0.4, occ=0.9, dist=0.4, ang=0.9), dict(free=0.4, occ=0.6, dist=0.4, ang=0.13), dict(free=0.3, occ=0.8, dist=0.0, ang=0.0)]
    for s, st in enumerate(settings):
        ps = [oracle.CpuProcessor(k) for k in KINDS]
        for p in ps:
            p.set_update_factor_free(st["free"]); p.set_update_factor_occupied(st["occ"])
            p.set_map_update_min_dist_diff(st["dist"]); p.set_map_update_min_angle_diff(st["ang"])
        poses = [np.zeros(3, np.float32) for _ in ps]
        for k in range(n_scans):
            mwm = (k % 9 == 8)
            if k == 30:
                for p in ps:
                    p.reset()
                poses = [np.zeros(3, np.float32) for _ in ps]
            for i, p in enumerate(ps):
                poses[i] = p.update(pts[k, s], cnt[k, s], hint=poses[i], map_without_matching=mwm)
            assert np.array_equal(bits(poses[0]), bits(poses[1])), (s, k)
            if cnt[k, s] and not mwm and k not in (0, 30):
                assert np.array_equal(bits(ps[0].cov()), bits(ps[1].cov()))
            assert ps[0].update_index(0) == ps[1].update_index(0)
            assert np.array_equal(bits(ps[0].last_map_update_pose()), bits(ps[1].last_map_update_pose()))
        for l in range(3):
            assert same_cells(ps[0], ps[1], l)
        for p in ps:
            p.close()
