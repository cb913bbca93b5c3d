"""GPU edge cases against the CPU oracle, bit for bit: empty and ragged scans, beams leaving the map, pose hints,
map_without_matching (stale coarse containers), reset, stream subsets via stream_ids, the LaserScan-ranges entry
point, band-wise ray-casting of large bounding boxes, map upload/download/classification, and the
size-independent properties used at BASELINE sizes."""
import os

import numpy as np
import pytest

from hs_testutil import bits, make_streams

pytestmark = pytest.mark.gpu


def setup_pair(hs, oracle, n, size=1024, levels=3, res=0.05, max_points=360, free=0.4, occ=0.9, dist=0.0, ang=0.0, start=(0.5, 0.5)):
    eng = hs.Engine(res, size, size, start, levels, n_streams=n, max_points=max_points)
    procs = [oracle.CpuProcessor("port", res, size, size, start, levels) for _ in range(n)]
    for p in procs + [eng]:
        p.set_update_factor_free(free); p.set_update_factor_occupied(occ)
        p.set_map_update_min_dist_diff(dist); p.set_map_update_min_angle_diff(ang)
    return eng, procs


def assert_same_maps(eng, procs, levels, streams=None):
    for s, p in enumerate(procs):
        gs = s if streams is None else streams[s]
        for l in range(levels):
            g, c = eng.download_level(gs, l), p.download_level(l)
            assert np.array_equal(g["updateIndex"], c["updateIndex"]), "stream %d level %d markers" % (s, l)
            assert np.array_equal(bits(g["logOddsVal"]), bits(c["logOddsVal"])), "stream %d level %d log-odds" % (s, l)


def test_ragged_empty_hints_flags_and_reset(hs, oracle):
    n, n_scans = 5, 36
    sc, ranges, pts, cnt, truth = make_streams(hs, n, n_scans, seed0=2024)
    ranges = ranges.copy()
    rng = np.random.default_rng(1)
    ranges[rng.random(ranges.shape) < 0.2] = np.inf
    ranges[5, 0, :] = np.inf; ranges[6, 0, :] = np.inf          # empty scans
    ranges[9, 2, 40:] = 0.01                                     # nearly empty
    pts, cnt = hs.scan_to_points_host(ranges, sc, 20.0)
    eng, procs = setup_pair(hs, oracle, n, dist=0.4, ang=0.9)
    poses = np.zeros((n, 3), np.float32)
    for k in range(n_scans):
        flags = np.array([(k % 8 == 7) and (s % 2 == 0) for s in range(n)], np.uint8)
        if k == 20:
            eng.reset(); [p.reset() for p in procs]; poses[:] = 0
        if k == 28:
            eng.reset(3); procs[3].reset(); poses[3] = 0
        hints = poses + (rng.normal(0, 0.01, poses.shape).astype(np.float32) if k % 5 == 4 else 0)
        hints = hints.astype(np.float32)
        out, cov = eng.update_batch(pts[k], cnt[k], pose_hints=hints, map_without_matching=flags)
        for s, p in enumerate(procs):
            ref = p.update(pts[k, s], cnt[k, s], hint=hints[s], map_without_matching=bool(flags[s]))
            assert np.array_equal(bits(out[s]), bits(ref)), (k, s, out[s], ref)
            if cnt[k, s] and not flags[s]:
                assert np.array_equal(bits(cov[s]), bits(p.cov()))
            assert eng.update_index(s, 0) == p.update_index(0)
        poses = out
        assert np.array_equal(bits(eng.last_map_update_poses()), bits(np.stack([p.last_map_update_pose() for p in procs])))
    assert_same_maps(eng, procs, 3)
    eng.close()


def test_beams_leaving_the_map_and_offset_origo(hs, oracle):
    """Small map + start near a corner: many end points and the start fall outside -> whole beams dropped."""
    n = 3
    sc, ranges, pts, cnt, truth = make_streams(hs, n, 12, seed0=99, scale_to_map=20.0)
    eng, procs = setup_pair(hs, oracle, n, size=128, levels=3, start=(0.2, 0.9))
    origos = np.array([[0.0, 0.0], [3.5, -2.25], [2000.0, 0.0]], np.float32)   # last: start cell outside the map
    poses = np.zeros((n, 3), np.float32)
    for k in range(12):
        out, _ = eng.update_batch(pts[k], cnt[k], origos=origos, pose_hints=poses)
        for s, p in enumerate(procs):
            ref = p.update(pts[k, s], cnt[k, s], origo=origos[s], hint=poses[s])
            assert np.array_equal(bits(out[s]), bits(ref))
        poses = out
    assert_same_maps(eng, procs, 3)
    assert (eng.download_level(2, 0)["updateIndex"] == -1).all()      # nothing was ever integrated for stream 2
    eng.close()


def test_stream_ids_subset_and_permutation(hs, oracle):
    n_eng, n_scans = 6, 10
    sc, ranges, pts, cnt, truth = make_streams(hs, 3, n_scans, seed0=7)
    ids = np.array([4, 0, 2], np.int32)
    eng, procs = setup_pair(hs, oracle, 3)
    eng.close()
    eng = hs.Engine(n_streams=n_eng, max_points=360)
    eng.set_update_factor_free(0.4); eng.set_update_factor_occupied(0.9)
    eng.set_map_update_min_dist_diff(0.0); eng.set_map_update_min_angle_diff(0.0)
    for k in range(n_scans):
        out, cov = eng.update_batch(pts[k], cnt[k], stream_ids=ids)
        for s, p in enumerate(procs):
            ref = p.update(pts[k, s], cnt[k, s], hint=(out[s] * 0 if k == 0 else prev[s]))
            assert np.array_equal(bits(out[s]), bits(ref))
            assert np.array_equal(bits(cov[s]), bits(p.cov())) or k == 0
        prev = out
    assert_same_maps(eng, procs, 3, streams=list(ids))
    for untouched in (1, 3, 5):
        assert (eng.download_level(untouched, 0)["updateIndex"] == -1).all()
        assert not eng.poses(untouched, 1).any()
    with pytest.raises(hs.HsError):
        eng.update_batch(pts[0], cnt[0], stream_ids=np.array([0, 1, 6], np.int32))
    # duplicate ids would put two CTAs on one stream's maps and counters: refused by every host entry point, state untouched
    before = eng.poses()
    dup = np.array([2, 5, 2], np.int32)
    for call in (lambda: eng.update_batch(pts[0], cnt[0], stream_ids=dup), lambda: eng.match_batch(pts[0], cnt[0], stream_ids=dup),
                 lambda: eng.raycast_batch(pts[0], cnt[0], truth[0], stream_ids=dup),
                 lambda: eng.raycast_batch(pts[0], cnt[0], truth[0], stream_ids=dup, refresh_coarse=False)):
        with pytest.raises(hs.HsError) as err:
            call()
        assert err.value.status == hs.HS_ERR_INVALID_ARG
    assert np.array_equal(bits(before), bits(eng.poses()))
    eng.close()


def test_ranges_entry_point_equals_points_entry_point(hs, oracle):
    """k_scan_to_points (LaserScan ranges on the device) == host conversion + update, incl. dropped ranges."""
    n, n_scans = 4, 15
    sc, ranges, pts, cnt, truth = make_streams(hs, n, n_scans, seed0=31)
    ranges = ranges.copy()
    rng = np.random.default_rng(2)
    ranges[rng.random(ranges.shape) < 0.25] = np.inf
    ranges[3, 1, :] = 0.0
    pts, cnt = hs.scan_to_points_host(ranges, sc, 20.0)
    eng, procs = setup_pair(hs, oracle, n)
    with pytest.raises(hs.HsError):
        eng.update_batch_ranges(ranges[0])
    eng.set_scan_geometry(sc.n_beams, sc.angle_min, sc.angle_increment, sc.range_min, sc.range_max)
    for k in range(n_scans):
        out, cov = eng.update_batch_ranges(ranges[k])
        for s, p in enumerate(procs):
            ref = p.update(pts[k, s], cnt[k, s], hint=p.pose())
            assert np.array_equal(bits(out[s]), bits(ref)), (k, s)
    assert_same_maps(eng, procs, 3)
    eng.close()
    # beam counts that are not a multiple of 4 take the kernel's scalar path (the 360-beam case above takes the float4 path)
    for nb in (359, 358, 5):
        with hs.Engine(n_streams=n, max_points=360) as e2, hs.Engine(n_streams=n, max_points=360) as e3:
            e2.set_scan_geometry(nb, sc.angle_min, sc.angle_increment, sc.range_min, sc.range_max)
            for k in range(3):
                out, _ = e2.update_batch_ranges(ranges[k][:, :nb])
                p2, c2 = hs.scan_to_points_host(ranges[k][:, :nb], sc, 20.0, max_points=360)
                ref, _ = e3.update_batch(p2, c2)
                assert np.array_equal(bits(out), bits(ref)), (nb, k)


def test_large_bounding_box_is_processed_in_bands(hs, oracle):
    """Scans whose bounding box exceeds the shared-memory cell map (here: a 2048^2 @0.01 m grid, box ~1000^2 cells,
    i.e. ~6 bands of rows on level 0) are ray-cast band by band; results stay bit-exact."""
    n, n_scans = 2, 6
    sc, ranges, pts, cnt, truth = make_streams(hs, n, n_scans, seed0=1234, scale_to_map=100.0)
    eng, procs = setup_pair(hs, oracle, n, size=2048, levels=2, res=0.01)
    for k in range(n_scans):
        out, _ = eng.update_batch(pts[k], cnt[k])
        for s, p in enumerate(procs):
            ref = p.update(pts[k, s], cnt[k, s], hint=p.pose())
            assert np.array_equal(bits(out[s]), bits(ref)), (k, s)
    assert_same_maps(eng, procs, 2)
    eng.close()


def test_teacher_forced_bands_all_octants(hs, oracle):
    """Teacher-forced integration of hand-made scans on a 4096^2 grid: long beams in every octant (incl. axis-parallel and
    exactly diagonal ones, shared end cells and beams leaving the grid) from several start cells, so that every band of the
    box sees x- and y-dominant traces entering and leaving it. Markers and log-odds bits must equal the CPU oracle's."""
    size, res, mp = 4096, 0.025, 256
    rng = np.random.default_rng(5)
    n = 4
    ang = np.linspace(0.0, 2.0 * np.pi, 200, endpoint=False)
    pts = np.zeros((n, mp, 2), np.float32)
    cnt = np.zeros(n, np.int32)
    for s in range(n):
        rad = rng.uniform(30.0, 1900.0, ang.size) if s < 3 else np.full(ang.size, 1200.0)
        p = np.stack([np.cos(ang) * rad, np.sin(ang) * rad], 1)
        extra = np.array([[1500, 0], [-1500, 0], [0, 1500], [0, -1500], [1000, 1000], [-1000, 1000], [1000, -1000], [-1000, -1000],
                          [1000, 1000], [3000, 10], [10, -3000], [0.2, 0.1], [1, 0]], np.float32)
        p = np.concatenate([p, extra]).astype(np.float32)
        pts[s, :len(p)] = p
        cnt[s] = len(p)
    poses = np.array([[0, 0, 0], [3.3, -2.1, 0.7], [-11.0, 7.5, -2.9], [0.01, 0.02, 3.1]], np.float32)
    origos = np.array([[0, 0], [4.0, -2.5], [0, 0], [-7.5, 3.25]], np.float32)
    with hs.Engine(res, size, size, (0.5, 0.5), 3, n_streams=n, max_points=mp) as eng:
        procs = [oracle.CpuProcessor("port", res, size, size, (0.5, 0.5), 3) for _ in range(n)]
        for rep in range(2):
            eng.raycast_batch(pts, cnt, poses, origos=origos)
            for s, p in enumerate(procs):
                p.raycast(pts[s], poses[s], n=int(cnt[s]), origo=origos[s])
        assert_same_maps(eng, procs, 3)
        for p in procs:
            p.close()


def test_published_map_info_matches_reference_geometry(hs, oracle):
    """setServiceGetMapData (HectorMappingRos.cpp:544-560): origin = getWorldCoords(0,0) - cellLength/2 per level."""
    for res, size, start, levels in ((0.05, 1024, (0.5, 0.5), 3), (0.025, 2048, (0.3, 0.7), 4), (0.1, 500, (0.5, 0.5), 2)):
        p = oracle.CpuProcessor("port", res, size, size, start, levels)
        with hs.Engine(res, size, size, start, levels, n_streams=1, max_points=8) as eng:
            for l in range(levels):
                mi = eng.map_info(l)
                sx, sy, cl = p.level_dims(l)
                w0 = p.map_to_world(l, np.zeros(3, np.float32))
                half = np.float32(cl) * np.float32(0.5)
                assert mi["width"] == sx and mi["height"] == sy and np.float32(mi["resolution"]) == np.float32(cl)
                assert np.float32(mi["origin"][0]) == np.float32(w0[0]) - half and np.float32(mi["origin"][1]) == np.float32(w0[1]) - half
        p.close()


def test_upload_download_classify_roundtrip(hs):
    with hs.Engine(0.05, 64, 64, (0.5, 0.5), 2, n_streams=2, max_points=8) as eng:
        rng = np.random.default_rng(0)
        for level, n in ((0, 64 * 64), (1, 32 * 32)):
            cells = np.zeros(n, hs.CELL_DTYPE)
            cells["logOddsVal"] = rng.normal(0, 1, n).astype(np.float32)
            cells["logOddsVal"][::7] = 0.0
            cells["updateIndex"] = rng.integers(-1, 100, n)
            eng.upload_level(1, level, cells)
            back = eng.download_level(1, level)
            assert np.array_equal(back, cells)
            occ = eng.classify_level(1, level)          # publishMap: isFree -> 0, isOccupied -> 100, else -1
            v = cells["logOddsVal"]
            assert np.array_equal(occ, np.where(v < 0, 0, np.where(v > 0, 100, -1)).astype(np.int8))
            untouched = eng.download_level(0, level)
            assert (untouched["logOddsVal"] == 0).all() and (untouched["updateIndex"] == -1).all()


def test_match_only_and_pose_hypotheses(hs, oracle):
    """hs_match_batch (no integration) and BASELINE config 4: 4096 pose hypotheses through getCompleteHessianDerivs."""
    n_scans = 25
    sc, ranges, pts, cnt, truth = make_streams(hs, 1, n_scans + 1, seed0=4096)
    eng, procs = setup_pair(hs, oracle, 1)
    p = procs[0]
    for k in range(n_scans):
        eng.update_batch(pts[k], cnt[k]); p.update(pts[k, 0], cnt[k, 0], hint=p.pose())
    before = [eng.download_level(0, l) for l in range(3)]
    hint = p.pose() + np.array([0.05, -0.03, 0.02], np.float32)
    out, cov = eng.match_batch(pts[n_scans], cnt[n_scans], pose_hints=hint[None])
    # oracle: matchData level by level (MapRepMultiMap::matchData)
    ref = hint.astype(np.float32)
    for level in (2, 1, 0):
        scaled = (pts[n_scans, 0, :cnt[n_scans, 0]] * np.float32(1.0 / (1 << level))).astype(np.float32)
        ref, rcov = p.match_level(level, ref, scaled, 5 if level == 0 else 3)
    assert np.array_equal(bits(out[0]), bits(ref)) and np.array_equal(bits(cov[0]), bits(rcov))
    for l in range(3):
        assert np.array_equal(eng.download_level(0, l), before[l])         # match only: maps untouched
    rng = np.random.default_rng(9)
    base = p.world_to_map(0, p.pose())
    hyp = (base[None] + rng.uniform(-1, 1, (4096, 3)) * np.array([20.0, 20.0, 0.5])).astype(np.float32)
    scan = pts[n_scans, 0, :cnt[n_scans, 0]]
    H, d = eng.hessian_derivs_batch(0, 0, hyp, scan)
    Hc, dc = p.hessian_derivs_batch(0, hyp, scan)
    assert np.array_equal(bits(H), bits(Hc)) and np.array_equal(bits(d), bits(dc))
    eng.close()


def test_checkpoint_restore_continues_bit_exact(hs, oracle):
    """hs_export_stream_state + hs_download_level of every level is a complete checkpoint: restored into another slot of
    another engine (after a reset elsewhere moved the counters), the stream continues bit for bit like the original --
    including the update markers, whose counter is not reset by reset() -- and like the CPU oracle."""
    n_scans, cut = 24, 11
    sc, ranges, pts, cnt, truth = make_streams(hs, 1, n_scans, seed0=4711)
    eng, procs = setup_pair(hs, oracle, 1, dist=0.2, ang=0.3)
    for k in range(cut):
        if k == 5:
            eng.reset(); procs[0].reset()        # counters keep counting across reset()
        out, _ = eng.update_batch(pts[k], cnt[k])
        procs[0].update(pts[k, 0], cnt[k, 0], hint=procs[0].pose())
    state, coarse = eng.export_stream_state(0)
    levels = [eng.download_level(0, l) for l in range(3)]
    with hs.Engine(n_streams=3, max_points=360) as other:
        other.set_update_factor_free(0.4); other.set_update_factor_occupied(0.9)
        other.set_map_update_min_dist_diff(0.2); other.set_map_update_min_angle_diff(0.3)
        other.import_stream_state(2, state, coarse)
        for l in range(3):
            other.upload_level(2, l, levels[l])
        for k in range(cut, n_scans):
            a, ca = eng.update_batch(pts[k], cnt[k])
            b, cb = other.update_batch(pts[k], cnt[k], stream_ids=[2])
            ref = procs[0].update(pts[k, 0], cnt[k, 0], hint=procs[0].pose())
            assert np.array_equal(bits(a[0]), bits(b[0])) and np.array_equal(bits(a[0]), bits(ref)) and np.array_equal(bits(ca), bits(cb)), k
        assert_same_maps(other, procs, 3, streams=[2])
        assert other.update_index(2, 0) == eng.update_index(0, 0) == procs[0].update_index(0)
    eng.close()
    for p in procs:
        p.close()


def test_offpath_members_on_random_maps(hs, oracle):
    """getResidualForState / getLikelihoodForState / getCovarianceForPose / interpMapValue(+WithDerivatives) and the
    occupancy ray query on random maps, incl. hypotheses and coordinates outside the map limits, empty point sets aside."""
    rng = np.random.default_rng(21)
    size = 256
    with hs.Engine(0.05, size, size, (0.5, 0.5), 2, n_streams=2, max_points=512) as eng:
        p = oracle.CpuProcessor("port", 0.05, size, size, (0.5, 0.5), 2)
        for l in range(2):
            n = (size >> l) ** 2
            cells = np.zeros(n, hs.CELL_DTYPE)
            cells["logOddsVal"] = rng.normal(0, 2, n).astype(np.float32)
            cells["logOddsVal"][rng.random(n) < 0.5] = 0.0
            cells["updateIndex"] = -1
            eng.upload_level(1, l, cells); p.upload_level(l, cells)
            lim = size >> l
            pts = rng.uniform(-0.3 * lim, 0.3 * lim, (417, 2)).astype(np.float32)
            poses = np.concatenate([rng.uniform(-0.1 * lim, 1.1 * lim, (64, 2)), rng.uniform(-4, 4, (64, 1))], 1).astype(np.float32)
            r, lh = eng.likelihood_batch(1, l, poses, pts)
            r_ref = np.array([p.residual_for_state(l, q, pts) for q in poses], np.float32)
            lh_ref = np.array([p.likelihood_for_state(l, q, pts) for q in poses], np.float32)
            assert np.array_equal(bits(r), bits(r_ref)) and np.array_equal(bits(lh), bits(lh_ref))
            for q in poses[:4]:
                cm, cw = eng.covariance_for_pose(1, l, q, pts)
                cm_ref, cw_ref = p.covariance_for_pose(l, q, pts)
                assert np.array_equal(bits(cm), bits(cm_ref)) and np.array_equal(bits(cw), bits(cw_ref))
            coords = rng.uniform(-3, lim + 3, (300, 2)).astype(np.float32)
            coords[:4] = [(0, 0), (lim - 2, lim - 2), (lim - 1.999, 5), (5, -0.0001)]
            iv, dv = eng.interp_map_value_batch(1, l, coords)
            iv_ref = np.array([p.interp_map_value(l, c[0], c[1]) for c in coords], np.float32)
            dv_ref = np.stack([p.interp_with_derivs(l, c[0], c[1]) for c in coords])
            assert np.array_equal(bits(iv), bits(iv_ref)) and np.array_equal(bits(dv), bits(dv_ref))
            v = cells["logOddsVal"].reshape(lim, lim)
            occ = np.where(v < 0, 0, np.where(v > 0, 100, -1)).astype(np.int8)
            occ_sparse = occ.copy()
            begins = rng.integers(-2, lim + 2, (200, 2)).astype(np.int32)
            ends = rng.integers(-2, lim + 2, (200, 2)).astype(np.int32)
            dist, hit = eng.occupancy_ray_batch(1, l, begins, ends)
            for i in range(200):
                d_ref, h_ref = oracle.check_occupancy_bresenhami("port", occ_sparse, begins[i], ends[i])
                assert dist[i] == d_ref and np.array_equal(hit[i], h_ref if h_ref is not None else (-1, -1)), i
        p.close()


def test_pinned_host_batches_zero_copy_and_chunked_are_identical(hs, monkeypatch):
    """With pinned host buffers hs_update_batch lets the matcher read the scans straight from host memory (zero-copy, the
    default) or, with HSGPU_E2E_ZERO_COPY=0, cuts the batch into chunks on several CUDA streams (copy of chunk c+1 under the
    kernels of chunk c), or (HSGPU_E2E_DEFER=1/2, the default is 1) returns as soon as the poses are in the caller's buffer
    and lets the map integration run on while the next call's points are prefetched. Results must not depend on the
    transport: zero-copy, 1, 2, 3 and 4 chunks, deferred with a DMA or a copy-kernel prefetch, with and without stream_ids,
    all bit-identical -- and the poses must be valid the moment the call returns."""
    import torch
    n, n_scans = 300, 5
    sc, ranges, pts, cnt, truth = make_streams(hs, n, n_scans, seed0=9000)
    h_pts = torch.from_numpy(pts).pin_memory()
    h_cnt = torch.from_numpy(cnt).pin_memory()
    ids = torch.from_numpy(np.random.default_rng(3).permutation(n).astype(np.int32)).pin_memory()
    results = []
    for chunks in ("1", "2", "3", "4", "zero-copy", "defer-1", "defer-2"):
        monkeypatch.setenv("HSGPU_E2E_ZERO_COPY", "0" if chunks.isdigit() else "1")
        monkeypatch.setenv("HSGPU_E2E_CHUNKS", chunks if chunks.isdigit() else "2")
        monkeypatch.setenv("HSGPU_E2E_DEFER", chunks[-1] if chunks.startswith("defer") else "0")
        for use_ids in (False, True):
            with hs.Engine(n_streams=n, max_points=360) as eng:
                eng.set_update_factor_free(0.4); eng.set_update_factor_occupied(0.9)
                eng.set_map_update_min_dist_diff(0.0); eng.set_map_update_min_angle_diff(0.0)
                out = torch.full((n_scans, n, 3), float("nan"), dtype=torch.float32).pin_memory()
                at_return = []
                for k in range(n_scans):
                    eng.update_batch_raw(n, h_pts[k].data_ptr(), h_cnt[k].data_ptr(), out[k].data_ptr(),
                                         ids_ptr=ids.data_ptr() if use_ids else None)
                    at_return.append(out[k].numpy().copy())   # no further synchronisation: valid when the call returns
                poses = np.stack(at_return)
                if use_ids:   # slot i updated stream ids[i]: bring back to stream order
                    inv = np.empty(n, np.int64); inv[ids.numpy()] = np.arange(n)
                    final = eng.poses()
                    maps = [eng.download_level(int(ids[s]), 0)["logOddsVal"].copy() for s in (0, n // 2, n - 1)]
                else:
                    final = eng.poses()
                    maps = [eng.download_level(s, 0)["logOddsVal"].copy() for s in (0, n // 2, n - 1)]
                results.append((chunks, use_ids, poses, final, maps))
    base = results[0]
    for chunks, use_ids, poses, final, maps in results[1:]:
        assert np.array_equal(bits(poses), bits(base[2])), (chunks, use_ids)     # slot i always carries scan stream i's points
        for a, b in zip(maps, base[4]):
            assert np.array_equal(bits(a), bits(b)), (chunks, use_ids)


def test_point_cloud_entry_point_equals_oracle_pipeline(hs, oracle):
    """hs_update_batch_clouds = rosPointCloudToDataContainer on the GPU + update(): against the oracle's cloud conversion
    followed by the CPU processor (points AND offset origo), bit for bit, ragged clouds incl. an empty one."""
    from test_scangen import _laser_transform
    rng = np.random.default_rng(31)
    n, n_scans, max_cloud = 5, 8, 400
    sc, ranges, pts, cnt, truth = make_streams(hs, n, n_scans, seed0=606)
    eng, procs = setup_pair(hs, oracle, n, max_points=360)
    filt = (np.float32(0.16), np.float32(900.0), np.float32(-1.0), np.float32(1.0))
    for k in range(n_scans):
        clouds = np.zeros((n, max_cloud, 3), np.float32)
        counts = np.zeros(n, np.int32)
        xf = np.zeros((n, 12), np.float64)
        for s in range(n):
            m = int(cnt[k, s]) if not (k == 3 and s == 2) else 0
            # laser-frame cloud in metres from the synthetic scan points (cells / 20), plus a z spread that the z window cuts
            clouds[s, :m, :2] = pts[k, s, :m] / np.float32(20.0)
            clouds[s, :m, 2] = rng.uniform(-1.2, 1.2, m).astype(np.float32)
            counts[s] = m
            t = _laser_transform(rng); t[:9] = np.eye(3).reshape(-1); t[9:] = (0.1 * s, -0.05 * s, 0.3)   # pure offset mount
            xf[s] = t
        out, cov = eng.update_batch_clouds(clouds, counts, xf, *filt)
        for s, p in enumerate(procs):
            cp, co = oracle.cloud_to_points(clouds[s, :counts[s]], xf[s], np.float32(20.0), *filt)
            ref = p.update(cp, len(cp), origo=co, hint=p.pose())
            assert np.array_equal(bits(out[s]), bits(ref)), (k, s)
    assert_same_maps(eng, procs, 3)
    eng.close()
    for p in procs:
        p.close()


def test_odd_grid_sizes_use_the_generic_sweep(hs, oracle):
    """Grid widths that are not multiples of 4 (500 -> 250 -> 125 cells; 251 -> 125 -> 62) cannot use the 16-byte quad sweep:
    the per-cell path must give the same bits, closed loop and teacher-forced, rectangular maps and off-centre start included."""
    for size_x, size_y, levels, res, start in ((500, 500, 3, 0.05, (0.5, 0.5)), (251, 333, 3, 0.1, (0.4, 0.6))):
        n, n_scans = 3, 10
        sc, ranges, pts, cnt, truth = make_streams(hs, n, n_scans, seed0=8800 + size_x, scale_to_map=1.0 / res)
        eng = hs.Engine(res, size_x, size_y, start, levels, n_streams=n, max_points=360)
        procs = [oracle.CpuProcessor("port", res, size_x, size_y, start, levels) for _ in range(n)]
        for p in procs + [eng]:
            p.set_update_factor_free(0.4); p.set_update_factor_occupied(0.9)
            p.set_map_update_min_dist_diff(0.0); p.set_map_update_min_angle_diff(0.0)
        for k in range(n_scans):
            out, _ = eng.update_batch(pts[k], cnt[k])
            for s, p in enumerate(procs):
                ref = p.update(pts[k, s], cnt[k, s], hint=p.pose())
                assert np.array_equal(bits(out[s]), bits(ref)), (size_x, k, s)
        eng.raycast_batch(pts[0], cnt[0], out)
        for s, p in enumerate(procs):
            p.raycast(pts[0, s], out[s], n=int(cnt[0, s]))
        assert_same_maps(eng, procs, levels)
        eng.close()
        for p in procs:
            p.close()


def test_pipelined_submission_equals_synchronous_updates(hs):
    """hs_submit_batch / hs_wait: scans of the same streams submitted back to back without waiting (inputs double-buffered
    on a copy stream) give bit for bit the poses and maps of synchronous hs_update_batch calls -- full batches, a stream
    subset with stream_ids, per-scan pose buffers -- and pageable buffers degrade to the synchronous call."""
    import torch
    n, n_scans = 200, 12
    sc, ranges, pts, cnt, truth = make_streams(hs, n, n_scans, seed0=12345)
    h_pts = torch.from_numpy(pts).pin_memory()
    h_cnt = torch.from_numpy(cnt).pin_memory()
    sub = torch.from_numpy(np.arange(0, n, 3, dtype=np.int32)).pin_memory()
    sub_pts = torch.from_numpy(np.ascontiguousarray(pts[:, ::3])).pin_memory()
    sub_cnt = torch.from_numpy(np.ascontiguousarray(cnt[:, ::3])).pin_memory()
    def make():
        eng = hs.Engine(n_streams=n, max_points=360)
        eng.set_update_factor_free(0.4); eng.set_update_factor_occupied(0.9)
        eng.set_map_update_min_dist_diff(0.0); eng.set_map_update_min_angle_diff(0.0)
        return eng
    with make() as a, make() as b:
        ref = np.zeros((n_scans, n, 3), np.float32)
        for k in range(n_scans):
            if k % 4 == 3:   # every 4th scan only for a third of the streams
                out, _ = a.update_batch(pts[k, ::3], cnt[k, ::3], stream_ids=sub.numpy(), want_cov=False)
                ref[k, ::3] = out
            else:
                ref[k], _ = a.update_batch(pts[k], cnt[k], want_cov=False)
        got = torch.zeros((n_scans, n, 3), dtype=torch.float32).pin_memory()
        got_sub = torch.zeros((n_scans, len(sub), 3), dtype=torch.float32).pin_memory()
        for k in range(n_scans):   # all submitted before anything is waited for
            if k % 4 == 3:
                b.submit_batch_raw(len(sub), sub_pts[k].data_ptr(), sub_cnt[k].data_ptr(), got_sub[k].data_ptr(), ids_ptr=sub.data_ptr())
            else:
                b.submit_batch_raw(n, h_pts[k].data_ptr(), h_cnt[k].data_ptr(), got[k].data_ptr())
        b.wait()
        for k in range(n_scans):
            if k % 4 == 3:
                assert np.array_equal(bits(got_sub[k].numpy()), bits(ref[k, ::3])), k
            else:
                assert np.array_equal(bits(got[k].numpy()), bits(ref[k])), k
        assert np.array_equal(bits(a.poses()), bits(b.poses()))
        for s_ in (0, 3, n - 1):
            for l in range(3):
                assert np.array_equal(a.download_level(s_, l), b.download_level(s_, l))
        # pageable buffers: same results through the synchronous fallback
        out_pageable = np.zeros((n, 3), np.float32)
        extra_pts, extra_cnt = np.ascontiguousarray(pts[0]), np.ascontiguousarray(cnt[0])
        b.submit_batch_raw(n, extra_pts.ctypes.data, extra_cnt.ctypes.data, out_pageable.ctypes.data)
        ref_extra, _ = a.update_batch(pts[0], cnt[0], want_cov=False)
        assert np.array_equal(bits(out_pageable), bits(ref_extra))


def test_fuzz_random_scans_teacher_forced_and_matched(hs, oracle):
    """Randomised configurations (grid sizes incl. odd and rectangular ones, 1-4 levels, scan lengths around the matcher's
    register-variant boundaries, random points with duplicates / zero-length / out-of-map beams, random poses and origos,
    both update-factor sets): two teacher-forced integrations, one match-only call and one full update per configuration,
    everything bit for bit against the oracle."""
    # HSGPU_FUZZ_SEED / HSGPU_FUZZ_TRIALS: longer campaigns with other seeds (tests/tools/gpu_fuzz_campaign.sh)
    rng = np.random.default_rng(int(os.environ.get("HSGPU_FUZZ_SEED", "20261020")))
    n_trials = int(os.environ.get("HSGPU_FUZZ_TRIALS", "24"))
    sizes = [(64, 64), (128, 96), (250, 250), (333, 200), (512, 512), (1024, 1024), (1000, 1000), (2048, 2048)]
    n_checked = 0
    for trial in range(n_trials):
        sx, sy = sizes[trial % len(sizes)]
        levels = int(rng.integers(1, 5))
        while (min(sx, sy) >> (levels - 1)) < 8:
            levels -= 1
        res = float(rng.choice([0.025, 0.05, 0.1]))
        mp = int(rng.choice([17, 128, 129, 256, 257, 360, 384, 385, 700]))
        n = 3
        start = (float(rng.uniform(0.3, 0.7)), float(rng.uniform(0.3, 0.7)))
        free, occ = (0.4, 0.9) if trial % 2 else (0.4, 0.6)
        eng = hs.Engine(res, sx, sy, start, levels, n_streams=n, max_points=mp)
        procs = [oracle.CpuProcessor("port", res, sx, sy, start, levels) for _ in range(n)]
        for p in procs + [eng]:
            p.set_update_factor_free(free); p.set_update_factor_occupied(occ)
            p.set_map_update_min_dist_diff(0.0); p.set_map_update_min_angle_diff(0.0)
        reach = 0.45 * min(sx, sy)
        pts = np.zeros((n, mp, 2), np.float32)
        cnt = rng.integers(0, mp + 1, n).astype(np.int32)
        cnt[0] = mp
        for s in range(n):
            ang = rng.uniform(0, 2 * np.pi, mp)
            rad = rng.uniform(0.0, reach * 1.3, mp)          # some beams end outside the map
            pts[s, :, 0] = np.cos(ang) * rad; pts[s, :, 1] = np.sin(ang) * rad
            pts[s, ::7] = pts[s, 1]                            # shared end cells
            pts[s, ::11] = 0.0                                 # zero-length beams (begin == end)
        poses = np.concatenate([rng.uniform(-0.1, 0.1, (n, 2)) * res * min(sx, sy), rng.uniform(-3.2, 3.2, (n, 1))], 1).astype(np.float32)
        origos = rng.uniform(-3, 3, (n, 2)).astype(np.float32)
        for rep in range(2):
            eng.raycast_batch(pts, cnt, poses, origos=origos)
            for s, p in enumerate(procs):
                p.raycast(pts[s], poses[s], n=int(cnt[s]), origo=origos[s])
        assert_same_maps(eng, procs, levels)
        hint = poses + rng.uniform(-0.05, 0.05, (n, 3)).astype(np.float32)
        got, gcov = eng.match_batch(pts, cnt, origos=origos, pose_hints=hint)
        out, cov = eng.update_batch(pts, cnt, origos=origos, pose_hints=hint)
        # A singular Hessian (few or degenerate points) gives inf/NaN poses; the reference then indexes its grid with
        # (int)NaN and crashes, so parity is only defined -- and the oracle only called -- where the GPU result is finite
        # (the arithmetic is identical, so the oracle's is finite too). The engine itself treats NaN coordinates as out of map.
        ok = np.isfinite(got).all(axis=1) & np.isfinite(out).all(axis=1) & np.isfinite(gcov.reshape(n, -1)).all(axis=1)
        for s, p in enumerate(procs):
            if not ok[s]:
                continue
            ref = p.update(pts[s], int(cnt[s]), origo=origos[s], hint=hint[s])
            assert np.array_equal(bits(got[s]), bits(ref)), (trial, s, "match")
            assert np.array_equal(bits(out[s]), bits(ref)), (trial, s, "update")
        if ok.all():
            assert_same_maps(eng, procs, levels)
        n_checked = n_checked + int(ok.sum())
        eng.close()
        for p in procs:
            p.close()
    assert n_checked >= (40 * n_trials) // 24, n_checked   # most of the 3 x n_trials random streams give a finite match


def test_pinned_ranges_entry_point_deferred_and_synchronous_identical(hs, monkeypatch):
    """hs_update_batch_ranges with pinned buffers: the deferred-integration transport (returns after the match, ranges of the
    next call prefetched by DMA) and the synchronous one give the same poses -- valid at return -- and the same maps as the
    points entry point fed with the host conversion of the same ranges."""
    import torch
    n, n_scans = 200, 6
    sc, ranges, pts, cnt, truth = make_streams(hs, n, n_scans, seed0=515)
    ranges = ranges.copy()
    ranges[np.random.default_rng(2).random(ranges.shape) < 0.2] = np.inf
    pts, cnt = hs.scan_to_points_host(ranges, sc, 20.0)
    h_rng = torch.from_numpy(ranges).pin_memory()
    results = []
    for defer in ("0", "1", "points"):
        monkeypatch.setenv("HSGPU_E2E_DEFER", "0" if defer == "points" else defer)
        with hs.Engine(n_streams=n, max_points=360) as eng:
            eng.set_update_factor_free(0.4); eng.set_update_factor_occupied(0.9)
            eng.set_map_update_min_dist_diff(0.0); eng.set_map_update_min_angle_diff(0.0)
            eng.set_scan_geometry(sc.n_beams, sc.angle_min, sc.angle_increment, sc.range_min, sc.range_max)
            out = torch.full((n_scans, n, 3), float("nan"), dtype=torch.float32).pin_memory()
            at_return = []
            for k in range(n_scans):
                if defer == "points":
                    o, _ = eng.update_batch(pts[k], cnt[k], want_cov=False)
                    at_return.append(o.copy())
                else:
                    eng.update_batch_ranges_raw(n, h_rng[k].data_ptr(), out[k].data_ptr())
                    at_return.append(out[k].numpy().copy())
            maps = [eng.download_level(s, l) for s in (0, n - 1) for l in (0, 2)]
            results.append((np.stack(at_return), eng.poses().copy(), [(m["logOddsVal"].copy(), m["updateIndex"].copy()) for m in maps]))
    for r in results[:-1]:
        assert np.array_equal(bits(r[0]), bits(results[-1][0])) and np.array_equal(bits(r[1]), bits(results[-1][1]))
        for (a, ai), (b, bi) in zip(r[2], results[-1][2]):
            assert np.array_equal(bits(a), bits(b)) and np.array_equal(ai, bi)


@pytest.mark.parametrize("max_points", [40, 200, 360, 700, 900])
def test_every_matcher_instantiation_with_ragged_counts(hs, oracle, max_points):
    """The matcher is compiled once per points-per-thread class (max_points <= 256: 2, <= 384: 3, <= 768: 6; above: k_match2)
    and its cache-miss pass uses full-warp ballots: scans whose length ends inside a warp, inside the first warp, exactly on a
    warp or CTA boundary, or at 0 / 1 / max_points points must neither hang nor differ from the oracle."""
    n, n_scans = 12, 5
    sc, ranges, pts, cnt, truth = make_streams(hs, n, 2 * n_scans, seed0=6100 + max_points)
    # scans longer than 360 points: two consecutive scans of a stream back to back (same room, nearly the same pose)
    long_pts = np.concatenate([pts[0::2], pts[1::2]], axis=2)[:, :, :max_points].copy()      # [n_scans][n][<=720][2]
    if long_pts.shape[2] < max_points:
        long_pts = np.concatenate([long_pts, long_pts[:, :, :max_points - long_pts.shape[2]]], axis=2)
    lengths = [0, 30, 31, 32, 33, 64, 127, 128, 129, max_points - 1, max_points, max_points // 2]
    counts = np.array([[min(lengths[(s + k) % len(lengths)], max_points) for s in range(n)] for k in range(n_scans)], np.int32)
    counts[0] = max_points     # a full first scan, so that there is a map to match the ragged ones against
    eng, procs = setup_pair(hs, oracle, n, max_points=max_points)
    for k in range(n_scans):
        out, cov = eng.update_batch(np.ascontiguousarray(long_pts[k]), counts[k])
        for s, p in enumerate(procs):
            ref = p.update(long_pts[k, s], counts[k, s], hint=p.pose())
            assert np.array_equal(bits(out[s]), bits(ref)), (k, s, counts[k, s])
            if counts[k, s]:
                assert np.array_equal(bits(cov[s]), bits(p.cov())), (k, s, counts[k, s])
    assert_same_maps(eng, procs, 3)
    # 1..3 points: a singular Hessian, where the reference itself indexes its grid with (int)NaN -- no oracle, just no hang
    tiny = np.array([1 + (s % 3) for s in range(n)], np.int32)
    eng.update_batch(np.ascontiguousarray(long_pts[0]), tiny)
    eng.synchronize()
    eng.close()
    for p in procs:
        p.close()
