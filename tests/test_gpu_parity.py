"""GPU parity tests proper: the CUDA path (through the C ABI) against the CPU oracle on identical inputs.

Bars (BASELINE.json north_star): Bresenham cell sets and update markers bit-exact (teacher-forced);
matched poses within 1e-4 m / 1e-4 rad per scan; log-odds within 1e-5. On top of that the engine's
strict-summation mode must reproduce the oracle bit for bit in closed loop.
"""
import numpy as np
import pytest

from hs_testutil import bits, make_streams

pytestmark = pytest.mark.gpu

NODE = dict(free=0.4, occ=0.9, dist=0.4, ang=0.9)  # HectorMappingRos.cpp:72-76 defaults


def configure(p, free, occ, dist, ang):
    p.set_update_factor_free(free)
    p.set_update_factor_occupied(occ)
    p.set_map_update_min_dist_diff(dist)
    p.set_map_update_min_angle_diff(ang)


def cpu_streams(oracle, n, size=1024, levels=3, res=0.05, **kw):
    procs = [oracle.CpuProcessor("port", res, size, size, (0.5, 0.5), levels) for _ in range(n)]
    for p in procs:
        configure(p, **kw)
    return procs


def compare_maps(eng, procs, levels, exact=True, tol=1e-5):
    n_diff_idx = 0
    max_dv = 0.0
    for s, p in enumerate(procs):
        for l in range(levels):
            g = eng.download_level(s, l)
            c = p.download_level(l)
            n_diff_idx += int((g["updateIndex"] != c["updateIndex"]).sum())
            if exact:
                assert np.array_equal(bits(g["logOddsVal"]), bits(c["logOddsVal"])), "stream %d level %d log-odds bits differ" % (s, l)
            else:
                max_dv = max(max_dv, float(np.abs(g["logOddsVal"] - c["logOddsVal"]).max()))
    return n_diff_idx, max_dv


def test_hessian_derivs_parity(hs, oracle):
    """getCompleteHessianDerivs on a map built by the oracle: strict mode bit-exact, fast mode ~1e-6 relative."""
    n_scans = 40
    sc, ranges, pts, cnt, truth = make_streams(hs, 1, n_scans)
    cpu = cpu_streams(oracle, 1, dist=0.0, ang=0.0, free=0.4, occ=0.9)[0]
    pose = np.zeros(3, np.float32)
    for k in range(n_scans):
        pose = cpu.update(pts[k, 0], cnt[k, 0], hint=pose)
    with hs.Engine(n_streams=1, max_points=360) as eng:
        for l in range(3):
            eng.upload_level(0, l, cpu.download_level(l))
        rng = np.random.default_rng(7)
        for level in range(3):
            base = cpu.world_to_map(level, pose)
            poses = base[None, :] + rng.uniform(-1, 1, (256, 3)).astype(np.float32) * np.array([20, 20, 0.5], np.float32) / (1 << level)
            poses = poses.astype(np.float32)
            p = (pts[n_scans - 1, 0, :cnt[n_scans - 1, 0]] * np.float32(1.0 / (1 << level))).astype(np.float32)
            Hc, dc = cpu.hessian_derivs_batch(level, poses, p)
            eng.set_strict_summation(True)
            Hs, ds = eng.hessian_derivs_batch(0, level, poses, p)
            assert np.array_equal(bits(Hs), bits(Hc)) and np.array_equal(bits(ds), bits(dc)), "strict mode must be bit-exact"
            eng.set_strict_summation(False)
            Hf, df = eng.hessian_derivs_batch(0, level, poses, p)
            scale = np.abs(Hc).max(axis=(1, 2), keepdims=True) + 1e-6
            assert np.abs(Hf - Hc).max() / scale.max() < 1e-5
            assert np.abs((Hf - Hc) / scale).max() < 1e-5
            dscale = np.abs(dc).max(axis=1, keepdims=True) + 1e-3
            assert np.abs((df - dc) / dscale).max() < 1e-4


def test_hessian_derivs_4096_hypotheses_host_and_device_entry_points(hs, oracle):
    """BASELINE configs[3] at its full size: 4096 pose hypotheses of one 360-beam scan through getCompleteHessianDerivs,
    strict mode bit-exact against the oracle; every poses-per-warp instantiation; the device-pointer entry point; and the
    probability block plane follows the map when it changes (update, upload, reset)."""
    import os
    import torch
    n_scans, n_hyp = 30, 4096
    sc, ranges, pts, cnt, truth = make_streams(hs, 2, n_scans + 6, seed0=31337)
    cpus = cpu_streams(oracle, 2, dist=0.0, ang=0.0, free=0.4, occ=0.9)
    rng = np.random.default_rng(11)
    old = os.environ.get("HSGPU_HD_PPW")
    try:
        for ppw in ("0", "1", "2", "3"):
            os.environ["HSGPU_HD_PPW"] = ppw
            with hs.Engine(n_streams=2, max_points=360) as eng:
                configure(eng, dist=0.0, ang=0.0, free=0.4, occ=0.9)
                if ppw == "0":
                    for k in range(n_scans):
                        eng.update_batch(pts[k], cnt[k])
                        for s, c in enumerate(cpus):
                            c.update(pts[k, s], cnt[k, s], hint=c.pose())
                else:
                    for s, c in enumerate(cpus):
                        for l in range(3):
                            eng.upload_level(s, l, c.download_level(l))
                for s in (1, 0):
                    base = cpus[s].world_to_map(0, cpus[s].pose())
                    poses = (base[None, :] + rng.uniform(-1, 1, (n_hyp, 3)).astype(np.float32) * np.array([20, 20, 0.5], np.float32)).astype(np.float32)
                    poses[7] = (-50.0, 3000.0, 0.3)          # every point out of the map: H = 0, dTr = 0
                    k = n_scans + s
                    p = pts[k, s, :cnt[k, s]].copy()
                    Hc, dc = cpus[s].hessian_derivs_batch(0, poses, p)
                    Hg, dg = eng.hessian_derivs_batch(s, 0, poses, p)
                    assert np.array_equal(bits(Hg), bits(Hc)) and np.array_equal(bits(dg), bits(dc)), (ppw, s)
                    assert not Hg[7].any() and not dg[7].any()
                    if ppw in ("0", "3"):
                        d_pose = torch.from_numpy(poses).cuda(); d_pts = torch.from_numpy(p).cuda()
                        d_H = torch.zeros((n_hyp, 9), device="cuda"); d_d = torch.zeros((n_hyp, 3), device="cuda")
                        eng.hessian_derivs_batch_device(s, 0, d_pose.data_ptr(), n_hyp, d_pts.data_ptr(), p.shape[0], d_H.data_ptr(), d_d.data_ptr())
                        eng.synchronize()
                        assert np.array_equal(bits(d_H.cpu().numpy().reshape(n_hyp, 3, 3)), bits(Hc)) and np.array_equal(bits(d_d.cpu().numpy()), bits(dc))
                if ppw == "0":
                    # the plane must follow the map: another update, then an upload of a foreign map, then a reset
                    k = n_scans + 2
                    eng.update_batch(pts[k], cnt[k])
                    for s, c in enumerate(cpus):
                        c.update(pts[k, s], cnt[k, s], hint=c.pose())
                    sub = poses[:300]
                    Hc, dc = cpus[0].hessian_derivs_batch(0, sub, p)
                    Hg, dg = eng.hessian_derivs_batch(0, 0, sub, p)
                    assert np.array_equal(bits(Hg), bits(Hc)) and np.array_equal(bits(dg), bits(dc))
                    eng.upload_level(0, 0, cpus[1].download_level(0))
                    Hc1, dc1 = cpus[1].hessian_derivs_batch(0, sub, p)
                    Hg, dg = eng.hessian_derivs_batch(0, 0, sub, p)
                    assert np.array_equal(bits(Hg), bits(Hc1)) and np.array_equal(bits(dg), bits(dc1))
                    eng.reset(0)
                    Hg, dg = eng.hessian_derivs_batch(0, 0, sub, p)
                    blank = oracle.CpuProcessor("port")
                    Hb, db = blank.hessian_derivs_batch(0, sub, p)
                    assert np.array_equal(bits(Hg), bits(Hb)) and np.array_equal(bits(dg), bits(db))
                    # fast (tree) mode stays within tolerance at the full size
                    eng.upload_level(0, 0, cpus[0].download_level(0))
                    eng.set_strict_summation(False)
                    Hf, df = eng.hessian_derivs_batch(0, 0, poses, p)
                    Hc, dc = cpus[0].hessian_derivs_batch(0, poses, p)
                    scale = np.abs(Hc).max(axis=(1, 2), keepdims=True) + 1e-6
                    assert np.abs((Hf - Hc) / scale).max() < 1e-5
    finally:
        if old is None:
            os.environ.pop("HSGPU_HD_PPW", None)
        else:
            os.environ["HSGPU_HD_PPW"] = old


def test_raycast_teacher_forced_bitexact(hs, oracle):
    """updateByScan with oracle-given poses: every cell's marker AND log-odds bits identical."""
    n_streams, n_scans = 6, 30
    sc, ranges, pts, cnt, truth = make_streams(hs, n_streams, n_scans, seed0=77)
    procs = cpu_streams(oracle, n_streams, **NODE)
    with hs.Engine(n_streams=n_streams, max_points=360) as eng:
        configure(eng, **NODE)
        rng = np.random.default_rng(3)
        for k in range(n_scans):
            poses = (truth[k] + rng.normal(0, 0.01, truth[k].shape)).astype(np.float32)
            for s, p in enumerate(procs):
                p.raycast(pts[k, s], poses[s], cnt[k, s])
            eng.raycast_batch(pts[k], cnt[k], poses)
        n_diff, _ = compare_maps(eng, procs, 3, exact=True)
        assert n_diff == 0
        for s, p in enumerate(procs):
            assert eng.update_index(s, 0) == p.update_index(0) == n_scans - 1


def test_integrate_uses_the_containers_of_the_last_match(hs, oracle):
    """hs_integrate_batch = MapRepMultiMap::updateByScan proper (MapRepMultiMap.h:134-147): match scan A, integrate scan B ->
    level 0 gets B, the coarse levels get A (stale); before any match they get nothing. Against the reference itself
    (oracle/_ref) when it is on the box. hs_raycast_batch (teacher forcing) refreshes the containers and must differ."""
    kind = "ref" if oracle.have("ref") else "port"
    n = 3
    sc, ranges, pts, cnt, truth = make_streams(hs, n, 12, seed0=5150)
    procs = [oracle.CpuProcessor(kind) for _ in range(n)]
    for p in procs:
        configure(p, **NODE)
    with hs.Engine(n_streams=n, max_points=360) as eng, hs.Engine(n_streams=n, max_points=360) as eng_tf:
        configure(eng, **NODE); configure(eng_tf, **NODE)
        # no match yet: the coarse containers are empty, only level 0 is integrated
        for s, p in enumerate(procs):
            p.integrate(pts[0, s], truth[0, s], cnt[0, s])
        eng.raycast_batch(pts[0], cnt[0], truth[0], refresh_coarse=False)
        eng_tf.raycast_batch(pts[0], cnt[0], truth[0])
        for k in range(1, 11, 2):
            hint = truth[k]
            for s, p in enumerate(procs):
                p.refresh_coarse(pts[k, s], cnt[k, s])               # matchData of scan k (pose discarded)
                p.integrate(pts[k + 1, s], truth[k + 1, s], cnt[k + 1, s])   # updateByScan of scan k+1
            eng.match_batch(pts[k], cnt[k], pose_hints=hint)
            eng.raycast_batch(pts[k + 1], cnt[k + 1], truth[k + 1], refresh_coarse=False)
            eng_tf.raycast_batch(pts[k + 1], cnt[k + 1], truth[k + 1])
        n_diff, _ = compare_maps(eng, procs, 3, exact=True)
        assert n_diff == 0
        coarse_differs = any(not np.array_equal(eng.download_level(s, l)["updateIndex"], eng_tf.download_level(s, l)["updateIndex"])
                             for s in range(n) for l in (1, 2))
        assert coarse_differs, "the stale-container path must be distinguishable from teacher forcing"
        assert all(np.array_equal(eng.download_level(s, 0)["updateIndex"], eng_tf.download_level(s, 0)["updateIndex"]) for s in range(n))


def test_short_stream_against_the_reference_itself(hs, oracle):
    """Closes the chain on the GPU box: the CUDA path against oracle/_ref (the unmodified reference headers) directly, not via
    the C restatement. Skipped only where _ref was never built."""
    if not oracle.have("ref"):
        pytest.skip("oracle/_ref not built on this box")
    n_streams, n_scans = 4, 60
    sc, ranges, pts, cnt, truth = make_streams(hs, n_streams, n_scans, seed0=2468)
    procs = [oracle.CpuProcessor("ref") for _ in range(n_streams)]
    for p in procs:
        configure(p, **NODE)
    _, cpu_poses, cpu_cov = oracle.run_streams(procs, pts, cnt, threads=2, want_cov=True)
    with hs.Engine(n_streams=n_streams, max_points=360) as eng:
        configure(eng, **NODE)
        for k in range(n_scans):
            poses, cov = eng.update_batch(pts[k], cnt[k])
            assert np.array_equal(bits(poses), bits(cpu_poses[k])), k
            assert np.array_equal(bits(cov.reshape(n_streams, 9)), bits(cpu_cov[k])), k
        n_diff, _ = compare_maps(eng, procs, 3, exact=True)
        assert n_diff == 0


def test_closed_loop_strict_bitexact(hs, oracle):
    """Full update() loop, strict summation: poses, covariances and maps equal the oracle bit for bit."""
    n_streams, n_scans = 8, 120
    sc, ranges, pts, cnt, truth = make_streams(hs, n_streams, n_scans, seed0=500)
    kw = [dict(NODE), dict(free=0.4, occ=0.9, dist=0.0, ang=0.0)]
    for cfg in kw:
        procs = cpu_streams(oracle, n_streams, **cfg)
        secs, cpu_poses, cpu_cov = oracle.run_streams(procs, pts, cnt, threads=4, want_cov=True)
        with hs.Engine(n_streams=n_streams, max_points=360) as eng:
            configure(eng, **cfg)
            eng.set_strict_summation(True)
            for k in range(n_scans):
                poses, cov = eng.update_batch(pts[k], cnt[k])
                assert np.array_equal(bits(poses), bits(cpu_poses[k])), "scan %d poses differ: %s vs %s" % (k, poses, cpu_poses[k])
                assert np.array_equal(bits(cov.reshape(-1, 9)), bits(cpu_cov[k]))
            n_diff, _ = compare_maps(eng, procs, 3, exact=True)
            assert n_diff == 0
            st = eng.stats()
            assert st["integrations"] == sum(p.integrations() for p in procs)


def test_closed_loop_fast_tolerance(hs, oracle):
    """Production (fast) summation: pose within 1e-4 m / 1e-4 rad per scan, log-odds within 1e-5."""
    n_streams, n_scans = 16, 200
    sc, ranges, pts, cnt, truth = make_streams(hs, n_streams, n_scans, seed0=900)
    procs = cpu_streams(oracle, n_streams, **NODE)
    secs, cpu_poses, _ = oracle.run_streams(procs, pts, cnt, threads=8)
    with hs.Engine(n_streams=n_streams, max_points=360) as eng:
        configure(eng, **NODE)
        worst = np.zeros(3)
        for k in range(n_scans):
            poses, _ = eng.update_batch(pts[k], cnt[k], want_cov=False)
            d = np.abs(poses - cpu_poses[k])
            d[:, 2] = np.abs((d[:, 2] + np.pi) % (2 * np.pi) - np.pi)
            worst = np.maximum(worst, d.max(axis=0))
        assert worst[0] < 1e-4 and worst[1] < 1e-4 and worst[2] < 1e-4, worst
        n_diff, max_dv = compare_maps(eng, procs, 3, exact=False)
        print("fast mode: worst pose diff", worst, "cells with different marker", n_diff, "max |dlogodds|", max_dv)
        # boundary-flipped end points are reported, not hidden: allow a handful of cells per stream
        assert n_diff <= 4 * n_streams


def test_both_matcher_variants_are_bit_identical(hs, oracle, monkeypatch):
    """k_match (register cache, scans <= 768 points) and k_match2 (shared-memory cache + dense probability pass, longer
    scans) are the same arithmetic: forced onto the same 360-point streams they must agree with the oracle bit for bit."""
    n, n_scans = 6, 25
    sc, ranges, pts, cnt, truth = make_streams(hs, n, n_scans, seed0=77)
    procs = cpu_streams(oracle, n, dist=0.0, ang=0.0, free=0.4, occ=0.9)
    ref = np.stack([np.stack([p.update(pts[k, s], cnt[k, s], hint=p.pose()) for s, p in enumerate(procs)]) for k in range(n_scans)])
    for variant in ("1", "2", "3"):
        monkeypatch.setenv("HSGPU_MATCH_VARIANT", variant)
        with hs.Engine(n_streams=n, max_points=360) as eng:
            configure(eng, dist=0.0, ang=0.0, free=0.4, occ=0.9)
            for k in range(n_scans):
                out, cov = eng.update_batch(pts[k], cnt[k])
                assert np.array_equal(bits(out), bits(ref[k])), (variant, k)
            n_diff, _ = compare_maps(eng, procs, 3)
            assert n_diff == 0
    for p in procs:
        p.close()


def test_long_closed_loop_64_streams_1000_scans(hs, oracle):
    """SURVEY section 8(d) parity gate at its full length: 64 streams x 1000 scans (one full loop of the trajectory, node
    parameters: gate 0.4 m / 0.9 rad, factors 0.4 / 0.9), closed loop. Every pose of every scan and the final maps must equal
    the CPU oracle bit for bit -- far inside the 1e-4 m / 1e-4 rad / 1e-5 log-odds bar -- and the update indices must agree
    (so exactly the same scans passed the gate)."""
    import os
    n_streams, n_scans = 64, 1000
    sc, ranges, pts, cnt, truth = make_streams(hs, n_streams, n_scans, seed0=31000)
    procs = cpu_streams(oracle, n_streams, **NODE)
    secs, cpu_poses, _ = oracle.run_streams(procs, pts, cnt, threads=os.cpu_count() or 4)
    with hs.Engine(n_streams=n_streams, max_points=360) as eng:
        configure(eng, **NODE)
        for k in range(n_scans):
            poses, _ = eng.update_batch(pts[k], cnt[k], want_cov=False)
            assert np.array_equal(bits(poses), bits(cpu_poses[k])), "scan %d: max |dpose| %g" % (k, np.abs(poses - cpu_poses[k]).max())
        n_diff, _ = compare_maps(eng, procs, 3, exact=True)
        assert n_diff == 0
        for s in (0, 17, 63):
            assert eng.update_index(s, 0) == procs[s].update_index(0)
        st = eng.stats()
        assert st["integrations"] == sum(p.integrations() for p in procs)
        assert 64 * 10 < st["integrations"] < 64 * 400      # the 0.4 m / 0.9 rad gate lets only a fraction of the scans through
        # the matcher stays locked on: final poses within 15 cm of the simulated truth (relative to the start pose)
        err = np.linalg.norm(poses[:, :2] - truth[-1, :, :2], axis=1)
        assert err.max() < 0.15, err.max()
    for p in procs:
        p.close()


def test_multi_start_match_hypotheses(hs, oracle):
    """hs_match_hypotheses (BASELINE configs[3], relocalisation-style multi-start match): the full coarse-to-fine matchData of
    one scan from 512 pose hints against one stream's maps equals the reference's per-level ScanMatcher::matchData chained
    coarse to fine (MapRepMultiMap.h:116-132), bit for bit, and leaves the stream's state untouched."""
    n_scans = 60
    sc, ranges, pts, cnt, truth = make_streams(hs, 2, n_scans + 1, seed0=2718)
    cpu = cpu_streams(oracle, 1, dist=0.0, ang=0.0, free=0.4, occ=0.9)[0]
    with hs.Engine(n_streams=2, max_points=360) as eng:
        configure(eng, dist=0.0, ang=0.0, free=0.4, occ=0.9)
        pose = None
        for k in range(n_scans):
            out, _ = eng.update_batch(pts[k], cnt[k])
            pose = cpu.update(pts[k, 1], cnt[k, 1], hint=cpu.pose())
        assert np.array_equal(bits(out[1]), bits(pose))
        rng = np.random.default_rng(9)
        n_hyp = 512
        hints = (pose[None, :] + rng.uniform(-1, 1, (n_hyp, 3)) * np.array([0.3, 0.3, 0.15])).astype(np.float32)
        scan = pts[n_scans, 1, :cnt[n_scans, 1]]
        before = (eng.poses().copy(), eng.covariances().copy(), eng.download_level(1, 0).copy(), eng.update_index(1, 0))
        got, gcov = eng.match_hypotheses(1, scan, hints)
        after = (eng.poses(), eng.covariances(), eng.download_level(1, 0), eng.update_index(1, 0))
        assert np.array_equal(bits(before[0]), bits(after[0])) and np.array_equal(bits(before[1]), bits(after[1]))
        assert np.array_equal(before[2], after[2]) and before[3] == after[3]
        n_ok = 0
        for i in range(n_hyp):
            p, c = hints[i], None
            for level in (2, 1, 0):
                p, c = cpu.match_level(level, p, scan * np.float32(1.0 / (1 << level)), 5 if level == 0 else 3)
            if np.isfinite(p).all():
                assert np.array_equal(bits(got[i]), bits(p)), i
                assert np.array_equal(bits(gcov[i]), bits(c)), i
                n_ok += 1
        assert n_ok > n_hyp * 0.9
        # small batches take the matcher's ordinary gather path (no probability block planes): same bits
        few, fcov = eng.match_hypotheses(1, scan, hints[:9])
        assert np.array_equal(bits(few), bits(got[:9])) and np.array_equal(bits(fcov), bits(gcov[:9]))
        # the hypotheses near the true pose converge onto the matcher's own answer for that scan
        ref_pose = cpu.update(pts[n_scans, 1], cnt[n_scans, 1], hint=cpu.pose())
        near = np.linalg.norm(hints[:, :2] - pose[None, :2], axis=1) < 0.05
        near &= np.abs(hints[:, 2] - pose[2]) < 0.03
        assert near.sum() >= 1 and np.abs(got[near, :2] - ref_pose[None, :2]).max() < 0.02
    cpu.close()
