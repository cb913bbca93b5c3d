"""The CPU oracle against the frozen golden fixtures (tests/golden/streams.*), which were produced by the
reference itself (oracle/_ref, unmodified hector_slam_lib headers) with tests/golden/make_golden.py.
Everything is compared bit for bit: per-scan poses and covariances, update indices, SHA-256 of every
level's {logOddsVal, updateIndex} cells, and function-level Hessian/gradient known answers."""
import hashlib
import json
import os

import numpy as np
import pytest

from hs_testutil import bits

HERE = os.path.dirname(os.path.abspath(__file__))


def load_golden():
    z = np.load(os.path.join(HERE, "golden", "streams.npz"))
    with open(os.path.join(HERE, "golden", "streams.json")) as f:
        meta = json.load(f)
    return z, meta


def replay(oracle, kind, case, z, m):
    cfg, params = m["cfg"], m["params"]
    p = oracle.CpuProcessor(kind, cfg["res"], cfg["size"], cfg["size"], (0.5, 0.5), cfg["levels"])
    p.set_update_factor_free(params["free"]); p.set_update_factor_occupied(params["occ"])
    p.set_map_update_min_dist_diff(params["dist"]); p.set_map_update_min_angle_diff(params["ang"])
    pts, cnt = z[case + "/points"], z[case + "/counts"]
    pose = np.zeros(3, np.float32)
    for k in range(m["n_scans"]):
        mwm = bool(m["mwm_every"] and k % m["mwm_every"] == m["mwm_every"] - 1)
        pose = p.update(pts[k], cnt[k], hint=pose, map_without_matching=mwm)
        assert np.array_equal(bits(pose), bits(z[case + "/poses"][k])), "%s scan %d pose" % (case, k)
        if cnt[k] > 0 and not mwm:
            assert np.array_equal(bits(p.cov().reshape(-1)), bits(z[case + "/cov"][k])), "%s scan %d cov" % (case, k)
        assert p.update_index(0) == z[case + "/update_index"][k]
    return p


@pytest.mark.parametrize("case", ["c1_node_params", "c1_gate0_mwm", "c3_small"])
@pytest.mark.parametrize("kind", ["port", "ref"])
def test_stream_golden(oracle, have_ref, kind, case):
    if kind == "ref" and not have_ref:
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    z, meta = load_golden()
    m = meta["cases"][case]
    p = replay(oracle, kind, case, z, m)
    for l in range(m["cfg"]["levels"]):
        cells = p.download_level(l)
        assert hashlib.sha256(cells.tobytes()).hexdigest() == m["level_digests_sha256"][l], "%s level %d cells" % (case, l)
        assert int((cells["logOddsVal"] != 0).sum()) == m["nonzero_cells"][l]
    # function-level known answers on the final map
    pts, cnt = z[case + "/points"], z[case + "/counts"]
    for l in range(m["cfg"]["levels"]):
        scaled = (pts[-1, :cnt[-1]] * np.float32(1.0 / (1 << l))).astype(np.float32)
        H, d = p.hessian_derivs_batch(l, z["%s/kat_l%d_poses" % (case, l)], scaled)
        assert np.array_equal(bits(H), bits(z["%s/kat_l%d_H" % (case, l)]))
        assert np.array_equal(bits(d), bits(z["%s/kat_l%d_dTr" % (case, l)]))
        # off-path members: getResidualForState / getLikelihoodForState / getCovarianceForPose / interpMapValue
        hyp = z["%s/kat_l%d_poses" % (case, l)]
        r = np.array([p.residual_for_state(l, h, scaled) for h in hyp], np.float32)
        lh = np.array([p.likelihood_for_state(l, h, scaled) for h in hyp], np.float32)
        assert np.array_equal(bits(r), bits(z["%s/kat_l%d_resid" % (case, l)])) and np.array_equal(bits(lh), bits(z["%s/kat_l%d_lh" % (case, l)]))
        for i in range(3):
            cm, cw = p.covariance_for_pose(l, hyp[i], scaled)
            assert np.array_equal(bits(cm), bits(z["%s/kat_l%d_covmap" % (case, l)][i]))
            assert np.array_equal(bits(cw), bits(z["%s/kat_l%d_covworld" % (case, l)][i]))
        iv = np.array([p.interp_map_value(l, c[0], c[1]) for c in z["%s/kat_l%d_coords" % (case, l)]], np.float32)
        assert np.array_equal(bits(iv), bits(z["%s/kat_l%d_interp" % (case, l)]))
    # occupancy ray queries (hector_map_tools) on the classified level-0 map
    size = m["cfg"]["size"]
    v0 = p.download_level(0)["logOddsVal"].reshape(size, size)
    occ = np.where(v0 < 0, 0, np.where(v0 > 0, 100, -1)).astype(np.int8)
    for e, d_ref, h_ref in zip(z[case + "/ray_ends"], z[case + "/ray_dist"], z[case + "/ray_hit"]):
        d_got, h_got = oracle.check_occupancy_bresenhami(kind, occ, z[case + "/ray_begin"], e)
        assert d_got == d_ref and (h_got is None or np.array_equal(h_got, h_ref))
    assert (z[case + "/ray_dist"] >= 0).sum() >= 3, "golden rays should hit something"
    p.close()


@pytest.mark.gpu
@pytest.mark.parametrize("case", ["c1_node_params", "c1_gate0_mwm", "c3_small"])
def test_stream_golden_gpu(hs, case):
    """The CUDA path (default reference-order sums) against the reference's own golden outputs, bit for bit."""
    z, meta = load_golden()
    m = meta["cases"][case]
    cfg, params = m["cfg"], m["params"]
    pts, cnt = z[case + "/points"], z[case + "/counts"]
    with hs.Engine(cfg["res"], cfg["size"], cfg["size"], (0.5, 0.5), cfg["levels"], n_streams=1, max_points=pts.shape[1]) as eng:
        eng.set_update_factor_free(params["free"]); eng.set_update_factor_occupied(params["occ"])
        eng.set_map_update_min_dist_diff(params["dist"]); eng.set_map_update_min_angle_diff(params["ang"])
        pose = np.zeros((1, 3), np.float32)
        for k in range(m["n_scans"]):
            mwm = bool(m["mwm_every"] and k % m["mwm_every"] == m["mwm_every"] - 1)
            pose, cov = eng.update_batch(pts[k:k + 1], cnt[k:k + 1], pose_hints=pose, map_without_matching=[mwm])
            assert np.array_equal(bits(pose[0]), bits(z[case + "/poses"][k])), "%s scan %d pose %s vs %s" % (case, k, pose[0], z[case + "/poses"][k])
            if cnt[k] > 0 and not mwm:
                assert np.array_equal(bits(cov.reshape(-1)), bits(z[case + "/cov"][k]))
            assert eng.update_index(0, 0) == z[case + "/update_index"][k]
        for l in range(cfg["levels"]):
            cells = eng.download_level(0, l)
            assert hashlib.sha256(cells.tobytes()).hexdigest() == m["level_digests_sha256"][l], "%s level %d cells" % (case, l)
        for l in range(cfg["levels"]):
            scaled = (pts[-1, :cnt[-1]] * np.float32(1.0 / (1 << l))).astype(np.float32)
            H, d = eng.hessian_derivs_batch(0, l, z["%s/kat_l%d_poses" % (case, l)], scaled)
            assert np.array_equal(bits(H), bits(z["%s/kat_l%d_H" % (case, l)]))
            assert np.array_equal(bits(d), bits(z["%s/kat_l%d_dTr" % (case, l)]))
            # off-path OccGridMapUtil members and the hector_map_tools ray query, against the reference's golden outputs
            hyp = z["%s/kat_l%d_poses" % (case, l)]
            r, lh = eng.likelihood_batch(0, l, hyp, scaled)
            assert np.array_equal(bits(r), bits(z["%s/kat_l%d_resid" % (case, l)])) and np.array_equal(bits(lh), bits(z["%s/kat_l%d_lh" % (case, l)]))
            for i in range(3):
                cm, cw = eng.covariance_for_pose(0, l, hyp[i], scaled)
                assert np.array_equal(bits(cm), bits(z["%s/kat_l%d_covmap" % (case, l)][i]))
                assert np.array_equal(bits(cw), bits(z["%s/kat_l%d_covworld" % (case, l)][i]))
            coords = z["%s/kat_l%d_coords" % (case, l)]
            iv, dv = eng.interp_map_value_batch(0, l, coords)
            assert np.array_equal(bits(iv), bits(z["%s/kat_l%d_interp" % (case, l)]))
            assert np.array_equal(bits(dv[:, 0]), bits(iv))   # interpMapValueWithDerivatives' value equals interpMapValue
        ends = z[case + "/ray_ends"]
        dist, hit = eng.occupancy_ray_batch(0, 0, np.repeat(z[case + "/ray_begin"][None, :], len(ends), 0), ends)
        assert np.array_equal(dist, z[case + "/ray_dist"]) and np.array_equal(hit, z[case + "/ray_hit"])
