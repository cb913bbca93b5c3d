"""Generates the frozen golden fixtures under tests/golden/ from the REFERENCE ITSELF
(oracle/_ref/libhsref.so = the unmodified hector_slam_lib headers under /root/reference compiled with the
mini-Eigen shim, see oracle/Makefile). Run here (where /root/reference exists):

    python tests/golden/make_golden.py

The reference ships no tests or golden vectors for this path, so these files are what
pins the C restatement (oracle/hs_oracle.c) and, through it, the CUDA path on machines without the reference.
Inputs are stored next to the outputs so that nothing depends on regenerating the synthetic scans.
"""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

import workload  # noqa: E402  (host-side scan generator + conversion only: no CUDA library is loaded)
from oracle import pyoracle  # noqa: E402


def cell_digest(cells):
    return hashlib.sha256(np.ascontiguousarray(cells).tobytes()).hexdigest()


def run_stream(kind, cfg, pts, cnt, params, mwm_every=0):
    p = pyoracle.CpuProcessor(kind, cfg["res"], cfg["size"], cfg["size"], (0.5, 0.5), cfg["levels"])
    p.set_update_factor_free(params["free"]); p.set_update_factor_occupied(params["occ"])
    p.set_map_update_min_dist_diff(params["dist"]); p.set_map_update_min_angle_diff(params["ang"])
    n = pts.shape[0]
    poses = np.zeros((n, 3), np.float32); covs = np.zeros((n, 9), np.float32); upd = np.zeros(n, np.int32)
    pose = np.zeros(3, np.float32)
    for k in range(n):
        mwm = bool(mwm_every and k % mwm_every == mwm_every - 1)
        pose = p.update(pts[k], cnt[k], hint=pose, map_without_matching=mwm)
        poses[k] = pose; covs[k] = p.cov().reshape(-1); upd[k] = p.update_index(0)
    digests = [cell_digest(p.download_level(l)) for l in range(cfg["levels"])]
    return p, poses, covs, upd, digests


def main():
    if not pyoracle.have("ref"):
        pyoracle.build(ref=True)
    assert pyoracle.have("ref"), "oracle/_ref/libhsref.so is required (needs /root/reference)"
    meta = {"generator": "tests/golden/make_golden.py", "source": "oracle/_ref/libhsref.so (unmodified reference headers + mini-Eigen shim)", "cases": {}}
    arrays = {}

    cases = {
        # BASELINE configs[0]/[1] geometry: RPLidar-style 360-beam scans, 3 levels 1024^2 @0.05 m, ROS-node parameters
        "c1_node_params": dict(preset=workload.SCENE_RPLIDAR_ROOM, seed=0xC0FFEE, n_scans=40, cfg=dict(res=0.05, size=1024, levels=3),
                               params=dict(free=0.4, occ=0.9, dist=0.4, ang=0.9), mwm_every=0),
        # every scan integrated (gate 0/0), library-default factors 0.4/0.6, with map_without_matching every 7th scan
        "c1_gate0_mwm": dict(preset=workload.SCENE_RPLIDAR_ROOM, seed=0xC0FFEE + 1, n_scans=30, cfg=dict(res=0.05, size=1024, levels=3),
                             params=dict(free=0.4, occ=0.6, dist=0.0, ang=0.0), mwm_every=7),
        # BASELINE configs[2] shape, reduced: 1440-beam 30 m scans, 4 levels 2048^2 @0.025 m
        "c3_small": dict(preset=workload.SCENE_HIRES_HALL, seed=0xC0FFEE + 2, n_scans=8, cfg=dict(res=0.025, size=2048, levels=4),
                         params=dict(free=0.4, occ=0.9, dist=0.0, ang=0.0), mwm_every=0),
    }
    for name, c in cases.items():
        sc = workload.scene_preset(c["preset"])
        ranges, _ = workload.scangen_batch(sc, c["seed"], 1, 0, c["n_scans"], truth=False)
        ranges = ranges[:, 0]
        if name == "c1_gate0_mwm":   # ragged input: knock out ranges so that counts vary, plus one empty scan
            rng = np.random.default_rng(5)
            mask = rng.random(ranges.shape) < 0.15
            ranges = ranges.copy(); ranges[mask] = np.inf
            ranges[11, :] = 0.05     # below range_min: DataContainer stays empty
        pts, cnt = workload.scan_to_points_host(ranges, sc, 1.0 / c["cfg"]["res"])
        p, poses, covs, upd, digests = run_stream("ref", c["cfg"], pts, cnt, c["params"], c["mwm_every"])
        arrays[name + "/points"] = pts; arrays[name + "/counts"] = cnt
        arrays[name + "/poses"] = poses; arrays[name + "/cov"] = covs; arrays[name + "/update_index"] = upd
        # function-level KATs on the final map
        rng = np.random.default_rng(11)
        for level in range(c["cfg"]["levels"]):
            base = p.world_to_map(level, poses[-1])
            hyp = (base[None, :] + rng.uniform(-1, 1, (12, 3)) * np.array([6.0, 6.0, 0.3]) / (1 << level)).astype(np.float32)
            scaled = (pts[-1, :cnt[-1]] * np.float32(1.0 / (1 << level))).astype(np.float32)
            H, d = p.hessian_derivs_batch(level, hyp, scaled)
            arrays["%s/kat_l%d_poses" % (name, level)] = hyp
            arrays["%s/kat_l%d_H" % (name, level)] = H
            arrays["%s/kat_l%d_dTr" % (name, level)] = d
            # off-path OccGridMapUtil members (OccGridMapUtil.h:106-285) on the same hypotheses
            arrays["%s/kat_l%d_resid" % (name, level)] = np.array([p.residual_for_state(level, h, scaled) for h in hyp], np.float32)
            arrays["%s/kat_l%d_lh" % (name, level)] = np.array([p.likelihood_for_state(level, h, scaled) for h in hyp], np.float32)
            cc = [p.covariance_for_pose(level, h, scaled) for h in hyp[:3]]
            arrays["%s/kat_l%d_covmap" % (name, level)] = np.stack([c[0] for c in cc])
            arrays["%s/kat_l%d_covworld" % (name, level)] = np.stack([c[1] for c in cc])
            coords = (hyp[:, :2] + rng.uniform(-40, 40, (12, 2)) / (1 << level)).astype(np.float32)
            arrays["%s/kat_l%d_coords" % (name, level)] = coords
            arrays["%s/kat_l%d_interp" % (name, level)] = np.array([p.interp_map_value(level, c[0], c[1]) for c in coords], np.float32)
        # hector_map_tools occupancy ray queries on the published (classified) level-0 map
        size = c["cfg"]["size"]
        v0 = p.download_level(0)["logOddsVal"].reshape(size, size)
        occ = np.where(v0 < 0, 0, np.where(v0 > 0, 100, -1)).astype(np.int8)
        b0 = p.world_to_map(0, poses[-1])[:2].astype(np.int32)
        ends = (b0[None, :] + rng.integers(-size // 3, size // 3, (24, 2))).astype(np.int32)
        ends[0] = b0                      # begin == end
        ends[1] = (size + 5, b0[1])       # end outside the map
        rays = [pyoracle.check_occupancy_bresenhami("ref", occ, b0, e) for e in ends]
        arrays[name + "/ray_begin"] = b0
        arrays[name + "/ray_ends"] = ends
        arrays[name + "/ray_dist"] = np.array([r[0] for r in rays], np.float32)
        arrays[name + "/ray_hit"] = np.array([r[1] if r[1] is not None else (-1, -1) for r in rays], np.int32)
        nz = [int((p.download_level(l)["logOddsVal"] != 0).sum()) for l in range(c["cfg"]["levels"])]
        meta["cases"][name] = {"cfg": c["cfg"], "params": c["params"], "mwm_every": c["mwm_every"], "n_scans": c["n_scans"],
                               "level_digests_sha256": digests, "nonzero_cells": nz}
        print(name, "final pose", poses[-1], "nonzero cells", nz, "counts", cnt.min(), cnt.max())
        p.close()
    np.savez_compressed(os.path.join(HERE, "streams.npz"), **arrays)
    with open(os.path.join(HERE, "streams.json"), "w") as f:
        json.dump(meta, f, indent=1)
    print("wrote", os.path.join(HERE, "streams.npz"), os.path.getsize(os.path.join(HERE, "streams.npz")) // 1024, "KiB")


if __name__ == "__main__":
    main()
