"""Small closed-loop + teacher-forced + off-path run for compute-sanitizer (memcheck / racecheck / initcheck):
    compute-sanitizer --tool racecheck python tests/tools/gpu_sanitizer_run.py
Covers k_match, k_match2, k_raycast (single band and banded), k_scan_to_points, k_cloud_to_points, k_residual, k_interp,
k_occupancy_ray, pack/unpack/classify."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import tfe_slam_ucl_b200 as hs
import workload

def run(variant, size, res, levels, preset, mp, n=3, scans=3):
    os.environ["HSGPU_MATCH_VARIANT"] = variant
    sc = workload.scene_preset(preset)
    ranges, _ = workload.scangen_batch(sc, 77, n, 0, scans, truth=False)
    pts, cnt = hs.scan_to_points_host(ranges, sc, 1.0 / res)
    with hs.Engine(res, size, size, (0.5, 0.5), levels, n_streams=n, max_points=mp) as eng:
        eng.set_update_factor_free(0.4); eng.set_update_factor_occupied(0.9)
        eng.set_map_update_min_dist_diff(0.0); eng.set_map_update_min_angle_diff(0.0)
        eng.set_scan_geometry(sc.n_beams, sc.angle_min, sc.angle_increment, sc.range_min, sc.range_max)
        for k in range(scans):
            poses, _ = eng.update_batch(pts[k], cnt[k]) if k % 2 == 0 else eng.update_batch_ranges(ranges[k])
        eng.raycast_batch(pts[0], cnt[0], poses)
        p = pts[0, 0, :cnt[0, 0]]
        hyp = np.array([[size / 2, size / 2, 0.1], [size / 2 + 3, size / 2 - 2, -0.2]], np.float32)
        eng.hessian_derivs_batch(0, 0, hyp, p); eng.likelihood_batch(0, 0, hyp, p); eng.covariance_for_pose(0, 0, hyp[0], p)
        eng.interp_map_value_batch(0, 0, hyp[:, :2])
        eng.occupancy_ray_batch(0, 0, np.array([[size // 2, size // 2]] * 4, np.int32), np.array([[size - 5, size // 2], [5, 9], [size // 2, 3], [size // 2, size // 2]], np.int32))
        eng.classify_level(0, levels - 1); eng.upload_level(1, 0, eng.download_level(0, 0))
        st, cp = eng.export_stream_state(0); eng.import_stream_state(2, st, cp)
        clouds = np.zeros((n, 64, 3), np.float32); clouds[:, :, 0] = 2.0
        xf = np.tile(np.concatenate([np.eye(3).reshape(-1), [0.1, 0.0, 0.2]]), (n, 1))
        eng.update_batch_clouds(clouds, np.full(n, 64, np.int32), xf, 0.16, 900.0, -1.0, 1.0)
        print("ok", variant, size, eng.stats()["kernel_launches"])

run("1", 1024, 0.05, 3, workload.SCENE_RPLIDAR_ROOM, 360)
run("2", 1024, 0.05, 3, workload.SCENE_RPLIDAR_ROOM, 360)
run("2", 2048, 0.025, 3, workload.SCENE_HIRES_HALL, 1440, n=3, scans=2)     # banded ray-cast
