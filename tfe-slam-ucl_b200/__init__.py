"""hsgpu -- B200-native batched scan-to-map engine behind the hector_slam_lib API.

This Python module is only a ctypes binding over the C ABI in ``include/hsgpu.h`` (the product is the
CUDA engine in ``csrc/`` built into ``libhsgpu.so``); it exists for the test-suite and ``bench.py``.
There is no Python or CPU implementation of the hot path here: if ``libhsgpu.so`` is missing the import
of the library fails loudly, and without a CUDA device ``Engine(...)`` raises ``HsError``.

Reference interface mirrored: ``hectorslam::HectorSlamProcessor``
(src/hector_slam/hector_mapping/include/hector_slam_lib/slam_main/HectorSlamProcessor.h:50-154).
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libhsgpu.so")
CSRC = os.path.join(_HERE, "csrc")

HS_OK = 0
HS_ERR_INVALID_ARG = -1
HS_ERR_NO_DEVICE = -2
HS_ERR_OUT_OF_MEMORY = -3
HS_ERR_NOT_CONFIGURED = -4
HS_ALL_STREAMS = -1

CELL_DTYPE = np.dtype([("logOddsVal", np.float32), ("updateIndex", np.int32)])


class HsError(RuntimeError):
    def __init__(self, status, what):
        self.status = status
        super().__init__("%s failed: %s (status %d)" % (what, status_string(status), status))


class HsConfig(ctypes.Structure):
    _fields_ = [("map_resolution", ctypes.c_float), ("map_size_x", ctypes.c_int32), ("map_size_y", ctypes.c_int32),
                ("start_x", ctypes.c_float), ("start_y", ctypes.c_float), ("levels", ctypes.c_int32),
                ("n_streams", ctypes.c_int32), ("max_points", ctypes.c_int32), ("device", ctypes.c_int32)]


class HsStats(ctypes.Structure):
    _fields_ = [("updates", ctypes.c_uint64), ("integrations", ctypes.c_uint64), ("angle_clamps", ctypes.c_uint64),
                ("kernel_launches", ctypes.c_uint64), ("h2d_bytes", ctypes.c_uint64), ("d2h_bytes", ctypes.c_uint64),
                ("cell_visits", ctypes.c_uint64)]


class HsMapInfo(ctypes.Structure):
    _fields_ = [("origin_x", ctypes.c_float), ("origin_y", ctypes.c_float), ("resolution", ctypes.c_float),
                ("width", ctypes.c_int32), ("height", ctypes.c_int32)]


class HsStreamState(ctypes.Structure):
    _fields_ = [("last_scan_match_pose", ctypes.c_float * 3), ("last_map_update_pose", ctypes.c_float * 3),
                ("last_scan_match_cov", ctypes.c_float * 9), ("curr_update_index", ctypes.c_int32),
                ("last_update_index", ctypes.c_int32), ("coarse_count", ctypes.c_int32), ("coarse_origo", ctypes.c_float * 2)]


class HsKernelTimes(ctypes.Structure):
    _fields_ = [("ms", ctypes.c_double * 8), ("launches", ctypes.c_uint64 * 8)]


KERNEL_NAMES = ("k_scan_to_points", "k_match", "k_raycast", "k_hessian_derivs", "k_build_pblk")


def build(verbose=False):
    """Compile libhsgpu.so for sm_100a (nvcc cross-compiles without a GPU)."""
    res = subprocess.run(["make", "-C", CSRC], capture_output=True, text=True)
    if verbose or res.returncode != 0:
        print(res.stdout)
        print(res.stderr)
    if res.returncode != 0:
        raise RuntimeError("building libhsgpu.so failed")
    return LIB_PATH


_lib = None
_P = ctypes.c_void_p
_I = ctypes.c_int
_F = ctypes.c_float


def lib():
    """The loaded C-ABI library. Raises if it has not been built: there is no fallback path."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError("libhsgpu.so is missing (%s): run `python -c 'import __graft_entry__ as g; g.build()'` "
                          "or `make -C tfe-slam-ucl_b200/csrc`; hsgpu has no CPU fallback" % LIB_PATH)
    L = ctypes.CDLL(LIB_PATH)
    sig = {
        "hs_create": (_I, [_P, _P]), "hs_destroy": (_I, [_P]), "hs_reset": (_I, [_P, _I]),
        "hs_set_update_factor_free": (_I, [_P, _F]), "hs_set_update_factor_occupied": (_I, [_P, _F]),
        "hs_set_map_update_min_dist_diff": (_I, [_P, _F]), "hs_set_map_update_min_angle_diff": (_I, [_P, _F]),
        "hs_set_strict_summation": (_I, [_P, _I]),
        "hs_get_scale_to_map": (_F, [_P]), "hs_get_map_levels": (_I, [_P]), "hs_get_n_streams": (_I, [_P]),
        "hs_get_max_points": (_I, [_P]), "hs_get_level_dims": (_I, [_P, _I, _P, _P, _P]),
        "hs_set_cuda_stream": (_I, [_P, _P]), "hs_synchronize": (_I, [_P]), "hs_status_string": (ctypes.c_char_p, [_I]),
        "hs_update_batch": (_I, [_P, _I, _P, _P, _P, _P, _P, _P, _P, _P]),
        "hs_update_batch_device": (_I, [_P, _I, _P, _P, _P, _P, _P, _P]),
        "hs_submit_batch": (_I, [_P, _I, _P, _P, _P, _P, _P, _P, _P]), "hs_wait": (_I, [_P]),
        "hs_set_scan_geometry": (_I, [_P, _I, _F, _F, _F, _F]),
        "hs_update_batch_ranges": (_I, [_P, _I, _P, _P, _P, _P, _P, _P]),
        "hs_update_batch_ranges_device": (_I, [_P, _I, _P, _P, _P, _P]),
        "hs_scan_to_points_host": (_I, [_P, _I, _F, _F, _F, _F, _F, _P]),
        "hs_scan_to_points_host_batch": (_I, [_P, _I, _I, _F, _F, _F, _F, _F, _I, _P, _P, _I]),
        "hs_get_poses": (_I, [_P, _I, _I, _P]), "hs_get_covariances": (_I, [_P, _I, _I, _P]),
        "hs_get_last_map_update_poses": (_I, [_P, _I, _I, _P]), "hs_device_pose_ptr": (_P, [_P]),
        "hs_hessian_derivs_batch": (_I, [_P, _I, _I, _P, _I, _P, _I, _P, _P]),
        "hs_hessian_derivs_batch_device": (_I, [_P, _I, _I, _P, _I, _P, _I, _P, _P]),
        "hs_match_batch": (_I, [_P, _I, _P, _P, _P, _P, _P, _P, _P]),
        "hs_match_hypotheses": (_I, [_P, _I, _P, _I, _P, _I, _P, _P]),
        "hs_raycast_batch": (_I, [_P, _I, _P, _P, _P, _P, _P]),
        "hs_integrate_batch": (_I, [_P, _I, _P, _P, _P, _P, _P]),
        "hs_download_level": (_I, [_P, _I, _I, _P]), "hs_upload_level": (_I, [_P, _I, _I, _P]),
        "hs_get_update_index": (_I, [_P, _I, _I, _P]), "hs_classify_level": (_I, [_P, _I, _I, _P]),
        "hs_get_stats": (_I, [_P, _P]), "hs_get_map_info": (_I, [_P, _I, _P]),
        "hs_update_batch_clouds": (_I, [_P, _I, _P, _P, _P, _I, _P, _F, _F, _F, _F, _P, _P, _P, _P]),
        "hs_cloud_to_points_host": (_I, [_P, _I, _P, _F, _F, _F, _F, _F, _I, _P, _P]),
        "hs_likelihood_batch": (_I, [_P, _I, _I, _P, _I, _P, _I, _P, _P]),
        "hs_covariance_for_pose": (_I, [_P, _I, _I, _P, _P, _I, _P, _P]),
        "hs_interp_map_value_batch": (_I, [_P, _I, _I, _P, _I, _P, _P]),
        "hs_occupancy_ray_batch": (_I, [_P, _I, _I, _P, _P, _I, _P, _P]),
        "hs_export_stream_state": (_I, [_P, _I, _P, _P]), "hs_import_stream_state": (_I, [_P, _I, _P, _P]),
        "hs_device_count": (_I, []), "hs_get_map_epoch": (ctypes.c_uint64, [_P]), "hs_copy_stream_maps": (_I, [_P, _I, _P, _I]),
        "hs_profile_enable": (_I, [_P, _I]), "hs_profile_collect": (_I, [_P, _P]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)
        fn.restype = res
        fn.argtypes = args
    _lib = L
    return L


def status_string(status):
    return lib().hs_status_string(int(status)).decode()


def _check(status, what):
    if status != HS_OK:
        raise HsError(status, what)


def _arr(a, dtype, shape=None):
    if a is None:
        return None
    a = np.ascontiguousarray(a, dtype=dtype)
    if shape is not None:
        a = a.reshape(shape)
    return a


def _ptr(a):
    return None if a is None else a.ctypes.data


# ---- host twins of the ingestion kernels --------------------------------------------------------------

def scan_to_points_host(ranges, scene, scale_to_map, max_points=None):
    """rosLaserScanToDataContainer over a [..., n_beams] array -> (points [..., max_points, 2], counts [...]).
    `scene`: anything with angle_min / angle_increment / range_min / range_max (the LaserScan header fields)."""
    ranges = np.ascontiguousarray(ranges, np.float32)
    lead = ranges.shape[:-1]
    nb = ranges.shape[-1]
    mp = max_points or nb
    flat = ranges.reshape(-1, nb)
    pts = np.empty((flat.shape[0], mp, 2), np.float32)
    cnt = np.empty(flat.shape[0], np.int32)
    _check(lib().hs_scan_to_points_host_batch(flat.ctypes.data, flat.shape[0], nb, scene.angle_min, scene.angle_increment,
                                              scene.range_min, scene.range_max, scale_to_map, mp, pts.ctypes.data,
                                              cnt.ctypes.data, os.cpu_count() or 1), "hs_scan_to_points_host_batch")
    return pts.reshape(lead + (mp, 2)), cnt.reshape(lead)


def cloud_to_points_host(cloud_xyz, transform12, scale_to_map, sqr_min_dist, sqr_max_dist, z_min, z_max, max_points=None):
    """rosPointCloudToDataContainer for one cloud on the host -> (points [k][2], origo [2])."""
    cloud_xyz = np.ascontiguousarray(cloud_xyz, np.float32).reshape(-1, 3)
    t = np.ascontiguousarray(transform12, np.float64).reshape(12)
    mp = max_points or max(cloud_xyz.shape[0], 1)
    pts = np.zeros((mp, 2), np.float32)
    origo = np.zeros(2, np.float32)
    k = lib().hs_cloud_to_points_host(cloud_xyz.ctypes.data, cloud_xyz.shape[0], t.ctypes.data, scale_to_map, sqr_min_dist,
                                      sqr_max_dist, z_min, z_max, mp, pts.ctypes.data, origo.ctypes.data)
    if k < 0:
        raise HsError(k, "hs_cloud_to_points_host")
    return pts[:k].copy(), origo


# ---- the engine ------------------------------------------------------------------------------------

class Engine:
    """n_streams independent HectorSlamProcessor states on one GPU."""

    def __init__(self, map_resolution=0.05, map_size_x=1024, map_size_y=1024, start=(0.5, 0.5), levels=3, n_streams=1,
                 max_points=360, device=0):
        self._h = ctypes.c_void_p()
        self.cfg = HsConfig(map_resolution, map_size_x, map_size_y, start[0], start[1], levels, n_streams, max_points, device)
        _check(lib().hs_create(ctypes.byref(self.cfg), ctypes.byref(self._h)), "hs_create")
        self.n_streams = n_streams
        self.max_points = max_points
        self.levels = levels

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            lib().hs_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # configuration (HectorSlamProcessor.h:136-139)
    def set_update_factor_free(self, f): _check(lib().hs_set_update_factor_free(self._h, f), "hs_set_update_factor_free")
    def set_update_factor_occupied(self, f): _check(lib().hs_set_update_factor_occupied(self._h, f), "hs_set_update_factor_occupied")
    def set_map_update_min_dist_diff(self, d): _check(lib().hs_set_map_update_min_dist_diff(self._h, d), "hs_set_map_update_min_dist_diff")
    def set_map_update_min_angle_diff(self, a): _check(lib().hs_set_map_update_min_angle_diff(self._h, a), "hs_set_map_update_min_angle_diff")
    def set_strict_summation(self, on): _check(lib().hs_set_strict_summation(self._h, int(on)), "hs_set_strict_summation")
    def set_cuda_stream(self, handle): _check(lib().hs_set_cuda_stream(self._h, handle), "hs_set_cuda_stream")
    def synchronize(self): _check(lib().hs_synchronize(self._h), "hs_synchronize")
    def reset(self, stream=HS_ALL_STREAMS): _check(lib().hs_reset(self._h, stream), "hs_reset")
    def scale_to_map(self): return lib().hs_get_scale_to_map(self._h)
    def map_levels(self): return lib().hs_get_map_levels(self._h)

    def level_dims(self, level):
        sx, sy, cl = ctypes.c_int(), ctypes.c_int(), ctypes.c_float()
        _check(lib().hs_get_level_dims(self._h, level, ctypes.byref(sx), ctypes.byref(sy), ctypes.byref(cl)), "hs_get_level_dims")
        return sx.value, sy.value, cl.value

    def set_scan_geometry(self, n_beams, angle_min, angle_increment, range_min, range_max):
        _check(lib().hs_set_scan_geometry(self._h, n_beams, angle_min, angle_increment, range_min, range_max), "hs_set_scan_geometry")

    # the hot path
    def update_batch(self, points, counts, origos=None, pose_hints=None, map_without_matching=None, stream_ids=None,
                     want_cov=True):
        counts = _arr(counts, np.int32)
        n = counts.shape[0]
        points = _arr(points, np.float32, (n, self.max_points, 2))
        origos = _arr(origos, np.float32, (n, 2))
        pose_hints = _arr(pose_hints, np.float32, (n, 3))
        flags = _arr(map_without_matching, np.uint8, (n,))
        ids = _arr(stream_ids, np.int32, (n,))
        poses = np.empty((n, 3), np.float32)
        cov = np.empty((n, 9), np.float32) if want_cov else None
        _check(lib().hs_update_batch(self._h, n, _ptr(ids), points.ctypes.data, counts.ctypes.data, _ptr(origos),
                                     _ptr(pose_hints), _ptr(flags), poses.ctypes.data, _ptr(cov)), "hs_update_batch")
        return poses, (cov.reshape(n, 3, 3) if want_cov else None)

    def update_batch_raw(self, n, points_ptr, counts_ptr, poses_out_ptr, origos_ptr=None, hints_ptr=None, flags_ptr=None,
                         ids_ptr=None, cov_out_ptr=None):
        """hs_update_batch with raw host addresses (pinned buffers in bench.py)."""
        _check(lib().hs_update_batch(self._h, n, ids_ptr, points_ptr, counts_ptr, origos_ptr, hints_ptr, flags_ptr,
                                     poses_out_ptr, cov_out_ptr), "hs_update_batch")

    def submit_batch_raw(self, n, points_ptr, counts_ptr, poses_out_ptr=None, origos_ptr=None, hints_ptr=None, flags_ptr=None, ids_ptr=None):
        """hs_submit_batch with raw PINNED host addresses: returns at once, results are in place after wait()."""
        _check(lib().hs_submit_batch(self._h, n, ids_ptr, points_ptr, counts_ptr, origos_ptr, hints_ptr, flags_ptr, poses_out_ptr),
               "hs_submit_batch")

    def wait(self): _check(lib().hs_wait(self._h), "hs_wait")

    def update_batch_device(self, n, points_dptr, counts_dptr, origos_dptr=None, hints_dptr=None, flags_dptr=None, ids_dptr=None):
        _check(lib().hs_update_batch_device(self._h, n, ids_dptr, points_dptr, counts_dptr, origos_dptr, hints_dptr, flags_dptr),
               "hs_update_batch_device")

    def update_batch_ranges(self, ranges, pose_hints=None, map_without_matching=None, stream_ids=None, want_cov=True):
        ranges = np.ascontiguousarray(ranges, np.float32)
        n = ranges.shape[0]
        pose_hints = _arr(pose_hints, np.float32, (n, 3))
        flags = _arr(map_without_matching, np.uint8, (n,))
        ids = _arr(stream_ids, np.int32, (n,))
        poses = np.empty((n, 3), np.float32)
        cov = np.empty((n, 9), np.float32) if want_cov else None
        _check(lib().hs_update_batch_ranges(self._h, n, _ptr(ids), ranges.ctypes.data, _ptr(pose_hints), _ptr(flags),
                                            poses.ctypes.data, _ptr(cov)), "hs_update_batch_ranges")
        return poses, (cov.reshape(n, 3, 3) if want_cov else None)

    def update_batch_clouds(self, clouds_xyz, cloud_counts, transforms12, sqr_min_dist, sqr_max_dist, z_min, z_max,
                            pose_hints=None, map_without_matching=None, stream_ids=None, want_cov=True):
        """rosPointCloudToDataContainer on the GPU + update(): clouds_xyz [n][max_cloud][3], transforms12 [n][12] (double)."""
        cloud_counts = _arr(cloud_counts, np.int32)
        n = cloud_counts.shape[0]
        clouds_xyz = np.ascontiguousarray(clouds_xyz, np.float32).reshape(n, -1, 3)
        transforms12 = np.ascontiguousarray(transforms12, np.float64).reshape(n, 12)
        pose_hints = _arr(pose_hints, np.float32, (n, 3))
        flags = _arr(map_without_matching, np.uint8, (n,))
        ids = _arr(stream_ids, np.int32, (n,))
        poses = np.empty((n, 3), np.float32)
        cov = np.empty((n, 9), np.float32) if want_cov else None
        _check(lib().hs_update_batch_clouds(self._h, n, _ptr(ids), clouds_xyz.ctypes.data, cloud_counts.ctypes.data, clouds_xyz.shape[1],
                                            transforms12.ctypes.data, sqr_min_dist, sqr_max_dist, z_min, z_max, _ptr(pose_hints),
                                            _ptr(flags), poses.ctypes.data, _ptr(cov)), "hs_update_batch_clouds")
        return poses, (cov.reshape(n, 3, 3) if want_cov else None)

    def update_batch_ranges_raw(self, n, ranges_ptr, poses_out_ptr, hints_ptr=None, flags_ptr=None, ids_ptr=None, cov_out_ptr=None):
        _check(lib().hs_update_batch_ranges(self._h, n, ids_ptr, ranges_ptr, hints_ptr, flags_ptr, poses_out_ptr, cov_out_ptr),
               "hs_update_batch_ranges")

    def update_batch_ranges_device(self, n, ranges_dptr, hints_dptr=None, flags_dptr=None, ids_dptr=None):
        _check(lib().hs_update_batch_ranges_device(self._h, n, ids_dptr, ranges_dptr, hints_dptr, flags_dptr),
               "hs_update_batch_ranges_device")

    def match_batch(self, points, counts, origos=None, pose_hints=None, stream_ids=None):
        counts = _arr(counts, np.int32)
        n = counts.shape[0]
        points = _arr(points, np.float32, (n, self.max_points, 2))
        origos = _arr(origos, np.float32, (n, 2))
        pose_hints = _arr(pose_hints, np.float32, (n, 3))
        ids = _arr(stream_ids, np.int32, (n,))
        poses = np.empty((n, 3), np.float32)
        cov = np.empty((n, 9), np.float32)
        _check(lib().hs_match_batch(self._h, n, _ptr(ids), points.ctypes.data, counts.ctypes.data, _ptr(origos),
                                    _ptr(pose_hints), poses.ctypes.data, cov.ctypes.data), "hs_match_batch")
        return poses, cov.reshape(n, 3, 3)

    def match_hypotheses(self, stream, points, pose_hints):
        """Multi-start matchData of one scan against one stream's maps: (poses [n][3], cov [n][3][3]); no state is touched."""
        points = _arr(points, np.float32).reshape(-1, 2)
        pose_hints = _arr(pose_hints, np.float32).reshape(-1, 3)
        n = pose_hints.shape[0]
        poses = np.empty((n, 3), np.float32)
        cov = np.empty((n, 9), np.float32)
        _check(lib().hs_match_hypotheses(self._h, stream, points.ctypes.data, points.shape[0], pose_hints.ctypes.data, n,
                                         poses.ctypes.data, cov.ctypes.data), "hs_match_hypotheses")
        return poses, cov.reshape(n, 3, 3)

    def raycast_batch(self, points, counts, poses_world, origos=None, stream_ids=None, refresh_coarse=True):
        """refresh_coarse=True: hs_raycast_batch (teacher forcing); False: hs_integrate_batch (MapRepMultiMap::updateByScan
        proper: the coarse levels integrate the containers the last matchData left behind)."""
        counts = _arr(counts, np.int32)
        n = counts.shape[0]
        points = _arr(points, np.float32, (n, self.max_points, 2))
        origos = _arr(origos, np.float32, (n, 2))
        poses_world = _arr(poses_world, np.float32, (n, 3))
        ids = _arr(stream_ids, np.int32, (n,))
        fn, name = (lib().hs_raycast_batch, "hs_raycast_batch") if refresh_coarse else (lib().hs_integrate_batch, "hs_integrate_batch")
        _check(fn(self._h, n, _ptr(ids), points.ctypes.data, counts.ctypes.data, _ptr(origos), poses_world.ctypes.data), name)

    def hessian_derivs_batch(self, stream, level, poses_map, points):
        poses_map = _arr(poses_map, np.float32).reshape(-1, 3)
        points = _arr(points, np.float32).reshape(-1, 2)
        n = poses_map.shape[0]
        H = np.empty((n, 9), np.float32)
        d = np.empty((n, 3), np.float32)
        _check(lib().hs_hessian_derivs_batch(self._h, stream, level, poses_map.ctypes.data, n, points.ctypes.data,
                                             points.shape[0], H.ctypes.data, d.ctypes.data), "hs_hessian_derivs_batch")
        return H.reshape(n, 3, 3), d

    def hessian_derivs_batch_device(self, stream, level, poses_dptr, n_poses, points_dptr, n_points, H9_dptr, dTr3_dptr):
        """Device pointers in and out, asynchronous on the engine's stream."""
        _check(lib().hs_hessian_derivs_batch_device(self._h, stream, level, poses_dptr, n_poses, points_dptr, n_points, H9_dptr, dTr3_dptr),
               "hs_hessian_derivs_batch_device")

    # off-path OccGridMapUtil members (OccGridMapUtil.h:106-285) and hector_map_tools ray queries
    def likelihood_batch(self, stream, level, poses_map, points):
        """(getResidualForState, getLikelihoodForState) per pose hypothesis."""
        poses_map = _arr(poses_map, np.float32).reshape(-1, 3)
        points = _arr(points, np.float32).reshape(-1, 2)
        n = poses_map.shape[0]
        r = np.empty(n, np.float32)
        lh = np.empty(n, np.float32)
        _check(lib().hs_likelihood_batch(self._h, stream, level, poses_map.ctypes.data, n, points.ctypes.data, points.shape[0],
                                         r.ctypes.data, lh.ctypes.data), "hs_likelihood_batch")
        return r, lh

    def covariance_for_pose(self, stream, level, pose_map, points):
        pose_map = _arr(pose_map, np.float32).reshape(3)
        points = _arr(points, np.float32).reshape(-1, 2)
        cm = np.empty(9, np.float32)
        cw = np.empty(9, np.float32)
        _check(lib().hs_covariance_for_pose(self._h, stream, level, pose_map.ctypes.data, points.ctypes.data, points.shape[0],
                                            cm.ctypes.data, cw.ctypes.data), "hs_covariance_for_pose")
        return cm.reshape(3, 3), cw.reshape(3, 3)

    def interp_map_value_batch(self, stream, level, coords_map):
        """(interpMapValue [n], interpMapValueWithDerivatives [n][3])."""
        coords_map = _arr(coords_map, np.float32).reshape(-1, 2)
        n = coords_map.shape[0]
        v = np.empty(n, np.float32)
        d = np.empty((n, 3), np.float32)
        _check(lib().hs_interp_map_value_batch(self._h, stream, level, coords_map.ctypes.data, n, v.ctypes.data, d.ctypes.data),
               "hs_interp_map_value_batch")
        return v, d

    def occupancy_ray_batch(self, stream, level, begin_cells, end_cells):
        """checkOccupancyBresenhami per ray: (distance in cells or -1 [n], hit cell or (-1,-1) [n][2])."""
        b = _arr(begin_cells, np.int32).reshape(-1, 2)
        en = _arr(end_cells, np.int32).reshape(-1, 2)
        n = b.shape[0]
        dist = np.empty(n, np.float32)
        hit = np.empty((n, 2), np.int32)
        _check(lib().hs_occupancy_ray_batch(self._h, stream, level, b.ctypes.data, en.ctypes.data, n, dist.ctypes.data, hit.ctypes.data),
               "hs_occupancy_ray_batch")
        return dist, hit

    def export_stream_state(self, stream):
        st = HsStreamState()
        pts = np.zeros((self.max_points, 2), np.float32)
        _check(lib().hs_export_stream_state(self._h, stream, ctypes.byref(st), pts.ctypes.data), "hs_export_stream_state")
        return st, pts

    def import_stream_state(self, stream, state, coarse_points):
        coarse_points = _arr(coarse_points, np.float32, (self.max_points, 2))
        _check(lib().hs_import_stream_state(self._h, stream, ctypes.byref(state), coarse_points.ctypes.data), "hs_import_stream_state")

    # state access
    def poses(self, first=0, n=None):
        n = self.n_streams - first if n is None else n
        out = np.empty((n, 3), np.float32)
        _check(lib().hs_get_poses(self._h, first, n, out.ctypes.data), "hs_get_poses")
        return out

    def covariances(self, first=0, n=None):
        n = self.n_streams - first if n is None else n
        out = np.empty((n, 9), np.float32)
        _check(lib().hs_get_covariances(self._h, first, n, out.ctypes.data), "hs_get_covariances")
        return out.reshape(n, 3, 3)

    def last_map_update_poses(self, first=0, n=None):
        n = self.n_streams - first if n is None else n
        out = np.empty((n, 3), np.float32)
        _check(lib().hs_get_last_map_update_poses(self._h, first, n, out.ctypes.data), "hs_get_last_map_update_poses")
        return out

    def device_pose_ptr(self):
        return lib().hs_device_pose_ptr(self._h)

    def download_level(self, stream, level):
        sx, sy, _ = self.level_dims(level)
        out = np.empty(sx * sy, CELL_DTYPE)
        _check(lib().hs_download_level(self._h, stream, level, out.ctypes.data), "hs_download_level")
        return out

    def upload_level(self, stream, level, cells):
        cells = np.ascontiguousarray(cells, CELL_DTYPE)
        sx, sy, _ = self.level_dims(level)
        assert cells.size == sx * sy
        _check(lib().hs_upload_level(self._h, stream, level, cells.ctypes.data), "hs_upload_level")

    def update_index(self, stream, level=0):
        v = ctypes.c_int()
        _check(lib().hs_get_update_index(self._h, stream, level, ctypes.byref(v)), "hs_get_update_index")
        return v.value

    def map_info(self, level=0):
        """nav_msgs/OccupancyGrid metadata of a level (HectorMappingRos::setServiceGetMapData)."""
        mi = HsMapInfo()
        _check(lib().hs_get_map_info(self._h, level, ctypes.byref(mi)), "hs_get_map_info")
        return {"origin": (mi.origin_x, mi.origin_y), "resolution": mi.resolution, "width": mi.width, "height": mi.height}

    def classify_level(self, stream, level):
        sx, sy, _ = self.level_dims(level)
        out = np.empty(sx * sy, np.int8)
        _check(lib().hs_classify_level(self._h, stream, level, out.ctypes.data), "hs_classify_level")
        return out

    def map_epoch(self):
        return int(lib().hs_get_map_epoch(self._h))

    def copy_stream_maps_from(self, dst_stream, src_engine, src_stream):
        """All levels of src_engine's stream src_stream into this engine's dst_stream (same or peer GPU)."""
        _check(lib().hs_copy_stream_maps(self._h, dst_stream, src_engine._h, src_stream), "hs_copy_stream_maps")

    def profile_enable(self, on=True):
        _check(lib().hs_profile_enable(self._h, int(on)), "hs_profile_enable")

    def profile_collect(self):
        t = HsKernelTimes()
        _check(lib().hs_profile_collect(self._h, ctypes.byref(t)), "hs_profile_collect")
        return {KERNEL_NAMES[i]: {"ms": t.ms[i], "launches": int(t.launches[i])} for i in range(len(KERNEL_NAMES))}

    def stats(self):
        s = HsStats()
        _check(lib().hs_get_stats(self._h, ctypes.byref(s)), "hs_get_stats")
        return {k: getattr(s, k) for k, _ in HsStats._fields_}
