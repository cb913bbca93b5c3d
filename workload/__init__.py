"""Synthetic scan streams: ctypes binding over workload/libhsscangen.so (include/hsscangen.h).

The workload generator shared by the tests and by both arms of bench.py. It is neither the product (it never loads
libhsgpu.so) nor the oracle: deterministic LaserScan range arrays of a sensor driving a closed loop through a room with
box obstacles, plus a copy of the product's host conversion helper
(rosLaserScanToDataContainer, HectorMappingRos.cpp:483-507) so that CPU-only runs need no CUDA library.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libhsscangen.so")

SCENE_RPLIDAR_ROOM = 0
SCENE_HIRES_HALL = 1


class SceneConfig(ctypes.Structure):
    _fields_ = [("n_beams", ctypes.c_int32), ("angle_min", ctypes.c_float), ("angle_increment", ctypes.c_float),
                ("range_min", ctypes.c_float), ("range_max", ctypes.c_float), ("noise_sigma", ctypes.c_float),
                ("room_w", ctypes.c_float), ("room_h", ctypes.c_float), ("n_boxes", ctypes.c_int32),
                ("box_min", ctypes.c_float), ("box_max", ctypes.c_float), ("traj_rx", ctypes.c_float),
                ("traj_ry", ctypes.c_float), ("traj_period", ctypes.c_int32)]


def build():
    res = subprocess.run(["make", "-C", _HERE], capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("building libhsscangen.so failed:\n" + res.stdout + res.stderr)
    return LIB_PATH


_lib = None
_P, _I, _F = ctypes.c_void_p, ctypes.c_int, ctypes.c_float


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            build()
        L = ctypes.CDLL(LIB_PATH)
        L.hs_scene_preset.restype = _I
        L.hs_scene_preset.argtypes = [_I, _P]
        L.hs_scangen_batch.restype = _I
        L.hs_scangen_batch.argtypes = [_P, ctypes.c_uint64, _I, _I, _I, _P, _P, _I]
        L.hs_scan_to_points_host_batch.restype = _I
        L.hs_scan_to_points_host_batch.argtypes = [_P, _I, _I, _F, _F, _F, _F, _F, _I, _P, _P, _I]
        _lib = L
    return _lib


def _check(status, what):
    if status != 0:
        raise RuntimeError("%s failed with status %d" % (what, status))


def scene_preset(preset=SCENE_RPLIDAR_ROOM):
    sc = SceneConfig()
    _check(lib().hs_scene_preset(preset, ctypes.byref(sc)), "hs_scene_preset")
    return sc


def scangen_batch(scene, seed0, n_streams, first_scan, n_scans, threads=None, truth=True, out=None):
    """ranges [n_scans][n_streams][n_beams] (and truth poses [n_scans][n_streams][3])."""
    threads = threads or (os.cpu_count() or 1)
    ranges = out if out is not None else np.empty((n_scans, n_streams, scene.n_beams), np.float32)
    tr = np.empty((n_scans, n_streams, 3), np.float32) if truth else None
    _check(lib().hs_scangen_batch(ctypes.byref(scene), seed0, n_streams, first_scan, n_scans, ranges.ctypes.data,
                                  None if tr is None else tr.ctypes.data, threads), "hs_scangen_batch")
    return ranges, tr


def scan_to_points_host(ranges, scene, scale_to_map, max_points=None):
    """rosLaserScanToDataContainer over a [..., n_beams] array -> (points [..., max_points, 2], counts [...])."""
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
