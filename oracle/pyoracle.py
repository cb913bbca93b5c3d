"""TEST INFRASTRUCTURE ONLY -- ctypes loaders for the CPU oracle.

  kind="port": oracle/libhsoracle.so, the plain-C restatement (oracle/hs_oracle.c), prefix hso_
  kind="ref" : oracle/_ref/libhsref.so, the UNMODIFIED reference headers + mini-Eigen shim, prefix hsref_

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this.
The product (tfe-slam-ucl_b200/libhsgpu.so) never does.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
PORT_LIB = os.path.join(_HERE, "libhsoracle.so")
REF_LIB = os.path.join(_HERE, "_ref", "libhsref.so")
CELL_DTYPE = np.dtype([("logOddsVal", np.float32), ("updateIndex", np.int32)])

_P, _I, _F = ctypes.c_void_p, ctypes.c_int, ctypes.c_float
_libs = {}


def build(ref=True):
    """Compile the C restatement and, when /root/reference is present, oracle/_ref."""
    targets = ["port"] + (["ref"] if ref else [])
    res = subprocess.run(["make", "-C", _HERE] + targets, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("building the oracle failed:\n" + res.stdout + res.stderr)


def have(kind):
    return os.path.exists(PORT_LIB if kind == "port" else REF_LIB)


def _load(kind):
    if kind in _libs:
        return _libs[kind]
    path, pre = (PORT_LIB, "hso_") if kind == "port" else (REF_LIB, "hsref_")
    if not os.path.exists(path):
        if kind == "port":
            build(ref=False)
        else:
            raise FileNotFoundError(path)
    L = ctypes.CDLL(path)
    sig = {
        "create": (_P, [_F, _I, _I, _F, _F, _I]), "destroy": (None, [_P]), "reset": (None, [_P]),
        "set_update_factor_free": (None, [_P, _F]), "set_update_factor_occupied": (None, [_P, _F]),
        "set_map_update_min_dist_diff": (None, [_P, _F]), "set_map_update_min_angle_diff": (None, [_P, _F]),
        "get_scale_to_map": (_F, [_P]), "get_map_levels": (_I, [_P]),
        "update": (None, [_P, _P, _I, _F, _F, _P, _I]),
        "get_pose": (None, [_P, _P]), "get_last_map_update_pose": (None, [_P, _P]), "get_cov": (None, [_P, _P]),
        "level_dims": (None, [_P, _I, _P, _P, _P]), "get_update_index": (_I, [_P, _I]),
        "download_level": (None, [_P, _I, _P]), "upload_level": (None, [_P, _I, _P]),
        "set_cell": (None, [_P, _I, _I, _F, _I]),
        "world_to_map": (None, [_P, _I, _P, _P]), "map_to_world": (None, [_P, _I, _P, _P]),
        "hessian_derivs": (None, [_P, _I, _P, _P, _I, _P, _P]),
        "hessian_derivs_batch": (None, [_P, _I, _P, _I, _P, _I, _P, _P]),
        "interp_with_derivs": (None, [_P, _I, _F, _F, _P]),
        "match_level": (None, [_P, _I, _P, _P, _I, _I, _P, _P]),
        "update_by_scan_level": (None, [_P, _I, _P, _I, _F, _F, _P]),
        "refresh_coarse": (None, [_P, _P, _I, _F, _F]),
        "update_by_scan_all": (None, [_P, _P, _I, _F, _F, _P]),
        "pose_difference_larger_than": (_I, [_P, _P, _F, _F]),
        "normalize_angle": (_F, [_F]),
        "run_streams": (ctypes.c_double, [_P, _I, _I, _I, _P, _P, _I, _I, _P, _P]),
        "interp_map_value": (_F, [_P, _I, _F, _F]),
        "residual_for_state": (_F, [_P, _I, _P, _P, _I]),
        "likelihood_for_state": (_F, [_P, _I, _P, _P, _I]),
        "covariance_for_pose": (None, [_P, _I, _P, _P, _I, _P, _P]),
        "check_occupancy_bresenhami": (_F, [_P, _I, _I, _I, _I, _I, _I, _P]),
    }
    if kind == "port":
        sig["get_cell_visits"] = (ctypes.c_ulonglong, [_P])
        sig["get_integrations"] = (ctypes.c_ulonglong, [_P])
        sig["scan_to_points"] = (_I, [_P, _I, _F, _F, _F, _F, _F, _P])
        sig["cloud_to_points"] = (_I, [_P, _I, _P, _F, _F, _F, _F, _F, _P, _P])
    fns = {}
    for name, (res, args) in sig.items():
        fn = getattr(L, pre + name)
        fn.restype = res
        fn.argtypes = args
        fns[name] = fn
    _libs[kind] = fns
    return fns


def _f32(a):
    return np.ascontiguousarray(a, np.float32)


class CpuProcessor:
    """One hectorslam::HectorSlamProcessor on the CPU (the reference itself or its C restatement)."""

    def __init__(self, kind="port", map_resolution=0.05, map_size_x=1024, map_size_y=1024, start=(0.5, 0.5), levels=3):
        self.kind = kind
        self.f = _load(kind)
        self.h = self.f["create"](map_resolution, map_size_x, map_size_y, start[0], start[1], levels)
        if not self.h:
            raise RuntimeError("oracle create failed")
        self.levels = levels

    def close(self):
        if getattr(self, "h", None):
            self.f["destroy"](self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def reset(self): self.f["reset"](self.h)
    def set_update_factor_free(self, v): self.f["set_update_factor_free"](self.h, v)
    def set_update_factor_occupied(self, v): self.f["set_update_factor_occupied"](self.h, v)
    def set_map_update_min_dist_diff(self, v): self.f["set_map_update_min_dist_diff"](self.h, v)
    def set_map_update_min_angle_diff(self, v): self.f["set_map_update_min_angle_diff"](self.h, v)
    def scale_to_map(self): return self.f["get_scale_to_map"](self.h)

    def update(self, points, n=None, origo=(0.0, 0.0), hint=(0.0, 0.0, 0.0), map_without_matching=False):
        points = _f32(points)
        n = points.reshape(-1, 2).shape[0] if n is None else int(n)
        hint = _f32(hint)
        self.f["update"](self.h, points.ctypes.data, n, float(origo[0]), float(origo[1]), hint.ctypes.data, int(map_without_matching))
        return self.pose()

    def pose(self):
        out = np.zeros(3, np.float32)
        self.f["get_pose"](self.h, out.ctypes.data)
        return out

    def last_map_update_pose(self):
        out = np.zeros(3, np.float32)
        self.f["get_last_map_update_pose"](self.h, out.ctypes.data)
        return out

    def cov(self):
        out = np.zeros(9, np.float32)
        self.f["get_cov"](self.h, out.ctypes.data)
        return out.reshape(3, 3)

    def level_dims(self, level):
        sx, sy, cl = ctypes.c_int(), ctypes.c_int(), ctypes.c_float()
        self.f["level_dims"](self.h, level, ctypes.byref(sx), ctypes.byref(sy), ctypes.byref(cl))
        return sx.value, sy.value, cl.value

    def update_index(self, level=0): return self.f["get_update_index"](self.h, level)

    def download_level(self, level):
        sx, sy, _ = self.level_dims(level)
        out = np.empty(sx * sy, CELL_DTYPE)
        self.f["download_level"](self.h, level, out.ctypes.data)
        return out

    def upload_level(self, level, cells):
        cells = np.ascontiguousarray(cells, CELL_DTYPE)
        self.f["upload_level"](self.h, level, cells.ctypes.data)

    def set_cell(self, level, index, logodds, update_index=-1): self.f["set_cell"](self.h, level, int(index), float(logodds), int(update_index))

    def world_to_map(self, level, pose):
        pose = _f32(pose); out = np.zeros(3, np.float32)
        self.f["world_to_map"](self.h, level, pose.ctypes.data, out.ctypes.data)
        return out

    def map_to_world(self, level, pose):
        pose = _f32(pose); out = np.zeros(3, np.float32)
        self.f["map_to_world"](self.h, level, pose.ctypes.data, out.ctypes.data)
        return out

    def hessian_derivs_batch(self, level, poses_map, points):
        poses_map = _f32(poses_map).reshape(-1, 3)
        points = _f32(points).reshape(-1, 2)
        n = poses_map.shape[0]
        H = np.zeros((n, 9), np.float32); d = np.zeros((n, 3), np.float32)
        self.f["hessian_derivs_batch"](self.h, level, poses_map.ctypes.data, n, points.ctypes.data, points.shape[0], H.ctypes.data, d.ctypes.data)
        return H.reshape(n, 3, 3), d

    def interp_with_derivs(self, level, x, y):
        out = np.zeros(3, np.float32)
        self.f["interp_with_derivs"](self.h, level, float(x), float(y), out.ctypes.data)
        return out

    def match_level(self, level, hint_world, points, max_iter):
        points = _f32(points).reshape(-1, 2); hint_world = _f32(hint_world)
        pose = np.zeros(3, np.float32); cov = np.zeros(9, np.float32)
        self.f["match_level"](self.h, level, hint_world.ctypes.data, points.ctypes.data, points.shape[0], max_iter, pose.ctypes.data, cov.ctypes.data)
        return pose, cov.reshape(3, 3)

    def update_by_scan_level(self, level, points, pose_world, origo=(0.0, 0.0)):
        points = _f32(points).reshape(-1, 2); pose_world = _f32(pose_world)
        self.f["update_by_scan_level"](self.h, level, points.ctypes.data, points.shape[0], float(origo[0]), float(origo[1]), pose_world.ctypes.data)

    def raycast(self, points, pose_world, n=None, origo=(0.0, 0.0)):
        """refresh coarse containers from `points`, then MapRepMultiMap::updateByScan + onMapUpdated."""
        points = _f32(points); pose_world = _f32(pose_world)
        n = points.reshape(-1, 2).shape[0] if n is None else int(n)
        self.f["refresh_coarse"](self.h, points.ctypes.data, n, float(origo[0]), float(origo[1]))
        self.f["update_by_scan_all"](self.h, points.ctypes.data, n, float(origo[0]), float(origo[1]), pose_world.ctypes.data)

    def refresh_coarse(self, points, n=None, origo=(0.0, 0.0)):
        """MapRepMultiMap::matchData run for its side effect on dataContainers (MapRepMultiMap.h:127); pose discarded."""
        points = _f32(points)
        n = points.reshape(-1, 2).shape[0] if n is None else int(n)
        self.f["refresh_coarse"](self.h, points.ctypes.data, n, float(origo[0]), float(origo[1]))

    def integrate(self, points, pose_world, n=None, origo=(0.0, 0.0)):
        """MapRepMultiMap::updateByScan + onMapUpdated (MapRepMultiMap.h:134-147): the coarse levels use whatever the last
        matchData left in dataContainers."""
        points = _f32(points); pose_world = _f32(pose_world)
        n = points.reshape(-1, 2).shape[0] if n is None else int(n)
        self.f["update_by_scan_all"](self.h, points.ctypes.data, n, float(origo[0]), float(origo[1]), pose_world.ctypes.data)

    # off-path OccGridMapUtil members (OccGridMapUtil.h:106-285); poses in MAP coordinates, points scaled to the level
    def interp_map_value(self, level, x, y): return float(self.f["interp_map_value"](self.h, level, float(x), float(y)))

    def residual_for_state(self, level, pose_map, points):
        points = _f32(points).reshape(-1, 2); pose_map = _f32(pose_map)
        return float(self.f["residual_for_state"](self.h, level, pose_map.ctypes.data, points.ctypes.data, points.shape[0]))

    def likelihood_for_state(self, level, pose_map, points):
        points = _f32(points).reshape(-1, 2); pose_map = _f32(pose_map)
        return float(self.f["likelihood_for_state"](self.h, level, pose_map.ctypes.data, points.ctypes.data, points.shape[0]))

    def covariance_for_pose(self, level, pose_map, points):
        """(covMatrixMap, getCovMatrixWorldCoords(covMatrixMap)); entries the reference leaves unset in the latter are 0 here."""
        points = _f32(points).reshape(-1, 2); pose_map = _f32(pose_map)
        cm = np.zeros(9, np.float32); cw = np.zeros(9, np.float32)
        self.f["covariance_for_pose"](self.h, level, pose_map.ctypes.data, points.ctypes.data, points.shape[0], cm.ctypes.data, cw.ctypes.data)
        return cm.reshape(3, 3), cw.reshape(3, 3)

    def cell_visits(self): return int(self.f["get_cell_visits"](self.h))
    def integrations(self): return int(self.f["get_integrations"](self.h))


def cloud_to_points(cloud_xyz, transform12, scale_to_map, sqr_min_dist, sqr_max_dist, z_min, z_max):
    """rosPointCloudToDataContainer (C restatement only: the node needs ROS/tf): (points [k][2], origo [2])."""
    cloud_xyz = _f32(cloud_xyz).reshape(-1, 3)
    t = np.ascontiguousarray(transform12, np.float64).reshape(12)
    pts = np.zeros((max(cloud_xyz.shape[0], 1), 2), np.float32)
    origo = np.zeros(2, np.float32)
    k = _load("port")["cloud_to_points"](cloud_xyz.ctypes.data, cloud_xyz.shape[0], t.ctypes.data, scale_to_map, sqr_min_dist,
                                         sqr_max_dist, z_min, z_max, pts.ctypes.data, origo.ctypes.data)
    return pts[:k].copy(), origo


def check_occupancy_bresenhami(kind, grid, begin, end):
    """HectorMapTools::DistanceMeasurementProvider::checkOccupancyBresenhami on an int8 occupancy grid [sy][sx]:
    (distance in cells or -1, hit cell or None)."""
    grid = np.ascontiguousarray(grid, np.int8)
    hit = np.full(2, -1, np.int32)
    d = _load(kind)["check_occupancy_bresenhami"](grid.ctypes.data, grid.shape[1], grid.shape[0], int(begin[0]), int(begin[1]),
                                                  int(end[0]), int(end[1]), hit.ctypes.data)
    return float(d), (hit.copy() if d >= 0 else None)


def pose_difference_larger_than(kind, p1, p2, d, a):
    f = _load(kind)
    p1 = _f32(p1); p2 = _f32(p2)
    return bool(f["pose_difference_larger_than"](p1.ctypes.data, p2.ctypes.data, d, a))


def normalize_angle(kind, a):
    return _load(kind)["normalize_angle"](float(a))


def scan_to_points(ranges, n_beams, angle_min, angle_increment, range_min, range_max, scale_to_map):
    f = _load("port")
    ranges = _f32(ranges)
    out = np.zeros((n_beams, 2), np.float32)
    n = f["scan_to_points"](ranges.ctypes.data, n_beams, angle_min, angle_increment, range_min, range_max, scale_to_map, out.ctypes.data)
    return out[:n].copy()


def run_streams(procs, points, counts, threads=1, stream_major=True, want_poses=True, want_cov=False):
    """points [n_scans][n_streams][max_pts][2], counts [n_scans][n_streams]; hint = previous pose.
    Returns (seconds, poses [n_scans][n_streams][3] | None, cov | None)."""
    kind = procs[0].kind
    f = _load(kind)
    points = _f32(points); counts = np.ascontiguousarray(counts, np.int32)
    n_scans, n_streams, max_pts = points.shape[0], points.shape[1], points.shape[2]
    handles = (ctypes.c_void_p * n_streams)(*[p.h for p in procs])
    poses = np.zeros((n_scans, n_streams, 3), np.float32) if want_poses else None
    cov = np.zeros((n_scans, n_streams, 9), np.float32) if want_cov else None
    secs = f["run_streams"](handles, n_streams, n_scans, max_pts, points.ctypes.data, counts.ctypes.data, threads, int(stream_major),
                            poses.ctypes.data if want_poses else None, cov.ctypes.data if want_cov else None)
    return secs, poses, cov
