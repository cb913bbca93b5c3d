// TEST INFRASTRUCTURE ONLY -- never linked into or called by the product path.
//
// C-ABI wrapper around the UNMODIFIED reference headers of hector_slam_lib, compiled from
// where they lie under /root/reference (include path given on the command line, see
// oracle/Makefile) against the mini-Eigen / tf shims in oracle/shim. Output goes to
// oracle/_ref/libhsref.so (git-ignored). It is used
//   * to pin the plain-C restatement in oracle/hs_oracle.c (bit-exact comparison),
//   * to generate the golden fixtures under tests/golden/ (tests/golden/make_golden.py),
//   * optionally as the CPU baseline ("kind": "reference") in bench.py.
// Reference entry points exercised: HectorSlamProcessor (slam_main/HectorSlamProcessor.h:50-154),
// OccGridMapUtil::getCompleteHessianDerivs (map/OccGridMapUtil.h:64-104) and its off-path members
// (getResidualForState, getLikelihoodForState, getCovarianceForPose, interpMapValue: :106-285),
// HectorMapTools::DistanceMeasurementProvider (hector_map_tools/HectorMapTools.h:118-236),
// OccGridMapBase::updateByScan (map/OccGridMapBase.h:121-168),
// util::poseDifferenceLargerThan (util/UtilFunctions.h:73-92).

// <math.h>/<stdlib.h> first: upstream tf/LinearMath/Scalar.h does the same in the real ROS
// build, which makes unqualified abs/sin/cos/exp/log resolve to float overloads everywhere.
#include <math.h>
#include <stdlib.h>
#include <climits>
#include <cstring>
#include <iostream>
#include <vector>
#include <thread>
#include <chrono>

#include <Eigen/Core>
#include <Eigen/Geometry>
#include <tf/LinearMath/Transform.h>

#include "slam_main/HectorSlamProcessor.h"
#include "hector_map_tools/HectorMapTools.h"   // DistanceMeasurementProvider (occupancy ray queries)

namespace {

typedef hectorslam::OccGridMapUtilConfig<hectorslam::GridMap> Util;

struct MultiMapAccess : hectorslam::MapRepMultiMap {
  static std::vector<hectorslam::MapProcContainer>& containers(hectorslam::MapRepMultiMap& m) {
    return m.*(&MultiMapAccess::mapContainer);
  }
};

struct Proc : hectorslam::HectorSlamProcessor {
  Proc(float res, int sx, int sy, float startx, float starty, int levels)
      : hectorslam::HectorSlamProcessor(res, sx, sy, Eigen::Vector2f(startx, starty), levels) {}
  hectorslam::MapRepMultiMap& multi() { return *static_cast<hectorslam::MapRepMultiMap*>(mapRep); }
  hectorslam::MapProcContainer& level(int l) { return MultiMapAccess::containers(multi())[l]; }
  const Eigen::Vector3f& lastUpdatePose() const { return lastMapUpdatePose; }
  hectorslam::DataContainer scratch;
};

void fill(hectorslam::DataContainer& dc, const float* pts, int n, float ox, float oy) {
  dc.clear();
  dc.setOrigo(Eigen::Vector2f(ox, oy));
  for (int i = 0; i < n; ++i) dc.add(Eigen::Vector2f(pts[2 * i], pts[2 * i + 1]));
}

struct QuietCout {
  QuietCout() { std::cout.setstate(std::ios_base::failbit); }
} quiet_cout;

}  // namespace

extern "C" {

struct hsref_cell { float logOddsVal; int updateIndex; };

void* hsref_create(float res, int sx, int sy, float startx, float starty, int levels) {
  return new Proc(res, sx, sy, startx, starty, levels);
}
void hsref_destroy(void* h) { delete static_cast<Proc*>(h); }
void hsref_reset(void* h) { static_cast<Proc*>(h)->reset(); }
void hsref_set_update_factor_free(void* h, float f) { static_cast<Proc*>(h)->setUpdateFactorFree(f); }
void hsref_set_update_factor_occupied(void* h, float f) { static_cast<Proc*>(h)->setUpdateFactorOccupied(f); }
void hsref_set_map_update_min_dist_diff(void* h, float d) { static_cast<Proc*>(h)->setMapUpdateMinDistDiff(d); }
void hsref_set_map_update_min_angle_diff(void* h, float a) { static_cast<Proc*>(h)->setMapUpdateMinAngleDiff(a); }
float hsref_get_scale_to_map(void* h) { return static_cast<Proc*>(h)->getScaleToMap(); }
int hsref_get_map_levels(void* h) { return static_cast<Proc*>(h)->getMapLevels(); }

void hsref_update(void* h, const float* pts, int n, float ox, float oy, const float* hint, int map_without_matching) {
  Proc* p = static_cast<Proc*>(h);
  fill(p->scratch, pts, n, ox, oy);
  p->update(p->scratch, Eigen::Vector3f(hint[0], hint[1], hint[2]), map_without_matching != 0);
}
void hsref_get_pose(void* h, float* out3) {
  const Eigen::Vector3f& v = static_cast<Proc*>(h)->getLastScanMatchPose();
  out3[0] = v[0]; out3[1] = v[1]; out3[2] = v[2];
}
void hsref_get_last_map_update_pose(void* h, float* out3) {
  const Eigen::Vector3f& v = static_cast<Proc*>(h)->lastUpdatePose();
  out3[0] = v[0]; out3[1] = v[1]; out3[2] = v[2];
}
// row-major 3x3
void hsref_get_cov(void* h, float* out9) {
  const Eigen::Matrix3f& m = static_cast<Proc*>(h)->getLastScanMatchCovariance();
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) out9[3 * i + j] = m(i, j);
}
void hsref_level_dims(void* h, int level, int* sx, int* sy, float* cell_length) {
  const hectorslam::GridMap& g = static_cast<Proc*>(h)->getGridMap(level);
  *sx = g.getSizeX(); *sy = g.getSizeY(); *cell_length = g.getCellLength();
}
int hsref_get_update_index(void* h, int level) { return static_cast<Proc*>(h)->getGridMap(level).getUpdateIndex(); }
void hsref_download_level(void* h, int level, hsref_cell* out) {
  const hectorslam::GridMap& g = static_cast<Proc*>(h)->getGridMap(level);
  int n = g.getSizeX() * g.getSizeY();
  for (int i = 0; i < n; ++i) { out[i].logOddsVal = g.getCell(i).logOddsVal; out[i].updateIndex = g.getCell(i).updateIndex; }
}
void hsref_upload_level(void* h, int level, const hsref_cell* in) {
  Proc* p = static_cast<Proc*>(h);
  hectorslam::GridMap& g = p->level(level).getGridMap();
  int n = g.getSizeX() * g.getSizeY();
  for (int i = 0; i < n; ++i) { g.getCell(i).logOddsVal = in[i].logOddsVal; g.getCell(i).updateIndex = in[i].updateIndex; }
  p->level(level).resetCachedData();
}
void hsref_set_cell(void* h, int level, int index, float logodds, int update_index) {
  Proc* p = static_cast<Proc*>(h);
  hectorslam::GridMap& g = p->level(level).getGridMap();
  g.getCell(index).logOddsVal = logodds; g.getCell(index).updateIndex = update_index;
  p->level(level).resetCachedData();
}
void hsref_world_to_map(void* h, int level, const float* w, float* m) {
  Eigen::Vector3f r = static_cast<Proc*>(h)->getGridMap(level).getMapCoordsPose(Eigen::Vector3f(w[0], w[1], w[2]));
  m[0] = r[0]; m[1] = r[1]; m[2] = r[2];
}
void hsref_map_to_world(void* h, int level, const float* m, float* w) {
  Eigen::Vector3f r = static_cast<Proc*>(h)->getGridMap(level).getWorldCoordsPose(Eigen::Vector3f(m[0], m[1], m[2]));
  w[0] = r[0]; w[1] = r[1]; w[2] = r[2];
}

// One evaluation of the reference's Hessian/gradient accumulation on one level.
// pose is in MAP coordinates of that level, pts are already scaled to that level.
void hsref_hessian_derivs(void* h, int level, const float* pose_map, const float* pts, int n, float* H9, float* dTr3) {
  Proc* p = static_cast<Proc*>(h);
  fill(p->scratch, pts, n, 0.f, 0.f);
  Eigen::Matrix3f H; Eigen::Vector3f dTr;
  p->level(level).gridMapUtil->getCompleteHessianDerivs(Eigen::Vector3f(pose_map[0], pose_map[1], pose_map[2]), p->scratch, H, dTr);
  for (int i = 0; i < 3; ++i) { for (int j = 0; j < 3; ++j) H9[3 * i + j] = H(i, j); dTr3[i] = dTr[i]; }
}
void hsref_hessian_derivs_batch(void* h, int level, const float* poses_map, int n_poses, const float* pts, int n, float* H9, float* dTr3) {
  Proc* p = static_cast<Proc*>(h);
  fill(p->scratch, pts, n, 0.f, 0.f);
  Util* util = p->level(level).gridMapUtil;
  for (int k = 0; k < n_poses; ++k) {
    Eigen::Matrix3f H; Eigen::Vector3f dTr;
    util->getCompleteHessianDerivs(Eigen::Vector3f(poses_map[3 * k], poses_map[3 * k + 1], poses_map[3 * k + 2]), p->scratch, H, dTr);
    for (int i = 0; i < 3; ++i) { for (int j = 0; j < 3; ++j) H9[9 * k + 3 * i + j] = H(i, j); dTr3[3 * k + i] = dTr[i]; }
  }
}
// interpMapValueWithDerivatives at map coordinates (level grid)
void hsref_interp_with_derivs(void* h, int level, float x, float y, float* out3) {
  Proc* p = static_cast<Proc*>(h);
  Eigen::Vector3f r = p->level(level).gridMapUtil->interpMapValueWithDerivatives(Eigen::Vector2f(x, y));
  out3[0] = r[0]; out3[1] = r[1]; out3[2] = r[2];
}
// ScanMatcher::matchData on one level (pts already scaled to that level); world pose in/out.
void hsref_match_level(void* h, int level, const float* hint_world, const float* pts, int n, int max_iter, float* pose_out, float* cov9) {
  Proc* p = static_cast<Proc*>(h);
  fill(p->scratch, pts, n, 0.f, 0.f);
  Eigen::Matrix3f cov = Eigen::Matrix3f::Zero();
  Eigen::Vector3f r = p->level(level).matchData(Eigen::Vector3f(hint_world[0], hint_world[1], hint_world[2]), p->scratch, cov, max_iter);
  for (int i = 0; i < 3; ++i) { pose_out[i] = r[i]; for (int j = 0; j < 3; ++j) cov9[3 * i + j] = cov(i, j); }
}
// OccGridMapBase::updateByScan on one level (pts/origo already scaled to that level).
void hsref_update_by_scan_level(void* h, int level, const float* pts, int n, float ox, float oy, const float* pose_world) {
  Proc* p = static_cast<Proc*>(h);
  fill(p->scratch, pts, n, ox, oy);
  p->level(level).updateByScan(p->scratch, Eigen::Vector3f(pose_world[0], pose_world[1], pose_world[2]));
  p->level(level).resetCachedData();
}
// MapRepMultiMap::updateByScan over all levels with a given (teacher-forced) pose. Coarse levels use
// the containers refreshed by the last matchData call, exactly as in the reference (MapRepMultiMap.h:134-147),
// so call hsref_refresh_coarse first when no match preceded.
void hsref_refresh_coarse(void* h, const float* pts, int n, float ox, float oy) {
  // MapRepMultiMap::matchData is the only place that refreshes dataContainers (MapRepMultiMap.h:127);
  // run it for its side effect and discard the pose.
  Proc* p = static_cast<Proc*>(h);
  fill(p->scratch, pts, n, ox, oy);
  Eigen::Matrix3f cov;
  p->multi().matchData(Eigen::Vector3f(0.f, 0.f, 0.f), p->scratch, cov);
}
void hsref_update_by_scan_all(void* h, const float* pts, int n, float ox, float oy, const float* pose_world) {
  Proc* p = static_cast<Proc*>(h);
  fill(p->scratch, pts, n, ox, oy);
  p->multi().updateByScan(p->scratch, Eigen::Vector3f(pose_world[0], pose_world[1], pose_world[2]));
  p->multi().onMapUpdated();
}
// Off-path OccGridMapUtil members (map/OccGridMapUtil.h:106-285); pose in MAP coordinates, pts scaled to the level.
float hsref_interp_map_value(void* h, int level, float x, float y) {
  return static_cast<Proc*>(h)->level(level).gridMapUtil->interpMapValue(Eigen::Vector2f(x, y));
}
float hsref_residual_for_state(void* h, int level, const float* pose_map, const float* pts, int n) {
  Proc* p = static_cast<Proc*>(h);
  fill(p->scratch, pts, n, 0.f, 0.f);
  return p->level(level).gridMapUtil->getResidualForState(Eigen::Vector3f(pose_map[0], pose_map[1], pose_map[2]), p->scratch);
}
float hsref_likelihood_for_state(void* h, int level, const float* pose_map, const float* pts, int n) {
  Proc* p = static_cast<Proc*>(h);
  fill(p->scratch, pts, n, 0.f, 0.f);
  return p->level(level).gridMapUtil->getLikelihoodForState(Eigen::Vector3f(pose_map[0], pose_map[1], pose_map[2]), p->scratch);
}
void hsref_covariance_for_pose(void* h, int level, const float* pose_map, const float* pts, int n, float* cov_map9, float* cov_world9) {
  Proc* p = static_cast<Proc*>(h);
  fill(p->scratch, pts, n, 0.f, 0.f);
  Util* util = p->level(level).gridMapUtil;
  Eigen::Matrix3f c = util->getCovarianceForPose(Eigen::Vector3f(pose_map[0], pose_map[1], pose_map[2]), p->scratch);
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) cov_map9[3 * i + j] = c(i, j);
  if (cov_world9) {
    Eigen::Matrix3f w = util->getCovMatrixWorldCoords(c);
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) cov_world9[3 * i + j] = w(i, j);
  }
}

// HectorMapTools::DistanceMeasurementProvider::checkOccupancyBresenhami on a caller-supplied occupancy grid.
float hsref_check_occupancy_bresenhami(const signed char* grid, int size_x, int size_y, int x0, int y0, int x1, int y1, int* hit_xy) {
  std::shared_ptr<nav_msgs::OccupancyGrid> map(new nav_msgs::OccupancyGrid());
  map->info.resolution = 1.0f;
  map->info.width = (uint32_t)size_x;
  map->info.height = (uint32_t)size_y;
  map->info.origin.position.x = 0.0; map->info.origin.position.y = 0.0; map->info.origin.position.z = 0.0;
  map->data.assign(grid, grid + (size_t)size_x * size_y);
  HectorMapTools::DistanceMeasurementProvider dmp;
  dmp.setMap(map);
  Eigen::Vector2i hit(-1, -1);
  float d = dmp.checkOccupancyBresenhami(Eigen::Vector2i(x0, y0), Eigen::Vector2i(x1, y1), &hit);
  if (hit_xy && d >= 0.0f) { hit_xy[0] = hit[0]; hit_xy[1] = hit[1]; }
  return d;
}

int hsref_pose_difference_larger_than(const float* p1, const float* p2, float d, float a) {
  return util::poseDifferenceLargerThan(Eigen::Vector3f(p1[0], p1[1], p1[2]), Eigen::Vector3f(p2[0], p2[1], p2[2]), d, a) ? 1 : 0;
}
float hsref_normalize_angle(float a) { return util::normalize_angle(a); }

// ---------------------------------------------------------------------------------------------
// Multi-stream runner (CPU baseline): one HectorSlamProcessor per stream, streams statically
// partitioned over `threads` std::threads. pts: [n_scans][n_streams][max_pts][2], counts: [n_scans][n_streams].
// `stream_major` picks the loop order (see below). Pose hint = previous matched pose (HectorMappingRos.cpp:254). Returns wall seconds spent in update().
// poses_out (optional): [n_scans][n_streams][3]; cov_out (optional): [n_scans][n_streams][9].
double hsref_run_streams(void** procs, int n_streams, int n_scans, int max_pts, const float* pts, const int* counts,
                         int threads, int stream_major, float* poses_out, float* cov_out) {
  if (threads < 1) threads = 1;
  auto work = [&](int t) {
    int s0 = (int)((long long)n_streams * t / threads), s1 = (int)((long long)n_streams * (t + 1) / threads);
    hectorslam::DataContainer dc;
    // stream_major: finish one stream before the next (map stays cache-hot: the CPU's best case);
    // otherwise scan-major, the order a live batched system would see.
    long long total = (long long)n_scans * (s1 - s0);
    for (long long it = 0; it < total; ++it) {
      int k = stream_major ? (int)(it % n_scans) : (int)(it / (s1 - s0));
      int s = s0 + (stream_major ? (int)(it / n_scans) : (int)(it % (s1 - s0)));
      {
        Proc* p = static_cast<Proc*>(procs[s]);
        size_t slot = (size_t)k * n_streams + s;
        fill(dc, pts + slot * max_pts * 2, counts[slot], 0.f, 0.f);
        Eigen::Vector3f hint = p->getLastScanMatchPose();
        p->update(dc, hint);
        if (poses_out) { const Eigen::Vector3f& v = p->getLastScanMatchPose(); for (int i = 0; i < 3; ++i) poses_out[slot * 3 + i] = v[i]; }
        if (cov_out) { const Eigen::Matrix3f& m = p->getLastScanMatchCovariance(); for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) cov_out[slot * 9 + 3 * i + j] = m(i, j); }
      }
    }
  };
  auto t0 = std::chrono::steady_clock::now();
  std::vector<std::thread> pool;
  for (int t = 1; t < threads; ++t) pool.emplace_back(work, t);
  work(0);
  for (auto& th : pool) th.join();
  return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}

}  // extern "C"
