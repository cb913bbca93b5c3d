// hsgpu::BatchEngine -- C++ RAII wrapper over the C ABI (hsgpu.h) for callers that drive MANY HectorSlamProcessor states at
// once. Where include/hsgpu/HectorSlamProcessor.h is the one-stream drop-in for the reference's class
// (HSL/slam_main/HectorSlamProcessor.h:50-154), this is the same interface widened to a batch: every method maps 1:1 to a
// method of the reference class applied to n streams, errors become exceptions (the C ABI's status codes), and there is no
// CPU fallback -- construction throws without a usable CUDA device.
#ifndef HSGPU_BATCHENGINE_H
#define HSGPU_BATCHENGINE_H

#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

#include "../hsgpu.h"

namespace hsgpu {

class Error : public std::runtime_error {
 public:
  Error(int status, const char* what) : std::runtime_error(std::string(what) + ": " + hs_status_string(status)), status_(status) {}
  int status() const { return status_; }

 private:
  int status_;
};

class BatchEngine {
 public:
  // HectorSlamProcessor(mapResolution, mapSizeX, mapSizeY, startCoords, multi_res_size) x n_streams (HectorSlamProcessor.h:54-64)
  BatchEngine(float mapResolution, int mapSizeX, int mapSizeY, float startX, float startY, int multi_res_size, int n_streams,
              int max_points, int device = 0)
      : e_(nullptr) {
    hs_config cfg;
    cfg.map_resolution = mapResolution; cfg.map_size_x = mapSizeX; cfg.map_size_y = mapSizeY;
    cfg.start_x = startX; cfg.start_y = startY; cfg.levels = multi_res_size;
    cfg.n_streams = n_streams; cfg.max_points = max_points; cfg.device = device;
    check(hs_create(&cfg, &e_), "hs_create");
  }
  ~BatchEngine() { if (e_) hs_destroy(e_); }
  BatchEngine(const BatchEngine&) = delete;
  BatchEngine& operator=(const BatchEngine&) = delete;

  // setters of HectorSlamProcessor (HectorSlamProcessor.h:136-139), applied to every stream
  void setUpdateFactorFree(float f) { check(hs_set_update_factor_free(e_, f), "hs_set_update_factor_free"); }
  void setUpdateFactorOccupied(float f) { check(hs_set_update_factor_occupied(e_, f), "hs_set_update_factor_occupied"); }
  void setMapUpdateMinDistDiff(float d) { check(hs_set_map_update_min_dist_diff(e_, d), "hs_set_map_update_min_dist_diff"); }
  void setMapUpdateMinAngleDiff(float a) { check(hs_set_map_update_min_angle_diff(e_, a), "hs_set_map_update_min_angle_diff"); }
  float getScaleToMap() const { return hs_get_scale_to_map(e_); }
  int getMapLevels() const { return hs_get_map_levels(e_); }
  int streams() const { return hs_get_n_streams(e_); }
  int maxPoints() const { return hs_get_max_points(e_); }
  void reset(int stream = HS_ALL_STREAMS) { check(hs_reset(e_, stream), "hs_reset"); }

  // update() for n streams (HectorSlamProcessor.h:71-113): points [n][maxPoints][2] in level-0 cells, counts [n]; optional
  // stream_ids / origos / poseHints (default: each stream's last pose) / map_without_matching; poses_out [n][3], cov_out [n][9].
  void update(int n, const float* points, const int32_t* counts, float* poses_out, float* cov_out = nullptr,
              const int32_t* stream_ids = nullptr, const float* origos = nullptr, const float* poseHints = nullptr,
              const uint8_t* map_without_matching = nullptr) {
    check(hs_update_batch(e_, n, stream_ids, points, counts, origos, poseHints, map_without_matching, poses_out, cov_out), "hs_update_batch");
  }
  // the same without waiting (pinned buffers; see hs_submit_batch), and the matching wait
  void submit(int n, const float* points, const int32_t* counts, float* poses_out, const int32_t* stream_ids = nullptr,
              const float* origos = nullptr, const float* poseHints = nullptr, const uint8_t* map_without_matching = nullptr) {
    check(hs_submit_batch(e_, n, stream_ids, points, counts, origos, poseHints, map_without_matching, poses_out), "hs_submit_batch");
  }
  void wait() { check(hs_wait(e_), "hs_wait"); }
  // multi-start matchData of one scan against one stream's maps (relocalisation): n_hyp world pose hints in, as many matched
  // poses (and level-0 Hessians) out; the stream's state is not touched (hs_match_hypotheses)
  void matchHypotheses(int stream, const float* points, int n_points, const float* poseHints, int n_hyp, float* poses_out,
                       float* cov_out = nullptr) {
    check(hs_match_hypotheses(e_, stream, points, n_points, poseHints, n_hyp, poses_out, cov_out), "hs_match_hypotheses");
  }

  // getLastScanMatchPose / getLastScanMatchCovariance (HectorSlamProcessor.h:126-127) of streams [first, first + n)
  std::vector<float> lastScanMatchPoses(int first, int n) {
    std::vector<float> out((size_t)n * 3);
    check(hs_get_poses(e_, first, n, out.data()), "hs_get_poses");
    return out;
  }
  std::vector<float> lastScanMatchCovariances(int first, int n) {
    std::vector<float> out((size_t)n * 9);
    check(hs_get_covariances(e_, first, n, out.data()), "hs_get_covariances");
    return out;
  }
  // getGridMap(level) of one stream (HectorSlamProcessor.h:131): cells row-major, LogOddsCell layout
  std::vector<hs_cell> gridMap(int stream, int level) {
    int sx = 0, sy = 0;
    check(hs_get_level_dims(e_, level, &sx, &sy, nullptr), "hs_get_level_dims");
    std::vector<hs_cell> cells((size_t)sx * sy);
    check(hs_download_level(e_, stream, level, cells.data()), "hs_download_level");
    return cells;
  }
  int updateIndex(int stream, int level = 0) {
    int v = 0;
    check(hs_get_update_index(e_, stream, level, &v), "hs_get_update_index");
    return v;
  }
  hs_engine* handle() { return e_; }

 private:
  static void check(int status, const char* what) { if (status != HS_OK) throw Error(status, what); }
  hs_engine* e_;
};

}  // namespace hsgpu
#endif
