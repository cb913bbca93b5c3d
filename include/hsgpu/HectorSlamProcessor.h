// Drop-in replacement for hector_slam_lib's  slam_main/HectorSlamProcessor.h  backed by the hsgpu CUDA engine.
//
// Include this header INSTEAD of the reference's "slam_main/HectorSlamProcessor.h"
// (src/hector_slam/hector_mapping/include/hector_slam_lib/slam_main/HectorSlamProcessor.h:50-154): it defines
// hectorslam::HectorSlamProcessor with the same constructor, methods, argument meaning and (absent) error
// behaviour, plus the small part of hectorslam::GridMap / DataContainer that the only caller, HectorMappingRos,
// touches. All arithmetic happens on the GPU through the C ABI in hsgpu.h; this class is
// a batch-of-one host wrapper that keeps a lazily refreshed host mirror of each level's cells for getGridMap().
//
// Differences from the reference, by design:
//   * DrawInterface / HectorDebugInfoInterface pointers are accepted; the per-update draw calls of
//     HectorSlamProcessor::update (:96-112) are issued, the matcher-internal ones (ScanMatcher.h:56-66,100-109) are not.
//   * the "HectorSM map lvl" banner (MapRepMultiMap.h:60) and the "SearchDir angle change too large" message
//     (ScanMatcher.h:211,214) are not printed; the latter is counted (hs_stats.angle_clamps).
//   * a DataContainer larger than HSGPU_FACADE_MAX_POINTS throws std::length_error (the reference has no limit).
//   * construction throws std::runtime_error when no CUDA device is usable: there is no CPU fallback.
// Needs Eigen (Vector2f/Vector3f/Matrix3f/Vector2i) exactly as the reference header does.
#ifndef HSGPU_HECTORSLAMPROCESSOR_H
#define HSGPU_HECTORSLAMPROCESSOR_H

#include <Eigen/Core>

#include <cstdlib>
#include <mutex>
#include <stdexcept>
#include <string>
#include <vector>

#include "../hsgpu.h"

#ifndef HSGPU_FACADE_MAX_POINTS
#define HSGPU_FACADE_MAX_POINTS 2048
#endif

// ---- interfaces of hector_slam_lib/util (same declarations as the reference: DrawInterface.h:32-42,
//      HectorDebugInfoInterface.h:32-38, MapLockerInterface.h:32-37) ----
#ifndef drawinterface_h__
#define drawinterface_h__
class DrawInterface {
 public:
  virtual void drawPoint(const Eigen::Vector2f& pointWorldFrame) = 0;
  virtual void drawArrow(const Eigen::Vector3f& poseWorld) = 0;
  virtual void drawCovariance(const Eigen::Vector2f& mean, const Eigen::Matrix2f& cov) = 0;
  virtual void setScale(double scale) = 0;
  virtual void setColor(double r, double g, double b, double a = 1.0) = 0;
  virtual void sendAndResetData() = 0;
};
#endif
#ifndef hectordebuginfointerface_h__
#define hectordebuginfointerface_h__
class HectorDebugInfoInterface {
 public:
  virtual void sendAndResetData() = 0;
  virtual void addHessianMatrix(const Eigen::Matrix3f& hessian) = 0;
  virtual void addPoseLikelihood(float lh) = 0;
};
#endif
#ifndef maplockerinterface_h__
#define maplockerinterface_h__
class MapLockerInterface {
 public:
  virtual void lockMap() = 0;
  virtual void unlockMap() = 0;
};
#endif

namespace hectorslam {

// DataPointContainer<Eigen::Vector2f> (scan/DataPointContainer.h:36-96)
#ifndef __DataPointContainer_h_
#define __DataPointContainer_h_
template <typename DataPointType>
class DataPointContainer {
 public:
  DataPointContainer(int size = 1000) { dataPoints.reserve(size); }
  void setFrom(const DataPointContainer& other, float factor) {
    origo = other.getOrigo() * factor;
    dataPoints = other.dataPoints;
    for (size_t i = 0; i < dataPoints.size(); ++i) dataPoints[i] *= factor;
  }
  void add(const DataPointType& dataPoint) { dataPoints.push_back(dataPoint); }
  void clear() { dataPoints.clear(); }
  int getSize() const { return (int)dataPoints.size(); }
  const DataPointType& getVecEntry(int index) const { return dataPoints[index]; }
  DataPointType getOrigo() const { return origo; }
  void setOrigo(const DataPointType& origoIn) { origo = origoIn; }

 protected:
  std::vector<DataPointType> dataPoints;
  DataPointType origo;
};
typedef DataPointContainer<Eigen::Vector2f> DataContainer;
#endif

// Host mirror of one level: the part of GridMap (OccGridMapBase<LogOddsCell,...>, GridMapBase) its callers use.
class GridMap {
 public:
  int getSizeX() const { return dims_[0]; }
  int getSizeY() const { return dims_[1]; }
  const Eigen::Vector2i& getMapDimensions() const { return dims_; }
  float getCellLength() const { return cell_length_; }
  float getScaleToMap() const { return scale_; }
  int getUpdateIndex() const { return update_index_; }
  // LogOddsCell::isOccupied / isFree (GridMapLogOdds.h:73-81)
  bool isOccupied(int index) const { return cells_[index].logOddsVal > 0.0f; }
  bool isFree(int index) const { return cells_[index].logOddsVal < 0.0f; }
  bool isOccupied(int x, int y) const { return isOccupied(y * dims_[0] + x); }
  bool isFree(int x, int y) const { return isFree(y * dims_[0] + x); }
  const hs_cell& getCell(int index) const { return cells_[index]; }
  const hs_cell& getCell(int x, int y) const { return cells_[y * dims_[0] + x]; }
  bool hasGridValue(int x, int y) const { return (x >= 0) && (y >= 0) && (x < dims_[0]) && (y < dims_[1]); }
  // GridMapBase::getWorldCoords / getMapCoords / *Pose (GridMapBase.h:210-239), same fp32 operations
  Eigen::Vector2f getWorldCoords(const Eigen::Vector2f& m) const { return Eigen::Vector2f(a_ * m[0] + twx_, a_ * m[1] + twy_); }
  Eigen::Vector2f getMapCoords(const Eigen::Vector2f& w) const { return Eigen::Vector2f(scale_ * w[0] + tmx_, scale_ * w[1] + tmy_); }
  Eigen::Vector3f getWorldCoordsPose(const Eigen::Vector3f& m) const { return Eigen::Vector3f(a_ * m[0] + twx_, a_ * m[1] + twy_, m[2]); }
  Eigen::Vector3f getMapCoordsPose(const Eigen::Vector3f& w) const { return Eigen::Vector3f(scale_ * w[0] + tmx_, scale_ * w[1] + tmy_, w[2]); }

 private:
  friend class HectorSlamProcessor;   // fills the mirror (also for hsgpu/MapRepGpu.h, which derives from it)
  Eigen::Vector2i dims_;
  float cell_length_, scale_, tmx_, tmy_, a_, twx_, twy_;
  int update_index_;
  int mirrored_index_;
  std::vector<hs_cell> cells_;
};

class HectorSlamProcessor {
 public:
  HectorSlamProcessor(float mapResolution, int mapSizeX, int mapSizeY, const Eigen::Vector2f& startCoords, int multi_res_size,
                      DrawInterface* drawInterfaceIn = 0, HectorDebugInfoInterface* debugInterfaceIn = 0)
      : engine_(0), drawInterface(drawInterfaceIn), debugInterface(debugInterfaceIn) {
    hs_config cfg;
    cfg.map_resolution = mapResolution;
    cfg.map_size_x = mapSizeX;
    cfg.map_size_y = mapSizeY;
    cfg.start_x = startCoords[0];
    cfg.start_y = startCoords[1];
    cfg.levels = multi_res_size;
    cfg.n_streams = 1;
    cfg.max_points = HSGPU_FACADE_MAX_POINTS;
    const char* dev = std::getenv("HSGPU_DEVICE");
    cfg.device = dev ? std::atoi(dev) : 0;
    int st = hs_create(&cfg, &engine_);
    if (st != HS_OK) throw std::runtime_error(std::string("hsgpu: hs_create failed: ") + hs_status_string(st));
    levels_.resize(multi_res_size);
    mutexes_.assign(multi_res_size, (MapLockerInterface*)0);
    // level geometry with the reference's arithmetic (MapRepMultiMap.h:48-72, GridMapBase.h:265-280)
    float res = mapResolution;
    const float offx = (mapResolution * static_cast<float>(mapSizeX)) * startCoords[0];
    const float offy = (mapResolution * static_cast<float>(mapSizeY)) * startCoords[1];
    for (int i = 0; i < multi_res_size; ++i) {
      GridMap& g = levels_[i];
      int sx, sy;
      float cl;
      hs_get_level_dims(engine_, i, &sx, &sy, &cl);
      g.dims_ = Eigen::Vector2i(sx, sy);
      g.cell_length_ = res;
      g.scale_ = 1.0f / res;
      g.tmx_ = g.scale_ * offx;
      g.tmy_ = g.scale_ * offy;
      const float invdet = 1.0f / (g.scale_ * g.scale_);
      g.a_ = g.scale_ * invdet;
      g.twx_ = -(g.a_ * g.tmx_);
      g.twy_ = -(g.a_ * g.tmy_);
      g.update_index_ = -1;
      g.mirrored_index_ = -2;
      hs_cell blank;
      blank.logOddsVal = 0.0f;
      blank.updateIndex = -1;
      g.cells_.assign((size_t)sx * sy, blank);
      res *= 2.0f;
    }
    lastScanMatchPose = Eigen::Vector3f::Zero();
    lastScanMatchCov = Eigen::Matrix3f::Zero();
    this->reset();
    this->setMapUpdateMinDistDiff(0.4f * 1.0f);
    this->setMapUpdateMinAngleDiff(0.13f * 1.0f);
  }

  ~HectorSlamProcessor() {
    hs_destroy(engine_);
    for (size_t i = 0; i < mutexes_.size(); ++i) delete mutexes_[i];  // MapProcContainer::cleanup owns the mutex (MapProcContainer.h:56-65)
  }

  void update(const DataContainer& dataContainer, const Eigen::Vector3f& poseHintWorld, bool map_without_matching = false) {
    const int n = dataContainer.getSize();
    if (n > HSGPU_FACADE_MAX_POINTS) throw std::length_error("hsgpu: DataContainer larger than HSGPU_FACADE_MAX_POINTS");
    if (points_.size() < (size_t)2 * HSGPU_FACADE_MAX_POINTS) points_.resize((size_t)2 * HSGPU_FACADE_MAX_POINTS);
    for (int i = 0; i < n; ++i) {
      const Eigen::Vector2f& p = dataContainer.getVecEntry(i);
      points_[2 * i] = p[0];
      points_[2 * i + 1] = p[1];
    }
    const Eigen::Vector2f o = dataContainer.getOrigo();
    const float origo[2] = {o[0], o[1]};
    const float hint[3] = {poseHintWorld[0], poseHintWorld[1], poseHintWorld[2]};
    const int32_t count = n;
    const uint8_t flag = map_without_matching ? 1 : 0;
    float pose[3], cov[9];
    {
      std::lock_guard<std::mutex> guard(engine_mutex_);
      for (size_t i = 0; i < mutexes_.size(); ++i) if (mutexes_[i]) mutexes_[i]->lockMap();
      int st = hs_update_batch(engine_, 1, 0, points_.data(), &count, origo, hint, &flag, pose, cov);
      for (size_t i = 0; i < mutexes_.size(); ++i) if (mutexes_[i]) mutexes_[i]->unlockMap();
      if (st != HS_OK) throw std::runtime_error(std::string("hsgpu: hs_update_batch failed: ") + hs_status_string(st));
      int ui = -1;
      hs_get_update_index(engine_, 0, 0, &ui);
      for (size_t i = 0; i < levels_.size(); ++i) levels_[i].update_index_ = ui;
    }
    lastScanMatchPose = Eigen::Vector3f(pose[0], pose[1], pose[2]);
    if (!map_without_matching && n != 0)
      for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) lastScanMatchCov(i, j) = cov[3 * i + j];

    if (drawInterface) {  // HectorSlamProcessor.h:96-107
      const GridMap& gridMapRef = levels_[0];
      drawInterface->setColor(1.0, 0.0, 0.0);
      drawInterface->setScale(0.15);
      drawInterface->drawPoint(gridMapRef.getWorldCoords(Eigen::Vector2f::Zero()));
      drawInterface->drawPoint(gridMapRef.getWorldCoords(Eigen::Vector2f((float)(gridMapRef.getSizeX() - 1), (float)(gridMapRef.getSizeY() - 1))));
      drawInterface->drawPoint(Eigen::Vector2f(1.0f, 1.0f));
      drawInterface->sendAndResetData();
    }
    if (debugInterface) debugInterface->sendAndResetData();
  }

  void reset() {
    std::lock_guard<std::mutex> guard(engine_mutex_);
    lastScanMatchPose = Eigen::Vector3f::Zero();
    hs_reset(engine_, HS_ALL_STREAMS);
    for (size_t i = 0; i < levels_.size(); ++i) levels_[i].mirrored_index_ = -2;  // force a refresh: cells were cleared
  }

  const Eigen::Vector3f& getLastScanMatchPose() const { return lastScanMatchPose; }
  const Eigen::Matrix3f& getLastScanMatchCovariance() const { return lastScanMatchCov; }
  float getScaleToMap() const { return hs_get_scale_to_map(engine_); }
  int getMapLevels() const { return hs_get_map_levels(engine_); }

  // The returned reference stays valid for the processor's lifetime; its cells are refreshed from the GPU here
  // when the level's update index changed since the last call (the reference's caller reads cells only after
  // checking getUpdateIndex(), HectorMappingRos.cpp:440).
  const GridMap& getGridMap(int mapLevel = 0) const {
    GridMap& g = levels_[mapLevel];
    std::lock_guard<std::mutex> guard(engine_mutex_);
    if (g.mirrored_index_ != g.update_index_) {
      if (mutexes_[mapLevel]) mutexes_[mapLevel]->lockMap();
      hs_download_level(engine_, 0, mapLevel, g.cells_.data());
      if (mutexes_[mapLevel]) mutexes_[mapLevel]->unlockMap();
      g.mirrored_index_ = g.update_index_;
    }
    return g;
  }

  void addMapMutex(int i, MapLockerInterface* mapMutex) {
    delete mutexes_[i];
    mutexes_[i] = mapMutex;
  }
  MapLockerInterface* getMapMutex(int i) { return mutexes_[i]; }

  void setUpdateFactorFree(float free_factor) { hs_set_update_factor_free(engine_, free_factor); }
  void setUpdateFactorOccupied(float occupied_factor) { hs_set_update_factor_occupied(engine_, occupied_factor); }
  void setMapUpdateMinDistDiff(float minDist) { hs_set_map_update_min_dist_diff(engine_, minDist); }
  void setMapUpdateMinAngleDiff(float angleChange) { hs_set_map_update_min_angle_diff(engine_, angleChange); }

  hs_engine* engine() { return engine_; }  // escape hatch to the C ABI (statistics, device variants)

 protected:
  void setLevelUpdateIndex(size_t level, int updateIndex) { levels_[level].update_index_ = updateIndex; }   // for hsgpu/MapRepGpu.h

  hs_engine* engine_;
  mutable std::mutex engine_mutex_;
  mutable std::vector<GridMap> levels_;
  std::vector<MapLockerInterface*> mutexes_;
  std::vector<float> points_;

  Eigen::Vector3f lastScanMatchPose;
  Eigen::Matrix3f lastScanMatchCov;

  DrawInterface* drawInterface;
  HectorDebugInfoInterface* debugInterface;

 private:
  HectorSlamProcessor(const HectorSlamProcessor&);
  HectorSlamProcessor& operator=(const HectorSlamProcessor&);
};

}  // namespace hectorslam

#endif
