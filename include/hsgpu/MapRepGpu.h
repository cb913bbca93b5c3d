// hectorslam::MapRepGpu -- the reference's seam one level BELOW HectorSlamProcessor, backed by the hsgpu CUDA engine.
//
// The reference splits its processor into HectorSlamProcessor (gate, bookkeeping) and a MapRepresentationInterface
// (slam_main/MapRepresentationInterface.h:38-62) whose only implementation is MapRepMultiMap (MapRepMultiMap.h:44-171).
// A caller that keeps the reference's HectorSlamProcessor logic -- or its own -- and only wants the map representation
// on the GPU constructs a MapRepGpu where the reference constructs a MapRepMultiMap (HectorSlamProcessor.h:57):
//
//   mapRep = new MapRepGpu(mapResolution, mapSizeX, mapSizeY, multi_res_size, startCoords, drawInterfaceIn, debugInterfaceIn);
//
//   matchData     -> hs_match_batch   (MapRepMultiMap::matchData, coarse to fine; refreshes the coarse containers)
//   updateByScan  -> hs_integrate_batch (MapRepMultiMap::updateByScan: level 0 from the given scan, the coarse levels from
//                    the containers the last matchData left behind, MapRepMultiMap.h:134-147)
//   onMapUpdated  -> nothing to do    (there is no GridMapCacheArray to invalidate)
//   getGridMap    -> lazily refreshed host mirror, as in hsgpu/HectorSlamProcessor.h
//
// As in the reference, updateByScan of a scan that was not the last one matched integrates STALE (or, before the first
// matchData, empty) containers on the coarse levels -- e.g. HectorSlamProcessor::update(..., map_without_matching = true)
// (tests/cpp/maprep_test.cpp drives exactly that against the CPU reference).
// Needs Eigen exactly as the reference header does. No CPU fallback: the constructor throws without a CUDA device.
#ifndef HSGPU_MAPREPGPU_H
#define HSGPU_MAPREPGPU_H

#include "HectorSlamProcessor.h"

namespace hectorslam {

// slam_main/MapRepresentationInterface.h:38-62 (same declarations, same include guard)
#ifndef _hectormaprepresentationinterface_h__
#define _hectormaprepresentationinterface_h__
class MapRepresentationInterface {
 public:
  virtual ~MapRepresentationInterface() {}
  virtual void reset() = 0;
  virtual float getScaleToMap() const = 0;
  virtual int getMapLevels() const = 0;
  virtual const GridMap& getGridMap(int mapLevel = 0) const = 0;
  virtual void addMapMutex(int i, MapLockerInterface* mapMutex) = 0;
  virtual MapLockerInterface* getMapMutex(int i) = 0;
  virtual void onMapUpdated() = 0;
  virtual Eigen::Vector3f matchData(const Eigen::Vector3f& beginEstimateWorld, const DataContainer& dataContainer, Eigen::Matrix3f& covMatrix) = 0;
  virtual void updateByScan(const DataContainer& dataContainer, const Eigen::Vector3f& robotPoseWorld) = 0;
  virtual void setUpdateFactorFree(float free_factor) = 0;
  virtual void setUpdateFactorOccupied(float occupied_factor) = 0;
};
#endif

class MapRepGpu : public MapRepresentationInterface {
 public:
  // MapRepMultiMap(float, int, int, unsigned int numDepth, const Vector2f& startCoords, DrawInterface*, HectorDebugInfoInterface*)
  // (MapRepMultiMap.h:47)
  MapRepGpu(float mapResolution, int mapSizeX, int mapSizeY, unsigned int numDepth, const Eigen::Vector2f& startCoords,
            DrawInterface* drawInterfaceIn = 0, HectorDebugInfoInterface* debugInterfaceIn = 0)
      : proc_(mapResolution, mapSizeX, mapSizeY, startCoords, (int)numDepth, drawInterfaceIn, debugInterfaceIn) {}

  virtual void reset() { proc_.reset(); }
  virtual float getScaleToMap() const { return proc_.getScaleToMap(); }
  virtual int getMapLevels() const { return proc_.getMapLevels(); }
  virtual const GridMap& getGridMap(int mapLevel = 0) const { return proc_.getGridMap(mapLevel); }
  virtual void addMapMutex(int i, MapLockerInterface* mapMutex) { proc_.addMapMutex(i, mapMutex); }
  virtual MapLockerInterface* getMapMutex(int i) { return proc_.getMapMutex(i); }
  virtual void onMapUpdated() {}

  virtual Eigen::Vector3f matchData(const Eigen::Vector3f& beginEstimateWorld, const DataContainer& dataContainer, Eigen::Matrix3f& covMatrix) {
    return proc_.matchOnly(beginEstimateWorld, dataContainer, covMatrix);
  }
  virtual void updateByScan(const DataContainer& dataContainer, const Eigen::Vector3f& robotPoseWorld) {
    proc_.integrateOnly(dataContainer, robotPoseWorld);
  }
  virtual void setUpdateFactorFree(float free_factor) { proc_.setUpdateFactorFree(free_factor); }
  virtual void setUpdateFactorOccupied(float occupied_factor) { proc_.setUpdateFactorOccupied(occupied_factor); }

  hs_engine* engine() { return proc_.engine(); }

 private:
  // the processor façade owns the engine, the level mirrors and the map mutexes; this class adds the two split entry points
  struct Proc : HectorSlamProcessor {
    Proc(float res, int sx, int sy, const Eigen::Vector2f& start, int levels, DrawInterface* d, HectorDebugInfoInterface* g)
        : HectorSlamProcessor(res, sx, sy, start, levels, d, g) {}

    void pack(const DataContainer& dc, float origo[2], int32_t& count) {
      const int n = dc.getSize();
      if (n > HSGPU_FACADE_MAX_POINTS) throw std::length_error("hsgpu: DataContainer larger than HSGPU_FACADE_MAX_POINTS");
      if (points_.size() < (size_t)2 * HSGPU_FACADE_MAX_POINTS) points_.resize((size_t)2 * HSGPU_FACADE_MAX_POINTS);
      for (int i = 0; i < n; ++i) {
        const Eigen::Vector2f& p = dc.getVecEntry(i);
        points_[2 * i] = p[0];
        points_[2 * i + 1] = p[1];
      }
      const Eigen::Vector2f o = dc.getOrigo();
      origo[0] = o[0];
      origo[1] = o[1];
      count = n;
    }

    // MapRepMultiMap::matchData (MapRepMultiMap.h:116-132); an empty scan returns the estimate and leaves covMatrix alone
    // (ScanMatcher.h:56, :189)
    Eigen::Vector3f matchOnly(const Eigen::Vector3f& beginEstimateWorld, const DataContainer& dc, Eigen::Matrix3f& covMatrix) {
      float origo[2], pose[3], cov[9];
      int32_t count;
      pack(dc, origo, count);
      const float hint[3] = {beginEstimateWorld[0], beginEstimateWorld[1], beginEstimateWorld[2]};
      std::lock_guard<std::mutex> guard(engine_mutex_);
      int st = hs_match_batch(engine_, 1, 0, points_.data(), &count, origo, hint, pose, cov);
      if (st != HS_OK) throw std::runtime_error(std::string("hsgpu: hs_match_batch failed: ") + hs_status_string(st));
      if (count != 0)
        for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) covMatrix(i, j) = cov[3 * i + j];
      return Eigen::Vector3f(pose[0], pose[1], pose[2]);
    }

    // MapRepMultiMap::updateByScan (MapRepMultiMap.h:134-147) under the map mutexes, as HectorSlamProcessor::update takes them
    void integrateOnly(const DataContainer& dc, const Eigen::Vector3f& robotPoseWorld) {
      float origo[2];
      int32_t count;
      pack(dc, origo, count);
      const float pose[3] = {robotPoseWorld[0], robotPoseWorld[1], robotPoseWorld[2]};
      std::lock_guard<std::mutex> guard(engine_mutex_);
      for (size_t i = 0; i < mutexes_.size(); ++i) if (mutexes_[i]) mutexes_[i]->lockMap();
      int st = hs_integrate_batch(engine_, 1, 0, points_.data(), &count, origo, pose);
      for (size_t i = 0; i < mutexes_.size(); ++i) if (mutexes_[i]) mutexes_[i]->unlockMap();
      if (st != HS_OK) throw std::runtime_error(std::string("hsgpu: hs_integrate_batch failed: ") + hs_status_string(st));
      int ui = -1;
      hs_get_update_index(engine_, 0, 0, &ui);
      for (size_t i = 0; i < levels_.size(); ++i) setLevelUpdateIndex(i, ui);
    }
  };
  mutable Proc proc_;
};

}  // namespace hectorslam

#endif
