// Drives hectorslam::MapRepGpu (include/hsgpu/MapRepGpu.h) through the reference's MapRepresentationInterface the way the
// reference's HectorSlamProcessor::update drives MapRepMultiMap (HectorSlamProcessor.h:71-95): matchData from the last pose,
// the poseDifferenceLargerThan gate, updateByScan + onMapUpdated -- the gate and the bookkeeping are host code here, as in
// the reference. Dumps poses, covariances and the final level cells for comparison with the CPU oracle.
// usage: maprep_test <in.bin> <out.bin> [mwm_period]   (formats as facade_test.cpp; gate thresholds 0.4 m / 0.9 rad)
// mwm_period P > 0: every scan k with k % P == P - 1 takes update()'s map_without_matching = true branch
// (HectorSlamProcessor.h:78-82, :89): no matchData, the pose hint becomes the pose, updateByScan runs unconditionally -- so
// the coarse levels integrate the containers the LAST matched scan left behind (MapRepMultiMap.h:134-147).
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <limits>
#include <vector>

#include "hsgpu/MapRepGpu.h"

// util::poseDifferenceLargerThan (util/UtilFunctions.h:73-92): float norm, the angle wrapped in double, float fabs
static bool pose_difference_larger_than(const Eigen::Vector3f& pose1, const Eigen::Vector3f& pose2, float distanceDiffThresh, float angleDiffThresh) {
  const float dx = pose1[0] - pose2[0], dy = pose1[1] - pose2[1];
  if (std::sqrt(dx * dx + dy * dy) > distanceDiffThresh) return true;
  float angleDiff = pose1[2] - pose2[2];
  const double pi = 3.14159265358979323846;
  if ((double)angleDiff > pi) angleDiff = (float)((double)angleDiff - pi * 2.0f);
  else if ((double)angleDiff < -pi) angleDiff = (float)((double)angleDiff + pi * 2.0f);
  return std::fabs(angleDiff) > angleDiffThresh;
}

int main(int argc, char** argv) {
  if (argc < 3) return 2;
  FILE* fi = std::fopen(argv[1], "rb");
  FILE* fo = std::fopen(argv[2], "wb");
  if (!fi || !fo) return 3;
  const int mwm_period = argc > 3 ? std::atoi(argv[3]) : 0;
  int n_scans = 0, max_pts = 0;
  if (std::fread(&n_scans, 4, 1, fi) != 1 || std::fread(&max_pts, 4, 1, fi) != 1) return 4;
  hectorslam::MapRepresentationInterface* mapRep = new hectorslam::MapRepGpu(0.05f, 1024, 1024, 3, Eigen::Vector2f(0.5f, 0.5f), 0, 0);
  mapRep->setUpdateFactorFree(0.4f);
  mapRep->setUpdateFactorOccupied(0.9f);
  if (mapRep->getMapLevels() != 3 || mapRep->getScaleToMap() != 20.0f) return 5;
  const float fmax = std::numeric_limits<float>::max();
  Eigen::Vector3f lastScanMatchPose(0.0f, 0.0f, 0.0f), lastMapUpdatePose(fmax, fmax, fmax);   // HectorSlamProcessor::reset (:115-124)
  Eigen::Matrix3f lastScanMatchCov;
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) lastScanMatchCov(i, j) = 0.0f;
  std::vector<float> buf((size_t)max_pts * 2);
  hectorslam::DataContainer dc;
  int integrations = 0;
  for (int k = 0; k < n_scans; ++k) {
    int count = 0;
    if (std::fread(&count, 4, 1, fi) != 1 || std::fread(buf.data(), 4, buf.size(), fi) != buf.size()) return 6;
    dc.clear();
    dc.setOrigo(Eigen::Vector2f::Zero());
    for (int i = 0; i < count; ++i) dc.add(Eigen::Vector2f(buf[2 * i], buf[2 * i + 1]));
    const bool map_without_matching = mwm_period > 0 && (k % mwm_period) == mwm_period - 1;
    const Eigen::Vector3f poseHintWorld = lastScanMatchPose;
    Eigen::Vector3f newPoseEstimateWorld;
    if (!map_without_matching) newPoseEstimateWorld = mapRep->matchData(poseHintWorld, dc, lastScanMatchCov);
    else newPoseEstimateWorld = poseHintWorld;
    lastScanMatchPose = newPoseEstimateWorld;
    if (pose_difference_larger_than(newPoseEstimateWorld, lastMapUpdatePose, 0.4f, 0.9f) || map_without_matching) {
      mapRep->updateByScan(dc, newPoseEstimateWorld);
      mapRep->onMapUpdated();
      lastMapUpdatePose = newPoseEstimateWorld;
      ++integrations;
    }
    const Eigen::Vector3f& p = lastScanMatchPose;
    const Eigen::Matrix3f& c = lastScanMatchCov;
    float row[12] = {p[0], p[1], p[2], c(0, 0), c(0, 1), c(0, 2), c(1, 0), c(1, 1), c(1, 2), c(2, 0), c(2, 1), c(2, 2)};
    std::fwrite(row, 4, 12, fo);
  }
  for (int l = 0; l < mapRep->getMapLevels(); ++l) {
    const hectorslam::GridMap& g = mapRep->getGridMap(l);
    int hdr[3] = {g.getSizeX(), g.getSizeY(), g.getUpdateIndex()};
    std::fwrite(hdr, 4, 3, fo);
    std::fwrite(&g.getCell(0), sizeof(hs_cell), (size_t)g.getSizeX() * g.getSizeY(), fo);
  }
  std::fprintf(stderr, "%d scans, %d integrated\n", n_scans, integrations);
  delete mapRep;
  std::fclose(fi);
  std::fclose(fo);
  return 0;
}
