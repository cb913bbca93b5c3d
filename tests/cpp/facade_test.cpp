// Drives the drop-in hectorslam::HectorSlamProcessor facade the way HectorMappingRos::scanCallback does
// (HectorMappingRos.cpp:253-254: fill a DataContainer, update(dc, getLastScanMatchPose())), and dumps poses,
// covariances and the final level cells for comparison with the CPU oracle (tests/test_facade_cpp.py).
// usage: facade_test <in.bin> <out.bin>
// in : int32 n_scans, int32 max_pts, then per scan: int32 count, float[max_pts*2]
// out: per scan float[3] pose + float[9] cov, then per level: int32 sx, int32 sy, int32 updateIndex, cells
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "hsgpu/HectorSlamProcessor.h"

struct CountingLock : MapLockerInterface {
  int locks = 0, unlocks = 0;
  void lockMap() override { ++locks; }
  void unlockMap() override { ++unlocks; }
};

int main(int argc, char** argv) {
  if (argc < 3) return 2;
  FILE* fi = std::fopen(argv[1], "rb");
  FILE* fo = std::fopen(argv[2], "wb");
  if (!fi || !fo) return 3;
  int n_scans = 0, max_pts = 0;
  if (std::fread(&n_scans, 4, 1, fi) != 1 || std::fread(&max_pts, 4, 1, fi) != 1) return 4;
  hectorslam::HectorSlamProcessor proc(0.05f, 1024, 1024, Eigen::Vector2f(0.5f, 0.5f), 3);
  proc.setUpdateFactorFree(0.4f);
  proc.setUpdateFactorOccupied(0.9f);
  proc.setMapUpdateMinDistDiff(0.4f);
  proc.setMapUpdateMinAngleDiff(0.9f);
  CountingLock* lock = new CountingLock();  // ownership passes to the processor, as in the reference
  proc.addMapMutex(0, lock);
  if (proc.getMapLevels() != 3 || proc.getScaleToMap() != 20.0f || proc.getMapMutex(0) != lock) return 5;
  std::vector<float> buf((size_t)max_pts * 2);
  hectorslam::DataContainer dc;
  for (int k = 0; k < n_scans; ++k) {
    int count = 0;
    if (std::fread(&count, 4, 1, fi) != 1 || std::fread(buf.data(), 4, buf.size(), fi) != buf.size()) return 6;
    dc.clear();
    dc.setOrigo(Eigen::Vector2f::Zero());
    for (int i = 0; i < count; ++i) dc.add(Eigen::Vector2f(buf[2 * i], buf[2 * i + 1]));
    if (k == n_scans / 2) {  // reset in the middle, like the node's reset service
      proc.reset();
    }
    proc.update(dc, proc.getLastScanMatchPose());
    const Eigen::Vector3f& p = proc.getLastScanMatchPose();
    const Eigen::Matrix3f& c = proc.getLastScanMatchCovariance();
    float row[12] = {p[0], p[1], p[2], c(0, 0), c(0, 1), c(0, 2), c(1, 0), c(1, 1), c(1, 2), c(2, 0), c(2, 1), c(2, 2)};
    std::fwrite(row, 4, 12, fo);
  }
  for (int l = 0; l < proc.getMapLevels(); ++l) {
    const hectorslam::GridMap& g = proc.getGridMap(l);
    int hdr[3] = {g.getSizeX(), g.getSizeY(), g.getUpdateIndex()};
    std::fwrite(hdr, 4, 3, fo);
    int n = g.getSizeX() * g.getSizeY();
    int free_cells = 0, occ_cells = 0;
    for (int i = 0; i < n; ++i) { free_cells += g.isFree(i); occ_cells += g.isOccupied(i); }
    std::fwrite(&g.getCell(0), sizeof(hs_cell), (size_t)n, fo);
    std::fprintf(stderr, "level %d: %dx%d updateIndex %d free %d occupied %d\n", l, hdr[0], hdr[1], hdr[2], free_cells, occ_cells);
  }
  Eigen::Vector2f w = proc.getGridMap(0).getWorldCoords(Eigen::Vector2f(512.0f, 512.0f));
  std::fprintf(stderr, "world(512,512) = %g %g; locks %d unlocks %d\n", w[0], w[1], lock->locks, lock->unlocks);
  if (lock->locks != lock->unlocks || lock->locks < n_scans) return 7;
  std::fclose(fi);
  std::fclose(fo);
  return 0;
}
