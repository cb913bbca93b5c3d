// Drives hsgpu::BatchEngine (include/hsgpu/BatchEngine.h): n streams x k scans through update(), dumps the poses of every scan
// and the final level-0 cells of the first and last stream for comparison with the CPU oracle (tests/test_facade_cpp.py).
// usage: batch_test <in.bin> <out.bin>
// in : int32 n_streams, int32 n_scans, int32 max_pts, then per scan: int32 counts[n_streams], float points[n_streams][max_pts][2]
// out: per scan float poses[n_streams][3]; then int32 updateIndex + cells of level 0 for stream 0 and stream n_streams-1
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "hsgpu/BatchEngine.h"

int main(int argc, char** argv) {
  if (argc < 3) return 2;
  FILE* fi = std::fopen(argv[1], "rb");
  FILE* fo = std::fopen(argv[2], "wb");
  if (!fi || !fo) return 3;
  int hdr[3];
  if (std::fread(hdr, 4, 3, fi) != 3) return 4;
  const int n = hdr[0], n_scans = hdr[1], max_pts = hdr[2];
  try {
    hsgpu::BatchEngine eng(0.05f, 1024, 1024, 0.5f, 0.5f, 3, n, max_pts);
    eng.setUpdateFactorFree(0.4f);
    eng.setUpdateFactorOccupied(0.9f);
    eng.setMapUpdateMinDistDiff(0.0f);
    eng.setMapUpdateMinAngleDiff(0.0f);
    if (eng.getMapLevels() != 3 || eng.getScaleToMap() != 20.0f || eng.streams() != n || eng.maxPoints() != max_pts) return 5;
    std::vector<int32_t> counts((size_t)n);
    std::vector<float> pts((size_t)n * max_pts * 2), poses((size_t)n * 3);
    for (int k = 0; k < n_scans; ++k) {
      if (std::fread(counts.data(), 4, counts.size(), fi) != counts.size() || std::fread(pts.data(), 4, pts.size(), fi) != pts.size()) return 6;
      if (k % 2 == 0) {
        eng.update(n, pts.data(), counts.data(), poses.data());
      } else {   // pageable buffers: submit() degrades to the synchronous call
        eng.submit(n, pts.data(), counts.data(), poses.data());
        eng.wait();
      }
      std::fwrite(poses.data(), 4, poses.size(), fo);
      std::vector<float> again = eng.lastScanMatchPoses(0, n);
      for (size_t i = 0; i < again.size(); ++i)
        if (again[i] != poses[i]) return 7;
    }
    for (int s : {0, n - 1}) {
      int ui = eng.updateIndex(s, 0);
      std::vector<hs_cell> cells = eng.gridMap(s, 0);
      std::fwrite(&ui, 4, 1, fo);
      std::fwrite(cells.data(), sizeof(hs_cell), cells.size(), fo);
    }
    bool threw = false;
    try { eng.gridMap(n, 0); } catch (const hsgpu::Error& err) { threw = err.status() == HS_ERR_INVALID_ARG; }
    if (!threw) return 8;
  } catch (const hsgpu::Error& err) {
    std::fprintf(stderr, "hsgpu error: %s\n", err.what());
    return 1;
  }
  std::fclose(fi);
  std::fclose(fo);
  return 0;
}
