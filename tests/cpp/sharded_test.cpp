// Drives hsgpu::ShardedEngine (include/hsgpu/ShardedEngine.h) next to one hsgpu::BatchEngine holding the same streams:
// every scan's poses, sampled final maps, and the sharded hypothesis entry points must
// equal the single-engine results bit for bit. Shard d runs on device d mod (visible devices), so `n_shards` > 1 exercises the
// fan-out, the concatenation and the guest-stream replication even on a single GPU; on a multi-GPU box it spans the GPUs.
// usage: sharded_test <in.bin> <out.bin> [n_shards (0 = one per visible device)]
// in : int32 n_streams, int32 n_scans, int32 max_pts, then per scan: int32 counts[n_streams], float points[n_streams][max_pts][2]
// out: per scan float poses[n_streams][3] (of the sharded engine)
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "hsgpu/ShardedEngine.h"

static bool same_bits(const float* a, const float* b, size_t n) { return std::memcmp(a, b, n * sizeof(float)) == 0; }

int main(int argc, char** argv) {
  if (argc < 3) return 2;
  FILE* fi = std::fopen(argv[1], "rb");
  FILE* fo = std::fopen(argv[2], "wb");
  if (!fi || !fo) return 3;
  const int n_shards = argc > 3 ? std::atoi(argv[3]) : 0;
  int hdr[3];
  if (std::fread(hdr, 4, 3, fi) != 3) return 4;
  const int n = hdr[0], n_scans = hdr[1], max_pts = hdr[2];
  try {
    hsgpu::ShardedEngine sh(0.05f, 1024, 1024, 0.5f, 0.5f, 3, n, max_pts, n_shards);
    hsgpu::BatchEngine one(0.05f, 1024, 1024, 0.5f, 0.5f, 3, n, max_pts);
    sh.setUpdateFactorFree(0.4f); sh.setUpdateFactorOccupied(0.9f); sh.setMapUpdateMinDistDiff(0.0f); sh.setMapUpdateMinAngleDiff(0.0f);
    one.setUpdateFactorFree(0.4f); one.setUpdateFactorOccupied(0.9f); one.setMapUpdateMinDistDiff(0.0f); one.setMapUpdateMinAngleDiff(0.0f);
    // the partition covers every stream exactly once
    int covered = 0;
    for (int d = 0; d < sh.shards(); ++d) {
      int first, count;
      hsgpu::shard_range(n, sh.shards(), d, first, count);
      if (first != covered) return 5;
      covered += count;
    }
    if (covered != n || sh.shardOf(0) != 0 || sh.shardOf(n - 1) != sh.shards() - 1 || sh.shardOf(n) != -1) return 5;

    std::vector<int32_t> counts((size_t)n);
    std::vector<float> pts((size_t)n * max_pts * 2), poses((size_t)n * 3), cov((size_t)n * 9), poses1((size_t)n * 3), cov1((size_t)n * 9);
    std::vector<float> last_pts;
    int last_cnt = 0;
    for (int k = 0; k < n_scans; ++k) {
      if (std::fread(counts.data(), 4, counts.size(), fi) != counts.size() || std::fread(pts.data(), 4, pts.size(), fi) != pts.size()) return 6;
      sh.update(pts.data(), counts.data(), poses.data(), cov.data());
      one.update(n, pts.data(), counts.data(), poses1.data(), cov1.data());
      if (!same_bits(poses.data(), poses1.data(), poses.size())) { std::fprintf(stderr, "poses differ at scan %d\n", k); return 7; }
      if (k > 0 && !same_bits(cov.data(), cov1.data(), cov.size())) { std::fprintf(stderr, "covariances differ at scan %d\n", k); return 7; }
      std::fwrite(poses.data(), 4, poses.size(), fo);
      const int s = n / 2;
      last_cnt = counts[s];
      last_pts.assign(pts.begin() + (size_t)s * max_pts * 2, pts.begin() + (size_t)s * max_pts * 2 + (size_t)last_cnt * 2);
    }
    std::vector<float> all = sh.lastScanMatchPoses();
    if (!same_bits(all.data(), poses.data(), all.size())) return 8;
    for (int s : {0, n / 2, n - 1})
      for (int l = 0; l < 3; ++l) {
        std::vector<hs_cell> a = sh.gridMap(s, l), b = one.gridMap(s, l);
        if (a.size() != b.size() || std::memcmp(a.data(), b.data(), a.size() * sizeof(hs_cell)) != 0) { std::fprintf(stderr, "map %d/%d differs\n", s, l); return 9; }
      }

    // hypotheses of stream n/2, sharded vs one engine; twice, with an update in between (the replicas must follow the maps)
    const int s = n / 2, n_hyp = 1031;     // deliberately not divisible by the shard count
    for (int round = 0; round < 2; ++round) {
      std::vector<float> hints((size_t)n_hyp * 3), pm((size_t)n_hyp * 3);
      unsigned lcg = 12345u + round;
      for (int i = 0; i < n_hyp; ++i)
        for (int c = 0; c < 3; ++c) {
          lcg = lcg * 1664525u + 1013904223u;
          const float u = (float)((lcg >> 8) & 0xFFFF) / 65535.0f * 2.0f - 1.0f;
          hints[3 * i + c] = poses[3 * s + c] + u * (c == 2 ? 0.3f : 0.5f);
          pm[3 * i + c] = c == 2 ? hints[3 * i + c] : 20.0f * hints[3 * i + c] + 512.0f;   // level-0 map coordinates
        }
      std::vector<float> H((size_t)n_hyp * 9), d((size_t)n_hyp * 3), H1((size_t)n_hyp * 9), d1((size_t)n_hyp * 3);
      sh.hessianDerivsSharded(s, 0, pm.data(), n_hyp, last_pts.data(), last_cnt, H.data(), d.data());
      int st = hs_hessian_derivs_batch(one.handle(), s, 0, pm.data(), n_hyp, last_pts.data(), last_cnt, H1.data(), d1.data());
      if (st != HS_OK) return 10;
      if (!same_bits(H.data(), H1.data(), H.size()) || !same_bits(d.data(), d1.data(), d.size())) { std::fprintf(stderr, "sharded H/dTr differ (round %d)\n", round); return 10; }
      std::vector<float> mp((size_t)n_hyp * 3), mc((size_t)n_hyp * 9), mp1((size_t)n_hyp * 3), mc1((size_t)n_hyp * 9);
      sh.matchHypothesesSharded(s, last_pts.data(), last_cnt, hints.data(), n_hyp, mp.data(), mc.data());
      one.matchHypotheses(s, last_pts.data(), last_cnt, hints.data(), n_hyp, mp1.data(), mc1.data());
      if (!same_bits(mp.data(), mp1.data(), mp.size()) || !same_bits(mc.data(), mc1.data(), mc.size())) { std::fprintf(stderr, "sharded multi-start match differs (round %d)\n", round); return 11; }
      if (round == 0) {   // change the maps: the cached replicas are now stale
        sh.update(pts.data(), counts.data(), poses.data(), nullptr);
        one.update(n, pts.data(), counts.data(), poses1.data(), nullptr);
        if (!same_bits(poses.data(), poses1.data(), poses.size())) return 12;
      }
    }
    bool threw = false;
    try { sh.gridMap(n, 0); } catch (const hsgpu::Error& err) { threw = err.status() == HS_ERR_INVALID_ARG; }
    if (!threw) return 13;
    std::fprintf(stderr, "sharded ok: %d streams over %d shard(s) on %d device(s), %d scans, %d hypotheses\n", n, sh.shards(), hs_device_count(), n_scans, n_hyp);
  } catch (const hsgpu::Error& err) {
    std::fprintf(stderr, "hsgpu error: %s\n", err.what());
    return 1;
  }
  std::fclose(fi);
  std::fclose(fo);
  return 0;
}
