// hsgpu::ShardedEngine -- the multi-GPU host layer for a single C++ process.
//
// north_star: "Work is partitioned across the 8 GPUs of one box by sharding the independent streams. No NCCL is needed on the
// hot path because there is no cross-stream reduction." Streams share nothing (own grids, pose, update counters), so global
// stream g lives on exactly one shard; one hs_engine per shard, one persistent host thread per shard (its own CUDA context
// binding: every hs_* call selects its engine's device), update() fans the slot ranges out and the matched poses come back
// by plain concatenation into the caller's arrays. Shard d runs on device (d mod visible devices), so the same code path can
// be exercised with several shards on one GPU.
//
// The one mode with a real exchange step is BASELINE configs[3] sharded: ONE stream's maps, thousands of
// pose hypotheses. The owner's level grids are replicated into a guest stream of every other shard (hs_copy_stream_maps:
// cudaMemcpyPeerAsync over NVLink, 10.5 MiB for 3 levels of 1024^2, cached until the owner's maps change), every shard
// evaluates a contiguous block of the hypotheses, and the n x 12 floats land in the caller's arrays.
//
// Needs only the C ABI (hsgpu.h) and the C++11 standard library. No CPU fallback: construction throws hsgpu::Error when no
// CUDA device is usable.
#ifndef HSGPU_SHARDEDENGINE_H
#define HSGPU_SHARDEDENGINE_H

#include <condition_variable>
#include <cstdint>
#include <functional>
#include <memory>
#include <mutex>
#include <thread>
#include <vector>

#include "BatchEngine.h"

namespace hsgpu {

// Contiguous, balanced block of [0, n_total) owned by `rank` of `world` (the same partition as tfe-slam-ucl_b200/sharding.py).
inline void shard_range(int n_total, int world, int rank, int& first, int& count) {
  const int base = n_total / world, rem = n_total % world;
  count = base + (rank < rem ? 1 : 0);
  first = rank * base + (rank < rem ? rank : rem);
}

class ShardedEngine {
 public:
  // n_streams_total HectorSlamProcessor states over n_shards shards (0 = one per visible device).
  ShardedEngine(float mapResolution, int mapSizeX, int mapSizeY, float startX, float startY, int multi_res_size, int n_streams_total,
                int max_points, int n_shards = 0)
      : n_total_(n_streams_total), max_points_(max_points), levels_(multi_res_size) {
    const int n_dev = hs_device_count();
    if (n_dev <= 0) throw Error(HS_ERR_NO_DEVICE, "hsgpu::ShardedEngine");
    if (n_shards <= 0) n_shards = n_dev;
    if (n_shards > n_streams_total) n_shards = n_streams_total > 0 ? n_streams_total : 1;
    shards_.resize(n_shards);
    for (int d = 0; d < n_shards; ++d) {
      Shard& s = shards_[d];
      shard_range(n_streams_total, n_shards, d, s.first, s.count);
      s.device = d % n_dev;
      hs_config cfg;
      cfg.map_resolution = mapResolution; cfg.map_size_x = mapSizeX; cfg.map_size_y = mapSizeY;
      cfg.start_x = startX; cfg.start_y = startY; cfg.levels = multi_res_size;
      cfg.n_streams = s.count + 1;          // + one guest stream: the replica of another shard's stream (hypothesis sharding)
      cfg.max_points = max_points; cfg.device = s.device;
      int st = hs_create(&cfg, &s.engine);
      if (st != HS_OK) { destroy(); throw Error(st, "hs_create"); }
      s.guest_owner = -1;
      s.guest_epoch = 0;
    }
    for (int d = 0; d < n_shards; ++d) workers_.emplace_back(new Worker());
  }
  ~ShardedEngine() {
    workers_.clear();   // joins the threads
    destroy();
  }
  ShardedEngine(const ShardedEngine&) = delete;
  ShardedEngine& operator=(const ShardedEngine&) = delete;

  int shards() const { return (int)shards_.size(); }
  int streams() const { return n_total_; }
  int shardOf(int stream) const {
    for (size_t d = 0; d < shards_.size(); ++d)
      if (stream >= shards_[d].first && stream < shards_[d].first + shards_[d].count) return (int)d;
    return -1;
  }
  int deviceOf(int stream) const { const int d = shardOf(stream); return d < 0 ? -1 : shards_[d].device; }
  hs_engine* engineOfShard(int d) { return shards_[d].engine; }

  // setters of HectorSlamProcessor (HectorSlamProcessor.h:136-139), applied to every stream of every shard
  void setUpdateFactorFree(float f) { for (Shard& s : shards_) check(hs_set_update_factor_free(s.engine, f), "hs_set_update_factor_free"); }
  void setUpdateFactorOccupied(float f) { for (Shard& s : shards_) check(hs_set_update_factor_occupied(s.engine, f), "hs_set_update_factor_occupied"); }
  void setMapUpdateMinDistDiff(float v) { for (Shard& s : shards_) check(hs_set_map_update_min_dist_diff(s.engine, v), "hs_set_map_update_min_dist_diff"); }
  void setMapUpdateMinAngleDiff(float v) { for (Shard& s : shards_) check(hs_set_map_update_min_angle_diff(s.engine, v), "hs_set_map_update_min_angle_diff"); }
  void reset() { parallel([&](int d) { return hs_reset(shards_[d].engine, HS_ALL_STREAMS); }, "hs_reset"); }

  // HectorSlamProcessor::update (HectorSlamProcessor.h:71-113) for ALL streams: slot g = global stream g.
  // points [n_total][max_points][2], counts [n_total]; optional origos [n_total][2], poseHints [n_total][3] (default: each
  // stream's last pose), map_without_matching [n_total]; poses_out [n_total][3], cov_out [n_total][9] or null.
  // Every shard's slice goes through hs_update_batch on its own thread; results are written straight into the caller's arrays.
  void update(const float* points, const int32_t* counts, float* poses_out, float* cov_out = nullptr, const float* origos = nullptr,
              const float* poseHints = nullptr, const uint8_t* map_without_matching = nullptr) {
    const size_t row = (size_t)max_points_ * 2;
    parallel([&](int d) {
      const Shard& s = shards_[d];
      const size_t f = (size_t)s.first;
      return hs_update_batch(s.engine, s.count, nullptr, points + f * row, counts + f, origos ? origos + f * 2 : nullptr,
                             poseHints ? poseHints + f * 3 : nullptr, map_without_matching ? map_without_matching + f : nullptr,
                             poses_out ? poses_out + f * 3 : nullptr, cov_out ? cov_out + f * 9 : nullptr);
    }, "hs_update_batch");
  }

  // getLastScanMatchPose of all streams, concatenated in global stream order
  std::vector<float> lastScanMatchPoses() {
    std::vector<float> out((size_t)n_total_ * 3);
    parallel([&](int d) { return hs_get_poses(shards_[d].engine, 0, shards_[d].count, out.data() + (size_t)shards_[d].first * 3); }, "hs_get_poses");
    return out;
  }
  // getGridMap(level) of one global stream
  std::vector<hs_cell> gridMap(int stream, int level) {
    const int d = shardOf(stream);
    if (d < 0) throw Error(HS_ERR_INVALID_ARG, "hsgpu::ShardedEngine::gridMap");
    int sx = 0, sy = 0;
    check(hs_get_level_dims(shards_[d].engine, level, &sx, &sy, nullptr), "hs_get_level_dims");
    std::vector<hs_cell> cells((size_t)sx * sy);
    check(hs_download_level(shards_[d].engine, stream - shards_[d].first, level, cells.data()), "hs_download_level");
    return cells;
  }
  void synchronize() { parallel([&](int d) { return hs_synchronize(shards_[d].engine); }, "hs_synchronize"); }

  // ---- BASELINE configs[3] sharded --------------------------------------------------------------------
  // OccGridMapUtil::getCompleteHessianDerivs of `stream`'s level for n_poses hypotheses (map coordinates), the hypotheses in
  // contiguous blocks over the shards: H9 [n_poses][9], dTr3 [n_poses][3]. Host pointers.
  void hessianDerivsSharded(int stream, int level, const float* poses_map, int n_poses, const float* points, int n_points, float* H9,
                            float* dTr3) {
    const int owner = replicate(stream);
    parallel([&](int d) {
      int first, count;
      shard_range(n_poses, (int)shards_.size(), d, first, count);
      if (count == 0) return (int)HS_OK;
      const Shard& s = shards_[d];
      const int local = d == owner ? stream - s.first : s.count;    // the owner reads its own stream, the others their guest replica
      return hs_hessian_derivs_batch(s.engine, local, level, poses_map + (size_t)first * 3, count, points, n_points, H9 + (size_t)first * 9,
                                     dTr3 + (size_t)first * 3);
    }, "hs_hessian_derivs_batch");
  }
  // MapRepMultiMap::matchData of one scan against `stream`'s maps from n_hyp world pose hints, sharded the same way.
  void matchHypothesesSharded(int stream, const float* points, int n_points, const float* poseHints, int n_hyp, float* poses_out,
                              float* cov_out = nullptr) {
    const int owner = replicate(stream);
    parallel([&](int d) {
      int first, count;
      shard_range(n_hyp, (int)shards_.size(), d, first, count);
      if (count == 0) return (int)HS_OK;
      const Shard& s = shards_[d];
      const int local = d == owner ? stream - s.first : s.count;
      return hs_match_hypotheses(s.engine, local, points, n_points, poseHints + (size_t)first * 3, count, poses_out + (size_t)first * 3,
                                 cov_out ? cov_out + (size_t)first * 9 : nullptr);
    }, "hs_match_hypotheses");
  }

 private:
  struct Shard {
    hs_engine* engine = nullptr;
    int device = 0, first = 0, count = 0;
    int guest_owner = -1;        // global stream whose maps the guest slot holds
    uint64_t guest_epoch = 0;    // ... as of this map epoch of the owner's engine
  };
  // one persistent thread per shard: takes a job, runs it, reports the status
  struct Worker {
    Worker() : stop(false), has_job(false), done(true), status(HS_OK), th([this] { loop(); }) {}
    ~Worker() {
      { std::lock_guard<std::mutex> g(m); stop = true; }
      cv.notify_all();
      th.join();
    }
    void submit(std::function<int()> j) {
      { std::lock_guard<std::mutex> g(m); job = std::move(j); has_job = true; done = false; }
      cv.notify_all();
    }
    int wait() {
      std::unique_lock<std::mutex> g(m);
      cv.wait(g, [this] { return done; });
      return status;
    }
    void loop() {
      for (;;) {
        std::function<int()> j;
        {
          std::unique_lock<std::mutex> g(m);
          cv.wait(g, [this] { return stop || has_job; });
          if (stop && !has_job) return;
          j = std::move(job);
          has_job = false;
        }
        const int st = j();
        { std::lock_guard<std::mutex> g(m); status = st; done = true; }
        cv.notify_all();
      }
    }
    std::mutex m;
    std::condition_variable cv;
    bool stop, has_job, done;
    int status;
    std::function<int()> job;
    std::thread th;
  };

  // runs fn(d) for every shard on that shard's thread and waits for all of them; the first failure is thrown
  template <typename F>
  void parallel(F fn, const char* what) {
    for (size_t d = 0; d < shards_.size(); ++d) workers_[d]->submit([fn, d] { return fn((int)d); });
    int bad = HS_OK;
    for (size_t d = 0; d < shards_.size(); ++d) {
      const int st = workers_[d]->wait();
      if (st != HS_OK && bad == HS_OK) bad = st;
    }
    check(bad, what);
  }

  // makes sure every non-owner shard's guest slot holds the current maps of `stream`; returns the owner shard
  int replicate(int stream) {
    const int owner = shardOf(stream);
    if (owner < 0) throw Error(HS_ERR_INVALID_ARG, "hsgpu::ShardedEngine: stream out of range");
    Shard& o = shards_[owner];
    const uint64_t epoch = hs_get_map_epoch(o.engine);
    parallel([&, owner, epoch](int d) {
      Shard& s = shards_[d];
      if (d == owner || (s.guest_owner == stream && s.guest_epoch == epoch)) return (int)HS_OK;
      const int st = hs_copy_stream_maps(s.engine, s.count, o.engine, stream - o.first);
      if (st == HS_OK) { s.guest_owner = stream; s.guest_epoch = epoch; }
      return st;
    }, "hs_copy_stream_maps");
    return owner;
  }

  void destroy() {
    for (Shard& s : shards_) { if (s.engine) hs_destroy(s.engine); s.engine = nullptr; }
  }
  static void check(int status, const char* what) { if (status != HS_OK) throw Error(status, what); }

  int n_total_, max_points_, levels_;
  std::vector<Shard> shards_;
  std::vector<std::unique_ptr<Worker>> workers_;
};

}  // namespace hsgpu
#endif
