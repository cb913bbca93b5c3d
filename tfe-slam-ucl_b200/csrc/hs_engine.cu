// hsgpu engine: host side of the C ABI declared in include/hsgpu.h. Owns the device memory of all streams
// (grids resident in HBM for the engine's lifetime) and sequences the kernels of hs_kernels.cuh on one CUDA
// stream. There is deliberately no CPU implementation of the hot path in this library: every entry point
// either runs the CUDA kernels or returns an error status.
#include <cuda_runtime.h>

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <new>
#include <vector>

#include "hs_kernels.cuh"
#include "hsgpu.h"

#define HS_MAX_CHUNKS 4

struct hs_engine {
  hs_config cfg;
  HsParams P;
  int device;
  cudaStream_t own_stream;
  cudaStream_t stream;
  // host batches are cut into chunks whose host-to-device copies overlap the kernels of the previous chunk
  int e2e_chunks, e2e_first_pct, e2e_zero_copy, e2e_defer, sm_count = 148;
  int ray_alternate = 1; unsigned ray_launches = 0;   // k_raycast slot order alternates between launches
  int prof_every = 1; unsigned prof_tick = 0;   // kernel events around every prof_every-th update only
  cudaEvent_t ev_match = nullptr;
  cudaStream_t aux_stream[HS_MAX_CHUNKS - 1];
  cudaEvent_t ev_fork, ev_join[HS_MAX_CHUNKS - 1];
  // staging (device) for the host-pointer entry points
  float* d_points;
  int32_t* d_counts;
  float* d_origos;
  float* d_hints;
  uint8_t* d_flags;
  int32_t* d_ids;
  float* d_ranges;
  float* d_out;      // [n_streams][9] gather scratch
  // pipelined submission (hs_submit_batch): two sets of input staging buffers filled by DMA on a copy stream while the
  // previous batch computes; allocated on first use
  struct PipeBuf { float* points; float* ranges; int32_t* counts; float* origos; float* hints; uint8_t* flags; int32_t* ids; cudaEvent_t staged, consumed; };
  PipeBuf pipe[2];
  cudaStream_t copy_stream;
  uint64_t submits;
  float2* d_beam_cs; // scan geometry
  int n_beams;
  float range_min, max_range_for_container, angle_min, angle_increment, range_max;
  // misc scratch for single-function entry points
  void* d_scratch;
  size_t scratch_bytes;
  // bookkeeping
  uint64_t updates, launches, h2d_bytes, d2h_bytes;
  float free_factor, occ_factor;
  int match_variant;         // matcher variant: 2 = k_match2 (shared-memory cache, dense probability pass) for long scans, k_match
                             // (register cache) for scans of <= 768 points; HSGPU_MATCH_VARIANT=1|2 overrides
  int log2_seg_levels;       // k_raycast walk segment lengths (log2): level 0 | coarser levels << 8
  // optional per-kernel timing with CUDA events on the launching stream (hs_profile_*)
  int profiling;
  std::vector<cudaEvent_t>* prof_events;  // pairs: start, stop
  std::vector<int>* prof_ids;
  size_t prof_used;
  std::vector<uint64_t>* id_bitmap;        // check_ids scratch
  // probability block plane of one (stream, level) for the hypothesis kernels (ensure_pblk)
  float4* pblk;
  int pblk_stream;
  uint64_t pblk_epoch, map_epoch;          // map_epoch: bumped whenever cells of any stream may have been written
  int hd_ppw;                              // HSGPU_HD_PPW: poses per warp in k_hessian_derivs (0 = by batch size)
  int match_ct_pitch;                      // HSGPU_MATCH_CT_PITCH (A/B): use the compile-time-pitch instantiation of k_match when it applies
  int hyp_use_pblk;                        // HSGPU_HYP_PBLK (A/B): multi-start matching through the probability block planes
  int match2_threads;                      // HSGPU_MATCH2_THREADS: 128 / 256 / 512 threads per k_match2 CTA (0 = by batch size)
  int ray_big;                             // k_raycast as one 1024-thread CTA per SM with a 144 KB map (long scans, launch_raycast)
  void* h_pinned;                          // engine-owned pinned host staging (ensure_pinned)
  size_t pinned_bytes;
};

enum { HS_K_SCAN_TO_POINTS = 0, HS_K_MATCH = 1, HS_K_RAYCAST = 2, HS_K_HESSIAN = 3, HS_K_BUILD_PBLK = 4, HS_K_COUNT = HS_PROFILE_KERNELS };

struct ProfScope {
  hs_engine* e;
  cudaEvent_t stop;
  ProfScope(hs_engine* eng, int id) : e(eng), stop(nullptr) {
    if (!e->profiling) return;
    if (e->prof_used + 2 > e->prof_events->size()) {
      cudaEvent_t a, b;
      if (cudaEventCreate(&a) != cudaSuccess || cudaEventCreate(&b) != cudaSuccess) return;
      e->prof_events->push_back(a);
      e->prof_events->push_back(b);
    }
    cudaEvent_t start = (*e->prof_events)[e->prof_used];
    stop = (*e->prof_events)[e->prof_used + 1];
    e->prof_used += 2;
    e->prof_ids->push_back(id);
    cudaEventRecord(start, e->stream);
  }
  ~ProfScope() {
    if (stop) cudaEventRecord(stop, e->stream);
  }
};

#define HS_CK(call)                                  \
  do {                                               \
    cudaError_t err__ = (call);                      \
    if (err__ != cudaSuccess) return (int)err__;     \
  } while (0)

static inline int hs_launch_check(hs_engine* e) {
  e->launches++;
  cudaError_t err = cudaGetLastError();
  return err == cudaSuccess ? HS_OK : (int)err;
}
#define HS_LAUNCHED(e)                 \
  do {                                 \
    int st__ = hs_launch_check(e);     \
    if (st__ != HS_OK) return st__;    \
  } while (0)

static int raycast_log2_hash(int max_points) {
  int l = 4;
  while ((1 << l) < 2 * max_points) ++l;
  return l;
}
static size_t raycast_smem_for(int max_points, size_t map_bytes = HS_MAP_BYTES) {
  return map_bytes + 2 * sizeof(int) * ((size_t)1 << raycast_log2_hash(max_points)) + (size_t)max_points * sizeof(HsWalkRec) + (size_t)max_points * sizeof(HsBeamRec) + 2 * (size_t)max_points * sizeof(unsigned short) + 16;
}
static size_t match2_smem_for(int max_points) {
  const size_t pitch = (size_t)(((max_points + 27) & ~31) + 4);   // hs_match2_pitch
  return HS_NPROD * pitch * sizeof(float) + (size_t)max_points * (sizeof(float4) + sizeof(float2) + sizeof(int) + sizeof(unsigned short)) + 16;
}

// The ray-cast launch: one CTA per (slot, level). Scans of up to 768 points: 384 threads, a 48 KB cell map, three CTAs per SM.
// Longer scans (BASELINE configs[2]: 1440 beams) need ~80 KB of hash table and beam records per CTA, so only ONE CTA fits an
// SM: it then takes the whole SM -- 1024 threads and a 144 KB map (590 k cells per band instead of 197 k).
// Measured and dropped on B200 (configs[1]): the coarse levels as a second, concurrent launch of 128-thread CTAs with a 12 KB
// map (k_raycast 163 -> 182 us: the small CTAs are slower than the slots they free).
static int launch_raycast(hs_engine* e, const HsParams& Pc, int n, const int32_t* ids, const float2* points, const int32_t* counts,
                          const float2* origos, cudaStream_t st) {
  const int log2_hash = raycast_log2_hash(e->P.max_points);
#ifdef HS_EXPERIMENTS
  if (const char* ol = getenv("HSGPU_DEBUG_ONLY_LEVEL")) {   // timing experiment (wrong maps): one level's CTAs alone
    const int l = atoi(ol);
    k_raycast<HS_RAYCAST_THREADS, HS_MAP_BYTES, HS_RAYCAST_CTAS><<<n, HS_RAYCAST_THREADS, raycast_smem_for(e->P.max_points), st>>>(
        Pc, n, ids, points, counts, origos, log2_hash, e->log2_seg_levels, l, 1);
    return hs_launch_check(e);
  }
#endif
  if (e->ray_big)
    k_raycast<HS_RAYCAST_THREADS_BIG, HS_MAP_BYTES_BIG, 1><<<n * e->P.n_levels, HS_RAYCAST_THREADS_BIG, raycast_smem_for(e->P.max_points, HS_MAP_BYTES_BIG), st>>>(
        Pc, n, ids, points, counts, origos, log2_hash, e->log2_seg_levels, 0, e->P.n_levels);
  else
    k_raycast<HS_RAYCAST_THREADS, HS_MAP_BYTES, HS_RAYCAST_CTAS><<<n * e->P.n_levels, HS_RAYCAST_THREADS, raycast_smem_for(e->P.max_points), st>>>(
        Pc, n, ids, points, counts, origos, log2_hash, e->log2_seg_levels, 0, e->P.n_levels);
  return hs_launch_check(e);
}
static int warp_grid(int n_warps, int threads) { return (n_warps * 32 + threads - 1) / threads; }

// GridMapLogOddsFunctions::probToLogOdds (HSL/map/GridMapLogOdds.h:198-203), host libm like the reference.
static float prob_to_log_odds(float prob) {
  volatile float odds = prob / (1.0f - prob);
  return logf(odds);
}

// GridMapBase::setMapTransformation (HSL/map/GridMapBase.h:265-280): Scaling(s)*Translation(off) and its
// affine inverse with Eigen 3.3's 2x2 inverse ((1/det) * adjugate, t' = (-Linv) t), fp32, unfused.
static void level_transform(HsLevel* L, float offx, float offy, float cell_length) {
  volatile float s = 1.0f / cell_length;
  volatile float tmx = s * offx, tmy = s * offy;
  volatile float m00 = s, m01 = 0.0f, m10 = 0.0f, m11 = s;
  volatile float p0 = m00 * m11, p1 = m10 * m01;
  volatile float det = p0 - p1;
  volatile float invdet = 1.0f / det;
  volatile float i00 = m11 * invdet, i10 = -m10 * invdet, i01 = -m01 * invdet, i11 = m00 * invdet;
  volatile float q0 = (-i00) * tmx, q1 = (-i01) * tmy, q2 = (-i10) * tmx, q3 = (-i11) * tmy;
  volatile float twx = q0 + q1, twy = q2 + q3;
  (void)i11;
  L->scale = s;
  L->tmx = tmx;
  L->tmy = tmy;
  L->a = i00;
  L->twx = twx;
  L->twy = twy;
}

static int ensure_scratch(hs_engine* e, size_t bytes) {
  if (bytes <= e->scratch_bytes) return HS_OK;
  if (e->d_scratch) cudaFree(e->d_scratch);
  e->d_scratch = nullptr;
  e->scratch_bytes = 0;
  HS_CK(cudaMalloc(&e->d_scratch, bytes));
  e->scratch_bytes = bytes;
  return HS_OK;
}

extern "C" {

const char* hs_status_string(int status) {
  switch (status) {
    case HS_OK: return "ok";
    case HS_ERR_INVALID_ARG: return "invalid argument";
    case HS_ERR_NO_DEVICE: return "no usable CUDA device (hsgpu has no CPU fallback)";
    case HS_ERR_OUT_OF_MEMORY: return "out of memory";
    case HS_ERR_NOT_CONFIGURED: return "scan geometry not set (hs_set_scan_geometry)";
    default: break;
  }
  if (status > 0) return cudaGetErrorString((cudaError_t)status);
  return "unknown hsgpu status";
}

int hs_create(const hs_config* cfg, hs_engine** out) {
  if (!cfg || !out) return HS_ERR_INVALID_ARG;
  *out = nullptr;
  if (cfg->levels < 1 || cfg->levels > HS_MAX_LEVELS || cfg->n_streams < 1 || cfg->max_points < 1 || cfg->max_points > 5600 ||
      cfg->map_size_x < 2 || cfg->map_size_y < 2 || cfg->map_size_x > 16384 || cfg->map_size_y > 16384 || !(cfg->map_resolution > 0.0f))
    return HS_ERR_INVALID_ARG;
  if ((cfg->map_size_x >> (cfg->levels - 1)) < 2 || (cfg->map_size_y >> (cfg->levels - 1)) < 2) return HS_ERR_INVALID_ARG;
  int n_dev = 0;
  cudaError_t derr = cudaGetDeviceCount(&n_dev);
  if (derr != cudaSuccess || n_dev <= 0) return HS_ERR_NO_DEVICE;
  if (cfg->device < 0 || cfg->device >= n_dev) return HS_ERR_INVALID_ARG;
  HS_CK(cudaSetDevice(cfg->device));

  // smem budgets of the two per-scan kernels grow with max_points: refuse what the device cannot launch BEFORE allocating
  int optin_smem = 0;
  {
    int optin = 0;
    HS_CK(cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, cfg->device));
    optin_smem = optin;
    const size_t match_smem = cfg->max_points > 768 ? match2_smem_for(cfg->max_points)
                                                    : (size_t)hs_match2_pitch(cfg->max_points) * HS_NPROD * sizeof(float) + 6 * 4 * 160 * sizeof(float) + 256;
    if (raycast_smem_for(cfg->max_points) > (size_t)optin || match_smem > (size_t)optin) return HS_ERR_INVALID_ARG;
  }
  // value-initialisation: every member zero, then the default member initialisers (ray_alternate, prof_every, sm_count)
  hs_engine* e = new (std::nothrow) hs_engine();
  if (!e) return HS_ERR_OUT_OF_MEMORY;
  e->prof_events = new std::vector<cudaEvent_t>();
  e->prof_ids = new std::vector<int>();
  e->id_bitmap = new std::vector<uint64_t>();
  e->cfg = *cfg;
  e->device = cfg->device;
  HsParams& P = e->P;
  P.n_levels = cfg->levels;
  P.n_streams = cfg->n_streams;
  P.max_points = cfg->max_points;
  { const char* d = getenv("HSGPU_DEBUG_SKIP"); P.dbg_skip = d ? atoi(d) : 0; }   // timing experiments only: results are wrong when set
  {
    int n_sm = 148;
    cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, cfg->device);
    e->sm_count = n_sm;
    P.n_sm = n_sm;
    P.l2_keep_levels = 0xFF;
    P.reverse_order = 0;
    if (const char* ra = getenv("HSGPU_RAYCAST_ALTERNATE")) e->ray_alternate = atoi(ra) != 0;
    // The matcher's evict-last gathers live in the persisting part of L2. Measured on B200 (k_raycast us per launch): 0 MB 170,
    // 8 MB 167, 16 MB 164, 24 MB (this driver's default) 165, 32 MB 173, 64 MB 271 -- a large set-aside starves the
    // ray-caster's streaming traffic. A process that has chosen a size keeps it; only an unset (0) limit is raised.
    if (const char* pm = getenv("HSGPU_L2_PERSIST_MB")) {
      cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, (size_t)atoi(pm) << 20);
    } else {
      size_t cur = 0;
      if (cudaDeviceGetLimit(&cur, cudaLimitPersistingL2CacheSize) == cudaSuccess && cur == 0)
        cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, (size_t)16 << 20);
    }
    cudaGetLastError();
    if (const char* k = getenv("HSGPU_MATCH_L2_LEVELS")) P.l2_keep_levels = atoi(k);
    P.tail_slots = 2 * n_sm;   // measured on B200 (148 SMs): 0 -> 179 us, 148 -> 178, 296 -> 176, 444 -> 178.5, 600 -> 181
  }
  if (const char* pw = getenv("HSGPU_HD_PPW")) e->hd_ppw = atoi(pw);
  if (const char* ts = getenv("HSGPU_RAYCAST_TAIL_SLOTS")) P.tail_slots = atoi(ts) < 0 ? 0 : atoi(ts);
  P.strict = 1;  // reference-order sums by default: bitwise the CPU reference's poses

  // MapRepMultiMap ctor (HSL/slam_main/MapRepMultiMap.h:48-72): same world offset for every level,
  // resolution doubles and the grid halves (integer division) per level.
  float res = cfg->map_resolution;
  int sx = cfg->map_size_x, sy = cfg->map_size_y;
  volatile float total_x = res * (float)sx;
  volatile float mid_x = total_x * cfg->start_x;
  volatile float total_y = res * (float)sy;
  volatile float mid_y = total_y * cfg->start_y;
  unsigned long long off = 0;
  for (int i = 0; i < cfg->levels; ++i) {
    HsLevel& L = P.lv[i];
    L.sx = sx;
    L.sy = sy;
    L.limx = (float)sx - 2.0f;
    L.limy = (float)sy - 2.0f;
    L.pt_factor = (float)(1.0 / pow(2.0, (double)i));
    L.max_iter = (i == 0) ? 5 : 3;
    L.cell_offset = off;
    level_transform(&L, mid_x, mid_y, res);
    off += (unsigned long long)sx * (unsigned long long)sy;
    sx /= 2;
    sy /= 2;
    res *= 2.0f;
  }
  P.cells_per_stream = off;
  {
    int s0 = 4, s1 = 2;   // 16-step segments on the fine grid, 4-step segments on the coarse ones (measured on B200: 5,4 217 us; 4,2 214; 0,0 224)
    const char* env = getenv("HSGPU_RAYCAST_SEG_LOG2");   // tuning hook: "<level0>,<coarse>"
    if (env) sscanf(env, "%d,%d", &s0, &s1);
    s0 = s0 < 0 ? 0 : (s0 > 12 ? 12 : s0);
    s1 = s1 < 0 ? 0 : (s1 > 12 ? 12 : s1);
    e->log2_seg_levels = s0 | (s1 << 8);
  }
  e->free_factor = 0.4f;
  e->occ_factor = 0.6f;
  P.log_odds_free = prob_to_log_odds(0.4f);
  P.log_odds_occ = prob_to_log_odds(0.6f);
  P.min_dist = 0.4f * 1.0f;
  P.min_angle = 0.13f * 1.0f;

  const size_t ns = (size_t)cfg->n_streams, mp = (size_t)cfg->max_points;
  const size_t cells = ns * (size_t)P.cells_per_stream;
  int st = HS_OK;
#define HS_ALLOC(ptr, bytes)                                            \
  if (st == HS_OK) {                                                    \
    cudaError_t err__ = cudaMalloc((void**)&(ptr), (bytes));            \
    if (err__ != cudaSuccess) st = (err__ == cudaErrorMemoryAllocation) ? HS_ERR_OUT_OF_MEMORY : (int)err__; \
  }
  HS_ALLOC(P.val, cells * sizeof(float));
  HS_ALLOC(P.idx, cells * sizeof(int));
  HS_ALLOC(P.mk, cells * sizeof(uint8_t));
  HS_ALLOC(P.mk_base, ns * HS_MAX_LEVELS * sizeof(int));
  HS_ALLOC(P.last_pose, ns * 3 * sizeof(float));
  HS_ALLOC(P.last_update_pose, ns * 3 * sizeof(float));
  HS_ALLOC(P.cov, ns * 9 * sizeof(float));
  HS_ALLOC(P.cur_update_index, ns * sizeof(int));
  HS_ALLOC(P.last_update_index, ns * sizeof(int));
  HS_ALLOC(P.coarse_pts, ns * mp * 2 * sizeof(float));
  HS_ALLOC(P.coarse_cnt, ns * sizeof(int));
  HS_ALLOC(P.coarse_origo, ns * 2 * sizeof(float));
  HS_ALLOC(P.slot_pose, ns * 3 * sizeof(float));
  HS_ALLOC(P.slot_sc, ns * 2 * sizeof(float));
  HS_ALLOC(P.slot_integrate, ns * sizeof(int));
  HS_ALLOC(P.slot_mark_base, ns * sizeof(int));
  HS_ALLOC(P.counters, 4 * sizeof(unsigned long long));
  HS_ALLOC(e->d_points, ns * mp * 2 * sizeof(float));
  HS_ALLOC(e->d_counts, ns * sizeof(int32_t));
  HS_ALLOC(e->d_origos, ns * 2 * sizeof(float));
  HS_ALLOC(e->d_hints, ns * 3 * sizeof(float));
  HS_ALLOC(e->d_flags, ns * sizeof(uint8_t));
  HS_ALLOC(e->d_ids, ns * sizeof(int32_t));
  HS_ALLOC(e->d_out, ns * 9 * sizeof(float));
#undef HS_ALLOC
  if (st == HS_OK) {
    cudaError_t err = cudaStreamCreateWithFlags(&e->own_stream, cudaStreamNonBlocking);
    if (err != cudaSuccess) st = (int)err;
  }
  if (st != HS_OK) {
    cudaGetLastError();
    hs_destroy(e);
    return st;
  }
  e->stream = e->own_stream;
  {
    e->e2e_chunks = 2;   // measured on B200 (1024 streams x 360 points, pinned buffers): see DESIGN.md section 5
    if (const char* c = getenv("HSGPU_E2E_CHUNKS")) e->e2e_chunks = atoi(c);
    e->e2e_chunks = e->e2e_chunks < 1 ? 1 : (e->e2e_chunks > HS_MAX_CHUNKS ? HS_MAX_CHUNKS : e->e2e_chunks);
    e->e2e_first_pct = 100;
    e->e2e_zero_copy = 1;
    if (const char* z = getenv("HSGPU_E2E_ZERO_COPY")) e->e2e_zero_copy = atoi(z) != 0;
    e->e2e_defer = 1;    // 0: the call returns after the map integration; 1/2: after the match (inputs prefetched by DMA / by a copy kernel)
    if (const char* z = getenv("HSGPU_E2E_DEFER")) e->e2e_defer = atoi(z);
    if (const char* c = getenv("HSGPU_E2E_FIRST_PCT")) e->e2e_first_pct = atoi(c) < 5 ? 5 : (atoi(c) > 100 ? 100 : atoi(c));
    cudaError_t err = cudaEventCreateWithFlags(&e->ev_fork, cudaEventDisableTiming);
    for (int c = 0; c < HS_MAX_CHUNKS - 1 && err == cudaSuccess; ++c) {
      err = cudaStreamCreateWithFlags(&e->aux_stream[c], cudaStreamNonBlocking);
      if (err == cudaSuccess) err = cudaEventCreateWithFlags(&e->ev_join[c], cudaEventDisableTiming);
    }
    if (err != cudaSuccess) {
      hs_destroy(e);
      return (int)err;
    }
  }
  {
    int smem = hs_match2_pitch(cfg->max_points) * HS_NPROD * (int)sizeof(float);  // k_match's per-CTA product buffer
    if (smem > 48 * 1024) {
      cudaError_t e1 = cudaFuncSetAttribute(k_match<true, 0, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
      cudaError_t e2 = cudaFuncSetAttribute(k_match<false, 0, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
      if (e1 != cudaSuccess || e2 != cudaSuccess) {
        hs_destroy(e);
        return (int)(e1 != cudaSuccess ? e1 : e2);
      }
    }
    // measured on B200, per 1024 scans: 360 points: k_match 101 us, k_match with warp-compacted misses (3) 95 us, k_match2 127 us;
    // 1440 points: k_match 400 us vs k_match2 353 us
    e->match_variant = cfg->max_points > 768 ? 2 : 3;
    if (const char* mv = getenv("HSGPU_MATCH_VARIANT")) {
      const int v = atoi(mv);
      if (v >= 1 && v <= 3) e->match_variant = v;
    }
    int smem2 = (int)match2_smem_for(cfg->max_points);
    if (smem2 > 48 * 1024) {
      cudaError_t e4 = cudaSuccess;
#define HS_M2_ATTR(S_, T_) if (e4 == cudaSuccess) e4 = cudaFuncSetAttribute(k_match2<S_, T_>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem2)
      HS_M2_ATTR(true, 128); HS_M2_ATTR(true, 256); HS_M2_ATTR(true, 512); HS_M2_ATTR(false, 128); HS_M2_ATTR(false, 256); HS_M2_ATTR(false, 512);
#undef HS_M2_ATTR
      if (e4 != cudaSuccess) {
        hs_destroy(e);
        return (int)e4;
      }
    }
    if (const char* mt = getenv("HSGPU_MATCH2_THREADS")) e->match2_threads = atoi(mt);
    int smem_r = (int)raycast_smem_for(cfg->max_points);  // k_raycast's cell map + hash table + beam records
    cudaError_t e3 = cudaFuncSetAttribute(k_raycast<HS_RAYCAST_THREADS, HS_MAP_BYTES, HS_RAYCAST_CTAS>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_r);
    // long scans: ~80 KB of tables per CTA leave room for one or two CTAs per SM only -- use the big-CTA instantiation if it fits
    e->ray_big = cfg->max_points > 768 && raycast_smem_for(cfg->max_points, HS_MAP_BYTES_BIG) <= (size_t)optin_smem;
    if (const char* sp = getenv("HSGPU_RAYCAST_BIG")) e->ray_big = atoi(sp) != 0 && raycast_smem_for(cfg->max_points, HS_MAP_BYTES_BIG) <= (size_t)optin_smem;
    if (e3 == cudaSuccess && e->ray_big)
      e3 = cudaFuncSetAttribute(k_raycast<HS_RAYCAST_THREADS_BIG, HS_MAP_BYTES_BIG, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                (int)raycast_smem_for(cfg->max_points, HS_MAP_BYTES_BIG));
    if (e3 != cudaSuccess) {
      hs_destroy(e);
      return (int)e3;
    }
    e->hyp_use_pblk = 1;
    if (const char* sp = getenv("HSGPU_HYP_PBLK")) e->hyp_use_pblk = atoi(sp);
    e->match_ct_pitch = 1;
    if (const char* sp = getenv("HSGPU_MATCH_CT_PITCH")) e->match_ct_pitch = atoi(sp);
  }
  if (getenv("HSGPU_DEBUG_TIMES")) {   // timing experiments only
    if (cudaMalloc((void**)&P.dbg_times, ns * HS_MAX_LEVELS * 8 * sizeof(unsigned long long)) != cudaSuccess) P.dbg_times = nullptr;
  }
  cudaMemsetAsync(P.counters, 0, 4 * sizeof(unsigned long long), e->stream);
  cudaMemsetAsync(P.slot_integrate, 0, ns * sizeof(int), e->stream);
  cudaMemsetAsync(P.coarse_pts, 0, ns * mp * 2 * sizeof(float), e->stream);
  k_reset_state<<<(cfg->n_streams + 127) / 128, 128, 0, e->stream>>>(P, 0, cfg->n_streams, 1);
  e->launches++;
  k_clear_cells<<<148 * 8, 256, 0, e->stream>>>(P.val, P.idx, P.mk, cells);
  e->launches++;
  cudaError_t err = cudaStreamSynchronize(e->stream);
  if (err != cudaSuccess) {
    hs_destroy(e);
    return (int)err;
  }
  *out = e;
  return HS_OK;
}

int hs_destroy(hs_engine* e) {
  if (!e) return HS_OK;
  cudaSetDevice(e->device);
  if (e->own_stream) cudaStreamSynchronize(e->own_stream);
  if (e->P.dbg_times) {   // timing experiments only: dump the stamps of the last ray-cast launch
    const char* path = getenv("HSGPU_DEBUG_TIMES");
    size_t n = (size_t)e->P.n_streams * HS_MAX_LEVELS * 8;
    std::vector<unsigned long long> h(n);
    cudaDeviceSynchronize();
    if (path && cudaMemcpy(h.data(), e->P.dbg_times, n * sizeof(unsigned long long), cudaMemcpyDeviceToHost) == cudaSuccess) {
      FILE* f = fopen(path, "wb");
      if (f) { fwrite(h.data(), sizeof(unsigned long long), n, f); fclose(f); }
    }
    cudaFree(e->P.dbg_times);
  }
  HsParams& P = e->P;
  void* ptrs[] = {P.val, P.idx, P.mk, P.mk_base, P.last_pose, P.last_update_pose, P.cov, P.cur_update_index, P.last_update_index,
                  P.coarse_pts, P.coarse_cnt, P.coarse_origo, P.slot_pose, P.slot_sc, P.slot_integrate, P.slot_mark_base, P.counters,
                  e->d_points, e->d_counts, e->d_origos, e->d_hints, e->d_flags, e->d_ids, e->d_ranges, e->d_out,
                  e->d_beam_cs, e->d_scratch, e->pblk};
  for (void* p : ptrs)
    if (p) cudaFree(p);
  if (e->copy_stream) { cudaStreamSynchronize(e->copy_stream); cudaStreamDestroy(e->copy_stream); }
  for (int b = 0; b < 2; ++b) {
    void* pp[] = {e->pipe[b].points, e->pipe[b].ranges, e->pipe[b].counts, e->pipe[b].origos, e->pipe[b].hints, e->pipe[b].flags, e->pipe[b].ids};
    for (void* q : pp) if (q) cudaFree(q);
    if (e->pipe[b].staged) cudaEventDestroy(e->pipe[b].staged);
    if (e->pipe[b].consumed) cudaEventDestroy(e->pipe[b].consumed);
  }
  for (int c = 0; c < HS_MAX_CHUNKS - 1; ++c) {
    if (e->aux_stream[c]) { cudaStreamSynchronize(e->aux_stream[c]); cudaStreamDestroy(e->aux_stream[c]); }
    if (e->ev_join[c]) cudaEventDestroy(e->ev_join[c]);
  }
  if (e->ev_fork) cudaEventDestroy(e->ev_fork);
  if (e->ev_match) cudaEventDestroy(e->ev_match);
  if (e->own_stream) cudaStreamDestroy(e->own_stream);
  if (e->prof_events) {
    for (cudaEvent_t ev : *e->prof_events) cudaEventDestroy(ev);
    delete e->prof_events;
    delete e->prof_ids;
  }
  delete e->id_bitmap;
  if (e->h_pinned) cudaFreeHost(e->h_pinned);
  delete e;
  return HS_OK;
}

int hs_reset(hs_engine* e, int stream) {
  if (!e) return HS_ERR_INVALID_ARG;
  if (stream != HS_ALL_STREAMS && (stream < 0 || stream >= e->P.n_streams)) return HS_ERR_INVALID_ARG;
  HS_CK(cudaSetDevice(e->device));
  e->map_epoch++;
  int first = stream == HS_ALL_STREAMS ? 0 : stream;
  int n = stream == HS_ALL_STREAMS ? e->P.n_streams : 1;
  k_reset_state<<<(n + 127) / 128, 128, 0, e->stream>>>(e->P, first, n, 0);
  HS_LAUNCHED(e);
  size_t cells = (size_t)n * (size_t)e->P.cells_per_stream;
  size_t base = (size_t)first * (size_t)e->P.cells_per_stream;
  k_clear_cells<<<148 * 8, 256, 0, e->stream>>>(e->P.val + base, e->P.idx + base, e->P.mk + base, cells);
  HS_LAUNCHED(e);
  return HS_OK;
}

int hs_set_update_factor_free(hs_engine* e, float f) {
  if (!e) return HS_ERR_INVALID_ARG;
  e->free_factor = f;
  e->P.log_odds_free = prob_to_log_odds(f);
  return HS_OK;
}
int hs_set_update_factor_occupied(hs_engine* e, float f) {
  if (!e) return HS_ERR_INVALID_ARG;
  e->occ_factor = f;
  e->P.log_odds_occ = prob_to_log_odds(f);
  return HS_OK;
}
int hs_set_map_update_min_dist_diff(hs_engine* e, float d) {
  if (!e) return HS_ERR_INVALID_ARG;
  e->P.min_dist = d;
  return HS_OK;
}
int hs_set_map_update_min_angle_diff(hs_engine* e, float a) {
  if (!e) return HS_ERR_INVALID_ARG;
  e->P.min_angle = a;
  return HS_OK;
}
int hs_set_strict_summation(hs_engine* e, int on) {
  if (!e) return HS_ERR_INVALID_ARG;
  e->P.strict = on ? 1 : 0;
  return HS_OK;
}
float hs_get_scale_to_map(const hs_engine* e) { return e ? e->P.lv[0].scale : 0.0f; }
int hs_get_map_levels(const hs_engine* e) { return e ? e->P.n_levels : 0; }
int hs_get_n_streams(const hs_engine* e) { return e ? e->P.n_streams : 0; }
int hs_get_max_points(const hs_engine* e) { return e ? e->P.max_points : 0; }
int hs_get_level_dims(const hs_engine* e, int level, int* sx, int* sy, float* cell_length) {
  if (!e || level < 0 || level >= e->P.n_levels) return HS_ERR_INVALID_ARG;
  if (sx) *sx = e->P.lv[level].sx;
  if (sy) *sy = e->P.lv[level].sy;
  if (cell_length) {
    float res = e->cfg.map_resolution;
    for (int i = 0; i < level; ++i) res *= 2.0f;
    *cell_length = res;
  }
  return HS_OK;
}
int hs_set_cuda_stream(hs_engine* e, void* cuda_stream) {
  if (!e) return HS_ERR_INVALID_ARG;
  HS_CK(cudaSetDevice(e->device));
  HS_CK(cudaStreamSynchronize(e->stream));
  e->stream = cuda_stream ? (cudaStream_t)cuda_stream : e->own_stream;
  return HS_OK;
}
int hs_synchronize(hs_engine* e) {
  if (!e) return HS_ERR_INVALID_ARG;
  HS_CK(cudaSetDevice(e->device));
  HS_CK(cudaStreamSynchronize(e->stream));
  return HS_OK;
}

// ---- the hot path ----------------------------------------------------------------------------

static int check_batch(const hs_engine* e, int n) {
  if (!e) return HS_ERR_INVALID_ARG;
  if (n < 0 || n > e->P.n_streams) return HS_ERR_INVALID_ARG;
  return HS_OK;
}

// match (+gate) then ray-cast, all pointers on the device
// Slots [slot0, slot0 + n) of the batch on CUDA stream `st` (the engine's stream, or one of its auxiliary streams when a
// host batch is cut into chunks); the pointers are those of the WHOLE batch.
static int run_update_device(hs_engine* e, int n, const int32_t* ids, const float* points, const int32_t* counts,
                             const float* origos, const float* hints, const uint8_t* flags, int mode,
                             cudaStream_t st = nullptr, int slot0 = 0, float* stage = nullptr, int32_t* stage_cnt = nullptr,
                             float* host_pose_out = nullptr, cudaEvent_t ev_matched = nullptr) {
  if (n == 0) return HS_OK;
  if (!st) st = e->stream;
  HsParams Pc = e->P;
  Pc.slot0 = slot0;
  Pc.stage_pts = (float2*)stage;   // non-NULL: `points` is pinned host memory, the matcher stages each scan into `stage`
  Pc.stage_cnt = stage ? stage_cnt : nullptr;   // non-NULL: `counts` is pinned host memory too
  Pc.host_pose_out = host_pose_out;
  const float* points_ray = stage ? stage : points;
  const int32_t* counts_ray = (stage && stage_cnt) ? stage_cnt : counts;
  struct StreamSwap {   // ProfScope records on e->stream
    hs_engine* e; cudaStream_t saved;
    StreamSwap(hs_engine* en, cudaStream_t s) : e(en), saved(en->stream) { e->stream = s; }
    ~StreamSwap() { e->stream = saved; }
  } swap(e, st);
  struct ProfGate {   // sampled kernel timing: the events of the other updates stay out of the stream
    hs_engine* e; int saved;
    explicit ProfGate(hs_engine* en) : e(en), saved(en->profiling) {
      if (saved && e->prof_every > 1 && (e->prof_tick++ % (unsigned)e->prof_every) != 0) e->profiling = 0;
    }
    ~ProfGate() { e->profiling = saved; }
  } gate(e);
  {
    ProfScope ps(e, HS_K_MATCH);
    const int threads = HS_MATCH_THREADS;
    if (e->match_variant == 2) {
      const size_t smem2 = match2_smem_for(e->P.max_points);
      // few long scans: more threads per scan (measured on B200, 256 x 1440 points: see DESIGN.md)
      int t2 = n >= 4 * e->sm_count ? 128 : (n >= 2 * e->sm_count ? 256 : 512);
      if (e->match2_threads == 128 || e->match2_threads == 256 || e->match2_threads == 512) t2 = e->match2_threads;
#define HS_LAUNCH_MATCH2(STRICT_)                                                                                                          \
  do {                                                                                                                                     \
    if (t2 == 512) k_match2<STRICT_, 512><<<n, 512, smem2, st>>>(Pc, n, ids, (const float2*)points, counts, (const float2*)origos, hints, flags, mode); \
    else if (t2 == 256) k_match2<STRICT_, 256><<<n, 256, smem2, st>>>(Pc, n, ids, (const float2*)points, counts, (const float2*)origos, hints, flags, mode); \
    else k_match2<STRICT_, 128><<<n, 128, smem2, st>>>(Pc, n, ids, (const float2*)points, counts, (const float2*)origos, hints, flags, mode); \
  } while (0)
      if (e->P.strict) HS_LAUNCH_MATCH2(true); else HS_LAUNCH_MATCH2(false);
#undef HS_LAUNCH_MATCH2
    } else {
    int ppt = (e->P.max_points + threads - 1) / threads;   // register-resident points (+ cell caches) per thread
    ppt = ppt <= 2 ? 2 : (ppt <= 3 ? 3 : (ppt <= 6 ? 6 : 0));
    const bool wc = e->match_variant == 3 && ppt > 0;   // cache misses of a warp processed densely (hs_points_phase_wc)
    const size_t smem = (size_t)hs_match2_pitch(e->P.max_points) * HS_NPROD * sizeof(float) +
                        (wc ? (size_t)(threads / 32) * ppt * 160 * sizeof(float) + 32 * sizeof(uint64_t) : 0);
#define HS_LAUNCH_MATCH(STRICT_, PPT_)                                                               \
  do {                                                                                               \
    if (wc) k_match<STRICT_, PPT_, (PPT_ > 0 ? 1 : 0)><<<n, threads, smem, st>>>(Pc, n, ids, (const float2*)points, counts, \
                                                        (const float2*)origos, hints, flags, mode); \
    else k_match<STRICT_, PPT_, 0><<<n, threads, smem, st>>>(Pc, n, ids, (const float2*)points, counts, \
                                                        (const float2*)origos, hints, flags, mode); \
  } while (0)
#define HS_LAUNCH_MATCH_S(STRICT_)                                    \
  do {                                                                \
    if (ppt == 2) HS_LAUNCH_MATCH(STRICT_, 2);                        \
    else if (ppt == 3 && wc && Pc.hyp_pblk)                                                                                        \
      k_match<STRICT_, 3, 2, 0><<<n, threads, smem, st>>>(Pc, n, ids, (const float2*)points, counts, (const float2*)origos, hints, flags, mode); \
    else if (ppt == 3 && wc && e->match_ct_pitch && hs_match2_pitch(e->P.max_points) == 388)                                      \
      k_match<STRICT_, 3, 1, 388><<<n, threads, smem, st>>>(Pc, n, ids, (const float2*)points, counts, (const float2*)origos, hints, flags, mode); \
    else if (ppt == 3) HS_LAUNCH_MATCH(STRICT_, 3);                   \
    else if (ppt == 6) HS_LAUNCH_MATCH(STRICT_, 6);                   \
    else HS_LAUNCH_MATCH(STRICT_, 0);                                 \
  } while (0)
    if (e->P.strict) HS_LAUNCH_MATCH_S(true); else HS_LAUNCH_MATCH_S(false);
#undef HS_LAUNCH_MATCH_S
#undef HS_LAUNCH_MATCH
    }
  }
  HS_LAUNCHED(e);
  if (ev_matched) HS_CK(cudaEventRecord(ev_matched, st));
  if (mode == 0) {
    e->map_epoch++;
    Pc.reverse_order = e->ray_alternate ? (int)(e->ray_launches++ & 1u) : 0;
    {
      ProfScope ps(e, HS_K_RAYCAST);
      int rc = launch_raycast(e, Pc, n, ids, (const float2*)points_ray, counts_ray, (const float2*)origos, st);
      if (rc != HS_OK) return rc;
    }
  }
  e->updates += (uint64_t)n;
  return HS_OK;
}

__global__ void k_gather_rows(const float* __restrict__ src, const int32_t* __restrict__ ids, int n, int width, float* __restrict__ dst) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n * width) return;
  int r = i / width, c = i - r * width;
  dst[i] = src[(size_t)ids[r] * width + c];
}

static int fetch_rows(hs_engine* e, const float* d_src, const int32_t* d_ids, int first, int n, int width, float* host_out) {
  if (!host_out || n == 0) return HS_OK;
  size_t bytes = (size_t)n * width * sizeof(float);
  if (d_ids) {
    k_gather_rows<<<(n * width + 255) / 256, 256, 0, e->stream>>>(d_src, d_ids, n, width, e->d_out);
    HS_LAUNCHED(e);
    HS_CK(cudaMemcpyAsync(host_out, e->d_out, bytes, cudaMemcpyDeviceToHost, e->stream));
  } else {
    HS_CK(cudaMemcpyAsync(host_out, d_src + (size_t)first * width, bytes, cudaMemcpyDeviceToHost, e->stream));
  }
  e->d2h_bytes += bytes;
  return HS_OK;
}

static int stage_common(hs_engine* e, int n, const int32_t* ids, const int32_t* counts, const float* origos,
                        const float* hints, const uint8_t* flags) {
  if (ids) { HS_CK(cudaMemcpyAsync(e->d_ids, ids, (size_t)n * sizeof(int32_t), cudaMemcpyHostToDevice, e->stream)); e->h2d_bytes += (size_t)n * 4; }
  if (counts) { HS_CK(cudaMemcpyAsync(e->d_counts, counts, (size_t)n * sizeof(int32_t), cudaMemcpyHostToDevice, e->stream)); e->h2d_bytes += (size_t)n * 4; }
  if (origos) { HS_CK(cudaMemcpyAsync(e->d_origos, origos, (size_t)n * 2 * sizeof(float), cudaMemcpyHostToDevice, e->stream)); e->h2d_bytes += (size_t)n * 8; }
  if (hints) { HS_CK(cudaMemcpyAsync(e->d_hints, hints, (size_t)n * 3 * sizeof(float), cudaMemcpyHostToDevice, e->stream)); e->h2d_bytes += (size_t)n * 12; }
  if (flags) { HS_CK(cudaMemcpyAsync(e->d_flags, flags, (size_t)n, cudaMemcpyHostToDevice, e->stream)); e->h2d_bytes += (size_t)n; }
  return HS_OK;
}

// stream ids of a batch must be in range and DISTINCT: two slots on one stream would race on its grids and counters
static int check_ids(hs_engine* e, int n, const int32_t* ids) {
  if (!ids) return HS_OK;
  std::vector<uint64_t>& seen = *e->id_bitmap;
  seen.assign(((size_t)e->P.n_streams + 63) / 64, 0ULL);
  for (int i = 0; i < n; ++i) {
    const int32_t s = ids[i];
    if (s < 0 || s >= e->P.n_streams) return HS_ERR_INVALID_ARG;
    const uint64_t bit = 1ULL << (s & 63);
    if (seen[(size_t)s >> 6] & bit) return HS_ERR_INVALID_ARG;
    seen[(size_t)s >> 6] |= bit;
  }
  return HS_OK;
}

int hs_update_batch_device(hs_engine* e, int n, const int32_t* stream_ids, const float* points, const int32_t* counts,
                           const float* origos, const float* pose_hints, const uint8_t* map_without_matching) {
  int st = check_batch(e, n);
  if (st != HS_OK) return st;
  if (n > 0 && (!points || !counts)) return HS_ERR_INVALID_ARG;
  HS_CK(cudaSetDevice(e->device));
  return run_update_device(e, n, stream_ids, points, counts, origos, pose_hints, map_without_matching, 0);
}

static int pipe_init(hs_engine* e);

// Copies `n16` 16-byte words of pinned host memory (mapped, read over PCIe) into device memory. One warp per CTA and at
// most 32 registers per thread, so that a CTA fits into the registers k_raycast leaves free on an SM (3 x 384 threads x 56
// registers of 64 K): launched on a high-priority stream it runs underneath the previous batch's ray-cast.
__global__ void __launch_bounds__(64, 32) k_prefetch_host(const float4* __restrict__ src, float4* __restrict__ dst, unsigned n16,
                                                          const int32_t* __restrict__ cnt_src, int32_t* __restrict__ cnt_dst, int n) {
  const unsigned stride = gridDim.x * blockDim.x;
  unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < (unsigned)n) cnt_dst[i] = cnt_src[i];
  for (; i + 3 * stride < n16; i += 4 * stride) {
    const float4 a = __ldcs(src + i), b = __ldcs(src + i + stride), c = __ldcs(src + i + 2 * stride), d = __ldcs(src + i + 3 * stride);
    dst[i] = a; dst[i + stride] = b; dst[i + 2 * stride] = c; dst[i + 3 * stride] = d;
  }
  for (; i < n16; i += stride) dst[i] = __ldcs(src + i);
}

static int update_host(hs_engine* e, int n, const int32_t* ids, const float* points, const int32_t* counts, const float* origos,
                       const float* hints, const uint8_t* flags, float* poses_out, float* cov_out, int mode) {
  int st = check_batch(e, n);
  if (st != HS_OK) return st;
  if (n == 0) return HS_OK;
  if (!points || !counts) return HS_ERR_INVALID_ARG;
  if ((st = check_ids(e, n, ids)) != HS_OK) return st;
  HS_CK(cudaSetDevice(e->device));
  const size_t row = (size_t)e->P.max_points * 2;   // floats per slot
  const size_t pbytes = (size_t)n * row * sizeof(float);
  // Is the caller's memory pinned (then it is mapped into the device's address space, UVA)?
  auto mapped = [](const void* p) -> const void* {
    cudaPointerAttributes attr;
    const bool ok = cudaPointerGetAttributes(&attr, p) == cudaSuccess && attr.type == cudaMemoryTypeHost && attr.devicePointer;
    cudaGetLastError();
    return ok ? attr.devicePointer : nullptr;
  };
  const float* m_points = (const float*)mapped(points);
  // counts pinned too: read zero-copy as well (the matcher then fetches whole rows, so that the point loads do not depend on
  // the count's round trip), which takes a small DMA copy and its latency off the front of the call
  const int32_t* m_counts = (m_points && e->e2e_zero_copy) ? (const int32_t*)mapped(counts) : nullptr;
  // poses_out pinned: the matcher writes the poses straight into it (no D2H copy behind the last kernel)
  float* m_poses = (m_points && e->e2e_zero_copy && poses_out && !ids) ? (float*)mapped(poses_out) : nullptr;
  if (mode == 0 && e->e2e_defer && m_points && m_counts && m_poses && !cov_out) {
    // Deferred integration: the call returns as soon as the poses are in the caller's buffer; the map integration of this
    // batch keeps running on the engine's stream, and every later call on the engine is ordered behind it. The next
    // call's points are then fetched into a second staging set WHILE that ray-cast runs (copy stream), so that a caller
    // looping over update() pays for the PCIe transfer only once, on the first call.
    if ((st = pipe_init(e)) != HS_OK) return st;
    if (!e->ev_match) HS_CK(cudaEventCreateWithFlags(&e->ev_match, cudaEventDisableTiming));
    hs_engine::PipeBuf& B = e->pipe[e->submits & 1];
    cudaStream_t cp = e->copy_stream;
    HS_CK(cudaStreamWaitEvent(cp, B.consumed, 0));
    if (e->e2e_defer == 2 && (n * row) % 4 == 0) {
      k_prefetch_host<<<2 * e->sm_count, 32, 0, cp>>>((const float4*)m_points, (float4*)B.points, (unsigned)(n * row / 4), m_counts, B.counts, n);
      HS_LAUNCHED(e);
    } else {
      HS_CK(cudaMemcpyAsync(B.points, points, pbytes, cudaMemcpyHostToDevice, cp));
      HS_CK(cudaMemcpyAsync(B.counts, counts, (size_t)n * sizeof(int32_t), cudaMemcpyHostToDevice, cp));
    }
    HS_CK(cudaEventRecord(B.staged, cp));
    if ((st = stage_common(e, n, nullptr, nullptr, origos, hints, flags)) != HS_OK) return st;
    HS_CK(cudaStreamWaitEvent(e->stream, B.staged, 0));
    st = run_update_device(e, n, nullptr, B.points, B.counts, origos ? e->d_origos : nullptr, hints ? e->d_hints : nullptr,
                           flags ? e->d_flags : nullptr, 0, e->stream, 0, nullptr, nullptr, m_poses, e->ev_match);
    if (st != HS_OK) return st;
    HS_CK(cudaEventRecord(B.consumed, e->stream));
    e->submits++;
    e->h2d_bytes += pbytes + (size_t)n * sizeof(int32_t);
    e->d2h_bytes += (size_t)n * 3 * sizeof(float);
    HS_CK(cudaEventSynchronize(e->ev_match));
    return HS_OK;
  }
  if ((st = stage_common(e, n, ids, m_counts ? nullptr : counts, origos, hints, flags)) != HS_OK) return st;
  // The points are the bulk of the input (2.9 KB per 360-beam scan).
  // Pinned host memory is mapped into the device's address space (UVA): the matcher CTAs then fetch their scans straight
  // from the caller's buffer over PCIe -- no staging copy to wait for, the transfer of one stream's points overlaps the
  // matching of the others -- and leave a device copy for the ray-caster. With HSGPU_E2E_ZERO_COPY=0 pinned batches are
  // instead copied in chunks on auxiliary streams (chunk c+1's copy under chunk c's kernels; measured slower on B200, see
  // DESIGN.md section 5). Pageable memory is copied in one piece (its copies are staged by the driver anyway).
  const int32_t* d_ids = ids ? e->d_ids : nullptr;
  const float* d_or = origos ? e->d_origos : nullptr;
  const float* d_hi = hints ? e->d_hints : nullptr;
  const uint8_t* d_fl = flags ? e->d_flags : nullptr;
  const bool pinned = m_points != nullptr;
  if (pinned && e->e2e_zero_copy) {
    // The batch still goes out as `chunks` launches on separate CUDA streams: the CTAs of the first chunk put their PCIe
    // reads on the wire first, so its scans are matched (and its ray-cast runs) while the later chunks' points still arrive.
    const int chunks = (e->e2e_chunks > 1 && n >= 64 * e->e2e_chunks) ? e->e2e_chunks : 1;
    if (chunks > 1) HS_CK(cudaEventRecord(e->ev_fork, e->stream));
    for (int c = 0; c < chunks; ++c) {
      const int off = (int)((long long)n * c / chunks), cnt = (int)((long long)n * (c + 1) / chunks) - off;
      cudaStream_t cs = c == 0 ? e->stream : e->aux_stream[c - 1];
      if (c > 0) HS_CK(cudaStreamWaitEvent(cs, e->ev_fork, 0));
      st = run_update_device(e, cnt, d_ids, m_points, m_counts ? m_counts : e->d_counts, d_or, d_hi, d_fl, mode, cs, off, e->d_points,
                             m_counts ? e->d_counts : nullptr, m_poses);
      if (st != HS_OK) return st;
      if (c > 0) {
        HS_CK(cudaEventRecord(e->ev_join[c - 1], cs));
        HS_CK(cudaStreamWaitEvent(e->stream, e->ev_join[c - 1], 0));
      }
    }
  } else {
    int chunks = 1;
    if (pinned && e->e2e_chunks > 1 && n >= 64 * e->e2e_chunks) chunks = e->e2e_chunks;
    if (chunks > 1) HS_CK(cudaEventRecord(e->ev_fork, e->stream));   // after the small arrays and everything queued before
    for (int c = 0; c < chunks; ++c) {
      // chunk boundaries: the first chunk may be smaller than the others (e2e_first_pct of an equal share), so that the
      // first kernels start after a shorter copy
      auto bound = [&](int k) -> int {
        if (k <= 0) return 0;
        if (k >= chunks) return n;
        const long long first = (long long)n * e->e2e_first_pct / (100LL * chunks);
        return (int)(first + ((long long)n - first) * (k - 1) / (chunks - 1));
      };
      const int off = bound(c), cnt = bound(c + 1) - off;
      if (cnt <= 0) continue;
      cudaStream_t cs = c == 0 ? e->stream : e->aux_stream[c - 1];
      if (c > 0) HS_CK(cudaStreamWaitEvent(cs, e->ev_fork, 0));
      HS_CK(cudaMemcpyAsync(e->d_points + off * row, points + off * row, (size_t)cnt * row * sizeof(float), cudaMemcpyHostToDevice, cs));
      st = run_update_device(e, cnt, d_ids, e->d_points, e->d_counts, d_or, d_hi, d_fl, mode, cs, off);
      if (st != HS_OK) return st;
      if (c > 0) {
        HS_CK(cudaEventRecord(e->ev_join[c - 1], cs));
        HS_CK(cudaStreamWaitEvent(e->stream, e->ev_join[c - 1], 0));
      }
    }
  }
  e->h2d_bytes += pbytes;
  if (pinned && e->e2e_zero_copy && m_poses) e->d2h_bytes += (size_t)n * 3 * sizeof(float);
  else if ((st = fetch_rows(e, e->P.last_pose, ids ? e->d_ids : nullptr, 0, n, 3, poses_out)) != HS_OK) return st;
  if (cov_out) {
    if (ids && poses_out) HS_CK(cudaStreamSynchronize(e->stream));  // d_out is reused by the second gather
    if ((st = fetch_rows(e, e->P.cov, ids ? e->d_ids : nullptr, 0, n, 9, cov_out)) != HS_OK) return st;
  }
  HS_CK(cudaStreamSynchronize(e->stream));
  return HS_OK;
}

int hs_update_batch(hs_engine* e, int n, const int32_t* stream_ids, const float* points, const int32_t* counts,
                    const float* origos, const float* pose_hints, const uint8_t* map_without_matching,
                    float* poses_out, float* cov_out) {
  return update_host(e, n, stream_ids, points, counts, origos, pose_hints, map_without_matching, poses_out, cov_out, 0);
}

int hs_match_batch(hs_engine* e, int n, const int32_t* stream_ids, const float* points, const int32_t* counts,
                   const float* origos, const float* pose_hints, float* poses_out, float* cov_out) {
  return update_host(e, n, stream_ids, points, counts, origos, pose_hints, nullptr, poses_out, cov_out, 1);
}

// ---- pipelined submission ---------------------------------------------------------------------------

static int pipe_init(hs_engine* e) {
  if (e->copy_stream) return HS_OK;
  const size_t ns = (size_t)e->P.n_streams, mp = (size_t)e->P.max_points;
  int prio_lo = 0, prio_hi = 0;
  HS_CK(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
  HS_CK(cudaStreamCreateWithPriority(&e->copy_stream, cudaStreamNonBlocking, prio_hi));
  for (int b = 0; b < 2; ++b) {
    hs_engine::PipeBuf& B = e->pipe[b];
    HS_CK(cudaMalloc((void**)&B.points, ns * mp * 2 * sizeof(float)));
    HS_CK(cudaMalloc((void**)&B.ranges, ns * mp * sizeof(float)));   // n_beams <= max_points
    HS_CK(cudaMalloc((void**)&B.counts, ns * sizeof(int32_t)));
    HS_CK(cudaMalloc((void**)&B.origos, ns * 2 * sizeof(float)));
    HS_CK(cudaMalloc((void**)&B.hints, ns * 3 * sizeof(float)));
    HS_CK(cudaMalloc((void**)&B.flags, ns * sizeof(uint8_t)));
    HS_CK(cudaMalloc((void**)&B.ids, ns * sizeof(int32_t)));
    HS_CK(cudaEventCreateWithFlags(&B.staged, cudaEventDisableTiming));
    HS_CK(cudaEventCreateWithFlags(&B.consumed, cudaEventDisableTiming));
  }
  return HS_OK;
}

// update() for n streams WITHOUT waiting for the result: the inputs are copied by DMA on a copy stream into one of two
// staging sets -- so batch k+1's transfer runs while batch k is matched and ray-cast -- and the kernels are queued behind it.
// All host buffers must be pinned and stay untouched until hs_wait(); poses_out (pinned, or NULL) is written by the
// matcher itself. Pose hints default to each stream's last pose on the device, so consecutive scans of the same streams
// can be submitted back to back. Anything else falls back to the synchronous hs_update_batch.
int hs_submit_batch(hs_engine* e, int n, const int32_t* stream_ids, const float* points, const int32_t* counts, const float* origos,
                    const float* pose_hints, const uint8_t* map_without_matching, float* poses_out) {
  int st = check_batch(e, n);
  if (st != HS_OK) return st;
  if (n == 0) return HS_OK;
  if (!points || !counts) return HS_ERR_INVALID_ARG;
  if ((st = check_ids(e, n, stream_ids)) != HS_OK) return st;
  HS_CK(cudaSetDevice(e->device));
  auto mapped = [](const void* p) -> void* {
    cudaPointerAttributes attr;
    const bool ok = p && cudaPointerGetAttributes(&attr, p) == cudaSuccess && attr.type == cudaMemoryTypeHost && attr.devicePointer;
    cudaGetLastError();
    return ok ? attr.devicePointer : nullptr;
  };
  float* m_poses = (float*)mapped(poses_out);
  const bool all_pinned = mapped(points) && mapped(counts) && (!poses_out || m_poses) && (!stream_ids || mapped(stream_ids)) &&
                          (!origos || mapped(origos)) && (!pose_hints || mapped(pose_hints)) && (!map_without_matching || mapped(map_without_matching));
  if (!all_pinned) return hs_update_batch(e, n, stream_ids, points, counts, origos, pose_hints, map_without_matching, poses_out, nullptr);
  if ((st = pipe_init(e)) != HS_OK) return st;
  hs_engine::PipeBuf& B = e->pipe[e->submits & 1];
  const size_t row = (size_t)e->P.max_points * 2;
  cudaStream_t cp = e->copy_stream;
  HS_CK(cudaStreamWaitEvent(cp, B.consumed, 0));   // the batch that last used this staging set has been computed (no-op the first time)
  HS_CK(cudaMemcpyAsync(B.points, points, (size_t)n * row * sizeof(float), cudaMemcpyHostToDevice, cp));
  HS_CK(cudaMemcpyAsync(B.counts, counts, (size_t)n * sizeof(int32_t), cudaMemcpyHostToDevice, cp));
  size_t bytes = (size_t)n * (row * sizeof(float) + 4);
  if (stream_ids) { HS_CK(cudaMemcpyAsync(B.ids, stream_ids, (size_t)n * 4, cudaMemcpyHostToDevice, cp)); bytes += (size_t)n * 4; }
  if (origos) { HS_CK(cudaMemcpyAsync(B.origos, origos, (size_t)n * 8, cudaMemcpyHostToDevice, cp)); bytes += (size_t)n * 8; }
  if (pose_hints) { HS_CK(cudaMemcpyAsync(B.hints, pose_hints, (size_t)n * 12, cudaMemcpyHostToDevice, cp)); bytes += (size_t)n * 12; }
  if (map_without_matching) { HS_CK(cudaMemcpyAsync(B.flags, map_without_matching, (size_t)n, cudaMemcpyHostToDevice, cp)); bytes += (size_t)n; }
  e->h2d_bytes += bytes;
  HS_CK(cudaEventRecord(B.staged, cp));
  HS_CK(cudaStreamWaitEvent(e->stream, B.staged, 0));
  st = run_update_device(e, n, stream_ids ? B.ids : nullptr, B.points, B.counts, origos ? B.origos : nullptr, pose_hints ? B.hints : nullptr,
                         map_without_matching ? B.flags : nullptr, 0, e->stream, 0, nullptr, nullptr, stream_ids ? nullptr : m_poses);
  if (st != HS_OK) return st;
  if (m_poses && stream_ids) {   // poses of a stream subset: gathered on the device, then written to the pinned buffer
    k_gather_rows<<<(n * 3 + 255) / 256, 256, 0, e->stream>>>(e->P.last_pose, B.ids, n, 3, m_poses);
    HS_LAUNCHED(e);
  }
  if (m_poses) e->d2h_bytes += (size_t)n * 12;
  HS_CK(cudaEventRecord(B.consumed, e->stream));
  e->submits++;
  return HS_OK;
}

// Waits until everything submitted so far (and every other call on this engine) has completed and its outputs are in place.
int hs_wait(hs_engine* e) { return hs_synchronize(e); }

int hs_set_scan_geometry(hs_engine* e, int n_beams, float angle_min, float angle_increment, float range_min, float range_max) {
  if (!e || n_beams < 1 || n_beams > e->P.max_points) return HS_ERR_INVALID_ARG;
  HS_CK(cudaSetDevice(e->device));
  HS_CK(cudaStreamSynchronize(e->stream));
  if (e->d_beam_cs) { cudaFree(e->d_beam_cs); e->d_beam_cs = nullptr; }
  if (e->d_ranges) { cudaFree(e->d_ranges); e->d_ranges = nullptr; }
  // rosLaserScanToDataContainer (HectorMappingRos.cpp:487, :505): float accumulation of the beam angle;
  // cos/sin from the host libm, as the reference's caller computes them.
  std::vector<float2> cs((size_t)n_beams);
  volatile float angle = angle_min;
  for (int i = 0; i < n_beams; ++i) {
    float a = angle;
    cs[i] = make_float2(cosf(a), sinf(a));
    angle = angle + angle_increment;
  }
  HS_CK(cudaMalloc((void**)&e->d_beam_cs, (size_t)n_beams * sizeof(float2)));
  HS_CK(cudaMalloc((void**)&e->d_ranges, (size_t)e->P.n_streams * n_beams * sizeof(float)));
  HS_CK(cudaMemcpy(e->d_beam_cs, cs.data(), (size_t)n_beams * sizeof(float2), cudaMemcpyHostToDevice));
  e->n_beams = n_beams;
  e->angle_min = angle_min;
  e->angle_increment = angle_increment;
  e->range_min = range_min;
  e->range_max = range_max;
  volatile float mr = range_max - 0.1f;
  e->max_range_for_container = mr;
  return HS_OK;
}

int hs_update_batch_ranges_device(hs_engine* e, int n, const int32_t* stream_ids, const float* ranges,
                                  const float* pose_hints, const uint8_t* map_without_matching) {
  int st = check_batch(e, n);
  if (st != HS_OK) return st;
  if (!e->d_beam_cs) return HS_ERR_NOT_CONFIGURED;
  if (n == 0) return HS_OK;
  if (!ranges) return HS_ERR_INVALID_ARG;
  HS_CK(cudaSetDevice(e->device));
  const int threads = 128;
  {
    ProfScope ps(e, HS_K_SCAN_TO_POINTS);
    k_scan_to_points<<<warp_grid(n, threads), threads, 0, e->stream>>>(ranges, n, e->n_beams, e->d_beam_cs, e->range_min,
                                                                       e->max_range_for_container, e->P.lv[0].scale,
                                                                       e->P.max_points, (float2*)e->d_points, e->d_counts);
  }
  HS_LAUNCHED(e);
  return run_update_device(e, n, stream_ids, e->d_points, e->d_counts, nullptr, pose_hints, map_without_matching, 0);
}

int hs_update_batch_ranges(hs_engine* e, int n, const int32_t* stream_ids, const float* ranges, const float* pose_hints,
                           const uint8_t* map_without_matching, float* poses_out, float* cov_out) {
  int st = check_batch(e, n);
  if (st != HS_OK) return st;
  if (!e->d_beam_cs) return HS_ERR_NOT_CONFIGURED;
  if (n == 0) return HS_OK;
  if (!ranges) return HS_ERR_INVALID_ARG;
  if ((st = check_ids(e, n, stream_ids)) != HS_OK) return st;
  HS_CK(cudaSetDevice(e->device));
  size_t rbytes = (size_t)n * e->n_beams * sizeof(float);
  if (e->e2e_defer && e->e2e_zero_copy && !stream_ids && !cov_out && poses_out) {   // deferred integration, see update_host
    cudaPointerAttributes ar, ap;
    const bool pinned = cudaPointerGetAttributes(&ar, ranges) == cudaSuccess && ar.type == cudaMemoryTypeHost && ar.devicePointer &&
                        cudaPointerGetAttributes(&ap, poses_out) == cudaSuccess && ap.type == cudaMemoryTypeHost && ap.devicePointer;
    cudaGetLastError();
    if (pinned) {
      if ((st = pipe_init(e)) != HS_OK) return st;
      if (!e->ev_match) HS_CK(cudaEventCreateWithFlags(&e->ev_match, cudaEventDisableTiming));
      hs_engine::PipeBuf& B = e->pipe[e->submits & 1];
      HS_CK(cudaStreamWaitEvent(e->copy_stream, B.consumed, 0));
      HS_CK(cudaMemcpyAsync(B.ranges, ranges, rbytes, cudaMemcpyHostToDevice, e->copy_stream));
      HS_CK(cudaEventRecord(B.staged, e->copy_stream));
      if ((st = stage_common(e, n, nullptr, nullptr, nullptr, pose_hints, map_without_matching)) != HS_OK) return st;
      HS_CK(cudaStreamWaitEvent(e->stream, B.staged, 0));
      {
        ProfScope ps(e, HS_K_SCAN_TO_POINTS);
        k_scan_to_points<<<warp_grid(n, 128), 128, 0, e->stream>>>(B.ranges, n, e->n_beams, e->d_beam_cs, e->range_min, e->max_range_for_container,
                                                                   e->P.lv[0].scale, e->P.max_points, (float2*)B.points, B.counts);
      }
      HS_LAUNCHED(e);
      st = run_update_device(e, n, nullptr, B.points, B.counts, nullptr, pose_hints ? e->d_hints : nullptr,
                             map_without_matching ? e->d_flags : nullptr, 0, e->stream, 0, nullptr, nullptr, (float*)ap.devicePointer, e->ev_match);
      if (st != HS_OK) return st;
      HS_CK(cudaEventRecord(B.consumed, e->stream));
      e->submits++;
      e->h2d_bytes += rbytes;
      e->d2h_bytes += (size_t)n * 3 * sizeof(float);
      HS_CK(cudaEventSynchronize(e->ev_match));
      return HS_OK;
    }
  }
  const float* d_src = e->d_ranges;
  {   // pinned host memory: k_scan_to_points reads the ranges straight from the caller's buffer (zero-copy, see update_host)
    cudaPointerAttributes attr;
    if (e->e2e_zero_copy && cudaPointerGetAttributes(&attr, ranges) == cudaSuccess && attr.type == cudaMemoryTypeHost && attr.devicePointer)
      d_src = (const float*)attr.devicePointer;
    cudaGetLastError();
  }
  if (d_src == e->d_ranges) HS_CK(cudaMemcpyAsync(e->d_ranges, ranges, rbytes, cudaMemcpyHostToDevice, e->stream));
  e->h2d_bytes += rbytes;
  if ((st = stage_common(e, n, stream_ids, nullptr, nullptr, pose_hints, map_without_matching)) != HS_OK) return st;
  st = hs_update_batch_ranges_device(e, n, stream_ids ? e->d_ids : nullptr, d_src, pose_hints ? e->d_hints : nullptr,
                                     map_without_matching ? e->d_flags : nullptr);
  if (st != HS_OK) return st;
  if ((st = fetch_rows(e, e->P.last_pose, stream_ids ? e->d_ids : nullptr, 0, n, 3, poses_out)) != HS_OK) return st;
  if (cov_out) {
    if (stream_ids && poses_out) HS_CK(cudaStreamSynchronize(e->stream));
    if ((st = fetch_rows(e, e->P.cov, stream_ids ? e->d_ids : nullptr, 0, n, 9, cov_out)) != HS_OK) return st;
  }
  HS_CK(cudaStreamSynchronize(e->stream));
  return HS_OK;
}

// ---- point-cloud ingestion (the node's default path, HectorMappingRos.cpp:273-282, :509-542) -------------------

int hs_cloud_to_points_host(const float* cloud_xyz, int n_cloud, const double* transform12, float scale_to_map,
                            float sqr_min_dist, float sqr_max_dist, float z_min, float z_max, int max_points,
                            float* points_out, float* origo_out) {
  if (n_cloud < 0 || (!cloud_xyz && n_cloud > 0) || !transform12 || !points_out || max_points < 0) return HS_ERR_INVALID_ARG;
  const HsCloudFilter F = {sqr_min_dist, sqr_max_dist, z_min, z_max};
  int count = 0;
  for (int i = 0; i < n_cloud; ++i) {
    float px, py;
    if (hs_cloud_point(cloud_xyz + 3 * (size_t)i, transform12, F, scale_to_map, px, py)) {
      if (count < max_points) { points_out[2 * count] = px; points_out[2 * count + 1] = py; }
      ++count;
    }
  }
  if (origo_out) { origo_out[0] = (float)transform12[9] * scale_to_map; origo_out[1] = (float)transform12[10] * scale_to_map; }
  return count < max_points ? count : max_points;
}

int hs_update_batch_clouds(hs_engine* e, int n, const int32_t* stream_ids, const float* clouds_xyz, const int32_t* cloud_counts,
                           int max_cloud, const double* transforms12, float sqr_min_dist, float sqr_max_dist, float z_min, float z_max,
                           const float* pose_hints, const uint8_t* map_without_matching, float* poses_out, float* cov_out) {
  int st = check_batch(e, n);
  if (st != HS_OK) return st;
  if (n == 0) return HS_OK;
  if (!clouds_xyz || !cloud_counts || !transforms12 || max_cloud < 1) return HS_ERR_INVALID_ARG;
  if ((st = check_ids(e, n, stream_ids)) != HS_OK) return st;
  HS_CK(cudaSetDevice(e->device));
  auto align = [](size_t x) { return (x + 255) & ~(size_t)255; };
  const size_t b_c = (size_t)n * max_cloud * 3 * sizeof(float), b_n = (size_t)n * sizeof(int32_t), b_t = (size_t)n * 12 * sizeof(double);
  const size_t o_n = align(b_c), o_t = o_n + align(b_n);
  if ((st = ensure_scratch(e, o_t + b_t)) != HS_OK) return st;
  char* base = (char*)e->d_scratch;
  HS_CK(cudaMemcpyAsync(base, clouds_xyz, b_c, cudaMemcpyHostToDevice, e->stream));
  HS_CK(cudaMemcpyAsync(base + o_n, cloud_counts, b_n, cudaMemcpyHostToDevice, e->stream));
  HS_CK(cudaMemcpyAsync(base + o_t, transforms12, b_t, cudaMemcpyHostToDevice, e->stream));
  e->h2d_bytes += b_c + b_n + b_t;
  if ((st = stage_common(e, n, stream_ids, nullptr, nullptr, pose_hints, map_without_matching)) != HS_OK) return st;
  const HsCloudFilter F = {sqr_min_dist, sqr_max_dist, z_min, z_max};
  {
    ProfScope ps(e, HS_K_SCAN_TO_POINTS);
    k_cloud_to_points<<<warp_grid(n, 128), 128, 0, e->stream>>>((const float*)base, (const int32_t*)(base + o_n), n, max_cloud,
                                                                (const double*)(base + o_t), F, e->P.lv[0].scale, e->P.max_points,
                                                                (float2*)e->d_points, e->d_counts, (float2*)e->d_origos);
  }
  HS_LAUNCHED(e);
  const int32_t* d_ids = stream_ids ? e->d_ids : nullptr;
  st = run_update_device(e, n, d_ids, e->d_points, e->d_counts, e->d_origos, pose_hints ? e->d_hints : nullptr,
                         map_without_matching ? e->d_flags : nullptr, 0);
  if (st != HS_OK) return st;
  if ((st = fetch_rows(e, e->P.last_pose, d_ids, 0, n, 3, poses_out)) != HS_OK) return st;
  if (cov_out) {
    if (stream_ids && poses_out) HS_CK(cudaStreamSynchronize(e->stream));
    if ((st = fetch_rows(e, e->P.cov, d_ids, 0, n, 9, cov_out)) != HS_OK) return st;
  }
  HS_CK(cudaStreamSynchronize(e->stream));
  return HS_OK;
}

static int get_rows(hs_engine* e, const float* d_src, int first, int n, int width, float* out) {
  if (!e || !out || first < 0 || n < 0 || first + n > e->P.n_streams) return HS_ERR_INVALID_ARG;
  HS_CK(cudaSetDevice(e->device));
  int st = fetch_rows(e, d_src, nullptr, first, n, width, out);
  if (st != HS_OK) return st;
  HS_CK(cudaStreamSynchronize(e->stream));
  return HS_OK;
}
int hs_get_poses(hs_engine* e, int first, int n, float* poses_out) { return get_rows(e, e ? e->P.last_pose : nullptr, first, n, 3, poses_out); }
int hs_get_covariances(hs_engine* e, int first, int n, float* cov_out) { return get_rows(e, e ? e->P.cov : nullptr, first, n, 9, cov_out); }
int hs_get_last_map_update_poses(hs_engine* e, int first, int n, float* poses_out) {
  return get_rows(e, e ? e->P.last_update_pose : nullptr, first, n, 3, poses_out);
}
const float* hs_device_pose_ptr(const hs_engine* e) { return e ? e->P.last_pose : nullptr; }

// ---- single-function entry points --------------------------------------------------------------

// The probability block planes of one stream, all levels (see k_build_pblk): built on the engine's stream when first needed
// and kept until the stream's maps may have changed (e->map_epoch is bumped by every entry point that writes cells).
static int ensure_pblk(hs_engine* e, int stream) {
  if (e->pblk && e->pblk_stream == stream && e->pblk_epoch == e->map_epoch) return HS_OK;
  if (!e->pblk) {
    cudaError_t err = cudaMalloc((void**)&e->pblk, (size_t)e->P.cells_per_stream * sizeof(float4));
    if (err != cudaSuccess) { cudaGetLastError(); e->pblk = nullptr; return err == cudaErrorMemoryAllocation ? HS_ERR_OUT_OF_MEMORY : (int)err; }
  }
  const HsLevel& L0 = e->P.lv[0];
  dim3 grid((L0.sx + HS_PB_TX - 1) / HS_PB_TX, (L0.sy + HS_PB_TY - 1) / HS_PB_TY, e->P.n_levels);
  {
    ProfScope ps(e, HS_K_BUILD_PBLK);
    k_build_pblk<<<grid, HS_PB_TX * HS_PB_TY, 0, e->stream>>>(e->P, stream, e->pblk);
  }
  HS_LAUNCHED(e);
  e->pblk_stream = stream;
  e->pblk_epoch = e->map_epoch;
  return HS_OK;
}

int hs_hessian_derivs_batch_device(hs_engine* e, int stream, int level, const float* poses_map, int n_poses, const float* points,
                                   int n_points, float* H9, float* dTr3) {
  if (!e || stream < 0 || stream >= e->P.n_streams || level < 0 || level >= e->P.n_levels || n_poses < 0 || n_points < 0)
    return HS_ERR_INVALID_ARG;
  if (n_poses == 0) return HS_OK;
  if (!poses_map || (!points && n_points > 0) || !H9 || !dTr3) return HS_ERR_INVALID_ARG;
  HS_CK(cudaSetDevice(e->device));
  int st = ensure_pblk(e, stream);
  if (st != HS_OK) return st;
  // poses per warp: 3 fills 27 of the chain phase's 32 lanes and divides the serial part's issue slots by three, but the
  // kernel is latency-bound until every SM holds its 64 warps: measured on B200 at 4096 hypotheses x 360 points (strict /
  // tree, us per launch): 1 pose per warp 19.0 / 14.3, 2: 20.8 / 16.8, 3: 23.1 / 18.9
  int ppw = n_poses >= 3 * 64 * e->sm_count ? 3 : (n_poses >= 2 * 64 * e->sm_count ? 2 : 1);
  if (e->hd_ppw >= 1 && e->hd_ppw <= 3) ppw = e->hd_ppw;
  const int warps_per_cta = HS_HD_THREADS / 32;
  const size_t smem = ((size_t)2 * ((n_points + 1) & ~1) + (size_t)warps_per_cta * ppw * HS_NPROD * HS_HD_PITCH) * sizeof(float);
  if (smem > 200 * 1024) return HS_ERR_INVALID_ARG;
  const int n_groups = (n_poses + ppw - 1) / ppw;
  int grid = (n_groups + warps_per_cta - 1) / warps_per_cta;
  const int max_grid = e->sm_count * 16;
  if (grid > max_grid) grid = max_grid;
  const HsLevel& L = e->P.lv[level];
  ProfScope ps(e, HS_K_HESSIAN);
#define HS_LAUNCH_HD(STRICT_, PPW_)                                                                                                     \
  do {                                                                                                                                  \
    if (smem > 48 * 1024) HS_CK(cudaFuncSetAttribute(k_hessian_derivs<STRICT_, PPW_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    k_hessian_derivs<STRICT_, PPW_><<<grid, HS_HD_THREADS, smem, e->stream>>>(e->pblk + L.cell_offset, L, poses_map, n_poses, (const float2*)points,    \
                                                                             n_points, H9, dTr3);                                      \
  } while (0)
  if (e->P.strict) {
    if (ppw == 3) HS_LAUNCH_HD(true, 3); else if (ppw == 2) HS_LAUNCH_HD(true, 2); else HS_LAUNCH_HD(true, 1);
  } else {
    if (ppw == 3) HS_LAUNCH_HD(false, 3); else if (ppw == 2) HS_LAUNCH_HD(false, 2); else HS_LAUNCH_HD(false, 1);
  }
#undef HS_LAUNCH_HD
  HS_LAUNCHED(e);
  return HS_OK;
}

static int ensure_pinned(hs_engine* e, size_t bytes) {
  if (bytes <= e->pinned_bytes) return HS_OK;
  if (e->h_pinned) cudaFreeHost(e->h_pinned);
  e->h_pinned = nullptr;
  e->pinned_bytes = 0;
  HS_CK(cudaHostAlloc(&e->h_pinned, bytes, cudaHostAllocDefault));
  e->pinned_bytes = bytes;
  return HS_OK;
}

int hs_hessian_derivs_batch(hs_engine* e, int stream, int level, const float* poses_map, int n_poses, const float* points,
                            int n_points, float* H9, float* dTr3) {
  if (!e || stream < 0 || stream >= e->P.n_streams || level < 0 || level >= e->P.n_levels || n_poses < 0 || n_points < 0)
    return HS_ERR_INVALID_ARG;
  if (n_poses == 0) return HS_OK;
  if (!poses_map || (!points && n_points > 0) || !H9 || !dTr3) return HS_ERR_INVALID_ARG;
  HS_CK(cudaSetDevice(e->device));
  // One staging block each way, mirrored in an engine-owned PINNED host buffer: poses | points in, H | dTr out. The caller's
  // (usually pageable) arrays are memcpy'd to / from it, so the call costs two DMA transfers instead of four driver-staged
  // copies (measured on B200, 4096 hypotheses x 360 points: 95 -> see DESIGN.md us per call).
  size_t b_poses = (size_t)n_poses * 3 * sizeof(float), b_pts = (size_t)(n_points > 0 ? n_points : 1) * 2 * sizeof(float);
  size_t b_h = (size_t)n_poses * 9 * sizeof(float), b_d = (size_t)n_poses * 3 * sizeof(float);
  auto align = [](size_t x) { return (x + 255) & ~(size_t)255; };
  size_t o_pts = align(b_poses), o_h = o_pts + align(b_pts), o_d = o_h + align(b_h), total = o_d + align(b_d);
  int st = ensure_scratch(e, total);
  if (st != HS_OK) return st;
  if ((st = ensure_pinned(e, total)) != HS_OK) return st;
  char* base = (char*)e->d_scratch;
  char* hbase = (char*)e->h_pinned;
  memcpy(hbase, poses_map, b_poses);
  if (n_points > 0) memcpy(hbase + o_pts, points, (size_t)n_points * 2 * sizeof(float));
  HS_CK(cudaMemcpyAsync(base, hbase, o_pts + (size_t)n_points * 2 * sizeof(float), cudaMemcpyHostToDevice, e->stream));
  e->h2d_bytes += b_poses + (size_t)n_points * 8;
  st = hs_hessian_derivs_batch_device(e, stream, level, (const float*)base, n_poses, (const float*)(base + o_pts), n_points,
                                      (float*)(base + o_h), (float*)(base + o_d));
  if (st != HS_OK) return st;
  HS_CK(cudaMemcpyAsync(hbase + o_h, base + o_h, (o_d - o_h) + b_d, cudaMemcpyDeviceToHost, e->stream));
  e->d2h_bytes += b_h + b_d;
  HS_CK(cudaStreamSynchronize(e->stream));
  memcpy(H9, hbase + o_h, b_h);
  memcpy(dTr3, hbase + o_d, b_d);
  return HS_OK;
}

// Relocalisation-style multi-start matching (BASELINE configs[3]): MapRepMultiMap::matchData (HSL/slam_main/MapRepMultiMap.h:116-132)
// of ONE scan against ONE stream's maps from n_hyp different pose hints at once. Nothing of the stream's state is touched
// (no pose, no coarse containers, no integration). poses_out [n_hyp][3] (world), cov_out [n_hyp][9] (the last level-0 Hessian,
// as getLastScanMatchCovariance would return); either may be NULL.
int hs_match_hypotheses(hs_engine* e, int stream, const float* points, int n_points, const float* pose_hints, int n_hyp,
                        float* poses_out, float* cov_out) {
  if (!e || stream < 0 || stream >= e->P.n_streams || n_points < 0 || n_points > e->P.max_points || n_hyp < 0) return HS_ERR_INVALID_ARG;
  if (n_hyp == 0) return HS_OK;
  if (!pose_hints || (!points && n_points > 0)) return HS_ERR_INVALID_ARG;
  HS_CK(cudaSetDevice(e->device));
  auto align = [](size_t x) { return (x + 255) & ~(size_t)255; };
  const size_t b_pts = (size_t)e->P.max_points * 2 * sizeof(float), b_h = (size_t)n_hyp * 3 * sizeof(float), b_c = (size_t)n_hyp * 9 * sizeof(float);
  const size_t o_cnt = align(b_pts), o_hint = o_cnt + 256, o_pose = o_hint + align(b_h), o_cov = o_pose + align(b_h);
  int st = ensure_scratch(e, o_cov + b_c);
  if (st != HS_OK) return st;
  char* base = (char*)e->d_scratch;
  const int32_t cnt = n_points;
  if (n_points > 0) HS_CK(cudaMemcpyAsync(base, points, (size_t)n_points * 2 * sizeof(float), cudaMemcpyHostToDevice, e->stream));
  HS_CK(cudaMemcpyAsync(base + o_cnt, &cnt, sizeof(cnt), cudaMemcpyHostToDevice, e->stream));
  HS_CK(cudaMemcpyAsync(base + o_hint, pose_hints, b_h, cudaMemcpyHostToDevice, e->stream));
  HS_CK(cudaMemsetAsync(base + o_cov, 0, b_c, e->stream));   // an empty scan leaves the covariance untouched (ScanMatcher.h:189)
  e->h2d_bytes += (size_t)n_points * 8 + 4 + b_h;
  // thousands of hypotheses read one stream's maps: the register-cache matcher takes its 2x2 blocks from the stream's
  // probability block planes (one conversion per cell for all of them, ensure_pblk) -- 360-point-class scans only
  const bool use_pblk = e->match_variant == 3 && (e->P.max_points + HS_MATCH_THREADS - 1) / HS_MATCH_THREADS == 3 && n_hyp >= 64 && e->hyp_use_pblk;
  if (use_pblk && (st = ensure_pblk(e, stream)) != HS_OK) return st;
  const int saved_stream = e->P.hyp_stream;
  e->P.hyp_pblk = use_pblk ? e->pblk : nullptr;
  e->P.hyp_pose = (float*)(base + o_pose);
  e->P.hyp_cov = (float*)(base + o_cov);
  e->P.hyp_stream = stream;
  st = run_update_device(e, n_hyp, nullptr, (const float*)base, (const int32_t*)(base + o_cnt), nullptr, (const float*)(base + o_hint), nullptr, 1);
  e->P.hyp_pose = nullptr;
  e->P.hyp_cov = nullptr;
  e->P.hyp_pblk = nullptr;
  e->P.hyp_stream = saved_stream;
  e->updates -= (uint64_t)n_hyp;   // hypotheses are not stream updates
  if (st != HS_OK) return st;
  if (poses_out) { HS_CK(cudaMemcpyAsync(poses_out, base + o_pose, b_h, cudaMemcpyDeviceToHost, e->stream)); e->d2h_bytes += b_h; }
  if (cov_out) { HS_CK(cudaMemcpyAsync(cov_out, base + o_cov, b_c, cudaMemcpyDeviceToHost, e->stream)); e->d2h_bytes += b_c; }
  HS_CK(cudaStreamSynchronize(e->stream));
  return HS_OK;
}

static int integrate_host(hs_engine* e, int n, const int32_t* stream_ids, const float* points, const int32_t* counts,
                          const float* origos, const float* poses_world, int refresh_coarse) {
  int st = check_batch(e, n);
  if (st != HS_OK) return st;
  if (n == 0) return HS_OK;
  if (!points || !counts || !poses_world) return HS_ERR_INVALID_ARG;
  if ((st = check_ids(e, n, stream_ids)) != HS_OK) return st;
  HS_CK(cudaSetDevice(e->device));
  size_t pbytes = (size_t)n * e->P.max_points * 2 * sizeof(float);
  HS_CK(cudaMemcpyAsync(e->d_points, points, pbytes, cudaMemcpyHostToDevice, e->stream));
  e->h2d_bytes += pbytes;
  if ((st = stage_common(e, n, stream_ids, counts, origos, poses_world, nullptr)) != HS_OK) return st;
  const int32_t* d_ids = stream_ids ? e->d_ids : nullptr;
  const float2* d_or = origos ? (const float2*)e->d_origos : nullptr;
  const int threads = 128;
  e->map_epoch++;
  k_force_integrate<<<warp_grid(n, threads), threads, 0, e->stream>>>(e->P, n, d_ids, (const float2*)e->d_points, e->d_counts, d_or, e->d_hints,
                                                                      refresh_coarse);
  HS_LAUNCHED(e);
  {
    ProfScope ps(e, HS_K_RAYCAST);
    int rc = launch_raycast(e, e->P, n, d_ids, (const float2*)e->d_points, e->d_counts, d_or, e->stream);
    if (rc != HS_OK) return rc;
  }
  HS_CK(cudaStreamSynchronize(e->stream));
  return HS_OK;
}

int hs_raycast_batch(hs_engine* e, int n, const int32_t* stream_ids, const float* points, const int32_t* counts,
                     const float* origos, const float* poses_world) {
  return integrate_host(e, n, stream_ids, points, counts, origos, poses_world, 1);
}

int hs_integrate_batch(hs_engine* e, int n, const int32_t* stream_ids, const float* points, const int32_t* counts,
                       const float* origos, const float* poses_world) {
  return integrate_host(e, n, stream_ids, points, counts, origos, poses_world, 0);
}

// ---- off-path OccGridMapUtil members, occupancy ray queries, stream state ------------------------

static int check_stream_level(const hs_engine* e, int stream, int level) {
  if (!e || stream < 0 || stream >= e->P.n_streams || level < 0 || level >= e->P.n_levels) return HS_ERR_INVALID_ARG;
  return HS_OK;
}

int hs_likelihood_batch(hs_engine* e, int stream, int level, const float* poses_map, int n_poses, const float* points,
                        int n_points, float* residual_out, float* likelihood_out) {
  int st = check_stream_level(e, stream, level);
  if (st != HS_OK) return st;
  if (n_poses < 0 || n_points < 0) return HS_ERR_INVALID_ARG;
  if (n_poses == 0) return HS_OK;
  if (!poses_map || (!points && n_points > 0)) return HS_ERR_INVALID_ARG;
  HS_CK(cudaSetDevice(e->device));
  auto align = [](size_t x) { return (x + 255) & ~(size_t)255; };
  const size_t b_poses = (size_t)n_poses * 3 * sizeof(float), b_pts = (size_t)(n_points > 0 ? n_points : 1) * 2 * sizeof(float);
  const size_t b_out = (size_t)n_poses * sizeof(float);
  const size_t o_pts = align(b_poses), o_r = o_pts + align(b_pts), o_l = o_r + align(b_out), total = o_l + align(b_out);
  if ((st = ensure_scratch(e, total)) != HS_OK) return st;
  char* base = (char*)e->d_scratch;
  HS_CK(cudaMemcpyAsync(base, poses_map, b_poses, cudaMemcpyHostToDevice, e->stream));
  if (n_points > 0) HS_CK(cudaMemcpyAsync(base + o_pts, points, (size_t)n_points * 2 * sizeof(float), cudaMemcpyHostToDevice, e->stream));
  e->h2d_bytes += b_poses + (size_t)n_points * 8;
  const size_t smem = (size_t)(n_points > 0 ? n_points : 1) * sizeof(float);
  if (smem > 200 * 1024) return HS_ERR_INVALID_ARG;
  if (smem > 48 * 1024) {
    HS_CK(cudaFuncSetAttribute(k_residual<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    HS_CK(cudaFuncSetAttribute(k_residual<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  }
  if (e->P.strict)
    k_residual<true><<<n_poses, HS_MATCH_THREADS, smem, e->stream>>>(e->P, stream, level, (const float*)base, n_poses, (const float2*)(base + o_pts),
                                                                    n_points, (float*)(base + o_r), (float*)(base + o_l));
  else
    k_residual<false><<<n_poses, HS_MATCH_THREADS, smem, e->stream>>>(e->P, stream, level, (const float*)base, n_poses, (const float2*)(base + o_pts),
                                                                     n_points, (float*)(base + o_r), (float*)(base + o_l));
  HS_LAUNCHED(e);
  if (residual_out) { HS_CK(cudaMemcpyAsync(residual_out, base + o_r, b_out, cudaMemcpyDeviceToHost, e->stream)); e->d2h_bytes += b_out; }
  if (likelihood_out) { HS_CK(cudaMemcpyAsync(likelihood_out, base + o_l, b_out, cudaMemcpyDeviceToHost, e->stream)); e->d2h_bytes += b_out; }
  HS_CK(cudaStreamSynchronize(e->stream));
  return HS_OK;
}

// OccGridMapUtil::getCovarianceForPose (HSL/map/OccGridMapUtil.h:106-160) + getCovMatrixWorldCoords (:162-189). The seven
// sigma-point likelihoods come from k_residual; the 3x3 arithmetic follows Eigen 3.3's evaluation order (fp32, unfused):
// likelihoods.sum() = (a0 + (a1 + a2)) + ((a3 + a4) + (a5 + a6)); mean += sigma_i * lh_i; cov += (lh_i * inv) * (d d^T).
int hs_covariance_for_pose(hs_engine* e, int stream, int level, const float* pose_map, const float* points, int n_points,
                           float* cov_map9, float* cov_world9) {
  if (!pose_map || !cov_map9) return HS_ERR_INVALID_ARG;
  const float dtx = 1.5f, dty = 1.5f, dang = 0.05f;
  const float x = pose_map[0], y = pose_map[1], ang = pose_map[2];
  volatile float sp[7][3] = {{x + dtx, y, ang}, {x - dtx, y, ang}, {x, y + dty, ang}, {x, y - dty, ang},
                             {x, y, ang + dang}, {x, y, ang - dang}, {x, y, ang}};
  float spf[21], lh[7];
  for (int i = 0; i < 7; ++i) for (int k = 0; k < 3; ++k) spf[3 * i + k] = sp[i][k];
  int st = hs_likelihood_batch(e, stream, level, spf, 7, points, n_points, nullptr, lh);
  if (st != HS_OK) return st;
  volatile float s12 = lh[1] + lh[2], s012 = lh[0] + s12, s34 = lh[3] + lh[4], s56 = lh[5] + lh[6], s3456 = s34 + s56;
  volatile float sum = s012 + s3456;
  volatile float inv = 1 / sum;
  volatile float mean[3] = {0.0f, 0.0f, 0.0f};
  for (int i = 0; i < 7; ++i)
    for (int k = 0; k < 3; ++k) { volatile float t = sp[i][k] * lh[i]; mean[k] = mean[k] + t; }
  for (int k = 0; k < 3; ++k) mean[k] = mean[k] * inv;
  volatile float cov[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  for (int i = 0; i < 7; ++i) {
    volatile float d[3] = {sp[i][0] - mean[0], sp[i][1] - mean[1], sp[i][2] - mean[2]};
    volatile float w = lh[i] * inv;
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 3; ++c) { volatile float o = d[r] * d[c]; volatile float t = w * o; cov[3 * r + c] = cov[3 * r + c] + t; }
  }
  for (int k = 0; k < 9; ++k) cov_map9[k] = cov[k];
  if (cov_world9) {
    float res = e->cfg.map_resolution;
    for (int i = 0; i < level; ++i) res *= 2.0f;          // getCellLength() of the level
    volatile float scale = res, scale_sq = scale * scale;
    for (int k = 0; k < 9; ++k) cov_world9[k] = 0.0f;
    cov_world9[0] = cov[0] * scale_sq;
    cov_world9[4] = cov[4] * scale_sq;
    cov_world9[3] = cov[3] * scale_sq; cov_world9[1] = cov_world9[3];
    cov_world9[6] = cov[6] * scale;    cov_world9[2] = cov_world9[6];
    cov_world9[7] = cov[7] * scale;    cov_world9[5] = cov_world9[7];
    cov_world9[8] = cov[8];
  }
  return HS_OK;
}

int hs_interp_map_value_batch(hs_engine* e, int stream, int level, const float* coords_map, int n, float* values_out, float* derivs_out) {
  int st = check_stream_level(e, stream, level);
  if (st != HS_OK) return st;
  if (n < 0) return HS_ERR_INVALID_ARG;
  if (n == 0) return HS_OK;
  if (!coords_map) return HS_ERR_INVALID_ARG;
  HS_CK(cudaSetDevice(e->device));
  auto align = [](size_t x) { return (x + 255) & ~(size_t)255; };
  const size_t b_in = (size_t)n * 2 * sizeof(float), b_out = (size_t)n * 4 * sizeof(float);
  const size_t o_out = align(b_in);
  if ((st = ensure_scratch(e, o_out + b_out)) != HS_OK) return st;
  char* base = (char*)e->d_scratch;
  HS_CK(cudaMemcpyAsync(base, coords_map, b_in, cudaMemcpyHostToDevice, e->stream));
  e->h2d_bytes += b_in;
  k_interp<<<(n + 127) / 128, 128, 0, e->stream>>>(e->P, stream, level, (const float2*)base, n, (float4*)(base + o_out));
  HS_LAUNCHED(e);
  std::vector<float> host((size_t)n * 4);
  HS_CK(cudaMemcpyAsync(host.data(), base + o_out, b_out, cudaMemcpyDeviceToHost, e->stream));
  e->d2h_bytes += b_out;
  HS_CK(cudaStreamSynchronize(e->stream));
  for (int i = 0; i < n; ++i) {
    if (values_out) values_out[i] = host[4 * (size_t)i + 3];
    if (derivs_out) { derivs_out[3 * (size_t)i] = host[4 * (size_t)i]; derivs_out[3 * (size_t)i + 1] = host[4 * (size_t)i + 1]; derivs_out[3 * (size_t)i + 2] = host[4 * (size_t)i + 2]; }
  }
  return HS_OK;
}

int hs_occupancy_ray_batch(hs_engine* e, int stream, int level, const int32_t* begin_cells, const int32_t* end_cells, int n,
                           float* dist_out, int32_t* hit_cells_out) {
  int st = check_stream_level(e, stream, level);
  if (st != HS_OK) return st;
  if (n < 0) return HS_ERR_INVALID_ARG;
  if (n == 0) return HS_OK;
  if (!begin_cells || !end_cells || !dist_out) return HS_ERR_INVALID_ARG;
  HS_CK(cudaSetDevice(e->device));
  auto align = [](size_t x) { return (x + 255) & ~(size_t)255; };
  const size_t b_c = (size_t)n * 2 * sizeof(int32_t), b_d = (size_t)n * sizeof(float);
  const size_t o_e = align(b_c), o_d = o_e + align(b_c), o_h = o_d + align(b_d);
  if ((st = ensure_scratch(e, o_h + b_c)) != HS_OK) return st;
  char* base = (char*)e->d_scratch;
  HS_CK(cudaMemcpyAsync(base, begin_cells, b_c, cudaMemcpyHostToDevice, e->stream));
  HS_CK(cudaMemcpyAsync(base + o_e, end_cells, b_c, cudaMemcpyHostToDevice, e->stream));
  HS_CK(cudaMemsetAsync(base + o_h, 0xFF, b_c, e->stream));   // hit cells default to (-1, -1)
  e->h2d_bytes += 2 * b_c;
  k_occupancy_ray<<<(n + 127) / 128, 128, 0, e->stream>>>(e->P, stream, level, (const int2*)base, (const int2*)(base + o_e), n,
                                                          (float*)(base + o_d), (int2*)(base + o_h));
  HS_LAUNCHED(e);
  HS_CK(cudaMemcpyAsync(dist_out, base + o_d, b_d, cudaMemcpyDeviceToHost, e->stream));
  if (hit_cells_out) HS_CK(cudaMemcpyAsync(hit_cells_out, base + o_h, b_c, cudaMemcpyDeviceToHost, e->stream));
  e->d2h_bytes += b_d + (hit_cells_out ? b_c : 0);
  HS_CK(cudaStreamSynchronize(e->stream));
  return HS_OK;
}

// Everything of one stream that is not a grid cell: with hs_download_level / hs_upload_level of every level this is a
// complete checkpoint of a HectorSlamProcessor (the reference has none; GridMapBase::operator= copies a grid, GridMapBase.h:182-205).
int hs_export_stream_state(hs_engine* e, int stream, hs_stream_state* out, float* coarse_points_out) {
  if (!e || !out || stream < 0 || stream >= e->P.n_streams) return HS_ERR_INVALID_ARG;
  HS_CK(cudaSetDevice(e->device));
  const HsParams& P = e->P;
  const size_t s = (size_t)stream;
  HS_CK(cudaMemcpyAsync(out->last_scan_match_pose, P.last_pose + 3 * s, 3 * sizeof(float), cudaMemcpyDeviceToHost, e->stream));
  HS_CK(cudaMemcpyAsync(out->last_map_update_pose, P.last_update_pose + 3 * s, 3 * sizeof(float), cudaMemcpyDeviceToHost, e->stream));
  HS_CK(cudaMemcpyAsync(out->last_scan_match_cov, P.cov + 9 * s, 9 * sizeof(float), cudaMemcpyDeviceToHost, e->stream));
  HS_CK(cudaMemcpyAsync(&out->curr_update_index, P.cur_update_index + s, sizeof(int), cudaMemcpyDeviceToHost, e->stream));
  HS_CK(cudaMemcpyAsync(&out->last_update_index, P.last_update_index + s, sizeof(int), cudaMemcpyDeviceToHost, e->stream));
  HS_CK(cudaMemcpyAsync(&out->coarse_count, P.coarse_cnt + s, sizeof(int), cudaMemcpyDeviceToHost, e->stream));
  HS_CK(cudaMemcpyAsync(out->coarse_origo, P.coarse_origo + 2 * s, 2 * sizeof(float), cudaMemcpyDeviceToHost, e->stream));
  if (coarse_points_out)
    HS_CK(cudaMemcpyAsync(coarse_points_out, P.coarse_pts + s * P.max_points * 2, (size_t)P.max_points * 2 * sizeof(float), cudaMemcpyDeviceToHost, e->stream));
  HS_CK(cudaStreamSynchronize(e->stream));
  return HS_OK;
}

int hs_import_stream_state(hs_engine* e, int stream, const hs_stream_state* in, const float* coarse_points) {
  if (!e || !in || stream < 0 || stream >= e->P.n_streams) return HS_ERR_INVALID_ARG;
  if (in->coarse_count < 0 || in->coarse_count > e->P.max_points || (in->coarse_count > 0 && !coarse_points)) return HS_ERR_INVALID_ARG;
  HS_CK(cudaSetDevice(e->device));
  const HsParams& P = e->P;
  const size_t s = (size_t)stream;
  HS_CK(cudaMemcpyAsync(P.last_pose + 3 * s, in->last_scan_match_pose, 3 * sizeof(float), cudaMemcpyHostToDevice, e->stream));
  HS_CK(cudaMemcpyAsync(P.last_update_pose + 3 * s, in->last_map_update_pose, 3 * sizeof(float), cudaMemcpyHostToDevice, e->stream));
  HS_CK(cudaMemcpyAsync(P.cov + 9 * s, in->last_scan_match_cov, 9 * sizeof(float), cudaMemcpyHostToDevice, e->stream));
  HS_CK(cudaMemcpyAsync(P.cur_update_index + s, &in->curr_update_index, sizeof(int), cudaMemcpyHostToDevice, e->stream));
  HS_CK(cudaMemcpyAsync(P.last_update_index + s, &in->last_update_index, sizeof(int), cudaMemcpyHostToDevice, e->stream));
  HS_CK(cudaMemcpyAsync(P.coarse_cnt + s, &in->coarse_count, sizeof(int), cudaMemcpyHostToDevice, e->stream));
  HS_CK(cudaMemcpyAsync(P.coarse_origo + 2 * s, in->coarse_origo, 2 * sizeof(float), cudaMemcpyHostToDevice, e->stream));
  if (coarse_points)
    HS_CK(cudaMemcpyAsync(P.coarse_pts + s * P.max_points * 2, coarse_points, (size_t)P.max_points * 2 * sizeof(float), cudaMemcpyHostToDevice, e->stream));
  HS_CK(cudaStreamSynchronize(e->stream));
  return HS_OK;
}

// ---- map access ----------------------------------------------------------------------------------

static int check_level(const hs_engine* e, int stream, int level) {
  if (!e || stream < 0 || stream >= e->P.n_streams || level < 0 || level >= e->P.n_levels) return HS_ERR_INVALID_ARG;
  return HS_OK;
}

int hs_download_level(hs_engine* e, int stream, int level, hs_cell* cells_out) {
  int st = check_level(e, stream, level);
  if (st != HS_OK) return st;
  if (!cells_out) return HS_ERR_INVALID_ARG;
  HS_CK(cudaSetDevice(e->device));
  const HsLevel& L = e->P.lv[level];
  size_t n = (size_t)L.sx * L.sy;
  if ((st = ensure_scratch(e, n * sizeof(hs_cell))) != HS_OK) return st;
  size_t base = (size_t)stream * (size_t)e->P.cells_per_stream + (size_t)L.cell_offset;
  k_pack_cells<<<148 * 4, 256, 0, e->stream>>>(e->P.val + base, e->P.idx + base, e->P.mk + base,
                                               e->P.mk_base + (size_t)stream * HS_MAX_LEVELS + level, (hs_cell*)e->d_scratch, n);
  HS_LAUNCHED(e);
  HS_CK(cudaMemcpyAsync(cells_out, e->d_scratch, n * sizeof(hs_cell), cudaMemcpyDeviceToHost, e->stream));
  e->d2h_bytes += n * sizeof(hs_cell);
  HS_CK(cudaStreamSynchronize(e->stream));
  return HS_OK;
}

int hs_upload_level(hs_engine* e, int stream, int level, const hs_cell* cells_in) {
  int st = check_level(e, stream, level);
  if (st != HS_OK) return st;
  if (!cells_in) return HS_ERR_INVALID_ARG;
  HS_CK(cudaSetDevice(e->device));
  const HsLevel& L = e->P.lv[level];
  size_t n = (size_t)L.sx * L.sy;
  if ((st = ensure_scratch(e, n * sizeof(hs_cell))) != HS_OK) return st;
  size_t base = (size_t)stream * (size_t)e->P.cells_per_stream + (size_t)L.cell_offset;
  e->map_epoch++;
  HS_CK(cudaMemcpyAsync(e->d_scratch, cells_in, n * sizeof(hs_cell), cudaMemcpyHostToDevice, e->stream));
  e->h2d_bytes += n * sizeof(hs_cell);
  k_unpack_cells<<<148 * 4, 256, 0, e->stream>>>((const hs_cell*)e->d_scratch, e->P.val + base, e->P.idx + base, e->P.mk + base, n);
  HS_LAUNCHED(e);
  HS_CK(cudaStreamSynchronize(e->stream));
  return HS_OK;
}

int hs_get_update_index(hs_engine* e, int stream, int level, int* update_index_out) {
  int st = check_level(e, stream, level);
  if (st != HS_OK) return st;
  if (!update_index_out) return HS_ERR_INVALID_ARG;
  HS_CK(cudaSetDevice(e->device));
  HS_CK(cudaMemcpyAsync(update_index_out, e->P.last_update_index + stream, sizeof(int), cudaMemcpyDeviceToHost, e->stream));
  HS_CK(cudaStreamSynchronize(e->stream));
  return HS_OK;
}

int hs_classify_level(hs_engine* e, int stream, int level, int8_t* occupancy_out) {
  int st = check_level(e, stream, level);
  if (st != HS_OK) return st;
  if (!occupancy_out) return HS_ERR_INVALID_ARG;
  HS_CK(cudaSetDevice(e->device));
  const HsLevel& L = e->P.lv[level];
  size_t n = (size_t)L.sx * L.sy;
  if ((st = ensure_scratch(e, n)) != HS_OK) return st;
  size_t base = (size_t)stream * (size_t)e->P.cells_per_stream + (size_t)L.cell_offset;
  k_classify<<<148 * 4, 256, 0, e->stream>>>(e->P.val + base, (int8_t*)e->d_scratch, n);
  HS_LAUNCHED(e);
  HS_CK(cudaMemcpyAsync(occupancy_out, e->d_scratch, n, cudaMemcpyDeviceToHost, e->stream));
  e->d2h_bytes += n;
  HS_CK(cudaStreamSynchronize(e->stream));
  return HS_OK;
}

// HectorMappingRos::setServiceGetMapData (HectorMappingRos.cpp:544-560): origin = getWorldCoords(0,0) - cellLength/2.
int hs_get_map_info(const hs_engine* e, int level, hs_map_info* out) {
  if (!e || !out || level < 0 || level >= e->P.n_levels) return HS_ERR_INVALID_ARG;
  const HsLevel& L = e->P.lv[level];
  float res = e->cfg.map_resolution;
  for (int i = 0; i < level; ++i) res *= 2.0f;
  volatile float zx = L.a * 0.0f, zy = L.a * 0.0f;
  volatile float ox = zx + L.twx, oy = zy + L.twy;     // worldTmap * (0, 0)
  volatile float half = res * 0.5f;
  out->origin_x = ox - half;
  out->origin_y = oy - half;
  out->resolution = res;
  out->width = L.sx;
  out->height = L.sy;
  return HS_OK;
}

// ---- multi-GPU plumbing for single-process callers (include/hsgpu/ShardedEngine.h) ---------------------------------

int hs_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}

uint64_t hs_get_map_epoch(const hs_engine* e) { return e ? e->map_epoch : 0; }

// Replicates all levels of one stream's maps (both planes) from one engine into a stream of another engine with the same map
// geometry -- on the same GPU or on a peer (cudaMemcpyPeerAsync over NVLink) -- ordered after everything queued on the source
// engine and before everything queued later on the destination engine. Returns when the copy has completed.
int hs_copy_stream_maps(hs_engine* dst, int dst_stream, hs_engine* src, int src_stream) {
  if (!dst || !src || dst_stream < 0 || dst_stream >= dst->P.n_streams || src_stream < 0 || src_stream >= src->P.n_streams) return HS_ERR_INVALID_ARG;
  if (dst->P.n_levels != src->P.n_levels || dst->P.cells_per_stream != src->P.cells_per_stream || dst->P.lv[0].sx != src->P.lv[0].sx ||
      dst->P.lv[0].sy != src->P.lv[0].sy)
    return HS_ERR_INVALID_ARG;
  if (dst == src && dst_stream == src_stream) return HS_OK;
  const size_t cells = (size_t)src->P.cells_per_stream;
  HS_CK(cudaSetDevice(src->device));
  k_fold_markers<<<src->P.n_levels, 512, 0, src->stream>>>(src->P, src_stream);   // the int32 marker plane is complete afterwards
  HS_LAUNCHED(src);
  HS_CK(cudaStreamSynchronize(src->stream));       // the source maps are final
  HS_CK(cudaSetDevice(dst->device));
  dst->map_epoch++;
  HS_CK(cudaMemcpyPeerAsync(dst->P.val + (size_t)dst_stream * cells, dst->device, src->P.val + (size_t)src_stream * cells, src->device,
                            cells * sizeof(float), dst->stream));
  HS_CK(cudaMemcpyPeerAsync(dst->P.idx + (size_t)dst_stream * cells, dst->device, src->P.idx + (size_t)src_stream * cells, src->device,
                            cells * sizeof(int), dst->stream));
  HS_CK(cudaMemsetAsync(dst->P.mk + (size_t)dst_stream * cells, 0, cells, dst->stream));
  HS_CK(cudaStreamSynchronize(dst->stream));
  return HS_OK;
}

int hs_get_stats(hs_engine* e, hs_stats* out) {
  if (!e || !out) return HS_ERR_INVALID_ARG;
  HS_CK(cudaSetDevice(e->device));
  unsigned long long c[4];
  HS_CK(cudaMemcpyAsync(c, e->P.counters, sizeof(c), cudaMemcpyDeviceToHost, e->stream));
  HS_CK(cudaStreamSynchronize(e->stream));
  out->updates = e->updates;
  out->integrations = c[1];
  out->angle_clamps = c[2];
  out->cell_visits = c[3];
  out->kernel_launches = e->launches;
  out->h2d_bytes = e->h2d_bytes;
  out->d2h_bytes = e->d2h_bytes;
  return HS_OK;
}

int hs_profile_enable(hs_engine* e, int on) {
  if (!e) return HS_ERR_INVALID_ARG;
  HS_CK(cudaSetDevice(e->device));
  HS_CK(cudaStreamSynchronize(e->stream));
  e->profiling = on ? 1 : 0;
  e->prof_every = on > 1 ? on : 1;
  e->prof_tick = 0;
  e->prof_used = 0;
  e->prof_ids->clear();
  return HS_OK;
}

int hs_profile_collect(hs_engine* e, hs_kernel_times* out) {
  if (!e || !out) return HS_ERR_INVALID_ARG;
  HS_CK(cudaSetDevice(e->device));
  HS_CK(cudaStreamSynchronize(e->stream));
  memset(out, 0, sizeof(*out));
  for (size_t i = 0; i < e->prof_ids->size(); ++i) {
    float ms = 0.0f;
    HS_CK(cudaEventElapsedTime(&ms, (*e->prof_events)[2 * i], (*e->prof_events)[2 * i + 1]));
    int id = (*e->prof_ids)[i];
    if (id >= 0 && id < HS_K_COUNT) {
      out->ms[id] += (double)ms;
      out->launches[id] += 1;
    }
  }
  e->prof_used = 0;
  e->prof_ids->clear();
  return HS_OK;
}

}  // extern "C"
