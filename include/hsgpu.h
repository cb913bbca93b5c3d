/* hsgpu -- B200-native batched scan-to-map engine behind the hector_slam_lib API.
 *
 * Thin C ABI (plain pointers and sizes, no C++/torch types) over the sm_100a CUDA engine in
 * tfe-slam-ucl_b200/csrc/. One engine lives on one GPU and holds `n_streams` independent
 * hectorslam::HectorSlamProcessor states (multi-resolution log-odds grids, poses, update
 * counters) resident in HBM. Every entry point cites the reference interface it replaces;
 * HSL/ = src/hector_slam/hector_mapping/include/hector_slam_lib/ of the reference tree.
 *
 * Conventions
 *   - every function returns an int status: HS_OK (0), a negative HS_ERR_* code, or a positive
 *     cudaError_t value; hs_status_string() explains any of them. There is NO CPU fallback:
 *     without a usable CUDA device hs_create fails with HS_ERR_NO_DEVICE / the CUDA error.
 *   - "host" pointers are caller-owned host memory (pinned memory makes the copies asynchronous);
 *     "device" pointers (the *_device entry points) are caller-owned device memory on the engine's GPU.
 *   - calls on one engine must be serialised by the caller (the reference is not re-entrant either:
 *     HSL/matcher/ScanMatcher.h keeps H/dTr as members). Different engines are independent.
 *   - poses are (x [m], y [m], theta [rad]) in world coordinates; points are in level-0 map cells
 *     relative to the sensor (what HectorMappingRos::rosLaserScanToDataContainer produces).
 *   - 3x3 matrices are row-major.
 */
#ifndef HSGPU_H
#define HSGPU_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HS_OK 0
#define HS_ERR_INVALID_ARG (-1)
#define HS_ERR_NO_DEVICE (-2)
#define HS_ERR_OUT_OF_MEMORY (-3)
#define HS_ERR_NOT_CONFIGURED (-4)

#define HS_MAX_LEVELS 8
#define HS_ALL_STREAMS (-1)

typedef struct hs_engine hs_engine;

/* Same 8-byte layout as LogOddsCell (HSL/map/GridMapLogOdds.h:37-103). */
typedef struct { float logOddsVal; int32_t updateIndex; } hs_cell;

/* Constructor arguments of HectorSlamProcessor (HSL/slam_main/HectorSlamProcessor.h:54-64) plus
 * the batch geometry. Library defaults are applied as in the reference: update factors 0.4 / 0.6
 * (HSL/map/GridMapLogOdds.h:117-118), map-update gate 0.4 m / 0.13 rad (HectorSlamProcessor.h:62-63). */
typedef struct {
  float map_resolution; /* mapResolution [m per level-0 cell] */
  int32_t map_size_x;   /* mapSizeX [cells, level 0] */
  int32_t map_size_y;   /* mapSizeY */
  float start_x;        /* startCoords.x (fraction of the map, 0.5 = centre) */
  float start_y;        /* startCoords.y */
  int32_t levels;       /* multi_res_size, 1..HS_MAX_LEVELS */
  int32_t n_streams;    /* independent processors held by this engine */
  int32_t max_points;   /* capacity of one DataContainer (beams per scan); bounded by the ray-caster's shared memory
                           (48 KB map + 8 B x hash size + 32 B per point <= the device's opt-in limit: 3679 on B200),
                           larger values are refused with HS_ERR_INVALID_ARG */
  int32_t device;       /* CUDA device ordinal */
} hs_config;

/* ---- lifetime / configuration ---------------------------------------------------------- */

/* HectorSlamProcessor ctor + MapRepMultiMap ctor (HSL/slam_main/MapRepMultiMap.h:48-72).
 * Process-wide side effect: if the device's cudaLimitPersistingL2CacheSize is 0, it is raised to 16 MB (the matcher's
 * evict-last gathers live there; HSGPU_L2_PERSIST_MB sets another size, a limit the process has already chosen is kept). */
int hs_create(const hs_config* cfg, hs_engine** out);
/* ~HectorSlamProcessor (HectorSlamProcessor.h:66-69). */
int hs_destroy(hs_engine* e);
/* HectorSlamProcessor::reset (HectorSlamProcessor.h:115-124); stream = HS_ALL_STREAMS or an index.
 * As in the reference the per-grid update counters keep counting across resets. */
int hs_reset(hs_engine* e, int stream);
/* setUpdateFactorFree / setUpdateFactorOccupied (HectorSlamProcessor.h:136-137), all streams. */
int hs_set_update_factor_free(hs_engine* e, float free_factor);
int hs_set_update_factor_occupied(hs_engine* e, float occupied_factor);
/* setMapUpdateMinDistDiff / setMapUpdateMinAngleDiff (HectorSlamProcessor.h:138-139), all streams. */
int hs_set_map_update_min_dist_diff(hs_engine* e, float min_dist);
int hs_set_map_update_min_angle_diff(hs_engine* e, float min_angle);
/* Debug/parity switch: accumulate the Hessian and gradient sums in the reference's sequential point order
 * (OccGridMapUtil.h:76-98) instead of per-lane partial sums + warp butterfly. Bitwise CPU-identical, slower. */
int hs_set_strict_summation(hs_engine* e, int on);
/* getScaleToMap / getMapLevels (HectorSlamProcessor.h:128-130). */
float hs_get_scale_to_map(const hs_engine* e);
int hs_get_map_levels(const hs_engine* e);
int hs_get_n_streams(const hs_engine* e);
int hs_get_max_points(const hs_engine* e);
/* GridMapBase::getSizeX/getSizeY/getCellLength of one level (HSL/map/GridMapBase.h:60-61,303-306). */
int hs_get_level_dims(const hs_engine* e, int level, int* size_x, int* size_y, float* cell_length);
/* Launch all work of this engine on an externally owned cudaStream_t (NULL = engine-owned stream). */
int hs_set_cuda_stream(hs_engine* e, void* cuda_stream);
/* Block until all work queued by this engine has finished. */
int hs_synchronize(hs_engine* e);
const char* hs_status_string(int status);

/* ---- the hot path ------------------------------------------------------------------------ */

/* HectorSlamProcessor::update (HectorSlamProcessor.h:71-113) for n streams in one launch sequence:
 * MapRepMultiMap::matchData -> poseDifferenceLargerThan gate -> MapRepMultiMap::updateByScan.
 *   stream_ids           [n] engine stream index of each slot, or NULL for 0..n-1. Ids must be in range and DISTINCT (two
 *                        slots on one stream would race on its maps): every host-pointer entry point checks this and
 *                        returns HS_ERR_INVALID_ARG; the *_device variants cannot look at device memory without a
 *                        synchronisation and leave the requirement to the caller
 *   points               [n][max_points][2] DataContainer points (level-0 cells), only counts[i] used
 *   counts               [n] DataContainer::getSize()
 *   origos               [n][2] DataContainer::getOrigo(), or NULL for (0,0)
 *   pose_hints           [n][3] poseHintWorld, or NULL to use each stream's last scan-match pose
 *   map_without_matching [n] bool flags, or NULL for false
 *   poses_out            [n][3] getLastScanMatchPose() after the update, or NULL
 *   cov_out              [n][9] getLastScanMatchCovariance() (the last level-0 Hessian), or NULL
 * Host variant: copies inputs H2D, runs, copies outputs D2H and returns when they are valid.
 * Deferred integration: when points, counts and poses_out are pinned host memory (cudaHostAlloc / cudaHostRegister),
 * stream_ids and cov_out are NULL, the call returns as soon as the poses are in poses_out; this batch's updateByScan keeps
 * running on the engine's stream, every later call on the engine (another update, a map or pose read, hs_synchronize) is
 * ordered behind it, and the input buffers may be reused on return. The next call's points are then fetched while that
 * integration runs. HSGPU_E2E_DEFER=0 restores "returns after the integration". Consequence for device-resident consumers:
 * work the caller launches on ANOTHER CUDA stream (a kernel reading hs_device_pose_ptr, a peer copy of the grids) is not
 * ordered behind the deferred integration -- call hs_synchronize first, or run it on the stream given to hs_set_cuda_stream. */
int hs_update_batch(hs_engine* e, int n, const int32_t* stream_ids, const float* points, const int32_t* counts,
                    const float* origos, const float* pose_hints, const uint8_t* map_without_matching,
                    float* poses_out, float* cov_out);
/* Device variant: all pointers are device memory; asynchronous on the engine's stream; results stay
 * in the engine (hs_get_poses / hs_get_covariances, or hs_device_pose_ptr). */
int hs_update_batch_device(hs_engine* e, int n, const int32_t* stream_ids, const float* points, const int32_t* counts,
                           const float* origos, const float* pose_hints, const uint8_t* map_without_matching);

/* Pipelined submission -- update() for n streams WITHOUT waiting for the result (no reference counterpart: the reference's
 * update() is synchronous; this is what a server feeding queued scans uses). The inputs are copied by DMA on a copy stream
 * into one of two staging sets, so that batch k+1's transfer overlaps batch k's matching and integration; poses_out (or
 * NULL) is written by the matcher. ALL host buffers must be pinned and must not be touched until hs_wait(); with pageable
 * memory the call degrades to the synchronous hs_update_batch. Pose hints default to each stream's last pose ON THE
 * DEVICE, so consecutive scans of the same streams can be submitted back to back. Results are identical to hs_update_batch. */
int hs_submit_batch(hs_engine* e, int n, const int32_t* stream_ids, const float* points, const int32_t* counts,
                    const float* origos, const float* pose_hints, const uint8_t* map_without_matching, float* poses_out);
/* Returns when everything submitted so far has completed and its outputs are in place. */
int hs_wait(hs_engine* e);

/* LaserScan ingestion (HectorMappingRos::rosLaserScanToDataContainer, src/hector_slam/hector_mapping/src/
 * HectorMappingRos.cpp:483-507) fused in front of update: ranges [n][n_beams] -> points on device.
 * hs_set_scan_geometry fixes sensor_msgs/LaserScan's angle_min, angle_increment, range_min, range_max. */
int hs_set_scan_geometry(hs_engine* e, int n_beams, float angle_min, float angle_increment, float range_min, float range_max);
int hs_update_batch_ranges(hs_engine* e, int n, const int32_t* stream_ids, const float* ranges,
                           const float* pose_hints, const uint8_t* map_without_matching, float* poses_out, float* cov_out);
int hs_update_batch_ranges_device(hs_engine* e, int n, const int32_t* stream_ids, const float* ranges,
                                  const float* pose_hints, const uint8_t* map_without_matching);
/* Host helper with the same arithmetic (for callers that want DataContainer points). Returns the count. */
int hs_scan_to_points_host(const float* ranges, int n_beams, float angle_min, float angle_increment,
                           float range_min, float range_max, float scale_to_map, float* points_out);
/* The same for n_scans scans with `threads` host threads: points_out [n_scans][max_points][2] (zero padded),
 * counts_out [n_scans]. */
int hs_scan_to_points_host_batch(const float* ranges, int n_scans, int n_beams, float angle_min, float angle_increment,
                                 float range_min, float range_max, float scale_to_map, int max_points,
                                 float* points_out, int32_t* counts_out, int threads);

/* The node's default ingestion path (use_tf_scan_transformation, HectorMappingRos.cpp:273-282):
 * rosPointCloudToDataContainer (HectorMappingRos.cpp:509-542) for n clouds on the GPU, then update(). clouds_xyz
 * [n][max_cloud][3] laser-frame points, cloud_counts [n]; transforms12 [n][12] = the laser->base tf::Transform as
 * row-major 3x3 basis followed by the origin, in double like tf; filters = the node's parameters
 * (laser_min_dist^2, laser_max_dist^2, laser_z_min_value, laser_z_max_value, :98-108). origo = laser position * scaleToMap.
 * Points beyond max_points are dropped. Host pointers. */
int hs_update_batch_clouds(hs_engine* e, int n, const int32_t* stream_ids, const float* clouds_xyz, const int32_t* cloud_counts,
                           int max_cloud, const double* transforms12, float sqr_min_dist, float sqr_max_dist, float z_min, float z_max,
                           const float* pose_hints, const uint8_t* map_without_matching, float* poses_out, float* cov_out);
/* Host helper with the same arithmetic for one cloud (callers that want DataContainer points + origo). Returns the count. */
int hs_cloud_to_points_host(const float* cloud_xyz, int n_cloud, const double* transform12, float scale_to_map,
                            float sqr_min_dist, float sqr_max_dist, float z_min, float z_max, int max_points,
                            float* points_out, float* origo_out);

/* getLastScanMatchPose / getLastScanMatchCovariance (HectorSlamProcessor.h:126-127) of streams
 * [first, first+n) into host memory. hs_get_last_map_update_poses exposes lastMapUpdatePose. */
int hs_get_poses(hs_engine* e, int first, int n, float* poses_out);
int hs_get_covariances(hs_engine* e, int first, int n, float* cov_out);
int hs_get_last_map_update_poses(hs_engine* e, int first, int n, float* poses_out);
/* Device-resident [n_streams][3] array of last scan-match poses (valid for the engine's lifetime). */
const float* hs_device_pose_ptr(const hs_engine* e);

/* ---- single-function entry points (parity tests, relocalisation, teacher forcing) ------- */

/* OccGridMapUtil::getCompleteHessianDerivs (HSL/map/OccGridMapUtil.h:64-104) of ONE stream's level for
 * n_poses pose hypotheses (MAP coordinates of that level) over one point set already scaled to that
 * level: H9 [n_poses][9], dTr3 [n_poses][3]. All host pointers. */
int hs_hessian_derivs_batch(hs_engine* e, int stream, int level, const float* poses_map, int n_poses,
                            const float* points, int n_points, float* H9, float* dTr3);
/* The same with every pointer in DEVICE memory on the engine's GPU: asynchronous on the engine's stream, no copies and no
 * synchronisation (a relocaliser scoring hypotheses in a loop keeps poses and results on the device). The level's
 * log-odds are converted to probabilities once (a plane of 2x2 probability blocks, the GPU form of the reference's
 * GridMapCacheArray) and reused by later calls until the stream's maps change. */
int hs_hessian_derivs_batch_device(hs_engine* e, int stream, int level, const float* poses_map, int n_poses,
                                   const float* points, int n_points, float* H9, float* dTr3);
/* Relocalisation-style multi-start matching (BASELINE configs[3]): MapRepMultiMap::matchData (MapRepMultiMap.h:116-132) of ONE
 * scan (points [n_points][2], level-0 cells) against ONE stream's maps from n_hyp world pose hints [n_hyp][3] at once.
 * The stream's state is not touched (no pose, no coarse containers, no integration). poses_out [n_hyp][3], cov_out
 * [n_hyp][9] (the last level-0 Hessian, what getLastScanMatchCovariance would return); either may be NULL. Host pointers. */
int hs_match_hypotheses(hs_engine* e, int stream, const float* points, int n_points, const float* pose_hints, int n_hyp,
                        float* poses_out, float* cov_out);
/* MapRepMultiMap::matchData only (HSL/slam_main/MapRepMultiMap.h:116-132): no gate, no integration,
 * last scan-match pose/covariance are updated like update() would. Host pointers. */
int hs_match_batch(hs_engine* e, int n, const int32_t* stream_ids, const float* points, const int32_t* counts,
                   const float* origos, const float* pose_hints, float* poses_out, float* cov_out);
/* MapRepMultiMap::updateByScan + onMapUpdated (MapRepMultiMap.h:107-114,134-147) with caller-given
 * (teacher-forced) world poses: coarse-level containers are first refreshed from `points` exactly as
 * matchData's setFrom would (MapRepMultiMap.h:127). Does not touch lastMapUpdatePose. Host pointers. */
int hs_raycast_batch(hs_engine* e, int n, const int32_t* stream_ids, const float* points, const int32_t* counts,
                     const float* origos, const float* poses_world);
/* MapRepMultiMap::updateByScan + onMapUpdated (MapRepMultiMap.h:107-114,134-147) exactly: level 0 integrates `points`,
 * level i > 0 integrates dataContainers[i-1] AS THE LAST matchData LEFT IT (possibly a different scan, or empty when no
 * match has run yet) -- what a caller sees that drives the reference's MapRepresentationInterface::updateByScan itself,
 * e.g. HectorSlamProcessor::update(..., map_without_matching = true). Same arguments as hs_raycast_batch. Host pointers. */
int hs_integrate_batch(hs_engine* e, int n, const int32_t* stream_ids, const float* points, const int32_t* counts,
                       const float* origos, const float* poses_world);

/* nav_msgs/OccupancyGrid metadata of a level as HectorMappingRos::setServiceGetMapData fills it
 * (HectorMappingRos.cpp:544-560): origin = getWorldCoords(0,0) - cellLength/2, resolution = getCellLength(), width/height
 * = grid size. With hs_classify_level this is the published map. */
typedef struct {
  float origin_x, origin_y;   /* info.origin.position [m] (orientation is identity) */
  float resolution;           /* info.resolution [m/cell] */
  int32_t width, height;      /* info.width, info.height [cells] */
} hs_map_info;
int hs_get_map_info(const hs_engine* e, int level, hs_map_info* out);

/* Off-path members of OccGridMapUtil (HSL/map/OccGridMapUtil.h:106-285; not called by update(), used for relocalisation
 * scoring). Poses are MAP coordinates of the level, points already scaled to it; all host pointers.
 * getResidualForState (:216-233) and getLikelihoodForState (:201-214) for n_poses hypotheses; either output may be NULL. */
int hs_likelihood_batch(hs_engine* e, int stream, int level, const float* poses_map, int n_poses, const float* points,
                        int n_points, float* residual_out, float* likelihood_out);
/* getCovarianceForPose (:106-160): 7 sigma points (+-1.5 cells, +-0.05 rad) weighted by their likelihoods -> cov_map9;
 * cov_world9 (may be NULL) = getCovMatrixWorldCoords (:162-189) of it. */
int hs_covariance_for_pose(hs_engine* e, int stream, int level, const float* pose_map, const float* points, int n_points,
                           float* cov_map9, float* cov_world9);
/* interpMapValue (:241-285) -> values_out [n] and/or interpMapValueWithDerivatives (:287-347) -> derivs_out [n][3] at
 * n map coordinates [n][2] of the level. */
int hs_interp_map_value_batch(hs_engine* e, int stream, int level, const float* coords_map, int n, float* values_out, float* derivs_out);
/* HectorMapTools::DistanceMeasurementProvider::checkOccupancyBresenhami
 * (src/hector_slam/hector_map_tools/include/hector_map_tools/HectorMapTools.h:148-232) for n rays on the occupancy
 * classification of a level (cell == 100 <=> log-odds > 0): begin/end cells [n][2] (x, y); dist_out [n] = distance to the
 * first occupied cell in cells (truncated, as the reference) or -1; hit_cells_out [n][2] (may be NULL) = that cell or (-1,-1).
 * getDist's world<->cell conversion (HectorMapTools.h:131-146) is the caller's (origin/resolution of the published map). */
int hs_occupancy_ray_batch(hs_engine* e, int stream, int level, const int32_t* begin_cells, const int32_t* end_cells, int n,
                           float* dist_out, int32_t* hit_cells_out);

/* ---- map access --------------------------------------------------------------------------- */

/* getGridMap(level) contents (HectorSlamProcessor.h:131): sx*sy cells, row-major cell[y*sx + x]. */
int hs_download_level(hs_engine* e, int stream, int level, hs_cell* cells_out);
int hs_upload_level(hs_engine* e, int stream, int level, const hs_cell* cells_in);
/* GridMapBase::getUpdateIndex (HSL/map/GridMapBase.h:329). */
int hs_get_update_index(hs_engine* e, int stream, int level, int* update_index_out);
/* The occupancy classification HectorMappingRos::publishMap performs per cell (HectorMappingRos.cpp:447-470):
 * -1 unknown, 0 free, 100 occupied; sx*sy int8 values. */
int hs_classify_level(hs_engine* e, int stream, int level, int8_t* occupancy_out);

/* Everything of one stream that is not a grid cell (HectorSlamProcessor.h:141-153 members, the update counters of
 * OccGridMapBase.h:265 / GridMapBase.h:399, and the coarse-level containers of MapRepMultiMap.h:151). Together with
 * hs_download_level / hs_upload_level of every level this checkpoints and restores a stream bit for bit. */
typedef struct {
  float last_scan_match_pose[3];
  float last_map_update_pose[3];
  float last_scan_match_cov[9];
  int32_t curr_update_index;
  int32_t last_update_index;
  int32_t coarse_count;
  float coarse_origo[2];
} hs_stream_state;
/* coarse_points: [max_points][2] level-0 points behind MapRepMultiMap::dataContainers (may be NULL on export). */
int hs_export_stream_state(hs_engine* e, int stream, hs_stream_state* out, float* coarse_points_out);
int hs_import_stream_state(hs_engine* e, int stream, const hs_stream_state* in, const float* coarse_points);

/* ---- multi-GPU plumbing for single-process callers (include/hsgpu/ShardedEngine.h) ------ */

/* Number of usable CUDA devices (0 without a GPU: there is no CPU fallback). */
int hs_device_count(void);
/* Counter that changes whenever cells of any stream of the engine may have been written (update, integrate, upload, reset,
 * hs_copy_stream_maps): lets a caller cache replicas of a stream's maps. */
uint64_t hs_get_map_epoch(const hs_engine* e);
/* Replicates all levels of src's stream src_stream into dst's stream dst_stream (same map geometry; engines on the same GPU
 * or on peers: cudaMemcpyPeerAsync). What SURVEY 8(e) row 2 needs to shard the pose hypotheses of ONE stream over several
 * GPUs: the owner's maps are copied into a guest stream of every other engine. Per-stream state other than cells (poses,
 * update counters, coarse containers) is not copied. Returns when the copy has completed. */
int hs_copy_stream_maps(hs_engine* dst, int dst_stream, hs_engine* src, int src_stream);

/* ---- bookkeeping -------------------------------------------------------------------------- */

typedef struct {
  uint64_t updates;          /* stream-updates processed (scans matched or passed through) */
  uint64_t integrations;     /* stream-updates that passed the gate and were ray-cast into the maps */
  uint64_t angle_clamps;     /* Gauss-Newton steps whose rotation was clamped to +-0.2 rad
                                (the reference prints "SearchDir angle change too large", ScanMatcher.h:211,214) */
  uint64_t kernel_launches;  /* CUDA kernels launched by this engine so far */
  uint64_t h2d_bytes, d2h_bytes;
  uint64_t cell_visits;      /* sum over integrated levels and valid beams of max(|dx|,|dy|) + 1 grid cells
                                (the ray-caster's algorithmic work) */
} hs_stats;
int hs_get_stats(hs_engine* e, hs_stats* out);

/* Per-kernel device time, measured with CUDA events recorded on the launching stream around every hot-path
 * kernel while enabled. Index: 0 k_scan_to_points, 1 k_match / k_match2, 2 k_raycast, 3 k_hessian_derivs, 4 k_build_pblk
 * (the probability block plane behind the hypothesis entry points); the rest is reserved. */
#define HS_PROFILE_KERNELS 8
typedef struct {
  double ms[HS_PROFILE_KERNELS];
  uint64_t launches[HS_PROFILE_KERNELS];
} hs_kernel_times;
/* on: 0 = off, 1 = around every kernel, n > 1 = around the kernels of every n-th update only (the event pairs keep
 * consecutive kernels from overlapping their launch and drain, ~6 us per pair, so a sampled measurement perturbs less). */
int hs_profile_enable(hs_engine* e, int on);
/* Waits for the stream, sums the elapsed time per kernel since the last collect/enable, and clears the log. */
int hs_profile_collect(hs_engine* e, hs_kernel_times* out);

/* The deterministic synthetic scan streams the tests and bench.py feed the engine are NOT part of this library:
 * include/hsscangen.h, workload/libhsscangen.so. */

#ifdef __cplusplus
}
#endif
#endif /* HSGPU_H */
