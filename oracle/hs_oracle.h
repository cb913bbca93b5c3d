/* TEST INFRASTRUCTURE ONLY -- CPU oracle for the hector_mapping scan-to-map path.
 *
 * Plain-C restatement of hector_slam_lib (reference: /root/reference/src/hector_slam/
 * hector_mapping/include/hector_slam_lib/, abbreviated HSL/ below). Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load
 * this library; the product (libhsgpu.so) never links or calls it.
 *
 * Parity pin: this restatement is checked bit-for-bit (poses, Hessians, every cell's
 * {logOddsVal, updateIndex}) against oracle/_ref/libhsref.so -- the unmodified reference
 * headers compiled with a mini-Eigen shim -- by tests/test_oracle_vs_ref.py and against
 * the frozen fixtures in tests/golden/. The reference itself ships no tests or golden
 * vectors for this path, and Eigen (unpinned, absent) is restated
 * from its published 3.3 semantics: see DESIGN.md "Oracle and parity pin".
 */
#ifndef HS_ORACLE_H
#define HS_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

typedef struct hso_proc hso_proc;
typedef struct { float logOddsVal; int updateIndex; } hso_cell; /* HSL/map/GridMapLogOdds.h:99-100 */

hso_proc* hso_create(float res, int sx, int sy, float startx, float starty, int levels);
void hso_destroy(hso_proc* p);
void hso_reset(hso_proc* p);
void hso_set_update_factor_free(hso_proc* p, float f);
void hso_set_update_factor_occupied(hso_proc* p, float f);
void hso_set_map_update_min_dist_diff(hso_proc* p, float d);
void hso_set_map_update_min_angle_diff(hso_proc* p, float a);
float hso_get_scale_to_map(hso_proc* p);
int hso_get_map_levels(hso_proc* p);

void hso_update(hso_proc* p, const float* pts, int n, float ox, float oy, const float* hint, int map_without_matching);
void hso_get_pose(hso_proc* p, float* out3);
void hso_get_last_map_update_pose(hso_proc* p, float* out3);
void hso_get_cov(hso_proc* p, float* out9);
void hso_level_dims(hso_proc* p, int level, int* sx, int* sy, float* cell_length);
int hso_get_update_index(hso_proc* p, int level);
void hso_download_level(hso_proc* p, int level, hso_cell* out);
void hso_upload_level(hso_proc* p, int level, const hso_cell* in);
void hso_set_cell(hso_proc* p, int level, int index, float logodds, int update_index);
void hso_world_to_map(hso_proc* p, int level, const float* w, float* m);
void hso_map_to_world(hso_proc* p, int level, const float* m, float* w);

void hso_hessian_derivs(hso_proc* p, int level, const float* pose_map, const float* pts, int n, float* H9, float* dTr3);
void hso_hessian_derivs_batch(hso_proc* p, int level, const float* poses_map, int n_poses, const float* pts, int n, float* H9, float* dTr3);
void hso_interp_with_derivs(hso_proc* p, int level, float x, float y, float* out3);
void hso_match_level(hso_proc* p, int level, const float* hint_world, const float* pts, int n, int max_iter, float* pose_out, float* cov9);
void hso_update_by_scan_level(hso_proc* p, int level, const float* pts, int n, float ox, float oy, const float* pose_world);
void hso_refresh_coarse(hso_proc* p, const float* pts, int n, float ox, float oy);
void hso_update_by_scan_all(hso_proc* p, const float* pts, int n, float ox, float oy, const float* pose_world);
/* Off-path OccGridMapUtil members (HSL/map/OccGridMapUtil.h:106-285); poses in MAP coordinates of the level,
 * pts already scaled to the level. */
float hso_interp_map_value(hso_proc* p, int level, float x, float y);
float hso_residual_for_state(hso_proc* p, int level, const float* pose_map, const float* pts, int n);
float hso_likelihood_for_state(hso_proc* p, int level, const float* pose_map, const float* pts, int n);
void hso_covariance_for_pose(hso_proc* p, int level, const float* pose_map, const float* pts, int n, float* cov_map9, float* cov_world9);
/* HectorMapTools::DistanceMeasurementProvider::checkOccupancyBresenhami (hector_map_tools/HectorMapTools.h:148-232) */
float hso_check_occupancy_bresenhami(const signed char* grid, int size_x, int size_y, int x0, int y0, int x1, int y1, int* hit_xy);
int hso_pose_difference_larger_than(const float* p1, const float* p2, float d, float a);
float hso_normalize_angle(float a);

/* Bookkeeping for the roofline's algorithmic integrate bytes:
 * sum over integrated levels/beams of (max(|dx|,|dy|) + 1) cell visits, and integrated scans. */
unsigned long long hso_get_cell_visits(hso_proc* p);
unsigned long long hso_get_integrations(hso_proc* p);

/* rosLaserScanToDataContainer (src/hector_slam/hector_mapping/src/HectorMappingRos.cpp:483-507).
 * Returns the number of points written to pts_out (x,y interleaved, level-0 cell units). */
int hso_scan_to_points(const float* ranges, int n_beams, float angle_min, float angle_increment,
                       float range_min, float range_max, float scale_to_map, float* pts_out);

/* rosPointCloudToDataContainer (HectorMappingRos.cpp:509-542); transform12 = row-major 3x3 basis + origin of the
 * laser->base tf::Transform (double). Returns the number of points written; origo_out[2] = laser position * scaleToMap. */
int hso_cloud_to_points(const float* cloud_xyz, int n_cloud, const double* transform12, float scale_to_map,
                        float sqr_min_dist, float sqr_max_dist, float z_min, float z_max, float* pts_out, float* origo_out);

/* Multi-stream CPU runner, same contract as hsref_run_streams in ref_driver.cpp. */
double hso_run_streams(hso_proc** procs, int n_streams, int n_scans, int max_pts, const float* pts, const int* counts,
                       int threads, int stream_major, float* poses_out, float* cov_out);

#ifdef __cplusplus
}
#endif
#endif
