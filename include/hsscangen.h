/* hsscangen -- deterministic synthetic LaserScan streams: the workload generator shared by the
 * tests and by both arms of bench.py. Host code only, built into workload/libhsscangen.so; NOT part of the product
 * library (libhsgpu.so knows nothing about scenes) and NOT part of the oracle. The library also carries a copy of the
 * product's host conversion helpers hs_scan_to_points_host / hs_scan_to_points_host_batch (declared in hsgpu.h, source
 * tfe-slam-ucl_b200/csrc/hs_hostconv.c) so that CPU-only runs need no CUDA library at all. */
#ifndef HSSCANGEN_H
#define HSSCANGEN_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HSG_OK 0
#define HSG_ERR_INVALID_ARG (-1)
#define HSG_ERR_OUT_OF_MEMORY (-3)

typedef struct {
  int32_t n_beams;
  float angle_min, angle_increment;
  float range_min, range_max;
  float noise_sigma;       /* Gaussian range noise [m] */
  float room_w, room_h;    /* axis-aligned room centred at the origin [m] */
  int32_t n_boxes;         /* axis-aligned box obstacles */
  float box_min, box_max;  /* box edge length range [m] */
  float traj_rx, traj_ry;  /* ellipse radii of the closed-loop trajectory [m] */
  int32_t traj_period;     /* scans per loop */
} hs_scene_config;

#define HS_SCENE_RPLIDAR_ROOM 0 /* 360 beams 0..359 deg, 0.15-12 m, 10 m x 8 m room, 6 boxes (configs 1,2,5) */
#define HS_SCENE_HIRES_HALL 1   /* 1440 beams, 30 m, 40 m x 30 m hall, 12 boxes (config 3) */
int hs_scene_preset(int preset, hs_scene_config* out);
/* ranges [n_scans][n_streams][n_beams] for streams seed0 + s (s in [0, n_streams)) and scans
 * [first_scan, first_scan+n_scans); truth_poses [n_scans][n_streams][3] (relative to each stream's
 * first pose) or NULL. Deterministic in (cfg, seed, scan, beam); `threads` host threads. */
int hs_scangen_batch(const hs_scene_config* cfg, uint64_t seed0, int n_streams, int first_scan, int n_scans,
                     float* ranges, float* truth_poses, int threads);

#ifdef __cplusplus
}
#endif
#endif /* HSSCANGEN_H */
