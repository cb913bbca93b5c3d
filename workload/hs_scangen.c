/* Deterministic synthetic LaserScan streams -- the WORKLOAD GENERATOR of the tests and of bench.py (both arms), built into
 * workload/libhsscangen.so. Not part of the product library (libhsgpu.so) and not part of the oracle.
 *
 * Produces what the callers of the hot path feed it: sensor_msgs/LaserScan range arrays shaped like
 * the RPLidar driver's output (reference: src/rplidar_ros/src/node.cpp:76-82, :331-338 -- beams at
 * angle_min + i*angle_increment, range_min 0.15, range_max = scan-mode maximum), taken from a sensor that
 * drives a smooth closed loop through a rectangular room with box obstacles.
 * Everything is a pure function of (config, stream seed, scan index, beam index): no sequential RNG
 * state, so any batch of scans can be generated independently and in parallel.
 */
#define _GNU_SOURCE
#include "hsscangen.h"

#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

#define HS_MAX_BOXES 32

static uint64_t splitmix64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ULL;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ULL;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBULL;
  return x ^ (x >> 31);
}
static double u01(uint64_t h) { return ((double)(h >> 11) + 0.5) * (1.0 / 9007199254740992.0); }

typedef struct { double x0, y0, x1, y1; } seg_t;

typedef struct {
  int n_seg;
  seg_t seg[4 + 4 * HS_MAX_BOXES];
  double rx, ry, cx, cy, phase, dir;
  double start[3];
} scene_t;

static void traj_pose(const scene_t* sc, const hs_scene_config* cfg, int k, double* pose) {
  double phi = sc->phase + sc->dir * 2.0 * M_PI * (double)k / (double)cfg->traj_period;
  pose[0] = sc->cx + sc->rx * cos(phi);
  pose[1] = sc->cy + sc->ry * sin(phi);
  pose[2] = atan2(sc->dir * sc->ry * cos(phi), -sc->dir * sc->rx * sin(phi));
}

static void add_box(scene_t* sc, double cx, double cy, double w, double h) {
  double x0 = cx - 0.5 * w, x1 = cx + 0.5 * w, y0 = cy - 0.5 * h, y1 = cy + 0.5 * h;
  seg_t s[4] = { { x0, y0, x1, y0 }, { x1, y0, x1, y1 }, { x1, y1, x0, y1 }, { x0, y1, x0, y0 } };
  for (int i = 0; i < 4; ++i) sc->seg[sc->n_seg++] = s[i];
}

static void build_scene(const hs_scene_config* cfg, uint64_t seed, scene_t* sc) {
  uint64_t h = splitmix64(seed ^ 0xC0FFEEULL);
  memset(sc, 0, sizeof(*sc));
  sc->rx = cfg->traj_rx * (0.9 + 0.2 * u01(h = splitmix64(h)));
  sc->ry = cfg->traj_ry * (0.9 + 0.2 * u01(h = splitmix64(h)));
  sc->cx = 0.1 * cfg->room_w * (u01(h = splitmix64(h)) - 0.5);
  sc->cy = 0.1 * cfg->room_h * (u01(h = splitmix64(h)) - 0.5);
  sc->phase = 2.0 * M_PI * u01(h = splitmix64(h));
  sc->dir = (u01(h = splitmix64(h)) < 0.5) ? -1.0 : 1.0;
  add_box(sc, 0.0, 0.0, cfg->room_w, cfg->room_h); /* the room walls */
  traj_pose(sc, cfg, 0, sc->start);
  int n_boxes = cfg->n_boxes > HS_MAX_BOXES ? HS_MAX_BOXES : cfg->n_boxes;
  for (int b = 0; b < n_boxes; ++b) {
    for (int attempt = 0; attempt < 200; ++attempt) {
      double w = cfg->box_min + (cfg->box_max - cfg->box_min) * u01(h = splitmix64(h));
      double hh = cfg->box_min + (cfg->box_max - cfg->box_min) * u01(h = splitmix64(h));
      double bx = (cfg->room_w - w - 0.4) * (u01(h = splitmix64(h)) - 0.5);
      double by = (cfg->room_h - hh - 0.4) * (u01(h = splitmix64(h)) - 0.5);
      double clear = 0.5 + 0.5 * sqrt(w * w + hh * hh);
      int ok = 1;
      for (int k = 0; k < 128 && ok; ++k) { /* keep the whole path clear of the box */
        double phi = 2.0 * M_PI * k / 128.0;
        double px = sc->cx + sc->rx * cos(phi), py = sc->cy + sc->ry * sin(phi);
        if (hypot(px - bx, py - by) < clear) ok = 0;
      }
      if (ok && hypot(sc->start[0] - bx, sc->start[1] - by) < 1.0 + 0.5 * sqrt(w * w + hh * hh)) ok = 0;
      if (ok) { add_box(sc, bx, by, w, hh); break; }
    }
  }
}

static double cast_ray(const scene_t* sc, double px, double py, double dx, double dy) {
  double best = INFINITY;
  for (int i = 0; i < sc->n_seg; ++i) {
    const seg_t* s = &sc->seg[i];
    double ex = s->x1 - s->x0, ey = s->y1 - s->y0;
    double den = dx * ey - dy * ex;
    if (fabs(den) < 1e-12) continue;
    double wx = s->x0 - px, wy = s->y0 - py;
    double t = (wx * ey - wy * ex) / den;
    double u = (wx * dy - wy * dx) / den;
    if (t > 1e-9 && u >= 0.0 && u <= 1.0 && t < best) best = t;
  }
  return best;
}

static void gen_scan(const hs_scene_config* cfg, const scene_t* sc, uint64_t seed, int k, float* ranges, float* truth) {
  double pose[3];
  traj_pose(sc, cfg, k, pose);
  if (truth) { /* pose relative to the stream's first pose: the frame the SLAM estimate lives in */
    double c = cos(sc->start[2]), s = sin(sc->start[2]);
    double dx = pose[0] - sc->start[0], dy = pose[1] - sc->start[1];
    truth[0] = (float)(c * dx + s * dy);
    truth[1] = (float)(-s * dx + c * dy);
    double dth = pose[2] - sc->start[2];
    while (dth > M_PI) dth -= 2.0 * M_PI;
    while (dth < -M_PI) dth += 2.0 * M_PI;
    truth[2] = (float)dth;
  }
  for (int i = 0; i < cfg->n_beams; ++i) {
    double beam = (double)cfg->angle_min + (double)i * (double)cfg->angle_increment;
    double a = pose[2] + beam;
    double r = cast_ray(sc, pose[0], pose[1], cos(a), sin(a));
    uint64_t hh = splitmix64(seed * 0x100000001B3ULL + ((uint64_t)(uint32_t)k << 20) + (uint64_t)i);
    double u1 = u01(hh), u2 = u01(splitmix64(hh));
    double g = sqrt(-2.0 * log(u1)) * cos(2.0 * M_PI * u2);
    r += cfg->noise_sigma * g;
    ranges[i] = (r > (double)cfg->range_max || !(r == r)) ? INFINITY : (float)r;
  }
}

int hs_scene_preset(int preset, hs_scene_config* out) {
  if (!out) return HSG_ERR_INVALID_ARG;
  memset(out, 0, sizeof(*out));
  if (preset == HS_SCENE_RPLIDAR_ROOM) {
    out->n_beams = 360;
    out->angle_min = 0.0f;
    out->angle_increment = (float)((359.0 * M_PI / 180.0) / 359.0);
    out->range_min = 0.15f; out->range_max = 12.0f;
    out->noise_sigma = 0.01f;
    out->room_w = 10.0f; out->room_h = 8.0f;
    out->n_boxes = 6; out->box_min = 0.5f; out->box_max = 1.5f;
    out->traj_rx = 2.5f; out->traj_ry = 1.8f;
    out->traj_period = 1000;
    return HSG_OK;
  }
  if (preset == HS_SCENE_HIRES_HALL) {
    out->n_beams = 1440;
    out->angle_min = 0.0f;
    out->angle_increment = (float)((359.75 * M_PI / 180.0) / 1439.0);
    out->range_min = 0.15f; out->range_max = 30.0f;
    out->noise_sigma = 0.01f;
    out->room_w = 40.0f; out->room_h = 30.0f;
    out->n_boxes = 12; out->box_min = 1.0f; out->box_max = 4.0f;
    out->traj_rx = 8.0f; out->traj_ry = 6.0f;
    out->traj_period = 4000;
    return HSG_OK;
  }
  return HSG_ERR_INVALID_ARG;
}

typedef struct {
  const hs_scene_config* cfg; uint64_t seed0; int n_streams, first_scan, n_scans; float* ranges; float* truth;
  int threads, tid;
} gen_arg;

static void* gen_worker(void* vp) {
  gen_arg* a = (gen_arg*)vp;
  int s0 = (int)((long long)a->n_streams * a->tid / a->threads), s1 = (int)((long long)a->n_streams * (a->tid + 1) / a->threads);
  scene_t sc;
  for (int s = s0; s < s1; ++s) {
    uint64_t seed = a->seed0 + (uint64_t)s;
    build_scene(a->cfg, seed, &sc);
    for (int k = 0; k < a->n_scans; ++k) {
      size_t slot = (size_t)k * a->n_streams + s;
      gen_scan(a->cfg, &sc, seed, a->first_scan + k, a->ranges + slot * a->cfg->n_beams, a->truth ? a->truth + slot * 3 : NULL);
    }
  }
  return NULL;
}

int hs_scangen_batch(const hs_scene_config* cfg, uint64_t seed0, int n_streams, int first_scan, int n_scans,
                     float* ranges, float* truth_poses, int threads) {
  if (!cfg || !ranges || n_streams < 0 || n_scans < 0 || cfg->n_beams <= 0 || cfg->traj_period <= 0) return HSG_ERR_INVALID_ARG;
  if (threads < 1) threads = 1;
  if (threads > 256) threads = 256;
  if (threads > n_streams) threads = n_streams > 0 ? n_streams : 1;
  gen_arg* args = (gen_arg*)malloc(sizeof(gen_arg) * threads);
  pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * threads);
  if (!args || !th) { free(args); free(th); return HSG_ERR_OUT_OF_MEMORY; }
  for (int t = 0; t < threads; ++t) {
    gen_arg a = { cfg, seed0, n_streams, first_scan, n_scans, ranges, truth_poses, threads, t };
    args[t] = a;
    if (t > 0) pthread_create(&th[t], NULL, gen_worker, &args[t]);
  }
  gen_worker(&args[0]);
  for (int t = 1; t < threads; ++t) pthread_join(th[t], NULL);
  free(args); free(th);
  return HSG_OK;
}
