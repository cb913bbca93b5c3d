/* Host twins of the engine's ingestion kernel (part of libhsgpu.so's C ABI, declared in include/hsgpu.h): the arithmetic of
 * HectorMappingRos::rosLaserScanToDataContainer (src/hector_slam/hector_mapping/src/HectorMappingRos.cpp:483-507) for callers
 * that want DataContainer points on the host. Plain C, no CUDA. workload/libhsscangen.so compiles this file in as well, so that
 * the CPU-only arms (bench.py --impl reference, golden generation) never have to map libhsgpu.so. */
#define _GNU_SOURCE
#include "hsgpu.h"

#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

/* HectorMappingRos::rosLaserScanToDataContainer (HectorMappingRos.cpp:483-507): float angle accumulation,
 * strict (range_min, range_max - 0.1) window, dist *= scaleToMap, (cos(angle)*dist, sin(angle)*dist). */
int hs_scan_to_points_host(const float* ranges, int n_beams, float angle_min, float angle_increment,
                           float range_min, float range_max, float scale_to_map, float* points_out) {
  float angle = angle_min;
  float max_range_for_container = range_max - 0.1f;
  int n = 0;
  for (int i = 0; i < n_beams; ++i) {
    float dist = ranges[i];
    if ((dist > range_min) && (dist < max_range_for_container)) {
      dist *= scale_to_map;
      points_out[2 * n] = cosf(angle) * dist;
      points_out[2 * n + 1] = sinf(angle) * dist;
      ++n;
    }
    angle += angle_increment;
  }
  return n;
}

/* Threaded batch form of hs_scan_to_points_host: ranges [n_scans][n_beams] -> points [n_scans][max_points][2]
 * (zero padded), counts [n_scans]. */
typedef struct {
  const float* ranges; int n_scans, n_beams; float amin, ainc, rmin, rmax, scale; int max_points; float* pts; int32_t* cnt;
  int threads, tid;
} s2p_arg;

static void* s2p_worker(void* vp) {
  s2p_arg* a = (s2p_arg*)vp;
  int k0 = (int)((long long)a->n_scans * a->tid / a->threads), k1 = (int)((long long)a->n_scans * (a->tid + 1) / a->threads);
  float* tmp = (float*)malloc(sizeof(float) * 2 * (size_t)a->n_beams);
  for (int k = k0; k < k1; ++k) {
    int c = hs_scan_to_points_host(a->ranges + (size_t)k * a->n_beams, a->n_beams, a->amin, a->ainc, a->rmin, a->rmax, a->scale, tmp);
    if (c > a->max_points) c = a->max_points;
    float* out = a->pts + (size_t)k * a->max_points * 2;
    memcpy(out, tmp, sizeof(float) * 2 * (size_t)c);
    memset(out + 2 * (size_t)c, 0, sizeof(float) * 2 * (size_t)(a->max_points - c));
    a->cnt[k] = c;
  }
  free(tmp);
  return NULL;
}

int hs_scan_to_points_host_batch(const float* ranges, int n_scans, int n_beams, float angle_min, float angle_increment,
                                 float range_min, float range_max, float scale_to_map, int max_points,
                                 float* points_out, int32_t* counts_out, int threads) {
  if (!ranges || !points_out || !counts_out || n_scans < 0 || n_beams < 1 || max_points < 1) return HS_ERR_INVALID_ARG;
  if (threads < 1) threads = 1;
  if (threads > 256) threads = 256;
  if (threads > n_scans) threads = n_scans > 0 ? n_scans : 1;
  s2p_arg* args = (s2p_arg*)malloc(sizeof(s2p_arg) * threads);
  pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * threads);
  if (!args || !th) { free(args); free(th); return HS_ERR_OUT_OF_MEMORY; }
  for (int t = 0; t < threads; ++t) {
    s2p_arg a = { ranges, n_scans, n_beams, angle_min, angle_increment, range_min, range_max, scale_to_map, max_points, points_out, counts_out, threads, t };
    args[t] = a;
    if (t > 0) pthread_create(&th[t], NULL, s2p_worker, &args[t]);
  }
  s2p_worker(&args[0]);
  for (int t = 1; t < threads; ++t) pthread_join(th[t], NULL);
  free(args); free(th);
  return HS_OK;
}
