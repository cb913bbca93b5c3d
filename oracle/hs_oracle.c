/* TEST INFRASTRUCTURE ONLY -- see hs_oracle.h. Plain-C restatement of the hector_slam_lib
 * scan-to-map path; every function cites the reference lines it follows (HSL/ =
 * /root/reference/src/hector_slam/hector_mapping/include/hector_slam_lib/).
 * Build: gcc -O2 -ffp-contract=off (no FMA contraction, like the catkin default build).
 * All arithmetic is fp32 in the reference's operation order unless noted. */
#define _GNU_SOURCE
#include "hs_oracle.h"

#include <float.h>
#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#define HSO_MAX_LEVELS 8

typedef struct { float val; int index; } cache_elem; /* HSL/map/GridMapCacheArray.h:34-39 */

typedef struct {
  int sx, sy;
  float cell_length;
  float scale_to_map;
  /* mapTworld / worldTmap (HSL/map/GridMapBase.h:265-280), stored as linear 2x2 + translation */
  float mtw_l[4], mtw_t[2];
  float wtm_l[4], wtm_t[2];
  float limit_x, limit_y; /* mapLimitsf = dims - 2 (HSL/map/MapDimensionProperties.h:69-74) */
  hso_cell* cells;
  cache_elem* cache;
  int cache_epoch;      /* GridMapCacheArray::currCacheIndex */
  int curr_update_index; /* OccGridMapBase::currUpdateIndex (HSL/map/OccGridMapBase.h:265) */
  int mark_free, mark_occ;
  int last_update_index; /* GridMapBase::lastUpdateIndex (HSL/map/GridMapBase.h:399) */
  float log_odds_free, log_odds_occ;
} level_t;

typedef struct { float* xy; int n; int cap; float ox, oy; } container_t; /* HSL/scan/DataPointContainer.h */

struct hso_proc {
  int n_levels;
  level_t lv[HSO_MAX_LEVELS];
  container_t coarse[HSO_MAX_LEVELS]; /* MapRepMultiMap::dataContainers[level-1] */
  float last_map_update_pose[3];
  float last_scan_match_pose[3];
  float last_cov[9]; /* row-major */
  float min_dist, min_angle;
  unsigned long long cell_visits, integrations;
};

/* ------------------------------------------------------------------ helpers ---- */

static void container_reserve(container_t* c, int n) {
  if (n > c->cap) { c->xy = (float*)realloc(c->xy, sizeof(float) * 2 * (size_t)n); c->cap = n; }
}

/* GridMapLogOddsFunctions::probToLogOdds (HSL/map/GridMapLogOdds.h:198-203) */
static float prob_to_log_odds(float prob) {
  float odds = prob / (1.0f - prob);
  return logf(odds);
}

/* GridMapLogOddsFunctions::getGridProbability (HSL/map/GridMapLogOdds.h:163-166) */
static float grid_probability(float log_odds) {
  float odds = expf(log_odds);
  return odds / (odds + 1.0f);
}

/* GridMapBase::setMapTransformation (HSL/map/GridMapBase.h:265-280) with Eigen 3.3's
 * DiagonalMatrix*Translation and Affine inverse (2x2 inverse * (1/det), t' = (-Linv) t). */
static void set_map_transformation(level_t* g, float offx, float offy, float cell_length) {
  g->cell_length = cell_length;
  g->scale_to_map = 1.0f / cell_length;
  float s = g->scale_to_map;
  g->mtw_l[0] = s; g->mtw_l[1] = 0.0f; g->mtw_l[2] = 0.0f; g->mtw_l[3] = s; /* row-major [00 01 10 11] */
  g->mtw_t[0] = s * offx; g->mtw_t[1] = s * offy;
  float m00 = g->mtw_l[0], m01 = g->mtw_l[1], m10 = g->mtw_l[2], m11 = g->mtw_l[3];
  float det = m00 * m11 - m10 * m01;
  float invdet = 1.0f / det;
  float i00 = m11 * invdet, i10 = -m10 * invdet, i01 = -m01 * invdet, i11 = m00 * invdet;
  g->wtm_l[0] = i00; g->wtm_l[1] = i01; g->wtm_l[2] = i10; g->wtm_l[3] = i11;
  g->wtm_t[0] = (-i00) * g->mtw_t[0] + (-i01) * g->mtw_t[1];
  g->wtm_t[1] = (-i10) * g->mtw_t[0] + (-i11) * g->mtw_t[1];
}

/* Affine2f * Vector2f = linear*v + translation */
static void affine_apply(const float* l, const float* t, float x, float y, float* ox, float* oy) {
  *ox = (l[0] * x + l[1] * y) + t[0];
  *oy = (l[2] * x + l[3] * y) + t[1];
}

/* GridMapBase::getMapCoordsPose / getWorldCoordsPose (HSL/map/GridMapBase.h:226-239) */
static void map_coords_pose(const level_t* g, const float* w, float* m) {
  affine_apply(g->mtw_l, g->mtw_t, w[0], w[1], &m[0], &m[1]);
  m[2] = w[2];
}
static void world_coords_pose(const level_t* g, const float* m, float* w) {
  affine_apply(g->wtm_l, g->wtm_t, m[0], m[1], &w[0], &w[1]);
  w[2] = m[2];
}

/* util::normalize_angle_pos / normalize_angle (HSL/util/UtilFunctions.h:37-49): the
 * literals 2.0f*M_PI are double, so fmod runs in double; results narrow to float. */
static float normalize_angle_pos(float angle) {
  return (float)fmod(fmod((double)angle, 2.0f * M_PI) + 2.0f * M_PI, 2.0f * M_PI);
}
static float normalize_angle(float angle) {
  float a = normalize_angle_pos(angle);
  if (a > M_PI) a -= 2.0f * M_PI;
  return a;
}

/* util::poseDifferenceLargerThan (HSL/util/UtilFunctions.h:73-92); abs() is the float overload */
static int pose_difference_larger_than(const float* p1, const float* p2, float dist_thresh, float ang_thresh) {
  float dx = p1[0] - p2[0], dy = p1[1] - p2[1];
  if (sqrtf(dx * dx + dy * dy) > dist_thresh) return 1;
  float angle_diff = p1[2] - p2[2];
  if (angle_diff > M_PI) angle_diff -= M_PI * 2.0f;
  else if (angle_diff < -M_PI) angle_diff += M_PI * 2.0f;
  if (fabsf(angle_diff) > ang_thresh) return 1;
  return 0;
}

/* ------------------------------------------------------------- construction ---- */

/* GridMapBase::clear / LogOddsCell::resetGridCell (HSL/map/GridMapBase.h:69-87, GridMapLogOdds.h:89-93) */
static void level_clear(level_t* g) {
  size_t n = (size_t)g->sx * g->sy;
  for (size_t i = 0; i < n; ++i) { g->cells[i].logOddsVal = 0.0f; g->cells[i].updateIndex = -1; }
}

/* HectorSlamProcessor::reset (HSL/slam_main/HectorSlamProcessor.h:115-124) -> MapProcContainer::reset
 * (HSL/slam_main/MapProcContainer.h:67-71). currUpdateIndex / lastUpdateIndex are NOT reset. */
void hso_reset(hso_proc* p) {
  p->last_map_update_pose[0] = p->last_map_update_pose[1] = p->last_map_update_pose[2] = FLT_MAX;
  p->last_scan_match_pose[0] = p->last_scan_match_pose[1] = p->last_scan_match_pose[2] = 0.0f;
  for (int l = 0; l < p->n_levels; ++l) { level_clear(&p->lv[l]); p->lv[l].cache_epoch++; }
}

/* HectorSlamProcessor ctor (HSL/slam_main/HectorSlamProcessor.h:54-64), MapRepMultiMap ctor
 * (HSL/slam_main/MapRepMultiMap.h:48-72), GridMapBase ctor (HSL/map/GridMapBase.h:95-108),
 * OccGridMapBase ctor (OccGridMapBase.h:48-53), GridMapLogOddsFunctions ctor (GridMapLogOdds.h:115-119),
 * GridMapCacheArray (GridMapCacheArray.h:46-52, 119-133). */
hso_proc* hso_create(float res, int sx, int sy, float startx, float starty, int levels) {
  if (levels < 1 || levels > HSO_MAX_LEVELS) return NULL;
  hso_proc* p = (hso_proc*)calloc(1, sizeof(hso_proc));
  p->n_levels = levels;
  float total_x = res * (float)sx, mid_x = total_x * startx;
  float total_y = res * (float)sy, mid_y = total_y * starty;
  for (int i = 0; i < levels; ++i) {
    level_t* g = &p->lv[i];
    g->sx = sx; g->sy = sy;
    g->limit_x = (float)sx - 2.0f; g->limit_y = (float)sy - 2.0f;
    size_t n = (size_t)sx * sy;
    g->cells = (hso_cell*)malloc(sizeof(hso_cell) * n);
    g->cache = (cache_elem*)malloc(sizeof(cache_elem) * n);
    for (size_t k = 0; k < n; ++k) g->cache[k].index = -1;
    g->cache_epoch = 0;
    g->curr_update_index = 0; g->mark_occ = -1; g->mark_free = -1;
    g->last_update_index = -1;
    g->log_odds_free = prob_to_log_odds(0.4f);
    g->log_odds_occ = prob_to_log_odds(0.6f);
    set_map_transformation(g, mid_x, mid_y, res);
    level_clear(g);
    sx /= 2; sy /= 2;
    res *= 2.0f;
  }
  hso_reset(p);
  p->min_dist = 0.4f * 1.0f;
  p->min_angle = 0.13f * 1.0f;
  return p;
}

void hso_destroy(hso_proc* p) {
  if (!p) return;
  for (int l = 0; l < p->n_levels; ++l) { free(p->lv[l].cells); free(p->lv[l].cache); }
  for (int l = 0; l < HSO_MAX_LEVELS; ++l) free(p->coarse[l].xy);
  free(p);
}

/* MapRepMultiMap::setUpdateFactorFree/Occupied (HSL/slam_main/MapRepMultiMap.h:149-167) */
void hso_set_update_factor_free(hso_proc* p, float f) { for (int l = 0; l < p->n_levels; ++l) p->lv[l].log_odds_free = prob_to_log_odds(f); }
void hso_set_update_factor_occupied(hso_proc* p, float f) { for (int l = 0; l < p->n_levels; ++l) p->lv[l].log_odds_occ = prob_to_log_odds(f); }
void hso_set_map_update_min_dist_diff(hso_proc* p, float d) { p->min_dist = d; }
void hso_set_map_update_min_angle_diff(hso_proc* p, float a) { p->min_angle = a; }
float hso_get_scale_to_map(hso_proc* p) { return p->lv[0].scale_to_map; }
int hso_get_map_levels(hso_proc* p) { return p->n_levels; }

/* ------------------------------------------------------------------ matching ---- */

/* GridMapCacheArray::containsCachedData / cacheData (HSL/map/GridMapCacheArray.h:80-102) around
 * OccGridMapBase::getGridProbabilityMap (OccGridMapBase.h:74). Pure memoisation. */
static float cached_probability(level_t* g, int index) {
  cache_elem* e = &g->cache[index];
  if (e->index == g->cache_epoch) return e->val;
  float v = grid_probability(g->cells[index].logOddsVal);
  e->index = g->cache_epoch;
  e->val = v;
  return v;
}

/* OccGridMapUtil::interpMapValueWithDerivatives (HSL/map/OccGridMapUtil.h:287-347).
 * Note the reference's gradient blend: dx blends rows with the x factors, dy blends columns
 * with the y factors (lines 344-345) -- kept as is. */
static void interp_with_derivs(level_t* g, float cx, float cy, float* out3) {
  if ((cx < 0.0f) || (cx > g->limit_x) || (cy < 0.0f) || (cy > g->limit_y)) { /* MapDimensionProperties.h:65-68 */
    out3[0] = 0.0f; out3[1] = 0.0f; out3[2] = 0.0f;
    return;
  }
  int ix = (int)cx, iy = (int)cy;
  float fx = cx - (float)ix, fy = cy - (float)iy;
  int index = iy * g->sx + ix;
  float i0 = cached_probability(g, index);
  ++index;
  float i1 = cached_probability(g, index);
  index += g->sx - 1;
  float i2 = cached_probability(g, index);
  ++index;
  float i3 = cached_probability(g, index);
  float dx1 = i0 - i1, dx2 = i2 - i3;
  float dy1 = i0 - i2, dy2 = i1 - i3;
  float x_inv = 1.0f - fx, y_inv = 1.0f - fy;
  out3[0] = ((i0 * x_inv + i1 * fx) * y_inv) + ((i2 * x_inv + i3 * fx) * fy);
  out3[1] = -((dx1 * x_inv) + (dx2 * fx));
  out3[2] = -((dy1 * y_inv) + (dy2 * fy));
}

/* OccGridMapUtil::interpMapValue (HSL/map/OccGridMapUtil.h:241-285): value only. */
static float interp_map_value(level_t* g, float cx, float cy) {
  if ((cx < 0.0f) || (cx > g->limit_x) || (cy < 0.0f) || (cy > g->limit_y)) return 0.0f;
  int ix = (int)cx, iy = (int)cy;
  float fx = cx - (float)ix, fy = cy - (float)iy;
  int index = iy * g->sx + ix;
  float i0 = cached_probability(g, index);
  ++index;
  float i1 = cached_probability(g, index);
  index += g->sx - 1;
  float i2 = cached_probability(g, index);
  ++index;
  float i3 = cached_probability(g, index);
  float x_inv = 1.0f - fx, y_inv = 1.0f - fy;
  return ((i0 * x_inv + i1 * fx) * y_inv) + ((i2 * x_inv + i3 * fx) * fy);
}

/* OccGridMapUtil::getResidualForState (HSL/map/OccGridMapUtil.h:216-233): sequential fp32 sum of 1 - interp. */
static float residual_for_state(level_t* g, const float* pose, const float* pts, int n) {
  float sin_rot = sinf(pose[2]);
  float cos_rot = cosf(pose[2]);
  float l[4] = { cos_rot, -sin_rot, sin_rot, cos_rot };
  float t[2] = { pose[0], pose[1] };
  float residual = 0.0f;
  for (int i = 0; i < n; ++i) {
    float qx, qy;
    affine_apply(l, t, pts[2 * i], pts[2 * i + 1], &qx, &qy);
    float funval = 1.0f - interp_map_value(g, qx, qy);
    residual += funval;
  }
  return residual;
}

/* getLikelihoodForState / getLikelihoodForResidual (OccGridMapUtil.h:201-214): 1 - residual / n. */
static float likelihood_for_state(level_t* g, const float* pose, const float* pts, int n) {
  float resid = residual_for_state(g, pose, pts, n);
  float sizef = (float)n;
  return 1 - (resid / sizef);
}

/* OccGridMapUtil::getCovarianceForPose (OccGridMapUtil.h:106-160): 7 sigma points weighted by their likelihoods.
 * Eigen semantics as restated in oracle/shim: likelihoods.sum() is the unrolled 7-term redux
 * (a0 + (a1 + a2)) + ((a3 + a4) + (a5 + a6)); mean += sigma * lh elementwise;
 * cov += (lh * inv) * (d * d^T) with the outer product formed first. cov row-major. */
static void covariance_for_pose(level_t* g, const float* pose, const float* pts, int n, float* cov9) {
  const float dtx = 1.5f, dty = 1.5f, dang = 0.05f;
  float x = pose[0], y = pose[1], ang = pose[2];
  float sp[7][3] = { { x + dtx, y, ang }, { x - dtx, y, ang }, { x, y + dty, ang }, { x, y - dty, ang },
                     { x, y, ang + dang }, { x, y, ang - dang }, { x, y, ang } };
  float lh[7];
  for (int i = 0; i < 7; ++i) lh[i] = likelihood_for_state(g, sp[i], pts, n);
  float sum = (lh[0] + (lh[1] + lh[2])) + ((lh[3] + lh[4]) + (lh[5] + lh[6]));
  float inv = 1 / sum;
  float mean[3] = { 0.0f, 0.0f, 0.0f };
  for (int i = 0; i < 7; ++i)
    for (int k = 0; k < 3; ++k) mean[k] += sp[i][k] * lh[i];
  for (int k = 0; k < 3; ++k) mean[k] *= inv;
  for (int k = 0; k < 9; ++k) cov9[k] = 0.0f;
  for (int i = 0; i < 7; ++i) {
    float d[3] = { sp[i][0] - mean[0], sp[i][1] - mean[1], sp[i][2] - mean[2] };
    float w = lh[i] * inv;
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 3; ++c) cov9[3 * r + c] += w * (d[r] * d[c]);
  }
}

/* OccGridMapUtil::getCovMatrixWorldCoords (OccGridMapUtil.h:162-189). Only the entries the reference assigns. */
static void cov_matrix_world_coords(const level_t* g, const float* m, float* w) {
  float scale = g->cell_length;
  float scale_sq = scale * scale;
  w[0] = m[0] * scale_sq;
  w[4] = m[4] * scale_sq;
  w[3] = m[3] * scale_sq;
  w[1] = w[3];
  w[6] = m[6] * scale;
  w[2] = w[6];
  w[7] = m[7] * scale;
  w[5] = w[7];
  w[8] = m[8];
}

/* OccGridMapUtil::getCompleteHessianDerivs (HSL/map/OccGridMapUtil.h:64-104). H row-major 3x3. */
static void complete_hessian_derivs(level_t* g, const float* pose, const float* pts, int n, float* H, float* dTr) {
  float sin_rot = sinf(pose[2]);
  float cos_rot = cosf(pose[2]);
  /* getTransformForState: Translation2f(x,y) * Rotation2Df(theta) -> linear [c -s; s c] */
  float l[4] = { cos_rot, -sin_rot, sin_rot, cos_rot };
  float t[2] = { pose[0], pose[1] };
  for (int i = 0; i < 9; ++i) H[i] = 0.0f;
  dTr[0] = dTr[1] = dTr[2] = 0.0f;
  for (int i = 0; i < n; ++i) {
    float px = pts[2 * i], py = pts[2 * i + 1];
    float qx, qy, tp[3];
    affine_apply(l, t, px, py, &qx, &qy);
    interp_with_derivs(g, qx, qy, tp);
    float fun_val = 1.0f - tp[0];
    dTr[0] += tp[1] * fun_val;
    dTr[1] += tp[2] * fun_val;
    float rot_deriv = ((-sin_rot * px - cos_rot * py) * tp[1] + (cos_rot * px - sin_rot * py) * tp[2]);
    dTr[2] += rot_deriv * fun_val;
    H[0] += tp[1] * tp[1];
    H[4] += tp[2] * tp[2];
    H[8] += rot_deriv * rot_deriv;
    H[1] += tp[1] * tp[2];
    H[2] += tp[1] * rot_deriv;
    H[5] += tp[2] * rot_deriv;
  }
  H[3] = H[1];
  H[6] = H[2];
  H[7] = H[5];
}

/* Eigen 3.3 Matrix3f::inverse (cofactor method; see oracle/shim/Eigen/LU). m, r row-major. */
static float cofactor3(const float* m, int i, int j) {
  int i1 = (i + 1) % 3, i2 = (i + 2) % 3, j1 = (j + 1) % 3, j2 = (j + 2) % 3;
  return m[3 * i1 + j1] * m[3 * i2 + j2] - m[3 * i1 + j2] * m[3 * i2 + j1];
}
static void inverse3(const float* m, float* r) {
  float c00 = cofactor3(m, 0, 0), c10 = cofactor3(m, 1, 0), c20 = cofactor3(m, 2, 0);
  float det = c00 * m[0] + (c10 * m[3] + c20 * m[6]);
  float invdet = 1.0f / det;
  r[0] = c00 * invdet; r[1] = c10 * invdet; r[2] = c20 * invdet;
  r[3] = cofactor3(m, 0, 1) * invdet; r[4] = cofactor3(m, 1, 1) * invdet; r[5] = cofactor3(m, 2, 1) * invdet;
  r[6] = cofactor3(m, 0, 2) * invdet; r[7] = cofactor3(m, 1, 2) * invdet; r[8] = cofactor3(m, 2, 2) * invdet;
}

/* ScanMatcher::estimateTransformationLogLh + updateEstimatedPose (HSL/matcher/ScanMatcher.h:194-226) */
static int estimate_transformation(level_t* g, float* estimate, const float* pts, int n, float* H, float* dTr) {
  complete_hessian_derivs(g, estimate, pts, n, H, dTr);
  if ((H[0] != 0.0f) && (H[4] != 0.0f)) {
    float inv[9], sd[3];
    inverse3(H, inv);
    for (int i = 0; i < 3; ++i) sd[i] = inv[3 * i] * dTr[0] + (inv[3 * i + 1] * dTr[1] + inv[3 * i + 2] * dTr[2]);
    if (sd[2] > 0.2f) sd[2] = 0.2f;
    else if (sd[2] < -0.2f) sd[2] = -0.2f;
    estimate[0] += sd[0]; estimate[1] += sd[1]; estimate[2] += sd[2];
    return 1;
  }
  return 0;
}

/* ScanMatcher::matchData (HSL/matcher/ScanMatcher.h:54-190): always 1 + maxIterations evaluations,
 * no convergence exit; cov = last H; empty container returns the hint and leaves cov untouched. */
static void match_data(level_t* g, const float* begin_world, const float* pts, int n, int max_iter, float* out_world, float* cov9) {
  if (n != 0) {
    float estimate[3], H[9], dTr[3];
    map_coords_pose(g, begin_world, estimate);
    estimate_transformation(g, estimate, pts, n, H, dTr);
    for (int i = 0; i < max_iter; ++i) estimate_transformation(g, estimate, pts, n, H, dTr);
    estimate[2] = normalize_angle(estimate[2]);
    memcpy(cov9, H, sizeof(float) * 9);
    world_coords_pose(g, estimate, out_world);
    return;
  }
  out_world[0] = begin_world[0]; out_world[1] = begin_world[1]; out_world[2] = begin_world[2];
}

/* DataPointContainer::setFrom (HSL/scan/DataPointContainer.h:46-58) */
static void container_set_from(container_t* c, const float* pts, int n, float ox, float oy, float factor) {
  c->ox = ox * factor; c->oy = oy * factor;
  container_reserve(c, n);
  c->n = n;
  for (int i = 0; i < 2 * n; ++i) c->xy[i] = pts[i] * factor;
}

/* MapRepMultiMap::matchData (HSL/slam_main/MapRepMultiMap.h:116-132) */
static void multi_match_data(hso_proc* p, const float* begin_world, const float* pts, int n, float ox, float oy, float* out_world, float* cov9) {
  float tmp[3] = { begin_world[0], begin_world[1], begin_world[2] };
  for (int index = p->n_levels - 1; index >= 0; --index) {
    float res[3];
    if (index == 0) {
      match_data(&p->lv[0], tmp, pts, n, 5, res, cov9);
    } else {
      float factor = (float)(1.0 / pow(2.0, (double)index));
      container_set_from(&p->coarse[index - 1], pts, n, ox, oy, factor);
      match_data(&p->lv[index], tmp, p->coarse[index - 1].xy, p->coarse[index - 1].n, 3, res, cov9);
    }
    tmp[0] = res[0]; tmp[1] = res[1]; tmp[2] = res[2];
  }
  out_world[0] = tmp[0]; out_world[1] = tmp[1]; out_world[2] = tmp[2];
}

/* --------------------------------------------------------------- integration ---- */

/* OccGridMapBase::bresenhamCellFree + GridMapLogOddsFunctions::updateSetFree
 * (HSL/map/OccGridMapBase.h:216-224, GridMapLogOdds.h:146-151) */
static inline void bresenham_cell_free(level_t* g, unsigned int offset) {
  hso_cell* cell = &g->cells[offset];
  if (cell->updateIndex < g->mark_free) {
    cell->logOddsVal += g->log_odds_free;
    cell->updateIndex = g->mark_free;
  }
}

/* OccGridMapBase::bresenhamCellOcc + updateUnsetFree/updateSetOccupied
 * (HSL/map/OccGridMapBase.h:226-241, GridMapLogOdds.h:135-156) */
static inline void bresenham_cell_occ(level_t* g, unsigned int offset) {
  hso_cell* cell = &g->cells[offset];
  if (cell->updateIndex < g->mark_occ) {
    if (cell->updateIndex == g->mark_free) cell->logOddsVal -= g->log_odds_free;
    if (cell->logOddsVal < 50.0f) cell->logOddsVal += g->log_odds_occ;
    cell->updateIndex = g->mark_occ;
  }
}

/* OccGridMapBase::bresenham2D (HSL/map/OccGridMapBase.h:243-260) */
static void bresenham_2d(level_t* g, unsigned int abs_da, unsigned int abs_db, int error_b, int offset_a, int offset_b, unsigned int offset) {
  bresenham_cell_free(g, offset);
  unsigned int end = abs_da - 1;
  for (unsigned int i = 0; i < end; ++i) {
    offset += offset_a;
    error_b += abs_db;
    if ((unsigned int)error_b >= abs_da) {
      offset += offset_b;
      error_b -= abs_da;
    }
    bresenham_cell_free(g, offset);
  }
}

/* util::sign (HSL/util/UtilFunctions.h:56-59): sign(0) = -1 */
static inline int sign_i(int x) { return x > 0 ? 1 : -1; }

/* OccGridMapBase::updateLineBresenhami (HSL/map/OccGridMapBase.h:170-214) */
static void update_line_bresenhami(hso_proc* p, level_t* g, int x0, int y0, int x1, int y1) {
  if ((x0 < 0) || (x0 >= g->sx) || (y0 < 0) || (y0 >= g->sy)) return;
  if ((x1 < 0) || (x1 >= g->sx) || (y1 < 0) || (y1 >= g->sy)) return;
  int dx = x1 - x0, dy = y1 - y0;
  unsigned int abs_dx = abs(dx), abs_dy = abs(dy);
  int offset_dx = sign_i(dx);
  int offset_dy = sign_i(dy) * g->sx;
  unsigned int start_offset = y0 * g->sx + x0;
  if (abs_dx >= abs_dy) {
    int error_y = abs_dx / 2;
    bresenham_2d(g, abs_dx, abs_dy, error_y, offset_dx, offset_dy, start_offset);
    p->cell_visits += abs_dx + 1;
  } else {
    int error_x = abs_dy / 2;
    bresenham_2d(g, abs_dy, abs_dx, error_x, offset_dy, offset_dx, start_offset);
    p->cell_visits += abs_dy + 1;
  }
  unsigned int end_offset = y1 * g->sx + x1;
  bresenham_cell_occ(g, end_offset);
}

/* OccGridMapBase::updateByScan (HSL/map/OccGridMapBase.h:121-168). float->int conversions
 * truncate toward zero (Vector2i(float,float) ctor at :137, cast<int>() at :155). */
static void update_by_scan(hso_proc* p, level_t* g, const float* pts, int n, float ox, float oy, const float* pose_world) {
  g->mark_free = g->curr_update_index + 1;
  g->mark_occ = g->curr_update_index + 2;
  float map_pose[3];
  map_coords_pose(g, pose_world, map_pose);
  float s = sinf(map_pose[2]), c = cosf(map_pose[2]);
  float l[4] = { c, -s, s, c };
  float t[2] = { map_pose[0], map_pose[1] };
  float bx, by;
  affine_apply(l, t, ox, oy, &bx, &by);
  int begin_x = (int)(bx + 0.5f), begin_y = (int)(by + 0.5f);
  for (int i = 0; i < n; ++i) {
    float ex, ey;
    affine_apply(l, t, pts[2 * i], pts[2 * i + 1], &ex, &ey);
    ex += 0.5f; ey += 0.5f;
    int end_x = (int)ex, end_y = (int)ey;
    if (begin_x != end_x || begin_y != end_y) update_line_bresenhami(p, g, begin_x, begin_y, end_x, end_y);
  }
  g->last_update_index++; /* setUpdated() */
  g->curr_update_index += 3;
}

/* MapRepMultiMap::updateByScan (HSL/slam_main/MapRepMultiMap.h:134-147): level 0 gets the scan,
 * level i>0 gets dataContainers[i-1] as last refreshed by matchData (possibly stale/empty). */
static void multi_update_by_scan(hso_proc* p, const float* pts, int n, float ox, float oy, const float* pose_world) {
  for (int i = 0; i < p->n_levels; ++i) {
    if (i == 0) update_by_scan(p, &p->lv[0], pts, n, ox, oy, pose_world);
    else update_by_scan(p, &p->lv[i], p->coarse[i - 1].xy, p->coarse[i - 1].n, p->coarse[i - 1].ox, p->coarse[i - 1].oy, pose_world);
  }
  p->integrations++;
}

/* MapRepMultiMap::onMapUpdated (HSL/slam_main/MapRepMultiMap.h:107-114) */
static void on_map_updated(hso_proc* p) { for (int l = 0; l < p->n_levels; ++l) p->lv[l].cache_epoch++; }

/* HectorSlamProcessor::update (HSL/slam_main/HectorSlamProcessor.h:71-113) */
void hso_update(hso_proc* p, const float* pts, int n, float ox, float oy, const float* hint, int map_without_matching) {
  float new_pose[3];
  if (!map_without_matching) multi_match_data(p, hint, pts, n, ox, oy, new_pose, p->last_cov);
  else { new_pose[0] = hint[0]; new_pose[1] = hint[1]; new_pose[2] = hint[2]; }
  p->last_scan_match_pose[0] = new_pose[0]; p->last_scan_match_pose[1] = new_pose[1]; p->last_scan_match_pose[2] = new_pose[2];
  if (pose_difference_larger_than(new_pose, p->last_map_update_pose, p->min_dist, p->min_angle) || map_without_matching) {
    multi_update_by_scan(p, pts, n, ox, oy, new_pose);
    on_map_updated(p);
    p->last_map_update_pose[0] = new_pose[0]; p->last_map_update_pose[1] = new_pose[1]; p->last_map_update_pose[2] = new_pose[2];
  }
}

/* ------------------------------------------------------------------ accessors ---- */

void hso_get_pose(hso_proc* p, float* out3) { memcpy(out3, p->last_scan_match_pose, sizeof(float) * 3); }
void hso_get_last_map_update_pose(hso_proc* p, float* out3) { memcpy(out3, p->last_map_update_pose, sizeof(float) * 3); }
void hso_get_cov(hso_proc* p, float* out9) { memcpy(out9, p->last_cov, sizeof(float) * 9); }
void hso_level_dims(hso_proc* p, int level, int* sx, int* sy, float* cell_length) {
  *sx = p->lv[level].sx; *sy = p->lv[level].sy; *cell_length = p->lv[level].cell_length;
}
int hso_get_update_index(hso_proc* p, int level) { return p->lv[level].last_update_index; }
void hso_download_level(hso_proc* p, int level, hso_cell* out) {
  memcpy(out, p->lv[level].cells, sizeof(hso_cell) * (size_t)p->lv[level].sx * p->lv[level].sy);
}
void hso_upload_level(hso_proc* p, int level, const hso_cell* in) {
  memcpy(p->lv[level].cells, in, sizeof(hso_cell) * (size_t)p->lv[level].sx * p->lv[level].sy);
  p->lv[level].cache_epoch++;
}
void hso_set_cell(hso_proc* p, int level, int index, float logodds, int update_index) {
  p->lv[level].cells[index].logOddsVal = logodds; p->lv[level].cells[index].updateIndex = update_index;
  p->lv[level].cache_epoch++;
}
void hso_world_to_map(hso_proc* p, int level, const float* w, float* m) { map_coords_pose(&p->lv[level], w, m); }
void hso_map_to_world(hso_proc* p, int level, const float* m, float* w) { world_coords_pose(&p->lv[level], m, w); }

void hso_hessian_derivs(hso_proc* p, int level, const float* pose_map, const float* pts, int n, float* H9, float* dTr3) {
  complete_hessian_derivs(&p->lv[level], pose_map, pts, n, H9, dTr3);
}
void hso_hessian_derivs_batch(hso_proc* p, int level, const float* poses_map, int n_poses, const float* pts, int n, float* H9, float* dTr3) {
  for (int k = 0; k < n_poses; ++k) complete_hessian_derivs(&p->lv[level], poses_map + 3 * k, pts, n, H9 + 9 * k, dTr3 + 3 * k);
}
void hso_interp_with_derivs(hso_proc* p, int level, float x, float y, float* out3) { interp_with_derivs(&p->lv[level], x, y, out3); }
void hso_match_level(hso_proc* p, int level, const float* hint_world, const float* pts, int n, int max_iter, float* pose_out, float* cov9) {
  for (int i = 0; i < 9; ++i) cov9[i] = 0.0f;
  match_data(&p->lv[level], hint_world, pts, n, max_iter, pose_out, cov9);
}
void hso_update_by_scan_level(hso_proc* p, int level, const float* pts, int n, float ox, float oy, const float* pose_world) {
  update_by_scan(p, &p->lv[level], pts, n, ox, oy, pose_world);
  p->lv[level].cache_epoch++;
}
void hso_refresh_coarse(hso_proc* p, const float* pts, int n, float ox, float oy) {
  for (int index = p->n_levels - 1; index >= 1; --index)
    container_set_from(&p->coarse[index - 1], pts, n, ox, oy, (float)(1.0 / pow(2.0, (double)index)));
}
void hso_update_by_scan_all(hso_proc* p, const float* pts, int n, float ox, float oy, const float* pose_world) {
  multi_update_by_scan(p, pts, n, ox, oy, pose_world);
  on_map_updated(p);
}
float hso_interp_map_value(hso_proc* p, int level, float x, float y) { return interp_map_value(&p->lv[level], x, y); }
float hso_residual_for_state(hso_proc* p, int level, const float* pose_map, const float* pts, int n) { return residual_for_state(&p->lv[level], pose_map, pts, n); }
float hso_likelihood_for_state(hso_proc* p, int level, const float* pose_map, const float* pts, int n) { return likelihood_for_state(&p->lv[level], pose_map, pts, n); }
void hso_covariance_for_pose(hso_proc* p, int level, const float* pose_map, const float* pts, int n, float* cov_map9, float* cov_world9) {
  covariance_for_pose(&p->lv[level], pose_map, pts, n, cov_map9);
  if (cov_world9) cov_matrix_world_coords(&p->lv[level], cov_map9, cov_world9);
}

/* HectorMapTools::DistanceMeasurementProvider::checkOccupancyBresenhami + bresenham2D
 * (src/hector_slam/hector_map_tools/include/hector_map_tools/HectorMapTools.h:148-232) on an occupancy grid
 * {-1, 0, 100}: first cell == 100 within min(5000, abs_da) steps from the begin cell (the end cell itself is not tested);
 * returns (float)(int)|begin - hit| or -1. hit_xy is written only on a hit. */
float hso_check_occupancy_bresenhami(const signed char* grid, int size_x, int size_y, int x0, int y0, int x1, int y1, int* hit_xy) {
  if ((x0 < 0) || (x0 >= size_x) || (y0 < 0) || (y0 >= size_y)) return -1.0f;
  if ((x1 < 0) || (x1 >= size_x) || (y1 < 0) || (y1 >= size_y)) return -1.0f;
  int dx = x1 - x0, dy = y1 - y0;
  unsigned int abs_dx = (unsigned int)abs(dx), abs_dy = (unsigned int)abs(dy);
  int offset_dx = dx > 0 ? 1 : -1;
  int offset_dy = (dy > 0 ? 1 : -1) * size_x;
  unsigned int offset = (unsigned int)(y0 * size_x + x0);
  unsigned int abs_da, abs_db;
  int offset_a, offset_b;
  if (abs_dx >= abs_dy) { abs_da = abs_dx; abs_db = abs_dy; offset_a = offset_dx; offset_b = offset_dy; }
  else { abs_da = abs_dy; abs_db = abs_dx; offset_a = offset_dy; offset_b = offset_dx; }
  int error_b = (int)(abs_da / 2);
  unsigned int end = abs_da < 5000u ? abs_da : 5000u;
  int end_offset = -1;
  for (unsigned int i = 0; i < end; ++i) {
    if (grid[offset] == 100) { end_offset = (int)offset; break; }
    offset += (unsigned int)offset_a;
    error_b += (int)abs_db;
    if ((unsigned int)error_b >= abs_da) { offset += (unsigned int)offset_b; error_b -= (int)abs_da; }
  }
  if (end_offset == -1) return -1.0f;
  int ex = end_offset % size_x, ey = end_offset / size_x;
  float fx = (float)(x0 - ex), fy = (float)(y0 - ey);
  int dist_map = (int)sqrtf(fx * fx + fy * fy);   /* ((begin - end).cast<float>()).norm() truncated to int */
  if (hit_xy) { hit_xy[0] = ex; hit_xy[1] = ey; }
  return (float)dist_map;
}

int hso_pose_difference_larger_than(const float* p1, const float* p2, float d, float a) { return pose_difference_larger_than(p1, p2, d, a); }
float hso_normalize_angle(float a) { return normalize_angle(a); }
unsigned long long hso_get_cell_visits(hso_proc* p) { return p->cell_visits; }
unsigned long long hso_get_integrations(hso_proc* p) { return p->integrations; }

/* HectorMappingRos::rosLaserScanToDataContainer (src/hector_slam/hector_mapping/src/HectorMappingRos.cpp:483-507):
 * float angle accumulation, strict range window (min, max - 0.1), scale, (cos*dist, sin*dist). */
int hso_scan_to_points(const float* ranges, int n_beams, float angle_min, float angle_increment,
                       float range_min, float range_max, float scale_to_map, float* pts_out) {
  float angle = angle_min;
  float max_range_for_container = range_max - 0.1f;
  int n = 0;
  for (int i = 0; i < n_beams; ++i) {
    float dist = ranges[i];
    if ((dist > range_min) && (dist < max_range_for_container)) {
      dist *= scale_to_map;
      pts_out[2 * n] = cosf(angle) * dist;
      pts_out[2 * n + 1] = sinf(angle) * dist;
      ++n;
    }
    angle += angle_increment;
  }
  return n;
}

/* ------------------------------------------------------------- stream runner ---- */

/* HectorMappingRos::rosPointCloudToDataContainer (src/hector_slam/hector_mapping/src/HectorMappingRos.cpp:509-542).
 * tf is a ROS dependency absent from /root/reference and from this image (PARITY UNPINNED for its arithmetic): restated
 * from tf/LinearMath/Transform.h / Vector3.h / Matrix3x3.h (tfScalar = double): Transform * v = Vector3(basis[0].dot(v) +
 * origin.x, basis[1].dot(v) + origin.y, basis[2].dot(v) + origin.z), dot = x*x' + y*y' + z*z' evaluated left to right.
 * transform12 = row-major basis, then origin. Returns the number of points written (x,y interleaved), origo_out = laser
 * position * scaleToMap. */
int hso_cloud_to_points(const float* cloud_xyz, int n_cloud, const double* transform12, float scale_to_map,
                        float sqr_min_dist, float sqr_max_dist, float z_min, float z_max, float* pts_out, float* origo_out) {
  const double* b = transform12;
  const double* o = transform12 + 9;
  origo_out[0] = (float)o[0] * scale_to_map;
  origo_out[1] = (float)o[1] * scale_to_map;
  int count = 0;
  for (int i = 0; i < n_cloud; ++i) {
    float cx = cloud_xyz[3 * i], cy = cloud_xyz[3 * i + 1], cz = cloud_xyz[3 * i + 2];
    float dist_sqr = cx * cx + cy * cy;
    if ((dist_sqr > sqr_min_dist) && (dist_sqr < sqr_max_dist)) {
      if ((cx < 0.0f) && (dist_sqr < 0.50f)) continue;
      double vx = cx, vy = cy, vz = cz;
      double px = (b[0] * vx + b[1] * vy + b[2] * vz) + o[0];
      double py = (b[3] * vx + b[4] * vy + b[5] * vz) + o[1];
      double pz = (b[6] * vx + b[7] * vy + b[8] * vz) + o[2];
      float z_laser = (float)(pz - o[2]);
      if (z_laser > z_min && z_laser < z_max) {
        pts_out[2 * count] = (float)px * scale_to_map;
        pts_out[2 * count + 1] = (float)py * scale_to_map;
        ++count;
      }
    }
  }
  return count;
}


typedef struct {
  hso_proc** procs; int n_streams, n_scans, max_pts; const float* pts; const int* counts;
  int threads, tid, stream_major; float* poses_out; float* cov_out;
} run_arg;

static void* run_worker(void* vp) {
  run_arg* a = (run_arg*)vp;
  int s0 = (int)((long long)a->n_streams * a->tid / a->threads), s1 = (int)((long long)a->n_streams * (a->tid + 1) / a->threads);
  long long total = (long long)a->n_scans * (s1 - s0);
  for (long long it = 0; it < total; ++it) {
    int k = a->stream_major ? (int)(it % a->n_scans) : (int)(it / (s1 - s0));
    int s = s0 + (a->stream_major ? (int)(it / a->n_scans) : (int)(it % (s1 - s0)));
    hso_proc* p = a->procs[s];
    size_t slot = (size_t)k * a->n_streams + s;
    float hint[3] = { p->last_scan_match_pose[0], p->last_scan_match_pose[1], p->last_scan_match_pose[2] };
    hso_update(p, a->pts + slot * a->max_pts * 2, a->counts[slot], 0.0f, 0.0f, hint, 0);
    if (a->poses_out) memcpy(a->poses_out + slot * 3, p->last_scan_match_pose, sizeof(float) * 3);
    if (a->cov_out) memcpy(a->cov_out + slot * 9, p->last_cov, sizeof(float) * 9);
  }
  return NULL;
}

double hso_run_streams(hso_proc** procs, int n_streams, int n_scans, int max_pts, const float* pts, const int* counts,
                       int threads, int stream_major, float* poses_out, float* cov_out) {
  if (threads < 1) threads = 1;
  if (threads > 1024) threads = 1024;
  run_arg* args = (run_arg*)malloc(sizeof(run_arg) * threads);
  pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * threads);
  struct timespec t0, t1;
  clock_gettime(CLOCK_MONOTONIC, &t0);
  for (int t = 0; t < threads; ++t) {
    run_arg a = { procs, n_streams, n_scans, max_pts, pts, counts, threads, t, stream_major, poses_out, cov_out };
    args[t] = a;
    if (t > 0) pthread_create(&th[t], NULL, run_worker, &args[t]);
  }
  run_worker(&args[0]);
  for (int t = 1; t < threads; ++t) pthread_join(th[t], NULL);
  clock_gettime(CLOCK_MONOTONIC, &t1);
  free(args); free(th);
  return (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
}
