// sm_100a kernels of the hsgpu scan-to-map engine. Compiled with -fmad=false: every fp32 product/sum
// below rounds separately, in the reference's operation order (the CPU reference is built without FMA
// contraction), and fusion happens only inside hs_libm_exact.h where glibc fuses.
//
// HSL/ = src/hector_slam/hector_mapping/include/hector_slam_lib/ of the reference tree.
//
// Data layout in HBM (per engine):
//   val[stream][level-concatenated cells]  float   log-odds  (LogOddsCell::logOddsVal)
//   idx[stream][level-concatenated cells]  int32   update marker (LogOddsCell::updateIndex)
// i.e. the reference's AoS {float,int} cell split into two planes: the matcher gathers touch only
// `val` (8 cells per 32 B sector instead of 4), the ray-caster's "already marked this scan?" test touches
// only `idx`. Levels of one stream are contiguous; rows are row-major cell[y*W + x] as in
// GridMapBase::getCell (HSL/map/GridMapBase.h:141-149).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <float.h>

#include "hs_libm_exact.h"
#include "hsgpu.h"

struct HsLevel {
  int sx, sy;
  float scale;        // scaleToMap = 1/cellLength                (GridMapBase.h:270)
  float tmx, tmy;     // mapTworld.translation = scale * offset    (GridMapBase.h:272)
  float a;            // worldTmap.linear diagonal = scale * (1/(scale*scale))
  float twx, twy;     // worldTmap.translation                     (GridMapBase.h:279)
  float limx, limy;   // mapLimitsf = dims - 2                     (MapDimensionProperties.h:69-74)
  float pt_factor;    // (float)(1.0 / pow(2.0, level))            (MapRepMultiMap.h:127)
  int max_iter;       // 5 on level 0, 3 otherwise                 (MapRepMultiMap.h:125,128)
  unsigned long long cell_offset;  // first cell of this level inside a stream's slab
};

struct HsParams {
  int n_levels;
  int n_streams;
  int max_points;
  int strict;                       // 1: accumulate H/dTr in the reference's sequential point order
  unsigned long long cells_per_stream;
  HsLevel lv[HS_MAX_LEVELS];
  float log_odds_free, log_odds_occ;  // GridMapLogOddsFunctions (GridMapLogOdds.h:187-203), computed on the host
  float min_dist, min_angle;          // paramMinDistanceDiffForMapUpdate / AngleDiff (HectorSlamProcessor.h:138-139)
  // grids
  float* val;
  int* idx;
  // per-stream state
  float* last_pose;          // [n_streams][3] lastScanMatchPose
  float* last_update_pose;   // [n_streams][3] lastMapUpdatePose
  float* cov;                // [n_streams][9] lastScanMatchCov (row-major)
  int* cur_update_index;     // [n_streams]    OccGridMapBase::currUpdateIndex (same for all levels of a stream)
  int* last_update_index;    // [n_streams]    GridMapBase::lastUpdateIndex
  float* coarse_pts;         // [n_streams][max_points][2] level-0 points behind MapRepMultiMap::dataContainers
  int* coarse_cnt;           // [n_streams]
  float* coarse_origo;       // [n_streams][2]
  // per-slot (one slot = one stream update in the current batch)
  float* slot_pose;          // [n_streams][3] pose the ray-caster integrates with
  int* slot_integrate;       // [n_streams]    gate result
  int* slot_mark_base;       // [n_streams]    currUpdateIndex at the start of this scan's integration
  unsigned long long* counters;  // [0] unused [1] integrations [2] angle clamps [3] ray-cast cell visits
};

// ------------------------------------------------------------------------------------------------
// small helpers

__device__ __forceinline__ int hs_slot_stream(const int32_t* __restrict__ ids, int slot) { return ids ? ids[slot] : slot; }

// GridMapBase::getMapCoordsPose (GridMapBase.h:235-239): mapTworld * v, linear = diag(s, s).
__device__ __forceinline__ void hs_world_to_map(const HsLevel& L, float wx, float wy, float& mx, float& my) {
  mx = L.scale * wx + L.tmx;
  my = L.scale * wy + L.tmy;
}
// GridMapBase::getWorldCoordsPose (GridMapBase.h:226-230): worldTmap * v.
__device__ __forceinline__ void hs_map_to_world(const HsLevel& L, float mx, float my, float& wx, float& wy) {
  wx = L.a * mx + L.twx;
  wy = L.a * my + L.twy;
}

// GridMapLogOddsFunctions::getGridProbability (GridMapLogOdds.h:163-166): odds/(odds+1), odds = expf(v).
// The reference memoises this per cell (GridMapCacheArray.h); the cache is semantically invisible and is
// not reproduced: the value is recomputed from the log-odds on every gather.
__device__ __forceinline__ float hs_probability(float v) {
  float odds = hs_expf(v);
  return odds / (odds + 1.0f);
}

// util::normalize_angle (HSL/util/UtilFunctions.h:37-49), double fmod exactly as the reference.
__device__ __forceinline__ float hs_normalize_angle(float angle) {
  const double two_pi = 2.0f * 3.14159265358979323846;
  float a = (float)fmod(fmod((double)angle, two_pi) + two_pi, two_pi);
  if ((double)a > 3.14159265358979323846) a = (float)((double)a - two_pi);
  return a;
}

// util::poseDifferenceLargerThan (UtilFunctions.h:73-92)
__device__ __forceinline__ bool hs_pose_difference_larger_than(float x1, float y1, float t1, float x2, float y2, float t2,
                                                               float dist_thresh, float ang_thresh) {
  float dx = x1 - x2, dy = y1 - y2;
  if (sqrtf(dx * dx + dy * dy) > dist_thresh) return true;
  float angle_diff = t1 - t2;
  const double pi = 3.14159265358979323846;
  if ((double)angle_diff > pi) angle_diff = (float)((double)angle_diff - pi * 2.0f);
  else if ((double)angle_diff < -pi) angle_diff = (float)((double)angle_diff + pi * 2.0f);
  return fabsf(angle_diff) > ang_thresh;
}

// ------------------------------------------------------------------------------------------------
// k_scan_to_points: HectorMappingRos::rosLaserScanToDataContainer (HectorMappingRos.cpp:483-507).
// One warp per scan; beam_cs holds (cosf(angle_i), sinf(angle_i)) with angle_i the reference's float
// accumulation, precomputed on the host when the scan geometry is set. Valid beams are compacted in
// order (ballot + popc), as the reference's push_back loop does.
__global__ void __launch_bounds__(128) k_scan_to_points(const float* __restrict__ ranges, int n_slots, int n_beams,
                                                        const float2* __restrict__ beam_cs, float range_min,
                                                        float max_range_for_container, float scale_to_map, int max_points,
                                                        float2* __restrict__ points, int* __restrict__ counts) {
  int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  if (warp >= n_slots) return;
  const float* r = ranges + (size_t)warp * n_beams;
  float2* out = points + (size_t)warp * max_points;
  int base = 0;
  for (int i0 = 0; i0 < n_beams; i0 += 32) {
    int i = i0 + lane;
    float dist = (i < n_beams) ? r[i] : 0.0f;
    bool valid = (i < n_beams) && (dist > range_min) && (dist < max_range_for_container);
    unsigned m = __ballot_sync(0xffffffffu, valid);
    if (valid) {
      float d = dist * scale_to_map;
      float2 cs = beam_cs[i];
      int pos = base + __popc(m & ((1u << lane) - 1u));
      if (pos < max_points) out[pos] = make_float2(cs.x * d, cs.y * d);
    }
    base += __popc(m);
  }
  if (lane == 0) counts[warp] = base < max_points ? base : max_points;
}

// ------------------------------------------------------------------------------------------------
// Matching: OccGridMapUtil::getCompleteHessianDerivs + interpMapValueWithDerivatives
// (HSL/map/OccGridMapUtil.h:64-104, :287-347).

struct HsAcc {  // dTr[3] and the upper triangle of H
  float d0, d1, d2, h00, h11, h22, h01, h02, h12;
};

// Per-point terms: (gx, gy, rot, funVal). Out-of-bounds points contribute zeros like the reference.
__device__ __forceinline__ void hs_point_terms(const float* __restrict__ val, const HsLevel& L, float ex, float ey,
                                               float sn, float cs, float px, float py,
                                               float& gx, float& gy, float& rot, float& fun_val) {
  float qx = (cs * px + (-sn) * py) + ex;  // Affine2f * v = linear*v + translation
  float qy = (sn * px + cs * py) + ey;
  float m = 0.0f;
  gx = 0.0f;
  gy = 0.0f;
  if (qx >= 0.0f && qx <= L.limx && qy >= 0.0f && qy <= L.limy) {  // !pointOutOfMapBounds
    int ix = (int)qx, iy = (int)qy;
    float fx = qx - (float)ix, fy = qy - (float)iy;
    const float* p = val + (size_t)iy * L.sx + ix;
    float v0 = __ldg(p), v1 = __ldg(p + 1), v2 = __ldg(p + L.sx), v3 = __ldg(p + L.sx + 1);
    float i0 = hs_probability(v0), i1 = hs_probability(v1), i2 = hs_probability(v2), i3 = hs_probability(v3);
    float dx1 = i0 - i1, dx2 = i2 - i3;
    float dy1 = i0 - i2, dy2 = i1 - i3;
    float x_inv = 1.0f - fx, y_inv = 1.0f - fy;
    m = ((i0 * x_inv + i1 * fx) * y_inv) + ((i2 * x_inv + i3 * fx) * fy);
    gx = -((dx1 * x_inv) + (dx2 * fx));  // the reference blends with the x factors here (OccGridMapUtil.h:344)
    gy = -((dy1 * y_inv) + (dy2 * fy));  // ... and with the y factors here (:345)
  }
  fun_val = 1.0f - m;
  rot = ((-sn * px - cs * py) * gx + (cs * px - sn * py) * gy);
}

__device__ __forceinline__ void hs_acc_add(HsAcc& a, float gx, float gy, float rot, float fv) {
  a.d0 += gx * fv;
  a.d1 += gy * fv;
  a.d2 += rot * fv;
  a.h00 += gx * gx;
  a.h11 += gy * gy;
  a.h22 += rot * rot;
  a.h01 += gx * gy;
  a.h02 += gx * rot;
  a.h12 += gy * rot;
}

// One evaluation for one warp: all lanes return identical sums.
//   STRICT = false: each lane accumulates its points (i = lane, lane+32, ...), then a xor-butterfly.
//   STRICT = true : terms are computed in parallel but added in the reference's order i = 0..n-1
//                   (every lane walks the 32 shuffled terms of each round), giving bitwise the CPU sums.
template <bool STRICT>
__device__ __forceinline__ HsAcc hs_eval_warp(const float* __restrict__ val, const HsLevel& L, float ex, float ey, float eth,
                                              const float2* __restrict__ pts, int n, float pt_factor, int lane) {
  float sn = hs_sinf(eth), cs = hs_cosf(eth);
  HsAcc a = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (!STRICT) {
#pragma unroll 2
    for (int i = lane; i < n; i += 32) {
      float2 p = pts[i];
      float gx, gy, rot, fv;
      hs_point_terms(val, L, ex, ey, sn, cs, p.x * pt_factor, p.y * pt_factor, gx, gy, rot, fv);
      hs_acc_add(a, gx, gy, rot, fv);
    }
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) {
      a.d0 += __shfl_xor_sync(0xffffffffu, a.d0, m);
      a.d1 += __shfl_xor_sync(0xffffffffu, a.d1, m);
      a.d2 += __shfl_xor_sync(0xffffffffu, a.d2, m);
      a.h00 += __shfl_xor_sync(0xffffffffu, a.h00, m);
      a.h11 += __shfl_xor_sync(0xffffffffu, a.h11, m);
      a.h22 += __shfl_xor_sync(0xffffffffu, a.h22, m);
      a.h01 += __shfl_xor_sync(0xffffffffu, a.h01, m);
      a.h02 += __shfl_xor_sync(0xffffffffu, a.h02, m);
      a.h12 += __shfl_xor_sync(0xffffffffu, a.h12, m);
    }
  } else {
    for (int i0 = 0; i0 < n; i0 += 32) {
      int i = i0 + lane;
      float gx = 0.f, gy = 0.f, rot = 0.f, fv = 0.f;
      if (i < n) {
        float2 p = pts[i];
        hs_point_terms(val, L, ex, ey, sn, cs, p.x * pt_factor, p.y * pt_factor, gx, gy, rot, fv);
      }
      int cnt = min(32, n - i0);
      for (int l = 0; l < cnt; ++l) {
        float g0 = __shfl_sync(0xffffffffu, gx, l), g1 = __shfl_sync(0xffffffffu, gy, l);
        float r0 = __shfl_sync(0xffffffffu, rot, l), f0 = __shfl_sync(0xffffffffu, fv, l);
        hs_acc_add(a, g0, g1, r0, f0);
      }
    }
  }
  return a;
}

// ScanMatcher::estimateTransformationLogLh (HSL/matcher/ScanMatcher.h:194-221) with Eigen 3.3's
// Matrix3f::inverse (cofactors of column 0 -> det -> 1/det, entries = cofactor*invdet) and the
// 3-term reductions a0 + (a1 + a2). Returns true when the rotation step was clamped.
__device__ __forceinline__ bool hs_gauss_newton_step(const HsAcc& a, float& ex, float& ey, float& eth) {
  if (!((a.h00 != 0.0f) && (a.h11 != 0.0f))) return false;
  // symmetric H: m01 = m10 = h01, m02 = m20 = h02, m12 = m21 = h12
  const float m00 = a.h00, m01 = a.h01, m02 = a.h02, m11 = a.h11, m12 = a.h12, m22 = a.h22;
  const float m10 = m01, m20 = m02, m21 = m12;
  // cofactor(i,j) = m(i+1,j+1)*m(i+2,j+2) - m(i+1,j+2)*m(i+2,j+1), indices mod 3
  float c00 = m11 * m22 - m12 * m21;
  float c10 = m21 * m02 - m22 * m01;
  float c20 = m01 * m12 - m02 * m11;
  float det = c00 * m00 + (c10 * m10 + c20 * m20);
  float invdet = 1.0f / det;
  float c01 = m12 * m20 - m10 * m22;
  float c11 = m22 * m00 - m20 * m02;
  float c21 = m02 * m10 - m00 * m12;
  float c02 = m10 * m21 - m11 * m20;
  float c12 = m20 * m01 - m21 * m00;
  float c22 = m00 * m11 - m01 * m10;
  float r00 = c00 * invdet, r01 = c10 * invdet, r02 = c20 * invdet;
  float r10 = c01 * invdet, r11 = c11 * invdet, r12 = c21 * invdet;
  float r20 = c02 * invdet, r21 = c12 * invdet, r22 = c22 * invdet;
  float sx = r00 * a.d0 + (r01 * a.d1 + r02 * a.d2);
  float sy = r10 * a.d0 + (r11 * a.d1 + r12 * a.d2);
  float st = r20 * a.d0 + (r21 * a.d1 + r22 * a.d2);
  bool clamped = false;
  if (st > 0.2f) { st = 0.2f; clamped = true; }
  else if (st < -0.2f) { st = -0.2f; clamped = true; }
  ex += sx;
  ey += sy;
  eth += st;
  return clamped;
}

// k_match: HectorSlamProcessor::update minus the ray-cast (HSL/slam_main/HectorSlamProcessor.h:71-95):
//   MapRepMultiMap::matchData (MapRepMultiMap.h:116-132) -> ScanMatcher::matchData per level, coarse to fine
//   (ScanMatcher.h:54-190), then the poseDifferenceLargerThan gate and the per-stream bookkeeping.
// One warp owns one stream for the whole coarse-to-fine solve; nothing returns to the host in between.
// mode: 0 = update (match + gate), 1 = match only (no gate, no integration).
template <bool STRICT>
__global__ void __launch_bounds__(128) k_match(HsParams P, int n_slots, const int32_t* __restrict__ stream_ids,
                                               const float2* __restrict__ points, const int32_t* __restrict__ counts,
                                               const float2* __restrict__ origos, const float* __restrict__ hints,
                                               const uint8_t* __restrict__ mwm_flags, int mode) {
  int slot = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  if (slot >= n_slots) return;
  int s = hs_slot_stream(stream_ids, slot);
  int n = counts[slot];
  n = n < 0 ? 0 : (n > P.max_points ? P.max_points : n);
  const float2* pts = points + (size_t)slot * P.max_points;
  bool mwm = mwm_flags ? (mwm_flags[slot] != 0) : false;
  float wx, wy, wth;
  if (hints) { wx = hints[3 * slot]; wy = hints[3 * slot + 1]; wth = hints[3 * slot + 2]; }
  else { wx = P.last_pose[3 * s]; wy = P.last_pose[3 * s + 1]; wth = P.last_pose[3 * s + 2]; }

  if (!mwm) {
    // dataContainers[level-1].setFrom(dataContainer, 2^-level) (MapRepMultiMap.h:127): keep the level-0
    // points of the last matched scan; the power-of-two factor is applied (exactly) where they are used.
    float2* cp = reinterpret_cast<float2*>(P.coarse_pts) + (size_t)s * P.max_points;
    for (int i = lane; i < n; i += 32) cp[i] = pts[i];
    if (lane == 0) {
      P.coarse_cnt[s] = n;
      float2 o = origos ? origos[slot] : make_float2(0.f, 0.f);
      P.coarse_origo[2 * s] = o.x;
      P.coarse_origo[2 * s + 1] = o.y;
    }
    if (n != 0) {
      const float* val_s = P.val + (size_t)s * P.cells_per_stream;
      HsAcc a = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      unsigned clamps = 0;
      for (int level = P.n_levels - 1; level >= 0; --level) {
        const HsLevel& L = P.lv[level];
        const float* val = val_s + L.cell_offset;
        float ex, ey, eth = wth;
        hs_world_to_map(L, wx, wy, ex, ey);
        for (int it = 0; it <= L.max_iter; ++it) {  // 1 + maxIterations evaluations, no early exit
          a = hs_eval_warp<STRICT>(val, L, ex, ey, eth, pts, n, L.pt_factor, lane);
          clamps += hs_gauss_newton_step(a, ex, ey, eth) ? 1u : 0u;
        }
        eth = hs_normalize_angle(eth);
        hs_map_to_world(L, ex, ey, wx, wy);
        wth = eth;
      }
      if (lane == 0) {
        float* c = P.cov + 9 * (size_t)s;  // covMatrix = H of the last evaluation (ScanMatcher.h:184)
        c[0] = a.h00; c[1] = a.h01; c[2] = a.h02;
        c[3] = a.h01; c[4] = a.h11; c[5] = a.h12;
        c[6] = a.h02; c[7] = a.h12; c[8] = a.h22;
        if (clamps) atomicAdd(&P.counters[2], (unsigned long long)clamps);
      }
    }
  }
  if (lane == 0) {
    P.last_pose[3 * s] = wx; P.last_pose[3 * s + 1] = wy; P.last_pose[3 * s + 2] = wth;
    int integrate = 0;
    if (mode == 0) {
      float ux = P.last_update_pose[3 * s], uy = P.last_update_pose[3 * s + 1], uth = P.last_update_pose[3 * s + 2];
      integrate = (hs_pose_difference_larger_than(wx, wy, wth, ux, uy, uth, P.min_dist, P.min_angle) || mwm) ? 1 : 0;
      if (integrate) {
        P.last_update_pose[3 * s] = wx; P.last_update_pose[3 * s + 1] = wy; P.last_update_pose[3 * s + 2] = wth;
        int cur = P.cur_update_index[s];
        P.slot_mark_base[slot] = cur;
        P.cur_update_index[s] = cur + 3;        // OccGridMapBase.h:167
        P.last_update_index[s] += 1;            // setUpdated() (OccGridMapBase.h:164)
        P.slot_pose[3 * slot] = wx; P.slot_pose[3 * slot + 1] = wy; P.slot_pose[3 * slot + 2] = wth;
        atomicAdd(&P.counters[1], 1ULL);
      }
    }
    P.slot_integrate[slot] = integrate;
  }
}

// k_force_integrate: teacher-forced integration set-up (hs_raycast_batch): refresh the coarse containers from
// the given points, take the pose from the caller and open an update epoch, as updateByScan would.
__global__ void __launch_bounds__(128) k_force_integrate(HsParams P, int n_slots, const int32_t* __restrict__ stream_ids,
                                                         const float2* __restrict__ points, const int32_t* __restrict__ counts,
                                                         const float2* __restrict__ origos, const float* __restrict__ poses) {
  int slot = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  if (slot >= n_slots) return;
  int s = hs_slot_stream(stream_ids, slot);
  int n = counts[slot];
  n = n < 0 ? 0 : (n > P.max_points ? P.max_points : n);
  const float2* pts = points + (size_t)slot * P.max_points;
  float2* cp = reinterpret_cast<float2*>(P.coarse_pts) + (size_t)s * P.max_points;
  for (int i = lane; i < n; i += 32) cp[i] = pts[i];
  if (lane == 0) {
    P.coarse_cnt[s] = n;
    float2 o = origos ? origos[slot] : make_float2(0.f, 0.f);
    P.coarse_origo[2 * s] = o.x;
    P.coarse_origo[2 * s + 1] = o.y;
    int cur = P.cur_update_index[s];
    P.slot_mark_base[slot] = cur;
    P.cur_update_index[s] = cur + 3;
    P.last_update_index[s] += 1;
    P.slot_pose[3 * slot] = poses[3 * slot]; P.slot_pose[3 * slot + 1] = poses[3 * slot + 1]; P.slot_pose[3 * slot + 2] = poses[3 * slot + 2];
    P.slot_integrate[slot] = 1;
    atomicAdd(&P.counters[1], 1ULL);
  }
}

// k_hessian_derivs: getCompleteHessianDerivs for many pose hypotheses against one stream's level
// (BASELINE config 4). One warp per pose; pts are already scaled to the level.
template <bool STRICT>
__global__ void __launch_bounds__(128) k_hessian_derivs(HsParams P, int stream, int level, const float* __restrict__ poses_map,
                                                        int n_poses, const float2* __restrict__ pts, int n_pts,
                                                        float* __restrict__ H9, float* __restrict__ dTr3) {
  int k = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  if (k >= n_poses) return;
  const HsLevel& L = P.lv[level];
  const float* val = P.val + (size_t)stream * P.cells_per_stream + L.cell_offset;
  HsAcc a = hs_eval_warp<STRICT>(val, L, poses_map[3 * k], poses_map[3 * k + 1], poses_map[3 * k + 2], pts, n_pts, 1.0f, lane);
  if (lane == 0) {
    float* h = H9 + 9 * (size_t)k;
    h[0] = a.h00; h[1] = a.h01; h[2] = a.h02;
    h[3] = a.h01; h[4] = a.h11; h[5] = a.h12;
    h[6] = a.h02; h[7] = a.h12; h[8] = a.h22;
    float* d = dTr3 + 3 * (size_t)k;
    d[0] = a.d0; d[1] = a.d1; d[2] = a.d2;
  }
}

// ------------------------------------------------------------------------------------------------
// Integration: OccGridMapBase::updateByScan / updateLineBresenhami / bresenham2D / bresenhamCellFree /
// bresenhamCellOcc (HSL/map/OccGridMapBase.h:121-260) + the log-odds updates (GridMapLogOdds.h:135-156).
//
// The reference walks beams sequentially; its result per cell is a function of two facts only:
//   io = smallest beam index whose END POINT is the cell, jf = smallest beam index whose FREE trace crosses it.
//   io exists: idx = markOcc, v' = (jf < io ? (v + f) - f : v), then v' += o if v' < 50.
//   only jf:   idx = markFree, v += f.                      
// Every stored updateIndex is < markFree when a scan starts (currUpdateIndex grows by 3 per scan), so the
// marker word is free to serve as scratch during the scan as long as its final value is exact. One CTA
// owns one (stream, level) and runs three barrier-separated phases over all beams:
//   A  end points:  atomicMin(idx[end], TEMP | beam<<1)        -> io per end cell
//   B  free traces: cell still < markFree  -> atomicMax(idx, markFree); the unique winner adds f to val
//                   cell holds TEMP and this beam < io -> atomicOr(idx, 1)  (the "free came first" flag)
//   C  end points:  the beam that owns io applies undo/occupied to val and stores markOcc.
// A beam's own trace never contains its own end point (bresenham2D visits abs_da cells from the start).

#define HS_TEMP_BASE ((int)0x80000000)  // TEMP values are < -1, every legitimate marker is >= -1

struct HsBeam {
  int valid;            // begin != end and both inside the grid
  int k0;               // start cell offset
  int kend;             // end cell offset
  unsigned abs_da, abs_db;
  int off_a, off_b;
};

__device__ __forceinline__ HsBeam hs_make_beam(const HsLevel& L, int bx, int by, float mpx, float mpy, float sn, float cs,
                                              float px, float py) {
  HsBeam b;
  float exf = ((cs * px + (-sn) * py) + mpx) + 0.5f;
  float eyf = ((sn * px + cs * py) + mpy) + 0.5f;
  int ex = (int)exf, ey = (int)eyf;  // truncation toward zero, like cast<int>() (OccGridMapBase.h:152-155)
  b.valid = (bx != ex || by != ey) && bx >= 0 && bx < L.sx && by >= 0 && by < L.sy && ex >= 0 && ex < L.sx && ey >= 0 && ey < L.sy;
  int dx = ex - bx, dy = ey - by;
  unsigned abs_dx = (unsigned)abs(dx), abs_dy = (unsigned)abs(dy);
  int off_dx = dx > 0 ? 1 : -1;             // util::sign: sign(0) = -1 (UtilFunctions.h:56-59)
  int off_dy = (dy > 0 ? 1 : -1) * L.sx;
  b.k0 = by * L.sx + bx;
  b.kend = ey * L.sx + ex;
  if (abs_dx >= abs_dy) { b.abs_da = abs_dx; b.abs_db = abs_dy; b.off_a = off_dx; b.off_b = off_dy; }
  else { b.abs_da = abs_dy; b.abs_db = abs_dx; b.off_a = off_dy; b.off_b = off_dx; }
  return b;
}

__device__ __forceinline__ void hs_mark_free(float* __restrict__ val, int* __restrict__ idx, int k, int beam, int mark_free, float f) {
  int w = __ldcg(idx + k);
  if (w < -1) {                                   // TEMP: an end point lives here
    int io = (w - HS_TEMP_BASE) >> 1;
    if (beam < io && !(w & 1)) atomicOr(idx + k, 1);
  } else if (w < mark_free) {
    int old = atomicMax(idx + k, mark_free);
    if (old < mark_free) val[k] = val[k] + f;     // updateSetFree, exactly once per cell and scan
  }
}

__global__ void __launch_bounds__(256) k_raycast(HsParams P, int n_slots, const int32_t* __restrict__ stream_ids,
                                                 const float2* __restrict__ points, const int32_t* __restrict__ counts,
                                                 const float2* __restrict__ origos) {
  int slot = blockIdx.x / P.n_levels;
  int level = blockIdx.x - slot * P.n_levels;
  if (slot >= n_slots) return;
  if (!P.slot_integrate[slot]) return;
  int s = hs_slot_stream(stream_ids, slot);
  const HsLevel& L = P.lv[level];
  float* val = P.val + (size_t)s * P.cells_per_stream + L.cell_offset;
  int* idx = P.idx + (size_t)s * P.cells_per_stream + L.cell_offset;

  // level 0 integrates the scan itself, level i > 0 the container last refreshed by matchData
  // (MapRepMultiMap.h:134-147) -- possibly stale or empty when map_without_matching was used.
  const float2* pts;
  int n;
  float ox, oy;
  if (level == 0) {
    pts = points + (size_t)slot * P.max_points;
    n = counts[slot];
    n = n < 0 ? 0 : (n > P.max_points ? P.max_points : n);
    float2 o = origos ? origos[slot] : make_float2(0.f, 0.f);
    ox = o.x; oy = o.y;
  } else {
    pts = reinterpret_cast<const float2*>(P.coarse_pts) + (size_t)s * P.max_points;
    n = P.coarse_cnt[s];
    ox = P.coarse_origo[2 * s] * L.pt_factor;
    oy = P.coarse_origo[2 * s + 1] * L.pt_factor;
  }
  const int mark_base = P.slot_mark_base[slot];
  const int mark_free = mark_base + 1, mark_occ = mark_base + 2;
  const float f = P.log_odds_free, o = P.log_odds_occ;

  float mpx, mpy;
  hs_world_to_map(L, P.slot_pose[3 * slot], P.slot_pose[3 * slot + 1], mpx, mpy);
  float th = P.slot_pose[3 * slot + 2];
  float sn = hs_sinf(th), cs = hs_cosf(th);
  int bx = (int)(((cs * ox + (-sn) * oy) + mpx) + 0.5f);  // Vector2i(float, float): truncation (OccGridMapBase.h:137)
  int by = (int)(((sn * ox + cs * oy) + mpy) + 0.5f);
  const float pf = L.pt_factor;

  // phase A: end points
  unsigned visits = 0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    float2 p = pts[i];
    HsBeam b = hs_make_beam(L, bx, by, mpx, mpy, sn, cs, p.x * pf, p.y * pf);
    if (b.valid) {
      atomicMin(idx + b.kend, HS_TEMP_BASE + (i << 1));
      visits += b.abs_da + 1u;
    }
  }
  visits = __reduce_add_sync(0xffffffffu, visits);
  if ((threadIdx.x & 31) == 0 && visits) atomicAdd(&P.counters[3], (unsigned long long)visits);
  __syncthreads();
  // phase B: free traces (bresenham2D, OccGridMapBase.h:243-260)
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    float2 p = pts[i];
    HsBeam b = hs_make_beam(L, bx, by, mpx, mpy, sn, cs, p.x * pf, p.y * pf);
    if (!b.valid) continue;
    int k = b.k0;
    int err = (int)(b.abs_da / 2);
    hs_mark_free(val, idx, k, i, mark_free, f);
    unsigned end = b.abs_da - 1;
    for (unsigned t = 0; t < end; ++t) {
      k += b.off_a;
      err += (int)b.abs_db;
      if ((unsigned)err >= b.abs_da) { k += b.off_b; err -= (int)b.abs_da; }
      hs_mark_free(val, idx, k, i, mark_free, f);
    }
  }
  __syncthreads();
  // phase C: end points
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    float2 p = pts[i];
    HsBeam b = hs_make_beam(L, bx, by, mpx, mpy, sn, cs, p.x * pf, p.y * pf);
    if (!b.valid) continue;
    int w = __ldcg(idx + b.kend);
    if (w < -1 && ((w - HS_TEMP_BASE) >> 1) == i) {
      float v = val[b.kend];
      if (w & 1) v = (v + f) - f;        // updateSetFree by an earlier beam, then updateUnsetFree
      if (v < 50.0f) v = v + o;          // updateSetOccupied
      val[b.kend] = v;
      idx[b.kend] = mark_occ;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// maintenance kernels

// GridMapBase::clear -> LogOddsCell::resetGridCell (GridMapBase.h:69-87, GridMapLogOdds.h:89-93)
__global__ void k_clear_cells(float* __restrict__ val, int* __restrict__ idx, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) { val[i] = 0.0f; idx[i] = -1; }
}

// HectorSlamProcessor::reset state (HectorSlamProcessor.h:115-124); update counters keep counting.
__global__ void k_reset_state(HsParams P, int first, int n, int init_counters) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int s = first + i;
  for (int k = 0; k < 3; ++k) { P.last_update_pose[3 * s + k] = FLT_MAX; P.last_pose[3 * s + k] = 0.0f; }
  if (init_counters) {
    for (int k = 0; k < 9; ++k) P.cov[9 * s + k] = 0.0f;
    P.cur_update_index[s] = 0;
    P.last_update_index[s] = -1;
    P.coarse_cnt[s] = 0;
    P.coarse_origo[2 * s] = 0.0f; P.coarse_origo[2 * s + 1] = 0.0f;
  }
}

// interleave / de-interleave one level between the SoA planes and the reference's {float,int} cells
__global__ void k_pack_cells(const float* __restrict__ val, const int* __restrict__ idx, hs_cell* __restrict__ out, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) { hs_cell c; c.logOddsVal = val[i]; c.updateIndex = idx[i]; out[i] = c; }
}
__global__ void k_unpack_cells(const hs_cell* __restrict__ in, float* __restrict__ val, int* __restrict__ idx, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) { hs_cell c = in[i]; val[i] = c.logOddsVal; idx[i] = c.updateIndex; }
}
// HectorMappingRos::publishMap classification (HectorMappingRos.cpp:447-470): isFree -> 0, isOccupied -> 100, else -1
__global__ void k_classify(const float* __restrict__ val, int8_t* __restrict__ out, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) { float v = val[i]; out[i] = (v < 0.0f) ? (int8_t)0 : ((v > 0.0f) ? (int8_t)100 : (int8_t)-1); }
}
