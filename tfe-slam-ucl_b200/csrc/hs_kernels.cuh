// sm_100a kernels of the hsgpu scan-to-map engine. Compiled with -fmad=false: every fp32 product/sum
// below rounds separately, in the reference's operation order (the CPU reference is built without FMA
// contraction), and fusion happens only inside hs_libm_exact.h where glibc fuses.
//
// HSL/ = src/hector_slam/hector_mapping/include/hector_slam_lib/ of the reference tree.
//
// Data layout in HBM (per engine):
//   val[stream][level-concatenated cells]  float   log-odds  (LogOddsCell::logOddsVal)
//   idx[stream][level-concatenated cells]  int32   update marker (LogOddsCell::updateIndex), cold plane
//   mk [stream][level-concatenated cells]  uint8   update marker, hot plane: what the ray-caster writes. A scan marks ~26 k
//        cells, most of them in partly marked 32-byte sectors whose write makes HBM read the sector first; with one byte per
//        cell a sector holds 32 cells instead of 8. b != 0: updateIndex = mk_base[stream][level] + b (markers grow by 3 per
//        integrated scan, so a base lasts 84 scans); b == 0: the int32 plane holds the marker. When the next scan's markers
//        would not fit a byte, the (stream, level)'s ray-cast CTA first folds the byte plane into the int32 plane and
//        re-bases (hs_fold_markers). Readers (k_pack_cells) merge the two planes.
// i.e. the reference's AoS {float,int} cell split into two planes: the matcher gathers touch only
// `val` (8 cells per 32 B sector instead of 4), the ray-caster's "already marked this scan?" test touches
// only `idx`. Levels of one stream are contiguous; rows are row-major cell[y*W + x] as in
// GridMapBase::getCell (HSL/map/GridMapBase.h:141-149).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <float.h>
#include <limits.h>

#include <type_traits>

#include "hs_libm_exact.h"
#include "hsgpu.h"

// Timing-experiment switches (HSGPU_DEBUG_SKIP bits, HSGPU_DEBUG_TIMES stamps) exist only in builds made with
// `make EXPERIMENTS=1` (which also compiles the matcher's phase clocks in); the product build has none of these branches.
#ifdef HS_EXPERIMENTS
#define HS_DBG(P_, mask) ((P_).dbg_skip & (mask))
#else
#define HS_DBG(P_, mask) 0
#endif

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
  float2* stage_pts;                // matcher: non-NULL = `points` is pinned HOST memory read over PCIe; each CTA copies its scan here
                                    // (device, [slot][max_points]) in one pass and everything else reads the copy
  int32_t* stage_cnt;               // ... and its point count here ([slot]); NULL = `counts` is already device memory
  // multi-start matching (hs_match_hypotheses): non-NULL = every slot matches the SAME scan (slot 0's points) against the
  // maps of stream `hyp_stream` from its own hint, touches no stream state, and writes pose / covariance here ([slot][3], [slot][9])
  float* hyp_pose;
  float* hyp_cov;
  int hyp_stream;
  const float4* hyp_pblk;           // ... WC == 2 instantiation: the stream's probability block planes (k_build_pblk) instead of val
  float* host_pose_out;             // non-NULL: the caller's pinned poses_out [slot][3]; the matcher's tail writes the pose there too
  int n_sm;                         // SM count of the device
  int reverse_order;                // k_raycast: this launch walks the slots from the last to the first
  int l2_keep_levels;               // bit l: the matcher's gathers on level l carry an L2 evict-last hint
  int tail_slots;                   // k_raycast: the last `tail_slots` slots of a launch are dispatched level-major
  int slot0;                        // first slot of this launch (a host batch may be cut into chunks that run on several CUDA streams)
  int dbg_skip;                     // timing experiments only (HSGPU_DEBUG_SKIP): bit0 skip walk, bit1 skip sweep, bit2 skip end points
  unsigned long long* dbg_times;    // timing experiments only (HSGPU_DEBUG_TIMES): 8 globaltimer stamps per ray-cast CTA
  int strict;                       // 1: accumulate H/dTr in the reference's sequential point order
  unsigned long long cells_per_stream;
  HsLevel lv[HS_MAX_LEVELS];
  float log_odds_free, log_odds_occ;  // GridMapLogOddsFunctions (GridMapLogOdds.h:187-203), computed on the host
  float min_dist, min_angle;          // paramMinDistanceDiffForMapUpdate / AngleDiff (HectorSlamProcessor.h:138-139)
  // grids
  float* val;
  int* idx;                  // update markers, int32 plane: authoritative where the byte plane holds 0
  uint8_t* mk;               // update markers, hot byte plane: b != 0 means updateIndex = mk_base[stream][level] + b
  int* mk_base;              // [n_streams][HS_MAX_LEVELS]
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
  float* slot_sc;            // [n_streams][2] sinf / cosf of that pose's angle (computed once, by the gate)
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
#define HS_SCAN_INFLIGHT 12   // 384 beams per pass (scalar path)
#define HS_SCAN_INFLIGHT4 3   // 3 x 128 beams per pass (float4 path)
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
  if ((n_beams & 3) == 0 && (reinterpret_cast<uintptr_t>(ranges) & 15) == 0) {
    // float4 path: a lane owns 4 consecutive beams (16-byte range loads, 32-byte cos/sin loads); the ordered compaction is a
    // warp prefix sum over the lanes' valid counts plus the position inside the lane
    const float4* r4 = reinterpret_cast<const float4*>(r);
    const float4* cs4 = reinterpret_cast<const float4*>(beam_cs);   // two (cos, sin) pairs per float4
    const int n4 = n_beams >> 2;
    for (int c0 = 0; c0 < n4; c0 += 32 * HS_SCAN_INFLIGHT4) {
      float4 d_in[HS_SCAN_INFLIGHT4];
#pragma unroll
      for (int u = 0; u < HS_SCAN_INFLIGHT4; ++u) {   // all loads of a chunk first: `ranges` may be pinned host memory (zero-copy)
        const int q = c0 + u * 32 + lane;
        d_in[u] = q < n4 ? r4[q] : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int u = 0; u < HS_SCAN_INFLIGHT4; ++u) {
        const int q = c0 + u * 32 + lane;
        const float d[4] = {d_in[u].x, d_in[u].y, d_in[u].z, d_in[u].w};
        unsigned vm = 0u;
#pragma unroll
        for (int t = 0; t < 4; ++t)
          if (q < n4 && d[t] > range_min && d[t] < max_range_for_container) vm |= 1u << t;
        const int mine = __popc(vm);
        int incl = mine;
#pragma unroll
        for (int s = 1; s < 32; s <<= 1) {
          const int t = __shfl_up_sync(0xffffffffu, incl, s);
          if (lane >= s) incl += t;
        }
        int pos = base + incl - mine;
        if (vm) {
          const float4 ca = cs4[2 * q], cb = cs4[2 * q + 1];
          const float cx[4] = {ca.x, ca.z, cb.x, cb.z}, sy[4] = {ca.y, ca.w, cb.y, cb.w};
#pragma unroll
          for (int t = 0; t < 4; ++t)
            if (vm & (1u << t)) {
              const float dd = d[t] * scale_to_map;
              if (pos < max_points) out[pos] = make_float2(cx[t] * dd, sy[t] * dd);
              ++pos;
            }
        }
        base += __shfl_sync(0xffffffffu, incl, 31);
      }
    }
    if (lane == 0) counts[warp] = base < max_points ? base : max_points;
    return;
  }
  for (int c0 = 0; c0 < n_beams; c0 += 32 * HS_SCAN_INFLIGHT) {
    // all loads of a chunk are issued before the first is used: `ranges` may be the caller's pinned host buffer read over
    // PCIe (zero-copy), where one round trip costs microseconds
    float d_in[HS_SCAN_INFLIGHT];
#pragma unroll
    for (int u = 0; u < HS_SCAN_INFLIGHT; ++u) {
      const int i = c0 + u * 32 + lane;
      d_in[u] = (i < n_beams) ? r[i] : 0.0f;
    }
#pragma unroll
    for (int u = 0; u < HS_SCAN_INFLIGHT; ++u) {
      const int i = c0 + u * 32 + lane;
      const float dist = d_in[u];
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
  }
  if (lane == 0) counts[warp] = base < max_points ? base : max_points;
}

// ------------------------------------------------------------------------------------------------
// k_cloud_to_points: HectorMappingRos::rosPointCloudToDataContainer (HectorMappingRos.cpp:509-542), the node's default
// ingestion path (use_tf_scan_transformation): laser-frame points (x, y, z) -> DataContainer points in the base frame,
// scaled to level-0 cells, with the reference's filters (squared planar distance window, rear points closer than
// sqrt(0.5) m dropped, z window in the laser frame) and origo = laser position * scaleToMap. tf (absent here) works in
// double: T * v = (basis.row(i) . v) + origin_i, products summed left to right (tf/LinearMath/Transform.h, Vector3.h).
// One warp per cloud; xf [n][12] = row-major 3x3 basis then origin. Ordered compaction like the reference's push_back.
struct HsCloudFilter { float sqr_min_dist, sqr_max_dist, z_min, z_max; };

__host__ __device__ __forceinline__ bool hs_cloud_point(const float* p3, const double* xf, const HsCloudFilter& F, float scale_to_map,
                                                        float& ox, float& oy) {
  const float x = p3[0], y = p3[1], z = p3[2];
  const float dist_sqr = x * x + y * y;
  if (!((dist_sqr > F.sqr_min_dist) && (dist_sqr < F.sqr_max_dist))) return false;
  if ((x < 0.0f) && (dist_sqr < 0.50f)) return false;
  const double vx = (double)x, vy = (double)y, vz = (double)z;
  const double bx = ((xf[0] * vx + xf[1] * vy) + xf[2] * vz) + xf[9];
  const double by = ((xf[3] * vx + xf[4] * vy) + xf[5] * vz) + xf[10];
  const double bz = ((xf[6] * vx + xf[7] * vy) + xf[8] * vz) + xf[11];
  const float z_laser = (float)(bz - xf[11]);
  if (!(z_laser > F.z_min && z_laser < F.z_max)) return false;
  ox = (float)bx * scale_to_map;
  oy = (float)by * scale_to_map;
  return true;
}

__global__ void __launch_bounds__(128) k_cloud_to_points(const float* __restrict__ clouds, const int32_t* __restrict__ cloud_counts, int n_slots,
                                                         int max_cloud, const double* __restrict__ xf, HsCloudFilter F, float scale_to_map,
                                                         int max_points, float2* __restrict__ points, int* __restrict__ counts,
                                                         float2* __restrict__ origos) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= n_slots) return;
  const float* c = clouds + (size_t)warp * max_cloud * 3;
  const double* t = xf + (size_t)warp * 12;
  float2* out = points + (size_t)warp * max_points;
  int n = cloud_counts[warp];
  n = n < 0 ? 0 : (n > max_cloud ? max_cloud : n);
  int base = 0;
  for (int i0 = 0; i0 < n; i0 += 32) {
    const int i = i0 + lane;
    float px = 0.0f, py = 0.0f;
    const bool valid = (i < n) && hs_cloud_point(c + 3 * (size_t)i, t, F, scale_to_map, px, py);
    const unsigned m = __ballot_sync(0xffffffffu, valid);
    if (valid) {
      const int pos = base + __popc(m & ((1u << lane) - 1u));
      if (pos < max_points) out[pos] = make_float2(px, py);
    }
    base += __popc(m);
  }
  if (lane == 0) {
    counts[warp] = base < max_points ? base : max_points;
    origos[warp] = make_float2((float)t[9] * scale_to_map, (float)t[10] * scale_to_map);   // Vector2f(laserPos.x(), laserPos.y()) * scaleToMap
  }
}

// ------------------------------------------------------------------------------------------------
// Matching: OccGridMapUtil::getCompleteHessianDerivs + interpMapValueWithDerivatives
// (HSL/map/OccGridMapUtil.h:64-104, :287-347).

struct HsAcc {  // dTr[3] and the upper triangle of H
  float d0, d1, d2, h00, h11, h22, h01, h02, h12;
};

// The four cell probabilities around a point, remembered between evaluations: Gauss-Newton steps on one level are
// mostly sub-cell, so the same 2x2 cell block is hit again and its log-odds -> probability conversions (4 double
// precision expf + 4 divisions) and gathers can be skipped. Exactly the same values either way.
struct HsCellCache {
  int key;          // cell offset of the block's first cell on the current level, -1 = empty
  float p0, p1, p2, p3;
};

// Per-point terms: (gx, gy, rot, funVal). Out-of-bounds points contribute zeros like the reference.
template <bool CACHED>
__device__ __forceinline__ void hs_point_terms(const float* __restrict__ val, const HsLevel& L, float ex, float ey,
                                               float sn, float cs, float px, float py, HsCellCache& cc,
                                               float& gx, float& gy, float& rot, float& fun_val) {
  float qx = (cs * px + (-sn) * py) + ex;  // Affine2f * v = linear*v + translation
  float qy = (sn * px + cs * py) + ey;
  float m = 0.0f;
  gx = 0.0f;
  gy = 0.0f;
  if (qx >= 0.0f && qx <= L.limx && qy >= 0.0f && qy <= L.limy) {  // !pointOutOfMapBounds
    int ix = (int)qx, iy = (int)qy;
    float fx = qx - (float)ix, fy = qy - (float)iy;
    const int k = iy * L.sx + ix;
    float i0, i1, i2, i3;
    if (CACHED && k == cc.key) {
      i0 = cc.p0; i1 = cc.p1; i2 = cc.p2; i3 = cc.p3;
    } else {
      const float* p = val + k;
      float v0 = __ldg(p), v1 = __ldg(p + 1), v2 = __ldg(p + L.sx), v3 = __ldg(p + L.sx + 1);
      i0 = hs_probability(v0); i1 = hs_probability(v1); i2 = hs_probability(v2); i3 = hs_probability(v3);
      if (CACHED) { cc.key = k; cc.p0 = i0; cc.p1 = i1; cc.p2 = i2; cc.p3 = i3; }
    }
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

// timing experiments only (build with `make EXPERIMENTS=1`, run with HSGPU_DEBUG_TIMES): thread 0 of a matcher CTA accumulates
// clock64() deltas per phase. Compiles to nothing in the product build.
struct HsMatchTimer {
#ifdef HS_EXPERIMENTS
  long long t, acc[5];
  unsigned long long* out;
  __device__ __forceinline__ void begin(unsigned long long* o) { out = o; if (out && threadIdx.x == 0) { t = clock64(); for (int k = 0; k < 5; ++k) acc[k] = 0; } }
  __device__ __forceinline__ void mark(int k) { if (out && threadIdx.x == 0) { long long c = clock64(); acc[k] += c - t; t = c; } }
  __device__ __forceinline__ void end() { if (out && threadIdx.x == 0) for (int k = 0; k < 5; ++k) out[(size_t)blockIdx.x * 8 + k] = (unsigned long long)acc[k]; }
#else
  __device__ __forceinline__ void begin(unsigned long long*) {}
  __device__ __forceinline__ void mark(int) {}
  __device__ __forceinline__ void end() {}
#endif
};

// One evaluation of getCompleteHessianDerivs by a whole CTA. The per-point work (transform, 4 gathers,
// probabilities, the 9 products) is spread over all threads; the products go to shared memory as
// prod[i][9] (stride 9 words: conflict-free for both the per-point writes and the per-sum reads).
// Warp 0 then forms the nine sums and returns them in all of its lanes:
//   STRICT = true  (default): lanes 0..8 each own one sum and add the n products in the reference's order
//                   i = 0..n-1 (OccGridMapUtil.h:76-98) -- one dependent FADD chain per sum, bitwise the
//                   CPU result. fp32 addition is not associative and hector's Gauss-Newton amplifies
//                   last-bit differences of H/dTr to >1e-4 m on some streams, so this is the mode that
//                   meets the pose-parity bar.
//   STRICT = false: every lane sums the points i = lane (mod 32) and a xor-butterfly combines them.
// Must be called by all threads of the CTA; contains two __syncthreads. Only the serial warp's return value is defined.
#define HS_NPROD 9

// Products are stored SUM-MAJOR: prod[k * pitch + i] = k-th product of point i, so that the lane that owns sum k reads
// four consecutive terms of its chain per 16-byte shared load; pitch == 4 (mod 32) keeps those reads conflict-free.
__host__ __device__ __forceinline__ int hs_match2_pitch(int max_points) {   // >= max_points, == 4 (mod 32)
  return ((max_points + 27) & ~31) + 4;
}

__device__ __forceinline__ void hs_store_products(float* __restrict__ q, int pitch, float gx, float gy, float rot, float fv) {
  q[0] = gx * fv;             // dTr[0]
  q[pitch] = gy * fv;         // dTr[1]
  q[2 * pitch] = rot * fv;    // dTr[2]
  q[3 * pitch] = gx * gx;     // H(0,0)
  q[4 * pitch] = gy * gy;     // H(1,1)
  q[5 * pitch] = rot * rot;   // H(2,2)
  q[6 * pitch] = gx * gy;     // H(0,1)
  q[7 * pitch] = gx * rot;    // H(0,2)
  q[8 * pitch] = gy * rot;    // H(1,2)
}

// Gather of a log-odds value with an L2 evict-last hint: the cells around the scan end points are read again by the next
// scan's match (the robot moves a few cells per scan), while the ray-caster streams half a gigabyte of other lines through
// L2 in between.
__device__ __forceinline__ float hs_ldg_evict_last(const float* p, unsigned long long policy) {
  float v;
  asm volatile("ld.global.nc.L2::cache_hint.f32 %0, [%1], %2;" : "=f"(v) : "l"(p), "l"(policy));
  return v;
}
__device__ __forceinline__ unsigned long long hs_policy_evict_normal() {
  unsigned long long policy;
  asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(policy));
  return policy;
}
__device__ __forceinline__ unsigned long long hs_policy_evict_last() {
  unsigned long long policy;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(policy));
  return policy;
}

// Points phase of one evaluation with the cache misses of a WARP processed densely. With the register cache only a few
// lanes of a warp miss in most evaluations (measured: 3.5 of 32 per point slot), yet every miss drags the whole warp through
// four inlined probability computations. Here the lanes first publish the block addresses of their misses to a per-warp list
// in shared memory; the list x 4 corners is then dealt out to all 32 lanes (one gather, one expf, one division per item, four
// consecutive lanes = the four corners of one block), and the owners read their four probabilities back as one float4.
// Same values, same order of the per-point arithmetic; only who computes a probability changes.
//   w_addr: this warp's PPT*32 ints, w_prob: this warp's PPT*128 floats (16-byte aligned), s_tab: expf table in shared memory
// PB: `val` is not the log-odds plane but the level's probability BLOCK plane (k_build_pblk; multi-start matching, where
// thousands of hypotheses read one stream's maps): a miss is one 16-byte load of the finished 2x2 block, no dense pass.
template <int PPT, bool PB = false>
__device__ __forceinline__ void hs_points_phase_wc(const float* __restrict__ val, const HsLevel& L, float ex, float ey, float sn, float cs,
                                                   const float2* preg, HsCellCache* cache, int n, float pt_factor,
                                                   float* __restrict__ prod, int pitch, int* __restrict__ w_addr,
                                                   float* __restrict__ w_prob, const uint64_t* __restrict__ s_tab, int i0, int istride,
                                                   bool keep_in_l2) {
  const int lane = threadIdx.x & 31;
  const unsigned lt = (1u << lane) - 1u;
  unsigned bal[PPT];
  unsigned mine = 0;   // bit j: slot j missed
  int total = 0;
#pragma unroll
  for (int j = 0; j < PPT; ++j) {
    const int i = i0 + j * istride;
    const float px = preg[j].x * pt_factor, py = preg[j].y * pt_factor;
    const float qx = (cs * px + (-sn) * py) + ex;  // Affine2f * v = linear*v + translation
    const float qy = (sn * px + cs * py) + ey;
    bool m = false;
    if (i < n && qx >= 0.0f && qx <= L.limx && qy >= 0.0f && qy <= L.limy) {  // !pointOutOfMapBounds
      const int k = (int)qy * L.sx + (int)qx;
      if (k != cache[j].key) { cache[j].key = k; m = true; }
    }
    if (PB) {
      if (m) {
        const float4 b = __ldg(reinterpret_cast<const float4*>(val) + cache[j].key);
        cache[j].p0 = b.x; cache[j].p1 = b.y; cache[j].p2 = b.z; cache[j].p3 = b.w;
      }
      continue;
    }
    bal[j] = __ballot_sync(0xffffffffu, m);
    if (m) {
      mine |= 1u << j;
      w_addr[total + __popc(bal[j] & lt)] = cache[j].key;
    }
    total += __popc(bal[j]);
  }
  if (!PB && total) {   // warp-uniform
    __syncwarp();
    // item = (miss, corner): lane l always handles corner l & 3 of miss (l >> 2) + 8 * round
    const int coff = (lane & 1) + ((lane & 2) ? L.sx : 0);
    const int m0 = lane >> 2;
    const unsigned long long pol = keep_in_l2 ? hs_policy_evict_last() : hs_policy_evict_normal();
    for (int mb = 0; mb < total; mb += 32) {   // 32 misses = 128 items per turn, four gathers in flight per lane
      float v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int mi = mb + u * 8 + m0;
        v[u] = 0.0f;
        if (mi < total) v[u] = hs_ldg_evict_last(val + w_addr[mi] + coff, pol);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        if (mb + u * 8 < total) {   // warp-uniform
          const float odds = hs_expf_tab(v[u], s_tab);
          const float pr = odds / (odds + 1.0f);
          if (mb + u * 8 + m0 < total) w_prob[((mb + u * 8) << 2) + lane] = pr;
        }
      }
    }
    __syncwarp();
    int off = 0;
#pragma unroll
    for (int j = 0; j < PPT; ++j) {
      if (mine & (1u << j)) {
        const float4 p4 = reinterpret_cast<const float4*>(w_prob)[off + __popc(bal[j] & lt)];
        cache[j].p0 = p4.x; cache[j].p1 = p4.y; cache[j].p2 = p4.z; cache[j].p3 = p4.w;
      }
      off += __popc(bal[j]);
    }
  }
#pragma unroll
  for (int j = 0; j < PPT; ++j) {
    const int i = i0 + j * istride;
    if (i < n) {
      const float px = preg[j].x * pt_factor, py = preg[j].y * pt_factor;
      const float qx = (cs * px + (-sn) * py) + ex;
      const float qy = (sn * px + cs * py) + ey;
      float m = 0.0f, gx = 0.0f, gy = 0.0f;
      if (qx >= 0.0f && qx <= L.limx && qy >= 0.0f && qy <= L.limy) {
        const int ix = (int)qx, iy = (int)qy;
        const float fx = qx - (float)ix, fy = qy - (float)iy;
        const float v0 = cache[j].p0, v1 = cache[j].p1, v2 = cache[j].p2, v3 = cache[j].p3;   // intensities[0..3] (:318-335)
        const float dx1 = v0 - v1, dx2 = v2 - v3;
        const float dy1 = v0 - v2, dy2 = v1 - v3;
        const float x_inv = 1.0f - fx, y_inv = 1.0f - fy;
        m = ((v0 * x_inv + v1 * fx) * y_inv) + ((v2 * x_inv + v3 * fx) * fy);
        gx = -((dx1 * x_inv) + (dx2 * fx));  // the reference blends with the x factors here (OccGridMapUtil.h:344)
        gy = -((dy1 * y_inv) + (dy2 * fy));  // ... and with the y factors here (:345)
      }
      const float fv = 1.0f - m;
      const float rot = ((-sn * px - cs * py) * gx + (cs * px - sn * py) * gy);
      hs_store_products(prod + i, pitch, gx, gy, rot, fv);
    }
  }
}

// PPT > 0: the thread's points (i = tid + j*blockDim, j < PPT) are held in registers by the caller (they are the
// same for all evaluations of a scan); PPT == 0: read them from `pts` (any n).
template <bool STRICT, int PPT, int WC = 0>
__device__ __forceinline__ HsAcc hs_eval_cta(const float* __restrict__ val, const HsLevel& L, float ex, float ey, float sn, float cs,
                                             const float2* __restrict__ pts, const float2* preg, HsCellCache* cache, int n,
                                             float pt_factor, float* __restrict__ prod, int pitch, int serial_warp = 0, HsMatchTimer* mt = nullptr,
                                             int* w_addr = nullptr, float* w_prob = nullptr, const uint64_t* s_tab = nullptr,
                                             int i0 = 0, int istride = 0, bool keep_in_l2 = false) {
  if (PPT > 0 && WC) {
    if ((i0 & ~31) < n)   // warp-uniform (the phase uses full-warp ballots): skip warps whose first point is past the scan
      hs_points_phase_wc<(PPT > 0 ? PPT : 1), WC == 2>(val, L, ex, ey, sn, cs, preg, cache, n, pt_factor, prod, pitch, w_addr, w_prob, s_tab, i0, istride, keep_in_l2);
  } else if (PPT > 0) {
#pragma unroll
    for (int j = 0; j < (PPT > 0 ? PPT : 1); ++j) {
      const int i = threadIdx.x + j * blockDim.x;
      if (i < n) {
        float gx, gy, rot, fv;
        hs_point_terms<true>(val, L, ex, ey, sn, cs, preg[j].x * pt_factor, preg[j].y * pt_factor, cache[j], gx, gy, rot, fv);
        hs_store_products(prod + i, pitch, gx, gy, rot, fv);
      }
    }
  } else {
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
      float2 p = pts[i];
      float gx, gy, rot, fv;
      HsCellCache none;
      hs_point_terms<false>(val, L, ex, ey, sn, cs, p.x * pt_factor, p.y * pt_factor, none, gx, gy, rot, fv);
      hs_store_products(prod + i, pitch, gx, gy, rot, fv);
    }
  }
  __syncthreads();
  if (mt) mt->mark(2);
  HsAcc a = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if ((threadIdx.x >> 5) == serial_warp) {
    const int lane = threadIdx.x & 31;
    float acc = 0.0f;
    if (STRICT) {
      if (lane < HS_NPROD) {
        const float* q = prod + lane * pitch;
        const float4* q4 = reinterpret_cast<const float4*>(q);
        const int n4 = n >> 2;
        for (int i = 0; i < n4; ++i) {  // loads are independent of the chain and issue ahead of it
          const float4 t = q4[i];
          acc += t.x; acc += t.y; acc += t.z; acc += t.w;
        }
        for (int i = n4 << 2; i < n; ++i) acc += q[i];
      }
      a.d0 = __shfl_sync(0xffffffffu, acc, 0);
      a.d1 = __shfl_sync(0xffffffffu, acc, 1);
      a.d2 = __shfl_sync(0xffffffffu, acc, 2);
      a.h00 = __shfl_sync(0xffffffffu, acc, 3);
      a.h11 = __shfl_sync(0xffffffffu, acc, 4);
      a.h22 = __shfl_sync(0xffffffffu, acc, 5);
      a.h01 = __shfl_sync(0xffffffffu, acc, 6);
      a.h02 = __shfl_sync(0xffffffffu, acc, 7);
      a.h12 = __shfl_sync(0xffffffffu, acc, 8);
    } else {
      float r[HS_NPROD];
#pragma unroll
      for (int k = 0; k < HS_NPROD; ++k) r[k] = 0.0f;
      for (int i = lane; i < n; i += 32) {
#pragma unroll
        for (int k = 0; k < HS_NPROD; ++k) r[k] += prod[k * pitch + i];
      }
#pragma unroll
      for (int m = 16; m >= 1; m >>= 1) {
#pragma unroll
        for (int k = 0; k < HS_NPROD; ++k) r[k] += __shfl_xor_sync(0xffffffffu, r[k], m);
      }
      a.d0 = r[0]; a.d1 = r[1]; a.d2 = r[2]; a.h00 = r[3]; a.h11 = r[4]; a.h22 = r[5]; a.h01 = r[6]; a.h02 = r[7]; a.h12 = r[8];
    }
  }
  if (mt) mt->mark(3);
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

// Tail of HectorSlamProcessor::update (HectorSlamProcessor.h:84-95), one thread per stream: store the matched pose, the
// poseDifferenceLargerThan gate, and -- if the scan is to be integrated -- the update epoch the ray-caster works in.
__device__ __forceinline__ void hs_match_tail(const HsParams& P, int slot, int s, float wx, float wy, float wth, bool mwm, int mode) {
  P.last_pose[3 * s] = wx; P.last_pose[3 * s + 1] = wy; P.last_pose[3 * s + 2] = wth;
  if (P.host_pose_out) { P.host_pose_out[3 * slot] = wx; P.host_pose_out[3 * slot + 1] = wy; P.host_pose_out[3 * slot + 2] = wth; }
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
      P.slot_sc[2 * slot] = hs_sinf(wth); P.slot_sc[2 * slot + 1] = hs_cosf(wth);   // Rotation2Df(pose[2]) of updateByScan (OccGridMapBase.h:130)
      atomicAdd(&P.counters[1], 1ULL);
    }
  }
  P.slot_integrate[slot] = integrate;
}

// k_match: HectorSlamProcessor::update minus the ray-cast (HSL/slam_main/HectorSlamProcessor.h:71-95):
//   MapRepMultiMap::matchData (MapRepMultiMap.h:116-132) -> ScanMatcher::matchData per level, coarse to fine
//   (ScanMatcher.h:54-190), then the poseDifferenceLargerThan gate and the per-stream bookkeeping.
// One CTA owns one stream for the whole coarse-to-fine solve (14 evaluations for 3 levels); the pose estimate
// lives in shared memory and nothing returns to the host in between. One warp carries the serial part
// (sums, 3x3 solve, sin/cos of the new angle); the other warps only help with the per-point gathers.
// mode: 0 = update (match + gate), 1 = match only (no gate, no integration).
// dynamic shared memory: 9 * pitch floats (products, sum-major) | WC: per warp PPT*128 floats + PPT*32 ints | WC: 32 x 8 B expf table.
// WC = 1: the cache misses of a warp are processed densely (hs_points_phase_wc); 0: every thread converts its own misses;
// 2 (multi-start matching only): misses are single 16-byte loads from the stream's probability block planes.
#define HS_MATCH_THREADS 128   // measured on B200 at 1024 streams: 64 -> 160 us, 96 -> 122, 128 -> 109, 192/256 -> 143 per launch

// 7 CTAs per SM (<= 72 registers) keeps all 1024 streams of BASELINE configs[1] resident in ONE wave on 148 SMs; at 73+
// registers the launch needs a second wave and takes 1.5x as long (measured: 94 -> 140 us).
// PITCH > 0: the pitch of the product rows is a compile-time constant (the engine instantiates 388 = scans of 357..384 points,
// BASELINE's 360-beam configurations): the nine stores per point and the chain's loads then address shared memory with
// immediates. The per-warp miss lists and the expf table sit IN FRONT of the products at offsets that depend on PPT only.
template <bool STRICT, int PPT, int WC, int PITCH = 0>
__global__ void __launch_bounds__(HS_MATCH_THREADS, (PPT > 0 && PPT <= 3) ? 7 : 1) k_match(HsParams P, int n_slots, const int32_t* __restrict__ stream_ids,
                                                            const float2* __restrict__ points, const int32_t* __restrict__ counts,
                                                            const float2* __restrict__ origos, const float* __restrict__ hints,
                                                            const uint8_t* __restrict__ mwm_flags, int mode) {
  extern __shared__ __align__(16) float hs_sm[];
  __shared__ float s_est[5];  // ex, ey, eth, sin(eth), cos(eth)
  __shared__ HsLevel s_lv;    // the current level's limx, limy, pt_factor, sx, max_iter (the other fields are not copied)
  __shared__ const float* s_val;   // ... and the first cell of its log-odds plane
  const int slot = P.slot0 + blockIdx.x;
  const int tid = threadIdx.x;
  if ((int)blockIdx.x >= n_slots) return;
  // dynamic shared memory: WC: 32 x 8 B expf table | per warp PPT*32 ints (block addresses) | per warp PPT*128 floats
  // (probabilities) -- then the products, 9 rows of `pitch` floats
  constexpr int wc_ppt = PPT > 0 ? PPT : 1;
  constexpr int n_warps_ct = HS_MATCH_THREADS / 32;
  constexpr int aux_floats = WC ? 64 + n_warps_ct * wc_ppt * 160 : 0;
  const int pitch = PITCH > 0 ? PITCH : hs_match2_pitch(P.max_points);
  uint64_t* s_tab = reinterpret_cast<uint64_t*>(hs_sm);
  int* w_addr = reinterpret_cast<int*>(hs_sm + 64) + (tid >> 5) * (wc_ppt * 32);
  float* w_prob = hs_sm + 64 + n_warps_ct * wc_ppt * 32 + (tid >> 5) * (wc_ppt * 128);
  float* hs_prod = hs_sm + aux_floats;
  if (WC && tid < 32) s_tab[tid] = hs_exp2f_tab_dev[tid];   // visible after the first __syncthreads below
  // the serial part (sums, 3x3 solve, sin/cos) is carried by ONE warp, chosen by the CTA index. CTAs are dealt out to the
  // SMs breadth first and 148 = 0 mod 4, so the CTAs sharing an SM pick the SAME warp number and their serial warps share
  // one SM sub-partition (warp w issues from scheduler w mod 4), where their 4-cycle FADD chains interleave. Measured on
  // B200: spreading them over the four sub-partitions ((blockIdx + blockIdx / n_sm) & 3) costs 5-8 us per launch.
  const int sw = blockIdx.x & ((blockDim.x >> 5) - 1);
  const bool serial = (tid >> 5) == sw, leader = tid == (sw << 5);
  const bool hyp = P.hyp_pose != nullptr;
  const int s = hyp ? P.hyp_stream : hs_slot_stream(stream_ids, slot);
  // point i = i0 + j * istride of the scan is this thread's j-th. (Measured on B200: leaving the serial warp without points,
  // so that its SM sub-partition only carries the latency-critical chains, and giving the other three warps 4 points per
  // thread, costs 18 us per launch -- the points phase needs all four sub-partitions.)
  const int i0 = tid, istride = (int)blockDim.x;
  const float2* pts = points + (hyp ? (size_t)0 : (size_t)slot * P.max_points);
  // register-resident points: requested BEFORE the count is known (a row always has max_points slots), so that the prologue
  // is one memory round trip (count, hint and points together) instead of count -> points
  float2 preg[PPT > 0 ? PPT : 1];
  if (PPT > 0 && !P.stage_pts) {
#pragma unroll
    for (int j = 0; j < (PPT > 0 ? PPT : 1); ++j) {
      const int i = i0 + j * istride;
      preg[j] = i < P.max_points ? pts[i] : make_float2(0.f, 0.f);
    }
  }
  int n = counts[hyp ? 0 : slot];
  n = n < 0 ? 0 : (n > P.max_points ? P.max_points : n);
  if (P.stage_pts) {   // zero-copy input: fetch this scan from the caller's pinned host buffer once
    if (P.stage_cnt && tid == 0) P.stage_cnt[slot] = n;
    float2* stg = P.stage_pts + (size_t)slot * P.max_points;
    // counts on the host too: fetch the whole row, so that the point loads do not wait for the count's round trip
    const int n_fetch = P.stage_cnt ? P.max_points : n;
    for (int i = tid; i < n_fetch; i += blockDim.x) stg[i] = pts[i];
    __syncthreads();
    pts = stg;
    if (PPT > 0) {
#pragma unroll
      for (int j = 0; j < (PPT > 0 ? PPT : 1); ++j) {
        const int i = i0 + j * istride;
        preg[j] = i < n ? pts[i] : make_float2(0.f, 0.f);
      }
    }
  }
  const bool mwm = mwm_flags ? (mwm_flags[slot] != 0) : false;
  float wx, wy, wth;
  if (hints) { wx = hints[3 * slot]; wy = hints[3 * slot + 1]; wth = hints[3 * slot + 2]; }
  else { wx = P.last_pose[3 * s]; wy = P.last_pose[3 * s + 1]; wth = P.last_pose[3 * s + 2]; }
  if (PPT > 0) {
#pragma unroll
    for (int j = 0; j < (PPT > 0 ? PPT : 1); ++j)
      if (i0 + j * istride >= n) preg[j] = make_float2(0.f, 0.f);
  }

  if (!mwm) {
    // dataContainers[level-1].setFrom(dataContainer, 2^-level) (MapRepMultiMap.h:127): keep the level-0
    // points of the last matched scan; the power-of-two factor is applied (exactly) where they are used.
    float2* cp = reinterpret_cast<float2*>(P.coarse_pts) + (size_t)s * P.max_points;
    if (!hyp) {
      if (PPT > 0) {
#pragma unroll
        for (int j = 0; j < (PPT > 0 ? PPT : 1); ++j) {
          const int i = i0 + j * istride;
          if (i < n) cp[i] = preg[j];
        }
      } else {
        for (int i = tid; i < n; i += blockDim.x) cp[i] = pts[i];
      }
    }
    if (tid == 0 && !hyp) {
      P.coarse_cnt[s] = n;
      float2 o = origos ? origos[slot] : make_float2(0.f, 0.f);
      P.coarse_origo[2 * s] = o.x;
      P.coarse_origo[2 * s + 1] = o.y;
    }
    if (n != 0) {
      const float* val_s = P.val + (size_t)s * P.cells_per_stream;
      HsAcc a = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      unsigned clamps = 0;
      HsMatchTimer mt;
      mt.begin(P.dbg_times);
      HsCellCache cache[PPT > 0 ? PPT : 1];
      for (int level = P.n_levels - 1; level >= 0; --level) {
        const HsLevel& Lc = P.lv[level];   // constant bank, indexed by a run-time level: only touched here, once per level
        float ex = 0.f, ey = 0.f, eth = wth;
#pragma unroll
        for (int j = 0; j < (PPT > 0 ? PPT : 1); ++j) cache[j].key = -1;   // a new level is a new grid
        if (serial) {
          hs_world_to_map(Lc, wx, wy, ex, ey);
          if (leader) {
            s_est[0] = ex; s_est[1] = ey; s_est[2] = eth; s_est[3] = hs_sinf(eth); s_est[4] = hs_cosf(eth);
            // what every evaluation needs of the level goes to shared memory: with `P.lv[level]` read in place, register-indexed
            // constant loads (LDC c[0x0][R+imm]) sat in front of every use -- the loop bound's was the hottest instruction of
            // the kernel (13.5 % of the stall samples) -- and the plane pointer was re-derived inside the miss loop. Same-box
            // A/B on B200, 1024 streams: 99.1 us (constant bank) -> 96.0 (level fields here) -> 95.6 (+ the plane pointer)
            s_lv.limx = Lc.limx; s_lv.limy = Lc.limy; s_lv.pt_factor = Lc.pt_factor; s_lv.sx = Lc.sx; s_lv.max_iter = Lc.max_iter;
            s_val = WC == 2 ? reinterpret_cast<const float*>(P.hyp_pblk + Lc.cell_offset) : val_s + Lc.cell_offset;
          }
        }
        __syncthreads();
        const HsLevel& L = s_lv;   // read with immediate-address LDS wherever the evaluation needs a field
        const float* val = s_val;
        const int max_iter = L.max_iter;
        const bool keep_l2 = ((P.l2_keep_levels >> level) & 1) != 0;
        for (int it = 0; it <= max_iter; ++it) {  // 1 + maxIterations evaluations, no early exit
          const float cex = s_est[0], cey = s_est[1], csn = s_est[3], ccs = s_est[4];
          a = hs_eval_cta<STRICT, PPT, WC>(val, L, cex, cey, csn, ccs, pts, preg, cache, n, L.pt_factor, hs_prod, pitch, sw, &mt,
                                           w_addr, w_prob, s_tab, i0, istride, keep_l2);
          if (serial) {
            clamps += hs_gauss_newton_step(a, ex, ey, eth) ? 1u : 0u;
            if (leader) { s_est[0] = ex; s_est[1] = ey; s_est[2] = eth; s_est[3] = hs_sinf(eth); s_est[4] = hs_cosf(eth); }
          }
          __syncthreads();
          mt.mark(4);
        }
        if (serial) {
          eth = hs_normalize_angle(eth);
          hs_map_to_world(Lc, ex, ey, wx, wy);
          wth = eth;
        }
      }
      if (leader) {
        float* c = hyp ? P.hyp_cov + 9 * (size_t)slot : P.cov + 9 * (size_t)s;  // covMatrix = H of the last evaluation (ScanMatcher.h:184)
        c[0] = a.h00; c[1] = a.h01; c[2] = a.h02;
        c[3] = a.h01; c[4] = a.h11; c[5] = a.h12;
        c[6] = a.h02; c[7] = a.h12; c[8] = a.h22;
        if (clamps) atomicAdd(&P.counters[2], (unsigned long long)clamps);
      }
      mt.end();
    }
  }
  if (leader) {
    if (hyp) { P.hyp_pose[3 * slot] = wx; P.hyp_pose[3 * slot + 1] = wy; P.hyp_pose[3 * slot + 2] = wth; }
    else hs_match_tail(P, slot, s, wx, wy, wth, mwm, mode);
  }
}

// ------------------------------------------------------------------------------------------------
// k_match2: the same computation as k_match, organised so that the expensive part -- log-odds -> probability
// (a double-precision expf and an IEEE division per cell) -- runs DENSE:
//   step 1  every point: transform, bounds test, 2x2 block address k; the block's four probabilities are memoised in
//           shared memory (pcache[i], key kcell[i]); points whose block changed since the last evaluation go to a miss list
//           (Gauss-Newton steps on one level are mostly sub-cell, so after the first evaluation most points hit)
//   step 2  the miss list x 4 cells is dealt out to ALL threads: gather (four loads in flight per thread), expf, division
//   step 3  every point: bilinear value and gradient from pcache, the nine products -> prodT[9][pitch] (sum-major, so the
//           strict sums read 4 consecutive products per 16-byte shared load)
//   sums    as in hs_eval_cta: lanes 0..8 of warp 0, one dependent FADD chain per sum in the reference's point order.
// Points, keys and probabilities live in shared memory, so any scan length works with one instantiation.
// dynamic shared memory: 9 * pitch floats | max_points float4 | max_points float2 | max_points int | max_points ushort.

template <bool STRICT>
__device__ __forceinline__ HsAcc hs_eval_cta2(const float* __restrict__ val, const HsLevel& L, float ex, float ey, float sn, float cs,
                                              const float2* __restrict__ spts, int n, float pt_factor, float* __restrict__ prodT, int pitch,
                                              float4* __restrict__ pcache, int* __restrict__ kcell, unsigned short* __restrict__ miss,
                                              int* __restrict__ s_nmiss, int serial_warp, HsMatchTimer& mt) {
  const int T = blockDim.x, tid = threadIdx.x, lane = tid & 31;
  const int W = L.sx;
  // step 1: block addresses, miss list
  for (int i0 = 0; i0 < n; i0 += T) {
    const int i = i0 + tid;
    bool m = false;
    if (i < n) {
      const float2 p = spts[i];
      const float px = p.x * pt_factor, py = p.y * pt_factor;
      const float qx = (cs * px + (-sn) * py) + ex;  // Affine2f * v = linear*v + translation
      const float qy = (sn * px + cs * py) + ey;
      if (qx >= 0.0f && qx <= L.limx && qy >= 0.0f && qy <= L.limy) {  // !pointOutOfMapBounds
        const int k = (int)qy * W + (int)qx;
        if (k != kcell[i]) { kcell[i] = k; m = true; }
      }
    }
    const unsigned b = __ballot_sync(0xffffffffu, m);
    int base = 0;
    if (lane == 0 && b) base = atomicAdd(s_nmiss, __popc(b));
    base = __shfl_sync(0xffffffffu, base, 0);
    if (m) miss[base + __popc(b & ((1u << lane) - 1u))] = (unsigned short)i;
  }
  __syncthreads();
  mt.mark(0);
  // step 2: probabilities of the missed blocks, one (block, corner) item per thread and turn
  {
    const int items = *s_nmiss << 2;
    float* pc = reinterpret_cast<float*>(pcache);
    for (int q0 = tid; q0 < items; q0 += 4 * T) {
      float v[4];
      int dst[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int q = q0 + u * T;
        dst[u] = -1;
        if (q < items) {
          const int i = miss[q >> 2], c = q & 3;
          dst[u] = (i << 2) | c;
          v[u] = __ldg(val + kcell[i] + (c & 1) + ((c >> 1) ? W : 0));
        }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (dst[u] >= 0) pc[dst[u]] = hs_probability(v[u]);
    }
  }
  __syncthreads();
  mt.mark(1);
  // step 3: per-point terms and the nine products
  for (int i = tid; i < n; i += T) {
    const float2 p = spts[i];
    const float px = p.x * pt_factor, py = p.y * pt_factor;
    const float qx = (cs * px + (-sn) * py) + ex;
    const float qy = (sn * px + cs * py) + ey;
    float m = 0.0f, gx = 0.0f, gy = 0.0f;
    if (qx >= 0.0f && qx <= L.limx && qy >= 0.0f && qy <= L.limy) {
      const int ix = (int)qx, iy = (int)qy;
      const float fx = qx - (float)ix, fy = qy - (float)iy;
      const float4 c4 = pcache[i];
      const float i0 = c4.x, i1 = c4.y, i2 = c4.z, i3 = c4.w;
      const float dx1 = i0 - i1, dx2 = i2 - i3;
      const float dy1 = i0 - i2, dy2 = i1 - i3;
      const float x_inv = 1.0f - fx, y_inv = 1.0f - fy;
      m = ((i0 * x_inv + i1 * fx) * y_inv) + ((i2 * x_inv + i3 * fx) * fy);
      gx = -((dx1 * x_inv) + (dx2 * fx));  // the reference blends with the x factors here (OccGridMapUtil.h:344)
      gy = -((dy1 * y_inv) + (dy2 * fy));  // ... and with the y factors here (:345)
    }
    const float fv = 1.0f - m;
    const float rot = ((-sn * px - cs * py) * gx + (cs * px - sn * py) * gy);
    float* q = prodT + i;
    q[0] = gx * fv;             // dTr[0]
    q[pitch] = gy * fv;         // dTr[1]
    q[2 * pitch] = rot * fv;    // dTr[2]
    q[3 * pitch] = gx * gx;     // H(0,0)
    q[4 * pitch] = gy * gy;     // H(1,1)
    q[5 * pitch] = rot * rot;   // H(2,2)
    q[6 * pitch] = gx * gy;     // H(0,1)
    q[7 * pitch] = gx * rot;    // H(0,2)
    q[8 * pitch] = gy * rot;    // H(1,2)
  }
  __syncthreads();
  mt.mark(2);
  if (tid == 0) *s_nmiss = 0;   // every thread has read it (step 2) before the barrier above; next use is after the caller's barrier
  HsAcc a = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if ((tid >> 5) == serial_warp) {
    float r9[HS_NPROD];
    if (STRICT) {
      float acc = 0.0f;
      if (lane < HS_NPROD) {
        const float* q = prodT + lane * pitch;
        const float4* q4 = reinterpret_cast<const float4*>(q);
        const int n4 = n >> 2;
        for (int i = 0; i < n4; ++i) {   // loads are independent of the chain and issue ahead of it
          const float4 t = q4[i];
          acc += t.x; acc += t.y; acc += t.z; acc += t.w;
        }
        for (int i = n4 << 2; i < n; ++i) acc += q[i];
      }
#pragma unroll
      for (int k = 0; k < HS_NPROD; ++k) r9[k] = __shfl_sync(0xffffffffu, acc, k);
    } else {
#pragma unroll
      for (int k = 0; k < HS_NPROD; ++k) r9[k] = 0.0f;
      for (int i = lane; i < n; i += 32) {
#pragma unroll
        for (int k = 0; k < HS_NPROD; ++k) r9[k] += prodT[k * pitch + i];
      }
#pragma unroll
      for (int m = 16; m >= 1; m >>= 1) {
#pragma unroll
        for (int k = 0; k < HS_NPROD; ++k) r9[k] += __shfl_xor_sync(0xffffffffu, r9[k], m);
      }
    }
    a.d0 = r9[0]; a.d1 = r9[1]; a.d2 = r9[2]; a.h00 = r9[3]; a.h11 = r9[4]; a.h22 = r9[5]; a.h01 = r9[6]; a.h02 = r9[7]; a.h12 = r9[8];
  }
  mt.mark(3);
  return a;
}

// THREADS per CTA: 128 when there are enough scans to fill the device with one 4-warp CTA each, 512 for small batches of long
// scans (BASELINE configs[2]: 256 scans of 1440 points on 148 SMs) -- the point phases scale with the threads, the chains do not.
template <bool STRICT, int THREADS>
__global__ void __launch_bounds__(THREADS, 2) k_match2(HsParams P, int n_slots, const int32_t* __restrict__ stream_ids,
                                                             const float2* __restrict__ points, const int32_t* __restrict__ counts,
                                                             const float2* __restrict__ origos, const float* __restrict__ hints,
                                                             const uint8_t* __restrict__ mwm_flags, int mode) {
  extern __shared__ __align__(16) float hs_m2[];
  __shared__ float s_est[5];  // ex, ey, eth, sin(eth), cos(eth)
  __shared__ HsLevel s_lv;    // the current level's limx, limy, pt_factor, sx, max_iter
  __shared__ const float* s_val;
  __shared__ int s_nmiss;
  const int pitch = hs_match2_pitch(P.max_points);
  float* prodT = hs_m2;
  float4* pcache = reinterpret_cast<float4*>(prodT + HS_NPROD * pitch);
  float2* spts = reinterpret_cast<float2*>(pcache + P.max_points);
  int* kcell = reinterpret_cast<int*>(spts + P.max_points);
  unsigned short* miss = reinterpret_cast<unsigned short*>(kcell + P.max_points);
  const int slot = P.slot0 + blockIdx.x;
  const int tid = threadIdx.x;
  if ((int)blockIdx.x >= n_slots) return;
  // the serial part (sums, 3x3 solve, sin/cos) is carried by ONE warp, chosen by the CTA index. CTAs are dealt out to the
  // SMs breadth first and 148 = 0 mod 4, so the CTAs sharing an SM pick the SAME warp number and their serial warps share
  // one SM sub-partition (warp w issues from scheduler w mod 4), where their 4-cycle FADD chains interleave. Measured on
  // B200: spreading them over the four sub-partitions ((blockIdx + blockIdx / n_sm) & 3) costs 5-8 us per launch.
  const int sw = blockIdx.x & ((blockDim.x >> 5) - 1);
  const bool serial = (tid >> 5) == sw, leader = tid == (sw << 5);
  const bool hyp = P.hyp_pose != nullptr;
  const int s = hyp ? P.hyp_stream : hs_slot_stream(stream_ids, slot);
  int n = counts[hyp ? 0 : slot];
  n = n < 0 ? 0 : (n > P.max_points ? P.max_points : n);
  const float2* pts = points + (hyp ? (size_t)0 : (size_t)slot * P.max_points);
  if (P.stage_pts) {   // zero-copy input: fetch this scan from the caller's pinned host buffer once
    if (P.stage_cnt && tid == 0) P.stage_cnt[slot] = n;
    float2* stg = P.stage_pts + (size_t)slot * P.max_points;
    // counts on the host too: fetch the whole row, so that the point loads do not wait for the count's round trip
    const int n_fetch = P.stage_cnt ? P.max_points : n;
    for (int i = tid; i < n_fetch; i += blockDim.x) stg[i] = pts[i];
    __syncthreads();
    pts = stg;
  }
  const bool mwm = mwm_flags ? (mwm_flags[slot] != 0) : false;
  float wx, wy, wth;
  if (hints) { wx = hints[3 * slot]; wy = hints[3 * slot + 1]; wth = hints[3 * slot + 2]; }
  else { wx = P.last_pose[3 * s]; wy = P.last_pose[3 * s + 1]; wth = P.last_pose[3 * s + 2]; }

  if (!mwm) {
    // dataContainers[level-1].setFrom(dataContainer, 2^-level) (MapRepMultiMap.h:127): keep the level-0
    // points of the last matched scan; the power-of-two factor is applied (exactly) where they are used.
    float2* cp = reinterpret_cast<float2*>(P.coarse_pts) + (size_t)s * P.max_points;
    for (int i = tid; i < n; i += blockDim.x) { const float2 p = pts[i]; if (!hyp) cp[i] = p; spts[i] = p; }
    if (tid == 0) {
      if (!hyp) {
        P.coarse_cnt[s] = n;
        float2 o = origos ? origos[slot] : make_float2(0.f, 0.f);
        P.coarse_origo[2 * s] = o.x;
        P.coarse_origo[2 * s + 1] = o.y;
      }
      s_nmiss = 0;
    }
    if (n != 0) {
      const float* val_s = P.val + (size_t)s * P.cells_per_stream;
      HsAcc a = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      unsigned clamps = 0;
      HsMatchTimer mt;
      mt.begin(P.dbg_times);
      for (int level = P.n_levels - 1; level >= 0; --level) {
        const HsLevel& Lc = P.lv[level];   // constant bank, indexed by a run-time level: only touched here, once per level
        float ex = 0.f, ey = 0.f, eth = wth;
        for (int i = tid; i < n; i += blockDim.x) kcell[i] = -1;   // a new level is a new grid
        if (serial) {
          hs_world_to_map(Lc, wx, wy, ex, ey);
          if (leader) {
            s_est[0] = ex; s_est[1] = ey; s_est[2] = eth; s_est[3] = hs_sinf(eth); s_est[4] = hs_cosf(eth);
            s_lv.limx = Lc.limx; s_lv.limy = Lc.limy; s_lv.pt_factor = Lc.pt_factor; s_lv.sx = Lc.sx; s_lv.max_iter = Lc.max_iter;   // see k_match
            s_val = val_s + Lc.cell_offset;
          }
        }
        __syncthreads();
        const HsLevel& L = s_lv;
        const float* val = s_val;
        const int max_iter = s_lv.max_iter;
        for (int it = 0; it <= max_iter; ++it) {  // 1 + maxIterations evaluations, no early exit
          const float cex = s_est[0], cey = s_est[1], csn = s_est[3], ccs = s_est[4];
          a = hs_eval_cta2<STRICT>(val, L, cex, cey, csn, ccs, spts, n, L.pt_factor, prodT, pitch, pcache, kcell, miss, &s_nmiss, sw, mt);
          if (serial) {
            clamps += hs_gauss_newton_step(a, ex, ey, eth) ? 1u : 0u;
            if (leader) { s_est[0] = ex; s_est[1] = ey; s_est[2] = eth; s_est[3] = hs_sinf(eth); s_est[4] = hs_cosf(eth); }
          }
          __syncthreads();
          mt.mark(4);
        }
        if (serial) {
          eth = hs_normalize_angle(eth);
          hs_map_to_world(Lc, ex, ey, wx, wy);
          wth = eth;
        }
      }
      if (leader) {
        float* c = hyp ? P.hyp_cov + 9 * (size_t)slot : P.cov + 9 * (size_t)s;  // covMatrix = H of the last evaluation (ScanMatcher.h:184)
        c[0] = a.h00; c[1] = a.h01; c[2] = a.h02;
        c[3] = a.h01; c[4] = a.h11; c[5] = a.h12;
        c[6] = a.h02; c[7] = a.h12; c[8] = a.h22;
        if (clamps) atomicAdd(&P.counters[2], (unsigned long long)clamps);
      }
      mt.end();
    }
  }
  if (leader) {
    if (hyp) { P.hyp_pose[3 * slot] = wx; P.hyp_pose[3 * slot + 1] = wy; P.hyp_pose[3 * slot + 2] = wth; }
    else hs_match_tail(P, slot, s, wx, wy, wth, mwm, mode);
  }
}

// k_force_integrate: integration set-up with a caller-given pose (hs_raycast_batch / hs_integrate_batch): take the pose from
// the caller and open an update epoch, as updateByScan would. refresh_coarse != 0 (teacher forcing): the coarse containers
// are first refreshed from the given points, as matchData's setFrom would have; 0: they stay as the last matchData left
// them -- MapRepMultiMap::updateByScan proper (MapRepMultiMap.h:134-147).
__global__ void __launch_bounds__(128) k_force_integrate(HsParams P, int n_slots, const int32_t* __restrict__ stream_ids,
                                                         const float2* __restrict__ points, const int32_t* __restrict__ counts,
                                                         const float2* __restrict__ origos, const float* __restrict__ poses,
                                                         int refresh_coarse) {
  int slot = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  if (slot >= n_slots) return;
  int s = hs_slot_stream(stream_ids, slot);
  int n = counts[slot];
  n = n < 0 ? 0 : (n > P.max_points ? P.max_points : n);
  const float2* pts = points + (size_t)slot * P.max_points;
  float2* cp = reinterpret_cast<float2*>(P.coarse_pts) + (size_t)s * P.max_points;
  if (refresh_coarse) for (int i = lane; i < n; i += 32) cp[i] = pts[i];
  if (lane == 0) {
    if (refresh_coarse) {
      P.coarse_cnt[s] = n;
      float2 o = origos ? origos[slot] : make_float2(0.f, 0.f);
      P.coarse_origo[2 * s] = o.x;
      P.coarse_origo[2 * s + 1] = o.y;
    }
    int cur = P.cur_update_index[s];
    P.slot_mark_base[slot] = cur;
    P.cur_update_index[s] = cur + 3;
    P.last_update_index[s] += 1;
    P.slot_pose[3 * slot] = poses[3 * slot]; P.slot_pose[3 * slot + 1] = poses[3 * slot + 1]; P.slot_pose[3 * slot + 2] = poses[3 * slot + 2];
    P.slot_sc[2 * slot] = hs_sinf(poses[3 * slot + 2]); P.slot_sc[2 * slot + 1] = hs_cosf(poses[3 * slot + 2]);
    P.slot_integrate[slot] = 1;
    atomicAdd(&P.counters[1], 1ULL);
  }
}

// ------------------------------------------------------------------------------------------------
// BASELINE configs[3]: getCompleteHessianDerivs (OccGridMapUtil.h:64-104) of ONE stream's level for thousands of pose
// hypotheses of one scan (relocalisation-style multi-start).
//
// Every hypothesis gathers from the same map, and log-odds -> probability (GridMapLogOddsFunctions::getGridProbability,
// GridMapLogOdds.h:163-166: a double-precision expf and an IEEE division per cell) is a pure function of the cell. So the
// conversion is done ONCE per cell into a "probability block plane" -- the B200 form of the reference's GridMapCacheArray
// (HSL/map/GridMapCacheArray.h), laid out for the gather instead of per cell:
//     pblk[k] = ( P(k), P(k+1), P(k+W), P(k+W+1) )       16 bytes: the 2x2 block interpMapValueWithDerivatives reads
//                                                         at a point whose integer cell is k (OccGridMapUtil.h:300-335)
// A point evaluation is then ONE 16-byte load (one 32-byte sector) and the bilinear arithmetic; 4096 hypotheses x 360 points
// do 1.47 M block loads instead of 5.9 M cell loads + 5.9 M expf/divisions. The plane of a level (16 MB for 1024^2) stays in
// the 126 MB L2 while the hypotheses run; the engine keeps it until the stream's maps change.

// k_build_pblk: all levels of one stream in one launch (blockIdx.z = level; the plane uses the stream slab's cell offsets).
// One 32 x 8 tile of blocks per CTA; the tile's 33 x 9 probabilities are computed once into shared memory.
#define HS_PB_TX 32
#define HS_PB_TY 8
__global__ void __launch_bounds__(HS_PB_TX * HS_PB_TY) k_build_pblk(HsParams P, int stream, float4* __restrict__ pblk_stream) {
  __shared__ float sp[HS_PB_TY + 1][HS_PB_TX + 1];
  const HsLevel& L = P.lv[blockIdx.z];
  const int W = L.sx, H = L.sy;
  const int x0 = blockIdx.x * HS_PB_TX, y0 = blockIdx.y * HS_PB_TY;
  if (x0 >= W || y0 >= H) return;
  const float* __restrict__ val = P.val + (size_t)stream * P.cells_per_stream + L.cell_offset;
  float4* __restrict__ pblk = pblk_stream + L.cell_offset;
  for (int t = threadIdx.x; t < (HS_PB_TY + 1) * (HS_PB_TX + 1); t += blockDim.x) {
    const int ty = t / (HS_PB_TX + 1), tx = t - ty * (HS_PB_TX + 1);
    const int x = x0 + tx, y = y0 + ty;
    float pr = 0.0f;
    if (x < W && y < H) pr = hs_probability(__ldg(val + (size_t)y * W + x));
    sp[ty][tx] = pr;
  }
  __syncthreads();
  const int tx = threadIdx.x & (HS_PB_TX - 1), ty = threadIdx.x / HS_PB_TX;
  const int x = x0 + tx, y = y0 + ty;
  if (x < W && y < H) pblk[(size_t)y * W + x] = make_float4(sp[ty][tx], sp[ty][tx + 1], sp[ty + 1][tx], sp[ty + 1][tx + 1]);
}

// Per-point terms from the four probabilities of the block (the arithmetic of hs_point_terms, OccGridMapUtil.h:318-346).
__device__ __forceinline__ void hs_block_terms(const float4 b, float fx, float fy, float sn, float cs, float px, float py, bool inside,
                                               float& gx, float& gy, float& rot, float& fv) {
  float m = 0.0f;
  gx = 0.0f;
  gy = 0.0f;
  if (inside) {
    const float dx1 = b.x - b.y, dx2 = b.z - b.w;
    const float dy1 = b.x - b.z, dy2 = b.y - b.w;
    const float x_inv = 1.0f - fx, y_inv = 1.0f - fy;
    m = ((b.x * x_inv + b.y * fx) * y_inv) + ((b.z * x_inv + b.w * fx) * fy);
    gx = -((dx1 * x_inv) + (dx2 * fx));  // the reference blends with the x factors here (OccGridMapUtil.h:344)
    gy = -((dy1 * y_inv) + (dy2 * fy));  // ... and with the y factors here (:345)
  }
  fv = 1.0f - m;
  rot = ((-sn * px - cs * py) * gx + (cs * px - sn * py) * gy);
}

// k_hessian_derivs: every WARP works on its own group of PPW hypotheses, no CTA barrier after the points are staged.
//   point phase  the 32 lanes evaluate a chunk of HS_HD_CHUNK points for each of the PPW poses (all block loads of the chunk
//                are issued before the first is used) and leave the nine products per point in the warp's shared-memory
//                buffer, sum-major: w[(j*9 + k) * pitch + e]
//   chain phase  STRICT: lane j*9 + k (27 of 32 lanes busy for PPW = 3) extends sum k of pose j by the chunk's products in
//                the reference's point order (OccGridMapUtil.h:76-98), four terms per 16-byte shared load -- the dependent
//                FADD chain of one pose costs a warp instruction per term, so carrying three poses per warp divides the
//                serial part's issue slots by three
//                tree mode: no shared memory, every lane keeps 9 partial sums per pose over its points, xor-butterfly at the end
// poses are MAP coordinates of the level, pts already scaled to it. dynamic smem: n_pts float2 (padded to 16 B) | per warp
// PPW * 9 * (HS_HD_CHUNK + 4) floats.
#ifndef HS_HD_CHUNK
#define HS_HD_CHUNK 64
#endif
#define HS_HD_PITCH (HS_HD_CHUNK + 4)   // == 4 (mod 32): conflict-free 16-byte reads by lanes that own different sums
#define HS_HD_THREADS 128

template <bool STRICT, int PPW>
__global__ void __launch_bounds__(HS_HD_THREADS) k_hessian_derivs(const float4* __restrict__ pblk, HsLevel L, const float* __restrict__ poses_map,
                                                                  int n_poses, const float2* __restrict__ pts, int n_pts,
                                                                  float* __restrict__ H9, float* __restrict__ dTr3) {
  extern __shared__ __align__(16) float hs_hd[];
  float2* spts = reinterpret_cast<float2*>(hs_hd);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float* w = hs_hd + 2 * ((n_pts + 1) & ~1) + warp * (PPW * HS_NPROD * HS_HD_PITCH);
  for (int i = threadIdx.x; i < n_pts; i += blockDim.x) spts[i] = pts[i];
  __syncthreads();
  const int W = L.sx;
  const float limx = L.limx, limy = L.limy;
  const int n_groups = (n_poses + PPW - 1) / PPW;
  constexpr int R = HS_HD_CHUNK / 32;
  for (int g = blockIdx.x * (HS_HD_THREADS / 32) + warp; g < n_groups; g += gridDim.x * (HS_HD_THREADS / 32)) {
    // lane j < PPW loads pose j of the group and computes its sin/cos once; everybody else gets them by shuffle
    float mx = 0.0f, my = 0.0f, msn = 0.0f, mcs = 1.0f;
    if (lane < PPW && g * PPW + lane < n_poses) {
      const float* pm = poses_map + 3 * (size_t)(g * PPW + lane);
      mx = pm[0]; my = pm[1];
      const float th = pm[2];
      msn = hs_sinf(th); mcs = hs_cosf(th);
    }
    float ex[PPW], ey[PPW], sn[PPW], cs[PPW];
#pragma unroll
    for (int j = 0; j < PPW; ++j) {
      ex[j] = __shfl_sync(0xffffffffu, mx, j); ey[j] = __shfl_sync(0xffffffffu, my, j);
      sn[j] = __shfl_sync(0xffffffffu, msn, j); cs[j] = __shfl_sync(0xffffffffu, mcs, j);
    }
    float acc = 0.0f;                    // STRICT: sum (lane % 9) of pose (lane / 9)
    float part[STRICT ? 1 : PPW][HS_NPROD];   // tree mode: this lane's partial sums
    if (!STRICT) {
#pragma unroll
      for (int j = 0; j < PPW; ++j)
#pragma unroll
        for (int k = 0; k < HS_NPROD; ++k) part[j][k] = 0.0f;
    }
    // software pipeline over the chunks: the block loads of chunk c+1 are issued BEFORE the chain phase of chunk c, so the L2
    // round trip of the gathers hides behind the serial adds instead of preceding them
    float4 blk[PPW][R];
    float fxs[PPW][R], fys[PPW][R];
    unsigned inside = 0u;
    auto gather = [&](int c0) {   // point phase, step 1: block addresses of chunk c0, all loads in flight
      inside = 0u;
#pragma unroll
      for (int j = 0; j < PPW; ++j) {
#pragma unroll
        for (int r = 0; r < R; ++r) {
          const int i = c0 + r * 32 + lane;
          blk[j][r] = make_float4(0.f, 0.f, 0.f, 0.f);
          fxs[j][r] = 0.0f; fys[j][r] = 0.0f;
          if (i < n_pts) {
            const float2 p = spts[i];
            const float qx = (cs[j] * p.x + (-sn[j]) * p.y) + ex[j];  // Affine2f * v = linear*v + translation
            const float qy = (sn[j] * p.x + cs[j] * p.y) + ey[j];
            if (qx >= 0.0f && qx <= limx && qy >= 0.0f && qy <= limy) {  // !pointOutOfMapBounds
              const int ix = (int)qx, iy = (int)qy;
              fxs[j][r] = qx - (float)ix; fys[j][r] = qy - (float)iy;
              blk[j][r] = __ldg(pblk + (size_t)iy * W + ix);
              inside |= 1u << (j * R + r);
            }
          }
        }
      }
    };
    gather(0);
    for (int c0 = 0; c0 < n_pts; c0 += HS_HD_CHUNK) {
      const int len = min(HS_HD_CHUNK, n_pts - c0);
      // step 2: terms and products of chunk c0
#pragma unroll
      for (int j = 0; j < PPW; ++j) {
#pragma unroll
        for (int r = 0; r < R; ++r) {
          const int e = r * 32 + lane;
          if (c0 + e < n_pts) {
            const float2 p = spts[c0 + e];
            float gx, gy, rot, fv;
            hs_block_terms(blk[j][r], fxs[j][r], fys[j][r], sn[j], cs[j], p.x, p.y, (inside >> (j * R + r)) & 1u, gx, gy, rot, fv);
            if (STRICT) {
              hs_store_products(w + j * HS_NPROD * HS_HD_PITCH + e, HS_HD_PITCH, gx, gy, rot, fv);
            } else {
              float* a = part[STRICT ? 0 : j];
              a[0] += gx * fv; a[1] += gy * fv; a[2] += rot * fv; a[3] += gx * gx; a[4] += gy * gy; a[5] += rot * rot;
              a[6] += gx * gy; a[7] += gx * rot; a[8] += gy * rot;
            }
          }
        }
      }
      if (c0 + HS_HD_CHUNK < n_pts) gather(c0 + HS_HD_CHUNK);   // in flight during the chain phase below
      if (STRICT) {
        __syncwarp();
        if (lane < PPW * HS_NPROD) {
          const float* q = w + lane * HS_HD_PITCH;
          const float4* q4 = reinterpret_cast<const float4*>(q);
          const int n4 = len >> 2;
          for (int e = 0; e < n4; ++e) {   // loads are independent of the chain and issue ahead of it
            const float4 t = q4[e];
            acc += t.x; acc += t.y; acc += t.z; acc += t.w;
          }
          for (int e = n4 << 2; e < len; ++e) acc += q[e];
        }
        __syncwarp();
      }
    }
    // results: sum order d0 d1 d2 h00 h11 h22 h01 h02 h12 (HsAcc)
    if (STRICT) {
      const int j = lane / HS_NPROD, k = lane - j * HS_NPROD;
      const int p = g * PPW + j;
      if (lane < PPW * HS_NPROD && p < n_poses) {
        float* h = H9 + 9 * (size_t)p;
        float* d = dTr3 + 3 * (size_t)p;
        switch (k) {
          case 0: d[0] = acc; break;
          case 1: d[1] = acc; break;
          case 2: d[2] = acc; break;
          case 3: h[0] = acc; break;
          case 4: h[4] = acc; break;
          case 5: h[8] = acc; break;
          case 6: h[1] = acc; h[3] = acc; break;
          case 7: h[2] = acc; h[6] = acc; break;
          default: h[5] = acc; h[7] = acc; break;
        }
      }
    } else {
#pragma unroll
      for (int j = 0; j < PPW; ++j) {
        float* a = part[STRICT ? 0 : j];
#pragma unroll
        for (int m = 16; m >= 1; m >>= 1) {
#pragma unroll
          for (int k = 0; k < HS_NPROD; ++k) a[k] += __shfl_xor_sync(0xffffffffu, a[k], m);
        }
        const int p = g * PPW + j;
        if (lane == 0 && p < n_poses) {
          float* h = H9 + 9 * (size_t)p;
          float* d = dTr3 + 3 * (size_t)p;
          d[0] = a[0]; d[1] = a[1]; d[2] = a[2];
          h[0] = a[3]; h[1] = a[6]; h[2] = a[7];
          h[3] = a[6]; h[4] = a[4]; h[5] = a[8];
          h[6] = a[7]; h[7] = a[8]; h[8] = a[5];
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Off-path OccGridMapUtil members (HSL/map/OccGridMapUtil.h:106-285) and the occupancy ray query of hector_map_tools.

// OccGridMapUtil::interpMapValue (OccGridMapUtil.h:241-285): bilinear probability, 0 outside the map limits.
__device__ __forceinline__ float hs_interp_map_value(const float* __restrict__ val, const HsLevel& L, float qx, float qy) {
  if (!(qx >= 0.0f && qx <= L.limx && qy >= 0.0f && qy <= L.limy)) return 0.0f;
  const int ix = (int)qx, iy = (int)qy;
  const float fx = qx - (float)ix, fy = qy - (float)iy;
  const float* p = val + iy * L.sx + ix;
  const float i0 = hs_probability(__ldg(p)), i1 = hs_probability(__ldg(p + 1));
  const float i2 = hs_probability(__ldg(p + L.sx)), i3 = hs_probability(__ldg(p + L.sx + 1));
  const float x_inv = 1.0f - fx, y_inv = 1.0f - fy;
  return ((i0 * x_inv + i1 * fx) * y_inv) + ((i2 * x_inv + i3 * fx) * fy);
}

// k_residual: getResidualForState / getLikelihoodForState (OccGridMapUtil.h:201-233) for many pose hypotheses against one
// stream's level. One CTA per pose; the threads evaluate 1 - interpMapValue per point into shared memory, then the
// residual is summed in the reference's point order (STRICT: one dependent FADD chain, bitwise the CPU result) or by a
// tree. dynamic smem: n_pts floats.
template <bool STRICT>
__global__ void __launch_bounds__(HS_MATCH_THREADS) k_residual(HsParams P, int stream, int level, const float* __restrict__ poses_map, int n_poses,
                                                               const float2* __restrict__ pts, int n_pts, float* __restrict__ residual_out,
                                                               float* __restrict__ likelihood_out) {
  extern __shared__ float hs_fun[];
  const int k = blockIdx.x;
  if (k >= n_poses) return;
  const HsLevel& L = P.lv[level];
  const float* val = P.val + (size_t)stream * P.cells_per_stream + L.cell_offset;
  const float ex = poses_map[3 * k], ey = poses_map[3 * k + 1], th = poses_map[3 * k + 2];
  const float sn = hs_sinf(th), cs = hs_cosf(th);
  for (int i = threadIdx.x; i < n_pts; i += blockDim.x) {
    const float2 p = pts[i];
    const float qx = (cs * p.x + (-sn) * p.y) + ex;
    const float qy = (sn * p.x + cs * p.y) + ey;
    hs_fun[i] = 1.0f - hs_interp_map_value(val, L, qx, qy);
  }
  __syncthreads();
  if (threadIdx.x < 32) {
    float residual = 0.0f;
    if (STRICT) {
      if (threadIdx.x == 0) {
        int i = 0;
        for (; i + 8 <= n_pts; i += 8) {
          const float t0 = hs_fun[i], t1 = hs_fun[i + 1], t2 = hs_fun[i + 2], t3 = hs_fun[i + 3];
          const float t4 = hs_fun[i + 4], t5 = hs_fun[i + 5], t6 = hs_fun[i + 6], t7 = hs_fun[i + 7];
          residual += t0; residual += t1; residual += t2; residual += t3; residual += t4; residual += t5; residual += t6; residual += t7;
        }
        for (; i < n_pts; ++i) residual += hs_fun[i];
      }
    } else {
      for (int i = threadIdx.x; i < n_pts; i += 32) residual += hs_fun[i];
#pragma unroll
      for (int m = 16; m >= 1; m >>= 1) residual += __shfl_xor_sync(0xffffffffu, residual, m);
    }
    if (threadIdx.x == 0) {
      residual_out[k] = residual;
      const float sizef = (float)n_pts;                 // getLikelihoodForResidual (OccGridMapUtil.h:208-214)
      likelihood_out[k] = 1.0f - (residual / sizef);
    }
  }
}

// k_interp: interpMapValueWithDerivatives / interpMapValue (OccGridMapUtil.h:241-347) at arbitrary map coordinates of one
// stream's level; out4 [n][4] = (value, d/dx, d/dy, interpMapValue) -- the first three are the reference's Vector3f.
__global__ void __launch_bounds__(128) k_interp(HsParams P, int stream, int level, const float2* __restrict__ coords, int n, float4* __restrict__ out4) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const HsLevel& L = P.lv[level];
  const float* val = P.val + (size_t)stream * P.cells_per_stream + L.cell_offset;
  const float2 c = coords[i];
  float gx, gy, rot, fv;
  HsCellCache none;
  hs_point_terms<false>(val, L, c.x, c.y, 0.0f, 1.0f, 0.0f, 0.0f, none, gx, gy, rot, fv);   // identity rotation: q = (0,0) + (ex,ey)
  const float m = hs_interp_map_value(val, L, c.x, c.y);
  out4[i] = make_float4(m, gx, gy, m);
}

// k_occupancy_ray: HectorMapTools::DistanceMeasurementProvider::checkOccupancyBresenhami + bresenham2D
// (src/hector_slam/hector_map_tools/include/hector_map_tools/HectorMapTools.h:148-232) for many rays on one stream's
// level. The reference walks the published int8 grid, where a cell is 100 exactly when its log-odds are > 0
// (HectorMappingRos.cpp:447-470), so the walk tests the log-odds plane directly. One thread per ray:
// dist_out = (float)(int)|begin - hit| in cells or -1, hit_out = hit cell (untouched on a miss).
__global__ void __launch_bounds__(128) k_occupancy_ray(HsParams P, int stream, int level, const int2* __restrict__ begins, const int2* __restrict__ ends,
                                                       int n, float* __restrict__ dist_out, int2* __restrict__ hit_out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const HsLevel& L = P.lv[level];
  const float* val = P.val + (size_t)stream * P.cells_per_stream + L.cell_offset;
  const int sx = L.sx, sy = L.sy;
  const int x0 = begins[i].x, y0 = begins[i].y, x1 = ends[i].x, y1 = ends[i].y;
  float d = -1.0f;
  if (x0 >= 0 && x0 < sx && y0 >= 0 && y0 < sy && x1 >= 0 && x1 < sx && y1 >= 0 && y1 < sy) {
    const int dx = x1 - x0, dy = y1 - y0;
    const unsigned abs_dx = (unsigned)abs(dx), abs_dy = (unsigned)abs(dy);
    const int off_dx = dx > 0 ? 1 : -1, off_dy = (dy > 0 ? 1 : -1) * sx;
    unsigned abs_da, abs_db;
    int off_a, off_b;
    if (abs_dx >= abs_dy) { abs_da = abs_dx; abs_db = abs_dy; off_a = off_dx; off_b = off_dy; }
    else { abs_da = abs_dy; abs_db = abs_dx; off_a = off_dy; off_b = off_dx; }
    int err = (int)(abs_da / 2);
    int offset = y0 * sx + x0;
    const unsigned end = abs_da < 5000u ? abs_da : 5000u;   // bresenham2D's max_length argument is the constant 5000 (:189,:193)
    for (unsigned t = 0; t < end; ++t) {
      if (__ldg(val + offset) > 0.0f) {     // data[offset] == 100
        const int hx = offset % sx, hy = offset / sx;
        const float fx = (float)(x0 - hx), fy = (float)(y0 - hy);
        d = (float)(int)sqrtf(fx * fx + fy * fy);
        hit_out[i] = make_int2(hx, hy);
        break;
      }
      offset += off_a;
      err += (int)abs_db;
      if ((unsigned)err >= abs_da) { offset += off_b; err -= (int)abs_da; }
    }
  }
  dist_out[i] = d;
}

// ------------------------------------------------------------------------------------------------
// Integration: OccGridMapBase::updateByScan / updateLineBresenhami / bresenham2D / bresenhamCellFree /
// bresenhamCellOcc (HSL/map/OccGridMapBase.h:121-260) + the log-odds updates (GridMapLogOdds.h:135-156).
//
// The reference walks beams sequentially; its result per cell is a function of two facts only:
//   io = smallest beam index whose END POINT is the cell, jf = smallest beam index whose FREE trace crosses it.
//   io exists: idx = markOcc, v' = (jf < io ? (v + f) - f : v), then v' += o if v' < 50.
//   only jf:   idx = markFree, v += f.                      
// Every stored updateIndex is < markFree when a scan starts (currUpdateIndex grows by 3 per scan): a cell whose
// marker is rewritten during a scan needs no look at its old marker, "touched this scan" is the only fact it holds.
//
// One CTA owns one (stream, level) and runs barrier-separated phases over all beams; everything the beams
// share while the scan is being traced lives in SHARED memory, the grids in HBM only see the final,
// coalesced updates:
//   prologue  every global input of the CTA is requested before the gate flag is tested (one memory round trip); the hash
//             table and the cell map are cleared while those loads are in flight
//   A1+A2  per beam: end cell, Bresenham parameters -> beam record; bounding box of the scan; end cell -> shared hash
//          table (cell -> smallest beam index io, via atomicMin)
//   A3  beams that share an end cell have identical traces (same start, same end): only the owner io of each end
//       cell -- the smallest index, which is also the one that decides every "free came first" question -- is put
//       on the walk list (on the coarse levels this removes more than half of the beams), its end-cell bit is set in
//       the cell map, its trace is cut into segments, and the load of its end cell's log-odds is issued (used in C)
//   B1  free cells: the segments are dealt out evenly to the threads, each runs the reference's bresenham2D recurrence
//       (OccGridMapBase.h:243-260) -- but in a 2-bit-per-cell map of the bounding box held in SHARED memory
//       (free bit set with a shared-memory atomicOr, skipped when already set): the scattered cell visits of a
//       scan never reach L2/HBM. A free cell that is some beam's end cell (end bit) looks io up in the hash
//       table and, if this beam < io, sets the entry's "free came first" flag.
//   B2  coalesced sweep of the box, one aligned 4-cell quad per thread and turn: val += f as one 16-byte fp32 reduction
//       performed in L2 (-0.0f for the unmarked cells of the quad), idx = markFree  (exactly once per cell)
//   C   end points: the beam that owns io applies undo-free / occupied to val and stores markOcc.
// A bounding box with more cells than the map holds (HS_MAP_CELLS) is processed in horizontal bands of rows: a trace is
// monotone in y, so its steps inside a band are one interval with a closed form; segments, B1 and B2 run once per band.
// No scratch value is ever stored in the grids. A beam's own trace never contains its own end point (bresenham2D visits
// abs_da cells from the start).

#define HS_HASH_EMPTY (-1)
#ifndef HS_MAP_BYTES
#define HS_MAP_BYTES 32768                 // shared-memory cell map: 2 bits per cell, 16 cells per 32-bit word
#endif                                     // (low half-word: crossed by a free trace; high half-word: some beam's end cell)
#ifndef HS_RAYCAST_THREADS
#define HS_RAYCAST_THREADS 384
#endif
#ifndef HS_RAYCAST_CTAS
#define HS_RAYCAST_CTAS 4                  // CTAs per SM the default instantiation is compiled for (register cap 65536 / (CTAS * THREADS))
#endif
// long scans (hash table + beam records of ~80 KB: one CTA per SM anyway): the CTA takes the whole SM
#define HS_MAP_BYTES_BIG 147456
#define HS_RAYCAST_THREADS_BIG 1024

struct HsBeam {
  int valid;            // begin != end and both inside the grid
  int ex, ey;           // end cell
  int kend;             // end cell offset
  unsigned abs_da, abs_db;
  int a_is_x, sa, sb;   // dominant axis, step signs along the dominant / minor axis
};

// updateByScan's per-beam arithmetic (OccGridMapBase.h:145-158) + updateLineBresenhami's set-up (:170-206)
__device__ __forceinline__ HsBeam hs_make_beam(const HsLevel& L, int bx, int by, float mpx, float mpy, float sn, float cs,
                                              float px, float py) {
  HsBeam b;
  float exf = ((cs * px + (-sn) * py) + mpx) + 0.5f;
  float eyf = ((sn * px + cs * py) + mpy) + 0.5f;
  int ex = (int)exf, ey = (int)eyf;  // truncation toward zero, like cast<int>() (OccGridMapBase.h:152-155)
  b.ex = ex;
  b.ey = ey;
  b.valid = (bx != ex || by != ey) && bx >= 0 && bx < L.sx && by >= 0 && by < L.sy && ex >= 0 && ex < L.sx && ey >= 0 && ey < L.sy;
  int dx = ex - bx, dy = ey - by;
  unsigned abs_dx = (unsigned)abs(dx), abs_dy = (unsigned)abs(dy);
  int sdx = dx > 0 ? 1 : -1;             // util::sign: sign(0) = -1 (UtilFunctions.h:56-59)
  int sdy = dy > 0 ? 1 : -1;
  b.kend = ey * L.sx + ex;
  if (abs_dx >= abs_dy) { b.abs_da = abs_dx; b.abs_db = abs_dy; b.a_is_x = 1; b.sa = sdx; b.sb = sdy; }
  else { b.abs_da = abs_dy; b.abs_db = abs_dx; b.a_is_x = 0; b.sa = sdy; b.sb = sdx; }
  return b;
}

// Beam record kept in shared memory between the phases (one per scan point), 12 bytes.
struct HsBeamRec {
  int kend;                    // end cell offset in the grid
  unsigned short da, db;       // abs_da (0 marks a dropped beam), abs_db
  unsigned int flags;          // bit0: dominant axis is x, bit1: step along a is +, bit2: step along b is +, bits 3..: end row
};

// Walk record: one per end-cell owner beam, everything the free-cell walk needs in one 16-byte shared load.
struct __align__(16) HsWalkRec {
  unsigned dadb;               // abs_da | abs_db << 16
  unsigned offs;               // cell-address step along the dominant axis (low 16 bits, signed) | along the minor axis << 16
  int beam;                    // beam index (the reference's loop position; decides "free came first") | x-dominant << 30 | y grows << 31
  int seg_start;               // index of this beam's first walk segment in the current band (ascending over the walk list)
};

// exact x / d for 0 <= x < 2^22, 0 < d < 2^22: approximate float quotient (MUFU.RCP, <= 2 ulp) + one-step correction
__device__ __forceinline__ int hs_udiv_small(int x, int d, int& rem) {
  int q = (int)__fdividef((float)x, (float)d);
  int r = x - q * d;
  if (r < 0) { --q; r += d; }
  else if (r >= d) { ++q; r -= d; }
  rem = r;
  return q;
}

// 32-bit shared-window accesses for the walk loop (keeps the address arithmetic out of the generic space)
__device__ __forceinline__ unsigned hs_lds_u32(unsigned addr) {
  unsigned v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ void hs_reds_or(unsigned addr, unsigned v) {
  asm volatile("red.shared.or.b32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
// explicit global-space stores (for pointers that were made opaque to the compiler, which would otherwise use generic stores)
__device__ __forceinline__ void hs_stg_u8(uint8_t* p, unsigned v) { asm volatile("st.global.u8 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ void hs_stg_u32(uint8_t* p, unsigned v) { asm volatile("st.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ float4 hs_lds_f4(unsigned addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ unsigned hs_atoms_or(unsigned addr, unsigned v) {
  unsigned old;
  asm volatile("atom.shared.or.b32 %0, [%1], %2;" : "=r"(old) : "r"(addr), "r"(v) : "memory");
  return old;
}
// shared-window address of a generic pointer, computed ONCE (volatile: the compiler may not re-derive it per use)
__device__ __forceinline__ unsigned hs_shared_addr(const void* p) {
  unsigned a;
  asm volatile("{ .reg .u64 t; cvta.to.shared.u64 t, %1; cvt.u32.u64 %0, t; }" : "=r"(a) : "l"(p));
  return a;
}

__device__ __forceinline__ void hs_stamp(const HsParams& P, int k) {
#ifdef HS_EXPERIMENTS
  if (P.dbg_times && threadIdx.x == 0) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    P.dbg_times[(size_t)blockIdx.x * 8 + k] = t;
  }
#else
  (void)P; (void)k;
#endif
}

// End-cell hash table in shared memory: open addressing, linear probing, size = power of two >= 2 * max_points.
__device__ __forceinline__ int hs_hash_slot(int cell, int log2_size) { return (int)(((unsigned)cell * 2654435761u) >> (32 - log2_size)); }

// returns the slot of `cell`, or -1 if the cell is not an end cell
__device__ __forceinline__ int hs_hash_find(const int* __restrict__ keys, int mask, int log2_size, int cell) {
  int h = hs_hash_slot(cell, log2_size);
  for (;;) {
    int k = keys[h];
    if (k == cell) return h;
    if (k == HS_HASH_EMPTY) return -1;
    h = (h + 1) & mask;
  }
}

// a free trace reaches a cell that is some beam's end point: only "did a free trace of an EARLIER beam cross it" matters
__device__ __forceinline__ void hs_free_hits_end_cell(const int* __restrict__ hkeys, int* __restrict__ hvals, int hash_mask, int log2_hash,
                                                      int cell, int beam) {
  int h = hs_hash_find(hkeys, hash_mask, log2_hash, cell);
  if (h >= 0) {
    int hv = hvals[h];
    if (beam < (hv >> 1) && !(hv & 1)) atomicOr(&hvals[h], 1);
  }
}

// Folds the byte marker plane of n cells into the int32 plane (b != 0: idx = base + b; b = 0) -- all threads of the CTA.
// 16 cells per 16-byte load where the slab is aligned; almost every word is zero except around the robot.
__device__ __forceinline__ void hs_fold_markers(uint8_t* __restrict__ mk, int* __restrict__ idx, size_t n, int base, int tid, int T) {
  if ((reinterpret_cast<uintptr_t>(mk) & 15) == 0) {
    uint4* m4 = reinterpret_cast<uint4*>(mk);
    const size_t n16 = n >> 4;
    for (size_t i = tid; i < n16; i += T) {
      const uint4 w = m4[i];
      if (w.x | w.y | w.z | w.w) {
        const unsigned ww[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
        for (int q = 0; q < 4; ++q)
#pragma unroll
          for (int b = 0; b < 4; ++b) {
            const unsigned v = (ww[q] >> (8 * b)) & 0xFFu;
            if (v) idx[(i << 4) + q * 4 + b] = base + (int)v;
          }
        m4[i] = make_uint4(0u, 0u, 0u, 0u);
      }
    }
    for (size_t i = (n16 << 4) + tid; i < n; i += T) {
      const unsigned v = mk[i];
      if (v) { idx[i] = base + (int)v; mk[i] = 0; }
    }
  } else {
    for (size_t i = tid; i < n; i += T) {
      const unsigned v = mk[i];
      if (v) { idx[i] = base + (int)v; mk[i] = 0; }
    }
  }
}

// k_fold_markers: the same for every level of one stream (before the int32 plane is copied or handed out as it is).
// One CTA per level; the bases stay as they are (with all bytes zero they mean nothing).
__global__ void __launch_bounds__(512) k_fold_markers(HsParams P, int stream) {
  const int level = blockIdx.x;
  if (level >= P.n_levels) return;
  const HsLevel& L = P.lv[level];
  const size_t off = (size_t)stream * P.cells_per_stream + L.cell_offset;
  hs_fold_markers(P.mk + off, P.idx + off, (size_t)L.sx * L.sy, P.mk_base[(size_t)stream * HS_MAX_LEVELS + level], threadIdx.x, blockDim.x);
}

// dynamic shared memory: HS_MAP_BYTES (cell map) | hash keys[T] | hash vals[T] | max_points * sizeof(HsWalkRec)
// | max_points * sizeof(HsBeamRec) | 2 * max_points * sizeof(unsigned short), T = hash_size (power of two >= 2 * max_points).
// log2_seg_levels: the free-cell walk is cut into segments of 2^k steps that are dealt out evenly to the threads
// (k = low byte on level 0, next byte on the coarser levels).
template <int THREADS, int MAP_BYTES, int MIN_CTAS>
__global__ void __launch_bounds__(THREADS, MIN_CTAS) k_raycast(HsParams P, int n_slots, const int32_t* __restrict__ stream_ids,
                                                               const float2* __restrict__ points, const int32_t* __restrict__ counts,
                                                               const float2* __restrict__ origos, int log2_hash, int log2_seg_levels,
                                                               int level_first, int level_count) {
  constexpr int MAP_CELLS = MAP_BYTES * 4;
  extern __shared__ __align__(16) unsigned char hs_smem[];
  const int hash_size = 1 << log2_hash, hash_mask = hash_size - 1;
  unsigned* cmap = reinterpret_cast<unsigned*>(hs_smem);                   // 16 cells per word: free bits | end bits << 16
  int* hkeys = reinterpret_cast<int*>(hs_smem + MAP_BYTES);             // end cell (grid offset) or EMPTY
  int* hvals = hkeys + hash_size;                                          // (io << 1) | free-came-first flag
  HsWalkRec* walk = reinterpret_cast<HsWalkRec*>(hvals + hash_size);       // owner beams
  HsBeamRec* beams = reinterpret_cast<HsBeamRec*>(walk + P.max_points);
  unsigned short* tlo = reinterpret_cast<unsigned short*>(beams + P.max_points);   // per owner: first step inside the current band
  unsigned short* tlen = tlo + P.max_points;                                        // ... and the number of steps inside it
  __shared__ int s_box[4];  // xmin, ymin, xmax, ymax over the start cell and all valid end cells
  __shared__ float4 s_addlut[16];   // sweep: the four addends of a quad by its 4 map bits
  __shared__ unsigned long long s_alloc;   // walk-list allocator: owners << 32 | segments (single band)
  __shared__ int s_items;
  __shared__ int s_wsum[32];
  __shared__ unsigned s_visits;
  __shared__ unsigned s_inv_pitch;   // floor((2^32 - 1) / pitch) of the cell map

  // slot-major launch order: the long fine-grid CTAs and the short coarse-grid CTAs of a scan are dispatched together, which
  // keeps the mix of shared-memory-bound (walk) and HBM-bound (sweep) work on every SM even (measured: level-major order
  // is 20 % slower)
  int level, slot;
  {
    // ... except for the last ~wave of slots, which is dispatched level-major (fine grids first): the launch then drains
    // through the short coarse-grid CTAs instead of waiting for a few long fine-grid ones
    // this launch covers levels [level_first, level_first + level_count) of every slot
    const int n_tail = min(n_slots, P.tail_slots), n_head = n_slots - n_tail;
    int local;
    // (hs_udiv_small: the generic 32-bit division is ~35 instructions, and every thread of the launch runs this)
    if ((int)blockIdx.x < n_head * level_count) {
      const unsigned b = blockIdx.x;   // level_count <= HS_MAX_LEVELS: divisions by constants (multiply-high)
      switch (level_count) {
        case 1: local = (int)b; break;
        case 2: local = (int)(b >> 1); break;
        case 3: local = (int)(b / 3u); break;
        case 4: local = (int)(b >> 2); break;
        default: local = (int)(b / (unsigned)level_count); break;
      }
      level = (int)b - local * level_count;
    } else {
      int rem;
      level = hs_udiv_small((int)blockIdx.x - n_head * level_count, n_tail, rem);
      local = n_head + rem;
    }
    if (local >= n_slots || level >= level_count) return;
    level += level_first;
    // consecutive launches walk the slots in opposite directions: the grids of the streams integrated LAST by one launch
    // are still in L2 when the next launch starts with them
    if (P.reverse_order) local = n_slots - 1 - local;
    slot = P.slot0 + local;
  }
  hs_stamp(P, 0);
  // Everything this CTA needs from global memory is requested up front, BEFORE the gate flag is tested: the loads are
  // independent, so the prologue costs one memory round trip instead of a chain of three (flag -> parameters -> points).
  const int integrate = P.slot_integrate[slot];
  const int s = hs_slot_stream(stream_ids, slot);
  const HsLevel& L = P.lv[level];
  const int s_map = HS_DBG(P, 256) ? (s & 7) : s;   // timing experiment (wrong results): all CTAs share 8 streams' grids, which fit in L2
  float* val = P.val + (size_t)s_map * P.cells_per_stream + L.cell_offset;
  int* idx = P.idx + (size_t)s_map * P.cells_per_stream + L.cell_offset;
  uint8_t* mk = P.mk + (size_t)s_map * P.cells_per_stream + L.cell_offset;
  const int mk_base_old = P.mk_base[(size_t)s * HS_MAX_LEVELS + level];
  const int W = L.sx;
  constexpr int T = THREADS;   // == blockDim.x: divisions and strides by T compile to multiplications / immediates

  // level 0 integrates the scan itself, level i > 0 the container last refreshed by matchData
  // (MapRepMultiMap.h:134-147) -- possibly stale or empty when map_without_matching was used.
  const float2* pts;
  int n;
  float ox, oy;
  if (level == 0) {
    pts = points + (size_t)slot * P.max_points;
    n = counts[slot];
    float2 o = origos ? origos[slot] : make_float2(0.f, 0.f);
    ox = o.x; oy = o.y;
  } else {
    pts = reinterpret_cast<const float2*>(P.coarse_pts) + (size_t)s * P.max_points;
    n = P.coarse_cnt[s];
    ox = P.coarse_origo[2 * s] * L.pt_factor;
    oy = P.coarse_origo[2 * s + 1] * L.pt_factor;
  }
  const float2 p_first = (int)threadIdx.x < P.max_points ? pts[threadIdx.x] : make_float2(0.f, 0.f);   // this thread's first beam
  const int mark_base = P.slot_mark_base[slot];
  const float pose_x = P.slot_pose[3 * slot], pose_y = P.slot_pose[3 * slot + 1];
  const float sn = P.slot_sc[2 * slot], cs = P.slot_sc[2 * slot + 1];   // sinf / cosf of the pose angle, from the gate
  if (!integrate) return;
  n = n < 0 ? 0 : (n > P.max_points ? P.max_points : n);
  const int mark_free = mark_base + 1, mark_occ = mark_base + 2;
  const float f = P.log_odds_free, o = P.log_odds_occ;
  // markers of this scan as bytes relative to the (stream, level)'s base; if they do not fit (every 84 integrated scans, or
  // after the counters were set from outside), fold the byte plane into the int32 plane first and start a new base
  int mk_base = mk_base_old;
  if ((unsigned)(mark_free - mk_base - 1) > 253u) {   // CTA-uniform
    hs_fold_markers(mk, idx, (size_t)L.sx * L.sy, mk_base_old, threadIdx.x, THREADS);
    mk_base = mark_base;
    __syncthreads();
    if (threadIdx.x == 0) P.mk_base[(size_t)s * HS_MAX_LEVELS + level] = mk_base;
  }
  const unsigned b_free = (unsigned)(mark_free - mk_base), b_occ = (unsigned)(mark_occ - mk_base);   // 1 .. 255

  float mpx, mpy;
  hs_world_to_map(L, pose_x, pose_y, mpx, mpy);
  const int bx = (int)(((cs * ox + (-sn) * oy) + mpx) + 0.5f);  // Vector2i(float, float): truncation (OccGridMapBase.h:137)
  const int by = (int)(((sn * ox + cs * oy) + mpy) + 0.5f);
  const float pf = L.pt_factor;
  const int lane = threadIdx.x & 31;

  // clear the hash table, reset the box and the allocators
  {
    int4* k4 = reinterpret_cast<int4*>(hkeys);
    int4* v4 = reinterpret_cast<int4*>(hvals);
    for (int i = threadIdx.x; i < (hash_size >> 2); i += T) {
      k4[i] = make_int4(HS_HASH_EMPTY, HS_HASH_EMPTY, HS_HASH_EMPTY, HS_HASH_EMPTY);
      v4[i] = make_int4(INT_MAX, INT_MAX, INT_MAX, INT_MAX);
    }
  }
  {   // ... and the whole cell map (8 16-byte stores per thread, while the prologue's global loads are in flight): the box
      // is not known yet, and phase A3 sets end-cell bits right after the next barrier
    uint4* m4 = reinterpret_cast<uint4*>(hs_smem);
    for (int i = threadIdx.x; i < MAP_BYTES / 16; i += T) m4[i] = make_uint4(0u, 0u, 0u, 0u);
  }
  if (threadIdx.x == 0) { s_box[0] = INT_MAX; s_box[1] = INT_MAX; s_box[2] = INT_MIN; s_box[3] = INT_MIN; s_alloc = 0ULL; s_visits = 0u; }
  if (threadIdx.x < 16) {
    const unsigned b = threadIdx.x;
    s_addlut[b] = make_float4((b & 1u) ? f : -0.0f, (b & 2u) ? f : -0.0f, (b & 4u) ? f : -0.0f, (b & 8u) ? f : -0.0f);
  }
  __syncthreads();

  // phases A1 + A2: one thread per beam -- beam record, bounding box, and end point ownership (io = smallest beam index
  // per end cell) through the shared hash table
  {
    unsigned visits = 0;
    int xmin = INT_MAX, ymin = INT_MAX, xmax = INT_MIN, ymax = INT_MIN;
    for (int i = threadIdx.x; i < n; i += T) {
      const float2 p = (i == (int)threadIdx.x) ? p_first : pts[i];
      HsBeam b = hs_make_beam(L, bx, by, mpx, mpy, sn, cs, p.x * pf, p.y * pf);
      HsBeamRec r;
      r.kend = b.kend; r.da = 0; r.db = (unsigned short)b.abs_db;
      r.flags = (unsigned)((b.a_is_x ? 1 : 0) | (b.sa > 0 ? 2 : 0) | (b.sb > 0 ? 4 : 0)) | ((unsigned)b.ey << 3);   // + the end row: no division by W later
      if (b.valid) {
        visits += b.abs_da + 1u;
        r.da = (unsigned short)b.abs_da;
        xmin = min(xmin, b.ex); xmax = max(xmax, b.ex); ymin = min(ymin, b.ey); ymax = max(ymax, b.ey);
        int h = hs_hash_slot(b.kend, log2_hash);
        for (;;) {
          int old = atomicCAS(&hkeys[h], HS_HASH_EMPTY, b.kend);
          if (old == HS_HASH_EMPTY || old == b.kend) break;
          h = (h + 1) & hash_mask;
        }
        atomicMin(&hvals[h], i << 1);
      }
      beams[i] = r;
    }
    visits = __reduce_add_sync(0xffffffffu, visits);
    xmin = __reduce_min_sync(0xffffffffu, xmin); ymin = __reduce_min_sync(0xffffffffu, ymin);
    xmax = __reduce_max_sync(0xffffffffu, xmax); ymax = __reduce_max_sync(0xffffffffu, ymax);
    if (lane == 0 && visits) {
      atomicAdd(&s_visits, visits);
      atomicMin(&s_box[0], min(xmin, bx)); atomicMin(&s_box[1], min(ymin, by));
      atomicMax(&s_box[2], max(xmax, bx)); atomicMax(&s_box[3], max(ymax, by));
    }
  }
  __syncthreads();
  if (s_box[2] < s_box[0]) return;  // no valid beam (uniform for the CTA)
  if (threadIdx.x == 0) atomicAdd(&P.counters[3], (unsigned long long)s_visits);

  // geometry of the cell map: rows of `pitch` cells starting at a 32-aligned x0; the row pitch in words is odd so
  // that a vertical run of cells spreads over the shared-memory banks. A box with more cells than the map holds is
  // processed in horizontal bands of `band_rows` rows, one after the other (hi-res grids, long ranges).
  const int x0 = s_box[0] & ~31, y0 = s_box[1], x1 = s_box[2], y1 = s_box[3];
  const int bh = y1 - y0 + 1;
  const int pw = (((x1 - x0) >> 4) + 1) | 1;   // 16-cell words per map row
  const int pitch = pw << 4;                   // cells per map row
  int band_rows = bh, n_bands = 1;
  if (bh * pitch > MAP_CELLS) {   // (rare for 360-beam scans: the divisions are skipped)
    int div_rem;
    band_rows = min(bh, hs_udiv_small(MAP_CELLS, pitch, div_rem));   // >= 1: hs_create limits the grid width
    n_bands = hs_udiv_small(bh + band_rows - 1, band_rows, div_rem);
  }

  if (threadIdx.x == 0) s_inv_pitch = 0xFFFFFFFFu / (unsigned)pitch;   // read after the barrier behind phase A3
  const bool single_band = n_bands == 1;
  const int log2_seg = level == 0 ? (log2_seg_levels & 0xFF) : (log2_seg_levels >> 8);   // per level: fine grid / coarse grids
  const int seg = 1 << log2_seg;
  hs_stamp(P, 1);

  // phase A3: walk list = the owner beam of every end cell. Beams that share an end cell have identical traces (same
  // start, same end), and the owner -- the smallest index -- is the one that decides every "free came first" question.
  float v_end = 0.0f;   // log-odds of the end cell owned by beam threadIdx.x (first pass of A3)
  int my_pos = 0, my_seg0 = 0, my_nseg = 0;   // this thread's owner beam of the first pass: list position, first segment, segments
  for (int i0 = 0; i0 < n; i0 += T) {
    const int i = i0 + threadIdx.x;
    bool owner = false;
    HsBeamRec r;
    r.kend = 0; r.da = 0; r.db = 0; r.flags = 0;
    if (i < n) {
      r = beams[i];
      if (r.da != 0) {
        int h = hs_hash_find(hkeys, hash_mask, log2_hash, r.kend);
        owner = (hvals[h] >> 1) == i;
      }
    }
    // one 64-bit shared atomicAdd per warp hands out list positions and (single band) segment indices together, so
    // seg_start ascends along the list
    const int nseg = (owner && single_band) ? ((int)r.da + seg - 1) >> log2_seg : 0;
    int incl = nseg;
    if (single_band) {
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += t;
      }
    }
    const unsigned m = __ballot_sync(0xffffffffu, owner);
    unsigned long long base = 0ULL;
    if (lane == 31 && m) base = atomicAdd(&s_alloc, ((unsigned long long)__popc(m) << 32) | (unsigned long long)(unsigned)incl);
    base = __shfl_sync(0xffffffffu, base, 31);
    if (owner) {
      const int sa = (r.flags & 2) ? 1 : -1, sb = (r.flags & 4) ? 1 : -1;
      const int off_a = (r.flags & 1) ? sa : sa * pitch;
      const int off_b = (r.flags & 1) ? sb * pitch : sb;
      HsWalkRec w;
      w.dadb = (unsigned)r.da | ((unsigned)r.db << 16);
      w.offs = ((unsigned)off_a & 0xFFFFu) | ((unsigned)off_b << 16);
      const unsigned xdom = r.flags & 1u;   // y grows along the trace: sign of the minor step (x-dominant) or of the major step
      const unsigned y_up = xdom ? ((r.flags >> 2) & 1u) : ((r.flags >> 1) & 1u);
      w.beam = (int)((unsigned)i | (xdom << 30) | (y_up << 31));
      w.seg_start = (int)(unsigned)(base & 0xFFFFFFFFULL) + incl - nseg;
      const int pos = (int)(base >> 32) + __popc(m & ((1u << lane) - 1u));
      walk[pos] = w;
      if (i0 == 0) { my_pos = pos; my_seg0 = w.seg_start; my_nseg = nseg; }
      // phase C needs this end cell's log-odds. Nothing else touches an end cell during the scan (the sweep only adds -0.0f
      // to it), so the load is issued now and consumed after the walk and the sweep: its latency never shows.
      if (i0 == 0) v_end = __ldcg(val + r.kend);
      else asm volatile("prefetch.global.L2 [%0];" ::"l"(val + r.kend));
      if (single_band) {   // the whole trace lies in the one band; mark the end cell in the (already cleared) map
        const int ey = (int)(r.flags >> 3), ex = r.kend - ey * W;
        const int kb = (ey - y0) * pitch + (ex - x0);
        atomicOr(&cmap[kb >> 4], 0x10000u << (kb & 15));
      }
    }
  }
  __syncthreads();
  const int n_walk = (int)(s_alloc >> 32);
  // Start table (single band, one pass of A3): walk slot u begins at segment u * per, and the owner beam that holds that segment
  // leaves its list position there -- instead of every thread binary-searching the walk list for it (5 % of the kernel's
  // instructions). tlo / tlen (contiguous, 2 * max_points entries) are not used by a single-band scan.
  unsigned short* wstart = tlo;
  const bool fast_start = single_band && n <= T && 2 * P.max_points >= T;
  if (fast_start) {
    const int per0 = ((int)(unsigned)(s_alloc & 0xFFFFFFFFULL) + T - 1) / T;
    if (my_nseg > 0) {
      int rem;
      int u = hs_udiv_small(my_seg0 + per0 - 1, per0, rem);   // first slot that starts at or behind this beam's first segment
      for (int js = u * per0; js < my_seg0 + my_nseg; js += per0, ++u) wstart[u] = (unsigned short)my_pos;
    }
    __syncthreads();
  }
  hs_stamp(P, 2);

  const int warp = threadIdx.x >> 5, n_warps = T >> 5;
  const bool vec_ok = ((W & 3) == 0) && ((reinterpret_cast<uintptr_t>(val) & 15) == 0);
  const unsigned cmap_s = hs_shared_addr(cmap);

  auto run_bands = [&](auto single_c) {
  constexpr bool SINGLE = decltype(single_c)::value;
  for (int band = 0; band < (SINGLE ? 1 : n_bands); ++band) {
    const int yb0 = SINGLE ? y0 : y0 + band * band_rows;
    const int rows = SINGLE ? bh : min(band_rows, y1 - yb0 + 1);
    const int yb1 = yb0 + rows - 1;
    if (!SINGLE) {   // clear the part of the cell map this band uses
      uint4* m4 = reinterpret_cast<uint4*>(hs_smem);
      const int n4 = (pw * rows + 3) >> 2;
      for (int i = threadIdx.x; i < n4; i += T) m4[i] = make_uint4(0u, 0u, 0u, 0u);
      if (threadIdx.x == 0) s_items = 0;
    }
    if (!SINGLE) __syncthreads();

    // per owner: end bit of its end cell, the steps [t_lo, t_lo + len) of its trace that lie in this band (a trace is
    // monotone in y), cut into segments of `seg` steps; seg_start = exclusive prefix sum over the walk list
    for (int w0 = 0; w0 < (SINGLE ? 0 : n_walk); w0 += T) {
      const int wi = w0 + threadIdx.x;
      int nseg = 0;
      if (wi < n_walk) {
        const HsWalkRec w = walk[wi];
        const int da = (int)(w.dadb & 0xFFFFu), db = (int)(w.dadb >> 16);
        const HsBeamRec br = beams[w.beam & 0x3FFFFFFF];
        const int ey = (int)(br.flags >> 3), ex = br.kend - ey * W;
        if (ey >= yb0 && ey <= yb1) {
          const int kb = (ey - yb0) * pitch + (ex - x0);
          atomicOr(&cmap[kb >> 4], 0x10000u << (kb & 15));
        }
        int t_lo = 0, t_hi = da - 1;
        {
          const bool y_up = (w.beam >> 31) & 1;
          // rows of the band in units of "y steps from the start cell along the trace"
          int r_lo = y_up ? yb0 - by : by - yb1, r_hi = y_up ? yb1 - by : by - yb0;
          if (w.beam & 0x40000000) {   // x-dominant: row(t) = floor((da/2 + t*db) / da)
            const int half = da >> 1;
            const int r_max = (half + (da - 1) * db) / da;
            if (r_lo > r_max || r_hi < 0) { t_hi = -1; }
            else {
              if (r_lo > 0) t_lo = (r_lo * da - half + db - 1) / db;                 // db > 0 here (r_max >= r_lo > 0)
              if (r_hi < r_max) t_hi = ((r_hi + 1) * da - half + db - 1) / db - 1;
            }
          } else {                     // y-dominant: row(t) = t
            t_lo = max(r_lo, 0);
            t_hi = min(r_hi, da - 1);
          }
        }
        const int len = t_hi >= t_lo ? t_hi - t_lo + 1 : 0;
        tlo[wi] = (unsigned short)t_lo;
        tlen[wi] = (unsigned short)len;
        nseg = (len + seg - 1) >> log2_seg;
      }
      int incl = nseg;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += t;
      }
      if (lane == 31) s_wsum[warp] = incl;
      __syncthreads();
      int off = s_items;
      for (int k = 0; k < warp; ++k) off += s_wsum[k];
      if (wi < n_walk) walk[wi].seg_start = off + incl - nseg;
      __syncthreads();
      if (threadIdx.x == T - 1) s_items = off + incl;
    }
    if (!SINGLE) __syncthreads();
    const int n_items = SINGLE ? (int)(unsigned)(s_alloc & 0xFFFFFFFFULL) : s_items;

    // phase B1: free cells with the reference's bresenham2D recurrence (OccGridMapBase.h:243-260). Every thread takes
    // the same number of consecutive walk segments (so the longest beam does not set the CTA's pace); a thread that
    // starts inside a beam enters the recurrence through its closed form after t steps:
    //   cell(t) = k0 + t*off_a + floor((abs_da/2 + t*abs_db) / abs_da) * off_b,  err(t) = (abs_da/2 + t*abs_db) mod abs_da.
    // Marks go to the 2-bit-per-cell map of the band in SHARED memory: the scattered cell visits of a scan never reach
    // L2/HBM.
    if (!HS_DBG(P, 1)) {
      const int k0 = (by - yb0) * pitch + (bx - x0);   // map address of the start cell (may lie outside this band)
      const int per = (n_items + T - 1) / T;
      // neighbouring lanes take segments ~37 beams apart: lanes that walk adjacent beams at the same radius would hit
      // the same 16-cell map words in the same instruction and serialise in the shared-memory atomics
      const int uslot = (int)((threadIdx.x * 37u) % (unsigned)T);
      int j = uslot * per;
      const int j1 = min(j + per, n_items);
      if (j < j1) {
        int lo = 0, hi = n_walk - 1;       // largest p with walk[p].seg_start <= j
        if (SINGLE && fast_start) {
          lo = (int)wstart[uslot];
        } else {
          while (lo < hi) {
            const int mid = (lo + hi + 1) >> 1;
            if (walk[mid].seg_start <= j) lo = mid; else hi = mid - 1;
          }
        }
        int p = lo;
        HsWalkRec w = walk[p];
        int da = (int)(w.dadb & 0xFFFFu), db = (int)(w.dadb >> 16);
        int off_a = (int)(short)(w.offs & 0xFFFFu), off_b = ((int)w.offs) >> 16;
        int beam = w.beam & 0x3FFFFFFF;
        const int s_off = (j - w.seg_start) << log2_seg;
        int t0 = (SINGLE ? 0 : (int)tlo[p]) + s_off;
        int rem = (SINGLE ? da : (int)tlen[p]) - s_off;
        int e0 = (da >> 1) + t0 * db;
        int err;
        int q0 = e0 < (1 << 22) ? hs_udiv_small(e0, da, err) : e0 / da;   // (the generic division: hi-res grids with long, steep traces only)
        int k = k0 + t0 * off_a + q0 * off_b;
        err = e0 - q0 * da;
        for (; j < j1; ++j) {
          const int steps = min(seg, rem);
          for (int t = 0; t < steps; ++t) {
            // ONE shared-memory atomic per step: it sets the free bit unconditionally and returns the word, whose end bit says
            // whether the cell is some beam's end point. (The load / test / conditional-reduction form cost 8 more instructions
            // per step, and near the sensor, where most cells are already marked by other beams, nearly every warp iteration ran
            // both sides of the test: 27 lanes -> 16 + 11, ncu.) A free bit that lands on an end cell is masked out by the sweep.
            const unsigned waddr = cmap_s + (((unsigned)k >> 4) << 2);
            unsigned fbit;   // 1 << (k & 15) as a one-bit mask (BMSK: no register for the constant 1)
            asm("bmsk.clamp.b32 %0, %1, 1;" : "=r"(fbit) : "r"((unsigned)k & 15u));
            const unsigned word = hs_atoms_or(waddr, fbit);
            if (word & (fbit << 16)) {   // the free trace crosses a cell that is some beam's end point (one warp step in five)
              // k -> (row, column) of the map: multiplication by floor(2^32 / pitch) and one correction (k < 2^20: the product's
              // error is below one)
              int yy = (int)__umulhi((unsigned)k, s_inv_pitch);
              int xx = k - yy * pitch;
              if (xx >= pitch) { ++yy; xx -= pitch; }
              hs_free_hits_end_cell(hkeys, hvals, hash_mask, log2_hash, (yy + yb0) * W + (xx + x0), beam);
            }
            // bresenham2D's update (OccGridMapBase.h:252-258) as predicated adds: 5 instructions, 2 of them on the alu pipe --
            // the compiler makes add / compare / select / select / subtract / 3-input add (4 alu) of the C form, and the alu pipe
            // (one warp instruction per 2 cycles per scheduler) is the walk's busiest: 160.9 -> 159.5 us per launch, same box
            asm volatile("{ .reg .pred p;\n\t"
                         "add.s32 %1, %1, %2;\n\t"       // err += db
                         "add.s32 %0, %0, %4;\n\t"       // k += off_a
                         "setp.ge.s32 p, %1, %3;\n\t"
                         "@p sub.s32 %1, %1, %3;\n\t"    // err -= da
                         "@p add.s32 %0, %0, %5;\n\t"    // k += off_b
                         "}" : "+r"(k), "+r"(err) : "r"(db), "r"(da), "r"(off_a), "r"(off_b));
          }
          rem -= steps;
          if (rem == 0 && j + 1 < j1) {
            if (SINGLE) ++p; else do { ++p; } while (tlen[p] == 0);
            w = walk[p];
            da = (int)(w.dadb & 0xFFFFu); db = (int)(w.dadb >> 16);
            off_a = (int)(short)(w.offs & 0xFFFFu); off_b = ((int)w.offs) >> 16;
            beam = w.beam & 0x3FFFFFFF;
            if (SINGLE) {
              k = k0; err = da >> 1; rem = da;
            } else {
              t0 = (int)tlo[p]; rem = (int)tlen[p];
              e0 = (da >> 1) + t0 * db;
              q0 = t0 ? e0 / da : 0;
              k = k0 + t0 * off_a + q0 * off_b; err = e0 - q0 * da;
            }
          }
        }
      }
    }
    __syncthreads();
    if (band == 0) hs_stamp(P, 3);

    // phase B2: updateSetFree exactly once per marked cell -- coalesced sweep of the band.
    if (HS_DBG(P, 2)) {
    } else if (vec_ok) {
      // fast path: the band is a list of aligned 4-cell quads (one float4 of log-odds, 4 map bits), dealt out to the
      // threads in row-major order (a warp's 32 quads are 512 contiguous bytes of a row), four quads in flight per
      // thread. Unmarked components are written back unchanged, hence the barrier before phase C touches the end cells.
      const int qpr = ((x1 - x0) >> 2) + 1;     // quads per box row (x0 is 32-aligned, W a multiple of 4)
      const int nq = qpr * rows;
      const unsigned addlut_s = hs_shared_addr(s_addlut);
      // the two plane pointers are made opaque here: the compiler otherwise re-derives each of them from the kernel parameters
      // in front of every access (9 instructions per non-empty quad iteration)
      float* valq = val;
      uint8_t* mkq = mk;
      asm volatile("" : "+l"(valq), "+l"(mkq));
      const unsigned bf4 = b_free * 0x01010101u;
      // a thread's quads are T apart in the row-major list; its position advances incrementally: c = quad in the row,
      // ni = nibble index into the cell map (row pitch pw*4 nibbles), gq = cell offset of the quad in the grid
      int d_c;
      const int d_ry = hs_udiv_small(T, qpr, d_c);   // one stride of T quads, as (rows, quads)
      const int ni_step = d_ry * (pw << 2) + d_c, gq_step = d_ry * W + (d_c << 2);
      const int ni_wrap = (pw << 2) - qpr, gq_wrap = W - (qpr << 2);
      int c;
      const int ry0 = hs_udiv_small(threadIdx.x, qpr, c);
      int ni = ry0 * (pw << 2) + c, gq = (yb0 + ry0) * W + x0 + (c << 2);
      for (int q0 = threadIdx.x; q0 < nq; q0 += T) {
        const unsigned mw = cmap[ni >> 2];
        const unsigned bits = ((mw & ~(mw >> 16)) >> ((unsigned)(ni & 3) << 2)) & 0xFu;   // free and not an end cell
        if (bits) {
          // val += f on the marked cells of the quad as ONE 16-byte floating-point reduction performed in L2 (IEEE round to
          // nearest, like the FADD of updateSetFree; nothing is loaded into or written back from registers). Unmarked cells
          // get -0.0f added, which leaves every value -- including +-0.0 -- bit for bit as it was.
          const float4 a = hs_lds_f4(addlut_s + (bits << 4));   // (bit i ? f : -0.0f) for the quad's four cells: one 16-byte shared load instead of 8 selects
          asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(valq + gq), "f"(a.x), "f"(a.y), "f"(a.z), "f"(a.w) : "memory");
          if (HS_DBG(P, 8)) {   // timing experiment (wrong markers): no free-marker stores
          } else {
            // the quad's marker address is formed ONCE (left to itself the compiler re-derives the 64-bit sum in front of every
            // predicated byte store: 3 instructions per byte, 12 % of the kernel's instructions), the byte stores of a partly
            // marked quad are four predicated stores at immediate offsets
            uint8_t* mq = mkq + gq;
            asm volatile("" : "+l"(mq));
            if (bits == 0xFu || HS_DBG(P, 64)) {   // bit 6, timing experiment (wrong markers): every quad as one store
              hs_stg_u32(mq, bf4);
            } else {
              asm volatile("{ .reg .pred p0, p1, p2, p3; .reg .b32 t;\n\t"
                           "and.b32 t, %1, 1; setp.ne.u32 p0, t, 0;\n\t"
                           "and.b32 t, %1, 2; setp.ne.u32 p1, t, 0;\n\t"
                           "and.b32 t, %1, 4; setp.ne.u32 p2, t, 0;\n\t"
                           "and.b32 t, %1, 8; setp.ne.u32 p3, t, 0;\n\t"
                           "@p0 st.global.u8 [%0], %2;\n\t"
                           "@p1 st.global.u8 [%0+1], %2;\n\t"
                           "@p2 st.global.u8 [%0+2], %2;\n\t"
                           "@p3 st.global.u8 [%0+3], %2;\n\t"
                           "}" ::"l"(mq), "r"(bits), "r"(b_free) : "memory");
            }
          }
        }
        c += d_c; ni += ni_step; gq += gq_step;
        if (c >= qpr) { c -= qpr; ni += ni_wrap; gq += gq_wrap; }
      }
    } else {
      // generic path (any grid width): a lane owns one cell per 32-cell segment
      const int segs = ((x1 - x0) >> 5) + 1;
      for (int ry = warp; ry < rows; ry += n_warps) {
        const int row = (yb0 + ry) * W + x0 + lane;
        const unsigned* mrow = cmap + ry * pw + (lane >> 4);
        for (int sg = 0; sg < segs; ++sg) {
          const unsigned mw = mrow[sg * 2];
          const bool on = (x0 + (sg << 5) + lane <= x1) && (((mw & ~(mw >> 16)) >> (lane & 15)) & 1u);   // free and not an end cell
          if (on) {
            val[row + (sg << 5)] = val[row + (sg << 5)] + f;
            mk[row + (sg << 5)] = (uint8_t)b_free;
          }
        }
      }
    }
    __syncthreads();   // the map is cleared for the next band / phase C touches cells a quad store may rewrite
  }
  };
  if (single_band) run_bands(std::true_type{}); else run_bands(std::false_type{});

  hs_stamp(P, 4);
  // phase C: end points -- the owner beam applies undo-free / occupied and stores markOcc
  // (end cells are disjoint from the cells B2 marks: their free bit is never set)
  for (int i = threadIdx.x; i < (HS_DBG(P, 4) ? 0 : n); i += T) {   // same beam-to-thread mapping as phase A3
    const HsBeamRec r = beams[i];
    if (r.da == 0) continue;
    const int h = hs_hash_find(hkeys, hash_mask, log2_hash, r.kend);
    const int hv = hvals[h];
    if ((hv >> 1) != i) continue;      // not the owner of its end cell
    float v = i < T ? v_end : val[r.kend];
    if (hv & 1) v = (v + f) - f;       // updateSetFree by an earlier beam, then updateUnsetFree
    if (v < 50.0f) v = v + o;          // updateSetOccupied
    val[r.kend] = v;
    mk[r.kend] = (uint8_t)b_occ;
  }
#ifdef HS_EXPERIMENTS
  if (P.dbg_times) { __syncthreads(); hs_stamp(P, 5); }
#endif
}

// ------------------------------------------------------------------------------------------------
// maintenance kernels

// GridMapBase::clear -> LogOddsCell::resetGridCell (GridMapBase.h:69-87, GridMapLogOdds.h:89-93)
__global__ void k_clear_cells(float* __restrict__ val, int* __restrict__ idx, uint8_t* __restrict__ mk, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) { val[i] = 0.0f; idx[i] = -1; mk[i] = 0; }
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
    for (int l = 0; l < HS_MAX_LEVELS; ++l) P.mk_base[(size_t)s * HS_MAX_LEVELS + l] = 0;
    P.coarse_cnt[s] = 0;
    P.coarse_origo[2 * s] = 0.0f; P.coarse_origo[2 * s + 1] = 0.0f;
  }
}

// interleave / de-interleave one level between the SoA planes and the reference's {float,int} cells
// (the marker is the byte plane's where that holds one, the int32 plane's otherwise)
__global__ void k_pack_cells(const float* __restrict__ val, const int* __restrict__ idx, const uint8_t* __restrict__ mk,
                             const int* __restrict__ mk_base, hs_cell* __restrict__ out, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t stride = (size_t)gridDim.x * blockDim.x;
  const int base = *mk_base;
  for (; i < n; i += stride) {
    hs_cell c;
    const unsigned b = mk[i];
    c.logOddsVal = val[i];
    c.updateIndex = b ? base + (int)b : idx[i];
    out[i] = c;
  }
}
__global__ void k_unpack_cells(const hs_cell* __restrict__ in, float* __restrict__ val, int* __restrict__ idx, uint8_t* __restrict__ mk, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) { hs_cell c = in[i]; val[i] = c.logOddsVal; idx[i] = c.updateIndex; mk[i] = 0; }
}
// HectorMappingRos::publishMap classification (HectorMappingRos.cpp:447-470): isFree -> 0, isOccupied -> 100, else -1
__global__ void k_classify(const float* __restrict__ val, int8_t* __restrict__ out, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) { float v = val[i]; out[i] = (v < 0.0f) ? (int8_t)0 : ((v > 0.0f) ? (int8_t)100 : (int8_t)-1); }
}
