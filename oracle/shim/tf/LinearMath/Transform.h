// TEST INFRASTRUCTURE ONLY -- stand-in for ROS tf's <tf/LinearMath/Transform.h>, which
// hector_slam_lib/util/UtilFunctions.h:33 includes but which is not present in this image.
// Upstream tf/LinearMath/Scalar.h includes <math.h> and <stdlib.h>; that is what makes the
// reference's unqualified abs()/sin()/cos()/exp()/log() calls resolve to the *float*
// overloads (UtilFunctions.h:88, OccGridMapUtil.h:70-71, GridMapLogOdds.h:165,202).
// The reference headers also rely on transitive includes for UINT_MAX, memcpy, std::cout.
#ifndef HS_ORACLE_TF_TRANSFORM_SHIM
#define HS_ORACLE_TF_TRANSFORM_SHIM

#include <math.h>
#include <stdlib.h>
#include <climits>
#include <cfloat>
#include <cstring>
#include <iostream>
#include <vector>

namespace geometry_msgs {
struct Quaternion { double x, y, z, w; };
}

namespace tf {
class Quaternion {
 public:
  Quaternion(double x, double y, double z, double w) : x_(x), y_(y), z_(z), w_(w) {}
  double x_, y_, z_, w_;
};
static inline double getYaw(const Quaternion& q) {
  return atan2(2.0 * (q.w_ * q.z_ + q.x_ * q.y_), 1.0 - 2.0 * (q.y_ * q.y_ + q.z_ * q.z_));
}
}  // namespace tf

#endif
