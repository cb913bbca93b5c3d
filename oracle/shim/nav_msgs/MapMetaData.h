// -*- C++ -*-
// TEST INFRASTRUCTURE ONLY -- stand-in for ROS nav_msgs/MapMetaData.h (ROS is not in this image): just the fields
// hector_map_tools/HectorMapTools.h reads (info.resolution, width, height, origin.position).
#ifndef HS_ORACLE_SHIM_NAV_MSGS_MAPMETADATA_H
#define HS_ORACLE_SHIM_NAV_MSGS_MAPMETADATA_H
#include <stdint.h>
namespace geometry_msgs_shim {
struct Point { double x, y, z; };
struct QuaternionD { double x, y, z, w; };
struct Pose { Point position; QuaternionD orientation; };
}  // namespace geometry_msgs_shim
namespace nav_msgs {
struct MapMetaData {
  float resolution;
  uint32_t width, height;
  geometry_msgs_shim::Pose origin;
};
}  // namespace nav_msgs
#endif
