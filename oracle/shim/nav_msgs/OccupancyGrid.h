// -*- C++ -*-
// TEST INFRASTRUCTURE ONLY -- stand-in for ROS nav_msgs/OccupancyGrid.h (ROS is not in this image).
#ifndef HS_ORACLE_SHIM_NAV_MSGS_OCCUPANCYGRID_H
#define HS_ORACLE_SHIM_NAV_MSGS_OCCUPANCYGRID_H
#include <memory>
#include <vector>
#include <nav_msgs/MapMetaData.h>
namespace nav_msgs {
struct OccupancyGrid {
  MapMetaData info;
  std::vector<int8_t> data;
};
typedef std::shared_ptr<OccupancyGrid> OccupancyGridPtr;
typedef std::shared_ptr<const OccupancyGrid> OccupancyGridConstPtr;
}  // namespace nav_msgs
#endif
