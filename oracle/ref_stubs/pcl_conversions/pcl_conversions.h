// Stub of pcl_conversions::toPCL(stamp).  Test infrastructure only.
#pragma once
#include <cstdint>
#include <pcl/common/common_headers.h>
#include <ros/ros.h>
namespace pcl_conversions {
inline void toPCL(const ros::Time& stamp, uint64_t& pcl_stamp) { pcl_stamp = (uint64_t)(stamp.ns / 1000); }
}  // namespace pcl_conversions
