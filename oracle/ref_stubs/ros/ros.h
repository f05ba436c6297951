// Stub of the ros::Time / ros::Duration types named in the reference's headers.  Test infrastructure only.
#pragma once
#include <cstdint>
#include <iomanip>
#include <iostream>
#include <sstream>
#include <string>
namespace ros {
struct Duration { int64_t ns = 0; int64_t toNSec() const { return ns; } };
struct Time { int64_t ns = 0; double toSec() const { return 1e-9 * (double)ns; } };
inline Duration operator-(const Time& a, const Time& b) { Duration d; d.ns = a.ns - b.ns; return d; }
}  // namespace ros
namespace std_msgs { struct Header { ros::Time stamp; std::string frame_id; uint32_t seq = 0; }; }
namespace sensor_msgs {
struct Image { std_msgs::Header header; };
typedef std::shared_ptr<Image> ImagePtr;
typedef std::shared_ptr<const Image> ImageConstPtr;
namespace image_encodings { static const std::string TYPE_8SC1 = "8SC1", TYPE_8UC1 = "8UC1", MONO8 = "mono8"; }
}  // namespace sensor_msgs
