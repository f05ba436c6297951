// Stub of cv_bridge::CvImage (header + encoding + cv::Mat).  Test infrastructure only.
#pragma once
#include <memory>
#include <string>
#include <opencv2/core.hpp>
#include <ros/ros.h>
namespace cv_bridge {
class CvImage {
 public:
  std_msgs::Header header;
  std::string encoding;
  cv::Mat image;
  sensor_msgs::ImagePtr toImageMsg() const { auto m = std::make_shared<sensor_msgs::Image>(); m->header = header; return m; }
};
typedef std::shared_ptr<CvImage> CvImagePtr;
typedef std::shared_ptr<const CvImage> CvImageConstPtr;
}  // namespace cv_bridge
