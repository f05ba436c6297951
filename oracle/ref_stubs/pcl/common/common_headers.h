// Stub of pcl::PointXYZI / pcl::PointCloud: a vector of points.  Test infrastructure only.
#pragma once
#include <cstdint>
#include <memory>
#include <string>
#include <vector>
namespace pcl {
struct PointXYZI { float x = 0.f, y = 0.f, z = 0.f, intensity = 0.f; };
struct PCLHeader { uint32_t seq = 0; uint64_t stamp = 0; std::string frame_id; };
template <typename PointT>
class PointCloud {
 public:
  typedef std::shared_ptr<PointCloud<PointT>> Ptr;
  typedef std::shared_ptr<const PointCloud<PointT>> ConstPtr;
  PCLHeader header;
  std::vector<PointT> points;
  uint32_t width = 0, height = 0;
  void push_back(const PointT& p) { points.push_back(p); width = (uint32_t)points.size(); height = 1; }
  size_t size() const { return points.size(); }
  void resize(size_t n) { points.resize(n); }
  bool empty() const { return points.empty(); }
};
}  // namespace pcl
