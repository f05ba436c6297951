// ref_shim.cc -- C entry points around the reference's OWN filter classes.  TEST INFRASTRUCTURE ONLY.
//
// Linked with /root/reference/src/cfear_radarodometry/radar_filters.cpp and cfar.cpp compiled UNMODIFIED from where they
// lie (Makefile target _ref/libcfear_ref.so; the stub headers under ref_stubs/ stand in for OpenCV / cv_bridge / PCL /
// ROS types and carry no CFEAR arithmetic).  Calls are made exactly like radarDriver::Process does
// (radar_driver.cpp:48-61): float parameters of radarDriver::Parameters (radar_driver.h:40-46) are passed to the int /
// double constructor arguments by implicit conversion.  Used by tests/ to pin the oracle restatement
// (oracle/cfear_oracle.cc) and the CUDA path to the reference source: rows a1-a3 of SURVEY.md section 8 and CA-CFAR.
#include <cstdint>
#include <cstring>

#include "cfear_radarodometry/cfar.h"
#include "cfear_radarodometry/radar_filters.h"

namespace {

cv_bridge::CvImagePtr make_image(const uint8_t* img, int A, int R) {
  cv_bridge::CvImagePtr p = boost::make_shared<cv_bridge::CvImage>();
  p->image = cv::Mat::zeros(A, R, CV_8UC1);
  std::memcpy(p->image.data, img, (size_t)A * R);
  p->image.fill_guard();
  return p;
}

// dense_filtered_ / dense_filtered_peaks_ are protected members (radar_filters.h): read them through a derived class.
struct Probe : public CFEAR_Radarodometry::StructuredKStrongest {
  using StructuredKStrongest::StructuredKStrongest;
  const std::vector<std::vector<intensity_range>>& kept() const { return dense_filtered_; }
  const std::vector<std::vector<intensity_range>>& peaks() const { return dense_filtered_peaks_; }
};

void dump(const std::vector<std::vector<CFEAR_Radarodometry::StructuredKStrongest::intensity_range>>& v, int A, int k,
          int32_t* idx_out, int32_t* cnt_out) {
  for (int a = 0; a < A; ++a) {
    const int n = a < (int)v.size() ? (int)v[a].size() : 0;
    cnt_out[a] = n;
    for (int j = 0; j < k; ++j) idx_out[(size_t)a * k + j] = j < n ? v[a][j].second : -1;
  }
}

int copy_cloud(const pcl::PointCloud<pcl::PointXYZI>::Ptr& c, float* out, int cap) {
  const int n = c ? (int)c->size() : 0;
  for (int i = 0; i < n && i < cap; ++i) {
    out[4 * i + 0] = c->points[i].x; out[4 * i + 1] = c->points[i].y;
    out[4 * i + 2] = c->points[i].z; out[4 * i + 3] = c->points[i].intensity;
  }
  return n;
}

}  // namespace

extern "C" {

// StructuredKStrongest(cv_polar_image, par.z_min, par.k_strongest, par.min_distance, par.range_res) + both
// getPeaksFilteredPointCloud calls of radarDriver::Process.  idx/cnt: kept bins per row in the reference's stored
// order; pidx/pcnt: the AxialNonMaxSupress survivors; cloud / peaks clouds [cap][4] floats, returns their sizes.
int ref_kstrongest(const uint8_t* img, int A, int R, float z_min, int k, float min_distance, float range_res,
                   int32_t* idx_out, int32_t* cnt_out, int32_t* pidx_out, int32_t* pcnt_out,
                   float* cloud_out, int32_t* ncloud, float* peaks_out, int32_t* npeaks, int cap) {
  cv_bridge::CvImagePtr im = make_image(img, A, R);
  Probe filt(im, z_min, k, min_distance, range_res);
  pcl::PointCloud<pcl::PointXYZI>::Ptr cloud(new pcl::PointCloud<pcl::PointXYZI>());
  pcl::PointCloud<pcl::PointXYZI>::Ptr peaks(new pcl::PointCloud<pcl::PointXYZI>());
  filt.getPeaksFilteredPointCloud(cloud, false);
  filt.getPeaksFilteredPointCloud(peaks, true);
  if (idx_out && cnt_out) dump(filt.kept(), A, k, idx_out, cnt_out);
  if (pidx_out && pcnt_out) dump(filt.peaks(), A, k, pidx_out, pcnt_out);
  if (ncloud) *ncloud = cloud_out ? copy_cloud(cloud, cloud_out, cap) : (int)cloud->size();
  if (npeaks) *npeaks = peaks_out ? copy_cloud(peaks, peaks_out, cap) : (int)peaks->size();
  return 0;
}

// AzimuthCACFAR filter(par.window_size, par.false_alarm_rate, par.nb_guard_cells, par.range_res, par.z_min,
//                      par.min_distance, 400.0); filter.getFilteredPointCloud(cv_polar_image, cloud)   radar_driver.cpp:52-56
int ref_cfar(const uint8_t* img, int A, int R, int window_size, float false_alarm_rate, int nb_guard_cells, float range_res,
             float z_min, float min_distance, double max_distance, float* cloud_out, int cap) {
  cv_bridge::CvImagePtr im = make_image(img, A, R);
  AzimuthCACFAR filter(window_size, false_alarm_rate, nb_guard_cells, range_res, z_min, min_distance, max_distance);
  pcl::PointCloud<pcl::PointXYZI>::Ptr cloud(new pcl::PointCloud<pcl::PointXYZI>());
  filter.getFilteredPointCloud(im, cloud);
  return copy_cloud(cloud, cloud_out, cap);
}

}  // extern "C"
