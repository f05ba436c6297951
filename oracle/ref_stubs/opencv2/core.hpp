// Stub of the part of <opencv2/core.hpp> the reference's radar_filters.cpp / cfar.cpp use: an 8UC1 matrix with
// unchecked element access (cv::Mat::at in a release build), Mat::zeros, Mat::row.  Test infrastructure only.
#pragma once
#include <algorithm>
#include <cassert>
#include <cmath>
#include <cstddef>
#include <cstring>
#include <memory>
#include <vector>
typedef unsigned char uchar;
#define CV_8UC1 0
namespace cv {
class Mat {
 public:
  int rows = 0, cols = 0;
  Mat() {}
  // Storage has `guard` bytes on either side of the rows*cols payload.  The reference's AxialNonMaxSupress
  // (radar_filters.cpp:260) indexes up to 3 bins outside a row; in the contiguous cv::Mat that lands in the
  // neighbouring row, and for the first / last row outside the allocation (undefined behaviour upstream).  The guard
  // makes those reads defined: the harness fills it by replicating the first / last payload byte, which is the
  // "clamp to the flat buffer" convention SURVEY.md A.1 documents.
  static constexpr int guard = 64;
  static Mat zeros(int r, int c, int /*type*/) {
    Mat m; m.rows = r; m.cols = c;
    m.store_ = std::make_shared<std::vector<uchar>>((size_t)r * c + 2 * guard, (uchar)0);
    m.data = m.store_->data() + guard;
    return m;
  }
  template <typename T> T& at(int r, int c) { return reinterpret_cast<T*>(data)[(ptrdiff_t)r * cols + c]; }
  template <typename T> const T& at(int r, int c) const { return reinterpret_cast<const T*>(data)[(ptrdiff_t)r * cols + c]; }
  template <typename T> T& at(int i) { return reinterpret_cast<T*>(data)[i]; }
  template <typename T> const T& at(int i) const { return reinterpret_cast<const T*>(data)[i]; }
  Mat row(int r) const { Mat m; m.rows = 1; m.cols = cols; m.store_ = store_; m.data = data + (ptrdiff_t)r * cols; return m; }
  void fill_guard() {
    if (!store_ || rows * cols == 0) return;
    std::memset(data - guard, data[0], guard);
    std::memset(data + (size_t)rows * cols, data[(size_t)rows * cols - 1], guard);
  }
  uchar* data = nullptr;
 private:
  std::shared_ptr<std::vector<uchar>> store_;
};
}  // namespace cv
