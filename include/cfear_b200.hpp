// cfear_b200.hpp -- C++ host mirror of the reference's per-scan API over the C ABI (cfear_b200.h).
//
// Same class names, argument meaning and error behaviour as the reference
// (dan11003/CFEAR_Radarodometry_code_public), minus ROS / PCL / Eigen / Ceres types, which are replaced by
// the small PODs below.  A maintainer of the reference swaps the bodies of radarDriver::Process,
// MapPointNormal::MapPointNormal and n_scan_normal_reg::Register for calls into this layer.
//
//   radarDriver            include/cfear_radarodometry/radar_driver.h:30-118, src/.../radar_driver.cpp:23-176
//   cell, MapPointNormal   include/cfear_radarodometry/pointnormal.h:45-199, src/.../pointnormal.cpp:7-297
//   Registration,
//   n_scan_normal_reg      include/cfear_radarodometry/registration.h:48-131, n_scan_normal.h:28-81,
//                          src/.../n_scan_normal.cpp:82-187
//   OdometryKeyframeFuser  include/cfear_radarodometry/odometrykeyframefuser.h, src/.../odometrykeyframefuser.cpp:62-259,470-494
//   vectorToAffine3d, Affine3dToVectorXYeZ   registration.cpp:130-144, utils.cpp:115-122
//
// Header-only; link with libcfear_b200.so.  Everything computes on the GPU; with no CUDA device the first
// call throws std::runtime_error (there is no CPU fallback).
#ifndef CFEAR_B200_HPP_
#define CFEAR_B200_HPP_

#include <algorithm>
#include <array>
#include <cassert>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iomanip>
#include <iostream>
#include <map>
#include <memory>
#include <mutex>
#include <sstream>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "cfear_b200.h"

namespace CFEAR_Radarodometry {

// ---- minimal stand-ins for the Eigen / PCL / sensor_msgs types on the reference's signatures ------------
struct Vector2d {
  double v[2];
  Vector2d() : v{0, 0} {}
  Vector2d(double x, double y) : v{x, y} {}
  double& operator()(int i) { return v[i]; }
  double operator()(int i) const { return v[i]; }
  double dot(const Vector2d& o) const { return v[0] * o.v[0] + v[1] * o.v[1]; }
  double norm() const { return std::sqrt(dot(*this)); }
};
struct Matrix2d {
  double m[4];
  Matrix2d() : m{0, 0, 0, 0} {}
  double& operator()(int r, int c) { return m[2 * r + c]; }
  double operator()(int r, int c) const { return m[2 * r + c]; }
};
struct Matrix6d {
  double m[36];
  Matrix6d() { for (double& x : m) x = 0; }
  static Matrix6d Identity() { Matrix6d I; for (int i = 0; i < 6; ++i) I.m[7 * i] = 1; return I; }
  double& operator()(int r, int c) { return m[6 * r + c]; }
  double operator()(int r, int c) const { return m[6 * r + c]; }
};
// Rigid transform as the 4x4 homogeneous matrix Eigen::Affine3d stores (row-major here).
struct Affine3d {
  double m[16];
  Affine3d() { for (int i = 0; i < 16; ++i) m[i] = (i % 5 == 0) ? 1.0 : 0.0; }
  static Affine3d Identity() { return Affine3d(); }
  double& operator()(int r, int c) { return m[4 * r + c]; }
  double operator()(int r, int c) const { return m[4 * r + c]; }
  Affine3d operator*(const Affine3d& o) const {
    Affine3d r;
    for (int i = 0; i < 4; ++i)
      for (int j = 0; j < 4; ++j) {
        double s = 0;
        for (int k = 0; k < 4; ++k) s += m[4 * i + k] * o.m[4 * k + j];
        r.m[4 * i + j] = s;
      }
    return r;
  }
  Affine3d inverse() const {   // rigid inverse
    Affine3d r;
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) r.m[4 * i + j] = m[4 * j + i];
    for (int i = 0; i < 3; ++i) r.m[4 * i + 3] = -(r.m[4 * i] * m[3] + r.m[4 * i + 1] * m[7] + r.m[4 * i + 2] * m[11]);
    return r;
  }
};
// registration.cpp:130-144
inline Affine3d vectorToAffine3d(const std::vector<double>& vek) {
  assert(vek.size() == 3);
  Affine3d T;
  const double c = std::cos(vek[2]), s = std::sin(vek[2]);
  T(0, 0) = c; T(0, 1) = -s; T(1, 0) = s; T(1, 1) = c; T(0, 3) = vek[0]; T(1, 3) = vek[1];
  return T;
}
inline Affine3d vectorToAffine3d(double x, double y, double yaw) { return vectorToAffine3d(std::vector<double>{x, y, yaw}); }
// utils.cpp:115-122: (x, y, eulerAngles(0,1,2)[2]); for a planar pose that angle is atan2(R(1,0), R(1,1)).
inline void Affine3dToVectorXYeZ(const Affine3d& T, std::vector<double>& par) {
  if (par.size() != 3) par.resize(3, 0);
  par[0] = T(0, 3); par[1] = T(1, 3); par[2] = std::atan2(T(1, 0), T(1, 1));
}

typedef cfear_point PointXYZI;                       // pcl::PointXYZI (x, y, z, intensity)
struct PointCloud {                                  // pcl::PointCloud<pcl::PointXYZI>
  std::vector<PointXYZI> points;
  uint64_t stamp = 0;
  size_t size() const { return points.size(); }
  void clear() { points.clear(); }
  void push_back(const PointXYZI& p) { points.push_back(p); }
};
typedef std::shared_ptr<PointCloud> CloudPtr;
struct PolarImage {                                  // sensor_msgs::Image, 8UC1: rows = azimuths, cols = range bins
  int rows = 0, cols = 0;
  const uint8_t* data = nullptr;
  uint64_t stamp = 0;
};

typedef enum weight_options { Uniform = 0, Sim_N = 1, Sim_direciton = 2, Sim_scale = 3, Combined_weights = 4 } weightoption;   // registration.h:50
typedef enum costmetric { P2P, P2L, P2D } cost_metric;                                                                     // registration.h:55
typedef enum losstype { None, Huber, Cauchy, SoftLOne, Combined, Tukey } loss_type;                                        // registration.h:60

inline cost_metric Str2Cost(const std::string& s) { return s == "P2P" ? P2P : (s == "P2D" ? P2D : P2L); }
inline loss_type Str2loss(const std::string& s) {
  return s == "Huber" ? Huber : s == "Cauchy" ? Cauchy : s == "SoftLOne" ? SoftLOne : s == "Combined" ? Combined : s == "Tukey" ? Tukey : None;
}

// ---- the one device context all mirror objects share (the reference keeps comparable process-wide state:
// statics in pointnormal.cpp:4-5, types.cpp:74, statistics.cpp:6) ------------------------------------------
class Backend {
 public:
  // Must be called before the first object is built if the defaults (400x3360, k=12, 64 cell-set slots) do not fit.
  static void Configure(const cfear_config& cfg) {
    Backend& b = get();
    std::lock_guard<std::mutex> l(b.mu_);
    if (b.ctx_) throw std::runtime_error("cfear Backend already created");
    b.cfg_ = cfg; b.configured_ = true;
  }
  static Backend& get() { static Backend b; return b; }
  cfear_ctx* ctx() {
    std::lock_guard<std::mutex> l(mu_);
    if (!ctx_) {
      if (!configured_) { cfear_default_config(&cfg_); cfg_.max_cellsets = 64; cfg_.max_keyframes = 8; }
      if (cfear_create(&cfg_, &ctx_) != CFEAR_OK) throw std::runtime_error(std::string("cfear_create: ") + cfear_last_error());
      for (int s = cfg_.max_cellsets - 1; s >= 0; --s) free_.push_back(s);
    }
    return ctx_;
  }
  cfear_config& cfg() { ctx(); return cfg_; }
  void apply() { if (cfear_update_config(ctx(), &cfg_) != CFEAR_OK) throw std::runtime_error(cfear_last_error()); }
  int acquire_slot() {
    ctx();
    std::lock_guard<std::mutex> l(mu_);
    if (free_.empty()) throw std::runtime_error("cfear Backend: out of cell-set slots (raise cfear_config.max_cellsets)");
    int s = free_.back(); free_.pop_back(); return s;
  }
  void release_slot(int s) { std::lock_guard<std::mutex> l(mu_); free_.push_back(s); }
  ~Backend() { if (ctx_) cfear_destroy(ctx_); }

 private:
  Backend() {}
  std::mutex mu_;
  cfear_config cfg_;
  bool configured_ = false;
  cfear_ctx* ctx_ = nullptr;
  std::vector<int> free_;
};

// ---- radarDriver (radar_driver.h:30-118) ---------------------------------------------------------------------
typedef enum filter_type { kstrong, CACFAR } filtertype;   // radar_driver.h:24

class radarDriver {
 public:
  class Parameters {
   public:
    float z_min = 60;
    float range_res = 0.0438f;
    int azimuths = 400, k_strongest = 12;
    int nb_guard_cells = 20, window_size = 10;
    float false_alarm_rate = 0.01f;
    float min_distance = 2.5f, max_distance = 200;
    std::string dataset = "oxford";
    filtertype filter_type_ = kstrong;
    std::string ToString() {
      std::ostringstream s;
      s << "range res, " << range_res << std::endl << "z min, " << z_min << std::endl << "min distance, " << min_distance << std::endl
        << "max distance, " << max_distance << std::endl << "k strongest, " << k_strongest << std::endl << "dataset, " << dataset << std::endl
        << "filter type, kstrong" << std::endl;
      return s.str();
    }
  };
  radarDriver(const Parameters& pars, bool disable_callback = false) : par(pars) {
    (void)disable_callback;
    cloud_filtered_ = std::make_shared<PointCloud>();
    cloud_filtered_peaks_ = std::make_shared<PointCloud>();
  }
  // radar_driver.cpp:163-176 -> CallbackOxford :99-111 -> Process :48-73.  Hands out the driver's internal clouds
  // (aliasing, like the reference).  A NULL image terminates the process (radar_driver.cpp:101-104).
  void CallbackOffline(const PolarImage& radar_image_polar, CloudPtr& cloud, CloudPtr& cloud_peaks) {
    if (radar_image_polar.data == nullptr) { std::cerr << "Radar image NULL" << std::endl; std::exit(0); }
    Backend& b = Backend::get();
    cfear_config& cfg = b.cfg();
    if (radar_image_polar.rows != cfg.azimuths || radar_image_polar.cols != cfg.range_bins || par.k_strongest != cfg.k_strongest)
      throw std::runtime_error("radarDriver: image shape / k_strongest differ from the Backend configuration");
    cfg.z_min = par.z_min; cfg.range_res = par.range_res; cfg.min_distance = par.min_distance;
    b.apply();
    polar_image = radar_image_polar;                                    // cv_polar_image (radar_driver.h:92)
    const int cap = cfg.azimuths * cfg.k_strongest;
    cloud_filtered_ = std::make_shared<PointCloud>();
    cloud_filtered_peaks_ = std::make_shared<PointCloud>();
    if (par.filter_type_ == CACFAR) {                                   // radar_driver.cpp:52-56 (the peaks cloud stays empty)
      cfear_cfar_params cp; cp.window_size = par.window_size; cp.nb_guard_cells = par.nb_guard_cells;
      cp.false_alarm_rate = par.false_alarm_rate; cp.max_distance = 400.0;
      int32_t n = 0;
      int capacity = 4 * cap;
      for (int attempt = 0; attempt < 2; ++attempt) {
        cloud_filtered_->points.resize(capacity);
        const int rc = cfear_cfar_filter(b.ctx(), radar_image_polar.data, 1, &cp, cloud_filtered_->points.data(), capacity, &n);
        if (rc == CFEAR_OK) break;
        if (rc != CFEAR_ERR_CAPACITY || attempt == 1) throw std::runtime_error(std::string("cfear_cfar_filter: ") + cfear_last_error());
        capacity = n;
      }
      cloud_filtered_->points.resize(n);
      cloud_filtered_->stamp = radar_image_polar.stamp;
      cloud = cloud_filtered_; cloud_peaks = cloud_filtered_peaks_;
      return;
    }
    cloud_filtered_->points.resize(cap); cloud_filtered_peaks_->points.resize(cap);
    int32_t n = 0, np = 0;
    if (cfear_filter(b.ctx(), radar_image_polar.data, 1, nullptr, nullptr, cloud_filtered_->points.data(), &n,
                     cloud_filtered_peaks_->points.data(), &np) != CFEAR_OK)
      throw std::runtime_error(std::string("cfear_filter: ") + cfear_last_error());
    cloud_filtered_->points.resize(n); cloud_filtered_peaks_->points.resize(np);
    cloud_filtered_->stamp = cloud_filtered_peaks_->stamp = radar_image_polar.stamp;
    cloud = cloud_filtered_; cloud_peaks = cloud_filtered_peaks_;
  }
  PolarImage polar_image;   // latest radar image (cv_polar_image)

 private:
  Parameters par;
  CloudPtr cloud_filtered_, cloud_filtered_peaks_;
};

// Compensate(cloud, Tmotion, ccw)  utils.cpp:96-113 (in place, like the reference)
inline void Compensate(PointCloud& cloud, const Affine3d& Tmotion, bool ccw) {
  std::vector<double> mot;
  Affine3dToVectorXYeZ(Tmotion, mot);
  if (cloud.points.empty()) return;
  if (cfear_compensate(Backend::get().ctx(), cloud.points.data(), (int)cloud.points.size(), mot.data(), ccw ? 1 : 0) != CFEAR_OK)
    throw std::runtime_error(std::string("cfear_compensate: ") + cfear_last_error());
}

// The two clouds of a frame (filtered + peaks, odometrykeyframefuser.cpp:148-149) in ONE device round trip: the points are
// independent, so concatenating them changes nothing but the number of copies and synchronisations per frame.
inline void Compensate(PointCloud& cloud, PointCloud& cloud_peaks, const Affine3d& Tmotion, bool ccw) {
  const size_t n1 = cloud.points.size(), n2 = cloud_peaks.points.size();
  const cfear_config& g = Backend::get().cfg();
  const size_t cap = (size_t)g.max_batch * (size_t)g.azimuths * (size_t)g.k_strongest;      // what cfear_compensate accepts
  if (n1 == 0 || n2 == 0 || n1 + n2 > cap) { Compensate(cloud, Tmotion, ccw); Compensate(cloud_peaks, Tmotion, ccw); return; }
  std::vector<double> mot;
  Affine3dToVectorXYeZ(Tmotion, mot);
  std::vector<PointXYZI> both(n1 + n2);
  std::copy(cloud.points.begin(), cloud.points.end(), both.begin());
  std::copy(cloud_peaks.points.begin(), cloud_peaks.points.end(), both.begin() + n1);
  if (cfear_compensate(Backend::get().ctx(), both.data(), (int)both.size(), mot.data(), ccw ? 1 : 0) != CFEAR_OK)
    throw std::runtime_error(std::string("cfear_compensate: ") + cfear_last_error());
  std::copy(both.begin(), both.begin() + n1, cloud.points.begin());
  std::copy(both.begin() + n1, both.end(), cloud_peaks.points.begin());
}

// ---- cell / MapPointNormal (pointnormal.h:45-199) -------------------------------------------------------------
class cell {
 public:
  double GetPlanarity() { return scale_; }
  Vector2d u_;
  Matrix2d cov_;
  double scale_ = 0;
  Vector2d snormal_;
  double avg_intensity_ = 0;
  size_t Nsamples_ = 0;
  bool valid_ = true;
};

class MapPointNormal;
typedef std::shared_ptr<MapPointNormal> MapNormalPtr;

class MapPointNormal {
 public:
  static double& downsample_factor() { static double f = 1; return f; }   // pointnormal.cpp:5
  // pointnormal.cpp:65-90.  An empty cloud terminates the process like the reference (:72-75).
  MapPointNormal(const CloudPtr& cld, float radius, const Vector2d& origin = Vector2d(0, 0), const bool weight_intensity = false,
                 const bool raw = false)
      : input_(cld), radius_(radius) {
    if (input_->size() == 0) { std::cout << "error, cloud empty" << std::endl; std::exit(0); }
    if (origin(0) != 0.0 || origin(1) != 0.0) throw std::runtime_error("MapPointNormal: only origin (0,0) is supported (the only value the reference passes)");
    Backend& b = Backend::get();
    slot_ = b.acquire_slot();
    int32_t nc = 0;
    if (raw) {                                   // :76-82 identity cell per point
      std::vector<cfear_cell> cs(input_->size());
      for (size_t i = 0; i < cs.size(); ++i) {
        cfear_cell& c = cs[i];
        c.mean[0] = input_->points[i].x; c.mean[1] = input_->points[i].y; c.normal[0] = 1; c.normal[1] = 0;
        c.cov[0] = 0.1; c.cov[1] = 0; c.cov[2] = 0; c.cov[3] = 0.1; c.planarity = 1.0; c.avg_intensity = 1.0; c.nsamples = 1; c.pad = 0;
      }
      if (cfear_cells_upload(b.ctx(), slot_, cs.data(), (int)cs.size()) != CFEAR_OK) fail("cfear_cells_upload");
      nc = (int32_t)cs.size();
    } else {
      cfear_config& cfg = b.cfg();
      cfg.radius = radius; cfg.weight_intensity = weight_intensity ? 1 : 0; cfg.downsample_factor = downsample_factor();
      b.apply();
      if (cfear_surface_points(b.ctx(), input_->points.data(), (int)input_->size(), slot_, &nc) != CFEAR_OK) fail("cfear_surface_points");
    }
    std::vector<cfear_cell> cs(nc > 0 ? nc : 1);
    if (cfear_cells_download(b.ctx(), slot_, cs.data(), (int)cs.size(), &nc) != CFEAR_OK) fail("cfear_cells_download");
    cells.resize(nc);
    for (int i = 0; i < nc; ++i) {
      cell& c = cells[i];
      c.u_ = Vector2d(cs[i].mean[0], cs[i].mean[1]); c.snormal_ = Vector2d(cs[i].normal[0], cs[i].normal[1]);
      for (int k = 0; k < 4; ++k) c.cov_.m[k] = cs[i].cov[k];
      c.scale_ = cs[i].planarity; c.avg_intensity_ = cs[i].avg_intensity; c.Nsamples_ = (size_t)cs[i].nsamples;
    }
  }
  ~MapPointNormal() { if (slot_ >= 0) Backend::get().release_slot(slot_); }
  MapPointNormal(const MapPointNormal&) = delete;
  MapPointNormal& operator=(const MapPointNormal&) = delete;

  std::vector<cell> GetCells() { return cells; }
  cell& GetCell(const size_t i) { return cells[i]; }
  Vector2d GetMean2d(const size_t i) { return cells[i].u_; }
  Matrix2d GetCov2d(const size_t i) { return cells[i].cov_; }
  Vector2d GetNormal2d(const size_t i) { return cells[i].snormal_; }
  size_t GetSize() { return cells.size(); }
  CloudPtr GetScan() { return input_; }
  // pointnormal.cpp:238-254: 0 or 1 index
  std::vector<int> GetClosestIdx(const Vector2d& p, double d) {
    int32_t idx = -1;
    const double q[2] = {p(0), p(1)};
    if (cfear_nearest(Backend::get().ctx(), slot_, q, 1, d, &idx) != CFEAR_OK) fail("cfear_nearest");
    return idx >= 0 ? std::vector<int>{idx} : std::vector<int>();
  }
  // pointnormal.cpp:139-143 with GetRelTimeStamp of utils.h:28-32: where in the sweep (-0.5 .. 0.5) the cell was observed
  double GetCellRelTimeStamp(const size_t index, const bool ccw) {
    const double x = cells[index].u_(0), y = cells[index].u_(1);
    const double a = std::atan2(y, x);
    const double d = ((a > 0.00001 ? a : (2 * M_PI + a)) / (2 * M_PI));
    return ccw ? -(d - 0.5) : (d - 0.5);
  }
  int slot() const { return slot_; }

 private:
  static void fail(const char* what) { throw std::runtime_error(std::string(what) + ": " + cfear_last_error()); }
  std::vector<cell> cells;
  CloudPtr input_;
  float radius_;
  int slot_ = -1;
};

// ---- Registration / n_scan_normal_reg (registration.h:63-131, n_scan_normal.h:28-81) ----------------------------
struct SolverSummary {            // the fields of ceres::Solver::Summary the reference reads (n_scan_normal.cpp:123-166)
  double final_cost = 0;
  int num_residuals = 0, num_residual_blocks = 0;
  int num_inner_iterations = 0;   // summed over the outer iterations
  bool usable = true;
  bool IsSolutionUsable() const { return usable; }
};

typedef std::pair<int, int> int_pair;     // registration.h:44

class Registration {
 public:
  virtual ~Registration() {}
  virtual bool Register(std::vector<MapNormalPtr>& scans, std::vector<Affine3d>& Tsrc, std::vector<Matrix6d>& reg_cov, bool soft_constraints = false) = 0;
  virtual double getScore() { return score_; }
  class Weights {                         // registration.h:88-101, registration.cpp:67-76
   public:
    Weights(double N1, double N2, double sim_dir, double plan1, double plan2) : N1_(N1), N2_(N2), sim_dir_(sim_dir), plan1_(plan1), plan2_(plan2) {}
    double GetWeight(const weightoption opt) {
      switch (opt) {
        case Uniform: return 1.0;
        case Sim_N: return Similarity(N1_, N2_);
        case Sim_direciton: return sim_dir_;
        case Sim_scale: return Similarity(plan1_, plan2_);
        case Combined_weights: return GetWeight(Sim_N) + GetWeight(Sim_direciton) + GetWeight(Sim_scale);
        default: return 1.0;
      }
    }
    double Similarity(const double x, const double y) { return 2 * std::min(x, y) / (x + y); }
    double N1_, N2_;
    double sim_dir_;
    double plan1_, plan2_;
  };
  weightoption weight_opt_ = Uniform;
  // the last outer iteration's data association, keyed (target scan index, source scan index): pairs (target cell, source
  // cell) and their weight terms, in source-cell order (registration.h:103-106, n_scan_normal.cpp:255-256).  Filled by
  // Register() when keep_associations_ is set (they cost a device->host copy the pose path does not need).
  std::map<int_pair, std::vector<Weights> > weight_associations_;
  std::map<int_pair, std::vector<int_pair> > scan_associations_;
  bool keep_associations_ = false;
  size_t itr_ = 0;
  SolverSummary summary_;

 protected:
  cost_metric cost_ = P2L;
  loss_type loss_ = Huber;
  double loss_limit_ = 0.1;
  double radius_ = 2.0;           // registration.h:122
  double score_ = 0;
};

class n_scan_normal_reg : public Registration {
 public:
  n_scan_normal_reg() {}
  n_scan_normal_reg(const cost_metric& cost, loss_type loss = Huber, double loss_limit = 0.1, const weightoption opt = Uniform) {
    cost_ = cost; loss_ = loss; loss_limit_ = loss_limit; weight_opt_ = opt;
  }
  void SetD2dPar(const double cov_scale, const double regularization) { cov_scale_ = cov_scale; regularization_ = regularization; }
  void SetParameters(unsigned int max_itr_association, unsigned int max_itr_solver) { max_itr_association_ = max_itr_association; max_itr_solver_ = max_itr_solver; }
  double getScore() { return score_; }
  void getScore(double& score, int& num_residuals) { score = score_; num_residuals = summary_.num_residuals; }

  // n_scan_normal.cpp:82-187.  scans.back() is the free block, the others are fixed keyframes
  // (mode_ is always incremental_last_to_previous).  Tsrc / reg_cov are rewritten like the reference.
  bool Register(std::vector<MapNormalPtr>& scans, std::vector<Affine3d>& Tsrc, std::vector<Matrix6d>& reg_cov, bool soft_constraints = false) {
    const size_t n_scans = scans.size();
    assert(Tsrc.size() == n_scans && reg_cov.size() == n_scans);
    Backend& b = Backend::get();
    cfear_config& cfg = b.cfg();
    cfg.cost = cost_ == P2P ? CFEAR_COST_P2P : (cost_ == P2L ? CFEAR_COST_P2L : CFEAR_COST_P2D);
    cfg.loss = (int)loss_; cfg.loss_limit = loss_limit_; cfg.weight_opt = (int)weight_opt_;
    cfg.cov_scale = cov_scale_; cfg.regularization = regularization_; cfg.reg_radius = radius_;
    cfg.max_outer = (int)max_itr_association_; cfg.min_outer = (int)min_itr_; cfg.max_inner = (int)max_itr_solver_;
    cfg.solver_mode = CFEAR_SOLVER_CERES_LM;
    b.apply();
    std::vector<int32_t> slots(n_scans);
    std::vector<double> poses(3 * n_scans);
    std::vector<double> par;
    for (size_t i = 0; i < n_scans; ++i) {
      assert(scans[i] != nullptr);
      slots[i] = scans[i]->slot();
      Affine3dToVectorXYeZ(Tsrc[i], par);                     // :88-92
      poses[3 * i] = par[0]; poses[3 * i + 1] = par[1]; poses[3 * i + 2] = par[2];
    }
    double cov36[36];
    cfear_reg_stats st;
    double L9[9];
    if (soft_constraints && !PriorSqrtInformation(reg_cov.back(), L9))        // :374 Cov6to3(cov).inverse().llt().matrixL()
      throw std::runtime_error("n_scan_normal_reg: the covariance of the guess is not positive definite");
    scan_associations_.clear(); weight_associations_.clear();                 // :103-104
    std::vector<int32_t> assoc; std::vector<double> sim;
    const size_t max_cells = (size_t)(cfg.max_cells > 0 ? cfg.max_cells : cfg.azimuths * cfg.k_strongest);
    if (keep_associations_) { assoc.assign((n_scans - 1) * max_cells, -1); sim.assign((n_scans - 1) * max_cells, 0.0); }
    if (cfear_register_batch_ex(b.ctx(), 1, slots.data(), (int)n_scans, poses.data(), cov36, &st, keep_associations_ ? assoc.data() : nullptr,
                                keep_associations_ ? sim.data() : nullptr, soft_constraints ? L9 : nullptr) != CFEAR_OK)
      throw std::runtime_error(std::string("cfear_register_batch_ex: ") + cfear_last_error());
    if (keep_associations_) {
      MapNormalPtr& src = scans[n_scans - 1];
      for (size_t i = 0; i + 1 < n_scans; ++i) {
        const int_pair scan_pair((int)i, (int)n_scans - 1);
        for (size_t j = 0; j < src->GetSize(); ++j) {
          const int m = assoc[i * max_cells + j];
          if (m < 0) continue;
          weight_associations_[scan_pair].push_back(Weights((double)src->GetCell(j).Nsamples_, (double)scans[i]->GetCell(m).Nsamples_,
                                                            sim[i * max_cells + j], src->GetCell(j).GetPlanarity(), scans[i]->GetCell(m).GetPlanarity()));
          scan_associations_[scan_pair].push_back(std::make_pair(m, (int)j));
        }
      }
    }
    itr_ = (size_t)st.outer_iterations;
    summary_.final_cost = st.final_cost; summary_.num_residuals = st.num_residuals; summary_.num_residual_blocks = st.num_blocks;
    summary_.num_inner_iterations = st.inner_iterations; summary_.usable = st.usable != 0;
    // Tsrc[i] = vectorToAffine3d(parameters[i]) happens after every successful solve (:119-121,177-178); a later failing
    // iteration leaves the pose of the last good one in place
    if (st.pose_written)
      for (size_t i = 0; i < n_scans; ++i) Tsrc[i] = vectorToAffine3d(poses[3 * i], poses[3 * i + 1], poses[3 * i + 2]);
    const bool reached_cov = st.num_residuals > 1 && st.usable;   // the `if(success)` block at :163 (the loop ended without a failure)
    if (reached_cov) {
      score_ = st.score;
      Matrix6d d; d(0, 0) = 0.01; d(1, 1) = 0.01; d(5, 5) = 0.0001;   // :171-175
      for (size_t i = 0; i < n_scans; ++i) reg_cov[i] = d;
      Matrix6d c; for (int i = 0; i < 36; ++i) c.m[i] = cov36[i];
      if (st.success) reg_cov.back() = c;                              // GetCovariance :392-433
    }
    return st.success != 0;
  }

  // n_scan_normal.cpp:187-213.  Cost of the problem built at the poses given (nothing is optimised).  The per-residual
  // values stay on the device: `residuals` is sized to the number of scalar residuals and zero-filled (the only caller,
  // approximateCovarianceBySampling, reads the score alone).
  bool GetCost(std::vector<MapNormalPtr>& scans, std::vector<Affine3d>& Tsrc, double& score, std::vector<double>& residuals) {
    std::vector<std::vector<Affine3d> > one(1, Tsrc);
    std::vector<double> scores; std::vector<int> nres; std::vector<bool> ok;
    GetCostBatch(scans, one, scores, nres, ok);
    score = scores[0];
    residuals.assign((size_t)std::max(nres[0], 0), 0.0);
    if (ok[0]) score_ = score / std::max(nres[0], 1);               // :211
    return ok[0];
  }
  // The same for many pose hypotheses of one set of scans in a single launch (what the sampling loop of
  // odometrykeyframefuser.cpp:293-320 does one call at a time).
  void GetCostBatch(std::vector<MapNormalPtr>& scans, const std::vector<std::vector<Affine3d> >& Tsrc_samples,
                    std::vector<double>& scores, std::vector<int>& num_residuals, std::vector<bool>& ok) {
    const size_t n_scans = scans.size(), ns = Tsrc_samples.size();
    assert(n_scans >= 2);
    Backend& b = Backend::get();
    cfear_config& cfg = b.cfg();
    cfg.cost = cost_ == P2P ? CFEAR_COST_P2P : (cost_ == P2L ? CFEAR_COST_P2L : CFEAR_COST_P2D);
    cfg.loss = (int)loss_; cfg.loss_limit = loss_limit_; cfg.weight_opt = (int)weight_opt_;
    cfg.cov_scale = cov_scale_; cfg.regularization = regularization_; cfg.reg_radius = radius_;
    b.apply();
    std::vector<int32_t> slots(ns * n_scans);
    std::vector<double> poses(ns * n_scans * 3), par;
    for (size_t s = 0; s < ns; ++s) {
      assert(Tsrc_samples[s].size() == n_scans);
      for (size_t i = 0; i < n_scans; ++i) {
        slots[s * n_scans + i] = scans[i]->slot();
        Affine3dToVectorXYeZ(Tsrc_samples[s][i], par);
        for (int k = 0; k < 3; ++k) poses[(s * n_scans + i) * 3 + k] = par[k];
      }
    }
    scores.assign(ns, 0.0); num_residuals.assign(ns, 0); ok.assign(ns, false);
    std::vector<int32_t> nr(ns), okv(ns);
    const size_t max_batch = (size_t)cfg.max_batch;
    for (size_t s0 = 0; s0 < ns; s0 += max_batch) {
      const int nb = (int)std::min(max_batch, ns - s0);
      if (cfear_get_cost_batch(b.ctx(), nb, slots.data() + s0 * n_scans, (int)n_scans, poses.data() + s0 * n_scans * 3,
                               scores.data() + s0, nr.data() + s0, okv.data() + s0) != CFEAR_OK)
        throw std::runtime_error(std::string("cfear_get_cost_batch: ") + cfear_last_error());
    }
    for (size_t s = 0; s < ns; ++s) { num_residuals[s] = nr[s]; ok[s] = okv[s] != 0; }
  }
  // Cov6to3 (registration.cpp:123-129) -> inverse -> lower Cholesky factor, row-major 3x3; false if not positive definite
  static bool PriorSqrtInformation(const Matrix6d& C, double L[9]) {
    const double a = C(0, 0), b = C(0, 1), c = C(0, 5), d = C(1, 0), e = C(1, 1), f = C(1, 5), g = C(5, 0), h = C(5, 1), i = C(5, 5);
    const double det = a * (e * i - f * h) - b * (d * i - f * g) + c * (d * h - e * g);
    if (!(std::fabs(det) > 0.0) || !std::isfinite(det)) return false;
    const double I[9] = {(e * i - f * h) / det, (c * h - b * i) / det, (b * f - c * e) / det,
                         (f * g - d * i) / det, (a * i - c * g) / det, (c * d - a * f) / det,
                         (d * h - e * g) / det, (b * g - a * h) / det, (a * e - b * d) / det};
    for (int k = 0; k < 9; ++k) L[k] = 0.0;
    if (!(I[0] > 0.0)) return false;                        // llt() reads the lower triangle
    L[0] = std::sqrt(I[0]);
    L[3] = I[3] / L[0]; L[6] = I[6] / L[0];
    const double d1 = I[4] - L[3] * L[3];
    if (!(d1 > 0.0)) return false;
    L[4] = std::sqrt(d1);
    L[7] = (I[7] - L[6] * L[3]) / L[4];
    const double d2 = I[8] - L[6] * L[6] - L[7] * L[7];
    if (!(d2 > 0.0)) return false;
    L[8] = std::sqrt(d2);
    return true;
  }
  // n_scan_normal.cpp:435-441 (summary_ of the last Register; three parameters in the reduced problem)
  bool GetCovarianceScaler(double& cov_scale) {
    if (summary_.num_residuals - 3 == 0) return false;
    cov_scale = summary_.final_cost / (summary_.num_residuals - 3);
    return true;
  }

 private:
  double cov_scale_ = 1;
  double regularization_ = 0.01;
  double max_itr_association_ = 8, min_itr_ = 3;
  unsigned int max_itr_solver_ = 20;
};

// ---- small dense linear algebra for the covariance-by-sampling fit (the reference uses Eigen there) ------------------
namespace detail {
// min |A q - c|_2, A row-major m x n with m >= n, by Householder QR on column-equilibrated A.  False if rank deficient.
inline bool lstsq_qr(std::vector<double> A, int m, int n, std::vector<double> c, double* q) {
  std::vector<double> cs(n, 1.0);
  for (int j = 0; j < n; ++j) {
    double mx = 0; for (int i = 0; i < m; ++i) mx = std::max(mx, std::fabs(A[(size_t)i * n + j]));
    if (mx == 0) return false;
    cs[j] = 1.0 / mx;
    for (int i = 0; i < m; ++i) A[(size_t)i * n + j] *= cs[j];
  }
  for (int k = 0; k < n; ++k) {
    double nrm = 0; for (int i = k; i < m; ++i) nrm += A[(size_t)i * n + k] * A[(size_t)i * n + k];
    nrm = std::sqrt(nrm);
    if (nrm < 1e-13) return false;
    const double alpha = A[(size_t)k * n + k] > 0 ? -nrm : nrm;
    std::vector<double> v(m, 0.0);
    for (int i = k; i < m; ++i) v[i] = A[(size_t)i * n + k];
    v[k] -= alpha;
    double vv = 0; for (int i = k; i < m; ++i) vv += v[i] * v[i];
    if (vv == 0) continue;
    for (int j = k; j < n; ++j) {
      double d = 0; for (int i = k; i < m; ++i) d += v[i] * A[(size_t)i * n + j];
      d = 2 * d / vv;
      for (int i = k; i < m; ++i) A[(size_t)i * n + j] -= d * v[i];
    }
    double d = 0; for (int i = k; i < m; ++i) d += v[i] * c[i];
    d = 2 * d / vv;
    for (int i = k; i < m; ++i) c[i] -= d * v[i];
  }
  for (int k = n - 1; k >= 0; --k) {
    double t = c[k];
    for (int j = k + 1; j < n; ++j) t -= A[(size_t)k * n + j] * q[j];
    q[k] = t / A[(size_t)k * n + k];
  }
  for (int j = 0; j < n; ++j) q[j] *= cs[j];
  return true;
}
// eigenvalues of a symmetric 3x3 (cyclic Jacobi)
inline void eig3_sym(const double H[9], double ev[3]) {
  double a[9]; for (int i = 0; i < 9; ++i) a[i] = H[i];
  for (int sweep = 0; sweep < 50; ++sweep) {
    const double off = std::fabs(a[1]) + std::fabs(a[2]) + std::fabs(a[5]);
    if (off < 1e-300 || off < 1e-18 * (std::fabs(a[0]) + std::fabs(a[4]) + std::fabs(a[8]))) break;
    for (int p = 0; p < 3; ++p) for (int q = p + 1; q < 3; ++q) {
      if (a[3 * p + q] == 0) continue;
      const double theta = (a[3 * q + q] - a[3 * p + p]) / (2 * a[3 * p + q]);
      const double t = (theta >= 0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1));
      const double c = 1 / std::sqrt(t * t + 1), sn = t * c;
      for (int k = 0; k < 3; ++k) { const double akp = a[3 * k + p], akq = a[3 * k + q]; a[3 * k + p] = c * akp - sn * akq; a[3 * k + q] = sn * akp + c * akq; }
      for (int k = 0; k < 3; ++k) { const double apk = a[3 * p + k], aqk = a[3 * q + k]; a[3 * p + k] = c * apk - sn * aqk; a[3 * q + k] = sn * apk + c * aqk; }
    }
  }
  ev[0] = a[0]; ev[1] = a[4]; ev[2] = a[8];
}
inline bool inv3(const double H[9], double I[9]) {
  const double c00 = H[4] * H[8] - H[5] * H[7], c01 = H[5] * H[6] - H[3] * H[8], c02 = H[3] * H[7] - H[4] * H[6];
  const double det = H[0] * c00 + H[1] * c01 + H[2] * c02;
  if (det == 0 || !std::isfinite(det)) return false;
  I[0] = c00 / det; I[1] = (H[2] * H[7] - H[1] * H[8]) / det; I[2] = (H[1] * H[5] - H[2] * H[4]) / det;
  I[3] = c01 / det; I[4] = (H[0] * H[8] - H[2] * H[6]) / det; I[5] = (H[2] * H[3] - H[0] * H[5]) / det;
  I[6] = c02 / det; I[7] = (H[1] * H[6] - H[0] * H[7]) / det; I[8] = (H[0] * H[4] - H[1] * H[3]) / det;
  return true;
}
}  // namespace detail

// ---- OdometryKeyframeFuser (odometrykeyframefuser.h:66-293, odometrykeyframefuser.cpp:23-259, 397-416, 470-494) -----
// Per-scan orchestration: motion compensation with the previous motion, constant-velocity guess, registration against
// the sliding window of keyframes, sanity check, keyframe insertion, optional covariance by cost sampling.  ROS
// publishing / TF / pose-graph logging are out of scope.  Quirk kept from the reference: the return value of Register() lands in a
// shadowed local (odometrykeyframefuser.cpp:184-186), so the pose in T_vek.back() is used whether or not it succeeded.
struct RadarScan {                         // the members of types.h:93-143 the pose path touches
  Affine3d T;                              // GetPose()
  MapNormalPtr cloud_normal_;
  uint64_t stamp = 0;
};
typedef std::vector<RadarScan> PoseScanVector;

class OdometryKeyframeFuser {
 public:
  class Parameters {
   public:
    std::string cost_type = "P2L";
    weightoption weight_opt = Uniform;
    int submap_scan_size = 3;
    bool weight_intensity_ = false;
    bool use_guess = true, disable_registration = false, soft_constraint = false;
    bool compensate = true, radar_ccw = false;
    bool use_keyframe = true, use_raw_pointcloud = false;
    double res = 3.5;
    double min_keyframe_dist_ = 1.5, min_keyframe_rot_deg_ = 5;
    std::string loss_type_ = "Huber";
    double loss_limit_ = 0.1;
    double covar_scale_ = 1.0;
    double regularization_ = 0.0;
    // covariance by cost sampling (odometrykeyframefuser.h:104-110)
    bool estimate_cov_by_sampling = false;
    double cov_sampling_xy_range = 0.4;
    double cov_sampling_yaw_range = 0.0043625;
    unsigned int cov_sampling_samples_per_axis = 3;
    double cov_sampling_covariance_scaler = 4.0;
  };

  OdometryKeyframeFuser(const Parameters& pars, bool disable_callback = false) : par(pars) {
    (void)disable_callback;
    assert(par.res > 0.05 && par.submap_scan_size >= 1);
    radar_reg = std::make_shared<n_scan_normal_reg>(Str2Cost(par.cost_type), Str2loss(par.loss_type_), par.loss_limit_, par.weight_opt);
    radar_reg->SetD2dPar(par.covar_scale_, par.regularization_);
    cov_current = Matrix6d::Identity();
  }

  // odometrykeyframefuser.cpp:397-416
  void pointcloudCallback(CloudPtr& cloud_filtered, CloudPtr& cloud_filtered_peaks, Affine3d& Tcurr, const uint64_t& t) {
    updated = false;
    processFrame(cloud_filtered, cloud_filtered_peaks, t);
    nr_callbacks_++;
    Tcurr = Tcurrent;
  }
  void pointcloudCallback(CloudPtr& cloud_filtered, CloudPtr& cloud_filtered_peaks, Affine3d& Tcurr, const uint64_t& t, Matrix6d& cov_curr) {
    pointcloudCallback(cloud_filtered, cloud_filtered_peaks, Tcurr, t);
    cov_curr = cov_current;
  }

  // :62-73
  static bool KeyFrameBasedFuse(const Affine3d& diff, bool use_keyframe, double min_keyframe_dist, double min_keyframe_rot_deg) {
    if (!use_keyframe) return true;
    const double yaw = std::atan2(diff(1, 0), diff(1, 1));          // eulerAngles(0,1,2) of a planar rotation, normalised
    const double tn = std::sqrt(diff(0, 3) * diff(0, 3) + diff(1, 3) * diff(1, 3) + diff(2, 3) * diff(2, 3));
    return tn > min_keyframe_dist || std::fabs(yaw) > (min_keyframe_rot_deg * M_PI / 180.0);
  }
  // :76-94
  static bool AccelerationVelocitySanityCheck(const Affine3d& Tmot_prev, const Affine3d& Tmot_curr) {
    const double dt = 0.25, vel_limit = 200, acc_limit = 200;
    const double vx = Tmot_curr(0, 3) / dt, vy = Tmot_curr(1, 3) / dt;
    const double ax = (Tmot_curr(0, 3) - Tmot_prev(0, 3)) / (dt * dt), ay = (Tmot_curr(1, 3) - Tmot_prev(1, 3)) / (dt * dt);
    if (std::sqrt(ax * ax + ay * ay) > acc_limit) return false;
    if (std::sqrt(vx * vx + vy * vy) > vel_limit) return false;
    return true;
  }

  // odometrykeyframefuser.cpp:261-380.  The samples_per_axis^3 cost samples (yaw-major, then x, then y, like the
  // reference's loops) are evaluated in ONE cfear_get_cost_batch launch; the quadric fit, convexity check and scaling
  // follow the reference (its bdcSvd least-squares solve is a Householder QR here).  Writing the samples to a csv file
  // (cov_samples_to_file_as_well) is not mirrored.
  bool approximateCovarianceBySampling(std::vector<MapNormalPtr>& scans_vek, const std::vector<Affine3d>& T_vek, Matrix6d& cov_sampled) {
    const Affine3d T_best_guess = T_vek.back();
    const double xy_sample_range = par.cov_sampling_xy_range * 0.5, theta_range = par.cov_sampling_yaw_range * 0.5;
    const unsigned int n = par.cov_sampling_samples_per_axis;
    const std::vector<double> xy_samples = linspace(-xy_sample_range, xy_sample_range, (int)n);
    const std::vector<double> theta_samples = linspace(-theta_range, theta_range, (int)n);
    std::vector<std::vector<Affine3d> > samples;
    std::vector<double> sx, sy, syaw;
    const double yaw0 = std::atan2(T_best_guess(1, 0), T_best_guess(0, 0));
    for (unsigned int it = 0; it < n; ++it)
      for (unsigned int ix = 0; ix < n; ++ix)
        for (unsigned int iy = 0; iy < n; ++iy) {
          std::vector<Affine3d> T(T_vek);
          T.back() = vectorToAffine3d(T_best_guess(0, 3) + xy_samples[ix], T_best_guess(1, 3) + xy_samples[iy], yaw0 + theta_samples[it]);
          samples.push_back(T);
          sx.push_back(xy_samples[ix]); sy.push_back(xy_samples[iy]); syaw.push_back(theta_samples[it]);
        }
    std::vector<double> cost; std::vector<int> nres; std::vector<bool> ok;
    radar_reg->GetCostBatch(scans_vek, samples, cost, nres, ok);
    const int m = (int)samples.size();
    if (m < 10) return false;
    // f(x,y,z) = a x^2 + b y^2 + c z^2 + d xy + e yz + f zx + g x + h y + i z + j
    std::vector<double> A((size_t)m * 10);
    for (int r = 0; r < m; ++r) {
      const double x = sx[r], y = sy[r], z = syaw[r];
      const double row[10] = {x * x, y * y, z * z, x * y, y * z, z * x, x, y, z, 1.0};
      for (int k = 0; k < 10; ++k) A[(size_t)r * 10 + k] = row[k];
    }
    double q[10];
    if (!detail::lstsq_qr(A, m, 10, cost, q)) return false;
    const double H[9] = {2 * q[0], q[3], q[5], q[3], 2 * q[1], q[4], q[5], q[4], 2 * q[2]};
    double ev[3];
    detail::eig3_sym(H, ev);
    if (!(ev[0] > 0.0) || !(ev[1] > 0.0) || !(ev[2] > 0.0)) return false;   // not convex: sampling not used for this scan
    double score_scale = 1.0, Hi[9];
    if (!radar_reg->GetCovarianceScaler(score_scale) || !detail::inv3(H, Hi)) return false;
    const double f = 2.0 * score_scale * par.cov_sampling_covariance_scaler;
    cov_sampled = Matrix6d::Identity();
    cov_sampled(0, 0) = f * Hi[0]; cov_sampled(0, 1) = f * Hi[1]; cov_sampled(1, 0) = f * Hi[3]; cov_sampled(1, 1) = f * Hi[4];
    cov_sampled(5, 5) = f * Hi[8];
    cov_sampled(0, 5) = f * Hi[2]; cov_sampled(1, 5) = f * Hi[5]; cov_sampled(5, 0) = f * Hi[6]; cov_sampled(5, 1) = f * Hi[7];
    return true;
  }
  // odometrykeyframefuser.cpp:497-523
  static std::vector<double> linspace(double start, double end, int num) {
    std::vector<double> v;
    if (num == 0) return v;
    if (num == 1) { v.push_back(start); return v; }
    const double delta = (end - start) / (num - 1);
    for (int i = 0; i < num - 1; ++i) v.push_back(start + delta * i);
    v.push_back(end);
    return v;
  }

  bool updated = false;
  size_t nr_callbacks_ = 0, frame_nr_ = 0;
  double distance_traveled = 0;
  Affine3d Tcurrent, T_prev, Tmot;
  Matrix6d cov_current;
  PoseScanVector keyframes_;
  std::shared_ptr<n_scan_normal_reg> radar_reg;

 private:
  // :143-259
  void processFrame(CloudPtr& cloud, CloudPtr& cloud_peaks, const uint64_t& t) {
    const Affine3d TprevMot(Tmot);
    if (par.compensate) {
      if (cloud_peaks) Compensate(*cloud, *cloud_peaks, TprevMot, par.radar_ccw);     // :148-149, one round trip for both
      else Compensate(*cloud, TprevMot, par.radar_ccw);
    }
    std::vector<Matrix6d> cov_vek;
    std::vector<MapNormalPtr> scans_vek;
    std::vector<Affine3d> T_vek;
    MapNormalPtr Pcurrent(new MapPointNormal(cloud, (float)par.res, Vector2d(0, 0), par.weight_intensity_, par.use_raw_pointcloud));
    const Affine3d Tguess = par.use_guess ? T_prev * TprevMot : T_prev;
    if (keyframes_.empty()) {                                        // :171-177
      RadarScan scan; scan.T = Affine3d::Identity(); scan.cloud_normal_ = Pcurrent; scan.stamp = t;
      updated = true;
      AddToReference(keyframes_, scan, (size_t)par.submap_scan_size);
      return;
    }
    for (size_t i = 0; i < keyframes_.size(); ++i) {                 // FormatScans :478-494
      cov_vek.push_back(Matrix6d::Identity()); scans_vek.push_back(keyframes_[i].cloud_normal_); T_vek.push_back(keyframes_[i].T);
    }
    cov_vek.push_back(Matrix6d::Identity()); scans_vek.push_back(Pcurrent); T_vek.push_back(Tguess);
    if (!par.disable_registration) (void)radar_reg->Register(scans_vek, T_vek, cov_vek, par.soft_constraint);   // result ignored (:184-186)
    Tcurrent = T_vek.back();
    cov_current = cov_vek.back();
    const Affine3d Tmot_current = T_prev.inverse() * Tcurrent;
    if (!AccelerationVelocitySanityCheck(Tmot, Tmot_current)) Tcurrent = Tguess;    // :197-199
    Tmot = T_prev.inverse() * Tcurrent;
    if (par.estimate_cov_by_sampling && !par.disable_registration) {              // :203-208
      Matrix6d cov_sampled;
      if (approximateCovarianceBySampling(scans_vek, T_vek, cov_sampled)) { cov_current = cov_sampled; cov_vek.back() = cov_sampled; }
    }
    const Affine3d Tkeydiff = keyframes_.back().T.inverse() * Tcurrent;
    const bool fuse = KeyFrameBasedFuse(Tkeydiff, par.use_keyframe, par.min_keyframe_dist_, par.min_keyframe_rot_deg_);
    if (fuse) {                                                      // `success && fuse` with success always true
      distance_traveled += std::sqrt(Tkeydiff(0, 3) * Tkeydiff(0, 3) + Tkeydiff(1, 3) * Tkeydiff(1, 3));
      frame_nr_++;
      RadarScan scan; scan.T = Tcurrent; scan.cloud_normal_ = Pcurrent; scan.stamp = t;
      AddToReference(keyframes_, scan, (size_t)par.submap_scan_size);
      updated = true;
    }
    T_prev = Tcurrent;
  }
  // :470-476
  static void AddToReference(PoseScanVector& reference, RadarScan& scan, size_t submap_scan_size) {
    reference.push_back(scan);
    if (reference.size() > submap_scan_size) reference.erase(reference.begin());
  }
  Parameters par;
};

// KITTI-format pose row (eval_trajectory.cpp:169-183 via MatToString types.cpp:64-73): the 3x4 row-major matrix,
// fixed notation with 6 decimals (std::fixed << std::showpoint), space separated.
inline std::string MatToString(const Affine3d& T) {
  std::ostringstream s;
  s << std::fixed << std::showpoint;
  s.precision(6);
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 4; ++c) { s << T(r, c); if (!(r == 2 && c == 3)) s << " "; }
  return s.str();
}

// ---- trajectory files (eval_trajectory.cpp:169-233): est/NN.txt (KITTI), tum_NN.txt, cov_NN.txt ---------------------
struct PoseStamped { Affine3d pose; Matrix6d cov; uint32_t sec = 0, nsec = 0; };
typedef std::vector<PoseStamped> poseStampedVector;

// Eigen::Quaterniond(rotation matrix) -- QuaternionBase::operator=(MatrixBase), same branch structure
inline void RotationToQuaternion(const Affine3d& T, double& qx, double& qy, double& qz, double& qw) {
  double q[3];
  double t = T(0, 0) + T(1, 1) + T(2, 2);
  if (t > 0) {
    t = std::sqrt(t + 1.0); qw = 0.5 * t; t = 0.5 / t;
    q[0] = (T(2, 1) - T(1, 2)) * t; q[1] = (T(0, 2) - T(2, 0)) * t; q[2] = (T(1, 0) - T(0, 1)) * t;
  } else {
    int i = 0;
    if (T(1, 1) > T(0, 0)) i = 1;
    if (T(2, 2) > T(i, i)) i = 2;
    const int j = (i + 1) % 3, k = (j + 1) % 3;
    t = std::sqrt(T(i, i) - T(j, j) - T(k, k) + 1.0);
    q[i] = 0.5 * t; t = 0.5 / t;
    qw = (T(k, j) - T(j, k)) * t; q[j] = (T(j, i) + T(i, j)) * t; q[k] = (T(k, i) + T(i, k)) * t;
  }
  qx = q[0]; qy = q[1]; qz = q[2];
}

class EvalTrajectory {
 public:
  void CallbackESTEigen(const Affine3d& pose, const Matrix6d& cov, uint32_t sec, uint32_t nsec) {   // eval_trajectory.cpp:56-62
    PoseStamped p; p.pose = pose; p.cov = cov; p.sec = sec; p.nsec = nsec; est_vek.push_back(p);
  }
  // eval_trajectory.cpp:74-143 (DatasetToSequence): the Oxford Radar RobotCar sequence names in evaluation order ->
  // "00" .. "31"; anything else maps to "01" like the reference's fall-through.
  static std::string SequenceToFileName(const std::string& sequence) {
    static const char* const names[] = {
        "2019-01-10-11-46-21-radar-oxford-10k", "2019-01-10-12-32-52-radar-oxford-10k", "2019-01-10-14-02-34-radar-oxford-10k",
        "2019-01-10-14-36-48-radar-oxford-10k-partial", "2019-01-10-14-50-05-radar-oxford-10k", "2019-01-10-15-19-41-radar-oxford-10k",
        "2019-01-11-12-26-55-radar-oxford-10k", "2019-01-11-13-24-51-radar-oxford-10k", "2019-01-11-14-02-26-radar-oxford-10k",
        "2019-01-11-14-37-14-radar-oxford-10k", "2019-01-14-12-05-52-radar-oxford-10k", "2019-01-14-12-41-28-radar-oxford-10k",
        "2019-01-14-13-38-21-radar-oxford-10k", "2019-01-14-14-15-12-radar-oxford-10k", "2019-01-14-14-48-55-radar-oxford-10k",
        "2019-01-15-12-01-32-radar-oxford-10k", "2019-01-15-12-52-32-radar-oxford-10k-partial", "2019-01-15-13-06-37-radar-oxford-10k",
        "2019-01-15-13-53-14-radar-oxford-10k", "2019-01-15-14-24-38-radar-oxford-10k", "2019-01-16-11-53-11-radar-oxford-10k",
        "2019-01-16-13-09-37-radar-oxford-10k", "2019-01-16-13-42-28-radar-oxford-10k", "2019-01-16-14-15-33-radar-oxford-10k",
        "2019-01-17-11-46-31-radar-oxford-10k", "2019-01-17-12-48-25-radar-oxford-10k", "2019-01-17-13-26-39-radar-oxford-10k",
        "2019-01-17-14-03-00-radar-oxford-10k", "2019-01-18-12-42-34-radar-oxford-10k", "2019-01-18-14-14-42-radar-oxford-10k",
        "2019-01-18-14-46-59-radar-oxford-10k", "2019-01-18-15-20-12-radar-oxford-10k"};
    for (int i = 0; i < 32; ++i)
      if (sequence == names[i]) { char b[8]; snprintf(b, sizeof b, "%02d", i); return b; }
    return "01";
  }
  static void Write(const std::string& path, const poseStampedVector& v) {                          // :169-183
    std::ofstream f(path);
    for (size_t i = 0; i < v.size(); ++i) f << MatToString(v[i].pose) << std::endl;
  }
  static void WriteTUM(const std::string& path, const poseStampedVector& v) {                        // :185-212
    std::ofstream f(path);
    for (size_t i = 0; i < v.size(); ++i) {
      f << v[i].sec << "." << std::setfill('0') << std::setw(9) << v[i].nsec << " " << std::setw(0);
      f << std::fixed << std::setprecision(4);
      f << v[i].pose(0, 3) << " " << v[i].pose(1, 3) << " " << v[i].pose(2, 3) << " ";
      double qx, qy, qz, qw; RotationToQuaternion(v[i].pose, qx, qy, qz, qw);
      f << std::defaultfloat;
      f << qx << " " << qy << " " << qz << " " << qw;
      f << std::endl;
    }
  }
  static void WriteCov(const std::string& path, const poseStampedVector& v) {                        // :214-233
    std::ofstream f(path);
    for (size_t i = 0; i < v.size(); ++i) {
      f << v[i].sec << "." << std::setfill('0') << std::setw(9) << v[i].nsec << " " << std::setw(0);
      for (int k = 0; k < 36; ++k) { f << v[i].cov.m[k]; if (k != 35) f << " "; }   // Eigen IOFormat(StreamPrecision, DontAlignCols, " ", " ")
      f << std::endl;
    }
  }
  poseStampedVector est_vek;
};

}  // namespace CFEAR_Radarodometry
#endif  // CFEAR_B200_HPP_
