// offline_odometry on the B200 path: the radarReader loop of the reference (src/offline_odometry.cpp:56-131) with its
// command-line options (src/offline_odometry.cpp:153-288), ROS-free.  The rosbag of sensor_msgs::Image messages is
// replaced by a raw frame file (what cfear_radarodometry_code_public_b200/io.py writes from Oxford / MulRan PNGs):
//
//   "CFRS" | int32 n_frames, azimuths, range_bins | n_frames x { uint64 stamp_ns | azimuths*range_bins uint8 }
//
// Per frame:  driver.CallbackOffline(image, cloud, peaks);  fuser.pointcloudCallback(cloud, peaks, Tcurrent, stamp, cov);
//             eval.CallbackESTEigen(Tcurrent, cov, stamp)   -- then est/<NN>.txt (KITTI rows), optionally TUM / cov files.
//
//   g++ -std=c++14 -O2 -I include examples/offline_odometry.cpp -o offline_odometry
//       -L cfear_radarodometry_code_public_b200 -lcfear_b200 -Wl,-rpath,cfear_radarodometry_code_public_b200
//   ./offline_odometry --frames seq.cfrs --est_directory out --cost_type P2L --submap_scan_size 4 --res 3 --k_strongest 12
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <map>
#include <string>
#include <vector>

#include "cfear_b200.hpp"

using namespace CFEAR_Radarodometry;

static std::map<std::string, std::string> ParseArgs(int argc, char** argv) {
  std::map<std::string, std::string> a;
  for (int i = 1; i < argc; ++i) {
    std::string k = argv[i];
    if (k.rfind("--", 0) != 0) { std::cerr << "unexpected argument " << k << std::endl; exit(2); }
    k = k.substr(2);
    const size_t eq = k.find('=');
    if (eq != std::string::npos) { a[k.substr(0, eq)] = k.substr(eq + 1); continue; }
    if (i + 1 < argc && std::string(argv[i + 1]).rfind("--", 0) != 0) a[k] = argv[++i];
    else a[k] = "true";
  }
  return a;
}
static bool ToBool(const std::string& s) { return s == "true" || s == "1" || s == "True"; }

static int run(int argc, char** argv) {
  std::map<std::string, std::string> vm = ParseArgs(argc, argv);
  if (vm.count("help") || !vm.count("frames")) {
    std::cout << "offline_odometry --frames <file.cfrs> [--est_directory DIR] [--sequence NAME] [--res 3.5] [--range-res 0.0438]\n"
                 "  [--min_distance 2.5] [--submap_scan_size 3] [--weight_intensity true] [--k_strongest 12] [--z-min 65]\n"
                 "  [--registered_min_keyframe_dist 1.5] [--radar_ccw false] [--disable_compensate false] [--cost_type P2L]\n"
                 "  [--loss_type Huber] [--loss_limit 0.1] [--covar_scale 1] [--regularization 1] [--weight_option 0]\n"
                 "  [--covar_sampling false] [--covar_XY_sample_range 0.4] [--covar_yaw_sample_range 0.0043625]\n"
                 "  [--covar_samples_per_axis 3] [--covar_sampling_scale 4] [--filter-type kstrong|CA-CFAR]\n"
                 "  [--false-alarm-rate 0.01] [--nb-guard-cells 10] [--nb-window-cells 10] [--tum true] [--cov true]\n";
    return vm.count("help") ? 0 : 2;
  }
  auto num = [&](const char* k, double d) { return vm.count(k) ? atof(vm[k].c_str()) : d; };
  auto str = [&](const char* k, const char* d) { return vm.count(k) ? vm[k] : std::string(d); };
  auto flag = [&](const char* k, bool d) { return vm.count(k) ? ToBool(vm[k]) : d; };

  // defaults and option -> parameter mapping of src/offline_odometry.cpp:153-288
  OdometryKeyframeFuser::Parameters odom_pars;
  radarDriver::Parameters rad_pars;
  odom_pars.res = num("res", 3.5);
  rad_pars.range_res = (float)num("range-res", 0.0438);
  rad_pars.min_distance = (float)num("min_distance", 2.5);
  rad_pars.max_distance = (float)num("max_distance", 200);
  odom_pars.submap_scan_size = (int)num("submap_scan_size", 3);
  odom_pars.weight_intensity_ = flag("weight_intensity", true);
  rad_pars.k_strongest = (int)num("k_strongest", 12);
  odom_pars.min_keyframe_dist_ = num("registered_min_keyframe_dist", 1.5);
  rad_pars.z_min = (float)num("z-min", 65);
  odom_pars.radar_ccw = flag("radar_ccw", false);
  odom_pars.soft_constraint = flag("soft_constraint", false);
  odom_pars.compensate = !flag("disable_compensate", false);
  odom_pars.cost_type = str("cost_type", "P2L");
  odom_pars.loss_type_ = str("loss_type", "Huber");
  odom_pars.loss_limit_ = num("loss_limit", 0.1);
  odom_pars.covar_scale_ = num("covar_scale", 1);
  odom_pars.regularization_ = num("regularization", 1);
  odom_pars.weight_opt = (weightoption)(int)num("weight_option", 0);
  odom_pars.estimate_cov_by_sampling = flag("covar_sampling", false);
  odom_pars.cov_sampling_xy_range = num("covar_XY_sample_range", 0.4);
  odom_pars.cov_sampling_yaw_range = num("covar_yaw_sample_range", 0.0043625);
  odom_pars.cov_sampling_samples_per_axis = (unsigned)num("covar_samples_per_axis", 3);
  odom_pars.cov_sampling_covariance_scaler = num("covar_sampling_scale", 4);
  rad_pars.dataset = str("dataset", "oxford");
  rad_pars.filter_type_ = str("filter-type", "kstrong") == "CA-CFAR" ? CACFAR : kstrong;
  rad_pars.false_alarm_rate = (float)num("false-alarm-rate", 0.01);
  rad_pars.nb_guard_cells = (int)num("nb-guard-cells", 10);
  rad_pars.window_size = (int)num("nb-window-cells", 10);
  const std::string est_dir = str("est_directory", ".");
  const std::string sequence = str("sequence", "2019-01-10-12-32-52-radar-oxford-10k");

  FILE* f = fopen(vm["frames"].c_str(), "rb");
  if (!f) { std::cerr << "cannot open " << vm["frames"] << std::endl; return 3; }
  char magic[4]; int32_t hdr[3];
  if (fread(magic, 1, 4, f) != 4 || memcmp(magic, "CFRS", 4) != 0 || fread(hdr, 4, 3, f) != 3) { std::cerr << "not a CFRS frame file" << std::endl; return 4; }
  const int n = hdr[0], A = hdr[1], R = hdr[2];
  rad_pars.azimuths = A;
  std::cout << "Loading frames from: " << vm["frames"] << " (" << n << " x " << A << " x " << R << ")" << std::endl;
  std::cout << rad_pars.ToString();

  // The device context is shaped by the data and the options: image geometry from the frame header, k, and room for the
  // keyframe window (+ current scan + the sets MapPointNormal objects of the last frames still hold).
  cfear_config cfg;
  cfear_default_config(&cfg);
  cfg.azimuths = A; cfg.range_bins = R; cfg.k_strongest = rad_pars.k_strongest;
  cfg.max_keyframes = std::max(odom_pars.submap_scan_size, 1);
  cfg.max_cellsets = cfg.max_keyframes + 8;
  cfg.max_batch = std::max(1, (int)(odom_pars.cov_sampling_samples_per_axis * odom_pars.cov_sampling_samples_per_axis *
                                     odom_pars.cov_sampling_samples_per_axis));
  Backend::Configure(cfg);

  radarDriver driver(rad_pars, true);
  OdometryKeyframeFuser fuser(odom_pars, true);
  EvalTrajectory eval;
  // the frame buffer is this program's own, so it is page-locked: the image then goes to the device as one DMA at link
  // rate instead of being staged by the driver (0.1 ms of a 0.6 ms frame); pageable memory works as well
  const size_t img_bytes = (size_t)A * R;
  std::vector<uint8_t> img_pageable;
  uint8_t* img_data = static_cast<uint8_t*>(cfear_alloc_pinned(img_bytes));
  const bool img_pinned = img_data != nullptr;
  if (!img_pinned) { img_pageable.resize(img_bytes); img_data = img_pageable.data(); }
  double tot = 0;
  for (int frame = 0; frame < n; ++frame) {
    uint64_t stamp = 0;
    if (fread(&stamp, 8, 1, f) != 1 || fread(img_data, 1, img_bytes, f) != img_bytes) { std::cerr << "truncated frame file" << std::endl; return 4; }
    const auto t0 = std::chrono::steady_clock::now();
    PolarImage pim; pim.rows = A; pim.cols = R; pim.data = img_data; pim.stamp = stamp;
    CloudPtr cloud_filtered, cloud_filtered_peaks;
    driver.CallbackOffline(pim, cloud_filtered, cloud_filtered_peaks);                                   // offline_odometry.cpp:103
    Affine3d Tcurrent; Matrix6d cov_current;
    fuser.pointcloudCallback(cloud_filtered, cloud_filtered_peaks, Tcurrent, stamp, cov_current);         // :108
    eval.CallbackESTEigen(Tcurrent, cov_current, (uint32_t)(stamp / 1000000000ull), (uint32_t)(stamp % 1000000000ull));   // :117
    const double d = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    tot += d;
    std::cout << "Frame: " << frame << ", dur: " << d << ", avg: " << (frame + 1) / tot << std::endl;          // :124
  }
  fclose(f);
  if (img_pinned) cfear_free_pinned(img_data);
  // EvalTrajectory::Save (eval_trajectory.cpp:145-167): <est_directory>/<NN>.txt with NN from the sequence name
  const std::string nn = EvalTrajectory::SequenceToFileName(sequence);
  EvalTrajectory::Write(est_dir + "/" + nn + ".txt", eval.est_vek);
  if (flag("tum", false)) EvalTrajectory::WriteTUM(est_dir + "/" + nn + "_tum.txt", eval.est_vek);
  if (flag("cov", false)) EvalTrajectory::WriteCov(est_dir + "/" + nn + "_cov.txt", eval.est_vek);
  std::cout << "Trajectory saved to: " << est_dir << "/" << nn << ".txt (" << eval.est_vek.size() << " poses, "
            << fuser.frame_nr_ << " keyframes, " << fuser.distance_traveled << " m)" << std::endl;
  return 0;
}

int main(int argc, char** argv) {
  try {
    return run(argc, argv);
  } catch (const std::exception& e) {
    std::cerr << "offline_odometry: " << e.what() << std::endl;
    return 1;
  }
}
