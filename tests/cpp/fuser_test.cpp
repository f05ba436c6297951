// Sequence replay through the C++ mirror exactly like radarReader's loop (src/offline_odometry.cpp:73-127):
//   driver.CallbackOffline(img, cloud, peaks);  fuser.pointcloudCallback(cloud, peaks, Tcurrent, stamp, cov);
// usage: fuser_test <in.bin> <out.bin> <est.txt>
//   in : int32 nscans, A, R, submap_scan_size, weight_opt, weight_intensity; float res; char cost[4] (P2L/P2D/P2P);
//        double regularization; images
//   out: per scan 3 f64 (x, y, yaw) + int32 updated ; est.txt: KITTI rows (eval_trajectory.cpp:169-183)
#include <cstdio>
#include <cstring>
#include <fstream>
#include <vector>

#include "cfear_b200.hpp"

using namespace CFEAR_Radarodometry;

int main(int argc, char** argv) {
  if (argc < 4) return 2;
  FILE* f = fopen(argv[1], "rb");
  if (!f) return 3;
  int32_t h[6]; float res; char cost[4]; double regularization;
  if (fread(h, 4, 6, f) != 6 || fread(&res, 4, 1, f) != 1 || fread(cost, 1, 4, f) != 4 || fread(&regularization, 8, 1, f) != 1) return 4;
  const int n = h[0], A = h[1], R = h[2];
  std::vector<uint8_t> img((size_t)A * R);

  radarDriver::Parameters rad_pars;
  OdometryKeyframeFuser::Parameters odom_pars;
  odom_pars.submap_scan_size = h[3]; odom_pars.weight_opt = (weightoption)h[4]; odom_pars.weight_intensity_ = h[5] != 0;
  odom_pars.res = res; odom_pars.cost_type = std::string(cost, 3); odom_pars.regularization_ = regularization;
  radarDriver driver(rad_pars, true);
  OdometryKeyframeFuser fuser(odom_pars, true);

  FILE* o = fopen(argv[2], "wb");
  std::ofstream est(argv[3]);
  for (int i = 0; i < n; ++i) {
    if (fread(img.data(), 1, img.size(), f) != img.size()) return 4;
    PolarImage pim; pim.rows = A; pim.cols = R; pim.data = img.data(); pim.stamp = (uint64_t)i;
    CloudPtr cloud, cloud_peaks;
    driver.CallbackOffline(pim, cloud, cloud_peaks);
    Affine3d Tcurrent; Matrix6d cov_current;
    fuser.pointcloudCallback(cloud, cloud_peaks, Tcurrent, pim.stamp, cov_current);
    std::vector<double> par; Affine3dToVectorXYeZ(Tcurrent, par);
    fwrite(par.data(), 8, 3, o);
    int32_t up = fuser.updated ? 1 : 0; fwrite(&up, 4, 1, o);
    est << MatToString(Tcurrent) << std::endl;
  }
  fclose(o); fclose(f);
  printf("fuser_test: %d scans, %zu keyframes in window, distance %.3f\n", n, fuser.keyframes_.size(), fuser.distance_traveled);
  return 0;
}
