// OdometryKeyframeFuser::approximateCovarianceBySampling of the mirror, driven like processFrame drives it
// (odometrykeyframefuser.cpp:186-208): Register, then sample the cost around the registered pose.
// usage: covsample_test <in.bin> <out.bin>   (input format of mirror_test)
//   out: int32 reg_ok, cov_ok, nres; final_cost (f64); pose (3 f64); sampled cov 36 f64; cost at the registered pose (f64)
#include <cstdio>
#include <vector>

#include "cfear_b200.hpp"

using namespace CFEAR_Radarodometry;

int main(int argc, char** argv) {
  if (argc < 3) return 2;
  FILE* f = fopen(argv[1], "rb");
  if (!f) return 3;
  int32_t hdr[3]; float radius; int32_t opt[3]; double regularization;
  if (fread(hdr, 4, 3, f) != 3 || fread(&radius, 4, 1, f) != 1 || fread(opt, 4, 3, f) != 3 || fread(&regularization, 8, 1, f) != 1) return 4;
  const int ns = hdr[0], A = hdr[1], R = hdr[2];
  std::vector<std::vector<uint8_t>> imgs(ns, std::vector<uint8_t>((size_t)A * R));
  for (auto& im : imgs) if (fread(im.data(), 1, im.size(), f) != im.size()) return 4;
  std::vector<double> poses(3 * ns); double mot[3];
  if (fread(poses.data(), 8, poses.size(), f) != poses.size() || fread(mot, 8, 3, f) != 3) return 4;
  fclose(f);

  radarDriver::Parameters rad_pars;
  radarDriver driver(rad_pars, true);
  OdometryKeyframeFuser::Parameters par;
  par.cost_type = opt[0] == 0 ? "P2P" : (opt[0] == 1 ? "P2L" : "P2D");
  par.loss_type_ = "Huber"; par.loss_limit_ = 0.1; par.weight_opt = (weightoption)opt[2];
  par.covar_scale_ = 1.0; par.regularization_ = regularization; par.res = radius;
  par.estimate_cov_by_sampling = true;
  OdometryKeyframeFuser fuser(par, true);

  std::vector<MapNormalPtr> scans_vek;
  std::vector<Affine3d> T_vek;
  std::vector<Matrix6d> cov_vek;
  for (int i = 0; i < ns; ++i) {
    PolarImage img; img.rows = A; img.cols = R; img.data = imgs[i].data();
    CloudPtr cloud, cloud_peaks;
    driver.CallbackOffline(img, cloud, cloud_peaks);
    if (i == ns - 1) Compensate(*cloud, vectorToAffine3d(mot[0], mot[1], mot[2]), false);
    scans_vek.push_back(MapNormalPtr(new MapPointNormal(cloud, radius, Vector2d(0, 0), true, false)));
    T_vek.push_back(vectorToAffine3d(poses[3 * i], poses[3 * i + 1], poses[3 * i + 2]));
    cov_vek.push_back(Matrix6d::Identity());
  }
  const bool ok = fuser.radar_reg->Register(scans_vek, T_vek, cov_vek, false);
  const double final_cost = fuser.radar_reg->summary_.final_cost;
  const int32_t nres = fuser.radar_reg->summary_.num_residuals;
  Matrix6d cov_sampled;
  const bool cov_ok = fuser.approximateCovarianceBySampling(scans_vek, T_vek, cov_sampled);
  double cost_here = 0; std::vector<double> residuals;
  fuser.radar_reg->GetCost(scans_vek, T_vek, cost_here, residuals);
  std::vector<double> par_out;
  Affine3dToVectorXYeZ(T_vek.back(), par_out);

  FILE* o = fopen(argv[2], "wb");
  int32_t ih[3] = {ok ? 1 : 0, cov_ok ? 1 : 0, nres};
  fwrite(ih, 4, 3, o); fwrite(&final_cost, 8, 1, o); fwrite(par_out.data(), 8, 3, o); fwrite(cov_sampled.m, 8, 36, o); fwrite(&cost_here, 8, 1, o);
  fclose(o);
  printf("covsample_test: reg=%d cov=%d residuals=%d cost=%.6f (%zu residual slots)\n", (int)ok, (int)cov_ok, nres, cost_here, residuals.size());
  return 0;
}
