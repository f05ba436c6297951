// Drives the C++ host mirror (include/cfear_b200.hpp) exactly the way offline_odometry / OdometryKeyframeFuser
// drive the reference classes (src/offline_odometry.cpp:103-108, odometrykeyframefuser.cpp:146-196):
//   driver.CallbackOffline(img, cloud, peaks) -> Compensate -> new MapPointNormal(...) -> radar_reg->Register(...)
// usage: mirror_test <in.bin> <out.bin>
//   in : int32 nscan, A, R; float radius; int32 cost, loss, weight_opt; double regularization; int32 flags (1 = soft
//        constraints, 2 = keep the association tables); nscan images (A*R u8); nscan poses (x,y,yaw f64, last = guess); mot (3 f64)
//   out: int32 ok, itr, nres, ncells_last, npts_last; pose (3 f64); score (f64); cov 36 f64; closest idx of (10, 0);
//        int32 n_assoc (all scan pairs); f64 sum of target indices, sum of source indices, sum of sim_dir_, sum of GetWeight();
//        f64 GetCellRelTimeStamp(0, false), GetCellRelTimeStamp(0, true)
#include <cstdio>
#include <vector>

#include "cfear_b200.hpp"

using namespace CFEAR_Radarodometry;

int main(int argc, char** argv) {
  if (argc < 3) return 2;
  FILE* f = fopen(argv[1], "rb");
  if (!f) return 3;
  int32_t hdr[3]; float radius; int32_t opt[3]; double regularization; int32_t flags;
  if (fread(hdr, 4, 3, f) != 3 || fread(&radius, 4, 1, f) != 1 || fread(opt, 4, 3, f) != 3 || fread(&regularization, 8, 1, f) != 1 ||
      fread(&flags, 4, 1, f) != 1) return 4;
  const int ns = hdr[0], A = hdr[1], R = hdr[2];
  std::vector<std::vector<uint8_t>> imgs(ns, std::vector<uint8_t>((size_t)A * R));
  for (auto& im : imgs) if (fread(im.data(), 1, im.size(), f) != im.size()) return 4;
  std::vector<double> poses(3 * ns); double mot[3];
  if (fread(poses.data(), 8, poses.size(), f) != poses.size() || fread(mot, 8, 3, f) != 3) return 4;
  fclose(f);

  radarDriver::Parameters rad_pars;               // defaults: z_min 60, k 12, range_res 0.0438, min_distance 2.5
  radarDriver driver(rad_pars, true);
  n_scan_normal_reg radar_reg((cost_metric)opt[0], (loss_type)opt[1], 0.1, (weightoption)opt[2]);
  radar_reg.SetD2dPar(1.0, regularization);

  std::vector<MapNormalPtr> scans_vek;
  std::vector<Affine3d> T_vek;
  std::vector<Matrix6d> cov_vek;
  size_t npts_last = 0;
  for (int i = 0; i < ns; ++i) {
    PolarImage img; img.rows = A; img.cols = R; img.data = imgs[i].data();
    CloudPtr cloud, cloud_peaks;
    driver.CallbackOffline(img, cloud, cloud_peaks);
    if (i == ns - 1) Compensate(*cloud, vectorToAffine3d(mot[0], mot[1], mot[2]), false);   // only the current scan moves
    npts_last = cloud->size();
    scans_vek.push_back(MapNormalPtr(new MapPointNormal(cloud, radius, Vector2d(0, 0), true, false)));
    T_vek.push_back(vectorToAffine3d(poses[3 * i], poses[3 * i + 1], poses[3 * i + 2]));
    cov_vek.push_back(Matrix6d::Identity());
  }
  radar_reg.keep_associations_ = (flags & 2) != 0;
  if (flags & 1) { cov_vek.back()(0, 0) = 0.04; cov_vek.back()(1, 1) = 0.09; cov_vek.back()(0, 1) = cov_vek.back()(1, 0) = 0.01; cov_vek.back()(5, 5) = 0.0004;
                   cov_vek.back()(0, 5) = cov_vek.back()(5, 0) = 0.001; }
  const bool ok = radar_reg.Register(scans_vek, T_vek, cov_vek, (flags & 1) != 0);
  std::vector<double> par;
  Affine3dToVectorXYeZ(T_vek.back(), par);
  std::vector<int> near = scans_vek.back()->GetClosestIdx(Vector2d(10.0, 0.0), 50.0);

  FILE* o = fopen(argv[2], "wb");
  int32_t ih[5] = {ok ? 1 : 0, (int32_t)radar_reg.itr_, radar_reg.summary_.num_residuals, (int32_t)scans_vek.back()->GetSize(), (int32_t)npts_last};
  fwrite(ih, 4, 5, o);
  fwrite(par.data(), 8, 3, o);
  const double score = radar_reg.getScore();
  fwrite(&score, 8, 1, o);
  fwrite(cov_vek.back().m, 8, 36, o);
  int32_t ni = near.empty() ? -1 : near[0];
  fwrite(&ni, 4, 1, o);
  int32_t n_assoc = 0; double sums[4] = {0, 0, 0, 0};
  for (auto& kv : radar_reg.scan_associations_) {
    std::vector<Registration::Weights>& w = radar_reg.weight_associations_[kv.first];
    if (w.size() != kv.second.size() || kv.first.second != ns - 1) return 5;
    for (size_t i = 0; i < kv.second.size(); ++i) {
      ++n_assoc; sums[0] += kv.second[i].first; sums[1] += kv.second[i].second; sums[2] += w[i].sim_dir_;
      sums[3] += w[i].GetWeight(radar_reg.weight_opt_);
    }
  }
  fwrite(&n_assoc, 4, 1, o); fwrite(sums, 8, 4, o);
  const double ts[2] = {scans_vek.back()->GetCellRelTimeStamp(0, false), scans_vek.back()->GetCellRelTimeStamp(0, true)};
  fwrite(ts, 8, 2, o);
  fclose(o);
  printf("mirror_test: ok=%d itr=%zu residuals=%d cells=%zu pose=(%.6f %.6f %.6f)\n", (int)ok, radar_reg.itr_, radar_reg.summary_.num_residuals,
         scans_vek.back()->GetSize(), par[0], par[1], par[2]);
  return 0;
}
