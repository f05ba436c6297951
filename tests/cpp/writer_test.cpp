// Trajectory file formats of the mirror (no GPU work): writes est / tum / cov files for two poses.
#include "cfear_b200.hpp"
using namespace CFEAR_Radarodometry;
int main(int argc, char** argv) {
  if (argc < 4) return 2;
  EvalTrajectory ev;
  Matrix6d c = Matrix6d::Identity(); c(0, 0) = 0.0123456789; c(0, 5) = -1.5e-7; c(5, 5) = 1e-4;
  ev.CallbackESTEigen(vectorToAffine3d(1.23456789, -2.5, 0.3), c, 1547120000u, 625u);
  ev.CallbackESTEigen(vectorToAffine3d(100.0, 0.000012345, -2.9), Matrix6d::Identity(), 1547120001u, 250000000u);
  EvalTrajectory::Write(argv[1], ev.est_vek);
  EvalTrajectory::WriteTUM(argv[2], ev.est_vek);
  EvalTrajectory::WriteCov(argv[3], ev.est_vek);
  return 0;
}
