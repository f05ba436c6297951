// =============================================================================
// cfear_oracle.cc -- CPU restatement of CFEAR's per-scan hot path.
//
// TEST INFRASTRUCTURE ONLY.  This file is the *checker* for the CUDA path in
// cfear_radarodometry_code_public_b200/csrc.  Only tests/, __graft_entry__.smoke()
// and bench.py's cpu_baseline / --impl reference legs may load it.  The product
// never links or calls it.
//
// PARITY UNPINNED: the reference (dan11003/CFEAR_Radarodometry_code_public)
// ships no tests / golden vectors, and cannot be compiled here (ROS1, PCL/FLANN,
// Ceres, Eigen, OpenCV are absent).  The arithmetic that
// lives in un-vendored, un-pinned third-party libraries is restated from their
// published algorithms:
//   * PCL VoxelGrid / KdTreeFLANN (PCL 1.8-1.10, FLANN 1.9 L2_Simple<float>)
//   * Eigen SelfAdjointEigenSolver<Matrix2d>           (closed form here)
//   * Ceres Solver trust-region LM (>= 1.12 minimizer loop, defaults),
//     HuberLoss/CauchyLoss/SoftLOneLoss/TukeyLoss/ScaledLoss/ComposedLoss,
//     Corrector (rho'' <= 0 branch) and Covariance.
// Every function cites the reference file:line it follows
// (paths relative to /root/reference).
//
// Build: see oracle/Makefile  (g++ -O3 -ffp-contract=off, no FMA contraction so
// the fp32 coordinates match a plain x86-64 -O3 build of the reference).
// =============================================================================
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <limits>
#include <memory>
#include <thread>
#include <vector>

namespace {

// ----------------------------------------------------------------------------
// A.1  k-strongest  (src/cfear_radarodometry/radar_filters.cpp:209-237)
// Per azimuth row keep the k lexicographically largest (intensity, range)
// pairs with intensity >= z_min; stored ascending.
// ----------------------------------------------------------------------------
typedef std::pair<uint8_t, int> intensity_range;

void kstrongest_row(const uint8_t* row, int R, uint8_t zmin, int k,
                    std::vector<intensity_range>& out) {
  out.clear();
  for (int range = 0; range < R; ++range) {
    const uint8_t intensity = row[range];
    if (intensity < zmin) continue;                       // :217
    if (out.empty()) {
      out.push_back(std::make_pair(intensity, range));    // :222
    } else {
      const intensity_range p = std::make_pair(intensity, range);
      auto it = std::lower_bound(out.cbegin(), out.cend(), p);   // :225
      out.insert(it, p);
      if ((size_t)out.size() > (size_t)k) out.erase(out.begin());  // :227-228
    }
  }
}

// ----------------------------------------------------------------------------
// utils.h:28-32  GetRelTimeStamp
// ----------------------------------------------------------------------------
inline double rel_time_stamp(double x, double y, bool ccw) {
  double a = atan2(y, x);
  double d = ((a > 0.00001 ? a : (2 * M_PI + a)) / (2 * M_PI));
  return ccw ? -(d - 0.5) : (d - 0.5);
}

// ----------------------------------------------------------------------------
// A.3  2x2 symmetric eigen decomposition (closed form; restates what
// Eigen::SelfAdjointEigenSolver<Matrix2d> returns at pointnormal.cpp:39-45:
// eigenvalues ascending, unit eigenvectors; signs are unspecified upstream,
// the normal's sign is fixed afterwards by the flip toward the origin).
// ----------------------------------------------------------------------------
struct Eig2 { double lmin, lmax, nx, ny, ox, oy; };

Eig2 eig2_sym(double a, double b, double d) {
  Eig2 e;
  const double t = 0.5 * (a - d);
  const double m = 0.5 * (a + d);
  const double h = std::sqrt(t * t + b * b);
  e.lmax = m + h;
  e.lmin = m - h;
  double vx, vy;  // eigenvector of lmax
  if (h == 0.0) { vx = 1.0; vy = 0.0; }
  else if (t >= 0.0) { vx = t + h; vy = b; }
  else { vx = b; vy = h - t; }
  const double nrm = std::sqrt(vx * vx + vy * vy);
  if (nrm > 0.0) { vx /= nrm; vy /= nrm; } else { vx = 1.0; vy = 0.0; }
  e.ox = vx; e.oy = vy;        // orth_normal  (col(1), lambda_max)
  e.nx = -vy; e.ny = vx;       // snormal_     (col(0), lambda_min)
  return e;
}

struct Cell {
  double ux, uy;
  double cxx, cxy, cyx, cyy;
  double scale;            // planarity  log(1+cond/2)
  double nx, ny, ox, oy;
  double lmin, lmax;
  double sum_intensity, avg_intensity;
  int nsamples;
  bool valid;
};

// pointnormal.cpp:7-63  cell::cell + cell::ComputeNormal
Cell make_cell(const float* xyzi, const std::vector<int>& idx, bool weight_intensity,
               double origin_x, double origin_y) {
  Cell c;
  const size_t N = idx.size();
  c.nsamples = (int)N;
  std::vector<double> w(N), x(2 * N);
  double wsum = 0.0;
  for (size_t i = 0; i < N; ++i) {
    const float* p = xyzi + 4 * (size_t)idx[i];
    x[2 * i] = (double)p[0];
    x[2 * i + 1] = (double)p[1];
    w[i] = weight_intensity ? std::max((double)p[3] - 60.0, 0.0) : 1.0;   // :15
  }
  for (size_t i = 0; i < N; ++i) wsum += w[i];                             // :18
  c.sum_intensity = wsum;
  c.avg_intensity = wsum / (double)N;                                      // :19
  for (size_t i = 0; i < N; ++i) w[i] = w[i] / wsum;                       // :21
  double ux = 0.0, uy = 0.0;
  for (size_t i = 0; i < N; ++i) { ux += w[i] * x[2 * i]; uy += w[i] * x[2 * i + 1]; }  // :23-24
  for (size_t i = 0; i < N; ++i) { x[2 * i] -= ux; x[2 * i + 1] -= uy; }  // :26-27
  double cxx = 0, cxy = 0, cyx = 0, cyy = 0;                               // cov = x^T * (w .* x)  :29-33
  for (size_t i = 0; i < N; ++i) {
    const double wx = w[i] * x[2 * i], wy = w[i] * x[2 * i + 1];
    cxx += x[2 * i] * wx;     cxy += x[2 * i] * wy;
    cyx += x[2 * i + 1] * wx; cyy += x[2 * i + 1] * wy;
  }
  c.ux = ux; c.uy = uy; c.cxx = cxx; c.cxy = cxy; c.cyx = cyx; c.cyy = cyy;
  // ComputeNormal :37-63
  Eig2 e = eig2_sym(cxx, cyx, cyy);   // SelfAdjointEigenSolver reads the lower triangle
  c.lmin = e.lmin; c.lmax = e.lmax;
  c.nx = e.nx; c.ny = e.ny; c.ox = e.ox; c.oy = e.oy;
  const double cond = std::fabs(c.lmax / c.lmin);
  const double det = c.lmax * c.lmin;
  const bool reasonable = (cond <= 10000) && (det > 0.00001) && c.lmin > 0 && c.lmax > 0;   // :56
  c.scale = std::log(1.0 + cond / 2);                                      // :57
  const double px = origin_x - ux, py = origin_y - uy;
  if (c.nx * px + c.ny * py < 0) { c.nx = -c.nx; c.ny = -c.ny; }          // :59-61
  c.valid = reasonable;
  return c;
}

// ----------------------------------------------------------------------------
// A.2  PCL VoxelGrid centroids (pointnormal.cpp:277-280) -- restated from
// pcl/filters/impl/voxel_grid.hpp (applyFilter): fp32 keys & centroids,
// ascending voxel index output.  Within a voxel points are accumulated in
// ascending input order (upstream order is unspecified: unstable std::sort).
// ----------------------------------------------------------------------------
struct VoxelOut { std::vector<float> cx, cy, cz, ci; std::vector<int> vid; int divx, divy, minbx, minby; };

void voxel_grid(const float* xyzi, int n, float leaf, VoxelOut& out) {
  const float inv = 1.0f / leaf;
  float minx = FLT_MAX, miny = FLT_MAX, minz = FLT_MAX, maxx = -FLT_MAX, maxy = -FLT_MAX, maxz = -FLT_MAX;
  for (int i = 0; i < n; ++i) {
    const float* p = xyzi + 4 * (size_t)i;
    minx = std::min(minx, p[0]); maxx = std::max(maxx, p[0]);
    miny = std::min(miny, p[1]); maxy = std::max(maxy, p[1]);
    minz = std::min(minz, p[2]); maxz = std::max(maxz, p[2]);
  }
  const int minbx = (int)std::floor(minx * inv), maxbx = (int)std::floor(maxx * inv);
  const int minby = (int)std::floor(miny * inv), maxby = (int)std::floor(maxy * inv);
  const int minbz = (int)std::floor(minz * inv), maxbz = (int)std::floor(maxz * inv);
  const int divx = maxbx - minbx + 1, divy = maxby - minby + 1;
  (void)maxbz;
  std::vector<std::pair<int, int>> keyed(n);
  for (int i = 0; i < n; ++i) {
    const float* p = xyzi + 4 * (size_t)i;
    const int ijk0 = (int)(std::floor(p[0] * inv) - (float)minbx);
    const int ijk1 = (int)(std::floor(p[1] * inv) - (float)minby);
    const int ijk2 = (int)(std::floor(p[2] * inv) - (float)minbz);
    keyed[i] = std::make_pair(ijk0 + ijk1 * divx + ijk2 * divx * divy, i);
  }
  std::sort(keyed.begin(), keyed.end());   // (voxel, input index) -> deterministic
  out = VoxelOut();
  out.divx = divx; out.divy = divy; out.minbx = minbx; out.minby = minby;
  size_t i = 0;
  while (i < keyed.size()) {
    size_t j = i;
    float sx = 0.f, sy = 0.f, sz = 0.f, si = 0.f;
    while (j < keyed.size() && keyed[j].first == keyed[i].first) {
      const float* p = xyzi + 4 * (size_t)keyed[j].second;
      sx += p[0]; sy += p[1]; sz += p[2]; si += p[3];
      ++j;
    }
    const float cnt = (float)(j - i);
    out.cx.push_back(sx / cnt); out.cy.push_back(sy / cnt);
    out.cz.push_back(sz / cnt); out.ci.push_back(si / cnt);
    out.vid.push_back(keyed[i].first);
    i = j;
  }
}

// Uniform bucket grid over fp32 points: exact radius / nearest queries with the
// fp32 L2_Simple distance ((dx*dx + dy*dy) [+ dz*dz]) used by FLANN.
struct BucketGrid {
  float g = 1.f, ox = 0.f, oy = 0.f;
  int nx = 0, ny = 0;
  std::vector<int> start, items;
  std::vector<float> px, py;
  void build(const float* x, const float* y, int n, float cell) {
    g = cell; px.assign(x, x + n); py.assign(y, y + n);
    float minx = FLT_MAX, miny = FLT_MAX, maxx = -FLT_MAX, maxy = -FLT_MAX;
    for (int i = 0; i < n; ++i) { minx = std::min(minx, x[i]); maxx = std::max(maxx, x[i]); miny = std::min(miny, y[i]); maxy = std::max(maxy, y[i]); }
    if (n == 0) { minx = miny = maxx = maxy = 0.f; }
    ox = minx; oy = miny;
    nx = (int)std::floor((maxx - ox) / g) + 1; ny = (int)std::floor((maxy - oy) / g) + 1;
    start.assign((size_t)nx * ny + 1, 0); items.resize(n);
    std::vector<int> cid(n);
    for (int i = 0; i < n; ++i) { cid[i] = cx(x[i]) + cy(y[i]) * nx; start[cid[i] + 1]++; }
    for (size_t c = 0; c < (size_t)nx * ny; ++c) start[c + 1] += start[c];
    std::vector<int> fill(start.begin(), start.end() - 1);
    for (int i = 0; i < n; ++i) items[fill[cid[i]]++] = i;
  }
  int cx(float x) const { int c = (int)std::floor((x - ox) / g); return std::min(std::max(c, 0), nx - 1); }
  int cy(float y) const { int c = (int)std::floor((y - oy) / g); return std::min(std::max(c, 0), ny - 1); }
  int cxu(float x) const { return (int)std::floor((x - ox) / g); }
  int cyu(float y) const { return (int)std::floor((y - oy) / g); }
};

// A.2 step 3: FLANN radius search, strict d2 < r2, results sorted by distance
void radius_search(const BucketGrid& G, float qx, float qy, float r, std::vector<int>& out) {
  const float r2 = (float)((double)r * (double)r);
  const int span = (int)std::ceil(r / G.g) + 1;
  const int cx0 = G.cxu(qx), cy0 = G.cyu(qy);
  std::vector<std::pair<float, int>> found;
  for (int yy = std::max(cy0 - span, 0); yy <= std::min(cy0 + span, G.ny - 1); ++yy)
    for (int xx = std::max(cx0 - span, 0); xx <= std::min(cx0 + span, G.nx - 1); ++xx) {
      const int c = xx + yy * G.nx;
      for (int s = G.start[c]; s < G.start[c + 1]; ++s) {
        const int i = G.items[s];
        const float dx = qx - G.px[i], dy = qy - G.py[i];
        float d2 = dx * dx; d2 += dy * dy; d2 += 0.f * 0.f;     // z == 0
        if (d2 < r2) found.push_back(std::make_pair(d2, i));
      }
    }
  std::sort(found.begin(), found.end());
  out.resize(found.size());
  for (size_t i = 0; i < found.size(); ++i) out[i] = found[i].second;
}

// pointnormal.cpp:238-254  GetClosestIdx: exact fp32 1-NN, accepted iff d2 < d*d.
// Ties (cells whose fp32 means coincide -- CFEAR's cell sets hold such duplicates): smallest index.  Which copy the
// reference returns is decided by FLANN's tree walk; g_nn_tie_largest flips the choice to the largest index so that
// tests can bound what that freedom does to a registration (tests/test_flann_pin.py).
static int g_nn_tie_largest = 0;
int nearest_within(const BucketGrid& G, double pxd, double pyd, double radius) {
  const float qx = (float)pxd, qy = (float)pyd;     // :241-242
  const int span = (int)std::ceil(radius / G.g) + 1;
  const int cx0 = G.cxu(qx), cy0 = G.cyu(qy);
  float best = FLT_MAX; int besti = -1;
  for (int yy = std::max(cy0 - span, 0); yy <= std::min(cy0 + span, G.ny - 1); ++yy)
    for (int xx = std::max(cx0 - span, 0); xx <= std::min(cx0 + span, G.nx - 1); ++xx) {
      const int c = xx + yy * G.nx;
      for (int s = G.start[c]; s < G.start[c + 1]; ++s) {
        const int i = G.items[s];
        const float dx = qx - G.px[i], dy = qy - G.py[i];
        float d2 = dx * dx; d2 += dy * dy;
        if (d2 < best || (d2 == best && (g_nn_tie_largest ? i > besti : i < besti))) { best = d2; besti = i; }
      }
    }
  if (besti >= 0 && (double)best < radius * radius) return besti;   // :250
  return -1;
}

// ----------------------------------------------------------------------------
// Cell set handed to registration (what MapPointNormal exposes: GetMean2d,
// GetNormal2d, GetCov2d, GetCell(i).Nsamples_, GetPlanarity, kd_cells)
// ----------------------------------------------------------------------------
struct CellSet {
  int n = 0;
  const double* mean = nullptr;     // 2n
  const double* normal = nullptr;   // 2n
  const double* cov = nullptr;      // 4n row-major
  const double* planarity = nullptr;
  const int32_t* nsamples = nullptr;
  BucketGrid grid;                  // fp32 means  (pointnormal.cpp:151-162)
  void build_index() {
    std::vector<float> fx(n), fy(n);
    for (int i = 0; i < n; ++i) { fx[i] = (float)mean[2 * i]; fy[i] = (float)mean[2 * i + 1]; }
    grid.build(fx.data(), fy.data(), n, 2.0f);
  }
};

struct RegCfg {
  int cost;            // 0 P2P, 1 P2L, 2 P2D   (registration.h:55)
  int loss;            // 0 None,1 Huber,2 Cauchy,3 SoftLOne,4 Combined,5 Tukey (registration.h:60)
  double loss_limit;
  int weight_opt;      // registration.h:50
  double cov_scale, regularization;   // n_scan_normal.h:72-73, SetD2dPar
  double radius;       // registration.h:122 (2.0)
  int max_outer, min_outer, max_inner;  // 8, 3, 20
  int solver_mode;     // 0 ceres_lm, 1 gn_fixed
  int gn_iters;
  // Register(..., soft_constraints = true), n_scan_normal.cpp:373-377: residual block alpha * L * (guess - x) without a loss
  // (mahalanobisDistanceError, n_scan_normal.h:259-290; applyOnTheLeft(L), i.e. L and not L^T as in P2D), L =
  // Cov6to3(cov).inverse().llt().matrixL() row-major 3x3, alpha = sqrt(#source cells), guess = the pose handed in.
  const double* prior_L = nullptr;
  double prior_guess[3] = {0, 0, 0};
  double prior_alpha = 0.0;
};

struct Residual {      // one residual block  (n_scan_normal.cpp:266-320)
  double px, py;       // src mean (local)
  double qx, qy;       // target mean in world
  double a, b, c;      // P2L: a=nx, b=ny ; P2D: L = [[a,0],[b,c]]
  double w;            // ScaledLoss weight
};

// Ceres loss functions (ceres/loss_function.cc), rho[0..2]
void loss_eval(int loss, double a, double s, double rho[3]) {
  switch (loss) {
    case 1: {  // HuberLoss
      const double b = a * a;
      if (s > b) { const double r = std::sqrt(s); rho[0] = 2.0 * a * r - b; rho[1] = std::max(DBL_MIN, a / r); rho[2] = -rho[1] / (2.0 * s); }
      else { rho[0] = s; rho[1] = 1.0; rho[2] = 0.0; }
      return; }
    case 2: {  // CauchyLoss
      const double b = a * a, c = 1.0 / b;
      const double sum = 1.0 + s * c, inv = 1.0 / sum;
      rho[0] = b * std::log(sum); rho[1] = std::max(DBL_MIN, inv); rho[2] = -c * (inv * inv);
      return; }
    case 3: {  // SoftLOneLoss
      const double b = a * a, c = 1.0 / b;
      const double sum = 1.0 + s * c, tmp = std::sqrt(sum);
      rho[0] = 2.0 * b * (tmp - 1.0); rho[1] = std::max(DBL_MIN, 1.0 / tmp); rho[2] = -(c * rho[1]) / (2.0 * sum);
      return; }
    case 4: {  // ComposedLoss(Huber(1), Cauchy(1))  registration.cpp:88-92
      double g[3], f[3];
      loss_eval(2, 1.0, s, g); loss_eval(1, 1.0, g[0], f);
      rho[0] = f[0]; rho[1] = f[1] * g[1]; rho[2] = f[2] * g[1] * g[1] + f[1] * g[2];
      return; }
    case 5: {  // TukeyLoss (Ceres 2.x form)
      const double a2 = a * a;
      if (s <= a2) { const double v = 1.0 - s / a2, v2 = v * v; rho[0] = a2 / 3.0 * (1.0 - v2 * v); rho[1] = v2; rho[2] = -2.0 / a2 * v; }
      else { rho[0] = a2 / 3.0; rho[1] = 0.0; rho[2] = 0.0; }
      return; }
    default:   // None: ScaledLoss(nullptr, w): rho = s
      rho[0] = s; rho[1] = 1.0; rho[2] = 0.0; return;
  }
}

struct Eval { double cost; double H[6]; double g[3]; };  // H = J~^T J~ (xx,xy,xt,yy,yt,tt), g = J~^T r~

// Evaluate cost (and optionally normal equations) at x.  Residual / Jacobian
// per n_scan_normal.h:180-255,330-361 ; loss via ScaledLoss(w) + Corrector with
// rho'' <= 0  => rows scaled by sqrt(w rho').
void evaluate(const RegCfg& cfg, const std::vector<Residual>& res, const double x[3], bool with_jac, Eval& ev) {
  const double cs = std::cos(x[2]), sn = std::sin(x[2]);
  double cost = 0.0;
  double H[6] = {0, 0, 0, 0, 0, 0}, g[3] = {0, 0, 0};
  for (const Residual& r : res) {
    const double rx = cs * r.px - sn * r.py, ry = sn * r.px + cs * r.py;   // R(psi) p
    const double ex = rx + x[0] - r.qx, ey = ry + x[1] - r.qy;            // transformed src - target
    const double dpx = -ry, dpy = rx;                                     // d(Rp)/dpsi
    double r0, r1 = 0.0, J0[3], J1[3] = {0, 0, 0};
    int nr;
    if (cfg.cost == 1) {            // P2L  r = v.dot(n)
      r0 = ex * r.a + ey * r.b; nr = 1;
      J0[0] = r.a; J0[1] = r.b; J0[2] = dpx * r.a + dpy * r.b;
    } else if (cfg.cost == 2) {     // P2D  r = L * e   (applyOnTheLeft, n_scan_normal.h:241)
      r0 = r.a * ex; r1 = r.b * ex + r.c * ey; nr = 2;
      J0[0] = r.a; J0[1] = 0.0; J0[2] = r.a * dpx;
      J1[0] = r.b; J1[1] = r.c; J1[2] = r.b * dpx + r.c * dpy;
    } else {                        // P2P  r = tar - src
      r0 = -ex; r1 = -ey; nr = 2;
      J0[0] = -1.0; J0[1] = 0.0; J0[2] = -dpx;
      J1[0] = 0.0; J1[1] = -1.0; J1[2] = -dpy;
    }
    const double s = r0 * r0 + r1 * r1;
    double rho[3];
    loss_eval(cfg.loss, cfg.loss_limit, s, rho);
    cost += 0.5 * r.w * rho[0];
    if (with_jac) {
      const double wr = r.w * rho[1];     // (sqrt(w rho'))^2
      H[0] += wr * J0[0] * J0[0]; H[1] += wr * J0[0] * J0[1]; H[2] += wr * J0[0] * J0[2];
      H[3] += wr * J0[1] * J0[1]; H[4] += wr * J0[1] * J0[2]; H[5] += wr * J0[2] * J0[2];
      g[0] += wr * J0[0] * r0; g[1] += wr * J0[1] * r0; g[2] += wr * J0[2] * r0;
      if (nr == 2) {
        H[0] += wr * J1[0] * J1[0]; H[1] += wr * J1[0] * J1[1]; H[2] += wr * J1[0] * J1[2];
        H[3] += wr * J1[1] * J1[1]; H[4] += wr * J1[1] * J1[2]; H[5] += wr * J1[2] * J1[2];
        g[0] += wr * J1[0] * r1; g[1] += wr * J1[1] * r1; g[2] += wr * J1[2] * r1;
      }
    }
  }
  if (cfg.prior_L) {                                  // the soft prior's 3 rows: r = alpha L (guess - x), J = -alpha L
    const double* L = cfg.prior_L; const double al = cfg.prior_alpha;
    const double d[3] = {cfg.prior_guess[0] - x[0], cfg.prior_guess[1] - x[1], cfg.prior_guess[2] - x[2]};
    for (int row = 0; row < 3; ++row) {
      const double J[3] = {-al * L[3 * row + 0], -al * L[3 * row + 1], -al * L[3 * row + 2]};
      const double r = -(J[0] * d[0] + J[1] * d[1] + J[2] * d[2]);
      cost += 0.5 * r * r;
      if (with_jac) {
        H[0] += J[0] * J[0]; H[1] += J[0] * J[1]; H[2] += J[0] * J[2]; H[3] += J[1] * J[1]; H[4] += J[1] * J[2]; H[5] += J[2] * J[2];
        g[0] += J[0] * r; g[1] += J[1] * r; g[2] += J[2] * r;
      }
    }
  }
  ev.cost = cost;
  if (with_jac) { for (int i = 0; i < 6; ++i) ev.H[i] = H[i]; for (int i = 0; i < 3; ++i) ev.g[i] = g[i]; }
}

// Solve symmetric positive-definite 3x3 (xx,xy,xt,yy,yt,tt) A y = b by Cholesky.
bool chol3_solve(const double A[6], const double b[3], double y[3]) {
  const double l00 = std::sqrt(A[0]);
  if (!(l00 > 0.0) || !std::isfinite(l00)) return false;
  const double l10 = A[1] / l00, l20 = A[2] / l00;
  const double d1 = A[3] - l10 * l10;
  if (!(d1 > 0.0)) return false;
  const double l11 = std::sqrt(d1);
  const double l21 = (A[4] - l20 * l10) / l11;
  const double d2 = A[5] - l20 * l20 - l21 * l21;
  if (!(d2 > 0.0)) return false;
  const double l22 = std::sqrt(d2);
  const double z0 = b[0] / l00;
  const double z1 = (b[1] - l10 * z0) / l11;
  const double z2 = (b[2] - l20 * z0 - l21 * z1) / l22;
  y[2] = z2 / l22;
  y[1] = (z1 - l21 * y[2]) / l11;
  y[0] = (z0 - l10 * y[1] - l20 * y[2]) / l00;
  return std::isfinite(y[0]) && std::isfinite(y[1]) && std::isfinite(y[2]);
}

struct SolveSummary {
  double initial_cost = 0, final_cost = 0;
  int n_iterations = 0;          // summary.iterations.size()
  double last_relative_decrease = 0;
  bool usable = true;
  int n_successful = 0;
};

// ----------------------------------------------------------------------------
// A.6  ceres::Solve with the reference's options (n_scan_normal.cpp:9,443-452;
// registration.cpp:5): TRUST_REGION / LEVENBERG_MARQUARDT, defaults otherwise.
// Restates ceres/trust_region_minimizer.cc + levenberg_marquardt_strategy.cc.
// ----------------------------------------------------------------------------
void ceres_lm_solve(const RegCfg& cfg, const std::vector<Residual>& res, double x[3], SolveSummary& sum) {
  const double kFunctionTol = 1e-6, kGradientTol = 1e-10, kParameterTol = 1e-8;
  const double kMinRelDecrease = 1e-3, kMinDiag = 1e-6, kMaxDiag = 1e32;
  const double kMaxRadius = 1e16, kMinRadius = 1e-32;
  double radius = 1e4, decrease_factor = 2.0;
  bool reuse_diagonal = false;
  int invalid_in_a_row = 0;
  sum = SolveSummary();

  Eval ev;
  evaluate(cfg, res, x, true, ev);                 // IterationZero
  double x_cost = ev.cost;
  sum.initial_cost = x_cost;
  double scale[3];                                 // jacobi_scaling, once
  scale[0] = 1.0 / (1.0 + std::sqrt(ev.H[0]));
  scale[1] = 1.0 / (1.0 + std::sqrt(ev.H[3]));
  scale[2] = 1.0 / (1.0 + std::sqrt(ev.H[5]));
  double x_norm = std::sqrt(x[0] * x[0] + x[1] * x[1] + x[2] * x[2]);
  double min_cost = x_cost;                        // SetSummaryFinalCost
  double gmax = std::max(std::fabs(ev.g[0]), std::max(std::fabs(ev.g[1]), std::fabs(ev.g[2])));
  sum.n_iterations = 1; sum.last_relative_decrease = 0.0;
  sum.final_cost = min_cost;
  if (gmax <= kGradientTol) return;                // CONVERGENCE at iteration 0
  double diag[3] = {0, 0, 0};
  for (int it = 1;; ++it) {
    // scaled normal equations  Hs = S H S, gs = S g
    double Hs[6] = {ev.H[0] * scale[0] * scale[0], ev.H[1] * scale[0] * scale[1], ev.H[2] * scale[0] * scale[2],
                    ev.H[3] * scale[1] * scale[1], ev.H[4] * scale[1] * scale[2], ev.H[5] * scale[2] * scale[2]};
    double gs[3] = {ev.g[0] * scale[0], ev.g[1] * scale[1], ev.g[2] * scale[2]};
    if (!reuse_diagonal) {
      diag[0] = std::min(std::max(Hs[0], kMinDiag), kMaxDiag);
      diag[1] = std::min(std::max(Hs[3], kMinDiag), kMaxDiag);
      diag[2] = std::min(std::max(Hs[5], kMinDiag), kMaxDiag);
    }
    double A[6] = {Hs[0] + diag[0] / radius, Hs[1], Hs[2], Hs[3] + diag[1] / radius, Hs[4], Hs[5] + diag[2] / radius};
    double y[3];
    double nb[3] = {-gs[0], -gs[1], -gs[2]};
    bool ok = chol3_solve(A, nb, y);
    reuse_diagonal = true;
    double model_change = 0.0;
    if (ok) {
      // -(Js y).(r + Js y / 2) = -y.gs - 0.5 y^T Hs y
      const double Hy0 = Hs[0] * y[0] + Hs[1] * y[1] + Hs[2] * y[2];
      const double Hy1 = Hs[1] * y[0] + Hs[3] * y[1] + Hs[4] * y[2];
      const double Hy2 = Hs[2] * y[0] + Hs[4] * y[1] + Hs[5] * y[2];
      model_change = -(y[0] * gs[0] + y[1] * gs[1] + y[2] * gs[2]) - 0.5 * (y[0] * Hy0 + y[1] * Hy1 + y[2] * Hy2);
    }
    if (!ok || !(model_change > 0.0)) {            // invalid step
      if (++invalid_in_a_row >= 5) { sum.usable = false; return; }   // FAILURE
      radius /= decrease_factor; decrease_factor *= 2.0; reuse_diagonal = true;   // StepIsInvalid -> StepRejected(0)
      sum.n_iterations++; sum.last_relative_decrease = 0.0;
      min_cost = std::min(min_cost, x_cost); sum.final_cost = min_cost;
      if (it >= cfg.max_inner) return;
      if (radius <= kMinRadius) return;
      continue;
    }
    invalid_in_a_row = 0;
    const double delta[3] = {y[0] * scale[0], y[1] * scale[1], y[2] * scale[2]};
    const double xc[3] = {x[0] + delta[0], x[1] + delta[1], x[2] + delta[2]};
    Eval evc;
    evaluate(cfg, res, xc, false, evc);
    const double cand_cost = evc.cost;
    const double step_norm = std::sqrt(delta[0] * delta[0] + delta[1] * delta[1] + delta[2] * delta[2]);
    if (step_norm <= kParameterTol * (x_norm + kParameterTol)) return;          // not recorded, x kept
    const double cost_change = x_cost - cand_cost;
    if (std::fabs(cost_change) <= kFunctionTol * x_cost) return;                // not recorded, x kept
    const double rel = cost_change / model_change;
    sum.n_iterations++; sum.last_relative_decrease = rel;
    if (rel > kMinRelDecrease) {                                                // HandleSuccessfulStep
      x[0] = xc[0]; x[1] = xc[1]; x[2] = xc[2];
      x_norm = std::sqrt(x[0] * x[0] + x[1] * x[1] + x[2] * x[2]);
      evaluate(cfg, res, x, true, ev);
      x_cost = ev.cost;
      gmax = std::max(std::fabs(ev.g[0]), std::max(std::fabs(ev.g[1]), std::fabs(ev.g[2])));
      radius = radius / std::max(1.0 / 3.0, 1.0 - std::pow(2.0 * rel - 1.0, 3));
      radius = std::min(kMaxRadius, radius);
      decrease_factor = 2.0; reuse_diagonal = false;
      sum.n_successful++;
      min_cost = std::min(min_cost, x_cost); sum.final_cost = min_cost;
      if (it >= cfg.max_inner) return;
      if (gmax <= kGradientTol) return;
    } else {                                                                    // HandleUnsuccessfulStep
      radius /= decrease_factor; decrease_factor *= 2.0; reuse_diagonal = true;
      min_cost = std::min(min_cost, cand_cost); sum.final_cost = min_cost;     // iteration cost = candidate cost
      if (it >= cfg.max_inner) return;
      if (radius <= kMinRadius) return;
    }
  }
}

// registration.cpp:67-76
double assoc_weight(int opt, double n1, double n2, double simdir, double p1, double p2) {
  auto sim = [](double x, double y) { return 2 * std::min(x, y) / (x + y); };
  switch (opt) {
    case 0: return 1.0;
    case 1: return sim(n1, n2);
    case 2: return simdir;
    case 3: return sim(p1, p2);
    case 4: return sim(n1, n2) + simdir + sim(p1, p2);
    default: return 1.0;
  }
}

// A.4  AddScanPairCost for every keyframe  (n_scan_normal.cpp:215-326, 359-367)
void build_problem(const RegCfg& cfg, const std::vector<CellSet*>& scans, const std::vector<double>& poses,
                   int itr, std::vector<Residual>& res, std::vector<int32_t>* assoc /* per (kf, src) tar idx */,
                   std::vector<double>* assoc_sim = nullptr /* per (kf, src) direction similarity, Weights::sim_dir_ */) {
  res.clear();
  const int ns = (int)scans.size();
  const CellSet& src = *scans[ns - 1];
  const double* xs = &poses[3 * (ns - 1)];
  const double cs_s = std::cos(xs[2]), sn_s = std::sin(xs[2]);
  const double angle_outlier = std::cos(M_PI / 6.0);
  const double curr_radius = (itr == 1) ? 2 * cfg.radius : cfg.radius;     // :222
  if (assoc) assoc->assign((size_t)(ns - 1) * src.n, -1);
  if (assoc_sim) assoc_sim->assign((size_t)(ns - 1) * src.n, 0.0);
  for (int i = 0; i < ns - 1; ++i) {
    const CellSet& tar = *scans[i];
    const double* xt = &poses[3 * i];
    const double ct = std::cos(xt[2]), st = std::sin(xt[2]);
    // Tsrctotar = Ttar^-1 * Tsrc   :224
    const double rc = ct * cs_s + st * sn_s;     // cos(psi_s - psi_t)
    const double rs = ct * sn_s - st * cs_s;     // sin(psi_s - psi_t)
    const double dx = xs[0] - xt[0], dy = xs[1] - xt[1];
    const double tx = ct * dx + st * dy, ty = -st * dx + ct * dy;
    for (int j = 0; j < src.n; ++j) {
      const double mx = src.mean[2 * j], my = src.mean[2 * j + 1];
      const double qx = rc * mx - rs * my + tx, qy = rs * mx + rc * my + ty;      // :240
      const int m = nearest_within(tar.grid, qx, qy, curr_radius);               // :241
      if (m < 0) continue;
      const double nsx = src.normal[2 * j], nsy = src.normal[2 * j + 1];
      const double ntx_ = rc * nsx - rs * nsy, nty_ = rs * nsx + rc * nsy;        // :244
      const double sim = std::max(ntx_ * tar.normal[2 * m] + nty_ * tar.normal[2 * m + 1], 0.0);   // :246
      if (!(sim > angle_outlier)) continue;                                      // :247
      Residual r;
      r.w = assoc_weight(cfg.weight_opt, (double)src.nsamples[j], (double)tar.nsamples[m], sim,
                         src.planarity[j], tar.planarity[m]);                    // :249-255, 274
      r.px = mx; r.py = my;
      const double tmx = tar.mean[2 * m], tmy = tar.mean[2 * m + 1];
      r.qx = ct * tmx - st * tmy + xt[0]; r.qy = st * tmx + ct * tmy + xt[1];    // Ttar*tar_mean
      r.a = r.b = r.c = 0.0;
      if (cfg.cost == 1) {            // :279-289
        const double nx = tar.normal[2 * m], ny = tar.normal[2 * m + 1];
        r.a = ct * nx - st * ny; r.b = st * nx + ct * ny;
      } else if (cfg.cost == 2) {     // :290-300
        const double* C = tar.cov + 4 * (size_t)m;
        // R C R^T
        const double a00 = ct * C[0] - st * C[2], a01 = ct * C[1] - st * C[3];
        const double a10 = st * C[0] + ct * C[2], a11 = st * C[1] + ct * C[3];
        double s00 = a00 * ct - a01 * st, s01 = a00 * st + a01 * ct;
        double s10 = a10 * ct - a11 * st, s11 = a10 * st + a11 * ct;
        s00 = (cfg.regularization + s00) * cfg.cov_scale; s11 = (cfg.regularization + s11) * cfg.cov_scale;
        s01 = s01 * cfg.cov_scale; s10 = s10 * cfg.cov_scale;
        const double det = s00 * s11 - s01 * s10;
        const double i00 = s11 / det, i01 = -s01 / det, i10 = -s10 / det, i11 = s00 / det;   // tar_cov.inverse()
        (void)i01;
        // llt().matrixL(): uses the lower triangle
        const double l00 = std::sqrt(i00);
        const double l10 = i10 / l00;
        const double l11 = std::sqrt(i11 - l10 * l10);
        r.a = l00; r.b = l10; r.c = l11;
      }
      res.push_back(r);
      if (assoc) (*assoc)[(size_t)i * src.n + j] = m;
      if (assoc_sim) (*assoc_sim)[(size_t)i * src.n + j] = sim;
    }
  }
}

int num_scalar_residuals(const RegCfg& cfg, size_t nblocks) { return (int)nblocks * (cfg.cost == 1 ? 1 : 2); }

struct RegStats {
  int32_t success;          // Register() return value
  int32_t outer_iterations; // itr_ after the loop (n_scan_normal.cpp:161)
  int32_t inner_iterations; // sum of summary.iterations.size()-1
  int32_t num_residuals;    // scalar residuals of the last problem
  int32_t num_blocks;       // residual blocks of the last problem
  int32_t usable;
  double final_cost;
  double score;
  int32_t pose_written;     // some solve was usable: Tsrc rewritten from the parameters (n_scan_normal.cpp:119-121)
  int32_t reserved;
};

// A.5  n_scan_normal_reg::Register   (n_scan_normal.cpp:82-187)
bool do_register(const RegCfg& cfg_in, std::vector<CellSet*>& scans, std::vector<double>& poses, double cov36[36],
                 RegStats& st, std::vector<int32_t>* last_assoc, std::vector<double>* last_sim = nullptr) {
  const int ns = (int)scans.size();
  st = RegStats();
  double* x = &poses[3 * (ns - 1)];
  RegCfg cfg = cfg_in;
  if (cfg.prior_L) {                                                       // :95-96 guess, :375 alpha
    cfg.prior_guess[0] = x[0]; cfg.prior_guess[1] = x[1]; cfg.prior_guess[2] = x[2];
    cfg.prior_alpha = std::sqrt((double)scans[ns - 1]->n);
    if (cfg.solver_mode != 0) cfg.prior_L = nullptr;                       // Register()'s ceres_lm loop only
  }
  std::vector<Residual> res;
  SolveSummary sum;
  bool success = true;
  int inner_total = 0;
  if (cfg.solver_mode == 1) {
    // gn_fixed: N undamped Gauss-Newton/IRLS iterations, re-associating each one
    // (BASELINE config 2).  Not a reference mode; defined by this repo.
    int it;
    for (it = 1; it <= cfg.gn_iters; ++it) {
      build_problem(cfg, scans, poses, it, res, last_assoc, last_sim);
      if (num_scalar_residuals(cfg, res.size()) <= 1) { success = false; break; }
      Eval ev; evaluate(cfg, res, x, true, ev);
      double y[3], nb[3] = {-ev.g[0], -ev.g[1], -ev.g[2]};
      if (!chol3_solve(ev.H, nb, y)) { success = false; break; }
      x[0] += y[0]; x[1] += y[1]; x[2] += y[2];
      st.pose_written = 1;
      sum.final_cost = ev.cost;
      inner_total++;
    }
    st.outer_iterations = it;
    if (success) { Eval ev; evaluate(cfg, res, x, false, ev); sum.final_cost = ev.cost; }
  } else {
    double prev_par[3] = {x[0], x[1], x[2]};
    double prev_score = DBL_MAX;
    int itr;
    for (itr = 1; itr <= cfg.max_outer && success; ++itr) {               // :102
      build_problem(cfg, scans, poses, itr, res, last_assoc, last_sim);  // :105
      if (num_scalar_residuals(cfg, res.size()) <= 1) { success = false; break; }   // :370, :114
      const double x_in[3] = {x[0], x[1], x[2]};
      ceres_lm_solve(cfg, res, x, sum);                                  // :117
      success = sum.usable;                                              // :451
      // ceres writes the state back to the parameter blocks only if the solution is usable (solver.cc Minimize)
      if (!success) { x[0] = x_in[0]; x[1] = x_in[1]; x[2] = x_in[2]; }
      else st.pose_written = 1;                                          // :119-121
      inner_total += sum.n_iterations - 1;
      const double current_score = sum.final_cost;                       // :123
      const double rel_improvement = (prev_score - current_score) / prev_score;
      if (itr > cfg.min_outer) {                                         // :134
        if (prev_score < current_score) { x[0] = prev_par[0]; x[1] = prev_par[1]; x[2] = prev_par[2]; break; }
        else if (rel_improvement < 0.00001) break;
        else if (sum.last_relative_decrease < 0.00001 || sum.n_iterations == 1) break;
      }
      prev_score = current_score;
      prev_par[0] = x[0]; prev_par[1] = x[1]; prev_par[2] = x[2];
    }
    st.outer_iterations = itr;
  }
  st.inner_iterations = inner_total;
  st.num_blocks = (int)res.size();
  st.num_residuals = num_scalar_residuals(cfg, res.size());
  if (cfg.prior_L && st.num_residuals > 1) { st.num_blocks += 1; st.num_residuals += 3; }   // summary_.num_residuals counts the prior
  st.usable = sum.usable ? 1 : 0;
  st.final_cost = sum.final_cost;
  for (int i = 0; i < 36; ++i) cov36[i] = 0.0;
  if (!success) { st.success = 0; return false; }
  st.score = sum.final_cost / st.num_residuals;                           // :166
  // reg_cov defaults :171-175 (the caller applies them to the other scans)
  cov36[0] = 0.01; cov36[7] = 0.01; cov36[35] = 0.0001;
  // GetCovariance :392-433  (ceres::Covariance at the current parameters, loss applied)
  Eval ev; evaluate(cfg, res, x, true, ev);
  // inverse of H by Cholesky; rank deficiency -> false
  double inv[9];
  bool ok = true;
  for (int c = 0; c < 3 && ok; ++c) {
    double e[3] = {0, 0, 0}, y[3]; e[c] = 1.0;
    ok = chol3_solve(ev.H, e, y);
    inv[0 + c] = y[0]; inv[3 + c] = y[1]; inv[6 + c] = y[2];
  }
  if (!ok) { st.success = 0; return false; }                              // :411-412
  if (st.num_residuals - 3 == 0) { st.success = 0; return false; }       // :417
  const double f = 30 * (sum.final_cost / (st.num_residuals - 3));       // :418
  for (int i = 0; i < 36; ++i) cov36[i] = 0.0;
  for (int i = 0; i < 6; ++i) cov36[i * 6 + i] = 1.0;                     // :426
  cov36[0] = f * inv[0]; cov36[1] = f * inv[1]; cov36[6] = f * inv[3]; cov36[7] = f * inv[4];   // :427
  cov36[35] = f * inv[8];                                                 // :428
  cov36[5] = f * inv[2];                                                  // :429
  cov36[30] = f * inv[6];                                                 // :430
  st.success = 1;
  return true;
}

}  // namespace

// =============================================================================
// C interface (ctypes / tests / bench cpu_baseline)
// =============================================================================
extern "C" {

// radar_filters.cpp:198-237.  idx_out: A*k range bins, ascending (intensity,range), -1 padded.
int orc_kstrongest(const uint8_t* img, int A, int R, int z_min, int k, int32_t* idx_out, int32_t* cnt_out) {
  if (k < 1) return -1;
  const uint8_t zmin = (uint8_t)z_min;                                     // :212
  std::vector<intensity_range> row;
  for (int a = 0; a < A; ++a) {
    kstrongest_row(img + (size_t)a * R, R, zmin, k, row);
    cnt_out[a] = (int)row.size();
    for (int j = 0; j < k; ++j) idx_out[(size_t)a * k + j] = j < (int)row.size() ? row[j].second : -1;
  }
  return 0;
}

// radar_filters.cpp:238-298  AxialNonMaxSupress.  Out-of-row reads (the reference
// indexes the flat cv::Mat buffer unchecked for 3<=r<6 and R-6<=r<R-3) are
// restated as reads of the flat image buffer, clamped to [0, A*R).
int orc_peaks(const uint8_t* img, int A, int R, int k, const int32_t* idx, const int32_t* cnt,
              int32_t* pidx_out, int32_t* pcnt_out) {
  const int W = 3;
  const long total = (long)A * R;
  for (int a = 0; a < A; ++a) {
    int np = 0;
    auto score = [&](int r_n, bool computed_ok) -> uint16_t {
      if (!computed_ok) return 0;
      uint16_t s = 0;
      for (int r_nn = r_n - W; r_nn <= r_n + W; ++r_nn) {
        long flat = (long)a * R + r_nn;
        flat = std::min(std::max(flat, 0L), total - 1);
        s = (uint16_t)(s + (uint16_t)img[flat]);
      }
      return s;
    };
    // which r_n have a computed score: those within +-W of an in-bounds kept bin (:251-263)
    std::vector<char> have(R + 2 * W + 2, 0);
    for (int j = 0; j < cnt[a]; ++j) {
      const int r = idx[(size_t)a * k + j];
      if (r < W || r >= R - W) continue;
      for (int r_n = r - W; r_n <= r + W; ++r_n) have[r_n + W] = 1;
    }
    auto sc = [&](int r_n) -> uint16_t {
      const int h = r_n + W;
      const bool ok = h >= 0 && h < (int)have.size() && have[h];
      return score(r_n, ok);      // score[] default-inserts 0 for unknown keys (:271-276)
    };
    for (int j = 0; j < cnt[a]; ++j) {
      const int r = idx[(size_t)a * k + j];
      bool largest = true;
      const uint16_t pthis = sc(r);
      for (int i = 1; i <= W; ++i) {
        const uint16_t pnext = sc(r + i), pprev = sc(r - i);
        if (pprev > pthis || pthis < pnext) { largest = false; break; }     // :282
      }
      if (largest) pidx_out[(size_t)a * k + np++] = r;
    }
    pcnt_out[a] = np;
    for (int j = np; j < k; ++j) pidx_out[(size_t)a * k + j] = -1;
  }
  return 0;
}

// radar_filters.cpp:309-337.  xyzi_out: up to A*k points (x,y,z,intensity) fp32.  Returns n.
int orc_cloud(const uint8_t* img, int A, int R, int k, const int32_t* idx, const int32_t* cnt,
              float min_distance_f, float range_res_f, float* xyzi_out) {
  const double min_distance = (double)min_distance_f, range_res = (double)range_res_f;   // radar_driver.cpp:58 (float -> double)
  const int min_range_bin = (int)std::ceil(min_distance / range_res);      // :315
  int n = 0;
  for (int a = 0; a < A; ++a) {
    const double theta = ((double)(a + 1) / A) * 2. * M_PI;                // :317
    if (cnt[a] == 0) continue;
    const double cos_t = std::cos(theta), sin_t = std::sin(theta);
    const double half = range_res / 2.0;
    for (int j = 0; j < cnt[a]; ++j) {
      const int range = idx[(size_t)a * k + j];
      if (range > min_range_bin) {                                         // :327
        float* p = xyzi_out + 4 * (size_t)n++;
        p[0] = (float)((half + range_res * range) * cos_t);
        p[1] = (float)((half + range_res * range) * sin_t);
        p[2] = 0.f;
        p[3] = (float)img[(size_t)a * R + range];
      }
    }
  }
  return n;
}

// utils.cpp:96-113 Compensate
void orc_compensate(float* xyzi, int n, const double* mot, int ccw) {
  for (int i = 0; i < n; ++i) {
    float* p = xyzi + 4 * (size_t)i;
    const double d = rel_time_stamp((double)p[0], (double)p[1], ccw != 0);
    const double s1 = std::sin(d * mot[2]), c1 = std::cos(d * mot[2]);      // utils.cpp:130-139
    const double x = (double)p[0], y = (double)p[1];
    const double tx = c1 * x + (-s1) * y + d * mot[0];
    const double ty = s1 * x + c1 * y + d * mot[1];
    p[0] = (float)tx; p[1] = (float)ty;
  }
}

// Voxel centroids only (for tests).  Returns number of voxels; outputs sized >= n.
int orc_voxel_centroids(const float* xyzi, int n, float radius, double downsample_factor,
                        float* cx, float* cy, float* ci, int32_t* vid, int32_t* dims4) {
  VoxelOut vo;
  const float leaf = (float)((double)radius / downsample_factor);
  voxel_grid(xyzi, n, leaf, vo);
  for (size_t i = 0; i < vo.cx.size(); ++i) { cx[i] = vo.cx[i]; cy[i] = vo.cy[i]; ci[i] = vo.ci[i]; vid[i] = vo.vid[i]; }
  dims4[0] = vo.divx; dims4[1] = vo.divy; dims4[2] = vo.minbx; dims4[3] = vo.minby;
  return (int)vo.cx.size();
}

// pointnormal.cpp:65-90, 265-297.  Outputs sized >= n.  Returns number of valid cells.
int orc_surface_points(const float* xyzi, int n, float radius, double downsample_factor, int weight_intensity,
                       double origin_x, double origin_y,
                       double* mean, double* normal, double* cov, double* planarity,
                       int32_t* nsamples, double* avg_intensity, double* lambdas /*2 per cell or null*/) {
  if (n <= 0) return 0;
  VoxelOut vo;
  const float leaf = (float)((double)radius / downsample_factor);          // :279
  voxel_grid(xyzi, n, leaf, vo);
  std::vector<float> px(n), py(n);
  for (int i = 0; i < n; ++i) { px[i] = xyzi[4 * (size_t)i]; py[i] = xyzi[4 * (size_t)i + 1]; }
  BucketGrid G; G.build(px.data(), py.data(), n, radius);
  std::vector<int> nb;
  int nc = 0;
  for (size_t v = 0; v < vo.cx.size(); ++v) {
    radius_search(G, vo.cx[v], vo.cy[v], radius, nb);
    if ((int)nb.size() >= 6) {                                             // :291
      Cell c = make_cell(xyzi, nb, weight_intensity != 0, origin_x, origin_y);
      if (c.valid) {                                                       // :293
        mean[2 * nc] = c.ux; mean[2 * nc + 1] = c.uy;
        normal[2 * nc] = c.nx; normal[2 * nc + 1] = c.ny;
        cov[4 * nc] = c.cxx; cov[4 * nc + 1] = c.cxy; cov[4 * nc + 2] = c.cyx; cov[4 * nc + 3] = c.cyy;
        planarity[nc] = c.scale; nsamples[nc] = c.nsamples; avg_intensity[nc] = c.avg_intensity;
        if (lambdas) { lambdas[2 * nc] = c.lmin; lambdas[2 * nc + 1] = c.lmax; }
        ++nc;
      }
    }
  }
  return nc;
}

// pointnormal.cpp:238-254 on a set of cell means.
void orc_set_nn_tie_largest(int on) { g_nn_tie_largest = on; }

void orc_nearest(const double* means, int n, const double* queries, int nq, double radius, int32_t* out) {
  CellSet cs; cs.n = n; cs.mean = means; cs.build_index();
  for (int q = 0; q < nq; ++q) out[q] = nearest_within(cs.grid, queries[2 * q], queries[2 * q + 1], radius);
}

// cfg_i: cost, loss, weight_opt, max_outer, min_outer, max_inner, solver_mode, gn_iters
// cfg_d: loss_limit, cov_scale, regularization, radius
// Cell sets are concatenated: offsets[nscans+1] index into mean/normal/cov/...
int orc_register(const int32_t* cfg_i, const double* cfg_d, int nscans, const int32_t* offsets,
                 const double* mean, const double* normal, const double* cov, const double* planarity,
                 const int32_t* nsamples, double* poses /*nscans*3, last in/out*/, double* cov36,
                 void* stats_out /*RegStats*/, int32_t* assoc_out /*(nscans-1)*n_src or null*/) {
  RegCfg cfg;
  cfg.cost = cfg_i[0]; cfg.loss = cfg_i[1]; cfg.weight_opt = cfg_i[2];
  cfg.max_outer = cfg_i[3]; cfg.min_outer = cfg_i[4]; cfg.max_inner = cfg_i[5];
  cfg.solver_mode = cfg_i[6]; cfg.gn_iters = cfg_i[7];
  cfg.loss_limit = cfg_d[0]; cfg.cov_scale = cfg_d[1]; cfg.regularization = cfg_d[2]; cfg.radius = cfg_d[3];
  std::vector<CellSet> sets(nscans);
  std::vector<CellSet*> ptrs(nscans);
  for (int i = 0; i < nscans; ++i) {
    const int o = offsets[i];
    sets[i].n = offsets[i + 1] - o;
    sets[i].mean = mean + 2 * (size_t)o; sets[i].normal = normal + 2 * (size_t)o; sets[i].cov = cov + 4 * (size_t)o;
    sets[i].planarity = planarity + o; sets[i].nsamples = nsamples + o;
    if (i < nscans - 1) sets[i].build_index();
    ptrs[i] = &sets[i];
  }
  std::vector<double> p(poses, poses + 3 * (size_t)nscans);
  RegStats st;
  std::vector<int32_t> assoc;
  bool ok = do_register(cfg, ptrs, p, cov36, st, assoc_out ? &assoc : nullptr);
  for (int i = 0; i < 3 * nscans; ++i) poses[i] = p[i];
  if (stats_out) std::memcpy(stats_out, &st, sizeof(RegStats));
  if (assoc_out) std::copy(assoc.begin(), assoc.end(), assoc_out);
  return ok ? 1 : 0;
}

// orc_register plus the similarity table and the soft prior (prior_L: row-major 3x3 lower-triangular, may be null)
int orc_register_ex(const int32_t* cfg_i, const double* cfg_d, int nscans, const int32_t* offsets, const double* mean,
                    const double* normal, const double* cov, const double* planarity, const int32_t* nsamples,
                    double* poses, double* cov36, void* stats_out, int32_t* assoc_out, double* sim_out, const double* prior_L) {
  RegCfg cfg;
  cfg.cost = cfg_i[0]; cfg.loss = cfg_i[1]; cfg.weight_opt = cfg_i[2];
  cfg.max_outer = cfg_i[3]; cfg.min_outer = cfg_i[4]; cfg.max_inner = cfg_i[5];
  cfg.solver_mode = cfg_i[6]; cfg.gn_iters = cfg_i[7];
  cfg.loss_limit = cfg_d[0]; cfg.cov_scale = cfg_d[1]; cfg.regularization = cfg_d[2]; cfg.radius = cfg_d[3];
  cfg.prior_L = prior_L;
  std::vector<CellSet> sets(nscans);
  std::vector<CellSet*> ptrs(nscans);
  for (int i = 0; i < nscans; ++i) {
    const int o = offsets[i];
    sets[i].n = offsets[i + 1] - o;
    sets[i].mean = mean + 2 * (size_t)o; sets[i].normal = normal + 2 * (size_t)o; sets[i].cov = cov + 4 * (size_t)o;
    sets[i].planarity = planarity + o; sets[i].nsamples = nsamples + o;
    if (i < nscans - 1) sets[i].build_index();
    ptrs[i] = &sets[i];
  }
  std::vector<double> p(poses, poses + 3 * (size_t)nscans);
  RegStats st;
  std::vector<int32_t> assoc; std::vector<double> sim;
  bool ok = do_register(cfg, ptrs, p, cov36, st, &assoc, &sim);
  for (int i = 0; i < 3 * nscans; ++i) poses[i] = p[i];
  if (stats_out) std::memcpy(stats_out, &st, sizeof(RegStats));
  if (assoc_out) std::copy(assoc.begin(), assoc.end(), assoc_out);
  if (sim_out) std::copy(sim.begin(), sim.end(), sim_out);
  return ok ? 1 : 0;
}

int orc_regstats_size() { return (int)sizeof(RegStats); }

// n_scan_normal_reg::GetCost (n_scan_normal.cpp:187-213): BuildOptimizationProblem at the poses given, then
// ceres::Problem::Evaluate with default options = 1/2 sum w rho(s) over the residual blocks.  itr_ is whatever the last
// Register left (>= 2), so the association radius is cfg.radius (:222).  Returns 1 / 0 like GetCost's bool.
int orc_get_cost(const int32_t* cfg_i, const double* cfg_d, int nscans, const int32_t* offsets,
                 const double* mean, const double* normal, const double* cov, const double* planarity,
                 const int32_t* nsamples, const double* poses, double* cost_out, int32_t* num_residuals_out) {
  RegCfg cfg;
  cfg.cost = cfg_i[0]; cfg.loss = cfg_i[1]; cfg.weight_opt = cfg_i[2];
  cfg.max_outer = cfg_i[3]; cfg.min_outer = cfg_i[4]; cfg.max_inner = cfg_i[5];
  cfg.solver_mode = cfg_i[6]; cfg.gn_iters = cfg_i[7];
  cfg.loss_limit = cfg_d[0]; cfg.cov_scale = cfg_d[1]; cfg.regularization = cfg_d[2]; cfg.radius = cfg_d[3];
  std::vector<CellSet> sets(nscans);
  std::vector<CellSet*> ptrs(nscans);
  for (int i = 0; i < nscans; ++i) {
    const int o = offsets[i];
    sets[i].n = offsets[i + 1] - o;
    sets[i].mean = mean + 2 * (size_t)o; sets[i].normal = normal + 2 * (size_t)o; sets[i].cov = cov + 4 * (size_t)o;
    sets[i].planarity = planarity + o; sets[i].nsamples = nsamples + o;
    if (i < nscans - 1) sets[i].build_index();
    ptrs[i] = &sets[i];
  }
  std::vector<double> p(poses, poses + 3 * (size_t)nscans);
  std::vector<Residual> res;
  build_problem(cfg, ptrs, p, 2, res, nullptr);
  const int nr = num_scalar_residuals(cfg, res.size());
  if (num_residuals_out) *num_residuals_out = nr;
  *cost_out = 0.0;
  if (nr <= 1) return 0;                                                         // :205-208
  Eval ev;
  evaluate(cfg, res, &p[3 * (size_t)(nscans - 1)], false, ev);
  *cost_out = ev.cost;
  return 1;
}

// Evaluate cost + normal equations for a fixed association at x (unit tests of the LM pieces).
double orc_eval_cost(const int32_t* cfg_i, const double* cfg_d, int nres, const double* res8 /*px,py,qx,qy,a,b,c,w*/,
                     const double* x, double* H6, double* g3) {
  RegCfg cfg; cfg.cost = cfg_i[0]; cfg.loss = cfg_i[1]; cfg.loss_limit = cfg_d[0];
  std::vector<Residual> res(nres);
  for (int i = 0; i < nres; ++i) { const double* r = res8 + 8 * (size_t)i; res[i] = Residual{r[0], r[1], r[2], r[3], r[4], r[5], r[6], r[7]}; }
  Eval ev; evaluate(cfg, res, x, true, ev);
  for (int i = 0; i < 6; ++i) H6[i] = ev.H[i];
  for (int i = 0; i < 3; ++i) g3[i] = ev.g[i];
  return ev.cost;
}

// ----------------------------------------------------------------------------
// Whole per-scan path for many independent problems (CPU baseline; mirrors
// radar_driver.cpp:48-61 -> odometrykeyframefuser.cpp:146-186 for one scan):
//   polar -> k-strongest -> cloud -> Compensate -> MapPointNormal -> Register
// against K resident keyframe cell sets.  nthreads workers over independent
// scans = the reference's own process-level parallel model
// (launch/oxford/eval/utils/worker:86-87).
// pipe_i: A, R, k, z_min, weight_intensity, compensate, ccw, K(keyframes)
// pipe_f: min_distance, range_res, radius(res)  ; downsample_factor = 1
// Keyframe cell sets: for problem b, keyframe i -> set id kf_ids[b*K+i] into the
// concatenated arrays via kf_offsets.
// ----------------------------------------------------------------------------
int orc_pipeline_batch(int nthreads, int nprob, const int32_t* pipe_i, const float* pipe_f,
                       const int32_t* cfg_i, const double* cfg_d,
                       const uint8_t* polar /*nprob*A*R*/, const double* mot /*nprob*3*/,
                       const int32_t* kf_ids, const int32_t* kf_offsets,
                       const double* kf_mean, const double* kf_normal, const double* kf_cov,
                       const double* kf_planarity, const int32_t* kf_nsamples,
                       double* poses /*nprob*(K+1)*3, last in/out*/, double* cov36 /*nprob*36*/,
                       void* stats /*nprob RegStats*/, int32_t* ncells_out /*nprob*/, int32_t* npts_out,
                       double* stage_ms /*3: filter, build_normals, register (summed over scans)*/) {
  const int A = pipe_i[0], R = pipe_i[1], k = pipe_i[2], zmin = pipe_i[3];
  const int wint = pipe_i[4], comp = pipe_i[5], ccw = pipe_i[6], K = pipe_i[7];
  std::atomic<int> next(0);
  std::atomic<long> t_filter(0), t_normals(0), t_reg(0);
  auto worker = [&]() {
    std::vector<int32_t> idx((size_t)A * k), cnt(A);
    std::vector<float> cloud((size_t)A * k * 4);
    std::vector<double> mean, normal, cov, plan, avg;
    std::vector<int32_t> ns;
    for (;;) {
      const int b = next.fetch_add(1);
      if (b >= nprob) break;
      auto t0 = std::chrono::steady_clock::now();
      const uint8_t* img = polar + (size_t)b * A * R;
      orc_kstrongest(img, A, R, zmin, k, idx.data(), cnt.data());
      int n = orc_cloud(img, A, R, k, idx.data(), cnt.data(), pipe_f[0], pipe_f[1], cloud.data());
      auto t1 = std::chrono::steady_clock::now();
      if (comp) orc_compensate(cloud.data(), n, mot + 3 * (size_t)b, ccw);
      mean.resize(2 * (size_t)n + 2); normal.resize(2 * (size_t)n + 2); cov.resize(4 * (size_t)n + 4);
      plan.resize(n + 1); avg.resize(n + 1); ns.resize(n + 1);
      int nc = orc_surface_points(cloud.data(), n, pipe_f[2], 1.0, wint, 0.0, 0.0, mean.data(), normal.data(),
                                  cov.data(), plan.data(), ns.data(), avg.data(), nullptr);
      auto t2 = std::chrono::steady_clock::now();
      if (npts_out) npts_out[b] = n;
      if (ncells_out) ncells_out[b] = nc;
      RegCfg cfg;
      cfg.cost = cfg_i[0]; cfg.loss = cfg_i[1]; cfg.weight_opt = cfg_i[2];
      cfg.max_outer = cfg_i[3]; cfg.min_outer = cfg_i[4]; cfg.max_inner = cfg_i[5];
      cfg.solver_mode = cfg_i[6]; cfg.gn_iters = cfg_i[7];
      cfg.loss_limit = cfg_d[0]; cfg.cov_scale = cfg_d[1]; cfg.regularization = cfg_d[2]; cfg.radius = cfg_d[3];
      std::vector<CellSet> sets(K + 1);
      std::vector<CellSet*> ptrs(K + 1);
      for (int i = 0; i < K; ++i) {
        const int id = kf_ids[(size_t)b * K + i];
        const int o = kf_offsets[id];
        sets[i].n = kf_offsets[id + 1] - o;
        sets[i].mean = kf_mean + 2 * (size_t)o; sets[i].normal = kf_normal + 2 * (size_t)o; sets[i].cov = kf_cov + 4 * (size_t)o;
        sets[i].planarity = kf_planarity + o; sets[i].nsamples = kf_nsamples + o;
        sets[i].build_index();     // the reference builds this kd-tree when the keyframe was a current scan
        ptrs[i] = &sets[i];
      }
      sets[K].n = nc; sets[K].mean = mean.data(); sets[K].normal = normal.data(); sets[K].cov = cov.data();
      sets[K].planarity = plan.data(); sets[K].nsamples = ns.data();
      sets[K].build_index();       // pointnormal.cpp:86 (built for every scan)
      ptrs[K] = &sets[K];
      std::vector<double> p(poses + (size_t)b * (K + 1) * 3, poses + (size_t)(b + 1) * (K + 1) * 3);
      RegStats st;
      do_register(cfg, ptrs, p, cov36 + 36 * (size_t)b, st, nullptr);
      for (int i = 0; i < 3 * (K + 1); ++i) poses[(size_t)b * (K + 1) * 3 + i] = p[i];
      if (stats) std::memcpy((char*)stats + sizeof(RegStats) * (size_t)b, &st, sizeof(RegStats));
      auto t3 = std::chrono::steady_clock::now();
      t_filter += std::chrono::duration_cast<std::chrono::nanoseconds>(t1 - t0).count();
      t_normals += std::chrono::duration_cast<std::chrono::nanoseconds>(t2 - t1).count();
      t_reg += std::chrono::duration_cast<std::chrono::nanoseconds>(t3 - t2).count();
    }
  };
  if (nthreads <= 1) worker();
  else {
    std::vector<std::thread> th;
    for (int t = 0; t < nthreads; ++t) th.emplace_back(worker);
    for (auto& t : th) t.join();
  }
  if (stage_ms) { stage_ms[0] = t_filter.load() * 1e-6; stage_ms[1] = t_normals.load() * 1e-6; stage_ms[2] = t_reg.load() * 1e-6; }
  return 0;
}

}  // extern "C"

// ----------------------------------------------------------------------------
// Sequence replay: OdometryKeyframeFuser::processFrame semantics
// (odometrykeyframefuser.cpp:143-259, 62-94, 470-494) on top of the per-scan
// path above, one sequence, one scan at a time -- the execution model of
// offline_odometry (src/offline_odometry.cpp:73-127).  Planar rigid transforms
// are kept as the 2x3 matrices an Eigen::Affine3d product would hold.
// ----------------------------------------------------------------------------
namespace {
struct T2 { double r00 = 1, r01 = 0, r10 = 0, r11 = 1, x = 0, y = 0; };
T2 t2_from(double x, double y, double yaw) { T2 t; t.r00 = std::cos(yaw); t.r01 = -std::sin(yaw); t.r10 = std::sin(yaw); t.r11 = std::cos(yaw); t.x = x; t.y = y; return t; }
T2 t2_mul(const T2& a, const T2& b) {
  T2 r;
  r.r00 = a.r00 * b.r00 + a.r01 * b.r10; r.r01 = a.r00 * b.r01 + a.r01 * b.r11;
  r.r10 = a.r10 * b.r00 + a.r11 * b.r10; r.r11 = a.r10 * b.r01 + a.r11 * b.r11;
  r.x = a.r00 * b.x + a.r01 * b.y + a.x; r.y = a.r10 * b.x + a.r11 * b.y + a.y;
  return r;
}
T2 t2_inv(const T2& a) {
  T2 r; r.r00 = a.r00; r.r01 = a.r10; r.r10 = a.r01; r.r11 = a.r11;
  r.x = -(r.r00 * a.x + r.r01 * a.y); r.y = -(r.r10 * a.x + r.r11 * a.y);
  return r;
}
double t2_yaw(const T2& a) { return std::atan2(a.r10, a.r11); }   // utils.cpp:115-122 (eulerAngles(0,1,2)[2], planar)

struct OwnedCells {
  std::vector<double> mean, normal, cov, plan; std::vector<int32_t> ns; CellSet cs;
  void finish(int n) { cs.n = n; cs.mean = mean.data(); cs.normal = normal.data(); cs.cov = cov.data(); cs.planarity = plan.data(); cs.nsamples = ns.data(); cs.build_index(); }
};
}  // namespace

extern "C" {
// pipe_i: A, R, k, z_min, weight_intensity, compensate, ccw, submap_scan_size, use_guess
// pipe_f: min_distance, range_res, radius(res)
// kf_d:   min_keyframe_dist, min_keyframe_rot_deg
int orc_odometry_sequence(int nscans, const int32_t* pipe_i, const float* pipe_f, const double* kf_d,
                          const int32_t* cfg_i, const double* cfg_d, const uint8_t* polar,
                          double* poses_out /*nscans*3*/, int32_t* keyframe_out /*nscans*/, void* stats_out /*nscans RegStats*/,
                          int32_t* ncells_out /*nscans*/) {
  const int A = pipe_i[0], R = pipe_i[1], k = pipe_i[2], zmin = pipe_i[3];
  const int wint = pipe_i[4], comp = pipe_i[5], ccw = pipe_i[6], submap = pipe_i[7], use_guess = pipe_i[8];
  RegCfg cfg;
  cfg.cost = cfg_i[0]; cfg.loss = cfg_i[1]; cfg.weight_opt = cfg_i[2];
  cfg.max_outer = cfg_i[3]; cfg.min_outer = cfg_i[4]; cfg.max_inner = cfg_i[5];
  cfg.solver_mode = cfg_i[6]; cfg.gn_iters = cfg_i[7];
  cfg.loss_limit = cfg_d[0]; cfg.cov_scale = cfg_d[1]; cfg.regularization = cfg_d[2]; cfg.radius = cfg_d[3];
  T2 T_prev, Tmot, Tcurrent;
  std::vector<std::pair<T2, std::shared_ptr<OwnedCells>>> keyframes;
  std::vector<int32_t> idx((size_t)A * k), cnt(A);
  std::vector<float> cloud((size_t)A * k * 4);
  for (int s = 0; s < nscans; ++s) {
    const uint8_t* img = polar + (size_t)s * A * R;
    orc_kstrongest(img, A, R, zmin, k, idx.data(), cnt.data());
    const int n = orc_cloud(img, A, R, k, idx.data(), cnt.data(), pipe_f[0], pipe_f[1], cloud.data());
    const T2 TprevMot = Tmot;                                              // :146
    if (comp) { const double mot[3] = {TprevMot.x, TprevMot.y, t2_yaw(TprevMot)}; orc_compensate(cloud.data(), n, mot, ccw); }
    auto cur = std::make_shared<OwnedCells>();
    cur->mean.resize(2 * (size_t)n + 2); cur->normal.resize(2 * (size_t)n + 2); cur->cov.resize(4 * (size_t)n + 4);
    cur->plan.resize(n + 1); cur->ns.resize(n + 1);
    std::vector<double> avg(n + 1);
    const int nc = orc_surface_points(cloud.data(), n, pipe_f[2], 1.0, wint, 0.0, 0.0, cur->mean.data(), cur->normal.data(),
                                      cur->cov.data(), cur->plan.data(), cur->ns.data(), avg.data(), nullptr);   // :161
    cur->finish(nc);
    if (ncells_out) ncells_out[s] = nc;
    RegStats st = RegStats();
    keyframe_out[s] = 0;
    const T2 Tguess = use_guess ? t2_mul(T_prev, TprevMot) : T_prev;       // :164-168
    if (keyframes.empty()) {                                               // :171-177
      keyframes.push_back(std::make_pair(T2(), cur));
      keyframe_out[s] = 1;
      poses_out[3 * s] = Tcurrent.x; poses_out[3 * s + 1] = Tcurrent.y; poses_out[3 * s + 2] = t2_yaw(Tcurrent);
      if (stats_out) std::memcpy((char*)stats_out + sizeof(RegStats) * (size_t)s, &st, sizeof(RegStats));
      continue;
    }
    std::vector<CellSet*> scans; std::vector<double> poses;                // FormatScans :478-494
    for (auto& kf : keyframes) { scans.push_back(&kf.second->cs); poses.push_back(kf.first.x); poses.push_back(kf.first.y); poses.push_back(t2_yaw(kf.first)); }
    scans.push_back(&cur->cs); poses.push_back(Tguess.x); poses.push_back(Tguess.y); poses.push_back(t2_yaw(Tguess));
    double cov36[36];
    const std::vector<double> poses_in = poses;
    do_register(cfg, scans, poses, cov36, st, nullptr);                    // :186 (return value ignored :184-186)
    // Tsrc is rewritten from the parameters after every usable solve (n_scan_normal.cpp:119-121,177-178) and keeps the
    // pose of the last one if a later outer iteration fails
    const bool wrote = st.pose_written != 0;
    const size_t L = poses.size() - 3;
    Tcurrent = wrote ? t2_from(poses[L], poses[L + 1], poses[L + 2]) : Tguess;       // :195
    const T2 Tmot_current = t2_mul(t2_inv(T_prev), Tcurrent);
    {                                                                      // AccelerationVelocitySanityCheck :76-94, :197-199
      const double dt = 0.25;
      const double vel = std::sqrt(Tmot_current.x * Tmot_current.x + Tmot_current.y * Tmot_current.y) / dt;
      const double ax = (Tmot_current.x - Tmot.x) / (dt * dt), ay = (Tmot_current.y - Tmot.y) / (dt * dt);
      if (std::sqrt(ax * ax + ay * ay) > 200 || vel > 200) Tcurrent = Tguess;
    }
    Tmot = t2_mul(t2_inv(T_prev), Tcurrent);                               // :200
    const T2 Tkeydiff = t2_mul(t2_inv(keyframes.back().first), Tcurrent);  // :227
    const bool fuse = std::sqrt(Tkeydiff.x * Tkeydiff.x + Tkeydiff.y * Tkeydiff.y) > kf_d[0] ||
                      std::fabs(t2_yaw(Tkeydiff)) > kf_d[1] * M_PI / 180.0;   // :62-73
    if (fuse) {                                                            // :234-249, AddToReference :470-476
      keyframes.push_back(std::make_pair(Tcurrent, cur));
      if ((int)keyframes.size() > submap) keyframes.erase(keyframes.begin());
      keyframe_out[s] = 1;
    }
    T_prev = Tcurrent;                                                     // :257
    poses_out[3 * s] = Tcurrent.x; poses_out[3 * s + 1] = Tcurrent.y; poses_out[3 * s + 2] = t2_yaw(Tcurrent);
    if (stats_out) std::memcpy((char*)stats_out + sizeof(RegStats) * (size_t)s, &st, sizeof(RegStats));
  }
  return 0;
}
}  // extern "C"

// ----------------------------------------------------------------------------
// CA-CFAR: AzimuthCACFAR::getFilteredPointCloud (cfar.cpp:35-83), getMean (:73-83),
// getCAScalingFactor (:12-16), constructed as in radar_driver.cpp:54.
// ----------------------------------------------------------------------------
extern "C" int orc_cfar(const uint8_t* img, int A, int R, int window_size, double false_alarm_rate, int nb_guard_cells,
                        float range_res_f, float static_threshold_f, float min_distance_f, double max_distance,
                        float* xyzi_out, int capacity) {
  const double range_resolution = (double)range_res_f, static_threshold = (double)static_threshold_f, min_distance = (double)min_distance_f;
  const double N = (double)(window_size * 2);
  const double scaling_factor = N * (std::pow(false_alarm_rate, -1. / N) - 1.);
  auto get_mean = [&](const uint8_t* az, int start_idx, int end_idx) {
    double sum = 0., n = 0.;
    for (int i = start_idx; i < end_idx; i++) { sum += std::pow(double(az[i]), 2.); n += 1.; }
    return sum / n;
  };
  int n = 0;
  for (int azimuth_nb = 0; azimuth_nb < A; azimuth_nb++) {
    const uint8_t* az = img + (size_t)azimuth_nb * R;
    const double theta = (double(azimuth_nb + 1) / A) * 2. * M_PI;
    for (int range_bin = 0; range_bin < R; range_bin++) {
      const double range = range_resolution * double(range_bin);
      const double intensity = double(az[range_bin]);
      if (range > min_distance && range < max_distance && intensity > static_threshold) {
        const int ts = std::max(0, range_bin - nb_guard_cells - window_size), te = range_bin - nb_guard_cells;
        const double trailing_mean = get_mean(az, ts, te);
        const int fs = range_bin + nb_guard_cells, fe = std::min(R, range_bin + nb_guard_cells + window_size);
        const double forwarding_mean = get_mean(az, fs, fe);
        const double mean = (trailing_mean + forwarding_mean) / 2.0;
        const double threshold = scaling_factor * mean;
        if (std::pow(intensity, 2.) > threshold) {
          if (n < capacity) {
            float* p = xyzi_out + 4 * (size_t)n;
            p[0] = (float)(range * std::cos(theta)); p[1] = (float)(range * std::sin(theta)); p[2] = 0.f; p[3] = (float)intensity;
          }
          ++n;
        }
      }
    }
  }
  return n;
}
