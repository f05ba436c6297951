// K7: azimuth CA-CFAR detector -- the reference's alternative filter (--filter-type CA-CFAR),
// AzimuthCACFAR::getFilteredPointCloud (src/cfear_radarodometry/cfar.cpp:35-83) as called by
// radarDriver::Process (radar_driver.cpp:52-56).
//
// One CTA per azimuth row: the row goes to shared memory once, an exclusive prefix sum of the squared
// intensities (exact in uint32: 255^2 * R < 2^32 for R <= 65535) turns both window means into two differences,
// so every bin is tested with the reference's own double arithmetic: mean = (S_t/N_t + S_f/N_f)/2,
// threshold = scaling*mean, detect iff I^2 > threshold (empty windows give 0/0 = NaN and never detect, like the
// reference).  Pass 0 counts detections per row, k7_offsets scans the counts per image, pass 1 writes the points in
// (row, bin) order -- the reference's push_back order.
#pragma once
#include "common.cuh"

namespace cfear {

struct CfarParams {
  const uint8_t* polar;     // [nscans][A][R]
  int nrows, A, R;
  int window, guard;
  double scaling;           // N (alpha^(-1/N) - 1), N = 2*window  (cfar.cpp:12-16,31), host libm
  double range_res, static_threshold, min_distance, max_distance;
  const double2* cs;        // [A] (cos, sin) of theta = 2 pi (a+1)/A
  int32_t* rowcnt;          // [nrows]
  const int32_t* rowoff;    // [nrows] exclusive offsets inside each image (pass 1)
  float4* cloud;            // [nscans][cap]
  int cap;
  int pass;
};

constexpr int K7_THREADS = 256;

__global__ void __launch_bounds__(K7_THREADS) k7_cfar(const CfarParams p) {
  extern __shared__ uint32_t s_pre[];            // [R+1] exclusive prefix of squares, then reused flags
  __shared__ int s_warp[33];
  const int row = blockIdx.x;
  const int tid = threadIdx.x, T = blockDim.x;
  const int R = p.R;
  const uint8_t* img = p.polar + (size_t)row * R;
  for (int r = tid; r < R; r += T) { const uint32_t v = img[r]; s_pre[r] = v * v; }
  if (tid == 0) s_pre[R] = 0;
  __syncthreads();
  block_array_excl_scan(reinterpret_cast<int*>(s_pre), R + 1, s_warp);   // s_pre[i] = sum_{j<i} I_j^2
  const int a = row % p.A;
  const double2 cs = p.cs[a];
  const int scan = row / p.A;
  // each thread owns a contiguous chunk of bins so that the ordered write needs one block scan
  const int chunk = (R + T - 1) / T;
  const int lo = min(tid * chunk, R), hi = min(lo + chunk, R);
  auto detect = [&](int b) -> bool {
    const double range = p.range_res * (double)b;                        // cfar.cpp:44
    const double intensity = (double)img[b];
    if (!(range > p.min_distance && range < p.max_distance && intensity > p.static_threshold)) return false;   // :46
    const int ts = max(0, b - p.guard - p.window), te = b - p.guard;     // :49-50
    const int fs = b + p.guard, fe = min(R, b + p.guard + p.window);     // :53-54
    const double nt = te > ts ? (double)(te - ts) : 0.0, nf = fe > fs ? (double)(fe - fs) : 0.0;
    const double st = te > ts ? (double)(s_pre[te] - s_pre[ts]) : 0.0, sf = fe > fs ? (double)(s_pre[fe] - s_pre[fs]) : 0.0;
    const double mean = (st / nt + sf / nf) / 2.0;                       // :51,55,57  (0/0 -> NaN -> no detection)
    const double threshold = p.scaling * mean;                           // :59
    return intensity * intensity > threshold;                            // :60-61
  };
  int n = 0;
  for (int b = lo; b < hi; ++b) n += detect(b) ? 1 : 0;
  int total;
  int base = block_excl_scan(n, s_warp, &total);
  if (p.pass == 0) {
    if (tid == 0) p.rowcnt[row] = total;
    return;
  }
  float4* out = p.cloud + (size_t)scan * p.cap;
  base += p.rowoff[row];
  for (int b = lo; b < hi; ++b) {
    if (detect(b)) {
      if (base < p.cap) {
        const double range = p.range_res * (double)b;
        out[base] = make_float4((float)(range * cs.x), (float)(range * cs.y), 0.f, (float)img[b]);   // :63-67
      }
      ++base;
    }
  }
}

// per image: exclusive scan of the row counts -> row offsets, total
__global__ void __launch_bounds__(512) k7_offsets(const int32_t* rowcnt, int A, int32_t* rowoff, int32_t* npts) {
  extern __shared__ int s_off[];
  __shared__ int s_warp[33];
  const int scan = blockIdx.x;
  for (int a = threadIdx.x; a < A; a += blockDim.x) s_off[a] = rowcnt[(size_t)scan * A + a];
  __syncthreads();
  const int n = block_array_excl_scan(s_off, A, s_warp);
  for (int a = threadIdx.x; a < A; a += blockDim.x) rowoff[(size_t)scan * A + a] = s_off[a];
  if (threadIdx.x == 0) npts[scan] = n;
}

}  // namespace cfear
