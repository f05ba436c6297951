// K2/K3/K4: cloud assembly (+ motion compensation), voxel centroids, oriented surface points and the
// nearest-neighbour index over the cell means -- one CTA per scan, everything between the row-cloud
// read and the cell-set write stays in shared memory.
//
// Replaces, for one scan:
//   Compensate(cloud, Tmot, ccw)                         utils.cpp:96-113, utils.h:28-32
//   MapPointNormal::MapPointNormal / ComputeNormals      pointnormal.cpp:65-90, 265-297
//     pcl::VoxelGrid (leaf = radius/downsample_factor)   pointnormal.cpp:277-280
//     kd-tree radiusSearch(centroid, radius) >= 6        pointnormal.cpp:291
//     cell::cell + cell::ComputeNormal                   pointnormal.cpp:7-63
//   MapPointNormal::ComputeSearchTreeFromCells           pointnormal.cpp:151-162
//
// Data flow (smem unless the cloud is too large, then the same code runs on global scratch):
//   bufA <- TMA bulk copy (cp.async.bulk + mbarrier) of the scan's padded row cloud [A][k] float4
//   bufA: in-place left compaction (row order = the reference's push_back order) + Compensate
//   hist: voxel histogram -> exclusive scan -> scatter bufA -> bufB (bucketed by voxel)
//   per voxel: order by input index, fp32 centroid (sequential sum like VoxelGrid) -> cxy list (in bufA)
//   per centroid: exact fp32 radius test over the 3x3 voxel neighbourhood in bufB, fp64 weighted
//   mean / covariance / closed-form 2x2 eigen -> validity -> ordered compaction into the cell slot
//   fp32 means -> bucket grid (counting sort through hist) -> gstart / gxy / gidx of the slot
#pragma once
#include <type_traits>
#include <cstdio>
#include <cuda_fp16.h>
#include "common.cuh"
#include "tma.cuh"

namespace cfear {

#ifndef CFEAR_K3_THREADS
#define CFEAR_K3_THREADS 512   // 16 warps, 64 registers: two CTAs per SM (109 KB of shared memory each), 256 scans in one wave.  Alone:
                               // 384 x 2: 0.112 ms / 256 scans, 512 x 2: 0.105, 768 x 1: 0.123, 256 x 2: 0.129; with four steps in flight
                               // the whole step takes 0.379 / 0.373 / 0.398 ms (profiles/r02e_inflight_k3_ab.txt)
#endif
#ifndef CFEAR_K3_MINBLOCKS
#define CFEAR_K3_MINBLOCKS 2
#endif
constexpr int K3_THREADS = CFEAR_K3_THREADS;
#ifndef CFEAR_K3_THREADS_WIDE
#define CFEAR_K3_THREADS_WIDE 1024
#endif
constexpr int K3_THREADS_WIDE = CFEAR_K3_THREADS_WIDE;
#ifndef CFEAR_K3_LPC
#define CFEAR_K3_LPC 2
#endif
constexpr int K3_LANES_PER_CENTROID = CFEAR_K3_LPC;   // 8 -> 4: 0.164 -> 0.142 ms (fixed cost per centroid amortised over twice as many per warp); 2: 0.140, 1: 0.160, 16: 0.218
constexpr int K3_HIST_CAP = 16384;        // voxel / NN-grid bins kept in shared memory

struct K3Params {
  // input mode 0: padded row clouds from K1;  mode 1: ready clouds (already compensated) in `cloud`
  int mode;
  int A, k;                    // rows per scan, slots per row (mode 0)
  const float4* rowcloud;      // [nscans][A][k]
  const int32_t* rowcnt;       // [nscans][A]
  const double* mot;           // [nscans][3] previous motion (x, y, yaw) or nullptr = no compensation
  const double2* cs;           // [A] (cos, sin) of the row angles 2 pi (a+1) / A (the table K1 uses), mode 0
  int ccw;
  float4* cloud;               // [nscans][cap_pts] out (mode 0, may be null) / in (mode 1)
  int32_t* npts;               // [nscans] out (mode 0) / in (mode 1)
  int cap_pts;                 // capacity of one cloud (A*k)
  const int32_t* slots;        // [nscans] destination cell-set slot
  float radius;                // "res"
  float leaf;                  // float(radius / downsample_factor)
  int weight_intensity;
  double origin_x, origin_y;
  float nn_cell;               // bucket size of the NN grid (metres)
  int pts_in_smem;             // 1: bufA/bufB in shared memory
  float4* g_bufA; float4* g_bufB;   // [nscans][cap_pts] global fallback
  int* g_hist; int g_hist_cap;      // [nscans][g_hist_cap+1] global fallback for large grids
  int32_t* status;             // [nscans] 0 ok, 1 voxel grid over capacity
  double2* cell_tmp;           // [nscans][cap_pts][4] per-centroid raw moments (scratch)
  CellPool pool;
};

// utils.h:28-32 GetRelTimeStamp.  (a / 2 pi is formed as a * (1 / 2 pi): an ulp of the time stamp moves a compensated
// coordinate by 1e-16 m; Compensate is FP64-issue bound and the division is a fifth of its instructions.)
__device__ __forceinline__ double rel_time_stamp(double x, double y, bool ccw) {
  const double a = atan2(y, x);
  const double d = (a > 0.00001 ? a : (2 * M_PI + a)) * (1.0 / (2 * M_PI));
  return ccw ? -(d - 0.5) : (d - 0.5);
}

// The same time stamp for a point of azimuth row a (mode 0): the point is the fp32 rounding of rho (cos, sin)(theta_a),
// so atan2(y, x) = theta_a + eps with eps = cross((cos, sin)(theta_a), (x, y)) / |(x, y)| to first order (|eps| ~ 1e-7,
// the cubic term is 1e-22): a dozen FP64 instructions instead of the ~120 of atan2, accurate to ~1e-14 rad -- a
// compensated coordinate moves by < 1e-13 m, eight orders below an fp32 ulp at these ranges.  theta_a lies in (0, 2 pi],
// which is already the interval utils.h:30 maps atan2's result to (its 1e-5 threshold is never reached: theta_1 = 0.0157).
#ifndef CFEAR_K3_ATAN2
__device__ __forceinline__ double rel_time_stamp_row(double x, double y, double2 cs, double theta, bool ccw) {
  const double cross = fma(cs.x, y, -(cs.y * x));
  const double r2 = fma(x, x, y * y);
  double ri; asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(ri) : "d"(r2));
  const double h = 0.5 * r2;
  double e = fma(-h * ri, ri, 0.5); ri = fma(ri, e, ri);
  e = fma(-h * ri, ri, 0.5); ri = fma(ri, e, ri);
  const double d = fma(cross, ri, theta) * (1.0 / (2 * M_PI));
  return ccw ? -(d - 0.5) : (d - 0.5);
}
#endif

// closed-form symmetric 2x2 eigen decomposition: (a b; b d) -> ascending eigenvalues, unit eigenvectors
struct Eig2 { double lmin, lmax, nx, ny; };
__device__ __forceinline__ Eig2 eig2_sym(double a, double b, double d) {
  Eig2 e;
  const double t = 0.5 * (a - d), m = 0.5 * (a + d);
  const double h = sqrt(t * t + b * b);
  e.lmax = m + h; e.lmin = m - h;
  double vx, vy;
  if (h == 0.0) { vx = 1.0; vy = 0.0; }
  else if (t >= 0.0) { vx = t + h; vy = b; }
  else { vx = b; vy = h - t; }
  const double nrm = sqrt(vx * vx + vy * vy);
  if (nrm > 0.0) { const double inrm = 1.0 / nrm; vx *= inrm; vy *= inrm; } else { vx = 1.0; vy = 0.0; }
  e.nx = -vy; e.ny = vx;
  return e;
}

__device__ __forceinline__ float block_min_f(float v, float* s_red) {
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) v = fminf(v, __shfl_xor_sync(FULL, v, d));
  __syncthreads();
  if (lane_id() == 0) s_red[warp_id()] = v;
  __syncthreads();
  float r = s_red[0];
  for (int w = 1; w < (int)(blockDim.x >> 5); ++w) r = fminf(r, s_red[w]);
  return r;
}
__device__ __forceinline__ float block_max_f(float v, float* s_red) { return -block_min_f(-v, s_red); }

// Bounding box of the block's points in one reduction: (min x, min y, max x, max y) -- the maxima travel negated, so one
// min tree serves all four.  s_red4: >= 4 * 32 floats.  Two barriers instead of the eight of four separate reductions.
__device__ __forceinline__ void block_bounds(float& mnx, float& mny, float& mxx, float& mxy, float* s_red4) {
  float v0 = mnx, v1 = mny, v2 = -mxx, v3 = -mxy;
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) {
    v0 = fminf(v0, __shfl_xor_sync(FULL, v0, d)); v1 = fminf(v1, __shfl_xor_sync(FULL, v1, d));
    v2 = fminf(v2, __shfl_xor_sync(FULL, v2, d)); v3 = fminf(v3, __shfl_xor_sync(FULL, v3, d));
  }
  __syncthreads();
  if (lane_id() == 0) { float* o = s_red4 + 4 * warp_id(); o[0] = v0; o[1] = v1; o[2] = v2; o[3] = v3; }
  __syncthreads();
  const int nw = (int)(blockDim.x >> 5), l = lane_id();
  const float4 r = l < nw ? *reinterpret_cast<const float4*>(s_red4 + 4 * l) : make_float4(3.0e38f, 3.0e38f, 3.0e38f, 3.0e38f);
  v0 = r.x; v1 = r.y; v2 = r.z; v3 = r.w;
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) {
    v0 = fminf(v0, __shfl_xor_sync(FULL, v0, d)); v1 = fminf(v1, __shfl_xor_sync(FULL, v1, d));
    v2 = fminf(v2, __shfl_xor_sync(FULL, v2, d)); v3 = fminf(v3, __shfl_xor_sync(FULL, v3, d));
  }
  mnx = v0; mny = v1; mxx = -v2; mxy = -v3;
}

// Histogram / prefix-sum table with 16-bit entries (counts and offsets are bounded by the points of one scan, <= 65535),
// two per 32-bit word so that shared-memory atomics can update them: half the footprint of an int table.
struct Hist16 {
  uint32_t* w;
  __device__ __forceinline__ int get(int v) const { return reinterpret_cast<const uint16_t*>(w)[v]; }
  __device__ __forceinline__ void set(int v, int x) const { reinterpret_cast<uint16_t*>(w)[v] = (uint16_t)x; }
  // += 1, returns the previous value
  __device__ __forceinline__ int inc(int v) const {
    const int sh = (v & 1) << 4;
    return (int)((atomicAdd(&w[v >> 1], 1u << sh) >> sh) & 0xffffu);
  }
  // entries [0, n) <- 0 (whole words: entry n, if it shares the last word, is cleared too)
  __device__ __forceinline__ void zero(int n) const {
    for (int i = threadIdx.x; i < (n + 1) >> 1; i += blockDim.x) w[i] = 0u;
  }
  // in-place exclusive scan of entries [0, n); returns the total.  Whole block, contains __syncthreads().
  __device__ inline int excl_scan(int n, int* s_warp) const {
    const int T = blockDim.x;
    const int chunk = (((n + T - 1) / T) + 1) & ~1;                 // even: no two threads share a word
    const int lo = min((int)threadIdx.x * chunk, n), hi = min(lo + chunk, n);
    int s = 0;
    for (int i = lo; i < hi; ++i) s += get(i);
    int total;
    int base = block_excl_scan(s, s_warp, &total);
    for (int i = lo; i < hi; ++i) { const int t = get(i); set(i, base); base += t; }
    __syncthreads();
    return total;
  }
};

// Build the bucket grid over n fp32 means (fm in shared or global memory) into slot arrays.
// hist: >= hist_cap+1 ints of scratch.  Whole block participates.
__device__ inline void build_nn_grid(const CellPool& pool, int slot, const float2* fm, int n, float nn_cell,
                                     Hist16 hist, int hist_cap, int* s_warp, float* s_red) {
  const int tid = threadIdx.x, T = blockDim.x;
  float mnx = 3.0e38f, mny = 3.0e38f, mxx = -3.0e38f, mxy = -3.0e38f;
  for (int i = tid; i < n; i += T) {
    const float2 p = fm[i];
    mnx = fminf(mnx, p.x); mxx = fmaxf(mxx, p.x); mny = fminf(mny, p.y); mxy = fmaxf(mxy, p.y);
  }
  block_bounds(mnx, mny, mxx, mxy, s_red);
  if (n == 0) { mnx = mny = mxx = mxy = 0.f; }
  float g = nn_cell;
  int nx, ny;
  for (;;) {
    nx = (int)floorf((mxx - mnx) / g) + 1; ny = (int)floorf((mxy - mny) / g) + 1;
    if ((long long)nx * ny <= hist_cap) break;
    g *= 2.f;
  }
  const float inv = 1.0f / g;
  const int nb = nx * ny;
  hist.zero(nb + 1);
  __syncthreads();
  for (int i = tid; i < n; i += T) {
    const float2 p = fm[i];
    int bx = (int)floorf((p.x - mnx) * inv), by = (int)floorf((p.y - mny) * inv);
    bx = min(max(bx, 0), nx - 1); by = min(max(by, 0), ny - 1);
    hist.inc(bx + by * nx);
  }
  __syncthreads();
  hist.excl_scan(nb + 1, s_warp);                   // hist[b] = start of bucket b, hist[nb] = n
  uint16_t* gstart = pool.gstart + (size_t)slot * pool.grid_stride;
  for (int b = tid; b <= nb; b += T) gstart[b] = (uint16_t)hist.get(b);
  __syncthreads();
  float4* gpt = pool.gpt + (size_t)slot * pool.max_cells;
  for (int i = tid; i < n; i += T) {
    const float2 p = fm[i];
    int bx = (int)floorf((p.x - mnx) * inv), by = (int)floorf((p.y - mny) * inv);
    bx = min(max(bx, 0), nx - 1); by = min(max(by, 0), ny - 1);
    const int pos = hist.inc(bx + by * nx);
    // .w: the cell's normal as two fp16 -- lets the registration's 30 degree normal gate decide from shared memory in
    // all but the borderline cases (k5_register.cuh, normal_gate)
    const double2 nrm = pool.normal[(size_t)slot * pool.max_cells + i];
    const __half2 h = __floats2half2_rn((float)nrm.x, (float)nrm.y);
    gpt[pos] = make_float4(p.x, p.y, __int_as_float(i), __uint_as_float(*reinterpret_cast<const unsigned*>(&h)));
  }
  if (tid == 0) {
    NNGrid G; G.ox = mnx; G.oy = mny; G.inv_g = inv; G.g = g; G.nx = nx; G.ny = ny;
    pool.grid[slot] = G;
  }
  __syncthreads();
}

// -DCFEAR_K3_PROFILE: clock64 phase probes, printed by thread 0 of a few scans
#ifdef CFEAR_K3_PROFILE
#define K3P(i) if (threadIdx.x == 0) { k3t[i] = clock64(); }
#else
#define K3P(i)
#endif

// PTS_SMEM (= K3Params::pts_in_smem, fixed per context): a compile-time fact, so that every access to the point buffer and to
// the shared-memory histogram is a shared-memory instruction (LDS / ATOMS) instead of a generic one.
// NT = K3_THREADS_WIDE: the same kernel as one 1024-thread CTA per SM, launched for batches of at most one scan per SM
// (sequence replay with few sequences, the drop-in classes' single-scan calls), where a scan's latency is the step.
// Every loop strides by blockDim.x and no sum depends on the thread count, so the cells are bit-identical.
template <bool PTS_SMEM, int NT = K3_THREADS>
__global__ void __launch_bounds__(NT, NT == K3_THREADS ? CFEAR_K3_MINBLOCKS : 1) k3_surface_points(const K3Params p) {
#ifdef CFEAR_K3_PROFILE
  long long k3t[16];
#endif
  K3P(0)
  extern __shared__ __align__(128) unsigned char dyn_smem[];
  __shared__ int s_warp[33];
  __shared__ __align__(16) float s_red[128];
  __shared__ __align__(8) uint64_t s_bar;
  __shared__ int s_misc[4];

  const int scan = blockIdx.x;
  const int tid = threadIdx.x, T = blockDim.x;
  const int slot = p.slots[scan];
  const int cap = p.cap_pts;

  // bufA (the points, read over and over by the neighbourhood pass) lives in shared memory when a scan's cloud fits;
  // bufB (scatter target of the voxel sort, then the centroid / cell-mean lists: touched a few times per element) is
  // global scratch that stays in L2.  With the 16-bit histogram that is 109 KB per CTA: two CTAs per SM.
  float4* bufA; float4* bufB = p.g_bufB + (size_t)scan * cap; int* s_hist;
  if constexpr (PTS_SMEM) {
    bufA = reinterpret_cast<float4*>(dyn_smem);
    s_hist = reinterpret_cast<int*>(bufA + cap);
  } else {
    bufA = p.g_bufA + (size_t)scan * cap;
    s_hist = reinterpret_cast<int*>(dyn_smem);
  }

  // ---- stage the input into bufA ------------------------------------------------------------------
  int n = 0;
  if (p.mode == 0) {
    const int nslots = p.A * p.k;
    const float4* src = p.rowcloud + (size_t)scan * nslots;
    if constexpr (PTS_SMEM) {
      if (tid == 0) mbar_init(&s_bar, 1);
      __syncthreads();
      if (tid == 0) {
        const uint32_t bytes = (uint32_t)nslots * 16u;
        mbar_expect_tx(&s_bar, bytes);
        tma_bulk_g2s(bufA, src, bytes, &s_bar);
      }
    }
    // row offsets while the copy is in flight: s_hist[a] = exclusive prefix of rowcnt
    const int32_t* rc = p.rowcnt + (size_t)scan * p.A;
    for (int a = tid; a < p.A; a += T) s_hist[a] = rc[a];
    __syncthreads();
    n = block_array_excl_scan(s_hist, p.A, s_warp);
    if constexpr (PTS_SMEM) mbar_wait(&s_bar, 0);
    // in-place left compaction, chunk by chunk (dest index <= source index), + Compensate
    const double* mot = p.mot ? p.mot + 3 * (size_t)scan : nullptr;
    const double m0 = mot ? mot[0] : 0.0, m1 = mot ? mot[1] : 0.0, m2 = mot ? mot[2] : 0.0;
    float4* outc = p.cloud ? p.cloud + (size_t)scan * cap : nullptr;
    for (int s0 = 0; s0 < nslots; s0 += T) {
      const int s = s0 + tid;
      bool have = false; float4 pt = make_float4(0, 0, 0, 0); int dst = 0;
      if (s < nslots) {
        const int a = s / p.k, j = s - a * p.k;
        const int start = s_hist[a];
        const int cnt = ((a + 1 < p.A) ? s_hist[a + 1] : n) - start;
        if (j < cnt) {
          have = true; dst = start + j;
          if constexpr (PTS_SMEM) pt = bufA[s]; else pt = src[s];
          if (mot) {                                       // utils.cpp:96-113
            const double x = (double)pt.x, y = (double)pt.y;
#ifdef CFEAR_K3_ATAN2
            const double d = rel_time_stamp(x, y, p.ccw != 0);
#else
            const double d = rel_time_stamp_row(x, y, p.cs[a], (double)(a + 1) * (2 * M_PI / (double)p.A), p.ccw != 0);
#endif
            double s1, c1; sincos_small(d * m2, &s1, &c1);
            const double tx = c1 * x + (-s1) * y + d * m0;
            const double ty = s1 * x + c1 * y + d * m1;
            pt.x = (float)tx; pt.y = (float)ty;
          }
        }
      }
      __syncthreads();
      if (have) { bufA[dst] = pt; if (outc) outc[dst] = pt; }
    }
    __syncthreads();
    K3P(1)
    if (tid == 0 && p.npts) p.npts[scan] = n;
  } else {
    n = p.npts[scan];
    const float4* src = p.cloud + (size_t)scan * cap;
    if constexpr (PTS_SMEM) {
      if (tid == 0) mbar_init(&s_bar, 1);
      __syncthreads();
      if (n > 0) {
        if (tid == 0) {
          const uint32_t bytes = (uint32_t)n * 16u;
          mbar_expect_tx(&s_bar, bytes);
          tma_bulk_g2s(bufA, src, bytes, &s_bar);
        }
        mbar_wait(&s_bar, 0);
      }
    } else {
      for (int i = tid; i < n; i += T) bufA[i] = src[i];
    }
    __syncthreads();
  }

  if (tid == 0) p.status[scan] = 0;
  if (n == 0) {                                            // reference exits on an empty cloud (pointnormal.cpp:72-75)
    if (tid == 0) {
      p.pool.ncells[slot] = 0;
      NNGrid G; G.ox = G.oy = 0.f; G.g = p.nn_cell; G.inv_g = 1.f / p.nn_cell; G.nx = G.ny = 1;
      p.pool.grid[slot] = G;
      p.pool.gstart[(size_t)slot * p.pool.grid_stride] = 0;
      p.pool.gstart[(size_t)slot * p.pool.grid_stride + 1] = 0;
    }
    return;
  }

  // ---- VoxelGrid: bounds, voxel ids (fp32, restating pcl voxel_grid.hpp applyFilter) ---------------
  float mnx = 3.0e38f, mny = 3.0e38f, mxx = -3.0e38f, mxy = -3.0e38f;
  for (int i = tid; i < n; i += T) {
    const float4 q = bufA[i];
    mnx = fminf(mnx, q.x); mxx = fmaxf(mxx, q.x); mny = fminf(mny, q.y); mxy = fmaxf(mxy, q.y);
  }
  block_bounds(mnx, mny, mxx, mxy, s_red);
  K3P(2)
  const float inv = 1.0f / p.leaf;
  const int minbx = (int)floorf(mnx * inv), maxbx = (int)floorf(mxx * inv);
  const int minby = (int)floorf(mny * inv), maxby = (int)floorf(mxy * inv);
  const int divx = maxbx - minbx + 1, divy = maxby - minby + 1;
  const long long nbins_ll = (long long)divx * divy;
  const bool hist_global = nbins_ll + 1 > K3_HIST_CAP;         // block-uniform; very sparse scans only
  if (hist_global && nbins_ll + 1 > p.g_hist_cap) {
    if (tid == 0) { p.status[scan] = 1; p.pool.ncells[slot] = 0; }
    return;
  }
  // The rest of the kernel exists twice, once per home of the voxel histogram, so that in the usual case the compiler
  // knows it is shared memory (LDS / ATOMS instead of generic loads and atomics in the sort and the neighbourhood pass).
  auto rest = [&](auto HG) {
  Hist16 hist;
  if constexpr (decltype(HG)::value) hist.w = reinterpret_cast<uint32_t*>(p.g_hist + (size_t)scan * (p.g_hist_cap + 1));
  else hist.w = reinterpret_cast<uint32_t*>(s_hist);
  const int nbins = (int)nbins_ll;
  hist.zero(nbins + 1);
  __syncthreads();
  const float fminbx = (float)minbx, fminby = (float)minby;
  for (int i = tid; i < n; i += T) {
    const float4 q = bufA[i];
    const int i0 = (int)(floorf(q.x * inv) - fminbx), i1 = (int)(floorf(q.y * inv) - fminby);
    hist.inc(i0 + i1 * divx);
  }
  __syncthreads();
  K3P(3)
  hist.excl_scan(nbins + 1, s_warp);                       // hist[v] = start of voxel v
  K3P(4)
  for (int i = tid; i < n; i += T) {
    float4 q = bufA[i];
    const int i0 = (int)(floorf(q.x * inv) - fminbx), i1 = (int)(floorf(q.y * inv) - fminby);
    const int pos = hist.inc(i0 + i1 * divx);              // afterwards hist[v] = end of voxel v
    q.z = __int_as_float(i);                               // z is identically 0 on this path: carry the input index
    bufB[pos] = q;
  }
  __syncthreads();
  K3P(5)
  // restore input order inside every voxel (VoxelGrid sums a voxel's points in a fixed order; the oracle's is
  // ascending input index): rank by counting, one thread per point, neighbours in a warp share the segment
  for (int a = tid; a < n; a += T) {
    const float4 q = bufB[a];
    const int i0 = (int)(floorf(q.x * inv) - fminbx), i1 = (int)(floorf(q.y * inv) - fminby);
    const int v = i0 + i1 * divx;
    const int s = v ? hist.get(v - 1) : 0, e = hist.get(v);
    const int me = __float_as_int(q.z);
    int rank = 0;
    for (int b = s; b < e; ++b) rank += (__float_as_int(bufB[b].z) < me);
    bufA[s + rank] = q;
  }
  __syncthreads();
  K3P(6)
  float4* pts = bufA;                                      // points bucketed by voxel, input order inside a voxel

  // ---- ordered list of non-empty voxels, sequential fp32 centroid per voxel ---------------------------
  float2* cxy = reinterpret_cast<float2*>(bufB);           // bufB is free now: centroid list, later fp32 cell means
  int* vlist = reinterpret_cast<int*>(cxy + cap);          // non-empty voxel ids, ascending
  {
    const int chunk = (nbins + T - 1) / T;
    const int lo = min(tid * chunk, nbins), hi = min(lo + chunk, nbins);
    int nonempty = 0;
    for (int v = lo; v < hi; ++v) nonempty += (hist.get(v) > (v ? hist.get(v - 1) : 0));
    int total;
    int base = block_excl_scan(nonempty, s_warp, &total);
    for (int v = lo; v < hi; ++v)
      if (hist.get(v) > (v ? hist.get(v - 1) : 0)) vlist[base++] = v;
    if (tid == 0) s_misc[0] = total;
    __syncthreads();
  }
  K3P(7)
  const int nvox = s_misc[0];
  for (int c = tid; c < nvox; c += T) {
    const int v = vlist[c];
    const int s = v ? hist.get(v - 1) : 0, e = hist.get(v);
    float sx = 0.f, sy = 0.f;
    for (int a = s; a < e; ++a) { sx += pts[a].x; sy += pts[a].y; }
    const float cnt = (float)(e - s);
    cxy[c] = make_float2(sx / cnt, sy / cnt);
  }
  __syncthreads();

  // ---- per centroid: radius neighbourhood -> cell.  LPC (4) lanes share one centroid and groups pull centroids
  // from a shared counter (dense neighbourhoods cluster in voxel order, static assignment would idle most
  // of the block).  One pass accumulates the weight / first / second moments about the centroid q
  // (exact in fp64: q and the points are fp32), from which the weighted mean and the central covariance of
  // pointnormal.cpp:21-33 follow.  Cells land in a per-scan scratch at their centroid index; an ordered
  // compaction of the valid ones follows. ---------------------------------------------------------------
  K3P(8)
  const float r = p.radius;
  const float r2 = (float)((double)r * (double)r);
  const float rq = r * 1.0001f + 1e-4f;                    // bin-range margin (the d2 test itself is exact)
  const size_t cbase = (size_t)slot * p.pool.max_cells;
  constexpr int LPC = K3_LANES_PER_CENTROID;               // lanes sharing one centroid (power of two <= 32)
  const int sl = tid & (LPC - 1);
  const int gleader = lane_id() & ~(LPC - 1);
  const unsigned gmask = (LPC == 32 ? 0xffffffffu : ((1u << LPC) - 1u)) << gleader;
  const bool wint = p.weight_intensity != 0;
  double2* tmp = p.cell_tmp + (size_t)scan * cap * 4;                        // [cap][4] double2: raw moments
  if (tid == 0) s_misc[1] = 0;
  __syncthreads();
  for (;;) {
    int c = 0;
    if (sl == 0) c = atomicAdd(&s_misc[1], 1);
    c = __shfl_sync(gmask, c, gleader);
    if (c >= nvox) break;
    const float2 q = cxy[c];
    int bx0 = (int)(floorf((q.x - rq) * inv) - fminbx), bx1 = (int)(floorf((q.x + rq) * inv) - fminbx);
    int by0 = (int)(floorf((q.y - rq) * inv) - fminby), by1 = (int)(floorf((q.y + rq) * inv) - fminby);
    bx0 = max(bx0, 0); by0 = max(by0, 0); bx1 = min(bx1, divx - 1); by1 = min(by1, divy - 1);
    int gN = 0;
    double S0 = 0.0, S1x = 0.0, S1y = 0.0, Sxx = 0.0, Sxy = 0.0, Syy = 0.0;
    auto take = [&](int a) {
      const float4 pt = pts[a];
      const float dx = q.x - pt.x, dy = q.y - pt.y;
      float d2 = dx * dx; d2 += dy * dy;                    // FLANN L2_Simple in fp32, strict d2 < r2
      if (d2 < r2) {
        ++gN;
        // pointnormal.cpp:15  max(intensity - 60, 0): intensities are small integers, so the fp32 form is the same number
        const double w = wint ? (double)fmaxf(pt.w - 60.0f, 0.0f) : 1.0;
        const double ex = (double)pt.x - (double)q.x, ey = (double)pt.y - (double)q.y;
        const double wx = w * ex, wy = w * ey;
        S0 += w; S1x += wx; S1y += wy; Sxx = fma(wx, ex, Sxx); Sxy = fma(wx, ey, Sxy); Syy = fma(wy, ey, Syy);
      }
    };
#ifndef CFEAR_K3_NESTED
    if (by1 - by0 <= 2) {
      // the usual 3 x 3 voxel neighbourhood (leaf = radius): the point runs of the (up to) three voxel rows are walked as ONE
      // flattened range, so the groups of a warp -- each on its own centroid -- stay in a single loop instead of diverging
      // at every row boundary
      const bool h1 = by0 + 1 <= by1, h2 = by0 + 2 <= by1;
      const int o0 = by0 * divx, o1 = o0 + divx, o2 = o1 + divx;
      const int s0 = (bx0 + o0) ? hist.get(bx0 + o0 - 1) : 0, e0 = (by0 <= by1) ? hist.get(bx1 + o0) : s0;
      const int s1 = h1 ? hist.get(bx0 + o1 - 1) : 0, e1 = h1 ? hist.get(bx1 + o1) : 0;
      const int s2 = h2 ? hist.get(bx0 + o2 - 1) : 0, e2 = h2 ? hist.get(bx1 + o2) : 0;
      const int n0 = e0 - s0, n01 = n0 + (e1 - s1), total = n01 + (e2 - s2);
      for (int it = sl; it < total; it += LPC) take(it < n0 ? s0 + it : (it < n01 ? s1 + (it - n0) : s2 + (it - n01)));
    } else
#endif
    for (int by = by0; by <= by1; ++by) {
      const int b_lo = bx0 + by * divx, b_hi = bx1 + by * divx;
      const int s = b_lo ? hist.get(b_lo - 1) : 0, e = hist.get(b_hi);
      for (int a = s + sl; a < e; a += LPC) take(a);
    }
#pragma unroll
    for (int d = 1; d < LPC; d <<= 1) {
      gN += __shfl_xor_sync(gmask, gN, d);
      S0 += __shfl_xor_sync(gmask, S0, d); S1x += __shfl_xor_sync(gmask, S1x, d); S1y += __shfl_xor_sync(gmask, S1y, d);
      Sxx += __shfl_xor_sync(gmask, Sxx, d); Sxy += __shfl_xor_sync(gmask, Sxy, d); Syy += __shfl_xor_sync(gmask, Syy, d);
    }
    if (sl == 0) {                                          // raw moments; finalised one centroid per thread below
      double2* o = tmp + (size_t)c * 4;
      o[0] = make_double2(S0, S1x); o[1] = make_double2(S1y, Sxx); o[2] = make_double2(Sxy, Syy);
      o[3] = make_double2(__longlong_as_double((long long)gN), 0.0);
    }
  }
  __syncthreads();
  K3P(9)
  int ncells = 0;                                          // block-uniform running count
  for (int c0 = 0; c0 < nvox; c0 += T) {
    const int c = c0 + tid;
    bool valid = false;
    double ux = 0, uy = 0, nx_ = 0, ny_ = 0, cxx = 0, cxy_ = 0, cyy = 0, scale = 0, avgI = 0;
    int gN = 0;
    if (c < nvox) {
      const double2* o = tmp + (size_t)c * 4;
      const double2 m0 = o[0], m1 = o[1], m2 = o[2];
      gN = (int)__double_as_longlong(o[3].x);
      if (gN >= 6) {                                        // pointnormal.cpp:291
        const float2 q = cxy[c];
        const double S0 = m0.x, iS0 = 1.0 / S0;             // one reciprocal for the five quotients (FP64-issue bound phase)
        const double mdx = m0.y * iS0, mdy = m1.x * iS0;    // weighted mean relative to q
        ux = (double)q.x + mdx; uy = (double)q.y + mdy;
        cxx = m1.y * iS0 - mdx * mdx; cxy_ = m2.x * iS0 - mdx * mdy; cyy = m2.y * iS0 - mdy * mdy;
        const Eig2 eg = eig2_sym(cxx, cxy_, cyy);           // ComputeNormal (:37-63)
        const double cond = fabs(eg.lmax / eg.lmin);
        const double det = eg.lmax * eg.lmin;
        valid = (cond <= 10000) && (det > 0.00001) && eg.lmin > 0 && eg.lmax > 0;
        nx_ = eg.nx; ny_ = eg.ny;
        if (nx_ * (p.origin_x - ux) + ny_ * (p.origin_y - uy) < 0) { nx_ = -nx_; ny_ = -ny_; }
        scale = log(1.0 + cond / 2);                        // scale_
        avgI = S0 / (double)gN;                             // avg_intensity_
      }
    }
    int total;
    const int pos = ncells + block_excl_scan(valid ? 1 : 0, s_warp, &total);   // (the scan's barrier orders the cxy reads above before the writes below)
    if (valid && pos < p.pool.max_cells) {
      p.pool.mean[cbase + pos] = make_double2(ux, uy);
      p.pool.normal[cbase + pos] = make_double2(nx_, ny_);
      p.pool.cov[cbase + pos] = make_double4(cxx, cxy_, cxy_, cyy);
      p.pool.planarity[cbase + pos] = scale;
      p.pool.avg_intensity[cbase + pos] = avgI;
      p.pool.nsamples[cbase + pos] = gN;
      cxy[pos] = make_float2((float)ux, (float)uy);        // pointnormal.cpp:153-157 (pos <= c: safe in place)
    }
    ncells += total;
    __syncthreads();
  }
  K3P(10)
  ncells = min(ncells, p.pool.max_cells);
  if (tid == 0) p.pool.ncells[slot] = ncells;
  __syncthreads();

  // ---- NN index over the fp32 means --------------------------------------------------------------
  Hist16 ghist; ghist.w = reinterpret_cast<uint32_t*>(s_hist);
  build_nn_grid(p.pool, slot, cxy, ncells, p.nn_cell, ghist, min(K3_HIST_CAP, p.pool.grid_cap) - 1, s_warp, s_red);
#ifdef CFEAR_K3_PROFILE
  K3P(11)
  if (tid == 0 && (scan & 63) == 5)
    printf("K3 scan %d n=%d nvox=%d ncells=%d nbins=%d | stage+compact %lld bounds %lld hist %lld scan %lld scatter %lld reorder %lld voxlist %lld centroids %lld moments %lld finalise %lld nngrid %lld | total %lld\n",
           scan, n, nvox, ncells, nbins, k3t[1] - k3t[0], k3t[2] - k3t[1], k3t[3] - k3t[2], k3t[4] - k3t[3], k3t[5] - k3t[4], k3t[6] - k3t[5],
           k3t[7] - k3t[6], k3t[8] - k3t[7], k3t[9] - k3t[8], k3t[10] - k3t[9], k3t[11] - k3t[10], k3t[11] - k3t[0]);
#endif
  };
  if (hist_global) rest(std::true_type{}); else rest(std::false_type{});
}

// NN index for an uploaded cell set (cfear_cells_upload): one CTA per slot.
struct K4Params { CellPool pool; const int32_t* slots; float nn_cell; };

__global__ void __launch_bounds__(K3_THREADS, CFEAR_K3_MINBLOCKS) k4_build_index(const K4Params p) {
  extern __shared__ __align__(128) unsigned char dyn_smem[];
  __shared__ int s_warp[33];
  __shared__ __align__(16) float s_red[128];
  const int slot = p.slots[blockIdx.x];
  const int n = p.pool.ncells[slot];
  Hist16 s_hist; s_hist.w = reinterpret_cast<uint32_t*>(dyn_smem);
  float2* fm = p.pool.fm_scratch + (size_t)slot * p.pool.max_cells;
  const double2* mean = p.pool.mean + (size_t)slot * p.pool.max_cells;
  for (int i = threadIdx.x; i < n; i += blockDim.x) fm[i] = make_float2((float)mean[i].x, (float)mean[i].y);
  __syncthreads();
  build_nn_grid(p.pool, slot, fm, n, p.nn_cell, s_hist, min(K3_HIST_CAP, p.pool.grid_cap) - 1, s_warp, s_red);
}

}  // namespace cfear
