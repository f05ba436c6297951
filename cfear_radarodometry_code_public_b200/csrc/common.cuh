// Shared device helpers and device-side data layout for the CFEAR hot path (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace cfear {

constexpr unsigned FULL = 0xffffffffu;

// Kernel launch with a per-launch scheduling priority (cudaLaunchAttributePriority; numerically lower = served first by
// the block scheduler, 0 = the default).  With several steps in flight the registration kernel -- the long, latency-bound
// end of every step's chain -- is given precedence over the next steps' filter kernels, which fill what it leaves.
template <typename... KArgs, typename... Args>
inline cudaError_t launch_with_priority(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, int prio, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributePriority; at[0].val.priority = prio;
  cfg.attrs = at; cfg.numAttrs = prio != 0 ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// ------------------------------------------------------------------------------------------------
// Device-resident cell-set pool (SoA).  Slot s occupies [s*max_cells, (s+1)*max_cells) of each array.
// This is what MapPointNormal holds for the pose path (pointnormal.h:66-73,196-199): cells + the
// 2-D nearest-neighbour index over fp32 cell means (kd_cells), here a uniform bucket grid.
// ------------------------------------------------------------------------------------------------
struct NNGrid {            // per slot
  float ox, oy, inv_g, g;  // origin, 1/cell, cell
  int nx, ny;              // dims (nx*ny <= grid_cap)
};

struct CellPool {
  int max_cells;
  int grid_cap;            // max grid cells per slot
  int* ncells;             // [slots]
  double2* mean;           // u_
  double2* normal;         // snormal_
  double4* cov;            // cov_ row-major (xx, xy, yx, yy)
  double* planarity;       // scale_
  double* avg_intensity;
  int* nsamples;
  // NN index: uniform bucket grid over the fp32 means
  NNGrid* grid;            // [slots]
  uint16_t* gstart;        // [slots][grid_stride]  bucket start offsets (nx*ny+1 used); u16: max_cells <= 65535
  float4* gpt;             // [slots][max_cells]    (x, y, cell index as int bits, normal as 2 x fp16) sorted by bucket
  float2* fm_scratch;      // [slots][max_cells]    fp32 means staging for the index build of uploaded sets
  int grid_stride;         // u16 entries per slot (grid_cap + 8: keeps every slot 16-byte aligned for bulk copies)
};

// ------------------------------------------------------------------------------------------------
// warp / block primitives
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }
__device__ __forceinline__ int warp_id() { return threadIdx.x >> 5; }

__device__ __forceinline__ int warp_incl_scan(int v) {
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    int t = __shfl_up_sync(FULL, v, d);
    if (lane_id() >= d) v += t;
  }
  return v;
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(FULL, v, d);
  return v;
}

// sin / cos of small angles (Compensate: time stamp in [-0.5, 0.5] times the inter-scan yaw; K5: the yaw component of an
// LM step): Taylor polynomials, exact to an ulp below 1/16 rad, the library routine above.
__device__ __forceinline__ void sincos_small(double x, double* s, double* c) {
  if (fabs(x) < 0.0625) {
    const double z = x * x;
    double ps = fma(z, -1.0 / 39916800.0, 1.0 / 362880.0);
    ps = fma(z, ps, -1.0 / 5040.0); ps = fma(z, ps, 1.0 / 120.0); ps = fma(z, ps, -1.0 / 6.0);
    *s = fma(x * z, ps, x);
    double pc = fma(z, 1.0 / 479001600.0, -1.0 / 3628800.0);
    pc = fma(z, pc, 1.0 / 40320.0); pc = fma(z, pc, -1.0 / 720.0); pc = fma(z, pc, 1.0 / 24.0); pc = fma(z, pc, -0.5);
    *c = fma(z, pc, 1.0);
  } else {
    sincos(x, s, c);
  }
}

// Block-wide exclusive scan of one int per thread.  s_warp: >= 33 ints of shared memory.
// Returns the exclusive prefix; *total gets the block sum.  Contains __syncthreads().
__device__ __forceinline__ int block_excl_scan(int v, int* s_warp, int* total) {
  const int incl = warp_incl_scan(v);
  const int nw = (blockDim.x + 31) >> 5;
  __syncthreads();                         // protect s_warp reuse
  if (lane_id() == 31) s_warp[warp_id()] = incl;
  __syncthreads();
  if (warp_id() == 0) {
    int w = lane_id() < nw ? s_warp[lane_id()] : 0;
    int wi = warp_incl_scan(w);
    s_warp[lane_id()] = wi - w;            // exclusive warp offsets
    if (lane_id() == 31) s_warp[32] = wi;  // total
  }
  __syncthreads();
  *total = s_warp[32];
  return s_warp[warp_id()] + incl - v;
}

// In-place exclusive scan of an int array of length n (global or shared) by the whole block.
// After the call a[i] = sum_{j<i} old a[j]; returns the total.  a must have room for n entries.
__device__ inline int block_array_excl_scan(int* a, int n, int* s_warp) {
  const int T = blockDim.x;
  const int chunk = (n + T - 1) / T;
  const int lo = min(threadIdx.x * chunk, n), hi = min(lo + chunk, n);
  int s = 0;
  for (int i = lo; i < hi; ++i) s += a[i];
  int total;
  int base = block_excl_scan(s, s_warp, &total);
  for (int i = lo; i < hi; ++i) { int t = a[i]; a[i] = base; base += t; }
  __syncthreads();
  return total;
}

}  // namespace cfear
