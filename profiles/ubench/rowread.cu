// How fast can 256 x 400 rows of 3360 B be streamed from HBM?  Upper bounds for K1's access pattern.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint4 ldnc(const uint8_t* p) { uint4 r; asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p)); return r; }
// (a) K1's layout: warp per row, 7 x 16 B per lane in flight, 8 rows per CTA
template <int MINB>
__global__ void __launch_bounds__(256, MINB) warp_per_row(const uint8_t* img, int nrows, int R, uint32_t* out) {
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= nrows) return;
  const uint8_t* base = img + (size_t)row * R;
  const int nvec = R >> 4;
  uint4 d[7];
#pragma unroll
  for (int i = 0; i < 7; ++i) { const int v = i * 32 + lane; d[i] = v < nvec ? ldnc(base + 16 * v) : make_uint4(0, 0, 0, 0); }
  uint32_t s = 0;
#pragma unroll
  for (int i = 0; i < 7; ++i) s += d[i].x ^ d[i].y ^ d[i].z ^ d[i].w;
  s = __reduce_add_sync(0xffffffffu, s);
  if (lane == 0) out[row] = s;
}
// (b) flat grid-stride, 4 x 16 B per thread in flight
__global__ void __launch_bounds__(256) flat(const uint4* img, size_t nvec, uint32_t* out) {
  uint32_t s = 0;
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (size_t)gridDim.x * blockDim.x;
  for (; i + 3 * stride < nvec; i += 4 * stride) {
    uint4 a = ldnc((const uint8_t*)(img + i)), b = ldnc((const uint8_t*)(img + i + stride)), c = ldnc((const uint8_t*)(img + i + 2 * stride)), d = ldnc((const uint8_t*)(img + i + 3 * stride));
    s += a.x ^ a.y ^ a.z ^ a.w ^ b.x ^ b.y ^ b.z ^ b.w ^ c.x ^ c.y ^ c.z ^ c.w ^ d.x ^ d.y ^ d.z ^ d.w;
  }
  for (; i < nvec; i += stride) { uint4 a = img[i]; s += a.x ^ a.y ^ a.z ^ a.w; }
  if (s == 0x12345678u) out[0] = s;
}
// (c) persistent CTAs, TMA bulk copies of 8-row groups into a 4-stage shared-memory ring, warps consume from smem
__device__ __forceinline__ uint32_t sa(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
template <int STAGES, int ROWS>
__global__ void __launch_bounds__(256) tma_ring(const uint8_t* img, int ngroups, int R, uint32_t* out) {
  extern __shared__ __align__(128) uint8_t sm[];
  __shared__ __align__(8) uint64_t full[STAGES];
  const uint32_t gbytes = (uint32_t)ROWS * R;
  if (threadIdx.x == 0) { for (int s = 0; s < STAGES; ++s) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(sa(&full[s]))); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  __syncthreads();
  const int first = blockIdx.x, step = gridDim.x;
  auto issue = [&](int it) {
    const int g = first + it * step;
    if (g < ngroups) {
      const int s = it % STAGES;
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(sa(&full[s])), "r"(gbytes) : "memory");
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(sa(sm + (size_t)s * gbytes)), "l"(img + (size_t)g * gbytes), "r"(gbytes), "r"(sa(&full[s])) : "memory");
    }
  };
  if (threadIdx.x == 0) for (int it = 0; it < STAGES - 1; ++it) issue(it);
  uint32_t acc = 0;
  for (int it = 0;; ++it) {
    const int g = first + it * step;
    if (g >= ngroups) break;
    if (threadIdx.x == 0) issue(it + STAGES - 1);
    const int s = it % STAGES; const uint32_t ph = (it / STAGES) & 1;
    asm volatile("{\n.reg .pred P1;\nW: mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n@P1 bra D;\nbra W;\nD:\n}" ::"r"(sa(&full[s])), "r"(ph) : "memory");
    const uint4* p = reinterpret_cast<const uint4*>(sm + (size_t)s * gbytes);
    for (uint32_t v = threadIdx.x; v < gbytes / 16; v += 256) { uint4 a = p[v]; acc += a.x ^ a.y ^ a.z ^ a.w; }
    __syncthreads();
  }
  if (acc == 0x12345678u) out[0] = acc;
}
int main() {
  const int nscan = 256, A = 400, R = 3360; const size_t bytes = (size_t)nscan * A * R;
  uint8_t* img; uint32_t* out; cudaMalloc(&img, bytes); cudaMalloc(&out, nscan * A * 4); cudaMemset(img, 30, bytes);
  uint8_t* flush; cudaMalloc(&flush, 256 << 20);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  auto run = [&](const char* name, auto f) {
    float best = 1e9f;
    for (int rep = 0; rep < 6; ++rep) { cudaMemsetAsync(flush, rep, 256 << 20); cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1); if (rep) best = fminf(best, ms); }
    printf("%-44s %.4f ms  %.0f GB/s  (%s)\n", name, best, bytes / best / 1e6, cudaGetErrorString(cudaGetLastError()));
  };
  const int nrows = nscan * A;
  run("warp per row, 7x16B/lane, 4 CTA/SM", [&] { warp_per_row<4><<<(nrows + 7) / 8, 256>>>(img, nrows, R, out); });
  run("warp per row, 7x16B/lane, 6 CTA/SM", [&] { warp_per_row<6><<<(nrows + 7) / 8, 256>>>(img, nrows, R, out); });
  run("warp per row, 7x16B/lane, 8 CTA/SM", [&] { warp_per_row<8><<<(nrows + 7) / 8, 256>>>(img, nrows, R, out); });
  run("flat grid-stride 148x8 CTAs, 4x16B/thread", [&] { flat<<<148 * 8, 256>>>((const uint4*)img, bytes / 16, out); });
  cudaFuncSetAttribute(tma_ring<4, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * 8 * R);
  run("TMA ring 4 stages x 8 rows, 2 CTA/SM", [&] { tma_ring<4, 8><<<148 * 2, 256, 4 * 8 * R>>>(img, nrows / 8, R, out); });
  cudaFuncSetAttribute(tma_ring<3, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 3 * 16 * R);
  run("TMA ring 3 stages x 16 rows, 1 CTA/SM", [&] { tma_ring<3, 16><<<148, 256, 3 * 16 * R>>>(img, nrows / 16, R, out); });
  return 0;
}
