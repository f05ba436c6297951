// libcfear_b200.so -- C ABI (include/cfear_b200.h) over the sm_100a kernels of the CFEAR per-scan hot path.
// No CPU fallback: every entry point either runs the CUDA kernels or fails with CFEAR_ERR_CUDA /
// CFEAR_ERR_NO_DEVICE.  Nothing here touches oracle/.
#include "../../include/cfear_b200.h"

#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <new>
#include <string>
#include <vector>

#include "common.cuh"
#include "k1_kstrongest.cuh"
#include "k3_surface.cuh"
#include "k5_launch.cuh"
#include "k6_fuser.cuh"
#include "k7_cfar.cuh"

using namespace cfear;

static thread_local std::string g_err;

#define CK(expr)                                                                                  \
  do {                                                                                            \
    cudaError_t e__ = (expr);                                                                     \
    if (e__ != cudaSuccess) {                                                                     \
      g_err = std::string(#expr) + ": " + cudaGetErrorString(e__);                                \
      return CFEAR_ERR_CUDA;                                                                      \
    }                                                                                             \
  } while (0)

static_assert(sizeof(cfear_reg_stats) == sizeof(RegStatsDev), "stats layout");
static_assert(sizeof(cfear_point) == sizeof(float4), "point layout");

namespace cfear {

// Row-padded points [nscans][A][k] -> dense cloud [nscans][cap] in row order (radar_filters.cpp:316-336 push_back order).
__global__ void __launch_bounds__(512) k2_compact_rows(const float4* rowpts, const int32_t* rowcnt, int A, int k, int cap,
                                                       float4* out, int32_t* nout) {
  __shared__ int s_warp[33];
  extern __shared__ int s_off[];
  const int scan = blockIdx.x;
  const int32_t* rc = rowcnt + (size_t)scan * A;
  for (int a = threadIdx.x; a < A; a += blockDim.x) s_off[a] = rc[a];
  __syncthreads();
  const int n = block_array_excl_scan(s_off, A, s_warp);
  const float4* src = rowpts + (size_t)scan * A * k;
  float4* dst = out + (size_t)scan * cap;
  for (int s = threadIdx.x; s < A * k; s += blockDim.x) {
    const int a = s / k, j = s - a * k;
    const int start = s_off[a];
    const int cnt = ((a + 1 < A) ? s_off[a + 1] : n) - start;
    if (j < cnt) dst[start + j] = src[s];
  }
  if (threadIdx.x == 0) nout[scan] = n;
}

// Compensate (utils.cpp:96-113) on a dense cloud.
__global__ void k2b_compensate(float4* cloud, int n, double m0, double m1, double m2, int ccw) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float4 pt = cloud[i];
  const double x = (double)pt.x, y = (double)pt.y;
  const double d = rel_time_stamp(x, y, ccw != 0);
  double s1, c1; sincos_small(d * m2, &s1, &c1);
  const double tx = c1 * x + (-s1) * y + d * m0;
  const double ty = s1 * x + c1 * y + d * m1;
  pt.x = (float)tx; pt.y = (float)ty;
  cloud[i] = pt;
}

__global__ void k4b_nearest(CellPool pool, int slot, const double* q, int nq, double radius, int32_t* out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nq) return;
  const NNGrid G = pool.grid[slot];
  GridView V; V.gs = pool.gstart + (size_t)slot * pool.grid_stride; V.gp = pool.gpt + (size_t)slot * pool.max_cells;
  uint32_t n16; out[i] = nn_query(V, G, q[2 * i], q[2 * i + 1], nn_radius(radius), &n16);
}

// AoS cfear_cell <-> pool SoA
struct CellAoS { double mean[2], normal[2], cov[4], planarity, avg_intensity; int32_t nsamples, pad; };
__global__ void k_cells_gather(CellPool pool, int slot, CellAoS* out, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const size_t b = (size_t)slot * pool.max_cells + i;
  CellAoS c;
  c.mean[0] = pool.mean[b].x; c.mean[1] = pool.mean[b].y;
  c.normal[0] = pool.normal[b].x; c.normal[1] = pool.normal[b].y;
  const double4 v = pool.cov[b];
  c.cov[0] = v.x; c.cov[1] = v.y; c.cov[2] = v.z; c.cov[3] = v.w;
  c.planarity = pool.planarity[b]; c.avg_intensity = pool.avg_intensity[b];
  c.nsamples = pool.nsamples[b]; c.pad = 0;
  out[i] = c;
}
__global__ void k_cells_scatter(CellPool pool, int slot, const CellAoS* in, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i == 0) pool.ncells[slot] = n;
  if (i >= n) return;
  const size_t b = (size_t)slot * pool.max_cells + i;
  const CellAoS c = in[i];
  pool.mean[b] = make_double2(c.mean[0], c.mean[1]);
  pool.normal[b] = make_double2(c.normal[0], c.normal[1]);
  pool.cov[b] = make_double4(c.cov[0], c.cov[1], c.cov[2], c.cov[3]);
  pool.planarity[b] = c.planarity; pool.avg_intensity[b] = c.avg_intensity;
  pool.nsamples[b] = c.nsamples;
}

}  // namespace cfear

static_assert(sizeof(cfear_cell) == sizeof(CellAoS), "cell layout");

#ifndef CFEAR_NN_CELL
#define CFEAR_NN_CELL 6.0f             // bucket size (metres) of the nearest-neighbour grid over a cell set's means: the finest whose four keyframe grids
                                       // still fit K5's shared memory at three CTAs per SM (4 m is 3 % faster per CTA but needs two CTAs per SM;
                                       // profiles/r02m_nn_cell_ab.txt, profiles/r02v_k5_three_per_sm_ab.txt)
#endif
constexpr int CFEAR_MAX_TICKETS = 8;   // steps that may be in flight between submit and wait
constexpr int CFEAR_NPIPES = 8;        // most device-resident steps that may overlap (cfear_config.steps_in_flight, default 5)

// Everything one step writes between K1 and K5.  Set 0 aliases the context's own buffers (every stream-ordered entry
// point uses it on the context stream); sets 1..CFEAR_NPIPES have their own streams so that consecutive device-resident
// steps can overlap: K1 / K3 of step i+1 fill the SMs that K5 of step i leaves idle in its tail.
struct PipeBufs {
  cudaStream_t stream = nullptr;
  cudaEvent_t in = nullptr, done = nullptr;
  int32_t *d_kidx = nullptr, *d_kcnt = nullptr, *d_rowcnt = nullptr, *d_npts = nullptr, *d_status = nullptr, *d_slots = nullptr;
  float4 *d_rowcloud = nullptr, *d_bufA = nullptr, *d_bufB = nullptr;
  int* d_ghist = nullptr; double2 *d_celltmp = nullptr, *d_res = nullptr;
};

struct cfear_ctx {
  cfear_config cfg;
  cudaStream_t stream = nullptr, copy_stream = nullptr;
  std::vector<cudaEvent_t> ev;      // 4 per timed step: before K1, after K1, after K3, after K5
  int ev_used = 0;                  // timed steps recorded since the last cfear_stage_timing call
  cudaEvent_t* evset = nullptr;     // current step's 4 events
  std::vector<cudaEvent_t> chunk_ev, k1_done, ticket_ev;   // host-buffer path: H2D done / staging area read / step done
  cudaEvent_t polar_free = nullptr; bool polar_dirty = true; int next_ticket = 0;
  int64_t launches = 0;
  int prio[3] = {0, 0, 0};          // launch priorities of K1, K3, K5 (experiment: CFEAR_PRIO="k1,k3,k5")
  int timing = 0;
  float stage_ms[3] = {0, 0, 0};
  int cap_pts = 0, max_cells = 0, grid_cap = 0, res_cap = 0;
  int pts_in_smem = 0; size_t k3_smem = 0, k4_smem = 0; int k5_smem = 0, k5_smem_mid = 0, k5_smem_wide = 0, num_sms = 0, k3_wide = 0;
  int g_hist_cap = 0;
  std::vector<void*> allocs;
  // device buffers
  uint8_t* d_polar = nullptr; double2* d_cs = nullptr;
  int32_t *d_kidx = nullptr, *d_kcnt = nullptr, *d_rowcnt = nullptr, *d_rowpeakcnt = nullptr, *d_npts = nullptr, *d_npeaks = nullptr;
  float4 *d_rowcloud = nullptr, *d_rowpeaks = nullptr, *d_cloud = nullptr, *d_peaks = nullptr, *d_bufA = nullptr, *d_bufB = nullptr;
  int* d_ghist = nullptr; int32_t* d_status = nullptr; double2* d_celltmp = nullptr;
  double* d_mot = nullptr; int32_t *d_slots = nullptr, *d_curslots = nullptr, *d_kfslots = nullptr;
  double *d_poses = nullptr, *d_cov36 = nullptr; cfear_reg_stats* d_stats = nullptr; int32_t* d_assoc = nullptr;
  double *d_assoc_sim = nullptr, *d_softL = nullptr;
  double2* d_res = nullptr; double* d_queries = nullptr; int32_t* d_qout = nullptr; CellAoS* d_cellaos = nullptr;
  CellPool pool;
  std::vector<double2> h_cs;
  PipeBufs pb[CFEAR_NPIPES + 1];    // [0] aliases the buffers above on the context stream; [1..] the overlapped steps' sets
  bool pipes_ready = false, inflight = false;
  int npipes = 4, next_pipe = 0, last_pipe = 0;
  int launch_conc = 1;              // steps this launch shares the GPU with (cfear_odometry_step_batch_dev_submit: steps_in_flight)
  int pipe_user = 0;                // who enqueued on the internal streams last: 1 = overlapped batch steps, 2 = sequence replay

  template <typename T> int alloc(T** p, size_t count) {
    void* q = nullptr;
    cudaError_t e = cudaMalloc(&q, std::max<size_t>(count, 1) * sizeof(T));
    if (e != cudaSuccess) { g_err = std::string("cudaMalloc: ") + cudaGetErrorString(e); return CFEAR_ERR_CUDA; }
    allocs.push_back(q);
    // every buffer starts zeroed: the padded tails of row-padded outputs (k-strongest rows with fewer than k candidates,
    // cell arrays beyond ncells) are copied and staged as whole blocks, and should not carry another context's bytes
    // (allocation happens at creation and on the first use of a feature; the context's streams are non-blocking, i.e. not
    // ordered against the legacy stream the memset runs on, hence the synchronisation)
    e = cudaMemset(q, 0, std::max<size_t>(count, 1) * sizeof(T));
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { g_err = std::string("cudaMemset: ") + cudaGetErrorString(e); return CFEAR_ERR_CUDA; }
    *p = reinterpret_cast<T*>(q);
    return CFEAR_OK;
  }
};

#define AL(p, n)                                    \
  do {                                              \
    int rc__ = c->alloc(&(p), (size_t)(n));         \
    if (rc__ != CFEAR_OK) { cfear_destroy(c); return rc__; } \
  } while (0)

extern "C" {

const char* cfear_last_error(void) { return g_err.c_str(); }
const char* cfear_version(void) { return "cfear_b200 0.1 (sm_100a)"; }

void cfear_default_config(cfear_config* cfg) {
  memset(cfg, 0, sizeof(*cfg));
  cfg->device = 0; cfg->max_batch = 1; cfg->azimuths = 400; cfg->range_bins = 3360; cfg->k_strongest = 12;
  cfg->z_min = 60.f; cfg->range_res = 0.0438f; cfg->min_distance = 2.5f; cfg->radius = 3.5f;   // radar_driver.h:40-48
  cfg->downsample_factor = 1.0; cfg->weight_intensity = 1; cfg->compensate = 1; cfg->radar_ccw = 0;
  cfg->cost = CFEAR_COST_P2L; cfg->loss = CFEAR_LOSS_HUBER; cfg->weight_opt = CFEAR_WEIGHT_UNIFORM;
  cfg->solver_mode = CFEAR_SOLVER_CERES_LM; cfg->loss_limit = 0.1; cfg->cov_scale = 1.0; cfg->regularization = 1.0;
  cfg->reg_radius = 2.0; cfg->max_outer = 8; cfg->min_outer = 3; cfg->max_inner = 20; cfg->gn_iters = 10;
  cfg->max_keyframes = 4; cfg->max_cellsets = 8; cfg->max_cells = 0; cfg->steps_in_flight = 0;
}

void cfear_destroy(cfear_ctx* c) {
  if (!c) return;
  cudaSetDevice(c->cfg.device);
  if (c->stream) cudaStreamSynchronize(c->stream);
  for (int i = 1; i <= CFEAR_NPIPES; ++i) if (c->pb[i].stream) cudaStreamSynchronize(c->pb[i].stream);
  for (void* p : c->allocs) cudaFree(p);
  for (auto& e : c->ev) if (e) cudaEventDestroy(e);
  for (auto& e : c->chunk_ev) cudaEventDestroy(e);
  for (auto& e : c->k1_done) cudaEventDestroy(e);
  for (auto& e : c->ticket_ev) cudaEventDestroy(e);
  if (c->polar_free) cudaEventDestroy(c->polar_free);
  for (int i = 1; i <= CFEAR_NPIPES; ++i) {
    PipeBufs& B = c->pb[i];
    if (B.stream) cudaStreamDestroy(B.stream);
    if (B.in) cudaEventDestroy(B.in);
    if (B.done) cudaEventDestroy(B.done);
  }
  if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
  if (c->stream) cudaStreamDestroy(c->stream);
  delete c;
}

int cfear_create(const cfear_config* cfg, cfear_ctx** out) {
  if (!cfg || !out) { g_err = "null argument"; return CFEAR_ERR_ARG; }
  *out = nullptr;
  if (cfg->azimuths < 1 || cfg->range_bins < 1 || cfg->k_strongest < 1 || cfg->k_strongest > K1_MAXK ||
      cfg->max_batch < 1 || cfg->max_keyframes < 1 || cfg->max_keyframes + 1 > K5_MAXSCANS || cfg->max_cellsets < 1 ||
      cfg->range_bins > 65535 || !(cfg->radius > 0.f) || !(cfg->downsample_factor > 0.0) || cfg->max_cells > 65535 ||
      (cfg->max_cells <= 0 && (long long)cfg->azimuths * cfg->k_strongest > 65535)) {
    g_err = "invalid configuration (need 1<=k<=64, range_bins<=65535, max_keyframes<=64, radius>0, max_cells<=65535)";
    return CFEAR_ERR_ARG;
  }
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= cfg->device) {
    g_err = "no CUDA device: the cfear_b200 path has no CPU fallback";
    return CFEAR_ERR_NO_DEVICE;
  }
  cfear_ctx* c = new (std::nothrow) cfear_ctx();
  if (!c) { g_err = "out of host memory"; return CFEAR_ERR_ARG; }
  c->cfg = *cfg;
  if (const char* e = getenv("CFEAR_PRIO")) sscanf(e, "%d,%d,%d", &c->prio[0], &c->prio[1], &c->prio[2]);
  c->npipes = cfg->steps_in_flight > 0 ? std::min(cfg->steps_in_flight, CFEAR_NPIPES) : 5;
  // from here on a CUDA failure must not leak the half-built context
#define CKC(expr)                                                                                 \
  do {                                                                                            \
    cudaError_t e__ = (expr);                                                                     \
    if (e__ != cudaSuccess) {                                                                     \
      g_err = std::string(#expr) + ": " + cudaGetErrorString(e__);                                \
      cfear_destroy(c);                                                                           \
      return CFEAR_ERR_CUDA;                                                                      \
    }                                                                                             \
  } while (0)
  CKC(cudaSetDevice(cfg->device));
  CKC(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
  CKC(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
  CKC(cudaEventCreateWithFlags(&c->polar_free, cudaEventDisableTiming));
  const int A = cfg->azimuths, R = cfg->range_bins, k = cfg->k_strongest, B = cfg->max_batch;
  c->cap_pts = A * k;
  c->max_cells = cfg->max_cells > 0 ? cfg->max_cells : A * k;
  c->grid_cap = K3_HIST_CAP;
  c->res_cap = cfg->max_keyframes * c->max_cells;
  c->g_hist_cap = 1 << 18;
  // shared-memory plan of K3
  int max_optin = 0;
  CKC(cudaDeviceGetAttribute(&max_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, cfg->device));
  // 16-bit histogram entries, two per word (k3_surface.cuh, Hist16); the same area first holds the A row counts as ints
  const size_t hist_bytes = std::max((size_t)((K3_HIST_CAP + 2) / 2) * 4, (size_t)(A + 1) * sizeof(int));
  const size_t full = (size_t)c->cap_pts * 16 + hist_bytes;
  c->pts_in_smem = (full + 2048 <= (size_t)max_optin) ? 1 : 0;
  c->k3_smem = c->pts_in_smem ? full : hist_bytes;
  c->k4_smem = hist_bytes;
  if (c->pts_in_smem) {
    CKC(cudaFuncSetAttribute(k3_surface_points<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c->k3_smem));
    CKC(cudaFuncSetAttribute(k3_surface_points<true, K3_THREADS_WIDE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c->k3_smem));
    c->k3_wide = 1;
  }
  else CKC(cudaFuncSetAttribute(k3_surface_points<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c->k3_smem));
  CKC(cudaFuncSetAttribute(k4_build_index, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)hist_bytes));
  // (the CA-CFAR kernel's shared-memory attribute is set by cfear_cfar_filter, which is the only place that needs it)
  {   // K5: K5_MINBLOCKS CTAs per SM, each with 1 KB reserved by the system and ~1 KB of static shared memory
    int per_sm = 0;
    CKC(cudaDeviceGetAttribute(&per_sm, cudaDevAttrMaxSharedMemoryPerMultiprocessor, cfg->device));
    const int fit = per_sm / K5_MINBLOCKS - 2048;
    c->k5_smem = std::min(K5_SMEM_BYTES > 0 ? K5_SMEM_BYTES : fit, max_optin - 1024);
    // the wide form (one 384-thread CTA per SM, batches of at most one problem per SM) takes what one CTA may have
    CKC(cudaDeviceGetAttribute(&c->num_sms, cudaDevAttrMultiProcessorCount, cfg->device));
    c->k5_smem_wide = max_optin - 2048;
    c->k5_smem_mid = std::min(per_sm / 2 - 2048, max_optin - 2048);     // two 192-thread CTAs per SM
  }
  CKC(k5_set_smem_cost0(c->k5_smem, c->k5_smem_mid, c->k5_smem_wide)); CKC(k5_set_smem_cost1(c->k5_smem, c->k5_smem_mid, c->k5_smem_wide));
  CKC(k5_set_smem_cost2(c->k5_smem, c->k5_smem_mid, c->k5_smem_wide));

  const size_t rows = (size_t)B * A;
  AL(c->d_polar, rows * R);
  AL(c->d_cs, A);
  AL(c->d_kidx, rows * k); AL(c->d_kcnt, rows); AL(c->d_rowcnt, rows); AL(c->d_rowpeakcnt, rows);
  AL(c->d_rowcloud, rows * k); AL(c->d_rowpeaks, rows * k);
  AL(c->d_cloud, (size_t)B * c->cap_pts); AL(c->d_peaks, (size_t)B * c->cap_pts);
  AL(c->d_npts, B); AL(c->d_npeaks, B); AL(c->d_status, B);
  if (!c->pts_in_smem) AL(c->d_bufA, (size_t)B * c->cap_pts);
  AL(c->d_bufB, (size_t)B * c->cap_pts);
  AL(c->d_ghist, (size_t)B * (c->g_hist_cap + 1));
  AL(c->d_celltmp, (size_t)B * c->cap_pts * 4);
  AL(c->d_mot, (size_t)B * 3);
  const int ns = cfg->max_keyframes + 1;
  AL(c->d_slots, (size_t)B * ns); AL(c->d_curslots, B); AL(c->d_kfslots, (size_t)B * ns);
  AL(c->d_poses, (size_t)B * ns * 3); AL(c->d_cov36, (size_t)B * 36); AL(c->d_stats, B);
  AL(c->d_res, (size_t)B * c->res_cap * 4);
  AL(c->d_cellaos, c->max_cells);
  // cell pool
  CellPool& P = c->pool;
  const size_t S = (size_t)cfg->max_cellsets, M = (size_t)c->max_cells;
  P.max_cells = c->max_cells; P.grid_cap = c->grid_cap;
  AL(P.ncells, S); AL(P.mean, S * M); AL(P.normal, S * M); AL(P.cov, S * M); AL(P.planarity, S * M);
  AL(P.avg_intensity, S * M); AL(P.nsamples, S * M); AL(P.grid, S);
  P.grid_stride = P.grid_cap + 8;
  AL(P.gstart, S * P.grid_stride); AL(P.gpt, S * M); AL(P.fm_scratch, S * M);
  CKC(cudaMemsetAsync(P.ncells, 0, S * sizeof(int), c->stream));
  CKC(cudaMemsetAsync(P.gstart, 0, S * P.grid_stride * sizeof(uint16_t), c->stream));
  {
    std::vector<NNGrid> g(S);
    for (auto& x : g) { x.ox = x.oy = 0.f; x.g = 4.f; x.inv_g = 0.25f; x.nx = x.ny = 1; }
    CKC(cudaMemcpyAsync(P.grid, g.data(), S * sizeof(NNGrid), cudaMemcpyHostToDevice, c->stream));
    CKC(cudaStreamSynchronize(c->stream));
  }
  // theta = 2 pi (a+1)/A  (radar_filters.cpp:317), host libm like the reference
  c->h_cs.resize(A);
  for (int a = 0; a < A; ++a) {
    const double theta = ((double)(a + 1) / A) * 2. * M_PI;
    c->h_cs[a] = make_double2(cos(theta), sin(theta));
  }
  CKC(cudaMemcpyAsync(c->d_cs, c->h_cs.data(), A * sizeof(double2), cudaMemcpyHostToDevice, c->stream));
  CKC(cudaStreamSynchronize(c->stream));
#undef CKC
  {   // buffer set 0 = the context's own buffers on the context stream
    PipeBufs& B0 = c->pb[0];
    B0.stream = c->stream;
    B0.d_kidx = c->d_kidx; B0.d_kcnt = c->d_kcnt; B0.d_rowcnt = c->d_rowcnt; B0.d_npts = c->d_npts; B0.d_status = c->d_status;
    B0.d_slots = c->d_slots; B0.d_rowcloud = c->d_rowcloud; B0.d_bufA = c->d_bufA; B0.d_bufB = c->d_bufB;
    B0.d_ghist = c->d_ghist; B0.d_celltmp = c->d_celltmp; B0.d_res = c->d_res;
  }
  *out = c;
  return CFEAR_OK;
}

int cfear_update_config(cfear_ctx* c, const cfear_config* cfg) {
  if (!c || !cfg) { g_err = "null argument"; return CFEAR_ERR_ARG; }
  const cfear_config& o = c->cfg;
  if (cfg->device != o.device || cfg->max_batch != o.max_batch || cfg->azimuths != o.azimuths ||
      cfg->range_bins != o.range_bins || cfg->k_strongest != o.k_strongest || cfg->max_keyframes != o.max_keyframes ||
      cfg->max_cellsets != o.max_cellsets || cfg->max_cells != o.max_cells || cfg->steps_in_flight != o.steps_in_flight) {
    g_err = "cfear_update_config: structural fields (device, max_batch, azimuths, range_bins, k_strongest, max_*) are fixed at create";
    return CFEAR_ERR_ARG;
  }
  if (!(cfg->radius > 0.f) || !(cfg->downsample_factor > 0.0) || cfg->cost < 0 || cfg->cost > 2 || cfg->loss < 0 || cfg->loss > 5) {
    g_err = "cfear_update_config: invalid radius / downsample_factor / cost / loss";
    return CFEAR_ERR_ARG;
  }
  c->cfg = *cfg;
  return CFEAR_OK;
}

int64_t cfear_launch_count(const cfear_ctx* c) { return c ? c->launches : 0; }
void* cfear_stream(cfear_ctx* c) { return c ? (void*)c->stream : nullptr; }
int cfear_sync(cfear_ctx* c) {
  if (!c) return CFEAR_ERR_ARG;
  CK(cudaSetDevice(c->cfg.device));
  if (c->inflight) {
    for (int i = 1; i <= std::max(c->npipes, 2); ++i) CK(cudaStreamWaitEvent(c->stream, c->pb[i].done, 0));
    c->inflight = false;
  }
  CK(cudaStreamSynchronize(c->stream));
  return CFEAR_OK;
}

void* cfear_alloc_pinned(size_t bytes) {
  void* p = nullptr;
  if (cudaHostAlloc(&p, bytes, cudaHostAllocDefault) != cudaSuccess) { g_err = "cudaHostAlloc failed"; return nullptr; }
  return p;
}
void cfear_free_pinned(void* p) { if (p) cudaFreeHost(p); }
void* cfear_alloc_device(cfear_ctx* c, size_t bytes) {
  if (!c) return nullptr;
  cudaSetDevice(c->cfg.device);
  void* p = nullptr;
  if (cudaMalloc(&p, bytes) != cudaSuccess) { g_err = "cudaMalloc failed"; return nullptr; }
  return p;
}
void cfear_free_device(cfear_ctx* c, void* p) { if (c && p) { cudaSetDevice(c->cfg.device); cudaFree(p); } }
int cfear_memcpy_h2d(cfear_ctx* c, void* dst, const void* src, size_t bytes) {
  if (!c) return CFEAR_ERR_ARG;
  CK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  return CFEAR_OK;
}
int cfear_memcpy_d2h(cfear_ctx* c, void* dst, const void* src, size_t bytes) {
  if (!c) return CFEAR_ERR_ARG;
  CK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  return CFEAR_OK;
}

}  // extern "C"

// ---- launch helpers (device-resident arguments) ----------------------------------------------------
static int launch_k1(cfear_ctx* c, const PipeBufs& B, const uint8_t* d_polar, int nscans) {
  K1Params p;
  p.polar = d_polar; p.nrows = nscans * c->cfg.azimuths; p.A = c->cfg.azimuths; p.R = c->cfg.range_bins;
  p.polar_end = d_polar + (size_t)p.nrows * p.R;
  p.zmin = (int)(uint8_t)(int)c->cfg.z_min;                                  // radar_driver.cpp:58, radar_filters.cpp:198,212
  p.k = c->cfg.k_strongest;
  const double range_res = (double)c->cfg.range_res, min_distance = (double)c->cfg.min_distance;
  p.min_range_bin = (int)ceil(min_distance / range_res);                      // radar_filters.cpp:315
  p.range_res = range_res; p.cs = c->d_cs;
  p.kidx = B.d_kidx; p.kcnt = B.d_kcnt; p.rowcloud = B.d_rowcloud; p.rowcnt = B.d_rowcnt;
  k1_launch(p, B.stream, c->prio[0]);
  c->launches++;
  CK(cudaGetLastError());
  return CFEAR_OK;
}

static int launch_peaks(cfear_ctx* c, const uint8_t* d_polar, int nscans) {
  PeaksParams p;
  p.polar = d_polar; p.nrows = nscans * c->cfg.azimuths; p.A = c->cfg.azimuths; p.R = c->cfg.range_bins;
  p.total = (long)p.nrows * p.R; p.k = c->cfg.k_strongest;
  const double range_res = (double)c->cfg.range_res, min_distance = (double)c->cfg.min_distance;
  p.min_range_bin = (int)ceil(min_distance / range_res);
  p.range_res = range_res; p.cs = c->d_cs; p.kidx = c->d_kidx; p.kcnt = c->d_kcnt;
  p.rowpeaks = c->d_rowpeaks; p.rowpeakcnt = c->d_rowpeakcnt;
  k1b_peaks<<<(p.nrows + 7) / 8, 256, 0, c->stream>>>(p);
  c->launches++;
  CK(cudaGetLastError());
  return CFEAR_OK;
}

// CFEAR_K3_WIDE=0 / CFEAR_K5_WIDE=0 in the environment keep small batches on the batch-sized forms of K3 / K5 (512 / 128
// threads).  Read at every launch so that the tests can run one context both ways; the choice never changes K3's results
// and moves K5's sums by rounding only.
static bool wide_allowed(const char* name) {
  const char* v = getenv(name);
  return !(v && v[0] == '0');
}
// One wide CTA per SM pays when every scan / problem in flight can have an SM of its own: a stream-ordered launch of at
// most num_sms of them, or overlapped steps whose sum stays near that (measured, five steps in flight: 32-problem steps
// 428 k scans/s wide vs 338 k batch-sized, 64-problem steps 549 k vs 572 k, 128: 550 k vs 767 k; profiles/r04g_*.txt).
static bool wide_batch(const cfear_ctx* c, int n) {
  return c->launch_conc <= 1 ? n <= c->num_sms : 2 * n * c->launch_conc <= 3 * c->num_sms;
}

static int launch_k3(cfear_ctx* c, const PipeBufs& B, int mode, int nscans, const double* d_mot, const int32_t* d_slots, bool write_cloud, int off = 0) {
  K3Params p;
  p.mode = mode; p.A = c->cfg.azimuths; p.k = c->cfg.k_strongest;
  p.rowcloud = B.d_rowcloud; p.rowcnt = B.d_rowcnt;
  p.mot = (c->cfg.compensate && mode == 0) ? d_mot : nullptr; p.ccw = c->cfg.radar_ccw; p.cs = c->d_cs;
  p.cloud = (mode == 1 || write_cloud) ? c->d_cloud : nullptr; p.npts = B.d_npts; p.cap_pts = c->cap_pts;
  p.slots = d_slots; p.radius = c->cfg.radius;
  p.leaf = (float)((double)c->cfg.radius / c->cfg.downsample_factor);        // pointnormal.cpp:279
  p.weight_intensity = c->cfg.weight_intensity; p.origin_x = 0.0; p.origin_y = 0.0;   // odometrykeyframefuser.cpp:161
  p.nn_cell = CFEAR_NN_CELL;
  p.pts_in_smem = c->pts_in_smem; p.g_bufA = B.d_bufA; p.g_bufB = B.d_bufB;
  p.g_hist = B.d_ghist; p.g_hist_cap = c->g_hist_cap; p.status = B.d_status; p.cell_tmp = B.d_celltmp; p.pool = c->pool;
  if (off) {   // sub-batch: every per-scan array starts at scan `off` (d_mot / d_slots are passed already offset)
    const size_t o = (size_t)off;
    p.rowcloud += o * p.A * p.k; p.rowcnt += o * p.A;
    if (p.cloud) p.cloud += o * p.cap_pts;
    p.npts += o; p.status += o; p.cell_tmp += o * p.cap_pts * 4;
    if (p.g_bufA) p.g_bufA += o * p.cap_pts;
    p.g_bufB += o * p.cap_pts;
    p.g_hist += o * (p.g_hist_cap + 1);
  }
  if (c->pts_in_smem && c->k3_wide && wide_batch(c, nscans) && wide_allowed("CFEAR_K3_WIDE"))          // at most one scan per SM: one 1024-thread CTA each
    CK(launch_with_priority(k3_surface_points<true, K3_THREADS_WIDE>, nscans, K3_THREADS_WIDE, c->k3_smem, B.stream, c->prio[1], p));
  else if (c->pts_in_smem) CK(launch_with_priority(k3_surface_points<true>, nscans, K3_THREADS, c->k3_smem, B.stream, c->prio[1], p));
  else CK(launch_with_priority(k3_surface_points<false>, nscans, K3_THREADS, c->k3_smem, B.stream, c->prio[1], p));
  c->launches++;
  CK(cudaGetLastError());
  return CFEAR_OK;
}

static int launch_k5(cfear_ctx* c, const PipeBufs& B, int nprob, int nscans, const int32_t* d_slots, double* d_poses, double* d_cov36,
                     cfear_reg_stats* d_stats, int32_t* d_assoc, int off = 0, const int32_t* d_nscans_pp = nullptr,
                     int solver_mode_override = -1, double* d_assoc_sim = nullptr, const double* d_softL = nullptr) {
  RegParams p;
  p.pool = c->pool; p.nprob = nprob; p.nscans = nscans; p.nscans_pp = d_nscans_pp; p.slots = d_slots; p.poses = d_poses; p.cov36 = d_cov36;
  p.stats = d_stats; p.assoc = d_assoc; p.assoc_sim = d_assoc_sim; p.soft_L = d_softL; p.res = B.d_res + (size_t)off * c->res_cap * 4; p.res_cap = c->res_cap;
  p.cost = c->cfg.cost; p.loss = c->cfg.loss; p.weight_opt = c->cfg.weight_opt; p.solver_mode = c->cfg.solver_mode;
  p.max_outer = c->cfg.max_outer; p.min_outer = c->cfg.min_outer; p.max_inner = c->cfg.max_inner; p.gn_iters = c->cfg.gn_iters;
  p.loss_limit = c->cfg.loss_limit; p.cov_scale = c->cfg.cov_scale; p.regularization = c->cfg.regularization;
  p.radius = c->cfg.reg_radius;
  if (solver_mode_override >= 0) p.solver_mode = solver_mode_override;
  if (p.cost < 0 || p.cost > 2 || p.loss < 0 || p.loss > 5) { g_err = "unknown cost / loss type"; return CFEAR_ERR_ARG; }
  p.smem_bytes = c->k5_smem;
  // launch form: 2 = one 384-thread CTA per SM (at most one problem per SM), 1 = two 192-thread CTAs per SM (a stream-ordered
  // launch of at most two problems per SM), 0 = three 128-thread CTAs per SM.  CFEAR_K5_FORM=0|1|2 forces one (tests, A/B runs).
  int form = 0;
  if (wide_allowed("CFEAR_K5_WIDE")) {
    if (wide_batch(c, nprob)) form = 2;
    else if (c->launch_conc <= 1 && nprob <= 2 * c->num_sms) form = 1;
  }
  if (const char* f = getenv("CFEAR_K5_FORM")) { if (f[0] >= '0' && f[0] <= '2') form = f[0] - '0'; }
  const int smem_form = form == 2 ? c->k5_smem_wide : c->k5_smem_mid;
  bool launched = false;
  switch (p.cost) {
    case 0: launched = k5_launch_cost0(p, nprob, c->k5_smem, form, smem_form, B.stream, c->prio[2]); break;
    case 1: launched = k5_launch_cost1(p, nprob, c->k5_smem, form, smem_form, B.stream, c->prio[2]); break;
    case 2: launched = k5_launch_cost2(p, nprob, c->k5_smem, form, smem_form, B.stream, c->prio[2]); break;
  }
  if (!launched) { g_err = "this build has no instantiation for the requested cost / loss"; return CFEAR_ERR_ARG; }
  c->launches++;
  CK(cudaGetLastError());
  return CFEAR_OK;
}

static int begin_timed_step(cfear_ctx* c) {
  if (!c->timing) return CFEAR_OK;
  const int cap_steps = 4096;
  if (c->ev_used >= cap_steps) c->ev_used = cap_steps - 1;
  while ((int)c->ev.size() < 4 * (c->ev_used + 1)) {
    cudaEvent_t e; CK(cudaEventCreate(&e));
    c->ev.push_back(e);
  }
  c->evset = c->ev.data() + 4 * c->ev_used;
  c->ev_used++;
  return CFEAR_OK;
}

// Stream-ordered temporaries that are released on every exit path of an entry point (CK / RC return early on errors).
struct TempBufs {
  cudaStream_t stream;
  std::vector<void*> ptrs;
  explicit TempBufs(cudaStream_t s) : stream(s) {}
  template <typename T> cudaError_t get(T** p, size_t bytes) {
    void* q = nullptr;
    cudaError_t e = cudaMallocAsync(&q, std::max<size_t>(bytes, 1), stream);
    if (e == cudaSuccess) { ptrs.push_back(q); *p = reinterpret_cast<T*>(q); }
    return e;
  }
  ~TempBufs() { for (void* q : ptrs) cudaFreeAsync(q, stream); }
};

static int check_slot(cfear_ctx* c, int slot) {
  if (slot < 0 || slot >= c->cfg.max_cellsets) { g_err = "cell-set slot out of range"; return CFEAR_ERR_ARG; }
  return CFEAR_OK;
}
#define RC(expr) do { int rc__ = (expr); if (rc__ != CFEAR_OK) return rc__; } while (0)
// The overlapped device-resident steps (cfear_odometry_step_batch_dev_submit) run on their own streams; every other
// entry point works on the context stream and first makes it wait, on the device, for the steps still in flight.
static int join_pipes(cfear_ctx* c) {
  if (!c->inflight) return CFEAR_OK;
  for (int i = 1; i <= std::max(c->npipes, 2); ++i) CK(cudaStreamWaitEvent(c->stream, c->pb[i].done, 0));
  c->inflight = false;
  return CFEAR_OK;
}
#define ENTER_NOJOIN(c) do { if (!(c)) { g_err = "null context"; return CFEAR_ERR_ARG; } CK(cudaSetDevice((c)->cfg.device)); } while (0)
#define ENTER(c) do { ENTER_NOJOIN(c); RC(join_pipes(c)); } while (0)

extern "C" {

int cfear_filter(cfear_ctx* c, const uint8_t* polar, int nscans, int32_t* idx_out, int32_t* cnt_out,
                 cfear_point* cloud_out, int32_t* npts_out, cfear_point* peaks_out, int32_t* npeaks_out) {
  ENTER(c);
  if (!polar || nscans < 0) { g_err = "null image"; return CFEAR_ERR_ARG; }     // radar_driver.cpp:75-78
  if (nscans > c->cfg.max_batch) { g_err = "nscans exceeds max_batch"; return CFEAR_ERR_CAPACITY; }
  if (nscans == 0) return CFEAR_OK;
  const int A = c->cfg.azimuths, R = c->cfg.range_bins, k = c->cfg.k_strongest;
  const size_t rows = (size_t)nscans * A;
  c->polar_dirty = true;
  CK(cudaMemcpyAsync(c->d_polar, polar, rows * R, cudaMemcpyHostToDevice, c->stream));
  RC(launch_k1(c, c->pb[0], c->d_polar, nscans));
  if (idx_out) CK(cudaMemcpyAsync(idx_out, c->d_kidx, rows * k * sizeof(int32_t), cudaMemcpyDeviceToHost, c->stream));
  if (cnt_out) CK(cudaMemcpyAsync(cnt_out, c->d_kcnt, rows * sizeof(int32_t), cudaMemcpyDeviceToHost, c->stream));
  if (cloud_out || npts_out) {
    k2_compact_rows<<<nscans, 512, A * sizeof(int), c->stream>>>(c->d_rowcloud, c->d_rowcnt, A, k, c->cap_pts, c->d_cloud, c->d_npts);
    c->launches++;
    CK(cudaGetLastError());
    if (cloud_out) CK(cudaMemcpyAsync(cloud_out, c->d_cloud, (size_t)nscans * c->cap_pts * sizeof(float4), cudaMemcpyDeviceToHost, c->stream));
    if (npts_out) CK(cudaMemcpyAsync(npts_out, c->d_npts, nscans * sizeof(int32_t), cudaMemcpyDeviceToHost, c->stream));
  }
  if (peaks_out || npeaks_out) {
    RC(launch_peaks(c, c->d_polar, nscans));
    k2_compact_rows<<<nscans, 512, A * sizeof(int), c->stream>>>(c->d_rowpeaks, c->d_rowpeakcnt, A, k, c->cap_pts, c->d_peaks, c->d_npeaks);
    c->launches++;
    CK(cudaGetLastError());
    if (peaks_out) CK(cudaMemcpyAsync(peaks_out, c->d_peaks, (size_t)nscans * c->cap_pts * sizeof(float4), cudaMemcpyDeviceToHost, c->stream));
    if (npeaks_out) CK(cudaMemcpyAsync(npeaks_out, c->d_npeaks, nscans * sizeof(int32_t), cudaMemcpyDeviceToHost, c->stream));
  }
  CK(cudaStreamSynchronize(c->stream));
  return CFEAR_OK;
}

int cfear_kstrongest(cfear_ctx* c, const uint8_t* polar, int nscans, int32_t* idx_out, int32_t* cnt_out) {
  return cfear_filter(c, polar, nscans, idx_out, cnt_out, nullptr, nullptr, nullptr, nullptr);
}

int cfear_compensate(cfear_ctx* c, cfear_point* cloud, int n, const double mot[3], int ccw) {
  ENTER(c);
  if (n < 0 || (n > 0 && !cloud) || !mot) { g_err = "bad cloud"; return CFEAR_ERR_ARG; }
  if (n > c->cfg.max_batch * c->cap_pts) { g_err = "cloud exceeds capacity"; return CFEAR_ERR_CAPACITY; }
  if (n == 0) return CFEAR_OK;
  CK(cudaMemcpyAsync(c->d_cloud, cloud, (size_t)n * sizeof(float4), cudaMemcpyHostToDevice, c->stream));
  k2b_compensate<<<(n + 255) / 256, 256, 0, c->stream>>>(c->d_cloud, n, mot[0], mot[1], mot[2], ccw);
  c->launches++;
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(cloud, c->d_cloud, (size_t)n * sizeof(float4), cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  return CFEAR_OK;
}

int cfear_surface_points(cfear_ctx* c, const cfear_point* cloud, int n, int slot, int32_t* ncells) {
  ENTER(c);
  RC(check_slot(c, slot));
  if (n < 0 || (n > 0 && !cloud)) { g_err = "bad cloud"; return CFEAR_ERR_ARG; }
  if (n > c->cap_pts) { g_err = "cloud exceeds azimuths*k_strongest points"; return CFEAR_ERR_CAPACITY; }
  if (n > 0) CK(cudaMemcpyAsync(c->d_cloud, cloud, (size_t)n * sizeof(float4), cudaMemcpyHostToDevice, c->stream));
  const int32_t n32 = n, s32 = slot;
  CK(cudaMemcpyAsync(c->d_npts, &n32, sizeof(int32_t), cudaMemcpyHostToDevice, c->stream));
  CK(cudaMemcpyAsync(c->d_curslots, &s32, sizeof(int32_t), cudaMemcpyHostToDevice, c->stream));
  RC(launch_k3(c, c->pb[0], 1, 1, nullptr, c->d_curslots, false));
  int32_t st = 0, nc = 0;
  CK(cudaMemcpyAsync(&st, c->d_status, sizeof(int32_t), cudaMemcpyDeviceToHost, c->stream));
  CK(cudaMemcpyAsync(&nc, c->pool.ncells + slot, sizeof(int32_t), cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  if (st != 0) { g_err = "voxel grid exceeds capacity (extent / leaf too large)"; return CFEAR_ERR_CAPACITY; }
  if (ncells) *ncells = nc;
  return CFEAR_OK;
}

int cfear_scans_to_cells_batch(cfear_ctx* c, int nscans, const uint8_t* polar, const double* mot, const int32_t* slots,
                               int32_t* npts_out, int32_t* ncells_out) {
  ENTER(c);
  if (!polar || !slots) { g_err = "null argument"; return CFEAR_ERR_ARG; }
  if (nscans < 0 || nscans > c->cfg.max_batch) { g_err = "nscans exceeds max_batch"; return CFEAR_ERR_CAPACITY; }
  if (nscans == 0) return CFEAR_OK;
  for (int i = 0; i < nscans; ++i) RC(check_slot(c, slots[i]));
  const size_t img = (size_t)c->cfg.azimuths * c->cfg.range_bins;
  c->polar_dirty = true;
  CK(cudaMemcpyAsync(c->d_polar, polar, (size_t)nscans * img, cudaMemcpyHostToDevice, c->stream));
  CK(cudaMemcpyAsync(c->d_curslots, slots, (size_t)nscans * sizeof(int32_t), cudaMemcpyHostToDevice, c->stream));
  const bool have_mot = mot != nullptr && c->cfg.compensate;
  if (have_mot) CK(cudaMemcpyAsync(c->d_mot, mot, (size_t)nscans * 3 * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  c->last_pipe = 0;
  RC(launch_k1(c, c->pb[0], c->d_polar, nscans));
  RC(launch_k3(c, c->pb[0], 0, nscans, have_mot ? c->d_mot : nullptr, c->d_curslots, false));
  std::vector<int32_t> st(nscans);
  CK(cudaMemcpyAsync(st.data(), c->d_status, (size_t)nscans * sizeof(int32_t), cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  for (int i = 0; i < nscans; ++i)
    if (st[i] != 0) { g_err = "voxel grid exceeds capacity (extent / leaf too large)"; return CFEAR_ERR_CAPACITY; }
  if (npts_out || ncells_out) RC(cfear_last_counts(c, nscans, slots, npts_out, ncells_out));
  return CFEAR_OK;
}

int cfear_cells_count(cfear_ctx* c, int slot, int32_t* ncells) {
  ENTER(c);
  RC(check_slot(c, slot));
  if (!ncells) { g_err = "null output"; return CFEAR_ERR_ARG; }
  CK(cudaMemcpyAsync(ncells, c->pool.ncells + slot, sizeof(int32_t), cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  return CFEAR_OK;
}

int cfear_cells_download(cfear_ctx* c, int slot, cfear_cell* out, int capacity, int32_t* ncells) {
  ENTER(c);
  RC(check_slot(c, slot));
  int32_t n = 0;
  CK(cudaMemcpyAsync(&n, c->pool.ncells + slot, sizeof(int32_t), cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  if (ncells) *ncells = n;
  if (!out) return CFEAR_OK;
  if (capacity < n) { g_err = "output capacity too small"; return CFEAR_ERR_CAPACITY; }
  if (n == 0) return CFEAR_OK;
  k_cells_gather<<<(n + 255) / 256, 256, 0, c->stream>>>(c->pool, slot, c->d_cellaos, n);
  c->launches++;
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(out, c->d_cellaos, (size_t)n * sizeof(CellAoS), cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  return CFEAR_OK;
}

int cfear_cells_upload(cfear_ctx* c, int slot, const cfear_cell* cells, int n) {
  ENTER(c);
  RC(check_slot(c, slot));
  if (n < 0 || (n > 0 && !cells)) { g_err = "bad cells"; return CFEAR_ERR_ARG; }
  if (n > c->max_cells) { g_err = "cell set exceeds max_cells"; return CFEAR_ERR_CAPACITY; }
  if (n > 0) CK(cudaMemcpyAsync(c->d_cellaos, cells, (size_t)n * sizeof(CellAoS), cudaMemcpyHostToDevice, c->stream));
  k_cells_scatter<<<(std::max(n, 1) + 255) / 256, 256, 0, c->stream>>>(c->pool, slot, c->d_cellaos, n);
  c->launches++;
  CK(cudaGetLastError());
  const int32_t s32 = slot;
  CK(cudaMemcpyAsync(c->d_curslots, &s32, sizeof(int32_t), cudaMemcpyHostToDevice, c->stream));
  K4Params p; p.pool = c->pool; p.slots = c->d_curslots; p.nn_cell = CFEAR_NN_CELL;
  k4_build_index<<<1, K3_THREADS, c->k4_smem, c->stream>>>(p);
  c->launches++;
  CK(cudaGetLastError());
  CK(cudaStreamSynchronize(c->stream));
  return CFEAR_OK;
}

int cfear_nearest(cfear_ctx* c, int slot, const double* queries_xy, int nq, double radius, int32_t* out_idx) {
  ENTER(c);
  RC(check_slot(c, slot));
  if (nq < 0 || (nq > 0 && (!queries_xy || !out_idx))) { g_err = "bad queries"; return CFEAR_ERR_ARG; }
  if (nq == 0) return CFEAR_OK;
  TempBufs tmp(c->stream);
  double* dq = nullptr; int32_t* dout = nullptr;
  CK(tmp.get(&dq, (size_t)nq * 2 * sizeof(double)));
  CK(tmp.get(&dout, (size_t)nq * sizeof(int32_t)));
  CK(cudaMemcpyAsync(dq, queries_xy, (size_t)nq * 2 * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  k4b_nearest<<<(nq + 255) / 256, 256, 0, c->stream>>>(c->pool, slot, dq, nq, radius, dout);
  c->launches++;
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(out_idx, dout, (size_t)nq * sizeof(int32_t), cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  return CFEAR_OK;
}

int cfear_register_batch_ex(cfear_ctx* c, int nprob, const int32_t* slots, int nscans, double* poses, double* cov36,
                            cfear_reg_stats* stats, int32_t* assoc_out, double* assoc_sim_out, const double* prior_sqrt_info) {
  ENTER(c);
  if (nprob < 0 || !slots || !poses) { g_err = "null argument"; return CFEAR_ERR_ARG; }
  if (nscans < 2 || nscans > c->cfg.max_keyframes + 1) { g_err = "nscans must be in [2, max_keyframes+1]"; return CFEAR_ERR_ARG; }   // n_scan_normal.cpp:190
  if (nprob > c->cfg.max_batch) { g_err = "nprob exceeds max_batch"; return CFEAR_ERR_CAPACITY; }
  if (prior_sqrt_info && c->cfg.solver_mode != CFEAR_SOLVER_CERES_LM) { g_err = "the soft prior belongs to Register()'s ceres_lm loop"; return CFEAR_ERR_ARG; }
  if (nprob == 0) return CFEAR_OK;
  for (size_t i = 0; i < (size_t)nprob * nscans; ++i) RC(check_slot(c, slots[i]));
  CK(cudaMemcpyAsync(c->d_slots, slots, (size_t)nprob * nscans * sizeof(int32_t), cudaMemcpyHostToDevice, c->stream));
  CK(cudaMemcpyAsync(c->d_poses, poses, (size_t)nprob * nscans * 3 * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  const size_t assoc_cap = (size_t)c->cfg.max_batch * c->cfg.max_keyframes * c->max_cells;
  const size_t assoc_n = (size_t)nprob * (nscans - 1) * c->max_cells;
  if (assoc_out || assoc_sim_out) {                    // the similarity is only meaningful next to the association table
    if (!c->d_assoc) RC(c->alloc(&c->d_assoc, assoc_cap));
    CK(cudaMemsetAsync(c->d_assoc, 0xff, assoc_n * sizeof(int32_t), c->stream));
  }
  if (assoc_sim_out) {
    if (!c->d_assoc_sim) RC(c->alloc(&c->d_assoc_sim, assoc_cap));
    CK(cudaMemsetAsync(c->d_assoc_sim, 0, assoc_n * sizeof(double), c->stream));
  }
  if (prior_sqrt_info) {
    if (!c->d_softL) RC(c->alloc(&c->d_softL, (size_t)c->cfg.max_batch * 9));
    CK(cudaMemcpyAsync(c->d_softL, prior_sqrt_info, (size_t)nprob * 9 * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  }
  RC(launch_k5(c, c->pb[0], nprob, nscans, c->d_slots, c->d_poses, c->d_cov36, c->d_stats, (assoc_out || assoc_sim_out) ? c->d_assoc : nullptr,
               0, nullptr, -1, assoc_sim_out ? c->d_assoc_sim : nullptr, prior_sqrt_info ? c->d_softL : nullptr));
  CK(cudaMemcpyAsync(poses, c->d_poses, (size_t)nprob * nscans * 3 * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  if (cov36) CK(cudaMemcpyAsync(cov36, c->d_cov36, (size_t)nprob * 36 * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  if (stats) CK(cudaMemcpyAsync(stats, c->d_stats, (size_t)nprob * sizeof(cfear_reg_stats), cudaMemcpyDeviceToHost, c->stream));
  if (assoc_out) CK(cudaMemcpyAsync(assoc_out, c->d_assoc, assoc_n * sizeof(int32_t), cudaMemcpyDeviceToHost, c->stream));
  if (assoc_sim_out) CK(cudaMemcpyAsync(assoc_sim_out, c->d_assoc_sim, assoc_n * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  return CFEAR_OK;
}

int cfear_register_batch(cfear_ctx* c, int nprob, const int32_t* slots, int nscans, double* poses, double* cov36,
                         cfear_reg_stats* stats, int32_t* assoc_out) {
  return cfear_register_batch_ex(c, nprob, slots, nscans, poses, cov36, stats, assoc_out, nullptr, nullptr);
}

// n_scan_normal_reg::GetCost for nprob independent (cell sets, poses) problems in one launch.
int cfear_get_cost_batch(cfear_ctx* c, int nprob, const int32_t* slots, int nscans, const double* poses,
                         double* cost_out, int32_t* num_residuals_out, int32_t* ok_out) {
  ENTER(c);
  if (nprob < 0 || !slots || !poses || !cost_out) { g_err = "null argument"; return CFEAR_ERR_ARG; }
  if (nscans < 2 || nscans > c->cfg.max_keyframes + 1) { g_err = "nscans must be in [2, max_keyframes+1]"; return CFEAR_ERR_ARG; }   // n_scan_normal.cpp:189
  if (nprob > c->cfg.max_batch) { g_err = "nprob exceeds max_batch"; return CFEAR_ERR_CAPACITY; }
  if (nprob == 0) return CFEAR_OK;
  for (size_t i = 0; i < (size_t)nprob * nscans; ++i) RC(check_slot(c, slots[i]));
  CK(cudaMemcpyAsync(c->d_slots, slots, (size_t)nprob * nscans * sizeof(int32_t), cudaMemcpyHostToDevice, c->stream));
  CK(cudaMemcpyAsync(c->d_poses, poses, (size_t)nprob * nscans * 3 * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  RC(launch_k5(c, c->pb[0], nprob, nscans, c->d_slots, c->d_poses, c->d_cov36, c->d_stats, nullptr, 0, nullptr, CFEAR_SOLVER_COST_ONLY));
  std::vector<cfear_reg_stats> st((size_t)nprob);
  CK(cudaMemcpyAsync(st.data(), c->d_stats, (size_t)nprob * sizeof(cfear_reg_stats), cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  for (int i = 0; i < nprob; ++i) {
    cost_out[i] = st[i].final_cost;
    if (num_residuals_out) num_residuals_out[i] = st[i].num_residuals;
    if (ok_out) ok_out[i] = st[i].success;
  }
  return CFEAR_OK;
}

int cfear_register(cfear_ctx* c, const int32_t* slots, int nscans, double* poses, double* cov36, cfear_reg_stats* stats) {
  return cfear_register_batch(c, 1, slots, nscans, poses, cov36, stats, nullptr);
}

// d_kf_slots [nprob][K], d_cur_slots [nprob] -> d_slots [nprob][K+1]
__global__ void k_merge_slots(const int32_t* kf, const int32_t* cur, int K, int nprob, int32_t* out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nprob * (K + 1)) return;
  const int b = i / (K + 1), j = i - b * (K + 1);
  out[i] = (j < K) ? kf[b * K + j] : cur[b];
}

static int step_dev(cfear_ctx* c, const PipeBufs& B, int nprob, const uint8_t* d_polar, const double* d_mot, const int32_t* d_kf_slots, int K,
                    const int32_t* d_cur_slots, double* d_poses, double* d_cov36, cfear_reg_stats* d_stats) {
  RC(begin_timed_step(c));
  if (c->timing) CK(cudaEventRecord(c->evset[0], B.stream));
  RC(launch_k1(c, B, d_polar, nprob));
  if (c->timing) CK(cudaEventRecord(c->evset[1], B.stream));
  RC(launch_k3(c, B, 0, nprob, d_mot, d_cur_slots, false));
  if (c->timing) CK(cudaEventRecord(c->evset[2], B.stream));
  k_merge_slots<<<(nprob * (K + 1) + 255) / 256, 256, 0, B.stream>>>(d_kf_slots, d_cur_slots, K, nprob, B.d_slots);
  c->launches++;
  CK(cudaGetLastError());
  RC(launch_k5(c, B, nprob, K + 1, B.d_slots, d_poses, d_cov36, d_stats, nullptr));
  if (c->timing) CK(cudaEventRecord(c->evset[3], B.stream));
  return CFEAR_OK;
}

int cfear_odometry_step_batch_dev(cfear_ctx* c, int nprob, const uint8_t* d_polar, const double* d_mot,
                                  const int32_t* d_kf_slots, int K, const int32_t* d_cur_slots,
                                  double* d_poses, double* d_cov36, cfear_reg_stats* d_stats) {
  ENTER(c);
  if (!d_polar || !d_kf_slots || !d_cur_slots || !d_poses || !d_cov36 || !d_stats) { g_err = "null argument"; return CFEAR_ERR_ARG; }
  if (K < 1 || K > c->cfg.max_keyframes) { g_err = "K must be in [1, max_keyframes]"; return CFEAR_ERR_ARG; }
  if (nprob < 0 || nprob > c->cfg.max_batch) { g_err = "nprob exceeds max_batch"; return CFEAR_ERR_CAPACITY; }
  if (nprob == 0) return CFEAR_OK;
  c->last_pipe = 0;
  return step_dev(c, c->pb[0], nprob, d_polar, d_mot, d_kf_slots, K, d_cur_slots, d_poses, d_cov36, d_stats);
}

static int ensure_tickets(cfear_ctx* c) {
  while ((int)c->ticket_ev.size() < CFEAR_MAX_TICKETS) {
    cudaEvent_t e; CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    c->ticket_ev.push_back(e);
  }
  return CFEAR_OK;
}

// The overlapped batch steps and the sequence replay use the internal streams / scratch sets differently: when the user
// changes, what is in flight is joined into the context stream first (the next step then orders itself after it).
static int set_pipe_user(cfear_ctx* c, int user) {
  if (c->pipe_user != user) { RC(join_pipes(c)); c->pipe_user = user; }
  return CFEAR_OK;
}

// Streams, events and the extra buffer sets of the overlapped device-resident steps, created on first use.
static int ensure_pipes(cfear_ctx* c) {
  if (c->pipes_ready) return CFEAR_OK;
  const int A = c->cfg.azimuths, k = c->cfg.k_strongest, B = c->cfg.max_batch;
  const size_t rows = (size_t)B * A;
  for (int i = 1; i <= std::max(c->npipes, 2); ++i) {       // the sequence replay always uses two
    PipeBufs& P = c->pb[i];
    CK(cudaStreamCreateWithFlags(&P.stream, cudaStreamNonBlocking));
    CK(cudaEventCreateWithFlags(&P.in, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&P.done, cudaEventDisableTiming));
    CK(cudaEventRecord(P.done, P.stream));
#define ALP(p, n) do { int rc__ = c->alloc(&(p), (size_t)(n)); if (rc__ != CFEAR_OK) return rc__; } while (0)
    ALP(P.d_kidx, rows * k); ALP(P.d_kcnt, rows); ALP(P.d_rowcnt, rows); ALP(P.d_rowcloud, rows * k);
    ALP(P.d_npts, B); ALP(P.d_status, B);
    if (!c->pts_in_smem) ALP(P.d_bufA, (size_t)B * c->cap_pts);
    ALP(P.d_bufB, (size_t)B * c->cap_pts);
    ALP(P.d_ghist, (size_t)B * (c->g_hist_cap + 1));
    ALP(P.d_celltmp, (size_t)B * c->cap_pts * 4);
    ALP(P.d_slots, (size_t)B * (c->cfg.max_keyframes + 1));
    ALP(P.d_res, (size_t)B * c->res_cap * 4);
#undef ALP
  }
  c->pipes_ready = true;
  return CFEAR_OK;
}

// Device-resident step, overlapped: step i goes to stream i mod CFEAR_NPIPES with its own K1..K5 scratch, so the filter
// and surface-point kernels of one step run while the registration of the previous one is still in its tail (K5 lasts as
// long as its slowest problem and leaves most SMs idle by then).  Inputs are ordered after the work already enqueued on
// the context stream; outputs are complete at cfear_odometry_step_batch_wait(ticket) (host) /
// cfear_stream_wait_ticket (context stream) / cfear_join / cfear_sync.
int cfear_odometry_step_batch_dev_submit(cfear_ctx* c, int nprob, const uint8_t* d_polar, const double* d_mot,
                                         const int32_t* d_kf_slots, int K, const int32_t* d_cur_slots,
                                         double* d_poses, double* d_cov36, cfear_reg_stats* d_stats, int32_t* ticket_out) {
  ENTER_NOJOIN(c);
  if (!d_polar || !d_kf_slots || !d_cur_slots || !d_poses || !d_cov36 || !d_stats) { g_err = "null argument"; return CFEAR_ERR_ARG; }
  if (K < 1 || K > c->cfg.max_keyframes) { g_err = "K must be in [1, max_keyframes]"; return CFEAR_ERR_ARG; }
  if (nprob < 0 || nprob > c->cfg.max_batch) { g_err = "nprob exceeds max_batch"; return CFEAR_ERR_CAPACITY; }
  RC(ensure_tickets(c));
  RC(ensure_pipes(c));
  RC(set_pipe_user(c, 1));
  const int ticket = c->next_ticket++ % CFEAR_MAX_TICKETS;
  if (ticket_out) *ticket_out = ticket;
  const int pi = 1 + c->next_pipe;
  c->next_pipe = (c->next_pipe + 1) % c->npipes;
  PipeBufs& B = c->pb[pi];
  CK(cudaEventRecord(B.in, c->stream));
  CK(cudaStreamWaitEvent(B.stream, B.in, 0));
  if (nprob > 0) {
    c->launch_conc = c->npipes;                          // this step shares the GPU with the other steps in flight
    const int rc = step_dev(c, B, nprob, d_polar, d_mot, d_kf_slots, K, d_cur_slots, d_poses, d_cov36, d_stats);
    c->launch_conc = 1;
    if (rc != CFEAR_OK) return rc;
  }
  CK(cudaEventRecord(B.done, B.stream));
  CK(cudaEventRecord(c->ticket_ev[ticket], B.stream));
  c->inflight = true; c->last_pipe = pi;
  return CFEAR_OK;
}

int cfear_stream_wait_ticket(cfear_ctx* c, int32_t ticket) {
  ENTER_NOJOIN(c);
  if (ticket < 0 || ticket >= (int)c->ticket_ev.size()) { g_err = "unknown ticket"; return CFEAR_ERR_ARG; }
  CK(cudaStreamWaitEvent(c->stream, c->ticket_ev[ticket], 0));
  return CFEAR_OK;
}

int cfear_join(cfear_ctx* c) {
  ENTER(c);
  return CFEAR_OK;
}

// Host-buffer path, asynchronous half: enqueues the whole step (H2D in sub-batches on the copy stream, K1 -> K3 -> K5
// per sub-batch on the compute stream, D2H of the results) and records the ticket's event.  Nothing here waits for the
// device, so a caller that submits step i+1 before waiting for step i keeps the PCIe link busy while the
// registration tail of step i (the slowest problem of its last sub-batch) finishes.
int cfear_odometry_step_batch_submit(cfear_ctx* c, int nprob, const uint8_t* polar, const double* mot,
                                     const int32_t* kf_slots, int K, const int32_t* cur_slots,
                                     double* poses, double* cov36, cfear_reg_stats* stats,
                                     int32_t* npts_out, int32_t* ticket_out) {
  ENTER(c);
  if (!polar || !kf_slots || !cur_slots || !poses) { g_err = "null argument"; return CFEAR_ERR_ARG; }
  if (K < 1 || K > c->cfg.max_keyframes) { g_err = "K must be in [1, max_keyframes]"; return CFEAR_ERR_ARG; }
  if (nprob < 0 || nprob > c->cfg.max_batch) { g_err = "nprob exceeds max_batch"; return CFEAR_ERR_CAPACITY; }
  for (int i = 0; i < nprob * K; ++i) RC(check_slot(c, kf_slots[i]));
  for (int i = 0; i < nprob; ++i) RC(check_slot(c, cur_slots[i]));
  RC(ensure_tickets(c));
  c->last_pipe = 0;
  const int ticket = c->next_ticket++ % CFEAR_MAX_TICKETS;
  if (ticket_out) *ticket_out = ticket;
  if (nprob == 0) { CK(cudaEventRecord(c->ticket_ev[ticket], c->stream)); return CFEAR_OK; }
  const int A = c->cfg.azimuths, R = c->cfg.range_bins;
  const size_t img = (size_t)A * R;
  int32_t* d_kf = c->d_kfslots;
  CK(cudaMemcpyAsync(d_kf, kf_slots, (size_t)nprob * K * sizeof(int32_t), cudaMemcpyHostToDevice, c->stream));
  CK(cudaMemcpyAsync(c->d_curslots, cur_slots, (size_t)nprob * sizeof(int32_t), cudaMemcpyHostToDevice, c->stream));
  CK(cudaMemcpyAsync(c->d_poses, poses, (size_t)nprob * (K + 1) * 3 * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  const bool have_mot = mot != nullptr && c->cfg.compensate;
  if (have_mot) CK(cudaMemcpyAsync(c->d_mot, mot, (size_t)nprob * 3 * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  // Sub-batches of `chunk` scans: the copy stream moves sub-batch i+1 host->device while the compute stream
  // runs K1 -> K3 -> K5 on sub-batch i.  The image staging area of sub-batch i is free again as soon as ITS K1 has
  // read it (k1_done[i]), not when the whole previous step is done.
  const int chunk = 32;
  const int nchunks = (nprob + chunk - 1) / chunk;
  while ((int)c->chunk_ev.size() < nchunks) {
    cudaEvent_t e; CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    c->chunk_ev.push_back(e);
    cudaEvent_t f; CK(cudaEventCreateWithFlags(&f, cudaEventDisableTiming));
    c->k1_done.push_back(f);
  }
  RC(begin_timed_step(c));
  const int timing = c->timing;
  if (timing) {   // stages interleave on this path: the whole step is reported under [2], nothing under [0], [1]
    CK(cudaEventRecord(c->evset[0], c->stream));
    CK(cudaEventRecord(c->evset[1], c->stream));
    CK(cudaEventRecord(c->evset[2], c->stream));
  }
  k_merge_slots<<<(nprob * (K + 1) + 255) / 256, 256, 0, c->stream>>>(d_kf, c->d_curslots, K, nprob, c->d_slots);
  c->launches++;
  CK(cudaGetLastError());
  if (c->polar_dirty) {       // another entry point wrote the staging area on the compute stream since the last step
    CK(cudaEventRecord(c->polar_free, c->stream));
    CK(cudaStreamWaitEvent(c->copy_stream, c->polar_free, 0));
    c->polar_dirty = false;
  }
  for (int ch = 0; ch < nchunks; ++ch) {
    const int b0 = ch * chunk, nb = std::min(chunk, nprob - b0);
    CK(cudaStreamWaitEvent(c->copy_stream, c->k1_done[ch], 0));              // no-op until the event has been recorded once
    CK(cudaMemcpyAsync(c->d_polar + b0 * img, polar + b0 * img, nb * img, cudaMemcpyHostToDevice, c->copy_stream));
    CK(cudaEventRecord(c->chunk_ev[ch], c->copy_stream));
    CK(cudaStreamWaitEvent(c->stream, c->chunk_ev[ch], 0));
    K1Params p;                                                              // K1 on this sub-batch
    p.polar = c->d_polar + b0 * img; p.nrows = nb * A; p.A = A; p.R = R;
    p.polar_end = c->d_polar + (size_t)nprob * img;
    p.zmin = (int)(uint8_t)(int)c->cfg.z_min; p.k = c->cfg.k_strongest;
    const double range_res = (double)c->cfg.range_res, min_distance = (double)c->cfg.min_distance;
    p.min_range_bin = (int)ceil(min_distance / range_res); p.range_res = range_res; p.cs = c->d_cs;
    const size_t r0 = (size_t)b0 * A;
    p.kidx = c->d_kidx + r0 * p.k; p.kcnt = c->d_kcnt + r0; p.rowcloud = c->d_rowcloud + r0 * p.k; p.rowcnt = c->d_rowcnt + r0;
    k1_launch(p, c->stream);
    c->launches++;
    CK(cudaGetLastError());
    CK(cudaEventRecord(c->k1_done[ch], c->stream));
    RC(launch_k3(c, c->pb[0], 0, nb, have_mot ? c->d_mot + 3 * (size_t)b0 : nullptr, c->d_curslots + b0, false, b0));
    RC(launch_k5(c, c->pb[0], nb, K + 1, c->d_slots + (size_t)b0 * (K + 1), c->d_poses + (size_t)b0 * (K + 1) * 3,
                 c->d_cov36 + (size_t)b0 * 36, c->d_stats + b0, nullptr, b0));
  }
  if (timing) CK(cudaEventRecord(c->evset[3], c->stream));
  CK(cudaMemcpyAsync(poses, c->d_poses, (size_t)nprob * (K + 1) * 3 * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  if (cov36) CK(cudaMemcpyAsync(cov36, c->d_cov36, (size_t)nprob * 36 * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  if (stats) CK(cudaMemcpyAsync(stats, c->d_stats, (size_t)nprob * sizeof(cfear_reg_stats), cudaMemcpyDeviceToHost, c->stream));
  if (npts_out) CK(cudaMemcpyAsync(npts_out, c->d_npts, (size_t)nprob * sizeof(int32_t), cudaMemcpyDeviceToHost, c->stream));
  CK(cudaEventRecord(c->ticket_ev[ticket], c->stream));
  return CFEAR_OK;
}

int cfear_odometry_step_batch_wait(cfear_ctx* c, int32_t ticket) {
  ENTER_NOJOIN(c);
  if (ticket < 0 || ticket >= (int)c->ticket_ev.size()) { g_err = "unknown ticket"; return CFEAR_ERR_ARG; }
  CK(cudaEventSynchronize(c->ticket_ev[ticket]));
  return CFEAR_OK;
}

int cfear_odometry_step_batch(cfear_ctx* c, int nprob, const uint8_t* polar, const double* mot,
                              const int32_t* kf_slots, int K, const int32_t* cur_slots,
                              double* poses, double* cov36, cfear_reg_stats* stats,
                              int32_t* npts_out, int32_t* ncells_out) {
  int32_t ticket = -1;
  RC(cfear_odometry_step_batch_submit(c, nprob, polar, mot, kf_slots, K, cur_slots, poses, cov36, stats, npts_out, &ticket));
  RC(cfear_odometry_step_batch_wait(c, ticket));
  if (nprob > 0 && ncells_out) RC(cfear_last_counts(c, nprob, cur_slots, nullptr, ncells_out));
  return CFEAR_OK;
}

int cfear_stage_timing(cfear_ctx* c, int enable, float ms_out[3]) {
  ENTER(c);
  int steps = 0;
  if (ms_out) {
    ms_out[0] = ms_out[1] = ms_out[2] = 0.f;
    if (c->timing && c->ev_used > 0) {
      CK(cudaStreamSynchronize(c->stream));
      for (int s = 0; s < c->ev_used; ++s)
        for (int i = 0; i < 3; ++i) {
          float ms = 0.f;
          if (cudaEventElapsedTime(&ms, c->ev[4 * s + i], c->ev[4 * s + i + 1]) != cudaSuccess) { ms = 0.f; (void)cudaGetLastError(); }
          ms_out[i] += ms;
        }
      steps = c->ev_used;
    }
  }
  c->ev_used = 0;
  c->timing = enable;
  return steps;       /* >= 0: number of timed steps summed into ms_out */
}

int cfear_last_counts(cfear_ctx* c, int nprob, const int32_t* cur_slots, int32_t* npts_out, int32_t* ncells_out) {
  ENTER(c);
  if (nprob < 0 || nprob > c->cfg.max_batch) { g_err = "nprob exceeds max_batch"; return CFEAR_ERR_CAPACITY; }
  if (npts_out) CK(cudaMemcpyAsync(npts_out, c->pb[c->last_pipe].d_npts, (size_t)nprob * sizeof(int32_t), cudaMemcpyDeviceToHost, c->stream));
  if (ncells_out) {
    if (!cur_slots) { g_err = "null slots"; return CFEAR_ERR_ARG; }
    std::vector<int32_t> all(c->cfg.max_cellsets);
    CK(cudaMemcpyAsync(all.data(), c->pool.ncells, all.size() * sizeof(int32_t), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    for (int i = 0; i < nprob; ++i) { RC(check_slot(c, cur_slots[i])); ncells_out[i] = all[cur_slots[i]]; }
  }
  CK(cudaStreamSynchronize(c->stream));
  return CFEAR_OK;
}


// ---- CA-CFAR (alternative filter) ---------------------------------------------------------------------------------
int cfear_cfar_filter(cfear_ctx* c, const uint8_t* polar, int nscans, const cfear_cfar_params* cp, cfear_point* cloud_out,
                      int capacity_per_scan, int32_t* npts_out) {
  ENTER(c);
  if (!polar || !cp || !npts_out || nscans < 0 || capacity_per_scan < 0) { g_err = "bad argument"; return CFEAR_ERR_ARG; }
  if (nscans > c->cfg.max_batch) { g_err = "nscans exceeds max_batch"; return CFEAR_ERR_CAPACITY; }
  if (cp->window_size < 1 || cp->nb_guard_cells < 0 || !(cp->false_alarm_rate > 0.0)) { g_err = "bad CFAR parameters"; return CFEAR_ERR_ARG; }
  if (nscans == 0) return CFEAR_OK;
  const int A = c->cfg.azimuths, R = c->cfg.range_bins;
  const size_t rows = (size_t)nscans * A;
  TempBufs tmp(c->stream);
  int32_t *d_cnt = nullptr, *d_off = nullptr, *d_n = nullptr; float4* d_out = nullptr;
  CK(tmp.get(&d_cnt, rows * 4));
  CK(tmp.get(&d_off, rows * 4));
  CK(tmp.get(&d_n, (size_t)nscans * 4));
  CK(tmp.get(&d_out, (size_t)nscans * capacity_per_scan * sizeof(float4)));
  c->polar_dirty = true;
  CK(cudaMemcpyAsync(c->d_polar, polar, rows * R, cudaMemcpyHostToDevice, c->stream));
  CfarParams p;
  p.polar = c->d_polar; p.nrows = (int)rows; p.A = A; p.R = R; p.window = cp->window_size; p.guard = cp->nb_guard_cells;
  const double N = (double)(2 * cp->window_size);
  p.scaling = N * (pow(cp->false_alarm_rate, -1. / N) - 1.);                 // cfar.cpp:12-16, 31
  p.range_res = (double)c->cfg.range_res; p.static_threshold = (double)c->cfg.z_min;   // radar_driver.cpp:54 passes the float parameters
  p.min_distance = (double)c->cfg.min_distance; p.max_distance = cp->max_distance;
  p.cs = c->d_cs; p.rowcnt = d_cnt; p.rowoff = d_off; p.cloud = d_out; p.cap = capacity_per_scan; p.pass = 0;
  const size_t smem = (size_t)(R + 1) * 4;
  CK(cudaFuncSetAttribute(k7_cfar, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  k7_cfar<<<(int)rows, K7_THREADS, smem, c->stream>>>(p);
  k7_offsets<<<nscans, 512, A * sizeof(int), c->stream>>>(d_cnt, A, d_off, d_n);
  p.pass = 1;
  k7_cfar<<<(int)rows, K7_THREADS, smem, c->stream>>>(p);
  c->launches += 3;
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(npts_out, d_n, (size_t)nscans * 4, cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  int rc = CFEAR_OK;
  for (int i = 0; i < nscans; ++i)
    if (npts_out[i] > capacity_per_scan) { g_err = "CFAR cloud exceeds capacity_per_scan (npts_out holds the required sizes)"; rc = CFEAR_ERR_CAPACITY; }
  if (rc == CFEAR_OK && cloud_out)
    for (int i = 0; i < nscans; ++i)
      CK(cudaMemcpyAsync(cloud_out + (size_t)i * capacity_per_scan, d_out + (size_t)i * capacity_per_scan,
                         (size_t)npts_out[i] * sizeof(float4), cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  return rc;
}

// ---- lock-step replay of many independent sequences (OdometryKeyframeFuser semantics on the device) ---------------
struct cfear_seq {
  cfear_ctx* ctx;
  SeqParams p;
  int slot_base;
  int parity = 0;                                   // K1 output set of the next step
  cudaEvent_t k1_done[2] = {nullptr, nullptr}, k3_done[2] = {nullptr, nullptr};
  std::vector<void*> allocs;
};

void cfear_seq_destroy(cfear_seq* s) {
  if (!s) return;
  cfear_sync(s->ctx);
  for (void* q : s->allocs) cudaFree(q);
  for (int i = 0; i < 2; ++i) { if (s->k1_done[i]) cudaEventDestroy(s->k1_done[i]); if (s->k3_done[i]) cudaEventDestroy(s->k3_done[i]); }
  delete s;
}

int cfear_seq_create(cfear_ctx* c, int nseq, int slot_base, int max_steps, const cfear_seq_params* sp, cfear_seq** out) {
  ENTER(c);
  if (!sp || !out || nseq < 1 || max_steps < 1) { g_err = "bad argument"; return CFEAR_ERR_ARG; }
  *out = nullptr;
  if (nseq > c->cfg.max_batch) { g_err = "nseq exceeds max_batch"; return CFEAR_ERR_CAPACITY; }
  if (sp->submap_scan_size < 1 || sp->submap_scan_size > c->cfg.max_keyframes) { g_err = "submap_scan_size must be in [1, max_keyframes]"; return CFEAR_ERR_ARG; }
  const int kmax = c->cfg.max_keyframes;
  if (slot_base < 0 || slot_base + nseq * (kmax + 1) > c->cfg.max_cellsets) { g_err = "sequences need nseq*(max_keyframes+1) cell-set slots from slot_base"; return CFEAR_ERR_CAPACITY; }
  RC(ensure_pipes(c));                      // streams / scratch sets of the replay are created here, not in the first step
  cfear_seq* s = new (std::nothrow) cfear_seq();
  if (!s) { g_err = "out of host memory"; return CFEAR_ERR_ARG; }
  s->ctx = c; s->slot_base = slot_base;
  SeqParams& P = s->p;
  P.nseq = nseq; P.submap = sp->submap_scan_size; P.kmax = kmax; P.use_guess = sp->use_guess; P.use_keyframe = sp->use_keyframe;
  P.max_steps = max_steps; P.min_keyframe_dist = sp->min_keyframe_dist; P.min_keyframe_rot_deg = sp->min_keyframe_rot_deg;
  auto al = [&](void** q, size_t bytes) { if (cudaMalloc(q, bytes) != cudaSuccess) return false; s->allocs.push_back(*q); return true; };
  bool ok = al((void**)&P.state, sizeof(SeqState) * nseq) && al((void**)&P.kf_pose, sizeof(T2) * (size_t)nseq * kmax) &&
            al((void**)&P.kf_slot, 4 * (size_t)nseq * kmax) && al((void**)&P.nscans_pp, 4 * (size_t)nseq) &&
            al((void**)&P.traj, 24 * (size_t)nseq * max_steps) && al((void**)&P.kf_flag, 4 * (size_t)nseq * max_steps) &&
            al((void**)&P.traj_stats, sizeof(cfear_reg_stats) * (size_t)nseq * max_steps);
  if (!ok) { g_err = "cudaMalloc failed"; cfear_seq_destroy(s); return CFEAR_ERR_CUDA; }
  P.mot = c->d_mot; P.cur_slots = c->d_curslots; P.slots = c->d_slots; P.poses = c->d_poses; P.stats = c->d_stats;
  for (int i = 0; i < 2; ++i) {
    if (cudaEventCreateWithFlags(&s->k1_done[i], cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&s->k3_done[i], cudaEventDisableTiming) != cudaSuccess) { g_err = "cudaEventCreate failed"; cfear_seq_destroy(s); return CFEAR_ERR_CUDA; }
  }
  k6_init<<<(nseq + 127) / 128, 128, 0, c->stream>>>(P, slot_base);
  c->launches++;
  CK(cudaGetLastError());
  CK(cudaMemsetAsync(P.traj, 0, 24 * (size_t)nseq * max_steps, c->stream));
  CK(cudaMemsetAsync(P.kf_flag, 0, 4 * (size_t)nseq * max_steps, c->stream));
  CK(cudaMemsetAsync(P.traj_stats, 0, sizeof(cfear_reg_stats) * (size_t)nseq * max_steps, c->stream));
  *out = s;
  return CFEAR_OK;
}

// One time step of every sequence.  The filter of scan t+1 needs nothing from scan t (src/offline_odometry.cpp:103: the
// CallbackOffline of the next frame), only Compensate and the registration do (previous motion, keyframe window), so K1
// runs on its own stream, one step ahead, into alternating output sets, and hides under the registration of the previous
// scan:   filter stream:  [H2D] K1(t) | [H2D] K1(t+1) ...        main stream:  k6_pre(t) K3(t) K5(t) k6_post(t) | ...
static int seq_step_common(cfear_seq* s, const uint8_t* polar, bool host) {
  cfear_ctx* c = s->ctx;
  const SeqParams& P = s->p;
  const int B = P.nseq;
  RC(ensure_pipes(c));
  RC(set_pipe_user(c, 2));
  PipeBufs& F = c->pb[1];                              // filter stream
  PipeBufs& M = c->pb[2];                              // everything that depends on the previous scan
  const int par = s->parity; s->parity ^= 1;
  PipeBufs W = c->pb[1 + par];                         // K1 -> K3 -> K5 scratch of this step
  CK(cudaEventRecord(F.in, c->stream));                // inputs are ordered after the context stream's work so far
  CK(cudaStreamWaitEvent(F.stream, F.in, 0));
  CK(cudaStreamWaitEvent(M.stream, F.in, 0));
  CK(cudaStreamWaitEvent(F.stream, s->k3_done[par], 0));   // this output set was last read by K3 two steps ago
  const uint8_t* d_polar = polar;
  if (host) {
    c->polar_dirty = true;
    CK(cudaMemcpyAsync(c->d_polar, polar, (size_t)B * c->cfg.azimuths * c->cfg.range_bins, cudaMemcpyHostToDevice, F.stream));
    d_polar = c->d_polar;
  }
  c->last_pipe = 1 + par;
  W.stream = F.stream;
  RC(launch_k1(c, W, d_polar, B));
  CK(cudaEventRecord(s->k1_done[par], F.stream));
  k6_pre<<<(B + 127) / 128, 128, 0, M.stream>>>(P);
  c->launches++;
  CK(cudaGetLastError());
  CK(cudaStreamWaitEvent(M.stream, s->k1_done[par], 0));
  W.stream = M.stream;
  RC(launch_k3(c, W, 0, B, c->d_mot, c->d_curslots, false));
  CK(cudaEventRecord(s->k3_done[par], M.stream));
  RC(launch_k5(c, W, B, P.kmax + 1, c->d_slots, c->d_poses, c->d_cov36, c->d_stats, nullptr, 0, P.nscans_pp));
  k6_post<<<(B + 127) / 128, 128, 0, M.stream>>>(P);
  c->launches++;
  CK(cudaGetLastError());
  CK(cudaEventRecord(F.done, F.stream));
  CK(cudaEventRecord(M.done, M.stream));
  c->inflight = true;
  return CFEAR_OK;
}

int cfear_seq_step_dev(cfear_seq* s, const uint8_t* d_polar) {
  if (!s || !d_polar) { g_err = "null argument"; return CFEAR_ERR_ARG; }
  ENTER_NOJOIN(s->ctx);
  return seq_step_common(s, d_polar, false);
}

int cfear_seq_step(cfear_seq* s, const uint8_t* polar) {
  if (!s || !polar) { g_err = "null argument"; return CFEAR_ERR_ARG; }
  ENTER_NOJOIN(s->ctx);
  return seq_step_common(s, polar, true);
}

int cfear_seq_read(cfear_seq* s, int step_from, int nsteps, double* poses_out, int32_t* keyframe_out, cfear_reg_stats* stats_out) {
  if (!s) { g_err = "null argument"; return CFEAR_ERR_ARG; }
  cfear_ctx* c = s->ctx;
  ENTER(c);
  const SeqParams& P = s->p;
  if (step_from < 0 || nsteps < 0 || step_from + nsteps > P.max_steps) { g_err = "step range outside [0, max_steps)"; return CFEAR_ERR_ARG; }
  for (int b = 0; b < P.nseq; ++b) {
    const size_t src = (size_t)b * P.max_steps + step_from, dst = (size_t)b * nsteps;
    if (poses_out) CK(cudaMemcpyAsync(poses_out + dst * 3, P.traj + src * 3, 24 * (size_t)nsteps, cudaMemcpyDeviceToHost, c->stream));
    if (keyframe_out) CK(cudaMemcpyAsync(keyframe_out + dst, P.kf_flag + src, 4 * (size_t)nsteps, cudaMemcpyDeviceToHost, c->stream));
    if (stats_out) CK(cudaMemcpyAsync(stats_out + dst, reinterpret_cast<cfear_reg_stats*>(P.traj_stats) + src, sizeof(cfear_reg_stats) * (size_t)nsteps, cudaMemcpyDeviceToHost, c->stream));
  }
  CK(cudaStreamSynchronize(c->stream));
  return CFEAR_OK;
}

}  // extern "C"
