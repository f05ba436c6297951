/*
 * cfear_b200.h -- C ABI of the B200-native CFEAR per-scan hot path.
 *
 * Drop-in boundary for the reference's radar_driver / MapPointNormal /
 * n_scan_normal_reg path (dan11003/CFEAR_Radarodometry_code_public).  Plain C:
 * POD structs, caller-allocated outputs, int status (0 = ok, <0 = error, see
 * cfear_last_error).  No ROS / PCL / Eigen / Ceres / torch types.
 * Citations are file:line into the reference tree.
 *
 * Each entry point names the reference interface it replaces:
 *
 *   cfear_kstrongest / cfear_filter     StructuredKStrongest::StructuredKStrongest + FilterKstrongest
 *                                       (src/cfear_radarodometry/radar_filters.cpp:198-237),
 *                                       getPeaksFilteredPointCloud(cloud,false) (:300-337),
 *                                       as called by radarDriver::Process (radar_driver.cpp:48-61)
 *                                       behind radarDriver::CallbackOffline (radar_driver.h:90).
 *   cfear_compensate                    Compensate(cloud, Tmotion, ccw) (utils.cpp:96-113, utils.h:28-32)
 *   cfear_surface_points                MapPointNormal::MapPointNormal (pointnormal.h:118, pointnormal.cpp:65-90,
 *                                       ComputeNormals :265-297, cell::cell :7-63, ComputeSearchTreeFromCells :151-162)
 *   cfear_cells_download / _upload      MapPointNormal accessors GetSize/GetMean2d/GetNormal2d/GetCov2d/
 *                                       GetCell(i).Nsamples_/GetPlanarity (pointnormal.h:124-172)
 *   cfear_nearest                       MapPointNormal::GetClosestIdx (pointnormal.cpp:238-254)
 *   cfear_register / _batch             n_scan_normal_reg::Register (n_scan_normal.h:37, n_scan_normal.cpp:82-187)
 *                                       incl. BuildOptimizationProblem/AddScanPairCost (:215-391),
 *                                       SolveOptimizationProblem (:443-452), GetCovariance (:392-433)
 *   cfear_get_cost_batch                n_scan_normal_reg::GetCost (n_scan_normal.h:41, n_scan_normal.cpp:187-213)
 *   cfear_odometry_step_batch[_dev]     one radarReader loop body (src/offline_odometry.cpp:103-108) for many
 *                                       independent scans: CallbackOffline -> Compensate -> MapPointNormal -> Register
 */
#ifndef CFEAR_B200_H_
#define CFEAR_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* cost_metric, registration.h:55 */
enum { CFEAR_COST_P2P = 0, CFEAR_COST_P2L = 1, CFEAR_COST_P2D = 2 };
/* loss_type, registration.h:60 */
enum { CFEAR_LOSS_NONE = 0, CFEAR_LOSS_HUBER = 1, CFEAR_LOSS_CAUCHY = 2, CFEAR_LOSS_SOFTLONE = 3,
       CFEAR_LOSS_COMBINED = 4, CFEAR_LOSS_TUKEY = 5 };
/* weightoption, registration.h:50 */
enum { CFEAR_WEIGHT_UNIFORM = 0, CFEAR_WEIGHT_SIM_N = 1, CFEAR_WEIGHT_SIM_DIRECTION = 2,
       CFEAR_WEIGHT_SIM_SCALE = 3, CFEAR_WEIGHT_COMBINED = 4 };
/* inner solver: faithful Ceres trust-region LM loop, or fixed-count Gauss-Newton/IRLS */
enum { CFEAR_SOLVER_CERES_LM = 0, CFEAR_SOLVER_GN_FIXED = 1, CFEAR_SOLVER_COST_ONLY = 2 /* internal: cfear_get_cost_batch */ };

/* status codes */
enum { CFEAR_OK = 0, CFEAR_ERR_ARG = -1, CFEAR_ERR_CUDA = -2, CFEAR_ERR_CAPACITY = -3, CFEAR_ERR_NO_DEVICE = -4 };

/* Parameters: radarDriver::Parameters (radar_driver.h:35-48), OdometryKeyframeFuser::Parameters
 * (odometrykeyframefuser.h), n_scan_normal_reg ctor + SetD2dPar (n_scan_normal.h:35,53,72-75),
 * Registration::radius_ (registration.h:122). */
typedef struct cfear_config {
  int32_t device;            /* CUDA device ordinal */
  int32_t max_batch;         /* max scans / problems per call */
  int32_t azimuths;          /* A: image rows (400) */
  int32_t range_bins;        /* R: image cols (3360) */
  int32_t k_strongest;       /* 12 */
  float   z_min;             /* 60; truncated to int then uchar like radar_driver.cpp:58 / radar_filters.cpp:212 */
  float   range_res;         /* 0.0438 (float, widened to double like the reference) */
  float   min_distance;      /* 2.5 */
  float   radius;            /* "res": surface point radius / voxel leaf (3.5) */
  double  downsample_factor; /* MapPointNormal::downsample_factor (1) */
  int32_t weight_intensity;  /* 1 */
  int32_t compensate;        /* 1 */
  int32_t radar_ccw;         /* 0 */
  int32_t cost;              /* CFEAR_COST_* */
  int32_t loss;              /* CFEAR_LOSS_* */
  int32_t weight_opt;        /* CFEAR_WEIGHT_* */
  int32_t solver_mode;       /* CFEAR_SOLVER_* */
  double  loss_limit;        /* 0.1 */
  double  cov_scale;         /* 1 */
  double  regularization;    /* 1 */
  double  reg_radius;        /* 2.0 (association radius; doubled on outer iteration 1) */
  int32_t max_outer;         /* 8 */
  int32_t min_outer;         /* 3 */
  int32_t max_inner;         /* 20 */
  int32_t gn_iters;          /* 10 (CFEAR_SOLVER_GN_FIXED only) */
  int32_t max_keyframes;     /* K max = submap_scan_size */
  int32_t max_cellsets;      /* number of device-resident cell-set slots */
  int32_t max_cells;         /* capacity of one cell set; 0 -> azimuths*k_strongest */
  int32_t steps_in_flight;   /* cfear_odometry_step_batch_dev_submit: internal streams / scratch sets to rotate through (0 -> 5, max 8) */
} cfear_config;

/* pcl::PointXYZI as written by getPeaksFilteredPointCloud (radar_filters.cpp:328-333) */
typedef struct cfear_point { float x, y, z, intensity; } cfear_point;

/* One oriented surface point: the pose-relevant members of class cell (pointnormal.h:66-73) */
typedef struct cfear_cell {
  double mean[2];        /* u_ */
  double normal[2];      /* snormal_ (flipped toward origin) */
  double cov[4];         /* cov_, row-major */
  double planarity;      /* scale_ = log(1 + cond/2) */
  double avg_intensity;  /* avg_intensity_ */
  int32_t nsamples;      /* Nsamples_ */
  int32_t pad;
} cfear_cell;

/* What Register() leaves behind: return value, itr_ (registration.h:107), ceres summary bits, score_ */
typedef struct cfear_reg_stats {
  int32_t success;           /* Register() return */
  int32_t outer_iterations;  /* itr_ after the loop (logged as "itrs", n_scan_normal.cpp:161) */
  int32_t inner_iterations;  /* sum over outer its of summary_.iterations.size()-1 */
  int32_t num_residuals;     /* scalar residuals of the last problem */
  int32_t num_blocks;        /* residual blocks of the last problem */
  int32_t usable;            /* summary_.IsSolutionUsable() of the last solve */
  double  final_cost;        /* summary_.final_cost */
  double  score;             /* score_ = final_cost / num_residuals */
  int32_t pose_written;      /* 1 if some solve of this Register() call was usable, i.e. Tsrc was rewritten from the parameters
                                (n_scan_normal.cpp:119-121); 0: the pose handed in is returned untouched */
  int32_t reserved;
} cfear_reg_stats;

typedef struct cfear_ctx cfear_ctx;

void cfear_default_config(cfear_config* cfg);
int  cfear_create(const cfear_config* cfg, cfear_ctx** out);
void cfear_destroy(cfear_ctx* ctx);
/* Change the non-structural parameters (filter thresholds, radius, weights, cost/loss, solver limits) of a live
 * context; device, max_batch, azimuths, range_bins, k_strongest and the max_* capacities must equal the create-time
 * values.  This is what lets one context serve MapPointNormal(radius, weight_intensity) / n_scan_normal_reg(cost, loss,
 * ...) objects constructed with different arguments, like the reference. */
int  cfear_update_config(cfear_ctx* ctx, const cfear_config* cfg);
const char* cfear_last_error(void);
const char* cfear_version(void);
/* number of kernel launches issued by this context so far (bench.py's gpu_launches) */
int64_t cfear_launch_count(const cfear_ctx* ctx);

/* ---- filter ---------------------------------------------------------------------------------------------------- */
/* polar: nscans images, A x R uint8 row-major (rows = azimuth), HOST memory.
 * idx_out [nscans][A][k] range bins ascending by (intensity, range), -1 padded; cnt_out [nscans][A]. */
int cfear_kstrongest(cfear_ctx* ctx, const uint8_t* polar, int nscans, int32_t* idx_out, int32_t* cnt_out);
/* Same plus the filtered cloud (min-range cut, polar->Cartesian; NOT motion compensated, like CallbackOffline).
 * cloud_out [nscans][A*k] points, npts_out [nscans]; idx_out/cnt_out may be NULL.
 * peaks_out/npeaks_out (may be NULL): the AxialNonMaxSupress cloud (radar_filters.cpp:238-298). */
int cfear_filter(cfear_ctx* ctx, const uint8_t* polar, int nscans, int32_t* idx_out, int32_t* cnt_out,
                 cfear_point* cloud_out, int32_t* npts_out, cfear_point* peaks_out, int32_t* npeaks_out);
/* In-place Compensate on a host cloud. mot = (x, y, yaw) of Tmotion. */
int cfear_compensate(cfear_ctx* ctx, cfear_point* cloud, int n, const double mot[3], int ccw);

/* ---- surface points -------------------------------------------------------------------------------------------- */
/* Build the cell set of one (already compensated) host cloud into device slot `slot`; returns ncells via *ncells. */
int cfear_surface_points(cfear_ctx* ctx, const cfear_point* cloud, int n, int slot, int32_t* ncells);
/* radarDriver::CallbackOffline -> Compensate -> MapPointNormal for nscans HOST images at once: cell set i lands in
 * slots[i] (mot [nscans][3] may be NULL = no compensation).  This is how keyframe sets become device resident. */
int cfear_scans_to_cells_batch(cfear_ctx* ctx, int nscans, const uint8_t* polar, const double* mot,
                               const int32_t* slots, int32_t* npts_out, int32_t* ncells_out);
int cfear_cells_count(cfear_ctx* ctx, int slot, int32_t* ncells);
int cfear_cells_download(cfear_ctx* ctx, int slot, cfear_cell* out, int capacity, int32_t* ncells);
/* Upload an arbitrary cell set (also builds its nearest-neighbour index). */
int cfear_cells_upload(cfear_ctx* ctx, int slot, const cfear_cell* cells, int n);
/* GetClosestIdx for nq query points (x,y doubles) against slot; out_idx[q] = cell index or -1. */
int cfear_nearest(cfear_ctx* ctx, int slot, const double* queries_xy, int nq, double radius, int32_t* out_idx);

/* ---- registration ---------------------------------------------------------------------------------------------- */
/* slots[nscans]: cell-set slots, last = current scan (free block), others = keyframes (fixed).
 * poses [nscans][3] (x, y, yaw) in/out (only the last changes); cov36: 6x6 row-major reg_cov.back();
 * returns CFEAR_OK even if Register() itself returned false (see stats->success). */
int cfear_register(cfear_ctx* ctx, const int32_t* slots, int nscans, double* poses, double* cov36,
                   cfear_reg_stats* stats);
/* nprob independent problems, each with nscans sets: slots [nprob][nscans], poses [nprob][nscans][3],
 * cov36 [nprob][36], stats [nprob]. assoc_out (may be NULL): [nprob][nscans-1][max_cells] target index of the
 * last outer iteration's association per (keyframe, src cell), -1 = none (scan_associations_, registration.h:105). */
int cfear_register_batch(cfear_ctx* ctx, int nprob, const int32_t* slots, int nscans, double* poses,
                         double* cov36, cfear_reg_stats* stats, int32_t* assoc_out);

/* The same with the rest of Register()'s observable state:
 *   assoc_sim_out (may be NULL)  [nprob][nscans-1][max_cells]: the direction similarity max(n_src' . n_tar, 0) of each
 *       associated pair -- with assoc_out, the cell sets' Nsamples_ / planarity this is the content of the public members
 *       scan_associations_ / weight_associations_ (registration.h:103-106, n_scan_normal.cpp:255-256);
 *   prior_sqrt_info (may be NULL) [nprob][9]: row-major 3x3 lower-triangular factor L of the guess prior, Register(...,
 *       soft_constraints = true): L = Cov6to3(reg_cov.back()).inverse().llt().matrixL() (n_scan_normal.cpp:373-377); the block
 *       r = sqrt(#source cells) L (guess - x) (mahalanobisDistanceError, n_scan_normal.h:259-290) joins every outer iteration's
 *       problem, guess = the pose handed in for the last scan. */
int cfear_register_batch_ex(cfear_ctx* ctx, int nprob, const int32_t* slots, int nscans, double* poses,
                            double* cov36, cfear_reg_stats* stats, int32_t* assoc_out, double* assoc_sim_out,
                            const double* prior_sqrt_info);

/* n_scan_normal_reg::GetCost (n_scan_normal.h:41, n_scan_normal.cpp:187-213) for nprob independent problems: associate
 * once at the poses given (all fixed, the last set is the source) with the registration radius, return the robustified
 * cost 1/2 sum w rho(|r|^2) (ceres::Problem::Evaluate with default options), the number of scalar residuals and
 * ok = 0 where GetCost returns false (<= 1 residual).  slots [nprob][nscans], poses [nprob][nscans][3].
 * num_residuals_out / ok_out may be NULL.  This is what approximateCovarianceBySampling
 * (odometrykeyframefuser.cpp:261-380) calls once per pose sample; here all samples go in one launch. */
int cfear_get_cost_batch(cfear_ctx* ctx, int nprob, const int32_t* slots, int nscans, const double* poses,
                         double* cost_out, int32_t* num_residuals_out, int32_t* ok_out);

/* ---- whole per-scan path, many independent scans ----------------------------------------------------------------- */
/* For b in [0,nprob): polar[b] -> k-strongest -> cloud -> Compensate(mot[b]) -> surface points into cur_slots[b]
 * -> Register against kf_slots[b][0..K-1] with poses[b][0..K-1] fixed and poses[b][K] the guess (in/out).
 * HOST buffers; host<->device copies are part of the call. ncells_out/npts_out may be NULL. */
int cfear_odometry_step_batch(cfear_ctx* ctx, int nprob, const uint8_t* polar, const double* mot,
                              const int32_t* kf_slots, int K, const int32_t* cur_slots,
                              double* poses, double* cov36, cfear_reg_stats* stats,
                              int32_t* npts_out, int32_t* ncells_out);
/* The same step split in two so that consecutive steps pipeline: _submit enqueues the host->device copies, the kernels
 * and the device->host copies of the results and returns at once with a ticket; _wait blocks until that step's results
 * are in the caller's buffers.  Up to 8 steps may be in flight; every buffer passed to _submit (inputs and outputs,
 * ideally pinned) must stay valid and untouched until its _wait.  Submitting step i+1 before waiting for step i keeps
 * the PCIe link busy during the registration tail of step i.  cfear_odometry_step_batch == _submit + _wait. */
int cfear_odometry_step_batch_submit(cfear_ctx* ctx, int nprob, const uint8_t* polar, const double* mot,
                                     const int32_t* kf_slots, int K, const int32_t* cur_slots,
                                     double* poses, double* cov36, cfear_reg_stats* stats,
                                     int32_t* npts_out, int32_t* ticket_out);
int cfear_odometry_step_batch_wait(cfear_ctx* ctx, int32_t ticket);
/* Same with every buffer already resident on the device (d_ prefix = device pointer), asynchronous on the
 * context's stream (cfear_sync to wait).  d_poses [nprob][K+1][3] in/out, d_cov36 [nprob][36] (reg_cov.back(),
 * GetCovariance layout), d_stats [nprob]. */
int cfear_odometry_step_batch_dev(cfear_ctx* ctx, int nprob, const uint8_t* d_polar, const double* d_mot,
                                  const int32_t* d_kf_slots, int K, const int32_t* d_cur_slots,
                                  double* d_poses, double* d_cov36, cfear_reg_stats* d_stats);
/* The same step, overlapped with its neighbours: step i runs on internal stream i mod cfear_config.steps_in_flight with its
 * own scratch, so that the
 * filter / surface-point kernels of step i+1 execute while the registration of step i is in its tail (the radarReader loop
 * of src/offline_odometry.cpp:103-108 has no dependency from CallbackOffline of frame t+1 on the pose of frame t; for
 * independent batches neither has the MapPointNormal build).  Ordering contract:
 *   - inputs must be complete with respect to the context stream (cfear_stream) at the time of the call;
 *   - every buffer of the step (d_polar .. d_stats) and the cell-set slots it names belong to the step until it is
 *     complete: cfear_odometry_step_batch_wait(ticket) on the host, cfear_stream_wait_ticket(ticket) on the context
 *     stream, or cfear_join / cfear_sync / any other entry point (all of which join first);
 *   - d_cur_slots of a step must not name a slot (keyframe or current) of a step still in flight: rotate as many sets
 *     as steps are kept in flight.
 * cfear_odometry_step_batch_dev == this + cfear_join, i.e. fully ordered on the context stream. */
int cfear_odometry_step_batch_dev_submit(cfear_ctx* ctx, int nprob, const uint8_t* d_polar, const double* d_mot,
                                         const int32_t* d_kf_slots, int K, const int32_t* d_cur_slots,
                                         double* d_poses, double* d_cov36, cfear_reg_stats* d_stats, int32_t* ticket_out);
/* Makes the context stream wait (on the device; the host does not block) for one step / for every step in flight. */
int cfear_stream_wait_ticket(cfear_ctx* ctx, int32_t ticket);
int cfear_join(cfear_ctx* ctx);
int cfear_sync(cfear_ctx* ctx);
/* The CUDA stream (cudaStream_t as void*) the context launches on, for event timing by the caller. */
void* cfear_stream(cfear_ctx* ctx);
/* Device time (ms, CUDA events on the context stream) per stage, SUMMED over every cfear_odometry_step_batch[_dev]
 * call since the previous cfear_stage_timing call: [0] k-strongest (+ its H2D chunks on the host-buffer path),
 * [1] surface points, [2] registration.  Returns the number of steps summed (>= 0) or a negative status.
 * `enable` switches recording for the following steps (it adds 4 event records per step). */
int cfear_stage_timing(cfear_ctx* ctx, int enable, float ms_out[3]);
/* Per-scan counts of the most recent step (device->host copy): npts/ncells of cur_slots order. */
int cfear_last_counts(cfear_ctx* ctx, int nprob, const int32_t* cur_slots, int32_t* npts_out, int32_t* ncells_out);


/* ---- CA-CFAR (alternative filter, --filter-type CA-CFAR) ------------------------------------------------------------ */
/* AzimuthCACFAR::getFilteredPointCloud (src/cfear_radarodometry/cfar.cpp:35-83) as dispatched by radarDriver::Process
 * (radar_driver.cpp:52-56): window / guard cells / false-alarm rate from radarDriver::Parameters (radar_driver.h:43-44);
 * static threshold = z_min, range_res and min_distance come from the context configuration; max_distance is 400 in the
 * reference's call.  cloud_out [nscans][capacity_per_scan] in (azimuth, bin) order, npts_out [nscans].  Returns
 * CFEAR_ERR_CAPACITY (npts_out filled with the needed sizes) if a cloud does not fit. */
typedef struct cfear_cfar_params {
  int32_t window_size;        /* 10 */
  int32_t nb_guard_cells;     /* 20 */
  double  false_alarm_rate;   /* 0.01 */
  double  max_distance;       /* 400.0 */
} cfear_cfar_params;
int cfear_cfar_filter(cfear_ctx* ctx, const uint8_t* polar, int nscans, const cfear_cfar_params* params,
                      cfear_point* cloud_out, int capacity_per_scan, int32_t* npts_out);

/* ---- lock-step replay of many independent sequences ------------------------------------------------------------- */
/* OdometryKeyframeFuser::processFrame (odometrykeyframefuser.cpp:143-259) for nseq sequences advancing together:
 * Compensate with the previous motion, constant-velocity guess, Register against the sliding window of keyframes,
 * AccelerationVelocitySanityCheck (:76-94), KeyFrameBasedFuse (:62-73), AddToReference (:470-476).  Poses, motion, the
 * keyframe window (a ring of cell-set slots) and the trajectory table stay on the device; a step is enqueued with no
 * host synchronisation.  Parameters: OdometryKeyframeFuser::Parameters (odometrykeyframefuser.h:92-107). */
typedef struct cfear_seq_params {
  int32_t submap_scan_size;      /* 3 */
  int32_t use_guess;             /* 1 (forced true by offline_odometry.cpp:273) */
  int32_t use_keyframe;          /* 1 */
  int32_t reserved;
  double  min_keyframe_dist;     /* 1.5 */
  double  min_keyframe_rot_deg;  /* 5 */
} cfear_seq_params;
typedef struct cfear_seq cfear_seq;
/* Sequence b owns the cell-set slots [slot_base + b*(max_keyframes+1), +max_keyframes+1). */
int  cfear_seq_create(cfear_ctx* ctx, int nseq, int slot_base, int max_steps, const cfear_seq_params* params, cfear_seq** out);
void cfear_seq_destroy(cfear_seq* seq);
/* One time step for every sequence: polar [nseq][A][R] (HOST / DEVICE).  Asynchronous: the step is enqueued on the
 * library's internal streams (the filter of a scan runs one step ahead, under the registration of the previous scan) after
 * whatever the context stream holds at the time of the call; the image buffer must stay valid and untouched until the
 * step is complete (cfear_join / cfear_sync / cfear_seq_read or any other entry point). */
int  cfear_seq_step(cfear_seq* seq, const uint8_t* polar);
int  cfear_seq_step_dev(cfear_seq* seq, const uint8_t* d_polar);
/* Waits for the enqueued steps and copies out steps [step_from, step_from+nsteps): poses_out [nseq][nsteps][3],
 * keyframe_out [nseq][nsteps] (1 = scan became a keyframe), stats_out [nseq][nsteps] (any may be NULL). */
int  cfear_seq_read(cfear_seq* seq, int step_from, int nsteps, double* poses_out, int32_t* keyframe_out, cfear_reg_stats* stats_out);

/* ---- memory helpers for callers that keep buffers resident (bench / replay harness) ------------------------------- */
/* Page-locked host memory (fast host<->device copies for the host-buffer entry points). */
void* cfear_alloc_pinned(size_t bytes);
void  cfear_free_pinned(void* p);
/* Raw device memory on the context's device + blocking copies, for the *_dev entry points. */
void* cfear_alloc_device(cfear_ctx* ctx, size_t bytes);
void  cfear_free_device(cfear_ctx* ctx, void* p);
int   cfear_memcpy_h2d(cfear_ctx* ctx, void* dst_dev, const void* src_host, size_t bytes);
int   cfear_memcpy_d2h(cfear_ctx* ctx, void* dst_host, const void* src_dev, size_t bytes);

#ifdef __cplusplus
}
#endif
#endif /* CFEAR_B200_H_ */
