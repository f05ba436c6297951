// K6: OdometryKeyframeFuser::processFrame bookkeeping on the device, for B independent sequences advancing in
// lock-step (odometrykeyframefuser.cpp:143-259, 62-94, 470-494).  With this the whole per-scan loop of
// offline_odometry (src/offline_odometry.cpp:73-127) is enqueued per time step as
//   k6_pre -> K1 -> K3 -> K5 -> k6_post
// with no host round trip: poses, previous motion, the sliding keyframe window (a ring of cell-set slots) and the
// trajectory table live in HBM.
#pragma once
#include "common.cuh"

namespace cfear {

struct T2 { double r00, r01, r10, r11, x, y; };   // planar rigid transform as the 2x3 block of an Affine3d

__device__ __forceinline__ T2 t2_identity() { T2 t; t.r00 = 1; t.r01 = 0; t.r10 = 0; t.r11 = 1; t.x = 0; t.y = 0; return t; }
__device__ __forceinline__ T2 t2_from(double x, double y, double yaw) {
  T2 t; double s, c; sincos(yaw, &s, &c);
  t.r00 = c; t.r01 = -s; t.r10 = s; t.r11 = c; t.x = x; t.y = y; return t;
}
__device__ __forceinline__ T2 t2_mul(const T2& a, const T2& b) {
  T2 r;
  r.r00 = a.r00 * b.r00 + a.r01 * b.r10; r.r01 = a.r00 * b.r01 + a.r01 * b.r11;
  r.r10 = a.r10 * b.r00 + a.r11 * b.r10; r.r11 = a.r10 * b.r01 + a.r11 * b.r11;
  r.x = a.r00 * b.x + a.r01 * b.y + a.x; r.y = a.r10 * b.x + a.r11 * b.y + a.y;
  return r;
}
__device__ __forceinline__ T2 t2_inv(const T2& a) {
  T2 r; r.r00 = a.r00; r.r01 = a.r10; r.r10 = a.r01; r.r11 = a.r11;
  r.x = -(r.r00 * a.x + r.r01 * a.y); r.y = -(r.r10 * a.x + r.r11 * a.y);
  return r;
}
__device__ __forceinline__ double t2_yaw(const T2& a) { return atan2(a.r10, a.r11); }   // utils.cpp:115-122, planar

struct SeqState {                 // one sequence
  T2 T_prev, Tmot, Tguess;
  int nkf;                        // keyframes in the window (<= submap)
  int cur_slot;                   // slot the next scan's cell set is written to
  int step;                       // scans processed so far
  int pad;
};

struct SeqParams {
  int nseq, submap, kmax;         // kmax = row stride - 1 of the slot / pose tables
  int use_guess, use_keyframe, max_steps;
  double min_keyframe_dist, min_keyframe_rot_deg;
  SeqState* state;                // [nseq]
  T2* kf_pose;                    // [nseq][kmax]  window, oldest first
  int32_t* kf_slot;               // [nseq][kmax]
  // per-step tables consumed by K3 / K5
  double* mot;                    // [nseq][3]
  int32_t* cur_slots;             // [nseq]
  int32_t* slots;                 // [nseq][kmax+1]
  int32_t* nscans_pp;             // [nseq]
  double* poses;                  // [nseq][kmax+1][3]
  const void* stats;              // [nseq] cfear_reg_stats (K5 output)
  // trajectory
  double* traj;                   // [nseq][max_steps][3]
  int32_t* kf_flag;               // [nseq][max_steps]
  void* traj_stats;               // [nseq][max_steps] cfear_reg_stats
};

struct RegStatsK6 { int32_t success, outer_iterations, inner_iterations, num_residuals, num_blocks, usable; double final_cost, score; int32_t pose_written, reserved; };

// before the scan is processed: previous motion for Compensate, guess, registration tables
__global__ void k6_pre(const SeqParams P) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= P.nseq) return;
  SeqState S = P.state[b];
  const T2 TprevMot = S.Tmot;                                                   // :146
  P.mot[3 * b + 0] = TprevMot.x; P.mot[3 * b + 1] = TprevMot.y; P.mot[3 * b + 2] = t2_yaw(TprevMot);
  S.Tguess = P.use_guess ? t2_mul(S.T_prev, TprevMot) : S.T_prev;               // :164-168
  const int stride = P.kmax + 1;
  for (int i = 0; i < S.nkf; ++i) {                                             // FormatScans :478-494
    const T2 T = P.kf_pose[(size_t)b * P.kmax + i];
    P.slots[(size_t)b * stride + i] = P.kf_slot[(size_t)b * P.kmax + i];
    double* p = P.poses + ((size_t)b * stride + i) * 3;
    p[0] = T.x; p[1] = T.y; p[2] = t2_yaw(T);
  }
  P.slots[(size_t)b * stride + S.nkf] = S.cur_slot;
  double* p = P.poses + ((size_t)b * stride + S.nkf) * 3;
  p[0] = S.Tguess.x; p[1] = S.Tguess.y; p[2] = t2_yaw(S.Tguess);
  P.cur_slots[b] = S.cur_slot;
  P.nscans_pp[b] = S.nkf + 1;
  P.state[b].Tguess = S.Tguess;
}

// after registration: pose bookkeeping, sanity check, keyframe decision, window update, trajectory row
__global__ void k6_post(const SeqParams P) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= P.nseq) return;
  SeqState S = P.state[b];
  const int stride = P.kmax + 1;
  const RegStatsK6 st = reinterpret_cast<const RegStatsK6*>(P.stats)[b];
  T2 Tcurrent = t2_identity();
  int fuse = 0;
  if (S.nkf == 0) {                                                             // first scan: keyframe at identity (:171-177)
    P.kf_pose[(size_t)b * P.kmax] = t2_identity();
    P.kf_slot[(size_t)b * P.kmax] = S.cur_slot;
    S.nkf = 1; fuse = 1;
    S.cur_slot = S.cur_slot + 1;                                                // slots of a sequence are contiguous
  } else {
    // Tsrc is rewritten from the parameters after every usable solve (n_scan_normal.cpp:119-121, 177-178) and keeps the
    // last such pose if a later outer iteration fails; the fuser ignores Register()'s return value (:184-186)
    const bool wrote = st.pose_written != 0;
    const double* p = P.poses + ((size_t)b * stride + S.nkf) * 3;
    Tcurrent = wrote ? t2_from(p[0], p[1], p[2]) : S.Tguess;                    // :195
    const T2 Tmot_current = t2_mul(t2_inv(S.T_prev), Tcurrent);
    {                                                                           // AccelerationVelocitySanityCheck :76-94
      const double dt = 0.25;
      const double vel = sqrt(Tmot_current.x * Tmot_current.x + Tmot_current.y * Tmot_current.y) / dt;
      const double ax = (Tmot_current.x - S.Tmot.x) / (dt * dt), ay = (Tmot_current.y - S.Tmot.y) / (dt * dt);
      if (sqrt(ax * ax + ay * ay) > 200 || vel > 200) Tcurrent = S.Tguess;      // :197-199
    }
    S.Tmot = t2_mul(t2_inv(S.T_prev), Tcurrent);                                // :200
    const T2 last = P.kf_pose[(size_t)b * P.kmax + S.nkf - 1];
    const T2 Tkeydiff = t2_mul(t2_inv(last), Tcurrent);                         // :227
    fuse = !P.use_keyframe || sqrt(Tkeydiff.x * Tkeydiff.x + Tkeydiff.y * Tkeydiff.y) > P.min_keyframe_dist ||
           fabs(t2_yaw(Tkeydiff)) > P.min_keyframe_rot_deg * M_PI / 180.0;      // :62-73
    if (fuse) {                                                                 // AddToReference :470-476
      const int newslot = S.cur_slot;
      if (S.nkf < P.submap) {
        P.kf_pose[(size_t)b * P.kmax + S.nkf] = Tcurrent;
        P.kf_slot[(size_t)b * P.kmax + S.nkf] = newslot;
        S.nkf += 1;
        S.cur_slot = newslot + 1;                                               // next never-used slot of this sequence
      } else {
        const int evicted = P.kf_slot[(size_t)b * P.kmax];
        for (int i = 0; i + 1 < S.nkf; ++i) {
          P.kf_pose[(size_t)b * P.kmax + i] = P.kf_pose[(size_t)b * P.kmax + i + 1];
          P.kf_slot[(size_t)b * P.kmax + i] = P.kf_slot[(size_t)b * P.kmax + i + 1];
        }
        P.kf_pose[(size_t)b * P.kmax + S.nkf - 1] = Tcurrent;
        P.kf_slot[(size_t)b * P.kmax + S.nkf - 1] = newslot;
        S.cur_slot = evicted;                                                   // the evicted keyframe's slot takes the next scan
      }
    }
    S.T_prev = Tcurrent;                                                        // :257
  }
  if (S.step < P.max_steps) {
    double* t = P.traj + ((size_t)b * P.max_steps + S.step) * 3;
    t[0] = Tcurrent.x; t[1] = Tcurrent.y; t[2] = t2_yaw(Tcurrent);
    P.kf_flag[(size_t)b * P.max_steps + S.step] = fuse;
    reinterpret_cast<RegStatsK6*>(P.traj_stats)[(size_t)b * P.max_steps + S.step] = st;
  }
  S.step += 1;
  P.state[b] = S;
}

__global__ void k6_init(const SeqParams P, int slot_base) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= P.nseq) return;
  SeqState S;
  S.T_prev = t2_identity(); S.Tmot = t2_identity(); S.Tguess = t2_identity();
  S.nkf = 0; S.cur_slot = slot_base + b * (P.kmax + 1); S.step = 0; S.pad = 0;
  P.state[b] = S;
}

}  // namespace cfear
