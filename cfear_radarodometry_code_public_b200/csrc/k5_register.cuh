// K5/K6: scan-to-keyframes registration, one CTA per independent problem, no host round trips.
//
// Replaces n_scan_normal_reg::Register (n_scan_normal.cpp:82-187) with everything below it:
//   BuildOptimizationProblem / AddScanPairCost       n_scan_normal.cpp:344-391, 215-326
//     MapPointNormal::GetClosestIdx                  pointnormal.cpp:238-254  (exact fp32 1-NN, d2 < r*r)
//     Weights::GetWeight                             registration.cpp:67-76
//   cost functors P2L / P2D / P2P (autodiff == the analytic Jacobians used here)  n_scan_normal.h:180-255, 330-361
//   Registration::GetLoss + ScaledLoss               registration.cpp:78-97, n_scan_normal.cpp:277
//   SolveOptimizationProblem -> ceres::Solve         n_scan_normal.cpp:443-452 (trust-region LM, Ceres defaults)
//   GetCovariance                                    n_scan_normal.cpp:392-433
//
// Per outer iteration the block (1) associates every (keyframe, source cell) pair through the keyframe's
// bucket grid and compacts the accepted pairs, in (keyframe, cell) order, into a residual list, then
// (2) runs the LM loop: every evaluation is a block-wide pass over the list with the 6 J^T J + 3 J^T r
// + cost sums reduced by warp shuffles and a fixed-order cross-warp sum (bit-reproducible), while the
// scalar trust-region logic is executed redundantly by every thread on identical inputs.
#pragma once
#include <cuda_fp16.h>
#include "common.cuh"
#include "tma.cuh"

namespace cfear {

#ifndef CFEAR_K5_THREADS
#define CFEAR_K5_THREADS 128
#endif
constexpr int K5_THREADS = CFEAR_K5_THREADS;   // 4 warps per problem, THREE problems resident per SM: 168 registers (no spills) and 74 KB of
                                               // shared memory per CTA (50-byte residual records, per-scan tables sized by the problem, 6 m NN grid).
                                               // A CTA alone is slower than the 2-per-SM forms (0.381 ms / 256 problems vs 0.343 for 192 threads,
                                               // 0.347 for 128 threads with 202 registers), but with four steps in flight the step takes 0.337 ms
                                               // instead of 0.370 / 0.356 (profiles/r02q_k5_threads_ab.txt, profiles/r02v_k5_three_per_sm_ab.txt)
// The WIDE form: the same kernel with one 384-thread CTA per SM, launched when a batch has no more problems than the GPU
// has SMs (sequence replay with few sequences, the drop-in classes' one-scan-at-a-time calls).  There the SM would hold one
// 128-thread CTA and the step lasts as long as one problem does; twelve warps cut the association and evaluation passes
// to a third (the register budget is the same 65536 / 384 = 170 per thread).  Only the product instantiations (AUX = false).
constexpr int K5_THREADS_WIDE = 384;
// The MID form: 192 threads, two CTAs per SM, for a stream-ordered launch of more problems than SMs but at most two per
// SM (bench.py's 256-problem step when nothing overlaps it): the third 128-thread slot would stay empty there.
constexpr int K5_THREADS_MID = 192;
#ifndef CFEAR_K5_MINBLOCKS
#define CFEAR_K5_MINBLOCKS (CFEAR_K5_THREADS > 256 ? 1 : (CFEAR_K5_THREADS > 128 ? 2 : 3))
#endif
constexpr int K5_MINBLOCKS = CFEAR_K5_MINBLOCKS;   // resident CTAs per SM the register and shared-memory budgets are set for
constexpr int K5_MAXSCANS = 65;      // K+1 <= 65
#ifndef CFEAR_K5_SMEM_KB
#define CFEAR_K5_SMEM_KB 0                     // 0: what lets CFEAR_K5_MINBLOCKS CTAs share an SM (cfear_create)
#endif
constexpr int K5_SMEM_BYTES = CFEAR_K5_SMEM_KB * 1024;   // dynamic smem per CTA

struct RegParams {
  CellPool pool;
  int nprob, nscans;             // nscans = K+1 cell sets per problem, last = current scan (row stride of slots / poses)
  const int32_t* nscans_pp;      // optional [nprob]: cell sets actually used by problem b (<= nscans); the current scan is entry nscans_pp[b]-1
  const int32_t* slots;          // [nprob][nscans]
  double* poses;                 // [nprob][nscans][3] in/out
  double* cov36;                 // [nprob][36]
  void* stats;                   // [nprob] cfear_reg_stats
  int32_t* assoc;                // [nprob][nscans-1][max_cells] or null   } association outputs and the soft prior are served by
  double* assoc_sim;             // same shape: direction similarity of the pair  } the AUX instantiations only (k5_launch_cost*)
  const double* soft_L;          // [nprob][9] or null: row-major lower-triangular sqrt information of the guess prior
  double2* res;                  // [nprob][4][res_cap] residual scratch, SoA: (px,py) (qx,qy) (a,b) (c,w)
  int res_cap;
  int smem_bytes;                // dynamic shared memory given to the kernel
  int cost, loss, weight_opt, solver_mode;
  int max_outer, min_outer, max_inner, gn_iters;
  double loss_limit, cov_scale, regularization, radius;
};

struct RegStatsDev {             // == cfear_reg_stats
  int32_t success, outer_iterations, inner_iterations, num_residuals, num_blocks, usable;
  double final_cost, score;
  int32_t pose_written, reserved;
};

// 1 / sqrt(s) for s in the normal range (callers guarantee s > loss_limit^2 > 0): the 22-bit hardware seed and two
// Newton steps, 9 FP64-pipe instructions instead of the ~25 (plus special-case branches) of the library rsqrt().
// Within 2 ulp of the exact value.
__device__ __forceinline__ double rsqrt_normal(double s) {
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(s));
  const double h = 0.5 * s;
  double e = fma(-h * y, y, 0.5);
  y = fma(y, e, y);
  e = fma(-h * y, y, 0.5);
  return fma(y, e, y);
}

// 1 / x for x in the normal range: hardware seed and two Newton steps, no special-case branches (so independent
// chains interleave); within 2 ulp.  Used where the operand is known to be a well-scaled positive quantity.
__device__ __forceinline__ double rcp_normal(double x) {
  double y;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  double e = fma(-x, y, 1.0);
  y = fma(y, e, y);
  e = fma(-x, y, 1.0);
  return fma(y, e, y);
}

// ceres/loss_function.cc (rho[2] is only needed by Ceres' corrector when rho'' > 0, which none of these have)
template <int LOSS>
__device__ __forceinline__ void loss_eval(double a, double s, double rho[3]) {
  if constexpr (LOSS == 1) {            // HuberLoss
    const double b = a * a;
    // sqrt(s) = s * rsqrt(s), a / sqrt(s) = a * rsqrt(s): one reciprocal square root instead of a square root and a
    // division per outlier (each within an ulp of the library forms; the unused rho[2] is dead code).  Branch-free:
    // in a warp some residual is always an outlier, so the reciprocal square root is paid anyway, and without the branch
    // the two residuals of an unrolled iteration interleave.
    const bool outlier = s > b;
    const double ri = rsqrt_normal(outlier ? s : b);
    rho[0] = outlier ? fma(2.0 * a, s * ri, -b) : s;
    rho[1] = outlier ? fmax(2.2250738585072014e-308, a * ri) : 1.0;
    rho[2] = outlier ? -rho[1] / (2.0 * s) : 0.0;
  } else if constexpr (LOSS == 2) {     // CauchyLoss
    const double b = a * a, c = 1.0 / b;
    const double sum = 1.0 + s * c, inv = 1.0 / sum;
    rho[0] = b * log(sum); rho[1] = fmax(2.2250738585072014e-308, inv); rho[2] = -c * (inv * inv);
  } else if constexpr (LOSS == 3) {     // SoftLOneLoss
    const double b = a * a, c = 1.0 / b;
    const double sum = 1.0 + s * c, tmp = sqrt(sum);
    rho[0] = 2.0 * b * (tmp - 1.0); rho[1] = fmax(2.2250738585072014e-308, 1.0 / tmp); rho[2] = -(c * rho[1]) / (2.0 * sum);
  } else if constexpr (LOSS == 4) {     // ComposedLoss(Huber(1), Cauchy(1))  registration.cpp:88-92
    double g[3], f[3];
    loss_eval<2>(1.0, s, g); loss_eval<1>(1.0, g[0], f);
    rho[0] = f[0]; rho[1] = f[1] * g[1]; rho[2] = f[2] * g[1] * g[1] + f[1] * g[2];
  } else if constexpr (LOSS == 5) {     // TukeyLoss
    const double a2 = a * a;
    if (s <= a2) { const double v = 1.0 - s / a2, v2 = v * v; rho[0] = a2 / 3.0 * (1.0 - v2 * v); rho[1] = v2; rho[2] = -2.0 / a2 * v; }
    else { rho[0] = a2 / 3.0; rho[1] = 0.0; rho[2] = 0.0; }
  } else {                              // None: ScaledLoss(nullptr, w)
    rho[0] = s; rho[1] = 1.0; rho[2] = 0.0;
  }
}

struct EvalOut { double cost, H[6], g[3]; };

// Residual list of one problem: per accepted correspondence the world-frame target q, the cost's constants (a, b) and
// (c, w) -- three double2 fields, SoA -- and the index j of its source cell (16 bits); the source mean p is read through
// the index from the source set's means (a shared-memory copy in the normal case), so a record takes 50 bytes instead of
// 64.  The first cap_s entries live in shared memory, the overflow in the problem's global scratch.
struct ResList {
  double2* s; uint16_t* sj; int cap_s;     // shared part: field f (0 q, 1 ab, 2 cw) at s + f*cap_s, source index at sj
  double2* g; uint16_t* gj; int cap_g;     // global part: field f at g + f*cap_g, source index at gj (indexed by the residual's global position)
  const double2* src_mean;                 // p of residual r = src_mean[j(r)]
  __device__ __forceinline__ double2 ld(int f, int r) const { return r < cap_s ? s[f * cap_s + r] : g[(size_t)f * cap_g + r]; }
  __device__ __forceinline__ double2 ldp(int r) const { return src_mean[r < cap_s ? sj[r] : gj[r]]; }
  __device__ __forceinline__ void st(int r, int j, double2 q, double2 ab, double2 cw) const {
    if (r < cap_s) { s[r] = q; s[cap_s + r] = ab; s[2 * cap_s + r] = cw; sj[r] = (uint16_t)j; }
    else { g[r] = q; g[(size_t)cap_g + r] = ab; g[2 * (size_t)cap_g + r] = cw; gj[r] = (uint16_t)j; }
  }
};
// -DCFEAR_K5_PROFILE: clock64 probes (thread 0's view), reported through unused covariance slots; profiles/ab_stage.py --prof
#ifdef CFEAR_K5_PROFILE
#define PROF_T(v) const long long v = clock64()
#define PROF_ADD(acc, a, b) (acc) += (b) - (a)
#define PROF_ARG , prof
#define PROF_PARAM , long long* prof
#else
#define PROF_ARG
#define PROF_PARAM
#define PROF_T(v)
#define PROF_ADD(acc, a, b)
#endif


// Row j of a residual block, loss-corrected weight wr = w rho'(s): adds wr j^T j to the packed upper triangle
// acc[1..6] = (xx, xy, xt, yy, yt, tt) and wr j r to acc[7..9].  Z0 / Z1: the row's d/dx / d/dy entry is the constant 0
// (its products are not formed at all).  Explicit FMAs: the file is compiled with -fmad=false for the fp32 / point
// arithmetic that has to round like the reference's x86 build, but these sums have no such contract (their order
// already differs from a sequential CPU sum) and the FP64 pipe (64 lanes per SM) is what bounds an evaluation pass.
template <bool Z0, bool Z1>
__device__ __forceinline__ void add_row(double wr, double j0, double j1, double j2, double r, double acc[10]) {
  if constexpr (!Z0) {
    const double u0 = wr * j0;
    acc[1] = fma(u0, j0, acc[1]);
    if constexpr (!Z1) acc[2] = fma(u0, j1, acc[2]);
    acc[3] = fma(u0, j2, acc[3]);
    acc[7] = fma(u0, r, acc[7]);
  }
  if constexpr (!Z1) {
    const double u1 = wr * j1;
    acc[4] = fma(u1, j1, acc[4]);
    acc[5] = fma(u1, j2, acc[5]);
    acc[8] = fma(u1, r, acc[8]);
  }
  const double u2 = wr * j2;
  acc[6] = fma(u2, j2, acc[6]);
  acc[9] = fma(u2, r, acc[9]);
}

// One residual block's contribution at x (cs = cos psi, sn = sin psi): cost and (optionally) normal equations.
// Residuals / Jacobians: n_scan_normal.h:180-255, 330-361; loss: ScaledLoss(w) around the base loss with
// Ceres' corrector in its rho'' <= 0 form (rows scaled by sqrt(w rho')).
template <int COST, int LOSS, bool JAC>
__device__ __forceinline__ void accumulate(double loss_limit, double cs, double sn, const double x[3], double2 p, double2 q,
                                           double2 ab, double2 cw, double acc[10]) {
  const double rx = fma(cs, p.x, -(sn * p.y)), ry = fma(sn, p.x, cs * p.y);
  const double ex = rx + x[0] - q.x, ey = ry + x[1] - q.y;
  // d(R p)/d psi = (-ry, rx)
  const double a = ab.x, b = ab.y, c = cw.x, w = cw.y;
  double r0, r1 = 0.0;
  if constexpr (COST == 1) r0 = fma(ex, a, ey * b);
  else if constexpr (COST == 2) { r0 = a * ex; r1 = fma(b, ex, c * ey); }
  else { r0 = -ex; r1 = -ey; }
  const double s = fma(r0, r0, r1 * r1);
  double rho[3];
  loss_eval<LOSS>(loss_limit, s, rho);
  acc[0] = fma(0.5 * w, rho[0], acc[0]);
  if constexpr (JAC) {
    const double wr = w * rho[1];
    if constexpr (COST == 1) {
      add_row<false, false>(wr, a, b, fma(rx, b, -(ry * a)), r0, acc);                  // J = (a, b, dpx a + dpy b)
    } else if constexpr (COST == 2) {
      add_row<false, true>(wr, a, 0.0, -(a * ry), r0, acc);                             // J0 = (a, 0, a dpx)
      add_row<false, false>(wr, b, c, fma(c, rx, -(b * ry)), r1, acc);                  // J1 = (b, c, b dpx + c dpy)
    } else {
      add_row<false, true>(wr, -1.0, 0.0, ry, r0, acc);                                 // J0 = (-1, 0, -dpx)
      add_row<true, false>(wr, 0.0, -1.0, -rx, r1, acc);                                // J1 = (0, -1, -dpy)
    }
  }
}

// ---- evaluation service -------------------------------------------------------------------------------------------
// The scalar trust-region arithmetic (a few hundred dependent fp64 instructions per LM iteration: a 3x3 Cholesky,
// divisions, square roots, sincos of the candidate yaw) runs in WARP 0 ONLY; replicated in every warp it saturates the
// FP64 pipe shared by the two resident problems and becomes the longest part of an iteration.  Warp 0 publishes the
// point to evaluate (x, cos psi, sin psi) in shared memory and raises barrier A; every warp (warp 0 included) then
// accumulates its share of the residual list, reduces it by xor-butterfly, lane 0 stores the 10 partial sums, barrier B;
// warp 0 adds the K5_WARPS partials in fixed order.  Bit-reproducible: fixed per-thread order, fixed tree, fixed order.
struct LMShared {
  double bc[5];            // x0, x1, psi, cos psi, sin psi of the requested evaluation
  int ctl, pad;            // 1 = evaluate, 0 = done
  double out_x[3];         // results broadcast at the end of a solve
  double out_final_cost, out_last_rel;
  int out_niter, out_usable, out_success, out_inner;
  // soft-constraint prior on the free block (n_scan_normal.cpp:373-377, mahalanobisDistanceError n_scan_normal.h:259-290):
  // r = alpha L (guess - x), alpha = sqrt(#source cells), no loss.  M = alpha^2 L^T L is its constant J^T J.
  int soft, pad2;
  double guess[3], aL[6], M[6];          // aL = alpha * (l00 l10 l11 l20 l21 l22), M packed like EvalOut::H
  // normal equations at the accepted point of the running LM solve: kept here rather than in warp 0's registers (they are
  // read once per linear solve; 20 registers less for every thread of the kernel)
  double acc_H[6], acc_g[3];
};

// Barriers A ("the point to evaluate -- or the stop mark -- is published") and B ("every warp's partial sums are stored")
// of an evaluation round.  Every warp of the block executes them from the SAME place (eval_round below, inside loops
// that all warps run): round 1 used a producer / consumer form in which warp 0 and the serving warps reached the same
// named barrier from different code locations -- legal PTX, but compute-sanitizer's synccheck reports it as "divergent
// thread(s) in block".  With the shared loop all three sanitizer tools are clean.
template <int NT> __device__ __forceinline__ void bar_a() { asm volatile("bar.sync 1, %0;" ::"n"(NT) : "memory"); }
template <int NT> __device__ __forceinline__ void bar_b() { asm volatile("bar.sync 2, %0;" ::"n"(NT) : "memory"); }

// Warp reduction of the 10 sums by recursive halving: at each step a lane hands the half of its values its partner
// keeps to that partner, so 5+3+2+1+1 = 12 exchanges replace the 50 of ten separate butterflies.  Ends with sum i in the
// even lane whose bits (16,8,4,2) encode i; fixed tree, so the result is bit-reproducible.
__device__ __forceinline__ double xchg_add(bool upper, double lo, double hi, int d) {
  const double send = upper ? lo : hi, keep = upper ? hi : lo;
  return keep + __shfl_xor_sync(FULL, send, d);
}
__device__ __forceinline__ void warp_reduce10(const double acc[10], double* s_part_warp) {
  const int l = lane_id();
  const bool b4 = l & 16, b3 = l & 8, b2 = l & 4, b1 = l & 2;
  const double v0 = xchg_add(b4, acc[0], acc[5], 16), v1 = xchg_add(b4, acc[1], acc[6], 16), v2 = xchg_add(b4, acc[2], acc[7], 16),
               v3 = xchg_add(b4, acc[3], acc[8], 16), v4 = xchg_add(b4, acc[4], acc[9], 16);
  const double w0 = xchg_add(b3, v0, v3, 8), w1 = xchg_add(b3, v1, v4, 8), w2 = xchg_add(b3, v2, 0.0, 8);   // (v0 v1 v2 | v3 v4 -)
  const double u0 = xchg_add(b2, w0, w2, 4), u1 = xchg_add(b2, w1, 0.0, 4);                                 // (w0 w1 | w2 -)
  double t = xchg_add(b1, u0, u1, 2);                                                                       // (u0 | u1)
  t += __shfl_xor_sync(FULL, t, 1);
  const bool pad = b2 && (b1 || b3);
  if (!(l & 1) && !pad) s_part_warp[(b4 ? 5 : 0) + (b3 ? 3 + (b1 ? 1 : 0) : (b2 ? 2 : 0) + (b1 ? 1 : 0))] = t;
}

// every warp: this thread's share of the list at (x, cs, sn) -> per-warp partial sums in s_part[warp][10]
template <int COST, int LOSS, int NT>
__device__ __forceinline__ void eval_contrib(double loss_limit, const ResList& res, int nres, const double x[3], double cs,
                                             double sn, double* s_part) {
  double acc[10];
#pragma unroll
  for (int i = 0; i < 10; ++i) acc[i] = 0.0;
  int r = threadIdx.x;
  if (nres <= res.cap_s) {                   // the whole list is in shared memory (the normal case): no per-load path select
    const double2* f1 = res.s, * f2 = res.s + res.cap_s, * f3 = res.s + 2 * res.cap_s;
    const uint16_t* fj = res.sj; const double2* sm = res.src_mean;
    for (; r + NT < nres; r += 2 * NT) {
      const int r2 = r + NT;
      const int j0 = fj[r], j1 = fj[r2];
      const double2 q0 = f1[r], ab0 = f2[r], cw0 = f3[r];
      const double2 q1 = f1[r2], ab1 = f2[r2], cw1 = f3[r2];
      const double2 p0 = sm[j0], p1 = sm[j1];
      accumulate<COST, LOSS, true>(loss_limit, cs, sn, x, p0, q0, ab0, cw0, acc);
      accumulate<COST, LOSS, true>(loss_limit, cs, sn, x, p1, q1, ab1, cw1, acc);
    }
    if (r < nres) accumulate<COST, LOSS, true>(loss_limit, cs, sn, x, sm[fj[r]], f1[r], f2[r], f3[r], acc);
  } else {
    for (; r + NT < nres; r += 2 * NT) {
      const int r2 = r + NT;
      const double2 p0 = res.ldp(r), q0 = res.ld(0, r), ab0 = res.ld(1, r), cw0 = res.ld(2, r);
      const double2 p1 = res.ldp(r2), q1 = res.ld(0, r2), ab1 = res.ld(1, r2), cw1 = res.ld(2, r2);
      accumulate<COST, LOSS, true>(loss_limit, cs, sn, x, p0, q0, ab0, cw0, acc);
      accumulate<COST, LOSS, true>(loss_limit, cs, sn, x, p1, q1, ab1, cw1, acc);
    }
    if (r < nres) accumulate<COST, LOSS, true>(loss_limit, cs, sn, x, res.ldp(r), res.ld(0, r), res.ld(1, r), res.ld(2, r), acc);
  }
  warp_reduce10(acc, s_part + warp_id() * 10);
}

// One evaluation round, executed by EVERY warp of the block.  Warp 0 owns the point y (all its lanes hold it) and the
// decision to stop; it publishes either in shared memory, barrier A, every warp accumulates its share of the residual
// list and stores its 10 partial sums, barrier B, warp 0 adds the K5_WARPS partials in fixed order (bit-reproducible) and,
// in the AUX instantiations, the soft prior's block.  Returns false (block-uniform) on the stop mark; `ev` is valid in warp 0.
template <int COST, int LOSS, bool AUX, int NT>
__device__ __forceinline__ bool eval_round(double loss_limit, const ResList& res, int nres, bool w0, bool stop, const double y[3],
                                           EvalOut& ev, LMShared* sh, double* s_part PROF_PARAM) {
  PROF_T(te0);
  double yy[3] = {y[0], y[1], y[2]}, cs = 1.0, sn = 0.0;
  if (w0) {
    if (!stop) sincos(yy[2], &sn, &cs);
    if (lane_id() == 0) { sh->bc[0] = yy[0]; sh->bc[1] = yy[1]; sh->bc[2] = yy[2]; sh->bc[3] = cs; sh->bc[4] = sn; sh->ctl = stop ? 0 : 1; }
  }
  PROF_T(te1);
  bar_a<NT>();
  if (sh->ctl == 0) return false;
  if (!w0) { yy[0] = sh->bc[0]; yy[1] = sh->bc[1]; yy[2] = sh->bc[2]; cs = sh->bc[3]; sn = sh->bc[4]; }
  eval_contrib<COST, LOSS, NT>(loss_limit, res, nres, yy, cs, sn, s_part);
  PROF_T(te2);
  bar_b<NT>();
  PROF_T(te3);
  PROF_ADD(prof[4], te0, te1); PROF_ADD(prof[5], te1, te2); PROF_ADD(prof[6], te2, te3); PROF_ADD(prof[7], 0, 1);
  if (w0) {
    // every lane adds the K5_WARPS partials of every sum in warp order (broadcast shared-memory reads, no shuffles)
    double tot[10];
#pragma unroll
    for (int i = 0; i < 10; ++i) tot[i] = s_part[i];
#pragma unroll
    for (int w = 1; w < NT / 32; ++w) {
#pragma unroll
      for (int i = 0; i < 10; ++i) tot[i] += s_part[w * 10 + i];
    }
    ev.cost = tot[0];
#pragma unroll
    for (int i = 0; i < 6; ++i) ev.H[i] = tot[1 + i];
#pragma unroll
    for (int i = 0; i < 3; ++i) ev.g[i] = tot[7 + i];
    if constexpr (AUX) {
      if (sh->soft) {                              // + the prior's residual block (3 rows, Jacobian -alpha L)
        const double d0 = sh->guess[0] - yy[0], d1 = sh->guess[1] - yy[1], d2 = sh->guess[2] - yy[2];
        const double* L = sh->aL;
        const double r0 = L[0] * d0, r1 = L[1] * d0 + L[2] * d1, r2 = L[3] * d0 + L[4] * d1 + L[5] * d2;
        ev.cost += 0.5 * (r0 * r0 + r1 * r1 + r2 * r2);
#pragma unroll
        for (int i = 0; i < 6; ++i) ev.H[i] += sh->M[i];
        ev.g[0] -= L[0] * r0 + L[1] * r1 + L[3] * r2;
        ev.g[1] -= L[2] * r1 + L[4] * r2;
        ev.g[2] -= L[5] * r2;
      }
    }
  }
  return true;
}

// Symmetric positive-definite 3x3 solve (xx,xy,xt,yy,yt,tt) by Cholesky; one reciprocal square root per pivot.
__device__ __forceinline__ bool chol3_solve(const double A[6], const double b[3], double y[3]) {
  if (!(A[0] > 0.0) || !isfinite(A[0])) return false;
  const double r00 = rsqrt_normal(A[0]);
  const double l10 = A[1] * r00, l20 = A[2] * r00;
  const double d1 = A[3] - l10 * l10;
  if (!(d1 > 0.0)) return false;
  const double r11 = rsqrt_normal(d1);
  const double l21 = (A[4] - l20 * l10) * r11;
  const double d2 = A[5] - l20 * l20 - l21 * l21;
  if (!(d2 > 0.0)) return false;
  const double r22 = rsqrt_normal(d2);
  const double z0 = b[0] * r00;
  const double z1 = (b[1] - l10 * z0) * r11;
  const double z2 = (b[2] - l20 * z0 - l21 * z1) * r22;
  y[2] = z2 * r22;
  y[1] = (z1 - l21 * y[2]) * r11;
  y[0] = (z0 - l10 * y[1] - l20 * y[2]) * r00;
  return isfinite(y[0]) && isfinite(y[1]) && isfinite(y[2]);
}

// The same solve by the adjugate (explicit FMAs): every cofactor is independent, so the dependent chain is
// products -> determinant -> one reciprocal -> scale, a third of the Cholesky chain (three reciprocal square roots in
// sequence).  Positive definiteness is checked on the leading principal minors (the Cholesky pivots' signs).  Used in the
// LM loop, where the damped, Jacobi-scaled matrix is well conditioned; CFEAR_K5_CHOL restores the Cholesky solve there.
__device__ __forceinline__ bool adj3_solve(const double A[6], const double b[3], double y[3]) {
  const double a00 = A[0], a01 = A[1], a02 = A[2], a11 = A[3], a12 = A[4], a22 = A[5];
  const double c00 = fma(a11, a22, -(a12 * a12)), c01 = fma(a02, a12, -(a01 * a22)), c02 = fma(a01, a12, -(a02 * a11));
  const double c11 = fma(a00, a22, -(a02 * a02)), c12 = fma(a01, a02, -(a00 * a12)), c22 = fma(a00, a11, -(a01 * a01));
  const double det = fma(a00, c00, fma(a01, c01, a02 * c02));
  if (!(a00 > 0.0) || !(c22 > 0.0) || !(det > 0.0) || !isfinite(det)) return false;
  const double inv = rcp_normal(det);
  const double n0 = fma(c00, b[0], fma(c01, b[1], c02 * b[2]));
  const double n1 = fma(c01, b[0], fma(c11, b[1], c12 * b[2]));
  const double n2 = fma(c02, b[0], fma(c12, b[1], c22 * b[2]));
  y[0] = n0 * inv; y[1] = n1 * inv; y[2] = n2 * inv;
  return isfinite(y[0]) && isfinite(y[1]) && isfinite(y[2]);
}

struct SolveSum { double final_cost; int n_iterations; double last_rel; bool usable; };

// warp 0 (all lanes hold the same values): the accepted point's normal equations -> shared memory
__device__ __forceinline__ void store_accepted(LMShared* sh, const EvalOut& e) {
  __syncwarp();                                  // earlier reads of the previous values are done
  if (lane_id() == 0) {
#pragma unroll
    for (int i = 0; i < 6; ++i) sh->acc_H[i] = e.H[i];
#pragma unroll
    for (int i = 0; i < 3; ++i) sh->acc_g[i] = e.g[i];
  }
  __syncwarp();
}

// Trust-region LM with Ceres defaults (trust_region_minimizer.cc / levenberg_marquardt_strategy.cc).  Called by EVERY
// warp: the loop body is one evaluation round (all warps) followed by the scalar trust-region logic, which only warp 0
// executes (all lanes redundantly) -- the others go straight to the next round's barrier.  The candidate point is
// evaluated with its Jacobian in the same pass, so an accepted step needs no second pass over the residuals.  x and sum
// are valid in warp 0; LMShared::acc_H is left holding J^T J (loss-corrected, unscaled) at the returned x, which is what
// GetCovariance needs.
template <int COST, int LOSS, bool AUX, int NT>
__device__ __forceinline__ void lm_solve(const RegParams& P, const ResList& res, int nres, bool w0, double x[3], SolveSum& sum,
                                         LMShared* sh, double* s_part PROF_PARAM) {
  const double kFunctionTol = 1e-6, kGradientTol = 1e-10, kParameterTol = 1e-8;
  const double kMinRelDecrease = 1e-3, kMinDiag = 1e-6, kMaxDiag = 1e32;
  const double kMaxRadius = 1e16, kMinRadius = 1e-32;
  double radius = 1e4, decrease_factor = 2.0;
  // 1 / radius is carried along multiplicatively (the oracle divides the diagonal by the radius; both are the same damping
  // to an ulp) so that no division sits between an accepted step and the next linear solve
  double inv_radius = 1e-4;
  bool reuse_diagonal = false, first = true, stop = false;
  int invalid_in_a_row = 0, it = 0;
  sum.usable = true; sum.n_iterations = 1; sum.last_rel = 0.0; sum.final_cost = 0.0;
  double scale[3] = {1, 1, 1}, diag[3] = {0, 0, 0};
  double x_cost = 0.0, x_n2 = 0.0, min_cost = 0.0, model_change = 0.0, inv_model_change = 0.0;
  double xc[3] = {x[0], x[1], x[2]}, delta[3] = {0, 0, 0};              // the first round evaluates x itself
  for (;;) {
    EvalOut evc;                                                        // (dead across the round: not loop-carried)
    if (!eval_round<COST, LOSS, AUX, NT>(P.loss_limit, res, nres, w0, stop, xc, evc, sh, s_part PROF_ARG)) break;
    if (!w0) continue;
    if (first) {                                                        // iteration 0: cost, gradient, Jacobi scaling at x
      first = false;
      store_accepted(sh, evc);
      x_cost = evc.cost;
      scale[0] = 1.0 / (1.0 + sqrt(evc.H[0]));
      scale[1] = 1.0 / (1.0 + sqrt(evc.H[3]));
      scale[2] = 1.0 / (1.0 + sqrt(evc.H[5]));
      x_n2 = x[0] * x[0] + x[1] * x[1] + x[2] * x[2];
      min_cost = x_cost;
      sum.final_cost = min_cost;
      if (fmax(fabs(evc.g[0]), fmax(fabs(evc.g[1]), fabs(evc.g[2]))) <= kGradientTol) { stop = true; continue; }
    } else {                                                            // the candidate of iteration `it` has been evaluated
      const double cand_cost = evc.cost;
      // parameter tolerance |delta| <= tol (|x| + tol): the square roots are only taken when the squares come close
      const double step_n2 = delta[0] * delta[0] + delta[1] * delta[1] + delta[2] * delta[2];
      if (step_n2 <= 4.0 * kParameterTol * kParameterTol * (x_n2 + kParameterTol * kParameterTol)) {      // (a+b)^2 <= 2a^2+2b^2, doubled again
        if (sqrt(step_n2) <= kParameterTol * (sqrt(x_n2) + kParameterTol)) { stop = true; continue; }
      }
      const double cost_change = x_cost - cand_cost;
      if (fabs(cost_change) <= kFunctionTol * x_cost) { stop = true; continue; }
      const double rel = cost_change * inv_model_change;
      sum.n_iterations++; sum.last_rel = rel;
      if (rel > kMinRelDecrease) {
        x[0] = xc[0]; x[1] = xc[1]; x[2] = xc[2];
        x_n2 = x[0] * x[0] + x[1] * x[1] + x[2] * x[2];
        store_accepted(sh, evc);
        x_cost = evc.cost;
        const double gmax = fmax(fabs(evc.g[0]), fmax(fabs(evc.g[1]), fabs(evc.g[2])));
        // radius /= max(1/3, 1 - (2 rel - 1)^3).  A step with rel >= 0.94 has (2 rel - 1)^3 > 2/3, i.e. the maximum picks
        // 1/3 whatever the last bits of rel are: that case needs neither the quotient nor the cube.
        double f = 1.0 / 3.0;
        if (!(cost_change >= 0.94 * model_change)) { const double t = 2.0 * rel - 1.0; f = fmax(1.0 / 3.0, 1.0 - t * t * t); }
        radius = fmin(kMaxRadius, radius * rcp_normal(f));           // f in [1/3, 2]
        inv_radius = fmax(1.0 / kMaxRadius, inv_radius * f);
        decrease_factor = 2.0; reuse_diagonal = false;
        min_cost = fmin(min_cost, x_cost); sum.final_cost = min_cost;
        if (it >= P.max_inner || gmax <= kGradientTol) { stop = true; continue; }
      } else {
        radius /= decrease_factor; inv_radius *= decrease_factor; decrease_factor *= 2.0; reuse_diagonal = true;
        min_cost = fmin(min_cost, cand_cost); sum.final_cost = min_cost;
        if (it >= P.max_inner || radius <= kMinRadius) { stop = true; continue; }
      }
    }
    // the next candidate: linear solves until a valid step comes out (an invalid one shrinks the radius and counts as an iteration)
    for (;;) {
      ++it;
      const double* aH = sh->acc_H; const double* ag = sh->acc_g;
      const double Hs[6] = {aH[0] * scale[0] * scale[0], aH[1] * scale[0] * scale[1], aH[2] * scale[0] * scale[2],
                            aH[3] * scale[1] * scale[1], aH[4] * scale[1] * scale[2], aH[5] * scale[2] * scale[2]};
      const double gs[3] = {ag[0] * scale[0], ag[1] * scale[1], ag[2] * scale[2]};
      if (!reuse_diagonal) {
        diag[0] = fmin(fmax(Hs[0], kMinDiag), kMaxDiag);
        diag[1] = fmin(fmax(Hs[3], kMinDiag), kMaxDiag);
        diag[2] = fmin(fmax(Hs[5], kMinDiag), kMaxDiag);
      }
      const double A[6] = {Hs[0] + diag[0] * inv_radius, Hs[1], Hs[2], Hs[3] + diag[1] * inv_radius, Hs[4], Hs[5] + diag[2] * inv_radius};
      double y[3];
      const double nb[3] = {-gs[0], -gs[1], -gs[2]};
#ifdef CFEAR_K5_CHOL
      const bool ok = chol3_solve(A, nb, y);
#else
      const bool ok = adj3_solve(A, nb, y);
#endif
      reuse_diagonal = true;
      model_change = 0.0;
      if (ok) {
        const double Hy0 = Hs[0] * y[0] + Hs[1] * y[1] + Hs[2] * y[2];
        const double Hy1 = Hs[1] * y[0] + Hs[3] * y[1] + Hs[4] * y[2];
        const double Hy2 = Hs[2] * y[0] + Hs[4] * y[1] + Hs[5] * y[2];
        model_change = -(y[0] * gs[0] + y[1] * gs[1] + y[2] * gs[2]) - 0.5 * (y[0] * Hy0 + y[1] * Hy1 + y[2] * Hy2);
      }
      if (!ok || !(model_change > 0.0)) {
        if (++invalid_in_a_row >= 5) { sum.usable = false; stop = true; break; }
        radius /= decrease_factor; inv_radius *= decrease_factor; decrease_factor *= 2.0; reuse_diagonal = true;
        sum.n_iterations++; sum.last_rel = 0.0;
        min_cost = fmin(min_cost, x_cost); sum.final_cost = min_cost;
        if (it >= P.max_inner || radius <= kMinRadius) { stop = true; break; }
        continue;
      }
      invalid_in_a_row = 0;
      delta[0] = y[0] * scale[0]; delta[1] = y[1] * scale[1]; delta[2] = y[2] * scale[2];
      xc[0] = x[0] + delta[0]; xc[1] = x[1] + delta[1]; xc[2] = x[2] + delta[2];
      inv_model_change = rcp_normal(model_change);       // model_change > 0; off the chain that follows the evaluation
      break;
    }
  }
}

// One keyframe's NN index as seen by the block: bucket starts + packed points, in shared memory when staged.
struct GridView { const uint16_t* gs; const float4* gp; };

// GetClosestIdx: exact fp32 nearest neighbour through the bucket grid; ties -> smallest cell index.  *nrm16 receives
// the winner's normal as packed fp16 pair (the .w of its grid point).
// A query sees only a handful of candidates (2-3 on the bench workload) spread over 2-3 bucket rows, so the walk is
// organised for latency rather than throughput: the row ranges of three rows are fetched together, their candidates
// are taken four at a time through one flattened index (four independent loads in flight), and the running minimum
// is a single 64-bit key  d2 bits : cell index : position  whose unsigned order is exactly "smaller distance, then
// smaller cell index" (d2 >= 0, so its bit pattern is monotonic).
// The acceptance test of pointnormal.cpp:250, (double)d2 < radius * radius with d2 a float, as a float comparison: d2 < T
// with T the smallest float >= radius^2 (any float below T is below radius^2, and T itself is not).
struct NNRadius { float rq, thr; };
__device__ __forceinline__ NNRadius nn_radius(double radius) {
  NNRadius r;
  r.rq = (float)radius * 1.0001f + 1e-3f;                   // bucket-range margin only
  const double r2 = radius * radius;
  float t = (float)r2;
  if ((double)t < r2) t = __uint_as_float(__float_as_uint(t) + 1u);      // next float up (r2 > 0, finite)
  r.thr = t;
  return r;
}
__device__ __forceinline__ int nn_query(const GridView& V, const NNGrid& G, double pxd, double pyd, NNRadius rad, uint32_t* nrm16) {
  const float qx = (float)pxd, qy = (float)pyd;             // pointnormal.cpp:241-242
  const float rq = rad.rq;
  int bx0 = __float2int_rd((qx - rq - G.ox) * G.inv_g), bx1 = __float2int_rd((qx + rq - G.ox) * G.inv_g);
  int by0 = __float2int_rd((qy - rq - G.oy) * G.inv_g), by1 = __float2int_rd((qy + rq - G.oy) * G.inv_g);
  bx0 = max(bx0, 0); by0 = max(by0, 0); bx1 = min(bx1, G.nx - 1); by1 = min(by1, G.ny - 1);
  *nrm16 = 0;
  if (bx0 > bx1) return -1;
  unsigned long long best = ~0ull;
  for (int byb = by0; byb <= by1; byb += 3) {
    const bool h1 = byb + 1 <= by1, h2 = byb + 2 <= by1;
    const int o0 = byb * G.nx, o1 = o0 + G.nx, o2 = o1 + G.nx;
    const int s0 = V.gs[bx0 + o0], e0 = V.gs[bx1 + o0 + 1];
    const int s1 = h1 ? V.gs[bx0 + o1] : 0, e1 = h1 ? V.gs[bx1 + o1 + 1] : 0;
    const int s2 = h2 ? V.gs[bx0 + o2] : 0, e2 = h2 ? V.gs[bx1 + o2 + 1] : 0;
    const int n0 = e0 - s0, n01 = n0 + (e1 - s1), total = n01 + (e2 - s2);
    for (int c0 = 0; c0 < total; c0 += 4) {
      unsigned long long key[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int c = c0 + k;
        const int a = c < n0 ? s0 + c : (c < n01 ? s1 + (c - n0) : s2 + (c - n01));
        key[k] = ~0ull;
        if (c < total) {
          const float4 m = V.gp[a];
          const float dx = qx - m.x, dy = qy - m.y;
          float d2 = dx * dx; d2 += dy * dy;
          key[k] = ((unsigned long long)__float_as_uint(d2) << 32) | ((unsigned long long)(__float_as_uint(m.z) & 0xffffu) << 16) | (unsigned)a;
        }
      }
      best = min(best, min(min(key[0], key[1]), min(key[2], key[3])));
    }
  }
  if (best == ~0ull) return -1;
  const float bd2 = __uint_as_float((uint32_t)(best >> 32));
  const int besti = (int)((best >> 16) & 0xffffu);
  *nrm16 = __float_as_uint(V.gp[(int)(best & 0xffffu)].w);
  if (bd2 < rad.thr) return besti;                          // pointnormal.cpp:250
  return -1;
}

// The normal gate of n_scan_normal.cpp:244-247, max(n_src' . n_tar, 0) > cos(30 deg) in fp64, decided from the fp16 copy
// of the target normal whenever the fp32 dot product is further from the threshold than the copy's rounding error can
// explain; only the borderline pairs fetch the fp64 normal.  Same decisions as the fp64 test, by construction.
__device__ __forceinline__ bool normal_gate(double ntx, double nty, uint32_t nrm16, const double2* tar_normal, double thr) {
  const float2 t = __half22float2(*reinterpret_cast<const __half2*>(&nrm16));
  const float sx = (float)ntx, sy = (float)nty;
  const float dot = sx * t.x + sy * t.y;
  const float band = (fabsf(sx) + fabsf(sy)) * (6e-4f * (fabsf(t.x) + fabsf(t.y)) + 1e-4f);
  const float th = (float)thr;
  if (dot > th + band) return true;
  if (dot < th - band) return false;
  const double2 n = *tar_normal;                                        // borderline (or non-finite): exact test
  return fmax(ntx * n.x + nty * n.y, 0.0) > thr;
}

__device__ __forceinline__ double sim_ratio(double x, double y) { return 2 * fmin(x, y) * rcp_normal(x + y); }

constexpr int K5_TILE_MAX = 4096;     // (keyframe, source cell) pairs associated per tile (2 x u16 of shared memory each)
constexpr uint16_t K5_NONE = 0xffffu;

// Everything one association pass needs, block-uniform.
struct AssocCtx {
  const int32_t* slots;       // s_slots
  const double* pose;         // s_pose [nscans][5]: x y yaw cos sin
  const NNGrid* grid;         // s_grid
  const GridView* view;       // s_view
  int K, n_src;
  size_t sbase;               // first cell of the source set in the pool
  const double2* src_mean;    // [n_src] means / normals of the source cells: shared-memory copies when they fit,
  const double2* src_normal;  //         else the pool's arrays
  double x[3], cs_s, sn_s;    // pose of the current scan
  double radius;              // association radius of this outer iteration
  NNRadius nnr;               // its float forms (nn_radius)
};

// Relative transform Ttar^-1 * Tsrc of pair (keyframe i): rotation (rc, rs) and translation (tx, ty).   n_scan_normal.cpp:224
struct RelT { double rc, rs, tx, ty, ct, st, px, py; };
__device__ __forceinline__ RelT rel_transform(const AssocCtx& C, int i) {
  const double* pt = C.pose + 5 * i;
  RelT T; T.ct = pt[3]; T.st = pt[4]; T.px = pt[0]; T.py = pt[1];
  T.rc = T.ct * C.cs_s + T.st * C.sn_s; T.rs = T.ct * C.sn_s - T.st * C.cs_s;
  const double dx = C.x[0] - pt[0], dy = C.x[1] - pt[1];
  T.tx = T.ct * dx + T.st * dy; T.ty = -T.st * dx + T.ct * dy;
  return T;
}

// The residual record of an accepted pair: source mean p, world-frame target q, the cost's constants (a, b, c)
// and the weight w.   n_scan_normal.cpp:262-311, registration.cpp:67-76
template <int COST>
__device__ __forceinline__ void make_record(const RegParams& P, const RelT& T, double sim, double2 mu, double2 tm, double2 ntar,
                                            double4 Cv, double n1, double n2, double p1, double p2, double2& rp, double2& rq,
                                            double2& rab, double2& rcw) {
  double w = 1.0;
  if (P.weight_opt == 1) w = sim_ratio(n1, n2);
  else if (P.weight_opt == 2) w = sim;
  else if (P.weight_opt == 3) w = sim_ratio(p1, p2);
  else if (P.weight_opt == 4) w = sim_ratio(n1, n2) + sim + sim_ratio(p1, p2);
  const double ct = T.ct, st = T.st;
  rp = mu;
  rq = make_double2(ct * tm.x - st * tm.y + T.px, st * tm.x + ct * tm.y + T.py);
  rab = make_double2(0, 0); rcw = make_double2(0, w);
  if constexpr (COST == 1) {                                                   // :279-289
    rab = make_double2(ct * ntar.x - st * ntar.y, st * ntar.x + ct * ntar.y);
  } else if constexpr (COST == 2) {                                            // :290-300
    const double a00 = ct * Cv.x - st * Cv.z, a01 = ct * Cv.y - st * Cv.w;
    const double a10 = st * Cv.x + ct * Cv.z, a11 = st * Cv.y + ct * Cv.w;
    double s00 = a00 * ct - a01 * st, s01 = a00 * st + a01 * ct;
    double s10 = a10 * ct - a11 * st, s11 = a10 * st + a11 * ct;
    s00 = (P.regularization + s00) * P.cov_scale; s11 = (P.regularization + s11) * P.cov_scale;
    s01 = s01 * P.cov_scale; s10 = s10 * P.cov_scale;
    // inverse and its lower Cholesky factor with one reciprocal, one reciprocal square root and one square root
    const double idet = rcp_normal(s00 * s11 - s01 * s10);
    const double i00 = s11 * idet, i10 = -(s10 * idet), i11 = s00 * idet;
    const double rl = rsqrt_normal(i00);
    const double l00 = i00 * rl, l10 = i10 * rl;
    const double t11 = i11 - l10 * l10;
    const double l11 = t11 * rsqrt_normal(t11);                                  // sqrt; t11 > 0 for a positive definite block
    rab = make_double2(l00, l10); rcw.x = l11;
  }
}

// One outer iteration's association pass (n_scan_normal.cpp:215-326), two phases per tile of pairs:
//   1. every (keyframe i, source cell j) pair: transform, exact NN through the keyframe's bucket grid, 30 degree normal
//      gate -> s_nn[pair] = target cell or NONE.  Shared memory only (source cells, grids, fp16 target normals) except
//      for borderline gate decisions; no block-wide synchronisation inside.
//   2. one block scan gives every accepted pair its position, in (keyframe, cell) order; the residual records are then
//      built position by position (two in flight per thread) and stored straight into the residual list.
// Returns the number of residual blocks (block-uniform).
template <int COST, bool AUX, int NT>
__device__ __forceinline__ int build_problem(const RegParams& P, const AssocCtx& C, const ResList& res, int32_t* assoc, double* assoc_sim,
                                             int* s_warp, uint16_t* s_nn, uint16_t* s_list, int tile PROF_PARAM) {
  const int T = NT, tid = threadIdx.x;
  const int n_src = C.n_src, npairs = C.K * n_src;
  const double angle_outlier = cos(M_PI / 6.0);
  const int cap = res.cap_g;
  int nres = 0;
  for (int t0 = 0; t0 < npairs; t0 += tile) {
    const int nt = min(tile, npairs - t0);
    PROF_T(tp0);
    // ---- phase 1 ----
    // pair t = i * n_src + j, advanced by T per round: (i, j) are carried along (no division) and the relative transform
    // is recomputed only when the keyframe changes
#ifndef CFEAR_K5_P1DIV
    int pi = (t0 + tid) / n_src, pj = (t0 + tid) - pi * n_src;
    const int di = T / n_src, dj = T - di * n_src;
    int ri = -1;
    RelT R;
#endif
    for (int u = tid; u < nt; u += T) {
      PROF_T(tq0);
#ifdef CFEAR_K5_P1DIV
      const int t = t0 + u, i = t / n_src, j = t - i * n_src;
      const RelT R = rel_transform(C, i);
#else
      const int i = pi, j = pj;
      pi += di; pj += dj;
      if (pj >= n_src) { pj -= n_src; ++pi; }
      if (i != ri) { R = rel_transform(C, i); ri = i; }
#endif
      const double2 mu = C.src_mean[j];
      const double2 nsrc = C.src_normal[j];
      const double qx = R.rc * mu.x - R.rs * mu.y + R.tx, qy = R.rs * mu.x + R.rc * mu.y + R.ty;   // :240
      uint32_t n16;
#ifdef CFEAR_K5_PROFILE
      const long long tq1 = clock64() + (long long)(__double_as_longlong(qx) & 0);               // after the transform
#endif
      const int m = nn_query(C.view[i], C.grid[i], qx, qy, C.nnr, &n16);                          // :241
#ifdef CFEAR_K5_PROFILE
      const long long tq2 = clock64() + (long long)(m & 0);
      prof[13] += tq1 - tq0; prof[14] += tq2 - tq1; prof[16] += 1;
#endif
      bool valid = false;
      if (m >= 0) {
        const double ntx = R.rc * nsrc.x - R.rs * nsrc.y, nty = R.rs * nsrc.x + R.rc * nsrc.y;    // :244
        valid = normal_gate(ntx, nty, n16, P.pool.normal + (size_t)C.slots[i] * P.pool.max_cells + m, angle_outlier);   // :246-247
      }
      s_nn[u] = valid ? (uint16_t)m : K5_NONE;
      if constexpr (AUX) {                     // tables of THIS pass: pairs an earlier pass accepted and this one does not are cleared
        if (assoc) assoc[(size_t)i * P.pool.max_cells + j] = valid ? m : -1;
        if (assoc_sim) assoc_sim[(size_t)i * P.pool.max_cells + j] = 0.0;
      }
#ifdef CFEAR_K5_PROFILE
      prof[15] += clock64() - tq2;
#endif
    }
    PROF_T(tp1);
    __syncthreads();
    PROF_T(tp2);
    // ---- positions: contiguous chunk per thread, one block scan ----
    const int chunk = (nt + T - 1) / T;
    const int lo = min(tid * chunk, nt), hi = min(lo + chunk, nt);
    int cnt = 0;
    for (int u = lo; u < hi; ++u) cnt += (s_nn[u] != K5_NONE);
    int total;
    int base = block_excl_scan(cnt, s_warp, &total);
    for (int u = lo; u < hi; ++u) if (s_nn[u] != K5_NONE) s_list[base++] = (uint16_t)u;
    __syncthreads();
    PROF_T(tp3);
    // ---- phase 2 ----
    for (int r0 = tid; r0 < total; r0 += 2 * T) {
      int ii[2] = {0, 0}, jj[2] = {0, 0};
      bool have[2];
      double2 mu[2], nsrc[2], tm[2], ntar[2];
      double4 Cv[2];
      double n1[2] = {0, 0}, n2[2] = {0, 0}, p1[2] = {0, 0}, p2[2] = {0, 0};
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int r = r0 + h * T;
        have[h] = r < total;
        const int u = have[h] ? s_list[r] : s_list[r0];
        const int t = t0 + u, i = t / n_src, j = t - i * n_src;
        ii[h] = i; jj[h] = j;
        const size_t sb = C.sbase + j;
        const size_t tb = (size_t)C.slots[i] * P.pool.max_cells + s_nn[u];
        // everything the record may need is requested at once: one L2 round trip for both pairs
        mu[h] = C.src_mean[j]; nsrc[h] = C.src_normal[j];
        tm[h] = P.pool.mean[tb]; ntar[h] = P.pool.normal[tb];
        Cv[h] = make_double4(0, 0, 0, 0);
        if constexpr (COST == 2) Cv[h] = P.pool.cov[tb];
        if (P.weight_opt == 1 || P.weight_opt == 4) { n1[h] = (double)P.pool.nsamples[sb]; n2[h] = (double)P.pool.nsamples[tb]; }
        if (P.weight_opt == 3 || P.weight_opt == 4) { p1[h] = P.pool.planarity[sb]; p2[h] = P.pool.planarity[tb]; }
      }
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int r = r0 + h * T;
        if (have[h]) {
          const RelT R = rel_transform(C, ii[h]);
          const double ntx = R.rc * nsrc[h].x - R.rs * nsrc[h].y, nty = R.rs * nsrc[h].x + R.rc * nsrc[h].y;
          const double sim = fmax(ntx * ntar[h].x + nty * ntar[h].y, 0.0);
          if constexpr (AUX) { if (assoc_sim) assoc_sim[(size_t)ii[h] * P.pool.max_cells + jj[h]] = sim; }
          double2 rp, rq, rab, rcw;
          make_record<COST>(P, R, sim, mu[h], tm[h], ntar[h], Cv[h], n1[h], n2[h], p1[h], p2[h], rp, rq, rab, rcw);
          const int pos = nres + r;
          (void)rp;                                    // the source mean is re-read through its index (ResList)
          if (pos < cap) res.st(pos, jj[h], rq, rab, rcw);
        }
      }
    }
    nres += total;
    PROF_T(tp4);
    __syncthreads();                 // s_nn / s_list are reused by the next tile; residual list visible to the block
    PROF_T(tp5);
    PROF_ADD(prof[8], tp0, tp1); PROF_ADD(prof[9], tp1, tp2); PROF_ADD(prof[10], tp2, tp3); PROF_ADD(prof[11], tp3, tp4); PROF_ADD(prof[12], tp4, tp5);
  }
  return min(nres, cap);
}

// Keyframe NN indices (bucket starts + packed fp32 means) -> shared memory by TMA bulk copies, keyframe by keyframe
// while they fit in [base, base + room); the rest is read through L1/L2.  Thread 0 plans the views once (plan = true)
// and issues the copies; every thread then waits on the mbarrier's current phase.
__device__ __forceinline__ void stage_grids(const RegParams& P, const int32_t* s_slots, const NNGrid* s_grid, GridView* s_view,
                                            unsigned char* base, uint32_t room, int K, uint64_t* bar, uint32_t phase,
                                            bool plan, uint32_t* used) {
  if (threadIdx.x == 0) {
    uint32_t off = 0, tx = 0;
    for (int i = 0; i < K; ++i) {
      const int sl = s_slots[i];
      const NNGrid G = s_grid[i];
      const int n = P.pool.ncells[sl];
      const uint32_t gs_bytes = (uint32_t)(((G.nx * G.ny + 1) * 2 + 15) & ~15);
      const uint32_t gp_bytes = (uint32_t)n * 16u;
      if (off + gs_bytes + gp_bytes <= room) {
        if (plan) {
          GridView V;
          V.gs = reinterpret_cast<const uint16_t*>(base + off);
          V.gp = reinterpret_cast<const float4*>(base + off + gs_bytes);
          s_view[i] = V;
        }
        off += gs_bytes + gp_bytes; tx += gs_bytes + gp_bytes;
      } else if (plan) {
        GridView V;
        V.gs = P.pool.gstart + (size_t)sl * P.pool.grid_stride;
        V.gp = P.pool.gpt + (size_t)sl * P.pool.max_cells;
        s_view[i] = V;
      }
    }
    if (plan) *used = off;
    mbar_expect_tx(bar, tx);
    off = 0;
    for (int i = 0; i < K; ++i) {
      const int sl = s_slots[i];
      const NNGrid G = s_grid[i];
      const int n = P.pool.ncells[sl];
      const uint32_t gs_bytes = (uint32_t)(((G.nx * G.ny + 1) * 2 + 15) & ~15);
      const uint32_t gp_bytes = (uint32_t)n * 16u;
      if (off + gs_bytes + gp_bytes <= room) {
        tma_bulk_g2s(base + off, P.pool.gstart + (size_t)sl * P.pool.grid_stride, gs_bytes, bar);
        if (gp_bytes) tma_bulk_g2s(base + off + gs_bytes, P.pool.gpt + (size_t)sl * P.pool.max_cells, gp_bytes, bar);
        off += gs_bytes + gp_bytes;
      }
    }
  }
  __syncthreads();
  mbar_wait(bar, phase);
}

// Orders this thread's earlier generic-proxy accesses to shared memory before later async-proxy (TMA) writes to it.
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// AUX = false: the ceres_lm solver only (the product path).  AUX = true: the two auxiliary modes -- gn_fixed and cost
// only -- which live in their own instantiation so that their code does not cost the main kernel registers.
template <int COST, int LOSS, bool AUX, int NT = K5_THREADS>
__global__ void __launch_bounds__(NT, NT == K5_THREADS ? CFEAR_K5_MINBLOCKS : (NT == K5_THREADS_MID ? 2 : 1)) k5_register(const RegParams P) {
  extern __shared__ __align__(128) unsigned char dyn_smem[];
  __shared__ int s_warp[33];
  __shared__ double s_part[(NT / 32) * 10];
  __shared__ LMShared s_lm;
  __shared__ __align__(8) uint64_t s_bar;
  __shared__ uint32_t s_grid_bytes;

  const int prob = blockIdx.x;
  const int stride = P.nscans;
  const int ns = P.nscans_pp ? P.nscans_pp[prob] : P.nscans, K = ns - 1;
  const int tid = threadIdx.x;
  double* poses = P.poses + (size_t)prob * stride * 3;
  if (ns < 2) {                      // nothing to register against (first scan of a sequence): Register() would assert
    if (tid == 0) {
      RegStatsDev st; st.success = 0; st.outer_iterations = 0; st.inner_iterations = 0; st.num_residuals = 0; st.num_blocks = 0;
      st.usable = 0; st.final_cost = 0.0; st.score = 0.0; st.pose_written = 0; st.reserved = 0;
      reinterpret_cast<RegStatsDev*>(P.stats)[prob] = st;
      for (int i = 0; i < 36; ++i) P.cov36[(size_t)prob * 36 + i] = 0.0;
    }
    return;
  }
  // per-scan tables (pose + cos / sin, NN grid header, grid view, slot) at the start of the dynamic allocation, sized by
  // the problem's scan count (a static array for the 65 scans the ABI allows would take 5.4 KB from every CTA)
  double* s_pose = reinterpret_cast<double*>(dyn_smem);                         // [ns][5]
  NNGrid* s_grid = reinterpret_cast<NNGrid*>(s_pose + 5 * ns);                  // [ns]
  GridView* s_view = reinterpret_cast<GridView*>(s_grid + ns);                  // [ns]
  int32_t* s_slots = reinterpret_cast<int32_t*>(s_view + ns);                   // [ns]
  const uint32_t tab_bytes = (uint32_t)((ns * (40 + sizeof(NNGrid) + sizeof(GridView) + 4) + 127) & ~127u);
  if (tid < ns) {
    const int sl = P.slots[(size_t)prob * stride + tid];
    s_slots[tid] = sl;
    s_grid[tid] = P.pool.grid[sl];
    const double yaw = poses[3 * tid + 2];
    double s, c; sincos(yaw, &s, &c);
    s_pose[5 * tid + 0] = poses[3 * tid + 0]; s_pose[5 * tid + 1] = poses[3 * tid + 1];
    s_pose[5 * tid + 2] = yaw; s_pose[5 * tid + 3] = c; s_pose[5 * tid + 4] = s;
  }
  if (tid == 0) {
    mbar_init(&s_bar, 1);
    // the soft prior is only set up by the ceres_lm branch; gn_fixed's covariance round (eval_round<.., AUX>) must not see
    // whatever the previous CTA left in this shared-memory word
    if constexpr (AUX) s_lm.soft = 0;
  }
  __syncthreads();

  // Shared-memory plan:  [ tables | s_nn | s_list | source means, normals | U ]  with U = the rest of the dynamic allocation.
  //  * problems whose pairs fit one tile (the normal case) OVERLAY U: during association it holds the keyframes' NN
  //    grids, during the LM solve the residual list (written there directly by phase 2 of the association); the grids
  //    are re-staged from L2 by TMA at the start of the next outer iteration.  Every evaluation of the solve -- the
  //    inner loop of the whole kernel -- then reads shared memory only.
  //  * larger problems keep the grids resident in U and put the residual list in what is left, overflowing to the
  //    problem's global scratch.
  AssocCtx C;
  C.slots = s_slots; C.pose = s_pose; C.grid = s_grid; C.view = s_view; C.K = K;
  const int src_slot = s_slots[K];
  C.n_src = P.pool.ncells[src_slot];
  C.sbase = (size_t)src_slot * P.pool.max_cells;
  const int npairs = K * C.n_src;
  const int tile = min(max((npairs + 7) & ~7, 8), K5_TILE_MAX);
  const bool overlay = npairs <= K5_TILE_MAX;
  uint16_t* s_nn = reinterpret_cast<uint16_t*>(dyn_smem + tab_bytes);
  uint16_t* s_list = s_nn + tile;
  uint32_t u_off = tab_bytes + (uint32_t)((tile * 4 + 127) & ~127);
  // the source cells' means and normals (read by every pair of every association pass) are copied to shared memory
  // once when they take at most a quarter of the allocation
  C.src_mean = P.pool.mean + C.sbase; C.src_normal = P.pool.normal + C.sbase;
  if ((uint32_t)C.n_src * 32u <= (uint32_t)P.smem_bytes / 4) {
    double2* sm = reinterpret_cast<double2*>(dyn_smem + u_off);
    double2* sn = sm + C.n_src;
    for (int j = tid; j < C.n_src; j += NT) { sm[j] = C.src_mean[j]; sn[j] = C.src_normal[j]; }
    C.src_mean = sm; C.src_normal = sn;
    u_off += (uint32_t)((C.n_src * 32 + 127) & ~127);
  }
  unsigned char* U = dyn_smem + u_off;
  const uint32_t u_room = (uint32_t)P.smem_bytes - u_off;
  uint32_t bar_phase = 0;
  stage_grids(P, s_slots, s_grid, s_view, U, u_room, K, &s_bar, bar_phase, true, &s_grid_bytes);
  bar_phase ^= 1;
  bool grids_resident = true;

#ifdef CFEAR_K5_PROFILE
  long long prof[20] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
  // 0 association 1 solve 2 grid staging | warp 0, per evaluation: 4 sincos+publish 5 own share 6 wait for the others 7 count |
  // association: 8 phase 1 (own work) 9 wait 10 scan + list 11 phase 2 (own work) 12 wait
  const long long tk0 = clock64();
#endif
  double x[3] = {s_pose[5 * K + 0], s_pose[5 * K + 1], s_pose[5 * K + 2]};
  ResList res;
  res.g = P.res + (size_t)prob * P.res_cap * 4; res.cap_g = P.res_cap;
  res.gj = reinterpret_cast<uint16_t*>(res.g + 3 * (size_t)P.res_cap);          // the fourth field's storage holds the source indices
  res.src_mean = C.src_mean;
  {
    unsigned char* lbase = U; int room = (int)u_room;
    if (!overlay) { const uint32_t g_end = (s_grid_bytes + 127) & ~127u; lbase = U + g_end; room = (int)u_room - (int)g_end; }
    res.cap_s = max(min((room / 50) & ~7, P.res_cap), 0);                       // 3 x 16 B + 2 B per record, field arrays 16-byte aligned
    res.s = reinterpret_cast<double2*>(lbase);
    res.sj = reinterpret_cast<uint16_t*>(res.s + 3 * res.cap_s);
  }
  int32_t* assoc = (AUX && P.assoc) ? P.assoc + (size_t)prob * (stride - 1) * P.pool.max_cells : nullptr;
  double* assoc_sim = (AUX && P.assoc_sim) ? P.assoc_sim + (size_t)prob * (stride - 1) * P.pool.max_cells : nullptr;
  constexpr int per_block = (COST == 1) ? 1 : 2;
  const bool w0 = warp_id() == 0;              // the warp that runs the scalar solver logic (see eval_round / lm_solve)
  LMShared* sh = &s_lm;

  // association at pose x for outer iteration itr (radius doubled on the first, n_scan_normal.cpp:222)
  auto associate = [&](int itr) -> int {
    PROF_T(ta0);
    if (overlay && !grids_resident) {
      fence_proxy_async_smem();                // our reads of the residual list precede the TMA writes over it
      __syncthreads();
      stage_grids(P, s_slots, s_grid, s_view, U, u_room, K, &s_bar, bar_phase, false, nullptr);
      bar_phase ^= 1;
    }
    PROF_T(ta1);
    C.x[0] = x[0]; C.x[1] = x[1]; C.x[2] = x[2];
    sincos(x[2], &C.sn_s, &C.cs_s);
    C.radius = (itr == 1) ? 2 * P.radius : P.radius;
    C.nnr = nn_radius(C.radius);
    const int n = build_problem<COST, AUX, NT>(P, C, res, assoc, assoc_sim, s_warp, s_nn, s_list, tile PROF_ARG);
    if (overlay) grids_resident = false;
    PROF_T(ta2);
    PROF_ADD(prof[2], ta0, ta1); PROF_ADD(prof[0], ta1, ta2);
    return n;
  };

  SolveSum sum; sum.final_cost = 0.0; sum.n_iterations = 0; sum.last_rel = 0.0; sum.usable = true;
  bool have_H = false;                         // LMShared::acc_H holds J^T J at x (block-uniform)
  bool success = true, pose_written = false;
  int inner_total = 0, nres = 0, outer = 0;
  if (AUX && P.solver_mode == 2) {
    // cost only -- n_scan_normal_reg::GetCost (n_scan_normal.cpp:187-213): one association at the registration radius
    // (itr_ is past 1 whenever GetCost runs after a Register) and one evaluation of 1/2 sum w rho(s); no solve, the
    // pose is returned untouched.  Used by the fuser's covariance-by-sampling (odometrykeyframefuser.cpp:261-380).
    nres = associate(2);
    outer = 0;
    if (nres * per_block <= 1) success = false;                                 // :205-208
    else {
      EvalOut ev;
      eval_round<COST, LOSS, false, NT>(P.loss_limit, res, nres, w0, false, x, ev, sh, s_part PROF_ARG);
      if (tid == 0) sh->out_final_cost = ev.cost;
      __syncthreads();
      sum.final_cost = sh->out_final_cost;
      __syncthreads();
    }
  } else if (AUX && P.solver_mode == 1) {
    // gn_fixed: N undamped Gauss-Newton / IRLS iterations, re-associating before each one
    int it;
    for (it = 1; it <= P.gn_iters; ++it) {
      nres = associate(it);
      if (nres * per_block <= 1) { success = false; break; }
      {
        EvalOut ev;
        eval_round<COST, LOSS, false, NT>(P.loss_limit, res, nres, w0, false, x, ev, sh, s_part PROF_ARG);
        if (w0) {
          double y[3]; const double nb[3] = {-ev.g[0], -ev.g[1], -ev.g[2]};
          const bool okw = chol3_solve(ev.H, nb, y);
          if (lane_id() == 0) {
            sh->out_success = okw ? 1 : 0; sh->out_final_cost = ev.cost;
            sh->out_x[0] = x[0] + (okw ? y[0] : 0.0); sh->out_x[1] = x[1] + (okw ? y[1] : 0.0); sh->out_x[2] = x[2] + (okw ? y[2] : 0.0);
          }
        }
        __syncthreads();
      }
      const bool ok = sh->out_success != 0;
      x[0] = sh->out_x[0]; x[1] = sh->out_x[1]; x[2] = sh->out_x[2];
      sum.final_cost = sh->out_final_cost;
      __syncthreads();                           // outputs consumed before the next episode rewrites them
      if (!ok) { success = false; break; }
      pose_written = true;
      inner_total++;
    }
    outer = it;
    if (success) {
      EvalOut ev;
      eval_round<COST, LOSS, false, NT>(P.loss_limit, res, nres, w0, false, x, ev, sh, s_part PROF_ARG);
      if (tid == 0) sh->out_final_cost = ev.cost;
      __syncthreads();
      sum.final_cost = sh->out_final_cost;
      __syncthreads();
    }
  } else {
    double prev_par[3] = {x[0], x[1], x[2]};
    double prev_score = 1.7976931348623157e308;
    int itr;
    have_H = true;
    if constexpr (AUX) {
      if (tid == 0) {
        sh->soft = P.soft_L != nullptr;
        if (P.soft_L) {                                                         // :373-377
          const double* Lr = P.soft_L + (size_t)prob * 9;
          const double alpha = sqrt((double)C.n_src);
          const double l00 = Lr[0], l10 = Lr[3], l11 = Lr[4], l20 = Lr[6], l21 = Lr[7], l22 = Lr[8];
          sh->guess[0] = x[0]; sh->guess[1] = x[1]; sh->guess[2] = x[2];
          sh->aL[0] = alpha * l00; sh->aL[1] = alpha * l10; sh->aL[2] = alpha * l11;
          sh->aL[3] = alpha * l20; sh->aL[4] = alpha * l21; sh->aL[5] = alpha * l22;
          const double a2 = alpha * alpha;                                      // M = alpha^2 L^T L  (xx xy xt yy yt tt)
          sh->M[0] = a2 * (l00 * l00 + l10 * l10 + l20 * l20); sh->M[1] = a2 * (l10 * l11 + l20 * l21); sh->M[2] = a2 * (l20 * l22);
          sh->M[3] = a2 * (l11 * l11 + l21 * l21); sh->M[4] = a2 * (l21 * l22); sh->M[5] = a2 * (l22 * l22);
        }
      }
      __syncthreads();
    }
    for (itr = 1; itr <= P.max_outer && success; ++itr) {                       // n_scan_normal.cpp:102
      nres = associate(itr);
      if (nres * per_block <= 1) { success = false; break; }                    // :370, :114
      PROF_T(ts0);
      const double x_in[3] = {x[0], x[1], x[2]};
      lm_solve<COST, LOSS, AUX, NT>(P, res, nres, w0, x, sum, sh, s_part PROF_ARG);          // :117
      if (tid == 0) {
        sh->out_x[0] = x[0]; sh->out_x[1] = x[1]; sh->out_x[2] = x[2];
        sh->out_final_cost = sum.final_cost; sh->out_last_rel = sum.last_rel;
        sh->out_niter = sum.n_iterations; sh->out_usable = sum.usable ? 1 : 0;
      }
      __syncthreads();
      x[0] = sh->out_x[0]; x[1] = sh->out_x[1]; x[2] = sh->out_x[2];
      sum.final_cost = sh->out_final_cost; sum.last_rel = sh->out_last_rel;
      sum.n_iterations = sh->out_niter; sum.usable = sh->out_usable != 0;
      __syncthreads();                           // outputs consumed before the next episode rewrites them
      PROF_T(ts1);
      PROF_ADD(prof[1], ts0, ts1);
      success = sum.usable;
      // an unusable solution is not written back (ceres restores the parameter blocks): the pose of the last good solve stays
      if (!success) { x[0] = x_in[0]; x[1] = x_in[1]; x[2] = x_in[2]; }
      else pose_written = true;                                                 // :119-121
      inner_total += sum.n_iterations - 1;
      const double current_score = sum.final_cost;
      const double rel_improvement = (prev_score - current_score) / prev_score;
      if (itr > P.min_outer) {                                                  // :134-149
        if (prev_score < current_score) { x[0] = prev_par[0]; x[1] = prev_par[1]; x[2] = prev_par[2]; have_H = false; break; }
        else if (rel_improvement < 0.00001) break;
        else if (sum.last_rel < 0.00001 || sum.n_iterations == 1) break;
      }
      prev_score = current_score;
      prev_par[0] = x[0]; prev_par[1] = x[1]; prev_par[2] = x[2];
    }
    outer = itr;
  }

  RegStatsDev st;
  st.outer_iterations = outer; st.inner_iterations = inner_total;
  st.num_blocks = nres; st.num_residuals = nres * per_block;
  if constexpr (AUX) { if (P.solver_mode == 0 && P.soft_L && nres * per_block > 1) { st.num_blocks += 1; st.num_residuals += 3; } }   // the prior's block
  st.usable = sum.usable ? 1 : 0; st.final_cost = sum.final_cost; st.score = 0.0; st.success = 0;
  st.pose_written = pose_written ? 1 : 0; st.reserved = 0;
  double cov[36];
#pragma unroll
  for (int i = 0; i < 36; ++i) cov[i] = 0.0;
  if (AUX && success && P.solver_mode == 2) {
    st.score = sum.final_cost / st.num_residuals;                               // score_ = score / max(#residuals, 1)  :211
    st.success = 1;
  } else if (success) {
    st.score = sum.final_cost / st.num_residuals;                               // :166
    cov[0] = 0.01; cov[7] = 0.01; cov[35] = 0.0001;                             // :171-175
    EvalOut ev;                                                                 // GetCovariance :392-433
    if (have_H) {                              // the last solve ended at x: its normal equations are the ones wanted
#pragma unroll
      for (int i = 0; i < 6; ++i) ev.H[i] = sh->acc_H[i];
    } else {                                   // (block-uniform) one more round at the restored pose
      eval_round<COST, LOSS, AUX, NT>(P.loss_limit, res, nres, w0, false, x, ev, sh, s_part PROF_ARG);
    }
    if (w0) {
      double inv[9]; bool ok = true;
      for (int c = 0; c < 3 && ok; ++c) {
        double e[3] = {0, 0, 0}, y[3]; e[c] = 1.0;
        ok = chol3_solve(ev.H, e, y);
        inv[0 + c] = y[0]; inv[3 + c] = y[1]; inv[6 + c] = y[2];
      }
      if (ok && st.num_residuals - 3 != 0) {
        const double f = 30 * (sum.final_cost / (st.num_residuals - 3));        // :418
#pragma unroll
        for (int i = 0; i < 36; ++i) cov[i] = 0.0;
        for (int i = 0; i < 6; ++i) cov[i * 6 + i] = 1.0;
        cov[0] = f * inv[0]; cov[1] = f * inv[1]; cov[6] = f * inv[3]; cov[7] = f * inv[4];
        cov[35] = f * inv[8]; cov[5] = f * inv[2]; cov[30] = f * inv[6];
        st.success = 1;
      }
    }
  }
  if (tid == 0) {
    poses[3 * K + 0] = x[0]; poses[3 * K + 1] = x[1]; poses[3 * K + 2] = x[2];
    reinterpret_cast<RegStatsDev*>(P.stats)[prob] = st;
    double* c36 = P.cov36 + (size_t)prob * 36;
    for (int i = 0; i < 36; ++i) c36[i] = cov[i];
#ifdef CFEAR_K5_PROFILE
    c36[13] = (double)prof[0]; c36[14] = (double)prof[1]; c36[16] = (double)prof[2]; c36[15] = (double)(clock64() - tk0);
    c36[8] = (double)prof[4]; c36[9] = (double)prof[5]; c36[10] = (double)prof[6]; c36[11] = (double)prof[7];
    c36[23] = (double)prof[13]; c36[24] = (double)prof[14]; c36[25] = (double)prof[15]; c36[26] = (double)prof[16];
    c36[17] = (double)prof[8]; c36[18] = (double)prof[9]; c36[19] = (double)prof[10]; c36[20] = (double)prof[11]; c36[22] = (double)prof[12];
#endif
  }
}

}  // namespace cfear
