// K1: k-strongest filter + polar->Cartesian cloud rows.
// Replaces StructuredKStrongest::FilterKstrongest (radar_filters.cpp:209-237) and
// getPeaksFilteredPointCloud (radar_filters.cpp:309-337) of the reference.
//
// One warp per azimuth row.  The row (R uint8, 3360 B on Navtech data) is streamed once from HBM
// with 16-byte non-allocating loads (7 in flight per lane), tested against z_min with SWAR byte
// compares, the (sparse) candidates are emitted warp-cooperatively as keys (intensity<<16 | range) into
// a per-warp shared-memory list, and the k largest keys are selected exactly:
//   lexicographic (intensity, range) order == integer order of the key, so ties go to the larger
//   range bin exactly like the reference's sorted-insert / erase-front loop.
// Algorithmic HBM bytes per row: R (read) + 4k+4 (indices) + 16k+4 (cloud row).
#pragma once
#include <algorithm>
#include "common.cuh"

namespace cfear {

#ifndef CFEAR_K1_MINBLOCKS
#define CFEAR_K1_MINBLOCKS 4   // resident CTAs per SM the register budget is set for: 64 registers (the next row's vectors stay in registers while a row is finished)
#endif
#ifndef CFEAR_K1_WARPS
#define CFEAR_K1_WARPS 8
#endif
constexpr int K1_WARPS = CFEAR_K1_WARPS;       // warps per CTA
#ifndef CFEAR_K1_RPW
#define CFEAR_K1_RPW 8
#endif
constexpr int K1_RPW = CFEAR_K1_RPW;   // most rows per warp (the loads of row i+1 fly while row i is finished); k1_launch picks K1Params::rpw
constexpr int K1_CAP = 256;       // candidate keys per warp kept in shared memory
constexpr int K1_TILES = 7;       // uint4 per lane per super-tile (7*512 B = 3584 B >= one 3360-bin Navtech row in registers);
constexpr int K1_TILES_WIDE = 8;  // rows of up to 4096 bytes (Oxford: 3768 bins) in one super-tile
constexpr int K1_MAXK = 64;       // k_strongest <= 64

struct K1Params {
  const uint8_t* polar;      // [nrows][R]
  const uint8_t* polar_end;  // polar + nrows*R
  int nrows;                 // nscans * A
  int A, R;
  uint32_t A_magic;          // ceil(2^32 / A): row -> azimuth without a division (k1_launch fills it)
  int rpw;                   // rows per warp, 1..K1_RPW (k1_launch fills it)
  uint32_t one;              // 1, unknown to the compiler (k1_launch fills it): x * one + c is an IMAD on the FMA pipe
  int zmin;                  // already uchar(int(z_min))
  int k;
  int min_range_bin;         // ceil(min_distance / range_res)
  double range_res;
  const double2* cs;         // [A] (cos theta, sin theta), theta = 2 pi (a+1)/A, host libm
  int32_t* kidx;             // [nrows][k]
  int32_t* kcnt;             // [nrows]
  float4* rowcloud;          // [nrows][k]  points passing the min-range cut, ascending (intensity, range)
  int32_t* rowcnt;           // [nrows]
};

__device__ __forceinline__ uint4 ld_stream16(const uint8_t* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}

// 16 bytes at p (16-B aligned); bytes outside [lo, hi) read as 0 without touching memory.
__device__ __forceinline__ uint4 load16_guarded(const uint8_t* p, const uint8_t* lo, const uint8_t* hi) {
  if (p >= lo && p + 16 <= hi) return ld_stream16(p);
  uint32_t w[4] = {0u, 0u, 0u, 0u};
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const uint8_t* q = p + i;
    if (q >= lo && q < hi) w[i >> 2] |= (uint32_t)(*q) << (8 * (i & 3));
  }
  return make_uint4(w[0], w[1], w[2], w[3]);
}

// Test of one 16-byte vector: does ANY byte reach z_min (z_min >= 1; z_min == 0 is handled by the caller)?
// z_min < 128: bit 7 of (x + c) | x per byte with c = 0x80 - z_min.  A byte >= 128 passes through the "| x"; below that
// byte + c cannot overflow, so bit 7 says byte >= z_min.  The low 7 bits are NOT masked off first, so a carry out of the
// byte below (inside the 32-bit word) may add one and let a byte equal to z_min - 1 pass -- but a byte only carries when
// it is >= 128 + z_min, i.e. a candidate itself, so the answer for the vector is exact; which bytes are the candidates is
// decided by the per-byte test of the drain.  Two instructions per word (one add, one three-input OR) instead of three
// plus the merges of an exact packed flag word.
// ZHI (z_min >= 128): bit 7 of x & ((x & 0x7f..) + c), exact.
template <bool ZHI>
__device__ __forceinline__ bool any_maybe_ge(const uint4& d, uint32_t addc, uint32_t one) {
  uint32_t m;
  if (ZHI) {
    m = d.x & ((d.x & 0x7f7f7f7fu) + addc);
    m |= d.y & ((d.y & 0x7f7f7f7fu) + addc);
    m |= d.z & ((d.z & 0x7f7f7f7fu) + addc);
    m |= d.w & ((d.w & 0x7f7f7f7fu) + addc);
  } else {
#ifdef CFEAR_K1_IMAD_ADD
    // the adds as multiply-adds by an opaque 1: IMAD runs on the FMA pipe, which this kernel leaves idle
    m = (d.x * one + addc) | d.x;
    m |= (d.y * one + addc) | d.y;
    m |= (d.z * one + addc) | d.z;
    m |= (d.w * one + addc) | d.w;
#else
    m = (d.x + addc) | d.x;
    m |= (d.y + addc) | d.y;
    m |= (d.z + addc) | d.z;
    m |= (d.w + addc) | d.w;
#endif
  }
  return (m & 0x80808080u) != 0u;
}

constexpr int K1_QCAP = 96;       // vectors queued per warp before a drain is forced (a drain is due once more than QCAP - 32 wait)

// Per-warp shared memory, one block so that a single 32-bit base address (kept in a register) reaches every part at an
// immediate offset: the queue of vectors that may hold candidates (+ a sentinel pair), the range bin of byte 0 of each,
// the candidate keys and the two small lists of the selection.
struct __align__(16) K1Warp {
  uint4 qvec[K1_QCAP + 2];
  int qbin[K1_QCAP + 2];
  uint32_t cand[K1_CAP];
  uint32_t sel[K1_MAXK];
  uint32_t out[K1_MAXK];
};
constexpr uint32_t K1_QBIN = (K1_QCAP + 2) * 16;            // byte offsets inside K1Warp
constexpr uint32_t K1_CAND = K1_QBIN + (K1_QCAP + 2) * 4;

__device__ __forceinline__ void sts128(uint32_t a, const uint4& v) {
  asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" :: "r"(a), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
template <uint32_t OFF>
__device__ __forceinline__ void sts32(uint32_t a, uint32_t v) { asm volatile("st.shared.u32 [%0+%2], %1;" :: "r"(a), "r"(v), "n"(OFF) : "memory"); }
template <uint32_t OFF>
__device__ __forceinline__ uint32_t lds32(uint32_t a) { uint32_t v; asm volatile("ld.shared.u32 %0, [%1+%2];" : "=r"(v) : "r"(a), "n"(OFF) : "memory"); return v; }
__device__ __forceinline__ uint32_t lds8(uint32_t a) { uint32_t v; asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(a) : "memory"); return v; }

// Drain the queue: two queued vectors per step, lane l tests byte (l & 15) of the first (l < 16) or second (l >= 16)
// exactly (byte >= z_min and the byte's range bin inside the row) and the accepted bytes are appended to the candidate
// list as keys (intensity << 16 | range) by one ballot.  Entry nq is a sentinel whose bins lie outside the row.
// wb: shared-space address of the warp's K1Warp.
__device__ __forceinline__ void k1_drain(uint32_t wb, int nq, int& C, int zmin, int R, int lane, uint32_t lanes_below) {
  if (lane == 0) sts32<K1_QBIN>(wb + 4u * (uint32_t)nq, (uint32_t)R);
  __syncwarp();
  uint32_t a_byte = wb + (uint32_t)lane;                       // byte (lane & 15) of entry e + (lane >> 4)
  uint32_t a_bin = wb + 4u * (uint32_t)(lane >> 4);
  const int jb = lane & 15;
  for (int e = 0; e < nq; e += 2, a_byte += 32u, a_bin += 8u) {
    const uint32_t byte = lds8(a_byte);
    const int bin = (int)lds32<K1_QBIN>(a_bin) + jb;
    const bool ok = (int)byte >= zmin && (uint32_t)bin < (uint32_t)R;
    const uint32_t bal = __ballot_sync(FULL, ok);
    const int pos = C + __popc(bal & lanes_below);
    if (ok && pos < K1_CAP) sts32<K1_CAND>(wb + 4u * (uint32_t)pos, (byte << 16) | (uint32_t)bin);
    C += __popc(bal);
  }
  __syncwarp();                                  // the queue is rewritten after this
}

// ALIGNED: every row starts on a 16-byte boundary and R % 16 == 0 (Navtech 3360-bin rows), so no vector straddles a row.
#ifndef CFEAR_K1_MINBLOCKS_UNALIGNED
#define CFEAR_K1_MINBLOCKS_UNALIGNED 4
#endif
template <bool ALIGNED, bool ZHI, int TILES = K1_TILES>
__global__ void __launch_bounds__(K1_WARPS * 32, ALIGNED ? CFEAR_K1_MINBLOCKS : CFEAR_K1_MINBLOCKS_UNALIGNED) k1_kstrongest(const K1Params p) {
  __shared__ K1Warp s_w[K1_WARPS];
  const int lane = lane_id();
  K1Warp& W = s_w[warp_id()];
  uint32_t* cand = W.cand;
  uint32_t* sel = W.sel;
  uint32_t* outk = W.out;
  uint32_t wb = (uint32_t)__cvta_generic_to_shared(&W);
  asm volatile("mov.u32 %0, %0;" : "+r"(wb));   // opaque: keep the base in a register instead of recomputing it at every use

  const int R = p.R, k = p.k;
  const uint32_t addc = (0x80u - (uint32_t)(p.zmin & 0x7f)) * 0x01010101u;
  const uint32_t lanes_below = (1u << lane) - 1u;
  // A warp takes p.rpw rows, K1_WARPS apart (the warps of a CTA read consecutive rows at every step), and requests the
  // next row's vectors as soon as the current row's have been looked at: the drain, the selection and the outputs of a
  // row run under the next row's loads.  No block-level synchronisation anywhere in this kernel.
  uint4 d[TILES];
  bool have_d = false;                         // d[] already holds the first super-tile of this row (warp-uniform)
  for (int it = 0; it < p.rpw; ++it) {
  const int grow = (blockIdx.x * p.rpw + it) * K1_WARPS + warp_id();
  if (grow >= p.nrows) break;
  const uint8_t* row = p.polar + (size_t)grow * R;
  const int off = ALIGNED ? 0 : (int)((uintptr_t)row & 15);
  const uint8_t* base = row - off;
  const int nvec = (off + R + 15) >> 4;
  const bool edge_row = grow == 0 || grow == p.nrows - 1;      // warp-uniform

  // ---- pass 1: stream the row, collect the candidates ---------------------------------------------
  // Candidates are sparse (a few vectors per row hold any).  The streaming part only decides, per 16-byte vector,
  // whether it MAY hold one (any_maybe_ge); such vectors go, as they are, onto a per-warp shared-memory queue (one
  // ballot per 32 vectors) and the queue is drained cooperatively with one lane per byte (k1_drain), which is where the
  // exact test, the row head / tail of unaligned rows and the keys are done.
  int C = 0;                                   // warp-uniform candidate count
  if (p.zmin == 0) {
    C = R;                                     // every byte is a candidate: the selection below reads the row itself
    if (R <= K1_CAP) {                         // ... unless the row is short enough for the candidate list
      for (int j = lane; j < R; j += 32) cand[j] = ((uint32_t)row[j] << 16) | (uint32_t)j;
      __syncwarp();
    }
  } else {
    int nq = 0;                                // warp-uniform queue length
    for (int v0 = 0; v0 < nvec; v0 += 32 * TILES) {
      // all tiles but the last lie inside the row and nothing reaches outside the buffer (only the first vector of the
      // first row and the last vector of the last row of an unaligned buffer can): unpredicated loads at immediate
      // offsets, the last tile at a clamped address (its lanes past the row re-read the row's last vector and are
      // dropped by the v < nvec test below)
      const bool fast = (v0 + 32 * (TILES - 1) < nvec) && (ALIGNED || !edge_row);      // warp-uniform
      if (have_d && v0 == 0) {
        // requested while the previous row was being finished
      } else if (fast) {
        const uint8_t* pl = base + 16 * (size_t)(v0 + lane);
#pragma unroll
        for (int i = 0; i < TILES - 1; ++i) d[i] = ld_stream16(pl + 512 * i);
        d[TILES - 1] = ld_stream16(base + 16 * (size_t)min(v0 + 32 * (TILES - 1) + lane, nvec - 1));
      } else {
#pragma unroll
        for (int i = 0; i < TILES; ++i) {
          const int v = v0 + i * 32 + lane;
          if (ALIGNED || !edge_row) d[i] = (v < nvec) ? ld_stream16(base + 16 * (size_t)v) : make_uint4(0, 0, 0, 0);
          else d[i] = (v < nvec) ? load16_guarded(base + 16 * (size_t)v, p.polar, p.polar_end) : make_uint4(0, 0, 0, 0);
        }
      }
#pragma unroll
      for (int i = 0; i < TILES; ++i) {
        const int v = v0 + i * 32 + lane;
        // vectors past the row hold zeros (never a candidate: z_min >= 1) except in the last tile of the fast path
        const bool has = any_maybe_ge<ZHI>(d[i], addc, p.one) && (i < TILES - 1 || v < nvec);
        const uint32_t bal = __ballot_sync(FULL, has);
        if (bal) {                                 // warp-uniform; nq <= K1_QCAP - 32 here
          if (has) {
            const uint32_t slot = (uint32_t)(nq + __popc(bal & lanes_below));
            sts128(wb + 16u * slot, d[i]);
            sts32<K1_QBIN>(wb + 4u * slot, (uint32_t)(v * 16 - off));   // range bin of byte 0 (negative in an unaligned row's head)
          }
          nq += __popc(bal);
          if (nq > K1_QCAP - 32) { k1_drain(wb, nq, C, p.zmin, R, lane, lanes_below); nq = 0; }
        }
      }
      have_d = false;
      if (it + 1 < p.rpw && v0 + 32 * TILES >= nvec) {
        // this row's last vectors have been looked at: request the next row's first super-tile (fast form only)
        const int gnext = grow + K1_WARPS;
        if (gnext < p.nrows) {
          const uint8_t* rown = row + (size_t)K1_WARPS * R;
          const int offn = ALIGNED ? 0 : (int)((uintptr_t)rown & 15);
          const uint8_t* basen = rown - offn;
          const int nvecn = (offn + R + 15) >> 4;
          if (32 * (TILES - 1) < nvecn && (ALIGNED || gnext != p.nrows - 1)) {
            const uint8_t* pl = basen + 16 * (size_t)lane;
#pragma unroll
            for (int i = 0; i < TILES - 1; ++i) d[i] = ld_stream16(pl + 512 * i);
            d[TILES - 1] = ld_stream16(basen + 16 * (size_t)min(32 * (TILES - 1) + lane, nvecn - 1));
            have_d = true;
          }
        }
      }
    }
    k1_drain(wb, nq, C, p.zmin, R, lane, lanes_below);
  }

  // ---- pass 2: exact top-k of the keys -----------------------------------------------------------
  const int kk = min(k, C);
  const uint32_t* work = cand;
  int nwork = C;
  if (C > K1_MAXK) {
    // radix select: largest T with #(key >= T) >= k.  Keys are distinct, so #(key >= T) == k.
    uint32_t T = 0;
    const bool in_smem = (C <= K1_CAP);
    for (int bit = 23; bit >= 0; --bit) {
      const uint32_t Tt = T | (1u << bit);
      int c = 0;
      if (in_smem) {
        for (int j = lane; j < C; j += 32) c += (cand[j] >= Tt);
      } else {   // candidate list overflowed (saturated row): re-read the row (L1/L2 resident)
        for (int r = lane; r < R; r += 32) c += ((((uint32_t)row[r]) << 16 | (uint32_t)r) >= Tt);
      }
      c = __reduce_add_sync(FULL, c);
      if (c >= k) T = Tt;
    }
    int nb = 0;
    const int n_iter = in_smem ? C : R;
    for (int j0 = 0; j0 < n_iter; j0 += 32) {
      const int j = j0 + lane;
      uint32_t key = 0;
      if (j < n_iter) key = in_smem ? cand[j] : (((uint32_t)row[j]) << 16 | (uint32_t)j);
      const bool s = (j < n_iter) && key >= T;
      const unsigned bal = __ballot_sync(FULL, s);
      if (s) sel[nb + __popc(bal & ((1u << lane) - 1))] = key;
      nb += __popc(bal);
    }
    __syncwarp();
    work = sel;
    nwork = k;
  }
  if (nwork <= 32) {
    // the usual case, one key per lane: bitonic sort across the warp (15 shuffle + min/max steps), ascending, empty
    // lanes (0) first, so the kk largest sit in the top kk lanes already in output order
    uint32_t key = (lane < nwork) ? work[lane] + 1u : 0u;          // +1: 0 is "absent"
    __syncwarp();
#pragma unroll
    for (int k2 = 2; k2 <= 32; k2 <<= 1) {
#pragma unroll
      for (int j = k2 >> 1; j > 0; j >>= 1) {
        const uint32_t other = __shfl_xor_sync(FULL, key, j);
        const bool up = (lane & k2) == 0, lower = (lane & j) == 0;
        key = (lower == up) ? min(key, other) : max(key, other);
      }
    }
    if (lane >= 32 - kk) outk[lane - (32 - kk)] = key - 1u;
  } else {
    // <= 64 distinct keys, two per lane: the kk largest, one warp-wide max reduction (REDUX) each, written in
    // ascending order (position kk-1-t for the t-th largest)
    uint32_t k0 = work[lane] + 1u;                                 // +1: 0 is the "taken / absent" mark
    uint32_t k1 = (lane + 32 < nwork) ? work[lane + 32] + 1u : 0u;
    __syncwarp();
    for (int t = 0; t < kk; ++t) {
      const uint32_t mx = __reduce_max_sync(FULL, max(k0, k1));
      if (k0 == mx) k0 = 0u;
      if (k1 == mx) k1 = 0u;
      if (lane == 0) outk[kk - 1 - t] = mx - 1u;
    }
  }
  __syncwarp();

  // ---- outputs: index set (parity target) + cloud row ---------------------------------------------
  int a = grow - (int)(__umulhi((uint32_t)grow, p.A_magic) * (uint32_t)p.A);     // grow % A (exact while grow * A < 2^32)
  if ((uint32_t)a >= (uint32_t)p.A) a = grow % p.A;
  const double2 cs = p.cs[a];
  const double half = p.range_res / 2.0;
  int ncloud = 0;
  for (int j0 = 0; j0 < k; j0 += 32) {
    const int j = j0 + lane;
    const bool have = j < kk;
    const uint32_t key = have ? outk[j] : 0u;
    const int r = (int)(key & 0xffffu);
    if (j < k) p.kidx[(size_t)grow * k + j] = have ? r : -1;
    const bool pass = have && r > p.min_range_bin;            // radar_filters.cpp:327
    const unsigned bal = __ballot_sync(FULL, pass);
    if (pass) {
      // (range_res/2 + range_res*range) * cos/sin in double, rounded once to float (:329-330); no FMA contraction
      const double rho = __dadd_rn(half, __dmul_rn(p.range_res, (double)r));
      float4 pt;
      pt.x = __double2float_rn(__dmul_rn(rho, cs.x));
      pt.y = __double2float_rn(__dmul_rn(rho, cs.y));
      pt.z = 0.f;
      pt.w = (float)(key >> 16);
      p.rowcloud[(size_t)grow * k + ncloud + __popc(bal & ((1u << lane) - 1))] = pt;
    }
    ncloud += __popc(bal);
  }
  if (lane == 0) { p.kcnt[grow] = kk; p.rowcnt[grow] = ncloud; }
  __syncwarp();                                // the lists are rewritten by the next row
  }
}

// Launch with the instantiation the data allows: aligned rows (every row on a 16-byte boundary, no vector straddles a
// row) and the z_min half (>= 128 or not) are compile-time.
inline void k1_launch(const K1Params& p_in, cudaStream_t stream, int prio = 0) {
  K1Params p = p_in;
  p.A_magic = (uint32_t)(0xffffffffull / (unsigned long long)p.A + 1ull);
  p.one = 1u;
  // rows per warp: as many as leave about two waves of CTAs on a B200 (148 SMs x 4 CTAs); small launches (one scan: 400
  // rows) keep one row per warp, where latency is what counts
  p.rpw = std::max(1, std::min(K1_RPW, p.nrows / (K1_WARPS * 148 * 4 * 2)));
  const int grid = (p.nrows + K1_WARPS * p.rpw - 1) / (K1_WARPS * p.rpw);
  const bool aligned = ((uintptr_t)p.polar & 15) == 0 && (p.R & 15) == 0;
  const bool zhi = p.zmin >= 128;
  const bool wide = ((15 + p.R + 15) >> 4) > 32 * K1_TILES;      // a row (plus its alignment slack) exceeds one 7-tile super-tile
  if (aligned) { if (zhi) launch_with_priority(k1_kstrongest<true, true>, grid, K1_WARPS * 32, 0, stream, prio, p); else launch_with_priority(k1_kstrongest<true, false>, grid, K1_WARPS * 32, 0, stream, prio, p); }
  else if (wide) { if (zhi) launch_with_priority(k1_kstrongest<false, true, K1_TILES_WIDE>, grid, K1_WARPS * 32, 0, stream, prio, p); else launch_with_priority(k1_kstrongest<false, false, K1_TILES_WIDE>, grid, K1_WARPS * 32, 0, stream, prio, p); }
  else { if (zhi) launch_with_priority(k1_kstrongest<false, true>, grid, K1_WARPS * 32, 0, stream, prio, p); else launch_with_priority(k1_kstrongest<false, false>, grid, K1_WARPS * 32, 0, stream, prio, p); }
}

// ------------------------------------------------------------------------------------------------
// Peaks cloud: StructuredKStrongest::AxialNonMaxSupress (radar_filters.cpp:238-298).  Off the pose
// path (the reference stores the peaks cloud but never uses it for registration); one thread per
// kept bin, scores recomputed from the image.  Reads outside the row index the flat image buffer
// like the reference's unchecked cv::Mat::at, clamped to the buffer.
// ------------------------------------------------------------------------------------------------
struct PeaksParams {
  const uint8_t* polar; long total;  // nscans*A*R bytes
  int nrows, A, R, k, min_range_bin; double range_res; const double2* cs;
  const int32_t* kidx; const int32_t* kcnt;
  float4* rowpeaks; int32_t* rowpeakcnt;   // [nrows][k], [nrows]
};

__device__ __forceinline__ int peaks_score(const PeaksParams& p, long rowbase, int r_n) {
  int s = 0;
#pragma unroll
  for (int d = -3; d <= 3; ++d) {
    long f = rowbase + r_n + d;
    f = f < 0 ? 0 : (f >= p.total ? p.total - 1 : f);
    s += p.polar[f];
  }
  return s;
}

__global__ void __launch_bounds__(256) k1b_peaks(const PeaksParams p) {
  const int grow = blockIdx.x * (blockDim.x >> 5) + warp_id();
  if (grow >= p.nrows) return;
  const int lane = lane_id();
  const int W = 3, R = p.R, k = p.k;
  const int cnt = p.kcnt[grow];
  const long rowbase = (long)grow * R;
  const int a = grow % p.A;
  const double2 cs = p.cs[a];
  const double half = p.range_res / 2.0;
  int nout = 0;
  for (int j0 = 0; j0 < k; j0 += 32) {
    const int j = j0 + lane;
    bool keep = false;
    int r = 0;
    if (j < cnt) {
      r = p.kidx[(size_t)grow * k + j];
      // score[x] exists iff x is within +-W of an in-guard kept bin (:251-263); missing keys read as 0 (:271-276)
      auto have = [&](int x) {
        for (int m = 0; m < cnt; ++m) {
          const int q = p.kidx[(size_t)grow * k + m];
          if (q >= W && q < R - W && x >= q - W && x <= q + W) return true;
        }
        return false;
      };
      auto sc = [&](int x) { return have(x) ? peaks_score(p, rowbase, x) : 0; };
      const int pthis = sc(r);
      keep = true;
      for (int i = 1; i <= W; ++i) {
        const int pnext = sc(r + i), pprev = sc(r - i);
        if (pprev > pthis || pthis < pnext) { keep = false; break; }   // :282
      }
    }
    const bool pass = keep && r > p.min_range_bin;
    const unsigned bal = __ballot_sync(FULL, pass);
    if (pass) {
      const double rho = __dadd_rn(half, __dmul_rn(p.range_res, (double)r));
      float4 pt;
      pt.x = __double2float_rn(__dmul_rn(rho, cs.x));
      pt.y = __double2float_rn(__dmul_rn(rho, cs.y));
      pt.z = 0.f;
      pt.w = (float)p.polar[rowbase + r];
      p.rowpeaks[(size_t)grow * k + nout + __popc(bal & ((1u << lane) - 1))] = pt;
    }
    nout += __popc(bal);
  }
  if (lane == 0) p.rowpeakcnt[grow] = nout;
}

}  // namespace cfear
