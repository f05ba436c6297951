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
#include "common.cuh"

namespace cfear {

#ifndef CFEAR_K1_MINBLOCKS
#define CFEAR_K1_MINBLOCKS 6   // resident CTAs per SM the register budget is set for (40 registers, no spills; 4 -> 6: 0.115 -> 0.109 ms)
#endif
constexpr int K1_WARPS = 8;       // warps (rows) per CTA
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

// bit7 of each byte set iff byte >= z_min.  addc = (0x80 - (zmin&0x7f)) * 0x01010101.  ZHI: z_min >= 128.
template <bool ZHI>
__device__ __forceinline__ uint32_t ge_flags(uint32_t x, uint32_t addc) {
  const uint32_t a = (x & 0x7f7f7f7fu) + addc;     // bit7: low 7 bits >= low 7 bits of zmin; no cross-byte carry
  return ZHI ? (x & a & 0x80808080u) : ((x | a) & 0x80808080u);
}

// flags of a uint4 packed in one word: byte b of word w -> bit 8b + 7 - w
// The kernel is bound by the integer ALU pipe (LOP3 / SHF / ISETP at half rate), so the three shift-and-or merges are
// written as multiply-high-and-add (x >> n == umulhi(x, 2^(32-n)); the flag bits of different words never collide, so
// + is |): IMAD.HI runs on the FMA pipe, which this kernel leaves idle.
template <bool ZHI>
__device__ __forceinline__ uint32_t ge_flags16(const uint4& d, uint32_t addc) {
  uint32_t m = ge_flags<ZHI>(d.x, addc);
  m = __umulhi(ge_flags<ZHI>(d.y, addc), 0x80000000u) + m;
  m = __umulhi(ge_flags<ZHI>(d.z, addc), 0x40000000u) + m;
  m = __umulhi(ge_flags<ZHI>(d.w, addc), 0x20000000u) + m;
  return m;
}

// ALIGNED: every row starts on a 16-byte boundary and R % 16 == 0 (Navtech 3360-bin rows), so no vector straddles a row.
#ifndef CFEAR_K1_MINBLOCKS_UNALIGNED
#define CFEAR_K1_MINBLOCKS_UNALIGNED 5   // the head / tail masks and the 8-tile form need 48 registers: at 40 they spill (3768-bin rows:
                                         // 0.371 ns per KB of image with 6 CTAs and spills, 0.316 with 5 CTAs; aligned 3360-bin rows: 0.282)
#endif
template <bool ALIGNED, bool ZHI, int TILES = K1_TILES>
__global__ void __launch_bounds__(K1_WARPS * 32, ALIGNED ? CFEAR_K1_MINBLOCKS : CFEAR_K1_MINBLOCKS_UNALIGNED) k1_kstrongest(const K1Params p) {
  __shared__ uint32_t s_cand[K1_WARPS][K1_CAP];
  __shared__ uint32_t s_sel[K1_WARPS][K1_MAXK];
  __shared__ uint32_t s_out[K1_WARPS][K1_MAXK];
  __shared__ uint2 s_queue[K1_WARPS][TILES * 32 + 2];  // vectors of one super-tile that hold candidates (+ a zero sentinel)
  const int grow = blockIdx.x * K1_WARPS + warp_id();
  if (grow >= p.nrows) return;                 // no block-level sync in this kernel
  const int lane = lane_id();
  uint32_t* cand = s_cand[warp_id()];
  uint32_t* sel = s_sel[warp_id()];
  uint32_t* outk = s_out[warp_id()];

  const int R = p.R, k = p.k;
  const uint8_t* row = p.polar + (size_t)grow * R;
  const int off = (int)((uintptr_t)row & 15);
  const uint8_t* base = row - off;
  const int nvec = (off + R + 15) >> 4;
  const uint32_t addc = (0x80u - (uint32_t)(p.zmin & 0x7f)) * 0x01010101u;
  const bool edge_row = grow == 0 || grow == p.nrows - 1;      // warp-uniform

  // ---- pass 1: stream the row, collect the candidates ---------------------------------------------
  // Candidates are sparse (a few vectors per row hold any), so they are emitted cooperatively: every lane whose vector
  // has flags pushes (flags, first range bin) onto a per-warp queue (one ballot per 32 vectors), then the warp drains
  // the queue two vectors at a time with lane l handling byte (l & 15) of the first (l < 16) or second (l >= 16).
  // Only range bins are emitted here; the intensities are re-read (L2) once the list is complete.
  const int jb = lane & 15, jw = jb >> 2;
  const uint32_t mybit = 1u << (8 * (jb & 3) + 7 - jw);                          // flag bit of byte jb (see ge_flags16)
  const uint32_t lowmask = 0x01010101u * (0x100u - (0x100u >> jw)) |             // flag bits of the bytes before jb
                           ((0x80808080u >> jw) & ((1u << (8 * (jb & 3))) - 1u));
  const uint32_t lanes_below = (1u << lane) - 1u;
  uint2* q = s_queue[warp_id()];
  const bool hi = lane >= 16;
  int C = 0;                                   // warp-uniform candidate count
  for (int v0 = 0; v0 < nvec; v0 += 32 * TILES) {
    uint32_t g[TILES];
    {
      uint4 d[TILES];
#pragma unroll
      for (int i = 0; i < TILES; ++i) {
        const int v = v0 + i * 32 + lane;
        // only the first vector of the first row and the last vector of the last row can reach outside the buffer;
        // tiles that lie entirely inside the row (warp-uniform test) load without a per-lane predicate
        if (ALIGNED || !edge_row) {
          if (ALIGNED && v0 + i * 32 + 32 <= nvec) d[i] = ld_stream16(base + 16 * (size_t)v);
          else d[i] = (v < nvec) ? ld_stream16(base + 16 * (size_t)v) : make_uint4(0, 0, 0, 0);
        } else d[i] = (v < nvec) ? load16_guarded(base + 16 * (size_t)v, p.polar, p.polar_end) : make_uint4(0, 0, 0, 0);
      }
#pragma unroll
      for (int i = 0; i < TILES; ++i) {
        // a super-tile past the end of the row (3768-bin Oxford rows need 236 vectors: one full super-tile of 224 and 12
        // more) skips the flag arithmetic of its empty tiles (warp-uniform test)
        if (!ALIGNED && v0 + i * 32 >= nvec) { g[i] = 0; continue; }
        const int v = v0 + i * 32 + lane;
        uint32_t m = ge_flags16<ZHI>(d[i], addc);
        const int b0 = v * 16 - off;             // range bin of byte 0 of this uint4
        if (v >= nvec) m = 0;
        else if (!ALIGNED && (b0 < 0 || b0 + 16 > R)) {        // row head / tail: drop bytes of neighbouring rows
          // valid bytes j in [jlo, jhi) -> 16-bit mask -> the packed flag layout (byte b of word w at bit 8b + 7 - w):
          // a nibble's bits go to the four byte lanes by one multiply (n * 0x00204081 puts bit b at 8b)
          const int jlo = max(0, -b0), jhi = min(16, R - b0);
          const uint32_t jm = (jhi > jlo) ? ((0xffffu >> (16 - jhi)) & (0xffffu << jlo)) : 0u;
          uint32_t keep = 0;
#pragma unroll
          for (int w = 0; w < 4; ++w) keep |= ((((jm >> (4 * w)) & 0xfu) * 0x00204081u) & 0x01010101u) << (7 - w);
          m &= keep;
        }
        g[i] = m;
      }
    }
    int nq = 0;                                // warp-uniform queue length
#pragma unroll
    for (int i = 0; i < TILES; ++i) {
      const bool has = g[i] != 0;
      const uint32_t bal = __ballot_sync(FULL, has);
      if (bal) {                                 // warp-uniform
        if (has) q[nq + __popc(bal & lanes_below)] = make_uint2(g[i], (uint32_t)((v0 + i * 32 + lane) * 16 - off));
        nq += __popc(bal);
      }
    }
    if (lane == 0) q[nq] = make_uint2(0u, 0u);   // sentinel: an odd queue drains its last entry beside an empty one
    __syncwarp();
    for (int e = 0; e < nq; e += 2) {
      const uint2 e1 = q[e];
      const uint2 e2 = q[e + 1];
      const uint32_t mm = hi ? e2.x : e1.x;
      const int c1 = __popc(e1.x);
      if (mm & mybit) {
        const int pos = C + (hi ? c1 : 0) + __popc(mm & lowmask);
        if (pos < K1_CAP) cand[pos] = (hi ? e2.y : e1.y) + (uint32_t)jb;
      }
      C += c1 + __popc(e2.x);
    }
    __syncwarp();                                // the queue is rewritten by the next super-tile
  }
  __syncwarp();
  if (C <= K1_CAP) {                             // range bins -> keys (intensity << 16 | range)
    for (int j = lane; j < C; j += 32) { const uint32_t r = cand[j]; cand[j] = ((uint32_t)row[r] << 16) | r; }
  }
  __syncwarp();

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
}

// Launch with the instantiation the data allows: aligned rows (every row on a 16-byte boundary, no vector straddles a
// row) and the z_min half (>= 128 or not) are compile-time.
inline void k1_launch(const K1Params& p_in, cudaStream_t stream) {
  K1Params p = p_in;
  p.A_magic = (uint32_t)(0xffffffffull / (unsigned long long)p.A + 1ull);
  const int grid = (p.nrows + K1_WARPS - 1) / K1_WARPS;
  const bool aligned = ((uintptr_t)p.polar & 15) == 0 && (p.R & 15) == 0;
  const bool zhi = p.zmin >= 128;
  const bool wide = ((15 + p.R + 15) >> 4) > 32 * K1_TILES;      // a row (plus its alignment slack) exceeds one 7-tile super-tile
  if (aligned) { if (zhi) k1_kstrongest<true, true><<<grid, K1_WARPS * 32, 0, stream>>>(p); else k1_kstrongest<true, false><<<grid, K1_WARPS * 32, 0, stream>>>(p); }
  else if (wide) { if (zhi) k1_kstrongest<false, true, K1_TILES_WIDE><<<grid, K1_WARPS * 32, 0, stream>>>(p); else k1_kstrongest<false, false, K1_TILES_WIDE><<<grid, K1_WARPS * 32, 0, stream>>>(p); }
  else { if (zhi) k1_kstrongest<false, true><<<grid, K1_WARPS * 32, 0, stream>>>(p); else k1_kstrongest<false, false><<<grid, K1_WARPS * 32, 0, stream>>>(p); }
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
