// bsq_k1h.cuh -- K1h: STD LOCAL for SHORT read pairs with small integer scores,
// two pairs per lane in packed 16-bit arithmetic (sm_100a DPX .16x2 forms).
//
// Same mapping as K1 (a group of 8 lanes owns a table of up to 152 columns,
// lane l owns 19 consecutive columns -- or 16 lanes x 16 columns for up to 256 --
// walks the rows, lags its left neighbour by one row) but every register holds the SAME cell of TWO problems
// -- problem A in the low halfword, problem B in the high one -- so that
// one VIADD.16x2 / VIMNMX3.S16x2.RELU / VIADDMNMX.U16x2 instruction advances two
// DP cells.  The recurrence is pwlib's (see bsq_kernels.cuh): candidates B, D, I,
// M in that order, every candidate tying for the maximum kept, a gap extends for
// free only out of a cell whose tied-max set holds the same gap op.
//
//   state per column : G2 = H + ge  (both problems), nD2 = "D not in the tie set"
//   D2  = G2[up]   + nD2[up]   * go          (one IMAD: halfword products cannot carry)
//   I2  = G2[left] + nI2[left] * go
//   M2  = G2[diag] + (subst - ge)             (table holds subst - ge, both letters pairs)
//   h2  = max3(D2, I2, M2, 0)                 B is allowed everywhere and scores 0 (RELU)
//   nX2 = min((X2 - h2) mod 2^16, 1)          0 iff candidate X ties for the maximum
//   pointer nibble = 15 - (nB + 2 nD + 4 nI + 8 nM), written in K1's (skewed) layout so that the
//   traceback kernels are shared.
//
// Selected by the host only when it is exact: every score an integer,
// (n + m + 4) * max|score| < 4000 (no 16-bit wrap with a -12000 "does not exist"
// sentinel), go <= 0, ge <= 0 and ge + go <= 0 (padding columns can then never win
// the end-cell search), alphabet <= 4 letters, m <= 255.  The two problems of a lane group
// need the same geometry, not the same shape: the group walks max(nA, nB) rows and the
// shorter problem sees padding rows / columns (letter 4, score "does not exist"), whose cells
// can only decay (H <= the cell above / to the left) and therefore never become the first
// maximum; their pointer words are not stored.  Results are identical to K1's and to pwlib's.
//   replaces: dptable_solve's fill loop pw.c:59-110 + _std_find_optimal (:339-355)
#pragma once
#include "bsq_kernels.cuh"
#include "bsq_launch.h"

#define BSQ_K1H_NINF (-12000)

#ifdef BSQ_SIMT_EMU
static inline unsigned bsq_h2(int lo, int hi) { return ((unsigned)hi << 16) | ((unsigned)lo & 0xffffu); }
static inline int bsq_h2lo(unsigned v) { return (int)(short)(v & 0xffffu); }
static inline int bsq_h2hi(unsigned v) { return (int)(short)(v >> 16); }
static inline unsigned bsq_vadd2(unsigned a, unsigned b) { return bsq_h2(bsq_h2lo(a) + bsq_h2lo(b), bsq_h2hi(a) + bsq_h2hi(b)); }
static inline unsigned bsq_vmax3relu2(unsigned a, unsigned b, unsigned c) {
  auto f = [](int x, int y, int z) { int m = x > y ? x : y; m = m > z ? m : z; return m > 0 ? m : 0; };
  return bsq_h2(f(bsq_h2lo(a), bsq_h2lo(b), bsq_h2lo(c)), f(bsq_h2hi(a), bsq_h2hi(b), bsq_h2hi(c)));
}
static inline unsigned bsq_vaddminu2(unsigned a, unsigned b, unsigned c) {  // min((a+b) mod 2^16, c), unsigned halfwords
  auto f = [](unsigned x, unsigned y, unsigned z) { unsigned s = (x + y) & 0xffffu; return s < z ? s : z; };
  return (f(a >> 16, b >> 16, c >> 16) << 16) | f(a & 0xffffu, b & 0xffffu, c & 0xffffu);
}
static inline unsigned bsq_vminu2(unsigned a, unsigned b) {
  auto f = [](unsigned x, unsigned y) { return x < y ? x : y; };
  return (f(a >> 16, b >> 16) << 16) | f(a & 0xffffu, b & 0xffffu);
}
static inline unsigned bsq_perm(unsigned a, unsigned b, unsigned sel) {   // __byte_perm
  const unsigned long long v = ((unsigned long long)b << 32) | a;
  unsigned r = 0;
  for (int i = 0; i < 4; i++) r |= (unsigned)((v >> (8 * ((sel >> (4 * i)) & 7))) & 0xffu) << (8 * i);
  return r;
}
static inline unsigned bsq_vbmax2(unsigned a, unsigned b, bool* ph, bool* pl) {  // max, pred = (a >= b)
  *pl = bsq_h2lo(a) >= bsq_h2lo(b);
  *ph = bsq_h2hi(a) >= bsq_h2hi(b);
  return bsq_h2(*pl ? bsq_h2lo(a) : bsq_h2lo(b), *ph ? bsq_h2hi(a) : bsq_h2hi(b));
}
#else
__device__ __forceinline__ unsigned bsq_h2(int lo, int hi) { return ((unsigned)hi << 16) | ((unsigned)lo & 0xffffu); }
__device__ __forceinline__ int bsq_h2lo(unsigned v) { return (int)(short)(v & 0xffffu); }
__device__ __forceinline__ int bsq_h2hi(unsigned v) { return (int)(short)(v >> 16); }
__device__ __forceinline__ unsigned bsq_vadd2(unsigned a, unsigned b) { return __vadd2(a, b); }
__device__ __forceinline__ unsigned bsq_vmax3relu2(unsigned a, unsigned b, unsigned c) { return __vimax3_s16x2_relu(a, b, c); }
__device__ __forceinline__ unsigned bsq_vaddminu2(unsigned a, unsigned b, unsigned c) { return __viaddmin_u16x2(a, b, c); }
__device__ __forceinline__ unsigned bsq_vminu2(unsigned a, unsigned b) { return __vminu2(a, b); }
__device__ __forceinline__ unsigned bsq_vbmax2(unsigned a, unsigned b, bool* ph, bool* pl) { return __vibmax_s16x2(a, b, ph, pl); }
__device__ __forceinline__ unsigned bsq_perm(unsigned a, unsigned b, unsigned sel) { return __byte_perm(a, b, sel); }
#endif

// (G, K) is one of K1's single-pass geometries: (8, 19) for m <= 151, (16, 16) for m <= 255
// SAME: every pair of the chunk has one shape (fixed-length reads): the per-problem row / column
// bounds fold into one
template <int G, int K, int GOS, bool SAME>
__global__ void __launch_bounds__(128) bsq_k1h_fill(const BsqFillArgs A) {
  constexpr int WPL = (K + 7) / 8, WST = BSQ_K1_WPL(K);
  static_assert(GOS <= 0, "K1h handles go <= 0");
  BSQ_BAIL_IF_BAD(A);
  // (origin letter of A, of B) x (mutant letter of A, of B); letter 4 = padding row / column:
  // subst - ge of both problems, packed
  __shared__ unsigned s_sub2[25 * 25];
  const int L = A.L;
  const int ge = A.ge_i, go = A.go_i;
  for (int i = threadIdx.x; i < 25 * 25; i += blockDim.x) {
    const int r = i / 25, c = i - r * 25;
    const int oA = r / 5, oB = r - oA * 5, mA = c / 5, mB = c - mA * 5;
    const int* sub = (const int*)A.subst;
    const int vA = (mA < L && oA < L) ? sub[oA * L + mA] - ge : BSQ_K1H_NINF;
    const int vB = (mB < L && oB < L) ? sub[oB * L + mB] - ge : BSQ_K1H_NINF;
    s_sub2[i] = bsq_h2(vA, vB);
  }
  __syncthreads();

  const int tid = threadIdx.x;
  const int l = tid % G;
  const unsigned lane = (unsigned)tid & 31u;
  const unsigned gmask = ((1u << G) - 1u) << (lane & ~(unsigned)(G - 1));
  const unsigned NINF2 = bsq_h2(BSQ_K1H_NINF, BSQ_K1H_NINF);
  const unsigned ONE2 = 0x00010001u;
  const unsigned ge2 = bsq_h2(ge, ge);
  const unsigned go16 = (unsigned)go & 0xffffu;  // halfword two's complement: (0/1) * go16 per halfword
  const int nduos = A.nduos;

  for (;;) {
    int q = 0;
    if (l == 0) q = atomicAdd(A.counter, 1);
    q = __shfl_sync(gmask, q, 0, G);
    if (q >= nduos) break;
    // problems (2q, 2q + 1) of the chunk; an odd one out aligns itself twice
    const int pa = 2 * q, pb = (2 * q + 1 < A.npairs) ? 2 * q + 1 : 2 * q;
    const BsqPair PA = A.pairs[pa];
    const BsqPair PB = A.pairs[pb];
    const int nA = PA.n, mA = PA.m, nB = SAME ? nA : PB.n, mB = SAME ? mA : PB.m;
    const int n = nA > nB ? nA : nB;               // rows the group walks
    const int y0 = l * K;
    int moff[K];                                   // column's letter pair -> index into a table row
#pragma unroll
    for (int k = 0; k < K; k++) {
      const int y = y0 + k;
      const int a = (y >= 1 && y <= mA) ? (int)A.codes[PA.m_off + y - 1] : (y > mA ? 4 : 0);
      const int b = (y >= 1 && y <= mB) ? (int)A.codes[PB.m_off + y - 1] : (y > mB ? 4 : 0);
      moff[k] = a * 5 + b;
    }
    unsigned G2[K], nD2[K];
#pragma unroll
    for (int k = 0; k < K; k++) { G2[k] = NINF2; nD2[k] = 0u; }
    unsigned lastG2 = NINF2, lastN2 = 0u, prev_in = NINF2;
    unsigned best2 = 0u;                           // running maximum, starts at H(0,0) = 0 (:343)
    int kA = 0, kB = 0, xA = -1, xB = -1;          // where it was last raised (first maximum)
    uint32_t* prowA = A.ptr + PA.ptr_off + (long long)l * WST;
    uint32_t* prowB = A.ptr + PB.ptr_off + (long long)l * WST;

    // letters of the next row are fetched one row ahead, substitution scores four cells ahead:
    // with 4 warps per scheduler nothing else would hide those latencies
    const uint8_t* oseqA = A.codes + PA.o_off;
    const uint8_t* oseqB = A.codes + PB.o_off;
    int la = 0, lb = 0;                            // origin letters of the row about to be computed (raw:
                                                   // combining them right after the load would wait for it)
    // A row's pointer words are stored one iteration LATER, behind the next row's shuffles and
    // first score loads: shuffles, shared-memory loads and global stores share one in-order queue,
    // and a store that misses L2 would otherwise hold up the shuffle that starts the next row.
    uint32_t pwa[WPL], pwb[WPL];
    int px = -1;
#pragma unroll
    for (int w = 0; w < WPL; w++) { pwa[w] = 0u; pwb[w] = 0u; }
    for (int t = 0; t <= n + G - 1; t++) {
      const int x = t - l;
      unsigned inG2 = __shfl_up_sync(gmask, lastG2, 1, G);
      unsigned inN2 = __shfl_up_sync(gmask, lastN2, 1, G);
      if (l == 0) { inG2 = NINF2; inN2 = 0u; }
      const int rowi = (la * 5 + lb) * 25;
      int na, nb_;                                 // letters of row x + 1, used one iteration later (4: past the end)
      if (SAME) {
        na = 0; nb_ = 0;
        if (x >= 0 && x < n) { na = (int)oseqA[x]; nb_ = (int)oseqB[x]; }
      } else {
        na = 4; nb_ = 4;
        if (x < nA) na = (x >= 0) ? (int)oseqA[x] : 0;
        if (x < nB) nb_ = (x >= 0) ? (int)oseqB[x] : 0;
      }
      const bool active = (x >= 0 && x <= n);
      const unsigned* srow = s_sub2 + rowi;
      unsigned sq[4];
#pragma unroll
      for (int k = 0; k < 4; k++) sq[k] = srow[moff[k]];
#ifndef BSQ_SIMT_EMU
      asm volatile("" ::: "memory");
#endif
      if (px >= 0) {
        if (SAME || px <= nA) bsq_store_words<WPL, WST>(prowA + (long long)(px + l) * (G * WST), pwa);
        if (pb != pa && (SAME || px <= nB)) bsq_store_words<WPL, WST>(prowB + (long long)(px + l) * (G * WST), pwb);
        px = -1;
      }
      if (active) {
        unsigned left = inG2, nI2 = inN2, diag = prev_in;
        const unsigned before = best2;
        uint32_t wa[WPL], wb[WPL];
        constexpr int NACC = (K + 3) / 4;
        unsigned acc[NACC];                        // 4 cells' not-in-tie-set nibbles per halfword (A low, B high)
#pragma unroll
        for (int q = 0; q < NACC; q++) acc[q] = 0u;
#pragma unroll
        for (int k = 0; k < K; k++) {
          const unsigned up = G2[k];
          const unsigned sig = sq[k & 3];
          if (k + 4 < K) sq[k & 3] = srow[moff[k + 4]];
          unsigned D2 = up, I2 = left;
          if (GOS != 0) {
            D2 = bsq_vadd2(up, nD2[k] * go16);
            I2 = bsq_vadd2(left, nI2 * go16);
          }
          const unsigned M2 = bsq_vadd2(diag, sig);
          const unsigned h2 = bsq_vmax3relu2(D2, I2, M2);
          const unsigned mh2 = bsq_vadd2(~h2, ONE2);           // -h per halfword
          const unsigned nd = bsq_vaddminu2(D2, mh2, ONE2);
          const unsigned ni = bsq_vaddminu2(I2, mh2, ONE2);
          const unsigned nm = bsq_vaddminu2(M2, mh2, ONE2);
          const unsigned nb = bsq_vminu2(mh2, ONE2);            // B (score 0) ties iff h == 0
          // four multiply-adds into the accumulator (the sum first and one shift would be five
          // instructions: the doubling of nd gets its own)
          {
            const unsigned sh = 4u * (unsigned)(k & 3);   // k is an unrolled loop index: a constant
            acc[k >> 2] = nb * (1u << sh) + acc[k >> 2];
            acc[k >> 2] = nd * (2u << sh) + acc[k >> 2];
            acc[k >> 2] = ni * (4u << sh) + acc[k >> 2];
            acc[k >> 2] = nm * (8u << sh) + acc[k >> 2];
          }
          nD2[k] = nd;
          nI2 = ni;
          diag = up;
          left = bsq_vadd2(h2, ge2);
          G2[k] = left;
          bool keepB, keepA;                                     // _std_find_optimal :343-353, first strict maximum
          best2 = bsq_vbmax2(best2, h2, &keepB, &keepA);
          kA = keepA ? kA : k;
          kB = keepB ? kB : k;
        }
        lastG2 = left;
        lastN2 = nI2;
        if ((best2 ^ before) & 0xffffu) xA = x;
        if ((best2 ^ before) >> 16) xB = x;
        // a row word holds 8 cells of ONE problem: the low halves of two accumulators for A, the
        // high halves for B (one byte permute each)
#pragma unroll
        for (int w = 0; w < WPL; w++) {
          const unsigned lo = acc[2 * w], hi = (2 * w + 1 < NACC) ? acc[2 * w + 1] : 0u;
          wa[w] = bsq_perm(lo, hi, 0x5410u);
          wb[w] = bsq_perm(lo, hi, 0x7632u);
        }
#pragma unroll
        for (int w = 0; w < WPL; w++) {
          const uint32_t full = (w == WPL - 1 && (K & 7)) ? ((1u << (4 * (K & 7))) - 1u) : 0xffffffffu;
          pwa[w] = ~wa[w] & full;
          pwb[w] = ~wb[w] & full;
        }
        px = x;
      }
      prev_in = (x >= 0) ? inG2 : NINF2;
      la = na;
      lb = nb_;
    }
    if (px >= 0) {
      if (SAME || px <= nA) bsq_store_words<WPL, WST>(prowA + (long long)(px + l) * (G * WST), pwa);
      if (pb != pa && (SAME || px <= nB)) bsq_store_words<WPL, WST>(prowB + (long long)(px + l) * (G * WST), pwb);
    }

    // group-wide first maximum of each problem: higher score, then smaller (x, y)
    int sA = bsq_h2lo(best2), sB = bsq_h2hi(best2);
    int a1 = (xA >= 0) ? xA : BSQ_INT_MAX, a2 = (xA >= 0) ? y0 + kA : BSQ_INT_MAX;
    int b1 = (xB >= 0) ? xB : BSQ_INT_MAX, b2 = (xB >= 0) ? y0 + kB : BSQ_INT_MAX;
    bsq_group_best<int, G>(gmask, sA, a1, a2);
    bsq_group_best<int, G>(gmask, sB, b1, b2);
    if (l == 0) {
      {
        const int oi = PA.pad;
        const bool none = (a1 == BSQ_INT_MAX);
        A.out_score[oi] = (double)sA;
        A.out_ex[oi] = none ? -1 : a1;
        A.out_ey[oi] = none ? -1 : a2;
        A.out_status[oi] = none ? BSQ_NO_ALIGNMENT : BSQ_OK;
      }
      if (pb != pa) {
        const int oi = PB.pad;
        const bool none = (b1 == BSQ_INT_MAX);
        A.out_score[oi] = (double)sB;
        A.out_ex[oi] = none ? -1 : b1;
        A.out_ey[oi] = none ? -1 : b2;
        A.out_status[oi] = none ? BSQ_NO_ALIGNMENT : BSQ_OK;
      }
    }
  }
}
