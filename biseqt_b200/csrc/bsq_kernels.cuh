// bsq_kernels.cuh -- sm_100a DP fill + traceback kernels of libbsqgpu.
//
// What is computed is pwlib's recurrence, not textbook Gotoh: per cell the
// candidates B, D, I, M are formed in that order (pw.c:77-80), EVERY candidate
// tying for the maximum is kept (pw.c:89-108) and a gap candidate extends for
// free only out of a source cell whose tied-max set already contains the same
// gap op (_pw_internals.c:266-278).  All choices of a cell carry the same
// score, so a cell is (H, 4-bit mask); the kernels keep H in registers and
// write only the masks (4 bit/cell) to HBM for the traceback kernel.
//
// Arithmetic (template parameter S):
//   ScF64  IEEE double, the reference's own type.  Adds are issued in the
//          reference's association -- (H + ge) then (+ go) as in
//          _pw_internals.c:269-273, H + subst as in :234 -- with __dadd_rn so
//          that nothing is contracted or re-associated; ties are detected with
//          == on doubles exactly like pw.c:94.
//   ScI32  exact integer path, selected by the host only when every score is an
//          integer and no partial sum can leave 30 bits; then double arithmetic
//          in the reference is exact too and the results are identical.
//
// "Does not exist" (outside the table / band) is encoded as -inf (F64) or
// -2^30 (I32): such a neighbour yields a candidate that can never tie with a
// real one, which removes every boundary branch from the inner loops.
#pragma once
#include "bsq_common.h"
#include "bsq_launch.h"
#include "../../include/bsqgpu.h"

#define BSQ_INT_MAX 0x7fffffff
// B_LOCAL end-cell keys of the integer K2 kernel (BLOCAL == 2)
#define BSQ_K2_KEYBITS 15
#define BSQ_K2_KSTEP(s, s0) (((1 << BSQ_K2_KEYBITS) - 1) - (((s) - (s0)) >> 1))

// every fill / traceback kernel starts with this: a letter code >= alphabet_len would index the
// shared-memory score table out of bounds, so after bsq_check_codes raised the flag nothing runs
#define BSQ_BAIL_IF_BAD(A) do { if ((A).bad != nullptr && *(volatile const int*)(A).bad != 0) return; } while (0)

// a lane's NW pointer words of one row / step -> ONE aligned vector store of STRIDE words
template <int NW, int STRIDE>
__device__ __forceinline__ void bsq_store_words(uint32_t* dst, const uint32_t (&w)[NW]) {
  static_assert(STRIDE == 1 || STRIDE == 2 || STRIDE == 4, "power-of-two words per lane");
  static_assert(NW <= STRIDE, "more words than the slot holds");
  if (STRIDE == 1) {
    dst[0] = w[0];
  } else if (STRIDE == 2) {
    uint2 v;
    v.x = w[0];
    v.y = NW > 1 ? w[NW > 1 ? 1 : 0] : 0u;
    *(uint2*)dst = v;
  } else {
    uint4 v;
    v.x = w[0];
    v.y = NW > 1 ? w[NW > 1 ? 1 : 0] : 0u;
    v.z = NW > 2 ? w[NW > 2 ? 2 : 0] : 0u;
    v.w = NW > 3 ? w[NW > 3 ? 3 : 0] : 0u;
    *(uint4*)dst = v;
  }
}

struct ScF64 {
  typedef double T;
  static constexpr bool IS_INT = false;
  static __device__ __forceinline__ T ninf() {
    return __longlong_as_double((long long)0xFFF0000000000000ULL);
  }
  static __device__ __forceinline__ T add(T a, T b) { return __dadd_rn(a, b); }
  // (a) + go when the gap opens, (a) when it extends.  The "opens" flag travels as the
  // high word of the double 1.0 / 0.0 and is applied with one DFMA: fma(1,go,a) is the
  // correctly rounded a+go (the same value as the reference's second add,
  // _pw_internals.c:272) and fma(0,go,a) == a, so no select is needed.
  static __device__ __forceinline__ T open(T a, int g, T go) { return __fma_rn(__hiloint2double(g, 0), go, a); }
  static __device__ __forceinline__ int flag(bool extends) { return extends ? 0 : 0x3FF00000; }
  static __device__ __forceinline__ double to_f64(T v) { return v; }
  static __device__ __forceinline__ T go(const BsqFillArgs& A) { return A.go_d; }
  static __device__ __forceinline__ T ge(const BsqFillArgs& A) { return A.ge_d; }
};

struct ScI32 {
  typedef int T;
  static constexpr bool IS_INT = true;
  static __device__ __forceinline__ T ninf() { return -(1 << 30); }
  static __device__ __forceinline__ T add(T a, T b) { return a + b; }
  static __device__ __forceinline__ T open(T a, int g, T go) { return g * go + a; }  // one IMAD
  static __device__ __forceinline__ int flag(bool extends) { return extends ? 0 : 1; }
  static __device__ __forceinline__ double to_f64(T v) { return (double)v; }
  static __device__ __forceinline__ T go(const BsqFillArgs& A) { return A.go_i; }
  static __device__ __forceinline__ T ge(const BsqFillArgs& A) { return A.ge_i; }
};

// Decides whether a gap candidate out of a source cell with tied-max set `nib`
// extends (score H+ge) or opens (score (H+ge)+go); `own` is the gap's own bit.
//   go < 0 : extends iff own is in the set           (max over k picks k==own)
//   go > 0 : extends iff nothing but own is in the set
//   go = 0 : both give H+ge; the second add is skipped altogether
template <int GOS>
__device__ __forceinline__ bool bsq_extends(uint32_t nib, uint32_t own) {
  if (GOS < 0) return (nib & own) != 0u;
  if (GOS > 0) return (nib & (0xFu ^ own)) == 0u;
  return true;
}

// One DP cell.  up/left/diag are H of (x-1,y), (x,y-1), (x-1,y-1); fD / fI say
// whether the D / I candidate extends; bval is 0 where B is allowed, else -inf.
template <class S, int GOS, bool HASB>
__device__ __forceinline__ void bsq_cell(typename S::T up, bool fD, typename S::T left,
                                         bool fI, typename S::T diag, typename S::T sig,
                                         typename S::T bval, typename S::T go,
                                         typename S::T ge, typename S::T& H,
                                         uint32_t& mask) {
  typedef typename S::T T;
  T D = S::add(up, ge);      // _pw_internals.c:269-270
  T I = S::add(left, ge);
  if (GOS != 0) {            // :271-273
    if (!fD) D = S::add(D, go);
    if (!fI) I = S::add(I, go);
  }
  const T M = S::add(diag, sig);  // :234
  T h = D > I ? D : I;
  h = M > h ? M : h;
  if (HASB) h = bval > h ? bval : h;
  uint32_t mk = (D == h ? BSQ_MD : 0u) | (I == h ? BSQ_MI : 0u) | (M == h ? BSQ_MM : 0u);
  if (HASB) mk |= (bval == h ? BSQ_MB : 0u);
  H = h;
  mask = mk;
}

// The same cell with the open/extend decision carried as S::flag() words (see
// ScF64::open) and the tie set returned as four predicates, which keeps the
// ALU-pipe work per cell minimal (the pipe that bounds these kernels).
template <class S, int GOS, bool HASB>
__device__ __forceinline__ void bsq_cell2(typename S::T up, int gD, typename S::T left, int gI,
                                          typename S::T diag, typename S::T sig, typename S::T bval,
                                          typename S::T go, typename S::T ge, typename S::T& H,
                                          bool& eB, bool& eD, bool& eI, bool& eM) {
  typedef typename S::T T;
  T D = S::add(up, ge);      // _pw_internals.c:269-270
  T I = S::add(left, ge);
  if (GOS != 0) {            // :271-273
    D = S::open(D, gD, go);
    I = S::open(I, gI, go);
  }
  const T M = S::add(diag, sig);  // :234
  T h = D > I ? D : I;
  h = M > h ? M : h;
  if (HASB) h = bval > h ? bval : h;
  eD = (D == h);
  eI = (I == h);
  eM = (M == h);
  eB = HASB ? (bval == h) : false;
  H = h;
}

// 0 if h == x, else 1 (h >= x).  Written in PTX on the device: left to the compiler,
// nvcc 12.9 folds "h - x == 0" into the predicate output of the VIMNMX that produced
// h, which is wrong when the two max operands tie (found by the GPU parity tests).
__device__ __forceinline__ int bsq_nz(int h, int x) {
#ifdef BSQ_SIMT_EMU
  return (unsigned)h - (unsigned)x != 0u;
#else
  unsigned r;
  asm("{\n\t.reg .u32 t;\n\tsub.u32 t, %1, %2;\n\tmin.u32 %0, t, 1;\n\t}" : "=r"(r) : "r"(h), "r"(x));
  return (int)r;
#endif
}

// 0 if x == h, else 1, given mh = -h (h >= x always): (x - h) mod 2^32 is 0 or >= 2^31, so one
// unsigned add-min (VIADDMNMX.U32) per candidate and ONE negation per cell.  (bsq_nz above needs
// a negation per candidate: measured 2-3 IMAD per cell in the issue-bound K2 loop.)
__device__ __forceinline__ int bsq_nzu(unsigned mh, int x) { return (int)__viaddmin_u32((unsigned)x, mh, 1u); }

// Integer cell without predicates: nX = min(h - X, 1) is 0 iff candidate X ties
// for the maximum (h >= X always), so the tie set, the open/extend flags and the
// pointer nibble are plain integer arithmetic -- subtractions and multiply-adds
// that ptxas places on the FMA pipe (IMAD) -- and the ALU pipe, which bounds
// these kernels, is left with the DPX max (VIMNMX3) and the four min() ops.
template <int GOS, bool HASB, bool BZERO>
__device__ __forceinline__ void bsq_cell_i32(int up, int gD, int left, int gI, int diag, int sig, int bval,
                                             int go, int ge, int& H, int& nB, int& nD, int& nI, int& nM) {
  int D = up + ge;
  int I = left + ge;
  if (GOS != 0) {
    D = gD * go + D;
    I = gI * go + I;
  }
  const int M = diag + sig;
  int h;
  if (HASB && BZERO) h = __vimax3_s32_relu(D, I, M);
  else {
    h = __vimax3_s32(D, I, M);
    if (HASB) h = max(h, bval);
  }
  // differences in unsigned arithmetic: a -2^30 sentinel candidate may sit 2^31 below h
  const unsigned mh = 0u - (unsigned)h;
  nD = bsq_nzu(mh, D);
  nI = bsq_nzu(mh, I);
  nM = bsq_nzu(mh, M);
  nB = HASB ? bsq_nzu(mh, BZERO ? 0 : bval) : 1;
  H = h;
}

// "the gap must open" flag out of a cell whose NOT-in-tie-set indicators are given
template <int GOS>
__device__ __forceinline__ int bsq_gapflag_i32(int nOwn, int nO1, int nO2, int nO3) {
  if (GOS < 0) return nOwn;                                              // extends iff own op in the set
  if (GOS > 0) return (int)min((unsigned)(nOwn + 3 - nO1 - nO2 - nO3), 1u);  // extends iff nothing else is
  return 0;
}

// open/extend flag of a gap out of a cell with tie set (eB,eD,eI,eM); own = the gap's op
template <class S, int GOS>
__device__ __forceinline__ int bsq_gapflag(bool eOwn, bool eO1, bool eO2, bool eO3) {
  if (GOS < 0) return S::flag(eOwn);                           // extends iff own op in the set
  if (GOS > 0) return S::flag(eOwn && !(eO1 || eO2 || eO3));   // extends iff nothing else is
  return 0;
}

// lexicographic "better end cell": higher score, then earlier scan position
template <class T>
__device__ __forceinline__ bool bsq_better(T s, int k1, int k2, T bs, int b1, int b2) {
  return s > bs || (s == bs && (k1 < b1 || (k1 == b1 && k2 < b2)));
}

// group-wide argmax with the reference's scan-order tie-break
template <class T, int G>
__device__ __forceinline__ void bsq_group_best(unsigned gmask, T& s, int& k1, int& k2) {
#pragma unroll
  for (int off = 1; off < G; off <<= 1) {
    T os = __shfl_xor_sync(gmask, s, off, G);
    int o1 = __shfl_xor_sync(gmask, k1, off, G);
    int o2 = __shfl_xor_sync(gmask, k2, off, G);
    if (bsq_better<T>(os, o1, o2, s, k1, k2)) { s = os; k1 = o1; k2 = o2; }
  }
}

// ===========================================================================
// K1: STD frame.  A group of G lanes owns one pair; lane l owns K consecutive
// columns (y), walks the rows (x) and lags its left neighbour by one row, so
// the (x, y-1) dependency arrives by one shuffle per row.  Tables wider than
// G*K columns are swept in passes, the boundary column going through a small
// per-group scratch slab.  Pointer layout: [pass][t = x + lane][lane][word], 8
// cells/word: skewed by the lane index like the computation itself, so that at
// every step the G lanes of a group write ONE contiguous run of full sectors
// (unskewed, each lane would write 16 bytes of a different row: 8x the L2 write
// requests, half-written sectors, read-modify-write in DRAM).
//   replaces: dptable_solve's fill loop pw.c:59-110 + _std_find_optimal
// ===========================================================================
// EALL: 0 = end cell not searched per cell, 1 = every cell may end the alignment
// (LOCAL / START_ANCHORED), 2 = same but padding columns must be masked out
// explicitly (only needed for score sets where a gap does not lower the score).
template <class S, int K, int GOS, bool ALLB, int EALL, bool BZ>
__device__ __forceinline__ void bsq_k1_row(
    typename S::T (&Hup)[K], int (&gD)[K], uint32_t (&cw)[(K + 7) / 8], const int (&moff)[K],
    const char* srow, typename S::T inH, int inG, typename S::T diag_in,
    typename S::T bfirst, typename S::T brest, typename S::T go, typename S::T ge, int xk,
    int kv, typename S::T& best, int& bpos, typename S::T& outH, int& outG) {
  typedef typename S::T T;
  constexpr int WPL = (K + 7) / 8;
  T left = inH;
  int gI = inG;
  T diag = diag_in;
#pragma unroll
  for (int w = 0; w < WPL; w++) cw[w] = 0u;
  if constexpr (S::IS_INT) {
    // end-cell search by packed key: (H << 5 | 31 - k) is maximal at the row's first
    // maximum, so one IMAD + one max per cell and one compare per row do the job
    int rowkey = -2147483647 - 1;
#pragma unroll
    for (int k = 0; k < K; k++) {
      const int up = Hup[k];
      const int sig = *(const int*)(srow + moff[k]);
      int H, nB, nD, nI, nM;
      if (ALLB || k == 0) {
        if (ALLB && BZ) bsq_cell_i32<GOS, true, true>(up, gD[k], left, gI, diag, sig, 0, go, ge, H, nB, nD, nI, nM);
        else bsq_cell_i32<GOS, true, false>(up, gD[k], left, gI, diag, sig, k == 0 ? bfirst : brest, go, ge, H, nB, nD, nI, nM);
      } else {
        bsq_cell_i32<GOS, false, false>(up, gD[k], left, gI, diag, sig, 0, go, ge, H, nB, nD, nI, nM);
      }
      const int notmask = nB + 2 * nD + 4 * nI + 8 * nM;
      cw[k >> 3] += (uint32_t)(15 - notmask) << (4 * (k & 7));
      gD[k] = bsq_gapflag_i32<GOS>(nD, nB, nI, nM);
      gI = bsq_gapflag_i32<GOS>(nI, nB, nD, nM);
      diag = up;
      left = H;
      Hup[k] = H;
      if (EALL == 1) rowkey = max(rowkey, H * 32 + (31 - k));
      else if (EALL == 2) { if (k < kv) rowkey = max(rowkey, H * 32 + (31 - k)); }
    }
    if (EALL != 0) {
      if ((rowkey >> 5) > best) { best = rowkey >> 5; bpos = xk + (31 - (rowkey & 31)); }
    }
  } else {
#pragma unroll
  for (int k = 0; k < K; k++) {
    const T up = Hup[k];
    const T sig = *(const T*)(srow + moff[k]);
    T H;
    bool eB, eD, eI, eM;
    if (ALLB || k == 0)
      bsq_cell2<S, GOS, true>(up, gD[k], left, gI, diag, sig, k == 0 ? bfirst : brest, go, ge, H, eB, eD, eI, eM);
    else
      bsq_cell2<S, GOS, false>(up, gD[k], left, gI, diag, sig, bfirst, go, ge, H, eB, eD, eI, eM);
    uint32_t w = cw[k >> 3];
    if (eB) w += BSQ_MB << (4 * (k & 7));
    if (eD) w += BSQ_MD << (4 * (k & 7));
    if (eI) w += BSQ_MI << (4 * (k & 7));
    if (eM) w += BSQ_MM << (4 * (k & 7));
    cw[k >> 3] = w;
    gD[k] = bsq_gapflag<S, GOS>(eD, eB, eI, eM);
    gI = bsq_gapflag<S, GOS>(eI, eB, eD, eM);
    diag = up;
    left = H;
    Hup[k] = H;
    if (EALL == 1) {  // _std_find_optimal :343-353, row-major first strict max
      if (H > best) { best = H; bpos = xk + k; }
    } else if (EALL == 2) {
      if (H > best && k < kv) { best = H; bpos = xk + k; }
    }
  }
  }
  outH = left;
  outG = gI;
}

template <class S, int G, int K, bool BALL, int EALL, int GOS>
// (minBlocks = 3 for the F64 instantiations is a scheduling hint: ptxas keeps ~160 registers either way,
// but orders the row loop better with it -- 586 -> 625 GCUPS on 150 bp log-odds pairs; 2 gives 619, none 586.
// The integer ones and K1h are fastest without a hint: profiles/r02_launch_bounds_sweep.txt.)
__global__ void __launch_bounds__(128, S::IS_INT ? 0 : 3) bsq_k1_fill(const BsqFillArgs A) {
  typedef typename S::T T;
  static_assert(K <= 32, "end-cell positions are packed as x*32 + k");
  BSQ_BAIL_IF_BAD(A);
  constexpr int WPL = (K + 7) / 8;          // words holding a lane's K nibbles
  constexpr int WST = BSQ_K1_WPL(K);        // words a lane's row occupies in the table
  constexpr int BW = G * K;
  // substitution table with one extra column (letter L = "padding column") holding
  // -inf, so cells right of the table can never win or tie anything
  __shared__ T s_sub[BSQ_MAX_L * (BSQ_MAX_L + 1)];
  const int L = A.L, Ls = L + 1;
  for (int i = threadIdx.x; i < L * Ls; i += blockDim.x) {
    const int o = i / Ls, mm = i - o * Ls;
    s_sub[i] = (mm < L) ? ((const T*)A.subst)[o * L + mm] : S::ninf();
  }
  __syncthreads();

  const int tid = threadIdx.x;
  const int l = tid % G;
  const unsigned lane = (unsigned)tid & 31u;
  const unsigned gmask = (G == 32) ? 0xffffffffu : (((1u << G) - 1u) << (lane & ~(unsigned)(G - 1)));
  const long long group = ((long long)blockIdx.x * blockDim.x + tid) / G;
  const T NINF = S::ninf();
  const T ZERO = (T)0;
  const T go = S::go(A), ge = S::ge(A);
  const int bmode = A.bmode, emode = A.emode;  // EDGE/ORIGIN and EDGE/CORNER are runtime
  T* colh = (T*)(A.col_h + group * A.col_stride);
  uint8_t* colf = A.col_f + group * A.col_stride;

  for (;;) {
    int p = 0;
    if (l == 0) p = atomicAdd(A.counter, 1);
    p = __shfl_sync(gmask, p, 0, G);
    if (p >= A.npairs) break;
    const BsqPair P = A.pairs[p];
    const int n = P.n, m = P.m;
    const int npass = (m + BW) / BW;  // ceil((m+1)/BW)

    // running end-cell record: E_ALL -> (x, y); E_EDGE -> (scan key, 0); E_CORNER -> none
    T best = EALL ? ZERO : NINF;
    int bk1 = BSQ_INT_MAX, bk2 = BSQ_INT_MAX;

    for (int pass = 0; pass < npass; pass++) {
      const int y0 = pass * BW + l * K;
      int moff[K];  // byte offset of my column's letter inside a table row
#pragma unroll
      for (int k = 0; k < K; k++) {
        const int y = y0 + k;
        const int let = (y >= 1 && y <= m) ? (int)A.codes[P.m_off + y - 1] : (y > m ? L : 0);
        moff[k] = let * (int)sizeof(T);
      }
      int kv = m - y0 + 1;
      kv = kv < 0 ? 0 : (kv > K ? K : kv);
      const int km = m - y0;  // which of my columns is column m (if in [0,K))
      T Hup[K];
      int gD[K];
      uint32_t cw[WPL];
#pragma unroll
      for (int k = 0; k < K; k++) { Hup[k] = NINF; gD[k] = 0; }
      T lastH = NINF;
      int lastG = 0;
      T prev_in = NINF;
      T pbest = ZERO;  // per-pass record for E_ALL (merged lexicographically below)
      int pbpos = BSQ_INT_MAX;
      uint32_t* prow = A.ptr + P.ptr_off + ((long long)pass * (n + G) * G + l) * WST;
      const bool more = (pass + 1 < npass);

      for (int t = 0; t <= n + G - 1; t++) {
        const int x = t - l;
        T inH = __shfl_up_sync(gmask, lastH, 1, G);
        int inG = __shfl_up_sync(gmask, lastG, 1, G);
        const bool active = (x >= 0 && x <= n);
        if (l == 0) {
          inH = NINF;
          inG = 0;
          if (pass > 0 && active) { inH = colh[x]; inG = S::flag(colf[x] == 0); }
        }
        if (active) {
          const int o = (x >= 1) ? (int)A.codes[P.o_off + x - 1] : 0;
          const char* srow = (const char*)s_sub + o * Ls * (int)sizeof(T);
          T outH;
          int outG;
          if (BALL || x == 0) {
            T bfirst, brest;
            if (BALL) { bfirst = ZERO; brest = ZERO; }
            else {
              brest = (bmode >= BSQ_B_EDGE) ? ZERO : NINF;
              bfirst = (bmode >= BSQ_B_EDGE || y0 == 0) ? ZERO : NINF;
            }
            bsq_k1_row<S, K, GOS, true, EALL, BALL>(Hup, gD, cw, moff, srow, inH, inG, prev_in, bfirst, brest, go, ge,
                                              x * 32, kv, pbest, pbpos, outH, outG);
          } else {
            const T bfirst = (bmode == BSQ_B_EDGE && y0 == 0) ? ZERO : NINF;
            bsq_k1_row<S, K, GOS, false, EALL, false>(Hup, gD, cw, moff, srow, inH, inG, prev_in, bfirst, NINF, go, ge,
                                               x * 32, kv, pbest, pbpos, outH, outG);
          }
          lastH = outH;
          lastG = outG;
          bsq_store_words<WPL, WST>(prow + (long long)t * (G * WST), cw);  // skewed: row t = x + l
          if (more && l == G - 1) { colh[x] = outH; colf[x] = (uint8_t)(outG != 0); }
          if (!EALL) {
            if (km >= 0 && km < K) {  // I own column m
              T Hm = NINF;
#pragma unroll
              for (int k = 0; k < K; k++) if (k == km) Hm = Hup[k];
              if (emode == BSQ_E_CORNER) {
                if (x == n) { best = Hm; bk1 = 0; bk2 = 0; }
              } else if (x < n) {  // right column, rows above the last: key = x (:321-326)
                if (bsq_better<T>(Hm, x, 0, best, bk1, bk2)) { best = Hm; bk1 = x; bk2 = 0; }
              }
            }
            if (emode == BSQ_E_EDGE && x == n) {  // bottom row, left to right: key = n + y
#pragma unroll
              for (int k = 0; k < K; k++) {
                if (k < kv && bsq_better<T>(Hup[k], n + y0 + k, 0, best, bk1, bk2)) {
                  best = Hup[k]; bk1 = n + y0 + k; bk2 = 0;
                }
              }
            }
          }
        }
        prev_in = (x >= 0) ? inH : NINF;
      }
      if (EALL) {
        const int pbx = (pbpos == BSQ_INT_MAX) ? BSQ_INT_MAX : (pbpos >> 5);
        const int pby = (pbpos == BSQ_INT_MAX) ? BSQ_INT_MAX : (y0 + (pbpos & 31));
        if (bsq_better<T>(pbest, pbx, pby, best, bk1, bk2)) { best = pbest; bk1 = pbx; bk2 = pby; }
      }
      if (more) __syncwarp(gmask);  // boundary column visible to lane 0 of the next pass
    }

    bsq_group_best<T, G>(gmask, best, bk1, bk2);
    if (l == 0) {
      int ex, ey, st = BSQ_OK;
      if (EALL) {
        ex = bk1; ey = bk2;
        if (bk1 == BSQ_INT_MAX) { st = BSQ_NO_ALIGNMENT; ex = -1; ey = -1; }
      } else if (emode == BSQ_E_EDGE) {
        if (bk1 < n) { ex = bk1; ey = m; } else { ex = n; ey = bk1 - n; }
      } else { ex = n; ey = m; }
      const int oi = P.pad;  // index of this pair in the caller's arrays
      A.out_score[oi] = S::to_f64(best);
      A.out_ex[oi] = ex;
      A.out_ey[oi] = ey;
      A.out_status[oi] = st;
    }
  }
}

// ===========================================================================
// K2: BANDED frame.  A group of G lanes owns one pair; lane l owns C (even)
// consecutive diagonals d = x - y and the group sweeps anti-diagonals s = x+y
// in lockstep: a cell on diagonal d at step s needs d-1 and d+1 from step s-1
// and d from step s-2, so every lane updates C/2 cells per step and only the
// strip edges cross lanes (one shuffle up on even steps, one down on odd).
// Pointer layout: [step s - s0][lane][word], C/2 nibbles per lane per step.
//
// The sweep runs in three phases.  While some band cell can still lie on a
// table edge, outside the table, or past the end of its diagonal ("careful"
// steps) every cell is gated individually.  In between ("interior" steps: every
// band diagonal has a cell with x >= 1 and y >= 1) nothing is gated, B is
// either impossible (B_GLOBAL, B_OVERLAP) or always 0 (B_LOCAL), and for DNA-
// sized alphabets the letters come from two nibble-packed shift registers (the
// origin window slides one way, the mutant window the other, like a systolic
// array) so that a cell's substitution score is one shared-memory load indexed
// by a nibble of their OR.
//   replaces: dptable_solve's fill loop pw.c:59-110 in (d,a) coordinates
//   (_pw_internals.c:87-153) + _banded_find_optimal (:364-414)
// ===========================================================================
// State per owned diagonal c: H[c] (latest cell), fD[c] / fI[c] = open/extend flag
// words (S::flag) of the D candidate of the cell below it / the I candidate of the
// cell right of it -- i.e. what diagonals c+1 and c-1 will need one step later.

// careful step: any cell may be off-table, on an edge, or a padding diagonal
template <class S, int C, int BLOCAL, int GOS, int Q>
__device__ __forceinline__ void bsq_k2_step_careful(
    int btype, typename S::T (&H)[C], int (&fD)[C], int (&fI)[C], typename S::T (&best)[BLOCAL ? C : 1],
    int (&bests)[BLOCAL == 1 ? C : 1], typename S::T edgeH, int edgeF, const typename S::T* s_sub, int L,
    const uint8_t* oseq, const uint8_t* mseq, int n, int m, int s, int d0, int nreal,
    typename S::T go, typename S::T ge, uint32_t (&cur)[(C / 2 + 7) / 8], const int kstep) {
  typedef typename S::T T;
  constexpr int HC = C / 2;
  constexpr int WPS = (HC + 7) / 8;
  const T NINF = S::ninf();
#pragma unroll
  for (int w = 0; w < WPS; w++) cur[w] = 0u;
  // Q == 0: cells c = 0,2,..; c = 0 takes (x-1,y) from the lane below (edge*)
  // Q == 1: cells c = 1,3,..; c = C-1 takes (x,y-1) from the lane above (edge*)
#pragma unroll
  for (int hc = 0; hc < HC; hc++) {
    const int c = 2 * hc + Q;
    const int d = d0 + c;
    const int x = (s + d) >> 1, y = (s - d) >> 1;
    const bool intab = (c < nreal) && x >= 0 && y >= 0 && x <= n && y <= m;
    T up, left;
    int gD, gI;
    if (c == 0) { up = edgeH; gD = edgeF; }
    else { up = H[c - 1]; gD = fD[c - 1]; }
    if (c == C - 1) { left = edgeH; gI = edgeF; }
    else { left = H[c + 1]; gI = fI[c + 1]; }
    const T diag = H[c];
    const int o = (x >= 1 && x <= n) ? (int)oseq[x - 1] : 0;
    const int mm = (y >= 1 && y <= m) ? (int)mseq[y - 1] : 0;
    const T sig = s_sub[o * L + mm];
    bool ballow;  // _alnchoice_B, _pw_internals.c:182-197
    if (BLOCAL) ballow = true;
    else if (btype == BSQ_B_OVERLAP) ballow = (x == 0 || y == 0);
    else ballow = (x == 0 && y == 0);
    const T bval = (intab && ballow) ? (T)0 : NINF;
    T h;
    bool eB, eD, eI, eM;
    bsq_cell2<S, GOS, true>(up, gD, left, gI, diag, sig, bval, go, ge, h, eB, eD, eI, eM);
    const uint32_t mk = (eB ? BSQ_MB : 0u) | (eD ? BSQ_MD : 0u) | (eI ? BSQ_MI : 0u) | (eM ? BSQ_MM : 0u);
    cur[hc >> 3] |= mk << (4 * (hc & 7));
    if (intab) {  // cells of one step never read each other, so commit in place
      H[c] = h;
      fD[c] = bsq_gapflag<S, GOS>(eD, eB, eI, eM);
      fI[c] = bsq_gapflag<S, GOS>(eI, eB, eD, eM);
      if (BLOCAL == 1) {  // :397-406 restricted to one diagonal (a ascends with s)
        if (h > best[c]) { best[c] = h; bests[c] = s; }
      }
    }
    if constexpr (BLOCAL == 2) {   // the same record as ONE packed key (see bsq_k2_fill), kept only for table cells
      const int key = (int)h * (1 << BSQ_K2_KEYBITS) + kstep;
      best[c] = (intab && key > (int)best[c]) ? key : best[c];
    }
  }
}

// interior step: every band cell is a regular cell (x >= 1, y >= 1, in table).
// comb: nibble hc = 4*origin letter + mutant letter of cell hc (PACKED), else
// the letters are read per cell.  fixc: local index of the first diagonal past
// the band in this lane (or -1): it is reset to -inf so the band edge stays shut.
template <class S, int C, int BLOCAL, int GOS, int Q, bool PACKED>
__device__ __forceinline__ void bsq_k2_step_fast(
    typename S::T (&H)[C], int (&fD)[C], int (&fI)[C], typename S::T (&best)[BLOCAL ? C : 1],
    int (&bests)[BLOCAL == 1 ? C : 1], typename S::T edgeH, int edgeF,
    const typename S::T* s_sub, int L, unsigned long long comb, const uint8_t* oseq,
    const uint8_t* mseq, int s, int d0, int fixc, typename S::T go, typename S::T ge,
    uint32_t* pout, const int kstep, const int n, const int m) {
  // pout: the lane's staging slot of this step in shared memory (words BSQ_FILL_THREADS apart)
  typedef typename S::T T;
  constexpr int HC = C / 2;
  constexpr int WPS = (HC + 7) / 8;
  constexpr int SH = (sizeof(T) == 8) ? 3 : 2;  // log2 of the table entry size
  uint32_t cur[WPS];
#pragma unroll
  for (int w = 0; w < WPS; w++) {
    cur[w] = 0u;
    if (S::IS_INT) {   // every candidate of every cell ties (B only where it exists); the cells subtract
      const int ncell = (HC - 8 * w) < 8 ? (HC - 8 * w) : 8;
      const uint32_t each = BLOCAL ? 0xFFFFFFFFu : 0xEEEEEEEEu;
      cur[w] = ncell >= 8 ? each : (each & ((1u << (4 * (ncell > 0 ? ncell : 0))) - 1u));
    }
  }
  const uint32_t comb_lo = (uint32_t)comb, comb_hi = (uint32_t)(comb >> 32);
#pragma unroll
  for (int hc = 0; hc < HC; hc++) {
    const int c = 2 * hc + Q;
    T up, left;
    int gD, gI;
    if (c == 0) { up = edgeH; gD = edgeF; }
    else { up = H[c - 1]; gD = fD[c - 1]; }
    if (c == C - 1) { left = edgeH; gI = edgeF; }
    else { left = H[c + 1]; gI = fI[c + 1]; }
    const T diag = H[c];
    T sig;
    if (PACKED) {
      const uint32_t cw = (hc < 8) ? comb_lo : comb_hi;
      const int sh = 4 * (hc & 7) - SH;  // nibble -> byte offset into the 16-entry table
      const uint32_t off = (sh >= 0 ? (cw >> (sh >= 0 ? sh : 0)) : (cw << (sh < 0 ? -sh : 0))) & (0xFu << SH);
      sig = *(const T*)((const char*)s_sub + off);
    } else {
      // (a lane's diagonals past the band are computed and thrown away: their cells may lie outside
      // the table, so the letter indices are clamped -- found by the AddressSanitizer run of the
      // emulator build, tools/asan_cases.py)
      const int d = d0 + c;
      int xi = ((s + d) >> 1) - 1, yi = ((s - d) >> 1) - 1;
      xi = xi < 0 ? 0 : (xi > n - 1 ? n - 1 : xi);
      yi = yi < 0 ? 0 : (yi > m - 1 ? m - 1 : yi);
      sig = s_sub[(int)oseq[xi] * L + (int)mseq[yi]];
    }
    if constexpr (S::IS_INT) {
      // Integer interior cell, the issue-bound inner loop of the banded fill.  In this phase the state
      // array holds G = H + ge (the caller converts on entry and exit, like K1h) and the score table
      // holds subst - ge, so the three candidates cost one add (two more multiply-adds for affine
      // gaps); the tie flags nX = min((X - h) mod 2^32, 1) share one negation; the pointer word
      // starts out as "every candidate ties" and each flag subtracts its bit with one multiply-add.
      int D = up, I = left;
      if (GOS != 0) { D = gD * go + D; I = gI * go + I; }
      const int M = diag + sig;
      const int h = BLOCAL ? __vimax3_s32_relu(D, I, M) : __vimax3_s32(D, I, M);
      const unsigned mh = 0u - (unsigned)h;
      const int nD = bsq_nzu(mh, D), nI = bsq_nzu(mh, I), nM = bsq_nzu(mh, M);
      const int nB = BLOCAL ? (int)min((unsigned)h, 1u) : 1;   // B (score 0) ties iff h == 0
      const uint32_t SHB = 4u * (uint32_t)(hc & 7);            // hc: an unrolled loop index, a constant
      uint32_t w = cur[hc >> 3];
      if (BLOCAL) w = (uint32_t)nB * (0u - (1u << SHB)) + w;
      w = (uint32_t)nD * (0u - (2u << SHB)) + w;
      w = (uint32_t)nI * (0u - (4u << SHB)) + w;
      w = (uint32_t)nM * (0u - (8u << SHB)) + w;
      cur[hc >> 3] = w;
      H[c] = h + ge;
      fD[c] = bsq_gapflag_i32<GOS>(nD, nB, nI, nM);
      fI[c] = bsq_gapflag_i32<GOS>(nI, nB, nD, nM);
      if (BLOCAL == 1) {
        if (h > best[c]) { best[c] = h; bests[c] = s; }
      } else if (BLOCAL == 2) {   // (score, earliest step) as one key: a multiply-add and a max
        best[c] = max(best[c], h * (1 << BSQ_K2_KEYBITS) + kstep);
      }
    } else {
    T h;
    bool eB, eD, eI, eM;
    bsq_cell2<S, GOS, BLOCAL != 0>(up, gD, left, gI, diag, sig, (T)0, go, ge, h, eB, eD, eI, eM);
    uint32_t w = cur[hc >> 3];
    if (BLOCAL) { if (eB) w += BSQ_MB << (4 * (hc & 7)); }
    if (eD) w += BSQ_MD << (4 * (hc & 7));
    if (eI) w += BSQ_MI << (4 * (hc & 7));
    if (eM) w += BSQ_MM << (4 * (hc & 7));
    cur[hc >> 3] = w;
    H[c] = h;
    fD[c] = bsq_gapflag<S, GOS>(eD, eB, eI, eM);
    fI[c] = bsq_gapflag<S, GOS>(eI, eB, eD, eM);
    if (BLOCAL == 1) {
      if (h > best[c]) { best[c] = h; bests[c] = s; }
    }
    }
  }
  if (fixc >= 0 && (fixc & 1) == Q) {  // one lane per group, every other step
    const T NINF = S::ninf();
    switch (fixc >> 1) {
#define BSQ_FIX(J) case J: if (J < HC) H[(2 * J + Q) < C ? (2 * J + Q) : 0] = NINF; break;
      BSQ_FIX(0) BSQ_FIX(1) BSQ_FIX(2) BSQ_FIX(3) BSQ_FIX(4) BSQ_FIX(5) BSQ_FIX(6) BSQ_FIX(7)
      BSQ_FIX(8) BSQ_FIX(9) BSQ_FIX(10) BSQ_FIX(11) BSQ_FIX(12) BSQ_FIX(13) BSQ_FIX(14) BSQ_FIX(15)
#undef BSQ_FIX
    }
  }
#pragma unroll
  for (int w = 0; w < WPS; w++) pout[w * BSQ_FILL_THREADS] = cur[w];
}

// minBlocks hint of bsq_k2_fill (0 = none): the 32 x 16 integer geometry (bands of 257 - 512 diagonals,
// config 4) lands at 128 registers = 4 blocks per SM on its own; capped at 96 it runs 5 and fills 9 %
// faster (1171 -> 1278 GCUPS on 4096 config-4 pairs).  Every hint tried on the 8 x 18 integer and the
// 16 x 10 F64 geometries (config 3) cost 0 - 20 %, so it is set per instantiation
// (profiles/r02_launch_bounds_sweep.txt).
template <class S, int G, int C, int BLOCAL, int GOS>
constexpr int bsq_k2_minblocks() {
  if (S::IS_INT && G == 32 && C == 16 && BLOCAL == 0) return 5;
  if (S::IS_INT && G == 8 && C == 18 && BLOCAL == 2 && GOS == 0) return 5;   // banded LOCAL, packed keys: 1006 -> 1025 GCUPS
  return 0;
}

template <class S, int G, int C, int BLOCAL, int GOS>
__global__ void __launch_bounds__(128, bsq_k2_minblocks<S, G, C, BLOCAL, GOS>()) bsq_k2_fill(const BsqFillArgs A) {
  typedef typename S::T T;
  static_assert(C % 2 == 0 && C <= 32, "lanes own an even number of diagonals, at most 32");
  constexpr int HC = C / 2;
  constexpr int WPS = (HC + 7) / 8;
  constexpr int NB = BLOCAL ? C : 1;
  constexpr int NBS = BLOCAL == 1 ? C : 1;
  static_assert(BLOCAL != 2 || S::IS_INT, "packed end-cell keys are an integer form");
  constexpr int TB = BSQ_K2_T(C);          // steps per 32-byte unit of the pointer table
  constexpr int UW = BSQ_K2_UNIT_WORDS;
  static_assert(TB >= 2 && TB % 2 == 0 && TB * WPS == UW, "a block of steps fills exactly one unit");
  BSQ_BAIL_IF_BAD(A);
  __shared__ uint32_t s_unit[UW][BSQ_FILL_THREADS];   // careful steps stage their block here
  __shared__ T s_sub[BSQ_MAX_L * BSQ_MAX_L];
  const int L = A.L;
  // DNA-sized alphabets: the table is re-laid out 4 x 4 so that a nibble
  // (4*o + m) indexes it directly
  const bool packed = (L <= 4);
  for (int i = threadIdx.x; i < BSQ_MAX_L * BSQ_MAX_L; i += blockDim.x) s_sub[i] = (T)0;
  __syncthreads();
  // the integer interior steps keep G = H + ge as their state and read subst - ge (see bsq_k2_step_fast)
  __shared__ T s_subm[S::IS_INT ? BSQ_MAX_L * BSQ_MAX_L : 1];
  for (int i = threadIdx.x; i < L * L; i += blockDim.x) {
    const int o = i / L, mm = i - o * L;
    s_sub[packed ? (o * 4 + mm) : i] = ((const T*)A.subst)[i];
  }
  __syncthreads();
  if (S::IS_INT) {
    for (int i = threadIdx.x; i < BSQ_MAX_L * BSQ_MAX_L; i += blockDim.x) s_subm[i] = S::add(s_sub[i], (T)0 - S::ge(A));
    __syncthreads();
  }
  const T* s_fast = S::IS_INT ? s_subm : s_sub;
  const int Lt = packed ? 4 : L;  // row stride of s_sub

  const int tid = threadIdx.x;
  const int l = tid % G;
  const unsigned lane = (unsigned)tid & 31u;
  const unsigned gmask = (G == 32) ? 0xffffffffu : (((1u << G) - 1u) << (lane & ~(unsigned)(G - 1)));
  const T NINF = S::ninf();
  const T go = S::go(A), ge = S::ge(A);
  const int btype = A.btype;

  // Each group runs its own loop over pairs (group-mask shuffles only).  A
  // warp-synchronous variant that votes on careful/interior per warp was measured
  // slower on B200 (more gated steps, more live state), so groups are independent.
  for (;;) {
    int p = 0;
    if (l == 0) p = atomicAdd(A.counter, 1);
    p = __shfl_sync(gmask, p, 0, G);
    if (p >= A.npairs) break;
    const BsqPair P = A.pairs[p];
    const int n = P.n, m = P.m, dmin = P.dmin, dmax = P.dmax;
    const int W = dmax - dmin + 1;
    const uint8_t* oseq = A.codes + P.o_off;
    const uint8_t* mseq = A.codes + P.m_off;
    const int d0 = dmin + l * C;       // my first diagonal
    int nreal = W - l * C;             // how many of my diagonals are inside the band
    nreal = nreal < 0 ? 0 : (nreal > C ? C : nreal);
    const int fixc = (nreal < C && W - l * C >= 0) ? nreal : -1;  // first diagonal past the band, if mine
    // last anti-diagonal holding a cell: max over d of 2*min(n, m+d) - d, at d* = clamp(n-m)
    int dstar = n - m;
    dstar = dstar < dmin ? dmin : (dstar > dmax ? dmax : dstar);
    const int s_end = 2 * (n < m + dstar ? n : m + dstar) - dstar;
    const int s0 = P.s0;               // (s0 - dmin) is even: step parity == loop parity
    const int s_last = s0 + 2 * ((s_end - s0) / 2) + 1;   // the last step the pair loops below execute
    // interior steps: all band diagonals have x >= 1, y >= 1 and have not ended
    const int admin = dmin < 0 ? -dmin : dmin, admax = dmax < 0 ? -dmax : dmax;
    const int si_lo = (admin > admax ? admin : admax) + 2;
    const int e_lo = 2 * (n < m + dmin ? n : m + dmin) - dmin, e_hi = 2 * (n < m + dmax ? n : m + dmax) - dmax;
    const int si_hi = e_lo < e_hi ? e_lo : e_hi;
    int sA = s0;
    if (si_lo > s0) sA = s0 + 2 * ((si_lo - s0 + 1) / 2);
    int sB = si_hi - 1;                // last pair start s with s + 1 <= si_hi, s == s0 (mod 2)
    if (((sB - s0) & 1) != 0) sB -= 1;

    T H[C];
    int fD[C], fI[C];
    // B_LOCAL's end cell (_banded_find_optimal :395-408: every cell, diagonals ascending, first strict
    // maximum) needs, per diagonal, its best score and the earliest step that reached it.  BLOCAL == 1:
    // two registers per diagonal and compare + two selects per cell.  BLOCAL == 2 (host: integer scores
    // below 2^(31 - KEYBITS), n + m below 2^(KEYBITS + 1)): ONE key score * 2^KEYBITS + (2^KEYBITS - 1 -
    // (s - s0) / 2) per diagonal, one multiply-add and one max per cell -- its maximum is the highest
    // score at the earliest step.
    T best[NB];
    int bests[NBS];
#pragma unroll
    for (int c = 0; c < C; c++) { H[c] = NINF; fD[c] = 0; fI[c] = 0; }
#pragma unroll
    for (int c = 0; c < NB; c++) best[c] = (BLOCAL == 2) ? (T)-1 : NINF;
#pragma unroll
    for (int c = 0; c < NBS; c++) bests[c] = 0;
    // my 32-byte unit of block blk: unit0 + blk * (G * UW)
    uint32_t* unit0 = A.ptr + P.ptr_off + (long long)l * UW;

    int s = s0;
    // A careful step leaves its nibbles in cur[]; they go into the lane's staging unit in shared
    // memory (word-major: no bank conflicts) at step (s - s0) mod TB of the block, and a full block
    // (or the pair's last one) leaves with two 16-byte stores.
    // Every step leaves its WPS words in the lane's staging unit in shared memory (word-major: no
    // bank conflicts); a full block of TB steps (or the pair's last one) leaves with two 16-byte stores.
    auto flush_unit = [&](const int blk) {
      uint4 lo, hi;
      lo.x = s_unit[0][tid]; lo.y = s_unit[1][tid]; lo.z = s_unit[2][tid]; lo.w = s_unit[3][tid];
      hi.x = s_unit[4][tid]; hi.y = s_unit[5][tid]; hi.z = s_unit[6][tid]; hi.w = s_unit[7][tid];
      uint4* dst = (uint4*)(unit0 + (long long)blk * (G * UW));
      dst[0] = lo;
      dst[1] = hi;
    };
    auto careful_put = [&](const uint32_t (&cur)[WPS], const int sc) {
      const int rel = sc - s0, blk = rel / TB, i = rel - blk * TB;
#pragma unroll
      for (int w = 0; w < WPS; w++) s_unit[i * WPS + w][tid] = cur[w];
      if (i == TB - 1 || sc >= s_last) flush_unit(blk);
    };
#define BSQ_K2_CAREFUL_PAIR()                                                                         \
    {                                                                                                 \
      uint32_t cur[WPS];                                                                              \
      T eH = __shfl_up_sync(gmask, H[C - 1], 1, G);                                                   \
      int eF = __shfl_up_sync(gmask, fD[C - 1], 1, G);                                                \
      if (l == 0) { eH = NINF; eF = 0; }                                                              \
      bsq_k2_step_careful<S, C, BLOCAL, GOS, 0>(btype, H, fD, fI, best, bests, eH, eF, s_sub, Lt,     \
                                                oseq, mseq, n, m, s, d0, nreal, go, ge, cur,          \
                                                BSQ_K2_KSTEP(s, s0));                                 \
      careful_put(cur, s);                                                                            \
      eH = __shfl_down_sync(gmask, H[0], 1, G);                                                       \
      eF = __shfl_down_sync(gmask, fI[0], 1, G);                                                      \
      if (l == G - 1) { eH = NINF; eF = 0; }                                                          \
      bsq_k2_step_careful<S, C, BLOCAL, GOS, 1>(btype, H, fD, fI, best, bests, eH, eF, s_sub, Lt,     \
                                                oseq, mseq, n, m, s + 1, d0, nreal, go, ge, cur,      \
                                                BSQ_K2_KSTEP(s + 1, s0));                             \
      careful_put(cur, s + 1);                                                                        \
    }
    // careful steps up to the first block boundary at or past sA
    for (; (s < sA || ((s - s0) % TB) != 0) && s <= s_end; s += 2) BSQ_K2_CAREFUL_PAIR()

    // interior steps in whole blocks: pair starts s .. sBB, the last block ending at or before sB + 1
    const int sBB = (s + TB - 2 <= sB) ? s + ((sB + 2 - s) / TB) * TB - 2 : s - 2;
    if (s <= sBB) {
      // letter windows for the pair of steps (s, s+1); cell hc of the even step is
      // (x0 + hc, y0 - hc), of the odd step (x0 + 1 + hc, y0 - hc).
      // nibble j of ow: 4*o[x0+j-1], of mw: m[y0-j-1] (j < HC); onx: 4*o[x0+HC-1],
      // the nibble that slides into ow for the odd step
      int x0 = (s + d0) >> 1, y0 = (s - d0) >> 1;
      unsigned long long ow = 0ull, mw = 0ull;
      unsigned onx = 0u;
      if (packed) {
#pragma unroll
        for (int j = 0; j < HC; j++) {
          const int xi = x0 + j - 1;
          const unsigned v = ((unsigned)xi < (unsigned)n) ? (unsigned)oseq[xi] : 0u;
          ow |= (unsigned long long)(v << 2) << (4 * j);
        }
        {
          const int xi = x0 + HC - 1;
          onx = ((unsigned)xi < (unsigned)n) ? ((unsigned)oseq[xi] << 2) : 0u;
        }
#pragma unroll
        for (int j = 0; j < HC; j++) {
          const int yi = y0 - j - 1;
          const unsigned v = ((unsigned)yi < (unsigned)m) ? (unsigned)mseq[yi] : 0u;
          mw |= (unsigned long long)v << (4 * j);
        }
      }
      if (S::IS_INT) {   // H -> G = H + ge for the interior phase
#pragma unroll
        for (int c = 0; c < C; c++) H[c] = S::add(H[c], ge);
      }
      int blk = (s - s0) / TB, ib = 0;     // block of steps, step inside it (0 here: blocks are entered whole)
      for (; s <= sBB; s += 2) {
        unsigned no = 0u, nm = 0u;  // letters entering the windows after this pair
        if (packed) {
          const int xi = x0 + HC, yi = y0;
          if ((unsigned)xi < (unsigned)n) no = oseq[xi];
          if ((unsigned)yi < (unsigned)m) nm = mseq[yi];
        }
        const unsigned long long ow_odd = (ow >> 4) | ((unsigned long long)onx << (4 * (HC - 1)));
        uint32_t* pe = &s_unit[ib * WPS][tid];
        uint32_t* po = pe + WPS * BSQ_FILL_THREADS;
        T eH = __shfl_up_sync(gmask, H[C - 1], 1, G);
        int eF = __shfl_up_sync(gmask, fD[C - 1], 1, G);
        if (l == 0) { eH = NINF; eF = 0; }
        if (packed)
          bsq_k2_step_fast<S, C, BLOCAL, GOS, 0, true>(H, fD, fI, best, bests, eH, eF, s_fast, Lt, ow | mw, oseq,
                                                       mseq, s, d0, fixc, go, ge, pe, BSQ_K2_KSTEP(s, s0), n, m);
        else
          bsq_k2_step_fast<S, C, BLOCAL, GOS, 0, false>(H, fD, fI, best, bests, eH, eF, s_fast, Lt, 0ull, oseq,
                                                        mseq, s, d0, fixc, go, ge, pe, BSQ_K2_KSTEP(s, s0), n, m);
        eH = __shfl_down_sync(gmask, H[0], 1, G);
        eF = __shfl_down_sync(gmask, fI[0], 1, G);
        if (l == G - 1) { eH = NINF; eF = 0; }
        if (packed)
          bsq_k2_step_fast<S, C, BLOCAL, GOS, 1, true>(H, fD, fI, best, bests, eH, eF, s_fast, Lt, ow_odd | mw,
                                                       oseq, mseq, s + 1, d0, fixc, go, ge, po, BSQ_K2_KSTEP(s + 1, s0), n, m);
        else
          bsq_k2_step_fast<S, C, BLOCAL, GOS, 1, false>(H, fD, fI, best, bests, eH, eF, s_fast, Lt, 0ull, oseq,
                                                        mseq, s + 1, d0, fixc, go, ge, po, BSQ_K2_KSTEP(s + 1, s0), n, m);
        if (packed) {
          ow = ow_odd;
          onx = no << 2;
          mw = (mw << 4) | (unsigned long long)nm;
          x0++; y0++;
        }
        ib += 2;
        if (ib == TB) { flush_unit(blk); blk++; ib = 0; }
      }
      if (S::IS_INT) {   // back to H for the careful steps and the end-cell search
#pragma unroll
        for (int c = 0; c < C; c++) H[c] = S::add(H[c], (T)0 - ge);
      }
    }
    for (; s <= s_end; s += 2) BSQ_K2_CAREFUL_PAIR()
#undef BSQ_K2_CAREFUL_PAIR

    // end cell, _banded_find_optimal (:364-414): diagonals ascending, strict >
    T bs = NINF;
    int bdi = BSQ_INT_MAX, ba = 0;
    if (!BLOCAL && btype == BSQ_B_GLOBAL) {
      const int dg = (n - m) - dmin - l * C;  // my local index of the diagonal of (n,m)
#pragma unroll
      for (int c = 0; c < C; c++) if (c == dg) { bs = H[c]; bdi = l * C + c; }
    } else {
#pragma unroll
      for (int c = 0; c < C; c++) {
        if (c < nreal) {
          T cand = BLOCAL ? best[c < NB ? c : 0] : H[c];
          int step = BLOCAL == 1 ? bests[c < NBS ? c : 0] : 0;
          if (BLOCAL == 2) {   // unpack (score, step); no key: no cell of this diagonal in the table
            const int key = (int)cand;
            step = s0 + 2 * ((1 << BSQ_K2_KEYBITS) - 1 - (key & ((1 << BSQ_K2_KEYBITS) - 1))) + (c & 1);
            cand = key < 0 ? NINF : (T)(key >> BSQ_K2_KEYBITS);
          }
          if (cand > bs) { bs = cand; bdi = l * C + c; ba = step; }
        }
      }
    }
    bsq_group_best<T, G>(gmask, bs, bdi, ba);
    if (l == 0) {
      int ex = -1, ey = -1, st = BSQ_OK;
      if (bdi == BSQ_INT_MAX) st = BSQ_NO_ALIGNMENT;
      else {
        const int d = dmin + bdi;
        if (BLOCAL) { ex = (ba + d) >> 1; ey = (ba - d) >> 1; }  // ba = step s of the best cell
        else { ex = n < m + d ? n : m + d; ey = ex - d; }        // last cell of the diagonal (:383)
      }
      const int oi = P.pad;
      A.out_score[oi] = S::to_f64(bs);
      A.out_ex[oi] = ex;
      A.out_ey[oi] = ey;
      A.out_status[oi] = st;
    }
  }
}
