// bsq_k3.cuh -- K3: BANDED frame with a band of ANY width (config 5: one
// 1 Mb x 1 Mb pair, 4097 diagonals), intra-task wavefront.
//
// The band is cut into column strips of 32*K mutant positions.  One warp owns
// a strip and walks it like K1 walks a table: lane l owns K consecutive columns
// and lags its left neighbour by one row, so (x, y-1) arrives by one shuffle per
// row.  A strip only visits the rows its columns have inside the band
// (x in [Y+dmin, Y+32K-1+dmax]); cells outside band or table are forced to
// "does not exist" (-inf).  Strip j+1 consumes the right boundary column of strip
// j through a per-pair slab in global memory whose 8-byte entries validate
// themselves (each carries the producing strip's number, so no fence and no flag
// word is needed: an entry is either the old one or completely the new one), and
// fetches them 32 rows at a time, one batch ahead.  Successive strips thus run as
// a software pipeline across the SMs: about (W + 32K) / 32K strips are in flight.  Strips of all pairs of a
// chunk form ONE ordered work list claimed with an atomic counter, strip j of every pair before
// strip j+1 of any (BsqFillArgs::k3_order) -- a strip's predecessor is therefore always held by a
// resident warp (no deadlock) and the pipelines of all pairs run side by side.
//
// The single-pair case is bound by the dependency chain, not by ALU issue: the
// last strip cannot finish before ~m * (1 + 1/K) row steps have elapsed.
//
// Pointer layout (4 bit/cell): [strip][t][lane], one PU (K nibbles) per lane and
// time step t = x - x_lo(strip) + lane, i.e. skewed so that a warp writes one
// contiguous 32*sizeof(PU) segment per step.
//   replaces: dptable_solve's fill loop pw.c:59-110 over _xlim/_ylim
//   (_pw_internals.c:116-153) + _banded_find_optimal (:364-414)
#pragma once
#include "bsq_kernels.cuh"
#include "bsq_launch.h"

#ifdef BSQ_SIMT_EMU
#define BSQ_LD_CG(p) (*(p))
#define BSQ_ST_CG(p, v) (*(p) = (v))
#define BSQ_BACKOFF(ns) emu::yield()
#else
#define BSQ_LD_CG(p) __ldcg(p)
#define BSQ_ST_CG(p, v) __stcg((p), (v))
#define BSQ_BACKOFF(ns) __nanosleep(ns)
#endif

struct BsqFalse { static constexpr bool value = false; };
struct BsqTrue { static constexpr bool value = true; };

template <int K> struct BsqK3Unit;
template <> struct BsqK3Unit<2> { typedef uint8_t type; };
template <> struct BsqK3Unit<4> { typedef uint16_t type; };
template <> struct BsqK3Unit<8> { typedef uint32_t type; };

// one cell, band-gated: H = -inf when the cell is outside band/table
template <class S, int GOS, bool HASB>
__device__ __forceinline__ void bsq_k3_cell(typename S::T up, int gD, typename S::T left, int gI,
                                            typename S::T diag, typename S::T sig, typename S::T bval,
                                            typename S::T go, typename S::T ge, bool inb, typename S::T& H,
                                            int& fD, int& fI, uint32_t& nib) {
  typedef typename S::T T;
  T h;
  if constexpr (S::IS_INT) {
    int nB, nD, nI, nM;
    bsq_cell_i32<GOS, HASB, false>(up, gD, left, gI, diag, sig, bval, go, ge, h, nB, nD, nI, nM);
    nib = (uint32_t)(15 - (nB + 2 * nD + 4 * nI + 8 * nM));
    fD = bsq_gapflag_i32<GOS>(nD, nB, nI, nM);
    fI = bsq_gapflag_i32<GOS>(nI, nB, nD, nM);
  } else {
    bool eB, eD, eI, eM;
    bsq_cell2<S, GOS, HASB>(up, gD, left, gI, diag, sig, bval, go, ge, h, eB, eD, eI, eM);
    nib = (eB ? BSQ_MB : 0u) | (eD ? BSQ_MD : 0u) | (eI ? BSQ_MI : 0u) | (eM ? BSQ_MM : 0u);
    fD = bsq_gapflag<S, GOS>(eD, eB, eI, eM);
    fI = bsq_gapflag<S, GOS>(eI, eB, eD, eM);
  }
  H = inb ? h : S::ninf();
}

// Boundary-column slab entries: 64-bit words that are written and read whole.
//   I32: one word per row   = H (32) | opens (1) | tag (31)
//   F64: two words per row  = lo32(H) | tag (32)   and   hi32(H) | opens (1) | tag (31)
// tag = 1 + index of the strip that wrote the entry; a reader of strip j accepts tag == j.
template <class S> struct BsqK3Slab;
template <> struct BsqK3Slab<ScI32> {
  static constexpr int EPR = 1;
  struct Raw { unsigned long long e0; };
  static __device__ __forceinline__ void put(unsigned long long* row, int H, int opens, unsigned tag) {
    BSQ_ST_CG(row, (unsigned long long)(unsigned)H | ((unsigned long long)(opens ? 1u : 0u) << 32) |
                       ((unsigned long long)tag << 33));
  }
  // the same entry from the flag word the kernels carry (S::flag(): 0 / 1) and tag2 = 2 * tag
  static __device__ __forceinline__ void put_g(unsigned long long* row, int H, int g, unsigned tag2) {
    BSQ_ST_CG((uint2*)row, make_uint2((unsigned)H, tag2 + (unsigned)g));
  }
  static __device__ __forceinline__ Raw load(const unsigned long long* row) { Raw r; r.e0 = BSQ_LD_CG(row); return r; }
  static __device__ __forceinline__ bool decode(const Raw& r, unsigned tag, int& H, int& opens) {
    H = (int)(unsigned)r.e0;
    opens = (int)((r.e0 >> 32) & 1ull);
    return (unsigned)(r.e0 >> 33) == tag;
  }
};
template <> struct BsqK3Slab<ScF64> {
  static constexpr int EPR = 2;
  struct Raw { unsigned long long e0, e1; };
  static __device__ __forceinline__ void put(unsigned long long* row, double H, int opens, unsigned tag) {
    const unsigned long long b = (unsigned long long)__double_as_longlong(H);
    BSQ_ST_CG(row, (b & 0xffffffffull) | ((unsigned long long)tag << 32));
    BSQ_ST_CG(row + 1, (b >> 32) | ((unsigned long long)(2u * tag + (opens ? 1u : 0u)) << 32));
  }
  static __device__ __forceinline__ void put_g(unsigned long long* row, double H, int g, unsigned tag2) {  // g: 0 / 0x3FF00000
    BSQ_ST_CG((uint2*)row, make_uint2((unsigned)__double2loint(H), tag2 >> 1));
    BSQ_ST_CG((uint2*)(row + 1), make_uint2((unsigned)__double2hiint(H), tag2 + ((unsigned)g >> 29)));
  }
  static __device__ __forceinline__ Raw load(const unsigned long long* row) {
    Raw r; r.e0 = BSQ_LD_CG(row); r.e1 = BSQ_LD_CG(row + 1); return r;
  }
  static __device__ __forceinline__ bool decode(const Raw& r, unsigned tag, double& H, int& opens) {
    H = __longlong_as_double((long long)((r.e1 << 32) | (r.e0 & 0xffffffffull)));
    opens = (int)((r.e1 >> 32) & 1ull);
    return (unsigned)(r.e0 >> 32) == tag && (unsigned)(r.e1 >> 33) == tag;
  }
};

// BT = banded_alntype (BSQ_B_GLOBAL 0 / BSQ_B_LOCAL 1 / BSQ_B_OVERLAP 2)
template <class S, int K, int R, int BT, int GOS>
__global__ void __launch_bounds__(BSQ_K3_THREADS) bsq_k3_fill(const BsqFillArgs A) {
  static_assert(R == 1 || R == 2 || R == 4 || R == 8, "R must divide the 32-row boundary batch");
  typedef typename S::T T;
  typedef typename BsqK3Unit<K>::type PU;
  typedef BsqK3Slab<S> Slab;
  constexpr int G = 32, BW = G * K;
  constexpr int NB = (BT == BSQ_B_LOCAL) ? K : 1;
  BSQ_BAIL_IF_BAD(A);
  __shared__ T s_sub[BSQ_MAX_L * BSQ_MAX_L];
  __shared__ T s_inH[BSQ_K3_THREADS / 32][32];    // the batch of boundary values lane 0 consumes
  __shared__ int s_inG[BSQ_K3_THREADS / 32][32];
  const int L = A.L;
  for (int i = threadIdx.x; i < L * L; i += blockDim.x) s_sub[i] = ((const T*)A.subst)[i];
  __syncthreads();

  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const unsigned FULL = 0xffffffffu;
  const T NINF = S::ninf();
  const T ZERO = (T)0;
  const T go = S::go(A), ge = S::ge(A);
  unsigned long long* slab_all = (unsigned long long*)A.col_h;

  for (;;) {
    int g = 0;
    if (lane == 0) g = atomicAdd(A.counter, 1);
    g = __shfl_sync(FULL, g, 0);
    if (g >= A.k3_total) break;
#ifdef BSQ_K3_PROFILE
    long long pf_claim, pf_ready = 0, pf_wait = 0;
    int pf_polls = 0;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(pf_claim));
#endif
    const BsqK3Item item = A.k3_order[g];   // a strip's predecessor is always earlier in the list
    const int p = item.pair, j = item.strip;
    const BsqPair P = A.pairs[p];
    const BsqK3Pair Q = A.k3[p];
    const int n = P.n, m = P.m, dmin = P.dmin, dmax = P.dmax;
    const int Wm1 = dmax - dmin;
    const int Y = j * BW;
    const int ylast = (Y + BW - 1 < m) ? (Y + BW - 1) : m;
    int x_lo = Y + dmin;
    x_lo = x_lo < 0 ? 0 : x_lo;
    int x_hi = ylast + dmax;
    x_hi = x_hi > n ? n : x_hi;
    const bool has_succ = (j + 1 < Q.nstrips);
    // rows whose (x, Y-1) lies inside the band, i.e. the rows the predecessor strip wrote to
    // the slab: [pred_lo, pred_hi] (empty for strip 0)
    int pred_hi = -1;
    if (j > 0) { pred_hi = Y - 1 + dmax; pred_hi = pred_hi > n ? n : pred_hi; }
    int pred_lo = Y - BW + dmin;
    pred_lo = pred_lo < 0 ? 0 : pred_lo;
    unsigned long long* slab = slab_all + Q.slab_off * Slab::EPR;
    const unsigned tag_in = (unsigned)j, tag_out = (unsigned)j + 1u;

    T bs = NINF;                     // lane's best end-cell candidate of this strip
    int bk1 = BSQ_INT_MAX, bk2 = BSQ_INT_MAX;

    if (x_lo <= x_hi) {
      const int y0 = Y + lane * K;
      int moff[K];
#pragma unroll
      for (int k = 0; k < K; k++) {
        const int y = y0 + k;
        moff[k] = (y >= 1 && y <= m) ? (int)A.codes[P.m_off + y - 1] : 0;
      }
      int kv = m - y0 + 1;           // my columns inside the table
      kv = kv < 0 ? 0 : (kv > K ? K : kv);
      const int km = m - y0;         // which of my columns is column m (if in [0,K))
      T Hup[K];
      int gD[K];
      T best[NB];
      int bestx[NB];
#pragma unroll
      for (int k = 0; k < K; k++) { Hup[k] = NINF; gD[k] = 0; }
#pragma unroll
      for (int k = 0; k < NB; k++) { best[k] = NINF; bestx[k] = 0; }
      T lastH[R];                    // my rightmost column's cells of the rows of the last step
      int lastG[R];
#pragma unroll
      for (int r = 0; r < R; r++) { lastH[r] = NINF; lastG[r] = 0; }
      T prev_in = NINF;
      if (lane == 0 && j > 0) {      // (x_lo - 1, Y - 1): the diagonal neighbour of my first cell
        const int xd = x_lo - 1;
        if (xd >= pred_lo && xd <= pred_hi) {
          unsigned ns = 32;
          int f;
          while (!Slab::decode(Slab::load(slab + (long long)xd * Slab::EPR), tag_in, prev_in, f)) { BSQ_BACKOFF(ns); if (ns < 1024) ns *= 2; }
        }
      }
      // A time step is R rows: lane l does rows x_lo + (t - l) R + r, r < R, of its K columns, i.e.
      // lags its left neighbour by R rows.  The pointer layout does not depend on R.
      const int nsteps = (x_hi - x_lo + R) / R + G - 1;
      PU* prow = (PU*)(A.ptr + P.ptr_off) + (long long)j * P.s0 * G + lane;
      const uint8_t* oseq = A.codes + P.o_off;
      // software pipeline of the letter / score loads: row x uses o[x-1]
      auto letter = [&](int xr) -> int { return (xr >= 1 && xr <= n) ? (int)oseq[xr - 1] : 0; };
      int o1[R];                     // letters of the rows of the next step
      T sig[R][K];
#pragma unroll
      for (int r = 0; r < R; r++) {
        o1[r] = letter(x_lo - lane * R + R + r);
        const int ob = letter(x_lo - lane * R + r) * L;
#pragma unroll
        for (int k = 0; k < K; k++) sig[r][k] = s_sub[ob + moff[k]];
      }
      typename Slab::Raw nx = {};    // prefetched (raw) boundary entry of row x_lo + (next batch) + lane
      bool nx_have = false;

      // One time step.  INT (interior): every cell of every lane lies inside band and table, no lane
      // is on row 0 / row n / column m and the letters two steps ahead exist -- all gating, the B
      // rule of the edges and the end-cell tests fold away (the same split as K2's three phases),
      // and the R x K block is computed in anti-diagonal order: the cells of one anti-diagonal do
      // not depend on each other, which is the only instruction-level parallelism a lone warp has.
      auto step = [&](const int t, auto interior_tag) {
        constexpr bool INT = decltype(interior_tag)::value;
        const int x0 = x_lo + (t - lane) * R;
        if (((t * R) & 31) == 0 && j > 0) {
          // boundary values of rows x_lo + tR .. x_lo + tR + 31: mine is row xr (fetched one batch
          // ahead); rows past pred_hi lie outside the band
          const int xr = x_lo + t * R + lane;
          const bool need = (xr <= pred_hi);
          T bH = NINF;
          int bF = 0;
          bool ok = !need || (nx_have && Slab::decode(nx, tag_in, bH, bF));
          unsigned ns = 32;
#ifdef BSQ_K3_PROFILE
          const long long pf_c0 = clock64();
#endif
          while (!__all_sync(FULL, ok)) {
            if (!ok) {
              ok = Slab::decode(Slab::load(slab + (long long)xr * Slab::EPR), tag_in, bH, bF);
              if (!ok) { BSQ_BACKOFF(ns); if (ns < 1024) ns *= 2; }
            }
#ifdef BSQ_K3_PROFILE
            pf_polls++;
#endif
          }
#ifdef BSQ_K3_PROFILE
          pf_wait += clock64() - pf_c0;
          if (t == 0) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(pf_ready));
#endif
          s_inH[wib][lane] = need ? bH : NINF;
          s_inG[wib][lane] = need ? S::flag(bF == 0) : 0;
          __syncwarp();
          const int xn = xr + 32;
          nx_have = (xn <= pred_hi);
          if (nx_have) nx = Slab::load(slab + (long long)xn * Slab::EPR);
        }
        T inH[R];
        int inG[R];
#pragma unroll
        for (int r = 0; r < R; r++) {
          inH[r] = __shfl_up_sync(FULL, lastH[r], 1);
          inG[r] = __shfl_up_sync(FULL, lastG[r], 1);
        }
        if (lane == 0) {
#pragma unroll
          for (int r = 0; r < R; r++) {
            inH[r] = NINF;
            inG[r] = 0;
            if (j > 0) { inH[r] = s_inH[wib][(t * R + r) & 31]; inG[r] = s_inG[wib][(t * R + r) & 31]; }
          }
        }
        // loads for the next two steps, off the dependency chain
        int o2[R];
        T sign[R][K];
#pragma unroll
        for (int r = 0; r < R; r++) {
          o2[r] = INT ? (int)oseq[x0 + 2 * R + r - 1] : letter(x0 + 2 * R + r);
          const int ob = o1[r] * L;
#pragma unroll
          for (int k = 0; k < K; k++) sign[r][k] = s_sub[ob + moff[k]];
        }
        if (INT) {
          constexpr bool HB = (BT == BSQ_B_LOCAL);
          T left[R], diag[R];
          int gI[R];
          uint32_t cw[R];
#pragma unroll
          for (int r = 0; r < R; r++) { left[r] = inH[r]; gI[r] = inG[r]; diag[r] = r == 0 ? prev_in : inH[r - 1]; cw[r] = 0u; }
#pragma unroll
          for (int d = 0; d < R + K - 1; d++) {
#pragma unroll
            for (int r = 0; r < R; r++) {
              const int k = d - r;
              if (k >= 0 && k < K) {
                const T up = Hup[k];
                T H;
                int fD, fI;
                uint32_t nib;
                bsq_k3_cell<S, GOS, HB>(up, gD[k], left[r], gI[r], diag[r], sig[r][k], ZERO, go, ge, true, H, fD, fI, nib);
                cw[r] |= nib << (4 * k);
                gD[k] = fD; gI[r] = fI; diag[r] = up; left[r] = H; Hup[k] = H;
                if (BT == BSQ_B_LOCAL) { if (H > best[k < NB ? k : 0]) { best[k < NB ? k : 0] = H; bestx[k < NB ? k : 0] = x0 + r; } }
              }
            }
          }
#pragma unroll
          for (int r = 0; r < R; r++) {
            lastH[r] = left[r];
            lastG[r] = gI[r];
            prow[(long long)(t * R + r - lane * (R - 1)) * G] = (PU)cw[r];
            if (lane == G - 1 && has_succ) Slab::put(slab + (long long)(x0 + r) * Slab::EPR, left[r], gI[r] != 0, tag_out);
          }
        } else {
#pragma unroll
          for (int r = 0; r < R; r++) {
            const int x = x0 + r;
            const bool active = (t >= lane) && (x <= x_hi);
            if (active) {
              const int e = x - y0 - dmin;   // cell k is inside the band iff 0 <= e - k <= W - 1
              // B (_alnchoice_B, _pw_internals.c:182-197): B_LOCAL anywhere, B_OVERLAP on x==0 or y==0,
              // B_GLOBAL at (0,0) only
              const bool rowb = (BT == BSQ_B_LOCAL) || (BT == BSQ_B_OVERLAP && x == 0);
              const T brest = rowb ? ZERO : NINF;
              const T bfirst = (rowb || (y0 == 0 && (BT == BSQ_B_OVERLAP || x == 0))) ? ZERO : NINF;
              T left = inH[r];
              int gI = inG[r];
              T diag = r == 0 ? prev_in : inH[r - 1];
              uint32_t cw = 0u;
              if (rowb) {
#pragma unroll
                for (int k = 0; k < K; k++) {
                  const bool inb = ((unsigned)(e - k) <= (unsigned)Wm1) && (k < kv);
                  const T up = Hup[k];
                  T H;
                  int fD, fI;
                  uint32_t nib;
                  bsq_k3_cell<S, GOS, true>(up, gD[k], left, gI, diag, sig[r][k], k == 0 ? bfirst : brest, go, ge, inb,
                                            H, fD, fI, nib);
                  cw |= nib << (4 * k);
                  gD[k] = fD; gI = fI; diag = up; left = H; Hup[k] = H;
                  if (BT == BSQ_B_LOCAL) { if (H > best[k < NB ? k : 0]) { best[k < NB ? k : 0] = H; bestx[k < NB ? k : 0] = x; } }
                }
              } else {
#pragma unroll
                for (int k = 0; k < K; k++) {
                  const bool inb = ((unsigned)(e - k) <= (unsigned)Wm1) && (k < kv);
                  const T up = Hup[k];
                  T H;
                  int fD, fI;
                  uint32_t nib;
                  if (k == 0) bsq_k3_cell<S, GOS, true>(up, gD[k], left, gI, diag, sig[r][k], bfirst, go, ge, inb, H, fD, fI, nib);
                  else bsq_k3_cell<S, GOS, false>(up, gD[k], left, gI, diag, sig[r][k], ZERO, go, ge, inb, H, fD, fI, nib);
                  cw |= nib << (4 * k);
                  gD[k] = fD; gI = fI; diag = up; left = H; Hup[k] = H;
                }
              }
              lastH[r] = left;
              lastG[r] = gI;
              prow[(long long)(x - x_lo + lane) * G] = (PU)cw;
              if (lane == G - 1 && has_succ) Slab::put(slab + (long long)x * Slab::EPR, left, gI != 0, tag_out);
              // end cell, _banded_find_optimal (_pw_internals.c:364-414)
              if (BT == BSQ_B_GLOBAL) {
                if (x == n && km >= 0 && km < K) {
#pragma unroll
                  for (int k = 0; k < K; k++) if (k == km) { bs = Hup[k]; bk1 = n - m; bk2 = n; }
                }
              } else if (BT == BSQ_B_OVERLAP) {  // last cell of every diagonal, diagonals ascending, strict >
                if (x == n) {
#pragma unroll
                  for (int k = 0; k < K; k++) {
                    const T c = Hup[k];
                    if (c > NINF && bsq_better<T>(c, n - (y0 + k), 0, bs, bk1, bk2)) { bs = c; bk1 = n - (y0 + k); bk2 = 0; }
                  }
                } else if (km >= 0 && km < K) {
                  T c = NINF;
#pragma unroll
                  for (int k = 0; k < K; k++) if (k == km) c = Hup[k];
                  if (c > NINF && bsq_better<T>(c, x - m, 0, bs, bk1, bk2)) { bs = c; bk1 = x - m; bk2 = 0; }
                }
              }
            }
          }
        }
        if (INT || t >= lane) prev_in = inH[R - 1];
#pragma unroll
        for (int r = 0; r < R; r++) {
          o1[r] = o2[r];
#pragma unroll
          for (int k = 0; k < K; k++) sig[r][k] = sign[r][k];
        }
      };
      // interior steps [t_a, t_b]: with c0 = x_lo - (Y + dmin) >= 0, cell (r, k) of lane l sits at
      // e - k = c0 + t R + r - l (R + K) - k in the band: >= 0 for lane 31, <= W - 1 for lane 0; lane 0
      // (the lowest rows) stays two steps clear of n, lane 31 is past row 0; the strip lies strictly
      // left of column m; B_OVERLAP's column 0 (strip 0) is never interior
      int t_a = nsteps, t_b = -1;
      if (Y + BW - 1 < m && !(BT == BSQ_B_OVERLAP && j == 0)) {
        const int c0 = x_lo - (Y + dmin);
        t_a = (31 * (R + K) + K - 1 - c0 + R - 1) / R;
        t_a = t_a < 32 ? 32 : t_a;
        t_b = (Wm1 - c0 - (R - 1)) / R;
        t_b = ((n - 3 * R + 1 - x_lo) / R < t_b) ? (n - 3 * R + 1 - x_lo) / R : t_b;
        t_b = ((x_hi - x_lo - (R - 1)) / R < t_b) ? (x_hi - x_lo - (R - 1)) / R : t_b;
        if (t_b < t_a) { t_a = nsteps; t_b = -1; }
      }
      {
        int t = 0;
        for (; t < t_a && t < nsteps; t++) step(t, BsqFalse{});
        for (; t <= t_b; t++) step(t, BsqTrue{});
        for (; t < nsteps; t++) step(t, BsqFalse{});
      }
      if (BT == BSQ_B_LOCAL) {  // every cell, (d, a) ascending; per column d ascends with x
#pragma unroll
        for (int k = 0; k < NB; k++) {
          const T c = best[k];
          const int d = bestx[k] - (y0 + k);
          if (c > NINF && bsq_better<T>(c, d, bestx[k], bs, bk1, bk2)) { bs = c; bk1 = d; bk2 = bestx[k]; }
        }
      }
    }

    bsq_group_best<T, 32>(FULL, bs, bk1, bk2);
    int last = 0;
    if (lane == 0) {
      BsqK3Res r;
      r.score = S::to_f64(bs);
      r.k1 = bk1;
      r.k2 = bk2;
#ifdef BSQ_K3_PROFILE
      r.t_claim = pf_claim; r.t_ready = pf_ready; r.wait_clk = pf_wait; r.polls = pf_polls;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(r.t_end));
      asm volatile("mov.u32 %0, %%smid;" : "=r"(r.sm));
#endif
      A.k3_res[Q.strip_base + j] = r;
      __threadfence();
      last = (atomicAdd(A.k3_done + p, 1) == Q.nstrips - 1);
    }
    last = __shfl_sync(FULL, last, 0);
    if (last) {  // this warp finished the pair's last open strip: pick the end cell
      __threadfence();
      double fs = ScF64::ninf();
      int f1 = BSQ_INT_MAX, f2 = BSQ_INT_MAX;
      for (int i = lane; i < Q.nstrips; i += 32) {
        const BsqK3Res* rp = A.k3_res + Q.strip_base + i;
        const double sc = BSQ_LD_CG(&rp->score);
        const int k1 = BSQ_LD_CG(&rp->k1), k2 = BSQ_LD_CG(&rp->k2);
        if (k1 != BSQ_INT_MAX && (f1 == BSQ_INT_MAX || bsq_better<double>(sc, k1, k2, fs, f1, f2))) { fs = sc; f1 = k1; f2 = k2; }
      }
      bsq_group_best<double, 32>(FULL, fs, f1, f2);
      if (lane == 0) {
        int ex = -1, ey = -1, st = BSQ_OK;
        if (f1 == BSQ_INT_MAX) { st = BSQ_NO_ALIGNMENT; fs = 0.0; }
        else if (BT == BSQ_B_LOCAL) { ex = f2; ey = f2 - f1; }
        else if (BT == BSQ_B_GLOBAL) { ex = n; ey = m; }
        else { ex = n < m + f1 ? n : m + f1; ey = ex - f1; }
        const int oi = P.pad;
        A.out_score[oi] = fs;
        A.out_ex[oi] = ex;
        A.out_ey[oi] = ey;
        A.out_status[oi] = st;
      }
    }
  }
}


// ===========================================================================
// K3L: the strip pipeline tuned for LATENCY -- one (or a few) pairs, where the fill time is the
// length of the dependency chain and not the cell count (config 5: one 1 Mb x 1 Mb pair; config 1:
// one 1 kb x 1 kb STD GLOBAL pair, planned as a full-width band).  With strips of BW columns, rows
// of lag R per lane and a boundary hand-over every B rows, strip j+1 starts
//     (BW + B) / R + 31  steps
// after strip j, and the whole fill takes m / BW of those lags: only the FIRST ~100 steps of every
// strip (its ramp: lanes entering the band one after the other) are on the critical path, each run
// by a lone warp that issues in order.  So the step is built for instruction count and has no
// branch (a lone warp pays 25 - 50 cycles to resolve one):
//   * 64-column strips (K = 2), one row per step, hand-over every 8 rows;
//   * groups of 8 unrolled steps in two flavours picked by one uniform test per group -- LEAN
//     (every lane is above the lower end of its columns for the whole group and the group holds
//     no cell where an alignment may begin: cells are gated by "t >= first step" alone, pointer
//     bytes are stored unconditionally -- slots of idle lanes exist in the layout and are never
//     read --, no end-cell bookkeeping) and FULL (everything gated; first group of a GLOBAL pair,
//     the first rows of OVERLAP, the tail of every strip);
//   * the shuffle is a rotation: lane 31 sends the NEXT boundary value (staged in shared memory by
//     the batch) instead of its own, so lane 0 needs no special case;
//   * the origin letter of a row enters with the boundary batch and rides the skew one step
//     ahead of its row, pre-shifted into a byte offset that is OR-ed with the lane's column base
//     in a transposed, power-of-two-padded score table: one LOP3 + one LDS per cell, off the chain.
// Same pointer layout ([strip][t][lane], one byte = two nibbles) and slab protocol as bsq_k3_fill
// with K = 2, so bsq_k3_walk<2> serves both.
//   replaces: dptable_solve's fill loop pw.c:59-110 over _xlim/_ylim
//   (_pw_internals.c:116-153) + _banded_find_optimal (:364-414)
// ===========================================================================
#define BSQ_K3L_BATCH 8
// flavours of a group of BSQ_K3L_BATCH steps
#define BSQ_K3L_FULL 0    // everything gated and tracked
#define BSQ_K3L_RAMP 1    // columns still beginning: gated by their first step
#define BSQ_K3L_INNER 2   // every column inside: nothing gated
#define BSQ_K3L_TAIL 3    // columns ending: gated by their last step, last values kept
template <int V> struct BsqInt { static constexpr int value = V; };

#ifdef BSQ_SIMT_EMU
#define BSQ_SH_ALIGN(n)
#else
#define BSQ_SH_ALIGN(n) __align__(n)
#endif

// score table entry at byte offset `a` (device: `a` already holds the table's shared-window address)
template <class T>
__device__ __forceinline__ T bsq_lds_off(const T* table, uint32_t a) {
#ifdef BSQ_SIMT_EMU
  return *(const T*)((const char*)table + a);
#else
  (void)table;
  T v;
  if constexpr (sizeof(T) == 8) asm("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a));
  else asm("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(a));
  return v;
#endif
}

// one cell, ungated.  HB: 0 = no B candidate, 1 = bval (0 or -inf), 2 = B allowed everywhere
template <class S, int GOS, int HB>
__device__ __forceinline__ void bsq_k3l_cell(typename S::T up, int gD, typename S::T left, int gI,
                                             typename S::T diag, typename S::T sig, typename S::T bval,
                                             typename S::T go, typename S::T ge, typename S::T& h,
                                             int& fD, int& fI, uint32_t& nib) {
  typedef typename S::T T;
  if constexpr (S::IS_INT) {
    int nB, nD, nI, nM;
    bsq_cell_i32<GOS, HB != 0, HB == 2>(up, gD, left, gI, diag, sig, bval, go, ge, h, nB, nD, nI, nM);
    nib = (uint32_t)(15 - (nB + 2 * nD + 4 * nI + 8 * nM));
    fD = bsq_gapflag_i32<GOS>(nD, nB, nI, nM);
    fI = bsq_gapflag_i32<GOS>(nI, nB, nD, nM);
  } else {
    bool eB, eD, eI, eM;
    bsq_cell2<S, GOS, HB != 0>(up, gD, left, gI, diag, sig, HB == 2 ? (T)0 : bval, go, ge, h, eB, eD, eI, eM);
    nib = (eB ? BSQ_MB : 0u) | (eD ? BSQ_MD : 0u) | (eI ? BSQ_MI : 0u) | (eM ? BSQ_MM : 0u);
    fD = bsq_gapflag<S, GOS>(eD, eB, eI, eM);
    fI = bsq_gapflag<S, GOS>(eI, eB, eD, eM);
  }
}

// (minBlocks = 1 in the launch bounds is a scheduling hint with a measured effect: without it ptxas
// keeps the kernel at 56 - 64 registers and puts the consumers of the two shuffles 4 - 6 instructions
// behind them; with it, it spends ~100 registers and 10 - 20 instructions of independent work in
// between -- config 5 fills in 119 ms instead of 139.  tools/shfl_distance.py reads that distance
// off the SASS.)
template <class S, int BT, int GOS>
__global__ void __launch_bounds__(BSQ_K3_THREADS, 1) bsq_k3l_fill(const BsqFillArgs A) {
  typedef typename S::T T;
  typedef uint8_t PU;
  typedef BsqK3Slab<S> Slab;
  constexpr int K = 2, G = 32, BW = G * K, NBT = BSQ_K3L_BATCH;
  constexpr int NB = (BT == BSQ_B_LOCAL) ? K : 1;
  // the word lanes pass along: open/extend flag of the sender's right cell | (origin letter of
  // the receiver's NEXT row) << LSH, i.e. the letter as a byte offset into a score-table column
  constexpr int LSH = (sizeof(T) == 8) ? 3 : 2;
  constexpr int LETMASK = (BSQ_MAX_L - 1) << LSH;
  BSQ_BAIL_IF_BAD(A);
  // transposed and padded: entry (origin letter o, mutant letter c) at [(c << lp) | o], 2^lp >= L
  __shared__ BSQ_SH_ALIGN(BSQ_MAX_L * BSQ_MAX_L * sizeof(T)) T s_subT[BSQ_MAX_L * BSQ_MAX_L];
  // boundary ring, filled by the feeder warp: slot (r mod RING) holds row x_lo + r of the strip's
  // left boundary column: H of (row, Y - 1) and its flag | the letter of row + 1
  constexpr int RING = 8 * NBT;
  __shared__ T s_rH[RING];
  __shared__ int s_rP[RING];
  __shared__ int s_ctl[4];        // [0] claimed strip, [1] rows staged (feeder), [2] groups consumed (worker)
  volatile int* ctl = s_ctl;
  const int L = A.L;
  int lp = 0;
  while ((1 << lp) < L) lp++;
  for (int i = threadIdx.x; i < L * L; i += blockDim.x) s_subT[((i % L) << lp) | (i / L)] = ((const T*)A.subst)[i];
  __syncthreads();
#ifdef BSQ_SIMT_EMU
  const uint32_t tbase = 0u;
#else
  const uint32_t tbase = (uint32_t)__cvta_generic_to_shared(s_subT);
#endif

  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int srcl = (lane + G - 1) & (G - 1);
  const bool lane31 = (lane == G - 1);
  const unsigned FULL = 0xffffffffu;
  const T NINF = S::ninf();
  const T ZERO = (T)0;
  const T go = S::go(A), ge = S::ge(A);
  unsigned long long* slab_all = (unsigned long long*)A.col_h;

  for (;;) {
    __syncthreads();                 // both warps are done with the previous strip and its ring
    if (threadIdx.x == 0) { s_ctl[0] = atomicAdd(A.counter, 1); s_ctl[1] = 0; s_ctl[2] = 0; }
    __syncthreads();
    const int g = s_ctl[0];
    if (g >= A.k3_total) break;
#ifdef BSQ_K3_PROFILE
    long long pf_claim, pf_ready = 0, pf_wait = 0;
    int pf_polls = 0;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(pf_claim));
#endif
    const BsqK3Item item = A.k3_order[g];   // a strip's predecessor is always earlier in the list
    const int p = item.pair, j = item.strip;
    const BsqPair P = A.pairs[p];
    const BsqK3Pair Q = A.k3[p];
    const int n = P.n, m = P.m, dmin = P.dmin, dmax = P.dmax;
    const int Y = j * BW;
    const int ylast = (Y + BW - 1 < m) ? (Y + BW - 1) : m;
    int x_lo = Y + dmin;
    x_lo = x_lo < 0 ? 0 : x_lo;
    int x_hi = ylast + dmax;
    x_hi = x_hi > n ? n : x_hi;
    const bool has_succ = (j + 1 < Q.nstrips);
    int pred_hi = -1;                // rows the predecessor strip wrote to the slab: [pred_lo, pred_hi]
    if (j > 0) { pred_hi = Y - 1 + dmax; pred_hi = pred_hi > n ? n : pred_hi; }
    int pred_lo = Y - BW + dmin;
    pred_lo = pred_lo < 0 ? 0 : pred_lo;
    unsigned long long* slab = slab_all + Q.slab_off * Slab::EPR;
    const unsigned tag_in = (unsigned)j, tag2_out = 2u * ((unsigned)j + 1u);

    if (wib != 0) {
      // The feeder warp: polls the predecessor's boundary column in the slab (32 rows per round,
      // rows past pred_hi lie outside the band; strip 0 has no predecessor: pred_hi = -1), stages
      // the contiguous valid prefix with the origin letters into the ring and publishes the count.
      // The worker warp never touches global memory for its inputs.
      if (x_lo <= x_hi) {
        const int nsteps = (x_hi - x_lo + 1) + G - 1;
        const int nrows = ((nsteps + NBT - 1) / NBT) * NBT;
        const uint8_t* oseq = A.codes + P.o_off;
        const int nm1 = n > 0 ? n - 1 : 0;
        int r0 = 0;
        int idle = 0;
        while (r0 < nrows) {
          const int r = r0 + lane;
          const int xr = x_lo + r;
          const bool need = (r < nrows) && (xr <= pred_hi);
          const int let = (int)oseq[xr < nm1 ? xr : nm1];        // the letter of row xr + 1 (rows past n: unused)
          T bH = NINF;
          int bF = 0;
          bool ok = true;
          if (need) ok = Slab::decode(Slab::load(slab + (long long)xr * Slab::EPR), tag_in, bH, bF);
          const unsigned okm = __ballot_sync(FULL, ok);
          int c = (okm == FULL) ? 32 : (__ffs((int)~okm) - 1);
          const int room = __shfl_sync(FULL, ctl[2], 0) * NBT + RING - r0;   // slots the worker has released (one read: uniform)
          c = c < room ? c : room;
          c = c < nrows - r0 ? c : nrows - r0;
          if (c > 0) {
            if (lane < c) {
              s_rH[r & (RING - 1)] = need ? bH : NINF;
              s_rP[r & (RING - 1)] = (need ? S::flag(bF == 0) : 0) | (let << LSH);
            }
            __threadfence_block();
            __syncwarp();
            r0 += c;
            if (lane == 0) ctl[1] = r0;
            idle = 0;
          } else {
#ifdef BSQ_SIMT_EMU
            emu::yield();
#else
            if (++idle > 64) __nanosleep(100);
#endif
          }
        }
      }
      continue;
    }

    T bs = NINF;                     // lane's best end-cell candidate of this strip
    int bk1 = BSQ_INT_MAX, bk2 = BSQ_INT_MAX;

    if (x_lo <= x_hi) {
      const int y0 = Y + lane * K;
      int mlet[K];
      int ta[K];                     // column k is inside band, table and strip rows for steps ta .. ta + span
      unsigned span[K];
      int xlast[K];                  // its last such row (or -1)
      uint32_t colb[K];              // byte address of its score-table column
      int tend[K];                   // ta + span (or -1)
      int thi = 0x7fffffff;          // the last step at which every existing column of the strip is still inside
      int tmax_a = 0;                // the first step at which every existing column has begun
      bool allcols = true;           // all 64 columns exist
#pragma unroll
      for (int k = 0; k < K; k++) {
        const int y = y0 + k;
        mlet[k] = (y >= 1 && y <= m) ? (int)A.codes[P.m_off + y - 1] : 0;
        colb[k] = tbase + ((uint32_t)mlet[k] << (lp + LSH));
        int xa = y + dmin, xb = y + dmax;
        xa = xa < x_lo ? x_lo : xa;
        xb = xb > x_hi ? x_hi : xb;
        const bool any = (y <= m) && (xa <= xb);
        ta[k] = any ? (xa - x_lo + lane) : 0x40000000;
        span[k] = any ? (unsigned)(xb - xa) : 0u;
        xlast[k] = any ? xb : -1;
        tend[k] = any ? ta[k] + (xb - xa) : -1;
        if (any) { thi = min(thi, tend[k]); tmax_a = max(tmax_a, ta[k]); }
        allcols = allcols && any;
      }
#pragma unroll
      for (int d = 16; d >= 1; d >>= 1) {
        thi = min(thi, __shfl_xor_sync(FULL, thi, d));
        tmax_a = max(tmax_a, __shfl_xor_sync(FULL, tmax_a, d));
      }
      allcols = __all_sync(FULL, allcols);
      // LEAN groups: no cell where an alignment may begin other than "everywhere" (B_LOCAL), no
      // column m above the last row (B_OVERLAP's other end), every cell below its column's last step
      bool lean_strip = true;
      int tlo_lean = 0;
      if (BT == BSQ_B_GLOBAL) tlo_lean = (j == 0) ? NBT : 0;                       // (0, 0)
      if (BT == BSQ_B_OVERLAP) { lean_strip = (j > 0) && (Y + BW - 1 < m); tlo_lean = (x_lo == 0) ? G : 0; }  // y == 0, x == 0
      const int km = m - y0;         // which of my columns is column m (if in [0,K))
      T Hup[K], Hend[K];
      int gD[K];
      T best[NB];
      int bestt[NB];
#pragma unroll
      for (int k = 0; k < K; k++) { Hup[k] = NINF; Hend[k] = NINF; gD[k] = 0; }
#pragma unroll
      for (int k = 0; k < NB; k++) { best[k] = NINF; bestt[k] = 0; }
      T lastH = NINF;                // what I send: my right column's cell of the previous step (lane 31: the
      int mypay = 0;                 // next boundary value), and its flag | letter word
      T prev_in = NINF;
      if (lane == 0 && j > 0) {      // (x_lo - 1, Y - 1): the diagonal neighbour of my first cell
        const int xd = x_lo - 1;
        if (xd >= pred_lo && xd <= pred_hi) {
          unsigned ns = 32;
          int f;
          while (!Slab::decode(Slab::load(slab + (long long)xd * Slab::EPR), tag_in, prev_in, f)) { BSQ_BACKOFF(ns); if (ns < 256) ns *= 2; }
        }
      }
      const int nsteps = (x_hi - x_lo + 1) + G - 1;
      const uint8_t* oseq = A.codes + P.o_off;
      T sig[K];
      {
        const int curlet = (lane == 0 && x_lo >= 1) ? (int)oseq[x_lo - 1] : 0;   // x_lo <= n
#pragma unroll
        for (int k = 0; k < K; k++) sig[k] = s_subT[(mlet[k] << lp) | curlet];
      }
      T* rH = s_rH;                  // ring slots of the current group
      int* rP = s_rP;

      PU* pst = (PU*)(A.ptr + P.ptr_off) + (long long)j * P.s0 * G + lane;       // + t * G
      unsigned long long* pslab = slab + (long long)(x_lo - lane) * Slab::EPR;     // slab + x * EPR
      const bool put31 = lane31 && has_succ;
      T gor = go, ger = ge;
#ifndef BSQ_SIMT_EMU
      // keep the loop's scalars in registers (the compiler would re-read the constant bank every
      // step to save registers nobody else wants: one warp per scheduler)
      if constexpr (S::IS_INT) asm volatile("" : "+r"(gor), "+r"(ger));
      else asm volatile("" : "+d"(gor), "+d"(ger));
#endif

      auto group = [&](const int t0, auto flavour) {
        constexpr int FL = decltype(flavour)::value;
        constexpr bool FULLG = (FL == BSQ_K3L_FULL);
        // The two shuffles of a step are issued as soon as the values they send are final -- before
        // the pointer byte, the slab entry and the end-cell bookkeeping of the step they follow -- so
        // that this off-chain work runs while they are in flight (the lone warp issues in order).
        T inH = __shfl_sync(FULL, lastH, srcl);
        int pay = __shfl_sync(FULL, mypay, srcl);
#pragma unroll
        for (int ti = 0; ti < NBT; ti++) {
          const int t = t0 + ti;           // steps past nsteps - 1 (the last group) are inactive for every lane
          const uint32_t lo = (uint32_t)(pay & LETMASK);
          int gI = pay & ~LETMASK;
          T sign[K];
#pragma unroll
          for (int k = 0; k < K; k++) sign[k] = bsq_lds_off<T>(s_subT, lo | colb[k]);
          // B (_alnchoice_B, _pw_internals.c:182-197): B_LOCAL anywhere, B_OVERLAP on x==0 or y==0,
          // B_GLOBAL at (0,0) only
          const int x = x_lo + t - lane;
          T bfirst = NINF, brest = NINF;
          if (FULLG) {
            const bool rowb = (BT == BSQ_B_LOCAL) || (BT == BSQ_B_OVERLAP && x == 0);
            brest = rowb ? ZERO : NINF;
            bfirst = (rowb || (y0 == 0 && (BT == BSQ_B_OVERLAP || x == 0))) ? ZERO : NINF;
          }
          T left = inH, diag = prev_in;
          uint32_t cw = 0u;
#pragma unroll
          for (int k = 0; k < K; k++) {
            const bool inb = FL == BSQ_K3L_RAMP ? (t >= ta[k]) : FL == BSQ_K3L_TAIL ? (t <= tend[k])
                           : FL == BSQ_K3L_INNER ? true : ((unsigned)(t - ta[k]) <= span[k]);
            const T up = Hup[k];
            T h;
            int fD, fI;
            uint32_t nib;
            if (!FULLG)
              bsq_k3l_cell<S, GOS, (BT == BSQ_B_LOCAL) ? 2 : 0>(up, gD[k], left, gI, diag, sig[k], ZERO, gor, ger, h, fD, fI, nib);
            else if (k == 0 || BT != BSQ_B_GLOBAL)
              bsq_k3l_cell<S, GOS, 1>(up, gD[k], left, gI, diag, sig[k], k == 0 ? bfirst : brest, gor, ger, h, fD, fI, nib);
            else
              bsq_k3l_cell<S, GOS, 0>(up, gD[k], left, gI, diag, sig[k], ZERO, gor, ger, h, fD, fI, nib);
            const T H = (FL == BSQ_K3L_INNER) ? h : (inb ? h : NINF);
            cw |= nib << (4 * k);
            gD[k] = fD; gI = fI; diag = up; left = H; Hup[k] = H;
            if (FULLG || FL == BSQ_K3L_TAIL) Hend[k] = inb ? H : Hend[k];
            if (BT == BSQ_B_LOCAL) { if (H > best[k < NB ? k : 0]) { best[k < NB ? k : 0] = H; bestt[k < NB ? k : 0] = t; } }
          }
          const T sentH = inH;         // (x, y0 - 1): next step's diagonal neighbour
          lastH = left;
          mypay = (int)lo | gI;
          if (ti + 1 < NBT) {
            if (lane31) { lastH = rH[ti + 1]; mypay = rP[ti + 1]; }
            inH = __shfl_sync(FULL, lastH, srcl);
            pay = __shfl_sync(FULL, mypay, srcl);
          }
          pst[ti * G] = (PU)cw;        // idle lanes write slots nobody reads (the stride is padded to whole groups)
          bool pput;                   // lane 31 is on a row of the strip: t >= 31 and x <= x_hi
          if (FL == BSQ_K3L_RAMP) pput = put31 && (t >= G - 1);
          else if (FL == BSQ_K3L_INNER) pput = put31;
          else if (FL == BSQ_K3L_TAIL) pput = put31 && (t < nsteps);
          else {
            const bool active = (unsigned)(t - lane) <= (unsigned)(x_hi - x_lo);   // t >= lane and x <= x_hi
            pput = put31 && active;
            if (BT == BSQ_B_OVERLAP) {
              if (active && x < n && km >= 0 && km < K) {   // column m above the last row (:380-390)
                T c = NINF;
#pragma unroll
                for (int k = 0; k < K; k++) if (k == km) c = Hup[k];
                if (c > NINF && bsq_better<T>(c, x - m, 0, bs, bk1, bk2)) { bs = c; bk1 = x - m; bk2 = 0; }
              }
            }
          }
          if (pput) Slab::put_g(pslab + ti * Slab::EPR, left, gI, tag2_out);
          prev_in = sentH;
#pragma unroll
          for (int k = 0; k < K; k++) sig[k] = sign[k];
        }
      };

      for (int t0 = 0; t0 < nsteps; t0 += NBT) {
        {
          // boundary values of rows x_lo + t0 .. x_lo + t0 + NBT - 1: wait for the feeder
          const int want = t0 + NBT;
#ifdef BSQ_K3_PROFILE
          const long long pf_c0 = clock64();
#endif
          while (ctl[1] < want) {
#ifdef BSQ_SIMT_EMU
            emu::yield();
#endif
#ifdef BSQ_K3_PROFILE
            pf_polls++;
#endif
          }
#ifdef BSQ_K3_PROFILE
          pf_wait += clock64() - pf_c0;
          if (t0 == 0) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(pf_ready));
#endif
          rH = s_rH + (t0 & (RING - 1));
          rP = s_rP + (t0 & (RING - 1));
          if (lane31) { lastH = rH[0]; mypay = rP[0]; }
        }
        const bool early = t0 + NBT - 1 < thi;      // every column still above its last step
        const bool begun = allcols && t0 >= tmax_a; // every column past its first
        if (!(lean_strip && t0 >= tlo_lean) || !(early || begun)) group(t0, BsqInt<BSQ_K3L_FULL>{});
        else if (!begun) group(t0, BsqInt<BSQ_K3L_RAMP>{});
        else if (early) group(t0, BsqInt<BSQ_K3L_INNER>{});
        else group(t0, BsqInt<BSQ_K3L_TAIL>{});
        pst += NBT * G;
        pslab += NBT * Slab::EPR;
        __syncwarp();                          // lane 31 has read the group's slots
        if (lane == 0) ctl[2] = t0 / NBT + 1;
      }
      // end cell, _banded_find_optimal (_pw_internals.c:364-414)
      if (BT == BSQ_B_GLOBAL) {
#pragma unroll
        for (int k = 0; k < K; k++) if (k == km && xlast[k] == n) { bs = Hend[k]; bk1 = n - m; bk2 = n; }
      } else if (BT == BSQ_B_OVERLAP) {   // last cell of every diagonal, diagonals ascending, strict >
#pragma unroll
        for (int k = 0; k < K; k++) {
          const T c = Hend[k];
          if (xlast[k] == n && c > NINF && bsq_better<T>(c, n - (y0 + k), 0, bs, bk1, bk2)) { bs = c; bk1 = n - (y0 + k); bk2 = 0; }
        }
      } else {                            // every cell, (d, a) ascending; per column d ascends with x
#pragma unroll
        for (int k = 0; k < NB; k++) {
          const T c = best[k];
          const int bx = x_lo + bestt[k] - lane;
          const int d = bx - (y0 + k);
          if (c > NINF && bsq_better<T>(c, d, bx, bs, bk1, bk2)) { bs = c; bk1 = d; bk2 = bx; }
        }
      }
    }

    bsq_group_best<T, 32>(FULL, bs, bk1, bk2);
    int last = 0;
    if (lane == 0) {
      BsqK3Res r;
      r.score = S::to_f64(bs);
      r.k1 = bk1;
      r.k2 = bk2;
#ifdef BSQ_K3_PROFILE
      r.t_claim = pf_claim; r.t_ready = pf_ready; r.wait_clk = pf_wait; r.polls = pf_polls;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(r.t_end));
      asm volatile("mov.u32 %0, %%smid;" : "=r"(r.sm));
#endif
      A.k3_res[Q.strip_base + j] = r;
      __threadfence();
      last = (atomicAdd(A.k3_done + p, 1) == Q.nstrips - 1);
    }
    last = __shfl_sync(FULL, last, 0);
    if (last) {  // this warp finished the pair's last open strip: pick the end cell
      __threadfence();
      double fs = ScF64::ninf();
      int f1 = BSQ_INT_MAX, f2 = BSQ_INT_MAX;
      for (int i = lane; i < Q.nstrips; i += 32) {
        const BsqK3Res* rp = A.k3_res + Q.strip_base + i;
        const double sc = BSQ_LD_CG(&rp->score);
        const int k1 = BSQ_LD_CG(&rp->k1), k2 = BSQ_LD_CG(&rp->k2);
        if (k1 != BSQ_INT_MAX && (f1 == BSQ_INT_MAX || bsq_better<double>(sc, k1, k2, fs, f1, f2))) { fs = sc; f1 = k1; f2 = k2; }
      }
      bsq_group_best<double, 32>(FULL, fs, f1, f2);
      if (lane == 0) {
        int ex = -1, ey = -1, st = BSQ_OK;
        if (f1 == BSQ_INT_MAX) { st = BSQ_NO_ALIGNMENT; fs = 0.0; }
        else if (BT == BSQ_B_LOCAL) { ex = f2; ey = f2 - f1; }
        else if (BT == BSQ_B_GLOBAL) { ex = n; ey = m; }
        else { ex = n < m + f1 ? n : m + f1; ey = ex - f1; }
        const int oi = P.pad;
        A.out_score[oi] = fs;
        A.out_ex[oi] = ex;
        A.out_ey[oi] = ey;
        A.out_status[oi] = st;
      }
    }
  }
}
