// bsq_post.cuh -- traceback, transcript packing and the table_scores debug
// kernel.  Included by exactly one translation unit (bsq_host.cu).
#pragma once
#include "bsq_kernels.cuh"
#include "bsq_k3.cuh"

// ===========================================================================
// K4: traceback (pw.c:116-151).  One thread per pair walks the 4-bit masks
// from the end cell: the current op is the lowest set bit on arrival by M or at
// the end cell (base = choices[0], _pw_internals.c:235) and the first maximiser
// of _pw_internals.c:266-278 on arrival by a gap.  Ops are written backwards
// into the right end of the pair's slot; a second kernel packs them.
// ===========================================================================
__device__ __forceinline__ uint32_t bsq_lowest(uint32_t mk) { return mk & (0u - mk); }

__device__ __forceinline__ uint32_t bsq_gap_base(uint32_t mk, uint32_t own, int gosign) {
  if (gosign < 0) return (mk & own) ? own : bsq_lowest(mk);
  if (gosign > 0) return (mk & (0xFu ^ own)) ? bsq_lowest(mk & (0xFu ^ own)) : own;
  return bsq_lowest(mk);
}

__device__ __forceinline__ uint32_t bsq_get_mask(const BsqTraceArgs& A, const BsqPair& P, int x, int y) {
  long long idx;
  int sh;
  if (A.mode == BSQ_STD_MODE) {
    const int G = A.G, K = A.K, BW = G * K, WPL = BSQ_K1_WPL(K);
    const int pass = y / BW, r = y - pass * BW, l = r / K, k = r - l * K;
    idx = P.ptr_off + (((long long)pass * (P.n + G) + x + l) * G + l) * WPL + (k >> 3);
    sh = 4 * (k & 7);
  } else {   // [block of TB steps][lane][32-byte unit], nibble (step in block) * C/2 + hc of the unit
    const int G = A.G, C = A.K, WPS = BSQ_K2_WPS(C), TB = BSQ_K2_T(C);
    const int rel = x + y - P.s0, blk = rel / TB, i = rel - blk * TB;
    const int di = (x - y) - P.dmin, l = di / C, c = di - l * C, hc = c >> 1;
    idx = P.ptr_off + ((long long)blk * G + l) * BSQ_K2_UNIT_WORDS + i * WPS + (hc >> 3);
    sh = 4 * (hc & 7);
  }
  return (A.ptr[idx] >> sh) & 0xFu;
}

__global__ void __launch_bounds__(128) bsq_k4_traceback(const BsqTraceArgs A) {
  BSQ_BAIL_IF_BAD(A);
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= A.npairs) return;
  const BsqPair P = A.pairs[p];
  const int oi = P.pad;  // index of this pair in the caller's arrays
  if (A.status[oi] != BSQ_OK) { A.txlen[oi] = 0; A.sx[oi] = -1; A.sy[oi] = -1; return; }
  const int ex = A.ex[oi], ey = A.ey[oi];
  int x = ex, y = ey, len = 0;
  const uint8_t* oseq = A.codes + P.o_off;
  const uint8_t* mseq = A.codes + P.m_off;
  char* wr = A.slots + P.tx_slot + (long long)P.n + P.m;  // one past the slot end
  uint32_t op = bsq_lowest(bsq_get_mask(A, P, x, y));
  while (op != BSQ_MB && op != 0u && len < P.n + P.m) {
    char ch;
    if (op == BSQ_MM) {
      ch = (oseq[x - 1] == mseq[y - 1]) ? 'M' : 'S';  // _pw_internals.c:232
      x--; y--;
      op = bsq_lowest(bsq_get_mask(A, P, x, y));
    } else if (op == BSQ_MD) {
      ch = 'D'; x--;
      op = bsq_gap_base(bsq_get_mask(A, P, x, y), BSQ_MD, A.gosign);
    } else {
      ch = 'I'; y--;
      op = bsq_gap_base(bsq_get_mask(A, P, x, y), BSQ_MI, A.gosign);
    }
    len++;
    wr[-len] = ch;
  }
  A.txlen[oi] = len;
  A.sx[oi] = x;
  A.sy[oi] = y;
  // pw.c:121-138: pos = (ex+ey) - len; pos <= 0 -> PANICK; no ops -> NULL
  A.status[oi] = (ex + ey - len <= 0) ? BSQ_REF_WOULD_ABORT : (len == 0 ? BSQ_EMPTY : BSQ_OK);
}

// ---------------------------------------------------------------------------
// K4w: the same walk for LONG pairs, a group of GW = 8 or 32 lanes per pair (4 or 1 pairs per warp).
// A thread-per-pair walk is a chain of dependent HBM loads (~1 us each); here the
// group stages the next BSQ_TILE_WORDS of pointer rows below the current position
// into shared memory with coalesced 16-byte loads and every lane of the group then
// performs the same walk out of shared memory (lane 0 emits); the rows below the tile are
// prefetched into L2 meanwhile.  Big batches use 8 lanes per pair (more walks in flight per
// issue slot), small ones 32 (groups of one warp diverge and would slow each other down).  The position inside a row is kept
// incrementally (lane, column) so the walk needs no divisions.
//   FAMILY 1: K1 layout [pass][x+lane][lane][word]   row id = pass*(n+G) + x + lane
//   FAMILY 2: K2 layout [s - s0][lane][word]    row id = s - s0
// ---------------------------------------------------------------------------
#define BSQ_TILE_WORDS 512     // per pair
#define BSQ_WALK_WARPS 4
#define BSQ_LET_WIN 128

template <int GOS>
__device__ __forceinline__ uint32_t bsq_gap_base_t(uint32_t mk, uint32_t own) {
  if (GOS < 0) return (mk & own) ? own : bsq_lowest(mk);
  if (GOS > 0) return (mk & (0xFu ^ own)) ? bsq_lowest(mk & (0xFu ^ own)) : own;
  return bsq_lowest(mk);
}

template <int FAMILY, int GOS, int BSQ_WALK_GW>
__global__ void __launch_bounds__(32 * BSQ_WALK_WARPS) bsq_k4_walk(const BsqTraceArgs A) {
  constexpr int BSQ_WALK_PPW = 32 / BSQ_WALK_GW;
  BSQ_BAIL_IF_BAD(A);
  __shared__ uint32_t s_tile[BSQ_WALK_WARPS * BSQ_WALK_PPW][BSQ_TILE_WORDS];
  __shared__ uint8_t s_let[BSQ_WALK_WARPS * BSQ_WALK_PPW][2][BSQ_LET_WIN];  // staged letters: origin, mutant
  const int wib = threadIdx.x >> 5;
  const int lane = threadIdx.x & (BSQ_WALK_GW - 1);              // lane inside my group of GW
  const int grp = (threadIdx.x & 31) / BSQ_WALK_GW;              // my pair's slot inside the warp
  const unsigned gmask = (BSQ_WALK_GW == 32) ? 0xffffffffu : (((1u << BSQ_WALK_GW) - 1u) << (grp * BSQ_WALK_GW));
  const int slot = wib * BSQ_WALK_PPW + grp;
  const int p = blockIdx.x * (BSQ_WALK_WARPS * BSQ_WALK_PPW) + slot;
  if (p >= A.npairs) return;  // the whole group
  const BsqPair P = A.pairs[p];
  const int oi = P.pad;
  if (A.status[oi] != BSQ_OK) {
    if (lane == 0) { A.txlen[oi] = 0; A.sx[oi] = -1; A.sy[oi] = -1; }
    return;
  }
  uint32_t* tile = s_tile[slot];
  const int G = A.G, K = A.K;
  const int WP = (FAMILY == 1) ? BSQ_K1_WPL(K) : (K / 2 + 7) / 8;  // words per lane per row
  const int RB = G * WP;                                        // words per row
  const int R = BSQ_TILE_WORDS / RB;                            // rows per tile
  const int ex = A.ex[oi], ey = A.ey[oi];
  int x = ex, y = ey, len = 0;
  int l, k;            // lane / column (K1) or lane / local diagonal (K2) of (x,y)
  long long rid;       // row id of (x,y)
  if (FAMILY == 1) {
    const int BW = G * K, pass = y / BW, r = y - pass * BW;
    l = r / K; k = r - l * K;
    rid = (long long)pass * (P.n + G) + x + l;
  } else {
    const int di = (x - y) - P.dmin;
    l = di / K; k = di - l * K;
    rid = (long long)(x + y) - P.s0;
  }
  const uint32_t* table = A.ptr + P.ptr_off;
  const uint8_t* oseq = A.codes + P.o_off;
  const uint8_t* mseq = A.codes + P.m_off;
  char* wr = A.slots + P.tx_slot + (long long)P.n + P.m;
  long long tlo = 1;            // first staged row; the tile holds rows [tlo, tlo + rows)
  int roff = -1;                // word offset of row `rid` inside the tile (< 0: not staged)
  int rows = 0;
  int colw, sh;                 // word inside the row and bit position of my nibble
#define BSQ_RECOL() { const int kk_ = (FAMILY == 1) ? k : (k >> 1); colw = l * WP + (kk_ >> 3); sh = 4 * (kk_ & 7); }
  BSQ_RECOL();
  uint32_t arrive = BSQ_MM;     // how we got here: M (or the end cell) -> lowest bit
  const int maxlen = P.n + P.m;
  uint8_t* obuf = s_let[slot][0];
  uint8_t* mbuf = s_let[slot][1];
  int xb = x + 1, yb = y + 1;   // obuf holds o[xb .. xb+WIN), mbuf m[yb .. yb+WIN); empty (forces a load)
  if constexpr (BSQ_WALK_GW == 32) {
  // ---- one pair per warp (small batches): latency bound, so runs of diagonal moves are examined
  // out of preloaded nibbles
  constexpr int BATCH = 8;       // M/S steps examined per round out of preloaded words
  const int mstep = (FAMILY == 1) ? RB : 2 * RB;   // a diagonal move goes back 1 row (K1) / 2 rows (K2)
  bool finished = false;
  while (!finished) {
    if (roff < 0 || roff >= rows * RB) {  // stage the rows at and below `rid`
      long long lo = rid - R + 1;
      if (lo < 0) lo = 0;
      tlo = lo;
      rows = (int)(rid - lo + 1);
      const int nw = rows * RB;            // multiple of 4 words
      const uint4* src = (const uint4*)(table + tlo * RB);
      __syncwarp(gmask);
      for (int i = lane; i < nw / 4; i += BSQ_WALK_GW) ((uint4*)tile)[i] = src[i];
#ifndef BSQ_SIMT_EMU
      {  // the walk goes on below this tile: pull those rows into L2 while this tile is walked
        const long long plo = tlo - R > 0 ? tlo - R : 0;
        const char* pf = (const char*)(table + plo * RB) + 128 * lane;
        if (tlo > 0 && 128 * lane < (int)(tlo - plo) * RB * 4) asm volatile("prefetch.global.L2 [%0];" ::"l"(pf));
        if (BSQ_WALK_GW < 16 && tlo > 0 && 128 * (lane + BSQ_WALK_GW) < (int)(tlo - plo) * RB * 4)
          asm volatile("prefetch.global.L2 [%0];" ::"l"(pf + 128 * BSQ_WALK_GW));
      }
#endif
      __syncwarp(gmask);
      roff = (int)(rid - tlo) * RB;
    }
    if (x - BATCH < xb && xb > 0) {  // the walk only moves towards smaller x and y
      xb = x - BSQ_LET_WIN; if (xb < 0) xb = 0;
      __syncwarp(gmask);
      for (int i = lane; i < BSQ_LET_WIN; i += BSQ_WALK_GW) if (xb + i < P.n) obuf[i] = oseq[xb + i];
      __syncwarp(gmask);
    }
    if (y - BATCH < yb && yb > 0) {
      yb = y - BSQ_LET_WIN; if (yb < 0) yb = 0;
      __syncwarp(gmask);
      for (int i = lane; i < BSQ_LET_WIN; i += BSQ_WALK_GW) if (yb + i < P.m) mbuf[i] = mseq[yb + i];
      __syncwarp(gmask);
    }
    // preload the nibbles a run of up to BATCH diagonal moves would visit: the loads are
    // independent of each other, so the chain per move is a few ALU ops, not a memory access
    int jmax = roff / mstep + 1;
    if (jmax > BATCH) jmax = BATCH;
    if (FAMILY == 1 && jmax > k + 1) jmax = k + 1;   // stay inside my lane's columns
    uint32_t nib[BATCH];
#pragma unroll
    for (int j = 0; j < BATCH; j++) {
      if (j < jmax) {
        int cwj = colw, shj = sh;
        if (FAMILY == 1) { cwj = l * WP + ((k - j) >> 3); shj = 4 * ((k - j) & 7); }
        nib[j] = (tile[roff - j * mstep + cwj] >> shj) & 0xFu;
      } else {
        nib[j] = 0u;
      }
    }
    int taken = 0;           // diagonal moves taken in this round
    uint32_t op = 0u;
#pragma unroll
    for (int j = 0; j < BATCH; j++) {
      if (j < jmax && taken == j) {
        const uint32_t mk = nib[j];
        op = (j == 0 && arrive != BSQ_MM) ? bsq_gap_base_t<GOS>(mk, arrive) : bsq_lowest(mk);
        if (op == BSQ_MM && len < maxlen) {
          const char ch = (obuf[x - 1 - xb] == mbuf[y - 1 - yb]) ? 'M' : 'S';
          x--; y--; len++;
          if (lane == 0) wr[-len] = ch;
          arrive = BSQ_MM;
          taken = j + 1;
        }
      }
    }
    if (taken > 0) {  // account for the diagonal moves
      rid -= (FAMILY == 1) ? taken : 2 * taken;
      roff -= taken * mstep;
      if (FAMILY == 1) {
        k -= taken;
        if (k < 0) {  // into the previous lane: its rows are skewed by one more
          k = K - 1;
          if (--l < 0) { l = G - 1; rid -= (P.n + 1); roff = -1; }  // (pass-1)*(n+G) + x + G-1
          else { rid -= 1; roff -= RB; }
        }
        BSQ_RECOL();
      }
      if (taken == jmax) continue;  // the run may go on: next round
    }
    // here `op` is the op found at the current cell and it is not a diagonal move
    if (taken == 0 && jmax <= 0) continue;   // nothing staged (cannot happen, defensive)
    if (op == BSQ_MB || op == 0u || len >= maxlen || op == BSQ_MM) { finished = true; break; }
    char ch;
    if (op == BSQ_MD) {
      ch = 'D'; x--;
      rid -= 1; roff -= RB;
      if (FAMILY == 2) { if (--k < 0) { k = K - 1; l--; } BSQ_RECOL(); }
    } else {
      ch = 'I'; y--;
      if (FAMILY == 1) {
        if (--k < 0) {
          k = K - 1;
          if (--l < 0) { l = G - 1; rid -= (P.n + 1); roff = -1; }
          else { rid -= 1; roff -= RB; }
        }
      } else {
        rid -= 1; roff -= RB;
        if (++k == K) { k = 0; l++; }
      }
      BSQ_RECOL();
    }
    arrive = op;
    len++;
    if (lane == 0) wr[-len] = ch;
  }
  } else {
  // One move per iteration, written without data-dependent branches: the groups of a warp walk
  // different pairs and would otherwise serialise each other on every gap.
  for (;;) {
    if (roff < 0 || roff >= rows * RB) {  // stage the rows at and below `rid`
      long long lo = rid - R + 1;
      if (lo < 0) lo = 0;
      tlo = lo;
      rows = (int)(rid - lo + 1);
      const int nw = rows * RB;            // multiple of 4 words
      const uint4* src = (const uint4*)(table + tlo * RB);
      __syncwarp(gmask);
      for (int i = lane; i < nw / 4; i += BSQ_WALK_GW) ((uint4*)tile)[i] = src[i];
#ifndef BSQ_SIMT_EMU
      {  // the walk goes on below this tile: pull those rows into L2 while this tile is walked
        const long long plo = tlo - R > 0 ? tlo - R : 0;
        const char* pf = (const char*)(table + plo * RB) + 128 * lane;
        if (tlo > 0 && 128 * lane < (int)(tlo - plo) * RB * 4) asm volatile("prefetch.global.L2 [%0];" ::"l"(pf));
        if (BSQ_WALK_GW < 16 && tlo > 0 && 128 * (lane + BSQ_WALK_GW) < (int)(tlo - plo) * RB * 4)
          asm volatile("prefetch.global.L2 [%0];" ::"l"(pf + 128 * BSQ_WALK_GW));
      }
#endif
      __syncwarp(gmask);
      roff = (int)(rid - tlo) * RB;
    }
    if (x - 1 < xb && xb > 0) {  // the walk only moves towards smaller x and y
      xb = x - BSQ_LET_WIN; if (xb < 0) xb = 0;
      __syncwarp(gmask);
      for (int i = lane; i < BSQ_LET_WIN; i += BSQ_WALK_GW) if (xb + i < P.n) obuf[i] = oseq[xb + i];
      __syncwarp(gmask);
    }
    if (y - 1 < yb && yb > 0) {
      yb = y - BSQ_LET_WIN; if (yb < 0) yb = 0;
      __syncwarp(gmask);
      for (int i = lane; i < BSQ_LET_WIN; i += BSQ_WALK_GW) if (yb + i < P.m) mbuf[i] = mseq[yb + i];
      __syncwarp(gmask);
    }
    const uint32_t mk = (tile[roff + colw] >> sh) & 0xFu;
    const uint32_t op = (arrive == BSQ_MM) ? bsq_lowest(mk) : bsq_gap_base_t<GOS>(mk, arrive);
    if (op == BSQ_MB || op == 0u || len >= maxlen) break;
    const bool isM = (op == BSQ_MM), isD = (op == BSQ_MD), isI = !(isM || isD);
    const int xi = x - 1 - xb, yi = y - 1 - yb;   // clamped: only meaningful for a diagonal move
    const char ms = (obuf[xi < 0 ? 0 : xi] == mbuf[yi < 0 ? 0 : yi]) ? 'M' : 'S';
    const char ch = isM ? ms : (isD ? 'D' : 'I');
    len++;
    if (lane == 0) wr[-len] = ch;
    arrive = op;
    x -= isI ? 0 : 1;
    y -= isD ? 0 : 1;
    if (FAMILY == 2) {   // [s][lane][word]: s = x + y, local diagonal k moves with x - y
      const int dr = isM ? 2 : 1;
      rid -= dr;
      roff -= dr * RB;
      k += isI ? 1 : (isD ? -1 : 0);
      if (k < 0) { k = K - 1; l--; }
      if (k == K) { k = 0; l++; }
    } else {             // [pass][x + lane][lane][word]
      int dr = isI ? 0 : 1;
      if (!isD && --k < 0) {                 // M and I leave the column: maybe the lane, maybe the pass
        k = K - 1;
        if (--l < 0) { l = G - 1; rid -= (P.n + 1); roff = -1 - (1 << 28); }  // (pass-1)*(n+G) + x + G-1: restage
        else dr += 1;                        // the previous lane's rows are skewed by one more
      }
      rid -= dr;
      roff -= dr * RB;
    }
    BSQ_RECOL();
  }
  }
#undef BSQ_RECOL
  if (lane == 0) {
    A.txlen[oi] = len;
    A.sx[oi] = x;
    A.sy[oi] = y;
    A.status[oi] = (ex + ey - len <= 0) ? BSQ_REF_WOULD_ABORT : (len == 0 ? BSQ_EMPTY : BSQ_OK);
  }
}

// ---------------------------------------------------------------------------
// K4t: the walk over K2 tables, one THREAD per pair.
// The table is [block of TB steps][lane][32-byte unit] (bsq_common.h): the cells a path visits
// while it crosses one block of anti-diagonals on one lane's diagonals share ONE sector, so a walk
// costs a DRAM access per ~TB/2 ops (the first layout, a word per lane and step, cost one per op:
// 14 ms for config 3's 32768 pairs whether a thread or 8 lanes walked a pair -- random 32-byte
// reads out of 42 GB, or all 42 GB streamed through shared memory).  Addressing is incremental
// (no divisions in the loop); the unit of the next block down is prefetched when a block is
// entered; with PACKED2 the M / S decision reads the letters out of two cached words of the 2-bit
// packed copy of codes[] (one load per 16 letters and sequence); four ops leave per 32-bit store.
// ---------------------------------------------------------------------------
#define BSQ_WALKT_PF 4

template <int GOS, bool PACKED2>
__global__ void __launch_bounds__(128) bsq_k2_walk_thread(const BsqTraceArgs A) {
  BSQ_BAIL_IF_BAD(A);
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= A.npairs) return;
  const BsqPair P = A.pairs[p];
  const int oi = P.pad;
  if (A.status[oi] != BSQ_OK) { A.txlen[oi] = 0; A.sx[oi] = -1; A.sy[oi] = -1; return; }
  const int G = A.G, C = A.K, WPS = BSQ_K2_WPS(C), TB = BSQ_K2_T(C);
  constexpr int UW = BSQ_K2_UNIT_WORDS;
  const int ex = A.ex[oi], ey = A.ey[oi];
  int x = ex, y = ey, len = 0;
  const int di = (x - y) - P.dmin;
  int l = di / C, c = di - l * C;
  const int rel = x + y - P.s0;
  int blk = rel / TB, i = rel - blk * TB;          // block of steps, step inside it
  const uint32_t* pu = A.ptr + P.ptr_off + ((long long)blk * G + l) * UW;   // my lane's unit of that block
  const long long bstride = (long long)G * UW;
  const uint8_t* oseq = A.codes + P.o_off;
  const uint8_t* mseq = A.codes + P.m_off;
  char* wr_end = A.slots + P.tx_slot + (long long)P.n + P.m;      // one past the slot end
  const int maxlen = P.n + P.m;
  uint32_t arrive = BSQ_MM;     // how we got here: M (or the end cell) -> lowest bit
  uint32_t acc = 0u;            // the last `filled` ops, latest in the low byte
  int filled = 0;
  uint32_t ow = 0u, mw = 0u;    // cached words of the packed letters
  long long owi = -1, mwi = -1;
  // the units of the next BSQ_WALKT_PF blocks down (same lane: a path drifts by a diagonal or two per
  // block) are requested ahead: a block is crossed in a few hundred cycles, DRAM answers in ~a thousand
#ifndef BSQ_SIMT_EMU
#pragma unroll
  for (int k = 1; k <= BSQ_WALKT_PF; k++) if (blk >= k) asm volatile("prefetch.global.L2 [%0];" ::"l"(pu - k * bstride));
#endif
  for (;;) {
    const int hc = c >> 1;
    const uint32_t mk = (pu[i * WPS + (hc >> 3)] >> (4 * (hc & 7))) & 0xFu;
    const uint32_t op = (arrive == BSQ_MM) ? bsq_lowest(mk) : bsq_gap_base_t<GOS>(mk, arrive);
    if (op == BSQ_MB || op == 0u || len >= maxlen) break;
    const bool isM = (op == BSQ_MM), isD = (op == BSQ_MD), isI = !(isM || isD);
    const int xi = x > 0 ? x - 1 : 0, yi = y > 0 ? y - 1 : 0;       // only meaningful for a diagonal move
    bool same;
    if (PACKED2) {
      const long long go_ = P.o_off + xi, gm_ = P.m_off + yi;
      if ((go_ >> 4) != owi) { owi = go_ >> 4; ow = A.codes2[owi]; }
      if ((gm_ >> 4) != mwi) { mwi = gm_ >> 4; mw = A.codes2[mwi]; }
      same = (((ow >> (2 * (int)(go_ & 15))) ^ (mw >> (2 * (int)(gm_ & 15)))) & 3u) == 0u;
    } else {
      same = oseq[xi] == mseq[yi];
    }
    const uint32_t ch = (uint32_t)(unsigned char)(isM ? (same ? 'M' : 'S') : (isD ? 'D' : 'I'));
    acc = (acc << 8) | ch;
    len++;
    filled++;
    {
      char* a = wr_end - len;
      if ((((unsigned long long)a) & 3ull) == 0ull) {     // a word boundary: flush
        if (filled == 4) *(uint32_t*)a = acc;
        else for (int j = 0; j < filled; j++) a[j] = (char)((acc >> (8 * j)) & 0xffu);
        filled = 0;
      }
    }
    arrive = op;
    x -= isI ? 0 : 1;
    y -= isD ? 0 : 1;
    i -= isM ? 2 : 1;                            // s = x + y drops by 2 on a diagonal move
    if (i < 0) {                                 // into the block below
      i += TB;
      blk--;
      pu -= bstride;
#ifndef BSQ_SIMT_EMU
      if (blk >= BSQ_WALKT_PF) asm volatile("prefetch.global.L2 [%0];" ::"l"(pu - BSQ_WALKT_PF * bstride));
#endif
    }
    if (!isM) {                                  // the local diagonal moves with x - y
      c += isI ? 1 : -1;
      if (c < 0) { c = C - 1; l--; pu -= UW; }
      if (c == C) { c = 0; l++; pu += UW; }
    }
  }
  if (filled > 0) {
    char* a = wr_end - len;
    for (int j = 0; j < filled; j++) a[j] = (char)((acc >> (8 * j)) & 0xffu);
  }
  A.txlen[oi] = len;
  A.sx[oi] = x;
  A.sy[oi] = y;
  A.status[oi] = (ex + ey - len <= 0) ? BSQ_REF_WOULD_ABORT : (len == 0 ? BSQ_EMPTY : BSQ_OK);
}

// ---------------------------------------------------------------------------
// K4c: the walk over K2 tables by a GROUP of GW lanes per pair (32: a warp per pair, the default;
// 8: four pairs per warp, measured slower).  The group stages a WINDOW of the table along the path
// -- the units of the path's own lane for the BSQ_WALKC_NBLK blocks of steps at and below the
// current one, 32 bytes per block instead of the 256 - 1024-byte row of blocks -- with independent
// 16-byte loads (GW requests in flight instead of one: a thread-per-pair walker is bound by the
// latency of a DRAM access per block on long pairs, 41 ms on config 4).  Then lane q examines cell
// (x-q, y-q) of the current diagonal (same lane of the table, two steps further down each), a ballot
// finds how far the run of diagonal moves goes, those lanes emit their M / S together, and only the
// op that ends the run is handled serially.  A gap moves the path to the next diagonal; when that
// belongs to the neighbouring lane the window is staged again (every few hundred ops).
// ---------------------------------------------------------------------------
#define BSQ_WALKC_NBLK 32
#define BSQ_WALKC_WARPS 4

template <int GOS, int GW>
__global__ void __launch_bounds__(32 * BSQ_WALKC_WARPS) bsq_k2_walk_coop(const BsqTraceArgs A) {
  constexpr int PPW = 32 / GW;                               // pairs per warp
  constexpr int UW = BSQ_K2_UNIT_WORDS;
  BSQ_BAIL_IF_BAD(A);
  __shared__ uint4 s_win[BSQ_WALKC_WARPS * PPW][BSQ_WALKC_NBLK * 2];   // [block][half unit] of ONE lane
  const int wib = threadIdx.x >> 5;
  const int lane = threadIdx.x & (GW - 1);
  const int grp = (threadIdx.x & 31) / GW;
  const unsigned gmask = (GW == 32) ? 0xffffffffu : (((1u << GW) - 1u) << (grp * GW));
  const unsigned lmask = (GW == 32) ? 0xffffffffu : ((1u << GW) - 1u);
  const int slot = wib * PPW + grp;
  const int p = blockIdx.x * (BSQ_WALKC_WARPS * PPW) + slot;
  if (p >= A.npairs) return;  // the whole group
  const BsqPair P = A.pairs[p];
  const int oi = P.pad;
  if (A.status[oi] != BSQ_OK) {
    if (lane == 0) { A.txlen[oi] = 0; A.sx[oi] = -1; A.sy[oi] = -1; }
    return;
  }
  const int G = A.G, C = A.K, WPS = BSQ_K2_WPS(C), TB = BSQ_K2_T(C);
  const int ex = A.ex[oi], ey = A.ey[oi];
  int x = ex, y = ey, len = 0;
  const int di0 = (x - y) - P.dmin;
  int l = di0 / C, c = di0 - l * C;                          // lane / local diagonal of (x, y)
  int rel = x + y - P.s0;                                    // step of (x, y)
  const uint32_t* table = A.ptr + P.ptr_off;
  const uint8_t* oseq = A.codes + P.o_off;
  const uint8_t* mseq = A.codes + P.m_off;
  char* wr = A.slots + P.tx_slot + (long long)P.n + P.m;
  const int maxlen = P.n + P.m;
  const uint32_t* win = (const uint32_t*)s_win[slot];
  int wlo = 1, whi = 0, wl = -1;                             // window: blocks [wlo, whi] of lane wl
  uint32_t arrive = BSQ_MM;
  for (;;) {
    const int blk = rel / TB;
    if (blk < wlo || blk > whi || l != wl) {                 // stage the window at and below (blk, l)
      whi = blk;
      wlo = blk - BSQ_WALKC_NBLK + 1;
      wlo = wlo < 0 ? 0 : wlo;
      wl = l;
      const int nq = (whi - wlo + 1) * 2;                    // 16-byte quantities to fetch
      __syncwarp(gmask);
      for (int k = lane; k < nq; k += GW)
        s_win[slot][k] = *(const uint4*)(table + ((long long)(wlo + (k >> 1)) * G + wl) * UW + 4 * (k & 1));
      __syncwarp(gmask);
    }
    // lane q looks at (x-q, y-q): the same diagonal (lane l, nibble c/2), step rel - 2q
    const int relq = rel - 2 * lane, blkq = relq >= 0 ? relq / TB : -1, iq = relq - blkq * TB;
    const bool valid = (relq >= 0) && (blkq >= wlo) && (x - lane >= 0) && (y - lane >= 0) && (len + lane < maxlen);
    const int hc = c >> 1;
    uint32_t nib = 0u;
    if (valid) nib = (win[(blkq - wlo) * UW + iq * WPS + (hc >> 3)] >> (4 * (hc & 7))) & 0xFu;
    const uint32_t opq = (lane == 0 && arrive != BSQ_MM) ? bsq_gap_base_t<GOS>(nib, arrive) : bsq_lowest(nib);
    const unsigned vmask = (__ballot_sync(gmask, valid) >> (grp * GW)) & lmask;
    const unsigned mmask = (__ballot_sync(gmask, valid && opq == BSQ_MM) >> (grp * GW)) & lmask;
    const int nv = __popc(vmask);                            // valid cells form a prefix
    const unsigned inv = ~mmask & lmask;
    const int run = inv == 0u ? GW : (__ffs((int)inv) - 1);
    if (nv == 0) break;                                      // len == maxlen (or nothing left)
    if (lane < run) {
      const char ch = (oseq[x - 1 - lane] == mseq[y - 1 - lane]) ? 'M' : 'S';
      wr[-(len + lane + 1)] = ch;
    }
    if (run > 0) { x -= run; y -= run; len += run; rel -= 2 * run; arrive = BSQ_MM; }
    if (run == nv) continue;                                 // the run reached the edge of the window
    const uint32_t op = __shfl_sync(gmask, opq, run, GW);    // the cell that ended the run: not a diagonal move
    if (op == BSQ_MB || op == 0u || op == BSQ_MM || len >= maxlen) break;
    char ch;
    if (op == BSQ_MD) { ch = 'D'; x--; if (--c < 0) { c = C - 1; l--; } }
    else { ch = 'I'; y--; if (++c == C) { c = 0; l++; } }
    rel -= 1;
    arrive = op;
    len++;
    if (lane == 0) wr[-len] = ch;
  }
  if (lane == 0) {
    A.txlen[oi] = len;
    A.sx[oi] = x;
    A.sy[oi] = y;
    A.status[oi] = (ex + ey - len <= 0) ? BSQ_REF_WOULD_ABORT : (len == 0 ? BSQ_EMPTY : BSQ_OK);
  }
}

// ---------------------------------------------------------------------------
// K3 traceback (pw.c:116-151): one warp per pair walks the [strip][t][lane] table.
//
// A path is mostly runs of diagonal moves, so the warp examines 32 cells of the current diagonal
// AT ONCE: lane q looks up the tied-max nibble of (x-q, y-q), a ballot finds how far the run of
// "arrived by M, leaves by M" goes, those lanes emit their M / S in one coalesced store, and only
// the op that ends the run (a gap, or B) is handled serially.  Config 5 (10^6 ops, ~2 10^4 gap
// runs) takes ~5 10^4 rounds instead of 10^6 dependent steps.
//
// The pointer rows come through a small cache of shared-memory tiles filled by the bulk-copy engine
// (cp.async.bulk + mbarrier, one elected lane issues).  A walk crosses a strip in (K + 1) * 32 steps
// and then continues in the strip to its left, ~BW rows further up: on entering a strip the tiles
// of the next two strips are requested around the step where a diagonal path would enter them
// (gaps move that by a few steps; the tile has room), and the steps below the current tile when
// the walk gets near its lower edge.  Config 5 crosses 15 625 strips: without the look-ahead each
// crossing waited for HBM (~1 us: 23 of the 40 ms its traceback took).
// ---------------------------------------------------------------------------
#define BSQ_K3_TILE_BYTES 8192
#define BSQ_K3_NTILES 4
#define BSQ_K3_LET_WIN 1024

#ifndef BSQ_SIMT_EMU
__device__ __forceinline__ uint32_t bsq_smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void bsq_mbar_init(unsigned long long* bar) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bsq_smem_addr(bar)));
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
// one lane: expect `bytes`, start the bulk copy global -> shared, completion counted on the barrier
__device__ __forceinline__ void bsq_bulk_load(void* dst, const void* src, uint32_t bytes, unsigned long long* bar) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bsq_smem_addr(bar)), "r"(bytes) : "memory");
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(bsq_smem_addr(dst)), "l"(src), "r"(bytes), "r"(bsq_smem_addr(bar)) : "memory");
}
__device__ __forceinline__ void bsq_mbar_wait(unsigned long long* bar, uint32_t phase) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t}" ::"r"(bsq_smem_addr(bar)), "r"(phase) : "memory");
}
#endif

template <int K, int GOS>
__global__ void __launch_bounds__(32, 1) bsq_k3_walk(const BsqTraceArgs A) {
  typedef typename BsqK3Unit<K>::type PU;
  constexpr int G = 32, BW = G * K, NT = BSQ_K3_NTILES;
  constexpr int LOGK = (K == 2) ? 1 : (K == 4 ? 2 : 3);
  constexpr int STEP_BYTES = G * (int)sizeof(PU);
  constexpr int TSTEPS = BSQ_K3_TILE_BYTES / STEP_BYTES;
  constexpr int SLACK = TSTEPS / 8;   // steps above the predicted entry step a look-ahead tile covers
  BSQ_BAIL_IF_BAD(A);
  alignas(128) __shared__ unsigned char s_tile[NT][BSQ_K3_TILE_BYTES];
  alignas(16) __shared__ uint8_t s_let[2][BSQ_K3_LET_WIN];
#ifndef BSQ_SIMT_EMU
  alignas(8) __shared__ unsigned long long s_bar[NT];
#endif
  const int lane = threadIdx.x & 31;
  const unsigned FULL = 0xffffffffu;
  const int p = blockIdx.x;
  if (p >= A.npairs) return;
  const BsqPair P = A.pairs[p];
  const int oi = P.pad;
  if (A.status[oi] != BSQ_OK) {
    if (lane == 0) { A.txlen[oi] = 0; A.sx[oi] = -1; A.sy[oi] = -1; }
    return;
  }
#ifndef BSQ_SIMT_EMU
  if (lane == 0) {
#pragma unroll
    for (int i = 0; i < NT; i++) bsq_mbar_init(&s_bar[i]);
  }
  __syncwarp();
#endif
  const int ex = A.ex[oi], ey = A.ey[oi];
  int x = ex, y = ey, len = 0;
  int j = y / BW;                // strip
  int c = y - j * BW;            // column inside the strip: lane = c >> LOGK, nibble = c & (K-1)
  int xlo = j * BW + P.dmin;
  xlo = xlo < 0 ? 0 : xlo;
  const int tstride = P.s0;
  const unsigned char* table = (const unsigned char*)(A.ptr + P.ptr_off);
  const uint8_t* oseq = A.codes + P.o_off;
  const uint8_t* mseq = A.codes + P.m_off;
  char* wr = A.slots + P.tx_slot + (long long)P.n + P.m;
  const int maxlen = P.n + P.m;
  // the tile cache: slot i holds (or is receiving: spend) steps [slo, shi] of strip sj; all uniform
  int sj[NT], slo[NT], shi[NT];
  bool spend[NT];
  uint32_t sphase[NT];
#pragma unroll
  for (int i = 0; i < NT; i++) { sj[i] = -1; slo[i] = 1; shi[i] = 0; spend[i] = false; sphase[i] = 0u; }
  int cur = 0, tlo = 1, thi = 0; // the slot being walked and its step range (empty: look one up)
  int rr = 0;                    // next slot to overwrite (round robin, never `cur`)
  int ahead_for = -1;            // strip whose look-ahead requests have been issued
  int ahead_lo = 0x7fffffff;     // lowest strip a look-ahead tile has been requested for
  uint8_t* obuf = s_let[0];
  uint8_t* mbuf = s_let[1];
  int xb = x + 1, yb = y + 1;    // obuf holds o[xb .. xb+WIN), mbuf m[yb .. yb+WIN); empty
  uint32_t arrive = BSQ_MM;
  auto wait_slot = [&](int b) {  // b: a compile-time index after unrolling
#ifndef BSQ_SIMT_EMU
    if (spend[b]) { bsq_mbar_wait(&s_bar[b], sphase[b]); sphase[b] ^= 1u; }
#endif
    spend[b] = false;
  };
  // steps [lo, hi] of strip js -> slot b; asynchronous when the copy engine can take it
  auto fetch = [&](int b, int js, int lo, int hi) {
    wait_slot(b);                // nothing may still be landing in a slot that is reused
    sj[b] = js; slo[b] = lo; shi[b] = hi;
    const unsigned char* src = table + ((long long)js * tstride + lo) * STEP_BYTES;
    const uint32_t bytes = (uint32_t)(hi - lo + 1) * STEP_BYTES;
#ifndef BSQ_SIMT_EMU
    if ((((unsigned long long)src) & 15ull) == 0ull) {
      if (lane == 0) bsq_bulk_load(s_tile[b], src, bytes, &s_bar[b]);
      spend[b] = true;
      return;
    }
#endif
    __syncwarp();
    for (uint32_t i = lane; i < bytes; i += 32) s_tile[b][i] = src[i];
    __syncwarp();
  };
  auto find = [&](int js, int t) {
    int f = -1;
#pragma unroll
    for (int i = 0; i < NT; i++) if (sj[i] == js && t >= slo[i] && t <= shi[i]) f = i;
    return f;
  };
  // request steps [.., hi] of strip js unless a slot already covers step `t`
  auto want = [&](int js, int t, int hi) {
    if (js < 0 || find(js, t) >= 0) return;
    hi = hi > tstride - 1 ? tstride - 1 : hi;
    int lo = hi - TSTEPS + 1;
    lo = lo < 0 ? 0 : lo;
    if (t < lo || t > hi) return;
    if (rr == cur) rr = (rr + 1) % NT;
    const int b = rr;
    rr = (rr + 1) % NT;
#pragma unroll
    for (int i = 0; i < NT; i++) if (i == b) fetch(i, js, lo, hi);
  };
  // letters [b, b + WIN) of a sequence of `len` into a window, b <= pos - (WIN - 16) chosen so that the
  // source is 16-byte aligned: every lane has all its loads in flight before it stores (a byte loop
  // waits for HBM once per iteration: 14 us per refill, a quarter of config 5's walk)
  auto refill = [&](uint8_t* buf, const uint8_t* seq, int pos, int len) {
    int b = pos - (BSQ_K3_LET_WIN - 16);
    b -= (int)((unsigned long long)(seq + b) & 15ull);
    __syncwarp();
    if (b >= 0) {
      uint4 v[BSQ_K3_LET_WIN / 512];
#pragma unroll
      for (int q = 0; q < BSQ_K3_LET_WIN / 512; q++) {
        const int o = q * 512 + lane * 16;
        v[q] = make_uint4(0u, 0u, 0u, 0u);
        if (b + o < len) v[q] = *(const uint4*)(seq + b + o);   // may read up to 15 bytes past `len`: codes[] is padded
      }
#pragma unroll
      for (int q = 0; q < BSQ_K3_LET_WIN / 512; q++) *(uint4*)(buf + q * 512 + lane * 16) = v[q];
    } else {                     // the window reaches the start of the sequence
      b = 0;
      for (int i = lane; i < BSQ_K3_LET_WIN; i += 32) if (i < len) buf[i] = seq[i];
    }
    __syncwarp();
    return b;
  };
  // The round is written for a lone warp that issues in order and pays 25 - 50 cycles for every
  // branch it resolves: all the rare upkeep (tile lookup, look-ahead requests, letter windows) sits
  // behind ONE test, and moves are applied with selects instead of branches.
  auto cross = [&]() {           // c < 0: into the strip on the left (select form, no branch)
    const bool over = c < 0;
    j -= over ? 1 : 0;
    c += over ? BW : 0;
    int xl = j * BW + P.dmin;
    xl = xl < 0 ? 0 : xl;
    xlo = over ? xl : xlo;
    tlo = over ? 1 : tlo;
    thi = over ? 0 : thi;
  };
  for (;;) {
    const int t = x - xlo + (c >> LOGK);
    const bool upkeep = (t < tlo) | (t > thi) | (ahead_for != j) | ((t - tlo < 40) & (tlo > 0)) |
                        ((x - 32 < xb) & (xb > 0)) | ((y - 32 < yb) & (yb > 0));
    if (upkeep) {
      if (t < tlo || t > thi) {  // the current tile does not hold this step
        __syncwarp();
        int f = find(j, t);
        if (f < 0) {             // nor does any other: fetch it now (no room above: the walk goes up)
          if (rr == cur && thi >= tlo) rr = (rr + 1) % NT;
          f = rr;
          rr = (rr + 1) % NT;
          int lo = t - TSTEPS + 1;
          lo = lo < 0 ? 0 : lo;
#pragma unroll
          for (int i = 0; i < NT; i++) if (i == f) fetch(i, j, lo, t);
        }
#pragma unroll
        for (int i = 0; i < NT; i++) if (i == f) { wait_slot(i); tlo = slo[i]; thi = shi[i]; }
        cur = f;
        __syncwarp();
      }
      if (ahead_for != j) {      // entered a strip: where a diagonal path enters the next two
        ahead_for = j;
        int xe = x - (c + 1);    // row at which column -1 of this strip = column BW-1 of strip j-1 is reached
#pragma unroll
        for (int a = 1; a <= 2; a++) {
          const int js = j - a;
          if (js >= 0 && js < ahead_lo && xe >= 0) {   // (strip j-1 was asked for one crossing ago)
            ahead_lo = js;
            int xl = js * BW + P.dmin;
            xl = xl < 0 ? 0 : xl;
            const int te = xe - xl + (G - 1);
            if (te >= 0) want(js, te < tstride - 1 ? te : tstride - 1, te + SLACK);
          }
          xe -= BW;
        }
      } else if (t - tlo < 40 && tlo > 0) {
        want(j, tlo - 1, tlo - 1);  // the walk is heading for the steps below this tile
      }
      if (x - 32 < xb && xb > 0) xb = refill(obuf, oseq, x, P.n);   // the walk only moves towards smaller x and y
      if (y - 32 < yb && yb > 0) yb = refill(mbuf, mseq, y, P.m);
    }
    // lane q looks at (x-q, y-q): column c-q of this strip, step t-q-(lanes crossed)
    const PU* tile = (const PU*)s_tile[cur];
    const int cq = c - lane, lq = cq >> LOGK;
    const int tq = x - lane - xlo + lq;
    const bool valid = (cq >= 0) && (tq >= tlo) && (len + lane < maxlen);
    uint32_t nib = 0u;
    if (valid) nib = ((uint32_t)tile[(tq - tlo) * G + lq] >> (4 * (cq & (K - 1)))) & 0xFu;
    const uint32_t opq = (lane == 0 && arrive != BSQ_MM) ? bsq_gap_base_t<GOS>(nib, arrive) : bsq_lowest(nib);
    const unsigned vmask = __ballot_sync(FULL, valid);
    const unsigned mmask = __ballot_sync(FULL, valid && opq == BSQ_MM);
    const int nv = __popc(vmask);                       // valid cells form a prefix
    const int run = (~mmask == 0u) ? 32 : (__ffs((int)~mmask) - 1);
    if (lane < run) {
      const char ch = (obuf[x - 1 - lane - xb] == mbuf[y - 1 - lane - yb]) ? 'M' : 'S';
      wr[-(len + lane + 1)] = ch;
    }
    x -= run; y -= run; len += run; c -= run;           // run == 0: nothing moves
    arrive = run > 0 ? BSQ_MM : arrive;
    cross();
    // the cell that ended the run (if it was staged: run < nv) holds an op that is not a diagonal move
    const uint32_t op = __shfl_sync(FULL, opq, run & 31);
    if (run == nv) {
      if (nv == 0) break;                               // len == maxlen
      continue;                                         // the run reached the edge of what was staged (or of the strip)
    }
    if (op == BSQ_MB || op == 0u || op == BSQ_MM || len >= maxlen) break;
    const bool del = (op == BSQ_MD);
    x -= del ? 1 : 0;
    y -= del ? 0 : 1;
    c -= del ? 0 : 1;
    cross();
    arrive = op;
    len++;
    if (lane == 0) wr[-len] = del ? 'D' : 'I';
  }
  // nothing may still be landing in shared memory at exit
#pragma unroll
  for (int i = 0; i < NT; i++) wait_slot(i);
  if (lane == 0) {
    A.txlen[oi] = len;
    A.sx[oi] = x;
    A.sy[oi] = y;
    A.status[oi] = (ex + ey - len <= 0) ? BSQ_REF_WOULD_ABORT : (len == 0 ? BSQ_EMPTY : BSQ_OK);
  }
}

// Input check dptable_init never needed (the reference indexes a double** the caller sized,
// pwlib.h:72-77): every letter the batch references must be < alphabet_len, because the fill
// kernels index a fixed-size shared-memory score table with it.  One warp per pair; raises *bad.
__global__ void __launch_bounds__(128) bsq_check_codes(const uint8_t* codes, const BsqPair* pairs, int npairs,
                                                       int L, int* bad) {
  const int warp = (int)(((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
  const int lane = threadIdx.x & 31;
  if (warp >= npairs) return;
  const BsqPair P = pairs[warp];
  const uint8_t* o = codes + P.o_off;
  const uint8_t* m = codes + P.m_off;
  int worst = 0;
  for (int i = lane; i < P.n; i += 32) worst = max(worst, (int)o[i]);
  for (int i = lane; i < P.m; i += 32) worst = max(worst, (int)m[i]);
  if (worst >= L) *bad = 1;
}

// BSQ_F_CODES_2BIT: four letters per byte over PCIe, one byte per letter for the kernels (the
// sequences are < 2 % of the bytes the fill moves, so the kernels keep their byte loads and the
// link carries a quarter).  Thread t expands packed bytes [4t, 4t+4) into 16 letters.
__global__ void __launch_bounds__(256) bsq_unpack_2bit(const uint8_t* packed, long long nbytes, uint8_t* out) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long b0 = 4 * t;
  if (b0 >= nbytes) return;
  if (b0 + 4 <= nbytes && (((unsigned long long)packed | (unsigned long long)out) & 15ull) == 0ull) {
    const uint32_t w = *(const uint32_t*)(packed + b0);
    uint4 v;
    v.x = (w & 3u) | ((w & 0xCu) << 6) | ((w & 0x30u) << 12) | ((w & 0xC0u) << 18);
    const uint32_t w1 = w >> 8, w2 = w >> 16, w3 = w >> 24;
    v.y = (w1 & 3u) | ((w1 & 0xCu) << 6) | ((w1 & 0x30u) << 12) | ((w1 & 0xC0u) << 18);
    v.z = (w2 & 3u) | ((w2 & 0xCu) << 6) | ((w2 & 0x30u) << 12) | ((w2 & 0xC0u) << 18);
    v.w = (w3 & 3u) | ((w3 & 0xCu) << 6) | ((w3 & 0x30u) << 12) | ((w3 & 0xC0u) << 18);
    *(uint4*)(out + 4 * b0) = v;
  } else {
    for (long long b = b0; b < b0 + 4 && b < nbytes; b++) {
      const uint32_t w = packed[b];
      out[4 * b] = (uint8_t)(w & 3u); out[4 * b + 1] = (uint8_t)((w >> 2) & 3u);
      out[4 * b + 2] = (uint8_t)((w >> 4) & 3u); out[4 * b + 3] = (uint8_t)(w >> 6);
    }
  }
}

// the reverse: letters [16 w, 16 w + 16) -> 32-bit word w of the 2-bit packed copy that the thread
// walker reads (bsq_k2_walk_thread); `letters` bytes in all, words [w_lo, w_hi)
__global__ void __launch_bounds__(256) bsq_pack_2bit(const uint8_t* codes, long long letters, long long w_lo, long long w_hi,
                                                     uint32_t* out) {
  const long long w = w_lo + (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= w_hi) return;
  uint32_t acc = 0u;
  if (16 * w + 16 <= letters && (((unsigned long long)codes) & 15ull) == 0ull) {
    const uint4 v = *(const uint4*)(codes + 16 * w);
    const uint32_t x4[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int k = 0; k < 4; k++) {
      const uint32_t b = x4[k];
      acc |= ((b & 3u) | ((b >> 6) & 0xCu) | ((b >> 12) & 0x30u) | ((b >> 18) & 0xC0u)) << (8 * k);
    }
  } else {
    for (int k = 0; k < 16; k++) {
      const long long q = 16 * w + k;
      if (q < letters) acc |= ((uint32_t)codes[q] & 3u) << (2 * k);
    }
  }
  out[w] = acc;
}

// BSQ_F_TX_2BIT: ops [16 w, 16 w + 16) of the packed byte transcripts -> 32-bit word w, for words
// [w_lo, w_hi); ops at or past `nops` read as 0.  M = 0, S = 1, I = 2, D = 3.
__global__ void __launch_bounds__(256) bsq_tx_to_2bit(const char* ops, long long nops, long long w_lo, long long w_hi,
                                                      uint32_t* out) {
  const long long w = w_lo + (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= w_hi) return;
  uint32_t acc = 0u;
#pragma unroll
  for (int k = 0; k < 16; k++) {
    const long long q = 16 * w + k;
    if (q < nops) {
      const char c = ops[q];
      const uint32_t code = (c == 'M') ? 0u : (c == 'S') ? 1u : (c == 'I') ? 2u : 3u;
      acc |= code << (2 * k);
    }
  }
  out[w] = acc;
}

// exclusive scan of txlen (int32) into int64 offsets, three small kernels
#define BSQ_SCAN_TILE 2048
__global__ void __launch_bounds__(256) bsq_scan_tile_sums(const int* in, int n, long long* tile_sums) {
  __shared__ long long sh[256];
  const int base = blockIdx.x * BSQ_SCAN_TILE;
  long long acc = 0;
  for (int i = threadIdx.x; i < BSQ_SCAN_TILE; i += 256)
    if (base + i < n) acc += in[base + i];
  sh[threadIdx.x] = acc;
  __syncthreads();
  for (int off = 128; off > 0; off >>= 1) {
    if ((int)threadIdx.x < off) sh[threadIdx.x] += sh[threadIdx.x + off];
    __syncthreads();
  }
  if (threadIdx.x == 0) tile_sums[blockIdx.x] = sh[0];
}
// single block: tile_sums -> exclusive prefix (+ *base), writes grand total back to *base_total
__global__ void __launch_bounds__(256) bsq_scan_tiles(long long* tile_sums, int ntiles, const long long* base_in,
                                                      long long* total_out) {
  __shared__ long long carry;
  __shared__ long long sh[256];
  if (threadIdx.x == 0) carry = base_in ? *base_in : 0;
  __syncthreads();
  for (int base = 0; base < ntiles; base += 256) {
    const int i = base + threadIdx.x;
    long long v = (i < ntiles) ? tile_sums[i] : 0;
    sh[threadIdx.x] = v;
    __syncthreads();
    for (int off = 1; off < 256; off <<= 1) {  // Hillis-Steele inclusive
      long long t = ((int)threadIdx.x >= off) ? sh[threadIdx.x - off] : 0;
      __syncthreads();
      sh[threadIdx.x] += t;
      __syncthreads();
    }
    if (i < ntiles) tile_sums[i] = carry + sh[threadIdx.x] - v;
    __syncthreads();
    if (threadIdx.x == 255) carry += sh[255];
    __syncthreads();
  }
  if (threadIdx.x == 0) *total_out = carry;
}
__global__ void __launch_bounds__(256) bsq_scan_apply(const int* in, int n, const long long* tile_sums,
                                                      long long* out) {
  __shared__ long long sh[256];
  __shared__ long long carry;
  const int base = blockIdx.x * BSQ_SCAN_TILE;
  if (threadIdx.x == 0) carry = tile_sums[blockIdx.x];
  __syncthreads();
  for (int j = 0; j < BSQ_SCAN_TILE; j += 256) {
    const int i = base + j + threadIdx.x;
    long long v = (i < n) ? in[i] : 0;
    sh[threadIdx.x] = v;
    __syncthreads();
    for (int off = 1; off < 256; off <<= 1) {
      long long t = ((int)threadIdx.x >= off) ? sh[threadIdx.x - off] : 0;
      __syncthreads();
      sh[threadIdx.x] += t;
      __syncthreads();
    }
    if (i < n) out[i] = carry + sh[threadIdx.x] - v;
    __syncthreads();
    if (threadIdx.x == 255) carry += sh[255];
    __syncthreads();
  }
}

// uniform batches: pair record e from the caller's offset arrays (processing order == caller order)
__global__ void __launch_bounds__(256) bsq_make_pairs_uniform(const long long* o_off, const long long* m_off,
                                                              long long E, const BsqUniform U, BsqPair* out) {
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= E) return;
  BsqPair bp;
  if (o_off) { bp.o_off = o_off[e] + U.off_base; bp.m_off = m_off[e] + U.off_base; }
  else { bp.o_off = U.off_base + e * ((long long)U.n + U.m); bp.m_off = bp.o_off + U.n; }   // pairs back to back in codes[]
  bp.ptr_off = (e % U.ppc) * U.words;
  bp.tx_slot = e * ((long long)U.n + U.m);
  bp.n = U.n; bp.m = U.m;
  bp.dmin = U.dmin; bp.dmax = U.dmax;
  bp.s0 = U.s0;
  bp.pad = (int)e;
  out[e] = bp;
}

// pack transcripts: slot tail -> packed[off]; one warp per pair and segment of BSQ_PACK_SEG ops
// (blockIdx.y = segment: config 5's single 2 10^6-op transcript took one warp 7 ms)
#define BSQ_PACK_SEG 8192
__global__ void __launch_bounds__(256) bsq_pack_transcripts(const BsqPair* pairs, int npairs,
                                                            const int* txlen, const long long* off,
                                                            const char* slots, char* packed) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= npairs) return;
  const int oi = pairs[warp].pad;
  const int len = txlen[oi];
  const long long lo = (long long)blockIdx.y * BSQ_PACK_SEG;
  if (lo >= len) return;
  const int hi = (int)(lo + BSQ_PACK_SEG < len ? lo + BSQ_PACK_SEG : len);
  const char* src = slots + pairs[warp].tx_slot + (long long)pairs[warp].n + pairs[warp].m - len;
  char* dst = packed + off[oi];
  for (int i = (int)lo + lane; i < hi; i += 32) dst[i] = src[i];
}

// ===========================================================================
// K5: Alignment post-processing (pw.py:367-448) of a traced batch, one thread per
// pair over the packed transcripts.  calculate_score adds in the reference's order:
// every run of M/S is summed on its own (Python's sum() starts from 0) and then added
// to the score; a gap run adds go + ge * len.
// ===========================================================================
__global__ void __launch_bounds__(128) bsq_k5_stats(const BsqPair* pairs, int npairs, const uint8_t* codes,
                                                    const long long* txoff, const int* txlen, const char* packed,
                                                    const int* sx, const int* sy, const int* status,
                                                    const double* subst, int L, double go, double ge,
                                                    bsq_aln_stats* out) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= npairs) return;
  const BsqPair P = pairs[e];
  const int oi = P.pad;
  bsq_aln_stats r;
  r.n_match = r.n_subst = r.n_ins = r.n_del = 0;
  r.trunc_lo = -1; r.trunc_hi = -1; r.trunc_start_x = -1; r.trunc_start_y = -1;
  r.rescore = 0.0;
  const int st = status[oi];
  const int len = txlen[oi];
  if ((st == BSQ_OK || st == BSQ_REF_WOULD_ABORT) && len > 0) {
    const char* tx = packed + txoff[oi];
    const uint8_t* o = codes + P.o_off;
    const uint8_t* m = codes + P.m_off;
    int i = sx[oi], j = sy[oi];
    double score = 0.0;
    int t = 0;
    int first_m = -1, last_m = -1, fx = 0, fy = 0;
    while (t < len) {
      const char op = tx[t];
      int run = 1;
      while (t + run < len && tx[t + run] == op) run++;
      if (op == 'M' || op == 'S') {
        double acc = 0.0;
        for (int k = 0; k < run; k++) acc = __dadd_rn(acc, subst[(int)o[i + k] * L + (int)m[j + k]]);
        score = __dadd_rn(score, acc);
        if (op == 'M') {
          r.n_match += run;
          if (first_m < 0) { first_m = t; fx = i; fy = j; }
          last_m = t + run - 1;
        } else {
          r.n_subst += run;
        }
        i += run; j += run;
      } else {
        score = __dadd_rn(score, __dadd_rn(go, __dmul_rn(ge, (double)run)));
        if (op == 'I') { r.n_ins += run; j += run; } else { r.n_del += run; i += run; }
      }
      t += run;
    }
    r.rescore = score;
    if (first_m >= 0 && first_m < last_m) { r.trunc_lo = first_m; r.trunc_hi = last_m; r.trunc_start_x = fx; r.trunc_start_y = fy; }
  }
  out[oi] = r;
}

// ===========================================================================
// table_scores(): H of every cell of one STD problem (pw.py:278-285).  A
// debug/teaching path: one thread block, anti-diagonal sweep over global
// memory, F64 only.
// ===========================================================================
template <int GOS>
__global__ void __launch_bounds__(256) bsq_table_scores_kernel(const uint8_t* o, int n, const uint8_t* t,
                                                               int m, const double* subst, int L, double go,
                                                               double ge, int bmode, double* Hout,
                                                               uint8_t* mask) {
  const double NINF = ScF64::ninf();
  const long long Wd = (long long)m + 1;
  for (int s = 0; s <= n + m; s++) {
    const int xlo = s - m > 0 ? s - m : 0, xhi = s < n ? s : n;
    for (int x = xlo + (int)threadIdx.x; x <= xhi; x += (int)blockDim.x) {
      const int y = s - x;
      const double up = x > 0 ? Hout[(x - 1) * Wd + y] : NINF;
      const double left = y > 0 ? Hout[x * Wd + y - 1] : NINF;
      const double diag = (x > 0 && y > 0) ? Hout[(x - 1) * Wd + y - 1] : NINF;
      const uint32_t upn = x > 0 ? mask[(x - 1) * Wd + y] : 0u;
      const uint32_t leftn = y > 0 ? mask[x * Wd + y - 1] : 0u;
      const double sig = (x > 0 && y > 0) ? subst[(int)o[x - 1] * L + (int)t[y - 1]] : 0.0;
      bool ballow = (x == 0 && y == 0) || bmode == BSQ_B_ALL || (bmode == BSQ_B_EDGE && (x == 0 || y == 0));
      double h;
      uint32_t mk;
      bsq_cell<ScF64, GOS, true>(up, bsq_extends<GOS>(upn, BSQ_MD), left, bsq_extends<GOS>(leftn, BSQ_MI),
                                 diag, sig, ballow ? 0.0 : NINF, go, ge, h, mk);
      Hout[x * Wd + y] = h;
      mask[x * Wd + y] = (uint8_t)mk;
    }
    __syncthreads();
  }
}
