/*
 * lean_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * A CPU restatement of the reference pwlib dynamic-programming recurrence in
 * "compact state" form (one score H + a 4-bit tied-max mask per cell).  It is
 * the checker for the CUDA path: only tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py may load it.  The product
 * (biseqt_b200/) never links, imports or executes anything in oracle/.
 *
 * Parity pin: this file is validated against the *compiled reference itself*
 * (oracle/_ref/pwlib.so, built by oracle/Makefile from the untouched sources
 * under /root/reference/biseqt/pwlib/) by tests/test_oracle_vs_pwlib.py on
 * tens of thousands of random problems over every alignment mode/type, and
 * against the golden vectors of the reference's own docs/tests
 * (tests/golden/, tests/test_oracle_golden.py).
 *
 * What each piece restates (paths relative to /root/reference/biseqt/pwlib):
 *   table geometry / band clipping      _pw_internals.c:8-62   (lo_geometry)
 *   fill order, x-major then y           pw.c:59-61, _pw_internals.c:116-153
 *   candidate order B, D, I, M           pw.c:77-80
 *   keep every candidate == max          pw.c:89-108
 *   B candidate (start allowed?)         _pw_internals.c:161-209 (lo_b_allowed)
 *   M candidate                          _pw_internals.c:217-245
 *   D / I candidates                     _pw_internals.c:247-299 (lo_gap_cand)
 *   end cell, STD                        _pw_internals.c:303-360
 *   end cell, BANDED                     _pw_internals.c:364-414
 *   traceback                            pw.c:116-151
 *
 * A reference cell is a list of choices {op, score, base*}; every stored
 * choice ties for the cell maximum, so the list collapses to (H, mask) with
 * mask bits in the reference's fixed candidate order B<D<I<M and choices[0]
 * being the lowest set bit.  max_new_mins > 0 is not supported (the reference
 * reads uninitialised fields there, _pw_internals.c:236-243).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <limits.h>

#define LO_B 1
#define LO_D 2
#define LO_I 4
#define LO_M 8

/* alnmode (pwlib.h:30-33) */
#define LO_STD_MODE 0
#define LO_BANDED_MODE 1
/* std_alntype (pwlib.h:39-53) */
#define LO_GLOBAL 0
#define LO_LOCAL 1
#define LO_START_ANCHORED 2
#define LO_END_ANCHORED 3
#define LO_OVERLAP 4
#define LO_START_ANCHORED_OVERLAP 5
#define LO_END_ANCHORED_OVERLAP 6
/* banded_alntype (pwlib.h:59-65) */
#define LO_B_GLOBAL 0
#define LO_B_LOCAL 1
#define LO_B_OVERLAP 2

/* result classes; the same numbering as include/bsqgpu.h */
#define LO_OK 0              /* alignment found, transcript non-empty          */
#define LO_NO_ALIGNMENT 1    /* dptable_solve returned {-1,-1} (pw.py:269-271) */
#define LO_EMPTY 2           /* dptable_traceback returned NULL (pw.c:135-138) */
#define LO_INIT_FAIL 3       /* dptable_init returned -1 (pw.py:250-251)       */
#define LO_REF_WOULD_ABORT 4 /* reference hits PANICK/exit(1) (pw.c:132-134)   */

typedef struct {
  int mode, type;
  int n, m;        /* frame lengths: origin_range.j-.i, mutant_range.j-.i */
  int dmin, dmax;  /* after clipping, BANDED only */
  int L;
  const double *subst;
  double go, ge;
  const uint8_t *o, *t; /* letters of the frame */
  int64_t *rowoff;      /* cell index of (x, ylo(x)) */
  int64_t ncells;
} lo_prob;

static inline int lo_ylo(const lo_prob *P, int x) {
  if (P->mode == LO_STD_MODE) return 0;
  return x - P->dmax > 0 ? x - P->dmax : 0; /* _pw_internals.c:146 */
}
static inline int lo_yhi(const lo_prob *P, int x) { /* inclusive */
  if (P->mode == LO_STD_MODE) return P->m;
  return P->m > x - P->dmin ? x - P->dmin : P->m; /* _pw_internals.c:147 */
}
static inline int lo_xlo(const lo_prob *P) {
  if (P->mode == LO_STD_MODE) return 0;
  return P->dmin > 0 ? P->dmin : 0; /* _pw_internals.c:127 */
}
static inline int lo_xhi(const lo_prob *P) { /* inclusive */
  if (P->mode == LO_STD_MODE) return P->n;
  return P->n > P->m + P->dmax ? P->m + P->dmax : P->n; /* _pw_internals.c:128 */
}
static inline int lo_in_table(const lo_prob *P, int x, int y) {
  /* _cellpos_valid (_pw_internals.c:76-85) restated in xy terms */
  if (x < 0 || y < 0 || x > P->n || y > P->m) return 0;
  if (P->mode == LO_BANDED_MODE) {
    int d = x - y;
    if (d < P->dmin || d > P->dmax) return 0;
  }
  return 1;
}
static inline int64_t lo_idx(const lo_prob *P, int x, int y) {
  return P->rowoff[x] + (y - lo_ylo(P, x));
}

/* _alnchoice_B, _pw_internals.c:161-209 */
static inline int lo_b_allowed(const lo_prob *P, int x, int y) {
  if (P->mode == LO_STD_MODE) {
    if (x == 0 && y == 0) return 1;
    if (P->type == LO_LOCAL || P->type == LO_END_ANCHORED) return 1;
    if ((P->type == LO_OVERLAP || P->type == LO_END_ANCHORED_OVERLAP) &&
        (x == 0 || y == 0))
      return 1;
    return 0;
  }
  if (P->type == LO_B_GLOBAL && x == 0 && y == 0) return 1;
  if (P->type == LO_B_OVERLAP && (x == 0 || y == 0)) return 1;
  if (P->type == LO_B_LOCAL) return 1;
  return 0;
}

/* _alnchoice_ID, _pw_internals.c:266-281: first strict max over the tied
 * choices of the source cell, starting from -INT_MAX.  `own` is LO_D or LO_I.
 * Returns the candidate score and which source choice it chains to. */
static inline double lo_gap_cand(double Hsrc, int masksrc, int own, double go,
                                 double ge, int *base_bit) {
  double max_score = -(double)INT_MAX;
  int base = 0, first = 0;
  for (int bit = 1; bit <= 8; bit <<= 1) {
    if (!(masksrc & bit)) continue;
    if (!first) { first = bit; base = bit; } /* base_idx = 0 initially */
    double s = Hsrc + ge;
    if (bit != own) s += go;
    if (s > max_score) { max_score = s; base = bit; }
  }
  *base_bit = base;
  return max_score;
}

static inline int lo_lowest(int mask) { return mask & -mask; }

/*
 * Solve + trace back one problem.
 *   origin/mutant : letter indices (uint8) of the FRAME (caller slices)
 *   transcript    : caller buffer of >= n+m+1 bytes (NUL terminated on return)
 *   hout          : optional (n+1)*(m+1) doubles, STD mode only (table_scores)
 * Returns the status class; on LO_OK / LO_REF_WOULD_ABORT all outputs are set
 * (start_x/start_y relative to the frame).
 */
int lean_align(int mode, int type, int L, const double *subst, double go,
               double ge, const uint8_t *origin, int n, const uint8_t *mutant,
               int m, int dmin, int dmax, double *score, int *end_x, int *end_y,
               int *start_x, int *start_y, char *transcript, int *tx_len,
               int64_t *ncells_out, double *hout) {
  lo_prob P;
  memset(&P, 0, sizeof P);
  P.mode = mode; P.type = type; P.n = n; P.m = m; P.L = L;
  P.subst = subst; P.go = go; P.ge = ge; P.o = origin; P.t = mutant;
  *score = 0; *end_x = *end_y = -1; *start_x = *start_y = -1; *tx_len = 0;
  if (transcript) transcript[0] = 0;
  if (ncells_out) *ncells_out = 0;

  if (mode == LO_BANDED_MODE) { /* _banded_table_init_dims, :22-62 */
    int dend = n - m;
    if (dmax > n) dmax = n;
    if (dmin < -m) dmin = -m;
    if (type == LO_B_GLOBAL &&
        (dend > dmax || dend < dmin || (int64_t)dmax * dmin > 0))
      return LO_INIT_FAIL;
    if (1 + dmax - dmin < 0) return LO_INIT_FAIL;
    P.dmin = dmin; P.dmax = dmax;
    if (1 + dmax - dmin == 0) return LO_NO_ALIGNMENT; /* zero rows: nothing to scan */
  }

  int xlo = lo_xlo(&P), xhi = lo_xhi(&P);
  P.rowoff = (int64_t *)malloc((size_t)(n + 2) * sizeof(int64_t));
  int64_t acc = 0;
  for (int x = 0; x <= n; x++) {
    P.rowoff[x] = acc;
    if (x >= xlo && x <= xhi) {
      int a = lo_ylo(&P, x), b = lo_yhi(&P, x);
      if (b >= a) acc += (b - a + 1);
    }
  }
  P.ncells = acc;
  if (ncells_out) *ncells_out = acc;
  uint8_t *mask = (uint8_t *)calloc((size_t)(acc > 0 ? acc : 1), 1);
  /* two rolling rows of H, indexed by absolute y */
  double *Hp = (double *)malloc((size_t)(m + 2) * sizeof(double));
  double *Hc = (double *)malloc((size_t)(m + 2) * sizeof(double));

  /* running end-cell search state (same scan order as the reference, which
   * scans after the fill; the orders are restated case by case below) */
  double best = 0; int bx = -1, by = -1;
  int track_all_std = (mode == LO_STD_MODE &&
                       (type == LO_LOCAL || type == LO_START_ANCHORED));
  /* LOCAL scan starts from H(0,0) (_pw_internals.c:342) which is always 0 */
  if (track_all_std) best = 0.0;

  /* banded scans are diagonal-major; keep per-diagonal records */
  int W = (mode == LO_BANDED_MODE) ? (P.dmax - P.dmin + 1) : 0;
  double *dbest = NULL, *dlast = NULL; int *dbest_a = NULL, *dlast_a = NULL;
  if (mode == LO_BANDED_MODE) {
    dbest = (double *)malloc(sizeof(double) * W);
    dlast = (double *)malloc(sizeof(double) * W);
    dbest_a = (int *)malloc(sizeof(int) * W);
    dlast_a = (int *)malloc(sizeof(int) * W);
    for (int i = 0; i < W; i++) { dbest_a[i] = -1; dlast_a[i] = -1; dbest[i] = dlast[i] = 0; }
  }
  /* STD overlap scan: right column top->bottom (rows < n), then bottom row */
  double *colH = NULL;
  if (mode == LO_STD_MODE && (type == LO_OVERLAP || type == LO_START_ANCHORED_OVERLAP))
    colH = (double *)malloc(sizeof(double) * (size_t)(n + 1));

  for (int x = xlo; x <= xhi; x++) { /* pw.c:59 */
    int ylo = lo_ylo(&P, x), yhi = lo_yhi(&P, x);
    for (int y = ylo; y <= yhi; y++) { /* pw.c:61 */
      double cand[4]; int cbit[4]; int nc = 0;
      if (lo_b_allowed(&P, x, y)) { cand[nc] = 0.0; cbit[nc] = LO_B; nc++; }
      if (lo_in_table(&P, x - 1, y)) { /* D, from (x-1,y) */
        int ms = mask[lo_idx(&P, x - 1, y)];
        if (ms) { int bb; cand[nc] = lo_gap_cand(Hp[y], ms, LO_D, go, ge, &bb); cbit[nc] = LO_D; nc++; }
      }
      if (lo_in_table(&P, x, y - 1)) { /* I, from (x,y-1) */
        int ms = mask[lo_idx(&P, x, y - 1)];
        if (ms) { int bb; cand[nc] = lo_gap_cand(Hc[y - 1], ms, LO_I, go, ge, &bb); cbit[nc] = LO_I; nc++; }
      }
      if (lo_in_table(&P, x - 1, y - 1)) { /* M/S, from (x-1,y-1), :217-245 */
        int ms = mask[lo_idx(&P, x - 1, y - 1)];
        if (ms) {
          cand[nc] = Hp[y - 1] + subst[(int)origin[x - 1] * L + (int)mutant[y - 1]];
          cbit[nc] = LO_M; nc++;
        }
      }
      int mk = 0; double H = 0;
      if (nc > 0) { /* pw.c:89-103 */
        double mx = cand[0];
        for (int k = 0; k < nc; k++) {
          if (cand[k] == mx) mk |= cbit[k];
          else if (cand[k] > mx) { mx = cand[k]; mk = cbit[k]; }
        }
        H = mx;
      }
      mask[lo_idx(&P, x, y)] = (uint8_t)mk;
      Hc[y] = H;
      if (hout && mode == LO_STD_MODE) hout[(int64_t)x * (m + 1) + y] = H;
      if (mode == LO_BANDED_MODE && !mk) dlast_a[(x - y) - P.dmin] = -1;
      if (!mk) continue;
      if (track_all_std) { /* :343-353 row-major strict > */
        if (H > best) { best = H; bx = x; by = y; }
      }
      if (colH && y == m) colH[x] = H;
      if (mode == LO_BANDED_MODE) {
        int di = (x - y) - P.dmin, a = x < y ? x : y;
        /* within a diagonal `a` ascends with x, so "first strict max along the
         * diagonal" here + "first strict max over diagonals" below is the
         * d-major, a-minor scan of :397-406 */
        if (dbest_a[di] < 0 || H > dbest[di]) { dbest[di] = H; dbest_a[di] = a; }
        dlast[di] = H; dlast_a[di] = a; /* geometric last cell j=row_lens-1, :383 */
      }
    }
    double *tmp = Hp; Hp = Hc; Hc = tmp; /* Hp now holds row x */
  }
  /* after the loop Hp holds the last filled row (x = xhi) */

  int ex = -1, ey = -1; double escore = 0;
  if (mode == LO_STD_MODE) {
    if (type == LO_GLOBAL || type == LO_END_ANCHORED || type == LO_END_ANCHORED_OVERLAP) {
      if (mask[lo_idx(&P, n, m)]) { ex = n; ey = m; escore = Hp[m]; }
    } else if (type == LO_OVERLAP || type == LO_START_ANCHORED_OVERLAP) {
      double mx = -(double)INT_MAX; /* :320-337 */
      for (int i = 0; i < n; i++)
        if (mask[lo_idx(&P, i, m)] && colH[i] > mx) { mx = colH[i]; ex = i; ey = m; }
      for (int j = 0; j <= m; j++)
        if (mask[lo_idx(&P, n, j)] && Hp[j] > mx) { mx = Hp[j]; ex = n; ey = j; }
      escore = mx;
    } else { ex = bx; ey = by; escore = best; }
  } else {
    if (type == LO_B_GLOBAL) {
      if (lo_in_table(&P, n, m) && mask[lo_idx(&P, n, m)]) { ex = n; ey = m; escore = Hp[m]; }
    } else if (type == LO_B_OVERLAP) { /* :380-395 */
      double mx = -(double)INT_MAX;
      for (int i = 0; i < W; i++) {
        if (dlast_a[i] < 0) continue;
        /* reference looks at the geometric last cell of the diagonal; every
         * in-band cell is filled so dlast is that cell */
        if (dlast[i] > mx) {
          mx = dlast[i]; int d = i + P.dmin, a = dlast_a[i];
          ex = a + (d > 0 ? d : 0); ey = a - (d > 0 ? 0 : d);
        }
      }
      escore = mx;
    } else { /* B_LOCAL :396-412 */
      double mx = -(double)INT_MAX;
      for (int i = 0; i < W; i++) {
        if (dbest_a[i] < 0) continue;
        if (dbest[i] > mx) {
          mx = dbest[i]; int d = i + P.dmin, a = dbest_a[i];
          ex = a + (d > 0 ? d : 0); ey = a - (d > 0 ? 0 : d);
        }
      }
      escore = mx;
    }
  }

  int status;
  if (ex < 0 || ey < 0) {
    status = LO_NO_ALIGNMENT;
  } else {
    *end_x = ex; *end_y = ey; *score = escore;
    /* traceback, pw.c:116-151; written backwards then reversed in place */
    int x = ex, y = ey, len = 0;
    int op = lo_lowest(mask[lo_idx(&P, x, y)]);
    while (op != LO_B) {
      int src_mask, bb = 0;
      if (op == LO_M) {
        transcript[len++] = (origin[x - 1] == mutant[y - 1]) ? 'M' : 'S';
        x--; y--;
        op = lo_lowest(mask[lo_idx(&P, x, y)]); /* base = choices[0], :235 */
      } else if (op == LO_D) {
        transcript[len++] = 'D'; x--;
        src_mask = mask[lo_idx(&P, x, y)];
        (void)lo_gap_cand(0.0, src_mask, LO_D, go, ge, &bb); /* tie-break only depends on go */
        op = bb;
      } else { /* LO_I */
        transcript[len++] = 'I'; y--;
        src_mask = mask[lo_idx(&P, x, y)];
        (void)lo_gap_cand(0.0, src_mask, LO_I, go, ge, &bb);
        op = bb;
      }
    }
    for (int i = 0; i < len / 2; i++) {
      char c = transcript[i]; transcript[i] = transcript[len - 1 - i]; transcript[len - 1 - i] = c;
    }
    transcript[len] = 0;
    *tx_len = len; *start_x = x; *start_y = y;
    /* pw.c:121-138: pos = (ex+ey) - len; pos <= 0 -> PANICK, else len==0 -> NULL */
    if (ex + ey - len <= 0) status = LO_REF_WOULD_ABORT;
    else if (len == 0) status = LO_EMPTY;
    else status = LO_OK;
  }

  free(P.rowoff); free(mask); free(Hp); free(Hc);
  free(dbest); free(dlast); free(dbest_a); free(dlast_a); free(colH);
  return status;
}

/* number of cells the reference's fill loop visits (pw.c:59-61) */
int64_t lean_count_cells(int mode, int n, int m, int dmin, int dmax) {
  if (mode == LO_STD_MODE) return (int64_t)(n + 1) * (m + 1);
  if (dmax > n) dmax = n;
  if (dmin < -m) dmin = -m;
  if (dmax < dmin) return 0;
  int64_t acc = 0;
  int xlo = dmin > 0 ? dmin : 0, xhi = n > m + dmax ? m + dmax : n;
  for (int x = xlo; x <= xhi; x++) {
    int a = x - dmax > 0 ? x - dmax : 0, b = m > x - dmin ? x - dmin : m;
    if (b >= a) acc += b - a + 1;
  }
  return acc;
}

/* Batch driver used as the "port" CPU baseline and by tests: same packed
 * layout as the C ABI in include/bsqgpu.h. Single-threaded on purpose (callers
 * fan out over processes). tx_buf receives NUL-free transcripts at tx_off[i]
 * (slot capacity o_len+m_len). */
int lean_align_batch(int mode, int type, int L, const double *subst, double go,
                     double ge, int64_t npairs, const uint8_t *codes,
                     const int64_t *o_off, const int32_t *o_len,
                     const int64_t *m_off, const int32_t *m_len,
                     const int32_t *dmin, const int32_t *dmax, double *score,
                     int32_t *end_x, int32_t *end_y, int32_t *start_x,
                     int32_t *start_y, int32_t *status, int32_t *tx_len,
                     const int64_t *tx_off, char *tx_buf) {
  for (int64_t p = 0; p < npairs; p++) {
    int n = o_len[p], m = m_len[p];
    char *tmp = (char *)malloc((size_t)n + m + 2);
    int ex, ey, sx, sy, tl;
    int st = lean_align(mode, type, L, subst, go, ge, codes + o_off[p], n,
                        codes + m_off[p], m, dmin ? dmin[p] : 0,
                        dmax ? dmax[p] : 0, &score[p], &ex, &ey, &sx, &sy, tmp,
                        &tl, NULL, NULL);
    end_x[p] = ex; end_y[p] = ey; start_x[p] = sx; start_y[p] = sy;
    status[p] = st; tx_len[p] = tl;
    if (tx_buf && tl > 0) memcpy(tx_buf + tx_off[p], tmp, (size_t)tl);
    free(tmp);
  }
  return 0;
}
