/*
 * bsqgpu.h -- C ABI of libbsqgpu.so, the B200 (sm_100a) replacement for the
 * DP hot path of biseqt's pwlib.
 *
 * Every entry point is `extern "C"`, takes plain pointers and sizes, never
 * throws, never prints and never calls exit().  The reference's own FFI
 * surface for this path is four functions over one heap-allocated table
 * (/root/reference/biseqt/pwlib/pwlib.h:204-242):
 *
 *     int        dptable_init(dptable*);                 pwlib.h:218
 *     intpair    dptable_solve(dptable*);                pwlib.h:242
 *     alignment* dptable_traceback(dptable*, intpair);   pwlib.h:228
 *     void       dptable_free(dptable*);                 pwlib.h:225
 *
 * which biseqt/pw.py binds with cffi (pw.py:45-69) one pair at a time.  The
 * struct-level ABI cannot be kept (pw.py:272,282 read cells[i][j].choices[0]
 * straight out of the C heap), so the drop-in boundary is batch-oriented and
 * each function below names the reference call(s) it replaces.  A batch of 1
 * is what biseqt_b200.pw.Aligner issues.
 *
 * Letters are alphabet indices, one uint8 per letter (the reference passes
 * int* built from Sequence.contents, pw.py:211-212).  All lengths/ranges are
 * the alignment FRAME: callers apply origin_range/mutant_range by pointing
 * o_off/m_off at the frame start (pwlib does exactly that, _pw_internals.c:
 * 226-229) and add the frame offset to the reported start (pw.c:146-147).
 */
#ifndef BSQGPU_H
#define BSQGPU_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* alnmode, pwlib.h:30-33 */
#define BSQ_STD_MODE 0
#define BSQ_BANDED_MODE 1
/* std_alntype, pwlib.h:39-53 */
#define BSQ_GLOBAL 0
#define BSQ_LOCAL 1
#define BSQ_START_ANCHORED 2
#define BSQ_END_ANCHORED 3
#define BSQ_OVERLAP 4
#define BSQ_START_ANCHORED_OVERLAP 5
#define BSQ_END_ANCHORED_OVERLAP 6
/* banded_alntype, pwlib.h:59-65 */
#define BSQ_B_GLOBAL 0
#define BSQ_B_LOCAL 1
#define BSQ_B_OVERLAP 2

/* per-pair status (what the reference would have done with this pair) */
#define BSQ_OK 0              /* alignment found, transcript non-empty              */
#define BSQ_NO_ALIGNMENT 1    /* dptable_solve -> {-1,-1}; Aligner.solve() is None  */
#define BSQ_EMPTY 2           /* dptable_traceback -> NULL (empty transcript)       */
#define BSQ_INIT_FAIL 3       /* dptable_init -> -1; Aligner.__enter__ raises       */
#define BSQ_REF_WOULD_ABORT 4 /* reference PANICKs/exit(1) in traceback (pw.c:132); *
                               * score/end/start/transcript here are still valid    */
#define BSQ_UNSUPPORTED 5     /* outside what this build handles (see last_error)   */

/* call-level return codes */
#define BSQ_SUCCESS 0
#define BSQ_ERR_ARG -1        /* bad argument (message in bsq_last_error)           */
#define BSQ_ERR_CUDA -2       /* CUDA runtime / no device                           */
#define BSQ_ERR_NOMEM -3      /* device or host memory                              */
#define BSQ_ERR_CAPACITY -4   /* tx_buf too small; needed size in *tx_total         */

/* arithmetic the fill kernels ran in */
#define BSQ_DTYPE_F64 0       /* general path: IEEE double, same association as pwlib */
#define BSQ_DTYPE_I32 1       /* exact fast path: every score an integer, no overflow */
#define BSQ_DTYPE_I16X2 2     /* reported only: the I32 path's short-read form, two problems per  *
                               * register in packed 16-bit arithmetic (same exact results)        */
#define BSQ_DTYPE_AUTO -1

typedef struct bsq_ctx bsq_ctx;     /* one per GPU; single-owner                    */
typedef struct bsq_batch bsq_batch; /* a solved batch whose tables are still in HBM */

/* alnscores + alnprob minus the per-pair parts (pwlib.h:71-77,113-123) */
typedef struct {
  int32_t mode;          /* BSQ_STD_MODE | BSQ_BANDED_MODE            (alnprob.mode)   */
  int32_t type;          /* std_alntype | banded_alntype              (*_alnparams.type) */
  int32_t alphabet_len;  /* L; letters are 0..L-1                                      */
  int32_t max_new_mins;  /* must be <= 0 (the reference's >0 path is undefined)        */
  double go;             /* alnscores.gap_open_score                                   */
  double ge;             /* alnscores.gap_extend_score                                 */
  const double* subst;   /* L*L row-major, subst[o*L+m]            (alnscores.subst_scores) */
  int32_t dtype;         /* BSQ_DTYPE_AUTO (default), or force F64 / I32               */
  int32_t flags;         /* BSQ_F_* below, 0 = one byte per letter in, one byte per op out */
} bsq_params;

/* bsq_params.flags: compact forms of the two big arrays that cross PCIe (alphabets of <= 4
 * letters; the reference passes int* at 4 bytes per letter, pwlib.h:85-86).
 *   BSQ_F_CODES_2BIT  codes[] holds four letters per byte: letter j sits in bits 2*(j&3) .. +1 of
 *                     byte j>>2.  o_off / m_off / ncodes count LETTERS, as always.
 *   BSQ_F_TX_2BIT     tx_buf[] holds four ops per byte (M=0, S=1, I=2, D=3): op k of the batch sits
 *                     in bits 2*(k&3) .. +1 of byte k>>2.  tx_off[] / *tx_total count OPS, as always
 *                     (transcript i = ops tx_off[i] .. tx_off[i+1]); tx_cap stays in BYTES and
 *                     must hold ceil(total ops / 16) * 4 of them. */
#define BSQ_F_CODES_2BIT 1
#define BSQ_F_TX_2BIT 2

/* kernel-level timing of the last bsq_align_batch / bsq_solve_batch call,
 * measured with cudaEvents on the library's own streams */
typedef struct {
  double fill_ms;        /* sum of DP fill kernel durations                      */
  double traceback_ms;   /* sum of traceback + compaction kernel durations       */
  double total_ms;       /* first H2D enqueue to last D2H complete (device time) */
  int64_t cells;         /* DP cells the reference would have filled (pw.c:59-61) */
  int64_t fill_launches; /* kernels launched inside the call (all kinds)         */
  int64_t h2d_bytes;
  int64_t d2h_bytes;
  int64_t ptr_bytes;     /* bytes of traceback-pointer tables written to HBM     */
  int32_t dtype;         /* BSQ_DTYPE_* actually used                            */
  int32_t chunks;
  double plan_ms;        /* host: classification + ordering + chunking (wall clock) */
  double call_ms;        /* host: whole call, wall clock                            */
} bsq_timing;

int bsq_device_count(void);
/* device: CUDA ordinal.  Returns NULL on failure (bsq_last_error(NULL)). */
bsq_ctx* bsq_create(int device);
void bsq_destroy(bsq_ctx* ctx);
const char* bsq_last_error(const bsq_ctx* ctx);
const char* bsq_version(void);

/* Test / tuning hooks are BSQ_* environment variables (tools/README.md) read ONCE, in
 * bsq_create; nothing on the call path reads the environment.  bsq_reload_hooks reads them
 * again for this context (tests flip a hook between two calls on one context). */
void bsq_reload_hooks(bsq_ctx* ctx);

/* pinned host memory so that H2D/D2H run at link speed without a bounce copy */
void* bsq_host_alloc(int64_t bytes);
void bsq_host_free(void* p);

/*
 * dptable_init + dptable_solve + dptable_traceback + dptable_free for a whole
 * batch, pipelined over chunks that fit the pointer-table arena.
 *
 *  in : codes[]               all letters, uint8 (or 2-bit packed: BSQ_F_CODES_2BIT)
 *       o_off/o_len, m_off/m_len   frame start + length of origin / mutant.  o_off and m_off
 *                             may both be NULL: the pairs then lie back to back in codes[],
 *                             origin_0 | mutant_0 | origin_1 | ... (nothing to upload for them)
 *       dmin/dmax             BANDED only (NULL in STD); banded_alnparams.dmin/dmax,
 *                             clipped to [-m, n] like _pw_internals.c:29-36
 *  out: score[npairs]         choices[0].score of the end cell (pw.py:272)
 *       end_x/end_y           end cell, frame-relative xy (dptable_solve's intpair,
 *                             mapped through _xy_from_cellpos for BANDED)
 *       start_x/start_y       alignment start, frame-relative (alignment.origin_idx /
 *                             mutant_idx minus the frame offset, pw.c:146-147)
 *       status[npairs]        BSQ_OK ... BSQ_UNSUPPORTED
 *       tx_off[npairs+1]      transcript i is tx_buf[tx_off[i] .. tx_off[i+1])
 *       tx_buf[tx_cap]        'M','S','I','D' bytes, packed, no terminators
 *       *tx_total             total transcript bytes (also set on BSQ_ERR_CAPACITY)
 *  Any of end_*, start_*, tx_* may be NULL (tx_off/tx_buf NULL = scores only,
 *  no traceback work is done).
 */
int bsq_align_batch(bsq_ctx* ctx, const bsq_params* params, int64_t npairs,
                    const uint8_t* codes, int64_t ncodes,
                    const int64_t* o_off, const int32_t* o_len,
                    const int64_t* m_off, const int32_t* m_len,
                    const int32_t* dmin, const int32_t* dmax,
                    double* score, int32_t* end_x, int32_t* end_y,
                    int32_t* start_x, int32_t* start_y, int32_t* status,
                    int64_t* tx_off, char* tx_buf, int64_t tx_cap,
                    int64_t* tx_total);

/*
 * Two-phase form mirroring Aligner.solve() / Aligner.traceback():
 * bsq_solve_batch = dptable_init + dptable_solve and keeps the pointer tables
 * resident (the whole batch must fit the arena, else BSQ_ERR_NOMEM);
 * bsq_traceback_batch = dptable_traceback; bsq_release = dptable_free.
 */
int bsq_solve_batch(bsq_ctx* ctx, const bsq_params* params, int64_t npairs,
                    const uint8_t* codes, int64_t ncodes,
                    const int64_t* o_off, const int32_t* o_len,
                    const int64_t* m_off, const int32_t* m_len,
                    const int32_t* dmin, const int32_t* dmax,
                    double* score, int32_t* end_x, int32_t* end_y,
                    int32_t* status, bsq_batch** batch);
int bsq_traceback_batch(bsq_batch* batch, int32_t* start_x, int32_t* start_y,
                        int32_t* status, int64_t* tx_off, char* tx_buf,
                        int64_t tx_cap, int64_t* tx_total);
void bsq_release(bsq_batch* batch);

/*
 * Device-resident form of the same path, for pipelines and for measurement:
 * bsq_batch_upload plans the batch and copies the inputs to HBM once;
 * bsq_batch_run executes fill (+ traceback + transcript packing when
 * with_traceback != 0) with every input and output resident in HBM and returns
 * when the device is done (bsq_get_timing then holds the CUDA-event times);
 * bsq_batch_fetch copies the results to host buffers (any pointer may be NULL).
 * keep_tables != 0 keeps all pointer tables resident (needed for a later
 * bsq_traceback_batch); 0 lets the scheduler reuse one arena chunk by chunk.
 */
int bsq_batch_upload(bsq_ctx* ctx, const bsq_params* params, int64_t npairs,
                     const uint8_t* codes, int64_t ncodes,
                     const int64_t* o_off, const int32_t* o_len,
                     const int64_t* m_off, const int32_t* m_len,
                     const int32_t* dmin, const int32_t* dmax,
                     int32_t keep_tables, bsq_batch** batch);
int bsq_batch_run(bsq_batch* batch, int32_t with_traceback);
int bsq_batch_fetch(bsq_batch* batch, double* score, int32_t* end_x, int32_t* end_y,
                    int32_t* start_x, int32_t* start_y, int32_t* status,
                    int64_t* tx_off, char* tx_buf, int64_t tx_cap, int64_t* tx_total);

/*
 * Alignment post-processing for every pair of a traced, resident batch, on the device
 * (the step right after the DP in every caller, e.g. experiments/blot_stats.py:460-470):
 *   n_match/n_subst/n_ins/n_del   op counts; Alignment.projected_len (pw.py:367-389) is
 *                                 n_match+n_subst+n_del on the origin, +n_ins on the mutant
 *   rescore                       Alignment.calculate_score (pw.py:391-428) of the transcript with
 *                                 the batch's own scores, summed in the reference's order
 *   trunc_*                       Alignment.truncate_to_match (pw.py:430-448): transcript range
 *                                 [trunc_lo, trunc_hi] and the new frame-relative start, or
 *                                 trunc_lo = -1 where the reference returns None
 * Pairs without a transcript (status != BSQ_OK / BSQ_REF_WOULD_ABORT) get zeros and trunc_lo = -1.
 */
typedef struct {
  int32_t n_match, n_subst, n_ins, n_del;
  int32_t trunc_lo, trunc_hi;
  int32_t trunc_start_x, trunc_start_y;
  double rescore;
} bsq_aln_stats;
int bsq_batch_stats(bsq_batch* batch, bsq_aln_stats* out);

/* Aligner.table_scores() (pw.py:278-285): H of every cell of ONE STD problem,
 * (n+1)*(m+1) doubles row-major. Debug/teaching path, not a hot path. */
int bsq_table_scores(bsq_ctx* ctx, const bsq_params* params,
                     const uint8_t* origin, int32_t n,
                     const uint8_t* mutant, int32_t m, double* out);

int bsq_get_timing(const bsq_ctx* ctx, bsq_timing* out);

/*
 * Batched synthetic pairs on the device: rand_seq (stochastics.py:28-41) and
 * MutationProcess.mutate (stochastics.py:143-201) for npairs pairs at once.  Pair i is a
 * random genome chunk; origin = chunk[0, origin_len); the mutant is mutate(chunk[off, off +
 * origin_len)) with off ~ U[0, max_offset] (0: the mutant is mutated from the origin itself);
 * fixed_mutant_len > 0 trims / pads (random letters) every mutant to that length (fixed-length
 * reads).  Randomness is counter based on (seed, pair): a pair does not depend on its batch.
 *  out: codes[] = origin_0 | mutant_0 | origin_1 | ... (+16 zero bytes), *ncodes bytes in all
 *       (BSQ_ERR_CAPACITY with *ncodes set if codes_cap is too small); o_off/o_len/m_off/m_len
 *       as bsq_align_batch takes them; src_off[i] = off (may be NULL); tx_off[npairs+1] /
 *       tx_buf: the planted edit transcripts mutate() returns (both may be NULL).
 */
typedef struct {
  uint64_t seed;
  int32_t alphabet_len;
  int32_t origin_len;
  int32_t max_offset;
  int32_t fixed_mutant_len;
  double subst_prob;     /* probability of any substitution (MutationProcess(subst_probs=float)) */
  double go_prob;        /* gap-open probability  */
  double ge_prob;        /* gap-extend probability, >= go_prob (stochastics.py:140-141) */
} bsq_gen_params;
int bsq_generate_pairs(bsq_ctx* ctx, const bsq_gen_params* gen, int64_t npairs,
                       uint8_t* codes, int64_t codes_cap, int64_t* ncodes,
                       int64_t* o_off, int32_t* o_len, int64_t* m_off, int32_t* m_len,
                       int32_t* src_off, int64_t* tx_off, char* tx_buf, int64_t tx_cap);

/*
 * Pipe-rate microbenchmarks: the measured denominators of the ALU roofline
 * (P_pipe = peak issue rate of the score datatype's ALU
 * op).  *ops_per_sec receives thread-level operations per second for the whole
 * GPU, measured with CUDA events over a dependent-free instruction stream.
 */
#define BSQ_MB_DADD 0      /* add.rn.f64                                   */
#define BSQ_MB_DMAX 1      /* setp.gt.f64 + selp.f64 (one "max" of doubles) */
#define BSQ_MB_DSETP 2     /* setp.eq.f64 feeding a predicated integer op   */
#define BSQ_MB_IADD 3      /* add.s32                                       */
#define BSQ_MB_VIMAX3 4    /* DPX three-way max (VIMNMX3)                   */
#define BSQ_MB_VIADDMAX 5  /* DPX add+max (VIADDMNMX)                       */
#define BSQ_MB_SHFL 6      /* shfl.sync.up.b32                              */
#define BSQ_MB_LDS64 7     /* ld.shared.b64, conflict free                  */
#define BSQ_MB_VADD2 8     /* VIADD.16x2, counted as two halfword adds            */
#define BSQ_MB_COUNT 9
int bsq_microbench(bsq_ctx* ctx, int32_t which, double* ops_per_sec);

/*
 * ONE caller batch over several GPUs of one box (pairs are independent: no collective, no peer
 * traffic; the reference's callers hand one process a list of segment pairs, experiments/
 * blot_stats.py:365-459).  The batch is cut into contiguous ranges of equal work (DP cells), one
 * per GPU; one host thread + one context per GPU runs bsq_align_batch on its range and writes
 * straight into disjoint slices of the caller's arrays (each GPU only receives the window of
 * codes[] its pairs reference).  Same arguments and semantics as bsq_align_batch except for the
 * transcripts: every GPU packs its own region of tx_buf, so transcript i is ops
 * tx_off[i] .. tx_off[i] + tx_len[i] (tx_len is an extra output) and tx_cap must hold the worst
 * case, sum(o_len + m_len) + 64 * bsq_multi_size() ops (a quarter of that in bytes with
 * BSQ_F_TX_2BIT).  devices == NULL means ordinals 0 .. ndev-1; an ordinal may repeat (several
 * contexts on one GPU).
 */
typedef struct bsq_multi bsq_multi;
bsq_multi* bsq_multi_create(int32_t ndev, const int32_t* devices);
void bsq_multi_destroy(bsq_multi* me);
int32_t bsq_multi_size(const bsq_multi* me);
const char* bsq_multi_last_error(const bsq_multi* me);
int bsq_multi_align_batch(bsq_multi* me, const bsq_params* params, int64_t npairs,
                          const uint8_t* codes, int64_t ncodes,
                          const int64_t* o_off, const int32_t* o_len,
                          const int64_t* m_off, const int32_t* m_len,
                          const int32_t* dmin, const int32_t* dmax,
                          double* score, int32_t* end_x, int32_t* end_y,
                          int32_t* start_x, int32_t* start_y, int32_t* status,
                          int64_t* tx_off, int32_t* tx_len, char* tx_buf, int64_t tx_cap);
/* shard k of the last bsq_multi_align_batch: the range of pairs it took, its device, its timing */
int bsq_multi_get_shard(const bsq_multi* me, int32_t k, int64_t* first, int64_t* count,
                        int32_t* device, bsq_timing* timing);

/* pure host helpers (usable without a GPU) ------------------------------- */

/* cells the reference's fill loop visits for one problem (pw.c:59-61 over
 * _xlim/_ylim, _pw_internals.c:116-153); dmin/dmax ignored in STD mode */
int64_t bsq_count_cells(int32_t mode, int32_t n, int32_t m, int32_t dmin, int32_t dmax);

/* Pair-sharding for multi-GPU runs (no collectives on the hot path): LPT-deal
 * pairs, heaviest first, to the least-loaded of nshards; shard_of[i] in
 * [0,nshards).  load[] (nshards, optional) receives cells per shard. */
int bsq_partition_pairs(int32_t mode, int64_t npairs, const int32_t* o_len,
                        const int32_t* m_len, const int32_t* dmin,
                        const int32_t* dmax, int32_t nshards,
                        int32_t* shard_of, int64_t* load);

/* 1 if the I32 exact path applies to these scores and lengths */
int bsq_scores_are_int32_exact(const bsq_params* params, int32_t max_n, int32_t max_m);

#ifdef __cplusplus
}
#endif
#endif /* BSQGPU_H */
