// bsq_host.cu -- C ABI of libbsqgpu (include/bsqgpu.h) and the host-side
// batch scheduler: classifies every pair the way dptable_init would
// (_pw_internals.c:22-62), picks the kernel geometry, orders pairs by size,
// cuts the batch into chunks whose pointer tables fit the HBM arena, and
// drives fill -> traceback -> pack on the context's stream.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

#include "../../include/bsqgpu.h"
#include "bsq_post.cuh"
#include "bsq_gen.cuh"
#include "bsq_launch.h"

int bsq_microbench_impl(int sms, cudaStream_t st, int which, double* out, std::string* err);

namespace {

thread_local std::string g_err;

struct DevBuf {
  void* p = nullptr;
  size_t cap = 0;
  bool ensure(size_t n) {
    if (n <= cap && p) return true;
    if (p) { cudaFree(p); p = nullptr; cap = 0; }
    size_t want = n + n / 16 + 256;
    if (cudaMalloc(&p, want) != cudaSuccess) {
      p = nullptr;
      cudaGetLastError();
      want = n ? n : 256;
      if (cudaMalloc(&p, want) != cudaSuccess) { p = nullptr; cudaGetLastError(); return false; }
    }
    cap = want;
    return true;
  }
  void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
  template <class T> T* as() const { return (T*)p; }
};

struct PinBuf {
  void* p = nullptr;
  size_t cap = 0;
  bool ensure(size_t n) {
    if (n <= cap && p) return true;
    if (p) { cudaFreeHost(p); p = nullptr; cap = 0; }
    size_t want = n + n / 16 + 256;
    if (cudaHostAlloc(&p, want, cudaHostAllocDefault) != cudaSuccess) { p = nullptr; cudaGetLastError(); return false; }
    cap = want;
    return true;
  }
  void release() { if (p) cudaFreeHost(p); p = nullptr; cap = 0; }
  template <class T> T* as() const { return (T*)p; }
};

struct BufSet {  // everything one batch needs in HBM
  DevBuf codes, pairs, ptr, slots, packed, score, ex, ey, sx, sy, status, txlen, txoff, tiles,
      subst, counters, colh, colf, total, k3pairs, k3sync, k3res, k3order, ooff, moff, flags, codes2, packed2;
  PinBuf hpairs, hsmall;
  size_t bytes() const {  // HBM currently held
    const DevBuf* all[] = {&codes, &pairs, &ptr, &slots, &packed, &score, &ex, &ey, &sx, &sy, &status,
                           &txlen, &txoff, &tiles, &subst, &counters, &colh, &colf, &total, &k3pairs, &k3sync, &k3res, &k3order, &ooff, &moff, &flags, &codes2, &packed2};
    size_t n = 0;
    for (const DevBuf* b : all) n += b->cap;
    return n;
  }
  void release() {
    DevBuf* all[] = {&codes, &pairs, &ptr, &slots, &packed, &score, &ex, &ey, &sx, &sy, &status,
                     &txlen, &txoff, &tiles, &subst, &counters, &colh, &colf, &total, &k3pairs, &k3sync, &k3res, &k3order, &ooff, &moff, &flags, &codes2, &packed2};
    for (DevBuf* b : all) b->release();
    hpairs.release();
    hsmall.release();
  }
};

struct Chunk {
  int64_t first, count;  // range in processing order
  int family;            // 1 = K1 (STD), 2 = K2 (BANDED)
  int cfg;
  int64_t ptr_words;
  int max_n_multipass;   // K1: longest origin among pairs needing > 1 pass (0 if none)
  int64_t idx_lo, idx_hi; // caller indices [idx_lo, idx_hi) covered (identity order only)
  int64_t code_end;      // one past the last byte of codes[] this chunk reads
  bool duo = false;      // K1h: pairs (2q, 2q+1) of the chunk share a lane group, packed 16-bit arithmetic
  int64_t k3_strips = 0; // K3: strips in the chunk's work list
  int64_t k3_slab = 0;   // K3: boundary-column slab entries (sum of n+1)
  int64_t k3_active = 0; // K3: strips in flight when every pair's pipeline is full
  int64_t k3_order_off = 0;  // K3: where the chunk's strip work list starts in Plan::k3order
};

// largest pointer table (words) of a batch that is aligned in caller order instead of sorted:
// short reads only (a 512 x 512 STD table has 34816 words)
#define BSQ_INORDER_MAX_WORDS 65536

// three slots: the fill of chunk c+2 need not wait for the traceback of chunk c, which shares
// the SMs with fill c+1 and often takes longer than a fill
#define BSQ_ARENA_SLOTS 3

inline bool chunk_is_k3(const Chunk& ch) { return ch.family == 2 && ch.cfg >= BSQ_K2_NCFG; }

struct Override { int64_t idx; int status; };

// Test / tuning hooks.  Environment variables, read ONCE per context in bsq_create (listed in
// tools/README.md); nothing on the call path reads the environment.
struct Hooks {
  bool force_k3 = false;        // BSQ_FORCE_K3: K3 for every band (test)
  bool no_duo = false;          // BSQ_NO_DUO: never use K1h (test)
  bool always_sort = false;     // BSQ_ALWAYS_SORT: no caller-order planners (test)
  bool serial_scan = false;     // BSQ_SERIAL_SCAN: serial chunk-cutting scan (test)
  bool plan_debug = false;      // BSQ_PLAN_DEBUG: planner / chunk timeline on stderr
  bool no_std_band = false;     // BSQ_NO_STD_BAND: big STD GLOBAL pairs stay on K1 (test)
  int64_t uniform_min = 4096;   // BSQ_UNIFORM_MIN: pairs from which the uniform plan / K1h are used
  int64_t plan_block = 4096;    // BSQ_PLAN_BLOCK: block size of the caller-order planner
  int target_chunks = -1;       // BSQ_TARGET_CHUNKS: chunks to cut a batch into (< 0: automatic)
  int k2_cfg = -1;              // BSQ_K2_CFG: forced K2 geometry
  int k3_cfg = -1;              // BSQ_K3_CFG: forced K3 geometry
  int std_band_max = 4096;      // BSQ_STD_BAND_MAX: most STD GLOBAL pairs of a batch planned as full-width bands
  char trace_kernel = 0;        // BSQ_TRACE_KERNEL: 't'hread / 'w'arp / 'n'arrow / 'p'air-thread (banded)
  size_t arena_cap_bytes = 0;   // BSQ_ARENA_CAP_BYTES: upper bound for the pointer arena
  void read_env() {
    auto num = [](const char* name, long long dflt) { const char* e = getenv(name); return e ? atoll(e) : dflt; };
    force_k3 = getenv("BSQ_FORCE_K3") != nullptr;
    no_duo = getenv("BSQ_NO_DUO") != nullptr;
    always_sort = getenv("BSQ_ALWAYS_SORT") != nullptr;
    serial_scan = getenv("BSQ_SERIAL_SCAN") != nullptr;
    plan_debug = getenv("BSQ_PLAN_DEBUG") != nullptr;
    no_std_band = getenv("BSQ_NO_STD_BAND") != nullptr;
    uniform_min = num("BSQ_UNIFORM_MIN", 4096);
    plan_block = std::max<long long>(1, num("BSQ_PLAN_BLOCK", 4096));
    target_chunks = (int)num("BSQ_TARGET_CHUNKS", -1);
    k2_cfg = (int)num("BSQ_K2_CFG", -1);
    k3_cfg = (int)num("BSQ_K3_CFG", -1);
    std_band_max = (int)num("BSQ_STD_BAND_MAX", 4096);
    const char* tk = getenv("BSQ_TRACE_KERNEL");
    trace_kernel = tk ? tk[0] : 0;
    arena_cap_bytes = (size_t)num("BSQ_ARENA_CAP_BYTES", 0);
  }
};
// pairs per planner thread before another one is worth waking; process-wide like the worker pool
// (BSQ_PAR_GRAIN, read in bsq_create: a test hook so that small batches exercise the threaded passes)
int64_t g_par_grain = 32768;

struct PlanItem {  // per-pair classification, filled in parallel
  int64_t words;   // pointer-table words
  int64_t cells;
  int32_t cfg;     // kernel geometry, or -1 - status for pairs that never reach the GPU
  int32_t aux;     // K1: n if multi-pass else 0; K2: first step s0
  int32_t lo, hi;  // clipped band
};


struct Plan {
  bsq_params prm;
  std::vector<double> subst;
  int64_t N = 0;
  int dtype = BSQ_DTYPE_F64;
  int gosign = 0;
  int bmode = 0, emode = 0;
  BsqPair* pairs = nullptr;        // processing order; lives in pinned host memory (BufSet.hpairs)
  int64_t E = 0;                   // number of pair records
  std::vector<Chunk> chunks;
  std::vector<Override> overrides; // pairs never sent to the GPU
  int64_t cells = 0;
  int64_t slot_bytes = 0;
  int64_t max_chunk_words = 0;
  int64_t total_words = 0;
  bool identity = false;  // processing order == caller order (no sort was needed)
  bool codes2_ok = false; // BufSet.codes2 holds the 2-bit packed copy of this batch's letters
  int64_t max_nm = 0;     // largest n + m of the batch
  double score_bound = 0; // (max n + max m + 4) * largest |score|: no partial sum leaves it
  int64_t half_words = 0; // > 0: BSQ_ARENA_SLOTS arena slots of this size, chunk c uses slot c % SLOTS so that
                          // trace(c) overlaps the fills of the following chunks
  // uniform batches (every pair the same shape, e.g. fixed-length reads): the pair records are
  // built ON THE DEVICE from the caller's offset arrays (bsq_make_pairs_uniform); pairs == nullptr
  bool uniform_dev = false;
  BsqUniform uni;
  const int64_t* src_o_off = nullptr;
  const int64_t* src_m_off = nullptr;
  std::vector<BsqK3Pair> k3;  // per pair record (processing order); empty when no pair uses K3
  std::vector<BsqK3Item> k3order;  // the strip work lists of the K3 chunks, one after the other
  int64_t k3_max_strips = 0, k3_max_slab = 0, k3_max_pairs = 0;  // maxima over chunks
};

// grid.y of bsq_pack_transcripts: segments of the longest possible transcript (n + m ops)
inline unsigned pack_segments(const Plan& P) {
  const int64_t seg = (P.max_nm + BSQ_PACK_SEG - 1) / BSQ_PACK_SEG;
  return (unsigned)std::min<int64_t>(std::max<int64_t>(seg, 1), 65535);
}

}  // namespace

struct bsq_ctx {
  int device = 0;
  int sms = 1;
  cudaStream_t st = nullptr;       // compute
  cudaStream_t st_copy = nullptr;  // H2D / D2H, so copies overlap kernels
  cudaStream_t st_copy2 = nullptr; // pair records / score table: must not queue behind the sequence bytes
  cudaStream_t st_d2h = nullptr;   // results back to the host while sequence bytes still travel the other way
  cudaStream_t st_trace = nullptr; // traceback + packing of chunk c while chunk c+1 fills
  PinBuf htotals;                  // pinned: running transcript totals per chunk
  std::vector<PlanItem> plan_items;  // planner scratch, kept across calls (no page faults)
  std::vector<int64_t> plan_order, plan_ptr_off, plan_slot_off, plan_sorted;
  std::vector<int64_t> plan_ooff, plan_moff;   // offsets of a back-to-back batch given without o_off / m_off
  std::vector<int32_t> plan_band_lo, plan_band_hi;   // full bands of STD GLOBAL pairs planned as wide bands
  std::vector<char> plan_gather;
  std::string err;
  bsq_timing tm;
  BufSet bufs;
  std::vector<cudaEvent_t> events;
  Hooks hooks;                 // test / tuning hooks, read from the environment once in bsq_create
  // cudaMemGetInfo costs 0.1 - 50 ms (measured on the bench box): the one-shot path remembers
  // "free bytes + our cached arena" and only asks again after a resident batch came or went or
  // an allocation failed
  double mem_avail = -1.0;         // HBM that would be free if this context held nothing; < 0: unknown
  // buffer sets of released resident / two-phase batches, kept for the next one: an Aligner used
  // one pair at a time (the reference's own calling pattern) must not pay ~15 cudaMalloc +
  // cudaFree + a pinned allocation per pair (measured: 49 ms per 100 x 100 pair without this)
  std::vector<BufSet*> pool;
  std::vector<bsq_batch*> live;    // resident batches that own a buffer set right now
};

struct bsq_batch {
  bsq_ctx* ctx = nullptr;
  Plan plan;
  BufSet* own = nullptr; // used when own_bufs (taken from / returned to the context's pool)
  BufSet* B = nullptr;  // the buffers in use (own or the context's cached set)
  bool own_bufs = false;
  bool single_chunk = false;  // all pointer tables resident at once
  bool filled = false, traced = false;
};

// How many chunks to cut a batch into so that traceback/copies of one chunk overlap the
// fill of the next.  Only short pairs profit: their chunks still hold many waves of work,
// and their thread-per-pair walk fits beside the fill; long pairs are cut by capacity only.
static int overlap_chunks(const bsq_ctx* ctx, int64_t npairs, const int32_t* o_len, const int32_t* m_len) {
  if (ctx->hooks.target_chunks >= 0) return ctx->hooks.target_chunks;  // tuning hook
  if (npairs < 32768) return 1;
  int64_t tot = 0;
  const int64_t step = std::max<int64_t>(npairs / 1024, 1);
  int64_t cnt = 0;
  for (int64_t i = 0; i < npairs; i += step) { tot += (int64_t)o_len[i] + m_len[i]; cnt++; }
  if (tot / std::max<int64_t>(cnt, 1) >= 1024) return 1;
  return (int)std::min<int64_t>(8, npairs / 16384);
}

namespace {

#define BSQ_CK(ctx, call)                                                          \
  do {                                                                             \
    cudaError_t e_ = (call);                                                       \
    if (e_ != cudaSuccess) {                                                       \
      (ctx)->err = std::string(#call) + ": " + cudaGetErrorString(e_);             \
      return BSQ_ERR_CUDA;                                                         \
    }                                                                              \
  } while (0)

int fail(bsq_ctx* ctx, int code, const std::string& msg) {
  if (ctx) ctx->err = msg;
  g_err = msg;
  return code;
}

// ---- geometry helpers (pure host) ---------------------------------------

// cells of one diagonal d of an (n+1)x(m+1) table: row_lens[], _pw_internals.c:56
inline int64_t diag_len(int n, int m, int d) {
  return 1 + (int64_t)(d > 0 ? 0 : d) + (int64_t)((int64_t)n - d > m ? m : n - d);
}

int64_t count_cells(int mode, int n, int m, int dmin, int dmax) {
  if (mode == BSQ_STD_MODE) return (int64_t)(n + 1) * (m + 1);
  if (dmax > n) dmax = n;
  if (dmin < -m) dmin = -m;
  if (dmax < dmin) return 0;
  // sum of diag_len over [dmin,dmax], closed form over the three linear pieces
  auto sum_range = [&](int64_t a, int64_t b, auto f) -> int64_t {  // f linear in d
    if (b < a) return 0;
    return (f(a) + f(b)) * (b - a + 1) / 2;
  };
  const int64_t lo = std::min(0, n - m), hi = std::max(0, n - m);
  int64_t acc = 0;
  {  // d in [dmin, min(dmax, lo-1)]: len rises with slope 1
    int64_t a = dmin, b = std::min<int64_t>(dmax, lo - 1);
    acc += sum_range(a, b, [&](int64_t d) { return diag_len(n, m, (int)d); });
  }
  {  // d in [max(dmin,lo), min(dmax,hi)]: plateau (constant)
    int64_t a = std::max<int64_t>(dmin, lo), b = std::min<int64_t>(dmax, hi);
    acc += sum_range(a, b, [&](int64_t d) { return diag_len(n, m, (int)d); });
  }
  {  // d in [max(dmin,hi+1), dmax]: falls with slope -1
    int64_t a = std::max<int64_t>(dmin, hi + 1), b = dmax;
    acc += sum_range(a, b, [&](int64_t d) { return diag_len(n, m, (int)d); });
  }
  return acc;
}

inline bool is_integer(double v) { return std::isfinite(v) && v == std::floor(v); }

// Bound on |any partial score|: every path has at most n+m ops, each adding at
// most maxabs (a gap op adds ge and possibly go).
bool scores_int32_exact(const bsq_params* p, int64_t max_n, int64_t max_m) {
  const int L = p->alphabet_len;
  double maxabs = std::fabs(p->go) + std::fabs(p->ge);
  if (!is_integer(p->go) || !is_integer(p->ge)) return false;
  for (int i = 0; i < L * L; i++) {
    if (!is_integer(p->subst[i])) return false;
    maxabs = std::max(maxabs, std::fabs(p->subst[i]));
  }
  const double R = (double)(max_n + max_m + 4) * maxabs;
  return R < (double)(1 << 25);  // H * 32 + column must fit an int (packed end-cell key)
}

void std_modes(int type, int* bmode, int* emode) {
  // _alnchoice_B (_pw_internals.c:165-181) and _std_find_optimal (:309-355)
  switch (type) {
    case BSQ_GLOBAL: *bmode = BSQ_B_ORIGIN; *emode = BSQ_E_CORNER; break;
    case BSQ_LOCAL: *bmode = BSQ_B_ALL; *emode = BSQ_E_ALL; break;
    case BSQ_START_ANCHORED: *bmode = BSQ_B_ORIGIN; *emode = BSQ_E_ALL; break;
    case BSQ_END_ANCHORED: *bmode = BSQ_B_ALL; *emode = BSQ_E_CORNER; break;
    case BSQ_OVERLAP: *bmode = BSQ_B_EDGE; *emode = BSQ_E_EDGE; break;
    case BSQ_START_ANCHORED_OVERLAP: *bmode = BSQ_B_ORIGIN; *emode = BSQ_E_EDGE; break;
    default: *bmode = BSQ_B_EDGE; *emode = BSQ_E_CORNER; break;  // END_ANCHORED_OVERLAP
  }
}

int pick_k1_cfg(int n, int m) {
  int best = 0;
  double bestc = 1e300;
  for (int c = 0; c < BSQ_K1_NCFG; c++) {
    const int G = BSQ_K1_CFGS[c].G, K = BSQ_K1_CFGS[c].K, BW = G * K;
    const double passes = (double)((m + BW) / BW);
    const double cost = passes * ((double)n + G) * BW / (G == 32 ? 1.0 : 1.0);
    if (cost < bestc) { bestc = cost; best = c; }
  }
  return best;
}

// Geometry for a band of W diagonals: least estimated time = lane-cells / relative
// speed.  Measured on B200 (profiles/r01_k2_geometry.txt): for doubles the
// register footprint of C >= 16 diagonals per lane costs occupancy, so lanes with
// fewer diagonals win even at lower lane efficiency; int32 prefers tight fits.
int pick_k2_cfg(int W, bool f64, int forced) {
  if (forced >= 0 && forced < BSQ_K2_NCFG && BSQ_K2_CFGS[forced].G * BSQ_K2_CFGS[forced].C >= W) return forced;
  int best = -1;
  double bestc = 1e300;
  for (int c = 0; c < BSQ_K2_NCFG; c++) {
    const int G = BSQ_K2_CFGS[c].G, C = BSQ_K2_CFGS[c].C, cap = G * C;
    if (cap < W) continue;
    double speed = 1.0;
    if (f64) speed = C <= 10 ? 1.0 : (C <= 16 ? 0.85 : 0.78);
    else speed = C <= 10 ? 0.97 : 1.0;
    const double cost = cap / speed;
    if (cost < bestc) { bestc = cost; best = c; }
  }
  return best;
}

int k1_op(int dtype, int cfg, int ball, int eall, int gos, int op, int grid, cudaStream_t st, const BsqFillArgs* A) {
  return dtype == BSQ_DTYPE_I32 ? bsq_k1_i32_op(cfg, ball, eall, gos, op, grid, st, A)
                                : bsq_k1_f64_op(cfg, ball, eall, gos, op, grid, st, A);
}
int k2_op(int dtype, int cfg, int blocal, int gos, int op, int grid, cudaStream_t st, const BsqFillArgs* A) {
  return dtype == BSQ_DTYPE_I32 ? bsq_k2_i32_op(cfg, blocal, gos, op, grid, st, A)
                                : bsq_k2_f64_op(cfg, blocal, gos, op, grid, st, A);
}

int k3_op(int dtype, int cfg, int btype, int gos, int grid, cudaStream_t st, const BsqFillArgs* A) {
  return dtype == BSQ_DTYPE_I32 ? bsq_k3_i32_op(cfg, btype, gos, grid, st, A)
                                : bsq_k3_f64_op(cfg, btype, gos, grid, st, A);
}



// ---- planning -------------------------------------------------------------

// A small persistent worker pool: the planner's passes over a million pairs take well under a
// millisecond each, less than spawning threads for them would.
class WorkerPool {
 public:
  // never destroyed: a forked child (or exit() from foreign code) must not try to join threads
  // that do not exist there
  static WorkerPool& get() { static WorkerPool* p = new WorkerPool(); return *p; }
  int size() const { return (int)th_.size() + 1; }
  // runs job(t) for t in [0, T) -- t = 0 on the caller -- and returns when all are done
  void run(int T, const std::function<void(int)>& job) {
    if (T <= 1) { job(0); return; }
    std::lock_guard<std::mutex> one_caller(run_mu_);  // contexts on different host threads share the pool
    {
      std::lock_guard<std::mutex> lk(mu_);
      job_ = &job;
      want_ = T;
      pending_ = T - 1;
      gen_++;
    }
    cv_.notify_all();
    job(0);
    std::unique_lock<std::mutex> lk(mu_);
    done_.wait(lk, [&] { return pending_ == 0; });
    job_ = nullptr;
  }
 private:
  WorkerPool() {
    unsigned hw = std::thread::hardware_concurrency();
    const int n = (int)std::max(1u, std::min(hw, 8u)) - 1;
    for (int i = 0; i < n; i++) { th_.emplace_back([this, i] { loop(i + 1); }); th_.back().detach(); }
  }
  void loop(int id) {
    uint64_t seen = 0;
    for (;;) {
      const std::function<void(int)>* job = nullptr;
      {
        std::unique_lock<std::mutex> lk(mu_);
        cv_.wait(lk, [&] { return gen_ != seen; });
        seen = gen_;
        if (stop_) return;
        if (id < want_) job = job_;
      }
      if (job) {
        (*job)(id);
        std::lock_guard<std::mutex> lk(mu_);
        if (--pending_ == 0) done_.notify_all();
      }
    }
  }
  std::vector<std::thread> th_;
  std::mutex mu_, run_mu_;
  std::condition_variable cv_, done_;
  const std::function<void(int)>* job_ = nullptr;
  int want_ = 0, pending_ = 0;
  uint64_t gen_ = 0;
  bool stop_ = false;
};

inline int64_t parallel_grain() { return g_par_grain > 0 ? g_par_grain : 32768; }
// bsq_multi runs one host thread per GPU, each planning its own shard: those plan single-threaded
// (N planners x 8 pool threads would only queue up on the pool's lock)
thread_local bool g_plan_serial = false;
inline int parallel_threads(int64_t n) {
  if (g_plan_serial) return 1;
  return (int)std::min<int64_t>(WorkerPool::get().size(), std::max<int64_t>(n / parallel_grain(), 1));
}

template <class F>
void parallel_for(int64_t n, F f) {
  WorkerPool& pool = WorkerPool::get();
  const int T = parallel_threads(n);
  if (T <= 1) { f(0, (int64_t)0, n); return; }
  pool.run(T, [&](int t) { f(t, n * t / T, n * (t + 1) / T); });
}

int make_plan(bsq_ctx* ctx, Plan& P, PinBuf& hpairs, const bsq_params* prm, int64_t N, int64_t ncodes,
              const int64_t* o_off, const int32_t* o_len, const int64_t* m_off,
              const int32_t* m_len, const int32_t* dmin, const int32_t* dmax,
              int64_t arena_words, bool single_chunk, int target_chunks = 0, int64_t ob = 0) {
  // ob: added to every caller offset (bsq_multi hands each GPU a window of the caller's codes[]);
  // with implicit back-to-back offsets it is where the first pair starts
  if (!prm || !prm->subst) return fail(ctx, BSQ_ERR_ARG, "params/subst is NULL");
  const int L = prm->alphabet_len;
  if (L < 1 || L > BSQ_MAX_L) return fail(ctx, BSQ_ERR_ARG, "alphabet_len must be in [1,32]");
  if (prm->mode != BSQ_STD_MODE && prm->mode != BSQ_BANDED_MODE) return fail(ctx, BSQ_ERR_ARG, "bad alnmode");
  if (prm->mode == BSQ_STD_MODE && (prm->type < 0 || prm->type > 6)) return fail(ctx, BSQ_ERR_ARG, "bad std_alntype");
  if (prm->mode == BSQ_BANDED_MODE && (prm->type < 0 || prm->type > 2)) return fail(ctx, BSQ_ERR_ARG, "bad banded_alntype");
  if (prm->mode == BSQ_BANDED_MODE && (!dmin || !dmax)) return fail(ctx, BSQ_ERR_ARG, "BANDED needs dmin/dmax");
  if (prm->max_new_mins > 0)
    return fail(ctx, BSQ_ERR_ARG, "max_new_mins > 0 is undefined behaviour in the reference (_pw_internals.c:236-243); not supported");
  if (!std::isfinite(prm->go) || !std::isfinite(prm->ge)) return fail(ctx, BSQ_ERR_ARG, "go/ge must be finite");
  if ((prm->flags & BSQ_F_CODES_2BIT) && L > 4) return fail(ctx, BSQ_ERR_ARG, "BSQ_F_CODES_2BIT needs an alphabet of at most 4 letters");
  if (prm->flags & ~(BSQ_F_CODES_2BIT | BSQ_F_TX_2BIT)) return fail(ctx, BSQ_ERR_ARG, "unknown bits in params.flags");
  // One big STD GLOBAL problem -- the reference's own calling pattern, config 1 -- or a few thousand
  // of them are latency bound on K1 (one warp walks the whole table, two passes for 1 kb: 3.3 ms for
  // 16 pairs, 10 ms for 1024).  GLOBAL is B_GLOBAL with the full band [-m, n] (same B rule, same end
  // cell (n,m), same traceback: _pw_internals.c:165-181 vs :182-197, :309-315 vs :366-378), so they
  // are planned as wide bands and take the strip pipeline (K3L): every 64 columns another warp,
  // the pipelines of all pairs side by side (0.19 ms for 16 pairs, 2.6 ms for 1024).  Beyond a few
  // thousand pairs K1 has enough warps of its own.
  bsq_params eff = *prm;
  bool std_as_band = false;
  if (prm->mode == BSQ_STD_MODE && prm->type == BSQ_GLOBAL && N >= 1 && N <= ctx->hooks.std_band_max && !ctx->hooks.no_std_band) {
    std_as_band = true;
    for (int64_t i = 0; i < N; i++) if (o_len[i] < 256 || m_len[i] < 256) std_as_band = false;
  }
  if (std_as_band) {
    ctx->plan_band_lo.resize((size_t)N);
    ctx->plan_band_hi.resize((size_t)N);
    for (int64_t i = 0; i < N; i++) { ctx->plan_band_lo[(size_t)i] = -m_len[i]; ctx->plan_band_hi[(size_t)i] = o_len[i]; }
    dmin = ctx->plan_band_lo.data();
    dmax = ctx->plan_band_hi.data();
    eff.mode = BSQ_BANDED_MODE;
    eff.type = BSQ_B_GLOBAL;
    prm = &eff;
  }
  P.prm = *prm;
  P.subst.assign(prm->subst, prm->subst + L * L);
  for (double& v : P.subst) {
    if (!std::isfinite(v)) return fail(ctx, BSQ_ERR_ARG, "subst scores must be finite");
    if (v == 0.0) v = 0.0;  // canonicalise -0.0
  }
  P.prm.subst = nullptr;
  P.N = N;
  P.gosign = prm->go < 0 ? -1 : (prm->go > 0 ? 1 : 0);
  if (prm->mode == BSQ_STD_MODE) std_modes(prm->type, &P.bmode, &P.emode);
  const int family = prm->mode == BSQ_STD_MODE ? 1 : 2;
  const int btype = prm->type;

  // one parallel pass over the lengths: maxima (for the arithmetic / overflow decisions), input
  // validation, and whether every pair has the same shape
  // o_off == m_off == NULL: the pairs lie back to back in codes[].  A uniform batch never needs the
  // offsets on the host (the device derives them); every other plan gets them materialised first.
  const bool implicit = (o_off == nullptr);
  auto materialize_offsets = [&]() {
    std::vector<int64_t>& io = ctx->plan_ooff;
    std::vector<int64_t>& im = ctx->plan_moff;
    if ((int64_t)io.size() < N) { io.resize((size_t)N); im.resize((size_t)N); }
    int64_t pos = 0;   // local offsets: ob is added wherever they are read, like for explicit ones
    for (int64_t i = 0; i < N; i++) { io[(size_t)i] = pos; pos += o_len[i]; im[(size_t)i] = pos; pos += m_len[i]; }
    o_off = io.data();
    m_off = im.data();
  };
  struct Scan { int64_t max_n = 0, max_m = 0, tot = 0; int bad = 0, uniform = 1; };
  std::vector<Scan> scans(16);
  {
    const int n0 = N ? o_len[0] : 0, m0 = N ? m_len[0] : 0;
    const int lo0 = (N && family == 2) ? dmin[0] : 0, hi0 = (N && family == 2) ? dmax[0] : 0;
    parallel_for(N, [&](int t, int64_t i0, int64_t i1) {
      Scan sc;
      for (int64_t i = i0; i < i1; i++) {
        const int n = o_len[i], m = m_len[i];
        if (n < 0 || m < 0 || (!implicit && (o_off[i] + ob < 0 || m_off[i] + ob < 0 || o_off[i] + ob + n > ncodes || m_off[i] + ob + m > ncodes))) { sc.bad = 1; continue; }
        sc.max_n = std::max<int64_t>(sc.max_n, n);
        sc.max_m = std::max<int64_t>(sc.max_m, m);
        sc.tot += (int64_t)n + m;
        if (n != n0 || m != m0 || (family == 2 && (dmin[i] != lo0 || dmax[i] != hi0))) sc.uniform = 0;
      }
      scans[t] = sc;
    });
  }
  int64_t scan_n = 0, scan_m = 0, scan_tot = 0;
  bool all_uniform = N > 0;
  for (const Scan& sc : scans) {
    if (sc.bad) return fail(ctx, BSQ_ERR_ARG, "negative length or sequence range outside codes[]");
    scan_n = std::max(scan_n, sc.max_n); scan_m = std::max(scan_m, sc.max_m); scan_tot += sc.tot;
    if (!sc.uniform) all_uniform = false;
  }
  if (implicit && (ob < 0 || ob + scan_tot > ncodes)) return fail(ctx, BSQ_ERR_ARG, "back-to-back pairs do not fit codes[]");
  P.max_nm = scan_n + scan_m;
  {
    double maxabs = std::fabs(prm->go) + std::fabs(prm->ge);
    for (double v : P.subst) maxabs = std::max(maxabs, std::fabs(v));
    P.score_bound = (double)(scan_n + scan_m + 4) * maxabs;
  }
  // geometry depends on the arithmetic; decide it up front from the scores and maxima
  bool f64_geom = true;
  {
    bsq_params t0 = *prm;
    t0.subst = P.subst.data();
    f64_geom = (prm->dtype == BSQ_DTYPE_F64) || !(prm->dtype == BSQ_DTYPE_I32 || scores_int32_exact(&t0, scan_n, scan_m));
  }
  P.uniform_dev = false;
  P.pairs = nullptr;
  const Hooks& hk = ctx->hooks;
  const bool force_k3 = hk.force_k3 || std_as_band;
  // strip geometry: cfg 0 (bsq_k3l_fill, 64-column strips, a worker and a feeder warp per strip) for
  // every plan -- config 5 fills in 118 ms against 367 ms with cfg 1, config 1 (a full-width band) in
  // 0.2 ms against 3.04 ms on K1, and it also wins on throughput (10 pairs of config 5: 1.35 s against
  // 3.39 s).  cfg 1 / 2 (bsq_k3_fill, 4 / 8 columns x 2 rows per lane and step) stay selectable.
  const int k3_cfg = (hk.k3_cfg >= 0 && hk.k3_cfg < BSQ_K3_NCFG) ? hk.k3_cfg : 0;
  // K1h (two short problems per lane group, packed 16-bit arithmetic): what the scores and
  // the alignment type must satisfy; the shape conditions are per pair (PlanItem.lo for STD)
  double duo_maxabs = std::fabs(prm->go) + std::fabs(prm->ge);
  for (double v : P.subst) duo_maxabs = std::max(duo_maxabs, std::fabs(v));
  bool duo_scores_ok = false;
  {
    bsq_params t1 = *prm;
    t1.subst = P.subst.data();
    const bool i32 = prm->dtype != BSQ_DTYPE_F64 && scores_int32_exact(&t1, scan_n, scan_m);
    duo_scores_ok = family == 1 && i32 && P.bmode == BSQ_B_ALL && P.emode == BSQ_E_ALL && prm->go <= 0.0 &&
                    prm->ge <= 0.0 && L <= 4 && !hk.no_duo;
  }
  auto duo_shape_ok = [&](int cfg, int64_t passes, int n, int m) -> bool {
    return duo_scores_ok && (cfg == 0 || cfg == 1) && passes == 1 && (double)(n + m + 4) * duo_maxabs < 4000.0;
  };
  // what dptable_init would make of one problem shape, the kernel geometry, table size, cells
  auto classify = [&](int n, int m, int dlo, int dhi) -> PlanItem {
    PlanItem it{};
    if (family == 1) {
      it.cfg = pick_k1_cfg(n, m);
      const int G = BSQ_K1_CFGS[it.cfg].G, K = BSQ_K1_CFGS[it.cfg].K, BW = G * K, WPL = BSQ_K1_WPL(K);
      const int64_t passes = (m + BW) / BW;
      it.words = passes * (int64_t)(n + G) * G * WPL;  // rows t = x + lane, see bsq_kernels.cuh
      it.aux = passes > 1 ? n : 0;
      it.cells = (int64_t)(n + 1) * (m + 1);
      it.lo = duo_shape_ok(it.cfg, passes, n, m) ? 1 : 0;   // STD has no band: lo flags K1h eligibility
    } else {
      int lo = dlo, hi = dhi;
      if (hi > n) hi = n;   // _pw_internals.c:29-36
      if (lo < -m) lo = -m;
      const int dend = n - m;
      const int64_t W = (int64_t)hi - lo + 1;
      if (btype == BSQ_B_GLOBAL && (dend > hi || dend < lo || (int64_t)hi * lo > 0)) it.cfg = -1 - BSQ_INIT_FAIL;
      else if (W < 0) it.cfg = -1 - BSQ_INIT_FAIL;
      else if (W == 0) it.cfg = -1 - BSQ_NO_ALIGNMENT;
      else {
        it.cfg = force_k3 ? -1 : pick_k2_cfg((int)std::min<int64_t>(W, 1 << 30), f64_geom, ctx->hooks.k2_cfg);
        if (it.cfg < 0) it.cfg = BSQ_K2_NCFG + k3_cfg;
      }
      if (it.cfg >= BSQ_K2_NCFG) {  // K3: strips of G*K columns, tstride time steps each
        it.lo = lo; it.hi = hi;
        const int K = BSQ_K3_CFGS[it.cfg - BSQ_K2_NCFG].K, G = 32, BW = G * K;
        const int64_t nstrips = ((int64_t)m + BW) / BW;
        // time steps per strip, padded to whole groups of 8: K3L stores every step of a group
        // unconditionally (slots of idle lanes and of the padding are never read)
        const int64_t tstride = (std::min<int64_t>((int64_t)n + 1, W + BW - 1) + G - 1 + 7) / 8 * 8;
        it.words = nstrips * tstride * G * (K / 2) / 4;
        it.aux = (int32_t)tstride;
        it.cells = count_cells(BSQ_BANDED_MODE, n, m, lo, hi);
      } else if (it.cfg >= 0) {
        it.lo = lo; it.hi = hi;
        const int G = BSQ_K2_CFGS[it.cfg].G, C = BSQ_K2_CFGS[it.cfg].C, TB = BSQ_K2_T(C);
        const int smin = (lo <= 0 && hi >= 0) ? 0 : std::min(std::abs(lo), std::abs(hi));
        const int sbase = smin - ((smin - lo) & 1);
        const int dstar = std::max(lo, std::min(hi, dend));
        const int s_end = 2 * std::min(n, m + dstar) - dstar;
        const int64_t steps = (int64_t)(s_end - sbase + 1 + 1) / 2 * 2 + 2;
        it.words = ((steps + TB - 1) / TB) * G * BSQ_K2_UNIT_WORDS;   // [block of TB steps][lane][32-byte unit]
        it.aux = sbase;
        it.cells = count_cells(BSQ_BANDED_MODE, n, m, lo, hi);
      }
    }
    return it;
  };

  // ---- fast path: one shape for the whole batch (fixed-length reads).  No per-pair host work:
  // chunk boundaries are arithmetic and the pair records are generated on the device.
  if (all_uniform && N >= hk.uniform_min && N < ((int64_t)1 << 31)) {
    const PlanItem it = classify(o_len[0], m_len[0], family == 2 ? dmin[0] : 0, family == 2 ? dmax[0] : 0);
    bsq_params tmp = *prm;
    tmp.subst = P.subst.data();
    const bool i32ok = scores_int32_exact(&tmp, scan_n, scan_m);
    if (prm->dtype == BSQ_DTYPE_I32 && !i32ok)
      return fail(ctx, BSQ_ERR_ARG, "dtype I32 forced but scores are not int32-exact for these lengths");
    const int dtype = (prm->dtype == BSQ_DTYPE_F64) ? BSQ_DTYPE_F64 : (i32ok ? BSQ_DTYPE_I32 : BSQ_DTYPE_F64);
    bool range_ok = true;
    if (dtype == BSQ_DTYPE_F64) {
      double maxabs = std::fabs(prm->go) + std::fabs(prm->ge);
      for (double v : P.subst) maxabs = std::max(maxabs, std::fabs(v));
      range_ok = (double)(scan_n + scan_m + 4) * maxabs < 2147483647.0;
    }
    const bool k3 = family == 2 && it.cfg >= BSQ_K2_NCFG;
    if (it.cfg >= 0 && !k3 && range_ok && it.words > 0) {
      P.dtype = dtype;
      P.overrides.clear();
      P.k3.clear();
      P.k3_max_strips = P.k3_max_slab = P.k3_max_pairs = 0;
      P.identity = true;
      P.E = N;
      P.cells = it.cells * N;
      const int64_t tot_words = it.words * N;
      if (target_chunks > 1 && !single_chunk) {
        const int64_t per = std::max<int64_t>(std::max<int64_t>((tot_words + target_chunks - 1) / target_chunks, (int64_t)8 << 20), it.words);
        arena_words = std::min(arena_words, per);
      }
      if (it.words > arena_words) return fail(ctx, BSQ_ERR_NOMEM, "a single pair's traceback table exceeds the device arena");
      if (single_chunk && tot_words > arena_words)
        return fail(ctx, BSQ_ERR_NOMEM, "batch does not fit the device arena in two-phase mode; use bsq_align_batch");
      const int64_t ppc = single_chunk ? N : std::max<int64_t>(1, std::min<int64_t>(N, arena_words / it.words));
      const int64_t NC = (N + ppc - 1) / ppc;
      P.chunks.assign((size_t)NC, Chunk{});
      for (int64_t c = 0; c < NC; c++) {
        Chunk& ch = P.chunks[(size_t)c];
        ch.first = c * ppc;
        ch.count = std::min<int64_t>(ppc, N - ch.first);
        ch.family = family;
        ch.cfg = it.cfg;
        ch.ptr_words = ch.count * it.words;
        ch.max_n_multipass = family == 1 ? it.aux : 0;
        ch.idx_lo = ch.first;
        ch.idx_hi = ch.first + ch.count;
        ch.code_end = 0;
      }
      if (implicit) {
        for (int64_t c = 0; c < NC; c++) P.chunks[(size_t)c].code_end = ob + std::min<int64_t>(N, (c + 1) * ppc) * ((int64_t)o_len[0] + m_len[0]);
      } else {  // last byte of codes[] each chunk reads (pairs need not be laid out in order)
        std::vector<int64_t> ends((size_t)NC * 16, 0);
        const int n = o_len[0], m = m_len[0];
        parallel_for(N, [&](int t, int64_t i0, int64_t i1) {
          for (int64_t i = i0; i < i1; i++) {
            int64_t& e = ends[(size_t)(i / ppc) * 16 + t];
            e = std::max(e, ob + std::max(o_off[i] + n, m_off[i] + m));
          }
        });
        for (int64_t c = 0; c < NC; c++)
          for (int t = 0; t < 16; t++) P.chunks[(size_t)c].code_end = std::max(P.chunks[(size_t)c].code_end, ends[(size_t)c * 16 + t]);
      }
      // K1h: short reads, small integer scores, STD LOCAL -> two problems per lane
      for (Chunk& ch : P.chunks) ch.duo = (family == 1 && it.lo == 1 && dtype == BSQ_DTYPE_I32);
      P.max_chunk_words = std::min(ppc, N) * it.words;
      P.total_words = tot_words;
      P.slot_bytes = N * ((int64_t)o_len[0] + m_len[0]);
      P.half_words = (!single_chunk && NC >= 2) ? P.max_chunk_words : 0;
      P.uniform_dev = true;
      P.src_o_off = o_off;
      P.src_m_off = m_off;
      P.uni.n = o_len[0]; P.uni.m = m_len[0];
      P.uni.dmin = family == 2 ? it.lo : 0; P.uni.dmax = family == 2 ? it.hi : 0;
      P.uni.s0 = family == 2 ? it.aux : 0;
      P.uni.words = it.words;
      P.uni.ppc = single_chunk ? ((int64_t)1 << 62) : ppc;
      P.uni.off_base = ob;
      return BSQ_SUCCESS;
    }
  }
  if (implicit) materialize_offsets();
  const bool plan_dbg = hk.plan_debug;
  const auto tp0 = std::chrono::steady_clock::now();
  auto lap = [&](const char* what) {
    if (plan_dbg)
      fprintf(stderr, "[bsq] make_plan %s at %.3f ms\n", what,
              std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - tp0).count());
  };

  // ---- fast path: ragged short reads (STD frame, one single-pass geometry for every pair, tables
  // within 2x of each other).  Pairs keep the caller's order -- the fill kernels pull them from a
  // queue, K1h pairs neighbours of any shape -- so there is no sort and results leave chunk by
  // chunk.  Two light passes over blocks of 4096 pairs: (A) per-block sums and checks, read-only;
  // a serial walk over the blocks cuts the chunks; (B) the pair records with their final offsets,
  // written once, sequentially.  (The general path below costs 3x the memory traffic.)
  const int64_t RB = hk.plan_block;
  if (family == 1 && !all_uniform && N >= 4 * RB && N < ((int64_t)1 << 31) &&
      !hk.always_sort) {
    bsq_params tmp = *prm;
    tmp.subst = P.subst.data();
    const bool i32ok = scores_int32_exact(&tmp, scan_n, scan_m);
    if (prm->dtype == BSQ_DTYPE_I32 && !i32ok)
      return fail(ctx, BSQ_ERR_ARG, "dtype I32 forced but scores are not int32-exact for these lengths");
    const int dtype = (prm->dtype == BSQ_DTYPE_F64) ? BSQ_DTYPE_F64 : (i32ok ? BSQ_DTYPE_I32 : BSQ_DTYPE_F64);
    bool range_ok = true;
    if (dtype == BSQ_DTYPE_F64) {
      double maxabs = std::fabs(prm->go) + std::fabs(prm->ge);
      for (double v : P.subst) maxabs = std::max(maxabs, std::fabs(v));
      range_ok = (double)(scan_n + scan_m + 4) * maxabs < 2147483647.0;
    }
    const int cfg = pick_k1_cfg((int)scan_n, (int)scan_m);   // geometry of the largest table
    const int G = BSQ_K1_CFGS[cfg].G, K = BSQ_K1_CFGS[cfg].K, BW = G * K, WPL = BSQ_K1_WPL(K);
    const int64_t NBLK = (N + RB - 1) / RB;
    struct Blk { int64_t words, len, cells, code_end; int32_t min_n, flags; };   // flags: 1 other geometry / passes, 2 K1h-eligible seen, 4 not eligible seen
    std::vector<Blk> blk((size_t)NBLK);
    const int TT = parallel_threads(N);
    auto over_blocks = [&](auto f) {
      if (TT <= 1) { f((int64_t)0, NBLK); return; }
      WorkerPool::get().run(TT, [&](int t) { f(NBLK * t / TT, NBLK * (t + 1) / TT); });
    };
    over_blocks([&](int64_t b0, int64_t b1) {
      for (int64_t b = b0; b < b1; b++) {
        Blk k{0, 0, 0, 0, INT32_MAX, 0};
        const int64_t i1 = std::min(N, (b + 1) * RB);
        for (int64_t i = b * RB; i < i1; i++) {
          const int n = o_len[i], m = m_len[i];
          if (m + BW >= 2 * BW || pick_k1_cfg(n, m) != cfg) k.flags |= 1;
          k.flags |= duo_shape_ok(cfg, 1, n, m) ? 2 : 4;
          k.words += (int64_t)(n + G) * G * WPL;
          k.len += (int64_t)n + m;
          k.cells += (int64_t)(n + 1) * (m + 1);
          k.code_end = std::max(k.code_end, ob + std::max(o_off[i] + n, m_off[i] + m));
          k.min_n = std::min(k.min_n, n);
        }
        blk[(size_t)b] = k;
      }
    });
    int flags = 0, min_n = INT32_MAX;
    int64_t tot_words = 0, cells = 0, max_blk = 0;
    for (const Blk& k : blk) { flags |= k.flags; min_n = std::min(min_n, k.min_n); tot_words += k.words; cells += k.cells; max_blk = std::max(max_blk, k.words); }
    const int64_t maxw = (int64_t)(scan_n + G) * G * WPL, minw = (int64_t)(min_n + G) * G * WPL;
    int64_t cap = arena_words;
    if (target_chunks > 1 && !single_chunk)
      cap = std::min(cap, std::max<int64_t>(std::max<int64_t>((tot_words + target_chunks - 1) / target_chunks, (int64_t)8 << 20), maxw));
    const bool duo_all = (flags & 6) == 2, duo_none = (flags & 2) == 0 || dtype != BSQ_DTYPE_I32;
    if (!(flags & 1) && range_ok && maxw <= 2 * minw && maxw <= BSQ_INORDER_MAX_WORDS && (duo_all || duo_none) && (single_chunk ? tot_words <= arena_words : max_blk <= cap)) {
      P.dtype = dtype;
      P.overrides.clear();
      P.k3.clear();
      P.k3_max_strips = P.k3_max_slab = P.k3_max_pairs = 0;
      P.identity = true;
      P.E = N;
      P.cells = cells;
      if (!hpairs.ensure(sizeof(BsqPair) * (size_t)N)) return fail(ctx, BSQ_ERR_NOMEM, "pinned host allocation failed");
      P.pairs = hpairs.as<BsqPair>();
      BsqPair* out = P.pairs;
      // chunks: whole blocks, as many as fit `cap` words; remember where each block's tables and
      // transcript slots start
      std::vector<int64_t> blk_ptr((size_t)NBLK), blk_slot((size_t)NBLK);
      P.chunks.clear();
      P.max_chunk_words = 0;
      Chunk cur{0, 0, family, cfg, 0, 0, 0, 0, 0};
      int64_t gw = 0, gl = 0;
      for (int64_t b = 0; b < NBLK; b++) {
        const Blk& k = blk[(size_t)b];
        if (!single_chunk && cur.count > 0 && cur.ptr_words + k.words > cap) {
          cur.idx_hi = b * RB;
          P.chunks.push_back(cur);
          cur = Chunk{b * RB, 0, family, cfg, 0, 0, b * RB, 0, 0};
        }
        blk_ptr[(size_t)b] = single_chunk ? gw : cur.ptr_words;
        blk_slot[(size_t)b] = gl;
        cur.ptr_words += k.words;
        cur.count += std::min(N, (b + 1) * RB) - b * RB;
        cur.code_end = std::max(cur.code_end, k.code_end);
        gw += k.words;
        gl += k.len;
        P.max_chunk_words = std::max(P.max_chunk_words, cur.ptr_words);
      }
      cur.idx_hi = N;
      P.chunks.push_back(cur);
      const bool duo = duo_all && dtype == BSQ_DTYPE_I32 && N >= hk.uniform_min;
      for (Chunk& ch : P.chunks) ch.duo = duo;
      // Inside a block the pairs are ordered by origin length (a counting sort of <= RB cached
      // entries): the lane groups of a warp -- and the two problems of a K1h duo -- then walk nearly
      // the same number of rows.  A chunk still covers a contiguous range of caller indices.
      const int nkeys = (int)scan_n - min_n + 1;
      over_blocks([&](int64_t b0, int64_t b1) {
        std::vector<int> cnt((size_t)nkeys + 1), perm((size_t)RB);
        for (int64_t b = b0; b < b1; b++) {
          int64_t pw = blk_ptr[(size_t)b], sl = blk_slot[(size_t)b];
          const int64_t ib = b * RB, i1 = std::min(N, (b + 1) * RB);
          std::fill(cnt.begin(), cnt.end(), 0);
          for (int64_t i = ib; i < i1; i++) cnt[(size_t)((int)scan_n - o_len[i]) + 1]++;
          for (int k = 0; k < nkeys; k++) cnt[(size_t)k + 1] += cnt[(size_t)k];
          for (int64_t i = ib; i < i1; i++) perm[(size_t)cnt[(size_t)((int)scan_n - o_len[i])]++] = (int)(i - ib);
          for (int64_t r = ib; r < i1; r++) {
            const int64_t i = ib + perm[(size_t)(r - ib)];
            const int n = o_len[i], m = m_len[i];
            BsqPair bp;
            bp.o_off = o_off[i] + ob;
            bp.m_off = m_off[i] + ob;
            bp.ptr_off = pw;
            bp.tx_slot = sl;
            bp.n = n;
            bp.m = m;
            bp.dmin = bp.dmax = bp.s0 = 0;
            bp.pad = (int)i;
            out[r] = bp;
            pw += (int64_t)(n + G) * G * WPL;
            sl += (int64_t)n + m;
          }
        }
      });
      P.total_words = tot_words;
      P.slot_bytes = gl;
      P.half_words = (!single_chunk && P.chunks.size() >= 2) ? P.max_chunk_words : 0;
      lap("in-order plan");
      return BSQ_SUCCESS;
    }
  }
  // ---- phase 1 (parallel): classify every pair like dptable_init would ----------
  std::vector<PlanItem>& items = ctx->plan_items;
  if ((int64_t)items.size() < N) items.resize((size_t)N);
  struct Part {
    int64_t max_n = 0, max_m = 0, cells = 0, tot_words = 0, maxw = 0, n_duo_ok = 0, n_invalid = 0;
    int64_t minw = INT64_MAX;
    int bad = 0;
    int cfg0 = -1, mixed = 0;   // one kernel geometry (and K1h eligibility) for all valid pairs?
  };
  std::vector<Part> parts(16);
  parallel_for(N, [&](int t, int64_t i0, int64_t i1) {
    Part pt;
    int pn = -1, pm = -1, plo = 0, phi = 0;
    PlanItem prev{};
    for (int64_t i = i0; i < i1; i++) {
      const int n = o_len[i], m = m_len[i];
      if (n < 0 || m < 0 || o_off[i] + ob < 0 || m_off[i] + ob < 0 || o_off[i] + ob + n > ncodes || m_off[i] + ob + m > ncodes) { pt.bad = 1; continue; }
      pt.max_n = std::max<int64_t>(pt.max_n, n);
      pt.max_m = std::max<int64_t>(pt.max_m, m);
      const int dlo = family == 2 ? dmin[i] : 0, dhi = family == 2 ? dmax[i] : 0;
      if (!(n == pn && m == pm && dlo == plo && dhi == phi)) {  // else: same shape as the previous pair
        prev = classify(n, m, dlo, dhi);
        pn = n; pm = m; plo = dlo; phi = dhi;
      }
      items[i] = prev;
      if (prev.cfg >= 0) {
        pt.cells += prev.cells;
        pt.tot_words += prev.words;
        pt.maxw = std::max(pt.maxw, prev.words);
        pt.minw = std::min(pt.minw, prev.words);
        const int tag = prev.cfg * 2 + ((family == 1 && prev.lo) ? 1 : 0);
        if (pt.cfg0 < 0) pt.cfg0 = tag;
        else if (pt.cfg0 != tag) pt.mixed = 1;
        if (family == 1 && prev.lo) pt.n_duo_ok++;
      } else {
        pt.n_invalid++;
      }
    }
    parts[t] = pt;
  });
  int64_t max_n = 0, max_m = 0, n_invalid = 0;
  Part all;
  P.cells = 0;
  for (const Part& pt : parts) {
    if (pt.bad) return fail(ctx, BSQ_ERR_ARG, "negative length or sequence range outside codes[]");
    max_n = std::max(max_n, pt.max_n); max_m = std::max(max_m, pt.max_m); P.cells += pt.cells;
    all.tot_words += pt.tot_words; all.maxw = std::max(all.maxw, pt.maxw); all.n_duo_ok += pt.n_duo_ok;
    all.minw = std::min(all.minw, pt.minw);
    if (pt.cfg0 >= 0) { if (all.cfg0 < 0) all.cfg0 = pt.cfg0; else if (all.cfg0 != pt.cfg0) all.mixed = 1; }
    all.mixed |= pt.mixed;
    n_invalid += pt.n_invalid;
  }

  bsq_params tmp = *prm;
  tmp.subst = P.subst.data();
  const bool i32ok = scores_int32_exact(&tmp, max_n, max_m);
  if (prm->dtype == BSQ_DTYPE_I32 && !i32ok)
    return fail(ctx, BSQ_ERR_ARG, "dtype I32 forced but scores are not int32-exact for these lengths");
  P.dtype = (prm->dtype == BSQ_DTYPE_F64) ? BSQ_DTYPE_F64 : (i32ok ? BSQ_DTYPE_I32 : BSQ_DTYPE_F64);
  bool range_ok = true;
  if (P.dtype == BSQ_DTYPE_F64) {
    // pwlib clamps gap candidates at -INT_MAX (_pw_internals.c:267); the kernels
    // do not, so refuse score ranges in which the clamp could ever bind.
    double maxabs = std::fabs(prm->go) + std::fabs(prm->ge);
    for (double v : P.subst) maxabs = std::max(maxabs, std::fabs(v));
    range_ok = (double)(max_n + max_m + 4) * maxabs < 2147483647.0;
  }

  lap("classified");
  // ---- phase 2 (serial, O(N) light): order, chunk boundaries, prefix sums -----------
  P.overrides.clear();
  std::vector<int64_t>& order = ctx->plan_order;  // valid pairs in caller order (only built if some are not)
  order.clear();
  bool uniform = true;
  int64_t first_words = -1, tot_words = 0, maxw = 0, n_duo_ok = 0;
  int first_cfg = -1;
  const bool have_order = !(range_ok && n_invalid == 0);   // every pair valid: position e IS caller index e
  if (have_order) {
    order.reserve((size_t)N);
    for (int64_t i = 0; i < N; i++) {
      const PlanItem& it = items[i];
      if (!range_ok) { P.overrides.push_back({i, BSQ_UNSUPPORTED}); continue; }
      if (it.cfg < 0) { P.overrides.push_back({i, -1 - it.cfg}); continue; }
      if (family == 1) n_duo_ok += it.lo;
      if (first_words < 0) { first_words = it.words; first_cfg = it.cfg; }
      else if (it.words != first_words || it.cfg != first_cfg) uniform = false;
      tot_words += it.words;
      maxw = std::max(maxw, it.words);
      order.push_back(i);
    }
  } else {
    tot_words = all.tot_words; maxw = all.maxw; n_duo_ok = all.n_duo_ok;
    // Caller order is kept (no sort, results leave chunk by chunk) for ragged SHORT reads: STD
    // frame, one kernel geometry, tables within 2x of each other and small.  The fill kernels pull
    // pairs from a queue and K1h pairs neighbours of any shape, so the order costs ~10 % of the fill
    // (lane groups of one warp run different row counts) where the sort would cost more than that.
    // Long pairs are always ordered: measured on config 3 (32768 banded pairs of ~10 kb) caller
    // order made the fill 26 % slower, and their plan is < 3 % of the call either way.
    uniform = family == 1 && !all.mixed && all.cfg0 >= 0 && all.maxw <= 2 * all.minw &&
              all.maxw <= BSQ_INORDER_MAX_WORDS && !hk.always_sort;
  }
  if (!range_ok) P.cells = 0;
  P.identity = uniform;
  if (target_chunks > 1 && !single_chunk) {  // cut evenly so that copies overlap compute
    int64_t per = std::max<int64_t>((tot_words + target_chunks - 1) / target_chunks, (int64_t)8 << 20);
    per = std::max(per, maxw);
    arena_words = std::min(arena_words, per);
  }
  if (maxw > arena_words) return fail(ctx, BSQ_ERR_NOMEM, "a single pair's traceback table exceeds the device arena");
  if (single_chunk && tot_words > arena_words)
    return fail(ctx, BSQ_ERR_NOMEM, "batch does not fit the device arena in two-phase mode; use bsq_align_batch");
  const int64_t E = have_order ? (int64_t)order.size() : N;
  P.E = E;
  if (!hpairs.ensure(sizeof(BsqPair) * (size_t)std::max<int64_t>(E, 1))) return fail(ctx, BSQ_ERR_NOMEM, "pinned host allocation failed");
  P.pairs = hpairs.as<BsqPair>();
  BsqPair* out = P.pairs;
  // what the chunk-cutting scan below reads, in processing order
  struct Gathered { int64_t words, code_end; int32_t len, cfg, aux, band_w, duo, pad_; };
  std::vector<char>& gbuf = ctx->plan_gather;
  if (gbuf.size() < sizeof(Gathered) * (size_t)std::max<int64_t>(E, 1)) gbuf.resize(sizeof(Gathered) * (size_t)std::max<int64_t>(E, 1));
  Gathered* gth = (Gathered*)gbuf.data();
  // (words, geometry, duo flag) of every pair in processing order, 8 bytes each: all the serial
  // chunk-cutting scan has to read
  std::vector<int64_t>& ckbuf = ctx->plan_sorted;
  if ((int64_t)ckbuf.size() < E + 1) ckbuf.resize((size_t)E + 1);
  uint64_t* ckey = (uint64_t*)ckbuf.data();

  // Processing order: by kernel geometry, heaviest first inside a geometry (LPT).  A stable counting
  // sort on (geometry, size class) -- 8 classes per octave of table size, so pairs of one class
  // differ by < 12.5 % -- whose scatter pass writes the pair records (minus the two offsets the
  // scan assigns) straight to their final place in pinned memory: sequential reads of the caller's
  // arrays, one random 56-byte write per pair.  (A comparison sort plus gathers through the
  // permutation cost 110 ms per 10^6 ragged pairs, more than aligning them.)
  // Short reads that qualify for K1h (PlanItem.lo) are keyed ahead of the rest, so they form their
  // own chunks, in which pairs (2q, 2q+1) share a lane group: neighbours of one size class, i.e.
  // the longer problem of a duo has < 12.5 % more rows than the shorter.  Worth it from a few
  // thousand candidates on.
  const bool use_duo = family == 1 && P.dtype == BSQ_DTYPE_I32 && n_duo_ok >= hk.uniform_min &&
                       !hk.no_duo;
  constexpr int NCLS = 512;
  const int nkeys_std = (BSQ_K2_NCFG + BSQ_K3_NCFG + 1) * NCLS;
  const int nkeys_duo = use_duo ? 2 * NCLS : 0;   // K1h runs on K1 geometries 0 and 1
  const int nkeys = nkeys_std + nkeys_duo;
  const int TMAX = parallel_threads(E);   // the ranges parallel_for(E, ..) hands out below
  std::vector<int64_t> hist;
  auto key_of = [&](const PlanItem& it) -> int {
    const uint64_t w = (uint64_t)std::max<int64_t>(it.words, 1);
    const int lg = 63 - __builtin_clzll(w);
    const int frac = lg >= 3 ? (int)((w >> (lg - 3)) & 7) : 0;
    return ((use_duo && it.lo) ? 0 : nkeys_duo) + it.cfg * NCLS + (NCLS - 1 - (lg * 8 + frac));
  };
  if (!uniform) {
    hist.assign((size_t)TMAX * nkeys, 0);
    parallel_for(E, [&](int t, int64_t e0, int64_t e1) {
      int64_t* h = hist.data() + (size_t)t * nkeys;
      for (int64_t e = e0; e < e1; e++) { const int64_t i = have_order ? order[e] : e; h[key_of(items[i])]++; }
    });
    int64_t run = 0;  // exclusive prefix over (key, thread): thread t's pairs of a key go behind thread t-1's
    for (int k = 0; k < nkeys; k++)
      for (int t = 0; t < TMAX; t++) { int64_t& h = hist[(size_t)t * nkeys + k]; const int64_t c = h; h = run; run += c; }
  }
  parallel_for(E, [&](int t, int64_t e0, int64_t e1) {
    int64_t* h = uniform ? nullptr : hist.data() + (size_t)t * nkeys;
    for (int64_t e = e0; e < e1; e++) {
      const int64_t i = have_order ? order[e] : e;
      const PlanItem& it = items[i];
      const int64_t pos = uniform ? e : h[key_of(it)]++;
      BsqPair bp;
      bp.o_off = o_off[i] + ob;
      bp.m_off = m_off[i] + ob;
      bp.ptr_off = 0;
      bp.tx_slot = 0;
      bp.n = o_len[i];
      bp.m = m_len[i];
      bp.dmin = family == 2 ? it.lo : 0;
      bp.dmax = family == 2 ? it.hi : 0;
      bp.s0 = family == 2 ? it.aux : 0;
      bp.pad = (int)i;  // original index (N < 2^31 enforced by caller)
      out[pos] = bp;
      Gathered gq;
      gq.words = it.words;
      gq.code_end = ob + std::max(o_off[i] + o_len[i], m_off[i] + m_len[i]);
      gq.len = o_len[i] + m_len[i];
      gq.cfg = it.cfg;
      gq.aux = it.aux;
      gq.band_w = it.hi - it.lo + 1;
      gq.duo = (use_duo && it.lo) ? 1 : 0;
      gq.pad_ = 0;
      gth[pos] = gq;
      ckey[pos] = (uint64_t)it.words | ((uint64_t)(it.cfg & 0xff) << 48) | ((uint64_t)gq.duo << 56);
    }
  });
  lap("ordered");

  // ---- the chunk-cutting scan (serial, sequential memory): pointer-table and transcript-slot
  // offsets, chunk boundaries
  P.chunks.clear();
  P.max_chunk_words = 0;
  Chunk cur{0, 0, family, -1, 0, 0, 0, 0, 0};
  int64_t global_words = 0, slot = 0;
  P.k3.clear();
  P.k3_max_strips = P.k3_max_slab = P.k3_max_pairs = 0;
  auto close_chunk = [&](Chunk& c) {
    P.k3_max_strips = std::max(P.k3_max_strips, c.k3_strips);
    P.k3_max_slab = std::max(P.k3_max_slab, c.k3_slab);
    if (chunk_is_k3(c)) P.k3_max_pairs = std::max(P.k3_max_pairs, c.count);
    P.chunks.push_back(c);
  };
  if (family == 2) {
    bool any_k3 = false;
    for (int64_t e = 0; e < E && !any_k3; e++) any_k3 = gth[e].cfg >= BSQ_K2_NCFG;
    if (any_k3) P.k3.assign((size_t)E, BsqK3Pair{0, 0, 0});
  }
  lap("scan starts");
  const bool serial_scan = !P.k3.empty() || E < 2 * parallel_grain() || hk.serial_scan;
  if (serial_scan) {
  for (int64_t e = 0; e < E; e++) {
    const Gathered& it = gth[e];
    const bool new_chunk = cur.count > 0 && (it.cfg != cur.cfg || (it.duo != 0) != cur.duo ||
                                             (!single_chunk && cur.ptr_words + it.words > arena_words));
    if (new_chunk) {
      cur.idx_hi = out[e].pad;
      close_chunk(cur);
      cur = Chunk{e, 0, family, -1, 0, 0, out[e].pad, 0, 0};
    }
    if (cur.count == 0) { cur.cfg = it.cfg; cur.duo = it.duo != 0; }
    out[e].ptr_off = single_chunk ? global_words : cur.ptr_words;
    out[e].tx_slot = slot;
    slot += it.len;
    if (family == 1) cur.max_n_multipass = std::max(cur.max_n_multipass, it.aux);
    cur.code_end = std::max<int64_t>(cur.code_end, it.code_end);
    if (family == 2 && it.cfg >= BSQ_K2_NCFG) {
      const int BW = 32 * BSQ_K3_CFGS[it.cfg - BSQ_K2_NCFG].K;
      const int64_t nstrips = ((int64_t)out[e].m + BW) / BW;
      if (cur.k3_strips + nstrips >= ((int64_t)1 << 30))
        return fail(ctx, BSQ_ERR_ARG, "too many wide-band strips in one chunk");
      P.k3[(size_t)e] = BsqK3Pair{cur.k3_slab, (int)cur.k3_strips, (int)nstrips};
      cur.k3_strips += nstrips;
      cur.k3_slab += (int64_t)out[e].n + 1;
      cur.k3_active += std::min<int64_t>(nstrips, ((int64_t)it.band_w + BW + 32) / BW + 2);
    }
    cur.ptr_words += it.words;
    global_words += it.words;
    cur.count++;
    P.max_chunk_words = std::max(P.max_chunk_words, cur.ptr_words);
  }
  } else {
    // (1) serial, over the 8-byte keys only: where the chunks begin
    std::vector<int64_t> chunk_gw;   // pointer-table words of all pairs before the chunk
    for (int64_t e = 0; e < E; e++) {
      const uint64_t ck = ckey[e];
      const int64_t words = (int64_t)(ck & 0xffffffffffffull);
      const int cfg = (int)((ck >> 48) & 0xff);
      const bool duo = (ck >> 56) != 0;
      const bool new_chunk = cur.count > 0 && (cfg != cur.cfg || duo != cur.duo ||
                                               (!single_chunk && cur.ptr_words + words > arena_words));
      if (new_chunk) {
        cur.idx_hi = out[e].pad;
        close_chunk(cur);
        cur = Chunk{e, 0, family, -1, 0, 0, out[e].pad, 0, 0};
      }
      if (cur.count == 0) { cur.cfg = cfg; cur.duo = duo; chunk_gw.push_back(global_words); }
      cur.ptr_words += words;
      global_words += words;
      cur.count++;
      P.max_chunk_words = std::max(P.max_chunk_words, cur.ptr_words);
    }
    std::vector<Chunk> done = P.chunks;   // closed so far; the open one is appended for the passes below
    done.push_back(cur);
    const int64_t NCH = (int64_t)done.size();
    // (2) parallel: per-range sums, then every pair's offsets and the per-chunk maxima
    const int TT = parallel_threads(E);
    std::vector<int64_t> sumW((size_t)TT + 1, 0), sumL((size_t)TT + 1, 0);
    parallel_for(E, [&](int t, int64_t e0, int64_t e1) {
      int64_t w = 0, l = 0;
      for (int64_t e = e0; e < e1; e++) { w += (int64_t)(ckey[e] & 0xffffffffffffull); l += gth[e].len; }
      sumW[(size_t)t + 1] = w;
      sumL[(size_t)t + 1] = l;
    });
    for (int t = 0; t < TT; t++) { sumW[(size_t)t + 1] += sumW[(size_t)t]; sumL[(size_t)t + 1] += sumL[(size_t)t]; }
    std::vector<int64_t> mx_code((size_t)TT * NCH, 0);
    std::vector<int> mx_aux((size_t)TT * NCH, 0);
    parallel_for(E, [&](int t, int64_t e0, int64_t e1) {
      int64_t gw = sumW[(size_t)t], gl = sumL[(size_t)t];
      int64_t c = (int64_t)(std::upper_bound(done.begin(), done.end(), e0, [](int64_t e, const Chunk& ch) { return e < ch.first; }) - done.begin()) - 1;
      int64_t* mc = mx_code.data() + (size_t)t * NCH;
      int* ma = mx_aux.data() + (size_t)t * NCH;
      for (int64_t e = e0; e < e1; e++) {
        while (c + 1 < NCH && e >= done[(size_t)c + 1].first) c++;
        const Gathered& it = gth[e];
        out[e].ptr_off = single_chunk ? gw : gw - chunk_gw[(size_t)c];
        out[e].tx_slot = gl;
        gw += it.words;
        gl += it.len;
        mc[c] = std::max(mc[c], it.code_end);
        if (family == 1) ma[c] = std::max(ma[c], it.aux);
      }
    });
    slot = sumL[(size_t)TT];
    for (int64_t c = 0; c < NCH; c++) {
      Chunk& ch = (c + 1 < NCH) ? P.chunks[(size_t)c] : cur;
      for (int t = 0; t < TT; t++) {
        ch.code_end = std::max(ch.code_end, mx_code[(size_t)t * NCH + c]);
        ch.max_n_multipass = std::max(ch.max_n_multipass, mx_aux[(size_t)t * NCH + c]);
      }
    }
  }
  cur.idx_hi = N;
  if (cur.count > 0) close_chunk(cur);
  // K3 work lists: strip j of every pair of the chunk before strip j+1 of any, so that the strip
  // pipelines of all its pairs run side by side (claimed in list order: a strip's predecessor is
  // always taken before it).  Pairs by falling strip count: level j is a prefix of that order.
  P.k3order.clear();
  if (!P.k3.empty()) {
    std::vector<int> by_strips;
    for (Chunk& ch : P.chunks) {
      if (!chunk_is_k3(ch)) continue;
      ch.k3_order_off = (int64_t)P.k3order.size();
      by_strips.resize((size_t)ch.count);
      for (int64_t i = 0; i < ch.count; i++) by_strips[(size_t)i] = (int)i;
      const BsqK3Pair* kp = P.k3.data() + ch.first;
      std::stable_sort(by_strips.begin(), by_strips.end(), [&](int a, int b) { return kp[a].nstrips > kp[b].nstrips; });
      size_t live = by_strips.size();
      for (int j = 0; live > 0; j++) {
        while (live > 0 && kp[by_strips[live - 1]].nstrips <= j) live--;
        for (size_t i = 0; i < live; i++) P.k3order.push_back(BsqK3Item{by_strips[i], j});
      }
    }
  }
  if (!P.chunks.empty()) P.chunks.front().idx_lo = 0;
  P.total_words = global_words;
  P.slot_bytes = slot;
  P.half_words = (!single_chunk && P.chunks.size() >= 2) ? P.max_chunk_words : 0;
  lap("chunked");
  return BSQ_SUCCESS;
}

cudaEvent_t get_event(bsq_ctx* ctx, size_t i) {
  while (ctx->events.size() <= i) {
    cudaEvent_t e;
    cudaEventCreate(&e);
    ctx->events.push_back(e);
  }
  return ctx->events[i];
}

// HBM held by this context outside the buffer set `mine` (which the caller is about to reuse)
size_t held_elsewhere(const bsq_ctx* ctx, const BufSet* mine) {
  size_t n = 0;
  if (&ctx->bufs != mine) n += ctx->bufs.bytes();
  for (const BufSet* p : ctx->pool) if (p != mine) n += p->bytes();
  for (const bsq_batch* b : ctx->live) if (b->own && b->own != mine) n += b->own->bytes();
  return n;
}

// words the pointer arena of a plan may take: 85 % of what is free for the buffer set `mine`,
// minus the plan's other buffers.  cudaMemGetInfo costs 0.1 - 50 ms (measured on the bench box),
// so the context remembers "free + everything we hold" and asks again only after a failed
// allocation (callers then retry once).
int64_t arena_budget_words(bsq_ctx* ctx, int64_t other_bytes, const BufSet* mine) {
  if (ctx->mem_avail < 0.0) {
    size_t fr = 0, tot = 0;
    if (cudaMemGetInfo(&fr, &tot) != cudaSuccess) { fr = (size_t)1 << 30; cudaGetLastError(); }
    ctx->mem_avail = (double)fr + (double)held_elsewhere(ctx, nullptr);
  }
  double budget = (ctx->mem_avail - (double)held_elsewhere(ctx, mine)) * 0.85 - (double)other_bytes;
  if (ctx->hooks.arena_cap_bytes && budget > (double)ctx->hooks.arena_cap_bytes) budget = (double)ctx->hooks.arena_cap_bytes;
  if (budget < (double)(1 << 16)) budget = (double)(1 << 16);
  return (int64_t)(budget / 4.0);
}

void drop_pool(bsq_ctx* ctx) {
  for (BufSet* p : ctx->pool) { p->release(); delete p; }
  ctx->pool.clear();
}

// pair records: a plain H2D copy, or (uniform batches) generated on the device from the
// caller's offset arrays
int upload_pairs(bsq_ctx* ctx, BufSet& B, const Plan& P, cudaStream_t st) {
  const int64_t E = P.E;
  if (E <= 0) return BSQ_SUCCESS;
  if (!P.uniform_dev) {
    BSQ_CK(ctx, cudaMemcpyAsync(B.pairs.p, P.pairs, sizeof(BsqPair) * (size_t)E, cudaMemcpyHostToDevice, st));
    ctx->tm.h2d_bytes += (int64_t)sizeof(BsqPair) * E;
    return BSQ_SUCCESS;
  }
  const bool have_off = P.src_o_off != nullptr;   // else: back-to-back pairs, offsets are arithmetic
  if (have_off) {
    if (!B.ooff.ensure(8 * (size_t)E) || !B.moff.ensure(8 * (size_t)E)) return fail(ctx, BSQ_ERR_NOMEM, "device allocation failed (offsets)");
    BSQ_CK(ctx, cudaMemcpyAsync(B.ooff.p, P.src_o_off, 8 * (size_t)E, cudaMemcpyHostToDevice, st));
    BSQ_CK(ctx, cudaMemcpyAsync(B.moff.p, P.src_m_off, 8 * (size_t)E, cudaMemcpyHostToDevice, st));
    ctx->tm.h2d_bytes += 16 * E;
  }
  const int grid = (int)((E + 255) / 256);
  BSQ_LAUNCH(bsq_make_pairs_uniform, grid, 256, 0, st, have_off ? (const long long*)B.ooff.as<long long>() : (const long long*)nullptr,
             have_off ? (const long long*)B.moff.as<long long>() : (const long long*)nullptr, (long long)E, P.uni, B.pairs.as<BsqPair>());
  BSQ_CK(ctx, cudaGetLastError());
  ctx->tm.fill_launches++;
  return BSQ_SUCCESS;
}

// bytes of codes[] as the caller holds them (one letter per byte, or four: BSQ_F_CODES_2BIT)
inline int64_t codes_in_bytes(const bsq_params& prm, int64_t ncodes) {
  return (prm.flags & BSQ_F_CODES_2BIT) ? (ncodes + 3) / 4 : ncodes;
}

// the thread-per-pair banded walker reads letters out of a 2-bit packed copy of codes[]
inline bool want_packed_copy(const bsq_params& prm) { return prm.alphabet_len <= 4 && prm.mode == BSQ_BANDED_MODE; }

// H2D of bytes [b0, b1) of the caller's codes[] on stream st; packed input is expanded to one byte
// per letter right behind the copy (same stream), so that every kernel reads B.codes
int upload_codes(bsq_ctx* ctx, BufSet& B, const bsq_params& prm, const uint8_t* codes, int64_t b0, int64_t b1, cudaStream_t st) {
  if (b1 <= b0) return BSQ_SUCCESS;
  if (!(prm.flags & BSQ_F_CODES_2BIT)) {
    BSQ_CK(ctx, cudaMemcpyAsync((char*)B.codes.p + b0, codes + b0, (size_t)(b1 - b0), cudaMemcpyHostToDevice, st));
    if (want_packed_copy(prm)) {   // letters [b0, b1) -> words [b0 / 16, ceil(b1 / 16)); slices start on 256-byte boundaries
      const long long w_lo = b0 / 16, w_hi = (b1 + 15) / 16;
      const int grid = (int)((w_hi - w_lo + 255) / 256);
      BSQ_LAUNCH(bsq_pack_2bit, grid, 256, 0, st, (const uint8_t*)B.codes.as<uint8_t>(), (long long)b1, w_lo, w_hi, B.codes2.as<uint32_t>());
      BSQ_CK(ctx, cudaGetLastError());
      ctx->tm.fill_launches++;
    }
    return BSQ_SUCCESS;
  }
  BSQ_CK(ctx, cudaMemcpyAsync((char*)B.codes2.p + b0, codes + b0, (size_t)(b1 - b0), cudaMemcpyHostToDevice, st));
  const long long nb = b1 - b0;
  const int grid = (int)((nb / 4 + 1 + 255) / 256);
  BSQ_LAUNCH(bsq_unpack_2bit, grid, 256, 0, st, (const uint8_t*)(B.codes2.as<uint8_t>() + b0), nb, B.codes.as<uint8_t>() + 4 * b0);
  BSQ_CK(ctx, cudaGetLastError());
  ctx->tm.fill_launches++;
  return BSQ_SUCCESS;
}
inline bool ensure_codes(BufSet& B, const bsq_params& prm, int64_t ncodes) {
  if (!B.codes.ensure((size_t)ncodes + 64)) return false;
  if (!(prm.flags & BSQ_F_CODES_2BIT) && !want_packed_copy(prm)) return true;
  return B.codes2.ensure((size_t)(ncodes + 3) / 4 + 64);
}
// packed input of a 4-letter alphabet cannot hold a letter >= alphabet_len
inline bool codes_need_check(const bsq_params& prm) { return !((prm.flags & BSQ_F_CODES_2BIT) && prm.alphabet_len >= 4); }

// every letter referenced by pair records [first, first + count) must be < alphabet_len; raises
// B.flags[0] on the device (the fill / traceback kernels then return at once)
int launch_check_codes(bsq_ctx* ctx, BufSet& B, const Plan& P, int64_t first, int64_t count, cudaStream_t st) {
  if (count <= 0 || !codes_need_check(P.prm)) return BSQ_SUCCESS;
  const int grid = (int)((count * 32 + 127) / 128);
  BSQ_LAUNCH(bsq_check_codes, grid, 128, 0, st, (const uint8_t*)B.codes.as<uint8_t>(),
             (const BsqPair*)(B.pairs.as<BsqPair>() + first), (int)count, P.prm.alphabet_len, B.flags.as<int>());
  BSQ_CK(ctx, cudaGetLastError());
  ctx->tm.fill_launches++;
  return BSQ_SUCCESS;
}

// K3 chunks: per-pair strip records, progress/done words, per-strip results, slabs
int stage_k3(bsq_ctx* ctx, BufSet& B, const Plan& P, cudaStream_t st) {
  if (P.k3.empty()) return BSQ_SUCCESS;
  const bool ok = B.k3pairs.ensure(sizeof(BsqK3Pair) * P.k3.size()) &&
                  B.k3sync.ensure(4 * (size_t)(P.k3_max_pairs + 1)) &&
                  B.k3res.ensure(sizeof(BsqK3Res) * (size_t)(P.k3_max_strips + 1)) &&
                  B.k3order.ensure(sizeof(BsqK3Item) * (P.k3order.size() + 1)) &&
                  B.colh.ensure(16 * (size_t)(P.k3_max_slab + 1));
  if (!ok) return fail(ctx, BSQ_ERR_NOMEM, "device allocation failed (wide-band strips)");
  BSQ_CK(ctx, cudaMemcpyAsync(B.k3pairs.p, P.k3.data(), sizeof(BsqK3Pair) * P.k3.size(), cudaMemcpyHostToDevice, st));
  BSQ_CK(ctx, cudaMemcpyAsync(B.k3order.p, P.k3order.data(), sizeof(BsqK3Item) * P.k3order.size(), cudaMemcpyHostToDevice, st));
  ctx->tm.h2d_bytes += (int64_t)(sizeof(BsqK3Pair) * P.k3.size() + sizeof(BsqK3Item) * P.k3order.size());
  return BSQ_SUCCESS;
}

// upload everything the kernels read; size the output arrays
int stage_inputs(bsq_ctx* ctx, BufSet& B, Plan& P, const uint8_t* codes, int64_t ncodes, bool single_chunk) {
  const int64_t N = P.N, E = (int64_t)P.E;
  const int L = P.prm.alphabet_len;
  bool ok = ensure_codes(B, P.prm, ncodes) && B.pairs.ensure(sizeof(BsqPair) * (size_t)std::max<int64_t>(E, 1)) &&
            B.score.ensure(8 * (size_t)N) && B.ex.ensure(4 * (size_t)N) && B.ey.ensure(4 * (size_t)N) &&
            B.sx.ensure(4 * (size_t)N) && B.sy.ensure(4 * (size_t)N) && B.status.ensure(4 * (size_t)N) &&
            B.txlen.ensure(4 * (size_t)N) && B.subst.ensure(8 * (size_t)L * L) &&
            B.counters.ensure(4 * (P.chunks.size() + 1)) && B.total.ensure(16) && B.flags.ensure(16) &&
            B.ptr.ensure(4 * (size_t)(single_chunk ? P.total_words : (P.half_words > 0 ? BSQ_ARENA_SLOTS * P.half_words : P.max_chunk_words)) + 64);
  if (!ok) { ctx->mem_avail = -1.0; return fail(ctx, BSQ_ERR_NOMEM, "device allocation failed"); }
  cudaStream_t st = ctx->st;
  {
    const int rcu = upload_codes(ctx, B, P.prm, codes, 0, codes_in_bytes(P.prm, ncodes), st);
    if (rcu != BSQ_SUCCESS) return rcu;
    P.codes2_ok = (P.prm.flags & BSQ_F_CODES_2BIT) || want_packed_copy(P.prm);
  }
  ctx->tm.h2d_bytes += codes_in_bytes(P.prm, ncodes);
  {
    const int rcp = upload_pairs(ctx, B, P, st);
    if (rcp != BSQ_SUCCESS) return rcp;
  }
  if (P.dtype == BSQ_DTYPE_I32) {
    std::vector<int> si((size_t)L * L);
    for (int i = 0; i < L * L; i++) si[i] = (int)P.subst[i];
    BSQ_CK(ctx, cudaMemcpyAsync(B.subst.p, si.data(), 4 * (size_t)L * L, cudaMemcpyHostToDevice, st));
    BSQ_CK(ctx, cudaStreamSynchronize(st));  // si is a temporary
  } else {
    BSQ_CK(ctx, cudaMemcpyAsync(B.subst.p, P.subst.data(), 8 * (size_t)L * L, cudaMemcpyHostToDevice, st));
  }
  BSQ_CK(ctx, cudaMemsetAsync(B.counters.p, 0, 4 * (P.chunks.size() + 1), st));
  BSQ_CK(ctx, cudaMemsetAsync(B.txlen.p, 0, 4 * (size_t)N, st));
  BSQ_CK(ctx, cudaMemsetAsync(B.flags.p, 0, 16, st));
  {
    const int rck = launch_check_codes(ctx, B, P, 0, E, st);
    if (rck != BSQ_SUCCESS) return rck;
  }
  return stage_k3(ctx, B, P, st);
}

int launch_fill(bsq_ctx* ctx, BufSet& B, const Plan& P, size_t ci, size_t* evi, cudaStream_t stream, bool leave_room) {
  const Chunk& ch = P.chunks[ci];
  BsqFillArgs A;
  memset(&A, 0, sizeof A);
  A.codes = B.codes.as<uint8_t>();
  A.pairs = B.pairs.as<BsqPair>() + ch.first;
  A.npairs = (int)ch.count;
  A.L = P.prm.alphabet_len;
  A.subst = B.subst.p;
  A.go_d = P.prm.go; A.ge_d = P.prm.ge;
  A.go_i = (int)P.prm.go; A.ge_i = (int)P.prm.ge;
  A.bmode = P.bmode; A.emode = P.emode; A.btype = P.prm.type;
  A.ptr = B.ptr.as<uint32_t>() + (P.half_words > 0 ? (int64_t)(ci % BSQ_ARENA_SLOTS) * P.half_words : 0);
  A.out_score = B.score.as<double>();
  A.out_ex = B.ex.as<int>();
  A.out_ey = B.ey.as<int>();
  A.out_status = B.status.as<int>();
  A.counter = B.counters.as<int>() + ci;
  A.bad = B.flags.as<int>();
  if (chunk_is_k3(ch)) {
    A.k3 = B.k3pairs.as<BsqK3Pair>() + ch.first;
    A.k3_order = B.k3order.as<BsqK3Item>() + ch.k3_order_off;
    A.k3_total = (int)ch.k3_strips;
    A.k3_done = B.k3sync.as<int>();
    A.k3_res = B.k3res.as<BsqK3Res>();
    A.col_h = B.colh.as<double>();
    BSQ_CK(ctx, cudaMemsetAsync(B.k3sync.p, 0, 4 * (size_t)ch.count, stream));
    BSQ_CK(ctx, cudaMemsetAsync(B.colh.p, 0, (P.dtype == BSQ_DTYPE_I32 ? 8 : 16) * (size_t)ch.k3_slab, stream));
    // workers: twice the strips in flight with full pipelines (the rest would only poll), at
    // most 16 warps per SM; any number is deadlock-free because strips are claimed in order
    // K3 blocks hold two independent strip workers, K3L blocks one worker and its feeder warp
    const int wpb = (ch.cfg - BSQ_K2_NCFG == 0) ? 1 : BSQ_K3_THREADS / 32;
    int64_t workers = std::min<int64_t>(ch.k3_strips, std::max<int64_t>(2 * ch.k3_active, 64));
    workers = std::min<int64_t>(workers, (int64_t)ctx->sms * 16);
    const int grid = (int)std::max<int64_t>((workers + wpb - 1) / wpb, 1);
    cudaEvent_t e0 = get_event(ctx, (*evi)++), e1 = get_event(ctx, (*evi)++);
    BSQ_CK(ctx, cudaEventRecord(e0, stream));
    if (k3_op(P.dtype, ch.cfg - BSQ_K2_NCFG, P.prm.type, P.gosign, grid, stream, &A) < 0)
      return fail(ctx, BSQ_ERR_ARG, "no kernel variant for this configuration");
    BSQ_CK(ctx, cudaGetLastError());
    BSQ_CK(ctx, cudaEventRecord(e1, stream));
    ctx->tm.fill_launches++;
    ctx->tm.ptr_bytes += 4 * ch.ptr_words;
    return BSQ_SUCCESS;
  }
  if (ch.duo) {
    int occ = bsq_k1h_op(ch.cfg, P.gosign, P.uniform_dev ? 1 : 0, BSQ_OP_OCCUPANCY, 0, 0, nullptr);
    if (occ < 0) return fail(ctx, BSQ_ERR_ARG, "no kernel variant for this configuration");
    const int gpb = BSQ_FILL_THREADS / BSQ_K1_CFGS[ch.cfg].G;
    const int64_t duos = (ch.count + 1) / 2;
    A.nduos = (int)duos;
    const int occ_use = (leave_room && occ >= 4) ? occ - 1 : occ;
    int grid = (int)std::min<int64_t>((duos + gpb - 1) / gpb, (int64_t)ctx->sms * occ_use);
    if (grid < 1) grid = 1;
    cudaEvent_t e0 = get_event(ctx, (*evi)++), e1 = get_event(ctx, (*evi)++);
    BSQ_CK(ctx, cudaEventRecord(e0, stream));
    if (bsq_k1h_op(ch.cfg, P.gosign, P.uniform_dev ? 1 : 0, BSQ_OP_LAUNCH, grid, stream, &A) < 0)
      return fail(ctx, BSQ_ERR_ARG, "no kernel variant for this configuration");
    BSQ_CK(ctx, cudaGetLastError());
    BSQ_CK(ctx, cudaEventRecord(e1, stream));
    ctx->tm.fill_launches++;
    ctx->tm.ptr_bytes += 4 * ch.ptr_words;
    ctx->tm.dtype = BSQ_DTYPE_I16X2;
    return BSQ_SUCCESS;
  }
  const int G = ch.family == 1 ? BSQ_K1_CFGS[ch.cfg].G : BSQ_K2_CFGS[ch.cfg].G;
  // per-cell end search: ungated when every gap move lowers (or keeps) the score, so that
  // the -inf padding columns can never beat a real cell; explicitly gated otherwise
  const bool monotone = (P.prm.ge <= 0.0) && (P.prm.ge + P.prm.go <= 0.0);
  const int ball = P.bmode == BSQ_B_ALL, eall = (P.emode == BSQ_E_ALL) ? (monotone ? 1 : 2) : 0;
  // B_LOCAL: 2 = packed end-cell keys (integer scores below 2^16, n + m below 2^16), 1 = general
  int blocal = P.prm.type == BSQ_B_LOCAL;
  if (blocal && ch.family == 2 && P.dtype == BSQ_DTYPE_I32 && P.max_nm < 65000 && P.score_bound < 65000.0) blocal = 2;
  int occ = ch.family == 1 ? k1_op(P.dtype, ch.cfg, ball, eall, P.gosign, BSQ_OP_OCCUPANCY, 0, 0, nullptr)
                           : k2_op(P.dtype, ch.cfg, blocal, P.gosign, BSQ_OP_OCCUPANCY, 0, 0, nullptr);
  if (occ < 0) return fail(ctx, BSQ_ERR_ARG, "no kernel variant for this configuration");
  const int groups_per_block = BSQ_FILL_THREADS / G;
  int64_t want_blocks = (ch.count + groups_per_block - 1) / groups_per_block;
  // when a traceback kernel is to run beside this fill, leave one block's worth of
  // registers per SM free (the fill is ALU-bound well below full occupancy)
  const int occ_use = (leave_room && occ >= 4) ? occ - 1 : occ;
  int grid = (int)std::min<int64_t>(want_blocks, (int64_t)ctx->sms * occ_use);
  if (grid < 1) grid = 1;
  if (ch.family == 1 && ch.max_n_multipass > 0) {
    const int64_t stride = (int64_t)ch.max_n_multipass + 2;
    const int64_t groups = (int64_t)grid * groups_per_block;
    if (!B.colh.ensure(8 * (size_t)(stride * groups)) || !B.colf.ensure((size_t)(stride * groups)))
      return fail(ctx, BSQ_ERR_NOMEM, "device allocation failed (boundary columns)");
    A.col_h = B.colh.as<double>();
    A.col_f = B.colf.as<uint8_t>();
    A.col_stride = stride;
  }
  cudaEvent_t e0 = get_event(ctx, (*evi)++), e1 = get_event(ctx, (*evi)++);
  BSQ_CK(ctx, cudaEventRecord(e0, stream));
  int rc = ch.family == 1 ? k1_op(P.dtype, ch.cfg, ball, eall, P.gosign, BSQ_OP_LAUNCH, grid, stream, &A)
                          : k2_op(P.dtype, ch.cfg, blocal, P.gosign, BSQ_OP_LAUNCH, grid, stream, &A);
  if (rc < 0) return fail(ctx, BSQ_ERR_ARG, "no kernel variant for this configuration");
  BSQ_CK(ctx, cudaGetLastError());
  BSQ_CK(ctx, cudaEventRecord(e1, stream));
  ctx->tm.fill_launches++;
  ctx->tm.ptr_bytes += 4 * ch.ptr_words;
  return BSQ_SUCCESS;
}

int launch_trace(bsq_ctx* ctx, BufSet& B, const Plan& P, size_t ci, cudaStream_t stream) {
  const Chunk& ch = P.chunks[ci];
  BsqTraceArgs T;
  memset(&T, 0, sizeof T);
  T.codes = B.codes.as<uint8_t>();
  T.pairs = B.pairs.as<BsqPair>() + ch.first;
  T.npairs = (int)ch.count;
  T.mode = P.prm.mode;
  const bool k3 = chunk_is_k3(ch);
  T.family = k3 ? 3 : ch.family;
  T.G = k3 ? 32 : (ch.family == 1 ? BSQ_K1_CFGS[ch.cfg].G : BSQ_K2_CFGS[ch.cfg].G);
  T.K = k3 ? BSQ_K3_CFGS[ch.cfg - BSQ_K2_NCFG].K : (ch.family == 1 ? BSQ_K1_CFGS[ch.cfg].K : BSQ_K2_CFGS[ch.cfg].C);
  T.gosign = P.gosign;
  T.ptr = B.ptr.as<uint32_t>() + (P.half_words > 0 ? (int64_t)(ci % BSQ_ARENA_SLOTS) * P.half_words : 0);
  T.ex = B.ex.as<int>();
  T.ey = B.ey.as<int>();
  T.status = B.status.as<int>();
  T.sx = B.sx.as<int>();
  T.sy = B.sy.as<int>();
  T.txlen = B.txlen.as<int>();
  T.slots = B.slots.as<char>();
  T.bad = B.flags.as<int>();
  T.codes2 = (P.codes2_ok && B.codes2.p) ? B.codes2.as<uint32_t>() : nullptr;
  // long pairs: one warp per pair walking shared-memory tiles; short pairs: one
  // thread per pair (there are enough of them to hide the load latency)
  const long long avg_path = P.slot_bytes / std::max<long long>((long long)P.E, 1);
  const char force = ctx->hooks.trace_kernel;  // test hook: 't'hread / 'w'arp / 'n'arrow
  bool warp_walk = avg_path >= 1024;
  if (force == 't') warp_walk = false;
  if (force == 'w' || force == 'n') warp_walk = true;  // "narrow": 8 lanes per pair
  if (k3) {  // one warp per pair over the [strip][t][lane] table
    const int grid = (int)ch.count;
#define BSQ_WALK3(KK) { if (P.gosign < 0) BSQ_LAUNCH((bsq_k3_walk<KK, -1>), grid, 32, 0, stream, T); \
                        else if (P.gosign == 0) BSQ_LAUNCH((bsq_k3_walk<KK, 0>), grid, 32, 0, stream, T); \
                        else BSQ_LAUNCH((bsq_k3_walk<KK, 1>), grid, 32, 0, stream, T); }
    if (T.K == 2) BSQ_WALK3(2) else if (T.K == 4) BSQ_WALK3(4) else BSQ_WALK3(8)
#undef BSQ_WALK3
  } else if (ch.family == 2 && force == 'p') {
    // banded tables, a thread per pair straight out of global memory (hook: kept for comparison)
    const int grid = (int)((ch.count + 127) / 128);
#define BSQ_WALKT(GS) { if (T.codes2) BSQ_LAUNCH((bsq_k2_walk_thread<GS, true>), grid, 128, 0, stream, T); \
                        else BSQ_LAUNCH((bsq_k2_walk_thread<GS, false>), grid, 128, 0, stream, T); }
    if (P.gosign < 0) BSQ_WALKT(-1) else if (P.gosign == 0) BSQ_WALKT(0) else BSQ_WALKT(1)
#undef BSQ_WALKT
  } else if (ch.family == 2 && force != 't') {
    // banded tables ([block of steps][lane][unit]): a warp per pair; 8 lanes per pair ('n' hook) measured
    // slower at every batch size (config 3, 32768 pairs: 9.3 ms against 6.7 ms; config 4: 17 against 4 ms)
    const bool narrow = (force == 'n');
    const int ppb = BSQ_WALKC_WARPS * (narrow ? 4 : 1);  // pairs per block
    const int grid = (int)((ch.count + ppb - 1) / ppb);
#define BSQ_WALKC(GS) { if (narrow) BSQ_LAUNCH((bsq_k2_walk_coop<GS, 8>), grid, 32 * BSQ_WALKC_WARPS, 0, stream, T); \
                        else BSQ_LAUNCH((bsq_k2_walk_coop<GS, 32>), grid, 32 * BSQ_WALKC_WARPS, 0, stream, T); }
    if (P.gosign < 0) BSQ_WALKC(-1) else if (P.gosign == 0) BSQ_WALKC(0) else BSQ_WALKC(1)
#undef BSQ_WALKC
  } else if (warp_walk) {
    // lanes per pair: 8 when there are enough pairs to fill the GPU four times over, else a warp
    const bool narrow = (ch.count >= (int64_t)ctx->sms * 16 * 4 || force == 'n') && force != 'w';
    const int ppb = BSQ_WALK_WARPS * (narrow ? 4 : 1);  // pairs per block
    const int grid = (int)((ch.count + ppb - 1) / ppb);
#define BSQ_WALK(F, GS) { if (narrow) BSQ_LAUNCH((bsq_k4_walk<F, GS, 8>), grid, 32 * BSQ_WALK_WARPS, 0, stream, T); \
                          else BSQ_LAUNCH((bsq_k4_walk<F, GS, 32>), grid, 32 * BSQ_WALK_WARPS, 0, stream, T); }
    if (P.gosign < 0) BSQ_WALK(1, -1) else if (P.gosign == 0) BSQ_WALK(1, 0) else BSQ_WALK(1, 1)
#undef BSQ_WALK
  } else {
    // small blocks so they fit beside a resident fill kernel's register allocation
    const int grid = (int)((ch.count + 63) / 64);
    BSQ_LAUNCH(bsq_k4_traceback, grid, 64, 0, stream, T);
  }
  BSQ_CK(ctx, cudaGetLastError());
  ctx->tm.fill_launches++;
  return BSQ_SUCCESS;
}

// scan txlen (original order) -> txoff, pack slots -> packed; all on device
int pack_transcripts(bsq_ctx* ctx, BufSet& B, const Plan& P) {
  const int64_t N = P.N, E = (int64_t)P.E;
  const int ntiles = (int)((N + BSQ_SCAN_TILE - 1) / BSQ_SCAN_TILE);
  if (!B.txoff.ensure(8 * (size_t)(N + 1)) || !B.tiles.ensure(8 * (size_t)std::max(ntiles, 1)) ||
      !B.packed.ensure((size_t)std::max<int64_t>(P.slot_bytes, 1)))
    return fail(ctx, BSQ_ERR_NOMEM, "device allocation failed (transcripts)");
  cudaStream_t st = ctx->st;
  BSQ_LAUNCH(bsq_scan_tile_sums, ntiles, 256, 0, st, (const int*)B.txlen.as<int>(), (int)N, B.tiles.as<long long>());
  BSQ_LAUNCH(bsq_scan_tiles, 1, 256, 0, st, B.tiles.as<long long>(), ntiles, (const long long*)nullptr,
             B.total.as<long long>());
  BSQ_LAUNCH(bsq_scan_apply, ntiles, 256, 0, st, (const int*)B.txlen.as<int>(), (int)N,
             (const long long*)B.tiles.as<long long>(), B.txoff.as<long long>());
  ctx->tm.fill_launches += 3;
  if (E > 0) {
    const dim3 grid((unsigned)((E * 32 + 255) / 256), pack_segments(P));
    BSQ_LAUNCH(bsq_pack_transcripts, grid, 256, 0, st, (const BsqPair*)B.pairs.as<BsqPair>(), (int)E,
               (const int*)B.txlen.as<int>(), (const long long*)B.txoff.as<long long>(),
               (const char*)B.slots.as<char>(), B.packed.as<char>());
    ctx->tm.fill_launches++;
  }
  BSQ_CK(ctx, cudaGetLastError());
  return BSQ_SUCCESS;
}

// BSQ_F_TX_2BIT: words [w_lo, w_hi) of the 2-bit form of the packed byte transcripts (16 ops per
// 32-bit word), converted on stream st and copied to tx_buf right behind
int ship_tx_2bit(bsq_ctx* ctx, BufSet& B, long long nops, long long w_lo, long long w_hi, char* tx_buf, cudaStream_t st) {
  if (w_hi <= w_lo) return BSQ_SUCCESS;
  const int grid = (int)((w_hi - w_lo + 255) / 256);
  BSQ_LAUNCH(bsq_tx_to_2bit, grid, 256, 0, st, (const char*)B.packed.as<char>(), nops, w_lo, w_hi, B.packed2.as<uint32_t>());
  BSQ_CK(ctx, cudaGetLastError());
  ctx->tm.fill_launches++;
  BSQ_CK(ctx, cudaMemcpyAsync(tx_buf + 4 * w_lo, (const char*)B.packed2.p + 4 * w_lo, (size_t)(4 * (w_hi - w_lo)), cudaMemcpyDeviceToHost, st));
  ctx->tm.d2h_bytes += 4 * (w_hi - w_lo);
  return BSQ_SUCCESS;
}

int fetch_transcripts(bsq_ctx* ctx, BufSet& B, const Plan& P, int64_t* tx_off, char* tx_buf,
                      int64_t tx_cap, int64_t* tx_total) {
  const int64_t N = P.N;
  cudaStream_t st = ctx->st;
  long long total = 0;
  BSQ_CK(ctx, cudaMemcpyAsync(&total, B.total.p, 8, cudaMemcpyDeviceToHost, st));
  BSQ_CK(ctx, cudaMemcpyAsync(tx_off, B.txoff.p, 8 * (size_t)N, cudaMemcpyDeviceToHost, st));
  BSQ_CK(ctx, cudaStreamSynchronize(st));
  tx_off[N] = total;
  if (tx_total) *tx_total = total;
  ctx->tm.d2h_bytes += 8 * N + 8;
  if (P.prm.flags & BSQ_F_TX_2BIT) {
    const long long words = (total + 15) / 16;
    if (4 * words > tx_cap) return fail(ctx, BSQ_ERR_CAPACITY, "tx_buf too small for the batch's transcripts");
    if (!B.packed2.ensure((size_t)(4 * words + 16))) return fail(ctx, BSQ_ERR_NOMEM, "device allocation failed (2-bit transcripts)");
    const int rc2 = ship_tx_2bit(ctx, B, total, 0, words, tx_buf, st);
    if (rc2 != BSQ_SUCCESS) return rc2;
    BSQ_CK(ctx, cudaStreamSynchronize(st));
    return BSQ_SUCCESS;
  }
  if (total > tx_cap) return fail(ctx, BSQ_ERR_CAPACITY, "tx_buf too small for the batch's transcripts");
  if (total > 0) {
    BSQ_CK(ctx, cudaMemcpyAsync(tx_buf, B.packed.p, (size_t)total, cudaMemcpyDeviceToHost, st));
    BSQ_CK(ctx, cudaStreamSynchronize(st));
    ctx->tm.d2h_bytes += total;
  }
  return BSQ_SUCCESS;
}

void apply_overrides(const Plan& P, double* score, int32_t* ex, int32_t* ey, int32_t* sx, int32_t* sy,
                     int32_t* status) {
  for (const Override& o : P.overrides) {
    if (score) score[o.idx] = 0.0;
    if (ex) ex[o.idx] = -1;
    if (ey) ey[o.idx] = -1;
    if (sx) sx[o.idx] = -1;
    if (sy) sy[o.idx] = -1;
    if (status) status[o.idx] = o.status;
  }
}

int collect_timing(bsq_ctx* ctx, size_t nev, cudaEvent_t t0, cudaEvent_t t1) {
  double fill = 0;
  for (size_t i = 2; i + 1 < nev; i += 2) {  // events 0/1 bracket the whole call
    float ms = 0;
    if (cudaEventElapsedTime(&ms, ctx->events[i], ctx->events[i + 1]) == cudaSuccess) fill += ms;
  }
  float tot = 0;
  cudaEventElapsedTime(&tot, t0, t1);
  ctx->tm.fill_ms = fill;
  ctx->tm.total_ms = tot;
  ctx->tm.traceback_ms = 0;
  return 0;
}

}  // namespace

// ===========================================================================
// C ABI
// ===========================================================================
extern "C" {

const char* bsq_version(void) {
#ifdef BSQ_SIMT_EMU
  return "bsqgpu 0.1 (simt-emu test build)";
#else
  return "bsqgpu 0.1 (sm_100a)";
#endif
}

int bsq_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}

const char* bsq_last_error(const bsq_ctx* ctx) { return ctx ? ctx->err.c_str() : g_err.c_str(); }

bsq_ctx* bsq_create(int device) {
  int n = bsq_device_count();
  if (n <= 0) { g_err = "no CUDA device available (libbsqgpu has no CPU fallback)"; return nullptr; }
  if (device < 0 || device >= n) { g_err = "device ordinal out of range"; return nullptr; }
  if (cudaSetDevice(device) != cudaSuccess) { g_err = "cudaSetDevice failed"; return nullptr; }
  bsq_ctx* ctx = new bsq_ctx();
  ctx->device = device;
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) { g_err = "cudaGetDeviceProperties failed"; delete ctx; return nullptr; }
  ctx->sms = prop.multiProcessorCount;
  if (cudaStreamCreateWithFlags(&ctx->st, cudaStreamNonBlocking) != cudaSuccess ||
      cudaStreamCreateWithFlags(&ctx->st_copy, cudaStreamNonBlocking) != cudaSuccess ||
      cudaStreamCreateWithFlags(&ctx->st_copy2, cudaStreamNonBlocking) != cudaSuccess ||
      cudaStreamCreateWithFlags(&ctx->st_d2h, cudaStreamNonBlocking) != cudaSuccess ||
      cudaStreamCreateWithPriority(&ctx->st_trace, cudaStreamNonBlocking, -1) != cudaSuccess) {
    g_err = "cudaStreamCreate failed"; delete ctx; return nullptr;
  }
  memset(&ctx->tm, 0, sizeof ctx->tm);
  bsq_reload_hooks(ctx);
  return ctx;
}

void bsq_destroy(bsq_ctx* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->st);
  cudaStreamSynchronize(ctx->st_trace);
  // batches that outlive their context (Python: an Engine closed before its SolvedBatch /
  // ResidentBatch objects are collected): give their buffers back now and orphan them, so that a
  // later bsq_release / bsq_traceback_batch / bsq_batch_* on them touches nothing of this context
  for (bsq_batch* b : ctx->live) {
    if (b->own) { b->own->release(); delete b->own; b->own = nullptr; }
    b->B = nullptr;
    b->ctx = nullptr;
  }
  ctx->live.clear();
  ctx->bufs.release();
  drop_pool(ctx);
  ctx->htotals.release();
  for (cudaEvent_t e : ctx->events) cudaEventDestroy(e);
  cudaStreamDestroy(ctx->st);
  cudaStreamDestroy(ctx->st_copy);
  cudaStreamDestroy(ctx->st_copy2);
  cudaStreamDestroy(ctx->st_d2h);
  cudaStreamDestroy(ctx->st_trace);
  delete ctx;
}

void bsq_reload_hooks(bsq_ctx* ctx) {
  if (!ctx) return;
  ctx->hooks.read_env();
  const char* e = getenv("BSQ_PAR_GRAIN");
  g_par_grain = e ? atoll(e) : 32768;
}

void* bsq_host_alloc(int64_t bytes) {
  void* p = nullptr;
  if (bytes < 0) return nullptr;
  if (cudaHostAlloc(&p, (size_t)(bytes ? bytes : 1), cudaHostAllocDefault) != cudaSuccess) { cudaGetLastError(); return nullptr; }
  return p;
}
void bsq_host_free(void* p) { if (p) cudaFreeHost(p); }

int bsq_get_timing(const bsq_ctx* ctx, bsq_timing* out) {
  if (!ctx || !out) return BSQ_ERR_ARG;
  *out = ctx->tm;
  return BSQ_SUCCESS;
}

int64_t bsq_count_cells(int32_t mode, int32_t n, int32_t m, int32_t dmin, int32_t dmax) {
  return count_cells(mode, n, m, dmin, dmax);
}

int bsq_scores_are_int32_exact(const bsq_params* params, int32_t max_n, int32_t max_m) {
  if (!params || !params->subst) return 0;
  return scores_int32_exact(params, max_n, max_m) ? 1 : 0;
}

int bsq_partition_pairs(int32_t mode, int64_t npairs, const int32_t* o_len, const int32_t* m_len,
                        const int32_t* dmin, const int32_t* dmax, int32_t nshards, int32_t* shard_of,
                        int64_t* load) {
  if (npairs < 0 || nshards < 1 || !o_len || !m_len || !shard_of) return BSQ_ERR_ARG;
  if (mode == BSQ_BANDED_MODE && (!dmin || !dmax)) return BSQ_ERR_ARG;
  std::vector<int64_t> w((size_t)npairs), order((size_t)npairs);
  for (int64_t i = 0; i < npairs; i++) {
    w[i] = count_cells(mode, o_len[i], m_len[i], dmin ? dmin[i] : 0, dmax ? dmax[i] : 0);
    order[i] = i;
  }
  std::stable_sort(order.begin(), order.end(), [&](int64_t a, int64_t b) { return w[a] > w[b]; });
  std::vector<int64_t> ld((size_t)nshards, 0);
  for (int64_t r = 0; r < npairs; r++) {
    const int64_t i = order[r];
    int bestk = 0;
    for (int k = 1; k < nshards; k++) if (ld[k] < ld[bestk]) bestk = k;
    shard_of[i] = bestk;
    ld[bestk] += w[i];
  }
  if (load) for (int k = 0; k < nshards; k++) load[k] = ld[k];
  return BSQ_SUCCESS;
}

static int check_inputs(bsq_ctx* ctx, int64_t npairs, const uint8_t* codes, const int64_t* o_off,
                        const int32_t* o_len, const int64_t* m_off, const int32_t* m_len) {
  if (npairs < 0 || npairs >= ((int64_t)1 << 31)) return fail(ctx, BSQ_ERR_ARG, "npairs out of range");
  if (npairs > 0 && (!codes || !o_len || !m_len)) return fail(ctx, BSQ_ERR_ARG, "NULL input array");
  if ((o_off == nullptr) != (m_off == nullptr)) return fail(ctx, BSQ_ERR_ARG, "o_off and m_off must both be given or both be NULL");
  return BSQ_SUCCESS;
}

// plan + H2D.  own_bufs: the batch gets its own HBM buffers (two-phase / resident
// use); otherwise the context's cached buffers are borrowed (one-shot).
static int batch_upload(bsq_ctx* ctx, bsq_batch* b, bool own_bufs, bool single_chunk, const bsq_params* params,
                        int64_t npairs, const uint8_t* codes, int64_t ncodes, const int64_t* o_off,
                        const int32_t* o_len, const int64_t* m_off, const int32_t* m_len,
                        const int32_t* dmin, const int32_t* dmax) {
  b->ctx = ctx;
  b->own_bufs = own_bufs;
  if (own_bufs && !b->own) {
    if (!ctx->pool.empty()) { b->own = ctx->pool.back(); ctx->pool.pop_back(); }
    else b->own = new BufSet();
    ctx->live.push_back(b);
  }
  b->B = own_bufs ? b->own : &ctx->bufs;
  b->plan.N = npairs;
  if (npairs == 0) return BSQ_SUCCESS;
  // a resident batch must not be squeezed into what the one-shot path's cached arena leaves free
  // (measured: config 3 cut into 34 small chunks instead of 2, fill 3x slower): give that back first
  if (own_bufs && ctx->bufs.bytes() > ((size_t)1 << 30)) {
    cudaStreamSynchronize(ctx->st);
    cudaStreamSynchronize(ctx->st_trace);
    ctx->bufs.release();
  }
  int64_t slot_guess = 0;
  for (int64_t i = 0; i < npairs; i++) slot_guess += (int64_t)o_len[i] + m_len[i];
  const int64_t other = ncodes + 2 * slot_guess + npairs * (int64_t)(sizeof(BsqPair) + 48) + (64 << 20);
  const int64_t arena_words = arena_budget_words(ctx, other, b->B);
  int rc = make_plan(ctx, b->plan, b->B->hpairs, params, npairs, ncodes, o_off, o_len, m_off, m_len, dmin, dmax,
                     single_chunk ? arena_words : arena_words / BSQ_ARENA_SLOTS, single_chunk,
                     single_chunk ? 0 : overlap_chunks(ctx, npairs, o_len, m_len));
  if (rc != BSQ_SUCCESS) return rc;
  b->single_chunk = single_chunk;
  return stage_inputs(ctx, *b->B, b->plan, codes, ncodes, single_chunk);
}

// Enqueues fill (compute stream) and traceback (+ optional per-chunk work, trace
// stream) for every chunk.  With two arena halves the traceback of chunk c runs
// beside the fill of chunk c+1: the fill is ALU-bound, the walk latency-bound.
// done[ci] is recorded after the chunk's last kernel.
extern "C++" {
template <class BeforeFill, class AfterTrace>
static int run_chunks(bsq_ctx* ctx, BufSet& B, const Plan& P, bool do_fill, bool do_trace, size_t* evi,
                      std::vector<size_t>* fill_ev0, std::vector<cudaEvent_t>* done, BeforeFill before_fill,
                      AfterTrace after_trace) {
  const size_t NC = P.chunks.size();
  cudaStream_t sc = ctx->st, stt = ctx->st_trace;
  const bool overlap = do_fill && do_trace && P.half_words > 0;
  std::vector<cudaEvent_t> efill(NC), etrace(NC);
  if (fill_ev0) fill_ev0->assign(NC, 0);
  if (done) done->assign(NC, nullptr);
  cudaEvent_t estart = get_event(ctx, (*evi)++);
  BSQ_CK(ctx, cudaEventRecord(estart, sc));
  if (do_trace) BSQ_CK(ctx, cudaStreamWaitEvent(stt, estart, 0));  // the trace stream joins the timeline
  for (size_t ci = 0; ci < NC; ci++) {
    int rc = before_fill(ci);
    if (rc != BSQ_SUCCESS) return rc;
    if (do_fill) {
      if (overlap && ci >= BSQ_ARENA_SLOTS) BSQ_CK(ctx, cudaStreamWaitEvent(sc, etrace[ci - BSQ_ARENA_SLOTS], 0));  // that arena slot is free again
      if (fill_ev0) (*fill_ev0)[ci] = *evi;
      rc = launch_fill(ctx, B, P, ci, evi, sc, overlap);
      if (rc != BSQ_SUCCESS) return rc;
      efill[ci] = get_event(ctx, (*evi)++);
      BSQ_CK(ctx, cudaEventRecord(efill[ci], sc));
    }
    if (do_trace) {
      cudaStream_t ts = overlap ? stt : sc;
      if (overlap) BSQ_CK(ctx, cudaStreamWaitEvent(ts, efill[ci], 0));
      rc = launch_trace(ctx, B, P, ci, ts);
      if (rc == BSQ_SUCCESS) rc = after_trace(ci, ts);
      if (rc != BSQ_SUCCESS) return rc;
      etrace[ci] = get_event(ctx, (*evi)++);
      BSQ_CK(ctx, cudaEventRecord(etrace[ci], ts));
      if (done) (*done)[ci] = etrace[ci];
    } else if (done) {
      (*done)[ci] = efill[ci];
    }
  }
  if (overlap) {  // the compute stream rejoins: everything after this sees all traces done
    for (size_t ci = (NC >= BSQ_ARENA_SLOTS ? NC - BSQ_ARENA_SLOTS : 0); ci < NC; ci++) BSQ_CK(ctx, cudaStreamWaitEvent(sc, etrace[ci], 0));
  }
  return BSQ_SUCCESS;
}
}  // extern "C++"

// fill (+ traceback + pack) entirely on the device; returns after the stream drained
static int batch_run(bsq_batch* b, bool do_fill, bool do_trace) {
  bsq_ctx* ctx = b->ctx;
  Plan& P = b->plan;
  BufSet& B = *b->B;
  if (P.N == 0) return BSQ_SUCCESS;
  if (!do_fill && !b->single_chunk)
    return fail(ctx, BSQ_ERR_ARG, "traceback-only run needs resident tables (two-phase batch)");
  cudaEvent_t t0 = get_event(ctx, 0), t1 = get_event(ctx, 1);
  size_t evi = 2;
  int rc = BSQ_SUCCESS;
  if (do_trace && !B.slots.ensure((size_t)std::max<int64_t>(P.slot_bytes, 1)))
    return fail(ctx, BSQ_ERR_NOMEM, "device allocation failed (transcript slots)");
  BSQ_CK(ctx, cudaEventRecord(t0, ctx->st));
  if (do_fill) BSQ_CK(ctx, cudaMemsetAsync(B.counters.p, 0, 4 * (P.chunks.size() + 1), ctx->st));
  std::vector<size_t> fill_ev0;
  rc = run_chunks(ctx, B, P, do_fill, do_trace, &evi, &fill_ev0, nullptr, [](size_t) { return BSQ_SUCCESS; },
                  [](size_t, cudaStream_t) { return BSQ_SUCCESS; });
  if (rc != BSQ_SUCCESS) { cudaStreamSynchronize(ctx->st); cudaStreamSynchronize(ctx->st_trace); return rc; }
  if (do_trace) {
    rc = pack_transcripts(ctx, B, P);
    if (rc != BSQ_SUCCESS) return rc;
    b->traced = true;
  }
  BSQ_CK(ctx, cudaEventRecord(t1, ctx->st));
  BSQ_CK(ctx, cudaStreamSynchronize(ctx->st));
  BSQ_CK(ctx, cudaStreamSynchronize(ctx->st_trace));
  collect_timing(ctx, 2, t0, t1);
  if (do_fill) {
    double fill = 0;
    for (size_t ci = 0; ci < fill_ev0.size(); ci++) {
      float ms = 0;
      if (cudaEventElapsedTime(&ms, ctx->events[fill_ev0[ci]], ctx->events[fill_ev0[ci] + 1]) == cudaSuccess) fill += ms;
    }
    ctx->tm.fill_ms = fill;
  }
  ctx->tm.cells = P.cells;
  ctx->tm.dtype = (!P.chunks.empty() && P.chunks[0].duo) ? BSQ_DTYPE_I16X2 : P.dtype;
  ctx->tm.chunks = (int32_t)P.chunks.size();
  if (do_fill) b->filled = true;
  return BSQ_SUCCESS;
}

static int batch_fetch(bsq_batch* b, double* score, int32_t* end_x, int32_t* end_y, int32_t* start_x,
                       int32_t* start_y, int32_t* status, int64_t* tx_off, char* tx_buf, int64_t tx_cap,
                       int64_t* tx_total) {
  bsq_ctx* ctx = b->ctx;
  Plan& P = b->plan;
  BufSet& B = *b->B;
  if (tx_total) *tx_total = 0;
  if (P.N == 0) { if (tx_off) tx_off[0] = 0; return BSQ_SUCCESS; }
  cudaStream_t st = ctx->st;
  const size_t N = (size_t)P.N;
  if (score) BSQ_CK(ctx, cudaMemcpyAsync(score, B.score.p, 8 * N, cudaMemcpyDeviceToHost, st));
  if (end_x) BSQ_CK(ctx, cudaMemcpyAsync(end_x, B.ex.p, 4 * N, cudaMemcpyDeviceToHost, st));
  if (end_y) BSQ_CK(ctx, cudaMemcpyAsync(end_y, B.ey.p, 4 * N, cudaMemcpyDeviceToHost, st));
  if (b->traced && start_x) BSQ_CK(ctx, cudaMemcpyAsync(start_x, B.sx.p, 4 * N, cudaMemcpyDeviceToHost, st));
  if (b->traced && start_y) BSQ_CK(ctx, cudaMemcpyAsync(start_y, B.sy.p, 4 * N, cudaMemcpyDeviceToHost, st));
  if (status) BSQ_CK(ctx, cudaMemcpyAsync(status, B.status.p, 4 * N, cudaMemcpyDeviceToHost, st));
  ctx->tm.d2h_bytes += (int64_t)N * (8 * (score != 0) + 4 * ((status != 0) + (end_x != 0) + (end_y != 0) +
                                                             (start_x != 0) + (start_y != 0)));
  int rc = BSQ_SUCCESS;
  if (b->traced && tx_off && tx_buf) rc = fetch_transcripts(ctx, B, P, tx_off, tx_buf, tx_cap, tx_total);
  BSQ_CK(ctx, cudaStreamSynchronize(st));
  apply_overrides(P, score, end_x, end_y, b->traced ? start_x : nullptr, b->traced ? start_y : nullptr, status);
  return rc;
}

// One-shot path, pipelined: the sequence bytes start moving on the copy stream
// before the host has planned anything; the batch is cut into ~8 chunks; the
// compute stream runs fill -> traceback -> pack chunk after chunk while the copy
// stream ships finished chunks back, so H2D, kernels and D2H overlap.
static double now_ms() {
  return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

static int align_batch_once(bsq_ctx* ctx, const bsq_params* params, int64_t npairs, const uint8_t* codes,
                            int64_t ncodes, const int64_t* o_off, const int32_t* o_len, const int64_t* m_off,
                            const int32_t* m_len, const int32_t* dmin, const int32_t* dmax, double* score,
                            int32_t* end_x, int32_t* end_y, int32_t* start_x, int32_t* start_y, int32_t* status,
                            int64_t* tx_off, char* tx_buf, int64_t tx_cap, int64_t* tx_total, int64_t off_base);

static int align_batch_retry(bsq_ctx* ctx, const bsq_params* params, int64_t npairs, const uint8_t* codes,
                             int64_t ncodes, const int64_t* o_off, const int32_t* o_len, const int64_t* m_off,
                             const int32_t* m_len, const int32_t* dmin, const int32_t* dmax, double* score,
                             int32_t* end_x, int32_t* end_y, int32_t* start_x, int32_t* start_y, int32_t* status,
                             int64_t* tx_off, char* tx_buf, int64_t tx_cap, int64_t* tx_total, int64_t off_base) {
  if (!ctx) return BSQ_ERR_ARG;
  int rc = align_batch_once(ctx, params, npairs, codes, ncodes, o_off, o_len, m_off, m_len, dmin, dmax, score, end_x,
                            end_y, start_x, start_y, status, tx_off, tx_buf, tx_cap, tx_total, off_base);
  if (rc == BSQ_ERR_NOMEM) {  // the remembered free-memory figure may be stale: ask the driver and try once more
    drop_pool(ctx);
    ctx->mem_avail = -1.0;
    rc = align_batch_once(ctx, params, npairs, codes, ncodes, o_off, o_len, m_off, m_len, dmin, dmax, score, end_x,
                          end_y, start_x, start_y, status, tx_off, tx_buf, tx_cap, tx_total, off_base);
  }
  return rc;
}

int bsq_align_batch(bsq_ctx* ctx, const bsq_params* params, int64_t npairs, const uint8_t* codes,
                    int64_t ncodes, const int64_t* o_off, const int32_t* o_len, const int64_t* m_off,
                    const int32_t* m_len, const int32_t* dmin, const int32_t* dmax, double* score,
                    int32_t* end_x, int32_t* end_y, int32_t* start_x, int32_t* start_y, int32_t* status,
                    int64_t* tx_off, char* tx_buf, int64_t tx_cap, int64_t* tx_total) {
  return align_batch_retry(ctx, params, npairs, codes, ncodes, o_off, o_len, m_off, m_len, dmin, dmax, score, end_x,
                           end_y, start_x, start_y, status, tx_off, tx_buf, tx_cap, tx_total, 0);
}

static int align_batch_once(bsq_ctx* ctx, const bsq_params* params, int64_t npairs, const uint8_t* codes,
                            int64_t ncodes, const int64_t* o_off, const int32_t* o_len, const int64_t* m_off,
                            const int32_t* m_len, const int32_t* dmin, const int32_t* dmax, double* score,
                            int32_t* end_x, int32_t* end_y, int32_t* start_x, int32_t* start_y, int32_t* status,
                            int64_t* tx_off, char* tx_buf, int64_t tx_cap, int64_t* tx_total, int64_t off_base) {
  if (!ctx) return BSQ_ERR_ARG;
  const double t_call = now_ms();
  ctx->err.clear();
  int rc = check_inputs(ctx, npairs, codes, o_off, o_len, m_off, m_len);
  if (rc != BSQ_SUCCESS) return rc;
  if (!score || !status) return fail(ctx, BSQ_ERR_ARG, "score/status must not be NULL");
  BSQ_CK(ctx, cudaSetDevice(ctx->device));
  memset(&ctx->tm, 0, sizeof ctx->tm);
  if (tx_total) *tx_total = 0;
  if (npairs == 0) { if (tx_off) tx_off[0] = 0; return BSQ_SUCCESS; }
  const bool want_tx = (tx_off && tx_buf);
  const bool want_trace = want_tx || start_x || start_y;
  BufSet& B = ctx->bufs;
  cudaStream_t sc = ctx->st, sx = ctx->st_copy;
  const size_t N = (size_t)npairs;

  // ---- 0. sequence bytes: H2D in slices on the copy stream, one event per slice.  Only the first
  // slice is enqueued before the plan exists: the H2D engine serves its queue in order, and the
  // (small) pair records must not wait behind all the sequence bytes (measured: the first fill
  // started 7 ms late).  The other slices follow the pair records.
  const int NSL = 8;
  size_t evi = 2;  // events 0/1 bracket the call on the compute stream
  if (!params) return fail(ctx, BSQ_ERR_ARG, "params is NULL");
  if ((params->flags & BSQ_F_CODES_2BIT) && params->alphabet_len > 4) return fail(ctx, BSQ_ERR_ARG, "BSQ_F_CODES_2BIT needs an alphabet of at most 4 letters");
  if (!ensure_codes(B, *params, ncodes)) return fail(ctx, BSQ_ERR_NOMEM, "device allocation failed (codes)");
  cudaEvent_t ev_code[NSL];
  int64_t slice_end[NSL];      // in LETTERS (what Chunk.code_end counts)
  int64_t slice_b0[NSL], slice_b1[NSL];   // in bytes of the caller's codes[]
  const int64_t in_bytes = codes_in_bytes(*params, ncodes);
  const int64_t lpb = (params->flags & BSQ_F_CODES_2BIT) ? 4 : 1;   // letters per input byte
  const int64_t slice = ((in_bytes + NSL - 1) / NSL + 255) / 256 * 256;
  int nsl = 0;
  for (int64_t off = 0; off < in_bytes && nsl < NSL; off += slice, nsl++) {
    const int64_t len = std::min<int64_t>(slice, in_bytes - off);
    ev_code[nsl] = get_event(ctx, evi++);
    slice_b0[nsl] = off;
    slice_b1[nsl] = off + len;
    slice_end[nsl] = std::min<int64_t>(ncodes, (off + len) * lpb);
  }
  auto enqueue_slices = [&](int from, int to) -> int {
    for (int i = from; i < to && i < nsl; i++) {
      const int rcu = upload_codes(ctx, B, *params, codes, slice_b0[i], slice_b1[i], sx);
      if (rcu != BSQ_SUCCESS) return rcu;
      BSQ_CK(ctx, cudaEventRecord(ev_code[i], sx));
    }
    return BSQ_SUCCESS;
  };
  rc = enqueue_slices(0, 1);
  if (rc != BSQ_SUCCESS) return rc;
  ctx->tm.h2d_bytes += in_bytes;

  // ---- 1. plan on the host while the bytes fly
  const double t_plan = now_ms();
  bsq_batch b;
  b.ctx = ctx; b.own_bufs = false; b.B = &B;
  Plan& P = b.plan;
  int64_t slot_guess = 0;
  {
    std::vector<int64_t> part(16, 0);
    parallel_for(npairs, [&](int t, int64_t i0, int64_t i1) {
      int64_t a = 0;
      for (int64_t i = i0; i < i1; i++) a += (int64_t)o_len[i] + m_len[i];
      part[t] = a;
    });
    for (int64_t v : part) slot_guess += v;
  }
  const int64_t other = ncodes + 2 * slot_guess + npairs * (int64_t)(sizeof(BsqPair) + 48) + (64 << 20);
  const double t_p1 = now_ms();
  const int64_t budget = arena_budget_words(ctx, other, &B);
  const double t_p2 = now_ms();
  rc = make_plan(ctx, P, B.hpairs, params, npairs, ncodes, o_off, o_len, m_off, m_len, dmin, dmax,
                 budget / BSQ_ARENA_SLOTS, false, overlap_chunks(ctx, npairs, o_len, m_len), off_base);
  if (rc != BSQ_SUCCESS) { cudaStreamSynchronize(sx); return rc; }
  ctx->tm.plan_ms = now_ms() - t_plan;
  P.codes2_ok = (params->flags & BSQ_F_CODES_2BIT) || want_packed_copy(*params);   // what enqueue_slices uploads
  if (ctx->hooks.plan_debug)
    fprintf(stderr, "[bsq] plan: memgetinfo %.3f ms, make_plan %.3f ms, h2d enqueue before plan %.3f ms\n",
            t_p2 - t_p1, now_ms() - t_p2, t_plan - t_call);
  const int64_t E = (int64_t)P.E;
  const int L = P.prm.alphabet_len;
  const size_t NC = P.chunks.size();
  const int ntiles_max = (int)((N + BSQ_SCAN_TILE - 1) / BSQ_SCAN_TILE) + (int)NC + 1;
  bool ok = B.pairs.ensure(sizeof(BsqPair) * (size_t)std::max<int64_t>(E, 1)) && B.score.ensure(8 * N) &&
            B.ex.ensure(4 * N) && B.ey.ensure(4 * N) && B.sx.ensure(4 * N) && B.sy.ensure(4 * N) &&
            B.status.ensure(4 * N) && B.txlen.ensure(4 * N) && B.subst.ensure(8 * (size_t)L * L + 64) &&
            B.counters.ensure(4 * (NC + 1)) && B.total.ensure(8 * (NC + 2)) && B.flags.ensure(16) &&
            B.ptr.ensure(4 * (size_t)(P.half_words > 0 ? BSQ_ARENA_SLOTS * P.half_words : P.max_chunk_words) + 64) &&
            ctx->htotals.ensure(8 * (NC + 2));
  if (ok && want_trace) ok = B.slots.ensure((size_t)std::max<int64_t>(P.slot_bytes, 1));
  if (ok && want_tx) ok = B.txoff.ensure(8 * (N + 1)) && B.tiles.ensure(8 * (size_t)ntiles_max) &&
                          B.packed.ensure((size_t)std::max<int64_t>(P.slot_bytes, 1));
  const bool tx2 = want_tx && (P.prm.flags & BSQ_F_TX_2BIT) != 0;
  if (ok && tx2) ok = B.packed2.ensure((size_t)(P.slot_bytes / 4 + 64));
  if (!ok) { ctx->mem_avail = -1.0; cudaStreamSynchronize(sx); return fail(ctx, BSQ_ERR_NOMEM, "device allocation failed"); }

  // ---- 2. pair records + score table right behind the first slice of sequence bytes, the rest of
  // the sequence bytes behind them
  cudaStream_t sp = sx;
  std::vector<int> subst_i;
  rc = upload_pairs(ctx, B, P, sp);
  if (rc != BSQ_SUCCESS) { cudaStreamSynchronize(sx); cudaStreamSynchronize(sp); return rc; }
  if (P.dtype == BSQ_DTYPE_I32) {
    subst_i.resize((size_t)L * L);
    for (int i = 0; i < L * L; i++) subst_i[i] = (int)P.subst[i];
    BSQ_CK(ctx, cudaMemcpyAsync(B.subst.p, subst_i.data(), 4 * (size_t)L * L, cudaMemcpyHostToDevice, sp));
  } else {
    BSQ_CK(ctx, cudaMemcpyAsync(B.subst.p, P.subst.data(), 8 * (size_t)L * L, cudaMemcpyHostToDevice, sp));
  }
  rc = stage_k3(ctx, B, P, sp);
  if (rc != BSQ_SUCCESS) { cudaStreamSynchronize(sx); cudaStreamSynchronize(sp); return rc; }
  cudaEvent_t ev_pairs = get_event(ctx, evi++);
  BSQ_CK(ctx, cudaEventRecord(ev_pairs, sp));
  rc = enqueue_slices(1, nsl);
  if (rc != BSQ_SUCCESS) { cudaStreamSynchronize(sx); return rc; }

  // ---- 3. compute stream: chunk after chunk
  cudaEvent_t t0 = get_event(ctx, 0), t1 = get_event(ctx, 1);
  BSQ_CK(ctx, cudaEventRecord(t0, sc));
  BSQ_CK(ctx, cudaMemsetAsync(B.counters.p, 0, 4 * (NC + 1), sc));
  BSQ_CK(ctx, cudaMemsetAsync(B.txlen.p, 0, 4 * N, sc));
  BSQ_CK(ctx, cudaMemsetAsync(B.total.p, 0, 8 * (NC + 2), sc));
  BSQ_CK(ctx, cudaMemsetAsync(B.flags.p, 0, 16, sc));
  BSQ_CK(ctx, cudaStreamWaitEvent(sc, ev_pairs, 0));
  const bool per_chunk_out = P.identity;  // chunk c covers caller indices [idx_lo, idx_hi)
  std::vector<cudaEvent_t> ev_done;
  std::vector<size_t> fill_ev0;
  int waited_slice = -1;
  long long* totals = B.total.as<long long>();  // totals[c] = transcript bytes before chunk c
  rc = run_chunks(
      ctx, B, P, true, want_trace, &evi, &fill_ev0, &ev_done,
      [&](size_t ci) -> int {  // the fill of chunk ci needs its slice of sequence bytes on the device
        const Chunk& ch = P.chunks[ci];
        int need = 0;
        while (need < nsl - 1 && slice_end[need] < ch.code_end) need++;
        for (int sidx = waited_slice + 1; sidx <= need && sidx < nsl; sidx++) BSQ_CK(ctx, cudaStreamWaitEvent(sc, ev_code[sidx], 0));
        waited_slice = std::max(waited_slice, need);
        return launch_check_codes(ctx, B, P, ch.first, ch.count, sc);
      },
      [&](size_t ci, cudaStream_t ts) -> int {  // pack this chunk's transcripts behind its traceback
        if (!(want_tx && per_chunk_out)) return BSQ_SUCCESS;
        const Chunk& ch = P.chunks[ci];
        const int64_t lo = ch.idx_lo, cnt = ch.idx_hi - ch.idx_lo;
        const int nt = (int)((cnt + BSQ_SCAN_TILE - 1) / BSQ_SCAN_TILE);
        long long* tiles = B.tiles.as<long long>() + (lo / BSQ_SCAN_TILE) + ci;  // disjoint scratch per chunk
        BSQ_LAUNCH(bsq_scan_tile_sums, nt, 256, 0, ts, (const int*)(B.txlen.as<int>() + lo), (int)cnt, tiles);
        BSQ_LAUNCH(bsq_scan_tiles, 1, 256, 0, ts, tiles, nt, (const long long*)(totals + ci), totals + ci + 1);
        BSQ_LAUNCH(bsq_scan_apply, nt, 256, 0, ts, (const int*)(B.txlen.as<int>() + lo), (int)cnt,
                   (const long long*)tiles, B.txoff.as<long long>() + lo);
        const dim3 grid((unsigned)((ch.count * 32 + 255) / 256), pack_segments(P));
        BSQ_LAUNCH(bsq_pack_transcripts, grid, 256, 0, ts, (const BsqPair*)(B.pairs.as<BsqPair>() + ch.first),
                   (int)ch.count, (const int*)B.txlen.as<int>(), (const long long*)B.txoff.as<long long>(),
                   (const char*)B.slots.as<char>(), B.packed.as<char>());
        ctx->tm.fill_launches += 4;
        if (cudaGetLastError() != cudaSuccess) return fail(ctx, BSQ_ERR_CUDA, "scan/pack launch failed");
        return BSQ_SUCCESS;
      });
  if (rc != BSQ_SUCCESS) { cudaStreamSynchronize(sc); cudaStreamSynchronize(ctx->st_trace); cudaStreamSynchronize(sx); return rc; }
  if (want_tx && !per_chunk_out) {
    rc = pack_transcripts(ctx, B, P);  // whole batch, caller order
    if (rc != BSQ_SUCCESS) { cudaStreamSynchronize(sc); cudaStreamSynchronize(sx); return rc; }
  }
  BSQ_CK(ctx, cudaEventRecord(t1, sc));

  // ---- 4. D2H stream: ship results back as chunks finish (PCIe is full duplex: the sequence bytes
  // of later chunks are still arriving on the H2D stream)
  cudaStream_t sd = ctx->st_d2h;
  long long* htot = ctx->htotals.as<long long>();
  int rc_cap = BSQ_SUCCESS;
  auto d2h_range = [&](int64_t lo, int64_t cnt) -> int {
    const size_t n4 = 4 * (size_t)cnt, o4 = 4 * (size_t)lo;
    BSQ_CK(ctx, cudaMemcpyAsync(score + lo, (char*)B.score.p + 8 * (size_t)lo, 8 * (size_t)cnt, cudaMemcpyDeviceToHost, sd));
    if (end_x) BSQ_CK(ctx, cudaMemcpyAsync(end_x + lo, (char*)B.ex.p + o4, n4, cudaMemcpyDeviceToHost, sd));
    if (end_y) BSQ_CK(ctx, cudaMemcpyAsync(end_y + lo, (char*)B.ey.p + o4, n4, cudaMemcpyDeviceToHost, sd));
    if (want_trace && start_x) BSQ_CK(ctx, cudaMemcpyAsync(start_x + lo, (char*)B.sx.p + o4, n4, cudaMemcpyDeviceToHost, sd));
    if (want_trace && start_y) BSQ_CK(ctx, cudaMemcpyAsync(start_y + lo, (char*)B.sy.p + o4, n4, cudaMemcpyDeviceToHost, sd));
    BSQ_CK(ctx, cudaMemcpyAsync(status + lo, (char*)B.status.p + o4, n4, cudaMemcpyDeviceToHost, sd));
    ctx->tm.d2h_bytes += cnt * (8 + 4 * (1 + (end_x != 0) + (end_y != 0) + (start_x != 0) + (start_y != 0)));
    return BSQ_SUCCESS;
  };
  if (per_chunk_out) {
    long long shipped = 0, shipped_words = 0;
    for (size_t ci = 0; ci < NC; ci++) {
      const Chunk& ch = P.chunks[ci];
      const int64_t lo = ch.idx_lo, cnt = ch.idx_hi - ch.idx_lo;
      BSQ_CK(ctx, cudaStreamWaitEvent(sd, ev_done[ci], 0));
      rc = d2h_range(lo, cnt);
      if (rc != BSQ_SUCCESS) return rc;
      if (want_tx) {
        BSQ_CK(ctx, cudaMemcpyAsync(tx_off + lo, B.txoff.as<long long>() + lo, 8 * (size_t)cnt, cudaMemcpyDeviceToHost, sd));
        BSQ_CK(ctx, cudaMemcpyAsync(htot + ci + 1, totals + ci + 1, 8, cudaMemcpyDeviceToHost, sd));
        BSQ_CK(ctx, cudaStreamSynchronize(sd));  // later chunks are already queued on the compute stream
        const long long upto = htot[ci + 1];
        ctx->tm.d2h_bytes += 8 * cnt + 8;
        if (tx2) {
          // whole 16-op words only; the word a chunk boundary cuts goes with the next chunk (its ops
          // are packed by then: the chunks' pack kernels run in order on one stream)
          const long long w_hi = (ci + 1 == NC) ? (upto + 15) / 16 : upto / 16;
          if (4 * w_hi > tx_cap) { rc_cap = BSQ_ERR_CAPACITY; }
          else if (rc_cap == BSQ_SUCCESS) {
            rc = ship_tx_2bit(ctx, B, upto, shipped_words, w_hi, tx_buf, sd);
            if (rc != BSQ_SUCCESS) return rc;
            shipped_words = std::max(shipped_words, w_hi);
          }
        } else if (upto > tx_cap) { rc_cap = BSQ_ERR_CAPACITY; }
        else if (upto > shipped) {
          BSQ_CK(ctx, cudaMemcpyAsync(tx_buf + shipped, B.packed.as<char>() + shipped, (size_t)(upto - shipped), cudaMemcpyDeviceToHost, sd));
          ctx->tm.d2h_bytes += upto - shipped;
        }
        shipped = upto;
      }
    }
    if (want_tx) { tx_off[N] = shipped; if (tx_total) *tx_total = shipped; }
  } else {
    BSQ_CK(ctx, cudaStreamWaitEvent(sd, t1, 0));
    rc = d2h_range(0, npairs);
    if (rc != BSQ_SUCCESS) return rc;
    if (want_tx) {
      BSQ_CK(ctx, cudaStreamSynchronize(sd));
      cudaStream_t keep = ctx->st;
      ctx->st = sd;  // fetch_transcripts issues on ctx->st
      rc_cap = fetch_transcripts(ctx, B, P, tx_off, tx_buf, tx_cap, tx_total);
      ctx->st = keep;
      if (rc_cap != BSQ_SUCCESS && rc_cap != BSQ_ERR_CAPACITY) return rc_cap;
    }
  }
  BSQ_CK(ctx, cudaStreamSynchronize(sd));
  BSQ_CK(ctx, cudaStreamSynchronize(sx));
  BSQ_CK(ctx, cudaStreamSynchronize(ctx->st_copy2));
  BSQ_CK(ctx, cudaStreamSynchronize(ctx->st_trace));
  int bad_codes = 0;
  BSQ_CK(ctx, cudaMemcpyAsync(&bad_codes, B.flags.p, 4, cudaMemcpyDeviceToHost, sc));
  BSQ_CK(ctx, cudaStreamSynchronize(sc));
  if (bad_codes) return fail(ctx, BSQ_ERR_ARG, "a letter code is >= alphabet_len");
  collect_timing(ctx, 2, t0, t1);
  // only the (fill start, fill end) event pairs count as fill time
  {
    double fill = 0;
    for (size_t ci = 0; ci < NC; ci++) {
      float ms = 0;
      if (cudaEventElapsedTime(&ms, ctx->events[fill_ev0[ci]], ctx->events[fill_ev0[ci] + 1]) == cudaSuccess) fill += ms;
    }
    ctx->tm.fill_ms = fill;
    if (ctx->hooks.plan_debug) {  // device timeline of the call, relative to its first event
      for (size_t ci = 0; ci < NC; ci++) {
        float a = 0, b = 0, c = 0;
        cudaEventElapsedTime(&a, t0, ctx->events[fill_ev0[ci]]);
        cudaEventElapsedTime(&b, t0, ctx->events[fill_ev0[ci] + 1]);
        if (ev_done.size() > ci && ev_done[ci]) cudaEventElapsedTime(&c, t0, ev_done[ci]);
        fprintf(stderr, "[bsq] chunk %zu: fill %.3f .. %.3f ms, done %.3f ms\n", ci, a, b, c);
      }
    }
  }
  ctx->tm.cells = P.cells;
  ctx->tm.dtype = (!P.chunks.empty() && P.chunks[0].duo) ? BSQ_DTYPE_I16X2 : P.dtype;
  ctx->tm.chunks = (int32_t)NC;
  apply_overrides(P, score, end_x, end_y, want_trace ? start_x : nullptr, want_trace ? start_y : nullptr, status);
  if (rc_cap == BSQ_ERR_CAPACITY) return fail(ctx, BSQ_ERR_CAPACITY, "tx_buf too small for the batch's transcripts");
  ctx->tm.call_ms = now_ms() - t_call;
  return BSQ_SUCCESS;
}

int bsq_batch_upload(bsq_ctx* ctx, const bsq_params* params, int64_t npairs, const uint8_t* codes,
                     int64_t ncodes, const int64_t* o_off, const int32_t* o_len, const int64_t* m_off,
                     const int32_t* m_len, const int32_t* dmin, const int32_t* dmax, int32_t keep_tables,
                     bsq_batch** batch) {
  if (!ctx) return BSQ_ERR_ARG;
  ctx->err.clear();
  if (!batch) return fail(ctx, BSQ_ERR_ARG, "batch out-pointer is NULL");
  *batch = nullptr;
  int rc = check_inputs(ctx, npairs, codes, o_off, o_len, m_off, m_len);
  if (rc != BSQ_SUCCESS) return rc;
  BSQ_CK(ctx, cudaSetDevice(ctx->device));
  memset(&ctx->tm, 0, sizeof ctx->tm);
  bsq_batch* b = new bsq_batch();
  rc = batch_upload(ctx, b, true, keep_tables != 0, params, npairs, codes, ncodes, o_off, o_len, m_off, m_len, dmin, dmax);
  if (rc == BSQ_ERR_NOMEM) {  // stale free-memory figure or hoarded buffer sets: ask the driver and try once more
    cudaStreamSynchronize(ctx->st);
    drop_pool(ctx);
    ctx->mem_avail = -1.0;
    rc = batch_upload(ctx, b, true, keep_tables != 0, params, npairs, codes, ncodes, o_off, o_len, m_off, m_len, dmin, dmax);
  }
  if (rc == BSQ_SUCCESS && npairs > 0) {
    int bad = 0;
    cudaError_t e = cudaMemcpyAsync(&bad, b->B->flags.p, 4, cudaMemcpyDeviceToHost, ctx->st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->st);
    if (e != cudaSuccess) { ctx->err = std::string("upload: ") + cudaGetErrorString(e); rc = BSQ_ERR_CUDA; }
    else if (bad) rc = fail(ctx, BSQ_ERR_ARG, "a letter code is >= alphabet_len");
  }
  if (rc != BSQ_SUCCESS) { cudaStreamSynchronize(ctx->st); bsq_release(b); return rc; }
  *batch = b;
  return BSQ_SUCCESS;
}

int bsq_batch_run(bsq_batch* b, int32_t with_traceback) {
  if (!b || !b->ctx) return BSQ_ERR_ARG;
  b->ctx->err.clear();
  BSQ_CK(b->ctx, cudaSetDevice(b->ctx->device));
  const int64_t h2d = b->ctx->tm.h2d_bytes;
  memset(&b->ctx->tm, 0, sizeof b->ctx->tm);
  b->ctx->tm.h2d_bytes = h2d;
  return batch_run(b, true, with_traceback != 0);
}

int bsq_batch_fetch(bsq_batch* b, double* score, int32_t* end_x, int32_t* end_y, int32_t* start_x,
                    int32_t* start_y, int32_t* status, int64_t* tx_off, char* tx_buf, int64_t tx_cap,
                    int64_t* tx_total) {
  if (!b || !b->ctx) return BSQ_ERR_ARG;
  b->ctx->err.clear();
  if (!b->filled) return fail(b->ctx, BSQ_ERR_ARG, "bsq_batch_fetch before bsq_batch_run");
  BSQ_CK(b->ctx, cudaSetDevice(b->ctx->device));
  return batch_fetch(b, score, end_x, end_y, start_x, start_y, status, tx_off, tx_buf, tx_cap, tx_total);
}

int bsq_solve_batch(bsq_ctx* ctx, const bsq_params* params, int64_t npairs, const uint8_t* codes,
                    int64_t ncodes, const int64_t* o_off, const int32_t* o_len, const int64_t* m_off,
                    const int32_t* m_len, const int32_t* dmin, const int32_t* dmax, double* score,
                    int32_t* end_x, int32_t* end_y, int32_t* status, bsq_batch** batch) {
  if (!ctx) return BSQ_ERR_ARG;
  if (!score || !status) return fail(ctx, BSQ_ERR_ARG, "score/status must not be NULL");
  int rc = bsq_batch_upload(ctx, params, npairs, codes, ncodes, o_off, o_len, m_off, m_len, dmin, dmax, 1, batch);
  if (rc != BSQ_SUCCESS) return rc;
  rc = batch_run(*batch, true, false);
  if (rc == BSQ_SUCCESS) rc = batch_fetch(*batch, score, end_x, end_y, nullptr, nullptr, status, nullptr, nullptr, 0, nullptr);
  if (rc != BSQ_SUCCESS) { bsq_release(*batch); *batch = nullptr; }
  return rc;
}

int bsq_traceback_batch(bsq_batch* b, int32_t* start_x, int32_t* start_y, int32_t* status, int64_t* tx_off,
                        char* tx_buf, int64_t tx_cap, int64_t* tx_total) {
  if (!b || !b->ctx) return BSQ_ERR_ARG;
  bsq_ctx* ctx = b->ctx;
  ctx->err.clear();
  if (!tx_off || !tx_buf) return fail(ctx, BSQ_ERR_ARG, "tx_off/tx_buf must not be NULL");
  if (!b->filled && b->plan.N > 0) return fail(ctx, BSQ_ERR_ARG, "bsq_traceback_batch before solve");
  BSQ_CK(ctx, cudaSetDevice(ctx->device));
  int rc = BSQ_SUCCESS;
  if (!b->traced) rc = batch_run(b, false, true);
  if (rc != BSQ_SUCCESS) return rc;
  return batch_fetch(b, nullptr, nullptr, nullptr, start_x, start_y, status, tx_off, tx_buf, tx_cap, tx_total);
}

void bsq_release(bsq_batch* b) {
  if (!b) return;
  bsq_ctx* ctx = b->ctx;
  if (ctx) { cudaSetDevice(ctx->device); cudaStreamSynchronize(ctx->st); cudaStreamSynchronize(ctx->st_trace); }
  if (b->own) {
    if (ctx) {
      for (size_t i = 0; i < ctx->live.size(); i++) if (ctx->live[i] == b) { ctx->live.erase(ctx->live.begin() + i); break; }
      // keep up to two modest buffer sets for the next batch; big ones go back to the device
      if (ctx->pool.size() < 2 && b->own->bytes() <= ((size_t)1 << 30)) { ctx->pool.push_back(b->own); b->own = nullptr; }
    }
    if (b->own) { b->own->release(); delete b->own; }
  }
  delete b;
}

// ---------------------------------------------------------------------------
// bsq_multi: one batch over several GPUs
// ---------------------------------------------------------------------------
}  // extern "C"

struct bsq_multi {
  struct Shard {
    bsq_ctx* ctx = nullptr;
    int device = 0;
    int64_t first = 0, count = 0;
    PinBuf txoff;          // the shard's own tx_off (count + 1 entries), rebased into the caller's array afterwards
    int rc = BSQ_SUCCESS;
    bsq_timing tm;
  };
  std::vector<Shard> shards;
  std::string err;
};

extern "C" {

bsq_multi* bsq_multi_create(int32_t ndev, const int32_t* devices) {
  if (ndev < 1 || ndev > 64) { g_err = "bsq_multi_create: need 1 <= ndev <= 64"; return nullptr; }
  bsq_multi* me = new bsq_multi();
  me->shards.resize((size_t)ndev);
  for (int k = 0; k < ndev; k++) {
    const int dev = devices ? devices[k] : k;
    me->shards[(size_t)k].device = dev;
    me->shards[(size_t)k].ctx = bsq_create(dev);
    memset(&me->shards[(size_t)k].tm, 0, sizeof(bsq_timing));
    if (!me->shards[(size_t)k].ctx) { bsq_multi_destroy(me); return nullptr; }   // g_err holds the reason
  }
  return me;
}

void bsq_multi_destroy(bsq_multi* me) {
  if (!me) return;
  for (auto& sh : me->shards) {
    if (sh.ctx) { cudaSetDevice(sh.device); sh.txoff.release(); bsq_destroy(sh.ctx); }
  }
  delete me;
}

int32_t bsq_multi_size(const bsq_multi* me) { return me ? (int32_t)me->shards.size() : 0; }
const char* bsq_multi_last_error(const bsq_multi* me) { return me ? me->err.c_str() : g_err.c_str(); }

int bsq_multi_get_shard(const bsq_multi* me, int32_t k, int64_t* first, int64_t* count, int32_t* device, bsq_timing* timing) {
  if (!me || k < 0 || k >= (int32_t)me->shards.size()) return BSQ_ERR_ARG;
  const bsq_multi::Shard& sh = me->shards[(size_t)k];
  if (first) *first = sh.first;
  if (count) *count = sh.count;
  if (device) *device = sh.device;
  if (timing) *timing = sh.tm;
  return BSQ_SUCCESS;
}

int bsq_multi_align_batch(bsq_multi* me, const bsq_params* params, int64_t npairs, const uint8_t* codes,
                          int64_t ncodes, const int64_t* o_off, const int32_t* o_len, const int64_t* m_off,
                          const int32_t* m_len, const int32_t* dmin, const int32_t* dmax, double* score,
                          int32_t* end_x, int32_t* end_y, int32_t* start_x, int32_t* start_y, int32_t* status,
                          int64_t* tx_off, int32_t* tx_len, char* tx_buf, int64_t tx_cap) {
  if (!me) return BSQ_ERR_ARG;
  me->err.clear();
  auto bad = [&](int code, const std::string& msg) { me->err = msg; g_err = msg; return code; };
  if (!params || npairs < 0 || npairs >= ((int64_t)1 << 31)) return bad(BSQ_ERR_ARG, "bad argument");
  if (npairs > 0 && (!codes || !o_len || !m_len || !score || !status)) return bad(BSQ_ERR_ARG, "NULL array");
  if ((o_off == nullptr) != (m_off == nullptr)) return bad(BSQ_ERR_ARG, "o_off and m_off must both be given or both be NULL");
  if (params->mode == BSQ_BANDED_MODE && (!dmin || !dmax)) return bad(BSQ_ERR_ARG, "BANDED needs dmin/dmax");
  const bool want_tx = tx_off && tx_buf;
  if (want_tx && !tx_len) return bad(BSQ_ERR_ARG, "bsq_multi_align_batch needs tx_len next to tx_off/tx_buf");
  const int K = (int)me->shards.size();
  const bool tx2 = (params->flags & BSQ_F_TX_2BIT) != 0, c2 = (params->flags & BSQ_F_CODES_2BIT) != 0;
  for (auto& sh : me->shards) { sh.first = sh.count = 0; sh.rc = BSQ_SUCCESS; memset(&sh.tm, 0, sizeof sh.tm); }
  if (npairs == 0) return BSQ_SUCCESS;

  // ---- cut [0, npairs) into K contiguous ranges of equal work.  Work of a pair ~ its DP cells
  // ((n+1)(m+1), or band width x shorter side); letters[] = prefix of n + m (windows of codes[] for
  // back-to-back batches, worst-case transcript room for every batch)
  const int T = parallel_threads(npairs);
  std::vector<double> pw((size_t)T + 1, 0.0);
  std::vector<int64_t> pl((size_t)T + 1, 0);
  auto weight = [&](int64_t i) -> double {
    const double n = o_len[i], m = m_len[i];
    if (params->mode == BSQ_STD_MODE) return (n + 1) * (m + 1);
    const double w = std::max(1.0, std::min<double>(dmax[i], n) - std::max<double>(dmin[i], -m) + 1);
    return w * (std::min(n, m) + 1);
  };
  parallel_for(npairs, [&](int t, int64_t i0, int64_t i1) {
    double w = 0; int64_t l = 0;
    for (int64_t i = i0; i < i1; i++) { w += weight(i); l += (int64_t)o_len[i] + m_len[i]; }
    pw[(size_t)t + 1] = w; pl[(size_t)t + 1] = l;
  });
  for (int t = 0; t < T; t++) { pw[(size_t)t + 1] += pw[(size_t)t]; pl[(size_t)t + 1] += pl[(size_t)t]; }
  const double total_w = pw[(size_t)T];
  std::vector<int64_t> cut((size_t)K + 1, npairs), let_at((size_t)K + 1, pl[(size_t)T]);
  cut[0] = 0; let_at[0] = 0;
  {
    int k = 1;
    for (int t = 0; t < T && k < K; t++) {   // walk the thread ranges that hold a boundary
      const int64_t i0 = npairs * t / T, i1 = npairs * (t + 1) / T;
      if (pw[(size_t)t + 1] < total_w * k / K) continue;
      double w = pw[(size_t)t]; int64_t l = pl[(size_t)t];
      for (int64_t i = i0; i < i1 && k < K; i++) {
        while (k < K && w >= total_w * k / K) { cut[(size_t)k] = i; let_at[(size_t)k] = l; k++; }
        w += weight(i); l += (int64_t)o_len[i] + m_len[i];
      }
    }
  }
  // transcript regions (in ops): R_0 = 0, R_{k+1} = R_k + roundup64(worst case of shard k)
  std::vector<int64_t> R((size_t)K + 1, 0);
  for (int k = 0; k < K; k++) R[(size_t)k + 1] = R[(size_t)k] + ((let_at[(size_t)k + 1] - let_at[(size_t)k] + 63) / 64) * 64;
  if (want_tx && (tx2 ? R[(size_t)K] / 4 : R[(size_t)K]) > tx_cap)
    return bad(BSQ_ERR_CAPACITY, "tx_buf must hold the worst case: sum(o_len + m_len) + 64 * bsq_multi_size() ops");

  auto run_shard = [&](int k) {
    bsq_multi::Shard& sh = me->shards[(size_t)k];
    const int64_t b = cut[(size_t)k], e = cut[(size_t)k + 1], n = e - b;
    sh.first = b; sh.count = n;
    if (n <= 0) return;
    g_plan_serial = true;
    cudaSetDevice(sh.device);
    // the window of codes[] this shard's pairs reference, starting on a 4-letter boundary
    int64_t lo, hi;
    if (!o_off) { lo = let_at[(size_t)k]; hi = let_at[(size_t)k + 1]; }
    else {
      lo = INT64_MAX; hi = 0;
      for (int64_t i = b; i < e; i++) {
        lo = std::min(lo, std::min(o_off[i], m_off[i]));
        hi = std::max(hi, std::max(o_off[i] + o_len[i], m_off[i] + m_len[i]));
      }
      if (lo < 0 || hi > ncodes) { sh.rc = BSQ_ERR_ARG; sh.ctx->err = "sequence range outside codes[]"; return; }
    }
    const int64_t ws = lo & ~(int64_t)3;
    const uint8_t* cw = codes + (c2 ? ws / 4 : ws);
    const int64_t base = o_off ? -ws : lo - ws;
    int64_t* loc = nullptr;
    if (want_tx) {
      if (!sh.txoff.ensure(8 * (size_t)(n + 1))) { sh.rc = BSQ_ERR_NOMEM; sh.ctx->err = "pinned host allocation failed"; return; }
      loc = sh.txoff.as<int64_t>();
    }
    const int64_t rcap = R[(size_t)k + 1] - R[(size_t)k];
    int64_t total = 0;
    sh.rc = align_batch_retry(sh.ctx, params, n, cw, hi - ws, o_off ? o_off + b : nullptr, o_len + b,
                              m_off ? m_off + b : nullptr, m_len + b, dmin ? dmin + b : nullptr, dmax ? dmax + b : nullptr,
                              score + b, end_x ? end_x + b : nullptr, end_y ? end_y + b : nullptr,
                              start_x ? start_x + b : nullptr, start_y ? start_y + b : nullptr, status + b, loc,
                              want_tx ? tx_buf + (tx2 ? R[(size_t)k] / 4 : R[(size_t)k]) : nullptr,
                              tx2 ? rcap / 4 : rcap, &total, base);
    sh.tm = sh.ctx->tm;
    if (sh.rc == BSQ_SUCCESS && want_tx) {
      for (int64_t i = 0; i < n; i++) { tx_off[b + i] = R[(size_t)k] + loc[i]; tx_len[b + i] = (int32_t)(loc[i + 1] - loc[i]); }
    }
    g_plan_serial = false;
  };
#ifdef BSQ_SIMT_EMU
  for (int k = 0; k < K; k++) run_shard(k);      // the emulator's fibers are not thread safe
#else
  {
    std::vector<std::thread> th;
    for (int k = 1; k < K; k++) th.emplace_back(run_shard, k);
    run_shard(0);
    for (auto& t : th) t.join();
  }
#endif
  if (want_tx) tx_off[npairs] = R[(size_t)K];
  for (int k = 0; k < K; k++) {
    const bsq_multi::Shard& sh = me->shards[(size_t)k];
    if (sh.rc != BSQ_SUCCESS)
      return bad(sh.rc, "shard " + std::to_string(k) + " (device " + std::to_string(sh.device) + "): " + sh.ctx->err);
  }
  return BSQ_SUCCESS;
}

#ifdef BSQ_K3_PROFILE
// tuning builds only (not declared in bsqgpu.h): raw per-strip records of the last K3 launch
int bsq_debug_k3_res(bsq_batch* b, void* out, int64_t bytes) {
  if (!b || !b->ctx || !b->B->k3res.p) return BSQ_ERR_ARG;
  if ((size_t)bytes > b->B->k3res.cap) bytes = (int64_t)b->B->k3res.cap;
  return cudaMemcpy(out, b->B->k3res.p, (size_t)bytes, cudaMemcpyDeviceToHost) == cudaSuccess ? (int)sizeof(BsqK3Res) : BSQ_ERR_CUDA;
}
#endif

int bsq_microbench(bsq_ctx* ctx, int32_t which, double* ops_per_sec) {
  if (!ctx || !ops_per_sec) return BSQ_ERR_ARG;
  ctx->err.clear();
  BSQ_CK(ctx, cudaSetDevice(ctx->device));
  return bsq_microbench_impl(ctx->sms, ctx->st, which, ops_per_sec, &ctx->err);
}

int bsq_generate_pairs(bsq_ctx* ctx, const bsq_gen_params* gp, int64_t npairs, uint8_t* codes, int64_t codes_cap,
                       int64_t* ncodes, int64_t* o_off, int32_t* o_len, int64_t* m_off, int32_t* m_len,
                       int32_t* src_off, int64_t* tx_off, char* tx_buf, int64_t tx_cap) {
  if (!ctx) return BSQ_ERR_ARG;
  ctx->err.clear();
  if (!gp || !ncodes || npairs < 0 || npairs >= ((int64_t)1 << 31)) return fail(ctx, BSQ_ERR_ARG, "bad argument");
  if (npairs > 0 && (!codes || !o_off || !o_len || !m_off || !m_len)) return fail(ctx, BSQ_ERR_ARG, "NULL output array");
  if ((tx_off == nullptr) != (tx_buf == nullptr)) return fail(ctx, BSQ_ERR_ARG, "tx_off and tx_buf go together");
  if (gp->alphabet_len < 1 || gp->alphabet_len > 255 || gp->origin_len < 0 || gp->max_offset < 0 || gp->fixed_mutant_len < 0)
    return fail(ctx, BSQ_ERR_ARG, "bad generator parameters");
  if (!(gp->subst_prob >= 0 && gp->subst_prob < 1 && gp->go_prob >= 0 && gp->ge_prob < 1 && gp->go_prob <= gp->ge_prob))
    return fail(ctx, BSQ_ERR_ARG, "need 0 <= subst_prob < 1 and 0 <= go_prob <= ge_prob < 1 (stochastics.py:128-141)");
  *ncodes = 0;
  if (tx_off) tx_off[0] = 0;
  if (npairs == 0) return BSQ_SUCCESS;
  BSQ_CK(ctx, cudaSetDevice(ctx->device));
  BsqGenArgs G;
  G.seed = gp->seed;
  G.L = gp->alphabet_len; G.n = gp->origin_len; G.max_offset = gp->max_offset; G.fixed_mlen = gp->fixed_mutant_len;
  G.subst_thr = (unsigned)std::floor(gp->subst_prob * 16777216.0);
  G.go_thr = (unsigned)std::floor(gp->go_prob * 16777216.0);
  G.ge_thr = (unsigned)std::floor(gp->ge_prob * 16777216.0);
  G.npairs = npairs;
  const size_t N = (size_t)npairs;
  cudaStream_t st = ctx->st;
  DevBuf dMlen, dTxlen, dOff, dBase, dTxbase, dCodes, dTx;
  auto cleanup = [&] { dMlen.release(); dTxlen.release(); dOff.release(); dBase.release(); dTxbase.release(); dCodes.release(); dTx.release(); };
  if (!dMlen.ensure(4 * N) || !dTxlen.ensure(4 * N) || !dOff.ensure(4 * N) || !dBase.ensure(8 * (N + 1)) || !dTxbase.ensure(8 * (N + 1))) {
    cleanup();
    return fail(ctx, BSQ_ERR_NOMEM, "device allocation failed (generator)");
  }
  const int gridp = (int)((npairs + 127) / 128);
  BSQ_LAUNCH(bsq_gen_count, gridp, 128, 0, st, G, dMlen.as<int>(), dTxlen.as<int>(), dOff.as<int>());
  std::vector<int32_t> txlen(N);
  std::vector<int32_t> offs(N);
  cudaError_t e = cudaGetLastError();
  if (e == cudaSuccess) e = cudaMemcpyAsync(m_len, dMlen.p, 4 * N, cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaMemcpyAsync(txlen.data(), dTxlen.p, 4 * N, cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaMemcpyAsync(offs.data(), dOff.p, 4 * N, cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  if (e != cudaSuccess) { cleanup(); ctx->err = std::string("generator: ") + cudaGetErrorString(e); return BSQ_ERR_CUDA; }
  std::vector<int64_t> base(N + 1), txbase(N + 1);
  int64_t pos = 0, tpos = 0;
  for (size_t i = 0; i < N; i++) {
    base[i] = pos; txbase[i] = tpos;
    o_off[i] = pos; o_len[i] = gp->origin_len;
    m_off[i] = pos + gp->origin_len;
    pos += (int64_t)gp->origin_len + m_len[i];
    tpos += txlen[i];
    if (src_off) src_off[i] = offs[i];
    if (tx_off) tx_off[i + 1] = tpos;
  }
  base[N] = pos; txbase[N] = tpos;
  *ncodes = pos + 16;
  if (pos + 16 > codes_cap || (tx_buf && tpos > tx_cap)) { cleanup(); return fail(ctx, BSQ_ERR_CAPACITY, "codes / tx_buf too small for the generated pairs"); }
  if (!dCodes.ensure((size_t)pos + 16) || (tx_buf && !dTx.ensure((size_t)std::max<int64_t>(tpos, 1)))) {
    cleanup();
    return fail(ctx, BSQ_ERR_NOMEM, "device allocation failed (generator output)");
  }
  e = cudaMemcpyAsync(dBase.p, base.data(), 8 * (N + 1), cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess) e = cudaMemcpyAsync(dTxbase.p, txbase.data(), 8 * (N + 1), cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess) e = cudaMemsetAsync((char*)dCodes.p + pos, 0, 16, st);
  if (e == cudaSuccess) {
    const long long total = (long long)npairs * gp->origin_len;
    const int grido = (int)std::max<long long>(1, std::min<long long>((total + 255) / 256, (long long)ctx->sms * 32));
    BSQ_LAUNCH(bsq_gen_origins, grido, 256, 0, st, G, (const long long*)dBase.as<long long>(), dCodes.as<uint8_t>());
    BSQ_LAUNCH(bsq_gen_write, gridp, 128, 0, st, G, (const long long*)dBase.as<long long>(), (const int*)dOff.as<int>(),
               dCodes.as<uint8_t>(), (const long long*)dTxbase.as<long long>(), tx_buf ? dTx.as<char>() : (char*)nullptr);
    e = cudaGetLastError();
  }
  if (e == cudaSuccess) e = cudaMemcpyAsync(codes, dCodes.p, (size_t)pos + 16, cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess && tx_buf && tpos > 0) e = cudaMemcpyAsync(tx_buf, dTx.p, (size_t)tpos, cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  cleanup();
  if (e != cudaSuccess) { ctx->err = std::string("generator: ") + cudaGetErrorString(e); return BSQ_ERR_CUDA; }
  return BSQ_SUCCESS;
}

int bsq_batch_stats(bsq_batch* b, bsq_aln_stats* out) {
  if (!b || !b->ctx) return BSQ_ERR_ARG;
  bsq_ctx* ctx = b->ctx;
  ctx->err.clear();
  if (!out) return fail(ctx, BSQ_ERR_ARG, "out must not be NULL");
  const Plan& P = b->plan;
  if (P.N == 0) return BSQ_SUCCESS;
  if (!b->traced) return fail(ctx, BSQ_ERR_ARG, "bsq_batch_stats needs a traced batch (run with traceback first)");
  BSQ_CK(ctx, cudaSetDevice(ctx->device));
  BufSet& B = *b->B;
  const int L = P.prm.alphabet_len;
  const size_t N = (size_t)P.N;
  DevBuf dS, dOut;
  if (!dS.ensure(8 * (size_t)L * L) || !dOut.ensure(sizeof(bsq_aln_stats) * N)) {
    dS.release(); dOut.release();
    return fail(ctx, BSQ_ERR_NOMEM, "device allocation failed (stats)");
  }
  cudaStream_t st = ctx->st;
  cudaError_t e = cudaMemcpyAsync(dS.p, P.subst.data(), 8 * (size_t)L * L, cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess) e = cudaMemsetAsync(dOut.p, 0, sizeof(bsq_aln_stats) * N, st);
  if (e == cudaSuccess && P.E > 0) {
    const int grid = (int)((P.E + 127) / 128);
    BSQ_LAUNCH(bsq_k5_stats, grid, 128, 0, st, (const BsqPair*)B.pairs.as<BsqPair>(), (int)P.E,
               (const uint8_t*)B.codes.as<uint8_t>(), (const long long*)B.txoff.as<long long>(),
               (const int*)B.txlen.as<int>(), (const char*)B.packed.as<char>(), (const int*)B.sx.as<int>(),
               (const int*)B.sy.as<int>(), (const int*)B.status.as<int>(), (const double*)dS.as<double>(), L,
               P.prm.go, P.prm.ge, dOut.as<bsq_aln_stats>());
    e = cudaGetLastError();
  }
  if (e == cudaSuccess) e = cudaMemcpyAsync(out, dOut.p, sizeof(bsq_aln_stats) * N, cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  dS.release(); dOut.release();
  if (e != cudaSuccess) { ctx->err = std::string("stats: ") + cudaGetErrorString(e); return BSQ_ERR_CUDA; }
  for (const Override& o : P.overrides) {  // pairs that never reached the device
    bsq_aln_stats z;
    memset(&z, 0, sizeof z);
    z.trunc_lo = z.trunc_hi = z.trunc_start_x = z.trunc_start_y = -1;
    out[o.idx] = z;
  }
  return BSQ_SUCCESS;
}

int bsq_table_scores(bsq_ctx* ctx, const bsq_params* params, const uint8_t* origin, int32_t n,
                     const uint8_t* mutant, int32_t m, double* out) {
  if (!ctx) return BSQ_ERR_ARG;
  ctx->err.clear();
  if (!params || !params->subst || !out || n < 0 || m < 0) return fail(ctx, BSQ_ERR_ARG, "bad argument");
  if (params->mode != BSQ_STD_MODE) return fail(ctx, BSQ_ERR_ARG, "table_scores is STD mode only (pw.py:280-281)");
  const int L = params->alphabet_len;
  if (L < 1 || L > BSQ_MAX_L) return fail(ctx, BSQ_ERR_ARG, "alphabet_len must be in [1,32]");
  BSQ_CK(ctx, cudaSetDevice(ctx->device));
  int bmode, emode;
  std_modes(params->type, &bmode, &emode);
  const size_t cells = (size_t)(n + 1) * (size_t)(m + 1);
  DevBuf dH, dM, dO, dT, dS;
  if (!dH.ensure(8 * cells) || !dM.ensure(cells) || !dO.ensure((size_t)n + 1) || !dT.ensure((size_t)m + 1) ||
      !dS.ensure(8 * (size_t)L * L)) {
    dH.release(); dM.release(); dO.release(); dT.release(); dS.release();
    return fail(ctx, BSQ_ERR_NOMEM, "device allocation failed (table_scores)");
  }
  cudaStream_t st = ctx->st;
  if (n) cudaMemcpyAsync(dO.p, origin, (size_t)n, cudaMemcpyHostToDevice, st);
  if (m) cudaMemcpyAsync(dT.p, mutant, (size_t)m, cudaMemcpyHostToDevice, st);
  cudaMemcpyAsync(dS.p, params->subst, 8 * (size_t)L * L, cudaMemcpyHostToDevice, st);
  const int gos = params->go < 0 ? -1 : (params->go > 0 ? 1 : 0);
  if (gos < 0)
    BSQ_LAUNCH(bsq_table_scores_kernel<-1>, 1, 256, 0, st, (const uint8_t*)dO.as<uint8_t>(), (int)n,
               (const uint8_t*)dT.as<uint8_t>(), (int)m, (const double*)dS.as<double>(), L, params->go,
               params->ge, bmode, dH.as<double>(), dM.as<uint8_t>());
  else if (gos == 0)
    BSQ_LAUNCH(bsq_table_scores_kernel<0>, 1, 256, 0, st, (const uint8_t*)dO.as<uint8_t>(), (int)n,
               (const uint8_t*)dT.as<uint8_t>(), (int)m, (const double*)dS.as<double>(), L, params->go,
               params->ge, bmode, dH.as<double>(), dM.as<uint8_t>());
  else
    BSQ_LAUNCH(bsq_table_scores_kernel<1>, 1, 256, 0, st, (const uint8_t*)dO.as<uint8_t>(), (int)n,
               (const uint8_t*)dT.as<uint8_t>(), (int)m, (const double*)dS.as<double>(), L, params->go,
               params->ge, bmode, dH.as<double>(), dM.as<uint8_t>());
  cudaError_t e = cudaGetLastError();
  if (e == cudaSuccess) e = cudaMemcpyAsync(out, dH.p, 8 * cells, cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  dH.release(); dM.release(); dO.release(); dT.release(); dS.release();
  if (e != cudaSuccess) { ctx->err = std::string("table_scores: ") + cudaGetErrorString(e); return BSQ_ERR_CUDA; }
  return BSQ_SUCCESS;
}

}  // extern "C"
