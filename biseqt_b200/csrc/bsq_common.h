// bsq_common.h -- shared host/device declarations of libbsqgpu.
#pragma once
#include <stdint.h>

#ifdef BSQ_SIMT_EMU
// tests/simt_emu: the same sources compiled for a CPU fiber emulator so that
// kernel logic can be checked without a GPU.  Never part of the product .so.
#include "simt_emu.h"
#define BSQ_LAUNCH(kern, grid, block, smem, stream, ...) \
  emu::launch(kern, dim3(grid), dim3(block), (size_t)(smem), __VA_ARGS__)
#define BSQ_HD
#else
#include <cuda_runtime.h>
#define BSQ_LAUNCH(kern, grid, block, smem, stream, ...) \
  kern<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__)
#define BSQ_HD __host__ __device__
#endif

// tied-max mask bits, in the reference's candidate order (pw.c:77-80)
#define BSQ_MB 1u
#define BSQ_MD 2u
#define BSQ_MI 4u
#define BSQ_MM 8u

// Words per lane per row of the K1 pointer layout, rounded up to a power of two: a lane then
// writes its row with ONE aligned vector store and a group's row is a whole number of full 32-byte
// sectors.  (Three separate 4-byte stores per lane put 8x the bytes on the L1->L2 path as partial
// sectors and cost more than the fill's arithmetic: profiles/r01_k1h_store_experiment.txt.)
#define BSQ_K1_WPL(K) ((((K) + 7) / 8) == 3 ? 4 : (((K) + 7) / 8))

#define BSQ_MAX_L 32  // largest alphabet the kernels stage in shared memory

// K2 pointer layout: [block of T steps][lane][32-byte unit].  A lane's nibbles of T consecutive
// anti-diagonals (WPS = ceil(C/16) words per step, T = 8 / WPS steps) fill ONE 32-byte sector,
// written with two 16-byte stores per block and read by the traceback as one sector per ~T/2 ops.
// (One word per lane and step in step-major order -- the first layout -- made every op of every walk
// a different DRAM sector: 14 ms of traceback behind a 24 ms fill on config 3, whether a thread or 8
// lanes walked a pair.)  Nibble of cell hc at step i of the block: word i * WPS + hc / 8, bits 4 (hc % 8).
#define BSQ_K2_UNIT_WORDS 8
#define BSQ_K2_WPS(C) (((C) / 2 + 7) / 8)
#define BSQ_K2_T(C) (BSQ_K2_UNIT_WORDS / BSQ_K2_WPS(C))

// where B (start of alignment) is allowed: _alnchoice_B, _pw_internals.c:161-209
enum { BSQ_B_ORIGIN = 0, BSQ_B_EDGE = 1, BSQ_B_ALL = 2 };
// where an alignment may end: _std_find_optimal / _banded_find_optimal, :303-414
enum { BSQ_E_CORNER = 0, BSQ_E_EDGE = 1, BSQ_E_ALL = 2 };
// sign of gap_open_score decides the D/I tie rules (_pw_internals.c:266-278)
enum { BSQ_GO_NEG = -1, BSQ_GO_ZERO = 0, BSQ_GO_POS = 1 };

struct BsqPair {         // one problem of a chunk, device resident
  long long o_off;       // frame start of origin in codes[]
  long long m_off;       // frame start of mutant in codes[]
  long long ptr_off;     // first uint32 word of this pair's pointer table
  long long tx_slot;     // first byte of this pair's transcript slot (cap n+m)
  int n, m;              // frame lengths
  int dmin, dmax;        // clipped band (BANDED)
  int s0;                // K2: first anti-diagonal x+y the kernel visits; K3: time steps per strip
  int pad;
};

// every pair of the batch has this shape (uniform batches: records are built on the device)
struct BsqUniform {
  int n, m, dmin, dmax, s0;
  long long words;       // pointer-table words per pair
  long long ppc;         // pairs per chunk (ptr_off restarts at every chunk)
  long long off_base;    // added to every caller offset (a shard's window of codes[]); with implicit
                         // back-to-back offsets: where the first pair starts
};

// K1h: the two problems (indices into the chunk's pair records, equal shapes; b == a for a
// problem without a partner) that share a lane group

// K3 (strip-pipelined wide bands): what a pair needs beyond BsqPair
struct BsqK3Pair {
  long long slab_off;    // first entry of this pair's boundary-column slab (n+1 entries)
  int strip_base;        // index of this pair's strip 0 in the chunk's per-strip result array
  int nstrips;           // ceil((m+1) / (32*K))
};
struct BsqK3Item {       // one entry of a chunk's strip work list: strip `strip` of the chunk's pair `pair`
  int pair, strip;
};
struct BsqK3Res {        // best end-cell candidate found inside one strip
  double score;
  int k1, k2;            // scan keys (diagonal, x); k1 == INT_MAX: none
#ifdef BSQ_K3_PROFILE       // tuning builds only: when the strip was claimed / could start / ended
  long long t_claim, t_ready, t_end;   // %globaltimer, ns
  long long wait_clk;                  // SM clocks spent waiting for boundary batches
  int polls, sm;
#endif
};

struct BsqFillArgs {
  const uint8_t* codes;
  const BsqPair* pairs;
  int npairs;
  int L;
  const void* subst;     // L*L scores already converted to the kernel's type
  double go_d, ge_d;
  int go_i, ge_i;
  int bmode, emode;      // STD: BSQ_B_* / BSQ_E_* of the alignment type
  int btype;             // BANDED: banded_alntype
  uint32_t* ptr;         // pointer-table arena
  double* out_score;
  int* out_ex;
  int* out_ey;
  int* out_status;
  int* counter;          // dynamic work queue head
  const int* bad;        // != 0: bsq_check_codes found a letter >= L; the kernel must not run
  // STD multi-pass boundary column scratch, one slab per resident group
  double* col_h;         // (also used for int scores, stored as double-sized slots)
  uint8_t* col_f;
  long long col_stride;  // entries per group
  // K1h: lane groups to serve; group q takes problems (2q, 2q+1) of the chunk
  int nduos;
  // K3: strips of all pairs of the chunk form one ordered work list; col_h holds one
  // boundary-column slab per pair (BsqK3Pair.slab_off rows of 1 (I32) / 2 (F64) 64-bit
  // self-validating entries), zeroed before the launch
  const BsqK3Pair* k3;
  const BsqK3Item* k3_order;  // [k3_total] the work list: strips of all pairs interleaved (strip j of every
                              // pair before strip j+1 of any), so that all pairs' pipelines run at once
  int k3_total;          // strips in the chunk
  int* k3_done;          // [npairs] strips finished per pair, zeroed before the launch
  BsqK3Res* k3_res;      // [k3_total]
};

struct BsqTraceArgs {
  const uint8_t* codes;
  const BsqPair* pairs;
  int npairs;
  int mode;              // BSQ_STD_MODE / BSQ_BANDED_MODE
  int G, K;              // geometry of the fill kernel that wrote the tables
  int gosign;
  const uint32_t* ptr;
  const int* ex;
  const int* ey;
  int* status;           // in: OK / NO_ALIGNMENT from the fill; out: final class
  int* sx;
  int* sy;
  int* txlen;
  char* slots;           // transcripts, right-aligned in [tx_slot, tx_slot+n+m)
  int family;            // 1 = K1, 2 = K2, 3 = K3 table layout
  const int* bad;        // != 0: the fill did not run (see BsqFillArgs.bad)
  const uint32_t* codes2; // the same letters, 16 per 32-bit word (alphabets <= 4), or NULL
};
