// bsq_gen.cuh -- batched synthetic pair generator on the device: rand_seq
// (stochastics.py:28-41) + MutationProcess.mutate (stochastics.py:143-201) for a
// whole batch of pairs.  The reference draws one np.random.choice per base in
// Python; 10^5 pairs of 10 kb (config 3) take minutes that way and milliseconds
// here.  Included by exactly one translation unit (bsq_host.cu).
//
// Randomness is counter based (splitmix64 of (seed, pair, stream, counter)), so a
// pair's letters do not depend on the batch it is generated in and the CPU
// restatement (oracle/gen_oracle.c) reproduces every byte.  Probabilities are
// compared as 24-bit integers (threshold = floor(p * 2^24), computed once on the host).
//
//   stream 0  letters of the pair's genome chunk: origin = chunk[0, n), the mutant is
//             mutated from chunk[off, off + n)
//   stream 1  the mutation state machine, one draw (u, v, letter) per pass of the
//             reference's while loop (stochastics.py:174-200)
//   stream 2  off ~ U[0, max_offset]
//   stream 3  pad letters (fixed_mutant_len: mutants trimmed / padded to one length)
#pragma once
#include "bsq_common.h"

struct BsqGenArgs {
  unsigned long long seed;
  int L, n, max_offset, fixed_mlen;
  unsigned subst_thr, go_thr, ge_thr;   // floor(p * 2^24)
  long long npairs;
};

BSQ_HD inline unsigned long long bsq_sm64(unsigned long long x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}
BSQ_HD inline unsigned long long bsq_gen_key(unsigned long long seed, long long pair, int stream) {
  return bsq_sm64(seed ^ bsq_sm64((unsigned long long)pair * 4ull + (unsigned long long)stream));
}
// letter j of the chunk of a pair whose stream-0 key is k0
BSQ_HD inline int bsq_gen_chunk_letter(unsigned long long k0, long long j, int L) {
  const unsigned long long h = bsq_sm64(k0 + (unsigned long long)(j >> 1));
  const unsigned long long f = (j & 1) ? (h >> 32) : (h & 0xffffffffull);
  return (int)((f * (unsigned long long)L) >> 32);
}
BSQ_HD inline int bsq_gen_offset(unsigned long long seed, long long pair, int max_offset) {
  if (max_offset <= 0) return 0;
  const unsigned long long h = bsq_sm64(bsq_gen_key(seed, pair, 2));
  return (int)(((h >> 32) * (unsigned long long)(max_offset + 1)) >> 32);
}

// One run of the mutation state machine for a pair.  WRITE = false: count only.  Returns the
// number of letters the machine emitted (before trimming / padding); *ntx = transcript ops.
template <bool WRITE>
BSQ_HD inline int bsq_gen_mutate_one(const BsqGenArgs& G, long long pair, int off, uint8_t* mut, char* tx, int* ntx) {
  const unsigned long long k0 = bsq_gen_key(G.seed, pair, 0), k1 = bsq_gen_key(G.seed, pair, 1);
  const int L = G.L, n = G.n;
  const int limit = G.fixed_mlen > 0 ? G.fixed_mlen : 0x7fffffff;   // stop once the reads are full
  int pos = 0, op = 0, len = 0, nt = 0;     // op: 0 none / M / S, 1 = I, 2 = D
  unsigned long long t = 0;
  while (pos < n && len < limit) {
    const unsigned long long h = bsq_sm64(k1 + t++);
    const unsigned u = (unsigned)(h >> 40), v = (unsigned)((h >> 16) & 0xffffffull), l16 = (unsigned)(h & 0xffffull);
    int emit = -1;
    char c = 0;
    if (op != 0) {                          // stochastics.py:176-184
      if (u < G.ge_thr) {
        if (op == 1) { emit = (int)((l16 * (unsigned)L) >> 16); c = 'I'; }
        else { pos++; c = 'D'; }
      } else {
        op = 0;                             // the gap ends; nothing is appended to the transcript
      }
    } else if (u < G.go_thr) {              // :187-194
      if (v < (1u << 23)) { op = 2; pos++; c = 'D'; }
      else { op = 1; emit = (int)((l16 * (unsigned)L) >> 16); c = 'I'; }
    } else {                                // :195-199
      const int src = bsq_gen_chunk_letter(k0, (long long)off + pos, L);
      int cp = src;
      if (v < G.subst_thr && L > 1) cp = (src + 1 + (int)((l16 * (unsigned)(L - 1)) >> 16)) % L;
      emit = cp;
      c = (cp == src) ? 'M' : 'S';
      pos++;
    }
    if (emit >= 0) {
      if (WRITE) mut[len] = (uint8_t)emit;
      len++;
    }
    if (c) {
      if (WRITE && tx) tx[nt] = c;
      nt++;
    }
  }
  if (ntx) *ntx = nt;
  return len;
}

#ifndef BSQ_GEN_HOST_ONLY
// pass A: lengths (one thread per pair)
__global__ void __launch_bounds__(128) bsq_gen_count(const BsqGenArgs G, int* mlen, int* txlen, int* off_out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= G.npairs) return;
  const int off = bsq_gen_offset(G.seed, i, G.max_offset);
  int nt = 0;
  const int len = bsq_gen_mutate_one<false>(G, i, off, nullptr, nullptr, &nt);
  mlen[i] = G.fixed_mlen > 0 ? G.fixed_mlen : len;
  txlen[i] = nt;
  off_out[i] = off;
}
// origins: every thread one letter (coalesced)
__global__ void __launch_bounds__(256) bsq_gen_origins(const BsqGenArgs G, const long long* base, uint8_t* codes) {
  const long long total = G.npairs * (long long)G.n;
  for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < total; q += (long long)gridDim.x * blockDim.x) {
    const long long i = q / G.n;
    const int j = (int)(q - i * G.n);
    codes[base[i] + j] = (uint8_t)bsq_gen_chunk_letter(bsq_gen_key(G.seed, i, 0), j, G.L);
  }
}
// pass B: mutants (+ transcripts), one thread per pair
__global__ void __launch_bounds__(128) bsq_gen_write(const BsqGenArgs G, const long long* base, const int* off_in,
                                                     uint8_t* codes, const long long* txbase, char* tx) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= G.npairs) return;
  uint8_t* mut = codes + base[i] + G.n;
  const int len = bsq_gen_mutate_one<true>(G, i, off_in[i], mut, tx ? tx + txbase[i] : nullptr, nullptr);
  if (G.fixed_mlen > 0) {   // pad short mutants with random letters (stream 3)
    const unsigned long long k3 = bsq_gen_key(G.seed, i, 3);
    for (int j = len; j < G.fixed_mlen; j++) mut[j] = (uint8_t)bsq_gen_chunk_letter(k3, j, G.L);
  }
}
#endif
