/*
 * gen_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement of the batched pair generator (biseqt_b200/csrc/bsq_gen.cuh), itself a
 * batched form of the reference's rand_seq (biseqt/stochastics.py:28-41) and
 * MutationProcess.mutate (stochastics.py:143-201): the state machine below follows the
 * reference's while loop line by line; the random draws come from a counter-based generator
 * (splitmix64 over (seed, pair, stream, counter)) so that both sides can be compared byte
 * for byte.  Only tests/ may load this file.
 */
#include <stdint.h>

static uint64_t sm64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ULL;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ULL;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBULL;
  return x ^ (x >> 31);
}
static uint64_t key(uint64_t seed, int64_t pair, int stream) {
  return sm64(seed ^ sm64((uint64_t)pair * 4u + (uint64_t)stream));
}
static int letter_at(uint64_t k, int64_t j, int L) {   /* uniform letter j of a stream */
  uint64_t h = sm64(k + (uint64_t)(j >> 1));
  uint64_t f = (j & 1) ? (h >> 32) : (h & 0xffffffffULL);
  return (int)((f * (uint64_t)L) >> 32);
}

/* One pair.  origin[n], mutant[cap], tx[cap_tx] are written; returns the mutant length.
 * thresholds are floor(p * 2^24).  *off_out = offset of the mutant's source in the chunk. */
int gen_pair(uint64_t seed, int64_t pair, int L, int n, int max_offset, int fixed_mlen, uint32_t subst_thr,
             uint32_t go_thr, uint32_t ge_thr, uint8_t *origin, uint8_t *mutant, char *tx, int *ntx, int *off_out) {
  uint64_t k0 = key(seed, pair, 0), k1 = key(seed, pair, 1), k3 = key(seed, pair, 3);
  int off = 0;
  if (max_offset > 0) off = (int)(((sm64(key(seed, pair, 2)) >> 32) * (uint64_t)(max_offset + 1)) >> 32);
  for (int j = 0; j < n; j++) origin[j] = (uint8_t)letter_at(k0, j, L);            /* rand_seq, :28-41 */
  int pos = 0, len = 0, nt = 0;
  int op = 0;                  /* 0: previous op is not an indel; 1: I; 2: D */
  uint64_t t = 0;
  int limit = fixed_mlen > 0 ? fixed_mlen : 0x7fffffff;
  while (pos < n && len < limit) {                                                 /* :174 */
    uint64_t h = sm64(k1 + t++);
    uint32_t u = (uint32_t)(h >> 40), v = (uint32_t)((h >> 16) & 0xffffff), l16 = (uint32_t)(h & 0xffff);
    if (op != 0) {                                                                 /* :175 */
      if (u < ge_thr) {                                                            /* :177 coin_toss(ge_prob) */
        if (op == 1) { mutant[len++] = (uint8_t)((l16 * (uint32_t)L) >> 16); tx[nt++] = 'I'; }   /* :178-179 */
        else { pos++; tx[nt++] = 'D'; }                                            /* :180-181 */
      } else {
        op = 0;                                                                    /* :183 */
      }
    } else if (u < go_thr) {                                                       /* :186 coin_toss(go_prob) */
      if (v < (1u << 23)) { op = 2; pos++; tx[nt++] = 'D'; }                       /* :188-190 */
      else { op = 1; mutant[len++] = (uint8_t)((l16 * (uint32_t)L) >> 16); tx[nt++] = 'I'; }     /* :191-193 */
    } else {                                                                       /* :194-198 */
      int src = letter_at(k0, (int64_t)off + pos, L);
      int cp = src;
      if (v < subst_thr && L > 1) cp = (src + 1 + (int)((l16 * (uint32_t)(L - 1)) >> 16)) % L;
      mutant[len++] = (uint8_t)cp;
      tx[nt++] = (cp == src) ? 'M' : 'S';
      pos++;
    }
  }
  if (fixed_mlen > 0) {
    for (int j = len; j < fixed_mlen; j++) mutant[j] = (uint8_t)letter_at(k3, j, L);
    len = fixed_mlen;
  }
  *ntx = nt;
  *off_out = off;
  return len;
}
