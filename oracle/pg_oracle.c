/* oracle/pg_oracle.c -- TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement (plain C) of the reference's pattern-growth split-read search.  It keeps the
 * reference's own formulation -- candidate positions bucketed by mismatch level and grown one
 * read base at a time -- so that it is an independent check of the data-parallel formulation
 * the CUDA kernels use.  Citations are file:line under /root/reference/.
 *
 * Parity status: PINNED.  tests/test_oracle_pinning.py compares this file against the reference
 * itself (oracle/_ref/libpindel_ref.so) wherever that library exists, and against the committed
 * dumps of the reference under tests/golden/ everywhere else.
 */
#include "pg_oracle.h"

#include <ctype.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define FWD 0
#define BWD 1
#define SENSE_ 0
#define ANTI_ 1
#define MAX_LEVELS 64

/* ---- tables: src/pindel.cpp:948-970 (Match2N, Convert2RC4N; every other entry is zero) ---- */
static inline int is_acgt(char c) { return c == 'A' || c == 'C' || c == 'G' || c == 'T'; }

static inline char rc4n(char c)
{
   switch (c) {
   case 'A': return 'T';
   case 'C': return 'G';
   case 'G': return 'C';
   case 'T': return 'A';
   case 'N': return 'N';
   default: return 0;
   }
}

/* Matches(), src/searcher.cpp:36-44 */
/* Algorithmic work counter for the op-roofline of SURVEY.md section 8(d): every execution of the inner statement of
 * CategorizePositions (src/searcher.cpp:48-63: one candidate position compared with one read character) adds one.
 * Plain (non-atomic) counter: meaningful with threads == 1 only. */
static unsigned long long g_char_compares = 0;
unsigned long long pgo_char_compares(int reset)
{
   unsigned long long v = g_char_compares;
   if (reset) g_char_compares = 0;
   return v;
}

static inline int matches(char readBase, char refBase)
{
   if (readBase != 'N') return refBase == readBase;
   return is_acgt(refBase); /* Match2N[ref] == 'N' */
}

/* ReverseComplement(), src/pindel.cpp:2037-2048 */
static void reverse_complement(const char *in, int n, char *out)
{
   for (int j = 0; j < n; j++) out[j] = rc4n(in[n - 1 - j]);
}

/* SPLIT_READ::setUnmatchedSeq(), src/pindel.cpp:142-169: trailing non-alphanumerics are stripped */
static int strip_trailing(char *s, int n)
{
   while (n > 0 && !isalnum((unsigned char)s[n - 1])) n--;
   return n;
}

/* ---- src/pindel.cpp:781-819 ---- */
static double prob_of_read_with_these_errors(unsigned length, unsigned numberOfErrors, double seqErrorRate)
{
   double chanceOfCorrect = 1.0 - seqErrorRate;
   unsigned numberOfCorrectBases = length - numberOfErrors;
   double matchedPart = pow(chanceOfCorrect, numberOfCorrectBases);
   double mismatchedPart = 1.0;
   for (unsigned i = 0; i < numberOfErrors; i++) {
      mismatchedPart *= (((length - i) * seqErrorRate) / (numberOfErrors - i));
   }
   return matchedPart * mismatchedPart;
}

void pgo_create_prob_table(double seq_error_rate, double sensitivity, uint32_t *out)
{
   for (unsigned length = 0; length < PGO_TABLE; length++) {
      double total = 0.0;
      out[length] = 0;
      for (unsigned e = 0; e <= length; e++) {
         total += prob_of_read_with_these_errors(length, e, seq_error_rate);
         if (total > sensitivity) {
            out[length] = e + 1;
            break;
         }
      }
   }
   out[0] = out[1] = out[2] = out[3] = 0;
}

void pgo_default_params(pgo_params *p)
{
   p->additional_mismatch = 1;
   p->min_close = 8;
   p->min_perfect_match = 3;
   p->max_range_index = 2;
   p->max_mismatch_rate = 0.02;
   pgo_create_prob_table(0.001 + 0.01, 0.95, p->max_mismatch); /* src/pindel.cpp:856 */
}

/* ---- CheckMismatches(), src/searcher.cpp:331-388 ----
 * ref/ref_len: chromosome of the UniquePoint; read/L: the read's current UnmatchedSeq. */
static int check_mismatches(const pgo_params *P, const char *ref, uint32_t ref_len, const char *read, int L,
                            const pgo_up *up)
{
   char cur[1024];
   int m = P->min_perfect_match;
   if (up->strand == SENSE_) memcpy(cur, read, (size_t)L);
   else reverse_complement(read, L, cur);
   uint32_t start;
   if (up->dir == FWD) {
      start = up->abs_loc - (uint32_t)up->len + 1u;
      if (up->len <= m) return 0;
      /* BP_On_Read = cur.substr(len-m, m); BP_On_Ref = ref.substr(AbsLoc-m+1, m) */
      for (int i = 0; i < m; i++) {
         uint32_t rp = up->abs_loc - (uint32_t)m + 1u + (uint32_t)i;
         if (rp >= ref_len) return 0;
         if (cur[up->len - m + i] != ref[rp]) return 0;
      }
   } else {
      start = up->abs_loc + (uint32_t)up->len - (uint32_t)L;
      if (L < up->len) return 0;
      /* BP_On_Read = cur.substr(L-len, m) (clipped at the string end); BP_On_Ref = ref.substr(AbsLoc, m) */
      int avail = up->len < m ? up->len : m;
      if (avail != m) return 0; /* strings of different length never compare equal */
      for (int i = 0; i < m; i++) {
         uint32_t rp = up->abs_loc + (uint32_t)i;
         if (rp >= ref_len) return 0;
         if (cur[L - up->len + i] != ref[rp]) return 0;
      }
   }
   float max_allowed = (float)((double)L * P->max_mismatch_rate); /* size_t * double -> float, :366 */
   short n = 0;
   for (int i = 0; i < L; i++) {
      char r = ref[start + (uint32_t)i];
      if (cur[i] == 'N') {
         if (!is_acgt(r)) n++; /* Match2N[ref] != 'N' */
      } else if (r != cur[i]) {
         n++;
      }
   }
   return (float)n >= max_allowed;
}

/* One candidate position being grown: the reference keeps them in per-level vectors
 * (PD[level], src/searcher.cpp:48-63); here each carries its level explicitly. */
typedef struct {
   uint32_t pos;
   uint8_t level;
   uint8_t minus; /* far end only: member of PD_Minus */
   uint16_t region;
} cand_t;

/* ---- GetCloseEndInner(), src/pindel.cpp:2250-2396, with CheckLeft_Close / CheckRight_Close
 * (src/searcher.cpp:153-197 / 247-286) and ExtendMatchClose (src/searcher.cpp:65-97) unrolled
 * from recursion into a loop over CurrentLength. ---- */
static int close_end_inner(const pgo_params *P, const char *ref, uint32_t ref_len, int chr_id, const char *seq, int L,
                           char matched_d, uint32_t matched_rel_pos, int insert_size, int range_index, pgo_up *out)
{
   const uint32_t *g = P->max_mismatch;
   int MAX = (int)g[L];
   int TM = MAX + P->additional_mismatch;
   int n_out = 0;
   char cur[1024];
   int dir;
   int start, end;
   char c0;
   if (matched_d == '+') {
      reverse_complement(seq, L, cur);
      start = (int)(matched_rel_pos + PGO_SPACER) - range_index * insert_size;
      end = start + (2 * range_index + 1) * insert_size;
      c0 = cur[0];
      dir = 1;
   } else if (matched_d == '-') {
      memcpy(cur, seq, (size_t)L);
      end = (int)(matched_rel_pos + PGO_SPACER) + range_index * insert_size;
      start = end - (2 * range_index + 1) * insert_size;
      c0 = cur[L - 1];
      dir = -1;
   } else {
      return 0;
   }
   int cap = end > start ? end - start : 0;
   cand_t *cands = (cand_t *)malloc(sizeof(cand_t) * (size_t)(cap + 1));
   int nc = 0;
   if (c0 != 'N') {
      for (int pos = start; pos < end; pos++) {
         if (pos >= 0 && (uint32_t)pos < ref_len && ref[pos] == c0) {
            cands[nc].pos = (uint32_t)pos;
            cands[nc].level = 0;
            nc++;
         }
      }
   }
   int BP_Start = P->min_close;
   int BP_End = L - 1;
   unsigned cnt[MAX_LEVELS];
   uint32_t who[MAX_LEVELS];
   for (int len = 1;; len++) {
      memset(cnt, 0, sizeof(cnt));
      for (int i = 0; i < nc; i++) {
         cnt[cands[i].level]++;
         who[cands[i].level] = cands[i].pos;
      }
      if (len >= BP_Start && len <= BP_End) {
         int m = 0;
         while (m <= MAX && cnt[m] == 0) m++;                   /* minimumNumberOfMismatches, :141-150 */
         if ((unsigned)m > g[len]) break;                        /* :167-169 */
         for (int i = 0; i <= MAX; i++) {
            if (cnt[i] == 1 && len >= BP_Start + i) {
               unsigned sum = 0;
               for (int k = 0; k <= i + P->additional_mismatch; k++) sum += cnt[k]; /* :173 */
               if (sum == 1 && (unsigned)i <= g[len]) {
                  pgo_up up;
                  up.abs_loc = who[i];
                  up.len = (int16_t)len;
                  up.mm = (int16_t)i;
                  up.chr = (int16_t)chr_id;
                  if (dir == 1) { up.dir = FWD; up.strand = ANTI_; }   /* :181 */
                  else { up.dir = BWD; up.strand = SENSE_; }           /* :273 */
                  if (check_mismatches(P, ref, ref_len, seq, L, &up)) {
                     out[n_out++] = up;
                     break;
                  }
               }
            }
         }
      }
      if (len >= BP_End) break;
      /* ExtendMatchClose: one growth step, CategorizePositions for every level with maxMM = TM */
      char ch = (dir == 1) ? cur[len] : cur[L - 1 - len];
      int w = 0;
      unsigned kept_le_max = 0;
      g_char_compares += (unsigned long long)nc;
      for (int i = 0; i < nc; i++) {
         uint32_t p = cands[i].pos + (uint32_t)dir;
         int lvl = cands[i].level;
         if (!matches(ch, ref[p])) {
            if (lvl < TM) lvl++;
            else continue; /* dropped, src/searcher.cpp:57 */
         }
         cands[w].pos = p;
         cands[w].level = (uint8_t)lvl;
         w++;
         if (lvl <= MAX) kept_le_max++;
      }
      nc = w;
      if (kept_le_max == 0) break; /* Sum over levels 0..MAX, :85-87 */
   }
   free(cands);
   return n_out;
}

/* ---- GetCloseEnd(), src/pindel.cpp:2531-2605 ---- */
int pgo_close_end(const pgo_params *p, const char *chr_seq, uint32_t chr_len, int chr_id, char *seq, int *len,
                  char matched_d, uint32_t matched_rel_pos, int insert_size, pgo_up *out, int *flips)
{
   int n = 0;
   int nflips = 0;
   for (int range_index = 0; range_index < 2; range_index++) {
      n = close_end_inner(p, chr_seq, chr_len, chr_id, seq, *len, matched_d, matched_rel_pos, insert_size, range_index, out);
      if (n == 0) {
         char tmp[1024];
         reverse_complement(seq, *len, tmp);
         memcpy(seq, tmp, (size_t)*len);
         *len = strip_trailing(seq, *len);
         seq[*len] = 0;
         nflips++;
         n = close_end_inner(p, chr_seq, chr_len, chr_id, seq, *len, matched_d, matched_rel_pos, insert_size, range_index, out);
      }
      if (n > 0) break;
   }
   if (flips) *flips = nflips;
   return n;
}

/* ---- CleanUniquePoints(), src/pindel.cpp:2904-2941 ---- */
int pgo_clean_unique_points(pgo_up *ups, int n)
{
   if (n <= 0) return 0;
   pgo_up last = ups[n - 1];
   int w = 0;
   uint32_t terminal = (last.dir == FWD) ? last.abs_loc - (uint32_t)last.len : last.abs_loc + (uint32_t)last.len;
   for (int i = 0; i < n; i++) {
      if (ups[i].chr != last.chr) continue;
      if (ups[i].dir != last.dir || ups[i].strand != last.strand) continue;
      uint32_t t = (last.dir == FWD) ? ups[i].abs_loc - (uint32_t)ups[i].len : ups[i].abs_loc + (uint32_t)ups[i].len;
      if (t == terminal) ups[w++] = ups[i];
   }
   return w;
}

/* ---- SearchFarEndAtPos(), src/farend_searcher.cpp:46-103, with CheckBoth (src/pindel.cpp:2823-2902)
 * and ExtendMatch (src/pindel.cpp:2673-2725) unrolled into a loop over CurrentLength.
 * Returns -1 when the reference would leave UP_Far untouched (no search / no seeds), else the
 * number of UniquePoints written to out. ---- */
static int far_end_at_pos(const pgo_params *P, const char *const *chr_seq, const uint32_t *chr_len, const char *seq,
                          int L, int close_maxlen, const pgo_region *regions, int n_regions, pgo_up *out)
{
   const uint32_t *g = P->max_mismatch;
   int MAX = (int)g[L];
   int TM = MAX + P->additional_mismatch;
   char base = seq[0];
   char base_rc = rc4n(base);
   if (base == 'N' || close_maxlen == 0) return -1; /* :54-56 */
   size_t cap = 0;
   for (int r = 0; r < n_regions; r++) {
      int s = (int)regions[r].start, e = (int)regions[r].end;
      if (s < 0) s = e - 1;
      if (e > s) cap += (size_t)(e - s);
   }
   cand_t *cands = (cand_t *)malloc(sizeof(cand_t) * (cap + 1));
   size_t nc = 0;
   for (int r = 0; r < n_regions; r++) {
      int s = (int)regions[r].start, e = (int)regions[r].end;
      if (s < 0) s = e - 1; /* :67-69 */
      const char *ref = chr_seq[regions[r].chr];
      for (int pos = s; pos < e; pos++) {
         char c = ref[pos];
         if (c == base) {
            cands[nc].pos = (uint32_t)pos; cands[nc].level = 0; cands[nc].minus = 0; cands[nc].region = (uint16_t)r; nc++;
         } else if (c == base_rc) {
            cands[nc].pos = (uint32_t)pos; cands[nc].level = 0; cands[nc].minus = 1; cands[nc].region = (uint16_t)r; nc++;
         }
      }
   }
   if (nc == 0) { /* NumberOfHits == 0, :88 */
      free(cands);
      return -1;
   }
   int n_out = 0;
   int BP_Start = 10; /* :89 */
   int BP_End = L - 1;
   unsigned cnt[MAX_LEVELS];
   size_t who[MAX_LEVELS];
   for (int len = 1;; len++) {
      memset(cnt, 0, sizeof(cnt));
      for (size_t i = 0; i < nc; i++) {
         cnt[cands[i].level]++;
         who[cands[i].level] = i;
      }
      if (len >= BP_Start && len <= BP_End) {
         int m = 0;
         while (m <= MAX && cnt[m] == 0) m++;             /* minimumNumberOfMismatches, src/pindel.cpp:2727-2740 */
         if ((unsigned)m > g[len]) break;                  /* :2836-2838 */
         unsigned fewer = 0;
         for (int nmm = 0; nmm <= MAX; nmm++) {
            if (fewer) break;                              /* :2840-2842 */
            unsigned sum = cnt[nmm];
            fewer = sum;
            if (sum == 1 && len >= BP_Start + nmm) {
               if (P->additional_mismatch > 0) {           /* :2851 */
                  unsigned total = 0;
                  for (int k = 0; k <= nmm + P->additional_mismatch; k++) total += cnt[k];
                  if (total == 1 && (unsigned)nmm <= g[len]) {
                     const cand_t *c = &cands[who[nmm]];
                     pgo_up up;
                     up.abs_loc = c->pos;
                     up.len = (int16_t)len;
                     up.mm = (int16_t)nmm;
                     up.chr = (int16_t)regions[c->region].chr;
                     if (!c->minus) { up.dir = FWD; up.strand = SENSE_; }   /* :2881 */
                     else { up.dir = BWD; up.strand = ANTI_; }              /* :2884 */
                     if (check_mismatches(P, chr_seq[up.chr], chr_len[up.chr], seq, L, &up)) {
                        out[n_out++] = up;
                        break;
                     }
                  }
               }
            }
         }
      }
      if (len >= BP_End) break;
      /* ExtendMatch: plus candidates grow +1 with read[len], minus candidates grow -1 with RC(read[len]) */
      char ch = seq[len];
      char ch_rc = rc4n(ch);
      size_t w = 0;
      g_char_compares += (unsigned long long)nc;
      for (size_t i = 0; i < nc; i++) {
         const char *ref = chr_seq[regions[cands[i].region].chr];
         uint32_t p;
         int ok;
         if (!cands[i].minus) { p = cands[i].pos + 1u; ok = matches(ch, ref[p]); }
         else { p = cands[i].pos - 1u; ok = matches(ch_rc, ref[p]); }
         int lvl = cands[i].level;
         if (!ok) {
            if (lvl < TM) lvl++;
            else continue;
         }
         cands[w] = cands[i];
         cands[w].pos = p;
         cands[w].level = (uint8_t)lvl;
         w++;
      }
      nc = w;
      if (nc == 0) break; /* AllEmpty, :2716 */
   }
   free(cands);
   return n_out;
}

static int max_len(const pgo_up *ups, int n) { return n > 0 ? ups[n - 1].len : 0; } /* SortedUniquePoints::MaxLen, :490-498 */

/* ---- SearchFarEnd(), src/pindel.cpp:1001-1074 ---- */
int pgo_far_end(const pgo_params *p, int n_chr, const char *const *chr_seq, const uint32_t *chr_len, int chr_id,
                const char *seq, int len, const pgo_up *close, int n_close, const pgo_region *regions, int n_regions,
                pgo_up *out)
{
   (void)n_chr;
   int n_far = 0;
   if (n_close == 0) return 0;
   pgo_up *tmp = (pgo_up *)malloc(sizeof(pgo_up) * (size_t)(len + 2));
   int close_max = max_len(close, n_close);
   if (n_regions != 0) {
      int k = far_end_at_pos(p, chr_seq, chr_len, seq, len, close_max, regions, n_regions, tmp);
      if (k >= 0 && max_len(tmp, k) >= max_len(out, n_far)) { /* NewUPFarIsBetter, farend_searcher.cpp:30-44 */
         memcpy(out, tmp, sizeof(pgo_up) * (size_t)k);
         n_far = k;
      }
      if (max_len(out, n_far) + close_max >= len) { free(tmp); return n_far; } /* goodFarEndFound, :480-483 */
   }
   uint32_t span = 64;
   uint32_t center = close[n_close - 1].abs_loc; /* getLastAbsLocCloseEnd, :474-477 */
   uint32_t size = chr_len[chr_id];
   for (int range_index = 1; range_index <= p->max_range_index + 1; range_index++) {
      pgo_region w;
      w.chr = chr_id;
      w.start = (center > span + PGO_SPACER) ? center - span : PGO_SPACER;
      w.end = (center + span + PGO_SPACER < size) ? center + span : size - PGO_SPACER;
      int k = far_end_at_pos(p, chr_seq, chr_len, seq, len, close_max, &w, 1, tmp);
      if (k >= 0 && max_len(tmp, k) >= max_len(out, n_far)) {
         memcpy(out, tmp, sizeof(pgo_up) * (size_t)k);
         n_far = k;
      }
      if (max_len(out, n_far) + close_max >= len) break;
      span *= 4;
   }
   free(tmp);
   return n_far;
}

static double now_s(void)
{
   struct timespec ts;
   clock_gettime(CLOCK_MONOTONIC, &ts);
   return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

typedef struct {
   int flip, L, ncr, ncc, nf;
   pgo_up *cr, *cc, *fa;
} read_result;

static void emit_up(int32_t *o, const pgo_up *u)
{
   o[0] = u->len; o[1] = (int32_t)u->abs_loc; o[2] = u->dir; o[3] = u->strand; o[4] = u->mm; o[5] = u->chr;
}

int64_t pgo_search_stream(const pgo_params *p, int n_chr, const char *const *chr_seq, const uint32_t *chr_len, int n,
                          const char *seq_blob, const uint32_t *seq_off, const char *matched_d,
                          const uint32_t *matched_rel_pos, const int16_t *insert_size, const int32_t *chr_id,
                          const pgo_region *regions, const uint32_t *region_off, int do_far, int threads,
                          int32_t **out_stream, double *seconds)
{
   read_result *res = (read_result *)calloc((size_t)n + 1, sizeof(read_result));
   char *seqs = (char *)malloc((size_t)n * 1024);
#ifdef _OPENMP
   omp_set_num_threads(threads > 0 ? threads : 1);
#else
   (void)threads;
#endif
   double t0 = now_s();
   #pragma omp parallel for schedule(dynamic, 64)
   for (int i = 0; i < n; i++) {
      read_result *r = &res[i];
      char *s = seqs + (size_t)i * 1024;
      int L = (int)(seq_off[i + 1] - seq_off[i]);
      if (L > 1000) L = 1000;
      memcpy(s, seq_blob + seq_off[i], (size_t)L);
      s[L] = 0;
      L = strip_trailing(s, L); /* first setUnmatchedSeq */
      int c = chr_id ? chr_id[i] : 0;
      r->cr = (pgo_up *)malloc(sizeof(pgo_up) * (size_t)(L + 2));
      int flips = 0;
      r->ncr = pgo_close_end(p, chr_seq[c], chr_len[c], c, s, &L, matched_d[i], matched_rel_pos[i], insert_size[i], r->cr, &flips);
      r->L = L;
      r->flip = (memcmp(s, seq_blob + seq_off[i], (size_t)L) != 0 || (uint32_t)L != seq_off[i + 1] - seq_off[i]) ? 1 : 0;
      r->cc = (pgo_up *)malloc(sizeof(pgo_up) * (size_t)(r->ncr + 1));
      memcpy(r->cc, r->cr, sizeof(pgo_up) * (size_t)r->ncr);
      r->ncc = pgo_clean_unique_points(r->cc, r->ncr);
   }
   double t1 = now_s();
   if (do_far) {
      #pragma omp parallel for schedule(dynamic, 64)
      for (int i = 0; i < n; i++) {
         read_result *r = &res[i];
         if (r->ncc == 0) continue;
         int c = chr_id ? chr_id[i] : 0;
         const pgo_region *rg = NULL;
         int nrg = 0;
         if (regions && region_off) { rg = regions + region_off[i]; nrg = (int)(region_off[i + 1] - region_off[i]); }
         r->fa = (pgo_up *)malloc(sizeof(pgo_up) * (size_t)(r->L + 2));
         r->nf = pgo_far_end(p, n_chr, chr_seq, chr_len, c, seqs + (size_t)i * 1024, r->L, r->cc, r->ncc, rg, nrg, r->fa);
      }
   }
   double t2 = now_s();
   if (seconds) { seconds[0] = t1 - t0; seconds[1] = t2 - t1; }
   size_t total = 0;
   for (int i = 0; i < n; i++) {
      size_t nrg = (regions && region_off && res[i].ncc) ? region_off[i + 1] - region_off[i] : 0;
      total += 6 + 6 * (size_t)(res[i].ncr + res[i].ncc + res[i].nf) + 3 * nrg;
   }
   int32_t *o = (int32_t *)malloc(sizeof(int32_t) * (total + 1));
   size_t q = 0;
   for (int i = 0; i < n; i++) {
      read_result *r = &res[i];
      size_t nrg = (regions && region_off && r->ncc) ? region_off[i + 1] - region_off[i] : 0;
      o[q++] = r->flip; o[q++] = r->L; o[q++] = r->ncr; o[q++] = r->ncc; o[q++] = r->nf; o[q++] = (int32_t)nrg;
      for (int j = 0; j < r->ncr; j++, q += 6) emit_up(o + q, &r->cr[j]);
      for (int j = 0; j < r->ncc; j++, q += 6) emit_up(o + q, &r->cc[j]);
      for (int j = 0; j < r->nf; j++, q += 6) emit_up(o + q, &r->fa[j]);
      for (size_t j = 0; j < nrg; j++) {
         const pgo_region *g = &regions[region_off[i] + j];
         o[q++] = g->chr; o[q++] = (int32_t)g->start; o[q++] = (int32_t)g->end;
      }
      free(r->cr); free(r->cc); free(r->fa);
   }
   free(res);
   free(seqs);
   *out_stream = o;
   return (int64_t)q;
}

void pgo_free(void *p) { free(p); }
