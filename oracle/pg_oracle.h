/* oracle/pg_oracle.h -- TEST INFRASTRUCTURE ONLY (CPU restatement of the reference algorithm).
 *
 * Plain-C restatement of Pindel's pattern-growth split-read search, written from the reference's
 * behaviour (each function cites the reference file:line it follows).  It is the CHECKER for the
 * CUDA path: only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may use it.
 * Parity of this restatement is pinned against the reference itself (oracle/_ref, built from the
 * untouched sources) by tests/test_oracle_pinning.py and the fixtures under tests/golden/.
 */
#ifndef PG_ORACLE_H
#define PG_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PGO_SPACER 100000u /* g_SpacerBeforeAfter, src/pindel.h:122 */
#define PGO_TABLE 500      /* MAX_LENGTH in createProbTable, src/pindel.cpp:801 */

typedef struct {
   int32_t additional_mismatch;   /* -a, userSettings->ADDITIONAL_MISMATCH (>= 1, src/pindel.cpp:927-930) */
   int32_t min_close;             /* -H, g_MinClose (src/pindel.cpp:918) */
   int32_t min_perfect_match;     /* -m, Min_Perfect_Match_Around_BP */
   int32_t max_range_index;       /* -x, MaxRangeIndex */
   double max_mismatch_rate;      /* -u, MaximumAllowedMismatchRate */
   uint32_t max_mismatch[PGO_TABLE]; /* g_maxMismatch, src/pindel.cpp:794 */
} pgo_params;

typedef struct {
   uint32_t abs_loc;  /* UniquePoint::AbsLoc (padded coordinate) */
   int16_t len;       /* UniquePoint::LengthStr */
   int16_t mm;        /* UniquePoint::Mismatches */
   uint8_t dir;       /* 0 = FORWARD '+', 1 = BACKWARD '-' */
   uint8_t strand;    /* 0 = SENSE '+',   1 = ANTISENSE '-' */
   int16_t chr;       /* index of UniquePoint::chromosome_p */
} pgo_up;

typedef struct {
   int32_t chr;
   uint32_t start, end; /* SearchWindow [start, end) in padded coordinates */
} pgo_region;

/* src/pindel.cpp:781-819 (createProbTable / probOfReadWithTheseErrors), double arithmetic */
void pgo_create_prob_table(double seq_error_rate, double sensitivity, uint32_t *out /* PGO_TABLE */);

/* Pindel defaults (src/fn_parameters.cpp:17-351): -a 1 -H 8 -m 3 -x 2 -u 0.02 -e 0.01 -E 0.95 */
void pgo_default_params(pgo_params *p);

/* GetCloseEnd (src/pindel.cpp:2531-2605).  seq is the read buffer (capacity >= *len+1), rewritten
 * in place when the reference flips it.  Returns the number of raw UniquePoints written to out
 * (pre-CleanUniquePoints), *flips = number of setUnmatchedSeq(ReverseComplement()) calls made. */
int pgo_close_end(const pgo_params *p, const char *chr_seq, uint32_t chr_len, int chr_id, char *seq, int *len,
                  char matched_d, uint32_t matched_rel_pos, int insert_size, pgo_up *out, int *flips);

/* CleanUniquePoints (src/pindel.cpp:2904-2941), in place; returns the new count. */
int pgo_clean_unique_points(pgo_up *ups, int n);

/* SearchFarEnd (src/pindel.cpp:1001-1074) for one read whose close end is known.
 * regions/n_regions = the BreakDancer search cluster of the read (may be empty).
 * Returns the number of far UniquePoints written to out. */
int pgo_far_end(const pgo_params *p, int n_chr, const char *const *chr_seq, const uint32_t *chr_len, int chr_id,
                const char *seq, int len, const pgo_up *close, int n_close, const pgo_region *regions, int n_regions,
                pgo_up *out);

/* Batch driver: GetCloseEnd -> CleanUniquePoints -> SearchFarEnd per read (OpenMP over reads).
 * Output stream (int32, malloc'ed, release with pgo_free), per read:
 *   [flip, L_after, n_close_raw, n_close_clean, n_far, n_regions, UPs..(6 ints: len, absLoc, dir, strand, mm, chr), regions..(3 ints)]
 * -- the same layout oracle/ref_hook.cpp emits for the reference. */
int64_t pgo_search_stream(const pgo_params *p, int n_chr, const char *const *chr_seq, const uint32_t *chr_len, int n,
                          const char *seq_blob, const uint32_t *seq_off, const char *matched_d,
                          const uint32_t *matched_rel_pos, const int16_t *insert_size, const int32_t *chr_id,
                          const pgo_region *regions, const uint32_t *region_off, int do_far, int threads,
                          int32_t **out_stream, double *seconds);
void pgo_free(void *p);

/* CategorizePositions inner iterations executed since the last reset (op-roofline of SURVEY.md 8(d)); threads == 1 only. */
unsigned long long pgo_char_compares(int reset);

#ifdef __cplusplus
}
#endif
#endif
