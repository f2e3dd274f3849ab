/* include/pindel_pg.h -- C ABI of the B200 pattern-growth split-read search (libpindel_pg.so).
 *
 * The reference (genome/pindel) has no plugin or FFI layer; the seam this library replaces is the pair
 * of C++ functions that the reference calls inside its OpenMP loops (citations under /root/reference/):
 *
 *   void GetCloseEnd(const std::string& CurrentChrSeq, SPLIT_READ& read)          src/pindel.h:523
 *        callers: ReadBuffer::flush  src/read_buffer.cpp:39-50 (BAM input, <= 50 000 reads per flush)
 *                 ReadInRead         src/reader.cpp:248-256, 300-306 (Pindel-text input, 1000 per batch)
 *   void SearchFarEnds(const std::string& chrSeq, std::vector<SPLIT_READ>& reads,
 *                      const Chromosome& chr)                                      src/pindel.h:741
 *        caller:  main               src/pindel.cpp:1888
 *
 * pg_close_end() is the batched GetCloseEnd, pg_far_end() the batched SearchFarEnds, pg_search() runs
 * both back to back on the device.  Plain pointers and sizes only; the caller owns every host buffer,
 * the library owns device memory and the packed reference.  There is no CPU fallback behind this ABI:
 * every entry point fails with PG_ERR_CUDA when no usable device exists.
 *
 * Coordinates are the reference's: positions index the padded in-memory chromosome string
 * (100 000 'N' before and after the sequence, src/pindel.cpp:305-309).
 */
#ifndef PINDEL_PG_H
#define PINDEL_PG_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PG_OK 0
#define PG_ERR_ARG 1        /* bad argument */
#define PG_ERR_CUDA 2       /* CUDA runtime error (including "no device") */
#define PG_ERR_STATE 3      /* call order (e.g. search before pg_set_reference) */
#define PG_ERR_RUNS_CAP 4   /* caller's run buffer too small: results->n_runs holds the required size; call pg_download again */
#define PG_ERR_UNSUPPORTED 5 /* a read is longer, or needs more mismatch levels, than the library supports */
#define PG_ERR_DATA 6        /* corrupt input data (pg_bgzf_inflate: not BGZF, invalid DEFLATE stream, CRC32 mismatch) */

typedef struct pg_ctx pg_ctx;

/* Pindel's command-line options that reach the search path (src/fn_parameters.cpp:17-351). */
typedef struct pg_settings {
   double seq_error_rate;        /* -e  Seq_Error_Rate              (default 0.01) */
   double sensitivity;           /* -E  sensitivity                 (default 0.95) */
   double max_mismatch_rate;     /* -u  MaximumAllowedMismatchRate  (default 0.02) */
   int32_t additional_mismatch;  /* -a  ADDITIONAL_MISMATCH         (default 1, minimum 1: src/pindel.cpp:927-930) */
   int32_t min_close;            /* -H  minClose / g_MinClose       (default 8) */
   int32_t min_perfect_match;    /* -m  Min_Perfect_Match_Around_BP (default 3) */
   int32_t max_range_index;      /* -x  MaxRangeIndex               (default 2, maximum 9: src/pindel.cpp:919-924) */
   int32_t device;               /* CUDA device ordinal */
   int32_t reserved;
} pg_settings;

/* One run of consecutive UniquePoints (src/pindel.h:137-158) found from the same seed with the same number of
 * mismatches: LengthStr = len0 .. len0+count-1, AbsLoc = abs_loc0 +/- (LengthStr - len0) (+ FORWARD, - BACKWARD). */
typedef struct pg_run {
   uint32_t abs_loc0;
   uint16_t len0;
   uint16_t count;
   uint8_t mm;          /* UniquePoint::Mismatches */
   uint8_t dirstrand;   /* bit0: Direction 0 FORWARD '+' / 1 BACKWARD '-';  bit1: Strand 0 SENSE '+' / 1 ANTISENSE '-' */
   int16_t chr;         /* index into the chromosomes given to pg_set_reference (UniquePoint::chromosome_p) */
} pg_run;

/* Expanded UniquePoint. */
typedef struct pg_up {
   uint32_t abs_loc;
   int16_t len;
   int16_t mm;
   uint8_t dir;         /* 0 FORWARD, 1 BACKWARD */
   uint8_t strand;      /* 0 SENSE, 1 ANTISENSE */
   int16_t chr;
} pg_up;

#define PG_READ_FLIPPED 1u      /* UnmatchedSeq must be reverse-complemented on the host (GetCloseEnd mutates the read, src/pindel.cpp:2545) */
#define PG_READ_UNSUPPORTED 2u  /* read not searched (see PG_ERR_UNSUPPORTED) */
#define PG_READ_OVERFLOW 4u     /* internal: never visible after a successful call */
#define PG_READ_NFLIPS(flags) (((flags) >> 8) & 3u)  /* times GetCloseEnd ran setUnmatchedSeq(ReverseComplement(seq)) on this read
                                                        (0..2); replay them on the host copy of the read, in order: a flip maps
                                                        non-ACGTN bases to NUL and strips the ones that end up trailing */

typedef struct pg_read_out {
   uint32_t run_off;        /* this read's runs: runs[run_off .. run_off + n_close_runs) = UP_Close (before CleanUniquePoints), */
   uint16_t n_close_runs;   /*                   then n_far_runs runs = UP_Far */
   uint16_t n_far_runs;
   uint16_t len_after;      /* ReadLength after the search (setUnmatchedSeq strips trailing non-alphanumerics, src/pindel.cpp:142-169) */
   uint16_t flags;          /* PG_READ_* */
} pg_read_out;

/* SearchWindow [start, end) on chromosome chr: one region of a BreakDancer search cluster
 * (g_bdData.getCorrespondingSearchWindowCluster, src/bddata.cpp:949-967). */
typedef struct pg_region {
   int32_t chr;
   uint32_t start, end;
} pg_region;

/* What SearchFarEnd needs to know about a read's close end when pg_far_end is called on its own. */
typedef struct pg_close_info {
   uint32_t last_abs_loc;   /* UP_Close.back().AbsLoc (getLastAbsLocCloseEnd, src/pindel.cpp:474-477) */
   uint16_t max_len;        /* UP_Close.MaxLen(); 0 = no close end, the read is skipped */
   uint16_t pad;
} pg_close_info;

/* A batch of anchored unmapped read-ends in SoA form (fields of SPLIT_READ, src/pindel.h:265-383). */
typedef struct pg_reads {
   int32_t n;
   const char *seq_blob;            /* UnmatchedSeq of every read, concatenated (ASCII) */
   const uint32_t *seq_off;         /* [n+1] offsets into seq_blob */
   const char *matched_d;           /* [n] MatchedD '+' / '-'            (pg_far_end: may be NULL) */
   const uint32_t *matched_rel_pos; /* [n] MatchedRelPos                 (pg_far_end: may be NULL) */
   const int16_t *insert_size;      /* [n] InsertSize                    (pg_far_end: may be NULL) */
   const int32_t *chr_id;           /* [n] chromosome of the anchor (FragName); NULL = all on chromosome 0 */
   const pg_close_info *close_info; /* [n] pg_far_end only */
   const pg_region *regions;        /* BreakDancer clusters of all reads, concatenated; NULL = none */
   const uint32_t *region_off;      /* [n+1] offsets into regions; NULL = none */
} pg_reads;

typedef struct pg_results {
   pg_read_out *out;        /* [n], caller-allocated */
   pg_run *runs;            /* [runs_cap], caller-allocated */
   uint32_t runs_cap;
   uint32_t n_runs;         /* out: runs written (or required, with PG_ERR_RUNS_CAP) */
} pg_results;

void pg_default_settings(pg_settings *s);
pg_ctx *pg_create(const pg_settings *s);          /* NULL on failure: see pg_create_error() */
const char *pg_create_error(void);
void pg_destroy(pg_ctx *ctx);
const char *pg_last_error(pg_ctx *ctx);

/* Use an existing CUDA stream (a cudaStream_t) for every copy and launch; NULL = the library's own stream. */
int pg_set_stream(pg_ctx *ctx, void *cuda_stream);

/* Chromosome strings exactly as the reference holds them in memory (Genome::loadChromosome, src/pindel.cpp:270-312):
 * upper-case, non-ACGT as 'N', padded with 100 000 'N' on both sides.  The library packs them to 2 bits per base
 * (+ a 1-bit N plane) on the device; the host strings are not referenced after the call returns. */
int pg_set_reference(pg_ctx *ctx, int32_t n_chr, const char *const *seq, const uint32_t *len);

/* g_maxMismatch as derived from -e/-E (createProbTable, src/pindel.cpp:781-819); out has 500 entries. */
int pg_max_mismatch_table(pg_ctx *ctx, uint32_t *out);

/* The reference seam, batched (host buffers in, host buffers out; copies are inside the call).  Batches of 256 k reads and
 * more are cut into chunks (at most 512 k reads) whose uploads, kernels and downloads overlap; the per-read argument checks
 * (offsets, chromosome numbers, read lengths) then run chunk by chunk, so an invalid read late in the batch is reported after
 * earlier chunks were searched: on any error the contents of `res` are undefined and nothing stays resident. */
int pg_close_end(pg_ctx *ctx, const pg_reads *in, pg_results *res);   /* GetCloseEnd   for every read */
int pg_far_end(pg_ctx *ctx, const pg_reads *in, pg_results *res);     /* SearchFarEnd  for every read with a close end */
int pg_search(pg_ctx *ctx, const pg_reads *in, pg_results *res);      /* both, without BreakDancer clusters unless given up front */

/* The same, split into its asynchronous steps (for callers that keep batches resident or overlap batches):
 * pg_upload copies a batch to the device, pg_launch runs mode (1 close, 2 far, 3 both) on the resident batch,
 * pg_download waits and copies the results back.  pg_launch may be repeated on the same resident batch. */
int pg_upload(pg_ctx *ctx, const pg_reads *in);
int pg_launch(pg_ctx *ctx, int32_t mode);
int pg_download(pg_ctx *ctx, pg_results *res);
int pg_sync(pg_ctx *ctx);

/* pg_upload + pg_launch in one call that leaves batch and results on the device (read them with pg_device_results /
 * pg_download, classify them with pg_classify_resident).  Large batches are uploaded in chunks while the kernel already
 * runs on the chunks that have arrived; the call returns once everything is queued.  The host buffers of `in` must stay
 * valid (and, to overlap at all, pinned: pg_host_alloc) until the results have been waited for. */
int pg_search_resident(pg_ctx *ctx, const pg_reads *in, int32_t mode);

/* Size limits of one call: seq_off / region_off / run offsets are 32-bit, so one batch holds less than 4 GiB of sequence
 * (28.6 M reads of 150 bp) and produces less than 2^32 runs; larger inputs are split by the caller (the reference itself
 * works window by window, src/pindel.cpp:1778-1985).  pg_upload / pg_search return PG_ERR_ARG when seq_off wraps. */

/* Several GPUs behind one handle ("device id(s) ... multi-GPU fan-out is internal"): a group owns one context per device,
 * the packed reference is replicated, and every batch call deals the reads to the devices in contiguous ranges -- one host
 * thread per device uploads its range behind its kernel, the per-device run counts become offsets, and every device copies
 * its records straight to their place in the caller's buffers.  Reads are independent units (each call mutates only its own
 * SPLIT_READ, src/pindel.cpp:1118-1121), so the results are exactly those of one device, in the caller's order.  This is the
 * form a single-process host (the reference's main()) uses; one process per GPU with an NCCL gather of the records is the other
 * (pindel_b200/shard.py, bench.py).  devices == NULL: 0 .. n_devices-1.  Errors as above, text from pg_group_last_error;
 * pg_group_context(g, k) gives the k-th device's context for the single-device calls (classification, BAM stages). */
typedef struct pg_group pg_group;
pg_group *pg_group_create(const pg_settings *s, int32_t n_devices, const int32_t *devices);   /* NULL: see pg_create_error() */
void pg_group_destroy(pg_group *g);
const char *pg_group_last_error(pg_group *g);
int32_t pg_group_size(pg_group *g);
pg_ctx *pg_group_context(pg_group *g, int32_t k);
int pg_group_set_reference(pg_group *g, int32_t n_chr, const char *const *seq, const uint32_t *len);
int pg_group_close_end(pg_group *g, const pg_reads *in, pg_results *res);
int pg_group_far_end(pg_group *g, const pg_reads *in, pg_results *res);
int pg_group_search(pg_group *g, const pg_reads *in, pg_results *res);

/* Bookkeeping the benchmarks report. */
uint64_t pg_kernel_launches(pg_ctx *ctx);      /* search-kernel launches since pg_create */
uint32_t pg_last_run_count(pg_ctx *ctx);       /* runs produced by the last completed launch (valid after pg_sync/pg_download) */

/* Statistics of the last completed launch (waits for it): reference positions scanned (sum of the window sizes of every
 * search executed, the W terms of the algorithmic-bytes formula in DESIGN.md) and UniquePoints produced. */
int pg_last_stats(pg_ctx *ctx, uint64_t *window_bases, uint64_t *unique_points);

/* Device pointers of the last launch's results (pg_read_out[n], pg_run[n_runs]; waits for that launch only), for
 * callers that hand the records to a collective (NCCL gather) without a host round trip.  Results rotate through three
 * slots: the pointers of launch k stay valid until launch k+3 is queued, so the records of launch k can travel while the
 * launches k+1 and k+2 run -- pg_device_results_previous returns launch k's after launch k+1 was queued (waits for launch
 * k only); the caller orders launch k+3 behind whatever still reads launch k's records.  PG_ERR_STATE when there is no
 * such launch. */
int pg_device_results(pg_ctx *ctx, void **out_dev, void **runs_dev, uint32_t *n_runs);
int pg_device_results_previous(pg_ctx *ctx, void **out_dev, void **runs_dev, uint32_t *n_runs, int32_t *n_reads);

/* Pinned host memory for batch buffers (optional; any host pointer is accepted by the calls above). */
void *pg_host_alloc(size_t bytes);
void pg_host_free(void *p);

/* ------------------------------------------------------------------------------------------------------------------
 * SV classification: the seven per-read drivers SearchSVs runs after the far-end search
 * (src/pindel.cpp:1140-1164) -- SearchDeletions / SearchShortInsertions ::Search (src/search_variant.cpp:49),
 * searchIndels (src/search_deletions_nt.cpp:26), searchTandemDuplications(NT) (src/search_tandem_duplications.cpp:28,
 * src/search_tandem_duplications_nt.cpp:27), searchInversions(NT) (src/search_inversions.cpp:29,
 * src/search_inversions_nt.cpp:27).  One call classifies every read of a window with all selected drivers; what each
 * driver did to a read comes back as records the host replays driver by driver (Vars[box].push_back, FutureReads_SR,
 * the SPLIT_READ fields), then calls the reference's SortOutput* as before.
 * ------------------------------------------------------------------------------------------------------------------ */
#define PG_SV_D 0
#define PG_SV_DI 1
#define PG_SV_TD 2
#define PG_SV_TD_NT 3
#define PG_SV_INV 4
#define PG_SV_INV_NT 5
#define PG_SV_SI 6
#define PG_SV_ALL_DRIVERS 0x7Fu

#define PG_SV_NONE 0     /* fields assigned, read not taken */
#define PG_SV_BOX 1      /* Vars[box].push_back(ReadIndex); Used = true */
#define PG_SV_FUTURE 2   /* saveReadForNextCycle (src/pindel.cpp:587-592) */
#define PG_SV_BREAK 3    /* breakpoint beyond the chromosome string: Used = true (src/search_variant.cpp:133-136) */

#define PG_SVF_CORE 1u   /* record carries Left, Right, BP, IndelSize, BPLeft, BPRight */
#define PG_SVF_NTSTR 2u  /* ... NT_str */
#define PG_SVF_NTSIZE 4u /* ... NT_size */
#define PG_SVF_KILLED 8u /* LeftMostTD / LeftMostINV set Used = true before the outcome (src/search_tandem_duplications.cpp:196-202) */

#define PG_NT_EMPTY 0    /* NT_str = "" */
#define PG_NT_SEQ 1      /* NT_str = rotl(UnmatchedSeq.substr(nt_start, nt_len), nt_rot) */
#define PG_NT_RC 2       /* NT_str = rotl(ReverseComplement(UnmatchedSeq).substr(nt_start, nt_len), nt_rot) */

#define PG_SVR_USED 1u       /* pg_sv_reads.rflags: SPLIT_READ::Used */
#define PG_SVR_OTHERCHR 2u   /*                     FragName != FarFragName */

typedef struct pg_sv_record {
   uint32_t read;        /* ReadIndex in the batch */
   uint8_t driver;       /* PG_SV_D .. PG_SV_SI */
   uint8_t outcome;      /* PG_SV_NONE / BOX / FUTURE / BREAK */
   uint8_t fields;       /* PG_SVF_* */
   uint8_t seq;          /* order among the records of (driver, read) */
   int32_t left, right;  /* SPLIT_READ::Left, Right   (src/pindel.h:350-351) */
   uint32_t bp_left, bp_right, indel_size;
   int16_t bp;
   uint16_t nt_size;
   uint16_t nt_start, nt_len, nt_rot;
   uint8_t nt_src;       /* PG_NT_* */
   uint8_t pad;
   uint32_t box;         /* TempBoxIndex (PG_SV_BOX) */
} pg_sv_record;

typedef struct pg_sv_window {
   uint32_t driver_mask;          /* bit d = run driver d; Analyze_TD / Analyze_INV clear theirs (src/pindel.cpp:1151-1158) */
   int32_t chr;                   /* chromosome of the window (pg_set_reference order) */
   uint32_t window_end;           /* SearchWindow::getEnd() */
   uint32_t num_boxes, box_size;  /* NumBoxes, BoxSize (src/pindel.cpp:1806-1810) */
   uint32_t region_start_defined, region_start, region_end_defined, region_end;   /* userSettings->getRegion() */
   int32_t min_inversion_size;    /* -v  MIN_IndelSize_Inversion (default 30) */
   int32_t min_matched_bases;     /* -M  Min_Num_Matched_Bases   (default 30) */
   int32_t reserved;
} pg_sv_window;

/* A window's reads after the far-end search: final UnmatchedSeq, anchor strand, insert size, flags, and the UP_Close
 * (after CleanUniquePoints; raw lists are accepted, the filter is idempotent) and UP_Far lists as runs. */
typedef struct pg_sv_reads {
   int32_t n_reads;
   const uint8_t *seq_blob;
   const uint32_t *seq_off;       /* [n+1] */
   const uint8_t *matched_d;
   const int16_t *insert_size;
   const uint8_t *rflags;         /* PG_SVR_* */
   const uint32_t *run_off;       /* [n+1] first run of each read: its UP_Close runs, then its UP_Far runs */
   const uint16_t *n_close_runs;  /* [n] */
   const pg_run *runs;
} pg_sv_reads;

typedef struct pg_sv_results {
   pg_sv_record *records;         /* caller-allocated, sorted by (driver, read, seq) on return */
   uint32_t records_cap;
   uint32_t n_records;            /* on return; if > records_cap the call returned PG_ERR_RUNS_CAP and nothing was copied */
} pg_sv_results;

int pg_classify(pg_ctx *ctx, const pg_sv_reads *in, const pg_sv_window *win, pg_sv_results *res);
/* The same on the batch that is still resident on the device after a far-end launch -- pg_far_end / pg_search on a batch
 * small enough for one launch (< 2 M reads), or pg_upload + pg_launch(2 or 3) + pg_download: sequences, anchors and the
 * launch's own run lists are used where they lie, nothing is uploaded again.  This is how a host chains the two seams of
 * one window: SearchFarEnds (src/pindel.cpp:1888) -> SearchSVs (src/pindel.cpp:1140).  ReadIndex in the records is the
 * index in that batch.
 *   after pg_launch(2) / pg_far_end: the sequences given to the launch are final (GetCloseEnd already ran); the reads'
 *       UP_Close lists come from the caller as runs (close_run_off [n+1], close_runs; raw or cleaned);
 *   after pg_launch(3) / pg_search: the launch's close runs are used and GetCloseEnd's flips are applied on the device;
 *       close_run_off / close_runs are ignored.
 * The batch must carry matched_d and insert_size.  rflags: [n] PG_SVR_USED per read, or NULL; FragName != FarFragName is
 * derived from the far end's chromosome (UpdateFarFragName, src/pindel.cpp:1262-1270). */
int pg_classify_resident(pg_ctx *ctx, const pg_sv_window *win, const uint8_t *rflags, const uint32_t *close_run_off,
                         const pg_run *close_runs, pg_sv_results *res);
uint64_t pg_classify_launches(pg_ctx *ctx);   /* classification-kernel launches since pg_create */
double pg_classify_last_kernel_ms(pg_ctx *ctx); /* CUDA-event duration of the last classification kernel */

/* The BAM record index of the inflated stream the last pg_bgzf_inflate left on the device: for every alignment record the
 * offset of its block_size field, refID, pos, end (pos + reference bases of the CIGAR, as bam_endpos), flag and mapq --
 * what htslib derives record by record behind sam_itr_next for its region test (src/reader.cpp:1258-1346).  The chain of
 * block_size fields is walked in 256 KiB segments in parallel from guessed record starts; the guesses are checked against
 * the walks and a wrong one is walked again, so the table is that of the serial walk.  first_record / n_ref (optional):
 * offset of the first record and number of references from the BAM header.  PG_ERR_STATE: nothing inflated;
 * PG_ERR_DATA: not a BAM stream or a truncated record chain; PG_ERR_RUNS_CAP: cap too small (*n_recs = needed). */
typedef struct pg_bam_rec {
   uint64_t off;
   int32_t tid, pos, end;
   uint16_t flag;
   uint8_t mapq, pad;
} pg_bam_rec;
int pg_bam_index(pg_ctx *ctx, pg_bam_rec *recs, uint64_t cap, uint64_t *n_recs, uint64_t *first_record, int32_t *n_ref);
/* The same table for a run of whole BGZF blocks cut out of a file -- what htslib reads for one window: the blocks between the
 * virtual offsets hts_itr_t holds for the region (iter->off[i].u / .v; coffset = voffset >> 16, uoffset = voffset & 0xffff) --
 * after pg_bgzf_inflate of just those blocks.  No BAM header is in front, so the caller passes first_record (the uoffset of the
 * chunk's first record inside the first block) and n_ref (bam_hdr_t::n_targets).  The last record may continue in a block that
 * was not read: the chain stops in front of it and *tail tells where it starts (== the stream length when nothing is cut off),
 * so the caller can read on from there.  pg_bam_rp_discovery / pg_bam_sr_candidates then work on this table as on a whole
 * file's (record offsets are relative to the run).  PG_ERR_DATA: the chain from first_record meets a block_size that cannot
 * be a record. */
int pg_bam_index_run(pg_ctx *ctx, uint64_t first_record, int32_t n_ref, pg_bam_rec *recs, uint64_t cap, uint64_t *n_recs, uint64_t *tail);
uint64_t pg_bam_launches(pg_ctx *ctx);
double pg_bam_last_ms(pg_ctx *ctx);            /* CUDA-event duration of the last pg_bam_index (kernels + its small copies) */
void pg_bam_last_segments(pg_ctx *ctx, uint32_t *segments, uint32_t *walked_again);

/* Discordant-pair discovery on the indexed stream (first half): what
 * ReadInBamReads_RP_Discovery (src/reader.cpp:418-480) collects for one BAM file and window -- every record the region
 * iterator delivers (refID == tid, pos < end, end-of-alignment > beg, up to the first record past the region) through
 * build_record_RP_Discovery (src/reader.cpp:982-1097).  rows: Reads_RP_Discovery in BAM order (mates ordered by position);
 * inter_rows: Reads_RP_Discovery_InterChr (pos_a = the read's position, pos_b = its mate's).  src / inter_src: the index
 * of the record in pg_bam_index's table, for the fields the host adds itself (ReadName, reference names, MQA = mapq, Tag).
 * n_scanned: records the iterator delivered (g_NumReadScanned's increase).  Needs pg_bgzf_inflate + pg_bam_index first. */
typedef struct pg_rp_query {
   int32_t tid;                   /* bam_name2id(header, FragName) */
   int32_t beg, end;              /* the window: sam_itr_queryi(idx, tid, window.getStart(), window.getEnd()) */
   int32_t insert_size;           /* the BAM file's insert size (configuration file) */
   uint32_t min_anchor_quality;   /* userSettings->minimalAnchorQuality (-A, default 0) */
} pg_rp_query;
struct pg_rp;
int pg_bam_rp_discovery(pg_ctx *ctx, const pg_rp_query *q, struct pg_rp *rows, uint32_t *src, uint64_t cap, uint64_t *n_rows,
                        struct pg_rp *inter_rows, uint32_t *inter_src, uint64_t inter_cap, uint64_t *n_inter, uint64_t *n_scanned);

/* The reader's split-read filter on the indexed stream (rest of row N3): what ReadInBamReads_SR
 * (src/reader.cpp:483-560) hands to its ReadBuffer for one BAM file and window.  Every record the region iterator delivers goes
 * through fetch_func_SR (src/reader.cpp:1099-1151): records are paired on their read name as they stream by, reads that
 * isWeirdRead accepts (src/reader.cpp:659-690) become candidates behind themselves and behind a mate isGoodAnchor accepts
 * (:561-612) -- build_record_SR (:799-900): sequence decoded, N-trimmed, reverse-complemented for reverse-strand reads, N
 * filter, the anchor's strand and position -- and reads isRefRead accepts (:619-654) are counted as reference-supporting
 * (build_record_RefRead, :903-923).  cands / refs come back in the order of the reference's addRead / push_back calls;
 * candidate k's UnmatchedSeq is seqs[k * stride .. + read_length) (stride: the longest sequence, returned; seqs holds
 * seq_cap bytes).  The host adds Name ("@" + the read's name + "/1" | "/2"), read_group (the anchor's RG tag), Tag, FragName and
 * InsertSize itself from the records `anchor` / `read` index.  n_scanned: records delivered (g_NumReadScanned's increase).
 * PG_ERR_RUNS_CAP: a buffer is too small (*n_cands, *n_refs, *stride hold the required sizes; pass NULL buffers to ask);
 * PG_ERR_DATA: a candidate's sequence is not shorter than insert_size -- the reference stops the program there
 * (src/reader.cpp:866-874) -- or a record of the window is malformed (its name, CIGAR and sequence do not fit its block_size:
 * htslib's bam_read1 fails on it); PG_ERR_STATE: no indexed stream is resident (pg_bgzf_inflate + pg_bam_index). */
typedef struct pg_sr_query {
   int32_t tid;                   /* bam_name2id(header, FragName) */
   int32_t beg, end;              /* the window: sam_itr_queryi(idx, tid, window.getStart(), window.getEnd()) */
   int32_t insert_size;           /* the BAM file's insert size (configuration file) */
   uint32_t min_anchor_quality;   /* userSettings->minimalAnchorQuality (-A) */
   uint32_t chr_len;              /* CurrentChrSeq->size() - 2 * g_SpacerBeforeAfter: MatchedRelPos is clamped to it */
   int32_t nm_limit;              /* userSettings->NM */
   uint32_t pad;
   double max_mismatch_rate;      /* userSettings->MaximumAllowedMismatchRate (-u) */
} pg_sr_query;
typedef struct pg_sr_cand {
   uint32_t anchor, read;         /* indices into pg_bam_index's table: mapped_read, unmapped_read of build_record_SR */
   uint32_t matched_rel_pos;      /* MatchedRelPos */
   uint32_t read_length;          /* UnmatchedSeq.size() */
   uint8_t matched_d;             /* MatchedD '+' / '-' */
   uint8_t ms;                    /* MS */
   uint16_t pad;
} pg_sr_cand;
typedef struct pg_sr_ref {
   uint32_t pos;                  /* REF_READ::Pos */
   int16_t read_length;           /* REF_READ::ReadLength */
   uint8_t mq, pad;               /* REF_READ::MQ */
} pg_sr_ref;
int pg_bam_sr_candidates(pg_ctx *ctx, const pg_sr_query *q, pg_sr_cand *cands, uint64_t cand_cap, uint64_t *n_cands, char *seqs, uint64_t seq_cap,
                         uint32_t *stride, pg_sr_ref *refs, uint64_t ref_cap, uint64_t *n_refs, uint64_t *n_scanned);
/* The same candidates, searched where they lie: the candidate rows become the resident batch of the search kernel on the
 * device (sequences at their stride, anchors from the rows, InsertSize = q->insert_size, chromosome chr_id of
 * pg_set_reference) and GetCloseEnd runs on them at once -- what ReadBuffer::addRead + flush (src/read_buffer.cpp:36-105)
 * do with the reads build_record_SR hands over, without the sequences travelling to the host and back.  Returns what
 * pg_bam_sr_candidates returns plus results->out[n_cands] / runs as pg_close_end does for the same reads in the same order
 * (UP_Close before CleanUniquePoints, PG_READ_NFLIPS).  The filter and the search run once per inflated stream, query and
 * chr_id: the size-asking call (NULL buffers, PG_ERR_RUNS_CAP with all sizes incl. results->n_runs) and the call that
 * fetches the rows share them, and a pg_bam_sr_candidates call with the same query costs only its copies.  The batch stays
 * resident (pg_classify_resident, pg_device_results). */
int pg_bam_sr_close_end(pg_ctx *ctx, const pg_sr_query *q, int32_t chr_id, pg_sr_cand *cands, uint64_t cand_cap, uint64_t *n_cands, char *seqs,
                        uint64_t seq_cap, uint32_t *stride, pg_sr_ref *refs, uint64_t ref_cap, uint64_t *n_refs, uint64_t *n_scanned, pg_results *res);

/* ------------------------------------------------------------------------------------------------------------------
 * The BreakDancer mask and the cluster lookup (last part): what turns the window's events
 * into the `regions` SearchFarEnd scans first (src/pindel.cpp:1006-1020).
 *   pg_bd_load_region = BDData::loadRegion (src/bddata.cpp:850-946) for SearchWindow(chr, start, end): `events` is the whole
 *                       event list in the reference's order (BDData::m_bdEvents after UpdateBD's sort, sortOnFirstBDCoordinate:
 *                       chromosome NAME, position, then the second coordinate), positions as the reference holds them (padded;
 *                       pos <= pos2 per coordinate -- swapped here if not, as the constructor does), chromosome indices of
 *                       pg_set_reference, chr_name_rank[c] = rank of chromosome c's name in std::string order (the cluster
 *                       builder sorts by name).  No per-position mask is built: the region is two arrays the lookup searches.
 *   far-end searches    pg_far_end / pg_launch(2) on a batch WITHOUT regions take every read's cluster from the loaded region on
 *                       the device (getCorrespondingSearchWindowCluster, src/bddata.cpp:949-967, at close_info.last_abs_loc)
 *                       until pg_bd_unload; batches that bring their own regions are unaffected.
 *   pg_bd_clusters    = the lookup alone, host arrays in and out (region_off [n+1], regions as in pg_reads): for hosts that
 *                       keep the clusters, and for the tests.
 * A close end at the last swept position or beyond reads mask entries the reference never wrote (its mask has one entry more
 * than its sweep fills, and the range check admits one more): defined here as "no cluster".  PG_ERR_UNSUPPORTED: a cluster of
 * more than 255 windows. */
typedef struct pg_bd_event {
   int32_t chr_a;
   uint32_t pos_a, pos2_a;        /* BreakDancerEvent::first  (.position, .position2) */
   int32_t chr_b;
   uint32_t pos_b, pos2_b;        /* BreakDancerEvent::second */
} pg_bd_event;
int pg_bd_load_region(pg_ctx *ctx, const pg_bd_event *events, uint32_t n_events, const int32_t *chr_name_rank, int32_t chr, uint32_t start, uint32_t end);
int pg_bd_unload(pg_ctx *ctx);
int pg_bd_clusters(pg_ctx *ctx, uint32_t n, const uint32_t *close_last_abs_loc, uint32_t *region_off, pg_region *regions, uint64_t cap, uint64_t *n_regions);
uint64_t pg_bd_launches(pg_ctx *ctx);
uint64_t pg_bd_lookups(pg_ctx *ctx);          /* reads whose cluster was looked up on the device since pg_create */

/* ------------------------------------------------------------------------------------------------------------------
 * Discordant-read-pair clustering: the two quadratic loops of BDData::UpdateBD
 * (src/bddata.cpp:646-659) that turn the window's discordant pairs into BreakDancer-like events for the far-end search.
 *   pg_rp_modify    = ModifyRP (src/bddata.cpp:385-431) after its std::sort: the caller sorts (Compare2RP,
 *                     src/bddata.cpp:239-253 -- std::sort's order among equal keys is the caller's library's business) and
 *                     passes the rows in that order; InitializeA1B1, the interval-merging double loop in serial row order,
 *                     the + ReadLength pass and the Visited mark for pairs closer than 500 come back.
 *   pg_rp_summarize = Summarize (src/bddata.cpp:470-575): identical events are merged into their first row
 *                     (NumberOfIdentical, Visited; merged_into tells the host whose Tags to append), Report for >= 5.
 * Rows of one call belong to one chromosome pair (ChrNameA == ChrNameB: Reads_RP_Discovery).
 * ------------------------------------------------------------------------------------------------------------------ */
typedef struct pg_rp {            /* RP_READ (src/pindel.h:215-262): what the clustering reads */
   uint32_t pos_a, pos_b;         /* PosA, PosB (== OriginalPosA, OriginalPosB on entry) */
   uint32_t insert_size;          /* InsertSize */
   int16_t read_length;           /* ReadLength */
   uint8_t da, db;                /* DA, DB: '+' / '-' */
} pg_rp;

typedef struct pg_rp_cluster {    /* what ModifyRP / Summarize leave in the RP_READ */
   uint32_t pos_a, pos_a1, pos_b, pos_b1;
   uint32_t n_identical;          /* NumberOfIdentical */
   uint32_t merged_into;          /* Summarize: the row that absorbed this one (append this row's Tags there), else its own index */
   uint8_t visited, report;       /* Visited, Report */
   uint16_t pad;
} pg_rp_cluster;

int pg_rp_modify(pg_ctx *ctx, int32_t n, const pg_rp *in, pg_rp_cluster *out);
int pg_rp_summarize(pg_ctx *ctx, int32_t n, const pg_rp *in, pg_rp_cluster *inout);
/* The inter-chromosomal variants, ModifyRP_InterChr (src/bddata.cpp:439-468, rows again in the order its sort leaves) and
 * Summarize_InterChr (src/bddata.cpp:577-624), for Reads_RP_Discovery_InterChr.  chr_a / chr_b: ChrNameA / ChrNameB as numbers
 * (equal numbers <=> equal names; below 2^24).  A row also meets the rows of the opposite orientation (Y, X, DB, DA). */
int pg_rp_modify_interchr(pg_ctx *ctx, int32_t n, const pg_rp *in, const int32_t *chr_a, const int32_t *chr_b, pg_rp_cluster *out);
int pg_rp_summarize_interchr(pg_ctx *ctx, int32_t n, const pg_rp *in, const int32_t *chr_a, const int32_t *chr_b, pg_rp_cluster *inout);
uint64_t pg_rp_launches(pg_ctx *ctx);          /* clustering-kernel launches since pg_create */
/* how the last pg_rp_modify split its rows: classes run by one warp each, classes run by one CTA each, rows in either */
void pg_rp_last_classes(pg_ctx *ctx, uint32_t *warp_classes, uint32_t *cta_classes, uint32_t *rows_in_classes);
double pg_rp_last_kernel_ms(pg_ctx *ctx);      /* CUDA-event duration of the last call's kernels */

/* ------------------------------------------------------------------------------------------------------------------
 * BGZF inflation (first half): what htslib's bgzf_read_block / inflate_block do behind
 * sam_itr_next (src/reader.cpp:1099-1151, 1258-1346), for all blocks of a file (or of any run of whole blocks) at once:
 * the host walks the block headers (BSIZE, ISIZE, CRC32), one device thread inflates one block (RFC 1951) into its place
 * in the inflated stream and, with check_crc, recomputes the payload's CRC32 as htslib does.
 *   out == NULL: the inflated stream stays on the device (pg_bgzf_device_data) for a consumer that parses it there;
 *   out != NULL: it is also copied to the caller's buffer (out_cap bytes; PG_ERR_RUNS_CAP if too small).
 * *out_len = inflated size.  PG_ERR_DATA: not a BGZF file / corrupt block (text in pg_last_error).
 * Damaged files: Huffman code sets are judged as zlib's inflate_table judges them -- over-subscribed sets are refused, and
 * incomplete ones too except the empty set and a single one-bit literal/length or distance code -- so a corrupted
 * dynamic block is refused here when zlib refuses it, with or without check_crc (tests/test_bgzf_emulator.py builds such
 * blocks bit by bit), and nothing is decoded from beyond the end of a block's compressed bytes (a cut-off stream, or a stored
 * block whose LEN reaches past it, is "compressed data exhausted").  On 400 000 randomly damaged streams of every block type the
 * decoder and zlib agree in every case on refuse / inflate and on the bytes (scripts/bgzf_corruption_campaign.py).
 * check_crc = 1 is what htslib does on every block and is the setting of the drop-in.  Block headers are judged as htslib's
 * check_header judges them (magic, CM = 8, FLG & 4, a 'B','C' subfield of length 2 in the extra field; the other flag bits are not
 * looked at), so a file refused here is one htslib refuses to read as BGZF.
 * ------------------------------------------------------------------------------------------------------------------ */
int64_t pg_bgzf_inflated_size(const uint8_t *file, size_t file_len);   /* host only: -1 if the header walk fails */
int pg_bgzf_inflate(pg_ctx *ctx, const uint8_t *file, size_t file_len, uint8_t *out, size_t out_cap, size_t *out_len, int32_t check_crc);
int pg_bgzf_device_data(pg_ctx *ctx, void **dev, size_t *len);         /* the last inflated stream; valid until the next call */
uint64_t pg_bgzf_launches(pg_ctx *ctx);
double pg_bgzf_last_kernel_ms(pg_ctx *ctx);

/* ------------------------------------------------------------------------------------------------------------------
 * Per-box read ordering and duplicate marking: bubblesortReads + markDuplicates
 * (src/reporter.cpp:932-972), which the D, SI and TD reporters run on every box with at least NumRead2ReportCutOff reads
 * (src/reporter.cpp:996-998, 1179-1180, 1425-1427), for all boxes of a driver at once.  The exchange sort's order among
 * equal keys is part of the output, so it is replayed, not replaced: order[box_off[b] + k] = which member of box b stands
 * at position k afterwards; nonunique[box_off[b] + k] = 1 when that read's UniqueRead is false afterwards.
 * keys: one per member, in box order; all reads of a call on one chromosome (smaller() compares FragName first; the caller
 * keeps boxes with mixed FragNames on the host).  name_id: equal <=> equal Name.
 *
 * pg_box_sort_kind: the same for every reporter.  The DI and inversion reporters carry the two loops inlined with their own
 * rules (SortOutputDI, src/reporter.cpp:1728-1775; OutputSorter::DoSortAndOutputInversions, src/output_sorter.cpp:74-131): a
 * STRICT exchange (equal reads are not swapped) under another comparator, and a duplicate scan on LeftMostPos / read end /
 * anchor strand from every first read.  For those kinds `left` carries LeftMostPos, and read_length / matched_d are read.
 * event (may be NULL): the cut of the surviving reads into events that every reporter makes next (the "event merge",
 * src/reporter.cpp:1016-1034, 1206-1236, 1446-1487, 1790-1813, src/output_sorter.cpp:149-205): per position of the final
 * order PG_BOX_EVENT_GOOD when that read enters GoodIndels (unique reads; every read for the inversion kinds) and
 * PG_BOX_EVENT_START when the reporter's event key changes there (D / TD: BPLeft, BPRight; SI: BPLeft, IndelSize; DI: BPLeft,
 * IndelSize, NT_size; INV: BPLeft + BPRight).  pg_box_sort = pg_box_sort_kind(PG_BOX_D, ..., NULL).
 * ------------------------------------------------------------------------------------------------------------------ */
typedef struct pg_sort_key {
   uint32_t bp_left, bp_right, indel_size;   /* BPLeft, BPRight, IndelSize */
   int32_t left, right;                      /* Left, Right (PG_BOX_D / _SI); LeftMostPos in `left` for the other kinds */
   uint32_t name_id;
   uint16_t nt_size;                         /* NT_size */
   int16_t bp;                               /* BP */
   uint8_t unique;                           /* UniqueRead on entry */
   uint8_t matched_d;                        /* MatchedD ('+' / '-'); PG_BOX_DI and the inversion kinds */
   uint16_t read_length;                     /* getReadLength(); PG_BOX_DI and the inversion kinds */
} pg_sort_key;
enum { PG_BOX_D = 0,        /* SortOutputD, SortAndOutputTandemDuplications */
       PG_BOX_SI = 1,       /* SortOutputSI */
       PG_BOX_DI = 2,       /* SortOutputDI */
       PG_BOX_INV = 3,      /* OutputSorter::SortAndOutputInversions */
       PG_BOX_INV_NT = 4 }; /* OutputSorter::SortAndOutputNonTemplateInversions */
#define PG_BOX_EVENT_GOOD 1u
#define PG_BOX_EVENT_START 2u
int pg_box_sort(pg_ctx *ctx, uint32_t n_boxes, const uint32_t *box_off, const pg_sort_key *keys, uint32_t *order, uint8_t *nonunique);
int pg_box_sort_kind(pg_ctx *ctx, int32_t kind, uint32_t n_boxes, const uint32_t *box_off, const pg_sort_key *keys, uint32_t *order, uint8_t *nonunique,
                     uint8_t *event);
uint64_t pg_box_sort_launches(pg_ctx *ctx);
double pg_box_sort_last_kernel_ms(pg_ctx *ctx);

/* Host helper: run-length encode UniquePoints (the inverse of pg_expand_runs) for callers that hold expanded lists.
 * Returns the number of runs (out may be NULL to just count). */
uint32_t pg_compress_ups(const pg_up *ups, uint32_t n, pg_run *out);

/* Host helpers: expand runs to UniquePoints (returns the count; out may be NULL to just count), and
 * CleanUniquePoints (src/pindel.cpp:2904-2941) in place (returns the new count). */
uint32_t pg_expand_runs(const pg_run *runs, uint32_t n_runs, pg_up *out);
uint32_t pg_clean_unique_points(pg_up *ups, uint32_t n);

#ifdef __cplusplus
}
#endif
#endif
