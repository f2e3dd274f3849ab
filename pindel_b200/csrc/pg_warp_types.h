// pindel_b200/csrc/pg_warp_types.h -- plain-data types shared by the warp program (pg_warp.cuh),
// the kernels, the C-ABI and host-only helpers.  No device code here.
#pragma once
#include <stdint.h>
#if !defined(__CUDACC__) && !defined(PG_EMULATE) && !defined(PG_HAVE_UINT2)
#define PG_HAVE_UINT2
struct uint2 { uint32_t x, y; };
#endif

// ------------------------------------------------------------------------------------------------
// Layout constants
// ------------------------------------------------------------------------------------------------
#define PG_SPACER 100000u   // g_SpacerBeforeAfter, src/pindel.h:122
#define PG_PADW 64          // guard words on both sides of every packed chromosome
#define PG_TILE_POS 2048    // reference positions per shared-memory tile
#define PG_TILE_WORDS 76    // 64 + 1 before + 1 after + alignment slack (multiple of 4)
#define PG_CHUNK_POS 1024   // positions per seed-list chunk (one 32-base word per lane)
#define PG_QCAP 64          // deferred ("slow") seed queue per warp
#define PG_IDX_BITS 20
#define PG_IDX_MASK 0xFFFFFu
#ifndef PG_FILTER_MIN_POS
#define PG_FILTER_MIN_POS 2048u  // far rings at least this long go through the threshold filter (pg_filter_tile): from the fourth window on (-x >= 3); measured: at 512 the filter saves instructions at -x 2 but its code no longer fits the instruction cache beside the rest
#endif
#define PG_FAR_BP_START 10  // src/farend_searcher.cpp:89
#define PG_TABLE 500        // g_maxMismatch size, src/pindel.cpp:801

// flags in PgReadOut.flags
#define PG_F_FLIPPED 1u     // UnmatchedSeq was reverse-complemented an odd number of times
#define PG_F_TOOLONG 2u     // read longer than the compiled limit -> not searched
#define PG_F_OVERFLOW 4u    // run pool exhausted -> results of this read dropped
#define PG_F_NFLIPS_SHIFT 8 // bits 8..9: how many times GetCloseEnd called setUnmatchedSeq(ReverseComplement(..)) (0..2)

struct PgChrom {
   const uint2 *planes;     // [nwords] {lo, hi} bit-planes, 32 bases per element; base i -> bit (i & 31) of element i >> 5
   const uint32_t *nplane;  // [nwords] 1 = reference base is not A/C/G/T
   uint32_t len;            // padded length (bases) == std::string size in the reference
   uint32_t nwords;
};

struct PgParams {
   int32_t add_mm;          // -a  ADDITIONAL_MISMATCH
   int32_t min_close;       // -H  g_MinClose
   int32_t min_perfect;     // -m  Min_Perfect_Match_Around_BP
   int32_t max_range_index; // -x  MaxRangeIndex
   uint32_t filter_min_pos; // far rings at least this long go through the threshold filter (PG_FILTER_MIN_POS; tuning knob, results do not depend on it)
   uint8_t g[512];          // g_maxMismatch[len] (src/pindel.cpp:794-819), host-computed in double
   uint16_t thr[512];       // thr[L] = least n with (float)n >= (float)(L * -u), src/searcher.cpp:366,383
};

// One run of consecutive UniquePoints from the same seed with the same mismatch count:
// lengths len0 .. len0+count-1; AbsLoc moves +1 per length for FORWARD, -1 for BACKWARD.
struct PgRun {
   uint32_t abs_loc0;       // UniquePoint::AbsLoc at len0
   uint16_t len0;
   uint16_t count;
   uint8_t mm;              // UniquePoint::Mismatches
   uint8_t dirstrand;       // bit0: 0 FORWARD / 1 BACKWARD, bit1: 0 SENSE / 1 ANTISENSE
   int16_t chr;
};

struct PgReadOut {
   uint32_t run_off;        // first run of this read in the run pool (close runs, then far runs)
   uint16_t n_close_runs;
   uint16_t n_far_runs;
   uint16_t len_after;      // ReadLength after the search (setUnmatchedSeq may strip, src/pindel.cpp:142-169)
   uint16_t flags;
};

struct PgRegion {
   int32_t chr;
   uint32_t start, end;     // [start, end) padded coordinates (SearchWindow)
};

// Far-end-only entry (the reference's SearchFarEnds seam, src/pindel.h:741): what the caller knows
// about the close end of each read.
struct PgCloseInfo {
   uint32_t last_abs_loc;   // UP_Close.back().AbsLoc   (getLastAbsLocCloseEnd, src/pindel.cpp:474-477)
   uint16_t max_len;        // UP_Close.MaxLen()        (0 = no close end -> read skipped)
   uint16_t pad;
};

struct PgBatch {
   int32_t n_reads;
   int32_t mode;                    // bit0 close end, bit1 far end (bit1 alone: close info supplied)
   const uint8_t *seq_blob;         // ASCII read bases
   const uint32_t *seq_off;         // [n+1]
   const uint8_t *matched_d;        // '+' / '-'
   const uint32_t *matched_rel_pos;
   const int16_t *insert_size;
   const int32_t *chr_id;           // may be null (all reads on chromosome 0)
   const PgCloseInfo *close_info;   // far-only mode
   const PgRegion *regions;         // BreakDancer search clusters, may be null
   const uint32_t *region_off;      // [n+1]
   PgReadOut *out;                  // [n]
   PgRun *run_pool;
   uint32_t run_pool_cap;
   uint32_t *counters;              // [8]: [0] next read, [1] runs allocated, [2] overflow flag, [4..5] window bases, [6..7] UniquePoints
   PgRun *scratch;                  // [n_warps * 3 * scratch_cap]
   uint32_t scratch_cap;
};

