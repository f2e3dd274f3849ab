// pindel_b200/csrc/pg_sv_types.h -- plain-data types of the SV-classification stage:
// the seven per-read drivers SearchSVs runs after the far-end search (src/pindel.cpp:1140-1164).  No device code here.
#pragma once
#include <stdint.h>
#include "pg_warp_types.h"

// Drivers, numbered in the order SearchSVs calls them (src/pindel.cpp:1146-1163).
#define PG_SV_D 0        // SearchDeletions::Search            src/search_variant.cpp:49, src/searchdeletions.cpp
#define PG_SV_DI 1       // searchIndels                       src/search_deletions_nt.cpp:26
#define PG_SV_TD 2       // searchTandemDuplications           src/search_tandem_duplications.cpp:28
#define PG_SV_TD_NT 3    // searchTandemDuplicationsNT         src/search_tandem_duplications_nt.cpp:27
#define PG_SV_INV 4      // searchInversions                   src/search_inversions.cpp:29
#define PG_SV_INV_NT 5   // searchInversionsNT                 src/search_inversions_nt.cpp:27
#define PG_SV_SI 6       // SearchShortInsertions::Search      src/search_variant.cpp:49, src/searchshortinsertions.cpp
#define PG_SV_NDRIVERS 7
#define PG_SV_ALL 0x7Fu

// PgSvRecord.outcome
#define PG_SV_NONE 0     // fields assigned, read stays unused (outside the region / box index too large / TD with BPLeft == 0)
#define PG_SV_BOX 1      // Vars[box].push_back(ReadIndex); Used = true
#define PG_SV_FUTURE 2   // saveReadForNextCycle: FutureReads_SR.push_back(read); Used = true   (src/pindel.cpp:587-592)
#define PG_SV_BREAK 3    // breakpoint beyond the chromosome string: Used = true, nothing stored (src/search_variant.cpp:133-136)

// PgSvRecord.fields
#define PG_SVF_CORE 1u   // Left, Right, BP, IndelSize, BPLeft, BPRight
#define PG_SVF_NTSTR 2u  // NT_str
#define PG_SVF_NTSIZE 4u // NT_size
#define PG_SVF_KILLED 8u // LeftMostTD / LeftMostINV found the breakpoint outside the chromosome: BP = BPLeft = BPRight = 1 and
                         // Used = true *before* the action of this record (src/search_tandem_duplications.cpp:196-202)

// PgSvRecord.nt_src
#define PG_NT_EMPTY 0    // ""
#define PG_NT_SEQ 1      // UnmatchedSeq.substr(nt_start, nt_len), rotated left by nt_rot
#define PG_NT_RC 2       // ReverseComplement(UnmatchedSeq).substr(nt_start, nt_len), rotated left by nt_rot

// per-read input flags
#define PG_SVR_USED 1u       // SPLIT_READ::Used on entry
#define PG_SVR_OTHERCHR 2u   // FragName != FarFragName (the drivers skip the read)

// One "fire" of a driver on a read: an arm's test passed, the SPLIT_READ fields were assigned (and adjusted by
// GetRealStart4Deletion/Insertion, LeftMostTD, LeftMostINV), and `outcome` says what happened next.  The host replays
// the records of one driver in (read, seq) order; the values are the reference's own, bit for bit.
struct PgSvRecord {
   uint32_t read;           // index in the batch (ReadIndex)
   uint8_t driver;          // PG_SV_*
   uint8_t outcome;         // PG_SV_NONE / BOX / FUTURE / BREAK
   uint8_t fields;          // PG_SVF_*
   uint8_t seq;             // order of this record among the records of (read, driver)
   int32_t left, right;     // SPLIT_READ::Left, Right
   uint32_t bp_left, bp_right, indel_size;
   int16_t bp;
   uint16_t nt_size;        // SPLIT_READ::NT_size (valid with PG_SVF_NTSIZE)
   uint16_t nt_start, nt_len, nt_rot;
   uint8_t nt_src;          // PG_NT_* (valid with PG_SVF_NTSTR)
   uint8_t pad;
   uint32_t box;            // TempBoxIndex for PG_SV_BOX
};

struct PgSvParams {
   uint32_t driver_mask;    // which drivers run (Analyze_TD / Analyze_INV switch theirs off, src/pindel.cpp:1151-1158)
   int32_t chr;             // chromosome of the window (index in pg_set_reference order)
   uint32_t window_end;     // SearchWindow::getEnd() of the current window
   uint32_t num_boxes;      // NumBoxes   (src/pindel.cpp:1810)
   uint32_t box_size;       // BoxSize    (src/pindel.cpp:1806-1809)
   uint32_t region_start_defined, region_start;   // userSettings->getRegion()   (src/pindel.cpp:567-585)
   uint32_t region_end_defined, region_end;
   int32_t min_inv;         // MIN_IndelSize_Inversion     (src/user_defined_settings.h:68)
   int32_t min_matched;     // Min_Num_Matched_Bases       (-M)
   uint16_t nt_thr[1032];   // nt_thr[k] = (short)(1 + Seq_Error_Rate * k), k = far length + close length, host-computed in double
                            //             (src/search_deletions_nt.cpp:50-51)
};

struct PgSvBatch {
   int32_t n_reads;
   const uint8_t *seq_blob;      // final UnmatchedSeq of every read (after GetCloseEnd's flips)
   const uint32_t *seq_off;      // [n+1]
   const uint8_t *matched_d;     // '+' / '-'
   const int16_t *insert_size;
   const uint8_t *rflags;        // PG_SVR_*
   const uint32_t *run_off;      // [n+1] first run of the read (UP_Close runs, then UP_Far runs)
   const uint16_t *n_close_runs; // [n]
   const PgRun *runs;
   // resident mode (pg_classify_resident): the lists of the search launch that is still on the device replace the three
   // arrays above -- far runs (and, after a close+far launch, close runs) come from res_pool through res_out; after a
   // far-end-only launch the close runs are the caller's (close_off / close_runs).  rflags may be NULL (nothing used).
   const PgReadOut *res_out;     // [n] or NULL
   const PgRun *res_pool;
   const uint32_t *close_off;    // [n+1] or NULL
   const PgRun *close_runs;
   const int32_t *chr_id;        // [n] chromosome of the anchor (FragName) or NULL = 0: FarFragName != FragName is derived
   const uint32_t *perm;         // [n] or NULL: thread t classifies read perm[t] ('+' anchors first: a warp's lanes walk the same arms)
   PgSvRecord *records;
   uint32_t records_cap;
   uint32_t *counters;           // [4]: [0] records allocated, [1] overflow flag, [2] pair tests (statistics)
   uint8_t *counts;              // [PG_SV_NDRIVERS * n_reads] records per (driver, read), driver-major; zeroed before the launch
};
