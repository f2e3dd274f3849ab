// pindel_b200/csrc/pg_bam_host.h -- host part of pg_bam_index: checking the segments' guessed record starts against the
// segments' walks (see pg_bam.cuh).  Shared by pindel_pg.cu and the CPU-side test harness (tests/emul/bam_emul.cpp).
#pragma once
#include <stdint.h>
#include "pg_bam.cuh"

#define PGB_SEGMENT (256u << 10)   // bytes of inflated stream per segment

// starts[0] is the true first record.  Follows the true chain segment by segment:
//   a segment the chain jumps over keeps no records (whatever it guessed);
//   a segment whose guess is the chain's next record is accepted, the chain continues at its landing offset;
//   otherwise the segment must be walked again from the chain's offset: returns 1 + s and that offset in *forced.
// Returns 0 when every segment is consistent and the chain ends exactly at n (*total = records), -1 for a corrupt chain.
static inline long pgb_validate(uint32_t n_seg, uint64_t first, uint64_t n, uint64_t seg, const uint64_t *starts, uint64_t *counts,
                                const uint64_t *landing, uint64_t *forced, uint64_t *total)
{
   uint64_t cur = first, sum = 0;
   for (uint32_t s = 0; s < n_seg; s++) {
      const uint64_t b1 = first + (uint64_t)(s + 1) * seg;
      if (cur >= n || cur >= b1) {
         counts[s] = 0;
         continue;
      }
      if (starts[s] != cur) {
         *forced = cur;
         return 1 + (long)s;
      }
      if (landing[s] == PGB_NONE) return -1;
      sum += counts[s];
      cur = landing[s];
   }
   if (cur != n) return -1;
   *total = sum;
   return 0;
}

// The same for a run of blocks (pgb_walk_run): the chain may end in front of a record that is cut off by the end of the run.
// A segment whose walk stopped before its own end (landing < min(segment end, n)) is where that happens: *tail = its landing
// offset, later segments keep nothing.  *tail == n when the run ends on a record boundary.
static inline long pgb_validate_run(uint32_t n_seg, uint64_t first, uint64_t n, uint64_t seg, const uint64_t *starts, uint64_t *counts,
                                    const uint64_t *landing, uint64_t *forced, uint64_t *total, uint64_t *tail)
{
   uint64_t cur = first, sum = 0;
   bool ended = false;
   for (uint32_t s = 0; s < n_seg; s++) {
      const uint64_t b1 = first + (uint64_t)(s + 1) * seg, lim = b1 < n ? b1 : n;
      if (ended || cur >= n || cur >= b1) {
         counts[s] = 0;
         continue;
      }
      if (starts[s] != cur) {
         *forced = cur;
         return 1 + (long)s;
      }
      if (landing[s] == PGB_NONE) return -1;
      sum += counts[s];
      cur = landing[s];
      if (cur < lim) ended = true;          // stopped in front of a record that does not fit
   }
   if (cur > n) return -1;
   *total = sum;
   *tail = cur;
   return 0;
}

// The settling loop of pg_bam_index / pg_bam_index_run, shared with the CPU-side harness so that its logic is tested without
// a device: walk_all() walks every segment from its guessed start (pass 0: counts, landing); validate(&forced) is
// pgb_validate / pgb_validate_run over the host copies; a segment whose guess was not the chain's next record gets the
// chain's offset as its start and walk_one(s) walks that segment alone -- the segments in front of it are settled and the
// ones behind it keep their walks, so a stream full of fake records costs one single-segment walk per bad guess instead of
// a walk over all segments (at most n_seg repairs: each one settles its segment for good).
// Returns 0, -1 for a corrupt chain, or the first non-zero value a walker returned.  *redone = segments walked again.
template <class Validate, class WalkAll, class WalkOne>
static inline int pgb_settle(uint64_t *starts, Validate validate, WalkAll walk_all, WalkOne walk_one, uint32_t *redone)
{
   *redone = 0;
   int rc = walk_all();
   if (rc != 0) return rc;
   for (;;) {
      uint64_t forced = 0;
      const long v = validate(&forced);
      if (v < 0) return -1;
      if (v == 0) return 0;
      const uint32_t s = (uint32_t)(v - 1);
      starts[s] = forced;
      if ((rc = walk_one(s)) != 0) return rc;
      ++*redone;
   }
}
