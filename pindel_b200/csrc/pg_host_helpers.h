// pindel_b200/csrc/pg_host_helpers.h -- the pure host helpers of the C ABI (no device work): run expansion,
// CleanUniquePoints and run-length coding of UniquePoint lists.  Shared by pindel_pg.cu and the CPU-side test harness
// (tests/emul/abi_emul.cpp).
#pragma once
#include <stdint.h>
#include "../../include/pindel_pg.h"

static inline uint32_t pgh_expand_runs(const pg_run *runs, uint32_t n_runs, pg_up *out)
{
   uint32_t k = 0;
   for (uint32_t r = 0; r < n_runs; r++) {
      const pg_run &q = runs[r];
      for (uint32_t j = 0; j < q.count; j++, k++) {
         if (!out) continue;
         pg_up u;
         u.len = (int16_t)(q.len0 + j);
         u.abs_loc = (q.dirstrand & 1u) ? q.abs_loc0 - j : q.abs_loc0 + j;
         u.mm = q.mm;
         u.dir = q.dirstrand & 1u;
         u.strand = (q.dirstrand >> 1) & 1u;
         u.chr = q.chr;
         out[k] = u;
      }
   }
   return k;
}

// CleanUniquePoints, src/pindel.cpp:2904-2941: keep the points that share chromosome, direction, strand and
// terminal (AbsLoc -/+ LengthStr) with the LAST point.
static inline uint32_t pgh_clean_unique_points(pg_up *ups, uint32_t n)
{
   if (!ups || n == 0) return 0;
   pg_up last = ups[n - 1];
   uint32_t terminal = last.dir == 0 ? last.abs_loc - (uint32_t)last.len : last.abs_loc + (uint32_t)last.len;
   uint32_t w = 0;
   for (uint32_t i = 0; i < n; i++) {
      if (ups[i].chr != last.chr || ups[i].dir != last.dir || ups[i].strand != last.strand) continue;
      uint32_t t = last.dir == 0 ? ups[i].abs_loc - (uint32_t)ups[i].len : ups[i].abs_loc + (uint32_t)ups[i].len;
      if (t == terminal) ups[w++] = ups[i];
   }
   return w;
}

// Run-length encode an expanded UniquePoint list: consecutive points continue a run when the length grows by one, the
// position moves by one in the run's direction and chromosome, direction, strand and mismatch count stay.
static inline uint32_t pgh_compress_ups(const pg_up *ups, uint32_t n, pg_run *out)
{
   uint32_t k = 0;
   for (uint32_t i = 0; i < n;) {
      uint32_t j = i + 1;
      while (j < n && j - i < 65535u && ups[j].len == ups[j - 1].len + 1 && ups[j].mm == ups[i].mm && ups[j].dir == ups[i].dir &&
             ups[j].strand == ups[i].strand && ups[j].chr == ups[i].chr &&
             ups[j].abs_loc == (ups[i].dir ? ups[j - 1].abs_loc - 1u : ups[j - 1].abs_loc + 1u))
         j++;
      if (out) {
         pg_run r;
         r.abs_loc0 = ups[i].abs_loc;
         r.len0 = (uint16_t)ups[i].len;
         r.count = (uint16_t)(j - i);
         r.mm = (uint8_t)ups[i].mm;
         r.dirstrand = (uint8_t)((ups[i].dir & 1) | ((ups[i].strand & 1) << 1));
         r.chr = ups[i].chr;
         out[k] = r;
      }
      k++;
      i = j;
   }
   return k;
}

