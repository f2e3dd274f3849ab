// tests/emul/bam_emul.cpp -- TEST INFRASTRUCTURE ONLY: the BAM record index of pindel_b200/csrc/pg_bam.cuh (the functions
// pg_bam_start_kernel / pg_bam_walk_kernel run) and the host's consistency check (pg_bam_host.h), compiled as plain C++ and
// driven segment by segment exactly like pg_bam_index drives the kernels.  No GPU.
#include <stdint.h>
#include <string.h>
#include <vector>
#include "../../pindel_b200/csrc/pg_bam.cuh"
#include "../../pindel_b200/csrc/pg_bam_host.h"

extern "C" {

int bamemu_rec_size(void) { return (int)sizeof(PgBamRec); }
static uint64_t g_last_walks = 0;
uint64_t bamemu_last_walks(void) { return g_last_walks; }

// seg: segment size (0 = the product's).  Returns the number of records (recs may be NULL to count), -1 not BAM,
// -2 corrupt chain.  stats[0] = segments, stats[1] = segments that had to be walked again, stats[2] = first record offset.
// bamemu_last_walks(): segment walks of pass 0 in the last call (= segments + walked again: a repair walks one segment).
long bamemu_index(const uint8_t *d, uint64_t n, uint64_t seg, PgBamRec *recs, uint64_t cap, uint64_t *stats)
{
   int32_t n_ref = 0;
   const uint64_t first = pgb_header(d, n, &n_ref);
   if (first == PGB_NONE) return -1;
   if (seg == 0) seg = PGB_SEGMENT;
   const uint32_t n_seg = first >= n ? 1u : (uint32_t)((n - first + seg - 1) / seg);
   std::vector<uint64_t> starts(n_seg), counts(n_seg), landing(n_seg), base(n_seg);
   for (uint32_t s = 0; s < n_seg; s++) {         // pg_bam_start_kernel
      if (s == 0) { starts[0] = first; continue; }
      const uint64_t b0 = first + (uint64_t)s * seg, b1 = b0 + seg < n ? b0 + seg : n;
      starts[s] = PGB_NONE;
      for (uint64_t c = b0; c < b1; c++)
         if (pgb_plausible_chain(d, n, c, n_ref)) { starts[s] = c; break; }
   }
   uint64_t total = 0;
   uint32_t redone = 0, walks = 0;
   auto walk_one = [&](uint32_t s) {              // pg_bam_walk_kernel, pass 0, one segment
      walks++;
      if (starts[s] == PGB_NONE) { counts[s] = 0; landing[s] = PGB_NONE; return 0; }
      const uint64_t b1 = first + (uint64_t)(s + 1) * seg;
      uint64_t k = 0;
      landing[s] = pgb_walk(d, n, starts[s], b1 < n ? b1 : n, NULL, &k);
      counts[s] = landing[s] == PGB_NONE ? 0 : k;
      return 0;
   };
   auto walk_all = [&]() { for (uint32_t s = 0; s < n_seg; s++) walk_one(s); return 0; };
   auto validate = [&](uint64_t *forced) { return pgb_validate(n_seg, first, n, seg, starts.data(), counts.data(), landing.data(), forced, &total); };
   if (pgb_settle(starts.data(), validate, walk_all, walk_one, &redone) != 0) return -2;
   g_last_walks = walks;
   if (stats) { stats[0] = n_seg; stats[1] = redone; stats[2] = first; }
   if (!recs) return (long)total;
   if (cap < total) return -3;
   uint64_t b = 0;
   for (uint32_t s = 0; s < n_seg; s++) { base[s] = b; b += counts[s]; }
   for (uint32_t s = 0; s < n_seg; s++) {         // pass 1
      if (counts[s] == 0) continue;
      const uint64_t b1 = first + (uint64_t)(s + 1) * seg;
      uint64_t k = 0;
      pgb_walk(d, n, starts[s], b1 < n ? b1 : n, recs + base[s], &k);
   }
   return (long)total;
}

// pg_bam_index_run: a headerless run of blocks whose last record may be cut off.  Returns the number of records, -2 for a
// chain that meets an impossible block_size; stats[3] = the tail offset.
long bamemu_index_run(const uint8_t *d, uint64_t n, uint64_t first, int32_t n_ref, uint64_t seg, PgBamRec *recs, uint64_t cap, uint64_t *stats)
{
   if (seg == 0) seg = PGB_SEGMENT;
   const uint32_t n_seg = first >= n ? 1u : (uint32_t)((n - first + seg - 1) / seg);
   std::vector<uint64_t> starts(n_seg), counts(n_seg), landing(n_seg), base(n_seg);
   for (uint32_t s = 0; s < n_seg; s++) {         // pg_bam_start_kernel
      if (s == 0) { starts[0] = first; continue; }
      const uint64_t b0 = first + (uint64_t)s * seg, b1 = b0 + seg < n ? b0 + seg : n;
      starts[s] = PGB_NONE;
      for (uint64_t c = b0; c < b1; c++)
         if (pgb_plausible_chain(d, n, c, n_ref)) { starts[s] = c; break; }
   }
   uint64_t total = 0, tail = n;
   uint32_t redone = 0, walks = 0;
   auto walk_one = [&](uint32_t s) {              // pg_bam_walk_run_kernel, pass 0, one segment
      walks++;
      if (starts[s] == PGB_NONE) { counts[s] = 0; landing[s] = PGB_NONE; return 0; }
      const uint64_t b1 = first + (uint64_t)(s + 1) * seg;
      uint64_t k = 0;
      landing[s] = pgb_walk_run(d, n, starts[s], b1 < n ? b1 : n, NULL, &k);
      counts[s] = landing[s] == PGB_NONE ? 0 : k;
      return 0;
   };
   auto walk_all = [&]() { for (uint32_t s = 0; s < n_seg; s++) walk_one(s); return 0; };
   auto validate = [&](uint64_t *forced) {
      return pgb_validate_run(n_seg, first, n, seg, starts.data(), counts.data(), landing.data(), forced, &total, &tail);
   };
   if (pgb_settle(starts.data(), validate, walk_all, walk_one, &redone) != 0) return -2;
   g_last_walks = walks;
   if (stats) { stats[0] = n_seg; stats[1] = redone; stats[2] = first; stats[3] = tail; }
   if (!recs) return (long)total;
   if (cap < total) return -3;
   uint64_t b = 0;
   for (uint32_t s = 0; s < n_seg; s++) { base[s] = b; b += counts[s]; }
   for (uint32_t s = 0; s < n_seg; s++) {         // pass 1
      if (counts[s] == 0) continue;
      const uint64_t b1 = first + (uint64_t)(s + 1) * seg;
      uint64_t k = 0;
      pgb_walk_run(d, n, starts[s], b1 < n ? b1 : n, recs + base[s], &k);
   }
   return (long)total;
}

// pg_bam_rp_discovery in the kernels' formulation: the stop index, the class of every record before it, positions by
// prefix count, scatter.  rows/src hold the intra rows then (from n_intra on) nothing; inter_* the inter-chromosomal ones.
// counts[0] = intra rows, [1] = inter rows, [2] = records delivered.
void bamemu_rp_discovery(const uint8_t *d, const PgBamRec *recs, uint32_t n_recs, const PgRpQuery *q, PgRpRow *rows, uint32_t *src,
                         PgRpRow *inter_rows, uint32_t *inter_src, uint64_t *counts)
{
   uint32_t stop = n_recs;
   for (uint32_t i = 0; i < n_recs; i++)
      if (pgb_itr_stops(recs[i], *q)) { stop = i; break; }
   uint64_t a = 0, b = 0, delivered = 0;
   for (uint32_t i = 0; i < stop; i++) {
      if (!pgb_itr_delivers(recs[i], *q)) continue;
      delivered++;
      PgRpRow row;
      const int cls = pgb_rp_discovery(d, recs[i], *q, row);
      if (cls == 1) { rows[a] = row; src[a++] = i; }
      else if (cls == 2) { inter_rows[b] = row; inter_src[b++] = i; }
   }
   counts[0] = a; counts[1] = b; counts[2] = delivered;
}
}
