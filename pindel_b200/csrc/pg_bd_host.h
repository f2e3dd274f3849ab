// pindel_b200/csrc/pg_bd_host.h -- host side of the BreakDancer mask (pg_bd.cuh): what BDData::loadRegion's lower_bound /
// upper_bound select (src/bddata.cpp:879-885) and the two arrays the per-read lookup searches.  O(m) per window.
#pragma once
#include <stdint.h>
#include <vector>

#include "pg_bd.cuh"

struct PgBdPrepared {
   std::vector<uint32_t> start, drop;
   std::vector<PgBdSel> sel;
   uint32_t z = 0, ws = 0, we = 0;
};

// events: the whole event list in the reference's order (sortOnFirstBDCoordinate, src/control_state.cpp:111-120: chromosome NAME,
// position, then the second coordinate), chr_rank[c]: rank of chromosome c's name in string order.  Returns 0, or -1 when the
// events of chromosome `chr` are not one ascending block.
static inline int pgbd_prepare(const PgBdEvent *events, uint32_t n_events, const int32_t *chr_rank, int32_t n_chr, int32_t chr, uint32_t start, uint32_t end,
                               PgBdPrepared &P)
{
   P.start.clear(); P.drop.clear(); P.sel.clear();
   P.ws = start >= 3u * PGBD_INSERT ? start - 3u * PGBD_INSERT : 0u;      // src/bddata.cpp:857-863
   P.we = end + 3u * PGBD_INSERT;                                         // :872
   P.z = 0;
   // the block of events whose first coordinate lies on `chr`
   uint32_t b0 = n_events, b1 = n_events;
   for (uint32_t i = 0; i < n_events; i++) {
      if (events[i].chr_a < 0 || events[i].chr_a >= n_chr || events[i].chr_b < 0 || events[i].chr_b >= n_chr) return -1;
      if (events[i].chr_a == chr) { if (b0 == n_events) b0 = i; b1 = i + 1; }
   }
   if (b0 == n_events) return 0;
   uint32_t prev = 0;
   std::vector<uint32_t> endw;
   for (uint32_t i = b0; i < b1; i++) {
      const PgBdEvent &e = events[i];
      if (e.chr_a != chr) return -1;
      const uint32_t pa = e.pos_a <= e.pos2_a ? e.pos_a : e.pos2_a, pa2 = e.pos_a <= e.pos2_a ? e.pos2_a : e.pos_a;   // src/control_state.cpp:25-34
      const uint32_t pb = e.pos_b <= e.pos2_b ? e.pos_b : e.pos2_b, pb2 = e.pos_b <= e.pos2_b ? e.pos2_b : e.pos_b;
      if (i > b0 && pa < prev) return -1;
      prev = pa;
      if (pa < P.ws || pa >= P.we) continue;        // lower_bound at (chr, ws), upper_bound at (chr, we) with an empty second coordinate
      P.start.push_back(pgbd_start_of_window(pa));
      endw.push_back(pgbd_end_of_window(pa, pa2));
      PgBdSel s;
      s.chr2 = e.chr_b; s.rank2 = (uint32_t)chr_rank[e.chr_b]; s.pos2 = pb; s.s_start = pgbd_start_of_window(pb); s.s_end = pgbd_end_of_window(pb, pb2);
      s.pos1 = pa;
      P.sel.push_back(s);
   }
   const uint32_t m = (uint32_t)P.start.size();
   while (P.z < m && P.start[P.z] < P.ws) P.z++;
   P.drop.resize(m);
   uint32_t run = 0;
   for (uint32_t s = 0; s < m; s++) {
      const uint32_t add = s + P.z < m ? P.start[s + P.z] : PGBD_NEVER;
      uint32_t d = PGBD_NEVER;
      if (add != PGBD_NEVER && run != PGBD_NEVER) {
         const uint32_t g = (endw[s] > add ? endw[s] : add) + 1u;
         d = g > run ? g : run;
      }
      P.drop[s] = d;
      run = d;
   }
   return 0;
}
