// pindel_b200/csrc/pg_rp_host.h -- host part of pg_rp_modify: the O(n log n) split of a window's discordant pairs into
// classes that cannot influence each other (the device then runs every class as an independent recurrence).  Kept apart
// from pindel_pg.cu so that the CPU suite exercises the same function (tests/emul/rp_emul.cpp).
#pragma once
#include <stdint.h>
#include <algorithm>
#include <utility>
#include <vector>
#include "pg_rp.cuh"

// Classes of rows that can interact (see pg_rp.cuh): equal strands, and connected through overlapping initial A intervals
// or overlapping initial B intervals.  Rows whose A or B interval is wider than 1000 never pass RecipicalOverlap, neither
// as first nor as second: they stay alone.  -> rows grouped by class (ascending inside a class), classes of one row left out;
// classes of more than warp_max rows go to the second list.
static inline void pg_rp_classes(int n, const PgRp *in, uint32_t warp_max, std::vector<uint32_t> &rows_small, std::vector<uint32_t> &off_small,
                                 std::vector<uint32_t> &rows_big, std::vector<uint32_t> &off_big)
{
   std::vector<PgRpIv> iv((size_t)n);
   std::vector<uint32_t> parent((size_t)n), live;
   live.reserve((size_t)n);
   for (int i = 0; i < n; i++) {
      iv[i] = pg_rp_init(in[i]);
      parent[i] = (uint32_t)i;
      if (!(pg_rp_iabs((int)(iv[i].a - iv[i].a1)) > 1000 || pg_rp_iabs((int)(iv[i].b - iv[i].b1)) > 1000)) live.push_back((uint32_t)i);
   }
   auto find = [&](uint32_t x) {
      while (parent[x] != x) { parent[x] = parent[parent[x]]; x = parent[x]; }
      return x;
   };
   auto unite = [&](uint32_t x, uint32_t y) {
      x = find(x); y = find(y);
      if (x != y) parent[x > y ? x : y] = x > y ? y : x;
   };
   std::vector<std::pair<uint64_t, uint32_t>> keyed(live.size());
   for (int side = 0; side < 2; side++) {
      // sweep over (strands, interval start): a row joins the running component while its start lies before the component's
      // end.  The sort runs on packed (strands << 32 | start, row) pairs: plain integer comparisons.
      for (size_t k = 0; k < live.size(); k++) {
         const uint32_t x = live[k];
         keyed[k].first = (uint64_t)((uint32_t)in[x].da << 8 | in[x].db) << 32 | (side ? iv[x].b : iv[x].a);
         keyed[k].second = x;
      }
      std::sort(keyed.begin(), keyed.end());
      uint32_t run_key = 0xFFFFFFFFu, run_end = 0, run_row = 0;
      for (size_t i = 0; i < keyed.size(); i++) {
         const uint32_t x = keyed[i].second;
         const uint32_t k = (uint32_t)(keyed[i].first >> 32), s0 = (uint32_t)keyed[i].first, s1 = side ? iv[x].b1 : iv[x].a1;
         if (k == run_key && s0 < run_end) {
            unite(x, run_row);
            if (s1 > run_end) run_end = s1;
         } else {
            run_key = k;
            run_end = s1;
            run_row = x;
         }
      }
   }
   // class lists: counting sort by root keeps the rows of a class ascending
   std::vector<uint32_t> count((size_t)n, 0), start((size_t)n, 0);
   for (uint32_t x : live) count[find(x)]++;
   rows_small.clear(); rows_big.clear();
   off_small.assign(1, 0); off_big.assign(1, 0);
   std::vector<uint32_t> slot((size_t)n, 0xFFFFFFFFu);
   for (int i = 0; i < n; i++) {      // roots in ascending order; a root is the smallest row of its class
      if (count[i] < 2) continue;
      std::vector<uint32_t> &rows = count[i] > warp_max ? rows_big : rows_small;
      std::vector<uint32_t> &off = count[i] > warp_max ? off_big : off_small;
      slot[i] = (uint32_t)rows.size();
      rows.resize(rows.size() + count[i]);
      off.push_back((uint32_t)rows.size());
   }
   for (uint32_t x : live) {      // live is ascending: it was filled in row order and never reordered
      const uint32_t r = find(x);
      if (count[r] < 2) continue;
      (count[r] > warp_max ? rows_big : rows_small)[slot[r]++] = x;
   }
}

// The same split for the inter-chromosomal rows (see pg_rp.cuh): flip[i] = 1 when row i is seen from its other end in the
// canonical orientation; classes = equal (chromosome pair, canonical strands), connected through overlapping canonical A
// or B intervals.  Every row is live here (no RecipicalOverlap that would make wide rows inert).  One list of classes.
static inline void pg_rp_classes_interchr(int n, const PgRp *in, const int32_t *chr_a, const int32_t *chr_b, std::vector<uint8_t> &flip,
                                          std::vector<uint32_t> &rows, std::vector<uint32_t> &off)
{
   struct Item {
      uint64_t key;
      uint32_t start, row;
      bool operator<(const Item &o) const { return key != o.key ? key < o.key : start != o.start ? start < o.start : row < o.row; }
   };
   std::vector<PgRpIv> iv((size_t)n);
   std::vector<uint64_t> key((size_t)n);
   std::vector<uint32_t> parent((size_t)n);
   flip.assign((size_t)n, 0);
   for (int i = 0; i < n; i++) {
      iv[i] = pg_rp_init(in[i]);
      flip[i] = chr_a[i] > chr_b[i];
      if (flip[i]) iv[i] = pg_rp_swap_sides(iv[i]);
      const uint32_t lo = (uint32_t)(flip[i] ? chr_b[i] : chr_a[i]), hi = (uint32_t)(flip[i] ? chr_a[i] : chr_b[i]);
      const uint32_t da = flip[i] ? in[i].db : in[i].da, db = flip[i] ? in[i].da : in[i].db;
      key[i] = (uint64_t)(lo & 0xFFFFFFu) << 40 | (uint64_t)(hi & 0xFFFFFFu) << 16 | (uint64_t)da << 8 | db;
      parent[i] = (uint32_t)i;
   }
   auto find = [&](uint32_t x) {
      while (parent[x] != x) { parent[x] = parent[parent[x]]; x = parent[x]; }
      return x;
   };
   auto unite = [&](uint32_t x, uint32_t y) {
      x = find(x); y = find(y);
      if (x != y) parent[x > y ? x : y] = x > y ? y : x;
   };
   std::vector<Item> items((size_t)n);
   for (int side = 0; side < 2; side++) {
      for (int i = 0; i < n; i++) { items[i].key = key[i]; items[i].start = side ? iv[i].b : iv[i].a; items[i].row = (uint32_t)i; }
      std::sort(items.begin(), items.end());
      uint64_t run_key = ~0ull;
      uint32_t run_end = 0, run_row = 0;
      for (int i = 0; i < n; i++) {
         const uint32_t x = items[i].row, s1 = side ? iv[x].b1 : iv[x].a1;
         if (items[i].key == run_key && items[i].start < run_end) {
            unite(x, run_row);
            if (s1 > run_end) run_end = s1;
         } else {
            run_key = items[i].key;
            run_end = s1;
            run_row = x;
         }
      }
   }
   std::vector<uint32_t> count((size_t)n, 0), slot((size_t)n, 0xFFFFFFFFu);
   for (int i = 0; i < n; i++) count[find((uint32_t)i)]++;
   rows.clear();
   off.assign(1, 0);
   for (int i = 0; i < n; i++) {
      if (count[i] < 2) continue;
      slot[i] = (uint32_t)rows.size();
      rows.resize(rows.size() + count[i]);
      off.push_back((uint32_t)rows.size());
   }
   for (int i = 0; i < n; i++) {
      const uint32_t r = find((uint32_t)i);
      if (count[r] >= 2) rows[slot[r]++] = (uint32_t)i;
   }
}
