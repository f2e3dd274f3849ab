// pindel_b200/csrc/pg_bd.cuh -- the BreakDancer mask and the cluster lookup (last part):
// BDData::loadRegion (src/bddata.cpp:850-946) sweeps every position of the search window (window +- 3000), keeps a
// contiguous range [S, E) of the events sorted by their first coordinate -- events enter at the back when the position
// reaches the start of their window, leave at the front once the position has passed the end of the FRONT event's window --
// and, whenever the range changes, builds the cluster of the range's second coordinates (createRegionCluster, :820-846);
// getCorrespondingSearchWindowCluster (:949-967) then maps a read's close-end position to the cluster of that position.
// The result is what SearchFarEnd scans first (src/pindel.cpp:1006-1020): the `regions` of the far-end kernel.
//
// Re-formulated without the per-position sweep.  With the selected events i = 0..m-1 in their sorted order, start_i / end_i the
// window of the first coordinate (position -+ 200), ws the first swept position and z = #{i : start_i < ws}:
//   * an event is appended only at the position that EQUALS its window start, and the range is contiguous, so the events
//     whose window starts left of ws are never appended themselves but the first z appends take THEIR slots:
//     E(q) = #{i : ws <= start_i <= q}, and slot s is filled at position add_s = start_{s+z};
//   * the front slot s is dropped at the first position p with p > end_s, p > add_s (it must have been filled before this
//     position's appends) and all slots before it dropped: drop_s = max(drop_{s-1}, max(end_s, add_s) + 1), a prefix
//     maximum; S(q) = #{s : drop_s <= q}.
// Both are binary searches per read over two arrays the host derives in O(m) when the region is loaded; the cluster of
// [S, E) -- sort by the second coordinate, merge neighbours whose windows touch -- is built per read on the fly (the ranges
// hold a handful of events).  The reference's order quirks come out exactly: ends that are not monotone in the sort order
// keep finished events in the range behind an unfinished front one; a merged window ends where its LAST member ends.
//
// __host__ __device__, so that the CPU suite runs the same code (tests/emul/bd_emul.cpp).
#pragma once
#include <stdint.h>

#include "pg_warp_types.h"

#ifdef __CUDACC__
#define PGBD_HD __host__ __device__ __forceinline__
#else
#define PGBD_HD inline
#endif

#define PGBD_SPAN 200u            // BreakDancerCoordinate::BREAKDANCER_WINDOWSPAN, src/control_state.h:46
#define PGBD_INSERT 1000u         // INSERT_SIZE of loadRegion, src/bddata.cpp:856
#define PGBD_NEVER 0xFFFFFFFFu

struct PgBdEvent {                // pg_bd_event (include/pindel_pg.h): one BreakDancerEvent, positions as the reference holds them (padded)
   int32_t chr_a;
   uint32_t pos_a, pos2_a;
   int32_t chr_b;
   uint32_t pos_b, pos2_b;
};

struct PgBdSel {                  // one selected event of the loaded region, what the lookup needs of it
   int32_t chr2;                  // second coordinate: chromosome index ...
   uint32_t rank2;                // ... rank of its name in string order (sortOnSecondBDCoordinate compares names)
   uint32_t pos2;                 // ... position (sort key)
   uint32_t s_start, s_end;       // ... startOfWindow / endOfWindow
   uint32_t pos1;                 // first coordinate's position (tie-break of the sort)
};

struct PgBdRegionView {           // the loaded region on the device / host
   const uint32_t *start;         // [m] startOfWindow of the first coordinate, non-decreasing
   const uint32_t *drop;          // [m] position at which slot s leaves the range (prefix maximum), PGBD_NEVER if never
   const PgBdSel *sel;            // [m]
   uint32_t m, z;
   uint32_t ws, we;               // swept positions [ws, we)
};

PGBD_HD uint32_t pgbd_start_of_window(uint32_t pos) { return pos >= PGBD_SPAN ? pos - PGBD_SPAN : 0u; }                 // src/control_state.cpp:71-83
PGBD_HD uint32_t pgbd_end_of_window(uint32_t pos, uint32_t pos2) { return ((pos2 > pos && pos2 > 0u) ? pos2 : pos) + PGBD_SPAN; }   // :90-98

// number of elements <= q in a non-decreasing array
PGBD_HD uint32_t pgbd_count_le(const uint32_t *a, uint32_t m, uint32_t q)
{
   uint32_t lo = 0, hi = m;
   while (lo < hi) {
      const uint32_t mid = (lo + hi) >> 1;
      if (a[mid] <= q) lo = mid + 1; else hi = mid;
   }
   return lo;
}

// The range [S, E) of slots that getCorrespondingSearchWindowCluster's cluster was built from for a close end at q; S == E: the
// empty cluster (also for positions outside the mask and for q == ws, src/bddata.cpp:952-963).  Positions at or beyond the last
// swept one read mask entries the reference never wrote: defined here as empty.
PGBD_HD void pgbd_range(const PgBdRegionView &R, uint32_t q, uint32_t &S, uint32_t &E)
{
   S = E = 0;
   if (q <= R.ws || q >= R.we) return;
   const uint32_t le = pgbd_count_le(R.start, R.m, q);
   E = le > R.z ? le - R.z : 0u;
   S = pgbd_count_le(R.drop, R.m, q);
   if (S > E) S = E;
}

PGBD_HD bool pgbd_less(const PgBdSel &a, uint32_t ia, const PgBdSel &b, uint32_t ib)      // sortOnSecondBDCoordinate, then the slot (equal elements)
{
   if (a.rank2 != b.rank2) return a.rank2 < b.rank2;
   if (a.pos2 != b.pos2) return a.pos2 < b.pos2;
   if (a.pos1 != b.pos1) return a.pos1 < b.pos1;
   return ia < ib;
}

// createRegionCluster over the slots [S, E): windows written to out (when not NULL, at most cap); returns their number.
// Selection in sorted order: r^2 comparisons for a range of r events (a handful).
PGBD_HD uint32_t pgbd_cluster(const PgBdRegionView &R, uint32_t S, uint32_t E, PgRegion *out, uint32_t cap)
{
   uint32_t n = 0;
   bool have_prev = false, open = false;
   uint32_t prev = 0;
   PgRegion cur;
   cur.chr = 0; cur.start = 0; cur.end = 0;
   for (uint32_t k = S; k < E; k++) {
      // the smallest slot that is greater than the previously taken one
      uint32_t best = PGBD_NEVER;
      for (uint32_t j = S; j < E; j++) {
         if (have_prev && !pgbd_less(R.sel[prev], prev, R.sel[j], j)) continue;
         if (best == PGBD_NEVER || pgbd_less(R.sel[j], j, R.sel[best], best)) best = j;
      }
      const PgBdSel e = R.sel[best];
      // regionsOverlap (src/bddata.cpp:814-818) against the PREVIOUS event of the sorted order, not against the window so far
      if (open && e.rank2 == R.sel[prev].rank2 && e.s_start <= R.sel[prev].s_end + 1u) {
         cur.end = e.s_end;
      } else {
         if (open) { if (out && n < cap) out[n] = cur; n++; }
         cur.chr = e.chr2; cur.start = e.s_start; cur.end = e.s_end;
         open = true;
      }
      prev = best; have_prev = true;
   }
   if (open) { if (out && n < cap) out[n] = cur; n++; }
   return n;
}

#ifdef __CUDACC__
// One thread per read: the number of windows of its cluster (capped at 255; the count kernel's byte array feeds pg_scan_*),
// then the windows themselves at the scanned offsets.
__global__ void pg_bd_count_kernel(PgBdRegionView R, uint32_t n, const PgCloseInfo *__restrict__ ci, const uint32_t *__restrict__ pos, uint8_t *__restrict__ counts,
                                   uint32_t *__restrict__ overflow)
{
   const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
   if (i >= n) return;
   uint32_t c = 0;
   const bool has = ci ? ci[i].max_len != 0 : true;
   if (has) {
      uint32_t S, E;
      pgbd_range(R, ci ? ci[i].last_abs_loc : pos[i], S, E);
      if (E > S) c = pgbd_cluster(R, S, E, nullptr, 0);
   }
   if (c > 255u) { atomicAdd(overflow, 1u); c = 255u; }
   counts[i] = (uint8_t)c;
}

__global__ void pg_bd_fill_kernel(PgBdRegionView R, uint32_t n, const PgCloseInfo *__restrict__ ci, const uint32_t *__restrict__ pos, const uint32_t *__restrict__ offsets,
                                  const uint8_t *__restrict__ counts, PgRegion *__restrict__ regions, uint32_t *__restrict__ region_off, const uint32_t *__restrict__ total)
{
   const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
   if (i > n) return;
   if (i == n) { region_off[n] = *total; return; }
   region_off[i] = offsets[i];
   if (counts[i] == 0) return;
   uint32_t S, E;
   pgbd_range(R, ci ? ci[i].last_abs_loc : pos[i], S, E);
   pgbd_cluster(R, S, E, regions + offsets[i], counts[i]);
}
#endif
