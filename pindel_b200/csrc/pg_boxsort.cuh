// pindel_b200/csrc/pg_boxsort.cuh -- per-box read ordering and duplicate marking on the device (SURVEY.md section 8(f),
// row N4): bubblesortReads and markDuplicates (src/reporter.cpp:932-972), which the D, SI and TD reporters run on every box
// that holds enough reads (src/reporter.cpp:996-998, 1179-1180, 1425-1427).
//
// bubblesortReads is an exchange sort -- for First, for Second > First: swap unless Reads[First] < Reads[Second] -- and the
// order it leaves among reads with equal keys (every split read of one event has the same BPLeft/BPRight/IndelSize/
// NT_size, often the same BP) is neither the input order nor its reverse, and it decides which read markDuplicates keeps
// and the order of the reads under every event in the output files.  So the device does not "sort": it replays the
// exchange sort and the duplicate scan, one warp per box (the boxes are the parallel axis: tens of thousands per window).
// One pass of the exchange sort (fixed First) carries a value "in hand" along the box: position j takes it over and gives
// its own value to the hand exactly when key[j] <= min(key of the hand at the start, key[First+1 .. j-1]) -- the weak
// prefix minima -- so a pass is a rotation along the chain of those positions, found 32 positions at a time with a
// warp prefix-minimum over 128-bit packed keys.  markDuplicates is a plain lane-parallel scan per First.
// pgs_sort_and_mark is the literal form (the first device version ran it, one thread per box: a 481-read box cost 34 ms).
//
// The other reporters carry the same two loops inlined with their own rules (SortOutputDI, src/reporter.cpp:1728-1775;
// OutputSorter::DoSortAndOutputInversions, src/output_sorter.cpp:74-131): the exchange there is STRICT (swap only when the
// first read is greater, equal reads stay), the comparators differ, and the duplicate scan compares LeftMostPos / read end /
// anchor strand from EVERY first read, unique or not.  A strict pass is the same rotation along the STRICT prefix minima.
// After the two loops every reporter walks the surviving reads once and cuts them into events wherever its own key changes
// (the "event merge", e.g. src/reporter.cpp:1455-1487); that cut is produced here too, as one flag byte per read.
#pragma once
#include <stdint.h>

#ifdef __CUDACC__
#define PGS_HD __host__ __device__ __forceinline__
#else
#define PGS_HD inline
#endif

struct PgSortKey {          // pg_sort_key (include/pindel_pg.h): what smaller() and markDuplicates read of a SPLIT_READ
   uint32_t bp_left, bp_right, indel_size;
   int32_t left, right;
   uint32_t name_id;        // equal ids <=> equal Name strings (the caller numbers the names)
   uint16_t nt_size;
   int16_t bp;
   uint8_t unique;          // UniqueRead on entry
   uint8_t matched_d;       // MatchedD ('+' / '-'): DI / INV duplicate rule
   uint16_t read_length;    // getReadLength(): DI / INV duplicate rule; `left` carries LeftMostPos for those kinds
};

// which reporter's loops: the comparator and swap rule of the exchange sort, the duplicate rule, who counts as a "good" read
// and the key whose change starts a new event
enum PgBoxKind {
   PGS_D = 0,        // SortOutputD / SortAndOutputTandemDuplications: bubblesortReads + markDuplicates; events on (BPLeft, BPRight)
   PGS_SI = 1,       // SortOutputSI: the same two functions; events on (BPLeft, IndelSize)
   PGS_DI = 2,       // SortOutputDI: strict, (BPLeft, BPRight, NT_size, BP); events on (BPLeft, IndelSize, NT_size)
   PGS_INV = 3,      // DoSortAndOutputInversions(false): strict, (BPLeft + BPRight, IndelSize descending, BPLeft, BPRight, BP); every read is good; events on BPLeft + BPRight
   PGS_INV_NT = 4    // DoSortAndOutputInversions(true): NT_size compared before BP
};
#define PGS_EV_GOOD 1u       // the read at this position enters GoodIndels
#define PGS_EV_START 2u      // ... and starts a new event there

// smaller (src/reporter.cpp:907-929) for reads of one chromosome (FragName equal: the caller checks)
PGS_HD bool pgs_smaller(const PgSortKey &a, const PgSortKey &b)
{
   if (a.bp_left != b.bp_left) return a.bp_left < b.bp_left;
   if (a.bp_right != b.bp_right) return a.bp_right < b.bp_right;
   if (a.indel_size != b.indel_size) return a.indel_size < b.indel_size;
   if (a.nt_size != b.nt_size) return a.nt_size < b.nt_size;
   if (a.bp != b.bp) return a.bp < b.bp;
   return false;
}

// "first is greater than second" of the inlined sorts (src/reporter.cpp:1733-1759, src/output_sorter.cpp:81-118)
PGS_HD bool pgs_greater(const PgSortKey &a, const PgSortKey &b, int kind)
{
   if (kind == PGS_DI) {
      if (a.bp_left != b.bp_left) return a.bp_left > b.bp_left;
      if (a.bp_right != b.bp_right) return a.bp_right > b.bp_right;
      if (a.nt_size != b.nt_size) return a.nt_size > b.nt_size;
      return a.bp > b.bp;
   }
   const uint32_t sa = a.bp_left + a.bp_right, sb = b.bp_left + b.bp_right;      // unsigned arithmetic, as in the reference
   if (sa != sb) return sa > sb;
   if (a.indel_size != b.indel_size) return a.indel_size < b.indel_size;         // larger ones first
   if (a.bp_left != b.bp_left) return a.bp_left > b.bp_left;
   if (a.bp_right != b.bp_right) return a.bp_right > b.bp_right;
   if (kind == PGS_INV_NT && a.nt_size != b.nt_size) return a.nt_size > b.nt_size;
   return a.bp > b.bp;
}

// the duplicate test of the inlined scans (src/reporter.cpp:1764-1772, src/output_sorter.cpp:122-129); `left` = LeftMostPos
PGS_HD bool pgs_same_fragment(const PgSortKey &a, const PgSortKey &b, int kind)
{
   if (kind == PGS_DI && a.read_length != b.read_length) return false;
   const uint32_t la = (uint32_t)a.left, lb = (uint32_t)b.left;
   if (!(la == lb || la + (uint32_t)a.read_length == lb + (uint32_t)b.read_length)) return false;
   return a.matched_d == b.matched_d;
}

// does read b continue the event that read a belongs to (a: any earlier good read of the event -- the keys are equalities)
PGS_HD bool pgs_same_event(const PgSortKey &a, const PgSortKey &b, int kind)
{
   switch (kind) {
   case PGS_D: return a.bp_left == b.bp_left && a.bp_right == b.bp_right;                                   // src/reporter.cpp:1216-1217, 1456-1459
   case PGS_SI: return a.bp_left == b.bp_left && a.indel_size == b.indel_size;                              // :1023-1024
   case PGS_DI: return a.bp_left == b.bp_left && a.indel_size == b.indel_size && a.nt_size == b.nt_size;    // :1800-1803
   default: return a.bp_left + a.bp_right == b.bp_left + b.bp_right;                                        // src/output_sorter.cpp:155
   }
}

// order[0..n): member slots of the box, identity on entry.  nonunique[k] and event[k] (may be NULL) refer to the k-th read of
// the final order.
PGS_HD void pgs_sort_and_mark(const PgSortKey *keys, uint32_t n, uint32_t *order, uint8_t *nonunique, int kind = PGS_D, uint8_t *event = nullptr)
{
   const bool strict = kind >= PGS_DI;
   for (uint32_t first = 0; first + 1 < n; first++)
      for (uint32_t second = first + 1; second < n; second++)
         if (strict ? pgs_greater(keys[order[first]], keys[order[second]], kind) : !pgs_smaller(keys[order[first]], keys[order[second]])) {
            const uint32_t t = order[first];
            order[first] = order[second];
            order[second] = t;
         }
   for (uint32_t k = 0; k < n; k++) nonunique[k] = keys[order[k]].unique ? 0 : 1;
   for (uint32_t first = 0; first + 1 < n; first++) {
      if (!strict && nonunique[first]) continue;
      const PgSortKey &a = keys[order[first]];
      for (uint32_t second = first + 1; second < n; second++) {
         const PgSortKey &b = keys[order[second]];
         if (strict ? pgs_same_fragment(a, b, kind) : (a.left == b.left && a.right == b.right && a.name_id == b.name_id)) nonunique[second] = 1;
      }
   }
   if (!event) return;
   bool have = false;
   uint32_t prev = 0;
   for (uint32_t k = 0; k < n; k++) {
      const bool good = kind >= PGS_INV || !nonunique[k];
      event[k] = 0;
      if (!good) continue;
      event[k] = (uint8_t)(PGS_EV_GOOD | ((!have || !pgs_same_event(keys[order[prev]], keys[order[k]], kind)) ? PGS_EV_START : 0u));
      prev = k; have = true;
   }
}

// the five fields smaller() compares, as one 128-bit unsigned number (BP is a short: biased)
struct PgKey128 {
   uint64_t hi, lo;
};
PGS_HD PgKey128 pgs_pack(const PgSortKey &k, int kind = PGS_D)
{
   PgKey128 p;
   const uint64_t bp = (uint64_t)(uint16_t)((int32_t)k.bp + 32768);
   if (kind <= PGS_SI) {
      p.hi = (uint64_t)k.bp_left << 32 | k.bp_right;
      p.lo = (uint64_t)k.indel_size << 32 | (uint64_t)k.nt_size << 16 | bp;
   } else if (kind == PGS_DI) {
      p.hi = (uint64_t)k.bp_left << 32 | k.bp_right;
      p.lo = (uint64_t)k.nt_size << 16 | bp;
   } else {         // BPRight follows from the sum and BPLeft
      p.hi = (uint64_t)(k.bp_left + k.bp_right) << 32 | (uint32_t)~k.indel_size;
      p.lo = (uint64_t)k.bp_left << 32 | (kind == PGS_INV_NT ? (uint64_t)k.nt_size << 16 : 0ull) | bp;
   }
   return p;
}
PGS_HD bool pgs_le(const PgKey128 &a, const PgKey128 &b) { return a.hi < b.hi || (a.hi == b.hi && a.lo <= b.lo); }
PGS_HD bool pgs_lt(const PgKey128 &a, const PgKey128 &b) { return a.hi < b.hi || (a.hi == b.hi && a.lo < b.lo); }

#ifdef __CUDACC__
__device__ __forceinline__ PgKey128 pgs_shfl(const PgKey128 &k, int src)
{
   PgKey128 r;
   r.hi = __shfl_sync(0xFFFFFFFFu, k.hi, src);
   r.lo = __shfl_sync(0xFFFFFFFFu, k.lo, src);
   return r;
}
__device__ __forceinline__ PgKey128 pgs_shfl_up(const PgKey128 &k, int d)
{
   PgKey128 r;
   r.hi = __shfl_up_sync(0xFFFFFFFFu, k.hi, d);
   r.lo = __shfl_up_sync(0xFFFFFFFFu, k.lo, d);
   return r;
}

// one warp per box
__global__ void __launch_bounds__(256) pg_box_sort_warp_kernel(uint32_t n_boxes, const uint32_t *__restrict__ box_off, const PgSortKey *__restrict__ keys,
                                                               uint32_t *__restrict__ order, uint8_t *__restrict__ nonunique, int kind, uint8_t *__restrict__ event)
{
   const bool strict = kind >= PGS_DI;
   const uint32_t b = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
   const int lane = threadIdx.x & 31;
   if (b >= n_boxes) return;
   const uint32_t o = box_off[b], n = box_off[b + 1] - o;
   const PgSortKey *bk = keys + o;
   uint32_t *ord = order + o;
   uint8_t *non = nonunique + o;
   for (uint32_t k = lane; k < n; k += 32) ord[k] = k;
   __syncwarp();
   for (uint32_t first = 0; first + 1 < n; first++) {
      uint32_t hand = ord[first];
      PgKey128 hand_key = pgs_pack(bk[hand], kind);
      for (uint32_t c0 = first + 1; c0 < n; c0 += 32) {
         const uint32_t j = c0 + (uint32_t)lane;
         const bool valid = j < n;
         const uint32_t slot = valid ? ord[j] : 0u;
         PgKey128 k;
         k.hi = k.lo = ~0ull;
         if (valid) k = pgs_pack(bk[slot], kind);
         PgKey128 pm = k;                                   // inclusive prefix minimum over the lanes
         for (int d = 1; d < 32; d <<= 1) {
            const PgKey128 t = pgs_shfl_up(pm, d);
            if (lane >= d && pgs_lt(t, pm)) pm = t;
         }
         PgKey128 ex = pgs_shfl_up(pm, 1);                  // minimum of the hand and the lanes before this one
         if (lane == 0 || pgs_lt(hand_key, ex)) ex = hand_key;
         const bool rec = valid && (strict ? pgs_lt(k, ex) : pgs_le(k, ex));
         const uint32_t R = __ballot_sync(0xFFFFFFFFu, rec);
         if (R) {
            const uint32_t below = R & ((1u << lane) - 1u);
            const uint32_t from_lane = __shfl_sync(0xFFFFFFFFu, slot, below ? 31 - __clz(below) : 0);
            if (rec) ord[j] = below ? from_lane : hand;     // takes what was in hand: the previous taker's value
            const int last = 31 - __clz(R);
            hand = __shfl_sync(0xFFFFFFFFu, slot, last);
            hand_key = pgs_shfl(k, last);
         }
      }
      if (lane == 0) ord[first] = hand;
      __syncwarp();
   }
   for (uint32_t k = lane; k < n; k += 32) non[k] = bk[ord[k]].unique ? 0 : 1;
   __syncwarp();
   for (uint32_t first = 0; first + 1 < n; first++) {
      if (!strict && non[first]) continue;
      const PgSortKey a = bk[ord[first]];
      for (uint32_t j = first + 1 + (uint32_t)lane; j < n; j += 32) {
         const PgSortKey &c = bk[ord[j]];
         if (strict ? pgs_same_fragment(a, c, kind) : (a.left == c.left && a.right == c.right && a.name_id == c.name_id)) non[j] = 1;
      }
      __syncwarp();
   }
   if (!event) return;
   // the event cut: every good read against the good read before it
   uint8_t *ev = event + o;
   uint32_t prev = 0xFFFFFFFFu;                              // position of the last good read of the chunks before this one
   for (uint32_t c0 = 0; c0 < n; c0 += 32) {
      const uint32_t j = c0 + (uint32_t)lane;
      const bool good = j < n && (kind >= PGS_INV || !non[j]);
      const uint32_t G = __ballot_sync(0xFFFFFFFFu, good);
      if (good) {
         const uint32_t below = G & ((1u << lane) - 1u);
         const uint32_t p = below ? c0 + (uint32_t)(31 - __clz(below)) : prev;
         const bool start = p == 0xFFFFFFFFu || !pgs_same_event(bk[ord[p]], bk[ord[j]], kind);
         ev[j] = (uint8_t)(PGS_EV_GOOD | (start ? PGS_EV_START : 0u));
      } else if (j < n) {
         ev[j] = 0;
      }
      if (G) prev = c0 + (uint32_t)(31 - __clz(G));
   }
}

// box b: keys[box_off[b] .. box_off[b+1]); order / nonunique are the box's slices of the output arrays (literal form)
__global__ void pg_box_sort_kernel(uint32_t n_boxes, const uint32_t *__restrict__ box_off, const PgSortKey *__restrict__ keys,
                                   uint32_t *__restrict__ order, uint8_t *__restrict__ nonunique, int kind, uint8_t *__restrict__ event)
{
   const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
   if (b >= n_boxes) return;
   const uint32_t o = box_off[b], n = box_off[b + 1] - o;
   for (uint32_t k = 0; k < n; k++) order[o + k] = k;
   pgs_sort_and_mark(keys + o, n, order + o, nonunique + o, kind, event ? event + o : nullptr);
}
#endif
