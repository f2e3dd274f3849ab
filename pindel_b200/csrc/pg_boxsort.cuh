// pindel_b200/csrc/pg_boxsort.cuh -- per-box read ordering and duplicate marking on the device (SURVEY.md section 8(f),
// row N4): bubblesortReads and markDuplicates (src/reporter.cpp:932-972), which the D, SI and TD reporters run on every box
// that holds enough reads (src/reporter.cpp:996-998, 1179-1180, 1425-1427).
//
// bubblesortReads is an exchange sort -- for First, for Second > First: swap unless Reads[First] < Reads[Second] -- and the
// order it leaves among reads with equal keys (every split read of one event has the same BPLeft/BPRight/IndelSize/
// NT_size, often the same BP) is neither the input order nor its reverse, and it decides which read markDuplicates keeps
// and the order of the reads under every event in the output files.  So the device does not "sort": it replays the
// exchange sort and the duplicate scan literally, one thread per box -- the boxes are the parallel axis (tens of
// thousands per window), the reads of a box a short serial loop.
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
   uint8_t pad[3];
};

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

// order[0..n): member slots of the box, identity on entry.  nonunique[k] refers to the k-th read of the final order.
PGS_HD void pgs_sort_and_mark(const PgSortKey *keys, uint32_t n, uint32_t *order, uint8_t *nonunique)
{
   for (uint32_t first = 0; first + 1 < n; first++)
      for (uint32_t second = first + 1; second < n; second++)
         if (!pgs_smaller(keys[order[first]], keys[order[second]])) {
            const uint32_t t = order[first];
            order[first] = order[second];
            order[second] = t;
         }
   for (uint32_t k = 0; k < n; k++) nonunique[k] = keys[order[k]].unique ? 0 : 1;
   for (uint32_t first = 0; first + 1 < n; first++) {
      if (nonunique[first]) continue;
      const PgSortKey &a = keys[order[first]];
      for (uint32_t second = first + 1; second < n; second++) {
         const PgSortKey &b = keys[order[second]];
         if (a.left == b.left && a.right == b.right && a.name_id == b.name_id) nonunique[second] = 1;
      }
   }
}

#ifdef __CUDACC__
// box b: keys[box_off[b] .. box_off[b+1]); order / nonunique are the box's slices of the output arrays
__global__ void pg_box_sort_kernel(uint32_t n_boxes, const uint32_t *__restrict__ box_off, const PgSortKey *__restrict__ keys,
                                   uint32_t *__restrict__ order, uint8_t *__restrict__ nonunique)
{
   const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
   if (b >= n_boxes) return;
   const uint32_t o = box_off[b], n = box_off[b + 1] - o;
   for (uint32_t k = 0; k < n; k++) order[o + k] = k;
   pgs_sort_and_mark(keys + o, n, order + o, nonunique + o);
}
#endif
