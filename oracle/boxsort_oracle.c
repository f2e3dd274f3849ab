/* oracle/boxsort_oracle.c -- TEST INFRASTRUCTURE ONLY (CPU restatement; the checker of pg_box_sort).
 *
 * Row N4 of SURVEY.md section 8(f): bubblesortReads and markDuplicates (src/reporter.cpp:932-972) with smaller()
 * (src/reporter.cpp:907-929) for reads of one chromosome, on plain arrays.  Pinned against the reference's own functions by
 * tests/test_boxsort.py (committed dumps tests/golden/boxsort_*.npz made through oracle/ref_hook.cpp, and live).
 *
 * bso_box_kind restates the same two loops as the DI and inversion reporters carry them inlined (SortOutputDI,
 * src/reporter.cpp:1728-1775; OutputSorter::DoSortAndOutputInversions, src/output_sorter.cpp:74-131) and the cut of the
 * surviving reads into events that every reporter makes next.  Those loops are not separate functions in the reference, so
 * they cannot be driven alone: they are pinned at the file level -- the pure reference binary beside the drop-in on inputs
 * dense in DI and inversion events (tests/dropin_scenarios.py, dense_events_live), where the drop-in's boxes go through the
 * device program that tests/test_boxsort.py holds equal to this restatement.
 */
#include <stdint.h>

typedef struct {      /* one read: the fields the two functions look at */
   uint32_t bp_left, bp_right, indel_size;
   int32_t left, right;
   uint32_t name_id;
   uint16_t nt_size;
   int16_t bp;
   uint8_t unique;
   uint8_t matched_d;     /* MatchedD */
   uint16_t read_length;  /* getReadLength(); for the DI / inversion loops `left` holds LeftMostPos */
} bso_read;

static int smaller(const bso_read *a, const bso_read *b)   /* src/reporter.cpp:907-929, FragName equal */
{
   if (a->bp_left != b->bp_left) return a->bp_left < b->bp_left;
   if (a->bp_right != b->bp_right) return a->bp_right < b->bp_right;
   if (a->indel_size != b->indel_size) return a->indel_size < b->indel_size;
   if (a->nt_size != b->nt_size) return a->nt_size < b->nt_size;
   if (a->bp != b->bp) return a->bp < b->bp;
   return 0;
}

/* reads: all reads; idx[0..n): VariantIndices of one box, reordered in place (bubblesortReads); unique[r] = UniqueRead of read
 * r, updated in place (markDuplicates). */
void bso_box(const bso_read *reads, uint8_t *unique, uint32_t *idx, uint32_t n)
{
   if (n == 0) return;
   for (uint32_t first = 0; first < n - 1; first++)
      for (uint32_t second = first + 1; second < n; second++)
         if (!smaller(&reads[idx[first]], &reads[idx[second]])) {
            uint32_t t = idx[first]; idx[first] = idx[second]; idx[second] = t;
         }
   for (uint32_t first = 0; first < n - 1; first++) {
      const bso_read *a = &reads[idx[first]];
      if (!unique[idx[first]]) continue;
      for (uint32_t second = first + 1; second < n; second++) {
         const bso_read *b = &reads[idx[second]];
         if (a->left == b->left && a->right == b->right && a->name_id == b->name_id) unique[idx[second]] = 0;
      }
   }
}

/* kind: 0 D / TD, 1 SI (bubblesortReads + markDuplicates as above), 2 DI, 3 INV, 4 INV_NT */
static int first_is_greater(const bso_read *a, const bso_read *b, int kind)
{
   if (kind == 2) {                                            /* src/reporter.cpp:1733-1759 */
      if (a->bp_left < b->bp_left) return 0;
      if (a->bp_left > b->bp_left) return 1;
      if (a->bp_right < b->bp_right) return 0;
      if (a->bp_right > b->bp_right) return 1;
      if (a->nt_size < b->nt_size) return 0;
      if (a->nt_size > b->nt_size) return 1;
      return a->bp > b->bp;
   }
   {                                                           /* src/output_sorter.cpp:81-118 */
      const uint32_t sa = a->bp_left + a->bp_right, sb = b->bp_left + b->bp_right;
      if (sa < sb) return 0;
      if (sa > sb) return 1;
      if (a->indel_size > b->indel_size) return 0;             /* IndelSize: larger ones first */
      if (a->indel_size < b->indel_size) return 1;
      if (a->bp_left < b->bp_left) return 0;
      if (a->bp_left > b->bp_left) return 1;
      if (a->bp_right < b->bp_right) return 0;
      if (a->bp_right > b->bp_right) return 1;
      if (kind == 4) {
         if (a->nt_size < b->nt_size) return 0;
         if (a->nt_size > b->nt_size) return 1;
      }
      return a->bp > b->bp;
   }
}

/* event[k]: bit 0 = the k-th read of the final order enters GoodIndels, bit 1 = it starts a new event */
void bso_box_kind(const bso_read *reads, uint8_t *unique, uint32_t *idx, uint32_t n, int kind, uint8_t *event)
{
   if (n == 0) return;
   if (kind <= 1) {
      bso_box(reads, unique, idx, n);
   } else {
      for (uint32_t first = 0; first < n - 1; first++)
         for (uint32_t second = first + 1; second < n; second++)
            if (first_is_greater(&reads[idx[first]], &reads[idx[second]], kind)) {
               uint32_t t = idx[first]; idx[first] = idx[second]; idx[second] = t;
            }
      for (uint32_t first = 0; first < n - 1; first++)
         for (uint32_t second = first + 1; second < n; second++) {
            const bso_read *a = &reads[idx[first]], *b = &reads[idx[second]];
            if (kind == 2 && a->read_length != b->read_length) continue;                                    /* src/reporter.cpp:1766 */
            if ((uint32_t)a->left == (uint32_t)b->left || (uint32_t)a->left + a->read_length == (uint32_t)b->left + b->read_length)
               if (a->matched_d == b->matched_d) unique[idx[second]] = 0;
         }
   }
   if (!event) return;
   /* GoodIndels and the event walk: D src/reporter.cpp:1436-1487, TD :1190-1236, SI :1002-1034, DI :1777-1813, INV src/output_sorter.cpp:134-205 */
   int have = 0;
   const bso_read *ev = 0;                                      /* the read OneIndelEvent was initialised from */
   for (uint32_t k = 0; k < n; k++) {
      const bso_read *r = &reads[idx[k]];
      event[k] = 0;
      if (kind < 3 && !unique[idx[k]]) continue;
      int same = 0;
      if (have) {
         switch (kind) {
         case 0: same = r->bp_left == ev->bp_left && r->bp_right == ev->bp_right; break;
         case 1: same = r->bp_left == ev->bp_left && r->indel_size == ev->indel_size; break;
         case 2: same = r->bp_left == ev->bp_left && r->indel_size == ev->indel_size && r->nt_size == ev->nt_size; break;
         default: same = r->bp_left + r->bp_right == ev->bp_left + ev->bp_right; break;
         }
      }
      event[k] = (uint8_t)(1 | (same ? 0 : 2));
      if (!same) ev = r;
      have = 1;
   }
}
