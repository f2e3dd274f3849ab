/* oracle/boxsort_oracle.c -- TEST INFRASTRUCTURE ONLY (CPU restatement; the checker of pg_box_sort).
 *
 * Row N4 of SURVEY.md section 8(f): bubblesortReads and markDuplicates (src/reporter.cpp:932-972) with smaller()
 * (src/reporter.cpp:907-929) for reads of one chromosome, on plain arrays.  Pinned against the reference's own functions by
 * tests/test_boxsort.py (committed dumps tests/golden/boxsort_*.npz made through oracle/ref_hook.cpp, and live).
 */
#include <stdint.h>

typedef struct {      /* one read: the fields the two functions look at */
   uint32_t bp_left, bp_right, indel_size;
   int32_t left, right;
   uint32_t name_id;
   uint16_t nt_size;
   int16_t bp;
   uint8_t unique;
   uint8_t pad[3];
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
