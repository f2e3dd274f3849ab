// tests/emul/boxsort_emul.cpp -- TEST INFRASTRUCTURE ONLY: pindel_b200/csrc/pg_boxsort.cuh (what every thread of
// pg_box_sort_kernel runs) compiled as plain C++, one "thread" per box in a loop.  No GPU.
#include <stdint.h>
#include "../../pindel_b200/csrc/pg_boxsort.cuh"

extern "C" {
int boxemu_key_size(void) { return (int)sizeof(PgSortKey); }
void boxemu_sort(uint32_t n_boxes, const uint32_t *box_off, const PgSortKey *keys, uint32_t *order, uint8_t *nonunique)
{
   for (uint32_t b = 0; b < n_boxes; b++) {
      const uint32_t o = box_off[b], n = box_off[b + 1] - o;
      for (uint32_t k = 0; k < n; k++) order[o + k] = k;
      pgs_sort_and_mark(keys + o, n, order + o, nonunique + o);
   }
}
}
