// tests/emul/boxsort_emul.cpp -- TEST INFRASTRUCTURE ONLY: pindel_b200/csrc/pg_boxsort.cuh (what every thread of
// pg_box_sort_kernel runs) compiled as plain C++, one "thread" per box in a loop.  No GPU.
#include <stdint.h>
#include "../../pindel_b200/csrc/pg_boxsort.cuh"

// pg_box_sort_warp_kernel with the 32 lanes written out as arrays: the same packed keys, the same weak-prefix-minimum test,
// the same hand-over of values, chunk by chunk.
static void warp_formulation(const PgSortKey *bk, uint32_t n, uint32_t *ord, uint8_t *non)
{
   for (uint32_t k = 0; k < n; k++) ord[k] = k;
   for (uint32_t first = 0; first + 1 < n; first++) {
      uint32_t hand = ord[first];
      PgKey128 hand_key = pgs_pack(bk[hand]);
      for (uint32_t c0 = first + 1; c0 < n; c0 += 32) {
         uint32_t slot[32];
         PgKey128 k[32], pm[32], ex[32];
         bool valid[32], rec[32];
         for (int lane = 0; lane < 32; lane++) {
            const uint32_t j = c0 + (uint32_t)lane;
            valid[lane] = j < n;
            slot[lane] = valid[lane] ? ord[j] : 0u;
            k[lane].hi = k[lane].lo = ~0ull;
            if (valid[lane]) k[lane] = pgs_pack(bk[slot[lane]]);
         }
         for (int lane = 0; lane < 32; lane++) { pm[lane] = k[lane]; if (lane && pgs_lt(pm[lane - 1], pm[lane])) pm[lane] = pm[lane - 1]; }
         uint32_t R = 0;
         for (int lane = 0; lane < 32; lane++) {
            ex[lane] = (lane == 0 || pgs_lt(hand_key, pm[lane - 1])) ? hand_key : pm[lane - 1];
            rec[lane] = valid[lane] && pgs_le(k[lane], ex[lane]);
            if (rec[lane]) R |= 1u << lane;
         }
         if (R) {
            for (int lane = 0; lane < 32; lane++) {
               if (!rec[lane]) continue;
               const uint32_t below = R & ((1u << lane) - 1u);
               ord[c0 + lane] = below ? slot[31 - __builtin_clz(below)] : hand;
            }
            const int last = 31 - __builtin_clz(R);
            hand = slot[last];
            hand_key = k[last];
         }
      }
      ord[first] = hand;
   }
   for (uint32_t k = 0; k < n; k++) non[k] = bk[ord[k]].unique ? 0 : 1;
   for (uint32_t first = 0; first + 1 < n; first++) {
      if (non[first]) continue;
      const PgSortKey a = bk[ord[first]];
      for (uint32_t j = first + 1; j < n; j++) {
         const PgSortKey &c = bk[ord[j]];
         if (a.left == c.left && a.right == c.right && a.name_id == c.name_id) non[j] = 1;
      }
   }
}

extern "C" {
int boxemu_key_size(void) { return (int)sizeof(PgSortKey); }
void boxemu_sort_warp_formulation(uint32_t n_boxes, const uint32_t *box_off, const PgSortKey *keys, uint32_t *order, uint8_t *nonunique)
{
   for (uint32_t b = 0; b < n_boxes; b++) warp_formulation(keys + box_off[b], box_off[b + 1] - box_off[b], order + box_off[b], nonunique + box_off[b]);
}
void boxemu_sort(uint32_t n_boxes, const uint32_t *box_off, const PgSortKey *keys, uint32_t *order, uint8_t *nonunique)
{
   for (uint32_t b = 0; b < n_boxes; b++) {
      const uint32_t o = box_off[b], n = box_off[b + 1] - o;
      for (uint32_t k = 0; k < n; k++) order[o + k] = k;
      pgs_sort_and_mark(keys + o, n, order + o, nonunique + o);
   }
}
}
