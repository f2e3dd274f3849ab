// tests/emul/boxsort_emul.cpp -- TEST INFRASTRUCTURE ONLY: pindel_b200/csrc/pg_boxsort.cuh (what every thread of
// pg_box_sort_kernel runs) compiled as plain C++, one "thread" per box in a loop.  No GPU.
#include <stdint.h>
#include "../../pindel_b200/csrc/pg_boxsort.cuh"

// pg_box_sort_warp_kernel with the 32 lanes written out as arrays: the same packed keys, the same weak-prefix-minimum test,
// the same hand-over of values, chunk by chunk.
static void warp_formulation(const PgSortKey *bk, uint32_t n, uint32_t *ord, uint8_t *non, int kind, uint8_t *ev)
{
   const bool strict = kind >= PGS_DI;
   for (uint32_t k = 0; k < n; k++) ord[k] = k;
   for (uint32_t first = 0; first + 1 < n; first++) {
      uint32_t hand = ord[first];
      PgKey128 hand_key = pgs_pack(bk[hand], kind);
      for (uint32_t c0 = first + 1; c0 < n; c0 += 32) {
         uint32_t slot[32];
         PgKey128 k[32], pm[32], ex[32];
         bool valid[32], rec[32];
         for (int lane = 0; lane < 32; lane++) {
            const uint32_t j = c0 + (uint32_t)lane;
            valid[lane] = j < n;
            slot[lane] = valid[lane] ? ord[j] : 0u;
            k[lane].hi = k[lane].lo = ~0ull;
            if (valid[lane]) k[lane] = pgs_pack(bk[slot[lane]], kind);
         }
         for (int lane = 0; lane < 32; lane++) { pm[lane] = k[lane]; if (lane && pgs_lt(pm[lane - 1], pm[lane])) pm[lane] = pm[lane - 1]; }
         uint32_t R = 0;
         for (int lane = 0; lane < 32; lane++) {
            ex[lane] = (lane == 0 || pgs_lt(hand_key, pm[lane - 1])) ? hand_key : pm[lane - 1];
            rec[lane] = valid[lane] && (strict ? pgs_lt(k[lane], ex[lane]) : pgs_le(k[lane], ex[lane]));
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
      if (!strict && non[first]) continue;
      const PgSortKey a = bk[ord[first]];
      for (uint32_t j = first + 1; j < n; j++) {
         const PgSortKey &c = bk[ord[j]];
         if (strict ? pgs_same_fragment(a, c, kind) : (a.left == c.left && a.right == c.right && a.name_id == c.name_id)) non[j] = 1;
      }
   }
   if (!ev) return;
   // the event cut, chunk by chunk with the ballot written out
   uint32_t prev = 0xFFFFFFFFu;
   for (uint32_t c0 = 0; c0 < n; c0 += 32) {
      uint32_t G = 0;
      for (int lane = 0; lane < 32; lane++) {
         const uint32_t j = c0 + (uint32_t)lane;
         if (j < n && (kind >= PGS_INV || !non[j])) G |= 1u << lane;
      }
      for (int lane = 0; lane < 32; lane++) {
         const uint32_t j = c0 + (uint32_t)lane;
         if (j >= n) continue;
         if (!((G >> lane) & 1u)) { ev[j] = 0; continue; }
         const uint32_t below = G & ((1u << lane) - 1u);
         const uint32_t p = below ? c0 + (uint32_t)(31 - __builtin_clz(below)) : prev;
         const bool start = p == 0xFFFFFFFFu || !pgs_same_event(bk[ord[p]], bk[ord[j]], kind);
         ev[j] = (uint8_t)(PGS_EV_GOOD | (start ? PGS_EV_START : 0u));
      }
      if (G) prev = c0 + (uint32_t)(31 - __builtin_clz(G));
   }
}

extern "C" {
int boxemu_key_size(void) { return (int)sizeof(PgSortKey); }
void boxemu_sort_warp_formulation_kind(int kind, uint32_t n_boxes, const uint32_t *box_off, const PgSortKey *keys, uint32_t *order, uint8_t *nonunique, uint8_t *event)
{
   for (uint32_t b = 0; b < n_boxes; b++)
      warp_formulation(keys + box_off[b], box_off[b + 1] - box_off[b], order + box_off[b], nonunique + box_off[b], kind, event ? event + box_off[b] : nullptr);
}
void boxemu_sort_warp_formulation(uint32_t n_boxes, const uint32_t *box_off, const PgSortKey *keys, uint32_t *order, uint8_t *nonunique)
{
   boxemu_sort_warp_formulation_kind(PGS_D, n_boxes, box_off, keys, order, nonunique, nullptr);
}
void boxemu_sort_kind(int kind, uint32_t n_boxes, const uint32_t *box_off, const PgSortKey *keys, uint32_t *order, uint8_t *nonunique, uint8_t *event)
{
   for (uint32_t b = 0; b < n_boxes; b++) {
      const uint32_t o = box_off[b], n = box_off[b + 1] - o;
      for (uint32_t k = 0; k < n; k++) order[o + k] = k;
      pgs_sort_and_mark(keys + o, n, order + o, nonunique + o, kind, event ? event + o : nullptr);
   }
}
void boxemu_sort(uint32_t n_boxes, const uint32_t *box_off, const PgSortKey *keys, uint32_t *order, uint8_t *nonunique)
{
   boxemu_sort_kind(PGS_D, n_boxes, box_off, keys, order, nonunique, nullptr);
}
}
