// tests/emul/sr_emul.cpp -- TEST INFRASTRUCTURE ONLY: the reader's split-read filter of pindel_b200/csrc/pg_sr.cuh compiled as
// plain C++ and driven step by step like pg_bam_sr_candidates drives the kernels: stop index, per-record analysis, names into
// the open-addressing table (in a shuffled order -- on the device the order of the atomics is arbitrary, the result must not
// depend on it), partner and calls per record, prefix counts, scatter.  No GPU.
#include <stdint.h>
#include <string.h>
#include <vector>
#include "../../pindel_b200/csrc/pg_sr.cuh"

extern "C" {

int sremu_sizes(int which) { return which == 0 ? (int)sizeof(PgSrQuery) : which == 1 ? (int)sizeof(PgSrCand) : which == 2 ? (int)sizeof(PgSrRef) : (int)sizeof(PgSrInfo); }

// counts: [0] candidates, [1] reference-supporting reads, [2] stride of seqs, [3] records delivered, [4] error bits.
// cands == NULL: counts only.  table_bits: log2 of the table size (0 = the product's choice); a small table forces collisions
// of slots, hash_mask (0 = none) is and-ed onto every key to force collisions of whole keys.
int sremu_candidates(const uint8_t *d, const PgBamRec *recs, uint32_t n_recs, const PgSrQuery *q, uint32_t shuffle_seed, uint32_t table_bits,
                     uint64_t hash_mask, PgSrCand *cands, uint8_t *seqs, PgSrRef *refs, uint64_t *counts)
{
   const PgRpQuery itr = pgsr_iterator(*q);
   uint32_t stop = n_recs;
   for (uint32_t i = 0; i < n_recs; i++)
      if (pgb_itr_stops(recs[i], itr)) { stop = i; break; }
   std::vector<PgSrInfo> info(n_recs);
   uint32_t delivered = 0, stride = 0;
   for (uint32_t i = 0; i < n_recs; i++) {              // pg_sr_info_kernel
      memset(&info[i], 0, sizeof(PgSrInfo));
      if (i < stop && pgb_itr_delivers(recs[i], itr)) {
         pgsr_analyse(d, recs[i], *q, info[i]);
         if (hash_mask && !(info[i].bits & PGSR_MALFORMED)) { info[i].key &= hash_mask; if (!info[i].key) info[i].key = 1; }
         delivered++;
         if (info[i].length > stride) stride = info[i].length;
      }
   }
   uint32_t bits = table_bits;
   if (bits == 0) { bits = 4; while ((1ull << bits) < 2ull * (stop ? stop : 1)) bits++; }
   const uint32_t mask = (uint32_t)((1ull << bits) - 1);
   if ((uint64_t)delivered > (uint64_t)mask) return -1;  // the table must keep a free slot
   std::vector<uint64_t> keys((size_t)mask + 1, 0);
   std::vector<uint32_t> head((size_t)mask + 1, PGSR_NONE), slot(n_recs, 0), next(n_recs, PGSR_NONE), order(n_recs);
   for (uint32_t i = 0; i < n_recs; i++) order[i] = i;
   uint64_t rng = 0x9E3779B97F4A7C15ull * (shuffle_seed + 1);
   if (shuffle_seed)
      for (uint32_t i = n_recs; i > 1; i--) {
         rng ^= rng << 13; rng ^= rng >> 7; rng ^= rng << 17;
         const uint32_t j = (uint32_t)(rng % i);
         const uint32_t t = order[i - 1]; order[i - 1] = order[j]; order[j] = t;
      }
   for (uint32_t k = 0; k < n_recs; k++) {              // pg_sr_insert_kernel
      const uint32_t i = order[k];
      if ((info[i].bits & (PGSR_DELIVERED | PGSR_MALFORMED)) != PGSR_DELIVERED) continue;
      uint32_t s = (uint32_t)(info[i].key >> 20) & mask;
      while (keys[s] != 0 && keys[s] != info[i].key) s = (s + 1u) & mask;
      keys[s] = info[i].key;
      slot[i] = s;
      next[i] = head[s];
      head[s] = i;
   }
   std::vector<uint32_t> partner(n_recs, PGSR_NONE), off_sr(n_recs, 0), off_ref(n_recs, 0);
   std::vector<uint8_t> emit(n_recs, 0);
   uint32_t err = 0, n_sr = 0, n_ref = 0;
   for (uint32_t i = 0; i < n_recs; i++) {              // pg_sr_pair_kernel + the two scans
      if (info[i].bits & PGSR_DELIVERED) {
         if (!(info[i].bits & PGSR_MALFORMED)) partner[i] = pgsr_partner(d, recs, next.data(), head[slot[i]], i);
         const uint32_t m = pgsr_emit_mask(info[i], partner[i] == PGSR_NONE ? (const PgSrInfo *)0 : &info[partner[i]], *q);
         err |= m & (PGSR_ERR_INSERT | PGSR_ERR_MALFORMED);
         emit[i] = (uint8_t)m;
      }
      off_sr[i] = n_sr; off_ref[i] = n_ref;
      n_sr += pgsr_count_sr(emit[i]); n_ref += pgsr_count_ref(emit[i]);
   }
   counts[0] = n_sr; counts[1] = n_ref; counts[2] = stride; counts[3] = delivered; counts[4] = err;
   if (!cands) return 0;
   for (uint32_t i = 0; i < n_recs; i++)                // pg_sr_scatter_kernel
      if (emit[i] & 0x3Fu) pgsr_scatter(d, recs, info.data(), *q, i, partner[i], emit[i], off_sr[i], off_ref[i], cands, seqs, stride, refs);
   return 0;
}
}
