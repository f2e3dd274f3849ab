// tests/emul/rp_emul.cpp -- TEST INFRASTRUCTURE ONLY: the pair predicates of pindel_b200/csrc/pg_rp.cuh (the code the
// read-pair clustering kernels run) compiled as plain C++, driven the way pg_rp_modify_kernel drives them -- per (DA, DB)
// group, rows in order, "first applicable row at or after the cursor", update, rescan -- and the grouping the way
// pg_rp_group_kernel / pg_rp_report_kernel do it (first earlier unvisited row with the same event; counters).  No GPU.
#include <stdint.h>
#include <string.h>
#include <vector>
#include "../../pindel_b200/csrc/pg_rp.cuh"
#include "../../pindel_b200/csrc/pg_rp_host.h"

extern "C" {

int rpemu_sizes(int which) { return which == 0 ? (int)sizeof(PgRp) : (int)sizeof(PgRpCluster); }

void rpemu_modify(int n, const PgRp *in, PgRpCluster *out)
{
   // the product's own class split (pg_rp_host.h); a small warp_max so that both lists are exercised
   std::vector<uint32_t> rows_small, off_small, rows_big, off_big;
   pg_rp_classes(n, in, 64u, rows_small, off_small, rows_big, off_big);
   for (int i = 0; i < n; i++) pg_rp_finish(in[i], pg_rp_init(in[i]), out[i], (uint32_t)i);   // pg_rp_single_kernel
   for (int list = 0; list < 2; list++) {
      const std::vector<uint32_t> &all = list ? rows_big : rows_small, &off = list ? off_big : off_small;
      for (size_t g = 0; g + 1 < off.size(); g++) {
         const uint32_t *rows = all.data() + off[g];
         const uint32_t ng = off[g + 1] - off[g];
         const uint8_t da = in[rows[0]].da, db = in[rows[0]].db;
         std::vector<PgRpIv> w(ng);
         for (uint32_t k = 0; k < ng; k++) w[k] = pg_rp_init(in[rows[k]]);
         for (uint32_t first = 0; first < ng; first++) {
            PgRpIv f = w[first];
            if (pg_rp_iabs((int)(f.a - f.a1)) > 1000 || pg_rp_iabs((int)(f.b - f.b1)) > 1000) continue;
            for (uint32_t cursor = 0; cursor < ng;) {
               uint32_t jm = ng;
               for (uint32_t j = cursor; j < ng; j++)
                  if (pg_rp_applicable(f, w[j], da, db)) { jm = j; break; }
               if (jm == ng) break;
               pg_rp_apply(f, w[jm], da, db);
               cursor = jm + 1;
            }
            w[first] = f;
         }
         for (uint32_t k = 0; k < ng; k++) pg_rp_finish(in[rows[k]], w[k], out[rows[k]], rows[k]);
      }
   }
}

// statistics of the class split: [0] classes, [1] rows in classes, [2] largest class
void rpemu_class_stats(int n, const PgRp *in, uint32_t *stats)
{
   std::vector<uint32_t> rows_small, off_small, rows_big, off_big;
   pg_rp_classes(n, in, 0xFFFFFFFFu, rows_small, off_small, rows_big, off_big);
   stats[0] = (uint32_t)off_small.size() - 1;
   stats[1] = (uint32_t)rows_small.size();
   stats[2] = 0;
   for (size_t g = 0; g + 1 < off_small.size(); g++) if (off_small[g + 1] - off_small[g] > stats[2]) stats[2] = off_small[g + 1] - off_small[g];
}

void rpemu_summarize(int n, const PgRp *in, PgRpCluster *st)
{
   if (n < 5) {
      for (int i = 0; i < n; i++) st[i].report = 0;
      return;
   }
   std::vector<uint32_t> merged((size_t)n), nid((size_t)n, 0);
   uint32_t counters[2] = {0, 0};
   for (int i = 0; i < n; i++) {
      uint32_t rep = (uint32_t)i;
      if (!st[i].visited)
         for (int j = 0; j < i; j++)
            if (!st[j].visited && pg_rp_same_event(st[i], in[i], st[j], in[j])) { rep = (uint32_t)j; break; }
      merged[i] = rep;
      if (!st[i].visited) {
         if (rep != (uint32_t)i || i <= n - 2) nid[rep]++;
         if (rep == (uint32_t)i && i <= n - 2) { counters[0]++; if ((uint32_t)i > counters[1]) counters[1] = (uint32_t)i; }
      }
   }
   for (int i = 0; i < n; i++) {
      PgRpCluster &o = st[i];
      if (o.visited) continue;
      if (merged[i] != (uint32_t)i) { o.visited = 1; o.merged_into = merged[i]; }
      else if (i <= n - 2) {
         o.n_identical = nid[i];
         if (counters[0] == 1u || (uint32_t)i != counters[1]) o.report = o.n_identical >= 5u ? 1 : 0;   // the last good row keeps its Report
      }
   }
}

// the inter-chromosomal variants in the kernels' formulation (pg_rp_modify_inter_warp_kernel, pg_rp_inter_*_kernel)
void rpemu_modify_interchr(int n, const PgRp *in, const int32_t *chr_a, const int32_t *chr_b, PgRpCluster *out)
{
   std::vector<uint8_t> flip;
   std::vector<uint32_t> all, off;
   pg_rp_classes_interchr(n, in, chr_a, chr_b, flip, all, off);
   for (int i = 0; i < n; i++) pg_rp_finish_inter(pg_rp_init(in[i]), out[i], (uint32_t)i);
   const uint32_t last_row = (uint32_t)(n - 1);
   for (size_t g = 0; g + 1 < off.size(); g++) {
      const uint32_t *rows = all.data() + off[g];
      const uint32_t ng = off[g + 1] - off[g];
      const bool f0 = flip[rows[0]] != 0;
      const uint8_t da = f0 ? in[rows[0]].db : in[rows[0]].da, db = f0 ? in[rows[0]].da : in[rows[0]].db;
      std::vector<PgRpIv> w(ng);
      for (uint32_t k = 0; k < ng; k++) { w[k] = pg_rp_init(in[rows[k]]); if (flip[rows[k]]) w[k] = pg_rp_swap_sides(w[k]); }
      for (uint32_t first = 0; first < ng; first++) {
         if (rows[first] == last_row) continue;
         PgRpIv f = w[first];
         for (uint32_t cursor = 0; cursor < ng;) {
            uint32_t jm = ng;
            for (uint32_t j = cursor; j < ng; j++)
               if (pg_rp_applicable_inter(f, w[j], da, db)) { jm = j; break; }
            if (jm == ng) break;
            pg_rp_apply(f, w[jm], da, db);
            cursor = jm + 1;
         }
         w[first] = f;
      }
      for (uint32_t k = 0; k < ng; k++) pg_rp_finish_inter(flip[rows[k]] ? pg_rp_swap_sides(w[k]) : w[k], out[rows[k]], rows[k]);
   }
}

void rpemu_summarize_interchr(int n, const PgRp *in, const int32_t *chr_a, const int32_t *chr_b, PgRpCluster *st)
{
   std::vector<uint32_t> merged((size_t)n), nid((size_t)n, 0);
   for (int i = 0; i < n; i++) {
      uint32_t rep = (uint32_t)i;
      if (!st[i].visited)
         for (int j = 0; j < i; j++)
            if (!st[j].visited && pg_rp_same_event_inter(st[i], in[i], chr_a[i], chr_b[i], st[j], in[j], chr_a[j], chr_b[j])) { rep = (uint32_t)j; break; }
      merged[i] = rep;
      if (!st[i].visited && rep != (uint32_t)i) nid[rep]++;
   }
   for (int i = 0; i < n; i++) {
      PgRpCluster &o = st[i];
      if (o.visited) continue;
      if (merged[i] != (uint32_t)i) { o.visited = 1; o.merged_into = merged[i]; }
      else if (i <= n - 2) { o.n_identical = nid[i]; o.report = o.n_identical >= 5u ? 1 : 0; }
   }
}
}
