/* oracle/rp_oracle.c -- TEST INFRASTRUCTURE ONLY (CPU restatement; the checker of pg_rp_modify / pg_rp_summarize).
 *
 * Row N2 of SURVEY.md section 8(f): Pindel's discordant-read-pair clustering, the part of BDData::UpdateBD
 * (src/bddata.cpp:646-659) that is quadratic in the number of pairs -- ModifyRP (src/bddata.cpp:385-431, without its
 * std::sort: rows arrive in the order that sort leaves) and Summarize (src/bddata.cpp:470-575).  Written from the
 * reference's behaviour, serial order (the reference's `omp for` over rows reads rows other threads are writing,
 * src/bddata.cpp:401-417, so only the one-thread order is defined).  Pinned against the reference's own functions by
 * tests/test_rp_oracle_pinning.py (live, through oracle/_ref/libpindel_ref.so) and tests/golden/rp_*.npz (committed).
 */
#include <stdint.h>
#include <stdlib.h>

typedef struct {          /* same layout as pg_rp (include/pindel_pg.h) */
   uint32_t pos_a, pos_b;
   uint32_t insert_size;
   int16_t read_length;
   uint8_t da, db;
} rpo_in;

typedef struct {          /* same layout as pg_rp_cluster */
   uint32_t pos_a, pos_a1, pos_b, pos_b1;
   uint32_t n_identical;
   uint32_t merged_into;
   uint8_t visited, report;
   uint16_t pad;
} rpo_out;

static int iabs(int x) { return x < 0 ? -x : x; }

/* InitializeA1B1, src/bddata.cpp:255-297 */
static void init_a1b1(const rpo_in *r, rpo_out *o)
{
   const unsigned distance = r->insert_size;
   const int rl = r->read_length;
   unsigned a = r->pos_a, b = r->pos_b, a1, b1;
   if (r->da == '+') {
      if (a > (unsigned)rl * 2) a = a - rl * 2; else a = 1;
      a1 = a + distance + rl * 2;
   } else {
      if (a > distance) a = a - distance; else a = 1;
      a1 = a + distance + rl;
   }
   if (r->db == '+') {
      if (b > (unsigned)rl * 2) b = b - rl * 2; else b = 1;
      b1 = b + distance + rl;
   } else {
      if (b > distance) b = b - distance; else b = 1;
      b1 = b + distance + rl;
   }
   o->pos_a = a; o->pos_a1 = a1; o->pos_b = b; o->pos_b1 = b1;
}

/* RecipicalOverlap, src/bddata.cpp:183-237 (DA/DB equality is tested by the caller's grouping as well) */
static int reciprocal_overlap(const rpo_in *fi, const rpo_out *f, const rpo_in *si, const rpo_out *s)
{
   const int distance = 1000;
   if (iabs((int)(f->pos_a - f->pos_a1)) > distance) return 0;
   if (iabs((int)(f->pos_b - f->pos_b1)) > distance) return 0;
   if (iabs((int)(s->pos_a - s->pos_a1)) > distance) return 0;
   if (iabs((int)(s->pos_b - s->pos_b1)) > distance) return 0;
   const float cutoff = 0.9f;
   unsigned fa = (f->pos_a + f->pos_a1) / 2, fb = (f->pos_b + f->pos_b1) / 2;
   if (fa > fb) { unsigned t = fa; fa = fb; fb = t; }
   unsigned sa = (s->pos_a + s->pos_a1) / 2, sb = (s->pos_b + s->pos_b1) / 2;
   if (sa > sb) { unsigned t = sa; sa = sb; sb = t; }
   if (fi->da != si->da || fi->db != si->db) return 0;
   if (fa > sb + 200 || fb + 200 < sa) return 0;
   if (fa <= sa && sb <= fb) {
      if ((double)(sb - sa) / (double)(fb - fa) >= cutoff) return 1;
   }
   if (sa <= fa && fb <= sb) {
      if ((double)(fb - fa) / (double)(sb - sa) >= cutoff) return 1;
   }
   if (fa <= sa && sa <= fb && fb <= sb) {
      if ((double)(fb - sa) / (double)(fb - fa) >= cutoff && (double)(fb - sa) / (double)(sb - sa) >= cutoff) return 1;
   }
   if (sa <= fa && fa <= sb && sb <= fb) {
      if ((double)(sb - fa) / (double)(fb - fa) >= cutoff && (double)(sb - fa) / (double)(sb - sa) >= cutoff) return 1;
   }
   return 0;
}

/* ProcessSameChromosomeSameStrand, src/bddata.cpp:299-353 (the > 10000 guards cannot trigger after RecipicalOverlap) */
static void process_pair(const rpo_in *fi, rpo_out *f, const rpo_out *s)
{
   if (s->pos_a1 - s->pos_a > 10000 || s->pos_b1 - s->pos_b > 10000) return;
   if ((fi->da == '+' && f->pos_a < s->pos_a && s->pos_a < f->pos_a1 && f->pos_a1 < s->pos_a1) ||
       (fi->da == '-' && f->pos_a < s->pos_a1 && s->pos_a1 < f->pos_a1 && s->pos_a < f->pos_a)) {
      f->pos_a = s->pos_a;
      f->pos_a1 = s->pos_a1;
   }
   if ((fi->db == '+' && f->pos_b < s->pos_b && s->pos_b < f->pos_b1 && f->pos_b1 < s->pos_b1) ||
       (fi->db == '-' && s->pos_b < f->pos_b && f->pos_b < s->pos_b1 && s->pos_b1 < f->pos_b1)) {
      f->pos_b = s->pos_b;
      f->pos_b1 = s->pos_b1;
   }
}

/* ModifyRP after its sort, src/bddata.cpp:396-430 */
void rpo_modify(int n, const rpo_in *in, rpo_out *out)
{
   for (int i = 0; i < n; i++) {
      init_a1b1(&in[i], &out[i]);
      out[i].n_identical = 0;
      out[i].merged_into = (uint32_t)i;
      out[i].visited = 0;
      out[i].report = 0;
      out[i].pad = 0;
   }
   for (int first = 0; first < n; first++)
      for (int second = 0; second < n; second++)
         if (reciprocal_overlap(&in[first], &out[first], &in[second], &out[second]))
            process_pair(&in[first], &out[first], &out[second]);   /* UpdateFirstBasedOnSecondIntraChromosome: same strands */
   for (int i = 0; i < n; i++) {
      if (in[i].da == '+') { out[i].pos_a += in[i].read_length; out[i].pos_a1 += in[i].read_length; }
      if (in[i].db == '+') { out[i].pos_b += in[i].read_length; out[i].pos_b1 += in[i].read_length; }
      if (iabs((int)(out[i].pos_a - out[i].pos_b)) < 500) out[i].visited = 1;   /* same chromosome by construction */
   }
}

static int same_event(const rpo_in *ai, const rpo_out *a, const rpo_in *bi, const rpo_out *b)
{
   return a->pos_a == b->pos_a && a->pos_b == b->pos_b && a->pos_a1 == b->pos_a1 && a->pos_b1 == b->pos_b1 && ai->da == bi->da &&
          ai->db == bi->db;
}

/* Summarize, src/bddata.cpp:470-575 */
void rpo_summarize(int n, const rpo_in *in, rpo_out *io)
{
   const unsigned cutoff = 5;
   if (n < 5) {
      for (int i = 0; i < n; i++) io[i].report = 0;
      return;
   }
   int *good = (int *)malloc(sizeof(int) * (size_t)n);
   int n_good = 0;
   for (int first = 0; first < n - 1; first++) {
      if (io[first].visited) continue;
      io[first].n_identical = 1;
      for (int second = first + 1; second < n; second++) {
         if (io[second].visited) continue;
         if (same_event(&in[first], &io[first], &in[second], &io[second])) {
            io[first].n_identical++;
            io[second].visited = 1;
            io[second].merged_into = (uint32_t)first;   /* its Tags are appended to the first row's */
         }
      }
      good[n_good++] = first;
   }
   if (n_good == 1) {
      io[good[0]].report = io[good[0]].n_identical >= cutoff;
   } else if (n_good > 1) {
      for (int a = 0; a < n_good - 1; a++) {       /* the last good row is never looked at */
         if (io[good[a]].visited) continue;
         for (int b = a + 1; b < n_good; b++) {
            if (io[good[b]].visited) continue;
            if (same_event(&in[good[a]], &io[good[a]], &in[good[b]], &io[good[b]])) {   /* cannot hold after the loop above */
               io[good[a]].n_identical += io[good[b]].n_identical;
               io[good[b]].visited = 1;
               io[good[b]].merged_into = (uint32_t)good[a];
            }
         }
         io[good[a]].report = io[good[a]].n_identical >= cutoff;
      }
   }
   free(good);
}

/* ---- the inter-chromosomal variants, src/bddata.cpp:439-468 (ModifyRP_InterChr after its sort) and :577-624
 * (Summarize_InterChr).  chr_a / chr_b: ChrNameA / ChrNameB as numbers (equal numbers <=> equal names). ---- */

/* UpdateFirstBasedOnSecondInterChromosome, src/bddata.cpp:362-382: no RecipicalOverlap here */
static void update_interchr(const rpo_in *fi, rpo_out *f, int32_t fa, int32_t fb, const rpo_in *si, const rpo_out *s, int32_t sa, int32_t sb)
{
   if (fa == sa && fb == sb) {
      if (fi->da == si->da && fi->db == si->db) process_pair(fi, f, s);
   } else if (fa == sb && fb == sa) {
      if (fi->da == si->db && fi->db == si->da) {
         rpo_out t = *s;       /* Current_AnotherSecond: the second row seen from its other end */
         t.pos_a = s->pos_b; t.pos_a1 = s->pos_b1; t.pos_b = s->pos_a; t.pos_b1 = s->pos_a1;
         process_pair(fi, f, &t);
      }
   }
}

void rpo_modify_interchr(int n, const rpo_in *in, const int32_t *chr_a, const int32_t *chr_b, rpo_out *out)
{
   for (int i = 0; i < n; i++) {
      init_a1b1(&in[i], &out[i]);
      out[i].n_identical = 0;
      out[i].merged_into = (uint32_t)i;
      out[i].visited = 0;
      out[i].report = 0;
      out[i].pad = 0;
   }
   for (int first = 0; first < n - 1; first++)          /* the last row is never updated, :456 */
      for (int second = 0; second < n; second++)
         update_interchr(&in[first], &out[first], chr_a[first], chr_b[first], &in[second], &out[second], chr_a[second], chr_b[second]);
}

void rpo_summarize_interchr(int n, const rpo_in *in, const int32_t *chr_a, const int32_t *chr_b, rpo_out *io)
{
   if (n == 0) return;
   for (int first = 0; first < n - 1; first++) {
      if (io[first].visited) continue;
      io[first].n_identical = 0;
      for (int second = first + 1; second < n; second++) {
         if (io[second].visited) continue;
         if (chr_a[first] == chr_a[second] && chr_b[first] == chr_b[second] && io[first].pos_a == io[second].pos_a &&
             io[first].pos_b == io[second].pos_b && in[first].da == in[second].da && in[first].db == in[second].db) {
            io[first].n_identical++;
            io[second].visited = 1;
            io[second].merged_into = (uint32_t)first;
         }
      }
      io[first].report = io[first].n_identical >= 5;
   }
}
