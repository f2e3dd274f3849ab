// pindel_b200/csrc/pg_rp.cuh -- discordant-read-pair clustering: the two quadratic loops
// of BDData::UpdateBD (src/bddata.cpp:646-659), ModifyRP (src/bddata.cpp:385-431) and Summarize (src/bddata.cpp:470-575).
//
// ModifyRP is a recurrence, not a map: row `first` is compared with every row `second` in index order and may take over
// second's A and/or B interval; rows before it already hold their final intervals, rows after it their initial ones, and its
// own interval changes while the inner loop runs.  (The reference runs the outer loop under `omp for`, reading rows other
// threads are writing; only the serial order is defined, and that is what is reproduced here, bit for bit.)  What makes
// it fast on a GPU without changing the order: a take-over needs the two A intervals (or the two B intervals) to overlap
// in one particular way, which almost no pair does, so the inner loop becomes a block-wide *find first applicable row
// at or after the cursor* -- 32 (warp) or 1024 (CTA) rows per step -- followed by the update and a rescan from the row
// after it.
// Rows interact only inside small classes, which the host derives in O(n log n) (pg_rp_modify in pindel_pg.cu): equal
// (DA, DB) (RecipicalOverlap demands it), and connected through overlapping A intervals or overlapping B intervals -- a
// take-over copies an interval that overlaps the current one, so a row's A interval never leaves the overlap component of
// its initial A interval, likewise for B; rows of different classes can therefore never pass an interval test.  Every
// class is an independent recurrence: one warp per class (pg_rp_modify_warp_kernel), one CTA for the rare class of more
// than 1024 rows (pg_rp_modify_kernel, intervals in shared memory).  Relative row order inside a class is kept.
//
// The pair predicates are __host__ __device__ so that the CPU suite runs the same code (tests/emul/rp_emul.cpp).
#pragma once
#include <stdint.h>

#ifdef __CUDACC__
#define PGRP_HD __host__ __device__ __forceinline__
#else
#define PGRP_HD inline
#endif

struct PgRp {            // pg_rp (include/pindel_pg.h)
   uint32_t pos_a, pos_b;
   uint32_t insert_size;
   int16_t read_length;
   uint8_t da, db;
};

struct PgRpCluster {     // pg_rp_cluster
   uint32_t pos_a, pos_a1, pos_b, pos_b1;
   uint32_t n_identical;
   uint32_t merged_into;
   uint8_t visited, report;
   uint16_t pad;
};

struct PgRpIv {          // the four interval ends of a row: A = [a, a1], B = [b, b1]
   uint32_t a, a1, b, b1;
};

PGRP_HD int pg_rp_iabs(int x) { return x < 0 ? -x : x; }

// InitializeA1B1, src/bddata.cpp:255-297
PGRP_HD PgRpIv pg_rp_init(const PgRp &r)
{
   const uint32_t distance = r.insert_size;
   const int rl = r.read_length;
   PgRpIv v;
   v.a = r.pos_a;
   v.b = r.pos_b;
   if (r.da == '+') {
      v.a = v.a > (uint32_t)rl * 2u ? v.a - (uint32_t)(rl * 2) : 1u;
      v.a1 = v.a + distance + (uint32_t)(rl * 2);
   } else {
      v.a = v.a > distance ? v.a - distance : 1u;
      v.a1 = v.a + distance + (uint32_t)rl;
   }
   if (r.db == '+') v.b = v.b > (uint32_t)rl * 2u ? v.b - (uint32_t)(rl * 2) : 1u;
   else v.b = v.b > distance ? v.b - distance : 1u;
   v.b1 = v.b + distance + (uint32_t)rl;
   return v;
}

// the interval tests of ProcessSameChromosomeSameStrand, src/bddata.cpp:306-337
PGRP_HD bool pg_rp_takes_a(const PgRpIv &f, const PgRpIv &s, uint8_t da)
{
   return (da == '+' && f.a < s.a && s.a < f.a1 && f.a1 < s.a1) || (da == '-' && f.a < s.a1 && s.a1 < f.a1 && s.a < f.a);
}
PGRP_HD bool pg_rp_takes_b(const PgRpIv &f, const PgRpIv &s, uint8_t db)
{
   return (db == '+' && f.b < s.b && s.b < f.b1 && f.b1 < s.b1) || (db == '-' && s.b < f.b && f.b < s.b1 && s.b1 < f.b1);
}

// RecipicalOverlap, src/bddata.cpp:183-237, for two rows of one (DA, DB) group.  The ratios are IEEE double divisions
// compared with the float constant 0.9f, as in the reference.
PGRP_HD bool pg_rp_reciprocal_overlap(const PgRpIv &f, const PgRpIv &s)
{
   if (pg_rp_iabs((int)(f.a - f.a1)) > 1000 || pg_rp_iabs((int)(f.b - f.b1)) > 1000) return false;
   if (pg_rp_iabs((int)(s.a - s.a1)) > 1000 || pg_rp_iabs((int)(s.b - s.b1)) > 1000) return false;
   const double cutoff = (double)0.9f;
   uint32_t fa = (f.a + f.a1) / 2u, fb = (f.b + f.b1) / 2u, sa = (s.a + s.a1) / 2u, sb = (s.b + s.b1) / 2u;
   if (fa > fb) { const uint32_t t = fa; fa = fb; fb = t; }
   if (sa > sb) { const uint32_t t = sa; sa = sb; sb = t; }
   if (fa > sb + 200u || fb + 200u < sa) return false;
   if (fa <= sa && sb <= fb && (double)(sb - sa) / (double)(fb - fa) >= cutoff) return true;
   if (sa <= fa && fb <= sb && (double)(fb - fa) / (double)(sb - sa) >= cutoff) return true;
   if (fa <= sa && sa <= fb && fb <= sb && (double)(fb - sa) / (double)(fb - fa) >= cutoff && (double)(fb - sa) / (double)(sb - sa) >= cutoff)
      return true;
   if (sa <= fa && fa <= sb && sb <= fb && (double)(sb - fa) / (double)(fb - fa) >= cutoff && (double)(sb - fa) / (double)(sb - sa) >= cutoff)
      return true;
   return false;
}

// does row s change row f?  (the cheap interval tests first: they fail for almost every pair)
PGRP_HD bool pg_rp_applicable(const PgRpIv &f, const PgRpIv &s, uint8_t da, uint8_t db)
{
   if (!(pg_rp_takes_a(f, s, da) || pg_rp_takes_b(f, s, db))) return false;
   if (!pg_rp_reciprocal_overlap(f, s)) return false;
   return !(s.a1 - s.a > 10000u || s.b1 - s.b > 10000u);   // the guard at src/bddata.cpp:301-305
}

PGRP_HD void pg_rp_apply(PgRpIv &f, const PgRpIv &s, uint8_t da, uint8_t db)
{
   if (pg_rp_takes_a(f, s, da)) { f.a = s.a; f.a1 = s.a1; }
   if (pg_rp_takes_b(f, s, db)) { f.b = s.b; f.b1 = s.b1; }
}

// the loop after the double loop, src/bddata.cpp:418-430
PGRP_HD void pg_rp_finish(const PgRp &r, const PgRpIv &v, PgRpCluster &o, uint32_t index)
{
   o.pos_a = v.a; o.pos_a1 = v.a1; o.pos_b = v.b; o.pos_b1 = v.b1;
   if (r.da == '+') { o.pos_a += (uint32_t)(int)r.read_length; o.pos_a1 += (uint32_t)(int)r.read_length; }
   if (r.db == '+') { o.pos_b += (uint32_t)(int)r.read_length; o.pos_b1 += (uint32_t)(int)r.read_length; }
   o.visited = pg_rp_iabs((int)(o.pos_a - o.pos_b)) < 500 ? 1 : 0;   // ChrNameA == ChrNameB for the rows of this call
   o.report = 0;
   o.n_identical = 0;
   o.merged_into = index;
   o.pad = 0;
}

PGRP_HD bool pg_rp_same_event(const PgRpCluster &x, const PgRp &xi, const PgRpCluster &y, const PgRp &yi)
{
   return x.pos_a == y.pos_a && x.pos_b == y.pos_b && x.pos_a1 == y.pos_a1 && x.pos_b1 == y.pos_b1 && xi.da == yi.da && xi.db == yi.db;
}

// ---- the inter-chromosomal variants: ModifyRP_InterChr (src/bddata.cpp:439-468) and Summarize_InterChr (:577-624) -------
// UpdateFirstBasedOnSecondInterChromosome (:362-382) has no RecipicalOverlap test, and a row (X, Y, DA, DB) also meets the
// rows (Y, X, DB, DA), seen "from their other end".  Every row is therefore put into a canonical orientation (smaller
// chromosome number first: its A and B sides swapped when needed); in that orientation two rows interact exactly when
// their chromosome pair and strands are equal, the take-over tests are the same on both sides, and the recurrence, the
// class split and the kernels are those of the intra-chromosomal case with a simpler predicate.
PGRP_HD PgRpIv pg_rp_swap_sides(const PgRpIv &v)
{
   PgRpIv r;
   r.a = v.b; r.a1 = v.b1; r.b = v.a; r.b1 = v.a1;
   return r;
}
PGRP_HD bool pg_rp_applicable_inter(const PgRpIv &f, const PgRpIv &s, uint8_t da, uint8_t db)
{
   if (!(pg_rp_takes_a(f, s, da) || pg_rp_takes_b(f, s, db))) return false;
   return !(s.a1 - s.a > 10000u || s.b1 - s.b > 10000u);   // ProcessSameChromosomeSameStrand's guard, src/bddata.cpp:301-305
}
PGRP_HD void pg_rp_finish_inter(const PgRpIv &v, PgRpCluster &o, uint32_t index)   // no + ReadLength pass, no Visited mark
{
   o.pos_a = v.a; o.pos_a1 = v.a1; o.pos_b = v.b; o.pos_b1 = v.b1;
   o.visited = 0;
   o.report = 0;
   o.n_identical = 0;
   o.merged_into = index;
   o.pad = 0;
}
PGRP_HD bool pg_rp_same_event_inter(const PgRpCluster &x, const PgRp &xi, int32_t xa, int32_t xb, const PgRpCluster &y, const PgRp &yi, int32_t ya,
                                    int32_t yb)
{
   return xa == ya && xb == yb && x.pos_a == y.pos_a && x.pos_b == y.pos_b && xi.da == yi.da && xi.db == yi.db;
}

#ifdef __CUDACC__
#define PG_RP_THREADS 1024

// One CTA per (DA, DB) group.  rows: indices of the group's rows in `in`, in order.  work: the group's intervals, in shared
// memory when the group fits (smem_rows), else in the global scratch the caller provides.
__global__ void __launch_bounds__(PG_RP_THREADS) pg_rp_modify_kernel(const PgRp *__restrict__ in, const uint32_t *__restrict__ group_rows,
                                                                     const uint32_t *__restrict__ group_off, uint4 *__restrict__ scratch,
                                                                     uint32_t smem_rows, PgRpCluster *__restrict__ out)
{
   extern __shared__ __align__(16) unsigned char pg_rp_smem[];
   __shared__ int hit[2];   // block-wide minimum, double-buffered: pass p uses hit[p & 1] and thread 0 re-arms it after the pass
   const uint32_t g0 = group_off[blockIdx.x], ng = group_off[blockIdx.x + 1] - g0;
   if (ng == 0) return;
   const uint32_t *rows = group_rows + g0;
   uint4 *w = ng <= smem_rows ? reinterpret_cast<uint4 *>(pg_rp_smem) : scratch + g0;
   const int tid = threadIdx.x;
   const uint8_t da = in[rows[0]].da, db = in[rows[0]].db;
   for (uint32_t k = tid; k < ng; k += PG_RP_THREADS) {
      const PgRpIv v = pg_rp_init(in[rows[k]]);
      w[k] = make_uint4(v.a, v.a1, v.b, v.b1);
   }
   if (tid == 0) hit[0] = hit[1] = 0x7FFFFFFF;
   __syncthreads();
   uint32_t pass = 0;
   for (uint32_t first = 0; first < ng; first++) {
      const uint4 q = w[first];
      PgRpIv f = {q.x, q.y, q.z, q.w};
      // RecipicalOverlap is false for every second when first's own intervals are wider than 1000
      if (pg_rp_iabs((int)(f.a - f.a1)) > 1000 || pg_rp_iabs((int)(f.b - f.b1)) > 1000) continue;
      bool changed = false;
      for (uint32_t cursor = 0; cursor < ng;) {
         const uint32_t j = cursor + (uint32_t)tid;
         bool app = false;
         if (j < ng) {
            const uint4 t = w[j];
            const PgRpIv s = {t.x, t.y, t.z, t.w};
            app = pg_rp_applicable(f, s, da, db);
         }
         if (!__syncthreads_or(app)) {
            cursor += PG_RP_THREADS;
            continue;
         }
         int *slot = &hit[pass & 1u];
         pass++;
         const uint32_t m = __ballot_sync(0xFFFFFFFFu, app);
         if (m && (tid & 31) == 0) atomicMin(slot, (int)(cursor + (uint32_t)tid + (uint32_t)__ffs(m) - 1u));
         __syncthreads();
         const uint32_t jm = (uint32_t)*slot;
         __syncthreads();
         if (tid == 0) *slot = 0x7FFFFFFF;   // used again two updates from now, with the next update's barriers in between
         const uint4 t = w[jm];
         const PgRpIv s = {t.x, t.y, t.z, t.w};
         pg_rp_apply(f, s, da, db);
         changed = true;
         cursor = jm + 1u;
      }
      if (changed) {   // uniform: every thread followed the same updates
         if (tid == 0) w[first] = make_uint4(f.a, f.a1, f.b, f.b1);
         __syncthreads();
      }
   }
   __syncthreads();
   for (uint32_t k = tid; k < ng; k += PG_RP_THREADS) {
      const uint4 t = w[k];
      const PgRpIv v = {t.x, t.y, t.z, t.w};
      PgRpCluster o;
      pg_rp_finish(in[rows[k]], v, o, rows[k]);
      out[rows[k]] = o;
   }
}

// One warp per class (<= 1024 rows; typically the pairs of one breakpoint).  scratch: class layout, like group_rows.
__global__ void __launch_bounds__(256) pg_rp_modify_warp_kernel(const PgRp *__restrict__ in, const uint32_t *__restrict__ group_rows,
                                                                const uint32_t *__restrict__ group_off, uint32_t n_groups,
                                                                uint4 *__restrict__ scratch, PgRpCluster *__restrict__ out)
{
   const uint32_t g = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31u;
   if (g >= n_groups) return;
   const uint32_t g0 = group_off[g], ng = group_off[g + 1] - g0;
   const uint32_t *rows = group_rows + g0;
   uint4 *w = scratch + g0;
   const uint8_t da = in[rows[0]].da, db = in[rows[0]].db;
   for (uint32_t k = lane; k < ng; k += 32u) {
      const PgRpIv v = pg_rp_init(in[rows[k]]);
      w[k] = make_uint4(v.a, v.a1, v.b, v.b1);
   }
   __syncwarp();
   for (uint32_t first = 0; first < ng; first++) {
      const uint4 q = w[first];
      PgRpIv f = {q.x, q.y, q.z, q.w};
      if (pg_rp_iabs((int)(f.a - f.a1)) > 1000 || pg_rp_iabs((int)(f.b - f.b1)) > 1000) continue;
      bool changed = false;
      for (uint32_t cursor = 0; cursor < ng;) {
         const uint32_t j = cursor + lane;
         bool app = false;
         if (j < ng) {
            const uint4 t = w[j];
            const PgRpIv s = {t.x, t.y, t.z, t.w};
            app = pg_rp_applicable(f, s, da, db);
         }
         const uint32_t m = __ballot_sync(0xFFFFFFFFu, app);
         if (!m) {
            cursor += 32u;
            continue;
         }
         const uint32_t jm = cursor + (uint32_t)__ffs(m) - 1u;
         const uint4 t = w[jm];
         const PgRpIv s = {t.x, t.y, t.z, t.w};
         pg_rp_apply(f, s, da, db);
         changed = true;
         cursor = jm + 1u;
      }
      if (changed) {
         if (lane == 0) w[first] = make_uint4(f.a, f.a1, f.b, f.b1);
         __syncwarp();
      }
   }
   __syncwarp();
   for (uint32_t k = lane; k < ng; k += 32u) {
      const uint4 t = w[k];
      const PgRpIv v = {t.x, t.y, t.z, t.w};
      PgRpCluster o;
      pg_rp_finish(in[rows[k]], v, o, rows[k]);
      out[rows[k]] = o;
   }
}

// rows that are alone in their class: InitializeA1B1 and the closing pass only
__global__ void pg_rp_single_kernel(int n, const PgRp *__restrict__ in, PgRpCluster *__restrict__ out)
{
   const int i = blockIdx.x * blockDim.x + threadIdx.x;
   if (i >= n) return;
   PgRpCluster o;
   pg_rp_finish(in[i], pg_rp_init(in[i]), o, (uint32_t)i);
   out[i] = o;
}

// Summarize, first loop (src/bddata.cpp:480-513): every unvisited row joins the first unvisited row with the same
// intervals and strands.  One warp per row scans the rows before it, 32 at a time, and stops at the first match; what it
// scans is one 64-bit key per row (a mix of the six compared fields; visited rows get a key nothing matches), so the
// quadratic part moves 8 bytes per pair, and the fields themselves are only read to confirm a key match.
// counters: [0] number of "good" rows (unvisited firsts with index <= n - 2), [1] the largest such index.
PGRP_HD uint64_t pg_rp_mix(uint64_t h, uint64_t v)
{
   h ^= v + 0x9E3779B97F4A7C15ull + (h << 6) + (h >> 2);
   h *= 0xFF51AFD7ED558CCDull;
   return h ^ (h >> 33);
}

__global__ void pg_rp_key_kernel(int n, const PgRp *__restrict__ in, const PgRpCluster *__restrict__ st, uint64_t *__restrict__ keys)
{
   const int i = blockIdx.x * blockDim.x + threadIdx.x;
   if (i >= n) return;
   const PgRpCluster o = st[i];
   uint64_t h = pg_rp_mix(0x243F6A8885A308D3ull, (uint64_t)o.pos_a << 32 | o.pos_a1);
   h = pg_rp_mix(h, (uint64_t)o.pos_b << 32 | o.pos_b1);
   h = pg_rp_mix(h, (uint64_t)in[i].da << 8 | in[i].db);
   keys[i] = o.visited ? ~(uint64_t)i : (h & 0x7FFFFFFFFFFFFFFFull);   // visited rows: top bit set, all different
}

__global__ void __launch_bounds__(256) pg_rp_group_kernel(int n, const PgRp *__restrict__ in, const PgRpCluster *__restrict__ st,
                                                          const uint64_t *__restrict__ keys, uint32_t *__restrict__ merged_into,
                                                          uint32_t *__restrict__ n_identical, uint32_t *__restrict__ counters)
{
   const int i = (int)((blockIdx.x * blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
   if (i >= n) return;
   const PgRpCluster me = st[i];
   const PgRp mi = in[i];
   const uint64_t mykey = keys[i];
   uint32_t rep = (uint32_t)i;
   if (!me.visited) {
      for (int c0 = 0; c0 < i; c0 += 32) {
         const int j = c0 + lane;
         bool eq = j < i && keys[j] == mykey;
         if (eq) {                                  // rare: confirm on the fields themselves
            const PgRpCluster o = st[j];
            eq = !o.visited && pg_rp_same_event(me, mi, o, in[j]);
         }
         const uint32_t m = __ballot_sync(0xFFFFFFFFu, eq);
         if (m) {
            rep = (uint32_t)(c0 + __ffs(m) - 1);
            break;
         }
      }
   }
   if (lane == 0) {
      merged_into[i] = rep;
      if (!me.visited) {
         if (rep != (uint32_t)i || i <= n - 2) atomicAdd(&n_identical[rep], 1u);
         if (rep == (uint32_t)i && i <= n - 2) {
            atomicAdd(&counters[0], 1u);
            atomicMax(&counters[1], (uint32_t)i);
         }
      }
   }
}

// Summarize, the rest (src/bddata.cpp:515-573): Report for the good rows -- all but the last one when there are several.
__global__ void pg_rp_report_kernel(int n, PgRpCluster *__restrict__ st, const uint32_t *__restrict__ merged_into,
                                    const uint32_t *__restrict__ n_identical, const uint32_t *__restrict__ counters)
{
   const int i = blockIdx.x * blockDim.x + threadIdx.x;
   if (i >= n) return;
   PgRpCluster o = st[i];
   const uint32_t rep = merged_into[i];
   if (!o.visited) {
      if (rep != (uint32_t)i) {
         o.visited = 1;
         o.merged_into = rep;
      } else if (i <= n - 2) {
         o.n_identical = n_identical[i];
         if (counters[0] == 1u || (uint32_t)i != counters[1]) o.report = o.n_identical >= 5u ? 1 : 0;   // the last good row keeps its Report
      }
   }
   st[i] = o;
}
// ---- inter-chromosomal kernels ------------------------------------------------------------------------------------
// every row: InitializeA1B1 only (rows that share a class with nobody, and the default for the rest)
__global__ void pg_rp_inter_single_kernel(int n, const PgRp *__restrict__ in, PgRpCluster *__restrict__ out)
{
   const int i = blockIdx.x * blockDim.x + threadIdx.x;
   if (i >= n) return;
   PgRpCluster o;
   pg_rp_finish_inter(pg_rp_init(in[i]), o, (uint32_t)i);
   out[i] = o;
}

// One warp per class, canonical orientation (flip[row] = the row's sides are swapped in it).  last_row: the global index
// that ModifyRP_InterChr's outer loop never reaches (src/bddata.cpp:456).
__global__ void __launch_bounds__(256) pg_rp_modify_inter_warp_kernel(const PgRp *__restrict__ in, const uint8_t *__restrict__ flip,
                                                                      const uint32_t *__restrict__ group_rows, const uint32_t *__restrict__ group_off,
                                                                      uint32_t n_groups, uint32_t last_row, uint4 *__restrict__ scratch,
                                                                      PgRpCluster *__restrict__ out)
{
   const uint32_t g = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31u;
   if (g >= n_groups) return;
   const uint32_t g0 = group_off[g], ng = group_off[g + 1] - g0;
   const uint32_t *rows = group_rows + g0;
   uint4 *w = scratch + g0;
   const bool f0 = flip[rows[0]] != 0;
   const uint8_t da = f0 ? in[rows[0]].db : in[rows[0]].da, db = f0 ? in[rows[0]].da : in[rows[0]].db;   // the class's strands, canonical
   for (uint32_t k = lane; k < ng; k += 32u) {
      PgRpIv v = pg_rp_init(in[rows[k]]);
      if (flip[rows[k]]) v = pg_rp_swap_sides(v);
      w[k] = make_uint4(v.a, v.a1, v.b, v.b1);
   }
   __syncwarp();
   for (uint32_t first = 0; first < ng; first++) {
      if (rows[first] == last_row) continue;
      const uint4 q = w[first];
      PgRpIv f = {q.x, q.y, q.z, q.w};
      bool changed = false;
      for (uint32_t cursor = 0; cursor < ng;) {
         const uint32_t j = cursor + lane;
         bool app = false;
         if (j < ng) {
            const uint4 t = w[j];
            const PgRpIv s = {t.x, t.y, t.z, t.w};
            app = pg_rp_applicable_inter(f, s, da, db);
         }
         const uint32_t m = __ballot_sync(0xFFFFFFFFu, app);
         if (!m) {
            cursor += 32u;
            continue;
         }
         const uint32_t jm = cursor + (uint32_t)__ffs(m) - 1u;
         const uint4 t = w[jm];
         const PgRpIv s = {t.x, t.y, t.z, t.w};
         pg_rp_apply(f, s, da, db);
         changed = true;
         cursor = jm + 1u;
      }
      if (changed) {
         if (lane == 0) w[first] = make_uint4(f.a, f.a1, f.b, f.b1);
         __syncwarp();
      }
   }
   __syncwarp();
   for (uint32_t k = lane; k < ng; k += 32u) {
      const uint4 t = w[k];
      PgRpIv v = {t.x, t.y, t.z, t.w};
      if (flip[rows[k]]) v = pg_rp_swap_sides(v);
      PgRpCluster o;
      pg_rp_finish_inter(v, o, rows[k]);
      out[rows[k]] = o;
   }
}

// Summarize_InterChr: keys over (ChrNameA, ChrNameB, PosA, PosB, DA, DB); NumberOfIdentical counts the *other* rows
__global__ void pg_rp_inter_key_kernel(int n, const PgRp *__restrict__ in, const int32_t *__restrict__ chr_a, const int32_t *__restrict__ chr_b,
                                       const PgRpCluster *__restrict__ st, uint64_t *__restrict__ keys)
{
   const int i = blockIdx.x * blockDim.x + threadIdx.x;
   if (i >= n) return;
   const PgRpCluster o = st[i];
   uint64_t h = pg_rp_mix(0x243F6A8885A308D3ull, (uint64_t)o.pos_a << 32 | o.pos_b);
   h = pg_rp_mix(h, (uint64_t)(uint32_t)chr_a[i] << 32 | (uint32_t)chr_b[i]);
   h = pg_rp_mix(h, (uint64_t)in[i].da << 8 | in[i].db);
   keys[i] = o.visited ? ~(uint64_t)i : (h & 0x7FFFFFFFFFFFFFFFull);
}

__global__ void __launch_bounds__(256) pg_rp_inter_group_kernel(int n, const PgRp *__restrict__ in, const int32_t *__restrict__ chr_a,
                                                                const int32_t *__restrict__ chr_b, const PgRpCluster *__restrict__ st,
                                                                const uint64_t *__restrict__ keys, uint32_t *__restrict__ merged_into,
                                                                uint32_t *__restrict__ n_identical)
{
   const int i = (int)((blockIdx.x * blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
   if (i >= n) return;
   const PgRpCluster me = st[i];
   const PgRp mi = in[i];
   const int32_t ma = chr_a[i], mb = chr_b[i];
   const uint64_t mykey = keys[i];
   uint32_t rep = (uint32_t)i;
   if (!me.visited) {
      for (int c0 = 0; c0 < i; c0 += 32) {
         const int j = c0 + lane;
         bool eq = j < i && keys[j] == mykey;
         if (eq) {
            const PgRpCluster o = st[j];
            eq = !o.visited && pg_rp_same_event_inter(me, mi, ma, mb, o, in[j], chr_a[j], chr_b[j]);
         }
         const uint32_t m = __ballot_sync(0xFFFFFFFFu, eq);
         if (m) {
            rep = (uint32_t)(c0 + __ffs(m) - 1);
            break;
         }
      }
   }
   if (lane == 0) {
      merged_into[i] = rep;
      if (!me.visited && rep != (uint32_t)i) atomicAdd(&n_identical[rep], 1u);   // the first row itself is not counted
   }
}

__global__ void pg_rp_inter_report_kernel(int n, PgRpCluster *__restrict__ st, const uint32_t *__restrict__ merged_into,
                                          const uint32_t *__restrict__ n_identical)
{
   const int i = blockIdx.x * blockDim.x + threadIdx.x;
   if (i >= n) return;
   PgRpCluster o = st[i];
   if (!o.visited) {
      const uint32_t rep = merged_into[i];
      if (rep != (uint32_t)i) {
         o.visited = 1;
         o.merged_into = rep;
      } else if (i <= n - 2) {                 // the outer loop stops before the last row, src/bddata.cpp:585
         o.n_identical = n_identical[i];
         o.report = o.n_identical >= 5u ? 1 : 0;
      }
   }
   st[i] = o;
}
#endif
