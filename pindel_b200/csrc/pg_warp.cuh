// pindel_b200/csrc/pg_warp.cuh
//
// The warp program of the B200 pattern-growth search: ONE warp owns ONE anchored read and runs
// the reference's whole per-read path on it --
//   GetCloseEnd        src/pindel.cpp:2531-2605  (GetCloseEndInner :2250-2396,
//                      CheckLeft_Close / CheckRight_Close src/searcher.cpp:153-197 / 247-286,
//                      ExtendMatchClose :65-97, CategorizePositions :48-63)
//   SearchFarEnd       src/pindel.cpp:1001-1074  (SearchFarEndAtPos src/farend_searcher.cpp:46-103,
//                      CheckBoth src/pindel.cpp:2823-2902, ExtendMatch :2673-2725)
//   CheckMismatches    src/searcher.cpp:331-388
// -- re-formulated for SIMT instead of translated:
//
//  * The reference grows per-mismatch-level position vectors one read base at a time.  Those
//    vectors are only ever consumed through their sizes and, when a size is 1, their single
//    element.  For a seed p let D_t(p) be the read index of its (t+1)-th mismatch (L if none):
//    p sits in a level <= t exactly for lengths len <= D_t(p).  So the number of candidates in
//    levels <= t at length len is #{p : D_t(p) >= len}, and everything CheckLeft/Right_Close and
//    CheckBoth test -- "first non-empty level", "exactly one candidate", "no competitor within
//    ADDITIONAL_MISMATCH levels" -- is decided by the LARGEST and SECOND-LARGEST D_t over all
//    seeds (plus which seed owns the largest), for each of the T = MAX+ADD+1 levels.
//  * One lane per reference position: the 2-bit (bit-plane) packed window is staged in shared
//    memory; a lane funnel-shifts its 32-base stretch out of two broadcast words, XORs it with
//    the read's bit-planes, and pulls its first T mismatch offsets out of the mismatch mask with
//    ffs -- no per-base loop, no dynamic memory, no recursion.  Lane-local top-2 keys are merged
//    with redux.sync; seeds that match too well to be decided inside 32 bases (the true locus)
//    are queued and finished cooperatively (lane c = 32-base chunk c, then lane t = level t).
//  * A final lane-per-length pass replays the reference's control flow (in-range abort,
//    uniqueness, minimum length, CheckMismatches) and emits run-length coded UniquePoints.
//
// The file is written against a tiny SIMT vocabulary (pg_lane, pg_ballot, pg_shfl, pg_redux_*,
// pg_syncwarp ...) so that the very same code also compiles for the CPU lane emulator in
// tests/emul/ (test infrastructure only: the shipped library contains the CUDA build alone).
#pragma once
#include <stdint.h>

#ifndef PG_EMULATE
#define PG_DEV __device__ __forceinline__
#define PG_PIN(x) asm volatile("" : "+r"(x))
#define PG_COLD __device__ __noinline__   // rarely executed helpers: one out-of-line copy keeps the kernel's code small
PG_DEV int pg_lane() { return (int)threadIdx.x; }      // one warp per CTA
PG_DEV uint32_t pg_ballot(bool p) { return __ballot_sync(0xffffffffu, p); }
PG_DEV bool pg_any(bool p) { return __any_sync(0xffffffffu, p); }
PG_DEV uint32_t pg_shfl(uint32_t v, int src) { return __shfl_sync(0xffffffffu, v, src); }
PG_DEV uint32_t pg_shfl_up(uint32_t v, int d) { return __shfl_up_sync(0xffffffffu, v, d); }
PG_DEV uint32_t pg_redux_max(uint32_t v) { return __reduce_max_sync(0xffffffffu, v); }
PG_DEV uint32_t pg_redux_add(uint32_t v) { return __reduce_add_sync(0xffffffffu, v); }
PG_DEV uint32_t pg_redux_or(uint32_t v) { return __reduce_or_sync(0xffffffffu, v); }
PG_DEV void pg_syncwarp() { __syncwarp(); }
PG_DEV int pg_popc(uint32_t v) { return __popc(v); }
PG_DEV int pg_ffs(uint32_t v) { return __ffs((int)v); }
PG_DEV int pg_clz(uint32_t v) { return __clz((int)v); }
PG_DEV uint32_t pg_bfind(uint32_t v) { uint32_t r; asm("bfind.u32 %0, %1;" : "=r"(r) : "r"(v)); return r; }   // index of the highest set bit (v != 0)
PG_DEV uint32_t pg_brev(uint32_t v) { return __brev(v); }
PG_DEV uint32_t pg_fsr(uint32_t lo, uint32_t hi, uint32_t s) { return __funnelshift_r(lo, hi, s); }
PG_DEV uint32_t pg_fsrc(uint32_t lo, uint32_t hi, uint32_t s) { return __funnelshift_rc(lo, hi, s); }
PG_DEV uint32_t pg_atomic_add(uint32_t *p, uint32_t v) { return atomicAdd(p, v); }
PG_DEV void pg_atomic_add64(unsigned long long *p, unsigned long long v) { atomicAdd(p, v); }
PG_DEV uint32_t pg_umax(uint32_t a, uint32_t b) { return max(a, b); }
PG_DEV uint32_t pg_umin(uint32_t a, uint32_t b) { return min(a, b); }
PG_DEV uint32_t pg_ldg32(const uint32_t *p) { return __ldg(p); }
PG_DEV uint2 pg_ldg64(const uint2 *p) { return __ldg(p); }
PG_DEV uint8_t pg_ldg8(const uint8_t *p) { return __ldg(p); }
#endif

#include "pg_warp_types.h"

// Per-read state of pg_process_read that lives across searches (uniform over the warp).
struct PgReadState {
   unsigned long long window_bases;
   int32_t chr_id;
   uint32_t n_regions;
   uint32_t close_runs, close_ups, close_max_len, close_last;   // the close end: PgResult of the search that found it
   uint32_t best_runs, best_ups, best_max_len, far_best;        // the best far end so far and the scratch third that holds it
   uint32_t have_prev, prev_ws, prev_we;                        // the previous far window
   uint32_t filter_ready;
};

// One growth step of the filter: read index `sh` compared with the reference base `sh` to the right of a plus seed and
// `sh` to the left of a minus seed (rsh = 32 - sh is the funnel shift that brings that base under the seed's bit).
struct alignas(16) PgFStep {
   uint32_t lo, hi;   // the read character's two code bits, each spread over a whole word
   uint32_t sh, rsh;
};

// ------------------------------------------------------------------------------------------------
// Per-warp shared memory
// ------------------------------------------------------------------------------------------------
template <int LW, int TL>
struct PgWarpSmem {
   // tile of the packed reference window (filled by TMA on the device)
   alignas(16) uint2 tile[PG_TILE_WORDS];
   alignas(16) uint32_t tile_n[PG_TILE_WORDS];
   alignas(8) uint64_t mbar;
   // read bit-planes: [0] forward order, [1] reversed order; planes lo, hi, N, X(other)
   uint32_t q[2][4][LW];
   // cooperative full-length mismatch masks: N-rule (M) and strict equality (S)
   uint32_t M[TL][LW];
   uint32_t S[TL][LW];
   uint32_t mw[LW];
   // per-level results of the current job
   uint32_t A[TL + 1], B[TL + 1], P[TL + 1], META[TL + 1], H[TL + 1], ALIAS[TL + 1], AM[TL + 1];
   uint32_t queue[PG_QCAP];
   uint16_t seedq[PG_CHUNK_POS];     // seeds of the current chunk: offset from the tile origin | strand << 15
   // threshold filter of the far rings (pg_filter_tile): the read's first characters as a step table, the number of
   // steps after which each level is tested, and the surviving seeds of the staged tile (one bit per position)
   alignas(16) PgFStep ftab[32];
   uint32_t fuse;                    // read indices 1..31 that have a table entry (not N / other, below L)
   uint32_t fcnt[TL];
   uint32_t fsp[PG_TILE_WORDS], fsm[PG_TILE_WORDS];
   PgReadState rs;                   // what pg_process_read keeps across the searches of one read
};

struct PgLevel {      // warp-level top-2 for level == lane
   uint32_t A, B, pos, meta;
};

PG_DEV void pg_level_merge(PgLevel &w, uint32_t D, uint32_t pos, uint32_t meta)
{
   if (D > w.A) { w.B = w.A; w.A = D; w.pos = pos; w.meta = meta; }
   else if (D > w.B) { w.B = D; }
}

PG_DEV void pg_key_update(uint32_t &a, uint32_t &b, uint32_t key)
{
   uint32_t nb = pg_umax(b, pg_umin(a, key));
   a = pg_umax(a, key);
   b = nb;
}

// 32 reference bases starting at absolute position pos, straight from the packed chromosome.
PG_DEV void pg_fetch32_global(const PgChrom &c, int64_t pos, uint32_t &lo, uint32_t &hi, uint32_t &nn)
{
   int64_t w = pos >> 5;             // may be negative inside the guard band
   uint32_t sh = (uint32_t)(pos & 31);
   uint2 a = pg_ldg64(c.planes + w), b = pg_ldg64(c.planes + w + 1);
   lo = pg_fsr(a.x, b.x, sh);
   hi = pg_fsr(a.y, b.y, sh);
   nn = pg_fsr(pg_ldg32(c.nplane + w), pg_ldg32(c.nplane + w + 1), sh);
}

// Describes which read characters a search consumes: chars[j] = plane[arr] (bit j) ^ inv.
struct PgChars {
   int arr;          // 0 forward planes, 1 reversed planes
   uint32_t inv;     // 0 or ~0: complement
};

template <int LW, int TL>
PG_DEV uint32_t pg_valid_mask(int L, int c)
{
   int rem = L - 32 * c;
   return rem >= 32 ? 0xffffffffu : (rem <= 0 ? 0u : ((1u << rem) - 1u));
}

// Mismatch words of chunk c for a seed: x = reference stretch in growth order.
template <int LW, int TL>
PG_DEV void pg_mism_words(const PgWarpSmem<LW, TL> *ws, PgChars ch, int c, int L, uint32_t xlo, uint32_t xhi, uint32_t xn,
                          uint32_t &M, uint32_t &S)
{
   uint32_t qlo = ws->q[ch.arr][0][c] ^ ch.inv, qhi = ws->q[ch.arr][1][c] ^ ch.inv;
   uint32_t qn = ws->q[ch.arr][2][c], qx = ws->q[ch.arr][3][c];
   uint32_t diff = (xlo ^ qlo) | (xhi ^ qhi);
   uint32_t v = pg_valid_mask<LW, TL>(L, c);
   M = (((diff & ~qn) | qx | xn)) & v;                       // Matches(), src/searcher.cpp:36-44
   S = ((((diff | xn) & ~qn) | (~xn & qn) | qx)) & v;        // plain char equality (BP_On_Read != BP_On_Ref)
}

// Reference stretch of chunk c (read indices 32c..32c+31) for a seed at pos growing in direction d.
PG_DEV void pg_seed_stretch(const PgChrom &chrom, uint32_t pos, int d, int c, uint32_t &lo, uint32_t &hi, uint32_t &nn)
{
   if (d > 0) {
      pg_fetch32_global(chrom, (int64_t)pos + 32 * c, lo, hi, nn);
   } else {
      pg_fetch32_global(chrom, (int64_t)pos - 32 * c - 31, lo, hi, nn);
      lo = pg_brev(lo); hi = pg_brev(hi); nn = pg_brev(nn);
   }
}

// Cooperative: lanes c < nchunks compute the full-length masks of one seed into ws->mw (N-rule) or
// ws->M[slot]/ws->S[slot]; returns the total N-rule mismatch count when want_slot >= 0.
template <int LW, int TL>
PG_COLD uint32_t pg_full_masks(PgWarpSmem<LW, TL> *ws, const PgChrom *chrom_p, int ch_arr, uint32_t ch_inv, uint32_t pos, int d,
                               int L, int slot)
{
   const PgChrom chrom = *chrom_p;
   PgChars ch;
   ch.arr = ch_arr; ch.inv = ch_inv;
   int lane = pg_lane();
   int nch = (L + 31) >> 5;
   uint32_t M = 0, S = 0;
   if (lane < nch) {
      uint32_t lo, hi, nn;
      pg_seed_stretch(chrom, pos, d, lane, lo, hi, nn);
      pg_mism_words<LW, TL>(ws, ch, lane, L, lo, hi, nn, M, S);
      if (slot >= 0) { ws->M[slot][lane] = M; ws->S[slot][lane] = S; }
      else ws->mw[lane] = M;
   }
   uint32_t h = pg_redux_add((uint32_t)pg_popc(M));
   pg_syncwarp();
   return h;
}

// ------------------------------------------------------------------------------------------------
// One search = one (or, for BreakDancer clusters, several) windows scanned with one character set.
// ------------------------------------------------------------------------------------------------
struct PgSearch {
   bool far;          // both strands (SearchFarEndAtPos) or one (GetCloseEndInner)
   PgChars chp;       // characters for the '+1' growth (close: the only set; far: plus strand)
   int d;             // close: growth direction; far: unused
   int bp_start;
   int L, MAX, T;
};

template <int LW, int TL>
struct PgWarpCtx {
   PgWarpSmem<LW, TL> *ws;
   const PgParams *prm;
   const PgChrom *chroms;
   PgLevel lvl;               // level == lane
   uint32_t qn;               // queued slow seeds (uniform)
   uint32_t any_seed;         // uniform
};

// device-specific: stage words [w0, w0+nw) of a chromosome into ws->tile / ws->tile_n
template <int LW, int TL>
PG_DEV void pg_tile_load(PgWarpSmem<LW, TL> *ws, const PgChrom &c, int64_t w0, int nw, uint32_t &phase);

template <int LW, int TL>
PG_DEV void pg_flush_queue(PgWarpCtx<LW, TL> &cx, const PgSearch &se, const PgChrom *chrom_p, int chr_id, uint32_t c0lo, uint32_t c0hi)
{
   PgWarpSmem<LW, TL> *ws = cx.ws;
   int lane = pg_lane();
   pg_syncwarp();
   for (uint32_t i = 0; i < cx.qn; i++) {
      uint32_t pos = ws->queue[i];
      int d = se.d;
      PgChars ch = se.chp;
      uint32_t minus = 0;
      if (se.far) {
         // plus seed iff ref[pos] == read[0] (farend_searcher.cpp:76), else it was queued as a minus seed
         uint2 wv = pg_ldg64(chrom_p->planes + (pos >> 5));
         uint32_t rlo = (wv.x >> (pos & 31)) & 1u, rhi = (wv.y >> (pos & 31)) & 1u;
         minus = (rlo != c0lo || rhi != c0hi) ? 1u : 0u;
         d = minus ? -1 : 1;
         ch.inv = minus ? ~se.chp.inv : se.chp.inv;
      }
      pg_full_masks<LW, TL>(ws, chrom_p, ch.arr, ch.inv, pos, d, se.L, -1);
      if (lane < se.T) {
         // D_lane = read index of the (lane+1)-th mismatch at index >= 1, or L
         int rank = lane;
         uint32_t D = (uint32_t)se.L;
         int nch = (se.L + 31) >> 5;
#pragma unroll 1
         for (int c = 0; c < nch; c++) {
            uint32_t w = ws->mw[c];
            if (c == 0) w &= ~1u;
            int pc = pg_popc(w);
            if (rank < pc) {
#pragma unroll 1
               for (int k = 0; k < rank; k++) w &= w - 1;
               D = (uint32_t)(32 * c + pg_ffs(w) - 1);
               break;
            }
            rank -= pc;
         }
         pg_level_merge(cx.lvl, D, pos, ((uint32_t)chr_id << 1) | minus);
      }
      pg_syncwarp();
   }
   cx.qn = 0;
}

// ------------------------------------------------------------------------------------------------
// Threshold filter for the far rings.
//
// A seed can only change the top-2 state (A_t, B_t) of some level t if its D_t exceeds the current second-largest B_t,
// i.e. if it has at most t mismatches among the read indices 1..B_t.  Once a few hundred positions have been scanned the
// B_t are large (B_0 ~ log4 of the seeds seen), so in the outer rings of SearchFarEnd's schedule (src/pindel.cpp:1025-1070)
// all but a handful of the seeds are irrelevant -- and that can be decided for 32 positions per lane at once, bit-sliced:
// the lane owns one 32-base word of the staged tile; for read index j the mismatch word of all 32 plus seeds is the
// reference planes funnel-shifted by j against the (uniform) read character, for the 32 minus seeds the planes shifted the
// other way against its complement; alive_t (seeds with <= t mismatches so far) is a thermometer counter updated with one
// LOP3 per level; after B_t steps alive_t is OR-ed into the survivors and level t is dropped (its members are survivors
// already).  Conservative by construction -- stale thresholds, skipped read N / other characters and reference N (seen as
// 'A') can only under-count mismatches -- and every survivor then takes the exact per-seed path, so the result is
// unchanged; the filter only decides how many seeds take that path.
// ------------------------------------------------------------------------------------------------
#ifdef PG_EMULATE
#ifndef PG_EMUL_TLS
#define PG_EMUL_TLS
#endif
static PG_EMUL_TLS unsigned long long pg_emul_filter_stats[3];    // host build only: seeds in, survivors out, tiles filtered (per emulated warp's thread)
#endif
template <int LW, int TL>
PG_COLD void pg_filter_prepare(PgWarpSmem<LW, TL> *ws, int arr, uint32_t inv, int L)
{
   const int lane = pg_lane();
   const uint32_t qlo = ws->q[arr][0][0] ^ inv, qhi = ws->q[arr][1][0] ^ inv;
   const uint32_t use = pg_valid_mask<LW, TL>(L, 0) & ~1u & ~(ws->q[arr][2][0] | ws->q[arr][3][0]);
   if ((use >> lane) & 1u) {
      PgFStep e;
      e.lo = ((qlo >> lane) & 1u) ? ~0u : 0u;
      e.hi = ((qhi >> lane) & 1u) ? ~0u : 0u;
      e.sh = (uint32_t)lane; e.rsh = 32u - (uint32_t)lane;
      ws->ftab[pg_popc(use & ((1u << lane) - 1u))] = e;
   }
   if (lane == 0) ws->fuse = use;
   pg_syncwarp();
}

// Thresholds of the staged tile from the warp-level state (lane t holds level t): fcnt[t] = number of table steps that cover
// the read indices 1..B_t, made non-increasing towards the lower levels (an earlier test only lets more seeds through).
template <int LW, int TL>
PG_DEV void pg_filter_thresholds(PgWarpCtx<LW, TL> &cx, int T)
{
   PgWarpSmem<LW, TL> *ws = cx.ws;
   const int lane = pg_lane();
   const uint32_t Bm = cx.lvl.B;
   uint32_t c = (uint32_t)pg_popc(ws->fuse & (Bm >= 31u ? ~0u : ((2u << Bm) - 1u)));
   if (lane >= T) c = 0xffffu;
#pragma unroll
   for (int d = 1; d < 32 && d < TL; d <<= 1) {
      uint32_t v = pg_shfl(c, (lane + d) & 31);
      if (lane + d < 32) c = pg_umin(c, v);
   }
   if (lane < TL) ws->fcnt[lane] = c;
   pg_syncwarp();
}

// Survivors of the positions [r0, r1) (relative to the tile origin) into ws->fsp / ws->fsm, one word per tile word.
// Returns (uniform) bit 0: the range holds a seed at all, bit 1: some seed survived.
template <int LW, int TL>
PG_COLD uint32_t pg_filter_tile(PgWarpSmem<LW, TL> *ws, uint32_t r0, uint32_t r1, uint32_t c0lom, uint32_t c0him, int T)
{
   const int lane = pg_lane();
   const uint32_t wlast = (r1 - 1u) >> 5;
   bool any = false, surv = false;
   for (uint32_t wb = r0 >> 5; wb <= wlast; wb += 32u) {
      const uint32_t wi = wb + (uint32_t)lane;
      if (wi <= wlast) {
         const uint2 Wm = ws->tile[wi - 1], W0 = ws->tile[wi], Wp = ws->tile[wi + 1];
         const uint32_t p0 = wi << 5;
         uint32_t inr = ~0u;
         if (p0 < r0) inr &= ~0u << (r0 - p0);
         if (r1 - p0 < 32u) inr &= (1u << (r1 - p0)) - 1u;
         const uint32_t ok = ~ws->tile_n[wi] & inr;
         const uint32_t dl = W0.x ^ c0lom, dh = W0.y ^ c0him;
         const uint32_t plus = ~(dl | dh) & ok, minus = (dl & dh) & ok;
         any = any || (plus | minus) != 0u;
         uint32_t ap[TL], am[TL];
#pragma unroll
         for (int t = 0; t < TL; t++) { ap[t] = plus; am[t] = minus; }
         uint32_t sp = 0, sm = 0;
         uint32_t k = 0;
#pragma unroll
         for (int t0 = 0; t0 < TL; t0++) {
            if (t0 < T) {
               const uint32_t kend = ws->fcnt[t0];
#pragma unroll 1
               for (; k < kend; k++) {
                  const PgFStep e = ws->ftab[k];
                  const uint32_t mp = (pg_fsr(W0.x, Wp.x, e.sh) ^ e.lo) | (pg_fsr(W0.y, Wp.y, e.sh) ^ e.hi);
                  const uint32_t mm = ~(pg_fsr(Wm.x, W0.x, e.rsh) ^ e.lo) | ~(pg_fsr(Wm.y, W0.y, e.rsh) ^ e.hi);
#pragma unroll
                  for (int t = TL - 1; t > t0; t--) {
                     ap[t] = (ap[t] & ~mp) | (ap[t - 1] & mp);
                     am[t] = (am[t] & ~mm) | (am[t - 1] & mm);
                  }
                  ap[t0] &= ~mp; am[t0] &= ~mm;
               }
               sp |= ap[t0]; sm |= am[t0];
            }
         }
         ws->fsp[wi] = sp; ws->fsm[wi] = sm;
         surv = surv || (sp | sm) != 0u;
#ifdef PG_EMULATE
         pg_emul_filter_stats[0] += (unsigned long long)pg_popc(plus | minus);
         pg_emul_filter_stats[1] += (unsigned long long)pg_popc(sp | sm);
#endif
      }
   }
#ifdef PG_EMULATE
   if (lane == 0) pg_emul_filter_stats[2]++;
#endif
   const uint32_t r = (pg_any(any) ? 1u : 0u) | (pg_any(surv) ? 2u : 0u);
   pg_syncwarp();
   return r;
}

// Scan window [s, e) of one chromosome (tile by tile) and fold its seeds into cx.lvl.
//
// Two phases per chunk of PG_CHUNK_POS positions, so that the expensive part runs on full warps:
//   (1) seed list: lane j owns word j of the chunk; which of its 32 positions are seeds is a pure word operation
//       (reference base == first read character, for the far end also == its complement), so a lane finds all of
//       its seeds without looking at the read; a warp prefix sum packs them, in position order, into ws->seedq;
//   (2) seeds are dealt 32 at a time to the lanes: funnel-shift the 32-base stretch out of the staged tile, XOR with
//       the read's bit-planes, pull the first TL mismatch indices out of the mask, update the lane-local top-2 keys.
//
// `filter` (far rings only): the seed list of a tile is first reduced to the seeds that can still beat a second-largest
// D_t of the state so far (pg_filter_tile); the survivors of the whole tile then share the passes of phase (2).
template <int LW, int TL>
PG_DEV void pg_scan_window(PgWarpCtx<LW, TL> &cx, const PgSearch &se, int chr_id, uint32_t s, uint32_t e, uint32_t &phase, bool filter)
{
   PgWarpSmem<LW, TL> *ws = cx.ws;
   const PgChrom chrom = cx.chroms[chr_id];
   const int lane = pg_lane();
   if (e > chrom.len) e = chrom.len;
   if (s >= e) return;
   // first character of each strand's character set
   const uint32_t q0lo = ws->q[se.chp.arr][0][0], q0hi = ws->q[se.chp.arr][1][0];
   const uint32_t q0n = ws->q[se.chp.arr][2][0], q0x = ws->q[se.chp.arr][3][0];
   if ((q0n | q0x) & 1u) return;            // first char N (or non-ACGTN): no seeds (pindel.cpp:2281, farend_searcher.cpp:54)
   const uint32_t c0lo = (q0lo ^ se.chp.inv) & 1u, c0hi = (q0hi ^ se.chp.inv) & 1u;
   const uint32_t c0lom = c0lo ? ~0u : 0u, c0him = c0hi ? ~0u : 0u;
   bool read_special = false;
   {
      uint32_t sp = 0;
      if (lane < LW) sp = ws->q[se.chp.arr][2][lane] | ws->q[se.chp.arr][3][lane];
      read_special = pg_any(sp != 0);
   }
   uint32_t valid0 = pg_valid_mask<LW, TL>(se.L, 0) & ~1u;
   PG_PIN(valid0);                              // keep it in a register: it is otherwise re-derived from L in every pass
   const bool far = se.far;
   const bool single_fwd = se.d > 0;
   // windows of several tiles are cut at multiples of the tile size, so that all tiles but the first are whole words
   const bool aligned_tiles = e - s > PG_TILE_POS;
   bool thr_valid = false;
   for (uint32_t ts = s, te; ts < e; ts = te) {
      te = aligned_tiles ? (ts & ~(uint32_t)(PG_TILE_POS - 1)) + PG_TILE_POS : e;
      if (te > e) te = e;
      const int64_t w0 = (((int64_t)(ts >> 5)) - 1) & ~(int64_t)3;
      const int64_t w1 = (int64_t)((te - 1) >> 5) + 2;          // exclusive
      const int nw = (int)(((w1 - w0) + 3) & ~(int64_t)3);
      pg_tile_load<LW, TL>(ws, chrom, w0, nw, phase);
      const uint32_t origin = (uint32_t)(w0 * 32);              // absolute position of tile word 0, bit 0 (w0 >= -4: taken modulo 2^32)
      bool filt = false;
      if (TL <= 16 && filter) {
         // the thresholds only move when a survivor was merged into the state
         if (!thr_valid) { pg_filter_thresholds<LW, TL>(cx, se.T); thr_valid = true; }
         const uint32_t fr = pg_filter_tile<LW, TL>(ws, ts - origin, te - origin, c0lom, c0him, se.T);
         if (fr & 1u) cx.any_seed = 1;
#ifndef PG_NO_TILE_SKIP
         if (!(fr & 2u)) continue;                               // nothing in this tile can change the state: next tile
#endif
         filt = true;
         thr_valid = false;
      }
      // does the tile hold any reference N?
      uint32_t nacc = 0;
#pragma unroll 1
      for (int i = lane; i < nw; i += 32) nacc |= ws->tile_n[i];
      const bool special = read_special || pg_any(nacc != 0);
      uint32_t ka_[TL], kb_[TL];     // lane-local top-2 keys of this tile, one pair per level
#pragma unroll
      for (int t = 0; t < TL; t++) { ka_[t] = 0; kb_[t] = 0; }
      // positions per lane in the seed-list phase: 32 (one word) for long tiles, 16 or 8 for the short windows that
      // dominate (close windows, inner far windows), so that all 32 lanes share the list building
      const uint32_t tlen = te - (ts & ~31u);
      const uint32_t gran = tlen <= 256u ? 8u : (tlen <= 512u ? 16u : 32u);
      const uint32_t gmask = gran == 32u ? ~0u : ((1u << gran) - 1u);
      // the seed queue is filled chunk by chunk; without the filter every chunk is processed on its own, with it the
      // (few) survivors of the whole tile share their passes
      uint32_t cs = ts & ~31u;
      uint32_t nq = 0;                                           // seeds waiting in ws->seedq (uniform)
      for (;;) {
         // ---- (1) seed list ----
         while (cs < te) {
            uint32_t plus = 0, minus = 0;
            const uint32_t p0 = cs + gran * (uint32_t)lane;      // first position owned by this lane
            if (p0 < te) {
               int wi = (int)((int64_t)(p0 >> 5) - w0);
               uint32_t sub = p0 & 31u;
               if (filt) {
                  plus = (ws->fsp[wi] >> sub) & gmask;
                  minus = (ws->fsm[wi] >> sub) & gmask;
               } else {
                  uint2 W = ws->tile[wi];
                  uint32_t notn = ~ws->tile_n[wi] >> sub;
                  uint32_t inr = gmask;                           // owned positions inside [ts, te)
                  if (p0 < ts) inr &= (ts - p0 >= 32u) ? 0u : (~0u << (ts - p0));
                  if (te - p0 < gran) inr &= (1u << (te - p0)) - 1u;
                  uint32_t dl = (W.x ^ c0lom) >> sub, dh = (W.y ^ c0him) >> sub;   // 0 bits where the base equals read[0]
                  plus = ~(dl | dh) & notn & inr;
                  if (far) minus = (dl & dh) & notn & inr;       // both planes differ: the complement of read[0]
                  else if (!single_fwd) { minus = plus; plus = 0; } // close end growing left: every seed is a "minus" seed
               }
            }
            uint32_t cnt = (uint32_t)(pg_popc(plus) + pg_popc(minus));
            uint32_t incl = cnt;                                  // inclusive prefix sum over the lanes
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
               uint32_t v = pg_shfl_up(incl, d);
               if (lane >= d) incl += v;
            }
            const uint32_t nseeds = pg_shfl(incl, 31);
            if (nq + nseeds > PG_CHUNK_POS) break;                // no room: process the queue first, then this chunk again
            if (nseeds) {
               cx.any_seed = 1;
               uint32_t o = nq + incl - cnt;
               uint32_t both = plus | minus;
               uint32_t rel = p0 - origin;
               while (both) {
                  uint32_t bit = (uint32_t)pg_ffs(both) - 1u;
                  both &= both - 1u;
                  ws->seedq[o++] = (uint16_t)((rel + bit) | (((minus >> bit) & 1u) << 15));
               }
               nq += nseeds;
            }
            cs += 32u * gran;
            if (!filt && nq) break;
         }
         if (nq == 0) break;                                      // cs >= te and nothing queued
         const uint32_t nseeds = nq;
         nq = 0;
         pg_syncwarp();
         // ---- (2) 32 seeds per pass ----
         for (uint32_t q0 = 0; q0 < nseeds; q0 += 32) {
            uint32_t qi = q0 + (uint32_t)lane;
            bool have = qi < nseeds;
            uint32_t ent = have ? (uint32_t)ws->seedq[qi] : 0u;
            uint32_t off = ent & 0x7fffu;
            bool bwd = (ent >> 15) != 0;
            // the 32 reference bases the seed grows over: [off, off+32) rightwards, [off-31, off] (then reversed) leftwards
            uint32_t first = bwd ? off - 31u : off;
            uint32_t shift = first & 31u;
            int wi = (int)(first >> 5);
            uint32_t inv = (bwd && far) ? ~se.chp.inv : se.chp.inv;
            uint2 A = ws->tile[wi], B = ws->tile[wi + 1];
            uint32_t xlo = pg_fsr(A.x, B.x, shift), xhi = pg_fsr(A.y, B.y, shift);
            if (bwd) { xlo = pg_brev(xlo); xhi = pg_brev(xhi); }
            uint32_t mism = (xlo ^ q0lo ^ inv) | (xhi ^ q0hi ^ inv);
            if (special) {
               uint32_t xn = pg_fsr(ws->tile_n[wi], ws->tile_n[wi + 1], shift);
               if (bwd) xn = pg_brev(xn);
               mism = (mism & ~q0n) | q0x | xn;
            }
            uint32_t m0 = mism & valid0;
            // "hot" seeds are decided inside their first 32 bases: all TL tracked levels get their mismatch index
            // from this one word (levels >= T carry values nobody reads); the rest are finished cooperatively
            bool hot = have && pg_popc(m0) >= TL;
            bool slow = have && !hot;
            if (hot) {
               // the mismatch indices upwards = the set bits from the lowest.  The walk itself is two ALU operations per level
               // (clear the lowest set bit); the bit's index (bfind, the slow pipe) hangs off that chain instead of being part
               // of it -- walking a bit-reversed word with clz put FLO + shift + and-not between consecutive levels
               uint32_t mm = m0;
#pragma unroll
               for (int t = 0; t < TL; t++) {
                  const uint32_t next = mm & (mm - 1u);
                  const uint32_t D = pg_bfind(mm ^ next);
                  mm = next;
                  pg_key_update(ka_[t], kb_[t], D * (1u << PG_IDX_BITS) + off);
               }
            }
            uint32_t slowmask = pg_ballot(slow);
            if (slowmask) {
               if (slow) ws->queue[cx.qn + (uint32_t)pg_popc(slowmask & ((1u << lane) - 1u))] = origin + off;
               cx.qn += (uint32_t)pg_popc(slowmask);
            }
            bool last_pass = q0 + 32u >= nseeds;
            if (cx.qn > PG_QCAP - 32 || (last_pass && cx.qn)) pg_flush_queue<LW, TL>(cx, se, cx.chroms + chr_id, chr_id, c0lo, c0hi);
         }
         pg_syncwarp();
      }
      // fold the tile's lane-local top-2 into the warp-level state of lane t (two redux.max per level; the merge
      // itself is written once, outside the unrolled loop, to keep the code small)
      uint32_t my_ka = 0, my_kb = 0;
#pragma unroll
      for (int t = 0; t < TL; t++) {
         uint32_t ka = pg_redux_max(ka_[t]);
         uint32_t mine = (ka_[t] == ka) ? kb_[t] : ka_[t];
         uint32_t kb = pg_redux_max(mine);
         if (lane == t) { my_ka = ka; my_kb = kb; }
      }
      if (my_ka != 0) {
         uint32_t off = my_ka & PG_IDX_MASK;
         uint32_t minus = 0;
         if (far) {
            uint2 wv = ws->tile[off >> 5];
            uint32_t rlo = (wv.x >> (off & 31)) & 1u, rhi = (wv.y >> (off & 31)) & 1u;
            minus = (rlo != c0lo || rhi != c0hi) ? 1u : 0u;
         }
         pg_level_merge(cx.lvl, my_ka >> PG_IDX_BITS, origin + off, ((uint32_t)chr_id << 1) | minus);
         if (my_kb != 0) pg_level_merge(cx.lvl, my_kb >> PG_IDX_BITS, 0xffffffffu, 0);
      }
      pg_syncwarp();
   }
}

// Result of one search, runs written to `dst`.
struct PgResult {
   uint32_t n_runs;
   uint32_t n_ups;
   uint32_t max_len;       // LengthStr of the last UniquePoint (SortedUniquePoints::MaxLen)
   uint32_t last_abs_loc;  // AbsLoc of the last UniquePoint
   uint32_t seeds;         // any seed at all (NumberOfHits > 0)
};

// The lane-per-length replay of CheckLeft/Right_Close / CheckBoth + CheckMismatches.
template <int LW, int TL>
PG_DEV PgResult pg_walk_emit(PgWarpCtx<LW, TL> &cx, const PgSearch &se, PgRun *dst, uint32_t dst_cap)
{
   PgWarpSmem<LW, TL> *ws = cx.ws;
   const PgParams *prm = cx.prm;
   int lane = pg_lane();
   PgResult res;
   res.n_runs = 0; res.n_ups = 0; res.max_len = 0; res.last_abs_loc = 0; res.seeds = cx.any_seed ? 1u : 0u;
   // levels >= T read as "no candidate" (A = B = 0)
   const bool mine = lane < se.T;
   if (lane <= TL) {
      ws->A[lane] = mine ? cx.lvl.A : 0u; ws->B[lane] = mine ? cx.lvl.B : 0u;
      ws->AM[lane] = lane <= se.MAX ? cx.lvl.A : 0xffffffffu;      // A_t of the levels that count towards m
      ws->P[lane] = cx.lvl.pos; ws->META[lane] = cx.lvl.meta;
   }
   // Full-length masks of the seeds that can own a UniquePoint: one per level <= MAX whose best seed is long enough
   // (need), shared between consecutive levels owned by the same seed (alias).  Lane t decides for level t.
   {
      const uint32_t pA = pg_shfl_up(cx.lvl.A, 1), pP = pg_shfl_up(cx.lvl.pos, 1), pM = pg_shfl_up(cx.lvl.meta, 1);
      const bool need = lane <= se.MAX && cx.lvl.A >= (uint32_t)(se.bp_start + lane);
      const bool alias = need && lane > 0 && pA >= (uint32_t)(se.bp_start + lane - 1) && pP == cx.lvl.pos && pM == cx.lvl.meta;
      uint32_t todo = pg_ballot(need && !alias);
      // the level whose masks level t uses: the nearest computed level at or below t
      if (lane <= TL) ws->ALIAS[lane] = need ? (uint32_t)(31 - pg_clz(todo & ((2u << lane) - 1u))) : (uint32_t)lane;
      pg_syncwarp();
      while (todo) {
         const int t = pg_ffs(todo) - 1;
         todo &= todo - 1u;
         const uint32_t meta = ws->META[t];
         const int d = se.far ? ((meta & 1u) ? -1 : 1) : se.d;
         const uint32_t inv = (se.far && (meta & 1u)) ? ~se.chp.inv : se.chp.inv;
         const uint32_t h = pg_full_masks<LW, TL>(ws, cx.chroms + (meta >> 1), se.chp.arr, inv, ws->P[t], d, se.L, t);
         if (lane == 0) ws->H[t] = h;
      }
      pg_syncwarp();
   }
   const int L = se.L, MAX = se.MAX, ADD = prm->add_mm, MP = prm->min_perfect;
   const uint32_t thr = prm->thr[L];
   bool dead = false;          // uniform: an earlier length aborted the growth
   // open run carried across rounds (uniform)
   uint32_t run_len0 = 0, run_cnt = 0, run_m = 0;
   int prev_valid = 0; uint32_t prev_m = 0;
   // A_t is non-decreasing in t: no length beyond A_MAX can pass the in-range test (m would be MAX+1 > g[len])
   int len_hi = (int)ws->A[MAX];
   if (len_hi > L - 1) len_hi = L - 1;
#pragma unroll 1
   for (int r0 = se.bp_start; r0 <= len_hi && !dead; r0 += 32) {
      int len = r0 + lane;
      bool inrange = len <= L - 1;
      int m = 0;
#pragma unroll
      for (int t = 0; t < TL; t++) m += (ws->AM[t] < (uint32_t)len) ? 1 : 0;
      PG_PIN(m);               // one register, computed once (the compiler otherwise re-derives it at every use)
      bool abort_here = inrange && ((uint32_t)m > (uint32_t)prm->g[len]);       // searcher.cpp:167-169, pindel.cpp:2836-2838
      uint32_t abmask = pg_ballot(abort_here);
      uint32_t live = abmask ? ((1u << (pg_ffs(abmask) - 1)) - 1u) : 0xffffffffu;
      if (abmask) dead = true;
      bool cand = inrange && ((live >> lane) & 1u) && m <= MAX;
      bool valid = false;
      if (cand) {
         // level m holds exactly one candidate, it is long enough, and nothing competes within ADD levels
         cand = ws->B[m] < (uint32_t)len && len >= se.bp_start + m && ws->B[m + ADD] < (uint32_t)len &&
                (!se.far || ADD > 0);
      }
      if (cand) {
         uint32_t meta = ws->META[m];
         int d = se.far ? ((meta & 1u) ? -1 : 1) : se.d;
         int slot = (int)ws->ALIAS[m];
         bool ok = (d > 0) ? (len > MP) : (len >= MP);                           // searcher.cpp:346 / :360 substr clipping
#pragma unroll 1
         for (int j = 0; ok && j < MP; j++) {
            int bit = len - MP + j;
            ok = !((ws->S[slot][bit >> 5] >> (bit & 31)) & 1u);
         }
         valid = ok && ws->H[slot] >= thr;                                       // searcher.cpp:383
      }
      uint32_t vmask = pg_ballot(valid);
      // run boundaries: a run continues while consecutive lengths are valid with the same level
      uint32_t pm = pg_shfl((uint32_t)m, (lane + 31) & 31);
      int pv = (int)((vmask >> ((lane + 31) & 31)) & 1u);
      if (lane == 0) { pm = prev_m; pv = prev_valid; }
      bool starts = valid && !(pv && pm == (uint32_t)m);
      uint32_t smask = pg_ballot(starts);
      // uniform replay over the bits of vmask / smask
      uint32_t rem = vmask;
      while (rem) {
         int bpos = pg_ffs(rem) - 1;
         uint32_t mm = pg_shfl((uint32_t)m, bpos);
         if ((smask >> bpos) & 1u) {
            if (run_cnt) {       // close the open run
               if (lane == 0 && res.n_runs < dst_cap) {
                  uint32_t meta = ws->META[run_m];
                  int d = se.far ? ((meta & 1u) ? -1 : 1) : se.d;
                  PgRun rr;
                  rr.abs_loc0 = ws->P[run_m] + (uint32_t)(d * (int)(run_len0 - 1));
                  rr.len0 = (uint16_t)run_len0; rr.count = (uint16_t)run_cnt; rr.mm = (uint8_t)run_m;
                  rr.dirstrand = se.far ? ((meta & 1u) ? 3u : 0u) : (se.d > 0 ? 2u : 1u);
                  rr.chr = (int16_t)(meta >> 1);
                  dst[res.n_runs] = rr;
               }
               res.n_runs++;
            }
            run_len0 = (uint32_t)(r0 + bpos); run_cnt = 0; run_m = mm;
         }
         // extend over the consecutive valid non-start bits following bpos
         uint32_t after = (bpos == 31) ? 0u : (~0u << (bpos + 1));
         uint32_t stop = (~vmask | smask) & after;          // first bit after bpos that is invalid or starts a new run
         int endb = stop ? (pg_ffs(stop) - 1) : 32;
         run_cnt += (uint32_t)(endb - bpos);
         res.n_ups += (uint32_t)(endb - bpos);
         res.max_len = (uint32_t)(r0 + endb - 1);
         rem &= (endb >= 32) ? 0u : (~0u << endb);
      }
      prev_valid = (int)((vmask >> 31) & 1u);
      prev_m = pg_shfl((uint32_t)m, 31);
      if (!prev_valid && run_cnt) {
         // a gap ends the run only logically; keep it open until the next start closes it (no merging across gaps:
         // the next valid length is flagged as a start because pv == 0)
      }
   }
   if (run_cnt) {
      if (lane == 0 && res.n_runs < dst_cap) {
         uint32_t meta = ws->META[run_m];
         int d = se.far ? ((meta & 1u) ? -1 : 1) : se.d;
         PgRun rr;
         rr.abs_loc0 = ws->P[run_m] + (uint32_t)(d * (int)(run_len0 - 1));
         rr.len0 = (uint16_t)run_len0; rr.count = (uint16_t)run_cnt; rr.mm = (uint8_t)run_m;
         rr.dirstrand = se.far ? ((meta & 1u) ? 3u : 0u) : (se.d > 0 ? 2u : 1u);
         rr.chr = (int16_t)(meta >> 1);
         dst[res.n_runs] = rr;
      }
      res.n_runs++;
   }
   if (res.n_ups) {
      uint32_t meta = ws->META[run_m];
      int d = se.far ? ((meta & 1u) ? -1 : 1) : se.d;
      res.last_abs_loc = ws->P[run_m] + (uint32_t)(d * (int)(res.max_len - 1));
   }
   pg_syncwarp();
   return res;
}

template <int LW, int TL>
PG_DEV void pg_search_begin(PgWarpCtx<LW, TL> &cx)
{
   cx.lvl.A = 0; cx.lvl.B = 0; cx.lvl.pos = 0; cx.lvl.meta = 0;
   cx.qn = 0; cx.any_seed = 0;
}

// Pack the read (ASCII) into bit-planes; returns the stripped length (setUnmatchedSeq, pindel.cpp:142-169).
template <int LW, int TL>
PG_COLD int pg_pack_read(PgWarpSmem<LW, TL> *ws, const uint8_t *seq, int len)
{
   int lane = pg_lane();
   if (lane < 8 * LW) (&ws->q[0][0][0])[lane] = 0u;         // 8 * LW plane words (<= 128)
   if (lane + 32 < 8 * LW) (&ws->q[0][0][0])[lane + 32] = 0u;
   if (lane + 64 < 8 * LW) (&ws->q[0][0][0])[lane + 64] = 0u;
   if (lane + 96 < 8 * LW) (&ws->q[0][0][0])[lane + 96] = 0u;
   pg_syncwarp();
   // one pass over the characters: the forward planes, and L = the length without trailing non-alphanumerics
   int L = 0;
   const int nraw = (len + 31) >> 5;
#pragma unroll 1
   for (int c = 0; c < nraw; c++) {
      const int i = 32 * c + lane;
      const uint8_t ch = (i < len) ? pg_ldg8(seq + i) : (uint8_t)0;
      const bool alnum = (ch >= '0' && ch <= '9') || (ch >= 'A' && ch <= 'Z') || (ch >= 'a' && ch <= 'z');
      const uint32_t mk = pg_ballot(alnum);
      if (mk) L = 32 * c + 32 - pg_clz(mk);
      const uint32_t code = (ch == 'C') ? 1u : (ch == 'G') ? 2u : (ch == 'T') ? 3u : 0u;
      const bool isn = (ch == 'N');
      const bool isx = !(ch == 'A' || ch == 'C' || ch == 'G' || ch == 'T' || ch == 'N');
      const uint32_t lo = pg_ballot((code & 1u) != 0), hi = pg_ballot((code & 2u) != 0);
      const uint32_t nn = pg_ballot(isn), xx = pg_ballot(isx);
      if (lane == 0) { ws->q[0][0][c] = lo; ws->q[0][1][c] = hi; ws->q[0][2][c] = nn; ws->q[0][3][c] = xx; }
   }
   pg_syncwarp();
   // cut the planes at L, then the reversed planes from the forward ones: with G = the whole LW-word string bit-reversed
   // (G word c = brev of forward word LW-1-c), reversed[i] = forward[L-1-i] = G[i + 32*LW - L]
   if (lane < LW) {
      const uint32_t vm = pg_valid_mask<LW, TL>(L, lane);
#pragma unroll
      for (int pl = 0; pl < 4; pl++) ws->q[0][pl][lane] &= vm;
   }
   pg_syncwarp();
   if (lane < LW) {
      const int sft = 32 * LW - L, c0 = lane + (sft >> 5);
      const uint32_t sb = (uint32_t)(sft & 31);
#pragma unroll
      for (int pl = 0; pl < 4; pl++) {
         const uint32_t g0 = c0 < LW ? pg_brev(ws->q[0][pl][LW - 1 - c0]) : 0u;
         const uint32_t g1 = c0 + 1 < LW ? pg_brev(ws->q[0][pl][LW - 2 - c0]) : 0u;
         ws->q[1][pl][lane] = pg_fsr(g0, g1, sb);
      }
   }
   pg_syncwarp();
   return L;
}

// setUnmatchedSeq(ReverseComplement(seq)) (pindel.cpp:2545): materialise RC(read) as the new forward read.
// Leading non-ACGTN characters become trailing NULs after the flip and are stripped; returns the new length.
template <int LW, int TL>
PG_COLD int pg_flip_read(PgWarpSmem<LW, TL> *ws, int L)
{
   int lane = pg_lane();
   // k = number of leading 'other' characters of the current read
   int k = 0;
   {
      bool done = false;
      for (int c = 0; c < LW && !done; c++) {
         uint32_t x = ws->q[0][3][c] & pg_valid_mask<LW, TL>(L, c);
         uint32_t nx = ~x & pg_valid_mask<LW, TL>(L, c);
         if (nx) { k += pg_ffs(nx) - 1; done = true; }
         else { int rem = L - 32 * c; k += rem >= 32 ? 32 : (rem > 0 ? rem : 0); if (rem <= 32) done = true; }
      }
   }
   int Ln = L - k;
   // new forward  F'[j] = comp(R[L-1-j]) = ~V[j]  (j < Ln);  new reversed V'[j] = comp(R[k+j]) = ~(F >> k)[j]
   uint32_t nf[4] = {0, 0, 0, 0}, nv[4] = {0, 0, 0, 0};
   if (lane < LW) {
      uint32_t vm = pg_valid_mask<LW, TL>(Ln, lane);
      int ws_ = lane + (k >> 5);
      uint32_t sh = (uint32_t)(k & 31);
      for (int pl = 0; pl < 4; pl++) {
         uint32_t v = ws->q[1][pl][lane];
         uint32_t f0 = (ws_ < LW) ? ws->q[0][pl][ws_] : 0u, f1 = (ws_ + 1 < LW) ? ws->q[0][pl][ws_ + 1] : 0u;
         uint32_t fs = pg_fsr(f0, f1, sh);
         if (pl < 2) { v = ~v; fs = ~fs; }
         nf[pl] = v & vm; nv[pl] = fs & vm;
      }
      // a base that is N or other carries code bits 0 in both planes before complement -> keep them 0
      nf[0] &= ~(nf[2] | nf[3]); nf[1] &= ~(nf[2] | nf[3]);
      nv[0] &= ~(nv[2] | nv[3]); nv[1] &= ~(nv[2] | nv[3]);
   }
   pg_syncwarp();
   if (lane < LW) {
      for (int pl = 0; pl < 4; pl++) { ws->q[0][pl][lane] = nf[pl]; ws->q[1][pl][lane] = nv[pl]; }
   }
   pg_syncwarp();
   return Ln;
}

// The whole per-read program: GetCloseEnd (pindel.cpp:2531-2605) then SearchFarEnd (pindel.cpp:1001-1074), written
// as one loop over "searches" so that the scan and the replay are instantiated exactly once.  What lives across searches
// (the close end's and the best far end's summaries, the previous window, the statistics) sits in the warp's shared memory
// (ws->rs: every lane writes and reads the same words), not in registers: the scan's hot loop needs all 64 of them.
template <int LW, int TL>
PG_DEV void pg_process_read(PgWarpCtx<LW, TL> &cx, const PgBatch &bt, int ri, PgRun *scratch, uint32_t &phase)
{
   PgWarpSmem<LW, TL> *ws = cx.ws;
   const PgParams *prm = cx.prm;
   int lane = pg_lane();
   const uint32_t cap = bt.scratch_cap;
   uint32_t o0 = bt.seq_off[ri], o1 = bt.seq_off[ri + 1];
   int rawlen = (int)(o1 - o0);
   PgReadOut ro;
   ro.run_off = 0; ro.n_close_runs = 0; ro.n_far_runs = 0; ro.len_after = 0; ro.flags = 0;
   if (rawlen > 32 * LW || rawlen >= PG_TABLE) {
      ro.flags = PG_F_TOOLONG; ro.len_after = (uint16_t)rawlen;
      if (lane == 0) bt.out[ri] = ro;
      return;
   }
   int L = pg_pack_read<LW, TL>(ws, bt.seq_blob + o0, rawlen);
   if ((int)prm->g[L] + prm->add_mm + 1 > TL) {      // more mismatch levels than this build tracks
      ro.flags = PG_F_TOOLONG; ro.len_after = (uint16_t)L;
      if (lane == 0) bt.out[ri] = ro;
      return;
   }
   PgReadState &rs = ws->rs;
   rs.chr_id = bt.chr_id ? bt.chr_id[ri] : 0;
   rs.n_regions = (bt.regions && bt.region_off) ? (bt.region_off[ri + 1] - bt.region_off[ri]) : 0u;
   rs.close_runs = 0; rs.close_ups = 0; rs.close_max_len = 0; rs.close_last = 0;
   rs.best_runs = 0; rs.best_ups = 0; rs.best_max_len = 0;
   rs.far_best = 1;            // scratch third that holds the best far end (the other of 1 / 2 receives the next one)
   rs.window_bases = 0;        // reference positions scanned for this read (bench statistics)
   rs.have_prev = 0; rs.prev_ws = 0; rs.prev_we = 0;
   rs.filter_ready = 0;        // ws->ftab describes this read
   uint32_t flips = 0;
   enum { ST_CLOSE = 0, ST_FAR_CLUSTER = 1, ST_FAR_WINDOW = 2, ST_DONE = 3 };
   int st = ST_DONE;
   int range_index = 0;        // GetCloseEnd: RangeIndex 0..1
   bool second_try = false;    // GetCloseEnd: the attempt after setUnmatchedSeq(ReverseComplement())
   int k = 1;                  // SearchFarEnd: rangeIndex 1..MaxRangeIndex+1
   uint32_t span = 64;         // START_SEARCH_SPAN
   if (bt.mode & 1) {
      st = ST_CLOSE;
   } else if ((bt.mode & 2) && bt.close_info) {
      const PgCloseInfo ci = bt.close_info[ri];
      rs.close_ups = ci.max_len ? 1u : 0u;
      rs.close_max_len = ci.max_len;
      rs.close_last = ci.last_abs_loc;
      if (ci.max_len && L > 0) st = rs.n_regions ? ST_FAR_CLUSTER : ST_FAR_WINDOW;
   }
   while (st != ST_DONE) {
      // ---- describe the search ----
      PgSearch se;
      se.L = L; se.MAX = prm->g[L]; se.T = se.MAX + prm->add_mm + 1;
      PgRegion two[2];
      PgRegion &one = two[0];
      const PgRegion *regs = two;
      int nreg = 1;
      bool skip = false;
      bool keep_state = false;     // far windows are nested: window k only scans the two rings around window k-1
      uint32_t alg_bases = 0;
      const int chr_id = rs.chr_id;
      if (st == ST_CLOSE) {
         // GetCloseEndInner, pindel.cpp:2250-2396
         se.far = false; se.bp_start = prm->min_close;
         const uint8_t md = bt.matched_d[ri];
         const uint32_t rp = bt.matched_rel_pos[ri];
         const int isz = bt.insert_size[ri];
         int64_t start, end;
         if (md == '+') {
            // CurrentReadSeq = ReverseComplement(read): chars[j] = comp(R[L-1-j]); seeded on its first char, grows right
            se.chp.arr = 1; se.chp.inv = ~0u; se.d = 1;
            start = (int64_t)rp + PG_SPACER - (int64_t)range_index * isz;
            end = start + (int64_t)(2 * range_index + 1) * isz;
         } else {
            // CurrentReadSeq = read, seeded on its last char, grows left: chars[j] = R[L-1-j]
            se.chp.arr = 1; se.chp.inv = 0u; se.d = -1;
            end = (int64_t)rp + PG_SPACER + (int64_t)range_index * isz;
            start = end - (int64_t)(2 * range_index + 1) * isz;
         }
         if (start < 0) start = 0;
         if (end > 0xffffffffll) end = 0xffffffffll;
         one.chr = chr_id; one.start = (uint32_t)start; one.end = (uint32_t)(end > start ? end : start);
         skip = !(md == '+' || md == '-') || L <= 0;
      } else {
         // SearchFarEndAtPos, farend_searcher.cpp:46-103
         se.far = true; se.bp_start = PG_FAR_BP_START; se.chp.arr = 0; se.chp.inv = 0u; se.d = 1;
         skip = (ws->q[0][2][0] & 1u) != 0;          // read[0] == 'N' -> returns at once (:54-56)
         if (st == ST_FAR_CLUSTER) { regs = bt.regions + bt.region_off[ri]; nreg = (int)rs.n_regions; }
         else {
            uint32_t center = rs.close_last, size = cx.chroms[chr_id].len;
            uint32_t wsx = (center > span + PG_SPACER) ? center - span : PG_SPACER;            // pindel.cpp:1033-1042
            uint32_t wex = (center + span + PG_SPACER < size) ? center + span : size - PG_SPACER;
            alg_bases = wex > wsx ? wex - wsx : 0u;
            one.chr = chr_id; one.start = wsx; one.end = wex;
            const uint32_t pws = rs.prev_ws, pwe = rs.prev_we, have_prev = rs.have_prev;
            pg_syncwarp();              // every lane has read the previous window before any lane replaces it
            if (have_prev && !skip && wsx <= pws && wex >= pwe && pwe >= pws) {
               // the reference searches the whole window again (pindel.cpp:1027-1029); the top-2 state of window k-1
               // is still exact for its positions, so only [wsx, prev_ws) and [prev_we, wex) are new
               keep_state = true;
               two[0].chr = chr_id; two[0].start = wsx; two[0].end = pws;
               two[1].chr = chr_id; two[1].start = pwe; two[1].end = wex;
               nreg = 2;
            }
            rs.have_prev = skip ? 0u : 1u; rs.prev_ws = wsx; rs.prev_we = wex;
         }
      }
      // ---- run it ----
      if (!keep_state) pg_search_begin<LW, TL>(cx);
      if (!skip) {
         for (int r = 0; r < nreg; r++) {
            PgRegion rg = regs[r];
            uint32_t s = rg.start, e = rg.end;
            if (se.far) {
               int64_t ss = (int32_t)rg.start, ee = (int32_t)rg.end;    // int Start/End, farend_searcher.cpp:65-69
               if (ss < 0) ss = ee - 1;
               if (ss < 0) ss = 0;
               if (ee < ss) ee = ss;
               s = (uint32_t)ss; e = (uint32_t)ee;
            }
            if (e > s) {
               // the outer rings of the far schedule go through the threshold filter (the state of the inner windows
               // supplies the thresholds); its step table is built once per read, here, from the final sequence
               const bool use_filter = TL <= 16 && se.far && keep_state && e - s >= prm->filter_min_pos;
               if (use_filter && !rs.filter_ready) {
                  pg_filter_prepare<LW, TL>(ws, se.chp.arr, se.chp.inv, L);
                  rs.filter_ready = 1;
               }
               pg_scan_window<LW, TL>(cx, se, rg.chr, s, e, phase, use_filter);
               if (st != ST_FAR_WINDOW) alg_bases += e - s;
            }
         }
         // far windows are accounted at the size the reference scans, rings or not
         const unsigned long long wb = rs.window_bases;
         pg_syncwarp();
         rs.window_bases = wb + alg_bases;
      }
      // close runs go to the first third of the warp's scratch, a new far end to the third that does not hold the best one
      PgRun *dst = scratch + (st == ST_CLOSE ? 0u : (3u - rs.far_best) * cap);
      PgResult res = pg_walk_emit<LW, TL>(cx, se, dst, cap);
      // ---- decide what comes next ----
      const uint32_t best_ml = rs.best_max_len, close_ml = rs.close_max_len, far_best = rs.far_best;
      pg_syncwarp();                    // reads of the shared state above, writes below
      if (st == ST_CLOSE) {
         rs.close_runs = res.n_runs; rs.close_ups = res.n_ups; rs.close_max_len = res.max_len; rs.close_last = res.last_abs_loc;
         if (res.n_ups) {
            st = ((bt.mode & 2) && L > 0) ? (rs.n_regions ? ST_FAR_CLUSTER : ST_FAR_WINDOW) : ST_DONE;
         } else if (!second_try) {
            L = pg_flip_read<LW, TL>(ws, L);         // Temp_One_Read.setUnmatchedSeq(ReverseComplement(..)), pindel.cpp:2545
            flips++;
            second_try = true;
         } else {
            second_try = false;
            if (++range_index >= 2) st = ST_DONE;    // MaxRange = 2
         }
      } else {
         const bool better = res.seeds && res.max_len >= best_ml;     // NewUPFarIsBetter, farend_searcher.cpp:30-44
         if (better) {
            rs.best_runs = res.n_runs; rs.best_ups = res.n_ups; rs.best_max_len = res.max_len;
            rs.far_best = 3u - far_best;
         }
         bool good = (better ? res.max_len : best_ml) + close_ml >= (uint32_t)L; // goodFarEndFound, pindel.cpp:480-483
         if (st == ST_FAR_CLUSTER) {
            st = good ? ST_DONE : ST_FAR_WINDOW;
         } else {
            span *= 4;
            if (good || ++k > prm->max_range_index + 1) st = ST_DONE;
         }
      }
   }
   // publish: one allocation in the run pool for close + far runs
   const uint32_t ncl = (bt.mode & 1) ? rs.close_runs : 0u, nfar = rs.best_runs;
   const PgRun *closeBuf = scratch, *farBest = scratch + rs.far_best * cap;
   uint32_t total = ncl + nfar;
   uint32_t off = 0;
   if (lane == 0 && total) off = pg_atomic_add(bt.counters + 1, total);
   off = pg_shfl(off, 0);
   ro.len_after = (uint16_t)L;
   ro.flags = (uint16_t)(((flips & 1u) ? PG_F_FLIPPED : 0u) | ((flips & 3u) << PG_F_NFLIPS_SHIFT));
   if (total && (off + total > bt.run_pool_cap || ncl > cap || nfar > cap)) {
      ro.flags |= PG_F_OVERFLOW;
      if (lane == 0) bt.counters[2] = 1;
   } else {
      ro.run_off = off; ro.n_close_runs = (uint16_t)ncl; ro.n_far_runs = (uint16_t)nfar;
      pg_syncwarp();
      for (uint32_t i = (uint32_t)lane; i < total; i += 32) bt.run_pool[off + i] = (i < ncl) ? closeBuf[i] : farBest[i - ncl];
   }
   if (lane == 0) {
      bt.out[ri] = ro;
      // statistics for the roofline arithmetic: counters[4..5] = window bases, [6..7] = UniquePoints (64-bit each)
      pg_atomic_add64((unsigned long long *)(bt.counters + 4), rs.window_bases);
      pg_atomic_add64((unsigned long long *)(bt.counters + 6), (unsigned long long)(((bt.mode & 1) ? rs.close_ups : 0u) + rs.best_ups));
   }
   pg_syncwarp();
}
