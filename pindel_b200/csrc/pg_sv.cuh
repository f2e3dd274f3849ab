// pindel_b200/csrc/pg_sv.cuh -- the SV-classification program: what Pindel's seven per-read drivers decide for one read
// (SURVEY.md section 8(f), row N1), written once and compiled twice: as the body of pg_sv_kernel (pindel_pg.cu, one
// thread per read) and, for the CPU-side unit tests, as plain C++ (tests/emul).
//
// The reference walks expanded UniquePoint vectors; here the lists stay run-length coded (PgRun, the search kernel's own
// output, so the two stages chain on the device without expansion).  Every driver is an instance of one pair
// enumerator: close entries in list order (or reversed), far *runs* in list order (or reversed) -- an arm's test pins
// either the far length (D, TD, INV: LengthStr sum == ReadLength) or the far position (SI: adjacent breakpoints), so a
// far run holds at most one candidate -- then the arm's field assignments, the breakpoint slide over the packed
// reference (GetRealStart4Deletion/Insertion, LeftMostTD, LeftMostINV), and the bin / region / box tail.
//
// A driver "fire" is reported as a PgSvRecord; the host replays records per driver (pg_sv_types.h).
//
// Reference: src/search_variant.cpp:49-266, src/searchdeletions.cpp:33-61, src/searchshortinsertions.cpp:33-61,
// src/search_deletions_nt.cpp:26-140, src/search_tandem_duplications.cpp:28-222, src/search_tandem_duplications_nt.cpp:27-131,
// src/search_inversions.cpp:29-318, src/search_inversions_nt.cpp:27-200, src/pindel.cpp:561-592,2094-2163.
#pragma once
#include "pg_sv_types.h"

#ifdef __CUDACC__
#define PGSV_HD __host__ __device__ __forceinline__
#define PGSV_FN __host__ __device__
#else
#define PGSV_HD inline
#define PGSV_FN inline
#endif

struct PgSvUP {
   int32_t len;     // LengthStr
   uint32_t loc;    // AbsLoc
   int32_t mm;      // Mismatches
   int32_t dir;     // 0 FORWARD, 1 BACKWARD
   int32_t strand;  // 0 SENSE, 1 ANTISENSE
};

PGSV_HD PgSvUP pg_sv_up(const PgRun &r, int j)
{
   PgSvUP u;
   u.len = (int32_t)r.len0 + j;
   u.dir = r.dirstrand & 1;
   u.strand = (r.dirstrand >> 1) & 1;
   u.loc = u.dir ? r.abs_loc0 - (uint32_t)j : r.abs_loc0 + (uint32_t)j;
   u.mm = r.mm;
   return u;
}

// CleanUniquePoints (src/pindel.cpp:2904-2941) keeps the points that extend the last one: same chromosome, direction,
// strand and the same AbsLoc -/+ LengthStr.  That value is constant inside a run, so whole runs stay or go.
PGSV_HD uint32_t pg_sv_term(const PgRun &r) { return (r.dirstrand & 1) ? r.abs_loc0 + r.len0 : r.abs_loc0 - r.len0; }
PGSV_HD bool pg_sv_close_kept(const PgRun &r, const PgRun &last)
{
   return r.chr == last.chr && r.dirstrand == last.dirstrand && pg_sv_term(r) == pg_sv_term(last);
}

// reference base at padded position pos: 0..3 = A C G T, 4 = N (anything else)
PGSV_HD int pg_sv_ref(const PgChrom &c, uint32_t pos)
{
   if (pos >= c.len) return 4; // (the reference would read past its string; callers never get here with real UniquePoints)
   const uint32_t w = pos >> 5, b = pos & 31;
   if ((c.nplane[w] >> b) & 1) return 4;
   const uint2 p = c.planes[w];
   return (int)(((p.x >> b) & 1) | (((p.y >> b) & 1) << 1));
}

PGSV_HD int pg_sv_code(uint8_t ch) { return ch == 'A' ? 0 : ch == 'C' ? 1 : ch == 'G' ? 2 : ch == 'T' ? 3 : ch == 'N' ? 4 : 5; }

struct PgSvRead {
   const uint8_t *seq;
   int32_t rl;              // ReadLength
   int32_t max_e;           // MAX_SNP_ERROR = g_maxMismatch[ReadLength]   (src/pindel.cpp:166)
   uint8_t matched_d;
   int32_t insert_size;
   const PgRun *cr;         // close runs
   int32_t ncr;
   const PgRun *fr;         // far runs
   int32_t nfr;
   uint32_t index;
   // SPLIT_READ fields the drivers assign (src/pindel.h:349-360)
   int32_t left, right;
   uint32_t bp_left, bp_right, indel;
   int32_t bp;              // short BP
   uint32_t nt_size;
   uint32_t nt_start, nt_len, nt_rot;
   uint32_t nt_src;
   bool used;
   uint32_t dirty;          // PG_SVF_* assigned since the last record
   uint32_t seq_no;
};

struct PgSvEnv {
   const PgSvParams *P;
   PgChrom chrom;
   PgSvRecord *records;
   uint32_t records_cap;
   uint32_t *counters;
   uint8_t *counts;         // optional [PG_SV_NDRIVERS * n_reads], driver-major: records of (driver, read); zeroed by the caller
   uint32_t n_reads;
};

PGSV_HD uint32_t pg_sv_alloc(uint32_t *ctr)
{
#if defined(__CUDA_ARCH__)
   return atomicAdd(ctr, 1u);
#else
   return (*ctr)++;
#endif
}

PGSV_FN void pg_sv_emit(const PgSvEnv &ev, PgSvRead &r, int driver, int outcome, uint32_t box)
{
   const uint32_t slot = pg_sv_alloc(&ev.counters[0]);
   if (slot < ev.records_cap) {
      PgSvRecord rec;
      rec.read = r.index;
      rec.driver = (uint8_t)driver;
      rec.outcome = (uint8_t)outcome;
      rec.fields = (uint8_t)r.dirty;
      rec.seq = (uint8_t)r.seq_no;
      rec.left = r.left;
      rec.right = r.right;
      rec.bp_left = r.bp_left;
      rec.bp_right = r.bp_right;
      rec.indel_size = r.indel;
      rec.bp = (int16_t)r.bp;
      rec.nt_size = (uint16_t)r.nt_size;
      rec.nt_start = (uint16_t)r.nt_start;
      rec.nt_len = (uint16_t)r.nt_len;
      rec.nt_rot = (uint16_t)r.nt_rot;
      rec.nt_src = (uint8_t)r.nt_src;
      rec.pad = 0;
      rec.box = box;
      ev.records[slot] = rec;
   } else {
      ev.counters[1] = 1u;
   }
   r.seq_no++;
   r.dirty = 0;
}

// character i of the current NT_str
PGSV_HD uint8_t pg_sv_nt_char(const PgSvRead &r, uint32_t i)
{
   uint32_t k = i + r.nt_rot;
   if (k >= r.nt_len) k -= r.nt_len;
   k += r.nt_start;
   if (r.nt_src == PG_NT_SEQ) return r.seq[k];
   const uint8_t ch = r.seq[(uint32_t)r.rl - 1 - k]; // ReverseComplement, Convert2RC4N (src/pindel.cpp:2037-2048)
   return ch == 'A' ? 'T' : ch == 'C' ? 'G' : ch == 'G' ? 'C' : ch == 'T' ? 'A' : ch == 'N' ? 'N' : 0;
}

PGSV_HD void pg_sv_set_nt(PgSvRead &r, uint32_t src, uint32_t start, uint32_t want) // std::string::substr(start, want)
{
   const uint32_t avail = (uint32_t)r.rl > start ? (uint32_t)r.rl - start : 0u;
   r.nt_len = want > avail ? avail : want;
   r.nt_src = r.nt_len ? src : PG_NT_EMPTY;
   r.nt_start = start;
   r.nt_rot = 0;
   r.dirty |= PG_SVF_NTSTR;
}

// GetRealStart4Deletion (src/pindel.cpp:2094-2116); only RealStart is used by the caller
PGSV_FN uint32_t pg_sv_slide_deletion(const PgChrom &c, uint32_t real_start, uint32_t real_end)
{
   uint32_t pos = real_start + PG_SPACER, end = real_end + PG_SPACER - 1;
   for (;;) {
      const int a = pg_sv_ref(c, pos);
      if (a == 4 || a != pg_sv_ref(c, end)) break;
      --pos;
      --end;
   }
   return pos - PG_SPACER;
}

// GetRealStart4Insertion (src/pindel.cpp:2131-2163): rotating the inserted string is a rotation offset here
PGSV_FN uint32_t pg_sv_slide_insertion(const PgChrom &c, PgSvRead &r, uint32_t real_end)
{
   uint32_t after = real_end + PG_SPACER;
   for (;;) {
      const int a = pg_sv_ref(c, after);
      if (a == 4 || a != pg_sv_code(pg_sv_nt_char(r, 0))) break;
      r.nt_rot = r.nt_rot + 1 == r.nt_len ? 0 : r.nt_rot + 1; // rotateForward
      after++;
   }
   uint32_t before = after - 1;
   for (;;) {
      const int a = pg_sv_ref(c, before);
      if (a == 4 || a != pg_sv_code(pg_sv_nt_char(r, r.nt_len - 1))) break;
      r.nt_rot = r.nt_rot == 0 ? r.nt_len - 1 : r.nt_rot - 1; // rotateBack
      before--;
   }
   return before - PG_SPACER;
}

// LeftMostTD (src/search_tandem_duplications.cpp:186-222); returns true when the read was "killed"
PGSV_FN bool pg_sv_left_most_td(const PgChrom &c, PgSvRead &r)
{
   uint32_t pos = r.bp_left + PG_SPACER, end = r.bp_right + PG_SPACER - 1;
   const uint32_t pos0 = pos;
   if (pos >= c.len || end >= c.len) {
      r.bp_left = 1;
      r.bp_right = 1;
      r.bp = 1;
      r.used = true;
      return true;
   }
   while (pos > 0 && pg_sv_ref(c, pos) == pg_sv_ref(c, end)) { // no 'N' test here: N == N keeps sliding
      --pos;
      --end;
   }
   int32_t diff = (int32_t)(pos0 - pos);
   if (diff > 0) {
      if (diff >= r.bp) diff = r.bp - 1;
      r.bp_left -= (uint32_t)diff;
      r.bp_right -= (uint32_t)diff;
      r.bp = (int16_t)(r.bp - diff);
   }
   return false;
}

// LeftMostINV (src/search_inversions.cpp:283-318)
PGSV_FN bool pg_sv_left_most_inv(const PgChrom &c, PgSvRead &r)
{
   uint32_t pos = r.bp_left + PG_SPACER + 1, end = r.bp_right + PG_SPACER - 1;
   const uint32_t pos0 = pos;
   if (c.len <= pos + PG_SPACER) {
      r.bp_left = 1;
      r.bp_right = 1;
      r.bp = 1;
      r.used = true;
      return true;
   }
   while (pos + 1 < c.len && end > 0) { // TheInput[PosIndex] == Convert2RC4N[TheInput[End]]: complement of A/C/G/T, N pairs with N
      const int a = pg_sv_ref(c, pos), b = pg_sv_ref(c, end);
      if (!(a == 4 ? b == 4 : (b != 4 && a == 3 - b))) break;
      ++pos;
      --end;
   }
   int32_t diff = (int16_t)(pos - pos0);
   if (diff > 0 && r.matched_d != '+') { // the '+' branch clamps DIFF and drops it
      if (diff + r.bp >= r.rl) diff = (int16_t)(r.rl - r.bp - 1);
      r.bp_left += (uint32_t)diff;
      r.bp_right -= (uint32_t)diff;
      r.bp = (int16_t)(r.bp + diff);
   }
   return false;
}

// readTransgressesBinBoundaries / saveReadForNextCycle / readInSpecifiedRegion / the box test (src/pindel.cpp:561-592,
// e.g. src/search_variant.cpp:147-166)
PGSV_FN void pg_sv_tail(const PgSvEnv &ev, PgSvRead &r, int driver, bool check_bin, int forced)
{
   const PgSvParams &P = *ev.P;
   int outcome = forced;   // PG_SV_BREAK from the caller, else decided here; PG_SV_NONE = nothing stored now
   uint32_t box = 0;
   if (outcome == PG_SV_NONE) {
      if (check_bin && r.bp_right > P.window_end - (uint32_t)(2 * r.insert_size)) outcome = PG_SV_FUTURE;
      else if (P.region_start_defined && r.bp_left + 1 < P.region_start) return;
      else if (P.region_end_defined && r.bp_left + 1 > P.region_end) return;
      else {
         box = r.bp_left / P.box_size;
         if (box >= P.num_boxes) return;
         outcome = PG_SV_BOX;
      }
   }
   pg_sv_emit(ev, r, driver, outcome, box);
   r.used = true;
}

// arms of the pair drivers
#define PG_ARM_D_PLUS 0
#define PG_ARM_D_MINUS 1
#define PG_ARM_SI_PLUS 2
#define PG_ARM_SI_MINUS 3
#define PG_ARM_TD_PLUS 4
#define PG_ARM_TD_MINUS 5
#define PG_ARM_INV0 6   // '+' anchor, far end downstream    (src/search_inversions.cpp:54-105)
#define PG_ARM_INV1 7   // '+' anchor inside the inverted block (:107-160)
#define PG_ARM_INV2 8   // '-' anchor                         (:166-214)
#define PG_ARM_INV3 9   // '-' anchor inside                  (:216-268)

// One passing pair: assign the fields, slide the breakpoint, decide.  Returns false when the far loop must stop (D/SI's
// "breakpoint beyond the chromosome" exit, src/search_variant.cpp:133-136).
PGSV_FN bool pg_sv_fire(const PgSvEnv &ev, PgSvRead &r, int driver, int arm, const PgSvUP &c, const PgSvUP &f)
{
   const uint32_t SP = PG_SPACER, rl = (uint32_t)r.rl, cl = (uint32_t)c.len, fl = (uint32_t)f.len;
   r.dirty |= PG_SVF_CORE;
   switch (arm) {
   case PG_ARM_D_PLUS:
   case PG_ARM_SI_PLUS:
      r.left = (int32_t)(c.loc - cl + 1);
      r.right = (int32_t)(f.loc + fl - 1);
      r.bp = c.len - 1;
      r.bp_left = c.loc - SP;
      r.bp_right = f.loc - SP;
      break;
   case PG_ARM_D_MINUS:
   case PG_ARM_SI_MINUS:
      r.left = (int32_t)(f.loc - fl + 1);
      r.right = (int32_t)(c.loc + cl - 1);
      r.bp = f.len - 1;
      r.bp_left = f.loc - SP;
      r.bp_right = c.loc - SP;
      break;
   case PG_ARM_TD_PLUS:
      r.right = (int32_t)(c.loc - cl + 1);
      r.left = (int32_t)(f.loc + fl - 1);
      r.bp = c.len - 1;
      r.indel = c.loc - f.loc + 1;
      r.bp_right = c.loc - SP;
      r.bp_left = f.loc - SP;
      break;
   case PG_ARM_TD_MINUS:
      r.right = (int32_t)(f.loc - fl + 1);
      r.left = (int32_t)(c.loc + cl - 1);
      r.bp = f.len - 1;
      r.indel = f.loc - c.loc + 1;
      r.bp_right = f.loc - SP;
      r.bp_left = c.loc - SP;
      break;
   case PG_ARM_INV0:
      r.left = (int32_t)((c.loc + 1) - cl);
      r.right = (int32_t)(f.loc - fl + rl);
      r.bp = c.len - 1;
      r.indel = f.loc - c.loc;
      r.bp_left = c.loc + 1 - SP;
      r.bp_right = f.loc - SP;
      break;
   case PG_ARM_INV1:
      r.right = (int32_t)(c.loc - cl + rl);
      r.left = (int32_t)(f.loc - fl + 1);
      r.bp = f.len - 1;
      r.indel = c.loc - f.loc;
      r.bp_right = c.loc - SP;
      r.bp_left = (f.loc + 1) - SP;
      break;
   case PG_ARM_INV2:
      r.left = (int32_t)(f.loc + fl - rl);
      r.right = (int32_t)(c.loc + cl - 1);
      r.bp = f.len - 1;
      r.indel = c.loc - f.loc;
      r.bp_left = f.loc - SP;
      r.bp_right = c.loc - 1 - SP;
      break;
   default: // PG_ARM_INV3
      r.right = (int32_t)(f.loc + fl - 1);
      r.left = (int32_t)(c.loc + cl - rl);
      r.bp = c.len - 1;
      r.indel = f.loc - c.loc;
      r.bp_left = c.loc - SP;
      r.bp_right = f.loc - 1 - SP;
      break;
   }
   bool check_bin = true;
   int forced = PG_SV_NONE;
   if (arm <= PG_ARM_SI_MINUS) {
      const bool plus = (arm == PG_ARM_D_PLUS || arm == PG_ARM_SI_PLUS);
      if (arm <= PG_ARM_D_MINUS) { // SearchDeletions::calculateIndelSize, getInsertedStr = ""
         r.indel = (uint32_t)((r.right - r.left) - (r.rl - 1));
         pg_sv_set_nt(r, PG_NT_EMPTY, 0, 0);
      } else { // SearchShortInsertions
         r.indel = (uint32_t)((r.rl - 1) - (r.right - r.left));
         pg_sv_set_nt(r, plus ? PG_NT_RC : PG_NT_SEQ, (uint32_t)(r.bp + 1), r.indel);
      }
      uint32_t real_left = r.bp_left;
      if (ev.chrom.len < real_left || ev.chrom.len < r.bp_right) {
         forced = PG_SV_BREAK;   // Used = true, the far loop stops (src/search_variant.cpp:133-136)
      } else {
         real_left = r.nt_len ? pg_sv_slide_insertion(ev.chrom, r, r.bp_right) : pg_sv_slide_deletion(ev.chrom, real_left, r.bp_right);
         int32_t diff = (int16_t)(r.bp_left - real_left);
         if ((r.bp - 1) < diff) diff = r.bp - 1;
         if (diff > 0) {
            r.bp = (int16_t)(r.bp - diff);
            r.bp_left -= (uint32_t)diff;
            r.bp_right -= (uint32_t)diff;
         }
      }
   } else if (arm <= PG_ARM_TD_MINUS) {
      if (r.bp_left == 0) return true; // src/search_tandem_duplications.cpp:88-90
      if (pg_sv_left_most_td(ev.chrom, r)) r.dirty |= PG_SVF_KILLED;
   } else {
      pg_sv_set_nt(r, PG_NT_EMPTY, 0, 0);
      r.nt_size = 0;
      r.dirty |= PG_SVF_NTSIZE;
      if (pg_sv_left_most_inv(ev.chrom, r)) r.dirty |= PG_SVF_KILLED;
      check_bin = arm <= PG_ARM_INV1; // the '-' arms have no bin test (src/search_inversions.cpp:199,251)
   }
   pg_sv_tail(ev, r, driver, check_bin, forced);
   return forced != PG_SV_BREAK;
}

// the test of one arm on a (close point, far point) pair
PGSV_HD bool pg_sv_pair_ok(const PgSvParams &P, const PgSvRead &r, int arm, const PgSvUP &c, const PgSvUP &f)
{
   const uint32_t mi = (uint32_t)P.min_inv;
   const bool sum_eq = f.len + c.len == r.rl;
   switch (arm) {
   case PG_ARM_D_PLUS: return sum_eq && f.loc > c.loc + 1;
   case PG_ARM_D_MINUS: return sum_eq && c.loc > f.loc + 1;
   case PG_ARM_SI_PLUS: return f.loc == c.loc + 1 && c.len + f.len < r.rl;
   case PG_ARM_SI_MINUS: return c.loc == f.loc + 1 && c.len + f.len < r.rl;
   case PG_ARM_TD_PLUS: return sum_eq && f.loc + (uint32_t)f.len < c.loc && f.loc + (uint32_t)c.len < c.loc;
   case PG_ARM_TD_MINUS: return sum_eq && c.loc + (uint32_t)c.len < f.loc && c.loc + (uint32_t)f.len < f.loc;
   case PG_ARM_INV0: return sum_eq && f.loc > c.loc + mi;
   case PG_ARM_INV1: return sum_eq && f.loc + mi < c.loc;
   case PG_ARM_INV2: return sum_eq && c.loc > f.loc + mi;
   default: return sum_eq && c.loc + mi < f.loc;
   }
}

// The pair enumerator shared by D, SI, TD and INV.  The reference nests four loops -- mismatch budget e, close points,
// far points, with the arm's test innermost -- and fires every passing pair until one of them takes the read.  Two facts
// make that cheap on run-length coded lists:
//   * inside a (close run, far run) pair the arm's test pins the far member: j = K + m * idx is linear in the close
//     member's index idx (length sum == ReadLength for D/TD/INV, adjacent positions for SI), so the close members that
//     have a far partner at all form one interval of the close run;
//   * a fire does not depend on e or on earlier fires, and a pair with c.mm + f.mm = s passes at every e >= s.  Firing
//     each pair once, at e == s, in the reference's order finds the same taking pair; when nothing takes the read the
//     fields keep the values of the reference's very last fire, which is the last pair in (close, far) order -- fired
//     once more at the end if it was not the last one fired.
struct PgSvSpan {
   int32_t lo, hi;   // close-run members, in iteration order, that have a partner in the far run (lo > hi: none)
   int32_t k, m;     // partner of close member idx: j = k + m * idx
};

PGSV_HD PgSvSpan pg_sv_span(const PgSvRead &r, int arm, const PgRun &crun, const PgRun &frun, bool close_desc, int far_dir)
{
   PgSvSpan s;
   int64_t K;
   if (arm == PG_ARM_SI_PLUS || arm == PG_ARM_SI_MINUS) {
      // want = c.loc + 1 (PLUS) or c.loc - 1 (MINUS), c.loc = abs_loc0 +/- idx; j = +/-(want - far abs_loc0), 32-bit wrap
      const uint32_t want0 = arm == PG_ARM_SI_PLUS ? crun.abs_loc0 + 1u : crun.abs_loc0 - 1u;
      K = (int64_t)(int32_t)(far_dir ? frun.abs_loc0 - want0 : want0 - frun.abs_loc0);
      const int cs = (crun.dirstrand & 1) ? -1 : 1;
      s.m = far_dir ? -cs : cs;
   } else {
      K = (int64_t)r.rl - (int64_t)crun.len0 - (int64_t)frun.len0;
      s.m = -1;
   }
   const int64_t fc = frun.count, cc = crun.count;
   int64_t a = s.m < 0 ? K - fc + 1 : -K, b = s.m < 0 ? K : fc - 1 - K;   // 0 <= K + m * idx <= fc - 1
   if (a < 0) a = 0;
   if (b > cc - 1) b = cc - 1;
   if (K > 0x7FFF0000ll || K < -0x7FFF0000ll) b = a - 1;   // the 32-bit value wraps nowhere near a run member
   s.k = (int32_t)K;
   if (close_desc) {
      s.lo = (int32_t)(cc - 1 - b);
      s.hi = (int32_t)(cc - 1 - a);
   } else {
      s.lo = (int32_t)a;
      s.hi = (int32_t)b;
   }
   return s;
}

PGSV_FN void pg_sv_pairs(const PgSvEnv &ev, PgSvRead &r, int driver, int arm, bool close_desc, bool far_desc, int far_dir)
{
   const PgRun &last_close = r.cr[r.ncr - 1];
   // the pair fired last, and the last pair in (close, far) order among those fired
   int32_t lf_ci = -1, lf_cj = -1, lf_fi = -1, mx_ci = -1, mx_cj = -1, mx_fi = -1;
   // passes e = 0 .. max_e fire the pairs with c.mm + f.mm == e; the extra pass replays the last pair in order (one call
   // site for the fire keeps the device code small: the arms share every instruction of it)
   for (int e = 0; e <= r.max_e + 1 && !r.used; e++) {
      const bool replay = e > r.max_e;
      if (replay && (mx_ci < 0 || (mx_ci == lf_ci && mx_cj == lf_cj && mx_fi == lf_fi))) break;
      for (int ci = replay ? mx_ci : 0; ci < (replay ? mx_ci + 1 : r.ncr) && !r.used; ci++) {
         const PgRun &crun = r.cr[close_desc ? r.ncr - 1 - ci : ci];
         if (!replay && ((int)crun.mm > e || !pg_sv_close_kept(crun, last_close))) continue;
         const int fmm = e - (int)crun.mm;   // pairs that pass for the first time at this e
         for (int cur = replay ? mx_cj : 0; cur < (replay ? mx_cj + 1 : (int)crun.count) && !r.used;) {
            int next = replay ? cur : 0x7FFFFFFF;
            for (int fi = 0; fi < r.nfr && !replay; fi++) {
               const PgRun &frun = r.fr[far_desc ? r.nfr - 1 - fi : fi];
               if ((int)frun.mm != fmm || (frun.dirstrand & 1) != far_dir) continue;
               const PgSvSpan sp = pg_sv_span(r, arm, crun, frun, close_desc, far_dir);
               if (sp.hi < cur || sp.hi < sp.lo) continue;
               const int at = sp.lo > cur ? sp.lo : cur;
               if (at < next) next = at;
            }
            if (next == 0x7FFFFFFF) break;
            const int idx = close_desc ? (int)crun.count - 1 - next : next;
            const PgSvUP c = pg_sv_up(crun, idx);
            for (int fi = replay ? mx_fi : 0; fi < (replay ? mx_fi + 1 : r.nfr) && !r.used; fi++) {
               const PgRun &frun = r.fr[far_desc ? r.nfr - 1 - fi : fi];
               if (!replay && ((int)frun.mm != fmm || (frun.dirstrand & 1) != far_dir)) continue;
               const PgSvSpan sp = pg_sv_span(r, arm, crun, frun, close_desc, far_dir);
               if (next < sp.lo || next > sp.hi) continue;
               const PgSvUP f = pg_sv_up(frun, sp.k + sp.m * idx);
               if (!pg_sv_pair_ok(*ev.P, r, arm, c, f)) continue;
               lf_ci = ci;
               lf_cj = next;
               lf_fi = fi;
               if (ci > mx_ci || (ci == mx_ci && (next > mx_cj || (next == mx_cj && fi > mx_fi)))) {
                  mx_ci = ci;
                  mx_cj = next;
                  mx_fi = fi;
               }
               if (!pg_sv_fire(ev, r, driver, arm, c, f)) break;
            }
            cur = next + 1;
         }
      }
   }
}

PGSV_HD PgSvUP pg_sv_first(const PgRun *runs) { return pg_sv_up(runs[0], 0); }
PGSV_HD PgSvUP pg_sv_last(const PgRun *runs, int n) { return pg_sv_up(runs[n - 1], (int)runs[n - 1].count - 1); }
// UP_Close[0] after CleanUniquePoints: the first run that survives the filter
PGSV_HD PgSvUP pg_sv_first_close(const PgSvRead &r)
{
   int k = 0;
   while (!pg_sv_close_kept(r.cr[k], r.cr[r.ncr - 1])) k++;
   return pg_sv_up(r.cr[k], 0);
}

// the shared shape of DI, TD_NT and INV_NT: one test on (last close point, last far point)
PGSV_FN void pg_sv_nt_driver(const PgSvEnv &ev, PgSvRead &r, int driver)
{
   const PgSvParams &P = *ev.P;
   const PgSvUP c = pg_sv_last(r.cr, r.ncr), f = pg_sv_last(r.fr, r.nfr);
   const uint32_t SP = PG_SPACER, rl = (uint32_t)r.rl, cl = (uint32_t)c.len, fl = (uint32_t)f.len, mi = (uint32_t)P.min_inv;
   const int32_t sum = f.len + c.len;
   const bool plus = r.matched_d == '+';
   if (!plus && r.matched_d != '-') return;
   if (driver == PG_SV_TD_NT && sum >= r.rl) return;
   if (f.mm + c.mm > (int32_t)(int16_t)P.nt_thr[sum < 1032 ? sum : 1031]) return;
   if (driver == PG_SV_INV_NT) {
      const PgSvUP c0 = pg_sv_first_close(r), f0 = pg_sv_first(r.fr);
      if (!(c0.strand != f0.strand && c0.dir == f0.dir)) return;
      if (f.dir != (plus ? 0 : 1)) return;
   } else if (f.dir != (plus ? 1 : 0)) return;
   const uint32_t nts = (uint32_t)(uint16_t)(r.rl - f.len - c.len);
   // DI (src/search_deletions_nt.cpp:55-124) and TD_NT (src/search_tandem_duplications_nt.cpp:56-117) make one test;
   // INV_NT (src/search_inversions_nt.cpp:56-186) runs the two tests of the anchor's strand one after the other,
   // whatever the first one did.
   if (driver == PG_SV_DI) {
      if (!(sum < r.rl && sum >= P.min_matched && (plus ? f.loc > c.loc + 1 : c.loc > f.loc + 1))) return;
   } else if (driver == PG_SV_TD_NT) {
      if (!(sum > P.min_matched)) return;
      if (plus ? !(f.loc + fl < c.loc && f.loc + cl < c.loc) : !(c.loc + cl < f.loc && c.loc + fl < f.loc)) return;
   } else if (!(sum < r.rl && sum >= P.min_matched)) return;
   const int rounds = driver == PG_SV_INV_NT ? 2 : 1;
   for (int k = 0; k < rounds; k++) {
      if (driver == PG_SV_DI) {
         if (plus) {
            r.left = (int32_t)(c.loc - cl + 1);
            r.right = (int32_t)(f.loc + fl - 1);
            r.bp = c.len - 1;
         } else {
            r.left = (int32_t)(f.loc - fl + 1);
            r.right = (int32_t)(c.loc + cl - 1);
            r.bp = f.len - 1;
         }
         pg_sv_set_nt(r, plus ? PG_NT_RC : PG_NT_SEQ, (uint32_t)(r.bp + 1), nts);
         r.indel = (uint32_t)((r.right - r.left) + (int32_t)nts - (r.rl - 1));
         r.bp_left = (plus ? c.loc : f.loc) - SP;
         r.bp_right = (plus ? f.loc : c.loc) - SP;
      } else if (driver == PG_SV_TD_NT) {
         if (plus) {
            r.right = (int32_t)(c.loc - cl + 1);
            r.left = (int32_t)(f.loc + fl - 1);
            r.bp = c.len - 1;
            r.indel = c.loc - f.loc + 1;
         } else {
            r.right = (int32_t)(f.loc - fl + 1);
            r.left = (int32_t)(c.loc + cl - 1);
            r.bp = f.len - 1;
            r.indel = f.loc - c.loc + 1;
         }
         pg_sv_set_nt(r, plus ? PG_NT_RC : PG_NT_SEQ, (uint32_t)(r.bp + 1), nts);
         r.bp_right = (plus ? c.loc : f.loc) - SP;
         r.bp_left = (plus ? f.loc : c.loc) - SP;
      } else {
         const int arm = plus ? (k == 0 ? PG_ARM_INV0 : PG_ARM_INV1) : (k == 0 ? PG_ARM_INV2 : PG_ARM_INV3);
         bool ok;
         switch (arm) {
         case PG_ARM_INV0: ok = f.loc > c.loc + mi; break;
         case PG_ARM_INV1: ok = f.loc + mi < c.loc; break;
         case PG_ARM_INV2: ok = c.loc > f.loc + mi; break;
         default: ok = c.loc + mi < f.loc; break;
         }
         if (!ok) continue;
         switch (arm) {
         case PG_ARM_INV0:
            r.left = (int32_t)((c.loc + 1) - cl);
            r.right = (int32_t)(f.loc - fl + rl);
            r.bp = c.len - 1;
            r.indel = f.loc - c.loc;
            pg_sv_set_nt(r, PG_NT_RC, (uint32_t)(r.bp + 1), nts);
            r.bp_left = c.loc + 1 - SP;
            r.bp_right = f.loc - SP;
            break;
         case PG_ARM_INV1:
            r.right = (int32_t)(c.loc - cl + rl);
            r.left = (int32_t)(f.loc - fl + 1);
            r.bp = f.len - 1;
            r.indel = c.loc - f.loc;
            pg_sv_set_nt(r, PG_NT_SEQ, (uint32_t)(r.bp + 1), nts);
            r.bp_right = c.loc - SP;
            r.bp_left = (f.loc + 1) - SP;
            break;
         case PG_ARM_INV2:
            r.left = (int32_t)(f.loc + fl - rl);
            r.right = (int32_t)(c.loc + cl - 1);
            r.bp = f.len - 1;
            r.indel = c.loc - f.loc;
            pg_sv_set_nt(r, PG_NT_SEQ, (uint32_t)(r.bp + 1), nts);
            r.bp_left = f.loc - SP;
            r.bp_right = c.loc - 1 - SP;
            break;
         default:
            r.right = (int32_t)(f.loc + fl - 1);
            r.left = (int32_t)(c.loc + cl - rl);
            r.bp = c.len - 1;
            r.indel = f.loc - c.loc;
            pg_sv_set_nt(r, PG_NT_RC, (uint32_t)(r.bp + 1), nts);
            r.bp_left = c.loc - SP;
            r.bp_right = f.loc - 1 - SP;
            break;
         }
      }
      r.dirty |= PG_SVF_CORE | PG_SVF_NTSIZE;
      r.nt_size = nts;
      pg_sv_tail(ev, r, driver, true, PG_SV_NONE);
   }
}

// All selected drivers on one read, in SearchSVs' order (src/pindel.cpp:1140-1164).
PGSV_FN void pg_sv_classify_read(const PgSvEnv &ev, PgSvRead &r, bool other_chr)
{
   const PgSvParams &P = *ev.P;
   if (other_chr || r.nfr == 0 || r.ncr == 0) return; // every driver skips such reads
   const bool plus = r.matched_d == '+', minus = r.matched_d == '-';
   for (int d = 0; d < PG_SV_NDRIVERS; d++) {
      if (!((P.driver_mask >> d) & 1u) || r.used) continue;
      r.seq_no = 0;
      int arm = -1;   // pair drivers: which arm runs, if any
      switch (d) {
      case PG_SV_D: arm = plus ? PG_ARM_D_PLUS : minus ? PG_ARM_D_MINUS : -1; break;
      case PG_SV_SI: arm = plus ? PG_ARM_SI_PLUS : minus ? PG_ARM_SI_MINUS : -1; break;
      case PG_SV_TD: arm = plus ? PG_ARM_TD_PLUS : minus ? PG_ARM_TD_MINUS : -1; break;
      case PG_SV_INV: {
         const PgSvUP c0 = pg_sv_first_close(r), cl = pg_sv_last(r.cr, r.ncr), f0 = pg_sv_first(r.fr), fl = pg_sv_last(r.fr, r.nfr);
         if (!(c0.strand != f0.strand && c0.dir == f0.dir)) break;
         const uint32_t mi = (uint32_t)P.min_inv;
         if (plus) arm = f0.loc > cl.loc + mi ? PG_ARM_INV0 : fl.loc + mi < c0.loc ? PG_ARM_INV1 : -1;
         else if (minus) arm = cl.loc > f0.loc + mi ? PG_ARM_INV2 : c0.loc + mi < fl.loc ? PG_ARM_INV3 : -1;
         break;
      }
      default: pg_sv_nt_driver(ev, r, d); break;
      }
      if (arm >= 0) {
         // iteration order of the close and far lists, and the far direction the arm accepts:
         //   D+/SI+/TD+ (close up, far down, BACKWARD)   D-/SI- (down, down, FORWARD)   TD- (down, up, FORWARD)
         //   INV0 (down, up, FORWARD)   INV1 (up, down, FORWARD)   INV2 (down, up, BACKWARD)   INV3 (up, down, BACKWARD)
         const bool close_desc = arm == PG_ARM_D_MINUS || arm == PG_ARM_SI_MINUS || arm == PG_ARM_TD_MINUS || arm == PG_ARM_INV0 ||
                                 arm == PG_ARM_INV2;
         const bool far_desc = !(arm == PG_ARM_TD_MINUS || arm == PG_ARM_INV0 || arm == PG_ARM_INV2);
         const int far_dir = (arm == PG_ARM_D_PLUS || arm == PG_ARM_SI_PLUS || arm == PG_ARM_TD_PLUS || arm == PG_ARM_INV2 ||
                              arm == PG_ARM_INV3) ? 1 : 0;
         pg_sv_pairs(ev, r, d, arm, close_desc, far_desc, far_dir);
      }
      if (r.dirty) pg_sv_emit(ev, r, d, PG_SV_NONE, 0); // fields assigned, read not taken: the host still gets the values
      if (ev.counts && r.seq_no) ev.counts[(size_t)d * ev.n_reads + r.index] = (uint8_t)r.seq_no;
   }
}
