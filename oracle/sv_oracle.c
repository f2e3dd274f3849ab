/* oracle/sv_oracle.c -- TEST INFRASTRUCTURE ONLY (never linked into the product, never a fallback).
 *
 * CPU restatement of Pindel's seven per-read SV drivers, in the order
 * SearchSVs runs them (src/pindel.cpp:1140-1164):
 *
 *   D       SearchVariant::Search + SearchDeletions        src/search_variant.cpp:49-266, src/searchdeletions.cpp:33-61
 *   DI      searchIndels                                   src/search_deletions_nt.cpp:26-140
 *   TD      searchTandemDuplications, LeftMostTD           src/search_tandem_duplications.cpp:28-222
 *   TD_NT   searchTandemDuplicationsNT                     src/search_tandem_duplications_nt.cpp:27-131
 *   INV     searchInversions, LeftMostINV                  src/search_inversions.cpp:29-318
 *   INV_NT  searchInversionsNT                             src/search_inversions_nt.cpp:27-200
 *   SI      SearchVariant::Search + SearchShortInsertions  src/search_variant.cpp:49-266, src/searchshortinsertions.cpp:33-61
 *
 * Like the reference it works on expanded UniquePoint lists and mutates per-read fields; what the drivers hand to the
 * reporter (Vars[box], FutureReads_SR) is dumped together with the state of every read after each driver.
 * PINNED: tests/test_sv_oracle_pinning.py compares these dumps with the reference's own drivers (oracle/ref_hook.cpp,
 * pgref_classify) on synthetic batches and on the COLO-829 slice.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define SP 100000u /* g_SpacerBeforeAfter, src/pindel.h:122 */

typedef struct {
   int32_t len;      /* LengthStr (short) */
   uint32_t loc;     /* AbsLoc */
   int32_t dir;      /* 0 FORWARD('+'), 1 BACKWARD('-')   -- Direction */
   int32_t strand;   /* 0 SENSE, 1 ANTISENSE */
   int32_t mm;       /* Mismatches */
} UP;

typedef struct {
   const char *seq;
   int32_t rl;       /* ReadLength */
   char matched_d;
   int16_t insert_size;
   int used, otherchr;
   const UP *c;
   int nc;
   const UP *f;
   int nf;
   /* SPLIT_READ fields the drivers assign (src/pindel.h:349-360) */
   int32_t Left, Right;
   int16_t BP;
   uint32_t BPLeft, BPRight, IndelSize;
   uint16_t NT_size;
   char *NT_str;
   int nt_len;
} Read;

typedef struct {
   uint32_t driver_mask;
   uint32_t window_end, num_boxes, box_size;
   uint32_t region_start_defined, region_start, region_end_defined, region_end;
   int32_t min_inv, min_matched;
   double seq_error_rate;
   int32_t max_snp[1024]; /* g_maxMismatch */
} SvParams;

typedef struct {
   int32_t *v;
   size_t n, cap;
} Stream;

static void put(Stream *s, int32_t x)
{
   if (s->n == s->cap) {
      s->cap = s->cap ? s->cap * 2 : 1 << 16;
      s->v = (int32_t *)realloc(s->v, s->cap * sizeof(int32_t));
   }
   s->v[s->n++] = x;
}

static char rc_char(char c) /* Convert2RC4N, src/pindel.cpp:2037-2048: A<->T C<->G N->N, everything else NUL */
{
   switch (c) {
   case 'A': return 'T';
   case 'C': return 'G';
   case 'G': return 'C';
   case 'T': return 'A';
   case 'N': return 'N';
   default: return 0;
   }
}

static void set_nt(Read *r, int from_rc, int start, uint32_t want) /* std::string::substr(start, want) of the read or its RC */
{
   int avail = r->rl - start;
   if (avail < 0) avail = 0;
   int n = want > (uint32_t)avail ? avail : (int)want;
   r->nt_len = n;
   for (int i = 0; i < n; i++) r->NT_str[i] = from_rc ? rc_char(r->seq[r->rl - 1 - (start + i)]) : r->seq[start + i];
}

/* src/pindel.cpp:2094-2116; only RealStart is consumed by the callers */
static void real_start_deletion(const char *chr, uint32_t chr_len, uint32_t *real_start, uint32_t real_end)
{
   if (chr_len < *real_start || chr_len < real_end) return;
   uint32_t pos = *real_start + SP, end = real_end + SP - 1;
   while (chr[pos] == chr[end] && chr[pos] != 'N') {
      --pos;
      --end;
   }
   *real_start = pos - SP;
}

/* src/pindel.cpp:2131-2163: the inserted string is rotated while the breakpoint slides */
static void real_start_insertion(const char *chr, uint32_t chr_len, char *s, int n, uint32_t *real_start, uint32_t real_end)
{
   if (chr_len < *real_start || chr_len < real_end) return;
   uint32_t after = real_end + SP;
   while (chr[after] == s[0] && chr[after] != 'N') {
      char first = s[0];
      memmove(s, s + 1, (size_t)n - 1);
      s[n - 1] = first;
      after++;
   }
   uint32_t before = after - 1;
   while (chr[before] == s[n - 1] && chr[before] != 'N') {
      char last = s[n - 1];
      memmove(s + 1, s, (size_t)n - 1);
      s[0] = last;
      before--;
   }
   *real_start = before - SP;
}

/* src/search_tandem_duplications.cpp:186-222 */
static void left_most_td(const char *chr, uint32_t chr_len, Read *r)
{
   uint32_t pos = r->BPLeft + SP, pos0 = pos, end = r->BPRight + SP - 1;
   if (pos >= chr_len || end >= chr_len) {
      r->BPLeft = 1;
      r->BPRight = 1;
      r->BP = 1;
      r->used = 1;
      return;
   }
   while (pos > 0 && chr[pos] == chr[end]) { /* (the reference has no lower bound; position 0 is never reached on padded input) */
      --pos;
      --end;
   }
   int diff = (int)(pos0 - pos);
   if (diff > 0) {
      if (diff >= r->BP) diff = r->BP - 1;
      r->BPLeft -= (uint32_t)diff;
      r->BPRight -= (uint32_t)diff;
      r->BP = (int16_t)(r->BP - diff);
   }
}

/* src/search_inversions.cpp:283-318 */
static void left_most_inv(const char *chr, uint32_t chr_len, Read *r)
{
   uint32_t pos = r->BPLeft + SP + 1, pos0 = pos, end = r->BPRight + SP - 1;
   if (chr_len <= pos + SP) {
      r->BPLeft = 1;
      r->BPRight = 1;
      r->BP = 1;
      r->used = 1;
      return;
   }
   while (pos + 1 < chr_len && end > 0 && chr[pos] == rc_char(chr[end])) { /* (bounds added; never reached on padded input) */
      ++pos;
      --end;
   }
   int16_t diff = (int16_t)(pos - pos0);
   if (diff > 0) {
      if (r->matched_d == '+') {
         /* the reference clamps DIFF here and then does not use it */
      } else {
         if (diff + r->BP >= r->rl) diff = (int16_t)(r->rl - r->BP - 1);
         r->BPLeft += (uint32_t)diff;
         r->BPRight -= (uint32_t)diff;
         r->BP = (int16_t)(r->BP + diff);
      }
   }
}

typedef struct {
   Stream *out;
   Stream vars, future;
   int n_vars, n_future;
   const SvParams *P;
   const char *chr;
   uint32_t chr_len;
   const Read *base;
} Ctx;

static uint32_t fnv(const char *s, int n)
{
   uint32_t h = 2166136261u;
   for (int i = 0; i < n; i++) h = (h ^ (uint8_t)s[i]) * 16777619u;
   return h;
}

static void put_state(Stream *s, const Read *r)
{
   uint32_t head = 0, tail = 0;
   for (int i = 0; i < 4 && i < r->nt_len; i++) head |= (uint32_t)(uint8_t)r->NT_str[i] << (8 * i);
   for (int i = 0; i < 4 && i < r->nt_len; i++) tail |= (uint32_t)(uint8_t)r->NT_str[r->nt_len - 1 - i] << (8 * i);
   put(s, r->Left);
   put(s, r->Right);
   put(s, r->BP);
   put(s, (int32_t)r->BPLeft);
   put(s, (int32_t)r->BPRight);
   put(s, (int32_t)r->IndelSize);
   put(s, r->NT_size);
   put(s, r->used);
   put(s, r->nt_len);
   put(s, (int32_t)fnv(r->NT_str, r->nt_len));
   put(s, (int32_t)head);
   put(s, (int32_t)tail);
}

/* the tail every arm shares: readTransgressesBinBoundaries (src/pindel.cpp:561-564), saveReadForNextCycle (:587-592),
 * readInSpecifiedRegion (:567-585), the box test (e.g. src/search_variant.cpp:165-175) */
static void bin_region_box(Ctx *cx, Read *r, int check_bin)
{
   const SvParams *P = cx->P;
   if (check_bin && r->BPRight > P->window_end - (uint32_t)(2 * (int)r->insert_size)) {
      put(&cx->future, (int32_t)(r - cx->base));
      put_state(&cx->future, r);
      cx->n_future++;
      r->used = 1;
      return;
   }
   int ok = 1;
   if (P->region_start_defined && r->BPLeft + 1 < P->region_start) ok = 0;
   if (P->region_end_defined && r->BPLeft + 1 > P->region_end) ok = 0;
   if (!ok) return;
   uint32_t box = r->BPLeft / P->box_size;
   if (box < P->num_boxes) {
      put(&cx->vars, (int32_t)box);
      put(&cx->vars, (int32_t)(r - cx->base));
      cx->n_vars++;
      r->used = 1;
   }
}

/* ---- D and SI: SearchVariant::Search, src/search_variant.cpp:49-266 ---- */
static void variant_search(Ctx *cx, Read *r, int is_si)
{
   if (r->otherchr) return;
   if (r->used || r->nf == 0) return;
   const int max_e = cx->P->max_snp[r->rl];
   const int plus = r->matched_d == '+';
   if (!plus && r->matched_d != '-') return;
   for (int e = 0; e <= max_e; e++) {
      if (r->used) break;
      for (int ci = plus ? 0 : r->nc - 1; plus ? ci < r->nc : ci >= 0; ci += plus ? 1 : -1) {
         if (r->used) break;
         const UP *c = &r->c[ci];
         if (c->mm > e) continue;
         for (int fi = r->nf - 1; fi >= 0; fi--) {
            if (r->used) break;
            const UP *f = &r->f[fi];
            if (f->mm > e) continue;
            if (f->mm + c->mm > e) continue;
            if (f->dir != (plus ? 1 : 0)) continue;
            int pass;
            if (!is_si) { /* src/searchdeletions.cpp:33-43 */
               pass = f->len + c->len == r->rl && (plus ? f->loc > c->loc + 1 : c->loc > f->loc + 1);
            } else { /* src/searchshortinsertions.cpp:33-43 */
               pass = (plus ? f->loc == c->loc + 1 : c->loc == f->loc + 1) && c->len + f->len < r->rl;
            }
            if (!pass) continue;
            if (plus) {
               r->Left = (int32_t)(c->loc - (uint32_t)c->len + 1);
               r->Right = (int32_t)(f->loc + (uint32_t)f->len - 1);
               r->BP = (int16_t)(c->len - 1);
            } else {
               r->Left = (int32_t)(f->loc - (uint32_t)f->len + 1);
               r->Right = (int32_t)(c->loc + (uint32_t)c->len - 1);
               r->BP = (int16_t)(f->len - 1);
            }
            if (!is_si) {
               r->IndelSize = (uint32_t)((r->Right - r->Left) - (r->rl - 1));
               r->nt_len = 0;
            } else {
               r->IndelSize = (uint32_t)((r->rl - 1) - (r->Right - r->Left));
               set_nt(r, plus, r->BP + 1, r->IndelSize);
            }
            r->BPLeft = (plus ? c->loc : f->loc) - SP;
            r->BPRight = (plus ? f->loc : c->loc) - SP;
            uint32_t real_left = r->BPLeft, real_right = r->BPRight;
            if (cx->chr_len < real_left || cx->chr_len < real_right) {
               r->used = 1;
               break;
            }
            if (r->nt_len) real_start_insertion(cx->chr, cx->chr_len, r->NT_str, r->nt_len, &real_left, real_right);
            else real_start_deletion(cx->chr, cx->chr_len, &real_left, real_right);
            int16_t diff = (int16_t)(r->BPLeft - real_left);
            diff = !((r->BP - 1) < diff) ? diff : (int16_t)(r->BP - 1);
            if (diff > 0) {
               r->BP = (int16_t)(r->BP - diff);
               r->BPLeft -= (uint32_t)diff;
               r->BPRight -= (uint32_t)diff;
            }
            bin_region_box(cx, r, 1);
         }
      }
   }
}

static int nt_mismatch_gate(const Ctx *cx, const UP *c, const UP *f) /* src/search_deletions_nt.cpp:50-53 */
{
   return f->mm + c->mm > (int16_t)(1 + cx->P->seq_error_rate * (f->len + c->len));
}

/* ---- DI: src/search_deletions_nt.cpp:26-140 ---- */
static void search_di(Ctx *cx, Read *r)
{
   if (r->used || r->nf == 0 || r->otherchr || r->nc == 0) return;
   const UP *c = &r->c[r->nc - 1], *f = &r->f[r->nf - 1];
   if (nt_mismatch_gate(cx, c, f)) return;
   const int sum = f->len + c->len;
   if (r->matched_d == '+') {
      if (f->dir != 1) return;
      if (!(sum < r->rl && sum >= cx->P->min_matched && f->loc > c->loc + 1)) return;
      r->Left = (int32_t)(c->loc - (uint32_t)c->len + 1);
      r->Right = (int32_t)(f->loc + (uint32_t)f->len - 1);
      r->BP = (int16_t)(c->len - 1);
      r->NT_size = (uint16_t)(r->rl - f->len - c->len);
      set_nt(r, 1, r->BP + 1, r->NT_size);
      r->IndelSize = (uint32_t)((r->Right - r->Left) + r->NT_size - (r->rl - 1));
      r->BPLeft = c->loc - SP;
      r->BPRight = f->loc - SP;
      bin_region_box(cx, r, 1);
   } else if (r->matched_d == '-') {
      if (f->dir != 0) return;
      if (!(sum < r->rl && sum >= cx->P->min_matched && c->loc > f->loc + 1)) return;
      r->Left = (int32_t)(f->loc - (uint32_t)f->len + 1);
      r->Right = (int32_t)(c->loc + (uint32_t)c->len - 1);
      r->BP = (int16_t)(f->len - 1);
      r->NT_size = (uint16_t)(r->rl - c->len - f->len);
      set_nt(r, 0, r->BP + 1, r->NT_size);
      r->IndelSize = (uint32_t)((r->Right - r->Left) - (r->rl - 1) + r->NT_size);
      r->BPLeft = f->loc - SP;
      r->BPRight = c->loc - SP;
      bin_region_box(cx, r, 1);
   }
}

/* ---- TD: src/search_tandem_duplications.cpp:28-184 ---- */
static void search_td(Ctx *cx, Read *r)
{
   if (r->used || r->nf == 0 || r->otherchr) return;
   const int max_e = cx->P->max_snp[r->rl];
   const int plus = r->matched_d == '+';
   if (!plus && r->matched_d != '-') return;
   for (int e = 0; e <= max_e; e++) {
      for (int ci = plus ? 0 : r->nc - 1; plus ? ci < r->nc : ci >= 0; ci += plus ? 1 : -1) {
         if (r->used) break;
         const UP *c = &r->c[ci];
         if (c->mm > e) continue;
         for (int fi = plus ? r->nf - 1 : 0; plus ? fi >= 0 : fi < r->nf; fi += plus ? -1 : 1) {
            if (r->used) break;
            const UP *f = &r->f[fi];
            if (f->mm > e) continue;
            if (f->mm + c->mm > e) continue;
            if (plus) {
               if (f->dir != 1) continue;
               if (!(f->len + c->len == r->rl && f->loc + (uint32_t)f->len < c->loc && f->loc + (uint32_t)c->len < c->loc)) continue;
               r->Right = (int32_t)(c->loc - (uint32_t)c->len + 1);
               r->Left = (int32_t)(f->loc + (uint32_t)f->len - 1);
               r->BP = (int16_t)(c->len - 1);
               r->IndelSize = c->loc - f->loc + 1;
               r->BPRight = c->loc - SP;
               r->BPLeft = f->loc - SP;
            } else {
               if (f->dir != 0) continue;
               if (!(c->len + f->len == r->rl && c->loc + (uint32_t)c->len < f->loc && c->loc + (uint32_t)f->len < f->loc)) continue;
               r->Right = (int32_t)(f->loc - (uint32_t)f->len + 1);
               r->Left = (int32_t)(c->loc + (uint32_t)c->len - 1);
               r->BP = (int16_t)(f->len - 1);
               r->IndelSize = f->loc - c->loc + 1;
               r->BPRight = f->loc - SP;
               r->BPLeft = c->loc - SP;
            }
            if (r->BPLeft == 0) continue;
            left_most_td(cx->chr, cx->chr_len, r);
            bin_region_box(cx, r, 1);
         }
      }
   }
}

/* ---- TD_NT: src/search_tandem_duplications_nt.cpp:27-131 ---- */
static void search_td_nt(Ctx *cx, Read *r)
{
   if (r->used || r->nf == 0 || r->otherchr || r->nc == 0) return;
   const UP *c = &r->c[r->nc - 1], *f = &r->f[r->nf - 1];
   if (f->len + c->len >= r->rl) return;
   if (nt_mismatch_gate(cx, c, f)) return;
   if (r->matched_d == '+') {
      if (f->dir != 1) return;
      if (!(f->loc + (uint32_t)f->len < c->loc && f->loc + (uint32_t)c->len < c->loc && f->len + c->len > cx->P->min_matched)) return;
      r->Right = (int32_t)(c->loc - (uint32_t)c->len + 1);
      r->Left = (int32_t)(f->loc + (uint32_t)f->len - 1);
      r->BP = (int16_t)(c->len - 1);
      r->IndelSize = c->loc - f->loc + 1;
      r->NT_size = (uint16_t)(r->rl - c->len - f->len);
      set_nt(r, 1, r->BP + 1, r->NT_size);
      r->BPRight = c->loc - SP;
      r->BPLeft = f->loc - SP;
      bin_region_box(cx, r, 1);
   } else if (r->matched_d == '-') {
      if (f->dir != 0) return;
      if (!(c->loc + (uint32_t)c->len < f->loc && c->loc + (uint32_t)f->len < f->loc && f->len + c->len > cx->P->min_matched)) return;
      r->Right = (int32_t)(f->loc - (uint32_t)f->len + 1);
      r->Left = (int32_t)(c->loc + (uint32_t)c->len - 1);
      r->BP = (int16_t)(f->len - 1);
      r->IndelSize = f->loc - c->loc + 1;
      r->NT_size = (uint16_t)(r->rl - c->len - f->len);
      set_nt(r, 0, r->BP + 1, r->NT_size);
      r->BPRight = f->loc - SP;
      r->BPLeft = c->loc - SP;
      bin_region_box(cx, r, 1);
   }
}

/* ---- INV: src/search_inversions.cpp:29-281 ---- */
static void search_inv(Ctx *cx, Read *r)
{
   if (r->used || r->nf == 0 || r->otherchr || r->nc == 0) return;
   if (!(r->c[0].strand != r->f[0].strand && r->c[0].dir == r->f[0].dir)) return;
   const uint32_t mi = (uint32_t)cx->P->min_inv;
   const int max_e = cx->P->max_snp[r->rl];
   int arm; /* 0: '+' anchor outside, 1: '+' anchor inside, 2: '-' outside, 3: '-' inside */
   if (r->matched_d == '+') {
      if (r->f[0].loc > r->c[r->nc - 1].loc + mi) arm = 0;
      else if (r->f[r->nf - 1].loc + mi < r->c[0].loc) arm = 1;
      else return;
   } else if (r->matched_d == '-') {
      if (r->c[r->nc - 1].loc > r->f[0].loc + mi) arm = 2;
      else if (r->c[0].loc + mi < r->f[r->nf - 1].loc) arm = 3;
      else return;
   } else return;
   const int close_desc = (arm == 0 || arm == 2), far_dir = (arm < 2) ? 0 : 1;
   for (int e = 0; e <= max_e; e++) {
      for (int ci = close_desc ? r->nc - 1 : 0; close_desc ? ci >= 0 : ci < r->nc; ci += close_desc ? -1 : 1) {
         if (r->used) break;
         const UP *c = &r->c[ci];
         if (c->mm > e) continue;
         for (int fi = close_desc ? 0 : r->nf - 1; close_desc ? fi < r->nf : fi >= 0; fi += close_desc ? 1 : -1) {
            if (r->used) break;
            const UP *f = &r->f[fi];
            if (f->mm > e) continue;
            if (f->mm + c->mm > e) continue;
            if (f->dir != far_dir) continue;
            if (f->len + c->len != r->rl) continue;
            if (arm == 0) {
               if (!(f->loc > c->loc + mi)) continue;
               r->Left = (int32_t)((c->loc + 1) - (uint32_t)c->len);
               r->Right = (int32_t)(f->loc - (uint32_t)f->len + (uint32_t)r->rl);
               r->BP = (int16_t)(c->len - 1);
               r->IndelSize = f->loc - c->loc;
               r->BPLeft = c->loc + 1 - SP;
               r->BPRight = f->loc - SP;
            } else if (arm == 1) {
               if (!(f->loc + mi < c->loc)) continue;
               r->Right = (int32_t)(c->loc - (uint32_t)c->len + (uint32_t)r->rl);
               r->Left = (int32_t)(f->loc - (uint32_t)f->len + 1);
               r->BP = (int16_t)(f->len - 1);
               r->IndelSize = c->loc - f->loc;
               r->BPRight = c->loc - SP;
               r->BPLeft = (f->loc + 1) - SP;
            } else if (arm == 2) {
               if (!(c->loc > f->loc + mi)) continue;
               r->Left = (int32_t)(f->loc + (uint32_t)f->len - (uint32_t)r->rl);
               r->Right = (int32_t)(c->loc + (uint32_t)c->len - 1);
               r->BP = (int16_t)(f->len - 1);
               r->IndelSize = c->loc - f->loc;
               r->BPLeft = f->loc - SP;
               r->BPRight = c->loc - 1 - SP;
            } else {
               if (!(c->loc + mi < f->loc)) continue;
               r->Right = (int32_t)(f->loc + (uint32_t)f->len - 1);
               r->Left = (int32_t)(c->loc + (uint32_t)c->len - (uint32_t)r->rl);
               r->BP = (int16_t)(c->len - 1);
               r->IndelSize = f->loc - c->loc;
               r->BPLeft = c->loc - SP;
               r->BPRight = f->loc - 1 - SP;
            }
            r->nt_len = 0;
            r->NT_size = 0;
            left_most_inv(cx->chr, cx->chr_len, r);
            bin_region_box(cx, r, arm < 2); /* the '-' arms go straight to the region test, src/search_inversions.cpp:199,251 */
         }
      }
   }
}

/* ---- INV_NT: src/search_inversions_nt.cpp:27-200; both arms of a strand are tested one after the other ---- */
static void search_inv_nt(Ctx *cx, Read *r)
{
   if (r->used || r->nf == 0 || r->otherchr || r->nc == 0) return;
   const UP *c = &r->c[r->nc - 1], *f = &r->f[r->nf - 1];
   if (nt_mismatch_gate(cx, c, f)) return;
   if (!(r->c[0].strand != r->f[0].strand && r->c[0].dir == r->f[0].dir)) return;
   const uint32_t mi = (uint32_t)cx->P->min_inv;
   const int sum = f->len + c->len;
   const int gate = sum < r->rl && sum >= cx->P->min_matched;
   if (r->matched_d == '+') {
      if (f->dir != 0) return;
      if (gate && f->loc > c->loc + mi) {
         r->Left = (int32_t)((c->loc + 1) - (uint32_t)c->len);
         r->Right = (int32_t)(f->loc - (uint32_t)f->len + (uint32_t)r->rl);
         r->BP = (int16_t)(c->len - 1);
         r->IndelSize = f->loc - c->loc;
         r->NT_size = (uint16_t)(r->rl - f->len - c->len);
         set_nt(r, 1, r->BP + 1, r->NT_size);
         r->BPLeft = c->loc + 1 - SP;
         r->BPRight = f->loc - SP;
         bin_region_box(cx, r, 1);
      }
      if (gate && f->loc + mi < c->loc) {
         r->Right = (int32_t)(c->loc - (uint32_t)c->len + (uint32_t)r->rl);
         r->Left = (int32_t)(f->loc - (uint32_t)f->len + 1);
         r->BP = (int16_t)(f->len - 1);
         r->IndelSize = c->loc - f->loc;
         r->NT_size = (uint16_t)(r->rl - f->len - c->len);
         set_nt(r, 0, r->BP + 1, r->NT_size);
         r->BPRight = c->loc - SP;
         r->BPLeft = (f->loc + 1) - SP;
         bin_region_box(cx, r, 1);
      }
   } else if (r->matched_d == '-') {
      if (f->dir != 1) return;
      if (gate && c->loc > f->loc + mi) {
         r->Left = (int32_t)(f->loc + (uint32_t)f->len - (uint32_t)r->rl);
         r->Right = (int32_t)(c->loc + (uint32_t)c->len - 1);
         r->BP = (int16_t)(f->len - 1);
         r->IndelSize = c->loc - f->loc;
         r->NT_size = (uint16_t)(r->rl - f->len - c->len);
         set_nt(r, 0, r->BP + 1, r->NT_size);
         r->BPLeft = f->loc - SP;
         r->BPRight = c->loc - 1 - SP;
         bin_region_box(cx, r, 1);
      }
      if (gate && c->loc + mi < f->loc) {
         r->Right = (int32_t)(f->loc + (uint32_t)f->len - 1);
         r->Left = (int32_t)(c->loc + (uint32_t)c->len - (uint32_t)r->rl);
         r->BP = (int16_t)(c->len - 1);
         r->IndelSize = f->loc - c->loc;
         r->NT_size = (uint16_t)(r->rl - f->len - c->len);
         set_nt(r, 1, r->BP + 1, r->NT_size);
         r->BPLeft = c->loc - SP;
         r->BPRight = f->loc - 1 - SP;
         bin_region_box(cx, r, 1);
      }
   }
}

static int cmp_box_read(const void *a, const void *b)
{
   const int32_t *x = (const int32_t *)a, *y = (const int32_t *)b;
   if (x[0] != y[0]) return x[0] < y[0] ? -1 : 1;
   return x[1] < y[1] ? -1 : (x[1] > y[1]);
}

/* Runs the drivers selected by P->driver_mask over n reads of one window, in the reference's order.
 * up: [total, 5] int32 rows (len, absLoc, dir, strand, mm); up_off[n+1]; n_close[n] = how many of a read's rows are UP_Close.
 * Output stream, per driver that ran:
 *   [driver, n_vars, n_future,  n_vars x (box, read),  n_future x (read, 12 state ints),  n x 12 state ints]
 * state = Left, Right, BP, BPLeft, BPRight, IndelSize, NT_size, Used, len(NT_str), fnv1a(NT_str), first 4 bytes, last 4 bytes. */
int64_t svo_classify(const SvParams *P, const char *chr, uint32_t chr_len, int n, const char *seq_blob, const uint32_t *seq_off,
                     const char *matched_d, const int16_t *insert_size, const uint8_t *rflags, const int32_t *up,
                     const uint32_t *up_off, const uint32_t *n_close, int32_t **out_stream)
{
   Read *reads = (Read *)calloc((size_t)n, sizeof(Read));
   UP *ups = (UP *)malloc(sizeof(UP) * (up_off[n] ? up_off[n] : 1));
   for (uint32_t i = 0; i < up_off[n]; i++) {
      ups[i].len = up[5 * i];
      ups[i].loc = (uint32_t)up[5 * i + 1];
      ups[i].dir = up[5 * i + 2];
      ups[i].strand = up[5 * i + 3];
      ups[i].mm = up[5 * i + 4];
   }
   for (int i = 0; i < n; i++) {
      Read *r = &reads[i];
      r->seq = seq_blob + seq_off[i];
      r->rl = (int32_t)(seq_off[i + 1] - seq_off[i]);
      r->matched_d = matched_d[i];
      r->insert_size = insert_size[i];
      r->used = rflags[i] & 1;
      r->otherchr = (rflags[i] >> 1) & 1;
      r->c = ups + up_off[i];
      r->nc = (int)n_close[i];
      r->f = r->c + r->nc;
      r->nf = (int)(up_off[i + 1] - up_off[i]) - r->nc;
      r->NT_str = (char *)calloc((size_t)r->rl + 1, 1);
   }
   Stream out = {0, 0, 0};
   for (int d = 0; d < 7; d++) {
      if (!((P->driver_mask >> d) & 1)) continue;
      Ctx cx;
      memset(&cx, 0, sizeof cx);
      cx.out = &out;
      cx.P = P;
      cx.chr = chr;
      cx.chr_len = chr_len;
      cx.base = reads;
      for (int i = 0; i < n; i++) {
         Read *r = &reads[i];
         switch (d) {
         case 0: variant_search(&cx, r, 0); break;
         case 1: search_di(&cx, r); break;
         case 2: search_td(&cx, r); break;
         case 3: search_td_nt(&cx, r); break;
         case 4: search_inv(&cx, r); break;
         case 5: search_inv_nt(&cx, r); break;
         case 6: variant_search(&cx, r, 1); break;
         }
      }
      put(&out, d);
      put(&out, cx.n_vars);
      put(&out, cx.n_future);
      /* the reference keeps one list per box: dump box-major, push order within a box (reads are visited in index order) */
      qsort(cx.vars.v, (size_t)cx.n_vars, 2 * sizeof(int32_t), cmp_box_read);
      for (size_t k = 0; k < cx.vars.n; k++) put(&out, cx.vars.v[k]);
      for (size_t k = 0; k < cx.future.n; k++) put(&out, cx.future.v[k]);
      for (int i = 0; i < n; i++) put_state(&out, &reads[i]);
      free(cx.vars.v);
      free(cx.future.v);
   }
   for (int i = 0; i < n; i++) free(reads[i].NT_str);
   free(reads);
   free(ups);
   *out_stream = out.v;
   return (int64_t)out.n;
}

void svo_free(void *p) { free(p); }
