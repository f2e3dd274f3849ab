// pindel_b200/csrc/pg_params.h -- host-side derivation of the search parameters the kernels read.
//
//  * g[len]  : the reference's g_maxMismatch table (createProbTable / probOfReadWithTheseErrors,
//              src/pindel.cpp:781-819), the binomial mismatch budget.  It is computed here in double
//              arithmetic with the same operation order, because its change points depend on pow()
//              rounding and every length must agree bit for bit with the CPU build.
//  * thr[L]  : CheckMismatches' acceptance threshold (src/searcher.cpp:366,383):
//              `(float)NumMismatches >= (float)(size * MaximumAllowedMismatchRate)` turned into the least
//              integer mismatch count that passes, so the device compares integers.
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>
#include "pg_warp_types.h"

static inline void pg_prob_table(double seq_error_rate_plus_snp, double sensitivity, uint32_t *out /* PG_TABLE */)
{
   for (unsigned length = 0; length < PG_TABLE; length++) {
      double total = 0.0;
      out[length] = 0;
      for (unsigned nerr = 0; nerr <= length; nerr++) {
         double chanceOfCorrect = 1.0 - seq_error_rate_plus_snp;
         unsigned correct = length - nerr;
         double matched = pow(chanceOfCorrect, correct);
         double mismatched = 1.0;
         for (unsigned i = 0; i < nerr; i++) mismatched *= (((length - i) * seq_error_rate_plus_snp) / (nerr - i));
         total += matched * mismatched;
         if (total > sensitivity) { out[length] = nerr + 1; break; }
      }
   }
   out[0] = out[1] = out[2] = out[3] = 0;
}

// Returns 0 on success, -1 when a table entry does not fit the device representation.
static inline int pg_fill_params(PgParams *p, int add_mm, int min_close, int min_perfect, int max_range_index,
                                 double max_mismatch_rate, const uint32_t *max_mismatch /* PG_TABLE */)
{
   memset(p, 0, sizeof(*p));
   p->add_mm = add_mm;
   p->min_close = min_close;
   p->min_perfect = min_perfect;
   p->max_range_index = max_range_index;
   p->filter_min_pos = PG_FILTER_MIN_POS;
   for (int i = 0; i < 512; i++) {
      uint32_t v = (i < PG_TABLE) ? max_mismatch[i] : 0u;
      if (v > 250u) return -1;
      p->g[i] = (uint8_t)v;
      float max_allowed = (float)((double)i * max_mismatch_rate);
      int n = 0;
      while ((float)n < max_allowed && n < 65535) n++;
      p->thr[i] = (uint16_t)n;
   }
   return 0;
}
