// tests/emul/sv_emul.cpp -- TEST INFRASTRUCTURE ONLY: pindel_b200/csrc/pg_sv.cuh (the body of pg_sv_kernel) compiled as
// plain C++ so the CPU suite can exercise the classification program without a GPU.  One "thread" per read, in a loop.
#include <stdint.h>
#include <string.h>
#include <vector>
#include "../../pindel_b200/csrc/pg_sv.cuh"
#include "../../pindel_b200/csrc/pg_pack.h"
#include "../../pindel_b200/csrc/pg_params.h"

extern "C" {

int svemu_record_size(void) { return (int)sizeof(PgSvRecord); }

// chr: padded ASCII chromosome string.  win: 11 uint32 (driver_mask, chr(ignored), window_end, num_boxes, box_size,
// region_start_defined, region_start, region_end_defined, region_end, min_inv, min_matched).  Returns the record count
// (records beyond cap are dropped; the count is still returned).
int64_t svemu_classify(const char *chr, uint32_t chr_len, double seq_error_rate, double sensitivity, const uint32_t *win, int n,
                       const uint8_t *seq_blob, const uint32_t *seq_off, const uint8_t *matched_d, const int16_t *insert_size,
                       const uint8_t *rflags, const uint32_t *run_off, const uint16_t *n_close_runs, const PgRun *runs,
                       PgSvRecord *records, uint32_t cap)
{
   PgPackedChrom pk;
   pg_pack_chromosome(chr, chr_len, PG_PADW, pk);
   PgChrom c;
   c.planes = (const uint2 *)pk.planes.data() + PG_PADW;
   c.nplane = pk.nplane.data() + PG_PADW;
   c.len = pk.len;
   c.nwords = pk.nwords;
   static PgSvParams P;
   memset(&P, 0, sizeof P);
   P.driver_mask = win[0];
   P.chr = 0;
   P.window_end = win[2];
   P.num_boxes = win[3];
   P.box_size = win[4];
   P.region_start_defined = win[5];
   P.region_start = win[6];
   P.region_end_defined = win[7];
   P.region_end = win[8];
   P.min_inv = (int32_t)win[9];
   P.min_matched = (int32_t)win[10];
   for (int k = 0; k < 1032; k++) P.nt_thr[k] = (uint16_t)(short)(1 + seq_error_rate * k);
   uint32_t table[PG_TABLE];
   pg_prob_table(0.001 + seq_error_rate, sensitivity, table);
   uint32_t counters[4] = {0, 0, 0, 0};
   PgSvEnv ev;
   ev.P = &P;
   ev.chrom = c;
   ev.records = records;
   ev.records_cap = cap;
   ev.counters = counters;
   ev.counts = NULL;
   ev.n_reads = (uint32_t)n;
   for (int i = 0; i < n; i++) {
      PgSvRead r;
      memset(&r, 0, sizeof r);
      r.seq = seq_blob + seq_off[i];
      r.rl = (int32_t)(seq_off[i + 1] - seq_off[i]);
      r.max_e = (int32_t)table[r.rl < PG_TABLE ? r.rl : PG_TABLE - 1];
      r.matched_d = matched_d[i];
      r.insert_size = insert_size[i];
      r.cr = runs + run_off[i];
      r.ncr = n_close_runs[i];
      r.fr = r.cr + r.ncr;
      r.nfr = (int32_t)(run_off[i + 1] - run_off[i]) - r.ncr;
      r.index = (uint32_t)i;
      r.nt_src = PG_NT_EMPTY;
      r.used = (rflags[i] & PG_SVR_USED) != 0;
      pg_sv_classify_read(ev, r, (rflags[i] & PG_SVR_OTHERCHR) != 0);
   }
   return (int64_t)counters[0];
}
}
