// tests/emul/bd_emul.cpp -- TEST INFRASTRUCTURE ONLY: the BreakDancer-mask functions of pindel_b200/csrc/pg_bd.cuh and
// pg_bd_host.h compiled for the host and driven in the kernels' formulation (count pass, offsets, fill pass).
#include <stdint.h>
#include <string.h>
#include <vector>

#include "../../pindel_b200/csrc/pg_bd_host.h"

extern "C" {
// -> total number of windows; region_off [n+1]; regions (chr, start, end) up to cap rows.  -1: bad event list.
long bdemu_clusters(const int32_t *events6, uint32_t n_events, const int32_t *chr_rank, int32_t n_chr, int32_t chr, uint32_t start, uint32_t end, uint32_t n_q,
                    const uint32_t *q, uint32_t *region_off, int32_t *regions3, uint32_t cap)
{
   std::vector<PgBdEvent> ev(n_events);
   for (uint32_t i = 0; i < n_events; i++) {
      const int32_t *e = events6 + 6 * i;
      ev[i].chr_a = e[0]; ev[i].pos_a = (uint32_t)e[1]; ev[i].pos2_a = (uint32_t)e[2]; ev[i].chr_b = e[3]; ev[i].pos_b = (uint32_t)e[4]; ev[i].pos2_b = (uint32_t)e[5];
   }
   PgBdPrepared P;
   if (pgbd_prepare(ev.data(), n_events, chr_rank, n_chr, chr, start, end, P) != 0) return -1;
   PgBdRegionView R;
   R.start = P.start.data(); R.drop = P.drop.data(); R.sel = P.sel.data(); R.m = (uint32_t)P.start.size(); R.z = P.z; R.ws = P.ws; R.we = P.we;
   uint32_t total = 0;
   for (uint32_t i = 0; i < n_q; i++) {                // count pass + scan
      uint32_t S, E;
      pgbd_range(R, q[i], S, E);
      region_off[i] = total;
      total += E > S ? pgbd_cluster(R, S, E, NULL, 0) : 0u;
   }
   region_off[n_q] = total;
   for (uint32_t i = 0; i < n_q; i++) {                // fill pass
      const uint32_t c = region_off[i + 1] - region_off[i];
      if (!c || region_off[i] + c > cap) continue;
      uint32_t S, E;
      pgbd_range(R, q[i], S, E);
      std::vector<PgRegion> tmp(c);
      pgbd_cluster(R, S, E, tmp.data(), c);
      for (uint32_t k = 0; k < c; k++) {
         regions3[3 * (region_off[i] + k)] = tmp[k].chr;
         regions3[3 * (region_off[i] + k) + 1] = (int32_t)tmp[k].start;
         regions3[3 * (region_off[i] + k) + 2] = (int32_t)tmp[k].end;
      }
   }
   return (long)total;
}
}
