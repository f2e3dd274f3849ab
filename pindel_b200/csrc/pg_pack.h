// pindel_b200/csrc/pg_pack.h -- host-side packing of the reference's in-memory chromosome strings
// (Genome::loadChromosome, src/pindel.cpp:270-312: upper-case A/C/G/T, everything else 'N', 100 000 'N'
// of padding on both sides) into the device layout of pg_warp.cuh: {lo, hi} bit-planes, 32 bases per
// uint2, plus a 1-bit "not A/C/G/T" plane, each with PG_PADW guard words (flagged N) on both sides.
#pragma once
#include <stdint.h>
#include <stddef.h>
#include <vector>

struct PgPackedChrom {
   std::vector<uint32_t> planes;   // 2 * (nwords + 2*PADW): lo, hi interleaved (== uint2 array)
   std::vector<uint32_t> nplane;   // nwords + 2*PADW
   uint32_t len = 0;
   uint32_t nwords = 0;
};

static inline void pg_pack_chromosome(const char *seq, uint32_t len, int padw, PgPackedChrom &out)
{
   uint32_t nwords = (len + 31u) / 32u;
   out.len = len;
   out.nwords = nwords;
   size_t total = (size_t)nwords + 2u * (size_t)padw;
   out.planes.assign(2 * total, 0u);
   out.nplane.assign(total, 0xffffffffu);            // guard band and tail bits read as N
   for (uint32_t w = 0; w < nwords; w++) {
      uint32_t lo = 0, hi = 0, nn = 0;
      uint32_t base = w * 32u;
      uint32_t cnt = (len - base) < 32u ? (len - base) : 32u;
      for (uint32_t j = 0; j < cnt; j++) {
         char c = seq[base + j];
         uint32_t code;
         switch (c) {
         case 'A': code = 0; break;
         case 'C': code = 1; break;
         case 'G': code = 2; break;
         case 'T': code = 3; break;
         default: code = 0; nn |= 1u << j; break;
         }
         lo |= (code & 1u) << j;
         hi |= (code >> 1) << j;
      }
      if (cnt < 32u) nn |= ~0u << cnt;
      size_t k = (size_t)padw + w;
      out.planes[2 * k] = lo;
      out.planes[2 * k + 1] = hi;
      out.nplane[k] = nn;
   }
}
