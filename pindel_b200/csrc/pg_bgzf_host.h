// pindel_b200/csrc/pg_bgzf_host.h -- host part of pg_bgzf_inflate: the walk over the BGZF block headers (SAM specification
// section 4.1: gzip member with FEXTRA, subfield 'B','C' carrying BSIZE = total block size - 1; trailer CRC32, ISIZE).
// Each hop reads 18 + 8 bytes, so the walk costs a few nanoseconds per 64 KiB block; everything else runs on the device.
#pragma once
#include <stdint.h>
#include <stddef.h>
#include <vector>
#include "pg_bgzf.cuh"

// -> number of blocks, or -1 when the bytes at `pos` are not a BGZF block header (bad_at = that offset).
static inline long pg_bgzf_walk(const uint8_t *file, size_t len, std::vector<PgzBlock> &blocks, size_t *bad_at)
{
   blocks.clear();
   size_t pos = 0;
   uint64_t out_off = 0;
   while (pos < len) {
      if (bad_at) *bad_at = pos;
      if (len - pos < 18 || file[pos] != 0x1f || file[pos + 1] != 0x8b || file[pos + 2] != 8 || !(file[pos + 3] & 4)) return -1;
      const uint32_t xlen = file[pos + 10] | (uint32_t)file[pos + 11] << 8;
      if (len - pos < 12 + (size_t)xlen) return -1;
      uint32_t bsize = 0;
      bool found = false;
      for (uint32_t x = 0; x + 4 <= xlen;) {      // the extra subfields: SI1 SI2 SLEN data
         const uint8_t *f = file + pos + 12 + x;
         const uint32_t slen = f[2] | (uint32_t)f[3] << 8;
         if (f[0] == 'B' && f[1] == 'C' && slen == 2 && x + 6 <= xlen) { bsize = f[4] | (uint32_t)f[5] << 8; found = true; }
         x += 4 + slen;
      }
      if (!found || (size_t)bsize + 1 > len - pos || bsize + 1 < 12 + xlen + 8) return -1;
      const uint8_t *tr = file + pos + bsize + 1 - 8;
      PgzBlock b;
      b.comp_off = pos + 12 + xlen;
      b.comp_len = bsize + 1 - 12 - xlen - 8;
      b.isize = tr[4] | (uint32_t)tr[5] << 8 | (uint32_t)tr[6] << 16 | (uint32_t)tr[7] << 24;
      b.out_off = out_off;
      b.crc = tr[0] | (uint32_t)tr[1] << 8 | (uint32_t)tr[2] << 16 | (uint32_t)tr[3] << 24;
      b.pad = 0;
      if (b.isize > 65536u) return -1;             // BGZF payloads are at most 64 KiB
      blocks.push_back(b);
      out_off += b.isize;
      pos += (size_t)bsize + 1;
   }
   return (long)blocks.size();
}
