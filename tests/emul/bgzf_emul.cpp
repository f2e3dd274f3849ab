// tests/emul/bgzf_emul.cpp -- TEST INFRASTRUCTURE ONLY: the BGZF inflater of pindel_b200/csrc/pg_bgzf.cuh (the code every
// thread of pg_bgzf_inflate_kernel runs) and the host's header walk, compiled as plain C++ so that the CPU suite can hold
// them against zlib without a GPU.
#include <stdint.h>
#include <string.h>
#include <vector>
#include "../../pindel_b200/csrc/pg_bgzf.cuh"
#include "../../pindel_b200/csrc/pg_bgzf_host.h"

extern "C" {

// total inflated size, or -1 (not BGZF)
long bgzfemu_size(const uint8_t *file, size_t len)
{
   std::vector<PgzBlock> blocks;
   if (pg_bgzf_walk(file, len, blocks, NULL) < 0) return -1;
   return blocks.empty() ? 0 : (long)(blocks.back().out_off + blocks.back().isize);
}

// file must be readable 8 bytes past len.  Returns the number of blocks (status[i] = PGZ_* per block, first max_status
// of them), or -1.
long bgzfemu_inflate(const uint8_t *file, size_t len, uint8_t *out, uint32_t *status, long max_status, int check_crc)
{
   uint32_t crc_table[256];
   for (uint32_t k = 0; k < 256; k++) crc_table[k] = pgz_crc_entry(k);
   std::vector<PgzBlock> blocks;
   if (pg_bgzf_walk(file, len, blocks, NULL) < 0) return -1;
   PgzScratch sc;
   for (size_t i = 0; i < blocks.size(); i++) {
      uint32_t produced = 0;
      int rc = pgz_inflate(file + blocks[i].comp_off, blocks[i].comp_len, out + blocks[i].out_off, blocks[i].isize, &sc, 0, 1, &produced);
      if (rc == PGZ_OK && check_crc) {      // the kernel's formulation: 32 slices, combined
         const uint32_t isize = blocks[i].isize, slice = (isize + 31u) / 32u;
         uint32_t crcs[32];
         for (uint32_t l = 0; l < 32; l++) {
            const uint32_t begin = l * slice, len = begin >= isize ? 0u : (isize - begin < slice ? isize - begin : slice);
            crcs[l] = pgz_crc32(out + blocks[i].out_off + begin, len, crc_table);
         }
         if (pgz_crc_combine_slices(crcs, 32u, slice, isize) != blocks[i].crc) rc = PGZ_ERR_CRC;
      }
      if ((long)i < max_status) status[i] = (uint32_t)rc;
   }
   return (long)blocks.size();
}

// one raw DEFLATE stream (no BGZF framing): for streams made with other settings than BGZF writers use
int bgzfemu_inflate_raw(const uint8_t *in, uint32_t in_len, uint8_t *out, uint32_t isize, uint32_t *produced)
{
   PgzScratch sc;
   return pgz_inflate(in, in_len, out, isize, &sc, 0, 1, produced);
}
}
