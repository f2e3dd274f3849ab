// pindel_b200/csrc/pg_bgzf.cuh -- BGZF block inflation on the device (first half: the
// "BGZF inflate" of BAM ingestion).  A BAM file is a series of BGZF blocks -- gzip members of at most 64 KiB of payload
// whose compressed size is in the header -- which is what makes the format parallel: every block is an independent
// DEFLATE stream (RFC 1951).  The reference reaches this through htslib's bgzf_read_block/inflate_block behind
// sam_itr_next (src/reader.cpp:1099-1151, 1258-1346); here one warp inflates one block.
//
// Written from RFC 1951/1952 and the SAM specification (section 4.1, BGZF).  The decoder is a plain table-driven one:
//   * per block (= per warp) two fast tables in shared memory: literal/length codes up to 9 bits, distance codes up to
//     6 bits, entry = symbol << 4 | code length; longer codes fall back to the canonical walk over the length counts;
//   * a 64-bit bit buffer refilled with aligned 32-bit loads;
//   * every lane of the warp runs the same decode on the same bits (the Huffman chain is serial; uniform execution costs
//     no extra issue slots and needs no broadcast), and the lanes split what is parallel: the bytes of a match are copied
//     32 at a time -- one trip to L2 per match instead of one per byte, which is what a thread-per-block decoder pays
//     (profiles/r1p_bgzf_kernel_thread_per_block.md: 6.5 GB/s, stalled on the store of every copied byte).
// __host__ __device__ throughout (the host build is the one-lane case), so the CPU suite runs the identical decoder
// against zlib (tests/emul/bgzf_emul.cpp).
#pragma once
#include <stdint.h>

#ifdef __CUDACC__
#define PGZ_HD __host__ __device__ __forceinline__
#define PGZ_FN __host__ __device__
#else
#define PGZ_HD inline
#define PGZ_FN inline
#endif

#define PGZ_LIT_BITS 9
#define PGZ_DIST_BITS 6
#define PGZ_LIT_SIZE (1 << PGZ_LIT_BITS)
#define PGZ_DIST_SIZE (1 << PGZ_DIST_BITS)

#define PGZ_OK 0
#define PGZ_ERR_HEADER 1      // block type 3, stored LEN/NLEN mismatch, bad code lengths
#define PGZ_ERR_CODE 2        // a bit pattern that is no code of the current tree
#define PGZ_ERR_DIST 3        // distance beyond the start of the output
#define PGZ_ERR_OUTPUT 4      // more output than ISIZE
#define PGZ_ERR_INPUT 5       // compressed data exhausted
#define PGZ_ERR_SIZE 6        // stream ended with fewer bytes than ISIZE
#define PGZ_ERR_CRC 7         // CRC32 of the payload differs from the trailer's

struct PgzBlock {             // one BGZF block, filled by the host's header walk (pg_bgzf_walk)
   uint64_t comp_off;         // offset of the DEFLATE data in the file (after the 18-byte header for standard blocks)
   uint64_t out_off;          // offset of this block's payload in the inflated stream
   uint32_t comp_len;         // DEFLATE bytes (BSIZE + 1 - header - 8)
   uint32_t isize;            // payload bytes (ISIZE)
   uint32_t crc;              // CRC32 of the payload (trailer)
   uint32_t pad;
};

// CRC-32 (RFC 1952 section 8): reflected polynomial 0xEDB88320
PGZ_HD uint32_t pgz_crc_entry(uint32_t i)
{
   uint32_t c = i;
   for (int k = 0; k < 8; k++) c = (c & 1u) ? 0xEDB88320u ^ (c >> 1) : c >> 1;
   return c;
}
PGZ_HD uint32_t pgz_crc32(const uint8_t *p, uint32_t n, const uint32_t *table)
{
   uint32_t c = 0xFFFFFFFFu;
   for (uint32_t i = 0; i < n; i++) c = table[(c ^ p[i]) & 0xFFu] ^ (c >> 8);
   return c ^ 0xFFFFFFFFu;
}

// CRC of a concatenation from the CRCs of its parts (the polynomial identity zlib's crc32_combine uses): with the bits of a
// CRC read as a polynomial over GF(2) modulo P, appending n zero bytes multiplies by x^(8n), so
// crc(A || B) = crc(A) * x^(8 |B|) mod P  xor  crc(B).  Lets the 32 lanes of a warp checksum 32 slices of a block.
PGZ_HD uint32_t pgz_multmodp(uint32_t a, uint32_t b)
{
   uint32_t m = 1u << 31, p = 0;
   for (;;) {
      if (a & m) {
         p ^= b;
         if ((a & (m - 1u)) == 0) break;
      }
      m >>= 1;
      b = (b & 1u) ? (b >> 1) ^ 0xEDB88320u : b >> 1;
   }
   return p;
}
PGZ_HD uint32_t pgz_xpow_bytes(uint32_t n)      // x^(8 n) mod P
{
   uint32_t p = 1u << 31, sq = 0x00800000u;    // x^0; x^8 (bit 31 - k holds the coefficient of x^k)
   while (n) {
      if (n & 1u) p = pgz_multmodp(sq, p);
      sq = pgz_multmodp(sq, sq);
      n >>= 1;
   }
   return p;
}
// crcs[s] = CRC-32 of slice s, slices of `slice` bytes (the last one shorter, trailing ones possibly empty)
PGZ_HD uint32_t pgz_crc_combine_slices(const uint32_t *crcs, uint32_t n_slices, uint32_t slice, uint32_t total)
{
   if (total == 0) return 0;
   const uint32_t full = pgz_xpow_bytes(slice);
   uint32_t acc = crcs[0];
   for (uint32_t s = 1; s < n_slices; s++) {
      const uint32_t begin = s * slice;
      if (begin >= total) break;
      const uint32_t len = total - begin < slice ? total - begin : slice;
      acc = pgz_multmodp(len == slice ? full : pgz_xpow_bytes(len), acc) ^ crcs[s];
   }
   return acc;
}

// canonical Huffman code: count[len], symbols sorted by (len, symbol)
struct PgzTree {
   uint16_t count[16];
   uint16_t symbol[288];
};

struct PgzScratch {           // working memory of one decoder (shared memory on the device): 2688 bytes
   uint16_t fast_lit[PGZ_LIT_SIZE];
   uint16_t fast_dist[PGZ_DIST_SIZE];
   PgzTree lit, dist;
   uint8_t lengths[320];
};

#if defined(__CUDA_ARCH__)
#define PGZ_SYNC() __syncwarp()
#else
#define PGZ_SYNC() ((void)0)
#endif

// The fast tables are read once per symbol.  On the device they are addressed through the shared window directly
// (ld.shared with a 32-bit address computed once per block) instead of through a generic pointer, whose translation the
// compiler otherwise repeats in front of every lookup (S2R SR_CgaCtaId + address arithmetic on the decode's critical path).
#if defined(__CUDA_ARCH__)
typedef uint32_t PgzFastRef;
PGZ_HD PgzFastRef pgz_fast_ref(const uint16_t *p) { return (uint32_t)__cvta_generic_to_shared(p); }
PGZ_HD uint32_t pgz_fast_load(PgzFastRef r, uint32_t i)
{
   uint16_t v;
   asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(r + 2u * i) : "memory");
   return v;
}
#else
typedef const uint16_t *PgzFastRef;
PGZ_HD PgzFastRef pgz_fast_ref(const uint16_t *p) { return p; }
PGZ_HD uint32_t pgz_fast_load(PgzFastRef r, uint32_t i) { return r[i]; }
#endif

struct PgzBits {
   const uint32_t *words;     // 4-byte aligned view of the input
   uint32_t next_word;        // index of the next word to load
   uint32_t n_words;          // words that may be loaded (input padded to a word)
   uint64_t buf;
   uint32_t cnt;              // valid bits in buf
   uint32_t overrun;          // words requested beyond the input (each delivers zeros)
};

PGZ_HD void pgz_refill(PgzBits &b)
{
   if (b.cnt <= 32u) {
      uint32_t w = 0u;
      if (b.next_word < b.n_words) w = b.words[b.next_word];
      else b.overrun++;
      b.next_word++;
      b.buf |= (uint64_t)w << b.cnt;
      b.cnt += 32u;
   }
}
PGZ_HD uint32_t pgz_peek(const PgzBits &b, uint32_t n) { return (uint32_t)b.buf & ((1u << n) - 1u); }
PGZ_HD void pgz_drop(PgzBits &b, uint32_t n) { b.buf >>= n; b.cnt -= n; }
PGZ_HD uint32_t pgz_take(PgzBits &b, uint32_t n) { const uint32_t v = pgz_peek(b, n); pgz_drop(b, n); return v; }

// Build the canonical tree of n code lengths and its fast table (bits-wide, indexed by the next input bits, LSB first).
// Returns 0, or PGZ_ERR_HEADER for a set of lengths zlib's table builder refuses: an over-subscribed set, and an
// incomplete one unless it is empty or -- for the literal/length and distance codes only (code_length_code = false) --
// a single code of length 1 (zlib's inflate_table: "left > 0 && (type == CODES || max != 1)").  In an accepted
// incomplete set an unassigned pattern decodes to PGZ_ERR_CODE.
PGZ_FN int pgz_build(const uint8_t *lengths, int n, PgzTree &t, uint16_t *fast, int bits, bool code_length_code = false)
{
   for (int l = 0; l < 16; l++) t.count[l] = 0;
   for (int s = 0; s < n; s++) t.count[lengths[s]]++;
   uint16_t offs[16];
   offs[1] = 0;
   int left = 1;
   for (int l = 1; l < 16; l++) {
      left <<= 1;
      left -= t.count[l];
      if (left < 0) return PGZ_ERR_HEADER;
      if (l < 15) offs[l + 1] = (uint16_t)(offs[l] + t.count[l]);
   }
   if (left > 0) {
      int longest = 15;
      while (longest > 0 && t.count[longest] == 0) longest--;
      if (longest != 0 && (code_length_code || longest != 1)) return PGZ_ERR_HEADER;
   }
   for (int s = 0; s < n; s++)
      if (lengths[s]) t.symbol[offs[lengths[s]]++] = (uint16_t)s;
   for (int i = 0; i < (1 << bits); i++) fast[i] = 0;
   // codes of length l are consecutive, in symbol order, starting at first[l]
   uint32_t code = 0, index = 0;
   for (int l = 1; l <= bits; l++) {
      for (uint32_t k = 0; k < t.count[l]; k++, code++, index++) {
         uint32_t rev = 0;
         for (int i = 0; i < l; i++) rev |= ((code >> i) & 1u) << (l - 1 - i);
         const uint16_t e = (uint16_t)(t.symbol[index] << 4 | l);
         for (uint32_t i = rev; i < (1u << bits); i += 1u << l) fast[i] = e;
      }
      code <<= 1;
   }
   return PGZ_OK;
}

// one symbol: fast table, else the canonical walk (RFC 1951 3.2.2) from the first length the table does not cover
PGZ_HD int pgz_decode(PgzBits &b, const PgzTree &t, PgzFastRef fast, int bits)
{
   const uint32_t e = pgz_fast_load(fast, pgz_peek(b, (uint32_t)bits));
   if (e & 15u) {
      pgz_drop(b, e & 15u);
      return e >> 4;
   }
   int code = 0, first = 0, index = 0;
   uint64_t v = b.buf;
   for (int l = 1; l < 16; l++) {
      code |= (int)(v & 1u);
      v >>= 1;
      const int c = t.count[l];
      if (code - c < first) {
         pgz_drop(b, (uint32_t)l);
         return t.symbol[index + (code - first)];
      }
      index += c;
      first += c;
      first <<= 1;
      code <<= 1;
   }
   return -1;
}

// Inflate one raw DEFLATE stream of in_len bytes at `in` into out[0, isize).  Called by all `nlanes` lanes of a decoder
// (32 on the device, 1 on the host) with identical arguments but `lane`; sc is the decoder's scratch.  Returns PGZ_OK or an
// error (the same in every lane); *produced = bytes written.
PGZ_FN int pgz_inflate(const uint8_t *in, uint32_t in_len, uint8_t *out, uint32_t isize, PgzScratch *sc, uint32_t lane, uint32_t nlanes,
                       uint32_t *produced)
{
   const uint8_t len_order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
   uint8_t *lengths = sc->lengths;
   const PgzFastRef fast_lit = pgz_fast_ref(sc->fast_lit), fast_dist = pgz_fast_ref(sc->fast_dist);
   PgzBits b;
   const uint32_t mis = (uint32_t)((uintptr_t)in & 3u);
   b.words = reinterpret_cast<const uint32_t *>(in - mis);
   b.n_words = (in_len + mis + 3u) / 4u;
   b.next_word = 0;
   b.buf = 0;
   b.cnt = 0;
   b.overrun = 0;
   pgz_refill(b);
   pgz_drop(b, 8u * mis);
   uint32_t op = 0;
   int rc = PGZ_OK;
   for (int last = 0; !last && rc == PGZ_OK;) {
      pgz_refill(b);
      last = (int)pgz_take(b, 1);
      const uint32_t type = pgz_take(b, 2);
      if (type == 0) {                       // stored: skip to the byte boundary, LEN, NLEN, bytes
         pgz_drop(b, b.cnt & 7u);
         pgz_refill(b);
         const uint32_t len = pgz_take(b, 16);
         pgz_refill(b);
         const uint32_t nlen = pgz_take(b, 16);
         if ((len ^ 0xFFFFu) != nlen) { rc = PGZ_ERR_HEADER; break; }
         if (op + len > isize) { rc = PGZ_ERR_OUTPUT; break; }
         for (uint32_t k = 0; k < len; k++) {
            pgz_refill(b);
            const uint8_t v = (uint8_t)pgz_take(b, 8);
            if (lane == 0) out[op] = v;
            op++;
         }
         if (b.overrun > 2u) { rc = PGZ_ERR_INPUT; break; }     // LEN reaches past the end of the input
         continue;
      }
      if (type == 3) { rc = PGZ_ERR_HEADER; break; }
      if (type == 1) {                       // fixed codes (RFC 1951 3.2.6)
         for (int s = 0; s < 144; s++) lengths[s] = 8;
         for (int s = 144; s < 256; s++) lengths[s] = 9;
         for (int s = 256; s < 280; s++) lengths[s] = 7;
         for (int s = 280; s < 288; s++) lengths[s] = 8;
         pgz_build(lengths, 288, sc->lit, sc->fast_lit, PGZ_LIT_BITS);
         for (int s = 0; s < 32; s++) lengths[s] = 5;      // 30 and 31 take part in the code (3.2.6) and are refused when met
         pgz_build(lengths, 32, sc->dist, sc->fast_dist, PGZ_DIST_BITS);
      } else {                               // dynamic codes (3.2.7)
         pgz_refill(b);
         const int nlen = (int)pgz_take(b, 5) + 257, ndist = (int)pgz_take(b, 5) + 1, ncode = (int)pgz_take(b, 4) + 4;
         if (nlen > 286 || ndist > 30) { rc = PGZ_ERR_HEADER; break; }
         uint8_t cl[19];
         for (int i = 0; i < 19; i++) cl[i] = 0;
         for (int i = 0; i < ncode; i++) {
            pgz_refill(b);
            cl[len_order[i]] = (uint8_t)pgz_take(b, 3);
         }
         // the code-length code borrows the literal tree's storage and the distance table's 64 entries
         if (pgz_build(cl, 19, sc->lit, sc->fast_dist, PGZ_DIST_BITS, true) != PGZ_OK) { rc = PGZ_ERR_HEADER; break; }
         int idx = 0;
         while (idx < nlen + ndist) {
            pgz_refill(b);
            const int sym = pgz_decode(b, sc->lit, fast_dist, PGZ_DIST_BITS);
            if (sym < 0) { rc = PGZ_ERR_CODE; break; }
            if (sym < 16) {
               lengths[idx++] = (uint8_t)sym;
               continue;
            }
            int prev = 0, rep;
            if (sym == 16) {
               if (idx == 0) { rc = PGZ_ERR_HEADER; break; }
               prev = lengths[idx - 1];
               rep = 3 + (int)pgz_take(b, 2);
            } else if (sym == 17) rep = 3 + (int)pgz_take(b, 3);
            else rep = 11 + (int)pgz_take(b, 7);
            if (idx + rep > nlen + ndist) { rc = PGZ_ERR_HEADER; break; }
            while (rep--) lengths[idx++] = (uint8_t)prev;
         }
         if (rc != PGZ_OK) break;
         if (lengths[256] == 0) { rc = PGZ_ERR_HEADER; break; }   // no end-of-block code
         if (pgz_build(lengths, nlen, sc->lit, sc->fast_lit, PGZ_LIT_BITS) != PGZ_OK) { rc = PGZ_ERR_HEADER; break; }
         if (pgz_build(lengths + nlen, ndist, sc->dist, sc->fast_dist, PGZ_DIST_BITS) != PGZ_OK) { rc = PGZ_ERR_HEADER; break; }
      }
      for (;;) {                             // the symbols of this block
         pgz_refill(b);
         int sym = pgz_decode(b, sc->lit, fast_lit, PGZ_LIT_BITS);
         if (sym < 0) { rc = PGZ_ERR_CODE; break; }
         if (sym < 256) {
            if (op >= isize) { rc = PGZ_ERR_OUTPUT; break; }
            if (lane == 0) out[op] = (uint8_t)sym;
            op++;
            continue;
         }
         if (sym == 256) break;
         sym -= 257;
         if (sym >= 29) { rc = PGZ_ERR_CODE; break; }
         // RFC 1951 3.2.5, the two tables in closed form: length codes 257..264 are 3..10, then four codes per extra bit,
         // code 285 is 258; distance codes 0..3 are 1..4, then two codes per extra bit
         const uint32_t lx = sym < 8 || sym == 28 ? 0u : ((uint32_t)sym >> 2) - 1u;
         const uint32_t len = (sym == 28 ? 258u : sym < 8 ? 3u + (uint32_t)sym : 3u + ((4u + ((uint32_t)sym & 3u)) << lx)) + pgz_take(b, lx);
         pgz_refill(b);
         const int ds = pgz_decode(b, sc->dist, fast_dist, PGZ_DIST_BITS);
         if (ds < 0 || ds >= 30) { rc = PGZ_ERR_CODE; break; }
         const uint32_t dx = ds < 4 ? 0u : ((uint32_t)ds >> 1) - 1u;
         const uint32_t d = (ds < 4 ? 1u + (uint32_t)ds : 1u + ((2u + ((uint32_t)ds & 1u)) << dx)) + pgz_take(b, dx);
         if (d > op) { rc = PGZ_ERR_DIST; break; }
         if (op + len > isize) { rc = PGZ_ERR_OUTPUT; break; }
         // every source byte lies before op (a match that overlaps its own output repeats its first d bytes), so the lanes
         // copy independently once the earlier stores of the other lanes are visible
         PGZ_SYNC();
         const uint8_t *src = out + op - d;
         if (d >= len) {                     // the usual case: no overlap, and most matches are shorter than a warp
            if (lane < len) out[op + lane] = src[lane];
            for (uint32_t k = lane + nlanes; k < len; k += nlanes) out[op + k] = src[k];
         } else {
            for (uint32_t k = lane; k < len; k += nlanes) out[op + k] = src[k % d];
         }
         op += len;
      }
      if (b.overrun > 2u) rc = PGZ_ERR_INPUT;   // the bit buffer may look one or two words past the end; decoding from there is an error
   }
   PGZ_SYNC();
   *produced = op;
   // Every bit the decoder consumed must lie inside the input: beyond it the bit buffer delivers zeros (the look-ahead of one
   // or two words is harmless, decoding from there is not), so a stream that stops early -- cut off, or a damaged length
   // that carries a block over the end -- is refused exactly, as zlib refuses it for want of input.
   if (rc == PGZ_OK && 32ull * b.next_word - b.cnt > 8ull * ((uint64_t)mis + in_len)) rc = PGZ_ERR_INPUT;
   if (rc == PGZ_OK && op != isize) rc = PGZ_ERR_SIZE;
   return rc;
}

#ifdef __CUDACC__
#define PGZ_WARPS 8      // decoders per CTA: 8 x 2688 B of scratch

// status[b] = PGZ_* of block b; one warp per block.  check_crc: also recompute the payload's CRC32 (what htslib's
// bgzf_read_block does) while the block is still in cache: every lane checksums one 32nd of the payload, lane 0 combines.
__global__ void __launch_bounds__(32 * PGZ_WARPS) pg_bgzf_inflate_kernel(const uint8_t *__restrict__ file, const PgzBlock *__restrict__ blocks,
                                                                        uint32_t n_blocks, uint8_t *__restrict__ out, uint32_t *__restrict__ status,
                                                                        int check_crc)
{
   __shared__ PgzScratch scratch[PGZ_WARPS];
   __shared__ uint32_t crc_table[256];
   for (uint32_t k = threadIdx.x; k < 256u; k += 32 * PGZ_WARPS) crc_table[k] = pgz_crc_entry(k);
   __syncthreads();
   const uint32_t w = threadIdx.x >> 5, lane = threadIdx.x & 31u;
   const uint32_t i = blockIdx.x * PGZ_WARPS + w;
   if (i >= n_blocks) return;
   const PgzBlock bl = blocks[i];
   uint32_t produced = 0;
   int rc = pgz_inflate(file + bl.comp_off, bl.comp_len, out + bl.out_off, bl.isize, &scratch[w], lane, 32u, &produced);
   if (rc == PGZ_OK && check_crc) {
      const uint32_t slice = (bl.isize + 31u) / 32u, begin = lane * slice;
      const uint32_t len = begin >= bl.isize ? 0u : (bl.isize - begin < slice ? bl.isize - begin : slice);
      uint32_t *crcs = reinterpret_cast<uint32_t *>(scratch[w].fast_lit);     // the tables are no longer needed
      crcs[lane] = pgz_crc32(out + bl.out_off + begin, len, crc_table);
      __syncwarp();
      if (lane == 0 && pgz_crc_combine_slices(crcs, 32u, slice, bl.isize) != bl.crc) rc = PGZ_ERR_CRC;
   }
   if (lane == 0) status[i] = (uint32_t)rc;
}
#endif
