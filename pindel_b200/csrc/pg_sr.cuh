// pindel_b200/csrc/pg_sr.cuh -- the reader's split-read filter on the indexed stream (rest of row N3).
//
// What ReadInBamReads_SR (src/reader.cpp:483-560) hands to its ReadBuffer for one BAM file and window: every record the
// region iterator delivers goes through fetch_func_SR (src/reader.cpp:1099-1151), which pairs records on their read name as
// they stream by (a record whose name is waiting completes the pair and frees the name; otherwise it waits) and calls
//    build_record_SR(anchor, read)      src/reader.cpp:799-900   for reads isWeirdRead accepts (:659-690), behind an anchor
//                                                                isGoodAnchor accepts (:561-612) or behind themselves
//    build_record_RefRead(anchor, read) src/reader.cpp:903-923   for reads isRefRead accepts (:619-654)
// in a fixed order per record.  The streaming hash map is the only serial part, and its outcome has a closed form: among the
// delivered records of one name in stream order, the 2nd, 4th, ... complete a pair with the one just before them.  So:
//    1. per record (pg_sr_info_kernel): name hash, the three predicates, the N statistics of the sequence
//    2. names into an open-addressing table, the records of one slot chained (pg_sr_insert_kernel)
//    3. per record (pg_sr_pair_kernel): rank among the records of its own name (names compared byte by byte, so a hash
//       collision only lengthens a chain), partner, the set of calls the reference makes at this record
//    4. positions by prefix count, then every call writes its row (pg_sr_scatter_kernel): the sequence decoded, N-trimmed,
//       reverse-complemented for reverse-strand reads, the anchor's strand and position
// The per-record functions are __host__ __device__; tests/emul/sr_emul.cpp runs them on the CPU in the kernels' formulation.
#pragma once
#include "pg_bam.cuh"

struct PgSrQuery {                // pg_sr_query (include/pindel_pg.h); the first five fields are PgRpQuery's: the same iterator
   int32_t tid, beg, end;
   int32_t insert_size;
   uint32_t min_anchor_quality;   // userSettings->minimalAnchorQuality (-A)
   uint32_t chr_len;              // CurrentChrSeq->size() - 2 * g_SpacerBeforeAfter: MatchedRelPos is clamped to it
   int32_t nm_limit;              // userSettings->NM
   uint32_t pad;
   double max_mismatch_rate;      // userSettings->MaximumAllowedMismatchRate (-u)
};

struct PgSrCand {                 // pg_sr_cand: one build_record_SR call that reached addRead
   uint32_t anchor, read;         // records (indices into pg_bam_index's table): mapped_read, unmapped_read
   uint32_t matched_rel_pos;      // MatchedRelPos
   uint32_t read_length;          // UnmatchedSeq.size() after setUnmatchedSeq; the bytes are at seqs[k * stride]
   uint8_t matched_d;             // '+' / '-'
   uint8_t ms;                    // MS = the anchor's mapping quality
   uint16_t pad;
};

struct PgSrRef {                  // pg_sr_ref: one build_record_RefRead call that was kept
   uint32_t pos;                  // REF_READ::Pos
   int16_t read_length;           // REF_READ::ReadLength (short)
   uint8_t mq, pad;
};

struct PgSrInfo {
   uint64_t key;                  // hash of the read name, never 0
   uint32_t lead, length;         // N's trimmed in front; `length` of build_record_SR after both trims
   uint8_t bits;                  // PGSR_*
   uint8_t pad[7];
};

#define PGSR_DELIVERED 1u         // the iterator delivers the record
#define PGSR_WEIRD 2u             // isWeirdRead
#define PGSR_ANCHOR 4u            // isGoodAnchor
#define PGSR_REFREAD 8u           // isRefRead
#define PGSR_SEQ_OK 16u           // passes build_record_SR's N filter and minimum length as the read
#define PGSR_QUAL_OK 32u          // passes build_record_SR's / build_record_RefRead's quality test
#define PGSR_MALFORMED 64u        // the variable-length fields do not fit the record's block_size (htslib's bam_read1 fails there)

#define PGSR_E_SELF 1u            // waiting record: build_record_SR(b1, b1)
#define PGSR_E0 2u                // completing record b1 with waiting b2: build_record_SR(b2, b2)
#define PGSR_E1 4u                //    build_record_SR(b1, b2)
#define PGSR_R0 8u                //    build_record_RefRead(b1, b2)
#define PGSR_E2 16u               //    build_record_SR(b2, b1)
#define PGSR_R1 32u               //    build_record_RefRead(b2, b1)
#define PGSR_ERR_INSERT 64u       // a call would stop the reference: InsertSize <= length (src/reader.cpp:866-874)
#define PGSR_ERR_MALFORMED 128u   // a delivered record is malformed: nothing is returned for the window
#define PGSR_NONE 0xFFFFFFFFu

// bam_aux_get(b, "NM") + bam_aux2i: walk the tag list; *found = the tag exists.  A malformed list ends the walk.
PGB_HD int32_t pgsr_aux_nm(const uint8_t *aux, const uint8_t *end, bool *found)
{
   *found = false;
   const uint8_t *p = aux;
   while (p + 3 <= end) {
      const uint8_t t0 = p[0], t1 = p[1], ty = p[2];
      p += 3;
      uint64_t n;
      switch (ty) {
      case 'A': case 'c': case 'C': n = 1; break;
      case 's': case 'S': n = 2; break;
      case 'i': case 'I': case 'f': n = 4; break;
      case 'd': n = 8; break;
      case 'Z': case 'H': {
         const uint8_t *e = p;
         while (e < end && *e) e++;
         if (e >= end) return 0;
         n = (uint64_t)(e - p) + 1;
         break;
      }
      case 'B': {
         if (p + 5 > end) return 0;
         const uint8_t sub = p[0];
         const uint64_t cnt = pgb_le32(p + 1);
         const uint64_t w = (sub == 'c' || sub == 'C') ? 1 : (sub == 's' || sub == 'S') ? 2 : (sub == 'i' || sub == 'I' || sub == 'f') ? 4 : 0;
         if (w == 0) return 0;
         n = 5 + w * cnt;
         break;
      }
      default: return 0;
      }
      if (n > (uint64_t)(end - p)) return 0;
      if (t0 == 'N' && t1 == 'M') {
         *found = true;
         switch (ty) {
         case 'c': return (int32_t)(int8_t)p[0];
         case 'C': return (int32_t)p[0];
         case 's': return (int32_t)(int16_t)pgb_le16(p);
         case 'S': return (int32_t)pgb_le16(p);
         case 'i': case 'I': return (int32_t)pgb_le32(p);
         default: return 0;                              // bam_aux2i of a non-integer tag
         }
      }
      p += n;
   }
   return 0;
}

// FNV-1a over the NUL-terminated read name (the key of the reference's khash is the C string)
PGB_HD uint64_t pgsr_name_key(const uint8_t *rec)
{
   const uint32_t l_name = rec[8];
   uint64_t h = 14695981039346656037ull;
   for (uint32_t k = 0; k < l_name && rec[32 + k]; k++) h = (h ^ rec[32 + k]) * 1099511628211ull;
   h ^= h >> 29;
   return h ? h : 1ull;
}

PGB_HD bool pgsr_same_name(const uint8_t *a, const uint8_t *b)
{
   const uint32_t la = a[8], lb = b[8];
   for (uint32_t k = 0;; k++) {
      const uint8_t ca = k < la ? a[32 + k] : 0, cb = k < lb ? b[32 + k] : 0;
      if (ca != cb) return false;
      if (ca == 0) return true;
   }
}

PGB_HD uint32_t pgsr_base(const uint8_t *seq, uint32_t i) { return (seq[i >> 1] >> ((i & 1u) ? 0 : 4)) & 15u; }

// everything the pairing step needs to know about one delivered record
PGB_HD void pgsr_analyse(const uint8_t *d, const PgBamRec &r, const PgSrQuery &q, PgSrInfo &info)
{
   const uint8_t *rec = d + r.off + 4;
   const uint32_t rec_len = pgb_le32(d + r.off);
   const uint32_t l_name = rec[8], n_cigar = pgb_le16(rec + 12), flag = r.flag, l_qseq = pgb_le32(rec + 16);
   const uint8_t *cig = rec + 32 + l_name, *seq = cig + 4ull * n_cigar;
   const uint8_t *aux = seq + (l_qseq + 1u) / 2u + l_qseq, *end = rec + rec_len;
   info.lead = info.length = 0;
   for (int k = 0; k < 7; k++) info.pad[k] = 0;
   if (l_name == 0 || 32ull + l_name + 4ull * n_cigar + (l_qseq + 1ull) / 2ull + l_qseq > rec_len) {      // never look inside such a record
      info.key = 1;
      info.bits = (uint8_t)(PGSR_DELIVERED | PGSR_MALFORMED);
      return;
   }
   info.key = pgsr_name_key(rec);
   uint32_t bits = PGSR_DELIVERED;
   const bool mapped = !(flag & 4u);
   bool has_nm = false;
   const int32_t nm = aux <= end ? pgsr_aux_nm(aux, end, &has_nm) : 0;
   int32_t non_m = 0;                  // bam_cigar2mismatch: bases of every operation but M
   bool gap_or_clip = false, indel = false;
   for (uint32_t k = 0; k < n_cigar; k++) {
      const uint32_t c = pgb_le32(cig + 4 * k), op = c & 15u;
      if (op != 0u) non_m += (int32_t)(c >> 4);
      if (op >= 1u && op <= 6u) gap_or_clip = true;                // I, D, N, S, H, P
      if (op == 1u || op == 2u) indel = true;
   }
   const bool flagged = (flag & (0x100u | 0x200u | 0x400u)) != 0;     // secondary, QC fail, duplicate
   // isWeirdRead
   if (!mapped || gap_or_clip || (has_nm && nm != 0) || nm + non_m > 0) bits |= PGSR_WEIRD;
   // isGoodAnchor
   if (mapped && (uint32_t)r.mapq >= q.min_anchor_quality && (q.min_anchor_quality == 0 || !flagged)) bits |= PGSR_ANCHOR;
   // isRefRead
   {
      bool ok = !flagged;
      if (ok && has_nm) {
         const int max_edits = (int)((double)(int32_t)l_qseq * q.max_mismatch_rate) + 1;
         if (nm > q.nm_limit || nm > max_edits) ok = false;
      }
      if (ok && n_cigar > 2 && indel) ok = false;                  // HasIndel
      if (ok && mapped && nm <= 2 && non_m <= 2) bits |= PGSR_REFREAD;
   }
   if ((int16_t)r.mapq >= (int16_t)q.min_anchor_quality) bits |= PGSR_QUAL_OK;
   if (bits & PGSR_WEIRD) {            // the N filter of build_record_SR, for when this record is the read of a call
      uint32_t lead = 0;
      while (lead < l_qseq && pgsr_base(seq, lead) == 15u) lead++;
      uint32_t length = l_qseq - lead;
      while (length > 0 && pgsr_base(seq, lead + length - 1) == 15u) length--;
      uint32_t n_count = 0;
      for (uint32_t k = 0; k < length; k++) n_count += pgsr_base(seq, lead + k) == 15u;
      info.lead = lead;
      info.length = length;
      if (!((int)n_count > (int)((double)(int)length * .10) || (int)length < 22)) bits |= PGSR_SEQ_OK;
   }
   info.bits = (uint8_t)bits;
}

// The calls fetch_func_SR makes at record `self`; partner = the waiting record it completes a pair with (NULL: self waits).
PGB_HD uint32_t pgsr_emit_mask(const PgSrInfo &self, const PgSrInfo *partner, const PgSrQuery &q)
{
   uint32_t m = 0;
   const uint32_t b1 = self.bits;
   if (b1 & PGSR_MALFORMED) return PGSR_ERR_MALFORMED;
   #define PGSR_BUILD(anchor_bits, read) \
      (((anchor_bits) & PGSR_QUAL_OK) && ((read).bits & PGSR_SEQ_OK) ? ((q.insert_size <= (int32_t)(read).length) ? PGSR_ERR_INSERT : 1u) : 0u)
   if (!partner) {
      if (b1 & PGSR_WEIRD) { const uint32_t b = PGSR_BUILD(b1, self); m |= b == 1u ? PGSR_E_SELF : b; }
      return m;
   }
   const uint32_t b2 = partner->bits;
   if (b2 & PGSR_WEIRD) { const uint32_t b = PGSR_BUILD(b2, *partner); m |= b == 1u ? PGSR_E0 : b; }
   if (b1 & PGSR_ANCHOR) {
      if (b2 & PGSR_WEIRD) { const uint32_t b = PGSR_BUILD(b1, *partner); m |= b == 1u ? PGSR_E1 : b; }
      if ((b2 & PGSR_REFREAD) && (b2 & PGSR_QUAL_OK)) m |= PGSR_R0;
   }
   if (b2 & PGSR_ANCHOR) {
      if (b1 & PGSR_WEIRD) { const uint32_t b = PGSR_BUILD(b2, self); m |= b == 1u ? PGSR_E2 : b; }
      if ((b1 & PGSR_REFREAD) && (b1 & PGSR_QUAL_OK)) m |= PGSR_R1;
   }
   #undef PGSR_BUILD
   return m;
}
PGB_HD uint32_t pgsr_count_sr(uint32_t m) { return ((m & PGSR_E_SELF) != 0) + ((m & PGSR_E0) != 0) + ((m & PGSR_E1) != 0) + ((m & PGSR_E2) != 0); }
PGB_HD uint32_t pgsr_count_ref(uint32_t m) { return ((m & PGSR_R0) != 0) + ((m & PGSR_R1) != 0); }

// Rank of record i among the delivered records of its name that precede it, through the chain of its table slot.
// -> the waiting record i completes a pair with, or PGSR_NONE when i itself waits.
PGB_HD uint32_t pgsr_partner(const uint8_t *d, const PgBamRec *recs, const uint32_t *next, uint32_t chain_head, uint32_t i)
{
   const uint8_t *mine = d + recs[i].off + 4;
   uint32_t rank = 0, pred = PGSR_NONE;
   for (uint32_t j = chain_head; j != PGSR_NONE; j = next[j]) {
      if (j >= i) continue;
      if (!pgsr_same_name(mine, d + recs[j].off + 4)) continue;
      rank++;
      if (pred == PGSR_NONE || j > pred) pred = j;
   }
   return (rank & 1u) ? pred : PGSR_NONE;
}

// build_record_SR(anchor, read) past its filters: the row and the sequence bytes (out_seq has room for info(read).length)
PGB_HD void pgsr_build(const uint8_t *d, const PgBamRec *recs, const PgSrInfo *info, const PgSrQuery &q, uint32_t anchor, uint32_t read, PgSrCand &c,
                       uint8_t *out_seq)
{
   const PgBamRec &ra = recs[anchor], &rr = recs[read];
   const uint8_t *rec = d + rr.off + 4;
   const uint8_t *seq = rec + 32 + rec[8] + 4ull * pgb_le16(rec + 12);
   const uint32_t lead = info[read].lead, length = info[read].length;
   uint32_t len = length;
   if (rr.flag & 16u) {                // ReverseComplement (Convert2RC4N): A<->T, C<->G, N, anything else NUL
      for (uint32_t k = 0; k < length; k++) {
         const uint32_t b = pgsr_base(seq, lead + length - 1 - k);
         out_seq[k] = b == 1u ? 'T' : b == 2u ? 'G' : b == 4u ? 'C' : b == 8u ? 'A' : b == 15u ? 'N' : 0;
      }
      while (len > 0 && out_seq[len - 1] == 0) len--;            // setUnmatchedSeq strips trailing non-alphanumerics
   } else {
      for (uint32_t k = 0; k < length; k++) {
         const uint32_t b = pgsr_base(seq, lead + k);
         out_seq[k] = (uint8_t)"=ACMGRSVTWYHKDBN"[b];
      }
      while (len > 0 && out_seq[len - 1] == '=') len--;
   }
   uint32_t pos = (uint32_t)ra.pos;
   if (ra.flag & 16u) {                // bam_cigar2len: M, I, S add, D subtracts
      const uint8_t *arec = d + ra.off + 4;
      const uint8_t *cig = arec + 32 + arec[8];
      const uint32_t n_cigar = pgb_le16(arec + 12);
      int32_t l = 0;
      for (uint32_t k = 0; k < n_cigar; k++) {
         const uint32_t cc = pgb_le32(cig + 4 * k), op = cc & 15u;
         if (op == 0u || op == 1u || op == 4u) l += (int32_t)(cc >> 4);
         if (op == 2u) l -= (int32_t)(cc >> 4);
      }
      pos += (uint32_t)l;
   }
   if (pos > q.chr_len) pos = q.chr_len;
   c.anchor = anchor;
   c.read = read;
   c.matched_rel_pos = pos;
   c.read_length = len;
   c.matched_d = (ra.flag & 16u) ? '-' : '+';
   c.ms = ra.mapq;
   c.pad = 0;
}

PGB_HD void pgsr_build_ref(const uint8_t *d, const PgBamRec &r, PgSrRef &o)
{
   o.pos = (uint32_t)r.pos;
   o.read_length = (int16_t)pgb_le32(d + r.off + 4 + 16);
   o.mq = r.mapq;
   o.pad = 0;
}

// every call of record i in the reference's order: rows at cands[off_sr ...] / refs[off_ref ...]
PGB_HD void pgsr_scatter(const uint8_t *d, const PgBamRec *recs, const PgSrInfo *info, const PgSrQuery &q, uint32_t i, uint32_t partner, uint32_t mask,
                         uint32_t off_sr, uint32_t off_ref, PgSrCand *cands, uint8_t *seqs, uint32_t stride, PgSrRef *refs)
{
   if (mask & PGSR_E_SELF) { pgsr_build(d, recs, info, q, i, i, cands[off_sr], seqs + (uint64_t)off_sr * stride); off_sr++; }
   if (mask & PGSR_E0) { pgsr_build(d, recs, info, q, partner, partner, cands[off_sr], seqs + (uint64_t)off_sr * stride); off_sr++; }
   if (mask & PGSR_E1) { pgsr_build(d, recs, info, q, i, partner, cands[off_sr], seqs + (uint64_t)off_sr * stride); off_sr++; }
   if (mask & PGSR_R0) pgsr_build_ref(d, recs[partner], refs[off_ref++]);
   if (mask & PGSR_E2) { pgsr_build(d, recs, info, q, partner, i, cands[off_sr], seqs + (uint64_t)off_sr * stride); off_sr++; }
   if (mask & PGSR_R1) pgsr_build_ref(d, recs[i], refs[off_ref++]);
}

PGB_HD PgRpQuery pgsr_iterator(const PgSrQuery &q)
{
   PgRpQuery r;
   r.tid = q.tid; r.beg = q.beg; r.end = q.end; r.insert_size = q.insert_size; r.min_anchor_quality = q.min_anchor_quality;
   return r;
}

#ifdef __CUDACC__
// counters: [0] stop index (pg_bam_stop_kernel), [1] records delivered, [2] longest `length` of a weird record, [3] error bits
__global__ void pg_sr_info_kernel(const uint8_t *__restrict__ d, const PgBamRec *__restrict__ recs, uint32_t n_recs, PgSrQuery q,
                                  uint32_t *__restrict__ counters, PgSrInfo *__restrict__ info)
{
   const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
   bool delivered = false;
   uint32_t length = 0;
   if (i < n_recs) {
      PgSrInfo mine;
      mine.key = 0; mine.lead = mine.length = 0; mine.bits = 0;
      for (int k = 0; k < 7; k++) mine.pad[k] = 0;
      if (i < counters[0]) {
         const PgBamRec r = recs[i];
         delivered = pgb_itr_delivers(r, pgsr_iterator(q));
         if (delivered) pgsr_analyse(d, r, q, mine);
      }
      length = mine.length;
      info[i] = mine;
   }
   const uint32_t m = __ballot_sync(0xFFFFFFFFu, delivered);
   for (int s = 16; s > 0; s >>= 1) { const uint32_t o = __shfl_xor_sync(0xFFFFFFFFu, length, s); length = o > length ? o : length; }
   if ((threadIdx.x & 31u) == 0) {
      if (m) atomicAdd(&counters[1], (uint32_t)__popc(m));
      if (length) atomicMax(&counters[2], length);
   }
}

// keys / head: the table (table_mask + 1 slots, keys zeroed, heads PGSR_NONE); slot[i]: where record i's name lives
__global__ void pg_sr_insert_kernel(const PgSrInfo *__restrict__ info, uint32_t n_recs, unsigned long long *__restrict__ keys, uint32_t *__restrict__ head,
                                    uint32_t table_mask, uint32_t *__restrict__ slot, uint32_t *__restrict__ next)
{
   const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
   if (i >= n_recs || (info[i].bits & (PGSR_DELIVERED | PGSR_MALFORMED)) != PGSR_DELIVERED) return;
   const unsigned long long key = info[i].key;
   uint32_t s = (uint32_t)(key >> 20) & table_mask;
   for (;;) {
      const unsigned long long prev = atomicCAS(&keys[s], 0ull, key);
      if (prev == 0ull || prev == key) break;
      s = (s + 1u) & table_mask;
   }
   slot[i] = s;
   next[i] = atomicExch(&head[s], i);
}

// partner[i], emit[i] (PGSR_E* / PGSR_R* bits), n_sr[i], n_ref[i] (byte counts for the scans)
__global__ void pg_sr_pair_kernel(const uint8_t *__restrict__ d, const PgBamRec *__restrict__ recs, uint32_t n_recs, PgSrQuery q,
                                  const PgSrInfo *__restrict__ info, const uint32_t *__restrict__ head, const uint32_t *__restrict__ slot,
                                  const uint32_t *__restrict__ next, uint32_t *__restrict__ partner, uint8_t *__restrict__ emit,
                                  uint8_t *__restrict__ n_sr, uint8_t *__restrict__ n_ref, uint32_t *__restrict__ counters)
{
   const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
   if (i >= n_recs) return;
   uint32_t mask = 0, p = PGSR_NONE;
   if (info[i].bits & PGSR_DELIVERED) {
      if (!(info[i].bits & PGSR_MALFORMED)) p = pgsr_partner(d, recs, next, head[slot[i]], i);
      mask = pgsr_emit_mask(info[i], p == PGSR_NONE ? (const PgSrInfo *)0 : &info[p], q);
      if (mask & (PGSR_ERR_INSERT | PGSR_ERR_MALFORMED)) atomicOr(&counters[3], mask & (PGSR_ERR_INSERT | PGSR_ERR_MALFORMED));
   }
   partner[i] = p;
   emit[i] = (uint8_t)mask;
   n_sr[i] = (uint8_t)pgsr_count_sr(mask);
   n_ref[i] = (uint8_t)pgsr_count_ref(mask);
}

__global__ void pg_sr_scatter_kernel(const uint8_t *__restrict__ d, const PgBamRec *__restrict__ recs, uint32_t n_recs, PgSrQuery q,
                                     const PgSrInfo *__restrict__ info, const uint32_t *__restrict__ partner, const uint8_t *__restrict__ emit,
                                     const uint32_t *__restrict__ off_sr, const uint32_t *__restrict__ off_ref, PgSrCand *__restrict__ cands,
                                     uint8_t *__restrict__ seqs, uint32_t stride, PgSrRef *__restrict__ refs)
{
   const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
   if (i >= n_recs || !(emit[i] & 0x3Fu)) return;
   pgsr_scatter(d, recs, info, q, i, partner[i], emit[i], off_sr[i], off_ref[i], cands, seqs, stride, refs);
}
#endif
