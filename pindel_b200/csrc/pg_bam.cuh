// pindel_b200/csrc/pg_bam.cuh -- the BAM record index on the device (between the BGZF
// inflate and the reader's filter).  What htslib does record by record behind sam_itr_next (src/reader.cpp:1258-1346:
// bam_read1's walk over the block_size chain, bam_endpos over the CIGAR for the region test) is done for the whole
// inflated stream at once: the table (offset, refID, pos, end, flag, mapq) of every alignment record.
//
// The chain offset[k+1] = offset[k] + 4 + block_size(offset[k]) is serial.  It is cut into fixed segments of the stream;
// each segment guesses its first record start -- the first offset where three consecutive records look like records
// (sizes consistent with l_read_name / n_cigar_op / l_seq, reference ids in range, the read name NUL-terminated) -- and
// walks from there.  The guesses are then checked against the walks: segment s must land exactly on the guess of the next
// segment that has one.  Segment 0 starts at the true first record, so a consistent set of walks is the true chain; any
// mismatch (a guess fooled by data that looks like a record) is repaired by walking that segment again from the chain's own
// offset (pg_bam_host.h).  Written from the SAM specification section 4.2; __host__ __device__ so that the CPU suite runs the same code (tests/emul/bam_emul.cpp).
#pragma once
#include <stdint.h>

#ifdef __CUDACC__
#define PGB_HD __host__ __device__ __forceinline__
#else
#define PGB_HD inline
#endif

#define PGB_NONE 0xFFFFFFFFFFFFFFFFull

struct PgBamRec {        // pg_bam_rec (include/pindel_pg.h)
   uint64_t off;         // offset of the record's block_size field in the inflated stream
   int32_t tid, pos;     // refID, 0-based leftmost position
   int32_t end;          // pos + reference bases the CIGAR consumes (1 if unmapped or none): bam_endpos
   uint16_t flag;
   uint8_t mapq;
   uint8_t pad;
};

PGB_HD uint32_t pgb_le32(const uint8_t *p) { return (uint32_t)p[0] | (uint32_t)p[1] << 8 | (uint32_t)p[2] << 16 | (uint32_t)p[3] << 24; }
PGB_HD uint32_t pgb_le16(const uint8_t *p) { return (uint32_t)p[0] | (uint32_t)p[1] << 8; }

// does a record header at offset c fit the format?  (necessary conditions only; see the validation above)
PGB_HD bool pgb_plausible(const uint8_t *d, uint64_t n, uint64_t c, int32_t n_ref)
{
   if (c + 36 > n) return false;
   const uint32_t bs = pgb_le32(d + c);
   if (bs < 32u + 2u || bs > (1u << 28) || c + 4 + (uint64_t)bs > n) return false;
   const uint8_t *r = d + c + 4;
   const int32_t tid = (int32_t)pgb_le32(r), pos = (int32_t)pgb_le32(r + 4), mtid = (int32_t)pgb_le32(r + 20), mpos = (int32_t)pgb_le32(r + 24);
   if (tid < -1 || tid >= n_ref || mtid < -1 || mtid >= n_ref || pos < -1 || mpos < -1) return false;
   const uint32_t l_name = r[8], n_cigar = pgb_le16(r + 12), l_seq = pgb_le32(r + 16);
   if (l_name < 2u || l_seq > (1u << 28)) return false;
   const uint64_t fixed = 32ull + l_name + 4ull * n_cigar + (l_seq + 1u) / 2u + l_seq;
   if (fixed > bs) return false;
   if (r[32 + l_name - 1] != 0) return false;                       // read name is NUL-terminated ...
   for (uint32_t k = 0; k + 1 < l_name; k++) {                       // ... and printable ([!-?A-~] in the specification)
      const uint8_t ch = r[32 + k];
      if (ch < 33 || ch > 126) return false;
   }
   return true;
}

// three consecutive plausible records (or fewer when the stream ends exactly on a record boundary)
PGB_HD bool pgb_plausible_chain(const uint8_t *d, uint64_t n, uint64_t c, int32_t n_ref)
{
   for (int k = 0; k < 3; k++) {
      if (!pgb_plausible(d, n, c, n_ref)) return false;
      c += 4 + (uint64_t)pgb_le32(d + c);
      if (c == n) return true;
   }
   return true;
}

// reference bases consumed by the CIGAR: M, D, N, =, X (bam_endpos); rec points at refID, rec_len = block_size.
// A CIGAR that would run past the record (corrupt input) is cut at the record's end.
PGB_HD int32_t pgb_ref_span(const uint8_t *rec, uint64_t rec_len)
{
   const uint32_t l_name = rec[8], flag = pgb_le16(rec + 14);
   uint32_t n_cigar = pgb_le16(rec + 12);
   const uint8_t *cig = rec + 32 + l_name;
   if (32ull + l_name + 4ull * n_cigar > rec_len) n_cigar = 32ull + l_name >= rec_len ? 0u : (uint32_t)((rec_len - 32ull - l_name) / 4ull);
   uint32_t span = 0;                   // (unsigned: the lengths of a damaged CIGAR may sum past 2^31; wraps like the device's add)
   if (!(flag & 4u)) {
      for (uint32_t k = 0; k < n_cigar; k++) {
         const uint32_t c = pgb_le32(cig + 4 * k), op = c & 15u;
         if (op == 0u || op == 2u || op == 3u || op == 7u || op == 8u) span += c >> 4;
      }
   }
   return (int32_t)span > 0 ? (int32_t)span : 1;
}

PGB_HD void pgb_fill(const uint8_t *d, uint64_t c, PgBamRec &r)
{
   const uint8_t *rec = d + c + 4;
   r.off = c;
   r.tid = (int32_t)pgb_le32(rec);
   r.pos = (int32_t)pgb_le32(rec + 4);
   r.end = (int32_t)((uint32_t)r.pos + (uint32_t)pgb_ref_span(rec, (uint64_t)pgb_le32(d + c)));
   r.flag = (uint16_t)pgb_le16(rec + 14);
   r.mapq = rec[9];
   r.pad = 0;
}

// Walk from `start` while the record starts before `limit` (the segment's end).  recs != NULL: store the records.
// Returns the landing offset (first offset >= limit, or n), PGB_NONE on a record that runs past the end of the stream.
PGB_HD uint64_t pgb_walk(const uint8_t *d, uint64_t n, uint64_t start, uint64_t limit, PgBamRec *recs, uint64_t *count)
{
   uint64_t c = start, k = 0;
   while (c < limit && c < n) {
      if (c + 4 > n) return PGB_NONE;
      const uint64_t next = c + 4 + (uint64_t)pgb_le32(d + c);
      if (next > n || next < c + 36) return PGB_NONE;
      if (recs) pgb_fill(d, c, recs[k]);
      k++;
      c = next;
   }
   *count = k;
   return c;
}

// The same walk over a *run* of BGZF blocks cut out of a file (the chunk of a window between two .bai virtual offsets): the
// last record may continue in a block that is not there.  The walk stops in front of the first record that does not fit and
// returns its offset (< limit and < n: the run's tail); PGB_NONE only for a block_size that cannot be a record.
PGB_HD uint64_t pgb_walk_run(const uint8_t *d, uint64_t n, uint64_t start, uint64_t limit, PgBamRec *recs, uint64_t *count)
{
   uint64_t c = start, k = 0;
   while (c < limit && c < n) {
      if (c + 4 > n) break;
      const uint64_t next = c + 4 + (uint64_t)pgb_le32(d + c);
      if (next < c + 36) return PGB_NONE;
      if (next > n) break;
      if (recs) pgb_fill(d, c, recs[k]);
      k++;
      c = next;
   }
   *count = k;
   return c;
}

// BAM header: magic, l_text, text, n_ref, (l_name, name, l_ref) per reference.  -> offset of the first record.
PGB_HD uint64_t pgb_header(const uint8_t *d, uint64_t n, int32_t *n_ref)
{
   if (n < 12 || d[0] != 'B' || d[1] != 'A' || d[2] != 'M' || d[3] != 1) return PGB_NONE;
   uint64_t p = 8ull + pgb_le32(d + 4);
   if (p + 4 > n) return PGB_NONE;
   const int32_t nr = (int32_t)pgb_le32(d + p);
   p += 4;
   if (nr < 0) return PGB_NONE;
   for (int32_t i = 0; i < nr; i++) {
      if (p + 4 > n) return PGB_NONE;
      p += 4ull + pgb_le32(d + p) + 4ull;
      if (p > n) return PGB_NONE;
   }
   *n_ref = nr;
   return p;
}

// ---- discordant-pair discovery on the indexed stream (first half) ----------------------
// ReadInBamReads_RP_Discovery (src/reader.cpp:418-480): every record the region iterator delivers goes through
// build_record_RP_Discovery (src/reader.cpp:982-1097), a stateless test that keeps paired, well-mapped reads whose mate is on
// another reference, too far away, or on the same strand, as one RP_READ row (mates ordered by position when both are on
// one reference).
struct PgRpQuery {       // pg_rp_query
   int32_t tid;          // bam_name2id(header, FragName)
   int32_t beg, end;     // window.getStart(), window.getEnd(): sam_itr_queryi's region
   int32_t insert_size;  // the BAM file's insert size from the configuration file
   uint32_t min_anchor_quality;   // userSettings->minimalAnchorQuality (-A)
};

struct PgRpRow {         // pg_rp (same layout as PgRp in pg_rp.cuh)
   uint32_t pos_a, pos_b;
   uint32_t insert_size;
   int16_t read_length;
   uint8_t da, db;
};

// the iterator ends at the first record that lies past the region (the stream is coordinate-sorted)
PGB_HD bool pgb_itr_stops(const PgBamRec &r, const PgRpQuery &q) { return r.tid != q.tid ? (r.tid > q.tid || r.tid < 0) : r.pos >= q.end; }
PGB_HD bool pgb_itr_delivers(const PgBamRec &r, const PgRpQuery &q) { return r.tid == q.tid && r.pos < q.end && r.end > (q.beg < 0 ? 0 : q.beg); }

// 0 = not kept, 1 = Reads_RP_Discovery (both mates on one reference), 2 = Reads_RP_Discovery_InterChr
PGB_HD int pgb_rp_discovery(const uint8_t *d, const PgBamRec &r, const PgRpQuery &q, PgRpRow &row)
{
   const uint8_t *rec = d + r.off + 4;
   const uint32_t flag = r.flag;
   if (!(flag & 1u)) return 0;                                   // BAM_FPAIRED
   if ((uint32_t)r.mapq < q.min_anchor_quality) return 0;
   if (flag & (4u | 8u)) return 0;                               // BAM_FUNMAP | BAM_FMUNMAP
   const int32_t mtid = (int32_t)pgb_le32(rec + 20), mpos = (int32_t)pgb_le32(rec + 24), isize = (int32_t)pgb_le32(rec + 28);
   const bool rev = (flag & 16u) != 0, mrev = (flag & 32u) != 0;
   const int32_t aisize = isize < 0 ? -isize : isize;
   if (!(r.tid != mtid || aisize > 3 * q.insert_size + 1000 || rev == mrev)) return 0;
   row.da = rev ? '-' : '+';
   row.db = mrev ? '-' : '+';
   row.pos_a = (uint32_t)r.pos;
   row.pos_b = (uint32_t)mpos;
   row.read_length = (int16_t)(int32_t)pgb_le32(rec + 16);       // short ReadLength = l_qseq
   row.insert_size = (uint32_t)q.insert_size;
   if (r.tid != mtid) return 2;
   if (!(row.pos_a < row.pos_b)) {                               // the mate comes first: the row is stored from its side
      const uint32_t p = row.pos_a; row.pos_a = row.pos_b; row.pos_b = p;
      const uint8_t s = row.da; row.da = row.db; row.db = s;
   }
   return 1;
}

#ifdef __CUDACC__
// info: [0] first record offset (PGB_NONE = not a BAM stream), [1] n_ref
__global__ void pg_bam_header_kernel(const uint8_t *__restrict__ d, uint64_t n, uint64_t *__restrict__ info)
{
   if (blockIdx.x == 0 && threadIdx.x == 0) {
      int32_t n_ref = 0;
      info[0] = pgb_header(d, n, &n_ref);
      info[1] = (uint64_t)(uint32_t)n_ref;
   }
}

// starts[s] for s >= 1: the first offset in [first + s * seg, first + (s + 1) * seg) that begins a plausible chain.  One warp
// per segment, 32 candidate offsets per step.
__global__ void __launch_bounds__(256) pg_bam_start_kernel(const uint8_t *__restrict__ d, uint64_t n, uint64_t first, uint64_t seg, uint32_t n_seg,
                                                           int32_t n_ref, uint64_t *__restrict__ starts)
{
   const uint32_t s = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31u;
   if (s >= n_seg) return;
   if (s == 0) {
      if (lane == 0) starts[0] = first;
      return;
   }
   const uint64_t b0 = first + (uint64_t)s * seg, b1 = b0 + seg < n ? b0 + seg : n;
   uint64_t found = PGB_NONE;
   for (uint64_t c0 = b0; c0 < b1; c0 += 32) {
      const uint64_t c = c0 + lane;
      const bool ok = c < b1 && pgb_plausible_chain(d, n, c, n_ref);
      const uint32_t m = __ballot_sync(0xFFFFFFFFu, ok);
      if (m) {
         found = c0 + (uint64_t)(__ffs(m) - 1);
         break;
      }
   }
   if (lane == 0) starts[s] = found;
}

// one thread per segment: walk from starts[s] to the segment's end.  pass 0: counts[s], landing[s]; pass 1: the records,
// at recs[base[s]].
__global__ void pg_bam_walk_kernel(const uint8_t *__restrict__ d, uint64_t n, uint64_t first, uint64_t seg, uint32_t n_seg,
                                   const uint64_t *__restrict__ starts, uint64_t *__restrict__ counts, uint64_t *__restrict__ landing,
                                   const uint64_t *__restrict__ base, PgBamRec *__restrict__ recs)
{
   const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
   if (s >= n_seg) return;
   const uint64_t st = starts[s];
   if (st == PGB_NONE) {
      if (!recs) { counts[s] = 0; landing[s] = PGB_NONE; }
      return;
   }
   const uint64_t b1 = first + (uint64_t)(s + 1) * seg;
   uint64_t k = 0;
   const uint64_t land = pgb_walk(d, n, st, b1 < n ? b1 : n, recs ? recs + base[s] : (PgBamRec *)0, &k);
   if (!recs) { counts[s] = land == PGB_NONE ? 0 : k; landing[s] = land; }
}
// pg_bam_walk_kernel for a run of blocks (pgb_walk_run: a truncated last record ends the chain instead of failing it)
__global__ void pg_bam_walk_run_kernel(const uint8_t *__restrict__ d, uint64_t n, uint64_t first, uint64_t seg, uint32_t n_seg,
                                       const uint64_t *__restrict__ starts, uint64_t *__restrict__ counts, uint64_t *__restrict__ landing,
                                       const uint64_t *__restrict__ base, PgBamRec *__restrict__ recs)
{
   const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
   if (s >= n_seg) return;
   const uint64_t st = starts[s];
   if (st == PGB_NONE) {
      if (!recs) { counts[s] = 0; landing[s] = PGB_NONE; }
      return;
   }
   const uint64_t b1 = first + (uint64_t)(s + 1) * seg;
   uint64_t k = 0;
   const uint64_t land = pgb_walk_run(d, n, st, b1 < n ? b1 : n, recs ? recs + base[s] : (PgBamRec *)0, &k);
   if (!recs) { counts[s] = land == PGB_NONE ? 0 : k; landing[s] = land; }
}
// counters: [0] index of the first record past the region (initialised to n_recs), [1] records delivered
__global__ void pg_bam_stop_kernel(const PgBamRec *__restrict__ recs, uint32_t n_recs, PgRpQuery q, uint32_t *__restrict__ counters)
{
   const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
   const bool stops = i < n_recs && pgb_itr_stops(recs[i], q);
   const uint32_t m = __ballot_sync(0xFFFFFFFFu, stops);
   if (m && (threadIdx.x & 31u) == 0) atomicMin(&counters[0], (i & ~31u) + (uint32_t)__ffs(m) - 1u);
}

// cls_intra[i] / cls_inter[i] = 1 when record i yields a row of that list (byte arrays, scanned into positions afterwards)
__global__ void pg_bam_rpd_flag_kernel(const uint8_t *__restrict__ d, const PgBamRec *__restrict__ recs, uint32_t n_recs, PgRpQuery q,
                                       uint32_t *__restrict__ counters, uint8_t *__restrict__ cls_intra, uint8_t *__restrict__ cls_inter)
{
   const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
   bool delivered = false;
   int cls = 0;
   if (i < n_recs && i < counters[0]) {
      const PgBamRec r = recs[i];
      delivered = pgb_itr_delivers(r, q);
      PgRpRow row;
      if (delivered) cls = pgb_rp_discovery(d, r, q, row);
   }
   if (i < n_recs) {
      cls_intra[i] = cls == 1;
      cls_inter[i] = cls == 2;
   }
   const uint32_t m = __ballot_sync(0xFFFFFFFFu, delivered);
   if (m && (threadIdx.x & 31u) == 0) atomicAdd(&counters[1], (uint32_t)__popc(m));
}

__global__ void pg_bam_rpd_scatter_kernel(const uint8_t *__restrict__ d, const PgBamRec *__restrict__ recs, uint32_t n_recs, PgRpQuery q,
                                          const uint8_t *__restrict__ cls_intra, const uint8_t *__restrict__ cls_inter,
                                          const uint32_t *__restrict__ off_intra, const uint32_t *__restrict__ off_inter,
                                          PgRpRow *__restrict__ rows, uint32_t *__restrict__ src, PgRpRow *__restrict__ inter_rows,
                                          uint32_t *__restrict__ inter_src)
{
   const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
   if (i >= n_recs || !(cls_intra[i] | cls_inter[i])) return;
   PgRpRow row;
   pgb_rp_discovery(d, recs[i], q, row);
   if (cls_intra[i]) { rows[off_intra[i]] = row; src[off_intra[i]] = i; }
   else { inter_rows[off_inter[i]] = row; inter_src[off_inter[i]] = i; }
}
#endif
